mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_gputest_g.log
cat gpurun_out/r02_gputest_g.log | tail -6
export KEXP_VARIANTS='{"base":[]}'
export KEXP_PDL=0
python tools/kexp.py run > gpurun_out/r02_kexp_g.jsonl 2> gpurun_out/r02_kexp_g.err; cat gpurun_out/r02_kexp_g.jsonl; tail -3 gpurun_out/r02_kexp_g.err
