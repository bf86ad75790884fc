mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_gputest_h.log
cat gpurun_out/r02_gputest_h.log | tail -6
export KEXP_VARIANTS='{"base":[]}'
export KEXP_KVARIANTS=8,4
timeout 300 python tools/kexp.py run > gpurun_out/r02_kexp_h.jsonl 2> gpurun_out/r02_kexp_h.err; cat gpurun_out/r02_kexp_h.jsonl; tail -3 gpurun_out/r02_kexp_h.err
