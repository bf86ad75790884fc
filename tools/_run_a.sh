mkdir -p gpurun_out
export KEXP_VARIANTS='{"base":[]}'
export KEXP_PDL=1,0
python tools/kexp.py run > gpurun_out/r02_kexp_c.jsonl 2> gpurun_out/r02_kexp_c.err
cat gpurun_out/r02_kexp_c.jsonl; tail -5 gpurun_out/r02_kexp_c.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_gputest_c.log
cat gpurun_out/r02_gputest_c.log
python bench.py --steps 200 --warmup 20 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; tail -3 gpurun_out/r02_bench_c.err; cat gpurun_out/r02_bench_c.json
