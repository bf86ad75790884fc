# the standard single-GPU check on a B200 box: GPU parity tests, smoke(), the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/gputest_1gpu.log; tail -2 gpurun_out/gputest_1gpu.log
timeout 120 python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/bench_1gpu.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])"
