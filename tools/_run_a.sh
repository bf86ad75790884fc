mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_gputest_d.log
cat gpurun_out/r02_gputest_d.log | tail -8
timeout 300 ncu --set full --import-source on --clock-control none -k regex:wl_step_quad -s 8 -c 2 -o gpurun_out/r02_v2_4096 python tools/sweep.py --sizes 4096 --steps 6 --warm 3 > gpurun_out/r02_ncu_d.log 2>&1
python bench.py --steps 200 --warmup 20 --no-extras > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; tail -3 gpurun_out/r02_bench_d.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_d.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['variants_ms_per_step'], d['flush_protocol']['step_us_median'], d['warm_l2_graph']['ms_per_step'])"
