mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_gputest_l.log; tail -4 gpurun_out/r02_gputest_l.log
for wk in elev hound4wd drift; do
  timeout 400 python bench.py --workload $wk --steps 100 --warmup 10 --no-extras > gpurun_out/r02_bench_${wk}_l.json 2> gpurun_out/r02_bench_${wk}_l.err; tail -2 gpurun_out/r02_bench_${wk}_l.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_${wk}_l.json')); print('$wk', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['flush_protocol']['step_us_median'], d['warm_l2_graph']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'])"
done
timeout 200 ncu --set full --import-source on --clock-control none -k regex:wl_step_duo -s 8 -c 1 -o gpurun_out/r02_ncu_drift_4096 python tools/sweep.py --sizes 4096 --steps 6 --warm 3 > gpurun_out/r02_ncu_l1.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"wl_step_quad|wl_scan" -s 16 -c 2 -o gpurun_out/r02_ncu_elev_4096 python tools/sweep.py --task elevation --sizes 4096 --steps 6 --warm 3 > gpurun_out/r02_ncu_l2.log 2>&1
python tools/next_rows_bench.py > gpurun_out/r02_next_rows.json 2> gpurun_out/r02_next_rows.err; cat gpurun_out/r02_next_rows.json | head -c 1500
