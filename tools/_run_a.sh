mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "drift or variant or graph or trajectory or rollout" 2>&1 | tail -3
KEXP_SETS=8 KEXP_VARIANTS='{"base":[]}' timeout 300 python tools/kexp.py run --envs 4096 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['variant'], 'cold', round(d['cold_us_median'],2), 'warm', round(d['warm_graph_us'],3))"
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_drift_k20.json 2> gpurun_out/r02_bench_drift_k20.err; tail -2 gpurun_out/r02_bench_drift_k20.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_drift_k20.json')); print('drift', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['flush_protocol']['step_us_median'], d['warm_l2_graph']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'])"
timeout 300 python bench.py --steps 100 --warmup 10 --no-extras > gpurun_out/r02_bench_drift_k100.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_drift_k100.json')); print('drift K=100', d['value'], d['ms_per_step'])"
timeout 300 python bench.py --workload hound4wd --steps 20 --warmup 5 --no-extras > gpurun_out/r02_bench_hound4wd_k20.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_hound4wd_k20.json')); print('hound K=20', d['value'], d['ms_per_step'])"
