# multi-GPU run: $1 = number of GPUs, $2 = what (tests|bench|all); every command has a SHORT timeout (a hung collective must not burn the budget)
N=${1:-2}; WHAT=${2:-all}
mkdir -p gpurun_out
export TORCH_NCCL_ASYNC_ERROR_HANDLING=1 NCCL_DEBUG=WARN
if [ "$WHAT" = tests ] || [ "$WHAT" = all ]; then
  for t in "two_process_nccl_gather_equals_single_rank and nccl" "two_process_nccl_gather_equals_single_rank and fanout" "two_process_fused_allreduce_adam"; do
    timeout 170 python -m pytest tests -m gpu -x -q -k "$t" 2>&1 | tail -12 > gpurun_out/r02_gputest_mg_n$N.log; echo "== $t"; tail -6 gpurun_out/r02_gputest_mg_n$N.log
  done
fi
if [ "$WHAT" = diag2 ]; then
  M=278 T=10 timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29877 tools/fanout_diag.py > gpurun_out/r02_fanout_diag_n${N}_t10_m278.json 2> gpurun_out/r02_fanout_diag_n${N}_t10_m278.err
  echo "== diag M=278 rc=$?"; grep -v "^\*\|^$\|OMP_NUM\|W0" gpurun_out/r02_fanout_diag_n${N}_t10_m278.err | tail -14
  for rep in 1 2; do
    timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2988$rep bench.py --gpus $N --steps 20 --warmup 5 --no-extras --gather fanout > gpurun_out/r02_bench_n${N}_fanout_k20_rep$rep.json 2> gpurun_out/r02_bench_n${N}_fanout_k20_rep$rep.err
    python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n${N}_fanout_k20_rep$rep.json')); print('fanout K=20', d['value'], d['ms_per_step'], d['collective']['ms_each_measured_alone'], d['per_rank'])" 2>&1 | tail -1
  done
fi
if [ "$WHAT" = diag ]; then
  for T in 10 64; do
    T=$T timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29877 tools/fanout_diag.py > gpurun_out/r02_fanout_diag_n${N}_t$T.json 2> gpurun_out/r02_fanout_diag_n${N}_t$T.err
    echo "== diag T=$T rc=$?"; grep -v "^\*\|^$\|OMP_NUM\|W0" gpurun_out/r02_fanout_diag_n${N}_t$T.err | tail -14
  done
  timeout 100 python -m pytest tests -m gpu -x -q -k "scan_pipeline or elevation_trajectory" 2>&1 | tail -4
  timeout 100 python bench.py --workload elev --steps 20 --warmup 5 --no-extras > gpurun_out/r02_bench_elev_pipe.json 2> gpurun_out/r02_bench_elev_pipe.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_elev_pipe.json')); print('elev', d['value'], d['ms_per_step'], d['roofline'])" 2>&1 | tail -1
fi
if [ "$WHAT" = bench ] || [ "$WHAT" = all ]; then
  for mode in ${MODES:-fanout ce nccl}; do
    for K in ${KS:-20}; do
      timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800 + K % 100 + 1)) bench.py --gpus $N --steps $K --warmup 5 --no-extras --gather $mode > gpurun_out/r02_bench_n${N}_${mode}_k$K.json 2> gpurun_out/r02_bench_n${N}_${mode}_k$K.err
      echo "== $mode K=$K rc=$?"; grep -v "^$" gpurun_out/r02_bench_n${N}_${mode}_k$K.err | tail -4
      python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n${N}_${mode}_k$K.json')); print('$mode K=$K', d['value'], d['ms_per_step'], d['config']['timing'][-40:], d['collective']['count'], d['collective']['ms_each_measured_alone'], d['collective'].get('bus_GBps'), d['per_rank'])" 2>&1 | tail -1
    done
  done
fi
if [ "$WHAT" = mcast ]; then
  timeout 170 python -m pytest tests -m gpu -x -q -k "two_process_nccl_gather_equals_single_rank" 2>&1 | tail -6
  MODES="mcast fanout nccl" KS="20" bash tools/_run_mg.sh $N bench
fi
