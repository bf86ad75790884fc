# multi-GPU run: $1 = number of GPUs, $2 = what (tests|bench|all); every command has a SHORT timeout (a hung collective must not burn the budget)
N=${1:-2}; WHAT=${2:-all}
mkdir -p gpurun_out
export TORCH_NCCL_ASYNC_ERROR_HANDLING=1 NCCL_DEBUG=WARN
if [ "$WHAT" = tests ] || [ "$WHAT" = all ]; then
  for t in "two_process_nccl_gather_equals_single_rank and nccl" "two_process_nccl_gather_equals_single_rank and fanout" "two_process_fused_allreduce_adam"; do
    timeout 170 python -m pytest tests -m gpu -x -q -k "$t" 2>&1 | tail -12 > gpurun_out/r02_gputest_mg_n$N.log; echo "== $t"; tail -6 gpurun_out/r02_gputest_mg_n$N.log
  done
fi
if [ "$WHAT" = bench ] || [ "$WHAT" = all ]; then
  for mode in fanout nccl; do
    for K in 20 300; do
      timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 298$K bench.py --gpus $N --steps $K --warmup 5 --no-extras --gather $mode > gpurun_out/r02_bench_n${N}_${mode}_k$K.json 2> gpurun_out/r02_bench_n${N}_${mode}_k$K.err
      echo "== $mode K=$K rc=$?"; grep -v "^$" gpurun_out/r02_bench_n${N}_${mode}_k$K.err | tail -4
      python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n${N}_${mode}_k$K.json')); print('$mode K=$K', d['value'], d['ms_per_step'], d['config']['timing'][-40:], d['collective']['count'], d['collective']['ms_each_measured_alone'], d['collective'].get('bus_GBps'), d['per_rank'])" 2>&1 | tail -1
    done
  done
fi
