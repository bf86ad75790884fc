# multi-GPU run: $1 = number of GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "two_process" 2>&1 | tail -8 > gpurun_out/r02_gputest_mg_n$N.log; cat gpurun_out/r02_gputest_mg_n$N.log | tail -5
for mode in fanout nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --gather $mode > gpurun_out/r02_bench_n${N}_${mode}_k20.json 2> gpurun_out/r02_bench_n${N}_${mode}_k20.err
  tail -4 gpurun_out/r02_bench_n${N}_${mode}_k20.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n${N}_${mode}_k20.json')); print('$mode K=20', d['value'], d['ms_per_step'], d['collective']['count'], d['collective']['ms_each_measured_alone'], d['collective'].get('bus_GBps'), d['per_rank'])"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29813 bench.py --gpus $N --steps 300 --warmup 20 --no-extras --gather $mode > gpurun_out/r02_bench_n${N}_${mode}_k300.json 2> gpurun_out/r02_bench_n${N}_${mode}_k300.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n${N}_${mode}_k300.json')); print('$mode K=300', d['value'], d['ms_per_step'], d['collective']['count'], d['collective']['ms_each_measured_alone'], d['collective'].get('bus_GBps'))"
done
