# N-GPU evidence: 2-GPU parity tests + scaling bench (one process per GPU, NCCL)
set -x
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
for n in $(seq 2 2 $N); do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  wc -l gpurun_out/bench_n$n.json; python -c "
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], {k:d[k] for k in ['value','ms_per_step','n_gpus']}, 'e2e', d['e2e']['value'], 'search ms',d['roofline']['ms_per_launch'], 'enc ms',d['roofline']['encoder']['ms'], d['train'] and d['train']['value'])" gpurun_out/bench_n$n.json; tail -3 gpurun_out/bench_n$n.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 3 --no-pipeline --train-steps 0 > gpurun_out/bench_n${n}_nopipe.json 2> gpurun_out/bench_n${n}_nopipe.err
  python -c "
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], {k:d[k] for k in ['value','ms_per_step','n_gpus']}, 'e2e', d['e2e']['value'])" gpurun_out/bench_n${n}_nopipe.json
done
