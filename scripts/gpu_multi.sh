# N-GPU evidence: 2-GPU parity tests + scaling bench (one process per GPU, NCCL)
set -x
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
for n in $(seq 2 2 $N); do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  tail -c 1500 gpurun_out/bench_n$n.json; tail -3 gpurun_out/bench_n$n.err
done
