# multi-GPU validation: usage  bash scripts/gpu_r2_multi.sh G   (run under gpurun --gpus G)
G=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "$G" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -5; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $G --steps 20 --warmup 3 > gpurun_out/bench_n$G.json 2> gpurun_out/bench_n$G.err; head -c 600 gpurun_out/bench_n$G.json; tail -4 gpurun_out/bench_n$G.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $G --steps 20 --warmup 3 --search-ctas 108 --train-steps 0 --no-real-regime > gpurun_out/bench_n${G}_c108.json 2> gpurun_out/bench_n${G}_c108.err; head -c 300 gpurun_out/bench_n${G}_c108.json
