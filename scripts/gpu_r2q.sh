# round 2, call Q: G GPUs -- parity tests (G=2), the scaling bench line, an uncapped alternative
G=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "$G" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -5; fi
show() { python - $1 <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'n_gpus', d['n_gpus'], 'q/s %.0f ms/step %.4f e2e %.0f (%.4f ms) | search %.4f | enc %.4f | train %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['encoder']['ms'], d.get('train') and round(d['train']['value'],1)), d['config'].get('pipeline','')[:120])
PY
}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $G --steps 20 --warmup 3 $EXTRA > gpurun_out/bench_n$G.json 2> gpurun_out/bench_n$G.err; show gpurun_out/bench_n$G.json; tail -4 gpurun_out/bench_n$G.err
if [ "$G" = "8" ]; then
  for n in 4; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; show gpurun_out/bench_n$n.json; tail -2 gpurun_out/bench_n$n.err
  done
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 8 --config c5 --steps 10 --warmup 3 --train-steps 0 --no-cpu-baseline > gpurun_out/bench_c5_n8.json 2> gpurun_out/bench_c5_n8.err; show gpurun_out/bench_c5_n8.json; tail -4 gpurun_out/bench_c5_n8.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29550 bench.py --gpus 8 --config c4 --steps 5 --warmup 3 --no-cpu-baseline --no-real-regime > gpurun_out/bench_c4_n8.json 2> gpurun_out/bench_c4_n8.err; show gpurun_out/bench_c4_n8.json; tail -4 gpurun_out/bench_c4_n8.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $G --steps 20 --warmup 3 --search-ctas 0 --cluster-rows 64 --train-steps 0 --no-real-regime --no-cpu-baseline > gpurun_out/bench_n${G}_c0.json 2> gpurun_out/bench_n${G}_c0.err; show gpurun_out/bench_n${G}_c0.json
fi
