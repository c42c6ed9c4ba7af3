# round 2, call AD: 2 GPUs -- parity test and the bench line with the two-stream / graph train step (DP all-reduce)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -2 gpurun_out/bench_n2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print('n2 q/s %.0f ms/step %.4f e2e %.0f | train %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['train']))
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --config c4 --steps 5 --warmup 3 --train-steps 30 --no-cpu-baseline --no-real-regime > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err; tail -2 gpurun_out/bench_c4_n2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_n2.json')); print('c4 n2 %.1f step/s %.3f ms' % (d['value'], d['ms_per_step']), d['config'])"
