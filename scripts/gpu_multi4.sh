mkdir -p gpurun_out
N=${1:-4}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
wc -l gpurun_out/bench_n$N.json; python -c "
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], {k:d[k] for k in ['value','ms_per_step','n_gpus']}, 'e2e', d['e2e']['value'], 'search ms',d['roofline']['ms_per_launch'], 'enc ms',d['roofline']['encoder']['ms'], d['train'] and d['train']['value'])" gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
