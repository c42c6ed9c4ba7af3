# round 2, call R: the other BASELINE configs on one GPU (c5 with one rank's 625k-row shard of the 5M index)
mkdir -p gpurun_out
show() { python - $1 <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d['roofline']
    print(sys.argv[1], 'q/s %.0f ms/step %.4f e2e %.0f | %s frac %.3f ms %.4f | enc %.4f ms %.1f TF/s | train %s | cpu %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['bound'], r['frac'], r['ms_per_launch'], r['encoder']['ms'], r['encoder']['achieved_tflops'], d.get('train') and round(d['train']['value'],1), d.get('cpu_baseline') and round(d['cpu_baseline']['value'],1)))
    print('   ', d['config']['workload'][:200])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
for c in c2 c3 c1; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; show gpurun_out/bench_$c.json; tail -2 gpurun_out/bench_$c.err
done
timeout 900 python bench.py --config c4 --steps 10 --warmup 3 --train-steps 10 --no-real-regime > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; show gpurun_out/bench_c4.json; tail -2 gpurun_out/bench_c4.err
timeout 900 python bench.py --config c5 --targets 625000 --steps 10 --warmup 3 --train-steps 0 > gpurun_out/bench_c5_shard.json 2> gpurun_out/bench_c5_shard.err; show gpurun_out/bench_c5_shard.json; tail -2 gpurun_out/bench_c5_shard.err
