# final evidence for one GPU: tests, bench, reference arm, ncu launch list, ncu --set full of the dominant kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 900 python bench.py --steps 20 --warmup 3 --no-pipeline --no-cpu-baseline --train-steps 0 > gpurun_out/bench_n1_nopipe.json 2> gpurun_out/bench_n1_nopipe.err; tail -c 300 gpurun_out/bench_n1_nopipe.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 700 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-pipeline --no-cpu-baseline --train-steps 0 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 2 -o gpurun_out/prof_scan python bench.py --steps 2 --warmup 3 --no-pipeline --no-cpu-baseline --train-steps 0 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ptable_kernel -s 2 -c 1 -o gpurun_out/prof_lstm python bench.py --steps 2 --warmup 3 --no-pipeline --no-cpu-baseline --train-steps 0 > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log | cut -c1-200
timeout 200 python scripts/tsv_bench.py 200000 256 > gpurun_out/tsv_bench.json 2>/dev/null; cut -c1-300 gpurun_out/tsv_bench.json
ls -la gpurun_out | head -40
