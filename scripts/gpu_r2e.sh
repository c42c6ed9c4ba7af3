# round 2, call E: ncu --set full of finalize / scan at the two shapes; quick re-checks
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_misc.py -q -m gpu 2>&1 | tail -4
for shape in 600x1000000 4800x125000; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalize_kernel -s 3 -c 1 -o gpurun_out/prof_fin_$shape python scripts/search_probe.py $shape > gpurun_out/ncu_fin_$shape.log 2>&1; tail -1 gpurun_out/ncu_fin_$shape.log | cut -c1-150
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 7 -c 1 -o gpurun_out/prof_scan_4800 python scripts/search_probe.py 4800x125000 > gpurun_out/ncu_scan_4800.log 2>&1; tail -1 gpurun_out/ncu_scan_4800.log | cut -c1-150
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 7 -c 1 -o gpurun_out/prof_scan_600 python scripts/search_probe.py 600x1000000 > gpurun_out/ncu_scan_600.log 2>&1; tail -1 gpurun_out/ncu_scan_600.log | cut -c1-150
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_probe.csv python scripts/search_probe.py 600x1000000 > /dev/null 2>&1; python scripts/launch_table.py gpurun_out/launches_probe.csv | head -12
