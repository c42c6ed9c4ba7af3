mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_entrypoints.py -x -q 2>&1 | tail -3
timeout 300 python scripts/search_probe.py 600x1000000 4800x125000 2>&1 | grep "Q="
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/probe_launches.csv python scripts/search_probe.py 600x1000000 > /dev/null 2>&1; python - <<PY
import csv
from collections import defaultdict
rows=[r for r in csv.reader(open("gpurun_out/probe_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
d=defaultdict(list)
for r in rows[1:]:
    if 'sse' in r[ki]: d[r[ki][:60]].append(float(r[vi].replace(',','')))
for k,v in d.items(): print(k, len(v), round(sum(v)/len(v)/1e3,1), 'us')
PY
