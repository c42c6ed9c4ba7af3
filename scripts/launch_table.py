"""ncu launch list (csv from `ncu --metrics gpu__time_duration.sum --csv`) -> per-kernel count / mean / total table."""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
d = OrderedDict()
for r in rows[1:]:
    try:
        d.setdefault(r[ki][:90], []).append(float(r[vi].replace(",", "")))
    except ValueError:
        pass
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print("%-90s n=%4d mean %9.1f us total %10.1f us  %5.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / tot))
print("TOTAL %.1f us over %d launches" % (tot / 1e3, sum(len(v) for v in d.values())))
