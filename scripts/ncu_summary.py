"""Turn `ncu --set full` reports (gpurun_out/*.ncu-rep) into the text summary kept under profiles/ and refresh
profiles/traffic.json (dram bytes per launch of the filter scan).  Run here (CPU): python scripts/ncu_summary.py"""
import csv, io, json, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
           "smsp__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_tensor.sum", "lts__t_bytes.sum"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units, data = r[0], r[1], r[2:]
    return hdr, units, data


def main():
    reps = sys.argv[1:] or [os.path.join(REPO, "gpurun_out", f) for f in ("prof_scan_600.ncu-rep", "prof_lstm.ncu-rep", "prof_small.ncu-rep", "prof_scan_4800.ncu-rep", "prof_fin_4800.ncu-rep")]
    lines = ["# ncu --set full --clock-control none captures (B200, round 2); per kernel launch, values with units"]
    traffic = None
    for rep in reps:
        if not os.path.exists(rep):
            continue
        hdr, units, data = rows(rep)
        ki = hdr.index("Kernel Name")
        for d in data:
            lines.append("== %s : %s" % (os.path.basename(rep), d[ki][:110]))
            vals = {}
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    lines.append("   %-78s %s %s" % (m, d[i], units[i]))
                    vals[m] = (d[i], units[i])
            if "scan_kernel<1" in d[ki] and "dram__bytes_read.sum" in vals and "prof_scan_600" in rep:
                tot = sum(float(vals[m][0].replace(",", "")) * UNIT.get(vals[m][1], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                traffic = int(tot)
    open(os.path.join(REPO, "profiles", "r02_ncu_full_summary.txt"), "w").write("\n".join(lines) + "\n")
    if traffic:
        json.dump({"targets_per_gpu": 1000000, "queries": 600, "scan_filter_dram_bytes": traffic,
                   "source": "profiles/r02_ncu_full_summary.txt: ncu --set full of scan_kernel<FILTER>, dram__bytes_read.sum + dram__bytes_write.sum per launch"},
                  open(os.path.join(REPO, "profiles", "traffic.json"), "w"))
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
