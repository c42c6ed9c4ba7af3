"""Probe (not a test, not a benchmark): the tcgen05 scan under different work decompositions (environment knobs of
search_tc.cu), one subprocess per configuration because the knobs are read once per process.
usage: python scripts/scan_configs.py [QxN ...]"""
import os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
shapes = sys.argv[1:] or ["600x1000000", "4800x125000"]
CONFIGS = [
    ("auto", {}),
    ("nospan", {"SSE_SCAN_SPAN": "0"}),
    ("pdl", {"SSE_SCAN_PDL": "1"}),
    ("pack0", {"SSE_SCAN_PACK": "0"}),
    ("fused", {"SSE_SCAN_FUSED": "1"}),
    ("fused pack0", {"SSE_SCAN_FUSED": "1", "SSE_SCAN_PACK": "0"}),
    ("tn64", {"SSE_SCAN_ACC1": "0"}),
    ("cluster", {"SSE_SCAN_CLUSTER": "1"}),
    ("nofollow", {"SSE_SCAN_FOLLOW": "0"}),
    ("follow_cost650", {"SSE_SCAN_COST": "650,1000"}),
    ("follow_cost500", {"SSE_SCAN_COST": "500,1000"}),
    ("div8", {"SSE_SCAN_SAMPLE_DIV": "8"}),
    ("div12", {"SSE_SCAN_SAMPLE_DIV": "12"}),
    ("div24", {"SSE_SCAN_SAMPLE_DIV": "24"}),
    ("cost100", {"SSE_SCAN_COST": "100,1100"}),
    ("cost500", {"SSE_SCAN_COST": "500,1000"}),
    ("cost780", {"SSE_SCAN_COST": "780,1000"}),
    ("cost1000", {"SSE_SCAN_COST": "1000,1000"}),
    ("cost1300", {"SSE_SCAN_COST": "1300,1000"}),
    ("cost780pdl", {"SSE_SCAN_COST": "780,1000", "SSE_SCAN_PDL": "1"}),
]
only = os.environ.get("SCAN_CONFIGS")
for name, env in CONFIGS:
    if only and name not in only.split(","):
        continue
    e = dict(os.environ); e.update(env)
    print("#### %s %s" % (name, env), flush=True)
    try:
        r = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "search_probe.py")] + shapes, env=e, capture_output=True, text=True, timeout=180)
        out = [l for l in r.stdout.splitlines() if "Q=" in l]
        print("\n".join(out) if out else "FAILED rc=%d: %s" % (r.returncode, (r.stderr or "")[-600:]), flush=True)
    except subprocess.TimeoutExpired:
        print("TIMEOUT", flush=True)
