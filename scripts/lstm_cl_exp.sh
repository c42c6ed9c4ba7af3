for d in 16 32 64 128; do echo "== sample div $d"; SSE_SCAN_SAMPLE_DIV=$d timeout 300 python scripts/search_probe.py 600x1000000 4800x125000 1200x500000 2>&1 | grep "Q="; done
