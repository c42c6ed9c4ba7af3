nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_encoder.py -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_search.py -q 2>&1 | tail -40
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
