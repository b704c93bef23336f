timeout 900 python -m pytest tests/test_gpu_sideinfo.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_sideinfo.py 2>&1 | tail -3
P3_TRACE=1 timeout 300 python tools/dbg/e2e_parts.py 2>&1 | tail -6
