timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_fuzz.py tests/test_gpu_sideinfo.py -x -q 2>&1 | tail -3
timeout 300 python tools/dbg/cli_pattern.py 2>&1 | tail -4
