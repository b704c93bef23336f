P3_TRACE=1 timeout 300 python tools/dbg/e2e_parts.py 2>&1 | tail -14
