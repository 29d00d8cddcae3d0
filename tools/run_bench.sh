WGA_TIMING=1 timeout 600 python tools/time_bvcomp.py eu-2015-host-shaped 2>&1 | tail -14
