for t in k2_batch=4 k2_batch=8 k2_batch=12 k2_batch=16 k2_batch=24; do echo $t; WGA_TUNING=$t python tools/prof_decode.py eu-2015-host-shaped 3 2>&1 | tail -1; done
