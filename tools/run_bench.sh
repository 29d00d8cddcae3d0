timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/prof_decode.py eu-2015-host-shaped 3 2>&1 | tail -1
for t in refill=10 unit=64 "refill=10,unit=64"; do echo $t; WGA_TUNING=$t python tools/prof_decode.py eu-2015-host-shaped 3 2>&1 | tail -1; done
