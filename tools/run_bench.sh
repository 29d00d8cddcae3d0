timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/prof_decode.py web-1m 3 2>&1 | tail -1
python tools/prof_decode.py eu-2015-host-shaped 3 2>&1 | tail -1
