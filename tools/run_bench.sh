timeout 1500 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -5
python tools/prof_decode.py eu-2015-host-shaped 3 2>&1 | tail -1
