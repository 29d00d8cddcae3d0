timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 1500 python bench.py --workload twitter-2010-shaped --random-nodes 10000000 --steps 5 > gpurun_out/bench_r02_twitter.json 2> gpurun_out/bench_r02_twitter.err; tail -4 gpurun_out/bench_r02_twitter.err
timeout 600 python bench.py --workload dblp-2011-shaped --steps 20 > gpurun_out/bench_r02_dblp.json 2> gpurun_out/bench_r02_dblp.err; tail -3 gpurun_out/bench_r02_dblp.err
