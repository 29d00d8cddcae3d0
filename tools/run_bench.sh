timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; tail -3 gpurun_out/bench_r02_final.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-verify --no-cpu-baseline --no-model-leg --random-nodes 0 --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
