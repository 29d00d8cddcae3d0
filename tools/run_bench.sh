# usage (on the GPU box, from the repo root): bash tools/run_bench.sh [workloads...]
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for wl in ${@:-web-1m eu-2015-host-shaped}; do
timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$wl.json')); print('$wl', round(d['value'],1),'Garcs/s', round(d['ms_per_step'],3),'ms', {k:round(v,3) for k,v in d['roofline']['stage_ms'].items()}, d['verified_bit_exact'], 'frac',round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],2), 'ra', d.get('random_access'))" || tail -5 gpurun_out/bench_$wl.err
done
