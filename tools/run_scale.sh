# usage (on a multi-GPU box): bash tools/run_scale.sh N
N=${1:-2}
timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q 2>&1 | tail -5
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err; tail -12 gpurun_out/bench_r02_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02_n$N.json'))
print(round(d['value'],1),'Garcs/s', round(d['ms_per_step'],3),'ms', d['scaling'], 'verified', d['verified_bit_exact'])
print('per_rank', d['details']['per_rank']); print('parity', d['details']['n_rank_model_parity_vs_oracle']); print('e2e', d['e2e'])
PY
