cd $GRAFT_REPO_ROOT
nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8_v8.json 2> gpurun_out/bench_n8_v8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n8_v8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','e2e','clocks')})
PY
