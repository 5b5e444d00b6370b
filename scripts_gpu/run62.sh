cd $GRAFT_REPO_ROOT
nproc; python -c "import os; print(os.cpu_count(), len(os.sched_getaffinity(0)))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_v7.json 2> gpurun_out/bench_n2_v7.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_v7.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','e2e')})
PY
OMP_NUM_THREADS=8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('OMP8', d['value'], d['e2e']['value'])"
