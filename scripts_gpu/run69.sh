cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/sweep_sharded.py 600 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -4
timeout 600 python tests/sweep_sharded.py 600 2>&1 | tail -2
