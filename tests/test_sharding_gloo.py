"""N > 1 host path on CPU: two gloo ranks shard segments with LPT, score their shard with a stub
scorer and all_gather the per-segment records; both ranks must end up with the full table."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200sep import sharding, synth
    lens = synth.segment_lengths(37, 9)
    costs = [sharding.frames_of(n) for n in lens]
    mine = sharding.lpt_partition(costs, world)[rank]
    # stub "metrics": deterministic function of the segment id (stands in for sisdr, sisdri, perm...)
    rec = np.array([[i * 1.5, -float(lens[i]), float(i % 3)] for i in mine], np.float64).reshape(len(mine), 3)
    table = sharding.gather_records(mine, rec, len(lens))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), table)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from b200sep import synth
    lens = synth.segment_lengths(37, 9)
    want = np.array([[i * 1.5, -float(lens[i]), float(i % 3)] for i in range(37)])
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npy"))
        assert got.shape == (37, 3) and np.array_equal(got, want)


def test_gather_without_process_group_is_identity():
    sys.path.insert(0, ROOT)
    from b200sep import sharding
    t = sharding.gather_records([2, 0], np.array([[1.0, 2.0], [3.0, 4.0]]), 4)
    assert np.array_equal(t[2], [1.0, 2.0]) and np.array_equal(t[0], [3.0, 4.0]) and np.isnan(t[1]).all()
