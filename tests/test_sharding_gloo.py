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


def _fake_separate(chunks, sr):
    return [[np.asarray(c, np.float32) * 0.5, np.roll(np.asarray(c, np.float32), 5) * 0.25,
             (np.asarray(c, np.float32) ** 3).astype(np.float32)] for c in chunks]


def _oracle_pit(mixes, refs, preds, variant="overlap3"):
    from oracle import pit_restated as pr
    from b200sep.pit import PitResult
    n, K = len(mixes), len(refs[0])
    sisdr, sisdri, perm = np.full(n, np.nan), np.full(n, np.nan), np.full((n, K), -1, np.int32)
    for i in range(n):
        b, d, idx = pr.compute_sdr_improvement_pit_k_o3(mixes[i], list(refs[i]), list(preds[i]))
        if idx:
            sisdr[i], sisdri[i], perm[i] = b, d, idx
    return PitResult(sisdr, sisdri, perm, np.zeros(n, bool), np.zeros((n, 3, 4)))


def _eval_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200sep import sharding, synth
    mixes, srcs = synth.make_batch(23, seed=5, n_src=3, lo_s=0.2, hi_s=1.0, with_sources=True)
    sisdr, sisdri, perm, _ = sharding.evaluate_sharded(mixes, srcs, 8000, _fake_separate, _oracle_pit, max_samples=20000)
    np.savez(os.path.join(out_dir, f"eval{rank}.npz"), sisdr=sisdr, sisdri=sisdri, perm=perm)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_evaluation_equals_single_process(tmp_path):
    """evaluate_sharded over 2 gloo ranks (LPT shards, length buckets, one all_gather) == the same loop in one process."""
    world = 2
    mp.spawn(_eval_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from b200sep import sharding, synth
    mixes, srcs = synth.make_batch(23, seed=5, n_src=3, lo_s=0.2, hi_s=1.0, with_sources=True)
    s1, si1, p1, _ = sharding.evaluate_sharded(mixes, srcs, 8000, _fake_separate, _oracle_pit, max_samples=1 << 30)
    assert not np.isnan(s1).any() and p1.shape == (23, 3)
    for r in range(world):
        g = np.load(os.path.join(str(tmp_path), f"eval{r}.npz"))
        assert np.array_equal(g["sisdr"], s1) and np.array_equal(g["sisdri"], si1) and np.array_equal(g["perm"], p1)


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
