"""BASELINE.json configs[3] / [4] as a multi-GPU sweep (run under torchrun on a GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/sweep_sharded.py [n_segments]

n synthetic 3-speaker overlap segments U[0.4, 4.0] s (+ the chunks of one 60 s mixture cut as the streaming path does) are
LPT-sharded over the ranks, separated and PIT-scored on each rank's GPU, and the per-segment records are all_gathered
over NCCL.  Rank 0 checks a sample of segments against the oracle (model + the restated reference PIT) and prints one
JSON line with the aggregate metrics (`metrics.sep_metrics` keys) and the sweep throughput."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200sep
from b200sep import metrics, sharding, synth, weights

n_seg = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 600
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mixes, srcs = synth.make_batch(n_seg, seed=4, n_src=3, lo_s=0.4, hi_s=4.0, with_sources=True)
long_mix, long_src = synth.make_mixture(9001, 60 * 8000, 3)
for p in (2.0, 5.0):
    n = synth.streaming_chunk_samples(p)
    for o in range(0, len(long_mix) - n + 1, n):
        mixes.append(long_mix[o:o + n]); srcs.append([s[o:o + n] for s in long_src])
sd, args = weights.random_state_dict(3, 0)
sep = b200sep.Separator(backend="b200", device=f"cuda:{local}", sample_rate=8000, n_src=3, state_dict=sd, model_args=args)
pit = lambda m, r, p, v: b200sep.pit_sisdr_batch(m, r, p, v, device=local)
sharding.evaluate_sharded(mixes[:8], srcs[:8], 8000, sep.separate_batch, pit)  # warm-up (also of NCCL)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.time()
fused = sep.engine.separate_and_score if "--two-step" not in sys.argv else None
sisdr, sisdri, perm, t_sep = sharding.evaluate_sharded(mixes, srcs, 8000, sep.separate_batch, pit, max_samples=1 << 20,
                                                       separate_and_score=fused)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.time() - t0
if rank == 0:
    assert not np.isnan(sisdr).any() and perm.shape == (len(mixes), 3)
    from oracle import asteroid_restated as ar, pit_restated as pr
    ref_model = ar.ConvTasNet(**args).eval(); ref_model.load_state_dict(sd)
    worst, same = 0.0, True
    for i in list(range(0, len(mixes), max(1, len(mixes) // 12)))[:12]:
        with torch.no_grad():
            est = ref_model(torch.from_numpy(mixes[i])[None])[0].numpy()
        best, sdri, idx = pr.compute_sdr_improvement_pit_k_o3(mixes[i], srcs[i], list(est))
        worst = max(worst, abs(sdri - sisdri[i])); same &= (list(perm[i]) == idx)
    audio = sum(len(m) for m in mixes) / 8000.0
    line = {"segments": len(mixes), "world": world, "device_resident_pit": fused is not None, "audio_s": round(audio, 1), "sweep_s": round(dt, 3),
            "audio_s_per_s_incl_host_pit_and_gather": round(audio / dt, 1), "rank0_time_sep_sec": round(t_sep, 3),
            "max_abs_dSISDRi_vs_oracle_on_sample_dB": round(worst, 4), "perm_identical_on_sample": bool(same)}
    line.update(metrics.sep_metrics(list(sisdr), list(sisdri)))
    print(json.dumps(line), flush=True)
    assert worst <= 0.05 and same
if world > 1:
    dist.destroy_process_group()
