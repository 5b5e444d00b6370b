#!/usr/bin/env python
"""bench.py -- separated audio-seconds per second of the overlap-segment separation stage.

    python bench.py --gpus N --steps K --warmup W            # the B200 backend (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)
    python bench.py --config {1,2,3,4} ...                   # BASELINE.json configs[1..4] (default 1)

config 1 (the headline, BASELINE.json configs[1]): a "step" is one pass of the hot path (Conv-TasNet n_src=3 forward,
8 kHz) over one batch of 64 synthetic overlap segments U[0.2, 4.0] s (~135 audio-seconds), seeded random-init
weights.  With N > 1 the ranks share ONE seeded pool of 64 N segments that `sharding.lpt_partition` deals out (weak
scaling: 64 segments per GPU, no data-path collective); `value` = audio-seconds of all ranks / max-over-ranks device
time of the K timed steps, which follow >= 2 s of the same load (the GPU is at its sustained clocks when they run).
Every default run also carries a `sweep` object: BASELINE.json configs[4] (10 k segments + 60 s streaming-chunked
mixtures) through the real multi-GPU path -- LPT shards, length buckets, separation + PIT on the device, ONE
all_gather of the per-segment records -- timed end to end.

config 2 / 3: separation + PIT SI-SDR(i) of the 2-source Libri2Mix-shaped set / of 3000 3-source segments;
config 4: the sweep as the metric line.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 8000
METRIC = "separated audio-sec/sec (8k, 3-src)"
UNIT = "audio-s/s"
FLOP_PER_FRAME = {3: 9_969_664, 2: 9_822_208}  # algorithmic, SURVEY.md 8(d)
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
GATE_DB = 40.0          # north_star: separated outputs >= 40 dB SI-SDR vs the reference's fp32 outputs
GATE_SDRI_DB = 0.05     # north_star: aggregate SI-SDRi within 0.05 dB
SUSTAIN_S = 2.0         # seconds of identical load in front of the timed steps

WORKLOADS = {
    1: "configs[1]: 64 synthetic 3-speaker overlap segments U[0.2,4.0] s per GPU, 8 kHz, ConvTasNet n_src=3 random-init",
    2: "configs[2]: ConvTasNet n_src=2, Libri2Mix-8k-shaped synthetic mixtures (64 utterances U[3,15] s + 128 OSD-like "
       "sub-segments >= 0.4 s), separation + PIT SI-SDRi (evaluate_with_sources variant)",
    3: "configs[3]: 3000 synthetic 3-speaker segments U[0.4,4.0] s, separation + batched PIT SI-SDR/SI-SDRi over 3! permutations",
    4: "configs[4]: 10000 synthetic overlap segments U[0.2,4.0] s + 32 60-s mixtures cut as the streaming path does "
       "(15360- / 39936-sample chunks, each chunk + one OSD sub-segment), LPT-sharded over the GPUs, separation + PIT on the "
       "device, one all_gather of the per-segment records",
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            with open(p) as f:
                d = json.load(f)
            return d, "measured"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled from a thread of this process every
    ~2 ms (nvidia_ml_py), with `nvidia-smi -lms` as the fallback when NVML cannot be imported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.handle = None
        self.samples = []      # (sm_mhz, reasons bitmask)
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may remap indices: go through the UUID of the CUDA device when torch exposes it
            self.handle = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            except Exception:
                self.handle = None
            if self.handle is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                phys = int(ids[gpu_index]) if len(ids) > gpu_index and ids[gpu_index].isdigit() else gpu_index
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(
                    n, "nvmlDeviceGetCurrentClocksEventReasons") else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            if self._thread is not None:
                self._thread.join(timeout=1.0)
            n = self.nvml
            names = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            sm = [s for s, _ in self.samples]
            reasons = sorted(k for k, bit in names.items() if any(r & bit for _, r in self.samples))
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------ workloads
class Pool:
    """A seeded global list of segments.  `lens` is known to every rank (cheap); waveforms are generated only for the
    indices a rank owns (`materialise`), from per-segment seeds, so every rank sees the same pool."""

    def __init__(self, n_src, items):
        self.n_src = n_src
        self.items = items           # (mixture seed, total samples of the mixture, start, length)
        self.lens = [it[3] for it in items]
        self._cache = {}

    def materialise(self, idx):
        from b200sep import synth
        mixes, refs = [], []
        for i in idx:
            seed, n_total, start, n = self.items[i]
            key = (seed, n_total)
            if key not in self._cache:
                if len(self._cache) > 8:
                    self._cache.clear()
                self._cache[key] = synth.make_mixture(seed, n_total, self.n_src, SR)
            m, s = self._cache[key]
            mixes.append(np_c(m[start:start + n]))
            refs.append([np_c(a[start:start + n]) for a in s])
        return mixes, refs


def np_c(a):
    import numpy as np
    return np.ascontiguousarray(a, dtype=np.float32)


def pool_config1(world):
    from b200sep import synth
    n = 64 * world
    lens = synth.segment_lengths(n, 2, 0.2, 4.0, SR)
    return Pool(3, [(2 * 100003 + i, ln, 0, ln) for i, ln in enumerate(lens)])  # world 1: synth.make_batch(64, seed=2)


def pool_config2():
    import numpy as np
    rng = np.random.default_rng(32)
    items = []
    for u in range(64):
        n_total = int(round(rng.uniform(3.0, 15.0) * SR))
        items.append((3_000_000 + u, n_total, 0, n_total))     # the whole utterance (offline_overlap_mvp.py path)
        for _ in range(2):                                     # OSD-like overlap sub-segments >= 0.4 s
            n = int(round(rng.uniform(0.4, min(4.0, n_total / SR)) * SR))
            s = int(rng.integers(0, n_total - n + 1))
            items.append((3_000_000 + u, n_total, s, n))
    return Pool(2, items)


def pool_config3():
    from b200sep import synth
    lens = synth.segment_lengths(3000, 4, 0.4, 4.0, SR)
    return Pool(3, [(4 * 100003 + i, ln, 0, ln) for i, ln in enumerate(lens)])


def pool_config4(n_seg=10000, n_long=32):
    import numpy as np
    from b200sep import synth
    lens = synth.segment_lengths(n_seg, 5, 0.2, 4.0, SR)
    items = [(5 * 100003 + i, ln, 0, ln) for i, ln in enumerate(lens)]
    rng = np.random.default_rng(55)
    n60 = 60 * SR
    for u in range(n_long):
        chunk = synth.streaming_chunk_samples(2.0 if u % 2 == 0 else 5.0, SR)  # streaming_overlap_3src.py:89-91
        for c0 in range(0, n60 - chunk + 1, chunk):
            items.append((6_000_000 + u, n60, c0, chunk))                      # full-chunk separation (core.py:186)
            n = int(rng.uniform(0.2, 0.8) * chunk)                             # one OSD overlap sub-segment (core.py:247)
            s = int(rng.integers(0, chunk - n + 1))
            items.append((6_000_000 + u, n60, c0 + s, max(n, 1600)))
    return Pool(3, items)


def cpu_sample_indices(lens, k):
    """Stratified-by-length subsample: every (n/k)-th segment of the length-sorted batch."""
    order = sorted(range(len(lens)), key=lambda i: (lens[i], i))
    step = max(1, len(order) // k)
    return [order[i] for i in range(step // 2, len(order), step)][:k]


def oracle_model(n_src, device="cpu"):
    import torch
    from b200sep import weights
    from oracle import asteroid_restated as ar
    sd, margs = weights.random_state_dict(n_src, 0)
    model = ar.ConvTasNet(**margs).eval()
    model.load_state_dict(sd)
    return model.to(device)


def time_cpu_port(mixes, idx, n_src=3, threads=None):
    """The reference's CPU path: per-segment loop, time.time() around every separate() call
    (/root/reference/scripts/osd/overlap3_core.py:689-691), torch CPU fp32, all host threads."""
    import torch
    torch.set_num_threads(threads or os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: undo it explicitly
    model = oracle_model(n_src)
    t_sep, audio = 0.0, 0.0
    with torch.no_grad():
        model(torch.from_numpy(mixes[idx[0]][:1600])[None])  # warm the allocator / thread pool
        for i in idx:
            t0 = time.time()
            x = torch.from_numpy(mixes[i]).to(dtype=torch.float32).unsqueeze(0)
            est = model(x)
            out = est.squeeze(0).cpu().numpy()
            _ = [out[s] for s in range(n_src)]
            t_sep += time.time() - t0
            audio += len(mixes[i]) / SR
    return audio / t_sep, audio, t_sep, torch.get_num_threads()


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port: asteroid is not installable
    here) on all host cores; every step = one per-segment pass over ALL 64 segments of configs[1]
    -- the same sample the GPU arm's `cpu_baseline` times."""
    if rank != 0:
        return
    pool = pool_config1(1)
    idx = list(range(len(pool.lens)))
    mixes, _ = pool.materialise(idx)
    for _ in range(1 if args.warmup else 0):
        time_cpu_port(mixes, idx[:8])
    t_tot, a_tot, nthr = 0.0, 0.0, 1
    for _ in range(args.steps):
        _, audio, t, nthr = time_cpu_port(mixes, idx)
        t_tot += t; a_tot += audio
    value = a_tot / t_tot
    sample = (f"all {len(idx)} segments of the configs[1] batch per step, one separate() call per segment "
              f"({a_tot / args.steps:.1f} audio-s per step)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[1], "segments_per_gpu": len(idx), "sample_rate": SR, "n_src": 3,
                   "precision": "fp32 (torch CPU)", "torch_threads": nthr, "host_cpus": os.cpu_count(),
                   "parallelism": "rank 0 only: the reference is a single-process CPU loop"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthr, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


KERNEL_CLASSES = ["encoder", "bottleneck_gemm", "in_conv_gemm", "dwconv", "res_skip_gemm", "apply_in_gemm", "dw_rs_gemm",
                  "final_skip", "mask_gemm", "decoder", "mask_dec"]


def launches_per_forward(n_layers, pipeline):
    """pipeline 1 (fused fp16 tensor-core mode): encoder, bottleneck, (K_B, K_A) per block, final skip, mask,
    decoder; pipeline 0: encoder, bottleneck, (in-conv, depthwise, res/skip) per block, mask, decoder."""
    if pipeline == 2:  # ... and mask GEMM + decoder fused into one kernel
        return 2 + 2 * n_layers + 1 + 1
    return 2 + 2 * n_layers + 1 + 2 if pipeline == 1 else 2 + 3 * n_layers + 2


def classify_launches(ms, n_layers, pipeline=0):
    per = {k: [] for k in KERNEL_CLASSES}
    if len(ms) != launches_per_forward(n_layers, pipeline):
        return per
    per["encoder"].append(ms[0]); per["bottleneck_gemm"].append(ms[1])
    if pipeline in (1, 2):
        for i in range(n_layers):
            per["apply_in_gemm"].append(ms[2 + 2 * i]); per["dw_rs_gemm"].append(ms[3 + 2 * i])
        if pipeline == 2:
            per["final_skip"].append(ms[-2]); per["mask_dec"].append(ms[-1])
            return per
        per["final_skip"].append(ms[-3])
    else:
        for i in range(n_layers):
            per["in_conv_gemm"].append(ms[2 + 3 * i]); per["dwconv"].append(ms[3 + 3 * i]); per["res_skip_gemm"].append(ms[4 + 3 * i])
    per["mask_gemm"].append(ms[-2]); per["decoder"].append(ms[-1])
    return per


def si_sdr_db(ref, est):
    from oracle import pit_restated as pr
    return pr.si_sdr_f64(ref, est)


def parity_check(model_cpu, mixes, outs, k=8):
    """>= 40 dB SI-SDR of `outs` (what the TIMED path produced) against the oracle's fp32 CPU forward, on k segments
    of the timed batch stratified by length.  Returns the worst value over segments and sources."""
    import torch
    lens = [int(m.shape[0]) for m in mixes]
    worst = 1e9
    for i in cpu_sample_indices(lens, k):
        with torch.no_grad():
            ref = model_cpu(torch.from_numpy(mixes[i])[None])[0].numpy()
        for s in range(ref.shape[0]):
            worst = min(worst, si_sdr_db(ref[s], outs[i][s]))
    return worst


def split_device_out(d_out, lens, n_src):
    got = d_out.cpu().numpy()
    outs, o = [], 0
    for n in lens:
        outs.append(got[o * n_src:(o + n) * n_src].reshape(n_src, n))
        o += n
    return outs


# ------------------------------------------------------------------------------------------------ sweep (configs 2-4)
def run_sweep(eng, pool, world, rank, dev, steps, variant, check_parity, model_cpu=None, sweep_batch_samples=2 << 20):
    """Separation + PIT of every segment of `pool` through sharding.evaluate_sharded (LPT shards, length buckets,
    device-resident separate_and_score, one all_gather).  Wall-clock per step bracketed by barriers, max over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from b200sep import sharding
    lens = pool.lens
    mine = sharding.lpt_partition([sharding.frames_of(v) for v in lens], world)[rank]
    mixes_l, refs_l = pool.materialise(mine)
    mixes = {i: m for i, m in zip(mine, mixes_l)}
    refs = {i: r for i, r in zip(mine, refs_l)}
    K = pool.n_src

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one():
        return sharding.evaluate_sharded(mixes, refs, SR, None, None, variant=variant, device=dev,
                                         score_stream=eng.score_stream, lens=lens, n_refs=K,
                                         max_samples=sweep_batch_samples)
    one()  # warm-up (workspaces, pinned staging, NCCL)
    times = []
    out = None
    for _ in range(steps):
        barrier()
        t0 = time.perf_counter()
        out = one()
        torch.cuda.synchronize(dev)
        times.append(time.perf_counter() - t0)
    t = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t[0])
    sisdr, sisdri, perm, _ = out
    audio = sum(lens) / SR
    res = {"workload_segments": len(lens), "audio_s": audio, "seconds_per_sweep": sec, "value": audio / sec, "unit": UNIT,
           "scaling": "strong", "ranks": world, "segments_this_rank": len(mine),
           "sep_sisdr_mean": float(np.nanmean(sisdr)), "sep_sisdri_mean": float(np.nanmean(sisdri)),
           "records_gathered": int(np.sum(~np.isnan(sisdr))),
           "api": "sharding.evaluate_sharded(score_stream=engine.score_stream): host arrays in, gathered records out, "
                  f"batches of <= {sweep_batch_samples} samples, two in flight"}
    if check_parity and rank == 0 and model_cpu is not None:
        # oracle flow on a bounded sample of THIS rank's shard: oracle forward + the reference's PIT restatement
        from oracle import pit_restated as pr
        fn = pr.compute_sdr_improvement_pit_k_o3 if variant == "overlap3" else None
        sample = [mine[j] for j in cpu_sample_indices([lens[i] for i in mine], 16)]
        d_i, agree = [], 0
        for i in sample:
            with torch.no_grad():
                est = model_cpu(torch.from_numpy(mixes[i])[None])[0].numpy()
            if variant == "overlap3":
                best, sdri, idx = fn(mixes[i], refs[i], [est[s] for s in range(est.shape[0])])
            elif K == 2:
                best, sdri, idx, _sw = pr.compute_sdr_improvement_pit_ews(mixes[i], refs[i][0], refs[i][1],
                                                                         [est[s] for s in range(est.shape[0])])
            else:
                best, sdri, idx = pr.compute_sdr_improvement_pit_k_ews(mixes[i], refs[i], [est[s] for s in range(est.shape[0])])
            d_i.append(sisdri[i] - sdri)
            agree += int(list(perm[i][:K]) == list(idx))
        res["parity"] = {"sample": len(sample), "delta_sisdri_mean_db": float(abs(np.mean(d_i))),
                         "max_abs_delta_sisdri_db": float(np.max(np.abs(d_i))), "perm_agree": agree,
                         "gate_db": GATE_SDRI_DB}
        assert abs(np.mean(d_i)) <= GATE_SDRI_DB, res["parity"]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32", "fp16_simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the configs[4] sweep object of a config-1 run")
    ap.add_argument("--no-eager", action="store_true", help="skip the torch-eager-on-B200 baseline")
    ap.add_argument("--sweep-segments", type=int, default=10000)
    ap.add_argument("--sustain-s", type=float, default=SUSTAIN_S, help="seconds of identical load in front of the timed steps (0: none, for ncu launch lists)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import b200sep
    from b200sep import weights, sharding
    from b200sep import _native as nat

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    peaks, peaks_kind = load_peaks()
    n_src = 2 if args.config == 2 else 3
    sd, margs = weights.random_state_dict(n_src, 0)
    sep = b200sep.Separator(backend="b200", device=f"cuda:{local_rank}", sample_rate=SR, n_src=n_src, state_dict=sd,
                            model_args=margs, precision=args.precision)
    eng = sep.engine
    n_layers = int(margs["n_blocks"]) * int(margs["n_repeats"])
    pipeline = int(nat.lib().b200sep_pipeline(eng._handle))
    model_cpu = oracle_model(n_src) if rank == 0 else None
    torch.set_num_threads(os.cpu_count() or 1)

    # ============================================================ configs 2-4: separation + PIT sweeps
    if args.config in (2, 3, 4):
        pool = {2: pool_config2, 3: pool_config3, 4: lambda: pool_config4(args.sweep_segments)}[args.config]()
        variant = "evalsrc" if args.config == 2 else "overlap3"
        sampler = ClockSampler(local_rank)
        sampler.start()
        res = run_sweep(eng, pool, world, rank, dev, max(1, min(args.steps, 5)), variant, True, model_cpu)
        clocks = sampler.stop()
        if rank == 0:
            frames = sum(sharding.frames_of(v) for v in pool.lens)
            line = {"metric": METRIC if n_src == 3 else "separated audio-sec/sec (8k, 2-src)", "value": res["value"], "unit": UNIT,
                    "n_gpus": world, "steps": max(1, min(args.steps, 5)), "warmup": 1, "ms_per_step": 1e3 * res["seconds_per_sweep"],
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": "f16" if args.precision != "fp32" else "f32", "data": "synthetic",
                    "config": {"workload": WORKLOADS[args.config], "segments": len(pool.lens), "audio_s": res["audio_s"],
                               "sample_rate": SR, "n_src": n_src, "precision": args.precision,
                               "parallelism": f"{world} rank(s): lpt_partition shards, one NCCL all_gather of records per sweep",
                               "l2": "every sweep streams > 1 GB of activations per batch: working set > 126 MB L2"},
                    "model_tflops": FLOP_PER_FRAME[n_src] * frames / res["seconds_per_sweep"] / 1e12,
                    "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 4 * sum(pool.lens) * (1 + n_src),
                            "d2h_bytes_per_step": 8 * 20 * len(pool.lens), "api": res["api"]},
                    "gpu_launches": None, "clocks": clocks, "sweep": res}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ============================================================ config 1: the headline
    pool = pool_config1(world)
    mine = sharding.lpt_partition([sharding.frames_of(v) for v in pool.lens], world)[rank]
    mixes, srcs = pool.materialise(mine)
    lens = [int(m.shape[0]) for m in mixes]
    total = int(sum(lens))
    audio_s = total / SR
    frames = sum((n - 16) // 8 + 1 for n in lens)
    # a large batch runs as concurrent tile-balanced sub-batches (each a full launch sequence on its own stream)
    n_sub = int(nat.lib().b200sep_split_count(eng._handle, nat.i32_array(lens), len(lens)))
    launches_per_step = launches_per_forward(n_layers, pipeline) * max(n_sub, 1)

    # ---------------- device-resident timing (inputs already in HBM) ----------------
    d_mix = torch.from_numpy(np.concatenate(mixes)).to(dev)
    ws = torch.empty(eng.workspace_bytes(lens) + 256, dtype=torch.uint8, device=dev)
    d_out = torch.empty(total * n_src, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def timed_steps(n):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for e0, e1 in evs:
            flush.zero_()  # evict L2 between timed steps (outside the event pair)
            e0.record()
            eng.separate_device(d_mix, lens, out=d_out, workspace=ws)
            e1.record()
        return evs

    for _ in range(max(args.warmup, 3)):
        eng.separate_device(d_mix, lens, out=d_out, workspace=ws)
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    # sustained lead-in: the same step, back to back, for >= SUSTAIN_S seconds (every step event-timed), then EXACTLY
    # `steps` timed steps with no gap in between: the headline is measured at the clocks the GPU holds under this load
    lead, t_lead0 = [], time.perf_counter()
    while time.perf_counter() - t_lead0 < args.sustain_s:
        lead.extend(timed_steps(10))
        torch.cuda.synchronize(dev)
    evs = timed_steps(args.steps)
    barrier()
    clocks = sampler.stop()
    eng.check()
    lead_ms = [e0.elapsed_time(e1) for e0, e1 in lead] or [float('nan')]
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    ms_per_step = sum(step_ms) / len(step_ms)
    t = torch.tensor([ms_per_step, audio_s, sum(lead_ms) / len(lead_ms), min(lead_ms[:10])], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_max, audio_all, lead_mean, burst_ms = float(tmax[0]), float(tsum[1]), float(tmax[2]), float(tmax[3])
    else:
        ms_max, audio_all, lead_mean, burst_ms = ms_per_step, audio_s, float(t[2]), float(t[3])
    value = audio_all / (ms_max * 1e-3)
    sustained = {"lead_in_steps": len(lead), "lead_in_seconds": args.sustain_s, "lead_in_ms_per_step": lead_mean,
                 "first_steps_ms": burst_ms, "value_lead_in": audio_all / (lead_mean * 1e-3),
                 "note": "the K timed steps follow the lead-in without a gap; first_steps_ms = fastest of the first 10 "
                         "lead-in steps (burst clocks)"}
    outs_dev = split_device_out(d_out, lens, n_src)  # what the last TIMED step wrote

    # ---------------- end to end through the public API (host buffers, H2D + D2H inside) -------
    e2e, outs = None, None
    if not args.no_e2e:
        for _ in range(max(args.warmup, 3)):  # same call pattern as the timed loop (the previous result is still
            outs = sep.separate_batch(mixes, SR)  # referenced while the next call runs: two pinned output buffers)
        barrier()
        n_e2e = max(args.steps, int(args.sustain_s / (ms_per_step * 1.3e-3)))
        # (a) one call at a time: Separator.separate_batch, every call synchronises before it returns
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            outs = sep.separate_batch(mixes, SR)
        torch.cuda.synchronize(dev)
        dt_sync = (time.perf_counter() - t0) / n_e2e
        assert len(outs) == len(mixes) and outs[0][0].shape[0] == lens[0]
        # (b) the streaming form of the same API: Separator.separate_stream keeps two batches in flight (own stream,
        # scratch and pinned buffers each), so the H2D / D2H copies of one step overlap the kernels of the next; every
        # step still gathers from the caller's NumPy arrays, copies 4 T bytes in and 4 S T bytes out
        import itertools
        for _o in sep.separate_stream(itertools.repeat(mixes, 4), SR):
            pass
        barrier()
        t0 = time.perf_counter()
        for _o in sep.separate_stream(itertools.repeat(mixes, n_e2e), SR):
            outs = _o
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / n_e2e
        assert len(outs) == len(mixes) and outs[0][0].shape[0] == lens[0]
        t = torch.tensor([dt, dt_sync], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": audio_all / float(t[0]), "unit": UNIT, "h2d_bytes_per_step": 4 * total,
               "d2h_bytes_per_step": 4 * total * n_src, "steps": n_e2e, "ms_per_step": 1e3 * float(t[0]),
               "api": "Separator.separate_stream(batches, sr): numpy in -> numpy out, two batches in flight "
                      "(each as ONE sub-batch: the batches in flight fill each other's launch tails)",
               "one_call_at_a_time": {"value": audio_all / float(t[1]), "ms_per_step": 1e3 * float(t[1]),
                                      "api": "Separator.separate_batch(chunks, sr)"}}

    # ---------------- parity of what was timed (rank 0): >= 40 dB vs the oracle's fp32 forward -------
    parity_db = None
    if rank == 0:
        parity_db = {"device_resident": parity_check(model_cpu, mixes, outs_dev), "gate_db": GATE_DB,
                     "segments_checked": 8, "fp32_reruns": eng.fallback_reruns}
        if outs is not None:
            parity_db["e2e"] = parity_check(model_cpu, mixes, [np.stack(o) for o in outs])
        assert parity_db["device_resident"] >= GATE_DB, parity_db
        assert parity_db.get("e2e", GATE_DB) >= GATE_DB, parity_db

    # ---------------- roofline pass: per-kernel CUDA-event times, same workload, GPU still under load ----------
    roofline, kernels = None, {}
    try:
        eng.profile_last_call(True)
        acc = {k: [] for k in KERNEL_CLASSES}
        reps = 5
        for _ in range(reps):
            flush.zero_()
            eng.separate_device(d_mix, lens, out=d_out, workspace=ws)
            per = classify_launches(eng.profile_read(), n_layers, pipeline)
            for k, v in per.items():
                acc[k].extend(v)
        eng.profile_last_call(False)
        es = 2 if args.precision != "fp32" else 4
        alg = {  # algorithmic work per launch over the VALID frames of the batch
            "encoder": ("hbm", 4 * total + 512 * frames * es),
            "bottleneck_gemm": ("tensor", 2 * frames * 512 * 128),
            "in_conv_gemm": ("tensor", 2 * frames * 128 * 512),
            "dwconv": ("hbm", 2 * 512 * frames * es),
            "res_skip_gemm": ("tensor", 2 * frames * 512 * 256),
            # fused block kernels: HBM bytes per frame = G_res 256 + x r/w 512 (fp16 residual stream) + h1 1024
            # (K_B); h1 1024 + G_res 256 + skip stash 256 (K_A); their GEMM FLOPs are reported beside as tensor_tflops;
            # skip merge: one 256 B stash slot per block read once + the fp16 mask operand
            "apply_in_gemm": ("hbm", 1792 * frames),
            "dw_rs_gemm": ("hbm", 1536 * frames),
            "final_skip": ("hbm", (256 * n_layers + 256) * frames),
            "mask_gemm": ("tensor", 2 * frames * 128 * 512 * n_src),
            "mask_dec": ("tensor", 2 * frames * (128 * 512 * n_src + 512 * 16 * n_src)),
            "decoder": ("hbm", n_src * 512 * frames * es + 4 * n_src * total),
        }
        step_total = sum(sum(v) / reps for v in acc.values() if v)
        # these launches run back to back inside a long loaded run: the SUSTAINED tensor peak is the denominator
        tpeak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        for k, v in acc.items():
            if not v:
                continue
            avg = sum(v) / len(v)
            bound, work = alg[k]
            if bound == "tensor":
                ach, peak, unit = work / (avg * 1e-3) / 1e12, tpeak, "TFLOP/s"
            else:
                ach, peak, unit = work / (avg * 1e-3) / 1e9, float(peaks["hbm_gbs"]), "GB/s"
            kernels[k] = {"launches_per_step": len(v) // reps, "avg_ms": avg, "share_of_step": (sum(v) / reps) / step_total,
                          "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak}
            gemm_flops = {"apply_in_gemm": 2 * frames * 128 * 512, "dw_rs_gemm": 2 * frames * 512 * 256}.get(k)
            if gemm_flops:
                kernels[k]["tensor_tflops"] = gemm_flops / (avg * 1e-3) / 1e12
                kernels[k]["tensor_frac_of_sustained_peak"] = kernels[k]["tensor_tflops"] / tpeak
        dom = max(kernels, key=lambda k: kernels[k]["share_of_step"])
        kd = kernels[dom]
        traffic = None
        try:  # DRAM bytes per launch of the same kernel on the same workload, from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f).get(dom, {}).get("bytes")
        except Exception:  # noqa: BLE001
            traffic = None
        roofline = {"kernel": dom, "bound": kd["bound"], "achieved": kd["achieved"], "peak": kd["peak"], "unit": kd["unit"],
                    "frac": kd["frac"], "traffic": traffic, "algorithmic_bytes_or_flops": alg[dom][1],
                    "peak_source": f"{peaks_kind} ({'bf16_tflops_sustained' if kd['bound'] == 'tensor' else 'hbm_gbs'})",
                    "share_of_step": kd["share_of_step"], "avg_launch_ms": kd["avg_ms"],
                    "note": "per-kernel times are CUDA-event brackets around every launch of the whole batch run as ONE "
                            "sub-batch (profiling serialises the launch sequence); the timed steps run it as "
                            f"{n_sub} concurrent sub-batches"}
        # ---- PIT kernel (replaces overlap3_core.py:48-117): 28 B/sample read once (K = N = 3), SURVEY.md 8(d) ----
        from b200sep.pit import pit_sisdr_device
        d_refs = torch.from_numpy(np.concatenate([np.concatenate(s) for s in srcs])).to(dev)
        d_est = d_out
        for _ in range(3):
            pit_sisdr_device(d_refs, d_est, d_mix, lens, n_src, n_src, "overlap3")
        import ctypes as C
        from b200sep import _native as nat
        nat.lib().b200sep_pit_profile_enable(1)   # CUDA events around the two launches inside the C entry point
        pk = []
        for _ in range(10):
            flush.zero_()
            pit_sisdr_device(d_refs, d_est, d_mix, lens, n_src, n_src, "overlap3")
            o2 = (C.c_float * 2)()
            nat.check(nat.lib().b200sep_pit_profile_read(o2), "b200sep_pit_profile_read")
            pk.append((float(o2[0]), float(o2[1])))
        nat.lib().b200sep_pit_profile_enable(0)
        pms = sum(k[0] for k in pk) / len(pk)
        pfin = sum(k[1] for k in pk) / len(pk)
        pbytes = (2 * n_src + 1) * 4 * total
        kernels["pit"] = {"launches_per_step": 2, "avg_ms": pms, "finalize_ms": pfin, "bound": "hbm",
                          "achieved": pbytes / (pms * 1e-3) / 1e9, "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                          "frac": pbytes / (pms * 1e-3) / 1e9 / float(peaks["hbm_gbs"]), "algorithmic_bytes": pbytes,
                          "note": "pit_moments kernel of b200sep_pit_sisdr_batch over the configs[1] batch (64 segments: "
                                  "a 30 MB stream, launch-latency sized; not part of the separation step -- "
                                  "tools/pit_target.py / --config 3 time it at 3000 segments)"}
    except Exception as e:  # noqa: BLE001
        roofline = {"error": str(e)}

    # ---------------- second baseline: the oracle in torch eager ON THE B200, called like the reference ----------
    eager = None
    if rank == 0 and not args.no_eager:
        try:
            mg = oracle_model(n_src, dev)
            with torch.no_grad():
                def sep_one(wav):  # Separator.separate, separation.py:76-89, device="cuda" (overlap3_core.py:120-124)
                    x = torch.from_numpy(wav).to(device=dev, dtype=torch.float32).unsqueeze(0)
                    return mg(x).squeeze(0).cpu().numpy()
                for i in range(3):
                    sep_one(mixes[i])
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                for m in mixes:
                    sep_one(m)
                torch.cuda.synchronize(dev)
                dt = time.perf_counter() - t0
            eager = {"value": audio_s / dt, "unit": UNIT, "seconds": dt, "segments": len(mixes),
                     "what": "oracle restatement (fp32, torch eager, library kernels) on this B200, one separate()-shaped call "
                             "per segment incl. H2D / D2H like /root/reference/src/osd/separation.py:76-89 with --provider cuda",
                     "allow_tf32_conv": bool(torch.backends.cudnn.allow_tf32)}
            del mg
        except Exception as e:  # noqa: BLE001
            eager = {"error": str(e)}

    # ---------------- configs[4] sweep through the real multi-GPU path ----------------
    sweep = None
    if not args.no_sweep:
        try:
            sweep = run_sweep(eng, pool_config4(args.sweep_segments), world, rank, dev, 2, "overlap3", False)
            sweep["workload"] = WORKLOADS[4]
        except Exception as e:  # noqa: BLE001
            sweep = {"error": str(e)}

    # ---------------- CPU baseline (oracle port on the host cores), rank 0 at N = 1 only -------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded sample of the same workload: whole passes over the batch until >= 10 s of CPU work (<= 4 passes)
        idx = list(range(len(lens)))
        a, tsec, passes, nthr = 0.0, 0.0, 0, 1
        while tsec < 10.0 and passes < 4:
            _, a1, t1, nthr = time_cpu_port(mixes, idx, n_src)
            a += a1; tsec += t1; passes += 1
        cpu_baseline = {"value": a / tsec, "unit": UNIT, "cores": nthr, "kind": "port", "host_cpus": os.cpu_count(),
                        "sample": f"{passes} pass(es) over all {len(lens)} segments of the step, one separate() call per segment "
                                  f"({a:.1f} audio-s, {tsec:.1f} s of CPU work)"}

    if rank == 0:
        whole_flops = FLOP_PER_FRAME[n_src] * frames * world  # LPT shards: every rank holds ~ the same frame count
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3) + len(lead),
            "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if args.precision != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[1], "segments_per_gpu": len(lens), "audio_s_per_gpu": audio_s, "sample_rate": SR,
                       "n_src": n_src, "precision": args.precision, "concurrent_sub_batches": n_sub,
                       "pipeline": {0: "unfused (3 kernels per TCN block)", 1: "fused (2 kernels per TCN block)",
                                    2: "fused (2 kernels per TCN block; mask GEMM + decoder in one kernel)"}[pipeline],
                       "parallelism": f"replica x{world}: one seeded pool of {64 * world} segments dealt out by sharding.lpt_partition, "
                                      "no data-path collective",
                       "range_conditioning": eng.conditioning is not None, "torch_threads": torch.get_num_threads(),
                       "l2": "L2 flushed (256 MiB write) between timed steps; per-step activation working set ~1 GB > 126 MB L2"},
            "model_tflops": whole_flops / (ms_max * 1e-3) / 1e12,
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "sustained": sustained,
            "parity_db": parity_db, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline,
            "torch_eager_b200": eager, "sweep": sweep,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
