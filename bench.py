#!/usr/bin/env python
"""bench.py -- separated audio-seconds per second of the overlap-segment separation stage.

    python bench.py --gpus N --steps K --warmup W            # the B200 backend (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one pass of the hot path (Conv-TasNet n_src=3 forward, 8 kHz) over one batch of
synthetic overlap segments: BASELINE.json configs[1] = 64 segments with durations U[0.2, 4.0] s
(~135 audio-seconds), seeded random-init weights.  With N > 1 every rank runs its own such batch
(independent segments, no data-path collective -> weak scaling); `value` = audio-seconds of all
ranks / max-over-ranks device time.  One JSON line is printed by rank 0.
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
N_SRC = 3
METRIC = "separated audio-sec/sec (8k, 3-src)"
UNIT = "audio-s/s"
FLOP_PER_FRAME = 9_969_664  # algorithmic, SURVEY.md 8(d)
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


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


def make_workload(rank: int, n_seg: int):
    from b200sep import synth
    mixes = synth.make_batch(n_seg, seed=2 + 1000 * rank, n_src=N_SRC, sr=SR, lo_s=0.2, hi_s=4.0)
    lens = [int(m.shape[0]) for m in mixes]
    return mixes, lens


def cpu_sample_indices(lens, k):
    """Stratified-by-length subsample: every (n/k)-th segment of the length-sorted batch."""
    order = sorted(range(len(lens)), key=lambda i: (lens[i], i))
    step = max(1, len(order) // k)
    return [order[i] for i in range(step // 2, len(order), step)][:k]


def time_cpu_port(mixes, idx, threads=None):
    """The reference's CPU path: per-segment loop, time.time() around every separate() call
    (/root/reference/scripts/osd/overlap3_core.py:689-691), torch CPU fp32, all host threads."""
    import numpy as np
    import torch
    from b200sep import weights
    from oracle import asteroid_restated as ar
    if threads:
        torch.set_num_threads(threads)
    sd, args = weights.random_state_dict(N_SRC, 0)
    model = ar.ConvTasNet(**args).eval()
    model.load_state_dict(sd)
    t_sep, audio = 0.0, 0.0
    with torch.no_grad():
        model(torch.from_numpy(mixes[idx[0]][:1600])[None])  # warm the allocator / thread pool
        for i in idx:
            t0 = time.time()
            x = torch.from_numpy(mixes[i]).to(dtype=torch.float32).unsqueeze(0)
            est = model(x)
            out = est.squeeze(0).cpu().numpy()
            _ = [out[s] for s in range(N_SRC)]
            t_sep += time.time() - t0
            audio += len(mixes[i]) / SR
    return audio / t_sep, audio, t_sep, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    mixes, lens = make_workload(0, args.segments)
    idx = cpu_sample_indices(lens, 8)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        time_cpu_port(mixes, idx[:2])
    t_tot, a_tot = 0.0, 0.0
    for _ in range(args.steps):
        _, audio, t, nthr = time_cpu_port(mixes, idx)
        t_tot += t; a_tot += audio
    value = a_tot / t_tot
    sample = f"{len(idx)} of the {len(lens)} segments per step, stratified by length ({a_tot / args.steps:.1f} audio-s per step)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 64 synthetic 3-speaker overlap segments U[0.2,4.0] s per GPU, 8 kHz, ConvTasNet n_src=3 random-init",
                   "segments_per_gpu": len(lens), "sample_rate": SR, "n_src": N_SRC, "precision": "fp32 (torch CPU)",
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32", "fp16_simt"])
    ap.add_argument("--segments", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    from b200sep import weights

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

    mixes, lens = make_workload(rank, args.segments)
    total = int(sum(lens))
    audio_s = total / SR
    frames = sum((n - 16) // 8 + 1 for n in lens)
    sd, margs = weights.random_state_dict(N_SRC, 0)
    sep = b200sep.Separator(backend="b200", device=f"cuda:{local_rank}", sample_rate=SR, n_src=N_SRC, state_dict=sd,
                            model_args=margs, precision=args.precision)
    eng = sep.engine
    n_layers = int(margs["n_blocks"]) * int(margs["n_repeats"])
    from b200sep import _native as nat
    pipeline = int(nat.lib().b200sep_pipeline(eng._handle))
    # a large batch runs as concurrent tile-balanced sub-batches (each a full launch sequence on its own stream)
    n_sub = int(nat.lib().b200sep_split_count(eng._handle, nat.i32_array(lens), len(lens)))
    launches_per_step = launches_per_forward(n_layers, pipeline) * max(n_sub, 1)

    # ---------------- device-resident timing (inputs already in HBM) ----------------
    d_mix = torch.from_numpy(np.concatenate(mixes)).to(dev)
    ws = torch.empty(eng.workspace_bytes(lens) + 256, dtype=torch.uint8, device=dev)
    d_out = torch.empty(total * N_SRC, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    for _ in range(max(args.warmup, 3)):
        eng.separate_device(d_mix, lens, out=d_out, workspace=ws)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for e0, e1 in evs:
        flush.zero_()  # evict L2 between timed steps (outside the event pair)
        e0.record()
        eng.separate_device(d_mix, lens, out=d_out, workspace=ws)
        e1.record()
    barrier()
    clocks = sampler.stop()
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    ms_per_step = sum(step_ms) / len(step_ms)
    t = torch.tensor([ms_per_step, audio_s], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_max, audio_all = float(tmax[0]), float(tsum[1])
    else:
        ms_max, audio_all = ms_per_step, audio_s
    value = audio_all / (ms_max * 1e-3)

    # ---------------- end to end through the public API (host buffers, H2D + D2H inside) -------
    e2e = None
    if not args.no_e2e:
        outs = None
        for _ in range(max(args.warmup, 3)):  # same call pattern as the timed loop (the previous result is still
            outs = sep.separate_batch(mixes, SR)  # referenced while the next call runs: two pinned output buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            outs = sep.separate_batch(mixes, SR)
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / args.steps
        assert len(outs) == len(mixes) and outs[0][0].shape[0] == lens[0]
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": audio_all / float(t[0]), "unit": UNIT, "h2d_bytes_per_step": 4 * total,
               "d2h_bytes_per_step": 4 * total * N_SRC, "api": "Separator.separate_batch(chunks, sr) numpy in -> numpy out"}

    # ---------------- roofline pass: per-kernel CUDA-event times, same workload ----------------
    peaks, peaks_kind = load_peaks()
    roofline, kernels = None, {}
    try:
        eng.profile_last_call(True)
        acc = {k: [] for k in KERNEL_CLASSES}
        for _ in range(3):
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
            "mask_gemm": ("tensor", 2 * frames * 128 * 512 * N_SRC),
            "mask_dec": ("tensor", 2 * frames * (128 * 512 * N_SRC + 512 * 16 * N_SRC)),
            "decoder": ("hbm", N_SRC * 512 * frames * es + 4 * N_SRC * total),
        }
        step_total = sum(sum(v) / 3.0 for v in acc.values() if v)
        for k, v in acc.items():
            if not v:
                continue
            avg = sum(v) / len(v)
            bound, work = alg[k]
            if bound == "tensor":
                ach, peak, unit = work / (avg * 1e-3) / 1e12, float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))), "TFLOP/s"
            else:
                ach, peak, unit = work / (avg * 1e-3) / 1e9, float(peaks["hbm_gbs"]), "GB/s"
            kernels[k] = {"launches_per_step": len(v) // 3, "avg_ms": avg, "share_of_step": (sum(v) / 3.0) / step_total,
                          "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak}
            gemm_flops = {"apply_in_gemm": 2 * frames * 128 * 512, "dw_rs_gemm": 2 * frames * 512 * 256}.get(k)
            if gemm_flops:
                kernels[k]["tensor_tflops"] = gemm_flops / (avg * 1e-3) / 1e12
        dom = max(kernels, key=lambda k: kernels[k]["share_of_step"])
        kd = kernels[dom]
        traffic = None
        try:  # DRAM bytes per launch of the same kernel on the same workload, from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f).get(dom, {}).get("bytes")
        except Exception:  # noqa: BLE001
            traffic = None
        roofline = {"kernel": dom, "bound": kd["bound"], "achieved": kd["achieved"], "peak": kd["peak"], "unit": kd["unit"],
                    "frac": kd["frac"], "traffic": traffic, "algorithmic_bytes_or_flops": alg[dom][1], "peak_source": f"{peaks_kind} ({'bf16_tflops_sustained' if kd['bound'] == 'tensor' else 'hbm_gbs'})",
                    "share_of_step": kd["share_of_step"], "avg_launch_ms": kd["avg_ms"],
                    "note": "per-kernel times are CUDA-event brackets around every launch of the whole batch run as ONE "
                            "sub-batch (profiling serialises the launch sequence); the timed steps run it as "
                            f"{n_sub} concurrent sub-batches"}
    except Exception as e:  # noqa: BLE001
        roofline = {"error": str(e)}

    # ---------------- CPU baseline (oracle port on the host cores), rank 0 at N = 1 only -------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded sample of the same workload: whole passes over the batch until >= 10 s of CPU work (<= 4 passes)
        idx = list(range(len(lens)))
        a, tsec, passes, nthr = 0.0, 0.0, 0, 1
        while tsec < 10.0 and passes < 4:
            _, a1, t1, nthr = time_cpu_port(mixes, idx)
            a += a1; tsec += t1; passes += 1
        cpu_baseline = {"value": a / tsec, "unit": UNIT, "cores": nthr, "kind": "port", "host_cpus": os.cpu_count(),
                        "sample": f"{passes} pass(es) over all {len(lens)} segments of the step, one separate() call per segment "
                                  f"({a:.1f} audio-s, {tsec:.1f} s of CPU work)"}

    if rank == 0:
        whole_flops = FLOP_PER_FRAME * frames * world  # ranks have statistically identical batches
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if args.precision != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: 64 synthetic 3-speaker overlap segments U[0.2,4.0] s per GPU, 8 kHz, ConvTasNet n_src=3 random-init",
                       "segments_per_gpu": len(lens), "audio_s_per_gpu": audio_s, "sample_rate": SR, "n_src": N_SRC,
                       "precision": args.precision, "concurrent_sub_batches": n_sub, "pipeline": {0: "unfused (3 kernels per TCN block)", 1: "fused (2 kernels per TCN block)", 2: "fused (2 kernels per TCN block; mask GEMM + decoder in one kernel)"}[pipeline], "parallelism": f"replica x{world} (segments sharded, no data-path collective)",
                       "l2": "L2 flushed (256 MiB write) between timed steps; per-step activation working set ~1 GB > 126 MB L2"},
            "model_tflops": whole_flops / (ms_max * 1e-3) / 1e12,
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
