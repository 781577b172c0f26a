"""Shared measurement helpers of bench.py / bench_phi.py / bench_configs.py: clock sampling during the timed
region, the driver-written peaks, CUDA-event timing with max-over-ranks, and the source hash that ties a
committed ncu traffic figure to the kernel sources it was measured on."""
import hashlib
import json
import os
import subprocess
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))


class ClockSampler:
    """nvidia-smi sampled every 100 ms while a timed region runs (B200_PROFILING.md's clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    """MEASURED_PEAKS.json (driver-written) or the fallbacks B200_PROFILING.md states."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    out = {"hbm_gbs": 6650.0, "bf16_tflops": 1600.0, "bf16_tflops_sustained": None,
           "source": "fallback (B200_PROFILING.md; MEASURED_PEAKS.json absent)"}
    if os.path.isfile(p):
        try:
            d = json.load(open(p))
            out.update(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                       bf16_tflops_sustained=float(d.get("bf16_tflops_sustained") or 0) or None,
                       source="MEASURED_PEAKS.json")
        except Exception:
            pass
    return out


def source_sha(files):
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(ROOT, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def committed_traffic(name, files):
    """DRAM bytes per launch of a kernel from profiles/traffic.json -- only when the sources the ncu capture was
    taken on are still the ones in the tree (else None: a stale figure is worse than none)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        ent = json.load(open(p))[name]
    except Exception:
        return None, "no committed ncu capture"
    if ent.get("source_sha") != source_sha(files):
        return None, "committed capture predates the current kernel sources"
    return ent.get("dram_bytes_per_launch"), ent.get("note", "")


def timed_ms(fn, steps, warmup, dev, world=1, sync=None):
    """average ms per call over `steps` calls (CUDA events on the current stream, after `warmup` calls;
    barrier + synchronize on both sides; max over ranks)."""
    import torch.distributed as dist

    def fence():
        if sync is not None:
            sync()
        elif world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        fn()
    fence()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    fence()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms[0])


def cublas_tf32_tflops(dev, n=8192, reps=5):
    """cuBLAS TF32 GEMM throughput measured in this run (the tensor-pipe denominator MEASURED_PEAKS.json lacks)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a, b = torch.randn(n, n, device=dev), torch.randn(n, n, device=dev)
        for _ in range(2):
            a @ b
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
