#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native hot path.

Workload (BASELINE.json configs[1]): AAL3-style ROI mean/max pooling, 170 labels,
batch 64 of synthetic 1x91x109x91 volumes per GPU; subject-sharded (weak
scaling, no data-path collective) across N GPUs.

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                     # reference torch-CPU expression on host cores
    torchrun ... bench.py --gpus N ...                       # N > 1, one rank per GPU

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPE = (91, 109, 91)
N_ROIS = 170
BATCH = 64
WORKLOAD = "roi_pool_aal3like_170labels_batch64_1x91x109x91_f32"
# dram__bytes_read.sum + dram__bytes_write.sum of one roi_stream_kernel launch, from the committed
# ncu --set full capture (profiles/); None until a capture of the current kernel is committed.
TRAFFIC_NCU = 238.5e6   # profiles/r01_roi_stream_ncu_full.csv: 231.95 MB read + 6.6 MB written per launch


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _once(self):
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            names = {
                "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            }
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._once()
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self._once()
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def stop(self) -> dict:
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self._stop.set()
        if self._thr:
            self._thr.join()
        self._once()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def reference_step(feats, onehot):
    """One pass of the reference's own per-batch ROI expression on host cores
    (image_features.py:111-114, restated verbatim in oracle/roi_oracle.py; the one-hot
    mask of :80-82 is built once outside the loop, as the reference does)."""
    from oracle.roi_oracle import reference_pool_torch

    return reference_pool_torch(feats, onehot)


def run_reference_arm(args, rank: int):
    import numpy as np
    import torch

    from oracle.roi_oracle import reference_onehot_torch, synthetic_atlas

    if rank != 0:
        return
    lab = reference_onehot_torch(synthetic_atlas(SHAPE, N_ROIS))
    sample = 1                                   # volumes per step: the (B,R,C,D,H,W) product is 614 MB per volume
    g = torch.Generator().manual_seed(0)
    x = torch.rand((sample, 1) + SHAPE, generator=g)
    for _ in range(max(args.warmup, 1)):
        reference_step(x, lab)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        reference_step(x, lab)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "roi_pool_volumes_per_sec", "value": v, "unit": "volumes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "labels": N_ROIS, "volume": list(SHAPE)},
        "cpu_baseline": {"value": v, "unit": "volumes/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} volume per step of the same workload, reference torch expression "
                                   "image_features.py:80-82,111-114 on host cores"},
        "e2e": {"value": v, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_cuda_arm(args, rank: int, local_rank: int, world: int):
    import numpy as np
    import torch

    from multimodal_ad_b200 import RoiPlan, _lib
    from multimodal_ad_b200.sharding import max_over_ranks
    from oracle.roi_oracle import synthetic_atlas

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    lab = synthetic_atlas(SHAPE, N_ROIS)
    V = lab.size
    plan = RoiPlan(lab, N_ROIS, tile=args.tile, stages=args.stages)
    # three input batches (3 x 231 MB > 126 MB L2) visited round-robin: every step streams from HBM
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    bufs = [torch.rand((BATCH, V), device=dev, generator=g) for _ in range(3)]
    lib = _lib.load()

    def step(i):
        return plan.pool(bufs[i % 3])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        out = step(i)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = _lib.launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        ms = max_over_ranks(ms, dev)
    value = world * BATCH * args.steps / (ms * 1e-3)

    # --- dominant kernel alone: K back-to-back launches of the streaming kernel only (no finalize pass),
    #     CUDA events on the launching stream (torch's current stream is the one the C-ABI call is given) ---
    for i in range(3):
        plan.stream_only(bufs[i % 3])
    torch.cuda.synchronize()
    ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ka.record()
    for i in range(args.steps):
        plan.stream_only(bufs[i % 3])
    kb.record()
    kb.synchronize()
    k_avg_ms = ka.elapsed_time(kb) / args.steps
    alg_bytes = plan.algorithmic_bytes(BATCH)
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9

    # --- end to end through the public host-buffer API: pinned host -> device, pool, results -> host ---
    e2e = None
    if rank == 0 or world > 1:
        xh = torch.rand((BATCH, V)).pin_memory()
        plan.pool_host(xh)
        n_e2e = max(1, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            res = plan.pool_host(xh)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            dt = max_over_ranks(dt, dev)
        e2e = {"value": world * BATCH * n_e2e / dt, "unit": "volumes/s",
               "h2d_bytes_per_step": BATCH * V * 4, "d2h_bytes_per_step": BATCH * N_ROIS * 12, "steps": n_e2e}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.roi_oracle import reference_onehot_torch

        xs = torch.rand((1, 1) + SHAPE)
        onehot = reference_onehot_torch(lab)
        reference_step(xs, onehot)
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 10.0:
            reference_step(xs, onehot)
            n += 1
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": n / dt, "unit": "volumes/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{n} single-volume passes (~10 s) of the reference torch expression "
                                  "(image_features.py:80-82,111-114) on the host cores"}

    if rank == 0:
        line = {
            "metric": "roi_pool_volumes_per_sec", "value": value, "unit": "volumes/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "labels": N_ROIS, "volume": list(SHAPE), "batch_per_gpu": BATCH,
                       "parallelism": f"subject-sharded x{world}, no collective", "tile": plan.tile,
                       "l2": "3 input batches of 231 MB visited round-robin (each > 126 MB L2)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": TRAFFIC_NCU, "peak_source": peak_src, "kernel": "roi_stream_kernel<256,16>",
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_avg_ms},
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
    else:
        run_cuda_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
