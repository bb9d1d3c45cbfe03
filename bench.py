#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native hot path.

Top-level line (BASELINE.json configs[2], the metric BASELINE.json leads with): ResNet3D-18 bf16 TRAINING, batch 16 per
GPU, synthetic 1x128^3 volumes, 3 classes, data parallel across N GPUs (gradient all-reduce overlapped with backward).
One step = forward + CE loss + backward + gradient all-reduce + clip + Adam (train_ResNet3D.py:207-218).

Nested objects on the same JSON line (each with its own roofline):
  roi_pool                    configs[1]: AAL3-style ROI mean/max pooling, 170 labels, batch 64 x 1x91x109x91 per GPU
  resnet3d18_train_91x109x91  the same training step on the reference's own volume (config/config_unet.json input_D/H/W)
  resnet3d50_train            configs[3] as the reference can run it (cfg_denseNet.json: model_type resnet, depth 50)
  unet3d_roi_extract          configs[4]'s image branch: UNet3D forward (eval) + on-device ROI pooling of the 64-channel map
  unet3d_train                UNet3D forward + backward + Adam (models/unet3d.py stacks through the same kernels)
  torch_gpu_baseline          stock PyTorch / cuDNN on the same GPU for the same ResNet3D-18 step (context, not the target)

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                     # the reference's PyTorch CPU path on the host cores
    torchrun ... bench.py --gpus N ...                       # N > 1, one rank per GPU

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import csv
import glob
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
ROI_WORKLOAD = "roi_pool_aal3like_170labels_batch64_1x91x109x91_f32"
RESNET_WORKLOAD = "resnet3d18_bf16_train_batch16_1x128^3_3class"
RESNET_BATCH, RESNET_SIZE = 16, 128


def peaks_json():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


def measured_peaks():
    d = peaks_json()
    if "hbm_gbs" in d:
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def tensor_peak():
    d = peaks_json()
    if "bf16_tflops_sustained" in d:
        return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"


def roi_traffic_from_profile():
    """dram__bytes_read.sum + dram__bytes_write.sum of one roi_stream_kernel launch, read from the newest committed
    `ncu --set full` summary under profiles/ (None when there is none)."""
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_roi_stream_ncu_full.csv"))):
        rd = wr = None
        with open(path) as f:
            for row in csv.reader(l for l in f if not l.startswith("#")):
                if len(row) >= 3 and row[0] == "dram__bytes_read.sum":
                    rd = float(row[2]) * (1e6 if row[1] == "Mbyte" else 1.0)
                if len(row) >= 3 and row[0] == "dram__bytes_write.sum":
                    wr = float(row[2]) * (1e6 if row[1] == "Mbyte" else 1.0)
        if rd is not None and wr is not None:
            best = (rd + wr, os.path.relpath(path, ROOT))
    return best if best else (None, None)


def step_tensor_pipe_from_profile():
    """Step-level sm__pipe_tensor_cycles_active (time-weighted over every launch of one training step), written by
    tools/step_tensor_pipe.py from an ncu pass and committed under profiles/; None when absent."""
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_resnet_step_tensor_pipe.json")))
    if not paths:
        return None
    with open(paths[-1]) as f:
        d = json.load(f)
    d["source"] = os.path.relpath(paths[-1], ROOT)
    return d


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


# ---------------------------------------------------------------------------------------------------------------------
# ResNet3D training step (configs[2], configs[0]'s volume, configs[3])
# ---------------------------------------------------------------------------------------------------------------------
def resnet_conv_flops_model(model, batch, shape):
    """Algorithmic conv FLOPs of one training step of any models/resnet.py network on 1 x D x H x W volumes: forward + wgrad
    for every convolution, + dgrad for all but the stem (walks the module tree, resnet.py:126-143).  Counts every tap of
    every output voxel, including the taps that fall into the zero padding (which the kernels skip)."""
    def out(n, k, s, p, d):
        return (n + 2 * p - d * (k - 1) - 1) // s + 1

    def vox(sp):
        return sp[0] * sp[1] * sp[2]

    s1 = tuple(out(n, 7, 2, 3, 1) for n in shape)
    sp = tuple(out(n, 3, 2, 1, 1) for n in s1)                        # after conv1 and the max-pool
    total = 2 * 2.0 * batch * vox(s1) * 64 * 343
    for layer in (model.layer1, model.layer2, model.layer3, model.layer4):
        for blk in layer:
            cur = sp
            for name in ("conv1", "conv2", "conv3"):
                conv = getattr(blk, name, None)
                if conv is None:
                    continue
                k, st, dil, pad = conv.kernel_size[0], conv.stride[0], conv.dilation[0], conv.padding[0]
                cur = tuple(out(n, k, st, pad, dil) for n in cur)
                total += 3 * 2.0 * batch * vox(cur) * conv.out_channels * conv.in_channels * k ** 3
            ds = blk.downsample
            if ds is not None and not callable(getattr(ds, "func", None)):
                total += 3 * 2.0 * batch * vox(cur) * ds[0].out_channels * ds[0].in_channels
            sp = cur
    return total


def run_resnet_train(args, rank, world, dev, dist, *, depth=18, batch=RESNET_BATCH, shape=(RESNET_SIZE,) * 3, nb_class=3,
                     steps=20, warmup=5, workload=RESNET_WORKLOAD, metric="resnet3d18_train_volumes_per_sec", do_e2e=True,
                     sample_clocks=False, local_rank=0):
    """One ResNet3D training workload: `warmup` untimed steps, then exactly `steps` timed steps (CUDA events, barrier +
    synchronize on both sides, max over ranks)."""
    import torch
    import torch.nn as nn

    from multimodal_ad_b200 import _lib
    from multimodal_ad_b200.models.Resnet3D import generate_model
    from multimodal_ad_b200.sharding import GradReducer, max_over_ranks

    torch.manual_seed(0)                               # same initial weights on every rank
    model = generate_model(model_depth=depth, input_W=shape[2], input_H=shape[1], input_D=shape[0], nb_class=nb_class,
                           pretrain_path=None, dropout_rate=0.5, device=dev)
    model.train()
    reducer = GradReducer(bucket_numel=int(os.environ.get("MMAD_BUCKET_NUMEL", str(1 << 23))))
    model.grad_reducer = reducer
    opt = torch.optim.Adam(model.parameters(), lr=1e-5, weight_decay=1e-4, fused=True, capturable=True)
    crit = nn.CrossEntropyLoss()
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    xs = [torch.rand((batch, 1) + tuple(shape), device=dev, generator=g) for _ in range(2)]   # 2 x 134 MB at 16 x 128^3
    ys = [torch.randint(0, nb_class, (batch,), device=dev, generator=g) for _ in range(2)]

    def step(i, x=None, y=None):
        out = model(xs[i % 2] if x is None else x)
        loss = crit(out, ys[i % 2] if y is None else y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        reducer.finish(model.parameters())
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        opt.step()
        return loss

    for i in range(max(3, warmup)):
        step(i)
    torch.cuda.synchronize()

    # The whole step (our launches, the torch loss / clip / Adam kernels and - under data parallel - the NCCL all-reduces the
    # GradReducer issues on its side stream) is captured once into a CUDA graph and replayed, so the step time does not depend
    # on the host's launch rate (eager enqueue costs 8-11 ms per step) and every N runs in the same launch mode.  Falls back to
    # the eager step when capture is unavailable (MMAD_BENCH_DP_GRAPH=0 keeps data parallel eager).
    graph, graph_note, graph_launches = None, "eager", 0
    want_graph = not getattr(args, "no_graph", False) and (world == 1 or os.environ.get("MMAD_BENCH_DP_GRAPH", "1") != "0")
    if want_graph:
        try:
            sx, sy = xs[0].clone(), ys[0].clone()
            opt.zero_grad(set_to_none=True)
            graph = torch.cuda.CUDAGraph()
            c0 = _lib.launch_count()
            with torch.cuda.graph(graph):
                gloss = crit(model(sx), sy)
                gloss.backward()
                reducer.finish(model.parameters())
                torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
                opt.step()
            graph_launches = _lib.launch_count() - c0
            graph.replay()
            torch.cuda.synchronize()
            graph_note = "cuda graph replay" + (" (NCCL all-reduces captured)" if world > 1 else "")
        except Exception as e:                             # noqa: BLE001 - any capture problem falls back to the eager step
            graph, graph_note = None, f"eager (graph capture failed: {type(e).__name__}: {str(e)[:80]})"
            torch.cuda.synchronize()
            opt.zero_grad(set_to_none=True)

    if graph is not None:
        def step(i, x=None, y=None):                       # noqa: F811 - same contract as the eager step above
            sx.copy_(xs[i % 2] if x is None else x, non_blocking=True)
            sy.copy_(ys[i % 2] if y is None else y, non_blocking=True)
            graph.replay()
            return gloss
        for i in range(2):
            step(i)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = None
    if sample_clocks:
        sampler = ClockSampler(local_rank)
        sampler.start()
    f0 = _lib.executed_mma_flops()
    l0 = _lib.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        loss = step(i)
    b.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler is not None else None
    ms = max_over_ranks(a.elapsed_time(b), dev) / steps
    executed = (_lib.executed_mma_flops() - f0) / steps
    launches = graph_launches * steps if graph is not None else (_lib.launch_count() - l0)
    flops = resnet_conv_flops_model(model, batch, shape)
    peak, peak_src = tensor_peak()
    ach = flops / (ms * 1e-3) / 1e12
    res = {
        "metric": metric, "value": world * batch / (ms * 1e-3), "unit": "volumes/s",
        "ms_per_step": ms, "steps": steps, "warmup": max(3, warmup), "dtype": "bf16", "scaling": "weak",
        "config": {"workload": workload, "batch_per_gpu": batch, "volume": list(shape),
                   "parallelism": f"data parallel x{world}, gradient all-reduce overlapped with backward",
                   "step": "forward + CE loss + backward + grad clip + Adam (train_ResNet3D.py:207-218)", "launch": graph_note,
                   "l2": f"2 input batches of {batch * shape[0] * shape[1] * shape[2] * 4 / 1e6:.0f} MB alternate; the step's "
                         "activations (several GB) stream through HBM every step"},
        "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_flops_per_step": flops,
                     "executed_flops_per_step": executed, "executed_tflops": executed / (ms * 1e-3) / 1e12,
                     "executed_frac": executed / (ms * 1e-3) / 1e12 / peak,
                     "note": "algorithmic = every tap of every output voxel (zero-padding taps included); executed = 2*M*N*16 per "
                             "tcgen05.mma actually issued (padding taps skipped, partially filled tiles counted whole)"},
        "gpu_launches": launches, "loss": float(loss.detach()),
    }
    if clocks is not None:
        res["clocks"] = clocks
    if do_e2e:
        # end to end: pinned host batch -> device, step, loss back on the host EVERY step (float(loss) synchronises, as the
        # reference's `loss.item()` does).  The copy of batch i+1 runs on a copy stream under step i, the way a pinned-memory
        # prefetching loader feeds a training loop; all copies are inside the timed region.
        xh = torch.rand((batch, 1) + tuple(shape)).pin_memory()
        yh = torch.randint(0, nb_class, (batch,)).pin_memory()
        copy_stream = torch.cuda.Stream(device=dev)
        bufs = [(torch.empty((batch, 1) + tuple(shape), device=dev), torch.empty((batch,), dtype=torch.int64, device=dev)) for _ in range(2)]
        evs = [torch.cuda.Event(), torch.cuda.Event()]

        def fetch(j):
            with torch.cuda.stream(copy_stream):
                bufs[j % 2][0].copy_(xh, non_blocking=True)
                bufs[j % 2][1].copy_(yh, non_blocking=True)
                evs[j % 2].record(copy_stream)

        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        n_e2e = max(1, min(steps, 10))
        t0 = time.perf_counter()
        fetch(0)
        for i in range(n_e2e):
            torch.cuda.current_stream().wait_event(evs[i % 2])
            if i + 1 < n_e2e:
                fetch(i + 1)                               # buffer (i+1) % 2 was last read by step i-1, which has completed
            float(step(i, bufs[i % 2][0], bufs[i % 2][1]).detach())
        dt = max_over_ranks(time.perf_counter() - t0, dev)
        res["e2e"] = {"value": world * batch * n_e2e / dt, "unit": "volumes/s",
                      "h2d_bytes_per_step": batch * shape[0] * shape[1] * shape[2] * 4 + batch * 8, "d2h_bytes_per_step": 4, "steps": n_e2e}
    del model, opt, xs
    torch.cuda.empty_cache()
    return res


def stock_torch_forward(m, x):
    """The reference's ResNet forward (resnet.py:54-69, 204-213) through the nn.Conv3d / BatchNorm3d / MaxPool3d modules'
    OWN forward (stock PyTorch / cuDNN kernels) - the model object is the drop-in's parameter container, nothing of this
    repo's CUDA library runs here.  BasicBlock networks only."""
    x = m.maxpool(m.relu(m.bn1(m.conv1(x))))
    for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
        for blk in layer:
            res = x
            out = blk.relu(blk.bn1(blk.conv1(x)))
            out = blk.bn2(blk.conv2(out))
            if blk.downsample is not None:
                res = blk.downsample(x)
            x = blk.relu(out + res)
    return m.conv_seg(x)


def torch_gpu_baseline(dev, batch=RESNET_BATCH, size=RESNET_SIZE, steps=5):
    """Same-box context number: the same ResNet3D-18 training step (same module tree, loss, clip, Adam) run by stock
    PyTorch + cuDNN on this GPU, (a) fp32 with TF32 allowed, (b) bf16 autocast with channels_last_3d.  Not the target and
    not the --impl reference arm: it tells what train_ResNet3D.py:73 (`net.to(device)`) gets for free on a B200."""
    import torch
    import torch.nn as nn

    from multimodal_ad_b200.models.Resnet3D import generate_model

    out = {"workload": RESNET_WORKLOAD, "batch": batch, "steps": steps,
           "what": "stock torch.nn modules (cuDNN) fwd + CE + bwd + clip + Adam, same module tree as the accelerated model"}
    x = torch.rand((batch, 1, size, size, size), device=dev)
    y = torch.randint(0, 3, (batch,), device=dev)
    for name, autocast, cl in (("tf32", False, False), ("bf16_autocast_channels_last_3d", True, True)):
        try:
            torch.manual_seed(0)
            model = generate_model(model_depth=18, input_W=size, input_H=size, input_D=size, nb_class=3, pretrain_path=None,
                                   dropout_rate=0.5, device=dev).train()
            if cl:
                model = model.to(memory_format=torch.channels_last_3d)
            opt = torch.optim.Adam(model.parameters(), lr=1e-5, weight_decay=1e-4, fused=True)
            crit = nn.CrossEntropyLoss()
            old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.benchmark = True
            xin = x.contiguous(memory_format=torch.channels_last_3d) if cl else x

            def step():
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    loss = crit(stock_torch_forward(model, xin).float(), y)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
                opt.step()
                return loss

            for _ in range(3):
                step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                step()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / steps
            out[name] = {"ms_per_step": ms, "volumes_per_sec": batch / (ms * 1e-3)}
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
            del model, opt
        except Exception as e:                             # noqa: BLE001 - a context number must not take the bench down
            out[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    return out


def resnet_cpu_reference(steps, warmup, size=RESNET_SIZE, batch=2, seconds_cap=None):
    """The reference's PyTorch CPU path for the headline workload: ResNet3D-18 forward + CE loss + backward, fp32, on the
    host cores (oracle/resnet_oracle.py restates resnet.py and is pinned bit-exact against it).  One step = a BOUNDED
    SAMPLE of the workload: batch `batch` instead of 16 of the same 1 x size^3 volumes (per-volume cost is batch
    independent on a CPU).  Returns volumes/s over exactly `steps` timed steps (or as many as fit in seconds_cap)."""
    import torch
    import torch.nn.functional as F

    from multimodal_ad_b200.models import resnet
    from oracle.resnet_oracle import classifier_head_oracle, resnet_features_oracle

    torch.set_num_threads(max(1, os.cpu_count() or 1))      # torchrun exports OMP_NUM_THREADS=1: use every host core anyway
    torch.manual_seed(0)
    m = resnet.resnet18(sample_input_D=size, sample_input_H=size, sample_input_W=size, num_seg_classes=1)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in m.state_dict().items()}
    fw, fb = torch.randn(3, 512, requires_grad=True), torch.zeros(3, requires_grad=True)
    x, y = torch.rand(batch, 1, size, size, size), torch.randint(0, 3, (batch,))

    def step():
        loss = F.cross_entropy(classifier_head_oracle(resnet_features_oracle(sd, x, [2, 2, 2, 2], True), fw, fb), y)
        loss.backward()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    n = 0
    while n < steps:
        step()
        n += 1
        if seconds_cap is not None and time.perf_counter() - t0 > seconds_cap:
            break
    dt = time.perf_counter() - t0
    return {"value": batch * n / dt, "unit": "volumes/s", "cores": torch.get_num_threads(), "kind": "port",
            "ms_per_step": 1e3 * dt / n, "steps": n,
            "sample": f"{n} step(s) of batch {batch} (instead of 16) x 1x{size}^3 fp32 forward + CE + backward through the pinned "
                      "restatement of models/resnet.py on the host cores"}


# ---------------------------------------------------------------------------------------------------------------------
# ROI pooling (configs[1])
# ---------------------------------------------------------------------------------------------------------------------
def reference_step(feats, onehot):
    """One pass of the reference's own per-batch ROI expression on host cores
    (image_features.py:111-114, restated verbatim in oracle/roi_oracle.py; the one-hot
    mask of :80-82 is built once outside the loop, as the reference does)."""
    from oracle.roi_oracle import reference_pool_torch

    return reference_pool_torch(feats, onehot)


def roi_cpu_reference(seconds=8.0):
    import torch

    from oracle.roi_oracle import reference_onehot_torch, synthetic_atlas

    onehot = reference_onehot_torch(synthetic_atlas(SHAPE, N_ROIS))
    xs = torch.rand((1, 1) + SHAPE)
    reference_step(xs, onehot)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        reference_step(xs, onehot)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "volumes/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} single-volume passes (~{seconds:.0f} s) of the reference torch expression "
                      "(image_features.py:80-82,111-114) on the host cores"}


def run_roi_pool(args, rank, world, dev, dist, steps, warmup, want_cpu):
    import torch

    from multimodal_ad_b200 import RoiPlan, _lib
    from multimodal_ad_b200.sharding import max_over_ranks
    from oracle.roi_oracle import synthetic_atlas                       # label-map generator only (outside every timed region)

    lab = synthetic_atlas(SHAPE, N_ROIS)
    V = lab.size
    plan = RoiPlan(lab, N_ROIS, tile=args.tile, stages=args.stages)
    # three input batches (3 x 231 MB > 126 MB L2) visited round-robin: every step streams from HBM
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    bufs = [torch.rand((BATCH, V), device=dev, generator=g) for _ in range(3)]

    def step(i):
        return plan.pool(bufs[i % 3])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(warmup, 3)):
        step(i)
    barrier()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the K steps are queued behind a short spin kernel (outside the event bracket) so that the device-side time of exactly
    # K steps is measured even when a slow host core cannot launch a 46 us step every 46 us
    torch.cuda._sleep(int(max(0.01, steps * 1.0e-4) * 1.9e9))
    ev0.record()
    for i in range(steps):
        step(i)
    ev1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        ms = max_over_ranks(ms, dev)
    value = world * BATCH * steps / (ms * 1e-3)

    # --- dominant kernel alone: K back-to-back launches of the streaming kernel only (no finalize pass),
    #     CUDA events on the launching stream (torch's current stream is the one the C-ABI call is given) ---
    for i in range(3):
        plan.stream_only(bufs[i % 3])
    torch.cuda.synchronize()
    ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(max(0.01, steps * 1.0e-4) * 1.9e9))
    ka.record()
    for i in range(steps):
        plan.stream_only(bufs[i % 3])
    kb.record()
    kb.synchronize()
    k_avg_ms = ka.elapsed_time(kb) / steps
    alg_bytes = plan.algorithmic_bytes(BATCH)
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9
    traffic, traffic_src = roi_traffic_from_profile()

    # --- end to end through the public host-buffer API: pinned host -> device, pool, results -> host ---
    xh = torch.rand((BATCH, V)).pin_memory()
    plan.pool_host(xh)
    n_e2e = max(1, min(steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        plan.pool_host(xh)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist is not None:
        dt = max_over_ranks(dt, dev)
    res = {
        "metric": "roi_pool_volumes_per_sec", "value": value, "unit": "volumes/s", "steps": steps, "ms_per_step": ms / steps,
        "dtype": "f32", "scaling": "weak",
        "config": {"workload": ROI_WORKLOAD, "labels": N_ROIS, "volume": list(SHAPE), "batch_per_gpu": BATCH,
                   "parallelism": f"subject-sharded x{world}, no collective", "tile": plan.tile,
                   "l2": "3 input batches of 231 MB visited round-robin (each > 126 MB L2)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernel": f"roi_stream_kernel<{plan.tile},16>", "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_avg_ms},
        "e2e": {"value": world * BATCH * n_e2e / dt, "unit": "volumes/s", "h2d_bytes_per_step": BATCH * V * 4,
                "d2h_bytes_per_step": BATCH * N_ROIS * 12, "steps": n_e2e},
        "gpu_launches": launches,
    }
    if want_cpu:
        res["cpu_baseline"] = roi_cpu_reference()
    del bufs, plan
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------------------------------------------------------
# UNet3D forward + on-device ROI features (configs[4]'s image branch; image_features.py:97-114)
# ---------------------------------------------------------------------------------------------------------------------
def unet3d_forward_flops(grid=(96, 112, 96)):
    """Algorithmic FLOPs of one UNet3D(1, 1) forward on one volume (unet3d.py:100-113 channel plan on the padded grid)."""
    v0 = grid[0] * grid[1] * grid[2]
    v = [v0, v0 // 8, v0 // 64, v0 // 512]
    f = 0.0
    enc = [(1, 32, 64), (64, 64, 128), (128, 128, 256), (256, 256, 512)]
    for lvl, (cin, mid, cout) in enumerate(enc):
        f += 2.0 * v[lvl] * 27 * (cin * mid + mid * cout)
    for lvl, (cup, cres) in ((2, (512, 256)), (1, (256, 128)), (0, (128, 64))):
        f += 2.0 * v[lvl] * cup * cup                      # transposed convolution: one tap per output voxel
        mid = cup // 2
        f += 2.0 * v[lvl] * 27 * ((cup + cres) * mid + mid * mid)
    f += 2.0 * v0 * 64
    return f


def unet_cpu_reference(seconds_cap=40.0):
    """CPU leg for the extraction path: oracle UNet3D forward (pinned bit-exact to models/unet3d.py) in eval mode + the ROI means of
    the hooked tensor, one 1x91x109x91 volume per pass, fp32 on the host cores.  The reference's own pooling expression
    (image_features.py:111-114) would materialise a (1,170,64,91,109,91) product = 39 GB per volume for this 64-channel map;
    the per-ROI restatement (oracle/unet_oracle.py, pinned to that expression on small cases) is timed instead."""
    import torch

    from multimodal_ad_b200.models import unet3d
    from oracle.roi_oracle import synthetic_atlas
    from oracle.unet_oracle import roi_features_oracle, unet3d_oracle

    torch.set_num_threads(max(1, os.cpu_count() or 1))
    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in unet3d.UNet3D(1, 1).state_dict().items()}
    lab = synthetic_atlas(SHAPE, N_ROIS)
    x = torch.rand((1, 1) + SHAPE)
    t0 = time.perf_counter()
    n = 0
    with torch.no_grad():
        while n < 1 or time.perf_counter() - t0 < seconds_cap / 2:
            hooked = {}
            unet3d_oracle(sd, x, False, hooked=hooked)
            roi_features_oracle(hooked["s_block1.conv2"], lab, N_ROIS)
            n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "volumes/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} single-volume pass(es) of the UNet3D forward (eval, fp32) + ROI means of the 64-channel map on the host cores"}


def unet_stock_torch_forward(m, x):
    """unet3d.py:137-157 through the nn modules' OWN forward (stock PyTorch / cuDNN): same module tree, nothing of this repo's
    CUDA library runs here.  Returns (out, the raw s_block1.conv2 output)."""
    import torch
    import torch.nn.functional as F

    d, h, w = x.shape[2:]
    td, th, tw = m.target
    cur = F.pad(x, (0, tw - w, 0, th - h, 0, td - d))
    skips = []
    for blk in (m.a_block1, m.a_block2, m.a_block3, m.bottleNeck):
        cur = blk.relu(blk.bn1(blk.conv1(cur)))
        cur = blk.relu(blk.bn2(blk.conv2(cur)))
        if not blk.bottleneck:
            skips.append(cur)
            cur = blk.pooling(cur)
    hook = None
    for blk in (m.s_block3, m.s_block2, m.s_block1):
        cur = torch.cat((blk.upconv1(cur), skips.pop()), 1)
        cur = blk.relu(blk.bn(blk.conv1(cur)))
        hook = blk.conv2(cur)
        cur = blk.relu(blk.bn(hook))
    return m.s_block1.conv3(cur)[:, :, :d, :h, :w], hook


def unet_torch_gpu_baseline(dev, batch=8, steps=3):
    """Same-box context for the extraction path: the same UNet3D module tree run by stock PyTorch + cuDNN in eval mode (bf16 autocast,
    channels_last_3d) + the ROI means of the 64-channel map as a torch index_add over labelled voxels (the reference's one-hot
    product would need 39 GB per volume).  Not the target and not the --impl reference arm."""
    import numpy as np
    import torch

    from multimodal_ad_b200.models import unet3d
    from oracle.roi_oracle import synthetic_atlas                       # label-map generator only

    out = {"batch": batch, "steps": steps, "what": "stock torch.nn modules (cuDNN), eval, bf16 autocast + channels_last_3d, + torch index_add ROI means"}
    try:
        torch.manual_seed(0)
        model = unet3d.UNet3D(1, 1).to(dev).eval().to(memory_format=torch.channels_last_3d)
        lab = torch.from_numpy(synthetic_atlas(SHAPE, N_ROIS).astype(np.int64)).to(dev).reshape(-1)
        idx = torch.nonzero(lab).squeeze(1)
        cnt = torch.bincount(lab, minlength=N_ROIS + 1)[1:].clamp_min(1).float()
        x = torch.rand((batch, 1) + SHAPE, device=dev)
        old = torch.backends.cudnn.benchmark
        torch.backends.cudnn.benchmark = True

        def step():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                o, hook = unet_stock_torch_forward(model, x)
                f = hook[..., :SHAPE[0], :SHAPE[1], :SHAPE[2]].float().reshape(batch, 64, -1)[:, :, idx]      # (B, 64, labelled voxels)
                acc = torch.zeros((batch, 64, N_ROIS + 1), device=dev).index_add_(2, lab[idx], f)
                return o, (acc[:, :, 1:] / cnt).permute(0, 2, 1)

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        out.update({"ms_per_step": ms, "volumes_per_sec": batch / (ms * 1e-3)})
        torch.backends.cudnn.benchmark = old
        del model
    except Exception as e:                                 # noqa: BLE001
        out["error"] = f"{type(e).__name__}: {e}"[:200]
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return out


def run_unet_roi_extract(args, rank, world, dev, dist, batch=8, steps=10, want_cpu=False):
    """image_features.py:97-114 on the accelerated path: UNet3D(1, 1).eval() forward on batch `batch` of 1x91x109x91 volumes, the
    64-channel s_block1.conv2 map pooled over a 170-label atlas without leaving the GPU; subject-sharded across ranks."""
    import torch

    from multimodal_ad_b200 import RoiPlan, _lib
    from multimodal_ad_b200.models import unet3d
    from multimodal_ad_b200.sharding import max_over_ranks
    from oracle.roi_oracle import synthetic_atlas                       # label-map generator only

    torch.manual_seed(0)
    model = unet3d.UNet3D(in_channels=1, num_classes=1).to(dev).eval()
    plan = RoiPlan(synthetic_atlas(SHAPE, N_ROIS), N_ROIS)
    g = torch.Generator(device=dev).manual_seed(300 + rank)
    xs = [torch.rand((batch, 1) + SHAPE, device=dev, generator=g) for _ in range(2)]

    def step(x):
        with torch.no_grad():
            return model.roi_features(x, plan)

    for i in range(3):
        step(xs[i % 2])
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    f0, l0 = _lib.executed_mma_flops(), _lib.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        step(xs[i % 2])
    b.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = max_over_ranks(a.elapsed_time(b), dev) / steps
    executed = (_lib.executed_mma_flops() - f0) / steps
    launches = _lib.launch_count() - l0
    # end to end: pinned host volumes in, network output + ROI features back on the host every step
    xh = torch.rand((batch, 1) + SHAPE).pin_memory()
    oh = torch.empty((batch, 1) + SHAPE).pin_memory()
    rh = torch.empty((batch, N_ROIS, 64)).pin_memory()
    xd = torch.empty((batch, 1) + SHAPE, device=dev)
    torch.cuda.synchronize()
    n_e2e = max(1, min(steps, 5))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        xd.copy_(xh, non_blocking=True)
        out, roi = step(xd)
        oh.copy_(out, non_blocking=True)
        rh.copy_(roi, non_blocking=True)
        torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0, dev)
    flops = unet3d_forward_flops() * batch
    peak, peak_src = tensor_peak()
    ach = flops / (ms * 1e-3) / 1e12
    res = {
        "metric": "unet3d_roi_extract_volumes_per_sec", "value": world * batch / (ms * 1e-3), "unit": "volumes/s", "ms_per_step": ms,
        "steps": steps, "dtype": "bf16", "scaling": "weak",
        "config": {"workload": "unet3d_eval_forward_plus_roi_mean_170labels_batch8_1x91x109x91 (image_features.py:97-114)",
                   "batch_per_gpu": batch, "volume": list(SHAPE), "padded_grid": [96, 112, 96],
                   "parallelism": f"subject-sharded x{world}, no collective", "launch": "eager"},
        "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_flops_per_step": flops, "executed_flops_per_step": executed,
                     "executed_frac": executed / (ms * 1e-3) / 1e12 / peak},
        "e2e": {"value": world * batch * n_e2e / dt, "unit": "volumes/s", "h2d_bytes_per_step": batch * SHAPE[0] * SHAPE[1] * SHAPE[2] * 4,
                "d2h_bytes_per_step": batch * (SHAPE[0] * SHAPE[1] * SHAPE[2] + N_ROIS * 64) * 4, "steps": n_e2e},
        "gpu_launches": launches,
    }
    if want_cpu:
        res["cpu_baseline"] = unet_cpu_reference()
    del model, xs, plan
    torch.cuda.empty_cache()
    if want_cpu:                                           # one GPU, rank 0: the stock PyTorch / cuDNN context number as well
        res["torch_gpu_baseline"] = unet_torch_gpu_baseline(dev, batch)
    torch.cuda.empty_cache()
    return res


def run_unet_train(args, rank, world, dev, dist, batch=4, steps=5):
    """UNet3D(1, 1) training step on batch `batch` of 1x91x109x91 volumes per GPU: forward + voxel-wise MSE loss + backward
    (+ one coalesced gradient all-reduce under data parallel) + Adam - the forward AND backward of unet3d.py's Conv3d + BatchNorm3d +
    ReLU stacks through the C-ABI kernels."""
    import torch

    from multimodal_ad_b200 import _lib
    from multimodal_ad_b200.models import unet3d
    from multimodal_ad_b200.sharding import max_over_ranks

    torch.manual_seed(0)
    model = unet3d.UNet3D(in_channels=1, num_classes=1).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-5, fused=True)
    g = torch.Generator(device=dev).manual_seed(400 + rank)
    x = torch.rand((batch, 1) + SHAPE, device=dev, generator=g)
    y = torch.rand((batch, 1) + SHAPE, device=dev, generator=g)
    params = [p for p in model.parameters()]

    def step():
        loss = ((model(x) - y) ** 2).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if dist is not None:
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        opt.step()
        return loss

    for _ in range(3):
        step()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    f0, l0 = _lib.executed_mma_flops(), _lib.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss = step()
    b.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = max_over_ranks(a.elapsed_time(b), dev) / steps
    executed = (_lib.executed_mma_flops() - f0) / steps
    flops = 3.0 * unet3d_forward_flops() * batch                          # forward + dgrad + wgrad of (nearly) every convolution
    peak, peak_src = tensor_peak()
    ach = flops / (ms * 1e-3) / 1e12
    res = {"metric": "unet3d_train_volumes_per_sec", "value": world * batch / (ms * 1e-3), "unit": "volumes/s", "ms_per_step": ms,
           "steps": steps, "dtype": "bf16", "scaling": "weak",
           "config": {"workload": f"unet3d_bf16_train_batch{batch}_1x91x109x91 (models/unet3d.py, padded grid 96x112x96)", "batch_per_gpu": batch,
                      "parallelism": f"data parallel x{world}" + (", one coalesced gradient all-reduce after backward" if world > 1 else ""),
                      "step": "forward + MSE loss + backward + Adam", "launch": "eager"},
           "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                        "peak_source": peak_src, "algorithmic_flops_per_step": flops, "executed_flops_per_step": executed,
                        "executed_frac": executed / (ms * 1e-3) / 1e12 / peak},
           "gpu_launches": _lib.launch_count() - l0, "loss": float(loss.detach())}
    del model, opt, x, y
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------------------------------------------------------
# arms
# ---------------------------------------------------------------------------------------------------------------------
def run_reference_arm(args, rank: int):
    """--impl reference: the reference's own CPU implementation of the headline path on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    r = resnet_cpu_reference(args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "resnet3d18_train_volumes_per_sec", "value": r["value"], "unit": "volumes/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": max(args.warmup, 1), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": RESNET_WORKLOAD, "batch_per_gpu": RESNET_BATCH, "volume": [RESNET_SIZE] * 3},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_extras:
        line["roi_pool"] = {"impl": "reference", "metric": "roi_pool_volumes_per_sec", "config": {"workload": ROI_WORKLOAD},
                            **roi_cpu_reference(5.0)}
    print(json.dumps(line), flush=True)


def run_cuda_arm(args, rank: int, local_rank: int, world: int):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    solo = rank == 0 and world == 1

    head = run_resnet_train(args, rank, world, dev, dist, steps=args.steps, warmup=args.warmup, sample_clocks=True, local_rank=local_rank)
    extras = {}

    def guarded(name, fn):
        """A nested workload must never take the headline line down on one GPU (under torchrun an exception on one rank would
        leave the others in a collective, so there it is allowed to propagate)."""
        if world > 1:
            extras[name] = fn()
            return
        try:
            extras[name] = fn()
        except Exception as e:                             # noqa: BLE001
            extras[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.synchronize()
            torch.cuda.empty_cache()

    if not args.no_extras:
        guarded("roi_pool", lambda: run_roi_pool(args, rank, world, dev, dist, max(args.steps, 20), max(args.warmup, 3),
                                                 want_cpu=solo and not args.no_cpu_baseline))
        guarded("resnet3d18_train_91x109x91", lambda: run_resnet_train(
            args, rank, world, dev, dist, shape=SHAPE, steps=min(args.steps, 10), warmup=3, do_e2e=False,
            workload="resnet3d18_bf16_train_batch16_1x91x109x91_3class (config/config_unet.json volume)",
            metric="resnet3d18_train_91x109x91_volumes_per_sec"))
        guarded("resnet3d50_train", lambda: run_resnet_train(
            args, rank, world, dev, dist, depth=50, batch=8, nb_class=2, steps=min(args.steps, 8), warmup=3, do_e2e=False,
            workload="resnet3d50_bottleneck_bf16_train_batch8_1x128^3 (cfg_denseNet.json: model_type resnet, depth 50)",
            metric="resnet3d50_train_volumes_per_sec"))
        guarded("unet3d_roi_extract", lambda: run_unet_roi_extract(args, rank, world, dev, dist, steps=min(args.steps, 10),
                                                                   want_cpu=solo and not args.no_cpu_baseline))
        guarded("unet3d_train", lambda: run_unet_train(args, rank, world, dev, dist, steps=min(args.steps, 5)))
        if solo:
            guarded("torch_gpu_baseline", lambda: torch_gpu_baseline(dev))
    cpu_baseline = None
    if solo and not args.no_cpu_baseline:
        r = resnet_cpu_reference(steps=64, warmup=1, seconds_cap=20.0)
        cpu_baseline = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        pipe = step_tensor_pipe_from_profile()
        if pipe is not None:
            head["roofline"]["tensor_pipe_active_pct_ncu"] = pipe
        line = {
            "metric": head["metric"], "value": head["value"], "unit": head["unit"], "n_gpus": world,
            "steps": head["steps"], "warmup": head["warmup"], "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": head["config"], "roofline": head["roofline"], "cpu_baseline": cpu_baseline, "e2e": head.get("e2e"),
            "gpu_launches": head["gpu_launches"], "clocks": head.get("clocks"), "loss": head["loss"],
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (no nested ROI / ResNet-50 / UNet / torch objects)")
    ap.add_argument("--no-graph", action="store_true", help="ResNet step: launch eagerly instead of replaying a CUDA graph")
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
