"""Atlas ROI pooling — host-side mirror of the reference's ROI feature path.

The reference computes ROI features inline (its models/ROI_pol.py is empty):
/root/reference/image_features.py:67-69 reads the atlas label volume,
:80-82 builds the one-hot mask and :111-114 does
``roi_feat = (feats[:,None] * onehot[None,:,None]).sum(spatial) / den`` giving
(B, R, C).  `ROIPool.forward` returns exactly that tensor; `ROIPool.pool`
adds the per-ROI max / argmax / counts the north star asks for.

All compute is in the CUDA library behind include/mmad_b200.h
(csrc/roi_pool.cu); a CPU tensor raises instead of falling back.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_int32, c_int64, c_void_p
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from .. import _lib


def _np_labels(labels) -> np.ndarray:
    if isinstance(labels, torch.Tensor):
        labels = labels.detach().cpu().numpy()
    lab = np.asarray(labels)
    if not np.issubdtype(lab.dtype, np.integer):
        lab = lab.astype(int)            # image_features.py:67 `.astype(int)`
    return np.ascontiguousarray(lab.reshape(-1), dtype=np.int32)


class RoiPlan:
    """One atlas, prepared for pooling (wraps `mmad_roi_plan`).

    labels : integer array (D, H, W) or flat, values in [0, n_rois], 0 = background.
    n_rois : number of ROI columns; default labels.max() (image_features.py:80-81).
    """

    def __init__(self, labels, n_rois: Optional[int] = None, *, tile: int = 256, stages: int = 0,
                 consumer_warps: int = 0, host_only: bool = False, device: Optional[torch.device] = None):
        lab = _np_labels(labels)
        if lab.size == 0:
            raise ValueError("empty label map")
        if n_rois is None:
            n_rois = int(lab.max())
        self.n_voxels = int(lab.size)
        self.n_rois = int(n_rois)
        self.tile = int(tile)
        self.host_only = bool(host_only)
        self.device = None
        self._h = c_void_p()
        lib = _lib.load()
        if host_only:
            rc = lib.mmad_roi_plan_create_ex(lab.ctypes.data, self.n_voxels, self.n_rois, tile, stages,
                                             consumer_warps, 1, byref(self._h))
        else:
            if not torch.cuda.is_available():
                raise _lib.MmadError("RoiPlan needs a CUDA device (no CPU fallback)")
            self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
            with torch.cuda.device(self.device):
                rc = lib.mmad_roi_plan_create_ex(lab.ctypes.data, self.n_voxels, self.n_rois, tile, stages,
                                                 consumer_warps, 0, byref(self._h))
        _lib.check(rc, "mmad_roi_plan_create")
        self._counts = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.load().mmad_roi_plan_destroy(h)
            except Exception:
                pass
            self._h = c_void_p()

    # -- plan facts -------------------------------------------------------------------
    @property
    def counts(self) -> np.ndarray:
        """Per-ROI voxel counts, computed on the GPU at plan creation (int32[R])."""
        if self._counts is None:
            out = np.empty(self.n_rois, np.int32)
            _lib.check(_lib.load().mmad_roi_plan_counts(self._h, out.ctypes.data), "mmad_roi_plan_counts")
            self._counts = out
        return self._counts

    def algorithmic_bytes(self, n_vols: int) -> int:
        return int(_lib.load().mmad_roi_pool_algorithmic_bytes(self._h, n_vols))

    def programme(self):
        """(words uint32[], offsets int32[n_tiles+1] in 16-byte units, stages, smem_bytes) — host-side, for tests."""
        lib = _lib.load()
        nw, nt, ns, sm, cw = c_int64(), c_int32(), c_int32(), c_int64(), c_int32()
        _lib.check(lib.mmad_roi_plan_programme(self._h, None, byref(nw), None, byref(nt), byref(ns), byref(sm),
                                               byref(cw)), "mmad_roi_plan_programme")
        self.consumer_warps = cw.value
        words = np.empty(nw.value, np.uint32)
        offs = np.empty(nt.value + 1, np.int32)
        _lib.check(lib.mmad_roi_plan_programme(self._h, words.ctypes.data, None, offs.ctypes.data, None, None, None,
                                               None), "mmad_roi_plan_programme")
        return words, offs, ns.value, sm.value

    def binding(self, n_vols: int, sms: int = 148) -> dict:
        """Work-item / partial-slot layout for n_vols volumes on `sms` SMs — host-side, for tests."""
        lib = _lib.load()
        ni, nsl, grid = c_int32(), c_int32(), c_int32()
        _lib.check(lib.mmad_roi_plan_binding(self._h, n_vols, sms, byref(ni), byref(nsl), byref(grid), None, None,
                                             None, None, None, None, None), "mmad_roi_plan_binding")
        n_groups = (n_vols + 31) // 32
        out = dict(n_items=ni.value, n_slots=nsl.value, grid=grid.value, n_groups=n_groups,
                   item_group=np.empty(ni.value, np.int32), item_t0=np.empty(ni.value, np.int32),
                   item_t1=np.empty(ni.value, np.int32), item_slot_ptr=np.empty(ni.value + 1, np.int32),
                   fin_ptr=np.empty(n_groups * self.n_rois + 1, np.int32))
        sl = np.empty(max(nsl.value, 1), np.uint8)
        sd = np.empty(max(nsl.value, 1), np.int32)
        _lib.check(lib.mmad_roi_plan_binding(self._h, n_vols, sms, None, None, None, out["item_group"].ctypes.data,
                                             out["item_t0"].ctypes.data, out["item_t1"].ctypes.data,
                                             out["item_slot_ptr"].ctypes.data, sl.ctypes.data, sd.ctypes.data,
                                             out["fin_ptr"].ctypes.data), "mmad_roi_plan_binding")
        out["slot_label"] = sl[:nsl.value]
        out["slot_dst"] = sd[:nsl.value]
        return out

    # -- compute ----------------------------------------------------------------------
    def _check_vols(self, vols: torch.Tensor) -> torch.Tensor:
        if self.host_only:
            raise _lib.MmadError("host-only plan cannot run kernels")
        if not isinstance(vols, torch.Tensor) or not vols.is_cuda:
            raise _lib.MmadError("ROI pooling runs on CUDA tensors only (no CPU fallback)")
        if vols.dtype != torch.float32:
            raise TypeError("ROI pooling expects float32 volumes")
        if vols.device != self.device:
            raise ValueError(f"volumes on {vols.device}, plan on {self.device}")
        v = vols.reshape(-1, self.n_voxels) if vols.numel() else vols.reshape(0, self.n_voxels)
        return v.contiguous()

    def pool_channels_last(self, feats: torch.Tensor, atlas_shape) -> torch.Tensor:
        """ROI means of a channels-last feature map that never left the GPU.

        feats: (N, Dp, Hp, Wp, 64) float32 CUDA, NDHWC contiguous - the fp32 side output the convolution epilogue writes for
        the layer image_features.py:58-60 hooks; atlas_shape (D, H, W) <= (Dp, Hp, Wp) is the plan's label grid (the crop of
        image_features.py:104 is folded into the kernel).  Returns roi_feat (N, R, 64) = image_features.py:114."""
        if self.host_only:
            raise _lib.MmadError("host-only plan cannot run kernels")
        if not isinstance(feats, torch.Tensor) or not feats.is_cuda:
            raise _lib.MmadError("ROI pooling runs on CUDA tensors only (no CPU fallback)")
        if feats.dtype != torch.float32 or feats.dim() != 5 or not feats.is_contiguous():
            raise TypeError("pool_channels_last expects a contiguous float32 (N, Dp, Hp, Wp, C) tensor")
        n, dp, hp, wp, c = feats.shape
        d, h, w = (int(t) for t in atlas_shape)
        out = torch.empty((n, self.n_rois, c), dtype=torch.float32, device=feats.device)
        with torch.cuda.device(feats.device):
            st = c_void_p(torch.cuda.current_stream(feats.device).cuda_stream)
            _lib.check(_lib.load().mmad_roi_pool_ndhwc_f32(self._h, c_void_p(feats.data_ptr()), n, dp, hp, wp, d, h, w, c,
                                                           c_void_p(out.data_ptr()), st), "mmad_roi_pool_ndhwc_f32")
        return out

    def pool(self, vols: torch.Tensor, want_max: bool = True):
        """vols (N, V) float32 CUDA -> mean (N, R), max (N, R), argmax (N, R) int32 (max/argmax None if not wanted)."""
        v = self._check_vols(vols)
        n = v.shape[0]
        mean = torch.empty((n, self.n_rois), dtype=torch.float32, device=v.device)
        mx = torch.empty((n, self.n_rois), dtype=torch.float32, device=v.device) if want_max else None
        arg = torch.empty((n, self.n_rois), dtype=torch.int32, device=v.device) if want_max else None
        if n:
            with torch.cuda.device(v.device):
                stream = torch.cuda.current_stream(v.device).cuda_stream
                rc = _lib.load().mmad_roi_pool_f32(self._h, v.data_ptr(), n, mean.data_ptr(),
                                                   mx.data_ptr() if want_max else None,
                                                   arg.data_ptr() if want_max else None, c_void_p(stream))
            _lib.check(rc, "mmad_roi_pool_f32")
        return mean, mx, arg

    def stream_only(self, vols: torch.Tensor) -> None:
        """Runs only the streaming kernel (no outputs) — for timing the dominant kernel alone."""
        v = self._check_vols(vols)
        with torch.cuda.device(v.device):
            stream = torch.cuda.current_stream(v.device).cuda_stream
            rc = _lib.load().mmad_roi_pool_f32(self._h, v.data_ptr(), v.shape[0], None, None, None, c_void_p(stream))
        _lib.check(rc, "mmad_roi_pool_f32")

    def mean_backward(self, grad_mean: torch.Tensor) -> torch.Tensor:
        g = grad_mean.contiguous().to(torch.float32)
        n = g.shape[0]
        out = torch.empty((n, self.n_voxels), dtype=torch.float32, device=g.device)
        if n:
            with torch.cuda.device(g.device):
                stream = torch.cuda.current_stream(g.device).cuda_stream
                rc = _lib.load().mmad_roi_pool_mean_backward_f32(self._h, g.data_ptr(), n, out.data_ptr(),
                                                                 c_void_p(stream))
            _lib.check(rc, "mmad_roi_pool_mean_backward_f32")
        return out

    def pool_host(self, vols):
        """HOST volumes (numpy array or CPU tensor, ideally pinned) -> numpy mean, max, argmax.
        The library chunks the host->device copy and overlaps it with the kernel."""
        if self.host_only:
            raise _lib.MmadError("host-only plan cannot run kernels")
        if isinstance(vols, torch.Tensor):
            if vols.is_cuda:
                raise ValueError("pool_host takes host memory")
            t = vols.reshape(-1, self.n_voxels).contiguous()
            ptr, n = t.data_ptr(), t.shape[0]
            if t.dtype != torch.float32:
                raise TypeError("float32 expected")
        else:
            t = np.ascontiguousarray(np.asarray(vols, np.float32).reshape(-1, self.n_voxels))
            ptr, n = t.ctypes.data, t.shape[0]
        mean = np.empty((n, self.n_rois), np.float32)
        mx = np.empty((n, self.n_rois), np.float32)
        arg = np.empty((n, self.n_rois), np.int32)
        with torch.cuda.device(self.device):
            rc = _lib.load().mmad_roi_pool_host_f32(self._h, ptr, n, mean.ctypes.data, mx.ctypes.data, arg.ctypes.data)
        _lib.check(rc, "mmad_roi_pool_host_f32")
        return mean, mx, arg


class _RoiMeanFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats2d, plan):
        ctx.plan = plan
        mean, _, _ = plan.pool(feats2d, want_max=False)
        return mean

    @staticmethod
    def backward(ctx, grad_mean):
        return ctx.plan.mean_backward(grad_mean), None


class ROIPool(nn.Module):
    """nn.Module form of image_features.py:62-114.

    >>> pool = ROIPool(aal_data)                 # (D,H,W) integer atlas, 0 = background
    >>> roi_feat = pool(feats64)                 # (B,C,D,H,W) float32 CUDA -> (B,R,C), image_features.py:114
    >>> out = pool.pool(feats64)                 # dict(mean, max, argmax (B,R,C); counts (R,))

    The feature grid may be LARGER than the atlas (96x112x96 UNet features against the 91x109x91 AAL grid): the
    reference's crop `feats64[..., :D, :H, :W]` (image_features.py:103-105) is then folded into the pooling plan.
    """

    def __init__(self, atlas_labels, num_rois: Optional[int] = None, *, tile: int = 256):
        super().__init__()
        lab = _np_labels(atlas_labels)
        self.spatial_shape = tuple(np.asarray(atlas_labels.shape if hasattr(atlas_labels, "shape") else lab.shape))
        self.num_rois = int(lab.max()) if num_rois is None else int(num_rois)
        self.tile = tile
        self.register_buffer("atlas", torch.from_numpy(lab.astype(np.int32)), persistent=True)
        self._plans = {}

    def _plan(self, device: torch.device, grid=None) -> RoiPlan:
        grid = self.spatial_shape if grid is None else tuple(grid)
        key = (device.type, device.index, grid)
        pl = self._plans.get(key)
        if pl is None:
            lab = self.atlas.cpu().numpy().reshape(self.spatial_shape)
            if grid != self.spatial_shape:
                # the feature map is larger than the atlas (the reference's UNet pads 91x109x91 to 96x112x96 and crops the
                # features back, image_features.py:103-105): pool the UNCROPPED map against the atlas embedded in that grid
                # with background labels - the same sums over the same voxels, without the 64-channel crop copy
                big = np.zeros(grid, dtype=lab.dtype)
                big[: lab.shape[0], : lab.shape[1], : lab.shape[2]] = lab
                lab = big
            pl = RoiPlan(lab, self.num_rois, tile=self.tile, device=device)
            self._plans[key] = pl
        return pl

    def _prep(self, feats: torch.Tensor):
        if feats.dim() != 5:
            raise ValueError("expected (B, C, D, H, W)")
        if not feats.is_cuda:
            raise _lib.MmadError("ROIPool runs on CUDA tensors only (no CPU fallback)")
        b, c = feats.shape[:2]
        grid = tuple(int(v) for v in feats.shape[2:])
        if len(self.spatial_shape) != 3 or any(g < a for g, a in zip(grid, self.spatial_shape)):
            raise ValueError(f"feature grid {grid} does not cover atlas {self.spatial_shape}")
        return b, c, self._plan(feats.device, grid), grid

    def forward(self, feats: torch.Tensor) -> torch.Tensor:
        b, c, plan, _ = self._prep(feats)
        mean = _RoiMeanFunction.apply(feats.reshape(b * c, -1), plan)        # (B*C, R)
        return mean.reshape(b, c, self.num_rois).permute(0, 2, 1)            # (B, R, C)  image_features.py:114

    @torch.no_grad()
    def pool(self, feats: torch.Tensor) -> dict:
        b, c, plan, grid = self._prep(feats)
        mean, mx, arg = plan.pool(feats.reshape(b * c, -1), want_max=True)
        if grid != self.spatial_shape:                                       # flat index in the padded grid -> in the atlas grid
            d_, h_, w_ = self.spatial_shape
            a = arg.long()
            dd, hh, ww = a // (grid[1] * grid[2]), (a // grid[2]) % grid[1], a % grid[2]
            arg = torch.where(arg >= 0, (dd * h_ + hh) * w_ + ww, a).to(arg.dtype)
        shp = lambda t: t.reshape(b, c, self.num_rois).permute(0, 2, 1)      # noqa: E731
        return dict(mean=shp(mean), max=shp(mx), argmax=shp(arg),
                    counts=torch.from_numpy(plan.counts.copy()))


def roi_pool(feats: torch.Tensor, atlas_labels, num_rois: Optional[int] = None) -> torch.Tensor:
    """Functional one-shot form: (B,C,D,H,W) x (D,H,W) labels -> (B,R,C) ROI means."""
    return ROIPool(atlas_labels, num_rois).to(feats.device)(feats)
