"""3-D ResNet image branch — drop-in for /root/reference/models/resnet.py.

Same names (`conv3x3x3`, `BasicBlock`, `Bottleneck`, `ResNet`, `resnet10` … `resnet200`), same constructor
arguments, same parameter / buffer names (so `state_dict()`s are interchangeable, resnet.py:112-202), same
initialisation (resnet.py:171-176).  `ResNet.forward` (resnet.py:204-215) runs the backbone
(conv1 → bn1 → relu → maxpool → layer1..4) on the CUDA library behind include/mmad_b200.h — implicit-GEMM Conv3d on
tcgen05/TMEM (csrc/conv3d_igemm.cu, csrc/conv3d_wgrad.cu) and the bandwidth kernels of csrc/nn_kernels.cu — inside ONE
`torch.autograd.Function`; `conv_seg` (the head the training scripts replace, train_ResNet3D.py:66-71) stays a normal
torch module applied to the backbone's output.

The nn.Conv3d / nn.BatchNorm3d objects are parameter containers only: their own forward is never called.  Activations
live as NDHWC bf16 between kernels; convolutions accumulate in fp32.  There is no CPU / cuDNN fallback: a CPU tensor raises.
BasicBlock and Bottleneck networks, shortcut types 'A' (detached, as in the reference) and 'B' are covered.
"""
from __future__ import annotations

import ctypes
from ctypes import c_void_p
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib

__all__ = ['ResNet', 'resnet10', 'resnet18', 'resnet34', 'resnet50', 'resnet101', 'resnet152', 'resnet200']


def conv3x3x3(in_planes, out_planes, stride=1, dilation=1):
    # resnet.py:14-23
    return nn.Conv3d(in_planes, out_planes, kernel_size=3, dilation=dilation, stride=stride, padding=dilation, bias=False)


def downsample_basic_block(x, planes, stride, no_cuda=False):
    # resnet.py:26-37 (shortcut type 'A'); kept for API parity, not on the accelerated path
    out = F.avg_pool3d(x, kernel_size=1, stride=stride)
    zero_pads = torch.zeros(out.size(0), planes - out.size(1), out.size(2), out.size(3), out.size(4),
                            dtype=out.dtype, device=out.device)
    return torch.cat([out, zero_pads], dim=1)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        self.conv1 = conv3x3x3(inplanes, planes, stride=stride, dilation=dilation)
        self.bn1 = nn.BatchNorm3d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3x3(planes, planes, dilation=dilation)
        self.bn2 = nn.BatchNorm3d(planes)
        self.downsample = downsample
        self.stride = stride
        self.dilation = dilation


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv3d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm3d(planes)
        self.conv2 = nn.Conv3d(planes, planes, kernel_size=3, stride=stride, dilation=dilation, padding=dilation, bias=False)
        self.bn2 = nn.BatchNorm3d(planes)
        self.conv3 = nn.Conv3d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm3d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride
        self.dilation = dilation


# ----------------------------------------------------------------------------------------------------------------
# kernel plumbing
# ----------------------------------------------------------------------------------------------------------------
def _p(t):
    return c_void_p(t.data_ptr()) if t is not None else None


_SIDE_STREAMS = {}


def _side_stream(device) -> torch.cuda.Stream:
    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


class _Run:
    """One forward or backward pass: library handle, stream, allocation helpers, thin kernel wrappers."""

    def __init__(self, device, use_side_stream=False):
        self.lib = _lib.load()
        self.dev = device
        self.main = torch.cuda.current_stream(device)
        self.stream = c_void_p(self.main.cuda_stream)
        # weight-gradient kernels are off the backward critical path: they go to a side stream so the HBM-bound
        # BatchNorm / ReLU kernels of the main chain run underneath the tensor-bound wgrad work
        self.side = _side_stream(device) if use_side_stream else None
        self._keep = []                                    # tensors in use on the side stream, released by join_side()
        self.tracked = []                                  # num_batches_tracked buffers of the BatchNorms this pass has updated

    def bump_tracked(self):
        if self.tracked:
            uniq = {}
            for t in self.tracked:                         # a BatchNorm shared by two convolutions (unet3d.py:69) counts twice
                uniq[id(t)] = (t, uniq.get(id(t), (t, 0))[1] + 1)
            torch._foreach_add_([t for t, _ in uniq.values()], [c for _, c in uniq.values()])
            self.tracked = []

    def empty(self, shape, dtype=torch.bfloat16):
        return torch.empty(shape, dtype=dtype, device=self.dev)

    def chk(self, rc, what):
        _lib.check(rc, what)

    # y = conv(x, w_fwd) [+ per-CTA BN statistic partials]
    def conv(self, x, w_fwd, cout, k, stride, pad, dil, want_stats):
        n, d, h, w, cin = x.shape
        do = (d + 2 * pad - dil * (k - 1) - 1) // stride + 1
        ho = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
        wo = (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
        y = self.empty((n, do, ho, wo, cout))
        part = None
        if want_stats:
            npart = self.lib.mmad_conv3d_stats_partials(n, d, h, w, cout, k, stride, pad, dil)
            part = self.empty((npart, cout, 2), torch.float32)
        self.chk(self.lib.mmad_conv3d_fwd_bf16(_p(x), _p(w_fwd), _p(y), _p(part), n, d, h, w, cin, cout, k, stride, pad, dil,
                                               self.stream), "mmad_conv3d_fwd_bf16")
        return y, part

    def stem_wgrad(self, xs, dy, n, d, h, w, out_dw):
        """conv1 weight gradient from the space-to-depth input (include/mmad_b200.h: mmad_stem_s2d_wgrad); side stream like wgrad."""
        nsplit = ctypes.c_int(0)
        elems = self.lib.mmad_stem_s2d_wgrad_workspace(n, d, h, w, ctypes.byref(nsplit))
        if elems < 0:
            raise _lib.MmadError("mmad_stem_s2d_wgrad_workspace: bad geometry")
        ws = self.empty((elems,), torch.float32)
        cs = self.stream
        if self.side is not None:
            self.side.wait_stream(self.main)
            cs = c_void_p(self.side.cuda_stream)
            self._keep += [xs, dy, ws]                     # see wgrad()
        self.chk(self.lib.mmad_stem_s2d_wgrad(_p(xs), _p(dy), _p(ws), n, d, h, w, cs), "mmad_stem_s2d_wgrad")
        self.chk(self.lib.mmad_stem_s2d_wgrad_reduce(_p(ws), nsplit.value, _p(out_dw), cs), "mmad_stem_s2d_wgrad_reduce")

    def wgrad(self, x, dy, cout, k, stride, pad, dil, out_dw, on_done=None):
        """out_dw: fp32 tensor in torch layout (Cout, Cin, k,k,k) (or any tensor of Cout*Cin*taps elements).
        Runs on the side stream when the run has one; on_done() is then called with that stream current (used to start
        the data-parallel all-reduce of this gradient behind its own kernels)."""
        n, d, h, w, cin = x.shape
        nsplit = ctypes.c_int(0)
        elems = self.lib.mmad_conv3d_wgrad_workspace(n, d, h, w, cin, cout, k, stride, pad, dil, ctypes.byref(nsplit))
        if elems < 0:
            raise _lib.MmadError("mmad_conv3d_wgrad_workspace: bad geometry")
        ws = self.empty((elems,), torch.float32)
        stream, cs = self.main, self.stream
        if self.side is not None:
            self.side.wait_stream(self.main)               # x and dy were produced on the main stream
            stream, cs = self.side, c_void_p(self.side.cuda_stream)
            # the operands belong to the main stream's allocator pool: hold them until join_side() instead of
            # record_stream() (recorded blocks cannot be reused until an event query succeeds, which made the allocator
            # fall back to cudaMalloc in the middle of the step)
            self._keep += [x, dy, ws]
        self.chk(self.lib.mmad_conv3d_wgrad_bf16(_p(x), _p(dy), _p(ws), n, d, h, w, cin, cout, k, stride, pad, dil, cs),
                 "mmad_conv3d_wgrad_bf16")
        self.chk(self.lib.mmad_wgrad_reduce(_p(ws), nsplit.value, _p(out_dw), cout, cin, k * k * k, cs), "mmad_wgrad_reduce")
        if on_done is not None:
            with torch.cuda.stream(stream):
                on_done()

    def join_side(self):
        if self.side is not None:
            self.main.wait_stream(self.side)
            self._keep.clear()                             # later main-stream work is ordered after the side stream's kernels

    def prep_weights(self, conv: nn.Conv3d, want_dgrad):
        cout, cin, k = conv.out_channels, conv.in_channels, conv.kernel_size[0]
        taps = k * k * k
        wf = self.empty((cout, taps, cin))
        wt = self.empty((cin, taps, cout)) if want_dgrad else None
        self.chk(self.lib.mmad_conv3d_prep_weights(_p(conv.weight.detach().contiguous()), _p(wf), _p(wt), cout, cin, taps, self.stream),
                 "mmad_conv3d_prep_weights")
        return wf, wt

    def prep_all(self, items):
        """items: [(key, fp32 weight (Cout, Cin, k,k,k), want_dgrad)] -> {key: (w_fwd [Cout][taps][Cin] bf16, w_dgrad [Cin][taps
        reversed][Cout] bf16 or None)} with TWO launches for all layers (mmad_conv3d_prep_weights_batched) instead of two per layer."""
        n = len(items)
        if n == 0:
            return {}
        ws = [w.detach().contiguous() for _, w, _ in items]
        sizes = [w.numel() for w in ws]
        fwd = self.empty((sum(sizes),))
        dgr = self.empty((sum(sz for sz, (_, _, wd) in zip(sizes, items) if wd),)) if any(wd for _, _, wd in items) else None
        P = ctypes.c_void_p * n
        I = ctypes.c_int * n
        pw, pf, pd, co, ci, tp = P(), P(), P(), I(), I(), I()
        out, of, od = {}, 0, 0
        for i, ((key, _, wd), w, sz) in enumerate(zip(items, ws, sizes)):
            cout, cin = w.shape[0], w.shape[1]
            taps = sz // (cout * cin)
            wf = fwd[of:of + sz].view(cout, taps, cin)
            of += sz
            wt = None
            if wd:
                wt = dgr[od:od + sz].view(cin, taps, cout)
                od += sz
            pw[i], pf[i], pd[i] = w.data_ptr(), wf.data_ptr(), (wt.data_ptr() if wt is not None else None)
            co[i], ci[i], tp[i] = cout, cin, taps
            out[key] = (wf, wt)
        self.chk(self.lib.mmad_conv3d_prep_weights_batched(n, pw, pf, pd, co, ci, tp, self.stream), "mmad_conv3d_prep_weights_batched")
        self._prep_keep = ws                               # the fp32 sources stay alive until the launches are enqueued (they are)
        return out

    def bn_params(self, bn: nn.BatchNorm3d, part, count, training):
        c = bn.num_features
        vec = self.empty((4, c), torch.float32)          # mean, invstd, scale, shift
        g, b = bn.weight.detach(), bn.bias.detach()
        if training:
            momentum = 0.1 if bn.momentum is None else bn.momentum
            track = bn.track_running_stats and bn.running_mean is not None
            self.chk(self.lib.mmad_bn_finalize(_p(part), part.shape[0], c, float(count), _p(g), _p(b), bn.eps, momentum,
                                               _p(bn.running_mean) if track else None, _p(bn.running_var) if track else None,
                                               _p(vec[0]), _p(vec[1]), _p(vec[2]), _p(vec[3]), self.stream), "mmad_bn_finalize")
            if track and bn.num_batches_tracked is not None:
                self.tracked.append(bn.num_batches_tracked)   # bumped together at the end of the pass (one launch, not one per layer)
        else:
            self.chk(self.lib.mmad_bn_eval_params(c, _p(g), _p(b), _p(bn.running_mean), _p(bn.running_var), bn.eps,
                                                  _p(vec[0]), _p(vec[1]), _p(vec[2]), _p(vec[3]), self.stream), "mmad_bn_eval_params")
        return vec

    def bn_apply(self, x, vec, relu, res=None, res_vec=None, also_f32=False):
        """-> bf16 activation (and, with also_f32, the same values in fp32 for the caller-facing output)."""
        rows, c = x.numel() // x.shape[-1], x.shape[-1]
        out = self.empty(x.shape)
        out32 = self.empty(x.shape, torch.float32) if also_f32 else None
        self.chk(self.lib.mmad_bn_apply(_p(x), _p(vec[2]), _p(vec[3]), _p(res), _p(res_vec[2]) if res_vec is not None else None,
                                        _p(res_vec[3]) if res_vec is not None else None, 1 if relu else 0,
                                        _p(out), _p(out32), rows, c, self.stream), "mmad_bn_apply")
        return (out, out32) if also_f32 else out

    def bn_bwd(self, dy, dy2, mask, x, vec, gamma, training, dy_is_f32=False, want_g=True, mask_from_x=False):
        """-> dx (bf16), g (bf16 or None), dgamma, dbeta (fp32).  mask_from_x: ReLU mask = relu(bn(x)) > 0, recomputed."""
        rows, c = x.numel() // x.shape[-1], x.shape[-1]
        npart = self.lib.mmad_bn_bwd_partials(rows)
        part = self.empty((npart, c, 2), torch.float32)
        g = self.empty(x.shape) if want_g else None
        self.chk(self.lib.mmad_bn_bwd_reduce(None if dy_is_f32 else _p(dy), _p(dy) if dy_is_f32 else None, _p(dy2), _p(mask), _p(x),
                                             _p(vec[0]), _p(vec[1]), _p(vec[2]) if mask_from_x else None,
                                             _p(vec[3]) if mask_from_x else None, _p(g), _p(part), rows, c, self.stream),
                 "mmad_bn_bwd_reduce")
        out = self.empty((2, c), torch.float32)           # dgamma, dbeta
        coef = self.empty((3, c), torch.float32)
        self.chk(self.lib.mmad_bn_bwd_finalize(_p(part), npart, c, float(rows), _p(gamma), _p(vec[0]), _p(vec[1]), 1 if training else 0,
                                               _p(out[0]), _p(out[1]), _p(coef), self.stream), "mmad_bn_bwd_finalize")
        dx = self.empty(x.shape)
        if want_g:
            self.chk(self.lib.mmad_bn_bwd_apply(_p(g), _p(x), _p(coef), _p(dx), rows, c, self.stream), "mmad_bn_bwd_apply")
        else:
            # no masked gradient was written: pass 2 reads the upstream gradient itself and, when the layer has a ReLU whose mask
            # comes from x, recomputes that mask (four tensor passes per BatchNorm backward instead of six)
            if dy_is_f32 or dy2 is not None or mask is not None:
                raise _lib.MmadError("bn_bwd(want_g=False) takes one bf16 upstream gradient and no stored mask")
            self.chk(self.lib.mmad_bn_bwd_apply_ex(_p(dy), _p(x), _p(coef), _p(vec[2]) if mask_from_x else None,
                                                   _p(vec[3]) if mask_from_x else None, _p(dx), rows, c, self.stream), "mmad_bn_bwd_apply_ex")
        return dx, g, out[0], out[1]


def _backbone_forward(model: "ResNet", x: torch.Tensor, training: bool, need_grad: bool):
    """Returns (features fp32 (N,D',H',W',C) NDHWC, tape for the backward pass)."""
    r = _Run(x.device)
    lib = r.lib
    n, cin, d, h, w = x.shape
    if cin != 1:
        raise _lib.MmadError("the accelerated stem expects 1 input channel (resnet.py:126-132)")
    x = x.contiguous().float()
    tape = {"blocks": [], "training": training}

    # ---- stem: conv1 7x7x7 s2 p3 as a space-to-depth implicit GEMM (no im2col matrix), bn1, relu, maxpool
    #      (resnet.py:205-208) ----
    k, s, p = 7, 2, 3
    do, ho, wo = (d + 2 * p - k) // s + 1, (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    rows = n * do * ho * wo
    xs_elems = lib.mmad_stem_s2d_elems(n, d, h, w)
    if xs_elems <= 0:
        raise _lib.MmadError(f"mmad_stem_s2d_elems: unsupported input geometry {(n, d, h, w)}")
    xs = r.empty((xs_elems,))
    r.chk(lib.mmad_stem_s2d_pack(_p(x), _p(xs), n, d, h, w, r.stream), "mmad_stem_s2d_pack")
    wstem = r.empty((64, 512))
    r.chk(lib.mmad_stem_s2d_prep_weights(_p(model.conv1.weight.detach().contiguous()), _p(wstem), r.stream), "mmad_stem_s2d_prep_weights")
    c0 = r.empty((n, do, ho, wo, 64))
    part = r.empty((lib.mmad_stem_s2d_stats_partials(n, d, h, w), 64, 2), torch.float32) if training else None
    r.chk(lib.mmad_stem_s2d_fwd(_p(xs), _p(wstem), _p(c0), _p(part), n, d, h, w, r.stream), "mmad_stem_s2d_fwd")
    v0 = r.bn_params(model.bn1, part, rows, training)
    pd, ph, pw = (do - 1) // 2 + 1, (ho - 1) // 2 + 1, (wo - 1) // 2 + 1
    p0 = r.empty((n, pd, ph, pw, 64))
    idx0 = torch.empty((n, pd, ph, pw, 64), dtype=torch.uint8, device=x.device)
    # bn1 + relu + maxpool fused: the 64-channel full-resolution activation is never written
    r.chk(lib.mmad_stem_bn_relu_maxpool_fwd(_p(c0), _p(v0[2]), _p(v0[3]), _p(p0), _p(idx0), n, do, ho, wo, 64, r.stream),
          "mmad_stem_bn_relu_maxpool_fwd")
    tape["stem"] = dict(xs=xs if need_grad else None, c0=c0, v0=v0, p0=p0, idx0=idx0, in_shape=(n, d, h, w))

    # ---- residual stages (resnet.py:209-212) ----
    cur = p0
    blocks = [b for layer in (model.layer1, model.layer2, model.layer3, model.layer4) for b in layer]
    items = []
    for blk in blocks:                                     # every convolution's weights re-laid in two launches
        if isinstance(blk, Bottleneck):
            items += [(blk.conv1, blk.conv1.weight, need_grad), (blk.conv2, blk.conv2.weight, need_grad and blk.conv2.stride[0] == 1),
                      (blk.conv3, blk.conv3.weight, need_grad)]
        else:
            items += [(blk.conv1, blk.conv1.weight, need_grad), (blk.conv2, blk.conv2.weight, need_grad)]
        if isinstance(blk.downsample, nn.Module):
            items.append((blk.downsample[0], blk.downsample[0].weight, need_grad))
    prepped = r.prep_all(items)
    for bi, blk in enumerate(blocks):
        last = bi == len(blocks) - 1
        if isinstance(blk, Bottleneck):
            # resnet.py:89-109: 1x1x1 reduce, 3x3x3 (stride / dilation), 1x1x1 expand (x4), residual
            st, dil = blk.conv2.stride[0], blk.conv2.dilation[0]
            planes, outc = blk.conv1.out_channels, blk.conv3.out_channels
            (w1f, w1t), (w2f, w2t), (w3f, w3t) = prepped[blk.conv1], prepped[blk.conv2], prepped[blk.conv3]
            c1, part1 = r.conv(cur, w1f, planes, 1, 1, 0, 1, training)
            v1 = r.bn_params(blk.bn1, part1, c1.numel() // planes, training)
            a1 = r.bn_apply(c1, v1, relu=True)
            c2, part2 = r.conv(a1, w2f, planes, 3, st, dil, dil, training)
            cnt = c2.numel() // planes
            v2 = r.bn_params(blk.bn2, part2, cnt, training)
            a2 = r.bn_apply(c2, v2, relu=True)
            c3, part3 = r.conv(a2, w3f, outc, 1, 1, 0, 1, training)
            v3 = r.bn_params(blk.bn3, part3, cnt, training)
            rec = dict(blk=blk, xin=cur, c1=c1, v1=v1, a1=a1, c2=c2, v2=v2, a2=a2, c3=c3, v3=v3, w1t=w1t, w2t=w2t, w3t=w3t,
                       stride=st, dil=dil)
            pre, vpre, width = c3, v3, outc
        else:
            st, dil = blk.conv1.stride[0], blk.conv1.dilation[0]
            planes = blk.conv1.out_channels
            (w1f, w1t), (w2f, w2t) = prepped[blk.conv1], prepped[blk.conv2]
            c1, part1 = r.conv(cur, w1f, planes, 3, st, dil, dil, training)
            cnt = c1.numel() // planes
            v1 = r.bn_params(blk.bn1, part1, cnt, training)
            a1 = r.bn_apply(c1, v1, relu=True)
            c2, part2 = r.conv(a1, w2f, planes, 3, 1, dil, dil, training)
            v2 = r.bn_params(blk.bn2, part2, cnt, training)
            rec = dict(blk=blk, xin=cur, c1=c1, v1=v1, a1=a1, c2=c2, v2=v2, w1t=w1t, w2t=w2t, stride=st, dil=dil)
            pre, vpre, width = c2, v2, planes
        if blk.downsample is not None and not isinstance(blk.downsample, nn.Module):
            # shortcut 'A' (resnet.py:26-37): subsample and zero-pad the channels - no parameters, plain tensor plumbing
            ra = r.empty(pre.shape)
            ra.zero_()
            ra[..., : cur.shape[-1]] = cur[:, ::st, ::st, ::st, :]
            out = r.bn_apply(pre, vpre, relu=True, res=ra, also_f32=last)
            rec.update(short_a=True)
        elif blk.downsample is not None:
            dconv, dbn = blk.downsample[0], blk.downsample[1]
            wdf, wdt = prepped[dconv]
            cd, partd = r.conv(cur, wdf, width, 1, dconv.stride[0], 0, 1, training)
            vd = r.bn_params(dbn, partd, cnt, training)
            out = r.bn_apply(pre, vpre, relu=True, res=cd, res_vec=vd, also_f32=last)
            rec.update(cd=cd, vd=vd, wdt=wdt)
        else:
            out = r.bn_apply(pre, vpre, relu=True, res=cur, also_f32=last)
        out32 = None
        if last:
            out, out32 = out
        rec["out"] = out                                   # bf16: next block's input and the ReLU mask of the backward
        tape["blocks"].append(rec)
        cur = out
    r.bump_tracked()
    return out32, tape


class _GradDict(dict):
    """{parameter: gradient}; hands each gradient to the model's GradReducer (if any) the moment it is enqueued, so
    the data-parallel all-reduce overlaps the rest of the backward pass."""

    def __init__(self, reducer, wanted=None):
        super().__init__()
        self.reducer = reducer
        self._wanted = wanted                              # ids of the parameters that require a gradient (None: all)

    def wanted(self, param) -> bool:
        return self._wanted is None or id(param) in self._wanted

    def __setitem__(self, k, v):
        if not self.wanted(k):
            return                                         # frozen parameter: no .grad, no all-reduce
        super().__setitem__(k, v)
        if self.reducer is not None:
            self.reducer.push(v)

    def set_quiet(self, k, v):
        """store without starting the all-reduce (the caller pushes it from the stream that produces the tensor)"""
        super().__setitem__(k, v)

    def pusher(self, v):
        return (lambda: self.reducer.push(v)) if self.reducer is not None else None

    def new_grad(self, param):
        """storage for a weight gradient: a view of the reducer's current bucket when data-parallel, else a fresh tensor"""
        return self.reducer.alloc_like(param) if self.reducer is not None else torch.empty_like(param)


def _backbone_backward(model: "ResNet", tape, grad_out: torch.Tensor, need_input_grad=False, wanted=None):
    """grad_out: gradient w.r.t. the (N,C,D',H',W')-shaped output view.  Returns {parameter: gradient}."""
    r = _Run(grad_out.device, use_side_stream=getattr(model, "wgrad_side_stream", True))
    lib = r.lib
    training = tape["training"]
    grads = _GradDict(getattr(model, "grad_reducer", None), wanted)
    last = tape["blocks"][-1]
    n, do, ho, wo, c = last["out"].shape
    # gradient of the NDHWC fp32 output; accept either memory order of the NCDHW-shaped gradient
    g_ndhwc = grad_out.permute(0, 2, 3, 4, 1)
    if g_ndhwc.is_contiguous() and grad_out.dtype == torch.float32:
        dy, dy_f32 = g_ndhwc, True
    else:
        src = grad_out.contiguous().float()
        dy = r.empty((n, do, ho, wo, c))
        r.chk(lib.mmad_ncs_f32_to_nsc_bf16(_p(src), _p(dy), n, c, do * ho * wo, r.stream), "mmad_ncs_f32_to_nsc_bf16")
        dy_f32 = False
    dy2 = None

    def wgrad_of(weight, x, dyc, cout, k, st, pad, dil):
        """weight gradient of one convolution, skipped (no kernels, no all-reduce bucket space) for a frozen parameter"""
        if not grads.wanted(weight):
            return
        gw = grads.new_grad(weight)
        r.wgrad(x, dyc, cout, k, st, pad, dil, gw, grads.pusher(gw))
        grads.set_quiet(weight, gw)

    def dgrad_3x3(dc, weight_t, conv, xin_shape, cin, cout, st, dil):
        """data gradient of a 3x3x3 convolution `conv` (cin -> cout): dc (N,Do,Ho,Wo,cout) -> (N,D,H,W,cin)"""
        if st == 1:
            return r.conv(dc, weight_t, cin, 3, 1, dil, dil, False)[0]
        if st == 2 and dil == 1:
            # stride 2: eight interleaved phase convolutions of dc (no zero insertion, 1/8 of the MACs)
            wph = r.empty((27 * cin * cout,))
            r.chk(lib.mmad_conv3d_prep_weights_s2(_p(conv.weight.detach().contiguous()), _p(wph), cout, cin, r.stream), "mmad_conv3d_prep_weights_s2")
            dx = r.empty(tuple(xin_shape[:4]) + (cin,))
            r.chk(lib.mmad_conv3d_dgrad_s2_bf16(_p(dc), _p(wph), _p(dx), xin_shape[0], xin_shape[1], xin_shape[2], xin_shape[3], cin, cout,
                                                r.stream), "mmad_conv3d_dgrad_s2_bf16")
            return dx
        if weight_t is None:
            weight_t = r.prep_weights(conv, True)[1]
        up = r.empty(tuple(xin_shape[:4]) + (cout,))
        r.chk(lib.mmad_upsample_zero2(_p(dc), _p(up), xin_shape[0], dc.shape[1], dc.shape[2], dc.shape[3], xin_shape[1], xin_shape[2],
                                      xin_shape[3], cout, r.stream), "mmad_upsample_zero2")
        return r.conv(up, weight_t, cin, 3, 1, 2 * dil - dil, dil, False)[0]      # pad' = dil*(k-1) - pad

    for rec in reversed(tape["blocks"]):
        blk, st, dil = rec["blk"], rec["stride"], rec["dil"]
        planes = blk.conv1.out_channels
        inpl = blk.conv1.in_channels
        xin = rec["xin"]
        if isinstance(blk, Bottleneck):
            outc = blk.conv3.out_channels
            # out = relu(bn3(c3) + res): g3 = dy * (out > 0)
            dc3, g2, dg, db = r.bn_bwd(dy, dy2, rec["out"], rec["c3"], rec["v3"], blk.bn3.weight.detach(), training, dy_is_f32=dy_f32)
            dy_f32 = False
            grads[blk.bn3.weight], grads[blk.bn3.bias] = dg, db
            wgrad_of(blk.conv3.weight, rec["a2"], dc3, outc, 1, 1, 0, 1)
            da2, _ = r.conv(dc3, rec["w3t"], planes, 1, 1, 0, 1, False)
            dc2, _, dg, db = r.bn_bwd(da2, None, None, rec["c2"], rec["v2"], blk.bn2.weight.detach(), training, want_g=False, mask_from_x=True)
            grads[blk.bn2.weight], grads[blk.bn2.bias] = dg, db
            wgrad_of(blk.conv2.weight, rec["a1"], dc2, planes, 3, st, dil, dil)
            da1 = dgrad_3x3(dc2, rec["w2t"], blk.conv2, rec["a1"].shape, planes, planes, st, dil)
            dc1, _, dg, db = r.bn_bwd(da1, None, None, rec["c1"], rec["v1"], blk.bn1.weight.detach(), training, want_g=False, mask_from_x=True)
            grads[blk.bn1.weight], grads[blk.bn1.bias] = dg, db
            wgrad_of(blk.conv1.weight, xin, dc1, planes, 1, 1, 0, 1)
            dx1, _ = r.conv(dc1, rec["w1t"], inpl, 1, 1, 0, 1, False)
            planes = outc                                  # width of the residual branch below
        else:
            # out = relu(bn2(c2) + res): g2 = dy * (out > 0)
            dc2, g2, dg, db = r.bn_bwd(dy, dy2, rec["out"], rec["c2"], rec["v2"], blk.bn2.weight.detach(), training, dy_is_f32=dy_f32)
            dy_f32 = False
            grads[blk.bn2.weight], grads[blk.bn2.bias] = dg, db
            wgrad_of(blk.conv2.weight, rec["a1"], dc2, planes, 3, 1, dil, dil)
            da1, _ = r.conv(dc2, rec["w2t"], planes, 3, 1, dil, dil, False)            # dgrad of conv2 (unit stride)
            dc1, _, dg, db = r.bn_bwd(da1, None, None, rec["c1"], rec["v1"], blk.bn1.weight.detach(), training, want_g=False, mask_from_x=True)
            grads[blk.bn1.weight], grads[blk.bn1.bias] = dg, db
            wgrad_of(blk.conv1.weight, rec["xin"], dc1, planes, 3, st, dil, dil)
            dx1 = dgrad_3x3(dc1, rec["w1t"], blk.conv1, xin.shape, inpl, planes, st, dil)      # dgrad of conv1
        if "cd" in rec:
            dconv, dbn = blk.downsample[0], blk.downsample[1]
            dcd, _, dg, db = r.bn_bwd(g2, None, None, rec["cd"], rec["vd"], dbn.weight.detach(), training, want_g=False)
            grads[dbn.weight], grads[dbn.bias] = dg, db
            wgrad_of(dconv.weight, xin, dcd, planes, 1, dconv.stride[0], 0, 1)
            if dconv.stride[0] == 1:
                dx2, _ = r.conv(dcd, rec["wdt"], inpl, 1, 1, 0, 1, False)
            else:
                up = r.empty(tuple(xin.shape[:4]) + (planes,))
                r.chk(lib.mmad_upsample_zero2(_p(dcd), _p(up), xin.shape[0], dcd.shape[1], dcd.shape[2], dcd.shape[3], xin.shape[1],
                                              xin.shape[2], xin.shape[3], planes, r.stream), "mmad_upsample_zero2")
                dx2, _ = r.conv(up, rec["wdt"], inpl, 1, 1, 0, 1, False)
        elif rec.get("short_a"):
            # shortcut 'A': the reference rebuilds the shortcut from `.data` (resnet.py:35, Variable(torch.cat([out.data, ...]))),
            # which DETACHES it - no gradient flows through the shortcut branch
            dx2 = None
        else:
            dx2 = g2                                                                # identity shortcut
        dy, dy2 = dx1, dx2

    if grads.reducer is not None:
        grads.reducer.flush()                              # the last large bucket is all-reduced under the stem's backward kernels
    # ---- stem backward: maxpool, relu+bn1 (mask recomputed from c0), wgrad of the space-to-depth GEMM.  No input gradient: the MRI
    #      volume is data ----
    stem = tape["stem"]
    n, d, h, w = stem["in_shape"]
    c0, v0 = stem["c0"], stem["v0"]
    dsum = r.empty(dy.shape)
    # dy + dy2 is needed as one tensor by the max-pool gather; fold the add into a bn_apply with scale 1 / shift 0
    ones = torch.ones((4, 64), dtype=torch.float32, device=dy.device)
    ones[3].zero_()
    r.chk(lib.mmad_bn_apply(_p(dy), _p(ones[2]), _p(ones[3]), _p(dy2), None, None, 0, _p(dsum), None, dy.numel() // 64, 64, r.stream),
          "mmad_bn_apply")
    da0 = r.empty(c0.shape)
    r.chk(lib.mmad_maxpool3d_bwd(_p(dsum), _p(stem["idx0"]), _p(da0), n, c0.shape[1], c0.shape[2], c0.shape[3], 64, r.stream),
          "mmad_maxpool3d_bwd")
    dc0, _, dg, db = r.bn_bwd(da0, None, None, c0, v0, model.bn1.weight.detach(), training, want_g=False, mask_from_x=True)
    grads[model.bn1.weight], grads[model.bn1.bias] = dg, db
    gw = None
    if grads.wanted(model.conv1.weight):
        gw = grads.new_grad(model.conv1.weight)
        r.stem_wgrad(stem["xs"], dc0, n, d, h, w, gw)
    r.join_side()                                          # every weight gradient is complete on the main stream from here
    if gw is not None:
        grads[model.conv1.weight] = gw
    return grads


def tape_stages(model: "ResNet", tape) -> dict:
    """Stored activations of a forward tape as {stage name: NCDHW fp32 tensor} (test / debugging aid; the names are
    the ones oracle/resnet_oracle.py accepts for `forced`)."""
    f = lambda t: t.float().permute(0, 4, 1, 2, 3)          # noqa: E731
    out = {"c0": f(tape["stem"]["c0"]), "p0": f(tape["stem"]["p0"])}
    names = [f"layer{li}.{bi}" for li, layer in enumerate((model.layer1, model.layer2, model.layer3, model.layer4), 1)
             for bi in range(len(layer))]
    for pre, rec in zip(names, tape["blocks"]):
        for k in ("c1", "a1", "c2", "a2", "c3", "cd", "out"):
            if k in rec:
                out[f"{pre}.{k}"] = f(rec[k])
    return out


class _BackboneFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, model, *params):
        feats, tape = _backbone_forward(model, x, model.training, True)
        ctx.model, ctx.tape, ctx.params = model, tape, params
        model._last_tape = tape if getattr(model, "keep_tape", False) else None
        return feats.permute(0, 4, 1, 2, 3)              # (N, C, D', H', W') view over NDHWC memory

    @staticmethod
    def backward(ctx, grad_out):
        # ctx.needs_input_grad[2:] follows the parameters' requires_grad: frozen layers get no wgrad kernels, no .grad and no
        # all-reduce (an optimizer over model.parameters() must not see - and weight-decay - a frozen tensor)
        wanted = {id(p) for p, need in zip(ctx.params, ctx.needs_input_grad[2:]) if need}
        if ctx.tape is None:
            raise RuntimeError("the activation tape of this forward pass was released by the first backward; set "
                               "model.retain_tape = True before the forward to run backward(retain_graph=True) twice")
        grads = _backbone_backward(ctx.model, ctx.tape, grad_out, wanted=wanted)
        if not getattr(ctx.model, "retain_tape", False):
            ctx.tape = None                                # like autograd's saved tensors without retain_graph: freed after one use
        red = getattr(ctx.model, "grad_reducer", None)
        if red is not None and red.active:
            # Data parallel: the reducer's asynchronous all-reduces average these tensors IN PLACE after this function
            # returns, so they must become the parameters' .grad themselves (autograd's AccumulateGrad would clone the
            # not-yet-reduced values).  Standard loop (zero_grad(set_to_none=True)): assign and report "no gradient".
            if all(p.grad is None for p in ctx.params if grads.get(p) is not None):
                for p in ctx.params:
                    g = grads.get(p)
                    if g is not None:
                        p.grad = g
                return (None, None) + (None,) * len(ctx.params)
            red.finish()                                   # gradient accumulation: hand autograd fully averaged tensors
        return (None, None) + tuple(grads.get(p) for p in ctx.params)


class ResNet(nn.Module):
    # resnet.py:112-215
    def __init__(self, block, layers, sample_input_D, sample_input_H, sample_input_W, num_seg_classes,
                 shortcut_type='B', no_cuda=False):
        self.inplanes = 64
        self.no_cuda = no_cuda
        super().__init__()
        self.block_type = block
        self.shortcut_type = shortcut_type
        self.conv1 = nn.Conv3d(1, 64, kernel_size=7, stride=(2, 2, 2), padding=(3, 3, 3), bias=False)
        self.bn1 = nn.BatchNorm3d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool3d(kernel_size=(3, 3, 3), stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0], shortcut_type)
        self.layer2 = self._make_layer(block, 128, layers[1], shortcut_type, stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], shortcut_type, stride=1, dilation=2)
        self.layer4 = self._make_layer(block, 512, layers[3], shortcut_type, stride=1, dilation=4)
        self.conv_seg = nn.Sequential(
            nn.ConvTranspose3d(512 * block.expansion, 32, 2, stride=2),
            nn.BatchNorm3d(32),
            nn.ReLU(inplace=True),
            nn.Conv3d(32, 32, kernel_size=3, stride=(1, 1, 1), padding=(1, 1, 1), bias=False),
            nn.BatchNorm3d(32),
            nn.ReLU(inplace=True),
            nn.Conv3d(32, num_seg_classes, kernel_size=1, stride=(1, 1, 1), bias=False))
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out')
            elif isinstance(m, nn.BatchNorm3d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, shortcut_type, stride=1, dilation=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            if shortcut_type == 'A':
                downsample = partial(downsample_basic_block, planes=planes * block.expansion, stride=stride,
                                     no_cuda=self.no_cuda)
            else:
                downsample = nn.Sequential(
                    nn.Conv3d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                    nn.BatchNorm3d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride=stride, dilation=dilation, downsample=downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, dilation=dilation))
        return nn.Sequential(*layers)

    def backbone_parameters(self):
        ps = [self.conv1.weight, self.bn1.weight, self.bn1.bias]
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                ps += [blk.conv1.weight, blk.bn1.weight, blk.bn1.bias, blk.conv2.weight, blk.bn2.weight, blk.bn2.bias]
                if isinstance(blk, Bottleneck):
                    ps += [blk.conv3.weight, blk.bn3.weight, blk.bn3.bias]
                if isinstance(blk.downsample, nn.Module):                  # shortcut 'B'; type 'A' is a parameter-free partial
                    ps += [blk.downsample[0].weight, blk.downsample[1].weight, blk.downsample[1].bias]
        return ps

    def features(self, x):
        """conv1 … layer4 (resnet.py:205-212) -> (N, 512, D/8.., H/8.., W/8..) fp32."""
        if not x.is_cuda:
            raise _lib.MmadError("multimodal_ad_b200 ResNet runs on CUDA tensors only (no CPU fallback)")
        if self.block_type not in (BasicBlock, Bottleneck):
            raise _lib.MmadError("accelerated path covers BasicBlock / Bottleneck networks")
        with torch.cuda.device(x.device):
            if torch.is_grad_enabled() and any(p.requires_grad for p in self.backbone_parameters()):
                return _BackboneFunction.apply(x, self, *self.backbone_parameters())
            feats, _ = _backbone_forward(self, x, self.training, False)
            return feats.permute(0, 4, 1, 2, 3)

    def forward(self, x):
        x = self.features(x)
        x = self.conv_seg(x)
        return x


def resnet10(**kwargs):
    return ResNet(BasicBlock, [1, 1, 1, 1], **kwargs)


def resnet18(**kwargs):
    return ResNet(BasicBlock, [2, 2, 2, 2], **kwargs)


def resnet34(**kwargs):
    return ResNet(BasicBlock, [3, 4, 6, 3], **kwargs)


def resnet50(**kwargs):
    return ResNet(Bottleneck, [3, 4, 6, 3], **kwargs)


def resnet101(**kwargs):
    return ResNet(Bottleneck, [3, 4, 23, 3], **kwargs)


def resnet152(**kwargs):
    return ResNet(Bottleneck, [3, 8, 36, 3], **kwargs)


def resnet200(**kwargs):
    return ResNet(Bottleneck, [3, 24, 36, 3], **kwargs)
