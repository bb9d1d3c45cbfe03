"""3-D U-Net image branch — drop-in for /root/reference/models/unet3d.py.

Same names (`Conv3DBlock`, `UpConv3DBlock`, `UNet3D`), same constructor arguments, same parameter / buffer names and the same
default initialisation stream (so `state_dict()`s are interchangeable, unet3d.py:100-113).  `UNet3D.forward` (unet3d.py:137-157:
zero-extend to 96x112x96, three analysis blocks, bottleneck, three synthesis blocks, crop back) runs on the CUDA library behind
include/mmad_b200.h inside ONE `torch.autograd.Function`:

  * a_block1.conv1 (1 -> 32 channels, K = 27)         direct convolution, fp32 operands (csrc/unet_kernels.cu)
  * every other 3x3x3 convolution                      implicit GEMM on tcgen05 / TMEM (csrc/conv3d_igemm.cu); the 32-channel
                                                       tensor is carried as 64 channels (upper half zero) so that every K slice is
                                                       one 128-byte row
  * ConvTranspose3d(2, 2)                              eight 1x1x1 phase GEMMs writing interleaved, straight into the first
                                                       channels of the concatenation buffer
  * torch.cat((up, residual), 1) (unet3d.py:77)        no copy: both producers write their channel slice of one NDHWC buffer,
                                                       which the next convolution reads as a plain tensor
  * BatchNorm3d + ReLU                                 training: statistics from the convolution epilogue, one apply pass;
                                                       eval without autograd (image_features.py:40-41): folded into the
                                                       producing convolution's epilogue - the pre-activation is never written
  * MaxPool3d(2, 2), the 1x1x1 head + crop back        csrc/unet_kernels.cu

Convolution biases (unet3d.py:37-40 uses the default bias=True): a bias followed by BatchNorm in training mode cancels in the
mean subtraction, so the convolution output is stored without it; it moves the running mean (added there) and its gradient is
identically zero.  In eval mode it is folded into the BatchNorm shift.

The raw output of `s_block1.conv2` (64 channels, bias included) is what image_features.py:58-60 grabs with a forward hook and
pools over the atlas.  The convolution epilogue writes it once more as fp32 NDHWC; forward hooks registered on
`model.s_block1.conv2` / `model` fire as in the reference (the hook sees an (N,64,96,112,96)-shaped view), and
`UNet3D.roi_features(vol, plan)` is the accelerated form of image_features.py:97-114: forward + ROI pooling with the feature
map never leaving the GPU.

The nn.Conv3d / nn.BatchNorm3d / nn.ConvTranspose3d objects are parameter containers only.  Activations live as NDHWC bf16
between kernels, accumulation is fp32.  No CPU / cuDNN fallback: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes
from ctypes import c_void_p

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib
from .resnet import _p, _Run

__all__ = ["Conv3DBlock", "UpConv3DBlock", "UNet3D"]


class Conv3DBlock(nn.Module):
    # unet3d.py:14-46
    def __init__(self, in_channels, out_channels, bottleneck=False) -> None:
        super().__init__()
        self.conv1 = nn.Conv3d(in_channels=in_channels, out_channels=out_channels // 2, kernel_size=(3, 3, 3), padding=1)
        self.bn1 = nn.BatchNorm3d(num_features=out_channels // 2)
        self.conv2 = nn.Conv3d(in_channels=out_channels // 2, out_channels=out_channels, kernel_size=(3, 3, 3), padding=1)
        self.bn2 = nn.BatchNorm3d(num_features=out_channels)
        self.relu = nn.ReLU()
        self.bottleneck = bottleneck
        if not bottleneck:
            self.pooling = nn.MaxPool3d(kernel_size=(2, 2, 2), stride=2)

    def forward(self, input):
        raise _lib.MmadError("Conv3DBlock is a parameter container here: run it through UNet3D.forward (the accelerated path)")


class UpConv3DBlock(nn.Module):
    # unet3d.py:51-84
    def __init__(self, in_channels, res_channels=0, last_layer=False, num_classes=None) -> None:
        super().__init__()
        assert (last_layer == False and num_classes == None) or (last_layer == True and num_classes != None), 'Invalid arguments'  # noqa: E711,E712
        self.upconv1 = nn.ConvTranspose3d(in_channels=in_channels, out_channels=in_channels, kernel_size=(2, 2, 2), stride=2)
        self.relu = nn.ReLU()
        self.bn = nn.BatchNorm3d(num_features=in_channels // 2)
        self.conv1 = nn.Conv3d(in_channels=in_channels + res_channels, out_channels=in_channels // 2, kernel_size=(3, 3, 3), padding=(1, 1, 1))
        self.conv2 = nn.Conv3d(in_channels=in_channels // 2, out_channels=in_channels // 2, kernel_size=(3, 3, 3), padding=(1, 1, 1))
        self.last_layer = last_layer
        if last_layer:
            self.conv3 = nn.Conv3d(in_channels=in_channels // 2, out_channels=num_classes, kernel_size=(1, 1, 1))

    def forward(self, input, residual=None):
        raise _lib.MmadError("UpConv3DBlock is a parameter container here: run it through UNet3D.forward (the accelerated path)")


# ----------------------------------------------------------------------------------------------------------------
# kernel plumbing
# ----------------------------------------------------------------------------------------------------------------
def _ptr(t, byte_offset=0):
    return c_void_p(t.data_ptr() + byte_offset)


class _UNetRun(_Run):
    """_Run (models/resnet.py) + the UNet-specific entry points."""

    def prep_w(self, w, want_dgrad):
        """torch-layout fp32 weight (Cout, Cin, k,k,k) -> forward layout [Cout][taps][Cin] bf16 (+ dgrad layout)."""
        cout, cin = w.shape[0], w.shape[1]
        taps = w.numel() // (cout * cin)
        w = w.contiguous()
        wf = self.empty((cout, taps, cin))
        wt = self.empty((cin, taps, cout)) if want_dgrad else None
        self.chk(self.lib.mmad_conv3d_prep_weights(_p(w), _p(wf), _p(wt), cout, cin, taps, self.stream), "mmad_conv3d_prep_weights")
        return wf, wt

    def conv_ex(self, x, w_fwd, cout, y_ptr, ldy, want_stats, scale=None, shift=None, relu=False, out_f32=None, f32_bias=None,
                k=3, stride=1, pad=1):
        n, d, h, w, cin = x.shape
        cin_tensor = 0
        if cin < 64:                                       # a 32-channel tensor: K slices stay 64 wide, TMA zero-fills the rest in flight
            cin_tensor, cin = cin, 64
        part = None
        if want_stats:
            npart = self.lib.mmad_conv3d_stats_partials(n, d, h, w, cout, k, stride, pad, 1)
            part = self.empty((npart, cout, 2), torch.float32)
        self.chk(self.lib.mmad_conv3d_fwd_ex_bf16(_p(x), _p(w_fwd), y_ptr, ldy, _p(part), _p(scale), _p(shift), 1 if relu else 0,
                                                  _p(out_f32), _p(f32_bias), n, d, h, w, cin, cout, k, stride, pad, 1, cin_tensor, self.stream),
                 "mmad_conv3d_fwd_ex_bf16")
        return part

    def wgrad_ex(self, x_ptr, ldx, dims, cin, dy, cout, k, stride, pad, out_dw, cin_total, ci_off, keep=()):
        """weight gradient for the channel slice [ci_off, ci_off + cin) of the input; out_dw: torch layout (cout, cin_total, taps)."""
        n, d, h, w = dims
        nsplit = ctypes.c_int(0)
        elems = self.lib.mmad_conv3d_wgrad_workspace(n, d, h, w, cin, cout, k, stride, pad, 1, ctypes.byref(nsplit))
        if elems < 0:
            raise _lib.MmadError("mmad_conv3d_wgrad_workspace: bad geometry")
        ws = self.empty((elems,), torch.float32)
        cs = self.stream
        if self.side is not None:
            self.side.wait_stream(self.main)
            cs = c_void_p(self.side.cuda_stream)
            self._keep += [dy, ws, *keep]
        self.chk(self.lib.mmad_conv3d_wgrad_ex_bf16(x_ptr, ldx, _p(dy), _p(ws), n, d, h, w, cin, cout, k, stride, pad, 1, cs),
                 "mmad_conv3d_wgrad_ex_bf16")
        self.chk(self.lib.mmad_wgrad_reduce_ex(_p(ws), nsplit.value, _p(out_dw), cout, cin, k * k * k, cin_total, ci_off, cs),
                 "mmad_wgrad_reduce_ex")


def _pad_vec(v, c, fill=0.0):
    return v if v.numel() == c else torch.cat([v, torch.full((c - v.numel(),), fill, dtype=v.dtype, device=v.device)])


def _bn_vec(r: _UNetRun, bn: nn.BatchNorm3d, part, count, bias, training, width=None):
    """(4, C) fp32: mean, invstd, scale, shift of `bn` applied to a convolution output stored WITHOUT its bias `bias`.
    width > bn.num_features: the tensor carries zero padding channels (gamma = beta = 0 there).  Returns (vec, padded gamma)."""
    c = bn.num_features
    width = width or c
    gamma, beta = _pad_vec(bn.weight.detach(), width), _pad_vec(bn.bias.detach(), width)
    vec = r.empty((4, width), torch.float32)
    lib = r.lib
    if training:
        momentum = 0.1 if bn.momentum is None else bn.momentum
        track = bn.track_running_stats and bn.running_mean is not None
        rm = _pad_vec(bn.running_mean, width) if track else None
        rv = _pad_vec(bn.running_var, width, 1.0) if track else None
        r.chk(lib.mmad_bn_finalize(_p(part), part.shape[0], width, float(count), _p(gamma), _p(beta), bn.eps, momentum, _p(rm), _p(rv),
                                   _p(vec[0]), _p(vec[1]), _p(vec[2]), _p(vec[3]), r.stream), "mmad_bn_finalize")
        if track:
            if width != c:
                bn.running_mean.copy_(rm[:c])
                bn.running_var.copy_(rv[:c])
            if bias is not None:
                bn.running_mean.add_(bias.detach(), alpha=momentum)    # the batch mean of (conv + bias) is the stored mean + bias
            if bn.num_batches_tracked is not None:
                r.tracked.append(bn.num_batches_tracked)    # bumped together at the end of the pass (a shared BatchNorm appears twice)
    else:
        # eval: xhat = (c + bias - running_mean) * invstd with c stored without the bias -> use mean' = running_mean - bias
        mean = bn.running_mean if bias is None else bn.running_mean - bias.detach()
        mean_p, var_p = _pad_vec(mean, width), _pad_vec(bn.running_var, width, 1.0)   # named: alive until the call is enqueued
        r.chk(lib.mmad_bn_eval_params(width, _p(gamma), _p(beta), _p(mean_p), _p(var_p),
                                      bn.eps, _p(vec[0]), _p(vec[1]), _p(vec[2]), _p(vec[3]), r.stream), "mmad_bn_eval_params")
    return vec, gamma


def _unet_forward(model: "UNet3D", x: torch.Tensor, training: bool, need_grad: bool, want_hook: bool):
    """Returns (out fp32 (N,K,D,H,W), hook fp32 (N,Dp,Hp,Wp,64) NDHWC or None, tape)."""
    r = _UNetRun(x.device)
    lib = r.lib
    n, cin, d, h, w = x.shape
    dp, hp, wp = model.target
    if cin != 1:
        raise _lib.MmadError("the accelerated UNet3D expects 1 input channel (image_features.py:40: UNet3D(in_channels=1, ...))")
    if d > dp or h > hp or w > wp:
        raise _lib.MmadError(f"input {(d, h, w)} exceeds the padded grid {model.target} (unet3d.py:116-123)")
    x = x.contiguous().float()
    fold = (not training) and (not need_grad)              # eval without autograd: BatchNorm + ReLU live in the conv epilogue
    blocks_a = [model.a_block1, model.a_block2, model.a_block3, model.bottleNeck]
    blocks_s = [model.s_block3, model.s_block2, model.s_block1]
    tape = {"training": training, "in_shape": (n, d, h, w), "x": x if need_grad else None, "enc": [], "dec": []}

    # every 3x3x3 convolution's weights (but the fp32 first layer's) re-laid in two launches
    items = []
    for bi, blk in enumerate(blocks_a):
        if bi:
            items.append((blk.conv1, blk.conv1.weight, need_grad))
        w2 = blk.conv2.weight.detach()
        items.append((blk.conv2, F.pad(w2, (0, 0, 0, 0, 0, 0, 0, 64 - w2.shape[1])) if bi == 0 else w2, need_grad))
    for blk in blocks_s:
        items += [(blk.conv1, blk.conv1.weight, need_grad), (blk.conv2, blk.conv2.weight, need_grad)]
    prepped = r.prep_all(items)

    def conv_bn_relu(xin, conv, bn, out_t, out_off, out_ld, w_pad_cin=None, hook=False):
        """relu(bn(conv(xin) + bias)) written to channels [out_off, out_off + Cout) of out_t (rows out_ld apart; 0 = dense)."""
        wf, wt = prepped[conv]
        cout = conv.out_channels
        rows = xin.numel() // xin.shape[-1]
        hook32 = r.empty(tuple(xin.shape[:4]) + (cout,), torch.float32) if hook else None
        y_ptr = _ptr(out_t, out_off * 2)
        if fold:
            vec, gamma = _bn_vec(r, bn, None, 0, conv.bias, False)
            r.conv_ex(xin, wf, cout, y_ptr, out_ld, False, scale=vec[2], shift=vec[3], relu=True, out_f32=hook32,
                      f32_bias=conv.bias.detach() if hook else None)
            return dict(conv=conv, bn=bn, xin=xin), hook32
        c = r.empty(tuple(xin.shape[:4]) + (cout,))
        part = r.conv_ex(xin, wf, cout, _ptr(c), 0, training, out_f32=hook32, f32_bias=conv.bias.detach() if hook else None)
        vec, gamma = _bn_vec(r, bn, part, rows, conv.bias, training)
        r.chk(lib.mmad_bn_apply_ex(_p(c), _p(vec[2]), _p(vec[3]), None, None, None, 1, y_ptr, out_ld, None, rows, cout, r.stream),
              "mmad_bn_apply_ex")
        return dict(conv=conv, bn=bn, xin=xin, c=c, vec=vec, gamma=gamma, wt=wt, w_pad_cin=w_pad_cin), hook32

    # ---- concatenation buffers (unet3d.py:77): [up | skip] per level, written in place by their producers ----
    c1, c2, c3 = blocks_a[0].conv2.out_channels, blocks_a[1].conv2.out_channels, blocks_a[2].conv2.out_channels
    ups = [blk.upconv1.out_channels for blk in blocks_s]                  # 512, 256, 128
    grids = [(dp, hp, wp), (dp // 2, hp // 2, wp // 2), (dp // 4, hp // 4, wp // 4), (dp // 8, hp // 8, wp // 8)]
    cats = [r.empty((n,) + grids[0] + (ups[2] + c1,)), r.empty((n,) + grids[1] + (ups[1] + c2,)), r.empty((n,) + grids[2] + (ups[0] + c3,))]
    cat_up = [ups[2], ups[1], ups[0]]

    # ---- analysis path (unet3d.py:34-46, 141-144) ----
    cur = None
    for li, blk in enumerate(blocks_a):
        g = grids[li]
        rows = n * g[0] * g[1] * g[2]
        rec = {"blk": blk, "grid": g}
        mid = blk.conv1.out_channels
        if li == 0:
            # conv1: 1 -> 32 channels, direct kernel on the fp32 volume; the output grid is the zero-extended one
            width = 64
            wsrc = blk.conv1.weight.detach().contiguous()
            if fold:
                # eval without autograd: BatchNorm + ReLU (and the bias) folded into the kernel, one pass, one tensor written
                # ... and only the 32 real channels: the next convolution reads 64-byte rows and TMA pads its K slices in flight
                vec, gamma = _bn_vec(r, blk.bn1, None, rows, blk.conv1.bias, False, width=width)
                a1 = r.empty((n,) + g + (mid,))
                r.chk(lib.mmad_conv3d_c1_fwd(_p(x), _p(wsrc), _p(a1), None, _p(vec[2]), _p(vec[3]), n, d, h, w, g[0], g[1], g[2], mid,
                                             r.stream), "mmad_conv3d_c1_fwd")
                rec["l1"] = dict(conv=blk.conv1, bn=blk.bn1, first=True)
            else:
                cc = r.empty((n,) + g + (width,))
                nb = lib.mmad_conv3d_c1_blocks(n, *g)
                part = r.empty((nb, width, 2), torch.float32) if training else None
                r.chk(lib.mmad_conv3d_c1_fwd(_p(x), _p(wsrc), _p(cc), _p(part), None, None, n, d, h, w, g[0], g[1], g[2], width,
                                             r.stream), "mmad_conv3d_c1_fwd")
                vec, gamma = _bn_vec(r, blk.bn1, part, rows, blk.conv1.bias, training, width=width)
                a1 = r.empty(cc.shape)
                r.chk(lib.mmad_bn_apply_ex(_p(cc), _p(vec[2]), _p(vec[3]), None, None, None, 1, _p(a1), 0, None, rows, width, r.stream),
                      "mmad_bn_apply_ex")
                rec["l1"] = dict(conv=blk.conv1, bn=blk.bn1, c=cc, vec=vec, gamma=gamma, first=True)
            pad_cin = width
        else:
            a1 = r.empty((n,) + g + (mid,))
            rec["l1"], _ = conv_bn_relu(cur, blk.conv1, blk.bn1, a1, 0, 0)
            pad_cin = None
        cout = blk.conv2.out_channels
        if blk.bottleneck:
            res = r.empty((n,) + g + (cout,))
            rec["l2"], _ = conv_bn_relu(a1, blk.conv2, blk.bn2, res, 0, 0, w_pad_cin=pad_cin)
            cur = res
        else:
            cat = cats[li]
            ccat = cat.shape[-1]
            rec["l2"], _ = conv_bn_relu(a1, blk.conv2, blk.bn2, cat, cat_up[li], ccat, w_pad_cin=pad_cin)
            pooled = r.empty((n, g[0] // 2, g[1] // 2, g[2] // 2, cout))
            idx = torch.empty(pooled.shape, dtype=torch.uint8, device=x.device) if need_grad else None
            r.chk(lib.mmad_maxpool3d_k2_fwd(_ptr(cat, cat_up[li] * 2), ccat, _p(pooled), _p(idx), n, g[0], g[1], g[2], cout, r.stream),
                  "mmad_maxpool3d_k2_fwd")
            rec["idx"] = idx
            rec["pooled"] = pooled
            cur = pooled
        tape["enc"].append(rec)

    # ---- synthesis path (unet3d.py:74-84, 147-149) ----
    hook32 = None
    for si, blk in enumerate(blocks_s):
        li = 2 - si                                          # level of the output grid
        g, cat = grids[li], cats[li]
        ccat, cup = cat.shape[-1], cat_up[li]
        gi = grids[li + 1]
        upc = blk.upconv1
        wph = r.empty((8, cup, upc.in_channels))
        r.chk(lib.mmad_convtranspose3d_prep_weights(_p(upc.weight.detach().contiguous()), _p(wph), upc.in_channels, cup, r.stream),
              "mmad_convtranspose3d_prep_weights")
        r.chk(lib.mmad_convtranspose3d_k2s2_fwd_bf16(_p(cur), _p(wph), _p(upc.bias.detach()), _p(cat), ccat, n, gi[0], gi[1], gi[2],
                                                     upc.in_channels, cup, r.stream), "mmad_convtranspose3d_k2s2_fwd_bf16")
        rec = {"blk": blk, "grid": g, "xin": cur, "cat": cat, "cup": cup}
        mid = blk.conv1.out_channels
        a1 = r.empty((n,) + g + (mid,))
        rec["l1"], _ = conv_bn_relu(cat, blk.conv1, blk.bn, a1, 0, 0)
        if fold and blk.last_layer and not getattr(model, "keep_tape", False):
            # eval without autograd: conv2 + bn + relu + conv3 + crop back as ONE kernel, the last activation is never written
            head = blk.conv3
            k = head.out_channels
            wf, _ = prepped[blk.conv2]
            vec, _ = _bn_vec(r, blk.bn, None, 0, blk.conv2.bias, False)
            hook32 = r.empty((n,) + g + (mid,), torch.float32) if want_hook else None
            out = r.empty((n, k, d, h, w), torch.float32)
            hw = head.weight.detach().reshape(k, -1).contiguous()
            r.chk(lib.mmad_conv3d_fwd_head_bf16(_p(a1), _p(wf), _p(vec[2]), _p(vec[3]), _p(hook32), _p(blk.conv2.bias.detach()), _p(hw),
                                                _p(head.bias.detach()), _p(out), k, n, g[0], g[1], g[2], d, h, w, mid, r.stream),
                  "mmad_conv3d_fwd_head_bf16")
            return out, hook32, None
        a2 = r.empty((n,) + g + (mid,))
        rec["l2"], hk = conv_bn_relu(a1, blk.conv2, blk.bn, a2, 0, 0, hook=blk.last_layer and want_hook)
        if hk is not None:
            hook32 = hk
        rec["a2"] = a2
        cur = a2
        tape["dec"].append(rec)

    # ---- head + crop back (unet3d.py:72, 126-135) ----
    head = blocks_s[-1].conv3
    k = head.out_channels
    out = r.empty((n, k, d, h, w), torch.float32)
    r.chk(lib.mmad_head1x1_fwd(_p(cur), _p(head.weight.detach().reshape(k, -1).contiguous()), _p(head.bias.detach()), _p(out), n, dp, hp, wp,
                               d, h, w, cur.shape[-1], k, r.stream), "mmad_head1x1_fwd")
    r.bump_tracked()
    if not need_grad and not getattr(model, "keep_tape", False):
        tape = None
    return out, hook32, tape


def _unet_backward(model: "UNet3D", tape, grad_out: torch.Tensor, wanted=None):
    """grad_out: gradient of the (N,K,D,H,W) output.  Returns {parameter: gradient}."""
    r = _UNetRun(grad_out.device, use_side_stream=getattr(model, "wgrad_side_stream", True))
    lib = r.lib
    training = tape["training"]
    n, d, h, w = tape["in_shape"]
    dp, hp, wp = model.target
    grads = {}
    post = []                                              # fix-ups that read side-stream results (run after join_side)

    def want(p):
        return wanted is None or id(p) in wanted

    def put(p, g):
        if want(p):
            grads[p] = g if p not in grads else grads[p] + g     # the shared BatchNorm of an up block collects two contributions

    def layer_backward(rec, dy, dy2, xin, dims, cat=None):
        """Backward of relu(bn(conv(xin) + bias)) given d(activation) = dy (+ dy2).  Returns dc (gradient of the conv output).
        Weight / bias / BatchNorm gradients are stored; the data gradient is left to the caller (it depends on the input)."""
        conv, bn = rec["conv"], rec["bn"]
        # one upstream gradient: four-pass form (no masked gradient written, the mask recomputed from c in both passes); the skip
        # tensors receive two (max-pool path + decoder path) and keep the form that materialises their masked sum
        dc, _, dgamma, dbeta = r.bn_bwd(dy, dy2, None, rec["c"], rec["vec"], rec["gamma"], training, want_g=dy2 is not None, mask_from_x=True)
        c = bn.num_features
        put(bn.weight, dgamma[:c])
        put(bn.bias, dbeta[:c])
        if want(conv.bias):
            # training: BatchNorm's backward output sums to zero over the batch - the bias gradient is identically 0;
            # eval: BatchNorm is affine, d(bias) = scale * sum(g)
            put(conv.bias, torch.zeros_like(conv.bias) if training else (rec["vec"][2] * dbeta)[:c].clone())
        return dc

    # ---- head ----
    head = model.s_block1.conv3
    k = head.out_channels
    last = tape["dec"][-1]
    a2 = last["a2"]
    go = grad_out.contiguous().float()
    da = r.empty(a2.shape)
    nb = lib.mmad_head1x1_bwd_blocks()
    part = r.empty((nb, k, 65), torch.float32)
    dwh = r.empty((k, 64), torch.float32)
    dbh = r.empty((k,), torch.float32)
    r.chk(lib.mmad_head1x1_bwd(_p(a2), _p(head.weight.detach().reshape(k, -1).contiguous()), _p(go), _p(da), _p(part), _p(dwh), _p(dbh),
                               n, dp, hp, wp, d, h, w, 64, k, r.stream), "mmad_head1x1_bwd")
    put(head.weight, dwh.view_as(head.weight))
    put(head.bias, dbh)

    # ---- synthesis path, last block first ----
    dskip = {}
    dcur = da                                                # gradient of the block's output activation
    for rec in reversed(tape["dec"]):
        blk, g, cat, cup = rec["blk"], rec["grid"], rec["cat"], rec["cup"]
        dims = (n,) + g
        ccat = cat.shape[-1]
        mid = blk.conv1.out_channels
        l1, l2 = rec["l1"], rec["l2"]
        # conv2: input a1 = l2["xin"]
        dc2 = layer_backward(l2, dcur, None, l2["xin"], dims)
        if want(blk.conv2.weight):
            gw = torch.empty_like(blk.conv2.weight)
            r.wgrad_ex(_p(l2["xin"]), mid, dims, mid, dc2, mid, 3, 1, 1, gw, mid, 0, keep=(l2["xin"],))
            put(blk.conv2.weight, gw)
        da1, _ = r.conv(dc2, l2["wt"], mid, 3, 1, 1, 1, False)
        # conv1: input = the concatenation buffer [up | skip]
        dc1 = layer_backward(l1, da1, None, cat, dims)
        if want(blk.conv1.weight):
            gw = torch.empty_like(blk.conv1.weight)
            if mid == 64:
                # 64 output channels: one halo weight-gradient GEMM (64 -> 64, all 27 taps from one input box) per 64-channel slab of
                # the concatenated input - 3 x 0.9 ms instead of 4.3 + 0.9 ms through the generic kernel at batch 4 x 96x112x96
                for c0 in range(0, ccat, 64):
                    r.wgrad_ex(_ptr(cat, c0 * 2), ccat, dims, 64, dc1, mid, 3, 1, 1, gw, ccat, c0, keep=(cat,))
            elif ccat % 128 == 0:
                r.wgrad_ex(_p(cat), ccat, dims, ccat, dc1, mid, 3, 1, 1, gw, ccat, 0, keep=(cat,))
            else:                                            # one weight-gradient GEMM per source of the concatenation
                r.wgrad_ex(_p(cat), ccat, dims, cup, dc1, mid, 3, 1, 1, gw, ccat, 0, keep=(cat,))
                r.wgrad_ex(_ptr(cat, cup * 2), ccat, dims, ccat - cup, dc1, mid, 3, 1, 1, gw, ccat, cup, keep=(cat,))
            put(blk.conv1.weight, gw)
        # data gradient of conv1, one convolution per source (rows [0, cup) and [cup, ccat) of the dgrad weights)
        wt = l1["wt"]                                        # [ccat][27][mid]
        dup = r.empty(dims + (cup,))
        npart = lib.mmad_conv3d_stats_partials(n, g[0], g[1], g[2], cup, 3, 1, 1, 1)
        spart = r.empty((npart, cup, 2), torch.float32)     # per-channel sums of d(up): the transposed convolution's bias gradient
        r.chk(lib.mmad_conv3d_fwd_bf16(_p(dc1), _p(wt), _p(dup), _p(spart), n, g[0], g[1], g[2], mid, cup, 3, 1, 1, 1, r.stream),
              "mmad_conv3d_fwd_bf16")
        dres = r.empty(dims + (ccat - cup,))
        r.chk(lib.mmad_conv3d_fwd_bf16(_p(dc1), _ptr(wt, cup * 27 * mid * 2), _p(dres), None, n, g[0], g[1], g[2], mid, ccat - cup, 3, 1, 1,
                                       1, r.stream), "mmad_conv3d_fwd_bf16")
        dskip[id(cat)] = dres
        # transposed convolution: bias, weight and data gradients
        upc = blk.upconv1
        xin = rec["xin"]                                     # coarse input (N, g/2, Cin)
        cin_up = upc.in_channels
        put(upc.bias, spart[:, :, 0].sum(0))
        if want(upc.weight):
            gw = torch.empty_like(upc.weight)                # (Cin, Cout, 2,2,2) == the wgrad layout with the roles swapped
            r.wgrad_ex(_p(dup), cup, dims, cup, xin, cin_up, 2, 2, 0, gw, cup, 0, keep=(dup,))
            put(upc.weight, gw)
        wf_t, _ = r.prep_w(upc.weight.detach().reshape(cin_up, cup, 8), False)      # [Cin][8 taps][Cout]: conv weights Cout' = Cin
        dcur, _ = r.conv(dup, wf_t, cin_up, 2, 2, 0, 1, False)

    # ---- analysis path, bottleneck first ----
    level_cat = {tuple(rec["grid"]): rec["cat"] for rec in tape["dec"]}
    for rec in reversed(tape["enc"]):
        blk, g = rec["blk"], rec["grid"]
        dims = (n,) + g
        l1, l2 = rec["l1"], rec["l2"]
        cout = blk.conv2.out_channels
        if blk.bottleneck:
            dy, dy2 = dcur, None
        else:
            dy = r.empty(dims + (cout,))
            r.chk(lib.mmad_maxpool3d_k2_bwd(_p(dcur), _p(rec["idx"]), _p(dy), n, g[0], g[1], g[2], cout, r.stream), "mmad_maxpool3d_k2_bwd")
            dy2 = dskip[id(level_cat[tuple(g)])]
        a1 = l2["xin"]
        cin2 = a1.shape[-1]                                  # 64 for a_block1 (32 real channels + zero padding)
        dc2 = layer_backward(l2, dy, dy2, a1, dims)
        if want(blk.conv2.weight):
            real = blk.conv2.in_channels
            gw = torch.empty((cout, cin2) + tuple(blk.conv2.weight.shape[2:]), dtype=torch.float32, device=dy.device)
            r.wgrad_ex(_p(a1), cin2, dims, cin2, dc2, cout, 3, 1, 1, gw, cin2, 0, keep=(a1,))
            if real == cin2:
                put(blk.conv2.weight, gw)
            else:
                # the padded gradient is produced on the side stream: slice it only after the join below
                post.append(lambda p=blk.conv2.weight, t=gw, c=real: put(p, t[:, :c].contiguous()))
        da1, _ = r.conv(dc2, l2["wt"], cin2, 3, 1, 1, 1, False)
        if l1.get("first"):
            # conv1 of a_block1: direct kernel; the MRI volume is data, no input gradient
            conv, bn = l1["conv"], l1["bn"]
            dc1, _, dgamma, dbeta = r.bn_bwd(da1, None, None, l1["c"], l1["vec"], l1["gamma"], training, want_g=False, mask_from_x=True)
            c = bn.num_features
            put(bn.weight, dgamma[:c])
            put(bn.bias, dbeta[:c])
            if want(conv.bias):
                put(conv.bias, torch.zeros_like(conv.bias) if training else (l1["vec"][2] * dbeta)[:c].clone())
            if want(conv.weight):
                nb = lib.mmad_conv3d_c1_wgrad_blocks(n, *g)
                ws = r.empty((nb, 32, 27), torch.float32)
                gw = torch.empty_like(conv.weight)
                r.chk(lib.mmad_conv3d_c1_wgrad(_p(tape["x"]), _p(dc1), _p(ws), n, d, h, w, g[0], g[1], g[2], r.stream), "mmad_conv3d_c1_wgrad")
                r.chk(lib.mmad_wgrad_reduce(_p(ws), nb, _p(gw), 32, 1, 27, r.stream), "mmad_wgrad_reduce")
                put(conv.weight, gw)
        else:
            xin = l1["xin"]
            cin1 = xin.shape[-1]
            mid = blk.conv1.out_channels
            dc1 = layer_backward(l1, da1, None, xin, dims)
            if want(blk.conv1.weight):
                gw = torch.empty_like(blk.conv1.weight)
                r.wgrad_ex(_p(xin), cin1, dims, cin1, dc1, mid, 3, 1, 1, gw, cin1, 0, keep=(xin,))
                put(blk.conv1.weight, gw)
            dcur, _ = r.conv(dc1, l1["wt"], cin1, 3, 1, 1, 1, False)
    r.join_side()
    for fn in post:
        fn()
    return grads


def tape_stages(model: "UNet3D", tape) -> dict:
    """Stored activations of a (training-mode) forward tape as {stage name: NCDHW fp32 tensor} (test / debugging aid; the names
    are the ones oracle/unet_oracle.py accepts for `forced`)."""
    f = lambda t: t.float().permute(0, 4, 1, 2, 3)          # noqa: E731
    out = {}
    names = ["a_block1", "a_block2", "a_block3", "bottleNeck"]
    for name, rec in zip(names, tape["enc"]):
        c1 = rec["blk"].conv1.out_channels
        if "c" in rec["l1"]:
            out[f"{name}.c1"] = f(rec["l1"]["c"][..., :c1])
        out[f"{name}.a1"] = f(rec["l2"]["xin"][..., :c1])
        if "c" in rec["l2"]:
            out[f"{name}.c2"] = f(rec["l2"]["c"])
        if "pooled" in rec:
            out[f"{name}.p"] = f(rec["pooled"])
    for name, rec in zip(["s_block3", "s_block2", "s_block1"], tape["dec"]):
        cup = rec["cup"]
        out[f"{name}.up"] = f(rec["cat"][..., :cup])
        skip = {"s_block3": "a_block3", "s_block2": "a_block2", "s_block1": "a_block1"}[name]
        out[f"{skip}.a2"] = f(rec["cat"][..., cup:])
        if "c" in rec["l1"]:
            out[f"{name}.c1"] = f(rec["l1"]["c"])
        out[f"{name}.a1"] = f(rec["l2"]["xin"])
        if "c" in rec["l2"]:
            out[f"{name}.c2"] = f(rec["l2"]["c"])
        out[f"{name}.a2"] = f(rec["a2"])
    # the bottleneck's output is the first transposed convolution's input
    out["bottleNeck.a2"] = f(tape["dec"][0]["xin"])
    return out


class _UNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, model, want_hook, *params):
        out, hook32, tape = _unet_forward(model, x, model.training, True, want_hook)
        ctx.model, ctx.tape, ctx.params = model, tape, params
        model._last_tape = tape if getattr(model, "keep_tape", False) else None
        if hook32 is None:
            hook32 = out.new_zeros(())
        ctx.mark_non_differentiable(hook32)
        return out, hook32

    @staticmethod
    def backward(ctx, grad_out, _grad_hook):
        if ctx.tape is None:
            raise RuntimeError("the activation tape of this forward pass was released by the first backward; set "
                               "model.retain_tape = True before the forward to run backward(retain_graph=True) twice")
        wanted = {id(p) for p, need in zip(ctx.params, ctx.needs_input_grad[3:]) if need}
        grads = _unet_backward(ctx.model, ctx.tape, grad_out, wanted=wanted)
        if not getattr(ctx.model, "retain_tape", False):
            ctx.tape = None
        return (None, None, None) + tuple(grads.get(p) for p in ctx.params)


class UNet3D(nn.Module):
    # unet3d.py:87-157
    def __init__(self, in_channels, num_classes, level_channels=[64, 128, 256], bottleneck_channel=512):  # noqa: B006 - reference signature
        super().__init__()
        c1, c2, c3 = level_channels
        self.a_block1 = Conv3DBlock(in_channels, c1)
        self.a_block2 = Conv3DBlock(c1, c2)
        self.a_block3 = Conv3DBlock(c2, c3)
        self.bottleNeck = Conv3DBlock(c3, bottleneck_channel, bottleneck=True)
        self.s_block3 = UpConv3DBlock(bottleneck_channel, res_channels=c3)
        self.s_block2 = UpConv3DBlock(c3, res_channels=c2)
        self.s_block1 = UpConv3DBlock(c2, res_channels=c1, num_classes=num_classes, last_layer=True)
        self.target = (96, 112, 96)                        # unet3d.py:117 (the default argument of _pad_to_target)

    @staticmethod
    def _pad_to_target(x, target=(96, 112, 96)):
        # unet3d.py:116-123; kept for API parity (the accelerated path zero-extends inside its first kernel)
        _, _, D, H, W = x.shape
        tD, tH, tW = target
        pad = (0, tW - W, 0, tH - H, 0, tD - D)
        return F.pad(x, pad), pad

    @staticmethod
    def _crop_back(y, pad):
        # unet3d.py:126-135
        _, _, Dp, Hp, Wp = y.shape
        Dl, Dr = pad[4], pad[5]
        Hl, Hr = pad[2], pad[3]
        Wl, Wr = pad[0], pad[1]
        return y[:, :, Dl: Dp - Dr if Dr else None, Hl: Hp - Hr if Hr else None, Wl: Wp - Wr if Wr else None]

    def _check(self, x):
        if not x.is_cuda:
            raise _lib.MmadError("multimodal_ad_b200 UNet3D runs on CUDA tensors only (no CPU fallback)")
        c = [self.a_block1.conv1.out_channels, self.a_block1.conv2.out_channels, self.a_block2.conv2.out_channels,
             self.a_block3.conv2.out_channels, self.bottleNeck.conv2.out_channels]
        if c != [32, 64, 128, 256, 512]:
            raise _lib.MmadError("the accelerated path covers the reference's level_channels=[64,128,256], bottleneck_channel=512")
        if any(t % 8 for t in self.target):
            raise _lib.MmadError("the padded grid must be divisible by 8 (three 2x2x2 poolings)")

    def _run(self, x, want_hook):
        self._check(x)
        conv2 = self.s_block1.conv2
        want_hook = want_hook or bool(conv2._forward_hooks)
        with torch.cuda.device(x.device):
            params = list(self.parameters())
            if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                out, hook32 = _UNetFunction.apply(x, self, want_hook, *params)
                hook32 = hook32 if hook32.dim() == 5 else None
            else:
                out, hook32, tape = _unet_forward(self, x, self.training, False, want_hook)
                self._last_tape = tape
        if hook32 is not None and conv2._forward_hooks:
            view = hook32.permute(0, 4, 1, 2, 3)           # (N, 64, 96, 112, 96) over NDHWC memory, like the reference's hook sees
            for hook in list(conv2._forward_hooks.values()):
                hook(conv2, (None,), view)
        return out, hook32

    def forward(self, x):
        return self._run(x, False)[0]

    def forward_with_features(self, x):
        """-> (out (N,K,D,H,W), feat64 (N,96,112,96,64) fp32 NDHWC): the network output and the raw s_block1.conv2 output the
        reference's script hooks (image_features.py:58-60), both on the GPU."""
        return self._run(x, True)

    def roi_features(self, x, plan, atlas_shape=None):
        """The loop body of image_features.py:97-114 without leaving the GPU: forward, crop of the 64-channel map to the atlas
        grid, ROI mean pooling.  plan: multimodal_ad_b200.RoiPlan of the atlas; atlas_shape defaults to the input's (D,H,W).
        Returns (out (N,K,D,H,W), roi_feat (N,R,64))."""
        out, feat = self._run(x, True)
        shape = tuple(x.shape[2:]) if atlas_shape is None else tuple(atlas_shape)
        return out, plan.pool_channels_last(feat, shape)
