"""GPU parity of the Conv3d / BatchNorm / ReLU / MaxPool kernels (through the C-ABI) against torch fp32 on the same
bf16-rounded operands.  Tolerances: bf16 storage => 2^-8 relative per stored element (north star: 2e-2 in bf16)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def run(built_lib):
    from multimodal_ad_b200.models.resnet import _Run

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _Run(torch.device("cuda", 0))


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


CONV_CASES = [
    (1, 1, 8, 16, 64, 64, 1, 1, 0, 1), (1, 2, 8, 16, 128, 128, 1, 1, 0, 1), (1, 4, 8, 16, 64, 256, 1, 1, 0, 1),
    (2, 8, 8, 8, 64, 64, 3, 1, 1, 1), (1, 8, 16, 16, 128, 256, 3, 1, 2, 2), (1, 16, 16, 16, 64, 128, 3, 2, 1, 1),
    (1, 16, 16, 16, 64, 128, 1, 2, 0, 1), (1, 5, 7, 9, 64, 64, 3, 1, 1, 1), (1, 6, 11, 23, 64, 512, 3, 1, 4, 4),
    (2, 32, 32, 32, 64, 64, 3, 1, 1, 1), (1, 1, 1, 5000, 384, 64, 1, 1, 0, 1),
    # W-halo kernel on a ragged grid (layer1 of a 91x109x91 volume); two-sample-deep tiles with padding skips through the
    # CTA-pair kernel (dilation 4), the single-CTA BN=256 kernel (one voxel tile) and BN=128; multi-tile BN=64 K-step groups
    (3, 23, 28, 23, 64, 64, 3, 1, 1, 1), (2, 16, 16, 16, 64, 256, 3, 1, 4, 4), (4, 8, 8, 8, 128, 512, 3, 1, 2, 2),
    (2, 4, 4, 4, 64, 256, 3, 1, 4, 4), (2, 16, 16, 16, 64, 128, 3, 1, 4, 4), (1, 1, 1, 40000, 128, 64, 1, 1, 0, 1),
    # multi-slab W-halo kernel (Cin = 128 / 192 -> 64: the concatenated input of unet3d.py's s_block1.conv1), ragged grid
    (2, 24, 28, 31, 128, 64, 3, 1, 1, 1), (1, 32, 32, 45, 192, 64, 3, 1, 1, 1),
]


@pytest.mark.parametrize("cfg", CONV_CASES)
def test_conv3d_forward_and_statistics(cfg, run):
    n, d, h, w, cin, cout, k, stride, pad, dil = cfg
    g = torch.Generator(device="cuda").manual_seed(sum(cfg))
    x = torch.randn((n, d, h, w, cin), device="cuda", generator=g).to(torch.bfloat16)
    wt = (torch.randn((cout, k ** 3, cin), device="cuda", generator=g) / (k ** 1.5 * cin ** 0.5)).to(torch.bfloat16)
    y, part = run.conv(x, wt, cout, k, stride, pad, dil, True)
    ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), wt.float().reshape(cout, k, k, k, cin).permute(0, 4, 1, 2, 3),
                   stride=stride, padding=pad, dilation=dil).permute(0, 2, 3, 4, 1)
    assert not torch.isnan(y.float()).any()
    # every element within one bf16 rounding step of the fp32 result (+ fp32 accumulation-order slack)
    assert torch.all((y.float() - ref).abs() <= 2 ** -8 * ref.abs() + 1e-4)
    yb = y.float().reshape(-1, cout)
    s = part.sum(0)
    assert torch.allclose(s[:, 0], yb.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[:, 1], (yb * yb).sum(0), rtol=1e-4, atol=1e-2)


WG_CASES = [
    (1, 4, 4, 4, 64, 64, 1, 1, 0, 1), (1, 4, 8, 8, 128, 128, 1, 1, 0, 1), (2, 8, 8, 8, 64, 64, 3, 1, 1, 1),
    (1, 8, 8, 16, 128, 256, 3, 1, 2, 2), (1, 16, 16, 16, 64, 128, 3, 2, 1, 1), (1, 16, 16, 16, 64, 128, 1, 2, 0, 1),
    (1, 5, 7, 9, 256, 512, 3, 1, 4, 4), (1, 1, 1, 4096, 384, 64, 1, 1, 0, 1),
    # CTA-pair kernel: odd unit count (phantom unit), batch-deep chunks + padding skips, odd batch, single unit
    (2, 16, 16, 16, 128, 128, 3, 1, 1, 1), (2, 16, 16, 16, 256, 512, 3, 1, 4, 4), (3, 8, 8, 8, 128, 256, 3, 1, 1, 1),
    (1, 16, 16, 16, 128, 256, 1, 1, 0, 1),
    # halo kernel (64 -> 64): ragged, and the layer1 shape of a 91x109x91 volume
    (1, 5, 7, 9, 64, 64, 3, 1, 1, 1), (2, 23, 28, 23, 64, 64, 3, 1, 1, 1),
    # stride 2 on odd extents (phase convolutions with ragged phases), and on the layer2.0 shape of a 32^3 pooled grid
    (1, 9, 11, 13, 64, 128, 3, 2, 1, 1), (2, 32, 32, 32, 64, 128, 3, 2, 1, 1),
]


@pytest.mark.parametrize("cfg", WG_CASES)
def test_conv3d_wgrad_and_dgrad(cfg, run):
    from multimodal_ad_b200.models.resnet import _p

    n, d, h, w, cin, cout, k, stride, pad, dil = cfg
    g = torch.Generator(device="cuda").manual_seed(sum(cfg) + 1)
    x = torch.randn((n, d, h, w, cin), device="cuda", generator=g).to(torch.bfloat16)
    wt = torch.randn((cout, cin, k, k, k), device="cuda", generator=g) / (k ** 1.5 * cin ** 0.5)
    do, ho, wo = [(v + 2 * pad - dil * (k - 1) - 1) // stride + 1 for v in (d, h, w)]
    dy = torch.randn((n, do, ho, wo, cout), device="cuda", generator=g).to(torch.bfloat16)
    xr = x.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    wr = wt.to(torch.bfloat16).float().requires_grad_(True)
    F.conv3d(xr, wr, stride=stride, padding=pad, dilation=dil).backward(dy.float().permute(0, 4, 1, 2, 3))
    gw = torch.empty_like(wt)
    run.wgrad(x, dy, cout, k, stride, pad, dil, gw)
    assert _rel(gw, wr.grad) < 1e-4                     # fp32 accumulation of exact bf16 products: 1e-4 relative (fp32/tf32 bound)
    if cin == 384:
        return
    taps = k ** 3
    wf, wtr = run.empty((cout, taps, cin)), run.empty((cin, taps, cout))
    run.chk(run.lib.mmad_conv3d_prep_weights(_p(wt), _p(wf), _p(wtr), cout, cin, taps, run.stream), "prep")
    assert torch.equal(wf.float(), wt.to(torch.bfloat16).float().reshape(cout, cin, taps).permute(0, 2, 1))
    src = dy
    if stride == 2:
        src = run.empty((n, d, h, w, cout))
        run.chk(run.lib.mmad_upsample_zero2(_p(dy), _p(src), n, do, ho, wo, d, h, w, cout, run.stream), "up")
        chk = torch.zeros_like(src)
        chk[:, ::2, ::2, ::2][:, :do, :ho, :wo] = dy
        assert torch.equal(src, chk)
    dx, _ = run.conv(src, wtr, cin, k, 1, dil * (k - 1) - pad, dil, False)
    ref = xr.grad.permute(0, 2, 3, 4, 1)
    assert torch.all((dx.float() - ref).abs() <= 2 ** -8 * ref.abs() + 1e-4)
    if stride == 2 and k == 3 and pad == 1 and dil == 1:
        # the phase-decomposed data gradient (no zero insertion) gives the same tensor
        wph = run.empty((27 * cin * cout,))
        run.chk(run.lib.mmad_conv3d_prep_weights_s2(_p(wt), _p(wph), cout, cin, run.stream), "prep s2")
        dx2 = torch.full((n, d, h, w, cin), float("nan"), device="cuda", dtype=torch.bfloat16)
        run.chk(run.lib.mmad_conv3d_dgrad_s2_bf16(_p(dy), _p(wph), _p(dx2), n, d, h, w, cin, cout, run.stream), "dgrad s2")
        assert torch.all((dx2.float() - ref).abs() <= 2 ** -8 * ref.abs() + 1e-4)


@pytest.mark.parametrize("c,rows", [(64, 4099), (128, 513), (256, 70), (512, 1000)])
def test_batchnorm_forward_backward(c, rows, run):
    from multimodal_ad_b200.models.resnet import _p

    g = torch.Generator(device="cuda").manual_seed(c + rows)
    x = (torch.randn((1, 1, 1, rows, c), device="cuda", generator=g) * 1.7 + 0.3).to(torch.bfloat16)
    res = torch.randn((1, 1, 1, rows, c), device="cuda", generator=g).to(torch.bfloat16)
    bn = torch.nn.BatchNorm3d(c).cuda()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    # statistics partials the conv epilogue would have produced (two fake CTAs)
    xf = x.float().reshape(rows, c)
    half = rows // 2
    part = torch.stack([torch.stack([xf[:half].sum(0), (xf[:half] ** 2).sum(0)], 1),
                        torch.stack([xf[half:].sum(0), (xf[half:] ** 2).sum(0)], 1)]).contiguous()
    vec = run.bn_params(bn, part, rows, True)
    out = run.bn_apply(x, vec, relu=True, res=res)
    # torch reference on the same rounded operands
    xt = xf.clone().requires_grad_(True)
    gam, bet = bn.weight.detach().clone().requires_grad_(True), bn.bias.detach().clone().requires_grad_(True)
    rm, rv = rm0.clone(), rv0.clone()
    yt = F.batch_norm(xt.t().reshape(1, c, rows), rm, rv, gam, bet, True, 0.1, bn.eps).reshape(c, rows).t()
    ot = F.relu(yt + res.float().reshape(rows, c))
    assert torch.allclose(vec[0], xf.mean(0), rtol=1e-5, atol=1e-5)
    assert torch.allclose(vec[1], 1.0 / torch.sqrt(xf.var(0, unbiased=False) + bn.eps), rtol=1e-4)
    assert torch.allclose(bn.running_mean, rm, rtol=1e-5, atol=1e-6) and torch.allclose(bn.running_var, rv, rtol=1e-4, atol=1e-6)
    run.bump_tracked()                                                     # the per-layer counters are bumped together at the end of a pass
    assert int(bn.num_batches_tracked) == 1
    o = out.float().reshape(rows, c)
    assert torch.all((o - ot).abs() <= 2 ** -8 * ot.abs() + 1e-5)
    # backward: dy (+ dy2), ReLU mask from the stored output
    dy = torch.randn((1, 1, 1, rows, c), device="cuda", generator=g).to(torch.bfloat16)
    dy2 = torch.randn((1, 1, 1, rows, c), device="cuda", generator=g).to(torch.bfloat16)
    dx, gk, dgam, dbet = run.bn_bwd(dy, dy2, out, x, vec, bn.weight.detach(), True)
    gsum = (dy.float() + dy2.float()).reshape(rows, c) * (o > 0)
    gsum_r = gsum.to(torch.bfloat16).float()
    assert torch.equal(gk.float().reshape(rows, c), gsum_r)
    # autograd of the same graph fed with the rounded g (what pass 2 consumes)
    yt2 = F.batch_norm(xt.t().reshape(1, c, rows), rm0.clone(), rv0.clone(), gam, bet, True, 0.1, bn.eps).reshape(c, rows).t()
    yt2.backward(gsum_r)
    assert _rel(dgam, gam.grad) < 1e-4 and _rel(dbet, bet.grad) < 1e-4
    assert torch.all((dx.float().reshape(rows, c) - xt.grad).abs() <= 2 ** -8 * xt.grad.abs() + 3e-3 * xt.grad.abs().max())
    assert _rel(dx, xt.grad) < 5e-3
    # eval mode is an affine map
    bn.eval()
    vec_e = run.bn_params(bn, None, rows, False)
    oe = run.bn_apply(x, vec_e, relu=False)
    ref_e = F.batch_norm(xf.t().reshape(1, c, rows), bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.1, bn.eps)
    assert _rel(oe.reshape(rows, c), ref_e.reshape(c, rows).t()) < 4e-3


@pytest.mark.parametrize("shape", [(2, 8, 8, 8), (1, 7, 9, 11), (1, 16, 4, 6)])
def test_maxpool_forward_backward(shape, run):
    from multimodal_ad_b200.models.resnet import _p

    n, d, h, w = shape
    c = 64
    g = torch.Generator(device="cuda").manual_seed(d * h * w)
    x = torch.randn((n, d, h, w, c), device="cuda", generator=g).to(torch.bfloat16)
    do, ho, wo = (d - 1) // 2 + 1, (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = run.empty((n, do, ho, wo, c))
    idx = torch.empty((n, do, ho, wo, c), dtype=torch.uint8, device="cuda")
    run.chk(run.lib.mmad_maxpool3d_fwd(_p(x), _p(y), _p(idx), n, d, h, w, c, run.stream), "maxpool fwd")
    xt = x.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    yt = F.max_pool3d(xt, 3, 2, 1)
    assert torch.equal(y.float(), yt.permute(0, 2, 3, 4, 1))
    dy = torch.randn((n, do, ho, wo, c), device="cuda", generator=g).to(torch.bfloat16)
    dx = run.empty((n, d, h, w, c))
    run.chk(run.lib.mmad_maxpool3d_bwd(_p(dy), _p(idx), _p(dx), n, d, h, w, c, run.stream), "maxpool bwd")
    yt.backward(dy.float().permute(0, 4, 1, 2, 3))
    ref = xt.grad.permute(0, 2, 3, 4, 1)
    assert torch.all((dx.float() - ref).abs() <= 2 ** -7 * ref.abs() + 1e-6)      # sum of <= 8 bf16 terms, rounded once


@pytest.mark.parametrize("c,rows", [(64, 5000), (256, 333)])
def test_batchnorm_backward_four_pass_form_equals_six_pass_form(c, rows, run):
    """want_g=False + mask_from_x: pass 1 writes no masked gradient, pass 2 recomputes the ReLU mask from x - same dx, dgamma, dbeta
    as the form that stores g (bit for bit: the masked gradient of a single bf16 upstream tensor is exactly representable)."""
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = (torch.randn((1, 1, 1, rows, c), device="cuda", generator=g) * 1.3).to(torch.bfloat16)
    dy = torch.randn((1, 1, 1, rows, c), device="cuda", generator=g).to(torch.bfloat16)
    vec = torch.stack([torch.randn(c, device="cuda", generator=g) * 0.1, torch.rand(c, device="cuda", generator=g) + 0.5,
                       torch.rand(c, device="cuda", generator=g) + 0.5, torch.randn(c, device="cuda", generator=g) * 0.3]).contiguous()
    gamma = vec[2] / vec[1]
    for training in (True, False):
        dx6, _, dg6, db6 = run.bn_bwd(dy, None, None, x, vec, gamma, training, want_g=True, mask_from_x=True)
        dx4, g4, dg4, db4 = run.bn_bwd(dy, None, None, x, vec, gamma, training, want_g=False, mask_from_x=True)
        assert g4 is None and torch.equal(dx4, dx6) and torch.equal(dg4, dg6) and torch.equal(db4, db6)
    # without a ReLU (downsample branch): the unmasked form
    dxa, _, _, _ = run.bn_bwd(dy, None, None, x, vec, gamma, True, want_g=True)
    dxb, _, _, _ = run.bn_bwd(dy, None, None, x, vec, gamma, True, want_g=False)
    assert torch.equal(dxa, dxb)


def test_layout_transpose(run):
    from multimodal_ad_b200.models.resnet import _p

    x = torch.randn((2, 70, 5 * 6 * 7), device="cuda")
    y = run.empty((2, 5 * 6 * 7, 70))
    run.chk(run.lib.mmad_ncs_f32_to_nsc_bf16(_p(x), _p(y), 2, 70, 5 * 6 * 7, run.stream), "transpose")
    assert torch.equal(y, x.permute(0, 2, 1).to(torch.bfloat16))


@pytest.mark.parametrize("shape", [(2, 8, 8, 8), (1, 7, 9, 11)])
def test_fused_stem_matches_unfused_kernels(shape, run):
    """bn+relu+maxpool forward against the separate bn_apply / maxpool kernels; recomputed ReLU mask in the backward."""
    from multimodal_ad_b200.models.resnet import _p

    n, d, h, w = shape
    c = 64
    g = torch.Generator(device="cuda").manual_seed(d * h + w)
    c0 = (torch.randn((n, d, h, w, c), device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    rows = n * d * h * w
    vec = torch.stack([torch.randn(c, device="cuda", generator=g) * 0.1, torch.rand(c, device="cuda", generator=g) + 0.5,
                       torch.rand(c, device="cuda", generator=g) + 0.5, torch.randn(c, device="cuda", generator=g) * 0.3]).contiguous()
    gamma = vec[2] / vec[1]
    do, ho, wo = (d - 1) // 2 + 1, (h - 1) // 2 + 1, (w - 1) // 2 + 1
    # unfused reference path
    a0 = run.bn_apply(c0, vec, relu=True)
    p_ref = run.empty((n, do, ho, wo, c))
    i_ref = torch.empty((n, do, ho, wo, c), dtype=torch.uint8, device="cuda")
    run.chk(run.lib.mmad_maxpool3d_fwd(_p(a0), _p(p_ref), _p(i_ref), n, d, h, w, c, run.stream), "mp")
    p0 = run.empty((n, do, ho, wo, c))
    idx = torch.empty((n, do, ho, wo, c), dtype=torch.uint8, device="cuda")
    run.chk(run.lib.mmad_stem_bn_relu_maxpool_fwd(_p(c0), _p(vec[2]), _p(vec[3]), _p(p0), _p(idx), n, d, h, w, c, run.stream), "fused fwd")
    assert torch.equal(p0, p_ref) and torch.equal(idx, i_ref)
    # backward: the ReLU mask recomputed from c0 equals the mask read from the stored activation
    da0 = torch.randn((n, d, h, w, c), device="cuda", generator=g).to(torch.bfloat16)
    dx_a, g_a, dg_a, db_a = run.bn_bwd(da0, None, a0, c0, vec, gamma, True)
    dx_b, g_b, dg_b, db_b = run.bn_bwd(da0, None, None, c0, vec, gamma, True, mask_from_x=True)
    assert torch.equal(g_a, g_b) and torch.equal(dx_a, dx_b) and torch.equal(dg_a, dg_b) and torch.equal(db_a, db_b)


@pytest.mark.parametrize("shape", [(1, 16, 16, 16), (2, 8, 8, 16), (1, 13, 11, 19), (2, 32, 32, 32), (1, 40, 31, 27)])
def test_stem_space_to_depth_forward_statistics_wgrad(shape, run):
    """conv1 (resnet.py:126-132) through mmad_stem_s2d_*: forward within one bf16 rounding step of torch fp32 on the same
    bf16 operands, BatchNorm partials exact to fp32 summation, weight gradient to 1e-4 (fp32 accumulation)."""
    from multimodal_ad_b200.models.resnet import _p

    n, d, h, w = shape
    lib = run.lib
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    x = torch.randn((n, 1, d, h, w), device="cuda", generator=g)
    wt = torch.randn((64, 1, 7, 7, 7), device="cuda", generator=g) / 343 ** 0.5
    do, ho, wo = (d - 1) // 2 + 1, (h - 1) // 2 + 1, (w - 1) // 2 + 1
    xs = run.empty((lib.mmad_stem_s2d_elems(n, d, h, w),))
    wk = run.empty((64, 512))
    y = run.empty((n, do, ho, wo, 64))
    part = run.empty((lib.mmad_stem_s2d_stats_partials(n, d, h, w), 64, 2), torch.float32)
    run.chk(lib.mmad_stem_s2d_pack(_p(x), _p(xs), n, d, h, w, run.stream), "pack")
    run.chk(lib.mmad_stem_s2d_prep_weights(_p(wt), _p(wk), run.stream), "prep")
    run.chk(lib.mmad_stem_s2d_fwd(_p(xs), _p(wk), _p(y), _p(part), n, d, h, w, run.stream), "fwd")
    # the packed tensors hold exactly the bf16 roundings of the operands
    xp = F.pad(x[:, 0], (3, 2 * (wo + 3) - w - 3, 3, 2 * (ho + 3) - h - 3, 3, 2 * (do + 3) - d - 3)).to(torch.bfloat16)
    xp = xp.view(n, do + 3, 2, ho + 3, 2, wo + 3, 2).permute(0, 1, 3, 5, 2, 4, 6).reshape(-1)
    assert torch.equal(xs, xp)
    xr = x.to(torch.bfloat16).float()
    wr = wt.to(torch.bfloat16).float().requires_grad_(True)
    ref = F.conv3d(xr, wr, stride=2, padding=3)
    yf = y.float().permute(0, 4, 1, 2, 3)
    assert torch.all((yf - ref).abs() <= 2 ** -8 * ref.abs() + 1e-4)
    st = part.double().sum(0)
    assert torch.allclose(st[:, 0], yf.double().sum((0, 2, 3, 4)), rtol=1e-5, atol=1e-3)
    assert torch.allclose(st[:, 1], (yf.double() ** 2).sum((0, 2, 3, 4)), rtol=1e-5)
    dy = torch.randn((n, do, ho, wo, 64), device="cuda", generator=g).to(torch.bfloat16)
    gw = torch.empty_like(wt)
    run.stem_wgrad(xs, dy, n, d, h, w, gw)
    ref.backward(dy.float().permute(0, 4, 1, 2, 3))
    assert _rel(gw, wr.grad) < 1e-4


def test_conv_kernel_variants_behind_environment_knobs():
    """The tuning knobs select other code paths of the same convolution (read once per process, hence a subprocess each): the
    single-CTA W-halo kernel, the static tile stride, streamed instead of resident weights in the CTA-pair W-halo kernel.  All must agree with torch."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import torch, torch.nn.functional as F
from multimodal_ad_b200.models.resnet import _Run
r = _Run(torch.device("cuda", 0))
g = torch.Generator(device="cuda").manual_seed(7)
for (n, d, h, w, cin) in ((2, 24, 28, 31, 64), (1, 32, 32, 45, 192)):
    x = torch.randn((n, d, h, w, cin), device="cuda", generator=g).to(torch.bfloat16)
    wt = (torch.randn((64, 27, cin), device="cuda", generator=g) / (27 * cin) ** 0.5).to(torch.bfloat16)
    y, part = r.conv(x, wt, 64, 3, 1, 1, 1, True)
    ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), wt.float().reshape(64, 3, 3, 3, cin).permute(0, 4, 1, 2, 3), padding=1).permute(0, 2, 3, 4, 1)
    assert torch.all((y.float() - ref).abs() <= 2 ** -8 * ref.abs() + 1e-4)
    assert torch.allclose(part.sum(0)[:, 0], y.float().reshape(-1, 64).sum(0), rtol=1e-4, atol=1e-2)
print("ok")
'''
    for knobs in ({"MMAD_CONV_HALO_PAIR": "0"}, {"MMAD_CONV_DYN": "0"}, {"MMAD_CONV_WRES": "0"}, {"MMAD_CONV_HALO": "0"}):
        env = dict(os.environ, PYTHONPATH=root, **knobs)
        res = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0 and "ok" in res.stdout, (knobs, res.stdout[-500:], res.stderr[-1500:])
