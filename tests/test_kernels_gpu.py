"""Kernel-level parity on the B200, through the C-ABI: each CUDA kernel against a plain PyTorch fp32/fp64
reference of the same op on the same (bf16-representable) inputs. Tolerances: the only error left is the bf16
rounding of the output (2^-9 relative per element, ~1.7e-3 in the 2-norm) or, for fp32 outputs, accumulation
order (1e-5)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gpu_util import L, conv_ref, gen, pack, ptr, relerr, st  # noqa: E402
from syntex_b200 import _lib  # noqa: E402

BF16_OUT = 2.5e-3


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


CONV_CASES = [  # n, H, W, cin, cout, force_bn
    (1, 16, 16, 64, 64, 0), (1, 16, 16, 64, 128, 128), (1, 32, 32, 128, 128, 64), (2, 8, 8, 128, 256, 0),
    (1, 20, 12, 64, 64, 0),      # ragged: H, W not multiples of the tile
    (3, 4, 4, 256, 512, 64),     # several images inside one 128-row tile
    (1, 2, 2, 512, 512, 0),      # tiny map (64^2 input at pool4 level is 4x4; this is smaller still)
    (1, 64, 64, 256, 256, 128), (1, 32, 32, 512, 512, 0), (1, 256, 256, 64, 64, 0), (5, 16, 16, 64, 64, 0)]


@pytest.mark.parametrize("n,H,W,cin,cout,bn", CONV_CASES)
def test_conv_forward_and_dgrad(n, H, W, cin, cout, bn):
    g = gen(0)
    x = torch.randn((n, H, W, cin), device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn((3, 3, cin, cout), device="cuda", generator=g) / np.sqrt(9 * cin)).float()
    b = torch.randn((cout,), device="cuda", generator=g).float()
    wf, _, wd = pack(w)
    wq = w.to(torch.bfloat16).float()
    out = torch.full((n, H, W, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(L().stx_conv_bf16(x.data_ptr(), n, H, W, cin, wf.data_ptr(), cout, 9, b.data_ptr(), None, None, 1,
                                 out.data_ptr(), bn, st()))
    ref = conv_ref(x.float(), wq, b).relu()
    assert torch.isfinite(out.float()).all()
    assert relerr(out.float(), ref) < BF16_OUT
    # input gradient = convolution with the rotated, transposed pack == conv_transpose2d
    gy = torch.randn((n, H, W, cout), device="cuda", generator=g).to(torch.bfloat16)
    dx = torch.full((n, H, W, cin), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(L().stx_conv_bf16(gy.data_ptr(), n, H, W, cout, wd.data_ptr(), cin, 9, None, None, None, 0,
                                 dx.data_ptr(), bn, st()))
    refd = F.conv_transpose2d(gy.float().permute(0, 3, 1, 2), wq.permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    assert relerr(dx.float(), refd) < BF16_OUT


VARIANTS = [0, 64, 128, 256, 1064, 2064, 1128, 2128, 2256, 100064, 100128]   # bn + 1000*mt + 100000*per_tap


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("n,H,W,cin,cout", [(2, 40, 24, 64, 256), (1, 16, 8, 128, 256), (3, 9, 17, 64, 256)])
def test_conv_kernel_variants_agree_with_reference(n, H, W, cin, cout, variant):
    """Every tile configuration of both kernels (halo-patch / per-tap) on ragged and multi-image shapes."""
    g = gen(10)
    x = torch.randn((n, H, W, cin), device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn((3, 3, cin, cout), device="cuda", generator=g) / np.sqrt(9 * cin)).float()
    b = torch.randn((cout,), device="cuda", generator=g).float()
    msk = torch.randn((n, H, W, cout), device="cuda", generator=g).to(torch.bfloat16)
    wf, _, _ = pack(w)
    out = torch.full((n, H, W, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(L().stx_conv_bf16(x.data_ptr(), n, H, W, cin, wf.data_ptr(), cout, 9, b.data_ptr(), None,
                                 msk.data_ptr(), 0, out.data_ptr(), variant, st()))
    ref = (conv_ref(x.float(), w.to(torch.bfloat16).float(), b)) * (msk.float() > 0)
    assert torch.isfinite(out.float()).all()
    assert relerr(out.float(), ref) < BF16_OUT


@pytest.mark.parametrize("n,H,W,cin,cout", [(4, 72, 40, 64, 256),      # ragged, 25 tiles per image: a pair spans two images
                                            (6, 64, 64, 256, 128), (2, 128, 96, 128, 128), (37, 16, 16, 512, 512),
                                            (4, 104, 120, 64, 64), (2, 128, 128, 128, 64)])   # N tile 64, two pairs per TPC
def test_conv_cta_pairs_match_single_cta_bitwise(n, H, W, cin, cout):
    """CTA-pair instances (tcgen05 cta_group::2, variant + 1000000): two 16x8 tiles against one weight tile, half of it
    staged per SM. Same products accumulated in the same order: bit-identical to the one-CTA instance."""
    g = gen(12)
    x = torch.randn((n, H, W, cin), device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn((3, 3, cin, cout), device="cuda", generator=g) / np.sqrt(9 * cin)).float()
    b = torch.randn((cout,), device="cuda", generator=g).float()
    msk = torch.randn((n, H, W, cout), device="cuda", generator=g).to(torch.bfloat16)
    wf, _, _ = pack(w)
    outs = []
    # (cout a multiple of 256: the pair instance takes N tiles of 256 - 128 weight rows per SM)
    for variant in (1000000, 1128 if cout % 128 == 0 else 1064):
        o = torch.full((n, H, W, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
        _lib.check(L().stx_conv_bf16(x.data_ptr(), n, H, W, cin, wf.data_ptr(), cout, 9, b.data_ptr(), None,
                                     msk.data_ptr(), 0, o.data_ptr(), variant, st()))
        outs.append(o)
    torch.cuda.synchronize()
    assert torch.isfinite(outs[0].float()).all()
    assert torch.equal(outs[0], outs[1])
    ref = (conv_ref(x.float(), w.to(torch.bfloat16).float(), b)) * (msk.float() > 0)
    assert relerr(outs[0].float(), ref) < BF16_OUT


@pytest.mark.parametrize("n,H,W,cin,cout", [(1, 256, 256, 64, 64), (3, 104, 120, 64, 64), (2, 128, 128, 128, 64),
                                            (1, 250, 199, 128, 64)])
def test_conv_resident_weights_match_the_ring_bitwise(n, H, W, cin, cout):
    """Resident-weight instances (conv_patch.cu, RES; opt-in through stx_debug_set(8, 1) since they measured no
    faster than the ring) against the streaming ring: same MMA order, so identical bits."""
    g = gen(21)
    x = torch.randn((n, H, W, cin), device="cuda", generator=g)
    xh = x.to(torch.bfloat16)
    xl = (x - xh.float()).to(torch.bfloat16)
    w = (torch.randn((3, 3, cin, cout), device="cuda", generator=g) / np.sqrt(9 * cin)).float()
    b = torch.randn((cout,), device="cuda", generator=g).float()
    msk = torch.randn((n, H, W, cout), device="cuda", generator=g).to(torch.bfloat16)
    wf, wl, _ = pack(w, split=True)
    outs = []
    for variant in (0, 1064):
        o = torch.full((n, H, W, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
        L().stx_debug_set(8, 1 if variant == 0 else 0)
        try:
            _lib.check(L().stx_conv_bf16(xh.data_ptr(), n, H, W, cin, wf.data_ptr(), cout, 9, b.data_ptr(), None,
                                         msk.data_ptr(), 0, o.data_ptr(), variant, st()))
        finally:
            L().stx_debug_set(8, 0)
        outs.append(o)
    assert torch.equal(outs[0], outs[1])
    ref = (conv_ref(xh.float(), w.to(torch.bfloat16).float(), b)) * (msk.float() > 0)
    assert relerr(outs[0].float(), ref) < BF16_OUT
    if cin == 64:       # the bf16x3 forward has a resident instance for 64 -> 64 only
        res = []
        for variant in (0, 64):
            oh = torch.full((n, H, W, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
            ol = torch.full_like(oh, float("nan"))
            L().stx_debug_set(8, 1 if variant == 0 else 0)
            try:
                _lib.check(L().stx_conv_bf16x3(xh.data_ptr(), xl.data_ptr(), n, H, W, cin, wf.data_ptr(),
                                               wl.data_ptr(), cout, b.data_ptr(), 1, oh.data_ptr(), ol.data_ptr(),
                                               variant, st()))
            finally:
                L().stx_debug_set(8, 0)
            res.append((oh, ol))
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
        ref3 = conv_ref(x.double(), w.double(), b.double()).relu()
        assert relerr(res[0][0].double() + res[0][1].double(), ref3) < 1e-4


def test_conv_epilogue_addend_mask_inplace():
    n, H, W, cin, cout = 2, 24, 16, 128, 64
    g = gen(1)
    x = torch.randn((n, H, W, cin), device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn((3, 3, cin, cout), device="cuda", generator=g) / np.sqrt(9 * cin)).float()
    add = torch.randn((n, H, W, cout), device="cuda", generator=g).to(torch.bfloat16)
    msk = torch.randn((n, H, W, cout), device="cuda", generator=g).to(torch.bfloat16)
    wf, _, _ = pack(w)
    ref = (conv_ref(x.float(), w.to(torch.bfloat16).float()) + add.float()) * (msk.float() > 0)
    out = add.clone()                     # accumulate in place: addend aliases out
    _lib.check(L().stx_conv_bf16(x.data_ptr(), n, H, W, cin, wf.data_ptr(), cout, 9, None, out.data_ptr(),
                                 msk.data_ptr(), 0, out.data_ptr(), 0, st()))
    assert relerr(out.float(), ref) < BF16_OUT
    assert torch.equal(out.float() == 0, ref == 0) or (out.float()[msk.float() <= 0] == 0).all()


def test_conv_1x1_is_a_plain_gemm():
    n, H, W, C = 1, 32, 32, 128
    g = gen(2)
    x = torch.randn((n, H, W, C), device="cuda", generator=g).to(torch.bfloat16)
    d = torch.randn((1, C, C), device="cuda", generator=g).to(torch.bfloat16)     # [tap=1][cout][cin]
    out = torch.empty((n, H, W, C), dtype=torch.bfloat16, device="cuda")
    _lib.check(L().stx_conv_bf16(x.data_ptr(), n, H, W, C, d.data_ptr(), C, 1, None, None, None, 0, out.data_ptr(),
                                 0, st()))
    ref = x.float().reshape(-1, C) @ d[0].float().t()
    assert relerr(out.float().reshape(-1, C), ref) < BF16_OUT


@pytest.mark.parametrize("n,H,W,cin,cout,bn", [(1, 16, 16, 64, 64, 0), (1, 32, 32, 128, 128, 128),
                                               (2, 8, 8, 128, 256, 64), (1, 64, 64, 256, 256, 0),
                                               (1, 32, 32, 512, 512, 0), (1, 20, 12, 64, 128, 0),
                                               (1, 32, 32, 128, 128, 100128), (2, 4, 4, 128, 128, 0)])
def test_conv_bf16x3_forward(n, H, W, cin, cout, bn):
    g = gen(3)
    x = torch.randn((n, H, W, cin), device="cuda", generator=g).relu()
    xh = x.to(torch.bfloat16)
    xl = (x - xh.float()).to(torch.bfloat16)
    w = (torch.randn((3, 3, cin, cout), device="cuda", generator=g) / np.sqrt(9 * cin)).float()
    b = torch.randn((cout,), device="cuda", generator=g).float()
    wh, wl, _ = pack(w, split=True)
    oh = torch.full((n, H, W, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    ol = torch.full_like(oh, float("nan"))
    _lib.check(L().stx_conv_bf16x3(xh.data_ptr(), xl.data_ptr(), n, H, W, cin, wh.data_ptr(), wl.data_ptr(), cout,
                                   b.data_ptr(), 1, oh.data_ptr(), ol.data_ptr(), bn, st()))
    ref = conv_ref(x.double(), w.double(), b.double()).relu()
    assert relerr(oh.float() + ol.float(), ref) < 3e-5          # ~16 mantissa bits end to end
    assert relerr(oh.float(), ref) < BF16_OUT
    assert ((oh.float() > 0) == (ref > 0)).float().mean() > 0.9999   # ReLU signs: what the split is for


@pytest.mark.parametrize("pre", [0, 1])
def test_conv1_1_forward_backward(pre):
    n, H, W = 2, 32, 48
    g = gen(4)
    img = torch.rand((n, H, W, 3), device="cuda", generator=g) if not pre else \
        torch.randn((n, H, W, 3), device="cuda", generator=g)
    w = (torch.randn((3, 3, 3, 64), device="cuda", generator=g) * 0.02).float()
    b = (torch.randn((64,), device="cuda", generator=g) * 0.1).float()
    mean = (ctypes.c_float * 3)(123.68, 116.779, 103.939)
    hi = torch.empty((n, H, W, 64), dtype=torch.bfloat16, device="cuda")
    lo = torch.empty_like(hi)
    _lib.check(L().stx_conv1_1_fwd(img.data_ptr(), n, H, W, pre, mean, w.data_ptr(), b.data_ptr(), hi.data_ptr(),
                                   lo.data_ptr(), st()))
    xi = img.double().clone().requires_grad_(True)
    xx = torch.sigmoid(xi) if pre else xi
    xp = xx * 255.0 - torch.tensor([123.68, 116.779, 103.939], device="cuda", dtype=torch.float64)
    ref = conv_ref(xp, w.double(), b.double()).relu()
    assert relerr(hi.float() + lo.float(), ref.detach()) < 2e-5
    assert relerr(hi.float(), ref.detach()) < BF16_OUT
    gq = torch.randn((n, H, W, 64), device="cuda", generator=g).to(torch.bfloat16)
    gm = (gq.float() * (ref.detach() > 0)).to(torch.bfloat16)
    gout = torch.empty((n, H, W, 3), dtype=torch.float32, device="cuda")
    _lib.check(L().stx_conv1_1_bwd(gm.data_ptr(), img.data_ptr(), n, H, W, pre, w.data_ptr(), gout.data_ptr(), st()))
    (gref,) = torch.autograd.grad(ref, xi, gm.double())
    assert relerr(gout, gref) < 1e-3


def test_avgpool_forward_backward():
    n, H, W, C = 2, 32, 48, 128
    g = gen(5)
    x = torch.randn((n, H, W, C), device="cuda", generator=g)
    xh = x.to(torch.bfloat16)
    xl = (x - xh.float()).to(torch.bfloat16)
    oh = torch.empty((n, H // 2, W // 2, C), dtype=torch.bfloat16, device="cuda")
    ol = torch.empty_like(oh)
    _lib.check(L().stx_avgpool2_fwd(xh.data_ptr(), xl.data_ptr(), n, H, W, C, oh.data_ptr(), ol.data_ptr(), st()))
    ref = F.avg_pool2d((xh.double() + xl.double()).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert relerr(oh.float() + ol.float(), ref) < 2e-5
    o1 = torch.empty_like(oh)
    _lib.check(L().stx_avgpool2_fwd(xh.data_ptr(), None, n, H, W, C, o1.data_ptr(), None, st()))
    assert relerr(o1.float(), F.avg_pool2d(xh.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)) < BF16_OUT
    gi = torch.randn((n, H // 2, W // 2, C), device="cuda", generator=g).to(torch.bfloat16)
    go = torch.empty((n, H, W, C), dtype=torch.bfloat16, device="cuda")
    _lib.check(L().stx_avgpool2_bwd(gi.data_ptr(), xh.data_ptr(), n, H, W, C, go.data_ptr(), st()))
    refb = (F.interpolate(gi.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
            * 0.25) * (xh.float() > 0)
    assert torch.equal(go.float(), refb)       # x0.25 and masking are exact in bf16
    _lib.check(L().stx_avgpool2_bwd(gi.data_ptr(), None, n, H, W, C, go.data_ptr(), st()))
    up = F.interpolate(gi.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1) * 0.25
    assert torch.equal(go.float(), up)         # no mask: plain AvgPoolGrad


GRAM_CASES = [(1, 256, 64), (1, 4096, 64), (2, 1024, 128), (1, 1024, 256), (1, 256, 512), (2, 196, 512),
              (1, 65536, 64), (1, 16, 512), (3, 100, 256), (1, 1, 64)]


@pytest.mark.parametrize("n,HW,C", GRAM_CASES)
def test_gram(n, HW, C):
    g = gen(6)
    f = torch.randn((n, HW, C), device="cuda", generator=g).relu().to(torch.bfloat16)
    ws = torch.empty(max(L().stx_gram_workspace_bytes(n, HW, C) // 4, 1), dtype=torch.float32, device="cuda")
    out = torch.full((n, C, C), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(L().stx_gram_bf16(f.data_ptr(), n, HW, C, ctypes.c_float(1e-6), out.data_ptr(), ws.data_ptr(), st()))
    ref = torch.einsum("nkc,nkd->ncd", f.double(), f.double()) / float(np.float32(HW) + np.float32(1e-6))
    assert relerr(out, ref) < 2e-6               # bf16 products are exact in fp32; only the summation order differs
    assert torch.equal(out, out.transpose(1, 2))  # mirrored from the upper triangle: exactly symmetric
    out2 = torch.empty_like(out)
    _lib.check(L().stx_gram_bf16(f.data_ptr(), n, HW, C, ctypes.c_float(1e-6), out2.data_ptr(), ws.data_ptr(), st()))
    assert torch.equal(out, out2)                 # fixed-order split-K reduction: bit-reproducible


def test_invalid_arguments_fail_loudly():
    x = torch.zeros((1, 8, 8, 48), dtype=torch.bfloat16, device="cuda")
    w = torch.zeros((9, 64, 48), dtype=torch.bfloat16, device="cuda")
    o = torch.zeros((1, 8, 8, 64), dtype=torch.bfloat16, device="cuda")
    rc = L().stx_conv_bf16(x.data_ptr(), 1, 8, 8, 48, w.data_ptr(), 64, 9, None, None, None, 0, o.data_ptr(), 0, st())
    assert rc == -1 and b"multiples of 64" in L().stx_last_error()
    rc = L().stx_conv_bf16(x.data_ptr(), 1, 8, 8, 64, w.data_ptr(), 64, 5, None, None, None, 0, o.data_ptr(), 0, st())
    assert rc == -1
    rc = L().stx_avgpool2_fwd(x.data_ptr(), None, 1, 7, 8, 64, o.data_ptr(), None, st())
    assert rc == -1
    with pytest.raises(_lib.SyntexLibraryError):
        _lib.check(L().stx_gram_bf16(x.data_ptr(), 1, 64, 96, ctypes.c_float(1e-6), o.data_ptr(), o.data_ptr(), st()))


def test_cta_pair_mma_probe():
    """tcgen05 cta_group::2: one cluster of two CTAs computes a 256 x 128 tile - each CTA holds 128 rows of A and HALF
    of B, the leader issues M = 256 MMAs, each CTA reads its rows of D from its own tensor memory."""
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn((256, 64), device="cuda", generator=g).to(torch.bfloat16).contiguous()
    b = torch.randn((128, 64), device="cuda", generator=g).to(torch.bfloat16).contiguous()
    out = torch.full((256, 128), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(L().stx_debug_pair_probe(a.data_ptr(), b.data_ptr(), out.data_ptr(), st()))
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    assert torch.isfinite(out).all()
    assert relerr(out, ref) < 1e-6
