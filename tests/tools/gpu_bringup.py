"""Bring-up diagnostics for the CUDA kernels on a B200 (prints errors per kernel; not a test).
Usage: python tests/tools/gpu_bringup.py [stage ...]   stages: conv gram elem plan lbfgs
"""
import ctypes
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from syntex_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = "cuda"


def st():
    return _lib.stream_ptr()


def relerr(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max())


def pack(w_hwio):
    cin, cout = w_hwio.shape[2], w_hwio.shape[3]
    wf = torch.empty((9, cout, cin), dtype=torch.bfloat16, device=dev)
    wd = torch.empty((9, cin, cout), dtype=torch.bfloat16, device=dev)
    _lib.check(L.stx_pack_conv_weights(w_hwio.contiguous().data_ptr(), cin, cout, wf.data_ptr(), wd.data_ptr(), st()))
    return wf, wd


def conv_case(n, H, W, cin, cout, bn=0, bias=True, relu=True, addend=False, mask=False, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn((n, H, W, cin), device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn((3, 3, cin, cout), device=dev, generator=g) / np.sqrt(9 * cin)).float()
    b = torch.randn((cout,), device=dev, generator=g).float() if bias else None
    add = torch.randn((n, H, W, cout), device=dev, generator=g).to(torch.bfloat16) if addend else None
    msk = torch.randn((n, H, W, cout), device=dev, generator=g).to(torch.bfloat16) if mask else None
    wf, wd = pack(w)
    out = torch.full((n, H, W, cout), float("nan"), dtype=torch.bfloat16, device=dev)
    rc = L.stx_conv_bf16(x.data_ptr(), n, H, W, cin, wf.data_ptr(), cout, 9, b.data_ptr() if bias else None,
                         add.data_ptr() if addend else None, msk.data_ptr() if mask else None, int(relu),
                         out.data_ptr(), bn, st())
    _lib.check(rc)
    torch.cuda.synchronize()
    wq = w.to(torch.bfloat16).float()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wq.permute(3, 2, 0, 1), b, padding=1).permute(0, 2, 3, 1)
    if addend:
        ref = ref + add.float()
    if relu:
        ref = ref.relu()
    if mask:
        ref = ref * (msk.float() > 0)
    e = relerr(out.float(), ref)
    # dgrad: conv with rotated/transposed weights == conv_transpose
    gin = torch.randn((n, H, W, cout), device=dev, generator=g).to(torch.bfloat16)
    dx = torch.full((n, H, W, cin), float("nan"), dtype=torch.bfloat16, device=dev)
    _lib.check(L.stx_conv_bf16(gin.data_ptr(), n, H, W, cout, wd.data_ptr(), cin, 9, None, None, None, 0,
                               dx.data_ptr(), bn, st()))
    torch.cuda.synchronize()
    refd = F.conv_transpose2d(gin.float().permute(0, 3, 1, 2), wq.permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    e2 = relerr(dx.float(), refd)
    return e, e2


def stage_conv():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cases = [(1, 16, 16, 64, 64, 0), (1, 16, 16, 64, 128, 128), (1, 32, 32, 128, 128, 64), (2, 8, 8, 128, 256, 0),
             (1, 20, 12, 64, 64, 0), (3, 4, 4, 256, 512, 64), (1, 64, 64, 256, 256, 128), (1, 256, 256, 64, 64, 0),
             (1, 32, 32, 512, 512, 0)]
    for (n, H, W, ci, co, bn) in cases:
        for opts in ({"bias": True, "relu": True}, {"bias": False, "relu": False, "addend": True, "mask": True}):
            try:
                e, e2 = conv_case(n, H, W, ci, co, bn, **opts)
                print("conv n%d %dx%d %d->%d bn%d %s: fwd rel %.2e max %.2e | dgrad rel %.2e max %.2e" % (
                    n, H, W, ci, co, bn, "epi" if "addend" in opts else "std", e[0], e[1], e2[0], e2[1]), flush=True)
            except Exception as ex:
                print("conv case", (n, H, W, ci, co, bn), "FAILED:", ex, flush=True)
                raise


def split_case(n, H, W, cin, cout, bn=0, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn((n, H, W, cin), device=dev, generator=g).relu()
    xh = x.to(torch.bfloat16); xl = (x - xh.float()).to(torch.bfloat16)
    w = (torch.randn((3, 3, cin, cout), device=dev, generator=g) / np.sqrt(9 * cin)).float()
    b = torch.randn((cout,), device=dev, generator=g).float()
    wh = torch.empty((9, cout, cin), dtype=torch.bfloat16, device=dev); wl = torch.empty_like(wh)
    _lib.check(L.stx_pack_conv_weights_split(w.contiguous().data_ptr(), cin, cout, wh.data_ptr(), wl.data_ptr(), None, st()))
    oh = torch.full((n, H, W, cout), float("nan"), dtype=torch.bfloat16, device=dev); ol = torch.full_like(oh, float("nan"))
    _lib.check(L.stx_conv_bf16x3(xh.data_ptr(), xl.data_ptr(), n, H, W, cin, wh.data_ptr(), wl.data_ptr(), cout, b.data_ptr(), 1,
                                 oh.data_ptr(), ol.data_ptr(), bn, st()))
    torch.cuda.synchronize()
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), b.double(), padding=1).relu().permute(0, 2, 3, 1)
    return relerr(oh.float() + ol.float(), ref), relerr(oh.float(), ref)


def stage_split():
    for (n, H, W, ci, co, bn) in [(1, 16, 16, 64, 64, 0), (1, 32, 32, 128, 128, 128), (2, 8, 8, 128, 256, 64),
                                  (1, 64, 64, 256, 256, 0), (1, 32, 32, 512, 512, 0), (1, 256, 256, 64, 64, 0)]:
        e, eh = split_case(n, H, W, ci, co, bn)
        print("split conv n%d %dx%d %d->%d bn%d: hi+lo rel %.2e max %.2e | hi only rel %.2e" % (n, H, W, ci, co, bn, e[0], e[1], eh[0]), flush=True)


def gram_case(n, HW, C, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    f = torch.randn((n, HW, C), device=dev, generator=g).relu().to(torch.bfloat16)
    ws = torch.empty(max(L.stx_gram_workspace_bytes(n, HW, C) // 4, 1), dtype=torch.float32, device=dev)
    out = torch.full((n, C, C), float("nan"), dtype=torch.float32, device=dev)
    _lib.check(L.stx_gram_bf16(f.data_ptr(), n, HW, C, ctypes.c_float(1e-6), out.data_ptr(), ws.data_ptr(), st()))
    torch.cuda.synchronize()
    ref = torch.einsum("nkc,nkd->ncd", f.double(), f.double()) / HW
    return relerr(out, ref)


def stage_gram():
    for (n, HW, C) in [(1, 256, 64), (1, 4096, 64), (2, 1024, 128), (1, 1024, 256), (1, 256, 512), (2, 196, 512),
                       (1, 65536, 64), (1, 16, 512)]:
        e = gram_case(n, HW, C)
        print("gram n%d HW%d C%d: rel %.2e max %.2e" % (n, HW, C, e[0], e[1]), flush=True)


def stage_elem():
    g = torch.Generator(device=dev).manual_seed(1)
    n, H, W = 2, 32, 48
    img = torch.rand((n, H, W, 3), device=dev, generator=g)
    w = (torch.randn((3, 3, 3, 64), device=dev, generator=g) * 0.02).float()
    b = torch.randn((64,), device=dev, generator=g).float() * 0.1
    mean = (ctypes.c_float * 3)(123.68, 116.779, 103.939)
    for pre in (0, 1):
        out = torch.empty((n, H, W, 64), dtype=torch.bfloat16, device=dev)
        lo = torch.empty((n, H, W, 64), dtype=torch.bfloat16, device=dev)
        _lib.check(L.stx_conv1_1_fwd(img.data_ptr(), n, H, W, pre, mean, w.data_ptr(), b.data_ptr(), out.data_ptr(), lo.data_ptr(), st()))
        xi = img.clone().requires_grad_(True)
        xx = torch.sigmoid(xi) if pre else xi
        xp = xx * 255.0 - torch.tensor([123.68, 116.779, 103.939], device=dev)
        ref = F.conv2d(xp.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=1).relu().permute(0, 2, 3, 1)
        print("conv1_1 fwd pre=%d:" % pre, relerr(out.float(), ref.detach()), "hi+lo:", relerr(out.float() + lo.float(), ref.detach()), flush=True)
        gq = torch.randn((n, H, W, 64), device=dev, generator=g).to(torch.bfloat16)
        gm = (gq.float() * (ref.detach() > 0)).to(torch.bfloat16)
        gout = torch.empty((n, H, W, 3), dtype=torch.float32, device=dev)
        _lib.check(L.stx_conv1_1_bwd(gm.data_ptr(), img.data_ptr(), n, H, W, pre, w.data_ptr(), gout.data_ptr(), st()))
        (gref,) = torch.autograd.grad(ref, xi, gm.float())
        print("conv1_1 bwd pre=%d:" % pre, relerr(gout, gref), flush=True)
    x = torch.randn((n, H, W, 128), device=dev, generator=g).to(torch.bfloat16)
    o = torch.empty((n, H // 2, W // 2, 128), dtype=torch.bfloat16, device=dev)
    _lib.check(L.stx_avgpool2_fwd(x.data_ptr(), None, n, H, W, 128, o.data_ptr(), None, st()))
    ref = F.avg_pool2d(x.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    print("avgpool fwd:", relerr(o.float(), ref), flush=True)
    gi = torch.randn((n, H // 2, W // 2, 128), device=dev, generator=g).to(torch.bfloat16)
    go = torch.empty((n, H, W, 128), dtype=torch.bfloat16, device=dev)
    _lib.check(L.stx_avgpool2_bwd(gi.data_ptr(), x.data_ptr(), n, H, W, 128, go.data_ptr(), st()))
    ref = (F.interpolate(gi.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1) * 0.25) * (x.float() > 0)
    print("avgpool bwd:", relerr(go.float(), ref), flush=True)


def stage_plan():
    import syntex_oracle as so
    from syntex_b200.gatys import Synthesizer
    raw = so.synthetic_vgg_weights(0)
    W = so.load_vgg_weights(raw)
    for S, prec in ((64, "bf16x3"), (128, "bf16x3"), (256, "bf16x3"), (256, "bf16"), (512, "bf16x3")):
        orc = so.SynthesizerOracle(W, S, dtype=torch.float64)
        tgt = so.synthetic_texture(S, seed=3)
        x0 = so.init_input((S, S, 3), False, 0)
        syn = Synthesizer({"vgg19_npy_path": raw, "image_size": S, "maxiter": 10, "maxcor": 20, "precision": prec})
        syn.build()
        print("---- S %d precision %s" % (S, prec), flush=True)
        t = time.time()
        go = orc.compute_gram(tgt, update=True)
        gg = syn.compute_gram(tgt, update=False)
        for k in gg:
            print("S%d gram %s rel %.2e" % (S, k, relerr(torch.as_tensor(gg[k]), torch.as_tensor(go[k]))[0]), flush=True)
        syn.compute_gram(tgt, update=True)
        ro = orc.step(x0)
        rg = syn.step(x0)
        print("S%d func gpu %.6e oracle %.6e rel %.2e" % (S, rg["func"].sum(), ro["func"].sum(),
              abs(rg["func"].sum() - ro["func"].sum()) / ro["func"].sum()))
        print("S%d grad rel %.3e (max abs %.3e, ref max %.3e)" % ((S,) + relerr(torch.as_tensor(rg["grad"]), torch.as_tensor(ro["grad"])) + (float(np.abs(ro["grad"]).max()),)), flush=True)
        # timing of the f/g evaluation through the host API
        for _ in range(3):
            syn.step(x0)
        torch.cuda.synchronize()
        t = time.time()
        for _ in range(10):
            syn.step(x0)
        print("S%d step_host %.3f ms" % (S, (time.time() - t) * 100), flush=True)


def stage_lbfgs():
    import syntex_oracle as so
    import lbfgs_model as lm
    from syntex_b200.gatys import Synthesizer
    raw = so.synthetic_vgg_weights(0)
    for S, IT, prec in ((64, 50, "bf16x3"), (256, 100, "bf16x3"), (256, 100, "bf16")):
        tgt = so.synthetic_texture(S, seed=3)
        x0 = so.init_input((S, S, 3), False, 0)
        print("---- S %d precision %s" % (S, prec), flush=True)
        syn = Synthesizer({"vgg19_npy_path": raw, "image_size": S, "maxiter": IT, "maxcor": 20, "precision": prec})
        syn.build()
        syn.compute_gram(tgt, update=True)
        f0 = syn.step(x0)["func"].sum()
        t = time.time()
        x, fun = syn.synthesize(x0, tgt)
        dt = time.time() - t
        f1 = syn.step(x)["func"].sum()
        print("S%d device lbfgs: f0 %.4e -> fun %.4e (re-eval %.4e) nit %s nfev %s  %.3f s  (%.1f it/s)" % (
            S, f0, fun, f1, syn.last_result["nit"], syn.last_result["nfev"], dt, IT / dt), flush=True)
        t = time.time()
        x, fun = syn.synthesize(x0, tgt)
        dt = time.time() - t
        print("S%d device lbfgs 2nd run: fun %.4e  %.3f s (%.1f it/s)" % (S, fun, dt, IT / dt), flush=True)
        # same algorithm on the host (numpy model) over the GPU f/g
        def fg(xv):
            r = syn.step(xv.reshape(1, S, S, 3))
            return float(r["func"].sum()), r["grad"].ravel()
        o = lm.minimize(fg, x0, IT)
        print("S%d model over gpu f/g: fun %.4e nit %d nfev %d" % (S, o.f, o.nit, o.nfev), flush=True)
        syn2 = Synthesizer({"vgg19_npy_path": raw, "image_size": S, "maxiter": IT, "maxcor": 20, "optimizer": "scipy"})
        syn2.build()
        t = time.time()
        x2, fun2 = syn2.synthesize(x0, tgt)
        print("S%d scipy over gpu f/g: fun %.4e nit %s nfev %s %.3f s" % (S, fun2, syn2.last_result["nit"], syn2.last_result["nfev"], time.time() - t), flush=True)


if __name__ == "__main__":
    stages = sys.argv[1:] or ["elem", "conv", "split", "gram", "plan", "lbfgs"]
    print(torch.cuda.get_device_name(0), flush=True)
    for s in stages:
        print("==== stage", s, flush=True)
        try:
            globals()["stage_" + s]()
        except Exception as ex:
            import traceback
            traceback.print_exc()
            print("stage", s, "aborted:", ex, flush=True)
            if "CUDA" in str(ex) or "cuda" in str(ex):
                break
