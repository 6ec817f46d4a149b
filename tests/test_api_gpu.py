"""The texture_utils / po.model mirrors on the B200 against the oracle."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import syntex_oracle as so  # noqa: E402
from gpu_util import relerr  # noqa: E402
from syntex_b200.po import SynTexModelDesc  # noqa: E402
from syntex_b200.texture_utils import Vgg19Extractor, build_gram, build_texture_loss  # noqa: E402


def test_extractor_build_matches_oracle(raw_weights):
    x = torch.rand((2, 64, 96, 3), generator=torch.Generator().manual_seed(0))
    ext = Vgg19Extractor(raw_weights)
    feats = ext.build(x.cuda())
    ref = so.vgg_build(x.double(), so.TorchWeights(so.load_vgg_weights(raw_weights), torch.float64))
    assert list(feats.keys()) == list(ref.keys()) and len(feats) == 16
    for k in ref:
        assert tuple(feats[k].shape) == tuple(ref[k].shape) and feats[k].dtype == torch.float32
        assert relerr(feats[k], ref[k]) < 1e-4, k
    part = ext.build(x.cuda(), topmost="pool2")
    assert list(part.keys())[-1] == "pool2" and len(part) == 6
    assert torch.equal(part["pool2"], feats["pool2"])


def test_build_gram_and_texture_loss(raw_weights):
    x = torch.rand((1, 64, 64, 3), generator=torch.Generator().manual_seed(1))
    # a smooth exemplar: its Grams differ strongly from those of the noise image (no cancellation in G - T)
    t = torch.as_tensor(so.synthetic_texture(64, seed=2)[None]).float()
    ext = Vgg19Extractor(raw_weights)
    coefs = OrderedDict([("conv1_1", 1e5), ("pool2", 2e5), ("pool4", 5e4)])
    fx, ft = ext.build(x.cuda()), ext.build(t.cuda())
    gt = OrderedDict((k, build_gram(ft[k])) for k in coefs)
    lo, ll, gl = build_texture_loss(fx, gt, coefs, calc_grad=True)
    W = so.TorchWeights(so.load_vgg_weights(raw_weights), torch.float64)
    rx = so.vgg_build(x.double(), W)
    rx = OrderedDict((k, v.detach().requires_grad_(True)) for k, v in rx.items())
    rt = so.vgg_build(t.double(), W)
    rgt = OrderedDict((k, so.build_gram(rt[k])) for k in coefs)
    rlo, rll, rgl = so.build_texture_loss(rx, rgt, coefs, calc_grad=True)
    assert tuple(lo.shape) == (1,)
    assert abs(float(lo[0]) - float(rlo[0].detach())) / float(rlo[0].detach()) < 5e-3
    for k in coefs:
        assert relerr(gt[k], rgt[k]) < 1e-3
        assert abs(float(ll[k][0]) - float(rll[k][0].detach())) / float(rll[k][0].detach()) < 5e-3
        assert tuple(gl[k].shape) == tuple(rgl[k].shape)
        assert relerr(gl[k], rgl[k]) < 1e-2
    lo2, _, none = build_texture_loss(fx, ft, coefs, is_gram_b=False)
    assert none is None and torch.allclose(lo2, lo, rtol=1e-6)
    with pytest.raises(NotImplementedError):
        build_gram(fx["pool1"], mask=torch.ones((1, 32, 32)))


def test_po_model_doors(raw_weights):
    model = SynTexModelDesc({"vgg19_npy_path": raw_weights})
    x = torch.rand((2, 64, 64, 3), generator=torch.Generator().manual_seed(3)).cuda()
    feat, gram = model._build_extractor(x)
    assert list(gram.keys()) == list(SynTexModelDesc.DEFAULT_COEFS.keys())
    assert tuple(gram["pool4"].shape) == (2, 512, 512)
    feat2, none = model._build_extractor(x, calc_gram=False)
    assert none is None
    lo, ll, g = model._build_loss(x, gram)
    assert float(lo.abs().max()) == 0.0 and g is None          # loss of an image against its own Grams
    y = torch.rand((2, 64, 64, 3), generator=torch.Generator().manual_seed(4)).cuda()
    lo, ll, g = model._build_loss(y, gram, calc_grad=True)
    assert (lo > 0).all() and set(g.keys()) == set(SynTexModelDesc.DEFAULT_COEFS.keys())


def test_stage_preparation_gradients_match_autograd(raw_weights):
    """po/improved_model.py:209-232: features to the stage's top layer + full VGG back-propagation of the loss to
    every Gram layer's activations."""
    from syntex_b200.po import build_stage_preparation
    x = torch.rand((2, 64, 64, 3), generator=torch.Generator().manual_seed(5))
    t = torch.as_tensor(np.stack([so.synthetic_texture(64, seed=7), so.synthetic_texture(64, seed=8)])).float()
    coefs = OrderedDict([("conv1_1", 1e5), ("pool1", 1e5), ("pool2", 1e5)])     # stage 3 of 5
    ext = Vgg19Extractor(raw_weights)
    W = so.TorchWeights(so.load_vgg_weights(raw_weights), torch.float64)
    rt = so.vgg_build(t.double(), W, topmost="pool2")
    rgt = OrderedDict((k, so.build_gram(rt[k])) for k in coefs)
    feat, loss_input, loss_layer, grads = build_stage_preparation(
        ext, x.cuda(), OrderedDict((k, v.float()) for k, v in rgt.items()), coefs)
    assert list(feat.keys()) == list(coefs.keys())          # the features of the stage's Gram layers
    xr = x.double().requires_grad_(True)
    rf = so.vgg_build(xr, W, topmost="pool2")
    lo, ll, _ = so.build_texture_loss(rf, rgt, coefs)
    ref = torch.autograd.grad(lo.sum(), [rf[k] for k in coefs])
    for (k, g), r in zip(grads.items(), ref):
        assert tuple(g.shape) == tuple(r.shape)
        assert relerr(g, r) < 1e-2, k
        assert abs(float(loss_layer[k][0]) - float(ll[k][0])) / float(ll[k][0]) < 5e-3
    want = sum(coefs[k] / 4 * ll[k] for k in list(coefs)[:-1]).detach()
    assert relerr(loss_input, want) < 5e-3


@pytest.mark.parametrize("is_preimage", [False, True])
def test_synthesizer_verbose_record_matches_oracle(raw_weights, is_preimage):
    """The verbose fetches of gatys/synthesizer.py:80-107: func, gi_all and per Gram layer
    l_k, a_k, gl_k, gi_k, for the image and for the pre-image (sigmoid) parametrisation."""
    from syntex_b200.gatys import Synthesizer
    size = 64
    syn = Synthesizer({"vgg19_npy_path": raw_weights, "image_size": size, "maxiter": 3, "maxcor": 20, "verbose": True})
    syn.build()
    target = so.synthetic_texture(size, seed=31)
    rng = np.random.RandomState(5)
    x = rng.uniform(-1, 1, (1, size, size, 3)) if is_preimage else rng.uniform(0.05, 0.95, (1, size, size, 3))
    syn.compute_gram(target, update=True)
    syn.reset_verbose()
    rec = syn._verbose_record(x, is_preimage)
    # oracle
    W = so.TorchWeights(so.load_vgg_weights(raw_weights), torch.float64)
    coefs = Synthesizer.DEFAULT_COEFS
    gt = OrderedDict((k, so.build_gram(v)) for k, v in so.vgg_build(torch.as_tensor(target[None]).double(), W).items()
                     if k in coefs)
    xr = torch.as_tensor(x).double().requires_grad_(True)
    img = torch.sigmoid(xr) if is_preimage else xr
    feat = so.vgg_build(img, W)
    lo, ll, gl = so.build_texture_loss(feat, gt, coefs, calc_grad=True)
    gi_all = torch.autograd.grad(lo.sum(), xr, retain_graph=True)[0]
    assert abs(rec["func"] - float(lo.sum())) / float(lo.sum()) < 5e-3
    assert abs(rec["gi_all"] - float(gi_all.abs().mean())) / float(gi_all.abs().mean()) < 1e-2
    for k in coefs:
        assert abs(rec["l_" + k] - float(ll[k].sum())) / float(ll[k].sum()) < 5e-3, k
        assert abs(rec["a_" + k] - float(feat[k].abs().mean())) / float(feat[k].abs().mean()) < 1e-4, k
        assert abs(rec["gl_" + k] - float(gl[k].abs().mean())) / float(gl[k].abs().mean()) < 1e-2, k
        gik = torch.autograd.grad(ll[k].sum(), xr, retain_graph=True)[0]
        assert abs(rec["gi_" + k] - float(gik.abs().mean())) / float(gik.abs().mean()) < 1e-2, k
    assert set(syn.verbose_dict.keys()) == {"func", "gi_all"} | {p + k for p in ("l_", "a_", "gl_", "gi_") for k in coefs}


@pytest.mark.parametrize("optimizer", ["device", "scipy"])
def test_synthesize_verbose_collects_one_record_per_iteration(raw_weights, optimizer, tmp_path):
    from syntex_b200.gatys import Synthesizer
    size = 64
    syn = Synthesizer({"vgg19_npy_path": raw_weights, "image_size": size, "maxiter": 4, "maxcor": 20, "verbose": True,
                       "optimizer": optimizer})
    syn.build()
    x0 = np.random.RandomState(0).uniform(1 / 255., 1 - 1 / 255., (size, size, 3))
    img, fun = syn.synthesize(x0, so.synthetic_texture(size, seed=2))
    n = len(syn.verbose_dict["func"])
    assert 1 <= n <= 4 and all(len(v) == n for v in syn.verbose_dict.values())
    assert syn.verbose_dict["func"][-1] < syn.verbose_dict["func"][0] or n == 1
    assert abs(syn.verbose_dict["func"][-1] - float(fun)) / float(fun) < 1e-3
    out = tmp_path / "verbose.npz"
    syn.save_verbose(str(out))
    z = np.load(str(out))
    assert "gi_pool4" in z.files and len(z["func"]) == n


@pytest.mark.parametrize("B,H,W,C", [(2, 32, 32, 64), (1, 16, 24, 128), (2, 8, 8, 512)])
def test_gram_grad_double_bwd_matches_closed_form(B, H, W, C):
    """d/dF <U, alpha F (F^T F / n - T)> = alpha [U D + F (F^T U + U^T F) / n]."""
    from syntex_b200.texture_utils import gram_grad_double_bwd
    g = torch.Generator().manual_seed(C + H)
    F_ = torch.randn((B, H, W, C), generator=g).relu()
    F_ = F_.to(torch.bfloat16).float()            # the op consumes bf16 features: compare on the same inputs
    U = torch.randn((B, H, W, C), generator=g).to(torch.bfloat16).float()
    T = torch.randn((B, C, C), generator=g)
    T = 0.5 * (T + T.transpose(1, 2))
    n = H * W + 1e-6
    Fd = F_.double().reshape(B, H * W, C).requires_grad_(True)
    D = Fd.transpose(1, 2) @ Fd / n - T.double()
    alpha = 4.0 / (C * C * n)
    obj = ((alpha * Fd @ D) * U.double().reshape(B, H * W, C)).sum()
    want = torch.autograd.grad(obj, Fd)[0].reshape(B, H, W, C)
    got = gram_grad_double_bwd(F_.cuda(), D.detach().float().cuda(), U.cuda())
    assert tuple(got.shape) == (B, H, W, C) and relerr(got, want) < 1e-2
