"""The CPU oracle against independent restatements, closed forms, finite differences and the committed golden
vectors. (The reference ships no vectors for this path: SURVEY.md section 4 — parity of the oracle itself is
pinned only by these self-consistency checks and by TF's documented op semantics.)"""
import os

import numpy as np
import pytest
import torch

import syntex_oracle as so

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def naive_conv_same_relu(x, w, b):
    """Loop restatement of relu(tf.nn.conv2d(x, w, [1,1,1,1], 'SAME') + b): cross-correlation, HWIO, zero pad 1."""
    H, W, Ci = x.shape
    Co = w.shape[3]
    xp = np.zeros((H + 2, W + 2, Ci))
    xp[1:-1, 1:-1] = x
    y = np.zeros((H, W, Co))
    for r in range(3):
        for s in range(3):
            y += xp[r:r + H, s:s + W, :] @ w[r, s]
    return np.maximum(y + b, 0.0)


def test_conv_layer_semantics():
    rng = np.random.RandomState(0)
    x = rng.normal(size=(5, 7, 3))
    w = rng.normal(size=(3, 3, 3, 4))
    b = rng.normal(size=(4,))
    got = so._conv_layer(torch.as_tensor(x).permute(2, 0, 1)[None], w, b)[0].permute(1, 2, 0).numpy()
    np.testing.assert_allclose(got, naive_conv_same_relu(x, w, b), rtol=1e-12, atol=1e-12)


def test_vgg_build_layers_and_shapes(raw_weights):
    W = so.load_vgg_weights(raw_weights)
    x = torch.rand((2, 32, 48, 3), dtype=torch.float64)
    layers = so.vgg_build(x, W)
    assert list(layers.keys()) == so.LAYER_NAMES
    assert tuple(layers["conv1_1"].shape) == (2, 32, 48, 64)
    assert tuple(layers["pool4"].shape) == (2, 2, 3, 512)
    sub = so.vgg_build(x, W, topmost="pool2")
    assert list(sub.keys())[-1] == "pool2"
    # first layer by hand: x*255 - mean_RGB, flipped conv1_1 weights
    pre = x[0].numpy() * 255.0 - np.array(so.VGG_MEAN_RGB)
    w0 = raw_weights["conv1_1"][0][:, :, ::-1, :].astype(np.float64)
    ref = naive_conv_same_relu(pre, w0, raw_weights["conv1_1"][1].astype(np.float64))
    np.testing.assert_allclose(layers["conv1_1"][0].numpy(), ref, rtol=1e-9, atol=1e-9)
    # avg pool is the mean of each 2x2 block
    c12 = layers["conv1_2"][0].numpy()
    np.testing.assert_allclose(layers["pool1"][0].numpy(), c12.reshape(16, 2, 24, 2, 64).mean(axis=(1, 3)), rtol=1e-12)


def test_rgb_flip_only_touches_conv1_1(raw_weights):
    a = so.load_vgg_weights(raw_weights, is_rgb_input=True)
    b = so.load_vgg_weights(raw_weights, is_rgb_input=False)
    np.testing.assert_array_equal(a["conv1_1"][0], b["conv1_1"][0][:, :, ::-1, :])
    np.testing.assert_array_equal(a["conv2_1"][0], b["conv2_1"][0])
    with pytest.raises(ValueError):
        so.load_vgg_weights("weights.txt")


def test_gram_closed_form_and_mask():
    rng = np.random.RandomState(1)
    f = rng.normal(size=(2, 4, 5, 6))
    g = so.build_gram(torch.as_tensor(f)).numpy()
    ref = np.einsum("bhwc,bhwd->bcd", f, f) / (20 + 1e-6)
    np.testing.assert_allclose(g, ref, rtol=1e-12)
    m = (rng.uniform(size=(2, 4, 5)) > 0.5).astype(np.float64)
    gm = so.build_gram(torch.as_tensor(f), mask=torch.as_tensor(m)).numpy()
    fm = f * m[..., None]
    refm = np.einsum("bhwc,bhwd->bcd", fm, fm) / (m.sum(axis=(1, 2))[:, None, None] + 1e-6)
    np.testing.assert_allclose(gm, refm, rtol=1e-12)
    # fp32: the eps is absorbed for HW >= 32 (ulp(32)/2 > 1e-6; at HW = 16 it still rounds up)
    assert np.float32(32) + np.float32(1e-6) == np.float32(32)


def test_texture_loss_and_layer_gradient_closed_form():
    rng = np.random.RandomState(2)
    coefs = {"a": 3.0, "b": 0.5}
    fa = {k: torch.as_tensor(rng.normal(size=(2, 3, 4, 5))).requires_grad_(True) for k in coefs}
    gb = {k: torch.as_tensor(rng.normal(size=(1, 5, 5))) for k in coefs}
    lo, ll, gl = so.build_texture_loss(fa, gb, coefs, calc_grad=True)
    for k in coefs:
        F = fa[k].detach().numpy().reshape(2, 12, 5)
        G = np.einsum("bnc,bnd->bcd", F, F) / (12 + 1e-6)
        D = G - gb[k].numpy()
        np.testing.assert_allclose(ll[k].detach().numpy(), (D ** 2).mean(axis=(1, 2)), rtol=1e-12)
        ref = (2.0 / (25 * (12 + 1e-6))) * np.einsum("bnc,bcd->bnd", F, D + D.transpose(0, 2, 1))
        np.testing.assert_allclose(gl[k].numpy().reshape(2, 12, 5), ref, rtol=1e-10)
    total = sum(coefs[k] / 4 * ll[k] for k in coefs)
    np.testing.assert_allclose(lo.detach().numpy(), total.detach().numpy(), rtol=1e-12)


def test_input_gradient_matches_finite_differences(raw_weights):
    S = 16
    W = so.load_vgg_weights(raw_weights)
    orc = so.SynthesizerOracle(W, S, dtype=torch.float64)
    orc.compute_gram(so.synthetic_texture(S, seed=1), update=True)
    for pre in (False, True):
        x0 = so.init_input((S, S, 3), pre, 0).astype(np.float32).astype(np.float64)
        r = orc.step(x0, is_preimage=pre)
        rng = np.random.RandomState(3)
        d = rng.normal(size=x0.shape)
        d /= np.linalg.norm(d)
        eps = 1e-4
        # steps must stay float32-representable (the oracle rounds its input to fp32 like the TF placeholder)
        xp = (x0 + eps * d).astype(np.float32).astype(np.float64)
        xm = (x0 - eps * d).astype(np.float32).astype(np.float64)
        fd = (orc.step(xp, is_preimage=pre)["func"].sum() - orc.step(xm, is_preimage=pre)["func"].sum())
        an = float((r["grad"][0] * (xp - xm)).sum())
        assert abs(fd - an) <= 5e-3 * abs(an) + 1e-6, (fd, an)


@pytest.mark.parametrize("S", [64, 128])
def test_oracle_matches_golden(S, raw_weights, exemplar):
    gold = np.load(os.path.join(GOLD, "path_S%d.npz" % S))
    W = so.load_vgg_weights(raw_weights)
    orc = so.SynthesizerOracle(W, S, dtype=torch.float64)
    tgt = so.resize_crop(exemplar / 255.0, S)
    x0 = so.init_input((S, S, 3), False, 0)
    gt = orc.compute_gram(tgt, update=True)
    r = orc.step(x0)
    np.testing.assert_allclose(r["func"], gold["func"], rtol=1e-9)
    np.testing.assert_allclose(np.linalg.norm(r["grad"]), gold["grad_norm"], rtol=1e-8)
    probe = r["grad"].ravel()[::max(1, r["grad"].size // 256)][:256]
    np.testing.assert_allclose(probe, gold["grad_probe"], rtol=1e-6, atol=1e-9 * np.abs(gold["grad_probe"]).max())
    for k in so.DEFAULT_COEFS:
        np.testing.assert_allclose(r["loss_layer"][k], gold["loss_" + k], rtol=1e-9)
        np.testing.assert_allclose(np.trace(gt[k][0]), gold["t_gram_trace_" + k], rtol=1e-10)
        np.testing.assert_allclose(gt[k][0][:8, :8], gold["t_gram_corner_" + k], rtol=1e-9, atol=1e-12)
    p0 = so.init_input((S, S, 3), True, 0)
    rp = orc.step(p0, is_preimage=True)
    np.testing.assert_allclose(rp["func"], gold["pre_func"], rtol=1e-9)


def test_fp32_oracle_close_to_fp64(raw_weights):
    S = 32
    W = so.load_vgg_weights(raw_weights)
    tgt, x0 = so.synthetic_texture(S, seed=1), so.init_input((S, S, 3), False, 0)
    out = {}
    for dt in (torch.float64, torch.float32):
        orc = so.SynthesizerOracle(W, S, dtype=dt)
        orc.compute_gram(tgt, update=True)
        out[dt] = orc.step(x0)
    np.testing.assert_allclose(out[torch.float32]["func"], out[torch.float64]["func"], rtol=1e-4)
    rel = np.linalg.norm(out[torch.float32]["grad"] - out[torch.float64]["grad"]) / np.linalg.norm(out[torch.float64]["grad"])
    assert rel < 1e-3


def test_resize_crop_and_init():
    img = np.random.RandomState(0).uniform(size=(350, 400, 3))
    out = so.resize_crop(img, 256)
    assert out.shape == (256, 256, 3)
    same = so.resize_crop(np.ones((256, 256, 3)), 256)
    np.testing.assert_array_equal(same, np.ones((256, 256, 3)))
    x = so.init_input((8, 8, 3), False, 0)
    assert x.min() >= 1 / 255 and x.max() <= 1 - 1 / 255
    p = so.init_input((8, 8, 3), True, 0)
    assert p.min() >= -1 and p.max() <= 1


def test_scipy_driver_reduces_loss(raw_weights):
    S = 32
    W = so.load_vgg_weights(raw_weights)
    orc = so.SynthesizerOracle(W, S, dtype=torch.float32)
    tgt, x0 = so.synthetic_texture(S, seed=1), so.init_input((S, S, 3), False, 0)
    orc.compute_gram(tgt, update=True)
    f0 = orc.step(x0)["func"].sum()
    x, f, res = orc.synthesize(x0, tgt, maxiter=15)
    assert res.nit == 15 and f < 0.05 * f0
    assert x.min() >= 0.0 and x.max() <= 1.0 and x.shape == (S, S, 3)
