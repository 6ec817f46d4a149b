"""CPU tests of the PO side: the oracle's differentiable texture door
against the closed-form second-order term, the generator layers against numpy/TF padding semantics, the
ProgressiveSynTex graph over the oracle, and the data-parallel trainer (gradient mean over gloo, world size 2)."""
import os
import socket
import sys
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import syntex_oracle as so
from syntex_b200.po.layers import SymConv, UpsampleNN, symmetric_pad
from syntex_b200.po.model import SynTexModelDesc
from syntex_b200.po.progressive_model import ProgressiveSynTex
from syntex_b200.po.trainer import SynTexTrainer

LAYERS = list(SynTexModelDesc.DEFAULT_COEFS.keys())


def test_symmetric_pad_matches_numpy():
    x = torch.arange(2 * 3 * 5 * 6, dtype=torch.float32).reshape(2, 3, 5, 6)
    for before, after in ((1, 1), (2, 2), (2, 1), (0, 3)):
        ref = np.pad(x.numpy(), ((0, 0), (0, 0), (before, after), (before, after)), mode="symmetric")
        assert np.array_equal(symmetric_pad(x, before, after).numpy(), ref)
    with pytest.raises(ValueError):
        symmetric_pad(x, 6)


def test_generator_layer_shapes():
    x = torch.randn(1, 8, 6, 6)
    assert SymConv(8, 4, 3)(x).shape == (1, 4, 6, 6)
    assert SymConv(8, 4, 5, act=False)(x).shape == (1, 4, 6, 6)
    assert UpsampleNN(8, 4)(x).shape == (1, 4, 12, 12)


def test_oracle_second_order_term_matches_closed_form(raw_weights):
    """d/dF <U, grad_layer> for grad_layer = alpha F (G - T): alpha [U D + F (F^T U + U^T F) / n]  (SURVEY 8 a-14)."""
    g = torch.Generator().manual_seed(0)
    B, H, W, C = 2, 6, 5, 8
    F = torch.randn((B, H, W, C), generator=g, dtype=torch.float64).requires_grad_(True)
    T = torch.randn((B, C, C), generator=g, dtype=torch.float64)
    T = T + T.transpose(1, 2)
    U = torch.randn((B, H, W, C), generator=g, dtype=torch.float64)
    _, _, gl = so.build_texture_loss({"k": F}, {"k": T}, {"k": 1.0}, calc_grad=True, create_graph=True)
    got = torch.autograd.grad((gl["k"] * U).sum(), F)[0]
    n = H * W + 1e-6
    alpha = 4.0 / (C * C * n)
    Fm, Um = F.detach().reshape(B, H * W, C), U.reshape(B, H * W, C)
    D = Fm.transpose(1, 2) @ Fm / n - T
    want = alpha * (Um @ D + Fm @ (Fm.transpose(1, 2) @ Um + Um.transpose(1, 2) @ Fm) / n)
    assert torch.allclose(got.reshape(B, H * W, C), want, rtol=1e-10, atol=1e-12)
    assert torch.allclose(gl["k"].detach().reshape(B, H * W, C), alpha * Fm @ D, rtol=1e-10, atol=1e-12)


def _tiny_model(raw_weights, seed=0, dtype=torch.float64):
    torch.manual_seed(seed)
    tex = so.OracleTexture(so.load_vgg_weights(raw_weights), LAYERS, dtype)
    model = ProgressiveSynTex({"vgg19_npy_path": raw_weights, "image_size": 32, "n_gpu": 1}, texture_fn=tex)
    return model.to(dtype)


def test_progressive_graph_over_oracle(raw_weights):
    model = _tiny_model(raw_weights)
    assert len(model.syn) == 5 and [len(s.layers) for s in model.syn] == [1, 2, 3, 4, 5]
    rng = np.random.RandomState(0)
    z = rng.uniform(-1, 1, (1, 32, 32, 3))
    t = so.synthetic_texture(32, seed=1)[None] * 255.0
    loss = model.build_graph(torch.as_tensor(z, dtype=torch.float64), torch.as_tensor(t, dtype=torch.float64))
    assert loss.ndim == 0 and torch.isfinite(loss)
    assert len(model.losses) == 6 and len(model.image_outputs) == 6
    assert list(model.loss_layer_output.keys()) == LAYERS
    # the objective skips the first loss (computed from noise): progressive_model.py:262-264
    assert torch.allclose(loss, sum(model.losses[1:]))
    loss.backward()
    for s, stage in enumerate(model.syn):
        gn = sum(float(p.grad.abs().sum()) for p in stage.parameters() if p.grad is not None)
        assert gn > 0, "stage %d received no gradient" % s
    viz = model.viz(torch.as_tensor(t / 255.0))
    assert viz.dtype == torch.uint8 and tuple(viz.shape) == (1, 32, 7 * 32, 3)
    opt = model.optimizer()
    assert opt.defaults["eps"] == 1e-3 and opt.defaults["lr"] == 5e-5


def test_parser_surface():
    ps = ProgressiveSynTex.get_parser()
    args = ps.parse_args(["--image-size", "512", "--n-gpu", "4"])
    assert args.get("image_size") == 512 and args.get("n_gpu") == 4 and args.get("n_block") == 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, raw_weights, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        model = _tiny_model(raw_weights, seed=100 + rank, dtype=torch.float64)   # different init per rank on purpose
        model._lr = 5e-5 * world
        trainer = SynTexTrainer(None, model, world)
        rng = np.random.RandomState(rank)
        z = torch.as_tensor(rng.uniform(-1, 1, (1, 32, 32, 3)), dtype=torch.float64)
        t = torch.as_tensor(so.synthetic_texture(32, seed=10 + rank)[None] * 255.0, dtype=torch.float64)
        p0 = trainer.buckets.flat_param.clone()
        trainer.run_step((z, t))
        out[rank] = (p0.numpy(), trainer.buckets.flat_grad.clone().numpy(), trainer.buckets.flat_param.clone().numpy(),
                     float(trainer.opt.param_groups[0]["lr"]))
    finally:
        dist.destroy_process_group()


def test_data_parallel_trainer_gloo_world2(raw_weights):
    """Replicas start identical (rank 0's weights), gradients are the MEAN over ranks (po/model.py:132), the
    learning rate carries the n_gpu factor (progressive_model.py:36-37), and the replicas stay identical."""
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_dp_worker, args=(world, port, raw_weights, out), nprocs=world, join=True)
        r0, r1 = out[0], out[1]
    assert np.array_equal(r0[0], r1[0])          # broadcast initial parameters
    assert np.array_equal(r0[1], r1[1])          # all-reduced gradients identical on both ranks
    assert np.array_equal(r0[2], r1[2])          # replicas identical after the step
    assert not np.array_equal(r0[0], r0[2])
    assert r0[3] == pytest.approx(1e-4)
    # the mean really is the mean: recompute both local gradients in this process
    locals_ = []
    for rank in range(world):
        model = _tiny_model(raw_weights, seed=100, dtype=torch.float64)
        trainer = SynTexTrainer(None, model, 1)
        assert np.array_equal(trainer.buckets.flat_param.numpy(), r0[0])
        rng = np.random.RandomState(rank)
        z = torch.as_tensor(rng.uniform(-1, 1, (1, 32, 32, 3)), dtype=torch.float64)
        t = torch.as_tensor(so.synthetic_texture(32, seed=10 + rank)[None] * 255.0, dtype=torch.float64)
        trainer.buckets.zero_grad()
        model.build_graph(z, t).backward()
        locals_.append(trainer.buckets.flat_grad.clone().numpy())
    mean = 0.5 * (locals_[0] + locals_[1])
    # float64 replicas (in float32 the 5-stage graph amplifies summation-order noise to ~10 % of the gradient norm)
    assert np.linalg.norm(r0[1] - mean) / np.linalg.norm(mean) < 1e-3
    assert np.linalg.norm(r0[1] - locals_[0]) / np.linalg.norm(mean) > 1e-2   # ... and is not just one rank's gradient


@pytest.mark.parametrize("opts", [{}, {"pad_type": "symmetric", "norm_type": "none", "act_type": "lrelu", "pre_act": True},
                                  {"pad_type": "zero", "norm_type": "batch", "stop_grad": True, "deconv_upsample": True}])
def test_adaptive_graph_over_oracle(raw_weights, opts):
    """po/adaptive_model.py: one shared-structure stage per unfolding, fed by the full-VGG-backprop gradients; with
    --stop-grad off the training loss reaches stage s through the gradients handed to stage s+1 (double backward)."""
    from syntex_b200.po.adaptive_model import AdaptiveSynTex
    torch.manual_seed(0)
    tex = so.OracleTexture(so.load_vgg_weights(raw_weights), LAYERS, torch.float64)
    args = {"vgg19_npy_path": raw_weights, "image_size": 32, "n_stage": 2}
    args.update(opts)
    model = AdaptiveSynTex(args, texture_fn=tex).double()
    rng = np.random.RandomState(0)
    z = torch.as_tensor(rng.uniform(-1, 1, (1, 32, 32, 3)))
    t = torch.as_tensor(so.synthetic_texture(32, seed=1)[None] * 255.0)
    loss = model.build_graph(z, t)
    assert torch.isfinite(loss) and len(model.losses) == 3 and len(model.image_outputs) == 3
    assert torch.allclose(loss, sum(model.losses[1:]))
    loss.backward()
    g0 = sum(float(p.grad.abs().sum()) for p in model.syn[0].parameters() if p.grad is not None)
    g1 = sum(float(p.grad.abs().sum()) for p in model.syn[1].parameters() if p.grad is not None)
    assert g0 > 0 and g1 > 0
    opt = model.optimizer()
    assert opt.defaults["betas"] == (0.5, 0.999) and opt.defaults["eps"] == 1e-3
    ps = AdaptiveSynTex.get_parser()
    a = ps.parse_args(["--pad-type", "zero", "--stop-grad"])
    assert a.get("pad_type") == "zero" and a.get("stop_grad") and a.get("n_stage") == 1


def test_adaptive_second_order_path_changes_the_gradient(raw_weights):
    """Stage 0's parameters influence stage 1's loss also through the gradients stage 1 is fed: --stop-grad removes
    exactly that path (adaptive_model.py:219-220)."""
    from syntex_b200.po.adaptive_model import AdaptiveSynTex
    tex = so.OracleTexture(so.load_vgg_weights(raw_weights), LAYERS, torch.float64)
    rng = np.random.RandomState(0)
    z = torch.as_tensor(rng.uniform(-1, 1, (1, 32, 32, 3)))
    t = torch.as_tensor(so.synthetic_texture(32, seed=1)[None] * 255.0)
    grads = []
    for stop in (False, True):
        torch.manual_seed(0)
        m = AdaptiveSynTex({"vgg19_npy_path": raw_weights, "image_size": 32, "n_stage": 2, "stop_grad": stop},
                           texture_fn=tex).double()
        m.build_graph(z, t).backward()
        grads.append(torch.cat([p.grad.flatten() for p in m.syn[0].parameters()]))
    rel = float((grads[0] - grads[1]).norm() / grads[0].norm())
    assert rel > 1e-3


def test_single_model_graph_over_oracle(raw_weights):
    """po/single_model.py (BASELINE config 3's model): the generator reads the VGG features of the five Gram layers;
    training reaches the generator of stage 0 through stage 1's features (VGG back-propagation of feature gradients)."""
    from syntex_b200.po.single_model import SingleSynTex
    torch.manual_seed(0)
    tex = so.OracleTexture(so.load_vgg_weights(raw_weights), LAYERS, torch.float64)
    model = SingleSynTex({"vgg19_npy_path": raw_weights, "image_size": 32, "n_stage": 2, "loss_scale": 1.0},
                         texture_fn=tex).double()
    rng = np.random.RandomState(0)
    z = torch.as_tensor(rng.uniform(-1, 1, (1, 32, 32, 3)))
    t = torch.as_tensor(so.synthetic_texture(32, seed=1)[None] * 255.0)
    loss = model.build_graph(z, t)
    assert torch.isfinite(loss) and len(model.losses) == 3
    loss.backward()
    for s in (0, 1):
        assert sum(float(p.grad.abs().sum()) for p in model.syn[s].parameters() if p.grad is not None) > 0
    opt = model.optimizer()
    assert opt.defaults["lr"] == 5e-5 and opt.defaults["betas"] == (0.9, 0.999)
    args = SingleSynTex.get_parser().parse_args([])
    assert args.get("loss_scale") == 0. and args.get("n_stage") == 1


@pytest.mark.parametrize("opts", [{}, {"pad_type": "zero", "norm_type": "batch", "act_type": "lrelu"},
                                  {"pad_type": "symmetric", "norm_type": "none", "pre_act": True, "stop_grad": True}])
def test_improved_graph_over_oracle(raw_weights, opts):
    """po/improved_model.py: five progressive stages fed by the full-VGG-backprop gradients of their first s+1 Gram
    layers; loss_input of a stage leaves out its last layer (:218-226); --stop-grad also cuts the pre-image path."""
    from syntex_b200.po.improved_model import ImprovedSynTex, test_only
    torch.manual_seed(0)
    tex = so.OracleTexture(so.load_vgg_weights(raw_weights), LAYERS, torch.float64)
    args = {"vgg19_npy_path": raw_weights, "image_size": 32}
    args.update(opts)
    model = ImprovedSynTex(args, texture_fn=tex).double()
    assert [len(s.layers) for s in model.syn] == [1, 2, 3, 4, 5]
    rng = np.random.RandomState(0)
    z = torch.as_tensor(rng.uniform(-1, 1, (1, 32, 32, 3)))
    t = torch.as_tensor(so.synthetic_texture(32, seed=1)[None] * 255.0)
    loss = model.build_graph(z, t)
    assert torch.isfinite(loss) and len(model.losses) == 6 and len(model.image_outputs) == 6
    assert float(model.losses[0]) == 0.0                      # one layer: nothing left after dropping the last
    assert torch.allclose(loss, sum(model.losses[1:]))
    # stage 1's loss_input is the conv1_1 term of ITS input image only
    x1 = model.image_outputs[1].detach()
    ll, _ = tex(x1, LAYERS[:1], tex.grams(t / 255.), 0, False)
    assert torch.allclose(model.losses[1], (1e5 / 4 * ll["conv1_1"]).mean(), rtol=1e-9)
    loss.backward()
    g = [sum(float(p.grad.abs().sum()) for p in st.parameters() if p.grad is not None) for st in model.syn]
    assert all(v > 0 for v in g)
    # BatchNorm uses batch statistics in every phase (improved_model.py:353-360): eval() changes nothing
    with torch.no_grad():
        a = model.build_graph(z, t)
        model.eval()
        b = model.build_graph(z, t)
    assert torch.equal(a, b)
    opt = model.optimizer()
    assert opt.defaults["betas"] == (0.5, 0.999) and opt.defaults["eps"] == 1e-3 and opt.defaults["lr"] == 2e-4
    ps = ImprovedSynTex.get_parser()
    a = ps.parse_args([])
    assert a.get("pad_type") == "zero" and a.get("norm_type") == "batch" and a.get("act_type") == "lrelu"
    outs, losses = test_only(model, [(z, t)])
    assert outs[0].shape == (1, 32, 32 * 7, 3) and outs[0].dtype == np.uint8 and np.isfinite(losses[0])


def test_improved_stop_grad_cuts_the_cross_stage_paths(raw_weights):
    from syntex_b200.po.improved_model import ImprovedSynTex
    tex = so.OracleTexture(so.load_vgg_weights(raw_weights), LAYERS, torch.float64)
    rng = np.random.RandomState(0)
    z = torch.as_tensor(rng.uniform(-1, 1, (1, 32, 32, 3)))
    t = torch.as_tensor(so.synthetic_texture(32, seed=1)[None] * 255.0)
    grads = []
    for stop in (False, True):
        torch.manual_seed(0)
        m = ImprovedSynTex({"vgg19_npy_path": raw_weights, "image_size": 32, "stop_grad": stop}, texture_fn=tex).double()
        m.build_graph(z, t).backward()
        grads.append(torch.cat([p.grad.flatten() for p in m.syn[0].parameters()]))
    rel = float((grads[0] - grads[1]).norm() / grads[0].norm())
    assert rel > 1e-3


def test_single_model_synthesize_returns_the_output_image_of_the_graph(raw_weights):
    """SingleSynTex.synthesize(): feed-forward synthesis without the loss of the result - the same image as the last
    entry of build_graph's image_outputs."""
    from syntex_b200.po.single_model import SingleSynTex
    torch.manual_seed(0)
    tex = so.OracleTexture(so.load_vgg_weights(raw_weights), LAYERS, torch.float64)
    model = SingleSynTex({"vgg19_npy_path": raw_weights, "image_size": 32, "n_stage": 2}, texture_fn=tex).double()
    rng = np.random.RandomState(0)
    z = torch.as_tensor(rng.uniform(-1, 1, (1, 32, 32, 3)))
    t = torch.as_tensor(so.synthetic_texture(32, seed=1)[None] * 255.0)
    with torch.no_grad():
        model.build_graph(z, t)
    out = model.synthesize(z, t)
    assert out.shape == (1, 32, 32, 3) and not out.requires_grad
    assert torch.equal(out, model.image_outputs[-1])
