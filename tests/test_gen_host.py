"""Host logic of the hand-written generator route (po/gen_ops.py) that needs no GPU: the phase decomposition that
turns "nearest x2 + symmetric pad 2 + 5x5 conv" (build_upsampling_nn, po/progressive_model.py:123-134) into one
edge-replicated 3x3 convolution of the low-resolution map plus a pixel shuffle, and the loud failure without CUDA."""
import pytest
import torch
import torch.nn.functional as F

from syntex_b200.po import gen_ops
from syntex_b200.po.layers import UpsampleNN, symmetric_pad


@pytest.mark.parametrize("cin,cout,h,w", [(3, 2, 5, 7), (8, 4, 6, 6), (1, 1, 1, 1)])
def test_upsample_phase_filters_equal_the_5x5_convolution(cin, cout, h, w):
    torch.manual_seed(0)
    m = UpsampleNN(cin, cout).double()
    x = torch.randn(2, cin, h, w, dtype=torch.float64, requires_grad=True)
    ref = m(x)
    w3 = gen_ops.upsample_weights(m.conv.conv.weight)
    assert w3.shape == (4 * cout, cin, 3, 3)
    y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), w3)               # [B, 4*cout, h, w]
    y = gen_ops.pixel_shuffle_nhwc(y.permute(0, 2, 3, 1).contiguous(), cout)   # [B, 2h, 2w, cout]
    assert torch.allclose(y.permute(0, 3, 1, 2), ref, atol=1e-12)
    # ... and so are the gradients that flow back to the 5x5 parameters through the phase combination
    g = torch.randn_like(ref)
    gw_ref, gx_ref = torch.autograd.grad(ref, [m.conv.conv.weight, x], g)
    gw, gx = torch.autograd.grad(y.permute(0, 3, 1, 2), [m.conv.conv.weight, x], g)
    assert torch.allclose(gw, gw_ref, atol=1e-10) and torch.allclose(gx, gx_ref, atol=1e-10)


@pytest.mark.parametrize("pad_type", ["SYMMETRIC", "REFLECT", "CONSTANT"])
def test_upsample_phase_filters_equal_the_3x3_convolution_after_tiling(pad_type):
    """build_upsampling_nnconv (po/improved_model.py:185-199): tile x2, pad 1 (any tf.pad mode), 3x3 conv."""
    from syntex_b200.po.adaptive_model import _Upsample
    torch.manual_seed(0)
    m = _Upsample(6, 4, pad_type, True).double()
    x = torch.randn(2, 6, 5, 7, dtype=torch.float64)
    ref = m(x)
    w3 = gen_ops.upsample_weights(m.conv.conv.weight)
    xp = F.pad(x, (1, 1, 1, 1)) if pad_type == "CONSTANT" else F.pad(x, (1, 1, 1, 1), mode="replicate")
    y = gen_ops.pixel_shuffle_nhwc(F.conv2d(xp, w3).permute(0, 2, 3, 1).contiguous(), 4)
    assert torch.allclose(y.permute(0, 3, 1, 2), ref, atol=1e-12)


def test_symmetric_pad_of_upsampled_map_is_edge_replication_of_the_low_resolution_map():
    x = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).reshape(2, 3, 4, 5)
    up = F.interpolate(x, scale_factor=2, mode="nearest")
    a = symmetric_pad(up, 2)
    b = F.interpolate(F.pad(x, (1, 1, 1, 1), mode="replicate"), scale_factor=2, mode="nearest")
    assert torch.equal(a, b)


def test_native_operators_fail_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    x = torch.zeros((1, 8, 8, 64), dtype=torch.bfloat16)
    with pytest.raises(Exception):
        gen_ops.conv3x3(x, torch.zeros(64, 64, 3, 3))
    with pytest.raises(Exception):
        gen_ops.instance_norm(x, torch.ones(64), torch.zeros(64))


def test_native_route_selection_rules():
    """Which generator stages the hand-written route claims (po/adaptive_model.py:stage_native_supported) and what
    happens off the GPU: `auto` falls back to stock torch, `native` refuses - there is no CPU route."""
    from syntex_b200.po.adaptive_model import _Stage as AdaptiveStage, run_generator
    from syntex_b200.po.improved_model import _Stage as ImprovedStage
    from syntex_b200.po.progressive_model import _Stage as ProgressiveStage

    ok = ImprovedStage(2, 1, False, True, "REFLECT", "batch", "lrelu", 3)
    assert ok.native_supported(1) and ok.native_supported(2)     # batch statistics: the batch folds into the pixel axis
    assert ImprovedStage(2, 1, True, True, "CONSTANT", "instance", "relu", 3).native_supported(4)
    assert ImprovedStage(2, 1, False, True, "SYMMETRIC", "none", "relu", 3).native_supported(4)
    assert not ImprovedStage(2, 1, False, True, "REFLECT", "instance", "relu", 5).native_supported(1)   # 5x5 gradient convs
    assert not ImprovedStage(2, 1, False, False, "REFLECT", "instance", "relu", 3).native_supported(1)  # deconvolution
    assert not AdaptiveStage(1, False, True, "REFLECT", "batch", "relu", 3).native_supported(1)         # running statistics
    assert ProgressiveStage(3, 1, False, True).native_supported() and not ProgressiveStage(3, 1, False, False).native_supported()

    class _Model:
        _generator, _door_is_cuda = "auto", True
    stage = ImprovedStage(1, 1, False, True, "REFLECT", "instance", "relu", 3)
    x = {"conv1_1": torch.randn(1, 8, 8, 64)}
    out = run_generator(_Model(), stage, x)                                # CPU tensors: the stock route
    assert out.shape == (1, 8, 8, 3)
    _Model._generator = "native"
    with pytest.raises(NotImplementedError):
        run_generator(_Model(), stage, x)
