"""The trainer's device pieces: the TensorFlow-rule Adam kernel (stx_adam_tf) against the formula of
tf.train.AdamOptimizer (the reference's optimiser, po/progressive_model.py:268-271), and the guard that keeps a
PO autograd node from back-propagating through a plan whose activations a later forward has overwritten."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import syntex_oracle as so  # noqa: E402
from syntex_b200.po.trainer import TfAdam  # noqa: E402


def _tf_adam_numpy(p, grads, lr, b1, b2, eps):
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    for t, g in enumerate(grads, start=1):
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        p = p - lr_t * m / (np.sqrt(v) + eps)          # epsilon-hat: NOT m_hat / (sqrt(v_hat) + eps)
    return p


@pytest.mark.parametrize("n", [4096, 1000003])
def test_adam_kernel_is_the_tensorflow_update_rule(n):
    rng = np.random.RandomState(0)
    p0 = rng.normal(size=n).astype(np.float32)
    grads = [rng.normal(size=n).astype(np.float32) * s for s in (1.0, 0.1, 3.0, 1e-3)]
    lr, b1, b2, eps = 2e-3, 0.9, 0.999, 1e-3
    want = _tf_adam_numpy(p0.astype(np.float64), [g.astype(np.float64) for g in grads], lr, b1, b2, eps)
    p = torch.from_numpy(p0.copy()).cuda()
    g = torch.zeros_like(p)
    opt = TfAdam(p, g, lr, (b1, b2), eps)
    for gi in grads:
        g.copy_(torch.from_numpy(gi))
        opt.step()
    got = p.cpu().numpy().astype(np.float64)
    assert np.abs(got - want).max() < 5e-6
    # and it is NOT torch.optim.Adam's rule at this epsilon: the two differ by far more than that after four steps
    q = torch.nn.Parameter(torch.from_numpy(p0.copy()).cuda())
    topt = torch.optim.Adam([q], lr=lr, betas=(b1, b2), eps=eps)
    for gi in grads:
        q.grad = torch.from_numpy(gi).cuda()
        topt.step()
    assert np.abs(q.detach().cpu().numpy() - want).max() > 1e-4


def test_backward_through_an_overwritten_plan_raises(raw_weights):
    """Plans are cached per (shape, layers, slot): a second forward with the same key overwrites the activations the
    first graph's backward would read. The node records the plan's generation and refuses."""
    from syntex_b200.po import vgg_texture
    from syntex_b200.texture_utils import Vgg19Extractor
    layers = ["conv1_1", "pool1"]
    ext = Vgg19Extractor(raw_weights)
    W = so.TorchWeights(so.load_vgg_weights(raw_weights), torch.float64)
    tgt = torch.as_tensor(so.synthetic_texture(64, seed=1)[None]).double()
    rt = so.vgg_build(tgt, W, topmost="pool1")
    gram_target = OrderedDict((k, so.build_gram(rt[k]).float().cuda()) for k in layers)
    g = torch.Generator().manual_seed(0)
    x1 = torch.rand((1, 64, 64, 3), generator=g).cuda().requires_grad_(True)
    x2 = torch.rand((1, 64, 64, 3), generator=g).cuda().requires_grad_(True)
    loss1, _ = vgg_texture(x1, ext, layers, gram_target, slot=0, calc_grad=False)
    loss2, _ = vgg_texture(x2, ext, layers, gram_target, slot=0, calc_grad=False)     # same slot: overwrites
    with pytest.raises(RuntimeError, match="re-run"):
        sum(loss1.values()).sum().backward()
    sum(loss2.values()).sum().backward()                                               # the live one is fine
    assert x2.grad is not None and torch.isfinite(x2.grad).all()
    loss3, _ = vgg_texture(x1, ext, layers, gram_target, slot=1, calc_grad=False)     # distinct slots coexist
    loss4, _ = vgg_texture(x2, ext, layers, gram_target, slot=2, calc_grad=False)
    (sum(loss3.values()).sum() + sum(loss4.values()).sum()).backward()
