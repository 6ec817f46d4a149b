"""Diagnostic: native vs stock generator stage, error per output / input / parameter gradient."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from syntex_b200.po.progressive_model import _Stage

def relerr(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300))
def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))

import torch.nn.functional as F
from syntex_b200.po import gen_ops


def emu_conv3x3(x, weight, bias=None, mode="rep", relu_in=False):
    """torch fp32 arithmetic with the kernel's rounding points: bf16 weights, bf16 output."""
    xin = x.float().permute(0, 3, 1, 2)
    if relu_in:
        xin = F.relu(xin)
    w = weight + (weight.to(torch.bfloat16).float() - weight).detach()
    if mode == "rep":
        y = F.conv2d(F.pad(xin, (1, 1, 1, 1), mode="replicate"), w, bias)
    else:
        y = F.conv2d(xin, w, bias, padding=1)
    return y.permute(0, 2, 3, 1).to(torch.bfloat16)


def emu_instance_norm(x, gamma, beta, eps=1e-5, relu=False, shortcut=None):
    y = F.instance_norm(x.float().permute(0, 3, 1, 2), None, None, gamma, beta, True, 0.0, eps)
    if relu:
        y = F.relu(y)
    y = y.permute(0, 2, 3, 1)
    if shortcut is not None:
        y = y + shortcut.float()
    return y.to(torch.bfloat16)


REAL = (gen_ops.conv3x3, gen_ops.instance_norm)

for n_loss_layer, pre_act in [(1, False), (2, False), (3, True), (5, False)]:
    torch.manual_seed(1)
    stage = _Stage(n_loss_layer, 2, pre_act, True).cuda()
    size = 128
    chans = {"conv1_1": 64, "pool1": 64, "pool2": 128, "pool3": 256, "pool4": 512}
    inputs, refs = {}, {}
    s = size
    for i, k in enumerate(stage.layers):
        if k != "conv1_1":
            s //= 2
        v = torch.randn((1, s, s, chans[k]), device="cuda").to(torch.bfloat16).float()
        inputs[k] = v.clone().requires_grad_(True)
        refs[k] = v.clone().requires_grad_(True)
    out = stage.forward_native(inputs)
    if os.environ.get("EMU", "1") == "1":
        gen_ops.conv3x3, gen_ops.instance_norm = emu_conv3x3, emu_instance_norm
        ref = stage.forward_native(refs)
        gen_ops.conv3x3, gen_ops.instance_norm = REAL
    else:
        ref = stage(refs)
    print("stage", n_loss_layer, pre_act, "out relerr %.4f" % relerr(out, ref.detach()), "|ref| %.4g" % float(ref.abs().mean()))
    dy = torch.randn_like(out)
    out.backward(dy)
    gn = {n: p.grad.clone() for n, p in stage.named_parameters()}
    stage.zero_grad()
    ref.backward(dy)
    if os.environ.get("PERTURB", "0") == "1":
        # how far do two runs of the SAME (emulated) arithmetic drift apart when the weights move by less than half a
        # bf16 ulp (1e-3 relative)? This is the conditioning of the comparison, not an error of either arm.
        g_ref = {n: p.grad.clone() for n, p in stage.named_parameters()}
        gi_ref = {k: refs[k].grad.clone() for k in stage.layers}
        stage.zero_grad()
        saved = {n: p.data.clone() for n, p in stage.named_parameters()}
        with torch.no_grad():
            for n, p in stage.named_parameters():
                p.mul_(1.0 + 1e-3 * torch.randn_like(p))
        refs2 = {k: v.detach().clone().requires_grad_(True) for k, v in refs.items()}
        gen_ops.conv3x3, gen_ops.instance_norm = emu_conv3x3, emu_instance_norm
        ref2 = stage.forward_native(refs2)
        gen_ops.conv3x3, gen_ops.instance_norm = REAL
        ref2.backward(dy)
        print("   [conditioning] out relerr %.4f" % relerr(ref2, ref.detach()))
        for k in stage.layers:
            print("   [conditioning] d input %-8s rel %.4f cos %.5f" % (k, relerr(refs2[k].grad, gi_ref[k]), cos(refs2[k].grad, gi_ref[k])))
        worst = sorted(((cos(p.grad, g_ref[n]), relerr(p.grad, g_ref[n]), n) for n, p in stage.named_parameters()))[:3]
        for c, r, n in worst:
            print("   [conditioning] d param %-40s rel %.4f cos %.5f" % (n, r, c))
        with torch.no_grad():
            for n, p in stage.named_parameters():
                p.copy_(saved[n])
                p.grad = g_ref[n]
    for k in stage.layers:
        print("   d input %-8s rel %.4f cos %.5f" % (k, relerr(inputs[k].grad, refs[k].grad), cos(inputs[k].grad, refs[k].grad)))
    worst = sorted(((cos(gn[n], p.grad), relerr(gn[n], p.grad), n) for n, p in stage.named_parameters()))[:6]
    for c, r, n in worst:
        print("   d param %-40s rel %.4f cos %.5f" % (n, r, c))
