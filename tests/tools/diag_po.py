"""Diagnostics for the PO training-graph parity (GPU hot path vs float64 oracle). Test tooling: lives under tests/
because it uses the oracle.  python tests/tools/diag_po.py [size]"""
import os, sys
from collections import OrderedDict
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import syntex_oracle as so
from syntex_b200.synthetic import synthetic_vgg_weights
from syntex_b200.po import SynTexModelDesc
from syntex_b200.po.progressive_model import ProgressiveSynTex
LAYERS = list(SynTexModelDesc.DEFAULT_COEFS.keys())
raw = synthetic_vgg_weights(0)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 64

class MovedOracle:
    def __init__(self, tex): self.tex = tex
    def grams(self, image, layers=None):
        return OrderedDict((k, v.float().cuda()) for k, v in self.tex.grams(image, layers).items())
    def __call__(self, image, layers, gram_target, slot, calc_grad):
        ll, gl = self.tex(image, layers, gram_target, slot, calc_grad)
        ll = OrderedDict((k, v.float().cuda()) for k, v in ll.items())
        gl = OrderedDict((k, v.float().cuda()) for k, v in gl.items()) if gl is not None else None
        return ll, gl

def run(tag, gpu_texture=None, tf32=True, dtype=torch.float32):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.manual_seed(5)
    tex = so.OracleTexture(so.load_vgg_weights(raw), LAYERS)
    cpu = ProgressiveSynTex({"vgg19_npy_path": raw, "image_size": size}, texture_fn=tex).double()
    gpu = ProgressiveSynTex({"vgg19_npy_path": raw, "image_size": size}, texture_fn=gpu_texture)
    gpu.load_state_dict({k: v.to(dtype) for k, v in cpu.state_dict().items()})
    gpu = gpu.to(dtype).cuda()
    rng = np.random.RandomState(0)
    z = rng.uniform(-1, 1, (1, size, size, 3)); t = so.synthetic_texture(size, seed=9)[None] * 255.0
    lc = cpu.build_graph(torch.as_tensor(z), torch.as_tensor(t))
    lg = gpu.build_graph(torch.as_tensor(z).to(dtype), torch.as_tensor(t).to(dtype))
    print(tag, "losses rel diff:", ["%.2e" % (abs(float(a) - float(b)) / float(a)) for a, b in zip(cpu.losses, gpu.losses)])
    print(tag, "image diffs:", ["%.2e" % float((a - b.double().cpu()).norm() / a.norm()) for a, b in zip(cpu.image_outputs, gpu.image_outputs)])
    lc.backward(); lg.backward()
    for s in (4, 3, 2, 1, 0):
        gc = torch.cat([p.grad.flatten() for p in cpu.syn[s].parameters()])
        gg = torch.cat([p.grad.flatten() for p in gpu.syn[s].parameters()]).double().cpu()
        print(tag, "stage %d: cosine %.4f  norm ratio %.4f" % (s, float((gc * gg).sum() / (gc.norm() * gg.norm())), float(gg.norm() / gc.norm())))

moved = MovedOracle(so.OracleTexture(so.load_vgg_weights(raw), LAYERS))
#run("A cuda-texture tf32", None, True)
#run("B cuda-texture fp32", None, False)
#run("C oracle-texture+gpu-generator fp32", moved, False)
# (fp64 generator variant needs double inputs throughout; not needed)

def single_stage():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(5)
    tex = so.OracleTexture(so.load_vgg_weights(raw), LAYERS)
    cpu = ProgressiveSynTex({"vgg19_npy_path": raw, "image_size": size}, texture_fn=tex).double()
    gpu = ProgressiveSynTex({"vgg19_npy_path": raw, "image_size": size})
    gpu.load_state_dict({k: v.float() for k, v in cpu.state_dict().items()})
    gpu = gpu.cuda()
    rng = np.random.RandomState(1)
    t = torch.as_tensor(so.synthetic_texture(size, seed=9)[None])
    gt_c = tex.grams(t); gt_g = gpu._texture_fn.grams(t.float().cuda())
    for s in range(5):
        z = torch.as_tensor(rng.uniform(-1, 1, (1, size, size, 3)))
        res = []
        for model, gt, zz in ((cpu, gt_c, z.clone().requires_grad_(True)), (gpu, gt_g, z.float().cuda().requires_grad_(True))):
            model.zero_grad()
            _, lo_in, _, pre_out = model.build_stage(zz, gt, s + 1)
            ll, _ = model._texture_fn(torch.sigmoid(pre_out), LAYERS, gt, 5, False)
            loss = sum(1e5 / 4 * ll[k] for k in LAYERS).mean() + lo_in.mean()
            loss.backward()
            gp = torch.cat([p.grad.flatten() for p in model.syn[s].parameters()]).double().cpu()
            res.append((float(loss), pre_out.detach().double().cpu(), gp, zz.grad.double().cpu()))
        (lc, pc, gc, zc), (lg, pg, gg, zg) = res
        cos = lambda a, b: float((a * b).sum() / (a.norm() * b.norm()))
        print("single stage %d: loss rel %.2e  pre_out rel %.2e  dparam cos %.4f ratio %.4f relerr %.3e | dz cos %.4f ratio %.4f relerr %.3e" % (
            s, abs(lc - lg) / lc, float((pc - pg).norm() / pc.norm()), cos(gc, gg), float(gg.norm() / gc.norm()), float((gc - gg).norm() / gc.norm()),
            cos(zc, zg), float(zg.norm() / zc.norm()), float((zc - zg).norm() / zc.norm())))
single_stage()
