import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from syntex_b200.po.adaptive_model import AdaptiveSynTex
from syntex_b200.synthetic import synthetic_vgg_weights, synthetic_texture
for gen in ("torch", "auto"):
    for norm in ("batch", "instance", "none"):
        torch.manual_seed(0)
        m = AdaptiveSynTex({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": 64, "n_stage": 2, "norm_type": norm,
                            "act_type": "lrelu", "pad_type": "reflect", "generator": gen}).cuda()
        z = np.random.RandomState(0).uniform(-1, 1, (1, 64, 64, 3)).astype(np.float32)
        t = (synthetic_texture(64, seed=1)[None] * 255).astype(np.float32)
        loss = m.build_graph(torch.as_tensor(z), torch.as_tensor(t))
        loss.backward()
        g = sum(float(p.grad.abs().sum()) for p in m.syn[0].parameters() if p.grad is not None)
        print(gen, norm, "loss %.4g grad0 %.4g" % (float(loss), g), flush=True)
