"""Localise batch-dependent backward errors: per-loss-layer total gradients, batch vs single (GPU bring-up)."""
import os, sys
from collections import OrderedDict
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200.po import build_stage_preparation
from syntex_b200.texture_utils import Vgg19Extractor, build_gram
from syntex_b200.synthetic import init_input, synthetic_texture, synthetic_vgg_weights
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
coefs = OrderedDict([("conv1_1", 1e5), ("pool1", 1e5), ("pool2", 1e5), ("pool3", 1e5), ("pool4", 1e5)])
from syntex_b200 import _lib
for kv in sys.argv[3:]:
    k, v = kv.split('=')
    _lib.check(_lib.lib().stx_debug_set(int(k), int(v)))
ext = Vgg19Extractor(synthetic_vgg_weights(0))
tg = torch.as_tensor(np.stack([synthetic_texture(S, seed=100 + j) for j in range(B)]).astype(np.float32)).cuda()
x0 = torch.as_tensor(np.stack([init_input((S, S, 3), False, j) for j in range(B)]).astype(np.float32)).cuda()
ft = ext.build(tg)
gt = OrderedDict((k, build_gram(ft[k])) for k in coefs)
_, _, llb, gb = build_stage_preparation(ext, x0, gt, coefs)
for j in sorted(set([0, B // 4, B // 2, 3 * B // 4, B - 1])):
    g1t = OrderedDict((k, v[j:j + 1].contiguous()) for k, v in gt.items())
    _, _, ll1, g1 = build_stage_preparation(ext, x0[j:j + 1].contiguous(), g1t, coefs)
    print("img %2d:" % j, " ".join("%s %.2e" % (k, float((gb[k][j] - g1[k][0]).norm() / g1[k][0].norm())) for k in coefs), flush=True)
