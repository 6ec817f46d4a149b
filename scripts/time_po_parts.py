"""Where a ProgressiveSynTex training step spends its time: hot path (VGG/Gram door) vs stock-torch generator."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200.po.progressive_model import ProgressiveSynTex
from syntex_b200.po.trainer import SynTexTrainer, SyntheticPOData
from syntex_b200.synthetic import synthetic_vgg_weights
from syntex_b200.po import texture_op
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.backends.cudnn.benchmark = True
model = ProgressiveSynTex({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": size,
                           "generator_bf16": os.environ.get("STX_PO_GENERATOR_BF16", "0") == "1"}).cuda().to(memory_format=torch.channels_last)
data = SyntheticPOData(1, size, seed=0)
batch = next(iter(data))
trainer = SynTexTrainer(data, model, 1)
F = texture_op._make_function() if texture_op._FUNCTION is None else texture_op._FUNCTION
texture_op._FUNCTION = F
acc = {"fwd": 0.0, "bwd": 0.0}
of, ob = F.forward, F.backward
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); acc[name] += time.perf_counter() - t
        return r
    return staticmethod(w)
for _ in range(3): trainer.run_step(batch)
F.forward = timed("fwd", of); F.backward = timed("bwd", ob)
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 5
for _ in range(n): trainer.run_step(batch)
torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / n
print("size %d: step %.1f ms (synchronised around the door) | hot path forward %.1f ms, backward %.1f ms | generator + optimiser %.1f ms"
      % (size, tot * 1e3, acc["fwd"] / n * 1e3, acc["bwd"] / n * 1e3, (tot - (acc["fwd"] + acc["bwd"]) / n) * 1e3))
