"""torch.profiler view of one ProgressiveSynTex training step (which stock ops the generator spends its time in)."""
import os, sys
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200.po.progressive_model import ProgressiveSynTex
from syntex_b200.po.trainer import SynTexTrainer, SyntheticPOData
from syntex_b200.synthetic import synthetic_vgg_weights
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.backends.cudnn.benchmark = True
model = ProgressiveSynTex({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": size}).cuda().to(memory_format=torch.channels_last)
data = SyntheticPOData(1, size, seed=0)
batch = next(iter(data))
graph = os.environ.get('STX_PO_GRAPH', '0') == '1'
trainer = SynTexTrainer(data, model, 1, cuda_graph=graph)
for _ in range(5): trainer.run_step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2): trainer.run_step(batch)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=int(os.environ.get('ROWS', '28')), max_name_column_width=70))
