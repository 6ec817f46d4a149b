"""PO configurations of BASELINE.json on the B200.

  python scripts/bench_po.py infer [batch] [size] [single|progressive]   config 3: feed-forward synthesis, images/s
  python -m torch.distributed.run --nproc-per-node N ... scripts/bench_po.py train [size] [steps]
                                                            config 5: ProgressiveSynTex training, per-GPU batch 1,
                                                            gradient-mean all-reduce over NCCL, images/s

The VGG19 / Gram work (forward, grad_layer, first- and second-order backward) and the generator (po/gen_ops.py) run
on the hand-written kernels; STX_PO_GENERATOR=torch selects the stock ATen/cuDNN generator, STX_PO_GRAPH=0/1 the
eager / CUDA-graph training step (default: graph; the NCCL all-reduce is captured with it). Synthetic textures, seeded random generator weights."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200 import _lib  # noqa: E402
from syntex_b200.po.progressive_model import ProgressiveSynTex  # noqa: E402
from syntex_b200.po.single_model import SingleSynTex  # noqa: E402
from syntex_b200.po.trainer import SynTexTrainer, SyntheticPOData  # noqa: E402
from syntex_b200.synthetic import synthetic_vgg_weights  # noqa: E402


def timed(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


GEN_BF16 = os.environ.get("STX_PO_GENERATOR_BF16", "0") == "1"
GEN = os.environ.get("STX_PO_GENERATOR", "auto")      # auto / native (hand-written kernels) / torch (stock ATen/cuDNN)


def gen_name():
    if GEN != "torch":
        return "native (tcgen05 convs incl. wgrad, fused InstanceNorm/ReLU/pad; NHWC bf16)"
    return "bf16 autocast" if GEN_BF16 else "fp32 (TF32 convs)"


def infer(batch, size, which="progressive"):
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    if which == "single":
        model = SingleSynTex({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": size, "generator": GEN})
    else:
        model = ProgressiveSynTex({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": size,
                                   "generator_bf16": GEN_BF16, "generator": GEN})
    model = model.cuda().eval().to(memory_format=torch.channels_last)
    z, t = next(iter(SyntheticPOData(batch, size, seed=0)))
    l0 = _lib.lib().stx_launch_count()

    def step():
        with torch.no_grad():
            model.build_graph(z, t)
    ms = timed(step, 2, 5)
    ms_images_only = None
    if hasattr(model, "synthesize"):
        ms_images_only = timed(lambda: model.synthesize(z, t), 2, 5)
    # VGG19 work per image: 6 forwards of the current image (5 stages to their last Gram layer + the final loss) and
    # one of the target, plus the per-layer F*D products
    print(json.dumps({"workload": "po_%s_inference" % which, "batch": batch, "size": size, "images_per_s": batch / ms * 1e3,
                      "ms_per_batch": ms,
                      "images_only": None if ms_images_only is None else {
                          "images_per_s": batch / ms_images_only * 1e3, "ms_per_batch": ms_images_only,
                          "call": "model.synthesize(): the output image without the loss of the result"},
                      "hot_path_launches_per_batch": (_lib.lib().stx_launch_count() - l0) // 7,
                      "losses": [float(l) for l in model.losses], "generator": gen_name(),
                      "note": "config 3; single = SingleSynTex on the five Gram layers (single_model.py as shipped is stale "
                              "against the current extractor, SURVEY section 2 #8); random generator weights: throughput only"}))


def train(size, steps):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    model = ProgressiveSynTex({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": size, "n_gpu": world,
                               "generator_bf16": GEN_BF16, "generator": GEN}).cuda().to(memory_format=torch.channels_last)
    data = SyntheticPOData(1, size, seed=rank)
    batch = next(iter(data))
    graph = os.environ.get("STX_PO_GRAPH", "1") == "1"
    trainer = SynTexTrainer(data, model, world, cuda_graph=graph)
    losses = []

    def step():
        losses.append(trainer.run_step(batch))
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / steps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # time of the exchange step alone (same buffers, nothing to overlap with)
    ar_ms = 0.0
    if world > 1:
        g = trainer.buckets.flat_grad
        ar_ms = timed(lambda: dist.all_reduce(g), 2, 5)
    if rank == 0:
        print(json.dumps({"workload": "po_progressive_training", "size": size, "n_gpus": world, "per_gpu_batch": 1,
                          "images_per_s": world / float(ms) * 1e3, "ms_per_step": float(ms),
                          "allreduce_bytes": trainer.buckets.allreduce_bytes, "allreduce_alone_ms": ar_ms,
                          "allreduce_share_if_exposed": ar_ms / float(ms),
                          "generator_params": int(trainer.buckets.flat_param.numel()),
                          "generator": gen_name(), "cuda_graph": graph,
                          "loss_first": float(losses[0]), "loss_last": float(losses[-1]),
                          "lr": float(trainer.opt.param_groups[0]["lr"])}))
    if world > 1:
        if graph:
            # a captured NCCL all-reduce keeps the communicator busy at teardown (destroy_process_group was seen to
            # hang until the launcher's timeout): drop the graph first; STX_PO_HARD_EXIT=1 skips the shutdown altogether
            sys.stdout.flush()
            trainer.release_graph()
            dist.barrier()
            torch.cuda.synchronize()
            if os.environ.get("STX_PO_HARD_EXIT", "0") == "1":
                os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "infer"
    if mode == "infer":
        infer(int(sys.argv[2]) if len(sys.argv) > 2 else 64, int(sys.argv[3]) if len(sys.argv) > 3 else 256,
              sys.argv[4] if len(sys.argv) > 4 else "progressive")
    else:
        train(int(sys.argv[2]) if len(sys.argv) > 2 else 512, int(sys.argv[3]) if len(sys.argv) > 3 else 10)
