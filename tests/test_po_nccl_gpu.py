"""The one exchange step of the path on real hardware: the NCCL
gradient-mean all-reduce of the data-parallel PO trainer against the same two samples evaluated on ONE GPU.
Needs two GPUs (skipped otherwise); the CPU twin over gloo is tests/test_po_model.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model(raw_weights, n_gpu):
    from syntex_b200.po.progressive_model import ProgressiveSynTex
    torch.manual_seed(0)
    return ProgressiveSynTex({"vgg19_npy_path": raw_weights, "image_size": 64, "n_gpu": n_gpu}).cuda()


def _sample(rank):
    from syntex_b200.synthetic import synthetic_texture
    rng = np.random.RandomState(rank)
    z = torch.as_tensor(rng.uniform(-1, 1, (1, 64, 64, 3)), dtype=torch.float32).cuda()
    t = torch.as_tensor(synthetic_texture(64, seed=10 + rank)[None] * 255.0, dtype=torch.float32).cuda()
    return z, t


def _worker(rank, world, port, raw_weights, out):
    import torch.distributed as dist
    from syntex_b200.po.trainer import SynTexTrainer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        model = _model(raw_weights, world)
        trainer = SynTexTrainer(None, model, world)
        p0 = trainer.buckets.flat_param.clone()
        trainer.run_step(_sample(rank))
        torch.cuda.synchronize()
        out[rank] = (p0.cpu().numpy(), trainer.buckets.flat_grad.cpu().numpy(), trainer.buckets.flat_param.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_nccl_gradient_mean_equals_the_single_gpu_gradient_of_both_samples(raw_weights):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from syntex_b200.po.trainer import SynTexTrainer
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), raw_weights, out), nprocs=world, join=True)
        r0, r1 = out[0], out[1]
    assert np.array_equal(r0[0], r1[0])              # identical replicas before ...
    assert np.array_equal(r0[1], r1[1])              # ... the all-reduced gradient is the same on both ranks ...
    assert np.array_equal(r0[2], r1[2])              # ... and so are the replicas after the step
    # the same two samples on one GPU, one after the other: mean of the two per-sample gradients
    torch.cuda.set_device(0)
    grads = []
    for rank in range(world):
        model = _model(raw_weights, 1)
        trainer = SynTexTrainer(None, model, 1)
        trainer.buckets.zero_grad()
        model.build_graph(*_sample(rank)).backward()
        grads.append(trainer.buckets.flat_grad.double().cpu().numpy())
    want = 0.5 * (grads[0] + grads[1])
    got = r0[1].astype(np.float64)
    rel = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert rel < 1e-5, rel      # same kernels, same inputs: the only difference is the summation inside NCCL
