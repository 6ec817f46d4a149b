"""BASELINE.json config 4: a batch of 64 independent 256x256 Gatys synthesis jobs sharded over the GPUs of one node.

  python scripts/run_config4.py [--jobs 64] [--maxiter 100] [--batch 8]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/run_config4.py

Job j: exemplar j mod 47 (the reference ships 47 flower-bed exemplars; here seeded synthetic textures stand in for
the JPEGs), seed j; job j runs on rank j mod N (syntex_b200/parallel.py). No collective on the data path; one
all_reduce gathers (fun, nit, nfev). Per-job results are identical for every N.
"""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200 import parallel
from syntex_b200.gatys import Synthesizer
from syntex_b200.synthetic import init_input, synthetic_texture, synthetic_vgg_weights

ap = argparse.ArgumentParser()
ap.add_argument("--jobs", type=int, default=64)
ap.add_argument("--maxiter", type=int, default=100)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=256)
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
syn = Synthesizer({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": a.size, "maxiter": a.maxiter, "maxcor": 20,
                   "batch": a.batch})
syn.build()


_jobs = {j: (init_input((a.size, a.size, 3), False, j), synthetic_texture(a.size, seed=1000 + j % 47))
         for j in parallel.shard_jobs(a.jobs, rank, world)}          # inputs prepared outside the timed region
_jobs.update({j: (init_input((a.size, a.size, 3), False, j), synthetic_texture(a.size, seed=1000 + j % 47))
              for j in range(a.batch) if j not in _jobs})


def make_job(j):
    return _jobs[j]


def synth(x0s, exs):
    x, fun = syn.synthesize(x0s, exs)
    x = np.asarray(x).reshape(a.batch, a.size, a.size, 3)
    return x, np.atleast_1d(fun), syn.last_result["nit"], syn.last_result["nfev"]


synth(*map(np.stack, zip(*[make_job(j) for j in range(a.batch)])))      # warm-up (kernels, graph capture)
if dist is not None:
    dist.barrier()
torch.cuda.synchronize()
t = time.perf_counter()
images, fun, nit, nfev = parallel.run_sharded_gatys(make_job, a.jobs, synth, a.batch, rank, world, dist)
torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
dt = time.perf_counter() - t
if rank == 0:
    print(json.dumps({"workload": "config4_%d_jobs_%d" % (a.jobs, a.size), "n_gpus": world, "seconds": dt,
                      "job_iterations_per_s": float(nit.sum()) / dt, "fun_checksum": float(np.sum(fun)),
                      "fun_first8": [round(float(f), 4) for f in fun[:8]], "nfev_mean": float(nfev.mean())}), flush=True)
if dist is not None:
    dist.destroy_process_group()
