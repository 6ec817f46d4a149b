"""world_size = 2 over gloo on CPU: the N > 1 host logic (job partition, result gather) of the sharded Gatys
workload. The f/g arithmetic is stubbed; what is under test is that every job runs exactly once and that results
are identical for every world size."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_jobs, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from syntex_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def make_job(j):
        rng = np.random.RandomState(j)
        return rng.uniform(size=(4, 4, 3)), rng.uniform(size=(4, 4, 3))

    def synth(x0s, exs):     # deterministic stand-in for Synthesizer.synthesize on a batch
        fun = ((x0s - exs) ** 2).sum(axis=(1, 2, 3))
        return 0.5 * (x0s + exs), fun, np.full(len(x0s), 100), np.full(len(x0s), 108)

    images, fun, nit, nfev = parallel.run_sharded_gatys(make_job, n_jobs, synth, batch=3, rank=rank, world=world,
                                                        dist=dist)
    np.savez(os.path.join(out_dir, "w%d_r%d.npz" % (world, rank)), fun=fun, nit=nit, nfev=nfev,
             mine=np.array(sorted(images)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_results_independent_of_world_size(tmp_path):
    n_jobs = 11
    for world in (1, 2):
        mp.spawn(_worker, args=(world, _free_port(), n_jobs, str(tmp_path)), nprocs=world, join=True)
    one = np.load(os.path.join(tmp_path, "w1_r0.npz"))
    a = np.load(os.path.join(tmp_path, "w2_r0.npz"))
    b = np.load(os.path.join(tmp_path, "w2_r1.npz"))
    np.testing.assert_array_equal(a["fun"], one["fun"])      # gathered on every rank, bit-identical to N = 1
    np.testing.assert_array_equal(b["fun"], one["fun"])
    assert sorted(list(a["mine"]) + list(b["mine"])) == list(range(n_jobs))
    assert list(a["mine"]) == list(range(0, n_jobs, 2)) and list(b["mine"]) == list(range(1, n_jobs, 2))
    assert np.all(a["nit"] == 100) and np.all(b["nfev"] == 108)
