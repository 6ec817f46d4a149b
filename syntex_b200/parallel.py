"""Multi-GPU partitioning of the hot path.

Gatys synthesis jobs are independent (own exemplar, seed, target Grams and L-BFGS history), so the job list is
partitioned statically, job j -> rank j mod N (the reference's analogue is the manual --batch-test-start/-end
slicing, gatys/synthesizer.py:203-204,237). There is NO collective on the data path; the only exchange is the
final gather of results. torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing.
"""
import numpy as np


def shard_jobs(n_jobs, rank, world):
    """Indices of the jobs rank `rank` of `world` runs (round robin; sizes differ by at most one)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    return list(range(rank, n_jobs, world))


def batches(indices, batch):
    """Split a rank's job list into per-launch batches of at most `batch` jobs."""
    return [indices[i:i + batch] for i in range(0, len(indices), batch)]


def gather_job_results(local, n_jobs, dist=None):
    """local: {job index: (fun: float, nit: int, nfev: int)} of this rank. Returns, on every rank, three arrays
    indexed by job. Uses one all_reduce(SUM) over zero-filled arrays (each job is owned by exactly one rank)."""
    fun = np.zeros(n_jobs, dtype=np.float64)
    nit = np.zeros(n_jobs, dtype=np.float64)
    nfev = np.zeros(n_jobs, dtype=np.float64)
    for j, (f, a, b) in local.items():
        fun[j], nit[j], nfev[j] = f, a, b
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        import torch
        t = torch.from_numpy(np.stack([fun, nit, nfev]))
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        fun, nit, nfev = t.cpu().numpy()
    return fun, nit.astype(np.int64), nfev.astype(np.int64)


def run_sharded_gatys(make_job, n_jobs, synthesize_batch, batch, rank=0, world=1, dist=None):
    """Run n_jobs independent synthesis jobs over `world` ranks.
    make_job(j) -> (x0 [S,S,3], exemplar [S,S,3]); synthesize_batch(x0s [b,S,S,3], exemplars [b,S,S,3]) ->
    (images [b,S,S,3], fun [b], nit [b], nfev [b]). Per-job results do not depend on `world` as long as the batch
    composition seen by a job does not change its arithmetic (it does not: every kernel treats images
    independently and reduces in a fixed order)."""
    mine = shard_jobs(n_jobs, rank, world)
    local, images = {}, {}
    for group in batches(mine, batch):
        x0s, exs = zip(*[make_job(j) for j in group])
        pad = batch - len(group)
        x0s, exs = list(x0s), list(exs)
        for _ in range(pad):            # keep the launch shape fixed; padded jobs are discarded
            x0s.append(x0s[-1])
            exs.append(exs[-1])
        out, fun, nit, nfev = synthesize_batch(np.stack(x0s), np.stack(exs))
        for i, j in enumerate(group):
            local[j] = (float(fun[i]), int(nit[i]), int(nfev[i]))
            images[j] = out[i]
    fun, nit, nfev = gather_job_results(local, n_jobs, dist)
    return images, fun, nit, nfev
