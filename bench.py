#!/usr/bin/env python
"""Benchmark of the SynTex hot path on B200: Gatys L-BFGS synthesis at 256x256.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--size S]

A *step* is one cycle of the device-resident optimiser over the batch: ONE f/g evaluation of every job in flight
(VGG19 forward + Gram loss + input-gradient backward) followed by the L-BFGS update of every job. `value` is the
number of L-BFGS iterations (`nit` of scipy / of the device optimiser) the jobs complete per second over exactly K
such steps; a job in a line search spends a step without completing an iteration (`fg_evals_per_iteration`, ~1.06).
No job reaches an iteration limit inside the timed region; after JOB_LEN = 200 steps every job is replaced by a fresh
one (restart from x0 with an empty history), so every job is busy in every step however long the run. The workload per GPU
is `batch` independent synthesis jobs (own exemplar, own seed, own history) run side by side; N GPUs run N x batch
jobs with no communication - weak scaling; at N = 8, batch = 8 this is BASELINE.json's config 4
(64 independent 256^2 jobs). Weights are seeded synthetic VGG19 weights, exemplars seeded synthetic textures.

One JSON line is printed by rank 0. `value` is device-timed with the state resident in HBM; `e2e` is the same
metric through Synthesizer.synthesize() with host buffers (H2D of x0 and exemplars, target Grams, optimisation,
D2H of the result). `--impl reference` times the CPU oracle (torch-CPU fp32 restatement of the TF graph + the
reference's own scipy L-BFGS-B call) on the host cores instead. At N = 1 the line also carries `po`: the PO half of
BASELINE.json's metric (ProgressiveSynTex training images/s at 512x512, SingleSynTex inference images/s at batch 64
of 256x256) measured in the same process after the Gatys timing (`--no-po` skips it).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gatys_lbfgs_job_iterations_per_sec_256"
JOB_LEN = 200       # optimiser cycles a synthesis job stays in flight before a fresh one takes its place (see run_ours)
UNIT = "iterations/s"

# Algorithmic FLOPs of the tcgen05 conv kernels per image per f/g evaluation (2*M*N*K per
# direction, conv1_2..conv4_4; conv1_1 runs on CUDA cores and is excluded), as a function of the image side.
_CONV_LAYERS = [  # (downsample, cin, cout)
    (1, 64, 64), (2, 64, 128), (2, 128, 128), (4, 128, 256), (4, 256, 256), (4, 256, 256), (4, 256, 256),
    (8, 256, 512), (8, 512, 512), (8, 512, 512), (8, 512, 512)]


def conv_flops_one_direction(size):
    return sum(2.0 * (size // d) ** 2 * 9 * ci * co for d, ci, co in _CONV_LAYERS)


def fg_eval_flops(size):
    """Whole f/g evaluation per image, canonical counting of SURVEY.md section 8 (94.41 GFLOP at 256^2)."""
    conv = conv_flops_one_direction(size) + 2.0 * size * size * 27 * 64
    gram = sum(2.0 * (size // d) ** 2 * c * c for d, c in ((1, 64), (2, 64), (4, 128), (8, 256), (16, 512)))
    return 2 * conv + 2 * gram


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "hbm_gbs": d.get("hbm_gbs"),
                "source": "MEASURED_PEAKS.json (bf16_tflops_sustained: kernels timed inside a long step)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback of B200_PROFILING.md"}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (recipe of B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()   # the exact process started above
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_inputs(size, batch, rank):
    from syntex_b200.synthetic import init_input, synthetic_texture
    base = rank * batch
    targets = np.stack([synthetic_texture(size, seed=100 + base + j) for j in range(batch)]).astype(np.float32)
    x0 = np.stack([init_input((size, size, 3), False, base + j) for j in range(batch)]).astype(np.float32)
    return x0, targets


# ----------------------------------------------------------------------------------------------- CPU reference arm
def run_reference(args, rank, world):
    """The reference's own CPU path for this workload: torch-CPU fp32 restatement of the TF graph (the TF 1.12
    runtime cannot be installed here) driven by the reference's scipy L-BFGS-B call, all host threads."""
    if rank != 0:
        return None
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import syntex_oracle as so
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    size = args.size
    x0, targets = make_inputs(size, 1, 0)
    weights = so.load_vgg_weights(so.synthetic_vgg_weights(0))
    orc = so.SynthesizerOracle(weights, size, dtype=torch.float32, maxcor=args.maxcor)
    stamps = []
    orc.synthesize(x0[0], targets[0], maxiter=args.warmup + args.steps, callback=lambda xk: stamps.append(time.time()))
    if len(stamps) < args.warmup + args.steps:
        raise RuntimeError("reference arm stopped after %d iterations" % len(stamps))
    t0 = stamps[args.warmup - 1] if args.warmup > 0 else stamps[0]
    dt = stamps[args.warmup + args.steps - 1] - t0
    steps = args.steps if args.warmup > 0 else args.steps - 1
    value = steps / dt
    sample = "1 job x %d timed L-BFGS-B iterations at %dx%d (jobs are independent: job-iterations/s does not depend on the job count)" % (steps, size, size)
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "nfev": orc.nfev}


def cpu_baseline_sample(args, seconds_budget=25.0):
    """Bounded CPU sample next to the GPU number (rank 0, N = 1): the oracle + scipy L-BFGS-B on all host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import syntex_oracle as so
    cores = os.cpu_count() or 1
    old = torch.get_num_threads()
    torch.set_num_threads(cores)
    size = args.size
    x0, targets = make_inputs(size, 1, 0)
    weights = so.load_vgg_weights(so.synthetic_vgg_weights(0))
    orc = so.SynthesizerOracle(weights, size, dtype=torch.float32, maxcor=args.maxcor)
    iters = 12
    stamps = []
    orc.synthesize(x0[0], targets[0], maxiter=iters + 2, callback=lambda xk: stamps.append(time.time()))
    dt = stamps[-1] - stamps[1]
    n = len(stamps) - 2
    out = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "1 job x %d L-BFGS-B iterations at %dx%d after 2 warm-up iterations (torch-CPU fp32 oracle + scipy)" % (n, size, size)}
    # ... and once with the reference's own thread setting: gatys/synthesizer.py:211-216 runs its TF session with
    # intra_op_parallelism_threads = inter_op_parallelism_threads = 2
    try:
        torch.set_num_threads(2)
        stamps2 = []
        orc2 = so.SynthesizerOracle(weights, size, dtype=torch.float32, maxcor=args.maxcor)
        orc2.synthesize(x0[0], targets[0], maxiter=6, callback=lambda xk: stamps2.append(time.time()))
        if len(stamps2) > 2:
            out["reference_thread_setting"] = {"value": (len(stamps2) - 2) / (stamps2[-1] - stamps2[1]), "unit": UNIT,
                                               "cores": 2, "sample": "%d iterations after 2 warm-up iterations" % (len(stamps2) - 2)}
    except Exception as e:      # a report beside the report: never fatal
        out["reference_thread_setting"] = {"error": "%s: %s" % (type(e).__name__, e)}
    torch.set_num_threads(old)
    return out


def workload_config(args):
    return {"workload": "gatys_lbfgs_%d" % args.size, "image_size": args.size, "jobs_per_gpu": args.batch,
            "gram_layers": "conv1_1,pool1,pool2,pool3,pool4", "maxcor": args.maxcor, "bounds": "[0,1]",
            "weights": "seeded synthetic VGG19 (He-normal, mean activation ~1)",
            "parallelism": "independent jobs per GPU, no collective",
            "l2": "inputs larger than L2: %d jobs x ~95 MB of activations + L-BFGS history per GPU" % args.batch}


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    from syntex_b200 import _lib
    from syntex_b200.gatys import Synthesizer
    from syntex_b200.synthetic import synthetic_vgg_weights

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on STDOUT (NCCL_DEBUG is set in this image): send its log to stderr so that
        # stdout carries only the one JSON line the contract asks for
        # (NCCL only honours NCCL_DEBUG_FILE above the VERSION level, where the banner is printed straight to stdout)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = _lib.lib()
    size, batch = args.size, args.batch
    x0, targets = make_inputs(size, batch, rank)
    syn = Synthesizer({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": size, "maxiter": args.steps,
                       "maxcor": args.maxcor, "batch": batch})
    syn.build()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- e2e: the public call with host buffers (also serves as warm-up of every kernel) -------------------
    syn.options["maxiter"] = max(args.warmup, 3)
    syn.synthesize(x0, targets)                      # untimed warm-up of the whole path
    # the user-facing call runs each job to its own iteration limit; JOB_LEN bounds it so that a long --steps does not
    # turn the e2e leg into a convergence study (near the noise floor of the loss a job's line search gives up)
    syn.options["maxiter"] = min(args.steps, JOB_LEN)
    barrier()
    t0 = time.perf_counter()
    x_e2e, fun_e2e = syn.synthesize(x0, targets)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    nit_e2e = float(np.mean(syn.last_result["nit"]))       # iterations completed per job (== maxiter unless a job stalled)
    h2d = x0.nbytes + targets.nbytes
    d2h = x0.nbytes + 4 * batch + 8 * batch

    # ---- device-resident timed region -------------------------------------------------------------------
    stream = _lib.stream_ptr()
    xd = torch.from_numpy(x0).cuda()
    syn.compute_gram(targets, update=True)
    opt = syn._get_lbfgs(True, 0.0, 1.0)
    fun = torch.empty(batch, dtype=torch.float32, device="cuda")
    nit = torch.empty(batch, dtype=torch.int32, device="cuda")
    nfev = torch.empty(batch, dtype=torch.int32, device="cuda")

    def counters():
        _lib.check(L.stx_lbfgs_result(opt, None, fun.data_ptr(), nit.data_ptr(), nfev.data_ptr(), stream))
        torch.cuda.synchronize()
        return nit.cpu().numpy().astype(np.int64), nfev.cpu().numpy().astype(np.int64)

    # A step is one cycle of the device optimiser: ONE f/g evaluation of every job in flight (VGG19 forward, Gram
    # loss, input-gradient backward) followed by the L-BFGS update of every job. No job ever reaches its iteration
    # limit inside the timed region, so all jobs are busy in every cycle; the throughput is the number of L-BFGS
    # iterations the jobs completed over those K cycles (a job in a line search spends a cycle without completing one).
    never = 1 << 30
    _lib.check(L.stx_lbfgs_reset(opt, xd.data_ptr(), 0, stream))
    _lib.check(L.stx_lbfgs_run(opt, never, max(args.warmup, 1), stream))      # W untimed warm-up cycles
    nit0, nfev0 = counters()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = L.stx_launch_count()
    t_mark0 = time.time()
    ev0.record()
    # Exactly K cycles, in job lengths of at most JOB_LEN: after JOB_LEN cycles every job is replaced by a fresh one
    # (same exemplar, restart from x0, empty history) - a stream of synthesis jobs. Without that, a long --steps drives
    # the 256x256 jobs to the noise floor of the loss, where line searches fail and the optimiser of a job gives up
    # (6 of 18 jobs within 600 cycles). The restart (one copy of x0, inside the timed region) and the counter read-back
    # at the job boundary are part of the measurement.
    done, iters_done, jobs_stalled = 0, 0, 0
    nit_prev, nfev_prev = nit0, nfev0
    while done < args.steps:
        c = min(JOB_LEN, args.steps - done)
        if done > 0:
            _lib.check(L.stx_lbfgs_reset(opt, xd.data_ptr(), 0, stream))
            nit_prev, nfev_prev = 0 * nit0, 0 * nfev0
        _lib.check(L.stx_lbfgs_run(opt, never, c, stream))
        done += c
        if done < args.steps:
            nit_c, nfev_c = counters()
            iters_done += int((nit_c - nit_prev).sum())
            jobs_stalled += int(((nfev_c - nfev_prev) != c).sum())
    ev1.record()
    barrier()
    t_mark1 = time.time()
    launches = L.stx_launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_mark0, t_mark1) if rank == 0 else None
    nit1, nfev1 = counters()
    fun_h = fun.cpu().numpy()
    # every job is evaluated in every cycle unless its optimiser gave up (line search failed along steepest descent):
    # such a job stops counting iterations, which only lowers the reported throughput; it is reported, not hidden
    iters_done += int((nit1 - nit_prev).sum())             # L-BFGS iterations completed by this rank's jobs
    jobs_stalled += int(((nfev1 - nfev_prev) != c).sum())

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    it_all = torch.tensor([iters_done], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(it_all, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])
    iters_all = float(it_all[0])                           # over all ranks

    # ---- roofline of the dominant kernel family (tcgen05 conv), measured live with CUDA events ------------
    roofline, kernel_ms = None, None
    if rank == 0:
        ms_cls = (ctypes.c_float * 8)()
        cnt_cls = (ctypes.c_int * 8)()
        g = torch.empty_like(xd)
        f = torch.empty(batch, dtype=torch.float32, device="cuda")
        acc = np.zeros(8)
        reps = 5
        for _ in range(reps):
            _lib.check(L.stx_plan_step_timed(syn._plan.handle, xd.data_ptr(), 0, f.data_ptr(), g.data_ptr(), ms_cls,
                                             cnt_cls, stream))
            acc += np.array(list(ms_cls))
        acc /= reps
        counts = list(cnt_cls)
        names = ["conv_fwd", "conv_dgrad", "gram_grad", "gram", "gram_finish", "conv1_1", "pool", "other"]
        kernel_ms = {n: {"ms_per_eval": round(float(a), 4), "launches": c} for n, a, c in zip(names, acc, counts)}
        conv_ms = float(acc[0] + acc[1])
        conv_launches = counts[0] + counts[1]
        flops = 2.0 * conv_flops_one_direction(size) * batch          # fwd + dgrad, algorithmic (2*M*N*K)
        peaks = measured_peaks()
        achieved = flops / (conv_ms * 1e-3) / 1e12
        # the accurate forward issues three bf16 MMAs per algorithmic MAC (bf16x3): tensor-pipe work actually executed
        exec_total = (3.0 + 1.0) * conv_flops_one_direction(size) * batch / (conv_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "conv_patch_kernel (persistent tcgen05 3x3 implicit GEMM, fwd+dgrad, %d launches per f/g evaluation)" % conv_launches,
                    "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["bf16_tflops"],
                    # dram__bytes_read + dram__bytes_write averaged over the 22 conv launches of one evaluation in the
                    # ncu --set full captures at batch 18 (profiles/r01_ncu_full_b18.md) and batch 9
                    # (profiles/r01_ncu_full_final_b9.md); only meaningful at those batches
                    "traffic": {18: 139.8e6, 9: 61.8e6}.get(batch) if size == 256 else None,
                    "traffic_note": "average DRAM bytes per conv launch, ncu --set full (profiles/r01_ncu_full_b18.md, "
                                    "r01_ncu_full_final_b9.md): at batch 18 the deep layers move 47-80 MB per launch "
                                    "(weights + one pass over their activations), the 64-channel layers 300-490 MB "
                                    "(hi/lo activation planes in and out)",
                    "executed_tflops": exec_total, "executed_frac": exec_total / peaks["bf16_tflops"],
                    "note": "achieved counts algorithmic FLOPs (2*M*N*K once per direction); the bf16x3 forward executes 3 MMAs per MAC, executed_* counts those",
                    "avg_launch_us": 1e3 * conv_ms / max(conv_launches, 1),
                    "algorithmic_gflop_per_launch_avg": flops / 1e9 / max(conv_launches, 1),
                    "peak_source": peaks["source"] + " of measured"}

    if rank != 0:
        return None
    jobs = batch * world
    value = iters_all / (ms_max * 1e-3)                    # job-iterations completed per second, whole job
    e2e_value = jobs * nit_e2e / (e2e_ms_max * 1e-3)
    evals_per_it = float(batch * args.steps) / max(iters_done, 1)      # this rank: f/g evaluations per completed iteration
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
           "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / max(nit_e2e, 1),
                   "d2h_bytes_per_step": d2h / max(nit_e2e, 1),
                   "call": "Synthesizer.synthesize(x0, exemplars) with host arrays, %d iterations per job" % round(nit_e2e)},
           "roofline": roofline, "kernel_ms_per_eval": kernel_ms,
           "fg_evals_per_iteration": evals_per_it, "jobs_stalled": jobs_stalled,
           "tflops_algorithmic": value * evals_per_it * fg_eval_flops(size) / 1e12,
           "final_loss_mean": float(fun_h.mean()),
           "per_gpu_value": value / world}
    return out


def po_sample():
    """po_sample_inprocess() in a child process: a failure there (even a crash of the CUDA context) cannot take the
    Gatys line with it."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--po-child"], capture_output=True, text=True,
                           timeout=420)
        lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": "PO child exited %d: %s" % (r.returncode, (r.stderr or "").strip()[-300:])}
        return json.loads(lines[-1])
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, e)}


def po_sample_inprocess():
    """The PO half of BASELINE.json's metric ("PO imgs/sec"), measured beside the Gatys line on one GPU: config 5
    (ProgressiveSynTex training at 512x512, per-GPU batch 1, whole step replayed as a CUDA graph) and config 3
    (SingleSynTex feed-forward synthesis, batch 64 of 256x256). Synthetic textures, seeded random generator weights,
    generator on the hand-written kernels (syntex_b200/po/gen_ops.py). scripts/bench_po.py is the stand-alone /
    multi-GPU version. Never raises: a failure is reported in the returned dict."""
    import torch
    out = {}
    try:
        from syntex_b200.po.progressive_model import ProgressiveSynTex
        from syntex_b200.po.single_model import SingleSynTex
        from syntex_b200.po.trainer import SynTexTrainer, SyntheticPOData
        from syntex_b200.synthetic import synthetic_vgg_weights

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

        torch.manual_seed(0)
        model = ProgressiveSynTex({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": 512}).cuda()
        data = SyntheticPOData(1, 512, seed=0)
        batch = next(iter(data))
        trainer = SynTexTrainer(data, model, 1, cuda_graph=True)
        ms = timed(lambda: trainer.run_step(batch), 4, 10)
        out["progressive_training_512"] = {"images_per_s": 1e3 / ms, "ms_per_step": ms, "per_gpu_batch": 1,
                                           "generator_params": int(trainer.buckets.flat_param.numel()),
                                           "step": "door + generator + backward + fused Adam as one CUDA graph"}
        del trainer, model
        torch.cuda.empty_cache()
        single = SingleSynTex({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": 256}).cuda().eval()
        z, t = next(iter(SyntheticPOData(64, 256, seed=0)))

        def infer():
            with torch.no_grad():
                single.build_graph(z, t)
        ms = timed(infer, 2, 5)
        out["single_inference_b64_256"] = {"images_per_s": 64e3 / ms, "ms_per_batch": ms}
        del single
        torch.cuda.empty_cache()
    except Exception as e:      # the Gatys line above is the contract; this leg is a report beside it
        out["error"] = "%s: %s" % (type(e).__name__, e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=18,
                    help="independent synthesis jobs per GPU (a multiple of 9: the deep layers then have 32 and 64 work "
                         "items per job for the 148 / 296 persistent CTAs, 97 %% full waves; 18 measured 4 %% above 9 - "
                         "4584 vs 4401 job-iterations/s, conv roofline 0.45 vs 0.40 - and 27 another 0.5 %%)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--maxcor", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-po", action="store_true", help="skip the PO images/s report beside the Gatys line (N = 1)")
    ap.add_argument("--po-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.po_child:
        print(json.dumps(po_sample_inprocess()), flush=True)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE %d" % (args.gpus, world))

    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    out = run_ours(args, rank, world, local_rank)
    if out is not None:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_sample(args)
        else:
            out["cpu_baseline"] = None
        if world == 1 and not args.no_po:
            out["po"] = po_sample()
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
