#!/usr/bin/env python
"""Benchmark of the SynTex hot path on B200: BASELINE.json's config 4 - 64 independent Gatys L-BFGS-B synthesis jobs
at 256x256, split 64 / N per GPU with no communication (strong scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--size S]

A *step* is one cycle of the device-resident optimiser over this rank's jobs: ONE f/g evaluation of every job in
flight (VGG19 forward + Gram loss + input-gradient backward) followed by the L-BFGS-B update of every job. `value` is
the number of L-BFGS-B iterations (`nit` of scipy / of the device optimiser) all jobs of all ranks complete per second
over exactly K such steps; a job in a line search spends a step without completing an iteration
(`fg_evals_per_iteration`, ~1.06). The warm-up runs at least maxcor + 5 cycles, so the timed cycles see a full
L-BFGS history. No job reaches an iteration limit inside the timed region; after JOB_LEN = 200 steps every job is
replaced by a fresh one (restart from x0 with an empty history, inside the timed region), so every job is busy in
every step however long the run. `--batch B` overrides the 64 / N split (B jobs per GPU, weak scaling).
Weights are seeded synthetic VGG19 weights, exemplars seeded synthetic textures.

One JSON line is printed by rank 0. `value` is device-timed with the state resident in HBM; `e2e` is the same
metric through Synthesizer.synthesize() with host buffers (H2D of x0 and exemplars, target Grams, optimisation,
D2H of the result). `--impl reference` times the CPU oracle (torch-CPU fp32 restatement of the TF graph + the
reference's own scipy L-BFGS-B call) on the host cores instead. Beside the headline, every N carries
  config1_b1            BASELINE.json config 1 (ONE 256x256 job, 100 iterations) on rank 0: iterations/s
  config2_b1            BASELINE.json config 2 (ONE 512x512 job, 500 iterations) on rank 0: iterations/s
  po_config5            ProgressiveSynTex training at 512x512, per-GPU batch 1, data parallel over the N ranks (NCCL
                        gradient-mean all-reduce): images/s and the all-reduce's share of the step
  gpu_library_baseline  the same f/g evaluation through stock torch (cuDNN / cuBLAS, TF32 and bf16) on this GPU
  po_config3            SingleSynTex inference, 64 images of 256x256 sharded 64 / N per GPU: images/s
and N = 1 also `cpu_baseline`.
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
TOTAL_JOBS = 64     # BASELINE.json config 4: 64 independent 256x256 jobs, 64 / N per GPU
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
        # the roofline leg times its kernels one launch at a time in a burst of a few hundred ms at full clocks: the
        # like-for-like denominator is the BURST cuBLAS figure, not the power-capped sustained one
        return {"bf16_tflops": d.get("bf16_tflops", d.get("bf16_tflops_sustained")),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained"), "hbm_gbs": d.get("hbm_gbs"),
                "source": "MEASURED_PEAKS.json (bf16_tflops, the burst figure: kernels timed alone in a short leg)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback of B200_PROFILING.md (burst)"}


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
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
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
    return {"workload": ("config4_%d_gatys_jobs_%d" % (args.jobs_total, args.size)) if args.strong
            else "gatys_lbfgs_%d" % args.size, "image_size": args.size, "jobs_total": args.jobs_total,
            "jobs_per_gpu": args.batch,
            "gram_layers": "conv1_1,pool1,pool2,pool3,pool4", "maxcor": args.maxcor, "bounds": "[0,1]",
            "weights": "seeded synthetic VGG19 (He-normal, mean activation ~1)",
            "parallelism": "independent jobs, %s, no collective" % (
                "64 / N per GPU (strong scaling)" if args.strong else "--batch per GPU (weak scaling)"),
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
    # W untimed warm-up cycles, never fewer than maxcor + 5: the timed cycles see a FULL L-BFGS history (the optimiser's
    # passes over S / Y are at their steady-state cost)
    warm_cycles = max(args.warmup, args.maxcor + 5)
    _lib.check(L.stx_lbfgs_run(opt, never, warm_cycles, stream))
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
        reps = 10
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
        # tensor-pipe time actually spent, in kind::f16 MMA units per algorithmic MAC: the f16+f8 forward issues one
        # kind::f16 and one kind::f8f6f4 MMA per K step (the latter twice the MACs in the same time) = 2 units, the
        # input-gradient pass 1
        # ('f16f8-lean': conv4_2..conv4_4, 14.49 of the forward's 45.9 GFLOP per image, run without the correction
        # product: 1 unit there, 2 elsewhere)
        fwd_units = {"auto": 2.0, "f16f8": 2.0, "f16f8-lean": 2.0 - 14.4955 / 45.9, "bf16x3": 3.0,
                     "bf16": 1.0}.get(syn.precision, 2.0)
        exec_total = (fwd_units + 1.0) * conv_flops_one_direction(size) * batch / (conv_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "conv_patch_kernel (persistent tcgen05 3x3 implicit GEMM, fwd+dgrad, %d launches per f/g evaluation)" % conv_launches,
                    "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["bf16_tflops"],
                    # dram__bytes_read + dram__bytes_write averaged over the 22 conv launches of one evaluation in the
                    # ncu --set full capture at 64 jobs (profiles/r02_ncu_full_conv_b64_final.md: 7.24 GB read + 4.56 GB
                    # written); only meaningful at that batch with the default (f16+f8) forward
                    "traffic": 536.3e6 if (size == 256 and batch == 64) else None,
                    "traffic_note": "average DRAM bytes per conv launch, ncu --set full at 64 jobs "
                                    "(profiles/r02_ncu_full_conv_b64_final.md): the deep layers move 160-500 MB per launch "
                                    "(weights + one pass over their activation planes), the 64-channel layers 0.9-2.1 GB "
                                    "(fp16 + e4m3 planes in, gradient / feature planes in and out; ReLU masks are one bit "
                                    "per element); 11.8 GB per evaluation = 1.8 ms at the measured HBM rate against 6.6 ms "
                                    "of conv time",
                    "executed_tflops": exec_total, "executed_frac": exec_total / peaks["bf16_tflops"],
                    "note": "achieved counts algorithmic FLOPs (2*M*N*K once per direction); the f16+f8 forward spends two kind::f16 MMA times per MAC (one f16 MMA + one f8 MMA over twice the K), executed_* counts those",
                    "avg_launch_us": 1e3 * conv_ms / max(conv_launches, 1),
                    "algorithmic_gflop_per_launch_avg": flops / 1e9 / max(conv_launches, 1),
                    "frac_of_sustained": achieved / peaks["bf16_tflops_sustained"] if peaks.get("bf16_tflops_sustained") else None,
                    "peak_source": peaks["source"] + " of measured"}

    if rank != 0:
        return None
    jobs = batch * world
    value = iters_all / (ms_max * 1e-3)                    # job-iterations completed per second, whole job
    e2e_value = jobs * nit_e2e / (e2e_ms_max * 1e-3)
    evals_per_it = float(batch * args.steps) / max(iters_done, 1)      # this rank: f/g evaluations per completed iteration
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "warmup_cycles_run": warm_cycles, "ms_per_step": ms_max / args.steps,
           "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
           "vs_baseline": None, "dtype": "f16+f8/bf16", "data": "synthetic",
           "config": workload_config(args),
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


def _timed_ms(fn, warmup, steps):
    import torch
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


def config1_b1(args, size=None, IT=100, label="config1"):
    """BASELINE.json config 1 on one GPU: ONE 256x256 job, 100 L-BFGS-B iterations (rank 0). Device-timed through the
    optimiser's C entry points and end to end through Synthesizer.synthesize() with host arrays. With size = 512 and
    IT = 500: config 2 (configs[1] of BASELINE.json: the TF32-tolerance case; the tolerance itself is checked by
    tests/test_path_gpu.py and tests/test_optimizer_parity_gpu.py)."""
    import torch
    from syntex_b200 import _lib
    from syntex_b200.gatys import Synthesizer
    from syntex_b200.synthetic import synthetic_vgg_weights
    try:
        L = _lib.lib()
        stream = _lib.stream_ptr()
        size = args.size if size is None else size
        x0, targets = make_inputs(size, 1, 0)
        syn = Synthesizer({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": size, "maxiter": IT,
                           "maxcor": args.maxcor, "batch": 1})
        syn.build()
        syn.synthesize(x0[0], targets[0])                 # warm-up of the whole path (graph capture included)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, fun = syn.synthesize(x0[0], targets[0])
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        nit_e2e, nfev_e2e = int(syn.last_result["nit"][0]), int(syn.last_result["nfev"][0])
        xd = torch.from_numpy(x0).cuda()
        opt = syn._get_lbfgs(True, 0.0, 1.0)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.check(L.stx_lbfgs_reset(opt, xd.data_ptr(), 0, stream))
        torch.cuda.synchronize()
        ev0.record()
        _lib.check(L.stx_lbfgs_run(opt, IT, 0, stream))
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        syn.close()
        return {"workload": "%s: 1 job, %dx%d, %d L-BFGS-B iterations" % (label, size, size, IT),
                "iterations_per_s": nit_e2e / (ms * 1e-3), "ms_per_iteration": ms / max(nit_e2e, 1),
                "e2e_iterations_per_s": nit_e2e / e2e_s, "nit": nit_e2e, "nfev": nfev_e2e, "final_loss": float(fun)}
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, e)}


def gpu_library_baseline(args, evals_per_iteration):
    """The same f/g evaluation (VGG19 forward to pool4, five Gram losses, input gradient by autograd) through STOCK
    torch on this GPU - cuDNN convolutions, cuBLAS batched GEMMs, eager launches - in TF32 (torch's fp32 default for
    convolutions on this part) and under bf16 autocast. The reference has no GPU kernels of its own; this is the
    GPU-library baseline SURVEY.md section 8d asks for beside the CPU one. Reported per job; iterations/s uses the
    evaluations per iteration measured in the headline run."""
    import torch
    import torch.nn.functional as F
    from syntex_b200.synthetic import synthetic_vgg_weights
    try:
        dev = "cuda"
        size = args.size
        batch = min(args.batch, 16)
        torch.backends.cudnn.benchmark = True
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        raw = synthetic_vgg_weights(0)
        order = ["conv1_1", "conv1_2", "P", "conv2_1", "conv2_2", "P", "conv3_1", "conv3_2", "conv3_3", "conv3_4", "P",
                 "conv4_1", "conv4_2", "conv4_3", "conv4_4", "P"]
        gram_after = {0, 2, 5, 10, 15}                  # conv1_1, pool1, pool2, pool3, pool4
        W = {}
        for k, (w, b) in raw.items():
            w = np.asarray(w)
            if k == "conv1_1":
                w = w[:, :, ::-1, :]
            W[k] = (torch.as_tensor(np.ascontiguousarray(w)).permute(3, 2, 0, 1).contiguous().to(dev)
                    .to(memory_format=torch.channels_last), torch.as_tensor(np.asarray(b)).to(dev))
        mean = torch.tensor([123.68, 116.779, 103.939], device=dev)
        x0, _ = make_inputs(size, batch, 0)
        x = torch.from_numpy(x0).to(dev)
        targets = None

        def grams(img):
            h = (img * 255.0 - mean).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
            out = []
            for i, name in enumerate(order):
                if name == "P":
                    h = F.avg_pool2d(h, 2, 2)
                else:
                    h = F.relu(F.conv2d(h, W[name][0].to(h.dtype), W[name][1].to(h.dtype), padding=1))
                if i in gram_after:
                    f = h.flatten(2).float()                       # [B, C, HW]
                    out.append(torch.bmm(f, f.transpose(1, 2)) / f.shape[2])
            return out

        with torch.no_grad():
            targets = [g * 0.9 for g in grams(x)]

        def fg(autocast):
            xi = x.detach().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                gs = grams(xi)
            loss = sum(1e5 / 4 * ((g - t) ** 2).mean(dim=(1, 2)) for g, t in zip(gs, targets)).sum()
            (grad,) = torch.autograd.grad(loss, xi)
            return grad

        out = {"batch": batch, "what": "torch %s eager: F.conv2d (cuDNN, channels_last) + torch.bmm Grams + autograd, "
                                        "per f/g evaluation of %d jobs" % (torch.__version__, batch)}
        for name, ac in (("tf32", False), ("bf16_autocast", True)):
            ms = _timed_ms(lambda: fg(ac), 3, 10)
            out[name] = {"ms_per_eval_per_job": ms / batch,
                         "job_iterations_per_s": batch / (ms * 1e-3) / evals_per_iteration}
        return out
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, e)}


def po_config5(rank, world, local_rank):
    """BASELINE.json config 5: ProgressiveSynTex training at 512x512, per-GPU batch 1, synchronous data parallel over
    the N ranks of this run (po/model.py:114-136: mean of the tower losses; NCCL gradient-mean all-reduce, bucketed
    per stage and overlapped with the backward pass; learning rate x N, po/progressive_model.py:36-37). The whole
    step - VGG19 door, generator on the hand-written kernels, back-propagation, all-reduce, Adam - replays as one
    CUDA graph. Returns images/s (max over ranks of the device time) and the all-reduce's share of the step:
    `allreduce_alone_ms` is the exchange timed alone on the same buffer, `exposed_ms` the step time minus the step time
    of the same replica without the exchange (N = 1 of this run's record)."""
    import torch
    out = {}
    trainer = None
    try:
        import torch.distributed as dist
        from syntex_b200.po.progressive_model import ProgressiveSynTex
        from syntex_b200.po.trainer import SynTexTrainer, SyntheticPOData
        from syntex_b200.synthetic import synthetic_vgg_weights
        torch.manual_seed(0)
        model = ProgressiveSynTex({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": 512, "n_gpu": world}).cuda()
        data = SyntheticPOData(1, 512, seed=rank)
        batch = next(iter(data))
        trainer = SynTexTrainer(data, model, world, cuda_graph=True)
        for _ in range(4):
            trainer.run_step(batch)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        steps = 10
        ms_local = _timed_ms(lambda: trainer.run_step(batch), 0, steps)
        t = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        ar_ms = 0.0
        if world > 1:
            g = trainer.buckets.flat_grad
            ar_ms = _timed_ms(lambda: dist.all_reduce(g), 2, 5)
        out = {"workload": "config5: ProgressiveSynTex training 512x512, per-GPU batch 1, data parallel",
               "n_gpus": world, "images_per_s": world / ms * 1e3, "ms_per_step": ms,
               "allreduce_bytes": int(trainer.buckets.allreduce_bytes) if world > 1 else 0,
               "allreduce_alone_ms": ar_ms, "allreduce_share_of_step": ar_ms / ms,
               "generator_params": int(trainer.buckets.flat_param.numel()),
               "step": "door + generator + backward + all-reduce + Adam (TF rule) as one CUDA graph"}
    except Exception as e:      # the Gatys line is the contract; this leg is a report beside it
        out = {"error": "%s: %s" % (type(e).__name__, e)}
    finally:
        # a captured NCCL all-reduce keeps the communicator busy at teardown: drop the graph before the group goes
        try:
            if trainer is not None:
                trainer.release_graph()
            if world > 1:
                import torch.distributed as dist
                dist.barrier()
            torch.cuda.synchronize()
        except Exception:
            pass
    return out


def po_config3(rank, world):
    """BASELINE.json config 3: SingleSynTex feed-forward synthesis, batch 64 of 256x256, the images sharded 64 / N per
    GPU with no communication; images/s over all ranks, max over ranks of the device time."""
    import torch
    try:
        from syntex_b200.po.single_model import SingleSynTex
        from syntex_b200.po.trainer import SyntheticPOData
        from syntex_b200.synthetic import synthetic_vgg_weights
        torch.manual_seed(0)
        per_rank = max(64 // max(world, 1), 1)
        single = SingleSynTex({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": 256}).cuda().eval()
        z, t = next(iter(SyntheticPOData(per_rank, 256, seed=rank)))

        def infer():
            with torch.no_grad():
                single.build_graph(z, t)
        ms = _timed_ms(infer, 2, 5)
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt[0])
        return {"workload": "config3: SingleSynTex inference, 64 images of 256x256, %d per GPU" % per_rank,
                "n_gpus": world, "images_per_s": per_rank * world * 1e3 / ms, "ms_per_batch": ms}
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=25)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0,
                    help="jobs per GPU; 0 (default) = BASELINE.json config 4: 64 jobs in total, 64 / N per GPU (strong "
                         "scaling). A value B > 0 runs B jobs on every GPU instead (weak scaling)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--maxcor", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-po", action="store_true", help="skip the PO sub-records (config 5, config 3)")
    ap.add_argument("--no-extras", action="store_true", help="headline only: no config1_b1 / gpu_library_baseline / PO")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE %d" % (args.gpus, world))
    args.strong = args.batch <= 0
    if args.strong:
        if TOTAL_JOBS % max(world, 1) != 0:
            raise SystemExit("config 4 splits %d jobs evenly: --gpus must divide it" % TOTAL_JOBS)
        args.batch = TOTAL_JOBS // max(world, 1)
    args.jobs_total = args.batch * max(world, 1)

    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    out = run_ours(args, rank, world, local_rank)
    extras = not args.no_extras
    if extras and rank == 0:
        out["config1_b1"] = config1_b1(args)
        if args.size == 256:
            out["config2_b1"] = config1_b1(args, size=512, IT=500, label="config2")
        out["gpu_library_baseline"] = gpu_library_baseline(args, out["fg_evals_per_iteration"])
    if extras and not args.no_po:
        po = po_config5(rank, world, local_rank)          # every rank takes part (NCCL)
        po3 = po_config3(rank, world)
        if rank == 0:
            out["po_config5"] = po
            out["po_config3"] = po3
    if out is not None:
        out["cpu_baseline"] = cpu_baseline_sample(args) if (world == 1 and not args.no_cpu_baseline) else None
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
