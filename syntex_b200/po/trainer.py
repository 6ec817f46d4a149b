"""SynTexTrainer (po/model.py:77-136): synchronous data-parallel training of a PO model.

Reference semantics: N replicated towers, loss = mean of the tower losses, one Adam step on shared variables, learning
rate pre-multiplied by n_gpu (po/progressive_model.py:36-37), per-GPU batch BATCH. Here: one process per GPU
(torch.distributed, NCCL over NVLink), identical replicas, and a gradient MEAN all-reduce - the only exchange step of
the path (SURVEY.md section 8 row e). Parameters and gradients of the generator live in two flat buffers, one
contiguous slice per stage; the all-reduce of a stage's slice is launched from a hook as soon as the last gradient
of that stage has been accumulated, so it overlaps the back-propagation of the earlier stages (backward visits the
stages last to first).
"""
import time

import numpy as np
import torch
import torch.distributed as dist


class _FlatBuckets(object):
    """Re-homes the parameters of `groups` (list of lists of nn.Parameter) into one flat buffer, gradients into
    another, group by group; all-reduces a group's gradient slice asynchronously once all its gradients are in."""

    def __init__(self, groups, world_size):
        self.world_size = world_size
        params = [p for g in groups for p in g]
        dev, dtype = params[0].device, params[0].dtype
        total = sum(p.numel() for p in params)
        self.flat_param = torch.empty(total, device=dev, dtype=dtype)
        self.flat_grad = torch.zeros(total, device=dev, dtype=dtype)
        self.slices, self.pending, self.handles = [], [], []
        off = 0
        for gi, g in enumerate(groups):
            start = off
            for p in g:
                n = p.numel()
                self.flat_param[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat_param[off:off + n].view_as(p.data)
                p.grad = self.flat_grad[off:off + n].view_as(p.data)
                off += n
                if world_size > 1:
                    p.register_post_accumulate_grad_hook(self._make_hook(gi))
            self.slices.append((start, off))
            self.pending.append(len(g))
        self.group_sizes = [len(g) for g in groups]
        self.allreduce_bytes = total * self.flat_grad.element_size()

    def _make_hook(self, gi):
        def hook(_param):
            self.pending[gi] -= 1
            if self.pending[gi] == 0:
                a, b = self.slices[gi]
                self.handles.append(dist.all_reduce(self.flat_grad[a:b], op=dist.ReduceOp.SUM, async_op=True))
        return hook

    def zero_grad(self):
        self.flat_grad.zero_()
        self.pending = list(self.group_sizes)
        self.handles = []

    def finish(self):
        """Wait for the in-flight all-reduces and turn the sums into means (mean of tower losses, po/model.py:132)."""
        if self.world_size > 1:
            for gi, left in enumerate(self.pending):      # groups whose hooks did not all fire (unused parameters)
                if left != 0:
                    a, b = self.slices[gi]
                    self.handles.append(dist.all_reduce(self.flat_grad[a:b], op=dist.ReduceOp.SUM, async_op=True))
            for h in self.handles:
                h.wait()
            self.flat_grad.mul_(1.0 / self.world_size)


class TfAdam(object):
    """tf.train.AdamOptimizer(lr, beta1, beta2, epsilon) over the trainer's flat parameter / gradient buffers
    (po/progressive_model.py:268-271): lr_t = lr sqrt(1 - b2^t) / (1 - b1^t), p -= lr_t m / (sqrt(v) + eps).
    NOT torch.optim.Adam, whose eps is added to sqrt(v / (1 - b2^t)): with the reference's eps = 1e-3 the effective
    epsilon differs by 1 / sqrt(1 - b2^t) (30x at t = 1). On CUDA one hand-written kernel (stx_adam_tf) updates the
    whole generator; the learning rate and step count are device scalars, so a captured step replays with the current
    schedule value. On CPU (the gloo tests) the same rule in torch ops."""

    def __init__(self, flat_param, flat_grad, lr, betas=(0.9, 0.999), eps=1e-3):
        self.p, self.g = flat_param, flat_grad
        self.m = torch.zeros_like(flat_param)
        self.v = torch.zeros_like(flat_param)
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.lr_t = torch.tensor(float(lr), dtype=torch.float32, device=flat_param.device)
        self.step_t = torch.zeros((), dtype=torch.float32, device=flat_param.device)
        self.param_groups = [{"lr": self.lr_t, "betas": self.betas, "eps": self.eps}]

    def step(self):
        b1, b2 = self.betas
        self.step_t.add_(1.0)
        if self.p.is_cuda:
            from .. import _lib
            _lib.check(_lib.lib().stx_adam_tf(self.p.data_ptr(), self.g.data_ptr(), self.m.data_ptr(),
                                              self.v.data_ptr(), self.p.numel(), self.lr_t.data_ptr(),
                                              self.step_t.data_ptr(), b1, b2, self.eps, _lib.stream_ptr()))
            return
        with torch.no_grad():
            t = self.step_t
            self.m.mul_(b1).add_(self.g, alpha=1.0 - b1)
            self.v.mul_(b2).addcmul_(self.g, self.g, value=1.0 - b2)
            lr_t = self.lr_t * torch.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
            self.p.sub_(lr_t * self.m / (self.v.sqrt() + self.eps))

    def state_dict(self):
        return {"m": self.m.detach().cpu(), "v": self.v.detach().cpu(), "step": float(self.step_t),
                "lr": float(self.lr_t)}

    def load_state_dict(self, state):
        self.m.copy_(state["m"])
        self.v.copy_(state["v"])
        self.step_t.fill_(float(state["step"]))
        self.lr_t.fill_(float(state["lr"]))


class SynTexTrainer(object):
    def __init__(self, input, model, num_gpu=1, cuda_graph=False, graph_warmup=2):
        """input: iterable of (pre_image_input, image_target) batches for THIS rank; model: a PO model
        (nn.Module with build_graph / optimizer / vars); num_gpu: world size (one process per GPU).
        cuda_graph (single-GPU, CUDA only): after `graph_warmup` eager steps the whole step - VGG door, generator,
        back-propagation, Adam - is captured once and replayed; a step launches ~2000 kernels from Python and is
        otherwise paced by the host, not the GPU."""
        self.input = input
        self.model = model
        self.num_gpu = max(int(num_gpu), 1)
        self.cuda_graph = bool(cuda_graph)
        self.graph_warmup = max(int(graph_warmup), 1)
        self._graph = None
        self._static = None
        self._stream = None
        if self.num_gpu > 1:
            if not dist.is_initialized():
                raise RuntimeError("SynTexTrainer(num_gpu=%d): torch.distributed is not initialised (launch with "
                                   "torch.distributed.run, one process per GPU)" % self.num_gpu)
            if dist.get_world_size() != self.num_gpu:
                raise ValueError("num_gpu=%d but world size is %d" % (self.num_gpu, dist.get_world_size()))
            for p in model.vars:                      # identical replicas: rank 0's initial values everywhere
                dist.broadcast(p.data, src=0)
        groups = [list(stage.parameters()) for stage in model.syn] if hasattr(model, "syn") else [list(model.vars)]
        self.buckets = _FlatBuckets(groups, self.num_gpu)
        # The model's optimizer() states the hyper-parameters (Adam, epsilon 1e-3); the update itself runs as ONE
        # launch over the flat buffers with TensorFlow's rule (see TfAdam).
        g0 = model.optimizer().param_groups[0]
        self.base_lr = float(g0["lr"])
        self.opt = TfAdam(self.buckets.flat_param, self.buckets.flat_grad, self.base_lr, g0["betas"], g0["eps"])
        self.global_step = 0
        self.allreduce_wait_s = 0.0

    def set_learning_rate(self, lr):
        for g in self.opt.param_groups:
            if torch.is_tensor(g["lr"]):
                g["lr"].fill_(float(lr))
            else:
                g["lr"] = lr

    def run_step(self, batch=None):
        """One synchronous training step; returns the (local) loss as a 0-d tensor."""
        if batch is None:
            batch = next(self._iter())
        pre_image_input, image_target = batch
        if not self.cuda_graph:
            return self._step_eager(pre_image_input, image_target)
        if self.global_step >= self.graph_warmup:
            return self._run_step_graphed(pre_image_input, image_target)
        # Warm-up steps of the graph mode run on the stream the capture will use: autograd binds every parameter's
        # AccumulateGrad node to the stream of its first forward, and a node bound to the legacy default stream
        # invalidates a capture on any other stream.
        if self._stream is None:
            self._stream = torch.cuda.Stream()
        self._stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._stream):
            loss = self._step_eager(pre_image_input, image_target)
        torch.cuda.current_stream().wait_stream(self._stream)
        return loss

    def _step_eager(self, pre_image_input, image_target):
        self.buckets.zero_grad()
        loss = self.model.build_graph(pre_image_input, image_target)
        loss.backward()
        self.buckets.finish()
        self.opt.step()
        self.global_step += 1
        return loss.detach()

    def _run_step_graphed(self, pre_image_input, image_target):
        dev = self.buckets.flat_param.device
        z = torch.as_tensor(pre_image_input).to(device=dev, dtype=torch.float32)
        t = torch.as_tensor(image_target).to(device=dev, dtype=torch.float32)
        if self._graph is None:
            self._static = [z.clone(), t.clone(), None]
            self._graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            step0 = self.global_step
            if self._stream is None:
                self._stream = torch.cuda.Stream()
            with torch.cuda.graph(self._graph, stream=self._stream):
                self._static[2] = self._step_eager(self._static[0], self._static[1])
            self.global_step = step0          # capturing does not execute the step
        elif z.shape != self._static[0].shape:
            raise ValueError("cuda_graph=True: the batch shape is fixed at capture time")
        self._static[0].copy_(z)
        self._static[1].copy_(t)
        self._graph.replay()
        self.global_step += 1
        return self._static[2].clone()

    def release_graph(self):
        """Drop the captured step (and the NCCL work it references). Call before torch.distributed is shut down: a
        live CUDA graph holding captured all-reduces keeps the communicator busy and destroy_process_group() waits
        for it."""
        if self._graph is not None:
            torch.cuda.synchronize()
            self._graph = None
            self._static = None
            import gc
            gc.collect()
            torch.cuda.synchronize()

    def _iter(self):
        if not hasattr(self, "_it"):
            self._it = iter(self.input)
        return self._it

    def train_with_defaults(self, callbacks=None, max_epoch=1, steps_per_epoch=1, session_init=None,
                            start_dec_epoch=None, log_every=0):
        """Epoch loop with the reference's schedule (progressive_model.py:344-367): constant learning rate until
        start_dec_epoch (default max_epoch // 2), then linear decay to 0 at max_epoch. callbacks: callables
        (trainer, epoch) run after every epoch."""
        steps_per_epoch = max(int(steps_per_epoch), 1)
        if start_dec_epoch is None:
            start_dec_epoch = max_epoch // 2
        for epoch in range(1, int(max_epoch) + 1):
            # tensorpack's ScheduledHyperParamSetter applies the value scheduled for epoch e AFTER epoch e has run,
            # i.e. epoch e trains with the value of epoch e - 1: the last epoch still has a non-zero rate.
            if epoch - 1 > start_dec_epoch and max_epoch > start_dec_epoch:
                frac = (max_epoch - (epoch - 1)) / float(max_epoch - start_dec_epoch)
                self.set_learning_rate(self.base_lr * max(frac, 0.0))
            t0 = time.time()
            last = None
            for _ in range(steps_per_epoch):
                last = self.run_step()
            if log_every and (epoch % log_every == 0) and (self.num_gpu == 1 or dist.get_rank() == 0):
                print("epoch %d: loss %.6g  (%.2f s)" % (epoch, float(last), time.time() - t0))
            for cb in (callbacks or []):
                cb(self, epoch)
        return self


class SyntheticPOData(object):
    """Endless (pre_image_input ~ U(-1,1), image_target in [0,255]) batches of synthetic textures: the bench/test
    stand-in for get_data (progressive_model.py:274-296), which joins RandomZData([size,size,3], -1, 1) with
    augmented exemplar crops."""

    def __init__(self, batch, size, seed=0, device="cuda"):
        self.batch, self.size, self.device = batch, size, device
        self.rng = np.random.RandomState(seed)

    def __iter__(self):
        from ..synthetic import synthetic_texture
        while True:
            z = self.rng.uniform(-1, 1, size=(self.batch, self.size, self.size, 3)).astype(np.float32)
            t = np.stack([synthetic_texture(self.size, seed=int(self.rng.randint(1 << 30))) for _ in range(self.batch)])
            yield (torch.as_tensor(z).to(self.device), torch.as_tensor(t * 255.0, dtype=torch.float32).to(self.device))
