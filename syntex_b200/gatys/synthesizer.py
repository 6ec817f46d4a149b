"""Host mirror of gatys/synthesizer.py: Gatys et al. optimisation-based texture synthesis.

Same class, method names, argument order and defaults as the reference (gatys/synthesizer.py:16-159). The TF
session/graph is replaced by one prepared CUDA plan (syntex_b200/csrc/plan.cu); `sess` is accepted and ignored.
Two optimisers drive the same f/g evaluation:
  args["optimizer"] = "device" (default)  the GPU-resident L-BFGS-B (csrc/lbfgs.cu): x, history and line
                                          search stay in HBM, no host round trip per iteration;
  args["optimizer"] = "scipy"             the reference's own call, scipy L-BFGS-B over Synthesizer.step with host
                                          buffers (synthesizer.py:149-150), for like-for-like comparisons.
args["batch"] (default 1) runs that many independent synthesis jobs side by side on one GPU (each image of the
batch has its own target Grams, history and line search); the reference's shape is the batch == 1 case.
"""
import ctypes
from collections import OrderedDict, defaultdict

import numpy as np

from .. import _lib
from ..aparse import ArgParser
from ..texture_utils import Vgg19Extractor
from ..texture_utils.vgg19 import LAYER_INDEX


def get_bounds(shape=None):
    """synthesizer.py:10-14 (vectorised: an [n,2] array instead of an n-element list of lists)."""
    n = int(np.prod(shape))
    return np.tile(np.array([[0., 1.]]), (n, 1))


import os

# Forward precision of the f/g plan when args["precision"] is not given; STX_PRECISION overrides (bring-up).
DEFAULT_PRECISION = os.environ.get("STX_PRECISION", "auto")


class Synthesizer(object):
    DEFAULT_COEFS = OrderedDict([   # synthesizer.py:17-23
        ("conv1_1", 1e5),
        ("pool1", 1e5),
        ("pool2", 1e5),
        ("pool3", 1e5),
        ("pool4", 1e5)
    ])

    def __init__(self, args, sess=None):
        self.vgg19_path = args.get("vgg19_npy_path")
        self.image_size = args.get("image_size")
        self.gram_only = args.get("gram_only")
        self.verbose = args.get("verbose")
        self.options = {
            "maxiter": args["maxiter"],
            "maxcor": args["maxcor"],
            "disp": args.get("disp", False),
            "ftol": 0.,
            "gtol": 0.
        }
        self.optimizer = args.get("optimizer") or "device"
        if self.optimizer not in ("device", "scipy"):
            raise ValueError("optimizer must be 'device' or 'scipy'")
        self.batch = int(args.get("batch") or 1)
        self.precision = args.get("precision") or DEFAULT_PRECISION
        if self.precision not in ("auto", "f16f8", "f16f8-lean", "bf16x3", "bf16"):
            raise ValueError("precision must be 'auto', 'f16f8', 'bf16x3' (accurate forwards), 'f16f8-lean' or 'bf16' "
                             "(faster forwards)")
        self.sess = sess
        self._plan = None
        self._lbfgs = {}
        self.last_result = None

    @staticmethod
    def get_parser(ps=None):
        ps = ArgParser(ps, name="synthesizer")
        ps.add("--vgg19-npy-path")
        ps.add("--image-size", type=int, default=256)
        ps.add_flag("--gram-only")
        #
        ps.add("--maxiter", type=int, default=1000,
               help="Max iterations of optimizer")
        ps.add("--maxcor", type=int, default=20,
               help="Max number of history gradients and points")
        ps.add_flag("--disp")
        ps.add_flag("--verbose")
        # additions of this implementation
        ps.add("--optimizer", type=str, default="device", help="'device' (GPU L-BFGS) or 'scipy' (reference)")
        ps.add("--batch", type=int, default=1, help="independent jobs run side by side on one GPU")
        ps.add("--precision", type=str, default=DEFAULT_PRECISION,
               help="'auto' / 'f16f8' / 'bf16x3' (accurate forwards), 'f16f8-lean' or 'bf16' (faster)")
        return ps

    def build(self):
        """synthesizer.py:54-107: the whole f/g graph, here one CUDA plan."""
        self.vgg19 = Vgg19Extractor(self.vgg19_path)
        self.shape_input = [self.batch, self.image_size, self.image_size, 3]
        coefs = OrderedDict() if self.gram_only and False else Synthesizer.DEFAULT_COEFS
        # 'auto': the f16 + f8 forward wherever it exists (every convolution's map >= 8 pixels wide: image side a
        # multiple of 16 and >= 64), else bf16x3 - the same accuracy class, two MMAs' time per MAC instead of three
        # 'f16f8-lean': conv4_2..conv4_4 without the correction product (STX_PLAN_F16F8_LEAN): one f/g evaluation stays
        # inside the bars (gradient unchanged, pool4 Gram 5e-4 instead of 1e-4) and the forward is 14 % shorter, but
        # the rounding noise of the loss comes back to where the line search of a LONG run feels it (1000 iterations:
        # 2 of 8 jobs stop early, final loss 5-7 % above scipy's) - opt-in, not the default
        f16f8 = self.precision in ("f16f8", "f16f8-lean") or (self.precision == "auto" and self.image_size >= 64
                                                              and self.image_size % 16 == 0)
        if f16f8 and self.precision == "f16f8-lean":
            f16f8 = "lean"
        self._plan = self.vgg19.get_plan(self.batch, self.image_size, self.image_size, "pool4", coefs,
                                         fast_bf16=(self.precision == "bf16"), f16f8=f16f8)
        self._have_target = False
        self.coef_names = list(coefs.keys())

    # ------------------------------------------------------------------ helpers
    def _to_batch(self, image):
        image = np.asarray(image)
        if image.ndim == 3:
            image = image[np.newaxis, ...]
        if image.shape[0] == 1 and self.batch > 1:
            image = np.repeat(image, self.batch, axis=0)
        if list(image.shape) != self.shape_input:
            raise ValueError("expected input of shape %s, got %s" % (self.shape_input, list(image.shape)))
        # a float64 array fed to the reference's float32 placeholder is rounded to fp32 the same way
        return np.ascontiguousarray(image, dtype=np.float32)

    def _set_target(self, image, is_preimage=False):
        """compute_gram(image, update=True) without the device-to-host copy of the Grams it returns (12.7 MB at batch
        9): what synthesize() needs."""
        import torch
        if self.gram_only:
            raise AttributeError("'Synthesizer' object has no attribute 'op_update_gram_target'")
        img = torch.from_numpy(self._to_batch(image)).cuda()
        L = _lib.lib()
        st = _lib.stream_ptr()
        _lib.check(L.stx_plan_forward(self._plan.handle, img.data_ptr(), int(bool(is_preimage)), 1, st))
        _lib.check(L.stx_plan_update_target_from_forward(self._plan.handle, st))
        self._have_target = True

    def compute_gram(self, image, mask=None, update=False, is_preimage=False):
        """synthesizer.py:109-117. Returns {layer: [B,C,C]} or, with update=True, assigns the target Grams
        (the reference's op_update_gram_target) and returns them as a list."""
        import torch
        img = torch.from_numpy(self._to_batch(image)).cuda()
        L = _lib.lib()
        st = _lib.stream_ptr()
        _lib.check(L.stx_plan_forward(self._plan.handle, img.data_ptr(), int(bool(is_preimage)), 1, st))
        grams = OrderedDict()
        for k, name in enumerate(Synthesizer.DEFAULT_COEFS):
            _, _, c = self._plan.layer_shape(LAYER_INDEX[name])
            g = torch.empty((self.batch, c, c), dtype=torch.float32, device="cuda")
            _lib.check(L.stx_plan_copy_gram(self._plan.handle, k, g.data_ptr(), st))
            grams[name] = g
        if update:
            if self.gram_only:
                raise AttributeError("'Synthesizer' object has no attribute 'op_update_gram_target'")
            _lib.check(L.stx_plan_update_target_from_forward(self._plan.handle, st))
            self._have_target = True
            return [g.cpu().numpy() for g in grams.values()]
        return OrderedDict((k, g.cpu().numpy()) for k, g in grams.items())

    def step(self, input, mask=None, target_image=None, is_preimage=False):
        """synthesizer.py:119-125: one f/g evaluation through host buffers.
        First call 'compute_gram' with 'update=True' to load target gram."""
        x = self._to_batch(input)
        func = np.empty((self.batch,), dtype=np.float32)
        grad = np.empty(self.shape_input, dtype=np.float32)
        _lib.check(_lib.lib().stx_plan_step_host(self._plan.handle, x.ctypes.data, int(bool(is_preimage)),
                                                 func.ctypes.data, grad.ctypes.data, _lib.stream_ptr()))
        return {"grad": grad, "func": func}

    def _get_lbfgs(self, bounded, lo, hi):
        key = (bool(bounded), float(lo), float(hi), int(self.options["maxcor"]))
        h = self._lbfgs.get(key)
        if h is None:
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().stx_lbfgs_create(ctypes.byref(h), self._plan.handle, key[3], int(key[0]),
                                                   ctypes.c_float(lo), ctypes.c_float(hi)))
            self._lbfgs[key] = h
        return h

    def synthesize(self, input_init, image_target, mask_input=None, mask_target=None, bounds=None,
                   is_preimage=False):
        """synthesizer.py:127-151. Returns (image [S,S,3] (or [B,S,S,3] for batch > 1), fun)."""
        if self.verbose:
            self.reset_verbose()
        self._set_target(image_target)
        x0 = self._to_batch(input_init)
        if self.optimizer == "scipy":
            return self._synthesize_scipy(x0, bounds, is_preimage)
        if bounds is not None:
            b = np.asarray(bounds, dtype=np.float64).reshape(-1, 2)
            lo, hi = float(b[0, 0]), float(b[0, 1])
            if not (np.all(b[:, 0] == lo) and np.all(b[:, 1] == hi)):
                raise NotImplementedError("device optimizer supports one [lower, upper] box for all pixels; "
                                          "use optimizer='scipy' for per-pixel bounds")
            bounded = True
        else:
            bounded, lo, hi = (not is_preimage), 0., 1.
        opt = self._get_lbfgs(bounded, lo, hi)
        if self.verbose:
            return self._synthesize_device_verbose(opt, x0, is_preimage)
        # the reference returns scipy's float64 vector; the library widens while copying out of its staging buffer
        x = np.empty(self.shape_input, dtype=np.float64)
        fun = np.empty((self.batch,), dtype=np.float32)
        nit = np.empty((self.batch,), dtype=np.int32)
        nfev = np.empty((self.batch,), dtype=np.int32)
        _lib.check(_lib.lib().stx_lbfgs_minimize_host_f64(opt, x0.ctypes.data, int(bool(is_preimage)),
                                                          int(self.options["maxiter"]), x.ctypes.data, fun.ctypes.data,
                                                          nit.ctypes.data, nfev.ctypes.data, _lib.stream_ptr()))
        self.last_result = {"nit": nit, "nfev": nfev, "fun": fun}
        if self.batch == 1:
            return x.reshape(self.shape_input[1:]), float(fun[0])
        return x, fun

    def _synthesize_scipy(self, x0, bounds, is_preimage):
        from scipy.optimize import minimize

        def f(x):
            x = x.reshape(self.shape_input)
            ret = self.step(x, is_preimage=is_preimage)
            # We want the magnitude of grad to be independent of the batchsize
            return [np.sum(ret["func"]), np.array(ret["grad"].ravel(), dtype=np.float64)]

        def verbose_callback(x, *args, **kwargs):
            self._verbose_record(x.reshape(self.shape_input), is_preimage)

        cb = verbose_callback if self.verbose else None
        if bounds is None and (not is_preimage):
            bounds = get_bounds(self.shape_input)
        options = {k: v for k, v in self.options.items() if k != "disp"}
        res = minimize(f, x0.astype(np.float64).ravel(), method="L-BFGS-B", jac=True,
                       bounds=bounds, options=options, callback=cb)
        self.last_result = {"nit": np.array([res.nit]), "nfev": np.array([res.nfev]), "fun": np.array([res.fun])}
        if self.batch == 1:
            return res["x"].reshape(self.shape_input[1:]), res["fun"]
        return res["x"].reshape(self.shape_input), res["fun"]

    def _verbose_record(self, x, is_preimage):
        """The verbose fetches of the reference graph (synthesizer.py:80-89, :99-107) at `x`, appended to
        verbose_dict: func, gi_all and, per Gram layer k, l_k (loss), a_k (mean |activation|), gl_k (mean
        |d loss_k / d feat_k|), gi_k (mean |d loss_k / d input|: one extra back-propagation per layer)."""
        import torch
        L = _lib.lib()
        st = _lib.stream_ptr()
        plan = self._plan
        xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda()
        B = self.batch
        K = len(self.coef_names)
        func = torch.empty((B,), dtype=torch.float32, device="cuda")
        grad = torch.empty_like(xd)
        layer_loss = torch.empty((K, B), dtype=torch.float32, device="cuda")
        pre = int(bool(is_preimage))
        _lib.check(L.stx_plan_step(plan.handle, xd.data_ptr(), pre, func.data_ptr(), grad.data_ptr(),
                                   layer_loss.data_ptr(), st))
        rec = {"func": float(func.sum()), "gi_all": float(grad.abs().mean())}
        diffs = []
        for k, name in enumerate(self.coef_names):
            h, w, c = plan.layer_shape(LAYER_INDEX[name])
            rec["l_" + name] = float(layer_loss[k].sum())
            feat = torch.empty((B, h, w, c), dtype=torch.float32, device="cuda")
            _lib.check(L.stx_plan_copy_feature_f32(plan.handle, LAYER_INDEX[name], feat.data_ptr(), st))
            rec["a_" + name] = float(feat.abs().mean())
            g = torch.empty((B, c, c), dtype=torch.float32, device="cuda")
            t = torch.empty_like(g)
            _lib.check(L.stx_plan_copy_gram(plan.handle, k, g.data_ptr(), st))
            _lib.check(L.stx_plan_copy_target_gram(plan.handle, k, t.data_ptr(), st))
            norm = float(np.float32(h * w) + np.float32(1e-6))
            diffs.append(((g - t) * (4.0 / (c * c * norm))).contiguous())
        for k, name in enumerate(self.coef_names):
            h, w, c = plan.layer_shape(LAYER_INDEX[name])
            gl = torch.empty((B, h, w, c), dtype=torch.float32, device="cuda")
            _lib.check(L.stx_plan_layer_grad(plan.handle, k, gl.data_ptr(), st))
            rec["gl_" + name] = float(gl.abs().mean())
        for k, name in enumerate(self.coef_names):
            # d sum(loss_layer[k]) / d input: back-propagate F_k * (alpha (G_k - T_k)) alone
            mats = [diffs[j] if j == k else torch.zeros_like(diffs[j]) for j in range(K)]
            Mp = (ctypes.c_void_p * K)(*[m.data_ptr() for m in mats])
            gk = torch.empty_like(xd)
            _lib.check(L.stx_plan_backward_custom(plan.handle, xd.data_ptr(), pre, Mp, None, gk.data_ptr(), st))
            rec["gi_" + name] = float(gk.abs().mean())
        for key, v in rec.items():
            self.verbose_dict[key].append(v)
        return rec

    def _synthesize_device_verbose(self, opt, x0, is_preimage):
        """The device optimiser advanced one iteration at a time so that the verbose record can be taken after every
        iteration, as the reference's scipy callback does (diagnostic mode: slower than the resident loop)."""
        import torch
        L = _lib.lib()
        st = _lib.stream_ptr()
        B = self.batch
        x0d = torch.as_tensor(x0).cuda()
        xd = torch.empty_like(x0d)
        fun = torch.empty((B,), dtype=torch.float32, device="cuda")
        nit = torch.empty((B,), dtype=torch.int32, device="cuda")
        nfev = torch.empty((B,), dtype=torch.int32, device="cuda")
        _lib.check(L.stx_lbfgs_reset(opt, x0d.data_ptr(), int(bool(is_preimage)), st))
        maxiter = int(self.options["maxiter"])
        for it in range(1, maxiter + 1):
            _lib.check(L.stx_lbfgs_run(opt, it, 0, st))
            _lib.check(L.stx_lbfgs_result(opt, xd.data_ptr(), fun.data_ptr(), nit.data_ptr(), nfev.data_ptr(), st))
            self._verbose_record(xd.cpu().numpy(), is_preimage)
            if int(nit.max()) < it:      # every job has stopped early (line search could not make progress)
                break
        self.last_result = {"nit": nit.cpu().numpy(), "nfev": nfev.cpu().numpy(), "fun": fun.cpu().numpy()}
        x = xd.cpu().numpy()
        if self.batch == 1:
            return x.reshape(self.shape_input[1:]).astype(np.float64), float(fun[0])
        return x.astype(np.float64), fun.cpu().numpy()

    def reset_verbose(self):
        if self.verbose:
            self.verbose_dict = defaultdict(list)

    def save_verbose(self, save_path):
        if self.verbose:
            np.savez(save_path, **self.verbose_dict)

    def close(self):
        for h in self._lbfgs.values():
            _lib.lib().stx_lbfgs_destroy(h)
        self._lbfgs = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def resize_crop(img, size, interp=None):
    """synthesizer.py:161-176: resize shortest edge and center crop."""
    import cv2
    if interp is None:
        interp = cv2.INTER_LINEAR
    assert img.ndim == 3
    h, w = img.shape[:2]
    if h < w:
        newh, neww = size, int(w * 1. * size / h + 0.5)
        h_off = 0
        w_off = (neww - size) // 2
    else:
        newh, neww = int(h * 1. * size / w + 0.5), size
        h_off = (newh - size) // 2
        w_off = 0
    ret = cv2.resize(img, (neww, newh), interpolation=interp)
    if img.ndim == 3 and ret.ndim == 2:
        ret = ret[:, :, np.newaxis]
    return ret[h_off:h_off + size, w_off:w_off + size, :]


def test_single(syn, image_target, is_preimage, seed=None):
    """synthesizer.py:178-188 (optionally seeded; the reference draws from the global numpy state)."""
    rng = np.random if seed is None else np.random.RandomState(seed)
    if is_preimage:
        input_init = rng.uniform(-1., 1., size=image_target.shape)
    else:
        input_init = rng.uniform(1. / 255, 1. - 1. / 255, size=image_target.shape)
    output, func = syn.synthesize(input_init, image_target, is_preimage=is_preimage)
    if is_preimage:
        image_syn = 1. / (1. + np.exp(-output))
    else:
        image_syn = output
    return image_syn, func
