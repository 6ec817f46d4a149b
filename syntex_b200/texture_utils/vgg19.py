"""Host mirror of the reference's texture_utils/vgg19.py. The reference builds TF graph ops; here `build` runs
the CUDA plan (conv1_1 direct kernel, tcgen05 implicit-GEMM convs, 2x2 average pools) eagerly on torch CUDA
tensors and returns the same OrderedDict layer -> [B,H,W,C]."""
import ctypes
import inspect
import os
from collections import OrderedDict

import numpy as np

from .. import _lib

VGG_MEAN = [103.939, 116.779, 123.68]  # BGR [0, 255]            (vgg19.py:9)
VGG_MEAN_RGB = list(reversed(VGG_MEAN))  # (vgg19.py:10)
PARAM_SHAPE = OrderedDict([  # vgg19.py:11-30; the CUDA path covers the network up to pool4
    ("conv1_1", [(3, 3, 3, 64), (64,)]),
    ("conv1_2", [(3, 3, 64, 64), (64,)]),
    ("pool1", []),
    ("conv2_1", [(3, 3, 64, 128), (128,)]),
    ("conv2_2", [(3, 3, 128, 128), (128,)]),
    ("pool2", []),
    ("conv3_1", [(3, 3, 128, 256), (256,)]),
    ("conv3_2", [(3, 3, 256, 256), (256,)]),
    ("conv3_3", [(3, 3, 256, 256), (256,)]),
    ("conv3_4", [(3, 3, 256, 256), (256,)]),
    ("pool3", []),
    ("conv4_1", [(3, 3, 256, 512), (512,)]),
    ("conv4_2", [(3, 3, 512, 512), (512,)]),
    ("conv4_3", [(3, 3, 512, 512), (512,)]),
    ("conv4_4", [(3, 3, 512, 512), (512,)]),
    ("pool4", []),
])
LAYER_NAMES = list(PARAM_SHAPE.keys())
LAYER_INDEX = {k: i for i, k in enumerate(LAYER_NAMES)}


def get_default_vgg19_path():
    """vgg19.py:79-88: vgg19_normalised.npz next to this module, or None."""
    path = inspect.getfile(get_default_vgg19_path)
    path = os.path.abspath(os.path.join(path, os.pardir))
    path = os.path.join(path, "vgg19_normalised.npz")
    if os.path.exists(path):
        print("VGG19 path is found at", path)
        return path
    print("VGG19 path is not found")
    return None


class _Plan(object):
    """Owner of one stx_plan handle."""

    FAST_BF16 = 1   # STX_PLAN_FAST_BF16
    EXPORT_LAYER_GRADS = 2   # STX_PLAN_EXPORT_LAYER_GRADS
    FULL_FEATURES = 4   # STX_PLAN_FULL_FEATURES
    F16F8 = 8   # STX_PLAN_F16F8
    F16F8_LEAN = 16   # STX_PLAN_F16F8_LEAN

    def __init__(self, vgg_handle, batch, height, width, topmost_idx, loss_layers, coefs, flags=0):
        self.handle = ctypes.c_void_p()
        n = len(loss_layers)
        ll = (ctypes.c_int * max(n, 1))(*loss_layers)
        cf = (ctypes.c_float * max(n, 1))(*coefs)
        _lib.check(_lib.lib().stx_plan_create(ctypes.byref(self.handle), vgg_handle, batch, height, width,
                                              topmost_idx, n, ll, cf, int(flags)))
        self.batch, self.height, self.width = batch, height, width
        self.topmost_idx = topmost_idx
        self.loss_layers = list(loss_layers)
        self.coefs = list(coefs)

    def layer_shape(self, layer_idx):
        h, w, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(_lib.lib().stx_plan_layer_shape(self.handle, layer_idx, ctypes.byref(h), ctypes.byref(w),
                                                   ctypes.byref(c)))
        return h.value, w.value, c.value

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().stx_plan_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:
            pass


class Vgg19(object):
    """vgg19.py:96-169. Same constructor signature; weights become packed bf16 device constants
    (non-trainable, as in the reference's default get_var path, vgg19.py:192-206)."""

    def __init__(self, vgg19_path=None, trainable=False, padding="SAME",
                 is_rgb_input=True, use_avg_pool=True, topmost="fc8", name="vgg19"):
        if vgg19_path is None:
            vgg19_path = get_default_vgg19_path()
        if isinstance(vgg19_path, dict):          # extension: weights passed directly ({layer: [w, b]})
            data_dict = vgg19_path
        elif vgg19_path is None:
            raise ValueError("vgg19_npy_path should be a 'npy' or 'npz' file (no default weights found)")
        elif vgg19_path.endswith(".npy"):
            data_dict = np.load(vgg19_path, encoding="latin1", allow_pickle=True).item()
        elif vgg19_path.endswith(".npz"):
            data_dict = np.load(vgg19_path, encoding="latin1", allow_pickle=True)
        else:
            raise ValueError("vgg19_npy_path should be a 'npy' or 'npz' file")
        assert padding in ["SAME", "VALID"]
        if trainable:
            raise NotImplementedError("syntex_b200: VGG weights are constants on this path (the reference never "
                                      "trains them: every caller uses trainable=False)")
        if padding != "SAME":
            raise NotImplementedError("syntex_b200: only SAME padding is built (no reference caller uses VALID)")
        if not use_avg_pool:
            raise NotImplementedError("syntex_b200: only average pooling is built (no reference caller uses max)")
        if topmost not in LAYER_INDEX:
            raise NotImplementedError("syntex_b200 covers VGG19 up to pool4; topmost=%r is outside the hot path"
                                      % (topmost,))
        self.trainable = trainable
        self.padding = padding
        self.is_rgb_input = is_rgb_input
        self.use_avg_pool = use_avg_pool
        self.topmost = topmost
        self.name = name

        # load parameters (vgg19.py:117-127)
        self.data_dict = dict()
        for layer in PARAM_SHAPE:
            if layer in data_dict:
                self.data_dict[layer] = [np.asarray(data_dict[layer][0], dtype=np.float32),
                                         np.asarray(data_dict[layer][1], dtype=np.float32)]
            if layer == topmost:
                break
        if is_rgb_input:
            w_conv1_1 = self.data_dict["conv1_1"][0]
            assert w_conv1_1.shape == PARAM_SHAPE["conv1_1"][0]
            self.data_dict["conv1_1"][0] = np.ascontiguousarray(w_conv1_1[:, :, ::-1, :])
        if hasattr(data_dict, "close"):
            data_dict.close()
        self._vgg = None
        self._plans = {}
        if not isinstance(vgg19_path, dict):
            print("Loaded vgg19 from", vgg19_path)

    # ------------------------------------------------------------------ native handles
    def _vgg_handle(self):
        if self._vgg is None:
            _lib.require_cuda()
            ws, bs = [], []
            for layer in PARAM_SHAPE:
                if not PARAM_SHAPE[layer]:
                    if layer == self.topmost:
                        break
                    continue
                if layer not in self.data_dict:
                    # the reference would fall back to truncated_normal(1e-3) here (vgg19.py:197-198); a missing
                    # layer in a weight file is a user error worth surfacing
                    raise ValueError("weights for layer %s missing from the VGG19 file" % layer)
                w, b = self.data_dict[layer]
                if tuple(w.shape) != PARAM_SHAPE[layer][0] or tuple(b.shape) != PARAM_SHAPE[layer][1]:
                    raise ValueError("bad weight shape for %s: %s %s" % (layer, w.shape, b.shape))
                ws.append(np.ascontiguousarray(w, dtype=np.float32))
                bs.append(np.ascontiguousarray(b, dtype=np.float32))
                if layer == self.topmost:
                    break
            n = len(ws)
            wp = (ctypes.c_void_p * n)(*[a.ctypes.data for a in ws])
            bp = (ctypes.c_void_p * n)(*[a.ctypes.data for a in bs])
            mean = VGG_MEAN_RGB if self.is_rgb_input else VGG_MEAN
            m3 = (ctypes.c_float * 3)(*mean)
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().stx_vgg_create(ctypes.byref(h), wp, bp, n, m3, _lib.stream_ptr()))
            self._vgg = h
        return self._vgg

    def get_plan(self, batch, height, width, topmost=None, coefs=None, fast_bf16=False, export_layer_grads=False,
                 slot=0, full_features=None, f16f8=False):
        """Cached stx_plan for this shape. coefs: OrderedDict layer name -> weight (Gram layers), or None.
        fast_bf16: single-MMA bf16 forward instead of the accurate bf16x3 forward (see include/syntex_b200.h).
        f16f8: the f16 + f8 forward (STX_PLAN_F16F8): same accuracy class as bf16x3 at two thirds of its tensor time;
        "lean" adds STX_PLAN_F16F8_LEAN (conv4_2..conv4_4 without the correction product).
        slot: distinct slots give distinct plans of the same shape (each keeps the activations of its own last
        forward - the PO training graph evaluates VGG once per stage and back-propagates through all of them)."""
        topmost = self.topmost if topmost is None else topmost
        if LAYER_INDEX[topmost] > LAYER_INDEX[self.topmost]:
            raise ValueError("topmost %s is above the layers loaded (%s)" % (topmost, self.topmost))
        coefs = OrderedDict() if coefs is None else coefs
        if full_features is None:       # plans without loss layers exist to export features: keep every low half
            full_features = len(coefs) == 0
        key = (batch, height, width, topmost, tuple((k, float(v)) for k, v in coefs.items()), bool(fast_bf16),
               bool(export_layer_grads), int(slot), bool(full_features), str(f16f8))
        plan = self._plans.get(key)
        if plan is None:
            plan = _Plan(self._vgg_handle(), batch, height, width, LAYER_INDEX[topmost],
                         [LAYER_INDEX[k] for k in coefs], [float(v) for v in coefs.values()],
                         (_Plan.FAST_BF16 if fast_bf16 else 0) |
                         (_Plan.EXPORT_LAYER_GRADS if export_layer_grads else 0) |
                         (_Plan.FULL_FEATURES if full_features else 0) |
                         (_Plan.F16F8 if (f16f8 and not fast_bf16) else 0) |
                         (_Plan.F16F8_LEAN if (f16f8 == "lean" and not fast_bf16) else 0))
            plan.coef_names = list(coefs.keys())
            self._plans[key] = plan
        return plan

    # ------------------------------------------------------------------ reference API
    def build(self, input_image, topmost=None, padding=None, name="vgg19"):
        """vgg19.py:132-169. input_image: torch CUDA tensor [B,H,W,3] (any float dtype), values in [0,1].
        Returns OrderedDict layer -> torch.float32 CUDA tensor [B,h,w,c] for every layer up to `topmost`.
        (Features are computed and stored in bf16 on the device; the fp32 tensors returned are exact copies.)"""
        import torch
        if padding is not None and padding != "SAME":
            raise NotImplementedError("syntex_b200: only SAME padding is built")
        if topmost is None:
            topmost = self.topmost
        x = input_image
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.asarray(x))
        if x.dim() != 4 or x.shape[-1] != 3:
            raise ValueError("input_image must be [batch, height, width, 3], got %s" % (tuple(x.shape),))
        _lib.require_cuda()
        x = x.to(device="cuda", dtype=torch.float32).contiguous()
        B, H, W, _ = x.shape
        plan = self.get_plan(B, H, W, topmost)
        L = _lib.lib()
        _lib.check(L.stx_plan_forward(plan.handle, x.data_ptr(), 0, 0, _lib.stream_ptr()))
        layers = OrderedDict()
        for layer in PARAM_SHAPE:
            idx = LAYER_INDEX[layer]
            h, w, c = plan.layer_shape(idx)
            out = torch.empty((B, h, w, c), dtype=torch.float32, device="cuda")
            _lib.check(L.stx_plan_copy_feature_f32(plan.handle, idx, out.data_ptr(), _lib.stream_ptr()))
            layers[layer] = out
            if layer == topmost:
                break
        print("Build vgg19 model:", name)
        return layers

    def get_var_count(self):
        count = 0
        for layer, (w, b) in self.data_dict.items():
            count += int(np.prod(w.shape)) + int(np.prod(b.shape))
        return count

    def save_npy(self, sess=None, npy_path="vgg19-save.npy"):
        """vgg19.py:208-221: {layer: {0: w, 1: b}} (weights as held in memory, i.e. after the RGB flip)."""
        data = {k: {0: v[0], 1: v[1]} for k, v in self.data_dict.items()}
        np.save(npy_path, data)
        print("Save param to", npy_path)
        return npy_path

    def __del__(self):
        try:
            self._plans.clear()
            if self._vgg is not None:
                _lib.lib().stx_vgg_destroy(self._vgg)
                self._vgg = None
        except Exception:
            pass
