"""PO checkpoints: tensorpack's ModelSaver writes `model-<global_step>` files into the
log directory and `SaverRestore` / `--test-only` load them (po/progressive_model.py:354-356, po/improved_model.py:453-484).
Same file naming here. Two formats:
  * torch serialisation (save_checkpoint / restore_checkpoint): everything, including the optimiser state;
  * TensorFlow's own checkpoint format (export_tf_checkpoint / restore_tf_checkpoint, through po/tf_bundle.py, no
    TensorFlow needed) under the reference's variable names, for the generators of po/progressive_model.py and po/single_model.py: a
    checkpoint trained with the reference loads here, and one written here loads with the reference's SaverRestore.
    restore_checkpoint() recognises such a file by its `.index` companion."""
import glob
import os
import re

import numpy as np
import torch

from . import tf_bundle


def save_checkpoint(folder, model, trainer=None, global_step=None):
    os.makedirs(folder, exist_ok=True)
    step = int(global_step if global_step is not None else (trainer.global_step if trainer is not None else 0))
    state = {"global_step": step, "model": {k: v.detach().cpu() for k, v in model.state_dict().items()}}
    if trainer is not None:
        state["optimizer"] = trainer.opt.state_dict()       # plain tensors and floats only (see restore)
        state["learning_rate"] = float(trainer.opt.param_groups[0]["lr"])
    path = os.path.join(folder, "model-%d" % step)
    torch.save(state, path)
    with open(os.path.join(folder, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "model-%d"\n' % step)
    return path


def latest_checkpoint(folder):
    files = [p for p in glob.glob(os.path.join(folder, "model-*")) if re.search(r"model-\d+$", p)]
    files += [p[:-len(".index")] for p in glob.glob(os.path.join(folder, "model-*.index"))
              if re.search(r"model-\d+\.index$", p)]
    return max(files, key=lambda p: int(p.rsplit("-", 1)[1])) if files else None


# ---------------------------------------------------------------------------------------------- TensorFlow format
# state_dict key of the progressive / single generator -> variable name in the reference's graph
# (po/progressive_model.py:145-206, po/single_model.py:142-190: scopes syn / stage<s> / <layer> / ...; tensorpack's Conv2D names its variables W
# and b, its InstanceNorm gamma and beta; INReLU's norm lives under the convolution's scope as `inorm`).
_L = r"(conv\d_\d|pool\d)"
_TF_RULES = [
    (r"^syn\.(\d+)\.grad_conv(\d)\.%s\.conv\.weight$" % _L, "syn/stage{0}/{2}/grad_conv{1}/W"),
    (r"^syn\.(\d+)\.grad_conv(\d)\.%s\.norm\.weight$" % _L, "syn/stage{0}/{2}/grad_conv{1}/inorm/gamma"),
    (r"^syn\.(\d+)\.grad_conv(\d)\.%s\.norm\.bias$" % _L, "syn/stage{0}/{2}/grad_conv{1}/inorm/beta"),
    (r"^syn\.(\d+)\.up\.%s\.conv\.conv\.weight$" % _L, "syn/stage{0}/{1}/up/W"),
    (r"^syn\.(\d+)\.post_inorm\.%s\.weight$" % _L, "syn/stage{0}/{1}/post_inorm/gamma"),
    (r"^syn\.(\d+)\.post_inorm\.%s\.bias$" % _L, "syn/stage{0}/{1}/post_inorm/beta"),
    (r"^syn\.(\d+)\.pre_inorm\.%s\.weight$" % _L, "syn/stage{0}/{1}/inorm/gamma"),
    (r"^syn\.(\d+)\.pre_inorm\.%s\.bias$" % _L, "syn/stage{0}/{1}/inorm/beta"),
    (r"^syn\.(\d+)\.res\.%s\.(\d+)\.conv(\d)\.conv\.weight$" % _L, "syn/stage{0}/{1}/res{2}/conv{3}/W"),
    (r"^syn\.(\d+)\.res\.%s\.(\d+)\.conv(\d)\.norm\.weight$" % _L, "syn/stage{0}/{1}/res{2}/conv{3}/inorm/gamma"),
    (r"^syn\.(\d+)\.res\.%s\.(\d+)\.conv(\d)\.norm\.bias$" % _L, "syn/stage{0}/{1}/res{2}/conv{3}/inorm/beta"),
    (r"^syn\.(\d+)\.res\.%s\.(\d+)\.inorm\.weight$" % _L, "syn/stage{0}/{1}/res{2}/inorm/gamma"),
    (r"^syn\.(\d+)\.res\.%s\.(\d+)\.inorm\.bias$" % _L, "syn/stage{0}/{1}/res{2}/inorm/beta"),
    (r"^syn\.(\d+)\.actlast_inorm\.weight$", "syn/stage{0}/inorm/gamma"),
    (r"^syn\.(\d+)\.actlast_inorm\.bias$", "syn/stage{0}/inorm/beta"),
    (r"^syn\.(\d+)\.convlast\.conv\.weight$", "syn/stage{0}/convlast/W"),
    (r"^syn\.(\d+)\.convlast\.conv\.bias$", "syn/stage{0}/convlast/b"),
]
_TF_RULES = [(re.compile(a), b) for a, b in _TF_RULES]

# po/improved_model.py:24-56, :131-330 and po/adaptive_model.py (same scopes): the activation's norm is `inorm`
# (InstanceNorm) or `bnorm` (BatchNorm: gamma, beta, mean/EMA, variance/EMA) by --norm-type, the up-sampling
# convolution lives one scope deeper (`up/conv/W`), the norm after the merge keeps the name `post_inorm` whatever its
# type, and the pre-activation variant's extra norms sit directly under the layer / the stage scope.
_NORM_PARAM = {"weight": "gamma", "bias": "beta", "running_mean": "mean/EMA", "running_var": "variance/EMA"}
_TF_RULES_V2 = [
    (r"^syn\.(\d+)\.grad_conv(\d)\.%s\.conv\.weight$" % _L, "syn/stage{0}/{2}/grad_conv{1}/W", None),
    (r"^syn\.(\d+)\.grad_conv(\d)\.%s\.norm\.(\w+)$" % _L, "syn/stage{0}/{2}/grad_conv{1}/{N}/{P}", 3),
    (r"^syn\.(\d+)\.up\.%s\.conv\.conv\.weight$" % _L, "syn/stage{0}/{1}/up/conv/W", None),
    (r"^syn\.(\d+)\.post_norm\.%s\.(\w+)$" % _L, "syn/stage{0}/{1}/post_inorm/{P}", 2),
    (r"^syn\.(\d+)\.pre_norm\.%s\.(\w+)$" % _L, "syn/stage{0}/{1}/{N}/{P}", 2),
    (r"^syn\.(\d+)\.res\.%s\.(\d+)\.conv(\d)\.conv\.weight$" % _L, "syn/stage{0}/{1}/res{2}/conv{3}/W", None),
    (r"^syn\.(\d+)\.res\.%s\.(\d+)\.conv(\d)\.norm\.(\w+)$" % _L, "syn/stage{0}/{1}/res{2}/conv{3}/{N}/{P}", 4),
    (r"^syn\.(\d+)\.res\.%s\.(\d+)\.norm\.(\w+)$" % _L, "syn/stage{0}/{1}/res{2}/{N}/{P}", 3),
    (r"^syn\.(\d+)\.last_norm\.(\w+)$", "syn/stage{0}/{N}/{P}", 1),
    (r"^syn\.(\d+)\.convlast\.conv\.weight$", "syn/stage{0}/convlast/W", None),
    (r"^syn\.(\d+)\.convlast\.conv\.bias$", "syn/stage{0}/convlast/b", None),
]
_TF_RULES_V2 = [(re.compile(a), b, c) for a, b, c in _TF_RULES_V2]


def tf_variable_name(key, norm_type=None):
    """Reference variable name of a state_dict key. norm_type None: the generators of po/progressive_model.py and
    po/single_model.py; "instance" / "batch" / "none": those of po/improved_model.py and po/adaptive_model.py.
    KeyError for a key that has no counterpart; None for one that TensorFlow does not store (num_batches_tracked)."""
    if norm_type is None:
        for rx, fmt in _TF_RULES:
            m = rx.match(key)
            if m:
                return fmt.format(*m.groups())
    else:
        if key.endswith("num_batches_tracked"):
            return None
        scope = {"instance": "inorm", "batch": "bnorm"}.get(norm_type, "nonorm")
        for rx, fmt, pidx in _TF_RULES_V2:
            m = rx.match(key)
            if m:
                g = m.groups()
                if pidx is None:
                    return fmt.format(*g)
                if g[pidx] not in _NORM_PARAM:
                    break
                return fmt.replace("{N}", scope).replace("{P}", _NORM_PARAM[g[pidx]]).format(*g)
    raise KeyError("no TensorFlow variable name known for parameter %r" % key)


def _names_for(model):
    """{reference variable name: state_dict key} of a generator."""
    norm_type = getattr(model, "_norm_type", None) if hasattr(model, "_act_type") else None
    if not getattr(model, "_nn_upsample", True):
        raise NotImplementedError("TensorFlow checkpoints: the deconvolution up-sampling variant is not mapped")
    out = {}
    for k in model.state_dict():
        n = tf_variable_name(k, norm_type)
        if n is not None:
            out[n] = k
    return out


def _to_tf(key, t):
    a = t.detach().cpu().float().numpy()
    return np.ascontiguousarray(a.transpose(2, 3, 1, 0)) if a.ndim == 4 else a      # OIHW -> HWIO


def _from_tf(a, like):
    a = np.asarray(a, dtype=np.float32)
    if like.dim() == 4:
        a = a.transpose(3, 2, 0, 1)                                                   # HWIO -> OIHW
    if tuple(a.shape) != tuple(like.shape):
        raise ValueError("shape %s in the checkpoint, %s in the model" % (a.shape, tuple(like.shape)))
    return torch.from_numpy(np.ascontiguousarray(a))


def export_tf_checkpoint(folder, model, global_step=0):
    """Write `folder/model-<step>.index` + `.data-00000-of-00001` in TensorFlow's checkpoint format under the
    reference's variable names (model parameters and `global_step`; no optimiser slots), plus the `checkpoint` state
    file tf.train.latest_checkpoint reads. Returns the checkpoint prefix."""
    names = _names_for(model)
    os.makedirs(folder, exist_ok=True)
    step = int(global_step)
    sd = model.state_dict()
    tensors = {n: _to_tf(k, sd[k]) for n, k in names.items()}
    tensors["global_step"] = np.array(step, dtype=np.int64)
    prefix = os.path.join(folder, "model-%d" % step)
    tf_bundle.write_bundle(prefix, tensors)
    with open(os.path.join(folder, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "model-%d"\nall_model_checkpoint_paths: "model-%d"\n' % (step, step))
    return prefix


def restore_tf_checkpoint(prefix, model, strict=True):
    """Load a TensorFlow checkpoint written by the reference (tensorpack ModelSaver) or by export_tf_checkpoint into
    the progressive generator, in place. Optimiser slots (`.../Adam`, `beta1_power`, ...) and summaries' moving
    averages in the file are ignored. Returns the global step (0 if the file has none)."""
    own = model.state_dict()
    wanted = _names_for(model)
    have = tf_bundle.list_variables(prefix)
    missing = sorted(n for n in wanted if n not in have)
    if missing and strict:
        raise KeyError("checkpoint %s lacks %d variables of the model, e.g. %s" % (prefix, len(missing), missing[:3]))
    data = tf_bundle.read_bundle(prefix, names=set(wanted) | {"global_step"})
    with torch.no_grad():
        for name, key in wanted.items():
            if name in data:
                own[key].copy_(_from_tf(data[name], own[key]))
    return int(data["global_step"]) if "global_step" in data else 0


def restore_checkpoint(path, model, trainer=None):
    """path: a `model-<step>` file or a folder holding some. Returns the global step."""
    if os.path.isdir(path):
        found = latest_checkpoint(path)
        if found is None:
            raise IOError("no model-<step> checkpoint in %s" % path)
        path = found
    if tf_bundle.is_bundle(path) and not os.path.isfile(path):      # a TensorFlow checkpoint prefix
        step = restore_tf_checkpoint(path, model)
        if trainer is not None:
            trainer.global_step = step
        return step
    # tensors, floats and strings only: nothing in a checkpoint needs unpickling arbitrary objects
    state = torch.load(path, map_location="cpu", weights_only=True)
    with torch.no_grad():
        own = model.state_dict()
        for k, v in state["model"].items():
            own[k].copy_(v)          # in place: the trainer's flat parameter buffer keeps aliasing the parameters
    if trainer is not None and "optimizer" in state:
        trainer.opt.load_state_dict(state["optimizer"])
        trainer.global_step = state["global_step"]
    return state["global_step"]
