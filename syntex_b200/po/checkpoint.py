"""PO checkpoints: tensorpack's ModelSaver writes `model-<global_step>` files into the
log directory and `SaverRestore` / `--test-only` load them (po/progressive_model.py:354-356, po/improved_model.py:453-484).
Same naming here, torch serialisation inside (TensorFlow checkpoint files are not readable without TensorFlow)."""
import glob
import os
import re

import torch


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
    return max(files, key=lambda p: int(p.rsplit("-", 1)[1])) if files else None


def restore_checkpoint(path, model, trainer=None):
    """path: a `model-<step>` file or a folder holding some. Returns the global step."""
    if os.path.isdir(path):
        found = latest_checkpoint(path)
        if found is None:
            raise IOError("no model-<step> checkpoint in %s" % path)
        path = found
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
