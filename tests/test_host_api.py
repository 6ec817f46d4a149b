"""Host-side mirror of the reference interface: signatures, defaults, loader behaviour and error convention
(no device work)."""
import inspect
import os

import numpy as np
import pytest

from syntex_b200 import parallel
from syntex_b200.aparse import ArgParser
from syntex_b200.gatys import Synthesizer, get_bounds, resize_crop
from syntex_b200.po import SynTexModelDesc
from syntex_b200.synthetic import save_vgg_npz
from syntex_b200.texture_utils import VGG_MEAN, VGG_MEAN_RGB, Vgg19, Vgg19Extractor, build_gram, build_texture_loss


def test_signatures_match_the_reference():
    # texture_utils/vgg19_extractor.py:5-6
    p = inspect.signature(Vgg19Extractor.__init__).parameters
    assert list(p)[1:] == ["vgg19_path", "trainable", "padding", "is_rgb_input", "use_avg_pool", "topmost", "name"]
    assert p["topmost"].default == "pool4" and p["padding"].default == "SAME" and p["use_avg_pool"].default is True
    assert inspect.signature(Vgg19.__init__).parameters["topmost"].default == "fc8"          # vgg19.py:97-98
    assert list(inspect.signature(Vgg19.build).parameters)[1:] == ["input_image", "topmost", "padding", "name"]
    assert list(inspect.signature(build_gram).parameters) == ["feat", "mask", "eps", "name"]   # utils.py:8
    p = inspect.signature(build_texture_loss).parameters                                      # utils.py:23-24
    assert list(p) == ["a", "b", "coefs", "is_gram_a", "is_gram_b", "calc_grad", "name"]
    assert p["is_gram_b"].default is True and p["is_gram_a"].default is False and p["calc_grad"].default is False
    # gatys/synthesizer.py:109,119,127
    assert list(inspect.signature(Synthesizer.compute_gram).parameters)[1:] == ["image", "mask", "update", "is_preimage"]
    assert list(inspect.signature(Synthesizer.step).parameters)[1:] == ["input", "mask", "target_image", "is_preimage"]
    assert list(inspect.signature(Synthesizer.synthesize).parameters)[1:] == [
        "input_init", "image_target", "mask_input", "mask_target", "bounds", "is_preimage"]
    # po/model.py:51,60
    assert list(inspect.signature(SynTexModelDesc._build_loss).parameters)[1:] == [
        "image_output", "gram_target", "calc_grad", "name"]
    assert list(inspect.signature(SynTexModelDesc._build_extractor).parameters)[1:] == ["image_input", "calc_gram", "name"]


REF = "/root/reference"


def _ref_signature(relpath, qualname):
    """(argument names, {name: default literal}) of a function in the reference's SOURCE, by `ast` (the reference
    cannot be imported here: it needs TensorFlow 1.12 / tensorpack)."""
    import ast
    tree = ast.parse(open(os.path.join(REF, relpath)).read())
    node = tree
    for part in qualname.split("."):
        node = next(n for n in ast.walk(node) if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name == part)
    names = [a.arg for a in node.args.args]
    defaults = {}
    for a, d in zip(reversed(node.args.args), reversed(node.args.defaults)):
        try:
            defaults[a.arg] = ast.literal_eval(d)
        except ValueError:
            pass
    return names, defaults


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present on this machine")
@pytest.mark.parametrize("relpath,qualname,ours", [
    ("texture_utils/vgg19_extractor.py", "Vgg19Extractor.__init__", Vgg19Extractor.__init__),
    ("texture_utils/vgg19.py", "Vgg19.build", Vgg19.build),
    ("texture_utils/utils.py", "build_gram", build_gram),
    ("texture_utils/utils.py", "build_texture_loss", build_texture_loss),
    ("gatys/synthesizer.py", "Synthesizer.compute_gram", Synthesizer.compute_gram),
    ("gatys/synthesizer.py", "Synthesizer.step", Synthesizer.step),
    ("gatys/synthesizer.py", "Synthesizer.synthesize", Synthesizer.synthesize),
    ("gatys/synthesizer.py", "get_bounds", get_bounds),
    ("gatys/synthesizer.py", "resize_crop", resize_crop),
    ("po/model.py", "SynTexModelDesc._build_loss", SynTexModelDesc._build_loss),
    ("po/model.py", "SynTexModelDesc._build_extractor", SynTexModelDesc._build_extractor),
])
def test_signatures_parsed_from_the_reference_sources(relpath, qualname, ours):
    """Same argument names in the same order, same literal defaults - read from the reference's own files."""
    names, defaults = _ref_signature(relpath, qualname)
    p = inspect.signature(ours).parameters
    assert list(p)[:len(names)] == names, (qualname, list(p), names)
    for k, v in defaults.items():
        if k == "name":          # TF name scopes: accepted and ignored here
            continue
        assert p[k].default == v, (qualname, k, p[k].default, v)


def test_constants():
    assert VGG_MEAN == [103.939, 116.779, 123.68] and VGG_MEAN_RGB == [123.68, 116.779, 103.939]
    assert list(Synthesizer.DEFAULT_COEFS.items()) == [("conv1_1", 1e5), ("pool1", 1e5), ("pool2", 1e5),
                                                        ("pool3", 1e5), ("pool4", 1e5)]
    assert Synthesizer.DEFAULT_COEFS == SynTexModelDesc.DEFAULT_COEFS


def test_parser_defaults():
    args = Synthesizer.get_parser().parse_args([])
    assert args["image_size"] == 256 and args["maxiter"] == 1000 and args["maxcor"] == 20
    assert args["gram_only"] is False and args["verbose"] is False and args["vgg19_npy_path"] is None
    ps = ArgParser(None, name="x")
    ps.add("--a", type=int, default=3)
    ps.add_flag("--b")
    assert ps.parse_args(["--b"]) == {"a": 3, "b": True}
    syn = Synthesizer({"maxiter": 7, "maxcor": 5, "image_size": 64})
    assert syn.options == {"maxiter": 7, "maxcor": 5, "disp": False, "ftol": 0., "gtol": 0.}


def test_loader_npz_roundtrip_flip_and_errors(tmp_path, raw_weights):
    path = os.path.join(tmp_path, "vgg19_normalised.npz")
    save_vgg_npz(path, raw_weights)
    v = Vgg19Extractor(path)
    assert v.topmost == "pool4" and len(v.data_dict) == 12
    np.testing.assert_array_equal(v.data_dict["conv1_1"][0], raw_weights["conv1_1"][0][:, :, ::-1, :])
    np.testing.assert_array_equal(v.data_dict["conv4_4"][0], raw_weights["conv4_4"][0])
    bgr = Vgg19Extractor(path, is_rgb_input=False)
    np.testing.assert_array_equal(bgr.data_dict["conv1_1"][0], raw_weights["conv1_1"][0])
    short = Vgg19(path, topmost="pool2")
    assert list(short.data_dict) == ["conv1_1", "conv1_2", "conv2_1", "conv2_2"]
    npy = v.save_npy(npy_path=os.path.join(tmp_path, "w.npy"))
    back = Vgg19Extractor(npy, is_rgb_input=False)          # saved after the flip
    np.testing.assert_array_equal(back.data_dict["conv1_1"][0], v.data_dict["conv1_1"][0])
    assert v.get_var_count() == sum(w.size + b.size for w, b in raw_weights.values())
    with pytest.raises(ValueError):
        Vgg19Extractor(os.path.join(tmp_path, "weights.txt"))          # vgg19.py:106
    with pytest.raises(AssertionError):
        Vgg19Extractor(path, padding="FULL")                           # vgg19.py:110
    for kw in ({"padding": "VALID"}, {"use_avg_pool": False}, {"trainable": True}, {"topmost": "fc8"}):
        with pytest.raises(NotImplementedError):
            Vgg19Extractor(path, **kw)


def test_bounds_and_resize_crop():
    b = get_bounds([1, 4, 4, 3])
    assert b.shape == (48, 2) and np.all(b[:, 0] == 0) and np.all(b[:, 1] == 1)
    img = np.random.RandomState(0).uniform(size=(300, 420, 3))
    assert resize_crop(img, 256).shape == (256, 256, 3)
    assert resize_crop(img.transpose(1, 0, 2), 128).shape == (128, 128, 3)


def test_no_cuda_no_fallback(raw_weights):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from syntex_b200._lib import SyntexLibraryError
    v = Vgg19Extractor(raw_weights)
    with pytest.raises(SyntexLibraryError):
        v.build(np.zeros((1, 32, 32, 3), np.float32))
    with pytest.raises(SyntexLibraryError):
        build_gram(np.zeros((1, 4, 4, 64), np.float32))


def test_job_sharding():
    for n, w in ((64, 8), (47, 4), (3, 8), (0, 2)):
        parts = [parallel.shard_jobs(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert parallel.batches(list(range(10)), 4) == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
    with pytest.raises(ValueError):
        parallel.shard_jobs(4, 2, 2)

    def make_job(j):
        return np.full((2, 2, 3), j, float), np.full((2, 2, 3), -j, float)

    def fake(x0s, exs):
        return x0s * 2, x0s[:, 0, 0, 0] + 0.5, np.full(len(x0s), 7), np.full(len(x0s), 8)

    images, fun, nit, nfev = parallel.run_sharded_gatys(make_job, 5, fake, batch=2)
    np.testing.assert_allclose(fun, np.arange(5) + 0.5)
    assert list(nit) == [7] * 5 and list(nfev) == [8] * 5 and sorted(images) == list(range(5))
