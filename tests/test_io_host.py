"""Host-side I/O around the hot path: image pipeline, weight files, checkpoints."""
import os

import numpy as np
import pytest
import torch

import syntex_oracle as so
from syntex_b200.gatys import driver
from syntex_b200.gatys.synthesizer import resize_crop
from syntex_b200.po import checkpoint, data
from syntex_b200.texture_utils import caffe2npz


def _write_images(folder, n, h, w, seed=0):
    import cv2
    os.makedirs(folder, exist_ok=True)
    rng = np.random.RandomState(seed)
    for i in range(n):
        img = (rng.rand(h, w, 3) * 255).astype(np.uint8)
        cv2.imwrite(os.path.join(folder, "img%02d.jpg" % i), img)


def test_resize_crop_matches_oracle_restatement():
    rng = np.random.RandomState(0)
    for shape in ((350, 500, 3), (500, 350, 3), (256, 256, 3)):
        img = rng.rand(*shape)
        a, b = resize_crop(img, 256), so.resize_crop(img, 256)
        assert a.shape == (256, 256, 3) and np.array_equal(a, b)


def test_driver_helpers(tmp_path):
    folder = str(tmp_path / "imgs")
    _write_images(folder, 3, 40, 60)
    open(os.path.join(folder, "notes.txt"), "w").write("x")
    assert driver.get_image_file_names(folder) == ["img00.jpg", "img01.jpg", "img02.jpg"]
    img = driver.imread_rgb(os.path.join(folder, "img00.jpg"))
    assert img.dtype == np.uint8 and img.shape == (40, 60, 3)
    t = driver.load_target(os.path.join(folder, "img00.jpg"), 32)
    assert t.shape == (32, 32, 3) and 0.0 <= t.min() and t.max() <= 1.0
    u8 = driver.get_plottable_data(np.array([[[0.0, 0.5, 1.2]]]))
    assert u8.dtype == np.uint8 and u8.tolist() == [[[0, 128, 255]]]
    out = str(tmp_path / "o.png")
    driver.imwrite_rgb(out, (t * 255).astype(np.uint8))
    back = driver.imread_rgb(out)
    assert np.array_equal(back, (t * 255).astype(np.uint8))           # png is lossless: RGB order round-trips
    args = driver.get_parser().parse_args(["--batch-test", "--batch-test-start", "1", "--image-size", "64"])
    assert args.get("batch_test") and args.get("batch_test_start") == 1 and args.get("batch_test_end") is None
    with pytest.raises(IOError):
        driver.imread_rgb(os.path.join(folder, "missing.jpg"))


def test_po_data_pipeline(tmp_path):
    root = str(tmp_path / "ds")
    _write_images(os.path.join(root, "train"), 5, 50, 80, seed=1)
    _write_images(os.path.join(root, "test"), 2, 80, 50, seed=2)
    it = iter(data.get_data(root, size=32, isTrain=True, seed=0))
    seen = [next(it) for _ in range(7)]           # endless: more batches than files
    for z, im in seen:
        assert z.shape == (1, 32, 32, 3) and im.shape == (1, 32, 32, 3) and im.dtype == np.float32
        assert -1.0 <= z.min() and z.max() <= 1.0 and 0.0 <= im.min() and im.max() <= 255.0 and im.max() > 1.0
    test_batches = list(data.get_data(root, size=32, isTrain=False))
    assert len(test_batches) == 2
    # centre crop of the 1.143x resize is deterministic
    again = list(data.get_data(root, size=32, isTrain=False))
    assert all(np.array_equal(a[1], b[1]) for a, b in zip(test_batches, again))
    assert data.resize_shortest_edge(np.zeros((50, 80, 3), np.uint8), 36).shape == (36, 58, 3)
    # ranks see disjoint files of the same shuffle
    a = data.ImageFolderData(root, 32, True, seed=3, rank=0, world=2)
    b = data.ImageFolderData(root, 32, True, seed=3, rank=0, world=1)
    assert len(a._order()) == 3 and len(b._order()) == 5
    with pytest.raises(IOError):
        data.get_data(str(tmp_path / "nothing"), 32)


def test_weight_file_round_trip(tmp_path, raw_weights):
    params = {"conv1_1": (np.random.RandomState(0).randn(64, 3, 3, 3), np.arange(64.0)), "odd": (np.zeros(3),)}
    out = caffe2npz.convert(params)
    assert list(out.keys()) == ["conv1_1"] and out["conv1_1"][0].shape == (3, 3, 3, 64)
    assert out["conv1_1"][0][1, 2, 0, 5] == np.float32(params["conv1_1"][0][5, 0, 1, 2])
    path = caffe2npz.save_dict(str(tmp_path / "w.npz"), raw_weights)
    back = caffe2npz.load_npz(path)
    assert set(back.keys()) == set(raw_weights.keys())
    for k in raw_weights:
        assert np.array_equal(back[k][0], raw_weights[k][0]) and np.array_equal(back[k][1], raw_weights[k][1])
    # the file Vgg19.__init__ reads (vgg19.py:101-106) and the .npy variant written by save_npy (vgg19.py:208-221)
    from syntex_b200.texture_utils import Vgg19Extractor
    ext = Vgg19Extractor(path)
    assert np.array_equal(ext.data_dict["conv4_4"][0], raw_weights["conv4_4"][0])
    npy = ext.save_npy(npy_path=str(tmp_path / "w.npy"))
    again = Vgg19Extractor(npy, is_rgb_input=False)     # already flipped once when it was saved
    assert np.array_equal(again.data_dict["conv1_1"][0], ext.data_dict["conv1_1"][0])
    with pytest.raises(ValueError):
        Vgg19Extractor(str(tmp_path / "w.txt"))


def test_po_checkpoint_round_trip(tmp_path, raw_weights):
    from syntex_b200.po.model import SynTexModelDesc
    from syntex_b200.po.progressive_model import ProgressiveSynTex
    from syntex_b200.po.trainer import SynTexTrainer
    layers = list(SynTexModelDesc.DEFAULT_COEFS.keys())
    tex = so.OracleTexture(so.load_vgg_weights(raw_weights), layers, torch.float32)
    torch.manual_seed(0)
    m1 = ProgressiveSynTex({"vgg19_npy_path": raw_weights, "image_size": 32}, texture_fn=tex)
    t1 = SynTexTrainer(None, m1, 1)
    t1.global_step = 4000
    folder = str(tmp_path / "train_log")
    p = checkpoint.save_checkpoint(folder, m1, t1)
    assert os.path.basename(p) == "model-4000" and checkpoint.latest_checkpoint(folder) == p
    checkpoint.save_checkpoint(folder, m1, global_step=200)
    assert checkpoint.latest_checkpoint(folder) == p
    torch.manual_seed(1)
    m2 = ProgressiveSynTex({"vgg19_npy_path": raw_weights, "image_size": 32}, texture_fn=tex)
    t2 = SynTexTrainer(None, m2, 1)
    assert not torch.equal(t1.buckets.flat_param, t2.buckets.flat_param)
    assert checkpoint.restore_checkpoint(folder, m2, t2) == 4000
    assert torch.equal(t1.buckets.flat_param, t2.buckets.flat_param)      # restored in place, aliasing kept
    assert t2.global_step == 4000
