"""The PO input pipeline (po/progressive_model.py:274-296, po/model.py:138-147): host code, SURVEY.md section 8 row f-3.

tensorpack's dataflow (ImageFromFile, imgaug.ResizeShortestEdge / RandomCrop / CenterCrop / Flip, JoinData, BatchData)
is absent from /root/reference; its behaviour is restated here with cv2 + numpy: every sample is
[z ~ U(-1, 1) of shape (size, size, 3), exemplar crop as float32 RGB in 0..255].
"""
import glob
import os

import numpy as np

from .model import RandomZData

BATCH = 1
TEST_BATCH = 1


def resize_shortest_edge(img, size, interp=None):
    """imgaug.ResizeShortestEdge(size): scale so that the shorter side becomes `size` (aspect ratio kept)."""
    import cv2
    h, w = img.shape[:2]
    scale = size * 1.0 / min(h, w)
    if h < w:
        newh, neww = size, int(scale * w + 0.5)
    else:
        newh, neww = int(scale * h + 0.5), size
    return cv2.resize(img, (neww, newh), interpolation=cv2.INTER_LINEAR if interp is None else interp)


def center_crop(img, size):
    h, w = img.shape[:2]
    h0, w0 = int((h - size) * 0.5), int((w - size) * 0.5)
    return img[h0:h0 + size, w0:w0 + size]


def random_crop(img, size, rng):
    h, w = img.shape[:2]
    h0 = 0 if h == size else rng.randint(h - size + 1)
    w0 = 0 if w == size else rng.randint(w - size + 1)
    return img[h0:h0 + size, w0:w0 + size]


class ImageFolderData(object):
    """Endless [z, image] batches from <datadir>/train (shuffled, random crop + horizontal flip) or one pass over
    <datadir>/test (centre crop): get_data(datadir, size, isTrain) of the reference."""

    def __init__(self, datadir, size, isTrain=True, batch=None, seed=None, rank=0, world=1):
        sub = "train" if isTrain else "test"
        self.files = sorted(glob.glob(os.path.join(datadir, sub, "*.jpg")))
        if not self.files:
            raise IOError("no *.jpg under %s" % os.path.join(datadir, sub))
        self.size, self.isTrain = size, isTrain
        self.batch = batch or (BATCH if isTrain else TEST_BATCH)
        # crops and flips: a stream per rank; the epoch permutation: ONE stream shared by all ranks (seed, epoch), so
        # that the strided shards below partition the same permutation and an epoch covers the dataset exactly once
        self.rng = np.random.RandomState(seed if seed is None else seed + 7919 * (rank + 1))
        self.shuffle_seed = 0 if seed is None else int(seed)
        self.epoch = 0
        self.rank, self.world = rank, world
        self.z = iter(RandomZData([size, size, 3], -1, 1, seed=None if seed is None else seed + 104729 * (rank + 1)))

    def _load(self, path):
        import cv2
        img = cv2.imread(path, cv2.IMREAD_COLOR)
        if img is None:
            raise IOError("cannot read image %s" % path)
        img = resize_shortest_edge(img[:, :, ::-1], int(self.size * 1.143))
        if self.isTrain:
            img = random_crop(img, self.size, self.rng)
            if self.rng.rand() < 0.5:
                img = img[:, ::-1]
        else:
            img = center_crop(img, self.size)
        return np.ascontiguousarray(img, dtype=np.float32)

    def _order(self):
        idx = np.arange(len(self.files))
        if self.isTrain:
            np.random.RandomState([self.shuffle_seed, self.epoch]).shuffle(idx)
            self.epoch += 1
        return idx[self.rank::self.world] if self.world > 1 and len(idx) >= self.world else idx

    def __iter__(self):
        zs, ims = [], []
        while True:
            for i in self._order():
                zs.append(np.asarray(next(self.z)[0], dtype=np.float32))
                ims.append(self._load(self.files[i]))
                if len(ims) == self.batch:
                    yield [np.stack(zs), np.stack(ims)]
                    zs, ims = [], []
            if not self.isTrain:
                if ims:
                    yield [np.stack(zs), np.stack(ims)]
                return


def get_data(datadir, size=224, isTrain=True, **kw):
    return ImageFolderData(datadir, size, isTrain, **kw)
