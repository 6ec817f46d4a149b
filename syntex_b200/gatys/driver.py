"""The driver of gatys/synthesizer.py:190-268 (`test()`): single-image and batch-test synthesis from image files.

Host code around the hot path. imageio and the un-vendored helper modules of the
reference (file_utils, syntex.visualization) are not available here: cv2 reads and writes the images, and the two
helpers are restated from their call sites. Batch test: image j of the (sorted) folder listing, restricted to
[--batch-test-start, --batch-test-end) - the reference's own way of sharding a folder over processes
(synthesizer.py:203-204, :237); `syntex_b200.parallel` does the same split automatically under torch.distributed.
"""
import os
import time

import numpy as np

from .synthesizer import Synthesizer, resize_crop, test_single

IMAGE_EXTENSIONS = (".jpg", ".jpeg", ".png", ".bmp")


def get_image_file_names(folder):
    """Sorted image file names of a folder (file_utils.get_image_file_names at synthesizer.py:230)."""
    return sorted(f for f in os.listdir(folder) if f.lower().endswith(IMAGE_EXTENSIONS))


def imread_rgb(path):
    """imageio.imread(path, pilmode="RGB") (synthesizer.py:239): uint8 [H,W,3], RGB."""
    import cv2
    img = cv2.imread(path, cv2.IMREAD_COLOR)
    if img is None:
        raise IOError("cannot read image %s" % path)
    return img[:, :, ::-1]


def get_plottable_data(image, scale=255.):
    """float image in [0,1] -> uint8 (syntex.visualization.get_plottable_data at synthesizer.py:257)."""
    return np.clip(np.asarray(image) * scale + 0.5, 0, 255).astype(np.uint8)


def imwrite_rgb(path, image_u8):
    import cv2
    if not cv2.imwrite(path, np.ascontiguousarray(image_u8[:, :, ::-1])):
        raise IOError("cannot write image %s" % path)


def get_parser():
    ps = Synthesizer.get_parser()
    ps.add("--image-path", type=str, default="../images/flower_beds_256/FlowerBeds0008_256.jpg")
    ps.add("--image-folder", type=str, default="../images/honeycombed/train")
    ps.add("--save-folder", type=str, default="train_log")
    ps.add_flag("--batch-test")
    ps.add("--batch-test-start", type=int, default=0)
    ps.add("--batch-test-end", type=int, default=None)
    ps.add_flag("--preimage")
    return ps


def load_target(path, image_size):
    return resize_crop(imread_rgb(path) / 255., image_size)


def run(args, log=print):
    """synthesizer.py:205-265. Returns a list of (name, final loss, seconds)."""
    is_preimage = args.get("preimage", False)
    image_size = args.get("image_size")
    syn = Synthesizer(args, None)
    syn.build()
    results = []
    if args.get("batch_test"):
        image_folder = args.get("image_folder")
        save_folder = args.get("save_folder")
        names = get_image_file_names(image_folder)[args.get("batch_test_start") or 0:args.get("batch_test_end")]
        if not os.path.exists(save_folder):
            os.makedirs(save_folder)
        for name in names:
            image_path = os.path.join(image_folder, name)
            image_target = load_target(image_path, image_size)
            start_time = time.time()
            log(image_path)
            image_syn, func = test_single(syn, image_target, is_preimage)
            seconds = time.time() - start_time
            log("Finished (%.2f sec)" % seconds)
            syn.save_verbose(os.path.join(save_folder, name.rsplit('.', 1)[0] + ".npz"))
            imwrite_rgb(os.path.join(save_folder, name), get_plottable_data(image_syn, scale=255.))
            results.append((name, float(func), seconds))
    else:
        image_path = args.get("image_path")
        image_target = load_target(image_path, image_size)
        start_time = time.time()
        log(image_path)
        image_syn, func = test_single(syn, image_target, is_preimage)
        seconds = time.time() - start_time
        log("Finished (%.2f sec)" % seconds)
        save_folder = args.get("save_folder")
        if save_folder:
            if not os.path.exists(save_folder):
                os.makedirs(save_folder)
            imwrite_rgb(os.path.join(save_folder, os.path.basename(image_path)), get_plottable_data(image_syn))
        results.append((os.path.basename(image_path), float(func), seconds))
    return results


if __name__ == "__main__":
    ps = get_parser()
    run(ps.parse_args())
