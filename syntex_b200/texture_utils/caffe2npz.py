"""The weight-file layout of texture_utils/caffe2npz.py:6-21 without Caffe.

The reference converts `vgg19_normalised.caffemodel` with pycaffe: every layer with two blobs becomes
`[w transposed OIHW -> HWIO (2,3,1,0), b]` under its layer name in an .npz of object arrays. pycaffe is not
available here; `convert(params)` does the same transposition for weights already read into numpy (for instance
from a torchvision / ONNX export of the same network), and `load_npz` is the reader Vgg19.__init__ uses."""
import numpy as np


def convert(params):
    """params: {layer: (w [Cout,Cin,kh,kw], b [Cout])} in Caffe/torch blob layout -> {layer: [w HWIO, b]} fp32."""
    output = dict()
    for k, p in params.items():
        if len(p) == 2:
            w = np.transpose(np.asarray(p[0]).astype(np.float32), [2, 3, 1, 0])
            b = np.asarray(p[1]).astype(np.float32)
            output[k] = [w, b]
        else:
            print("Skip layer", k, "with", len(p), "param blobs")
    return output


def save_dict(path, output):
    """np.savez(path, layer=[w, b], ...): object arrays, as caffe2npz.py:6-7 writes them."""
    arrays = {}
    for k, (w, b) in output.items():
        a = np.empty(2, dtype=object)
        a[0], a[1] = np.asarray(w, np.float32), np.asarray(b, np.float32)
        arrays[k] = a
    np.savez(path, **arrays)
    return path


def load_npz(path):
    with np.load(path, encoding="latin1", allow_pickle=True) as z:
        return {k: [np.asarray(z[k][0], np.float32), np.asarray(z[k][1], np.float32)] for k in z.files}
