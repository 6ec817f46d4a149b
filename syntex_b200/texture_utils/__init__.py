# same exports as the reference's texture_utils/__init__.py:1-2
from .utils import build_gram, build_texture_loss, gram_grad_double_bwd
from .vgg19_extractor import VGG_MEAN, VGG_MEAN_RGB, Vgg19Extractor
from .vgg19 import Vgg19, PARAM_SHAPE
