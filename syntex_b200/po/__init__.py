from .model import SynTexModelDesc, RandomZData
from .prep import build_stage_preparation
from .texture_op import CudaTexture, vgg_stage_preparation, vgg_texture
