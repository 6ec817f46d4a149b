from .model import SynTexModelDesc, RandomZData
from .texture_op import CudaTexture, build_stage_preparation, vgg_stage_preparation, vgg_texture
