from .model import SynTexModelDesc, RandomZData
from .prep import build_stage_preparation
