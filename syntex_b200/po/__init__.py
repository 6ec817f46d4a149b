from .model import SynTexModelDesc, RandomZData
