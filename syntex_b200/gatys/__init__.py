from .synthesizer import Synthesizer, get_bounds, resize_crop, test_single
