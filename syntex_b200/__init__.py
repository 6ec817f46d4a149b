"""syntex_b200 — B200-native implementation of the SynTex hot path (VGG19 features, Gram style loss, L-BFGS).

Mirrors the reference's Python surface for that path:
  syntex_b200.texture_utils   Vgg19, Vgg19Extractor, build_gram, build_texture_loss, VGG_MEAN, VGG_MEAN_RGB
  syntex_b200.gatys           Synthesizer
  syntex_b200.po              SynTexModelDesc (hot-path doors _build_extractor / _build_loss)
All arithmetic runs in the CUDA library built from syntex_b200/csrc (C ABI: include/syntex_b200.h).
"""
__version__ = "0.1.0"
