"""splatco_b200 — B200-native (sm_100a) differentiable render hot path for SplatCo.

Drop-in layers (same names as the reference imports):
  splatco_b200.diff_gaussian_rasterization : GaussianRasterizationSettings, GaussianRasterizer
  splatco_b200.gaussian_renderer           : render, prefilter_voxel, generate_neural_gaussians
Everything computes through libsplatco_b200.so (hand-written CUDA behind a C ABI, include/splatco_b200.h).
There is no CPU or PyTorch fallback: calling an op without the built library raises.
"""
__version__ = "0.1.0"
