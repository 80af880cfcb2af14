"""fmhr_b200 — B200-native (sm_100a) implementation of FMHR's HAM inverse-rendering hot path.

    fmhr_b200.dr      nvdiffrast.torch-shaped rasterize / interpolate / antialias (autograd Functions)
    fmhr_b200.utils   get_normals / get_radiance / get_matrix / laplacian_smoothing / NCC (models/utils.py names)
    fmhr_b200.ham     HamOptimizer: the fused phase-A / phase-B iteration
    fmhr_b200.synth   seeded synthetic inputs of the BASELINE.json shapes

All compute goes through libfmhr_b200.so (include/fmhr_b200.h); there is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
