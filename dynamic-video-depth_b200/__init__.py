"""dvd_b200 — B200-native hot path of google/dynamic-video-depth (per-video test-time optimisation step).

Only the path named by BASELINE.json.north_star lives here: the fused
un-project → scene-flow-advect → re-project → flow-warp → loss kernels, the scene-flow MLP,
the flat Adam, and the host-side mirror of the reference's plug-in surface.
The CUDA extension (csrc/ → libdvd_b200.so, C ABI in include/dvd_b200.h) is mandatory on the
compute path: there is no CPU fallback.
"""
__version__ = '0.1.0'
