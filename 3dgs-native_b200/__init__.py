"""gsb200 -- B200-native 3D Gaussian Splatting rasterizer behind the reference's operator API.

Sub-modules mirror the reference's module names for the hot path:
``forward.render_gaussians``, ``backward.backward``, ``optimizer.*``, ``loss.*``.
The CUDA library (csrc/libgsb200.so) is loaded lazily by ``_lib``; there is no CPU fallback.
"""
__version__ = "0.1.0"
