"""porous-cfd hot path, B200-native: the physics-informed training step of Gallinator/porous-cfd
(model forward over a point cloud, Navier-Stokes-Darcy residual loss, backward to parameter
gradients) behind the reference's model / loss module API, executed by hand-written sm_100a CUDA
kernels in `csrc/` through the C-ABI declared in `include/pcfd.h`.

Importing the package never touches CUDA; the native library is loaded on first use by
`porous_cfd_b200._lib` and raises if it is missing -- there is no CPU fallback.
"""
__version__ = '0.1.0'
