"""B200-native corruption pipeline: drop-in for scripts/augmentations.py of
ysbbin/Robust-Object-Detection (noise / motion blur / lowres / random one-of-three),
computed by hand-written sm_100a CUDA kernels behind the C ABI in include/rod_b200.h.

Submodules
  augmentations  the reference's module surface (same names, same signatures)
  batch          device-resident ragged batches (plans) + the fused training-path driver
  _native        ctypes binding of librod_b200.so (fails loudly when it is missing)
"""
__version__ = "0.1.0"
