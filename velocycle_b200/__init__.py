"""velocycle_b200: B200-native fused ELBO+gradient hot path for VeloCycle's phase / velocity SVI.

See DESIGN.md.  The CUDA library (``libvcb.so``, C-ABI in ``include/vcb.h``) is loaded lazily by
``velocycle_b200._lib``; there is no CPU fallback.
"""
__version__ = "0.1.0"
