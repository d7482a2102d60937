"""matx_b200 — B200 (sm_100a) engine for MatX's fused-elementwise / reduction hot path.

The product is `libmatx_b200.so` (C ABI in include/matx_b200.h, CUDA sources in matx_b200/csrc/).  This
package holds the build script and a Python mirror of the MatX operator surface (`matx_b200.ops`) that lowers
statements to the same C ABI the C++ header shim uses.  Importing `matx_b200.ops` loads the library and fails
loudly if it is not built; there is no CPU / PyTorch fallback.
"""
__all__ = ["build"]
