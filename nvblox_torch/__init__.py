"""Drop-in `nvblox_torch` package for mindmap's reconstruction hot path, backed by libnvbx.so.

Module paths, class names and method signatures mirror the reference wrapper
(submodules/nvblox/nvblox_torch/nvblox_torch/*.py) so that `mindmap/mapping/helpers/*.py` imports
resolve unchanged; underneath, the torch custom classes of `libpy_nvblox.so` are replaced by the C ABI
in include/nvbx_c_api.h (hand-written sm_100a kernels).  Only the path SURVEY.md section 8 names is
provided: TSDF + feature mapping, feature mesh export, layer views and TSDF/feature point queries.
"""
__version__ = '0.1.0+b200'
__git_sha__ = 'nvbx'
