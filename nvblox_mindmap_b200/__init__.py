"""B200-native reconstruction hot path for mindmap (nvblox_torch drop-in back end).

`nvblox_mindmap_b200` holds the CUDA kernels + C ABI (csrc/ -> libnvbx.so) and the ctypes binding;
the reference-facing Python surface lives in the top-level `nvblox_torch` package.
"""
from nvblox_mindmap_b200.params import NvbxParams, NvbxCounters  # noqa: F401

__version__ = '0.1.0'
