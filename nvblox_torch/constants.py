"""Library constants (reference: nvblox_torch/constants.py, py_nvblox.cu `Constants`).

The reference bakes the feature length in at compile time (NVBLOX_FEATURE_ARRAY_NUM_ELEMENTS,
nvblox/core/feature_array.h:22-28; mindmap builds with 768, docker/install_nvblox.sh:24-25).  Here it is a
run-time value: it defaults to 768, can be preset with the NVBLOX_FEATURE_ARRAY_NUM_ELEMENTS environment
variable and changed with `constants.set_feature_array_num_elements(n)` before a Mapper is created.
"""
import os


class _Constants:
    """Collection of nvblox constants."""

    def __init__(self) -> None:
        self._feature_elems = int(os.environ.get('NVBLOX_FEATURE_ARRAY_NUM_ELEMENTS', '768'))

    def feature_array_num_elements(self) -> int:
        return self._feature_elems

    def set_feature_array_num_elements(self, n: int) -> None:
        """(ours) choose the feature length used by Mappers created from now on; multiple of 8."""
        n = int(n)
        assert n > 0 and n % 8 == 0, 'feature length must be a positive multiple of 8'
        self._feature_elems = n

    def feature_array_element_size(self) -> int:
        return 2    # fp16

    def esdf_unknown_distance(self) -> float:
        return -1000.0    # kept for API compatibility; there is no ESDF layer on this path


constants = _Constants()
