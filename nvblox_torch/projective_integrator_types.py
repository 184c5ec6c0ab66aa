"""Integrator type enum (reference: nvblox_torch/projective_integrator_types.py)."""
from enum import Enum


class ProjectiveIntegratorType(Enum):
    """Kinds of projective (depth) integrators a map can be built with."""
    TSDF = 'tsdf'
    OCCUPANCY = 'occupancy'    # not on mindmap's path; rejected by Mapper
