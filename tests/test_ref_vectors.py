"""The CPU oracle against REFERENCE-COMPILED golden vectors (tests/golden/ref_nvcc_*.npz).

The vectors were produced on a B200 by the reference's own device code -- nvblox's kernels and the scalar
functions they call, compiled verbatim with nvblox's nvcc flags (oracle/ref_snippets/, generator
tests/golden/make_ref_vectors.py).  The oracle in its default "nvcc" floating-point model must reproduce every
one of them BIT FOR BIT; the individually-rounded "ieee" model of round 1 must not (it is what the reference
binary does not compute).  CPU only: nothing here needs a GPU or /root/reference.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import ref_scenarios as RS

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8).reshape(a.shape + (a.dtype.itemsize,)) if a.dtype.kind == 'f' else a


def _diff(a, b):
    assert a.shape == b.shape, f'shape {a.shape} vs golden {b.shape}'
    return int((_bits(a) != _bits(b)).sum())


@pytest.fixture(autouse=True)
def _nvcc_model():
    O.lib().orc_set_fp_model(1)
    yield
    O.lib().orc_set_fp_model(1)


def test_function_probes_bit_exact():
    """interpolatePixels<__half|float>, blendTwoArrays, interpolateVertex, getBlockAndVoxelIndexFromPositionInLayer,
    projectThreadVoxel, UpdateTsdfVoxelFunctor (all six weighting functions)."""
    gold = np.load(os.path.join(GOLD, 'ref_nvcc_functions.npz'))
    inp = RS.function_inputs(16)
    mine = RS.function_outputs_oracle(inp, 16)
    assert set(mine) == set(gold.files)
    for k in gold.files:
        assert _diff(mine[k], gold[k]) == 0, f'{k}: oracle differs from the reference-compiled output'


def test_function_probes_ieee_model_is_not_the_reference():
    """The uncontracted model of round 1 disagrees with the reference binary on every arithmetic probe."""
    gold = np.load(os.path.join(GOLD, 'ref_nvcc_functions.npz'))
    O.lib().orc_set_fp_model(0)
    mine = RS.function_outputs_oracle(RS.function_inputs(16), 16)
    for k in ('interp_half_out', 'interp_float_out', 'blend_out', 'vertex_out', 'bv_out_0', 'project_out_0',
              'tsdf_out_2'):
        assert _diff(mine[k], gold[k]) > 0, k


@pytest.mark.parametrize('name', list(RS.SCENARIOS))
def test_pipeline_bit_exact(name):
    """Whole frame sequences: per-frame block lists and synthetic depth images, the final TSDF / feature / colour
    layers, the un-welded marching-cubes vertices and the voxel the mesh paint picks -- all produced by the
    reference's kernels -- equal the oracle's."""
    gold = np.load(os.path.join(GOLD, f'ref_nvcc_pipeline_{name}.npz'))
    sc = RS.SCENARIOS[name]
    mine = RS.run_scenario(sc, 16, RS.OracleBackend(sc, 16))
    mine.update(RS.oracle_mesh_rows(sc, 16, mine))
    assert set(mine) == set(gold.files)
    for k in gold.files:
        assert _diff(mine[k], gold[k]) == 0, f'{name}/{k}: oracle differs from the reference kernels'
    assert (mine['feat_data'][..., -1] != 0).sum() > 1000 and len(mine['mesh_vertex_bits']) > 1000


def test_pipeline_c768_digests():
    """The headline channel count: the maps are too large to commit, so digests and a sample of rows."""
    import hashlib
    gold = np.load(os.path.join(GOLD, 'ref_nvcc_pipeline_cube_stacking_c768.npz'))
    dg = json.loads(str(gold['digests']))
    sc = RS.SCENARIOS['cube_stacking']
    mine = RS.run_scenario(sc, 768, RS.OracleBackend(sc, 768))
    for k, v in mine.items():
        assert hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest() == dg[k], k
    rows = mine['feat_data'].reshape(-1, 769)[gold['sample_rows']].view(np.uint16)
    assert np.array_equal(rows, gold['sample_values'])


def test_ieee_model_changes_a_block_list():
    """Contraction is not only a value-level matter: on the drill-in-box sequence the uncontracted model
    selects a different TSDF block list than the reference's kernel."""
    gold = np.load(os.path.join(GOLD, 'ref_nvcc_pipeline_drill_in_box.npz'))
    O.lib().orc_set_fp_model(0)
    sc = RS.SCENARIOS['drill_in_box']
    mine = RS.run_scenario(sc, 16, RS.OracleBackend(sc, 16))
    assert not np.array_equal(mine['frame_tsdf_counts'], gold['frame_tsdf_counts'])
