"""The selectable kernel variants that are OFF by default (tuning knobs read from the environment when the library
first uses them) produce the same bits as the shipped configuration: each one re-runs the reference-vector test of one
scenario, in all three pipelining modes, in a child process with the knob set (-m gpu).

  NVBX_TRACE_TEAM=1           eight lanes per ray (sphere_trace_team)
  NVBX_TRACE_SPEC=2           speculative march, two samples per round trip
  NVBX_TRACE_MARCH=64         64 marching threads per trace CTA
  NVBX_TRACE_CACHE=4          TSDF blocks staged per warp in shared memory (sphere_trace_ray_cached)
  NVBX_TRACE_FREE=0           no observed-free-space block flag
  NVBX_RAYCAST_EARLY_FLUSH=0 + NVBX_PIPE_GEOM=0 + NVBX_PIPE_GATHER_CTAS=4
                              raycast marks flushed after the wait, geometry kernel on the caller's stream
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [
    {'NVBX_TRACE_TEAM': '1'},
    {'NVBX_TRACE_SPEC': '2'},
    {'NVBX_TRACE_MARCH': '64'},
    {'NVBX_TRACE_CACHE': '4'},
    {'NVBX_TRACE_FREE': '0'},
    {'NVBX_RAYCAST_EARLY_FLUSH': '0', 'NVBX_PIPE_GEOM': '0', 'NVBX_PIPE_GATHER_CTAS': '4'},
]


@pytest.mark.parametrize('knobs', VARIANTS, ids=lambda k: ','.join(f'{a}={b}' for a, b in k.items()))
def test_variant_reproduces_reference_kernels(knobs):
    env = dict(os.environ)
    env.update(knobs)
    r = subprocess.run([sys.executable, '-m', 'pytest', 'tests/test_gpu_ref_vectors.py', '-q', '-x', '--no-header', '-m', 'gpu',
                        '-k', 'cube_stacking or stick_in_bin', '-p', 'no:cacheprovider'],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    tail = (r.stdout or '')[-1500:] + (r.stderr or '')[-500:]
    assert r.returncode == 0, tail
    assert ' passed' in r.stdout and 'failed' not in r.stdout, tail


@pytest.mark.parametrize('mode', ['2'])     # (mode 1 is a subset of what mode 2 exercises; it was run once by hand)
def test_whole_parity_suite_under_env_opt_in(mode):
    """NVBX_PIPELINING=1 / 2 (every Mapper of the process pipelines, 2: with asynchronous enqueue) under callers that
    know nothing about it: the whole CUDA-vs-oracle parity file (decay, clear, masks, colour frames, queries, block
    views, meshes, viewpoint cache, several maps per mapper ... in between frames) and mindmap's own mapping helpers
    still produce the oracle's bits -- every entry point orders itself behind what is queued or in flight."""
    env = dict(os.environ)
    env['NVBX_PIPELINING'] = mode
    r = subprocess.run([sys.executable, '-m', 'pytest', 'tests/test_gpu_parity.py', 'tests/test_gpu_queries.py',
                        'tests/test_gpu_reference_suite.py::test_mindmap_mapping_helpers_drive_the_drop_in',
                        '-q', '-x', '--no-header', '-m', 'gpu', '-p', 'no:cacheprovider'],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout or '')[-2500:] + (r.stderr or '')[-500:]
    assert r.returncode == 0, tail
    assert ' passed' in r.stdout and 'failed' not in r.stdout, tail
