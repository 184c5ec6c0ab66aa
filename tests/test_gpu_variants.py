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
