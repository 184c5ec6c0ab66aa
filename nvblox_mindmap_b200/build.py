"""Build libnvbx.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m nvblox_mindmap_b200.build [--force]

The library is written for Blackwell (sm_100a) only.  -fmad=false / -ffp-contract=off implement the
floating-point contract of csrc/nvbx_math.cuh (no fused multiply-add anywhere in the geometry), which
is what makes host, device and CPU oracle agree bit for bit.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libnvbx.so')
# tuning aid (NVBX_PROFILE=1): same sources with -DNVBX_PROFILE_COUNTERS (in-kernel step / cycle counters)
PROFILE_LIB_PATH = os.path.join(LIB_DIR, 'libnvbx_prof.so')
SOURCES = ['nvbx.cu']
DEPS = ['nvbx.cu', 'nvbx_kernels.cuh', 'nvbx_upsample.cuh', 'nvbx_mesh.cuh', 'nvbx_export.cuh', 'nvbx_map.cuh', 'nvbx_math.cuh', 'mc_tables.h',
        os.path.join('..', '..', 'include', 'nvbx_c_api.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-fmad=false', '-prec-div=true', '-prec-sqrt=true', '-ftz=false',
    '-Xcompiler', '-fPIC,-ffp-contract=off,-fno-fast-math,-O2',
    '-shared', '-cudart', 'shared',
]


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return 'nvcc'


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, profile: bool = False) -> str:
    out = PROFILE_LIB_PATH if profile else LIB_PATH
    if not force and not needs_build() and os.path.exists(out) and \
            os.path.getmtime(out) >= max(os.path.getmtime(os.path.join(CSRC, d)) for d in DEPS):
        return out
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
        (['-DNVBX_PROFILE_COUNTERS'] if profile else []) + ['-o', out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + res.stdout)
    if verbose:
        print(res.stdout)
    return out


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, profile='--profile' in sys.argv))
