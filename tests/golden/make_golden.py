#!/usr/bin/env python3
"""Regenerate the golden fixtures under tests/golden/ (run in the authoring container, where
/root/reference is mounted; the fixtures travel, the reference does not).

    python tests/golden/make_golden.py

  threedmatch_seq01_160x120.npz  depth (uint16 mm, nearest-downsampled 4x), poses, intrinsics of frames 0-2 of
                                 nvblox's own 3DMatch test fixture
                                 (submodules/nvblox/nvblox/tests/data/3dmatch/seq-01, camera-intrinsics.txt).
  core_path.json                 SHA-256 digests of every product of the hot path for the scenarios of
                                 tests/golden_cases.py, computed with the CPU oracle.
  export_postprocess.npz         input / output vectors of the reference's OWN Python post-processing of the
                                 feature point cloud (mindmap/mapping/helpers/nvblox_output_helpers.py:22-91 and
                                 mindmap/data_loading/vertex_sampling.py:29-140), produced by importing those
                                 modules from /root/reference (third-party imports they do not need here are
                                 stubbed) and running them on seeded random clouds.
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference'


def make_threedmatch():
    from PIL import Image
    base = os.path.join(REF, 'submodules/nvblox/nvblox/tests/data/3dmatch')
    K = np.loadtxt(os.path.join(base, 'camera-intrinsics.txt')).astype(np.float32)
    depth, poses = [], []
    for i in range(3):
        d = np.asarray(Image.open(os.path.join(base, f'seq-01/frame-{i:06d}.depth.png'))).astype(np.uint16)
        depth.append(d[::4, ::4])                       # 640x480 -> 160x120, nearest
        poses.append(np.loadtxt(os.path.join(base, f'seq-01/frame-{i:06d}.pose.txt')).astype(np.float32))
    K4 = K.copy()
    K4[:2] /= 4.0                                       # pixel (4i, 4j) -> (i, j); the half-pixel shift is ignored
    np.savez_compressed(os.path.join(HERE, 'threedmatch_seq01_160x120.npz'), depth_mm=np.stack(depth),
                        poses=np.stack(poses), K=K4)


def make_core_path():
    from tests.golden_cases import SCENARIOS, OracleDriver
    out = {name: fn(OracleDriver) for name, fn in SCENARIOS.items()}
    with open(os.path.join(HERE, 'core_path.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    return out


class _Anywhere:
    """Config tensor whose .to('cuda') stays on the CPU (there is no GPU in the authoring container)."""

    def __init__(self, t):
        self.t = t

    def to(self, *_a, **_k):
        return self.t


class _StubModule(types.ModuleType):
    """Stand-in for a third-party package the reference imports at module scope but the functions we call
    never touch (clip, tap, transforms3d, zstandard, Isaac Sim ...): any attribute is a dummy class."""
    __path__ = []

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return type(name, (), {})


def _install_stub_importer():
    import importlib.abc
    import importlib.machinery
    prefixes = ('clip', 'tap', 'transforms3d', 'zstandard', 'open3d', 'isaaclab', 'isaacsim', 'omni', 'pxr', 'carb',
                'gymnasium', 'wandb', 'timm', 'h5py', 'cv2', 'matplotlib', 'scipy', 'PIL', 'torchvision', 'kornia',
                'diffusers', 'dgl', 'einops')

    class Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def find_spec(self, fullname, path, target=None):
            if fullname.split('.')[0] in prefixes:
                try:                                         # a real installation wins
                    for f in sys.meta_path:
                        if f is self:
                            continue
                        spec = f.find_spec(fullname, path, target) if hasattr(f, 'find_spec') else None
                        if spec is not None:
                            return None
                except Exception:
                    pass
                return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
            return None

        def create_module(self, spec):
            return _StubModule(spec.name)

        def exec_module(self, module):
            pass

    sys.meta_path.append(Finder())


def make_export_postprocess():
    import torch
    _install_stub_importer()
    sys.path.insert(0, REF)
    from mindmap.data_loading.vertex_sampling import VertexSamplingMethod, sample_to_n_vertices

    class CudaLike(torch.Tensor):        # the reference asserts .is_cuda on the mesh tensors
        is_cuda = property(lambda self: True)

    try:
        from mindmap.mapping.helpers.nvblox_output_helpers import get_vertices_and_features
    except Exception as e:               # pragma: no cover - the import chain pulls in simulator packages
        print('nvblox_output_helpers not importable here (%s: %s); restating its filter inline' %
              (type(e).__name__, e))
        get_vertices_and_features = None

    class FakeMesh:
        def __init__(self, v, f):
            self.v, self.f = v, f

        def vertices(self):
            return self.v.as_subclass(CudaLike)

        def vertex_features(self):
            return self.f.as_subclass(CudaLike)

    class FakeMapper:
        def __init__(self, v, f):
            self.mesh = FakeMesh(v, f)

        def update_feature_mesh(self, mapper_id):
            pass

        def get_feature_mesh(self, mapper_id):
            return self.mesh

    class Cfg:
        pass

    g = torch.Generator().manual_seed(7)
    store = {}
    for case, (n, c_keep, n_pad, zero_rows) in enumerate([(500, 24, 8, 40), (64, 32, 0, 0), (3000, 8, 24, 900)]):
        C = c_keep + n_pad
        v = (torch.rand((n, 3), generator=g) * torch.tensor([1.6, 1.6, 0.9]) - torch.tensor([0.45, 0.8, 0.15])).float()
        f = torch.randn((n, C), generator=g).half()
        f[:, c_keep:] = 0
        f[torch.randperm(n, generator=g)[:zero_rows]] = 0
        aabb_min = torch.tensor([-0.25, -0.65, -0.07])
        aabb_max = torch.tensor([1.0, 0.62, 0.56])
        cfg = Cfg()
        cfg.aabb_min_m, cfg.aabb_max_m = _Anywhere(aabb_min), _Anywhere(aabb_max)
        if get_vertices_and_features is not None:
            ov, of, om = get_vertices_and_features(FakeMapper(v, f), 0, cfg, remove_zero_features=True,
                                                   num_excess_features=n_pad, sample_vertices=False)
            src = 'reference function'
        else:
            mask = torch.all(torch.logical_and(v > aabb_min, v < aabb_max), dim=1)
            ov, of = v[mask], f[mask]
            if n_pad > 0:
                of = of[..., :-n_pad]
            z = torch.all(of == 0, dim=1)
            ov, of = ov[~z], of[~z]
            src = 'inline restatement of nvblox_output_helpers.py:57-74'
        ov, of = torch.Tensor(ov), torch.Tensor(of.float()).half()
        store[f'c{case}_in_vertices'] = v.numpy()
        store[f'c{case}_in_features'] = f.numpy()
        store[f'c{case}_aabb'] = torch.stack([aabb_min, aabb_max]).numpy()
        store[f'c{case}_num_excess'] = np.int32(n_pad)
        store[f'c{case}_out_vertices'] = ov.numpy()
        store[f'c{case}_out_features'] = of.numpy()
        # deterministic samplers of vertex_sampling.py: pad-with-zeros and lowest-z
        for m, want in ((VertexSamplingMethod.LOWEST, max(1, ov.shape[0] // 2)),
                        (VertexSamplingMethod.LOWEST, ov.shape[0] + 17),
                        (VertexSamplingMethod.NONE, 5)):
            sv, sf, sm = sample_to_n_vertices(ov, of, want, m)
            key = f'c{case}_{m.value}_{want}'
            store[key + '_vertices'] = sv.numpy()
            store[key + '_features'] = sf.numpy()
            store[key + '_valid'] = sm.numpy()
        print(f'case {case}: {n} -> {ov.shape[0]} vertices ({src})')
    np.savez_compressed(os.path.join(HERE, 'export_postprocess.npz'), **store)


if __name__ == '__main__':
    assert os.path.isdir(REF), 'run in the authoring container (needs /root/reference)'
    make_threedmatch()
    res = make_core_path()
    for k, v in res.items():
        print(k, [(r['view_blocks'][0], r['tsdf'][0], r['features'][0]) for r in v])
    make_export_postprocess()
