"""Stage the REFERENCE's own test files for tests/test_gpu_reference_suite.py.

The reference's pytest files (nvblox_torch) and mindmap's mapping helpers are copied UNMODIFIED from
/root/reference into tests/ref_tests/_ref/ -- a git-ignored directory, so no reference source enters the
repository's history, but one that travels to the GPU box with the snapshot (like oracle/_ref and the built .so
files).  `__graft_entry__.build()` calls stage() whenever /root/reference is mounted; on the GPU box (no
/root/reference) the already staged copy is used.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, '_ref')
REF = '/root/reference'
NT_TESTS = 'submodules/nvblox/nvblox_torch/nvblox_torch/tests'

# (reference path, staged path)
FILES = [(f'{NT_TESTS}/{f}', f'nvblox_torch_tests/tests/{f}') for f in (
    'test_mapper_masking.py', 'test_layer.py', 'test_indexing.py', 'test_timing.py', 'test_mapper_add_frames.py',
    'helpers/camera_utils.py', 'helpers/scene_utils.py')] + [
    ('mindmap/mapping/helpers/nvblox_mapping_helpers.py', 'mindmap_pkg/mindmap/mapping/helpers/nvblox_mapping_helpers.py'),
    ('mindmap/mapping/nvblox_mapper_constants.py', 'mindmap_pkg/mindmap/mapping/nvblox_mapper_constants.py'),
    ('mindmap/image_processing/image_mask_operations.py', 'mindmap_pkg/mindmap/image_processing/image_mask_operations.py'),
    ('mindmap/tasks/tasks.py', 'mindmap_pkg/mindmap/tasks/tasks.py'),
]
# package markers (ours, empty)
INITS = ['nvblox_torch_tests/tests/__init__.py', 'nvblox_torch_tests/tests/helpers/__init__.py',
         'mindmap_pkg/mindmap/__init__.py', 'mindmap_pkg/mindmap/mapping/__init__.py',
         'mindmap_pkg/mindmap/mapping/helpers/__init__.py', 'mindmap_pkg/mindmap/image_processing/__init__.py',
         'mindmap_pkg/mindmap/tasks/__init__.py']


def stage(ref_root: str = REF) -> bool:
    """Copy the files; returns False (and stages nothing) when the reference is not mounted."""
    if not os.path.isdir(ref_root):
        return False
    for src, dst in FILES:
        out = os.path.join(STAGED, dst)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(ref_root, src), out)
    for rel in INITS:
        out = os.path.join(STAGED, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        open(out, 'a').close()
    return True


def is_staged() -> bool:
    return all(os.path.exists(os.path.join(STAGED, dst)) for _, dst in FILES)


if __name__ == '__main__':
    print('staged' if stage() else 'reference not mounted')
