"""Host logic of frame pipelining in the Python surface, without a GPU: which input tensors `Mapper` keeps alive.

While pipelining, the geometry / gather kernels of a feature frame read the frame (and its mask) on streams torch does
not know about, up to four feature frames later; with asynchronous enqueue every input of a queued call is read later
still.  A tensor dropped too early would go back to the caching allocator and could be handed out again, so the
wrapper holds on to exactly the tensors whose addresses it passed -- including the CONTIGUOUS COPY it made of a strided
mask, not the caller's original.  (The C library is replaced by a stub that accepts every call.)
"""
import ctypes as C

import pytest
import torch

import nvblox_torch.mapper as M


class _FakeTensor:
    """Enough of a CUDA tensor for the wrappers: shape / dtype / contiguity / address."""

    def __init__(self, shape, dtype, contiguous=True, ptr=0x1000):
        self.shape, self.dtype, self._contig, self._ptr, self.is_cuda = shape, dtype, contiguous, ptr, True

    def is_contiguous(self):
        return self._contig

    def contiguous(self):
        return _FakeTensor(self.shape, self.dtype, True, self._ptr + 0x100)

    def data_ptr(self):
        return self._ptr

    def dim(self):
        return len(self.shape)


class _StubLib:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def f(*a, **k):
            self.calls.append((name, a))
            return 0
        return f


@pytest.fixture
def mapper(monkeypatch):
    monkeypatch.setattr(M, 'check_integrator_inputs', lambda *a, **k: None)
    monkeypatch.setattr(M, 'current_stream_ptr', lambda d: 0)
    m = M.Mapper.__new__(M.Mapper)
    m._voxel_sizes, m._feature_channels, m._lib, m._handle, m._device = [0.02], 16, _StubLib(), C.c_void_p(1), 0
    m._pipelining, m._async_enqueue, m._last_export, m._held_frames = False, False, {}, {}
    yield m
    m._handle = None      # nothing to destroy


def _frames():
    T, K = torch.eye(4), torch.eye(3)
    d = _FakeTensor((8, 8), torch.float32, ptr=0x1000)
    f = _FakeTensor((8, 8, 16), torch.float16, ptr=0x2000)
    strided_mask = _FakeTensor((8, 8), torch.uint8, contiguous=False, ptr=0x3000)
    return T, K, d, f, strided_mask


def test_nothing_is_held_without_pipelining(mapper):
    T, K, d, f, mk = _frames()
    mapper.add_depth_frame(d, T, K, mk)
    mapper.add_feature_frame(f, T, K, mk)
    assert mapper._held_frames == {}


def test_pipelining_holds_feature_frames_and_the_mask_copy(mapper):
    T, K, d, f, mk = _frames()
    mapper.set_pipelining(True)
    assert mapper._lib.calls[-1][0] == 'nvbx_set_pipelining' and mapper._lib.calls[-1][1][1] == 1
    mapper.add_depth_frame(d, T, K, mk)           # consumed on the caller's stream: not held
    mapper.add_feature_frame(f, T, K, mk)
    held = list(mapper._held_frames[0])
    assert [t.data_ptr() for t in held] == [0x2000, 0x3100]     # the frame and the CONTIGUOUS copy of the mask
    name, args = mapper._lib.calls[-1]
    assert name == 'nvbx_integrate_features' and args[2] == 0x2000 and args[6] == 0x3100   # ... the address that was passed
    for _ in range(20):
        mapper.add_feature_frame(f, T, K)
    assert len(mapper._held_frames[0]) == M._HOLD_PIPELINED      # bounded
    mapper.pipeline_join()
    assert mapper._held_frames == {}


def test_async_enqueue_holds_every_input(mapper):
    T, K, d, f, mk = _frames()
    mapper.set_pipelining(True, async_enqueue=True)
    assert mapper._lib.calls[-1][1][1] == 2
    mapper.add_depth_frame(d, T, K, mk)
    mapper.add_feature_frame(f, T, K)
    assert [t.data_ptr() for t in mapper._held_frames[0]] == [0x1000, 0x3100, 0x2000]
    assert mapper._held_frames[0].maxlen == M._HOLD_ASYNC
    # a sequence call: each frame's OWN masks are held
    mk2 = _FakeTensor((8, 8), torch.uint8, contiguous=True, ptr=0x4000)
    mapper._held_frames = {}
    mapper.integrate_frames([d, d], [f, f], [T, T], K, depth_masks=[mk, None], feature_masks=[None, mk2])
    assert sorted(t.data_ptr() for t in mapper._held_frames[0]) == [0x1000, 0x1000, 0x2000, 0x2000, 0x3100, 0x4000]
    mapper.set_pipelining(False)
    assert mapper._held_frames == {} and not mapper._async_enqueue

