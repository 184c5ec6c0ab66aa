"""`Mapper`: the surface mindmap/mapping calls (reference: nvblox_torch/mapper.py:34-490).

Same constructor, method names, argument order, assertions and error behaviour as the reference
wrapper; every method forwards to one C-ABI entry point of libnvbx.so (include/nvbx_c_api.h), passing
`tensor.data_ptr()` and torch's current CUDA stream, so image tensors cross zero-copy and all kernels
run stream-ordered with the caller's torch work.  There is no CPU fallback: constructing a Mapper
without a Blackwell GPU raises.
"""
import collections
import ctypes as C
import os
from enum import Enum
from typing import List, Optional

import torch

from nvblox_mindmap_b200 import _capi
from nvblox_mindmap_b200.params import NvbxCounters
from nvblox_mindmap_b200.torch_interop import current_stream_ptr, device_view
from nvblox_torch.constants import constants
from nvblox_torch.layer import ColorLayer, FeatureLayer, TsdfLayer
from nvblox_torch.mapper_params import MapperParams
from nvblox_torch.mesh import ColorMesh, FeatureMesh
from nvblox_torch.projective_integrator_types import ProjectiveIntegratorType


_HOLD_PIPELINED = 12   # input tensors kept alive per map while pipelining: (frame + mask) x (ring of 4 + 2)
_HOLD_ASYNC = 48       # ... with asynchronous enqueue: (depth, mask, features, mask) x (ring of 4 + 8 queued)


class QueryType(Enum):
    """Enum used when querying layers."""
    TSDF = 'tsdf'
    FEATURE = 'feature'
    OCCUPANCY = 'occupancy'
    ESDF = 'esdf'
    ESDF_GRAD = 'esdf_with_gradients'
    COLOR = 'color'


def _pose16(t_w_c: torch.Tensor) -> int:
    """Address of the 16 row-major floats of a CPU pose tensor (read synchronously by the C ABI)."""
    if not t_w_c.is_contiguous():
        t_w_c = t_w_c.contiguous()
        _pose16.keepalive = t_w_c
    return t_w_c.data_ptr()


_F9 = C.c_float * 9


def _fxfycxcy(intrinsics: torch.Tensor):
    """(fx, fy, cx, cy) of a CPU float32 3x3 tensor without four tensor-indexing round trips."""
    if not intrinsics.is_contiguous():
        intrinsics = intrinsics.contiguous()
    k = _F9.from_address(intrinsics.data_ptr())
    return k[0], k[4], k[2], k[5]


_LOWRES_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _strides_like_channels_last(t: torch.Tensor) -> bool:
    """c10's `is_strides_like_channels_last` for a 4-d tensor (what `suggest_memory_format()` consults)."""
    sizes, strides = t.shape, t.stride()
    lo = 0
    for d in (1, 3, 2, 0):
        if sizes[d] == 0 or strides[d] < lo:
            return False
        if d == 0 and lo == strides[1]:
            return False
        lo = strides[d]
        if sizes[d] > 1:
            lo *= sizes[d]
    return True


def lowres_descriptor(features_bchw: torch.Tensor):
    """(tensor to keep alive, c, h, w, dtype code, layout code, torch-kernel code) of a backbone feature map
    shaped [1, c, h, w] or [c, h, w], decided the way `F.interpolate(..., mode='bilinear')` decides: a
    channels-last-strided input with >= 16 channels runs torch's NHWC kernel, everything else the NCHW one."""
    x = features_bchw if features_bchw.ndim == 4 else features_bchw[None]
    assert x.ndim == 4 and x.shape[0] == 1, 'expected a [1, c, h, w] (or [c, h, w]) feature map'
    assert x.dtype in _LOWRES_DTYPES, f'unsupported low-res dtype {x.dtype}'
    _, c, h, w = x.shape
    nhwc = _strides_like_channels_last(x)
    kernel = 1 if (nhwc and c >= 16) else 0
    if nhwc and x.stride() == (h * w * c, 1, w * c, c):
        layout = 1                         # dense HWC memory: crosses zero-copy
    else:
        x, layout = x.contiguous(), 0      # CHW
    return x, c, h, w, _LOWRES_DTYPES[x.dtype], layout, kernel


class Mapper:
    """Accumulates depth (+ feature) frames into voxel-block-hashed TSDF / feature maps on one GPU.

    Several independent maps (`mapper_id`) can live in one Mapper, e.g. static / dynamic
    (mindmap/mapping/helpers/nvblox_mapping_helpers.py:72-76).  `mapper_id = -1` addresses all maps
    where the reference allows it.
    """

    def __init__(self,
                 voxel_sizes_m,
                 integrator_types=ProjectiveIntegratorType.TSDF,
                 mapper_parameters: Optional[MapperParams] = None,
                 device: Optional[int] = None) -> None:
        voxel_sizes = [voxel_sizes_m] if isinstance(voxel_sizes_m, (float, int)) else list(voxel_sizes_m)
        if isinstance(integrator_types, ProjectiveIntegratorType):
            integrator_types = [integrator_types] * len(voxel_sizes)
        assert len(voxel_sizes) == len(integrator_types)
        for t in integrator_types:
            if t != ProjectiveIntegratorType.TSDF:
                raise NotImplementedError('only TSDF maps are on the accelerated path (SURVEY.md 2.3)')
        self._params = MapperParams(mapper_parameters) if mapper_parameters is not None else MapperParams()
        self._voxel_sizes = [float(v) for v in voxel_sizes]
        self._integrator_types = list(integrator_types)
        self._feature_channels = constants.feature_array_num_elements()
        if not torch.cuda.is_available():
            raise RuntimeError('nvblox_torch (B200 back end) needs a CUDA device: there is no CPU fallback')
        self._device = torch.cuda.current_device() if device is None else int(device)
        self._lib = _capi.load()
        nv = self._params.to_nvbx()
        sizes = (C.c_float * len(voxel_sizes))(*self._voxel_sizes)
        handle = C.c_void_p()
        _capi.check(self._lib.nvbx_create(len(voxel_sizes), sizes, C.byref(nv), self._feature_channels, self._device,
                                          C.byref(handle)))
        self._handle = handle
        self._pipelining = False
        self._async_enqueue = False
        self._last_export = {}    # mapper_id -> (vertices ptr, features ptr, rows, channels) of the last export_points
        self._held_frames = {}    # mapper_id -> deque of input tensors kept alive while pipelining (see _hold)
        # Opt-in without touching the caller's code (a drop-in under mindmap's own loop): NVBX_PIPELINING=1 turns frame
        # pipelining on for every Mapper of the process, =2 adds asynchronous enqueue.  The caller then promises what
        # set_pipelining's contract says: frames are not modified in place after they were handed over.
        mode = os.environ.get('NVBX_PIPELINING', '0').strip()
        if mode in ('1', '2'):
            self.set_pipelining(True, async_enqueue=(mode == '2'))
        elif mode not in ('', '0'):
            raise ValueError(f'NVBX_PIPELINING={mode!r}: expected 0, 1 or 2')

    def __del__(self):
        h = getattr(self, '_handle', None)
        if h:
            try:
                self._lib.nvbx_destroy(h)
            except Exception:    # interpreter shutdown
                pass
            self._handle = None

    # -- helpers ----------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return current_stream_ptr(self._device)

    @staticmethod
    def _mask_tensor(mask_frame: Optional[torch.Tensor], image: torch.Tensor) -> Optional[torch.Tensor]:
        """The mask as the C ABI reads it (contiguous uint8 on the device), or None."""
        if mask_frame is None:
            return None
        # py_mapper.cu:90-99: not-on-GPU / size mismatch are logged and the frame is skipped there;
        # here they raise (ALL_ON_GPU_OR_RETURN, checkImageDimensionsEqual).
        assert mask_frame.is_cuda, 'Mask frame should be on device.'
        assert mask_frame.dtype == torch.uint8, 'Mask frame should have type torch.uint8.'
        assert mask_frame.shape[0] == image.shape[0] and mask_frame.shape[1] == image.shape[1], \
            'Mask frame size should match the image.'
        return mask_frame if mask_frame.is_contiguous() else mask_frame.contiguous()

    @staticmethod
    def _mask_ptr(mask_frame: Optional[torch.Tensor], image: torch.Tensor) -> Optional[int]:
        m = Mapper._mask_tensor(mask_frame, image)
        return None if m is None else m.data_ptr()

    def params(self) -> MapperParams:
        return MapperParams(self._params)

    # -- frame integration --------------------------------------------------------------------------------
    def add_depth_frame(self,
                        depth_frame: torch.Tensor,
                        t_w_c: torch.Tensor,
                        intrinsics: torch.Tensor,
                        mask_frame: Optional[torch.Tensor] = None,
                        mapper_id: int = 0) -> None:
        """Integrate a (H, W) float32 CUDA depth frame; t_w_c (4,4) and intrinsics (3,3) are CPU tensors."""
        assert 0 <= mapper_id < len(self._voxel_sizes)
        check_integrator_inputs(depth_frame, t_w_c, intrinsics, 'Depth', 2, torch.float32)
        depth_frame = depth_frame if depth_frame.is_contiguous() else depth_frame.contiguous()
        mask_frame = self._mask_tensor(mask_frame, depth_frame)   # (a contiguous copy, if one was needed, is what is held)
        mask_ptr = None if mask_frame is None else mask_frame.data_ptr()
        fx, fy, cx, cy = _fxfycxcy(intrinsics)
        _capi.check(self._lib.nvbx_integrate_depth(
            self._handle, mapper_id, depth_frame.data_ptr(), depth_frame.shape[0], depth_frame.shape[1], mask_ptr,
            _pose16(t_w_c), fx, fy, cx, cy, self._stream()))
        if self._async_enqueue:
            self._hold(mapper_id, depth_frame, mask_frame)

    def add_color_frame(self,
                        color_frame: torch.Tensor,
                        t_w_c: torch.Tensor,
                        intrinsics: torch.Tensor,
                        mask_frame: Optional[torch.Tensor] = None,
                        mapper_id: int = 0) -> None:
        """Integrate a (H, W, 3) uint8 CUDA colour frame into the colour layer (reference mapper.py:112-129)."""
        assert 0 <= mapper_id < len(self._voxel_sizes)
        check_integrator_inputs(color_frame, t_w_c, intrinsics, 'Color', 3, torch.uint8, 3)
        color_frame = color_frame if color_frame.is_contiguous() else color_frame.contiguous()
        mask_ptr = self._mask_ptr(mask_frame, color_frame)
        fx, fy, cx, cy = _fxfycxcy(intrinsics)
        _capi.check(self._lib.nvbx_integrate_color(
            self._handle, mapper_id, color_frame.data_ptr(), color_frame.shape[0], color_frame.shape[1], mask_ptr,
            _pose16(t_w_c), fx, fy, cx, cy, self._stream()))

    def add_feature_frame(self,
                          feature_frame: torch.Tensor,
                          t_w_c: torch.Tensor,
                          intrinsics: torch.Tensor,
                          mask_frame: Optional[torch.Tensor] = None,
                          mapper_id: int = 0) -> None:
        """Integrate a (H, W, C) float16 CUDA feature frame, C == constants.feature_array_num_elements()."""
        assert 0 <= mapper_id < len(self._voxel_sizes)
        check_integrator_inputs(feature_frame, t_w_c, intrinsics, 'Feature', 3, torch.float16, self._feature_channels)
        feature_frame = feature_frame if feature_frame.is_contiguous() else feature_frame.contiguous()
        mask_frame = self._mask_tensor(mask_frame, feature_frame)
        mask_ptr = None if mask_frame is None else mask_frame.data_ptr()
        fx, fy, cx, cy = _fxfycxcy(intrinsics)
        _capi.check(self._lib.nvbx_integrate_features(
            self._handle, mapper_id, feature_frame.data_ptr(), feature_frame.shape[0], feature_frame.shape[1],
            feature_frame.shape[2], mask_ptr, _pose16(t_w_c), fx, fy, cx, cy, self._stream()))
        if self._pipelining:
            self._hold(mapper_id, feature_frame, mask_frame)

    def integrate_frames(self, depth_frames, feature_frames, poses, intrinsics, mapper_id: int = 0,
                         depth_masks=None, feature_masks=None) -> None:
        """(ours) A SEQUENCE of frames of one map in one call: frame k = add_depth_frame(depth_frames[k], ...) followed
        by add_feature_frame(feature_frames[k], ...) (feature_frames[k] may be None), in order, with one crossing of
        the Python / C boundary for the whole sequence (nvbx_integrate_frames_batch).  For replay / datagen loops
        whose frames are already resident; the map is identical to the one the per-frame calls build.
        `poses`: list of CPU [4,4] tensors; `intrinsics`: one CPU [3,3] tensor or a list."""
        from nvblox_mindmap_b200.params import NvbxFrameJob
        assert 0 <= mapper_id < len(self._voxel_sizes)
        n = len(depth_frames)
        assert len(feature_frames) == n and len(poses) == n
        jobs = (NvbxFrameJob * n)()
        stream = self._stream()
        keep = []
        for k in range(n):
            d, f, T = depth_frames[k], feature_frames[k], poses[k]
            K = intrinsics if isinstance(intrinsics, torch.Tensor) else intrinsics[k]
            check_integrator_inputs(d, T, K, 'Depth', 2, torch.float32)
            d = d if d.is_contiguous() else d.contiguous()
            j = jobs[k]
            j.mapper, j.map_id, j.stream = self._handle.value, mapper_id, stream
            j.height, j.width = int(d.shape[0]), int(d.shape[1])
            j.depth = d.data_ptr()
            dm = self._mask_tensor(None if depth_masks is None else depth_masks[k], d)
            j.depth_mask = None if dm is None else dm.data_ptr()
            fm = None
            if f is not None:
                check_integrator_inputs(f, T, K, 'Feature', 3, torch.float16, self._feature_channels)
                assert f.shape[0] == d.shape[0] and f.shape[1] == d.shape[1], 'Feature frame size should match the depth frame.'
                f = f if f.is_contiguous() else f.contiguous()
                j.channels = int(f.shape[2])
                j.features = f.data_ptr()
                fm = self._mask_tensor(None if feature_masks is None else feature_masks[k], f)
                j.feature_mask = None if fm is None else fm.data_ptr()
            Tc = T if T.is_contiguous() else T.contiguous()
            C.memmove(j.T_L_C, Tc.data_ptr(), 64)
            j.fx, j.fy, j.cx, j.cy = _fxfycxcy(K)
            keep.append((d, f, dm, fm, Tc))
        _capi.check(self._lib.nvbx_integrate_frames_batch(jobs, n, 1))
        if self._pipelining:
            for d, f, dm, fm, _ in keep[-(_HOLD_ASYNC // 4):]:
                self._hold(mapper_id, f, fm)
                if self._async_enqueue:
                    self._hold(mapper_id, d, dm)

    def _hold(self, mapper_id: int, *tensors) -> None:
        """Keep input tensors alive while the library may still read them outside torch's stream order: the gather of
        a feature frame runs on the map's own stream until up to four feature frames later (the ring depth of
        nvbx_map.cuh), and with asynchronous enqueue up to eight more calls may still be waiting to be issued.  Without
        this the caching allocator could hand the memory of a dropped tensor to a new one."""
        held = self._held_frames.get(mapper_id)
        n = _HOLD_ASYNC if self._async_enqueue else _HOLD_PIPELINED
        if held is None or held.maxlen != n:
            held = collections.deque(held or (), maxlen=n)
            self._held_frames[mapper_id] = held
        for t in tensors:
            if t is not None:
                held.append(t)

    def set_pipelining(self, on: bool = True, async_enqueue: bool = False) -> None:
        """(ours) Overlap the memory-bound gather of feature frame i with the latency-bound depth path of frame i + 1
        (include/nvbx_c_api.h: nvbx_set_pipelining).  Results are bit-identical.  While it is on, a feature frame must
        not be overwritten IN PLACE until four more feature frames have been added to that map, or until a call that
        reads the feature layer (decay / clear / mesh / block views / queries / `pipeline_join`); the frames
        themselves are kept alive here.  Meant for replay / datagen loops that hold their frames.

        `async_enqueue`: add_depth_frame / add_feature_frame only validate and queue their arguments; a worker thread
        of the library issues the CUDA work in order (the ~20 us of launches per frame leave the calling thread, so a
        Python loop keeps up with the ~35 us the device needs).  Every other method of the Mapper first waits for the
        queue, so the map reads the same; but a frame is no longer ordered against the CALLER'S OWN later work on the
        stream: depth frames and masks, too, must stay unmodified until the next joining call, and
        `torch.cuda.synchronize()` alone does not wait for queued frames -- call `pipeline_join()` first.  An error in
        a queued frame is raised by the next joining call."""
        _capi.check(self._lib.nvbx_set_pipelining(self._handle, (2 if async_enqueue else 1) if on else 0))
        self._pipelining = bool(on)
        self._async_enqueue = bool(on and async_enqueue)
        if not on:
            self._held_frames = {}

    def pipeline_join(self, mapper_id: int = -1) -> None:
        """(ours) Order torch's current stream behind every feature gather still in flight."""
        _capi.check(self._lib.nvbx_pipeline_join(self._handle, mapper_id, self._stream()))
        self._held_frames = {}

    def integrate_frame_from_host(self, depth, features, t_w_c, intrinsics, depth_mask=None, feature_mask=None,
                                  mapper_id: int = 0) -> None:
        """(ours) depth + feature frame from HOST tensors.  A pinned feature frame is fetched sparsely over PCIe
        (see `set_host_fetch_mode`) and must not be modified until the stream has passed this call."""
        assert 0 <= mapper_id < len(self._voxel_sizes)
        _check_host_frame_inputs(depth, t_w_c, intrinsics, depth_mask, feature_mask)
        assert features.dtype == torch.float16 and not features.is_cuda and features.dim() == 3 and \
            features.is_contiguous(), 'Feature frame should be a contiguous [H, W, C] float16 CPU tensor.'
        assert features.shape[0] == depth.shape[0] and features.shape[1] == depth.shape[1], \
            'Feature frame size should match the depth frame.'
        assert features.shape[2] == self._feature_channels, f'Feature should have {self._feature_channels} channels.'
        fx, fy, cx, cy = _fxfycxcy(intrinsics)
        _capi.check(self._lib.nvbx_integrate_frame_host(
            self._handle, mapper_id, depth.data_ptr(), features.data_ptr(), depth.shape[0], depth.shape[1],
            features.shape[2], None if depth_mask is None else depth_mask.data_ptr(),
            None if feature_mask is None else feature_mask.data_ptr(), _pose16(t_w_c), fx, fy, cx, cy,
            self._stream()))

    def set_host_fetch_mode(self, mode: str) -> None:
        """(ours) 'sparse' (default): a pinned host feature frame is read through its device mapping and only the
        pixels the frame's voxels sample cross PCIe; 'dense': the whole frame is copied."""
        _capi.check(self._lib.nvbx_set_host_fetch_mode(self._handle, {'sparse': 0, 'dense': 1}[mode]))

    def add_feature_frame_lowres(self,
                                 features_bchw: torch.Tensor,
                                 output_size,
                                 t_w_c: torch.Tensor,
                                 intrinsics: torch.Tensor,
                                 mask_frame: Optional[torch.Tensor] = None,
                                 mapper_id: int = 0) -> None:
        """(ours) Integrate the backbone's [1, c, h, w] CUDA feature map as if it had gone through
        mindmap's `FeatureExtractor.compute()` tail -- `scale_image(features_bchw, output_size)`, HWC, zero-pad to
        `constants.feature_array_num_elements()`, `.to(float16)` (feature_extraction.py:188-196,
        nvblox_mapping_helpers.py:255-261) -- and then `add_feature_frame`.  The stored features are bit-identical;
        the (H, W, C) frame is never materialised.  `mask_frame`, `intrinsics` refer to the `output_size` frame."""
        assert 0 <= mapper_id < len(self._voxel_sizes)
        assert features_bchw.is_cuda, 'Feature frame must be on GPU'
        assert not t_w_c.is_cuda and not intrinsics.is_cuda
        assert t_w_c.dtype == torch.float32 and intrinsics.dtype == torch.float32
        H, W = int(output_size[0]), int(output_size[1])
        x, c, h, w, dtype, layout, kernel = lowres_descriptor(features_bchw)
        assert c <= self._feature_channels
        mask_ptr = None
        if mask_frame is not None:
            assert mask_frame.is_cuda and mask_frame.dtype == torch.uint8 and tuple(mask_frame.shape) == (H, W)
            mask_frame = mask_frame if mask_frame.is_contiguous() else mask_frame.contiguous()
            mask_ptr = mask_frame.data_ptr()
        fx, fy, cx, cy = _fxfycxcy(intrinsics)
        _capi.check(self._lib.nvbx_integrate_features_lowres(
            self._handle, mapper_id, x.data_ptr(), h, w, c, dtype, layout, kernel, H, W, mask_ptr, _pose16(t_w_c),
            fx, fy, cx, cy, self._stream()))

    def upsample_features(self, features_bchw: torch.Tensor, output_size, mapper_id: int = 0) -> torch.Tensor:
        """(ours) The (H, W, C) float16 frame `add_feature_frame_lowres` integrates, materialised (parity checks,
        visualisation)."""
        H, W = int(output_size[0]), int(output_size[1])
        x, c, h, w, dtype, layout, kernel = lowres_descriptor(features_bchw)
        out = torch.empty((H, W, self._feature_channels), dtype=torch.float16, device=x.device)
        _capi.check(self._lib.nvbx_upsample_features(
            self._handle, mapper_id, x.data_ptr(), h, w, c, dtype, layout, kernel, H, W, out.data_ptr(),
            self._stream()))
        return out

    def integrate_frame_from_host_lowres(self, depth, features_bchw, t_w_c, intrinsics, depth_mask=None,
                                         feature_mask=None, mapper_id: int = 0) -> None:
        """(ours) depth + low-res feature map from HOST (ideally pinned) tensors: 2.5 MB of H2D per 512^2 x 768 frame
        instead of 385 MB."""
        assert 0 <= mapper_id < len(self._voxel_sizes)
        _check_host_frame_inputs(depth, t_w_c, intrinsics, depth_mask, feature_mask)
        assert not features_bchw.is_cuda, 'Low-res feature map should be on the CPU.'
        x, c, h, w, dtype, layout, kernel = lowres_descriptor(features_bchw)
        fx, fy, cx, cy = _fxfycxcy(intrinsics)
        _capi.check(self._lib.nvbx_integrate_frame_host_lowres(
            self._handle, mapper_id, depth.data_ptr(), x.data_ptr(), h, w, c, dtype, layout, kernel,
            depth.shape[0], depth.shape[1], None if depth_mask is None else depth_mask.data_ptr(),
            None if feature_mask is None else feature_mask.data_ptr(), _pose16(t_w_c), fx, fy, cx, cy,
            self._stream()))

    # -- map maintenance ------------------------------------------------------------------------------------
    def decay(self, mapper_id: int = -1) -> None:
        """Decay TSDF weights and release fully decayed blocks (with their feature / mesh blocks)."""
        assert -1 <= mapper_id < len(self._voxel_sizes)
        _capi.check(self._lib.nvbx_decay(self._handle, mapper_id, self._stream()))

    def clear(self, mapper_id: int = -1) -> None:
        assert -1 <= mapper_id < len(self._voxel_sizes)
        _capi.check(self._lib.nvbx_clear(self._handle, mapper_id, self._stream()))

    def update_esdf(self, mapper_id: int = -1) -> None:
        raise NotImplementedError('ESDF is not on the accelerated path (SURVEY.md 2.3)')

    def update_color_mesh(self, mapper_id: int = -1) -> None:
        """Re-mesh the blocks touched since the last colour-mesh update and refresh their vertex colours."""
        assert -1 <= mapper_id < len(self._voxel_sizes)
        _capi.check(self._lib.nvbx_update_color_mesh(self._handle, mapper_id, self._stream()))

    def update_feature_mesh(self, mapper_id: int = -1) -> None:
        """Re-mesh the blocks touched since the last update and refresh their vertex features."""
        assert -1 <= mapper_id < len(self._voxel_sizes)
        _capi.check(self._lib.nvbx_update_feature_mesh(self._handle, mapper_id, self._stream()))

    def get_color_mesh(self, mapper_id: int = 0) -> ColorMesh:
        """Serialised colour mesh as zero-copy device views: vertices [N,3] f32, colours [N,3] u8, triangles."""
        assert 0 <= mapper_id < len(self._voxel_sizes)
        v, c, t = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nv, nt = C.c_int64(), C.c_int64()
        _capi.check(self._lib.nvbx_get_color_mesh(self._handle, mapper_id, C.byref(v), C.byref(c), C.byref(t),
                                                  C.byref(nv), C.byref(nt)))
        d = self._device
        return ColorMesh(c_mesh={
            'vertices': device_view(v.value, (nv.value, 3), torch.float32, d, owner=self),
            'appearances': device_view(c.value, (nv.value, 3), torch.uint8, d, owner=self),
            'triangles': device_view(t.value, (nt.value, 3), torch.int32, d, owner=self),
        })

    def get_feature_mesh(self, mapper_id: int = 0) -> FeatureMesh:
        """Serialised feature mesh as zero-copy device views (no kernel, no copy)."""
        assert 0 <= mapper_id < len(self._voxel_sizes)
        v, f, t = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nv, nt = C.c_int64(), C.c_int64()
        _capi.check(self._lib.nvbx_get_feature_mesh(self._handle, mapper_id, C.byref(v), C.byref(f), C.byref(t),
                                                    C.byref(nv), C.byref(nt)))
        c, d = self._feature_channels, self._device
        return FeatureMesh(c_mesh={
            'vertices': device_view(v.value, (nv.value, 3), torch.float32, d, owner=self),
            'appearances': device_view(f.value, (nv.value, c), torch.float16, d, owner=self),
            'triangles': device_view(t.value, (nt.value, 3), torch.int32, d, owner=self),
        })

    # -- fused export post-processing (ours) -------------------------------------------------
    def export_points(self, mapper_id: int, aabb_min_m, aabb_max_m, num_excess_features: int = 0,
                      remove_zero_features: bool = True, vertices: Optional[torch.Tensor] = None,
                      features: Optional[torch.Tensor] = None):
        """AABB filter + excess-channel strip + zero-row drop of the feature point cloud in one device pass
        (mindmap/mapping/helpers/nvblox_output_helpers.py:57-74).  Works on the current feature mesh of
        `mapper_id`, or on explicit (vertices [N,3] f32, features [N,C] f16) CUDA tensors.  Returns zero-copy views
        (vertices [M,3] f32, features [M,C-excess] f16) valid until the next export on this mapper_id."""
        assert 0 <= mapper_id < len(self._voxel_sizes)
        lo = (C.c_float * 3)(*[float(v) for v in aabb_min_m])
        hi = (C.c_float * 3)(*[float(v) for v in aabb_max_m])
        ov, of = C.c_void_p(), C.c_void_p()
        if vertices is None:
            vp, fp_, n, ch = None, None, 0, self._feature_channels
        else:
            assert vertices.is_cuda and features.is_cuda and vertices.dtype == torch.float32 and \
                features.dtype == torch.float16 and vertices.shape[0] == features.shape[0]
            vertices, features = vertices.contiguous(), features.contiguous()
            vp, fp_, n, ch = vertices.data_ptr(), features.data_ptr(), vertices.shape[0], features.shape[1]
        count = _capi.check(self._lib.nvbx_export_points(
            self._handle, mapper_id, vp, fp_, n, ch, lo, hi, int(num_excess_features), int(bool(remove_zero_features)),
            C.byref(ov), C.byref(of), self._stream()))
        keep = ch - int(num_excess_features)
        d = self._device
        self._last_export[mapper_id] = (ov.value, of.value, count, keep)   # output_helpers.sample_to_n_vertices
        return (device_view(ov.value, (count, 3), torch.float32, d, owner=self),
                device_view(of.value, (count, keep), torch.float16, d, owner=self))

    def gather_points(self, mapper_id: int, indices: Optional[torch.Tensor], n_out: int, channels: int,
                      features_dtype: torch.dtype = torch.float16, n_rows: Optional[int] = None):
        """Rows `indices` (CPU int64 tensor; None = the first n_rows rows) of the last export, zero-padded to n_out
        rows, features optionally widened to float32 in the same pass (vertex_sampling.py:29-108)."""
        assert features_dtype in (torch.float16, torch.float32)
        dev = f'cuda:{self._device}'
        out_v = torch.empty((n_out, 3), dtype=torch.float32, device=dev)
        out_f = torch.empty((n_out, channels), dtype=features_dtype, device=dev)
        if indices is not None:
            indices = indices.to(torch.int64).contiguous()
            assert indices.is_cpu
            ip, n_idx = indices.data_ptr(), indices.shape[0]
        else:
            ip, n_idx = None, int(n_rows)
        _capi.check(self._lib.nvbx_gather_points(self._handle, mapper_id, ip, n_idx, n_out, out_v.data_ptr(),
                                                 out_f.data_ptr(), int(features_dtype == torch.float32),
                                                 self._stream()))
        return out_v, out_f

    # -- layer views ------------------------------------------------------------------------------------------
    def tsdf_layer_view(self, mapper_id: int = 0) -> TsdfLayer:
        assert 0 <= mapper_id < len(self._voxel_sizes)
        return TsdfLayer(voxel_size_m=self._voxel_sizes[mapper_id], c_layer=(self, mapper_id))

    def feature_layer_view(self, mapper_id: int = 0) -> FeatureLayer:
        assert 0 <= mapper_id < len(self._voxel_sizes)
        return FeatureLayer(voxel_size_m=self._voxel_sizes[mapper_id], c_layer=(self, mapper_id))

    def color_layer_view(self, mapper_id: int = 0) -> ColorLayer:
        assert 0 <= mapper_id < len(self._voxel_sizes)
        return ColorLayer(voxel_size_m=self._voxel_sizes[mapper_id], c_layer=(self, mapper_id))

    def save_map(self, map_fname: str, mapper_id: int) -> None:
        """Write the map's TSDF / colour / feature layers as an nvblox `.nvblx` layer cake (sqlite)."""
        assert 0 <= mapper_id < len(self._voxel_sizes)
        from nvblox_mindmap_b200 import disk_helpers
        disk_helpers.save_map(self, map_fname, mapper_id)

    def load_from_file(self, filename: str, mapper_id: int) -> None:
        """Replace the map by the layers of a `.nvblx` file and re-mesh it (reference mapper.py:271-279)."""
        assert 0 <= mapper_id < len(self._voxel_sizes)
        from nvblox_mindmap_b200 import disk_helpers
        return disk_helpers.load_map(self, filename, mapper_id)

    def num_mappers(self) -> int:
        return int(self._lib.nvbx_num_maps(self._handle))

    # -- queries ------------------------------------------------------------------------------------------------
    def _maybe_allocate(self, size, tensor=None, dtype=torch.float32, value=None) -> torch.Tensor:
        if tensor is None:
            dev = f'cuda:{self._device}'
            return torch.zeros(size, dtype=dtype, device=dev) if value is None else \
                torch.full(size, value, dtype=dtype, device=dev)
        assert tuple(tensor.shape) == tuple(size), f'Expected preallocated size: {size}.'
        return tensor

    def query_layer(self,
                    query_type: QueryType,
                    query: torch.Tensor,
                    output: Optional[torch.Tensor] = None,
                    mapper_id: int = -1) -> torch.Tensor:
        """Look up N positions: TSDF -> [N,2] (distance, weight); FEATURE -> [N,C+1] fp16 (features, weight).

        Rows whose position falls in no allocated block keep the pre-filled value (zeros).
        """
        assert -1 <= mapper_id < len(self._voxel_sizes)
        num_queries = query.shape[0]
        ok = query.is_cuda and query.dtype == torch.float32 and query.dim() == 2 and query.shape[1] == 3
        if query_type == QueryType.TSDF:
            output = self._maybe_allocate((num_queries, TsdfLayer.num_elements_per_voxel()), output)
            if mapper_id == -1:
                raise NotImplementedError('multi-mapper TSDF query is not on the accelerated path')
            if not (ok and output.is_cuda and output.dtype == torch.float32):
                raise ValueError(f'Query failed for: {query_type}')
            q = query if query.is_contiguous() else query.contiguous()
            assert output.is_contiguous()
            _capi.check(self._lib.nvbx_query_tsdf(self._handle, mapper_id, q.data_ptr(), num_queries,
                                                  output.data_ptr(), self._stream()))
            return output
        if query_type == QueryType.FEATURE:
            output = self._maybe_allocate((num_queries, self._feature_channels + 1), output, dtype=torch.float16)
            assert mapper_id >= 0, 'Only single mapper query is supported for features'
            if not (ok and output.is_cuda and output.dtype == torch.float16):
                raise ValueError(f'Query failed for: {query_type}')
            q = query if query.is_contiguous() else query.contiguous()
            assert output.is_contiguous()
            _capi.check(self._lib.nvbx_query_features(self._handle, mapper_id, q.data_ptr(), num_queries,
                                                      output.data_ptr(), self._stream()))
            return output
        raise NotImplementedError(f'Query type {query_type} not implemented')

    # -- accounting (ours) ----------------------------------------------------------------------------------------
    def counters(self, mapper_id: int = 0) -> dict:
        """Device-side work counters that define the algorithmic bytes."""
        c = NvbxCounters()
        _capi.check(self._lib.nvbx_get_counters(self._handle, mapper_id, C.byref(c), self._stream()))
        return c.as_dict()

    def reset_counters(self, mapper_id: int = 0) -> None:
        _capi.check(self._lib.nvbx_reset_counters(self._handle, mapper_id, self._stream()))

    def print_timing(self) -> str:
        from nvblox_torch.timer import timer_status_string
        return timer_status_string()


def _check_host_frame_inputs(depth, t_w_c, intrinsics, depth_mask, feature_mask) -> None:
    """Input contract of the host-frame entry points: the C side reads these buffers by raw pointer and size."""
    assert depth.dtype == torch.float32 and not depth.is_cuda and depth.dim() == 2 and depth.is_contiguous(), \
        'Depth frame should be a contiguous [H, W] float32 CPU tensor.'
    for name, mask in (('Depth', depth_mask), ('Feature', feature_mask)):
        if mask is not None:
            assert mask.dtype == torch.uint8 and not mask.is_cuda and mask.is_contiguous() and \
                tuple(mask.shape) == tuple(depth.shape), \
                f'{name} mask should be a contiguous uint8 CPU tensor of the depth frame\'s size.'
    assert t_w_c.is_cpu and t_w_c.dtype == torch.float32 and tuple(t_w_c.shape) == (4, 4), \
        't_w_c should be a 4x4 float32 CPU tensor.'
    assert intrinsics.is_cpu and intrinsics.dtype == torch.float32 and tuple(intrinsics.shape) == (3, 3), \
        'intrinsics should be a 3x3 float32 CPU tensor.'


def check_integrator_inputs(image: torch.Tensor,
                            t_w_c: torch.Tensor,
                            intrinsics: torch.Tensor,
                            image_type: str,
                            expected_dim: int,
                            expected_type: torch.dtype,
                            expected_num_channels: Optional[int] = None) -> None:
    """Input contract of the integrators (reference mapper.py:458-490): violations raise AssertionError."""
    assert image.dim() == expected_dim, f'{image_type} image should have dim == {expected_dim}.'
    assert image.is_cuda, f'{image_type} image should be on device.'
    assert image.dtype == expected_type, f'{image_type} image should have type {expected_type}.'
    assert intrinsics.is_cpu, f'{image_type} intrinsics should be on the CPU.'
    assert intrinsics.dtype == torch.float32, f'{image_type} intrinsics should have type torch.float32.'
    assert t_w_c.is_cpu, f'{image_type} t_w_c  should be on the CPU.'
    assert t_w_c.dtype == torch.float32, f'{image_type} pose should have type torch.float32.'
    assert tuple(t_w_c.shape) == (4, 4) and tuple(intrinsics.shape) == (3, 3), \
        f'{image_type}: pose must be 4x4 and intrinsics 3x3.'
    if expected_num_channels:
        assert image.shape[2] == expected_num_channels, \
            f'{image_type} should have {expected_num_channels} channels.'
