"""Multi-GPU partitioning of the hot path: independent map replicas, no collective.

A single map does not shard (every stage is keyed by one spatial hash and reads neighbours); across maps
(environments / episodes in datagen and batched open-loop eval) the work is embarrassingly parallel.
Map `m` lives on rank `m mod world_size`; the only cross-rank traffic is the timing barrier and, when a
caller wants every point cloud in one place, a variable-length gather of the exported clouds.
"""
import ctypes as _C
from typing import List, Sequence

_F9 = _C.c_float * 9


def maps_of_rank(n_maps: int, world_size: int, rank: int) -> List[int]:
    """Map ids owned by `rank` (round-robin: map m -> rank m mod world_size)."""
    assert world_size >= 1 and 0 <= rank < world_size
    return list(range(rank, n_maps, world_size))


def owner_of_map(map_id: int, world_size: int) -> int:
    return map_id % world_size


def aggregate_throughput(units_per_rank: Sequence[float], seconds_per_rank: Sequence[float]) -> float:
    """Whole-job throughput: all units processed by all ranks over the SLOWEST rank's time."""
    return float(sum(units_per_rank)) / max(seconds_per_rank)


def gather_clouds(vertices, features, group=None):
    """Optional: gather every rank's exported (vertices [N_i,3], features [N_i,C]) on all ranks with one
    variable-length all-gather (works with gloo on CPU tensors and nccl on CUDA tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([vertices.shape[0]], dtype=torch.int64, device=vertices.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes) if sizes else 0
    pad_v = torch.zeros((nmax, 3), dtype=vertices.dtype, device=vertices.device)
    pad_f = torch.zeros((nmax, features.shape[1]), dtype=features.dtype, device=features.device)
    pad_v[:vertices.shape[0]] = vertices
    pad_f[:features.shape[0]] = features
    out_v = [torch.zeros_like(pad_v) for _ in range(world)]
    out_f = [torch.zeros_like(pad_f) for _ in range(world)]
    dist.all_gather(out_v, pad_v, group=group)
    dist.all_gather(out_f, pad_f, group=group)
    return [v[:s] for v, s in zip(out_v, sizes)], [f[:s] for f, s in zip(out_f, sizes)]


class MapBatch:
    """Several independent maps on ONE GPU (the maps of this rank in batched datagen, BASELINE configs[3]): one
    `nvblox_torch.Mapper` handle and one CUDA stream per map, one frame per map per `integrate_frames` call.

    A single map's frame is five latency-bound launches around one memory-bound one and leaves most of a B200
    idle; frames of different maps overlap on the device, and `nvbx_integrate_frames_batch` issues their launches
    from a pool of host threads so that the host's enqueue rate does not bound the step (reference equivalent: a
    Python loop over per-environment mappers, mindmap/run_isaaclab_datagen.py:194-235).  Results are exactly those
    of `Mapper.add_depth_frame` + `Mapper.add_feature_frame` per map.
    """

    def __init__(self, n_maps: int, voxel_size_m: float, mapper_parameters=None, device=None, host_threads: int = 0):
        import ctypes as C
        import torch
        from nvblox_mindmap_b200 import _capi
        from nvblox_mindmap_b200.params import NvbxFrameJob
        from nvblox_torch.mapper import Mapper
        self._torch = torch
        self._C = C
        self._capi = _capi
        self._lib = _capi.load()
        self.mappers = [Mapper(voxel_sizes_m=float(voxel_size_m), mapper_parameters=mapper_parameters, device=device)
                        for _ in range(n_maps)]
        self._device = self.mappers[0]._device
        self.streams = [torch.cuda.Stream(device=self._device) for _ in range(n_maps)]
        self._jobs = (NvbxFrameJob * n_maps)()
        self._host_threads = int(host_threads)
        for k, m in enumerate(self.mappers):
            self._jobs[k].mapper = m._handle.value
            self._jobs[k].map_id = 0
            self._jobs[k].stream = self.streams[k].cuda_stream
        self._keep = None

    def __len__(self):
        return len(self.mappers)

    def integrate_frames(self, depth_frames, feature_frames, poses, intrinsics, depth_masks=None, feature_masks=None):
        """One (depth, feature) frame per map: lists of CUDA tensors (feature_frames[k] may be None), CPU poses
        [4,4] and intrinsics [3,3] (one per map, or a single tensor shared by all).  Enqueues on each map's own
        stream; the frames' producers must be ordered before those streams (call `wait_for_current_stream()` after
        producing them on torch's current stream)."""
        t = self._torch
        n = len(self.mappers)
        assert len(depth_frames) == n and len(feature_frames) == n and len(poses) == n
        keep = []
        for k in range(n):
            j = self._jobs[k]
            d, f = depth_frames[k], feature_frames[k]
            K = intrinsics if isinstance(intrinsics, t.Tensor) else intrinsics[k]
            assert d.is_cuda and d.dtype == t.float32 and d.dim() == 2, 'Depth frame should be a 2-d float32 CUDA tensor.'
            # The kernels read the frames on the per-map streams: a `.contiguous()` copy made here would be written on
            # torch's current stream with nothing ordering the two, so non-contiguous inputs are rejected.
            assert d.is_contiguous(), 'MapBatch: depth frames must be contiguous.'
            j.height, j.width = int(d.shape[0]), int(d.shape[1])
            j.depth = d.data_ptr()
            dm = None if depth_masks is None else depth_masks[k]
            assert dm is None or dm.is_contiguous(), 'MapBatch: mask frames must be contiguous.'
            j.depth_mask = None if dm is None else self.mappers[k]._mask_ptr(dm, d)
            if f is not None:
                assert f.is_cuda and f.dtype == t.float16 and f.dim() == 3 and f.shape[:2] == d.shape, \
                    'Feature frame should be a [H, W, C] float16 CUDA tensor of the depth frame\'s size.'
                assert f.is_contiguous(), 'MapBatch: feature frames must be contiguous.'
                j.channels = int(f.shape[2])
                j.features = f.data_ptr()
                fm = None if feature_masks is None else feature_masks[k]
                assert fm is None or fm.is_contiguous(), 'MapBatch: mask frames must be contiguous.'
                j.feature_mask = None if fm is None else self.mappers[k]._mask_ptr(fm, f)
            else:
                j.channels, j.features, j.feature_mask = 0, None, None
            p = poses[k]
            assert (not p.is_cuda) and p.dtype == t.float32 and tuple(p.shape) == (4, 4), 'T_W_C should be a 4x4 CPU tensor.'
            pc = p if p.is_contiguous() else p.contiguous()      # bound to a name: alive while its bytes are read
            self._C.memmove(j.T_L_C, pc.data_ptr(), 64)
            assert (not K.is_cuda) and K.dtype == t.float32 and tuple(K.shape) == (3, 3), 'K should be a 3x3 CPU tensor.'
            Kc = K if K.is_contiguous() else K.contiguous()
            k9 = _F9.from_address(Kc.data_ptr())
            j.fx, j.fy, j.cx, j.cy = float(k9[0]), float(k9[4]), float(k9[2]), float(k9[5])
            keep.append((d, f, dm, None if f is None else fm))
        self._keep = keep          # alive until the next call (the launches read them asynchronously)
        self._capi.check(self._lib.nvbx_integrate_frames_batch(self._jobs, n, self._host_threads))

    def wait_for_current_stream(self):
        """Order every map stream after torch's current stream (inputs produced there become visible)."""
        cur = self._torch.cuda.current_stream(self._device)
        for s in self.streams:
            s.wait_stream(cur)

    def join_current_stream(self):
        """Order torch's current stream after every map stream (results become visible there)."""
        cur = self._torch.cuda.current_stream(self._device)
        for s in self.streams:
            cur.wait_stream(s)
