"""Tuning aid: where the HOST time of one frame goes (Python surface vs C ABI + launches).  Frames are enqueued
without synchronising; a large queue depth is avoided by syncing every 64 frames outside the timers."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from nvblox_mindmap_b200 import _capi
    from nvblox_torch.constants import constants
    from nvblox_torch.mapper import Mapper
    lib = _capi.load()
    constants.set_feature_array_num_elements(bench.C_FEAT)
    mp, _ = bench.mapper_params()
    mapper = Mapper(voxel_sizes_m=bench.VOXEL, mapper_parameters=mp, device=0)
    n = 64
    K, frames = bench.poses_and_depths(n)
    K_t = torch.from_numpy(K)
    poses = [torch.from_numpy(T) for T, _ in frames]
    depths = [torch.from_numpy(d).cuda() for _, d in frames]
    feat = torch.randn((bench.H, bench.W, bench.C_FEAT), device='cuda').half()
    for i in range(n):
        mapper.add_depth_frame(depths[i], poses[i], K_t)
        mapper.add_feature_frame(feat, poses[i], K_t)
    torch.cuda.synchronize()

    def timed(fn, reps=8):
        best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter()
            for i in range(n):
                fn(i)
            dt = time.perf_counter() - t0
            torch.cuda.synchronize()
            best = min(best, dt / n)
        return 1e6 * best

    t_py_d = timed(lambda i: mapper.add_depth_frame(depths[i], poses[i], K_t))
    t_py_f = timed(lambda i: mapper.add_feature_frame(feat, poses[i], K_t))
    # the same calls straight through ctypes with pre-marshalled arguments
    h, s = mapper._handle, mapper._stream()
    F16 = C.c_float * 16
    pose_c = [F16(*p.flatten().tolist()) for p in poses]
    fx, fy, cx, cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    dptr = [d.data_ptr() for d in depths]
    fptr = feat.data_ptr()
    t_c_d = timed(lambda i: lib.nvbx_integrate_depth(h, 0, dptr[i], bench.H, bench.W, None, pose_c[i], fx, fy, cx, cy, s))
    t_c_f = timed(lambda i: lib.nvbx_integrate_features(h, 0, fptr, bench.H, bench.W, bench.C_FEAT, None, pose_c[i],
                                                        fx, fy, cx, cy, s))
    t_noop = timed(lambda i: lib.nvbx_kernel_launch_count())
    print(f'host us per call (best of 8 x {n}, enqueue only): '
          f'add_depth_frame {t_py_d:.1f} (C ABI alone {t_c_d:.1f}), add_feature_frame {t_py_f:.1f} (C ABI alone {t_c_f:.1f}), '
          f'empty ctypes call {t_noop:.2f}')

    # The whole replay in ONE C call (jobs marshalled beforehand): the device's own frame period, with the host
    # enqueueing natively -- what the per-frame Python surface can at best approach.
    from nvblox_mindmap_b200.params import NvbxFrameJob
    n_seq = 512
    jobs = (NvbxFrameJob * n_seq)()
    for k in range(n_seq):
        i = k % n
        j = jobs[k]
        j.mapper, j.map_id, j.stream = h.value, 0, s
        j.height, j.width, j.channels = bench.H, bench.W, bench.C_FEAT
        j.depth, j.features = dptr[i], fptr
        C.memmove(j.T_L_C, poses[i].contiguous().data_ptr(), 64)
        j.fx, j.fy, j.cx, j.cy = fx, fy, cx, cy
    for pipe in (0, 1):
        mapper.set_pipelining(bool(pipe))
        best = (1e9, 0.0)
        for _ in range(4):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            t0 = time.perf_counter()
            rc = lib.nvbx_integrate_frames_batch(jobs, n_seq, 1)
            th = time.perf_counter() - t0
            assert rc == 0
            mapper.pipeline_join()
            b.record()
            torch.cuda.synchronize()
            best = min(best, (a.elapsed_time(b) * 1e3 / n_seq, th * 1e6 / n_seq))
        print(f'one C call for {n_seq} frames, pipelining {pipe}: device {best[0]:.2f} us per frame, host {best[1]:.2f} us per frame')


if __name__ == '__main__':
    main()
