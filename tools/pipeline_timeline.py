"""Steady-state timeline of one frame's kernels in the PDL-chained pipeline (tuning aid, NOT a bench line).

Run with the profile build:  NVBX_PROFILE=1 python tools/pipeline_timeline.py [--frames 48] [--out FILE.md]

Every kernel stamps %globaltimer when its first CTA passes griddepcontrol.wait and when its last CTA ends
(csrc/nvbx_kernels.cuh, PROF_BEGIN / PROF_END).  The workload is bench.py's (cube-stacking replay, frames
enqueued back to back without a host sync), so the stamps show the overlap that isolated event timings hide:
where each kernel really starts relative to its predecessor's end, and how long the GPU idles between frames
(host-bound) -- plus the host's enqueue time per frame for comparison.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('NVBX_PROFILE', '1')

NAMES = ['raycast (post-wait: CTA 0 only once the marks are flushed early)', 'tsdf_update', 'trace_and_band', 'feature_geometry', 'feature_gather',
         'raycast (pre-wait: ray march + early flush, all CTAs)']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=48)
    ap.add_argument('--out', default=None)
    ap.add_argument('--pipelining', type=int, default=1, help='0 off, 1 on, 2 on + asynchronous enqueue')
    args = ap.parse_args()
    import torch
    import bench
    from nvblox_mindmap_b200 import _capi
    from nvblox_torch.constants import constants
    from nvblox_torch.mapper import Mapper
    lib = _capi.load()
    constants.set_feature_array_num_elements(bench.C_FEAT)
    mp, _ = bench.mapper_params()
    mapper = Mapper(voxel_sizes_m=bench.VOXEL, mapper_parameters=mp, device=0)
    mapper.set_pipelining(bool(args.pipelining), async_enqueue=(args.pipelining == 2))
    n_warm = 16
    n = n_warm + args.frames
    assert args.frames <= 56
    K, frames = bench.poses_and_depths(n)
    K_t = torch.from_numpy(K)
    poses = [torch.from_numpy(T) for T, _ in frames]
    depths = [torch.from_numpy(d).cuda() for _, d in frames]
    g = torch.Generator(device='cuda')
    feats = []
    for i in range(bench.N_FEATURE_BUFFERS):
        g.manual_seed(1000 + i)
        feats.append(torch.randn((bench.H, bench.W, bench.C_FEAT), generator=g, device='cuda').half())

    def step(i):
        mapper.add_depth_frame(depths[i], poses[i], K_t)
        mapper.add_feature_frame(feats[i % len(feats)], poses[i], K_t)

    for i in range(n_warm):
        step(i)
    mapper.reset_counters(0)          # frame numbers restart at 0
    buf = (C.c_uint64 * (64 * 8 * 2))()
    _capi.check(lib.nvbx_debug_profile_stamps(mapper._handle, None, 1))
    t0 = time.perf_counter()
    for i in range(n_warm, n):
        step(i)
    host_us = 1e6 * (time.perf_counter() - t0) / args.frames
    mapper.pipeline_join()
    _capi.check(lib.nvbx_debug_profile_stamps(mapper._handle, buf, 0))
    st = np.frombuffer(buf, dtype=np.uint64).reshape(64, 8, 2).astype(np.float64)

    lo, hi = 8, args.frames - 2      # steady state: skip the first frames after the sync and the drain
    rows = []
    origin = st[lo:hi, 0, 0]          # raycast post-wait begin of each frame
    for k in (5, 0, 1, 2, 3, 4):
        b = (st[lo:hi, k, 0] - origin) / 1e3
        e = (st[lo:hi, k, 1] - origin) / 1e3
        rows.append((NAMES[k], float(np.median(b)), float(np.median(e)), float(np.median(e - b))))
    period = float(np.median(np.diff(st[lo:hi, 0, 0]))) / 1e3
    gather_end_to_next = float(np.median(st[lo + 1:hi, 0, 0] - st[lo:hi - 1, 4, 1])) / 1e3
    lines = ['# Pipeline timeline (profile build, %globaltimer stamps; microseconds relative to the frame\'s raycast '
             'passing its wait)', '',
             f'workload: {bench.WORKLOAD}; {args.frames} frames enqueued back to back, medians over frames {lo}..{hi}',
             '', '| kernel | first CTA past wait | last CTA ended | span |', '|---|---:|---:|---:|']
    for name, b, e, d in rows:
        lines.append(f'| {name} | {b:.2f} | {e:.2f} | {d:.2f} |')
    lines += ['', f'frame period (raycast to raycast): {period:.2f} us  ->  {1e6 / period:.0f} frames/s',
              f'previous gather end -> this frame\'s raycast past wait: {gather_end_to_next:.2f} us',
              f'host enqueue per frame (Python + C ABI + 5 launches): {host_us:.2f} us']
    text = '\n'.join(lines)
    print(text)
    print(json.dumps({'rows': rows, 'period_us': period, 'host_enqueue_us': host_us}))
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, 'w') as f:
            f.write(text + '\n')


if __name__ == '__main__':
    main()
