"""Tuning aid: the end-to-end (host frames) pass of bench.py alone, sparse fetch.  NVBX_FETCH_HINT=0/1 selects the
host-read instruction of k_pixel_fetch."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from nvblox_torch.constants import constants
    from nvblox_torch.mapper import Mapper
    constants.set_feature_array_num_elements(bench.C_FEAT)
    mp, _ = bench.mapper_params()
    mapper = Mapper(voxel_sizes_m=bench.VOXEL, mapper_parameters=mp, device=0)
    n = 48
    K, frames = bench.poses_and_depths(n)
    K_t = torch.from_numpy(K)
    poses = [torch.from_numpy(T) for T, _ in frames]
    h_depth = [torch.from_numpy(d).pin_memory() for _, d in frames]
    g = torch.Generator(device='cuda')
    h_feat = []
    for i in range(3):
        g.manual_seed(1000 + i)
        h_feat.append(torch.randn((bench.H, bench.W, bench.C_FEAT), generator=g, device='cuda').half().cpu().pin_memory())
    for rep in range(3):
        for i in range(4):
            mapper.integrate_frame_from_host(h_depth[i], h_feat[i % 3], poses[i], K_t)
        torch.cuda.synchronize()
        px0 = mapper.counters(0)['host_pixels_fetched']
        t0 = time.perf_counter()
        for i in range(n):
            mapper.integrate_frame_from_host(h_depth[i], h_feat[i % 3], poses[i], K_t)
            c = mapper.counters(0)
        secs = time.perf_counter() - t0
        px = (c['host_pixels_fetched'] - px0) / n
        print(f"hint={os.environ.get('NVBX_FETCH_HINT', '0')} rep {rep}: {n / secs:.1f} frames/s, {px:.0f} px/frame, "
              f"{px * 2 * bench.C_FEAT * n / secs / 1e9:.1f} GB/s over PCIe", flush=True)


if __name__ == '__main__':
    main()
