"""BASELINE configs[3] on ONE GPU: M independent episode maps (one Mapper handle + one CUDA stream each) fed
round-robin by one host thread, so that the latency-bound kernels of one map run under another map's
memory-bound gather.  Prints frames/s (all maps) for M = 1, 2, 4, 8.  Not the headline metric (that is M = 1)."""
import json
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
    n_warm, n_timed = 32, 256
    n = n_warm + n_timed
    K, frames = bench.poses_and_depths(n)
    K_t = torch.from_numpy(K)
    poses = [torch.from_numpy(T) for T, _ in frames]
    depths = [torch.from_numpy(d).cuda() for _, d in frames]
    g = torch.Generator(device='cuda')
    feats = []
    for i in range(bench.N_FEATURE_BUFFERS):
        g.manual_seed(1000 + i)
        feats.append(torch.randn((bench.H, bench.W, bench.C_FEAT), generator=g, device='cuda').half())
    torch.cuda.synchronize()
    rows = []
    for M in (1, 2, 4, 8):
        mappers = [Mapper(voxel_sizes_m=bench.VOXEL, mapper_parameters=mp, device=0) for _ in range(M)]
        streams = [torch.cuda.Stream() for _ in range(M)]

        def step(i):
            for k in range(M):       # map k replays the orbit with a phase shift: different views, same cost
                j = (i + 7 * k) % n
                with torch.cuda.stream(streams[k]):
                    mappers[k].add_depth_frame(depths[j], poses[j], K_t)
                    mappers[k].add_feature_frame(feats[(i + k) % len(feats)], poses[j], K_t)

        for i in range(n_warm):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for i in range(n_warm, n):
            step(i)
        host = time.perf_counter() - t0
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        upd = sum(m.counters(0)['feature_voxels_updated'] for m in mappers)
        rows.append({'maps': M, 'frames_per_s': M * n_timed / (ms / 1e3), 'us_per_frame': 1e3 * ms / (M * n_timed),
                     'host_enqueue_us_per_frame': 1e6 * host / (M * n_timed), 'feature_voxels_updated': upd})
        print(json.dumps(rows[-1]), flush=True)
        # the same M maps, one HOST THREAD per map (ctypes releases the GIL inside the C ABI, so the launches of
        # different maps are issued concurrently; only the thin Python wrappers serialise)
        import threading
        start = threading.Barrier(M + 1)

        def worker(k):
            torch.cuda.set_device(0)
            with torch.cuda.stream(streams[k]):
                start.wait()
                for i in range(n_warm, n):
                    j = (i + 7 * k) % n
                    mappers[k].add_depth_frame(depths[j], poses[j], K_t)
                    mappers[k].add_feature_frame(feats[(i + k) % len(feats)], poses[j], K_t)

        threads = [threading.Thread(target=worker, args=(k,)) for k in range(M)]
        for t in threads:
            t.start()
        torch.cuda.synchronize()
        e0.record()
        t0 = time.perf_counter()
        start.wait()
        for t in threads:
            t.join()
        host = time.perf_counter() - t0
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        rows.append({'maps': M, 'host_threads': M, 'frames_per_s': M * n_timed / (ms / 1e3),
                     'us_per_frame': 1e3 * ms / (M * n_timed), 'host_enqueue_us_per_frame': 1e6 * host / (M * n_timed)})
        print(json.dumps(rows[-1]), flush=True)
        del mappers
    return rows


if __name__ == '__main__':
    main()
