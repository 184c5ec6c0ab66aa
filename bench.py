#!/usr/bin/env python3
"""bench.py -- feature frames integrated / s on the cube-stacking replay (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one camera frame: add_depth_frame + add_feature_frame
(mindmap's integrate_frame minus colour, nvblox_mapping_helpers.py:207-261) of a 512x512 depth +
768-channel fp16 feature frame into a 2 cm TSDF/feature map bounded by the cube-stacking workspace box,
camera on a 64-pose wrist orbit (moves > 1 mm / 0.1 deg per frame, so the viewpoint cache never hits).

  value     device-timed throughput with inputs resident in HBM (CUDA events on the launching stream).
  e2e       the same metric through the public API with HOST (pinned) frames: the H2D copy of depth +
            features and a D2H read of the map counters are inside the timed region of every step.
  roofline  the dominant kernel (k_feature_integrate): algorithmic bytes per launch (SURVEY 8(d) formula,
            N_upd from the device counters, distinct pixels per voxel from the oracle sample) over its
            live CUDA-event duration, against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline  the CPU oracle (a port of the reference algorithm, oracle/) timed on a bounded sample.

With N > 1 (torchrun, one process per GPU) every rank integrates its own independent map replica; there
is no collective on the data path, the timed region is bracketed by barriers and the
slowest rank's time is used.  `--impl reference` times the reference's CPU algorithm (oracle port, all
host threads) on rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import scenes as S  # noqa: E402

H = W = 512
C_FEAT = 768
VOXEL = 0.02
N_POSES = 64
N_FEATURE_BUFFERS = 6          # distinct 384 MiB feature frames cycled through (>> L2)
WORKLOAD = 'cube_stacking_replay: 1 wrist cam 512x512, C=768 fp16, 2 cm voxels, S-table scene, 64-pose orbit'


def bench_config(world):
    """The `config` object of the JSON line -- the SAME dict for --impl ours and --impl reference."""
    return {'workload': WORKLOAD, 'voxel_size_m': VOXEL, 'workspace': 'cube_stacking box',
            'maps_per_gpu': 1, 'parallelism': f'{world} independent map replica(s), no collective',
            'l2_policy': f'inputs larger than L2: {N_FEATURE_BUFFERS} distinct 384 MiB feature frames '
                         'cycled, a different one every step'}


def mapper_params():
    from tests.parity_utils import make_params
    return make_params(workspace=S.WS_CUBE_STACKING, max_dist=5.0, alpha=1.0, raycast_sub=1, decay=0.98)


def poses_and_depths(n):
    """(K, [(T_W_C, depth)] * n): the 64-pose orbit, repeated when n > 64 (rendered once per pose)."""
    K = S.intrinsics(W, H)
    uniq = []
    for i in range(min(n, N_POSES)):
        T = S.orbit_pose(i, N_POSES)
        uniq.append((T, S.render_depth(K, H, W, T, **S.S_TABLE)))
    return K, [uniq[i % N_POSES] for i in range(n)]


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region through NVML (every ~2 ms; the
    nvidia-smi CLI of the profiling recipe needs ~1 s to produce its first line, longer than the region)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.max_mhz = index, [], False, None
        self.armed = threading.Event()

    def run(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            bits = {'hw_slowdown': N.nvmlClocksEventReasonHwSlowdown,
                    'hw_thermal_slowdown': N.nvmlClocksEventReasonHwThermalSlowdown,
                    'sw_thermal_slowdown': N.nvmlClocksEventReasonSwThermalSlowdown,
                    'sw_power_cap': N.nvmlClocksEventReasonSwPowerCap}
            get_reasons = getattr(N, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
                N.nvmlDeviceGetCurrentClocksThrottleReasons
            self.armed.set()
            while not self.stop_flag:
                r = get_reasons(h)
                self.rows.append((float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)),
                                  [k for k, b in bits.items() if r & b]))
                time.sleep(0.002)
        except Exception as e:      # NVML missing: report no samples rather than fail the bench
            self.error = repr(e)
            self.armed.set()

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2.0)
        sm = [r[0] for r in self.rows]
        reasons = sorted({x for r in self.rows for x in r[1]})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': self.max_mhz,
                'reasons': reasons, 'samples': len(sm), 'source': 'NVML, 2 ms period, device-timed region only'}


def measured_peak_gbs():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def oracle_sample(n_frames, threads):
    """Time the CPU oracle on the first `n_frames` frames of the workload; also returns the per-voxel
    distinct-pixel ratio that enters the algorithmic-bytes formula."""
    from oracle import oracle as O
    O.set_threads(threads)
    _, op = mapper_params()
    K, frames = poses_and_depths(n_frames)
    m = O.OracleMapper(VOXEL, C_FEAT, op)
    t_total, upd, pix, cand = 0.0, 0, 0, 0
    for i, (T, depth) in enumerate(frames):
        feat = S.feature_frame(H, W, C_FEAT, 1000 + i)
        c0 = m.counters()
        t0 = time.perf_counter()
        m.add_depth_frame(depth, T, K)
        m.add_feature_frame(feat, T, K)
        t_total += time.perf_counter() - t0
        c = m.counters()
        upd += c['last_feature_voxels']
        pix += c['last_distinct_pixels']
        cand += c['feature_candidate_blocks'] - c0['feature_candidate_blocks']
    return {'frames': n_frames, 'seconds': t_total, 'fps': n_frames / t_total, 'n_upd': upd, 'u_px': pix,
            'n_cand': cand, 'threads': threads}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port) with all host threads, rank 0 only."""
    if rank != 0:
        return
    from oracle import oracle as O
    threads = len(os.sched_getaffinity(0))
    O.set_threads(threads)
    _, op = mapper_params()
    # bounded sample: the CPU path runs at a few frames/s; --warmup is honoured as given (up to 64 frames: a whole
    # orbit), the timed region is at most 128 steps so that the arm finishes within a few minutes
    n_warm, n_timed = min(args.warmup, 64), min(args.steps, 128)
    K, frames = poses_and_depths(n_warm + n_timed)
    m = O.OracleMapper(VOXEL, C_FEAT, op)
    feats = [S.feature_frame(H, W, C_FEAT, 1000 + i) for i in range(min(N_FEATURE_BUFFERS, len(frames)))]
    t_timed = 0.0
    for i, (T, depth) in enumerate(frames):
        t0 = time.perf_counter()
        m.add_depth_frame(depth, T, K)
        m.add_feature_frame(feats[i % len(feats)], T, K)
        if i >= n_warm:
            t_timed += time.perf_counter() - t0
    value = n_timed / t_timed
    line = {
        'impl': 'reference', 'metric': 'feature frames integrated/s (C=768, 512^2)', 'value': value,
        'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': n_timed, 'warmup': n_warm,
        'requested_steps': args.steps, 'requested_warmup': args.warmup,
        'ms_per_step': 1000.0 * t_timed / n_timed, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32 geometry + f16 features', 'data': 'synthetic',
        'config': bench_config(args.gpus),
        'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{n_timed} frames of the workload after {n_warm} warm-up frames (bounded); the '
                                   'reference itself cannot be built offline (Eigen/stdgpu/glog absent), so this is '
                                   'the oracle port of its algorithm with OpenMP over blocks'},
        'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import ctypes as C
    from nvblox_mindmap_b200 import _capi
    from nvblox_mindmap_b200.params import NvbxCounters
    from nvblox_torch.constants import constants
    from nvblox_torch.mapper import Mapper

    torch.cuda.set_device(local_rank)
    dev = f'cuda:{local_rank}'
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(dev))
    lib = _capi.load()
    constants.set_feature_array_num_elements(C_FEAT)
    mp, _ = mapper_params()
    mapper = Mapper(voxel_sizes_m=VOXEL, mapper_parameters=mp, device=local_rank)
    # replay: every frame of the sequence stays resident and unmodified, which is the contract of frame pipelining
    # (the gather of frame i on the map's own stream, the depth path of frame i + 1 underneath it; same results)
    mapper.set_pipelining(bool(args.pipelining), async_enqueue=(args.pipelining == 2))

    n_total = args.warmup + args.steps
    K, frames = poses_and_depths(max(n_total, 1))
    K_t = torch.from_numpy(K)
    poses = [torch.from_numpy(T) for T, _ in frames]
    depths = [torch.from_numpy(d).to(dev) for _, d in frames]
    g = torch.Generator(device=dev)
    feats = []
    for i in range(N_FEATURE_BUFFERS):       # synthetic N(0,1) features, per-map seed = rank
        g.manual_seed(1000 + i + 7919 * rank)
        feats.append(torch.randn((H, W, C_FEAT), generator=g, device=dev, dtype=torch.float32).half())
    torch.cuda.synchronize()

    def step(i):
        mapper.add_depth_frame(depths[i], poses[i], K_t)
        mapper.add_feature_frame(feats[i % N_FEATURE_BUFFERS], poses[i], K_t)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-timed pass: inputs resident in HBM ------------------------------------------------
    for i in range(args.warmup):
        step(i)
    mapper.reset_counters(0)
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.armed.wait(5.0)
    sampler.rows.clear()
    launches0 = int(lib.nvbx_kernel_launch_count())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()

    def pipe_waits():
        a, b = C.c_int64(0), C.c_int64(0)
        lib.nvbx_pipeline_wait_stats(C.byref(a), C.byref(b))
        return a.value, b.value

    w0 = pipe_waits()
    e0.record()
    t_host = time.perf_counter()
    for i in range(args.warmup, n_total):
        step(i)
    t_host = time.perf_counter() - t_host      # host time to ENQUEUE the steps (no sync inside)
    w1 = pipe_waits()
    # how long the host sat waiting for a ring slot (device-bound) -- the rest of t_host is its own work
    host_wait = {'waits_per_step': (w1[0] - w0[0]) / args.steps, 'wait_us_per_step': (w1[1] - w0[1]) / 1e3 / args.steps}
    mapper.pipeline_join()                     # the timed region ends when the last gather has finished
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = int(lib.nvbx_kernel_launch_count()) - launches0
    counters = mapper.counters(0)
    clocks = sampler.finish()
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.steps / (ms / 1000.0)

    if args.sweep_gather:
        if rank == 0:
            sweep_gather(args, lib, mapper, depths, poses, feats, K_t, step)
        return

    def sequence_pass(seq):
        """The same frames through Mapper.integrate_frames (one Python -> C crossing per `seq` frames)."""
        idx = list(range(args.warmup, n_total))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        t0 = time.perf_counter()
        for k in range(0, len(idx), seq):
            ii = idx[k:k + seq]
            mapper.integrate_frames([depths[i] for i in ii], [feats[i % N_FEATURE_BUFFERS] for i in ii],
                                    [poses[i] for i in ii], K_t)
        th = time.perf_counter() - t0
        mapper.pipeline_join()
        b.record()
        torch.cuda.synchronize()
        return {'frames_per_call': seq, 'frames_per_s': len(idx) / (a.elapsed_time(b) / 1000.0),
                'ms_per_step': a.elapsed_time(b) / len(idx), 'host_enqueue_ms_per_step': 1000.0 * th / len(idx)}

    if args.quick:
        if rank == 0:
            print(json.dumps({'sequence_api': [sequence_pass(q) for q in (1, 8, 32)]}), flush=True)
            th = [0.0, 0.0]
            for i in range(args.warmup, n_total):
                t0 = time.perf_counter()
                mapper.add_depth_frame(depths[i], poses[i], K_t)
                t1 = time.perf_counter()
                mapper.add_feature_frame(feats[i % N_FEATURE_BUFFERS], poses[i], K_t)
                th[0] += t1 - t0
                th[1] += time.perf_counter() - t1
            torch.cuda.synchronize()
            print(json.dumps({'host_us_depth_call': 1e6 * th[0] / args.steps, 'host_us_feature_call': 1e6 * th[1] / args.steps}),
                  flush=True)
            kms = kernel_time_ms(args, mapper, depths, poses, feats, K_t)
            print(json.dumps({'per_kernel_us': per_kernel_us(args, mapper, depths, poses, feats, K_t)}), flush=True)
            print(json.dumps({'quick': True, 'value': value, 'ms_per_step': ms / args.steps, 'kernel_ms': kms,
                              'host_enqueue_ms_per_step': 1000.0 * t_host / args.steps, 'host_wait': host_wait,
                              'gpu_launches': launches, 'profile': counters.get('profile')}), flush=True)
        return

    # ---- per-kernel time of the dominant kernel, live, with CUDA events around each call ---------------
    # (separate pass so that the event records do not perturb the number above)
    feat_ms, ev = [], [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                       for _ in range(args.steps)]
    for j, i in enumerate(range(args.warmup, n_total)):
        mapper.add_depth_frame(depths[i], poses[i], K_t)
        ev[j][0].record()
        mapper.add_feature_frame(feats[i % N_FEATURE_BUFFERS], poses[i], K_t)
        ev[j][1].record()
    torch.cuda.synchronize()
    feat_calls = np.array([a.elapsed_time(b) for a, b in ev])
    feat_call_ms = float(np.mean(feat_calls))
    feat_call_pct = [float(np.percentile(feat_calls, q)) for q in (10, 50, 90)]   # median, p10 / p90

    # ---- end-to-end pass: HOST frames through the public API ----------------------------------------------
    # Headline e2e = the default host entry: depth copied H2D, the pinned feature frame read through its device
    # mapping so that only the pixels the frame's voxels sample cross PCIe (nvbx_integrate_frame_host, sparse
    # mode; identical map, tests/test_gpu_host_frames.py).  h2d_bytes_per_step is what actually crossed the bus:
    # depth bytes + device-counted fetched pixels x 2C.  The whole-frame copy is timed next to it (`dense`).
    n_e2e = max(1, min(args.steps, 64))
    h_depth = [d.cpu().pin_memory() for d in depths[:n_e2e]]
    h_feat = [feats[i % N_FEATURE_BUFFERS].cpu().pin_memory() for i in range(min(n_e2e, 3))]

    def e2e_pass(mode, n):
        mapper.set_host_fetch_mode(mode)
        for i in range(min(3, n)):
            mapper.integrate_frame_from_host(h_depth[i], h_feat[i % len(h_feat)], poses[i], K_t)
        px0 = mapper.counters(0)['host_pixels_fetched']
        barrier()
        t0 = time.perf_counter()
        for i in range(n):
            mapper.integrate_frame_from_host(h_depth[i], h_feat[i % len(h_feat)], poses[i], K_t)
            c = mapper.counters(0)          # D2H read of the step's result (map counters) + stream sync
        barrier()
        secs = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([secs], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        return world * n / secs, (c['host_pixels_fetched'] - px0) / n

    dense_value, _ = e2e_pass('dense', min(n_e2e, 16))
    e2e_value, e2e_px = e2e_pass('sparse', n_e2e)
    d2h = C.sizeof(NvbxCounters)
    e2e_h2d = H * W * 4 + int(round(e2e_px * 2 * C_FEAT))

    # ---- end to end from the frame mindmap's extractor really produces: the backbone's
    # [1, 768, 32, 32] fp32 map crosses PCIe (3 MiB), the 32^2 -> 512^2 up-sampling is fused into the gather -------
    g_low = torch.Generator(device=dev)
    h_low = []
    for i in range(3):
        g_low.manual_seed(5000 + i)
        hwc = torch.randn((1, H // 16, W // 16, C_FEAT), generator=g_low, device=dev, dtype=torch.float32)
        h_low.append(hwc.cpu().pin_memory().permute(0, 3, 1, 2))      # [1, c, h, w] view of pinned HWC memory

    def e2e_lowres_pass(n):
        for i in range(min(3, n)):
            mapper.integrate_frame_from_host_lowres(h_depth[i], h_low[i % len(h_low)], poses[i], K_t)
        barrier()
        t0 = time.perf_counter()
        for i in range(n):
            mapper.integrate_frame_from_host_lowres(h_depth[i], h_low[i % len(h_low)], poses[i], K_t)
            mapper.counters(0)
        barrier()
        secs = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([secs], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        return world * n / secs

    e2e_lowres_value = e2e_lowres_pass(n_e2e)
    e2e_lowres_h2d = H * W * 4 + (H // 16) * (W // 16) * C_FEAT * 4

    # ---- BASELINE configs[3]: 64 maps sharded over the ranks (every rank takes part) ---------------------------------
    batched64 = batched_64_maps_stage(depths, poses, feats, K_t, local_rank, world, dist)

    # Everything below is rank 0's own work (CPU oracle sample, extra legs): the other ranks leave now instead of
    # spinning in a collective while rank 0 computes.
    if world > 1 and rank != 0:
        dist.destroy_process_group()
        return

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        n_cores = len(os.sched_getaffinity(0))
        sample = oracle_sample(4, 1)                    # ~6 s of single-thread CPU work
        sample_mt = oracle_sample(24, n_cores)          # ~4-20 s with every host thread
        n_upd = counters['feature_voxels_updated'] / args.steps
        px_per_voxel = sample_mt['u_px'] / max(1, sample_mt['n_upd'])
        px_per_voxel_dev = e2e_px / max(1.0, n_upd)   # device-counted on the sparse host pass (same poses)
        # Algorithmic bytes of the dominant kernel (k_feature_gather) per launch, alpha = 1:
        # 2C bytes per DISTINCT feature pixel read + 2(C+1) bytes per voxel row written.  N_upd comes from the
        # device counters of the timed region, distinct pixels per voxel from the oracle sample.
        b_feat = 2 * C_FEAT * px_per_voxel * n_upd + 2 * (C_FEAT + 1) * n_upd
        kms = kernel_time_ms(args, mapper, depths, poses, feats, K_t)
        achieved = b_feat / (kms * 1e-3) / 1e9 if kms else None
        traffic, traffic_src = ncu_traffic_bytes()
        sub = argparse.Namespace(steps=min(args.steps, 128), warmup=args.warmup)
        per_kernel = per_kernel_us(sub, mapper, depths, poses, feats, K_t)
        export = export_stage(mapper)
        fused = fused_upsample_stage(sub, mapper, depths, poses, K_t, dev)
        mapper.set_pipelining(False)
        kms_alone = kernel_time_ms(args, mapper, depths, poses, feats, K_t)   # the same kernel with nothing beside it
        # the same frames without frame pipelining (what a caller that frees / overwrites its frames right away gets)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_np = min(args.steps, 256)
        torch.cuda.synchronize()
        ea.record()
        for i in range(args.warmup, args.warmup + n_np):
            step(i)
        eb.record()
        torch.cuda.synchronize()
        value_no_pipe = n_np / (ea.elapsed_time(eb) / 1e3)
        # ... and with pipelining but synchronous enqueue (the calling Python thread issues the CUDA work itself)
        mapper.set_pipelining(True, async_enqueue=False)
        for i in range(args.warmup):
            step(i)
        mapper.pipeline_join()
        torch.cuda.synchronize()
        ea.record()
        for i in range(args.warmup, args.warmup + n_np):
            step(i)
        mapper.pipeline_join()
        eb.record()
        torch.cuda.synchronize()
        value_sync_enqueue = n_np / (ea.elapsed_time(eb) / 1e3)
        mapper.set_pipelining(bool(args.pipelining), async_enqueue=(args.pipelining == 2))
        batched = batched_maps_stage(depths, poses, feats, K_t, local_rank)
        drill = drill_in_box_stage(lib, feats, h_feat, local_rank, peak)
        cold = cold_start_stage(depths, poses, feats, K_t, local_rank)
        del mapper
        torch.cuda.empty_cache()
        stress = stress_stage(local_rank)
        line = {
            'metric': 'feature frames integrated/s (C=768, 512^2)', 'value': value, 'unit': 'frames/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 geometry + f16 features', 'data': 'synthetic',
            'config': bench_config(world),
            'e2e': {'value': e2e_value, 'unit': 'frames/s',
                    'h2d_bytes_per_step': e2e_h2d, 'd2h_bytes_per_step': d2h, 'steps': n_e2e,
                    'transfer': 'depth by cudaMemcpyAsync; pinned feature frame read sparsely through its device '
                                f'mapping ({e2e_px:.0f} of {H * W} pixels per step cross PCIe, each once)',
                    'dense_copy': {'value': dense_value, 'unit': 'frames/s',
                                   'h2d_bytes_per_step': H * W * 4 + H * W * C_FEAT * 2},
                    # what the host's memory system has to deliver for this figure (all ranks read pinned host memory
                    # through the same root complex / NUMA node: the aggregate is what saturates at N = 4 .. 8)
                    'host_read_GBps_aggregate': e2e_value * e2e_h2d / 1e9,
                    'host_read_GBps_per_rank': e2e_value * e2e_h2d / 1e9 / world},
            'gpu_launches': launches,
            'clocks': clocks,
            'e2e_lowres': {'value': e2e_lowres_value, 'unit': 'frames/s', 'h2d_bytes_per_step': e2e_lowres_h2d,
                           'd2h_bytes_per_step': d2h, 'steps': n_e2e,
                           'note': 'the frame mindmap\'s extractor produces: depth + the backbone\'s [1, 768, 32, 32] fp32 '
                                   'map from pinned host memory, up-sampling fused into the gather '
                                   '(Mapper.integrate_frame_from_host_lowres; bit-identical map, tests/test_gpu_upsample.py)'},
            'roofline': {'bound': 'hbm',
                         'kernel': 'k_feature_gather_dyn<3,256,5> (static deal, item prefetch), on its own stream with '
                                   'the next frame\'s depth path underneath (3 CTAs/SM)' if args.pipelining else
                                   'k_feature_gather_dyn<3,256,5> on 4 CTAs/SM (static deal, item prefetch)',
                         'achieved': achieved, 'peak': peak,
                         'unit': 'GB/s', 'frac': (achieved / peak) if achieved else None, 'traffic': traffic,
                         'traffic_source': traffic_src,
                         'traffic_kind': 'static: parsed from the committed ncu summary under profiles/, NOT measured in this run',
                         'alone': {'kernel_ms': kms_alone,
                                   'achieved': (b_feat / (kms_alone * 1e-3) / 1e9) if kms_alone else None,
                                   'frac': (b_feat / (kms_alone * 1e-3) / 1e9 / peak) if kms_alone else None,
                                   'note': 'the same launch with frame pipelining off (nothing runs beside it, 4 CTAs/SM)'},
                         'whole_frame': {'algorithmic_bytes': b_feat + 4 * H * W + 16 * counters['tsdf_voxels_updated'] / args.steps
                                         + 4096 * counters['feature_candidate_blocks'] / args.steps,
                                         'GBps': (b_feat + 4 * H * W + 16 * counters['tsdf_voxels_updated'] / args.steps
                                                  + 4096 * counters['feature_candidate_blocks'] / args.steps) / (ms / args.steps * 1e-3) / 1e9,
                                         'note': 'all five kernels of a frame over ms_per_step'},
                         'peak_source': peak_src, 'kernel_ms': kms, 'algorithmic_bytes_per_launch': b_feat,
                         'n_upd_per_frame': n_upd, 'distinct_pixels_per_voxel': px_per_voxel,
                         'distinct_pixels_per_voxel_device_counted': px_per_voxel_dev,
                         'frac_of_nominal_8TBps': (achieved / 8000.0) if achieved else None},
            'cpu_baseline': {'value': sample_mt['fps'], 'unit': 'frames/s', 'cores': n_cores, 'kind': 'port',
                             'sample': f"first {sample_mt['frames']} frames of the workload through the CPU oracle "
                                       f"(OpenMP over blocks, {n_cores} threads, {sample_mt['seconds']:.1f} s); "
                                       f"1 thread, first {sample['frames']} frames: {sample['fps']:.3f} frames/s"},
            'extra': {'feature_call_ms': feat_call_ms, 'feature_call_ms_p10_p50_p90': feat_call_pct, 'host_enqueue_ms_per_step': 1000.0 * t_host / args.steps, 'host_wait_for_ring_slot': host_wait,
                      'per_kernel_us_in_pipeline': per_kernel, 'export_stage': export,
                      'fused_upsample': fused, 'batched_maps_one_gpu': batched, 'drill_in_box': drill,
                      'batched_64_maps': batched64, 'stress': stress, 'cold_start': cold,
                      'pipelining': {'mode': int(args.pipelining),
                                     'frames_per_s_without': value_no_pipe,
                                     'frames_per_s_synchronous_enqueue': value_sync_enqueue,
                                     'host_wait_for_ring_slot': host_wait,
                                     'note': 'value is measured with Mapper.set_pipelining(True, async_enqueue=True) '
                                             '(mode 2): every frame of the replay stays resident and unmodified, so '
                                             'geometry + gather of frame i run on the map\'s own streams above the depth '
                                             'path of frame i + 1, and the CUDA calls are issued by the mapper\'s worker '
                                             'thread (bit-identical map); frames_per_s_synchronous_enqueue = mode 1 (the '
                                             'Python thread issues them itself: host-bound); frames_per_s_without = the '
                                             'default mode of the drop-in (frames may be freed right after the call)'},
                      'counters_per_step': {k: v / args.steps for k, v in counters.items()
                                            if isinstance(v, (int, float))}},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def kernel_time_ms(args, mapper, depths, poses, feats, K_t):
    """Average duration of the k_feature_integrate launch, measured live with CUDA events placed by the
    library around that kernel (nvbx_set_kernel_timing)."""
    import ctypes as C
    from nvblox_mindmap_b200 import _capi
    lib = _capi.load()
    if not hasattr(lib, 'nvbx_set_kernel_timing'):
        return None
    lib.nvbx_set_kernel_timing.argtypes = [C.c_void_p, C.c_int]
    lib.nvbx_get_kernel_timing.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.nvbx_set_kernel_timing(mapper._handle, 1)
    n_total = args.warmup + args.steps
    for i in range(args.warmup, n_total):
        mapper.add_depth_frame(depths[i], poses[i], K_t)
        mapper.add_feature_frame(feats[i % N_FEATURE_BUFFERS], poses[i], K_t)
    ms, n = C.c_double(), C.c_int64()
    lib.nvbx_get_kernel_timing(mapper._handle, 0, C.byref(ms), C.byref(n))
    lib.nvbx_set_kernel_timing(mapper._handle, 0)
    return (ms.value / n.value) if n.value else None


def sweep_gather(args, lib, mapper, depths, poses, feats, K_t, step):
    """Tuning aid (not a bench line): every schedule of the feature gather in ONE process, on the bench
    workload -- live kernel duration (library-placed CUDA events) and whole-frame device time."""
    import torch
    n_total = args.warmup + args.steps
    grid = [(0, 0, 1), (4, 0, 1), (6, 0, 1), (7, 0, 1), (7, -1, 1), (9, 0, 1), (10, 0, 1)]
    rows = []
    for rep in range(2):
        for (v, pm, tk) in grid:
            assert lib.nvbx_set_gather_tuning(v, pm, tk) == 0
            for i in range(8):
                step(i)
            kms = kernel_time_ms(args, mapper, depths, poses, feats, K_t)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for i in range(args.warmup, n_total):
                step(i)
            e1.record()
            torch.cuda.synchronize()
            rows.append({'rep': rep, 'variant': v, 'dyn_permille': pm, 'ticket': tk, 'gather_us': 1000.0 * kms,
                         'frame_us': 1000.0 * e0.elapsed_time(e1) / args.steps})
            print(json.dumps(rows[-1]), flush=True)
    h_feat = [feats[i].cpu().pin_memory() for i in range(2)]
    peak, _ = measured_peak_gbs()
    for (v, pm, tk) in [(7, 0, 1), (7, -1, 1), (10, 0, 1)]:
        assert lib.nvbx_set_gather_tuning(v, pm, tk) == 0
        d = drill_in_box_stage(lib, feats, h_feat, 0, peak)
        print(json.dumps({'drill_in_box': True, 'variant': v, 'dyn_permille': pm, 'frames_per_s': d['frames_per_s'],
                          'gather_us': 1000.0 * d['gather_kernel_ms'], 'gather_GBps': d['gather_GBps']}), flush=True)
    lib.nvbx_set_gather_tuning(7, 0, 4)


def drill_in_box_stage(lib, feats, h_feat, local_rank, peak):
    """BASELINE configs[2] as an extra data point (not the headline): head (static: viewpoint-cache hits) + wrist
    camera per step, 512x512, C = 768, 1 cm voxels, drill-in-box workspace.  ~4x the voxels per frame of the
    headline workload, i.e. what the gather reaches when a launch is long enough to amortise its ramp."""
    import ctypes as C
    import torch
    from nvblox_torch.mapper import Mapper
    from tests.parity_utils import make_params
    mp, _ = make_params(workspace=S.WS_DRILL_IN_BOX, decay=0.999)
    mapper = Mapper(voxel_sizes_m=0.01, mapper_parameters=mp, device=local_rank)
    dev = f'cuda:{local_rank}'
    K = S.intrinsics(W, H)
    K_t = torch.from_numpy(K)
    n_pose, n_warm, n_timed = 16, 8, 32
    T_head = S.look_at((-0.2, 0.0, 0.6), (0.4, 0.0, 0.05))
    cams = [(torch.from_numpy(T_head), torch.from_numpy(S.render_depth(K, H, W, T_head, **S.S_TABLE)))]
    for i in range(n_pose):
        T = S.orbit_pose(i, 64, 0.4, 0.45)
        cams.append((torch.from_numpy(T), torch.from_numpy(S.render_depth(K, H, W, T, **S.S_TABLE))))
    d_dev = [d.to(dev) for _, d in cams]
    d_pin = [d.pin_memory() for _, d in cams]

    def step(i, host=False):
        for c in (0, 1 + i % n_pose):
            f = (i + c) % len(h_feat if host else feats)
            if host:
                mapper.integrate_frame_from_host(d_pin[c], h_feat[f], cams[c][0], K_t)
            else:
                mapper.add_depth_frame(d_dev[c], cams[c][0], K_t)
                mapper.add_feature_frame(feats[f], cams[c][0], K_t)

    for i in range(n_warm):
        step(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    mapper.reset_counters(0)
    e0.record()
    for i in range(n_warm, n_warm + n_timed):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n_upd = mapper.counters(0)['feature_voxels_updated'] / (2 * n_timed)
    # the same steps with frame pipelining (frames resident: the gather of one camera frame under the depth path of
    # the next)
    mapper.set_pipelining(True)
    for i in range(n_warm, n_warm + 4):
        step(i)
    mapper.pipeline_join()
    torch.cuda.synchronize()
    e0.record()
    for i in range(n_warm, n_warm + n_timed):
        step(i)
    mapper.pipeline_join()
    e1.record()
    torch.cuda.synchronize()
    ms_pipe = e0.elapsed_time(e1)
    mapper.set_pipelining(False)
    lib.nvbx_set_kernel_timing(mapper._handle, 1)
    for i in range(n_warm, n_warm + n_timed):
        step(i)
    kms, n = C.c_double(), C.c_int64()
    lib.nvbx_get_kernel_timing(mapper._handle, 0, C.byref(kms), C.byref(n))
    lib.nvbx_set_kernel_timing(mapper._handle, 0)
    kms = kms.value / max(1, n.value)
    px0 = mapper.counters(0)['host_pixels_fetched']
    for i in range(n_warm, n_warm + 4):          # distinct pixels per frame, device-counted by the sparse host path
        step(i, host=True)
    px = (mapper.counters(0)['host_pixels_fetched'] - px0) / 8
    b = 2 * C_FEAT * px + 2 * (C_FEAT + 1) * n_upd
    gbs = b / (kms * 1e-3) / 1e9 if kms else None
    return {'workload': 'drill_in_box: head (static) + wrist cam 512x512, C=768, 1 cm voxels', 'frames_per_s':
            2 * n_timed / (ms / 1e3), 'ms_per_camera_frame': ms / (2 * n_timed),
            'frames_per_s_pipelined': 2 * n_timed / (ms_pipe / 1e3), 'n_upd_per_frame': n_upd,
            'distinct_pixels_per_frame': px, 'gather_kernel_ms': kms, 'gather_algorithmic_bytes': b,
            'gather_GBps': gbs, 'gather_frac_of_measured_peak': (gbs / peak) if gbs else None,
            'feature_blocks': mapper.feature_layer_view(0).num_blocks(),
            'note': 'extra data point, same kernels and event-timing method as the headline roofline'}


def batched_maps_stage(depths, poses, feats, K_t, local_rank, n_maps=8, n_timed=96):
    """BASELINE configs[3] on one GPU: n_maps independent episode maps (what one rank of the 64-map datagen job
    holds at 8 GPUs), one frame per map per step through MapBatch -> nvbx_integrate_frames_batch: one stream per
    map, launches issued by a pool of host threads, so that the latency-bound kernels of one map run under another
    map's gather.  The single-thread Python loop over the same maps is timed beside it."""
    import torch
    from nvblox_mindmap_b200.replicas import MapBatch
    mp, _ = mapper_params()
    batch = MapBatch(n_maps, VOXEL, mp, device=local_rank)
    n = len(depths)

    def frames(i):
        idx = [(i + 7 * k) % n for k in range(n_maps)]
        return ([depths[j] for j in idx], [feats[(i + k) % len(feats)] for k in range(n_maps)],
                [poses[j] for j in idx])

    def timed(step_fn):
        for i in range(8):
            step_fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        batch.wait_for_current_stream()
        t0 = time.perf_counter()
        for i in range(8, 8 + n_timed):
            step_fn(i)
        host = time.perf_counter() - t0
        batch.join_current_stream()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        return n_maps * n_timed / (ms / 1e3), 1e6 * host / (n_maps * n_timed)

    def step_batch(i):
        d, f, p = frames(i)
        batch.integrate_frames(d, f, p, K_t)

    def step_loop(i):
        d, f, p = frames(i)
        for k in range(n_maps):
            with torch.cuda.stream(batch.streams[k]):
                batch.mappers[k].add_depth_frame(d[k], p[k], K_t)
                batch.mappers[k].add_feature_frame(f[k], p[k], K_t)

    fps_loop, host_loop = timed(step_loop)
    fps_batch, host_batch = timed(step_batch)
    return {'maps': n_maps, 'frames_per_s': fps_batch, 'host_us_per_frame': host_batch,
            'python_loop': {'frames_per_s': fps_loop, 'host_us_per_frame': host_loop},
            'note': 'all maps on one GPU, one stream each; nvbx_integrate_frames_batch (host launch pool) vs a '
                    'single-thread Python loop over the same Mapper handles'}


def batched_64_maps_stage(depths, poses, feats, K_t, local_rank, world, dist, n_total_maps=64, n_timed=16):
    """BASELINE configs[3]: 64 independent episode maps sharded over the ranks (64 / world per rank, one rank per
    GPU, no collective on the data path), each map integrating the cube-stacking sequence from its own phase of the
    orbit, one frame per map per step through MapBatch -> nvbx_integrate_frames_batch.  Every rank runs its share;
    the job's time is the slowest rank's (device-timed, barrier on both sides)."""
    import torch
    from nvblox_mindmap_b200.replicas import MapBatch
    n_maps = max(1, n_total_maps // world)
    mp, _ = mapper_params()
    batch = MapBatch(n_maps, VOXEL, mp, device=local_rank)
    n = len(depths)

    def step(i):
        idx = [(i + 5 * k) % n for k in range(n_maps)]          # per-map phase of the orbit
        batch.integrate_frames([depths[j] for j in idx], [feats[(i + k) % len(feats)] for k in range(n_maps)],
                               [poses[j] for j in idx], K_t)

    # Warm-up = one earlier episode per map: a whole orbit grows every map's arenas to the workspace's block count,
    # then clear() starts the timed episode from an EMPTY map with the arenas kept -- what datagen does between
    # episodes (run_isaaclab_datagen.py:194-196).  Block allocation of the new episode is inside the timed region.
    for i in range(n):
        step(i)
    for mpr in batch.mappers:
        mpr.clear()
        mpr.reset_counters(0)
    batch.join_current_stream()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    batch.wait_for_current_stream()
    for i in range(n_timed):
        step(i)
    batch.join_current_stream()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    upd = sum(m.counters(0)['feature_voxels_updated'] for m in batch.mappers)
    if world > 1:
        t = torch.tensor([ms], device=f'cuda:{local_rank}', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        u = torch.tensor([float(upd)], device=f'cuda:{local_rank}', dtype=torch.float64)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        upd = float(u.item())
    total_frames = n_maps * world * n_timed
    out = {'maps_total': n_maps * world, 'maps_per_rank': n_maps, 'ranks': world, 'steps_per_map': n_timed,
           'frames': total_frames, 'frames_per_s': total_frames / (ms / 1e3), 'ms_slowest_rank': ms,
           'feature_voxels_updated_total': upd,
           'note': 'BASELINE configs[3]: independent maps, one frame per map per step (MapBatch, one stream per map, '
                   'host launch pool); every map starts the timed episode empty (clear() after a warm-up episode that '
                   'sized its arenas); feature frames are shared read-only inputs, poses are per-map'}
    del batch
    torch.cuda.empty_cache()
    return out


def stress_stage(local_rank, target_blocks=2_000_000, n_steps=16, keep=False):
    """BASELINE configs[4]: 4 cameras 1024x1024, 1024-channel features, 1 cm voxels in a 20 x 20 x 4 m workspace whose
    map is first populated to >= 2 M TSDF blocks (8 GB) by a depth-only fly-through (cameras on a grid looking along
    the four horizontal directions at a wall beyond the integration distance: every block of each frustum is
    observed free space), then `n_steps` timed steps of the 4-camera rig over the table scene (depth + features per
    camera, decay per step, as mindmap runs it) and ONE full export (feature mesh of every dirty block ->
    vertices + features)."""
    import torch
    from nvblox_torch.constants import constants
    from nvblox_torch.mapper import Mapper
    from tests.parity_utils import make_params
    from tests.test_gpu_stress import SCENE, WS_STRESS, rig_poses
    if torch.cuda.mem_get_info(local_rank)[0] < 110 * 2 ** 30:
        return {'skipped': 'needs 110 GiB of free HBM'}
    C_S, HS = 1024, 1024
    dev = f'cuda:{local_rank}'
    constants.set_feature_array_num_elements(C_S)
    mp, _ = make_params(workspace=WS_STRESS, max_dist=5.0, alpha=1.0, raycast_sub=1, decay=0.999)
    m = Mapper(voxel_sizes_m=0.01, mapper_parameters=mp, device=local_rank)
    K = S.intrinsics(HS, HS)
    K_t = torch.from_numpy(K)
    t_pop = time.perf_counter()
    far = torch.full((HS, HS), 6.0, device=dev, dtype=torch.float32)     # beyond the 5 m integration distance
    layer = m.tsdf_layer_view(0)
    n_pop, n_blocks = 0, 0
    spots = [(x, y) for r in (5.0, 0.0, 8.0) for x in (-r, r) for y in (-r, r)] + [(5.0, 0.0), (-5.0, 0.0), (0.0, 5.0), (0.0, -5.0)]
    for (x, y) in spots:
        for dx, dy in ((1, 0), (0, 1), (-1, 0), (0, -1)):
            eye = np.array([x, y, 1.5])
            T = S.look_at(eye, eye + np.array([dx, dy, 0.0]), up=(0.0, 0.0, 1.0))
            m.add_depth_frame(far, torch.from_numpy(T), K_t)
            n_pop += 1
        n_blocks = layer.num_blocks()
        if n_blocks >= target_blocks:
            break
    torch.cuda.synchronize()
    t_pop = time.perf_counter() - t_pop
    g = torch.Generator(device=dev)
    feats = []
    for i in range(4):
        g.manual_seed(7000 + i)
        feats.append(torch.randn((HS, HS, C_S), generator=g, device=dev, dtype=torch.float32).half())
    frames = []
    for step in range(n_steps + 1):
        for T in rig_poses(step):
            frames.append((torch.from_numpy(T), torch.from_numpy(S.render_depth(K, HS, HS, T, **SCENE)).to(dev)))

    def run_step(step):
        for cam in range(4):
            T, d = frames[4 * step + cam]
            m.add_depth_frame(d, T, K_t)
            m.add_feature_frame(feats[cam], T, K_t)
        m.decay()

    run_step(0)                                   # warm-up: feature arena growth happens here
    m.reset_counters(0)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for step in range(1, n_steps + 1):
        run_step(step)
    e1.record()
    m.update_feature_mesh(0)
    mesh = m.get_feature_mesh(0)
    n_v = int(mesh.vertices().shape[0])
    e2.record()
    torch.cuda.synchronize()
    c = m.counters(0)
    ms_steps, ms_export = e0.elapsed_time(e1), e1.elapsed_time(e2)
    n_frames = 4 * n_steps
    out = {'workload': 'stress: 4 cams 1024x1024, C=1024, 1 cm voxels, 20 x 20 x 4 m workspace',
           'tsdf_blocks': int(layer.num_blocks()), 'tsdf_bytes': int(layer.num_blocks()) * 4096,
           'populate': {'depth_frames': n_pop, 'seconds': t_pop, 'tsdf_blocks': int(n_blocks)},
           'feature_blocks': int(m.feature_layer_view(0).num_blocks()),
           'steps': n_steps, 'camera_frames': n_frames, 'camera_frames_per_s': n_frames / (ms_steps / 1e3),
           'ms_per_step_4_cams_plus_decay': ms_steps / n_steps,
           'feature_voxels_updated_per_frame': c['feature_voxels_updated'] / n_frames,
           'tsdf_voxels_updated_per_frame': c['tsdf_voxels_updated'] / n_frames,
           'export': {'vertices': n_v, 'ms': ms_export, 'vertices_per_s': n_v / (ms_export / 1e3),
                      'bytes': n_v * (12 + 2 * C_S), 'mesh_blocks_remeshed': c['mesh_blocks_remeshed']},
           'note': 'decay walks all TSDF blocks every step; the export re-meshes every block the decay marked '
                   '(BlocksToUpdateTracker semantics, ours-only beyond the reference tracker\'s 100 k-block cap, SURVEY Q7)'}
    if keep:      # tests/test_gpu_stress.py inspects the map itself
        return out, m
    del m, feats, frames, mesh
    torch.cuda.empty_cache()
    constants.set_feature_array_num_elements(C_FEAT)
    return out


def cold_start_stage(depths, poses, feats, K_t, local_rank, n=8):
    """What the warm timed region never shows: the first `n` frames of an episode -- block allocation and arena
    growth included -- on a FRESH Mapper (device arenas are allocated inside the region) and after clear() (arenas
    kept, every block re-allocated).  Wall clock around a synchronised region."""
    import torch
    from nvblox_torch.mapper import Mapper
    mp, _ = mapper_params()

    def first_frames(mapper):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            mapper.add_depth_frame(depths[i], poses[i], K_t)
            mapper.add_feature_frame(feats[i % len(feats)], poses[i], K_t)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    m = Mapper(voxel_sizes_m=VOXEL, mapper_parameters=mp, device=local_rank)
    fresh = first_frames(m)
    c = m.counters(0)
    m.clear()
    m.reset_counters(0)
    cleared = first_frames(m)
    c2 = m.counters(0)
    return {'frames': n, 'fresh_mapper_ms_per_frame': 1e3 * fresh / n, 'after_clear_ms_per_frame': 1e3 * cleared / n,
            'fresh_mapper_frames_per_s': n / fresh, 'after_clear_frames_per_s': n / cleared,
            'tsdf_blocks_allocated': c['tsdf_blocks_allocated'], 'feature_blocks_allocated': c['feature_blocks_allocated'],
            'after_clear_tsdf_blocks_allocated': c2['tsdf_blocks_allocated']}


def ncu_traffic_bytes():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` summary
    (profiles/r02am_feature_gather.md: dram__bytes_read.sum + dram__bytes_write.sum, first capture)."""
    path = os.path.join(ROOT, 'profiles', 'r02am_feature_gather.md')
    try:
        rd = wr = None
        for line in open(path):
            cells = [c.strip() for c in line.split('|')]
            if len(cells) > 3 and cells[1] == 'dram__bytes_read.sum' and rd is None:
                rd = float(cells[2]) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}[cells[3]]
            if len(cells) > 3 and cells[1] == 'dram__bytes_write.sum' and wr is None:
                wr = float(cells[2]) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}[cells[3]]
        if rd is not None and wr is not None:
            return rd + wr, 'profiles/r02am_feature_gather.md (ncu --set full, one launch of the same workload)'
    except Exception:
        pass
    return None, None


def fused_upsample_stage(args, mapper, depths, poses, K_t, dev):
    """SURVEY 8(f) N4, the frame as mindmap really produces it: the backbone hands over a [1, 768, 32, 32] fp32
    patch-feature map (RADIO's permuted view, i.e. channels-last).  `chained` = mindmap's own tail of
    FeatureExtractor.compute() (F.interpolate -> HWC -> .to(float16), feature_extraction.py:188-196 +
    nvblox_mapping_helpers.py:256) followed by add_depth_frame + add_feature_frame; `fused` = add_depth_frame +
    add_feature_frame_lowres (bit-identical map, tests/test_gpu_upsample.py).  Device-timed, plus the host-buffer
    end-to-end rate of the fused entry (2.5 MB of H2D per frame)."""
    import torch
    import torch.nn.functional as F
    n = max(8, min(args.steps, 256))
    w = min(args.warmup, 16)
    g = torch.Generator(device=dev)
    lows = []
    for i in range(N_FEATURE_BUFFERS):
        g.manual_seed(5000 + i)
        hwc = torch.randn((1, H // 16, W // 16, C_FEAT), generator=g, device=dev, dtype=torch.float32)
        lows.append(hwc.permute(0, 3, 1, 2))                       # [1, c, h, w] view, channels-last strides

    def chained(i):
        up = F.interpolate(lows[i % len(lows)], size=(H, W), mode='bilinear', align_corners=False)
        frame = up[0].permute(1, 2, 0).contiguous().to(dtype=torch.float16)
        mapper.add_depth_frame(depths[i], poses[i], K_t)
        mapper.add_feature_frame(frame, poses[i], K_t)

    def fused(i):
        mapper.add_depth_frame(depths[i], poses[i], K_t)
        mapper.add_feature_frame_lowres(lows[i % len(lows)], (H, W), poses[i], K_t)

    out = {}
    for name, fn in (('chained', chained), ('fused', fused)):
        for i in range(w):
            fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(w, w + n):
            fn(i % len(depths))
        e1.record()
        torch.cuda.synchronize()
        out[name + '_ms_per_frame'] = e0.elapsed_time(e1) / n
    out['frames'] = n
    out['fused_frames_per_s'] = 1000.0 / out['fused_ms_per_frame']
    out['chained_frames_per_s'] = 1000.0 / out['chained_ms_per_frame']
    sub = argparse.Namespace(steps=n, warmup=w)
    import ctypes as C
    from nvblox_mindmap_b200 import _capi
    lib = _capi.load()
    lib.nvbx_set_kernel_timing(mapper._handle, 1)
    for i in range(w, w + n):
        fused(i % len(depths))
    ms, cnt = C.c_double(), C.c_int64()
    lib.nvbx_get_kernel_timing(mapper._handle, 0, C.byref(ms), C.byref(cnt))
    lib.nvbx_set_kernel_timing(mapper._handle, 0)
    out['gather_up_kernel_us'] = 1000.0 * ms.value / cnt.value if cnt.value else None
    # host buffers -> map, through the public API
    n_e2e = min(n, 64)
    h_depth = [depths[i].cpu().pin_memory() for i in range(n_e2e)]
    h_low = [l.permute(0, 2, 3, 1).contiguous().cpu().pin_memory().permute(0, 3, 1, 2) for l in lows]
    for i in range(3):
        mapper.integrate_frame_from_host_lowres(h_depth[i], h_low[i % len(h_low)], poses[i], K_t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        mapper.integrate_frame_from_host_lowres(h_depth[i], h_low[i % len(h_low)], poses[i], K_t)
        mapper.counters(0)
    torch.cuda.synchronize()
    out['fused_e2e_frames_per_s'] = n_e2e / (time.perf_counter() - t0)
    out['fused_e2e_h2d_bytes_per_step'] = H * W * 4 + (H // 16) * (W // 16) * C_FEAT * 4
    out['note'] = ('not the headline metric: the headline integrates the materialised 384 MiB frame; this is the '
                   'same map content built from the backbone output directly (bit-identical, tests/test_gpu_upsample.py)')
    return out


def export_stage(mapper, n=8):
    """Per-step export as mindmap runs it (decay -> update_feature_mesh -> get_feature_mesh), device-timed."""
    import torch
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    nv = 0
    for a, b in ev:
        a.record()
        mapper.decay()
        mapper.update_feature_mesh(0)
        mesh = mapper.get_feature_mesh(0)
        nv = int(mesh.vertices().shape[0])
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return {'ms_median': ms[len(ms) // 2], 'vertices': nv,
            'note': 'decay marks every block dirty, so each update re-meshes the whole map (SURVEY 3.4)'}


def per_kernel_us(args, mapper, depths, poses, feats, K_t):
    """In-pipeline duration of every kernel (event pair around each launch; includes the launch gap)."""
    import ctypes as C
    from nvblox_mindmap_b200 import _capi
    lib = _capi.load()
    lib.nvbx_set_kernel_timing(mapper._handle, 2)
    n_total = args.warmup + args.steps
    for i in range(args.warmup, n_total):
        mapper.add_depth_frame(depths[i], poses[i], K_t)
        mapper.add_feature_frame(feats[i % N_FEATURE_BUFFERS], poses[i], K_t)
    buf = C.create_string_buffer(1 << 16)
    n = lib.nvbx_kernel_timing_report(mapper._handle, buf, len(buf))
    lib.nvbx_set_kernel_timing(mapper._handle, 0)
    rep = json.loads(buf.value.decode()) if n > 0 else {}
    return {k: round(1000.0 * v[0] / max(1, v[1]), 2) for k, v in rep.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1024)
    ap.add_argument('--warmup', type=int, default=64)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--sweep-gather', action='store_true', help='tuning aid: time every gather schedule')
    ap.add_argument('--stream-priority', type=int, default=0,
                    help='tuning aid: run the mapper on a torch stream of this priority (0 default, -1 .. -5 higher)')
    ap.add_argument('--quick', action='store_true', help='tuning aid: device-timed pass + kernel timing only')
    ap.add_argument('--pipelining', type=int, default=2,
                    help='frame pipelining in the device-timed pass: 0 off, 1 on, 2 on + asynchronous enqueue (Mapper.set_pipelining)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if args.stream_priority:
        import torch
        torch.cuda.set_device(local_rank)
        with torch.cuda.stream(torch.cuda.Stream(device=local_rank, priority=args.stream_priority)):
            run_ours(args, rank, world, local_rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
