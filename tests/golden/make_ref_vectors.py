"""Reference-COMPILED golden vectors (run on a B200: `gpurun -- python tests/golden/make_ref_vectors.py`).

oracle/_ref/libref_kernels_c<C>.so holds the REFERENCE's own device code -- nvblox's integrateBlocksKernel
(TSDF / feature / colour), combinedBlockIndicesInImageKernel, sphereTracingKernel, the marching-cubes and
mesh-appearance kernels and the scalar functions they call -- compiled verbatim with nvblox's nvcc flags
(oracle/ref_snippets/Makefile).  This script runs that code on seeded inputs and records what it produces:

  ref_nvcc_functions.npz     function-level probes (inputs + outputs)
  ref_nvcc_pipeline_*.npz    whole frame sequences: the oracle does the reference's HOST work, every DEVICE
                             stage is handed to the reference kernel through orc_set_ref_hooks
  ref_nvcc_report.json       what the un-hooked oracle (nvcc model and ieee model) differs by, for the log

Output goes to gpurun_out/ref_vectors/ (copied into tests/golden/ by hand and committed).  The GPU box has
no /root/reference: only the prebuilt .so travels.
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import oracle as O  # noqa: E402
from tests import ref_scenarios as RS  # noqa: E402

OUT = os.path.join(ROOT, 'gpurun_out', 'ref_vectors')
DEV = 'cuda:0'


def ref_lib(channels: int) -> C.CDLL:
    path = os.path.join(ROOT, 'oracle', '_ref', f'libref_kernels_c{channels}.so')
    L = C.CDLL(path)
    assert L.ref_feature_channels() == channels
    assert L.ref_sizeof_tsdf_block() == 4096
    assert L.ref_sizeof_feature_voxel() == 2 * (channels + 1), L.ref_sizeof_feature_voxel()
    assert L.ref_sizeof_color() == 3 and L.ref_sizeof_color_voxel() == 8
    return L


def dptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def i32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


F = C.c_float


# ---------------------------------------------------------------------------------------------------------
# function-level probes
# ---------------------------------------------------------------------------------------------------------
def function_vectors(L, channels):
    out = {}
    inp = RS.function_inputs(channels)
    n = len(inp['interp_xy'])
    xy, fh, ff = dev(inp['interp_xy']), dev(inp['interp_half_f']), dev(inp['interp_float_f'])
    oh = torch.zeros(n, dtype=torch.float16, device=DEV)
    assert L.ref_fn_interp_half(n, dptr(xy), dptr(fh), dptr(oh)) == 0
    of = torch.zeros(n, dtype=torch.float32, device=DEV)
    assert L.ref_fn_interp_float(n, dptr(xy), dptr(ff), dptr(of)) == 0
    out['interp_half_out'] = oh.cpu().numpy().view(np.uint16)
    out['interp_float_out'] = of.cpu().numpy()

    a, b, w = dev(inp['blend_a']), dev(inp['blend_b']), dev(inp['blend_w'])
    ob = torch.zeros_like(a)
    assert L.ref_fn_blend(a.shape[0], dptr(a), dptr(b), dptr(w), dptr(ob)) == 0
    out['blend_out'] = ob.cpu().numpy().view(np.uint16)

    v1, v2, sdf = dev(inp['vertex_v1']), dev(inp['vertex_v2']), dev(inp['vertex_sdf'])
    ov = torch.zeros_like(v1)
    assert L.ref_fn_interp_vertex(v1.shape[0], dptr(v1), dptr(v2), dptr(sdf), dptr(ov)) == 0
    out['vertex_out'] = ov.cpu().numpy()

    for k, bs in enumerate(inp['block_sizes']):
        p = dev(inp[f'bv_points_{k}'])
        o = torch.zeros((p.shape[0], 6), dtype=torch.int32, device=DEV)
        assert L.ref_fn_block_voxel(p.shape[0], F(bs), dptr(p), dptr(o)) == 0
        out[f'bv_out_{k}'] = o.cpu().numpy()

    for k in range(len(inp['project_poses'])):
        T = np.ascontiguousarray(inp['project_poses'][k], np.float32).reshape(16)
        fx, fy, cx, cy, W, H, bs, maxd = inp['project_cams'][k]
        bv = dev(inp[f'project_bv_{k}'])
        o = torch.zeros((bv.shape[0], 6), dtype=torch.float32, device=DEV)
        ok = torch.zeros(bv.shape[0], dtype=torch.int32, device=DEV)
        assert L.ref_fn_project(bv.shape[0], f32p(T), F(fx), F(fy), F(cx), F(cy), int(W), int(H), F(bs), F(maxd),
                                dptr(bv), dptr(o), dptr(ok)) == 0
        out[f'project_out_{k}'] = o.cpu().numpy()
        out[f'project_ok_{k}'] = ok.cpu().numpy()

    for mode in range(6):
        tin, act, vox = dev(inp['tsdf_in']), dev(inp['tsdf_active']), dev(inp['tsdf_voxels'].copy())
        upd = torch.zeros(tin.shape[0], dtype=torch.uint8, device=DEV)
        trunc, maxw, inv = inp['tsdf_params']
        assert L.ref_fn_tsdf_functor(tin.shape[0], F(trunc), F(maxw), F(inv), mode, dptr(tin), dptr(act), dptr(vox),
                                     dptr(upd)) == 0
        out[f'tsdf_out_{mode}'] = vox.cpu().numpy()
        out[f'tsdf_updated_{mode}'] = upd.cpu().numpy()
    return inp, out


# ---------------------------------------------------------------------------------------------------------
# hooks: the oracle's device stages -> the reference's kernels
# ---------------------------------------------------------------------------------------------------------
class RefHooks:
    def __init__(self, L, channels):
        self.L, self.C = L, channels
        OL = O.lib()
        fp, u8p, u16p, ip = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_int)
        self.t_ray = C.CFUNCTYPE(C.c_int, fp, fp, C.c_int, C.c_int, fp, F, F, F, C.c_int, ip, ip, u8p)
        self.t_tsdf = C.CFUNCTYPE(C.c_int, fp, fp, C.c_int, C.c_int, fp, u8p, F, F, F, F, F, C.c_int, ip, C.c_int, fp)
        self.t_trace = C.CFUNCTYPE(C.c_int, fp, fp, C.c_int, C.c_int, ip, C.c_int, fp, F, F, C.c_int, F, F, C.c_int, fp)
        self.t_feat = C.CFUNCTYPE(C.c_int, fp, fp, C.c_int, C.c_int, u16p, u8p, fp, C.c_int, F, F, F, F, F, ip,
                                  C.c_int, u16p)
        self.t_color = C.CFUNCTYPE(C.c_int, fp, fp, C.c_int, C.c_int, u8p, u8p, fp, C.c_int, F, F, F, F, F, ip,
                                   C.c_int, u8p, fp)
        self.cb = (self.t_ray(self.raycast), self.t_tsdf(self.tsdf), self.t_trace(self.trace),
                   self.t_feat(self.feat), self.t_color(self.color))
        OL.orc_set_ref_hooks.argtypes = [self.t_ray, self.t_tsdf, self.t_trace, self.t_feat, self.t_color]
        OL.orc_set_ref_hooks.restype = None
        self.OL = OL
        self.calls = dict(raycast=0, tsdf=0, trace=0, feat=0, color=0)

    def install(self):
        self.OL.orc_set_ref_hooks(*self.cb)

    def remove(self):
        self.OL.orc_set_ref_hooks(self.t_ray(0), self.t_tsdf(0), self.t_trace(0), self.t_feat(0), self.t_color(0))

    @staticmethod
    def arr(p, shape, dtype):
        n = int(np.prod(shape))
        if n == 0 or not p:
            return np.zeros(shape, dtype)
        return np.ctypeslib.as_array(p, shape=(n,)).view(dtype).reshape(shape)

    def raycast(self, T, cam, W, H, depth, bs, maxd, behind, sub, amn, asz, grid):
        self.calls['raycast'] += 1
        size = self.arr(asz, (3,), np.int32)
        n = int(size[0]) * int(size[1]) * int(size[2])
        d = dev(self.arr(depth, (H, W), np.float32))
        g = torch.zeros(n, dtype=torch.uint8, device=DEV)
        rc = self.L.ref_raycast_blocks(T, F(cam[0]), F(cam[1]), F(cam[2]), F(cam[3]), W, H, dptr(d), F(bs), F(maxd),
                                       F(behind), sub, amn, asz, dptr(g))
        self.arr(grid, (n,), np.uint8)[:] = g.cpu().numpy()
        return rc

    def tsdf(self, T, cam, W, H, depth, mask, bs, maxd, trunc, maxw, inv, mode, idx, n, vox):
        self.calls['tsdf'] += 1
        d = dev(self.arr(depth, (H, W), np.float32))
        m = dev(self.arr(mask, (H, W), np.uint8)) if mask else None
        v = self.arr(vox, (n, 1024), np.float32)
        store = dev(v)
        slots = np.arange(n, dtype=np.int32)
        rc = self.L.ref_integrate_tsdf(T, F(cam[0]), F(cam[1]), F(cam[2]), F(cam[3]), W, H, dptr(d), dptr(m), F(bs),
                                       F(maxd), F(trunc), F(maxw), F(inv), mode, idx, i32p(slots), n, dptr(store))
        v[:] = store.cpu().numpy()
        return rc

    def trace(self, T, cam, W, H, idx, n_all, vox, trunc, bs, steps, maxlen, eps, sub, out):
        self.calls['trace'] += 1
        store = dev(self.arr(vox, (n_all, 1024), np.float32))
        rows, cols = H // sub, W // sub
        o = torch.zeros((rows, cols), dtype=torch.float32, device=DEV)
        rc = self.L.ref_sphere_trace(T, F(cam[0]), F(cam[1]), F(cam[2]), F(cam[3]), W, H, idx, n_all, dptr(store),
                                     F(trunc), F(bs), steps, F(maxlen), F(eps), sub, dptr(o))
        self.arr(out, (rows, cols), np.float32)[:] = o.cpu().numpy()
        return rc

    def feat(self, T, cam, W, H, img, mask, synth, sub, bs, maxd, trunc, maxw, alpha, idx, n, fvox):
        self.calls['feat'] += 1
        Cc = self.C
        im = dev(self.arr(img, (H, W, Cc), np.uint16).view(np.float16))
        m = dev(self.arr(mask, (H, W), np.uint8)) if mask else None
        sd = dev(self.arr(synth, (H // sub, W // sub), np.float32))
        v = self.arr(fvox, (n, 512 * (Cc + 1)), np.uint16)
        store = dev(v.view(np.float16))
        slots = np.arange(n, dtype=np.int32)
        rc = self.L.ref_integrate_features(T, F(cam[0]), F(cam[1]), F(cam[2]), F(cam[3]), W, H, dptr(im), dptr(m),
                                           dptr(sd), sub, F(bs), F(maxd), F(trunc), F(maxw), F(alpha), idx,
                                           i32p(slots), n, dptr(store))
        v[:] = store.cpu().numpy().view(np.uint16)
        return rc

    def color(self, T, cam, W, H, img, mask, synth, sub, bs, maxd, trunc, maxw, alpha, idx, n, rgb, wgt):
        self.calls['color'] += 1
        im = dev(self.arr(img, (H, W, 3), np.uint8))
        m = dev(self.arr(mask, (H, W), np.uint8)) if mask else None
        sd = dev(self.arr(synth, (H // sub, W // sub), np.float32))
        r = self.arr(rgb, (n, 512, 3), np.uint8)
        w = self.arr(wgt, (n, 512), np.float32)
        vox = np.zeros((n, 512, 8), np.uint8)   # ColorVoxel {uint8 rgb[3]; pad; float weight}
        vox[:, :, :3] = r
        vox[:, :, 4:8] = w.reshape(n, 512, 1).view(np.uint8)
        store = dev(vox)
        slots = np.arange(n, dtype=np.int32)
        rc = self.L.ref_integrate_color(T, F(cam[0]), F(cam[1]), F(cam[2]), F(cam[3]), W, H, dptr(im), dptr(m),
                                        dptr(sd), sub, F(bs), F(maxd), F(trunc), F(maxw), F(alpha), idx, i32p(slots),
                                        n, dptr(store))
        back = store.cpu().numpy()
        r[:] = back[:, :, :3]
        w[:] = np.ascontiguousarray(back[:, :, 4:8]).view(np.float32).reshape(n, 512)
        return rc


def ref_mesh(L, idx, tsdf, block_size, voxel_size, min_weight, max_v=7680):
    """Reference marching cubes (K8 + K9, no weld) over every block: per-block vertex rows, canonically sorted."""
    n = len(idx)
    lut = {tuple(b): i for i, b in enumerate(idx.tolist())}
    nb = np.full((n, 8), -1, np.int32)
    for i, b in enumerate(idx.tolist()):
        for j in range(8):
            d = ((j & 4) >> 2, (j & 2) >> 1, j & 1)
            nb[i, j] = lut.get((b[0] + d[0], b[1] + d[1], b[2] + d[2]), -1)
    store = dev(tsdf.reshape(n, 1024))
    verts = np.zeros((n, max_v, 3), np.float32)
    norms = np.zeros((n, max_v, 3), np.float32)
    counts = np.zeros(n, np.int32)
    bi = np.ascontiguousarray(idx, np.int32)
    rc = L.ref_mesh_blocks(i32p(bi), i32p(nb), n, dptr(store), F(block_size), F(voxel_size), F(min_weight), max_v,
                           f32p(verts), f32p(norms), i32p(counts))
    assert rc == 0, rc
    return verts, counts


def ref_paint(L, idx, counts, verts, channels, block_size, max_v):
    """Reference closest-voxel paint (K11): the feature blocks hold their own voxel's linear index in channel 0, so
    the returned value IS the voxel the reference picked for every vertex."""
    n = len(idx)
    fb = np.zeros((n, 512, channels + 1), np.float16)
    fb[:, :, 0] = np.arange(512, dtype=np.float16)[None, :]
    store = dev(fb)
    vd = dev(verts)
    out = torch.zeros((n, max_v, channels), dtype=torch.float16, device=DEV)
    slots = np.arange(n, dtype=np.int32)
    bi = np.ascontiguousarray(idx, np.int32)
    rc = L.ref_paint_features(i32p(bi), i32p(slots), i32p(np.ascontiguousarray(counts, np.int32)), n, max_v,
                              dptr(store), F(block_size), dptr(vd), dptr(out))
    assert rc == 0, rc
    return out[:, :, 0].cpu().numpy().astype(np.int32)


def run_pipeline(name, L, channels, with_mesh=True):
    sc = RS.SCENARIOS[name]
    hooks = RefHooks(L, channels)
    O.lib().orc_set_fp_model(1)
    hooks.install()
    try:
        res = RS.run_scenario(sc, channels, RS.OracleBackend(sc, channels))
    finally:
        hooks.remove()
    res['hook_calls'] = json.dumps(hooks.calls)
    if with_mesh:
        idx, tsdf = res['tsdf_idx'], res['tsdf_data']
        vs = np.float32(sc['voxel_size'])
        bs = np.float32(vs * np.float32(8))
        max_v = 7680
        verts, counts = ref_mesh(L, idx, tsdf, bs, vs, np.float32(1e-4), max_v)
        paint = ref_paint(L, idx, counts, verts, channels, bs, max_v)
        rows, cnt = [], []
        for i in range(len(idx)):
            v = verts[i, :counts[i]]
            p = paint[i, :counts[i]]
            r = np.concatenate([v.view(np.uint32).astype(np.int64), p[:, None].astype(np.int64)], axis=1)
            r = r[np.lexsort(r.T[::-1])] if len(r) else r
            rows.append(r)
            cnt.append(len(r))
        allr = np.concatenate(rows) if rows else np.zeros((0, 4), np.int64)
        res['mesh_vertex_bits'] = allr[:, :3].astype(np.uint32)   # per block, rows sorted
        res['mesh_vertex_voxel'] = allr[:, 3].astype(np.int16)    # voxel picked by the reference's paint kernel
        res['mesh_counts'] = np.asarray(cnt, np.int32)
    return res


def compare(a, b):
    """Number of differing elements per key (shape mismatch = -1)."""
    out = {}
    for k in a:
        if k not in b or not isinstance(a[k], np.ndarray):
            continue
        if a[k].shape != b[k].shape:
            out[k] = -1
            continue
        x = a[k].view(np.uint8) if a[k].dtype.kind == 'f' else a[k]
        y = b[k].view(np.uint8) if b[k].dtype.kind == 'f' else b[k]
        if a[k].dtype.kind == 'f':
            x = x.reshape(-1, a[k].dtype.itemsize)
            y = y.reshape(-1, a[k].dtype.itemsize)
            out[k] = int((x != y).any(axis=1).sum())
        else:
            out[k] = int((x != y).sum())
    return out


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    report = {'gpu': torch.cuda.get_device_name(0)}
    L16 = ref_lib(16)
    inp, out = function_vectors(L16, 16)
    np.savez_compressed(os.path.join(OUT, 'ref_nvcc_functions.npz'), **out)
    report['functions'] = {}
    for model in (1, 0):
        O.lib().orc_set_fp_model(model)
        mine = RS.function_outputs_oracle(inp, 16)
        report['functions'][f'oracle_model_{model}_mismatches'] = compare(out, mine)
    O.lib().orc_set_fp_model(1)
    print(json.dumps(report['functions'], indent=1), flush=True)

    report['pipelines'] = {}
    for name in RS.SCENARIOS:
        res = run_pipeline(name, L16, 16)
        np.savez_compressed(os.path.join(OUT, f'ref_nvcc_pipeline_{name}.npz'),
                            **{k: v for k, v in res.items() if isinstance(v, np.ndarray)})
        entry = {'hook_calls': res['hook_calls'], 'sizes': {k: list(v.shape) for k, v in res.items()
                                                           if isinstance(v, np.ndarray)}}
        sc = RS.SCENARIOS[name]
        for model in (1, 0):
            O.lib().orc_set_fp_model(model)
            mine = RS.run_scenario(sc, 16, RS.OracleBackend(sc, 16))
            mine.update(RS.oracle_mesh_rows(sc, 16, mine))
            entry[f'oracle_model_{model}_mismatches'] = compare(res, mine)
        O.lib().orc_set_fp_model(1)
        report['pipelines'][name] = entry
        print(name, json.dumps(entry, indent=1), flush=True)

    # the headline shape: C = 768; the maps are too large to commit, so digests + a sample of rows
    L768 = ref_lib(768)
    res = run_pipeline('cube_stacking', L768, 768, with_mesh=False)
    dg = {k: digest(v) for k, v in res.items() if isinstance(v, np.ndarray)}
    rng = np.random.default_rng(7)
    fd = res['feat_data'].reshape(-1, 769)
    nz = np.flatnonzero(fd[:, 768].view(np.uint16) != 0)
    pick = np.sort(rng.choice(nz, size=min(256, len(nz)), replace=False)) if len(nz) else np.zeros(0, np.int64)
    np.savez_compressed(os.path.join(OUT, 'ref_nvcc_pipeline_cube_stacking_c768.npz'),
                        feat_idx=res['feat_idx'], tsdf_idx=res['tsdf_idx'], sample_rows=pick,
                        sample_values=fd[pick].view(np.uint16), digests=np.asarray(json.dumps(dg)))
    O.lib().orc_set_fp_model(1)
    mine = RS.run_scenario(RS.SCENARIOS['cube_stacking'], 768, RS.OracleBackend(RS.SCENARIOS['cube_stacking'], 768))
    report['pipelines']['cube_stacking_c768'] = {'oracle_model_1_mismatches': compare(res, mine),
                                                 'updated_feature_voxels': int(len(nz))}
    print('c768', json.dumps(report['pipelines']['cube_stacking_c768'], indent=1), flush=True)
    with open(os.path.join(OUT, 'ref_nvcc_report.json'), 'w') as f:
        json.dump(report, f, indent=1)
    print('done')


if __name__ == '__main__':
    main()
