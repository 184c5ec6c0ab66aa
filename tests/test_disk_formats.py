"""On-disk formats: CPU checks of the containers, GPU round trips through the map."""
import os
import pickle
import sqlite3
import struct

import numpy as np
import pytest

from tests import scenes as S


def test_zstd_frame_is_what_the_reference_reader_expects():
    """nvblox_to_disk_helpers.py:60-65 writes ZstdCompressor(level=1).compress(pickle) -- one zstd frame with the
    content size in its header; dataset.py:410-415 reads it back with a streaming decompressor."""
    from nvblox_mindmap_b200 import disk_helpers as D
    rng = np.random.default_rng(0)
    payload = pickle.dumps({'x': rng.standard_normal(5000).astype(np.float16)}, protocol=pickle.HIGHEST_PROTOCOL)
    blob = D.zstd_compress(payload, level=1)
    assert blob[:4] == D.ZSTD_MAGIC                                  # 0xFD2FB528 little endian
    fhd = blob[4]                                                    # frame header descriptor (RFC 8878 3.1.1.1)
    assert (fhd >> 6) != 0 or (fhd & 0x20)                           # Frame_Content_Size is present
    assert D.zstd_decompress(blob) == payload
    assert D.zstd_decompress(D.zstd_compress(b'', 1)) == b''
    with pytest.raises(RuntimeError):
        D.zstd_decompress(b'not a zstd frame at all')


def test_vertex_features_container_roundtrip(tmp_path):
    import torch
    from nvblox_mindmap_b200 import disk_helpers as D
    g = torch.Generator().manual_seed(1)
    v = torch.rand((321, 3), generator=g)
    f = torch.randn((321, 24), generator=g).half()
    path = str(tmp_path / f'0007.{D.VERTEX_FEATURES_FILE_NAME}')
    D.write_vertex_features(path, v, f)
    ob = D.read_vertex_features(path)
    assert set(ob) == {'vertices', 'features', 'channel_length'} and ob['channel_length'] == 24
    assert ob['vertices'].dtype == torch.float16 and ob['features'].dtype == torch.float16
    assert torch.equal(ob['vertices'], v.half()) and torch.equal(ob['features'], f)


def test_nvblx_schema(tmp_path):
    """Tables and metadata rows exactly as Serializer::createLayerTables / setLayerParameter* write them
    (serializer.cpp:170-205,236-271)."""
    from nvblox_mindmap_b200 import disk_helpers as D
    path = str(tmp_path / 'm.nvblx')
    db = sqlite3.connect(path)
    for name in D.NVBLX_LAYERS:
        D._create_layer_tables(db, name, float(np.float32(0.02) * np.float32(8)))
    db.execute('INSERT INTO tsdf_layer_data (index_x, index_y, index_z, data) VALUES (?,?,?,?)',
               (1, -2, 3, struct.pack('<2f', 0.5, 1.0) * 512))
    db.commit()
    db.close()
    db = sqlite3.connect(path)
    tables = {r[0] for r in db.execute("SELECT name FROM sqlite_master WHERE type='table';")}
    assert tables == {f'{n}_{s}' for n in D.NVBLX_LAYERS for s in ('metadata', 'data')}
    # the reference's getLayerNames / getLayerParameter* queries (serializer.cpp:284-316,338-349)
    names = [r[0] for r in db.execute("SELECT name FROM sqlite_master WHERE type='table' AND name NOT LIKE "
                                      "'sqlite_%' AND name LIKE '%_metadata';")]
    assert sorted(n.replace('_metadata', '') for n in names) == sorted(D.NVBLX_LAYERS)
    assert [r[0] for r in db.execute('SELECT param_name FROM tsdf_layer_metadata WHERE value_string IS NOT NULL;')] \
        == ['type']
    assert [r[0] for r in db.execute('SELECT param_name FROM tsdf_layer_metadata WHERE value_float IS NOT NULL;')] \
        == ['block_size']
    bs = db.execute("SELECT value_float FROM tsdf_layer_metadata WHERE param_name = 'block_size';").fetchone()[0]
    assert isinstance(bs, float) and abs(bs - 0.16) < 1e-6
    assert db.execute("SELECT value_string FROM color_layer_metadata WHERE param_name = 'type';").fetchone()[0] \
        == 'color_layer'
    db.close()
    assert D.nvblx_summary(path)['tsdf_layer'] == (pytest.approx(0.16, abs=1e-6), 1, 4096)
    assert D.nvblx_summary(path)['feature_layer'][1:] == (0, 0)


@pytest.mark.gpu
def test_nvblx_save_load_roundtrip(tmp_path):
    """save_map -> load_from_file into a fresh mapper reproduces every layer and both meshes of the oracle."""
    import torch
    from nvblox_mindmap_b200 import disk_helpers as D
    from nvblox_torch.mapper import Mapper
    from tests.parity_utils import Pair, make_params, orbit_frames
    C_feat = 16
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, alpha=0.5)
    pair = Pair(0.02, C_feat, mp, op)
    for i, T, K, depth, feat in orbit_frames(3, 64, 64, C_feat, S.S_TABLE):
        pair.depth(depth, T, K)
        pair.color(S.color_frame(64, 64, 30 + i), T, K)
        pair.features(feat, T, K)
    path = str(tmp_path / '0000.nvblox_map_static.nvblx')
    pair.gpu.save_map(path, 0)
    summ = D.nvblx_summary(path)
    n_tsdf, n_feat, n_col = (pair.cpu.num_blocks(k) for k in (0, 1, 2))
    assert summ['tsdf_layer'][1:] == (n_tsdf, n_tsdf * 4096)
    assert summ['color_layer'][1:] == (n_col, n_col * 4096)
    assert summ['feature_layer'][1:] == (n_feat, n_feat * 512 * (C_feat + 1) * 2)   # packed (C+1)-half rows
    assert summ['esdf_layer'][1:] == (0, 0)
    # blob == the reference's voxels array: check one TSDF block against the oracle byte for byte
    idx, data = pair.cpu.all_blocks(0)
    db = sqlite3.connect(path)
    blob = db.execute('SELECT data FROM tsdf_layer_data WHERE index_x=? AND index_y=? AND index_z=?',
                      tuple(int(v) for v in idx[0])).fetchone()[0]
    db.close()
    assert blob == data[0].tobytes()
    fresh = Mapper(voxel_sizes_m=0.02, mapper_parameters=mp)
    fresh.load_from_file(path, 0)
    pair.gpu = fresh
    assert pair.check_tsdf() == n_tsdf
    assert pair.check_features() == n_feat
    assert pair.check_color() == n_col
    # loadMap re-meshes the full layer (mapper.cpp:887-900): both meshes are there without an update call
    assert fresh.get_feature_mesh(0).vertices().shape[0] > 0 and fresh.get_color_mesh(0).vertices().shape[0] > 0
    pair.cpu.mark_all_dirty()
    assert pair.check_mesh() > 0
    pair.cpu.mark_all_dirty()
    assert pair.check_color_mesh() > 0
    assert fresh.load_from_file(str(tmp_path / 'missing.nvblx'), 0) is False


@pytest.mark.gpu
def test_save_feature_mesh_to_disk(tmp_path):
    import torch
    from nvblox_mindmap_b200 import disk_helpers as D
    from nvblox_mindmap_b200.output_helpers import get_vertices_and_features
    from tests.parity_utils import Pair, make_params, orbit_frames
    C_feat = 32
    mp, op = make_params(workspace=S.WS_CUBE_STACKING)
    pair = Pair(0.02, C_feat, mp, op)
    for i, T, K, depth, feat in orbit_frames(2, 64, 64, C_feat, S.S_TABLE):
        feat[..., 24:] = 0
        pair.depth(depth, T, K)
        pair.features(feat, T, K)

    class Cfg:
        aabb_min_m = torch.tensor(S.WS_CUBE_STACKING[0])
        aabb_max_m = torch.tensor(S.WS_CUBE_STACKING[1])

    v, f = D.save_feature_mesh_to_disk(pair.gpu, Cfg, 8, 12, str(tmp_path), include_dynamic=False)
    ob = D.read_vertex_features(os.path.join(str(tmp_path), '0012.nvblox_vertex_features.zst'))
    assert ob['channel_length'] == 24 and ob['features'].shape == (v.shape[0], 24) and v.shape[0] > 0
    assert torch.equal(ob['vertices'], v.half().cpu()) and torch.equal(ob['features'], f.cpu())
    with pytest.raises(AssertionError):
        D.save_feature_mesh_to_disk(pair.gpu, Cfg, 8, 13, str(tmp_path), include_dynamic=True)
