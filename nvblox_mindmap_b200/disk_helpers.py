"""On-disk formats of the path.

1. `NNNN.nvblox_vertex_features.zst` -- datagen's product: zstd level 1 of a pickled
   {'vertices': f16 [N,3], 'features': f16 [N,C_keep], 'channel_length': C_keep}
   (mindmap/mapping/helpers/nvblox_to_disk_helpers.py:22-67; read back by data_loading/dataset.py:410-415
   through a streaming zstd reader).  The cloud comes from the fused device export (output_helpers.py).
   The python `zstandard` package the reference uses is not in this image; the same frames are produced with
   libzstd.so.1 (the library `zstandard` wraps) bound with ctypes: ZSTD_compress writes the content size into the
   frame header exactly like ZstdCompressor(level=1).compress().

2. `.nvblx` layer cakes -- nvblox's sqlite container (NB/src/map_saving/serializer.cpp:117-205,320-370,
   layer_type_register + common_types_impl.h:75-86, block_serialization_impl.h:19-45): per layer a
   `<name>_metadata(param_name, value_string, value_int, value_float)` table holding `block_size` and `type`, and a
   `<name>_data(index_x, index_y, index_z, data BLOB)` table whose blobs are the raw `voxels` array of each block:
   TSDF 512 x (float distance, float weight); colour 512 x (r, g, b, pad, float weight); feature
   512 x (C halves + half weight) = 512 x (C+1) halves, packed.  PARITY: schema and blob layout follow the cited
   sources; no .nvblx written by a real nvblox build is available here to cross-check (parity unpinned beyond the
   round trip and the schema test).
"""
import ctypes as C
import ctypes.util
import os
import pickle
import sqlite3
from typing import Optional

import numpy as np
import torch

VERTEX_FEATURES_FILE_NAME = 'nvblox_vertex_features.zst'      # data_loading/item_names.py:12
ZSTD_MAGIC = b'\x28\xb5\x2f\xfd'

_zstd = None


def _libzstd():
    global _zstd
    if _zstd is None:
        name = ctypes.util.find_library('zstd') or 'libzstd.so.1'
        L = C.CDLL(name)
        L.ZSTD_compressBound.restype = C.c_size_t
        L.ZSTD_compressBound.argtypes = [C.c_size_t]
        L.ZSTD_compress.restype = C.c_size_t
        L.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.ZSTD_decompress.restype = C.c_size_t
        L.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ZSTD_getFrameContentSize.restype = C.c_ulonglong
        L.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
        L.ZSTD_isError.restype = C.c_uint
        L.ZSTD_isError.argtypes = [C.c_size_t]
        L.ZSTD_getErrorName.restype = C.c_char_p
        L.ZSTD_getErrorName.argtypes = [C.c_size_t]
        _zstd = L
    return _zstd


def zstd_compress(data: bytes, level: int = 1) -> bytes:
    L = _libzstd()
    cap = L.ZSTD_compressBound(len(data))
    dst = C.create_string_buffer(cap)
    n = L.ZSTD_compress(dst, cap, data, len(data), level)
    if L.ZSTD_isError(n):
        raise RuntimeError('ZSTD_compress: ' + L.ZSTD_getErrorName(n).decode())
    return dst.raw[:n]


def zstd_decompress(blob: bytes) -> bytes:
    L = _libzstd()
    size = L.ZSTD_getFrameContentSize(blob, len(blob))
    if size in (2 ** 64 - 1, 2 ** 64 - 2):
        raise RuntimeError('zstd frame without a content size (or not a zstd frame)')
    dst = C.create_string_buffer(max(1, size))
    n = L.ZSTD_decompress(dst, size, blob, len(blob))
    if L.ZSTD_isError(n):
        raise RuntimeError('ZSTD_decompress: ' + L.ZSTD_getErrorName(n).decode())
    return dst.raw[:n]


# ---- 1. feature point cloud ----------------------------------------------------------------------------------
def write_vertex_features(path: str, vertices: torch.Tensor, features: torch.Tensor) -> None:
    """The reference's container for one frame's cloud (nvblox_to_disk_helpers.py:53-65)."""
    assert vertices.shape[0] == features.shape[0]
    assert vertices.shape[1] == 3
    pc_ob = {
        'vertices': vertices.to(torch.float16).cpu(),
        'features': features.to(torch.float16).cpu(),
        'channel_length': features.shape[1],
    }
    with open(path, 'wb') as outfile:
        outfile.write(zstd_compress(pickle.dumps(pc_ob, protocol=pickle.HIGHEST_PROTOCOL), level=1))


def read_vertex_features(path: str) -> dict:
    """Dataset.unpickle_zst (data_loading/dataset.py:410-415)."""
    with open(path, 'rb') as f:
        return pickle.loads(zstd_decompress(f.read()))


def save_feature_mesh_to_disk(mapper, mapping_config, num_excess_features: int, frame_index: int,
                              save_directory: str, include_dynamic: bool, mapper_id: int = 0):
    """nvblox_to_disk_helpers.py:22-67 (same arguments; `mapper_id` defaults to MAPPER_TO_ID.STATIC = 0)."""
    from nvblox_mindmap_b200.output_helpers import get_vertices_and_features
    assert not include_dynamic, 'Dynamics are not supported for mesh encoding yet.'
    vertices, features, _ = get_vertices_and_features(mapper, mapper_id, mapping_config, remove_zero_features=True,
                                                      num_excess_features=num_excess_features,
                                                      sample_vertices=False)
    write_vertex_features(os.path.join(save_directory, f'{frame_index:04d}.{VERTEX_FEATURES_FILE_NAME}'),
                          vertices, features)
    return vertices, features


def save_serialized_nvblox_map_to_disk(mapper, save_directory: str, index: int, include_dynamic: bool) -> None:
    """nvblox_to_disk_helpers.py:70-93."""
    mapper.save_map(os.path.join(save_directory, f'{index:04d}.nvblox_map_static.nvblx'), 0)
    if include_dynamic:
        mapper.save_map(os.path.join(save_directory, f'{index:04d}.nvblox_map_dynamic.nvblx'), 1)


# ---- 2. .nvblx layer cake --------------------------------------------------------------------------------------
# every serialisable layer of the reference's cake gets its two tables, also when empty (mapper.cpp:48-51,
# common_types_impl.h:75-86); only tsdf / colour / feature ever hold blocks on this path
NVBLX_LAYERS = ('tsdf_layer', 'esdf_layer', 'color_layer', 'occupancy_layer', 'feature_layer', 'freespace_layer')


def _create_layer_tables(db: sqlite3.Connection, name: str, block_size: float) -> None:
    db.execute(f'CREATE TABLE {name}_metadata(param_name TEXT PRIMARY KEY UNIQUE NOT NULL,'
               'value_string TEXT,value_int INT,value_float FLOAT);')
    db.execute(f'CREATE TABLE {name}_data(index_x INT NOT NULL,index_y INT NOT NULL,index_z INT NOT NULL,'
               'data BLOB,PRIMARY KEY(index_x, index_y, index_z));')
    db.execute(f"INSERT INTO {name}_metadata (param_name, value_string) VALUES('type','{name}');")
    # std::to_string(float): fixed notation, six decimals (serializer.cpp:263-271)
    db.execute(f"INSERT INTO {name}_metadata (param_name, value_float) VALUES ('block_size','{block_size:f}');")


def save_map(mapper, map_fname: str, mapper_id: int) -> bool:
    """Mapper::saveLayerCake -> io::writeLayerCakeToFile (layer_cake_io.cpp:22-37): truncate and rewrite."""
    if os.path.exists(map_fname):
        os.remove(map_fname)
    voxel_size = np.float32(mapper._voxel_sizes[mapper_id])
    block_size = float(np.float32(voxel_size * np.float32(8)))
    C_feat = mapper._feature_channels
    db = sqlite3.connect(map_fname)
    try:
        views = {'tsdf_layer': mapper.tsdf_layer_view(mapper_id), 'color_layer': mapper.color_layer_view(mapper_id),
                 'feature_layer': mapper.feature_layer_view(mapper_id)}
        for name in NVBLX_LAYERS:
            _create_layer_tables(db, name, block_size)
            layer = views.get(name)
            if layer is None:
                continue
            rows = []
            for idx in layer.get_all_block_indices():
                x, y, z = (int(v) for v in idx)
                blk = layer.get_block_at_index(idx)
                if blk is None:
                    continue
                if name == 'color_layer':           # the full 8-byte ColorVoxel records, not just the RGB view
                    from nvblox_mindmap_b200.torch_interop import device_view
                    blk = device_view(blk.data_ptr(), (512, 8), torch.uint8, mapper._device, owner=mapper)
                elif name == 'feature_layer':       # packed (C+1)-half rows
                    blk = blk.contiguous()
                    assert blk.shape[-1] == C_feat + 1
                rows.append((x, y, z, blk.cpu().numpy().tobytes()))
            # one transaction per layer (serializer.cpp:152-161); python's sqlite3 opens it implicitly
            db.executemany(f'INSERT INTO {name}_data (index_x, index_y, index_z, data) VALUES (?,?,?,?)', rows)
            db.commit()
        db.commit()
    finally:
        db.close()
    return True


def load_map(mapper, filename: str, mapper_id: int) -> bool:
    """Mapper::loadMap (mapper.cpp:859-900): replace the map's layers by the file's, then re-mesh everything."""
    if not os.path.exists(filename):
        return False
    db = sqlite3.connect(f'file:{filename}?mode=ro', uri=True)
    try:
        names = [r[0][:-len('_metadata')] for r in db.execute(
            "SELECT name FROM sqlite_master WHERE type='table' AND name NOT LIKE 'sqlite_%' "
            "AND name LIKE '%_metadata';")]
        if 'tsdf_layer' not in names:
            return False                                       # "No TSDF layer could be loaded from file"
        bs = db.execute("SELECT value_float FROM tsdf_layer_metadata WHERE param_name = 'block_size';").fetchone()
        voxel_size = float(np.float32(bs[0]) / np.float32(8))
        if abs(voxel_size - mapper._voxel_sizes[mapper_id]) > 1e-7 * max(1.0, voxel_size):
            raise ValueError(f'{filename} holds a {voxel_size} m map; this mapper_id was created with '
                             f'{mapper._voxel_sizes[mapper_id]} m (voxel sizes are fixed at construction here)')
        mapper.clear(mapper_id)
        C_feat = mapper._feature_channels
        spec = {'tsdf_layer': (mapper.tsdf_layer_view(mapper_id), np.float32, (8, 8, 8, 2)),
                'color_layer': (mapper.color_layer_view(mapper_id), np.uint8, (512, 8)),
                'feature_layer': (mapper.feature_layer_view(mapper_id), np.float16, (8, 8, 8, C_feat + 1))}
        dev = f'cuda:{mapper._device}'
        for name in ('tsdf_layer', 'color_layer', 'feature_layer'):
            if name not in names:
                continue
            layer, dtype, shape = spec[name]
            for x, y, z, blob in db.execute(f'SELECT index_x,index_y,index_z,data FROM {name}_data;'):
                want = int(np.prod(shape)) * np.dtype(dtype).itemsize
                if len(blob) != want:
                    raise ValueError(f'{name} block ({x},{y},{z}) has {len(blob)} bytes, expected {want} '
                                     f'(feature length mismatch?)')
                idx = torch.tensor([x, y, z], dtype=torch.int32)
                layer.allocate_block_at_index(idx)
                view = layer.get_block_at_index(idx)
                data = torch.from_numpy(np.frombuffer(blob, dtype=dtype).reshape(shape).copy()).to(dev)
                if name == 'color_layer':
                    from nvblox_mindmap_b200.torch_interop import device_view
                    view = device_view(view.data_ptr(), (512, 8), torch.uint8, mapper._device, owner=mapper)
                view.copy_(data)
        from nvblox_mindmap_b200 import _capi
        _capi.check(_capi.load().nvbx_mark_all_dirty(mapper._handle, mapper_id, mapper._stream()))
        mapper.update_color_mesh(mapper_id)
        mapper.update_feature_mesh(mapper_id)
    finally:
        db.close()
    return True


def nvblx_summary(filename: str) -> dict:
    """{layer: (block_size, n_blocks, blob_bytes)} of a .nvblx file (inspection / tests; no GPU needed)."""
    db = sqlite3.connect(f'file:{filename}?mode=ro', uri=True)
    try:
        out = {}
        for (tname,) in db.execute("SELECT name FROM sqlite_master WHERE type='table' AND name LIKE '%_metadata';"):
            name = tname[:-len('_metadata')]
            bs = db.execute(f"SELECT value_float FROM {tname} WHERE param_name = 'block_size';").fetchone()
            n, nbytes = db.execute(f'SELECT COUNT(*), COALESCE(SUM(LENGTH(data)),0) FROM {name}_data;').fetchone()
            out[name] = (float(bs[0]) if bs else None, int(n), int(nbytes))
        return out
    finally:
        db.close()
