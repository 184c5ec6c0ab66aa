"""`nvblox_torch.visualization` (reference: nvblox_torch/visualization.py:98-131) -- the one helper mindmap imports
from it (`mindmap/paper/utils/utils.py`: get_voxel_mesh).  Viewer glue around block tensors, off the hot path; present
so that the reference's import lines resolve against the drop-in."""
from typing import Optional

import torch


def voxel_cubes(centers: torch.Tensor, voxel_size_m: float, colors: Optional[torch.Tensor] = None):
    """Cubes of edge 0.9 x voxel_size_m (the reference's size) around `centers` [N, 3] as plain tensors:
    (vertices [8N, 3], triangles [12N, 3], vertex colours [8N, 3] or None)."""
    assert centers.dim() == 2
    assert centers.shape[-1] == 3
    if colors is not None:
        assert colors.shape[-1] == 3
        assert colors.dim() == 2
        assert centers.shape[0] == colors.shape[0]
    h = 0.45 * float(voxel_size_m)
    dev = centers.device
    corners = torch.tensor([[-h, -h, -h], [h, -h, -h], [h, h, -h], [-h, h, -h],
                            [-h, -h, h], [h, -h, h], [h, h, h], [-h, h, h]], device=dev, dtype=centers.dtype)
    faces = torch.tensor([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4],
                          [1, 2, 6], [1, 6, 5], [2, 3, 7], [2, 7, 6], [3, 0, 4], [3, 4, 7]], device=dev,
                         dtype=torch.int64)
    n = centers.shape[0]
    verts = (centers[:, None, :] + corners[None, :, :]).reshape(-1, 3)
    tris = (faces[None, :, :] + 8 * torch.arange(n, device=dev)[:, None, None]).reshape(-1, 3)
    vcol = None if colors is None else colors[:, None, :].expand(n, 8, 3).reshape(-1, 3)
    return verts, tris, vcol


def get_voxel_mesh(centers: torch.Tensor, voxel_size_m: float, colors: Optional[torch.Tensor] = None):
    """An Open3D TriangleMesh of the voxel grid, as the reference returns.  Needs Open3D (a viewer dependency that is
    not part of this image); `voxel_cubes` gives the same geometry as tensors."""
    verts, tris, vcol = voxel_cubes(centers, voxel_size_m, colors)
    try:
        import open3d as o3d
    except ImportError as e:
        raise ImportError('get_voxel_mesh returns an Open3D mesh: install open3d, or use '
                          'nvblox_torch.visualization.voxel_cubes for plain tensors') from e
    mesh = o3d.geometry.TriangleMesh()
    mesh.vertices = o3d.utility.Vector3dVector(verts.detach().cpu().double().numpy())
    mesh.triangles = o3d.utility.Vector3iVector(tris.detach().cpu().numpy().astype('int32'))
    if vcol is not None:
        mesh.vertex_colors = o3d.utility.Vector3dVector(vcol.detach().cpu().double().numpy())
    mesh.compute_vertex_normals()
    return mesh
