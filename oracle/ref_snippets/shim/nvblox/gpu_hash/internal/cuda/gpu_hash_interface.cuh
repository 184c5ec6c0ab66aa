// Stand-in for NB/include/nvblox/gpu_hash/internal/cuda/gpu_hash_interface.cuh (a wrapper around
// stdgpu::unordered_map; stdgpu is fetched by nvblox's CMake and is absent offline).
// TEST INFRASTRUCTURE ONLY.  The reference's device code only ever calls find() / end() and reads
// it->second (gpu_indexing.cuh:29-64), so a dense lookup table over the map's index AABB gives the same
// answers; no arithmetic lives here.
#pragma once
#include <cuda_runtime.h>
#include <thrust/pair.h>

#include <vector>

#include "nvblox/core/hash.h"
#include "nvblox/core/types.h"

namespace nvblox {

template <typename BlockType>
using ConstIndexBlockPair = thrust::pair<const Index3D, BlockType*>;

template <typename BlockType>
struct Index3DDeviceHashMapType {
  typedef thrust::pair<Index3D, BlockType*> Entry;
  Entry* cells = nullptr;   // device, prod(size) entries, second == nullptr when absent
  int mn[3] = {0, 0, 0};
  int size[3] = {0, 0, 0};

  __device__ const Entry* end() const { return nullptr; }
  __device__ const Entry* find(const Index3D& idx) const {
    const int x = idx.x() - mn[0], y = idx.y() - mn[1], z = idx.z() - mn[2];
    if (x < 0 || y < 0 || z < 0 || x >= size[0] || y >= size[1] || z >= size[2]) return nullptr;
    const Entry* e = cells + ((size_t)z * size[1] + y) * size[0] + x;
    return e->second ? e : nullptr;
  }

  // host: entry i of idx (int[n][3]) is stored at storage + i
  int build(const int* idx, int n, BlockType* storage) {
    if (n <= 0) {
      size[0] = size[1] = size[2] = 0;
      return 0;
    }
    int mx[3];
    for (int k = 0; k < 3; ++k) mn[k] = mx[k] = idx[k];
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) {
        if (idx[3 * i + k] < mn[k]) mn[k] = idx[3 * i + k];
        if (idx[3 * i + k] > mx[k]) mx[k] = idx[3 * i + k];
      }
    for (int k = 0; k < 3; ++k) size[k] = mx[k] - mn[k] + 1;
    const size_t total = (size_t)size[0] * size[1] * size[2];
    std::vector<Entry> host(total);
    for (size_t c = 0; c < total; ++c) host[c].second = nullptr;
    for (int i = 0; i < n; ++i) {
      const int x = idx[3 * i] - mn[0], y = idx[3 * i + 1] - mn[1], z = idx[3 * i + 2] - mn[2];
      Entry& e = host[((size_t)z * size[1] + y) * size[0] + x];
      e.first = Index3D(idx[3 * i], idx[3 * i + 1], idx[3 * i + 2]);
      e.second = storage + i;
    }
    if (cudaMalloc(&cells, total * sizeof(Entry)) != cudaSuccess) return -1;
    if (cudaMemcpy(cells, host.data(), total * sizeof(Entry), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    return 0;
  }
  void destroy() {
    cudaFree(cells);
    cells = nullptr;
  }
};

}  // namespace nvblox
