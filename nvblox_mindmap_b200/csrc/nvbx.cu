// nvbx.cu -- host orchestration and the C ABI of libnvbx.so (include/nvbx_c_api.h).
//
// The host side does what cannot be done on the device: argument checks, the viewpoint-cache decision
// (pose/intrinsics comparison, view_calculator.cu:472-541), the view AABB (8 frustum corners), arena
// sizing, and enqueueing a FIXED sequence of launches per call.  It never reads a result back during
// frame integration or decay; the only synchronising calls are update_feature_mesh (vertex total),
// the layer-view getters and get_counters.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <functional>
#include <vector>

#include "../../include/nvbx_c_api.h"
#include "nvbx_export.cuh"
#include "nvbx_mesh.cuh"
#include "nvbx_upsample.cuh"

using namespace nvbx;

namespace {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return fail(_e == cudaErrorMemoryAllocation ? NVBX_ERR_OUT_OF_MEMORY : NVBX_ERR_CUDA,         \
                  "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);      \
  } while (0)

// Per-kernel live timing (tuning / profiles): mode 2 brackets EVERY launch with a CUDA event pair on the
// launching stream; nvbx_kernel_timing_report() sums them per kernel name.  Off (0) in normal operation.
struct LaunchRec {
  const char* name;
  cudaEvent_t a, b;
};
std::atomic<int> g_timing_mode{0};
std::mutex g_launch_recs_mu;  // launches may come from the batch entry's worker threads
std::vector<LaunchRec> g_launch_recs;
thread_local long long t_last_rec = -1;  // index of the record this thread opened last
inline void launch_rec_begin(const char* name, cudaStream_t stream) {
  t_last_rec = -1;
  if (g_timing_mode.load(std::memory_order_relaxed) < 2) return;
  LaunchRec r{name, nullptr, nullptr};
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, stream);
  std::lock_guard<std::mutex> g(g_launch_recs_mu);
  t_last_rec = (long long)g_launch_recs.size();
  g_launch_recs.push_back(r);
}
inline void launch_rec_end(cudaStream_t stream) {
  if (t_last_rec < 0) return;
  cudaEvent_t b = nullptr;
  {
    std::lock_guard<std::mutex> g(g_launch_recs_mu);
    if (t_last_rec < (long long)g_launch_recs.size()) b = g_launch_recs[(size_t)t_last_rec].b;
  }
  if (b) cudaEventRecord(b, stream);
  t_last_rec = -1;
}

// All kernels are launched with programmatic stream serialization (see pdl_prologue() in
// nvbx_kernels.cuh); NVBX_PDL=0 falls back to plain stream order.
bool use_pdl() {
  static const bool v = [] {
    const char* e = getenv("NVBX_PDL");
    return !(e && e[0] == '0');
  }();
  return v;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

#define LAUNCH(kernel, grid, block, smem, stream, ...)                                              \
  do {                                                                                              \
    launch_rec_begin(#kernel, (stream));                                                            \
    cudaError_t _le = launch_kernel(kernel, dim3(grid), dim3(block), (size_t)(smem), (stream), __VA_ARGS__); \
    launch_rec_end((stream));                                                                       \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                             \
    if (_le != cudaSuccess)                                                                         \
      return fail(NVBX_ERR_CUDA, "launch of %s failed: %s (%s:%d)", #kernel, cudaGetErrorString(_le), __FILE__, \
                  __LINE__);                                                                        \
  } while (0)

template <typename T>
struct DevBuf {  // grow-only device scratch buffer
  T* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n, cudaStream_t stream, bool keep = false) {
    if (n <= cap) return NVBX_OK;
    size_t ncap = std::max(n, cap + cap / 2 + 64);
    T* np = nullptr;
    CUDA_TRY(cudaMalloc(&np, ncap * sizeof(T)));
    if (p) {
      if (keep) CUDA_TRY(cudaMemcpyAsync(np, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));  // earlier kernels may still use the old buffer
      cudaFree(p);
    }
    p = np;
    cap = ncap;
    return NVBX_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// One viewpoint-cache entry: pose + camera on the host, the block-index list it produced on the device.
struct CacheEntry {
  Pose T;
  Cam cam;
  DevBuf<int3> idx;     // raycast cache only: the block-index list (depends on the depth image)
  int* d_count = nullptr;
  int bound = 0;        // host upper bound of *d_count (cells of the AABB it was built from)
  ViewGrid grid{};      // planes cache only: the frustum AABB the in-view test enumerates
};

struct ViewCache {  // deque semantics of ViewpointCache (kMaxCacheSize = 2), view_calculator.cu:519-541
  std::deque<CacheEntry*> live;
  std::vector<std::unique_ptr<CacheEntry>> pool;
  CacheEntry* lookup(const Pose& T, const Cam& cam) {
    for (CacheEntry* e : live)
      if (cameras_equal(cam, e->cam, T, e->T)) return e;
    return nullptr;
  }
  // entry to (over)write for a new result: the one that would be evicted, or a fresh one
  int acquire(CacheEntry** out, bool device_list = true) {
    if (live.size() == 2) {
      *out = live.back();
      live.pop_back();
      return NVBX_OK;
    }
    pool.emplace_back(new CacheEntry());
    CacheEntry* e = pool.back().get();
    if (device_list) CUDA_TRY(cudaMalloc(&e->d_count, sizeof(int)));
    *out = e;
    return NVBX_OK;
  }
  void store(CacheEntry* e) { live.push_front(e); }
  void release() {
    for (auto& e : pool) {
      e->idx.release();
      if (e->d_count) cudaFree(e->d_count);
    }
    pool.clear();
    live.clear();
  }
};

struct Map {
  float voxel_size = 0, block_size = 0;
  MapDev dev{};
  // slot table + hash (contiguous, re-allocated on growth)
  int slot_capacity = 0;
  int feat_capacity = 0;
  long long stray_ub = 0;      // blocks requested OUTSIDE the workspace grid (allocate_block_at_index, load_from_file):
                               // they live in the overflow hash and are not bounded by the workspace's cell count
  long long slot_used_ub = 0;  // host upper bounds of live ids (pessimistic between syncs)
  long long feat_used_ub = 0;
  std::vector<void*> tsdf_slabs, feat_slabs, color_slabs;
  float2** d_tsdf_table = nullptr;
  __half** d_feat_table = nullptr;
  uint2** d_color_table = nullptr;
  bool color_enabled = false;  // colour slabs exist (allocated by the first colour frame / colour block request)
  Ctrl* d_ctrl = nullptr;
  Ctrl* h_ctrl = nullptr;  // pinned mirror
  // the TSDF, colour and feature integrators of one Mapper share one raycasting and one planes viewpoint cache
  // (shareViewpointCaches, mapper.cpp:56-58)
  ViewCache raycast_cache, planes_cache;
  CacheEntry scratch_ray, scratch_planes;  // used when the cache is disabled
  DevBuf<unsigned> grid;   // view bitmap; all-zero between frames (k_view_compact_alloc cleans it)
  bool grid_dirty = false;  // a frame failed between marking and compaction
  unsigned* small_grid[3] = {nullptr, nullptr, nullptr};  // triple-buffered bitmap of small views (3 x kFusedBitmapWords)
  unsigned small_seq = 0;  // number of small-view depth frames enqueued (wraps); frame s marks small_grid[small_idx]
  int small_idx = 0;       // = (number of small-view depth frames) mod 3
  bool small_dirty = false;
  DevBuf<int> view_slots, cband_slots;
  DevBuf<int> band_slots2[kFrameRing], newfeat_slots2[kFrameRing];  // per feature frame, ring slot MapDev::fp
  int color_parity = 0;
  bool have_cband_list = false;
  DevBuf<FeatItem> items2[kFrameRing];  // feature work-item lists, ring slot MapDev::fp
  // Frame pipelining (nvbx_set_pipelining): the gather of feature frame i runs on `gstream`, ordered after the
  // frame's geometry kernel by `ev_trace`; `ev_gather[r]` marks the end of the last gather that used ring slot r (and
  // with it the last use of every buffer of that slot on `gstream`).
  cudaStream_t gstream = nullptr;
  // NVBX_PIPE_GEOM=2: the geometry kernel of frame i on a third, high-priority stream (behind the frame's trace / band
  // selection by `ev_trace`, ahead of its gather by `ev_geom`), so that the caller's stream goes straight on to frame
  // i + 1's raycast / TSDF update
  cudaStream_t cstream = nullptr;
  cudaEvent_t ev_geom = nullptr;
  cudaEvent_t ev_trace = nullptr, ev_gather[kFrameRing] = {};
  bool gather_pending[kFrameRing] = {};
  DevBuf<float> synth2[kFrameRing];  // synthetic depth images, ring slot MapDev::fp
  int synth_last = 0;       // which of them the last appearance frame used (debug hook)
  int synth_rows = 0, synth_cols = 0;
  // The synthetic depth image is a pure function of (pose, camera, truncation, TSDF contents): the colour and
  // feature frames of one mindmap step share all four, so the second one skips the sphere tracing.
  struct SynthKey {
    Pose T;
    Cam cam;
    float trunc = 0;
    int sub = 0;
    unsigned long long tsdf_version = 0;
    int buf = 0;  // which synth2[] holds the image
    bool valid = false;
  } synth_key;
  unsigned long long tsdf_version = 1;  // bumped by everything that may change TSDF voxels or the block set
  CacheEntry* last_depth_entry = nullptr;
  bool have_band_list = false;
  // mesh
  DevBuf<int> cnt_v, cnt_t, off_v, off_t;
  DevBuf<float> arena_v[2];
  DevBuf<__half> arena_f[2];
  DevBuf<int> arena_t[2];
  int arena_cur = 0;
  long long mesh_nv = 0, mesh_nt = 0;
  DevBuf<float> carena_v[2];  // colour mesh layer (its own geometry + uint8 rgb per vertex)
  DevBuf<uint8_t> carena_c[2];
  DevBuf<int> carena_t[2];
  int carena_cur = 0;
  long long cmesh_nv = 0, cmesh_nt = 0;
  // fused export post-processing (N2)
  DevBuf<uint8_t> exp_keep;
  DevBuf<int> exp_tile_cnt;
  DevBuf<long long> exp_tile_off, exp_idx;
  DevBuf<float> exp_v;
  DevBuf<__half> exp_f;
  long long* d_exp_total = nullptr;
  long long exp_count = 0;
  int exp_C_keep = 0;
  // host staging for nvbx_integrate_frame_host
  DevBuf<float> st_depth;
  DevBuf<__half> st_feat;
  DevBuf<unsigned> st_pixmap;  // one bit per feature pixel a host-frame work item reads (kept all-zero between frames)
  DevBuf<uint8_t> st_mask_d, st_mask_f;
  DevBuf<float> st_low;      // N4: [lh][lw][C] fp32 staged low-res feature map
  DevBuf<uint8_t> st_low_in;  // N4: raw upload of the host low-res map
  // misc single-value device scratch
  int* d_tmp_int = nullptr;
  unsigned long long* d_tmp_ptr = nullptr;
  DevBuf<int3> idx_out;
  DevBuf<unsigned long long> ptr_out;
  DevBuf<int> slot_out;
};

}  // namespace

// ---- asynchronous enqueue (nvbx_set_pipelining(m, 2)) -------------------------------------------------------
// The five launches + six event operations of a pipelined frame cost the calling thread ~20 us -- with the Python
// surface on top, more than the ~35 us the device needs per frame.  In this mode nvbx_integrate_depth /
// nvbx_integrate_features validate their arguments, copy them into a small FIFO and return; one worker thread per
// mapper issues the CUDA calls in order.  Every other entry point first waits until the FIFO has been issued (and
// reports an error a queued frame ran into), so the library's own calls keep their stream order; what the caller
// gives up is the order between a frame and the caller's OWN later work on the same stream (same contract as
// pipelining: inputs stay alive and unmodified until the next joining call).
struct AsyncJob {
  int kind;  // 0: depth frame, 1: feature frame
  int map_id;
  const void* image;
  int height, width, channels;
  const void* mask;
  float T[16];
  float fx, fy, cx, cy;
  void* stream;
};
class AsyncEnqueue {
 public:
  static constexpr int kCap = 8;  // frames the caller may run ahead of the worker
  explicit AsyncEnqueue(nvbx_mapper* m) : m_(m), th_([this] { loop(); }) {}
  ~AsyncEnqueue() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_work_.notify_all();
    th_.join();
  }
  void push(const AsyncJob& j) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_room_.wait(lk, [this] { return (int)q_.size() < kCap; });
    q_.push_back(j);
    ++outstanding_;
    lk.unlock();
    cv_work_.notify_one();
  }
  // Returns when everything pushed so far has been issued; the first error of a queued frame is returned ONCE.
  int flush(std::string* err) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [this] { return outstanding_ == 0; });
    const int rc = rc_;
    if (rc) *err = err_;
    rc_ = 0;
    err_.clear();
    return rc;
  }

 private:
  void loop();
  nvbx_mapper* m_;
  std::mutex mu_;
  std::condition_variable cv_work_, cv_room_, cv_done_;
  std::deque<AsyncJob> q_;
  int outstanding_ = 0;  // queued + being issued
  bool stop_ = false;
  int rc_ = 0;
  std::string err_;
  std::thread th_;  // last member: started when everything above is constructed
};

struct nvbx_mapper {
  std::unique_ptr<AsyncEnqueue> async;  // non-null while asynchronous enqueue is on
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events[2];
  int device = 0;
  int C = 0;
  int sm_count = 148;
  int host_fetch_mode = NVBX_HOST_FETCH_SPARSE;
  bool pipelining = false;  // nvbx_set_pipelining
  nvbx_params params{};
  std::vector<std::unique_ptr<Map>> maps;
};

namespace {

constexpr long long kMaxWorkspaceGridCells = 1LL << 24;  // 64 MiB of slot ids
constexpr int kFeatureChunkBlocks = 8192;                // band blocks per geometry/gather pass (64 MiB of items)

int persistent_grid(const nvbx_mapper* m, int ctas_per_sm) { return m->sm_count * ctas_per_sm; }

// Test aid (NVBX_POISON_ARENAS=1, set by tests/conftest.py): fresh slabs are filled with garbage so that a
// missing zero-initialisation of a new block cannot hide behind cudaMalloc handing out zero pages.
bool poison_arenas() {
  static const bool v = [] {
    const char* e = getenv("NVBX_POISON_ARENAS");
    return e && e[0] == '1';
  }();
  return v;
}

int upload_slab_tables(Map& mp, cudaStream_t stream) {
  if (!mp.tsdf_slabs.empty())
    CUDA_TRY(cudaMemcpyAsync(mp.d_tsdf_table, mp.tsdf_slabs.data(), mp.tsdf_slabs.size() * sizeof(void*),
                             cudaMemcpyHostToDevice, stream));
  if (!mp.feat_slabs.empty())
    CUDA_TRY(cudaMemcpyAsync(mp.d_feat_table, mp.feat_slabs.data(), mp.feat_slabs.size() * sizeof(void*),
                             cudaMemcpyHostToDevice, stream));
  if (!mp.color_slabs.empty())
    CUDA_TRY(cudaMemcpyAsync(mp.d_color_table, mp.color_slabs.data(), mp.color_slabs.size() * sizeof(void*),
                             cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));  // the host vectors may reallocate later
  return NVBX_OK;
}

template <typename T>
int regrow(T** p, size_t old_n, size_t new_n, cudaStream_t stream) {
  T* np = nullptr;
  CUDA_TRY(cudaMalloc(&np, new_n * sizeof(T)));
  if (*p && old_n) CUDA_TRY(cudaMemcpyAsync(np, *p, old_n * sizeof(T), cudaMemcpyDeviceToDevice, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  if (*p) cudaFree(*p);
  *p = np;
  return NVBX_OK;
}

// Colour payload slabs mirror the TSDF slabs one to one (payload id == slot id).
int grow_color_slabs(Map& mp, cudaStream_t stream) {
  const size_t bytes = (size_t)(1 << kTsdfSlabShift) * kVoxelsPerBlock * sizeof(uint2);
  while (mp.color_slabs.size() < mp.tsdf_slabs.size()) {
    void* s = nullptr;
    CUDA_TRY(cudaMalloc(&s, bytes));
    if (poison_arenas()) CUDA_TRY(cudaMemsetAsync(s, 0x7b, bytes, stream));
    mp.color_slabs.push_back(s);
  }
  return NVBX_OK;
}

// Grow the slot table (and TSDF slabs, hash) to hold `new_cap` block indices.
// ---- frame pipelining (nvbx_set_pipelining) -------------------------------------------------------------
// The gather of feature frame i may run on the map's own stream while the caller's stream already carries the
// depth path of frame i + 1 (raycast, TSDF update, sphere tracing + band selection, geometry): those kernels are
// latency-bound and touch nothing the memory-bound gather reads or writes (Ctrl counters and the item list are
// double-buffered by MapDev::fp, voxel weights are written by the geometry kernel).
int env_int(const char* name, int dflt);
std::atomic<long long> g_pipe_waits{0}, g_pipe_wait_ns{0};  // nvbx_pipeline_wait_stats
int pipeline_init(Map& mp) {
  if (mp.gstream) return NVBX_OK;
  int lo = 0, hi = 0;
  CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  static const int prio = env_int("NVBX_PIPE_PRIO", 0);  // tuning knob: 0 lowest (the short kernels go first), 1 highest
  CUDA_TRY(cudaStreamCreateWithPriority(&mp.gstream, cudaStreamNonBlocking, prio ? hi : lo));
  static const int gprio = env_int("NVBX_PIPE_GEOM_PRIO", 1);  // tuning knob: the geometry stream's priority
  CUDA_TRY(cudaStreamCreateWithPriority(&mp.cstream, cudaStreamNonBlocking, gprio ? hi : lo));
  CUDA_TRY(cudaEventCreateWithFlags(&mp.ev_geom, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&mp.ev_trace, cudaEventDisableTiming));
  for (int r = 0; r < kFrameRing; ++r) CUDA_TRY(cudaEventCreateWithFlags(&mp.ev_gather[r], cudaEventDisableTiming));
  return NVBX_OK;
}
// Order `stream` after every gather still in flight: called by whatever reads or frees feature blocks.
int pipeline_join(Map& mp, cudaStream_t stream) {
  for (int p = 0; p < kFrameRing; ++p)
    if (mp.gather_pending[p]) {
      CUDA_TRY(cudaStreamWaitEvent(stream, mp.ev_gather[p], 0));
      mp.gather_pending[p] = false;
    }
  return NVBX_OK;
}
// Host-side wait (before device memory the gather uses is re-allocated or freed).
int pipeline_drain(Map& mp) {
  bool any = false;
  for (int r = 0; r < kFrameRing; ++r) any |= mp.gather_pending[r];
  if (mp.gstream && any) CUDA_TRY(cudaStreamSynchronize(mp.gstream));
  return NVBX_OK;
}
// The frame about to be enqueued re-uses ring slot mp.dev.fp (item list, band lists, synthetic depth image, their
// counters): the gather that read them last -- kFrameRing feature frames ago -- must be COMPLETE, because nothing on
// the caller's stream orders the new kernels behind it.  It practically always is (the ring is four frames deep);
// when the gather stream really lags that far, the host waits here rather than the device.
int pipeline_slot_ready(Map& mp) {
  const int r = mp.dev.fp;
  if (!mp.gather_pending[r]) return NVBX_OK;
  const cudaError_t q = cudaEventQuery(mp.ev_gather[r]);
  if (q == cudaErrorNotReady) {
    const auto t0 = std::chrono::steady_clock::now();
    CUDA_TRY(cudaEventSynchronize(mp.ev_gather[r]));
    g_pipe_waits.fetch_add(1, std::memory_order_relaxed);
    g_pipe_wait_ns.fetch_add(
        std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(),
        std::memory_order_relaxed);
  } else if (q != cudaSuccess) {
    CUDA_TRY(q);
  }
  return NVBX_OK;  // gather_pending[r] stays set: a later join still has to order the CALLER'S stream behind it
}

int grow_slots(nvbx_mapper* m, Map& mp, int new_cap, cudaStream_t stream) {
  if (int prc = pipeline_drain(mp)) return prc;
  CUDA_TRY(cudaStreamSynchronize(stream));
  const int old = mp.slot_capacity;
  new_cap = ((new_cap + (1 << kTsdfSlabShift) - 1) >> kTsdfSlabShift) << kTsdfSlabShift;
  if ((new_cap >> kTsdfSlabShift) > kMaxSlabs) return fail(NVBX_ERR_OUT_OF_MEMORY, "TSDF arena limit reached");
  int rc;
  if ((rc = regrow(&mp.dev.blk_index, old, new_cap, stream))) return rc;
  if ((rc = regrow(&mp.dev.blk_layers, old, new_cap, stream))) return rc;
  if ((rc = regrow(&mp.dev.blk_feat, old, new_cap, stream))) return rc;
  if ((rc = regrow(&mp.dev.blk_dirty, old, new_cap, stream))) return rc;
  if ((rc = regrow(&mp.dev.blk_mesh, old, new_cap, stream))) return rc;
  if ((rc = regrow(&mp.dev.blk_cmesh, old, new_cap, stream))) return rc;
  if ((rc = regrow(&mp.dev.slot_free, old, new_cap, stream))) return rc;
  CUDA_TRY(cudaMemsetAsync(mp.dev.blk_layers + old, 0, new_cap - old, stream));
  while ((int)mp.tsdf_slabs.size() < (new_cap >> kTsdfSlabShift)) {
    void* s = nullptr;
    const size_t bytes = (size_t)(1 << kTsdfSlabShift) * kVoxelsPerBlock * sizeof(float2);
    CUDA_TRY(cudaMalloc(&s, bytes));
    if (poison_arenas()) CUDA_TRY(cudaMemsetAsync(s, 0x7b, bytes, stream));
    mp.tsdf_slabs.push_back(s);
  }
  if (mp.color_enabled && (rc = grow_color_slabs(mp, stream))) return rc;
  if ((rc = upload_slab_tables(mp, stream))) return rc;
  mp.slot_capacity = new_cap;
  mp.dev.slot_capacity = new_cap;
  // hash: next pow2 >= 4 * capacity
  unsigned hcap = 1024;
  while (hcap < 4u * (unsigned)new_cap) hcap <<= 1;
  if (mp.dev.ws_cells > 0) {  // grid-indexed map: the hash only holds strays -- sized from their number, load <= 0.25
    hcap = 1u << 16;
    while ((long long)hcap < 4 * mp.stray_ub) hcap <<= 1;
    if (mp.dev.hash_mask && mp.dev.hash_mask + 1 > hcap) hcap = mp.dev.hash_mask + 1;  // never shrinks
  }
  if (hcap != mp.dev.hash_mask + 1 || mp.dev.keys == nullptr) {
    if (mp.dev.keys) cudaFree(mp.dev.keys);
    if (mp.dev.vals) cudaFree(mp.dev.vals);
    mp.dev.keys = nullptr;
    mp.dev.vals = nullptr;
    CUDA_TRY(cudaMalloc(&mp.dev.keys, (size_t)hcap * sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&mp.dev.vals, (size_t)hcap * sizeof(int)));
    mp.dev.hash_mask = hcap - 1;
    LAUNCH(k_hash_clear, persistent_grid(m, 4), 256, 0, stream, mp.dev, 1);
    LAUNCH(k_hash_reinsert, persistent_grid(m, 4), 256, 0, stream, mp.dev, 1);
  }
  return NVBX_OK;
}

int grow_feats(nvbx_mapper* m, Map& mp, int new_cap, cudaStream_t stream) {
  if (int prc = pipeline_drain(mp)) return prc;
  CUDA_TRY(cudaStreamSynchronize(stream));
  const int old = mp.feat_capacity;
  new_cap = ((new_cap + (1 << kFeatSlabShift) - 1) >> kFeatSlabShift) << kFeatSlabShift;
  if ((new_cap >> kFeatSlabShift) > kMaxSlabs) return fail(NVBX_ERR_OUT_OF_MEMORY, "feature arena limit reached");
  int rc;
  if ((rc = regrow(&mp.dev.feat_free, old, new_cap, stream))) return rc;
  const size_t slab_bytes = (size_t)(1 << kFeatSlabShift) * kVoxelsPerBlock * (size_t)mp.dev.row * sizeof(__half);
  while ((int)mp.feat_slabs.size() < (new_cap >> kFeatSlabShift)) {
    void* s = nullptr;
    CUDA_TRY(cudaMalloc(&s, slab_bytes));
    if (poison_arenas()) CUDA_TRY(cudaMemsetAsync(s, 0x7b, slab_bytes, stream));
    mp.feat_slabs.push_back(s);
  }
  if ((rc = upload_slab_tables(mp, stream))) return rc;
  mp.feat_capacity = new_cap;
  mp.dev.feat_capacity = new_cap;
  return NVBX_OK;
}

// First use of the colour layer: allocate its slabs next to the TSDF slabs.
int enable_color(Map& mp, cudaStream_t stream) {
  if (mp.color_enabled) return NVBX_OK;
  int rc;
  if ((rc = grow_color_slabs(mp, stream))) return rc;
  mp.color_enabled = true;
  return upload_slab_tables(mp, stream);
}

int read_ctrl(Map& mp, cudaStream_t stream) {
  CUDA_TRY(cudaMemcpyAsync(mp.h_ctrl, mp.d_ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  if (mp.h_ctrl->overflow) return fail(NVBX_ERR_OUT_OF_MEMORY, "internal: a block arena overflowed on the device");
  return NVBX_OK;
}

// Workspace cell count (upper bound of distinct block indices when a bounding box is set), or -1.
long long workspace_cells(const nvbx_mapper* m, const Map& mp) {
  if (m->params.workspace_bounds_type != NVBX_WORKSPACE_BOUNDING_BOX) return -1;
  long long n = 1;
  for (int i = 0; i < 3; ++i) {
    const float lo = m->params.workspace_min[i], hi = m->params.workspace_max[i];
    if (!(hi >= lo)) return 0;
    const long long c = (long long)std::floor(hi / mp.block_size) - (long long)std::floor(lo / mp.block_size) + 1;
    n *= std::max(1LL, c);
    if (n > (1LL << 40)) return -1;
  }
  return n;
}

// Make sure `need` more block indices can be allocated by the next kernel without the host knowing
// how many really will be (see DESIGN.md "Arena sizing without read-backs").
int ensure_slots(nvbx_mapper* m, Map& mp, long long need, cudaStream_t stream) {
  const long long ws0 = workspace_cells(m, mp);
  const long long ws = ws0 >= 0 ? ws0 + mp.stray_ub : ws0;  // strays outside the box are extra
  auto bound = [&](long long v) { return ws >= 0 ? std::min(v, ws) : v; };
  if (bound(mp.slot_used_ub + need) <= mp.slot_capacity) {
    mp.slot_used_ub = bound(mp.slot_used_ub + need);
    return NVBX_OK;
  }
  int rc = read_ctrl(mp, stream);  // exact figure
  if (rc) return rc;
  mp.slot_used_ub = (long long)mp.h_ctrl->slot_high - mp.h_ctrl->slot_free_top;
  const long long want = bound(mp.slot_used_ub + need);
  if (want > mp.slot_capacity) {
    long long ncap = std::max(want, (long long)(mp.slot_capacity * std::max(1.5f, m->params.expansion_factor)));
    if (ws >= 0) ncap = std::min(ncap, std::max(ws, want));
    if (ncap > (1LL << 30)) return fail(NVBX_ERR_OUT_OF_MEMORY, "map too large (%lld blocks)", ncap);
    if ((rc = grow_slots(m, mp, (int)ncap, stream))) return rc;
  }
  mp.slot_used_ub = want;
  return NVBX_OK;
}
bool slots_fit(const nvbx_mapper* m, const Map& mp, long long need) {
  const long long ws = workspace_cells(m, mp);
  const long long v = mp.slot_used_ub + need;
  return (ws >= 0 ? std::min(v, ws) : v) <= mp.slot_capacity;
}
// Feature blocks are ~1 MB each (512 x (C + 8) halves), so the pessimistic bound "every candidate block of the
// view gets one" is only affordable for small maps: past kFeatPessimisticBytes the arena grows by the EXACT
// number of feature blocks the frame will allocate, counted on the device by `count_new` (k_band_count) and read
// back -- two synchronisations per frame, paid only by maps of that size while their bound does not fit.
constexpr long long kFeatPessimisticBytes = 8LL << 30;
constexpr long long kFeatExactSlackBytes = 1LL << 30;

int ensure_feats(nvbx_mapper* m, Map& mp, long long need, cudaStream_t stream,
                 const std::function<int()>& count_new = nullptr) {
  need = std::min(need, mp.slot_used_ub);  // a feature block needs a TSDF block
  auto bound = [&](long long v) { return std::min(v, mp.slot_used_ub); };
  if (bound(mp.feat_used_ub + need) <= mp.feat_capacity) {
    mp.feat_used_ub = bound(mp.feat_used_ub + need);
    return NVBX_OK;
  }
  int rc = read_ctrl(mp, stream);
  if (rc) return rc;
  mp.feat_used_ub = (long long)mp.h_ctrl->feat_high - mp.h_ctrl->feat_free_top;
  mp.slot_used_ub = (long long)mp.h_ctrl->slot_high - mp.h_ctrl->slot_free_top;
  need = std::min(need, mp.slot_used_ub);
  long long want = bound(mp.feat_used_ub + need);
  if (want > mp.feat_capacity) {
    const long long block_bytes = (long long)kVoxelsPerBlock * mp.dev.row * (long long)sizeof(__half);
    long long ncap;
    if (count_new && want * block_bytes > kFeatPessimisticBytes) {
      CUDA_TRY(cudaMemsetAsync(&mp.d_ctrl->list_count, 0, sizeof(int), stream));
      if ((rc = count_new())) return rc;
      if ((rc = read_ctrl(mp, stream))) return rc;
      want = mp.feat_used_ub + (long long)mp.h_ctrl->list_count;
      ncap = std::max(want, std::min((long long)(mp.feat_capacity * 1.25f), want + kFeatExactSlackBytes / block_bytes));
      ncap = std::min(ncap, std::max(want, mp.slot_used_ub));
    } else {
      ncap = std::max(want, (long long)(mp.feat_capacity * 1.5f));
      ncap = std::min(ncap, std::max(want, mp.slot_used_ub));
    }
    if (ncap > mp.feat_capacity) {
      size_t free_b = 0, total_b = 0;
      CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
      const long long grow_bytes = (ncap - mp.feat_capacity) * block_bytes;
      if (grow_bytes + (1LL << 30) > (long long)free_b)
        return fail(NVBX_ERR_OUT_OF_MEMORY, "feature arena: %lld more blocks (%.1f GiB) do not fit the %.1f GiB free",
                    ncap - mp.feat_capacity, grow_bytes / 1073741824.0, free_b / 1073741824.0);
      if ((rc = grow_feats(m, mp, (int)ncap, stream))) return rc;
    }
  }
  mp.feat_used_ub = want;
  return NVBX_OK;
}

struct GridSpec {
  ViewGrid g;
  bool empty;
};
// AABB -> block grid (view_calculator.cu:283-289); workspace clipping (workspace_bounds.cpp:20-61)
int make_grid(const nvbx_mapper* m, const Map& mp, Aabb a, GridSpec* out) {
  const nvbx_params& p = m->params;
  if (p.workspace_bounds_type == NVBX_WORKSPACE_HEIGHT_BOUNDS) {
    a.mn[2] = fmaxf(a.mn[2], p.workspace_min[2]);
    a.mx[2] = fminf(a.mx[2], p.workspace_max[2]);
  } else if (p.workspace_bounds_type == NVBX_WORKSPACE_BOUNDING_BOX) {
    for (int i = 0; i < 3; ++i) {
      a.mn[i] = fmaxf(p.workspace_min[i], a.mn[i]);
      a.mx[i] = fminf(p.workspace_max[i], a.mx[i]);
    }
  }
  out->empty = a.empty();
  if (out->empty) return NVBX_OK;
  V3 lo, hi;
  lo.x = a.mn[0];
  lo.y = a.mn[1];
  lo.z = a.mn[2];
  hi.x = a.mx[0];
  hi.y = a.mx[1];
  hi.z = a.mx[2];
  const I3 mn = block_index_from_position(mp.block_size, lo);
  const I3 mx = block_index_from_position(mp.block_size, hi);
  const long long sx = (long long)mx.x - mn.x + 1, sy = (long long)mx.y - mn.y + 1, sz = (long long)mx.z - mn.z + 1;
  const long long n = sx * sy * sz;
  if (sx <= 0 || sy <= 0 || sz <= 0 || n > (1LL << 30))
    return fail(NVBX_ERR_UNSUPPORTED, "view AABB spans %lld x %lld x %lld blocks; set workspace bounds", sx, sy, sz);
  out->g.mn = mn;
  out->g.sx = (int)sx;
  out->g.sy = (int)sy;
  out->g.sz = (int)sz;
  out->g.n_cells = (int)n;
  return NVBX_OK;
}

thread_local bool t_multi_map_batch = false;  // this thread is issuing one map's share of a multi-map batch
thread_local bool t_async_worker = false;  // this thread is a mapper's enqueue worker: never wait for the FIFO
// Wait until every asynchronously queued frame has been issued; report the error one of them ran into.
int async_flush(nvbx_mapper* m) {
  if (!m || !m->async || t_async_worker) return NVBX_OK;
  std::string err;
  const int rc = m->async->flush(&err);
  if (rc) return fail(rc, "a frame queued by the asynchronous enqueue failed: %s", err.c_str());
  return NVBX_OK;
}
int check_map(nvbx_mapper* m, int map_id, bool flush = true) {
  if (!m) return fail(NVBX_ERR_INVALID_ARGUMENT, "null mapper handle");
  if (flush)
    if (int frc = async_flush(m)) return frc;
  if (map_id < 0 || map_id >= (int)m->maps.size())
    return fail(NVBX_ERR_INVALID_ARGUMENT, "mapper_id %d out of range [0, %d)", map_id, (int)m->maps.size());
  cudaError_t e = cudaSetDevice(m->device);
  if (e != cudaSuccess) return fail(NVBX_ERR_CUDA, "cudaSetDevice(%d): %s", m->device, cudaGetErrorString(e));
  return NVBX_OK;
}

int init_map(nvbx_mapper* m, Map& mp, float voxel_size, cudaStream_t stream) {
  mp.voxel_size = voxel_size;
  mp.block_size = voxel_size * 8;  // voxelSizeToBlockSize
  mp.dev = MapDev{};
  mp.dev.block_size = mp.block_size;
  mp.dev.voxel_size = voxel_size;
  mp.dev.voxel_size_inv = (float)(1.0 / (double)(mp.block_size * (1.0f / 8.0f)));
  mp.dev.C = m->C;
  mp.dev.row = m->C + 8;
  CUDA_TRY(cudaMalloc(&mp.d_ctrl, sizeof(Ctrl)));
  CUDA_TRY(cudaMemsetAsync(mp.d_ctrl, 0, sizeof(Ctrl), stream));
  CUDA_TRY(cudaMallocHost(&mp.h_ctrl, sizeof(Ctrl)));
  std::memset(mp.h_ctrl, 0, sizeof(Ctrl));
  mp.dev.ctrl = mp.d_ctrl;
  CUDA_TRY(cudaMalloc(&mp.d_tsdf_table, kMaxSlabs * sizeof(void*)));
  CUDA_TRY(cudaMalloc(&mp.d_feat_table, kMaxSlabs * sizeof(void*)));
  CUDA_TRY(cudaMalloc(&mp.d_color_table, kMaxSlabs * sizeof(void*)));
  mp.dev.tsdf_slabs = mp.d_tsdf_table;
  mp.dev.feat_slabs = mp.d_feat_table;
  mp.dev.color_slabs = mp.d_color_table;
  CUDA_TRY(cudaMalloc(&mp.d_tmp_int, sizeof(int)));
  CUDA_TRY(cudaMalloc(&mp.d_tmp_ptr, sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc(&mp.d_exp_total, sizeof(long long)));
  CUDA_TRY(cudaMalloc(&mp.scratch_ray.d_count, sizeof(int)));
  CUDA_TRY(cudaMalloc(&mp.small_grid[0], 3 * kFusedBitmapWords * sizeof(unsigned)));
  CUDA_TRY(cudaMemsetAsync(mp.small_grid[0], 0, 3 * kFusedBitmapWords * sizeof(unsigned), stream));
  mp.small_grid[1] = mp.small_grid[0] + kFusedBitmapWords;
  mp.small_grid[2] = mp.small_grid[1] + kFusedBitmapWords;
  // level-1 index: direct-mapped grid over the workspace box (same rounding as make_grid)
  if (m->params.workspace_bounds_type == NVBX_WORKSPACE_BOUNDING_BOX) {
    Aabb a;
    for (int i = 0; i < 3; ++i) {
      a.mn[i] = m->params.workspace_min[i];
      a.mx[i] = m->params.workspace_max[i];
    }
    if (!a.empty()) {
      V3 lo = {a.mn[0], a.mn[1], a.mn[2]}, hi = {a.mx[0], a.mx[1], a.mx[2]};
      const I3 mn = block_index_from_position(mp.block_size, lo);
      const I3 mx = block_index_from_position(mp.block_size, hi);
      const long long sx = (long long)mx.x - mn.x + 1, sy = (long long)mx.y - mn.y + 1, sz = (long long)mx.z - mn.z + 1;
      if (sx > 0 && sy > 0 && sz > 0 && sx * sy * sz <= kMaxWorkspaceGridCells && key_in_range(mn.x, mn.y, mn.z) &&
          key_in_range(mx.x, mx.y, mx.z)) {
        const int cells = (int)(sx * sy * sz);
        CUDA_TRY(cudaMalloc(&mp.dev.ws_slot, (size_t)cells * sizeof(int)));
        CUDA_TRY(cudaMemsetAsync(mp.dev.ws_slot, 0xff, (size_t)cells * sizeof(int), stream));  // -1
        mp.dev.ws_mn = mn;
        mp.dev.ws_sx = (int)sx;
        mp.dev.ws_sy = (int)sy;
        mp.dev.ws_sz = (int)sz;
        mp.dev.ws_cells = cells;
      }
    }
  }
  int rc = grow_slots(m, mp, std::max(1024, m->params.num_preallocated_blocks), stream);
  if (rc) return rc;
  return grow_feats(m, mp, 1 << kFeatSlabShift, stream);
}

void destroy_map(Map& mp) {
  cudaDeviceSynchronize();
  auto F = [](auto*& p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  F(mp.dev.ws_slot);
  F(mp.dev.keys);
  F(mp.dev.vals);
  F(mp.dev.blk_index);
  F(mp.dev.blk_layers);
  F(mp.dev.blk_feat);
  F(mp.dev.blk_dirty);
  F(mp.dev.blk_mesh);
  F(mp.dev.blk_cmesh);
  F(mp.dev.slot_free);
  F(mp.dev.feat_free);
  for (void* s : mp.tsdf_slabs) cudaFree(s);
  for (void* s : mp.feat_slabs) cudaFree(s);
  for (void* s : mp.color_slabs) cudaFree(s);
  mp.tsdf_slabs.clear();
  mp.feat_slabs.clear();
  mp.color_slabs.clear();
  F(mp.d_tsdf_table);
  F(mp.d_feat_table);
  F(mp.d_color_table);
  F(mp.d_ctrl);
  if (mp.h_ctrl) cudaFreeHost(mp.h_ctrl);
  mp.h_ctrl = nullptr;
  F(mp.d_tmp_int);
  F(mp.d_tmp_ptr);
  F(mp.d_exp_total);
  mp.exp_keep.release();
  mp.exp_tile_cnt.release();
  mp.exp_tile_off.release();
  mp.exp_idx.release();
  mp.exp_v.release();
  mp.exp_f.release();
  mp.raycast_cache.release();
  mp.planes_cache.release();
  mp.scratch_ray.idx.release();
  mp.scratch_planes.idx.release();
  F(mp.scratch_ray.d_count);
  F(mp.small_grid[0]);
  mp.small_grid[1] = nullptr;
  mp.grid.release();
  mp.view_slots.release();
  for (int b = 0; b < kFrameRing; ++b) {
    mp.band_slots2[b].release();
    mp.newfeat_slots2[b].release();
    mp.synth2[b].release();
  }
  mp.cband_slots.release();
  for (int r = 0; r < kFrameRing; ++r) mp.items2[r].release();
  if (mp.gstream) {
    cudaStreamSynchronize(mp.gstream);
    cudaStreamDestroy(mp.gstream);
    cudaStreamSynchronize(mp.cstream);
    cudaStreamDestroy(mp.cstream);
    cudaEventDestroy(mp.ev_geom);
    cudaEventDestroy(mp.ev_trace);
    for (int r = 0; r < kFrameRing; ++r) cudaEventDestroy(mp.ev_gather[r]);
    mp.gstream = nullptr;
  }
  mp.cnt_v.release();
  mp.cnt_t.release();
  mp.off_v.release();
  mp.off_t.release();
  for (int i = 0; i < 2; ++i) {
    mp.arena_v[i].release();
    mp.arena_f[i].release();
    mp.arena_t[i].release();
    mp.carena_v[i].release();
    mp.carena_c[i].release();
    mp.carena_t[i].release();
  }
  mp.st_depth.release();
  mp.st_feat.release();
  mp.st_pixmap.release();
  mp.st_mask_d.release();
  mp.st_mask_f.release();
  mp.st_low.release();
  mp.st_low_in.release();
  mp.idx_out.release();
  mp.ptr_out.release();
  mp.slot_out.release();
}

Cam make_cam(float fx, float fy, float cx, float cy, int H, int W) {
  Cam c;
  c.fu = fx;
  c.fv = fy;
  c.cu = cx;
  c.cv = cy;
  c.width = W;
  c.height = H;
  return c;
}

int timing_begin(nvbx_mapper* m, int which, cudaStream_t stream) {
  if (!m->timing || g_timing_mode >= 2) return NVBX_OK;
  cudaEvent_t a, b;
  CUDA_TRY(cudaEventCreate(&a));
  CUDA_TRY(cudaEventCreate(&b));
  m->timing_events[which].emplace_back(a, b);
  CUDA_TRY(cudaEventRecord(a, stream));
  return NVBX_OK;
}
int timing_end(nvbx_mapper* m, int which, cudaStream_t stream) {
  if (!m->timing || g_timing_mode >= 2) return NVBX_OK;
  CUDA_TRY(cudaEventRecord(m->timing_events[which].back().second, stream));
  return NVBX_OK;
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
// Schedule of the feature gather (tuning hooks; profiles/r01b_gather_variants.md, r01i_gather_schedule.md):
// NVBX_GATHER_VARIANT / NVBX_GATHER_DYN / NVBX_GATHER_TICKET at load, nvbx_set_gather_tuning at run time.
constexpr int kDefaultGatherVariant = 7;  // 48-register static deal + item prefetch on 4 CTAs / SM (profiles/r01k_gather_schedule.md)
int g_gather_tuning[3] = {-1, -2, -1};
int gather_variant() {
  if (g_gather_tuning[0] < 0) g_gather_tuning[0] = std::max(0, env_int("NVBX_GATHER_VARIANT", kDefaultGatherVariant));
  return g_gather_tuning[0];
}
int gather_dyn_permille() {  // share of the units handed out by ticket (k_feature_gather_dyn)
  if (g_gather_tuning[1] < -1) g_gather_tuning[1] = std::min(1000, std::max(-1, env_int("NVBX_GATHER_DYN", 0)));
  return g_gather_tuning[1];
}
int gather_ticket_units() {
  if (g_gather_tuning[2] < 0) g_gather_tuning[2] = std::min(64, std::max(1, env_int("NVBX_GATHER_TICKET", 4)));
  return g_gather_tuning[2];
}

template <int CH>
int launch_gather(nvbx_mapper* m, Map& mp, const FeatFrame& ff, int last_chunk, cudaStream_t stream, bool pipelined) {
  int rc;
  if ((rc = timing_begin(m, 0, stream))) return rc;
  switch (gather_variant()) {  // <CH, units in flight per warp, resident CTAs per SM>
    case 1:
      LAUNCH((k_feature_gather<CH, 2, 3>), persistent_grid(m, 3), 256, 0, stream, mp.dev, mp.items2[mp.dev.fp].p, ff, last_chunk);
      break;
    case 2:
      LAUNCH((k_feature_gather<CH, 1, 6>), persistent_grid(m, 6), 256, 0, stream, mp.dev, mp.items2[mp.dev.fp].p, ff, last_chunk);
      break;
    case 3:
      LAUNCH((k_feature_gather<CH, 1, 8>), persistent_grid(m, 8), 256, 0, stream, mp.dev, mp.items2[mp.dev.fp].p, ff, last_chunk);
      break;
    case 5:
      LAUNCH((k_feature_gather_dyn<CH, 512, 2>), persistent_grid(m, 2), 512, 0, stream, mp.dev, mp.items2[mp.dev.fp].p,
             (int)mp.items2[mp.dev.fp].cap, ff, last_chunk, gather_dyn_permille(), gather_ticket_units());
      break;
    case 6:
      LAUNCH((k_feature_gather_dyn<CH, 256, 5>), persistent_grid(m, 5), 256, 0, stream, mp.dev, mp.items2[mp.dev.fp].p,
             (int)mp.items2[mp.dev.fp].cap, ff, last_chunk, gather_dyn_permille(), gather_ticket_units());
      break;
    case 0:
      LAUNCH((k_feature_gather<CH, 1, 4>), persistent_grid(m, 4), 256, 0, stream, mp.dev, mp.items2[mp.dev.fp].p, ff, last_chunk);
      break;
    case 10:  // cp.async.bulk + mbarrier staging, one 8-warp CTA per SM (C = 256 * CH only)
      if constexpr (CH > 0) {
        static const bool once = [] {
          cudaFuncSetAttribute(k_feature_gather_tma<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               TmaGather<CH>::kSmemBytes);
          return true;
        }();
        (void)once;
        LAUNCH((k_feature_gather_tma<CH>), persistent_grid(m, 1), kTmaWarps * 32, TmaGather<CH>::kSmemBytes, stream,
               mp.dev, mp.items2[mp.dev.fp].p, ff, last_chunk);
      } else {
        LAUNCH((k_feature_gather_dyn<CH, 256, 5>), persistent_grid(m, 4), 256, 0, stream, mp.dev, mp.items2[mp.dev.fp].p,
               (int)mp.items2[mp.dev.fp].cap, ff, last_chunk, gather_dyn_permille(), gather_ticket_units());
      }
      break;
    case 7: {  // the 48-register build on 4 CTAs / SM: leaves registers for two CTAs of the next frame's raycast
      static const int waves = std::max(1, env_int("NVBX_GATHER_WAVES", 1));  // tuning knob: grid = waves x resident CTAs
      static const int pipe_ctas = std::min(4, std::max(1, env_int("NVBX_PIPE_GATHER_CTAS", 3)));  // tuning knob
      static const int pad = [] {  // tuning knob: dynamic shared memory per CTA (caps the resident CTAs per SM)
        const int v = std::max(0, env_int("NVBX_GATHER_SMEM_PAD", 0));
        if (v > 48 * 1024)
          cudaFuncSetAttribute(k_feature_gather_dyn<CH, 256, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
        return v;
      }();
      // On its own stream (frame pipelining) the gather runs on THREE CTAs per SM: the 28 k registers it leaves free
      // hold one CTA of any kernel of the next frame's depth path (k_tsdf_update and k_trace_and_band need 20 k),
      // which is what lets those kernels run underneath it (25.7 k vs 24.2 k frames/s with four; r02 sweep).
      LAUNCH((k_feature_gather_dyn<CH, 256, 5>), persistent_grid(m, pipelined ? pipe_ctas : 4) * waves, 256, pad, stream, mp.dev, mp.items2[mp.dev.fp].p,
             (int)mp.items2[mp.dev.fp].cap, ff, last_chunk, gather_dyn_permille(), gather_ticket_units());
      break;
    }
    case 8:
      LAUNCH((k_feature_gather_dyn<CH, 256, 5>), persistent_grid(m, 3), 256, 0, stream, mp.dev, mp.items2[mp.dev.fp].p,
             (int)mp.items2[mp.dev.fp].cap, ff, last_chunk, gather_dyn_permille(), gather_ticket_units());
      break;
    case 9:
      LAUNCH((k_feature_gather_dyn<CH, 256, 6>), persistent_grid(m, 4), 256, 0, stream, mp.dev, mp.items2[mp.dev.fp].p,
             (int)mp.items2[mp.dev.fp].cap, ff, last_chunk, gather_dyn_permille(), gather_ticket_units());
      break;
    default:
      LAUNCH((k_feature_gather_dyn<CH, 256, 4>), persistent_grid(m, 4), 256, 0, stream, mp.dev, mp.items2[mp.dev.fp].p,
             (int)mp.items2[mp.dev.fp].cap, ff, last_chunk, gather_dyn_permille(), gather_ticket_units());
  }
  return timing_end(m, 0, stream);
}

template <int CH>
int launch_gather_up(nvbx_mapper* m, Map& mp, const FeatFrame& ff, const UpFrame& uf, int mode, int last_chunk,
                     cudaStream_t stream) {
  int rc;
  if ((rc = timing_begin(m, 0, stream))) return rc;
  switch (mode) {  // <CH, torch kernel flavour, resident CTAs per SM>
    case 1:
      LAUNCH((k_feature_gather_up<CH, 1, 3>), persistent_grid(m, 3), 256, 0, stream, mp.dev, mp.items2[mp.dev.fp].p, ff, uf,
             last_chunk);
      break;
    case 2:
      LAUNCH((k_feature_gather_up<CH, 2, 3>), persistent_grid(m, 3), 256, 0, stream, mp.dev, mp.items2[mp.dev.fp].p, ff, uf,
             last_chunk);
      break;
    default:
      LAUNCH((k_feature_gather_up<CH, 0, 3>), persistent_grid(m, 3), 256, 0, stream, mp.dev, mp.items2[mp.dev.fp].p, ff, uf,
             last_chunk);
  }
  return timing_end(m, 0, stream);
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char* nvbx_version(void) { return "nvbx 0.1.0 (sm_100a)"; }
const char* nvbx_last_error(void) { return g_last_error.c_str(); }
int64_t nvbx_kernel_launch_count(void) { return g_launches.load(); }
void nvbx_pipeline_wait_stats(int64_t* waits, int64_t* wait_ns) {
  if (waits) *waits = g_pipe_waits.load();
  if (wait_ns) *wait_ns = g_pipe_wait_ns.load();
}

void nvbx_default_params(nvbx_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->max_integration_distance_m = 7.0f;
  p->truncation_distance_vox = 4.0f;
  p->weighting_mode = NVBX_WEIGHT_INVERSE_SQUARE;
  p->max_weight = 5.0f;
  p->invalid_depth_decay_factor = -1.0f;
  p->appearance_measurement_weight = 0.8f;
  p->appearance_truncation_distance_vox = 4.0f;
  p->sphere_tracing_subsampling = 4;
  p->sphere_tracing_max_ray_length_m = 7.0f;
  p->sphere_tracing_max_steps = 100;
  p->sphere_tracing_surface_epsilon_vox = 0.1f;
  p->tsdf_decay_factor = 0.95f;
  p->tsdf_decayed_weight_threshold = 1e-3f;
  p->tsdf_set_free_distance_on_decayed = 0;
  p->tsdf_decayed_free_distance_vox = 4.0f;
  p->deallocate_decayed_blocks = 1;
  p->raycast_subsampling_factor = 4;
  p->workspace_bounds_type = NVBX_WORKSPACE_UNBOUNDED;
  p->workspace_min[0] = 0.0f;
  p->workspace_min[1] = 2.0f;
  p->workspace_min[2] = 0.0f;
  p->workspace_max[0] = 0.0f;
  p->workspace_max[1] = 2.0f;
  p->workspace_max[2] = 1.0f;
  p->cache_last_viewpoint = 1;
  p->mesh_min_weight = 1e-4f;
  p->mesh_weld_vertices = 1;
  p->mesh_cutoff_distance_vox = 5.0f;
  p->num_preallocated_blocks = 2048;
  p->expansion_factor = 2.0f;
  p->strict_blend = 0;
}

int nvbx_create(int n_maps, const float* voxel_sizes_m, const nvbx_params* params, int feature_channels, int device,
                nvbx_mapper** out) {
  if (!out) return fail(NVBX_ERR_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  if (n_maps <= 0 || !voxel_sizes_m) return fail(NVBX_ERR_INVALID_ARGUMENT, "need at least one map");
  if (feature_channels <= 0 || feature_channels % 8)
    return fail(NVBX_ERR_INVALID_ARGUMENT, "feature_channels must be a positive multiple of 8 (got %d)",
                feature_channels);
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(NVBX_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
  if (device < 0 || device >= n_dev) return fail(NVBX_ERR_INVALID_ARGUMENT, "device %d out of range", device);
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(NVBX_ERR_UNSUPPORTED, "device %s is sm_%d%d; libnvbx is built for sm_100a only", prop.name, prop.major,
                prop.minor);
  std::unique_ptr<nvbx_mapper> m(new nvbx_mapper());
  m->device = device;
  m->C = feature_channels;
  m->sm_count = prop.multiProcessorCount;
  if (params)
    m->params = *params;
  else
    nvbx_default_params(&m->params);
  const nvbx_params& p = m->params;
  if (!(p.appearance_measurement_weight > 0.0f && p.appearance_measurement_weight <= 1.0f))
    return fail(NVBX_ERR_INVALID_ARGUMENT, "appearance_measurement_weight must be in (0, 1]");
  if (p.raycast_subsampling_factor <= 0 || p.sphere_tracing_subsampling <= 0)
    return fail(NVBX_ERR_INVALID_ARGUMENT, "subsampling factors must be positive");
  // the mesh kernels need > 48 KiB of dynamic shared memory
  CUDA_TRY(cudaFuncSetAttribute(k_mesh_count<kMeshFeature>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(MeshSmem)));
  CUDA_TRY(cudaFuncSetAttribute(k_mesh_emit<kMeshFeature>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(MeshSmem)));
  CUDA_TRY(cudaFuncSetAttribute(k_mesh_count<kMeshColor>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(MeshSmem)));
  CUDA_TRY(cudaFuncSetAttribute(k_mesh_emit<kMeshColor>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(MeshSmem)));
  for (int i = 0; i < n_maps; ++i) {
    if (!(voxel_sizes_m[i] > 0.0f)) return fail(NVBX_ERR_INVALID_ARGUMENT, "voxel size must be positive");
    m->maps.emplace_back(new Map());
    int rc = init_map(m.get(), *m->maps[i], voxel_sizes_m[i], 0);
    if (rc) {
      for (auto& mp : m->maps) destroy_map(*mp);
      return rc;
    }
  }
  CUDA_TRY(cudaDeviceSynchronize());
  *out = m.release();
  return NVBX_OK;
}

void nvbx_destroy(nvbx_mapper* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  m->async.reset();  // issues what is still queued, then stops the worker
  for (auto& mp : m->maps) pipeline_drain(*mp);
  for (auto& mp : m->maps) destroy_map(*mp);
  delete m;
}

int nvbx_num_maps(const nvbx_mapper* m) { return m ? (int)m->maps.size() : 0; }
int nvbx_feature_channels(const nvbx_mapper* m) { return m ? m->C : 0; }
int nvbx_get_params(const nvbx_mapper* m, nvbx_params* out) {
  if (!m || !out) return fail(NVBX_ERR_INVALID_ARGUMENT, "null argument");
  *out = m->params;
  return NVBX_OK;
}
float nvbx_voxel_size(const nvbx_mapper* m, int map_id) {
  if (!m || map_id < 0 || map_id >= (int)m->maps.size()) return 0.0f;
  return m->maps[map_id]->voxel_size;
}

// ---- depth ---------------------------------------------------------------------------------------
int nvbx_integrate_depth(nvbx_mapper* m, int map_id, const void* depth, int height, int width, const void* mask,
                         const float* T_L_C_rm, float fx, float fy, float cx, float cy, void* stream_v) {
  int rc = check_map(m, map_id, /*flush=*/false);
  if (rc) return rc;
  if (!depth || !T_L_C_rm || height <= 0 || width <= 0) return fail(NVBX_ERR_INVALID_ARGUMENT, "bad depth frame");
  if (m->async && !t_async_worker) {  // asynchronous enqueue: the mapper's worker thread issues the launches
    AsyncJob j{0, map_id, depth, height, width, 0, mask, {}, fx, fy, cx, cy, stream_v};
    std::memcpy(j.T, T_L_C_rm, sizeof(j.T));
    m->async->push(j);
    return NVBX_OK;
  }
  cudaStream_t stream = (cudaStream_t)stream_v;
  Map& mp = *m->maps[map_id];
  const nvbx_params& p = m->params;
  const Pose T_L_C = pose_from_row_major(T_L_C_rm);
  const Cam cam = make_cam(fx, fy, cx, cy, height, width);
  const float trunc = p.truncation_distance_vox * mp.voxel_size;

  ViewSource vs{};
  int view_mode;
  CacheEntry* entry = p.cache_last_viewpoint ? mp.raycast_cache.lookup(T_L_C, cam) : nullptr;
  if (entry) {
    // cache hit: previous block list, blocks (re-)allocated where required (inside the TSDF kernel)
    if ((rc = ensure_slots(m, mp, entry->bound, stream))) return rc;
    view_mode = kViewFromEntry;
  } else {
    GridSpec gs;
    if ((rc = make_grid(m, mp, view_aabb(cam, T_L_C, 0.0f, p.max_integration_distance_m), &gs))) return rc;
    if (gs.empty) {
      mp.last_depth_entry = nullptr;
      return NVBX_OK;  // nothing in view, nothing cached (view_calculator.cu:275-279)
    }
    if (p.cache_last_viewpoint) {
      if ((rc = mp.raycast_cache.acquire(&entry))) return rc;
    } else {
      entry = &mp.scratch_ray;
    }
    const size_t n = (size_t)gs.g.n_cells;
    const size_t n_words = (n + 31) / 32;
    if ((rc = entry->idx.ensure(n, stream))) return rc;
    // Small views (<= 32 768 cells: every mindmap workspace) mark one half of a double-buffered bitmap and the
    // TSDF kernel compacts it itself; larger views use the growable bitmap + k_view_compact_alloc.
    const bool fused = n_words <= (size_t)kFusedBitmapWords;
    unsigned* bits;
    if (fused) {
      if (mp.small_dirty) {
        CUDA_TRY(cudaMemsetAsync(mp.small_grid[0], 0, 3 * kFusedBitmapWords * sizeof(unsigned), stream));
        mp.small_dirty = false;
      }
      bits = mp.small_grid[mp.small_idx];
    } else {
      const unsigned* old_grid = mp.grid.p;
      if ((rc = mp.grid.ensure(n_words, stream))) return rc;
      if (mp.grid.p != old_grid || mp.grid_dirty)  // a fresh (or abandoned) bitmap starts all-zero
        CUDA_TRY(cudaMemsetAsync(mp.grid.p, 0, mp.grid.cap * sizeof(unsigned), stream));
      if ((rc = mp.view_slots.ensure(n, stream))) return rc;
      bits = mp.grid.p;
    }
    entry->T = T_L_C;
    entry->cam = cam;
    entry->bound = gs.g.n_cells;
    const int s = p.raycast_subsampling_factor;
    const int n_rows = (int)std::ceil((float)(height + 1) / (float)s);
    const int n_cols = (int)std::ceil((float)(width + 1) / (float)s);
    const int tiles_x = (n_cols + 15) / 16, n_tiles = tiles_x * ((n_rows + 15) / 16);
    // one tile per CTA while they all fit on the machine at once (8 CTAs of 256 threads per SM): the hardware
    // scheduler then balances tiles of unequal cost; larger images loop
    const int rgrid = std::max(1, std::min(n_tiles, persistent_grid(m, 8)));
    (fused ? mp.small_dirty : mp.grid_dirty) = true;
    RaycastFrame rf;
    rf.T_L_C = T_L_C;
    rf.cam = cam;
    rf.depth = (const float*)depth;
    rf.rows = height;
    rf.cols = width;
    rf.block_size = mp.block_size;
    rf.max_dist = p.max_integration_distance_m;
    rf.behind = trunc;
    rf.sub = s;
    rf.g = gs.g;
    for (int i = 0; i < 3; ++i) {  // RayCaster constructor, pixel-independent half (ray_caster_impl.h:26-53)
      rf.s[i] = (T_L_C.t[i] / mp.block_size) / 1.0f;
      rf.start[i] = (int)floorf(rf.s[i]);
      rf.shifted[i] = rf.s[i] - (float)rf.start[i];
    }
    rf.lin0 = (int)((unsigned)(rf.start[0] - gs.g.mn.x) + (unsigned)(rf.start[1] - gs.g.mn.y) * (unsigned)gs.g.sx +
                    (unsigned)(rf.start[2] - gs.g.mn.z) * (unsigned)gs.g.sx * (unsigned)gs.g.sy);
    rf.tiles_x = tiles_x;
    rf.n_tiles = n_tiles;
    rf.seq = mp.dev.seq;
    if (n_words <= (size_t)kRayBitmapWords) {
      static const int early_flush = env_int("NVBX_RAYCAST_EARLY_FLUSH", 1);  // tuning knob
      LAUNCH(k_raycast_mark<true>, rgrid, 256, n_words * sizeof(unsigned), stream, rf, bits, entry->d_count,
             (fused && early_flush) ? 1 : 0, mp.small_seq, &mp.d_ctrl->bitmap_clean_seq);
    } else {
      LAUNCH(k_raycast_mark<false>, rgrid, 256, 0, stream, rf, bits, entry->d_count, 0, 0u,
             &mp.d_ctrl->bitmap_clean_seq);
    }
    const int cgrid = std::max(1, std::min(persistent_grid(m, 4), (int)((n_words + 7) / 8)));  // warp per word
    if (!slots_fit(m, mp, (long long)n)) {
      // The pessimistic bound (every cell of the AABB is new) does not fit: count the marked cells and
      // read that one int back.  Only happens while the arena is still growing (or with an unbounded
      // workspace); a bounded workspace reaches its cell count and never synchronises again.
      CUDA_TRY(cudaMemsetAsync(mp.d_tmp_int, 0, sizeof(int), stream));
      LAUNCH(k_count_marked, cgrid, 256, 0, stream, bits, (int)n_words, mp.d_tmp_int);
      int marked = 0;
      CUDA_TRY(cudaMemcpyAsync(&marked, mp.d_tmp_int, sizeof(int), cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      entry->bound = std::max(1, marked);
    }
    if ((rc = ensure_slots(m, mp, (long long)entry->bound, stream))) return rc;
    if (fused) {
      view_mode = kViewFromBitmap;
      vs.bits = bits;
      vs.clean_bits = mp.small_grid[(mp.small_idx + 2) % 3];
      vs.clean_seq = mp.small_seq + 2u;
      vs.g = gs.g;
    } else {
      LAUNCH(k_view_compact_alloc, cgrid, 256, 0, stream, mp.dev, bits, gs.g, entry->idx.p, entry->d_count,
             mp.view_slots.p);
      mp.grid_dirty = false;
      view_mode = kViewFromSlots;
      vs.view_slots = mp.view_slots.p;
    }
    if (p.cache_last_viewpoint) mp.raycast_cache.store(entry);
  }
  vs.entry_idx = entry->idx.p;
  vs.entry_count = entry->d_count;
  mp.last_depth_entry = entry;
  DepthFrame f;
  f.depth = (const float*)depth;
  f.mask = (const uint8_t*)mask;
  f.rows = height;
  f.cols = width;
  f.cam = cam;
  f.T_C_L = inverse(T_L_C);
  f.max_depth = p.max_integration_distance_m;
  f.trunc = trunc;
  f.max_weight = p.max_weight;
  f.invalid_decay = p.invalid_depth_decay_factor;
  f.weighting_mode = p.weighting_mode;
  const int tgrid = std::max(1, std::min(persistent_grid(m, 3), entry->bound));  // one block per CTA up to 3 CTAs/SM
  ++mp.tsdf_version;
  if ((rc = timing_begin(m, 1, stream))) return rc;
  if (view_mode == kViewFromBitmap) {
    LAUNCH(k_tsdf_update<kViewFromBitmap>, tgrid, 512, 0, stream, mp.dev, vs, f);
    ++mp.small_seq;  // the third this frame marked is wiped by the next frame's kernel
    mp.small_idx = (mp.small_idx + 1) % 3;
    mp.small_dirty = false;
  } else if (view_mode == kViewFromEntry) {
    LAUNCH(k_tsdf_update<kViewFromEntry>, tgrid, 512, 0, stream, mp.dev, vs, f);
  } else {
    LAUNCH(k_tsdf_update<kViewFromSlots>, tgrid, 512, 0, stream, mp.dev, vs, f);
  }
  return timing_end(m, 1, stream);
}

}  // extern "C"

// ---- appearance frames (features, colour) ----------------------------------------------------------
namespace {

// What the feature and the colour integrator share (ProjectiveAppearanceIntegrator<LayerType>::integrateFrame,
// projective_appearance_integrator.cu:72-169): the planes view of the frame (through the integrator's own
// viewpoint cache), the band-select / sphere-trace launch and the synthetic depth image.
struct AppearancePrep {
  bool empty = false;
  float trunc = 0;
  Pose T_L_C, T_C_L;
  Cam cam;
  int srows = 0, scols = 0, sub = 0;
  long long cand_bound = 0;
  const float* synth = nullptr;  // the frame's synthetic depth image (freshly traced or re-used)
};

// color_parity < 0: feature frame (band list -> mp.band_slots / ctrl->band_count, feature slots allocated);
// otherwise colour frame (band list -> mp.cband_slots / ctrl->cband_count[parity]).
int appearance_prepare(nvbx_mapper* m, Map& mp, ViewCache& cache, int height, int width, const float* T_L_C_rm,
                       float fx, float fy, float cx, float cy, int color_parity, cudaStream_t stream,
                       AppearancePrep* out) {
  const nvbx_params& p = m->params;
  int rc;
  const int sub = p.sphere_tracing_subsampling;
  if (width % sub || height % sub)
    return fail(NVBX_ERR_INVALID_ARGUMENT, "frame %dx%d is not divisible by the sphere-tracing subsampling %d", height,
                width, sub);
  const Pose T_L_C = pose_from_row_major(T_L_C_rm);
  const Cam cam = make_cam(fx, fy, cx, cy, height, width);
  const float trunc = p.appearance_truncation_distance_vox * mp.voxel_size;
  out->T_L_C = T_L_C;
  out->T_C_L = inverse(T_L_C);
  out->cam = cam;
  out->trunc = trunc;

  CacheEntry* entry = p.cache_last_viewpoint ? cache.lookup(T_L_C, cam) : nullptr;
  if (!entry) {
    GridSpec gs;
    if ((rc = make_grid(m, mp, view_aabb(cam, T_L_C, 1e-6f, p.max_integration_distance_m + trunc), &gs))) return rc;
    if (gs.empty) {
      out->empty = true;
      return NVBX_OK;
    }
    if (p.cache_last_viewpoint) {
      if ((rc = cache.acquire(&entry, false))) return rc;
    } else {
      entry = &mp.scratch_planes;
    }
    entry->T = T_L_C;
    entry->cam = cam;
    entry->grid = gs.g;
    entry->bound = gs.g.n_cells;
    if (p.cache_last_viewpoint) cache.store(entry);
  }
  // The in-view test runs with the pose / camera / AABB of the cache entry: on a hit these are the
  // cached ones, i.e. the block list the reference would have re-used (view_calculator.cu:401-405).
  PlanesView pv;
  pv.g = entry->grid;
  pv.T_C_L = inverse(entry->T);
  {
    const V3 vmin = ray_from_image_plane(entry->cam, -10.0f, -10.0f);
    const V3 vmax = ray_from_image_plane(entry->cam, (float)entry->cam.width + 10.0f, (float)entry->cam.height + 10.0f);
    pv.vmin_x = vmin.x;
    pv.vmin_y = vmin.y;
    pv.vmax_x = vmax.x;
    pv.vmax_y = vmax.y;
  }
  const long long cand_bound = std::min((long long)entry->bound, std::max(1LL, mp.slot_used_ub));
  out->cand_bound = cand_bound;
  int* band_list;
  if (color_parity < 0) {
    int tile_cells = 32;  // the tiling of the band-select launch below
    while (tile_cells < 256 && (entry->bound + tile_cells - 1) / tile_cells > 2 * m->sm_count) tile_cells <<= 1;
    const int n_tiles = (entry->bound + tile_cells - 1) / tile_cells;
    auto count_new = [&]() -> int {
      LAUNCH(k_band_count, std::max(1, std::min(n_tiles, persistent_grid(m, 4))), 256, 0, stream, mp.dev, pv, trunc,
             tile_cells, n_tiles);
      return NVBX_OK;
    };
    if ((rc = ensure_feats(m, mp, cand_bound, stream, count_new))) return rc;
    if ((rc = mp.band_slots2[mp.dev.fp].ensure((size_t)cand_bound, stream))) return rc;
    if ((rc = mp.newfeat_slots2[mp.dev.fp].ensure((size_t)cand_bound, stream))) return rc;
    const long long chunk_blocks = std::min<long long>(cand_bound, kFeatureChunkBlocks);
    if ((rc = mp.items2[mp.dev.fp].ensure((size_t)chunk_blocks * kVoxelsPerBlock, stream))) return rc;
    band_list = mp.band_slots2[mp.dev.fp].p;
  } else {
    if ((rc = enable_color(mp, stream))) return rc;
    if ((rc = mp.cband_slots.ensure((size_t)cand_bound, stream))) return rc;
    band_list = mp.cband_slots.p;
  }
  const int srows = height / sub, scols = width / sub;
  if ((rc = mp.synth2[mp.dev.fp].ensure((size_t)srows * scols, stream))) return rc;
  out->srows = srows;
  out->scols = scols;
  out->sub = height / srows;  // projective_integrator_impl.cuh:424-425

  TraceParams tp;
  tp.cam = cam;
  tp.T_L_C = T_L_C;
  tp.trunc = trunc;
  tp.max_steps = p.sphere_tracing_max_steps;
  tp.max_ray_length = p.sphere_tracing_max_ray_length_m;
  tp.eps = p.sphere_tracing_surface_epsilon_vox * mp.voxel_size;
  tp.sub = sub;
  tp.rows = srows;
  tp.cols = scols;
  tp.free_dist = p.truncation_distance_vox * mp.voxel_size;  // == DepthFrame::trunc of nvbx_integrate_depth
  static const int use_free = env_int("NVBX_TRACE_FREE", 1);
  tp.use_free = use_free;
  // Team march (eight lanes per ray, sphere_trace_team): bit-identical, measured 14.9 vs 16.1 us alone but SLOWER under
  // the gather of the previous frame (42.1 vs 39.3 us per frame): the long rays spend their steps next to the surface,
  // where every step depends on the value just read, so the extra lanes only add work.  Kept selectable.
  static const int team = env_int("NVBX_TRACE_TEAM", 0);
  tp.team = team;
  static const int cache_blocks = env_int("NVBX_TRACE_CACHE", 0);  // tuning knob: 0 scalar march, 2 / 4 blocks per warp in smem
  tp.cache_blocks = (!tp.team && (cache_blocks == 2 || cache_blocks == 4)) ? cache_blocks : 0;
  static const int march = env_int("NVBX_TRACE_MARCH", 256);
  tp.march = (!tp.cache_blocks && (march == 128 || march == 64 || march == 32)) ? march : 256;
  // synthetic depth already rendered for exactly this pose / camera / TSDF state (the other appearance frame of
  // the same step): skip the sphere tracing, keep the band selection
  Map::SynthKey& key = mp.synth_key;
  // (a traced image goes to synth2[fp]; the callers made sure that ring slot is free -- pipeline_slot_ready)
  const bool reuse = key.valid && key.tsdf_version == mp.tsdf_version && key.trunc == trunc && key.sub == sub &&
                     mp.synth_rows == srows && mp.synth_cols == scols && mp.synth2[key.buf].p != nullptr &&
                     std::memcmp(&key.T, &T_L_C, sizeof(Pose)) == 0 && std::memcmp(&key.cam, &cam, sizeof(Cam)) == 0;
  mp.synth_rows = srows;
  mp.synth_cols = scols;
  {
    // band-select tiles: small enough that a mindmap-sized AABB (a few hundred cells) spreads over many SMs
    int tile_cells = 32;
    while (tile_cells < 256 && (entry->bound + tile_cells - 1) / tile_cells > 2 * m->sm_count) tile_cells <<= 1;
    const int n_tiles = (entry->bound + tile_cells - 1) / tile_cells;
    // trace CTAs: 8 x 4 rays with eight lanes per ray (team march), or 16 x 16 rays with one thread per ray
    const int tile_h = tp.team ? 4 : (tp.march < 256 ? tp.march / 8 : 16);
    const int trace_tiles_x = (tp.team || tp.march < 256) ? (scols + 7) / 8 : (scols + 15) / 16;
    const int n_trace = reuse ? 0 : trace_tiles_x * ((srows + tile_h - 1) / tile_h);
    const int n_band = std::max(1, std::min(n_tiles, persistent_grid(m, 4)));
    key.valid = false;  // a failed launch leaves no valid image behind
    static const int spec = env_int("NVBX_TRACE_SPEC", 1), ilp = env_int("NVBX_BAND_ILP", 2);
    const size_t tb_smem = tp.cache_blocks ? (size_t)8 * tp.cache_blocks * (kVoxelsPerBlock * sizeof(float2) + sizeof(int)) : 0;
#define NVBX_TB(S, I)                                                                                                \
  do {                                                                                                               \
    static const bool big_smem_ok = [] {                                                                             \
      return cudaFuncSetAttribute(k_trace_and_band<S, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) == \
             cudaSuccess;                                                                                            \
    }();                                                                                                             \
    if (tb_smem > 48 * 1024 && !big_smem_ok) return fail(NVBX_ERR_CUDA, "cannot raise the tracer's shared memory");  \
    LAUNCH((k_trace_and_band<S, I>), n_trace + n_band, 256, tb_smem, stream, mp.dev, tp, mp.synth2[mp.dev.fp].p,     \
           trace_tiles_x, n_trace, pv, trunc, band_list, mp.newfeat_slots2[mp.dev.fp].p, tile_cells, n_tiles,        \
           color_parity);                                                                                            \
  } while (0)
    if (spec == 1 && ilp == 1) {
      NVBX_TB(1, 1);
    } else if (spec == 1) {
      NVBX_TB(1, 2);
    } else if (spec == 2 && ilp == 1) {
      NVBX_TB(2, 1);
    } else if (spec == 2) {
      NVBX_TB(2, 2);
    } else if (ilp == 1) {
      NVBX_TB(4, 1);
    } else {
      NVBX_TB(4, 2);
    }
#undef NVBX_TB
    key.T = T_L_C;
    key.cam = cam;
    key.trunc = trunc;
    key.sub = sub;
    key.tsdf_version = mp.tsdf_version;
    if (!reuse) key.buf = mp.dev.fp;
    key.valid = true;
  }
  mp.synth_last = key.buf;
  out->synth = mp.synth2[key.buf].p;
  return NVBX_OK;
}

}  // namespace

namespace {

// Stage a low-res backbone feature map (device pointer) as [lh][lw][C] fp32 and describe it (N4).  A dense HWC
// fp32 map that already has C channels is used in place.
int prepare_lowres(nvbx_mapper* m, Map& mp, const void* lowres, int low_h, int low_w, int low_c, int dtype, int layout,
                   int kernel, int height, int width, cudaStream_t stream, UpFrame* uf, int* mode) {
  if (!lowres || low_h <= 0 || low_w <= 0 || low_c <= 0 || height <= 0 || width <= 0)
    return fail(NVBX_ERR_INVALID_ARGUMENT, "bad low-res feature map");
  if (low_c > m->C)
    return fail(NVBX_ERR_INVALID_ARGUMENT, "low-res feature map has %d channels, the map stores %d", low_c, m->C);
  if (dtype < NVBX_LOWRES_F32 || dtype > NVBX_LOWRES_BF16 || (layout != NVBX_LOWRES_CHW && layout != NVBX_LOWRES_HWC) ||
      (kernel != NVBX_UPSAMPLE_TORCH_NCHW && kernel != NVBX_UPSAMPLE_TORCH_NHWC))
    return fail(NVBX_ERR_INVALID_ARGUMENT, "bad low-res dtype / layout / kernel selector");
  if (low_h == height && low_w == width)  // torch copies instead of interpolating (UpSampleBilinear2d.cu)
    return fail(NVBX_ERR_UNSUPPORTED, "low-res map already has the frame size: use nvbx_integrate_features");
  int rc;
  if (dtype == NVBX_LOWRES_F32 && layout == NVBX_LOWRES_HWC && low_c == m->C && !(((uintptr_t)lowres) & 15)) {
    uf->low = (const float*)lowres;
  } else {
    const size_t n = (size_t)low_h * low_w * m->C;
    if ((rc = mp.st_low.ensure(n, stream))) return rc;
    const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)persistent_grid(m, 8));
    LAUNCH(k_lowres_stage, grid, 256, 0, stream, lowres, dtype, layout == NVBX_LOWRES_CHW ? 1 : 0, low_h, low_w, low_c,
           m->C, mp.st_low.p);
    uf->low = mp.st_low.p;
  }
  uf->lh = low_h;
  uf->lw = low_w;
  uf->rh = (float)low_h / (float)height;  // area_pixel_compute_scale, align_corners = false, no scale_factor
  uf->rw = (float)low_w / (float)width;
  // which multiply-add contraction the chained torch kernel has (csrc/nvbx_upsample.cuh)
  *mode = dtype == NVBX_LOWRES_BF16 ? 2 : ((dtype == NVBX_LOWRES_F32 && kernel == NVBX_UPSAMPLE_TORCH_NHWC) ? 1 : 0);
  return NVBX_OK;
}

// features != nullptr: the [H, W, C] fp16 frame (a8).  Otherwise `uf` / `up_mode` describe the low-res map (N4).
struct HostFetch {  // nvbx_integrate_frame_host: `features` is a device buffer filled sparsely from this mapped host frame
  const uint4* host_img;
  unsigned* bitmap;
  int n_words;
};

int integrate_features_impl(nvbx_mapper* m, int map_id, const void* features, const UpFrame* uf, int up_mode,
                            int height, int width, const void* mask, const float* T_L_C_rm, float fx, float fy,
                            float cx, float cy, void* stream_v, const HostFetch* hf = nullptr) {
  int rc;
  const nvbx_params& p = m->params;
  cudaStream_t stream = (cudaStream_t)stream_v;
  Map& mp = *m->maps[map_id];
  mp.have_band_list = false;
  // A frame that cannot be pipelined (host-resident / low-res feature sources) runs its gather on the caller's
  // stream, behind every gather in flight; a pipelined one only needs its ring slot to be free (pipeline_slot_ready).
  bool pipe = m->pipelining && !hf && !uf && features != nullptr;
  if (!pipe) {
    if ((rc = pipeline_join(mp, stream))) return rc;
  } else if ((rc = pipeline_slot_ready(mp))) {
    return rc;
  }
  AppearancePrep prep;
  if ((rc = appearance_prepare(m, mp, mp.planes_cache, height, width, T_L_C_rm, fx, fy, cx, cy, -1, stream, &prep)))
    return rc;
  if (prep.empty) return NVBX_OK;
  const float trunc = prep.trunc;
  const Cam cam = prep.cam;
  const Pose T_C_L = prep.T_C_L;
  const int srows = prep.srows, scols = prep.scols;
  const long long cand_bound = prep.cand_bound;
  const long long chunk_blocks = std::min<long long>(cand_bound, kFeatureChunkBlocks);

  FeatFrame ff;
  ff.img = (const __half*)features;
  ff.mask = (const uint8_t*)mask;
  ff.synth = prep.synth;
  ff.rows = height;
  ff.cols = width;
  ff.srows = srows;
  ff.scols = scols;
  ff.sub = height / srows;  // projective_integrator_impl.cuh:424-425
  ff.cam = cam;
  ff.T_C_L = T_C_L;
  ff.max_depth = p.max_integration_distance_m;
  ff.trunc = trunc;
  ff.alpha = p.appearance_measurement_weight;
  ff.max_weight = p.max_weight;
  {
    float w1 = 1.0f - ff.alpha, w2 = ff.alpha;  // blendTwoArrays, projective_appearance_integrator.cu:286-305
    const float tot = w1 + w2;
    w1 /= tot;
    w2 /= tot;
    const __half h1 = __float2half_rn(w1), h2 = __float2half_rn(w2);
    std::memcpy(&ff.h_w1, &h1, 2);
    std::memcpy(&ff.h_w2, &h2, 2);
  }
  ff.read_old = (p.strict_blend || ff.alpha != 1.0f) ? 1 : 0;
  // Pixel rows loaded without allocating in L1 (ldg_nc_stream)?  Measured on the B200 (profiles/r02_pipelining.md):
  // a single map with voxels that project to several pixels (cube stacking 2 cm, the 1024^2 stress rig) gains
  // (gather alone 24.8 -> 23.3 us, stress step +11 %): the rows are used once and the L1 is better left to the depth
  // path; with voxels denser than the pixels (drill in box, 1 cm at 512^2: 1.6 distinct pixels per voxel) the L1 is what
  // merges the neighbours' repeated rows (0.67 -> 0.52 of the roofline without it), and with several maps on one GPU
  // the L2 is the bound and the merging helps too (-10 % without it).  Hence: stream when a voxel at 1 m covers at
  // least four pixels (fx x voxel size) and the frame does not come from a multi-map batch.  NVBX_GATHER_STREAM_LOADS=0 / 1 forces it.
  static const int stream_loads = env_int("NVBX_GATHER_STREAM_LOADS", -1);
  ff.stream_loads = stream_loads >= 0 ? (stream_loads != 0) : (fx * mp.voxel_size >= 4.0f && !t_multi_map_batch);
  if (pipe && cand_bound > chunk_blocks) {  // several geometry / gather passes share one item list: not pipelined
    pipe = false;
    if ((rc = pipeline_join(mp, stream))) return rc;
  }
  if (pipe && (rc = pipeline_init(mp))) return rc;
  cudaStream_t gs = pipe ? mp.gstream : stream;
  for (long long begin = 0; begin < cand_bound; begin += chunk_blocks) {
    const long long end = std::min(cand_bound, begin + chunk_blocks);
    const int last = end >= cand_bound ? 1 : 0;
    if (begin > 0) CUDA_TRY(cudaMemsetAsync(&mp.d_ctrl->item_count[mp.dev.fp], 0, sizeof(int), stream));
    const int ggrid = persistent_grid(m, 2);  // full grid: new feature blocks are zero-filled by all CTAs
    // tuning knob: where the geometry kernel of a pipelined frame runs -- 0: caller's stream, 1: the gather stream
    // (measured slower), 2: its own high-priority stream
    static const int geom_mode = env_int("NVBX_PIPE_GEOM", 2);
    const bool geom_on_gs = geom_mode == 1, geom_own = geom_mode == 2;
    cudaStream_t geo = stream;
    if (pipe && geom_on_gs) geo = gs;
    if (pipe && geom_own) geo = mp.cstream;
    if (pipe && geo != stream) {  // behind this frame's trace / band selection
      CUDA_TRY(cudaEventRecord(mp.ev_trace, stream));
      CUDA_TRY(cudaStreamWaitEvent(geo, mp.ev_trace, 0));
    }
    LAUNCH(k_feature_geometry, ggrid, 512, 0, geo, mp.dev, mp.band_slots2[mp.dev.fp].p, mp.newfeat_slots2[mp.dev.fp].p, ff,
           mp.items2[mp.dev.fp].p, (int)begin, (int)end);
    const int ch = (m->C % 256 == 0 && m->C / 256 >= 1 && m->C / 256 <= 4) ? m->C / 256 : 0;
    if (hf) {
      const int nvec = m->C / 8;
      uint4* dev_img = (uint4*)const_cast<void*>(features);
      LAUNCH(k_pixel_mark, persistent_grid(m, 2), 256, 0, stream, mp.dev, mp.items2[mp.dev.fp].p, hf->bitmap, width);
      const int fgrid = std::min(persistent_grid(m, 8), (hf->n_words + 7) / 8);
      if (ch == 3)
        LAUNCH(k_pixel_fetch<3>, fgrid, 256, 0, stream, mp.dev, hf->bitmap, hf->n_words, hf->host_img, dev_img, nvec);
      else if (ch == 4)
        LAUNCH(k_pixel_fetch<4>, fgrid, 256, 0, stream, mp.dev, hf->bitmap, hf->n_words, hf->host_img, dev_img, nvec);
      else
        LAUNCH(k_pixel_fetch<0>, fgrid, 256, 0, stream, mp.dev, hf->bitmap, hf->n_words, hf->host_img, dev_img, nvec);
    }
    if (uf) {
      if (ch == 3)
        rc = launch_gather_up<3>(m, mp, ff, *uf, up_mode, last, stream);
      else if (ch == 4)
        rc = launch_gather_up<4>(m, mp, ff, *uf, up_mode, last, stream);
      else
        rc = launch_gather_up<0>(m, mp, ff, *uf, up_mode, last, stream);
    } else {
      if (pipe && geo == stream) {
        CUDA_TRY(cudaEventRecord(mp.ev_trace, stream));
        CUDA_TRY(cudaStreamWaitEvent(gs, mp.ev_trace, 0));
      } else if (pipe && geo == mp.cstream) {
        CUDA_TRY(cudaEventRecord(mp.ev_geom, geo));
        CUDA_TRY(cudaStreamWaitEvent(gs, mp.ev_geom, 0));
      }
      if (ch == 3)
        rc = launch_gather<3>(m, mp, ff, last, gs, pipe);
      else if (ch == 4)
        rc = launch_gather<4>(m, mp, ff, last, gs, pipe);
      else if (ch == 2)
        rc = launch_gather<2>(m, mp, ff, last, gs, pipe);
      else if (ch == 1)
        rc = launch_gather<1>(m, mp, ff, last, gs, pipe);
      else
        rc = launch_gather<0>(m, mp, ff, last, gs, pipe);
      if (pipe && !rc) {
        CUDA_TRY(cudaEventRecord(mp.ev_gather[mp.dev.fp], gs));
        mp.gather_pending[mp.dev.fp] = true;
      }
    }
    if (rc) return rc;
  }
  mp.have_band_list = true;
  mp.dev.fp = (mp.dev.fp + 1) % kFrameRing;  // the next feature frame uses the next ring slot
  ++mp.dev.seq;
  return NVBX_OK;
}

}  // namespace

extern "C" {

int nvbx_integrate_features(nvbx_mapper* m, int map_id, const void* features, int height, int width, int channels,
                            const void* mask, const float* T_L_C_rm, float fx, float fy, float cx, float cy,
                            void* stream_v) {
  int rc = check_map(m, map_id, /*flush=*/false);
  if (rc) return rc;
  if (!features || !T_L_C_rm || height <= 0 || width <= 0) return fail(NVBX_ERR_INVALID_ARGUMENT, "bad feature frame");
  if (channels != m->C)
    return fail(NVBX_ERR_INVALID_ARGUMENT, "feature frame has %d channels, the map was created with %d", channels,
                m->C);
  if (((uintptr_t)features) & 15) return fail(NVBX_ERR_INVALID_ARGUMENT, "feature frame must be 16-byte aligned");
  if (m->async && !t_async_worker) {
    AsyncJob j{1, map_id, features, height, width, channels, mask, {}, fx, fy, cx, cy, stream_v};
    std::memcpy(j.T, T_L_C_rm, sizeof(j.T));
    m->async->push(j);
    return NVBX_OK;
  }
  return integrate_features_impl(m, map_id, features, nullptr, 0, height, width, mask, T_L_C_rm, fx, fy, cx, cy,
                                 stream_v);
}

}  // extern "C"
void AsyncEnqueue::loop() {
  t_async_worker = true;
  std::unique_lock<std::mutex> lk(mu_);
  for (;;) {
    cv_work_.wait(lk, [this] { return stop_ || !q_.empty(); });
    if (q_.empty()) return;  // stop requested and nothing left to issue
    const AsyncJob j = q_.front();
    q_.pop_front();
    const bool skip = rc_ != 0;  // after a failure the frames behind it are dropped, the error is reported by flush()
    lk.unlock();
    cv_room_.notify_one();
    int rc = NVBX_OK;
    if (!skip)
      rc = j.kind == 0 ? nvbx_integrate_depth(m_, j.map_id, j.image, j.height, j.width, j.mask, j.T, j.fx, j.fy, j.cx,
                                              j.cy, j.stream)
                       : nvbx_integrate_features(m_, j.map_id, j.image, j.height, j.width, j.channels, j.mask, j.T, j.fx,
                                                 j.fy, j.cx, j.cy, j.stream);
    lk.lock();
    if (rc < 0 && rc_ == 0) {
      rc_ = rc;
      err_ = g_last_error;  // this thread's message
    }
    if (--outstanding_ == 0) cv_done_.notify_all();
  }
}
extern "C" {

int nvbx_integrate_features_lowres(nvbx_mapper* m, int map_id, const void* lowres, int low_h, int low_w, int low_c,
                                   int dtype, int layout, int kernel, int height, int width, const void* mask,
                                   const float* T_L_C_rm, float fx, float fy, float cx, float cy, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  if (!T_L_C_rm) return fail(NVBX_ERR_INVALID_ARGUMENT, "bad feature frame");
  UpFrame uf;
  int mode = 0;
  if ((rc = prepare_lowres(m, *m->maps[map_id], lowres, low_h, low_w, low_c, dtype, layout, kernel, height, width,
                           (cudaStream_t)stream_v, &uf, &mode)))
    return rc;
  return integrate_features_impl(m, map_id, nullptr, &uf, mode, height, width, mask, T_L_C_rm, fx, fy, cx, cy,
                                 stream_v);
}

int nvbx_upsample_features(nvbx_mapper* m, int map_id, const void* lowres, int low_h, int low_w, int low_c, int dtype,
                           int layout, int kernel, int height, int width, void* out, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  if (!out || (((uintptr_t)out) & 15)) return fail(NVBX_ERR_INVALID_ARGUMENT, "output frame must be 16-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_v;
  Map& mp = *m->maps[map_id];
  UpFrame uf;
  int mode = 0;
  if ((rc = prepare_lowres(m, mp, lowres, low_h, low_w, low_c, dtype, layout, kernel, height, width, stream, &uf,
                           &mode)))
    return rc;
  const int grid = persistent_grid(m, 8);
  if (mode == 1)
    LAUNCH(k_upsample_materialise<1>, grid, 256, 0, stream, uf, height, width, m->C, (__half*)out);
  else if (mode == 2)
    LAUNCH(k_upsample_materialise<2>, grid, 256, 0, stream, uf, height, width, m->C, (__half*)out);
  else
    LAUNCH(k_upsample_materialise<0>, grid, 256, 0, stream, uf, height, width, m->C, (__half*)out);
  return NVBX_OK;
}

int nvbx_integrate_color(nvbx_mapper* m, int map_id, const void* rgb, int height, int width, const void* mask,
                         const float* T_L_C_rm, float fx, float fy, float cx, float cy, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  if (!rgb || !T_L_C_rm || height <= 0 || width <= 0) return fail(NVBX_ERR_INVALID_ARGUMENT, "bad colour frame");
  const nvbx_params& p = m->params;
  cudaStream_t stream = (cudaStream_t)stream_v;
  Map& mp = *m->maps[map_id];
  mp.have_cband_list = false;
  const int parity = mp.color_parity;
  if ((rc = pipeline_slot_ready(mp))) return rc;  // a traced synthetic depth image goes to synth2[fp]
  AppearancePrep prep;
  if ((rc = appearance_prepare(m, mp, mp.planes_cache, height, width, T_L_C_rm, fx, fy, cx, cy, parity, stream,
                               &prep)))
    return rc;
  if (prep.empty) return NVBX_OK;
  ColorFrame cf;
  cf.img = (const uint8_t*)rgb;
  cf.mask = (const uint8_t*)mask;
  cf.synth = prep.synth;
  cf.rows = height;
  cf.cols = width;
  cf.srows = prep.srows;
  cf.scols = prep.scols;
  cf.sub = prep.sub;
  cf.cam = prep.cam;
  cf.T_C_L = prep.T_C_L;
  cf.max_depth = p.max_integration_distance_m;
  cf.trunc = prep.trunc;
  cf.alpha = p.appearance_measurement_weight;
  cf.max_weight = p.max_weight;
  {
    float w1 = 1.0f - cf.alpha, w2 = cf.alpha;  // blendTwoArrays, projective_appearance_integrator.cu:286-305
    const float tot = w1 + w2;
    w1 /= tot;
    w2 /= tot;
    cf.w1 = __half2float(__float2half_rn(w1));  // weightedSum(uint8_t, ...) receives __float2half(weight)
    cf.w2 = __half2float(__float2half_rn(w2));
  }
  const int grid = (int)std::max(1LL, std::min<long long>(persistent_grid(m, 2), prep.cand_bound));
  LAUNCH(k_color_update, grid, 512, 0, stream, mp.dev, mp.cband_slots.p, cf, parity);
  mp.color_parity ^= 1;
  mp.have_cband_list = true;
  return NVBX_OK;
}

int nvbx_integrate_frame_host(nvbx_mapper* m, int map_id, const float* depth_host, const void* features_host,
                              int height, int width, int channels, const uint8_t* depth_mask_host,
                              const uint8_t* feature_mask_host, const float* T_L_C, float fx, float fy, float cx,
                              float cy, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  if (!depth_host || !features_host || !T_L_C || height <= 0 || width <= 0)
    return fail(NVBX_ERR_INVALID_ARGUMENT, "bad host frame (null buffer / pose or non-positive size)");
  cudaStream_t stream = (cudaStream_t)stream_v;
  Map& mp = *m->maps[map_id];
  const size_t px = (size_t)height * width;
  if ((rc = mp.st_depth.ensure(px, stream))) return rc;
  if (channels != m->C)
    return fail(NVBX_ERR_INVALID_ARGUMENT, "feature frame has %d channels, the map was created with %d", channels,
                m->C);
  if ((rc = mp.st_feat.ensure(px * channels, stream))) return rc;
  CUDA_TRY(cudaMemcpyAsync(mp.st_depth.p, depth_host, px * sizeof(float), cudaMemcpyHostToDevice, stream));
  // A pinned (device-mapped) frame is read sparsely by the GPU: only the pixels this frame's work items touch
  // cross PCIe (k_pixel_mark / k_pixel_fetch).  A pageable frame, or host_fetch_mode = dense, is copied whole.
  HostFetch hf{nullptr, nullptr, 0};
  if (m->host_fetch_mode == NVBX_HOST_FETCH_SPARSE && !(((uintptr_t)features_host) & 15)) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, features_host) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
        attr.devicePointer != nullptr) {
      const size_t n_words = (px + 31) / 32;
      const size_t had = mp.st_pixmap.cap;
      if ((rc = mp.st_pixmap.ensure(n_words, stream))) return rc;
      if (mp.st_pixmap.cap != had) CUDA_TRY(cudaMemsetAsync(mp.st_pixmap.p, 0, mp.st_pixmap.cap * sizeof(unsigned), stream));
      hf.host_img = (const uint4*)attr.devicePointer;
      hf.bitmap = mp.st_pixmap.p;
      hf.n_words = (int)n_words;
    } else {
      (void)cudaGetLastError();  // an unregistered pointer is not an error here
    }
  }
  if (!hf.host_img)
    CUDA_TRY(cudaMemcpyAsync(mp.st_feat.p, features_host, px * channels * sizeof(__half), cudaMemcpyHostToDevice, stream));
  const uint8_t* dm = nullptr;
  const uint8_t* fm = nullptr;
  if (depth_mask_host) {
    if ((rc = mp.st_mask_d.ensure(px, stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(mp.st_mask_d.p, depth_mask_host, px, cudaMemcpyHostToDevice, stream));
    dm = mp.st_mask_d.p;
  }
  if (feature_mask_host) {
    if ((rc = mp.st_mask_f.ensure(px, stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(mp.st_mask_f.p, feature_mask_host, px, cudaMemcpyHostToDevice, stream));
    fm = mp.st_mask_f.p;
  }
  if ((rc = nvbx_integrate_depth(m, map_id, mp.st_depth.p, height, width, dm, T_L_C, fx, fy, cx, cy, stream_v)))
    return rc;
  return integrate_features_impl(m, map_id, mp.st_feat.p, nullptr, 0, height, width, fm, T_L_C, fx, fy, cx, cy,
                                 stream_v, hf.host_img ? &hf : nullptr);
}

int nvbx_set_pipelining(nvbx_mapper* m, int on) {
  if (!m) return fail(NVBX_ERR_INVALID_ARGUMENT, "null mapper");
  if (on < 0 || on > 2) return fail(NVBX_ERR_INVALID_ARGUMENT, "pipelining mode %d (0: off, 1: on, 2: on + asynchronous enqueue)", on);
  if (int frc = async_flush(m)) return frc;
  if (on != 2) m->async.reset();
  if (!on)
    for (auto& mp : m->maps) {
      int rc = pipeline_drain(*mp);
      if (rc) return rc;
      for (int r = 0; r < kFrameRing; ++r) mp->gather_pending[r] = false;
    }
  m->pipelining = on != 0;
  if (on == 2 && !m->async) m->async.reset(new AsyncEnqueue(m));
  return NVBX_OK;
}
int nvbx_pipeline_join(nvbx_mapper* m, int map_id, void* stream_v) {
  if (!m) return fail(NVBX_ERR_INVALID_ARGUMENT, "null mapper");
  if (int frc = async_flush(m)) return frc;
  if (map_id < 0) {
    for (int i = 0; i < (int)m->maps.size(); ++i) {
      int rc = pipeline_join(*m->maps[i], (cudaStream_t)stream_v);
      if (rc) return rc;
    }
    return NVBX_OK;
  }
  int rc = check_map(m, map_id);
  if (rc) return rc;
  return pipeline_join(*m->maps[map_id], (cudaStream_t)stream_v);
}

int nvbx_set_host_fetch_mode(nvbx_mapper* m, int mode) {
  if (!m) return fail(NVBX_ERR_INVALID_ARGUMENT, "null mapper");
  if (mode != NVBX_HOST_FETCH_SPARSE && mode != NVBX_HOST_FETCH_DENSE)
    return fail(NVBX_ERR_INVALID_ARGUMENT, "bad host fetch mode %d", mode);
  m->host_fetch_mode = mode;
  return NVBX_OK;
}

int nvbx_integrate_frame_host_lowres(nvbx_mapper* m, int map_id, const float* depth_host, const void* lowres_host,
                                     int low_h, int low_w, int low_c, int dtype, int layout, int kernel, int height,
                                     int width, const uint8_t* depth_mask_host, const uint8_t* feature_mask_host,
                                     const float* T_L_C, float fx, float fy, float cx, float cy, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  if (!depth_host || !lowres_host || !T_L_C || low_h <= 0 || low_w <= 0 || low_c <= 0 || dtype < 0 || dtype > 2 ||
      height <= 0 || width <= 0)
    return fail(NVBX_ERR_INVALID_ARGUMENT, "bad host frame (null buffer / pose, bad dtype or non-positive size)");
  cudaStream_t stream = (cudaStream_t)stream_v;
  Map& mp = *m->maps[map_id];
  const size_t px = (size_t)height * width;
  const size_t low_bytes = (size_t)low_h * low_w * low_c * (dtype == NVBX_LOWRES_F32 ? 4 : 2);
  if ((rc = mp.st_depth.ensure(px, stream))) return rc;
  if ((rc = mp.st_low_in.ensure(low_bytes, stream))) return rc;
  CUDA_TRY(cudaMemcpyAsync(mp.st_depth.p, depth_host, px * sizeof(float), cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaMemcpyAsync(mp.st_low_in.p, lowres_host, low_bytes, cudaMemcpyHostToDevice, stream));
  const uint8_t* dm = nullptr;
  const uint8_t* fm = nullptr;
  if (depth_mask_host) {
    if ((rc = mp.st_mask_d.ensure(px, stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(mp.st_mask_d.p, depth_mask_host, px, cudaMemcpyHostToDevice, stream));
    dm = mp.st_mask_d.p;
  }
  if (feature_mask_host) {
    if ((rc = mp.st_mask_f.ensure(px, stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(mp.st_mask_f.p, feature_mask_host, px, cudaMemcpyHostToDevice, stream));
    fm = mp.st_mask_f.p;
  }
  if ((rc = nvbx_integrate_depth(m, map_id, mp.st_depth.p, height, width, dm, T_L_C, fx, fy, cx, cy, stream_v)))
    return rc;
  return nvbx_integrate_features_lowres(m, map_id, mp.st_low_in.p, low_h, low_w, low_c, dtype, layout, kernel, height,
                                        width, fm, T_L_C, fx, fy, cx, cy, stream_v);
}

// ---- decay / clear -------------------------------------------------------------------------------
static int decay_one(nvbx_mapper* m, Map& mp, cudaStream_t stream) {
  DecayParams dp;
  dp.factor = m->params.tsdf_decay_factor;
  dp.threshold = m->params.tsdf_decayed_weight_threshold;
  dp.set_free = m->params.tsdf_set_free_distance_on_decayed;
  dp.free_distance = m->params.tsdf_decayed_free_distance_vox * mp.voxel_size;
  dp.deallocate = m->params.deallocate_decayed_blocks;
  const int grid = std::max(1, (int)std::min<long long>(persistent_grid(m, 8), std::max(1LL, (long long)mp.slot_capacity)));
  ++mp.tsdf_version;
  LAUNCH(k_decay, grid, 256, 0, stream, mp.dev, dp);
  if (dp.deallocate) {
    LAUNCH(k_hash_clear, persistent_grid(m, 4), 256, 0, stream, mp.dev, 0);
    LAUNCH(k_hash_reinsert, persistent_grid(m, 4), 256, 0, stream, mp.dev, 0);
    LAUNCH(k_hash_rebuild_done, 1, 1, 0, stream, mp.dev);
  }
  return NVBX_OK;
}

int nvbx_decay(nvbx_mapper* m, int map_id, void* stream_v) {
  if (!m) return fail(NVBX_ERR_INVALID_ARGUMENT, "null mapper handle");
  if (int frc = async_flush(m)) return frc;
  if (map_id < 0) {
    for (int i = 0; i < (int)m->maps.size(); ++i) {
      int rc = nvbx_decay(m, i, stream_v);
      if (rc) return rc;
    }
    return NVBX_OK;
  }
  int rc = check_map(m, map_id);
  if (rc) return rc;
  if ((rc = pipeline_join(*m->maps[map_id], (cudaStream_t)stream_v))) return rc;
  return decay_one(m, *m->maps[map_id], (cudaStream_t)stream_v);
}

int nvbx_clear(nvbx_mapper* m, int map_id, void* stream_v) {
  if (!m) return fail(NVBX_ERR_INVALID_ARGUMENT, "null mapper handle");
  if (int frc = async_flush(m)) return frc;
  if (map_id < 0) {
    for (int i = 0; i < (int)m->maps.size(); ++i) {
      int rc = nvbx_clear(m, i, stream_v);
      if (rc) return rc;
    }
    return NVBX_OK;
  }
  int rc = check_map(m, map_id);
  if (rc) return rc;
  Map& mp = *m->maps[map_id];
  cudaStream_t stream = (cudaStream_t)stream_v;
  if ((rc = pipeline_join(mp, stream))) return rc;
  ++mp.tsdf_version;
  LAUNCH(k_clear_all, persistent_grid(m, 4), 256, 0, stream, mp.dev);
  mp.slot_used_ub = 0;
  mp.feat_used_ub = 0;
  mp.stray_ub = 0;
  mp.mesh_nv = 0;
  mp.mesh_nt = 0;
  mp.cmesh_nv = 0;
  mp.cmesh_nt = 0;
  // NOTE: the viewpoint caches and the to-update tracker survive, as in the reference
  // (py_mapper.cu:286-306 clears the layers only).
  return NVBX_OK;
}

int nvbx_mark_all_dirty(nvbx_mapper* m, int map_id, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  Map& mp = *m->maps[map_id];
  ++mp.tsdf_version;  // called after block payloads were written through layer views (load_from_file)
  LAUNCH(k_mark_all_dirty, persistent_grid(m, 4), 256, 0, (cudaStream_t)stream_v, mp.dev);
  return NVBX_OK;
}

// ---- mesh ----------------------------------------------------------------------------------------
static int update_mesh_one(nvbx_mapper* m, Map& mp, int kind, cudaStream_t stream) {
  int rc;
  const size_t cap = (size_t)mp.slot_capacity;
  if ((rc = mp.cnt_v.ensure(cap, stream))) return rc;
  if ((rc = mp.cnt_t.ensure(cap, stream))) return rc;
  if ((rc = mp.off_v.ensure(cap, stream))) return rc;
  if ((rc = mp.off_t.ensure(cap, stream))) return rc;
  if (kind == kMeshColor && (rc = enable_color(mp, stream))) return rc;
  MeshParams mpar;
  mpar.min_weight = m->params.mesh_min_weight;
  mpar.cutoff = m->params.mesh_cutoff_distance_vox * mp.voxel_size;
  mpar.weld = m->params.mesh_weld_vertices;
  const int grid = persistent_grid(m, 2);
  if (kind == kMeshFeature)
    LAUNCH(k_mesh_count<kMeshFeature>, grid, 512, sizeof(MeshSmem), stream, mp.dev, mpar, mp.cnt_v.p, mp.cnt_t.p);
  else
    LAUNCH(k_mesh_count<kMeshColor>, grid, 512, sizeof(MeshSmem), stream, mp.dev, mpar, mp.cnt_v.p, mp.cnt_t.p);
  LAUNCH(k_mesh_scan, 1, 1024, 0, stream, mp.dev, mp.cnt_v.p, mp.cnt_t.p, mp.off_v.p, mp.off_t.p, kind);
  if ((rc = read_ctrl(mp, stream))) return rc;  // the one synchronisation of the export path
  const long long nv = mp.h_ctrl->mesh_total_v, nt = mp.h_ctrl->mesh_total_t;
  mp.slot_used_ub = (long long)mp.h_ctrl->slot_high - mp.h_ctrl->slot_free_top;  // free refresh of the bounds
  mp.feat_used_ub = (long long)mp.h_ctrl->feat_high - mp.h_ctrl->feat_free_top;
  if (kind == kMeshFeature) {
    const int nxt = mp.arena_cur ^ 1;
    if ((rc = mp.arena_v[nxt].ensure((size_t)std::max(1LL, nv) * 3, stream))) return rc;
    if ((rc = mp.arena_f[nxt].ensure((size_t)std::max(1LL, nv) * m->C, stream))) return rc;
    if ((rc = mp.arena_t[nxt].ensure((size_t)std::max(1LL, nt), stream))) return rc;
    MeshArena prev{mp.arena_v[mp.arena_cur].p, mp.arena_f[mp.arena_cur].p, mp.arena_t[mp.arena_cur].p, nullptr};
    MeshArena out{mp.arena_v[nxt].p, mp.arena_f[nxt].p, mp.arena_t[nxt].p, nullptr};
    LAUNCH(k_mesh_emit<kMeshFeature>, grid, 512, sizeof(MeshSmem), stream, mp.dev, mpar, mp.off_v.p, mp.off_t.p, prev,
           out);
    mp.arena_cur = nxt;
    mp.mesh_nv = nv;
    mp.mesh_nt = nt;
  } else {
    const int nxt = mp.carena_cur ^ 1;
    if ((rc = mp.carena_v[nxt].ensure((size_t)std::max(1LL, nv) * 3, stream))) return rc;
    if ((rc = mp.carena_c[nxt].ensure((size_t)std::max(1LL, nv) * 3, stream))) return rc;
    if ((rc = mp.carena_t[nxt].ensure((size_t)std::max(1LL, nt), stream))) return rc;
    MeshArena prev{mp.carena_v[mp.carena_cur].p, nullptr, mp.carena_t[mp.carena_cur].p, mp.carena_c[mp.carena_cur].p};
    MeshArena out{mp.carena_v[nxt].p, nullptr, mp.carena_t[nxt].p, mp.carena_c[nxt].p};
    LAUNCH(k_mesh_emit<kMeshColor>, grid, 512, sizeof(MeshSmem), stream, mp.dev, mpar, mp.off_v.p, mp.off_t.p, prev,
           out);
    mp.carena_cur = nxt;
    mp.cmesh_nv = nv;
    mp.cmesh_nt = nt;
  }
  return NVBX_OK;
}

static int update_mesh(nvbx_mapper* m, int map_id, int kind, void* stream_v) {
  if (!m) return fail(NVBX_ERR_INVALID_ARGUMENT, "null mapper handle");
  if (map_id < 0) {
    for (int i = 0; i < (int)m->maps.size(); ++i) {
      int rc = update_mesh(m, i, kind, stream_v);
      if (rc) return rc;
    }
    return NVBX_OK;
  }
  int rc = check_map(m, map_id);
  if (rc) return rc;
  if ((rc = pipeline_join(*m->maps[map_id], (cudaStream_t)stream_v))) return rc;
  return update_mesh_one(m, *m->maps[map_id], kind, (cudaStream_t)stream_v);
}

int nvbx_update_feature_mesh(nvbx_mapper* m, int map_id, void* stream_v) {
  return update_mesh(m, map_id, kMeshFeature, stream_v);
}
int nvbx_update_color_mesh(nvbx_mapper* m, int map_id, void* stream_v) {
  return update_mesh(m, map_id, kMeshColor, stream_v);
}

int nvbx_get_feature_mesh(nvbx_mapper* m, int map_id, const void** vertices, const void** features,
                          const void** triangles, int64_t* n_vertices, int64_t* n_triangles) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  Map& mp = *m->maps[map_id];
  if (vertices) *vertices = mp.arena_v[mp.arena_cur].p;
  if (features) *features = mp.arena_f[mp.arena_cur].p;
  if (triangles) *triangles = mp.arena_t[mp.arena_cur].p;
  if (n_vertices) *n_vertices = mp.mesh_nv;
  if (n_triangles) *n_triangles = mp.mesh_nt / 3;
  return NVBX_OK;
}

int nvbx_get_color_mesh(nvbx_mapper* m, int map_id, const void** vertices, const void** colors,
                        const void** triangles, int64_t* n_vertices, int64_t* n_triangles) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  Map& mp = *m->maps[map_id];
  if (vertices) *vertices = mp.carena_v[mp.carena_cur].p;
  if (colors) *colors = mp.carena_c[mp.carena_cur].p;
  if (triangles) *triangles = mp.carena_t[mp.carena_cur].p;
  if (n_vertices) *n_vertices = mp.cmesh_nv;
  if (n_triangles) *n_triangles = mp.cmesh_nt / 3;
  return NVBX_OK;
}

// ---- fused export post-processing (N2) -----------------------------------------------------------------
int64_t nvbx_export_points(nvbx_mapper* m, int map_id, const void* vertices, const void* features, int64_t n,
                           int channels, const float* aabb_min, const float* aabb_max, int num_excess_features,
                           int remove_zero_features, const void** out_vertices, const void** out_features,
                           void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  if (!aabb_min || !aabb_max) return fail(NVBX_ERR_INVALID_ARGUMENT, "null AABB");
  Map& mp = *m->maps[map_id];
  cudaStream_t stream = (cudaStream_t)stream_v;
  if (!vertices) {  // the map's current feature mesh
    vertices = mp.arena_v[mp.arena_cur].p;
    features = mp.arena_f[mp.arena_cur].p;
    n = mp.mesh_nv;
    channels = m->C;
  }
  if (n < 0 || channels <= 0 || num_excess_features < 0 || num_excess_features > channels || (n > 0 && !features))
    return fail(NVBX_ERR_INVALID_ARGUMENT, "bad point cloud (n %lld, channels %d, excess %d)", (long long)n, channels,
                num_excess_features);
  ExportParams p;
  p.verts = (const float*)vertices;
  p.feats = (const __half*)features;
  p.n = n;
  p.C = channels;
  // features[..., :-num_excess] with num_excess == 0 keeps everything (nvblox_output_helpers.py:66-67)
  p.C_keep = channels - num_excess_features;
  for (int i = 0; i < 3; ++i) {
    p.mn[i] = aabb_min[i];
    p.mx[i] = aabb_max[i];
  }
  p.remove_zero = remove_zero_features ? 1 : 0;
  p.vec_in = (channels % 8 == 0) && (((uintptr_t)features & 15) == 0);
  p.vec_out = (p.C_keep % 8 == 0);
  mp.exp_count = 0;
  mp.exp_C_keep = p.C_keep;
  if (out_vertices) *out_vertices = nullptr;
  if (out_features) *out_features = nullptr;
  if (n == 0) return 0;
  const long long n_tiles = (n + kExportTile - 1) / kExportTile;
  if ((rc = mp.exp_keep.ensure((size_t)n, stream))) return rc;
  if ((rc = mp.exp_tile_cnt.ensure((size_t)n_tiles, stream))) return rc;
  if ((rc = mp.exp_tile_off.ensure((size_t)n_tiles, stream))) return rc;
  if ((rc = mp.exp_v.ensure((size_t)n * 3, stream))) return rc;
  if ((rc = mp.exp_f.ensure((size_t)n * (size_t)std::max(1, p.C_keep), stream))) return rc;
  const int grid = (int)std::min<long long>(n_tiles, persistent_grid(m, 8));
  LAUNCH(k_export_flag, grid, 256, 0, stream, p, mp.exp_keep.p, mp.exp_tile_cnt.p, n_tiles);
  LAUNCH(k_export_scan, 1, 1024, 0, stream, mp.exp_tile_cnt.p, mp.exp_tile_off.p, n_tiles, mp.d_exp_total);
  LAUNCH(k_export_scatter, grid, 256, 0, stream, p, mp.exp_keep.p, mp.exp_tile_off.p, n_tiles, mp.exp_v.p,
         mp.exp_f.p);
  long long total = 0;
  CUDA_TRY(cudaMemcpyAsync(&total, mp.d_exp_total, sizeof(total), cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  mp.exp_count = total;
  if (out_vertices) *out_vertices = mp.exp_v.p;
  if (out_features) *out_features = mp.exp_f.p;
  return total;
}

int nvbx_gather_points(nvbx_mapper* m, int map_id, const int64_t* indices, int64_t n_indices, int64_t n_out,
                       void* out_vertices, void* out_features, int features_f32, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  Map& mp = *m->maps[map_id];
  cudaStream_t stream = (cudaStream_t)stream_v;
  if (n_indices < 0 || n_out < n_indices || (n_out > 0 && (!out_vertices || !out_features)))
    return fail(NVBX_ERR_INVALID_ARGUMENT, "bad gather request (n_indices %lld, n_out %lld)", (long long)n_indices,
                (long long)n_out);
  if (!indices && n_indices > mp.exp_count)
    return fail(NVBX_ERR_INVALID_ARGUMENT, "identity gather of %lld rows from %lld exported points",
                (long long)n_indices, mp.exp_count);
  if (n_out == 0) return NVBX_OK;
  const long long* d_idx = nullptr;
  if (indices && n_indices > 0) {
    for (int64_t i = 0; i < n_indices; ++i)
      if (indices[i] < 0 || indices[i] >= mp.exp_count)
        return fail(NVBX_ERR_INVALID_ARGUMENT, "index %lld out of range [0, %lld)", (long long)indices[i], mp.exp_count);
    if ((rc = mp.exp_idx.ensure((size_t)n_indices, stream))) return rc;
    // pageable source: the copy is staged by the runtime before the call returns
    CUDA_TRY(cudaMemcpyAsync(mp.exp_idx.p, indices, (size_t)n_indices * sizeof(long long), cudaMemcpyHostToDevice,
                             stream));
    d_idx = mp.exp_idx.p;
  }
  const int C_keep = mp.exp_C_keep;
  const int vec = (C_keep % 8 == 0) && (((uintptr_t)out_features & 15) == 0) ? 1 : 0;
  const int grid = (int)std::max<long long>(1, std::min<long long>((n_out + 7) / 8, persistent_grid(m, 8)));
  if (features_f32)
    LAUNCH(k_export_gather<true>, grid, 256, 0, stream, mp.exp_v.p, mp.exp_f.p, C_keep, d_idx, (long long)n_indices,
           (long long)n_out, (float*)out_vertices, out_features, vec);
  else
    LAUNCH(k_export_gather<false>, grid, 256, 0, stream, mp.exp_v.p, mp.exp_f.p, C_keep, d_idx, (long long)n_indices,
           (long long)n_out, (float*)out_vertices, out_features, vec);
  return NVBX_OK;
}

// ---- layer views ---------------------------------------------------------------------------------
int64_t nvbx_num_blocks(nvbx_mapper* m, int map_id, int layer, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  Map& mp = *m->maps[map_id];
  if ((rc = read_ctrl(mp, (cudaStream_t)stream_v))) return rc;
  if (layer == NVBX_LAYER_COLOR) return mp.h_ctrl->n_color;
  return layer == NVBX_LAYER_TSDF ? mp.h_ctrl->n_tsdf : mp.h_ctrl->n_feat;
}
int64_t nvbx_num_allocated_blocks(nvbx_mapper* m, int map_id, int layer, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  (void)stream_v;
  Map& mp = *m->maps[map_id];
  if (layer == NVBX_LAYER_COLOR) return mp.color_enabled ? mp.slot_capacity : 0;
  return layer == NVBX_LAYER_TSDF ? mp.slot_capacity : mp.feat_capacity;
}
int64_t nvbx_num_allocated_bytes(nvbx_mapper* m, int map_id, int layer, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  (void)stream_v;
  Map& mp = *m->maps[map_id];
  if (layer == NVBX_LAYER_TSDF) return (int64_t)mp.slot_capacity * kVoxelsPerBlock * (int64_t)sizeof(float2);
  if (layer == NVBX_LAYER_COLOR)
    return mp.color_enabled ? (int64_t)mp.slot_capacity * kVoxelsPerBlock * (int64_t)sizeof(uint2) : 0;
  return (int64_t)mp.feat_capacity * kVoxelsPerBlock * (int64_t)mp.dev.row * (int64_t)sizeof(__half);
}

// Blocks of a layer in DESCENDING slot order.  Slot ids are handed out in increasing order, so this is reverse
// allocation order -- what iterating the reference's std::unordered_map gives for a freshly built layer (each new
// node goes to the head of the node list; NT tests/test_layer.py:101-138 rely on it) -- and it is deterministic,
// which the arrival order of the collecting kernel's atomics is not.
int64_t nvbx_get_all_blocks(nvbx_mapper* m, int map_id, int layer, int32_t* out_xyz, void** out_ptrs, int64_t capacity,
                            int64_t* voxel_stride_elems, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  Map& mp = *m->maps[map_id];
  cudaStream_t stream = (cudaStream_t)stream_v;
  if (voxel_stride_elems)
    *voxel_stride_elems = layer == NVBX_LAYER_TSDF ? 2 : (layer == NVBX_LAYER_COLOR ? 8 : mp.dev.row);
  if (layer == NVBX_LAYER_COLOR && !mp.color_enabled) return 0;
  if (layer == NVBX_LAYER_TSDF && out_ptrs) {  // the caller may write through the returned views
    ++mp.tsdf_version;
    LAUNCH(k_clear_free_bits, persistent_grid(m, 2), 256, 0, stream, mp.dev);
  }
  if (layer == NVBX_LAYER_FEATURE && out_ptrs && (rc = pipeline_join(mp, stream))) return rc;
  if ((rc = mp.idx_out.ensure((size_t)mp.slot_capacity, stream))) return rc;
  if ((rc = mp.ptr_out.ensure((size_t)mp.slot_capacity, stream))) return rc;
  if ((rc = mp.slot_out.ensure((size_t)mp.slot_capacity, stream))) return rc;
  CUDA_TRY(cudaMemsetAsync(&mp.d_ctrl->list_count, 0, sizeof(int), stream));
  LAUNCH(k_collect_blocks, persistent_grid(m, 4), 256, 0, stream, mp.dev, layer, mp.idx_out.p, mp.ptr_out.p,
         mp.slot_out.p, mp.slot_capacity);
  if ((rc = read_ctrl(mp, stream))) return rc;
  const int64_t n = mp.h_ctrl->list_count;
  if (capacity > 0 && n > 0 && (out_xyz || out_ptrs)) {
    std::vector<int3> idx((size_t)n);
    std::vector<unsigned long long> ptrs((size_t)n);
    std::vector<int> slots((size_t)n);
    CUDA_TRY(cudaMemcpyAsync(idx.data(), mp.idx_out.p, (size_t)n * sizeof(int3), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(ptrs.data(), mp.ptr_out.p, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                             stream));
    CUDA_TRY(cudaMemcpyAsync(slots.data(), mp.slot_out.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    std::vector<int> order((size_t)n);
    for (int64_t i = 0; i < n; ++i) order[(size_t)i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return slots[(size_t)a] > slots[(size_t)b]; });
    const int64_t c = std::min(n, capacity);
    for (int64_t i = 0; i < c; ++i) {
      const size_t k = (size_t)order[(size_t)i];
      if (out_xyz) {
        out_xyz[3 * i] = idx[k].x;
        out_xyz[3 * i + 1] = idx[k].y;
        out_xyz[3 * i + 2] = idx[k].z;
      }
      if (out_ptrs) out_ptrs[i] = (void*)ptrs[k];
    }
  }
  return n;
}

int64_t nvbx_get_block_indices(nvbx_mapper* m, int map_id, int layer, int32_t* out_xyz, int64_t capacity,
                               void* stream_v) {
  return nvbx_get_all_blocks(m, map_id, layer, out_xyz, nullptr, capacity, nullptr, stream_v);
}

int nvbx_get_block_ptr(nvbx_mapper* m, int map_id, int layer, int x, int y, int z, void** ptr,
                       int64_t* voxel_stride_elems, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  if (!ptr) return fail(NVBX_ERR_INVALID_ARGUMENT, "ptr is null");
  Map& mp = *m->maps[map_id];
  cudaStream_t stream = (cudaStream_t)stream_v;
  if (layer == NVBX_LAYER_COLOR && !mp.color_enabled) return fail(NVBX_ERR_NOT_FOUND, "block (%d, %d, %d) is not allocated", x, y, z);
  if (layer == NVBX_LAYER_TSDF) {  // the caller may write through the returned view
    ++mp.tsdf_version;
    LAUNCH(k_clear_free_bits, persistent_grid(m, 2), 256, 0, stream, mp.dev);
  }
  if (layer == NVBX_LAYER_FEATURE && (rc = pipeline_join(mp, stream))) return rc;
  LAUNCH(k_find_one, 1, 1, 0, stream, mp.dev, x, y, z, layer, mp.d_tmp_ptr);
  unsigned long long h = 0;
  CUDA_TRY(cudaMemcpyAsync(&h, mp.d_tmp_ptr, sizeof(h), cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  *ptr = (void*)h;
  if (voxel_stride_elems)
    *voxel_stride_elems = layer == NVBX_LAYER_TSDF ? 2 : (layer == NVBX_LAYER_COLOR ? 8 : mp.dev.row);
  if (!h) return fail(NVBX_ERR_NOT_FOUND, "block (%d, %d, %d) is not allocated", x, y, z);
  return NVBX_OK;
}

int nvbx_allocate_block(nvbx_mapper* m, int map_id, int layer, int x, int y, int z, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  if (!key_in_range(x, y, z)) return fail(NVBX_ERR_INVALID_ARGUMENT, "block index out of range");
  Map& mp = *m->maps[map_id];
  cudaStream_t stream = (cudaStream_t)stream_v;
  if ((rc = pipeline_join(mp, stream))) return rc;
  if (mp.dev.ws_cells > 0) {  // an index outside the workspace grid goes to the overflow hash: count it, keep the
                              // table's load below 0.25 (re-hash through grow_slots when it would not be)
    const I3 mn = mp.dev.ws_mn;
    const bool inside = x >= mn.x && x < mn.x + mp.dev.ws_sx && y >= mn.y && y < mn.y + mp.dev.ws_sy && z >= mn.z &&
                        z < mn.z + mp.dev.ws_sz;
    if (!inside) {
      ++mp.stray_ub;
      if (4 * mp.stray_ub > (long long)mp.dev.hash_mask + 1 && mp.slot_capacity > 0)
        if ((rc = grow_slots(m, mp, mp.slot_capacity, stream))) return rc;
    }
  }
  if ((rc = ensure_slots(m, mp, 1, stream))) return rc;
  ++mp.tsdf_version;
  if (layer == NVBX_LAYER_COLOR) {
    if ((rc = enable_color(mp, stream))) return rc;
  } else if (layer != NVBX_LAYER_TSDF) {
    // a stand-alone feature block: capacity is bounded by slots, so size it directly
    if (mp.feat_used_ub + 1 > mp.feat_capacity)
      if ((rc = grow_feats(m, mp, mp.feat_capacity + (1 << kFeatSlabShift), stream))) return rc;
    mp.feat_used_ub += 1;
  }
  LAUNCH(k_allocate_one, 1, 1, 0, stream, mp.dev, x, y, z, layer, mp.d_tmp_int);
  if (layer == NVBX_LAYER_FEATURE) LAUNCH(k_zero_one_feature_block, 64, 256, 0, stream, mp.dev, mp.d_tmp_int);
  return NVBX_OK;
}

// ---- queries -------------------------------------------------------------------------------------
int nvbx_query_tsdf(nvbx_mapper* m, int map_id, const void* xyz, int64_t n, void* out, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!xyz || !out))) return fail(NVBX_ERR_INVALID_ARGUMENT, "bad query buffers");
  if (n == 0) return NVBX_OK;
  Map& mp = *m->maps[map_id];
  LAUNCH(k_query_tsdf, (unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream_v, mp.dev, (const float*)xyz,
         (long long)n, (float2*)out);
  return NVBX_OK;
}
int nvbx_query_features(nvbx_mapper* m, int map_id, const void* xyz, int64_t n, void* out, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!xyz || !out))) return fail(NVBX_ERR_INVALID_ARGUMENT, "bad query buffers");
  if (n == 0) return NVBX_OK;
  Map& mp = *m->maps[map_id];
  if ((rc = pipeline_join(mp, (cudaStream_t)stream_v))) return rc;
  LAUNCH(k_query_features, (unsigned)((n * 32 + 127) / 128), 128, 0, (cudaStream_t)stream_v, mp.dev,
         (const float*)xyz, (long long)n, (__half*)out);
  return NVBX_OK;
}

// ---- accounting ----------------------------------------------------------------------------------
int nvbx_get_counters(nvbx_mapper* m, int map_id, nvbx_counters* out, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  if (!out) return fail(NVBX_ERR_INVALID_ARGUMENT, "out is null");
  Map& mp = *m->maps[map_id];
  if ((rc = pipeline_join(mp, (cudaStream_t)stream_v))) return rc;  // the geometry kernel counts on the gather stream
  if ((rc = read_ctrl(mp, (cudaStream_t)stream_v))) return rc;
  const unsigned long long* c = mp.h_ctrl->counters;
  std::memset(out, 0, sizeof(*out));
  out->depth_frames = (int64_t)c[kCntDepthFrames];
  out->feature_frames = (int64_t)c[kCntFeatureFrames];
  out->tsdf_blocks_in_view = (int64_t)c[kCntTsdfBlocksInView];
  out->tsdf_voxels_updated = (int64_t)c[kCntTsdfVoxelsUpdated];
  out->tsdf_blocks_allocated = (int64_t)c[kCntTsdfBlocksAllocated];
  out->feature_candidate_blocks = (int64_t)c[kCntFeatCandidateBlocks];
  out->feature_band_blocks = (int64_t)c[kCntFeatBandBlocks];
  out->feature_voxels_updated = (int64_t)c[kCntFeatVoxelsUpdated];
  out->feature_blocks_allocated = (int64_t)c[kCntFeatBlocksAllocated];
  out->blocks_deallocated = (int64_t)c[kCntBlocksDeallocated];
  out->mesh_blocks_remeshed = (int64_t)c[kCntMeshBlocksRemeshed];
  out->mesh_vertices = (int64_t)c[kCntMeshVertices];
  out->color_frames = (int64_t)c[kCntColorFrames];
  out->color_band_blocks = (int64_t)c[kCntColorBandBlocks];
  out->color_voxels_updated = (int64_t)c[kCntColorVoxelsUpdated];
  out->color_blocks_allocated = (int64_t)c[kCntColorBlocksAllocated];
  out->host_pixels_fetched = (int64_t)c[kCntHostPixelsFetched];
  for (int i = 0; i < 4; ++i) out->reserved[i] = (int64_t)c[kCntProfile0 + i];  // NVBX_PROFILE_COUNTERS builds only
  return NVBX_OK;
}
int nvbx_reset_counters(nvbx_mapper* m, int map_id, void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  Map& mp = *m->maps[map_id];
  if ((rc = pipeline_join(mp, (cudaStream_t)stream_v))) return rc;
  CUDA_TRY(cudaMemsetAsync(mp.d_ctrl->counters, 0, sizeof(unsigned long long) * kCntNum, (cudaStream_t)stream_v));
  mp.dev.seq = 0;  // the timeline stamps of profile builds are numbered from here
  return NVBX_OK;
}

// ---- batched frames: a small persistent pool of host threads issues the launches of different maps concurrently
namespace {
class LaunchPool {
 public:
  // Runs fn(0) .. fn(n - 1), fn(0) on the calling thread; returns when all have returned.
  void run(int n, const std::function<void(int)>& fn) {
    if (n <= 1) {
      if (n == 1) fn(0);
      return;
    }
    std::unique_lock<std::mutex> lk(mu_);
    while ((int)workers_.size() < n - 1) {
      const int id = (int)workers_.size() + 1;
      workers_.emplace_back([this, id] { loop(id); });
    }
    fn_ = &fn;
    n_active_ = n;
    pending_ = n - 1;
    ++epoch_;
    lk.unlock();
    cv_work_.notify_all();
    fn(0);
    lk.lock();
    cv_done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void loop(int id) {
    unsigned seen = 0;
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      cv_work_.wait(lk, [&] { return epoch_ != seen; });
      seen = epoch_;
      if (id >= n_active_) continue;
      const std::function<void(int)>* fn = fn_;
      lk.unlock();
      (*fn)(id);
      lk.lock();
      if (--pending_ == 0) cv_done_.notify_one();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_work_, cv_done_;
  std::vector<std::thread> workers_;
  const std::function<void(int)>* fn_ = nullptr;
  int n_active_ = 0, pending_ = 0;
  unsigned epoch_ = 0;
};
LaunchPool& launch_pool() {
  static LaunchPool* pool = new LaunchPool;  // never destroyed: its threads sleep until the process exits
  return *pool;
}
std::mutex g_batch_mu;  // one batch at a time per process (the pool's dispatch state is single-batch)
}  // namespace

static_assert(sizeof(nvbx_frame_job) == 152, "nvbx_frame_job layout is part of the ABI (params.py: NvbxFrameJob)");

int nvbx_integrate_frames_batch(nvbx_frame_job* jobs, int n_jobs, int n_threads) {
  if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return fail(NVBX_ERR_INVALID_ARGUMENT, "bad job list");
  if (n_jobs == 0) return NVBX_OK;
  // distinct handles in order of first appearance; every job of a handle goes to the same worker, in job order
  std::vector<nvbx_mapper*> handles;
  std::vector<int> owner(n_jobs);
  for (int i = 0; i < n_jobs; ++i) {
    if (!jobs[i].mapper) return fail(NVBX_ERR_INVALID_ARGUMENT, "job %d: null mapper handle", i);
    size_t k = 0;
    while (k < handles.size() && handles[k] != jobs[i].mapper) ++k;
    if (k == handles.size()) handles.push_back(jobs[i].mapper);
    owner[i] = (int)k;
    jobs[i].status = NVBX_OK;
  }
  const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
  const int n_workers = std::max(1, std::min({n_threads <= 0 ? hw : n_threads, (int)handles.size(), 64}));
  std::mutex err_mu;
  std::string first_error;
  int first_rc = NVBX_OK, first_job = n_jobs;
  auto work = [&](int w) {
    t_multi_map_batch = handles.size() > 1;
    for (int i = 0; i < n_jobs; ++i) {
      if (owner[i] % n_workers != w) continue;
      nvbx_frame_job& j = jobs[i];
      int rc = nvbx_integrate_depth(j.mapper, j.map_id, j.depth, j.height, j.width, j.depth_mask, j.T_L_C, j.fx, j.fy,
                                    j.cx, j.cy, j.stream);
      if (rc == NVBX_OK && j.features)
        rc = nvbx_integrate_features(j.mapper, j.map_id, j.features, j.height, j.width, j.channels, j.feature_mask,
                                     j.T_L_C, j.fx, j.fy, j.cx, j.cy, j.stream);
      j.status = rc;
      if (rc < 0) {
        std::lock_guard<std::mutex> g(err_mu);
        if (i < first_job) {
          first_job = i;
          first_rc = rc;
          first_error = g_last_error;  // this worker's thread-local message
        }
      }
    }
    t_multi_map_batch = false;
  };
  {
    std::lock_guard<std::mutex> g(g_batch_mu);
    launch_pool().run(n_workers, work);
  }
  if (first_rc < 0) return fail(first_rc, "job %d: %s", first_job, first_error.c_str());
  return NVBX_OK;
}

int nvbx_set_gather_tuning(int variant, int dyn_permille, int ticket_units) {
  if (variant < 0 || variant > 10 || dyn_permille < -1 || dyn_permille > 1000 || ticket_units < 1 || ticket_units > 64)
    return fail(NVBX_ERR_INVALID_ARGUMENT, "gather tuning out of range (variant 0..10, permille -1..1000, ticket 1..64)");
  g_gather_tuning[0] = variant;
  g_gather_tuning[1] = dyn_permille;
  g_gather_tuning[2] = ticket_units;
  return NVBX_OK;
}

int nvbx_set_kernel_timing(nvbx_mapper* m, int enabled) {
  if (!m) return fail(NVBX_ERR_INVALID_ARGUMENT, "null mapper handle");
  if (int frc = async_flush(m)) return frc;
  m->timing = enabled != 0;
  g_timing_mode = enabled;
  return NVBX_OK;
}

int64_t nvbx_kernel_timing_report(nvbx_mapper* m, char* json, int64_t capacity) {
  if (!m || !json || capacity <= 2) return fail(NVBX_ERR_INVALID_ARGUMENT, "bad timing report buffer");
  if (int frc = async_flush(m)) return frc;
  CUDA_TRY(cudaSetDevice(m->device));
  std::vector<std::pair<std::string, std::pair<double, long long>>> agg;
  std::lock_guard<std::mutex> recs_guard(g_launch_recs_mu);
  for (auto& r : g_launch_recs) {
    float ms = 0.f;
    if (r.b && cudaEventSynchronize(r.b) == cudaSuccess) cudaEventElapsedTime(&ms, r.a, r.b);
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
    std::string name = r.name;
    const size_t par = name.find('(');  // "(k_feature_gather<CH, 1, 4>)" -> strip the macro parentheses
    if (par == 0) name = name.substr(1, name.size() - 2);
    auto it = std::find_if(agg.begin(), agg.end(), [&](auto& e) { return e.first == name; });
    if (it == agg.end()) {
      agg.push_back({name, {ms, 1}});
    } else {
      it->second.first += ms;
      it->second.second += 1;
    }
  }
  g_launch_recs.clear();
  std::string out = "{";
  for (size_t i = 0; i < agg.size(); ++i) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s\"%s\": [%.6f, %lld]", i ? ", " : "", agg[i].first.c_str(), agg[i].second.first,
             agg[i].second.second);
    out += buf;
  }
  out += "}";
  if ((int64_t)out.size() + 1 > capacity) return fail(NVBX_ERR_INVALID_ARGUMENT, "timing report buffer too small");
  std::memcpy(json, out.c_str(), out.size() + 1);
  return (int64_t)out.size();
}
int nvbx_get_kernel_timing(nvbx_mapper* m, int which, double* total_ms, int64_t* launches) {
  if (!m || which < 0 || which > 1) return fail(NVBX_ERR_INVALID_ARGUMENT, "bad timing query");
  if (int frc = async_flush(m)) return frc;
  CUDA_TRY(cudaSetDevice(m->device));
  double sum = 0.0;
  int64_t n = 0;
  for (auto& ev : m->timing_events[which]) {
    CUDA_TRY(cudaEventSynchronize(ev.second));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, ev.first, ev.second));
    sum += ms;
    ++n;
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  m->timing_events[which].clear();
  if (total_ms) *total_ms = sum;
  if (launches) *launches = n;
  return NVBX_OK;
}

int64_t nvbx_debug_last_block_list(nvbx_mapper* m, int map_id, int which, int32_t* out_xyz, int64_t capacity,
                                   void* stream_v) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  Map& mp = *m->maps[map_id];
  cudaStream_t stream = (cudaStream_t)stream_v;
  if (which == 0) {
    if (!mp.last_depth_entry) return 0;
    int n = 0;
    CUDA_TRY(cudaMemcpyAsync(&n, mp.last_depth_entry->d_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (out_xyz && capacity > 0) {
      CUDA_TRY(cudaMemcpyAsync(out_xyz, mp.last_depth_entry->idx.p, (size_t)std::min<int64_t>(n, capacity) * sizeof(int3),
                               cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
    }
    return n;
  }
  if (which == 2 ? !mp.have_cband_list : !mp.have_band_list) return 0;
  if ((rc = pipeline_join(mp, stream))) return rc;  // the geometry kernel (gather stream) records last_band_count
  if ((rc = read_ctrl(mp, stream))) return rc;
  const int n = which == 2 ? mp.h_ctrl->last_cband_count : mp.h_ctrl->last_band_count;
  if (out_xyz && capacity > 0 && n > 0) {
    std::vector<int> slots((size_t)n);
    CUDA_TRY(cudaMemcpyAsync(slots.data(), which == 2 ? mp.cband_slots.p : mp.band_slots2[(mp.dev.fp + kFrameRing - 1) % kFrameRing].p, (size_t)n * sizeof(int),
                             cudaMemcpyDeviceToHost, stream));
    std::vector<int3> all((size_t)mp.slot_capacity);
    CUDA_TRY(cudaMemcpyAsync(all.data(), mp.dev.blk_index, all.size() * sizeof(int3), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    for (int64_t i = 0; i < std::min<int64_t>(n, capacity); ++i) {
      const int3 b = all[slots[i] & kSlotMask];
      out_xyz[3 * i] = b.x;
      out_xyz[3 * i + 1] = b.y;
      out_xyz[3 * i + 2] = b.z;
    }
  }
  return n;
}

int nvbx_debug_profile_stamps(nvbx_mapper* m, uint64_t* out, int reset) {
  if (!m) return fail(NVBX_ERR_INVALID_ARGUMENT, "null mapper handle");
  if (int frc = async_flush(m)) return frc;
#ifdef NVBX_PROFILE_COUNTERS
  CUDA_TRY(cudaSetDevice(m->device));
  CUDA_TRY(cudaDeviceSynchronize());
  constexpr int n = 64 * 2 * kProfKernels;
  if (out) CUDA_TRY(cudaMemcpyFromSymbol(out, g_prof, n * sizeof(unsigned long long)));
  if (out && getenv("NVBX_PROF_CTAS")) {  // per-CTA view of the raycast kernel of the last frames (stderr)
    static unsigned long long cta[4][256][5];
    CUDA_TRY(cudaMemcpyFromSymbol(cta, g_prof_cta, sizeof(cta)));
    for (int fr = 0; fr < 4; ++fr) {
      unsigned long long t0 = ~0ull;
      for (int c = 0; c < 256; ++c)
        if (cta[fr][c][0]) t0 = std::min(t0, cta[fr][c][0]);
      for (int c = 0; c < 256; ++c)
        if (cta[fr][c][0])
          fprintf(stderr, "raycast frame%%4=%d cta %3d sm %3llu: entry %+7.2f marched %+7.2f wait-passed %+7.2f us\n", fr, c,
                  cta[fr][c][4], (cta[fr][c][0] - t0) / 1e3, (cta[fr][c][1] - t0) / 1e3, (cta[fr][c][2] - t0) / 1e3);
    }
  }
  if (reset) {
    std::vector<unsigned long long> init(n);
    for (int i = 0; i < n; ++i) init[i] = (i & 1) ? 0ULL : ~0ULL;  // (min begin, max end) pairs
    CUDA_TRY(cudaMemcpyToSymbol(g_prof, init.data(), n * sizeof(unsigned long long)));
  }
  return 64 * kProfKernels;
#else
  (void)out;
  (void)reset;
  return fail(NVBX_ERR_UNSUPPORTED, "timeline stamps need the NVBX_PROFILE=1 build (libnvbx_prof.so)");
#endif
}

int nvbx_debug_last_synthetic_depth(nvbx_mapper* m, int map_id, const void** ptr, int* rows, int* cols) {
  int rc = check_map(m, map_id);
  if (rc) return rc;
  Map& mp = *m->maps[map_id];
  if (ptr) *ptr = mp.synth2[mp.synth_last].p;
  if (rows) *rows = mp.synth_rows;
  if (cols) *cols = mp.synth_cols;
  return NVBX_OK;
}

}  // extern "C"
