// api.cu — the extern "C" surface declared in include/pcgol_b200.h.
// Host entry points stage caller buffers through device memory (stream-ordered pool),
// run the device pipelines of voxelgrid.cu / index.cu / icp.cu and copy results back.
#include <algorithm>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "bvh.cuh"
#include "icp_math.cuh"

namespace pcg {

std::atomic<int64_t> g_launches{0};
std::atomic<int> g_profile{0};

namespace {
struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof_recs;
thread_local cudaEvent_t t_prof_start = nullptr;
thread_local const char* t_prof_name = nullptr;
}  // namespace

void prof_begin(const char* name, cudaStream_t s) {
  cudaEvent_t a;
  if (cudaEventCreate(&a) != cudaSuccess) return;
  cudaEventRecord(a, s);
  t_prof_start = a;
  t_prof_name = name;
}
void prof_end(cudaStream_t s) {
  if (!t_prof_start) return;
  cudaEvent_t b;
  if (cudaEventCreate(&b) != cudaSuccess) return;
  cudaEventRecord(b, s);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_recs.push_back(ProfRec{t_prof_name, t_prof_start, b});
  t_prof_start = nullptr;
}
static thread_local std::string t_error;
void set_error(const std::string& msg) { t_error = msg; }

void ensure_pool(int device) {
  static std::atomic<uint64_t> done{0};
  if (device < 0 || device >= 64) return;
  if (done.load(std::memory_order_acquire) & (1ull << device)) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t threshold = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
  done.fetch_or(1ull << device, std::memory_order_release);
}

// implemented in the other translation units
pcg_status voxelgrid_filter_device(const CloudView& v, const float leaf[3], const int64_t chunk[3], uint8_t* d_out,
                                   int64_t* n_out, cudaStream_t stream);
void minmax_device(const CloudView& v, float mn[3], float mx[3], cudaStream_t stream);
extern std::atomic<int> g_vg_path;
void minmax_packed_device(const CloudView& v, uint32_t index_base, long long* d_out6, cudaStream_t stream);
void voxelgrid_owner_order_device(const CloudView& v, const float leaf[3], const int64_t chunk[3], const float* mm6,
                                  const int64_t* cuts, int world, uint32_t* d_perm, int64_t* counts, uint8_t* d_send,
                                  cudaStream_t stream);
int64_t voxelgrid_chunk_histogram_device(const CloudView& v, const float leaf[3], const int64_t chunk[3],
                                         int64_t sample_step, int64_t* hist_out, int64_t cap, cudaStream_t stream,
                                         const float* mm6 = nullptr);
pcg_status voxelgrid_filter_chunks_device(const CloudView& v, const float leaf[3], const int64_t chunk[3],
                                          int64_t cid_lo, int64_t cid_hi, uint8_t* d_out, int64_t* n_out,
                                          cudaStream_t stream, const float* mm6 = nullptr);
void nearest_device(const Index& ix, const CloudView& q, float max_range, float min_dist_sq, int32_t* d_ids,
                    float* d_dist_sq, pcg_neighbor* d_aos, cudaStream_t stream);
void range_device(const Index& ix, const CloudView& q, float max_range, DevBuf<long long>& offsets,
                  DevBuf<pcg_neighbor>& out, int64_t* total_out, cudaStream_t stream);
void range_count_device(const Index& ix, const CloudView& q, float max_range, DevBuf<long long>& offsets,
                        int64_t* total_out, cudaStream_t stream);
void range_fill_device(const Index& ix, const CloudView& q, float max_range, const long long* d_offsets,
                       int64_t total, pcg_neighbor* d_out, cudaStream_t stream);
pcg_status icp_fit_device(const Index& base, const CloudView& tgt, const pcg_icp_params& prm, bool evaluate_only,
                          float trans[16], pcg_icp_stat* stat, cudaStream_t stream);
void icp_fit_pairs_device(int32_t count, const void* const* d_base, const int64_t* n_base,
                          const void* const* d_target, const int64_t* n_target, int64_t stride,
                          const int64_t xyz_off[3], const pcg_icp_params& prm, int device, float* trans_out,
                          pcg_icp_stat* stat_out, pcg_status* status_out, cudaStream_t stream);
void icp_partial_device(const Index& base, const CloudView& tgt, float max_dist, const float trans[16], bool first,
                        const uint32_t* d_order, double* d_partial16, cudaStream_t stream);
pcg_status icp_finish_host(const double partial16[16], const pcg_icp_params& prm, int32_t* iter, float trans[16],
                           pcg_evaluated* ev_out, int32_t* converged);
void icp_pairs_device(const Index& base, const CloudView& tgt, float max_dist, float min_dist_sq, int32_t* d_ids,
                      float* d_dsq, cudaStream_t stream);
struct IcpShard;
pcg_status icp_fit_multi_device(int n_dev, const Index* const* bases, const void* const* d_targets,
                                const int64_t* n_targets, int64_t stride, const int64_t xyz_off[3],
                                const pcg_icp_params& prm, float trans[16], pcg_icp_stat* stat);
IcpShard* icp_shard_new(const Index& base, const CloudView& tgt, const pcg_icp_params& prm, cudaStream_t stream);
void icp_shard_free(IcpShard* sh);
void icp_shard_partial(IcpShard& sh, double* d_partial16, cudaStream_t stream);
void icp_shard_finish(IcpShard& sh, const double* d_reduced16, cudaStream_t stream);
pcg_status icp_shard_result(IcpShard& sh, float trans[16], pcg_icp_stat* stat, int32_t* done, cudaStream_t stream);
struct CloudHeader;
struct Cloud;
void cloud_free(Cloud* c);
Cloud* cloud_unmarshal(const uint8_t* pcd, int64_t len, int device, cudaStream_t stream);
std::string cloud_marshal_header(const Cloud& c);
struct RegionGrowing;
RegionGrowing* region_growing_new_device(const Index& search, const CloudView& v, int64_t label_off,
                                         cudaStream_t stream);
void region_growing_free(RegionGrowing* rg);
int64_t region_growing_segment_device(const RegionGrowing& rg, const float p[3], float max_range,
                                      DevBuf<uint32_t>& d_result, cudaStream_t stream);
void region_growing_widen(const uint32_t* d_in, int64_t n, long long* d_out, cudaStream_t stream);
float debug_sequential_sum(const float* d_x, int64_t n, bool exact_path, cudaStream_t stream, float* stats3);

static void check_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw StatusError{PCG_E_NO_DEVICE, std::string("no usable CUDA device: ") + cudaGetErrorString(e)};
  if (device < 0 || device >= count) throw StatusError{PCG_E_INVALID_ARG, "device ordinal out of range"};
}

template <typename F>
static pcg_status guarded(F f) {
  try {
    return f();
  } catch (const StatusError& e) {
    set_error(e.msg);
    return e.s;
  } catch (const CudaError& e) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d in %s", (int)e.e, cudaGetErrorString(e.e), e.file, e.line,
             e.what);
    set_error(buf);
    cudaGetLastError();
    return (e.e == cudaErrorNoDevice || e.e == cudaErrorInsufficientDriver) ? PCG_E_NO_DEVICE : PCG_E_CUDA;
  } catch (const std::bad_alloc&) {
    set_error("host allocation failed");
    return PCG_E_TOO_LARGE;
  } catch (...) {
    set_error("unexpected exception");
    return PCG_E_CUDA;
  }
}

// Copies a host cloud to the device (whole records: other fields are needed by the filter).
struct StagedCloud {
  DevBuf<uint8_t> buf;
  CloudView view;
  StagedCloud(const void* data, int64_t n, int64_t stride, const int64_t off[3], cudaStream_t s) {
    check_view_args(data, n, stride, off);
    buf.alloc((size_t)std::max<int64_t>(1, n * stride), s);
    if (n) PCG_CUDA(cudaMemcpyAsync(buf.p, data, (size_t)(n * stride), cudaMemcpyHostToDevice, s));
    view = make_view(buf.p, n, stride, off);
  }
};

}  // namespace pcg
struct pcg_index {
  pcg::Index* ix;
};
namespace pcg {
// shared with the other translation units that export C entry points (cloud.cu)
pcg_status api_guard(const std::function<pcg_status()>& f) { return guarded(f); }
pcg_index* api_wrap_index(Index* ix) { return new pcg_index{ix}; }
Index* api_index_of(pcg_index* idx) { return idx->ix; }
void api_check_device(int device) { check_device(device); }

struct RangeResult {
  std::vector<int64_t> offsets;
  // plain heap memory: page-locking a result of hundreds of MB costs far more than the copy it would speed up
  pcg_neighbor* neighbors = nullptr;
  int64_t total = 0;
  ~RangeResult() { free(neighbors); }
};

}  // namespace pcg

using namespace pcg;

struct pcg_range_result {
  RangeResult r;
};
struct pcg_icp_shard {
  IcpShard* sh;
  int device;
};
struct pcg_region_growing {
  RegionGrowing* rg;
  int device;
};

extern "C" {

int32_t pcg_abi_version(void) { return PCGOL_B200_ABI_VERSION; }
const char* pcg_last_error(void) { return t_error.c_str(); }
const char* pcg_status_string(pcg_status s) {
  switch (s) {
    case PCG_OK: return "ok";
    case PCG_E_INVALID_ARG: return "invalid argument";
    case PCG_E_NO_POINT: return "no point";
    case PCG_E_REF_WOULD_PANIC: return "reference would panic (index out of range)";
    case PCG_E_REF_UNDEFINED: return "reference behaviour is implementation-specific for this input";
    case PCG_E_NOT_ENOUGH_PAIRS: return "not enough correspondence pairs";
    case PCG_E_CUDA: return "CUDA error";
    case PCG_E_NO_DEVICE: return "no usable CUDA device";
    case PCG_E_TOO_LARGE: return "problem too large";
    case PCG_E_PCD_SYNTAX: return "PCD syntax error";
    case PCG_E_PCD_EOF: return "PCD: unexpected end of data";
    case PCG_E_PCD_CORRUPT: return "PCD: corrupt compressed data";
    case PCG_E_INVALID_FIELD: return "invalid field name";
  }
  return "unknown status";
}
int32_t pcg_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return count;
}
int64_t pcg_kernel_launch_count(void) { return g_launches.load(); }

void pcg_profile_enable(int32_t on) { g_profile.store(on ? 1 : 0); }
void pcg_debug_set_vg_path(int32_t path) { g_vg_path.store(path); }

pcg_status pcg_debug_sequential_sum_f32(const float* x, int64_t n, int32_t device, int32_t exact_path, float* out) {
  return guarded([&]() -> pcg_status {
    if (!out || n < 0 || (n && !x)) throw StatusError{PCG_E_INVALID_ARG, "bad arguments"};
    check_device(device);
    DeviceGuard g(device);
    cudaStream_t s = cudaStreamPerThread;
    const int64_t n_pad = (n + 3) & ~(int64_t)3;
    DevBuf<float> d((size_t)std::max<int64_t>(4, n_pad), s);
    PCG_CUDA(cudaMemsetAsync(d.p, 0, d.bytes(), s));
    if (n) PCG_CUDA(cudaMemcpyAsync(d.p, x, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s));
    float stats[3] = {0, 0, 0};
    *out = debug_sequential_sum(d.p, n, exact_path != 0, s, stats);
    out[1] = stats[0];  // out is float[4]: sum, chunks on the integer path, chunks replayed, cycles of the walk
    out[2] = stats[1];
    out[3] = stats[2];
    return PCG_OK;
  });
}


// Synchronises the device(s), folds the recorded events into per-kernel totals and writes
// a JSON object {"kernel": {"launches": n, "total_ms": t}, ...}; clears the records.
int64_t pcg_profile_report(char* buf, int64_t cap) {
  std::vector<ProfRec> recs;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    recs.swap(g_prof_recs);
  }
  std::map<std::string, std::pair<int64_t, double>> agg;
  for (auto& r : recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.name];
      e.first++;
      e.second += ms;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  std::string out = "{";
  bool first = true;
  for (auto& kv : agg) {
    char line[512];
    std::string name = kv.first;
    for (auto& c : name)
      if (c == '"' || c == '\\') c = '_';
    snprintf(line, sizeof(line), "%s\"%s\": {\"launches\": %lld, \"total_ms\": %.6f}", first ? "" : ", ", name.c_str(),
             (long long)kv.second.first, kv.second.second);
    out += line;
    first = false;
  }
  out += "}";
  if (buf && cap > 0) {
    size_t ncopy = std::min<size_t>((size_t)cap - 1, out.size());
    memcpy(buf, out.data(), ncopy);
    buf[ncopy] = 0;
  }
  return (int64_t)out.size();
}

pcg_status pcg_host_alloc(void** out, int64_t bytes) {
  return guarded([&]() -> pcg_status {
    if (!out || bytes < 0) throw StatusError{PCG_E_INVALID_ARG, "pcg_host_alloc: bad arguments"};
    PCG_CUDA(cudaMallocHost(out, (size_t)std::max<int64_t>(1, bytes)));
    return PCG_OK;
  });
}
void pcg_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
pcg_status pcg_device_alloc(int32_t device, void** out, int64_t bytes) {
  return guarded([&]() -> pcg_status {
    if (!out || bytes < 0) throw StatusError{PCG_E_INVALID_ARG, "pcg_device_alloc: bad arguments"};
    check_device(device);
    DeviceGuard g(device);
    PCG_CUDA(cudaMalloc(out, (size_t)std::max<int64_t>(1, bytes)));
    return PCG_OK;
  });
}
void pcg_device_free(int32_t device, void* p) {
  if (!p) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  cudaFree(p);
  if (prev >= 0) cudaSetDevice(prev);
}
pcg_status pcg_memcpy_h2d(int32_t device, void* dst, const void* src, int64_t bytes) {
  return guarded([&]() -> pcg_status {
    check_device(device);
    DeviceGuard g(device);
    PCG_CUDA(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyHostToDevice));
    return PCG_OK;
  });
}
pcg_status pcg_memcpy_d2h(int32_t device, void* dst, const void* src, int64_t bytes) {
  return guarded([&]() -> pcg_status {
    check_device(device);
    DeviceGuard g(device);
    PCG_CUDA(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return PCG_OK;
  });
}
pcg_status pcg_device_synchronize(int32_t device) {
  return guarded([&]() -> pcg_status {
    check_device(device);
    DeviceGuard g(device);
    PCG_CUDA(cudaDeviceSynchronize());
    return PCG_OK;
  });
}

// ---- index -----------------------------------------------------------------------------
pcg_status pcg_index_build_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                               int32_t device, void* stream, pcg_index** out) {
  return guarded([&]() -> pcg_status {
    if (!out) throw StatusError{PCG_E_INVALID_ARG, "pcg_index_build: null output"};
    *out = nullptr;
    check_device(device);
    check_view_args(d_data, n, stride, xyz_off);
    DeviceGuard g(device);
    CloudView v = make_view(d_data, n, stride, xyz_off);
    Index* ix = index_build_device(v, device, (cudaStream_t)stream);
    *out = new pcg_index{ix};
    return PCG_OK;
  });
}

pcg_status pcg_index_build(const void* data, int64_t n, int64_t stride, const int64_t xyz_off[3], int32_t device,
                           pcg_index** out) {
  return guarded([&]() -> pcg_status {
    if (!out) throw StatusError{PCG_E_INVALID_ARG, "pcg_index_build: null output"};
    *out = nullptr;
    check_device(device);
    DeviceGuard g(device);
    cudaStream_t s = cudaStreamPerThread;
    StagedCloud c(data, n, stride, xyz_off, s);
    Index* ix = index_build_device(c.view, device, s);
    PCG_CUDA(cudaStreamSynchronize(s));
    *out = new pcg_index{ix};
    return PCG_OK;
  });
}

void pcg_index_free(pcg_index* idx) {
  if (!idx) return;
  index_free(idx->ix);
  delete idx;
}
int64_t pcg_index_len(const pcg_index* idx) { return idx ? idx->ix->n : 0; }
int32_t pcg_index_device(const pcg_index* idx) { return idx ? idx->ix->device : -1; }
int64_t pcg_index_device_bytes(const pcg_index* idx) { return idx ? idx->ix->bytes : 0; }

pcg_status pcg_debug_index_slots(const pcg_index* idx, float* out, int64_t cap_slots, int64_t* slots) {
  return guarded([&]() -> pcg_status {
    if (!idx || !slots) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    const int64_t total = (int64_t)idx->ix->leaves * kLeaf;
    *slots = idx->ix->n ? total : 0;
    if (!out || idx->ix->n == 0) return PCG_OK;
    if (cap_slots < total) throw StatusError{PCG_E_INVALID_ARG, "buffer too small"};
    DeviceGuard g(idx->ix->device);
    // device layout: one line per leaf, x[8] y[8] z[8] id[8] (bvh.cuh); the hook reports slots as {x, y, z, id}
    std::vector<float> raw((size_t)total * 4);
    PCG_CUDA(cudaMemcpy(raw.data(), idx->ix->pts, raw.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int64_t s = 0; s < total; s++)
      for (int c = 0; c < 4; c++) out[4 * s + c] = raw[(size_t)(s / kLeaf) * (4 * kLeaf) + (size_t)c * kLeaf + (size_t)(s % kLeaf)];
    return PCG_OK;
  });
}

pcg_status pcg_index_nearest_approx_dev(pcg_index* idx, const void* d_q, int64_t nq, int64_t q_stride,
                                        const int64_t q_xyz_off[3], float max_range, float min_dist_sq,
                                        int32_t* d_ids, float* d_dist_sq, void* stream) {
  return guarded([&]() -> pcg_status {
    if (!idx) throw StatusError{PCG_E_INVALID_ARG, "null index"};
    check_view_args(d_q, nq, q_stride, q_xyz_off);
    if (nq && (!d_ids || !d_dist_sq)) throw StatusError{PCG_E_INVALID_ARG, "null output"};
    if (!(min_dist_sq >= 0.f)) throw StatusError{PCG_E_INVALID_ARG, "MinDistSq must be >= 0"};
    DeviceGuard g(idx->ix->device);
    nearest_device(*idx->ix, make_view(d_q, nq, q_stride, q_xyz_off), max_range, min_dist_sq, d_ids, d_dist_sq,
                   nullptr, (cudaStream_t)stream);
    return PCG_OK;
  });
}

pcg_status pcg_index_nearest_dev(pcg_index* idx, const void* d_q, int64_t nq, int64_t q_stride,
                                 const int64_t q_xyz_off[3], float max_range, int32_t* d_ids, float* d_dist_sq,
                                 void* stream) {
  return pcg_index_nearest_approx_dev(idx, d_q, nq, q_stride, q_xyz_off, max_range, 0.f, d_ids, d_dist_sq, stream);
}

// Large host batches are cut into slices that travel on a few streams: the upload of one slice, the
// search of another and the download of a third overlap (two DMA engines + the SMs), so the call is bound
// by the slower of PCIe and the kernel instead of their sum.  (Pageable caller memory still works; the
// copies then serialise inside the driver.)
constexpr int64_t kNearestSlice = 1 << 20;
constexpr int kNearestStreams = 4;
struct SliceStreams {
  cudaStream_t s[kNearestStreams] = {nullptr, nullptr, nullptr, nullptr};
  int device = -1;
  ~SliceStreams() {
    // the runtime may already be shutting down when thread-local storage is torn down: leak on purpose
  }
  void ensure(int dev) {
    if (device == dev) return;
    for (int i = 0; i < kNearestStreams; i++) {
      if (s[i]) cudaStreamDestroy(s[i]);
      PCG_CUDA(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
    }
    device = dev;
  }
};

pcg_status pcg_index_nearest_approx(pcg_index* idx, const void* q, int64_t nq, int64_t q_stride,
                                    const int64_t q_xyz_off[3], float max_range, float min_dist_sq,
                                    pcg_neighbor* out) {
  return guarded([&]() -> pcg_status {
    if (!idx) throw StatusError{PCG_E_INVALID_ARG, "null index"};
    if (nq && !out) throw StatusError{PCG_E_INVALID_ARG, "null output"};
    if (!(min_dist_sq >= 0.f)) throw StatusError{PCG_E_INVALID_ARG, "MinDistSq must be >= 0"};
    check_view_args(q, nq, q_stride, q_xyz_off);
    DeviceGuard g(idx->ix->device);
    if (nq == 0) return PCG_OK;
    if (nq < 2 * kNearestSlice) {
      cudaStream_t s = cudaStreamPerThread;
      StagedCloud c(q, nq, q_stride, q_xyz_off, s);
      DevBuf<pcg_neighbor> d_out((size_t)nq, s);
      nearest_device(*idx->ix, c.view, max_range, min_dist_sq, nullptr, nullptr, d_out.p, s);
      PCG_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)nq * sizeof(pcg_neighbor), cudaMemcpyDeviceToHost, s));
      PCG_CUDA(cudaStreamSynchronize(s));
      return PCG_OK;
    }
    static thread_local SliceStreams streams;
    streams.ensure(idx->ix->device);
    const uint8_t* qb = (const uint8_t*)q;
    try {
      int k = 0;
      for (int64_t b = 0; b < nq; b += kNearestSlice, k++) {
        const int64_t m = std::min(kNearestSlice, nq - b);
        cudaStream_t s = streams.s[k % kNearestStreams];
        StagedCloud c(qb + b * q_stride, m, q_stride, q_xyz_off, s);  // stream-ordered: freed after the slice's work
        DevBuf<pcg_neighbor> d_out((size_t)m, s);
        nearest_device(*idx->ix, c.view, max_range, min_dist_sq, nullptr, nullptr, d_out.p, s);
        PCG_CUDA(cudaMemcpyAsync(out + b, d_out.p, (size_t)m * sizeof(pcg_neighbor), cudaMemcpyDeviceToHost, s));
      }
    } catch (...) {  // nothing may still be writing into the caller's buffer when the error is reported
      for (int i = 0; i < kNearestStreams; i++) cudaStreamSynchronize(streams.s[i]);
      throw;
    }
    for (int i = 0; i < kNearestStreams; i++) PCG_CUDA(cudaStreamSynchronize(streams.s[i]));
    return PCG_OK;
  });
}

pcg_status pcg_index_nearest(pcg_index* idx, const void* q, int64_t nq, int64_t q_stride,
                             const int64_t q_xyz_off[3], float max_range, pcg_neighbor* out) {
  return pcg_index_nearest_approx(idx, q, nq, q_stride, q_xyz_off, max_range, 0.f, out);
}

pcg_status pcg_index_delete_points(pcg_index* idx, const int64_t* ids, int64_t n) {
  return guarded([&]() -> pcg_status {
    if (!idx || n < 0 || (n && !ids)) throw StatusError{PCG_E_INVALID_ARG, "bad arguments"};
    const int64_t len = idx->ix->n;
    for (int64_t i = 0; i < n; i++) {
      if (ids[i] < 0 || ids[i] > len - 1) {  // kdtree.go:323-325
        char buf[96];
        snprintf(buf, sizeof(buf), "%lld does not correspond to any point in the tree", (long long)ids[i]);
        throw StatusError{PCG_E_INVALID_ARG, buf};
      }
    }
    if (n == 0) return PCG_OK;
    DeviceGuard g(idx->ix->device);
    cudaStream_t s = cudaStreamPerThread;
    DevBuf<int64_t> d_ids((size_t)n, s);
    PCG_CUDA(cudaMemcpyAsync(d_ids.p, ids, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    index_delete_points_device(*idx->ix, d_ids.p, n, s);
    PCG_CUDA(cudaStreamSynchronize(s));
    return PCG_OK;
  });
}

pcg_status pcg_index_range(pcg_index* idx, const void* q, int64_t nq, int64_t q_stride,
                           const int64_t q_xyz_off[3], float max_range, pcg_range_result** out) {
  return guarded([&]() -> pcg_status {
    if (!idx || !out) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    *out = nullptr;
    DeviceGuard g(idx->ix->device);
    cudaStream_t s = cudaStreamPerThread;
    StagedCloud c(q, nq, q_stride, q_xyz_off, s);
    DevBuf<long long> d_off;
    DevBuf<pcg_neighbor> d_nb;
    int64_t total = 0;
    range_device(*idx->ix, c.view, max_range, d_off, d_nb, &total, s);
    std::unique_ptr<pcg_range_result> r(new pcg_range_result());
    r->r.offsets.resize((size_t)nq + 1);
    r->r.total = total;
    static_assert(sizeof(long long) == sizeof(int64_t), "offset width");
    PCG_CUDA(cudaMemcpyAsync(r->r.offsets.data(), d_off.p, ((size_t)nq + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost,
                             s));
    if (total) {
      r->r.neighbors = (pcg_neighbor*)malloc((size_t)total * sizeof(pcg_neighbor));
      if (!r->r.neighbors) throw std::bad_alloc();
      PCG_CUDA(cudaMemcpyAsync(r->r.neighbors, d_nb.p, (size_t)total * sizeof(pcg_neighbor), cudaMemcpyDeviceToHost,
                               s));
    }
    PCG_CUDA(cudaStreamSynchronize(s));
    *out = r.release();
    return PCG_OK;
  });
}
// Two-call protocol with caller-owned (reusable, possibly pinned) buffers.
pcg_status pcg_index_range_count(pcg_index* idx, const void* q, int64_t nq, int64_t q_stride,
                                 const int64_t q_xyz_off[3], float max_range, int64_t* offsets) {
  return guarded([&]() -> pcg_status {
    if (!idx || !offsets) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    DeviceGuard g(idx->ix->device);
    cudaStream_t s = cudaStreamPerThread;
    StagedCloud c(q, nq, q_stride, q_xyz_off, s);
    DevBuf<long long> d_off;
    int64_t total = 0;
    range_count_device(*idx->ix, c.view, max_range, d_off, &total, s);
    PCG_CUDA(cudaMemcpyAsync(offsets, d_off.p, ((size_t)nq + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    PCG_CUDA(cudaStreamSynchronize(s));
    return PCG_OK;
  });
}

pcg_status pcg_index_range_fill(pcg_index* idx, const void* q, int64_t nq, int64_t q_stride,
                                const int64_t q_xyz_off[3], float max_range, const int64_t* offsets,
                                pcg_neighbor* out) {
  return guarded([&]() -> pcg_status {
    if (!idx || !offsets) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    if (nq < 0) throw StatusError{PCG_E_INVALID_ARG, "negative query count"};
    // the CSR offsets come from the caller: they must be the monotone table pcg_index_range_count produced
    if (offsets[0] != 0) throw StatusError{PCG_E_INVALID_ARG, "offsets[0] must be 0"};
    for (int64_t i = 0; i < nq; i++)
      if (offsets[i + 1] < offsets[i]) throw StatusError{PCG_E_INVALID_ARG, "offsets are not monotone"};
    const int64_t total = offsets[nq];
    if (total && !out) throw StatusError{PCG_E_INVALID_ARG, "null output"};
    if (nq == 0 || total == 0) return PCG_OK;
    DeviceGuard g(idx->ix->device);
    cudaStream_t s = cudaStreamPerThread;
    StagedCloud c(q, nq, q_stride, q_xyz_off, s);
    DevBuf<long long> d_off((size_t)nq + 1, s);
    PCG_CUDA(cudaMemcpyAsync(d_off.p, offsets, ((size_t)nq + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    DevBuf<pcg_neighbor> d_out((size_t)total, s);
    range_fill_device(*idx->ix, c.view, max_range, d_off.p, total, d_out.p, s);
    PCG_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)total * sizeof(pcg_neighbor), cudaMemcpyDeviceToHost, s));
    PCG_CUDA(cudaStreamSynchronize(s));
    return PCG_OK;
  });
}

int64_t pcg_range_total(const pcg_range_result* r) { return r ? r->r.total : 0; }
const int64_t* pcg_range_offsets(const pcg_range_result* r) { return r ? r->r.offsets.data() : nullptr; }
const pcg_neighbor* pcg_range_neighbors(const pcg_range_result* r) { return r ? r->r.neighbors : nullptr; }
void pcg_range_free(pcg_range_result* r) { delete r; }

// ---- region growing ---------------------------------------------------------------------
pcg_status pcg_region_growing_new(pcg_index* search, const void* data, int64_t n, int64_t stride,
                                  const int64_t xyz_off[3], int64_t label_off, pcg_region_growing** out) {
  return guarded([&]() -> pcg_status {
    if (!search || !out) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    *out = nullptr;
    check_view_args(data, n, stride, xyz_off);
    if (label_off < 0 || label_off + 4 > stride) throw StatusError{PCG_E_INVALID_ARG, "label offset outside the record"};
    if (n != search->ix->n)
      throw StatusError{PCG_E_INVALID_ARG, "the property accessor and the search must cover the same points"};
    DeviceGuard g(search->ix->device);
    cudaStream_t s = cudaStreamPerThread;
    StagedCloud c(data, n, stride, xyz_off, s);
    RegionGrowing* rg = region_growing_new_device(*search->ix, c.view, label_off, s);
    PCG_CUDA(cudaStreamSynchronize(s));
    *out = new pcg_region_growing{rg, search->ix->device};
    return PCG_OK;
  });
}

void pcg_region_growing_free(pcg_region_growing* rg) {
  if (!rg) return;
  region_growing_free(rg->rg);
  delete rg;
}

pcg_status pcg_region_growing_segment(pcg_region_growing* rg, const float p[3], float max_range, int64_t* indice,
                                      int64_t cap, int64_t* n_out) {
  return guarded([&]() -> pcg_status {
    if (!rg || !p || !n_out || cap < 0 || (cap && !indice)) throw StatusError{PCG_E_INVALID_ARG, "bad arguments"};
    *n_out = 0;
    DeviceGuard g(rg->device);
    cudaStream_t s = cudaStreamPerThread;
    DevBuf<uint32_t> d_result;
    const int64_t m = region_growing_segment_device(*rg->rg, p, max_range, d_result, s);
    *n_out = m;
    if (m > cap) throw StatusError{PCG_E_INVALID_ARG, "result buffer too small (n_out holds the size needed)"};
    if (m == 0) return PCG_OK;
    DevBuf<long long> wide((size_t)m, s);
    region_growing_widen(d_result.p, m, wide.p, s);
    PCG_CUDA(cudaMemcpyAsync(indice, wide.p, (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    PCG_CUDA(cudaStreamSynchronize(s));
    return PCG_OK;
  });
}

// ---- voxel grid -------------------------------------------------------------------------
static void check_vg_args(const float leaf[3], const int64_t chunk[3]) {
  if (!leaf || !chunk) throw StatusError{PCG_E_INVALID_ARG, "null leaf/chunk"};
  for (int k = 0; k < 3; k++)
    if (chunk[k] < 0) throw StatusError{PCG_E_INVALID_ARG, "negative chunk size"};
}

pcg_status pcg_voxelgrid_filter_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                                    const float leaf[3], const int64_t chunk[3], int32_t device, void* d_out,
                                    int64_t* n_out, void* stream) {
  return guarded([&]() -> pcg_status {
    if (!n_out) throw StatusError{PCG_E_INVALID_ARG, "null n_out"};
    *n_out = 0;
    check_device(device);
    check_view_args(d_data, n, stride, xyz_off);
    check_vg_args(leaf, chunk);
    if (n && !d_out) throw StatusError{PCG_E_INVALID_ARG, "null output"};
    DeviceGuard g(device);
    return voxelgrid_filter_device(make_view(d_data, n, stride, xyz_off), leaf, chunk, (uint8_t*)d_out, n_out,
                                   (cudaStream_t)stream);
  });
}

pcg_status pcg_voxelgrid_filter(const void* data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                                const float leaf[3], const int64_t chunk[3], int32_t device, void* out,
                                int64_t* n_out) {
  return guarded([&]() -> pcg_status {
    if (!n_out) throw StatusError{PCG_E_INVALID_ARG, "null n_out"};
    *n_out = 0;
    check_device(device);
    check_vg_args(leaf, chunk);
    if (n && !out) throw StatusError{PCG_E_INVALID_ARG, "null output"};
    DeviceGuard g(device);
    cudaStream_t s = cudaStreamPerThread;
    StagedCloud c(data, n, stride, xyz_off, s);
    DevBuf<uint8_t> d_out((size_t)std::max<int64_t>(1, n * stride), s);
    pcg_status rc = voxelgrid_filter_device(c.view, leaf, chunk, d_out.p, n_out, s);
    if (rc == PCG_OK && *n_out) {
      PCG_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)(*n_out * stride), cudaMemcpyDeviceToHost, s));
      PCG_CUDA(cudaStreamSynchronize(s));
    }
    return rc;
  });
}

pcg_status pcg_voxelgrid_chunk_histogram_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                                             const float leaf[3], const int64_t chunk[3], int32_t device,
                                             int64_t sample_step, int64_t* hist, int64_t cap, int64_t* n_chunks,
                                             void* stream) {
  return guarded([&]() -> pcg_status {
    if (!n_chunks) throw StatusError{PCG_E_INVALID_ARG, "null n_chunks"};
    *n_chunks = 0;
    check_device(device);
    check_view_args(d_data, n, stride, xyz_off);
    check_vg_args(leaf, chunk);
    DeviceGuard g(device);
    *n_chunks = voxelgrid_chunk_histogram_device(make_view(d_data, n, stride, xyz_off), leaf, chunk, sample_step, hist,
                                                 cap,
                                                 (cudaStream_t)stream);
    return PCG_OK;
  });
}

pcg_status pcg_voxelgrid_filter_chunks_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                                           const float leaf[3], const int64_t chunk[3], int64_t cid_lo, int64_t cid_hi,
                                           int32_t device, void* d_out, int64_t* n_out, void* stream) {
  return guarded([&]() -> pcg_status {
    if (!n_out) throw StatusError{PCG_E_INVALID_ARG, "null n_out"};
    *n_out = 0;
    check_device(device);
    check_view_args(d_data, n, stride, xyz_off);
    check_vg_args(leaf, chunk);
    if (n && !d_out) throw StatusError{PCG_E_INVALID_ARG, "null output"};
    DeviceGuard g(device);
    return voxelgrid_filter_chunks_device(make_view(d_data, n, stride, xyz_off), leaf, chunk, cid_lo, cid_hi,
                                          (uint8_t*)d_out, n_out, (cudaStream_t)stream);
  });
}

// ---- point-sharded Filter: the same steps with the bounds of the WHOLE cloud supplied by the caller ----
pcg_status pcg_minmax_packed_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                                 int32_t device, int64_t index_base, int64_t* d_out6, void* stream) {
  return guarded([&]() -> pcg_status {
    check_device(device);
    check_view_args(d_data, n, stride, xyz_off);
    if (!d_out6 || index_base < 0 || index_base + n >= ((int64_t)1 << 30))
      throw StatusError{PCG_E_INVALID_ARG, "null output / global point index out of range"};
    DeviceGuard g(device);
    minmax_packed_device(make_view(d_data, n, stride, xyz_off), (uint32_t)index_base, (long long*)d_out6,
                         (cudaStream_t)stream);
    return PCG_OK;
  });
}

pcg_status pcg_voxelgrid_chunk_histogram_mm_dev(const void* d_data, int64_t n, int64_t stride,
                                                const int64_t xyz_off[3], const float leaf[3], const int64_t chunk[3],
                                                const float mm6[6], int32_t device, int64_t sample_step, int64_t* hist,
                                                int64_t cap, int64_t* n_chunks, void* stream) {
  return guarded([&]() -> pcg_status {
    if (!n_chunks || !mm6) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    *n_chunks = 0;
    check_device(device);
    check_view_args(d_data, n, stride, xyz_off);
    check_vg_args(leaf, chunk);
    DeviceGuard g(device);
    *n_chunks = voxelgrid_chunk_histogram_device(make_view(d_data, n, stride, xyz_off), leaf, chunk, sample_step, hist,
                                                 cap, (cudaStream_t)stream, mm6);
    return PCG_OK;
  });
}

pcg_status pcg_voxelgrid_owner_order_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                                         const float leaf[3], const int64_t chunk[3], const float mm6[6],
                                         const int64_t* cuts, int32_t world, int32_t device, uint32_t* d_perm,
                                         int64_t* counts, void* d_send, void* stream) {
  return guarded([&]() -> pcg_status {
    if (!mm6 || !cuts || !counts || (n && !d_perm)) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    check_device(device);
    check_view_args(d_data, n, stride, xyz_off);
    check_vg_args(leaf, chunk);
    DeviceGuard g(device);
    voxelgrid_owner_order_device(make_view(d_data, n, stride, xyz_off), leaf, chunk, mm6, cuts, world, d_perm, counts,
                                 (uint8_t*)d_send, (cudaStream_t)stream);
    return PCG_OK;
  });
}

pcg_status pcg_voxelgrid_filter_chunks_mm_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                                              const float leaf[3], const int64_t chunk[3], const float mm6[6],
                                              int64_t cid_lo, int64_t cid_hi, int32_t device, void* d_out,
                                              int64_t* n_out, void* stream) {
  return guarded([&]() -> pcg_status {
    if (!n_out || !mm6) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    *n_out = 0;
    check_device(device);
    check_view_args(d_data, n, stride, xyz_off);
    check_vg_args(leaf, chunk);
    if (n && !d_out) throw StatusError{PCG_E_INVALID_ARG, "null output"};
    DeviceGuard g(device);
    return voxelgrid_filter_chunks_device(make_view(d_data, n, stride, xyz_off), leaf, chunk, cid_lo, cid_hi,
                                          (uint8_t*)d_out, n_out, (cudaStream_t)stream, mm6);
  });
}

pcg_status pcg_minmax_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3], int32_t device,
                          float mn[3], float mx[3], void* stream) {
  return guarded([&]() -> pcg_status {
    check_device(device);
    check_view_args(d_data, n, stride, xyz_off);
    if (n == 0) throw StatusError{PCG_E_NO_POINT, "no point"};
    DeviceGuard g(device);
    minmax_device(make_view(d_data, n, stride, xyz_off), mn, mx, (cudaStream_t)stream);
    return PCG_OK;
  });
}

// ---- icp ----------------------------------------------------------------------------------
pcg_status pcg_icp_pairs(pcg_index* base, const void* target, int64_t n, int64_t stride, const int64_t xyz_off[3],
                         float max_dist, int64_t* base_id, int64_t* target_id, float* dist_sq, int64_t* n_pairs) {
  return pcg_icp_pairs_approx(base, target, n, stride, xyz_off, max_dist, 0.f, base_id, target_id, dist_sq, n_pairs);
}

pcg_status pcg_icp_pairs_approx(pcg_index* base, const void* target, int64_t n, int64_t stride,
                                const int64_t xyz_off[3], float max_dist, float min_dist_sq, int64_t* base_id,
                                int64_t* target_id, float* dist_sq, int64_t* n_pairs) {
  return guarded([&]() -> pcg_status {
    if (!base || !n_pairs) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    if (n > 0 && (!base_id || !target_id || !dist_sq)) throw StatusError{PCG_E_INVALID_ARG, "null output"};
    if (!(min_dist_sq >= 0.f)) throw StatusError{PCG_E_INVALID_ARG, "MinDistSq must be >= 0"};
    *n_pairs = 0;
    DeviceGuard g(base->ix->device);
    cudaStream_t s = cudaStreamPerThread;
    StagedCloud c(target, n, stride, xyz_off, s);
    if (n == 0) return PCG_OK;
    DevBuf<int32_t> d_ids((size_t)n, s);
    DevBuf<float> d_dsq((size_t)n, s);
    icp_pairs_device(*base->ix, c.view, max_dist, min_dist_sq, d_ids.p, d_dsq.p, s);
    std::vector<int32_t> ids((size_t)n);
    std::vector<float> dsq((size_t)n);
    PCG_CUDA(cudaMemcpyAsync(ids.data(), d_ids.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    PCG_CUDA(cudaMemcpyAsync(dsq.data(), d_dsq.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    PCG_CUDA(cudaStreamSynchronize(s));
    int64_t m = 0;  // correspondence.go:27-35: unmatched targets are skipped, order kept
    for (int64_t i = 0; i < n; i++) {
      if (ids[(size_t)i] < 0) continue;
      base_id[m] = ids[(size_t)i];
      target_id[m] = i;
      dist_sq[m] = dsq[(size_t)i];
      m++;
    }
    *n_pairs = m;
    return PCG_OK;
  });
}

pcg_status pcg_icp_evaluate(pcg_index* base, const void* target, int64_t n, int64_t stride,
                            const int64_t xyz_off[3], float max_dist, int32_t min_pairs, int32_t mode,
                            pcg_evaluated* out, int64_t* n_pairs) {
  pcg_icp_params prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.max_dist = max_dist;
  prm.min_pairs = min_pairs;
  prm.mode = mode;
  return pcg_icp_evaluate_params(base, target, n, stride, xyz_off, &prm, out, n_pairs);
}

pcg_status pcg_icp_evaluate_params(pcg_index* base, const void* target, int64_t n, int64_t stride,
                                   const int64_t xyz_off[3], const pcg_icp_params* params, pcg_evaluated* out,
                                   int64_t* n_pairs) {
  return guarded([&]() -> pcg_status {
    if (!base || !out || !params) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    DeviceGuard g(base->ix->device);
    cudaStream_t s = cudaStreamPerThread;
    StagedCloud c(target, n, stride, xyz_off, s);
    const pcg_icp_params& prm = *params;
    pcg_icp_stat stat;
    pcg_status rc = icp_fit_device(*base->ix, c.view, prm, true, nullptr, &stat, s);
    *out = stat.evaluated;
    if (n_pairs) *n_pairs = stat.n_pairs;
    if (rc == PCG_E_NOT_ENOUGH_PAIRS) set_error("not enough correspondence pairs");
    return rc;
  });
}

pcg_status pcg_icp_fit_dev(pcg_index* base, const void* d_target, int64_t n, int64_t stride,
                           const int64_t xyz_off[3], const pcg_icp_params* params, float trans[16],
                           pcg_icp_stat* stat, void* stream) {
  return guarded([&]() -> pcg_status {
    if (!base || !params || !trans) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    check_view_args(d_target, n, stride, xyz_off);
    DeviceGuard g(base->ix->device);
    pcg_status rc = icp_fit_device(*base->ix, make_view(d_target, n, stride, xyz_off), *params, false, trans, stat,
                                   (cudaStream_t)stream);
    if (rc == PCG_E_NOT_ENOUGH_PAIRS) set_error("not enough correspondence pairs");
    return rc;
  });
}

pcg_status pcg_icp_fit(pcg_index* base, const void* target, int64_t n, int64_t stride, const int64_t xyz_off[3],
                       const pcg_icp_params* params, float trans[16], pcg_icp_stat* stat) {
  return guarded([&]() -> pcg_status {
    if (!base || !params || !trans) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    DeviceGuard g(base->ix->device);
    cudaStream_t s = cudaStreamPerThread;
    StagedCloud c(target, n, stride, xyz_off, s);
    pcg_status rc = icp_fit_device(*base->ix, c.view, *params, false, trans, stat, s);
    if (rc == PCG_E_NOT_ENOUGH_PAIRS) set_error("not enough correspondence pairs");
    return rc;
  });
}

pcg_status pcg_icp_fit_pairs_dev(int32_t count, const void* const* d_base, const int64_t* n_base,
                                 const void* const* d_target, const int64_t* n_target, int64_t stride,
                                 const int64_t xyz_off[3], const pcg_icp_params* params, int32_t device,
                                 float* trans_out, pcg_icp_stat* stat_out, pcg_status* status_out, void* stream) {
  return guarded([&]() -> pcg_status {
    if (count < 0 || (count && (!d_base || !n_base || !d_target || !n_target || !params)))
      throw StatusError{PCG_E_INVALID_ARG, "bad arguments"};
    check_device(device);
    DeviceGuard g(device);
    icp_fit_pairs_device(count, d_base, n_base, d_target, n_target, stride, xyz_off, *params, device, trans_out,
                         stat_out, status_out, (cudaStream_t)stream);
    return PCG_OK;
  });
}

pcg_status pcg_query_order_dev(pcg_index* idx, const void* d_q, int64_t n, int64_t stride, const int64_t xyz_off[3],
                               uint32_t* d_order, void* stream) {
  return guarded([&]() -> pcg_status {
    if (!idx || (n && !d_order)) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    check_view_args(d_q, n, stride, xyz_off);
    if (idx->ix->n == 0) throw StatusError{PCG_E_INVALID_ARG, "empty index has no bounding box to order by"};
    DeviceGuard g(idx->ix->device);
    query_order_device(*idx->ix, make_view(d_q, n, stride, xyz_off), d_order, (cudaStream_t)stream);
    return PCG_OK;
  });
}

pcg_status pcg_icp_partial_dev(pcg_index* base, const void* d_target, int64_t n, int64_t stride,
                               const int64_t xyz_off[3], float max_dist, const float trans[16], int32_t first,
                               const uint32_t* d_visit_order, double* d_partial16, void* stream) {
  return guarded([&]() -> pcg_status {
    if (!base || !trans || !d_partial16) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    check_view_args(d_target, n, stride, xyz_off);
    DeviceGuard g(base->ix->device);
    icp_partial_device(*base->ix, make_view(d_target, n, stride, xyz_off), max_dist, trans, first != 0, d_visit_order,
                       d_partial16, (cudaStream_t)stream);
    return PCG_OK;
  });
}

// ---- multi-GPU entry points for a single-process caller (device list = the devices of the index replicas) ----
pcg_status pcg_index_replicate(pcg_index* src, int32_t device, pcg_index** out) {
  return guarded([&]() -> pcg_status {
    if (!src || !out) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    *out = nullptr;
    check_device(device);
    *out = new pcg_index{index_replicate(*src->ix, device)};
    return PCG_OK;
  });
}

pcg_status pcg_icp_fit_multi_dev(int32_t n_dev, pcg_index* const* bases, const void* const* d_targets,
                                 const int64_t* n_targets, int64_t stride, const int64_t xyz_off[3],
                                 const pcg_icp_params* params, float trans[16], pcg_icp_stat* stat) {
  return guarded([&]() -> pcg_status {
    if (!bases || !d_targets || !n_targets || !params || !trans) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    if (n_dev < 1 || n_dev > 8) throw StatusError{PCG_E_INVALID_ARG, "1 to 8 devices"};
    const Index* ix[8];
    for (int r = 0; r < n_dev; r++) {
      if (!bases[r]) throw StatusError{PCG_E_INVALID_ARG, "null index"};
      ix[r] = bases[r]->ix;
    }
    return icp_fit_multi_device(n_dev, ix, d_targets, n_targets, stride, xyz_off, *params, trans, stat);
  });
}

pcg_status pcg_icp_fit_multi(int32_t n_dev, pcg_index* const* bases, const void* target, int64_t n, int64_t stride,
                             const int64_t xyz_off[3], const pcg_icp_params* params, float trans[16],
                             pcg_icp_stat* stat) {
  return guarded([&]() -> pcg_status {
    if (!bases || !params || !trans) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    if (n_dev < 1 || n_dev > 8) throw StatusError{PCG_E_INVALID_ARG, "1 to 8 devices"};
    check_view_args(target, n, stride, xyz_off);
    // contiguous slices of the target, one per device (the order of the float64 partial sums follows the devices)
    const Index* ix[8];
    void* d_tgt[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const void* d_ctgt[8];
    int64_t cnt[8];
    int prev = -1;
    cudaGetDevice(&prev);
    auto release = [&]() {
      for (int r = 0; r < n_dev; r++)
        if (d_tgt[r]) {
          cudaSetDevice(ix[r]->device);
          cudaFree(d_tgt[r]);
        }
      if (prev >= 0) cudaSetDevice(prev);
    };
    try {
      for (int r = 0; r < n_dev; r++) {
        if (!bases[r]) throw StatusError{PCG_E_INVALID_ARG, "null index"};
        ix[r] = bases[r]->ix;
      }
      for (int r = 0; r < n_dev; r++) {
        const int64_t lo = n * r / n_dev, hi = n * (r + 1) / n_dev;
        cnt[r] = hi - lo;
        PCG_CUDA(cudaSetDevice(ix[r]->device));
        PCG_CUDA(cudaMalloc(&d_tgt[r], (size_t)std::max<int64_t>(1, cnt[r] * stride)));
        if (cnt[r])
          PCG_CUDA(cudaMemcpyAsync(d_tgt[r], (const uint8_t*)target + lo * stride, (size_t)(cnt[r] * stride),
                                   cudaMemcpyHostToDevice, cudaStreamPerThread));
        d_ctgt[r] = d_tgt[r];
      }
      for (int r = 0; r < n_dev; r++) {
        PCG_CUDA(cudaSetDevice(ix[r]->device));
        PCG_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
      }
      const pcg_status rc = icp_fit_multi_device(n_dev, ix, d_ctgt, cnt, stride, xyz_off, *params, trans, stat);
      release();
      return rc;
    } catch (...) {
      release();
      throw;
    }
  });
}

pcg_status pcg_icp_shard_new(pcg_index* base, const void* d_target, int64_t n, int64_t stride, const int64_t xyz_off[3],
                             const pcg_icp_params* params, void* stream, pcg_icp_shard** out) {
  return guarded([&]() -> pcg_status {
    if (!base || !params || !out) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    *out = nullptr;
    check_view_args(d_target, n, stride, xyz_off);
    DeviceGuard g(base->ix->device);
    IcpShard* sh = icp_shard_new(*base->ix, make_view(d_target, n, stride, xyz_off), *params, (cudaStream_t)stream);
    *out = new pcg_icp_shard{sh, base->ix->device};
    return PCG_OK;
  });
}
void pcg_icp_shard_free(pcg_icp_shard* sh) {
  if (!sh) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(sh->device);
  cudaDeviceSynchronize();  // the stream-ordered workspace may still be in use
  icp_shard_free(sh->sh);
  if (prev >= 0) cudaSetDevice(prev);
  delete sh;
}
pcg_status pcg_icp_shard_partial(pcg_icp_shard* sh, double* d_partial16, void* stream) {
  return guarded([&]() -> pcg_status {
    if (!sh || !d_partial16) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    DeviceGuard g(sh->device);
    icp_shard_partial(*sh->sh, d_partial16, (cudaStream_t)stream);
    return PCG_OK;
  });
}
pcg_status pcg_icp_shard_finish(pcg_icp_shard* sh, const double* d_reduced16, void* stream) {
  return guarded([&]() -> pcg_status {
    if (!sh || !d_reduced16) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    DeviceGuard g(sh->device);
    icp_shard_finish(*sh->sh, d_reduced16, (cudaStream_t)stream);
    return PCG_OK;
  });
}
pcg_status pcg_icp_shard_result(pcg_icp_shard* sh, float trans[16], pcg_icp_stat* stat, int32_t* done, void* stream) {
  return guarded([&]() -> pcg_status {
    if (!sh || !trans) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    DeviceGuard g(sh->device);
    const pcg_status rc = icp_shard_result(*sh->sh, trans, stat, done, (cudaStream_t)stream);
    if (rc == PCG_E_NOT_ENOUGH_PAIRS) set_error("not enough correspondence pairs");
    return rc;
  });
}

pcg_status pcg_icp_finish(const double partial16[16], const pcg_icp_params* params, int32_t* iter, float trans[16],
                          pcg_evaluated* ev, int32_t* converged) {
  return guarded([&]() -> pcg_status {
    if (!partial16 || !params || !iter || !trans || !converged) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    pcg_status rc = icp_finish_host(partial16, *params, iter, trans, ev, converged);
    if (rc == PCG_E_NOT_ENOUGH_PAIRS) set_error("not enough correspondence pairs");
    return rc;
  });
}

}  // extern "C"
