// regiongrowing.cu — RegionGrowing.Segment (pc/segmentation/regiongrowing/regiongrowing.go:23-56)
// on top of the batched KDTree.Range of index.cu.
//
// The reference is a FIFO breadth-first search: the seed's Range neighbours are enqueued, every
// dequeued point whose label equals the label of the seed's nearest neighbour is emitted and its
// Range neighbours that were never enqueued before are appended.  A FIFO BFS is level-synchronous,
// so one level = ONE batched Range over the level's emitted points; the order in which the
// sequential loop would have enqueued the new points is the order of their FIRST occurrence in the
// concatenated (frontier order, then (DistSq, ID) inside a list) neighbour lists, which an atomicMin
// of the CSR position per point id recovers exactly.  Result: the reference's index list in the
// reference's order (for the canonical tie order of Range), never a host-side queue.
#include "bvh.cuh"

namespace pcg {

struct RegionGrowing {
  const Index* search = nullptr;  // not owned
  int device = 0;
  int64_t n = 0;
  float4* pl = nullptr;  // x, y, z, label bits per point id (Vec3At + Uint32At of the reference's accessors)
};

__global__ void __launch_bounds__(256)
    rg_pack_kernel(CloudView v, int32_t label_off, int label_aligned, float4* __restrict__ pl) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n) return;
  const float3 p = load_xyz(v, i);
  const uint8_t* r = v.data + i * v.stride + label_off;
  uint32_t lab;
  if (label_aligned)
    lab = __ldg((const uint32_t*)r);
  else
    lab = (uint32_t)r[0] | ((uint32_t)r[1] << 8) | ((uint32_t)r[2] << 16) | ((uint32_t)r[3] << 24);
  pl[i] = make_float4(p.x, p.y, p.z, __uint_as_float(lab));
}

// frontier -> flag of the points that are emitted / expanded (label == target)
__global__ void __launch_bounds__(256)
    rg_flag_kernel(const float4* __restrict__ pl, const uint32_t* __restrict__ frontier, uint32_t count,
                   uint32_t target, uint32_t* __restrict__ flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  flags[i] = __float_as_uint(pl[frontier[i]].w) == target ? 1u : 0u;
}

// compacts the flagged frontier points (order kept): their ids go to the result, their coordinates
// become the queries of this level
__global__ void __launch_bounds__(256)
    rg_emit_kernel(const float4* __restrict__ pl, const uint32_t* __restrict__ frontier, uint32_t count,
                   const uint32_t* __restrict__ flags, const long long* __restrict__ offs,
                   uint32_t* __restrict__ result, float* __restrict__ queries) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count || !flags[i]) return;
  const long long o = offs[i];
  const uint32_t id = frontier[i];
  const float4 p = pl[id];
  result[o] = id;
  queries[3 * o] = p.x;
  queries[3 * o + 1] = p.y;
  queries[3 * o + 2] = p.z;
}

// first occurrence of every point id in the concatenated neighbour lists (global CSR position)
__global__ void __launch_bounds__(256)
    rg_claim_kernel(const pcg_neighbor* __restrict__ nb, long long total, unsigned long long base,
                    unsigned long long* __restrict__ first_pos) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total) return;
  atomicMin(&first_pos[nb[j].id], base + (unsigned long long)j);
}
__global__ void __launch_bounds__(256)
    rg_new_flag_kernel(const pcg_neighbor* __restrict__ nb, long long total, unsigned long long base,
                       const unsigned long long* __restrict__ first_pos, uint32_t* __restrict__ flags) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total) return;
  flags[j] = first_pos[nb[j].id] == base + (unsigned long long)j ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
    rg_enqueue_kernel(const pcg_neighbor* __restrict__ nb, long long total, const uint32_t* __restrict__ flags,
                      const long long* __restrict__ offs, uint32_t* __restrict__ queue_tail) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total || !flags[j]) return;
  queue_tail[offs[j]] = (uint32_t)nb[j].id;
}
__global__ void rg_widen_kernel(const uint32_t* __restrict__ in, long long n, long long* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (long long)in[i];
}

RegionGrowing* region_growing_new_device(const Index& search, const CloudView& v, int64_t label_off,
                                         cudaStream_t stream) {
  search.wait(stream);
  RegionGrowing* rg = new RegionGrowing();
  rg->search = &search;
  rg->device = search.device;
  rg->n = v.n;
  if (v.n == 0) return rg;
  try {
    PCG_CUDA(cudaMallocAsync((void**)&rg->pl, (size_t)v.n * sizeof(float4), stream));
    const int aligned = ((v.stride & 3) == 0 && (label_off & 3) == 0 && (((uintptr_t)v.data) & 3) == 0) ? 1 : 0;
    PCG_LAUNCH(rg_pack_kernel, div_up(v.n, 256), 256, 0, stream, v, (int32_t)label_off, aligned, rg->pl);
  } catch (...) {
    if (rg->pl) cudaFree(rg->pl);
    delete rg;
    throw;
  }
  return rg;
}

void region_growing_free(RegionGrowing* rg) {
  if (!rg) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(rg->device);
  if (rg->pl) cudaFree(rg->pl);
  if (prev >= 0) cudaSetDevice(prev);
  delete rg;
}

static long long read_ll(const long long* d, cudaStream_t stream) {
  long long h = 0;
  PCG_CUDA(cudaMemcpyAsync(&h, d, sizeof(long long), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  return h;
}

// Segment: returns the number of indices; d_result (uint32[n]) holds them in the reference's order.
int64_t region_growing_segment_device(const RegionGrowing& rg, const float p[3], float max_range,
                                      DevBuf<uint32_t>& d_result, cudaStream_t stream) {
  const Index& ix = *rg.search;
  const uint32_t n = (uint32_t)rg.n;
  if (n == 0 || ix.n == 0) return 0;
  DevBuf<unsigned long long> first_pos(n, stream);
  PCG_CUDA(cudaMemsetAsync(first_pos.p, 0xff, first_pos.bytes(), stream));
  DevBuf<uint32_t> queue(n, stream);  // every point is enqueued at most once
  d_result.alloc(n, stream);
  DevBuf<float> queries((size_t)3 * std::max<uint32_t>(1, n), stream);
  static const int64_t kOff[3] = {0, 4, 8};

  uint32_t q_head = 0, q_tail = 0;  // the current frontier is queue[q_head, q_tail)
  int64_t emitted = 0;
  unsigned long long base = 0;
  uint32_t target = 0;
  bool seed_level = true;
  for (;;) {
    int64_t n_queries;
    if (seed_level) {  // regiongrowing.go:26: neighbors := r.search.Range(p, maxRange)
      PCG_CUDA(cudaMemcpyAsync(queries.p, p, 3 * sizeof(float), cudaMemcpyHostToDevice, stream));
      n_queries = 1;
    } else {
      const uint32_t count = q_tail - q_head;
      if (count == 0) break;
      DevBuf<uint32_t> flags(count, stream);
      DevBuf<long long> offs((size_t)count + 1, stream);
      PCG_LAUNCH(rg_flag_kernel, div_up(count, 256), 256, 0, stream, rg.pl, queue.p + q_head, count, target, flags.p);
      scan_counts(flags.p, offs.p, count, stream);
      PCG_LAUNCH(rg_emit_kernel, div_up(count, 256), 256, 0, stream, rg.pl, queue.p + q_head, count, flags.p, offs.p,
                 d_result.p + emitted, queries.p);
      n_queries = read_ll(offs.p + count, stream);
      emitted += n_queries;
      q_head = q_tail;
      if (n_queries == 0) break;
    }
    const CloudView qv = make_view(queries.p, n_queries, 12, kOff);
    DevBuf<long long> offsets;
    int64_t total = 0;
    range_count_device(ix, qv, max_range, offsets, &total, stream);
    if (total == 0) {
      if (seed_level) return 0;  // regiongrowing.go:27-29
      break;
    }
    DevBuf<pcg_neighbor> nb((size_t)total, stream);
    range_fill_device(ix, qv, max_range, offsets.p, total, nb.p, stream);
    if (seed_level) {  // regiongrowing.go:31: targetVal := Uint32At(neighbors[0].ID)
      pcg_neighbor first;
      PCG_CUDA(cudaMemcpyAsync(&first, nb.p, sizeof(first), cudaMemcpyDeviceToHost, stream));
      PCG_CUDA(cudaStreamSynchronize(stream));
      float4 pl;
      PCG_CUDA(cudaMemcpyAsync(&pl, rg.pl + first.id, sizeof(pl), cudaMemcpyDeviceToHost, stream));
      PCG_CUDA(cudaStreamSynchronize(stream));
      std::memcpy(&target, &pl.w, 4);
      seed_level = false;
    }
    // regiongrowing.go:47-52: append the neighbours that were never enqueued, in list order
    DevBuf<uint32_t> flags((size_t)total, stream);
    DevBuf<long long> offs((size_t)total + 1, stream);
    const int blocks = div_up(total, 256);
    PCG_LAUNCH(rg_claim_kernel, blocks, 256, 0, stream, nb.p, (long long)total, base, first_pos.p);
    PCG_LAUNCH(rg_new_flag_kernel, blocks, 256, 0, stream, nb.p, (long long)total, base, first_pos.p, flags.p);
    if (total >= ((int64_t)1 << 32)) throw StatusError{PCG_E_TOO_LARGE, "more than 2^32 neighbours in one BFS level"};
    scan_counts(flags.p, offs.p, (uint32_t)total, stream);
    PCG_LAUNCH(rg_enqueue_kernel, blocks, 256, 0, stream, nb.p, (long long)total, flags.p, offs.p, queue.p + q_tail);
    const long long fresh = read_ll(offs.p + total, stream);
    q_tail += (uint32_t)fresh;
    base += (unsigned long long)total;
  }
  return emitted;
}

void region_growing_widen(const uint32_t* d_in, int64_t n, long long* d_out, cudaStream_t stream) {
  if (n) PCG_LAUNCH(rg_widen_kernel, div_up(n, 256), 256, 0, stream, d_in, (long long)n, d_out);
}

}  // namespace pcg
