/* pcgol_b200.h — C ABI of the B200-native pcgol hot path (libpcgol_b200.so).
 *
 * This is the drop-in boundary: exactly what a cgo shim inside seqsense/pcgol
 * would bind to replace, for the data-parallel hot path only,
 *   - pc/storage/kdtree   (kdtree.New / KDTree.Nearest / KDTree.Range behind storage.Search)
 *   - pc/filter/voxelgrid (voxelGrid.Filter behind filter.Filter)
 *   - pc/registration/icp (NearestPointCorresponder.Pairs, PointToPointEvaluator.Evaluate,
 *                          gradientDescentUpdater.Update, PointToPointICPGradient.Fit)
 * Reference paths below are relative to the pcgol repository root.
 *
 * Conventions
 *   - Plain C types only.  Every call returns a pcg_status (0 == PCG_OK); the
 *     message of the last failure on the calling thread is pcg_last_error().
 *     Nothing aborts or throws across this boundary.
 *   - Clouds are passed the way pc.PointCloud stores them (pc/pointcloud.go:72-78):
 *     an interleaved little-endian record buffer, `stride` bytes per record,
 *     float32 x/y/z at byte offsets xyz_off[0..2].  A pc.Vec3Slice
 *     (pc/vec3slice.go:8) is stride 12, offsets {0,4,8}.
 *   - Host entry points copy their inputs to the device during the call and keep
 *     no host pointer after returning (cgo pointer-passing rule).  Buffers from
 *     pcg_host_alloc are pinned and copy at full PCIe rate.
 *   - `*_dev` entry points take device pointers that live on the index's /
 *     call's device, enqueue on `stream` (a cudaStream_t passed as void*, NULL =
 *     legacy default stream) and are used by pipelines that keep clouds resident
 *     in HBM, and by bench.py.
 *   - Every entry point selects its CUDA device explicitly, so calls may come
 *     from any OS thread (goroutines migrate).  Concurrent queries on one index
 *     are safe (KDTree.Nearest/Range are goroutine-safe, kdtree.go:44-50).
 *   - There is no CPU fallback: without a usable CUDA device every compute entry
 *     point fails with PCG_E_NO_DEVICE / PCG_E_CUDA.
 */
#ifndef PCGOL_B200_H_
#define PCGOL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: pcg_icp_params gained min_dist_sq + updater (appended); DeletePoint, MinDistSq search, Hessian,
 *    region growing and PCD entry points added. */
#define PCGOL_B200_ABI_VERSION 3

typedef int32_t pcg_status;
enum {
  PCG_OK = 0,
  PCG_E_INVALID_ARG = 1,
  /* pc/minmax.go:10-12: MinMaxVec3 on an empty cloud -> errors.New("no point") */
  PCG_E_NO_POINT = 2,
  /* The reference would panic here (index out of range on its dense voxel array,
   * pc/filter/voxelgrid/voxelgrid.go:46,151, or on its chunk table, :89). */
  PCG_E_REF_WOULD_PANIC = 3,
  /* The reference's result is implementation-specific here (float->int conversion
   * of a non-finite or out-of-range value). */
  PCG_E_REF_UNDEFINED = 4,
  /* pc/registration/icp/evaluator.go:15-17,92-106: ErrNotEnoughPairs */
  PCG_E_NOT_ENOUGH_PAIRS = 5,
  PCG_E_CUDA = 6,
  PCG_E_NO_DEVICE = 7,
  PCG_E_TOO_LARGE = 8,
  /* pc.Unmarshal (pc/io.go): strconv.ErrSyntax / header validation errors, io.EOF / io.ErrUnexpectedEOF,
   * lzf.ErrDataCorruption / "wrong uncompressed size" */
  PCG_E_PCD_SYNTAX = 9,
  PCG_E_PCD_EOF = 10,
  PCG_E_PCD_CORRUPT = 11,
  /* errors.New("invalid field name") (pc/pointcloud.go:115,189) */
  PCG_E_INVALID_FIELD = 12
};

/* Mirrors Go's storage.Neighbor{ID int; DistSq float32} (pc/storage/search.go:8-11)
 * as laid out on 64-bit targets (16 bytes), so a []storage.Neighbor can be passed
 * as is.  A miss is {ID:-1, DistSq:maxRange*maxRange} (kdtree.go:84-86,100-103). */
typedef struct pcg_neighbor {
  int64_t id;
  float dist_sq;
  uint32_t pad_;
} pcg_neighbor;

/* ---- library ------------------------------------------------------------- */
int32_t pcg_abi_version(void);
const char* pcg_last_error(void);
const char* pcg_status_string(pcg_status s);
int32_t pcg_device_count(void);
/* Pinned host memory (optional; any host pointer is accepted everywhere). */
pcg_status pcg_host_alloc(void** out, int64_t bytes);
void pcg_host_free(void* p);
/* Device memory on `device` for callers that keep clouds resident (Go has no CUDA runtime). */
pcg_status pcg_device_alloc(int32_t device, void** out, int64_t bytes);
void pcg_device_free(int32_t device, void* p);
pcg_status pcg_memcpy_h2d(int32_t device, void* dst_dev, const void* src_host, int64_t bytes);
pcg_status pcg_memcpy_d2h(int32_t device, void* dst_host, const void* src_dev, int64_t bytes);
pcg_status pcg_device_synchronize(int32_t device);
/* Number of kernels this library has launched on the calling thread's behalf, process-wide. */
int64_t pcg_kernel_launch_count(void);
/* Diagnostics: when enabled, every kernel launch is bracketed by CUDA events on its
 * stream; pcg_profile_report synchronises and writes {"kernel": {"launches", "total_ms"}}
 * as JSON into buf (returns the length needed) and clears the records. */
/* Test hook: the float32 sum of x[0..n) accumulated one element at a time in index order (the
 * order of PointToPointEvaluator.Evaluate, evaluator.go:122-145), computed on the device by the
 * strict-mode replay: exact_path 1 = binade-parallel exact replay, 0 = plain sequential chain.
 * out is float[4]: the sum, then (exact path only) chunks taken as one integer add, chunks replayed
 * element by element, and SM cycles of the in-order walk. */
pcg_status pcg_debug_sequential_sum_f32(const float* x, int64_t n, int32_t device, int32_t exact_path, float* out);
void pcg_profile_enable(int32_t on);
/* Test hook: which VoxelGrid pipeline pcg_voxelgrid_filter* runs. 0 = automatic (one cooperative kernel up to 1.2M
 * points, the multi-kernel LSD radix pipeline above), 1 = always the multi-kernel pipeline.  Both produce the same
 * bytes; the parity tests run both. */
void pcg_debug_set_vg_path(int32_t path);
int64_t pcg_profile_report(char* buf, int64_t cap);

/* ---- storage.Search: spatial index (replaces kdtree.New, kdtree.go:33-56) --- */
typedef struct pcg_index pcg_index;

pcg_status pcg_index_build(const void* data, int64_t n, int64_t stride, const int64_t xyz_off[3], int32_t device,
                           pcg_index** out);
pcg_status pcg_index_build_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                               int32_t device, void* stream, pcg_index** out);
void pcg_index_free(pcg_index* idx);
int64_t pcg_index_len(const pcg_index* idx);    /* Vec3RandomAccessor.Len */
int32_t pcg_index_device(const pcg_index* idx);
int64_t pcg_index_device_bytes(const pcg_index* idx); /* HBM held by the index */
/* Test hook: copies the index's point slots (leaves*8 float4: x, y, z, original id as bits; tail padding
 * has id 0xffffffff) to out; *slots receives their number.  Lets the tests check the build invariants. */
pcg_status pcg_debug_index_slots(const pcg_index* idx, float* out, int64_t cap_slots, int64_t* slots);

/* KDTree.Nearest for a batch (kdtree.go:83-92).  Exact search (MinDistSq == 0).
 * Result i is the (DistSq, ID)-lexicographic minimum over points with
 * DistSq < maxRange*maxRange, i.e. the reference's own brute-force oracle
 * (kdtree_test.go:955-968); DistSq is ((dx*dx+dy*dy)+dz*dz) in float32 with
 * d = point - query, no FMA (mat/vec3.go:18-20,38-40). */
pcg_status pcg_index_nearest(pcg_index* idx, const void* q, int64_t nq, int64_t q_stride,
                             const int64_t q_xyz_off[3], float max_range, pcg_neighbor* out);
/* Device variant: d_ids int32[nq] (-1 = miss), d_dist_sq float[nq]. */
pcg_status pcg_index_nearest_dev(pcg_index* idx, const void* d_q, int64_t nq, int64_t q_stride,
                                 const int64_t q_xyz_off[3], float max_range, int32_t* d_ids, float* d_dist_sq,
                                 void* stream);

/* KDTree.MinDistSq > 0 (kdtree.go:19-22,104,120,140): approximate search that may stop at the
 * first candidate with DistSq < min_dist_sq.  Every answer is the exact nearest neighbour or a real
 * point of the cloud with DistSq < min_dist_sq (and < maxRange^2), DistSq always belonging to the
 * returned ID — the contract the reference's answers satisfy; WHICH close point is returned depends
 * on the traversal there as here, so IDs are not comparable between implementations.  One
 * reference quirk is not reproduced: with maxRange^2 < MinDistSq the reference can report a miss
 * after looking at a single leaf (kdtree.go:100-106); this library still returns the exact answer.
 * min_dist_sq == 0 is the exact search. */
pcg_status pcg_index_nearest_approx(pcg_index* idx, const void* q, int64_t nq, int64_t q_stride,
                                    const int64_t q_xyz_off[3], float max_range, float min_dist_sq,
                                    pcg_neighbor* out);
pcg_status pcg_index_nearest_approx_dev(pcg_index* idx, const void* d_q, int64_t nq, int64_t q_stride,
                                        const int64_t q_xyz_off[3], float max_range, float min_dist_sq,
                                        int32_t* d_ids, float* d_dist_sq, void* stream);

/* KDTree.DeletePoint (kdtree.go:322-332) for a batch of ids: afterwards no search returns them
 * (naiveSearch.deletePoint, kdtree_test.go:1003-1005, is the oracle).  Len() is unchanged, deleting
 * an id twice is a no-op, and an id outside [0, Len()-1] fails with PCG_E_INVALID_ARG ("%d does not
 * correspond to any point in the tree") without deleting anything.  Like the reference, a delete
 * must not run concurrently with searches on the same index. */
pcg_status pcg_index_delete_points(pcg_index* idx, const int64_t* ids, int64_t n);

/* KDTree.Range for a batch (kdtree.go:148-161): all points with DistSq < maxRange^2
 * (strict), each list sorted by (DistSq, ID) — the canonical order of
 * kdtree_test.go:926-941.  CSR result owned by the library. */
typedef struct pcg_range_result pcg_range_result;
pcg_status pcg_index_range(pcg_index* idx, const void* q, int64_t nq, int64_t q_stride,
                           const int64_t q_xyz_off[3], float max_range, pcg_range_result** out);
/* Two-call form with caller-owned buffers (lets a Go caller allocate / reuse / pin the result):
 * count fills offsets[nq+1]; fill writes offsets[nq] neighbours, list i at [offsets[i], offsets[i+1]).
 * The offsets passed to fill must come from count with the same queries and range. */
pcg_status pcg_index_range_count(pcg_index* idx, const void* q, int64_t nq, int64_t q_stride,
                                 const int64_t q_xyz_off[3], float max_range, int64_t* offsets);
pcg_status pcg_index_range_fill(pcg_index* idx, const void* q, int64_t nq, int64_t q_stride,
                                const int64_t q_xyz_off[3], float max_range, const int64_t* offsets,
                                pcg_neighbor* out);
int64_t pcg_range_total(const pcg_range_result* r);
const int64_t* pcg_range_offsets(const pcg_range_result* r);       /* nq+1 entries */
const pcg_neighbor* pcg_range_neighbors(const pcg_range_result* r); /* total entries */
void pcg_range_free(pcg_range_result* r);

/* ---- pc.PointCloud resident in HBM + its PCD encodings (pc/pointcloud.go:9-78, pc/io.go) -------- */
/* The steps either side of the hot path: a cloud is parsed / uploaded once, then VoxelGrid, the index
 * build and ICP run on the resident records, and Marshal reads the result back once. */
typedef struct pcg_cloud pcg_cloud;
#define PCG_MAX_FIELDS 32
typedef struct pcg_cloud_header { /* pc.PointCloudHeader + Points + len(Data) */
  float version;
  int32_t n_fields;
  char fields[PCG_MAX_FIELDS][32];
  char type[PCG_MAX_FIELDS][8];
  int64_t size[PCG_MAX_FIELDS];
  int64_t count[PCG_MAX_FIELDS];
  int64_t width, height;
  int32_t n_viewpoint;
  float viewpoint[16];
  int64_t points;
  int64_t data_bytes;
} pcg_cloud_header;
/* pc.Unmarshal (io.go:32-45): ascii, binary and binary_compressed (LZF on the host, the field-major payload
 * is transposed into records on the device, reproducing io.go:205-228 including its handling of COUNT > 1).
 * Errors map to PCG_E_PCD_*; inputs on which the reference would panic (index out of range) give
 * PCG_E_REF_WOULD_PANIC. */
pcg_status pcg_pcd_unmarshal(const void* pcd, int64_t len, int32_t device, pcg_cloud** out);
/* pc.Marshal (io.go:232-285): header text + records, always "DATA binary".  *len receives the size needed;
 * the call fails with PCG_E_INVALID_ARG if cap is smaller. */
pcg_status pcg_pcd_marshal(const pcg_cloud* c, void* buf, int64_t cap, int64_t* len);
/* A cloud from a host record buffer (header->points records of stride sum(size*count)). */
pcg_status pcg_cloud_upload(const pcg_cloud_header* header, const void* data, int32_t device, pcg_cloud** out);
pcg_status pcg_cloud_get_header(const pcg_cloud* c, pcg_cloud_header* out);
pcg_status pcg_cloud_download(const pcg_cloud* c, void* data, int64_t cap);
const void* pcg_cloud_device_ptr(const pcg_cloud* c);
void pcg_cloud_free(pcg_cloud* c);
/* voxelGrid.Filter on a resident cloud: header cloned, Width = output points, Height = 1
 * (voxelgrid.go:119-128).  x/y/z are resolved like PointCloud.Vec3Iterator (pointcloud.go:130-171). */
pcg_status pcg_cloud_voxelgrid_filter(const pcg_cloud* in, const float leaf[3], const int64_t chunk[3],
                                      pcg_cloud** out);
/* kdtree.New(pp.Vec3Iterator()) on a resident cloud. */
pcg_status pcg_cloud_index_build(const pcg_cloud* c, pcg_index** out);
/* (pcg_cloud_icp_fit, the Fit with a resident target, is declared with the ICP entry points below.) */

/* ---- segmentation: RegionGrowing (pc/segmentation/regiongrowing/regiongrowing.go:11-56) ---- */
/* regiongrowing.New(search, propertyIter): `search` is the index, the property is a uint32 field at byte
 * label_off of the same records the index was built from (pc.Uint32RandomAccessor over the cloud).  The handle
 * keeps its own device copy of (x, y, z, label); it borrows `search`, which must outlive it. */
typedef struct pcg_region_growing pcg_region_growing;
pcg_status pcg_region_growing_new(pcg_index* search, const void* data, int64_t n, int64_t stride,
                                  const int64_t xyz_off[3], int64_t label_off, pcg_region_growing** out);
void pcg_region_growing_free(pcg_region_growing* rg);
/* RegionGrowing.Segment(p, maxRange) (regiongrowing.go:23-56): ids of the points reached from the Range
 * neighbours of p through chains of points closer than maxRange that all carry the label of p's nearest
 * neighbour, in the reference's breadth-first order (for Range lists in the canonical (DistSq, ID) order;
 * the reference leaves equal-DistSq ties unordered and its own test sorts the result,
 * regiongrowing_test.go:186).  Every level of the search is one batched Range on the device.
 * indice must hold `cap` entries; if the result is larger the call fails with *n_out = size needed
 * (pcg_index_len is always enough). */
pcg_status pcg_region_growing_segment(pcg_region_growing* rg, const float p[3], float max_range, int64_t* indice,
                                      int64_t cap, int64_t* n_out);

/* ---- filter.Filter: VoxelGrid (pc/filter/voxelgrid/voxelgrid.go:35-187) ------ */
/* chunk = Options.ChunkSize (option.go:7-18); any zero selects the un-chunked path
 * (voxelgrid.go:45-47).  `out` must hold n*stride bytes; *n_out receives the number of
 * records written.  Output records, their order (ascending chunk id, then ascending
 * voxel key x + xs*(y + ys*z)), the kept fields of each voxel's first point and the
 * centroid sum*(1/num)+vMin reproduce the reference bit for bit. */
pcg_status pcg_voxelgrid_filter(const void* data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                                const float leaf[3], const int64_t chunk[3], int32_t device, void* out,
                                int64_t* n_out);
pcg_status pcg_voxelgrid_filter_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                                    const float leaf[3], const int64_t chunk[3], int32_t device, void* d_out,
                                    int64_t* n_out, void* stream);
/* One large VoxelGrid over several GPUs.  Chunks are independent units in the reference (voxelgrid.go:102-116, outputs
 * concatenated in ascending chunk id), so with the cloud replicated every rank filters a range of chunk ids and the
 * ranks' outputs, concatenated in rank order, are the reference's output byte for byte.
 * pcg_voxelgrid_chunk_histogram_dev: points per chunk id (hist == NULL or cap too small: only *n_chunks is set) - what
 * a caller balances the ranges with; sample_step = 1 counts every point, s > 1 one run of 32 points out of every 32*s
 * (any partition of the chunk ids gives the reference's output, the histogram only balances the load).
 * pcg_voxelgrid_filter_chunks_dev: the Filter restricted to chunk ids [cid_lo, cid_hi); MinMaxVec3, the voxel grid and the chunk table are those of the whole cloud.  For an un-chunked
 * filter (any ChunkSize component zero; one chunk) the ids are the top <= 12 bits of the voxel key instead: voxels are
 * independent and emitted in ascending key order (voxelgrid.go:172-184), so key ranges concatenate the same way. */
pcg_status pcg_voxelgrid_chunk_histogram_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                                             const float leaf[3], const int64_t chunk[3], int32_t device,
                                             int64_t sample_step, int64_t* hist, int64_t cap, int64_t* n_chunks,
                                             void* stream);
pcg_status pcg_voxelgrid_filter_chunks_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                                           const float leaf[3], const int64_t chunk[3], int64_t cid_lo, int64_t cid_hi,
                                           int32_t device, void* d_out, int64_t* n_out, void* stream);
/* The same Filter with the POINTS sharded: every rank holds a contiguous slice of the cloud (slice r before slice
 * r+1 in point order) and nothing is replicated.  Steps, each a call below with a collective of the caller in between:
 *   pcg_minmax_packed_dev        the slice's six MinMaxVec3 accumulators as 64-bit words that ONE signed MIN
 *                                all-reduce combines (the maxima are stored negated): order-preserving value bits,
 *                                then the global index of the point (first occurrence wins across ranks,
 *                                minmax.go:17-22), then the sign of a zero; a NaN at global point 0 (which the
 *                                reference never replaces, minmax.go:13) is a marker word.  pcgol_b200/dist.py
 *                                (decode_minmax_words) turns the reduced words into the six floats (< 2^30 points)
 *   ..._chunk_histogram_mm_dev   points per chunk id of the slice under the bounds mm6 = {min xyz, max xyz} of the
 *                                whole cloud; summed over ranks it balances the chunk ranges
 *   ..._owner_order_dev          cuts[r] = first chunk id of rank r (cuts[0] = 0): d_perm = the slice's points ordered
 *                                by owner (stable), counts[r] = points for rank r, d_send (optional, n * stride
 *                                bytes) = the whole records in that order -> all-to-all
 *   ..._filter_chunks_mm_dev     the owner filters the records it received (sources in rank order = global point
 *                                order, which the stable sort keeps) under the same bounds; d_data must hold exactly
 *                                the points of its chunks [cid_lo, cid_hi) - nothing is selected any more
 * Outputs concatenated in rank order are the reference's output (voxelgrid.go:102-133). */
pcg_status pcg_minmax_packed_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                                 int32_t device, int64_t index_base, int64_t* d_out6, void* stream);
pcg_status pcg_voxelgrid_chunk_histogram_mm_dev(const void* d_data, int64_t n, int64_t stride,
                                                const int64_t xyz_off[3], const float leaf[3], const int64_t chunk[3],
                                                const float mm6[6], int32_t device, int64_t sample_step, int64_t* hist,
                                                int64_t cap, int64_t* n_chunks, void* stream);
pcg_status pcg_voxelgrid_owner_order_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                                         const float leaf[3], const int64_t chunk[3], const float mm6[6],
                                         const int64_t* cuts, int32_t world, int32_t device, uint32_t* d_perm,
                                         int64_t* counts, void* d_send, void* stream);
pcg_status pcg_voxelgrid_filter_chunks_mm_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3],
                                              const float leaf[3], const int64_t chunk[3], const float mm6[6],
                                              int64_t cid_lo, int64_t cid_hi, int32_t device, void* d_out,
                                              int64_t* n_out, void* stream);
/* MinMaxVec3 (pc/minmax.go:9-26) — first step of the filter, exposed for sharded runs. */
pcg_status pcg_minmax_dev(const void* d_data, int64_t n, int64_t stride, const int64_t xyz_off[3], int32_t device,
                          float mn[3], float mx[3], void* stream);

/* ---- pc/registration/icp ---------------------------------------------------- */
enum {
  /* The nine float32 sums of Evaluate are accumulated sequentially in target order,
   * exactly as evaluator.go:122-145 does: bit-identical trajectory. */
  PCG_ICP_STRICT = 0,
  /* Fixed-shape float64 tree reduction (deterministic, independent of GPU count up to
   * float64 rounding); differs from the reference by its float32 summation error. */
  PCG_ICP_FAST = 1,
  /* OR-ed into the mode of pcg_icp_evaluate: also fill Evaluated.Hessian (see pcg_evaluated). */
  PCG_ICP_WITH_HESSIAN = 0x100
};

enum {
  /* gradientDescentUpdater (updater.go:39-71): the reference's only updater. */
  PCG_UPDATER_GRADIENT_DESCENT = 0,
  /* Solves the 6x6 normal equations H d = -g accumulated by Evaluate (float64 Cholesky) and applies
   * the increment the way the reference applies its own (Translate * Rodrigues * trans,
   * updater.go:65-68); same convergence test on the gradient (updater.go:45-54) and the same
   * MaxIteration cap.  Weight is ignored.  Not in the reference: a consumer of the
   * Evaluated.Hessian / HasHessian() hooks it declares (evaluator.go:25-36,76). */
  PCG_UPDATER_GAUSS_NEWTON = 1
};

/* Zero means the reference default in every field. */
typedef struct pcg_icp_params {
  float max_dist;        /* NearestPointCorresponder.MaxDist (correspondence.go:18-20) */
  int32_t min_pairs;     /* PointToPointEvaluator.MinPairs, 0 -> 6 (evaluator.go:92-95) */
  float weight[6];       /* GradientDescentUpdaterFactory.Weight, all 0 -> 0.3 (updater.go:15,25-27) */
  float threshold[6];    /* .Threshold, all 0 -> 0.01 (updater.go:16,28-30) */
  int32_t max_iteration; /* .MaxIteration, 0 -> 20 (updater.go:31-33) */
  int32_t mode;          /* PCG_ICP_STRICT | PCG_ICP_FAST */
  /* ABI 2 */
  float min_dist_sq;     /* KDTree.MinDistSq of the base search (kdtree.go:19-22); 0 = exact.  The reference's own
                          * ICP test and benchmark set 0.01 (icp_test.go:58,118-120). */
  int32_t updater;       /* PCG_UPDATER_GRADIENT_DESCENT | PCG_UPDATER_GAUSS_NEWTON */
  /* ABI 3 */
  int32_t weight_fn;     /* PointToPointEvaluator.WeightFn (evaluator.go:19-23,72), see pcg_weight_fn; 0 = default */
  float weight_param;    /* threshold of PCG_WEIGHT_TRUNCATED / k^2 of PCG_WEIGHT_HUBER (a squared distance) */
} pcg_icp_params;

/* EvaluateWeightFn is a Go closure in the reference (evaluator.go:19); a closure cannot run on the device, so the ABI
 * offers the family real callers use, evaluated per pair in float32 with every operation rounded:
 *   PCG_WEIGHT_CONSTANT   w = 1                                 DefaultEvaluateWeightFn (evaluator.go:21-23)
 *   PCG_WEIGHT_TRUNCATED  w = dsq < param ? 1 : 0               hard rejection of far pairs (they still count as pairs)
 *   PCG_WEIGHT_HUBER      w = dsq <= param ? 1 : sqrt(param / dsq)   Huber weight with k^2 = param
 * The weight multiplies every term exactly as evaluator.go:130-144 does (Value += w*dsq, SumW += w, G += w*...).
 * The Hessian extension (PCG_ICP_WITH_HESSIAN / Gauss-Newton) needs PCG_WEIGHT_CONSTANT. */
typedef enum pcg_weight_fn { PCG_WEIGHT_CONSTANT = 0, PCG_WEIGHT_TRUNCATED = 1, PCG_WEIGHT_HUBER = 2 } pcg_weight_fn;

/* icp.Evaluated (evaluator.go:25-30).  The reference never writes Hessian (HasHessian() == false,
 * evaluator.go:76): it is zero unless PCG_ICP_WITH_HESSIAN / PCG_UPDATER_GAUSS_NEWTON is selected.
 * Then it is the Gauss-Newton Hessian of Value, 2f * sum J^T J with J = [I | -[pt]x] per pair (pt =
 * transformed target point, f = 1/sum w as in evaluator.go:156-159), accumulated in float64;
 * 6x6 symmetric, index = col*6 + row, parameter order (tx,ty,tz,rx,ry,rz) like Gradient. */
typedef struct pcg_evaluated {
  float value;
  float gradient[6];
  float hessian[36];
  float dist_rms;
} pcg_evaluated;

/* icp.Stat (stat.go:3-6) + the pair count of the last Evaluate. */
typedef struct pcg_icp_stat {
  pcg_evaluated evaluated;
  int32_t num_iteration;
  int32_t pad_;
  int64_t n_pairs;
} pcg_icp_stat;

/* NearestPointCorresponder.Pairs (correspondence.go:22-37): arrays sized n_target. */
pcg_status pcg_icp_pairs(pcg_index* base, const void* target, int64_t n, int64_t stride, const int64_t xyz_off[3],
                         float max_dist, int64_t* base_id, int64_t* target_id, float* dist_sq, int64_t* n_pairs);

/* Pairs over a base search with KDTree.MinDistSq = min_dist_sq (see pcg_index_nearest_approx). */
pcg_status pcg_icp_pairs_approx(pcg_index* base, const void* target, int64_t n, int64_t stride,
                                const int64_t xyz_off[3], float max_dist, float min_dist_sq, int64_t* base_id,
                                int64_t* target_id, float* dist_sq, int64_t* n_pairs);

/* PointToPointEvaluator.Evaluate (evaluator.go:91-189), default weight function. */
pcg_status pcg_icp_evaluate(pcg_index* base, const void* target, int64_t n, int64_t stride,
                            const int64_t xyz_off[3], float max_dist, int32_t min_pairs, int32_t mode,
                            pcg_evaluated* out, int64_t* n_pairs);

/* Evaluate with the full parameter block: max_dist, min_pairs, mode (| PCG_ICP_WITH_HESSIAN) and
 * min_dist_sq are used, the updater fields are not. */
pcg_status pcg_icp_evaluate_params(pcg_index* base, const void* target, int64_t n, int64_t stride,
                                   const int64_t xyz_off[3], const pcg_icp_params* params, pcg_evaluated* out,
                                   int64_t* n_pairs);

/* PointToPointICPGradient.Fit (icp.go:23-67) with the default gradient-descent updater
 * (updater.go:44-71).  trans is column-major (mat/mat4.go:8-10).  On
 * PCG_E_NOT_ENOUGH_PAIRS trans/stat hold the values reached so far (icp.go:51-53). */
pcg_status pcg_icp_fit(pcg_index* base, const void* target, int64_t n, int64_t stride, const int64_t xyz_off[3],
                       const pcg_icp_params* params, float trans[16], pcg_icp_stat* stat);
pcg_status pcg_icp_fit_dev(pcg_index* base, const void* d_target, int64_t n, int64_t stride,
                           const int64_t xyz_off[3], const pcg_icp_params* params, float trans[16],
                           pcg_icp_stat* stat, void* stream);

/* PointToPointICPGradient.Fit with a resident target. */
pcg_status pcg_cloud_icp_fit(pcg_index* base, const pcg_cloud* target, const pcg_icp_params* params, float trans[16],
                             pcg_icp_stat* stat);

/* Scan-pair farm: `count` independent (base, target) pairs, device resident, all on
 * `device`; builds one index per pair and runs the fits concurrently. status_out[i]
 * is the per-pair status. */
pcg_status pcg_icp_fit_pairs_dev(int32_t count, const void* const* d_base, const int64_t* n_base,
                                 const void* const* d_target, const int64_t* n_target, int64_t stride,
                                 const int64_t xyz_off[3], const pcg_icp_params* params, int32_t device,
                                 float* trans_out /* count*16 */, pcg_icp_stat* stat_out, pcg_status* status_out,
                                 void* stream);

/* One large ICP sharded over GPUs (target points split across ranks, base index
 * replicated).  Per iteration each rank calls pcg_icp_partial_dev on its slice, the
 * 16 doubles are sum-all-reduced by the caller (NCCL), then every rank applies
 * pcg_icp_finish identically.  d_partial16: {Value, SumW, G0..5, R, nPairs, 0...}. */
pcg_status pcg_icp_partial_dev(pcg_index* base, const void* d_target, int64_t n, int64_t stride,
                               const int64_t xyz_off[3], float max_dist, const float trans[16], int32_t first,
                               const uint32_t* d_visit_order /* optional, from pcg_query_order_dev */,
                               double* d_partial16, void* stream);
/* Morton visit order of a query / target cloud relative to the index (a permutation of
 * 0..n-1 in d_order): coherent warps for repeated searches over the same cloud. */
pcg_status pcg_query_order_dev(pcg_index* idx, const void* d_q, int64_t n, int64_t stride, const int64_t xyz_off[3],
                               uint32_t* d_order, void* stream);
/* The same sharded Fit with the loop resident on the device: no host round trip per iteration.  Every rank creates
 * a shard over its slice of the target, then per iteration enqueues pcg_icp_shard_partial (16 float64 into
 * d_partial16), all-reduces that buffer in stream order (NCCL) and enqueues pcg_icp_shard_finish, which applies the
 * tail of Evaluate + gradientDescentUpdater.Update to the device-resident state; once converged / failed the
 * remaining calls fall through.  pcg_icp_shard_result synchronises `stream` and returns the transform, Stat and
 * whether the loop has ended (status PCG_E_NOT_ENOUGH_PAIRS like Fit).  Float64 sums, gradient-descent updater. */
/* ---- multi-GPU from ONE process (the Go shim has no ranks): the device list is the list of index replicas.
 * pcg_index_replicate copies a built index to another device (NVLink peer copy; the replica is bit-identical).
 * pcg_icp_fit_multi* = PointToPointICPGradient.Fit (icp.go:23-67) with the target split over the replicas' devices:
 * every device runs the fast loop on its slice, one kernel per iteration; the last CTA of that kernel stores the
 * device's ten float64 partial sums into every peer's exchange buffer over NVLink (peer-mapped memory), waits on flags,
 * adds the slots in device order and applies the identical update - no library collective, no extra kernel and no
 * host round trip per iteration; the host enqueues all iterations and joins once.  Fast mode, gradient-descent
 * updater, exact nearest neighbour (min_dist_sq == 0).  The transform equals the single-GPU fast Fit up to the
 * rounding of the float64 sums.  n_dev == 1 is allowed (no exchange).  _dev: d_targets[r] (n_targets[r] records) is
 * device memory on the device of bases[r]; the host variant cuts `target` into contiguous slices itself. */
pcg_status pcg_index_replicate(pcg_index* src, int32_t device, pcg_index** out);
pcg_status pcg_icp_fit_multi_dev(int32_t n_dev, pcg_index* const* bases, const void* const* d_targets,
                                 const int64_t* n_targets, int64_t stride, const int64_t xyz_off[3],
                                 const pcg_icp_params* params, float trans[16], pcg_icp_stat* stat);
pcg_status pcg_icp_fit_multi(int32_t n_dev, pcg_index* const* bases, const void* target, int64_t n, int64_t stride,
                             const int64_t xyz_off[3], const pcg_icp_params* params, float trans[16],
                             pcg_icp_stat* stat);

typedef struct pcg_icp_shard pcg_icp_shard;
pcg_status pcg_icp_shard_new(pcg_index* base, const void* d_target, int64_t n, int64_t stride, const int64_t xyz_off[3],
                             const pcg_icp_params* params, void* stream, pcg_icp_shard** out);
void pcg_icp_shard_free(pcg_icp_shard* sh);
pcg_status pcg_icp_shard_partial(pcg_icp_shard* sh, double* d_partial16, void* stream);
pcg_status pcg_icp_shard_finish(pcg_icp_shard* sh, const double* d_reduced16, void* stream);
pcg_status pcg_icp_shard_result(pcg_icp_shard* sh, float trans[16], pcg_icp_stat* stat, int32_t* done, void* stream);

/* Host-side tail of Evaluate (evaluator.go:156-186) + Update (updater.go:44-71) from
 * all-reduced sums.  *iter is the updater's i; returns converged in *converged. */
pcg_status pcg_icp_finish(const double partial16[16], const pcg_icp_params* params, int32_t* iter, float trans[16],
                          pcg_evaluated* ev, int32_t* converged);

#ifdef __cplusplus
}
#endif
#endif /* PCGOL_B200_H_ */
