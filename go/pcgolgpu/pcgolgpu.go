package pcgolgpu

/*
#cgo LDFLAGS: -lpcgol_b200
#include <stdlib.h>
#include "pcgol_b200.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"io"
	"runtime"
	"strconv"
	"unsafe"

	"github.com/seqsense/pcgol/mat"
	"github.com/seqsense/pcgol/pc"
	"github.com/seqsense/pcgol/pc/filter"
	"github.com/seqsense/pcgol/pc/filter/voxelgrid"
	"github.com/seqsense/pcgol/pc/registration/icp"
	"github.com/seqsense/pcgol/pc/storage"
)

// Status errors that have no counterpart in the reference (it would panic there).
var ErrReferenceWouldPanic = errors.New("pcgolgpu: reference would panic (voxel/chunk index out of range)")

// ErrDataCorruption stands for lzf.ErrDataCorruption (github.com/zhuyie/golzf) of pc.Unmarshal.
var ErrDataCorruption = errors.New("lzf: data corruption")

func statusError(s C.pcg_status) error {
	switch s {
	case C.PCG_OK:
		return nil
	case C.PCG_E_NO_POINT:
		return errors.New("no point") // pc/minmax.go:10-12
	case C.PCG_E_NOT_ENOUGH_PAIRS:
		return icp.ErrNotEnoughPairs // pc/registration/icp/evaluator.go:15-17
	case C.PCG_E_REF_WOULD_PANIC:
		return ErrReferenceWouldPanic
	case C.PCG_E_PCD_SYNTAX: // wraps the class of error pc.Unmarshal returns (pc/io_test.go:108-165)
		return fmt.Errorf("%s: %w", C.GoString(C.pcg_last_error()), strconv.ErrSyntax)
	case C.PCG_E_PCD_EOF:
		return io.EOF
	case C.PCG_E_PCD_CORRUPT:
		return fmt.Errorf("%s: %w", C.GoString(C.pcg_last_error()), ErrDataCorruption)
	case C.PCG_E_INVALID_FIELD:
		return errors.New("invalid field name") // pc/pointcloud.go:115,189
	}
	return fmt.Errorf("pcgolgpu: status %d: %s", int(s), C.GoString(C.pcg_last_error()))
}

// pin keeps the calling goroutine on one OS thread until the returned function runs: pcg_last_error is
// thread-local in C, and a goroutine may migrate between the failing call and the call that fetches its message.
// Every method that may call statusError starts with `defer pin()()`.
func pin() func() {
	runtime.LockOSThread()
	return runtime.UnlockOSThread
}

// flatten turns any Vec3RandomAccessor into (pointer, n, stride, xyz offsets).
// Fast paths pass the caller's memory as is (no Go pointer is retained by C after the call):
// pc.Vec3Slice is []mat.Vec3 = contiguous [3]float32 records.
func flatten(ra pc.Vec3RandomAccessor) (unsafe.Pointer, C.int64_t, C.int64_t, [3]C.int64_t, []float32) {
	off := [3]C.int64_t{0, 4, 8}
	if v, ok := ra.(pc.Vec3Slice); ok {
		if len(v) == 0 {
			return nil, 0, 12, off, nil
		}
		return unsafe.Pointer(&v[0]), C.int64_t(len(v)), 12, off, nil
	}
	buf := make([]float32, 3*ra.Len()) // generic path: copy through Vec3At
	for i := 0; i < ra.Len(); i++ {
		p := ra.Vec3At(i)
		copy(buf[3*i:], p[:])
	}
	if len(buf) == 0 {
		return nil, 0, 12, off, buf
	}
	return unsafe.Pointer(&buf[0]), C.int64_t(ra.Len()), 12, off, buf
}

// Index implements storage.Search on the GPU (pc/storage/search.go:13-17).
type Index struct {
	pc.Vec3RandomAccessor // Vec3At / Len / RawIndexAt stay on the host accessor
	h                     *C.pcg_index

	// MinDistSq mirrors KDTree.MinDistSq (kdtree.go:19-22): larger than zero makes Nearest (and the ICP
	// correspondences searched through this index) the approximate search.
	MinDistSq float32
	shared    bool   // a With() copy: the handle belongs to the original
	owner     *Index // ... which the copy keeps reachable: the original's finalizer frees the device index
}

// With is KDTree.With (kdtree.go:58-65): a shallow copy sharing the device index.
func (k *Index) With(minDistSq float32) *Index {
	k2 := *k
	k2.MinDistSq = minDistSq
	k2.shared = true
	k2.owner = k // kdtree.New(ra).With(x) drops the original: without this reference the GC would free k.h under k2
	if k.owner != nil {
		k2.owner = k.owner
	}
	return &k2
}

// DeletePoint is KDTree.DeletePoint (kdtree.go:322-332): the point stops matching any search.
func (k *Index) DeletePoint(pID int) error {
	defer runtime.KeepAlive(k) // the finalizer must not free k.h while C uses it
	defer pin()()
	id := C.int64_t(pID)
	return statusError(C.pcg_index_delete_points(k.h, &id, 1))
}

var _ storage.Search = (*Index)(nil)

// NewIndex is the drop-in for kdtree.New (pc/storage/kdtree/kdtree.go:33).
func NewIndex(ra pc.Vec3RandomAccessor, device int) (*Index, error) {
	defer pin()()
	p, n, stride, off, keep := flatten(ra)
	idx := &Index{Vec3RandomAccessor: ra}
	s := C.pcg_index_build(p, n, stride, &off[0], C.int32_t(device), &idx.h)
	runtime.KeepAlive(keep)
	runtime.KeepAlive(ra)
	if err := statusError(s); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(idx, (*Index).Close)
	return idx, nil
}

func (k *Index) Close() {
	if k.h != nil && !k.shared {
		C.pcg_index_free(k.h)
	}
	k.h = nil
}

// Nearest is KDTree.Nearest (kdtree.go:83-92) as a batch of one.
func (k *Index) Nearest(p mat.Vec3, maxRange float32) storage.Neighbor {
	var out [1]storage.Neighbor
	k.NearestBatch(pc.Vec3Slice{p}, maxRange, out[:])
	return out[0]
}

// Range is KDTree.Range (kdtree.go:148-161), sorted by (DistSq, ID).
func (k *Index) Range(p mat.Vec3, maxRange float32) []storage.Neighbor {
	_, nb := k.RangeBatch(pc.Vec3Slice{p}, maxRange)
	return nb
}

// NearestBatch answers every query in one call. storage.Neighbor{ID int; DistSq float32} has
// the layout of C.pcg_neighbor on 64-bit targets, so `out` is written in place.
func (k *Index) NearestBatch(q pc.Vec3RandomAccessor, maxRange float32, out []storage.Neighbor) {
	defer runtime.KeepAlive(k) // the finalizer must not free k.h while C uses it
	defer pin()()
	p, n, stride, off, keep := flatten(q)
	if n == 0 {
		return
	}
	s := C.pcg_index_nearest_approx(k.h, p, n, stride, &off[0], C.float(maxRange), C.float(k.MinDistSq),
		(*C.pcg_neighbor)(unsafe.Pointer(&out[0])))
	runtime.KeepAlive(keep)
	runtime.KeepAlive(q)
	if s != C.PCG_OK {
		panic(statusError(s)) // Nearest has no error return in the reference
	}
}

// RangeBatch returns CSR offsets (len(q)+1) and the concatenated neighbour lists, using the
// two-call protocol so that the result lives in Go-allocated memory (count, then fill).
func (k *Index) RangeBatch(q pc.Vec3RandomAccessor, maxRange float32) ([]int64, []storage.Neighbor) {
	defer runtime.KeepAlive(k) // the finalizer must not free k.h while C uses it
	defer pin()()
	p, n, stride, off, keep := flatten(q)
	offs := make([]int64, int(n)+1)
	s := C.pcg_index_range_count(k.h, p, n, stride, &off[0], C.float(maxRange), (*C.int64_t)(unsafe.Pointer(&offs[0])))
	if s != C.PCG_OK {
		panic(statusError(s))
	}
	nb := make([]storage.Neighbor, offs[n])
	if len(nb) > 0 {
		s = C.pcg_index_range_fill(k.h, p, n, stride, &off[0], C.float(maxRange),
			(*C.int64_t)(unsafe.Pointer(&offs[0])), (*C.pcg_neighbor)(unsafe.Pointer(&nb[0])))
		if s != C.PCG_OK {
			panic(statusError(s))
		}
	}
	runtime.KeepAlive(keep)
	runtime.KeepAlive(q)
	return offs, nb
}

// voxelGrid implements filter.Filter (pc/filter/filter.go:7-9).
type voxelGrid struct {
	voxelgrid.Options
	device int
}

// NewVoxelGrid is the drop-in for voxelgrid.New (pc/filter/voxelgrid/voxelgrid.go:23-33).
func NewVoxelGrid(leafSize mat.Vec3, opts ...voxelgrid.Option) filter.Filter {
	vg := &voxelGrid{Options: voxelgrid.Options{LeafSize: leafSize}}
	for _, o := range opts {
		o(&vg.Options)
	}
	return vg
}

func xyzOffsets(pp *pc.PointCloud) ([3]C.int64_t, error) {
	var off [3]C.int64_t
	found := 0
	o := 0
	for i, name := range pp.Fields {
		switch name {
		case "xyz":
			return [3]C.int64_t{C.int64_t(o), C.int64_t(o + 4), C.int64_t(o + 8)}, nil
		case "x":
			off[0] = C.int64_t(o)
			found |= 1
		case "y":
			off[1] = C.int64_t(o)
			found |= 2
		case "z":
			off[2] = C.int64_t(o)
			found |= 4
		}
		o += pp.Size[i] * pp.Count[i]
	}
	if found != 7 {
		return off, errors.New("invalid field name") // pc/pointcloud.go:115
	}
	return off, nil
}

// Filter is voxelGrid.Filter (voxelgrid.go:35-134): same records, same order, same bits.
func (f *voxelGrid) Filter(pp *pc.PointCloud) (*pc.PointCloud, error) {
	defer pin()()
	off, err := xyzOffsets(pp)
	if err != nil {
		return nil, err
	}
	stride := pp.Stride()
	out := make([]byte, stride*pp.Points)
	leaf := [3]C.float{C.float(f.LeafSize[0]), C.float(f.LeafSize[1]), C.float(f.LeafSize[2])}
	chunk := [3]C.int64_t{C.int64_t(f.ChunkSize[0]), C.int64_t(f.ChunkSize[1]), C.int64_t(f.ChunkSize[2])}
	var in, outp unsafe.Pointer
	if pp.Points > 0 {
		in, outp = unsafe.Pointer(&pp.Data[0]), unsafe.Pointer(&out[0])
	}
	var n C.int64_t
	s := C.pcg_voxelgrid_filter(in, C.int64_t(pp.Points), C.int64_t(stride), &off[0], &leaf[0], &chunk[0],
		C.int32_t(f.device), outp, &n)
	runtime.KeepAlive(pp)
	if err := statusError(s); err != nil {
		return nil, err
	}
	newPc := &pc.PointCloud{PointCloudHeader: pp.Clone(), Points: int(n), Data: out[:int(n)*stride]}
	newPc.Width, newPc.Height = int(n), 1
	return newPc, nil
}

// NearestPointCorresponder implements icp.PointToPointCorresponder (correspondence.go:14-37).
type NearestPointCorresponder struct{ MaxDist float32 }

func (c *NearestPointCorresponder) Pairs(base storage.Search, target pc.Vec3RandomAccessor) []icp.PointToPointCorrespondence {
	defer pin()()
	idx := base.(*Index)
	defer runtime.KeepAlive(idx)
	p, n, stride, off, keep := flatten(target)
	baseID := make([]int64, int(n)+1)
	targetID := make([]int64, int(n)+1)
	dsq := make([]float32, int(n)+1)
	var m C.int64_t
	s := C.pcg_icp_pairs(idx.h, p, n, stride, &off[0], C.float(c.MaxDist),
		(*C.int64_t)(unsafe.Pointer(&baseID[0])), (*C.int64_t)(unsafe.Pointer(&targetID[0])), (*C.float)(unsafe.Pointer(&dsq[0])), &m)
	runtime.KeepAlive(keep)
	if s != C.PCG_OK {
		panic(statusError(s))
	}
	out := make([]icp.PointToPointCorrespondence, int(m))
	for i := range out {
		out[i] = icp.PointToPointCorrespondence{BaseID: int(baseID[i]), TargetID: int(targetID[i]), SquaredDistance: dsq[i]}
	}
	return out
}

// Mode selects the accumulation order of Evaluate: Strict reproduces the reference's
// sequential float32 sums bit for bit, Fast uses a fixed float64 tree.
type Mode int32

const (
	Strict      Mode = C.PCG_ICP_STRICT
	Fast        Mode = C.PCG_ICP_FAST
	WithHessian Mode = C.PCG_ICP_WITH_HESSIAN // OR-ed in: Evaluate also fills Evaluated.Hessian
)

// WeightFn selects one of the parametric weight functions the device can evaluate (pcg_weight_fn): the reference's
// EvaluateWeightFn (evaluator.go:19-23,72) is a closure, which cannot cross the boundary.
type WeightFn int32

const (
	WeightConstant  WeightFn = C.PCG_WEIGHT_CONSTANT  // DefaultEvaluateWeightFn: w = 1
	WeightTruncated WeightFn = C.PCG_WEIGHT_TRUNCATED // w = dsq < WeightParam ? 1 : 0
	WeightHuber     WeightFn = C.PCG_WEIGHT_HUBER     // w = dsq <= WeightParam ? 1 : sqrt(WeightParam/dsq)
)

// PointToPointEvaluator implements icp.Evaluator (evaluator.go:32-36,69-189).
type PointToPointEvaluator struct {
	Corresponder *NearestPointCorresponder
	MinPairs     int
	Mode         Mode
	WeightFn     WeightFn // zero value = the reference's default weight function
	WeightParam  float32  // a squared distance (threshold / Huber k^2)
}

func (PointToPointEvaluator) HasGradient() bool  { return true }
func (e PointToPointEvaluator) HasHessian() bool { return e.Mode&WithHessian != 0 } // false by default, like evaluator.go:76

func evaluatedFromC(e *C.pcg_evaluated) icp.Evaluated {
	var out icp.Evaluated
	out.Value = float32(e.value)
	for i := 0; i < 6; i++ {
		out.Gradient[i] = float32(e.gradient[i])
	}
	for i := 0; i < 36; i++ { // zero unless WithHessian / GaussNewton was selected
		out.Hessian[i] = float32(e.hessian[i])
	}
	out.DistRMS = float32(e.dist_rms)
	return out
}

func (e *PointToPointEvaluator) Evaluate(base storage.Search, target pc.Vec3RandomAccessor) (*icp.Evaluated, error) {
	defer pin()()
	idx := base.(*Index)
	defer runtime.KeepAlive(idx)
	p, n, stride, off, keep := flatten(target)
	var ev C.pcg_evaluated
	var np C.int64_t
	var prm C.pcg_icp_params
	prm.max_dist = C.float(e.Corresponder.MaxDist)
	prm.min_pairs = C.int32_t(e.MinPairs)
	prm.mode = C.int32_t(e.Mode)
	prm.min_dist_sq = C.float(idx.MinDistSq)
	prm.weight_fn = C.int32_t(e.WeightFn)
	prm.weight_param = C.float(e.WeightParam)
	s := C.pcg_icp_evaluate_params(idx.h, p, n, stride, &off[0], &prm, &ev, &np)
	runtime.KeepAlive(keep)
	if err := statusError(s); err != nil {
		return nil, err
	}
	out := evaluatedFromC(&ev)
	return &out, nil
}

// PointToPointICPGradient.Fit has the signature of icp.PointToPointICPGradient.Fit (icp.go:23);
// the whole loop (correspondence, Evaluate, Update, re-transform) runs on the device.
type PointToPointICPGradient struct {
	Evaluator      *PointToPointEvaluator
	UpdaterFactory *icp.GradientDescentUpdaterFactory
	// GaussNewton replaces the damped gradient step by the solution of the 6x6 normal equations accumulated
	// with Evaluated.Hessian (not in the reference; Weight is then ignored).
	GaussNewton bool
}

func (r *PointToPointICPGradient) Fit(base storage.Search, target pc.Vec3RandomAccessor) (mat.Mat4, icp.Stat, error) {
	defer pin()()
	idx := base.(*Index)
	defer runtime.KeepAlive(idx)
	p, n, stride, off, keep := flatten(target)
	var prm C.pcg_icp_params
	prm.max_dist = C.float(r.Evaluator.Corresponder.MaxDist)
	prm.min_pairs = C.int32_t(r.Evaluator.MinPairs)
	prm.mode = C.int32_t(r.Evaluator.Mode)
	prm.min_dist_sq = C.float(idx.MinDistSq)
	prm.weight_fn = C.int32_t(r.Evaluator.WeightFn)
	prm.weight_param = C.float(r.Evaluator.WeightParam)
	if r.GaussNewton {
		prm.updater = C.PCG_UPDATER_GAUSS_NEWTON
	}
	if f := r.UpdaterFactory; f != nil { // zero values select the reference defaults (updater.go:24-33)
		for i := 0; i < 6; i++ {
			prm.weight[i] = C.float(f.Weight[i])
			prm.threshold[i] = C.float(f.Threshold[i])
		}
		prm.max_iteration = C.int32_t(f.MaxIteration)
	}
	var trans mat.Mat4
	var st C.pcg_icp_stat
	s := C.pcg_icp_fit(idx.h, p, n, stride, &off[0], &prm, (*C.float)(unsafe.Pointer(&trans[0])), &st)
	runtime.KeepAlive(keep)
	stat := icp.Stat{Evaluated: evaluatedFromC(&st.evaluated), NumIteration: int(st.num_iteration)}
	return trans, stat, statusError(s) // on ErrNotEnoughPairs: (trans so far, stat, err) like icp.go:51-53
}

// Replicate copies the built index to another GPU (NVLink peer copy, bit-identical; pcg_index_replicate).
func (k *Index) Replicate(device int) (*Index, error) {
	defer pin()()
	defer runtime.KeepAlive(k)
	r := &Index{Vec3RandomAccessor: k.Vec3RandomAccessor, MinDistSq: k.MinDistSq}
	if err := statusError(C.pcg_index_replicate(k.h, C.int32_t(device), &r.h)); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(r, (*Index).Close)
	return r, nil
}

// FitMulti is Fit with the target split over the GPUs that hold the replicas (bases[0] and its Replicate copies):
// one process, no ranks - each device runs one persistent kernel and the ten partial sums of every iteration
// travel as peer-memory stores over NVLink (pcg_icp_fit_multi).  Fast mode (float64 partial sums).
func (r *PointToPointICPGradient) FitMulti(bases []*Index, target pc.Vec3RandomAccessor) (mat.Mat4, icp.Stat, error) {
	defer pin()()
	p, n, stride, off, keep := flatten(target)
	hs := make([]*C.pcg_index, len(bases))
	for i, b := range bases {
		hs[i] = b.h
	}
	var prm C.pcg_icp_params
	prm.max_dist = C.float(r.Evaluator.Corresponder.MaxDist)
	prm.min_pairs = C.int32_t(r.Evaluator.MinPairs)
	prm.mode = C.PCG_ICP_FAST
	prm.weight_fn = C.int32_t(r.Evaluator.WeightFn)
	prm.weight_param = C.float(r.Evaluator.WeightParam)
	if f := r.UpdaterFactory; f != nil {
		for i := 0; i < 6; i++ {
			prm.weight[i] = C.float(f.Weight[i])
			prm.threshold[i] = C.float(f.Threshold[i])
		}
		prm.max_iteration = C.int32_t(f.MaxIteration)
	}
	var trans mat.Mat4
	var st C.pcg_icp_stat
	// hs holds C pointers only (no Go pointers), so passing &hs[0] obeys the cgo pointer rules
	s := C.pcg_icp_fit_multi(C.int32_t(len(hs)), &hs[0], p, n, stride, &off[0], &prm, (*C.float)(unsafe.Pointer(&trans[0])), &st)
	runtime.KeepAlive(keep)
	runtime.KeepAlive(bases)
	stat := icp.Stat{Evaluated: evaluatedFromC(&st.evaluated), NumIteration: int(st.num_iteration)}
	return trans, stat, statusError(s)
}

// RegionGrowing is the drop-in for regiongrowing.New / Segment (pc/segmentation/regiongrowing/regiongrowing.go:11-56):
// every breadth-first level is one batched Range on the device.
type RegionGrowing struct {
	h *C.pcg_region_growing
	n int
}

// NewRegionGrowing takes the cloud the index was built from and the name of its uint32 property
// (what pp.Uint32Iterator(field) would iterate).
func NewRegionGrowing(search *Index, pp *pc.PointCloud, field string) (*RegionGrowing, error) {
	defer pin()()
	off, err := xyzOffsets(pp)
	if err != nil {
		return nil, err
	}
	labelOff := 0
	found := false
	for i, fn := range pp.Fields {
		if fn == field {
			found = true
			break
		}
		labelOff += pp.Size[i] * pp.Count[i]
	}
	if !found {
		return nil, errors.New("invalid field name") // pointcloud.go:189
	}
	rg := &RegionGrowing{n: pp.Points}
	var data unsafe.Pointer
	if len(pp.Data) > 0 {
		data = unsafe.Pointer(&pp.Data[0])
	}
	s := C.pcg_region_growing_new(search.h, data, C.int64_t(pp.Points), C.int64_t(pp.Stride()), &off[0],
		C.int64_t(labelOff), &rg.h)
	runtime.KeepAlive(pp)
	if err := statusError(s); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(rg, (*RegionGrowing).Close)
	return rg, nil
}

func (r *RegionGrowing) Close() {
	if r.h != nil {
		C.pcg_region_growing_free(r.h)
		r.h = nil
	}
}

// Segment is RegionGrowing.Segment (regiongrowing.go:23-56).
func (r *RegionGrowing) Segment(p mat.Vec3, maxRange float32) []int {
	defer runtime.KeepAlive(r)
	defer pin()()
	out := make([]int, r.n+1) // Go int == int64 on the 64-bit targets this library supports
	var m C.int64_t
	s := C.pcg_region_growing_segment(r.h, (*C.float)(unsafe.Pointer(&p[0])), C.float(maxRange),
		(*C.int64_t)(unsafe.Pointer(&out[0])), C.int64_t(r.n), &m)
	if s != C.PCG_OK {
		panic(statusError(s)) // Segment has no error return in the reference
	}
	return out[:m]
}

// DeviceCloud is a pc.PointCloud whose Data stays in HBM: Unmarshal -> VoxelGrid -> NewIndexFromCloud -> Fit
// without host round trips (pc/io.go:32-45,232-285 for the encodings).
type DeviceCloud struct{ h *C.pcg_cloud }

// Unmarshal is pc.Unmarshal over a byte slice; ascii, binary and binary_compressed.
func Unmarshal(pcd []byte, device int) (*DeviceCloud, error) {
	defer pin()()
	var p unsafe.Pointer
	if len(pcd) > 0 {
		p = unsafe.Pointer(&pcd[0])
	}
	c := &DeviceCloud{}
	if err := statusError(C.pcg_pcd_unmarshal(p, C.int64_t(len(pcd)), C.int32_t(device), &c.h)); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(c, (*DeviceCloud).Close)
	return c, nil
}

// Marshal is pc.Marshal ("DATA binary").
func (c *DeviceCloud) Marshal() ([]byte, error) {
	defer runtime.KeepAlive(c)
	defer pin()()
	var n C.int64_t
	C.pcg_pcd_marshal(c.h, nil, 0, &n)
	out := make([]byte, int(n))
	if n == 0 {
		return out, nil
	}
	return out, statusError(C.pcg_pcd_marshal(c.h, unsafe.Pointer(&out[0]), n, &n))
}

func (c *DeviceCloud) Close() {
	if c.h != nil {
		C.pcg_cloud_free(c.h)
		c.h = nil
	}
}

// VoxelGrid is voxelGrid.Filter on the resident cloud.
func (c *DeviceCloud) VoxelGrid(leaf mat.Vec3, chunk [3]int) (*DeviceCloud, error) {
	defer runtime.KeepAlive(c)
	defer pin()()
	ck := [3]C.int64_t{C.int64_t(chunk[0]), C.int64_t(chunk[1]), C.int64_t(chunk[2])}
	out := &DeviceCloud{}
	if err := statusError(C.pcg_cloud_voxelgrid_filter(c.h, (*C.float)(unsafe.Pointer(&leaf[0])), &ck[0], &out.h)); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(out, (*DeviceCloud).Close)
	return out, nil
}
