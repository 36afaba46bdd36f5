// Package pcgolgpu is the cgo shim a pcgol maintainer adds to route the data-parallel hot
// path (storage.Search Nearest/Range, filter.VoxelGrid, point-to-point ICP) to
// libpcgol_b200.so.  It exposes the SAME Go types as the reference:
//
//	pcgolgpu.NewIndex(ra)                 implements storage.Search   (replaces kdtree.New)
//	pcgolgpu.NewVoxelGrid(leaf, opts...)  implements filter.Filter    (replaces voxelgrid.New)
//	pcgolgpu.NearestPointCorresponder     implements icp.PointToPointCorresponder
//	pcgolgpu.PointToPointEvaluator        implements icp.Evaluator
//	pcgolgpu.PointToPointICPGradient      same Fit signature as icp.PointToPointICPGradient
//
// SOURCE ONLY: the build image has no Go toolchain, so this package has not been compiled
// here; the C ABI it binds (include/pcgol_b200.h) is exercised by the C++ mirror
// (include/pcgol_b200.hpp, tests/cpp) and the Python ctypes mirror (pcgol_b200/).
//
// Build:  CGO_CFLAGS="-I${REPO}/include" CGO_LDFLAGS="-L${REPO}/pcgol_b200 -lpcgol_b200" go build ./...
package pcgolgpu
