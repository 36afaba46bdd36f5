"""Pins the CPU oracle against every golden vector the reference's own tests hold
for the hot path.  Each test names the reference test it replays."""
import numpy as np
import pytest

f32 = np.float32


# ---------------------------------------------------------------- mat ------
def test_translate_elements_and_example(oracle):
    # mat/transform_test.go:9-34
    m = oracle.translate(4, 5, 6)
    assert m.tolist() == [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 4, 5, 6, 1]
    out = oracle.mat4_transform(oracle.translate(1, 2, 3), [[4, 5, 6]])
    assert out.tolist() == [[5, 7, 9]]


def test_mat4_mul_matches_f32_definition(oracle):
    # mat/mat4.go:16-28 : sum over k ascending, every op rounded to float32
    rng = np.random.default_rng(0)
    a = rng.standard_normal(16).astype(f32)
    b = rng.standard_normal(16).astype(f32)
    got = oracle.mat4_mul(a, b)
    exp = np.zeros(16, f32)
    for i in range(4):
        for j in range(4):
            s = f32(0)
            for k in range(4):
                s = f32(s + f32(a[4 * k + i] * b[4 * j + k]))
            exp[4 * j + i] = s
    assert got.tobytes() == exp.tobytes()


def test_rodrigues_vs_rotate(oracle):
    # pc/registration/icp/rodrigues_test.go:9-29 (eps 1e-3; sub-sampled sweep, same float32 stepping)
    eps = 0.001
    vals = []
    v = f32(-1.0)
    while v < 1:
        vals.append(v)
        v = f32(v + f32(0.02))
    vals = vals[::7]
    for vx in vals:
        for vy in vals:
            for vz in vals:
                vec = np.array([vx, vy, vz], f32)
                r = oracle.rodrigues(vec)
                nsq = f32(oracle.norm_sq(vec))
                norm = f32(np.sqrt(np.float64(nsq)))
                inv = f32(f32(1.0) / norm)
                vn = (vec * inv).astype(f32)
                exp = oracle.rotate(vn[0], vn[1], vn[2], norm)
                assert np.all(np.abs(r - exp) <= eps)


# ------------------------------------------------------------- kdtree ------
FIXTURE7 = np.array([[4, 1, 0], [2, 2, 1], [5, 0, 0], [3, 0, 0], [0, 1, 0], [1, 0, 0], [6, 2, 1]], f32)


def test_kdtree_structure(oracle):
    # pc/storage/kdtree/kdtree_test.go:124-155
    k = oracle.Search(FIXTURE7, "kdtree")
    root, ids, dim, left, right = k.dump()
    def node(i):
        if i < 0:
            return None
        return (int(ids[i]), int(dim[i]), node(left[i]), node(right[i]))
    exp = (3, 0,
           (4, 1, (5, 2, None, None), (1, 2, None, None)),
           (0, 1, (2, 2, None, None), (6, 2, None, None)))
    assert node(root) == exp


@pytest.mark.parametrize("pts,depth", [
    (FIXTURE7[:2], 2), (FIXTURE7[:3], 2), (FIXTURE7[:6], 3), (FIXTURE7, 3)])
def test_kdtree_max_depth(oracle, pts, depth):
    # kdtree_test.go:62-117
    assert oracle.Search(pts, "kdtree").max_depth() == depth


NEAREST_CASES = [  # kdtree_test.go:162-229  (p, nodeID, distSq, maxRange)
    ((5, 0, 0), 2, 0.0, 1.0),
    ((5, 0, 0.1), 2, 0.1 * 0.1, 1.0),
    ((4.9, 0.0, 0.0), 2, 0.1 * 0.1, 1.0),
    ((3, 0, 0), 3, 0.0, 1.0),
    ((3, 0, 0.1), 3, 0.1 * 0.1, 1.0),
    ((2.1, 1.9, 1), 1, 2 * 0.1 * 0.1, 1.0),
    ((2.1, 2.1, 1), 1, 2 * 0.1 * 0.1, 1.0),
    ((3.9, 1, 0), 0, 0.1 * 0.1, 1.0),
    ((4.1, 1, 0), 0, 0.1 * 0.1, 1.0),
    ((4.2, 1, 0), -1, 0.1 * 0.1, 0.1),
]


@pytest.mark.parametrize("kind", ["kdtree", "naive"])
@pytest.mark.parametrize("min_dist", [0.0, 0.001])
def test_nearest_golden(oracle, kind, min_dist):
    # kdtree_test.go:157-246 (eps 1e-5 on DistSq, exact ID)
    s = oracle.Search(FIXTURE7, kind, min_dist_sq=float(f32(min_dist) * f32(min_dist)))
    for p, nid, dsq, mr in NEAREST_CASES:
        ids, d = s.nearest([p], mr)
        assert ids[0] == nid, (p, mr)
        assert abs(float(d[0]) - dsq) <= 1e-5


def test_search_leaf_golden(oracle):
    # kdtree_test.go:250-279
    k = oracle.Search(FIXTURE7, "kdtree")
    assert k.search_leaf([5, 0, 0]) == 2
    assert k.search_leaf([2, 2, 1]) == 1
    assert k.search_leaf([4, 1, 0]) == 6


RANGE_FIXTURE = np.array([[0.0, 0.2, 0.0], [3.0, 0.0, 0.0], [0.2, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 5.0],
                          [0.5, 0.0, 0.0], [0.0, 0.0, 0.4]], f32)
RANGE_CASES = [  # kdtree_test.go:313-362
    ((10, 10, 10), 1, []),
    ((0, 0.2, 0), 0.05, [(0, 0.0)]),
    ((0, 0.2, 0), 0.3, [(0, 0.0), (2, 0.08)]),
    ((0, 0.2, 0), 0.45, [(0, 0.0), (2, 0.08), (6, 0.2)]),
    ((0, 0.2, 0), 0.6, [(0, 0.0), (2, 0.08), (6, 0.2), (5, 0.29)]),
]


@pytest.mark.parametrize("kind", ["kdtree", "naive"])
def test_range_golden(oracle, kind):
    s = oracle.Search(RANGE_FIXTURE, kind)
    for p, mr, exp in RANGE_CASES:
        off, ids, d = s.range([p], mr)
        assert off.tolist() == [0, len(exp)]
        assert ids.tolist() == [e[0] for e in exp]
        for got, e in zip(d, exp):
            assert abs(float(got) - e[1]) <= 1e-5


def test_kdtree_equals_naive_random_cloud(oracle):
    # kdtree_test.go:794-834 (Nearest: ID and DistSq bit-equal) and :887-924 (Range, canonical order)
    rng = np.random.default_rng(1234)
    for trial in range(20):
        pts = (rng.random((100, 3), dtype=f32) * f32(10.0)).astype(f32)
        k = oracle.Search(pts, "kdtree")
        nv = oracle.Search(pts, "naive")
        for _ in range(100):
            p = (rng.random(3, dtype=f32) * f32(10.0)).astype(f32)
            mr = float(rng.random(dtype=f32) * f32(10.0))
            a = k.nearest([p], mr)
            b = nv.nearest([p], mr)
            assert a[0][0] == b[0][0] and a[1].tobytes() == b[1].tobytes()
            ra = k.range([p], mr)
            rb = nv.range([p], mr)
            assert ra[0].tolist() == rb[0].tolist()
            assert ra[1].tolist() == rb[1].tolist()
            assert ra[2].tobytes() == rb[2].tobytes()


def test_kdtree_equals_naive_medium(oracle):
    rng = np.random.default_rng(7)
    pts = (rng.random((20000, 3), dtype=f32) * f32(10.0)).astype(f32)
    q = (rng.random((2000, 3), dtype=f32) * f32(12.0) - f32(1.0)).astype(f32)
    k = oracle.Search(pts, "kdtree")
    nv = oracle.Search(pts, "naive")
    for mr in (0.05, 0.3, 20.0):
        a = k.nearest(q, mr)
        b = nv.nearest(q, mr)
        assert np.array_equal(a[0], b[0]) and a[1].tobytes() == b[1].tobytes()


def test_empty_tree(oracle):
    # kdtree.go:84-86, :150-152
    for kind in ("kdtree", "naive"):
        s = oracle.Search(np.zeros((0, 3), f32), kind)
        ids, d = s.nearest([[1, 2, 3]], 0.5)
        assert ids[0] == -1 and d[0] == f32(0.25)
        off, ids, d = s.range([[1, 2, 3]], 0.5)
        assert off.tolist() == [0, 0]


# ------------------------------------------- kdtree: DeletePoint, MinDistSq ------
def _tree(k):
    root, ids, dim, left, right = k.dump()

    def node(i):
        if i < 0:
            return None
        return (int(ids[i]), int(dim[i]), node(left[i]), node(right[i]))

    return node(root)


def _leaf(i, d):
    return (i, d, None, None)


FULL_TREE = (3, 0, (4, 1, _leaf(5, 2), _leaf(1, 2)), (0, 1, _leaf(2, 2), _leaf(6, 2)))
AFTER_DEL_3 = (0, 0, (4, 1, _leaf(5, 2), _leaf(1, 2)), (6, 1, _leaf(2, 2), None))
DELETE_CASES = {  # kdtree_test.go:413-716: (pID, hasError, expected tree) applied in sequence
    "LeafThenNodeWithRightSubTree": [
        (5, False, (3, 0, (4, 1, None, _leaf(1, 2)), (0, 1, _leaf(2, 2), _leaf(6, 2)))),
        (4, False, (3, 0, _leaf(1, 1), (0, 1, _leaf(2, 2), _leaf(6, 2)))),
    ],
    "RootThenNodeWithLeftSubTree": [
        (3, False, AFTER_DEL_3),
        (6, False, (0, 0, (4, 1, _leaf(5, 2), _leaf(1, 2)), _leaf(2, 1))),
    ],
    "NodeWithBothLeftAndRightSubTrees": [
        (0, False, (3, 0, (4, 1, _leaf(5, 2), _leaf(1, 2)), (6, 1, _leaf(2, 2), None))),
    ],
    "TwiceTheSamePoint": [(3, False, AFTER_DEL_3), (3, False, AFTER_DEL_3)],
    "InvalidPointID": [(-1, True, FULL_TREE), (123, True, FULL_TREE)],
}


def test_find_minimum_golden(oracle):
    # kdtree_test.go:388-411
    k = oracle.Search(FIXTURE7, "kdtree")
    assert [k.find_minimum(d) for d in (0, 1, 2)] == [4, 3, 3]
    assert k.find_minimum(3) == -2  # "dim should be <3"


@pytest.mark.parametrize("name", sorted(DELETE_CASES))
def test_delete_point_golden_trees(oracle, name):
    # kdtree_test.go:413-729 (expected tree after every step, error flag)
    k = oracle.Search(FIXTURE7, "kdtree")
    for pid, has_error, exp in DELETE_CASES[name]:
        ok = k.delete_point(pid)
        assert ok == (not has_error)
        assert _tree(k) == exp


def test_delete_all_points_on_a_line(oracle):
    # kdtree_test.go:731-751
    pts = np.array([[4, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0]], f32)
    for kind in ("kdtree", "naive"):
        k = oracle.Search(pts, kind)
        for i in range(len(pts)):
            assert k.delete_point(i)
            ids, _ = k.nearest([pts[i]], 0.001)
            assert ids[0] < 0


def test_delete_point_random_cloud_equals_naive(oracle):
    # kdtree_test.go:864-885 + testNearestRandomCloud :794-834 (incl. findMinimum agreement)
    rng = np.random.default_rng(99)
    for trial in range(10):
        pts = (rng.random((100, 3), dtype=f32) * f32(10.0)).astype(f32)
        k = oracle.Search(pts, "kdtree")
        nv = oracle.Search(pts, "naive")
        for i in rng.permutation(100 // 3):
            assert k.delete_point(int(i)) and nv.delete_point(int(i))
        for d in range(3):
            a, b = k.find_minimum(d), nv.find_minimum(d)
            assert pts[a, d] == pts[b, d]
        for _ in range(100):
            p = (rng.random(3, dtype=f32) * f32(10.0)).astype(f32)
            mr = float(rng.random(dtype=f32) * f32(10.0))
            a, b = k.nearest([p], mr), nv.nearest([p], mr)
            assert a[0][0] == b[0][0] and a[1].tobytes() == b[1].tobytes()
            ra, rb = k.range([p], mr), nv.range([p], mr)
            assert ra[1].tolist() == rb[1].tolist() and ra[2].tobytes() == rb[2].tobytes()


def check_min_dist_contract(pts, q, max_range, min_dist_sq, ids, dsq, exact_ids, exact_dsq, allow_early_miss=False):
    """KDTree.MinDistSq (kdtree.go:19-22): the search may stop at the first candidate closer than
    sqrt(MinDistSq).  Every answer is therefore either the exact nearest neighbour or a real point
    with DistSq < MinDistSq (and inside maxRange); DistSq always belongs to the returned ID."""
    mr2 = f32(max_range) * f32(max_range)
    for i in range(len(q)):
        if ids[i] == exact_ids[i] and dsq[i] == exact_dsq[i]:
            continue
        if ids[i] < 0:
            # reference quirk (kdtree.go:100-106): with maxRange^2 < MinDistSq a first-leaf miss ends the search
            assert allow_early_miss and mr2 < f32(min_dist_sq), i
            assert dsq[i] == mr2
            continue
        d = pts[ids[i]] - q[i]
        want = f32(f32(f32(d[0] * d[0]) + f32(d[1] * d[1])) + f32(d[2] * d[2]))
        assert dsq[i] == want, i
        assert dsq[i] < f32(min_dist_sq) and dsq[i] <= mr2, i
        assert exact_dsq[i] <= dsq[i]


def test_min_dist_sq_contract_on_kdtree(oracle):
    # the property every MinDistSq answer of the reference satisfies; the GPU approximate mode is held to it too
    rng = np.random.default_rng(5)
    pts = (rng.random((5000, 3), dtype=f32) * f32(10.0)).astype(f32)
    q = (rng.random((3000, 3), dtype=f32) * f32(10.0)).astype(f32)
    nv = oracle.Search(pts, "naive")
    for mds in (0.01, 0.05, 0.5):
        k = oracle.Search(pts, "kdtree", min_dist_sq=mds)
        for mr in (0.1, 1.0, 20.0):
            ids, d = k.nearest(q, mr)
            eids, ed = nv.nearest(q, mr)
            check_min_dist_contract(pts, q, mr, mds, ids, d, eids, ed, allow_early_miss=True)
        approx_hits = int(np.sum(k.nearest(q, 20.0)[0] != nv.nearest(q, 20.0)[0]))
        if mds >= 0.05:
            assert approx_hits > 0  # the approximation really kicks in on this cloud


# ------------------------------------------------------ region growing ------
def region_growing_scene(seed=0):
    """regiongrowing_test.go:17-117: a floor and three boxes with labels 0/1/1/2 and +-0.01 noise.
    (Go's math/rand stream is not reproducible here; the expected sets do not depend on the noise.)"""
    def box(w, l, h, res):
        def axis(width):
            out, v = [], f32(-0.5) * f32(width)
            while v <= f32(0.5) * f32(width):
                out.append(v)
                v = f32(v + f32(res))
            return out
        return np.array([[a, b, c] for a in axis(w) for b in axis(l) for c in axis(h)], f32)

    objs = [((0, 0, 0), box(2, 2, 0.01, 0.1), 0), ((0, 0, 0.25), box(0.5, 0.5, 0.5, 0.1), 1),
            ((0, 0.6, 0.4), box(0.3, 0.3, 0.8, 0.1), 1), ((1.5, 0, 0.25), box(0.25, 0.25, 0.5, 0.05), 2)]
    rng = np.random.default_rng(seed)
    pts, labels, indice, cnt = [], [], [], 0
    for pos, p, lab in objs:
        v = (p + np.array(pos, f32) + rng.uniform(-0.01, 0.01, p.shape).astype(f32)).astype(f32)
        pts.append(v)
        labels += [lab] * len(p)
        indice.append(list(range(cnt, cnt + len(p))))
        cnt += len(p)
    return np.concatenate(pts), np.array(labels, np.uint32), indice


REGION_CASES = {  # regiongrowing_test.go:119-155: name -> (p, maxRange, objects whose indices are expected)
    "Label0": ((0.5, 0.1, 0), 0.15, [0]),
    "Label1FirstBox": ((0.25, 0.15, 0.15), 0.15, [1]),
    "Label1SecondBox": ((0, 0.45, 0.4), 0.15, [2]),
    "Label1BothBoxes": ((0, 0.45, 0.4), 0.3, [1, 2]),
    "Label3": ((1.4, 0.125, 0.2), 0.15, [3]),
    "StillOnlyLabel3": ((1.4, 0.125, 0.2), 0.5, [3]),
}


@pytest.mark.parametrize("kind", ["kdtree", "naive"])
def test_region_growing_golden(oracle, kind):
    # regiongrowing_test.go:157-191 (result compared after sort.Ints)
    pts, labels, indice = region_growing_scene()
    s = oracle.Search(pts, kind)
    for name, (p, mr, objs) in REGION_CASES.items():
        got = oracle.region_growing_segment(s, labels, p, mr)
        exp = sorted(i for o in objs for i in indice[o])
        assert sorted(got.tolist()) == exp, name
        assert len(set(got.tolist())) == len(got)
    assert len(oracle.region_growing_segment(s, labels, (10, 10, 10), 0.15)) == 0  # regiongrowing.go:27-29


# ------------------------------------------------------------ PCD I/O ------
def pcd_cases():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "pcd_cases.json")) as f:
        return json.load(f)["cases"]


PCD_ERRORS = {"strconv.ErrSyntax": "PcdSyntaxError", "io.EOF": "PcdEOF", "lzf.ErrDataCorruption": "PcdCorrupt"}


def test_pcd_unmarshal_golden():
    # pc/io_test.go:16-255 (vectors extracted by tools/gen_golden_pcd.py)
    from oracle import pcd

    for name, c in pcd_cases().items():
        raw = bytes.fromhex(c["pcd_hex"])
        if c["err"]:
            with pytest.raises(getattr(pcd, PCD_ERRORS[c["err"]])):
                pcd.unmarshal(raw)
            continue
        h, n, data = pcd.unmarshal(raw)
        assert n == len(c["expected"]) and h.fields == ["x", "y", "z", "xyz", "label"] and h.stride() == 28
        rec = np.frombuffer(data, np.uint8)[: n * 28].reshape(n, 28)
        xyz = rec[:, :12].copy().view("<f4")
        lab = rec[:, 24:28].copy().view("<u4").ravel()
        for i, e in enumerate(c["expected"]):
            assert tuple(xyz[i]) == (f32(e[0]), f32(e[1]), f32(e[2])), name  # Vec3.Equal
            assert lab[i] == e[3], name
    # the three encodings of the reference's fixture agree on x, y, z and label
    a = pcd.unmarshal(bytes.fromhex(pcd_cases()["Ascii"]["pcd_hex"]))[2]
    b = pcd.unmarshal(bytes.fromhex(pcd_cases()["Binary"]["pcd_hex"]))[2]
    assert a == b


def test_pcd_marshal_golden():
    # pc/io_test.go:257-345: default / provided Viewpoint survive Marshal -> Unmarshal; header text of io.go:252-273
    from oracle import pcd

    h = pcd.Header(fields=["x", "y", "z"], size=[4, 4, 4], count=[1, 1, 1], type=["F", "F", "F"], width=3, height=1)
    out = pcd.marshal(h, 3, bytes(36))
    assert out.startswith(b"VERSION 0.0\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 3\nHEIGHT 1\n"
                          b"VIEWPOINT 0.0000 0.0000 0.0000 1.0000 0.0000 0.0000 0.0000\nPOINTS 3\nDATA binary\n")
    h2, n2, d2 = pcd.unmarshal(out)
    assert h2.viewpoint == [0, 0, 0, 1, 0, 0, 0] and n2 == 3 and d2 == bytes(36)
    vp = [1, 2, 3, -0.333, 0.003, 0.895, -0.298]
    h.viewpoint = vp
    h3 = pcd.unmarshal(pcd.marshal(h, 3, bytes(36)))[0]
    assert h3.viewpoint == [float(f32(v)) for v in vp]
    # the Binary golden re-marshals to itself (it was produced by Marshal: 4-decimal viewpoint)
    raw = bytes.fromhex(pcd_cases()["Binary"]["pcd_hex"])
    hb, nb, db = pcd.unmarshal(raw)
    assert pcd.marshal(hb, nb, db) == raw


# ---------------------------------------------------------- voxelgrid ------
def _vg_cloud():
    # pc/filter/voxelgrid/voxelgrid_test.go:59-76 : fields x,y,z,label ; stride 16
    rec = np.zeros(6, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("label", "<u4")])
    pts = [(0.625, 1.875, 0.125, 1), (1.250, 1.250, 1.250, 2), (0.650, 1.875, 0.150, 3), (1.250, 0.000, 1.250, 4),
           (1.250, 1.275, 1.250, 5), (0.000, 3.000, 0.000, 6)]
    for i, p in enumerate(pts):
        rec[i] = p
    return rec


VG_CASES = {  # voxelgrid_test.go:18-54
    "Default": ((0, 0, 0), [(0.0, 3.0, 0.0), (0.6375, 1.875, 0.1375), (1.25, 0.0, 1.25), (1.25, 1.2625, 1.25)],
                [6, 1, 4, 2]),
    "WithChunkSize881": ((8, 8, 1),
                         [(0.0, 3.0, 0.0), (0.6375, 1.875, 0.1375), (1.25, 0.0, 1.25), (1.25, 1.2625, 1.25)],
                         [6, 1, 4, 2]),
    "WithChunkSize333": ((3, 3, 3),
                         [(0.6375, 1.875, 0.1375), (0.0, 3.0, 0.0), (1.25, 0.0, 1.25), (1.25, 1.2625, 1.25)],
                         [1, 6, 4, 2]),
}


@pytest.mark.parametrize("mode", ["dense", "sparse"])
@pytest.mark.parametrize("name", list(VG_CASES))
def test_voxelgrid_golden(oracle, name, mode):
    chunk, exp_pts, exp_labels = VG_CASES[name]
    rec = _vg_cloud()
    rc, out = oracle.voxelgrid_filter(rec.view(np.uint8), 16, (0, 4, 8), (0.125, 0.125, 0.125), chunk, mode=mode)
    assert rc == oracle.OK
    got = out.view(rec.dtype)
    assert len(got) == len(exp_pts)
    for g, e, l in zip(got, exp_pts, exp_labels):
        # exact float equality (Vec3.Equal), expected literals are float32 constants in Go
        assert (g["x"], g["y"], g["z"]) == (f32(e[0]), f32(e[1]), f32(e[2]))
        assert g["label"] == l


def test_voxelgrid_empty_is_no_point(oracle):
    # pc/minmax.go:10-12
    rc, out = oracle.voxelgrid_filter(np.zeros(0, np.uint8), 12, (0, 4, 8), (0.1, 0.1, 0.1))
    assert rc == oracle.E_NO_POINT and len(out) == 0


def test_voxelgrid_negative_min_unchunked_would_panic(oracle):
    # voxelgrid.go:46 passes vMax as size: a cloud with negative vMin indexes past the dense array
    pts = np.array([[-5, -5, -5], [1, 1, 1], [0.9, 0.9, 0.9]], f32)
    rc, _ = oracle.voxelgrid_filter(pts.view(np.uint8), 12, (0, 4, 8), (0.1, 0.1, 0.1))
    assert rc == oracle.E_REF_WOULD_PANIC
    rc, out = oracle.voxelgrid_filter(pts.view(np.uint8), 12, (0, 4, 8), (0.1, 0.1, 0.1), (16, 16, 16))
    assert rc == oracle.OK and len(out) // 12 == 3


@pytest.mark.parametrize("chunk", [(0, 0, 0), (4, 4, 4), (128, 128, 128), (7, 3, 5)])
def test_voxelgrid_sparse_equals_dense(oracle, chunk):
    rng = np.random.default_rng(5)
    for stride, off in ((12, (0, 4, 8)), (20, (4, 8, 12)), (28, (16, 0, 8))):
        n = 5000
        buf = rng.integers(0, 255, size=(n, stride), dtype=np.uint8)
        xyz = (rng.random((n, 3), dtype=f32) * np.array([4.0, 3.0, 1.5], f32)).astype(f32)
        xyz[:, :] -= xyz.min(axis=0)  # min == 0 exactly (needed by the un-chunked reference path)
        for k in range(3):
            buf[:, off[k]:off[k] + 4] = xyz[:, k:k + 1].copy().view(np.uint8)
        a = oracle.voxelgrid_filter(buf, stride, off, (0.1, 0.07, 0.13), chunk, mode="dense")
        b = oracle.voxelgrid_filter(buf, stride, off, (0.1, 0.07, 0.13), chunk, mode="sparse")
        assert a[0] == b[0] == oracle.OK
        assert a[1].tobytes() == b[1].tobytes()
        assert 0 < len(a[1]) // stride < n


# ---------------------------------------------------------------- icp ------
def test_corresponder_golden(oracle):
    # pc/registration/icp/correspondence_test.go:12-37
    base = np.array([[4, 1, 0], [1, 1, 0], [8, 1, 1], [-5, 0, 1], [0, 1, 0]], f32)
    tgt = np.array([[8, 1, 1], [-8, 1, 1], [2, 1, 0]], f32)
    for kind in ("kdtree", "naive"):
        b, t, d = oracle.icp_pairs(oracle.Search(base, kind), tgt, 3.0)
        assert b.tolist() == [2, 1] and t.tolist() == [0, 2] and d.tolist() == [0.0, 1.0]


def test_evaluator_golden(oracle):
    # pc/registration/icp/evaluator_test.go:11-77
    base = np.array([[0, 0, 0], [1, 1, 0], [2, 2, 0], [3, 1, 1], [4, 0, 0]], f32)
    delta = np.array([0.25, 0.125, -0.125], f32)
    target = (base[2:5] + delta).astype(f32)
    kdt = oracle.Search(base, "kdtree")
    rc, ev, npairs = oracle.icp_evaluate(kdt, target, 2.0, 3)
    assert rc == oracle.OK and npairs == 3
    assert ev[0] == f32(oracle.norm_sq(delta))  # exact equality, evaluator_test.go:40-42
    factor = f32(-0.1)
    dR = oracle.rodrigues((ev[4:7] * factor).astype(f32))
    rot = oracle.mat4_transform(dR, target)
    rc, ev2, _ = oracle.icp_evaluate(kdt, rot, 2.0, 3)
    assert rc == oracle.OK and ev2[0] < ev[0]
    tr = (target + (ev[1:4] * factor).astype(f32)).astype(f32)
    rc, ev3, _ = oracle.icp_evaluate(kdt, tr, 2.0, 3)
    assert rc == oracle.OK and ev3[0] < ev[0]


def test_evaluator_not_enough_pairs(oracle):
    # evaluator.go:92-106 : MinPairs 0 -> 6
    base = np.array([[0, 0, 0], [1, 1, 0], [2, 2, 0]], f32)
    rc, _, npairs = oracle.icp_evaluate(oracle.Search(base, "kdtree"), base, 2.0, 0)
    assert rc == oracle.E_NOT_ENOUGH_PAIRS and npairs == 3


def _icp_deltas(oracle):
    T, R, M = oracle.translate, oracle.rotate, oracle.mat4_mul
    return {  # pc/registration/icp/icp_test.go:44-59
        "Trans(0,0,0)": T(0, 0, 0),
        "Trans(0.25,0.125,-0.125)": T(0.25, 0.125, -0.125),
        "Trans(0.5,0.5,1)": T(0.5, 0.5, 1.0),
        "Trans(-0.5,-0.5,0)": T(-0.5, -0.5, 0.0),
        "Rot(1,0,0,0.2)": R(1, 0, 0, 0.2),
        "Rot(1,0,0,-0.2)": R(1, 0, 0, -0.2),
        "Rot(1,0,0,0.1)Trans(0.2,0,0)": M(R(1, 0, 0, 0.1), T(0.2, 0, 0)),
        "Rot(1,0,0,0.1)Trans(-0.2,0,0)": M(R(1, 0, 0, 0.1), T(-0.2, 0, 0)),
        "Trans(0.2,0,0)Rot(1,0,0,0.1)": M(T(0.2, 0, 0), R(1, 0, 0, 0.1)),
        "Trans(-0.2,0,0)Rot(1,0,0,0.1)": M(T(-0.2, 0, 0), R(1, 0, 0, 0.1)),
        "Rot(0,1,0,0.1)Trans(0.2,0,0)": M(R(0, 1, 0, 0.1), T(0.2, 0, 0)),
        "Rot(0,1,0,0.1)Trans(-0.2,0,0)": M(R(0, 1, 0, 0.1), T(-0.2, 0, 0)),
        "Trans(0.2,0,0)Rot(0,1,0,0.1)": M(T(0.2, 0, 0), R(0, 1, 0, 0.1)),
        "Trans(-0.2,0,0)Rot(0,1,0,0.1)": M(T(-0.2, 0, 0), R(0, 1, 0, 0.1)),
    }


@pytest.mark.parametrize("zoff", [0.0, 5.0])
def test_icp_fit_golden(oracle, zoff):
    # pc/registration/icp/icp_test.go:13-98 : MinDistSq=0.01, MaxDist=2, MinPairs=3, residual <= 0.05
    base = np.array([[-2.1, 0, 0], [-1, 1, 0], [0, 2, 0], [1, 1, 1], [2, 0, 0]], f32)
    base[:, 2] += f32(zoff)
    indices = [3, 1, 4, 0, 2]
    for name, delta in _icp_deltas(oracle).items():
        target = oracle.mat4_transform(delta, base[indices])
        kdt = oracle.Search(base, "kdtree", min_dist_sq=0.01)
        rc, trans, ev, iters = oracle.icp_fit(kdt, target, oracle.icp_params(2.0, 3))
        assert rc == oracle.OK, name
        assert 1 <= iters <= 20
        moved = oracle.mat4_transform(trans, target)
        residual = f32(0)
        for i, idx in enumerate(indices):
            residual = f32(residual + f32(oracle.norm_sq((moved[i] - base[idx]).astype(f32))))
        residual = f32(residual / f32(len(indices)))
        assert 0.05 >= residual, (name, residual)


def test_icp_update_defaults_and_flat(oracle):
    # updater.go:24-37 defaults (0.3 / 0.01 / 20), :45-54 flat gradient => unchanged + converged
    prm = oracle.icp_params(1.0)
    ident = oracle.translate(0, 0, 0)
    ev = np.array([0.5, 0.005, -0.005, 0.01, -0.01, 0.0, 0.0, 1.0], f32)
    t, conv = oracle.icp_update(prm, 0, ident, ev)
    assert conv and t.tobytes() == ident.tobytes()
    ev[1] = f32(0.5)
    t, conv = oracle.icp_update(prm, 0, ident, ev)
    assert not conv
    # delta_x = -(1-0/20) * 0.3 * 0.5 ; the rotation delta is within the small-angle branch
    assert t[12] == f32(f32(f32(-1.0) * f32(0.3)) * f32(0.5))
    t, conv = oracle.icp_update(prm, 19, ident, ev)
    assert conv  # i+1 >= MaxIteration
