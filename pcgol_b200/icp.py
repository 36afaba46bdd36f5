"""pc/registration/icp on the GPU: the same types as the reference package.

  NearestPointCorresponder.pairs      correspondence.go:22-37
  PointToPointEvaluator.evaluate      evaluator.go:91-189
  GradientDescentUpdaterFactory       updater.go:18-37
  PointToPointICPGradient.fit         icp.go:23-67
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .pc import as_vec3_buffer
from .storage import Index

STRICT = _lib.ICP_STRICT
FAST = _lib.ICP_FAST
WITH_HESSIAN = _lib.ICP_WITH_HESSIAN


class ErrNotEnoughPairs(RuntimeError):  # evaluator.go:15-17
    def __init__(self, trans=None, stat=None):
        super().__init__("not enough correspondence pairs")
        self.trans = trans
        self.stat = stat


@dataclass
class Evaluated:  # evaluator.go:25-30
    value: float
    gradient: np.ndarray
    hessian: np.ndarray
    dist_rms: float

    @staticmethod
    def _from_c(e: _lib.Evaluated) -> "Evaluated":
        return Evaluated(np.float32(e.value), np.array(list(e.gradient), np.float32),
                         np.array(list(e.hessian), np.float32), np.float32(e.dist_rms))


@dataclass
class Stat:  # stat.go:3-6
    evaluated: Evaluated
    num_iteration: int
    n_pairs: int = 0


def _off(off):
    return (C.c_int64 * 3)(*[int(o) for o in off])


@dataclass
class NearestPointCorresponder:
    max_dist: float

    def pairs(self, base: Index, target) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(base_id, target_id, squared_distance) of every matched target, in target order."""
        data, n, stride, off = as_vec3_buffer(target)
        b = np.empty(max(n, 1), np.int64)
        t = np.empty(max(n, 1), np.int64)
        d = np.empty(max(n, 1), np.float32)
        m = C.c_int64(0)
        _lib.check(_lib.lib.pcg_icp_pairs_approx(base._h, data.ctypes.data, n, stride, _off(off), self.max_dist,
                                                getattr(base, "min_dist_sq", 0.0), b.ctypes.data, t.ctypes.data,
                                                d.ctypes.data, C.byref(m)))
        return b[: m.value].copy(), t[: m.value].copy(), d[: m.value].copy()


WEIGHT_CONSTANT, WEIGHT_TRUNCATED, WEIGHT_HUBER = _lib.WEIGHT_CONSTANT, _lib.WEIGHT_TRUNCATED, _lib.WEIGHT_HUBER


@dataclass
class PointToPointEvaluator:
    corresponder: NearestPointCorresponder
    min_pairs: int = 0
    mode: int = STRICT
    #: PointToPointEvaluator.WeightFn (evaluator.go:19-23,72): a closure in Go, here one of the parametric family
    #: WEIGHT_CONSTANT (the default, w = 1) | WEIGHT_TRUNCATED (w = dsq < param) | WEIGHT_HUBER (k^2 = param)
    weight_fn: int = WEIGHT_CONSTANT
    weight_param: float = 0.0

    def has_gradient(self) -> bool:
        return True

    def has_hessian(self) -> bool:
        """False like the reference (evaluator.go:76) unless the mode carries WITH_HESSIAN."""
        return bool(self.mode & WITH_HESSIAN)

    def evaluate(self, base: Index, target) -> Evaluated:
        data, n, stride, off = as_vec3_buffer(target)
        ev = _lib.Evaluated()
        npairs = C.c_int64(0)
        p = _params(self, None, base)
        rc = _lib.lib.pcg_icp_evaluate_params(base._h, data.ctypes.data, n, stride, _off(off), C.byref(p),
                                              C.byref(ev), C.byref(npairs))
        if rc == _lib.E_NOT_ENOUGH_PAIRS:
            raise ErrNotEnoughPairs()
        _lib.check(rc)
        return Evaluated._from_c(ev)


@dataclass
class GradientDescentUpdaterFactory:
    weight: Sequence[float] = (0,) * 6
    threshold: Sequence[float] = (0,) * 6
    max_iteration: int = 0
    kind = _lib.UPDATER_GRADIENT_DESCENT


@dataclass
class GaussNewtonUpdaterFactory:
    """Not in the reference: consumes Evaluated.Hessian (the hook evaluator.go:25-36 declares) and
    solves the 6x6 normal equations per iteration; same convergence test and MaxIteration cap."""
    threshold: Sequence[float] = (0,) * 6
    max_iteration: int = 0
    weight: Sequence[float] = (0,) * 6  # ignored
    kind = _lib.UPDATER_GAUSS_NEWTON


def _params(evaluator: PointToPointEvaluator, uf, base: Optional[Index] = None) -> _lib.IcpParams:
    uf = uf or GradientDescentUpdaterFactory()
    p = _lib.IcpParams()
    p.max_dist = evaluator.corresponder.max_dist
    p.min_pairs = evaluator.min_pairs
    for k in range(6):
        p.weight[k] = uf.weight[k]
        p.threshold[k] = uf.threshold[k]
    p.max_iteration = uf.max_iteration
    p.mode = evaluator.mode
    p.updater = uf.kind
    p.weight_fn = evaluator.weight_fn
    p.weight_param = evaluator.weight_param
    # the base search's own MinDistSq applies to the correspondences, as with a *kdtree.KDTree in the reference
    p.min_dist_sq = getattr(base, "min_dist_sq", 0.0) if base is not None else 0.0
    return p


@dataclass
class PointToPointICPGradient:
    evaluator: PointToPointEvaluator
    updater_factory: Optional[object] = None  # GradientDescentUpdaterFactory | GaussNewtonUpdaterFactory

    def params(self, base: Optional[Index] = None) -> _lib.IcpParams:
        return _params(self.evaluator, self.updater_factory, base)

    def _finish(self, rc, trans, stat):
        st = Stat(Evaluated._from_c(stat.evaluated), int(stat.num_iteration), int(stat.n_pairs))
        if rc == _lib.E_NOT_ENOUGH_PAIRS:  # icp.go:51-53: (trans so far, stat, err)
            raise ErrNotEnoughPairs(trans, st)
        _lib.check(rc)
        return trans, st

    def fit(self, base: Index, target) -> Tuple[np.ndarray, Stat]:
        """Returns (trans float32[16] column-major, Stat); raises ErrNotEnoughPairs like the reference."""
        data, n, stride, off = as_vec3_buffer(target)
        p = self.params(base)
        trans = np.zeros(16, np.float32)
        stat = _lib.IcpStat()
        rc = _lib.lib.pcg_icp_fit(base._h, data.ctypes.data, n, stride, _off(off), C.byref(p), trans.ctypes.data,
                                  C.byref(stat))
        return self._finish(rc, trans, stat)

    def fit_dev(self, base: Index, d_target: int, n: int, stream: int = 0, stride: int = 12, off=(0, 4, 8)):
        p = self.params(base)
        trans = np.zeros(16, np.float32)
        stat = _lib.IcpStat()
        rc = _lib.lib.pcg_icp_fit_dev(base._h, d_target, n, stride, _off(off), C.byref(p), trans.ctypes.data,
                                      C.byref(stat), stream)
        return self._finish(rc, trans, stat)

    def fit_multi(self, bases: Sequence[Index], target) -> Tuple[np.ndarray, Stat]:
        """Fit with the target split over the devices of `bases` (replicas of one index, see Index.replicate):
        one process, NVLink peer exchange of the partial sums inside the per-device kernels.  Fast mode."""
        data, n, stride, off = as_vec3_buffer(target)
        p = self.params(bases[0])
        k = len(bases)
        hs = (C.c_void_p * k)(*[b._h for b in bases])
        trans = np.zeros(16, np.float32)
        stat = _lib.IcpStat()
        rc = _lib.lib.pcg_icp_fit_multi(k, hs, data.ctypes.data, n, stride, _off(off), C.byref(p), trans.ctypes.data,
                                        C.byref(stat))
        return self._finish(rc, trans, stat)

    def fit_multi_dev(self, bases: Sequence[Index], d_targets: Sequence[int], n_targets: Sequence[int],
                      stride: int = 12, off=(0, 4, 8)) -> Tuple[np.ndarray, Stat]:
        """Device-resident variant: d_targets[r] (n_targets[r] records) lives on the device of bases[r]."""
        p = self.params(bases[0])
        k = len(bases)
        hs = (C.c_void_p * k)(*[b._h for b in bases])
        pt = (C.c_void_p * k)(*d_targets)
        nt = (C.c_int64 * k)(*n_targets)
        trans = np.zeros(16, np.float32)
        stat = _lib.IcpStat()
        rc = _lib.lib.pcg_icp_fit_multi_dev(k, hs, pt, nt, stride, _off(off), C.byref(p), trans.ctypes.data,
                                            C.byref(stat))
        return self._finish(rc, trans, stat)

    def fit_pairs_dev(self, d_base, n_base, d_target, n_target, device: int = 0, stream: int = 0, stride: int = 12,
                      off=(0, 4, 8)):
        """Scan-pair farm: lists of device pointers / sizes. Returns (trans (k,16), num_iteration, status)."""
        k = len(d_base)
        p = self.params()
        pb = (C.c_void_p * k)(*d_base)
        pt = (C.c_void_p * k)(*d_target)
        nb = (C.c_int64 * k)(*n_base)
        nt = (C.c_int64 * k)(*n_target)
        trans = np.zeros((k, 16), np.float32)
        stats = (_lib.IcpStat * k)()
        status = np.zeros(k, np.int32)
        _lib.check(_lib.lib.pcg_icp_fit_pairs_dev(k, pb, nb, pt, nt, stride, _off(off), C.byref(p), device,
                                                 trans.ctypes.data, C.cast(stats, C.c_void_p), status.ctypes.data,
                                                 stream))
        iters = np.array([s.num_iteration for s in stats], np.int32)
        return trans, iters, status, stats
