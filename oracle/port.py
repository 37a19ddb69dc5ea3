"""ctypes binding of oracle/libcvt_oracle.so — the plain-C restatement ("port" oracle) of the
reference's CVT/RVD hot path (oracle/cvt_oracle.c). TEST INFRASTRUCTURE ONLY.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcvt_oracle.so")
_SRC = os.path.join(_HERE, "cvt_oracle.c")
_LIB = None

FLAG_EXHAUSTED, FLAG_TIE, FLAG_OVERFLOW = 1, 2, 4
COUNTER_NAMES = ("pairs", "planes", "plane_vertex", "intersections", "triangles", "nonempty_pairs", "sr_exits", "exhausted")

_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint32)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)
_qp = C.POINTER(C.c_uint64)


def build(force=False):
    """gcc -ffp-contract=off: no FMA contraction, like the reference's own flags."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-frounding-math",
                               _SRC, "-o", _SO, "-lm"])
    return _SO


def _lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            build()
        _LIB = C.CDLL(_SO)
        _LIB.orc_hlbfgs_rosenbrock.restype = C.c_int
    return _LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def facet_adjacency(T):
    T = np.ascontiguousarray(T, dtype=np.uint32)
    adj = np.empty(T.shape, dtype=np.int32)
    _lib().orc_facet_adjacency(C.c_uint32(T.shape[0]), _p(T, _up), _p(adj, _ip))
    return adj


def tet_adjacency(T):
    T = np.ascontiguousarray(T, dtype=np.uint32)
    adj = np.empty(T.shape, dtype=np.int32)
    _lib().orc_tet_adjacency(C.c_uint32(T.shape[0]), _p(T, _up), _p(adj, _ip))
    return adj


def _adjacency(T):
    """facet adjacency for triangles, tet adjacency for tetrahedra (volumetric mode)"""
    return tet_adjacency(T) if T.shape[1] == 4 else facet_adjacency(T)


class _Volumetric:
    """RestrictedVoronoiDiagram::set_volumetric around one oracle call (elements with 4 ids are tetrahedra)."""
    def __init__(self, T):
        self.on = T.shape[1] == 4

    def __enter__(self):
        _lib().orc_set_volumetric(C.c_int(1 if self.on else 0))

    def __exit__(self, *a):
        _lib().orc_set_volumetric(C.c_int(0))


def knn(x, k=20, ksize=None, kstride=None):
    """Neighbour lists with the semantics of Delaunay_NearestNeighbors (delaunay_nn.cpp:73-145)."""
    x = _f64(x)
    S, dim = x.shape
    kstride = kstride or k
    idx = np.empty((S, kstride), dtype=np.uint32)
    cnt = np.empty(S, dtype=np.uint32)
    sqd = np.empty((S, kstride))
    tie = np.zeros(S, dtype=np.uint8)
    ks = None if ksize is None else np.ascontiguousarray(ksize, dtype=np.uint32)
    _lib().orc_knn(C.c_int(dim), C.c_uint32(S), _p(x, _dp), C.c_uint32(k), _p(ks, _up), C.c_uint32(kstride),
                   _p(idx, _up), _p(cnt, _up), _p(sqd, _dp), _p(tie, _bp))
    return idx, cnt, sqd, tie


def nearest(x, q):
    x, q = _f64(x), _f64(q)
    out = np.empty(q.shape[0], dtype=np.uint32)
    _lib().orc_nearest(C.c_int(x.shape[1]), C.c_uint32(x.shape[0]), _p(x, _dp), C.c_uint32(q.shape[0]), _p(q, _dp), _p(out, _up))
    return out


class SurfaceEval:
    """Result of one evaluation (kNN rebuild + compute_centroids / compute_CVT_func_grad)."""
    pass


def surface_eval(V, T, x, mode, check_SR, k=20, kcap=None, ksize=None, weights=None, adj=None, want_pairs=False):
    V, x = _f64(V), _f64(x)
    T = np.ascontiguousarray(T, dtype=np.uint32)
    if adj is None:
        adj = _adjacency(T)
    S, dim = x.shape
    kcap = kcap or (k if not check_SR else max(4 * k, 64))
    r = SurfaceEval()
    r.m = np.zeros(S)
    r.mg = np.zeros((S, dim))
    r.g = np.zeros((S, dim))
    r.f_seed = np.zeros(S)
    f = C.c_double(0.0)
    r.flags = np.zeros(S, dtype=np.uint8)
    cnt = np.zeros(8, dtype=np.uint64)
    w = _f64(weights)
    ks = None if ksize is None else np.ascontiguousarray(ksize, dtype=np.uint32)
    cap = 64 * S if want_pairs else 0
    pairs = np.zeros((max(cap, 1), 2), dtype=np.uint32)
    npairs = C.c_uint64(0)
    with _Volumetric(T):
      rc = _lib().orc_surface_eval(
        C.c_int(dim), C.c_uint32(V.shape[0]), _p(V, _dp), C.c_uint32(T.shape[0]), _p(T, _up), _p(adj, _ip),
        _p(w, _dp), C.c_uint32(S), _p(x, _dp), C.c_uint32(k), C.c_uint32(kcap), _p(ks, _up),
        C.c_int(int(check_SR)), C.c_int(mode), _p(r.m, _dp), _p(r.mg, _dp), C.byref(f), _p(r.g, _dp),
        _p(r.f_seed, _dp), _p(r.flags, _bp), _p(cnt, _qp), _p(pairs, _up) if want_pairs else None,
        C.c_uint64(cap), C.byref(npairs))
    assert rc == 0
    r.f = f.value
    r.counters = dict(zip(COUNTER_NAMES, (int(c) for c in cnt)))
    r.pairs = pairs[:min(npairs.value, cap)] if want_pairs else None
    r.ksize = ks
    return r


def rdt(V, T, x, k=20, kcap=256, adj=None):
    """compute_RDT, simple mode (RVD.cpp:2353-2370), check_SR = true: (n, 3) rows (seed, bisector(0), bisector(1))."""
    V, x = _f64(V), _f64(x)
    T = np.ascontiguousarray(T, dtype=np.uint32)
    if adj is None:
        adj = _adjacency(T)
    S, dim = x.shape
    cap = 8 * S + 64
    tri = np.zeros((cap, 3), dtype=np.uint32)
    n = C.c_uint64(0)
    rc = _lib().orc_rdt(C.c_int(dim), C.c_uint32(V.shape[0]), _p(V, _dp), C.c_uint32(T.shape[0]), _p(T, _up), _p(adj, _ip),
                        C.c_uint32(S), _p(x, _dp), C.c_uint32(k), C.c_uint32(min(kcap, max(S - 1, 1))), _p(tri, _up),
                        C.c_uint64(cap), C.byref(n))
    assert rc == 0 and n.value <= cap
    return tri[:n.value].copy()


def rdt_volume(V, T, x, k=20, kcap=256, adj=None):
    """compute_RDT in volumetric mode (RVD.cpp:2308-2335): (n, 4) rows (seed, v1, v2, v3) in the reference's traversal order,
    reoriented as the reference does; second value: rows whose orientation determinant is below the rounding bound."""
    V, x = _f64(V), _f64(x)
    T = np.ascontiguousarray(T, dtype=np.uint32)
    if adj is None:
        adj = tet_adjacency(T)
    S = x.shape[0]
    cap = 10 * S + 64
    tet = np.zeros((cap, 4), dtype=np.uint32)
    n, unc = C.c_uint64(0), C.c_uint64(0)
    with _Volumetric(T):
        rc = _lib().orc_rdt_volume(C.c_uint32(V.shape[0]), _p(V, _dp), C.c_uint32(T.shape[0]), _p(T, _up), _p(adj, _ip),
                                   C.c_uint32(S), _p(x, _dp), C.c_uint32(k), C.c_uint32(min(kcap, max(S - 1, 1))), _p(tet, _up),
                                   C.c_uint64(cap), C.byref(n), C.byref(unc))
    assert rc == 0 and n.value <= cap
    return tet[:n.value].copy(), int(unc.value)


def rdt_multinerve(V, T, x, use_centroids=True, prefer_seeds=True, locked=None, k=20, kcap=256, adj=None):
    """compute_RDT with RDT_MULTINERVE (| RDT_RVC_CENTROIDS | RDT_PREFER_SEEDS), RVD.cpp:1901-2264, in the reference's own
    order: (triangles [n, 3] of component indices as emitted, vertices [nc, dim], seed of every vertex [nc])."""
    V, x = _f64(V), _f64(x)
    T = np.ascontiguousarray(T, dtype=np.uint32)
    if adj is None:
        adj = _adjacency(T)
    S, dim = x.shape
    tcap, vcap = 12 * S + 64, 2 * S + 64
    tri = np.zeros((tcap, 3), dtype=np.uint32)
    vert = np.zeros((vcap, dim))
    vseed = np.zeros(vcap, dtype=np.uint32)
    nt_, nv_ = C.c_uint64(0), C.c_uint64(0)
    lk = None if locked is None else np.ascontiguousarray(locked, dtype=np.uint8)
    rc = _lib().orc_rdt_multinerve(C.c_int(dim), C.c_uint32(V.shape[0]), _p(V, _dp), C.c_uint32(T.shape[0]), _p(T, _up), _p(adj, _ip),
                                   C.c_uint32(S), _p(x, _dp), C.c_uint32(k), C.c_uint32(min(kcap, max(S - 1, 1))),
                                   C.c_int(int(use_centroids)), C.c_int(int(prefer_seeds)), _p(lk, _bp),
                                   _p(tri, _up), C.c_uint64(tcap), C.byref(nt_), _p(vert, _dp), _p(vseed, _up), C.c_uint64(vcap), C.byref(nv_))
    assert rc == 0 and nt_.value <= tcap and nv_.value <= vcap
    return tri[:nt_.value].copy(), vert[:nv_.value].copy(), vseed[:nv_.value].copy()


def canonical_multinerve(tri, vert, vseed):
    """Order-independent form of a multinerve RDT: components sorted by (seed, coordinates), triangles rotated to their
    smallest vertex, duplicates dropped, rows sorted. The reference numbers components in traversal order."""
    # (seed, coordinates rounded to 1e-6): two sides that differ by rounding noise must order the components of one seed alike
    q = np.round(vert * 1e6)
    order = np.lexsort(tuple(q[:, c] for c in range(vert.shape[1] - 1, -1, -1)) + (vseed,))
    rank = np.empty(len(order), dtype=np.int64)
    rank[order] = np.arange(len(order))
    t = rank[tri.astype(np.int64)] if len(tri) else np.zeros((0, 3), dtype=np.int64)
    if len(t):
        k = np.argmin(t, axis=1)
        t = np.stack([t[np.arange(len(t)), (k + i) % 3] for i in range(3)], 1)
        t = np.unique(t, axis=0)
    return t, vert[order], vseed[order]


def initial_sampling(V, E, S, weights=None):
    """compute_initial_sampling_on_surface / _in_volume (RVD.cpp:1658-1698, mesh_sampling.h): (x [S, dim], element of each
    sample, ok). Elements with 4 ids are tetrahedra."""
    V = _f64(V)
    E = np.ascontiguousarray(E, dtype=np.uint32)
    dim = V.shape[1]
    x = np.zeros((S, dim))
    elem = np.zeros(S, dtype=np.uint32)
    w = None if weights is None else _f64(weights)
    rc = _lib().orc_initial_sampling(C.c_int(dim), C.c_uint32(V.shape[0]), _p(V, _dp), C.c_uint32(E.shape[0]), _p(E, _up),
                                     C.c_int(E.shape[1]), _p(w, _dp), C.c_uint32(S), _p(x, _dp), _p(elem, _up))
    assert rc in (0, 1)
    return x, elem, rc == 0


def lloyd(V, T, x, nb_iter, k=20, locked=None, weights=None, adj=None):
    V = _f64(V)
    x = np.array(x, dtype=np.float64, order="C", copy=True)
    T = np.ascontiguousarray(T, dtype=np.uint32)
    if adj is None:
        adj = _adjacency(T)
    S, dim = x.shape
    flags = np.zeros(S, dtype=np.uint8)
    lk = None if locked is None else np.ascontiguousarray(locked, dtype=np.uint8)
    w = _f64(weights)
    with _Volumetric(T):
      _lib().orc_lloyd(C.c_int(dim), C.c_uint32(V.shape[0]), _p(V, _dp), C.c_uint32(T.shape[0]), _p(T, _up),
                     _p(adj, _ip), _p(w, _dp), C.c_uint32(S), _p(x, _dp), C.c_uint32(k), C.c_uint32(nb_iter),
                     _p(lk, _bp), _p(flags, _bp))
    return x, flags


def newton(V, T, x, nb_iter, m=7, k=20, kcap=256, locked=None, weights=None, adj=None):
    V = _f64(V)
    x = np.array(x, dtype=np.float64, order="C", copy=True)
    T = np.ascontiguousarray(T, dtype=np.uint32)
    if adj is None:
        adj = _adjacency(T)
    S, dim = x.shape
    cap = nb_iter + 8
    fh = np.zeros(cap)
    gh = np.zeros(cap)
    nit, nfev = C.c_uint32(0), C.c_uint32(0)
    flags = np.zeros(S, dtype=np.uint8)
    lk = None if locked is None else np.ascontiguousarray(locked, dtype=np.uint8)
    w = _f64(weights)
    with _Volumetric(T):
      _lib().orc_newton(C.c_int(dim), C.c_uint32(V.shape[0]), _p(V, _dp), C.c_uint32(T.shape[0]), _p(T, _up),
                      _p(adj, _ip), _p(w, _dp), C.c_uint32(S), _p(x, _dp), C.c_uint32(k), C.c_uint32(kcap),
                      C.c_uint32(nb_iter), C.c_uint32(m), _p(lk, _bp), _p(fh, _dp), _p(gh, _dp), C.c_uint32(cap),
                      C.byref(nit), C.byref(nfev), _p(flags, _bp))
    n = min(nit.value, cap)
    return x, dict(f=fh[:n], gnorm=gh[:n], iters=nit.value, nfev=nfev.value, flags=flags)


def hlbfgs_rosenbrock(N=1000, M=5, max_iter=1000):
    """test_HLBFGS known answer: minimum f=0 at x=1 (geogram/src/tests/test_HLBFGS/main.cpp)."""
    x = np.empty(N)
    x[0::2] = -1.2
    x[1::2] = 1.0
    f = C.c_double(0.0)
    nfev = C.c_uint32(0)
    it = _lib().orc_hlbfgs_rosenbrock(C.c_uint32(N), C.c_uint32(M), _p(x, _dp), C.c_uint32(max_iter), C.byref(f), C.byref(nfev))
    return x, f.value, it, nfev.value
