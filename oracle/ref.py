"""ctypes binding of oracle/_ref/libref_driver.so (the UNMODIFIED reference built by
oracle/Makefile.ref from /root/reference). TEST INFRASTRUCTURE ONLY.

Each method names the reference function it drives; see oracle/ref_driver.cpp.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint32)
_ullp = C.POINTER(C.c_ulonglong)


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_driver.so"))


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_ref", "libref_driver.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref not built: run `make -f oracle/Makefile.ref -j8`")
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, C.c_uint, _dp, C.c_uint, _up, C.c_int, _dp, C.c_int, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_set_points.argtypes = [C.c_void_p, C.c_uint, _dp]
        L.ref_get_points.argtypes = [C.c_void_p, _dp]
        L.ref_initial_sampling.argtypes = [C.c_void_p, C.c_uint]
        L.ref_lock_point.argtypes = [C.c_void_p, C.c_uint]
        for n in ("ref_lloyd",):
            getattr(L, n).restype = C.c_double
            getattr(L, n).argtypes = [C.c_void_p, C.c_uint]
        L.ref_newton.restype = C.c_double
        L.ref_newton.argtypes = [C.c_void_p, C.c_uint, C.c_uint]
        L.ref_update_delaunay.restype = C.c_double
        L.ref_update_delaunay.argtypes = [C.c_void_p]
        L.ref_get_neighbors.argtypes = [C.c_void_p, C.c_uint, _up, _up]
        L.ref_nearest_vertex.restype = C.c_uint
        L.ref_nearest_vertex.argtypes = [C.c_void_p, _dp]
        L.ref_centroids.restype = C.c_double
        L.ref_centroids.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.ref_funcgrad.restype = C.c_double
        L.ref_funcgrad.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.ref_polygons.restype = C.c_ulonglong
        L.ref_polygons.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _ullp, _up, C.c_ulonglong]
        L.ref_rdt.restype = C.c_uint
        L.ref_rdt.argtypes = [C.c_void_p, C.c_int, _up, C.c_uint, _dp, C.c_uint, _up]
        L.ref_nb_threads.restype = C.c_int
        L.ref_counters.argtypes = [C.c_void_p, _up, _up]
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _u(a):
    return a.ctypes.data_as(_up)


class RefCVT:
    """GEO::CentroidalVoronoiTesselation over a procedural mesh (CVT.cpp:56-75)."""

    def __init__(self, vertices, elems, volumetric=False, weights=None, multithread=False, max_threads=0):
        L = _lib()
        self.V = np.ascontiguousarray(vertices, dtype=np.float64)
        self.E = np.ascontiguousarray(elems, dtype=np.uint32)
        self.dim = self.V.shape[1]
        self.volumetric = bool(volumetric)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self.h = L.ref_create(self.dim, self.V.shape[0], _d(self.V), self.E.shape[0], _u(self.E),
                              int(volumetric), None if w is None else _d(w), int(multithread), int(max_threads))
        self.S = 0

    def close(self):
        if self.h:
            _lib().ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def nb_threads():
        return _lib().ref_nb_threads()

    def set_points(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.S = x.shape[0]
        _lib().ref_set_points(self.h, self.S, _d(x))

    def points(self):
        x = np.empty((self.S, self.dim))
        _lib().ref_get_points(self.h, _d(x))
        return x

    def initial_sampling(self, S):
        self.S = S
        return _lib().ref_initial_sampling(self.h, S)

    def lock_point(self, i):
        _lib().ref_lock_point(self.h, i)

    def lloyd(self, n):
        return _lib().ref_lloyd(self.h, n)

    def newton(self, n, m=7):
        return _lib().ref_newton(self.h, n, m)

    def counters(self):
        a, b = C.c_uint32(0), C.c_uint32(0)
        _lib().ref_counters(self.h, C.byref(a), C.byref(b))
        return dict(funcgrad=a.value, newiteration=b.value)

    def update_delaunay(self):
        return _lib().ref_update_delaunay(self.h)

    def neighbors(self, kmax=64):
        idx = np.empty((self.S, kmax), dtype=np.uint32)
        cnt = np.empty(self.S, dtype=np.uint32)
        _lib().ref_get_neighbors(self.h, kmax, _u(idx), _u(cnt))
        return idx, cnt

    def nearest_vertex(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        return _lib().ref_nearest_vertex(self.h, _d(p))

    def centroids(self, check_SR=False):
        mg = np.zeros((self.S, self.dim))
        m = np.zeros(self.S)
        t = _lib().ref_centroids(self.h, int(check_SR), _d(mg), _d(m))
        return mg, m, t

    def funcgrad(self, check_SR=True):
        f = C.c_double(0.0)
        g = np.zeros((self.S, self.dim))
        t = _lib().ref_funcgrad(self.h, int(check_SR), C.byref(f), _d(g))
        return f.value, g, t

    def polygons(self, check_SR=False, want_pairs=False, pairs_cap=0):
        fs = np.zeros(self.S)
        ms = np.zeros(self.S)
        cnt = np.zeros(4, dtype=np.uint64)
        pairs = np.zeros((max(pairs_cap, 1), 2), dtype=np.uint32) if want_pairs else None
        n = _lib().ref_polygons(self.h, int(check_SR), _d(fs), _d(ms), cnt.ctypes.data_as(_ullp),
                                None if pairs is None else _u(pairs), pairs_cap)
        return fs, ms, cnt, (None if pairs is None else pairs[:min(n, pairs_cap)])

    def rdt(self, mode=0, cap=None):
        cap = cap or (8 * self.S + 64)
        per = 4 if self.volumetric else 3
        tri = np.empty((cap, per), dtype=np.uint32)
        vtx = np.empty((cap, self.dim))
        nv = C.c_uint(0)
        n = _lib().ref_rdt(self.h, mode, _u(tri), cap, _d(vtx), cap, C.byref(nv))
        return tri[:n].copy(), vtx[:nv.value].copy()
