"""ctypes binding of the C-ABI library (include/b200cvt.h, graphitethree_b200/libb200cvt.so).

There is no CPU fallback: importing works anywhere (so that CPU-only checks can verify the
exported symbols), but every compute call needs a CUDA device and raises B200CVTError
otherwise. The library must have been built in-tree by __graft_entry__.build().
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200cvt.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "b200cvt.h")

FLAG_EXHAUSTED, FLAG_TIE, FLAG_POLY_OVERFLOW, FLAG_KMAX = 1, 2, 4, 8
KMAX = 252

_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint32)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)
_fp = C.POINTER(C.c_float)
_qp = C.POINTER(C.c_uint64)

PROGRESS_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_double, C.c_double)
EXCHANGE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p)

_SIGNATURES = {
    "b200cvt_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "b200cvt_destroy": (None, [C.c_void_p]),
    "b200cvt_last_error": (C.c_char_p, []),
    "b200cvt_set_mesh": (C.c_int, [C.c_void_p, _dp, C.c_uint32, C.c_uint32, _up, _ip, C.c_uint32, _dp]),
    "b200cvt_set_seeds": (C.c_int, [C.c_void_p, _dp, C.c_uint32]),
    "b200cvt_initial_sampling": (C.c_int, [C.c_void_p, C.c_uint32, _dp, C.POINTER(C.c_int)]),
    "b200cvt_knn": (C.c_int, [C.c_void_p, C.c_uint32, _up, _up, _dp, _bp]),
    "b200cvt_nearest": (C.c_int, [C.c_void_p, _dp, C.c_uint32, _up]),
    "b200cvt_centroids": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp]),
    "b200cvt_funcgrad": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp]),
    "b200cvt_rdt": (C.c_int, [C.c_void_p, _up, C.c_uint64, C.POINTER(C.c_uint64)]),
    "b200cvt_rdt_multinerve": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _bp, _up, C.c_uint64, C.POINTER(C.c_uint64), _dp, _up,
                                         C.c_uint64, C.POINTER(C.c_uint64)]),
    "b200cvt_get_flags": (C.c_int, [C.c_void_p, _bp]),
    "b200cvt_get_seed_energy": (C.c_int, [C.c_void_p, _dp]),
    "b200cvt_get_stats": (C.c_int, [C.c_void_p, _qp]),
    "b200cvt_lloyd": (C.c_int, [C.c_void_p, C.c_uint32, _bp, _dp, C.c_uint32, PROGRESS_CB, C.c_void_p]),
    "b200cvt_newton": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, _bp, _dp, C.c_uint32, PROGRESS_CB, C.c_void_p, _up]),
    "b200cvt_set_partition": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "b200cvt_set_seeds_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "b200cvt_exchange_chunk_doubles": (C.c_uint64, [C.c_int, C.c_uint32, C.c_uint32]),
    "b200cvt_set_exchange": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, EXCHANGE_CB, C.c_void_p]),
    "b200cvt_set_locked": (C.c_int, [C.c_void_p, _bp, C.c_uint32]),
    "b200cvt_lloyd_device": (C.c_int, [C.c_void_p, C.c_uint32, PROGRESS_CB, C.c_void_p]),
    "b200cvt_newton_device": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, PROGRESS_CB, C.c_void_p, _up]),
    "b200cvt_get_seeds_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200cvt_get_seeds": (C.c_int, [C.c_void_p, _dp]),
    "b200cvt_get_timings": (C.c_int, [C.c_void_p, _fp]),
    "b200cvt_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200cvt_get_cumulative": (C.c_int, [C.c_void_p, _dp, _qp, C.c_int]),
    "b200cvt_measure_peaks": (C.c_int, [C.c_int, _dp, _dp, _dp]),
    "b200cvt_launch_count": (C.c_uint64, [C.c_void_p]),
    "b200cvt_comm_unique_id": (C.c_int, [_bp]),
    "b200cvt_comm_init": (C.c_int, [C.c_void_p, _bp, C.c_uint32, C.c_uint32]),
    "b200cvt_comm_destroy": (C.c_int, [C.c_void_p]),
    "b200cvt_group_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "b200cvt_group_destroy": (None, [C.c_void_p]),
    "b200cvt_group_size": (C.c_uint32, [C.c_void_p]),
    "b200cvt_group_member": (C.c_void_p, [C.c_void_p, C.c_uint32]),
    "b200cvt_group_set_mesh": (C.c_int, [C.c_void_p, _dp, C.c_uint32, C.c_uint32, _up, _ip, C.c_uint32, _dp]),
    "b200cvt_group_lloyd": (C.c_int, [C.c_void_p, C.c_uint32, _bp, _dp, C.c_uint32, PROGRESS_CB, C.c_void_p]),
    "b200cvt_group_newton": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, _bp, _dp, C.c_uint32, PROGRESS_CB, C.c_void_p, _up]),
}

COMM_ID_BYTES = 128

_LIB = None


class B200CVTError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200cvt error %d: %s" % (code, msg))
        self.code = code


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Loads the in-tree CUDA library; fails loudly if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(no CPU fallback exists for this path)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise B200CVTError(rc, lib().b200cvt_last_error().decode("utf-8", "replace"))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Handle:
    """Thin object wrapper over a b200cvt_handle."""

    def __init__(self, dim=3, volumetric=False, device=-1):
        self._h = C.c_void_p()
        self.dim = dim
        self.volumetric = bool(volumetric)
        self.S = 0
        _check(lib().b200cvt_create(device, dim, int(volumetric), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().b200cvt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_mesh(self, vertices, elems, adjacency=None, weights=None):
        V = _f64(vertices)
        E = np.ascontiguousarray(elems, dtype=np.uint32)
        adj = None if adjacency is None else np.ascontiguousarray(adjacency, dtype=np.int32)
        w = None if weights is None else _f64(weights)
        _check(lib().b200cvt_set_mesh(self._h, V.ctypes.data_as(_dp), V.shape[0], V.shape[1], E.ctypes.data_as(_up),
                                      None if adj is None else adj.ctypes.data_as(_ip), E.shape[0],
                                      None if w is None else w.ctypes.data_as(_dp)))

    def set_seeds(self, x):
        x = _f64(x)
        self.S = x.shape[0]
        _check(lib().b200cvt_set_seeds(self._h, x.ctypes.data_as(_dp), self.S))

    def initial_sampling(self, S):
        """compute_initial_sampling (RVD.cpp:1658-1698): (seeds [S, dim], ok); the seeds become the handle's current seeds."""
        x = np.empty((S, self.dim))
        ok = C.c_int(1)
        _check(lib().b200cvt_initial_sampling(self._h, S, x.ctypes.data_as(_dp), C.byref(ok)))
        self.S = S
        return x, bool(ok.value)

    def set_seeds_device(self, ptr, S):
        self.S = S
        _check(lib().b200cvt_set_seeds_device(self._h, C.c_void_p(ptr), S))

    def knn(self, k=20):
        idx = np.empty((self.S, k), dtype=np.uint32)
        cnt = np.empty(self.S, dtype=np.uint32)
        sqd = np.empty((self.S, k))
        fl = np.empty(self.S, dtype=np.uint8)
        _check(lib().b200cvt_knn(self._h, k, idx.ctypes.data_as(_up), cnt.ctypes.data_as(_up), sqd.ctypes.data_as(_dp),
                                 fl.ctypes.data_as(_bp)))
        return idx, cnt, sqd, fl

    def nearest(self, q):
        q = _f64(q)
        out = np.empty(q.shape[0], dtype=np.uint32)
        _check(lib().b200cvt_nearest(self._h, q.ctypes.data_as(_dp), q.shape[0], out.ctypes.data_as(_up)))
        return out

    def centroids(self, check_SR=False, mg=None, m=None):
        mg = np.zeros((self.S, self.dim)) if mg is None else mg
        m = np.zeros(self.S) if m is None else m
        _check(lib().b200cvt_centroids(self._h, int(check_SR), mg.ctypes.data_as(_dp), m.ctypes.data_as(_dp)))
        return mg, m

    def funcgrad(self, check_SR=True, g=None, f0=0.0):
        g = np.zeros((self.S, self.dim)) if g is None else g
        f = C.c_double(f0)
        _check(lib().b200cvt_funcgrad(self._h, int(check_SR), C.byref(f), g.ctypes.data_as(_dp)))
        return f.value, g

    def rdt(self):
        """compute_RDT, simple mode (RVD.cpp:2353-2370): (n, 3) original seed indices, rows sorted; volumetric handles:
        (n, 4) Delaunay tets whose Voronoi vertex lies inside the domain (RVD.cpp:2308-2335)."""
        n = C.c_uint64(0)
        _check(lib().b200cvt_rdt(self._h, None, 0, C.byref(n)))
        tri = np.empty((int(n.value), 4 if self.volumetric else 3), dtype=np.uint32)
        if n.value:
            _check(lib().b200cvt_rdt(self._h, tri.ctypes.data_as(_up), n.value, C.byref(n)))
        return tri

    def rdt_multinerve(self, use_centroids=True, prefer_seeds=True, locked=None):
        """compute_RDT with RDT_MULTINERVE (RVD.cpp:1901-2264): (triangles [n, 3], vertices [nc, dim], seed of each vertex)."""
        lk = None if locked is None else np.ascontiguousarray(locked, dtype=np.uint8)
        lkp = None if lk is None else lk.ctypes.data_as(_bp)
        nt, nv = C.c_uint64(0), C.c_uint64(0)
        _check(lib().b200cvt_rdt_multinerve(self._h, int(use_centroids), int(prefer_seeds), lkp, None, 0, C.byref(nt), None, None, 0, C.byref(nv)))
        tri = np.empty((int(nt.value), 3), dtype=np.uint32)
        vert = np.empty((int(nv.value), self.dim))
        vseed = np.empty(int(nv.value), dtype=np.uint32)
        _check(lib().b200cvt_rdt_multinerve(self._h, int(use_centroids), int(prefer_seeds), lkp, tri.ctypes.data_as(_up), nt.value,
                                            C.byref(nt), vert.ctypes.data_as(_dp), vseed.ctypes.data_as(_up), nv.value, C.byref(nv)))
        return tri, vert, vseed

    def flags(self):
        fl = np.empty(self.S, dtype=np.uint8)
        _check(lib().b200cvt_get_flags(self._h, fl.ctypes.data_as(_bp)))
        return fl

    def seed_energy(self):
        fs = np.empty(self.S)
        _check(lib().b200cvt_get_seed_energy(self._h, fs.ctypes.data_as(_dp)))
        return fs

    def stats(self):
        st = np.zeros(16, dtype=np.uint64)
        _check(lib().b200cvt_get_stats(self._h, st.ctypes.data_as(_qp)))
        d = dict(planes=int(st[0]), cuts=int(st[1]), triangles=int(st[2]), nonempty_pairs=int(st[3]),
                 redo_seeds=int(st[4]), candidate_pairs=int(st[5]), pair_cap=int(st[6]), grid_cells=int(st[7]))
        d["clip_planes"] = int(st[8])
        d["knn_queries"] = int(st[11])                            # seeds served by the last kNN launch (owned + halo when sharded)
        d["subdivision"] = dict(pieces=int(st[9]), pieces_uncertified=int(st[10]), seeds_collected=int(st[12]))
        # volumetric handles reuse the three slots: cells integrated by the cell-first path, cells sent to the (tet, seed) path,
        # bisectors applied to whole cells (vcell.cuh)
        d["volumetric_cells"] = dict(direct=int(st[9]), tet_path=int(st[10]), bisectors=int(st[12]))
        d["facets_skipped_far_from_owned_seeds"] = int(st[13])    # sharded runs
        d["facets_uncertified"] = int(st[14])                     # home list ended inside the distance bound
        d["facets_subdivided"] = int(st[15] & 0xffffffff)         # ... of which covered by small pieces
        d["facets_subdivision_gave_up"] = int(st[15] >> 32)       # ... budgets exceeded: whole-facet grid scan
        return d

    def lloyd(self, x, nb_iter, locked=None, callback=None):
        x = np.array(x, dtype=np.float64, order="C", copy=True)
        self.S = x.shape[0]
        lk = None if locked is None else np.ascontiguousarray(locked, dtype=np.uint8)
        cb = PROGRESS_CB(callback) if callback else PROGRESS_CB()
        _check(lib().b200cvt_lloyd(self._h, nb_iter, None if lk is None else lk.ctypes.data_as(_bp),
                                   x.ctypes.data_as(_dp), self.S, cb, None))
        return x

    def newton(self, x, nb_iter, m=7, locked=None, callback=None):
        x = np.array(x, dtype=np.float64, order="C", copy=True)
        self.S = x.shape[0]
        lk = None if locked is None else np.ascontiguousarray(locked, dtype=np.uint8)
        cb = PROGRESS_CB(callback) if callback else PROGRESS_CB()
        info = np.zeros(4, dtype=np.uint32)
        _check(lib().b200cvt_newton(self._h, nb_iter, m, None if lk is None else lk.ctypes.data_as(_bp),
                                    x.ctypes.data_as(_dp), self.S, cb, None, info.ctypes.data_as(_up)))
        return x, dict(iters=int(info[0]), nfev=int(info[1]), ls_info=int(info[2]))

    def comm_init(self, comm_id, rank, nranks):
        """One rank of an in-library communicator (NCCL + peer mailboxes); collective over the ranks."""
        cid = np.ascontiguousarray(comm_id, dtype=np.uint8)
        assert cid.size == COMM_ID_BYTES
        _check(lib().b200cvt_comm_init(self._h, cid.ctypes.data_as(_bp), rank, nranks))

    def comm_destroy(self):
        _check(lib().b200cvt_comm_destroy(self._h))

    def set_partition(self, rank, nranks):
        _check(lib().b200cvt_set_partition(self._h, rank, nranks))

    def set_exchange(self, slice_ptr, all_ptr, chunk_doubles, callback):
        """callback() -> 0 must all-gather the slice buffer of every rank into the all buffer."""
        self._xcb = EXCHANGE_CB(lambda user: int(callback()))   # keep alive
        _check(lib().b200cvt_set_exchange(self._h, C.c_void_p(slice_ptr), C.c_void_p(all_ptr), chunk_doubles, self._xcb, None))

    def set_locked(self, locked):
        lk = None if locked is None else np.ascontiguousarray(locked, dtype=np.uint8)
        _check(lib().b200cvt_set_locked(self._h, None if lk is None else lk.ctypes.data_as(_bp), self.S))

    def lloyd_device(self, nb_iter, callback=None):
        cb = PROGRESS_CB(callback) if callback else PROGRESS_CB()
        _check(lib().b200cvt_lloyd_device(self._h, nb_iter, cb, None))

    def newton_device(self, nb_iter, m=7, callback=None):
        cb = PROGRESS_CB(callback) if callback else PROGRESS_CB()
        info = np.zeros(4, dtype=np.uint32)
        _check(lib().b200cvt_newton_device(self._h, nb_iter, m, cb, None, info.ctypes.data_as(_up)))
        return dict(iters=int(info[0]), nfev=int(info[1]), ls_info=int(info[2]))

    def get_seeds_device(self, ptr):
        _check(lib().b200cvt_get_seeds_device(self._h, C.c_void_p(ptr)))

    def get_seeds(self):
        x = np.empty((self.S, self.dim))
        _check(lib().b200cvt_get_seeds(self._h, x.ctypes.data_as(_dp)))
        return x

    def timings(self):
        ms = np.zeros(6, dtype=np.float32)
        _check(lib().b200cvt_get_timings(self._h, ms.ctypes.data_as(_fp)))
        return dict(sort=float(ms[0]), knn=float(ms[1]), pairs=float(ms[2]), clip=float(ms[3]), update=float(ms[4]), total=float(ms[5]))

    def set_stream(self, cuda_stream_ptr):
        _check(lib().b200cvt_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def cumulative(self, reset=False):
        ms = np.zeros(6)
        n = C.c_uint64(0)
        _check(lib().b200cvt_get_cumulative(self._h, ms.ctypes.data_as(_dp), C.byref(n), int(reset)))
        return dict(sort=ms[0], knn=ms[1], pairs=ms[2], clip=ms[3], clip_kernel=ms[4], cells=ms[5], evals=int(n.value))

    def launch_count(self):
        return int(lib().b200cvt_launch_count(self._h))


def comm_unique_id():
    """128 bytes that identify a new communicator (rank 0 creates them, every rank passes them to Handle.comm_init)."""
    cid = np.zeros(COMM_ID_BYTES, dtype=np.uint8)
    _check(lib().b200cvt_comm_unique_id(cid.ctypes.data_as(_bp)))
    return cid


class Group:
    """One process, N GPUs: b200cvt_group_* (one host thread per GPU inside the library)."""

    def __init__(self, n_gpus, dim=3, volumetric=False):
        self._g = C.c_void_p()
        self.dim = dim
        _check(lib().b200cvt_group_create(n_gpus, dim, int(volumetric), C.byref(self._g)))
        self.size = int(lib().b200cvt_group_size(self._g))

    def close(self):
        if getattr(self, "_g", None):
            lib().b200cvt_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_mesh(self, vertices, elems, adjacency=None, weights=None):
        V = _f64(vertices)
        E = np.ascontiguousarray(elems, dtype=np.uint32)
        adj = None if adjacency is None else np.ascontiguousarray(adjacency, dtype=np.int32)
        w = None if weights is None else _f64(weights)
        _check(lib().b200cvt_group_set_mesh(self._g, V.ctypes.data_as(_dp), V.shape[0], V.shape[1], E.ctypes.data_as(_up),
                                            None if adj is None else adj.ctypes.data_as(_ip), E.shape[0],
                                            None if w is None else w.ctypes.data_as(_dp)))

    def lloyd(self, x, nb_iter, locked=None, callback=None):
        x = np.array(x, dtype=np.float64, order="C", copy=True)
        lk = None if locked is None else np.ascontiguousarray(locked, dtype=np.uint8)
        cb = PROGRESS_CB(callback) if callback else PROGRESS_CB()
        _check(lib().b200cvt_group_lloyd(self._g, nb_iter, None if lk is None else lk.ctypes.data_as(_bp),
                                         x.ctypes.data_as(_dp), x.shape[0], cb, None))
        return x

    def newton(self, x, nb_iter, m=7, locked=None, callback=None):
        x = np.array(x, dtype=np.float64, order="C", copy=True)
        lk = None if locked is None else np.ascontiguousarray(locked, dtype=np.uint8)
        cb = PROGRESS_CB(callback) if callback else PROGRESS_CB()
        info = np.zeros(4, dtype=np.uint32)
        _check(lib().b200cvt_group_newton(self._g, nb_iter, m, None if lk is None else lk.ctypes.data_as(_bp),
                                          x.ctypes.data_as(_dp), x.shape[0], cb, None, info.ctypes.data_as(_up)))
        return x, dict(iters=int(info[0]), nfev=int(info[1]), ls_info=int(info[2]))


def measure_peaks(device=-1):
    """(fp32 TFLOP/s, fp64 TFLOP/s, copy GB/s) measured with FMA / copy microbenchmarks."""
    a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
    _check(lib().b200cvt_measure_peaks(device, C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value
