"""Procedural input meshes and seed sets for the CVT/RVD hot path.

All meshes are (vertices float64 [nv, D], elements uint32 [ne, 3|4]). Numpy only; used
by tests, bench.py and the smoke test to build the same synthetic workloads on the CPU
oracle and the CUDA path.

Reference counterparts: Graphite's create_sphere (icosahedron + 1->4 splits,
OGF/mesh/commands/mesh_grob_shapes_commands.cpp:254-275), geogram's
set_anisotropy (geogram/mesh/mesh_geometry.cpp:144-164) and area-weighted random
sampling (geogram/mesh/mesh_sampling.h:119-199).
"""
import numpy as np

_T = (1.0 + 5.0 ** 0.5) / 2.0
_ICO_V = np.array([
    [-1, _T, 0], [1, _T, 0], [-1, -_T, 0], [1, -_T, 0],
    [0, -1, _T], [0, 1, _T], [0, -1, -_T], [0, 1, -_T],
    [_T, 0, -1], [_T, 0, 1], [-_T, 0, -1], [-_T, 0, 1]], dtype=np.float64)
_ICO_F = np.array([
    [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
    [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
    [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
    [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)


def _weld(V, F, decimals=12):
    """Merges duplicated vertices (exact after rounding) and drops unused ones."""
    key = np.round(V, decimals) + 0.0
    _, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    return V[first], inv[F].astype(np.uint32)


def icosphere(frequency, radius=1.0):
    """Class-I geodesic sphere: 20*frequency^2 triangles (C1: frequency 45 -> 40 500)."""
    n = int(frequency)
    base = _ICO_V / np.linalg.norm(_ICO_V[0])
    Vs, Fs, off = [], [], 0
    # barycentric lattice of one face
    ij = [(i, j) for i in range(n + 1) for j in range(n + 1 - i)]
    index = {p: k for k, p in enumerate(ij)}
    lat = np.array(ij, dtype=np.float64) / n
    tris = []
    for i in range(n):
        for j in range(n - i):
            tris.append((index[(i, j)], index[(i + 1, j)], index[(i, j + 1)]))
            if i + j < n - 1:
                tris.append((index[(i + 1, j)], index[(i + 1, j + 1)], index[(i, j + 1)]))
    tris = np.array(tris, dtype=np.int64)
    for f in _ICO_F:
        a, b, c = base[f[0]], base[f[1]], base[f[2]]
        P = a[None, :] * (1.0 - lat[:, :1] - lat[:, 1:2]) + b[None, :] * lat[:, :1] + c[None, :] * lat[:, 1:2]
        Vs.append(P)
        Fs.append(tris + off)
        off += P.shape[0]
    V = np.concatenate(Vs)
    F = np.concatenate(Fs)
    V, F = _weld(V, F, 9)
    V = V / np.linalg.norm(V, axis=1, keepdims=True) * radius
    return np.ascontiguousarray(V), np.ascontiguousarray(F)


def icosphere_split(p, radius=1.0):
    """Graphite create_sphere: icosahedron + p 1->4 splits, 20*4^p triangles."""
    V = _ICO_V / np.linalg.norm(_ICO_V[0])
    F = _ICO_F.copy()
    for _ in range(p):
        e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
        es = np.sort(e, axis=1)
        ue, inv = np.unique(es, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        mid = 0.5 * (V[ue[:, 0]] + V[ue[:, 1]])
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        nv = V.shape[0]
        V = np.concatenate([V, mid])
        nf = F.shape[0]
        m01, m12, m20 = inv[:nf] + nv, inv[nf:2 * nf] + nv, inv[2 * nf:] + nv
        F = np.concatenate([
            np.stack([F[:, 0], m01, m20], 1), np.stack([F[:, 1], m12, m01], 1),
            np.stack([F[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)])
    return np.ascontiguousarray(V * radius), np.ascontiguousarray(F.astype(np.uint32))


def _value_noise(P, seed=42, octaves=3, base_freq=2.0):
    """3-octave trilinear value noise on a hashed integer lattice, in [-1, 1]."""
    def h(ix, iy, iz, s):
        x = (ix.astype(np.int64) * 73856093) ^ (iy.astype(np.int64) * 19349663) ^ (iz.astype(np.int64) * 83492791) ^ (s * 2654435761)
        x = (x ^ (x >> 13)) * 1274126177
        x = x ^ (x >> 16)
        return ((x & 0xffffff).astype(np.float64) / float(0xffffff)) * 2.0 - 1.0
    out = np.zeros(P.shape[0])
    amp, freq, tot = 1.0, base_freq, 0.0
    for o in range(octaves):
        Q = P * freq
        I = np.floor(Q)
        t = Q - I
        t = t * t * (3.0 - 2.0 * t)
        ix, iy, iz = I[:, 0], I[:, 1], I[:, 2]
        acc = 0.0
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    w = (t[:, 0] if dx else 1 - t[:, 0]) * (t[:, 1] if dy else 1 - t[:, 1]) * (t[:, 2] if dz else 1 - t[:, 2])
                    acc = acc + w * h(ix + dx, iy + dy, iz + dz, seed + o)
        out += amp * acc
        tot += amp
        amp *= 0.5
        freq *= 2.0
    return out / tot


def noise_sphere(frequency, amplitude=0.1, seed=42):
    """C2: icosphere displaced radially by 3-octave value noise, amplitude 0.1."""
    V, F = icosphere(frequency)
    r = 1.0 + amplitude * _value_noise(V, seed)
    return np.ascontiguousarray(V * r[:, None]), F


def trefoil_tube(nu, nv, tube_radius=0.4):
    """C3: tube around the trefoil (sin t + 2 sin 2t, cos t - 2 cos 2t, -sin 3t); 2*nu*nv triangles."""
    t = np.linspace(0.0, 2.0 * np.pi, nu, endpoint=False)
    c = np.stack([np.sin(t) + 2 * np.sin(2 * t), np.cos(t) - 2 * np.cos(2 * t), -np.sin(3 * t)], 1)
    d = np.stack([np.cos(t) + 4 * np.cos(2 * t), -np.sin(t) + 4 * np.sin(2 * t), -3 * np.cos(3 * t)], 1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    dd = np.stack([-np.sin(t) - 8 * np.sin(2 * t), -np.cos(t) + 8 * np.cos(2 * t), 9 * np.sin(3 * t)], 1)
    nrm = dd - (dd * d).sum(1, keepdims=True) * d
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    bn = np.cross(d, nrm)
    a = np.linspace(0.0, 2.0 * np.pi, nv, endpoint=False)
    V = (c[:, None, :] + tube_radius * (np.cos(a)[None, :, None] * nrm[:, None, :] + np.sin(a)[None, :, None] * bn[:, None, :])).reshape(-1, 3)
    i = np.arange(nu)[:, None]
    j = np.arange(nv)[None, :]
    v00 = (i * nv + j)
    v10 = (((i + 1) % nu) * nv + j)
    v01 = (i * nv + (j + 1) % nv)
    v11 = (((i + 1) % nu) * nv + (j + 1) % nv)
    F = np.concatenate([np.stack([v00, v10, v11], -1).reshape(-1, 3), np.stack([v00, v11, v01], -1).reshape(-1, 3)])
    return np.ascontiguousarray(V), np.ascontiguousarray(F.astype(np.uint32))


def box_surface(n, size=(1.0, 1.0, 1.0)):
    """Closed box surface, n x n quads per side (12 n^2 triangles): sharp edges, 'CAD-like'."""
    sx, sy, sz = size
    Vs, Fs, off = [], [], 0
    u = np.linspace(0.0, 1.0, n + 1)
    U, W = np.meshgrid(u, u, indexing="ij")
    U, W = U.reshape(-1), W.reshape(-1)
    i = np.arange(n)[:, None]
    j = np.arange(n)[None, :]
    q00 = (i * (n + 1) + j).reshape(-1)
    q10 = ((i + 1) * (n + 1) + j).reshape(-1)
    q01 = (i * (n + 1) + j + 1).reshape(-1)
    q11 = ((i + 1) * (n + 1) + j + 1).reshape(-1)
    for axis in range(3):
        for side in (0, 1):
            P = np.zeros((U.shape[0], 3))
            a1, a2 = (axis + 1) % 3, (axis + 2) % 3
            P[:, axis] = side
            P[:, a1] = U
            P[:, a2] = W
            if side:
                T = np.concatenate([np.stack([q00, q10, q11], 1), np.stack([q00, q11, q01], 1)])
            else:
                T = np.concatenate([np.stack([q00, q11, q10], 1), np.stack([q00, q01, q11], 1)])
            Vs.append(P)
            Fs.append(T + off)
            off += P.shape[0]
    V, F = _weld(np.concatenate(Vs), np.concatenate(Fs), 12)
    return np.ascontiguousarray(V * np.array([sx, sy, sz])), np.ascontiguousarray(F)


def cad_like(n):
    """C4: L-shaped bracket made of axis-aligned faces (sharp creases); ~ 28 n^2 triangles.

    Built as the boundary of the voxel union {[0,2]x[0,1]x[0,1]} U {[0,1]x[1,2]x[0,1]} at
    resolution n per unit."""
    occ = np.zeros((2 * n, 2 * n, n), dtype=bool)
    occ[:, :n, :] = True
    occ[:n, n:, :] = True
    return voxel_boundary(occ, 1.0 / n)


def voxel_boundary(occ, h):
    """Triangulated boundary of a voxel set, outward orientation."""
    quads = []
    pad = np.pad(occ, 1)
    X, Y, Z = np.nonzero(occ)
    corner = {
        (0, 0): [(0, 0, 0), (0, 0, 1), (0, 1, 1), (0, 1, 0)], (0, 1): [(1, 0, 0), (1, 1, 0), (1, 1, 1), (1, 0, 1)],
        (1, 0): [(0, 0, 0), (1, 0, 0), (1, 0, 1), (0, 0, 1)], (1, 1): [(0, 1, 0), (0, 1, 1), (1, 1, 1), (1, 1, 0)],
        (2, 0): [(0, 0, 0), (0, 1, 0), (1, 1, 0), (1, 0, 0)], (2, 1): [(0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]}
    for axis in range(3):
        for side in (0, 1):
            d = [0, 0, 0]
            d[axis] = 1 if side else -1
            nb = pad[X + 1 + d[0], Y + 1 + d[1], Z + 1 + d[2]]
            sel = ~nb
            base = np.stack([X[sel], Y[sel], Z[sel]], 1)
            q = np.stack([base + np.array(c) for c in corner[(axis, side)]], 1)
            quads.append(q)
    Q = np.concatenate(quads)  # [nq, 4, 3] integer lattice corners
    nq = Q.shape[0]
    P = Q.reshape(-1, 3)
    dims = np.array(occ.shape) + 1
    key = (P[:, 0] * dims[1] + P[:, 1]) * dims[2] + P[:, 2]
    uk, inv = np.unique(key, return_inverse=True)
    inv = inv.reshape(nq, 4)
    V = np.stack([uk // (dims[1] * dims[2]), (uk // dims[2]) % dims[1], uk % dims[2]], 1).astype(np.float64) * h
    F = np.concatenate([inv[:, [0, 1, 2]], inv[:, [0, 2, 3]]]).astype(np.uint32)
    return np.ascontiguousarray(V), np.ascontiguousarray(F)


def kuhn_cube(n):
    """C5: unit cube split into n^3 cells x 6 Kuhn tetrahedra (positively oriented)."""
    g = np.arange(n + 1)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    V = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float64) / n

    def vid(i, j, k):
        return (i * (n + 1) + j) * (n + 1) + k
    c = np.arange(n)
    I, J, K = np.meshgrid(c, c, c, indexing="ij")
    I, J, K = I.reshape(-1), J.reshape(-1), K.reshape(-1)
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    tets = []
    for p in perms:
        o = np.stack([I, J, K], 1)
        steps = [o.copy()]
        cur = o.copy()
        for a in p:
            cur = cur.copy()
            cur[:, a] += 1
            steps.append(cur)
        ids = [vid(s[:, 0], s[:, 1], s[:, 2]) for s in steps]
        T = np.stack(ids, 1)
        # orientation: sign of permutation decides; fix by swapping two vertices when negative
        sign = np.linalg.det(np.eye(3)[list(p)])
        if sign < 0:
            T = T[:, [0, 1, 3, 2]]
        tets.append(T)
    return np.ascontiguousarray(V), np.ascontiguousarray(np.concatenate(tets).astype(np.uint32))


def vertex_normals(V, F):
    """Area-weighted vertex normals (geogram compute_normals, mesh_geometry.cpp:66-105)."""
    P = V[:, :3]
    n = np.cross(P[F[:, 1]] - P[F[:, 0]], P[F[:, 2]] - P[F[:, 0]])
    N = np.zeros_like(P)
    for k in range(3):
        np.add.at(N, F[:, k], n)
    l = np.linalg.norm(N, axis=1, keepdims=True)
    l[l == 0] = 1.0
    return N / l


def lift_anisotropic(V, F, anisotropy=0.04):
    """6D lift: (x, n * anisotropy * bbox_diagonal), as set_anisotropy (mesh_geometry.cpp:144-164)."""
    N = vertex_normals(V, F)
    diag = np.linalg.norm(V[:, :3].max(0) - V[:, :3].min(0))
    return np.ascontiguousarray(np.concatenate([V[:, :3], N * (anisotropy * diag)], 1))


def sample_surface(V, F, S, seed=1):
    """Uniform-by-area random points on the triangles (any dimension D)."""
    rng = np.random.default_rng(seed)
    P = V[F.astype(np.int64)]
    e1, e2 = P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]
    a = np.sqrt(np.maximum((e1 * e1).sum(1) * (e2 * e2).sum(1) - ((e1 * e2).sum(1)) ** 2, 0.0))
    cdf = np.cumsum(a)
    f = np.minimum(np.searchsorted(cdf, rng.random(S) * cdf[-1]), F.shape[0] - 1)
    u, v = rng.random(S), rng.random(S)
    flip = u + v > 1.0
    u[flip], v[flip] = 1.0 - u[flip], 1.0 - v[flip]
    return np.ascontiguousarray(P[f, 0] + u[:, None] * e1[f] + v[:, None] * e2[f])


def sample_volume(V, T, S, seed=1):
    """Uniform random points inside the tetrahedra."""
    rng = np.random.default_rng(seed)
    P = V[T.astype(np.int64)]
    vol = np.abs(np.einsum("ij,ij->i", np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), P[:, 3] - P[:, 0]))
    cdf = np.cumsum(vol)
    t = np.minimum(np.searchsorted(cdf, rng.random(S) * cdf[-1]), T.shape[0] - 1)
    w = -np.log(rng.random((S, 4)))
    w /= w.sum(1, keepdims=True)
    return np.ascontiguousarray((P[t] * w[:, :, None]).sum(1))
