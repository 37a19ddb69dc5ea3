"""CPU suite: the C-ABI library loads and exports every symbol include/b200cvt.h declares
(no compute call without a GPU), shapes and sharding helpers behave, gloo all-gather of the
sharded Lloyd step reproduces the single-process oracle."""
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

from graphitethree_b200 import capi, shapes, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    header = open(capi.HEADER_PATH).read()
    declared = set(re.findall(r"\b(b200cvt_[a-z_0-9]+)\s*\(", header))
    declared -= {"b200cvt_progress_cb", "b200cvt_exchange_cb"}
    L = capi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "missing export: " + name
    assert declared == set(capi.exported_symbols())


def test_no_oracle_or_cpu_fallback_in_product():
    # the product must not import, link or execute anything under oracle/
    for dirpath, _, files in os.walk(os.path.join(ROOT, "graphitethree_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "").replace("the oracle", "").replace("CPU oracle", "") or f == "shapes.py", f
                assert "import oracle" not in src and "from oracle" not in src, f
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "geogram" not in out


def test_create_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.B200CVTError) as ei:
        capi.Handle(3)
    assert ei.value.code == 2


def test_shapes():
    V, F = shapes.icosphere(45)
    assert F.shape[0] == 40500 and V.shape[0] == 20252            # C1 mesh, Euler characteristic 2
    V, F = shapes.icosphere_split(3)
    assert F.shape[0] == 20 * 4 ** 3
    V, F = shapes.trefoil_tube(40, 8)
    assert F.shape[0] == 2 * 40 * 8
    V, T = shapes.kuhn_cube(3)
    P = V[T.astype(np.int64)]
    vol = np.einsum("ij,ij->i", np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), P[:, 3] - P[:, 0]) / 6.0
    assert np.allclose(vol, 1.0 / (6 * 27)) and T.shape[0] == 162
    V, F = shapes.cad_like(3)
    E = np.sort(np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]]), 1)
    _, c = np.unique(E, axis=0, return_counts=True)
    assert (c == 2).all()                                          # closed manifold


def test_partition_arithmetic():
    for S, n in ((10, 3), (7, 8), (1000, 4), (5, 1)):
        cover = []
        for r in range(n):
            b, e = sharding.owned_range(S, r, n)
            cover += list(range(b, e))
        assert cover == list(range(S))
        assert sharding.chunk_doubles(3, S, n) == sharding.slice_len(S, n) * 4


_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from graphitethree_b200 import shapes, sharding
from oracle import port
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
V, F = shapes.icosphere(6)
X = shapes.sample_surface(V, F, 151, 3)
S, dim = X.shape
x = X.copy()
for it in range(3):
    # every rank holds all seeds; Morton order stands in for any fixed permutation here
    order = np.lexsort((x[:, 2], x[:, 1], x[:, 0])).astype(np.int64)
    e = port.surface_eval(V, F, x, 0, False)
    new = x.copy()
    ok = e.m > 1e-30
    new[ok] = (1.0 / e.m[ok])[:, None] * e.mg[ok]
    ex = sharding.TorchExchange(dim, S, rank, world, "cpu")
    b, en = sharding.owned_range(S, rank, world)
    contrib = np.zeros_like(new); contrib[order[b:en]] = new[order[b:en]]       # only the owned slice is trusted
    ex.slice.copy_(torch.from_numpy(sharding.pack_slice(contrib[order], e.m[order], rank, world)))
    assert ex() == 0
    x, m_sorted = sharding.unpack_all(ex.all.numpy(), S, dim, world, order)
    assert np.array_equal(m_sorted, e.m[order])
ref, _ = port.lloyd(V, F, X, 3)
assert np.array_equal(x, ref), np.abs(x - ref).max()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_sharded_lloyd_exchange_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port_no = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out


def test_bench_reference_arm_contract():
    # `bench.py --impl reference`: ONE JSON line on stdout (library banners go to stderr), the contract's keys, timed on
    # the host cores with the reference built from its own sources (or the oracle port when it is absent)
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--small", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "CVT seed-iterations/sec" and d["unit"] == "seed-iterations/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_header_is_plain_c_and_links(built, tmp_path):
    # include/b200cvt.h must be usable from C (the reference-side FFI is extern "C"): compile a C99 client that takes the
    # address of every declared entry point with -Wall -Werror and link it against the library
    header = open(capi.HEADER_PATH).read()
    names = sorted(set(re.findall(r"\b(b200cvt_[a-z_0-9]+)\s*\(", header)) - {"b200cvt_progress_cb", "b200cvt_exchange_cb"})
    src = tmp_path / "client.c"
    src.write_text('#include "b200cvt.h"\n#include <stdio.h>\nint main(void) {\n    const void* f[] = {\n'
                   + "".join("        (const void*)%s,\n" % n for n in names)
                   + '    };\n    printf("%zu\\n", sizeof(f) / sizeof(f[0]));\n    return b200cvt_last_error() == 0;\n}\n')
    exe = tmp_path / "client"
    root = os.path.dirname(capi.HEADER_PATH)
    libdir = os.path.dirname(capi.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-Wno-pedantic", "-I", root, str(src), "-o", str(exe),
                        "-L", libdir, "-lb200cvt", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and int(r.stdout.strip()) == len(names)
