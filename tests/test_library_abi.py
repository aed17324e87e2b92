"""CPU-side checks of the drop-in boundary: the shared library loads, exports every
symbol include/fxb200.h declares, the ctypes prototypes cover the header, struct
layouts match, and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import fxb200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fxb200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    syms = header_symbols()
    assert len(syms) >= 20
    lib = fxb200.load_library()
    out = subprocess.run(["nm", "-D", "--defined-only", fxb200.LIB_PATH], capture_output=True,
                         text=True, check=True).stdout
    exported = set(re.findall(r" T (fx_[a-z0-9_]+)", out))
    for s in syms:
        assert s in exported, "header declares %s but the library does not export it" % s
        assert s in fxb200.PROTOTYPES, "no ctypes prototype for %s" % s
        assert hasattr(lib, s)
    assert lib.fx_version() >= 100


def test_struct_layouts_match_header():
    assert ctypes.sizeof(fxb200.SolverOpts) == 48
    assert ctypes.sizeof(fxb200.Stats) == 13 * 8 + 7 * 8 + 6 * 4


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fxb200.FxError) as ei:
        fxb200.kpca(np.random.rand(32, 3), fxb200.GaussianKernel(1.0), k=2)
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "featureextraction.jl_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", txt, flags=re.M), f


def test_sharding_rules():
    from fxb200.sharding import row_partition, query_partition
    ld, n_ld, parts = row_partition(65536, 8)
    assert ld == 8192 and n_ld == 65536 and parts[3] == (3 * 8192, 8192)
    ld, n_ld, parts = row_partition(1000, 8)
    assert ld == 128 and n_ld == 1024 and parts[7] == (896, 104)
    ld, n_ld, parts = row_partition(500, 8)
    assert [p[1] for p in parts] == [128, 128, 128, 116, 0, 0, 0, 0]
    assert sum(c for _, c in query_partition(1000003, 8)) == 1000003
    with pytest.raises(ValueError):
        row_partition(0, 2)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("peer", [False, True])
def test_symmetric_gram_tile_ownership_covers_every_tile_once(world, peer):
    """The ownership rule of the symmetric Gram kernel (csrc/gram.cu tile_role, restated in
    sharding.gram_tile_roles): over all ranks every 64x64 cell of the global matrix is produced
    exactly once, either by the tile that computes it or by exactly one mirror store."""
    import numpy as np
    n = 128 * 3 * world - 5
    ld = fxb200.sharding.row_partition(n, world)[0]
    ty_n = ld // 128
    cells = np.zeros((2 * ty_n * world, 2 * ty_n * world), dtype=np.int64)   # 64x64 cells
    for r in range(world):
        role, target = fxb200.sharding.gram_tile_roles(n, world, r, peer)
        for ty in range(role.shape[0]):
            for tx in range(role.shape[1]):
                if role[ty, tx] == 0:
                    continue
                gi = 2 * (r * ty_n + ty)                     # first 64-row cell of the 128-row tile
                cells[gi:gi + 2, tx] += 1
                if role[ty, tx] == 2:                        # transposed: 64 rows x 128 columns
                    assert peer or target[ty, tx] == r
                    cells[tx, gi:gi + 2] += 1
    assert (cells == 1).all()
    ex = [fxb200.sharding.gram_tiles_executed(n, world, r, peer)[0] for r in range(world)]
    assert max(ex) - min(ex) <= ty_n                      # balanced across ranks


def _header_prototypes():
    """name -> (return class, [argument classes]) parsed from include/fxb200.h;
    classes: 'int', 'i64', 'f64', 'ptr', 'void', 'cstr'."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)

    def cls(t):
        t = t.strip()
        if "*" in t:
            return "cstr" if re.match(r"const\s+char\s*\*", t) else "ptr"
        base = re.sub(r"\bconst\b", "", t).split()
        if not base:
            return "void"
        # drop the parameter name if present
        ty = base[0] if len(base) == 1 else " ".join(base[:-1])
        return {"int": "int", "int32_t": "int", "int64_t": "i64", "double": "f64", "void": "void",
                "uint64_t": "i64", "size_t": "i64"}[ty]

    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(fx_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argl = [] if args.strip() in ("", "void") else [cls(a) for a in args.split(",")]
        protos[name] = (cls(ret + " x") if "*" not in ret else cls(ret), argl)
    return protos


def test_julia_wrapper_ccalls_match_header():
    """Julia is not in the image, so the wrapper cannot be executed here: check statically that
    every ccall names a declared symbol and passes the declared number and kinds of arguments."""
    protos = _header_prototypes()
    src = open(os.path.join(ROOT, "julia", "FeatureExtractionB200.jl")).read()
    jl = {"Cint": "int", "Int64": "i64", "Float64": "f64", "Cvoid": "void", "Cstring": "cstr"}
    calls = list(re.finditer(r"ccall\(\(:(fx_[a-z0-9_]+),\s*libfx\),\s*(\w+),\s*\(", src))
    assert len(calls) >= 10
    seen = set()
    for m in calls:
        name, ret = m.group(1), m.group(2)
        # argument type tuple: balanced parentheses from the '(' that ends the match
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        tup = src[m.end():i - 1]
        parts = [p.strip() for p in re.split(r",(?![^{}]*\})", tup) if p.strip()]
        kinds = ["ptr" if p.startswith(("Ptr{", "Ref{")) else jl[p] for p in parts]
        assert name in protos, "ccall of undeclared symbol %s" % name
        hret, hargs = protos[name]
        assert jl[ret] == hret, (name, ret, hret)
        assert kinds == hargs, "%s: Julia passes %s, header declares %s" % (name, kinds, hargs)
        seen.add(name)
    for required in ("fx_init", "fx_shutdown", "fx_kpca_fit_f64", "fx_kpca_transform_f64", "fx_klda_fit_f64",
                     "fx_model_create_f64", "fx_model_free", "fx_pca_f64", "fx_lda_f64", "fx_project_f64"):
        assert required in seen, required


def test_ctypes_prototypes_match_header():
    """The ctypes argument lists have the header's arity and pointer / scalar kinds."""
    protos = _header_prototypes()
    kind = {ctypes.c_int: "int", ctypes.c_int32: "int", ctypes.c_int64: "i64", ctypes.c_double: "f64",
            ctypes.c_uint64: "i64", ctypes.c_size_t: "i64"}
    for name, (restype, argtypes) in fxb200.PROTOTYPES.items():
        assert name in protos, name
        hret, hargs = protos[name]
        got = [kind.get(a, "ptr") for a in argtypes]
        assert got == hargs, "%s: ctypes %s vs header %s" % (name, got, hargs)
        if restype is None:
            assert hret == "void", name
        elif restype is ctypes.c_char_p:
            assert hret == "cstr", name
        else:
            assert kind.get(restype, "ptr") == hret, name


def build_c_smoke(tmp_path):
    """gcc -std=c99 -pedantic build of tests/host/abi_smoke.c against the in-tree library."""
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(fxb200.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-o", exe,
                        os.path.join(ROOT, "tests", "host", "abi_smoke.c"), "-L" + libdir, "-lfxb200", "-lm",
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_valid_c99_and_c_caller_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = build_c_smoke(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present (the GPU leg is tests/test_gpu_parity.py::test_plain_c_caller)")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stdout, r.stdout + r.stderr


def test_sharding_properties_hypothesis():
    """Random (n, world): the row partition covers [0, n) exactly with only trailing padding, the
    query partition is a contiguous cover, and the symmetric-Gram ownership stays exact-once and
    balanced (property test over the Python restatement; the C++ side is tied to it by
    tests/test_device_math_host.py)."""
    from hypothesis import given, settings, strategies as st
    import numpy as np
    from fxb200.sharding import row_partition, query_partition, gram_tile_roles

    @settings(max_examples=60, deadline=None)
    @given(n=st.integers(1, 4000), world=st.integers(1, 8), m=st.integers(0, 10**7))
    def prop(n, world, m):
        ld, n_ld, parts = row_partition(n, world)
        assert ld % 128 == 0 and n_ld == ld * world and n_ld >= n
        assert sum(r for _, r in parts) == n
        nxt = 0
        for row0, rows in parts:
            assert row0 % ld == 0 and 0 <= rows <= ld
            if rows:
                assert row0 == nxt        # valid rows are contiguous in the global index space
                nxt += rows
            if rows < ld:
                nxt = n                   # only the last non-empty rank is ragged
        qp = query_partition(m, world)
        assert sum(c for _, c in qp) == m and all(qp[i][0] + qp[i][1] == qp[i + 1][0] for i in range(world - 1))
        ty_n = ld // 128
        cells = np.zeros((2 * ty_n * world, 2 * ty_n * world), dtype=np.int64)
        ex = []
        for r in range(world):
            role, target = gram_tile_roles(n, world, r, peer=True)
            ex.append(int((role > 0).sum()))
            tys, txs = np.nonzero(role)
            for ty, tx in zip(tys, txs):
                gi = 2 * (r * ty_n + ty)
                cells[gi:gi + 2, tx] += 1
                if role[ty, tx] == 2:
                    assert tx // (2 * ty_n) == target[ty, tx]
                    cells[tx, gi:gi + 2] += 1
        assert (cells == 1).all()
        assert max(ex) - min(ex) <= max(2, ty_n)      # an odd checkerboard leaves one super-tile (2 tiles) over

    prop()


def test_null_context_is_an_argument_error_everywhere():
    """Every entry point that takes a context rejects NULL with a negative code (argument error,
    the analogue of common.jl:87-89) instead of crashing -- callable without a GPU."""
    lib = fxb200.load_library()
    for name, (restype, argtypes) in fxb200.PROTOTYPES.items():
        if restype is not ctypes.c_int or not argtypes or name in ("fx_init", "fx_init_distributed",
                                                                    "fx_comm_unique_id", "fx_version"):
            continue
        args = []
        for a in argtypes:
            if a in (ctypes.c_int, ctypes.c_int32, ctypes.c_int64):
                args.append(a(8))
            elif a is ctypes.c_double:
                args.append(a(0.0))
            else:
                args.append(None)                   # every pointer NULL, the context first of all
        rc = getattr(lib, name)(*args)
        if name == "fx_ctx_peer_access":
            assert rc == 0
        else:
            assert rc < 0, "%s(NULL, ...) returned %d" % (name, rc)
            assert lib.fx_last_error()
