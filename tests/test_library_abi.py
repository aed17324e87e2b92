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
