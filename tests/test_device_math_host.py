"""The Gaussian kernel map's branch-free exp (csrc/fx_common.cuh fx_exp_nonpos) is compiled for
the HOST from the very same source and measured against expl(): pins the 13-operation algorithm,
its constants and the 2^(j/64) hi/lo table on CPU, where no GPU is needed.  (On the device the
same code runs through the parity tests against the oracle's np.exp.)"""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_device_exp_against_long_double(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "exp_host")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(ROOT, "tests", "host", "exp_host.cu")],
                   check=True, capture_output=True, text=True)
    out = subprocess.run([exe, "4000000"], check=True, capture_output=True, text=True).stdout
    r = json.loads(out)
    assert r["max_ulp"] < 0.9, r                  # DESIGN.md / fx_common.cuh state 0.83 ulp
    assert r["exp0"] == 1.0
    assert abs(r["exp_tiny_pos"] - 1.0) <= 2.3e-16
    assert 0.0 < r["exp_minus_1e6"] < 1e-307      # clamped at x = -708, never 0 / subnormal / NaN
    assert r["non_monotone"] == 0
    assert r["max_abs_err_below_minus_708"] < 3.4e-308


@pytest.mark.parametrize("tiles_y", [3, 130])
def test_gram_tile_ownership_and_sweep_host_build(tmp_path, tiles_y):
    """csrc/gram_tiles.cuh (the header gram.cu uses) compiled for the host: bijective sweep,
    every cell of the global matrix produced exactly once over all ranks, distinct peers per sweep
    block; and its roles equal the Python restatement used by bench.py and the gloo tests."""
    import re
    import numpy as np
    import fxb200
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "gram_tiles_host")
    subprocess.run([gxx, "-x", "c++", "-std=c++17", "-O1", "-o", exe,
                    os.path.join(ROOT, "tests", "host", "gram_tiles_host.cu")], check=True, capture_output=True, text=True)
    r = subprocess.run([exe, str(tiles_y)], capture_output=True, text=True)
    assert r.returncode == 0 and "failures=0" in r.stdout, r.stdout[-2000:]
    lines = re.findall(r"sym=(\d) world=(\d+) rank=(\d+) executed=(\d+) checksum=(\d+)", r.stdout)
    assert len(lines) == 2 * sum(range(1, 9))
    for sym, world, rank, executed, checksum in (tuple(int(v) for v in l) for l in lines):
        role, target = fxb200.sharding.gram_tile_roles(128 * tiles_y * world, world, rank, peer=(sym == 2))
        ty, tx = np.indices(role.shape)
        want = int((role * (ty * 131 + tx * 7 + 1)).sum() + (1000003 * target[role == 2]).sum())
        assert (int((role > 0).sum()), want) == (executed, checksum), (sym, world, rank)
