"""The C ABI from a non-Python caller: tests/c_abi_driver.c includes only include/mpb200.h, is
compiled here with gcc against libmpb200.so, runs BASELINE config C1 (N = 1000, ISRR_2H) and its output bytes are
compared with the oracle."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

import fixtures as fx
from conftest import unpack_bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "motionplanning.jl_b200", "csrc")


def _compile(tmp):
    exe = os.path.join(tmp, "c_abi_driver")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi_driver.c"), "-o", exe, "-L", CSRC, "-lmpb200", "-Wl,-rpath," + CSRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_c_driver_compiles_against_the_header_only(tmp_path):
    """no GPU needed: the driver builds with -Wall -Werror from include/mpb200.h alone and links the library"""
    import mpb200
    mpb200.load()
    assert shutil.which("gcc")
    _compile(str(tmp_path))


@pytest.mark.gpu
def test_c_driver_runs_c1_and_matches_the_oracle(gpu, orc, tmp_path):
    mp = gpu
    N = 1000
    O = orc.Obstacles2D(fx.ISRR_2H)
    So = orc.StateSpace([0, 0], [1, 1])
    cand = fx.uniform_samples(3 * N, 2, 20240601)
    V = np.ascontiguousarray(np.vstack([[0.1, 0.1], cand[orc.states_free(O, So, cand)][:N - 2], [0.9, 0.9]]))
    r = fx.fmt_radius(N, 2)
    pk = mp.shapes2d.pack_obstacles(mp.obstaclesets.ISRR_2H())
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([N, 2, pk["n_gates"], pk["n_shapes"], len(pk["data"])], dtype=np.int64).tofile(f)
        np.array([r, 0.0, 0.0, 1.0, 1.0], dtype=np.float64).tofile(f)            # r, lo, hi
        V.tofile(f)                                                               # row per state == column-major d x N
        pk["gate_parent"].astype(np.int32).tofile(f)
        pk["gate_aabb"].astype(np.float64).tofile(f)
        pk["shape_kind"].astype(np.int32).tofile(f)
        pk["shape_gate"].astype(np.int32).tofile(f)
        pk["shape_off"].astype(np.int32).tofile(f)
        pk["data"].astype(np.float64).tofile(f)
    exe = _compile(str(tmp_path))
    res = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    raw = np.fromfile(fout, dtype=np.uint8)
    nnz, checks = (int(x) for x in raw[:16].view(np.int64))
    off = 16
    colptr = raw[off:off + 8 * (N + 1)].view(np.int64); off += 8 * (N + 1)
    rowval = raw[off:off + 8 * nnz].view(np.int64); off += 8 * nnz
    nzval = raw[off:off + 8 * nnz].view(np.float64); off += 8 * nnz
    pw, ew = (N + 63) // 64, (nnz + 63) // 64
    pbits = raw[off:off + 8 * pw].view(np.uint64); off += 8 * pw
    ebits = raw[off:off + 8 * ew].view(np.uint64); off += 8 * ew
    assert off == len(raw)
    ref = orc.KDTree(V).rball(r)
    assert np.array_equal(colptr, ref[0]) and np.array_equal(rowval, ref[1]) and nzval.tobytes() == ref[2].tobytes()
    assert np.array_equal(unpack_bits(pbits, N), orc.states_free(O, So, V))
    exp, cnt = orc.edges_free_csc(O, So, V, ref[0], ref[1])
    assert np.array_equal(unpack_bits(ebits, nnz), exp.astype(bool)) and checks == cnt
