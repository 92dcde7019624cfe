"""CPU tier: the __host__ __device__ building blocks of 2048_b200/csrc/b2048_device.cuh, compiled for
the host by tests/host_shim.cu, checked against the golden fixtures and the oracle.  This is the same
source the CUDA kernels are built from (packed moves, predicates, D4 images, features, Philox spawns)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

SRC = os.path.join(ROOT, "tests", "host_shim.cu")
HDR = os.path.join(ROOT, "2048_b200", "csrc", "b2048_device.cuh")
OUT = os.path.join(ROOT, "tests", "_build", "libhostshim.so")


@pytest.fixture(scope="module")
def hs():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["nvcc", "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets",
                        "-Xcompiler", "-fPIC", "-shared", SRC, "-o", OUT], check=True, capture_output=True)
    lib = C.CDLL(OUT)
    lib.hs_table_offset.restype = C.c_int64
    lib.hs_spawn_initial.restype = C.c_uint64
    lib.hs_spawn_initial.argtypes = [C.c_uint64, C.c_uint64]
    lib.hs_spawn_move.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
    lib.hs_spawn_move.restype = C.c_uint32
    lib.hs_spawn_sweep.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    lib.hs_spawn_sweep.restype = C.c_uint32
    lib.hs_spawn_nonempty_agrees.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    lib.hs_d4.argtypes = [C.c_uint64, C.c_void_p]
    lib.hs_move4.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64] + [C.c_void_p] * 4
    lib.hs_stats.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    lib.hs_features.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    lib.hs_features_fast.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    return lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def lut_of(hs):
    lut = np.zeros(65536, np.uint32)
    hs.hs_lut(ptr(lut))
    return lut


def test_lut_entries_match_reference_table(hs):
    """all 65,536 LUT entries decode to Game.table (game_logic.py:18-39): line, score, changed, overflow"""
    g = load_golden("move_table.npz")
    lut = lut_of(hs)
    lines = np.stack([(lut >> 12) & 15, (lut >> 8) & 15, (lut >> 4) & 15, lut & 15], axis=1).astype(np.uint8)
    a, b = (lut >> 16) & 15, (lut >> 20) & 15
    score = np.where(a > 0, 2 << a.astype(np.int64), 0) + np.where(b > 0, 2 << b.astype(np.int64), 0)
    ovf = (g["lines"] == 16).any(axis=1)
    assert np.array_equal(lines, np.minimum(g["lines"], 15))
    assert np.array_equal(score, g["score"].astype(np.int64))
    assert np.array_equal((lut >> 24) & 1, g["changed"])
    assert np.array_equal(((lut >> 25) & 1).astype(bool), ovf) and ovf.sum() == 767
    assert not (lut >> 26).any()


@pytest.mark.parametrize("split", [0, 1])
def test_move4_and_predicates(hs, orc, split):
    """packed 4-direction moves (global-LUT and shared-memory-split decode), game_over, counts"""
    g = load_golden("boards.npz")
    rows = g["boards"].astype(np.int32)
    boards = orc.pack_np(rows)
    m = len(boards)
    after = np.zeros((m, 4), np.uint64)
    gain = np.zeros((m, 4), np.uint32)
    flags = np.zeros(m, np.uint8)
    over = np.zeros(m, np.uint8)
    lut = lut_of(hs)
    hs.hs_move4(ptr(lut), split, ptr(boards), m, ptr(after), ptr(gain), ptr(flags), ptr(over))
    ref_after = g["after"].astype(np.int32)                 # [m,4,4,4], may contain 16
    ovf = (ref_after > 15).any(axis=(2, 3))
    assert np.array_equal(orc.unpack_np(after.reshape(-1)).reshape(m, 4, 4, 4), np.minimum(ref_after, 15))
    assert np.array_equal(gain.astype(np.int64), g["gain"].astype(np.int64))
    for d in range(4):
        assert np.array_equal((flags >> d) & 1, g["change"][:, d])
        assert np.array_equal(((flags >> (4 + d)) & 1).astype(bool), ovf[:, d])
    assert ovf.sum() >= 3
    assert np.array_equal(over, g["over"])
    stats = np.zeros((m, 4), np.uint8)
    hs.hs_stats(ptr(boards), m, ptr(stats))
    assert np.array_equal(stats[:, 0], g["n_empty"].astype(np.uint8))
    assert np.array_equal(stats[:, 1], g["n_pairs"].astype(np.uint8))
    assert np.array_equal(stats[:, 2], g["over"])
    assert np.array_equal(stats[:, 3], rows.reshape(m, -1).max(axis=1).astype(np.uint8))


def test_exhaustive_rows_in_all_directions(hs, orc):
    """every one of the 65,536 lines embedded as a row (left/right) and as a column (up/down)"""
    lines = np.arange(65536, dtype=np.uint64)
    rows = np.stack([(lines >> 12) & 15, (lines >> 8) & 15, (lines >> 4) & 15, lines & 15], axis=1).astype(np.int32)
    filler = np.array([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]], dtype=np.int32)
    b_row = np.concatenate([np.repeat(filler[None, :1], 65536, 0), rows[:, None, :],
                            np.repeat(filler[None, 1:], 65536, 0)], axis=1)       # line at row 1
    b_col = np.transpose(b_row, (0, 2, 1)).copy()                                   # line at column 1
    lut = lut_of(hs)
    for rows4 in (b_row, b_col):
        boards = orc.pack_np(rows4)
        m = len(boards)
        after = np.zeros((m, 4), np.uint64)
        gain = np.zeros((m, 4), np.uint32)
        flags = np.zeros(m, np.uint8)
        over = np.zeros(m, np.uint8)
        hs.hs_move4(ptr(lut), 0, ptr(boards), m, ptr(after), ptr(gain), ptr(flags), ptr(over))
        ra, rs, rc = orc.pre_move_batch(rows4)
        assert np.array_equal(orc.unpack_np(after.reshape(-1)).reshape(m, 4, 4, 4), np.minimum(ra, 15))
        assert np.array_equal(gain.astype(np.int64), rs)
        assert np.array_equal(np.stack([(flags >> d) & 1 for d in range(4)], 1), rc.astype(np.uint8))


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6])
def test_features(hs, orc, n):
    g = load_golden("boards.npz")
    ref = g[f"f_{n}"]
    boards = orc.pack_np(g["boards"][:len(ref)].astype(np.int32))
    feat = np.zeros_like(ref)
    assert hs.hs_features(n, ptr(boards), len(boards), ptr(feat)) == ref.shape[1]
    assert np.array_equal(feat, ref)
    assert [hs.hs_table_offset(n, i) for i in range(ref.shape[1] + 1)] == orc.table_offsets(n).tolist()
    if n >= 4:                                   # the shared-work routine of the evaluate kernels: same indices
        fast = np.zeros_like(ref)
        assert hs.hs_features_fast(n, ptr(boards), len(boards), ptr(fast)) == ref.shape[1]
        assert np.array_equal(fast, ref)
        rng = np.random.default_rng(n)
        rnd = rng.integers(0, 16, size=(20000, 16)).astype(np.int32)              # includes exponents 14, 15 (clamp)
        rb = orc.pack_np(rnd)
        a, b = np.zeros((len(rb), ref.shape[1]), np.int32), np.zeros((len(rb), ref.shape[1]), np.int32)
        hs.hs_features(n, ptr(rb), len(rb), ptr(a))
        hs.hs_features_fast(n, ptr(rb), len(rb), ptr(b))
        assert np.array_equal(a, b) and np.array_equal(a, orc.features_batch(n, rnd))


def test_d4_images_are_the_reference_set(hs, orc):
    """the 8 images of d4_image() == the set QAgent.update visits (r_learning.py:207-214)"""
    g = load_golden("d4.npz")
    ref = {tuple(i) for i in g["images"].astype(int)}
    out = np.zeros(8, np.uint64)
    hs.hs_d4(int(orc.pack_np(np.arange(16).reshape(1, 4, 4))[0]), ptr(out))
    got = {tuple(orc.unpack_np(out[s:s + 1]).ravel()) for s in range(8)}
    assert got == ref and len(got) == 8
    # update keys multiplicities through the packed path, n = 4
    sample = g["sample"].astype(np.int32)
    offs = g["offs_4"]
    toff = orc.table_offsets(4)
    for q, b in enumerate(sample):
        hs.hs_d4(int(orc.pack_np(b[None])[0]), ptr(out))
        feat = np.zeros((8, 17), np.int32)
        hs.hs_features(4, ptr(out), 8, ptr(feat))
        k, c = np.unique((feat + toff[None, :17]).ravel(), return_counts=True)
        assert np.array_equal(k, g["keys_4"][offs[q]:offs[q + 1]]) and np.array_equal(c, g["counts_4"][offs[q]:offs[q + 1]])


def test_philox_and_spawns_match_oracle(hs, orc, fx):
    out = np.zeros(4, np.uint32)
    c = np.array([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], np.uint32)
    k = np.array([0xa4093822, 0x299f31d0], np.uint32)
    hs.hs_philox(ptr(c), ptr(k), ptr(out))
    assert out.tolist() == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    for gid in list(range(200)) + [2 ** 40 + 7, 2 ** 63 + 1]:
        assert hs.hs_spawn_initial(99, gid) == int(orc.pack_np(orc.spawn_initial(99, gid)[None])[0])
    boards = fx.random_boards(400, seed=3)
    for i, row in enumerate(boards):
        b = np.array([orc.pack_np(row[None])[0]], np.uint64)
        res = hs.hs_spawn_move(5, i, i % 50 + 1, ptr(b))
        new, ores = orc.spawn_move(5, i, i % 50 + 1, row)
        assert b[0] == orc.pack_np(new[None])[0]
        assert res == (0xFFFF if ores < 0 else ores)
        b = np.array([orc.pack_np(row[None])[0]], np.uint64)
        hs.hs_spawn_sweep(7, 1000 + i, i % 4, ptr(b))
        r2 = row.reshape(16).copy()
        orc.lib().orc_spawn_sweep(7, 1000 + i, i % 4, r2.ctypes.data_as(C.POINTER(C.c_int32)))
        assert b[0] == orc.pack_np(r2.reshape(1, 4, 4))[0]
    # the search-free spawn used in the hot loops (boards with at least one tile) == the general one: 1..15 empty cells
    # in every position pattern class, every k, both tiles
    rs = np.random.RandomState(11)
    for _ in range(20000):
        cells = rs.randint(0, 16, 16) * (rs.random_sample(16) < rs.random_sample())
        if not cells.any():
            cells[rs.randint(16)] = 1 + rs.randint(15)
        b = int(orc.pack_np(cells.reshape(1, 4, 4).astype(np.int32))[0])
        assert hs.hs_spawn_nonempty_agrees(b, int(rs.randint(0, 2 ** 32, dtype=np.uint64)), int(rs.randint(0, 2 ** 32, dtype=np.uint64)))
    for pos in range(16):                                         # exactly one tile: 15 empties, every k
        b = 3 << (4 * pos)
        for k in range(15):
            r_pos = ((k << 32) // 15 + 1) & 0xFFFFFFFF
            assert hs.hs_spawn_nonempty_agrees(b, 0, r_pos) and hs.hs_spawn_nonempty_agrees(b, 2 ** 31, r_pos)
