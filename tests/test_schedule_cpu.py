"""The GEMM engine's work decomposition, checked on the CPU through the `prl_debug_schedule` test hook (the same
`decode_item` the kernels run, compiled for the host): every output tile is covered exactly once, vocab sweeps partition the
n-tiles of each m-tile, K-splits partition the k-blocks of each tile, and the static makespan is what the design states."""

import ctypes
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as g

    g.build()
    from prime_rl_cse_b200 import _lib

    return _lib.load()


def schedule(L, kind, M, N, K):
    dims = (ctypes.c_int32 * 8)()
    assert L.prl_debug_schedule(kind, M, N, K, None, 0, dims) == 0
    n = dims[3]
    items = (ctypes.c_int32 * (6 * max(n, 1)))()
    assert L.prl_debug_schedule(kind, M, N, K, items, n, dims) == 0
    return np.ctypeslib.as_array(items).reshape(-1, 6)[:n].copy(), dict(zip(
        ("m_tiles", "n_tiles", "k_blocks", "num_items", "workers", "k_splits", "n_splits", "rows_per_m_tile"), dims))


SHAPES = [  # (M, N, K)
    (16383, 151936, 4096), (16384, 151936, 4096), (32768, 151936, 4096), (16384, 151936, 2048), (4096, 151936, 1024),
    (129, 1003, 64), (1, 8, 8), (700, 20064, 512), (16384, 18992, 5120),
]


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_vocab_sweeps_partition_the_n_tiles(L, kind, M, N, K):
    items, d = schedule(L, kind, M, N, K)
    assert d["m_tiles"] == math.ceil(M / d["rows_per_m_tile"]) and d["n_tiles"] == math.ceil(N / 256) and d["k_blocks"] == math.ceil(K / 64)
    assert len(items) == d["m_tiles"] * d["n_splits"] and d["n_splits"] <= 64
    cover = np.zeros((d["m_tiles"], d["n_tiles"]), dtype=np.int32)
    for m, nb, nc, kb, kc, split in items:
        assert nc >= 1 and kb == 0 and kc == d["k_blocks"]
        cover[m, nb : nb + nc] += 1
    assert (cover == 1).all()
    # the number of vocab splits minimises the static makespan (rounds of the persistent workers x tiles per split)
    tps = int(items[:, 2].max())
    best = min(math.ceil(d["m_tiles"] * math.ceil(d["n_tiles"] / math.ceil(d["n_tiles"] / s)) / d["workers"]) * math.ceil(d["n_tiles"] / s)
               for s in range(1, min(d["n_tiles"], 64) + 1))
    assert math.ceil(len(items) / d["workers"]) * tps == best


@pytest.mark.parametrize("kind", [2, 3])
@pytest.mark.parametrize("M,N,K", [(16384, 4096, 151936), (16384, 2048, 151936), (4096, 1024, 151936), (32768, 4096, 151936),
                                   (151936, 4096, 16384), (16384, 4096, 16384), (129, 64, 1003), (600, 256, 20064), (1, 8, 8),
                                   (16384, 5120, 18992), (513, 136, 4104)])
def test_store_tiles_cover_every_tile_and_k_block_once(L, kind, M, N, K):
    items, d = schedule(L, kind, M, N, K)
    assert len(items) == d["m_tiles"] * d["n_tiles"] * d["k_splits"] and d["k_splits"] in (1, 2)
    cover = np.zeros((d["m_tiles"], d["n_tiles"], d["k_blocks"]), dtype=np.int16)
    for m, nb, nc, kb, kc, split in items:
        assert nc == 1 and kc >= 1
        cover[m, nb, kb : kb + kc] += 1
    assert (cover == 1).all()
    # K is split in two only when that shortens the static makespan by more than 3 % and K is long enough
    tiles, w = d["m_tiles"] * d["n_tiles"], d["workers"]
    gain = (math.ceil(2 * tiles / w) / 2) < 0.97 * math.ceil(tiles / w)
    assert (d["k_splits"] == 2) == (gain and d["k_blocks"] >= 16)
    if d["k_splits"] == 2:  # the two halves of a tile are neighbouring items: they run at the same time
        first = {}
        for i, (m, nb, nc, kb, kc, split) in enumerate(items):
            first.setdefault((m, nb), []).append(i)
        assert all(len(v) == 2 and v[1] - v[0] <= 4 for v in first.values())


def test_headline_shapes_are_not_k_split_and_small_ones_are(L):
    assert schedule(L, 3, 16384, 4096, 151936)[1]["k_splits"] == 1   # Qwen3-8B: 512 tiles on 74 CTA pairs, 6.9 waves
    assert schedule(L, 3, 32768, 4096, 151936)[1]["k_splits"] == 1
    assert schedule(L, 3, 16384, 5120, 75968)[1]["k_splits"] == 1    # Qwen3-32B, vocab-sharded x2
    assert schedule(L, 3, 16384, 2048, 151936)[1]["k_splits"] == 2   # Qwen3-1.7B: 256 tiles = 3.5 waves
    assert schedule(L, 3, 4096, 1024, 151936)[1]["k_splits"] == 2    # Qwen3-0.6B: 32 tiles < 74 workers
    assert L.prl_debug_schedule(9, 1, 1, 1, None, 0, (ctypes.c_int32 * 8)()) == -1


@pytest.mark.parametrize("peers,T_local,H,K", [(8, 16384, 5120, 18992), (2, 16384, 5120, 75968), (4, 2048, 4096, 37984), (8, 2048, 4096, 18992),
                                              (2, 512, 256, 1000), (8, 32, 64, 128)])
def test_fused_reduce_scatter_tile_order(L, peers, T_local, H, K):
    """prl_lmhead_bwd_scatter: every rank covers every output tile exactly once, starts at the NEXT rank's tokens, and -- when the
    token ranges are whole tiles -- at every position of the tile list the ranks target pairwise different owners (no rank's
    NVLink inbound is hit by all the others at once)."""
    M = peers * T_local
    orders = []
    for r in range(peers):
        dims = (ctypes.c_int32 * 8)()
        assert L.prl_debug_schedule_scatter(M, H, K, peers, r, None, 0, dims) == 0
        n = dims[3]
        items = (ctypes.c_int32 * (6 * n))()
        assert L.prl_debug_schedule_scatter(M, H, K, peers, r, items, n, dims) == 0
        it = np.ctypeslib.as_array(items).reshape(-1, 6).copy()
        m_tiles, n_tiles, rows_per_tile = dims[0], dims[1], dims[7]
        cover = np.zeros((m_tiles, n_tiles), dtype=np.int32)
        for m, nb, nc, kb, kc, split in it:
            assert nc == 1 and kb == 0 and kc == dims[2]
            cover[m, nb] += 1
        assert (cover == 1).all()
        assert dims[6] == T_local
        owner_of_first = min(it[0, 0] * rows_per_tile // T_local, peers - 1)
        if T_local % rows_per_tile == 0:
            assert owner_of_first == (r + 1) % peers
        orders.append(it[:, 0] * rows_per_tile // T_local)
    if T_local % rows_per_tile == 0:
        owners = np.stack(orders)  # [rank, item] -> owner of the tile that rank writes at that position
        assert all(len(set(owners[:, i])) == peers for i in range(owners.shape[1]))
    assert L.prl_debug_schedule_scatter(1000, 64, 64, 3, 0, None, 0, (ctypes.c_int32 * 8)()) == -1  # M % n_peers != 0
