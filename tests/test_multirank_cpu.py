"""world_size-2 `gloo` tests (CPU) of the multi-GPU HOST logic: dealing micro-batches to ranks, rank-local
loss scale, vocab-shard bounds, and the partial-statistics exchange whose merge must reproduce the
full-vocabulary log-probs / entropy (checked with the oracle's merge: no compute kernel runs on the CPU)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from prime_rl_cse_b200 import parallel as par

        # ---- token sharding: contiguous chunks, rank-local loss scale (batch.py:167-173, train.py:263-267)
        g = torch.Generator().manual_seed(0)
        mbs = []
        for i in range(6):
            T = 32 + 8 * i
            mbs.append({"id": i, "input_ids": torch.zeros(1, T, dtype=torch.long), "loss_mask": torch.rand(1, T, generator=g) < 0.5})
        mine = par.shard_micro_batches(mbs, rank, world)
        assert [m["id"] for m in mine] == list(range(rank * 3, rank * 3 + 3))
        assert par.local_loss_scale(mine, "token") == sum(int(m["loss_mask"].sum()) for m in mine)
        assert par.local_loss_scale(mine, "sequence") == 3
        try:
            par.shard_micro_batches(mbs[:5], rank, world)
            raise AssertionError("expected ValueError")
        except ValueError:
            pass

        # ---- vocab sharding: every rank computes the partial of ITS shard, exchanges 16 B/token, merges
        T, V, tau = 40, 1001, 0.7
        rng = np.random.default_rng(5)
        z = (rng.standard_normal((T, V)) * 3 / tau)
        ids = rng.integers(0, V, T)
        v0, v1 = par.vocab_shard_bounds(V, world, rank)
        part = torch.from_numpy(oracle.shard_partials(z[:, v0:v1], ids, v0).astype(np.float32))
        gathered = par.exchange_partials(part)
        assert gathered.shape == (world, T, 4)
        assert torch.equal(gathered[rank], part)
        logp, ent, lse = oracle.merge_partials(gathered.numpy())
        wl, we, wlse = oracle.logprobs_entropy(z * tau, ids, tau)  # logprobs_entropy divides by tau itself
        np.testing.assert_allclose(logp, wl, atol=2e-5)
        np.testing.assert_allclose(ent, we, atol=2e-5)
        np.testing.assert_allclose(lse, wlse, atol=2e-5)

        # ---- config 4 as rounds (VocabShardedLMHead): every rank brings ITS OWN T tokens; a rank sweeps all world*T tokens over its
        # vocab shard, the partials go to the token owners (all-to-all, 16 B per token and shard), owners merge; (lse, dlogp) of
        # all tokens are gathered for the backward with the dlogp of every row's position 0 forced to 0
        T = 24
        rng_r = [np.random.default_rng(100 + r) for r in range(world)]
        z_rows = [rg.standard_normal((T, V)) * 3 for rg in rng_r]          # un-shifted logits of every rank's packed row
        ids_rows = [rg.integers(0, V, T) for rg in rng_r]
        z_all = np.concatenate([oracle.shift_rows(zr) for zr in z_rows])    # the shift is per packed row (rl/loss.py:54-57)
        ids_all = np.concatenate(ids_rows)
        part_all = oracle.shard_partials(z_all[:, v0:v1], ids_all, v0).astype(np.float32).reshape(world, T, 4)
        mine = par.partials_to_owners(torch.from_numpy(part_all))
        assert mine.shape == (world, T, 4)
        logp, ent, lse = oracle.merge_partials(mine.numpy())
        wl, we, wlse = oracle.logprobs_entropy(oracle.shift_rows(z_rows[rank]), ids_rows[rank], 1.0)
        np.testing.assert_allclose(logp, wl, atol=2e-5)
        np.testing.assert_allclose(ent, we, atol=2e-5)
        assert abs(logp[0] + np.log(V)) < 1e-5  # position 0 of every row: uniform over the WHOLE vocabulary after the merge
        g_local = torch.from_numpy(rng_r[rank].standard_normal(T).astype(np.float32))
        lse_all, g_all = par.gather_token_scalars(torch.from_numpy(lse.astype(np.float32)), g_local)
        assert lse_all.shape == (world * T,) and torch.equal(g_all[rank * T + 1 : (rank + 1) * T], g_local[1:])
        assert all(float(g_all[r * T]) == 0.0 for r in range(world))
        np.testing.assert_allclose(lse_all[rank * T : (rank + 1) * T].numpy(), lse, atol=1e-6)
        ret[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        ret[rank] = f"{type(e).__name__}: {e}"
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=120)
        for p in procs:
            assert p.exitcode == 0, "worker died or timed out"
        assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)


def test_vocab_shard_bounds():
    from prime_rl_cse_b200 import vocab_shard_bounds

    assert [vocab_shard_bounds(151936, 8, r) for r in (0, 7)] == [(0, 18992), (132944, 151936)]
    b = [vocab_shard_bounds(10, 4, r) for r in range(4)]
    assert b == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert vocab_shard_bounds(4, 8, 7) == (4, 4)  # trailing ranks may own nothing (torch.chunk semantics)


def test_oracle_partials_roundtrip():
    rng = np.random.default_rng(1)
    T, V = 17, 333
    z = rng.standard_normal((T, V)) * 4
    ids = rng.integers(0, V, T)
    cuts = [0, 100, 100, 250, V]  # includes an empty shard
    parts = np.stack([oracle.shard_partials(z[:, a:b], ids, a) for a, b in zip(cuts[:-1], cuts[1:])])
    logp, ent, lse = oracle.merge_partials(parts)
    wl, we, wlse = oracle.logprobs_entropy(z, ids, 1.0)
    np.testing.assert_allclose(logp, wl, atol=1e-10)
    np.testing.assert_allclose(ent, we, atol=1e-10)
