"""The orchestrator -> trainer file format (rollouts/step_<n>/rank_<r>.pt) and the loader interface around it, on the CPU.
`tests/golden/rollouts/step_7/` was written by the reference's own code path and read back by its `DataLoader`
(tests/golden/make_golden.py::gen_rollout_files)."""

import os
import sys
import threading
import time

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
KEYS = ("input_ids", "position_ids", "advantages", "logprobs", "loss_mask")


@pytest.fixture(scope="module")
def io():
    from prime_rl_cse_b200 import rollout_io

    return rollout_io


def golden_rank(r):
    d = np.load(os.path.join(GOLDEN, "packed_layout.npz"))
    return [{k: d[f"rank{r}/mb{i}/{k}"] for k in KEYS} | {"temperature": float(d[f"rank{r}/mb{i}/temperature"])}
            for i in range(int(d[f"rank{r}/n"]))]


def same(mb, ref):
    return all(np.array_equal(mb[k].numpy(), ref[k]) and mb[k].dtype == torch.from_numpy(np.asarray(ref[k])).dtype for k in KEYS) \
        and float(mb["temperature"]) == pytest.approx(ref["temperature"])


@pytest.mark.parametrize("rank", [0, 1])
def test_loader_reads_the_reference_files(io, rank):
    dl = io.DataLoader(GOLDEN, 7, rank=rank)
    assert str(dl.get_rollout_path()).endswith(f"rollouts/step_7/rank_{rank}.pt")
    dl.wait_for_batch(timeout=1)
    got, want = dl.get_batch(), golden_rank(rank)
    assert len(got) == len(want) and all(same(a, b) for a, b in zip(got, want))
    assert dl.current_step == 8  # rl/data.py:74
    with pytest.raises(TimeoutError):
        dl.wait_for_batch(interval=0.01, timeout=0.05)  # step 8 was never written


def test_rank_defaults_to_the_environment(io, monkeypatch):
    monkeypatch.setenv("RANK", "1")
    assert io.DataLoader(GOLDEN, 7).get_rollout_path().name == "rank_1.pt"


def test_writer_round_trip_and_atomic_rename(io, tmp_path):
    ranks = [[{k: torch.from_numpy(np.asarray(v)) for k, v in mb.items() if k in KEYS} | {"temperature": mb["temperature"]}
              for mb in golden_rank(r)] for r in (0, 1)]
    paths = io.save_step(tmp_path, 3, ranks)
    assert [p.name for p in paths] == ["rank_0.pt", "rank_1.pt"] and not list((tmp_path / "rollouts" / "step_3").glob("*.tmp"))
    for r in (0, 1):
        got = io.DataLoader(tmp_path, 3, rank=r).get_batch()
        assert all(same(a, b) for a, b in zip(got, golden_rank(r)))
    # identical payload to the file the reference wrote
    ref = torch.load(os.path.join(GOLDEN, "rollouts", "step_7", "rank_0.pt"))
    mine = torch.load(paths[0])
    assert len(ref) == len(mine) and all(set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in KEYS) for a, b in zip(ref, mine))


def test_host_side_segment_lengths_match_the_oracle():
    """response_lengths_host (what the loader attaches so that the micro-step needs no device sync) == the oracle's restatement of
    get_response_lengths (trainer/utils.py:36-69) on the reference's fixtures and on random packed rows"""
    sys.path.insert(0, ROOT)
    import oracle
    from prime_rl_cse_b200.segments import response_lengths_host

    d = np.load(os.path.join(GOLDEN, "segments.npz"))
    for name in d["names"]:
        assert response_lengths_host(torch.from_numpy(d[f"{name}/position_ids"])) == d[f"{name}/lengths"].tolist(), name
    rng = np.random.default_rng(0)
    for _ in range(300):
        lens = rng.integers(1, 40, rng.integers(1, 12)).tolist()
        pos = np.concatenate([np.arange(n) for n in lens] + [np.zeros(rng.integers(0, 4), dtype=np.int64)])
        assert response_lengths_host(torch.from_numpy(pos).unsqueeze(0)) == oracle.response_lengths(pos)
    assert response_lengths_host(torch.zeros(0, dtype=torch.int64)) == [0] and response_lengths_host(torch.zeros(2, dtype=torch.int64)) == [2]


def test_rank_files_hold_only_their_own_tokens(io, tmp_path):
    """micro-batches that are slice VIEWS of one step-wide buffer (what prepare_batch(device_output=False) hands out) must not
    drag the whole buffer into every rank file: file size is bounded by the rank's own token count"""
    n_ranks, per_rank, T = 4, 3, 5000
    total = n_ranks * per_rank * T
    flat = {"input_ids": torch.arange(total), "position_ids": torch.arange(total) % T, "advantages": torch.randn(total),
            "logprobs": torch.randn(total), "loss_mask": torch.rand(total) < 0.5}
    ranks = []
    for r in range(n_ranks):
        mbs = []
        for i in range(per_rank):
            a = (r * per_rank + i) * T
            mbs.append({k: v[a : a + T].unsqueeze(0) for k, v in flat.items()} | {"temperature": 1.0})
        ranks.append(mbs)
    paths = io.save_step(tmp_path, 0, ranks)
    bytes_per_token = 8 + 8 + 4 + 4 + 1
    for r, p in enumerate(paths):
        assert os.path.getsize(p) < per_rank * T * bytes_per_token * 1.05 + 20000, os.path.getsize(p)
        got = io.DataLoader(tmp_path, 0, rank=r).get_batch()
        assert all(torch.equal(g[k], w[k]) for g, w in zip(got, ranks[r]) for k in KEYS)
        assert all(g[k].untyped_storage().nbytes() == g[k].numel() * g[k].element_size() for g in got for k in KEYS)


def test_wait_for_batch_sees_a_file_that_arrives_later(io, tmp_path):
    dl = io.DataLoader(tmp_path, 0, rank=0)
    mb = {k: torch.zeros(1, 4, dtype=dt) for k, dt in zip(KEYS, (torch.int64, torch.int64, torch.float32, torch.float32, torch.bool))}
    mb["temperature"] = 1.0
    t = threading.Thread(target=lambda: (time.sleep(0.2), io.save_step(tmp_path, 0, [[mb]])))
    t.start()
    dl.wait_for_batch(interval=0.02, timeout=5)
    t.join()
    assert len(dl.get_batch()) == 1


def test_malformed_files_are_rejected(io, tmp_path):
    p = io.rollout_path(tmp_path, 0, 0)
    p.parent.mkdir(parents=True)
    torch.save({"not": "a list"}, p)
    with pytest.raises(ValueError):
        io.DataLoader(tmp_path, 0, rank=0).get_batch()
    torch.save([{"input_ids": torch.zeros(1, 4, dtype=torch.int64)}], p)
    with pytest.raises(ValueError):
        io.DataLoader(tmp_path, 0, rank=0).get_batch()


def test_fake_loader_matches_the_reference_counts(io):
    fl = io.FakeDataLoader(batch_size=8, micro_batch_size=2, seq_len=16, world_size=2)
    fl.wait_for_batch()
    mbs = fl.get_batch()
    assert len(mbs) == 8 // 2 // 2 and all(tuple(mb["input_ids"].shape) == (1, 32) for mb in mbs)  # rl/data.py:30,44-47
    assert mbs[0]["position_ids"][0, :16].tolist() == list(range(16)) and mbs[0]["position_ids"][0, 16].item() == 0
    with pytest.raises(ValueError):
        io.FakeDataLoader(batch_size=3, micro_batch_size=2)


def test_single_buffer_staging_layout(io):
    mb = {k: torch.from_numpy(np.asarray(v)) for k, v in golden_rank(0)[0].items() if k in KEYS} | {"temperature": 0.6}
    buf, B, S = io.pack_host(mb, pin=False)
    assert buf.dtype == torch.uint8 and buf.numel() >= B * S * 25 and buf.numel() % 16 == 0
    back = io.unpack_device(buf, B, S, 0.6)  # the same views work on the host copy
    assert all(torch.equal(back[k], mb[k]) and back[k].dtype == mb[k].dtype for k in KEYS) and back["temperature"] == 0.6
    with pytest.raises(RuntimeError):
        next(io.upload_micro_batches([mb], "cpu"))  # no CPU path


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present (GPU box)")
def test_reference_loader_reads_our_files(io, tmp_path, monkeypatch):
    """Cross-check in the build container only: the reference's own DataLoader consumes what `save_step` wrote."""
    ranks = [[{k: torch.from_numpy(np.asarray(v)) for k, v in mb.items() if k in KEYS} | {"temperature": mb["temperature"]}
              for mb in golden_rank(r)] for r in (0, 1)]
    io.save_step(tmp_path, 5, ranks)
    code = (
        "import sys, torch; sys.path.insert(0, '/root/reference/src')\n"
        "from pathlib import Path\n"
        "from prime_rl.trainer.rl.data import DataLoader\n"
        f"b = DataLoader(Path({str(tmp_path)!r}), 5).get_batch()\n"
        "print(len(b), int(b[0]['input_ids'].sum()))\n")
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", LOCAL_WORLD_SIZE="2")
    import subprocess

    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    n, s = r.stdout.split()[-2:]
    assert int(n) == len(ranks[1]) and int(s) == int(ranks[1][0]["input_ids"].sum())
