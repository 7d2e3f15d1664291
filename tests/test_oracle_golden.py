"""Pin the CPU oracle against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py) and the known-answer vectors of SURVEY.md §8(c)."""

import os

import numpy as np
import pytest

import oracle
from oracle import LossCfg

G = os.path.join(os.path.dirname(__file__), "golden")
KINDS = ["drgrpo", "drgrpo-negclipped", "rloo", "opo"]


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


# ----------------------------------------------------------------- advantages
def test_advantages_known_answers():
    # KAT-A, SURVEY §8c (docstring of orchestrator/advantage.py:13-16,27-30)
    r, ln = [0.0, 0.0, 1.0, 1.0], [1, 2, 3, 4]
    np.testing.assert_array_equal(oracle.advantages(r, ln, 4, "drgrpo"), np.float32([-0.5, -0.5, 0.5, 0.5]))
    np.testing.assert_array_equal(oracle.advantages(r, ln, 4, "drgrpo-negclipped"), np.float32([0, 0, 0.5, 0.5]))
    np.testing.assert_array_equal(
        oracle.advantages(r, ln, 4, "rloo"), np.float32([-0.6666666865, -0.6666666865, 0.6666666865, 0.6666666865])
    )
    np.testing.assert_array_equal(
        oracle.advantages(r, ln, 4, "opo"), np.float32([-0.6999999881, -0.6999999881, 0.3000000119, 0.3000000119])
    )


def test_advantages_vs_reference_fixtures():
    d = load("advantages.npz")
    for name in d["names"]:
        r, ln, grp = d[f"{name}/rewards"], d[f"{name}/lengths"], int(d[f"{name}/group"])
        for k in KINDS:
            got = oracle.advantages(r, ln, grp, k)
            ref = d[f"{name}/{k}"]
            if name.startswith("tol_"):
                # reduction order differs from ATen's for these group sizes: <= 1e-6 rel (north_star)
                np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-6 * np.abs(r).max())
            else:
                np.testing.assert_array_equal(got, ref, err_msg=f"{name}/{k}")  # bit-exact (nan == nan)


def test_advantages_bad_group():
    with pytest.raises(AssertionError):
        oracle.advantages([1.0, 2.0, 3.0], [1, 1, 1], 2, "drgrpo")


# ------------------------------------------------------------------- segments
def test_segments_vs_reference_fixtures():
    d = load("segments.npz")
    for name in d["names"]:
        assert oracle.response_lengths(d[f"{name}/position_ids"]) == d[f"{name}/lengths"].tolist(), name
    # KAT-2
    assert oracle.response_lengths([0, 1, 2, 0, 1, 2, 3, 0, 1]) == [3, 4, 2]
    assert oracle.response_lengths([0, 1, 2, 0, 0, 1, 2]) == [7]


# ----------------------------------------------------------------------- loss
CFG = {
    "token": LossCfg("token", False, 8.0),
    "token_c2": LossCfg("token", False, 2.0),
    "seq": LossCfg("sequence", False, 8.0),
    "seqnorm": LossCfg("sequence", True, 8.0),
}


def test_loss_vs_reference_fixtures():
    d = load("loss.npz")
    n = 0
    for name in d["names"]:
        logp, old, adv, mask = (d[f"{name}/{k}"] for k in ("logp", "old", "adv", "mask"))
        lens = d[f"{name}/lens"].tolist()
        for cname in d["cfgs"]:
            keys = [k for k in d.files if k.startswith(f"{name}/{cname}/S") and k.endswith("/loss")]
            for k in keys:
                scale = int(k.split("/")[2][1:])
                pre = f"{name}/{cname}/S{scale}"
                loss, ratio, clipped, flag, dlogp = oracle.loss_fwd_bwd(logp, old, adv, mask, lens, CFG[cname], scale, np.float64)
                np.testing.assert_allclose(loss, d[pre + "/loss"], rtol=1e-5, atol=1e-6)
                np.testing.assert_allclose(ratio, d[pre + "/ratio"], rtol=2e-5)
                np.testing.assert_allclose(clipped, d[pre + "/clipped"], rtol=2e-5)
                np.testing.assert_array_equal(flag, d[pre + "/is_clipped"])
                np.testing.assert_allclose(dlogp, d[pre + "/dlogp"], rtol=2e-5, atol=1e-7)
                n += 1
    assert n >= 40


def test_loss_known_answers():
    # KAT-3 / KAT-4
    logp, old = [-1, -2, -0.5, -3, -0.1], [-1.5, -1, -3, -3, -0.2]
    adv, mask, lens = [0.5, 0.5, 0.5, -1, -1], [1, 1, 1, 0, 1], [3, 2]
    loss, ratio, clipped, flag, dlogp = oracle.loss_fwd_bwd(logp, old, adv, mask, lens, CFG["token"], 4)
    assert abs(loss - (-0.9757823348)) < 1e-6
    np.testing.assert_allclose(ratio, [1.6487212, 0.3678795, 12.1824942, 1.0, 1.1051710], rtol=1e-6)
    assert clipped[2] == 8.0 and flag.tolist() == [0, 0, 1, 0, 0]
    np.testing.assert_allclose(dlogp, [-0.2060902, -0.0459849, 0, 0, 0.2762927], rtol=1e-5, atol=1e-7)
    loss, ratio, *_, dlogp = oracle.loss_fwd_bwd(logp, old, adv, mask, lens, CFG["seq"], 2)
    assert abs(loss - (-1.2946784496)) < 1e-6
    np.testing.assert_allclose(ratio, [7.3890562, 1.1051710], rtol=1e-6)
    np.testing.assert_allclose(dlogp, [-1.8472641] * 3 + [0, 0.5525855], rtol=1e-5)
    loss, ratio, *_, dlogp = oracle.loss_fwd_bwd(logp, old, adv, mask, lens, CFG["seqnorm"], 2)
    assert abs(loss - 0.0656519234) < 1e-6
    np.testing.assert_allclose(dlogp, [-0.1623112] * 3 + [0, 0.5525855], rtol=1e-5)
    # loss_scale = 0 behaves as 1
    l0 = oracle.loss_fwd_bwd(logp, old, adv, mask, lens, CFG["token"], 0)[0]
    l1 = oracle.loss_fwd_bwd(logp, old, adv, mask, lens, CFG["token"], 1)[0]
    assert l0 == l1


# -------------------------------------------------------------------- LM head
def test_kat1_shift_logp_entropy():
    d = load("lmhead.npz")
    logits, ids = d["kat1/logits"], d["kat1/ids"]
    for tau in (1.0, 2.0):
        lp, ent, _ = oracle.logprobs_entropy(oracle.shift_rows(logits), ids, tau)
        np.testing.assert_allclose(lp, d[f"kat1/tau{tau}/logp"], atol=1e-6)
        np.testing.assert_allclose(ent, d[f"kat1/tau{tau}/entropy"], atol=1e-6)
    lp, ent, _ = oracle.logprobs_entropy(oracle.shift_rows(logits), ids, 1.0)
    np.testing.assert_allclose(lp, [-1.3862944, -0.4401898, -1.3862944], atol=1e-6)
    np.testing.assert_allclose(ent, [1.3862944, 0.9475369, 1.3862944], atol=1e-6)


def rel_l2(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_lmhead_chain_vs_reference_fixtures():
    d = load("lmhead.npz")
    for key in d["names"]:
        name, tau_s, cname = key.split("/")
        tau = float(tau_s[3:])
        is_bf16 = bool(d[f"{name}/is_bf16"])
        cfg = LossCfg() if cname == "token" else LossCfg("sequence", True, 8.0)
        res = oracle.micro_step(
            d[f"{name}/hidden"], d[f"{name}/weight"], d[f"{name}/ids"], d[f"{name}/pos"], d[f"{key}/old"],
            d[f"{name}/adv"], d[f"{name}/mask"], tau, cfg, int(d[f"{key}/scale"]), round_bf16=is_bf16,
        )
        assert res["lengths"] == d[f"{key}/lens"].tolist()
        # bf16 CPU matmul accumulates in a different order than fp64: a logit can land on the
        # other side of a bf16 rounding boundary (<= 1 bf16 ulp of the logit) -> looser atol.
        atol = 3e-2 if is_bf16 else 2e-5
        np.testing.assert_allclose(res["logp"], d[f"{key}/logp"], atol=atol)
        np.testing.assert_allclose(res["entropy"], d[f"{key}/entropy"], atol=atol)
        assert np.mean(np.abs(res["logp"] - d[f"{key}/logp"]) < 1e-3) > 0.97
        np.testing.assert_allclose(res["loss"], d[f"{key}/loss"], rtol=5e-3 if is_bf16 else 1e-4, atol=1e-5)
        assert rel_l2(res["d_hidden"], d[f"{key}/d_hidden"]) < 1e-2
        assert rel_l2(res["d_weight"], d[f"{key}/d_weight"]) < 1e-2


def test_bf16_round_matches_torch():
    import torch

    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 37
    ref = torch.from_numpy(x).to(torch.bfloat16).float().numpy()
    np.testing.assert_array_equal(oracle.bf16_round(x), ref)


def test_sft_cross_entropy_vs_reference_call_site():
    d = load("sft.npz")
    for name in d["names"]:
        bf = bool(d[f"{name}/is_bf16"])
        loss, dh, dw = oracle.sft_cross_entropy(d[f"{name}/hidden"][0], d[f"{name}/weight"], d[f"{name}/targets"], d[f"{name}/mask"], bf)
        np.testing.assert_allclose(loss, d[f"{name}/loss"], rtol=2e-2 if bf else 1e-5)  # the bf16 reference loss has 8 mantissa bits
        assert rel_l2(dh, d[f"{name}/d_hidden"][0]) < (2e-2 if bf else 1e-5)
        assert rel_l2(dw, d[f"{name}/d_weight"]) < (2e-2 if bf else 1e-5)


def test_packed_layout_from_reference_prepare_batch():
    """KAT-5: micro-batches produced by the reference orchestrator (prepare_batch) — layout facts the path relies on."""
    d = load("packed_layout.npz")
    assert int(d["n_ranks"]) == 2
    np.testing.assert_array_equal(oracle.advantages(d["rewards"], d["lens"], 4, "drgrpo"), d["advantages"])
    n_pad = 0
    for r in range(2):
        for i in range(int(d[f"rank{r}/n"])):
            pre = f"rank{r}/mb{i}"
            pos, mask, adv = d[f"{pre}/position_ids"][0], d[f"{pre}/loss_mask"][0], d[f"{pre}/advantages"][0]
            lens = oracle.response_lengths(pos)
            assert lens == d[f"{pre}/lengths"].tolist() and sum(lens) == pos.shape[0]
            assert lens == sorted(lens, reverse=True)  # first-fit-decreasing keeps each bin longest-first
            o = 0
            for n in lens:  # one advantage per rollout, broadcast over prompt AND completion (batch.py:44)
                assert np.all(adv[o : o + n] == adv[o]) and not mask[o]  # the first (prompt) token is never trained on
                o += n
            if not mask.any():
                n_pad += 1
                assert not adv.any()  # padding micro-batch: zero advantages, all-False mask (batch.py:153-161)
    assert n_pad == 1


def test_first_fit_decreasing_matches_reference_bins():
    """Host half of N4: the bin assignment reproduces the reference's packed micro-batches (segment lengths per bin)."""
    from prime_rl_cse_b200.batch import first_fit_decreasing

    d = load("packed_layout.npz")
    lens = (d["ro/prompt_len"] + d["ro/completion_len"]).tolist()
    temperature, batch_size, mbs, seq_len, nw = d["ro/args"]
    bins = first_fit_decreasing(lens, int(seq_len) * int(mbs))
    got = [[lens[i] for i in b] for b in bins]
    want = [d["rank0/mb0/lengths"].tolist(), d["rank0/mb1/lengths"].tolist(), d["rank1/mb0/lengths"].tolist()]
    assert got == want


def test_grpo_std_extension_matches_its_torch_definition():
    """"grpo" is NOT in the reference (its registry is Dr.GRPO: mean only); the product offers it because north_star names
    "mean/std".  Its definition -- (r - mean) / (std + 1e-6), unbiased std -- is pinned here against plain torch."""
    import torch

    assert "grpo" not in oracle.ADVANTAGE_KINDS and oracle.EXTENSION_ADVANTAGE_KINDS == ("grpo",)
    rng = np.random.default_rng(0)
    for G in (2, 5, 8, 16, 24, 32, 40):
        r = rng.standard_normal(G * 9).astype(np.float32)
        t = torch.from_numpy(r).view(-1, G)
        want = ((t - t.mean(1, keepdim=True)) / (t.std(1, keepdim=True) + 1e-6)).reshape(-1).numpy()
        np.testing.assert_allclose(oracle.advantages(r, np.ones(len(r), dtype=np.int64), G, "grpo"), want, rtol=2e-6, atol=1e-6)
    np.testing.assert_array_equal(oracle.advantages([1.0] * 4, [1] * 4, 4, "grpo"), np.zeros(4, dtype=np.float32))  # zero variance
    assert np.isnan(oracle.advantages([3.0], [1], 1, "grpo")).all()  # torch.std of one element is nan
    with pytest.raises(KeyError):
        oracle.advantages([1.0, 2.0], [1, 1], 2, "gae")


# ----------------------------------------------------------------- same-logits checker (tests/same_logits.py)
def test_torch_checker_matches_numpy_oracle():
    """the torch fp64 restatement used for device-side checks at the BASELINE shapes == the numpy oracle"""
    import torch

    from tests.same_logits import bf16_ulp, chain_from_block, logprobs_entropy_t, rn_bf16_violations

    rng = np.random.default_rng(0)
    T, H, V, tau = 97, 48, 531, 0.6
    hidden = oracle.bf16_round(rng.standard_normal((T, H)).astype(np.float32))
    weight = oracle.bf16_round((rng.standard_normal((V, H)) * 0.4).astype(np.float32))
    ids = rng.integers(0, V, T)
    logits = oracle.lm_head_logits(hidden, weight, True)
    wl, we, wlse = oracle.logprobs_entropy(oracle.shift_rows(logits), ids, tau)
    lp, en, ls = logprobs_entropy_t(torch.from_numpy(oracle.shift_rows(logits)), torch.from_numpy(ids), tau)
    np.testing.assert_allclose(lp.numpy(), wl, rtol=0, atol=1e-12)
    np.testing.assert_allclose(en.numpy(), we, rtol=0, atol=1e-12)
    np.testing.assert_allclose(ls.numpy(), wlse, rtol=0, atol=1e-12)
    # block form (un-shifted rows, padded pitch), chunked
    blk = torch.zeros(T, 576, dtype=torch.bfloat16)
    blk[:, :V] = torch.from_numpy(logits.astype(np.float32)).bfloat16()
    lp2, en2, ls2 = chain_from_block(blk, V, torch.from_numpy(ids), tau, shift=True, chunk=13)
    np.testing.assert_allclose(lp2.numpy(), wl, rtol=0, atol=1e-12)
    np.testing.assert_allclose(en2.numpy(), we, rtol=0, atol=1e-12)
    # ulp helper and the rounding check: the oracle's own block has no violation; a block that is 1 ulp off has
    assert bf16_ulp(torch.tensor([1.0, 1.5, 2.0, 0.75, -3.0])).tolist() == [2.0**-7, 2.0**-7, 2.0**-6, 2.0**-8, 2.0**-6]
    h_t, w_t = torch.from_numpy(hidden).bfloat16(), torch.from_numpy(weight).bfloat16()
    rows = torch.arange(T)
    ok = rn_bf16_violations(blk, h_t, w_t, rows, chunk=40)
    assert ok["n"] == T * V and ok["violations"] == 0 and ok["flips"] == 0
    off = blk.clone()
    off[5, 7] = (off[5, 7].double() + bf16_ulp(off[5, 7].double())).to(torch.bfloat16)
    bad = rn_bf16_violations(off, h_t, w_t, rows, verbose=False)
    assert bad["violations"] == 1 and bad["flips"] == 1 and bad["noise_budget_used"] > 1.0 and ok["noise_budget_used"] <= 0.0


def test_micro_step_from_given_logits_equals_recomputed():
    """oracle.micro_step(logits=...) with the oracle's own logits == oracle.micro_step()"""
    rng = np.random.default_rng(1)
    T, H, V, tau = 60, 32, 200, 0.8
    hidden = oracle.bf16_round(rng.standard_normal((T, H)).astype(np.float32))
    weight = oracle.bf16_round((rng.standard_normal((V, H)) * 0.3).astype(np.float32))
    ids = rng.integers(0, V, T)
    pos = np.concatenate([np.arange(25), np.arange(35)])
    old = -rng.random(T).astype(np.float32) * 5
    adv = rng.standard_normal(T).astype(np.float32)
    mask = rng.random(T) < 0.7
    a = oracle.micro_step(hidden, weight, ids, pos, old, adv, mask, tau, LossCfg(), int(mask.sum()))
    b = oracle.micro_step(hidden, weight, ids, pos, old, adv, mask, tau, LossCfg(), int(mask.sum()),
                          logits=oracle.lm_head_logits(hidden, weight, True))
    for k in ("logp", "entropy", "loss", "dlogp", "d_hidden", "d_weight"):
        np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]))
