"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  The CUDA path is called through the
C-ABI library and compared with the CPU oracle (`oracle/`) and with the golden fixtures produced by
the unmodified reference.  Tolerances are the ones BASELINE.json's north_star states:
advantages bit-exact; logp / entropy <= 1e-3 abs; loss <= 1e-4 rel; gradients <= 1e-2 rel-L2.

Fused LM-head path: the stated tolerances are held for EVERY token against the oracle started from the kernel's own bf16
logits (`tests/same_logits.py`: what is left out is only which way an accumulator on a bf16 rounding midpoint rounds, and
`rn_bf16_violations` proves the kernel's logits are the correctly rounded product up to accumulation noise).  Against the
exact-product oracle / the reference's fixtures the log-probs are within 1e-3 for >= 99.9 % of tokens (at most one token of a
row shorter than 1000: `allowed_flip_outliers`); tests/test_gpu_same_logits.py does the same at the BASELINE shapes."""

import math
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import LossCfg
from tests.same_logits import allowed_flip_outliers, assert_same_logits_micro_step, kernel_logits, rn_bf16_violations

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
KINDS = ["drgrpo", "drgrpo-negclipped", "rloo", "opo"]


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(4321)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(4321)


@pytest.fixture(scope="module")
def P():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import prime_rl_cse_b200 as p

    return p


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def cu(x, dtype=None):
    t = torch.as_tensor(np.asarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


# ------------------------------------------------------------------------------- advantages
def test_advantages_fixtures_bit_exact(P):
    d = load("advantages.npz")
    for name in d["names"]:
        r, ln, grp = d[f"{name}/rewards"], d[f"{name}/lengths"], int(d[f"{name}/group"])
        for k in KINDS:
            got = np.float32(P.compute_advantages(r.tolist(), ln.tolist(), grp, k))
            np.testing.assert_array_equal(got, oracle.advantages(r, ln, grp, k), err_msg=f"{name}/{k} vs oracle")
            if name.startswith("tol_"):
                np.testing.assert_allclose(got, d[f"{name}/{k}"], rtol=1e-6, atol=1e-6 * np.abs(r).max())
            else:
                np.testing.assert_array_equal(got, d[f"{name}/{k}"], err_msg=f"{name}/{k} vs reference")


def test_advantages_large_and_broadcast(P):
    rng = np.random.default_rng(5)
    for grp, n in [(16, 1 << 16), (32, 1 << 12), (7, 7 * 100), (1, 5), (8, 8 * 1001), (24, 24 * 333), (16, 16 * 3), (32, 32)]:
        r = rng.standard_normal(n).astype(np.float32)
        ln = rng.integers(1, 5000, n)
        for k in KINDS:
            got = P.compute_advantages_tensor(cu(r), cu(ln), grp, k).cpu().numpy()
            np.testing.assert_array_equal(got, oracle.advantages(r, ln, grp, k))
    # group sizes 8 / 16 / 24 / 32 take the 8-rollouts-per-lane kernel (16-byte loads) when the buffers are 16-byte aligned;
    # a view that is only 4-byte aligned must give the same bits through the generic kernel
    r = rng.standard_normal(16 * 50 + 1).astype(np.float32)
    ln = rng.integers(1, 5000, 16 * 50 + 1)
    for k in KINDS:
        got = P.compute_advantages_tensor(cu(r)[1:], cu(ln)[1:], 16, k).cpu().numpy()
        np.testing.assert_array_equal(got, oracle.advantages(r[1:], ln[1:], 16, k))
    # fused per-token broadcast (orchestrator/batch.py:44)
    n, grp = 64, 8
    r = rng.standard_normal(n).astype(np.float32)
    ln = rng.integers(1, 50, n)
    offs = np.concatenate([[0], np.cumsum(ln)])
    adv, tok = P.compute_advantages_tensor(cu(r), cu(ln), grp, "drgrpo", token_offsets=cu(offs))
    np.testing.assert_array_equal(tok.cpu().numpy(), np.repeat(adv.cpu().numpy(), ln))


def test_advantages_grpo_std_extension(P):
    """the std-normalised extension (not in the reference): bit-exact against the oracle's definition through all three
    kernels (8-rollouts-per-lane, small groups, one warp per group), close to torch's expression, per-token broadcast included"""
    rng = np.random.default_rng(11)
    for grp, n in [(8, 8 * 300), (16, 16 * 1000), (24, 24 * 50), (32, 32 * 64), (5, 5 * 77), (12, 12 * 40), (40, 40 * 30), (2, 2 * 9)]:
        r = rng.standard_normal(n).astype(np.float32)
        ln = rng.integers(1, 100, n)
        got = P.compute_advantages_tensor(cu(r), cu(ln), grp, "grpo").cpu().numpy()
        np.testing.assert_array_equal(got, oracle.advantages(r, ln, grp, "grpo"))
        t = torch.from_numpy(r).view(-1, grp)
        want = ((t - t.mean(1, keepdim=True)) / (t.std(1, keepdim=True) + 1e-6)).reshape(-1).numpy()
        np.testing.assert_allclose(got, want, rtol=5e-6, atol=2e-6)
    n, grp = 48, 16
    r = rng.standard_normal(n).astype(np.float32)
    ln = rng.integers(1, 30, n)
    offs = np.concatenate([[0], np.cumsum(ln)])
    adv, tok = P.compute_advantages_tensor(cu(r), cu(ln), grp, "grpo", token_offsets=cu(offs))
    np.testing.assert_array_equal(adv.cpu().numpy(), oracle.advantages(r, ln, grp, "grpo"))
    np.testing.assert_array_equal(tok.cpu().numpy(), np.repeat(adv.cpu().numpy(), ln))
    assert P.compute_advantages([1.0, 1.0, 1.0, 1.0], [1, 2, 3, 4], 4, "grpo") == [0.0, 0.0, 0.0, 0.0]


def test_advantages_errors(P):
    with pytest.raises(AssertionError):
        P.compute_advantages([1.0, 2.0, 3.0], [1, 1, 1], 2, "drgrpo")
    with pytest.raises(KeyError):
        P.compute_advantages([1.0, 2.0], [1, 1], 2, "gae")
    assert P.compute_advantages([], [], 4, "drgrpo") == []


# ------------------------------------------------------------------------------- segments
def test_segments_fixtures(P):
    d = load("segments.npz")
    for name in d["names"]:
        got = P.get_response_lengths(cu(d[f"{name}/position_ids"]))
        assert got == d[f"{name}/lengths"].tolist(), name


def test_segments_long_rows(P):
    rng = np.random.default_rng(3)
    for T_target in (4096, 16384, 32768 + 77):
        lens, tot = [], 0
        while tot < T_target:
            n = int(rng.integers(1, 3000))
            lens.append(n)
            tot += n
        pos = np.concatenate([np.arange(n) for n in lens] + [np.zeros(int(rng.integers(0, 9)), dtype=np.int64)])
        assert P.get_response_lengths(cu(pos).unsqueeze(0)) == oracle.response_lengths(pos)


def test_segments_more_than_the_first_pass_holds(P):
    """get_response_lengths fetches a table of at most 4095 segments in its one D2H copy and falls back to a full-size pass
    for rows with more (9000 two-token rollouts here; also the exact limit, and T < 3)"""
    for nseg in (4094, 4095, 4096, 9000):
        pos = np.tile(np.array([0, 1], dtype=np.int64), nseg)
        want = oracle.response_lengths(pos)
        assert len(want) == nseg and P.get_response_lengths(cu(pos).unsqueeze(0)) == want
    for pos in ([0], [0, 1], [0, 0]):
        assert P.get_response_lengths(cu(np.array(pos, dtype=np.int64)).unsqueeze(0)) == oracle.response_lengths(np.array(pos))


# ------------------------------------------------------------------------------- loss
CFG = {
    "token": ("token", False, 8.0),
    "token_c2": ("token", False, 2.0),
    "seq": ("sequence", False, 8.0),
    "seqnorm": ("sequence", True, 8.0),
}


@pytest.mark.parametrize("as_views", [True, False])
def test_loss_fixtures(P, as_views):
    d = load("loss.npz")
    for name in d["names"]:
        logp, old, adv, mask = (d[f"{name}/{k}"] for k in ("logp", "old", "adv", "mask"))
        lens = d[f"{name}/lens"].tolist()
        for cname in d["cfgs"]:
            cfg = P.LossConfig(*CFG[cname])
            for k in [k for k in d.files if k.startswith(f"{name}/{cname}/S") and k.endswith("/loss")]:
                scale = int(k.split("/")[2][1:])
                pre = f"{name}/{cname}/S{scale}"
                lp = cu(logp).requires_grad_(True)
                if as_views:  # what the trainer passes: split views of one tensor (train.py:293-296)
                    items = [lp.split(lens), cu(old).split(lens), cu(adv).split(lens), cu(mask).split(lens)]
                else:  # what the reference's unit tests pass: independent tensors
                    items = [[t.clone() for t in x.split(lens)] for x in (lp, cu(old), cu(adv), cu(mask))]
                loss, tens = P.compute_loss(*items, cfg, scale)
                assert loss.shape == ()
                loss.backward()
                np.testing.assert_allclose(loss.item(), d[pre + "/loss"], rtol=1e-4, atol=1e-6, err_msg=pre)
                np.testing.assert_allclose(tens["importance_ratio"].cpu().numpy(), d[pre + "/ratio"], rtol=1e-4)
                np.testing.assert_allclose(tens["clipped_importance_ratio"].cpu().numpy(), d[pre + "/clipped"], rtol=1e-4)
                np.testing.assert_array_equal(tens["is_clipped"].cpu().numpy(), d[pre + "/is_clipped"])
                np.testing.assert_allclose(lp.grad.cpu().numpy(), d[pre + "/dlogp"], rtol=1e-4, atol=1e-7, err_msg=pre)


def test_loss_reference_unit_test_shapes(P):
    # tests/unit/train/rl/test_loss.py:10-42 of the reference, with values checked against the oracle
    torch.manual_seed(0)
    for ratio_type, lens in (("token", [50, 30]), ("sequence", [40, 60])):
        logprobs = [torch.randn(n).cuda() for n in lens]
        old = [torch.randn(n).cuda() for n in lens]
        adv = [torch.randn(n).cuda() for n in lens]
        mask = [torch.ones(n, dtype=torch.bool).cuda() for n in lens]
        loss, _ = P.compute_loss(logprobs, old, adv, loss_mask=mask, loss_config=P.LossConfig(ratio_type=ratio_type, clip_ratio=10.0),
                                 loss_scale=1.0)
        assert loss.shape == ()
        want = oracle.loss_fwd_bwd(torch.cat(logprobs).cpu().numpy(), torch.cat(old).cpu().numpy(), torch.cat(adv).cpu().numpy(),
                                   torch.cat(mask).cpu().numpy(), lens, LossCfg(ratio_type, False, 10.0), 1.0, np.float64)[0]
        np.testing.assert_allclose(loss.item(), want, rtol=1e-4)


@pytest.mark.parametrize("ratio_type", ["token", "sequence"])
def test_loss_nan_propagation_matches_torch(P, ratio_type):
    """A NaN log-prob must surface exactly where the reference's torch ops put it (torch.clamp propagates NaN, fminf does not):
    NaN loss when the token is unmasked, finite loss when it is masked out, NaN gradient at the token itself either way in
    token mode (autograd: 0 * exp'(nan)).  Checked against the same torch ops under autograd (oracle/torch_port.py)."""
    from oracle import torch_port

    torch.manual_seed(3)
    lens = [40, 57, 31]
    T = sum(lens)
    old, adv = torch.randn(T).cuda(), torch.randn(T).cuda()
    for where, masked_out in ((5, False), (60, True), (T - 1, False)):
        base = (old + 0.1 * torch.randn(T).cuda())
        base[where] = float("nan")
        mask = torch.rand(T).cuda() < 0.8
        mask[where] = not masked_out
        lp_ref = base.clone().requires_grad_(True)
        want = torch_port.loss_cpu(lp_ref, old, adv, mask, lens, ratio_type, ratio_type == "sequence", 8.0, int(mask.sum()))
        want.backward()
        lp = base.clone().requires_grad_(True)
        loss, tens = P.compute_loss(lp.split(lens), old.split(lens), adv.split(lens), mask.split(lens),
                                    P.LossConfig(ratio_type, ratio_type == "sequence", 8.0), int(mask.sum()))
        loss.backward()
        assert math.isnan(loss.item()) == math.isnan(want.item()) == (not masked_out), (where, loss.item(), want.item())
        assert torch.equal(torch.isnan(lp.grad), torch.isnan(lp_ref.grad)), (ratio_type, where)
        ok = ~torch.isnan(lp_ref.grad)
        np.testing.assert_allclose(lp.grad[ok].cpu().numpy(), lp_ref.grad[ok].cpu().numpy(), rtol=1e-4, atol=1e-7)
        if masked_out:
            np.testing.assert_allclose(loss.item(), want.item(), rtol=1e-4)


def test_raw_wrappers_reject_operands_the_kernels_would_misread(P):
    """an fp32 / non-contiguous / mis-shaped operand must raise instead of being reinterpreted as bf16 by the TMA descriptors"""
    T, H, V = 64, 64, 512
    h, w, i = torch.randn(T, H).cuda().bfloat16(), torch.randn(V, H).cuda().bfloat16(), torch.randint(0, V, (T,)).cuda()
    lse, g = torch.zeros(T).cuda(), torch.zeros(T).cuda()
    bad = [(h.float(), w, i), (h, w.float(), i), (h.t().contiguous().t(), w, i), (h, w[:, :32], i), (h, w, i.int()), (h, w, i[:-1]),
           (h.cpu(), w, i)]
    for hh, ww, ii in bad:
        with pytest.raises((ValueError, RuntimeError)):
            P.lm_head_forward_raw(hh, ww, ii, 1.0)
        with pytest.raises((ValueError, RuntimeError)):
            P.lm_head_backward_raw(hh, ww, ii, lse, g, 1.0)
    with pytest.raises(ValueError):
        P.lm_head_backward_raw(h, w, i, lse.double(), g, 1.0)
    with pytest.raises(ValueError):
        P.recompute_logprobs(h.float().unsqueeze(0), w, i.unsqueeze(0), 1.0)
    P.lm_head_forward_raw(h, w, i, 1.0)  # the well-formed call goes through


def test_logits_stats_nan_and_minus_inf_like_torch(P):
    """materialised-logits helpers: -inf logits are fine for the log-prob (torch.logsumexp) and make the entropy NaN
    (softmax * z = 0 * -inf, rl/loss.py:45-46); a NaN logit makes both NaN -- as the reference's torch ops do"""
    torch.manual_seed(0)
    z = torch.randn(1, 6, 300).cuda()
    z[0, 1, 0] = float("-inf")      # the very first element a thread sees
    z[0, 2, 17] = float("-inf")
    z[0, 3, 100] = float("nan")
    ids = torch.randint(1, 300, (1, 6)).cuda()
    lp = P.selective_log_softmax(z, ids)
    ent = P.compute_entropy(z)
    ref_lp = torch.gather(z, -1, ids.unsqueeze(-1)).squeeze(-1) - torch.logsumexp(z, -1)
    pd = torch.softmax(z, -1)
    ref_ent = torch.logsumexp(z, -1) - (pd * z).sum(-1)
    assert torch.equal(torch.isnan(lp), torch.isnan(ref_lp)) and torch.equal(torch.isnan(ent), torch.isnan(ref_ent))
    ok = ~torch.isnan(ref_lp)
    assert (lp[ok] - ref_lp[ok]).abs().max().item() < 1e-5
    ok = ~torch.isnan(ref_ent)
    assert (ent[ok] - ref_ent[ok]).abs().max().item() < 1e-5
    # logging moments: torch.min / torch.max propagate NaN
    st = P.TensorStats(median=False)
    st.add("x", torch.tensor([1.0, float("nan"), 3.0]).cuda())
    got = st.compute_stats()
    assert math.isnan(got["x/min"]) and math.isnan(got["x/max"]) and math.isnan(got["x/mean"])


def test_loss_all_masked_and_big(P):
    # padding micro-batches: all-False mask, zero advantages, loss_scale 0 (orchestrator/batch.py:153-161)
    T = 1000
    z = torch.zeros(T).cuda()
    loss, tens = P.compute_loss([z.clone().requires_grad_(True)], [z], [z], [torch.zeros(T, dtype=torch.bool).cuda()], P.LossConfig(), 0)
    assert loss.item() == 0.0 and tens["importance_ratio"].shape == (T,)
    # config-5 sized row, sequence mode, one 32k segment; token mode 2^20 tokens
    rng = np.random.default_rng(11)
    for mode, T, lens in (("sequence", 32768, [32768]), ("token", 1 << 20, [1 << 19, (1 << 19) - 3, 3])):
        logp = -np.abs(rng.standard_normal(T)).astype(np.float32)
        old = (logp + 0.02 * rng.standard_normal(T) / (50 if mode == "sequence" else 1)).astype(np.float32)
        adv = rng.standard_normal(T).astype(np.float32)
        mask = rng.random(T) < 0.8
        cfg = P.LossConfig(mode, True, 8.0)
        lp = cu(logp).requires_grad_(True)
        loss, tens = P.compute_loss(lp.split(lens), cu(old).split(lens), cu(adv).split(lens), cu(mask).split(lens), cfg, int(mask.sum()))
        loss.backward()
        want = oracle.loss_fwd_bwd(logp, old, adv, mask, lens, LossCfg(mode, True, 8.0), int(mask.sum()), np.float64)
        np.testing.assert_allclose(loss.item(), want[0], rtol=1e-4)
        np.testing.assert_allclose(tens["importance_ratio"].cpu().numpy(), want[1], rtol=1e-4)
        assert rel_l2(lp.grad.cpu().numpy(), want[4]) < 1e-4


# ------------------------------------------------------------------------------- materialised-logits helpers
def test_logits_stats_vs_oracle(P):
    d = load("lmhead.npz")
    logits, ids = d["kat1/logits"], d["kat1/ids"]
    for tau in (1.0, 2.0):
        lp, ent = P.shifted_logprobs_entropy(cu(logits).unsqueeze(0), cu(ids).unsqueeze(0), tau)
        np.testing.assert_allclose(lp[0].cpu().numpy(), d[f"kat1/tau{tau}/logp"], atol=1e-5)
        np.testing.assert_allclose(ent[0].cpu().numpy(), d[f"kat1/tau{tau}/entropy"], atol=1e-5)
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((3, 17, 5003)) * 3).astype(np.float32)
    ids = rng.integers(0, 5003, (3, 17))
    lp = P.selective_log_softmax(cu(x), cu(ids)).cpu().numpy()
    ent = P.compute_entropy(cu(x)).cpu().numpy()
    assert ent.shape == (3, 17)
    for b in range(3):
        wl, we, _ = oracle.logprobs_entropy(x[b], ids[b], 1.0)
        np.testing.assert_allclose(lp[b], wl, atol=1e-4)
        np.testing.assert_allclose(ent[b], we, atol=1e-4)
    sh = P.shift_logits(cu(x)).cpu().numpy()
    for b in range(3):
        np.testing.assert_array_equal(sh[b], oracle.shift_rows(x[b]))


# ------------------------------------------------------------------------------- tcgen05 GEMM engine
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 520, 200), (1024, 768, 4096)])
def test_gemm_engine(P, a_mn, b_mn, M, N, K):
    from prime_rl_cse_b200 import _lib

    lib = _lib.load()
    torch.manual_seed(M + N + K)
    # pitches must be multiples of 8 elements for TMA; pad logical sizes where needed
    M8, N8, K8 = (M + 7) // 8 * 8, (N + 7) // 8 * 8, (K + 7) // 8 * 8
    A = torch.randn(M8, K8, device="cuda").bfloat16()
    B = torch.randn(N8, K8, device="cuda").bfloat16()
    a_st = A.t().contiguous() if a_mn else A
    b_st = B.t().contiguous() if b_mn else B
    C = torch.full((M8, N8), float("nan"), device="cuda")
    rc = lib.prl_gemm_bf16_f32(_lib.ptr(a_st), _lib.ptr(b_st), _lib.ptr(C), M8, N8, K8, a_mn, b_mn, 0, _lib.stream())
    _lib.check(rc, "gemm")
    want = A.float() @ B.float().t()
    torch.cuda.synchronize()
    err = (C - want).abs().max().item()
    assert err <= 2e-3 * max(1.0, want.abs().max().item()), f"max err {err}"
    rc = lib.prl_gemm_bf16_f32(_lib.ptr(a_st), _lib.ptr(b_st), _lib.ptr(C), M8, N8, K8, a_mn, b_mn, 1, _lib.stream())
    _lib.check(rc, "gemm acc")
    assert (C - 2 * want).abs().max().item() <= 4e-3 * max(1.0, want.abs().max().item())


def test_reference_trainer_chain_with_drop_in_functions_is_differentiable(P):
    """the reference's own chain (train.py:285-310) with ONLY the imports swapped — materialised logits,
    `shift_logits`, `/ temperature`, `selective_log_softmax`, `compute_loss`, `compute_entropy` — gives the torch chain's loss and
    gradients (`selective_log_softmax` is differentiable w.r.t. the logits)."""
    from oracle import torch_port

    T, H, V, tau = 300, 64, 1000, 0.7
    hidden, weight, ids = _mk(T, H, V, 77, peaked=True)
    rng = np.random.default_rng(2)
    lens = [120, 100, 80]
    pos = cu(np.concatenate([np.arange(n) for n in lens])).unsqueeze(0)
    adv, mask = cu(rng.standard_normal(T).astype(np.float32)), cu(rng.random(T) < 0.8)
    old = cu(-rng.random(T).astype(np.float32) * 3)
    i_d = ids.cuda()
    h0 = hidden.cuda().requires_grad_(True)
    w0 = weight.cuda().requires_grad_(True)
    loss0, lp0, ent0 = torch_port.micro_step_cpu(h0, w0, i_d, pos[0], old, adv, mask, tau, "token", False, 8.0, int(mask.sum()))

    for variant in ("reference-calls", "one-pass"):
        h = hidden.cuda().requires_grad_(True)
        w = weight.cuda().requires_grad_(True)
        logits = (h @ w.t()).float().unsqueeze(0).contiguous()                         # train.py:285
        if variant == "reference-calls":
            shifted = P.shift_logits(logits) / tau                                     # train.py:286-287
            lp = P.selective_log_softmax(shifted, i_d.unsqueeze(0))                    # train.py:288
            with torch.no_grad():
                ent = P.compute_entropy(shifted)                                       # train.py:302-303
        else:
            lp, ent = P.shifted_logprobs_entropy(logits, i_d.unsqueeze(0), tau)        # the same three steps in one pass
        seg = P.get_response_lengths(pos)                                              # train.py:291
        assert seg == lens
        loss, _ = P.compute_loss(lp.squeeze().split(seg), old.split(seg), adv.split(seg), mask.split(seg), P.LossConfig(), int(mask.sum()))
        loss.backward()                                                                # train.py:310
        assert (lp[0] - lp0).abs().max().item() < 1e-4 and (ent[0] - ent0).abs().max().item() < 1e-4
        assert abs(loss.item() - loss0.item()) <= 1e-4 * abs(loss0.item()) + 1e-7
        assert rel_l2(h.grad.float().cpu().numpy(), h0.grad.float().cpu().numpy()) < 2e-3, variant
        assert rel_l2(w.grad.float().cpu().numpy(), w0.grad.float().cpu().numpy()) < 2e-3, variant
    # low-precision logits (the reference's bf16 branch, loss.py:31-38): gradient comes back in the logits' dtype
    z = torch.randn(2, 9, 513, device="cuda").bfloat16().requires_grad_(True)
    idx = torch.randint(0, 513, (2, 9), device="cuda")
    gw = torch.randn(2, 9, device="cuda")
    (P.selective_log_softmax(z, idx) * gw).sum().backward()
    zr = z.detach().float().requires_grad_(True)
    (torch.log_softmax(zr, -1).gather(-1, idx.unsqueeze(-1)).squeeze(-1) * gw).sum().backward()
    assert z.grad.dtype == torch.bfloat16 and (z.grad.float() - zr.grad).abs().max().item() <= 2.0 ** -8 * zr.grad.abs().max().item() + 1e-6


# ------------------------------------------------------------------------------- fused LM head
def _mk(T, H, V, seed, peaked=False):
    g = torch.Generator().manual_seed(seed)
    hidden = torch.randn(T, H, generator=g).bfloat16()
    weight = (torch.randn(V, H, generator=g) * (3.0 / H**0.5)).bfloat16()
    ids = torch.randint(0, V, (T,), generator=g)
    if peaked:  # make the sampled token likely: add its embedding direction to the previous hidden state
        hidden[:-1] = (hidden[:-1].float() + 4.0 * weight[ids[1:]].float() / weight[ids[1:]].float().norm(dim=-1, keepdim=True) ).bfloat16()
    return hidden, weight, ids


@pytest.mark.parametrize("T,H,V,tau", [(48, 32, 160, 1.0), (300, 256, 5000, 0.6), (1000, 512, 3333 * 8, 1.0), (129, 1024, 151936, 0.6)])
def test_lmhead_forward_vs_oracle(P, T, H, V, tau):
    hidden, weight, ids = _mk(T, H, V, T + V, peaked=(T == 300))
    lp, ent = P.fused_lm_head_logprobs_entropy(hidden.cuda().unsqueeze(0), weight.cuda(), ids.cuda().unsqueeze(0), tau)
    logits = oracle.lm_head_logits(hidden.float().numpy(), weight.float().numpy(), True)
    wl, we, _ = oracle.logprobs_entropy(oracle.shift_rows(logits), ids.numpy(), tau)
    dl = np.abs(lp[0].cpu().numpy() - wl)
    de = np.abs(ent[0].cpu().numpy() - we)
    # tcgen05 and the fp64 oracle accumulate in different orders: a logit that sits on a bf16 rounding
    # midpoint may round the other way (SURVEY §7 "bf16 rounding"), moving logp by <= 1 bf16 ulp of the logit.
    assert int((dl >= 1e-3).sum()) <= allowed_flip_outliers(T, H) and dl.max() < 0.07, (dl.max(), int((dl >= 1e-3).sum()))
    assert de.max() < 1e-3, de.max()
    # ... and that is all there is: the kernel's logits are the correctly rounded product, and from THOSE logits every token matches
    klog, blk = kernel_logits(hidden, weight, ids, tau)
    rn = rn_bf16_violations(blk, hidden.cuda(), weight.cuda(), torch.arange(T - 1, device="cuda"))
    assert rn["violations"] == 0, rn
    kl, ke, _ = oracle.logprobs_entropy(oracle.shift_rows(klog), ids.numpy(), tau)
    assert np.abs(lp[0].cpu().numpy() - kl).max() < 5e-4 and np.abs(ent[0].cpu().numpy() - ke).max() < 5e-4
    # closed-form first row (shift_logits' zero row): logp = -ln V, entropy = ln V
    assert abs(lp[0, 0].item() + np.log(V)) < 1e-5 and abs(ent[0, 0].item() - np.log(V)) < 1e-5


def test_lmhead_forward_exact_rounding_free(P):
    # without bf16 rounding of the accumulator there is no boundary effect: every token within 1e-3 (in fact ~1e-5)
    T, H, V, tau = 257, 128, 4096 + 64, 0.7
    hidden, weight, ids = _mk(T, H, V, 9)
    from prime_rl_cse_b200 import lm_head_forward_raw

    lp, ent, lse, _ = lm_head_forward_raw(hidden.cuda(), weight.cuda(), ids.cuda(), 1.0 / tau, shift=True, round_bf16=False)
    logits = oracle.lm_head_logits(hidden.float().numpy(), weight.float().numpy(), False)
    wl, we, wlse = oracle.logprobs_entropy(oracle.shift_rows(logits), ids.numpy(), tau)
    np.testing.assert_allclose(lp.cpu().numpy(), wl, atol=1e-4)
    np.testing.assert_allclose(ent.cpu().numpy(), we, atol=1e-4)
    np.testing.assert_allclose(lse.cpu().numpy(), wlse, atol=1e-4)


def test_lmhead_fixtures_from_reference(P):
    d = load("lmhead.npz")
    for key in d["names"]:
        name, tau_s, cname = key.split("/")
        if not bool(d[f"{name}/is_bf16"]):
            continue
        tau = float(tau_s[3:])
        hidden, weight = cu(d[f"{name}/hidden"], torch.bfloat16), cu(d[f"{name}/weight"], torch.bfloat16)
        hidden.requires_grad_(True)
        weight.requires_grad_(True)
        mb = dict(input_ids=cu(d[f"{name}/ids"]).unsqueeze(0), position_ids=cu(d[f"{name}/pos"]).unsqueeze(0),
                  advantages=cu(d[f"{name}/adv"]).unsqueeze(0), logprobs=cu(d[f"{key}/old"]).unsqueeze(0),
                  loss_mask=cu(d[f"{name}/mask"]).unsqueeze(0), temperature=tau)
        cfg = P.LossConfig() if cname == "token" else P.LossConfig("sequence", True, 8.0)
        out = P.rl_micro_step(hidden.unsqueeze(0), weight, mb, cfg, int(d[f"{key}/scale"]))
        # every token, from the kernel's own logits
        klog, _ = kernel_logits(hidden.detach(), weight.detach(), mb["input_ids"][0], tau)
        same = oracle.micro_step(d[f"{name}/hidden"], d[f"{name}/weight"], d[f"{name}/ids"], d[f"{name}/pos"], d[f"{key}/old"],
                                 d[f"{name}/adv"], d[f"{name}/mask"], tau, LossCfg(*((cname,) if cname == "token" else ("sequence", True, 8.0))),
                                 int(d[f"{key}/scale"]), logits=klog)
        assert_same_logits_micro_step(out, same, hidden.grad, weight.grad)
        # the reference's own numbers (its GEMM rounds midpoint logits its own way)
        dl = np.abs(out["logprobs"][0].cpu().numpy() - d[f"{key}/logp"])
        n_out = int((dl >= 1e-3).sum())
        assert n_out <= allowed_flip_outliers(dl.size, hidden.shape[-1]) and dl.max() < 0.07, (key, dl.max(), n_out)
        np.testing.assert_allclose(out["loss"].item(), d[f"{key}/loss"], rtol=1e-4 if n_out == 0 else 5e-3, atol=1e-6)
        assert rel_l2(hidden.grad.float().cpu().numpy(), d[f"{key}/d_hidden"]) < 1e-2, key
        assert rel_l2(weight.grad.float().cpu().numpy(), d[f"{key}/d_weight"]) < 1e-2, key


@pytest.mark.parametrize("save", [False, True], ids=["recompute", "saved-logits"])
@pytest.mark.parametrize("T,H,V,tau,mode,vchunk", [
    (48, 32, 160, 1.0, "token", 0),
    (300, 256, 5000, 0.6, "token", 1024),
    (640, 512, 20000, 1.0, "sequence", 8192),
    (200, 1024, 151936, 0.6, "token", 0),
])
def test_micro_step_vs_oracle(P, T, H, V, tau, mode, vchunk, save):
    from prime_rl_cse_b200 import _lib as _lib_mod

    _lib_mod.set_knob("PRL_VCHUNK", vchunk if vchunk else 8192)  # knobs are read from the environment once; tests set them directly
    try:
        hidden, weight, ids = _mk(T, H, V, T + H, peaked=True)
        rng = np.random.default_rng(T)
        lens = [T // 3, T // 4, T - T // 3 - T // 4]
        pos = np.concatenate([np.arange(n) for n in lens])
        adv = np.concatenate([np.full(n, rng.standard_normal(), dtype=np.float32) for n in lens])
        mask = rng.random(T) < 0.75
        cfg_o = LossCfg(mode, mode == "sequence", 8.0)
        scale = int(mask.sum()) if mode == "token" else 1
        logits = oracle.lm_head_logits(hidden.float().numpy(), weight.float().numpy(), True)
        lp_true, _, _ = oracle.logprobs_entropy(oracle.shift_rows(logits), ids.numpy(), tau)
        old = (lp_true + 0.3 * rng.standard_normal(T) / (20 if mode == "sequence" else 1)).astype(np.float32)
        want = oracle.micro_step(hidden.float().numpy(), weight.float().numpy(), ids.numpy(), pos, old, adv, mask, tau, cfg_o, scale)

        h = hidden.cuda().requires_grad_(True)
        w = weight.cuda().requires_grad_(True)
        mb = dict(input_ids=ids.cuda().unsqueeze(0), position_ids=cu(pos).unsqueeze(0), advantages=cu(adv).unsqueeze(0),
                  logprobs=cu(old).unsqueeze(0), loss_mask=cu(mask).unsqueeze(0), temperature=tau)
        out = P.rl_micro_step(h.unsqueeze(0), w, mb, P.LossConfig(mode, mode == "sequence", 8.0), scale, save_logits=save)
        klog, _ = kernel_logits(hidden, weight, ids, tau)
        same = oracle.micro_step(hidden.float().numpy(), weight.float().numpy(), ids.numpy(), pos, old, adv, mask, tau, cfg_o, scale, logits=klog)
        assert_same_logits_micro_step(out, same, h.grad, w.grad)  # every token at the stated tolerances
        dl = np.abs(out["logprobs"][0].cpu().numpy() - want["logp"])
        n_out = int((dl >= 1e-3).sum())
        assert n_out <= allowed_flip_outliers(T, H) and dl.max() < 0.07, (dl.max(), n_out)
        np.testing.assert_allclose(out["loss"].item(), want["loss"], rtol=1e-4 if n_out == 0 else 5e-3, atol=1e-6)
        assert rel_l2(h.grad.float().cpu().numpy(), want["d_hidden"]) < 1e-2
        assert rel_l2(w.grad.float().cpu().numpy(), want["d_weight"]) < 1e-2
        assert h.grad.dtype == torch.bfloat16 and w.grad.dtype == torch.bfloat16
        assert float(h.grad[-1].abs().max()) == 0.0  # the dropped last position gets no gradient (loss.py:54)
    finally:
        _lib_mod.set_knob("PRL_VCHUNK", 16384)


@pytest.mark.parametrize("cta_group,epi_tma", [(1, 1), (2, 0), (4, 1), (1, 0)])
def test_engine_variants(P, cta_group, epi_tma):
    """The non-default engine variants (single CTA, two pairs with B multicast, register epilogue) stay correct."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PRL_CTA_GROUP=str(cta_group), PRL_EPI_TMA=str(epi_tma))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "variant_check.py")], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]


def test_fused_lm_head_module_in_hf_qwen3(P):
    """swap `model.lm_head` of a (tiny, random-init) HF Qwen3 for `FusedLMHead`; the model's
    `.logits` then carries (logprob, entropy) and matches the reference chain on the unmodified model."""
    transformers = pytest.importorskip("transformers")
    cfg = transformers.Qwen3Config(vocab_size=1024, hidden_size=128, intermediate_size=256, num_hidden_layers=2,
                                   num_attention_heads=4, num_key_value_heads=2, head_dim=32, max_position_embeddings=512,
                                   tie_word_embeddings=False)
    torch.manual_seed(0)
    model = transformers.Qwen3ForCausalLM(cfg).to(device="cuda", dtype=torch.bfloat16)
    T, tau = 200, 0.8
    ids = torch.randint(0, cfg.vocab_size, (1, T), device="cuda")
    pos = torch.cat([torch.arange(120), torch.arange(80)]).unsqueeze(0).cuda()

    # reference chain on the unmodified model (train.py:285-288, 302-303) in plain torch
    logits = model(input_ids=ids, position_ids=pos).logits.float().contiguous()
    shifted = torch.cat([torch.zeros_like(logits[:, :1]), logits[:, :-1]], dim=1) / tau
    ref_lp = torch.gather(shifted, -1, ids.unsqueeze(-1)).squeeze(-1) - torch.logsumexp(shifted, -1)
    ref_ent = torch.logsumexp(shifted, -1) - (torch.softmax(shifted, -1) * shifted).sum(-1)
    (ref_lp * 0.01).sum().backward()
    ref_gw = model.lm_head.weight.grad.float().clone()
    ref_ge = model.model.embed_tokens.weight.grad.float().clone()
    model.zero_grad()

    head = P.FusedLMHead.from_linear(model.lm_head)
    model.lm_head = head
    head.set_context(ids, tau)
    out = model(input_ids=ids, position_ids=pos).logits
    assert out.shape == (1, T, 2)
    lp, ent = out[..., 0], out[..., 1].detach()
    d = (lp - ref_lp).abs()
    assert int((d >= 1e-3).sum()) <= allowed_flip_outliers(T, cfg.hidden_size) and d.max().item() < 0.07
    assert int(((ent - ref_ent).abs() >= 1e-3).sum()) <= allowed_flip_outliers(T, cfg.hidden_size)
    (lp * 0.01).sum().backward()
    gw = model.lm_head.weight.grad.float()
    ge = model.model.embed_tokens.weight.grad.float()
    assert ((gw - ref_gw).norm() / ref_gw.norm()).item() < 1e-2
    assert ((ge - ref_ge).norm() / ref_ge.norm()).item() < 2e-2  # through the whole (bf16) transformer body
    with pytest.raises(RuntimeError):
        model(input_ids=ids, position_ids=pos)  # context must be set before every forward
    # with the micro-batch's loss_mask the backward skips masked positions: same gradients for a loss that ignores them
    mask = torch.rand(1, T, device="cuda") < 0.6
    grads = []
    for m in (None, mask):
        model.zero_grad()
        head.set_context(ids, tau, m)
        lp_m = model(input_ids=ids, position_ids=pos).logits[..., 0]
        (lp_m * mask * 0.01).sum().backward()
        grads.append((model.lm_head.weight.grad.float().clone(), model.model.embed_tokens.weight.grad.float().clone()))
    assert ((grads[1][0] - grads[0][0]).norm() / grads[0][0].norm()).item() < 5e-3
    assert ((grads[1][1] - grads[0][1]).norm() / grads[0][1].norm()).item() < 2e-2


def test_next_rows_sft_cross_entropy_and_logprob_recompute(P):
    """SURVEY §8f N3 (SFT cross-entropy on the fused LM head) and N2 (forward-only old-logprob recompute)."""
    d = load("sft.npz")
    hidden, weight = cu(d["sft_bf16/hidden"], torch.bfloat16), cu(d["sft_bf16/weight"], torch.bfloat16)
    targets, mask = cu(d["sft_bf16/targets"]), cu(d["sft_bf16/mask"])
    hidden.requires_grad_(True)
    weight.requires_grad_(True)
    loss = P.fused_linear_cross_entropy(hidden, weight, targets, mask)
    loss.backward()
    want, wdh, wdw = oracle.sft_cross_entropy(d["sft_bf16/hidden"][0], d["sft_bf16/weight"], d["sft_bf16/targets"], d["sft_bf16/mask"], True)
    np.testing.assert_allclose(loss.item(), want, rtol=2e-3)
    np.testing.assert_allclose(loss.item(), d["sft_bf16/loss"], rtol=2e-2)  # reference call site computes in bf16
    assert rel_l2(hidden.grad.float().cpu().numpy()[0], wdh) < 1e-2 and rel_l2(weight.grad.float().cpu().numpy(), wdw) < 1e-2
    assert rel_l2(hidden.grad.float().cpu().numpy(), d["sft_bf16/d_hidden"]) < 3e-2

    T, H, V, tau = 333, 256, 5000, 0.6
    h, w, ids = _mk(T, H, V, 5)
    logits = oracle.lm_head_logits(h.float().numpy(), w.float().numpy(), True)
    wl, _, _ = oracle.logprobs_entropy(oracle.shift_rows(logits), ids.numpy(), tau)
    old = torch.from_numpy(wl).float().cuda() + 0.1
    lp, err = P.recompute_logprobs(h.cuda().unsqueeze(0), w.cuda(), ids.cuda().unsqueeze(0), tau, old.unsqueeze(0))
    dl = np.abs(lp[0].cpu().numpy() - wl)
    assert int((dl >= 1e-3).sum()) <= allowed_flip_outliers(T, H) and dl.max() < 0.07
    klog, _ = kernel_logits(h, w, ids, tau)
    kl, _, _ = oracle.logprobs_entropy(oracle.shift_rows(klog), ids.numpy(), tau)
    assert np.abs(lp[0].cpu().numpy() - kl).max() < 5e-4  # every token, from the kernel's own logits
    np.testing.assert_allclose(err[0].cpu().numpy(), np.exp(lp[0].cpu().numpy() - old.cpu().numpy()), rtol=1e-5)


def test_next_row_on_device_logging_stats(P):
    """masked mean/std/min/max per logged key, accumulated across micro-batches on the device."""
    rng = np.random.default_rng(8)
    st = P.TensorStats()
    vals, probs = [], []
    for T in (1000, 16384, 77):
        x = rng.standard_normal(T).astype(np.float32) - 1.0
        m = rng.random(T) < 0.6
        st.add("entropy", cu(x), cu(m))
        st.add("probs", cu(x), cu(m), exp=True)
        st.add("all", cu(x))
        vals.append(x[m])
        probs.append(np.exp(x[m].astype(np.float64)))
    got = st.compute_stats()
    v, p = np.concatenate(vals).astype(np.float64), np.concatenate(probs)
    for key, ref in (("entropy", v), ("probs", p)):
        np.testing.assert_allclose(got[f"{key}/mean"], ref.mean(), rtol=1e-6)
        np.testing.assert_allclose(got[f"{key}/std"], ref.std(ddof=1), rtol=1e-5)
        np.testing.assert_allclose(got[f"{key}/min"], ref.min(), rtol=1e-6)
        np.testing.assert_allclose(got[f"{key}/max"], ref.max(), rtol=1e-6)
        assert got[f"{key}/count"] == len(ref)
    assert got["all/count"] == 1000 + 16384 + 77 and st.compute_stats() == {}
    # the median is torch.median's lower median of the same fp32 values, bit for bit
    assert got["entropy/median"] == torch.median(torch.from_numpy(np.concatenate(vals))).item()
    np.testing.assert_allclose(got["probs/median"], torch.median(torch.exp(cu(np.concatenate(vals)))).item(), rtol=2e-7)  # expf vs ATen exp


@pytest.mark.parametrize("mode", ["token", "sequence"])
def test_next_row_micro_step_logging_keys(P, mode):
    """`TensorStats.add_micro_step` logs what rl/train.py:313-333 logs, under the same keys; values checked against the
    boolean-index + torch statistics the reference computes (trainer/utils.py:246-250)."""
    T, H, V, tau = 300, 64, 1000, 0.9
    hidden, weight, ids = _mk(T, H, V, 3)
    rng = np.random.default_rng(3)
    lens = [100, 120, 80]
    mask = cu(rng.random(T) < 0.7).unsqueeze(0)
    mb = dict(input_ids=ids.cuda().unsqueeze(0), position_ids=cu(np.concatenate([np.arange(n) for n in lens])).unsqueeze(0),
              advantages=cu(rng.standard_normal(T).astype(np.float32)).unsqueeze(0),
              logprobs=cu(-rng.random(T).astype(np.float32) * 4).unsqueeze(0), loss_mask=mask, temperature=tau)
    st = P.TensorStats()
    ref = {}
    for _ in range(2):  # two micro-batches of one step
        out = P.rl_micro_step(hidden.cuda().unsqueeze(0).requires_grad_(True), weight.cuda().requires_grad_(True), mb,
                              P.LossConfig(mode, mode == "sequence", 8.0), 5)
        err = torch.exp(out["logprobs"] - mb["logprobs"])
        st.add_micro_step(out, mb, ratio_type=mode, recomputed_logprob_error=err)
        per = {"probs": torch.exp(out["logprobs"])[mask], "old_probs": torch.exp(mb["logprobs"])[mask], "entropy": out["entropy"][mask],
               "recomputed_logprob_error": err[mask], "loss": out["loss"].reshape(1)}
        for k in ("importance_ratio", "clipped_importance_ratio", "is_clipped"):
            per[k] = out[k].float() if mode == "sequence" else out[k].float()[mask.squeeze()]
        for k, v in per.items():
            ref.setdefault(k, []).append(v.float())
    got = st.compute_stats()
    for k, parts in ref.items():
        v = torch.cat(parts)
        assert got[f"{k}/count"] == v.numel()
        np.testing.assert_allclose(got[f"{k}/mean"], v.double().mean().item(), rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(got[f"{k}/median"], torch.median(v).item(), rtol=2e-7, atol=0)
        np.testing.assert_allclose(got[f"{k}/min"], v.min().item(), rtol=2e-7)
        np.testing.assert_allclose(got[f"{k}/max"], v.max().item(), rtol=2e-7)
        if v.numel() > 1:
            np.testing.assert_allclose(got[f"{k}/std"], v.double().std().item(), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("n,kind", [(1, "normal"), (2, "normal"), (7, "ties"), (100000, "ties"), (300001, "normal"), (4096, "signed-zero-inf"),
                                    (70000, "many-adds")])
def test_next_row_exact_median_by_radix_select(P, n, kind):
    """`{key}/median` == torch.median (lower median) over every masked value of every add: odd / even counts, heavy ties, negative
    values, +-0, +-inf, a buffer that has to grow, an all-masked add, and nan propagation."""
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, generator=g) * 3
    if kind == "ties":
        x = torch.round(x)  # a handful of distinct values, many duplicates
    if kind == "signed-zero-inf":
        x[::7] = 0.0
        x[1::7] = -0.0
        x[2::97] = float("inf")
        x[3::89] = float("-inf")
    mask = torch.rand(n, generator=g) < 0.7
    mask[0] = True
    st = P.TensorStats()
    pieces = 9 if kind == "many-adds" else 2
    for part, pm in zip(x.chunk(pieces), mask.chunk(pieces)):
        st.add("v", part.cuda(), pm.cuda())
    st.add("v", torch.ones(50).cuda(), torch.zeros(50, dtype=torch.bool).cuda())  # an all-masked micro-batch adds nothing
    st.add("w", x.cuda())                                                          # no mask
    got = st.compute_stats()
    assert got["v/median"] == torch.median(x[mask]).item() and got["v/count"] == int(mask.sum())
    assert got["w/median"] == torch.median(x).item()
    if kind == "normal" and n > 2:
        st.add("v", x.cuda(), mask.cuda())
        st.add("v", torch.tensor([float("nan")]).cuda())
        assert math.isnan(st.compute_stats()["v/median"])  # torch.median propagates nan
    assert P.TensorStats(median=False).compute_stats() == {}


@pytest.mark.parametrize("save", [False, True], ids=["recompute", "saved-logits"])
@pytest.mark.parametrize("T,H,V", [(1, 64, 512), (2, 8, 8), (3, 64, 255), (127, 72, 1000), (129, 200, 257), (513, 136, 4104)])
def test_micro_step_edge_shapes(P, T, H, V, save):
    """Ragged / tiny shapes: T = 1 (only shift_logits' zero row), V and H that are not multiples of the tile sizes
    (H % 8 == 0 is the only requirement: 16-byte TMA row pitch), single-token segments."""
    hidden, weight, ids = _mk(T, H, V, 1000 + T)
    rng = np.random.default_rng(T)
    lens = [T] if T < 3 else [1, T - 2, 1]
    pos = np.concatenate([np.arange(n) for n in lens])
    adv = rng.standard_normal(T).astype(np.float32)
    mask = np.ones(T, dtype=bool)
    tau = 0.9
    logits = oracle.lm_head_logits(hidden.float().numpy(), weight.float().numpy(), True)
    lp_true, _, _ = oracle.logprobs_entropy(oracle.shift_rows(logits), ids.numpy(), tau)
    old = (lp_true + 0.2 * rng.standard_normal(T)).astype(np.float32)
    want = oracle.micro_step(hidden.float().numpy(), weight.float().numpy(), ids.numpy(), pos, old, adv, mask, tau, LossCfg(), T)
    h = hidden.cuda().requires_grad_(True)
    w = weight.cuda().requires_grad_(True)
    mb = dict(input_ids=ids.cuda().unsqueeze(0), position_ids=cu(pos).unsqueeze(0), advantages=cu(adv).unsqueeze(0),
              logprobs=cu(old).unsqueeze(0), loss_mask=cu(mask).unsqueeze(0), temperature=tau)
    out = P.rl_micro_step(h.unsqueeze(0), w, mb, P.LossConfig(), T, response_lengths=want["lengths"], save_logits=save)
    klog, _ = kernel_logits(hidden, weight, ids, tau)
    same = oracle.micro_step(hidden.float().numpy(), weight.float().numpy(), ids.numpy(), pos, old, adv, mask, tau, LossCfg(), T, logits=klog)
    assert_same_logits_micro_step(out, same, h.grad if T > 1 else None, w.grad if T > 1 else None)
    dl = np.abs(out["logprobs"][0].cpu().numpy() - want["logp"])
    assert dl.max() < 0.07 and int((dl >= 1e-3).sum()) <= allowed_flip_outliers(T, H)
    if T > 1:
        assert rel_l2(h.grad.float().cpu().numpy(), want["d_hidden"]) < 1.5e-2
        assert rel_l2(w.grad.float().cpu().numpy(), want["d_weight"]) < 1.5e-2
    else:
        assert float(h.grad.abs().max()) == 0.0 and float(w.grad.abs().max()) == 0.0  # row 0 is a constant


def test_unsupported_shapes_raise(P):
    h = torch.zeros(1, 4, 12, dtype=torch.bfloat16, device="cuda")  # H % 8 != 0
    w = torch.zeros(16, 12, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError):
        P.fused_lm_head_logprobs_entropy(h, w, torch.zeros(1, 4, dtype=torch.long, device="cuda"), 1.0)
    with pytest.raises(ValueError):
        P.fused_lm_head_logprobs_entropy(h.float(), w.float(), torch.zeros(1, 4, dtype=torch.long, device="cuda"), 1.0)
    with pytest.raises(ValueError):
        P.compute_loss([torch.zeros(4, device="cuda")], [torch.zeros(3, device="cuda")], [torch.zeros(4, device="cuda")],
                       [torch.ones(4, dtype=torch.bool, device="cuda")], P.LossConfig(), 1)


def test_backward_output_modes(P):
    """d_weight in bf16 / fp32, stored or accumulated in the kernel; d_hidden in bf16 / fp32; W-only and hidden-only."""
    from prime_rl_cse_b200 import lm_head_backward_raw, lm_head_forward_raw

    T, H, V, tau = 384, 256, 9000, 0.8
    hidden, weight, ids = _mk(T, H, V, 21)
    h, w, i = hidden.cuda(), weight.cuda(), ids.cuda()
    g = torch.randn(T, device="cuda") / T
    logp, ent, lse, _ = lm_head_forward_raw(h, w, i, 1 / tau)
    dh_ref, dw_ref = lm_head_backward_raw(h, w, i, lse, g, 1 / tau, dweight=torch.zeros(V, H, device="cuda"), dweight_accumulate=False,
                                          dhidden_f32=True, vchunk=4096)
    assert dh_ref.dtype == torch.float32 and dw_ref.dtype == torch.float32

    def rel(a, b):
        return ((a.float() - b.float()).norm() / b.float().norm()).item()

    for vchunk in (0, 4096):
        dh, dw = lm_head_backward_raw(h, w, i, lse, g, 1 / tau, vchunk=vchunk)  # bf16 / bf16, stored
        assert dh.dtype == torch.bfloat16 and dw.dtype == torch.bfloat16
        assert rel(dh, dh_ref) < 5e-3 and rel(dw, dw_ref) < 5e-3
    acc32 = dw_ref.clone()
    lm_head_backward_raw(h, w, i, lse, g, 1 / tau, dweight=acc32, dweight_accumulate=True, want_dhidden=False)  # fp32 +=
    assert rel(acc32, 2 * dw_ref) < 1e-5
    acc16 = dw_ref.bfloat16()
    lm_head_backward_raw(h, w, i, lse, g, 1 / tau, dweight=acc16, dweight_accumulate=True, want_dhidden=False)  # bf16 +=
    assert rel(acc16, 2 * dw_ref) < 5e-3
    dh_only, none_w = lm_head_backward_raw(h, w, i, lse, g, 1 / tau, want_dweight=False)
    assert none_w is None and rel(dh_only, dh_ref) < 5e-3


@pytest.mark.parametrize("T,H,V,shift", [(129, 64, 1003, True), (700, 256, 20000 + 64, True), (384, 128, 4096, False), (1, 64, 300, True)])
def test_saved_logits_mode_is_the_recompute_mode_bit_for_bit(P, T, H, V, shift):
    """Saved-logits pair (prl_lmhead_fwd_save / prl_lmhead_bwd_saved): the forward statistics are bit-identical to the plain
    forward, the kept block holds the bf16 logits, the in-place logits -> dlogits pass reproduces the recompute epilogue's
    dlogits bit for bit, and (single vocab chunk: same K order) so do d_hidden / d_weight."""
    from prime_rl_cse_b200 import (lm_head_backward_raw, lm_head_backward_saved_raw, lm_head_forward_raw, lm_head_forward_save_raw)
    from prime_rl_cse_b200 import _lib

    tau = 0.8
    hidden, weight, ids = _mk(T, H, V, 7 * T + V, peaked=True)
    h, w, i = hidden.cuda(), weight.cuda(), ids.cuda()
    g = torch.randn(T, device="cuda") / T
    lp0, ent0, lse0, _ = lm_head_forward_raw(h, w, i, 1 / tau, shift=shift)
    lp1, ent1, lse1, _, saved = lm_head_forward_save_raw(h, w, i, 1 / tau, shift=shift)
    assert torch.equal(lp0, lp1) and torch.equal(ent0, ent1) and torch.equal(lse0, lse1)
    rows = T - int(shift)
    assert saved.shape[0] >= T and saved.shape[1] >= V and saved.shape[1] % 8 == 0
    want_logits = (h[:rows].float() @ w.float().t())
    got = saved[:rows, :V].float()
    ulp = 2.0 ** -7 * want_logits.abs() + 1e-4  # one bf16 ulp is at most 2^-7 relative; + fp32 accumulation-order noise
    assert bool(((got - want_logits).abs() <= ulp).all())  # = a correctly rounded bf16 of some fp32 accumulation order
    # recompute mode, dlogits only (phase 1), single vocab chunk: its workspace starts with the [T_pad, vc] bf16 dlogits block
    lib = _lib.load()
    vc = (V + 255) // 256 * 256
    ws = torch.zeros(lib.prl_lmhead_bwd_workspace_bytes(T, H, V, vc), dtype=torch.uint8, device="cuda")
    lm_head_backward_raw(h, w, i, lse0, g, 1 / tau, shift=shift, vchunk=vc, phase_mask=1, workspace=ws)
    dz_re = ws[: ((T + 127) // 128 * 128) * vc * 2].view(torch.bfloat16).view(-1, vc)[:T, :V]
    blk = saved.clone()
    lm_head_backward_saved_raw(h, w, i, lse0, g, 1 / tau, blk, shift=shift, phase_mask=1, want_dhidden=False, want_dweight=False)
    assert torch.equal(blk[:T, :V].float(), dz_re.float())  # float compare: -0.0 == 0.0 for the target-less last row
    if shift:
        assert float(blk[T - 1, :V].abs().max()) == 0.0
    # full backward, both modes
    dh0, dw0 = lm_head_backward_raw(h, w, i, lse0, g, 1 / tau, shift=shift, vchunk=vc)
    dh1, dw1 = lm_head_backward_saved_raw(h, w, i, lse0, g, 1 / tau, saved, shift=shift)
    assert torch.equal(dh0, dh1) and torch.equal(dw0, dw1)
    # many vocab chunks in the recompute mode: the fp32 partial sums are combined in another order, nothing else changes
    if V > 512 and T > 1:
        dh2, dw2 = lm_head_backward_raw(h, w, i, lse0, g, 1 / tau, shift=shift, vchunk=256)
        assert rel_l2(dh2.float().cpu().numpy(), dh1.float().cpu().numpy()) < 3e-3
        assert rel_l2(dw2.float().cpu().numpy(), dw1.float().cpu().numpy()) < 1e-3
    # fp32 accumulation into main_grad and fp32 d_hidden
    acc = torch.ones(V, H, device="cuda")
    _, _, _, _, saved = lm_head_forward_save_raw(h, w, i, 1 / tau, shift=shift)
    dh3, _ = lm_head_backward_saved_raw(h, w, i, lse0, g, 1 / tau, saved, shift=shift, dweight=acc, dweight_accumulate=True, dhidden_f32=True)
    assert dh3.dtype == torch.float32
    if T > 1:
        assert rel_l2((acc - 1).cpu().numpy(), dw1.float().cpu().numpy()) < 5e-3
        assert rel_l2(dh3.cpu().numpy(), dh1.float().cpu().numpy()) < 5e-3


def test_dhidden_k_split_matches_unsplit_and_is_deterministic(P):
    """Few output tiles (small T x H) and a long K: the d_hidden GEMM runs as two k-splits that reduce-add into a zeroed fp32
    accumulator.  Same result as the unsplit GEMM up to fp32 summation order, and run-to-run identical (two addends commute)."""
    from prime_rl_cse_b200 import lm_head_backward_raw, lm_head_backward_saved_raw, lm_head_forward_save_raw

    from prime_rl_cse_b200 import _lib as _lib_mod

    T, H, V, tau = 600, 256, 20000 + 64, 0.8
    hidden, weight, ids = _mk(T, H, V, 17, peaked=True)
    h, w, i = hidden.cuda(), weight.cuda(), ids.cuda()
    g = torch.randn(T, device="cuda") / T
    out = {}
    try:
        for ks in ("1", "2"):
            _lib_mod.set_knob("PRL_K_SPLITS", int(ks))
            for f32 in (False, True):
                lse = lm_head_forward_save_raw(h, w, i, 1 / tau)[2]
                a = lm_head_backward_raw(h, w, i, lse, g, 1 / tau, vchunk=V, dhidden_f32=f32)[0]
                b = lm_head_backward_raw(h, w, i, lse, g, 1 / tau, vchunk=4096, dhidden_f32=f32)[0]
                blk = lm_head_forward_save_raw(h, w, i, 1 / tau)[4]
                c = lm_head_backward_saved_raw(h, w, i, lse, g, 1 / tau, blk, dhidden_f32=f32)[0]
                blk = lm_head_forward_save_raw(h, w, i, 1 / tau)[4]
                c2 = lm_head_backward_saved_raw(h, w, i, lse, g, 1 / tau, blk, dhidden_f32=f32)[0]
                assert torch.equal(c, c2) and torch.equal(a, c)  # deterministic; single-chunk recompute == saved
                out[(ks, f32)] = (a.float(), b.float(), c.float())
    finally:
        _lib_mod.set_knob("PRL_K_SPLITS", 0)
    for f32 in (False, True):
        ref = out[("1", f32)][0]
        for t in out[("2", f32)] + out[("1", f32)][1:]:
            assert ((t - ref).norm() / ref.norm()).item() < 2e-3


def test_saved_logits_policy_and_second_backward(P):
    """Autograd integration: the default is the recompute backward (nothing of size [T, V] is stored); PRL_SAVE_LOGITS=auto
    decides by size; a second backward through a retained graph (the kept block is consumed by the first one) falls back to
    the recompute, says so once, and gives the same gradients."""
    import warnings

    from prime_rl_cse_b200 import saved_logits_policy

    assert os.environ.get("PRL_SAVE_LOGITS") is None
    assert not saved_logits_policy(16384, 151936) and not saved_logits_policy(16, 1000)  # default: recompute
    os.environ["PRL_SAVE_LOGITS"] = "auto"
    try:
        assert saved_logits_policy(16384, 151936) and not saved_logits_policy(10**6, 151936) and not saved_logits_policy(0, 1000)
    finally:
        os.environ.pop("PRL_SAVE_LOGITS")
    assert saved_logits_policy(10**6, 151936, True) and not saved_logits_policy(16, 1000, False)
    T, H, V = 300, 128, 3000
    hidden, weight, ids = _mk(T, H, V, 5)
    h = hidden.cuda().requires_grad_(True)
    w = weight.cuda().requires_grad_(True)
    lp, _ = P.fused_lm_head_logprobs_entropy(h.unsqueeze(0), w, ids.cuda().unsqueeze(0), 0.9, save_logits=True)
    loss = (lp[0] * torch.linspace(-1, 1, T, device="cuda")).sum()
    loss.backward(retain_graph=True)
    g1 = (h.grad.clone(), w.grad.clone())
    h.grad = w.grad = None
    import prime_rl_cse_b200.lm_head as lm_head_mod

    lm_head_mod._warned_fallback = False
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        loss.backward()
    assert any("recompute backward" in str(c.message) for c in caught)  # the 25 % cliff is announced (once per process)
    assert torch.equal(g1[0], h.grad) and torch.equal(g1[1], w.grad)
    with torch.no_grad():  # nothing is kept when no gradient can flow
        lp2, _ = P.fused_lm_head_logprobs_entropy(hidden.cuda().unsqueeze(0), weight.cuda(), ids.cuda().unsqueeze(0), 0.9, save_logits=True)
        assert torch.equal(lp2, lp.detach())


def test_micro_step_on_reference_packed_batches(P):
    """The packed micro-batches the reference orchestrator writes (prepare_batch, 2 train workers, incl. the all-masked padding
    batch) through the fused micro-step, each rank with its rank-local loss scale (train.py:263-267)."""
    d = load("packed_layout.npz")
    H, V = 64, 1000
    hidden_rng = torch.Generator().manual_seed(3)
    weight = (torch.randn(V, H, generator=hidden_rng) * 0.4).bfloat16()
    for r in range(int(d["n_ranks"])):
        mbs = []
        for i in range(int(d[f"rank{r}/n"])):
            pre = f"rank{r}/mb{i}"
            mbs.append({k: torch.from_numpy(d[f"{pre}/{k}"]) for k in ("input_ids", "position_ids", "advantages", "logprobs", "loss_mask")}
                       | {"temperature": float(d[f"{pre}/temperature"])})
        scale = P.local_loss_scale(mbs, "token")
        for mb in mbs:
            T = mb["input_ids"].shape[1]
            hidden = torch.randn(T, H, generator=hidden_rng).bfloat16()
            want = oracle.micro_step(hidden.float().numpy(), weight.float().numpy(), mb["input_ids"][0].numpy(), mb["position_ids"][0].numpy(),
                                     mb["logprobs"][0].numpy(), mb["advantages"][0].numpy(), mb["loss_mask"][0].numpy(), mb["temperature"],
                                     LossCfg(), scale)
            h = hidden.cuda().requires_grad_(True)
            w = weight.cuda().requires_grad_(True)
            out = P.rl_micro_step(h.unsqueeze(0), w, P.to_device(mb, "cuda"), P.LossConfig(), scale)
            assert P.get_response_lengths(mb["position_ids"].cuda()) == want["lengths"]
            klog, _ = kernel_logits(hidden, weight, mb["input_ids"][0], mb["temperature"])
            same = oracle.micro_step(hidden.float().numpy(), weight.float().numpy(), mb["input_ids"][0].numpy(), mb["position_ids"][0].numpy(),
                                     mb["logprobs"][0].numpy(), mb["advantages"][0].numpy(), mb["loss_mask"][0].numpy(), mb["temperature"],
                                     LossCfg(), scale, logits=klog)
            assert_same_logits_micro_step(out, same, h.grad, w.grad)
            dl = np.abs(out["logprobs"][0].cpu().numpy() - want["logp"])
            assert int((dl >= 1e-3).sum()) <= allowed_flip_outliers(T, H) and dl.max() < 0.07
            if mb["loss_mask"].any():
                assert rel_l2(h.grad.float().cpu().numpy(), want["d_hidden"]) < 1.5e-2
                assert rel_l2(w.grad.float().cpu().numpy(), want["d_weight"]) < 1.5e-2
            else:  # padding batch: exactly zero loss and gradients
                assert out["loss"].item() == 0.0 and float(h.grad.abs().max()) == 0.0 and float(w.grad.abs().max()) == 0.0


@pytest.mark.parametrize("save", [False, True], ids=["recompute", "saved-logits"])
def test_backward_skips_masked_positions_exactly(P, save):
    """`loss_mask`: the backward over the active positions only gives the same gradients as the full one (masked rows
    contribute zeros); no index tensors are built on the host."""
    T, H, V, tau = 900, 256, 6000, 0.7
    hidden, weight, ids = _mk(T, H, V, 31)
    mask = torch.rand(T) < 0.6
    mask[0] = True  # position 0 is active but has no source row: must be ignored
    g = (torch.randn(T) / T) * mask
    outs = []
    for lm in (None, mask.cuda().unsqueeze(0)):
        h = hidden.cuda().requires_grad_(True)
        w = weight.cuda().requires_grad_(True)
        lp, _ = P.fused_lm_head_logprobs_entropy(h.unsqueeze(0), w, ids.cuda().unsqueeze(0), tau, loss_mask=lm, save_logits=save)
        (lp[0] * g.cuda()).sum().backward()
        outs.append((h.grad.float(), w.grad.float()))
    (dh0, dw0), (dh1, dw1) = outs
    assert ((dh1 - dh0).norm() / dh0.norm()).item() < 2e-3 and ((dw1 - dw0).norm() / dw0.norm()).item() < 2e-3
    inactive_rows = (~mask[1:]).nonzero().reshape(-1).cuda()
    assert float(dh1[inactive_rows].abs().max()) == 0.0


@pytest.mark.parametrize("T,shift,p_active", [(1, 1, 1.0), (2, 1, 0.5), (700, 1, 0.6), (1024, 0, 0.3), (1025, 1, 0.0), (4099, 1, 1.0),
                                              (40000, 1, 0.75), (33, 0, 0.5)])
def test_row_partition_gather_scatter_kernels(P, T, shift, p_active):
    """rows.py / prl_partition_positions & co. against numpy: stable partition (active positions first), row gather, token and
    row scatter with exact zeros at masked rows."""
    from prime_rl_cse_b200 import rows

    rng = np.random.default_rng(T + shift)
    mask = rng.random(T) < p_active
    H = 64
    part = rows.partition_positions(cu(mask), bool(shift))
    idx = np.arange(shift, T)
    act, ina = idx[mask[shift:]], idx[~mask[shift:]]
    want_order = np.concatenate([act, ina])
    assert part.n_active == len(act) and part.rows == T - shift
    np.testing.assert_array_equal(part.order.cpu().numpy()[: T - shift], want_order)
    want_slot = np.full(T, -1, dtype=np.int64)
    want_slot[want_order] = np.arange(T - shift)
    np.testing.assert_array_equal(part.slot.cpu().numpy()[:T], want_slot)
    hidden = torch.randn(T, H).bfloat16()
    ids = torch.from_numpy(rng.integers(0, 1000, T))
    for n in sorted({0, len(act), T - shift}):
        h_p, i_p = rows.gather_rows(hidden.cuda(), ids.cuda(), part, n)
        assert torch.equal(h_p.cpu(), hidden[want_order[:n] - shift]) and torch.equal(i_p.cpu(), ids[want_order[:n]])
        x, y = torch.randn(T), torch.randn(T)
        (x_p, y_p) = rows.gather_tokens(part, n, x.cuda(), y.cuda())
        assert torch.equal(x_p.cpu(), x[want_order[:n]]) and torch.equal(y_p.cpu(), y[want_order[:n]])
        (x_back,) = rows.scatter_tokens(part, n, (x_p, -7.0))
        want = torch.zeros(T)
        want[want_order[:n]] = x[want_order[:n]]
        want[:shift] = -7.0
        assert torch.equal(x_back.cpu(), want)
        dh = rows.scatter_rows(h_p if n > 0 else None, part, n, H)
        want_dh = torch.zeros(T, H, dtype=torch.bfloat16)
        want_dh[want_order[:n] - shift] = hidden[want_order[:n] - shift]
        assert torch.equal(dh.cpu(), want_dh)


@pytest.mark.parametrize("T,H,V,peers,self_index,vchunk", [(300, 64, 5000, 3, 1, 1024), (512, 256, 151936, 8, 5, 0), (200, 128, 4096, 2, 0, 0),
                                                            (129, 64, 1000, 8, 7, 256), (64, 64, 321, 8, 3, 0)])
def test_d_weight_scatter_into_owner_shards_emulated_on_one_gpu(P, T, H, V, peers, self_index, vchunk):
    """prl_lmhead_bwd_dw_scatter with the owners' shards as separate buffers on ONE GPU: the concatenated shards must equal the
    classic fp32 d_W accumulator bit for bit (one adder per element here), for shard sizes that are not multiples of the
    32-row epilogue slab (slabs that straddle two owners are sent to both, clipped by the tensor maps), trailing owners
    with fewer rows (torch.chunk layout), rotated vocab-chunk order, and accumulation over two micro-batches."""
    import ctypes
    from types import SimpleNamespace

    from prime_rl_cse_b200 import lm_head_backward_dw_scatter_raw, lm_head_backward_raw, lm_head_forward_raw

    hidden, weight, ids = _mk(T, H, V, T + V + peers)
    h, w, i = hidden.cuda(), weight.cuda(), ids.cuda()
    per = -(-V // peers)
    shards = [torch.zeros(max(min(per, V - j * per), 0), H, device="cuda") for j in range(peers)]
    guard = [torch.full((64, H), 7.0, device="cuda") for _ in range(peers)]  # canaries: nothing may be written past a shard
    bufs = [torch.cat([s_, g_]) for s_, g_ in zip(shards, guard)]
    ptrs = (ctypes.c_void_p * peers)(*[b.data_ptr() if shards[j].shape[0] > 0 else 0 for j, b in enumerate(bufs)])
    acc = SimpleNamespace(V=V, H=H, world=peers, rank=self_index, rows_per_peer=per, buf=SimpleNamespace(ptrs=ptrs))
    want = torch.zeros(V, H, device="cuda")
    for step in range(2):
        g = torch.randn(T, device="cuda") / T
        _, _, lse, _ = lm_head_forward_raw(h, w, i, 1.0)
        dh = lm_head_backward_dw_scatter_raw(h, w, i, lse, g, 1.0, acc, vchunk=vchunk)
        dh_ref, _ = lm_head_backward_raw(h, w, i, lse, g, 1.0, dweight=want, dweight_accumulate=True, vchunk=vchunk)
        assert ((dh.float() - dh_ref.float()).norm() / dh_ref.float().norm()).item() < 2e-3  # chunk order differs: fp32 summation order
    got = torch.cat([b[: shards[j].shape[0]] for j, b in enumerate(bufs)])
    assert got.shape == want.shape and torch.equal(got, want)
    for j, b in enumerate(bufs):
        assert float((b[shards[j].shape[0]:] - 7.0).abs().max()) == 0.0, f"shard {j} overrun"
    # the "last micro-batch sends the local total" schedule: the shards receive local accumulator + this micro-batch's tiles
    for b, s_ in zip(bufs, shards):
        b[: s_.shape[0]].zero_()
    local = want.clone()  # stands for the first n - 1 micro-batches of a step
    g = torch.randn(T, device="cuda") / T
    _, _, lse, _ = lm_head_forward_raw(h, w, i, 1.0)
    lm_head_backward_dw_scatter_raw(h, w, i, lse, g, 1.0, acc, vchunk=vchunk, local_add=local, want_dhidden=False)
    lm_head_backward_raw(h, w, i, lse, g, 1.0, dweight=want, dweight_accumulate=True, vchunk=vchunk, want_dhidden=False)
    got = torch.cat([b[: shards[j].shape[0]] for j, b in enumerate(bufs)])
    assert ((got - want).abs().max() / want.abs().max()).item() < 1e-6  # (local + tile) + 0 vs local + tile: same sum, fp32
    for j, b in enumerate(bufs):
        assert float((b[shards[j].shape[0]:] - 7.0).abs().max()) == 0.0, f"shard {j} overrun"


@pytest.mark.parametrize("rows_per_peer,H,Vs,peers,self_index,vchunk", [(96, 64, 1000, 4, 1, 0), (32, 128, 640, 8, 7, 256), (544, 256, 18992, 3, 0, 0),
                                                                          (160, 64, 333, 2, 1, 128)])
def test_d_hidden_reduce_scatter_into_owner_buffers_emulated_on_one_gpu(P, rows_per_peer, H, Vs, peers, self_index, vchunk):
    """prl_lmhead_bwd_scatter (config 4: the reduce-scatter of d_hidden fused into the GEMM epilogue) with the owners'
    accumulators as separate buffers on ONE GPU: rows [j * rows_per_peer, (j + 1) * rows_per_peer) of the concatenated sweep must
    land in owner j's buffer and nowhere else (canary rows behind every buffer), equal to the plain backward's fp32 d_hidden
    partial of this vocab shard; the d_W shard is the plain one bit for bit; a second call accumulates (the owners zero their
    buffers per round, not the kernel)."""
    import ctypes

    from prime_rl_cse_b200 import _lib, lm_head_backward_raw, lm_head_forward_raw

    lib = _lib.load()
    M, v0 = peers * rows_per_peer, 4096  # the shard covers vocab rows [v0, v0 + Vs) of a larger vocabulary
    hidden, weight, _ = _mk(M, H, Vs, M + Vs + peers)
    h, w = hidden.cuda(), weight.cuda()
    ids = torch.randint(0, v0 + Vs + 500, (M,), generator=torch.Generator().manual_seed(5)).cuda()  # targets inside and outside the shard
    g = (torch.randn(M, generator=torch.Generator().manual_seed(6)) / M).cuda()
    g[::rows_per_peer] = 0.0  # position 0 of every owner's packed row (parallel.gather_token_scalars)
    _, _, lse, _ = lm_head_forward_raw(h, w, ids, 1.0, v_offset=v0, v_total=v0 + Vs + 500)
    lse = lse + 0.5  # stands for the merged lse over all shards (any value >= the shard's own)
    want_dw = torch.zeros(Vs, H, device="cuda")
    vc = vchunk if vchunk > 0 else (Vs + 255) // 256 * 256  # 0: the whole shard in one chunk, as VocabShardedLMHead.backward does
    want_dh, _ = lm_head_backward_raw(h, w, ids, lse, g, 1.0, v_offset=v0, dweight=want_dw, dweight_accumulate=True, vchunk=vc, dhidden_f32=True)
    bufs = [torch.cat([torch.zeros(rows_per_peer, H), torch.full((32, H), 7.0)]).cuda() for _ in range(peers)]
    ptrs = (ctypes.c_void_p * peers)(*[b.data_ptr() for b in bufs])
    dw = torch.zeros(Vs, H, device="cuda")
    nb = lib.prl_lmhead_bwd_workspace_bytes(M, H, Vs, vc)
    ws = torch.empty(max(int(nb), 16), dtype=torch.uint8, device="cuda")
    for rep in (1, 2):
        rc = lib.prl_lmhead_bwd_scatter(_lib.ptr(h), _lib.ptr(w), _lib.ptr(ids), _lib.ptr(lse), _lib.ptr(g), M, H, Vs, v0, 1.0, 1, 1, ptrs, peers,
                                        rows_per_peer, self_index, _lib.ptr(dw), _lib.PRL_DT_F32, 1, vc, 0, _lib.ptr(ws), ws.numel(), _lib.stream())
        _lib.check(rc, "prl_lmhead_bwd_scatter")
        got = torch.cat([b[:rows_per_peer] for b in bufs])
        # same products, but the fp32 addends of an element are combined in another order (the plain path may split K in two for
        # few output tiles, vocab chunks arrive in tile order, and the tensor core truncates its accumulator per K = 16 step):
        # equal to accumulation rounding -- a mis-routed slab would be an O(1) error
        assert ((got - rep * want_dh).norm() / want_dh.norm().clamp_min(1e-30)).item() < 1e-4
        for j, b in enumerate(bufs):
            assert float((b[rows_per_peer:] - 7.0).abs().max()) == 0.0, f"owner {j} overrun"
    assert ((dw - 2 * want_dw).abs().max() / want_dw.abs().max().clamp_min(1e-30)).item() < 1e-5
    # a row count that is not a multiple of the 32-row epilogue slab is refused, not mis-routed
    rc = lib.prl_lmhead_bwd_scatter(_lib.ptr(h), _lib.ptr(w), _lib.ptr(ids), _lib.ptr(lse), _lib.ptr(g), peers * 40, H, Vs, v0, 1.0, 1, 1, ptrs, peers, 40,
                                    self_index, _lib.ptr(dw), _lib.PRL_DT_F32, 1, vc, 0, _lib.ptr(ws), ws.numel(), _lib.stream())
    assert rc != 0


def test_micro_step_is_cuda_graph_capturable(P):
    """With the host-side segment lengths and active count attached by the loader (rollout_io.upload_micro_batches) the whole
    micro-step -- partition, gather, fused forward, packed loss, autograd backward through the recompute GEMMs, fp32 d_W
    accumulation -- issues no device-to-host synchronisation and no allocation outside torch's caching allocator, so it can be
    captured in a CUDA graph and replayed on new data in the same buffers: the replay must reproduce the eager results."""
    from prime_rl_cse_b200.segments import response_lengths_host

    T, H, V, tau = 640, 256, 4096, 0.8
    lens = [200, 248, 192]
    pos = torch.cat([torch.arange(n) for n in lens])
    mask = torch.ones(T, dtype=torch.bool)
    mask[:50] = False
    mask[200:260] = False
    mask[448:470] = False
    extra = {"response_lengths": response_lengths_host(pos.unsqueeze(0)), "n_active": int(mask[1:].sum())}
    assert extra["response_lengths"] == lens

    def data(seed):
        hidden, weight, ids = _mk(T, H, V, seed)
        g = torch.Generator().manual_seed(seed)
        adv = torch.cat([torch.full((n,), float(a)) for n, a in zip(lens, torch.randn(3, generator=g))])
        old = -torch.rand(T, generator=g) * 2
        return hidden, weight, ids, adv, old

    static = {k: v.cuda() for k, v in zip(("hidden", "weight", "ids", "adv", "old"), data(1))}
    h = static["hidden"].clone().requires_grad_(True)
    w = static["weight"]
    acc = torch.zeros(V, H, device="cuda")
    mb = dict(input_ids=static["ids"].unsqueeze(0), position_ids=pos.cuda().unsqueeze(0), advantages=static["adv"].unsqueeze(0),
              logprobs=static["old"].unsqueeze(0), loss_mask=mask.cuda().unsqueeze(0), temperature=tau, **extra)
    cfg, scale = P.LossConfig(), int(mask.sum())

    def step():
        h.grad = None
        out = P.rl_micro_step(h.unsqueeze(0), w.requires_grad_(True), mb, cfg, scale, main_grad=acc)
        return out["loss"], out["logprobs"], out["entropy"], h.grad

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # warm-up outside the capture (library knobs, function attributes, autograd's buffers)
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_loss, g_lp, g_ent, g_dh = step()
    for seed in (2, 3):
        new = data(seed)
        with torch.no_grad():
            for k, v in zip(("hidden", "ids", "adv", "old"), (new[0], new[2], new[3], new[4])):
                static[k].copy_(v.cuda())
            w.copy_(new[1].cuda())
            h.copy_(static["hidden"])
        acc.zero_()
        graph.replay()
        torch.cuda.synchronize()
        got = (g_loss.clone(), g_lp.clone(), g_ent.clone(), g_dh.clone(), acc.clone())
        acc.zero_()
        e_loss, e_lp, e_ent, e_dh = step()
        torch.cuda.synchronize()
        assert got[0].item() == e_loss.item() and torch.equal(got[1], e_lp) and torch.equal(got[2], e_ent)
        assert torch.equal(got[3], e_dh) and torch.equal(got[4], acc)
        assert float(got[3][:49].abs().max()) == 0.0 and float(got[3].abs().max()) > 0.0  # masked prompt rows: exact zeros


@pytest.mark.parametrize("save", [False, True], ids=["recompute", "saved-logits"])
def test_masked_micro_step_edge_masks(P, save):
    """all-masked rows (the orchestrator's padding batches, batch.py:153-161), a single active token, and an all-active row
    through rl_micro_step: exact zeros where nothing is active, same results as the unpartitioned path otherwise"""
    T, H, V, tau = 260, 64, 1000, 0.8
    hidden, weight, ids = _mk(T, H, V, 77)
    rng = np.random.default_rng(0)
    pos = np.concatenate([np.arange(100), np.arange(160)])
    adv = rng.standard_normal(T).astype(np.float32)
    old = -rng.random(T).astype(np.float32) * 3
    for kind in ("none", "one", "all"):
        mask = np.zeros(T, dtype=bool) if kind != "all" else np.ones(T, dtype=bool)
        if kind == "one":
            mask[137] = True
        res = []
        for skip in (False, True):
            h = hidden.cuda().requires_grad_(True)
            w = weight.cuda().requires_grad_(True)
            mb = dict(input_ids=ids.cuda().unsqueeze(0), position_ids=cu(pos).unsqueeze(0), advantages=cu(adv).unsqueeze(0),
                      logprobs=cu(old).unsqueeze(0), loss_mask=cu(mask).unsqueeze(0), temperature=tau)
            out = P.rl_micro_step(h.unsqueeze(0), w, mb, P.LossConfig(), max(int(mask.sum()), 1), skip_masked=skip, save_logits=save)
            res.append((out, h.grad.float(), w.grad.float()))
        (o0, dh0, dw0), (o1, dh1, dw1) = res
        assert torch.equal(o0["logprobs"], o1["logprobs"]) and torch.equal(o0["entropy"], o1["entropy"])
        assert o0["loss"].item() == o1["loss"].item()
        if kind == "none":
            assert float(dh1.abs().max()) == 0.0 and float(dw1.abs().max()) == 0.0 and o1["loss"].item() == 0.0
        else:
            assert ((dh1 - dh0).norm() / dh0.norm()).item() < 2e-3 and ((dw1 - dw0).norm() / dw0.norm()).item() < 2e-3
        if kind == "one":
            assert int((dh1.abs().sum(dim=1) > 0).sum()) == 1 and float(dh1[136].abs().max()) > 0.0  # only source row 136


@pytest.mark.parametrize("save", [False, True], ids=["recompute", "saved-logits"])
def test_forward_active_only(P, save):
    """Opt-in: forward restricted to the unmasked positions gives the same values there (and the same loss / gradients)."""
    T, H, V, tau = 700, 256, 6000, 0.7
    hidden, weight, ids = _mk(T, H, V, 41)
    rng = np.random.default_rng(4)
    mask = rng.random(T) < 0.55
    mask[0] = True
    lens = [300, 400]
    pos = np.concatenate([np.arange(n) for n in lens])
    adv = rng.standard_normal(T).astype(np.float32)
    old = -rng.random(T).astype(np.float32) * 3
    mb = dict(input_ids=ids.cuda().unsqueeze(0), position_ids=cu(pos).unsqueeze(0), advantages=cu(adv).unsqueeze(0),
              logprobs=cu(old).unsqueeze(0), loss_mask=cu(mask).unsqueeze(0), temperature=tau)
    res = []
    for skip_fwd in (False, True):
        h = hidden.cuda().requires_grad_(True)
        w = weight.cuda().requires_grad_(True)
        out = P.rl_micro_step(h.unsqueeze(0), w, mb, P.LossConfig(), int(mask.sum()), forward_skips_masked=skip_fwd, save_logits=save)
        res.append((out, h.grad.float(), w.grad.float()))
    (o0, dh0, dw0), (o1, dh1, dw1) = res
    m = torch.from_numpy(mask).cuda()
    assert (o0["logprobs"][0][m] - o1["logprobs"][0][m]).abs().max().item() < 2e-4
    assert (o0["entropy"][0][m] - o1["entropy"][0][m]).abs().max().item() < 2e-4
    assert float(o1["logprobs"][0][~m].abs().max()) == 0.0
    assert abs(o0["loss"].item() - o1["loss"].item()) <= 1e-4 * abs(o0["loss"].item()) + 1e-7
    assert ((dh1 - dh0).norm() / dh0.norm()).item() < 2e-3 and ((dw1 - dw0).norm() / dw0.norm()).item() < 2e-3


def test_loader_uploads_one_buffer_per_micro_batch(P):
    """The trainer-side data format: the rank file written by the reference (tests/golden/rollouts) streamed to the device with
    one pinned staging buffer + one H2D copy per micro-batch (replaces train.py:275-279), then through the fused micro-step."""
    dl = P.DataLoader(G, 7, rank=0)
    host = P.DataLoader(G, 7, rank=0).get_batch()
    dev = list(dl.get_batch_on_device("cuda"))
    assert len(dev) == len(host) and dl.current_step == 8
    H, V = 64, 1000
    gen = torch.Generator().manual_seed(9)
    w = (torch.randn(V, H, generator=gen) * 0.4).bfloat16().cuda().requires_grad_(True)
    for a, b in zip(dev, host):
        for k in ("input_ids", "position_ids", "advantages", "logprobs", "loss_mask"):
            assert a[k].is_cuda and a[k].dtype == b[k].dtype and torch.equal(a[k].cpu(), b[k])
        assert a["temperature"] == b["temperature"]
        T = a["input_ids"].shape[1]
        h = torch.randn(1, T, H, generator=gen).bfloat16().cuda().requires_grad_(True)
        ref = P.rl_micro_step(h, w, P.to_device(b, "cuda"), P.LossConfig(), 7, backward=False)
        out = P.rl_micro_step(h, w, a, P.LossConfig(), 7, backward=False)
        assert torch.equal(out["logprobs"], ref["logprobs"]) and torch.equal(out["loss"], ref["loss"])
        # the loader attached what it saw on the host, so that the micro-step needs no device-to-host sync: same values as the
        # device kernels give, and the same gradients with the backward over the active positions only
        assert a["response_lengths"] == P.get_response_lengths(a["position_ids"]) and a["n_active"] == int(b["loss_mask"][0, 1:].sum())
        g = []
        for mb in (a, P.to_device(b, "cuda")):
            h2 = h.detach().clone().requires_grad_(True)
            w2 = w.detach().clone().requires_grad_(True)
            P.rl_micro_step(h2, w2, mb, P.LossConfig(), 7)
            g.append((h2.grad, w2.grad))
        assert torch.equal(g[0][0], g[1][0]) and torch.equal(g[0][1], g[1][1])


def test_next_row_prepare_batch_on_device(P):
    """`prepare_batch` drop-in (host first-fit-decreasing + one packing kernel) reproduces the reference
    orchestrator's micro-batches bit for bit, including the padding micro-batch and the per-worker split."""
    from types import SimpleNamespace

    d = load("packed_layout.npz")
    pl, cl = d["ro/prompt_len"], d["ro/completion_len"]
    po, co = np.concatenate([[0], np.cumsum(pl)]), np.concatenate([[0], np.cumsum(cl)])
    rollouts = []
    for i in range(len(pl)):
        rollouts.append(SimpleNamespace(
            prompt_tokens=d["ro/prompt_tokens"][po[i]:po[i + 1]].tolist(), prompt_mask=d["ro/prompt_mask"][po[i]:po[i + 1]].tolist(),
            completion_tokens=d["ro/completion_tokens"][co[i]:co[i + 1]].tolist(),
            completion_mask=d["ro/completion_mask"][co[i]:co[i + 1]].tolist(),
            completion_logprobs=d["ro/completion_logprobs"][co[i]:co[i + 1]].tolist(), advantage=float(d["ro/advantage"][i])))
    temperature, batch_size, mbs, seq_len, nw = d["ro/args"]
    out = P.prepare_batch(rollouts, float(temperature), None, int(batch_size), int(mbs), int(seq_len), int(nw))
    assert len(out) == int(d["n_ranks"])
    for r, per_rank in enumerate(out):
        assert len(per_rank) == int(d[f"rank{r}/n"])
        for i, mb in enumerate(per_rank):
            assert mb["temperature"] == float(d[f"rank{r}/mb{i}/temperature"])
            for k in ("input_ids", "position_ids", "advantages", "logprobs", "loss_mask"):
                ref = d[f"rank{r}/mb{i}/{k}"]
                got = mb[k].numpy()
                assert got.dtype == ref.dtype and got.shape == ref.shape, (k, got.dtype, ref.dtype, got.shape, ref.shape)
                np.testing.assert_array_equal(got, ref, err_msg=f"rank{r}/mb{i}/{k}")
    with pytest.raises(ValueError):
        P.prepare_batch(rollouts, 1.0, None, 12, 1, 8, 2)  # a rollout longer than seq_len * micro_batch_size
