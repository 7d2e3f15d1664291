"""Fused LM head: hidden x W_lm^T -> (log-prob of the sampled token, entropy) without materialising
the [T, V] logits.  Replaces rl/train.py:285-288,302-303 of the reference
(src/prime_rl/trainer/rl/train.py) and the autograd chain behind `loss.backward()` (:310).

Two backward modes, same results:
  * recompute (DEFAULT; BASELINE.json north_star's design): nothing but lse[T] is kept, the backward recomputes the logits
    per vocab chunk (one more GEMM) -- the [T, V] logits never exist in HBM;
  * saved logits (OPT-IN: `save_logits=True` or PRL_SAVE_LOGITS=1|auto): the forward keeps ONE bf16 copy of the logits
    (T x V x 2 bytes, 4.6 GiB for a 16384-token micro-batch of Qwen3; the reference keeps the fp32 logits plus their
    gradient), the backward rewrites it in place into dlogits and runs two GEMMs: ~25 % faster, at that memory cost."""

from __future__ import annotations

import os
import warnings

import torch
from torch import nn

from . import _lib, rows


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _check_operands(hidden, weight, ids, what: str, **per_token) -> None:
    """The kernels read raw pointers through TMA descriptors: anything but contiguous bf16 [T, H] / [V, H] operands and
    contiguous int64 ids [T] would be silently reinterpreted, so every wrapper that reaches prl_lmhead_* checks them here."""
    _lib.require_cuda(hidden, weight, ids, *per_token.values())
    if hidden.dim() != 2 or weight.dim() != 2 or hidden.shape[1] != weight.shape[1]:
        raise ValueError(f"{what}: hidden must be [T, H] and weight [V, H] (got {tuple(hidden.shape)} and {tuple(weight.shape)})")
    for name, t in (("hidden", hidden), ("weight", weight)):
        if t.dtype != torch.bfloat16:
            raise ValueError(f"{what}: {name} must be bf16 (got {t.dtype}); cast it explicitly (FSDP param_dtype=bf16, model.py:103)")
        if not t.is_contiguous() or t.data_ptr() % 16 != 0:
            raise ValueError(f"{what}: {name} must be contiguous and 16-byte aligned")
    T = hidden.shape[0]
    if ids.dtype != torch.int64 or ids.dim() != 1 or ids.shape[0] != T or not ids.is_contiguous():
        raise ValueError(f"{what}: ids must be a contiguous int64 [T] tensor (T = {T})")
    for name, t in per_token.items():
        if t is not None and (t.dtype != torch.float32 or t.numel() != T or not t.is_contiguous()):
            raise ValueError(f"{what}: {name} must be a contiguous fp32 tensor with T = {T} elements")


def lm_head_forward_raw(hidden, weight, ids, inv_temp, shift=True, round_bf16=True, v_offset=0, v_total=None,
                        want_partials=False, workspace=None):
    """hidden [T,H] bf16, weight [V,H] bf16, ids [T] int64 -> (logp, entropy, lse, partials|None)."""
    _check_operands(hidden, weight, ids, "lm_head_forward_raw")
    lib = _lib.load()
    T, H = hidden.shape
    V = weight.shape[0]
    v_total = V if v_total is None else v_total
    dev = hidden.device
    logp = torch.empty(T, dtype=torch.float32, device=dev)
    ent = torch.empty(T, dtype=torch.float32, device=dev)
    lse = torch.empty(T, dtype=torch.float32, device=dev)
    parts = torch.empty(T, 4, dtype=torch.float32, device=dev) if want_partials else None
    nb = lib.prl_lmhead_fwd_workspace_bytes(T, H, V)
    ws = workspace if workspace is not None and workspace.numel() >= nb else _ws(nb, dev)
    rc = lib.prl_lmhead_fwd(_lib.ptr(hidden), _lib.ptr(weight), _lib.ptr(ids), T, H, V, v_offset, v_total, float(inv_temp),
                            1 if shift else 0, 1 if round_bf16 else 0, _lib.ptr(logp), _lib.ptr(ent), _lib.ptr(lse),
                            _lib.ptr(parts), _lib.ptr(ws), ws.numel(), _lib.stream())
    _lib.check(rc, "prl_lmhead_fwd")
    return logp, ent, lse, parts


def saved_logits_policy(T: int, V: int, override: bool | None = None, device=None) -> bool:
    """Whether the forward keeps the bf16 logits for the backward.  `override` wins; else the environment variable
    PRL_SAVE_LOGITS = 0 (default: recompute, nothing of size [T, V] is ever stored) | 1 | auto.  `auto` keeps them when the
    block is at most PRL_SAVE_LOGITS_MAX_GIB (default 12 GiB: a 32768-token packed row at V = 151936 is 9.3 GiB) AND at
    most half of the device memory that is free right now."""
    if override is not None:
        return bool(override)
    env = os.environ.get("PRL_SAVE_LOGITS", "0").strip().lower()
    if env in ("0", "off", "false", "no", ""):
        return False
    if env in ("1", "on", "true", "yes"):
        return True
    need = T * saved_logits_ld(V) * 2
    limit = float(os.environ.get("PRL_SAVE_LOGITS_MAX_GIB", "12")) * 2**30
    if not 0 < need <= limit:
        return False
    if torch.cuda.is_available():
        free, _total = torch.cuda.mem_get_info(device)
        return need <= free // 2
    return True


_warned_fallback = False


def _warn_fallback(why: str) -> None:
    """one line, once per process: the saved-logits backward was asked for but the recompute backward runs (~25 % slower)"""
    global _warned_fallback
    if not _warned_fallback:
        _warned_fallback = True
        warnings.warn(f"prime_rl_cse_b200: saved-logits backward unavailable ({why}); using the recompute backward "
                      "(one more GEMM per micro-batch). Reported once.", RuntimeWarning, stacklevel=3)


def saved_logits_ld(V: int) -> int:
    """Row pitch (elements) of the saved-logits block (prl_lmhead_saved_logits_ld)."""
    return int(_lib.load().prl_lmhead_saved_logits_ld(int(V)))


def lm_head_forward_save_raw(hidden, weight, ids, inv_temp, shift=True, v_offset=0, v_total=None, want_partials=False,
                             workspace=None, saved=None):
    """`lm_head_forward_raw` (bf16-rounded logits) that also keeps the logits: returns (logp, entropy, lse,
    partials|None, saved [T, ld] bf16).  `saved` may be a preallocated [T, ld] bf16 block, ld = saved_logits_ld(V)."""
    _check_operands(hidden, weight, ids, "lm_head_forward_save_raw")
    lib = _lib.load()
    T, H = hidden.shape
    V = weight.shape[0]
    v_total = V if v_total is None else v_total
    dev = hidden.device
    ld = saved_logits_ld(V)
    if saved is None:
        saved = torch.empty(T, ld, dtype=torch.bfloat16, device=dev)
    elif saved.dtype != torch.bfloat16 or saved.dim() != 2 or saved.shape[0] < T or saved.shape[1] < V or not saved.is_contiguous():
        raise ValueError("saved logits block must be a contiguous bf16 [>= T, >= V] tensor")
    logp = torch.empty(T, dtype=torch.float32, device=dev)
    ent = torch.empty(T, dtype=torch.float32, device=dev)
    lse = torch.empty(T, dtype=torch.float32, device=dev)
    parts = torch.empty(T, 4, dtype=torch.float32, device=dev) if want_partials else None
    nb = lib.prl_lmhead_fwd_workspace_bytes(T, H, V)
    ws = workspace if workspace is not None and workspace.numel() >= nb else _ws(nb, dev)
    rc = lib.prl_lmhead_fwd_save(_lib.ptr(hidden), _lib.ptr(weight), _lib.ptr(ids), T, H, V, v_offset, v_total, float(inv_temp),
                                 1 if shift else 0, _lib.ptr(logp), _lib.ptr(ent), _lib.ptr(lse), _lib.ptr(parts),
                                 _lib.ptr(saved), saved.shape[1], _lib.ptr(ws), ws.numel(), _lib.stream())
    _lib.check(rc, "prl_lmhead_fwd_save")
    return logp, ent, lse, parts, saved


def _dweight_arg(dweight, want_dweight, dweight_accumulate, V, H, dev):
    if want_dweight and dweight is None:
        dweight = torch.empty(V, H, dtype=torch.bfloat16, device=dev)
        dweight_accumulate = False
    dt = _lib.PRL_DT_BF16
    if dweight is not None:
        if dweight.dtype == torch.float32:
            dt = _lib.PRL_DT_F32
        elif dweight.dtype != torch.bfloat16:
            raise ValueError("d_weight must be bf16 or fp32")
        if tuple(dweight.shape) != (V, H) or not dweight.is_contiguous():
            raise ValueError("d_weight must be a contiguous [V, H] tensor")
    return dweight, dt, dweight_accumulate


def lm_head_backward_saved_raw(hidden, weight, ids, lse, dlogp, inv_temp, saved, shift=True, v_offset=0, want_dhidden=True,
                               dweight=None, dweight_accumulate=False, want_dweight=True, dhidden_f32=False, phase_mask=0,
                               workspace=None):
    """Backward from the logits kept by `lm_head_forward_save_raw`.  `saved` is overwritten (it holds the bf16 dlogits
    afterwards).  Returns (d_hidden, d_weight) like `lm_head_backward_raw`."""
    _check_operands(hidden, weight, ids, "lm_head_backward_saved_raw", lse=lse, dlogp=dlogp)
    lib = _lib.load()
    T, H = hidden.shape
    V = weight.shape[0]
    dev = hidden.device
    if saved.dtype != torch.bfloat16 or saved.dim() != 2 or saved.shape[0] < T or saved.shape[1] < V or not saved.is_contiguous():
        raise ValueError("saved logits block must be a contiguous bf16 [>= T, >= V] tensor")
    dh = None
    if want_dhidden:
        dh = torch.empty(T, H, dtype=torch.float32 if dhidden_f32 else torch.bfloat16, device=dev)
    dweight, dt, dweight_accumulate = _dweight_arg(dweight, want_dweight, dweight_accumulate, V, H, dev)
    nb = lib.prl_lmhead_bwd_saved_workspace_bytes(T, H, V) if want_dhidden else 0
    ws = None
    if nb > 0:
        ws = workspace if workspace is not None and workspace.numel() >= nb else _ws(nb, dev)
    rc = lib.prl_lmhead_bwd_saved(_lib.ptr(hidden), _lib.ptr(weight), _lib.ptr(ids), _lib.ptr(lse), _lib.ptr(dlogp), T, H, V, v_offset,
                                  float(inv_temp), 1 if shift else 0, _lib.ptr(saved), saved.shape[1], _lib.ptr(dh),
                                  1 if dhidden_f32 else 0, _lib.ptr(dweight), dt, 1 if dweight_accumulate else 0, phase_mask,
                                  _lib.ptr(ws), ws.numel() if ws is not None else 0, _lib.stream())
    _lib.check(rc, "prl_lmhead_bwd_saved")
    return dh, dweight


def lm_head_backward_raw(hidden, weight, ids, lse, dlogp, inv_temp, shift=True, round_bf16=True, v_offset=0,
                         want_dhidden=True, dweight=None, dweight_accumulate=False, want_dweight=True, vchunk=0,
                         dhidden_f32=False, phase_mask=0, workspace=None):
    """Returns (d_hidden [T,H] bf16|fp32|None, d_weight [V,H]).  `dweight` may be a preallocated bf16 or
    fp32 [V,H] buffer (with `dweight_accumulate` the kernel adds into it: fused gradient accumulation)."""
    _check_operands(hidden, weight, ids, "lm_head_backward_raw", lse=lse, dlogp=dlogp)
    lib = _lib.load()
    T, H = hidden.shape
    V = weight.shape[0]
    dev = hidden.device
    dh = None
    if want_dhidden:
        dh = torch.empty(T, H, dtype=torch.float32 if dhidden_f32 else torch.bfloat16, device=dev)
    dweight, dt, dweight_accumulate = _dweight_arg(dweight, want_dweight, dweight_accumulate, V, H, dev)
    nb = lib.prl_lmhead_bwd_workspace_bytes(T, H, V, vchunk)
    ws = workspace if workspace is not None and workspace.numel() >= nb else _ws(nb, dev)
    rc = lib.prl_lmhead_bwd(_lib.ptr(hidden), _lib.ptr(weight), _lib.ptr(ids), _lib.ptr(lse), _lib.ptr(dlogp), T, H, V, v_offset,
                            float(inv_temp), 1 if shift else 0, 1 if round_bf16 else 0, _lib.ptr(dh), 1 if dhidden_f32 else 0,
                            _lib.ptr(dweight), dt, 1 if dweight_accumulate else 0, vchunk, phase_mask, _lib.ptr(ws), ws.numel(), _lib.stream())
    _lib.check(rc, "prl_lmhead_bwd")
    return dh, dweight


def lm_head_backward_dw_scatter_raw(hidden, weight, ids, lse, dlogp, inv_temp, acc, shift=True, round_bf16=True, want_dhidden=True,
                                    vchunk=0, dhidden_f32=False, phase_mask=0, workspace=None, local_add=None):
    """Recompute backward whose d_weight GEMM adds its tiles straight into the ranks' shards of `acc` (a
    `peer.ShardedGradAccumulator`: d_W_lm delivered in the reduce-scattered layout, no step-end collective).  `local_add`
    (fp32 [V, H]): a local accumulator whose tiles are added to what is sent.  Returns d_hidden."""
    _check_operands(hidden, weight, ids, "lm_head_backward_dw_scatter_raw", lse=lse, dlogp=dlogp)
    lib = _lib.load()
    T, H = hidden.shape
    V = weight.shape[0]
    if V != acc.V or H != acc.H:
        raise ValueError(f"accumulator is for a [{acc.V}, {acc.H}] gradient, the weight is [{V}, {H}]")
    if local_add is not None and (local_add.dtype != torch.float32 or tuple(local_add.shape) != (V, H) or not local_add.is_contiguous()):
        raise ValueError("local_add must be a contiguous fp32 [V, H] tensor")
    dh = torch.empty(T, H, dtype=torch.float32 if dhidden_f32 else torch.bfloat16, device=hidden.device) if want_dhidden else None
    nb = lib.prl_lmhead_bwd_workspace_bytes(T, H, V, vchunk)
    ws = workspace if workspace is not None and workspace.numel() >= nb else _ws(nb, hidden.device)
    rc = lib.prl_lmhead_bwd_dw_scatter(_lib.ptr(hidden), _lib.ptr(weight), _lib.ptr(ids), _lib.ptr(lse), _lib.ptr(dlogp), T, H, V,
                                       float(inv_temp), 1 if shift else 0, 1 if round_bf16 else 0, _lib.ptr(dh), 1 if dhidden_f32 else 0,
                                       acc.buf.ptrs, acc.world, acc.rows_per_peer, acc.rank, _lib.ptr(local_add), vchunk, phase_mask,
                                       _lib.ptr(ws), ws.numel(), _lib.stream())
    _lib.check(rc, "prl_lmhead_bwd_dw_scatter")
    return dh


def _is_sharded_accumulator(main_grad) -> bool:
    return main_grad is not None and not torch.is_tensor(main_grad) and hasattr(main_grad, "rows_per_peer")


def _forward_any(h, w, i, inv_temp, shift, round_bf16, save):
    """(logp, ent, lse, saved|None): saved-logits forward when asked for and possible, else the plain one."""
    if save and round_bf16 and h.shape[0] > 0:
        try:
            logp, ent, lse, _, saved = lm_head_forward_save_raw(h, w, i, inv_temp, shift)
            return logp, ent, lse, saved
        except torch.OutOfMemoryError:
            _warn_fallback(f"no room for the {h.shape[0]} x {w.shape[0]} bf16 logits block")  # recompute mode needs lse only
    logp, ent, lse, _ = lm_head_forward_raw(h, w, i, inv_temp, shift, round_bf16)
    return logp, ent, lse, None


class _FusedLMHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, weight, ids, inv_temp, shift, round_bf16, main_grad, partition=None, forward_active_only=False,
                save_logits=None):
        lead = hidden.shape[:-1]
        h2 = hidden.detach().reshape(-1, hidden.shape[-1]).contiguous()
        w2 = weight.detach().contiguous()
        i2 = ids.detach().reshape(-1).contiguous().to(torch.int64)
        T, V = h2.shape[0], w2.shape[0]
        ctx.part = None           # rows.Partition when the row was processed in "active positions first" order
        ctx.saved_logits = None   # bf16 logits kept for the backward (consumed by it)
        ctx.had_saved = False
        needs_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        if _is_sharded_accumulator(main_grad):
            if save_logits:
                _warn_fallback("d_W_lm is accumulated in a ShardedGradAccumulator, which the recompute backward feeds")
            save_logits = False
        # Masked positions get no gradient (loss.py:104): with a partition of the row (active positions first, rows.py) the
        # backward -- and optionally the forward -- only visits the first n_active slots.  Nothing to gain when every
        # position is active, or when no backward will follow and the forward has to cover the whole row anyway.
        if partition is not None and T > 0 and (needs_grad or forward_active_only) and partition.n_active < partition.rows:
            part = partition
            n_fwd = part.n_active if forward_active_only else part.rows
            # log-prob / entropy are 0 at positions the forward skipped (nothing downstream reads them: the loss multiplies by
            # the mask and the trainer logs `[loss_mask]` slices only, rl/train.py:313-333); position 0 of a shifted row is the
            # closed form of shift_logits' all-zero row (rl/loss.py:55-56)
            lp_first, ent_first, lse_first = rows.first_row_values(V)
            if n_fwd > 0:
                h_p, i_p = rows.gather_rows(h2, i2, part, n_fwd)
                save = needs_grad and saved_logits_policy(n_fwd, V, save_logits)
                lp_p, ent_p, lse_p, ctx.saved_logits = _forward_any(h_p, w2, i_p, inv_temp, False, round_bf16, save)
                logp, ent = rows.scatter_tokens(part, n_fwd, (lp_p, lp_first), (ent_p, ent_first))
            else:
                h_p = h2.new_empty(0, h2.shape[1])
                i_p = i2.new_empty(0)
                lse_p = torch.empty(0, dtype=torch.float32, device=h2.device)
                logp, ent = rows.scatter_tokens(part, 0, (None, lp_first), (None, ent_first))
            ctx.part = part
            ctx.save_for_backward(h_p, w2, i_p, lse_p)
        else:
            save = needs_grad and saved_logits_policy(T, V, save_logits)
            logp, ent, lse, ctx.saved_logits = _forward_any(h2, w2, i2, inv_temp, shift, round_bf16, save)
            ctx.save_for_backward(h2, w2, i2, lse)
        ctx.cfg = (inv_temp, shift, round_bf16, hidden.shape, main_grad)
        logp, ent = logp.reshape(lead), ent.reshape(lead)
        ctx.mark_non_differentiable(ent)
        return logp, ent

    @staticmethod
    def backward(ctx, g_logp, _g_ent):
        h2, w2, i2, lse = ctx.saved_tensors
        inv_temp, shift, round_bf16, hshape, main_grad = ctx.cfg
        g = g_logp.detach().reshape(-1).contiguous().float()
        need_h, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        none_tail = (None,) * 8
        # the kept logits are consumed by the first backward; a second one (retain_graph) falls back to the recompute
        if ctx.saved_logits is None and ctx.had_saved:
            _warn_fallback("second backward through the same graph: the block was consumed by the first")
        ctx.had_saved = ctx.had_saved or ctx.saved_logits is not None
        saved, ctx.saved_logits = ctx.saved_logits, None
        dw_target = main_grad if (main_grad is not None and need_w) else None

        def run(h, i, l, gg, sh, block):
            if _is_sharded_accumulator(dw_target):  # d_weight goes to the owners' shards (directly, or through the local total)
                return dw_target.accumulate(h, w2, i, l, gg, inv_temp, sh, round_bf16, want_dhidden=need_h), None
            kw = dict(want_dhidden=need_h, dweight=dw_target, dweight_accumulate=dw_target is not None, want_dweight=need_w)
            if block is not None:
                return lm_head_backward_saved_raw(h, w2, i, l, gg, inv_temp, block, sh, **kw)
            return lm_head_backward_raw(h, w2, i, l, gg, inv_temp, sh, round_bf16, **kw)

        if ctx.part is not None:
            # Only the active positions contribute to d_hidden and d_W: they are the first n slots of the partitioned row the
            # forward saved (exactly the same sums, minus the all-zero terms); d_hidden goes back to source-row order.
            part, n = ctx.part, ctx.part.n_active
            dh = dw = None
            if n > 0:
                (g_p,) = rows.gather_tokens(part, n, g)
                dh_c, dw = run(h2[:n], i2[:n], lse[:n], g_p, False, saved)
            elif need_w and dw_target is None:
                dw = torch.zeros_like(w2)
            elif _is_sharded_accumulator(dw_target):
                dw_target.skip(w2)  # nothing to add, but the micro-batch counts towards the accumulator's step schedule
            if need_h:
                dh = rows.scatter_rows(dh_c if n > 0 else None, part, n, w2.shape[1])
            if dw_target is not None:
                dw = None
            return ((dh.reshape(hshape) if dh is not None else None), dw) + none_tail
        dh, dw = run(h2, i2, lse, g, shift, saved)
        if dw_target is not None:
            dw = None
        return ((dh.reshape(hshape) if dh is not None else None), dw) + none_tail


def fused_lm_head_logprobs_entropy(
    hidden: torch.Tensor,
    weight: torch.Tensor,
    input_ids: torch.Tensor,
    temperature: float,
    *,
    shift: bool = True,
    round_bf16: bool = True,
    main_grad: torch.Tensor | None = None,
    loss_mask: torch.Tensor | rows.Partition | None = None,
    forward_active_only: bool = False,
    save_logits: bool | None = None,
) -> tuple[torch.Tensor, torch.Tensor]:
    """hidden [..., T, H] bf16 (final-norm output), weight [V, H] bf16 (`lm_head.weight`), input_ids
    [..., T] -> (logprobs [..., T] fp32, differentiable w.r.t. hidden and weight; entropy [..., T] fp32,
    detached).  Equivalent to the reference's
        logits = lm_head(hidden).float(); shifted = shift_logits(logits) / temperature
        logprobs = selective_log_softmax(shifted, input_ids); entropy = compute_entropy(shifted)
    (rl/train.py:285-288,302-303) for one packed row.  `shift` applies to the whole flattened row as the
    reference's `shift_logits` does for batch size 1 (packed micro-batches, rl/data.py:13-23).
    `main_grad`: optional fp32/bf16 [V,H] buffer; the weight gradient is then ACCUMULATED into it by the
    kernel and `weight.grad` is left untouched (fused gradient accumulation across micro-batches).  It may also be a
    `peer.ShardedGradAccumulator`: the gradient is then accumulated directly in the reduce-scattered layout across the ranks of
    the node, inside the d_weight GEMM's epilogue (recompute backward only; no step-end reduce-scatter is left).
    `loss_mask`: optional bool [..., T] tensor (the micro-batch's `loss_mask`) or a `rows.Partition` built from it: positions
    with mask False receive no gradient from the loss (rl/loss.py:104), so the backward skips their rows (prompt / environment
    tokens) — identical gradients, proportionally less work; the row is reordered "active positions first" by the library's own
    kernels (rows.py).  By default the forward still covers all T positions, as the reference does; with
    `forward_active_only=True` it visits the active positions only and returns 0 for log-prob and entropy elsewhere.
    `save_logits`: False / None (default) recomputes the logits in the backward, so nothing of size [T, V] is ever stored;
    True keeps one bf16 copy of the logits (T x V x 2 bytes) between forward and backward so that the backward needs two
    GEMMs instead of three (opt-in; None follows PRL_SAVE_LOGITS, `saved_logits_policy`).  Both give the same gradients
    (same bf16 dlogits)."""
    _lib.require_cuda(hidden, weight, input_ids)
    if hidden.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise ValueError("fused_lm_head_logprobs_entropy needs bf16 hidden and weight (FSDP param_dtype=bf16, model.py:103)")
    if hidden.shape[:-1] != input_ids.shape:
        raise ValueError("hidden [..., T, H] and input_ids [..., T] disagree")
    if not float(temperature) > 0.0:
        raise ValueError("temperature must be positive (the online softmax tracks the maximum logit)")
    if hidden.dim() > 2 and hidden.shape[:-2].numel() != 1:
        raise ValueError("only one packed row per call (micro_batch_size == 1 for packed batches)")
    partition = None
    if loss_mask is not None:
        partition = loss_mask if isinstance(loss_mask, rows.Partition) else rows.partition_positions(loss_mask, bool(shift))
        if partition.T != input_ids.numel() or partition.shift != int(bool(shift)):
            raise ValueError("loss_mask / partition does not match input_ids [..., T] and `shift`")
    return _FusedLMHead.apply(hidden, weight, input_ids, 1.0 / float(temperature), bool(shift), bool(round_bf16), main_grad,
                              partition, bool(forward_active_only), save_logits)


class FusedLMHead(nn.Module):
    """Drop-in replacement MODULE for `model.lm_head` (SURVEY §7 "FSDP integration"): it owns the same
    `weight` Parameter, so FSDP2's `[lm_head, norm]` group hooks (model.py:126-131) keep firing, and its
    `forward(hidden)` returns a `[..., T, 2]` tensor (logprob, entropy) where the reference returns
    logits.  Call `set_context(input_ids, temperature)` right before the model forward; with the micro-batch's `loss_mask` as
    third argument the backward skips the masked (prompt / environment) positions, whose gradient is zero anyway."""

    def __init__(self, weight: nn.Parameter, main_grad: torch.Tensor | None = None, save_logits: bool | None = None):
        super().__init__()
        self.weight = weight
        self.main_grad = main_grad
        self.save_logits = save_logits  # None: recompute unless PRL_SAVE_LOGITS says otherwise (saved_logits_policy)
        self._ctx: tuple | None = None

    @classmethod
    def from_linear(cls, linear: nn.Linear) -> "FusedLMHead":
        if linear.bias is not None:
            raise ValueError("lm_head with bias is not supported (Qwen3/Llama heads are bias-free)")
        return cls(linear.weight)

    def set_context(self, input_ids: torch.Tensor, temperature: float, loss_mask: torch.Tensor | None = None) -> None:
        self._ctx = (input_ids, float(temperature), loss_mask)

    def forward(self, hidden: torch.Tensor) -> torch.Tensor:
        if self._ctx is None:
            raise RuntimeError("FusedLMHead.set_context(input_ids, temperature) must be called before forward")
        ids, temp, mask = self._ctx
        self._ctx = None
        logp, ent = fused_lm_head_logprobs_entropy(hidden, self.weight, ids, temp, main_grad=self.main_grad, loss_mask=mask,
                                                   save_logits=self.save_logits)
        return torch.stack([logp, ent], dim=-1)
