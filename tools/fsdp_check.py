"""FSDP2 integration of `FusedLMHead` (SURVEY §7 "FSDP integration"), 2 ranks, NCCL:

  * a tiny random-init HF Qwen3 (fp32 master weights) is wrapped exactly like the reference's `setup_fsdp`
    (src/prime_rl/trainer/model.py:102-135): every block, embed_tokens, the group [lm_head, norm] with
    reshard_after_forward=False, then the root; MixedPrecisionPolicy(param_dtype=bf16, reduce_dtype=fp32);
  * `model.lm_head` is swapped for `FusedLMHead` BEFORE sharding, so the group's forward/backward hooks see the module;
  * each rank runs ONE different packed micro-batch through `model(...)` -> (logprob, entropy) -> compute_loss -> backward;
  * the reduce-scattered gradients are compared with the average of the two ranks' gradients computed by the plain torch
    chain (materialised logits) on an unsharded bf16 copy of the same model.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/fsdp_check.py
"""
import os, sys
import torch
import torch.distributed as dist
from torch.distributed.fsdp import MixedPrecisionPolicy, fully_shard

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prime_rl_cse_b200 as P
import transformers


def make_model():
    cfg = transformers.Qwen3Config(vocab_size=2048, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=2, head_dim=32, max_position_embeddings=512, tie_word_embeddings=False)
    torch.manual_seed(0)
    return transformers.Qwen3ForCausalLM(cfg)


def micro_batch(rank, V, T=192):
    g = torch.Generator().manual_seed(100 + rank)
    lens = [T // 2, T - T // 2]
    mb = dict(input_ids=torch.randint(0, V, (1, T), generator=g), position_ids=torch.cat([torch.arange(n) for n in lens]).unsqueeze(0),
              advantages=torch.cat([torch.full((n,), float(torch.randn((), generator=g))) for n in lens]).unsqueeze(0),
              logprobs=-torch.rand(1, T, generator=g) * 2, loss_mask=(torch.rand(1, T, generator=g) < 0.8), temperature=0.8)
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in mb.items()}, lens


def torch_chain_grads(model_bf16, mb, lens):
    """reference chain (train.py:285-310) with plain torch ops on an unsharded bf16 model"""
    model_bf16.zero_grad()
    logits = model_bf16(input_ids=mb["input_ids"], position_ids=mb["position_ids"]).logits.float().contiguous()
    shifted = torch.cat([torch.zeros_like(logits[:, :1]), logits[:, :-1]], dim=1) / mb["temperature"]
    lp = (torch.gather(shifted, -1, mb["input_ids"].unsqueeze(-1)).squeeze(-1) - torch.logsumexp(shifted, -1))[0]
    mask, old, adv = mb["loss_mask"][0], mb["logprobs"][0], mb["advantages"][0]
    r = torch.exp(lp - old)
    loss = (-(torch.clamp(r, max=8.0)) * adv)[mask].sum() / max(int(mask.sum()), 1)
    loss.backward()
    return {n: p.grad.float().clone() for n, p in model_bf16.named_parameters()}, loss.detach()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    V = 2048
    # expected: average over ranks of the plain-chain gradients (every rank can compute both micro-batches)
    ref_model = make_model().cuda().bfloat16()
    exp, ref_losses = None, []
    for r in range(world):
        mb, lens = micro_batch(r, V)
        g, l = torch_chain_grads(ref_model, mb, lens)
        ref_losses.append(l)
        exp = g if exp is None else {k: exp[k] + g[k] for k in exp}
    exp = {k: v / world for k, v in exp.items()}
    del ref_model

    model = make_model().cuda()  # fp32 master weights, bf16 compute via the mixed-precision policy
    model.lm_head = P.FusedLMHead.from_linear(model.lm_head)
    mp = MixedPrecisionPolicy(param_dtype=torch.bfloat16, reduce_dtype=torch.float32)
    for blk in model.model.layers:
        fully_shard(blk, mp_policy=mp)
    fully_shard(model.model.embed_tokens, mp_policy=mp)
    fully_shard([model.lm_head, model.model.norm], mp_policy=mp, reshard_after_forward=False)
    fully_shard(model, mp_policy=mp)

    mb, lens = micro_batch(rank, V)
    model.lm_head.set_context(mb["input_ids"], mb["temperature"])
    out = model(input_ids=mb["input_ids"], position_ids=mb["position_ids"]).logits
    assert out.shape[-1] == 2, out.shape
    logprobs = out[..., 0]
    lengths = P.get_response_lengths(mb["position_ids"])
    assert lengths == lens, (lengths, lens)
    loss, _ = P.compute_loss(logprobs.squeeze().split(lengths), mb["logprobs"].squeeze().split(lengths),
                             mb["advantages"].squeeze().split(lengths), mb["loss_mask"].squeeze().split(lengths), P.LossConfig(),
                             int(mb["loss_mask"].sum()))
    loss.backward()
    torch.cuda.synchronize()

    ok = abs(loss.item() - ref_losses[rank].item()) <= 2e-2 * abs(ref_losses[rank].item()) + 1e-5
    worst = ("", 0.0)
    for n, p in model.named_parameters():
        assert p.grad is not None, f"no gradient for {n} (FSDP hooks did not fire?)"
        g = p.grad.full_tensor().float()
        e = exp[n]
        rel = ((g - e).norm() / e.norm().clamp_min(1e-12)).item()
        if rel > worst[1]:
            worst = (n, rel)
    ok = ok and worst[1] < 5e-2  # bf16 transformer body on both sides; lm_head itself is ~1e-3
    head_rel = ((dict(model.named_parameters())["lm_head.weight"].grad.full_tensor().float() - exp["lm_head.weight"]).norm()
                / exp["lm_head.weight"].norm()).item()
    print(f"rank {rank}: loss {loss.item():.6f} (torch chain {ref_losses[rank].item():.6f}); lm_head.weight grad rel err {head_rel:.2e}; "
          f"worst param {worst[0]} rel err {worst[1]:.2e} -> {'OK' if ok and head_rel < 1e-2 else 'MISMATCH'}", flush=True)
    flag = torch.tensor([0 if (ok and head_rel < 1e-2) else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
