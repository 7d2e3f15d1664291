import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prime_rl_cse_b200 import lm_head_forward_raw
T, H, V = int(os.environ.get("T", 16384)), int(os.environ.get("H", 4096)), int(os.environ.get("V", 151936))
torch.manual_seed(0)
hidden = torch.randn(T, H, device="cuda").bfloat16(); W = (torch.randn(V, H, device="cuda") * 0.02).bfloat16()
ids = torch.randint(0, V, (T,), device="cuda")
for _ in range(2):
    lm_head_forward_raw(hidden, W, ids, 1.0)
torch.cuda.synchronize()
