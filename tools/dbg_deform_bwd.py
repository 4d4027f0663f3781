import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import native_from_oracle
from oracle import pipeline as pl
from nersemble_b200 import ops
stage = sys.argv[1]
P = pl.random_params(n_timesteps=4, log2_hashmap_size=12, table_scale=0.5, time_std_scale=100.0, deform_last_scale=0.05)
NP = native_from_oracle(P, "cuda:0")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 333
g = torch.Generator().manual_seed(9)
pos = (P.aabb[0] + (torch.rand((n, 3), generator=g) * 0.9 + 0.05) * (P.aabb[1] - P.aabb[0])).cuda()
times = (torch.sort(torch.randint(0, 4, (n,), generator=g))[0].float() / 3).cuda()
kw = dict(positions=pos, sample_times=times)
print("fwd...", flush=True)
saved = ops.field_forward(NP, window_hash=None, window_deform=5.5, use_deformation=True, want=("offsets", "deform_acts"), **kw)
torch.cuda.synchronize(); print("fwd ok", saved["offsets"].abs().mean().item(), flush=True)
if stage == "bwd":
    d_xs = torch.randn((n, 3), device="cuda")
    gr = ops.deform_backward(NP, saved, d_xs, window_deform=5.5, loss_scale=64.0, **kw)
    torch.cuda.synchronize(); print("bwd ok", gr["d_r_w"].abs().mean().item(), gr["d_stem_w"][0].abs().mean().item(), flush=True)
