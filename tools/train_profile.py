"""torch.profiler table of one training step of bench config 3 / 5 (kernel breakdown).  usage: python tools/train_profile.py [3|5]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench as B
from tools import bench_configs as C
from nersemble_b200.nerfstudio_shim import RayBundle
cfg5 = len(sys.argv) > 1 and sys.argv[1] == "5"
dev = torch.device("cuda", 0)
S = B.synthetic_params()
if cfg5:
    S["aabb"] = torch.tensor(C.SEQ97_AABB)
    m = B.build_model(S, dev, disable_occupancy_grid=True, lambda_dist_loss=0.0).train()
    m.occupancy_grid.binaries[:] = True; m.occupancy_grid.occs.fill_(B.STEP)
else:
    m = B.build_model(S, dev).train()
    occ = B.blob_occupancy(seed=5)
    m.occupancy_grid.binaries[0] = occ.to(dev); m.occupancy_grid.occs.copy_((occ.flatten().float() * 0.05).to(dev))
opts, params = C._optimizers(m)
o, d, t = B.synthetic_rays(B.RAYS, 1000, dev)
rb = RayBundle(origins=o, directions=d, pixel_area=torch.ones(B.RAYS, 1, device=dev), camera_indices=torch.zeros(B.RAYS, 1, dtype=torch.long, device=dev), times=t)
batch = C._train_batch(dev, 7)
def step():
    for op in opts: op.zero_grad(set_to_none=True)
    out = m.get_outputs(rb)
    loss = sum(m.get_loss_dict(out, batch).values())
    loss.backward()
    for op in opts: op.step()
    return out
for _ in range(4): out = step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    out = step(); torch.cuda.synchronize()
print("kept samples", int(out["num_samples_per_ray"].sum()))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
print("==== sorted by CPU time (host-side cost of one step)")
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=45, max_name_column_width=90))
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"10 steps: host enqueue {1e2 * (t1 - t0):.2f} ms/step, wall {1e2 * (t2 - t0):.2f} ms/step")
