"""Small end-to-end exercise of every libnsb kernel, sized for compute-sanitizer (tools/sanitize.sh):
fixed + occupancy march, visibility mask, fused field forward (inference and training instantiations, ragged last tile,
several tiles per CTA so the mbarrier ring / tile counters wrap), composite fwd/bwd, losses, field / deformation
backward, rank-1 expand + fused Adam, occupancy update, component kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from nersemble_b200 import ops
from nersemble_b200.nerfstudio_shim import RayBundle
from nersemble_b200.optim import FusedFieldsAdam

dev = torch.device("cuda", 0)
R = int(os.environ.get("SAN_RAYS", "160"))
S = bench.synthetic_params(n_timesteps=4, log2T=12)
P = bench.native_params(S, dev)
o, d, t = bench.synthetic_rays(R, 3, dev)
ts, te, ri, info = ops.march_fixed(o, d, P.aabb, 100, bench.STEP, bench.NEAR)          # 16 000 samples = 125 tiles
out = ops.render_packed(P, o, d, t, ts, te, ri, info, window_hash=32.0, window_deform=7.0, training=False)
out2 = ops.render_packed(P, o, d, t, ts[:-37], te[:-37], ri[:-37],
                         torch.stack([info[:, 0], torch.clamp(info[:, 1] - torch.tensor([0] * (R - 1) + [37], device=dev), min=0)], -1),
                         window_hash=1.5, window_deform=3.3, training=False)                 # ragged last tile
x = torch.rand((500, 3), device=dev); codes = torch.randn((500, 32), device=dev) * 0.2
ops.hash_blend_forward(P, x, codes, window_hash=20.25)
ops.field_forward(P, window_hash=32.0, window_deform=7.0, positions=P.aabb[0].to(dev) + x * (P.aabb[1] - P.aabb[0]).to(dev),
                  sample_times=torch.rand(500, device=dev), want=("sigma",))
# plugin model: occupancy march + pre-pass + training step + occupancy update
m = bench.build_model(S, dev).train()
occ = bench.blob_occupancy(seed=5)
m.occupancy_grid.binaries[0] = occ.to(dev); m.occupancy_grid.occs.copy_((occ.flatten().float() * 0.05).to(dev))
g = m.get_param_groups()
opts = [FusedFieldsAdam(g["fields"], lr=5e-3, eps=1e-15), torch.optim.Adam(g["embeddings"] + [p for p in g["deformation_field"] if p.requires_grad], lr=1e-3)]
rb = RayBundle(origins=o, directions=d, pixel_area=torch.ones(R, 1, device=dev),
               camera_indices=torch.zeros(R, 1, dtype=torch.long, device=dev), times=t)
gen = torch.Generator().manual_seed(0)
batch = {"image": torch.rand((R, 3), generator=gen).to(dev), "alpha_map": torch.randint(0, 256, (R, 1), generator=gen).float().to(dev),
         "depth_maps": (torch.rand(R, generator=gen) * 4 + 7).to(dev)}
for it in range(2):
    for op in opts:
        op.zero_grad()
    outs = m.get_outputs(rb)
    loss = sum(m.get_loss_dict(outs, batch).values())
    loss.backward()
    for op in opts:
        op.step()
m.get_training_callbacks(None)[0].run_callback(0)
m.get_training_callbacks(None)[0].run_callback(4096)
m.eval()
with torch.no_grad():
    m.get_outputs(rb)
torch.cuda.synchronize()
print("sanitize_driver: ok, loss", float(loss), "rgb mean", float(out["rgb"].mean()), float(out2["rgb"].mean()))
