"""Time one full training step (fwd + losses + bwd + Adam [+ gradient all-reduce]) of the plugin model at the
BASELINE config-3/5 shape: 4096 rays, dense 256 samples/ray (2^20 samples), full-size tables, T = 24.
    python tools/train_step_bench.py [--steps 5]            (1 GPU)
    torchrun --nproc-per-node N tools/train_step_bench.py   (N GPUs: + one NCCL all-reduce of all gradients)"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.distributed as dist
import bench
from test_plugin_cpu import make_model
from nersemble_b200.nerfstudio_shim import RayBundle
from nersemble_b200.distributed import allreduce_gradients

ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=5); ap.add_argument("--profile", action="store_true"); ap.add_argument("--cprofile", action="store_true", help="host-side cProfile of 5 more steps"); ap.add_argument("--all-losses", action="store_true", help="config 3: all six losses with the script's default lambdas (synthetic depth maps)"); ap.add_argument("--torch-adam", action="store_true", help="dense table gradient + torch.optim.Adam (the reference's optimiser path) instead of FusedFieldsAdam"); args = ap.parse_args()
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr_ = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr_); dev = torch.device("cuda", lr_)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
m = (make_model(T=24, log2T=19) if args.all_losses else
     make_model(T=24, log2T=19, lambda_near_loss=0, lambda_empty_loss=0, lambda_depth_loss=0)).to(dev).train()
with torch.no_grad():   # trained-like scale so that densities are non-trivial
    m.field.hash_ensemble.tables.uniform_(-0.5, 0.5)
    m.time_embedding.weight.normal_(0, 0.18); m.time_embedding_deformation.weight.normal_(0, 0.09)
m.occupancy_grid.binaries[:] = True           # dense march (config 5: --disable_occupancy_grid-like sample count)
m.config.far_plane = 1e3
o, d, t = bench.synthetic_rays(4096, 1000 + rank, dev)
rb = RayBundle(origins=o, directions=d, pixel_area=torch.ones(4096, 1, device=dev),
               camera_indices=torch.zeros(4096, 1, dtype=torch.long, device=dev), times=t)
batch = {"image": torch.rand(4096, 3, device=dev), "alpha_map": torch.randint(0, 256, (4096, 1), device=dev).float()}
if args.all_losses:
    batch["depth_maps"] = (torch.rand(4096, device=dev) * 4 + 7) * (torch.rand(4096, device=dev) > 0.2)   # 0 = no depth
groups = m.get_param_groups()
from nersemble_b200.optim import FusedFieldsAdam
class _Opts:    # one optimiser per group, like nerfstudio's Optimizers
    def __init__(self):
        F = torch.optim.Adam if args.torch_adam else FusedFieldsAdam
        self.o = [F(groups["fields"], lr=5e-3, eps=1e-15), torch.optim.Adam(groups["embeddings"], lr=5e-3, eps=1e-15),
                  torch.optim.Adam([p for p in groups["deformation_field"] if p.requires_grad], lr=1e-3, eps=1e-15)]
    def zero_grad(self, set_to_none=True): [o.zero_grad(set_to_none) for o in self.o]
    def step(self): [o.step() for o in self.o]
opt = _Opts()
HE = [m.field.hash_ensemble]
m.sampler.eval()
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
rows = []
cpu_ms = []
for step in range(args.steps + 2):
    torch.cuda.synchronize(); t_cpu0 = time.perf_counter()
    e0 = ev(); opt.zero_grad(set_to_none=True)
    out = m.get_outputs(rb); e1 = ev()
    loss = sum(m.get_loss_dict(out, batch).values()); e2 = ev()
    loss.backward(); e3 = ev()
    allreduce_gradients([p for gr in groups.values() for p in gr], hash_ensembles=HE); e4 = ev()
    opt.step(); e5 = ev()
    t_cpu1 = time.perf_counter()            # host time to ENQUEUE the step (no sync inside = launch-bound floor)
    torch.cuda.synchronize()
    if step >= 2: cpu_ms.append((t_cpu1 - t_cpu0) * 1e3)
    if step >= 2:
        rows.append([e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e3.elapsed_time(e4), e4.elapsed_time(e5)])
if args.profile and rank == 0:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        opt.zero_grad(set_to_none=True)
        out = m.get_outputs(rb)
        loss = sum(m.get_loss_dict(out, batch).values())
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
if args.cprofile and rank == 0:
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(5):
        opt.zero_grad(set_to_none=True)
        out = m.get_outputs(rb)
        loss = sum(m.get_loss_dict(out, batch).values())
        loss.backward()
        opt.step()
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
avg = [sum(r[i] for r in rows) / len(rows) for i in range(5)]
n_samples = int(out["num_samples_per_ray"].sum().item())
if rank == 0:
    print(json.dumps({"losses": sorted(m.get_loss_dict(out, batch).keys()), "optimizer": "torch.optim.Adam" if args.torch_adam else "FusedFieldsAdam", "n_gpus": world, "samples_per_gpu": n_samples, "loss": loss.item(),
                      "ms": dict(forward=avg[0], losses=avg[1], backward=avg[2], allreduce=avg[3], adam=avg[4], total=sum(avg)),
                      "host_enqueue_ms": sum(cpu_ms) / len(cpu_ms), "it_per_s": 1000.0 / sum(avg), "M_samples_per_s": n_samples * world / sum(avg) / 1e3}))
if world > 1:
    dist.destroy_process_group()
