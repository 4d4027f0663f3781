"""BASELINE config 4: novel-view frames rendered with rays sharded across the ranks of one node.
Every rank renders a contiguous slice of each frame's rays with a full parameter replica (no data-path collective);
the per-ray RGB is all-gathered to every rank.  Prints frames/s, samples/s and the max |RGB| difference between the
sharded result and an unsharded render of the same frame on rank 0 (must be 0: the kernels are deterministic).
    python tools/render_frames.py [--height 272 --width 480 --frames 4]
    torchrun --nproc-per-node 8 tools/render_frames.py --height 1088 --width 1920 --frames 24"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.distributed as dist
from test_plugin_cpu import make_model
from nersemble_b200.distributed import gather_rays, shard_bounds
from nersemble_b200.nerfstudio_shim import RayBundle

ap = argparse.ArgumentParser()
ap.add_argument("--height", type=int, default=272); ap.add_argument("--width", type=int, default=480)
ap.add_argument("--frames", type=int, default=4); ap.add_argument("--log2T", type=int, default=19)
ap.add_argument("--chunk", type=int, default=1 << 19, help="eval_num_rays_per_chunk (reference scripts pass n_rays_eval; the loop is host-bound, 180 GB of HBM allow large chunks)")
args = ap.parse_args()
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lrank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lrank); dev = torch.device("cuda", lrank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)                                  # identical replicas on every rank
T = 24
m = make_model(T=T, log2T=args.log2T, eval_num_rays_per_chunk=args.chunk).to(dev).eval()
with torch.no_grad():
    m.field.hash_ensemble.tables.uniform_(-0.5, 0.5)
    m.time_embedding.weight.normal_(0, 0.18); m.time_embedding_deformation.weight.normal_(0, 0.09)
    ax = (torch.arange(128, device=dev).float() + 0.5) / 128
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    m.occupancy_grid.binaries[0] = ((X - 0.5) ** 2 + (Y - 0.5) ** 2 + (Z - 0.5) ** 2) < 0.33 ** 2   # a head-sized blob
H, W = args.height, args.width


def camera_rays(frame):
    ang = torch.tensor(2 * torch.pi * frame / max(args.frames, 1))
    o = torch.tensor([9.0 * torch.sin(ang), 0.0, 9.0 * torch.cos(ang)], device=dev)
    fwd = -o / o.norm(); right = torch.linalg.cross(fwd, torch.tensor([0.0, 1.0, 0.0], device=dev)); right = right / right.norm()
    up = torch.linalg.cross(right, fwd)
    ys, xs = torch.meshgrid(torch.linspace(0.25, -0.25, H, device=dev), torch.linspace(-0.25 * W / H, 0.25 * W / H, W, device=dev), indexing="ij")
    d = fwd[None, None] + xs[..., None] * right + ys[..., None] * up
    d = d / d.norm(dim=-1, keepdim=True)
    return o.expand(H * W, 3).contiguous(), d.reshape(-1, 3).contiguous(), torch.full((H * W, 1), frame / max(T - 1, 1), device=dev)


def render(o, d, t):
    n = o.shape[0]
    outs, samples = [], torch.zeros((), dtype=torch.long, device=dev)      # counted on the device: no extra host sync per chunk
    for i in range(0, n, m.config.eval_num_rays_per_chunk):
        sl = slice(i, i + m.config.eval_num_rays_per_chunk)
        rb = RayBundle(origins=o[sl], directions=d[sl], pixel_area=torch.ones_like(t[sl]),
                       camera_indices=torch.zeros_like(t[sl], dtype=torch.long), times=t[sl])
        out = m.get_outputs(rb)
        outs.append(out["rgb"]); samples += out["num_samples_per_ray"].sum()
    return torch.cat(outs), samples


max_diff, total_samples = 0.0, 0
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    lo0, hi0 = shard_bounds(H * W, rank, world)
    render(*[x[lo0:hi0] for x in camera_rays(0)])      # warm-up: one frame at the real chunk sizes (allocator, lazy module loads)
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); e0.record()
    frames = []
    for f in range(args.frames):
        o, d, t = camera_rays(f % T)
        lo, hi = shard_bounds(H * W, rank, world)
        rgb_local, ns = render(o[lo:hi], d[lo:hi], t[lo:hi])
        ns_dev = ns if f == 0 else ns_dev + ns
        frames.append(gather_rays(rgb_local, H * W))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if rank == 0:   # unsharded reference of the last frame: chunk boundaries differ, pixels must not
        full, _ = render(o, d, t)
        max_diff = (full - frames[-1]).abs().max().item()
total_samples = int(ns_dev)
tot = torch.tensor([float(total_samples), ms], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(tot[:1]); dist.all_reduce(tot[1:], op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"n_gpus": world, "frame": [H, W], "frames": args.frames, "ms_per_frame": tot[1].item() / args.frames,
                      "frames_per_s": args.frames / (tot[1].item() / 1e3), "M_samples_per_s": tot[0].item() / tot[1].item() / 1e3,
                      "samples_per_frame": tot[0].item() / args.frames, "max_abs_rgb_diff_vs_unsharded": max_diff}))
if world > 1:
    dist.destroy_process_group()
