import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from nersemble_b200 import ops
dev = torch.device("cuda", 0)
P = bench.native_params(bench.synthetic_params(), dev)
o, d, t = bench.synthetic_rays(bench.RAYS, 1000, dev)
ts, te, ri, info = ops.march_fixed(o, d, P.aabb, bench.SAMPLES_PER_RAY, bench.STEP, bench.NEAR)
kw = dict(origins=o, directions=d, ray_times=t, t_starts=ts, t_ends=te, ray_indices=ri)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
full = timeit(lambda: ops.field_forward(P, window_hash=32.0, window_deform=7.0, want=("sigma", "rgb", "offsets"), **kw))
nod = timeit(lambda: ops.field_forward(P, window_hash=32.0, window_deform=None, use_deformation=False, want=("sigma", "rgb"), **kw))
off = timeit(lambda: ops.field_forward(P, window_hash=32.0, window_deform=7.0, want=("offsets",), **kw))
print(f"field kernel ms: full {full:.3f}  no_deform {nod:.3f}  offsets_only {off:.3f}")
