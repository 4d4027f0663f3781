"""Per-role spill report for field_kernel_ws / render_kernel_ws: counts STL/LDL in the gather-warp region (between
USETMAXREG.DEALLOC and USETMAXREG.TRY_ALLOC in SASS order) and in the tensor-warp region.
usage: python tools/spill_report.py [nersemble_b200/csrc/nsb_field.o] [name-filter]"""
import re, subprocess, sys
obj = sys.argv[1] if len(sys.argv) > 1 else "nersemble_b200/csrc/nsb_field.o"
flt = sys.argv[2] if len(sys.argv) > 2 else "kernel_ws"
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, rows = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); rows[cur] = {"region": "pre", "pre": 0, "gather": 0, "tensor": 0}; continue
    if cur is None: continue
    r = rows[cur]
    if "USETMAXREG.DEALLOC" in line: r["region"] = "gather"
    elif "USETMAXREG.TRY_ALLOC" in line: r["region"] = "tensor"
    if re.search(r"\b(STL|LDL)\b|\bSTL\.|\bLDL\.", line): r[r["region"]] += 1
for k, r in rows.items():
    if flt in k:
        print(f"{k[7:60]:54s} pre {r['pre']:3d}  gather {r['gather']:3d}  tensor {r['tensor']:3d}")
