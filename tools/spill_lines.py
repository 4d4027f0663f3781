"""List local-memory (spill) instructions of one kernel with their source lines, per warp-role region.
usage: python tools/spill_lines.py <mangled-name-substring> [obj]"""
import re, subprocess, sys, tempfile, os, glob
pat = sys.argv[1]
obj = os.path.abspath(sys.argv[2] if len(sys.argv) > 2 else "nersemble_b200/csrc/nsb_field.o")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
cubin = glob.glob(d + "/*.cubin")[0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
sec = None; cur = None; region = None
for line in dis.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', line)
    if m: sec = m.group(1); cur = None; region = 'pre'; continue
    if line.lstrip().startswith('.section'): sec = None; continue
    if sec is None or pat not in sec: continue
    if "USETMAXREG.DEALLOC" in line: region = 'gather'
    elif "USETMAXREG.TRY_ALLOC" in line: region = 'tensor'
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.search(r'\b(STL|LDL)[.\w]*\s+(.*?);', line)
    if m and (len(sys.argv) < 4 or region == sys.argv[3]): print(region, cur, m.group(0))
