"""Static SASS instruction count per phase (between BAR.SYNCs) of a tile_kernel instantiation."""
import collections, re, subprocess, sys
lib = "digging-into-self-supervised-monocular-depth-estimation_b200/libmd2loss.so"
pat = sys.argv[1] if len(sys.argv) > 1 else "Li2ELb1"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
on = False; n = 0; ops = collections.Counter(); segs = []
for line in out.splitlines():
    if "Function :" in line:
        if on: segs.append((n, ops.most_common(6)))
        on = ("tile_kernel" in line and pat in line); n = 0; ops = collections.Counter(); 
        if on: print(line.strip()[:120])
        continue
    if not on: continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
    if not m: continue
    ins = m.group(1).split()
    op = (ins[1] if ins[0].startswith("@") else ins[0]).split(".")[0]
    n += 1; ops[op] += 1
    if op in ("BAR", "EXIT"):
        segs.append((n, ops.most_common(7))); n = 0; ops = collections.Counter()
for s in segs:
    if s[0] > 20: print(s)
