"""Dump the SASS of one phase (index as printed by ncu_phases.py) of a tile_kernel profile with per-instruction
execution counts (normalised by `norm` if given) and stall samples.  python tools/ncu_sass_dump.py rep phase [norm]"""
import csv, subprocess, sys
rep, ph = sys.argv[1], int(sys.argv[2])
norm = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {n: i for i, n in enumerate(hdr)}
cur = 0
for r in rows:
    if len(r) != len(hdr) or r[0] == "Address":
        continue
    sass = r[ix["Source"]].strip()
    toks = sass.split()
    if not toks:
        continue
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    if cur == ph:
        n = int(r[ix["Instructions Executed"]] or 0)
        print(f"{n / norm:9.2f} {int(r[ix['# Samples']] or 0):5d} {float(r[ix['Avg. Threads Executed']] or 0):5.1f}  {sass[:110]}")
    if op in ("BAR", "EXIT"):
        cur += 1
