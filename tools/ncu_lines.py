"""Per-source-line instruction counts / stall samples from an ncu report (needs -lineinfo)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr and r[0].isdigit() and len(r) > ie and r[ie].isdigit():
        agg.append((int(r[ie]), int(r[isamp]) if r[isamp].isdigit() else 0, cur_file, int(r[0]), r[1].strip()[:100]))
tot = sum(a[0] for a in agg); tots = sum(a[1] for a in agg)
print("total inst", tot, "samples", tots)
for a in sorted(agg, reverse=True)[:topn]:
    print(f"{100*a[0]/tot:5.2f}% inst {100*a[1]/max(tots,1):5.2f}% smp  {a[2]}:{a[3]:4d}  {a[4]}")
