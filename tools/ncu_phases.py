"""Per-phase breakdown of one tile_kernel profile: the SASS between two CTA barriers is one phase.

  python tools/ncu_phases.py gpurun_out/prof_<tag>.ncu-rep [out.json]

For every phase: warp instructions executed, share of the launch, average active threads per instruction,
stall samples (total and the top reasons), shared-memory bank-conflict wavefronts, and the opcode mix.
Reads the ncu source page (`--page source --print-source sass`); needs a report taken with --set full.
"""
import collections
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
phases, cur = [], None


def new_phase():
    return {"inst": 0, "thread_inst": 0, "samples": 0, "stalls": collections.Counter(), "ops": collections.Counter(),
            "static": 0, "smem_excess": 0, "smem_wavefronts": 0}


cur = new_phase()
for r in rows:
    if len(r) != len(hdr) or r[0] == "Address":
        continue
    sass = r[ix["Source"]].strip()
    toks = sass.split()
    if not toks:
        continue
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    n = int(r[ix["Instructions Executed"]] or 0)
    cur["static"] += 1
    cur["inst"] += n
    cur["thread_inst"] += int(r[ix["Thread Instructions Executed"]] or 0)
    cur["samples"] += int(r[ix["# Samples"]] or 0)
    cur["ops"][op] += n
    cur["smem_excess"] += int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
    cur["smem_wavefronts"] += int(r[ix["L1 Wavefronts Shared"]] or 0)
    for s in stall_cols:
        v = int(r[ix[s]] or 0)
        if v:
            cur["stalls"][s] += v
    if op in ("BAR", "EXIT"):
        phases.append(cur)
        cur = new_phase()
if cur["static"]:
    phases.append(cur)
tot = sum(p["inst"] for p in phases)
tots = sum(p["samples"] for p in phases)
res = []
print(f"total warp instructions {tot}, stall samples {tots}")
for i, p in enumerate(phases):
    if p["inst"] == 0:
        continue
    d = {"phase": i, "static_sass": p["static"], "warp_inst": p["inst"], "inst_share": p["inst"] / tot,
         "sample_share": p["samples"] / max(tots, 1), "avg_threads": p["thread_inst"] / max(p["inst"], 1),
         "smem_wavefronts": p["smem_wavefronts"], "smem_excess_wavefronts": p["smem_excess"],
         "top_stalls": {k: v / max(p["samples"], 1) for k, v in p["stalls"].most_common(5)},
         "top_ops": {k: v / p["inst"] for k, v in p["ops"].most_common(10)}}
    res.append(d)
    print(f"phase {i:2d}: static {p['static']:5d}  inst {p['inst'] / 1e6:7.1f}M ({100 * d['inst_share']:4.1f}%)  "
          f"samples {100 * d['sample_share']:4.1f}%  thr/inst {d['avg_threads']:4.1f}  "
          f"smem wf {p['smem_wavefronts'] / 1e6:6.1f}M (+{p['smem_excess'] / 1e6:.1f}M)")
    print("          stalls: " + ", ".join(f"{k[6:]} {100 * v:.0f}%" for k, v in d["top_stalls"].items()))
    print("          ops: " + ", ".join(f"{k} {100 * v:.0f}%" for k, v in d["top_ops"].items()))
if len(sys.argv) > 2:
    json.dump(res, open(sys.argv[2], "w"), indent=1)
