"""SASS opcode histogram of the shipped library (evidence that the cubins are sm_100a-native: TMA, packed fp32, ...).

    python tools/sass_histogram.py [profiles/<tag>_sass_histogram.json]
"""
import collections, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "digging-into-self-supervised-monocular-depth-estimation_b200", "libmd2loss.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
per_fn, total, fn = {}, collections.Counter(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        per_fn[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
    if not m or fn is None:
        continue
    ins = m.group(1).split()
    op = ins[1] if ins[0].startswith("@") else ins[0]
    per_fn[fn][op] += 1
    total[op] += 1
interesting = ["UTMALDG.4D", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "FFMA2", "FADD2", "FMUL2", "IDP.4A.U8.S8",
               "IDP.4A.U8.U8", "REDG.E.ADD.F32.FTZ.RN.STRONG.GPU", "RED.E.ADD.F32x4.FTZ.RN.STRONG.GPU", "LDS.128", "LDS.64", "STS.64",
               "LDG.E.CONSTANT", "MUFU.RCP", "BAR.SYNC.DEFER_BLOCKING", "HMMA", "WARPSYNC"]
def family(prefix):
    return {k: v for k, v in total.items() if k.startswith(prefix)}
tile = {f: c for f, c in per_fn.items() if "tile_kernel" in f}
main = next((f for f in tile if "TileILi2ELb1ELi32ELi16ELi320ELi0EEELb0" in f), None)
res = {"library": os.path.relpath(lib, ROOT), "architectures": archs, "functions": len(per_fn), "instructions": sum(total.values()),
       "families": {p: family(p) for p in ("UTMALDG", "SYNCS", "FFMA2", "FADD2", "IDP", "RED", "ATOM", "LDGSTS", "UBLKCP", "HMMA", "UTC")},
       "tile_kernel_S2_training": {"function": main, "instructions": sum(tile[main].values()) if main else None,
                                   "top": dict(tile[main].most_common(25)) if main else None},
       "top_overall": dict(total.most_common(40))}
print(json.dumps({k: res[k] for k in ("architectures", "functions", "instructions", "families")}, indent=1))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
