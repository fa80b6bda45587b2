"""Build libmd2loss.so with -Xptxas -v and print one line per tile_kernel instantiation (registers, spills)."""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import md2_b200.build as b
srcs = [os.path.join(b.CSRC, f) for f in ("md2_abi.cu", "md2_l1.cu", "md2_metrics.cu", "md2_pipeline.cu", "md2_jitter.cu", "md2_pad.cu", "md2_pool.cu")]
extra = sys.argv[1:]
r = subprocess.run([b.NVCC] + b.NVCC_FLAGS + ["-Xptxas", "-v"] + extra + srcs + ["-o", b.LIB], capture_output=True, text=True)
if r.returncode != 0:
    print(r.stderr[-3000:]); sys.exit(1)
lines = (r.stdout + r.stderr).splitlines()
seen = set()
for i, l in enumerate(lines):
    m = re.search(r"tile_kernelINS_4TileILi(\d)ELb(\d)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d)EEELb(\d)", l)
    if m and "Compiling" in l and m.group(6) == "0" and m.group(7) == "0":
        key = m.groups()
        if key in seen: continue
        seen.add(key)
        sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", lines[i + 2])
        rg = re.search(r"Used (\d+) registers", lines[i + 3])
        print(f"S={m.group(1)} bwd={m.group(2)} {m.group(3)}x{m.group(4)} nt={m.group(5)}: regs {rg.group(1) if rg else '?'} "
              f"stack {sp.group(1) if sp else '?'} spill st/ld {sp.group(2) if sp else '?'}/{sp.group(3) if sp else '?'}")
