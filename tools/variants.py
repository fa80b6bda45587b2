"""Build tile-shape / occupancy variants of libmd2loss.so into build/variants/ (experiments only)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "digging-into-self-supervised-monocular-depth-estimation_b200", "csrc")
SRCS = [os.path.join(CSRC, f) for f in ("md2_abi.cu", "md2_l1.cu", "md2_metrics.cu", "md2_pipeline.cu", "md2_jitter.cu")]
OUT = os.path.join(ROOT, "build", "variants")
os.makedirs(OUT, exist_ok=True)
VARIANTS = [a.split(",") for a in sys.argv[1:]] or [["32", "16", "256", "2"]]
procs = []
for tw, th, nt, minb in VARIANTS:
    name = f"libmd2loss_{tw}x{th}_{nt}_{minb}.so"
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "-shared", f"-DMD2_TW={tw}", f"-DMD2_TH={th}", f"-DMD2_NT={nt}", f"-DMD2_MINB={minb}", "-Xptxas", "-v",
           *SRCS, "-o", os.path.join(OUT, name)]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    out = p.communicate()[0]
    lines = out.splitlines()
    for i, l in enumerate(lines):
        if "Li2ELb1" in l and "Compiling" in l:
            print(name, lines[i + 1].strip(), lines[i + 2].strip())
    if p.returncode != 0:
        print(name, "FAILED", out[-800:])
