"""Build experiment variants of libmd2loss.so into build/variants/ (measurement only; never shipped).

  python tools/variants.py name1:-DFLAG_A,-DFLAG_B=3 name2:-DMD2_NT=288 ...

Each spec is `name:comma-separated nvcc flags`; the libraries are built in parallel and
tools/bench_variants.py times every one of them on the benchmark configuration."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "digging-into-self-supervised-monocular-depth-estimation_b200", "csrc")
SRCS = [os.path.join(CSRC, f) for f in ("md2_abi.cu", "md2_l1.cu", "md2_metrics.cu", "md2_pipeline.cu", "md2_jitter.cu", "md2_pad.cu", "md2_pool.cu")]
OUT = os.path.join(ROOT, "build", "variants")
os.makedirs(OUT, exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    flags = [f for f in flags.split(",") if f]
    lib = os.path.join(OUT, f"libmd2loss_{name}.so")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "-shared", "-ldl", "-Xptxas", "-v", *flags, *SRCS, "-o", lib]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    out = p.communicate()[0]
    lines = out.splitlines()
    for i, l in enumerate(lines):
        if re.search(r"TileILi2ELb1ELi32ELi\d+ELi\d+ELi0EEELb0", l) and "Compiling" in l:
            print(name, lines[i + 2].strip(), "|", lines[i + 3].strip())
    if p.returncode != 0:
        print(name, "FAILED", out[-1500:])
