#!/bin/sh
# Memory-safety check of the tile code without a GPU: the host emulation (tests/host_emu) built with
# AddressSanitizer, run over partial tiles and S = 1..4.  (compute-sanitizer is closed on the GPU pool.)
set -e
cd "$(dirname "$0")/.."
g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -std=c++17 -fPIC -shared \
    -I"digging-into-self-supervised-monocular-depth-estimation_b200/csrc" tests/host_emu/md2_emu.cpp -o /tmp/libmd2emu_asan.so
cat > /tmp/asan_run.py <<'PY'
import sys, ctypes
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from host_emu import emu
emu._lib = ctypes.CDLL('/tmp/libmd2emu_asan.so')
for n in ("md2_emu_forward", "md2_emu_forward_backward", "md2_emu_backward", "md2_emu_debug_warp"):
    getattr(emu._lib, n).restype = ctypes.c_int
from test_kernel_logic_emu import synth_args
for B, H, W, f, am in ((2, 40, 72, [0, -1, 1], True), (1, 32, 64, [0, 1], False), (1, 48, 64, [0, -1, 1, 's', 2], True),
                       (1, 24, 40, [0, -1, 1], True)):
    ns = 4 if H % 8 == 0 else 3
    a = synth_args(B, H, W, f, am, 'smooth', 3, num_scales=ns)
    o = emu.forward_backward(a); emu.forward(a); emu.backward(a, o['argmin']); emu.debug_warp(a, 1, 0)
    print(B, H, W, len(f) - 1, float(o['loss']))
print('asan run clean')
PY
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python /tmp/asan_run.py
