"""tcgen05.mma throughput of this GPU (own micro-benchmark kernel, csrc/tc_peak.cu): kind::f16 and kind::tf32, A operand in shared
memory (SS) or tensor memory (TS), N = 256 down to 16.  Writes gpurun_out/tcgen05_peaks.json (copied to profiles/ after a run):
    python tools/tc_peak.py"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ptgnn_b200 import _native as N  # noqa: E402

lib = N.lib()
fn = lib.ptgnn_b200_debug_tcgen05_peak
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]
torch.cuda.init()
torch.zeros(1, device="cuda")
out = {"gpu": torch.cuda.get_device_name(0), "unit": "TFLOP/s", "M": 128, "results": []}
for kind, kname, kk in ((0, "f16", 16), (2, "tf32", 8)):
    for ts in (0, 1):
        for n in (256, 128, 64, 48, 32, 16):
            iters = 40000 * 256 // n // (1 if kind == 0 else 1)
            best = 0.0
            for _ in range(3):
                ms, grid = ctypes.c_float(0), ctypes.c_int(0)
                rc = fn(kind, ts, n, iters, ctypes.byref(ms), ctypes.byref(grid))
                assert rc == 0, lib.ptgnn_b200_last_error()
                flops = 2.0 * 128 * n * kk * 4 * iters * grid.value
                best = max(best, flops / (ms.value * 1e-3) / 1e12)
            cyc = 2.0 * 128 * n * kk / (best * 1e12 / grid.value / 1.965e9)     # issue interval per MMA in cycles at 1965 MHz
            out["results"].append({"kind": kname, "a_operand": "tmem" if ts else "smem", "N": n, "tflops": round(best, 1),
                                   "ms": round(ms.value, 3), "cycles_per_mma_at_1965MHz": round(cyc, 1)})
            print(f"kind::{kname:4s} A in {'TMEM' if ts else 'SMEM'}  N={n:3d}: {best:7.1f} TFLOP/s   ({out['results'][-1]['cycles_per_mma_at_1965MHz']} cycles per MMA)")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "tcgen05_peaks.json"), "w") as f:
    json.dump(out, f, indent=1)
