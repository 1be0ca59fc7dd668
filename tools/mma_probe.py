"""tcgen05.mma SS-mode rate probe (csrc/probe_mma.cu): cycles per instruction and PFLOP/s vs N, accumulators in flight, cta_group."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from centerpose_b200 import _lib

L = _lib.lib()
L.cpb200_probe_mma.restype = ctypes.c_int
L.cpb200_probe_mma.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p]
torch.zeros(1, device="cuda")
st = torch.cuda.current_stream().cuda_stream
iters = 20000
for cg in (1, 2):
    for n in (64, 128, 256):
        for nacc in (1, 2):
            if nacc * n > 512:
                continue
            for stride in (0, 32):
                _lib.check(L.cpb200_probe_mma(n, cg, 2000, nacc, stride, st), "probe")
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); _lib.check(L.cpb200_probe_mma(n, cg, iters, nacc, stride, st), "probe"); e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                ns_per = ms * 1e6 / iters
                flops = 2.0 * 128 * cg * n * 16 * iters * (148 // cg) / (ms * 1e-3)
                print(f"cta_group {cg} N={n:3d} accumulators {nacc} k-stride {stride:2d} B: {ns_per:6.1f} ns / MMA (~{ns_per * 1.965:5.0f} cycles @1.965 GHz), "
                      f"{flops / 1e15:.3f} PFLOP/s chip-wide")
