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

# ---- the halo-window A operand of the 3x3 kernel and the split-operand K step (csrc/net_tc3.cu) --------------
L.cpb200_probe_mma_ex.restype = ctypes.c_int
L.cpb200_probe_mma_ex.argtypes = [ctypes.c_int] * 8 + [ctypes.c_void_p]
for n, alt in ((256, 0), (128, 0), (256, 1)):
    for sbo, shift in ((0, 0), (1280, 0), (1280, 128), (1280, 1280 + 256), (1152, 128)):
        args = (n, 1, iters, 2 if n <= 128 else 1, 32, sbo, shift, alt, st)
        _lib.check(L.cpb200_probe_mma_ex(n, 1, 2000, args[3], 32, sbo, shift, alt, st), "probe")
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.check(L.cpb200_probe_mma_ex(*args), "probe"); e1.record(); torch.cuda.synchronize()
        ns_per = e0.elapsed_time(e1) * 1e6 / iters
        what = "A_hi x [W_hi|W_lo] (N) + A_lo x W_hi (N/2) per step" if alt else "one MMA per step"
        print(f"N={n:3d} A window: group stride {sbo or 1024:4d} B, start +{shift:4d} B, {what}: {ns_per:6.1f} ns / step")

# ---- the conv kernels' issue loop: full-barrier wait, K steps, tcgen05.commit per stage (stand-in producer) -----
L.cpb200_probe_mma_pipe.restype = ctypes.c_int
L.cpb200_probe_mma_pipe.argtypes = [ctypes.c_int] * 7 + [ctypes.c_void_p]
stages_n = 4000
MODES = {0: "wait + fence + commit", 1: "no wait", 2: "arrive instead of commit", 3: "no wait, arrive instead of commit",
         4: "early wait", 8: "no fence", 12: "early wait, no fence"}
for n, alt, nacc in ((256, 1, 2), (128, 0, 2)):
    for ksteps in (4, 8):
        for mode, label in MODES.items():
            args = (n, stages_n, nacc, alt, 4, ksteps, mode, st)
            _lib.check(L.cpb200_probe_mma_pipe(n, 500, nacc, alt, 4, ksteps, mode, st), "probe"); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); _lib.check(L.cpb200_probe_mma_pipe(*args), "probe"); e1.record(); torch.cuda.synchronize()
            ns_stage = e0.elapsed_time(e1) * 1e6 / stages_n
            print(f"pipe N={n:3d} {'split K step' if alt else 'one MMA/step '} ring 4 x {ksteps:2d} K steps, {label:36s}: "
                  f"{ns_stage:7.1f} ns / stage = {ns_stage / ksteps:6.1f} ns / K step")

# ---- pipe2: unrolled K steps, early poll of the next stage, single-thread loop vs whole-warp loop + per-instruction elect
L.cpb200_probe_mma_pipe2.restype = ctypes.c_int
L.cpb200_probe_mma_pipe2.argtypes = [ctypes.c_int] * 7 + [ctypes.c_void_p]
for n, alt, nacc in ((256, 1, 2), (256, 0, 2), (128, 0, 2)):
    for ksteps in (4, 8):
        for uniform in (0, 1):
            args = (n, stages_n, nacc, alt, 4, ksteps, uniform, st)
            _lib.check(L.cpb200_probe_mma_pipe2(n, 500, nacc, alt, 4, ksteps, uniform, st), "probe"); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); _lib.check(L.cpb200_probe_mma_pipe2(*args), "probe"); e1.record(); torch.cuda.synchronize()
            ns_stage = e0.elapsed_time(e1) * 1e6 / stages_n
            print(f"pipe2 N={n:3d} {'split K step' if alt else 'one MMA/step '} ring 4 x {ksteps} K steps, "
                  f"{'whole warp + elect per instruction' if uniform else 'one elected thread            '}: "
                  f"{ns_stage:7.1f} ns / stage = {ns_stage / ksteps:6.1f} ns / K step")
