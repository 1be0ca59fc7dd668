#!/usr/bin/env python
"""Headline benchmark: images/sec of the centerpose inference hot path (DLA-34 backbone + six
heads + fused multi_pose_decode) on synthetic 512x512 batches, B=32 per GPU (BASELINE.json
configs[1]), weak scaling over N GPUs with an NCCL all-gather of the (B,100,56) detections.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  See the task contract for the keys; extra keys:
  roofline      network-level tensor roofline of the conv/DCN/head op program (the dominant
                kernels), measured live with CUDA events
  decode        the decode kernel's HBM numbers (algorithmic bytes / CUDA-event time)
  cpu_baseline  the oracle port (torch-CPU restatement of the reference) on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

ARCH = "dla_34"
IMG = 512
B_PER_GPU = 32
K_DET = 100
GFLOP_PER_IMG = 80.48            # SURVEY.md §8d (2*MAC of conv+deconv+DCN+heads, DLA-34 @512)
ARCH_GFLOP = {"dla_34": 80.48, "res_50": 86.85, "hrnet": 85.27, "mobilenetv3": 15.59}
ARCH_BATCH = {"dla_34": 32, "res_50": 16, "hrnet": 16, "mobilenetv3": 64}   # BASELINE.json configs[1..4] per GPU
# dram__bytes_read+write summed over the launches of one step (ncu, tools/traffic_list.sh;
# profiles/r01_traffic_per_kernel_v20.txt, profiles/r01_traffic_res50_b16_v1.txt)
NCU_TRAFFIC_BYTES = {("dla_34", 32): 9.359e9, ("res_50", 16): 4.659e9}
DECODE_BYTES_PER_IMG = 1230848
METRIC = "images/sec 512x512 DLA-34"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(device, precision):
    from centerpose_b200.config import default_cfg
    from centerpose_b200.model import create_model
    from oracle.init_recipe import conditioned_state_dict
    cfg = default_cfg(ARCH)
    cfg.B200.PRECISION = precision
    model = create_model(cfg.MODEL.NAME, cfg.MODEL.HEAD_CONV, cfg)
    sd = conditioned_state_dict(model.state_dict(), 317)     # random-init weights, conditioned
    model.load_state_dict(sd)
    return model.to(device).set_precision(precision), sd


def cpu_reference_step(sd, x):
    """The reference's CPU path restated by the oracle: forward -> sigmoid -> multi_pose_decode
    (lib/models/model.py:57-59, lib/detectors/multi_pose.py:32-55, lib/models/decode.py:235-308)."""
    from oracle import decode_ref, dla_ref
    hm, wh, hps, reg, hm_hp, hp_off = dla_ref.forward(sd, x, arch=ARCH)
    hm = hm.sigmoid_(); hm_hp = hm_hp.sigmoid_()
    return decode_ref.multi_pose_decode(hm.numpy(), wh.numpy(), hps.numpy(), reg.numpy(), hm_hp.numpy(),
                                        hp_off.numpy(), K=K_DET)


def cpu_threads():
    """Threads for the CPU arm: all host cores up to 32 — beyond that the reference's small convs
    and the per-tap DCN gathers oversubscribe (measured: 128 threads ran 10x slower than 8)."""
    return max(1, min(os.cpu_count() or 1, 32))


def time_cpu_baseline(sd, n_img, iters):
    from oracle.init_recipe import synth_images
    torch.set_num_threads(cpu_threads())
    x = synth_images(n_img, IMG, IMG, 317)
    cpu_reference_step(sd, x[:1])           # warm-up
    t0 = time.perf_counter()
    for _ in range(iters):
        cpu_reference_step(sd, x)
    dt = time.perf_counter() - t0
    return n_img * iters / dt, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port — the
    reference is Python and cannot travel to the GPU box), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from centerpose_b200.config import default_cfg
    from centerpose_b200.model import create_model
    from oracle.init_recipe import conditioned_state_dict, synth_images
    cfg = default_cfg(ARCH)
    sd = conditioned_state_dict(create_model(cfg.MODEL.NAME, cfg.MODEL.HEAD_CONV, cfg).state_dict(), 317)
    torch.set_num_threads(cpu_threads())
    n_img = 1
    x = synth_images(n_img, IMG, IMG, 317)
    for _ in range(args.warmup):
        cpu_reference_step(sd, x[:1])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(sd, x)
    dt = time.perf_counter() - t0
    val = n_img * args.steps / dt
    sample = f"{n_img} of the {B_PER_GPU} images of a step, per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "images/sec", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{ARCH} {IMG}x{IMG} forward + multi_pose_decode, CPU, {sample}"},
        "cpu_baseline": {"value": val, "unit": "images/sec", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CPB200_PRECISION", "bf16"), choices=["fp16x2", "bf16x2", "bf16", "fp32"])
    ap.add_argument("--arch", default="dla_34", choices=sorted(ARCH_GFLOP), help="dla_34 = the headline config")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    global ARCH, GFLOP_PER_IMG, B_PER_GPU, METRIC
    ARCH = args.arch; GFLOP_PER_IMG = ARCH_GFLOP[ARCH]; B_PER_GPU = args.batch or ARCH_BATCH[ARCH]
    args.batch = B_PER_GPU
    METRIC = "images/sec 512x512 " + {"dla_34": "DLA-34", "res_50": "ResNet-50", "hrnet": "HRNet-W32", "mobilenetv3": "MobileNetV3"}[ARCH]
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from centerpose_b200 import _lib, multi_pose_decode
    from oracle.init_recipe import synth_images
    model, sd = build_model(dev, args.precision)
    B = args.batch
    NROT = 4        # rotate input batches: 4 x 100 MB > the 126 MB L2
    host = [synth_images(B, IMG, IMG, 317 + 1000 * rank + i).pin_memory() for i in range(NROT)]
    resident = [h.to(dev) for h in host]
    dets_all = torch.empty((world * B, K_DET, 56), dtype=torch.float32, device=dev) if world > 1 else None
    host_dets = torch.empty((B, K_DET, 56), dtype=torch.float32).pin_memory()

    def step(x):
        hm, wh, hps, reg, hm_hp, hp_off = model(x)
        dets = multi_pose_decode(hm, wh, hps, reg=reg, hm_hp=hm_hp, hp_offset=hp_off, K=K_DET, apply_sigmoid=True)
        if world > 1:
            dist.all_gather_into_tensor(dets_all, dets)
            return dets_all
        return dets

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput ------------------------------------------------------
    for i in range(args.warmup):
        step(resident[i % NROT])
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = _lib.launch_count()
    ms = timed(lambda i: step(resident[i % NROT]), args.steps)
    launches = _lib.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with host buffers --------------------------------
    # Every step copies its own batch from pinned host memory (H2D, on a copy stream so that the copy of
    # step i+1 overlaps the kernels of step i) and reads its detections back (D2H) — both inside the timed region.
    copy_stream = torch.cuda.Stream(dev)
    dev_in = [torch.empty_like(resident[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def stage_in(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])            # the step that last read this slot is done
            dev_in[slot].copy_(host[i % NROT], non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_run(steps):
        for ev in consumed:
            ev.record(torch.cuda.current_stream(dev))
        stage_in(0)
        for i in range(steps):
            slot = i % 2
            if i + 1 < steps:
                stage_in(i + 1)
            torch.cuda.current_stream(dev).wait_event(ready[slot])
            d = step(dev_in[slot])
            consumed[slot].record(torch.cuda.current_stream(dev))
            host_dets.copy_(d[:B] if world > 1 else d, non_blocking=True)

    e2e_run(2)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e = float(tt.item())
    e2e_val = world * B * args.steps / (ms_e2e * 1e-3)

    # ---- network-only and decode-only kernel timings (roofline) -------------------------------
    model(resident[0])
    ms_net = timed(lambda i: model(resident[i % NROT]), args.steps) / args.steps
    outs = model(resident[0])
    ms_dec = timed(lambda i: multi_pose_decode(outs[0], outs[1], outs[2], reg=outs[3], hm_hp=outs[4],
                                               hp_offset=outs[5], K=K_DET, apply_sigmoid=True), args.steps) / args.steps
    peaks = measured_peaks()
    tf = GFLOP_PER_IMG * B / (ms_net * 1e-3) / 1e3
    roof = {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
            "frac": tf / peaks["bf16_tflops_sustained"], "traffic": NCU_TRAFFIC_BYTES.get((ARCH, B)),
            "kernel": "conv/DCN/head op program (all launches of one forward)", "ms_per_launch_set": ms_net,
            "peak_source": peaks["source"] + " sustained cuBLAS bf16"}
    dec_gbs = DECODE_BYTES_PER_IMG * B / (ms_dec * 1e-3) / 1e9
    decode = {"bound": "hbm", "achieved": dec_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
              "frac": dec_gbs / peaks["hbm_gbs"], "ms": ms_dec, "note": "B=%d maps fit in L2; see profiles/ for B=256/1024" % B}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, dt = time_cpu_baseline(sd, 1, 2)
        cpu = {"value": v, "unit": "images/sec", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "2 images (2 iterations x 1) of the same 512x512 workload, %.1f s" % dt}
    line = {
        "metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": f"{ARCH} {IMG}x{IMG} batch={B}/GPU forward + fused multi_pose_decode (K={K_DET})"
                               + (" + NCCL all-gather of detections" if world > 1 else ""),
                   "global_batch": world * B, "weights": "random-init, conditioned seed 317",
                   "l2": f"inputs rotate over {NROT} batches ({NROT * B * 3 * IMG * IMG * 4 / 1e6:.0f} MB > L2); "
                         "activations per step exceed L2"},
        "e2e": {"value": e2e_val, "unit": "images/sec", "h2d_bytes_per_step": B * 3 * IMG * IMG * 4,
                "d2h_bytes_per_step": B * K_DET * 56 * 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "decode": decode,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
