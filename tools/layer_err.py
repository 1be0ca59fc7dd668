"""Diagnostics (GPU): per-op error of one precision against the fp32 CUDA-core path, layer by layer, plus a probe of the
tensor core's accumulation rounding (all-positive operands expose a round-toward-zero bias).

    CPB200_NO_REUSE=1 python tools/layer_err.py [precision] [arch] [H] [W]
"""
import os
import sys

os.environ["CPB200_NO_REUSE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch                                                    # noqa: E402
import torch.nn.functional as F                                 # noqa: E402

from centerpose_b200.config import default_cfg                  # noqa: E402
from centerpose_b200.model import create_model                  # noqa: E402
from centerpose_b200.plan import PlanBuilder                    # noqa: E402
from oracle.init_recipe import conditioned_state_dict, synth_images   # noqa: E402

DEV = torch.device("cuda:0")


def run(arch, precision, x):
    cfg = default_cfg(arch)
    m = create_model(cfg.MODEL.NAME, cfg.MODEL.HEAD_CONV, cfg)
    m.load_state_dict(conditioned_state_dict(m.state_dict(), 317))
    m = m.to(DEV).set_precision(precision)
    outs = m(x)
    torch.cuda.synchronize()
    plan = next(iter(m._plans.values()))
    res = []
    for po in plan.pb.ops:
        d = po.dst
        if d.kind in ("act", "actf32", "f32"):
            res.append((po.type, po.cout, d.H, d.W, plan.tensor(d).float().clone()))
        else:
            res.append((po.type, po.cout, d.H, d.W, None))
    return res, [o.clone() for o in outs]


def accumulate_probe():
    """All-positive conv: a truncating accumulator shows a NEGATIVE mean signed error growing with K."""
    for prec in ("fp16x2", "bf16x2"):
        for ci, k in ((64, 1), (64, 3), (512, 3)):
            B, H, W, co = 2, 32, 32, 64
            g = torch.Generator().manual_seed(ci + k)
            x = torch.rand(B, ci, H, W, generator=g) + 0.5
            w = (torch.rand(co, ci, k, k, generator=g) + 0.5) / (ci * k * k)
            ref = F.conv2d(x.double(), w.double(), None, padding=k // 2)
            pb = PlanBuilder(B, 1, 1, prec, DEV)
            y = pb.conv([pb.external(x.permute(0, 2, 3, 1).contiguous().to(DEV))], w.to(DEV), torch.zeros(co, device=DEV), pad=k // 2)
            plan = pb.build(); plan.run(torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
            got = plan.tensor(y).double().permute(0, 3, 1, 2).cpu()
            rel = ((got - ref) / ref)
            ref32 = F.conv2d(x, w, None, padding=k // 2).double()
            rel32 = ((ref32 - ref) / ref)
            print(f"accumulate probe {prec} K={ci * k * k:5d}: mean signed rel err {rel.mean().item():+.3e}  rms {rel.pow(2).mean().sqrt().item():.3e}"
                  f"   (torch fp32 CPU: {rel32.mean().item():+.3e} / {rel32.pow(2).mean().sqrt().item():.3e})")


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x2"
    arch = sys.argv[2] if len(sys.argv) > 2 else "dla_34"
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    W = int(sys.argv[4]) if len(sys.argv) > 4 else H
    accumulate_probe()
    x = synth_images(1, H, W, 317).to(DEV)
    a, oa = run(arch, "fp32", x)
    b, ob = run(arch, prec, x)
    names = {1: "conv", 2: "stem", 3: "maxpool", 4: "up_add", 5: "dcn", 7: "upsample_add", 8: "dwconv", 9: "avgpool", 10: "scale_add", 11: "convert"}
    b = [t for t in b if t[0] != 11]
    a = [t for t in a if t[0] != 11]
    print(f"{arch} {H}x{W}: {len(a)} fp32 ops, {len(b)} {prec} ops")
    for i, (ta, tb) in enumerate(zip(a, b)):
        if ta[4] is None or tb[4] is None:
            continue
        if ta[4].shape != tb[4].shape:
            print(i, "shape mismatch", ta[:4], tb[:4]); continue
        d = (ta[4].double() - tb[4].double())
        rel = (d.norm() / (ta[4].double().norm() + 1e-30)).item()
        mean_signed = (d.sum() / (ta[4].double().abs().sum() + 1e-30)).item()
        print(f"op {i:3d} {names.get(ta[0], ta[0]):8s} cout {ta[1]:4d} {ta[2]:3d}x{ta[3]:<3d} relL2 {rel:.3e}  signed {mean_signed:+.2e}")
    for n, p, q in zip(("hm", "wh", "hps", "reg", "hm_hp", "hp_offset"), oa, ob):
        print(f"head {n:9s} relL2 {((p.double() - q.double()).norm() / p.double().norm()).item():.3e}  max abs {(p - q).abs().max().item():.3e}")


if __name__ == "__main__":
    main()
