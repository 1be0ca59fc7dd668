"""ORACLE (test infrastructure, not product): CPU restatements of the reference's
modulated deformable convolution (DCNv2) forward.

The reference implements this op ONLY in CUDA against the removed THC API and its CPU
entry point is ``AT_ERROR("Not implement on cpu")`` (``DCNv2/src/cpu/dcn_v2_cpu.cpp:23``),
so it cannot be compiled or run here ("unbuildable").  Two restatements:

  * :func:`dcn_v2_forward_loops` — literal float64 numpy loops, one statement per line of
    ``DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:25-54`` (``dmcn_im2col_bilinear``),
    ``:125-195`` (tap loop, offset/mask indexing, ``> -1 / < H`` bounds) and
    ``DCNv2/src/cuda/dcn_v2_cuda.cu:123-163`` (bias broadcast + GEMM).  Small cases only.
  * :func:`dcn_v2_forward` — the same arithmetic vectorised with torch CPU ops (fp32),
    used by the backbone oracle at full size.

Both are cross-checked in ``tests/test_oracle_dcn.py`` against each other, against
``torchvision.ops.deform_conv2d`` (third-party, torchvision 0.26.0 — what the reference
harness uses as ``_ext`` stub when generating golden vectors), and against the only
known-answer test the reference holds for this op: ``DCNv2/test.py:31-66``
``check_zero_offset`` (identity weights, zero offsets, mask 0.5 => 2*out == in).

:func:`dcn_module_forward` restates ``DCNv2/dcn_v2.py:117-127`` (``DCN.forward``: offset
conv -> chunk(3) -> cat(o1,o2) -> sigmoid(mask) -> modulated deformable conv).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _bilinear_loops(img, H, W, h, w):
    """dcn_v2_im2col_cuda.cu:25-54 (img is one (H,W) channel plane)."""
    h_low = int(np.floor(h)); w_low = int(np.floor(w))
    h_high = h_low + 1; w_high = w_low + 1
    lh = h - h_low; lw = w - w_low
    hh = 1 - lh; hw = 1 - lw
    v1 = img[h_low, w_low] if (h_low >= 0 and w_low >= 0) else 0.0
    v2 = img[h_low, w_high] if (h_low >= 0 and w_high <= W - 1) else 0.0
    v3 = img[h_high, w_low] if (h_high <= H - 1 and w_low >= 0) else 0.0
    v4 = img[h_high, w_high] if (h_high <= H - 1 and w_high <= W - 1) else 0.0
    return hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4


def dcn_v2_forward_loops(inp, weight, bias, offset, mask, stride=1, pad=1, dil=1):
    """Literal restatement (float64).  inp (B,C,H,W); offset (B,2*kh*kw,Ho,Wo) with channel
    2*tap = dh, 2*tap+1 = dw (dcn_v2_im2col_cuda.cu:170-171); mask (B,kh*kw,Ho,Wo);
    weight (Co,C,kh,kw); bias (Co)."""
    inp = np.asarray(inp, np.float64); weight = np.asarray(weight, np.float64)
    bias = np.asarray(bias, np.float64); offset = np.asarray(offset, np.float64)
    mask = np.asarray(mask, np.float64)
    B, C, H, W = inp.shape
    Co, _, kh, kw = weight.shape
    Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1      # dcn_v2_cuda.cu:85
    Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    col = np.zeros((B, C * kh * kw, Ho * Wo))
    for b in range(B):
        for c in range(C):
            for ho in range(Ho):
                for wo in range(Wo):
                    h_in = ho * stride - pad; w_in = wo * stride - pad
                    for i in range(kh):
                        for j in range(kw):
                            tap = i * kw + j
                            oh = offset[b, 2 * tap, ho, wo]; ow = offset[b, 2 * tap + 1, ho, wo]
                            m = mask[b, tap, ho, wo]
                            h_im = h_in + i * dil + oh; w_im = w_in + j * dil + ow
                            val = 0.0
                            if h_im > -1 and w_im > -1 and h_im < H and w_im < W:   # :180
                                val = _bilinear_loops(inp[b, c], H, W, h_im, w_im)
                            col[b, c * kh * kw + tap, ho * Wo + wo] = val * m
    wmat = weight.reshape(Co, C * kh * kw)
    out = np.einsum("ok,bkn->bon", wmat, col) + bias[None, :, None]   # dcn_v2_cuda.cu:123-163
    return out.reshape(B, Co, Ho, Wo)


def dcn_v2_forward(inp, weight, bias, offset, mask, stride=1, pad=1, dil=1):
    """Vectorised torch-CPU restatement of the same arithmetic (fp32).  All torch tensors."""
    B, C, H, W = inp.shape
    Co, _, kh, kw = weight.shape
    Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    dev, dt = inp.device, inp.dtype
    ys = (torch.arange(Ho, device=dev, dtype=dt) * stride - pad).view(1, Ho, 1)
    xs = (torch.arange(Wo, device=dev, dtype=dt) * stride - pad).view(1, 1, Wo)
    flat = inp.reshape(B, C, H * W)
    cols = []
    for i in range(kh):
        for j in range(kw):
            tap = i * kw + j
            h_im = ys + i * dil + offset[:, 2 * tap]          # (B,Ho,Wo)
            w_im = xs + j * dil + offset[:, 2 * tap + 1]
            inside = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
            h_low = torch.floor(h_im); w_low = torch.floor(w_im)
            lh = h_im - h_low; lw = w_im - w_low
            hh = 1 - lh; hw = 1 - lw
            h_low = h_low.long(); w_low = w_low.long()
            h_high = h_low + 1; w_high = w_low + 1
            acc = torch.zeros(B, C, Ho * Wo, device=dev, dtype=dt)
            for (hy, wx, wt) in ((h_low, w_low, hh * hw), (h_low, w_high, hh * lw),
                                 (h_high, w_low, lh * hw), (h_high, w_high, lh * lw)):
                ok = inside & (hy >= 0) & (hy <= H - 1) & (wx >= 0) & (wx <= W - 1)
                idx = (hy.clamp(0, H - 1) * W + wx.clamp(0, W - 1)).reshape(B, 1, Ho * Wo)
                v = flat.gather(2, idx.expand(B, C, Ho * Wo))
                acc = acc + v * (wt * ok.to(dt)).reshape(B, 1, Ho * Wo)
            cols.append(acc * mask[:, tap].reshape(B, 1, Ho * Wo))
    col = torch.stack(cols, dim=2).reshape(B, C * kh * kw, Ho * Wo)   # index c*9+tap
    out = torch.matmul(weight.reshape(Co, C * kh * kw), col)
    if bias is not None:
        out = out + bias.view(1, Co, 1)
    return out.reshape(B, Co, Ho, Wo)


def dcn_module_forward(x, weight, bias, om_weight, om_bias):
    """dcn_v2.py:117-127 (``DCN.forward``), 3x3 / stride 1 / pad 1 / 1 deformable group."""
    out = F.conv2d(x, om_weight, om_bias, stride=1, padding=1)
    o1, o2, m = torch.chunk(out, 3, dim=1)
    offset = torch.cat((o1, o2), dim=1)
    m = torch.sigmoid(m)
    return dcn_v2_forward(x, weight, bias, offset, m, 1, 1, 1)
