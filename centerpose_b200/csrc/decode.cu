// Fused multi_pose_decode for sm_100a: 3x3 max-pool NMS + per-channel top-K + gather +
// keypoint/candidate assignment in ONE launch.
//
// Replaces (reference file:line)  lib/models/decode.py:10-16 (_nms), :87-96 (_topk_channel),
// :99-115 (_topk), :235-308 (multi_pose_decode) and lib/models/utils.py:11-25 (gathers).
//
// HBM-bound design (DESIGN.md "decode kernel"): the only full-map reads are `heat`
// (1 channel) and `hm_hp` (J channels), each read exactly once with coalesced float4 loads
// (+ a one-row halo per row strip); the 2J+6 regression channels are touched only at the
// K centre cells / J*K candidate cells (sparse 4-byte gathers straight from NCHW — the
// reference instead materialises NHWC copies of all of them, utils.py:22).
//
// Grid: one CTA per (image, heat-map channel) = B*(1+J) CTAs of 256 threads.
//   phase 1  each CTA scans its H*W map: thread = (row strip, 4-column group), rolling
//            3-row window in registers, horizontal neighbours by warp shuffle; local maxima
//            above the channel's floor are compacted into shared memory (warp-ballot).
//   phase 2  exact top-K of the candidates: radix-select on the float bits when there are
//            more than 256, then a shared-memory bitonic sort of 64-bit (value,~index) keys.
//            Degenerate inputs (candidate overflow, massive ties, negative maps) take an
//            exact but slow K-round arg-max path so results stay defined everywhere.
//   phase 3  keypoint grouping for joint j is done by whichever of {centre CTA, joint-j CTA}
//            finishes second (one atomic per pair; deadlock-free, no second launch).
#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int CAP = 4096;      // candidate capacity per channel map (32 KB of smem)
constexpr int SORT_N = 256;    // bitonic sort capacity
constexpr int MAXK = CPB200_DECODE_MAX_K;
constexpr unsigned FULL = 0xffffffffu;

struct DecodeParams {
  const float *heat, *wh, *kps, *reg, *hm_hp, *hp_offset;
  float *out;
  float *tk_val;   // (B, 1+J, K)
  int *tk_idx;     // (B, 1+J, K)
  int *sync;       // (B, J) pair counters, zero between launches
  int B, H, W, J, K;
  int apply_sigmoid;
  float thresh;
};

struct Smem {
  float cval[CAP];
  int cidx[CAP];
  unsigned long long skey[SORT_N];
  int hist[256];
  int warp_tot[8];
  float topv[MAXK];
  int topi[MAXK];
  float gx[MAXK], gy[MAXK], gs[MAXK];
  float hd[MAXK];
  int hc[MAXK];
  unsigned long long red[8];
  int count, nsel, digit, krem, flags;
  unsigned todo_mask;
};

__device__ __forceinline__ float act(float v, bool sig) {
  return sig ? __fdividef(1.0f, 1.0f + __expf(-v)) : v;
}

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

__device__ __forceinline__ void push(Smem &s, bool pred, float val, int idx, int lane) {
  unsigned m = __ballot_sync(FULL, pred);
  if (m == 0) return;
  int leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(&s.count, __popc(m));
  base = __shfl_sync(FULL, base, leader);
  if (pred) {
    int pos = base + __popc(m & ((1u << lane) - 1u));
    if (pos < CAP) { s.cval[pos] = val; s.cidx[pos] = idx; }
  }
}

// ---- phase 1, vector path (W % 4 == 0, 16-byte aligned map) -----------------------------
__device__ void scan_vec(Smem &s, const float *__restrict__ map, int H, int W, bool sig,
                         float floorv) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int XG = W >> 2;
  int S = TPB / XG; if (S < 1) S = 1; if (S > H) S = H;
  const int RS = cpb::ceil_div(H, S);
  const int U = XG * S;
  const float NINF = -INFINITY;
  for (int u0 = 0; u0 < U; u0 += TPB) {
    const int u = u0 + tid;
    const bool on = u < U;
    const int g = on ? (u % XG) : 0, st = on ? (u / XG) : 0;
    const int r0 = st * RS;
    const int r1 = min(H, r0 + RS);
    const int x0 = g * 4;
    const bool lpad = (g == 0), rpad = (g == XG - 1);
    const bool lsh = (lane > 0) && !lpad, rsh = (lane < 31) && !rpad;
    const float *base = map + x0;

    auto load_row = [&](int y, float4 &v, float4 &h) {
      const bool ok = on && y >= 0 && y < H;
      if (ok) {
        v = __ldg(reinterpret_cast<const float4 *>(base + (size_t)y * W));
        if (sig) { v.x = act(v.x, true); v.y = act(v.y, true); v.z = act(v.z, true); v.w = act(v.w, true); }
      } else {
        v = make_float4(NINF, NINF, NINF, NINF);
      }
      float lv = __shfl_up_sync(FULL, v.w, 1);
      float rv = __shfl_down_sync(FULL, v.x, 1);
      float l = NINF, r = NINF;
      if (ok) {
        if (!lpad) l = lsh ? lv : act(__ldg(base + (size_t)y * W - 1), sig);
        if (!rpad) r = rsh ? rv : act(__ldg(base + (size_t)y * W + 4), sig);
      }
      h.x = max3(l, v.x, v.y); h.y = max3(v.x, v.y, v.z);
      h.z = max3(v.y, v.z, v.w); h.w = max3(v.z, v.w, r);
    };

    float4 vcur, hprev, hcur, tmp;
    load_row(r0 - 1, tmp, hprev);
    load_row(r0, vcur, hcur);
    for (int i0 = 0; i0 < RS; i0 += 4) {
      float4 nv[4], nh[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int y = r0 + i0 + q + 1;
        load_row((y <= r1) ? y : -1, nv[q], nh[q]);   // row r1 is the strip's lower halo
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int y = r0 + i0 + q;
        const bool rowok = on && (y < r1);
        float mx = max3(hprev.x, hcur.x, nh[q].x), my = max3(hprev.y, hcur.y, nh[q].y);
        float mz = max3(hprev.z, hcur.z, nh[q].z), mw = max3(hprev.w, hcur.w, nh[q].w);
        const int idx = y * W + x0;
        bool px = rowok && (vcur.x == mx), py = rowok && (vcur.y == my);
        bool pz = rowok && (vcur.z == mz), pw = rowok && (vcur.w == mw);
        if ((px && vcur.x < 0.f) || (py && vcur.y < 0.f) || (pz && vcur.z < 0.f) || (pw && vcur.w < 0.f))
          s.flags = 1;   // benign race: negative peak seen (only matters when count < K)
        push(s, px && vcur.x > floorv, vcur.x, idx, lane);
        push(s, py && vcur.y > floorv, vcur.y, idx + 1, lane);
        push(s, pz && vcur.z > floorv, vcur.z, idx + 2, lane);
        push(s, pw && vcur.w > floorv, vcur.w, idx + 3, lane);
        hprev = hcur; hcur = nh[q]; vcur = nv[q];
      }
    }
  }
}

// ---- post-NMS value of one cell straight from global memory (generic / slow paths) -------
__device__ __forceinline__ float nms_value(const float *__restrict__ map, int H, int W, int y,
                                           int x, bool sig, bool *is_peak) {
  const float v = act(__ldg(map + (size_t)y * W + x), sig);
  float m = v;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= W || (dy == 0 && dx == 0)) continue;
      m = fmaxf(m, act(__ldg(map + (size_t)yy * W + xx), sig));
    }
  }
  *is_peak = (m == v);
  return (m == v) ? v : 0.0f;
}

// ---- phase 1, scalar path (any W / alignment) -------------------------------------------
__device__ void scan_scalar(Smem &s, const float *__restrict__ map, int H, int W, bool sig,
                            float floorv) {
  const int N = H * W, lane = threadIdx.x & 31;
  const int iters = cpb::ceil_div(N, TPB);
  for (int it = 0; it < iters; ++it) {
    const int cell = it * TPB + threadIdx.x;
    bool pk = false; float v = 0.f;
    if (cell < N) v = nms_value(map, H, W, cell / W, cell % W, sig, &pk);
    if (pk && v < 0.f) s.flags = 1;
    push(s, pk && v > floorv, v, cell, lane);
  }
}

__device__ __forceinline__ unsigned long long make_key(float v, int idx) {
  unsigned u = __float_as_uint(v == 0.f ? 0.f : v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);
}
__device__ __forceinline__ float key_val(unsigned long long k) {
  unsigned u = (unsigned)(k >> 32);
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  return __uint_as_float(u);
}
__device__ __forceinline__ int key_idx(unsigned long long k) {
  return (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
}

__device__ void bitonic_desc(unsigned long long *key, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < P; t += TPB) {
        const int o = t ^ j;
        if (o > t) {
          const bool desc = ((t & k) == 0);
          unsigned long long a = key[t], b = key[o];
          if ((a < b) == desc) { key[t] = b; key[o] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// exact, general, slow: K rounds of block-wide arg-max over the post-NMS map
__device__ void select_slow(Smem &s, const float *__restrict__ map, int H, int W, bool sig, int K) {
  const int N = H * W, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  unsigned long long bound = ~0ull;
  for (int r = 0; r < K; ++r) {
    unsigned long long best = 0ull;
    for (int cell = tid; cell < N; cell += TPB) {
      bool pk;
      float v = nms_value(map, H, W, cell / W, cell % W, sig, &pk);
      unsigned long long k = make_key(v, cell);
      if (k < bound && k > best) best = k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long other = __shfl_xor_sync(FULL, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) s.red[wid] = best;
    __syncthreads();
    best = s.red[0];
#pragma unroll
    for (int w = 1; w < TPB / 32; ++w) best = s.red[w] > best ? s.red[w] : best;
    if (tid == 0) { s.topv[r] = key_val(best); s.topi[r] = key_idx(best); }
    bound = best;
    __syncthreads();
  }
}

// ---- phase 2: exact top-K of the compacted candidates ----------------------------------
// returns false when the fast path cannot decide (caller falls back to select_slow)
__device__ bool select_fast(Smem &s, int K, bool is_centre, int N) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = s.count;                         // caller guarantees n <= CAP
  int nsort;
  if (n <= SORT_N) {
    for (int i = tid; i < SORT_N; i += TPB)
      s.skey[i] = (i < n) ? make_key(s.cval[i], s.cidx[i]) : 0ull;
    nsort = n;
    __syncthreads();
  } else {
    // radix select (8-bit digits, MSB first) for the K-th largest value; candidates are > 0
    unsigned prefix = 0u, mask = 0u;
    int krem = K;
    for (int d = 3; d >= 0; --d) {
      const int shift = 8 * d;
      s.hist[tid] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += TPB) {
        unsigned u = __float_as_uint(s.cval[i]);
        if ((u & mask) == prefix) atomicAdd(&s.hist[(u >> shift) & 255u], 1);
      }
      __syncthreads();
      const int h = s.hist[tid];
      int suf = h;                                // inclusive suffix sum within the warp
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_down_sync(FULL, suf, o);
        if (lane + o < 32) suf += t;
      }
      if (lane == 0) s.warp_tot[wid] = suf;
      __syncthreads();
      for (int w = wid + 1; w < TPB / 32; ++w) suf += s.warp_tot[w];
      const int excl = suf - h;
      if (excl < krem && krem <= suf) { s.digit = tid; s.krem = krem - excl; }
      __syncthreads();
      prefix |= ((unsigned)s.digit) << shift;
      mask |= 255u << shift;
      krem = s.krem;
      __syncthreads();
    }
    if (tid == 0) s.nsel = 0;
    for (int i = tid; i < SORT_N; i += TPB) s.skey[i] = 0ull;
    __syncthreads();
    for (int i = tid; i < n; i += TPB) {
      if (__float_as_uint(s.cval[i]) >= prefix) {
        int pos = atomicAdd(&s.nsel, 1);
        if (pos < SORT_N) s.skey[pos] = make_key(s.cval[i], s.cidx[i]);
      }
    }
    __syncthreads();
    nsort = s.nsel;
    if (nsort > SORT_N) return false;             // massive ties at the K-th value
  }
  int P = 2;
  while (P < nsort) P <<= 1;
  bitonic_desc(s.skey, P);
  const int take = min(K, nsort);
  for (int i = tid; i < K; i += TPB) {
    if (i < take) { s.topv[i] = key_val(s.skey[i]); s.topi[i] = key_idx(s.skey[i]); }
    else if (!is_centre) { s.topv[i] = -1.0f; s.topi[i] = 0; }   // masked placeholder
  }
  __syncthreads();
  if (is_centre && take < K) {
    // fewer than K positive peaks: the reference's topk then returns zero-valued cells.
    // Canonical choice: the lowest flat indices that are not among the selected peaks.
    if (s.flags) return false;                    // negative peaks present: let the slow path rank
    const int need = K - take;
    int cand = tid;                               // 2K <= 256 candidates, one per thread
    bool free_cell = (cand < N) && (cand < 2 * K);
    if (free_cell) for (int i = 0; i < take; ++i) if (s.topi[i] == cand) { free_cell = false; break; }
    unsigned m = __ballot_sync(FULL, free_cell);
    if (lane == 0) s.warp_tot[wid] = __popc(m);
    __syncthreads();
    int rank = __popc(m & ((1u << lane) - 1u));
    for (int w = 0; w < wid; ++w) rank += s.warp_tot[w];
    if (free_cell && rank < need) { s.topv[take + rank] = 0.0f; s.topi[take + rank] = cand; }
    __syncthreads();
  }
  return true;
}

// ---- phase 3: assign keypoint candidates of joint j to the K people of image b ----------
__device__ void group_joint(Smem &s, const DecodeParams &p, int b, int j) {
  const int tid = threadIdx.x;
  const int K = p.K, J = p.J, W = p.W, N = p.H * p.W;
  const int C1 = 1 + J;
  const float *cv = p.tk_val + ((size_t)b * C1 + (j + 1)) * K;
  const int *ci = p.tk_idx + ((size_t)b * C1 + (j + 1)) * K;
  const float *pv = p.tk_val + ((size_t)b * C1) * K;
  const int *pi = p.tk_idx + ((size_t)b * C1) * K;
  __syncthreads();
  if (tid < K) {
    const float v = __ldcg(cv + tid);
    const int idx = __ldcg(ci + tid);
    if (v > p.thresh) {                                       // decode.py:282-285
      const float fx = (float)(idx % W), fy = (float)(idx / W);
      float ox = 0.5f, oy = 0.5f;
      if (p.hp_offset) {                                      // decode.py:272-277
        ox = __ldg(p.hp_offset + ((size_t)b * 2 + 0) * N + idx);
        oy = __ldg(p.hp_offset + ((size_t)b * 2 + 1) * N + idx);
      }
      s.gx[tid] = __fadd_rn(fx, ox); s.gy[tid] = __fadd_rn(fy, oy); s.gs[tid] = v;
    } else {
      s.gx[tid] = -10000.0f; s.gy[tid] = -10000.0f; s.gs[tid] = -1.0f;
    }
  }
  __syncthreads();
  const int pidx = tid & (MAXK - 1), half = tid >> 7;          // TPB == 2 * MAXK
  float kx = 0.f, ky = 0.f, l = 0.f, t = 0.f, r = 0.f, bt = 0.f, score = 0.f;
  float best_d = INFINITY; int best_c = 0;
  if (pidx < K) {
    const int idx = __ldcg(pi + pidx);
    score = __ldcg(pv + pidx);
    const float fx = (float)(idx % W), fy = (float)(idx / W);
    kx = __fadd_rn(__ldg(p.kps + ((size_t)b * 2 * J + 2 * j) * N + idx), fx);      // :244-247
    ky = __fadd_rn(__ldg(p.kps + ((size_t)b * 2 * J + 2 * j + 1) * N + idx), fy);
    float cx, cy;
    if (p.reg) {                                                                    // :248-255
      cx = __fadd_rn(fx, __ldg(p.reg + ((size_t)b * 2 + 0) * N + idx));
      cy = __fadd_rn(fy, __ldg(p.reg + ((size_t)b * 2 + 1) * N + idx));
    } else { cx = fx + 0.5f; cy = fy + 0.5f; }
    const float hw = __ldg(p.wh + ((size_t)b * 2 + 0) * N + idx) * 0.5f;           // :256-264
    const float hh = __ldg(p.wh + ((size_t)b * 2 + 1) * N + idx) * 0.5f;
    l = __fsub_rn(cx, hw); t = __fsub_rn(cy, hh); r = __fadd_rn(cx, hw); bt = __fadd_rn(cy, hh);
    const int hk = (K + 1) >> 1;
    const int c0 = half ? hk : 0, c1 = half ? K : hk;
    for (int c = c0; c < c1; ++c) {                                                 // :286-289
      const float dx = __fsub_rn(kx, s.gx[c]), dy = __fsub_rn(ky, s.gy[c]);
      const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      if (d < best_d) { best_d = d; best_c = c; }
    }
    if (half) { s.hd[pidx] = best_d; s.hc[pidx] = best_c; }
  }
  __syncthreads();
  if (half == 0 && pidx < K) {
    if (s.hd[pidx] < best_d) { best_d = s.hd[pidx]; best_c = s.hc[pidx]; }
    const float sx = s.gx[best_c], sy = s.gy[best_c], ss = s.gs[best_c];
    const bool rej = (sx < l) || (sx > r) || (sy < t) || (sy > bt) || (ss < p.thresh) ||
                     (best_d > __fmul_rn(fmaxf(__fsub_rn(bt, t), __fsub_rn(r, l)), 0.3f));   // :300-302
    const int row = 5 + 3 * J;
    float *o = p.out + ((size_t)b * K + pidx) * row;
    o[5 + 2 * j] = rej ? kx : sx;
    o[5 + 2 * j + 1] = rej ? ky : sy;
    o[5 + 2 * J + j] = ss;
    if (j == 0) { o[0] = l; o[1] = t; o[2] = r; o[3] = bt; o[4] = score; }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(TPB) decode_kernel(const DecodeParams p) {
  __shared__ Smem s;
  const int C1 = 1 + p.J;
  const int b = blockIdx.x / C1, ch = blockIdx.x % C1;
  const int tid = threadIdx.x;
  const int N = p.H * p.W;
  const bool is_centre = (ch == 0);
  const float *map = is_centre ? (p.heat + (size_t)b * N) : (p.hm_hp + ((size_t)b * p.J + (ch - 1)) * N);
  const bool sig = p.apply_sigmoid != 0;
  const float floorv = is_centre ? 0.0f : p.thresh;
  if (tid == 0) { s.count = 0; s.flags = 0; s.nsel = 0; s.todo_mask = 0; }
  __syncthreads();

  const bool vec_ok = ((p.W & 3) == 0) && ((reinterpret_cast<uintptr_t>(map) & 15) == 0);
  if (vec_ok) scan_vec(s, map, p.H, p.W, sig, floorv);
  else scan_scalar(s, map, p.H, p.W, sig, floorv);
  __syncthreads();

  bool done = false;
  if (s.count <= CAP) done = select_fast(s, p.K, is_centre, N);
  if (!done) {
    __syncthreads();
    select_slow(s, map, p.H, p.W, sig, p.K);
    if (!is_centre)
      for (int i = tid; i < p.K; i += TPB)
        if (!(s.topv[i] > p.thresh)) { s.topv[i] = -1.0f; s.topi[i] = 0; }
    __syncthreads();
  }

  // publish this channel's list, then pair up with the partner CTA(s)
  float *tv = p.tk_val + ((size_t)b * C1 + ch) * p.K;
  int *ti = p.tk_idx + ((size_t)b * C1 + ch) * p.K;
  for (int i = tid; i < p.K; i += TPB) { tv[i] = s.topv[i]; ti[i] = s.topi[i]; }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned todo = 0u;
    if (is_centre) {
      for (int j = 0; j < p.J; ++j)
        if (atomicAdd(&p.sync[(size_t)b * p.J + j], 1) == 1) todo |= (1u << j);
    } else {
      if (atomicAdd(&p.sync[(size_t)b * p.J + (ch - 1)], 1) == 1) todo = 1u << (ch - 1);
    }
    __threadfence();
    s.todo_mask = todo;
  }
  __syncthreads();
  const unsigned todo = s.todo_mask;
  for (int j = 0; j < p.J; ++j) {
    if (todo & (1u << j)) {
      group_joint(s, p, b, j);
      if (tid == 0) p.sync[(size_t)b * p.J + j] = 0;    // restore the zero state for the next launch
    }
  }
}

__global__ void sigmoid_kernel(float *x, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = 1.0f / (1.0f + expf(-x[i]));
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" size_t cpb200_decode_workspace_bytes(int B, int J, int K) {
  if (B <= 0 || J <= 0 || K <= 0) return 0;
  const size_t n = (size_t)B * (1 + J) * K;
  return align_up(n * 4, 16) + align_up(n * 4, 16) + align_up((size_t)B * J * 4, 16);
}

extern "C" int cpb200_multi_pose_decode(const float *heat, const float *wh, const float *kps,
                                        const float *reg, const float *hm_hp,
                                        const float *hp_offset, float *out, int B, int H, int W,
                                        int J, int K, int apply_sigmoid, void *workspace,
                                        size_t workspace_bytes, void *stream) {
  if (!heat || !wh || !kps || !out) return cpb::fail(CPB200_ERR_ARG, "decode: null tensor pointer");
  if (!hm_hp)
    return cpb::fail(CPB200_ERR_ARG, "decode: hm_hp is required (the reference's decode.py:307 "
                                     "uses hm_score unconditionally)");
  if (B <= 0 || H <= 0 || W <= 0) return cpb::fail(CPB200_ERR_ARG, "decode: bad shape B=%d H=%d W=%d", B, H, W);
  if (J < 1 || J > CPB200_DECODE_MAX_J) return cpb::fail(CPB200_ERR_ARG, "decode: J=%d outside [1,%d]", J, CPB200_DECODE_MAX_J);
  if (K < 1 || K > CPB200_DECODE_MAX_K) return cpb::fail(CPB200_ERR_ARG, "decode: K=%d outside [1,%d]", K, CPB200_DECODE_MAX_K);
  if ((long long)K > (long long)H * W)
    return cpb::fail(CPB200_ERR_ARG, "decode: selected index k out of range (K=%d > H*W=%d)", K, H * W);
  if ((long long)H * W > (1ll << 30)) return cpb::fail(CPB200_ERR_ARG, "decode: map too large");
  const size_t need = cpb200_decode_workspace_bytes(B, J, K);
  if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15))
    return cpb::fail(CPB200_ERR_ARG, "decode: workspace too small or misaligned (%zu < %zu)", workspace_bytes, need);
  if ((long long)B * (1 + J) > 2147483647ll) return cpb::fail(CPB200_ERR_ARG, "decode: batch too large");

  DecodeParams p;
  p.heat = heat; p.wh = wh; p.kps = kps; p.reg = reg; p.hm_hp = hm_hp; p.hp_offset = hp_offset;
  p.out = out;
  const size_t n = (size_t)B * (1 + J) * K;
  char *ws = static_cast<char *>(workspace);
  p.tk_val = reinterpret_cast<float *>(ws);
  p.tk_idx = reinterpret_cast<int *>(ws + align_up(n * 4, 16));
  p.sync = reinterpret_cast<int *>(ws + 2 * align_up(n * 4, 16));
  p.B = B; p.H = H; p.W = W; p.J = J; p.K = K;
  p.apply_sigmoid = apply_sigmoid;
  p.thresh = 0.1f;                                   // decode.py:267
  decode_kernel<<<B * (1 + J), TPB, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return cpb::check_launch("decode_kernel");
}

extern "C" int cpb200_sigmoid_inplace(float *x, size_t n, void *stream) {
  if (!x && n) return cpb::fail(CPB200_ERR_ARG, "sigmoid: null pointer");
  if (n == 0) return CPB200_OK;
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 148 * 16) blocks = 148 * 16;
  sigmoid_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n);
  return cpb::check_launch("sigmoid_kernel");
}
