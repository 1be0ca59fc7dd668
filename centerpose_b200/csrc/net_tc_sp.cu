// Small-channel 3x3 convolutions (Cin = 16 or 32, stride 1 or 2, pad 1) on tcgen05 with SIMT-fed operands.
//
// Why not TMA here: a tiled TMA box is fetched row by row, one request per (pixel row of the box) — measured
// ~2.5-3 ns per request per SM whatever its width.  With 16 or 32 channels a request is only 32 / 64 bytes, so
// the halo-reuse kernel (csrc/net_tc3.cu, 180 requests per tile) and the tap-per-stage kernel (csrc/net_tc.cu,
// 9 x 128 requests per tile for stride 2) spend 0.5-3.7 us per 128-pixel tile waiting for the copy engine:
// DLA-34 level0 (16->16 @512^2) 437 us and level1 (16->32 /2) 410 us against a ~100 us HBM floor.  Here producer
// THREADS fetch the halo with coalesced 16-byte loads (a tile row is one contiguous 320..1088-byte segment of the
// NHWC tensor) and store it straight into the swizzled K-major layout the UMMA descriptors read — the same
// scheme as the stem (csrc/net_stem_tc.cu): next tile's loads in flight while this tile is stored, several
// producer groups on alternate tiles, no CTA-wide barrier.
//
//   halo     stride 1: 18 x 10 pixels, row R = hy*10 + hx;  tap (r,s) = window starting (r*10 + s) rows later,
//            8-row-group stride 10 rows (as in net_tc3.cu).
//            stride 2: 33 x 17 pixels split into four parity planes of 17 x 9 (R = plane*153 + (hy/2)*9 + hx/2)
//            so that the 8 pixels of an output row are again 8 CONSECUTIVE operand rows: tap (r,s) = plane
//            (r&1, s&1), window offset ((r/2)*9 + s/2) rows, group stride 9 rows.
//   swizzle  32-byte (C=16) / 64-byte (C=32) pattern applied on absolute shared-memory address bits
//            (chunk bit(s) [4..] ^= address bits [7..]); stage bases are 1024-aligned.
//   weights  the ordinary tensor-core packing [tap][1 slab][N][C] bf16 (plan.py::_pack_conv_tc, bk = C), swizzled while being
//            copied to shared memory once per CTA.
//   split    P = 2 (CPB200_BF16X2 / CPB200_F16X2, tc_common.cuh): the stage ring is plane-granular — a tile occupies two
//            consecutive stages (hi halo, lo halo), fetched as two units of the producers' load pipeline; weights sit in
//            shared memory as [tap][hi tile | lo tile], so A_hi x [W_hi ; W_lo] is ONE tcgen05.mma of N = 2N (N <= 64
//            here) into two accumulator halves and A_lo x W_hi a second one; the epilogue adds the halves, applies
//            acc_scale / bias / residual / activation in fp32 and stores the hi and lo planes.
#include "tc_common.cuh"

#include <cstdlib>

namespace {

using namespace tc;

constexpr int SP_TH = 16, SP_TW = 8;
constexpr int SP_GROUPS = 2;
constexpr int SP_PT = 256;                                // producer threads per group
constexpr int SP_THREADS = SP_GROUPS * SP_PT + 5 * 32;    // + MMA warp + 4 epilogue warps
constexpr int SP_NACC = 4;

struct SpArgs {
  const __nv_bfloat16 *x;      // (B,H,W,C)            [P = 2: hi plane, lo plane x_plane elements later; 16-bit either format]
  const __nv_bfloat16 *w;      // [9][N][C]            [P = 2: [plane][9][N][C]]
  const __nv_bfloat16 *res;    // optional (B,Ho,Wo,N) [planes y_plane apart]
  __nv_bfloat16 *y;            // (B,Ho,Wo,N)          [planes y_plane apart]
  const float *bias;
  int B, H, W, Ho, Wo;
  int tiles_h, tiles_w, total_tiles;
  uint32_t act;
  uint32_t fmt;                // split mode: 0 = bf16 planes, 1 = fp16 planes
  float acc_scale;
  long long x_plane, y_plane;
};

template <int C, int S>
struct SpGeom {
  static constexpr int PIX_B = C * 2;                     // bytes per pixel row of the operand
  static constexpr int CH = PIX_B / 16;                   // 16-byte chunks per pixel
  static constexpr int HH = (SP_TH - 1) * S + 3, HW = (SP_TW - 1) * S + 3;       // halo extent in input pixels
  static constexpr int PLANE_W = (S == 1) ? HW : (HW + 1) / 2;                   // operand rows per halo row (per plane)
  static constexpr int PLANE_H = (S == 1) ? HH : (HH + 1) / 2;
  static constexpr int PLANE_ROWS = PLANE_W * PLANE_H;
  static constexpr int ROWS = (S == 1) ? PLANE_ROWS : 4 * PLANE_ROWS;
  static constexpr int STAGE_BYTES = ((ROWS * PIX_B + 1023) / 1024) * 1024;
  static constexpr int NCHUNK = HH * HW * CH;             // 16-byte chunks fetched per tile
  static constexpr int NLD = (NCHUNK + SP_PT - 1) / SP_PT;
  static constexpr uint32_t SWMASK = (C == 16) ? 1u : 3u;
  static constexpr uint32_t LAYOUT = (C == 16) ? 6u : 4u; // UMMA layout type: 32-byte / 64-byte swizzle
  static constexpr int NSTAGE = (STAGE_BYTES <= 12 * 1024) ? 6 : (STAGE_BYTES <= 20 * 1024 ? 5 : 4);
  // split mode: plane-granular stages, as many as fit beside the two weight planes (at most 8, at least 3)
  template <int N_>
  struct Split {
    static constexpr int K_ = (200 * 1024 - 2 * 9 * N_ * PIX_B) / STAGE_BYTES;
    static constexpr int NSTAGE = K_ > 8 ? 8 : K_;
  };
};

__device__ __forceinline__ uint32_t sp_swz(uint32_t off, uint32_t mask) {      // offset within a 1024-aligned region
  return off ^ (((off >> 7) & mask) << 4);
}

template <int C, int N, int S, int P>
__global__ void __launch_bounds__(SP_THREADS, 1) conv_sp_kernel(const SpArgs a) {
  using G = SpGeom<C, S>;
  constexpr int NSTAGE = P == 2 ? G::template Split<N>::NSTAGE : G::NSTAGE;
  static_assert(NSTAGE >= 3 && NSTAGE <= 8, "stage ring does not fit");
  constexpr int B_TILE_BYTES = N * G::PIX_B;               // one plane of one tap
  constexpr int B_TAP_BYTES = P * B_TILE_BYTES;            // [hi tile | lo tile]
  constexpr int ACC_COLS = P * N;                          // P = 2: two accumulator halves (hi x hi + lo x hi | hi x lo)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + NSTAGE * G::STAGE_BYTES;
  __shared__ __align__(8) uint64_t bars[2 * 8 + 2 * SP_NACC];
  __shared__ uint32_t s_tmem;
  __shared__ float s_bias[N];
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]);
  const uint32_t tfull0 = smem_u32(&bars[16]), tempty0 = smem_u32(&bars[16 + SP_NACC]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int MMA_WARP = SP_GROUPS * SP_PT / 32;
  constexpr uint32_t TMEM_COLS = (SP_NACC * ACC_COLS) < 32 ? 32u : (uint32_t)(SP_NACC * ACC_COLS);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full0 + 8 * s, SP_PT); mbar_init(empty0 + 8 * s, 1); }
    for (int s = 0; s < SP_NACC; ++s) { mbar_init(tfull0 + 8 * s, 1); mbar_init(tempty0 + 8 * s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // rows the producers never write (plane padding of the stride-2 layout) must hold finite values: zero everything once
  for (int i = threadIdx.x; i < NSTAGE * G::STAGE_BYTES / 16; i += SP_THREADS)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a_base + i * 16), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < 9 * B_TAP_BYTES / 16; i += SP_THREADS) {          // weights: dense [plane][tap][n][c] -> swizzled [tap][plane][n][c]
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(a.w) + i);
    constexpr int TILE_CH = B_TILE_BYTES / 16;
    const int blk = i / TILE_CH, within = i - blk * TILE_CH;                      // blk = plane * 9 + tap
    const int pl = blk / 9, tap = blk - 9 * pl;
    const uint32_t dst = b_base + sp_swz((uint32_t)((tap * P + pl) * TILE_CH + within) * 16u, G::SWMASK);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
  for (int i = threadIdx.x; i < N; i += SP_THREADS) s_bias[i] = a.bias ? __ldg(a.bias + i) : 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  auto decode_tile = [&](int t, int &n, int &h0, int &w0) {
    const int tw = t % a.tiles_w; t /= a.tiles_w;
    const int th = t % a.tiles_h; n = t / a.tiles_h;
    h0 = th * SP_TH; w0 = tw * SP_TW;
  };
  auto sbo_desc = [](uint32_t saddr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)G::LAYOUT << 61;
    return d;
  };

  if (warp < MMA_WARP) {
    // =============================== producers ===============================
    const int grp = threadIdx.x / SP_PT;
    const int p = threadIdx.x - grp * SP_PT;
    // chunk i = (hy*HW + hx)*CH + j of the halo: source offset (elements, relative to the halo origin) is tile
    // dependent only through (hi0, wi0); destination offset inside a stage is fixed -> precomputed.
    uint32_t doff[G::NLD];
    int hyx[G::NLD];                                        // hy << 16 | hx << 8 | j, or -1
#pragma unroll
    for (int q = 0; q < G::NLD; ++q) {
      const int i = p + q * SP_PT;
      if (i < G::NCHUNK) {
        const int j = i % G::CH, px = i / G::CH;
        const int hx = px % G::HW, hy = px / G::HW;
        int R;
        if (S == 1) R = hy * G::PLANE_W + hx;
        else R = ((hy & 1) * 2 + (hx & 1)) * G::PLANE_ROWS + (hy >> 1) * G::PLANE_W + (hx >> 1);
        doff[q] = sp_swz((uint32_t)(R * G::PIX_B + j * 16), G::SWMASK);
        hyx[q] = (hy << 16) | (hx << 8) | j;
      } else {
        doff[q] = 0; hyx[q] = -1;
      }
    }
    uint4 pre[G::NLD];
    auto fetch = [&](int t, int pl) {
      int n, h0, w0; decode_tile(t, n, h0, w0);
      const int hi0 = h0 * S - 1, wi0 = w0 * S - 1;
      const __nv_bfloat16 *xin = a.x + (size_t)n * a.H * a.W * C + (P == 2 ? (size_t)pl * a.x_plane : 0);
#pragma unroll
      for (int q = 0; q < G::NLD; ++q) {
        const int hy = hyx[q] >> 16, hx = (hyx[q] >> 8) & 255, j = hyx[q] & 255;
        const int hi = hi0 + hy, wi = wi0 + hx;
        const bool ok = hyx[q] >= 0 && hi >= 0 && hi < a.H && wi >= 0 && wi < a.W;
        pre[q] = ok ? __ldg(reinterpret_cast<const uint4 *>(xin + ((size_t)hi * a.W + wi) * C) + j) : make_uint4(0u, 0u, 0u, 0u);
      }
    };
    const int tstep = gridDim.x * SP_GROUPS;
    const int t_first = blockIdx.x + grp * gridDim.x;
    int it = grp;
    if (t_first < a.total_tiles) fetch(t_first, 0);
    for (int t = t_first; t < a.total_tiles; t += tstep, it += SP_GROUPS) {
#pragma unroll
      for (int pl = 0; pl < P; ++pl) {                       // a tile = P consecutive plane stages
        const int vs = it * P + pl;
        const int stage = vs % NSTAGE;
        const uint32_t phase = (uint32_t)(vs / NSTAGE) & 1u;
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        const uint32_t sa = a_base + stage * G::STAGE_BYTES;
#pragma unroll
        for (int q = 0; q < G::NLD; ++q)
          if (hyx[q] >= 0)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa + doff[q]), "r"(pre[q].x), "r"(pre[q].y), "r"(pre[q].z),
                         "r"(pre[q].w) : "memory");
        // the next unit's loads fly while the MMA warp consumes this one
        if (pl + 1 < P) fetch(t, pl + 1);
        else if (t + tstep < a.total_tiles) fetch(t + tstep, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(full0 + 8 * stage);
      }
    }
  } else if (warp == MMA_WARP) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = P == 1 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))
                                  : idesc_m128(N, a.fmt);
    const uint32_t idesc2 = idesc_m128(2 * N, a.fmt);       // P = 2: A_hi x [W_hi ; W_lo]
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t accphase = 0;
    for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
      mbar_wait(tempty0 + 8 * acc, accphase ^ 1);
#pragma unroll
      for (int pl = 0; pl < P; ++pl) {
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
          const uint64_t ad0 = sbo_desc(a_base + stage * G::STAGE_BYTES, G::PLANE_W * G::PIX_B);
          const uint64_t bd0 = sbo_desc(b_base, 8 * G::PIX_B);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, s = tap % 3;
            const int row0 = (S == 1) ? (r * G::PLANE_W + s)
                                      : (((r & 1) * 2 + (s & 1)) * G::PLANE_ROWS + (r >> 1) * G::PLANE_W + (s >> 1));
#pragma unroll
            for (int k = 0; k < C / 16; ++k)
              umma_bf16(d_tmem + (pl ? N : 0), ad0 + (uint32_t)(row0 * (G::PIX_B >> 4) + 2 * k), bd0 + (uint32_t)(tap * (B_TAP_BYTES >> 4) + 2 * k),
                        (P == 2 && pl == 0) ? idesc2 : idesc, (pl > 0 || tap > 0 || k > 0) ? 1u : 0u);   // lo x hi -> second half
          }
          umma_commit(empty0 + 8 * stage);
          if (pl == P - 1) umma_commit(tfull0 + 8 * acc);
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
      if (++acc == SP_NACC) { acc = 0; accphase ^= 1; }
    }
  } else {
    // =============================== epilogue (4 warps) ===============================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int ty = m >> 3, tx = m & 7;
    int acc = 0; uint32_t accphase = 0;
    for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
      int n, h0, w0; decode_tile(t, n, h0, w0);
      const int ho = h0 + ty, wo = w0 + tx;
      const bool ok = ho < a.Ho && wo < a.Wo;
      const size_t pix = ((size_t)n * a.Ho + ho) * a.Wo + wo;
      __nv_bfloat16 *o = a.y + pix * N;
      uint4 rr[P == 1 ? N / 8 : 1];
      if (P == 1 && a.res && ok) {
#pragma unroll
        for (int c = 0; c < N / 8; ++c) rr[c] = __ldg(reinterpret_cast<const uint4 *>(a.res + pix * N) + c);
      }
      mbar_wait(tfull0 + 8 * acc, accphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS;
#pragma unroll
      for (int c = 0; c < N / 16; ++c) {
        uint32_t v[16];
        tmem_ld16(taddr + c * 16, v);
        if constexpr (P == 2) {
          uint32_t v2[16];
          tmem_ld16(taddr + N + c * 16, v2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        } else {
          tmem_ld_wait();
        }
        if (ok) {
          float f[16];
          if constexpr (P == 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaf(__uint_as_float(v[j]), a.acc_scale, s_bias[c * 16 + j]);
            uint16_t *oh_ = reinterpret_cast<uint16_t *>(o) + c * 16;
            if (a.res) {
              const uint16_t *rh = reinterpret_cast<const uint16_t *>(a.res) + pix * N + c * 16;
              const uint4 h0 = __ldg(reinterpret_cast<const uint4 *>(rh)), h1 = __ldg(reinterpret_cast<const uint4 *>(rh) + 1);
              const uint4 l0 = __ldg(reinterpret_cast<const uint4 *>(rh + a.y_plane)), l1 = __ldg(reinterpret_cast<const uint4 *>(rh + a.y_plane) + 1);
              const uint32_t hw_[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
              const uint32_t lw_[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 x = join2(hw_[j], lw_[j], a.fmt);
                f[2 * j] += x.x; f[2 * j + 1] += x.y;
              }
            }
            uint32_t oh[8], ol[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              split2(cpb::act_fast(f[2 * j], a.act), cpb::act_fast(f[2 * j + 1], a.act), a.fmt, oh[j], ol[j]);
            st_global_32B(oh_, oh);
            st_global_32B(oh_ + a.y_plane, ol);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + s_bias[c * 16 + j];
            if (a.res) {
              const __nv_bfloat162 *rb0 = reinterpret_cast<const __nv_bfloat162 *>(&rr[2 * c]);
              const __nv_bfloat162 *rb1 = reinterpret_cast<const __nv_bfloat162 *>(&rr[2 * c + 1]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 x0 = __bfloat1622float2(rb0[j]), x1 = __bfloat1622float2(rb1[j]);
                f[2 * j] += x0.x; f[2 * j + 1] += x0.y; f[8 + 2 * j] += x1.x; f[8 + 2 * j + 1] += x1.y;
              }
            }
            uint4 o0, o1;
            __nv_bfloat162 *ob0 = reinterpret_cast<__nv_bfloat162 *>(&o0), *ob1 = reinterpret_cast<__nv_bfloat162 *>(&o1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              ob0[j] = __floats2bfloat162_rn(cpb::act_out<__nv_bfloat16>(f[2 * j], a.act), cpb::act_out<__nv_bfloat16>(f[2 * j + 1], a.act));
              ob1[j] = __floats2bfloat162_rn(cpb::act_out<__nv_bfloat16>(f[8 + 2 * j], a.act), cpb::act_out<__nv_bfloat16>(f[8 + 2 * j + 1], a.act));
            }
            const uint32_t ow[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
            st_global_32B(o + c * 16, ow);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      if (++acc == SP_NACC) { acc = 0; accphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int C, int N, int S, int P>
int launch_sp(const cpb200_op &op, cudaStream_t st) {
  using G = SpGeom<C, S>;
  constexpr int NSTAGE = P == 2 ? G::template Split<N>::NSTAGE : G::NSTAGE;
  SpArgs a;
  a.x = static_cast<const __nv_bfloat16 *>(op.src[0]); a.w = static_cast<const __nv_bfloat16 *>(op.weight);
  a.res = static_cast<const __nv_bfloat16 *>(op.res); a.y = static_cast<__nv_bfloat16 *>(op.dst); a.bias = op.bias;
  a.B = op.B; a.H = op.H; a.W = op.W; a.Ho = op.Ho; a.Wo = op.Wo;
  a.tiles_h = (op.Ho + SP_TH - 1) / SP_TH; a.tiles_w = (op.Wo + SP_TW - 1) / SP_TW;
  a.total_tiles = op.B * a.tiles_h * a.tiles_w;
  a.act = op.flags & CPB_ACT_MASK;
  a.fmt = op.act_dtype == CPB200_F16X2 ? 1u : 0u;
  a.acc_scale = op.acc_scale != 0.f ? op.acc_scale : 1.f;
  a.x_plane = (long long)op.B * op.H * op.W * C;
  a.y_plane = (long long)op.B * op.Ho * op.Wo * N;
  const size_t smem = 1024 + (size_t)NSTAGE * G::STAGE_BYTES + (size_t)P * 9 * N * G::PIX_B;
  static SmemAttrCache cache;
  if (int rc = ensure_smem(conv_sp_kernel<C, N, S, P>, smem, cache)) return rc;
  const int sms = tc::num_sms();
  const int grid = a.total_tiles < sms ? a.total_tiles : sms;
  conv_sp_kernel<C, N, S, P><<<grid, SP_THREADS, smem, st>>>(a);
  return cpb::check_launch("conv_sp_kernel");
}

template <int C, int S>
int dispatch_n(const cpb200_op &op, cudaStream_t st) {
  const bool split = op.act_dtype != CPB200_BF16;
  switch (op.cout) {
    case 16: return split ? launch_sp<C, 16, S, 2>(op, st) : launch_sp<C, 16, S, 1>(op, st);
    case 32: return split ? launch_sp<C, 32, S, 2>(op, st) : launch_sp<C, 32, S, 1>(op, st);
    case 64: return split ? launch_sp<C, 64, S, 2>(op, st) : launch_sp<C, 64, S, 1>(op, st);
  }
  return cpb::fail(CPB200_ERR_ARG, "conv_sp: cout %d", op.cout);
}

}  // namespace

namespace cpb {

bool sp_eligible(const cpb200_op &op) {
  static const bool enabled = []() { const char *e = getenv("CPB200_SP"); return !(e && e[0] == '0'); }();
  return enabled && op.type == CPB200_OP_CONV && (op.flags & CPB200_FLAG_TC) &&
         (op.act_dtype == CPB200_BF16 || op.act_dtype == CPB200_BF16X2 || op.act_dtype == CPB200_F16X2) && op.nsrc == 1 &&
         (op.src_pitch[0] == 0 || op.src_pitch[0] == op.cin[0]) &&
         (op.cin[0] == 16 || op.cin[0] == 32) && (op.cout == 16 || op.cout == 32 || op.cout == 64) && op.kh == 3 && op.kw == 3 &&
         op.pad_h == 1 && op.pad_w == 1 && (op.stride == 1 || op.stride == 2) &&
         op.Ho == (op.H + 2 - 3) / op.stride + 1 && op.Wo == (op.W + 2 - 3) / op.stride + 1 &&
         op.out_sy == 1 && op.out_sx == 1 && !op.out_oy && !op.out_ox && op.Hd == op.Ho && op.Wd == op.Wo &&
         !(op.flags & (CPB200_FLAG_OUT_NCHW_F32 | CPB200_FLAG_OUT_F32)) && op.Wo >= 8;
}

int sp_run(const cpb200_op &op, cudaStream_t st) {
  if (!sp_eligible(op)) return fail(CPB200_ERR_ARG, "conv_sp: unsupported shape");
  if (op.cin[0] == 16) return op.stride == 1 ? dispatch_n<16, 1>(op, st) : dispatch_n<16, 2>(op, st);
  return op.stride == 1 ? dispatch_n<32, 1>(op, st) : dispatch_n<32, 2>(op, st);
}

}  // namespace cpb
