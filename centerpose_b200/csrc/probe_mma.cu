// Hardware probe (diagnostics, not on the product path): sustained issue / execution rate of tcgen05.mma.kind::f16 with
// BOTH operands in shared memory (SS mode), as a function of N, the number of accumulators in flight, and cta_group.
// Every SM (or SM pair) runs `iters` back-to-back MMAs of shape (128 * cg) x N x 16 on zero operands; the host times the
// launch with CUDA events.  tools/mma_probe.py prints cycles per instruction and the implied FLOP/s.
#include "tc_common.cuh"

namespace {
using namespace tc;

struct MmaProbeArgs { int n, iters, nacc, stride_bytes, a_sbo, a_shift, alt, pipe, ksteps, mode; };

template <int CG>
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(const MmaProbeArgs a) {
  extern __shared__ __align__(1024) uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t ring[16];          // pipe mode: full[0..7], empty[8..15]
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += 128)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + i * 16), "r"(0u) : "memory");
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&ring[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const bool leader = CG == 1 || cluster_ctarank() == 0;
  if (warp == 1 && leader) {
    if (elect_one()) {
      const uint32_t idesc = idesc_mn(128 * CG, a.n, 1u);
      // K-major, 128-byte rows (SW128): A = 128 rows at base, B = n (or n/2 per CTA) rows at base + 32 KB; successive MMAs step
      // through `stride_bytes` so that operand fetches are not all the same lines
      // a_sbo != 0: the A operand is a shifted window of a halo tile (net_tc3.cu): 8-row groups a_sbo bytes apart
      // (10 pixels x 128 B = 1280 for a 16x8 tile of 64-channel pixels) starting a_shift bytes into the tile.
      uint64_t ad0 = make_desc(base + (uint32_t)a.a_shift, 128, 2);
      if (a.a_sbo) ad0 = (ad0 & ~((uint64_t)0x3FFF << 32)) | ((uint64_t)((uint32_t)a.a_sbo >> 4) << 32);
      const uint64_t bd0 = make_desc(base + 32 * 1024, 128, 2);
      const uint32_t idesc_half = idesc_mn(128 * CG, a.n / 2, 1u);
      if (a.pipe > 0 && CG == 1) {
        // pipe mode: the issue loop of the conv kernels — per "stage" wait a full barrier (armed by the stand-in producer
        // in warp 2), issue `ksteps` K steps, tcgen05.commit to the stage's empty barrier; `iters` counts stages.
        // mode bits: 1 = no full-barrier wait, 2 = plain mbarrier.arrive instead of tcgen05.commit, 4 = wait for the NEXT stage's
        // full barrier before the last K step of this one (its latency then overlaps queued MMAs), 8 = no fence after the wait
        int stage = 0; uint32_t phase = 0;
        const bool nowait = a.mode & 1, nocommit = a.mode & 2, early = a.mode & 4, nofence = a.mode & 8;
        if (early) mbar_wait(smem_u32(&ring[0]), 0);
        for (int i = 0; i < a.iters; ++i) {
          if (!nowait && !early) mbar_wait(smem_u32(&ring[stage]), phase);
          if (!nofence) tc_fence_after();
          const uint32_t d = tmem + (uint32_t)(i % a.nacc) * (uint32_t)a.n;
          int nstage = stage + 1; uint32_t nphase = phase;
          if (nstage == a.pipe) { nstage = 0; nphase ^= 1; }
          for (int k = 0; k < a.ksteps; ++k) {
            if (early && k == a.ksteps - 1 && i + 1 < a.iters) mbar_wait(smem_u32(&ring[nstage]), nphase);
            const uint32_t off = (uint32_t)(k * 32) >> 4;
            umma_bf16(d, ad0 + off, bd0 + off, idesc, (i >= a.nacc || k > 0) ? 1u : 0u);
            if (a.alt) umma_bf16(d + (uint32_t)a.n / 2, ad0 + off + 2, bd0 + off, idesc_half, 1u);
          }
          if (nocommit) mbar_arrive(smem_u32(&ring[8 + stage]));
          else umma_commit(smem_u32(&ring[8 + stage]));
          stage = nstage; phase = nphase;
        }
      } else
      for (int i = 0; i < a.iters; ++i) {
        const uint32_t off = ((uint32_t)(i & 3) * (uint32_t)a.stride_bytes) >> 4;
        const uint32_t d = tmem + (uint32_t)(i % a.nacc) * (uint32_t)a.n;
        if constexpr (CG == 2) umma_f16_cg2(d, ad0 + off, bd0 + off, idesc, i >= a.nacc ? 1u : 0u);
        else {
          // alt: the split-operand K step — A_hi x [W_hi|W_lo] (N) followed by A_lo x W_hi (N/2) into the upper half
          umma_bf16(d, ad0 + off, bd0 + off, idesc, i >= a.nacc ? 1u : 0u);
          if (a.alt) umma_bf16(d + (uint32_t)a.n / 2, ad0 + off + 2, bd0 + off, idesc_half, 1u);
        }
      }
      if constexpr (CG == 2) umma_commit_cg2(smem_u32(&bar), (uint16_t)1);
      else umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
  }
  if (warp == 2 && a.pipe > 0 && CG == 1 && (threadIdx.x & 31) == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < a.iters; ++i) {
      mbar_wait(smem_u32(&ring[8 + stage]), phase ^ 1);
      mbar_arrive(smem_u32(&ring[stage]));
      if (++stage == a.pipe) { stage = 0; phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if constexpr (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ---- pipe2: the issue loop written the way the production kernels should run it ------------------------------------
// K steps unrolled at compile time, no integer division, the NEXT stage's full barrier polled once before this stage's
// MMAs (its round trip overlaps queued tensor work), and — UNIFORM — the whole warp running the loop with uniform control
// flow and only the tcgen05 instructions themselves under elect.sync, so that descriptors live in uniform registers.
struct PipeArgs { int n, iters, nacc, alt, ring; };

template <int KS, bool UNIFORM>
__global__ void __launch_bounds__(128, 1) mma_pipe2_kernel(const PipeArgs a) {
  extern __shared__ __align__(1024) uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t ring[16];
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += 128)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + i * 16), "r"(0u) : "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&ring[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (warp == 1 && (UNIFORM || elect_one())) {
    const uint32_t idesc = idesc_mn(128, a.n, 1u), idesc_half = idesc_mn(128, a.n / 2, 1u);
    const uint64_t ad0 = make_desc(base, 128, 2), bd0 = make_desc(base + 32 * 1024, 128, 2);
    const uint32_t full0 = smem_u32(&ring[0]), empty0 = smem_u32(&ring[8]);
    int stage = 0, acc = 0; uint32_t phase = 0, first = 1;
    uint32_t peek = mbar_try_once(full0, 0);
    for (int i = 0; i < a.iters; ++i) {
      if (!peek) mbar_wait(full0 + 8 * stage, phase);
      tc_fence_after();
      const uint32_t d = tmem + (uint32_t)acc * (uint32_t)a.n;
      int ns = stage + 1; uint32_t np = phase;
      if (ns == a.ring) { ns = 0; np ^= 1; }
      peek = mbar_try_once(full0 + 8 * ns, np);
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        if (!UNIFORM || elect_one()) {
          umma_bf16(d, ad0 + 2 * k, bd0 + 2 * k, idesc, (k > 0 || !first) ? 1u : 0u);
          if (a.alt) umma_bf16(d + (uint32_t)a.n / 2, ad0 + 2 * k + 64, bd0 + 2 * k, idesc_half, 1u);
        }
      }
      if (!UNIFORM || elect_one()) umma_commit(empty0 + 8 * stage);
      stage = ns; phase = np;
      if (++acc == a.nacc) { acc = 0; first = 0; }
    }
    // drain: wait until the last commit has landed
    if (!UNIFORM || lane == 0) {
      int ls = (a.iters - 1) % a.ring;
      mbar_wait(empty0 + 8 * ls, (uint32_t)(((a.iters - 1) / a.ring) & 1));
    }
  } else if (warp == 2 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < a.iters; ++i) {
      if (i >= a.ring) mbar_wait(smem_u32(&ring[8 + stage]), phase ^ 1);
      mbar_arrive(smem_u32(&ring[stage]));
      if (++stage == a.ring) { stage = 0; phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}
}  // namespace

extern "C" int cpb200_probe_mma_ex(int n, int cg, int iters, int nacc, int stride_bytes, int a_sbo, int a_shift, int alt, void *stream);
extern "C" int cpb200_probe_mma_pipe(int n, int iters, int nacc, int alt, int stages, int ksteps, int mode, void *stream);
extern "C" int cpb200_probe_mma(int n, int cg, int iters, int nacc, int stride_bytes, void *stream) {
  return cpb200_probe_mma_ex(n, cg, iters, nacc, stride_bytes, 0, 0, 0, stream);
}
static int g_pipe = 0, g_ksteps = 0, g_mode = 0;
extern "C" int cpb200_probe_mma_pipe(int n, int iters, int nacc, int alt, int stages, int ksteps, int mode, void *stream) {
  if (stages < 1 || stages > 8 || ksteps < 1 || ksteps > 64) return cpb::fail(CPB200_ERR_ARG, "probe_mma_pipe: bad arguments");
  g_pipe = stages; g_ksteps = ksteps; g_mode = mode;
  const int rc = cpb200_probe_mma_ex(n, 1, iters, nacc, 32, 0, 0, alt, stream);
  g_pipe = 0; g_ksteps = 0; g_mode = 0;
  return rc;
}
extern "C" int cpb200_probe_mma_ex(int n, int cg, int iters, int nacc, int stride_bytes, int a_sbo, int a_shift, int alt, void *stream) {
  if (n < 16 || n > 256 || n % 32 || (cg != 1 && cg != 2) || nacc < 1 || nacc * n > 512 || iters < 1 || a_sbo % 16 || a_shift % 16 ||
      a_sbo > 1536 || a_shift > 4096)
    return cpb::fail(CPB200_ERR_ARG, "probe_mma: bad arguments");
  MmaProbeArgs a{n, iters, nacc, stride_bytes, a_sbo, a_shift, alt, g_pipe, g_ksteps, g_mode};
  const size_t smem = 100 * 1024;
  const int sms = tc::num_sms() & ~1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cg == 1) {
    static tc::SmemAttrCache c1;
    if (int rc = tc::ensure_smem(mma_probe_kernel<1>, smem, c1)) return rc;
    mma_probe_kernel<1><<<sms, 128, smem, st>>>(a);
  } else {
    static tc::SmemAttrCache c2;
    if (int rc = tc::ensure_smem(mma_probe_kernel<2>, smem, c2)) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)sms); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CPB_CUDA(cudaLaunchKernelEx(&cfg, mma_probe_kernel<2>, a));
  }
  return cpb::check_launch("mma_probe_kernel");
}

extern "C" int cpb200_probe_mma_pipe2(int n, int iters, int nacc, int alt, int ring, int ksteps, int uniform, void *stream) {
  if (n < 32 || n > 256 || n % 32 || nacc < 1 || nacc * n > 512 || iters < 1 || ring < 1 || ring > 8 || (ksteps != 4 && ksteps != 8))
    return cpb::fail(CPB200_ERR_ARG, "probe_mma_pipe2: bad arguments");
  PipeArgs a{n, iters, nacc, alt, ring};
  const size_t smem = 100 * 1024;
  const int sms = tc::num_sms();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define PIPE2(KS, U)                                                                  \
  {                                                                                   \
    static tc::SmemAttrCache c;                                                       \
    if (int rc = tc::ensure_smem(mma_pipe2_kernel<KS, U>, smem, c)) return rc;        \
    mma_pipe2_kernel<KS, U><<<sms, 128, smem, st>>>(a);                               \
  }
  if (ksteps == 4 && uniform) PIPE2(4, true)
  else if (ksteps == 4) PIPE2(4, false)
  else if (uniform) PIPE2(8, true)
  else PIPE2(8, false)
#undef PIPE2
  return cpb::check_launch("mma_pipe2_kernel");
}
