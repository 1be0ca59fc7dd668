// Hardware probe (diagnostics, not on the product path): sustained issue / execution rate of tcgen05.mma.kind::f16 with
// BOTH operands in shared memory (SS mode), as a function of N, the number of accumulators in flight, and cta_group.
// Every SM (or SM pair) runs `iters` back-to-back MMAs of shape (128 * cg) x N x 16 on zero operands; the host times the
// launch with CUDA events.  tools/mma_probe.py prints cycles per instruction and the implied FLOP/s.
#include "tc_common.cuh"

namespace {
using namespace tc;

struct MmaProbeArgs { int n, iters, nacc, stride_bytes; };

template <int CG>
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(const MmaProbeArgs a) {
  extern __shared__ __align__(1024) uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += 128)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + i * 16), "r"(0u) : "memory");
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
      const uint64_t ad0 = make_desc(base, 128, 2), bd0 = make_desc(base + 32 * 1024, 128, 2);
      for (int i = 0; i < a.iters; ++i) {
        const uint32_t off = ((uint32_t)(i & 3) * (uint32_t)a.stride_bytes) >> 4;
        const uint32_t d = tmem + (uint32_t)(i % a.nacc) * (uint32_t)a.n;
        if constexpr (CG == 2) umma_f16_cg2(d, ad0 + off, bd0 + off, idesc, i >= a.nacc ? 1u : 0u);
        else umma_bf16(d, ad0 + off, bd0 + off, idesc, i >= a.nacc ? 1u : 0u);
      }
      if constexpr (CG == 2) umma_commit_cg2(smem_u32(&bar), (uint16_t)1);
      else umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
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
}  // namespace

extern "C" int cpb200_probe_mma(int n, int cg, int iters, int nacc, int stride_bytes, void *stream) {
  if (n < 16 || n > 256 || n % 16 || (cg != 1 && cg != 2) || nacc < 1 || nacc * n > 512 || iters < 1)
    return cpb::fail(CPB200_ERR_ARG, "probe_mma: bad arguments");
  MmaProbeArgs a{n, iters, nacc, stride_bytes};
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
