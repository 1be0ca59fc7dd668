// tcgen05 / TMA tensor-core path (placeholder until the kernels land).
#include "common.cuh"
namespace cpb {
int tc_prepare_op(cpb200_op &op) { return fail(CPB200_ERR_STATE, "tensor-core path not built"); }
int tc_release_op(cpb200_op &op) { op.tc = nullptr; return CPB200_OK; }
int tc_run_op(const cpb200_op &op, cudaStream_t st) { return fail(CPB200_ERR_STATE, "tensor-core path not built"); }
}  // namespace cpb
