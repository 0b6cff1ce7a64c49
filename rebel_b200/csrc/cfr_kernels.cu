// CFR wave kernels, compiled as their own translation unit with -fmad=false: without fused multiply-adds every fp64
// operation is an individually rounded IEEE operation in the same order as the reference's scalar C++ (built without
// contraction), which makes the CFRB_STATE_F64 path reproduce the reference bit for bit between value-net calls.
#include "cfr_kernels.cuh"
#include "br_kernel.cuh"

namespace cfrb {

// Hand counts with a compile-time specialisation (1x4f, 1x5f, 1x6f, 2x3f, 2x4f); anything else takes the generic path.
#define CFRB_DISPATCH_H(H, CALL)          \
  switch (H) {                            \
    case 4: { CALL(4); break; }           \
    case 5: { CALL(5); break; }           \
    case 6: { CALL(6); break; }           \
    case 9: { CALL(9); break; }           \
    case 16: { CALL(16); break; }         \
    default: { CALL(0); break; }          \
  }

template <typename real>
cudaError_t cfr_configure(int group, int smem_bytes) {
  if (group != 32) return cudaSuccess;
  cudaError_t e = cudaSuccess;
#define CFRB_CFG(HC)                                                                                                      \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(cfr_iter_kernel<real, 32, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  CFRB_CFG(0) CFRB_CFG(4) CFRB_CFG(5) CFRB_CFG(6) CFRB_CFG(9) CFRB_CFG(16)
#undef CFRB_CFG
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(cfr_init_kernel<real, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
}

template <typename real>
void cfr_launch_init(const CfrDev<real>& p, int group, int blocks, int threads, size_t smem, cudaStream_t st, int scratch_per_group) {
  if (group == 32) cfr_init_kernel<real, 32><<<blocks, threads, smem, st>>>(p, scratch_per_group);
  else cfr_init_kernel<real, 256><<<blocks, 256, 0, st>>>(p, scratch_per_group);
}

template <typename real>
void cfr_launch_iter(const CfrDev<real>& p, int group, int blocks, int threads, size_t smem, cudaStream_t st, int iter, int do_b,
                     int do_f, int scratch_per_group) {
  if (group == 32) {
#define CFRB_CALL(HC) cfr_iter_kernel<real, 32, HC><<<blocks, threads, smem, st>>>(p, iter, do_b, do_f, scratch_per_group)
    CFRB_DISPATCH_H(p.H, CFRB_CALL)
#undef CFRB_CALL
  } else {
    cfr_iter_kernel<real, 256, 0><<<blocks, 256, 0, st>>>(p, iter, do_b, do_f, scratch_per_group);
  }
}

void br_launch(const BrDev& p, cudaStream_t st) { br_kernel<<<2, 1024, 0, st>>>(p); }

template <typename real>
cudaError_t cfr_configure_d2(int smem_bytes) {
  cudaError_t e = cudaSuccess;
#define CFRB_CFG(HC)                                                                                                      \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(cfr_iter_d2_kernel<real, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  CFRB_CFG(0) CFRB_CFG(4) CFRB_CFG(5) CFRB_CFG(6) CFRB_CFG(9) CFRB_CFG(16)
#undef CFRB_CFG
  return e;
}

// Launch with programmatic stream serialization (PDL): the grid may start before its predecessor in the stream has finished;
// the kernel orders itself with griddepcontrol.wait.  Captured into CUDA graphs as a programmatic dependency edge.
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), int blocks, int threads, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <typename real>
void cfr_launch_iter_d2(const CfrDev<real>& p, int blocks, int threads, size_t smem, cudaStream_t st, int iter, int do_b, int do_f,
                        int scratch_per_group) {
#define CFRB_CALL(HC) launch_pdl(cfr_iter_d2_kernel<real, HC>, blocks, threads, smem, st, p, iter, do_b, do_f, scratch_per_group)
  CFRB_DISPATCH_H(p.H, CFRB_CALL)
#undef CFRB_CALL
}

#define CFRB_INSTANTIATE(real)                                                                                             \
  template cudaError_t cfr_configure<real>(int, int);                                                                      \
  template void cfr_launch_init<real>(const CfrDev<real>&, int, int, int, size_t, cudaStream_t, int);                      \
  template void cfr_launch_iter<real>(const CfrDev<real>&, int, int, int, size_t, cudaStream_t, int, int, int, int);             \
  template cudaError_t cfr_configure_d2<real>(int);                                                                        \
  template void cfr_launch_iter_d2<real>(const CfrDev<real>&, int, int, size_t, cudaStream_t, int, int, int, int);
CFRB_INSTANTIATE(float)
CFRB_INSTANTIATE(double)

}  // namespace cfrb
