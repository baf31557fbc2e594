// Instantiation of the tcgen05 scan kernel for epilogue mode 2 (kModeSeed); see scan_tc_kernel.cuh.
#include "scan_tc_kernel.cuh"

namespace mcl {

cudaError_t tc_set_smem_attr_mode2() {
  cudaError_t e = cudaSuccess;
  const void* kernels[] = {(const void*)scan_tc_kernel<1, false, kModeSeed>, (const void*)scan_tc_kernel<2, false, kModeSeed>};
  for (const void* kfn : kernels)
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
  return e;
}

cudaError_t tc_launch_mode2(const cudaLaunchConfig_t* cfg, int cs, bool cap, const CUtensorMap& tm_q,
                            const CUtensorMap& tm_t, const TcParams& p) {
  (void)cap;   // ranking is by y: tanh is monotone, the seed needs no soft-cap variant
  if (cs == 2) return cudaLaunchKernelEx(cfg, scan_tc_kernel<2, false, kModeSeed>, tm_q, tm_t, p);
  return cudaLaunchKernelEx(cfg, scan_tc_kernel<1, false, kModeSeed>, tm_q, tm_t, p);
}

}  // namespace mcl
