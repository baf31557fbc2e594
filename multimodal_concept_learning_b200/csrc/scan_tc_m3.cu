// Instantiation of the tcgen05 scan kernel for epilogue mode 3 (kModeGrad); see scan_tc_kernel.cuh.
#include "scan_tc_kernel.cuh"

namespace mcl {

cudaError_t tc_set_smem_attr_mode3() {
  cudaError_t e = cudaSuccess;
  const void* kernels[] = {(const void*)scan_tc_kernel<1, false, kModeGrad>, (const void*)scan_tc_kernel<2, false, kModeGrad>,
                           (const void*)scan_tc_kernel<1, true, kModeGrad>, (const void*)scan_tc_kernel<2, true, kModeGrad>};
  for (const void* kfn : kernels)
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
  return e;
}

cudaError_t tc_launch_mode3(const cudaLaunchConfig_t* cfg, int cs, bool cap, const CUtensorMap& tm_q,
                            const CUtensorMap& tm_t, const TcParams& p) {
  if (cs == 2)
    return cap ? cudaLaunchKernelEx(cfg, scan_tc_kernel<2, true, kModeGrad>, tm_q, tm_t, p)
               : cudaLaunchKernelEx(cfg, scan_tc_kernel<2, false, kModeGrad>, tm_q, tm_t, p);
  return cap ? cudaLaunchKernelEx(cfg, scan_tc_kernel<1, true, kModeGrad>, tm_q, tm_t, p)
             : cudaLaunchKernelEx(cfg, scan_tc_kernel<1, false, kModeGrad>, tm_q, tm_t, p);
}

}  // namespace mcl
