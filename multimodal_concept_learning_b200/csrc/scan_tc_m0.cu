// Instantiation of the tcgen05 scan kernel for epilogue mode 0 (kModeTopK); see scan_tc_kernel.cuh.
#include "scan_tc_kernel.cuh"

namespace mcl {

cudaError_t tc_set_smem_attr_mode0() {
  cudaError_t e = cudaSuccess;
  const void* kernels[] = {(const void*)scan_tc_kernel<1, false, kModeTopK>, (const void*)scan_tc_kernel<2, false, kModeTopK>,
                           (const void*)scan_tc_kernel<1, true, kModeTopK>, (const void*)scan_tc_kernel<2, true, kModeTopK>};
  for (const void* kfn : kernels)
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
  return e;
}

cudaError_t tc_launch_mode0(const cudaLaunchConfig_t* cfg, int cs, bool cap, const CUtensorMap& tm_q,
                            const CUtensorMap& tm_t, const TcParams& p) {
  if (cs == 2)
    return cap ? cudaLaunchKernelEx(cfg, scan_tc_kernel<2, true, kModeTopK>, tm_q, tm_t, p)
               : cudaLaunchKernelEx(cfg, scan_tc_kernel<2, false, kModeTopK>, tm_q, tm_t, p);
  return cap ? cudaLaunchKernelEx(cfg, scan_tc_kernel<1, true, kModeTopK>, tm_q, tm_t, p)
             : cudaLaunchKernelEx(cfg, scan_tc_kernel<1, false, kModeTopK>, tm_q, tm_t, p);
}

}  // namespace mcl
