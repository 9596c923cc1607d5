// Launchers of the tcgen05 streaming path (vcb_umma.cuh).
#define VCB_UMMA_KERNELS
#include "vcb_umma.cuh"

namespace vcb {

cudaError_t vcb_launch_umma_tables(const umma::TableParams& tp, cudaStream_t st) {
  if (tp.n_chunks <= 0) return cudaSuccess;
  umma::vcb_umma_tables_kernel<<<(unsigned)tp.n_chunks, 64, 0, st>>>(tp);
  return cudaGetLastError();
}

cudaError_t vcb_launch_umma_stream(const umma::Params& sp, int n_tiles, int n_split, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(umma::vcb_umma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, umma::SM_TOTAL);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  umma::vcb_umma_stream_kernel<<<dim3((unsigned)n_tiles, (unsigned)n_split), umma::NTHR, umma::SM_TOTAL, st>>>(sp);
  return cudaGetLastError();
}

}  // namespace vcb
