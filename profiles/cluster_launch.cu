// Which cluster launch configurations does the B200 accept?  (debug aid for the CTA-pair GEMM launch)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o profiles/cluster_launch profiles/cluster_launch.cu
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void __launch_bounds__(576, 1) k(int* out) {
  extern __shared__ unsigned char sm[];
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  if (threadIdx.x == 0) { sm[0] = 1; atomicAdd(out + r, 1); }
}
static void attempt(dim3 grid, unsigned cx, unsigned cy, size_t smem, bool pdl, int* d) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cx; attr[0].val.clusterDim.y = cy; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 2 : 1;
  int nc = -1;
  cudaError_t eo = cudaOccupancyMaxActiveClusters(&nc, k, &cfg);
  cudaMemset(d, 0, 8);
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, d);
  cudaError_t e2 = cudaDeviceSynchronize();
  int h[2]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
  printf("grid (%u,%u,%u) cluster (%u,%u) smem %zu pdl %d: occupancy %s clusters %d | launch %s | sync %s | rank counts %d %d\n",
         grid.x, grid.y, grid.z, cx, cy, smem, (int)pdl, cudaGetErrorName(eo), nc, cudaGetErrorName(e), cudaGetErrorName(e2), h[0], h[1]);
  cudaGetLastError();
}
int main() {
  int* d; cudaMalloc(&d, 8);
  size_t smems[] = {1024, 100 * 1024, 197888, 226 * 1024};
  for (size_t s : smems) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s);
    printf("set max dyn smem %zu: %s\n", s, cudaGetErrorName(e));
    attempt(dim3(24, 6, 1), 1, 2, s, false, d);
    attempt(dim3(24, 6, 1), 1, 2, s, true, d);
    attempt(dim3(24, 6, 1), 2, 1, s, false, d);
    attempt(dim3(24, 6, 2), 1, 2, s, true, d);
  }
  return 0;
}
