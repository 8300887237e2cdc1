// Parameter blocks shared by the CUDA-core (conv_kernels.cu) and tcgen05 (gemm_tc.cu) implicit-GEMM kernels.
#pragma once
#include "vqs_common.cuh"

namespace vqs {

struct ConvParams {
  vqs_conv_gemm_desc d;
  int Ktot, Ntot;
  int a_vec;  // A rows are 16-byte aligned and Ktot % 4 == 0
  int cpb;    // 32-wide k-blocks per tap (tap-major A): Cred / 32
  int cred_real;   // channels that exist in X; d.Cred is the image's padded width when an operand image covers Cred % 32 != 0
  FastDiv divL, divCpb, divCred;
  // split-K (tensor-core engines, few-tile GEMMs with a bias-only epilogue): grid.z CTAs take kt_per_split k-blocks each
  // and write raw accumulators to partial[z][M][Ntot]; conv_splitk_epilogue_kernel sums them and adds the bias
  int splits, kt_per_split;
  float* partial;
};

struct WgradParams {
  vqs_wgrad_desc d;
  int Nw, Kred, splits, kt_per_split;
  float* partial;  // [splits][M*Nw] or NULL (direct)
  FastDiv divLa;
};


// tcgen05 launchers (gemm_tc.cu).  precision: 1 = single-pass TF32, 2 = 3xTF32 (fp32-accurate split).
int launch_conv_tc(const ConvParams& p, int precision, cudaStream_t st);
int launch_wgrad_tc(const WgradParams& p, int precision, cudaStream_t st);
bool conv_tc_supported(const ConvParams& p);
bool wgrad_tc_supported(const WgradParams& p);
bool wgrad_tc_pairs(const WgradParams& p);   // the tensor-core wgrad of this shape runs as CTA pairs (plan its split accordingly)
// TMA-fed wgrad (wgrad_tma.cu): k-blocks are 32 positions of one utterance; plan the split over wgrad_tma_kblocks(p)
bool wgrad_tma_supported(const WgradParams& p);
int wgrad_tma_kblocks(const WgradParams& p);
int launch_wgrad_tma(const WgradParams& p, float* workspace, cudaStream_t st);

}  // namespace vqs
