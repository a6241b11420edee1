// Internal declarations shared by the .cu translation units of libb200parse.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2p {

int set_error(const char* msg);   // records msg for b2p_last_error(); returns -1
void count_launch(int n = 1);     // bumps the kernel-launch counter read by b2p_launch_count()

#define B2P_CHECK_LAUNCH()                                          \
  do {                                                              \
    cudaError_t _e = cudaGetLastError();                            \
    if (_e != cudaSuccess) return b2p::set_error(cudaGetErrorString(_e)); \
    b2p::count_launch();                                            \
  } while (0)

// Descriptor of one dense contraction for gemm_launch().
struct ConvGemm {
  int mode;            // 0 = GEMM rows, 1 = conv3x3 s1 p1, 2 = conv3x3 s2 p1   (NHWC)
  int bf16;            // operand type: 0 = fp16, 1 = bf16
  const void* A;       // mode 0: [M][lda]; conv: NHWC input (channel slice allowed, lda = channels per pixel)
  long long lda;
  const void* B;       // weights [N][K] (conv: K = 9*Cin ordered (ky,kx,c))
  int M, N, K;         // mode 0 only: M rows, K reduction
  int batch, H, W, Cin;  // conv only: input geometry
  void* out;           // [pixels][ldc] fp16 or fp32
  long long ldc;
  int out_f32;
  const float* bias;   // [N] or null
  const void* res;     // residual, same dtype as out, or null
  long long ldr;
  int act;             // 0 none, 1 SiLU, 2 GELU(erf)
  int bn_max;          // 0 = auto
  int split_out;       // fp16 output written as [hi | hi | lo] (row stride ldc >= 3N): operand of a fp16x3 GEMM
};

int gemm_launch(const ConvGemm& d, cudaStream_t st);

}  // namespace b2p
