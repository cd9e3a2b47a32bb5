// Epilogue description + host entry of the generic batched tcgen05 TF32 GEMM (gemm.cu).
#pragma once
#include <cuda_runtime.h>

namespace hk {

struct GemmEpi {
  float* C;
  long long ldc, strideC;
  const float* D;
  long long ldd, strideD;
  const float* alpha_vec;
  const float* beta_vec;
  float alpha, beta, diag;
  int trans_c;
  int relu;            // bit0: ReLU, bit1: round the stored value to tf32
  float* C_lo;         // optional: store C as a (hi, lo) tf32 pair (3xTF32 operands for the next GEMM)
  const float* D_lo;   // optional: D given as a (hi, lo) pair
  const float* E;      // optional: raw partial product added to the accumulator before alpha (row-major [M][N] per batch)
  long long ldE, strideE;   // layout of E; 0 = same as C (ldc / strideC)
};

// C[b] = alpha_b * (A[b].B[b] + E[b]) + diag*I + beta_b * (D[b] + D_lo[b]);  see hk_gemm_tf32 in the public header.
// Dispatches on the precision mode (host.h): one TF32 pass, or 3xTF32 over internally split operands.
int gemm_tf32(const float* A, int a_mn, long long lda, long long strideA, const float* B, int b_mn, long long ldb,
              long long strideB, const GemmEpi& epi, int M, int N, int K, int batch, cudaStream_t stream);
// Always one TF32 pass (callers that manage (hi, lo) operand pairs themselves: the Newton-Schulz chain).
int gemm_tf32_1x(const float* A, int a_mn, long long lda, long long strideA, const float* B, int b_mn, long long ldb,
                 long long strideB, const GemmEpi& epi, int M, int N, int K, int batch, cudaStream_t stream);

// One launch of the 3xTF32 product for operands already held as tf32 (hi, lo) pairs:  C = epilogue(Ah.Bh + Al.Bh + Ah.Bl).
int gemm_tf32_pair(const float* Ah, const float* Al, int a_mn, long long lda, long long strideA, const float* Bh,
                   const float* Bl, int b_mn, long long ldb, long long strideB, const GemmEpi& epi, int M, int N, int K,
                   int batch, cudaStream_t stream);

}  // namespace hk
