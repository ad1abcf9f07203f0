// fp32 SIMT projector GEMM: Y = act(A W^T + bias), fp32 FFMA with fp32 accumulation.
//
// This is the 1e-5 parity path (BASELINE.json north_star: "within 1e-5 relative in fp32"): kind::tf32 tensor
// cores cannot meet that bound, so fp32 modules run here.  It also accepts bf16 storage (fp32 accumulate) so
// the tests can cross-check the tcgen05 kernel against an independent implementation on the GPU.  Not the
// benchmarked path.  Replaces nn.Linear(+nn.GELU) of merv/util/nn_utils.py:31-32,46-55.
#include "common.cuh"

namespace merv {

constexpr int SBM = 64, SBN = 64, SBK = 16;

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, long long lda, const T* __restrict__ W,
                                                        long long ldw, const T* __restrict__ bias, T* __restrict__ Y,
                                                        long long ldy, int M, int N, int K, int act) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Ws[SBK][SBN + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;  // 16 x 16 threads, 4 x 4 outputs each
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // each thread loads 4 elements of A and 4 of W per k-block: row = tid / 4, k = (tid % 4) * 4 .. +4
  const int lr = threadIdx.x / 4, lk = (threadIdx.x % 4) * 4;
  for (int k0 = 0; k0 < K; k0 += SBK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = k0 + lk + q;
      const int am = m0 + lr, wn = n0 + lr;
      As[lk + q][lr] = (am < M && k < K) ? to_float(A[(long long)am * lda + k]) : 0.f;
      Ws[lk + q][lr] = (wn < N && k < K) ? to_float(W[(long long)wn * ldw + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? to_float(bias[n]) : 0.f);
      if (act == MERV_ACT_GELU_ERF) v = gelu_erf(v);
      if constexpr (sizeof(T) == 2)
        Y[(long long)m * ldy + n] = __float2bfloat16_rn(v);
      else
        Y[(long long)m * ldy + n] = v;
    }
  }
}

int launch_gemm_simt(const void* A, long long lda, const void* W, long long ldw, const void* bias, void* Y, long long ldy,
                     int M, int N, int K, int act, int dtype, cudaStream_t s) {
  dim3 grid((N + SBN - 1) / SBN, (M + SBM - 1) / SBM);
  if (dtype == MERV_BF16) {
    using T = __nv_bfloat16;
    gemm_simt_kernel<T><<<grid, 256, 0, s>>>((const T*)A, lda, (const T*)W, ldw, (const T*)bias, (T*)Y, ldy, M, N, K, act);
  } else {
    using T = float;
    gemm_simt_kernel<T><<<grid, 256, 0, s>>>((const T*)A, lda, (const T*)W, ldw, (const T*)bias, (T*)Y, ldy, M, N, K, act);
  }
  MERV_CUDA_OK(cudaGetLastError());
  return MERV_OK;
}

}  // namespace merv
