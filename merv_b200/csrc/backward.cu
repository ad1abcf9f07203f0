// Backward of the module-by-module path: the gradients MERV's training step needs through the
// projectors and the fusion adapter (both trainable in every stage, merv/models/vidlms/merv.py:318-320,342-343,363-365).
// The backbones are frozen (merv.py:316,339,361), so no gradient flows into the patch features.
//
//   LinearProjector      y = P W^T + b           dW = dY^T P (tcgen05 GEMM on transposed copies), db = colsum(dY)
//   adapter (nn_utils.py:487-521, restated as  w = softmax_e(u . mean_t V_e),  out = sum_e w_e V_e):
//        dw_e   = <dOut, V_e>                     ds = softmax'(w) dw
//        dV_e   = w_e dOut + (ds_e / T) u         du = sum_{b,e} ds_e mean_t V_e
//        u = Wk^T q / sqrt(E), q = Wq Q^T + b_q:  dq = Wk du / sqrt(E), dWk = q du^T / sqrt(E), dWq = dq Q, db_q = dq, dQ = Wq^T dq
//   v_proj_weight / out_proj / b_k / b_v receive no (zero) gradient, exactly as in the reference where the attention
//   output is discarded and b_k cancels in the softmax.
// All reductions are fixed-order (deterministic).
#include "common.cuh"

namespace merv {

template <int kThreads>
__device__ __forceinline__ float bwd_block_sum(float v, float* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float r = (lane < kThreads / 32) ? smem[lane] : 0.f;
  return warp_sum(r);
}

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- y[c, r] = x[r, c] : 32 x 32 tiles through shared memory, coalesced both ways ------------------------
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ x, T* __restrict__ y, int R, int C, long long ldx, long long ldy) {
  __shared__ T tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i, c = c0 + tx;
    if (r < R && c < C) tile[ty + i][tx] = x[(long long)r * ldx + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, r = r0 + tx;
    if (r < R && c < C) y[(long long)c * ldy + r] = tile[tx][ty + i];
  }
}

// ---- out[n] = sum_m x[m, n]  (bias gradient): two fixed-order stages ---------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T* __restrict__ x, float* __restrict__ partial, int M, int N, long long ld, int rows_per_block) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  if (n >= N) return;
  float acc = 0.f;
  for (int m = m0; m < m1; ++m) acc += to_float(x[(long long)m * ld + n]);
  partial[(long long)blockIdx.y * N + n] = acc;
}
template <typename T>
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, T* __restrict__ out, int N, int nblocks) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  float acc = 0.f;
  for (int i = 0; i < nblocks; ++i) acc += partial[(long long)i * N + n];
  out[n] = from_float<T>(acc);
}

// ---- adapter backward, stage 1: dw[b,e] = <dOut[b], V_e[b]> and vbar[b,e,:] = mean_t V_e[b,t,:] ------------
struct BwdPtrs {
  const void* V[MERV_MAX_ENCODERS];
  void* dV[MERV_MAX_ENCODERS];
};

// grid (kblocks, E, B), 128 threads; thread owns one 16-byte vector of k and walks all T tokens
template <typename T>
__global__ void __launch_bounds__(128) mix_bwd_reduce_kernel(const __grid_constant__ BwdPtrs p, const T* __restrict__ dOut, float* __restrict__ dw_partial,
                                                             float* __restrict__ vbar, int E, int Ttok, int K) {
  constexpr int VEC = Vec16<T>::kN;
  __shared__ float red[32];
  const int kb = blockIdx.x, e = blockIdx.y, b = blockIdx.z;
  const int k0 = (kb * 128 + threadIdx.x) * VEC;
  float vs[VEC], dot = 0.f;
#pragma unroll
  for (int c = 0; c < VEC; ++c) vs[c] = 0.f;
  if (k0 < K) {
    const T* v = static_cast<const T*>(p.V[e]) + (long long)b * Ttok * K + k0;
    const T* g = dOut + (long long)b * Ttok * K + k0;
    for (int t = 0; t < Ttok; ++t) {
      float a[VEC], d[VEC];
      Vec16<T>::unpack(ldg_nc_v4(v + (long long)t * K), a);
      Vec16<T>::unpack(ldg_nc_v4(g + (long long)t * K), d);
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        vs[c] += a[c];
        dot = fmaf(a[c], d[c], dot);
      }
    }
    float* vb = vbar + ((long long)b * E + e) * K + k0;
    const float inv = 1.0f / float(Ttok);
#pragma unroll
    for (int c = 0; c < VEC; ++c) vb[c] = vs[c] * inv;
  }
  dot = bwd_block_sum<128>(dot, red);
  if (threadIdx.x == 0) dw_partial[((long long)b * E + e) * gridDim.x + kb] = dot;
}

// ---- stage 2 (one warp per video): ds = softmax backward ------------------------------------------------------
__global__ void __launch_bounds__(32) mix_bwd_scores_kernel(const float* __restrict__ dw_partial, const float* __restrict__ weights,
                                                            const float* __restrict__ dweights_out, float* __restrict__ ds, int E, int kblocks) {
  const int b = blockIdx.x, lane = threadIdx.x;
  float dw = 0.f, w = 0.f;
  if (lane < E) {
    for (int i = 0; i < kblocks; ++i) dw += dw_partial[((long long)b * E + lane) * kblocks + i];
    if (dweights_out != nullptr) dw += dweights_out[(long long)b * E + lane];
    w = weights[(long long)b * E + lane];
  }
  const float inner = warp_sum(w * dw);
  if (lane < E) ds[(long long)b * E + lane] = w * (dw - inner);
}

// ---- stage 3: dV_e[b,t,:] = w[b,e] dOut[b,t,:] + (ds[b,e] / T) u -----------------------------------------------
template <typename T, int E>
__global__ void __launch_bounds__(256) mix_bwd_dv_kernel(const __grid_constant__ BwdPtrs p, const T* __restrict__ dOut, const float* __restrict__ weights,
                                                         const float* __restrict__ ds, const float* __restrict__ u, int Ttok, int K) {
  constexpr int VEC = Vec16<T>::kN;
  const int b = blockIdx.y;
  float w[E], s[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    w[e] = weights[(long long)b * E + e];
    s[e] = ds[(long long)b * E + e] / float(Ttok);
  }
  const int kvec = K / VEC;
  const long long nvec = (long long)Ttok * kvec;
  const T* g = dOut + (long long)b * Ttok * K;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (long long)gridDim.x * 256) {
    const int kc = int(i % kvec) * VEC;
    float d[VEC], uu[VEC];
    Vec16<T>::unpack(ldg_nc_v4(g + i * VEC), d);
#pragma unroll
    for (int c = 0; c < VEC; c += 4) {
      const float4 u4 = __ldg(reinterpret_cast<const float4*>(u + kc + c));
      uu[c] = u4.x; uu[c + 1] = u4.y; uu[c + 2] = u4.z; uu[c + 3] = u4.w;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float o[VEC];
#pragma unroll
      for (int c = 0; c < VEC; ++c) o[c] = fmaf(w[e], d[c], s[e] * uu[c]);
      stg_na_v4(static_cast<T*>(p.dV[e]) + (long long)b * Ttok * K + i * VEC, Vec16<T>::pack(o));
    }
  }
}

// ---- query-vector backward ----------------------------------------------------------------------------------------
// du[k] = sum_{b,e} ds[b,e] vbar[b,e,k]
__global__ void __launch_bounds__(256) du_kernel(const float* __restrict__ ds, const float* __restrict__ vbar, float* __restrict__ du, int BE, int K) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= K) return;
  float acc = 0.f;
  for (int i = 0; i < BE; ++i) acc = fmaf(ds[i], vbar[(long long)i * K + k], acc);
  du[k] = acc;
}
// out[d] = scale * sum_k W[d,k] x[k] (+ bias[d]) : one warp per row
template <typename T, typename X>
__global__ void __launch_bounds__(256) gemv_n_kernel(const T* __restrict__ W, const X* __restrict__ x, const T* __restrict__ bias, float* __restrict__ out,
                                                     int D, int K, float scale) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= D) return;
  const int lane = threadIdx.x & 31;
  const T* w = W + (long long)row * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc += to_float(w[k]) * to_float(x[k]);
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc * scale + (bias ? to_float(bias[row]) : 0.f);
}
// out[j] = sum_d W[d,j] x[d]  (W [D, J] row-major), fp32 x
template <typename T>
__global__ void __launch_bounds__(256) gemv_tf_kernel(const T* __restrict__ W, const float* __restrict__ x, T* __restrict__ out, int D, int J) {
  __shared__ float part[8][33];
  const int jx = threadIdx.x & 31, dy = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + jx;
  float acc = 0.f;
  if (j < J)
    for (int d = dy; d < D; d += 8) acc += to_float(W[(long long)d * J + j]) * x[d];
  part[dy][jx] = acc;
  __syncthreads();
  if (dy == 0 && j < J) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][jx];
    out[j] = from_float<T>(s);
  }
}
// out[i, j] = scale * a[i] * b[j]
template <typename T, typename Bv>
__global__ void __launch_bounds__(256) outer_kernel(const float* __restrict__ a, const Bv* __restrict__ b, T* __restrict__ out, int I, int J, float scale) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)I * J) return;
  const int i = int(idx / J), j = int(idx % J);
  out[idx] = from_float<T>(scale * a[i] * to_float(b[j]));
}
template <typename T>
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ dq, T* __restrict__ dbias, int embed) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < 3 * embed) dbias[i] = from_float<T>(i < embed ? dq[i] : 0.f);  // b_k cancels in the softmax, b_v is dead
}

// ---- elementwise erf-GELU forward / backward for the MLP projectors' training path (nn.GELU(), nn_utils.py:48) --------
// y = gelu(z);  dz = dy * gelu'(z),  gelu'(z) = 0.5 (1 + erf(z / sqrt 2)) + z exp(-z^2 / 2) / sqrt(2 pi)
template <typename T, bool kBackward>
__global__ void __launch_bounds__(256) gelu_kernel(const T* __restrict__ z, const T* __restrict__ dy, T* __restrict__ out, long long nvec) {
  constexpr int VEC = Vec16<T>::kN;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (long long)gridDim.x * 256) {
    float a[VEC], g[VEC], o[VEC];
    Vec16<T>::unpack(ldg_nc_v4(z + i * VEC), a);
    if (kBackward) Vec16<T>::unpack(ldg_nc_v4(dy + i * VEC), g);
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      const float cdf = 0.5f * (1.0f + erff(a[c] * 0.70710678118654752440f));
      if (kBackward)
        o[c] = g[c] * (cdf + a[c] * 0.3989422804014327f * expf(-0.5f * a[c] * a[c]));
      else
        o[c] = a[c] * cdf;
    }
    stg_na_v4(out + i * VEC, Vec16<T>::pack(o));
  }
}

template <typename T>
static int launch_mix_dv(const BwdPtrs& p, const void* dOut, const float* weights, const float* ds, const float* u, int B, int E, int Ttok, int K,
                         cudaStream_t s) {
  constexpr int VEC = Vec16<T>::kN;
  const long long nvec = (long long)Ttok * (K / VEC);
  long long gx = (nvec + 255) / 256;
  const long long cap = (long long)sm_count() * 16 / (B > 0 ? B : 1) + 1;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, B);
  const T* g = static_cast<const T*>(dOut);
  switch (E) {
#define MERV_DV_CASE(n) case n: mix_bwd_dv_kernel<T, n><<<grid, 256, 0, s>>>(p, g, weights, ds, u, Ttok, K); break;
    MERV_DV_CASE(1) MERV_DV_CASE(2) MERV_DV_CASE(3) MERV_DV_CASE(4) MERV_DV_CASE(5) MERV_DV_CASE(6) MERV_DV_CASE(7) MERV_DV_CASE(8)
#undef MERV_DV_CASE
    default: return fail(MERV_E_ARG, "merv_mix_backward: E=%d", E);
  }
  MERV_CUDA_OK(cudaGetLastError());
  return MERV_OK;
}

}  // namespace merv

using namespace merv;

extern "C" int merv_transpose(const void* x, void* y, int R, int C, int64_t ldx, int64_t ldy, int dtype, void* stream) {
  MERV_REQUIRE(dtype == MERV_F32 || dtype == MERV_BF16, MERV_E_DTYPE, "merv_transpose: unknown dtype %d", dtype);
  MERV_REQUIRE(x && y, MERV_E_ARG, "merv_transpose: NULL pointer");
  MERV_REQUIRE(R >= 0 && C >= 0 && ldx >= C && ldy >= R, MERV_E_SHAPE, "merv_transpose: R=%d C=%d ldx=%lld ldy=%lld", R, C, (long long)ldx, (long long)ldy);
  if (int rc = require_sm100()) return rc;
  if (R == 0 || C == 0) return MERV_OK;
  dim3 grid((C + 31) / 32, (R + 31) / 32);
  MERV_REQUIRE(grid.y <= 65535, MERV_E_SHAPE, "merv_transpose: R=%d too large", R);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == MERV_BF16)
    transpose_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, R, C, ldx, ldy);
  else
    transpose_kernel<float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, R, C, ldx, ldy);
  MERV_CUDA_OK(cudaGetLastError());
  return MERV_OK;
}

extern "C" size_t merv_colsum_workspace(int M, int N) { return (size_t)((M + 255) / 256) * (size_t)(N > 0 ? N : 0); }

extern "C" int merv_colsum(const void* x, void* out, float* workspace, int M, int N, int64_t ld, int dtype, void* stream) {
  MERV_REQUIRE(dtype == MERV_F32 || dtype == MERV_BF16, MERV_E_DTYPE, "merv_colsum: unknown dtype %d", dtype);
  MERV_REQUIRE(x && out && workspace, MERV_E_ARG, "merv_colsum: NULL pointer");
  MERV_REQUIRE(M > 0 && N > 0 && ld >= N, MERV_E_SHAPE, "merv_colsum: M=%d N=%d ld=%lld", M, N, (long long)ld);
  if (int rc = require_sm100()) return rc;
  const int nblocks = (M + 255) / 256;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid((N + 255) / 256, nblocks);
  if (dtype == MERV_BF16) {
    colsum_partial_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, workspace, M, N, ld, 256);
    colsum_final_kernel<__nv_bfloat16><<<(N + 255) / 256, 256, 0, s>>>(workspace, (__nv_bfloat16*)out, N, nblocks);
  } else {
    colsum_partial_kernel<float><<<grid, 256, 0, s>>>((const float*)x, workspace, M, N, ld, 256);
    colsum_final_kernel<float><<<(N + 255) / 256, 256, 0, s>>>(workspace, (float*)out, N, nblocks);
  }
  MERV_CUDA_OK(cudaGetLastError());
  return MERV_OK;
}

extern "C" size_t merv_mix_backward_workspace(int B, int E, int T, int K, int embed) {
  (void)T;
  if (B <= 0 || E <= 0 || K <= 0) return 0;
  const size_t kblocks = (size_t)(K + 128 * 4 - 1) / (128 * 4);  // upper bound (fp32 vectors are the narrower ones)
  return (size_t)B * E * kblocks + (size_t)B * E * K + (size_t)B * E + (size_t)K + 2 * (size_t)embed;
}

// V_e [B, T, K] (all encoders must have T tokens), dOut [B, T, K], weights fp32 [B, E] (forward output), u fp32 [K];
// dweights_out (optional) fp32 [B, E]: gradient w.r.t. the returned weights.  Outputs: dV_e [B, T, K] (dtype) and the parameter
// gradients dQ [embed], dWq [embed, embed], dWk [embed, K], dbias [3 * embed] in `dtype`.
extern "C" int merv_mix_backward(const void* const* V, const void* dOut, const float* weights, const float* dweights_out, const float* u,
                                 const void* Q, const void* Wq, const void* Wk, const void* in_proj_bias, void* const* dV, void* dQ, void* dWq,
                                 void* dWk, void* dbias, float* workspace, size_t workspace_floats, int B, int E, int T, int K, int embed,
                                 int dtype, void* stream) {
  MERV_REQUIRE(dtype == MERV_F32 || dtype == MERV_BF16, MERV_E_DTYPE, "merv_mix_backward: unknown dtype %d", dtype);
  MERV_REQUIRE(V && dOut && weights && u && Q && Wq && Wk && dV && dQ && dWq && dWk && dbias && workspace, MERV_E_ARG, "merv_mix_backward: NULL pointer");
  MERV_REQUIRE(E >= 1 && E <= MERV_MAX_ENCODERS, MERV_E_ARG, "merv_mix_backward: E=%d", E);
  MERV_REQUIRE(B > 0 && T > 0 && K > 0 && embed > 0, MERV_E_SHAPE, "merv_mix_backward: B=%d T=%d K=%d embed=%d", B, T, K, embed);
  const int vec = dtype == MERV_BF16 ? 8 : 4;
  MERV_REQUIRE(K % vec == 0, MERV_E_SHAPE, "merv_mix_backward: K=%d must be a multiple of %d", K, vec);
  MERV_REQUIRE(workspace_floats >= merv_mix_backward_workspace(B, E, T, K, embed), MERV_E_ARG, "merv_mix_backward: workspace too small");
  if (int rc = require_sm100()) return rc;
  BwdPtrs p = {};
  for (int e = 0; e < E; ++e) {
    MERV_REQUIRE(V[e] && dV[e] && aligned16(V[e]) && aligned16(dV[e]), MERV_E_ALIGN, "merv_mix_backward: encoder %d: NULL or misaligned tensor", e);
    p.V[e] = V[e];
    p.dV[e] = dV[e];
  }
  const int kblocks = (K / vec + 127) / 128;
  float* dw_partial = workspace;
  float* vbar = dw_partial + (size_t)B * E * ((K + 511) / 512);
  float* ds = vbar + (size_t)B * E * K;
  float* du = ds + (size_t)B * E;
  float* q = du + K;
  float* dq = q + embed;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const float rs = 1.0f / sqrtf(float(embed));
  const long long nWk = (long long)embed * K, nWq = (long long)embed * embed;
  if (dtype == MERV_BF16) {
    using T_ = __nv_bfloat16;
    mix_bwd_reduce_kernel<T_><<<dim3(kblocks, E, B), 128, 0, s>>>(p, (const T_*)dOut, dw_partial, vbar, E, T, K);
    mix_bwd_scores_kernel<<<B, 32, 0, s>>>(dw_partial, weights, dweights_out, ds, E, kblocks);
    if (int rc = launch_mix_dv<T_>(p, dOut, weights, ds, u, B, E, T, K, s)) return rc;
    du_kernel<<<(K + 255) / 256, 256, 0, s>>>(ds, vbar, du, B * E, K);
    gemv_n_kernel<T_, T_><<<(embed + 7) / 8, 256, 0, s>>>((const T_*)Wq, (const T_*)Q, (const T_*)in_proj_bias, q, embed, embed, 1.0f);
    gemv_n_kernel<T_, float><<<(embed + 7) / 8, 256, 0, s>>>((const T_*)Wk, du, nullptr, dq, embed, K, rs);
    outer_kernel<T_, float><<<(unsigned)((nWk + 255) / 256), 256, 0, s>>>(q, du, (T_*)dWk, embed, K, rs);
    outer_kernel<T_, T_><<<(unsigned)((nWq + 255) / 256), 256, 0, s>>>(dq, (const T_*)Q, (T_*)dWq, embed, embed, 1.0f);
    gemv_tf_kernel<T_><<<(embed + 31) / 32, 256, 0, s>>>((const T_*)Wq, dq, (T_*)dQ, embed, embed);
    bias_grad_kernel<T_><<<(3 * embed + 255) / 256, 256, 0, s>>>(dq, (T_*)dbias, embed);
  } else {
    using T_ = float;
    mix_bwd_reduce_kernel<T_><<<dim3(kblocks, E, B), 128, 0, s>>>(p, (const T_*)dOut, dw_partial, vbar, E, T, K);
    mix_bwd_scores_kernel<<<B, 32, 0, s>>>(dw_partial, weights, dweights_out, ds, E, kblocks);
    if (int rc = launch_mix_dv<T_>(p, dOut, weights, ds, u, B, E, T, K, s)) return rc;
    du_kernel<<<(K + 255) / 256, 256, 0, s>>>(ds, vbar, du, B * E, K);
    gemv_n_kernel<T_, T_><<<(embed + 7) / 8, 256, 0, s>>>((const T_*)Wq, (const T_*)Q, (const T_*)in_proj_bias, q, embed, embed, 1.0f);
    gemv_n_kernel<T_, float><<<(embed + 7) / 8, 256, 0, s>>>((const T_*)Wk, du, nullptr, dq, embed, K, rs);
    outer_kernel<T_, float><<<(unsigned)((nWk + 255) / 256), 256, 0, s>>>(q, du, (T_*)dWk, embed, K, rs);
    outer_kernel<T_, T_><<<(unsigned)((nWq + 255) / 256), 256, 0, s>>>(dq, (const T_*)Q, (T_*)dWq, embed, embed, 1.0f);
    gemv_tf_kernel<T_><<<(embed + 31) / 32, 256, 0, s>>>((const T_*)Wq, dq, (T_*)dQ, embed, embed);
    bias_grad_kernel<T_><<<(3 * embed + 255) / 256, 256, 0, s>>>(dq, (T_*)dbias, embed);
  }
  MERV_CUDA_OK(cudaGetLastError());
  return MERV_OK;
}

// out = gelu(z) (dy == NULL) or out = dy * gelu'(z); n elements, contiguous, n % (16 / sizeof) == 0
extern "C" int merv_gelu(const void* z, const void* dy, void* out, int64_t n, int dtype, void* stream) {
  MERV_REQUIRE(dtype == MERV_F32 || dtype == MERV_BF16, MERV_E_DTYPE, "merv_gelu: unknown dtype %d", dtype);
  MERV_REQUIRE(z && out, MERV_E_ARG, "merv_gelu: NULL pointer");
  const int vec = dtype == MERV_BF16 ? 8 : 4;
  MERV_REQUIRE(n >= 0 && n % vec == 0, MERV_E_SHAPE, "merv_gelu: n=%lld must be a multiple of %d", (long long)n, vec);
  MERV_REQUIRE(aligned16(z) && aligned16(out) && (dy == nullptr || aligned16(dy)), MERV_E_ALIGN, "merv_gelu: pointers must be 16-byte aligned");
  if (int rc = require_sm100()) return rc;
  if (n == 0) return MERV_OK;
  const long long nvec = n / vec;
  long long blocks = (nvec + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == MERV_BF16) {
    using T_ = __nv_bfloat16;
    if (dy) gelu_kernel<T_, true><<<(unsigned)blocks, 256, 0, s>>>((const T_*)z, (const T_*)dy, (T_*)out, nvec);
    else gelu_kernel<T_, false><<<(unsigned)blocks, 256, 0, s>>>((const T_*)z, nullptr, (T_*)out, nvec);
  } else {
    if (dy) gelu_kernel<float, true><<<(unsigned)blocks, 256, 0, s>>>((const float*)z, (const float*)dy, (float*)out, nvec);
    else gelu_kernel<float, false><<<(unsigned)blocks, 256, 0, s>>>((const float*)z, nullptr, (float*)out, nvec);
  }
  MERV_CUDA_OK(cudaGetLastError());
  return MERV_OK;
}
