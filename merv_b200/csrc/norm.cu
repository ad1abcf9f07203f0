// LayerNorm over the channel dimension of (possibly segmented) rows — the two places the reference puts one on this path:
//   * nn.LayerNorm(vision_dim) in front of a projector (`pre_proj_layernorm=True`, merv/util/nn_utils.py:26-29,41-44,67-70,
//     91-94; constructed at merv/models/vidlms/merv.py:165-171);
//   * nn.LayerNorm(E * llm_dim) of feature_fusion == "concat_channel_ln" (merv.py:219-223), applied to the channel-wise
//     concatenation of the E projected token tensors (merv.py:603-606).  The kernel takes the E tensors as row SEGMENTS, so the
//     un-normalised concatenation is never written: every input element is read from HBM once, every output written once.
// Arithmetic follows ATen: fp32 statistics whatever the storage dtype, biased variance, y = (x - mean) * rstd * gamma + beta;
// two-pass (mean, then centred sum of squares) over values held in registers, so there is no E[x^2] - E[x]^2 cancellation.
// HBM-bound: algorithmic bytes per row = (in + out) * sum K_s * sizeof(T).
#include <cstdlib>

#include "common.cuh"

namespace merv {

struct LnSegments {
  const void* x[MERV_MAX_ENCODERS];
  long long ld[MERV_MAX_ENCODERS];
  int vend[MERV_MAX_ENCODERS];  // cumulative end of segment s in 16-byte vectors
  int nseg;
};

template <int kTPR>
__device__ __forceinline__ float row_sum(float v, float* red) {
  if constexpr (kTPR == 32) {
    return warp_sum(v);
  } else {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = (lane < kTPR / 32) ? red[lane] : 0.f;
    return warp_sum(r);
  }
}

// address of 16-byte vector i of row m (i counts vectors over the concatenated row)
template <typename T>
__device__ __forceinline__ const T* seg_vec_ptr(const LnSegments& p, long long m, int i) {
  constexpr int VEC = Vec16<T>::kN;
  int s = 0, v0 = 0;
#pragma unroll
  for (int k = 0; k < MERV_MAX_ENCODERS - 1; ++k) {
    if (k + 1 < p.nseg && i >= p.vend[k]) {
      s = k + 1;
      v0 = p.vend[k];
    }
  }
  return static_cast<const T*>(p.x[s]) + m * p.ld[s] + (long long)(i - v0) * VEC;
}

// the same split into (address in row 0, row stride): a thread's columns are fixed, so the forward kernel resolves the segment
// of each of its vectors once and pays one multiply-add per vector and row afterwards
template <typename T>
__device__ __forceinline__ void seg_col(const LnSegments& p, int i, const T*& col, int& ld) {
  constexpr int VEC = Vec16<T>::kN;
  const T* base = static_cast<const T*>(p.x[0]);
  long long l = p.ld[0];
  int v0 = 0;
#pragma unroll
  for (int k = 0; k < MERV_MAX_ENCODERS - 1; ++k) {
    if (k + 1 < p.nseg && i >= p.vend[k]) {
      base = static_cast<const T*>(p.x[k + 1]);
      l = p.ld[k + 1];
      v0 = p.vend[k];
    }
  }
  col = base + (long long)(i - v0) * VEC;
  ld = (int)l;
}

template <typename T> __device__ __forceinline__ uint4 vec_of_ones();
template <> __device__ __forceinline__ uint4 vec_of_ones<__nv_bfloat16>() { return make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u); }
template <> __device__ __forceinline__ uint4 vec_of_ones<float>() { return make_uint4(0x3F800000u, 0x3F800000u, 0x3F800000u, 0x3F800000u); }

// kTPR threads per row: 32 = one warp per row (8 rows per CTA), >= 256 = one CTA per row.  Persistent: a row group walks rows
// `first, first + stride, ...` and issues the loads of its NEXT row before it reduces and stores the current one.  A thread owns
// kMaxVec fixed 16-byte columns of the row (kMaxVec is trimmed to the row width by the dispatcher); the current row lives in
// registers as packed fp32 pairs, gamma / beta are staged in shared memory once per CTA (re-reading them per row would triple
// the L2 -> SM traffic of wide rows: they are as long as the row).
template <typename T, int kTPR, int kMaxVec>
__global__ void __launch_bounds__(kTPR > 256 ? kTPR : 256) layernorm_kernel(const __grid_constant__ LnSegments p, const T* __restrict__ gamma,
                                                                            const T* __restrict__ beta, T* __restrict__ y, long long ldy,
                                                                            int M, int nvec, float inv_n, float eps) {
  constexpr int VEC = Vec16<T>::kN;
  constexpr int P = Pairs<T>::kP;
  constexpr int kThreads = kTPR > 256 ? kTPR : 256;
  constexpr int kRows = kThreads / kTPR;
  typedef unsigned long long u64;
  __shared__ float red[32];
  extern __shared__ uint4 ln_gb[];  // gamma [nvec] then beta [nvec] (ones / zeros when absent)
  const int tr = threadIdx.x % kTPR;
  const int rg = threadIdx.x / kTPR;
  const long long stride = (long long)gridDim.x * kRows;
  const T* col[kMaxVec];
  int cld[kMaxVec];
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    col[j] = nullptr;
    cld[j] = 0;
    if (tr + j * kTPR < nvec) seg_col<T>(p, tr + j * kTPR, col[j], cld[j]);
  }
  uint4 nxt[kMaxVec];
  {
    const long long m0 = (long long)blockIdx.x * kRows + rg;
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      nxt[j] = make_uint4(0u, 0u, 0u, 0u);
      if (m0 < M && col[j] != nullptr) nxt[j] = ldg_nc_v4(col[j] + (long long)(int)m0 * (long long)cld[j]);
    }
  }
  for (int i = threadIdx.x; i < nvec; i += kThreads) {
    ln_gb[i] = gamma != nullptr ? ldg_v4(gamma + (long long)i * VEC) : vec_of_ones<T>();
    ln_gb[nvec + i] = beta != nullptr ? ldg_v4(beta + (long long)i * VEC) : make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const u64 one2 = f32x2_pack(1.0f, 1.0f);
  // `base` is CTA-uniform, so every thread of the CTA runs the same number of iterations (the reductions contain barriers)
  for (long long base = (long long)blockIdx.x * kRows; base < M; base += stride) {
    const long long m = base + rg;
    const bool live = m < M;
    u64 x[kMaxVec][P];
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) Pairs<T>::unpack(nxt[j], x[j]);  // vectors past the row end / rows past M are zero
    {
      const long long mn = m + stride;
#pragma unroll
      for (int j = 0; j < kMaxVec; ++j) {
        nxt[j] = make_uint4(0u, 0u, 0u, 0u);
        if (mn < M && col[j] != nullptr) nxt[j] = ldg_nc_v4(col[j] + (long long)(int)mn * (long long)cld[j]);
      }
    }
    u64 acc = f32x2_pack(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
#pragma unroll
      for (int q = 0; q < P; ++q) acc = f32x2_fma(x[j][q], one2, acc);
    }
    float s0, s1;
    f32x2_unpack(acc, s0, s1);
    const float mean = row_sum<kTPR>(s0 + s1, red) * inv_n;
    const u64 neg_mean2 = f32x2_pack(-mean, -mean);
    acc = f32x2_pack(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      if (col[j] != nullptr) {
#pragma unroll
        for (int q = 0; q < P; ++q) {
          const u64 d = f32x2_fma(x[j][q], one2, neg_mean2);
          acc = f32x2_fma(d, d, acc);
        }
      }
    }
    f32x2_unpack(acc, s0, s1);
    const float rstd = 1.0f / sqrtf(row_sum<kTPR>(s0 + s1, red) * inv_n + eps);
    const u64 rstd2 = f32x2_pack(rstd, rstd);
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int i = tr + j * kTPR;
      if (live && col[j] != nullptr) {
        u64 g[P], b[P], o[P];
        Pairs<T>::unpack(ln_gb[i], g);
        Pairs<T>::unpack(ln_gb[nvec + i], b);
#pragma unroll
        for (int q = 0; q < P; ++q) o[q] = f32x2_fma(f32x2_mul(f32x2_fma(x[j][q], one2, neg_mean2), rstd2), g[q], b[q]);
        stg_na_v4(y + m * ldy + (long long)i * VEC, Pairs<T>::pack(o));
      }
    }
  }
}

// Backward: with xhat = (x - mean) * rstd and g = dy * gamma,
//   dx = rstd * (g - mean_k(g) - xhat * mean_k(g * xhat)),   dgamma = sum_m dy * xhat,   dbeta = sum_m dy.
// The kernel writes dx (optional) and gx = dy * xhat; the two column sums are merv_colsum over gx and dy (fixed order).
template <typename T, int kTPR, int kMaxVec>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const __grid_constant__ LnSegments p, const T* __restrict__ dy, long long lddy,
                                                            const T* __restrict__ gamma, T* __restrict__ dx, long long lddx,
                                                            T* __restrict__ gx, long long ldgx, int M, int nvec, float inv_n, float eps) {
  constexpr int VEC = Vec16<T>::kN;
  constexpr int kRows = 256 / kTPR;
  __shared__ float red[32];
  const int tr = threadIdx.x % kTPR;
  const long long m = (long long)blockIdx.x * kRows + threadIdx.x / kTPR;
  const bool live = m < M;
  uint4 v[kMaxVec];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = tr + j * kTPR;
    v[j] = make_uint4(0u, 0u, 0u, 0u);
    if (live && i < nvec) {
      v[j] = ldg_nc_v4(seg_vec_ptr<T>(p, m, i));
      float f[VEC];
      Vec16<T>::unpack(v[j], f);
#pragma unroll
      for (int c = 0; c < VEC; ++c) sum += f[c];
    }
  }
  const float mean = row_sum<kTPR>(sum, red) * inv_n;
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = tr + j * kTPR;
    if (live && i < nvec) {
      float f[VEC];
      Vec16<T>::unpack(v[j], f);
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const float d = f[c] - mean;
        sq = fmaf(d, d, sq);
      }
    }
  }
  const float rstd = 1.0f / sqrtf(row_sum<kTPR>(sq, red) * inv_n + eps);
  // second sweep: the two row means of g and g * xhat (dy is re-read in the third sweep; it is L1/L2 resident by then)
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = tr + j * kTPR;
    if (live && i < nvec) {
      float f[VEC], d[VEC], g[VEC];
      Vec16<T>::unpack(v[j], f);
      Vec16<T>::unpack(ldg_v4(dy + m * lddy + (long long)i * VEC), d);
      if (gamma != nullptr) {
        Vec16<T>::unpack(ldg_v4(gamma + (long long)i * VEC), g);
      } else {
#pragma unroll
        for (int c = 0; c < VEC; ++c) g[c] = 1.f;
      }
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const float gg = d[c] * g[c];
        sg += gg;
        sgx = fmaf(gg, (f[c] - mean) * rstd, sgx);
      }
    }
  }
  const float mg = row_sum<kTPR>(sg, red) * inv_n;
  const float mgx = row_sum<kTPR>(sgx, red) * inv_n;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = tr + j * kTPR;
    if (live && i < nvec) {
      float f[VEC], d[VEC], g[VEC], o[VEC], h[VEC];
      Vec16<T>::unpack(v[j], f);
      Vec16<T>::unpack(ldg_v4(dy + m * lddy + (long long)i * VEC), d);
      if (gamma != nullptr) {
        Vec16<T>::unpack(ldg_v4(gamma + (long long)i * VEC), g);
      } else {
#pragma unroll
        for (int c = 0; c < VEC; ++c) g[c] = 1.f;
      }
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const float xh = (f[c] - mean) * rstd;
        h[c] = d[c] * xh;
        o[c] = rstd * (d[c] * g[c] - mg - xh * mgx);
      }
      if (gx != nullptr) stg_na_v4(gx + m * ldgx + (long long)i * VEC, Vec16<T>::pack(h));
      if (dx != nullptr) stg_na_v4(dx + m * lddx + (long long)i * VEC, Vec16<T>::pack(o));
    }
  }
}

static int fill_segments(const char* fn, const void* const* X, const int64_t* ldx, const int32_t* K, int nseg, int vec, LnSegments* p,
                         int* total) {
  MERV_REQUIRE(X && ldx && K, MERV_E_ARG, "%s: NULL pointer", fn);
  MERV_REQUIRE(nseg >= 1 && nseg <= MERV_MAX_ENCODERS, MERV_E_ARG, "%s: nseg=%d not in [1,%d]", fn, nseg, MERV_MAX_ENCODERS);
  int n = 0;
  *p = LnSegments{};
  for (int s = 0; s < nseg; ++s) {
    MERV_REQUIRE(K[s] > 0 && K[s] % vec == 0, MERV_E_SHAPE, "%s: segment %d has %d channels, need a positive multiple of %d", fn, s, K[s], vec);
    MERV_REQUIRE(X[s] && aligned16(X[s]) && ldx[s] >= K[s] && ldx[s] % vec == 0 && ldx[s] <= 0x7fffffffLL, MERV_E_ALIGN,
                 "%s: segment %d: NULL / misaligned pointer or row stride %lld", fn, s, (long long)ldx[s]);
    p->x[s] = X[s];
    p->ld[s] = ldx[s];
    n += K[s];
    p->vend[s] = n / vec;
  }
  p->nseg = nseg;
  *total = n;
  return MERV_OK;
}

}  // namespace merv

using namespace merv;

// backward: warp-per-row up to 8 vectors per lane, CTA-per-row up to 16 vectors per thread
#define MERV_LN_DISPATCH(KERNEL, T, ...)                                                                   \
  do {                                                                                                     \
    if (nvec <= 32 * 8)                                                                                    \
      KERNEL<T, 32, 8><<<(unsigned)((M + 7) / 8), 256, 0, s>>>(__VA_ARGS__);                                \
    else                                                                                                   \
      KERNEL<T, 256, 16><<<(unsigned)M, 256, 0, s>>>(__VA_ARGS__);                                          \
  } while (0)

// forward: persistent grid (resident CTAs x SMs), threads per row and vectors per thread trimmed to the row width
template <typename T, int kTPR, int kMaxVec>
static void launch_ln(cudaStream_t s, const LnSegments& p, const void* gamma, const void* beta, void* Y, long long ldy, int M, int nvec,
                      float inv_n, float eps) {
  constexpr int kThreads = kTPR > 256 ? kTPR : 256;
  constexpr int kRows = kThreads / kTPR;
  const size_t smem = (size_t)2 * nvec * sizeof(uint4);  // gamma + beta, <= 128 KB
  static bool opted_in[64] = {};  // per instantiation and device (function attributes are per device)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (!opted_in[dev]) {
    cudaFuncSetAttribute(layernorm_kernel<T, kTPR, kMaxVec>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * (int)sizeof(uint4));
    opted_in[dev] = true;
  }
  int resident = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, layernorm_kernel<T, kTPR, kMaxVec>, kThreads, smem) != cudaSuccess || resident < 1)
    resident = 1;
  const long long ctas = ((long long)M + kRows - 1) / kRows;
  const long long cap = (long long)sm_count() * resident;
  layernorm_kernel<T, kTPR, kMaxVec><<<(unsigned)(ctas < cap ? ctas : cap), kThreads, smem, s>>>(p, (const T*)gamma, (const T*)beta, (T*)Y, ldy,
                                                                                           M, nvec, inv_n, eps);
}

template <typename T, int kTPR>
static int launch_ln_tpr(cudaStream_t s, const LnSegments& p, const void* gamma, const void* beta, void* Y, long long ldy, int M, int nvec,
                         float inv_n, float eps) {
  const int per = (nvec + kTPR - 1) / kTPR;  // 16-byte vectors per thread
  constexpr int kMaxPer = kTPR >= 1024 ? 2 : kTPR >= 512 ? 4 : kTPR >= 256 ? 16 : 8;  // what fits the register budget of the CTA size
  if (per > kMaxPer) return 1;
  if (per <= 1) launch_ln<T, kTPR, 1>(s, p, gamma, beta, Y, ldy, M, nvec, inv_n, eps);
  else if (per <= 2) launch_ln<T, kTPR, 2>(s, p, gamma, beta, Y, ldy, M, nvec, inv_n, eps);
  else if (per <= 4) launch_ln<T, kTPR, 4>(s, p, gamma, beta, Y, ldy, M, nvec, inv_n, eps);
  else if constexpr (kMaxPer >= 8) {
    if (per <= 8) launch_ln<T, kTPR, 8>(s, p, gamma, beta, Y, ldy, M, nvec, inv_n, eps);
    else if constexpr (kMaxPer >= 16) launch_ln<T, kTPR, 16>(s, p, gamma, beta, Y, ldy, M, nvec, inv_n, eps);
  }
  return 0;
}

template <typename T>
static void launch_ln_any(cudaStream_t s, const LnSegments& p, const void* gamma, const void* beta, void* Y, long long ldy, int M, int nvec,
                          float inv_n, float eps) {
  // one warp per row while a lane holds <= 8 vectors, else one CTA per row
  // (measured on B200, bf16: 1024 channels 5.8 TB/s with a warp per row; 4 x 4096 channels 5.2 TB/s with 512 threads per row)
  int tpr = nvec <= 32 * 8 ? 32 : nvec <= 512 * 4 ? 512 : 256;
  if (const char* e = getenv("MERV_LN_TPR")) tpr = atoi(e);  // experiments: 32 | 256 | 512 | 1024
  int rc = 1;
  if (tpr == 32) rc = launch_ln_tpr<T, 32>(s, p, gamma, beta, Y, ldy, M, nvec, inv_n, eps);
  else if (tpr == 512) rc = launch_ln_tpr<T, 512>(s, p, gamma, beta, Y, ldy, M, nvec, inv_n, eps);
  else if (tpr == 1024) rc = launch_ln_tpr<T, 1024>(s, p, gamma, beta, Y, ldy, M, nvec, inv_n, eps);
  if (rc != 0) launch_ln_tpr<T, 256>(s, p, gamma, beta, Y, ldy, M, nvec, inv_n, eps);  // nvec <= 4096 always fits 256 x 16
}

extern "C" int merv_layernorm(const void* const* X, const int64_t* ldx, const int32_t* K, int nseg, const void* gamma, const void* beta,
                              float eps, void* Y, int64_t ldy, int M, int dtype, void* stream) {
  MERV_REQUIRE(dtype == MERV_F32 || dtype == MERV_BF16, MERV_E_DTYPE, "merv_layernorm: unknown dtype %d", dtype);
  const int vec = dtype == MERV_BF16 ? 8 : 4;
  LnSegments p;
  int total = 0;
  if (int rc = fill_segments("merv_layernorm", X, ldx, K, nseg, vec, &p, &total)) return rc;
  MERV_REQUIRE(Y && aligned16(Y) && ldy >= total && ldy % vec == 0, MERV_E_ALIGN, "merv_layernorm: Y NULL / misaligned or ldy=%lld < %d",
               (long long)ldy, total);
  MERV_REQUIRE((gamma == nullptr || aligned16(gamma)) && (beta == nullptr || aligned16(beta)), MERV_E_ALIGN,
               "merv_layernorm: gamma / beta must be 16-byte aligned");
  MERV_REQUIRE(M >= 0 && eps >= 0.f, MERV_E_SHAPE, "merv_layernorm: M=%d eps=%g", M, (double)eps);
  const int nvec = total / vec;
  MERV_REQUIRE(nvec <= 256 * 16, MERV_E_SHAPE, "merv_layernorm: rows of %d channels exceed the %d the kernel keeps in registers", total,
               256 * 16 * vec);
  if (int rc = require_sm100()) return rc;
  if (M == 0) return MERV_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const float inv_n = 1.0f / float(total);
  if (dtype == MERV_BF16)
    launch_ln_any<__nv_bfloat16>(s, p, gamma, beta, Y, ldy, M, nvec, inv_n, eps);
  else
    launch_ln_any<float>(s, p, gamma, beta, Y, ldy, M, nvec, inv_n, eps);
  MERV_CUDA_OK(cudaGetLastError());
  return MERV_OK;
}

extern "C" int merv_layernorm_backward(const void* const* X, const int64_t* ldx, const int32_t* K, int nseg, const void* dY, int64_t lddy,
                                       const void* gamma, float eps, void* dX, int64_t lddx, void* GX, int64_t ldgx, int M, int dtype,
                                       void* stream) {
  MERV_REQUIRE(dtype == MERV_F32 || dtype == MERV_BF16, MERV_E_DTYPE, "merv_layernorm_backward: unknown dtype %d", dtype);
  const int vec = dtype == MERV_BF16 ? 8 : 4;
  LnSegments p;
  int total = 0;
  if (int rc = fill_segments("merv_layernorm_backward", X, ldx, K, nseg, vec, &p, &total)) return rc;
  MERV_REQUIRE(dY && aligned16(dY) && lddy >= total && lddy % vec == 0, MERV_E_ALIGN, "merv_layernorm_backward: dY NULL / misaligned / lddy=%lld",
               (long long)lddy);
  MERV_REQUIRE(dX != nullptr || GX != nullptr, MERV_E_ARG, "merv_layernorm_backward: neither dX nor GX requested");
  MERV_REQUIRE(dX == nullptr || (aligned16(dX) && lddx >= total && lddx % vec == 0), MERV_E_ALIGN, "merv_layernorm_backward: dX misaligned / lddx=%lld",
               (long long)lddx);
  MERV_REQUIRE(GX == nullptr || (aligned16(GX) && ldgx >= total && ldgx % vec == 0), MERV_E_ALIGN, "merv_layernorm_backward: GX misaligned / ldgx=%lld",
               (long long)ldgx);
  MERV_REQUIRE(gamma == nullptr || aligned16(gamma), MERV_E_ALIGN, "merv_layernorm_backward: gamma must be 16-byte aligned");
  MERV_REQUIRE(M >= 0 && eps >= 0.f, MERV_E_SHAPE, "merv_layernorm_backward: M=%d eps=%g", M, (double)eps);
  const int nvec = total / vec;
  MERV_REQUIRE(nvec <= 256 * 16, MERV_E_SHAPE, "merv_layernorm_backward: rows of %d channels exceed the %d the kernel keeps in registers", total,
               256 * 16 * vec);
  if (int rc = require_sm100()) return rc;
  if (M == 0) return MERV_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const float inv_n = 1.0f / float(total);
  if (dtype == MERV_BF16) {
    using T = __nv_bfloat16;
    MERV_LN_DISPATCH(layernorm_bwd_kernel, T, p, (const T*)dY, lddy, (const T*)gamma, (T*)dX, lddx, (T*)GX, ldgx, M, nvec, inv_n, eps);
  } else {
    using T = float;
    MERV_LN_DISPATCH(layernorm_bwd_kernel, T, p, (const T*)dY, lddy, (const T*)gamma, (T*)dX, lddx, (T*)GX, ldgx, M, nvec, inv_n, eps);
  }
  MERV_CUDA_OK(cudaGetLastError());
  return MERV_OK;
}
