// LayerNorm over the channel dimension of (possibly segmented) rows — the two places the reference puts one on this path:
//   * nn.LayerNorm(vision_dim) in front of a projector (`pre_proj_layernorm=True`, merv/util/nn_utils.py:26-29,41-44,67-70,
//     91-94; constructed at merv/models/vidlms/merv.py:165-171);
//   * nn.LayerNorm(E * llm_dim) of feature_fusion == "concat_channel_ln" (merv.py:219-223), applied to the channel-wise
//     concatenation of the E projected token tensors (merv.py:603-606).  The kernel takes the E tensors as row SEGMENTS, so the
//     un-normalised concatenation is never written: every input element is read from HBM once, every output written once.
// Arithmetic follows ATen: fp32 statistics whatever the storage dtype, biased variance, y = (x - mean) * rstd * gamma + beta;
// two-pass (mean, then centred sum of squares) over values held in registers, so there is no E[x^2] - E[x]^2 cancellation.
// HBM-bound: algorithmic bytes per row = (in + out) * sum K_s * sizeof(T).
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

// kTPR threads per row (32: one warp per row, 8 rows per CTA; 256: one CTA per row); a thread keeps up to kMaxVec 16-byte
// vectors of its row in registers between the three passes.
template <typename T, int kTPR, int kMaxVec>
__global__ void __launch_bounds__(256) layernorm_kernel(const __grid_constant__ LnSegments p, const T* __restrict__ gamma,
                                                        const T* __restrict__ beta, T* __restrict__ y, long long ldy, int M, int nvec,
                                                        float inv_n, float eps) {
  constexpr int VEC = Vec16<T>::kN;
  constexpr int kRows = 256 / kTPR;
  __shared__ float red[32];
  const int tr = threadIdx.x % kTPR;
  const long long m = (long long)blockIdx.x * kRows + threadIdx.x / kTPR;
  const bool live = m < M;  // uniform per warp (kTPR == 32) or per CTA (kTPR == 256): the reductions below stay convergent
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
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = tr + j * kTPR;
    if (live && i < nvec) {
      float f[VEC], g[VEC], b[VEC];
      Vec16<T>::unpack(v[j], f);
      if (gamma != nullptr) {
        Vec16<T>::unpack(ldg_v4(gamma + (long long)i * VEC), g);
      } else {
#pragma unroll
        for (int c = 0; c < VEC; ++c) g[c] = 1.f;
      }
      if (beta != nullptr) {
        Vec16<T>::unpack(ldg_v4(beta + (long long)i * VEC), b);
      } else {
#pragma unroll
        for (int c = 0; c < VEC; ++c) b[c] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < VEC; ++c) f[c] = fmaf((f[c] - mean) * rstd, g[c], b[c]);
      stg_na_v4(y + m * ldy + (long long)i * VEC, Vec16<T>::pack(f));
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
    MERV_REQUIRE(X[s] && aligned16(X[s]) && ldx[s] >= K[s] && ldx[s] % vec == 0, MERV_E_ALIGN,
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

// warp-per-row up to 8 vectors per lane, CTA-per-row up to 16 vectors per thread
#define MERV_LN_DISPATCH(KERNEL, T, ...)                                                                   \
  do {                                                                                                     \
    if (nvec <= 32 * 8)                                                                                    \
      KERNEL<T, 32, 8><<<(unsigned)((M + 7) / 8), 256, 0, s>>>(__VA_ARGS__);                                \
    else                                                                                                   \
      KERNEL<T, 256, 16><<<(unsigned)M, 256, 0, s>>>(__VA_ARGS__);                                          \
  } while (0)

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
  if (dtype == MERV_BF16) {
    using T = __nv_bfloat16;
    MERV_LN_DISPATCH(layernorm_kernel, T, p, (const T*)gamma, (const T*)beta, (T*)Y, ldy, M, nvec, inv_n, eps);
  } else {
    using T = float;
    MERV_LN_DISPATCH(layernorm_kernel, T, p, (const T*)gamma, (const T*)beta, (T*)Y, ldy, M, nvec, inv_n, eps);
  }
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
