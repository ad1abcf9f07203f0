// Adaptive spatio-temporal average pooling, channel-last in -> channel-last out, all encoders in one launch.
//
// Replaces the rearrange -> AdaptiveAvgPool3d -> rearrange of AveragePooling3DProjector.forward
// (reference merv/util/nn_utils.py:320-329).  The reference permutes to channel-major, pools with ATen's
// generic kernel and permutes back through a transposed view; here the channel dimension stays the
// contiguous one end to end: one thread owns a 16-byte channel vector, so every global access is a fully
// coalesced 128-bit load/store, and the [B, T*S*S, C] result is directly the K-major A operand of the
// projector GEMM.  HBM-bound: compulsory traffic = input + pooled output (33.75 MB/video at merv-full).
//
// Work item = (video b, output frame t, group of output rows); window overlap re-reads (14 -> 8 has 2/3-wide
// overlapping windows) are served by L1/L2, not DRAM.  Optionally emits deterministic per-item column sums of
// the pooled tokens (fp32) from which the affine fast path derives the encoder scores without touching the
// projected tokens.
#include "common.cuh"

namespace merv {

struct PoolEnc {
  const void* x;
  void* y;
  float* colsum;
  int F, H, W, C, T, S;
  int rows_per_item, groups;  // output rows handled by one CTA; groups = ceil(S / rows_per_item)
  int items;                  // B * T * groups
  long long xbs, xfs, xts, ybs, yrs;
};
struct PoolParams {
  PoolEnc enc[MERV_MAX_ENCODERS];
};

__device__ __forceinline__ void window(int k, int n_in, int n_out, int& lo, int& hi) {
  lo = (k * n_in) / n_out;                  // floor(k * n_in / n_out)
  hi = ((k + 1) * n_in + n_out - 1) / n_out;  // ceil((k+1) * n_in / n_out)
}

template <typename T>
__global__ void __launch_bounds__(256) pool3d_kernel(const __grid_constant__ PoolParams p) {
  constexpr int VEC = Vec16<T>::kN;
  const PoolEnc& e = p.enc[blockIdx.y];
  const int item = blockIdx.x;
  if (item >= e.items) return;
  const int g = item % e.groups;
  const int t = (item / e.groups) % e.T;
  const int b = item / (e.groups * e.T);
  int f0, f1;
  window(t, e.F, e.T, f0, f1);
  const int i_begin = g * e.rows_per_item;
  const int i_end = min(e.S, i_begin + e.rows_per_item);
  const int nvec = e.C / VEC;
  const T* __restrict__ xb = static_cast<const T*>(e.x) + (long long)b * e.xbs;
  T* __restrict__ yb = static_cast<T*>(e.y) + (long long)b * e.ybs;

  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float csum[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) csum[c] = 0.f;
    for (int i = i_begin; i < i_end; ++i) {
      int h0, h1;
      window(i, e.H, e.S, h0, h1);
      for (int j = 0; j < e.S; ++j) {
        int w0, w1;
        window(j, e.W, e.S, w0, w1);
        float acc[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
        for (int f = f0; f < f1; ++f) {
          for (int h = h0; h < h1; ++h) {
            const T* row = xb + (long long)f * e.xfs + (long long)(h * e.W) * e.xts + (long long)v * VEC;
#pragma unroll 3
            for (int w = w0; w < w1; ++w) {
              float val[VEC];
              Vec16<T>::unpack(ldg_v4(row + (long long)w * e.xts), val);
#pragma unroll
              for (int c = 0; c < VEC; ++c) acc[c] += val[c];
            }
          }
        }
        const float cnt = float((f1 - f0) * (h1 - h0) * (w1 - w0));
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc[c] = acc[c] / cnt;
        const uint4 packed = Vec16<T>::pack(acc);
        const long long tok = (long long)(t * e.S + i) * e.S + j;
        *reinterpret_cast<uint4*>(yb + tok * e.yrs + (long long)v * VEC) = packed;
        // sums of the values the GEMM will actually consume (i.e. after rounding to the storage dtype)
        float rounded[VEC];
        Vec16<T>::unpack(packed, rounded);
#pragma unroll
        for (int c = 0; c < VEC; ++c) csum[c] += rounded[c];
      }
    }
    if (e.colsum != nullptr) {
      float* dst = e.colsum + ((long long)b * (e.T * e.groups) + (long long)t * e.groups + g) * e.C + (long long)v * VEC;
#pragma unroll
      for (int c = 0; c < VEC; c += 4)
        *reinterpret_cast<float4*>(dst + c) = make_float4(csum[c], csum[c + 1], csum[c + 2], csum[c + 3]);
    }
  }
}

static int rows_per_item_for(int T, int S, int B) {
  // Up to 4 row-groups per output frame: >= 4096 CTAs per encoder at the benchmark batch (several waves on 148
  // SMs) while the optional column-sum side output stays ~2 % of the traffic.  Deliberately independent of B
  // so the fp32 summation order of the column sums — hence every mixing weight — is bit-identical whether a
  // video is processed alone, in a batch of 64, or on another rank.
  (void)T;
  (void)B;
  const int groups = S < 4 ? S : 4;
  return (S + groups - 1) / groups;
}

}  // namespace merv

extern "C" int merv_pool3d_colsum_parts(int T, int S, int B) {
  if (T <= 0 || S <= 0 || B <= 0) return 0;
  const int rpi = merv::rows_per_item_for(T, S, B);
  return T * ((S + rpi - 1) / rpi);
}

extern "C" int merv_pool3d(const merv_pool_desc* enc, int num_encoders, int B, int dtype, void* stream) {
  using namespace merv;
  MERV_REQUIRE(enc != nullptr, MERV_E_ARG, "merv_pool3d: enc is NULL");
  MERV_REQUIRE(num_encoders >= 1 && num_encoders <= MERV_MAX_ENCODERS, MERV_E_ARG,
               "merv_pool3d: num_encoders=%d not in [1,%d]", num_encoders, MERV_MAX_ENCODERS);
  MERV_REQUIRE(dtype == MERV_F32 || dtype == MERV_BF16, MERV_E_DTYPE, "merv_pool3d: unknown dtype %d", dtype);
  MERV_REQUIRE(B >= 0, MERV_E_SHAPE, "merv_pool3d: B=%d", B);
  if (int rc = require_sm100()) return rc;
  if (B == 0) return MERV_OK;
  const int vec = dtype == MERV_BF16 ? 8 : 4;
  PoolParams p;
  int max_items = 0, max_vec = 0;
  for (int i = 0; i < num_encoders; ++i) {
    const merv_pool_desc& d = enc[i];
    MERV_REQUIRE(d.x && d.y, MERV_E_ARG, "merv_pool3d: encoder %d has a NULL tensor", i);
    MERV_REQUIRE(d.F > 0 && d.H > 0 && d.W > 0 && d.C > 0 && d.T > 0 && d.S > 0, MERV_E_SHAPE,
                 "merv_pool3d: encoder %d has a non-positive dimension (F=%d H=%d W=%d C=%d T=%d S=%d)", i, d.F,
                 d.H, d.W, d.C, d.T, d.S);
    MERV_REQUIRE(d.C % vec == 0, MERV_E_SHAPE, "merv_pool3d: encoder %d: C=%d must be a multiple of %d", i, d.C, vec);
    MERV_REQUIRE(aligned16(d.x) && aligned16(d.y) && (d.colsum == nullptr || aligned16(d.colsum)), MERV_E_ALIGN,
                 "merv_pool3d: encoder %d: base pointers must be 16-byte aligned", i);
    MERV_REQUIRE(d.x_batch_stride % vec == 0 && d.x_frame_stride % vec == 0 && d.x_token_stride % vec == 0 &&
                     d.y_batch_stride % vec == 0 && d.y_row_stride % vec == 0 && d.x_token_stride >= d.C &&
                     d.y_row_stride >= d.C,
                 MERV_E_ALIGN, "merv_pool3d: encoder %d: strides must be multiples of %d elements and >= C", i, vec);
    PoolEnc& e = p.enc[i];
    e.x = d.x; e.y = d.y; e.colsum = d.colsum;
    e.F = d.F; e.H = d.H; e.W = d.W; e.C = d.C; e.T = d.T; e.S = d.S;
    e.rows_per_item = rows_per_item_for(d.T, d.S, B);
    e.groups = (d.S + e.rows_per_item - 1) / e.rows_per_item;
    e.items = B * d.T * e.groups;
    e.xbs = d.x_batch_stride; e.xfs = d.x_frame_stride; e.xts = d.x_token_stride;
    e.ybs = d.y_batch_stride; e.yrs = d.y_row_stride;
    max_items = e.items > max_items ? e.items : max_items;
    max_vec = d.C / vec > max_vec ? d.C / vec : max_vec;
  }
  int threads = ((max_vec + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  dim3 grid(max_items, num_encoders);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == MERV_BF16)
    pool3d_kernel<__nv_bfloat16><<<grid, threads, 0, s>>>(p);
  else
    pool3d_kernel<float><<<grid, threads, 0, s>>>(p);
  MERV_CUDA_OK(cudaGetLastError());
  return MERV_OK;
}
