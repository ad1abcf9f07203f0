// Pooling INSIDE the fused projector GEMM ("pool assist"): the adaptive average pooling of AveragePooling3DProjector.forward
// (reference merv/util/nn_utils.py:320-329) is HBM-bound and leaves the tensor pipe idle, the projector GEMM is tensor-bound and leaves
// HBM idle, and as two launches they cannot be co-resident (the GEMM CTA owns the SM's shared memory and registers).  Here the two spare
// warps of the GEMM CTA (warps 2 and 3 of its producer warpgroup) pool the videos AHEAD of the tiles being multiplied:
//
//   * work items (video, encoder, frame, 64-channel chunk, half of the output rows) are handed out by ONE global atomic counter in
//     video-major order, so the videos complete in order whatever subset of the grid is resident (no CTA waits for a CTA that may not
//     have been scheduled yet: the resident assist warps can always drain the whole list);
//   * an item is loaded as two quarter slabs [4 input rows x W x 128 bytes] by 5-D TMA into a private double buffer (L2 evict-first:
//     the raw features are read exactly once) and reduced by the SAME separable packed-fp32 arithmetic, thread mapping and summation
//     order as pool3d.cu's pool_square consumer: pooled tokens and score partials are bit-identical to the standalone kernel's;
//   * the warp that completes a video's last item (global counter per video) reduces the video's score partials, does the softmax over
//     the encoders and writes the mixing weights and the per-video bias row exactly as mix.cu's scores_softmax_kernel does, then
//     publishes ready[video] with a release store;
//   * the GEMM's TMA producer acquires ready[video] (+ a generic->async proxy fence) before its first load of a video's pooled rows; the
//     epilogue's reads of the video's mixing weights and bias row are ordered behind that acquire by the mbarrier chain
//     (TMA complete_tx -> MMA -> tcgen05.commit -> epilogue) and bypass L1.
//
// The first `head` videos are pooled by the standalone kernel (they have to exist before the first tile can start).
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "tcgen05_util.cuh"

namespace merv {

// ---- pieces shared with pool3d.cu ---------------------------------------------------------------------------
template <int N_IN, int N_OUT>
struct PoolWin {  // adaptive pooling window k of an axis N_IN -> N_OUT: [floor(k N_IN / N_OUT), ceil((k + 1) N_IN / N_OUT))
  __host__ __device__ static constexpr int lo(int i) { return (i * N_IN) / N_OUT; }
  __host__ __device__ static constexpr int hi(int i) { return ((i + 1) * N_IN + N_OUT - 1) / N_OUT; }
};

__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
// predicated 128-bit shared load: zeros when !ok, no branch
__device__ __forceinline__ uint4 lds_v4_if(uint32_t addr, bool ok) {
  uint4 r;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %5, 0;\n\t"
      "mov.u32 %0, 0;\n\tmov.u32 %1, 0;\n\tmov.u32 %2, 0;\n\tmov.u32 %3, 0;\n\t"
      "@p ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr), "r"(uint32_t(ok)));
  return r;
}

// ---- arguments ----------------------------------------------------------------------------------------------
// build-time switch: per-warp cycle counters of the pooling warps (wait for TMA / pooling arithmetic / completion accounting), written to
// AssistArgs::prof when MERV_ASSIST_PROFILE=1 is also set in the environment.  Off in production builds (registers).
#ifndef MERV_ASSIST_PROFILE
#define MERV_ASSIST_PROFILE 0
#endif
constexpr int ASSIST_S = 8;                 // output grid 8 x 8 per frame
constexpr int ASSIST_CB = 64;               // channels per item: 128-byte rows of bf16
constexpr int ASSIST_ROWS = 4;              // input rows per quarter slab (two output rows; 16 -> 8: 4 rows, 14 -> 8: 4 rows with one shared)
constexpr int ASSIST_QUARTER_BYTES = 8192;  // 4 rows x 16 columns x 128 bytes (14 x 14 grids use 7168 of them)
constexpr int ASSIST_WARPS = 2;
constexpr int ASSIST_SMEM_BYTES = ASSIST_WARPS * 2 * ASSIST_QUARTER_BYTES;  // 32 KB: one ring stage of the CTA-pair GEMM
constexpr int ASSIST_SYNC_HEADER = 4;       // ints in front of the per-video counters: [0] = next work item
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;

struct AssistEnc {
  __nv_bfloat16* y;            // pooled tokens [B, T * 64, C]
  const float* score_vec;      // [C]
  float* score_partial;        // [B, parts]
  const int* batch_index;      // [B] or NULL
  const float* c;              // score constant or NULL
  const __nv_bfloat16* bias;   // [N] or NULL
  long long ybs, yrs;
  int H, C, T, nchunks;
  int item_begin;              // first item of this encoder inside a video's item list
  int parts;                   // score partials per video = T * nchunks * 4
};
struct AssistArgs {
  CUtensorMap m[MERV_MAX_SEGMENTS];  // x as (C, W, H, F, B), boxes of [64 channels, W, 4 rows, 1 frame, 1 video]
  AssistEnc enc[MERV_MAX_SEGMENTS];
  int n_enc, head, B, items_per_video, total_items, N, rows_per_video;
  int group;                   // consecutive items handed out per fetch (divides items_per_video)
  int dbg;                     // profile builds only (MERV_ASSIST_DBG): 1 = no score dot, 2 = no shared loads, 4 = no pooled-row stores
  int* prof;                   // optional per-warp cycle counters (MERV_ASSIST_PROFILE=1), 8 ints per (CTA, pooling warp); NULL otherwise
  int* sync;                   // [ASSIST_SYNC_HEADER + 2 B]: work counter, per-video completion counters, per-video ready flags
  float* scores;               // [B, E]
  float* weights;              // [B, E]
  __nv_bfloat16* weights_bf16; // [B, E] or NULL
  float* bias_mix;             // [B, N]
};

#ifdef __CUDACC__
// lane e (< E) passes its score; every lane gets the weight of lane e back in the same lane (shared with mix.cu)
__device__ __forceinline__ float warp_softmax_over_encoders(float score, int lane, int E) {
  const float s = lane < E ? score : -INFINITY;
  const float m = warp_max(s);
  const float ex = lane < E ? expf(s - m) : 0.f;
  const float den = warp_sum(ex);
  return ex / den;
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void assist_tma_load(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "l"(L2_EVICT_FIRST) : "memory");
}

// Two output rows (2 OQ, 2 OQ + 1) of one output column `wo` for one 16-byte channel vector `v`, from a quarter slab holding input rows
// lo(2 OQ) .. lo(2 OQ) + 3.  Arithmetic identical to pool_square_half (pool3d.cu): per input row the 2-3 taps of the column window are summed
// once, the row sum is added to the output rows whose window holds the row (ascending rows), then scaled by 1 / (hn wn); the score dot
// product continues the caller's packed accumulator in the order (output row, channel pair).
template <int H, int OQ>
__device__ __forceinline__ void assist_quarter(int oq, uint32_t slab, int v, int wo, __nv_bfloat16* __restrict__ yb, long long yrs, bool c_ok,
                                               const float* __restrict__ svp, unsigned long long& dot2, int dbg) {
  typedef unsigned long long u64;
  typedef __nv_bfloat16 T;
  constexpr int S = ASSIST_S, P = 4;
  constexpr int R_LO = PoolWin<H, S>::lo(2 * OQ), R_HI = PoolWin<H, S>::hi(2 * OQ + 1);
  static_assert(R_HI - R_LO <= ASSIST_ROWS, "a quarter slab holds four input rows");
  constexpr bool kThird = (H % S) != 0;
  const int w0 = (wo * H) / S;
  const int wn = ((wo + 1) * H + S - 1) / S - w0;
  const u64 one2 = f32x2_pack(1.0f, 1.0f);
  u64 acc[2][P];
#pragma unroll
  for (int o = 0; o < 2; ++o)
#pragma unroll
    for (int q = 0; q < P; ++q) acc[o][q] = f32x2_pack(0.f, 0.f);
  const uint32_t col = slab + uint32_t(w0) * 128u + uint32_t(v) * 16u;
#pragma unroll
  for (int r = R_LO; r < R_HI; ++r) {
    const uint32_t a = col + uint32_t((r - R_LO) * H * 128);
    u64 rs[P], t[P];
    Pairs<T>::unpack(lds_v4_if(a, !(dbg & 2)), rs);
    if constexpr (H / S >= 2 || kThird) {
      Pairs<T>::unpack(lds_v4_if(a + 128u, wn >= 2 && !(dbg & 2)), t);
#pragma unroll
      for (int q = 0; q < P; ++q) rs[q] = f32x2_fma(t[q], one2, rs[q]);
    }
    if constexpr (H / S + (kThird ? 2 : 0) >= 3) {
      Pairs<T>::unpack(lds_v4_if(a + 256u, wn >= 3 && !(dbg & 2)), t);
#pragma unroll
      for (int q = 0; q < P; ++q) rs[q] = f32x2_fma(t[q], one2, rs[q]);
    }
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      if (r >= PoolWin<H, S>::lo(2 * OQ + o) && r < PoolWin<H, S>::hi(2 * OQ + o)) {
#pragma unroll
        for (int q = 0; q < P; ++q) acc[o][q] = f32x2_fma(rs[q], one2, acc[o][q]);
      }
    }
  }
  u64 sv2[P];
#pragma unroll
  for (int q = 0; q < P; ++q) sv2[q] = f32x2_pack(0.f, 0.f);
  if (c_ok && !(dbg & 1)) {
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(svp)), s1 = __ldg(reinterpret_cast<const float4*>(svp) + 1);
    sv2[0] = f32x2_pack(s0.x, s0.y); sv2[1] = f32x2_pack(s0.z, s0.w); sv2[2] = f32x2_pack(s1.x, s1.y); sv2[3] = f32x2_pack(s1.z, s1.w);
  }
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    const int hn = PoolWin<H, S>::hi(2 * OQ + o) - PoolWin<H, S>::lo(2 * OQ + o);
    // 1 / (hn wn) (x 1 / frames of the temporal window: always one here) — wn is 1, 2 or 3: constants instead of a division per output
    const float inv = wn == 2 ? 1.0f / float(hn * 2) : wn == 3 ? 1.0f / float(hn * 3) : 1.0f / float(hn);
    const u64 inv2 = f32x2_pack(inv, inv);
    u64 out[P];
#pragma unroll
    for (int q = 0; q < P; ++q) out[q] = f32x2_mul(acc[o][q], inv2);
    const uint4 packed = Pairs<T>::pack(out);
    const int tok = (2 * oq + o) * S + wo;  // oq: the actual output-row pair (OQ only carries its window pattern)
    if (c_ok && !(dbg & 4)) *reinterpret_cast<uint4*>(yb + (long long)tok * yrs) = packed;
    u64 rr[P];
    Pairs<T>::unpack(packed, rr);  // dot with the values as stored (what the GEMM will consume)
#pragma unroll
    for (int q = 0; q < P; ++q) dot2 = f32x2_fma(sv2[q], rr[q], dot2);
  }
}

// The window pattern of an output-row pair relative to its quarter slab depends on the pair's parity only (16 -> 8: rows {0,1} | {2,3} for
// every pair; 14 -> 8: {0,1} | {1,2,3} for even pairs, {0,1,2} | {2,3} for odd ones), so three instantiations cover all eight pairs.
template <int H, int PAR> struct AssistWindowsMatch {
  static constexpr bool same(int oq) {
    for (int o = 0; o < 2; ++o)
      if (PoolWin<H, ASSIST_S>::lo(2 * oq + o) - PoolWin<H, ASSIST_S>::lo(2 * oq) != PoolWin<H, ASSIST_S>::lo(2 * PAR + o) - PoolWin<H, ASSIST_S>::lo(2 * PAR) ||
          PoolWin<H, ASSIST_S>::hi(2 * oq + o) - PoolWin<H, ASSIST_S>::lo(2 * oq) != PoolWin<H, ASSIST_S>::hi(2 * PAR + o) - PoolWin<H, ASSIST_S>::lo(2 * PAR))
        return false;
    return true;
  }
};
static_assert(AssistWindowsMatch<16, 0>::same(1) && AssistWindowsMatch<16, 0>::same(2) && AssistWindowsMatch<16, 0>::same(3), "16 -> 8: one pattern");
static_assert(AssistWindowsMatch<14, 0>::same(2) && AssistWindowsMatch<14, 1>::same(3), "14 -> 8: one pattern per parity");

// first input row of the quarter slab for output-row pair oq
__device__ __forceinline__ int assist_row0(int H, int oq) { return (2 * oq * H) / ASSIST_S; }

// One warp: everything mix.cu's scores_softmax_kernel does for video b, with the same summation trees (a 256-thread block there is 8
// "virtual warps" here), so scores, weights and the bias row come out bit-identical.
static __device__ __noinline__ void assist_finish_video(const AssistArgs& a, int b, int lane) {
  const int E = a.n_enc;
  float my_score = 0.f;
#pragma unroll 1
  for (int e = 0; e < E; ++e) {
    const AssistEnc& en = a.enc[e];
    const int n = en.parts;
    const float* src = en.score_partial + (long long)b * n;
    float red = 0.f;
#pragma unroll 1
    for (int vw = 0; vw < 8; ++vw) {
      float acc = 0.f;
      for (int i = vw * 32 + lane; i < n; i += 256) acc += __ldcg(src + i);
      acc = warp_sum(acc);
      if (lane == vw) red = acc;
    }
    const float r = warp_sum(red);
    const float sc = r / float(a.rows_per_video) + (en.c ? __ldg(en.c) : 0.f);
    if (lane == e) my_score = sc;
  }
  const float w = warp_softmax_over_encoders(lane < E ? my_score : 0.f, lane, E);
  if (lane < E) {
    a.scores[(long long)b * E + lane] = my_score;
    a.weights[(long long)b * E + lane] = w;
    if (a.weights_bf16 != nullptr) a.weights_bf16[(long long)b * E + lane] = __float2bfloat16_rn(w);
  }
  const int N = a.N;
  float we[MERV_MAX_SEGMENTS];
  const __nv_bfloat16* be[MERV_MAX_SEGMENTS];
#pragma unroll
  for (int e = 0; e < MERV_MAX_SEGMENTS; ++e) {
    we[e] = __shfl_sync(0xffffffffu, w, e);
    be[e] = e < E ? a.enc[e].bias : nullptr;
  }
#pragma unroll 4
  for (int nn = lane; nn < N; nn += 32) {
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < MERV_MAX_SEGMENTS; ++e)
      if (be[e] != nullptr) acc += we[e] * to_float(__ldg(be[e] + nn));
    a.bias_mix[(long long)b * N + nn] = acc;
  }
}

// a work item, warp-uniform: code = encoder | half << 2 | chunk << 3 | frame << 12
struct AssistItem {
  int code, b, src_b;
  __device__ __forceinline__ int e() const { return code & 3; }
  __device__ __forceinline__ int hf() const { return (code >> 2) & 1; }
  __device__ __forceinline__ int chunk() const { return (code >> 3) & 511; }
  __device__ __forceinline__ int t() const { return code >> 12; }
};

__device__ __forceinline__ AssistItem assist_decode(const AssistArgs& a, int idx) {
  AssistItem it;
  it.b = a.head + idx / a.items_per_video;
  const int r = idx % a.items_per_video;
  int e = 0;
#pragma unroll 1
  while (e + 1 < a.n_enc && r >= a.enc[e + 1].item_begin) ++e;
  const AssistEnc& en = a.enc[e];
  const int local = r - en.item_begin;
  const int ct = local >> 1;
  it.code = e | ((local & 1) << 2) | ((ct % en.nchunks) << 3) | ((ct / en.nchunks) << 12);
  it.src_b = en.batch_index ? __ldg(en.batch_index + it.b) : it.b;
  return it;
}

__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// the item after `it` inside the same video (groups never straddle a video): no integer division
__device__ __forceinline__ AssistItem assist_advance(const AssistArgs& a, AssistItem it) {
  int e = it.e(), hf = it.hf(), chunk = it.chunk(), t = it.t();
  if (hf == 0) {
    hf = 1;
  } else {
    hf = 0;
    const AssistEnc& en = a.enc[e];
    if (++chunk == en.nchunks) {
      chunk = 0;
      if (++t == en.T) {
        t = 0;
        ++e;
        const AssistEnc& en2 = a.enc[e];
        it.src_b = en2.batch_index ? __ldg(en2.batch_index + it.b) : it.b;
      }
    }
  }
  it.code = e | (hf << 2) | (chunk << 3) | (t << 12);
  return it;
}

// the loop of one assist warp.  `buf`: this warp's two quarter buffers (2 x ASSIST_QUARTER_BYTES), `bar`: its two mbarriers (count 1).
// Work is fetched in groups of a.group consecutive items (one video): one release fence + one completion atomic per group, whose
// result (am I the warp that completed the video?) is only looked at after the next group has been pooled.
__device__ __forceinline__ void assist_warp_loop(const AssistArgs& a, uint32_t buf, uint32_t bar, int lane, int gwarp) {
  typedef unsigned long long u64;
  int* next = a.sync;
  int* done = a.sync + ASSIST_SYNC_HEADER;
  int* ready = done + a.B;
  const int total = a.total_items, G = a.group;
  uint32_t phase = 0;  // both buffers flip together: an item uses each exactly once
#if MERV_ASSIST_PROFILE
  long long t_wait = 0, t_comp = 0, t_fin = 0, t_video = 0, t_all = clock64();
  int n_items = 0;
#define MERV_AP(stmt) stmt
#define MERV_DBG a.dbg
#else
#define MERV_AP(stmt)
#define MERV_DBG 0
#endif
  auto fetch_raw = [&]() {  // lane 0 only holds the result; broadcast it when it is needed (the atomic's latency stays hidden until then)
    int v = 0;
    if (lane == 0) v = atomicAdd(next, G);
    return v;
  };
  auto bcast = [&](int v) { return __shfl_sync(0xffffffffu, v, 0); };
  auto issue = [&](const AssistItem& it, int q) {
    if (lane == 0) {
      const AssistEnc& en = a.enc[it.e()];
      const uint32_t b = bar + 8u * uint32_t(q);
      mbar_expect_tx(b, uint32_t(ASSIST_ROWS * en.H * 128));
      assist_tma_load(&a.m[it.e()], b, buf + uint32_t(q) * ASSIST_QUARTER_BYTES, it.chunk() * ASSIST_CB, 0, assist_row0(en.H, 2 * it.hf() + q),
                      it.t(), it.src_b);
    }
  };
  auto complete = [&](int b, int old) {  // `old`: the video's completion count before this warp's group was added
    if (old + G == a.items_per_video) {
      MERV_AP(const long long t0 = clock64();)
      fence_acq_rel_gpu();
      assist_finish_video(a, b, lane);
      fence_acq_rel_gpu();
      __syncwarp();
      if (lane == 0) st_release_gpu(ready + b, 1);
      MERV_AP(t_video += clock64() - t0;)
    }
  };
  int cur = bcast(fetch_raw());      // first item of the current group
  int nxt_raw = fetch_raw();         // next group (lane 0)
  int pend_b = -1, pend_old_raw = 0; // completion atomic of the previous group, not yet looked at
  AssistItem ci = {}, ni = {};
  if (cur < total) {
    ci = assist_decode(a, cur);
    issue(ci, 0);
    issue(ci, 1);
  }
  const int v = lane & 7;
#pragma unroll 1
  while (cur < total) {
    int nxt = 0;
    const int group_b = ci.b;
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      // the item after this one: the next of the group, or the first of the next group
      bool has_next;
      if (g + 1 < G) {
        has_next = true;
        ni = assist_advance(a, ci);
      } else {
        nxt = bcast(nxt_raw);
        has_next = nxt < total;
        if (has_next) ni = assist_decode(a, nxt);
        nxt_raw = fetch_raw();
      }
      const AssistEnc& en = a.enc[ci.e()];
      const int c_base = ci.chunk() * ASSIST_CB + v * 8;
      const bool c_ok = c_base < en.C;
      __nv_bfloat16* yb = en.y + (long long)ci.b * en.ybs + (long long)ci.t() * (ASSIST_S * ASSIST_S) * en.yrs + c_base;
      const float* svp = en.score_vec + c_base;
      u64 dot_a = f32x2_pack(0.f, 0.f), dot_b = dot_a;
#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        MERV_AP(const long long t0 = clock64();)
        mbar_wait(bar + 8u * uint32_t(q), phase);
        MERV_AP(const long long t1 = clock64();)
        const uint32_t slab = buf + uint32_t(q) * ASSIST_QUARTER_BYTES;
        const int oq = 2 * ci.hf() + q;
#pragma unroll 1
        for (int vw = 0; vw < 2; ++vw) {
          const int wo = 4 * vw + (lane >> 3);
          u64 d = vw ? dot_b : dot_a;
          if (en.H == 16) assist_quarter<16, 0>(oq, slab, v, wo, yb, en.yrs, c_ok, svp, d, MERV_DBG);
          else if (oq & 1) assist_quarter<14, 1>(oq, slab, v, wo, yb, en.yrs, c_ok, svp, d, MERV_DBG);
          else assist_quarter<14, 0>(oq, slab, v, wo, yb, en.yrs, c_ok, svp, d, MERV_DBG);
          if (vw) dot_b = d; else dot_a = d;
        }
        __syncwarp();  // every lane has read the buffer before the next load may overwrite it
        if (has_next) issue(ni, q);
        MERV_AP(t_wait += t1 - t0; t_comp += clock64() - t1;)
      }
      phase ^= 1u;
#pragma unroll
      for (int vw = 0; vw < 2; ++vw) {
        float d0, d1;
        f32x2_unpack(vw ? dot_b : dot_a, d0, d1);
        float dot = 0.f;
        dot += d0 + d1;
        dot = warp_sum(dot);
        if (lane == 0) en.score_partial[(((long long)ci.b * en.T + ci.t()) * en.nchunks + ci.chunk()) * 4 + ci.hf() * 2 + vw] = dot;
      }
      MERV_AP(++n_items;)
      ci = ni;
    }
    MERV_AP(const long long t2 = clock64();)
    // the previous group's completion count has had a whole group of pooling to come back
    if (pend_b >= 0) complete(pend_b, bcast(pend_old_raw));
    fence_acq_rel_gpu();  // pooled rows and partials of this group: visible device-wide before the completion count moves
    __syncwarp();
    if (lane == 0) pend_old_raw = atomicAdd(done + group_b, G);
    pend_b = group_b;
    MERV_AP(t_fin += clock64() - t2;)
    cur = nxt;
  }
  if (pend_b >= 0) complete(pend_b, bcast(pend_old_raw));
#if MERV_ASSIST_PROFILE
  if (a.prof != nullptr && lane == 0) {
    int* pr = a.prof + gwarp * 8;
    pr[0] = n_items; pr[1] = int(t_wait >> 4); pr[2] = int(t_comp >> 4); pr[3] = int(t_fin >> 4); pr[4] = int(t_video >> 4);
    pr[5] = int((clock64() - t_all) >> 4);
  }
#endif
#undef MERV_AP
#undef MERV_DBG
  (void)gwarp;
}
#endif  // __CUDACC__

// host side (pool3d.cu): fills `args` for the pool-assist variant of merv_fused_forward; returns false (args untouched) when the
// configuration is not covered (then the caller runs the three-launch path)
bool build_pool_assist(const merv_fused_desc* d, int head, AssistArgs* args, int* rc);

}  // namespace merv
