// Cross attention of the attentive pooler ("attntv" resampler) on the 5th-gen tensor cores.
//   CrossAttention.forward, merv/util/nn_utils.py:393-412 (inside AttentivePooler.forward :229-238):
//       out[b, q, h*hd:(h+1)*hd] = softmax_k(scale * Q[q, h] . K[b, k, h]) V[b, k, h]
// K and V are the two halves of the `kv` Linear's output row ([K(C) | V(C)], heads contiguous inside each half, :398-399); the learned
// queries are the same for every batch entry (frame).  Both contractions are GEMM-shaped (2 * 2 * n_q * n_kv * C FLOP per frame) and run
// as tcgen05.mma with the accumulators in TMEM; the fp32 SIMT kernel of attention.cu remains the 1e-5 parity path (fp32 storage) and
// the fallback for shapes outside this kernel's tile (more than 128 queries or 256 keys per frame, head_dim not a multiple of 32).
//
// Work item = (frame b, head h); persistent CTAs, 6 warps:
//   warp 4  producer : TMA loads of Q_h [n_q x hd], K_h [n_kv x hd], V_h [n_kv x hd] through 4-D tensor maps (hd, head, row, frame): rows
//                      past n_q / n_kv and head dims past hd are zero-filled, so padding never needs masking on the operand side
//   warp 5  MMA      : S = Q_h K_h^T   (UMMA 128 x n_kv x 16, A and B K-major)            -> TMEM columns [0, 256)
//                      O = P V_h       (UMMA 128 x hd x 16, A = P K-major from shared memory, B = V_h read MN-major IN PLACE:
//                                       the same [keys x hd] image the TMA wrote, no transposed copy)  -> TMEM columns [256, 256 + hd)
//   warps 0-3 softmax: thread = query row: two passes over its S row in TMEM (tcgen05.ld): max, then exp2 / sum / bf16 P into shared
//                      memory in the 128-byte-swizzled K-major layout the second MMA reads; later O * (1 / sum) -> bf16 -> global.
// K is released as soon as the S MMAs have read it and V as soon as the PV MMAs have, so the loads of the next item overlap the
// softmax and the second MMA of the current one.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tcgen05_util.cuh"
#include "tmap.cuh"

namespace merv {

constexpr int AT_M = 128;                  // UMMA M: query rows (zero-padded)
constexpr int AT_KEYS = 256;               // keys per frame covered by the S tile
constexpr int AT_MAX_HD = 128;
constexpr int AT_Q_CHUNK = AT_M * 128;     // [128 rows x 64 head dims] bf16, 128-byte swizzle
constexpr int AT_KV_CHUNK = AT_KEYS * 128; // [256 keys x 64 head dims]
constexpr int AT_P_CHUNK = AT_M * 128;     // [128 rows x 64 keys]
constexpr int AT_HD_CHUNKS = AT_MAX_HD / 64;
constexpr int AT_Q_OFF = 0;
constexpr int AT_K_OFF = AT_Q_OFF + AT_HD_CHUNKS * AT_Q_CHUNK;
constexpr int AT_V_OFF = AT_K_OFF + AT_HD_CHUNKS * AT_KV_CHUNK;
constexpr int AT_P_OFF = AT_V_OFF + AT_HD_CHUNKS * AT_KV_CHUNK;
constexpr int AT_BAR_OFF = AT_P_OFF + (AT_KEYS / 64) * AT_P_CHUNK;
constexpr int AT_SMEM = 1024 + AT_BAR_OFF + 128;
static_assert(AT_SMEM <= 232448, "227 KB of shared memory per CTA");
constexpr int AT_THREADS = 192;
constexpr int AT_TMEM_COLS = 512;          // S: columns [0, 256), O: [256, 256 + hd)

struct AttnParams {
  __nv_bfloat16* out;
  long long ldo;
  int batches, heads, n_q, n_kv, hd;
  int q_per_batch;   // 1: every frame has its own queries (4th coordinate of the Q map = frame), 0: shared learned queries
  float scale_log2e; // softmax scale * log2(e)
};
struct AttnMaps {
  CUtensorMap q, kv;
};

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// MN-major operand with an explicit distance between its 64-element chunks (see umma_desc_mn_sw128)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128_lbo(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(AT_THREADS, 1) cross_attention_tcgen05_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t at_smem_raw[];
  const uint32_t raw = smem_u32(at_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* aligned = at_smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(aligned + AT_BAR_OFF);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t k_full = bar0, k_empty = bar0 + 8, v_full = bar0 + 16, v_empty = bar0 + 24, s_full = bar0 + 32, s_empty = bar0 + 40,
                 p_full = bar0 + 48, o_full = bar0 + 56, o_empty = bar0 + 64;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.batches * p.heads;
  const int hd_chunks = (p.hd + 63) / 64;

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&maps.q);
    prefetch_tmap(&maps.kv);
    mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(s_empty, 4); mbar_init(p_full, 4); mbar_init(o_full, 1); mbar_init(o_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(AT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + AT_KEYS;

  if (warp == 4) {
    if (lane == 0) {
      // ===== producer =====
      uint32_t it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int b = item / p.heads, h = item - b * p.heads;
        const uint32_t ph = it & 1u;
        mbar_wait(k_empty, ph ^ 1u);
        mbar_expect_tx(k_full, uint32_t(hd_chunks) * (AT_Q_CHUNK + AT_KV_CHUNK));
        for (int c = 0; c < hd_chunks; ++c) {
          tma_load_4d(&maps.q, k_full, base + AT_Q_OFF + c * AT_Q_CHUNK, c * 64, h, 0, p.q_per_batch ? b : 0);
          tma_load_4d(&maps.kv, k_full, base + AT_K_OFF + c * AT_KV_CHUNK, c * 64, h, 0, b);
        }
        mbar_wait(v_empty, ph ^ 1u);
        mbar_expect_tx(v_full, uint32_t(hd_chunks) * AT_KV_CHUNK);
        for (int c = 0; c < hd_chunks; ++c) tma_load_4d(&maps.kv, v_full, base + AT_V_OFF + c * AT_KV_CHUNK, c * 64, p.heads + h, 0, b);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const int n_keys16 = (p.n_kv + 15) & ~15;  // UMMA N of S and the k extent of P V
      const uint32_t idesc_s = umma_idesc_bf16(AT_M, n_keys16, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(AT_M, p.hd, 0, 1);  // B = V_h MN-major
      uint32_t it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1u;
        mbar_wait(k_full, ph);
        mbar_wait(s_empty, ph ^ 1u);
        tc_fence_after();
        for (int ks = 0; ks < p.hd / UMMA_K; ++ks) {
          const uint64_t a = umma_desc_sw128(base + AT_Q_OFF + (ks >> 2) * AT_Q_CHUNK) + 2 * (ks & 3);
          const uint64_t bd = umma_desc_sw128(base + AT_K_OFF + (ks >> 2) * AT_KV_CHUNK) + 2 * (ks & 3);
          umma_bf16(tmem_S, a, bd, idesc_s, ks != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        umma_commit(k_empty);  // Q_h / K_h may be overwritten by the next item's loads
        mbar_wait(p_full, ph);
        mbar_wait(v_full, ph);
        mbar_wait(o_empty, ph ^ 1u);
        tc_fence_after();
        for (int ks = 0; ks < n_keys16 / UMMA_K; ++ks) {
          const uint64_t a = umma_desc_sw128(base + AT_P_OFF + (ks >> 2) * AT_P_CHUNK) + 2 * (ks & 3);
          const uint64_t bd = umma_desc_mn_sw128_lbo(base + AT_V_OFF, AT_KV_CHUNK) + uint64_t(ks) * (2048u >> 4);
          umma_bf16(tmem_O, a, bd, idesc_o, ks != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(v_empty);
      }
    }
  } else {
    // ===== softmax + output (warp w owns TMEM lanes 32 w .. 32 w + 31 = query rows) =====
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t p_row = base + AT_P_OFF + uint32_t(row) * 128u;
    const uint32_t sw = uint32_t(row) & 7u;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int b = item / p.heads, h = item - b * p.heads;
      const uint32_t ph = it & 1u;
      mbar_wait(s_full, ph);
      tc_fence_after();
      float m = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < AT_KEYS / 32; ++c) {
        if (c * 32 >= p.n_kv) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < p.n_kv) m = fmaxf(m, __uint_as_float(v[i]));
      }
      const float mo = m * p.scale_log2e;
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < AT_KEYS / 32; ++c) {
        if (c * 32 >= ((p.n_kv + 15) & ~15)) break;  // the second MMA reads P up to the next multiple of 16 keys only
        uint32_t v[32];
        const bool live = c * 32 < p.n_kv;  // warp-uniform
        if (live) {
          tmem_ld32(tmem_S + lane_addr + c * 32, v);
          tmem_ld_wait();
        }
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float e0 = 0.f, e1 = 0.f;
          if (live && c * 32 + i < p.n_kv) e0 = exp2f(fmaf(__uint_as_float(v[i]), p.scale_log2e, -mo));
          if (live && c * 32 + i + 1 < p.n_kv) e1 = exp2f(fmaf(__uint_as_float(v[i + 1]), p.scale_log2e, -mo));
          sum += e0 + e1;
          packed[i >> 1] = pack_bf16x2(e0, e1);
        }
        // 32 keys = 64 bytes = four 16-byte chunks of this row inside the [128 x 64-key] P chunk c / 2
        const uint32_t chunk_base = p_row + uint32_t(c >> 1) * AT_P_CHUNK;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t c16 = uint32_t((c & 1) * 4 + j);
          sts_v4(chunk_base + ((c16 ^ sw) << 4), make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]));
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of P -> visible to the tensor core's shared-memory reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(p_full);
        mbar_arrive(s_empty);  // this warp is done with its rows of S
      }
      const float inv = 1.0f / sum;
      mbar_wait(o_full, ph);
      tc_fence_after();
      __nv_bfloat16* orow = p.out + ((long long)b * p.n_q + row) * p.ldo + (long long)h * p.hd;
#pragma unroll 1
      for (int c = 0; c < p.hd / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_O + lane_addr + c * 32, v);
        tmem_ld_wait();
        if (row < p.n_q) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(v[8 * j + i]) * inv;
            *reinterpret_cast<uint4*>(orow + c * 32 + 8 * j) = Vec16<__nv_bfloat16>::pack(o);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(AT_TMEM_COLS) : "memory");
  }
}

bool attention_tcgen05_supported(int n_q, int n_kv, int heads, int hd, long long ldq, long long q_batch_stride, long long ldkv, long long ldo) {
  const char* e = getenv("MERV_ATTN_IMPL");
  if (e != nullptr && strcmp(e, "simt") == 0) return false;
  (void)heads;
  return n_q <= AT_M && n_kv <= AT_KEYS && hd % 32 == 0 && hd <= AT_MAX_HD && ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0 && q_batch_stride % 8 == 0;
}

int launch_attention_tcgen05(const void* q, long long ldq, long long q_batch_stride, const void* kv, long long ldkv, void* out, long long ldo,
                             int batches, int n_q, int n_kv, int heads, int hd, float scale, cudaStream_t stream) {
  MERV_REQUIRE(aligned16(q) && aligned16(kv) && aligned16(out), MERV_E_ALIGN, "cross attention: operands must be 16-byte aligned");
  AttnMaps maps;
  AttnParams p = {};
  p.out = static_cast<__nv_bfloat16*>(out); p.ldo = ldo;
  p.batches = batches; p.heads = heads; p.n_q = n_q; p.n_kv = n_kv; p.hd = hd;
  p.q_per_batch = q_batch_stride != 0 ? 1 : 0;
  p.scale_log2e = scale * 1.4426950408889634f;
  {
    // Q as (head dim, head, query row, frame); rows past n_q and head dims past hd are zero-filled by TMA
    const unsigned long long dims[4] = {(unsigned long long)hd, (unsigned long long)heads, (unsigned long long)n_q,
                                        (unsigned long long)(p.q_per_batch ? batches : 1)};
    const unsigned long long strides[3] = {(unsigned long long)hd * 2, (unsigned long long)ldq * 2,
                                           (unsigned long long)(p.q_per_batch ? q_batch_stride : (long long)n_q * ldq) * 2};
    const unsigned box[4] = {64, 1, AT_M, 1};
    if (int rc = encode_tmap_cached(&maps.q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, q, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  {
    // kv rows [K(C) | V(C)] as (head dim, 2 x heads, key, frame): head h of K is "head" h, of V "head" heads + h
    const unsigned long long dims[4] = {(unsigned long long)hd, (unsigned long long)(2 * heads), (unsigned long long)n_kv, (unsigned long long)batches};
    const unsigned long long strides[3] = {(unsigned long long)hd * 2, (unsigned long long)ldkv * 2, (unsigned long long)n_kv * ldkv * 2};
    const unsigned box[4] = {64, 1, AT_KEYS, 1};
    if (int rc = encode_tmap_cached(&maps.kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, kv, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  static const cudaError_t attr_rc =
      cudaFuncSetAttribute(cross_attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
  MERV_REQUIRE(attr_rc == cudaSuccess, MERV_E_CUDA, "cudaFuncSetAttribute(max dynamic smem=%d) failed: %s", AT_SMEM, cudaGetErrorString(attr_rc));
  const long long items = (long long)batches * heads;
  const int sms = sm_count();
  const unsigned grid = unsigned(items < sms ? items : sms);
  cross_attention_tcgen05_kernel<<<grid, AT_THREADS, AT_SMEM, stream>>>(maps, p);
  MERV_CUDA_OK(cudaGetLastError());
  return MERV_OK;
}

}  // namespace merv
