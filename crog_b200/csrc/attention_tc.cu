// tcgen05 flash attention for the dense (unmasked) attentions of the forward: decoder self-attention
// (676 tokens, 8 heads; model/layers.py:313-318) and the attention pool (169 tokens, 32 heads;
// model/clip.py:123-140).  head_dim = 64, bf16 operands, fp32 softmax state.
//
// One CTA = 128 queries of one (sample, head); key tiles of 128.
//   warp 0       TMA: Q once, then K_j / V_j boxes ([128 rows x 64 ch], SWIZZLE_128B) straight out of the packed
//                qkv activation matrix (column offset = head * 64)
//   warp 1       MMA: S = Q K_j^T  (M=128, N=128, K=64; both operands K-major)   -> TMEM cols [0,128)
//                     O_j = P V_j  (M=128, N=64, K=128; A = P K-major from smem, B = V MN-major) -> TMEM cols [128,192)
//   warps 2..5   softmax: thread = query row; two passes over the S tile in TMEM (row max, then exp2 / sum),
//                P written as bf16 into shared memory in the SWIZZLE_128B K-major layout the tensor core reads,
//                running (max, sum) and the fp32 output row kept in registers: acc = acc * corr + O_j.
// Two CTAs are co-resident per SM (80 KB smem, 256 TMEM columns each), which overlaps one CTA's softmax with the
// other's MMAs without double-buffering S.
#include "tc_common.cuh"

namespace {

constexpr int QT = 128, KT = 128, HD = 64;
constexpr int SM_Q = 0, SM_K = 16384, SM_V = 32768, SM_P = 49152, SM_BAR = 81920, SM_TOTAL = SM_BAR + 128 + 1024;
constexpr int TMEM_COLS = 256, TM_S = 0, TM_O = 128;

__global__ void __launch_bounds__(192, 2) attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                               const __grid_constant__ CUtensorMap tmK,
                                                               const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ o,
                                                               int ldo, int Tq, int Tk, float scale_log2e,
                                                               const int64_t* __restrict__ pad_word) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 96);
  const uint32_t b_qfull = smem_u32(bars + 0), b_kfull = smem_u32(bars + 1), b_vfull = smem_u32(bars + 2),
                 b_kempty = smem_u32(bars + 3), b_vempty = smem_u32(bars + 4), b_sfull = smem_u32(bars + 5),
                 b_pfull = smem_u32(bars + 6), b_ofull = smem_u32(bars + 7);
  pdl_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const int ntiles = (Tk + KT - 1) / KT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(b_qfull, 1); mbar_init(b_kfull, 1); mbar_init(b_vfull, 1); mbar_init(b_kempty, 1); mbar_init(b_vempty, 1);
    mbar_init(b_sfull, 1); mbar_init(b_pfull, 128); mbar_init(b_ofull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tc_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  const uint32_t sQ = smem_u32(smem + SM_Q), sK = smem_u32(smem + SM_K), sV = smem_u32(smem + SM_V), sP = smem_u32(smem + SM_P);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(b_qfull, QT * HD * 2);
      tma_load_2d(sQ, &tmQ, h * HD, b * Tq + q0, b_qfull);
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait(b_kempty, (j & 1) ^ 1);
        mbar_expect_tx(b_kfull, KT * HD * 2);
        tma_load_2d(sK, &tmK, h * HD, b * Tk + j * KT, b_kfull);
        mbar_wait(b_vempty, (j & 1) ^ 1);
        mbar_expect_tx(b_vfull, KT * HD * 2);
        tma_load_2d(sV, &tmV, h * HD, b * Tk + j * KT, b_vfull);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(QT, KT);            // S: N = 128 keys
      constexpr uint32_t idesc_o = make_idesc(QT, HD, 0, 1);      // O: N = 64, B (= V) MN-major
      mbar_wait(b_qfull, 0);
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait(b_kfull, j & 1);
        tc_fence_after();
        const uint64_t dq = make_sdesc(sQ), dk = make_sdesc(sK);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_bf16(tmem + TM_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        tc_commit(b_kempty);
        tc_commit(b_sfull);
        mbar_wait(b_pfull, j & 1);  // P_j in smem (and S_j consumed)
        mbar_wait(b_vfull, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < KT / 16; ++k) {
          // A = P: two 64-key K-major atoms of 16 KB; B = V: 16 key rows per step = 2048 B
          const uint64_t dp = make_sdesc(sP + (k >> 2) * 16384) + 2 * (k & 3);
          const uint64_t dv = make_sdesc(sV + k * 2048, 16, 1024);
          tc_mma_bf16(tmem + TM_O, dp, dv, idesc_o, k != 0);
        }
        tc_commit(b_vempty);
        tc_commit(b_ofull);
      }
    }
  } else {
    const int qd = warp & 3, row = qd * 32 + lane;
    const uint32_t t_s = tmem + ((uint32_t)(qd * 32) << 16) + TM_S, t_o = tmem + ((uint32_t)(qd * 32) << 16) + TM_O;
    float acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
    float mx = -INFINITY, den = 0.f;
    uint8_t* prow = smem + SM_P + row * 128;
    const int sw = row & 7;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(b_sfull, j & 1);
      tc_fence_after();
      const int kvalid = Tk - j * KT;  // keys >= kvalid belong to the next sample / padding
      const bool full = kvalid >= KT && pad_word == nullptr;  // CTA-uniform: only the last key tile / padded words need masking
      uint32_t km[4] = {0u, 0u, 0u, 0u};  // bit i of km[c / 32]: key c + i takes part (in range and not a padding word)
      if (!full) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int key = cc * 32 + lane;
          bool ok = key < kvalid;
          if (ok && pad_word) ok = pad_word[(long long)b * Tk + j * KT + key] != 0;  // key_padding_mask = (word == 0)
          km[cc] = __ballot_sync(0xffffffffu, ok);
        }
      }
      float tmax = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < KT; c += 32) {
        if (c >= kvalid) break;
        uint32_t r[32];
        tc_ld32(t_s + c, r);
        tc_wait_ld();
        if (full || (c + 32 <= kvalid && pad_word == nullptr)) {
          float m0 = -INFINITY, m1 = -INFINITY;  // two independent FMNMX3 chains
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            m0 = fmaxf(fmaxf(m0, __uint_as_float(r[i])), __uint_as_float(r[i + 1]));
            m1 = fmaxf(fmaxf(m1, __uint_as_float(r[i + 2])), __uint_as_float(r[i + 3]));
          }
          tmax = fmaxf(fmaxf(tmax, m0), m1);
        } else {
          const uint32_t m = km[c >> 5];
#pragma unroll
          for (int i = 0; i < 32; ++i) if ((m >> i) & 1u) tmax = fmaxf(tmax, __uint_as_float(r[i]));
        }
      }
      const float nm = fmaxf(mx, tmax);
      const float corr = exp2f((mx - nm) * scale_log2e);
      const float nms = nm * scale_log2e;
      float psum = 0.f;
#pragma unroll 1
      for (int c = 0; c < KT; c += 32) {
        uint32_t pk[16];
        if (c < kvalid) {
          uint32_t r[32];
          tc_ld32(t_s + c, r);
          tc_wait_ld();
          if (full || (c + 32 <= kvalid && pad_word == nullptr)) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = exp2f(__uint_as_float(r[i]) * scale_log2e - nms);
              const float p1 = exp2f(__uint_as_float(r[i + 1]) * scale_log2e - nms);
              __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
              s0 += p0; s1 += p1;
              pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
            }
            psum += s0 + s1;
          } else {
            const uint32_t m = km[c >> 5];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = ((m >> i) & 1u) ? exp2f(__uint_as_float(r[i]) * scale_log2e - nms) : 0.f;
              const float p1 = ((m >> (i + 1)) & 1u) ? exp2f(__uint_as_float(r[i + 1]) * scale_log2e - nms) : 0.f;
              __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
              psum += p0 + p1;
              pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;  // padding keys: P = 0
        }
        // keys c..c+31 -> atom (c / 64), 16-byte units ((c % 64) / 8 + u), swizzled with the row
        uint8_t* base = prow + (c >> 6) * 16384;
        const int u0 = (c & 63) >> 3;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<uint4*>(base + (((u0 + u) ^ sw) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
      }
      den = den * corr + psum;
      mx = nm;
      tc_fence_before();
      fence_proxy_async_smem();  // generic-proxy writes of P -> visible to the tensor core (async proxy)
      mbar_arrive(b_pfull);
      mbar_wait(b_ofull, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t r[32];
        tc_ld32(t_o + c, r);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[c + i] = acc[c + i] * corr + __uint_as_float(r[i]);
      }
      tc_fence_before();
    }
    if (q0 + row < Tq) {
      const float inv = 1.f / den;
      bf16* op = o + ((long long)b * Tq + q0 + row) * ldo + h * HD;
#pragma unroll
      for (int d = 0; d < HD; d += 8) {
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = acc[d + i] * inv;
        store8(op + d, t);
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem, TMEM_COLS);
  }
}

}  // namespace

int crog_attention_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int B, int heads,
                      int Tq, int Tk, float scale, const int64_t* pad_word, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    CROG_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    attr = true;
  }
  CUtensorMap tmQ, tmK, tmV;
  int rc = crog_encode_2d_bf16(&tmQ, q, (uint64_t)heads * HD, (uint64_t)B * Tq, (uint64_t)ldq, QT);
  if (rc) return rc;
  rc = crog_encode_2d_bf16(&tmK, k, (uint64_t)heads * HD, (uint64_t)B * Tk, (uint64_t)ldk, KT);
  if (rc) return rc;
  rc = crog_encode_2d_bf16(&tmV, v, (uint64_t)heads * HD, (uint64_t)B * Tk, (uint64_t)ldv, KT);
  if (rc) return rc;
  dim3 grid((Tq + QT - 1) / QT, heads, B);
  crog_launch(attention_tc_kernel, grid, dim3(192), SM_TOTAL, stream, tmQ, tmK, tmV, (bf16*)o, ldo, Tq, Tk, scale * 1.4426950408889634f, pad_word);
  CROG_LAUNCH_OK("attention_tc");
  return CROG_OK;
}
