// tcgen05 flash attention for the dense (unmasked) attentions of the forward: decoder self-attention
// (676 tokens, 8 heads; model/layers.py:313-318) and the attention pool (169 tokens, 32 heads;
// model/clip.py:123-140).  head_dim = 64, bf16 operands, fp32 softmax state.
//
// One CTA = 128 queries of one (sample, head); key tiles of 128.
// 384 threads = three warpgroups; setmaxnreg moves registers from the producer warpgroup (warps 0-3) to the softmax ones (4-11).
//   warp 0       TMA: Q once, then K_j / V_j boxes ([128 rows x 64 ch], SWIZZLE_128B) straight out of the packed
//                qkv activation matrix (column offset = head * 64)
//   warp 1       MMA: S = Q K_j^T  (M=128, N=128, K=64; both operands K-major)   -> TMEM cols [0,128)
//                     O_j = P V_j  (M=128, N=64, K=128; A = P K-major from smem, B = V MN-major) -> TMEM cols [128,192)
//   warps 4..11  softmax: thread = query row x key half; its 64 keys of the S row are read from TMEM once,
//                row max (exchanged with the warp holding the other half), exp2 / sum, P written as bf16 into shared memory in the SWIZZLE_128B K-major layout the
//                tensor core reads.  The fp32 output row stays in TMEM (P V_j accumulates onto it) and is rescaled in
//                place only when a row maximum has grown by more than 2^8 (lazy rescaling): the per-tile
//                "read O, acc = acc * corr + O_j" round trip and its 64 registers are gone.
// Two CTAs are co-resident per SM (80 KB smem, 256 TMEM columns each), which overlaps one CTA's softmax with the
// other's MMAs without double-buffering S.
#include "tc_common.cuh"

namespace {

constexpr int QT = 128, KT = 128, HD = 64;
constexpr int SM_Q = 0, SM_K = 16384, SM_V = 32768, SM_P = 49152, SM_BAR = 81920, SM_XCH = SM_BAR + 128, SM_TOTAL = SM_XCH + 2048 + 1024;
constexpr int TMEM_COLS = 256, TM_S = 0, TM_O = 128;

// one MUFU.EX2 (exp2f without -use_fast_math adds a range test and two predicated multiplies per element for
// denormal results; softmax weights below 2^-126 may flush to zero)
// 64-thread named barrier of the two softmax warps that share TMEM lane quarter qd (ids 1..4; 0 is __syncthreads)
__device__ __forceinline__ void pair_sync(int qd) {
  switch (qd) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(384, 2) attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
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
    mbar_init(b_sfull, 1); mbar_init(b_pfull, 256); mbar_init(b_ofull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tc_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  const uint32_t sQ = smem_u32(smem + SM_Q), sK = smem_u32(smem + SM_K), sV = smem_u32(smem + SM_V), sP = smem_u32(smem + SM_P);

  if (warp < 4) {
    // producer warpgroup (TMA warp, MMA warp, two idle warps): hand its registers to the softmax warpgroup
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
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
          tc_mma_bf16(tmem + TM_O, dp, dv, idesc_o, (j | k) != 0);  // accumulate across key tiles
        }
        tc_commit(b_vempty);
        tc_commit(b_ofull);
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");  // 128 x 32 + 256 x 104 <= 384 x 80 registers of the CTA
    // Eight softmax warps: warps (4 + qd) and (8 + qd) share the 32 query rows of TMEM lane quarter qd and split the
    // 128 keys of a tile in halves (half = 0: keys 0..63, half = 1: keys 64..127), so four warps per scheduler hide the
    // MUFU / TMEM latencies.  Per tile the pair exchanges its partial row maxima through shared memory (one 64-thread
    // named barrier); partial row sums are only combined at the end.
    // The output row lives in TMEM, not in registers: P V_j accumulates onto it on the tensor core, and it is rescaled
    // there (tcgen05.ld -> FMUL -> tcgen05.st, each warp its 32 columns) only when the running maximum of some row of
    // the warp has grown by more than 2^RESCALE_LOG2 since the reference maximum `mx` was taken; otherwise the
    // (slightly stale) `mx` stays the reference, P <= 2^RESCALE_LOG2 and nothing is touched.
    constexpr float RESCALE_LOG2 = 8.f;
    const int qd = warp & 3, half = (warp - 4) >> 2, row = qd * 32 + lane;
    const uint32_t t_s = tmem + ((uint32_t)(qd * 32) << 16) + TM_S + half * 64;
    const uint32_t t_o = tmem + ((uint32_t)(qd * 32) << 16) + TM_O + half * 32;
    float* xch = reinterpret_cast<float*>(smem + SM_XCH);  // [2][128] partial maxima / sums
    float mx = -INFINITY, den = 0.f;
    uint8_t* prow = smem + SM_P + half * 16384 + row * 128;  // this warp's 64 keys are one 16 KB K-major atom of P
    const int sw = row & 7;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(b_sfull, j & 1);
      tc_fence_after();
      const int kvalid = Tk - j * KT - half * 64;  // of this warp's 64 keys; keys >= kvalid belong to the next sample / padding
      const bool full = kvalid >= 64 && pad_word == nullptr;  // warp-uniform: only the last key tile / padded words need masking
      uint32_t km[2] = {0xffffffffu, 0xffffffffu};  // bit i of km[c]: key 32 c + i of this half takes part
      if (!full) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int key = cc * 32 + lane;
          bool ok = key < kvalid;
          if (ok && pad_word) ok = pad_word[(long long)b * Tk + j * KT + half * 64 + key] != 0;  // key_padding_mask = (word == 0)
          km[cc] = __ballot_sync(0xffffffffu, ok);
        }
      }
      uint32_t r[2][32];
#pragma unroll
      for (int c = 0; c < 2; ++c)
        if (c * 32 < kvalid) tc_ld32(t_s + c * 32, r[c]);
      tc_wait_ld();
      float tmax = -INFINITY;
      if (full) {
        float m0 = -INFINITY, m1 = -INFINITY;  // two independent FMNMX3 chains
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            m0 = fmaxf(fmaxf(m0, __uint_as_float(r[c][i])), __uint_as_float(r[c][i + 1]));
            m1 = fmaxf(fmaxf(m1, __uint_as_float(r[c][i + 2])), __uint_as_float(r[c][i + 3]));
          }
        tmax = fmaxf(m0, m1);
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c * 32 < kvalid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) if ((km[c] >> i) & 1u) tmax = fmaxf(tmax, __uint_as_float(r[c][i]));
          }
      }
      // row maximum over both halves (the partner's next write of its slot comes after it has passed this tile's P
      // barrier, which needs this thread's arrival, i.e. after this read)
      xch[half * 128 + row] = tmax;
      pair_sync(qd);
      tmax = fmaxf(tmax, xch[(half ^ 1) * 128 + row]);
      if (j == 0) {
        mx = tmax;  // nothing accumulated yet: P V_0 overwrites the output row
      } else {
        const bool need = (tmax - mx) * scale_log2e > RESCALE_LOG2;
        if (__any_sync(0xffffffffu, need)) {  // same rows, same maxima: both warps of the pair take the same branch
          float corr = 1.f;
          if (need) { corr = exp2f((mx - tmax) * scale_log2e); mx = tmax; den *= corr; }
          mbar_wait(b_ofull, (j - 1) & 1);  // P V_{j-1} has landed in the output row
          tc_fence_after();
          uint32_t o32[32];
          tc_ld32(t_o, o32);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o32[i] = __float_as_uint(__uint_as_float(o32[i]) * corr);
          tc_st32(t_o, o32);
          tc_wait_st();
        }
      }
      const float nms = mx * scale_log2e;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
        if (c * 32 < kvalid) {
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = ex2_ftz(__uint_as_float(r[c][i]) * scale_log2e - nms);
              const float p1 = ex2_ftz(__uint_as_float(r[c][i + 1]) * scale_log2e - nms);
              __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
              s0 += p0; s1 += p1;
              pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
            }
          } else {
            const uint32_t m = km[c];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = ((m >> i) & 1u) ? ex2_ftz(__uint_as_float(r[c][i]) * scale_log2e - nms) : 0.f;
              const float p1 = ((m >> (i + 1)) & 1u) ? ex2_ftz(__uint_as_float(r[c][i + 1]) * scale_log2e - nms) : 0.f;
              __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
              s0 += p0; s1 += p1;
              pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;  // padding keys: P = 0
        }
        // keys 32c..32c+31 of this half -> 16-byte units (4 c + u) of the atom row, swizzled with the row
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<uint4*>(prow + (((4 * c + u) ^ sw) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
      }
      den += s0 + s1;
      tc_fence_before();
      fence_proxy_async_smem();  // generic-proxy writes of P -> visible to the tensor core (async proxy)
      mbar_arrive(b_pfull);
    }
    mbar_wait(b_ofull, (ntiles - 1) & 1);
    tc_fence_after();
    // total row sum = both halves' partial sums (relative to the same reference maximum)
    xch[256 + half * 128 + row] = den;
    pair_sync(qd);
    const float inv = 1.f / (den + xch[256 + (half ^ 1) * 128 + row]);
    uint32_t o32[32];
    tc_ld32(t_o, o32);
    tc_wait_ld();
    if (q0 + row < Tq) {
      bf16* op = o + ((long long)b * Tq + q0 + row) * ldo + h * HD + half * 32;
#pragma unroll
      for (int d = 0; d < 32; d += 8) {
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = __uint_as_float(o32[d + i]) * inv;
        store8(op + d, t);
      }
    }
    tc_fence_before();
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
  crog_launch(attention_tc_kernel, grid, dim3(384), SM_TOTAL, stream, tmQ, tmK, tmV, (bf16*)o, ldo, Tq, Tk, scale * 1.4426950408889634f, pad_word);
  CROG_LAUNCH_OK("attention_tc");
  return CROG_OK;
}
