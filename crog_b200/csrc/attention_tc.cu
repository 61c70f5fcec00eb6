// tcgen05 flash attention for the dense (unmasked) attentions of the forward: decoder self-attention
// (676 tokens, 8 heads; model/layers.py:313-318) and the attention pool (169 tokens, 32 heads;
// model/clip.py:123-140).  head_dim = 64, bf16 operands, fp32 softmax state.
//
// One CTA = 128 queries of one (sample, head); key tiles of 128.  256 threads = two warpgroups; setmaxnreg moves
// registers from the producer warpgroup (warps 0-3) to the softmax warpgroup (warps 4-7).
//   warp 0       TMA: Q once, then K_j / V_j boxes ([128 rows x 64 ch], SWIZZLE_128B) straight out of the packed
//                qkv activation matrix (column offset = head * 64)
//   warp 1       MMA: S = Q K_j^T  (M=128, N=128, K=64; both operands K-major)   -> TMEM cols [0,128)
//                     O += P V_j   (M=128, N=64, K=128; A = P K-major from smem, B = V MN-major) -> TMEM cols [128,192);
//                     a ragged last key tile contracts only over its 16-key steps that hold a valid key
//   warps 4..7   softmax: thread = query row; the whole 128-key S row is read from TMEM once (four tcgen05.ld in flight),
//                row max, one MUFU.EX2 per weight, P written as bf16 into shared memory in the SWIZZLE_128B K-major
//                layout the tensor core reads.  The fp32 output row stays in TMEM and is rescaled in place only when a
//                row maximum has grown by more than 2^8 (lazy rescaling).
// Two CTAs are co-resident per SM (80 KB smem, 256 TMEM columns each), which overlaps one CTA's softmax with the
// other's MMAs and TMEM loads.  Variants measured on B200 and rejected (profiles/README.md, attention log): eight softmax
// warps splitting the keys in halves, one CTA per SM with double-buffered S / K / V / P, software-pipelined S loads,
// and a polynomial exp2 on the FMA pipe for half of the weights.
#include "tc_common.cuh"

namespace {

constexpr int QT = 128, KT = 128, HD = 64;
constexpr int SM_Q = 0, SM_K = 16384, SM_V = 32768, SM_P = 49152, SM_BAR = 81920, SM_TOTAL = SM_BAR + 128 + 1024;
constexpr int TMEM_COLS = 256, TM_S = 0, TM_O = 128;

// one MUFU.EX2 (exp2f without -use_fast_math adds a range test and two predicated multiplies per element for
// denormal results; softmax weights below 2^-126 may flush to zero)
// 16-key contraction steps of the P V MMA for a tile with `kv` keys left (>= 1)
__device__ __forceinline__ int pv_ksteps(int kv) { return kv >= KT ? KT / 16 : (kv + 15) / 16; }
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(256, 2) attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
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

  if (warp < 4) {
    // producer warpgroup (TMA warp, MMA warp, two idle warps): hand its registers to the softmax warpgroup
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
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
        const int ksteps = pv_ksteps(Tk - j * KT);  // a ragged last tile contracts over its valid 16-key steps only
#pragma unroll
        for (int k = 0; k < KT / 16; ++k) {
          if (k >= ksteps) break;
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
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");  // 2 CTAs x (128 x 40 + 128 x 208) registers fit one SM
    // Four softmax warps, thread = query row.  The output row lives in TMEM, not in registers: P V_j accumulates onto
    // it on the tensor core, and it is rescaled there (tcgen05.ld -> FMUL -> tcgen05.st) only when the running maximum
    // of some row of the warp has grown by more than 2^RESCALE_LOG2 since the reference maximum `mx` was taken;
    // otherwise the (slightly stale) `mx` stays the reference, P <= 2^RESCALE_LOG2 and nothing is touched.  That frees
    // the registers to hold the whole 128-key S row, so the tile is read from TMEM once, all four loads in flight.
    constexpr float RESCALE_LOG2 = 8.f;
    const int qd = warp & 3, row = qd * 32 + lane;
    const uint32_t t_s = tmem + ((uint32_t)(qd * 32) << 16) + TM_S, t_o = tmem + ((uint32_t)(qd * 32) << 16) + TM_O;
    float mx = -INFINITY, den = 0.f;
    uint8_t* prow = smem + SM_P + row * 128;
    const int sw = row & 7;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(b_sfull, j & 1);
      tc_fence_after();
      const int kvalid = Tk - j * KT;             // keys >= kvalid belong to the next sample / padding
      const int kneed = pv_ksteps(kvalid) * 16;   // keys the P V MMA of this tile contracts over (whole 16-key steps)
      uint32_t r[4][32], km[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c * 32 < kneed) tc_ld32(t_s + c * 32, r[c]);
      tc_wait_ld();
      const bool full = kvalid >= KT && pad_word == nullptr;  // CTA-uniform; one straight-line block for the common case
      float m0 = -INFINITY, m1 = -INFINITY;  // two independent FMNMX3 chains
      if (full) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            m0 = fmaxf(fmaxf(m0, __uint_as_float(r[c][i])), __uint_as_float(r[c][i + 1]));
            m1 = fmaxf(fmaxf(m1, __uint_as_float(r[c][i + 2])), __uint_as_float(r[c][i + 3]));
          }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          km[c] = 0xffffffffu;  // bit i: key 32 c + i takes part
          if (c * 32 >= kneed) continue;
          if (c * 32 + 32 <= kvalid && pad_word == nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              m0 = fmaxf(fmaxf(m0, __uint_as_float(r[c][i])), __uint_as_float(r[c][i + 1]));
              m1 = fmaxf(fmaxf(m1, __uint_as_float(r[c][i + 2])), __uint_as_float(r[c][i + 3]));
            }
          } else {
            const int key = c * 32 + lane;
            bool ok = key < kvalid;
            if (ok && pad_word) ok = pad_word[(long long)b * Tk + j * KT + key] != 0;  // key_padding_mask = (word == 0)
            km[c] = __ballot_sync(0xffffffffu, ok);
#pragma unroll
            for (int i = 0; i < 32; ++i) if ((km[c] >> i) & 1u) m0 = fmaxf(m0, __uint_as_float(r[c][i]));
          }
        }
      }
      const float tmax = fmaxf(m0, m1);
      if (j == 0) {
        mx = tmax;  // nothing accumulated yet: P V_0 overwrites the output row
      } else {
        const bool need = (tmax - mx) * scale_log2e > RESCALE_LOG2;
        if (__any_sync(0xffffffffu, need)) {
          float corr = 1.f;
          if (need) { corr = exp2f((mx - tmax) * scale_log2e); mx = tmax; den *= corr; }
          mbar_wait(b_ofull, (j - 1) & 1);  // P V_{j-1} has landed in the output row
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < HD; c += 32) {
            uint32_t o32[32];
            tc_ld32(t_o + c, o32);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o32[i] = __float_as_uint(__uint_as_float(o32[i]) * corr);
            tc_st32(t_o + c, o32);
          }
          tc_wait_st();
        }
      }
      const float nms = mx * scale_log2e;
      float s0 = 0.f, s1 = 0.f;
      auto store_chunk = [&](int c, const uint32_t (&pk)[16]) {
        // keys 32c..32c+31 -> atom (c / 2), 16-byte units (4 (c % 2) + u), swizzled with the row
        uint8_t* base = prow + (c >> 1) * 16384;
        const int u0 = (c & 1) * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<uint4*>(base + (((u0 + u) ^ sw) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
      };
      if (full) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2_ftz(__uint_as_float(r[c][i]) * scale_log2e - nms);
            const float p1 = ex2_ftz(__uint_as_float(r[c][i + 1]) * scale_log2e - nms);
            __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
            s0 += p0; s1 += p1;
            pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
          }
          store_chunk(c, pk);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 32 >= kneed) continue;
          uint32_t pk[16];
          const uint32_t m = km[c];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ((m >> i) & 1u) ? ex2_ftz(__uint_as_float(r[c][i]) * scale_log2e - nms) : 0.f;
            const float p1 = ((m >> (i + 1)) & 1u) ? ex2_ftz(__uint_as_float(r[c][i + 1]) * scale_log2e - nms) : 0.f;
            __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
            s0 += p0; s1 += p1;
            pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
          }
          store_chunk(c, pk);
        }
      }
      den += s0 + s1;
      tc_fence_before();
      fence_proxy_async_smem();  // generic-proxy writes of P -> visible to the tensor core (async proxy)
      mbar_arrive(b_pfull);
    }
    mbar_wait(b_ofull, (ntiles - 1) & 1);
    tc_fence_after();
    const float inv = 1.f / den;
    const bool live = q0 + row < Tq;
    bf16* op = o + ((long long)b * Tq + q0 + row) * ldo + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 32) {
      uint32_t o32[32];
      tc_ld32(t_o + c, o32);
      tc_wait_ld();
      if (live) {
#pragma unroll
        for (int d = 0; d < 32; d += 8) {
          float t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = __uint_as_float(o32[d + i]) * inv;
          store8(op + c + d, t);
        }
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
  static DeviceOnce once;
  int dev_;
  if (once.need(&dev_)) {
    CROG_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    once.done(dev_);
  }
  CUtensorMap tmQ, tmK, tmV;
  int rc = crog_encode_2d_bf16(&tmQ, q, (uint64_t)heads * HD, (uint64_t)B * Tq, (uint64_t)ldq, QT);
  if (rc) return rc;
  rc = crog_encode_2d_bf16(&tmK, k, (uint64_t)heads * HD, (uint64_t)B * Tk, (uint64_t)ldk, KT);
  if (rc) return rc;
  rc = crog_encode_2d_bf16(&tmV, v, (uint64_t)heads * HD, (uint64_t)B * Tk, (uint64_t)ldv, KT);
  if (rc) return rc;
  dim3 grid((Tq + QT - 1) / QT, heads, B);
  crog_launch(attention_tc_kernel, grid, dim3(256), SM_TOTAL, stream, tmQ, tmK, tmV, (bf16*)o, ldo, Tq, Tk, scale * 1.4426950408889634f, pad_word);
  CROG_LAUNCH_OK("attention_tc");
  return CROG_OK;
}
