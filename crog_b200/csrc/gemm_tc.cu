// tcgen05 / TMEM / TMA implicit-GEMM for bf16 activations and weights (sm_100a), persistent and
// warp-specialised.
//
// One CTA per SM loops over 128 x BN output tiles (static round-robin, N fastest so concurrent CTAs share A):
//   warp 0      TMA producer   (cp.async.bulk.tensor 2D, SWIZZLE_128B, mbarrier complete_tx) over a STAGES-deep ring
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16) into one of two TMEM
//                              accumulator buffers, so tile i+1 is being multiplied while tile i drains
//   warps 2..9  epilogue       each warp owns 32 accumulator rows (its TMEM lane quarter) x half of the columns:
//                phase 1  tcgen05.ld -> registers (thread = row) -> +addmat, scale/bias, act, gate -> fp32 staging in smem
//                         (after which the accumulator buffer is handed back to the MMA warp)
//                phase 2  staging -> (+residual) -> HBM with 16-byte accesses that are contiguous along a row, eight
//                         lanes per 128-byte row segment, all loads of a pass issued before use
// The K loop runs over (tap, 64-channel chunk); for a 3x3 convolution on the zero-haloed ("padded") NHWC layout
// tap t is the same activation matrix shifted by a constant number of rows, so the A tile of every k-step is one plain
// 2D TMA box at row (row0 + shift_t); rows outside the tensor are zero-filled by TMA.
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 elements = 128 bytes = one SWIZZLE_128B atom row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + EPI_WARPS * 32;

template <int BN, int STAGES>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + ((B_BYTES + 1023) / 1024) * 1024;
  static constexpr int HALF = BN >= 32 ? BN / 2 : BN;        // columns per epilogue warp
  static constexpr int SUB = HALF > 32 ? 32 : HALF;           // columns staged at a time
  static constexpr int PITCH = SUB * 4 + 16;                  // staging row pitch (bytes): 16B-phase conflict free
  static constexpr int STG_BYTES = 32 * PITCH;                // per warp
  static constexpr int STG_OFF = STAGES * STAGE_BYTES;
  static constexpr int SB_OFF = STG_OFF + EPI_WARPS * STG_BYTES;  // scale/bias: [2 acc stages][2][BN] floats
  static constexpr int BAR_OFF = SB_OFF + 2 * 2 * BN * 4;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;          // + alignment slack
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory"); }

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB,
                                                                   const CrogGemm g, int n_tiles, int total_tiles) {
  using L = Cfg<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B tiles; plain pointer arithmetic keeps the shared address space so the
  // epilogue's staging / scale / bias accesses compile to LDS / STS instead of generic loads
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);  // full[S], empty[S], tfull[2], tempty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::BAR_OFF + (2 * STAGES + 4) * 8);
  float* s_sb = reinterpret_cast<float*>(smem + L::SB_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kchunks = g.cin / BK;
  const int num_kb = g.taps * kchunks;
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull0 = smem_u32(bars + 2 * STAGES),
                 tempty0 = smem_u32(bars + 2 * STAGES + 2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull0 + 8 * s, 1); mbar_init(tempty0 + 8 * s, EPI_WARPS * 32); }
    fence_barrier_init();
  }
  if (warp == 1) tc_alloc(smem_u32(tmem_slot), L::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int kbg = 0;  // k-blocks issued so far (ring position)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_t = tile % n_tiles, m_t = tile / n_tiles;
        const TileRows tr = tile_rows(g, m_t, BM);
        const int wrow0 = n_t * BN + (g.w_sample_stride > 0 ? tr.wsample * (int)(g.w_sample_stride / ((long long)g.taps * g.cin)) : 0);
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const int s = kbg % STAGES, it = kbg / STAGES;
          mbar_wait(empty0 + 8 * s, (it & 1) ^ 1);
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES), sb = sa + L::A_BYTES;
          const int tap = kb / kchunks, c0 = (kb % kchunks) * BK;
          mbar_expect_tx(full0 + 8 * s, L::A_BYTES + L::B_BYTES);
          tma_load_2d(sa, &tmA, c0, (int)(tr.row0 + tap_shift(g.taps, tap, g.W)), full0 + 8 * s);
          tma_load_2d(sb, &tmB, kb * BK, wrow0, full0 + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      int kbg = 0, i = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++i) {
        const int as = i & 1;
        mbar_wait(tempty0 + 8 * as, ((i >> 1) & 1) ^ 1);  // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const int s = kbg % STAGES, it = kbg / STAGES;
          mbar_wait(full0 + 8 * s, it & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES), sb = sa + L::A_BYTES;
          const uint64_t da = make_sdesc(sa), db = make_sdesc(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            tc_mma_bf16(tacc, da + (uint64_t)(k * UMMA_K * 2 / 16), db + (uint64_t)(k * UMMA_K * 2 / 16), idesc, (kb | k) != 0);
          tc_commit(empty0 + 8 * s);  // frees the smem slot once these MMAs retire
        }
        tc_commit(tfull0 + 8 * as);  // accumulator complete
      }
    }
  } else {
    const int ew = warp - 2;                 // 0..7
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int hsel = ew >> 2;                // which half of the tile's columns
    const bool active = (BN >= 32) || hsel == 0;
    uint8_t* stg = smem + L::STG_OFF + ew * L::STG_BYTES;
    const int et = threadIdx.x - 64;         // 0..255
    constexpr int UPR = L::SUB / 8;          // 16-byte (8-column) units per staged row
    constexpr int RPP = 32 / UPR;            // rows written per warp pass
    constexpr int PASSES = 32 / RPP;
    int i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++i) {
      const int as = i & 1;
      const int n_t = tile % n_tiles, m_t = tile / n_tiles;
      const TileRows tr = tile_rows(g, m_t, BM);
      const int n0 = n_t * BN;
      float* sc = s_sb + as * 2 * BN;
      float* bi = sc + BN;
      for (int c = et; c < BN; c += EPI_WARPS * 32) {
        const int n = n0 + c;
        sc[c] = (g.scale && n < g.N) ? g.scale[n] : 1.f;
        bi[c] = (g.bias && n < g.N) ? g.bias[n] : 0.f;
      }
      epi_bar_sync();
      const RowMap m = map_row(g, tr.row0 + q * 32 + lane, tr.row_end);
      const int my_orow = m.valid ? m.orow : -1;
      // phase-2 ownership: lane -> (row p*RPP + rsub, 8-column unit u) of every staged sub-block
      const int u = lane % UPR, rsub = lane / UPR;
      int orow[PASSES];
#pragma unroll
      for (int p = 0; p < PASSES; ++p) orow[p] = __shfl_sync(0xffffffffu, my_orow, p * RPP + rsub);
      // bf16 residual tiles are fetched before the accumulator is even ready, so their HBM latency hides behind the
      // MMA of this tile and the drain of the previous one
      constexpr bool PREFETCH = (BN <= 128);
      constexpr int NSUB = L::HALF / L::SUB;
      uint4 rpre[PREFETCH ? NSUB * PASSES : 1];
      const bool res_bf16 = g.residual != nullptr && g.out_dtype == CROG_BF16;
      if constexpr (PREFETCH) {
        if (res_bf16 && active) {
#pragma unroll
          for (int sb = 0; sb < NSUB; ++sb) {
            const int ncol = n0 + hsel * L::HALF + sb * L::SUB + u * 8;
#pragma unroll
            for (int p = 0; p < PASSES; ++p)
              rpre[sb * PASSES + p] = (ncol < g.N && orow[p] >= 0)
                  ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(g.residual) + (long long)orow[p] * g.res_ld + ncol))
                  : make_uint4(0, 0, 0, 0);
          }
        }
      }
      mbar_wait(tfull0 + 8 * as, (i >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + hsel * L::HALF;
#pragma unroll
      for (int sbi = 0; sbi < NSUB; ++sbi) {
        const int sub = sbi * L::SUB;
        const int cbase = hsel * L::HALF + sub;  // column of the tile where this staged block starts
        if (active) {
          // ---- phase 1: accumulator -> registers -> epilogue math -> fp32 staging (thread = row)
          if constexpr (L::SUB >= 32) {
            uint32_t r[32];
            tc_ld32(taddr + sub, r);
            tc_wait_ld();
            float acc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
            if (m.valid && n0 + cbase < g.N) epilogue_math<32>(g, m, n0 + cbase, acc, sc + cbase, bi + cbase);
            float4* dst = reinterpret_cast<float4*>(stg + lane * L::PITCH);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
          } else {
            uint32_t r[16];
            tc_ld16(taddr + sub, r);
            tc_wait_ld();
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(r[j]);
            if (m.valid) epilogue_math<16>(g, m, n0 + cbase, acc, sc + cbase, bi + cbase);
            float4* dst = reinterpret_cast<float4*>(stg + lane * L::PITCH);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
          }
        }
        if (sbi == NSUB - 1) {  // last read of this accumulator buffer: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(tempty0 + 8 * as);
        }
        __syncwarp();
        if (active) {
          // ---- phase 2: staging -> (+residual) -> global, row-contiguous 16-byte accesses
          const int ncol = n0 + cbase + u * 8;
          const bool colok = ncol < g.N;
          float v[PASSES][8];
#pragma unroll
          for (int p = 0; p < PASSES; ++p) {
            const float4* src = reinterpret_cast<const float4*>(stg + (p * RPP + rsub) * L::PITCH + u * 32);
            const float4 a = src[0], b = src[1];
            v[p][0] = a.x; v[p][1] = a.y; v[p][2] = a.z; v[p][3] = a.w; v[p][4] = b.x; v[p][5] = b.y; v[p][6] = b.z; v[p][7] = b.w;
          }
          if (g.residual) {
            if (g.out_dtype == CROG_BF16) {
              uint4 rr[PASSES];
#pragma unroll
              for (int p = 0; p < PASSES; ++p) {
                if constexpr (PREFETCH) rr[p] = rpre[sbi * PASSES + p];
                else
                  rr[p] = (colok && orow[p] >= 0)
                              ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(g.residual) + (long long)orow[p] * g.res_ld + ncol))
                              : make_uint4(0, 0, 0, 0);
              }
#pragma unroll
              for (int p = 0; p < PASSES; ++p) {
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rr[p]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = __bfloat1622float2(h2[j]);
                  v[p][2 * j] += f.x; v[p][2 * j + 1] += f.y;
                }
              }
            } else {
              float4 ra[PASSES], rb[PASSES];
#pragma unroll
              for (int p = 0; p < PASSES; ++p) {
                if (colok && orow[p] >= 0) {
                  const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(g.residual) + (long long)orow[p] * g.res_ld + ncol);
                  ra[p] = rp[0]; rb[p] = rp[1];
                } else {
                  ra[p] = make_float4(0, 0, 0, 0); rb[p] = ra[p];
                }
              }
#pragma unroll
              for (int p = 0; p < PASSES; ++p) {
                v[p][0] += ra[p].x; v[p][1] += ra[p].y; v[p][2] += ra[p].z; v[p][3] += ra[p].w;
                v[p][4] += rb[p].x; v[p][5] += rb[p].y; v[p][6] += rb[p].z; v[p][7] += rb[p].w;
              }
            }
            if (g.residual_relu) {
#pragma unroll
              for (int p = 0; p < PASSES; ++p)
#pragma unroll
                for (int j = 0; j < 8; ++j) v[p][j] = fmaxf(v[p][j], 0.f);
            }
          }
#pragma unroll
          for (int p = 0; p < PASSES; ++p) {
            if (!colok || orow[p] < 0) continue;
            if (g.out_dtype == CROG_BF16) store8(reinterpret_cast<bf16*>(g.out) + (long long)orow[p] * g.out_ld + ncol, v[p]);
            else store8(reinterpret_cast<float*>(g.out) + (long long)orow[p] * g.out_ld + ncol, v[p]);
          }
        }
        __syncwarp();  // staging is reused by the next sub-block / tile
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, L::TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
int g_num_sms = 0;

template <int BN, int STAGES>
int launch(const CrogGemm* g, cudaStream_t stream) {
  using L = Cfg<BN, STAGES>;
  static_assert(L::TOTAL <= 227 * 1024, "shared memory budget");
  static bool attr_set = false;  // per-process; device attribute is re-set cheaply if another device is used
  static int attr_dev = -1;
  int dev = 0;
  CROG_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_set || attr_dev != dev) {
    CROG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    CROG_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true; attr_dev = dev;
  }
  CUtensorMap tmA, tmB;
  const long long Ktot = (long long)g->taps * g->cin;
  int rc = crog_encode_2d_bf16(&tmA, g->a, (uint64_t)g->cin, (uint64_t)g->a_rows, (uint64_t)g->a_ld, BM);
  if (rc) return rc;
  uint64_t wrows = (uint64_t)g->N;
  if (g->w_sample_stride > 0) wrows = (uint64_t)(g->w_sample_stride / Ktot) * (uint64_t)(g->M / g->sample_rows);
  rc = crog_encode_2d_bf16(&tmB, g->w, (uint64_t)Ktot, wrows, (uint64_t)Ktot, BN);
  if (rc) return rc;
  const int n_tiles = (g->N + BN - 1) / BN;
  const int total = num_m_tiles(*g, BM) * n_tiles;
  if (total == 0) return CROG_OK;
  const int grid = total < g_num_sms ? total : g_num_sms;
  gemm_tc_kernel<BN, STAGES><<<grid, NUM_THREADS, L::TOTAL, stream>>>(tmA, tmB, *g, n_tiles, total);
  CROG_LAUNCH_OK("gemm_tc");
  return CROG_OK;
}

}  // namespace

int crog_gemm_tc(const CrogGemm* g, cudaStream_t stream) {
  CROG_REQUIRE(g->dtype == CROG_BF16, CROG_E_BADSHAPE, "gemm_tc: bf16 operands only");
  CROG_REQUIRE(g->cin % BK == 0, CROG_E_BADSHAPE, "gemm_tc: cin %d not a multiple of %d", g->cin, BK);
  CROG_REQUIRE(g->N % 8 == 0, CROG_E_BADSHAPE, "gemm_tc: N %d not a multiple of 8", g->N);
  CROG_REQUIRE(aligned16(g->a) && aligned16(g->w) && aligned16(g->out) && g->a_ld % 8 == 0, CROG_E_BADALIGN,
               "gemm_tc: operands must be 16B aligned");
  CROG_REQUIRE(!g->residual || aligned16(g->residual), CROG_E_BADALIGN, "gemm_tc: residual must be 16B aligned");
  if (g->w_sample_stride > 0)
    CROG_REQUIRE(g->w_sample_stride % ((long long)g->taps * g->cin) == 0 && g->sample_rows > 0, CROG_E_BADSHAPE,
                 "gemm_tc: per-sample weights need whole rows");
  if (g->N <= 16) return launch<16, 8>(g, stream);
  if (g->N <= 64) return launch<64, 6>(g, stream);
  // 128 x 256 tiles cut the L2 -> smem operand traffic per FLOP by 25 % ((BM+BN)/(BM*BN)); worth it when the
  // contraction is long enough to be tensor/L2 bound rather than epilogue bound
  if (g->N % 256 == 0 && (long long)g->taps * g->cin >= 1024) return launch<256, 3>(g, stream);
  return launch<128, 5>(g, stream);
}

int crog_encode_2d_bf16(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_rows) {
  static CrogEncodeTiledFn enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<CrogEncodeTiledFn>(p);
  }
  CROG_REQUIRE(enc != nullptr, CROG_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {64u, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CROG_REQUIRE(r == CUDA_SUCCESS, CROG_E_CUDA, "cuTensorMapEncodeTiled failed (%d): cols=%llu rows=%llu ld=%llu", (int)r,
               (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld_elems);
  return CROG_OK;
}
