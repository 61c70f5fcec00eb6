// tcgen05 / TMEM / TMA implicit-GEMM for bf16 activations and weights (sm_100a).
//
// One CTA computes a 128 x BN output tile:
//   warp 0      TMA producer   (cp.async.bulk.tensor 2D, SWIZZLE_128B, mbarrier complete_tx)
//   warp 1      TMEM allocator + MMA issuer (tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16)
//   warps 2..5  epilogue       (tcgen05.ld 32x32b -> registers -> scale/bias/act/gate/residual -> HBM)
// The K loop runs over (tap, 64-channel chunk); for a 3x3 convolution on the zero-haloed
// ("padded") NHWC layout tap t is the same activation matrix shifted by a constant number of
// rows, so the A tile of every k-step is one plain 2D TMA box at row (row0 + shift_t); rows
// outside the tensor are zero-filled by TMA.  STAGES smem slots form the TMA<->MMA ring; two
// CTAs are co-resident per SM so one tile's epilogue overlaps the other's main loop.
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 elements = 128 bytes = one SWIZZLE_128B atom row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;


template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + ((B_BYTES + 1023) / 1024) * 1024;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int SCALE_OFF = BAR_OFF + 256;
  static constexpr int TOTAL = SCALE_OFF + 2 * BN * 4 + 1024;  // + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const CrogGemm g, int n_tiles) {
  using L = SmemLayout<BN, STAGES>;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);  // full[S], empty[S], tmem_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::BAR_OFF + (2 * STAGES + 1) * 8);
  float* s_scale = reinterpret_cast<float*>(smem + L::SCALE_OFF);
  float* s_bias = s_scale + BN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_t = blockIdx.x % n_tiles, m_t = blockIdx.x / n_tiles;
  const TileRows tr = tile_rows(g, m_t, BM);
  const int n0 = n_t * BN;
  const int kchunks = g.cin / BK;
  const int num_kb = g.taps * kchunks;

  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull = smem_u32(bars + 2 * STAGES);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tc_alloc(smem_u32(tmem_slot), TMEM_COLS);
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < BN; i += 128) {
      const int n = n0 + i;
      s_scale[i] = (g.scale && n < g.N) ? g.scale[n] : 1.f;
      s_bias[i] = (g.bias && n < g.N) ? g.bias[n] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int wrow0 = n0 + (g.w_sample_stride > 0 ? tr.wsample * (int)(g.w_sample_stride / ((long long)g.taps * g.cin)) : 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES, it = kb / STAGES;
        mbar_wait(empty0 + 8 * s, (it & 1) ^ 1);
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES), sb = sa + L::A_BYTES;
        const int tap = kb / kchunks, c0 = (kb % kchunks) * BK;
        mbar_expect_tx(full0 + 8 * s, L::A_BYTES + L::B_BYTES);
        tma_load_2d(sa, &tmA, c0, (int)(tr.row0 + tap_shift(g.taps, tap, g.W)), full0 + 8 * s);
        tma_load_2d(sb, &tmB, kb * BK, wrow0, full0 + 8 * s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES, it = kb / STAGES;
        mbar_wait(full0 + 8 * s, it & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES), sb = sa + L::A_BYTES;
        const uint64_t da = make_sdesc(sa), db = make_sdesc(sb);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          tc_mma_bf16(tmem_base, da + (uint64_t)(k * UMMA_K * 2 / 16), db + (uint64_t)(k * UMMA_K * 2 / 16), idesc,
                      (kb | k) != 0);
        tc_commit(empty0 + 8 * s);  // frees the smem slot once these MMAs retire
      }
      tc_commit(tfull);  // accumulator complete
    }
  } else {
    // epilogue: warp w owns TMEM lanes [32*(w%4), +32) == tile rows
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const RowMap m = map_row(g, tr.row0 + row, tr.row_end);
    mbar_wait(tfull, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    if constexpr (BN >= 32) {
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tc_ld32(taddr + c, r);
        tc_wait_ld();
        if (m.valid && n0 + c < g.N) {
          float acc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
          epilogue_row<32>(g, m, n0 + c, acc, s_scale + c, s_bias + c);
        }
      }
    } else {
      uint32_t r[16];
      tc_ld16(taddr, r);
      tc_wait_ld();
      if (m.valid) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(r[j]);
        epilogue_row<16>(g, m, n0, acc, s_scale, s_bias);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
template <int BN, int STAGES>
int launch(const CrogGemm* g, cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES>;
  static bool attr_set = false;  // per-process; device attribute is re-set cheaply if another device is used
  static int attr_dev = -1;
  int dev = 0;
  CROG_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_set || attr_dev != dev) {
    CROG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
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
  const int grid = num_m_tiles(*g, BM) * n_tiles;
  if (grid == 0) return CROG_OK;
  gemm_tc_kernel<BN, STAGES><<<grid, NUM_THREADS, L::TOTAL, stream>>>(tmA, tmB, *g, n_tiles);
  CROG_LAUNCH_OK("gemm_tc");
  return CROG_OK;
}

}  // namespace

int crog_gemm_tc(const CrogGemm* g, cudaStream_t stream) {
  CROG_REQUIRE(g->dtype == CROG_BF16, CROG_E_BADSHAPE, "gemm_tc: bf16 operands only");
  CROG_REQUIRE(g->cin % BK == 0, CROG_E_BADSHAPE, "gemm_tc: cin %d not a multiple of %d", g->cin, BK);
  CROG_REQUIRE(aligned16(g->a) && aligned16(g->w) && g->a_ld % 8 == 0, CROG_E_BADALIGN, "gemm_tc: operands must be 16B aligned");
  if (g->w_sample_stride > 0)
    CROG_REQUIRE(g->w_sample_stride % ((long long)g->taps * g->cin) == 0 && g->sample_rows > 0, CROG_E_BADSHAPE,
                 "gemm_tc: per-sample weights need whole rows");
  if (g->N <= 16) return launch<16, 4>(g, stream);
  if (g->N <= 64) return launch<64, 4>(g, stream);
  return launch<128, 3>(g, stream);
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
