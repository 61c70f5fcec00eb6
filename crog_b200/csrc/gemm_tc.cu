// tcgen05 / TMEM / TMA implicit-GEMM for bf16 activations and weights (sm_100a), persistent and
// warp-specialised.
//
// One CTA per SM loops over 128 x BN output tiles (static round-robin, N fastest so concurrent CTAs share A):
//   warp 0      TMA producer   (cp.async.bulk.tensor 2D, SWIZZLE_128B, mbarrier complete_tx) over a STAGES-deep ring
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16) into one of two TMEM
//                              accumulator buffers, so tile i+1 is being multiplied while tile i drains
//   warps 2..   epilogue       EPI_WARPS / 4 independent groups of four warps (one, two or three); group g drains
//                              accumulator buffer g (tiles i = g, g + G, ...), so G tiles are in flight in the
//                              epilogue.  A warp owns 32 accumulator rows (its TMEM lane quarter) and walks the tile's
//                              columns in 128-byte chunks.  The tile's scale / bias slice is staged in shared memory
//                              once per tile by the group (named barrier) and applied as packed FFMA2; the residual
//                              add is FADD2, ReLU on bf16 outputs is applied to the packed pairs.
//     TMA epilogue   (output rows == enumerated rows: every token matrix, compact->compact and padded->padded
//                    convolutions)  tcgen05.ld -> registers (thread = row) -> +addmat, scale/bias (or the row-affine
//                    form of a folded LayerNorm), act, gate -> optional per-row (sum, sum^2) partials ->
//                    (+ residual chunk that a TMA load prefetched into the staging buffer NBUF-1 chunks ahead) ->
//                    packed into the SWIZZLE_128B staging tile -> one cp.async.bulk.tensor store per chunk.
//                    Halo rows of a padded output are written as zeros, which is what they must hold.
//     legacy epilogue (compact<->padded row remapping, per-sample weights, N <= 16)  fp32 staging, then
//                    row-contiguous 16-byte stores with the output row looked up per staged row.
// The K loop runs over (tap, 64-channel chunk); for a 3x3 convolution on the zero-haloed ("padded") NHWC layout
// tap t is the same activation matrix shifted by a constant number of rows, so the A tile of every k-step is one plain
// 2D TMA box at row (row0 + shift_t); rows outside the tensor are zero-filled by TMA.  A second activation operand
// (CrogGemm.a2) continues the K loop after the first one's chunks: two summed 1x1 convolutions as one contraction.
#include <stdlib.h>

#include "tc_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 elements = 128 bytes = one SWIZZLE_128B atom row
constexpr int UMMA_K = 16;
constexpr int STG_BUF = 4096;  // one staging tile: 32 rows x 128 bytes
enum { MODE_LEGACY = 0, MODE_TMA_BF16 = 1, MODE_TMA_F32 = 2 };

// CONV3 (3x3 convolutions with one 64-channel chunk and N <= BN): the whole weight matrix (9 taps x BN x 64) stays
// resident in shared memory, and a stage holds BM + 2 activation rows of one ky band, so the three kx taps are three
// row-shifted views (descriptor start + kx * 128 B) of ONE TMA box: operand traffic from L2 drops from 9 x (A + B) to
// 3 x A per tile.
// EPI_WARPS = 8: two independent 4-warp epilogue groups (group g drains tiles g, g+2, ...), for short contractions
// where the epilogue is the pace; EPI_WARPS = 4: one group drains every tile (alternating accumulator buffers), which
// frees 32 KB of staging for one more operand stage when the contraction is long and the TMA feed is the pace.
// PAIR: two CTAs of a cluster (one TPC) work on one 256 x BN tile with tcgen05 cta_group::2: each CTA stages its own
// 128 rows of A and HALF of the weights (BN/2 rows), the leader issues M=256 MMAs that read both halves, and each CTA's
// TMEM receives its own 128 x BN accumulator.  Per CTA and k-block that is 32 KB written + 32 KB read from shared
// memory instead of 48 + 48: at full tensor rate 125 B/clk instead of 187 B/clk against a 128 B/clk shared-memory port,
// which is what caps the single-CTA 128 x 256 tile at ~67 % tensor utilisation.
// BAND (3x3 convolutions on the zero-haloed layout, any cin, streamed weights): the activation operand is staged as one
// (BM + 2)-row BAND per (ky, 64-channel chunk) in its own ring of STAGES slots, and the three kx taps read it through
// row-shifted descriptor views (as CONV3 does), while the weight tiles stream through a second ring of NBST slots, one
// per (ky, kx, chunk).  The big 3x3 convolutions are bound by the L2 -> SM fabric (the launch list shows 11.3 TB/s of TMA
// traffic into the SMs for proj.vis.3, the measured cap is ~6300 B/clk = 12 TB/s at 1.9 GHz): fetching every activation
// row once per ky instead of once per tap removes two thirds of the activation half of that traffic.
// DUAL (CONV3 only): two MMA-issuing warps, one per TMEM accumulator buffer, take the tiles alternately.  A 128 x 64 tile
// is 36 short MMAs (32 clocks each); the issuing thread spends ~570 instructions per tile on descriptors, election loops
// and barrier polls and was measured as the pace of the kernel (it never waits; tensor pipe 39 % active).
template <int BN, int STAGES, int NBUF, bool CONV3, int EPI_WARPS, bool PAIR = false, int NBST = 0, bool DUAL = false>
struct Cfg {
  static_assert(!DUAL || (CONV3 && EPI_WARPS == 8), "dual issue: CONV3 with two accumulator buffers");
  static constexpr int EPI_WARP0 = DUAL ? 3 : 2;               // first epilogue warp
  static constexpr bool BAND = NBST > 0;
  static_assert(!(BAND && CONV3), "BAND streams the weights, CONV3 keeps them resident");
  static constexpr int NUM_THREADS = (DUAL ? 96 : 64) + EPI_WARPS * 32;
  static constexpr int A_ROWS = (CONV3 || BAND) ? BM + 2 : BM;
  static constexpr int A_TX = A_ROWS * BK * 2;                 // bytes one A box delivers
  static constexpr int A_BYTES = ((A_TX + 1023) / 1024) * 1024;
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;            // weight rows this CTA stages
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int BB_BYTES = ((B_BYTES + 1023) / 1024) * 1024;
  static constexpr int STAGE_BYTES = (CONV3 || BAND) ? A_BYTES : A_BYTES + BB_BYTES;
  static constexpr int BRING_OFF = STAGES * STAGE_BYTES;       // BAND: the weight ring follows the band ring
  static constexpr int BRES_BYTES = CONV3 ? 9 * B_BYTES : (BAND ? NBST * BB_BYTES : 0);   // resident weights / weight ring
  static constexpr int SUB = BN > 32 ? 32 : BN;               // legacy: columns staged at a time
  static constexpr int PITCH = SUB * 4 + 16;                  // legacy staging row pitch (bytes): 16B-phase conflict free
  static constexpr int STG_WARP = NBUF * STG_BUF;             // per epilogue warp (>= 32 * PITCH = 4608)
  static constexpr int BRES_OFF = STAGES * STAGE_BYTES;
  static constexpr int STG_OFF = BRES_OFF + BRES_BYTES;
  static constexpr int BAR_OFF = STG_OFF + EPI_WARPS * STG_WARP;
  static constexpr int NACC = EPI_WARPS / 4 < 2 ? 2 : EPI_WARPS / 4;  // TMEM accumulator buffers: one per epilogue group, at least two
  static constexpr int NBARS = 2 * STAGES + 2 * NACC + 1 + EPI_WARPS * NBUF + 2 * NBST;  // full[S], empty[S], tfull[NACC], tempty[NACC], bres, res[EPI][NBUF], bfull[NBST], bempty[NBST]
  // per epilogue group: this tile's scale / bias slice, double buffered by tile parity ([2][scale | bias][BN] floats);
  // configurations whose operand ring leaves no room for it (long contractions, where the epilogue hides under the
  // mainloop anyway) keep reading scale / bias through the read-only cache
  static constexpr int SB_BYTES = (EPI_WARPS / 4) * 2 * 2 * BN * 4;
  static constexpr int SB_OFF = ((BAR_OFF + NBARS * 8 + 16 + 15) / 16) * 16;
  static constexpr bool SB = SB_OFF + SB_BYTES + 1024 <= 227 * 1024;
  static constexpr int TOTAL = SB ? SB_OFF + SB_BYTES + 1024 : BAR_OFF + NBARS * 8 + 16 + 1024;   // + tmem slot + alignment slack
  static constexpr int TMEM_COLS = NACC * BN <= 32 ? 32 : (NACC * BN <= 64 ? 64 : (NACC * BN <= 128 ? 128 : (NACC * BN <= 256 ? 256 : 512)));
  static_assert(NACC * BN <= 512, "accumulator buffers exceed the 512 TMEM columns");
  static_assert(NBUF >= 2 && STG_WARP >= 32 * PITCH, "staging too small");
};

template <int BN, int STAGES, int NBUF, int MODE, bool CONV3, int EPI_WARPS, bool PAIR, int NBST, bool DUAL>
__global__ void __launch_bounds__((DUAL ? 96 : 64) + EPI_WARPS * 32, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmA2,
                                                                   const __grid_constant__ CUtensorMap tmB,
                                                                   const __grid_constant__ CUtensorMap tmOut,
                                                                   const __grid_constant__ CUtensorMap tmRes,
                                                                   const CrogGemm g, int n_tiles, int total_tiles) {
  using L = Cfg<BN, STAGES, NBUF, CONV3, EPI_WARPS, PAIR, NBST, DUAL>;
  constexpr bool BAND = NBST > 0;
  constexpr int GROUPS = EPI_WARPS / 4;
  constexpr int MT = PAIR ? 2 * BM : BM;  // rows of one tile (over the CTA pair)
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B tiles; plain pointer arithmetic keeps the shared address space so the
  // epilogue's staging accesses compile to LDS / STS instead of generic loads
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::BAR_OFF + L::NBARS * 8);

  pdl_launch();  // the successor may start its prologue as soon as this grid is resident
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Written as expressions, not variables: with named copies of blockIdx.x / gridDim.x nvcc schedules the single-CTA
  // CONV3 kernels 11 % slower (measured A/B on one box: stem.conv2 260 -> 292 us) although the SASS mix is identical.
#define rank (PAIR ? cluster_ctarank() : 0u)                          /* 0 = leader (issues the MMAs) */
#define tile0f (PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x)      /* both CTAs of a pair walk the same tiles */
#define tstepf (PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x)
  // g.reverse: the same walk from the last tile down (tile indices leave the range below 0: unsigned compares)
#define tile0 (g.reverse ? total_tiles - 1 - tile0f : tile0f)
#define tstep (g.reverse ? -tstepf : tstepf)
#define TILE_IN(t) ((unsigned)(t) < (unsigned)total_tiles)
  // (Measured negative, round 2: evict_first L2 hints on the activation / residual loads.  The hinted lines leave the L2
  // before their own re-reads - the three ky bands of a 3x3 tile, the n-tiles of one row block - and the forward's DRAM
  // reads rise from 12.3 to 14.3 GB, +0.25 ms.  evict_last on the output stores: 12.3 -> 12.0 GB, time unchanged.)
  const int kchunks = g.cin / BK;
  const int num_kb = CONV3 ? 3 : g.taps * kchunks + g.cin2 / BK;  // CONV3: one k-block per ky band; a2: its chunks follow a's
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull0 = smem_u32(bars + 2 * STAGES),
                 tempty0 = smem_u32(bars + 2 * STAGES + L::NACC), bres = smem_u32(bars + 2 * STAGES + 2 * L::NACC),
                 res0 = smem_u32(bars + 2 * STAGES + 2 * L::NACC + 1),
                 bfull0 = smem_u32(bars + 2 * STAGES + 2 * L::NACC + 1 + EPI_WARPS * NBUF), bempty0 = bfull0 + 8 * NBST;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    if (g.a2) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    if (MODE != MODE_LEGACY) {
      tma_prefetch_desc(&tmOut);
      if (g.residual) tma_prefetch_desc(&tmRes);
    }
    // pair: the leader's full barrier collects one arrive(+expect_tx) per CTA, its tempty the epilogue threads of both
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, PAIR ? 2 : 1); mbar_init(empty0 + 8 * s, 1); }
    for (int s = 0; s < L::NACC; ++s) { mbar_init(tfull0 + 8 * s, 1); mbar_init(tempty0 + 8 * s, (PAIR ? 2 : 1) * 4 * 32); }
    for (int s = 0; s < EPI_WARPS * NBUF; ++s) mbar_init(res0 + 8 * s, 1);
    for (int s = 0; s < NBST; ++s) { mbar_init(bfull0 + 8 * s, PAIR ? 2 : 1); mbar_init(bempty0 + 8 * s, 1); }
    mbar_init(bres, PAIR ? 2 : 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tc_alloc_pair(smem_u32(tmem_slot), L::TMEM_COLS);
    else tc_alloc(smem_u32(tmem_slot), L::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the predecessor's tail; its outputs are visible from here on

  if (warp == 0) {
    if (lane == 0) {
      int kbg = 0;  // k-blocks issued so far (ring position)
      if constexpr (CONV3) {
        if constexpr (PAIR) {  // each CTA keeps its half of the weight rows; both report to the leader's barrier
          const uint32_t lbres = mapa_shared(bres, 0);
          mbar_expect_tx_cluster(lbres, L::BRES_BYTES);
          for (int t = 0; t < 9; ++t) tma_load_2d_pair(smem_u32(smem + L::BRES_OFF + t * L::B_BYTES), &tmB, t * BK, (int)rank * L::B_ROWS, lbres);
        } else {
          mbar_expect_tx(bres, L::BRES_BYTES);
          for (int t = 0; t < 9; ++t) tma_load_2d(smem_u32(smem + L::BRES_OFF + t * L::B_BYTES), &tmB, t * BK, 0, bres);
        }
        for (int tile = tile0; TILE_IN(tile); tile += tstep) {
          const long long row0 = (long long)tile * MT + (PAIR ? (long long)rank * BM : 0);  // n_tiles == 1
          for (int ky = 0; ky < 3; ++ky, ++kbg) {
            const int s = kbg % STAGES, it = kbg / STAGES;
            mbar_wait(empty0 + 8 * s, (it & 1) ^ 1);
            if constexpr (PAIR) {
              const uint32_t lfull = mapa_shared(full0 + 8 * s, 0);
              mbar_expect_tx_cluster(lfull, L::A_TX);
              tma_load_2d_pair(smem_u32(smem + s * L::STAGE_BYTES), &tmA, 0, (int)(row0 + (ky - 1) * (g.W + 2) - 1), lfull);
            } else {
              mbar_expect_tx(full0 + 8 * s, L::A_TX);
              tma_load_2d(smem_u32(smem + s * L::STAGE_BYTES), &tmA, 0, (int)(row0 + (ky - 1) * (g.W + 2) - 1), full0 + 8 * s);
            }
          }
        }
      } else if constexpr (BAND) {
        int ia = 0, ib = 0;  // band / weight-tile ring positions
        for (int tile = tile0; TILE_IN(tile); tile += tstep) {
          const int n_t = tile % n_tiles, m_t = tile / n_tiles;
          const long long row0 = (long long)m_t * MT + (PAIR ? (long long)rank * BM : 0);
          const int wrow0 = n_t * BN + (PAIR ? (int)rank * L::B_ROWS : 0);
          for (int ky = 0; ky < 3; ++ky)
            for (int c = 0; c < kchunks; ++c) {
              {
                const int sA = ia % STAGES, it = ia / STAGES;
                ++ia;
                mbar_wait(empty0 + 8 * sA, (it & 1) ^ 1);
                const uint32_t sa = smem_u32(smem + sA * L::STAGE_BYTES);
                if constexpr (PAIR) {
                  const uint32_t lfull = mapa_shared(full0 + 8 * sA, 0);
                  mbar_expect_tx_cluster(lfull, L::A_TX);
                  tma_load_2d_pair(sa, &tmA, c * BK, (int)(row0 + (ky - 1) * (g.W + 2) - 1), lfull);
                } else {
                  mbar_expect_tx(full0 + 8 * sA, L::A_TX);
                  tma_load_2d(sa, &tmA, c * BK, (int)(row0 + (ky - 1) * (g.W + 2) - 1), full0 + 8 * sA);
                }
              }
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const int sB = ib % NBST, it = ib / NBST;
                ++ib;
                mbar_wait(bempty0 + 8 * sB, (it & 1) ^ 1);
                const uint32_t sb = smem_u32(smem + L::BRING_OFF + sB * L::BB_BYTES);
                const int wcol = (ky * 3 + kx) * g.cin + c * BK;
                if constexpr (PAIR) {
                  const uint32_t lfull = mapa_shared(bfull0 + 8 * sB, 0);
                  mbar_expect_tx_cluster(lfull, L::B_BYTES);
                  tma_load_2d_pair(sb, &tmB, wcol, wrow0, lfull);
                } else {
                  mbar_expect_tx(bfull0 + 8 * sB, L::B_BYTES);
                  tma_load_2d(sb, &tmB, wcol, wrow0, bfull0 + 8 * sB);
                }
              }
            }
        }
      } else
      for (int tile = tile0; TILE_IN(tile); tile += tstep) {
        const int n_t = tile % n_tiles, m_t = tile / n_tiles;
        TileRows tr = tile_rows(g, m_t, MT);
        if constexpr (PAIR) tr.row0 += (long long)rank * BM;  // shared weights only: tiles are plain row ranges
        const int wrow0 = n_t * BN + (g.w_sample_stride > 0 ? tr.wsample * (int)(g.w_sample_stride / ((long long)g.taps * g.cin)) : 0);
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const int s = kbg % STAGES, it = kbg / STAGES;
          mbar_wait(empty0 + 8 * s, (it & 1) ^ 1);
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES), sb = sa + L::A_BYTES;
          // k-block order.  3x3 convolutions: (ky, 64-channel chunk, kx) - the same order in every tile configuration
          // (single CTA, CTA pair, BAND, CONV3), so that the fp32 accumulation and hence every output bit is independent
          // of the configuration the autotuner picks; everything else: (tap, chunk), the second operand's chunks last.
          int tap, c0;
          if (g.taps == 9) { const int kyc = kb / 3; tap = (kyc / kchunks) * 3 + kb % 3; c0 = (kyc % kchunks) * BK; }
          else { tap = kb / kchunks; c0 = (kb % kchunks) * BK; }
          const int wcol = g.taps == 9 ? tap * g.cin + c0 : kb * BK;
          if constexpr (PAIR) {
            // both CTAs report to the LEADER's full barrier; this CTA stages its own rows of A and its half of the weights
            const uint32_t lfull = mapa_shared(full0 + 8 * s, 0);
            mbar_expect_tx_cluster(lfull, L::A_BYTES + L::B_BYTES);
            if (kb >= g.taps * kchunks) tma_load_2d_pair(sa, &tmA2, (kb - kchunks) * BK, (int)tr.row0, lfull);  // second operand (taps == 1)
            else tma_load_2d_pair(sa, &tmA, c0, (int)(tr.row0 + tap_shift(g.taps, tap, g.W)), lfull);
            tma_load_2d_pair(sb, &tmB, wcol, wrow0 + (int)rank * L::B_ROWS, lfull);
          } else {
            mbar_expect_tx(full0 + 8 * s, L::A_BYTES + L::B_BYTES);
            if (kb >= g.taps * kchunks) tma_load_2d(sa, &tmA2, (kb - kchunks) * BK, (int)tr.row0, full0 + 8 * s);  // second operand (taps == 1)
            else tma_load_2d(sa, &tmA, c0, (int)(tr.row0 + tap_shift(g.taps, tap, g.W)), full0 + 8 * s);
            tma_load_2d(sb, &tmB, wcol, wrow0, full0 + 8 * s);
          }
        }
      }
    }
  } else if (warp == 1 || (DUAL && warp == 2)) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc(MT, BN);
      // dual issue: this warp takes the tiles of its own parity (= its accumulator buffer); ring position 3 per tile
      constexpr int ISTEP = DUAL ? 2 : 1;
      int i = DUAL ? warp - 1 : 0;
      int kbg = DUAL ? 3 * i : 0;
      if constexpr (CONV3) mbar_wait(bres, 0);
      for (int tile = tile0 + i * tstep; TILE_IN(tile); tile += ISTEP * tstep, i += ISTEP, kbg += DUAL ? 3 : 0) {
        const int as = i % L::NACC;
        mbar_wait(tempty0 + 8 * as, ((i / L::NACC) & 1) ^ 1);  // epilogue group `as` has drained this accumulator buffer
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        if constexpr (CONV3) {
          if (g.tap_mask == 0) {  // all nine taps: the straight-line loop (kept apart from the masked form below: the
                                  // scheduling of this hot loop is sensitive to any extra control flow, -18 % measured)
            for (int ky = 0; ky < 3; ++ky, ++kbg) {
              const int s = kbg % STAGES, it = kbg / STAGES;
              mbar_wait(full0 + 8 * s, it & 1);
              tc_fence_after();
              const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                // rows [kx, kx + 128) of the band: the SWIZZLE_128B XOR is a function of the absolute shared address
                // (descriptor base_offset = 0), so a start address moved by whole 128-byte rows still reads the
                // pattern TMA wrote (verified on B200: tests/test_gpu_kernels.py::test_conv3x3_padded)
                const uint64_t da = make_sdesc(sa + kx * 128);
                const uint64_t db = make_sdesc(smem_u32(smem + L::BRES_OFF + (ky * 3 + kx) * L::B_BYTES));
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                  if constexpr (PAIR) tc_mma_bf16_pair(tacc, da + (uint64_t)(k * UMMA_K * 2 / 16), db + (uint64_t)(k * UMMA_K * 2 / 16), idesc, (ky | kx | k) != 0);
                  else tc_mma_bf16(tacc, da + (uint64_t)(k * UMMA_K * 2 / 16), db + (uint64_t)(k * UMMA_K * 2 / 16), idesc, (ky | kx | k) != 0);
                }
              }
              if constexpr (PAIR) tc_commit_pair(empty0 + 8 * s, 3);
              else tc_commit(empty0 + 8 * s);
            }
          } else {
            const uint32_t tmask = (uint32_t)g.tap_mask;
            uint32_t acc_on = 0;  // 0 until the tile's first MMA has been issued
            for (int ky = 0; ky < 3; ++ky, ++kbg) {
              const int s = kbg % STAGES, it = kbg / STAGES;
              mbar_wait(full0 + 8 * s, it & 1);
              tc_fence_after();
              const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                if (!((tmask >> (ky * 3 + kx)) & 1u)) continue;  // tap not contracted (its weights are zero)
                const uint64_t da = make_sdesc(sa + kx * 128);
                const uint64_t db = make_sdesc(smem_u32(smem + L::BRES_OFF + (ky * 3 + kx) * L::B_BYTES));
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                  if constexpr (PAIR) tc_mma_bf16_pair(tacc, da + (uint64_t)(k * UMMA_K * 2 / 16), db + (uint64_t)(k * UMMA_K * 2 / 16), idesc, acc_on | (uint32_t)k);
                  else tc_mma_bf16(tacc, da + (uint64_t)(k * UMMA_K * 2 / 16), db + (uint64_t)(k * UMMA_K * 2 / 16), idesc, acc_on | (uint32_t)k);
                }
                acc_on = 1;
              }
              if constexpr (PAIR) tc_commit_pair(empty0 + 8 * s, 3);
              else tc_commit(empty0 + 8 * s);
            }
          }
        } else if constexpr (BAND) {
          for (int kyc = 0; kyc < 3 * kchunks; ++kyc, ++kbg) {  // kbg counts bands here; the weight ring has its own cursor
            const int sA = kbg % STAGES, ita = kbg / STAGES;
            mbar_wait(full0 + 8 * sA, ita & 1);
            const uint32_t sa = smem_u32(smem + sA * L::STAGE_BYTES);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int ibm = kbg * 3 + kx, sB = ibm % NBST, itb = ibm / NBST;
              mbar_wait(bfull0 + 8 * sB, itb & 1);
              tc_fence_after();
              const uint64_t da = make_sdesc(sa + kx * 128);  // rows [kx, kx + 128) of the band (see CONV3)
              const uint64_t db = make_sdesc(smem_u32(smem + L::BRING_OFF + sB * L::BB_BYTES));
              if constexpr (PAIR) {
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                  tc_mma_bf16_pair(tacc, da + (uint64_t)(k * UMMA_K * 2 / 16), db + (uint64_t)(k * UMMA_K * 2 / 16), idesc, (kyc | kx | k) != 0);
                tc_commit_pair(bempty0 + 8 * sB, 3);
              } else {
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                  tc_mma_bf16(tacc, da + (uint64_t)(k * UMMA_K * 2 / 16), db + (uint64_t)(k * UMMA_K * 2 / 16), idesc, (kyc | kx | k) != 0);
                tc_commit(bempty0 + 8 * sB);
              }
            }
            if constexpr (PAIR) tc_commit_pair(empty0 + 8 * sA, 3);
            else tc_commit(empty0 + 8 * sA);
          }
        } else
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const int s = kbg % STAGES, it = kbg / STAGES;
          mbar_wait(full0 + 8 * s, it & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES), sb = sa + L::A_BYTES;
          const uint64_t da = make_sdesc(sa), db = make_sdesc(sb);
          if constexpr (PAIR) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              tc_mma_bf16_pair(tacc, da + (uint64_t)(k * UMMA_K * 2 / 16), db + (uint64_t)(k * UMMA_K * 2 / 16), idesc, (kb | k) != 0);
            tc_commit_pair(empty0 + 8 * s, 3);  // frees this slot in BOTH CTAs once these MMAs retire
          } else {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              tc_mma_bf16(tacc, da + (uint64_t)(k * UMMA_K * 2 / 16), db + (uint64_t)(k * UMMA_K * 2 / 16), idesc, (kb | k) != 0);
            tc_commit(empty0 + 8 * s);  // frees the smem slot once these MMAs retire
          }
        }
        if constexpr (PAIR) tc_commit_pair(tfull0 + 8 * as, 3);  // accumulator complete, in both CTAs
        else tc_commit(tfull0 + 8 * as);  // accumulator complete
      }
    }
  } else {
    const int ew = warp - L::EPI_WARP0;      // 0..EPI_WARPS-1
    const int grp = ew >> 2;                 // this warp's group drains tiles grp, grp + GROUPS, ... (buffer = tile parity)
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    uint8_t* stg = smem + L::STG_OFF + ew * L::STG_WARP;
    // scale / bias of the current tile in shared memory (see Cfg::SB): the group's 128 threads fetch the slice before
    // they wait for the accumulator, store it, and meet on the group's named barrier (id 1 + grp)
    float* sb_base = reinterpret_cast<float*>(smem + L::SB_OFF) + grp * (2 * 2 * BN);
    const int gt = (ew & 3) * 32 + lane;     // thread index inside the group
    auto stage_scale_bias = [&](int ord, int n0, uint32_t tfull_bar, uint32_t tfull_parity) -> const float* {
      float* sbt = sb_base + (ord & 1) * (2 * BN);
      if constexpr (L::SB) {
        // one column tile (N <= BN): every tile of this CTA uses the same scale / bias slice, and after the group's first
        // two tiles both parity buffers hold it - no global loads, no shared-memory writes, no group barrier (the load's
        // L2 round trip in front of every short tile was a quarter of the epilogue warps' stall samples)
        if (n_tiles == 1 && ord >= 2) {
          mbar_wait(tfull_bar, tfull_parity);
          return sbt;
        }
        float v[(2 * BN + 127) / 128];
#pragma unroll
        for (int u = 0; u < (2 * BN + 127) / 128; ++u) {
          const int idx = gt + u * 128, isb = idx >= BN, n = n0 + (isb ? idx - BN : idx);
          v[u] = isb ? 0.f : 1.f;
          const float* src = isb ? g.bias : g.scale;
          if (idx < 2 * BN && n < g.N && src) v[u] = __ldg(src + n);
        }
        mbar_wait(tfull_bar, tfull_parity);
#pragma unroll
        for (int u = 0; u < (2 * BN + 127) / 128; ++u)
          if (gt + u * 128 < 2 * BN) sbt[gt + u * 128] = v[u];
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else if (grp == 1) asm volatile("bar.sync 2, 128;" ::: "memory");
        else asm volatile("bar.sync 3, 128;" ::: "memory");
      } else {
        mbar_wait(tfull_bar, tfull_parity);
      }
      return sbt;
    };

    if constexpr (MODE != MODE_LEGACY) {
      // ------------------------------------------------------------ TMA epilogue (identity row mapping)
      constexpr int CW = MODE == MODE_TMA_BF16 ? 64 : 32;  // columns per 128-byte chunk
      constexpr int NCH = BN / CW;
      static_assert(BN % CW == 0, "tile must hold whole chunks");
      const bool has_res = g.residual != nullptr;
      // bf16 outputs: ReLU commutes with the rounding, so it is applied to the packed pairs (HMNMX2) after the pack
      const bool relu_late = L::SB && MODE == MODE_TMA_BF16 && g.act == CROG_ACT_RELU && !has_res && g.gate == nullptr &&
                             g.row_stats_out == nullptr;  // row statistics are taken of the clamped fp32 values
      const bool relu_packed = MODE == MODE_TMA_BF16 && (relu_late || (has_res && g.residual_relu));
      const uint32_t stg_u32 = smem_u32(stg), rbar0 = res0 + 8 * (ew * NBUF);
      const int sw = lane & 7;
      uint32_t nld = 0, ncs = 0;   // residual chunks requested / chunks consumed (warp-uniform)
      // prefetch cursor: tile ordinal, chunk in tile, and the tile's coordinates (recomputed once per tile: the two
      // integer divisions per chunk cost ~500 clocks of every 3800-clock chunk in the in-kernel trace)
      int pf_i = grp, pf_c = 0, pf_n0 = 0, pf_row = 0;
      bool pf_ok = false;
      auto pf_set_tile = [&]() {
        const int tile = tile0 + pf_i * tstep;
        pf_ok = TILE_IN(tile);
        if (pf_ok) { pf_n0 = (tile % n_tiles) * BN; pf_row = (tile / n_tiles) * MT + (int)rank * BM + q * 32; }
      };
      pf_set_tile();
      auto issue_prefetch = [&]() {
        if (!pf_ok) return;
        if (lane == 0) {
          const uint32_t b = nld % NBUF;
          mbar_expect_tx(rbar0 + 8 * b, STG_BUF);
          tma_load_2d(stg_u32 + b * STG_BUF, &tmRes, pf_n0 + pf_c * CW, pf_row, rbar0 + 8 * b);
        }
        ++nld;
        if (++pf_c == NCH || pf_n0 + pf_c * CW >= g.N) { pf_c = 0; pf_i += GROUPS; pf_set_tile(); }
      };
      if (has_res)
        for (int j = 0; j < NBUF - 1; ++j) issue_prefetch();
      for (int i = grp;; i += GROUPS) {
        const int tile = tile0 + i * tstep;
        if (!TILE_IN(tile)) break;
        const int buf = i % L::NACC;
        const uint32_t tfull = tfull0 + 8 * buf, tempty = tempty0 + 8 * buf;
        const int n0 = (tile % n_tiles) * BN, row0 = (tile / n_tiles) * MT + (int)rank * BM;
        RowMap m = map_row(g, (long long)row0 + q * 32 + lane, g.M);
        load_row_stats(g, m);
        const int nvc = min(NCH, (g.N - n0 + CW - 1) / CW);  // chunks with at least one real column
        const float* sbt = stage_scale_bias(i / GROUPS, n0, tfull, (i / L::NACC) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN;
        for (int c = 0; c < nvc; ++c) {
          float acc[CW];
          {  // both 32-column TMEM loads of a 64-column chunk in flight before the one wait (one TMEM round trip per chunk)
            uint32_t r[CW / 32][32];
#pragma unroll
            for (int h = 0; h < CW / 32; ++h) tc_ld32(taddr + c * CW + h * 32, r[h]);
            tc_wait_ld();
#pragma unroll
            for (int h = 0; h < CW / 32; ++h)
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[h * 32 + j] = __uint_as_float(r[h][j]);
          }
          if (c == nvc - 1) {  // last read of this accumulator buffer: hand it back to the (leader's) MMA warp
            tc_fence_before();
            if constexpr (PAIR) mbar_arrive_cluster(mapa_shared(tempty, 0));
            else mbar_arrive(tempty);
          }
          const int ncol = n0 + c * CW;
          if (m.valid) {
#pragma unroll
            for (int h = 0; h < CW / 32; ++h) {
              if (ncol + h * 32 < g.N) {
                float(&a32)[32] = *reinterpret_cast<float(*)[32]>(&acc[h * 32]);
                if constexpr (L::SB)
                  epilogue_math_smem<32>(g, m, ncol + h * 32, a32, sbt + c * CW + h * 32, sbt + BN + c * CW + h * 32, relu_late);
                else
                  epilogue_math<32>(g, m, ncol + h * 32, a32, g.scale ? g.scale + ncol + h * 32 : nullptr,
                                    g.bias ? g.bias + ncol + h * 32 : nullptr);
              }
            }
          }
          if constexpr (MODE == MODE_TMA_BF16) {
            if (g.row_stats_out && m.valid) {  // (sum, sum of squares) of this row's 64 outputs, for the LayerNorm folded into the next GEMM
              float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
              for (int j = 0; j < CW; j += 2) {
                s0 += acc[j]; s1 += acc[j + 1];
                q0 = fmaf(acc[j], acc[j], q0); q1 = fmaf(acc[j + 1], acc[j + 1], q1);
              }
              reinterpret_cast<float2*>(g.row_stats_out)[(long long)m.orow * g.row_stats_chunks + (ncol / CW)] = make_float2(s0 + s1, q0 + q1);
            }
          }
          const uint32_t b = ncs % NBUF;
          uint8_t* srow = stg + b * STG_BUF + lane * 128;
          if (has_res) {
            mbar_wait(rbar0 + 8 * b, (ncs / NBUF) & 1);
            if constexpr (MODE == MODE_TMA_BF16) {
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const uint4 rr = *reinterpret_cast<const uint4*>(srow + ((u ^ sw) << 4));
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = __bfloat1622float2(h2[j]);
                  fadd2(acc[u * 8 + 2 * j], acc[u * 8 + 2 * j + 1], f.x, f.y);
                }
              }
            } else {
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const float4 f = *reinterpret_cast<const float4*>(srow + ((u ^ sw) << 4));
                fadd2(acc[u * 4], acc[u * 4 + 1], f.x, f.y);
                fadd2(acc[u * 4 + 2], acc[u * 4 + 3], f.z, f.w);
              }
            }
            if (g.residual_relu && !relu_packed) {
#pragma unroll
              for (int j = 0; j < CW; ++j) acc[j] = fmaxf(acc[j], 0.f);
            }
          } else {
            // the store that last read this staging buffer (NBUF chunks ago) must have drained it
            if (lane == 0) bulk_wait_read<NBUF - 1>();
            __syncwarp();
          }
          if (!m.valid) {  // halo rows of a padded tensor hold zeros; rows >= M are clipped by the store
#pragma unroll
            for (int j = 0; j < CW; ++j) acc[j] = 0.f;
          }
          if constexpr (MODE == MODE_TMA_BF16) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              uint4 pk;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
              for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(acc[u * 8 + 2 * j], acc[u * 8 + 2 * j + 1]);
              if (relu_packed) {
                const __nv_bfloat162 z2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 4; ++j) h2[j] = __hmax2(h2[j], z2);
              }
              *reinterpret_cast<uint4*>(srow + ((u ^ sw) << 4)) = pk;
            }
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u)
              *reinterpret_cast<float4*>(srow + ((u ^ sw) << 4)) = make_float4(acc[u * 4], acc[u * 4 + 1], acc[u * 4 + 2], acc[u * 4 + 3]);
          }
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA engine (async proxy)
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmOut, stg_u32 + b * STG_BUF, ncol, row0 + q * 32);
            bulk_commit();
          }
          ++ncs;
          if (has_res) {
            // the next prefetch lands in the buffer of chunk ncs-2 (the previous one): its store must have drained it
            if (lane == 0) bulk_wait_read<1>();
            issue_prefetch();
          }
        }
      }
      if (lane == 0) bulk_wait_all();
    } else {
      // ------------------------------------------------------------ legacy epilogue (row remapping)
      constexpr int SUB = L::SUB;
      constexpr int NSUB = BN / SUB;
      constexpr int UPR = SUB / 8;             // 16-byte (8-column) units per staged row
      constexpr int RPP = 32 / UPR;            // rows written per warp pass
      constexpr int PASSES = 32 / RPP;
      const int u = lane % UPR, rsub = lane / UPR;
      for (int i = grp;; i += GROUPS) {
        const int tile = tile0 + i * tstep;
        if (!TILE_IN(tile)) break;
        const int buf = i % L::NACC;
        const uint32_t tfull = tfull0 + 8 * buf, tempty = tempty0 + 8 * buf;
        const int n_t = tile % n_tiles, m_t = tile / n_tiles;
        TileRows tr = tile_rows(g, m_t, MT);
        if constexpr (PAIR) tr.row0 += (long long)rank * BM;
        const int n0 = n_t * BN;
        RowMap m = map_row(g, tr.row0 + q * 32 + lane, tr.row_end);
        load_row_stats(g, m);
        const int my_orow = m.valid ? m.orow : -1;
        int orow[PASSES];
#pragma unroll
        for (int p = 0; p < PASSES; ++p) orow[p] = __shfl_sync(0xffffffffu, my_orow, p * RPP + rsub);
        const int nvs = min(NSUB, (g.N - n0 + SUB - 1) / SUB);
        const float* sbt = stage_scale_bias(i / GROUPS, n0, tfull, (i / L::NACC) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN;
        for (int sbi = 0; sbi < nvs; ++sbi) {
          const int cbase = sbi * SUB;
          // ---- phase 1: accumulator -> registers -> epilogue math -> fp32 staging (thread = row)
          float acc[SUB];
          if constexpr (SUB == 32) {
            uint32_t r[32];
            tc_ld32(taddr + cbase, r);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
          } else {
            uint32_t r[16];
            tc_ld16(taddr + cbase, r);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(r[j]);
          }
          if (sbi == nvs - 1) {
            tc_fence_before();
            if constexpr (PAIR) mbar_arrive_cluster(mapa_shared(tempty, 0));
            else mbar_arrive(tempty);
          }
          if (m.valid) {
            if constexpr (L::SB && SUB % 4 == 0) epilogue_math_smem<SUB>(g, m, n0 + cbase, acc, sbt + cbase, sbt + BN + cbase);
            else epilogue_math<SUB>(g, m, n0 + cbase, acc, g.scale ? g.scale + n0 + cbase : nullptr, g.bias ? g.bias + n0 + cbase : nullptr);
          }
          float4* dst = reinterpret_cast<float4*>(stg + lane * L::PITCH);
#pragma unroll
          for (int j = 0; j < SUB / 4; ++j) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
          __syncwarp();
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
              for (int p = 0; p < PASSES; ++p)
                rr[p] = (colok && orow[p] >= 0)
                            ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(g.residual) + (long long)orow[p] * g.res_ld + ncol))
                            : make_uint4(0, 0, 0, 0);
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
          __syncwarp();  // staging is reused by the next sub-block / tile
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // neither CTA may leave while the other can still signal its barriers / TMEM
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tc_dealloc_pair(tmem_base, L::TMEM_COLS);
    else tc_dealloc(tmem_base, L::TMEM_COLS);
  }
}

#undef rank
#undef tile0
#undef tstep
#undef tile0f
#undef tstepf
#undef TILE_IN

// ---------------------------------------------------------------- host side
int g_num_sms = 0;

template <int BN, int STAGES, int NBUF, int MODE, bool CONV3 = false, int EPI_WARPS = 8, bool PAIR = false, int NBST = 0, bool DUAL = false>
int launch(const CrogGemm* g, cudaStream_t stream) {
  using L = Cfg<BN, STAGES, NBUF, CONV3, EPI_WARPS, PAIR, NBST, DUAL>;
  static_assert(L::TOTAL <= 227 * 1024, "shared memory budget");
  static DeviceOnce once;  // function attributes are per device (one static per template instantiation)
  int dev = 0;
  if (once.need(&dev) || g_num_sms == 0) {
    CROG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, NBUF, MODE, CONV3, EPI_WARPS, PAIR, NBST, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    CROG_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    once.done(dev);
  }
  CUtensorMap tmA, tmA2, tmB, tmOut, tmRes;
  const long long Ktot = (long long)g->taps * g->cin + g->cin2;
  int rc = crog_encode_2d(&tmA, g->a, CROG_BF16, (uint64_t)g->cin, (uint64_t)g->a_rows, (uint64_t)g->a_ld, L::A_ROWS);
  if (rc) return rc;
  tmA2 = tmA;
  if (g->a2) {
    rc = crog_encode_2d(&tmA2, g->a2, CROG_BF16, (uint64_t)g->cin2, (uint64_t)g->a_rows, (uint64_t)g->a2_ld, L::A_ROWS);
    if (rc) return rc;
  }
  uint64_t wrows = (uint64_t)g->N;
  if (g->w_sample_stride > 0) wrows = (uint64_t)(g->w_sample_stride / Ktot) * (uint64_t)(g->M / g->sample_rows);
  rc = crog_encode_2d(&tmB, g->w, CROG_BF16, (uint64_t)Ktot, wrows, (uint64_t)Ktot, L::B_ROWS);
  if (rc) return rc;
  if (MODE != MODE_LEGACY) {
    rc = crog_encode_2d(&tmOut, g->out, g->out_dtype, (uint64_t)g->N, (uint64_t)g->M, (uint64_t)g->out_ld, 32);
    if (rc) return rc;
    tmRes = tmOut;
    if (g->residual) {
      rc = crog_encode_2d(&tmRes, g->residual, g->out_dtype, (uint64_t)g->N, (uint64_t)g->M, (uint64_t)g->res_ld, 32);
      if (rc) return rc;
    }
  } else {
    tmOut = tmA; tmRes = tmA;  // unused
  }
  const int n_tiles = (g->N + BN - 1) / BN;
  const int total = num_m_tiles(*g, PAIR ? 2 * BM : BM) * n_tiles;
  if (total == 0) return CROG_OK;
  if constexpr (PAIR) {
    // one cluster of two CTAs (an SM pair of one TPC) per tile stream
    int pairs = total < g_num_sms / 2 ? total : g_num_sms / 2;
    if (g->max_ctas > 0 && pairs > g->max_ctas / 2) pairs = g->max_ctas / 2 > 0 ? g->max_ctas / 2 : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(L::NUM_THREADS); cfg.dynamicSmemBytes = L::TOTAL; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CROG_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, NBUF, MODE, CONV3, EPI_WARPS, PAIR, NBST, DUAL>, tmA, tmA2, tmB, tmOut, tmRes, *g, n_tiles, total));
    return CROG_OK;
  }
  int grid = total < g_num_sms ? total : g_num_sms;
  if (g->max_ctas > 0 && grid > g->max_ctas) grid = g->max_ctas;
  crog_launch(gemm_tc_kernel<BN, STAGES, NBUF, MODE, CONV3, EPI_WARPS, PAIR, NBST, DUAL>, dim3(grid), dim3(L::NUM_THREADS), L::TOTAL, stream, tmA, tmA2, tmB, tmOut, tmRes, *g, n_tiles, total);
  CROG_LAUNCH_OK("gemm_tc");
  return CROG_OK;
}

template <int MODE>
int dispatch(const CrogGemm* g, cudaStream_t stream) {
  const long long Ktot = (long long)g->taps * g->cin + g->cin2;
  const bool conv3_ok = g->N <= 64 && g->taps == 9 && g->cin == BK && g->w_sample_stride == 0;
  const bool pair_ok = g->w_sample_stride == 0 && g->M > BM;  // shared weights, more than one CTA's worth of rows
  const bool band_ok = g->taps == 9 && g->in_padded && g->w_sample_stride == 0 && g->a2 == nullptr && g->H > 0;
  if (g->tile_cfg != CROG_TILE_AUTO) {
    // forced configuration (plan-time autotuner, tests): every one of them accumulates the k-blocks in the same order
    switch (g->tile_cfg) {
      case CROG_TILE_128x128: return launch<128, 4, 3, MODE>(g, stream);
      case CROG_TILE_128x256:
        CROG_REQUIRE(g->N > 128, CROG_E_BADSHAPE, "gemm_tc: 128x256 tiles need N > 128");
        return launch<256, 4, 2, MODE, false, 4>(g, stream);
      case CROG_TILE_128x256_E8:
        CROG_REQUIRE(g->N > 128, CROG_E_BADSHAPE, "gemm_tc: 128x256 tiles need N > 128");
        return launch<256, 3, 2, MODE>(g, stream);
      case CROG_TILE_PAIR_256x256:
        CROG_REQUIRE(pair_ok && g->N > 128, CROG_E_BADSHAPE, "gemm_tc: CTA-pair 256x256 tiles need shared weights, M > 128, N > 128");
        return launch<256, 6, 2, MODE, false, 4, true>(g, stream);
      case CROG_TILE_PAIR_256x256_E8:
        CROG_REQUIRE(pair_ok && g->N > 128, CROG_E_BADSHAPE, "gemm_tc: CTA-pair 256x256 tiles need shared weights, M > 128, N > 128");
        return launch<256, 4, 2, MODE, false, 8, true>(g, stream);
      case CROG_TILE_PAIR_256x128:
        CROG_REQUIRE(pair_ok && g->N > 64, CROG_E_BADSHAPE, "gemm_tc: CTA-pair 256x128 tiles need shared weights, M > 128, N > 64");
        return launch<128, 6, 2, MODE, false, 8, true>(g, stream);
      case CROG_TILE_128x64: return launch<64, 5, 3, MODE>(g, stream);
      case CROG_TILE_128x128_S3: return launch<128, 3, 3, MODE>(g, stream);
      case CROG_TILE_128x128_E12: return launch<128, 3, 2, MODE, false, 12>(g, stream);
      case CROG_TILE_CONV3:
        CROG_REQUIRE(conv3_ok, CROG_E_BADSHAPE, "gemm_tc: CONV3 tiles need a shared 3x3 kernel with cin == 64 and N <= 64");
        return launch<64, 5, 2, MODE, true>(g, stream);
      case CROG_TILE_CONV3_E12:
        CROG_REQUIRE(conv3_ok, CROG_E_BADSHAPE, "gemm_tc: CONV3 tiles need a shared 3x3 kernel with cin == 64 and N <= 64");
        return launch<64, 3, 2, MODE, true, 12>(g, stream);
      case CROG_TILE_CONV3_DUAL:
        CROG_REQUIRE(conv3_ok, CROG_E_BADSHAPE, "gemm_tc: CONV3 tiles need a shared 3x3 kernel with cin == 64 and N <= 64");
        return launch<64, 5, 2, MODE, true, 8, false, 0, true>(g, stream);
      case CROG_TILE_CONV3_PAIR:
        CROG_REQUIRE(conv3_ok && pair_ok, CROG_E_BADSHAPE, "gemm_tc: CONV3 pair tiles need a shared 3x3 kernel with cin == 64, N <= 64, M > 128");
        return launch<64, 5, 2, MODE, true, 8, true>(g, stream);
      case CROG_TILE_BAND_PAIR_256x256:
        CROG_REQUIRE(band_ok && pair_ok && g->N > 128, CROG_E_BADSHAPE, "gemm_tc: BAND pair 256x256 needs a shared 3x3 kernel on the padded layout, M > 128, N > 128");
        return launch<256, 3, 2, MODE, false, 4, true, 6>(g, stream);
      case CROG_TILE_BAND_PAIR_256x256_E8:
        CROG_REQUIRE(band_ok && pair_ok && g->N > 128, CROG_E_BADSHAPE, "gemm_tc: BAND pair 256x256 needs a shared 3x3 kernel on the padded layout, M > 128, N > 128");
        return launch<256, 3, 2, MODE, false, 8, true, 5>(g, stream);
      case CROG_TILE_BAND_PAIR_256x128:
        CROG_REQUIRE(band_ok && pair_ok && g->N > 64, CROG_E_BADSHAPE, "gemm_tc: BAND pair 256x128 needs a shared 3x3 kernel on the padded layout, M > 128, N > 64");
        return launch<128, 4, 2, MODE, false, 8, true, 8>(g, stream);
      case CROG_TILE_BAND_128x128:
        CROG_REQUIRE(band_ok && g->N > 64, CROG_E_BADSHAPE, "gemm_tc: BAND 128x128 needs a shared 3x3 kernel on the padded layout, N > 64");
        return launch<128, 3, 2, MODE, false, 8, false, 6>(g, stream);
      case CROG_TILE_BAND_128x256:
        CROG_REQUIRE(band_ok && g->N > 128, CROG_E_BADSHAPE, "gemm_tc: BAND 128x256 needs a shared 3x3 kernel on the padded layout, N > 128");
        return launch<256, 3, 2, MODE, false, 4, false, 4>(g, stream);
      default: CROG_REQUIRE(false, CROG_E_BADSHAPE, "gemm_tc: unknown tile_cfg %d", g->tile_cfg);
    }
  }
  if (conv3_ok && !getenv("CROG_GEMM_NO_CONV3")) return launch<64, 5, 2, MODE, true>(g, stream);
  if (g->N <= 64) return launch<64, 5, 3, MODE>(g, stream);
  // 3x3 convolutions on the zero-haloed layout: activation bands (a third of the activation traffic into the SMs, which
  // is what bounds them: measured +20-40 % on every such layer of the forward); CROG_GEMM_NO_BAND restores the per-tap loop
  if (band_ok && !getenv("CROG_GEMM_NO_BAND")) {
    if (g->N % 256 == 0) {
      if (pair_ok && g->M >= 2 * 256 * 74) return launch<256, 3, 2, MODE, false, 4, true, 6>(g, stream);
      return launch<256, 3, 2, MODE, false, 4, false, 4>(g, stream);
    }
    return launch<128, 3, 2, MODE, false, 8, false, 6>(g, stream);
  }
  // 128 x 256 tiles cut the L2 -> smem operand traffic per FLOP by 25 % ((BM+BN)/(BM*BN)); worth it when the
  // contraction is long enough to be tensor/L2 bound rather than epilogue bound
  if (g->N % 256 == 0 && Ktot >= 1024) {
    // long contractions: the TMA feed is the pace (96 B/clk/SM at full tensor rate), so a fourth 48 KB stage in flight
    // is worth more than a second epilogue group (measured 3-12 % per layer; CROG_GEMM_3STAGE restores the old config)
    // CTA pairs (cta_group::2) whenever there is at least one 256-row tile per pair: +2-5 % on the long convolutions
    // (proj.vis.3: 1350 -> 1422 TFLOP/s); CROG_GEMM_PAIR=0 falls back to single-CTA tiles
    const char* pe = getenv("CROG_GEMM_PAIR");  // read per call (launches are graph-captured; tests toggle it)
    if (g->w_sample_stride == 0 && g->M >= 256 * 74 && !(pe && pe[0] == '0')) return launch<256, 6, 2, MODE, false, 4, true>(g, stream);
    if (!getenv("CROG_GEMM_3STAGE")) return launch<256, 4, 2, MODE, false, 4>(g, stream);
    return launch<256, 3, 2, MODE>(g, stream);
  }
  // short contractions are epilogue bound: three operand stages leave room for the shared-memory scale / bias slice
  if (Ktot < 512) return launch<128, 3, 3, MODE>(g, stream);
  return launch<128, 4, 3, MODE>(g, stream);
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
  if (g->a2) {
    CROG_REQUIRE(g->taps == 1 && g->cin2 > 0 && g->cin2 % BK == 0 && g->a2_ld % 8 == 0 && aligned16(g->a2) && g->w_sample_stride == 0,
                 CROG_E_BADSHAPE, "gemm_tc: the second activation operand needs taps == 1, cin2 a multiple of %d, shared weights", BK);
  } else {
    CROG_REQUIRE(g->cin2 == 0, CROG_E_BADSHAPE, "gemm_tc: cin2 without a2");
  }
  if (g->row_stats_out || g->row_stats_in) {
    CROG_REQUIRE(g->in_padded == 0 && g->out_padded == 0 && g->out_sample_rows == 0 && g->w_sample_stride == 0, CROG_E_BADSHAPE,
                 "gemm_tc: folded LayerNorm needs the identity row mapping");
    CROG_REQUIRE(g->row_stats_chunks > 0 && g->row_stats_chunks % 2 == 0 && g->row_stats_width > 0, CROG_E_BADSHAPE,
                 "gemm_tc: row_stats_chunks (even) / row_stats_width");
    CROG_REQUIRE(!g->row_stats_in || (g->scale && g->bias), CROG_E_BADSHAPE, "gemm_tc: the folded LayerNorm consumer needs scale (= s) and bias (= c)");
    CROG_REQUIRE(!g->row_stats_out || (g->out_dtype == CROG_BF16 && g->N % 64 == 0 && g->row_stats_chunks == g->N / 64 && g->N > 16),
                 CROG_E_BADSHAPE, "gemm_tc: the row-statistics producer writes bf16 through the TMA epilogue, N a multiple of 64");
  }
  // TMA epilogue: output row == enumerated row (so the tile is one box of the output matrix), shared weights
  const bool identity = (g->H == 0) || (g->in_padded == g->out_padded);
  const int esz = g->out_dtype == CROG_BF16 ? 2 : 4;
  const bool tma_ok = identity && g->w_sample_stride == 0 && g->out_sample_rows == 0 && g->N > 16 && ((long long)g->out_ld * esz) % 16 == 0 &&
                      (!g->residual || ((long long)g->res_ld * esz) % 16 == 0) && !getenv("CROG_GEMM_LEGACY_EPILOGUE");
  CROG_REQUIRE(!g->row_stats_out || tma_ok, CROG_E_BADSHAPE, "gemm_tc: the row-statistics producer needs the TMA epilogue (16-byte aligned rows, identity row mapping)");
  if (tma_ok) return g->out_dtype == CROG_BF16 ? dispatch<MODE_TMA_BF16>(g, stream) : dispatch<MODE_TMA_F32>(g, stream);
  if (g->N <= 16) return launch<16, 8, 2, MODE_LEGACY>(g, stream);
  return dispatch<MODE_LEGACY>(g, stream);
}

int crog_encode_2d(CUtensorMap* tm, const void* base, int dtype, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_rows) {
  static CrogEncodeTiledFn enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<CrogEncodeTiledFn>(p);
  }
  CROG_REQUIRE(enc != nullptr, CROG_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const bool f32 = dtype == CROG_F32;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * (f32 ? 4u : 2u)};
  cuuint32_t box[2] = {f32 ? 32u : 64u, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CROG_REQUIRE(r == CUDA_SUCCESS, CROG_E_CUDA, "cuTensorMapEncodeTiled failed (%d): cols=%llu rows=%llu ld=%llu", (int)r,
               (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld_elems);
  return CROG_OK;
}
