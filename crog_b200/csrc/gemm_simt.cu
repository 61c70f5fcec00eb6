// CUDA-core implicit-GEMM (fp32 accumulate) with the same descriptor and epilogue as the
// tcgen05 kernel.  It is the fp32 ("strict parity") mode of the forward, and the
// cross-check for the tensor-core path; not a fallback: bf16 contractions default to tcgen05.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 32, PAD = 4;

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const CrogGemm g, int n_tiles) {
  __shared__ float As[BK][BM + PAD];
  __shared__ float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int n_t = blockIdx.x % n_tiles, m_t = blockIdx.x / n_tiles;
  const TileRows tr = tile_rows(g, m_t, BM);
  const int n0 = n_t * BN;
  const T* A = reinterpret_cast<const T*>(g.a);
  const T* Wt = reinterpret_cast<const T*>(g.w) + (long long)tr.wsample * g.w_sample_stride;
  const int Ktot = g.taps * g.cin;
  const int kchunks = g.cin / BK;

  const int lrow = tid >> 2, lk = (tid & 3) * 8;  // loader mapping: 64 rows x 4 chunks of 8
  const int ty = tid >> 4, tx = tid & 15;          // compute mapping: 16x16 threads, 4x4 each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const long long arow_base = tr.row0 + lrow;
  const int wn = n0 + lrow;
  for (int kb = 0; kb < g.taps * kchunks; ++kb) {
    const int tap = kb / kchunks, c0 = (kb % kchunks) * BK;
    {
      float v[8];
      const long long r = arow_base + tap_shift(g.taps, tap, g.W);
      if (arow_base < tr.row_end && r >= 0 && r < g.a_rows) load8(A + r * g.a_ld + c0 + lk, v);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) As[lk + i][lrow] = v[i];
      if (wn < g.N) load8(Wt + (long long)wn * Ktot + tap * g.cin + c0 + lk, v);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) Bs[lk + i][lrow] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // epilogue through smem so each thread finishes 8 consecutive columns of one row
  __shared__ float Cs[BM][BN + PAD];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Cs[ty * 4 + i][tx * 4 + j] = acc[i][j];
  __syncthreads();
  for (int it = tid; it < BM * (BN / 8); it += 256) {
    const int row = it / (BN / 8), cg = (it % (BN / 8)) * 8;
    RowMap m = map_row(g, tr.row0 + row, tr.row_end);
    load_row_stats(g, m);
    if (!m.valid || n0 + cg >= g.N) continue;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = Cs[row][cg + i];
    epilogue_row<8>(g, m, n0 + cg, v, g.scale ? g.scale + n0 + cg : nullptr, g.bias ? g.bias + n0 + cg : nullptr);
  }
}

}  // namespace

int crog_gemm_simt(const CrogGemm* g, cudaStream_t stream) {
  CROG_REQUIRE(g->cin % BK == 0, CROG_E_BADSHAPE, "gemm_simt: cin %d not a multiple of %d", g->cin, BK);
  CROG_REQUIRE(!g->a2 && g->cin2 == 0, CROG_E_BADSHAPE, "gemm_simt: the second activation operand exists on the tcgen05 path only");
  CROG_REQUIRE(!g->row_stats_out, CROG_E_BADSHAPE, "gemm_simt: the row-statistics producer exists on the tcgen05 path only");
  const int n_tiles = (g->N + BN - 1) / BN;
  const int grid = num_m_tiles(*g, BM) * n_tiles;
  if (grid == 0) return CROG_OK;
  if (g->dtype == CROG_F32) gemm_simt_kernel<float><<<grid, 256, 0, stream>>>(*g, n_tiles);
  else gemm_simt_kernel<bf16><<<grid, 256, 0, stream>>>(*g, n_tiles);
  CROG_LAUNCH_OK("gemm_simt");
  return CROG_OK;
}
