// Letterbox warps around the model (rows f-1 / f-2 of the scope table), bit-exact with OpenCV's
// cv2.warpAffine(..., INTER_CUBIC, BORDER_CONSTANT) as the reference calls it:
//
//   crog_warp_affine_cubic_f32  engine/crog_engine.py:387-391,499-517 — inverse letterbox of the float32 prediction
//                               (and target) maps from the network resolution to ori_size, borderValue = 0
//   crog_preprocess_u8          utils/dataset.py:843-866 — forward letterbox of the uint8 RGB image with the CLIP-mean
//                               border, then /255, -mean, /std into the NCHW float32 tensor forward() takes
//   crog_mask_iou               engine/crog_engine.py:500-501,515-518 — (pred > thr) vs (target != 0) pixel counts
//
// OpenCV's algorithm (imgwarp.cpp: warpAffine / WarpAffineInvoker / initInterTab2D / remapBicubic), restated in
// oracle/warp_affine.py and pinned there against the real cv2: source coordinates in 22.10 fixed point rounded to
// 1/32 pixel, weights = outer product of the 1-D cubic (A = -0.75) coefficients at that phase — float32 for float
// images, int16 (2^15 scale, forced to sum to 2^15) for uint8 — and a summation order that depends on whether the
// 4x4 footprint is inside the image.  Nothing here may be contracted into FMAs: every float operation is an explicit
// round-to-nearest intrinsic.
//
// Both kernels are pure gathers: one thread per destination pixel computes its coordinates and weights once and
// applies them to every plane / channel, reads go through L1/L2 (the source of a sample is 0.7-3.5 MB), writes are
// fully coalesced.  HBM bound: bytes = source read once + destination written once.
#include "common.cuh"

namespace {

constexpr int AB_BITS = 10, INTER_BITS = 5, TAB = 32, ROUND_DELTA = (1 << AB_BITS) / TAB / 2;
constexpr int COEF_BITS = 15, COEF_SCALE = 1 << COEF_BITS;

__device__ __forceinline__ int sat_int(double v) {  // saturate_cast<int>(double): round half to even
  if (!(v > -2147483648.0)) return (int)0x80000000;
  if (!(v < 2147483647.0)) return 0x7fffffff;
  return __double2int_rn(v);
}
__device__ __forceinline__ int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }
// c + a.lo16 * b.byte0 + a.hi16 * b.byte1 (lo) / ... byte2, byte3 (hi): signed 16-bit halves times UNSIGNED bytes
__device__ __forceinline__ int dp2a_lo_su(int a, uint32_t b, int c) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi_su(int a, uint32_t b, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// interpolateCubic(x = phase / 32, coeffs) in float32 without contraction
__device__ __forceinline__ void cubic_coeffs(int phase, float (&c)[4]) {
  const float A = -0.75f;
  const float x = __fmul_rn((float)phase, 1.0f / TAB);
  const float x1 = __fadd_rn(x, 1.f);
  c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), __fmul_rn(5.f, A)), x1), __fmul_rn(8.f, A)), x1),
                   __fmul_rn(4.f, A));
  const float a2 = __fadd_rn(A, 2.f), a3 = __fadd_rn(A, 3.f);
  c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(a2, x), a3), x), x), 1.f);
  const float xm = __fsub_rn(1.f, x);
  c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(a2, xm), a3), xm), xm), 1.f);
  c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
}

struct SrcCoord { int sx, sy, ax, ay; };

// Minv: the already inverted 2x3 matrix (float64).  (x, y) = destination pixel.
__device__ __forceinline__ SrcCoord src_coord(const double* __restrict__ M, int x, int y) {
  const int adelta = sat_int(__dmul_rn(__dmul_rn(M[0], (double)x), 1024.0));
  const int bdelta = sat_int(__dmul_rn(__dmul_rn(M[3], (double)x), 1024.0));
  const int X0 = sat_int(__dmul_rn(__dadd_rn(__dmul_rn(M[1], (double)y), M[2]), 1024.0)) + ROUND_DELTA;
  const int Y0 = sat_int(__dmul_rn(__dadd_rn(__dmul_rn(M[4], (double)y), M[5]), 1024.0)) + ROUND_DELTA;
  const int X = (int)((unsigned)X0 + (unsigned)adelta) >> (AB_BITS - INTER_BITS);
  const int Y = (int)((unsigned)Y0 + (unsigned)bdelta) >> (AB_BITS - INTER_BITS);
  SrcCoord c;
  c.sx = sat_short(X >> INTER_BITS) - 1;
  c.sy = sat_short(Y >> INTER_BITS) - 1;
  c.ax = X & (TAB - 1);
  c.ay = Y & (TAB - 1);
  return c;
}

// The same coordinates for a 32 x 8 pixel block: OpenCV itself computes adelta / bdelta once per destination column and
// X0 / Y0 once per row; a CTA does the 40 float64 evaluations once and a pixel adds two table entries.
struct BlockCoords { int ad[32], bd[32], x0[8], y0[8]; };
__device__ __forceinline__ void block_coords_fill(BlockCoords* t, const double* __restrict__ M, int xb, int yb, int tid) {
  if (tid < 32) {
    const double x = (double)(xb + tid);
    t->ad[tid] = sat_int(__dmul_rn(__dmul_rn(M[0], x), 1024.0));
    t->bd[tid] = sat_int(__dmul_rn(__dmul_rn(M[3], x), 1024.0));
  } else if (tid < 40) {
    const double y = (double)(yb + tid - 32);
    t->x0[tid - 32] = sat_int(__dmul_rn(__dadd_rn(__dmul_rn(M[1], y), M[2]), 1024.0)) + ROUND_DELTA;
    t->y0[tid - 32] = sat_int(__dmul_rn(__dadd_rn(__dmul_rn(M[4], y), M[5]), 1024.0)) + ROUND_DELTA;
  }
}
__device__ __forceinline__ SrcCoord block_coord(const BlockCoords* t, int tx, int ty) {
  const int X = (int)((unsigned)t->x0[ty] + (unsigned)t->ad[tx]) >> (AB_BITS - INTER_BITS);
  const int Y = (int)((unsigned)t->y0[ty] + (unsigned)t->bd[tx]) >> (AB_BITS - INTER_BITS);
  SrcCoord c;
  c.sx = sat_short(X >> INTER_BITS) - 1;
  c.sy = sat_short(Y >> INTER_BITS) - 1;
  c.ax = X & (TAB - 1);
  c.ay = Y & (TAB - 1);
  return c;
}

// Second form of the float warp: 32 x 8 pixel blocks (no index division), the 32 x 4 table of cubic coefficients built once
// per CTA in shared memory (a lookup instead of ~50 non-contractable flops per pixel), the 16 weight products of a pixel
// computed once for all planes (they are OpenCV's table entries, so their rounding is unchanged), plane-to-plane pointer
// increments, registers held to 64 for four CTAs per SM.  Same operations in the same order per output: bit-identical.
__global__ void __launch_bounds__(256, 4) warp_cubic_f32_v2_kernel(const float* __restrict__ src, int NP, int B, int Hs, int Ws,
                                                                   const double* __restrict__ minv, float* __restrict__ dst, int h,
                                                                   int w, float cv) {
  __shared__ float s_tab[TAB][4];
  __shared__ BlockCoords s_bc;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int b = blockIdx.z;
  if (tid < TAB) {
    float c4[4];
    cubic_coeffs(tid, c4);
    s_tab[tid][0] = c4[0]; s_tab[tid][1] = c4[1]; s_tab[tid][2] = c4[2]; s_tab[tid][3] = c4[3];
  }
  block_coords_fill(&s_bc, minv + b * 6, blockIdx.x * 32, blockIdx.y * 8, 255 - tid);  // the other end of the CTA
  __syncthreads();
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= w || y >= h) return;
  const SrcCoord c = block_coord(&s_bc, threadIdx.x, threadIdx.y);
  const long long splane = (long long)Hs * Ws, dplane = (long long)h * w;
  const long long sps = (long long)B * splane, dps = (long long)B * dplane;
  float* D = dst + (long long)b * dplane + (long long)y * w + x;
  const bool outside = c.sx >= Ws || c.sx + 4 <= 0 || c.sy >= Hs || c.sy + 4 <= 0;
  if (outside) {
    for (int pl = 0; pl < NP; ++pl, D += dps) *D = cv;
    return;
  }
  const float4 vx4 = *reinterpret_cast<const float4*>(s_tab[c.ax]), vy4 = *reinterpret_cast<const float4*>(s_tab[c.ay]);
  const float vx[4] = {vx4.x, vx4.y, vx4.z, vx4.w}, vy[4] = {vy4.x, vy4.y, vy4.z, vy4.w};
  const bool inside = c.sx >= 0 && c.sx < max(Ws - 3, 0) && c.sy >= 0 && c.sy < max(Hs - 3, 0);
  if (inside) {
    float wgt[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) wgt[i * 4 + j] = __fmul_rn(vy[i], vx[j]);
    const float* S = src + (long long)b * splane + (long long)c.sy * Ws + c.sx;
    for (int pl = 0; pl < NP; ++pl, S += sps, D += dps) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i * 4 + j] = __ldg(S + i * Ws + j);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float r = __fmul_rn(v[i * 4], wgt[i * 4]);
        r = __fadd_rn(r, __fmul_rn(v[i * 4 + 1], wgt[i * 4 + 1]));
        r = __fadd_rn(r, __fmul_rn(v[i * 4 + 2], wgt[i * 4 + 2]));
        r = __fadd_rn(r, __fmul_rn(v[i * 4 + 3], wgt[i * 4 + 3]));
        sum = i == 0 ? r : __fadd_rn(sum, r);
      }
      *D = sum;
    }
  } else {
    for (int pl = 0; pl < NP; ++pl, D += dps) {
      const float* S = src + ((long long)pl * B + b) * splane;
      float sum = __fmul_rn(cv, 1.f);
      for (int i = 0; i < 4; ++i) {
        const int yy = c.sy + i;
        if (yy < 0 || yy >= Hs) continue;
        for (int j = 0; j < 4; ++j) {
          const int xx = c.sx + j;
          if (xx < 0 || xx >= Ws) continue;
          sum = __fadd_rn(sum, __fmul_rn(__fsub_rn(__ldg(S + (long long)yy * Ws + xx), cv), __fmul_rn(vy[i], vx[j])));
        }
      }
      *D = sum;
    }
  }
}

// src [NP][B][Hs][Ws] -> dst [NP][B][h][w]; grid (ceil(w*h/256), B)
__global__ void __launch_bounds__(256) warp_cubic_f32_kernel(const float* __restrict__ src, int NP, int B, int Hs, int Ws,
                                                             const double* __restrict__ minv, float* __restrict__ dst, int h,
                                                             int w, float cv) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= h * w) return;
  const int y = p / w, x = p - y * w;
  const SrcCoord c = src_coord(minv + b * 6, x, y);
  const long long splane = (long long)Hs * Ws, dplane = (long long)h * w;
  const bool outside = c.sx >= Ws || c.sx + 4 <= 0 || c.sy >= Hs || c.sy + 4 <= 0;
  if (outside) {
    for (int pl = 0; pl < NP; ++pl) dst[((long long)pl * B + b) * dplane + p] = cv;
    return;
  }
  float vx[4], vy[4];
  cubic_coeffs(c.ax, vx);
  cubic_coeffs(c.ay, vy);
  const bool inside = c.sx >= 0 && c.sx < max(Ws - 3, 0) && c.sy >= 0 && c.sy < max(Hs - 3, 0);
  if (inside) {
    for (int pl = 0; pl < NP; ++pl) {
      const float* S = src + ((long long)pl * B + b) * splane + (long long)c.sy * Ws + c.sx;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float* R = S + i * Ws;
        float r = __fmul_rn(__ldg(R), __fmul_rn(vy[i], vx[0]));
        r = __fadd_rn(r, __fmul_rn(__ldg(R + 1), __fmul_rn(vy[i], vx[1])));
        r = __fadd_rn(r, __fmul_rn(__ldg(R + 2), __fmul_rn(vy[i], vx[2])));
        r = __fadd_rn(r, __fmul_rn(__ldg(R + 3), __fmul_rn(vy[i], vx[3])));
        sum = i == 0 ? r : __fadd_rn(sum, r);
      }
      dst[((long long)pl * B + b) * dplane + p] = sum;
    }
  } else {
    for (int pl = 0; pl < NP; ++pl) {
      const float* S = src + ((long long)pl * B + b) * splane;
      float sum = __fmul_rn(cv, 1.f);
      for (int i = 0; i < 4; ++i) {
        const int yy = c.sy + i;
        if (yy < 0 || yy >= Hs) continue;
        for (int j = 0; j < 4; ++j) {
          const int xx = c.sx + j;
          if (xx < 0 || xx >= Ws) continue;
          sum = __fadd_rn(sum, __fmul_rn(__fsub_rn(__ldg(S + (long long)yy * Ws + xx), cv), __fmul_rn(vy[i], vx[j])));
        }
      }
      dst[((long long)pl * B + b) * dplane + p] = sum;
    }
  }
}

// BicubicTab_i entry for the phase pair: saturate_cast<short>(v * 32768), then the entry of the {2,3}x{2,3} block that is
// largest (sum too small) or smallest (sum too large) absorbs the difference (initInterTab2D)
__device__ __forceinline__ void cubic_weights_i16(int ax, int ay, int (&wt)[16]) {
  float vx[4], vy[4];
  cubic_coeffs(ax, vx);
  cubic_coeffs(ay, vy);
  int isum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = sat_short(__float2int_rn(__fmul_rn(__fmul_rn(vy[i], vx[j]), (float)COEF_SCALE)));
      wt[i * 4 + j] = q;
      isum += q;
    }
  if (isum != COEF_SCALE) {
    const int diff = isum - COEF_SCALE;
    int Mk = 10, mk = 10;  // index 2*4+2
#pragma unroll
    for (int k1 = 2; k1 < 4; ++k1)
#pragma unroll
      for (int k2 = 2; k2 < 4; ++k2) {
        const int k = k1 * 4 + k2;
        if (wt[k] < wt[mk]) mk = k;
        else if (wt[k] > wt[Mk]) Mk = k;
      }
    if (diff < 0) wt[Mk] = (short)(wt[Mk] - diff);
    else wt[mk] = (short)(wt[mk] - diff);
  }
}

// OpenCV's BicubicTab_i: the 16 int16 weights of every (ay, ax) phase pair, built on the device by the same routine the
// per-pixel form used (so the table is bit-identical to it); 32 KB, rebuilt by every crog_preprocess_u8 call (~2 us).
__global__ void __launch_bounds__(256) cubic_table_i16_kernel(short* __restrict__ tab) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;  // = ay * 32 + ax
  if (t >= TAB * TAB) return;
  int wt[16];
  cubic_weights_i16(t % TAB, t / TAB, wt);
#pragma unroll
  for (int k = 0; k < 16; ++k) tab[t * 16 + k] = (short)wt[k];
}

// img [B][Ho][Wo][3] uint8 (RGB, HWC) -> out [B][3][S][S] float32 normalised; grid (ceil(S*S/256), B)
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const short* __restrict__ wtab, const uint8_t* __restrict__ img, int B, int Ho, int Wo,
                                                            const double* __restrict__ minv, float* __restrict__ out, int Sh,
                                                            int Sw, int cv0, int cv1, int cv2, float m0, float m1, float m2,
                                                            float s0, float s1, float s2, long long total_bytes) {
  // the result pixel is an integer 0..255: (v / 255 - mean) / std has 256 possible values per channel, computed once per
  // CTA with the reference's two IEEE divisions instead of six divisions per pixel
  __shared__ float lut[3][256];
  __shared__ BlockCoords s_bc;
  const int tid = threadIdx.y * 32 + threadIdx.x, b = blockIdx.z;
  {
    const float v = __fdiv_rn((float)tid, 255.f);
    lut[0][tid] = __fdiv_rn(__fsub_rn(v, m0), s0);
    lut[1][tid] = __fdiv_rn(__fsub_rn(v, m1), s1);
    lut[2][tid] = __fdiv_rn(__fsub_rn(v, m2), s2);
  }
  block_coords_fill(&s_bc, minv + b * 6, blockIdx.x * 32, blockIdx.y * 8, tid);
  __syncthreads();
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= Sw || y >= Sh) return;
  const int p = y * Sw + x;
  const SrcCoord c = block_coord(&s_bc, threadIdx.x, threadIdx.y);
  const int cval[3] = {cv0, cv1, cv2};
  int res[3] = {cv0, cv1, cv2};
  const bool outside = c.sx >= Wo || c.sx + 4 <= 0 || c.sy >= Ho || c.sy + 4 <= 0;
  if (!outside) {
    const int4* e = reinterpret_cast<const int4*>(wtab + (c.ay * TAB + c.ax) * 16);
    const int4 e0 = __ldg(e), e1 = __ldg(e + 1);
    const int q[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};  // the 16 taps as (even, odd) 16-bit pairs
    int acc[3] = {cv0 * COEF_SCALE, cv1 * COEF_SCALE, cv2 * COEF_SCALE};
    const uint8_t* S = img + (long long)b * Ho * Wo * 3;
    // Interior footprint (the common case): the 4 pixels x 3 channels of a footprint row are 12 contiguous bytes.  They are
    // fetched as four ALIGNED 32-bit words and realigned with funnel shifts - 16 word loads per pixel instead of 48 byte
    // loads; integer accumulation, so the order of the terms does not matter.  (The window may start up to 3 bytes before
    // the row's first byte and ends at most 15 bytes after it: it must lie inside the batch's allocation.)
    const long long off00 = ((long long)b * Ho * Wo + (long long)c.sy * Wo + c.sx) * 3;
    const bool fast = c.sx >= 0 && c.sx + 3 < Wo && c.sy >= 0 && c.sy + 3 < Ho && (reinterpret_cast<uintptr_t>(img) & 3) == 0 &&
                      ((off00 + 3LL * 3 * Wo) & ~3LL) + 16 <= total_bytes;
    if (fast) {
      // the weight table stores the taps as 16-bit pairs, which is the A operand of dp2a (two 16-bit x 8-bit products
      // per instruction): the row's bytes are regrouped per channel with two byte permutes, and a channel's four taps
      // are two dp2a.  sum(r * w) - cv * sum(w) + cv * SCALE is the same integer as sum((r - cv) * w) + cv * SCALE.
      int sw = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) sw = dp2a_lo_su(q[k], 0x0101u, sw);
      int a0 = cv0 * (COEF_SCALE - sw), a1 = cv1 * (COEF_SCALE - sw), a2 = cv2 * (COEF_SCALE - sw);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long off = off00 + (long long)i * Wo * 3;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(img + (off & ~3LL));
        const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3);
        const uint32_t sh = (uint32_t)(off & 3) * 8u;
        const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = __funnelshift_r(w2, w3, sh);
        // bytes: v0 = R0 G0 B0 R1, v1 = G1 B1 R2 G2, v2 = B2 R3 G3 B3
        const uint32_t rr = __byte_perm(__byte_perm(v0, v1, 0x0630), v2, 0x5210);  // R0 R1 R2 R3
        const uint32_t gg = __byte_perm(__byte_perm(v0, v1, 0x0741), v2, 0x6210);  // G0 G1 G2 G3
        const uint32_t bb = __byte_perm(__byte_perm(v0, v1, 0x0052), v2, 0x7410);  // B0 B1 B2 B3
        a0 = dp2a_hi_su(q[2 * i + 1], rr, dp2a_lo_su(q[2 * i], rr, a0));
        a1 = dp2a_hi_su(q[2 * i + 1], gg, dp2a_lo_su(q[2 * i], gg, a1));
        a2 = dp2a_hi_su(q[2 * i + 1], bb, dp2a_lo_su(q[2 * i], bb, a2));
      }
      acc[0] = a0; acc[1] = a1; acc[2] = a2;
    } else {
      int wt[16];
#pragma unroll
      for (int k = 0; k < 8; ++k) { wt[2 * k] = (int)(short)(q[k] & 0xffff); wt[2 * k + 1] = q[k] >> 16; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int yy = c.sy + i;
        if (yy < 0 || yy >= Ho) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int xx = c.sx + j;
          if (xx < 0 || xx >= Wo) continue;
          const uint8_t* px = S + ((long long)yy * Wo + xx) * 3;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) acc[ch] += ((int)__ldg(px + ch) - cval[ch]) * wt[i * 4 + j];
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const int v = (acc[ch] + (1 << (COEF_BITS - 1))) >> COEF_BITS;
      res[ch] = v < 0 ? 0 : (v > 255 ? 255 : v);
    }
  }
  const long long plane = (long long)Sh * Sw;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) out[((long long)b * 3 + ch) * plane + p] = lut[ch][res[ch]];
}

// inter/union pixel counts of (pred > thr) vs (target != 0), one CTA row per sample
__global__ void __launch_bounds__(256) mask_iou_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                       long long n, float thr, unsigned long long* __restrict__ counts) {
  const int b = blockIdx.y;
  const float* P = pred + (long long)b * n;
  const float* T = target + (long long)b * n;
  int inter = 0, uni = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const bool p = P[i] > thr, t = T[i] != 0.f;
    inter += p && t;
    uni += p || t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    inter += __shfl_xor_sync(0xffffffffu, inter, o);
    uni += __shfl_xor_sync(0xffffffffu, uni, o);
  }
  if ((threadIdx.x & 31) == 0 && (inter | uni)) {
    atomicAdd(counts + b * 2, (unsigned long long)inter);
    atomicAdd(counts + b * 2 + 1, (unsigned long long)uni);
  }
}

}  // namespace

extern "C" int crog_warp_affine_cubic_f32(const float* src, int32_t NP, int32_t B, int32_t Hs, int32_t Ws, const double* minv,
                                          float* dst, int32_t h, int32_t w, float border_value, void* stream) {
  CROG_REQUIRE(NP >= 1 && B >= 0 && Hs >= 1 && Ws >= 1 && h >= 1 && w >= 1, CROG_E_BADSHAPE, "warp_affine: bad shape");
  CROG_REQUIRE(B <= 65535 && (long long)h * w < (1LL << 31) && Hs < 32768 && Ws < 32768, CROG_E_BADSHAPE, "warp_affine: size limits");
  if (B == 0) return CROG_OK;
  if ((h + 7) / 8 <= 65535 && !getenv("CROG_WARP_V1")) {
    warp_cubic_f32_v2_kernel<<<dim3((w + 31) / 32, (h + 7) / 8, B), dim3(32, 8), 0, (cudaStream_t)stream>>>(src, NP, B, Hs, Ws, minv, dst, h, w,
                                                                                                       border_value);
    CROG_LAUNCH_OK("warp_affine_cubic_f32");
    return CROG_OK;
  }
  dim3 grid((unsigned)(((long long)h * w + 255) / 256), (unsigned)B);
  warp_cubic_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, NP, B, Hs, Ws, minv, dst, h, w, border_value);
  CROG_LAUNCH_OK("warp_affine_cubic_f32");
  return CROG_OK;
}

extern "C" int64_t crog_preprocess_workspace_bytes(void) { return (int64_t)TAB * TAB * 16 * sizeof(short); }

extern "C" int crog_preprocess_u8(const uint8_t* img, int32_t B, int32_t Ho, int32_t Wo, const double* minv, float* out,
                                  int32_t Sh, int32_t Sw, const double* border_rgb, const float* mean, const float* std_,
                                  void* workspace, void* stream) {
  CROG_REQUIRE(workspace != nullptr && aligned16(workspace), CROG_E_BADALIGN, "preprocess: workspace (crog_preprocess_workspace_bytes) must be 16B aligned");
  CROG_REQUIRE(B >= 0 && Ho >= 1 && Wo >= 1 && Sh >= 1 && Sw >= 1, CROG_E_BADSHAPE, "preprocess: bad shape");
  CROG_REQUIRE(B <= 65535 && Ho < 32768 && Wo < 32768, CROG_E_BADSHAPE, "preprocess: size limits");
  if (B == 0) return CROG_OK;
  int cv[3];
  for (int i = 0; i < 3; ++i) {  // saturate_cast<uchar>(double): round half to even, clamp
    double r = nearbyint(border_rgb[i]);
    cv[i] = r < 0 ? 0 : (r > 255 ? 255 : (int)r);
  }
  cubic_table_i16_kernel<<<(TAB * TAB + 255) / 256, 256, 0, (cudaStream_t)stream>>>((short*)workspace);
  CROG_LAUNCH_OK("cubic_table_i16");
  CROG_REQUIRE((Sh + 7) / 8 <= 65535, CROG_E_BADSHAPE, "preprocess: output too tall");
  dim3 grid((unsigned)((Sw + 31) / 32), (unsigned)((Sh + 7) / 8), (unsigned)B);
  preprocess_u8_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((const short*)workspace, img, B, Ho, Wo, minv, out, Sh, Sw, cv[0], cv[1], cv[2], mean[0],
                                                               mean[1], mean[2], std_[0], std_[1], std_[2],
                                                               (long long)B * Ho * Wo * 3);
  CROG_LAUNCH_OK("preprocess_u8");
  return CROG_OK;
}

extern "C" int crog_mask_iou(const float* pred, const float* target, int32_t B, int64_t n, float thr, int64_t* counts,
                             void* stream) {
  CROG_REQUIRE(B >= 0 && B <= 65535 && n >= 0, CROG_E_BADSHAPE, "mask_iou: bad shape");
  if (B == 0) return CROG_OK;
  CROG_CUDA_OK(cudaMemsetAsync(counts, 0, (size_t)B * 2 * sizeof(int64_t), (cudaStream_t)stream));
  int gx = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  mask_iou_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(pred, target, n, thr, (unsigned long long*)counts);
  CROG_LAUNCH_OK("mask_iou");
  return CROG_OK;
}
