// Exact integer geometry of the Jaccard tail, shared by the CUDA kernels (tail.cu) and a
// host-compiled self-check (tests/test_tail_geom_host.py compiles this header with g++ and
// compares it with the oracle on CPU, so the trickiest logic is validated without a GPU).
//
// Semantics restated from utils/grasp_eval.py:305-347 of the reference:
//   corners = cv2.boxPoints(((cx,cy),(w,h),-theta)) in float32, truncated toward zero (np.int0);
//   pixels  = skimage.draw.polygon(r = X, c = Y, shape=(480,640)) then rr<640, cc<480, i.e. the
//             lattice points (X,Y) in [0,479]^2 for which O'Rourke's point-in-polygon test with
//             the ray along +Y is non-zero (inside, on an edge, or a vertex).
// For one scanline X the test is evaluated for all Y at once:
//   r_cross(Y) = #{edges with exactly one endpoint strictly above X whose intercept Y* > Y}
//   l_cross(Y) = #{edges with exactly one endpoint strictly below X whose intercept Y* < Y}
//   painted(Y) = vertex(Y) | odd(r_cross) | odd(l_cross)
// and "Y* > Y" for integer Y is Y <= ceil(Y*)-1, "Y* < Y" is Y >= floor(Y*)+1, so each parity is
// an XOR of prefix / suffix bit masks.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define TG_HD __host__ __device__ __forceinline__
#else
#define TG_HD static inline
#endif

#define TG_RASTER 480      // raster window is [0,479]^2 (A.4 quirk: x >= 480 is dropped)
#define TG_WORDS 5         // 160-bit rows anchored at a multiple of 32 cover any 128-wide span
#define TG_MAXROWS 128
#define TG_COORD_LIM 16384 // |coord| below this keeps every product inside int32

struct TgRect {
  int X[4], Y[4];  // integer corners (X = image x = polygon "r", Y = image y = polygon "c")
  int x0, x1;      // clipped scanline range (inclusive); empty if x0 > x1
  int y0, y1;      // clipped bit range (inclusive)
  int yw0;         // first word index: y0 >> 5
  int fast;        // fits the 128-row x 160-bit fast path and int32 arithmetic
};

TG_HD int tg_floordiv(int p, int q) {  // q > 0
  int d = p / q;
  if ((p % q) != 0 && p < 0) --d;
  return d;
}
TG_HD int tg_ceildiv(int p, int q) {  // q > 0
  int d = p / q;
  if ((p % q) != 0 && p > 0) ++d;
  return d;
}

// Corner arithmetic of cv::RotatedRect::points in float32 without FMA contraction.
#if defined(__CUDA_ARCH__)
#define TG_MUL(a, b) __fmul_rn((a), (b))
#define TG_ADD(a, b) __fadd_rn((a), (b))
#define TG_SUB(a, b) __fsub_rn((a), (b))
#else
#define TG_MUL(a, b) ((float)((a) * (b)))
#define TG_ADD(a, b) ((float)((a) + (b)))
#define TG_SUB(a, b) ((float)((a) - (b)))
#endif

TG_HD void tg_box_points(float cx, float cy, float w, float h, float ang_deg, float* o8) {
  const double rad = (double)ang_deg * 3.141592653589793 / 180.0;
#if defined(__CUDA_ARCH__)
  double sn, cs;
  sincos(rad, &sn, &cs);  // one range reduction for both
#else
  const double sn = sin(rad), cs = cos(rad);
#endif
  const float b = TG_MUL((float)cs, 0.5f);
  const float a = TG_MUL((float)sn, 0.5f);
  o8[0] = TG_SUB(TG_SUB(cx, TG_MUL(a, h)), TG_MUL(b, w));
  o8[1] = TG_SUB(TG_ADD(cy, TG_MUL(b, h)), TG_MUL(a, w));
  o8[2] = TG_SUB(TG_ADD(cx, TG_MUL(a, h)), TG_MUL(b, w));
  o8[3] = TG_SUB(TG_SUB(cy, TG_MUL(b, h)), TG_MUL(a, w));
  o8[4] = TG_SUB(TG_MUL(2.f, cx), o8[0]);
  o8[5] = TG_SUB(TG_MUL(2.f, cy), o8[1]);
  o8[6] = TG_SUB(TG_MUL(2.f, cx), o8[2]);
  o8[7] = TG_SUB(TG_MUL(2.f, cy), o8[3]);
}

TG_HD int tg_trunc_clamp(float v) {  // np.int0 (truncate toward zero); clamp far-away values
  if (!(v > -1.0e9f)) return -1000000000;
  if (!(v < 1.0e9f)) return 1000000000;
  return (int)v;
}

// rect = (cx, cy, w, h, theta_deg) as doubles, like the reference's python floats.
TG_HD void tg_make_rect(const double* rect5, TgRect* R) {
  float p[8];
  tg_box_points((float)rect5[0], (float)rect5[1], (float)rect5[2], (float)rect5[3], (float)(-rect5[4]), p);
  int minx = 0x7fffffff, maxx = -0x7fffffff, miny = 0x7fffffff, maxy = -0x7fffffff, big = 0;
  for (int i = 0; i < 4; ++i) {
    R->X[i] = tg_trunc_clamp(p[2 * i]);
    R->Y[i] = tg_trunc_clamp(p[2 * i + 1]);
    if (R->X[i] < minx) minx = R->X[i];
    if (R->X[i] > maxx) maxx = R->X[i];
    if (R->Y[i] < miny) miny = R->Y[i];
    if (R->Y[i] > maxy) maxy = R->Y[i];
    if (R->X[i] <= -TG_COORD_LIM || R->X[i] >= TG_COORD_LIM || R->Y[i] <= -TG_COORD_LIM || R->Y[i] >= TG_COORD_LIM) big = 1;
  }
  R->x0 = minx < 0 ? 0 : minx;
  R->x1 = maxx > TG_RASTER - 1 ? TG_RASTER - 1 : maxx;
  R->y0 = miny < 0 ? 0 : miny;
  R->y1 = maxy > TG_RASTER - 1 ? TG_RASTER - 1 : maxy;
  R->yw0 = R->y0 >> 5;
  const int rows = R->x1 - R->x0 + 1;
  const int words = (R->y1 >> 5) - R->yw0 + 1;
  R->fast = !big && rows <= TG_MAXROWS && words <= TG_WORDS;
}

// bits [lo, hi] (absolute Y, inclusive) of word index wi (absolute: bit j of word wi is Y = 32*wi + j)
TG_HD uint32_t tg_word_range(int wi, int lo, int hi) {
  const int base = wi << 5;
  int a = lo - base, b = hi - base;
  if (a < 0) a = 0;
  if (b > 31) b = 31;
  if (a > b) return 0u;
  const uint32_t m_hi = (b == 31) ? 0xffffffffu : ((1u << (b + 1)) - 1u);
  const uint32_t m_lo = (a == 0) ? 0u : ((1u << a) - 1u);
  return m_hi & ~m_lo;
}

// Painted bits of scanline X for a rectangle with |coords| < TG_COORD_LIM: words [yw0, yw0+TG_WORDS),
// already clipped to [y0, y1].  out[TG_WORDS].
TG_HD void tg_row_mask(const TgRect* R, int X, uint32_t* out) {
  uint32_t rpar[TG_WORDS], lpar[TG_WORDS], vert[TG_WORDS];
  for (int w = 0; w < TG_WORDS; ++w) { rpar[w] = 0u; lpar[w] = 0u; vert[w] = 0u; }
  const int BIG = 1 << 29;
  int xb = R->X[3], yb = R->Y[3];  // previous vertex (pip starts from the last one)
  for (int i = 0; i < 4; ++i) {
    const int xa = R->X[i], ya = R->Y[i];
    // relative to the scanline: "y" of pip is the X axis here
    const int d0 = xa - X, d1 = xb - X;
    if (d0 == 0 && ya >= R->y0 && ya <= R->y1) {  // vertex on this scanline
      const int wi = (ya >> 5) - R->yw0;
      if (wi >= 0 && wi < TG_WORDS) vert[wi] |= 1u << (ya & 31);
    }
    const bool rs = (d0 > 0) != (d1 > 0);
    const bool ls = (d0 < 0) != (d1 < 0);
    if (rs || ls) {
      // intercept Y* = ya + (X - xa) * (yb - ya) / (xb - xa) = P / Q
      int Q = xb - xa;
      int P = ya * Q + (X - xa) * (yb - ya);
      if (Q < 0) { Q = -Q; P = -P; }
      const int fl = tg_floordiv(P, Q);             // one division serves both bounds
      const int ce = fl + ((P - fl * Q) != 0 ? 1 : 0);
      if (rs) {  // counts for Y <= ceil(Y*) - 1  -> prefix mask
        const int A = ce - 1;
        for (int w = 0; w < TG_WORDS; ++w) rpar[w] ^= tg_word_range(R->yw0 + w, -BIG, A);
      }
      if (ls) {  // counts for Y >= floor(Y*) + 1 -> suffix mask
        const int B = fl + 1;
        for (int w = 0; w < TG_WORDS; ++w) lpar[w] ^= tg_word_range(R->yw0 + w, B, BIG);
      }
    }
    xb = xa; yb = ya;
  }
  for (int w = 0; w < TG_WORDS; ++w)
    out[w] = (rpar[w] | lpar[w] | vert[w]) & tg_word_range(R->yw0 + w, R->y0, R->y1);
}

// Exact per-point test with 64-bit arithmetic (generic slow path): non-zero if (X,Y) is painted.
TG_HD int tg_point_painted(const TgRect* R, int X, int Y) {
  int rc = 0, lc = 0;
  long long x1 = (long long)R->Y[3] - Y, y1 = (long long)R->X[3] - X;  // pip's x = Y axis, y = X axis
  for (int i = 0; i < 4; ++i) {
    const long long x0 = (long long)R->Y[i] - Y, y0 = (long long)R->X[i] - X;
    if (x0 == 0 && y0 == 0) return 2;
    if ((y0 > 0) != (y1 > 0)) {
      const long long num = x0 * y1 - x1 * y0, den = y1 - y0;
      if ((num > 0 && den > 0) || (num < 0 && den < 0)) ++rc;
    }
    if ((y0 < 0) != (y1 < 0)) {
      const long long num = x0 * y1 - x1 * y0, den = y1 - y0;
      if ((num > 0 && den < 0) || (num < 0 && den > 0)) ++lc;
    }
    x1 = x0; y1 = y0;
  }
  if ((rc & 1) != (lc & 1)) return 3;
  return rc & 1;
}

TG_HD int tg_popc(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
