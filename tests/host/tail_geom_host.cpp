// Host-compiled self-check of crog_b200/csrc/tail_geom.h (same header the CUDA kernels use).
#include <math.h>
#include "../../crog_b200/csrc/tail_geom.h"

static int rect_area_fast(const TgRect* R, uint32_t* masks /*[128][5]*/) {
  int cnt = 0;
  for (int X = R->x0; X <= R->x1; ++X) {
    uint32_t* row = masks + (X - R->x0) * TG_WORDS;
    tg_row_mask(R, X, row);
    for (int w = 0; w < TG_WORDS; ++w) cnt += tg_popc(row[w]);
  }
  return cnt;
}
static int slow_count(const TgRect* A, const TgRect* B) {
  int x0 = B ? (A->x0 > B->x0 ? A->x0 : B->x0) : A->x0, x1 = B ? (A->x1 < B->x1 ? A->x1 : B->x1) : A->x1;
  int y0 = B ? (A->y0 > B->y0 ? A->y0 : B->y0) : A->y0, y1 = B ? (A->y1 < B->y1 ? A->y1 : B->y1) : A->y1;
  int cnt = 0;
  for (int X = x0; X <= x1; ++X)
    for (int Y = y0; Y <= y1; ++Y)
      if (tg_point_painted(A, X, Y) && (!B || tg_point_painted(B, X, Y))) ++cnt;
  return cnt;
}
extern "C" void tg_host_box_points(float cx, float cy, float w, float h, float ang, float* o8) { tg_box_points(cx, cy, w, h, ang, o8); }
// mode 0: fast path where available (as the kernel would), 1: force slow path
extern "C" void tg_host_counts(const double* rect_p, const double* rect_g, int mode, int* inter, int* uni, int* fast) {
  TgRect P, G;
  tg_make_rect(rect_p, &P);
  tg_make_rect(rect_g, &G);
  static uint32_t mp[TG_MAXROWS * TG_WORDS], mg[TG_MAXROWS * TG_WORDS];
  *fast = P.fast && G.fast;
  int ap, ag, in;
  if (mode == 0 && P.fast) ap = rect_area_fast(&P, mp); else ap = slow_count(&P, 0);
  if (mode == 0 && G.fast) ag = rect_area_fast(&G, mg); else ag = slow_count(&G, 0);
  if (mode == 0 && P.fast && G.fast) {
    in = 0;
    for (int X = G.x0; X <= G.x1; ++X) {
      if (X < P.x0 || X > P.x1) continue;
      const uint32_t* gr = mg + (X - G.x0) * TG_WORDS;
      const uint32_t* pr = mp + (X - P.x0) * TG_WORDS;
      int dw = G.yw0 - P.yw0;
      for (int w = 0; w < TG_WORDS; ++w) { int pw = w + dw; if (pw >= 0 && pw < TG_WORDS) in += tg_popc(gr[w] & pr[pw]); }
    }
  } else in = slow_count(&P, &G);
  *inter = in; *uni = ap + ag - in;
}
