/* ORACLE (test infrastructure; never shipped, never linked into the product).
 *
 * Plain-C restatement of the reference's grasp-decode + Jaccard tail,
 * utils/grasp_eval.py:289-374, written independently of oracle/grasp_tail.py so the
 * two can be checked against each other.  Third-party semantics restated
 * (absent from /root/reference): scikit-image 0.20.0 peak_local_max / draw.polygon,
 * scipy 1.9.1 maximum_filter, opencv 4.7.0 boxPoints, numpy 1.24.3 promotion
 * (SURVEY.md Appendix A).  PINNED (round 2) against the reference's own lines
 * 289-374 executed by oracle/make_golden_tail.py (tests/golden/tail_cases.npz) and
 * against OpenCV / scipy cross-checks; the two scikit-image internals that stay
 * unverifiable offline are listed in oracle/grasp_tail.py's header.
 *
 * Also used as bench.py's cpu_baseline ("port") for the tail micro-benchmark.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC, -ffp-contract=off so float32
 * box-point arithmetic is not fused).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CANVAS_H 480
#define CANVAS_W 640

/* ---- peak_local_max(min_distance=2, threshold_abs=thr, num_peaks=K), grasp_eval.py:292 */
typedef struct { float v; int idx; } cand_t;

static int cand_cmp(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
    if (x->v > y->v) return -1;
    if (x->v < y->v) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx); /* stable: row-major order on ties */
}

int oracle_peak_local_max(const float* img, int H, int W, float thr, int K, int* out_rc) {
    /* 5x5 max with edge replication, separable */
    float* tmp = (float*)malloc(sizeof(float) * (size_t)H * W);
    float* mx = (float*)malloc(sizeof(float) * (size_t)H * W);
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            float m = img[(size_t)r * W + c];
            for (int d = -2; d <= 2; ++d) {
                int cc = c + d; cc = cc < 0 ? 0 : (cc >= W ? W - 1 : cc);
                float v = img[(size_t)r * W + cc]; if (v > m) m = v;
            }
            tmp[(size_t)r * W + c] = m;
        }
    int all_equal = 1;
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            float m = tmp[(size_t)r * W + c];
            for (int d = -2; d <= 2; ++d) {
                int rr = r + d; rr = rr < 0 ? 0 : (rr >= H ? H - 1 : rr);
                float v = tmp[(size_t)rr * W + c]; if (v > m) m = v;
            }
            mx[(size_t)r * W + c] = m;
            if (img[(size_t)r * W + c] != m) all_equal = 0;
        }
    int n = 0;
    cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * (size_t)H * W);
    if (!all_equal) {
        for (int r = 2; r < H - 2; ++r)
            for (int c = 2; c < W - 2; ++c) {
                float v = img[(size_t)r * W + c];
                if (v == mx[(size_t)r * W + c] && v > thr) { cand[n].v = v; cand[n].idx = r * W + c; ++n; }
            }
    }
    qsort(cand, (size_t)n, sizeof(cand_t), cand_cmp);
    int kept = 0;
    for (int i = 0; i < n && kept < K; ++i) {
        int r = cand[i].idx / W, c = cand[i].idx % W, ok = 1;
        for (int j = 0; j < kept; ++j) {
            int dr = abs(out_rc[2 * j] - r), dc = abs(out_rc[2 * j + 1] - c);
            if ((dr > dc ? dr : dc) < 2) { ok = 0; break; }
        }
        if (ok) { out_rc[2 * kept] = r; out_rc[2 * kept + 1] = c; ++kept; }
    }
    free(tmp); free(mx); free(cand);
    return kept;
}

/* ---- detect_grasps, grasp_eval.py:289-302; rows [x, y, w*100, 20, angle_deg] float64 */
int oracle_detect_grasps(const float* q, const float* s, const float* c, const float* w,
                         int H, int W, int K, double* out_k5, int* out_rc) {
    int n = oracle_peak_local_max(q, H, W, 0.4f, K, out_rc);
    for (int i = 0; i < n; ++i) {
        int r = out_rc[2 * i], col = out_rc[2 * i + 1];
        size_t o = (size_t)r * W + col;
        float ang = (float)atan2((double)s[o], (double)c[o]) * 0.5f; /* float32 angle map / 2.0 */
        out_k5[5 * i + 0] = (double)col;
        out_k5[5 * i + 1] = (double)r;
        out_k5[5 * i + 2] = (double)w[o] * 100.0;
        out_k5[5 * i + 3] = 20.0;
        out_k5[5 * i + 4] = (double)ang / 3.141592653589793 * 180.0;
    }
    return n;
}

/* ---- cv2.boxPoints in float32 */
void oracle_box_points(float cx, float cy, float w, float h, float ang, float* o8) {
    double rad = (double)ang * 3.141592653589793 / 180.0;
    float b = (float)cos(rad) * 0.5f;
    float a = (float)sin(rad) * 0.5f;
    o8[0] = cx - a * h - b * w;  o8[1] = cy + b * h - a * w;
    o8[2] = cx + a * h - b * w;  o8[3] = cy - b * h - a * w;
    o8[4] = 2 * cx - o8[0];      o8[5] = 2 * cy - o8[1];
    o8[6] = 2 * cx - o8[2];      o8[7] = 2 * cy - o8[3];
}

/* ---- O'Rourke point-in-polygon on integer vertices (skimage pnpoly) */
static int pip(const long* xp, const long* yp, int n, long x, long y) {
    int rc = 0, lc = 0;
    long x1 = xp[n - 1] - x, y1 = yp[n - 1] - y;
    for (int i = 0; i < n; ++i) {
        long x0 = xp[i] - x, y0 = yp[i] - y;
        if (x0 == 0 && y0 == 0) return 2;
        if ((y0 > 0) != (y1 > 0)) {
            double t = (double)(x0 * y1 - x1 * y0) / (double)(y1 - y0);
            if (t > 0) ++rc;
        }
        if ((y0 < 0) != (y1 < 0)) {
            double t = (double)(x0 * y1 - x1 * y0) / (double)(y1 - y0);
            if (t < 0) ++lc;
        }
        x1 = x0; y1 = y0;
    }
    if ((rc & 1) != (lc & 1)) return 3;
    return rc & 1;
}

/* paint one rectangle (+1) on canvas[480][640] exactly as calculate_iou does */
static void paint(const double* rect5, unsigned char* canvas) {
    float pts[8];
    oracle_box_points((float)rect5[0], (float)rect5[1], (float)rect5[2], (float)rect5[3], (float)(-rect5[4]), pts);
    long bx[4], by[4];
    for (int i = 0; i < 4; ++i) { bx[i] = (long)pts[2 * i]; by[i] = (long)pts[2 * i + 1]; } /* np.int0 */
    /* polygon(r = bx, c = by, shape=(480,640)) */
    long minr = bx[0], maxr = bx[0], minc = by[0], maxc = by[0];
    for (int i = 1; i < 4; ++i) {
        if (bx[i] < minr) minr = bx[i]; if (bx[i] > maxr) maxr = bx[i];
        if (by[i] < minc) minc = by[i]; if (by[i] > maxc) maxc = by[i];
    }
    if (minr < 0) minr = 0; if (minc < 0) minc = 0;
    if (maxr > CANVAS_H - 1) maxr = CANVAS_H - 1;
    if (maxc > CANVAS_W - 1) maxc = CANVAS_W - 1;
    for (long r = minr; r <= maxr; ++r)
        for (long c = minc; c <= maxc; ++c)
            if (pip(by, bx, 4, c, r)) {
                /* rr = r (an x coordinate), cc = c (a y coordinate); keep rr<640, cc<480 */
                if (r < CANVAS_W && c < CANVAS_H) canvas[c * CANVAS_W + r] += 1;
            }
}

/* (intersection, union) pixel counts; returns 0 if angle-gated (counts left 0) */
int oracle_iou_counts(const double* rect_p, const double* rect_gt, long* inter, long* uni) {
    *inter = 0; *uni = 0;
    if (fabs(rect_p[4] - rect_gt[4]) > 30 && fabs(rect_p[4] + rect_gt[4]) > 30) return 0;
    unsigned char* canvas = (unsigned char*)calloc((size_t)CANVAS_H * CANVAS_W, 1);
    paint(rect_gt, canvas);
    paint(rect_p, canvas);
    for (size_t i = 0; i < (size_t)CANVAS_H * CANVAS_W; ++i) {
        if (canvas[i] > 0) ++*uni;
        if (canvas[i] == 2) ++*inter;
    }
    free(canvas);
    return 1;
}

/* calculate_jacquard_index (grasp_eval.py:362-374): edits gts in place (stride 6) */
int oracle_jacquard(const double* preds, int K, double* gts, int M) {
    for (int m = 0; m < M; ++m) {
        gts[6 * m + 3] = 20.0;
        double w = gts[6 * m + 2];
        gts[6 * m + 2] = w < 0 ? 0 : (w > 100 ? 100 : w);
    }
    double best = 0;
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k) {
            long in, un;
            oracle_iou_counts(preds + 5 * k, gts + 6 * m, &in, &un);
            double v = un <= 0 ? 0.0 : (double)in / (double)un;
            if (v > best) best = v;
        }
    return best > 0.25 ? 1 : 0;
}

/* Serial evaluation loop of engine/crog_engine.py:478-527 for a batch of maps:
 * for K in (1, 5): detect_grasps + calculate_jacquard_index.  counters = {correct@1,total@1,correct@5,total@5}.
 * out_grasps: [B,5,5] doubles (top-5 rows), out_n: [B], out_j: [B,2]. */
void oracle_tail_batch(const float* q, const float* s, const float* c, const float* w, int B, int H, int W,
                       double* gts /*[B,Mmax,6]*/, const int* gt_cnt, int Mmax,
                       double* out_grasps, int* out_n, int* out_j, long* counters) {
    size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
        int rc[10]; double g1[5]; int rc1[2];
        int n1 = oracle_detect_grasps(q + b * plane, s + b * plane, c + b * plane, w + b * plane, H, W, 1, g1, rc1);
        int j1 = oracle_jacquard(g1, n1, gts + (size_t)b * Mmax * 6, gt_cnt[b]);
        double* g5 = out_grasps + (size_t)b * 25;
        memset(g5, 0, sizeof(double) * 25);
        int n5 = oracle_detect_grasps(q + b * plane, s + b * plane, c + b * plane, w + b * plane, H, W, 5, g5, rc);
        int j5 = oracle_jacquard(g5, n5, gts + (size_t)b * Mmax * 6, gt_cnt[b]);
        out_n[b] = n5; out_j[2 * b] = j1; out_j[2 * b + 1] = j5;
        counters[0] += j1; counters[1] += 1; counters[2] += j5; counters[3] += 1;
    }
}
