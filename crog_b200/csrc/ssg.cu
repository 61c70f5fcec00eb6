// SSG (config 4) kernels: the layout / gather kernels the torchvision-style ResNet-50 + FPN needs on top of the
// implicit-GEMM (strided convolutions as gathers, 3x3/2 max-pool), the head finalisation, and the device side of
// ssg_post_processing (utils/grasp_eval.py:100-221 of the reference): score filter + box decode, Fast NMS
// (:55-93), prototype assembly + crop (utils/box_utils.py:150-171), bilinear resize to the original image,
// and the 17-tap Gaussian with float64 accumulation (skimage.filters.gaussian -> scipy.ndimage).
#include <math.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ long long prow(int b, int y, int x, int H, int W, int padded) {
  return padded ? ((long long)(b * (H + 2) + y + 1) * (W + 2) + x + 1) : ((long long)(b * H + y) * W + x);
}

// ------------------------------------------------------------------ stem: 7x7 / stride 2 / pad 3 patches -> [rows, Kp]
// K index = (ky*7 + kx)*cin + c (matches the [Cout, ky, kx, Cin] weight packing), zero padded to Kp.
// One thread per (output pixel, 8-wide K group), K fastest so every warp store is a contiguous run of 16-byte vectors.
// Tiled form (compile-time channel count): a CTA stages the 7 input rows x (2 * 64 + 5) columns x CIN planes behind 64
// consecutive output pixels of one output row in shared memory (coalesced plane reads), then writes the patch rows with
// 16-byte stores.  The index arithmetic is constant divisions only and every input value is read from HBM / L2 once per
// CTA instead of ~12 times through L1 (the generic kernel below: 3.8 ms for 64 x 544 x 544, a quarter of the SSG forward).
constexpr int S7_PX = 64, S7_SW = 2 * S7_PX + 5;
template <typename T, int CIN>
__global__ void __launch_bounds__(256) stem7_patches_tiled_kernel(const float* __restrict__ rgb, const float* __restrict__ depth, int B,
                                                                  int H, int W, int Kp, T* __restrict__ out) {
  __shared__ float s_in[7 * CIN + 1][S7_SW + 1];  // last row: zeros (the padding columns of a patch row read it)
  const int OH = (H + 6 - 7) / 2 + 1, OW = (W + 6 - 7) / 2 + 1;
  const int segs = (OW + S7_PX - 1) / S7_PX;
  const int seg = blockIdx.x % segs, oy = (blockIdx.x / segs) % OH, b = blockIdx.x / (segs * OH);
  const int ox0 = seg * S7_PX, ix0 = 2 * ox0 - 3, iy0 = 2 * oy - 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < 7 * CIN; r += 8) {  // a warp stages whole rows: no index divisions, coalesced plane reads
    const int c = r % CIN, ky = r / CIN, iy = iy0 + ky;
    const bool rowok = iy >= 0 && iy < H;
    const float* src = c < 3 ? rgb + ((long long)(b * 3 + c) * H + iy) * W : depth + ((long long)b * H + iy) * W;
    for (int col = lane; col < S7_SW; col += 32) {
      const int ix = ix0 + col;
      s_in[r][col] = (rowok && ix >= 0 && ix < W) ? __ldg(src + ix) : 0.f;
    }
  }
  for (int col = threadIdx.x; col < S7_SW + 1; col += 256) s_in[7 * CIN][col] = 0.f;
  __syncthreads();
  // thread = one 8-wide k group (its eight source offsets are computed once) x every 8th pixel of the segment
  const int groups = Kp / 8;
  T* orow = out + ((long long)(b * OH + oy) * OW + ox0) * Kp;
  const int npx = min(S7_PX, OW - ox0);
  for (int gk = lane; gk < groups; gk += 32) {
    int off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = gk * 8 + j;
      const int c = k % CIN, t = k / CIN, ky = t / 7, kx = t % 7;
      off[j] = k < 49 * CIN ? (ky * CIN + c) * (S7_SW + 1) + kx : 7 * CIN * (S7_SW + 1);
    }
    const float* base = &s_in[0][0];
    for (int p = warp; p < npx; p += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = base[off[j] + 2 * p];
      store8(orow + (long long)p * Kp + gk * 8, v);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) stem7_patches_kernel(const float* __restrict__ rgb, const float* __restrict__ depth, int B,
                                                            int H, int W, int cin, int Kp, T* __restrict__ out) {
  const int OH = (H + 6 - 7) / 2 + 1, OW = (W + 6 - 7) / 2 + 1;
  const int groups = Kp / 8;
  const long long total = (long long)B * OH * OW * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int gk = (int)(i % groups);
    const long long pix = i / groups;
    const int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH), b = (int)(pix / ((long long)OW * OH));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = gk * 8 + j;
      float val = 0.f;
      if (k < 49 * cin) {
        const int c = k % cin, t = k / cin, ky = t / 7, kx = t % 7;
        const int iy = 2 * oy + ky - 3, ix = 2 * ox + kx - 3;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W)
          val = c < 3 ? __ldg(rgb + ((long long)(b * 3 + c) * H + iy) * W + ix) : __ldg(depth + ((long long)b * H + iy) * W + ix);
      }
      v[j] = val;
    }
    store8(out + pix * Kp + gk * 8, v);
  }
}

// ------------------------------------------------------------------ nn.MaxPool2d(3, 2, 1) on NHWC (ssg.py:67,101)
template <typename T>
__global__ void __launch_bounds__(256) maxpool3s2_kernel(const T* __restrict__ in, int in_ld, int in_padded, T* __restrict__ out,
                                                         int out_ld, int out_padded, int B, int H, int W, int C) {
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1, cgs = C / 8;
  const int b = blockIdx.x / OH, oy = blockIdx.x % OH;
  T* op = out + prow(b, oy, 0, OH, OW, out_padded) * out_ld;
  for (int i = threadIdx.x; i < OW * cgs; i += 256) {
    const int ox = i / cgs, cg = i - ox * cgs;
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * oy + ky - 1;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = 2 * ox + kx - 1;
        if (ix < 0 || ix >= W) continue;
        float v[8];
        load8(in + prow(b, iy, ix, H, W, in_padded) * in_ld + cg * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
      }
    }
    store8(op + (long long)ox * out_ld + cg * 8, m);
  }
}

// ------------------------------------------------------------------ 3x3 / stride s / pad 1 patches of a zero-haloed NHWC tensor
// out[(b, oy, ox), tap*C + c] = in_padded[b, s*oy + ky, s*ox + kx, c]   (padded coordinates, so pad = 1 is the halo)
template <typename T>
__global__ void __launch_bounds__(256) patches3_kernel(const T* __restrict__ in, int in_ld, T* __restrict__ out, int B, int H, int W,
                                                       int C, int stride) {
  const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1, cgs = C / 8;
  const int b = blockIdx.x / OH, oy = blockIdx.x % OH;
  const int n = OW * 9 * cgs;
  T* op = out + (long long)(b * OH + oy) * OW * 9 * C;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int cg = i % cgs, t = (i / cgs) % 9, ox = i / (9 * cgs);
    const int ky = t / 3, kx = t % 3;
    const long long r = (long long)(b * (H + 2) + stride * oy + ky) * (W + 2) + stride * ox + kx;
    float f[8];
    load8(in + r * in_ld + cg * 8, f);
    store8(op + ((long long)ox * 9 + t) * C + cg * 8, f);
  }
}

// ------------------------------------------------------------------ extra resample modes: x[::2, ::2] and bilinear x2 align_corners=True
template <typename T, int MODE>  // 3 = subsample2, 4 = bilinear x2 (align_corners=True, ssg.py:159)
__global__ void __launch_bounds__(256) resample2_kernel(const T* __restrict__ in, int in_ld, int in_padded, T* __restrict__ out,
                                                        int out_ld, int out_padded, int B, int H, int W, int C) {
  const int OH = MODE == 3 ? (H - 1) / 2 + 1 : 2 * H, OW = MODE == 3 ? (W - 1) / 2 + 1 : 2 * W, cgs = C / 8;
  const int b = blockIdx.x / OH, oy = blockIdx.x % OH;
  T* op = out + prow(b, oy, 0, OH, OW, out_padded) * out_ld;
  if (MODE == 3) {
    for (int i = threadIdx.x; i < OW * cgs; i += 256) {
      const int ox = i / cgs, cg = i - ox * cgs;
      float v[8];
      load8(in + prow(b, 2 * oy, 2 * ox, H, W, in_padded) * in_ld + cg * 8, v);
      store8(op + (long long)ox * out_ld + cg * 8, v);
    }
  } else {
    // PyTorch: scale = (in - 1) / (out - 1); src = scale * dst; lambda1 = src - floor(src)
    const float sh = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f, sw = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    const float sy = sh * oy;
    const int y0 = (int)sy, y1 = min(y0 + 1, H - 1);
    const float ly = sy - y0, hy = 1.f - ly;
    for (int i = threadIdx.x; i < OW * cgs; i += 256) {
      const int ox = i / cgs, cg = i - ox * cgs;
      const float sx = sw * ox;
      const int x0 = (int)sx, x1 = min(x0 + 1, W - 1);
      const float lx = sx - x0, hx = 1.f - lx;
      float v[8], a[8], c[8], d[8];
      load8(in + prow(b, y0, x0, H, W, in_padded) * in_ld + cg * 8, v);
      load8(in + prow(b, y0, x1, H, W, in_padded) * in_ld + cg * 8, a);
      load8(in + prow(b, y1, x0, H, W, in_padded) * in_ld + cg * 8, c);
      load8(in + prow(b, y1, x1, H, W, in_padded) * in_ld + cg * 8, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = hy * (hx * v[j] + lx * a[j]) + ly * (hx * c[j] + lx * d[j]);
      store8(op + (long long)ox * out_ld + cg * 8, v);
    }
  }
}

// ------------------------------------------------------------------ head finalisation: softmax over classes + box slice
// in: [rows, ld] fp32 with columns [na*nc class logits | na*4 box deltas | pad]; one warp per (row, anchor).
__global__ void __launch_bounds__(256) ssg_heads_kernel(const float* __restrict__ in, int ld, long long rows, int na, int nc,
                                                        float* __restrict__ cls, float* __restrict__ box) {
  const int lane = threadIdx.x & 31;
  const long long w = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= rows * na) return;
  const long long r = w / na;
  const int a = (int)(w % na);
  const float* src = in + r * ld + a * nc;
  float mx = -INFINITY;
  for (int k = lane; k < nc; k += 32) mx = fmaxf(mx, src[k]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int k = lane; k < nc; k += 32) s += expf(src[k] - mx);
  s = warp_sum(s);
  for (int k = lane; k < nc; k += 32) cls[w * nc + k] = expf(src[k] - mx) / s;
  if (lane < 4) box[w * 4 + lane] = in[r * ld + na * nc + a * 4 + lane];
}

// ------------------------------------------------------------------ post-processing: score filter + box decode
// grasp_eval.py:113-137: keep = max_{c>=1} cls[n, c] > thr; centre-form decode with variances (0.1, 0.2) -> point form,
// clipped to [0, 1].  Every float op is a separately rounded fp32 op in the reference's order.
__global__ void ssg_decode_kernel(const float* __restrict__ cls, const float* __restrict__ box, const float* __restrict__ anchors,
                                  int N, int nc, float thr, int* __restrict__ keep, float* __restrict__ boxes) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  {  // blockIdx.y = image of a batch-wide launch ([B,N,nc] scores, [B,N,4] offsets; the anchors are shared)
    const long long img = blockIdx.y;
    cls += img * N * nc; box += img * N * 4; keep += img * N; boxes += img * N * 4;
  }
  float mx = -INFINITY;
  for (int c = 1; c < nc; ++c) mx = fmaxf(mx, cls[(long long)n * nc + c]);
  keep[n] = mx > thr;
  const float ax = anchors[n * 4], ay = anchors[n * 4 + 1], aw = anchors[n * 4 + 2], ah = anchors[n * 4 + 3];
  const float bx = box[n * 4], by = box[n * 4 + 1], bw = box[n * 4 + 2], bh = box[n * 4 + 3];
  const float cx = __fadd_rn(ax, __fmul_rn(__fmul_rn(bx, 0.1f), aw)), cy = __fadd_rn(ay, __fmul_rn(__fmul_rn(by, 0.1f), ah));
  const float w = __fmul_rn(aw, expf(__fmul_rn(bw, 0.2f))), h = __fmul_rn(ah, expf(__fmul_rn(bh, 0.2f)));
  const float x1 = __fsub_rn(cx, __fdiv_rn(w, 2.f)), y1 = __fsub_rn(cy, __fdiv_rn(h, 2.f));
  const float x2 = __fadd_rn(w, x1), y2 = __fadd_rn(h, y1);
  boxes[n * 4] = fminf(fmaxf(x1, 0.f), 1.f); boxes[n * 4 + 1] = fminf(fmaxf(y1, 0.f), 1.f);
  boxes[n * 4 + 2] = fminf(fmaxf(x2, 0.f), 1.f); boxes[n * 4 + 3] = fminf(fmaxf(y2, 0.f), 1.f);
}

// ------------------------------------------------------------------ Fast NMS (grasp_eval.py:55-93)
constexpr int NMS_T = 1024;       // threads per class CTA
constexpr int NMS_CAP = 24576;    // kept anchors held in shared memory (192 KB of 64-bit keys)
constexpr int NMS_SEL = 256;      // >= top_k

__device__ __forceinline__ unsigned long long score_key(float s, int idx) {
  // scores are softmax outputs (>= 0): the IEEE bit pattern orders them; ties -> lower index first (stable sort)
  return ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)idx);
}

// in-place bitonic sort, descending, n a power of two, all threads of the CTA participate
__device__ void bitonic_desc(unsigned long long* a, int n) {
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long x = a[i], y = a[l];
          const bool up = (i & k) == 0;  // descending runs first
          if (up ? (x < y) : (x > y)) { a[i] = y; a[l] = x; }
        }
      }
    }
  __syncthreads();
}

__device__ __forceinline__ float box_iou1(const float4 a, const float4 b) {
  // utils/box_utils.py:28-36, separately rounded fp32 ops
  const float iw = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f), ih = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
  const float ia = __fmul_rn(iw, ih);
  const float aa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y)), ab = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return __fdiv_rn(ia, __fsub_rn(__fadd_rn(aa, ab), ia));
}

// One CTA per foreground class: exact top_k of the kept anchors by (score desc, index asc) with a 64-bit radix select in
// shared memory, a bitonic sort of the selected keys, then the upper-triangular IoU column maximum.
__global__ void __launch_bounds__(NMS_T) ssg_nms_class_kernel(const float* __restrict__ cls, const int* __restrict__ keep,
                                                              const float* __restrict__ boxes, int N, int nc, int top_k, float iou_thr,
                                                              unsigned long long* __restrict__ cand, int* __restrict__ cand_keep,
                                                              int* __restrict__ cand_n, long long ws_stride) {
  {  // blockIdx.y = image of a batch-wide launch; every image has its own workspace slice (ws_stride bytes apart)
    const long long img = blockIdx.y;
    cls += img * N * nc; keep += img * N; boxes += img * N * 4;
    cand = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(cand) + img * ws_stride);
    cand_keep = reinterpret_cast<int*>(reinterpret_cast<char*>(cand_keep) + img * ws_stride);
    cand_n = reinterpret_cast<int*>(reinterpret_cast<char*>(cand_n) + img * ws_stride);
  }
  extern __shared__ unsigned long long s_keys[];  // [NMS_CAP]
  __shared__ unsigned long long s_sel[NMS_SEL];
  __shared__ float4 s_box[NMS_SEL];
  __shared__ int s_hist[256];
  __shared__ int s_cnt, s_nsel;
  __shared__ unsigned long long s_prefix;
  __shared__ int s_remaining;
  const int c = blockIdx.x + 1;  // class (0 = background is excluded)
  if (threadIdx.x == 0) { s_cnt = 0; s_nsel = 0; }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += NMS_T)
    if (keep[n]) {
      const int p = atomicAdd(&s_cnt, 1);
      if (p < NMS_CAP) s_keys[p] = score_key(cls[(long long)n * nc + c], n);
    }
  __syncthreads();
  const int n = min(s_cnt, NMS_CAP);
  unsigned long long T = 0ull;  // keys >= T are selected
  if (n > top_k) {
    if (threadIdx.x == 0) { s_prefix = 0ull; s_remaining = top_k; }
    for (int pass = 7; pass >= 0; --pass) {
      for (int i = threadIdx.x; i < 256; i += NMS_T) s_hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      const unsigned long long himask = pass == 7 ? 0ull : (~0ull << (8 * (pass + 1)));
      for (int i = threadIdx.x; i < n; i += NMS_T) {
        const unsigned long long k = s_keys[i];
        if ((k & himask) == prefix) atomicAdd(&s_hist[(int)((k >> (8 * pass)) & 0xff)], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int rem = s_remaining, b = 255;
        for (; b > 0; --b) {
          if (s_hist[b] >= rem) break;
          rem -= s_hist[b];
        }
        s_remaining = rem;  // rank of the wanted key inside bin b
        s_prefix = prefix | ((unsigned long long)b << (8 * pass));
      }
      __syncthreads();
    }
    T = s_prefix;
  }
  for (int i = threadIdx.x; i < NMS_SEL; i += NMS_T) s_sel[i] = 0ull;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += NMS_T) {
    const unsigned long long k = s_keys[i];
    if (k >= T && k != 0ull) { const int p = atomicAdd(&s_nsel, 1); if (p < NMS_SEL) s_sel[p] = k; }
  }
  __syncthreads();
  const int nsel = min(s_nsel, top_k);
  bitonic_desc(s_sel, NMS_SEL);
  for (int i = threadIdx.x; i < nsel; i += NMS_T) {
    const int a = (int)(0xffffffffu - (uint32_t)(s_sel[i] & 0xffffffffull));
    s_box[i] = *reinterpret_cast<const float4*>(boxes + (long long)a * 4);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nsel; j += NMS_T) {
    float m = 0.f;  // triu_(diagonal=1) leaves zeros on and below the diagonal, so the column maximum starts at 0
    bool isnan_ = false;
    const float4 bj = s_box[j];
    for (int i = 0; i < j; ++i) {
      const float v = box_iou1(s_box[i], bj);
      if (v != v) isnan_ = true;  // torch.max propagates NaN (0/0 for degenerate boxes); NaN <= thr is False
      m = fmaxf(m, v);
    }
    cand[(long long)blockIdx.x * top_k + j] = s_sel[j];
    cand_keep[(long long)blockIdx.x * top_k + j] = (!isnan_ && m <= iou_thr) ? 1 : 0;
  }
  if (threadIdx.x == 0) cand_n[blockIdx.x] = nsel;
}

// One CTA: flatten the kept candidates class-major (the order boolean indexing gives), stable sort by score, keep
// max_det, then the score > thr2 filter of grasp_eval.py:142-150 (kept only if at least one detection passes).
constexpr int MRG_CAP = 8192;
__global__ void __launch_bounds__(1024) ssg_nms_merge_kernel(const unsigned long long* __restrict__ cand, const int* __restrict__ cand_keep,
                                                             const int* __restrict__ cand_n, int ncls, int top_k, int max_det, float thr2,
                                                             int* __restrict__ det_n, int* __restrict__ det_anchor, int* __restrict__ det_class,
                                                             float* __restrict__ det_score, long long ws_stride) {
  {  // blockIdx.x = image of a batch-wide launch
    const long long img = blockIdx.x;
    cand = reinterpret_cast<const unsigned long long*>(reinterpret_cast<const char*>(cand) + img * ws_stride);
    cand_keep = reinterpret_cast<const int*>(reinterpret_cast<const char*>(cand_keep) + img * ws_stride);
    cand_n = reinterpret_cast<const int*>(reinterpret_cast<const char*>(cand_n) + img * ws_stride);
    det_n += img; det_anchor += img * max_det; det_class += img * max_det; det_score += img * max_det;
  }
  extern __shared__ unsigned long long s_k[];  // [MRG_CAP]
  __shared__ int s_pass;
  for (int i = threadIdx.x; i < MRG_CAP; i += blockDim.x) {
    unsigned long long k = 0ull;
    if (i < ncls * top_k) {
      const int c = i / top_k, j = i % top_k;
      if (j < cand_n[c] && cand_keep[i]) k = (cand[i] & 0xffffffff00000000ull) | (unsigned long long)(0xffffffffu - (uint32_t)i);
    }
    s_k[i] = k;
  }
  if (threadIdx.x == 0) s_pass = 0;
  bitonic_desc(s_k, MRG_CAP);
  int total = 0;
  for (int i = threadIdx.x; i < max_det; i += blockDim.x)
    if (s_k[i] != 0ull && __uint_as_float((uint32_t)(s_k[i] >> 32)) > thr2) atomicAdd(&s_pass, 1);
  __syncthreads();
  // count of non-empty entries among the first max_det
  __shared__ int s_tot;
  if (threadIdx.x == 0) s_tot = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < max_det; i += blockDim.x) if (s_k[i] != 0ull) atomicAdd(&s_tot, 1);
  __syncthreads();
  total = s_pass > 0 ? s_pass : s_tot;  // sorted descending, so the passing detections are a prefix
  for (int i = threadIdx.x; i < max_det; i += blockDim.x) {
    if (i < total) {
      const int flat = (int)(0xffffffffu - (uint32_t)(s_k[i] & 0xffffffffull));
      det_anchor[i] = (int)(0xffffffffu - (uint32_t)(cand[flat] & 0xffffffffull));
      det_class[i] = flat / top_k;  // 0-based foreground class (the reference adds 1 afterwards)
      det_score[i] = __uint_as_float((uint32_t)(s_k[i] >> 32));
    } else {
      det_anchor[i] = -1; det_class[i] = -1; det_score[i] = 0.f;
    }
  }
  if (threadIdx.x == 0) *det_n = total;
}

// ------------------------------------------------------------------ prototype assembly + crop (grasp_eval.py:171-181)
// out[d][k][y][x] = crop(act_k(protos[y, x, :] . coef_k(d)))  with k = 0 ins, 1 qua, 2 sin, 3 cos, 4 wid; sigmoid on 0, 1, 4.
__global__ void __launch_bounds__(256) ssg_lowres_kernel(const float* __restrict__ protos, int h, int w, int np_,
                                                         const float* __restrict__ coef, const float* __restrict__ gcoef,
                                                         const float* __restrict__ boxes, const int* __restrict__ det_anchor,
                                                         const int* __restrict__ det_n, float* __restrict__ out,
                                                         const int* __restrict__ inst_image, const int* __restrict__ inst_det, int N,
                                                         int max_det) {
  extern __shared__ float s_c[];  // [5][np]
  const int d = blockIdx.y;
  int a;
  if (inst_image != nullptr) {  // batch-wide call: instance d of the batch = detection inst_det[d] of image inst_image[d]
    const int b = inst_image[d];
    a = det_anchor[b * max_det + inst_det[d]];
    protos += (long long)b * h * w * np_;
    coef += (long long)b * N * np_;
    gcoef += (long long)b * N * 4 * np_;
    boxes += (long long)b * N * 4;
  } else {
    if (d >= *det_n) return;
    a = det_anchor[d];
  }
  for (int i = threadIdx.x; i < 5 * np_; i += blockDim.x) {
    const int k = i / np_, j = i % np_;
    s_c[i] = k == 0 ? coef[(long long)a * np_ + j] : gcoef[((long long)a * 4 + (k - 1)) * np_ + j];
  }
  __syncthreads();
  // sanitize_coordinates (box_utils.py:120-135) with padding = 1
  const float bx1 = boxes[a * 4], by1 = boxes[a * 4 + 1], bx2 = boxes[a * 4 + 2], by2 = boxes[a * 4 + 3];
  const float xa = __fmul_rn(bx1, (float)w), xb = __fmul_rn(bx2, (float)w), ya = __fmul_rn(by1, (float)h), yb = __fmul_rn(by2, (float)h);
  const float x1 = fmaxf(__fsub_rn(fminf(xa, xb), 1.f), 0.f), x2 = fminf(__fadd_rn(fmaxf(xa, xb), 1.f), (float)w);
  const float y1 = fmaxf(__fsub_rn(fminf(ya, yb), 1.f), 0.f), y2 = fminf(__fadd_rn(fmaxf(ya, yb), 1.f), (float)h);
  const int npix = h * w;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    const int y = p / w, x = p % w;
    const bool inside = (float)x >= x1 && (float)x < x2 && (float)y >= y1 && (float)y < y2;
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (inside) {
      const float* pr = protos + (long long)p * np_;
      for (int j = 0; j < np_; j += 4) {
        const float4 q = *reinterpret_cast<const float4*>(pr + j);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          acc[k] = fmaf(q.x, s_c[k * np_ + j], acc[k]); acc[k] = fmaf(q.y, s_c[k * np_ + j + 1], acc[k]);
          acc[k] = fmaf(q.z, s_c[k * np_ + j + 2], acc[k]); acc[k] = fmaf(q.w, s_c[k * np_ + j + 3], acc[k]);
        }
      }
      acc[0] = 1.f / (1.f + expf(-acc[0])); acc[1] = 1.f / (1.f + expf(-acc[1])); acc[4] = 1.f / (1.f + expf(-acc[4]));
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) out[((long long)d * 5 + k) * npix + p] = acc[k];
  }
}

// ------------------------------------------------------------------ F.interpolate(size=(S,S), bilinear, align_corners=False) + [:oh, :ow] crop
// planes [P][h][w] -> [P][oh][ow]; planes whose bit is set in bin_mask (by plane % planes_per_det) are thresholded > 0.5.
constexpr int BC_ROWS = 32;  // output rows per thread
constexpr int BC_WARPS = 8;  // row groups (warps) per CTA
// A warp owns a 128-column x BC_ROWS-row block of one plane; a thread 4 consecutive output columns (one 16-byte store
// per row).  The horizontal source indices / weights are computed once per column, the vertical ones once per warp
// (shared-memory table), and the two horizontally interpolated source rows slide down the columns: at scale 136 -> 640 a
// new source row (2 loads, 3 flops per column) is needed every ~4.7 output rows, so an output costs ~5 instructions
// instead of ~35 (4 loads + full address / weight arithmetic per output: 2.9 ms for the 3.1 GB of masks of a 64-image
// batch).
__global__ void __launch_bounds__(32 * BC_WARPS) bilinear_crop_kernel(const float* __restrict__ in, int h, int w, float* __restrict__ out,
                                                                      int oh, int ow, int S, const int* __restrict__ n_planes,
                                                                      int planes_per_det, uint32_t bin_mask, int det_stride,
                                                                      float* __restrict__ alt1) {
  __shared__ int s_y0[BC_WARPS][BC_ROWS];
  __shared__ float s_ly[BC_WARPS][BC_ROWS];
  const int pl = blockIdx.z;
  if (n_planes && pl >= *n_planes * planes_per_det) return;
  const int lane = threadIdx.x, rg = threadIdx.y;
  const float sc_h = (float)h / (float)S, sc_w = (float)w / (float)S;
  const int oyb = (blockIdx.y * BC_WARPS + rg) * BC_ROWS;
  if (oyb >= oh) return;  // warp-uniform
  {
    const float sy = fmaxf(__fsub_rn(__fmul_rn(sc_h, (float)(oyb + lane) + 0.5f), 0.5f), 0.f);
    const int y0 = min((int)sy, h - 1);
    s_y0[rg][lane] = y0;
    s_ly[rg][lane] = sy - y0;
  }
  __syncwarp();
  const int oxb = (blockIdx.x * 32 + lane) * 4;
  if (oxb >= ow) return;
  int x0[4], x1[4];
  float lx[4], hx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ox = min(oxb + i, ow - 1);  // columns past the edge (ow % 4 != 0) repeat the last one and are not stored
    const float sx = fmaxf(__fsub_rn(__fmul_rn(sc_w, (float)ox + 0.5f), 0.5f), 0.f);
    x0[i] = (int)sx; x1[i] = x0[i] + (x0[i] < w - 1);
    lx[i] = sx - x0[i]; hx[i] = 1.f - lx[i];
  }
  const float* src = in + (long long)pl * h * w;
  const int d = pl / planes_per_det, k = pl % planes_per_det;
  const bool bin = (bin_mask >> k) & 1u;
  // output is map-major: [planes_per_det][det_stride detections][oh][ow], so each map type is one contiguous batch
  float* dst = out + ((long long)k * det_stride + d) * oh * ow + oxb;
  if (alt1 != nullptr && k == 1) dst = alt1 + (long long)d * oh * ow + oxb;  // plane 1 (raw quality) to its own [det][oh][ow] buffer
  const bool vec = oxb + 4 <= ow && (ow & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
  int cur = -2;
  float top[4] = {0.f, 0.f, 0.f, 0.f}, bot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int j = 0; j < BC_ROWS; ++j) {
    const int oy = oyb + j;
    if (oy >= oh) break;
    const int y0 = s_y0[rg][j];  // warp-uniform
    if (y0 != cur) {
      const bool reuse = y0 == cur + 1 && cur >= 0, last = y0 >= h - 1;
      const float* r0 = src + y0 * w;
      const float* r1 = r0 + (last ? 0 : w);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        top[i] = reuse ? bot[i] : hx[i] * __ldg(r0 + x0[i]) + lx[i] * __ldg(r0 + x1[i]);
        bot[i] = last ? top[i] : hx[i] * __ldg(r1 + x0[i]) + lx[i] * __ldg(r1 + x1[i]);
      }
      cur = y0;
    }
    const float ly = s_ly[rg][j], hy = 1.f - ly;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = hy * top[i] + ly * bot[i];
      if (bin) v[i] = v[i] > 0.5f ? 1.f : 0.f;
    }
    float* o = dst + (long long)oy * ow;
    if (vec) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    else {
#pragma unroll
      for (int i = 0; i < 4; ++i) if (oxb + i < ow) o[i] = v[i];
    }
  }
}

// ------------------------------------------------------------------ Gaussian (sigma = 2 -> 17 taps), float64 accumulation, float32 per pass
struct GaussW { double w[33]; int r; };

// scipy correlate1d order for a symmetric kernel: centre tap first, then pairs from the outermost inwards; each product and
// sum is a separately rounded float64 operation (no FMA), the pass result is rounded to float32 once.
template <int AXIS>
__global__ void __launch_bounds__(256) gaussian_pass_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W,
                                                            const int* __restrict__ n_planes, int plane_stride_sel, int plane_sel, GaussW g) {
  // planes processed: pl = blockIdx.z * plane_stride_sel + plane_sel (lets the caller smooth only the quality planes)
  const int pz = blockIdx.z;
  if (n_planes && pz >= *n_planes) return;
  const long long base = ((long long)pz * plane_stride_sel + plane_sel) * H * W;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const float* src = in + base;
  const int r = g.r;
  auto at = [&](int o) -> double {
    if (AXIS == 0) { const int yy = min(max(y + o, 0), H - 1); return (double)__ldg(src + (long long)yy * W + x); }
    const int xx = min(max(x + o, 0), W - 1);
    return (double)__ldg(src + (long long)y * W + xx);
  };
  double acc = __dmul_rn(at(0), g.w[r]);
  for (int j = -r; j < 0; ++j) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(at(j), at(-j)), g.w[j + r]));
  out[base + (long long)y * W + x] = (float)acc;
}

// Fast forms for a compile-time radius (sigma = 2 -> R = 8): the same per-output operation order as the generic kernel above
// (so the results are bit-identical), but an output no longer costs 2R + 1 global loads.
//  * vertical pass: a thread owns one column and GV_TY consecutive rows; the GV_TY + 2R clamped source values sit in
//    registers and every output reads its window from them (3 loads per output instead of 17, all coalesced);
//  * horizontal pass: a CTA stages GH_ROWS row segments (+R on either side, edge-replicated) in shared memory with
//    coalesced loads; a thread then reads its 2R + 1 neighbours from shared memory (consecutive lanes, no conflicts).
constexpr int GV_TY = 8, GH_ROWS = 4;
template <int R>
__global__ void __launch_bounds__(256) gaussian_vert_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W,
                                                            const int* __restrict__ n_planes, int plane_stride_sel, int plane_sel, GaussW g) {
  const int pz = blockIdx.z;
  if (n_planes && pz >= *n_planes) return;
  const long long base = ((long long)pz * plane_stride_sel + plane_sel) * H * W;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y0 = blockIdx.y * GV_TY;
  if (x >= W) return;
  const float* src = in + base + x;
  double v[GV_TY + 2 * R];
#pragma unroll
  for (int j = 0; j < GV_TY + 2 * R; ++j) v[j] = (double)__ldg(src + (long long)min(max(y0 - R + j, 0), H - 1) * W);
#pragma unroll
  for (int t = 0; t < GV_TY; ++t) {
    if (y0 + t >= H) break;
    double acc = __dmul_rn(v[t + R], g.w[R]);
#pragma unroll
    for (int j = -R; j < 0; ++j) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(v[t + R + j], v[t + R - j]), g.w[j + R]));
    out[base + (long long)(y0 + t) * W + x] = (float)acc;
  }
}
template <int R>
__global__ void __launch_bounds__(256) gaussian_horz_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W,
                                                            const int* __restrict__ n_planes, int plane_stride_sel, int plane_sel, GaussW g) {
  __shared__ float tile[GH_ROWS][256 + 2 * R];
  const int pz = blockIdx.z;
  if (n_planes && pz >= *n_planes) return;
  const long long base = ((long long)pz * plane_stride_sel + plane_sel) * H * W;
  const int x0 = blockIdx.x * 256, y0 = blockIdx.y * GH_ROWS;
  for (int i = threadIdx.x; i < GH_ROWS * (256 + 2 * R); i += 256) {
    const int rr = i / (256 + 2 * R), cc = i % (256 + 2 * R);
    const int yy = min(y0 + rr, H - 1), xx = min(max(x0 - R + cc, 0), W - 1);
    tile[rr][cc] = __ldg(in + base + (long long)yy * W + xx);
  }
  __syncthreads();
  const int x = x0 + threadIdx.x;
  if (x >= W) return;
#pragma unroll
  for (int rr = 0; rr < GH_ROWS; ++rr) {
    if (y0 + rr >= H) break;
    const float* t = &tile[rr][threadIdx.x + R];
    double acc = __dmul_rn((double)t[0], g.w[R]);
#pragma unroll
    for (int j = -R; j < 0; ++j) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)t[j], (double)t[-j]), g.w[j + R]));
    out[base + (long long)(y0 + rr) * W + x] = (float)acc;
  }
}

// Both passes in one kernel: a CTA stages a (GF_TH + 2R) x (GF_TW + 2R) clamped input tile in shared memory, runs the
// vertical pass for all GF_TW + 2R columns into a second tile (rounded to float32 exactly where the two-pass form stores
// its intermediate), then the horizontal pass from that tile.  Same per-output operation order as above, so the results
// are bit-identical; the float32 intermediate never goes to HBM (3.8 GB -> 1.9 GB per 768 maps of 480 x 640) and a
// thread converts 3 values per output to float64 instead of 17.  What is left is the float64 pipe: 50 separately
// rounded DADD / DMUL per output.
constexpr int GF_TW = 128, GF_TH = 32;
template <int R>
__global__ void __launch_bounds__(256) gaussian_fused_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W,
                                                             const int* __restrict__ n_planes, int plane_stride_sel, int plane_sel, GaussW g) {
  constexpr int TWH = GF_TW + 2 * R;
  __shared__ float t_in[GF_TH + 2 * R][TWH];
  __shared__ __align__(16) float t_mid[GF_TH][TWH];
  const int pz = blockIdx.z;
  if (n_planes && pz >= *n_planes) return;
  const long long base = ((long long)pz * plane_stride_sel + plane_sel) * H * W;
  const int x0 = blockIdx.x * GF_TW, y0 = blockIdx.y * GF_TH;
  for (int i = threadIdx.x; i < (GF_TH + 2 * R) * TWH; i += 256) {
    const int rr = i / TWH, cc = i - rr * TWH;
    const int yy = min(max(y0 - R + rr, 0), H - 1), xx = min(max(x0 - R + cc, 0), W - 1);
    t_in[rr][cc] = __ldg(in + base + (long long)yy * W + xx);
  }
  __syncthreads();
  // vertical pass: task = (column, group of 8 rows); consecutive threads take consecutive columns
  for (int task = threadIdx.x; task < TWH * (GF_TH / 8); task += 256) {
    const int gy = task / TWH, c = task - gy * TWH;
    double v[8 + 2 * R];
#pragma unroll
    for (int j = 0; j < 8 + 2 * R; ++j) v[j] = (double)t_in[gy * 8 + j][c];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      double acc = __dmul_rn(v[t + R], g.w[R]);
#pragma unroll
      for (int j = -R; j < 0; ++j) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(v[t + R + j], v[t + R - j]), g.w[j + R]));
      t_mid[gy * 8 + t][c] = (float)acc;
    }
  }
  __syncthreads();
  // horizontal pass: task = (row, group of 8 columns); a thread's 8 + 2R inputs are six 16-byte shared-memory loads
  for (int task = threadIdx.x; task < GF_TH * (GF_TW / 8); task += 256) {
    const int r = task / (GF_TW / 8), gx = task - r * (GF_TW / 8);
    const int y = y0 + r, xb = x0 + gx * 8;
    if (y >= H || xb >= W) continue;
    double v[8 + 2 * R];
    const float4* src = reinterpret_cast<const float4*>(&t_mid[r][gx * 8]);
#pragma unroll
    for (int j = 0; j < (8 + 2 * R) / 4; ++j) {
      const float4 q = src[j];
      v[4 * j] = (double)q.x; v[4 * j + 1] = (double)q.y; v[4 * j + 2] = (double)q.z; v[4 * j + 3] = (double)q.w;
    }
    float o[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      double acc = __dmul_rn(v[t + R], g.w[R]);
#pragma unroll
      for (int j = -R; j < 0; ++j) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(v[t + R + j], v[t + R - j]), g.w[j + R]));
      o[t] = (float)acc;
    }
    float* dst = out + base + (long long)y * W + xb;
    if (xb + 8 <= W && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) if (xb + t < W) dst[t] = o[t];
    }
  }
}

}  // namespace

// ====================================================================== C ABI
extern "C" int crog_stem7_patches(const float* rgb, const float* depth, int32_t B, int32_t H, int32_t W, int32_t cin, int32_t Kp,
                                  void* out, int32_t out_dtype, void* stream) {
  CROG_REQUIRE((cin == 3 || cin == 4) && Kp % 8 == 0 && Kp >= 49 * cin, CROG_E_BADSHAPE, "stem7_patches: cin %d Kp %d", cin, Kp);
  CROG_REQUIRE(cin == 3 || depth != nullptr, CROG_E_BADSHAPE, "stem7_patches: depth plane missing");
  const long long total = (long long)B * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1) * (Kp / 8);
  if (total == 0) return CROG_OK;
  long long g = (total + 255) / 256;
  if (g > 148 * 64) g = 148 * 64;
  {
    const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
    const long long ctas = (long long)B * OH * ((OW + S7_PX - 1) / S7_PX);
    if (ctas <= 0x7fffffffLL && !getenv("CROG_STEM7_GENERIC")) {
      cudaStream_t s = (cudaStream_t)stream;
#define S7(T, C) stem7_patches_tiled_kernel<T, C><<<(int)ctas, 256, 0, s>>>(rgb, depth, B, H, W, Kp, (T*)out)
      if (out_dtype == CROG_F32) { if (cin == 4) S7(float, 4); else S7(float, 3); }
      else { if (cin == 4) S7(bf16, 4); else S7(bf16, 3); }
#undef S7
      CROG_LAUNCH_OK("stem7_patches");
      return CROG_OK;
    }
  }
  if (out_dtype == CROG_F32) stem7_patches_kernel<float><<<(int)g, 256, 0, (cudaStream_t)stream>>>(rgb, depth, B, H, W, cin, Kp, (float*)out);
  else stem7_patches_kernel<bf16><<<(int)g, 256, 0, (cudaStream_t)stream>>>(rgb, depth, B, H, W, cin, Kp, (bf16*)out);
  CROG_LAUNCH_OK("stem7_patches");
  return CROG_OK;
}

extern "C" int crog_maxpool3s2(const void* in, int32_t in_ld, int32_t in_padded, void* out, int32_t out_ld, int32_t out_padded,
                               int32_t B, int32_t H, int32_t W, int32_t C, int32_t dtype, void* stream) {
  CROG_REQUIRE(C % 8 == 0 && in_ld % 8 == 0 && out_ld % 8 == 0 && aligned16(in) && aligned16(out), CROG_E_BADSHAPE, "maxpool3s2: bad shape");
  if (B * H * W == 0) return CROG_OK;
  const int g = B * ((H - 1) / 2 + 1);
  if (dtype == CROG_F32) maxpool3s2_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)in, in_ld, in_padded, (float*)out, out_ld, out_padded, B, H, W, C);
  else maxpool3s2_kernel<bf16><<<g, 256, 0, (cudaStream_t)stream>>>((const bf16*)in, in_ld, in_padded, (bf16*)out, out_ld, out_padded, B, H, W, C);
  CROG_LAUNCH_OK("maxpool3s2");
  return CROG_OK;
}

extern "C" int crog_patches3(const void* in, int32_t in_ld, void* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t stride,
                             int32_t dtype, void* stream) {
  CROG_REQUIRE(C % 8 == 0 && in_ld % 8 == 0 && (stride == 1 || stride == 2) && aligned16(in) && aligned16(out), CROG_E_BADSHAPE, "patches3: bad shape");
  if (B * H * W == 0) return CROG_OK;
  const int g = B * ((H - 1) / stride + 1);
  if (dtype == CROG_F32) patches3_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)in, in_ld, (float*)out, B, H, W, C, stride);
  else patches3_kernel<bf16><<<g, 256, 0, (cudaStream_t)stream>>>((const bf16*)in, in_ld, (bf16*)out, B, H, W, C, stride);
  CROG_LAUNCH_OK("patches3");
  return CROG_OK;
}

int crog_resample2(const void* in, int in_ld, int in_padded, void* out, int out_ld, int out_padded, int B, int H, int W, int C, int mode,
                   int dtype, cudaStream_t s) {
  const int OH = mode == 3 ? (H - 1) / 2 + 1 : 2 * H;
  const int g = B * OH;
  if (g == 0) return CROG_OK;
#define RS2(T, M) resample2_kernel<T, M><<<g, 256, 0, s>>>((const T*)in, in_ld, in_padded, (T*)out, out_ld, out_padded, B, H, W, C)
  if (dtype == CROG_F32) { if (mode == 3) RS2(float, 3); else RS2(float, 4); }
  else { if (mode == 3) RS2(bf16, 3); else RS2(bf16, 4); }
#undef RS2
  CROG_LAUNCH_OK("resample2");
  return CROG_OK;
}

extern "C" int crog_ssg_heads(const float* in, int32_t ld, int64_t rows, int32_t na, int32_t nc, float* cls, float* box, void* stream) {
  CROG_REQUIRE(ld >= na * (nc + 4), CROG_E_BADSHAPE, "ssg_heads: ld %d too small", ld);
  if (rows == 0) return CROG_OK;
  const long long warps = rows * na;
  ssg_heads_kernel<<<(int)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(in, ld, rows, na, nc, cls, box);
  CROG_LAUNCH_OK("ssg_heads");
  return CROG_OK;
}

extern "C" int64_t crog_ssg_nms_workspace_bytes(int32_t num_classes, int32_t top_k) {
  const int64_t n = (int64_t)(num_classes - 1) * top_k;
  return n * 8 + n * 4 + (int64_t)(num_classes - 1) * 4 + 256;
}

static int ssg_fast_nms_launch(const float* cls, const int32_t* keep, const float* boxes, int32_t B, int32_t N, int32_t num_classes,
                               float iou_thr, int32_t top_k, int32_t max_det, float score_thr2, int32_t* det_n, int32_t* det_anchor,
                               int32_t* det_class, float* det_score, void* workspace, long long ws_stride, void* stream) {
  CROG_REQUIRE(N >= 0 && num_classes >= 2 && top_k >= 1 && top_k <= NMS_SEL && max_det >= 1 && max_det <= 1024, CROG_E_BADSHAPE,
               "ssg_fast_nms: N %d classes %d top_k %d max_det %d", N, num_classes, top_k, max_det);
  CROG_REQUIRE(N <= NMS_CAP && (num_classes - 1) * top_k <= MRG_CAP, CROG_E_BADSHAPE, "ssg_fast_nms: at most %d anchors and %d candidates", NMS_CAP, MRG_CAP);
  CROG_REQUIRE(aligned16(boxes) && aligned16(workspace) && ws_stride % 16 == 0, CROG_E_BADALIGN, "ssg_fast_nms: 16B alignment");
  CROG_REQUIRE(B >= 1 && B <= 65535, CROG_E_BADSHAPE, "ssg_fast_nms: batch %d", B);
  cudaStream_t s = (cudaStream_t)stream;
  const int ncls = num_classes - 1;
  unsigned long long* cand = (unsigned long long*)workspace;
  int* cand_keep = (int*)(cand + (size_t)ncls * top_k);
  int* cand_n = cand_keep + (size_t)ncls * top_k;
  static DeviceOnce once;
  int dev_;
  if (once.need(&dev_)) {
    CROG_CUDA_OK(cudaFuncSetAttribute(ssg_nms_class_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NMS_CAP * 8));
    CROG_CUDA_OK(cudaFuncSetAttribute(ssg_nms_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MRG_CAP * 8));
    once.done(dev_);
  }
  ssg_nms_class_kernel<<<dim3(ncls, B), NMS_T, NMS_CAP * 8, s>>>(cls, keep, boxes, N, num_classes, top_k, iou_thr, cand, cand_keep, cand_n,
                                                                 ws_stride);
  CROG_LAUNCH_OK("ssg_nms_class");
  ssg_nms_merge_kernel<<<B, 1024, MRG_CAP * 8, s>>>(cand, cand_keep, cand_n, ncls, top_k, max_det, score_thr2, det_n, det_anchor, det_class,
                                                    det_score, ws_stride);
  CROG_LAUNCH_OK("ssg_nms_merge");
  return CROG_OK;
}
extern "C" int crog_ssg_fast_nms(const float* cls, const int32_t* keep, const float* boxes, int32_t N, int32_t num_classes, float iou_thr,
                                 int32_t top_k, int32_t max_det, float score_thr2, int32_t* det_n, int32_t* det_anchor, int32_t* det_class,
                                 float* det_score, void* workspace, void* stream) {
  return ssg_fast_nms_launch(cls, keep, boxes, 1, N, num_classes, iou_thr, top_k, max_det, score_thr2, det_n, det_anchor, det_class, det_score,
                             workspace, 0, stream);
}
extern "C" int crog_ssg_detect_batched(const float* cls, const float* box, const float* anchors, int32_t B, int32_t N, int32_t num_classes,
                                       float score_thr, float iou_thr, int32_t top_k, int32_t max_det, float score_thr2, int32_t* keep,
                                       float* boxes, int32_t* det_n, int32_t* det_anchor, int32_t* det_class, float* det_score,
                                       void* workspace, int64_t workspace_stride, void* stream) {
  CROG_REQUIRE(N >= 1 && B >= 0 && B <= 65535, CROG_E_BADSHAPE, "ssg_detect_batched: B %d N %d", B, N);
  CROG_REQUIRE(workspace_stride >= crog_ssg_nms_workspace_bytes(num_classes, top_k), CROG_E_BADSHAPE, "ssg_detect_batched: workspace stride too small");
  if (B == 0) return CROG_OK;
  ssg_decode_kernel<<<dim3((N + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(cls, box, anchors, N, num_classes, score_thr, keep, boxes);
  CROG_LAUNCH_OK("ssg_decode");
  return ssg_fast_nms_launch(cls, keep, boxes, B, N, num_classes, iou_thr, top_k, max_det, score_thr2, det_n, det_anchor, det_class, det_score,
                             workspace, workspace_stride, stream);
}
extern "C" int crog_ssg_detect(const float* cls, const float* box, const float* anchors, int32_t N, int32_t num_classes, float score_thr,
                               float iou_thr, int32_t top_k, int32_t max_det, float score_thr2, int32_t* keep, float* boxes,
                               int32_t* det_n, int32_t* det_anchor, int32_t* det_class, float* det_score, void* workspace, void* stream) {
  CROG_REQUIRE(N >= 1, CROG_E_BADSHAPE, "ssg_detect: N %d", N);
  ssg_decode_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(cls, box, anchors, N, num_classes, score_thr, keep, boxes);
  CROG_LAUNCH_OK("ssg_decode");
  return crog_ssg_fast_nms(cls, keep, boxes, N, num_classes, iou_thr, top_k, max_det, score_thr2, det_n, det_anchor, det_class, det_score,
                           workspace, stream);
}

extern "C" int crog_ssg_masks(const float* protos, int32_t h, int32_t w, int32_t num_protos, const float* coef, const float* gcoef,
                              const float* boxes, const int32_t* det_anchor, const int32_t* det_n, int32_t max_det, float* lowres,
                              float* out, float* quality_raw, int32_t out_det_stride, int32_t out_h, int32_t out_w, int32_t resize_to,
                              void* stream) {
  CROG_REQUIRE(num_protos % 4 == 0 && num_protos <= 256 && aligned16(protos), CROG_E_BADSHAPE, "ssg_masks: num_protos %d", num_protos);
  CROG_REQUIRE(out_h <= resize_to && out_w <= resize_to && max_det * 5 <= 65535, CROG_E_BADSHAPE, "ssg_masks: bad output extent");
  if (out_det_stride <= 0) out_det_stride = max_det;
  CROG_REQUIRE(max_det >= 1, CROG_E_BADSHAPE, "ssg_masks: max_det");
  cudaStream_t s = (cudaStream_t)stream;
  dim3 g1((h * w + 255) / 256, max_det);
  ssg_lowres_kernel<<<g1, 256, 5 * num_protos * sizeof(float), s>>>(protos, h, w, num_protos, coef, gcoef, boxes, det_anchor, det_n, lowres,
                                                                    nullptr, nullptr, 0, 0);
  CROG_LAUNCH_OK("ssg_lowres");
  dim3 g2((out_w + 127) / 128, (out_h + BC_ROWS * BC_WARPS - 1) / (BC_ROWS * BC_WARPS), max_det * 5);
  bilinear_crop_kernel<<<g2, dim3(32, BC_WARPS), 0, s>>>(lowres, h, w, out, out_h, out_w, resize_to, det_n, 5, 1u, out_det_stride, quality_raw);
  CROG_LAUNCH_OK("ssg_resize");
  return CROG_OK;
}

extern "C" int crog_ssg_masks_batched(const float* protos, int32_t h, int32_t w, int32_t num_protos, const float* coef, const float* gcoef,
                                      const float* boxes, const int32_t* det_anchor, int32_t N, int32_t max_det,
                                      const int32_t* inst_image, const int32_t* inst_det, int32_t total, float* lowres, float* out,
                                      float* quality_raw, int32_t out_h, int32_t out_w, int32_t resize_to, void* stream) {
  CROG_REQUIRE(num_protos % 4 == 0 && num_protos <= 256 && aligned16(protos), CROG_E_BADSHAPE, "ssg_masks_batched: num_protos %d", num_protos);
  CROG_REQUIRE(out_h <= resize_to && out_w <= resize_to && N >= 1 && max_det >= 1, CROG_E_BADSHAPE, "ssg_masks_batched: bad extent");
  CROG_REQUIRE(inst_image != nullptr && inst_det != nullptr, CROG_E_BADSHAPE, "ssg_masks_batched: instance maps missing");
  if (total <= 0) return CROG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int CH = 13000;  // instances per launch (grid.z = 5 planes per instance <= 65535)
  const long long plane = (long long)out_h * out_w;
  for (int i0 = 0; i0 < total; i0 += CH) {
    const int n = total - i0 < CH ? total - i0 : CH;
    dim3 g1((h * w + 255) / 256, n);
    ssg_lowres_kernel<<<g1, 256, 5 * num_protos * sizeof(float), s>>>(protos, h, w, num_protos, coef, gcoef, boxes, det_anchor, nullptr,
                                                                      lowres + (long long)i0 * 5 * h * w, inst_image + i0, inst_det + i0, N,
                                                                      max_det);
    CROG_LAUNCH_OK("ssg_lowres");
    dim3 g2((out_w + 127) / 128, (out_h + BC_ROWS * BC_WARPS - 1) / (BC_ROWS * BC_WARPS), n * 5);
    bilinear_crop_kernel<<<g2, dim3(32, BC_WARPS), 0, s>>>(lowres + (long long)i0 * 5 * h * w, h, w, out + i0 * plane, out_h, out_w, resize_to, nullptr, 5,
                                            1u, total, quality_raw ? quality_raw + i0 * plane : nullptr);
    CROG_LAUNCH_OK("ssg_resize");
  }
  return CROG_OK;
}

extern "C" int crog_gaussian(const float* in, float* tmp, float* out, int32_t P, int32_t H, int32_t W, const double* weights_host,
                             int32_t radius, const int32_t* n_planes, int32_t plane_stride, int32_t plane_sel, void* stream) {
  const int r = radius;
  CROG_REQUIRE(weights_host != nullptr && r >= 1 && r <= 16, CROG_E_BADSHAPE, "gaussian: radius %d unsupported (1..16)", r);
  CROG_REQUIRE(P <= 65535 && H <= 65535, CROG_E_BADSHAPE, "gaussian: too many planes");
  if (P * H * W == 0) return CROG_OK;
  GaussW g;
  g.r = r;
  for (int i = 0; i <= 2 * r; ++i) g.w[i] = weights_host[i];
  cudaStream_t s = (cudaStream_t)stream;
  if (radius == 8 && in != out && !getenv("CROG_GAUSSIAN_GENERIC") && !getenv("CROG_GAUSSIAN_TWOPASS")) {  // sigma = 2, the reference's (grasp_eval.py:198)
    static_assert((GF_TW + 16) % 4 == 0 && GF_TH % 8 == 0 && GF_TW % 8 == 0, "tile shape");
    gaussian_fused_kernel<8><<<dim3((W + GF_TW - 1) / GF_TW, (H + GF_TH - 1) / GF_TH, P), 256, 0, s>>>(in, out, H, W, n_planes, plane_stride, plane_sel, g);
    CROG_LAUNCH_OK("gaussian_fused");
    return CROG_OK;
  }
  CROG_REQUIRE(tmp != nullptr, CROG_E_BADSHAPE, "gaussian: the two-pass forms (in-place call, radius != 8) need the tmp buffer");
  if (radius == 8 && !getenv("CROG_GAUSSIAN_GENERIC")) {
    gaussian_vert_kernel<8><<<dim3((W + 255) / 256, (H + GV_TY - 1) / GV_TY, P), 256, 0, s>>>(in, tmp, H, W, n_planes, plane_stride, plane_sel, g);
    CROG_LAUNCH_OK("gaussian_rows");
    gaussian_horz_kernel<8><<<dim3((W + 255) / 256, (H + GH_ROWS - 1) / GH_ROWS, P), 256, 0, s>>>(tmp, out, H, W, n_planes, plane_stride, plane_sel, g);
    CROG_LAUNCH_OK("gaussian_cols");
    return CROG_OK;
  }
  dim3 grid((W + 255) / 256, H, P);
  gaussian_pass_kernel<0><<<grid, 256, 0, s>>>(in, tmp, H, W, n_planes, plane_stride, plane_sel, g);
  CROG_LAUNCH_OK("gaussian_rows");
  gaussian_pass_kernel<1><<<grid, 256, 0, s>>>(tmp, out, H, W, n_planes, plane_stride, plane_sel, g);
  CROG_LAUNCH_OK("gaussian_cols");
  return CROG_OK;
}
