// Memory-bound glue kernels of the forward: layout / resampling, stem conv1, LayerNorm,
// embedding, EOT gather, projector weight folding, head split, sigmoid + bicubic.
// All are coalesced over the channel (innermost NHWC) dimension with 8-wide vector access.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ resample (copy / avgpool2 / bilinear x2)
__device__ __forceinline__ long long pix_row(int b, int y, int x, int H, int W, int padded) {
  return padded ? ((long long)(b * (H + 2) + y + 1) * (W + 2) + x + 1) : ((long long)(b * H + y) * W + x);
}

// One CTA per output row (b, oy): 32-bit index math only, threads sweep (ox, 8-channel group) with the channel
// group fastest so every warp access is a contiguous run of 16-byte vectors.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) resample_kernel(const T* __restrict__ in, int in_ld, int in_padded, T* __restrict__ out,
                                                       int out_ld, int out_padded, int B, int H, int W, int C) {
  pdl_launch();
  pdl_wait();
  const int OH = MODE == 1 ? H / 2 : (MODE == 2 ? 2 * H : H);
  const int OW = MODE == 1 ? W / 2 : (MODE == 2 ? 2 * W : W);
  const int cgs = C / 8;
  const int rows_per_b = MODE == 2 ? H + 1 : OH;  // bilinear: one CTA per output row PAIR
  const int b = blockIdx.x / rows_per_b, oy = blockIdx.x % rows_per_b;
  const T* ip = in;
  T* op = out + pix_row(b, MODE == 2 ? 0 : oy, 0, OH, OW, out_padded) * out_ld;
  const int n = OW * cgs;
  if (MODE == 0) {
    ip = in + pix_row(b, oy, 0, H, W, in_padded) * in_ld;
    for (int i = threadIdx.x; i < n; i += 256) {
      const int ox = i / cgs, cg = i - ox * cgs;
      float v[8];
      load8(ip + (long long)ox * in_ld + cg * 8, v);
      store8(op + (long long)ox * out_ld + cg * 8, v);
    }
  } else if (MODE == 1) {
    const T* r0 = in + pix_row(b, 2 * oy, 0, H, W, in_padded) * in_ld;
    const T* r1 = in + pix_row(b, 2 * oy + 1, 0, H, W, in_padded) * in_ld;
    for (int i = threadIdx.x; i < n; i += 256) {
      const int ox = i / cgs, cg = i - ox * cgs;
      float v[8], a[8], c[8], d[8];
      load8(r0 + (long long)(2 * ox) * in_ld + cg * 8, v);
      load8(r0 + (long long)(2 * ox + 1) * in_ld + cg * 8, a);
      load8(r1 + (long long)(2 * ox) * in_ld + cg * 8, c);
      load8(r1 + (long long)(2 * ox + 1) * in_ld + cg * 8, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (v[j] + a[j] + c[j] + d[j]) * 0.25f;
      store8(op + (long long)ox * out_ld + cg * 8, v);
    }
  } else {
    // F.interpolate(scale_factor=2, mode='bilinear', align_corners=False).  Output rows 2k-1 and 2k read input rows
    // k-1 and k (clamped) with weights (.75,.25) / (.25,.75), and the same holds for columns, so one thread turns a
    // 2x2 input neighbourhood (four 16-byte loads) into a 2x2 output block: a quarter of the loads and about a
    // third of the instructions of the one-output-per-thread form.  CTA = output row pair k in [0, H].
    const int k = oy;  // here blockIdx.x enumerates (b, k) with k in [0, H]  (OH := H + 1 in the launch)
    const int ya = max(k - 1, 0), yb = min(k, H - 1);
    const T* r0 = in + pix_row(b, ya, 0, H, W, in_padded) * in_ld;
    const T* r1 = in + pix_row(b, yb, 0, H, W, in_padded) * in_ld;
    const int oy0 = 2 * k - 1, oy1 = 2 * k;  // valid if in [0, 2H)
    T* o0 = oy0 >= 0 ? out + pix_row(b, oy0, 0, 2 * H, 2 * W, out_padded) * out_ld : nullptr;
    T* o1 = oy1 < 2 * H ? out + pix_row(b, oy1, 0, 2 * H, 2 * W, out_padded) * out_ld : nullptr;
    const int nj = (W + 1) * cgs;
    for (int i = threadIdx.x; i < nj; i += 256) {
      const int j = i / cgs, cg = i - j * cgs;
      const int xa = max(j - 1, 0), xb = min(j, W - 1);
      float p[8], q[8], r[8], t[8];  // (ya,xa) (ya,xb) (yb,xa) (yb,xb)
      load8(r0 + (long long)xa * in_ld + cg * 8, p);
      load8(r0 + (long long)xb * in_ld + cg * 8, q);
      load8(r1 + (long long)xa * in_ld + cg * 8, r);
      load8(r1 + (long long)xb * in_ld + cg * 8, t);
      const int ox0 = 2 * j - 1, ox1 = 2 * j;
      // the reference computes hy*(hx*a + lx*b) + ly*(hx*c + lx*d) with (l, h) = (.25, .75) for odd and (.75, .25) for
      // even output coordinates; at the clamped borders both taps coincide, so the same formula holds there
      float v[8];
#pragma unroll
      for (int half = 0; half < 4; ++half) {
        const bool oddy = (half & 2) == 0, oddx = (half & 1) == 0;  // half 0: (oy0, ox0), 1: (oy0, ox1), 2: (oy1, ox0), 3: (oy1, ox1)
        T* orow = oddy ? o0 : o1;
        const int ox = oddx ? ox0 : ox1;
        if (orow == nullptr || ox < 0 || ox >= 2 * W) continue;
        const float ly = (!oddy && k == 0) ? 0.f : (oddy ? 0.25f : 0.75f), hy = 1.f - ly;  // row / column 0: sy clamps to 0
        const float lx = (!oddx && j == 0) ? 0.f : (oddx ? 0.25f : 0.75f), hx = 1.f - lx;
        // odd outputs (2k-1): sy = k - 0.75 -> y0 = k-1, ly = .25; even (2k): sy = k - .25 -> y0 = k-1, ly = .75;
        // row 0 / col 0 clamp to sy = 0 (ly = 0): with ya == yb the blend is the same value
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = hy * (hx * p[c] + lx * q[c]) + ly * (hx * r[c] + lx * t[c]);
        store8(orow + (long long)ox * out_ld + cg * 8, v);
      }
    }
  }
}

// ------------------------------------------------------------------ stem conv1 (3->cout, 3x3, stride 2, pad 1)
// Thread = (two horizontally adjacent output pixels) x (8 output channels): the 3 x 5 x 3 input window is loaded once
// for both pixels (45 scalar loads instead of 54) and every weight vector read from shared memory feeds 16 FMAs
// (the one-pixel form was bound by its LDS stream at 4 FMAs per LDS.128).  Channel groups are the fastest thread
// index, so a pixel's cout channels are one contiguous store.  Only channels [0, cout) are written: the padding
// channels of the zero-initialised plan buffer stay zero.
template <typename T>
__global__ void __launch_bounds__(128) stem_conv1_kernel(const float* __restrict__ img, int B, int Hin, int Win,
                                                          const float* __restrict__ w, const float* __restrict__ scale,
                                                          const float* __restrict__ bias, int cout, T* __restrict__ out, int out_ld,
                                                          int pairs) {
  extern __shared__ __align__(16) float sw[];  // [27][cout] + scale + bias
  pdl_launch();  // the weights below are constants of the plan: staging them overlaps the predecessor's tail
  for (int i = threadIdx.x; i < 27 * cout; i += blockDim.x) {
    const int co = i % cout, k = i / cout;
    sw[i] = w[co * 27 + k];
  }
  for (int i = threadIdx.x; i < cout; i += blockDim.x) { sw[27 * cout + i] = scale[i]; sw[28 * cout + i] = bias[i]; }
  __syncthreads();
  pdl_wait();
  // thread = one output pixel pair, ALL channels: the 45 input values (3 channels x 3 rows x 5 columns) are loaded once
  // and reused by every 8-channel group (a thread per group re-loaded them cout/8 times and ran at a quarter of the
  // store bandwidth); in the pixel-pair layout a thread's two pixels are one contiguous 2*cout-channel row.
  const int OH = Hin / 2, OW = Win / 2, OWP = (OW + 1) / 2, cgs = cout / 8;
  const long long total = (long long)B * OH * OWP;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int oxp = (int)(t % OWP);
    t /= OWP;
    const int oy = (int)(t % OH), b = (int)(t / OH);
    const int ox = 2 * oxp;
    float x[3][3][5];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = 2 * oy + ky - 1;
        const float* rowp = img + ((long long)(b * 3 + ci) * Hin + iy) * Win;
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
          const int ix = 2 * ox + kx - 1;
          x[ci][ky][kx] = (iy >= 0 && iy < Hin && ix >= 0 && ix < Win) ? __ldg(rowp + ix) : 0.f;
        }
      }
    T* o0 = pairs ? out + ((long long)(b * (OH + 2) + oy + 1) * (OWP + 2) + oxp + 1) * out_ld
                  : out + pix_row(b, oy, ox, OH, OW, 1) * out_ld;
    T* o1 = pairs ? o0 + cout : o0 + out_ld;
    for (int cg = 0; cg < cgs; ++cg) {
      float v0[8], v1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { v0[j] = 0.f; v1[j] = 0.f; }
      const int c0 = cg * 8;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int k = ci * 9 + ky * 3 + kx;
            const float4 w0 = *reinterpret_cast<const float4*>(&sw[k * cout + c0]);
            const float4 w1 = *reinterpret_cast<const float4*>(&sw[k * cout + c0 + 4]);
            const float a = x[ci][ky][kx], c = x[ci][ky][kx + 2];
            // packed FFMA2: (v[2j], v[2j+1]) += (a, a) * (w[2j], w[2j+1]) — the same two fused multiply-adds, one instruction
            fma2_bcast(v0[0], v0[1], a, w0.x, w0.y); fma2_bcast(v0[2], v0[3], a, w0.z, w0.w);
            fma2_bcast(v0[4], v0[5], a, w1.x, w1.y); fma2_bcast(v0[6], v0[7], a, w1.z, w1.w);
            fma2_bcast(v1[0], v1[1], c, w0.x, w0.y); fma2_bcast(v1[2], v1[3], c, w0.z, w0.w);
            fma2_bcast(v1[4], v1[5], c, w1.x, w1.y); fma2_bcast(v1[6], v1[7], c, w1.z, w1.w);
          }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float sc = sw[27 * cout + c0 + j], bi = sw[28 * cout + c0 + j];
        v0[j] = fmaxf(v0[j] * sc + bi, 0.f);
        v1[j] = fmaxf(v1[j] * sc + bi, 0.f);
      }
      store8(o0 + c0, v0);
      if (pairs || ox + 1 < OW) store8(o1 + c0, v1);
    }
  }
}

// ------------------------------------------------------------------ LayerNorm (one warp per row)
template <typename TI, typename TO, int CH>  // D = CH * 256
__global__ void layernorm_kernel(const TI* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ residual, TO* __restrict__ out, long long rows, float eps) {
  constexpr int D = CH * 256;
  pdl_launch();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[CH][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    load8(x + row * D + (c * 32 + lane) * 8, v[c]);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[c][j];
  }
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = v[c][j] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (c * 32 + lane) * 8;
    float g8[8], b8[8], o8[8];
    load8(gamma + col, g8); load8(beta + col, b8);
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = (v[c][j] - mean) * rstd * g8[j] + b8[j];
    if (residual) {
      float r8[8]; load8(residual + row * D + col, r8);
#pragma unroll
      for (int j = 0; j < 8; ++j) o8[j] += r8[j];
    }
    store8(out + row * D + col, o8);
  }
}

// Two chained LayerNorms of the decoder (layers.py:318-329): y = residual + LN1(x) is the new fp32 residual stream, and
// the next sub-layer immediately normalises it again (z = LN2(y)).  One warp per row keeps y in registers, so the second
// normalisation costs no second pass over HBM; arithmetic and rounding are those of two layernorm_kernel launches.
template <typename TI, typename TO, int CH>
__global__ void layernorm_chain_kernel(const TI* __restrict__ x, const float* __restrict__ g1, const float* __restrict__ b1,
                                       const float* residual, float* y, const float* __restrict__ g2,
                                       const float* __restrict__ b2, TO* __restrict__ z, long long rows, float eps) {
  constexpr int D = CH * 256;
  pdl_launch();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[CH][8], r[CH][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    load8(x + row * D + (c * 32 + lane) * 8, v[c]);
    load8(residual + row * D + (c * 32 + lane) * 8, r[c]);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[c][j];
  }
  float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = v[c][j] - mean; q += d * d; }
  float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
  s = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (c * 32 + lane) * 8;
    float g8[8], b8[8];
    load8(g1 + col, g8); load8(b1 + col, b8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float o = (v[c][j] - mean) * rstd * g8[j] + b8[j];
      o += r[c][j];
      v[c][j] = o;
      s += o;
    }
    store8(y + row * D + col, v[c]);
  }
  mean = warp_sum(s) * (1.f / D);
  q = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = v[c][j] - mean; q += d * d; }
  rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (c * 32 + lane) * 8;
    float g8[8], b8[8], o8[8];
    load8(g2 + col, g8); load8(b2 + col, b8);
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = (v[c][j] - mean) * rstd * g8[j] + b8[j];
    store8(z + row * D + col, o8);
  }
}

// ------------------------------------------------------------------ text embedding / EOT gather
__global__ void embed_kernel(const int64_t* __restrict__ word, const float* __restrict__ emb, const float* __restrict__ pos,
                             float* __restrict__ out, int B, int L, int D, int vocab) {
  pdl_launch();
  pdl_wait();
  const int row = blockIdx.x, l = row % L;
  const long long id = word[row];
  if (id < 0 || id >= vocab) {  // nn.Embedding's device-side assert (clip.py:440): fail loudly, never read out of bounds
    if (threadIdx.x == 0) printf("crog_embed_tokens: token id %lld at row %d outside [0, %d)\n", id, row, vocab);
    __trap();
  }
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    const float4 e = *reinterpret_cast<const float4*>(emb + id * D + c);
    const float4 p = *reinterpret_cast<const float4*>(pos + (long long)l * D + c);
    *reinterpret_cast<float4*>(out + (long long)row * D + c) = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
  }
}

template <typename TI, typename TO>
__global__ void gather_eot_kernel(const int64_t* __restrict__ word, const TI* __restrict__ x, TO* __restrict__ out, int L, int D) {
  pdl_launch();
  pdl_wait();
  const int b = blockIdx.x;
  __shared__ int s_arg;
  if (threadIdx.x == 0) {
    int arg = 0; long long best = word[(long long)b * L];
    for (int l = 1; l < L; ++l) { const long long t = word[(long long)b * L + l]; if (t > best) { best = t; arg = l; } }
    s_arg = arg;
  }
  __syncthreads();
  const TI* src = x + ((long long)b * L + s_arg) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) out[(long long)b * D + c] = from_f<TO>(to_f(src[c]));
}

// ------------------------------------------------------------------ projector: txt linear + fold with vis.4
template <typename T>
__global__ void txt_linear_kernel(const T* __restrict__ state, const float* __restrict__ w, const float* __restrict__ bias,
                                  float* __restrict__ out, int B, int K, int O) {
  pdl_launch();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long wid = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)B * O) return;
  const int o = (int)(wid % O), b = (int)(wid / O);
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(to_f(state[(long long)b * K + k]), __ldg(w + (long long)o * K + k), acc);
  acc = warp_sum(acc);
  if (lane == 0) out[(long long)b * O + o] = acc + bias[o];
}

template <typename T>
__global__ void dynw_fold_kernel(const float* __restrict__ t, const float* __restrict__ vw, const float* __restrict__ vb,
                                 T* __restrict__ wfold, int B, int C, int NH, int rows_per_sample, int Cpad) {
  pdl_launch();
  pdl_wait();
  // one block per (b, h, tap); threads over j.  Output row of sample b: h*9 + tap.
  const int tap = blockIdx.x % 9, h = (blockIdx.x / 9) % NH, b = blockIdx.x / (9 * NH);
  const int O = 9 * C + 1;
  extern __shared__ float s_w[];  // w_dyn[b, :, tap]
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_w[c] = t[(long long)b * O + c * 9 + tap];
  __syncthreads();
  T* dst = wfold + ((long long)b * rows_per_sample + h * 9 + tap) * Cpad;
  for (int j = threadIdx.x; j < Cpad; j += blockDim.x) {
    float acc = 0.f;
    if (j < C) {
      for (int c = 0; c < C; ++c) acc = fmaf(s_w[c], __ldg(vw + (long long)(h * C + c) * C + j), acc);
    } else if (j == C) {
      for (int c = 0; c < C; ++c) acc = fmaf(s_w[c], __ldg(vb + h * C + c), acc);
      if (tap == 4) acc += t[(long long)b * O + 9 * C];
    }
    dst[j] = from_f<T>(acc);
  }
}

// out[h][b, y, x] = sum_tap Z[padded_row(b, y+ky-1, x+kx-1), h*9 + tap]: the nine shifted reads of the per-tap partial
// products that the per-sample 1x1 GEMM left in the zero-haloed Z matrix.
__global__ void dynconv_gather_kernel(const float* __restrict__ z, int ldz, float* __restrict__ out, int B, int H, int W, int NH) {
  pdl_launch();
  pdl_wait();
  const long long total = (long long)B * H * W;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / ((long long)W * H));
  const int PW = W + 2;
  const long long base = ((long long)b * (H + 2) + y) * PW + x;  // padded row of (y-1, x-1)
  for (int h = 0; h < NH; ++h) {
    float acc = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
      acc += __ldg(z + (base + (tap / 3) * PW + (tap % 3)) * ldz + h * 9 + tap);
    out[(long long)h * total + i] = acc;
  }
}

// Tiled form: a CTA owns GT_H x GT_W output pixels, stages the (GT_H + 2) x (GT_W + 2) rows of Z it needs in shared memory
// with coalesced 16-byte loads (a Z row is ldz = 48 contiguous floats; adjacent pixels are adjacent rows), then every thread
// sums its nine taps from shared memory.  The per-thread form above reads 45 scalars at a 192-byte stride between lanes:
// 32 cache lines per warp load, 95 us per 64 samples at 1.5 TB/s; same arithmetic, same summation order.
constexpr int GT_H = 8, GT_W = 32;
__global__ void __launch_bounds__(GT_H * GT_W) dynconv_gather_tiled_kernel(const float* __restrict__ z, int ldz, float* __restrict__ out, int B,
                                                                      int H, int W, int NH) {
  extern __shared__ float s_z[];  // [(GT_H + 2) * (GT_W + 2)][ldz + 1]
  pdl_launch();
  pdl_wait();
  const int b = blockIdx.z, y0 = blockIdx.y * GT_H, x0 = blockIdx.x * GT_W;
  const int PW = W + 2, pitch = ldz + 1;
  const int rows = (GT_H + 2) * (GT_W + 2), q4 = ldz / 4;
  for (int i = threadIdx.x; i < rows * q4; i += blockDim.x) {
    const int r = i / q4, c4 = i - r * q4;
    const int ty = r / (GT_W + 2), tx = r - ty * (GT_W + 2);
    const int py = y0 + ty, px = x0 + tx;  // padded coordinates of (y - 1 + ty, x - 1 + tx)
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (py < H + 2 && px < PW) v = __ldg(reinterpret_cast<const float4*>(z + (((long long)b * (H + 2) + py) * PW + px) * ldz) + c4);
    float* d = s_z + r * pitch + c4 * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const int ly = threadIdx.x / GT_W, lx = threadIdx.x % GT_W;
  const int y = y0 + ly, x = x0 + lx;
  if (y >= H || x >= W) return;
  const long long total = (long long)B * H * W, i = ((long long)b * H + y) * W + x;
  for (int h = 0; h < NH; ++h) {
    float acc = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
      acc += s_z[((ly + tap / 3) * (GT_W + 2) + lx + tap % 3) * pitch + h * 9 + tap];
    out[(long long)h * total + i] = acc;
  }
}

__global__ void split_heads_kernel(const float* __restrict__ in, int ld, float* __restrict__ out, long long rows, int NH) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= rows) return;
  for (int h = 0; h < NH; ++h) out[(long long)h * rows + i] = in[i * ld + h];
}

// ------------------------------------------------------------------ sigmoid + bicubic (align_corners=True, A=-0.75)
__device__ __forceinline__ float cc1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cc2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__global__ void sigmoid_bicubic_kernel(const float* __restrict__ in, float* __restrict__ out, int NP, int B, int Hin, int Win,
                                       int Hout, int Wout, uint32_t sig_mask, float sh, float sw) {
  const long long total = (long long)NP * B * Hout * Wout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wout);
    const int oy = (int)((i / Wout) % Hout);
    const long long pl = i / ((long long)Wout * Hout);  // plane index p*B + b
    const int p = (int)(pl / B);
    const bool sig = (sig_mask >> p) & 1u;
    const float ry = sh * oy, rx = sw * ox;
    const int iy = (int)floorf(ry), ix = (int)floorf(rx);
    const float ty = ry - iy, tx = rx - ix;
    const float wx[4] = {cc2(tx + 1.f), cc1(tx), cc1(1.f - tx), cc2(2.f - tx)};
    const float wy[4] = {cc2(ty + 1.f), cc1(ty), cc1(1.f - ty), cc2(2.f - ty)};
    const float* src = in + pl * (long long)Hin * Win;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), Hin - 1);
      float r = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int xx = min(max(ix - 1 + c, 0), Win - 1);
        float v = __ldg(src + (long long)yy * Win + xx);
        if (sig) v = 1.f / (1.f + expf(-v));
        r += v * wx[c];
      }
      acc += r * wy[a];
    }
    out[i] = acc;
  }
}

// Tiled variant: one CTA produces a 64 x 64 output tile from a <= 32 x 24 source tile staged (with the sigmoid
// applied once per source pixel) in shared memory; stores are coalesced 128 B rows.  (Measured at 5 x 64 maps
// 104 -> 416, L2 flushed: 32-row tiles 111 us, 64-row 88 us, 128-row 85 us.)  A thread owns one output column
// and BT_RPT consecutive output rows: the horizontal weights are computed once, and the horizontally filtered source rows
// slide as a 4-row window down the column (a new source row costs 4 shared-memory reads + 4 FMA, and at scale 1/4 only
// every fourth output row needs one), so an output costs ~15 instructions instead of ~100.
#ifndef CROG_BT_OH
#define CROG_BT_OH 64
#endif
constexpr int BT_OW = 64, BT_OH = CROG_BT_OH, BT_SW = 32, BT_SH = BT_OH / 4 + 8, BT_RPT = BT_OH / 4;  // rows per thread
__global__ void __launch_bounds__(256) sigmoid_bicubic_tiled_kernel(const float* __restrict__ in, float* __restrict__ out, int B,
                                                                    int Hin, int Win, int Hout, int Wout, uint32_t sig_mask,
                                                                    float sh, float sw) {
  __shared__ float tile[BT_SH][BT_SW + 1];
  __shared__ float4 s_wy[BT_OH];  // vertical weights of the tile's 32 output rows (the same for every column)
  __shared__ int s_sy[BT_OH];     // first source row of each output row, relative to the staged tile
  const int pl = blockIdx.z, p = pl / B;
  const bool sig = (sig_mask >> p) & 1u;
  const int ox0 = blockIdx.x * BT_OW, oy0 = blockIdx.y * BT_OH;
  const int ix_lo = (int)floorf(sw * ox0) - 1, iy_lo = (int)floorf(sh * oy0) - 1;
  const float* src = in + (long long)pl * Hin * Win;
  for (int i = threadIdx.x; i < BT_SH * BT_SW; i += 256) {
    const int ty = i / BT_SW, tx = i % BT_SW;
    const int yy = min(max(iy_lo + ty, 0), Hin - 1), xx = min(max(ix_lo + tx, 0), Win - 1);
    float v = __ldg(src + (long long)yy * Win + xx);
    if (sig) v = 1.f / (1.f + expf(-v));
    tile[ty][tx] = v;
  }
  if (threadIdx.x < BT_OH) {
    const float ry = sh * (oy0 + (int)threadIdx.x);
    const int iy = (int)floorf(ry);
    const float fy = ry - iy;
    s_wy[threadIdx.x] = make_float4(cc2(fy + 1.f), cc1(fy), cc1(1.f - fy), cc2(2.f - fy));
    s_sy[threadIdx.x] = iy - 1 - iy_lo;
  }
  __syncthreads();
  const int rg = threadIdx.x >> 6;  // row group: BT_RPT consecutive output rows
  const int ox = ox0 + (threadIdx.x & 63), oyb = oy0 + rg * BT_RPT;  // a warp = 32 consecutive columns, same rows
  if (ox >= Wout) return;
  const float rx = sw * ox;
  const int ix = (int)floorf(rx);
  const float fx = rx - ix;
  const float wx[4] = {cc2(fx + 1.f), cc1(fx), cc1(1.f - fx), cc2(2.f - fx)};
  const int sx = ix - 1 - ix_lo;
  auto hrow = [&](int row) {  // horizontally filtered source row (same summation order as the reference formulation)
    float r = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) r += tile[row][sx + c] * wx[c];
    return r;
  };
  float* dst = out + (long long)pl * Hout * Wout + ox;
  int cur = -1000;
  float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f;
#pragma unroll
  for (int j = 0; j < BT_RPT; ++j) {
    const int oy = oyb + j;
    if (oy >= Hout) break;
    const int sy = s_sy[rg * BT_RPT + j];  // warp-uniform
    const float4 wy = s_wy[rg * BT_RPT + j];
    if (sy != cur) {
      if (sy == cur + 1) { h0 = h1; h1 = h2; h2 = h3; h3 = hrow(sy + 3); }
      else { h0 = hrow(sy); h1 = hrow(sy + 1); h2 = hrow(sy + 2); h3 = hrow(sy + 3); }
      cur = sy;
    }
    float acc = 0.f;
    acc += h0 * wy.x;
    acc += h1 * wy.y;
    acc += h2 * wy.z;
    acc += h3 * wy.w;
    dst[(long long)oy * Wout] = acc;
  }
}

inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int crog_resample2(const void* in, int in_ld, int in_padded, void* out, int out_ld, int out_padded, int B, int H, int W, int C, int mode,
                   int dtype, cudaStream_t s);

extern "C" int crog_resample(const void* in, int32_t in_ld, int32_t in_padded, void* out, int32_t out_ld, int32_t out_padded,
                             int32_t B, int32_t H, int32_t W, int32_t C, int32_t mode, int32_t dtype, void* stream) {
  CROG_REQUIRE(C % 8 == 0 && in_ld % 8 == 0 && out_ld % 8 == 0, CROG_E_BADSHAPE, "resample: channels must be multiples of 8");
  CROG_REQUIRE(aligned16(in) && aligned16(out), CROG_E_BADALIGN, "resample: pointers must be 16B aligned");
  CROG_REQUIRE(mode >= 0 && mode <= 4, CROG_E_BADSHAPE, "resample: bad mode %d", mode);
  if (mode >= 3) return crog_resample2(in, in_ld, in_padded, out, out_ld, out_padded, B, H, W, C, mode, dtype, (cudaStream_t)stream);
  if (mode == 1) CROG_REQUIRE(H % 2 == 0 && W % 2 == 0, CROG_E_BADSHAPE, "avgpool2 needs even H,W");
  const int OH = mode == 1 ? H / 2 : (mode == 2 ? 2 * H : H), OW = mode == 1 ? W / 2 : (mode == 2 ? 2 * W : W);
  if ((long long)B * OH * OW * C == 0) return CROG_OK;
  const int g = mode == 2 ? B * (H + 1) : B * OH;
  cudaStream_t s = (cudaStream_t)stream;
#define RS(T, M) crog_launch(resample_kernel<T, M>, dim3(g), dim3(256), 0, s, (const T*)in, in_ld, in_padded, (T*)out, out_ld, out_padded, B, H, W, C)
  if (dtype == CROG_F32) { if (mode == 0) RS(float, 0); else if (mode == 1) RS(float, 1); else RS(float, 2); }
  else { if (mode == 0) RS(bf16, 0); else if (mode == 1) RS(bf16, 1); else RS(bf16, 2); }
#undef RS
  CROG_LAUNCH_OK("resample");
  return CROG_OK;
}

extern "C" int crog_stem_conv1(const float* img, int32_t B, int32_t Hin, int32_t Win, const float* w, const float* scale,
                               const float* bias, int32_t cout, void* out, int32_t out_ld, int32_t out_dtype, int32_t pixel_pairs,
                               void* stream) {
  CROG_REQUIRE(Hin % 2 == 0 && Win % 2 == 0 && out_ld % 8 == 0 && cout <= out_ld && cout % 8 == 0, CROG_E_BADSHAPE, "stem_conv1: bad shape");
  CROG_REQUIRE(!pixel_pairs || (Win % 4 == 0 && 2 * cout <= out_ld), CROG_E_BADSHAPE, "stem_conv1: the pixel-pair layout needs an even output width and out_ld >= 2*cout");
  const long long total = (long long)B * (Hin / 2) * ((Win / 2 + 1) / 2);  // one thread per output pixel pair
  if (total == 0) return CROG_OK;
  const int g = grid_for(total, 128);
  const size_t sm = (size_t)(29 * cout) * sizeof(float);
  if (out_dtype == CROG_F32)
    crog_launch(stem_conv1_kernel<float>, dim3(g), dim3(128), sm, (cudaStream_t)stream, img, B, Hin, Win, w, scale, bias, cout, (float*)out, out_ld, pixel_pairs);
  else
    crog_launch(stem_conv1_kernel<bf16>, dim3(g), dim3(128), sm, (cudaStream_t)stream, img, B, Hin, Win, w, scale, bias, cout, (bf16*)out, out_ld, pixel_pairs);
  CROG_LAUNCH_OK("stem_conv1");
  return CROG_OK;
}

extern "C" int crog_layernorm(const void* x, int32_t x_dtype, const float* gamma, const float* beta, const float* residual,
                              void* out, int32_t out_dtype, int64_t rows, int32_t D, float eps, void* stream) {
  CROG_REQUIRE(D == 512 || D == 2048 || D == 1024 || D == 256, CROG_E_BADSHAPE, "layernorm: D=%d unsupported", D);
  if (rows == 0) return CROG_OK;
  const int wpb = 8;
  const int g = (int)((rows + wpb - 1) / wpb);
  cudaStream_t s = (cudaStream_t)stream;
#define LN(TI, TO, CH) crog_launch(layernorm_kernel<TI, TO, CH>, dim3(g), dim3(wpb * 32), 0, s, (const TI*)x, gamma, beta, residual, (TO*)out, (long long)rows, eps)
#define LN_D(TI, TO) do { if (D == 256) LN(TI, TO, 1); else if (D == 512) LN(TI, TO, 2); else if (D == 1024) LN(TI, TO, 4); else LN(TI, TO, 8); } while (0)
  if (x_dtype == CROG_F32 && out_dtype == CROG_F32) LN_D(float, float);
  else if (x_dtype == CROG_F32) LN_D(float, bf16);
  else if (out_dtype == CROG_F32) LN_D(bf16, float);
  else LN_D(bf16, bf16);
#undef LN_D
#undef LN
  CROG_LAUNCH_OK("layernorm");
  return CROG_OK;
}

extern "C" int crog_layernorm_chain(const void* x, int32_t x_dtype, const float* g1, const float* b1, const float* residual,
                                    float* y, const float* g2, const float* b2, void* z, int32_t z_dtype, int64_t rows,
                                    int32_t D, float eps, void* stream) {
  CROG_REQUIRE(D == 512 || D == 256, CROG_E_BADSHAPE, "layernorm_chain: D=%d unsupported", D);
  CROG_REQUIRE(x && g1 && b1 && residual && y && g2 && b2 && z, CROG_E_BADSHAPE, "layernorm_chain: null operand");
  if (rows == 0) return CROG_OK;
  const int wpb = 8;
  const int g = (int)((rows + wpb - 1) / wpb);
  cudaStream_t s = (cudaStream_t)stream;
#define LNC(TI, TO, CH) crog_launch(layernorm_chain_kernel<TI, TO, CH>, dim3(g), dim3(wpb * 32), 0, s, (const TI*)x, g1, b1, residual, y, g2, b2, (TO*)z, (long long)rows, eps)
#define LNC_D(TI, TO) do { if (D == 256) LNC(TI, TO, 1); else LNC(TI, TO, 2); } while (0)
  if (x_dtype == CROG_F32 && z_dtype == CROG_F32) LNC_D(float, float);
  else if (x_dtype == CROG_F32) LNC_D(float, bf16);
  else if (z_dtype == CROG_F32) LNC_D(bf16, float);
  else LNC_D(bf16, bf16);
#undef LNC_D
#undef LNC
  CROG_LAUNCH_OK("layernorm_chain");
  return CROG_OK;
}

extern "C" int crog_embed_tokens(const int64_t* word, const float* emb, const float* pos, float* out, int32_t B, int32_t L,
                                 int32_t D, int32_t vocab, void* stream) {
  CROG_REQUIRE(D % 4 == 0, CROG_E_BADSHAPE, "embed: D %% 4");
  CROG_REQUIRE(vocab > 0, CROG_E_BADSHAPE, "embed: vocab > 0");
  if (B * L == 0) return CROG_OK;
  crog_launch(embed_kernel, dim3(B * L), dim3(128), 0, (cudaStream_t)stream, word, emb, pos, out, B, L, D, vocab);
  CROG_LAUNCH_OK("embed");
  return CROG_OK;
}

extern "C" int crog_gather_eot(const int64_t* word, const void* x, int32_t x_dtype, void* out, int32_t out_dtype, int32_t B,
                               int32_t L, int32_t D, void* stream) {
  if (B == 0) return CROG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (x_dtype == CROG_F32 && out_dtype == CROG_F32) crog_launch(gather_eot_kernel<float, float>, dim3(B), dim3(128), 0, s, word, (const float*)x, (float*)out, L, D);
  else if (x_dtype == CROG_F32) crog_launch(gather_eot_kernel<float, bf16>, dim3(B), dim3(128), 0, s, word, (const float*)x, (bf16*)out, L, D);
  else if (out_dtype == CROG_F32) crog_launch(gather_eot_kernel<bf16, float>, dim3(B), dim3(128), 0, s, word, (const bf16*)x, (float*)out, L, D);
  else crog_launch(gather_eot_kernel<bf16, bf16>, dim3(B), dim3(128), 0, s, word, (const bf16*)x, (bf16*)out, L, D);
  CROG_LAUNCH_OK("gather_eot");
  return CROG_OK;
}

extern "C" int crog_dynw_fold(const void* state, int32_t state_dtype, const float* txt_w, const float* txt_b, const float* v_w,
                              const float* v_b, float* scratch, void* wfold, int32_t dtype, int32_t B, int32_t word_dim,
                              int32_t C, int32_t NH, int32_t NH_pad, int32_t Cpad, void* stream) {
  CROG_REQUIRE(Cpad > C && 9 * NH <= NH_pad, CROG_E_BADSHAPE, "dynw_fold: need Cpad > C (ones channel) and rows_per_sample >= 9*NH");
  if (B == 0) return CROG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int O = 9 * C + 1;
  const long long warps = (long long)B * O;
  const int g1 = (int)((warps + 7) / 8);
  if (state_dtype == CROG_F32) crog_launch(txt_linear_kernel<float>, dim3(g1), dim3(256), 0, s, (const float*)state, txt_w, txt_b, scratch, B, word_dim, O);
  else crog_launch(txt_linear_kernel<bf16>, dim3(g1), dim3(256), 0, s, (const bf16*)state, txt_w, txt_b, scratch, B, word_dim, O);
  CROG_LAUNCH_OK("txt_linear");
  const int g2 = B * NH * 9;
  if (dtype == CROG_F32) crog_launch(dynw_fold_kernel<float>, dim3(g2), dim3(128), C * sizeof(float), s, scratch, v_w, v_b, (float*)wfold, B, C, NH, NH_pad, Cpad);
  else crog_launch(dynw_fold_kernel<bf16>, dim3(g2), dim3(128), C * sizeof(float), s, scratch, v_w, v_b, (bf16*)wfold, B, C, NH, NH_pad, Cpad);
  CROG_LAUNCH_OK("dynw_fold");
  return CROG_OK;
}

extern "C" int crog_dynconv_gather(const float* z, int32_t ldz, float* out, int32_t B, int32_t H, int32_t W, int32_t NH, void* stream) {
  const long long total = (long long)B * H * W;
  if (total == 0) return CROG_OK;
  const size_t smem = (size_t)(GT_H + 2) * (GT_W + 2) * (ldz + 1) * sizeof(float);
  if (ldz % 4 == 0 && aligned16(z) && smem <= 100 * 1024 && B <= 65535 && !getenv("CROG_GATHER_SIMPLE")) {
    static DeviceOnce once;
    int dev_;
    if (once.need(&dev_)) {
      CROG_CUDA_OK(cudaFuncSetAttribute(dynconv_gather_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      once.done(dev_);
    }
    crog_launch(dynconv_gather_tiled_kernel, dim3((W + GT_W - 1) / GT_W, (H + GT_H - 1) / GT_H, B), dim3(GT_H * GT_W), smem, (cudaStream_t)stream,
                z, ldz, out, B, H, W, NH);
    CROG_LAUNCH_OK("dynconv_gather");
    return CROG_OK;
  }
  crog_launch(dynconv_gather_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, z, ldz, out, B, H, W, NH);
  CROG_LAUNCH_OK("dynconv_gather");
  return CROG_OK;
}

extern "C" int crog_split_heads(const float* in, int32_t ld, float* out, int64_t rows, int32_t NH, void* stream) {
  if (rows == 0) return CROG_OK;
  split_heads_kernel<<<(int)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, ld, out, rows, NH);
  CROG_LAUNCH_OK("split_heads");
  return CROG_OK;
}

extern "C" int crog_sigmoid_bicubic(const float* in, float* out, int32_t NP, int32_t B, int32_t Hin, int32_t Win, int32_t Hout,
                                    int32_t Wout, uint32_t sigmoid_mask, void* stream) {
  const long long total = (long long)NP * B * Hout * Wout;
  if (total == 0) return CROG_OK;
  const float sh = Hout > 1 ? (float)(Hin - 1) / (float)(Hout - 1) : 0.f;
  const float sw = Wout > 1 ? (float)(Win - 1) / (float)(Wout - 1) : 0.f;
  // source extent of one output tile (+1 on each side for the 4-tap support, +1 for floor slack)
  const bool fits = (int)(sw * (BT_OW - 1)) + 5 <= BT_SW && (int)(sh * (BT_OH - 1)) + 5 <= BT_SH && (long long)NP * B <= 65535;
  if (fits) {
    dim3 grid((Wout + BT_OW - 1) / BT_OW, (Hout + BT_OH - 1) / BT_OH, NP * B);
    sigmoid_bicubic_tiled_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, B, Hin, Win, Hout, Wout, sigmoid_mask, sh, sw);
  } else {
    sigmoid_bicubic_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, NP, B, Hin, Win, Hout, Wout, sigmoid_mask, sh, sw);
  }
  CROG_LAUNCH_OK("sigmoid_bicubic");
  return CROG_OK;
}
