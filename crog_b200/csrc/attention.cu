// Softmax attention for head_dim 64 on CUDA cores, and the dispatcher of crog_attention:
//   * bf16 attentions with >= 128 queries or keys (attention pool, decoder self- and cross-attention) go to the tcgen05
//     kernel of attention_tc.cu;
//   * short sequences (the causal text tower, L <= 96) run attention_small_kernel: one CTA per (sample, head), one
//     thread per score, one warp per softmax row, one thread per output quad;
//   * everything else (fp32 mode with long sequences) runs the streaming kernel: one query per thread, K/V tiles of
//     64 keys staged in shared memory and broadcast to the warp.
// fp32 math throughout.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int KT = 64;   // keys per smem tile
constexpr int QB = 128;  // queries per block (one per thread)

template <typename T>
__global__ void __launch_bounds__(QB) attention_kernel(const T* __restrict__ q, int ldq, const T* __restrict__ k, int ldk,
                                                       const T* __restrict__ v, int ldv, T* __restrict__ o, int ldo, int Tq,
                                                       int Tk, float scale, int causal, const int64_t* __restrict__ pad_word) {
  __shared__ float Ks[KT][HD];
  __shared__ float Vs[KT][HD];
  __shared__ int s_masked[KT];
  pdl_launch();
  pdl_wait();
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * QB + threadIdx.x;
  const bool qvalid = qi < Tq;
  float qr[HD], acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  if (qvalid) {
    const T* qp = q + ((long long)b * Tq + qi) * ldq + h * HD;
#pragma unroll
    for (int d = 0; d < HD; d += 8) {
      float t[8]; load8(qp + d, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) qr[d + j] = t[j] * scale;
    }
  } else {
#pragma unroll
    for (int d = 0; d < HD; ++d) qr[d] = 0.f;
  }
  float mx = -INFINITY, den = 0.f;
  const int q_hi = min(blockIdx.x * QB + QB - 1, Tq - 1);
  const int k_end = causal ? min(Tk, q_hi + 1) : Tk;
  for (int k0 = 0; k0 < k_end; k0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * (HD / 8); i += QB) {
      const int kr = i / (HD / 8), d8 = (i % (HD / 8)) * 8;
      float tk[8], tv[8];
      if (k0 + kr < Tk) {
        load8(k + ((long long)b * Tk + k0 + kr) * ldk + h * HD + d8, tk);
        load8(v + ((long long)b * Tk + k0 + kr) * ldv + h * HD + d8, tv);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { tk[j] = 0.f; tv[j] = 0.f; }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { Ks[kr][d8 + j] = tk[j]; Vs[kr][d8 + j] = tv[j]; }
    }
    for (int i = threadIdx.x; i < KT; i += QB) {
      const int kk = k0 + i;
      s_masked[i] = (kk >= Tk) || (pad_word && pad_word[(long long)b * Tk + kk] == 0);
    }
    __syncthreads();
    const int kn = min(KT, k_end - k0);
    for (int j0 = 0; j0 < kn; j0 += 8) {
      float s[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = j0 + jj;
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int d = 0; d < HD; d += 8) {
          const float4 a = *reinterpret_cast<const float4*>(&Ks[j][d]);
          const float4 c = *reinterpret_cast<const float4*>(&Ks[j][d + 4]);
          d0 = fmaf(qr[d], a.x, d0); d1 = fmaf(qr[d + 1], a.y, d1); d0 = fmaf(qr[d + 2], a.z, d0); d1 = fmaf(qr[d + 3], a.w, d1);
          d0 = fmaf(qr[d + 4], c.x, d0); d1 = fmaf(qr[d + 5], c.y, d1); d0 = fmaf(qr[d + 6], c.z, d0); d1 = fmaf(qr[d + 7], c.w, d1);
        }
        const bool dead = (j >= kn) || s_masked[j] || (causal && (k0 + j) > qi);
        s[jj] = dead ? -INFINITY : (d0 + d1);
      }
      float tmax = s[0];
#pragma unroll
      for (int jj = 1; jj < 8; ++jj) tmax = fmaxf(tmax, s[jj]);
      if (tmax == -INFINITY) continue;  // all eight keys masked for this query
      const float nm = fmaxf(mx, tmax);
      const float corr = __expf(mx - nm);  // mx=-inf -> 0
      den *= corr;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] *= corr;
      mx = nm;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float p = __expf(s[jj] - nm);  // -inf -> 0
        den += p;
        const int j = min(j0 + jj, KT - 1);
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          const float4 vv = *reinterpret_cast<const float4*>(&Vs[j][d]);
          acc[d] = fmaf(p, vv.x, acc[d]); acc[d + 1] = fmaf(p, vv.y, acc[d + 1]);
          acc[d + 2] = fmaf(p, vv.z, acc[d + 2]); acc[d + 3] = fmaf(p, vv.w, acc[d + 3]);
        }
      }
    }
  }
  if (qvalid) {
    const float inv = 1.f / den;
    T* op = o + ((long long)b * Tq + qi) * ldo + h * HD;
#pragma unroll
    for (int d = 0; d < HD; d += 8) {
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = acc[d + j] * inv;
      store8(op + d, t);
    }
  }
}


// Small attention (text tower: L <= 77 queries and keys, 8 heads; clip.py:258-262): one CTA per (sample, head), every
// score and every output element computed by its own thread instead of one thread walking a whole query row
// (the streaming kernel above keeps 17 of 128 threads busy for ~2700 dependent instructions at L = 17: 41 us per layer,
// a third of the text tower).  Phase 1: S[i][j] = scale q_i . k_j for all visible pairs; phase 2: row softmax by one
// warp per row; phase 3: o[i][d..d+3] = sum_j P[i][j] v[j][d..d+3].  fp32 throughout; K rows padded against bank conflicts.
constexpr int SM_T = 96;      // max tokens
constexpr int SM_KP = HD + 4;  // padded K row (floats)
template <typename T>
__global__ void __launch_bounds__(256) attention_small_kernel(const T* __restrict__ q, int ldq, const T* __restrict__ k, int ldk,
                                                              const T* __restrict__ v, int ldv, T* __restrict__ o, int ldo, int Tq,
                                                              int Tk, float scale, int causal, const int64_t* __restrict__ pad_word) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                          // [Tq][HD]
  float* Ks = Qs + Tq * HD;                // [Tk][SM_KP]
  float* Vs = Ks + Tk * SM_KP;             // [Tk][HD]
  float* Ss = Vs + Tk * HD;                // [Tq][Tk + 1]
  pdl_launch();
  pdl_wait();
  const int b = blockIdx.y, h = blockIdx.x, tid = threadIdx.x, SP = Tk + 1;
  for (int i = tid; i < Tq * (HD / 8); i += 256) {
    const int r = i / (HD / 8), d8 = (i % (HD / 8)) * 8;
    float t[8];
    load8(q + ((long long)b * Tq + r) * ldq + h * HD + d8, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) Qs[r * HD + d8 + j] = t[j] * scale;
  }
  for (int i = tid; i < Tk * (HD / 8); i += 256) {
    const int r = i / (HD / 8), d8 = (i % (HD / 8)) * 8;
    float tk[8], tv[8];
    load8(k + ((long long)b * Tk + r) * ldk + h * HD + d8, tk);
    load8(v + ((long long)b * Tk + r) * ldv + h * HD + d8, tv);
#pragma unroll
    for (int j = 0; j < 8; ++j) { Ks[r * SM_KP + d8 + j] = tk[j]; Vs[r * HD + d8 + j] = tv[j]; }
  }
  __syncthreads();
  for (int p = tid; p < Tq * Tk; p += 256) {
    const int i = p / Tk, j = p - i * Tk;
    const bool dead = (causal && j > i) || (pad_word && pad_word[(long long)b * Tk + j] == 0);
    float d0 = 0.f, d1 = 0.f;
    if (!dead) {
      const float4* qa = reinterpret_cast<const float4*>(Qs + i * HD);
      const float4* ka = reinterpret_cast<const float4*>(Ks + j * SM_KP);
#pragma unroll
      for (int d = 0; d < HD / 4; d += 2) {
        const float4 a = qa[d], c = ka[d], a2 = qa[d + 1], c2 = ka[d + 1];
        d0 = fmaf(a.x, c.x, d0); d1 = fmaf(a.y, c.y, d1); d0 = fmaf(a.z, c.z, d0); d1 = fmaf(a.w, c.w, d1);
        d0 = fmaf(a2.x, c2.x, d0); d1 = fmaf(a2.y, c2.y, d1); d0 = fmaf(a2.z, c2.z, d0); d1 = fmaf(a2.w, c2.w, d1);
      }
    }
    Ss[i * SP + j] = dead ? -INFINITY : d0 + d1;
  }
  __syncthreads();
  for (int i = tid >> 5; i < Tq; i += 8) {  // one warp per row
    const int lane = tid & 31;
    float m = -INFINITY;
    for (int j = lane; j < Tk; j += 32) m = fmaxf(m, Ss[i * SP + j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < Tk; j += 32) {
      const float pj = __expf(Ss[i * SP + j] - m);  // -inf -> 0
      Ss[i * SP + j] = pj;
      sum += pj;
    }
    sum = warp_sum(sum);
    if (lane == 0) Ss[i * SP + Tk] = 1.f / sum;
  }
  __syncthreads();
  for (int p = tid; p < Tq * (HD / 4); p += 256) {
    const int i = p / (HD / 4), d4 = (p % (HD / 4)) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int jn = causal ? min(Tk, i + 1) : Tk;
    for (int j = 0; j < jn; ++j) {
      const float pj = Ss[i * SP + j];
      const float4 vv = *reinterpret_cast<const float4*>(Vs + j * HD + d4);
      acc.x = fmaf(pj, vv.x, acc.x); acc.y = fmaf(pj, vv.y, acc.y); acc.z = fmaf(pj, vv.z, acc.z); acc.w = fmaf(pj, vv.w, acc.w);
    }
    const float inv = Ss[i * SP + Tk];
    T* op = o + ((long long)b * Tq + i) * ldo + h * HD + d4;
    store4(op, acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
}

}  // namespace

int crog_attention_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int B, int heads,
                      int Tq, int Tk, float scale, const int64_t* pad_word, cudaStream_t stream);

extern "C" int crog_attention(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv, void* o,
                              int32_t ldo, int32_t B, int32_t heads, int32_t Tq, int32_t Tk, float scale, int32_t causal,
                              const int64_t* pad_word, int32_t dtype, void* stream) {
  CROG_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, CROG_E_BADSHAPE, "attention: ld %% 8");
  CROG_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), CROG_E_BADALIGN, "attention: 16B alignment");
  if (B == 0 || Tq == 0) return CROG_OK;
  CROG_REQUIRE(B <= 65535 && heads <= 65535, CROG_E_BADSHAPE, "attention: grid too large");
  cudaStream_t s = (cudaStream_t)stream;
  // bf16 attentions with many queries go to the tcgen05 kernel: the dense self-attentions, and the cross-attention
  // over <= 77 word tokens (one masked key tile; scores and P.V on the tensor core instead of 4352 LDS + FMA per
  // query on CUDA cores).  The causal text attention (L <= 77 queries) stays on CUDA cores.
  if (dtype == CROG_BF16 && !causal && (Tk >= 128 || (Tq >= 128 && !getenv("CROG_ATTN_CROSS_SIMT"))))
    return crog_attention_tc(q, ldq, k, ldk, v, ldv, o, ldo, B, heads, Tq, Tk, scale, (const int64_t*)pad_word, s);
  if (Tq <= SM_T && Tk <= SM_T && !getenv("CROG_ATTN_NO_SMALL")) {  // text tower (and any other short sequence)
    const size_t smem = ((size_t)Tq * HD + (size_t)Tk * SM_KP + (size_t)Tk * HD + (size_t)Tq * (Tk + 1)) * sizeof(float);
    static DeviceOnce once;
    int dev_;
    if (once.need(&dev_)) {
      const int mx = (SM_T * HD + SM_T * SM_KP + SM_T * HD + SM_T * (SM_T + 1)) * (int)sizeof(float);
      CROG_CUDA_OK(cudaFuncSetAttribute(attention_small_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
      CROG_CUDA_OK(cudaFuncSetAttribute(attention_small_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
      once.done(dev_);
    }
    dim3 g2(heads, B);
    if (dtype == CROG_F32)
      crog_launch(attention_small_kernel<float>, g2, dim3(256), smem, s, (const float*)q, ldq, (const float*)k, ldk, (const float*)v, ldv, (float*)o, ldo, Tq, Tk, scale, causal, pad_word);
    else
      crog_launch(attention_small_kernel<bf16>, g2, dim3(256), smem, s, (const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (bf16*)o, ldo, Tq, Tk, scale, causal, pad_word);
    CROG_LAUNCH_OK("attention_small");
    return CROG_OK;
  }
  dim3 grid((Tq + QB - 1) / QB, heads, B);
  if (dtype == CROG_F32)
    crog_launch(attention_kernel<float>, grid, dim3(QB), 0, s, (const float*)q, ldq, (const float*)k, ldk, (const float*)v, ldv, (float*)o, ldo, Tq, Tk, scale, causal, pad_word);
  else
    crog_launch(attention_kernel<bf16>, grid, dim3(QB), 0, s, (const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (bf16*)o, ldo, Tq, Tk, scale, causal, pad_word);
  CROG_LAUNCH_OK("attention");
  return CROG_OK;
}
