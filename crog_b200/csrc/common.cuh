// Shared device/host helpers for the crog_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/crog_b200.h"

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- host-side error plumbing
void crog_set_error(const char* fmt, ...);
#define CROG_FAIL(code, ...)          \
  do {                                \
    crog_set_error(__VA_ARGS__);      \
    return (code);                    \
  } while (0)
#define CROG_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) CROG_FAIL(code, __VA_ARGS__); \
  } while (0)
#define CROG_CUDA_OK(expr)                                                           \
  do {                                                                               \
    cudaError_t e__ = (expr);                                                        \
    if (e__ != cudaSuccess) CROG_FAIL(CROG_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

// Kernel function attributes (dynamic shared memory opt-in, carve-out) are PER DEVICE.  One bit per device ordinal, set
// after the attributes have been applied on that device; safe against concurrent host threads (two threads may both apply
// the same attributes once, which is harmless).  Usage:
//   static DeviceOnce once; int dev; if (once.need(&dev)) { cudaFuncSetAttribute(...); once.done(dev); }
struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  bool need(int* dev) const {
    *dev = 0;
    if (cudaGetDevice(dev) != cudaSuccess) return true;
    return *dev >= 64 || !((mask.load(std::memory_order_acquire) >> *dev) & 1ull);
  }
  void done(int dev) {
    if (dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release);
  }
};
#define CROG_LAUNCH_OK(name)                                                        \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) CROG_FAIL(CROG_E_CUDA, "launch %s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// With CROG_PDL=1 every kernel of the forward is launched with programmatic stream serialization (off by default: it
// measured slower under graph replay): it may become resident while its predecessor drains, runs its prologue (barrier init, TMEM allocation, descriptor prefetch, index math) and then
// blocks in pdl_wait() until the predecessor grid has completed and its writes are visible.  pdl_launch() at the top
// of a kernel lets ITS successor do the same.  Under stream capture these become programmatic graph edges.
// Rule: no global-memory access that depends on (or is depended on by) an earlier kernel before pdl_wait().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool crog_pdl_enabled();  // CROG_PDL=1 turns the launch attribute on (without it the device-side instructions are no-ops)
template <typename... KA, typename... A>
static inline cudaError_t crog_launch(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = crog_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KA>(args)...);
}

// ---------------------------------------------------------------- 8-wide vector access
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
__device__ __forceinline__ void store4(bf16* p, float a, float b, float c, float d) {
  uint2 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
  h[0] = __floats2bfloat162_rn(a, b); h[1] = __floats2bfloat162_rn(c, d);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- GEMM row mapping + epilogue
// (shared by the SIMT and tcgen05 kernels so both produce the same bytes for the same accumulators)
struct RowMap {
  int valid;        // row produces output
  int b;            // sample index
  int sp;           // interior pixel index (y*W+x) or row-in-sample
  int orow;         // output row (row counts fit 32 bits; multiply by the leading dimension in 64 bits)
  float rs, sh;     // folded LayerNorm of the input rows (CrogGemm.row_stats_in): rstd and -mean * rstd; 1 and 0 otherwise
};

__device__ __forceinline__ RowMap map_row(const CrogGemm& g, long long r64, long long row_end) {
  RowMap m;
  const int r = (int)r64;
  m.valid = r64 < row_end;
  m.b = 0; m.sp = 0; m.orow = r;
  m.rs = 1.f; m.sh = 0.f;
  if (!m.valid) return m;
  if (g.H == 0) {
    if (g.sample_rows > 0) { m.b = r / g.sample_rows; m.sp = r - m.b * g.sample_rows; }
    else m.sp = r;
    return m;
  }
  int y, x;
  if (g.in_padded) {
    const int PW = g.W + 2, P = (g.H + 2) * PW;
    m.b = r / P;
    const int rem = r - m.b * P;
    const int yy = rem / PW, xx = rem - yy * PW;
    m.valid = (yy >= 1) && (yy <= g.H) && (xx >= 1) && (xx <= g.W);
    y = yy - 1; x = xx - 1;
  } else {
    const int P = g.H * g.W;
    m.b = r / P;
    const int rem = r - m.b * P;
    y = rem / g.W; x = rem - y * g.W;
  }
  m.sp = y * g.W + x;
  if (g.out_padded) m.orow = (m.b * (g.H + 2) + y + 1) * (g.W + 2) + x + 1;
  else m.orow = m.b * (g.out_sample_rows > 0 ? g.out_sample_rows : g.H * g.W) + m.sp;
  return m;
}

// Row statistics of a folded LayerNorm: sums the producer's per-chunk (sum, sum of squares) pairs of this row.
__device__ __forceinline__ void load_row_stats(const CrogGemm& g, RowMap& m) {
  if (!g.row_stats_in || !m.valid) return;
  // 16-byte loads, eight in flight (a rolled loop of dependent 8-byte loads costs one L2 round trip per chunk)
  const float4* p = reinterpret_cast<const float4*>(g.row_stats_in + (long long)m.orow * g.row_stats_chunks * 2);
  float s = 0.f, q = 0.f;
  const int n4 = g.row_stats_chunks >> 1;  // two (sum, sum^2) pairs per load; row_stats_chunks is even (checked on the host), so every row is 16-byte aligned
#pragma unroll 8
  for (int i = 0; i < n4; ++i) { const float4 v = __ldg(p + i); s += v.x + v.z; q += v.y + v.w; }
  const float inv = 1.f / (float)g.row_stats_width;
  const float mean = s * inv;
  const float var = fmaxf(q * inv - mean * mean, 0.f);
  m.rs = rsqrtf(var + g.row_stats_eps);
  m.sh = -mean * m.rs;
}

__device__ __forceinline__ float quickgelu(float v) { return __fdividef(v, 1.f + __expf(-1.702f * v)); }

// Epilogue math for CNT (multiple of 8) consecutive columns starting at n0 of one valid row: everything except the
// residual add and the store.  sc/bi point at scale/bias for column n0 (global memory, 16-byte aligned; every lane of a
// warp reads the same addresses, so the loads are L1 broadcasts), or nullptr.  All mode tests are hoisted out of the per-element loops.
template <int CNT>
__device__ __forceinline__ void epilogue_math(const CrogGemm& g, const RowMap& m, int n0, float (&acc)[CNT], const float* sc,
                                              const float* bi) {
  const int nvalid = min(CNT, g.N - n0);
  if (g.addmat) {
    const float* ad = g.addmat + (long long)m.sp * g.N + n0;
    if (nvalid == CNT) {
#pragma unroll
      for (int j = 0; j < CNT; j += 4) {
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(ad + j));
        acc[j] += a4.x; acc[j + 1] += a4.y; acc[j + 2] += a4.z; acc[j + 3] += a4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CNT; ++j) if (j < nvalid) acc[j] += __ldg(ad + j);
    }
  }
  // scale and bias as one FFMA (a missing scale is 1, a missing bias 0, which keeps acc * s and acc + b exact), so
  // this path and the shared-memory one of the tcgen05 kernels (epilogue_math_smem) produce the same bits
  if (g.row_stats_in) {  // folded LayerNorm: acc * rstd_r + (sh_r * s_n + c_n); sc = s, bi = c (both required)
    if (nvalid == CNT) {
#pragma unroll
      for (int j = 0; j < CNT; j += 4) {
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + j)), b4 = __ldg(reinterpret_cast<const float4*>(bi + j));
        acc[j] = fmaf(acc[j], m.rs, fmaf(m.sh, s4.x, b4.x)); acc[j + 1] = fmaf(acc[j + 1], m.rs, fmaf(m.sh, s4.y, b4.y));
        acc[j + 2] = fmaf(acc[j + 2], m.rs, fmaf(m.sh, s4.z, b4.z)); acc[j + 3] = fmaf(acc[j + 3], m.rs, fmaf(m.sh, s4.w, b4.w));
      }
    } else {
#pragma unroll
      for (int j = 0; j < CNT; ++j) if (j < nvalid) acc[j] = fmaf(acc[j], m.rs, fmaf(m.sh, __ldg(sc + j), __ldg(bi + j)));
    }
  } else if (sc || bi) {
    if (nvalid == CNT) {
#pragma unroll
      for (int j = 0; j < CNT; j += 4) {
        const float4 s4 = sc ? __ldg(reinterpret_cast<const float4*>(sc + j)) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 b4 = bi ? __ldg(reinterpret_cast<const float4*>(bi + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[j] = fmaf(acc[j], s4.x, b4.x); acc[j + 1] = fmaf(acc[j + 1], s4.y, b4.y);
        acc[j + 2] = fmaf(acc[j + 2], s4.z, b4.z); acc[j + 3] = fmaf(acc[j + 3], s4.w, b4.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < CNT; ++j) if (j < nvalid) acc[j] = fmaf(acc[j], sc ? __ldg(sc + j) : 1.f, bi ? __ldg(bi + j) : 0.f);
    }
  }
  if (g.act == CROG_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < CNT; ++j) acc[j] = fmaxf(acc[j], 0.f);
  } else if (g.act == CROG_ACT_QUICKGELU) {
#pragma unroll
    for (int j = 0; j < CNT; ++j) acc[j] = quickgelu(acc[j]);
  } else if (g.act == CROG_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < CNT; ++j) acc[j] = tanhf(acc[j]);
  }
  if (g.gate) {
    const float* gt = g.gate + (long long)m.b * g.N + n0;
#pragma unroll
    for (int j = 0; j < CNT; ++j) if (j < nvalid)
      acc[j] = fmaxf((acc[j] * __ldg(gt + j)) * __ldg(g.scale2 + n0 + j) + __ldg(g.bias2 + n0 + j), 0.f);
  }
}

// Packed fp32 pairs (sm_100 FFMA2 / FADD2): two independent round-to-nearest operations per instruction, bit-identical
// to the scalar forms.  The epilogues issue about one instruction per four clocks per warp, so halving the count of
// the scale / bias and residual arithmetic is a direct saving.
__device__ __forceinline__ void ffma2(float& a0, float& a1, float s0, float s1, float b0, float b1) {
  unsigned long long a, sv, bv;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(sv) : "f"(s0), "f"(s1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bv) : "f"(b0), "f"(b1));
  asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(sv), "l"(bv));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
}
// (v0, v1) = (a, a) * (w0, w1) + (v0, v1)
__device__ __forceinline__ void fma2_bcast(float& v0, float& v1, float a, float w0, float w1) {
  unsigned long long v, av, wv;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(v0), "f"(v1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(wv) : "f"(w0), "f"(w1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(v) : "l"(av), "l"(wv));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(v));
}
__device__ __forceinline__ void fadd2(float& a0, float& a1, float b0, float b1) {
  unsigned long long a, bv;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bv) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(bv));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
}

// The same epilogue math with scale / bias staged in shared memory by the caller (s_sc / s_bi point at the CNT values of
// columns n0.., 16-byte aligned, already holding 1 / 0 where the layer has no scale / bias or the column is >= N): every
// lane reads the same addresses, so the loads are LDS broadcasts instead of four dependent L2 round trips per 64
// columns, and scale and bias are applied as one FFMA.  (In-kernel clock64 trace on B200: the __ldg version spends
// 1850 of 3800 clocks per 64-column chunk in this step.)
template <int CNT>
__device__ __forceinline__ void epilogue_math_smem(const CrogGemm& g, const RowMap& m, int n0, float (&acc)[CNT], const float* s_sc,
                                                   const float* s_bi, bool relu_later = false) {
  const int nvalid = min(CNT, g.N - n0);
  if (g.addmat) {
    const float* ad = g.addmat + (long long)m.sp * g.N + n0;
    if (nvalid == CNT) {
#pragma unroll
      for (int j = 0; j < CNT; j += 4) {
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(ad + j));
        acc[j] += a4.x; acc[j + 1] += a4.y; acc[j + 2] += a4.z; acc[j + 3] += a4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CNT; ++j) if (j < nvalid) acc[j] += __ldg(ad + j);
    }
  }
  if (g.row_stats_in) {  // folded LayerNorm (see epilogue_math): same arithmetic, s and c from shared memory
#pragma unroll
    for (int j = 0; j < CNT; j += 4) {
      const float4 s4 = *reinterpret_cast<const float4*>(s_sc + j), b4 = *reinterpret_cast<const float4*>(s_bi + j);
      acc[j] = fmaf(acc[j], m.rs, fmaf(m.sh, s4.x, b4.x)); acc[j + 1] = fmaf(acc[j + 1], m.rs, fmaf(m.sh, s4.y, b4.y));
      acc[j + 2] = fmaf(acc[j + 2], m.rs, fmaf(m.sh, s4.z, b4.z)); acc[j + 3] = fmaf(acc[j + 3], m.rs, fmaf(m.sh, s4.w, b4.w));
    }
  } else {
#pragma unroll
    for (int j = 0; j < CNT; j += 4) {
      const float4 s4 = *reinterpret_cast<const float4*>(s_sc + j);
      const float4 b4 = *reinterpret_cast<const float4*>(s_bi + j);
      ffma2(acc[j], acc[j + 1], s4.x, s4.y, b4.x, b4.y);
      ffma2(acc[j + 2], acc[j + 3], s4.z, s4.w, b4.z, b4.w);
    }
  }
  if (g.act == CROG_ACT_RELU) {
    if (!relu_later) {  // relu_later: the caller clamps the packed bf16 pairs instead (same bits, half the instructions)
#pragma unroll
      for (int j = 0; j < CNT; ++j) acc[j] = fmaxf(acc[j], 0.f);
    }
  } else if (g.act == CROG_ACT_QUICKGELU) {
#pragma unroll
    for (int j = 0; j < CNT; ++j) acc[j] = quickgelu(acc[j]);
  } else if (g.act == CROG_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < CNT; ++j) acc[j] = tanhf(acc[j]);
  }
  if (g.gate) {
    const float* gt = g.gate + (long long)m.b * g.N + n0;
#pragma unroll
    for (int j = 0; j < CNT; ++j) if (j < nvalid)
      acc[j] = fmaxf((acc[j] * __ldg(gt + j)) * __ldg(g.scale2 + n0 + j) + __ldg(g.bias2 + n0 + j), 0.f);
  }
}

// Full epilogue (math + residual + store) for one valid row; used by the CUDA-core GEMM.
template <int CNT>
__device__ __forceinline__ void epilogue_row(const CrogGemm& g, const RowMap& m, int n0, float (&acc)[CNT],
                                             const float* sc, const float* bi) {
  const int nvalid = min(CNT, g.N - n0);
  if (nvalid <= 0) return;
  epilogue_math<CNT>(g, m, n0, acc, sc, bi);
  const bool full = (nvalid == CNT);
  if (g.out_dtype == CROG_BF16) {
    bf16* o = reinterpret_cast<bf16*>(g.out) + (long long)m.orow * g.out_ld + n0;
    const bf16* rs = g.residual ? reinterpret_cast<const bf16*>(g.residual) + (long long)m.orow * g.res_ld + n0 : nullptr;
    if (full) {
#pragma unroll
      for (int j = 0; j < CNT; j += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = acc[j + i];
        if (rs) {
          float r8[8]; load8(rs + j, r8);
#pragma unroll
          for (int i = 0; i < 8; ++i) { v[i] += r8[i]; if (g.residual_relu) v[i] = fmaxf(v[i], 0.f); }
        }
        store8(o + j, v);
      }
    } else {
      for (int j = 0; j < nvalid; ++j) {
        float v = acc[j];
        if (rs) { v += to_f(rs[j]); if (g.residual_relu) v = fmaxf(v, 0.f); }
        o[j] = __float2bfloat16_rn(v);
      }
    }
  } else {
    float* o = reinterpret_cast<float*>(g.out) + (long long)m.orow * g.out_ld + n0;
    const float* rs = g.residual ? reinterpret_cast<const float*>(g.residual) + (long long)m.orow * g.res_ld + n0 : nullptr;
    if (full) {
#pragma unroll
      for (int j = 0; j < CNT; j += 4) {
        float4 v = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        if (rs) {
          float4 r4 = *reinterpret_cast<const float4*>(rs + j);
          v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
          if (g.residual_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        }
        *reinterpret_cast<float4*>(o + j) = v;
      }
    } else {
      for (int j = 0; j < nvalid; ++j) {
        float v = acc[j];
        if (rs) { v += rs[j]; if (g.residual_relu) v = fmaxf(v, 0.f); }
        o[j] = v;
      }
    }
  }
}

// tile -> rows. With per-sample weights tiles never straddle samples.
struct TileRows { long long row0, row_end; int wsample; };
__device__ __forceinline__ TileRows tile_rows(const CrogGemm& g, int m_tile, int BM) {
  TileRows t;
  if (g.w_sample_stride > 0) {
    const int tps = (g.sample_rows + BM - 1) / BM;
    t.wsample = m_tile / tps;
    t.row0 = (long long)t.wsample * g.sample_rows + (long long)(m_tile % tps) * BM;
    t.row_end = min((long long)(t.wsample + 1) * g.sample_rows, (long long)g.M);
  } else {
    t.wsample = 0; t.row0 = (long long)m_tile * BM; t.row_end = g.M;
  }
  return t;
}
static inline int num_m_tiles(const CrogGemm& g, int BM) {
  if (g.w_sample_stride > 0) return (g.M / g.sample_rows) * ((g.sample_rows + BM - 1) / BM);
  return (g.M + BM - 1) / BM;
}
__host__ __device__ __forceinline__ int tap_shift(int taps, int tap, int W) {
  return taps == 9 ? (tap / 3 - 1) * (W + 2) + (tap % 3 - 1) : 0;
}
