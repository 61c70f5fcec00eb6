// Grasp decode + Jaccard tail (utils/grasp_eval.py:289-374 of the reference) on the device.
//
//  K9  peak_scan_kernel   streaming 5x5 max-filter equality + threshold + border test; every warp
//                         owns a 128-column strip of a row band, keeps a 5-row window in registers
//                         (float4 per lane, neighbours by shuffle) and ballot-compacts candidates
//                         into a per-warp shared-memory list, then keeps its best TSEL keys.
//  K10 peak_select_kernel one warp per map: merges the per-warp lists, runs the greedy min-distance
//                         suppression in key order (value desc, row-major index asc), proves the
//                         result exact (or flags the map), gathers sin/cos/width and decodes grasps.
//      peak_exact_kernel  exact fallback for flagged maps (heavy plateaus): K full-map arg-max sweeps.
//  K11 jaccard_kernel     one CTA per sample: integer rasterisation of rotated rectangles as
//                         128 x 160-bit row masks in shared memory, popc(A&B) / popc(A|B), J flags,
//                         counters accumulated with one atomicAdd per sample.
#include <math.h>
#include <stdlib.h>

#include "tc_common.cuh"
#include "tail_geom.h"

namespace {

constexpr int BAND_SMALL = 32; // output rows per CTA band when few maps must fill the GPU
constexpr int BAND_LARGE = 104; // ... and when there are plenty (4 halo rows per band: 3.8 % instead of 12.5 % re-read)
constexpr int STRIP = 120;     // columns owned per warp (30 lanes x 4; lanes 0 and 31 carry the halo)
#ifndef CROG_CAPW
#define CROG_CAPW 512
#endif
constexpr int CAPW = CROG_CAPW;      // per-warp candidate list capacity
constexpr int TSEL = 32;       // keys kept per warp segment
constexpr int MAXK = 32;

struct SegHeader {
  int count;
  int truncated;
  unsigned long long worst_kept;  // valid when truncated
  int nonconst;                   // some owned pixel differs from pixel (0,0): the map is not "trivial" (A.1 step 2)
  int pad[3];
};
constexpr int SEG_BYTES = (int)sizeof(SegHeader) + TSEL * 8;

__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ unsigned long long make_key(float v, int idx) {
  return ((unsigned long long)f2ord(v) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)idx);
}
__device__ __forceinline__ int key_idx(unsigned long long k) { return (int)(0xffffffffu - (uint32_t)(k & 0xffffffffull)); }
__device__ __forceinline__ float key_val(unsigned long long k) { return ord2f((uint32_t)(k >> 32)); }
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, k, o);
    k = t > k ? t : k;
  }
  return k;
}
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
// largest float below x (finite x)
__device__ __forceinline__ float float_prev(float x) {
  if (x == 0.f) return -__uint_as_float(1u);
  const uint32_t u = __float_as_uint(x);
  return __uint_as_float(x > 0.f ? u - 1u : u + 1u);
}

// ---- warp-wide top-32 selection on 64-bit keys (bitonic networks over the 32 lanes, one key per lane and register).
// A selection is ~30 dependent shuffle steps (about 1k cycles) instead of `keep` serial arg-max sweeps over the list
// (about 400 cycles each): short enough to hide inside the scan's 20-row staging slack, so a warp that selects does not
// stall the CTA's bulk-copy ring.
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long k, int m) {
  return __shfl_xor_sync(0xffffffffu, k, m);
}
__device__ __forceinline__ unsigned long long sort32_desc(unsigned long long k, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const unsigned long long o = shfl_xor_u64(k, stride);
      const bool desc = (lane & size) == 0;      // direction of this lane's block (size 32: one descending block)
      const bool lower = (lane & stride) == 0;
      const unsigned long long mx = k > o ? k : o, mn = k > o ? o : k;
      k = (lower == desc) ? mx : mn;
    }
  }
  return k;
}
// a, b sorted descending over the lanes -> the 32 largest of the 64 keys, sorted descending
__device__ __forceinline__ unsigned long long merge_top32(unsigned long long a, unsigned long long b, int lane) {
  const unsigned long long br = shfl_xor_u64(b, 31);  // b reversed
  unsigned long long t = a > br ? a : br;              // bitonic sequence holding the top 32
#pragma unroll
  for (int stride = 16; stride > 0; stride >>= 1) {
    const unsigned long long o = shfl_xor_u64(t, stride);
    const bool lower = (lane & stride) == 0;
    const unsigned long long mx = t > o ? t : o, mn = t > o ? o : t;
    t = lower ? mx : mn;
  }
  return t;
}
// keep the best `keep` (<= 32) keys of list[0..n) (n <= CAPW; keys are unique and non-zero) at the front, descending
__device__ unsigned long long warp_select_top(unsigned long long* list, int n, int keep, unsigned long long* /*scratch*/, int lane) {
  unsigned long long acc = 0ull;
  for (int base = 0; base < n; base += 128) {  // four independent 32-key sorts in flight, then a merge tree
    unsigned long long k0 = base + lane < n ? list[base + lane] : 0ull;
    unsigned long long k1 = base + 32 + lane < n ? list[base + 32 + lane] : 0ull;
    unsigned long long k2 = base + 64 + lane < n ? list[base + 64 + lane] : 0ull;
    unsigned long long k3 = base + 96 + lane < n ? list[base + 96 + lane] : 0ull;
    k0 = sort32_desc(k0, lane);
    if (base + 32 < n) {
      k1 = sort32_desc(k1, lane);
      k0 = merge_top32(k0, k1, lane);
      if (base + 64 < n) {
        k2 = sort32_desc(k2, lane);
        if (base + 96 < n) { k3 = sort32_desc(k3, lane); k2 = merge_top32(k2, k3, lane); }
        k0 = merge_top32(k0, k2, lane);
      }
    }
    acc = base == 0 ? k0 : merge_top32(acc, k0, lane);
  }
  __syncwarp();
  if (lane < keep) list[lane] = acc;
  __syncwarp();
  return acc;  // lane i holds the i-th largest key (0: fewer than i + 1 keys)
}

// Cold path of the scan (kept out of line so its registers do not count against the streaming loop): keep the best tsel
// keys, then apply the value cut.  Returns the new list length; *bound rises to cover everything dropped.
__device__ __noinline__ int select_and_cut(unsigned long long* list, int count, int tsel, int kdist, int lane,
                                           unsigned long long* bound_io) {
  unsigned long long bound = *bound_io;
  const unsigned long long a = warp_select_top(list, count, tsel, nullptr, lane);
  if (count > tsel) { const unsigned long long w = list[tsel - 1]; bound = w > bound ? w : bound; count = tsel; }
  // value cut: the kdist-th distinct value among the kept (sorted) keys
  const uint32_t hv = (uint32_t)(a >> 32), pv = __shfl_up_sync(0xffffffffu, hv, 1);
  const bool valid = lane < count;
  const uint32_t mnew = __ballot_sync(0xffffffffu, valid && (lane == 0 || hv != pv));
  if (__popc(mnew) >= kdist) {
    const int pos = __fns(mnew, 0, kdist);
    const uint32_t cvv = __shfl_sync(0xffffffffu, hv, pos);
    const unsigned long long ck = (unsigned long long)cvv << 32;  // below every key of that value
    bound = ck > bound ? ck : bound;
    count = min(count, __popc(__ballot_sync(0xffffffffu, valid && hv >= cvv)));
  }
  *bound_io = bound;
  return count;
}

// Lanes 1..30 of a warp own 4 columns each (a 120-column strip); lanes 0 and 31 load the 4 columns on either side and
// only feed their neighbours through shuffles, so no lane needs extra halo loads.  Rows are software-pipelined five at a
// time (the next five float4 loads are in flight while the current five are processed).
__global__ void __launch_bounds__(256, 2) peak_scan_generic_kernel(const float* __restrict__ q, int H, int W, float thr, int nwarps,
                                                                int BAND, int tsel, uint8_t* __restrict__ ws) {
  extern __shared__ unsigned long long s_lists[];  // [warps][CAPW + TSEL]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= nwarps) return;
  const int band = blockIdx.x, b = blockIdx.y, nbands = gridDim.x;
  unsigned long long* list = s_lists + warp * (CAPW + TSEL);
  unsigned long long* top = list + CAPW;
  const float* img = q + (long long)b * H * W;
  const int c0 = warp * STRIP + (lane - 1) * 4;  // may be negative (lane 0 of warp 0) or >= W
  const bool owner = lane >= 1 && lane <= 30;
  const int y0 = band * BAND, y1 = min(y0 + BAND, H);
  const bool vec = (W % 4 == 0) && c0 >= 0 && c0 + 3 < W && (reinterpret_cast<uintptr_t>(q) & 15) == 0;
  const float NEG = -INFINITY;

  auto load_row = [&](int r) -> float4 {
    float4 v = make_float4(NEG, NEG, NEG, NEG);
    if (r >= 0 && r < H && r < y1 + 2) {
      const float* rowp = img + (long long)r * W;
      if (vec) v = __ldg(reinterpret_cast<const float4*>(rowp + c0));
      else {
        if (c0 >= 0 && c0 < W) v.x = __ldg(rowp + c0);
        if (c0 + 1 >= 0 && c0 + 1 < W) v.y = __ldg(rowp + c0 + 1);
        if (c0 + 2 >= 0 && c0 + 2 < W) v.z = __ldg(rowp + c0 + 2);
        if (c0 + 3 >= 0 && c0 + 3 < W) v.w = __ldg(rowp + c0 + 3);
      }
    }
    return v;
  };

  float4 raw[5], hm[5], cur[5];
#pragma unroll
  for (int u = 0; u < 5; ++u) { raw[u] = make_float4(NEG, NEG, NEG, NEG); hm[u] = raw[u]; cur[u] = load_row(y0 - 2 + u); }
  int count = 0, truncated = 0;
  const float p0 = __ldg(img);
  int nonconst = 0;

  for (int rbase = y0 - 2; rbase < y1 + 2; rbase += 5) {
    float4 nxt[5];
#pragma unroll
    for (int u = 0; u < 5; ++u) nxt[u] = load_row(rbase + 5 + u);
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int r = rbase + u;
      if (r >= y1 + 2) break;  // warp-uniform
      const float4 v = cur[u];
      const float l2 = __shfl_up_sync(0xffffffffu, v.z, 1), l1 = __shfl_up_sync(0xffffffffu, v.w, 1);
      const float r1 = __shfl_down_sync(0xffffffffu, v.x, 1), r2 = __shfl_down_sync(0xffffffffu, v.y, 1);
      if (owner && r >= y0 && r < y1) {  // strip min/max over owned pixels (trivial-image test)
        if (c0 < W) nonconst |= v.x != p0;
        if (c0 + 1 < W) nonconst |= v.y != p0;
        if (c0 + 2 < W) nonconst |= v.z != p0;
        if (c0 + 3 < W) nonconst |= v.w != p0;
      }
      const float mxyz = max3(v.x, v.y, v.z), myzw = max3(v.y, v.z, v.w);
      raw[u] = v;
      hm[u] = make_float4(max3(mxyz, l2, l1), max3(mxyz, l1, v.w), max3(myzw, v.x, r1), max3(myzw, r1, r2));
      // centre row r-2 now has its five horizontal maxima (rows r-4..r) in the ring
      const int rc = r - 2;
      if (rc >= y0 && rc < y1 && rc >= 2 && rc < H - 2) {
        const float4 ctr = raw[(u + 3) % 5];
        float4 vm;
        vm.x = max3(max3(hm[0].x, hm[1].x, hm[2].x), hm[3].x, hm[4].x);
        vm.y = max3(max3(hm[0].y, hm[1].y, hm[2].y), hm[3].y, hm[4].y);
        vm.z = max3(max3(hm[0].z, hm[1].z, hm[2].z), hm[3].z, hm[4].z);
        vm.w = max3(max3(hm[0].w, hm[1].w, hm[2].w), hm[3].w, hm[4].w);
        const bool k0 = owner && ctr.x == vm.x && ctr.x > thr && c0 >= 2 && c0 < W - 2;
        const bool k1 = owner && ctr.y == vm.y && ctr.y > thr && c0 + 1 >= 2 && c0 + 1 < W - 2;
        const bool k2 = owner && ctr.z == vm.z && ctr.z > thr && c0 + 2 < W - 2;
        const bool k3 = owner && ctr.w == vm.w && ctr.w > thr && c0 + 3 < W - 2;
        if (__ballot_sync(0xffffffffu, k0 | k1 | k2 | k3)) {
          if (count + 128 > CAPW) {  // make room: keep the best tsel so far
            __syncwarp();
            warp_select_top(list, count, tsel, top, lane);
            count = tsel; truncated = 1;
          }
          const uint32_t lt = (1u << lane) - 1u;
          const bool kk[4] = {k0, k1, k2, k3};
          const float cv[4] = {ctr.x, ctr.y, ctr.z, ctr.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t m = __ballot_sync(0xffffffffu, kk[j]);
            if (kk[j]) list[count + __popc(m & lt)] = make_key(cv[j], rc * W + c0 + j);
            count += __popc(m);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 5; ++u) cur[u] = nxt[u];
  }
  __syncwarp();
  if (count > tsel) { warp_select_top(list, count, tsel, top, lane); count = tsel; truncated = 1; }
  nonconst = __any_sync(0xffffffffu, nonconst);
  uint8_t* seg = ws + ((long long)(b * nbands + band) * nwarps + warp) * SEG_BYTES;
  if (lane == 0) {
    SegHeader h;
    h.count = count; h.truncated = truncated;
    h.worst_kept = truncated ? list[tsel - 1] : 0ull;
    h.nonconst = nonconst; h.pad[0] = h.pad[1] = h.pad[2] = 0;
    *reinterpret_cast<SegHeader*>(seg) = h;
  }
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(seg + sizeof(SegHeader));
  for (int i = lane; i < count; i += 32) keys[i] = list[i];
}


// Fast path (W % 4 == 0): bulk-copy staged.  A CTA owns a full-width band of rows, which is one contiguous byte
// range of the map, so a dedicated producer warp streams it into a shared-memory ring with cp.async.bulk
// (SCAN_R rows per stage, SCAN_NST stages, mbarrier complete_tx / empty handshakes) — ~100 KB in flight per SM
// without spending registers on prefetch — and the consumer warps (one 120-column strip each) run the lean filter:
//  * Emitted centres lie in rows [2, H-2) x columns [2, W-2), so no emitted 5x5 window ever leaves the map: the scan
//    never visits out-of-range rows, and lanes whose columns are out of range read a clamped (valid) column.
//  * A band visits rows [lo-2, hi+2) for its centre rows [lo, hi): four warm-up rows fill the register ring, every
//    later row emits, so the steady-state row has no row-range test; the column tests are folded into per-lane
//    thresholds (+inf for columns that may not emit).
//  * Per lane-row (4 pixels): one LDS.128, four shuffles, six FMNMX3 for the horizontal maxima; a stage's five rows are
//    loaded and voted as a batch, so a stage in which nothing can emit (the common case) is one warp-uniform branch.
//    Only rows that passed the threshold vote pay the eight FMNMX3 of the vertical maxima and the eight compares.
//  * The "trivial image" rule (A.1 step 2) costs nothing unless the map's four probe pixels are equal (TRACK).
//  * Density independence: the value cut (see scan_rows) is updated after every stage that appended, from the warp's
//    sorted distinct candidate values, and the emission thresholds follow it; on iid-uniform maps, where 4 % of all
//    pixels are 5x5 maxima above the threshold, the appends decay like (K + 1) / rows seen (23 % of the rows reach
//    the vertical maxima, 53 % of the stages take the slow branch).  Ties are bounded by the key-level cut of the
//    selections (tsel best keys, every SEL_SLACK appends).
//  * Code size matters here: with the append / selection logic inlined into each of the five unrolled row bodies the
//    hot loop was 30 KB of SASS and a quarter of the issue slots went to stall_no_instruction; the per-stage
//    bookkeeping now exists once, after the unrolled rows, and the selections are out of line (12 KB).
constexpr int SCAN_R = 5;    // rows per stage (= ring length, so ring slots are compile-time)
#ifndef CROG_SCAN_NST
#define CROG_SCAN_NST 4
#endif
constexpr int SCAN_NST = CROG_SCAN_NST;  // stages
constexpr int SEL_SLACK = 64; // keys a warp appends after a selection before it selects again

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}

// Cold paths of the scan, out of line: the streaming loop stays small enough for the instruction caches (with the
// append logic inlined in each of the five unrolled row bodies the loop was 30 KB of SASS and a quarter of all
// issue slots were lost to no_instruction stalls).
__device__ __noinline__ int scan_make_room(unsigned long long* list, int count, int tsel, int kdist, int lane,
                                           unsigned long long* bound_io) {
  __syncwarp();
  return select_and_cut(list, count, tsel, kdist, lane, bound_io);
}

template <bool TRACK>
__device__ __forceinline__ void scan_rows(const float* stage0, uint32_t full0, uint32_t empty0, int W, int row0, int nrows, int y0,
                                          int y1, int ccol, int c0, bool owner, float tx_, float ty_, float tz_, float tw_, float p0,
                                          int lane, unsigned long long* list, int& count,
                                          int& truncated, int& nonconst, const int tsel, const int kdist,
                                          unsigned long long& bound) {
  const float NEG = -INFINITY;
  // `bound`: every key this warp has dropped so far is below it (0: nothing dropped).  Two sources:
  //  * a selection that kept the best tsel keys: its worst kept key;
  //  * the VALUE cut: two candidates closer than the minimum distance are 5x5 maxima inside each other's window, hence
  //    equal, so the greedy pass only ever rejects a candidate because of an accepted one of the SAME value, and every
  //    distinct value yields at least one accepted peak.  Once the warp holds kdist = K + 1 distinct values, no
  //    candidate below the kdist-th of them can be reached before K peaks are accepted: it is dropped unseen (and the
  //    bound (value, lowest key) never flags a map, because the pass ends above it).
  // The warp keeps its distinct candidate values sorted across its lanes (`topv`), so the value cut follows every
  // append; on iid maps the appends then die out like kdist / rows seen.
  unsigned long long cut = bound;
  // per-lane emission thresholds (+inf for columns that may not emit), raised to just below the cut's value: the
  // threshold-first votes skip every row without a candidate that can still matter, so a dense map (most pixels above
  // the user threshold) costs little more than a sparse one.
  float tx = tx_, ty = ty_, tz = tz_, tw = tw_;
  int next_sel = SEL_SLACK;
  float topv = NEG;   // lane i: the i-th largest distinct candidate value this warp has appended (-inf: none yet)
  float vcut = NEG;   // the kdist-th of them once there are that many
  const bool vcut_on = kdist <= 32;
  float4 raw[5], hm[5];
#pragma unroll
  for (int u = 0; u < 5; ++u) { raw[u] = make_float4(NEG, NEG, NEG, NEG); hm[u] = raw[u]; }
  const uint32_t lt = (1u << lane) - 1u;
  const int nst = (nrows + SCAN_R - 1) / SCAN_R;
  for (int k = 0; k < nst; ++k) {
    const int s = k % SCAN_NST;
    mbar_wait(full0 + 8 * s, (k / SCAN_NST) & 1);
    const float* srow = stage0 + (long long)s * SCAN_R * W + ccol;
    // The stage's five rows are handled as a batch: all loads, then the threshold votes of the five centre rows (rows
    // 5k-2 .. 5k+2: two from the previous stage, three new) back to back, and - the common case - one warp-uniform
    // branch for the whole stage when no centre pixel can emit; the horizontal maxima of the five new rows are then
    // independent instruction streams.  (Row at a time, every row paid its own load -> shuffle -> vote -> branch
    // latency chain.)
    const int nvalid = min(SCAN_R, nrows - k * SCAN_R);  // rows of this stage that exist (the rest: stale ring data)
    float4 v[SCAN_R];
#pragma unroll
    for (int u = 0; u < SCAN_R; ++u) v[u] = *reinterpret_cast<const float4*>(srow + u * W);
    if (TRACK) {
#pragma unroll
      for (int u = 0; u < SCAN_R; ++u) {
        const int r = row0 + k * SCAN_R + u;  // rows past the stage's last valid one lie at or beyond y1
        if (owner && r >= y0 && r < y1) nonconst |= (v[u].x != p0) | (v[u].y != p0) | (v[u].z != p0) | (v[u].w != p0);
      }
    }
    bool pass[SCAN_R];
    bool any_pass = false;
#pragma unroll
    for (int u = 0; u < SCAN_R; ++u) {
      const float4 c = u < 2 ? raw[u + 3] : v[u - 2];  // centre row of row u = row u - 2
      pass[u] = __any_sync(0xffffffffu, (c.x > tx) | (c.y > ty) | (c.z > tz) | (c.w > tw)) && k * SCAN_R + u >= 4 && u < nvalid;
      any_pass |= pass[u];
    }
    if (!any_pass) {  // warp-uniform: nothing can emit in this stage, only the ring moves on (branch-free, five-way ILP)
#pragma unroll
      for (int u = 0; u < SCAN_R; ++u) {
        const float l2 = __shfl_up_sync(0xffffffffu, v[u].z, 1), l1 = __shfl_up_sync(0xffffffffu, v[u].w, 1);
        const float r1 = __shfl_down_sync(0xffffffffu, v[u].x, 1), r2 = __shfl_down_sync(0xffffffffu, v[u].y, 1);
        const float mxyz = max3(v[u].x, v[u].y, v[u].z), myzw = max3(v[u].y, v[u].z, v[u].w);
        raw[u] = v[u];
        hm[u] = make_float4(max3(mxyz, l2, l1), max3(mxyz, l1, v[u].w), max3(myzw, v[u].x, r1), max3(myzw, r1, r2));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * s);  // this warp is done reading the stage
      continue;
    }
    const int count0 = count;
#pragma unroll
    for (int u = 0; u < SCAN_R; ++u) {
      {
        const float l2 = __shfl_up_sync(0xffffffffu, v[u].z, 1), l1 = __shfl_up_sync(0xffffffffu, v[u].w, 1);
        const float r1 = __shfl_down_sync(0xffffffffu, v[u].x, 1), r2 = __shfl_down_sync(0xffffffffu, v[u].y, 1);
        const float mxyz = max3(v[u].x, v[u].y, v[u].z), myzw = max3(v[u].y, v[u].z, v[u].w);
        raw[u] = v[u];
        hm[u] = make_float4(max3(mxyz, l2, l1), max3(mxyz, l1, v[u].w), max3(myzw, v[u].x, r1), max3(myzw, r1, r2));
      }
      if (!pass[u]) continue;  // warp-uniform (voted with the thresholds of the stage's start: a superset)
      const float4 ctr = raw[(u + 3) % 5];  // centre row = this row - 2
      const float vx = max3(max3(hm[0].x, hm[1].x, hm[2].x), hm[3].x, hm[4].x);
      const float vy = max3(max3(hm[0].y, hm[1].y, hm[2].y), hm[3].y, hm[4].y);
      const float vz = max3(max3(hm[0].z, hm[1].z, hm[2].z), hm[3].z, hm[4].z);
      const float vw = max3(max3(hm[0].w, hm[1].w, hm[2].w), hm[3].w, hm[4].w);
      const float cv[4] = {ctr.x, ctr.y, ctr.z, ctr.w};
      const bool kk[4] = {ctr.x > tx && ctr.x == vx, ctr.y > ty && ctr.y == vy, ctr.z > tz && ctr.z == vz, ctr.w > tw && ctr.w == vw};
      if (!__any_sync(0xffffffffu, kk[0] | kk[1] | kk[2] | kk[3])) continue;
      if (count + 128 > CAPW) {  // rare: a stage of a plateau map
        count = scan_make_room(list, count, tsel, kdist, lane, &bound);
        cut = bound;
      }
      // exact filter on the 64-bit keys (value, then row-major index): a candidate that ties the cut's value but comes
      // later in the map is below the cut, so a plateau at the top value stops appending after tsel entries
      const uint32_t chi = (uint32_t)(cut >> 32), clo = (uint32_t)cut;
      const uint32_t lo0 = 0xffffffffu - (uint32_t)((row0 + k * SCAN_R + u - 2) * W + c0);  // low key word of column c0 (column j: lo0 - j)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t hi = f2ord(cv[j]);
        const bool kp = kk[j] && (hi > chi || (hi == chi && lo0 - j > clo));
        const uint32_t m = __ballot_sync(0xffffffffu, kp);
        if (kp) list[count + __popc(m & lt)] = ((unsigned long long)hi << 32) | (unsigned long long)(lo0 - j);
        count += __popc(m);
      }
    }
    if (count != count0) {  // once per stage that appended: value cut, thresholds, selection (one copy of this code)
      __syncwarp();
      if (vcut_on) {
        for (int i = count0; i < count; ++i) {  // enter the new values into the warp's sorted list of distinct values
          const float nv = key_val(list[i]);
          const uint32_t gt = __ballot_sync(0xffffffffu, topv > nv);
          if (__any_sync(0xffffffffu, topv == nv)) continue;
          const float up = __shfl_up_sync(0xffffffffu, topv, 1);
          const int pos = __popc(gt);  // the lanes holding larger values are exactly lanes [0, pos)
          if (lane == pos) topv = nv;
          else if (lane > pos) topv = up;
        }
        const float vc = __shfl_sync(0xffffffffu, topv, kdist - 1);
        if (vc > vcut) {  // "value > prev(vc)" == "value >= vc": no float lies between the two
          vcut = vc;
          const float pc = float_prev(vc);
          tx = fmaxf(tx, pc); ty = fmaxf(ty, pc); tz = fmaxf(tz, pc); tw = fmaxf(tw, pc);
        }
      }
      if (count >= next_sel) {  // lists only grow long on tie-heavy maps: keep the best tsel keys, cut at the worst kept
        count = scan_make_room(list, count, tsel, kdist, lane, &bound);
        cut = bound;
        next_sel = count + SEL_SLACK;
        if (cut) {
          const float pc = float_prev(key_val(cut));
          tx = fmaxf(tx, pc); ty = fmaxf(ty, pc); tz = fmaxf(tz, pc); tw = fmaxf(tw, pc);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * s);  // this warp is done reading the stage
  }
  truncated = bound != 0ull;
  if (vcut > NEG) {  // everything the raised thresholds skipped lies below the value cut
    const unsigned long long ck = (unsigned long long)f2ord(vcut) << 32;
    bound = ck > bound ? ck : bound;
  }
}

// blockDim = (nwarps + 1) * 32: warps 0..nwarps-1 consume, warp nwarps produces.
template <int MINB>  // resident CTAs per SM the register allocation is held to (2: no spills, 3: more warps in flight)
__global__ void __launch_bounds__(288, MINB) peak_scan_kernel(const float* __restrict__ q, int H, int W, float thr, int nwarps, int BAND,
                                                           int tsel, int kdist, uint8_t* __restrict__ ws) {
  extern __shared__ __align__(128) uint8_t s_raw[];
  float* stage0 = reinterpret_cast<float*>(s_raw);                                            // [SCAN_NST][SCAN_R][W]
  unsigned long long* s_lists = reinterpret_cast<unsigned long long*>(s_raw + (size_t)SCAN_NST * SCAN_R * W * 4);  // [warps][CAPW + TSEL]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lists + (size_t)nwarps * (CAPW + TSEL));  // full[NST], empty[NST]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int band = blockIdx.x, b = blockIdx.y, nbands = gridDim.x;
  const float* img = q + (long long)b * H * W;
  const int y0 = band * BAND, y1 = min(y0 + BAND, H);
  const int lo = max(y0, 2), hi = min(y1, H - 2);  // centre rows this band emits
  const float p0 = __ldg(img);
  const bool track = p0 == __ldg(img + 1) && p0 == __ldg(img + (H > 1 ? W : 0)) && p0 == __ldg(img + (long long)H * W - 1);  // CTA-uniform
  // rows to visit: [lo-2, hi+2) when the band emits (those rows all exist and include the band's own rows at the
  // top / bottom edge for the constant-map test); otherwise just the band's rows, and only when tracking
  const bool emits = hi > lo;
  const int row0 = emits ? lo - 2 : y0;
  const int nrows = emits ? hi - lo + 4 : (track ? y1 - y0 : 0);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + SCAN_NST);
  if (threadIdx.x == 0) {
    for (int s = 0; s < SCAN_NST; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, nwarps); }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == nwarps) {  // ---- producer
    if (lane == 0) {
      const int nst = (nrows + SCAN_R - 1) / SCAN_R;
      for (int k = 0; k < nst; ++k) {
        const int s = k % SCAN_NST;
        if (k >= SCAN_NST) {  // parked by the hardware until the consumers free the stage: no issue slots spent polling
          while (!mbar_try_wait_hint(empty0 + 8 * s, ((k / SCAN_NST) - 1) & 1, 100000u)) {}
        }
        const int rows = min(SCAN_R, nrows - k * SCAN_R);
        const uint32_t bytes = (uint32_t)rows * W * 4;
        mbar_expect_tx(full0 + 8 * s, bytes);
        bulk_load_1d(smem_u32(stage0 + (long long)s * SCAN_R * W), img + (long long)(row0 + k * SCAN_R) * W, bytes, full0 + 8 * s);
      }
    }
    return;
  }
  // ---- consumers
  unsigned long long* list = s_lists + warp * (CAPW + TSEL);
  const int c0 = warp * STRIP + (lane - 1) * 4;
  const bool inrange = c0 >= 0 && c0 + 3 < W;  // W % 4 == 0: a lane's four columns are all in or all out
  const bool owner = lane >= 1 && lane <= 30 && inrange;
  // candidate columns must lie in [2, W-2): columns that may not emit get an unreachable threshold
  const float INF = INFINITY;
  const float tx = (owner && c0 >= 2 && emits) ? thr : INF, ty = (owner && c0 + 1 >= 2 && emits) ? thr : INF;
  const float tz = (owner && c0 + 2 < W - 2 && emits) ? thr : INF, tw = (owner && c0 + 3 < W - 2 && emits) ? thr : INF;
  const int ccol = min(max(c0, 0), W - 4);
  int nonconst = track ? 0 : 1;
  int count = 0, truncated = 0;
  unsigned long long bound = 0ull;
  if (track) scan_rows<true>(stage0, full0, empty0, W, row0, nrows, y0, y1, ccol, c0, owner, tx, ty, tz, tw, p0, lane, list, count, truncated, nonconst, tsel, kdist, bound);
  else scan_rows<false>(stage0, full0, empty0, W, row0, nrows, y0, y1, ccol, c0, owner, tx, ty, tz, tw, p0, lane, list, count, truncated, nonconst, tsel, kdist, bound);
  __syncwarp();
  // one closing selection: best tsel keys, then the value cut (drops what was appended while the cut was still lower)
  if (count > min(tsel, kdist)) count = select_and_cut(list, count, tsel, kdist, lane, &bound);
  truncated = bound != 0ull;
  nonconst = __any_sync(0xffffffffu, nonconst);
  uint8_t* seg = ws + ((long long)(b * nbands + band) * nwarps + warp) * SEG_BYTES;
  if (lane == 0) {
    SegHeader h;
    h.count = count; h.truncated = truncated;
    h.worst_kept = bound;
    h.nonconst = nonconst; h.pad[0] = h.pad[1] = h.pad[2] = 0;
    *reinterpret_cast<SegHeader*>(seg) = h;
  }
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(seg + sizeof(SegHeader));
  for (int i = lane; i < count; i += 32) keys[i] = list[i];
}

// float32(atan2(double, double)) * 0.5f, then the NumPy-1.24 float64 promotions (App. A.3)
__device__ __forceinline__ void decode_grasp(const float* s, const float* c, const float* w, long long o, int row, int col,
                                             double* g) {
  const float ang = __fmul_rn((float)atan2((double)s[o], (double)c[o]), 0.5f);
  g[0] = (double)col; g[1] = (double)row;
  g[2] = (double)w[o] * 100.0;
  g[3] = 20.0;
  g[4] = (double)ang / 3.141592653589793 * 180.0;
}

__global__ void peak_select_kernel(const float* __restrict__ sin_m, const float* __restrict__ cos_m,
                                   const float* __restrict__ wid, int B, int H, int W, int K, int nseg,
                                   const uint8_t* __restrict__ ws, int* __restrict__ flags, int* __restrict__ peaks,
                                   int* __restrict__ n_peaks, double* __restrict__ grasps) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const uint8_t* base = ws + (long long)b * nseg * SEG_BYTES;
  // bound D: best "worst kept" key over truncated segments; min/max for the trivial-image rule
  unsigned long long D = 0ull;
  int nonconst = 0;
  for (int s = lane; s < nseg; s += 32) {
    const SegHeader* h = reinterpret_cast<const SegHeader*>(base + (long long)s * SEG_BYTES);
    if (h->truncated && h->worst_kept > D) D = h->worst_kept;
    nonconst |= h->nonconst;
  }
  D = warp_max_u64(D);
  const bool trivial = !__any_sync(0xffffffffu, nonconst);
  int acc_r[MAXK], acc_c[MAXK];
  int nacc = 0, flag = 0;
  unsigned long long prev = ~0ull;
  while (!trivial && nacc < K) {
    unsigned long long best = 0ull;
    for (int e = lane; e < nseg * TSEL; e += 32) {
      const int s = e / TSEL, i = e % TSEL;
      const SegHeader* h = reinterpret_cast<const SegHeader*>(base + (long long)s * SEG_BYTES);
      if (i < h->count) {
        const unsigned long long k = reinterpret_cast<const unsigned long long*>(base + (long long)s * SEG_BYTES + sizeof(SegHeader))[i];
        if (k < prev && k > best) best = k;
      }
    }
    best = warp_max_u64(best);
    if (best == 0ull) { if (D != 0ull) flag = 1; break; }  // exhausted; dropped candidates may remain
    if (best < D) { flag = 1; break; }                      // a dropped candidate could precede this one
    prev = best;
    const int idx = key_idx(best), r = idx / W, c = idx % W;
    bool ok = true;
    for (int j = 0; j < nacc; ++j) ok = ok && (max(abs(acc_r[j] - r), abs(acc_c[j] - c)) >= 2);
    if (ok) { acc_r[nacc] = r; acc_c[nacc] = c; ++nacc; }
  }
  if (lane == 0) {
    flags[b] = flag;
    n_peaks[b] = nacc;
    const long long plane = (long long)b * H * W;
    for (int j = 0; j < K; ++j) {
      double* g = grasps + ((long long)b * K + j) * 5;
      if (j < nacc) {
        peaks[((long long)b * K + j) * 2] = acc_r[j]; peaks[((long long)b * K + j) * 2 + 1] = acc_c[j];
        decode_grasp(sin_m + plane, cos_m + plane, wid + plane, (long long)acc_r[j] * W + acc_c[j], acc_r[j], acc_c[j], g);
      } else {
        peaks[((long long)b * K + j) * 2] = -1; peaks[((long long)b * K + j) * 2 + 1] = -1;
        g[0] = g[1] = g[2] = g[3] = g[4] = 0.0;
      }
    }
  }
}

// Exact fallback: one CTA per flagged map, K sweeps of "best remaining candidate".
__global__ void __launch_bounds__(256) peak_exact_kernel(const float* __restrict__ q, const float* __restrict__ sin_m,
                                                         const float* __restrict__ cos_m, const float* __restrict__ wid,
                                                         int H, int W, int K, float thr, const int* __restrict__ flags,
                                                         int* __restrict__ peaks, int* __restrict__ n_peaks,
                                                         double* __restrict__ grasps) {
  const int b = blockIdx.x;
  if (!flags[b]) return;
  __shared__ unsigned long long s_red[8];
  __shared__ int s_r[MAXK], s_c[MAXK];
  __shared__ int s_n;
  const float* img = q + (long long)b * H * W;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  unsigned long long prev = ~0ull;
  for (int sweep = 0; sweep < K * 9 + 1; ++sweep) {  // every acceptance rejects at most 8 others
    const int nacc = s_n;
    if (nacc >= K) break;
    unsigned long long best = 0ull;
    for (int p = threadIdx.x; p < H * W; p += blockDim.x) {
      const int r = p / W, c = p % W;
      if (r < 2 || r >= H - 2 || c < 2 || c >= W - 2) continue;
      const float v = img[p];
      if (!(v > thr)) continue;
      const unsigned long long k = make_key(v, p);
      if (k >= prev || k <= best) continue;
      bool ismax = true;
      for (int dy = -2; dy <= 2 && ismax; ++dy)
        for (int dx = -2; dx <= 2; ++dx)
          if (img[(r + dy) * W + c + dx] > v) { ismax = false; break; }
      if (ismax) best = k;
    }
    best = warp_max_u64(best);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
    __syncthreads();
    best = 0ull;
    for (int i = 0; i < (blockDim.x >> 5); ++i) best = s_red[i] > best ? s_red[i] : best;
    __syncthreads();
    if (best == 0ull) break;
    prev = best;
    if (threadIdx.x == 0) {
      const int idx = key_idx(best), r = idx / W, c = idx % W;
      bool ok = true;
      for (int j = 0; j < nacc; ++j) ok = ok && (max(abs(s_r[j] - r), abs(s_c[j] - c)) >= 2);
      if (ok) { s_r[nacc] = r; s_c[nacc] = c; s_n = nacc + 1; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int nacc = s_n;
    n_peaks[b] = nacc;
    const long long plane = (long long)b * H * W;
    for (int j = 0; j < K; ++j) {
      double* g = grasps + ((long long)b * K + j) * 5;
      if (j < nacc) {
        peaks[((long long)b * K + j) * 2] = s_r[j]; peaks[((long long)b * K + j) * 2 + 1] = s_c[j];
        decode_grasp(sin_m + plane, cos_m + plane, wid + plane, (long long)s_r[j] * W + s_c[j], s_r[j], s_c[j], g);
      } else {
        peaks[((long long)b * K + j) * 2] = -1; peaks[((long long)b * K + j) * 2 + 1] = -1;
        g[0] = g[1] = g[2] = g[3] = g[4] = 0.0;
      }
    }
  }
}

__global__ void angle_map_kernel(const float* __restrict__ s, const float* __restrict__ c, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __fmul_rn((float)atan2((double)s[i], (double)c[i]), 0.5f);
}

// ------------------------------------------------------------------ Jaccard
#ifndef CROG_JT
#define CROG_JT 128
#endif
constexpr int JT = CROG_JT;  // threads per sample CTA (>= MAXK, multiple of 32)
#ifndef JACCARD_CTAS_PER_SM
#define JACCARD_CTAS_PER_SM 8  // latency / barrier bound kernel: more resident samples per SM (64 registers, 24.5 KB smem each)
#endif

__device__ __forceinline__ int block_sum(int v, int* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
  for (int i = 0; i < (JT >> 5); ++i) t += s_red[i];
  return t;
}

#ifdef CROG_JAC_NOINLINE_RECT
__device__ __noinline__ void make_rect_nl(const double* rect5, TgRect* R) { tg_make_rect(rect5, R); }
#else
#define make_rect_nl tg_make_rect
#endif
#ifdef CROG_JAC_NOINLINE_SLOW
#define JAC_SLOW_ATTR __noinline__
#else
#define JAC_SLOW_ATTR
#endif

// pixel count of one rectangle / of the intersection of two, exact for any size (slow path)
__device__ JAC_SLOW_ATTR int slow_count(const TgRect* A, const TgRect* Bq, int* s_red) {
  const int x0 = Bq ? max(A->x0, Bq->x0) : A->x0, x1 = Bq ? min(A->x1, Bq->x1) : A->x1;
  const int y0 = Bq ? max(A->y0, Bq->y0) : A->y0, y1 = Bq ? min(A->y1, Bq->y1) : A->y1;
  int cnt = 0;
  if (x0 <= x1 && y0 <= y1) {
    const int wdt = y1 - y0 + 1, tot = (x1 - x0 + 1) * wdt;
    for (int p = threadIdx.x; p < tot; p += JT) {
      const int X = x0 + p / wdt, Y = y0 + p % wdt;
      if (tg_point_painted(A, X, Y) && (!Bq || tg_point_painted(Bq, X, Y))) ++cnt;
    }
  }
  return block_sum(cnt, s_red);
}

// One CTA per sample.  Predicted rectangles are rasterised once into shared-memory row masks; then every warp takes
// ground-truth rectangles round-robin (no CTA-wide barrier inside the loop): lane = scanline, row mask by exact integer
// scanline arithmetic, popc(A & B) against the predictions that pass the angle gate, warp-shuffle reductions.
__global__ void __launch_bounds__(JT, JACCARD_CTAS_PER_SM) jaccard_kernel(const double* __restrict__ grasps, const int* __restrict__ n_peaks, int K,
                                                     double* __restrict__ gt, const int* __restrict__ gt_count, int Mmax,
                                                     int* __restrict__ inter_out, int* __restrict__ uni_out,
                                                     int* __restrict__ j_flags, long long* __restrict__ counters, int edit_gt) {
  extern __shared__ uint32_t s_mask[];  // [K][TG_MAXROWS][TG_WORDS]
  __shared__ TgRect s_pred[MAXK];
  __shared__ int s_parea[MAXK];
  __shared__ TgRect s_gt;
  __shared__ TgRect s_gr[JT];      // phase-1 results: one ground-truth rectangle per thread
  __shared__ uint32_t s_pass[JT];
  __shared__ int s_list[JT];
  __shared__ int s_nlist;
  __shared__ int s_red[JT / 32];
  __shared__ int s_wcnt[(JT / 32) * MAXK];
  __shared__ int s_j1, s_jk, s_slow;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = min(n_peaks ? n_peaks[b] : K, K);
  const int M = min(gt_count[b], Mmax);
  double* G = gt + (long long)b * Mmax * 6;
  const double* P = grasps + (long long)b * K * 5;
  // grasp_eval.py:367-368, in place
  for (int m = tid; edit_gt && m < M; m += JT) {
    G[m * 6 + 3] = 20.0;
    const double w = G[m * 6 + 2];
    G[m * 6 + 2] = w < 0.0 ? 0.0 : (w > 100.0 ? 100.0 : w);  // NaN stays NaN like np.clip
  }
  if (n == 0 || M == 0) {  // nothing can match (block-uniform): J = 0, counters still count the sample
    if (inter_out)
      for (int i = tid; i < K * Mmax; i += JT) { inter_out[(long long)b * K * Mmax + i] = 0; uni_out[(long long)b * K * Mmax + i] = 0; }
    if (tid == 0) {
      j_flags[b * 2] = 0; j_flags[b * 2 + 1] = 0;
      if (counters) {
        atomicAdd(reinterpret_cast<unsigned long long*>(counters) + 1, 1ull);
        atomicAdd(reinterpret_cast<unsigned long long*>(counters) + 3, 1ull);
      }
    }
    return;
  }
  if (tid == 0) { s_j1 = 0; s_jk = 0; s_slow = 0; }
  if (tid < n) make_rect_nl(P + tid * 5, &s_pred[tid]);
  __syncthreads();
  int all_fast = 1;
  for (int k = 0; k < n; ++k) all_fast &= s_pred[k].fast;
  // predicted rectangles: row masks + areas.  Thread = scanline of every fast rectangle in turn (no barrier between
  // rectangles: per-warp partial areas go through shared memory once); oversized ones take the CTA-wide slow count.
  for (int k = 0; k < n; ++k) {
    const TgRect* R = &s_pred[k];
    if (!R->fast) continue;
    int cnt = 0;
    for (int r = tid; r < TG_MAXROWS; r += JT) {
      const int X = R->x0 + r;
      uint32_t* row = s_mask + ((long long)k * TG_MAXROWS + r) * TG_WORDS;
      if (X <= R->x1) {
        tg_row_mask(R, X, row);
#pragma unroll
        for (int w = 0; w < TG_WORDS; ++w) cnt += __popc(row[w]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) s_wcnt[warp * MAXK + k] = cnt;
  }
  __syncthreads();
  if (tid < n && s_pred[tid].fast) {
    int t = 0;
    for (int w = 0; w < JT / 32; ++w) t += s_wcnt[w * MAXK + tid];
    s_parea[tid] = t;
  }
  for (int k = 0; k < n && !all_fast; ++k) {  // block-uniform
    if (s_pred[k].fast) continue;
    const int cnt = slow_count(&s_pred[k], nullptr, s_red);
    if (tid == 0) s_parea[k] = cnt;
  }
  __syncthreads();
  if (inter_out) {
    for (int i = tid; i < K * Mmax; i += JT) {
      inter_out[(long long)b * K * Mmax + i] = 0; uni_out[(long long)b * K * Mmax + i] = 0;
    }
    __syncthreads();
  }
  if (all_fast) {
    // ---- phase 1 (thread per GT): angle gate, corner arithmetic (float64 trig) and bounding-box cull run once per
    // ground-truth rectangle, in parallel across the CTA; survivors are compacted into a work list.
    // ---- phase 2 (warp per surviving GT, no CTA-wide barrier inside): lane = scanline, row mask by exact integer
    // scanline arithmetic, popc(A & B) against the predictions that passed, warp-shuffle reductions.
    for (int m0 = 0; m0 < M; m0 += JT) {
      if (tid == 0) s_nlist = 0;
      __syncthreads();
      const int mt = m0 + tid;
      if (mt < M) {
        const double* g = G + mt * 6;
        uint32_t pass = 0;
        for (int k = 0; k < n; ++k) {
          const double tp = P[k * 5 + 4], tg = g[4];
          if (!(fabs(tp - tg) > 30.0 && fabs(tp + tg) > 30.0)) pass |= 1u << k;
        }
        if (pass) {
          TgRect Gr;
          make_rect_nl(g, &Gr);
          if (!Gr.fast) {
            s_slow = 1;  // oversized GT (only possible with edit_gt == 0): handled by the CTA-synchronous path below
          } else {
            if (!inter_out) {  // J only needs pairs that can intersect: drop gated predictions whose boxes miss this GT
              uint32_t keep = 0;
              for (int k = 0; k < n; ++k) if ((pass >> k) & 1u) {
                const TgRect* R = &s_pred[k];
                if (!(R->x1 < Gr.x0 || R->x0 > Gr.x1 || R->y1 < Gr.y0 || R->y0 > Gr.y1)) keep |= 1u << k;
              }
              pass = keep;
            }
            if (pass) {
              s_gr[tid] = Gr;
              s_pass[tid] = pass;
              s_list[atomicAdd(&s_nlist, 1)] = tid;
            }
          }
        }
      }
      __syncthreads();
      const int nl = s_nlist;
      for (int li = warp; li < nl; li += JT / 32) {
        const int t = s_list[li], m = m0 + t;
        const TgRect Gr = s_gr[t];
        const uint32_t pass = s_pass[t];
        int garea = 0;
        int inter[MAXK];
#pragma unroll 1
        for (int k = 0; k < n; ++k) inter[k] = 0;
        for (int X = Gr.x0 + lane; X <= Gr.x1; X += 32) {
          uint32_t grow[TG_WORDS];
          tg_row_mask(&Gr, X, grow);
#pragma unroll
          for (int w = 0; w < TG_WORDS; ++w) garea += __popc(grow[w]);
          for (int k = 0; k < n; ++k) {
            if (!((pass >> k) & 1u)) continue;
            const TgRect* R = &s_pred[k];
            if (X < R->x0 || X > R->x1) continue;
            const uint32_t* prow = s_mask + ((long long)k * TG_MAXROWS + (X - R->x0)) * TG_WORDS;
            const int dw = Gr.yw0 - R->yw0;
            int c = 0;
#pragma unroll
            for (int w = 0; w < TG_WORDS; ++w) {
              const int pw = w + dw;
              if (pw >= 0 && pw < TG_WORDS) c += __popc(grow[w] & prow[pw]);
            }
            inter[k] += c;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) garea += __shfl_xor_sync(0xffffffffu, garea, o);
        for (int k = 0; k < n; ++k) {
          if (!((pass >> k) & 1u)) continue;
          int it = inter[k];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) it += __shfl_xor_sync(0xffffffffu, it, o);
          const int uni = s_parea[k] + garea - it;
          if (lane == 0) {
            if (inter_out) { inter_out[((long long)b * K + k) * Mmax + m] = it; uni_out[((long long)b * K + k) * Mmax + m] = uni; }
            if (uni > 0 && 4LL * it > (long long)uni) { atomicOr(&s_jk, 1); if (k == 0) atomicOr(&s_j1, 1); }
          }
        }
      }
      __syncthreads();
    }
  }
  // ---- CTA-synchronous path: samples with oversized predictions, and oversized GT rectangles of any sample
  const bool need_slow = !all_fast || s_slow;  // block-uniform (read after the barrier above)
  for (int m = 0; need_slow && m < M; ++m) {
    const double* g = G + m * 6;
    uint32_t pass = 0;
    for (int k = 0; k < n; ++k) {
      const double tp = P[k * 5 + 4], tg = g[4];
      if (!(fabs(tp - tg) > 30.0 && fabs(tp + tg) > 30.0)) pass |= 1u << k;
    }
    if (!pass) continue;  // block-uniform
    __syncthreads();
    if (tid == 0) make_rect_nl(g, &s_gt);
    __syncthreads();
    const TgRect* Gr = &s_gt;
    if (all_fast && Gr->fast) continue;  // already handled by the warp path (block-uniform)
    const int garea = slow_count(Gr, nullptr, s_red);
    for (int k = 0; k < n; ++k) {
      if (!((pass >> k) & 1u)) continue;
      const int it = slow_count(&s_pred[k], Gr, s_red);
      const int uni = s_parea[k] + garea - it;
      if (tid == 0) {
        if (inter_out) { inter_out[((long long)b * K + k) * Mmax + m] = it; uni_out[((long long)b * K + k) * Mmax + m] = uni; }
        if (uni > 0 && 4LL * it > (long long)uni) { s_jk = 1; if (k == 0) s_j1 = 1; }
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    j_flags[b * 2] = s_j1; j_flags[b * 2 + 1] = s_jk;
    if (counters) {
      atomicAdd(reinterpret_cast<unsigned long long*>(counters) + 0, (unsigned long long)s_j1);
      atomicAdd(reinterpret_cast<unsigned long long*>(counters) + 1, 1ull);
      atomicAdd(reinterpret_cast<unsigned long long*>(counters) + 2, (unsigned long long)s_jk);
      atomicAdd(reinterpret_cast<unsigned long long*>(counters) + 3, 1ull);
    }
  }
}

// Band height: tall bands when there are enough maps to fill the GPU anyway (at least ~4 CTAs per SM), short ones otherwise.
inline void scan_geometry(int B, int H, int W, int* nbands, int* nwarps, int* band) {
  *nwarps = (W + STRIP - 1) / STRIP;
  const long long ctas_large = (long long)B * ((H + BAND_LARGE - 1) / BAND_LARGE);
  *band = ctas_large >= 4 * 148 ? BAND_LARGE : BAND_SMALL;
  *nbands = (H + *band - 1) / *band;
}

}  // namespace

extern "C" int64_t crog_detect_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t K) {
  int nb, nw, band;
  scan_geometry(B, H, W, &nb, &nw, &band);
  (void)K;
  return (int64_t)B * nb * nw * SEG_BYTES + (int64_t)B * 4 + 256;
}

extern "C" int crog_detect_grasps(const float* q, const float* sin_m, const float* cos_m, const float* wid, int32_t B, int32_t H,
                                  int32_t W, int32_t K, float threshold, int32_t* peaks, int32_t* n_peaks, double* grasps,
                                  void* workspace, void* stream) {
  CROG_REQUIRE(K >= 1 && K <= MAXK, CROG_E_BADSHAPE, "detect_grasps: 1 <= num_grasps <= %d", MAXK);
  CROG_REQUIRE(H >= 1 && W >= 1 && W <= 8 * STRIP && (long long)H * W < (1LL << 31), CROG_E_BADSHAPE, "detect_grasps: map %dx%d unsupported", H, W);
  CROG_REQUIRE(B <= 65535, CROG_E_BADSHAPE, "detect_grasps: at most 65535 maps per call");
  CROG_REQUIRE(aligned16(workspace), CROG_E_BADALIGN, "detect_grasps: the workspace must be 16B aligned");
  if (B == 0) return CROG_OK;
  int nb, nw, band;
  scan_geometry(B, H, W, &nb, &nw, &band);
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  int* flags = (int*)(ws + (((int64_t)B * nb * nw * SEG_BYTES + 15) / 16) * 16);
  const size_t smem_lists = (size_t)nw * (CAPW + TSEL) * 8;
  const size_t smem_fast = (size_t)SCAN_NST * SCAN_R * W * 4 + smem_lists + 2 * SCAN_NST * 8;
  static DeviceOnce once;
  int dev_;
  if (once.need(&dev_)) {
    CROG_CUDA_OK(cudaFuncSetAttribute(peak_scan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SCAN_NST * SCAN_R * 8 * STRIP * 4 + 8 * (CAPW + TSEL) * 8 + 2 * SCAN_NST * 8));
    CROG_CUDA_OK(cudaFuncSetAttribute(peak_scan_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SCAN_NST * SCAN_R * 8 * STRIP * 4 + 8 * (CAPW + TSEL) * 8 + 2 * SCAN_NST * 8));
    CROG_CUDA_OK(cudaFuncSetAttribute(peak_scan_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (CAPW + TSEL) * 8));
    once.done(dev_);
  }
  // keys kept per warp segment.  Strict 5x5 maxima are >= 3 apart, so without ties the answer is the global top K; ties
  // (plateaus, e.g. a quality map clipped at 1.0) make the greedy pass reject up to 8 candidates per accepted peak, and a
  // map whose segments kept too few keys to prove the result is flagged for the slow exact sweep.  32 keeps that rare
  // (measured: 2K + 2 = 12 keys flagged a quarter of the 'blobs' maps - 4.2 ms of exact sweeps per 4096 maps).
  static const int tsel_env = getenv("CROG_SCAN_TSEL") ? atoi(getenv("CROG_SCAN_TSEL")) : 0;
  const int tsel = tsel_env > 0 ? max(1, min(tsel_env, TSEL)) : TSEL;
  static const int occ_env = getenv("CROG_SCAN_OCC") ? atoi(getenv("CROG_SCAN_OCC")) : 0;
  // the bulk-copy staged scan needs 16-byte aligned rows (W % 4 == 0 and an aligned base); anything else - odd widths, a
  // plane of a larger tensor that starts at an odd offset - takes the generic kernel
  if (W % 4 == 0 && aligned16(q) && (long long)H * W >= 2 && !getenv("CROG_SCAN_GENERIC")) {
    if (occ_env != 3) peak_scan_kernel<2><<<dim3(nb, B), (nw + 1) * 32, smem_fast, s>>>(q, H, W, threshold, nw, band, tsel, K + 1, ws);
    else peak_scan_kernel<3><<<dim3(nb, B), (nw + 1) * 32, smem_fast, s>>>(q, H, W, threshold, nw, band, tsel, K + 1, ws);
  }
  else peak_scan_generic_kernel<<<dim3(nb, B), nw * 32, smem_lists, s>>>(q, H, W, threshold, nw, band, tsel, ws);
  CROG_LAUNCH_OK("peak_scan");
  peak_select_kernel<<<(B + 3) / 4, 128, 0, s>>>(sin_m, cos_m, wid, B, H, W, K, nb * nw, ws, flags, peaks, n_peaks, grasps);
  CROG_LAUNCH_OK("peak_select");
  peak_exact_kernel<<<B, 256, 0, s>>>(q, sin_m, cos_m, wid, H, W, K, threshold, flags, peaks, n_peaks, grasps);
  CROG_LAUNCH_OK("peak_exact");
  return CROG_OK;
}

extern "C" int crog_angle_map(const float* sin_m, const float* cos_m, float* out, int64_t n, void* stream) {
  if (n == 0) return CROG_OK;
  long long g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  angle_map_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(sin_m, cos_m, out, n);
  CROG_LAUNCH_OK("angle_map");
  return CROG_OK;
}

extern "C" int crog_jaccard(const double* grasps, const int32_t* n_peaks, int32_t K, double* gt, const int32_t* gt_count,
                            int32_t Mmax, int32_t B, int32_t* inter, int32_t* uni, int32_t* j_flags, int64_t* counters,
                            int32_t edit_gt, void* stream) {
  CROG_REQUIRE(K >= 1 && K <= MAXK, CROG_E_BADSHAPE, "jaccard: 1 <= K <= %d", MAXK);
  CROG_REQUIRE((inter == nullptr) == (uni == nullptr), CROG_E_BADSHAPE, "jaccard: inter/uni must both be given or both NULL");
  if (B == 0) return CROG_OK;
  const size_t smem = (size_t)K * TG_MAXROWS * TG_WORDS * 4;
  static DeviceOnce once_j;
  int dev_j;
  if (once_j.need(&dev_j)) {
    CROG_CUDA_OK(cudaFuncSetAttribute(jaccard_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXK * TG_MAXROWS * TG_WORDS * 4));
    CROG_CUDA_OK(cudaFuncSetAttribute(jaccard_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    once_j.done(dev_j);
  }
  jaccard_kernel<<<B, JT, smem, (cudaStream_t)stream>>>(grasps, n_peaks, K, gt, gt_count, Mmax, inter, uni, j_flags,
                                                        (long long*)counters, edit_gt);
  CROG_LAUNCH_OK("jaccard");
  return CROG_OK;
}
