/* crog_b200 C ABI — hand-written sm_100a kernels for CROG's batched referring-grasp
 * inference path (model forward + engine glue + grasp decode / Jaccard tail).
 *
 * The reference (HilbertXu/CROG) has no FFI boundary: its "operator API" is the Python
 * nn.Module surface (model/crog.py:47, model/__init__.py:6) and utils/grasp_eval.py
 * (:289,:305,:350,:362), everything below being torch -> cuDNN/cuBLAS or numpy/skimage on
 * the host.  This header is what a replacement binds instead; each entry
 * point cites the reference code it replaces.  Python binds it with ctypes
 * (crog_b200/_lib.py); see INTEGRATION.md for the reference-side stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer borrowed for the call unless it says "host";
 *  - calls are asynchronous on `stream` (a cudaStream_t passed as void*), never allocate,
 *    never synchronise, and are re-entrant across streams / devices;
 *  - return 0 on success, a negative CROG_E_* code otherwise; crog_last_error() gives the
 *    thread-local message;
 *  - there is no CPU fallback and no pre-sm_100 path (CROG_E_UNSUPPORTED_ARCH).
 */
#ifndef CROG_B200_H_
#define CROG_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  CROG_OK = 0,
  CROG_E_BADSHAPE = -1,
  CROG_E_BADALIGN = -2,
  CROG_E_UNSUPPORTED_ARCH = -3,
  CROG_E_CUDA = -4,
};
enum { CROG_F32 = 0, CROG_BF16 = 1 };
enum { CROG_ACT_NONE = 0, CROG_ACT_RELU = 1, CROG_ACT_QUICKGELU = 2, CROG_ACT_TANH = 3 };
enum { CROG_IMPL_AUTO = 0, CROG_IMPL_SIMT = 1, CROG_IMPL_TCGEN05 = 2 };

const char* crog_last_error(void);
int crog_abi_version(void);
/* 0 if the current device is sm_100 (B200), CROG_E_UNSUPPORTED_ARCH otherwise. */
int crog_check_device(void);

/* ------------------------------------------------------------------ contractions
 * One descriptor covers every dense contraction of the forward: 1x1 / 3x3 convolutions
 * as implicit GEMM over NHWC activations and all linear layers.
 *   replaces: nn.Conv2d+BatchNorm2d+ReLU (model/clip.py:44-57,208-213; model/layers.py:8-11),
 *             nn.Linear / MHA in/out projections (model/clip.py:239-265; model/layers.py:280-339),
 *             the FPN text gate (model/layers.py:376-379) and CoordConv (model/layers.py:19-44).
 *
 * out[map(r), n] = epi( sum_{t<taps} sum_{c<cin} A[r + shift_t, c] * Wt[n, t*cin + c] )
 *
 * Activations are row-major [rows, ld] matrices of NHWC pixels.  A "padded" tensor stores
 * every sample as (H+2)x(W+2) pixels with a zero halo, so the 9 taps of a 3x3/pad-1
 * convolution are nine row-shifted views of the same matrix (shift = (ky-1)*(W+2)+(kx-1));
 * halo rows are never written by the epilogue and must have been zeroed once.
 * epi(v): v += addmat[sp(r), n]; v = v*scale[n] + bias[n]; v = act(v);
 *         if gate: v = relu((v * gate[b(r), n]) * scale2[n] + bias2[n]);
 *         v += residual[map(r), n]   (then act2 = relu if residual_relu)
 */
typedef struct CrogGemm {
  const void* a;          /* activations [a_rows, a_ld] */
  int64_t a_rows;
  int32_t a_ld;           /* elements */
  int32_t cin;            /* channels contracted per tap (multiple of 64 for tcgen05) */
  int32_t taps;           /* 1 or 9 */
  int32_t dtype;          /* CROG_F32 | CROG_BF16: type of a and w */
  int32_t M;              /* rows enumerated */
  int32_t sample_rows;    /* enumerated rows per sample (0: no sample structure) */
  int32_t H, W;           /* interior spatial dims (0,0: plain matrix rows) */
  int32_t in_padded;      /* enumerated rows are padded-layout rows */
  int32_t out_padded;     /* output (and residual) rows are padded-layout rows */
  const void* w;          /* weights [N, taps*cin], K-major */
  int32_t N;
  int64_t w_sample_stride;/* elements between per-sample weight matrices (0: shared) */
  const float* scale;     /* [N] or NULL (=1) */
  const float* bias;      /* [N] or NULL (=0) */
  const float* addmat;    /* [addmat_rows, N] fp32 added before scale, or NULL */
  int32_t addmat_rows;    /* period: indexed by the interior pixel index / row-in-sample */
  int32_t act;            /* CROG_ACT_* */
  const float* gate;      /* [num_samples, N] or NULL */
  const float* scale2;    /* [N] (with gate) */
  const float* bias2;     /* [N] (with gate) */
  const void* residual;   /* [*, res_ld] of out_dtype, or NULL */
  int32_t res_ld;
  int32_t residual_relu;
  void* out;
  int32_t out_ld;
  int32_t out_dtype;      /* CROG_F32 | CROG_BF16 */
  int32_t impl;           /* CROG_IMPL_* */
  int32_t out_sample_rows;/* compact output: rows per sample in the OUTPUT matrix (0: H*W). Lets several feature
                             levels of one sample land back to back in one [B, sum_l H_l*W_l, N] tensor (torch.cat of
                             model/ssg.py:266-269): pass `out` already offset to the level's first row. */
  int32_t tile_cfg;       /* tcgen05 tile configuration: 0 = built-in heuristic, CROG_TILE_* = forced (the plan-time
                             autotuner times the applicable ones per layer; every configuration accumulates the k-blocks
                             in the same order, so results are bit-identical across them).  A configuration that does
                             not apply to the shape returns CROG_E_BADSHAPE. */
  /* LayerNorm folded into the next contraction (decoder FFN, model/layers.py:302-308: Linear -> ReLU -> LayerNorm ->
     Linear).  LN(h) W^T = rstd_r (h W'^T) - mu_r rstd_r s + c with W' = W * gamma (columns), s_n = sum_k W'[n,k],
     c = W beta + b: the producer GEMM also writes, per row and 64-column chunk, (sum h, sum h^2) of its fp32 outputs
     (row_stats_out: M x N/64 x 2 fp32 values, bf16 TMA-epilogue path only), and the consumer GEMM (weights W', scale = s,
     bias = c) reads them back (row_stats_in, row_stats_chunks pairs per row over row_stats_width values) and applies
     out = acc * rstd_r + (-mu_r rstd_r) * s_n + c_n.  Plain row matrices / compact layouts only (output row ==
     enumerated row). */
  /* Second activation operand, contracted after the first along K (K = cin + cin2, weights [N, cin + cin2]): fuses two
     1x1 convolutions that are summed, e.g. the last convolution of a bottleneck and its downsample branch
     (model/clip.py:44-57: out = relu(bn3(conv3(t)) + bn_d(conv_d(pool(x)))) with both BatchNorm scales folded into the
     weights).  Same row enumeration as `a`; taps == 1; tcgen05 path only. */
  const void* a2;
  int32_t a2_ld;
  int32_t cin2;
  float* row_stats_out;
  const float* row_stats_in;
  int32_t row_stats_chunks;
  int32_t row_stats_width;
  float row_stats_eps;
  int32_t max_ctas;       /* tcgen05 path: upper bound on the persistent grid (0: one CTA per SM).  The forward plan runs the
                             text tower on a few SMs BESIDE the memory-bound image front: both branches' persistent CTAs
                             need most of an SM's shared memory, so without disjoint SM budgets they serialise. */
  int32_t tap_mask;       /* 3x3 convolutions on the resident-weight path: bit (ky*3+kx) set = tap contracted; 0 = all nine.
                             Masked taps must have zero weights (other tile configurations still contract them): the
                             pixel-pair stem (two 32-channel pixels per 64-channel row) only reaches two of the three
                             horizontal pair offsets per output parity. */
  int32_t reverse;        /* tcgen05 path: 1 = the persistent CTAs walk the tiles from the last to the first.  Results are
                             unchanged; a consumer that starts where its producer finished finds the producer's last
                             ~60-100 MB of output still in the 126 MB L2 (the forward plan alternates directions). */
} CrogGemm;
enum {
  CROG_TILE_AUTO = 0,
  CROG_TILE_128x128 = 1,        /* one CTA, 128 x 128 tiles, 4 operand stages, two epilogue groups */
  CROG_TILE_128x256 = 2,        /* one CTA, 128 x 256 tiles, 4 stages, one epilogue group */
  CROG_TILE_128x256_E8 = 3,     /* one CTA, 128 x 256 tiles, 3 stages, two epilogue groups */
  CROG_TILE_PAIR_256x256 = 4,   /* CTA pair (cta_group::2), 256 x 256 tiles, 6 stages, one epilogue group per CTA */
  CROG_TILE_PAIR_256x256_E8 = 5,/* CTA pair, 256 x 256 tiles, 4 stages, two epilogue groups per CTA */
  CROG_TILE_PAIR_256x128 = 6,   /* CTA pair, 256 x 128 tiles, 6 stages, two epilogue groups per CTA */
  CROG_TILE_128x64 = 7,         /* one CTA, 128 x 64 tiles, 5 stages */
  CROG_TILE_CONV3 = 8,          /* 128 x 64 tiles, resident 3x3 weights, one TMA box per ky band (N <= 64, cin == 64) */
  CROG_TILE_128x128_S3 = 9,     /* one CTA, 128 x 128 tiles, 3 operand stages, two epilogue groups, scale / bias of the tile
                                   staged in shared memory (the 4-stage ring of CROG_TILE_128x128 leaves no room for it):
                                   the default for short contractions, where the epilogue is the pace */
  CROG_TILE_128x128_E12 = 10,   /* one CTA, 128 x 128 tiles, 3 stages, THREE epilogue groups (12 warps, three TMEM accumulators,
                                   2 staging buffers per warp): more warps per scheduler for epilogue-bound layers */
  /* BAND configurations (3x3 convolutions on the zero-haloed layout with shared weights): one (BM + 2)-row activation
     band per (ky, 64-channel chunk) serves the three kx taps through row-shifted views while the weight tiles stream
     through their own ring - a third of the activation traffic between L2 and the SMs, which is what bounds the large
     convolutions.  Same k-block order as every other configuration, hence bit-identical results. */
  CROG_TILE_BAND_PAIR_256x256 = 11,    /* CTA pair, 256 x 256 tiles, 3 band + 6 weight stages, one epilogue group per CTA */
  CROG_TILE_BAND_PAIR_256x256_E8 = 12, /* CTA pair, 256 x 256 tiles, 3 band + 5 weight stages, two epilogue groups per CTA */
  CROG_TILE_BAND_PAIR_256x128 = 13,    /* CTA pair, 256 x 128 tiles, 4 band + 8 weight stages, two epilogue groups per CTA */
  CROG_TILE_BAND_128x128 = 14,         /* one CTA, 128 x 128 tiles, 3 band + 6 weight stages, two epilogue groups */
  CROG_TILE_BAND_128x256 = 15,         /* one CTA, 128 x 256 tiles, 3 band + 4 weight stages, one epilogue group */
  CROG_TILE_CONV3_E12 = 16,            /* CROG_TILE_CONV3 with three epilogue groups and three band stages: a 128 x 64 tile's
                                          36 MMAs take ~1.2k clocks, its epilogue ~6k, so two groups leave the tensor pipe idle */
  CROG_TILE_CONV3_PAIR = 17,           /* CROG_TILE_CONV3 on a CTA pair (256 x 64 tiles, each CTA keeps half of the weight rows):
                                          one MMA-issuing thread feeds two SMs - the N = 64 MMAs are so short (32 clocks) that
                                          the single issuing thread, not the tensor pipe, paces the one-CTA form */
  CROG_TILE_CONV3_DUAL = 18,           /* CROG_TILE_CONV3 with TWO MMA-issuing warps (one per accumulator buffer, alternate tiles) */
  CROG_TILE_COUNT = 19
};
int crog_gemm(const CrogGemm* g, void* stream);

/* ------------------------------------------------------------------ layout / resampling
 * replaces: nn.AvgPool2d(2) (clip.py:23,35,184; layers.py:386), F.interpolate(scale_factor=2,
 * mode='bilinear') (layers.py:54,56,382,393), torch.cat along channels (writes a channel
 * slice: pass out already offset, out_ld = full width).  mode: 0 copy, 1 avgpool2, 2 bilinear x2 (align_corners=False),
 * 3 x[::2, ::2] (the input side of a stride-2 1x1 convolution, model/ssg.py:80-81), 4 bilinear x2 with
 * align_corners=True (nn.Upsample of model/ssg.py:159). */
int crog_resample(const void* in, int32_t in_ld, int32_t in_padded, void* out, int32_t out_ld,
                  int32_t out_padded, int32_t B, int32_t H, int32_t W, int32_t C, int32_t mode,
                  int32_t dtype, void* stream);

/* Stem conv1: 3x3 stride 2 pad 1 on NCHW fp32 images + folded BN + ReLU, writing the padded
 * NHWC layout (model/clip.py:165-170,208-211). w [cout,3,3,3] fp32. Only channels [0, cout) of the interior pixels are
 * written: halo pixels and the padding channels [cout, out_ld) must already be zero (the caller's zero-initialised buffer).
 * pixel_pairs != 0: the output is the PIXEL-PAIR layout instead: a zero-haloed [B, OH+2, OW/2+2] grid of rows holding two
 * horizontally adjacent pixels, channels [0,cout) = pixel 2s, [cout, 2*cout) = pixel 2s+1 (needs OW even, out_ld >= 2*cout):
 * a 32-channel tensor then fills 64-channel (128-byte) rows with no padding channels. */
int crog_stem_conv1(const float* img, int32_t B, int32_t Hin, int32_t Win, const float* w,
                    const float* scale, const float* bias, int32_t cout, void* out, int32_t out_ld,
                    int32_t out_dtype, int32_t pixel_pairs, void* stream);

/* LayerNorm over the last dim (clip.py:226-231, layers.py:288-311): y = LN(x)*g + b, optional
 * out = residual + y (fp32 residual stream).  x_dtype/out_dtype are CROG_F32|CROG_BF16. */
int crog_layernorm(const void* x, int32_t x_dtype, const float* gamma, const float* beta, const float* residual,
                   void* out, int32_t out_dtype, int64_t rows, int32_t D, float eps, void* stream);

/* Two chained LayerNorms of a decoder layer (layers.py:318-329): y = residual + LN(x; g1, b1) (the fp32 residual
 * stream; y may alias residual), z = LN(y; g2, b2).  Same arithmetic as two crog_layernorm calls, one pass over HBM. */
int crog_layernorm_chain(const void* x, int32_t x_dtype, const float* g1, const float* b1, const float* residual,
                         float* y, const float* g2, const float* b2, void* z, int32_t z_dtype, int64_t rows,
                         int32_t D, float eps, void* stream);

/* Token embedding + positional embedding (clip.py:440-443): out[b*L+l] = emb[word[b,l]] + pos[l], fp32.
 * `vocab` = rows of emb; an id outside [0, vocab) is a device-side assertion failure (message + trap), as with the
 * reference's nn.Embedding on CUDA — never a silent out-of-bounds read. */
int crog_embed_tokens(const int64_t* word, const float* emb, const float* pos, float* out, int32_t B,
                      int32_t L, int32_t D, int32_t vocab, void* stream);
/* EOT gather (clip.py:450-451): out[b] = x[b*L + argmax_l word[b,l]]. */
int crog_gather_eot(const int64_t* word, const void* x, int32_t x_dtype, void* out, int32_t out_dtype,
                    int32_t B, int32_t L, int32_t D, void* stream);

/* Softmax attention, head_dim 64 (clip.py:123-140,258-262; layers.py:313-333).
 * q [B*Tq, ldq], k/v [B*Tk, ldk/ldv], head h uses columns [h*64, h*64+64); o [B*Tq, ldo].
 * causal: key j visible to query i iff j <= i.  pad_word: optional int64 [B,Tk]; keys with id 0 masked. */
int crog_attention(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                   void* o, int32_t ldo, int32_t B, int32_t heads, int32_t Tq, int32_t Tk, float scale,
                   int32_t causal, const int64_t* pad_word, int32_t dtype, void* stream);

/* Projector text branch (layers.py:91-93) folded with vis.4 (layers.py:58,70-77):
 * t = Linear(state); w_dyn = t[:, :-1] as [C,3,3]; b_dyn = t[:, -1];
 * wfold[b, h*9 + tap, j] = sum_c w_dyn[b,c,tap] * V[h*C + c, j]   (j < C)
 * wfold[b, h*9 + tap, C] = sum_c w_dyn[b,c,tap] * vb[h*C + c] (+ b_dyn[b] on the centre tap), zero for j > C
 * wfold: [B, rows_per_sample, Cpad] of dtype (rows_per_sample >= 9*NH; rows >= 9*NH are left untouched).
 * With a constant-one channel at index C of the feature map, Z = feat x wfold[b]^T (a per-sample 1x1 GEMM) holds the
 * nine per-tap partial products of all NH heads, biases included, and crog_dynconv_gather sums the nine shifted
 * reads: the five grouped 3x3 dynamic convolutions of the reference in one pass over the features. */
int crog_dynw_fold(const void* state, int32_t state_dtype, const float* txt_w, const float* txt_b,
                   const float* v_w, const float* v_b, float* scratch /*[B, 9*C+1]*/, void* wfold,
                   int32_t dtype, int32_t B, int32_t word_dim, int32_t C, int32_t NH, int32_t rows_per_sample,
                   int32_t Cpad, void* stream);
/* z: zero-haloed [B*(H+2)*(W+2), ldz] fp32 with column h*9+tap; out: NH planes [NH][B,H,W] fp32 (the B x 1 x H x W logits). */
int crog_dynconv_gather(const float* z, int32_t ldz, float* out, int32_t B, int32_t H, int32_t W, int32_t NH, void* stream);

/* Element-wise dtype conversion (CROG_F32 <-> CROG_BF16) of n contiguous elements. */
int crog_cast(const void* in, int32_t in_dtype, void* out, int32_t out_dtype, int64_t n, void* stream);

/* [B*HW, ld] fp32 head columns -> NH planes [NH][B, HW] fp32 (the five B x 1 x 104 x 104 logits). */
int crog_split_heads(const float* in, int32_t ld, float* out, int64_t rows, int32_t NH, void* stream);

/* ------------------------------------------------------------------ engine glue
 * engine/crog_engine.py:183-211: sigmoid (planes flagged in sigmoid_mask bit i) + bicubic
 * (A=-0.75, align_corners=True) resize of NP planes [NP][B,Hin,Win] -> [NP][B,Hout,Wout], fp32. */
int crog_sigmoid_bicubic(const float* in, float* out, int32_t NP, int32_t B, int32_t Hin, int32_t Win,
                         int32_t Hout, int32_t Wout, uint32_t sigmoid_mask, void* stream);

/* ------------------------------------------------------------------ grasp decode + Jaccard tail
 * utils/grasp_eval.py:289-302 (detect_grasps = skimage peak_local_max(min_distance=2,
 * threshold_abs, num_peaks=K) + angle/width gather) batched over B maps of H x W fp32.
 *   peaks  [B,K,2] int32 (row, col), -1 padded;  n_peaks [B] int32
 *   grasps [B,K,5] float64 rows [x, y, width*100, 20, angle_deg]
 * workspace: crog_detect_workspace_bytes(B,H,W,K) bytes, 16-byte aligned.  Maps with W % 4 == 0 and a 16-byte aligned base are
 * streamed by the bulk-copy staged scan; any other width / alignment takes the generic scan kernel (same results). */
int64_t crog_detect_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t K);
int crog_detect_grasps(const float* q, const float* sin_m, const float* cos_m, const float* wid, int32_t B,
                       int32_t H, int32_t W, int32_t K, float threshold, int32_t* peaks, int32_t* n_peaks,
                       double* grasps, void* workspace, void* stream);
/* Whole-map angle field, the second return value of detect_grasps (grasp_eval.py:293). */
int crog_angle_map(const float* sin_m, const float* cos_m, float* out, int64_t n, void* stream);

/* utils/grasp_eval.py:305-374 batched: for every sample the K predicted rectangles against its
 * M ground-truth rectangles [B,Mmax,6] float64 (gt_count [B]); GT is edited in place like the
 * reference (h:=20, w:=clip(w,0,100)) when edit_gt != 0 (calculate_jacquard_index); with edit_gt == 0 the
 * rectangles are used as given (calculate_iou).  Outputs: inter/uni [B,K,Mmax] int32 pixel counts
 * (0,0 when angle-gated), j_flags [B,2] int32 = {J@1 (first prediction only), J@K (all)},
 * counters int64[4] += {correct@1, total@1, correct@K, total@K}.  inter/uni may be NULL. */
int crog_jaccard(const double* grasps, const int32_t* n_peaks, int32_t K, double* gt, const int32_t* gt_count,
                 int32_t Mmax, int32_t B, int32_t* inter, int32_t* uni, int32_t* j_flags, int64_t* counters,
                 int32_t edit_gt, void* stream);

/* ------------------------------------------------------------------ SSG (BASELINE config 4): model/ssg.py, grasp_eval.py:55-221
 * 7x7 / stride 2 / pad 3 stem (model/ssg.py:65,217-222) as a patch gather: out [B*OH*OW, Kp] of out_dtype, column
 * (ky*7+kx)*cin + c, zero padded to Kp; rgb [B,3,H,W] and depth [B,1,H,W] fp32 NCHW (depth may be NULL when cin == 3).
 * The convolution itself is then crog_gemm on the patch matrix. */
int crog_stem7_patches(const float* rgb, const float* depth, int32_t B, int32_t H, int32_t W, int32_t cin, int32_t Kp,
                       void* out, int32_t out_dtype, void* stream);
/* nn.MaxPool2d(kernel_size=3, stride=2, padding=1) on NHWC rows (model/ssg.py:67,101). */
int crog_maxpool3s2(const void* in, int32_t in_ld, int32_t in_padded, void* out, int32_t out_ld, int32_t out_padded,
                    int32_t B, int32_t H, int32_t W, int32_t C, int32_t dtype, void* stream);
/* 3x3 / stride s / pad 1 patches of a zero-haloed NHWC tensor -> compact [B*OH*OW, 9*C] (tap-major), the A operand of
 * the stride-2 3x3 convolutions (model/ssg.py:22,180-183). */
int crog_patches3(const void* in, int32_t in_ld, void* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t stride,
                  int32_t dtype, void* stream);
/* Head finalisation (model/ssg.py:136-139,272-275): in fp32 [rows, ld] with columns [na*nc logits | na*4 box deltas];
 * cls[rows*na, nc] = softmax over classes, box[rows*na, 4] = the deltas. */
int crog_ssg_heads(const float* in, int32_t ld, int64_t rows, int32_t na, int32_t nc, float* cls, float* box, void* stream);

/* Detection stage of ssg_post_processing for ONE image (grasp_eval.py:113-150 + fast_nms :55-93):
 * keep[n] = max_{c>=1} cls[n,c] > score_thr; boxes[N,4] = decoded point-form boxes clipped to [0,1] (all anchors);
 * per foreground class the top_k kept anchors by score (ties: lower anchor index), upper-triangular IoU suppression
 * at iou_thr, then the max_det best survivors over all classes, and finally the score > score_thr2 filter (applied
 * only if at least one detection passes, as in the reference).  Outputs (device): det_n, det_anchor[max_det],
 * det_class[max_det] (0-based foreground class; the reference reports class + 1), det_score[max_det]. */
int64_t crog_ssg_nms_workspace_bytes(int32_t num_classes, int32_t top_k);
/* fast_nms alone (grasp_eval.py:55-93) on already decoded boxes: cls [N, num_classes] (column 0 = background, ignored),
 * keep [N] (0/1 per anchor), boxes [N,4]; det_anchor indexes the N rows. */
int crog_ssg_fast_nms(const float* cls, const int32_t* keep, const float* boxes, int32_t N, int32_t num_classes, float iou_thr,
                      int32_t top_k, int32_t max_det, float score_thr2, int32_t* det_n, int32_t* det_anchor, int32_t* det_class,
                      float* det_score, void* workspace, void* stream);
int crog_ssg_detect(const float* cls, const float* box, const float* anchors, int32_t N, int32_t num_classes, float score_thr,
                    float iou_thr, int32_t top_k, int32_t max_det, float score_thr2, int32_t* keep, float* boxes,
                    int32_t* det_n, int32_t* det_anchor, int32_t* det_class, float* det_score, void* workspace, void* stream);
/* crog_ssg_detect for B images in three launches: cls [B,N,nc], box [B,N,4], keep [B,N], boxes [B,N,4], det_n [B],
 * det_anchor / det_class / det_score [B,max_det]; workspace = B slices of workspace_stride bytes (>= the per-image size,
 * multiple of 16).  Per image identical to crog_ssg_detect. */
int crog_ssg_detect_batched(const float* cls, const float* box, const float* anchors, int32_t B, int32_t N, int32_t num_classes,
                            float score_thr, float iou_thr, int32_t top_k, int32_t max_det, float score_thr2, int32_t* keep,
                            float* boxes, int32_t* det_n, int32_t* det_anchor, int32_t* det_class, float* det_score, void* workspace,
                            int64_t workspace_stride, void* stream);
/* Mask stage (grasp_eval.py:171-194): lowres[d][k] = crop(act_k(protos . coef_k(d))) for k = ins, qua, sin, cos, wid
 * (sigmoid on ins / qua / wid), [max_det, 5, h, w]; out[k][d] = bilinear resize to resize_to^2 (align_corners=False)
 * cropped to [out_h, out_w], map-major [5, out_det_stride, out_h, out_w] (out_det_stride <= 0: max_det; a larger stride
 * lets several images fill disjoint instance ranges of one batch-wide tensor); the instance plane is thresholded
 * (> 0.5 -> 1.0).  Only the first *det_n (<= max_det) detections are written.  quality_raw (may be NULL): when given,
 * the un-smoothed quality plane is written there as [max_det, out_h, out_w] instead of into out's plane 1, so that
 * crog_gaussian can smooth it INTO out's plane 1 out of place (its one-kernel form). */
int crog_ssg_masks(const float* protos, int32_t h, int32_t w, int32_t num_protos, const float* coef, const float* gcoef,
                   const float* boxes, const int32_t* det_anchor, const int32_t* det_n, int32_t max_det, float* lowres,
                   float* out, float* quality_raw, int32_t out_det_stride, int32_t out_h, int32_t out_w, int32_t resize_to,
                   void* stream);
/* The same mask stage for ALL instances of a batch in two launches: instance i (of `total`) is detection inst_det[i] of
 * image inst_image[i]; protos [B,h,w,np], coef [B,N,np], gcoef [B,N,4,np], boxes [B,N,4], det_anchor [B,max_det].
 * lowres [total,5,h,w]; out map-major [5,total,out_h,out_w]; quality_raw [total,out_h,out_w] or NULL (as above). */
int crog_ssg_masks_batched(const float* protos, int32_t h, int32_t w, int32_t num_protos, const float* coef, const float* gcoef,
                           const float* boxes, const int32_t* det_anchor, int32_t N, int32_t max_det, const int32_t* inst_image,
                           const int32_t* inst_det, int32_t total, float* lowres, float* out, float* quality_raw, int32_t out_h,
                           int32_t out_w, int32_t resize_to, void* stream);
/* skimage.filters.gaussian(map, sigma, preserve_range=True) of grasp_eval.py:198 = scipy.ndimage.gaussian_filter(mode='nearest'):
 * separable (rows first), float64 accumulation in scipy's tap order, float32 result per pass.  weights_host: the 2*radius+1
 * normalised float64 taps (HOST pointer; computed by the caller exactly as scipy does).  Planes smoothed:
 * p * plane_stride + plane_sel for p < min(P, *n_planes) (n_planes may be NULL); tmp and out use the same layout.  With
 * out != in and radius 8 (sigma 2) both passes run in one kernel (the float32 intermediate stays in shared memory) and tmp
 * may be NULL; out may alias in, which selects the two-pass form through tmp.  Bit-identical either way. */
int crog_gaussian(const float* in, float* tmp, float* out, int32_t P, int32_t H, int32_t W, const double* weights_host,
                  int32_t radius, const int32_t* n_planes, int32_t plane_stride, int32_t plane_sel, void* stream);

/* ------------------------------------------------------------------ letterbox warps around the model (OpenCV-exact)
 * cv2.warpAffine(src, M, (w,h), flags=cv2.INTER_CUBIC, borderValue=border_value) of engine/crog_engine.py:387-391,
 * 499-517 (the inverse letterbox of the prediction / target maps), batched: src [NP][B][Hs][Ws] fp32 -> dst
 * [NP][B][h][w] fp32.  minv [B,6] float64 = the ALREADY INVERTED 2x3 matrices (dst pixel -> src pixel; OpenCV inverts
 * M itself when WARP_INVERSE_MAP is not passed - the host does that in float64, crog_b200/utils/warp.py).
 * Bit-exact with OpenCV: 1/32-pixel fixed-point coordinates, float32 weight products and summation order. */
int crog_warp_affine_cubic_f32(const float* src, int32_t NP, int32_t B, int32_t Hs, int32_t Ws, const double* minv,
                               float* dst, int32_t h, int32_t w, float border_value, void* stream);
/* utils/dataset.py:843-866: cv2.warpAffine(img_u8, mat, input_size, INTER_CUBIC, borderValue=border_rgb) then
 * .float().div_(255.).sub_(mean).div_(std): img [B][Ho][Wo][3] uint8 RGB -> out [B][3][Sh][Sw] fp32.  minv as above;
 * border_rgb (3 doubles), mean, std_ (3 floats each) are HOST pointers.  Bit-exact (int16 fixed-point weights).
 * workspace: crog_preprocess_workspace_bytes() device bytes (OpenCV's 32 x 32 table of int16 weight sets, rebuilt by the call). */
int64_t crog_preprocess_workspace_bytes(void);
int crog_preprocess_u8(const uint8_t* img, int32_t B, int32_t Ho, int32_t Wo, const double* minv, float* out, int32_t Sh,
                       int32_t Sw, const double* border_rgb, const float* mean, const float* std_, void* workspace, void* stream);
/* engine/crog_engine.py:500-501,515-518: per sample pixel counts of (pred > thr) & (target != 0) and (pred > thr) |
 * (target != 0): counts [B,2] int64 = {inter, union} (zeroed by the call). */
int crog_mask_iou(const float* pred, const float* target, int32_t B, int64_t n, float thr, int64_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CROG_B200_H_ */
