/* libst_b200 -- C-ABI of the B200 (sm_100a) hot path of Soft-Truncation.
 *
 * Every entry point takes raw DEVICE pointers, explicit sizes and a cudaStream_t (passed as
 * void*), enqueues work on that stream and returns immediately: 0 = ok, otherwise an error code
 * whose text is available from st_last_error().  Nothing here allocates, synchronises or throws.
 * The caller (PyTorch on the host side) owns all memory.
 *
 * Layout convention: activations are NHWC ("pixels x channels", channels contiguous); dtype
 * codes are ST_F32 = 0 and ST_BF16 = 1.  All reductions accumulate in fp32.
 *
 * What each group replaces in the reference (file:line relative to the reference root):
 *   st_upfirdn2d        op/upfirdn2d.cpp:12-19, op/upfirdn2d_kernel.cu:209-368
 *   st_fused_bias_act   op/fused_bias_act.cpp:4-17, op/fused_bias_act_kernel.cu:18-98
 *   st_gemm             F.conv2d / nn.Linear / NIN einsum / attention einsums
 *                       (models/layers.py:100-124,546-555; models/layerspp.py:95-99) incl. their
 *                       autograd backward (dgrad / wgrad)
 *   st_gn_*             nn.GroupNorm + SiLU + Dropout (models/layerspp.py:232,244-245,258,275-278)
 *   st_resample2x       naive_upsample_2d / naive_downsample_2d (models/up_or_down_sampling.py:59-69)
 *   st_softmax_*        F.softmax in AttnBlockpp (models/layerspp.py:97)
 *   st_timestep_embedding / st_fourier_embedding
 *                       models/layers.py:515-529, models/layerspp.py:45-54
 *   st_dsm_perturb / st_dsm_loss
 *                       losses.py:116-132
 *   st_sumsq / st_adam_ema
 *                       losses.py:44-58 (clip_grad_norm_ + Adam) and models/ema.py:32-51
 *   st_pc_update / st_batch_norms
 *                       sampling.py:185-210, 263-292, 402-408
 */
#ifndef ST_B200_H_
#define ST_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ST_F32 0
#define ST_BF16 1

/* ------------------------------------------------------------------ library */
int st_version(void);
const char* st_last_error(void);
/* 1 when the tcgen05/TMA GEMM path can run on the current device (sm_100 + driver entry point). */
int st_tc_available(void);

/* ------------------------------------------------------------------ GEMM / implicit-GEMM conv
 * C[b][m][n] = alpha * ( sum_k A(b,m,k) * B(b,n,k) + bias[n] + rowbias[m / rows_per_rb][n]
 *                        + residual[b][m][n] )
 * or, with accumulate != 0,  C[b][m][n] += alpha * sum_k ...   (C must be fp32).
 *
 * a_mode / b_mode select how operand elements are addressed:
 *   ST_OP_STRIDED   X(b,i,k) = X[b*sXb + i*sXi + k*sXk]            (one of sXi, sXk must be 1)
 *   ST_OP_GATHER    implicit im2col of an NHWC tensor: for A, i is the output pixel and
 *                   k = tap*(C1+C2) + c;  for B (weight gradient), k is the pixel and
 *                   i = tap*(C1+C2) + c.  Taps walk a kh x kw window with "same" zero padding.
 *                   The channel axis may be the concatenation of two tensors (src, src2).
 *   ST_OP_DGRADW    (B only) conv weights W[co][tap][ci] read as B(n=ci, k=tap'*Co+co) =
 *                   W[co][ntaps-1-tap'][ci]: the data-gradient of a "same" convolution.
 */
#define ST_OP_STRIDED 0
#define ST_OP_GATHER 1
#define ST_OP_DGRADW 2

#define ST_BACKEND_AUTO 0
#define ST_BACKEND_SIMT 1
#define ST_BACKEND_TCGEN05 2

typedef struct st_gemm_args {
  int32_t a_mode, b_mode;
  int32_t in_dtype;       /* dtype of A, B and residual */
  int32_t out_dtype;      /* dtype of C */
  int32_t backend;
  int32_t accumulate;
  int32_t split_k;        /* >1 only with accumulate (fp32 atomics) */
  int32_t M, N, K, batch;
  int64_t sAm, sAk, sAb;
  int64_t sBn, sBk, sBb;
  int64_t sCm, sCb;       /* C row stride / batch stride (n stride is 1) */
  const void* A;
  const void* A2;         /* second gather source (channels C1..C1+C2) or NULL */
  const void* B;
  const void* B2;
  void* C;
  /* gather geometry */
  int32_t n_img, H, W, C1, C2, kh, kw;
  /* epilogue */
  const float* bias;      /* [N] or NULL */
  const float* rowbias;   /* [M / rows_per_rb][ld_rb] or NULL */
  int32_t rows_per_rb;
  int64_t ld_rb;
  const void* residual;   /* in_dtype, [b][m][n] with strides sRb, sRm or NULL */
  int64_t sRm, sRb;
  float alpha;
  /* GroupNorm statistics of C as a by-product of the epilogue (optional; tcgen05 backend, bf16 C, rows = pixels of
   * NHWC images of gn_hw pixels each): per block of R = min(gn_hw, 128) rows and per 4 adjacent channels, the sum and the
   * sum of squares of the stored values -> gn_part[M / R][N / 4][2] (fp32).  *gn_rows_out (HOST int, optional) receives
   * R when the launch emits them and 0 when it cannot (the caller then runs st_gn_stats).  st_gn_apply consumes them. */
  float* gn_part;
  int32_t gn_hw;
  int32_t* gn_rows_out;
  /* GroupNorm BACKWARD, phase 1, as the epilogue of the data-gradient GEMM (optional; tcgen05 backend, bf16 C, no bias /
   * residual / accumulate).  With dz_x set, the GEMM result dy = alpha * A B is the gradient with respect to
   * y = dropout(act(GroupNorm(x))) and the epilogue stores, instead of dy,
   *     dz = dy * keep * act'(u),   u = x * cst.x + cst.y   (the GroupNorm output before the activation)
   * and emits, per block of 32 rows and per 4 adjacent channels, the two sums the GroupNorm backward needs of the
   * STORED dz:  gn_part[M / 32][N / 4][2] = (sum gamma * dz, sum dz * (u - beta))  [= sum gamma*dz*xhat].
   * st_gn_bwd_dz_apply turns dz and these sums into dx: the reduction pass over (x, dy) of the two-pass backward
   * and all of its activation / dropout arithmetic run under the GEMM's main loop.
   *   dz_x     bf16 [M][N] rows of stride dz_ldx: the GroupNorm input (rows = pixels, gn_hw per image)
   *   dz_cst   fp32 [M / gn_hw][N][4]: (rstd*gamma, beta - mean*rstd*gamma, gamma, beta) per (image, channel)
   *            (st_gn_bwd_consts writes it)
   *   dz_keep  dropout keep flags, 1 bit per element ([M][N / 8] bytes, as st_gn_apply wrote them) or NULL
   *   dz_inv_keep  1 / (1 - p_drop) (1 without dropout);  dz_act  0 none, 1 SiLU
   * *gn_rows_out receives 32 when the launch did all this and 0 when it cannot (C then holds the plain dy and the
   * caller runs the ordinary backward). */
  const void* dz_x;
  int64_t dz_ldx;
  const float* dz_cst;
  const uint8_t* dz_keep;
  float dz_inv_keep;
  int32_t dz_act;
} st_gemm_args;

int st_gemm(const st_gemm_args* args, void* stream);
/* Number of bf16 problems the AUTO backend had to run on the fp32-FMA kernel because the tcgen05 path cannot express
 * them (alignment, channel counts, non power-of-two images) since the last reset - benchmarks and full-size tests assert
 * it is zero - and the reason given for the most recent one. */
int st_gemm_simt_fallbacks(int reset);
const char* st_gemm_simt_fallback_reason(void);

/* ------------------------------------------------------------------ fused attention core
 * AttnBlockpp's  w = softmax(q k^T * scale),  o = w v  (models/layerspp.py:95-99) in ONE tcgen05 kernel: the logits live
 * in tensor memory, the probabilities in shared memory; only q, k, v are read and o is written.
 *   qkv   bf16 [n_img * L][3C]: q | k | v of every pixel (the packed projection st_gemm produces)
 *   o     bf16 [n_img * L][C]
 *   p_out bf16 [n_img][L][L] or NULL: the normalised probabilities, kept for the backward pass (training)
 * st_attn_fwd_supported: 1 when the kernel can run the shape (sm_100, L = 256, C = 256, bf16); other shapes use the
 * st_gemm / st_softmax_fwd path. */
int st_attn_fwd_supported(int L, int C, int dtype);
int st_attn_fwd(const void* qkv, void* o, void* p_out, int n_img, int L, int C, float scale, void* stream);

/* ------------------------------------------------------------------ GroupNorm (+SiLU, +dropout)
 * x is NHWC [n_img][hw][C1] (+ optional second tensor [n_img][hw][C2] concatenated on channels),
 * G groups of (C1+C2)/G adjacent channels, statistics per (image, group).
 */
/* partial sums: part[n_img][splits][G][2] (sum, sum of squares) */
int st_gn_stats(const void* x1, const void* x2, int dtype, int n_img, int hw, int C1, int C2, int G,
                int splits, float* part, void* stream);
/* mean/rstd [n_img][G] from the partials */
int st_gn_finalize(const float* part, int n_img, int splits, int G, int64_t count, float eps,
                   float* mean, float* rstd, void* stream);
/* y = dropout( act( gamma*(x-mean)*rstd + beta ) );  act: 0 none, 1 SiLU.
 * dropout: keep-mask multiplies by 1/(1-p); `mask` (same dtype/shape as y, already scaled) is used
 * when non-NULL, else if p > 0 a counter-based RNG keyed by (seed, element index).  `keepbits` (optional,
 * n_img*hw*C/8 bytes) receives the drawn keep flags, one bit per element, for the backward kernels.
 * `part` != NULL (with `splits`, `count` = hw*(C/G), `eps`): the statistics are finalised inside the kernel from the
 * partial sums of st_gn_stats - no st_gn_finalize launch - and mean / rstd are OUTPUTS (kept for the backward).
 * `q1` != NULL (instead of `part`; with `qrows`, `count`, `eps`; `q2` for the second source): the statistics come from
 * the per-(qrows rows, 4 channels) sums the GEMMs that produced x1 / x2 emitted (st_gemm_args.gn_part) - no pass
 * over x for the statistics at all. */
int st_gn_apply(const void* x1, const void* x2, int dtype, int n_img, int hw, int C1, int C2, int G,
                const float* gamma, const float* beta, float* mean, float* rstd, int act,
                float p_drop, uint64_t seed, const void* mask, uint8_t* keepbits, void* y, const float* part,
                int splits, int64_t count, float eps, const float* q1, const float* q2, int qrows, void* stream);
/* Device-resident 64-bit addend of the `seed` argument of every st_gn_apply / st_gn_fwd_fused launch that draws its own
 * dropout mask (NULL = none, the default): the kernel keys its generator with seed + *ptr.  A training step captured
 * in a CUDA graph bakes `seed` into the launch; the host advances *ptr between replays so that every step draws a
 * fresh mask, exactly the seeds the eager path would have passed by value. */
int st_set_dropout_seed_offset(const void* dev_u64);
/* Statistics + apply in ONE launch: a thread-block cluster of `chunks` CTAs keeps an image resident in shared memory
 * (all of its cp.async copies in flight at once), exchanges per-group partial sums through distributed shared memory
 * and writes y from the resident copy - the tensor crosses HBM once in and once out.  mean / rstd [n_img][G] are
 * outputs (kept for the backward).  st_gn_fwd_fused_chunks returns the cluster size for a shape, 0 = the image does
 * not fit 16 CTAs x 64 KB (or the grid would under-fill the GPU, or ST_GN_FWD_FUSED=0): use st_gn_stats + st_gn_apply. */
int st_gn_fwd_fused_chunks(int n_img, int hw, int C);
int st_gn_fwd_fused(const void* x1, const void* x2, int dtype, int n_img, int hw, int C1, int C2, int G,
                    const float* gamma, const float* beta, float eps, int act, float p_drop, uint64_t seed,
                    const void* mask, uint8_t* keepbits, void* y, float* mean, float* rstd, int chunks, void* stream);
/* backward, pass 1: per (image, pixel split, channel) sums  red[n_img][splits][C][2] = (sum dz, sum dz*xhat) where
 * dz = dy * dropout_mask * act'(.)  */
int st_gn_bwd_reduce(const void* x1, const void* x2, const void* dy, int dtype, int n_img, int hw, int C1,
                     int C2, int G, const float* gamma, const float* beta, const float* mean,
                     const float* rstd, int act, float p_drop, uint64_t seed, const void* mask,
                     const uint8_t* keepbits, int splits, float* red, void* stream);
/* dgamma[c] += sum_r red[r][c][1], dbeta[c] += sum_r red[r][c][0], r over rows = n_img*splits */
int st_gn_bwd_params(const float* red, int rows, int C, float* dgamma, float* dbeta, void* stream);
/* backward, pass 2: dx = rstd*(gamma*dz - mean_g(gamma*dz) - xhat*mean_g(gamma*dz*xhat))
 *                        + extra_scale*extra,  written split over dx1 [..][C1] and dx2 [..][C2];
 * accum1/accum2 != 0 adds into the destination instead of overwriting.
 * `dgamma`/`dbeta` != NULL: the kernel also accumulates the parameter gradients from `red` (st_gn_bwd_params is
 * then not needed). */
int st_gn_bwd_apply(const void* x1, const void* x2, const void* dy, int dtype, int n_img, int hw, int C1,
                    int C2, int G, const float* gamma, const float* beta, const float* mean,
                    const float* rstd, int act, float p_drop, uint64_t seed, const void* mask,
                    const uint8_t* keepbits, int splits, const float* red, const void* extra, float extra_scale, void* dx1, int accum1,
                    void* dx2, int accum2, int chunks, float* csum, float* dgamma, float* dbeta, void* stream);
/* pixel chunks per image the backward-apply launch uses by default (grid.x); with `csum` != NULL the caller passes
 * this count explicitly and receives csum[n_img][chunks][C]: the column sums of the gradient contribution the
 * kernel produced (extra included, accumulated destination excluded). */
int st_gn_chunks(int n_img, int hw, int C);
/* Both backward passes in ONE launch: a thread-block cluster of `chunks` CTAs owns an image, reduces its pixel chunks
 * into red[n_img][chunks][C][2], synchronises (barrier.cluster) and produces dx: one launch and no
 * stream-ordered round trip of `red` (the second read of x / dy hits L2 only for what the resident wave leaves
 * there).  csum (optional) is [n_img][chunks][C].  The parameter gradients are left to the caller (st_gn_bwd_params or a kind-2 st_colsum_batched job over `red`, rows = n_img*chunks).
 * st_gn_bwd_fused_chunks returns the cluster size the library would use for this shape (`streams` = 2 + extra + an
 * accumulated destination), 0 = use the two-pass form (measured slower there, too few CTAs, or ST_GN_FUSED=0). */
int st_gn_bwd_fused_chunks(int n_img, int hw, int C, int dtype, int streams);
int st_gn_bwd_fused(const void* x1, const void* x2, const void* dy, int dtype, int n_img, int hw, int C1,
                    int C2, int G, const float* gamma, const float* beta, const float* mean,
                    const float* rstd, int act, float p_drop, uint64_t seed, const void* mask,
                    const uint8_t* keepbits, int chunks, float* red, const void* extra, float extra_scale,
                    void* dx1, int accum1, void* dx2, int accum2, float* csum, void* stream);

/* The backward with plain (not accumulated) destinations and x, dy (and `extra`, optional) read ONCE: a cluster of
 * `chunks` CTAs keeps the image resident in shared memory between the reduction and the apply phase (dz overwrites dy's copy), exchanges
 * the per-group sums through distributed shared memory and writes dx: 3 tensor passes over HBM instead of 5.  `red`
 * [n_img][chunks][C][2] and csum [n_img][chunks][C] as for st_gn_bwd_fused.  st_gn_bwd_resident_chunks: cluster size
 * for a shape and stream count (2, or 3 with `extra`), 0 = not applicable (an accumulated destination, C > 512, more
 * than 8 CTAs x 64 KB per image, too few CTAs, or ST_GN_BWD_RESIDENT=0). */
int st_gn_bwd_resident_chunks(int n_img, int hw, int C, int streams);
int st_gn_bwd_resident(const void* x1, const void* x2, const void* dy, int dtype, int n_img, int hw, int C1,
                       int C2, int G, const float* gamma, const float* beta, const float* mean,
                       const float* rstd, int act, float p_drop, uint64_t seed, const void* mask,
                       const uint8_t* keepbits, int chunks, float* red, const void* extra, float extra_scale,
                       void* dx1, void* dx2, float* csum, void* stream);

/* Both backward passes in ONE persistent launch whose phases chase each other through L2 (groupnorm_wave.cu): the work
 * items of the reduction pass and of the apply pass come from one queue, ordered so that the apply items of a group of
 * images (x + dy of a group = a fraction of L2) run while the reduction items of the next group stream from HBM; an
 * apply item waits on a per-image counter for the reduction items of its image.  HBM sees x and dy once and dx once
 * (3 passes instead of 5), without the load -> barrier -> store chain of the cluster-resident form.  Arguments as for
 * st_gn_bwd_fused, plus
 *   chunks, group  the plan st_gn_bwd_wave_plan returns (pixel chunks per image; images per group via *group_out);
 *                  chunks = 0: the tensor is too small to gain (or ST_GN_WAVE=0) - use the other forms
 *   red            [n_img][chunks][C][2] (required), csum [n_img][chunks][C] (optional)
 *   work           int32 [n_img + 2] counters, ZERO on entry; the kernel leaves them zero again.  reset != 0: the
 *                  library zeroes them first (cudaMemsetAsync on the stream). */
int st_gn_bwd_wave_plan(int n_img, int hw, int C, int dtype, int* group_out);
int st_gn_bwd_wave(const void* x1, const void* x2, const void* dy, int dtype, int n_img, int hw, int C1, int C2, int G,
                   const float* gamma, const float* beta, const float* mean, const float* rstd, int act, float p_drop,
                   uint64_t seed, const void* mask, const uint8_t* keepbits, int chunks, int group, float* red,
                   const void* extra, float extra_scale, void* dx1, int accum1, void* dx2, int accum2, float* csum,
                   int* work, int reset, void* stream);

/* The backward behind a data-gradient GEMM that already produced dz and its quad sums (st_gemm_args.dz_x):
 * st_gn_bwd_consts writes the per-(image, channel) table that epilogue reads, cst[n_img][C][4] =
 * (rstd*gamma, beta - mean*rstd*gamma, gamma, beta); st_gn_bwd_dz_apply is the ONE streaming pass that is left:
 *   dx = rstd*(gamma*dz - mean_g(gamma*dz) - xhat*mean_g(gamma*dz*xhat)) + extra_scale*extra  (split / accumulated as
 * in st_gn_bwd_apply), with qpart [n_img*hw/32][C/4][2] the sums the GEMM emitted.  It also leaves
 * red[n_img][chunks][C][2] = (sum dz, sum dz*xhat) per pixel chunk (parameter gradients via st_gn_bwd_params or a
 * batched column sum; NULL = not wanted) and csum[n_img][chunks][C] (optional) as st_gn_bwd_apply does.
 * hw must be a multiple of 32, C <= 1024. */
int st_gn_bwd_consts(const float* gamma, const float* beta, const float* mean, const float* rstd, int n_img, int C,
                     int G, float* cst, void* stream);
int st_gn_bwd_dz_apply(const void* x1, const void* x2, const void* dz, int dtype, int n_img, int hw, int C1, int C2,
                       int G, const float* gamma, const float* mean, const float* rstd, const float* qpart,
                       const void* extra, float extra_scale, void* dx1, int accum1, void* dx2, int accum2, int chunks,
                       float* red, float* csum, void* stream);

/* ------------------------------------------------------------------ training-batch preparation
 * Replaces the float pipeline of datasets.py:56-62,117,311-326 + run_lib.py:73-75: src uint8 [n_img][H][W][C] ->
 * dst fp32 [n_img][C][H][W] = a * deq(flip(src)/255) + b, deq(x) = (255 x + u)/256 when dequant == 1.  `flip`
 * (optional) holds one byte per image (non-zero = mirror left-right); `u` (optional, fp32 NCHW in [0,1)) replaces the
 * in-kernel generator keyed by (seed, element). */
int st_prep_batch(const uint8_t* src, const float* u, const uint8_t* flip, float* dst, int n_img, int C, int H,
                  int W, int dequant, uint64_t seed, float a, float b, void* stream);

/* ------------------------------------------------------------------ elementwise / small
 * All take element counts; pointers must be 16-byte aligned. */
int st_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);
/* out = alpha*a + beta*b (b may be NULL) */
int st_axpby(const void* a, const void* b, void* out, int dtype, float alpha, float beta, int64_t n,
             void* stream);
/* y = x*sigmoid(x);  dx = dy * silu'(x) */
int st_silu(const void* x, void* y, int dtype, int64_t n, void* stream);
int st_silu_bwd(const void* x, const void* dy, void* dx, int dtype, int64_t n, void* stream);
/* 2x nearest replicate (dir=+1: [n][H][W][C] -> [n][2H][2W][C]) or 2x2 box sum (dir=-1), times scale.
 * The input may be a channel concatenation of two tensors. */
int st_resample2x(const void* x1, const void* x2, void* y, int dtype, int n_img, int H, int W, int C1,
                  int C2, int dir, float scale, void* stream);
/* out[g][c] (+)= scale * sum over rows r in group g of x[r][c];  x is [groups*rows_per_group] rows of C columns,
 * consecutive rows ld elements apart */
int st_colsum(const void* x, int dtype, int64_t groups, int64_t rows_per_group, int C, int64_t ld, float scale,
              float* out, int accumulate, void* stream);
/* dst[off + (ci*taps + taps-1-t)*Co + co] = src[off + (co*taps + t)*Ci + ci] for every convolution weight
 * [Co][taps][Ci] listed in `table` (n_entries rows of {off, Co, taps, Ci}, device memory; `tile_prefix[e]` = number of
 * 32x32 tiles of the entries before e, total_tiles = their sum): the data gradient of a convolution
 * (torch conv2d backward w.r.t. input, reference models/layers.py ddpm_conv3x3 via autograd) is then a forward
 * convolution over dY with these weights. */
int st_transpose_conv_weights(const void* src, void* dst, int dtype, const int64_t* table, const int64_t* tile_prefix,
                              int n_entries, int64_t total_tiles, void* stream);
/* Many small fp32 column sums in ONE launch.  `jobs` is a device table of n_jobs records of 21 eight-byte fields:
 *   part[4] (device pointers), rows[4], ld[4] (elements), dst, n_parts, kind, C, groups, rows_per_group, ld_out,
 *   scale (double), accumulate.
 * kind 0: dst[c] (+)= scale * sum over parts p < n_parts and rows r < rows[p] of part[p][r*ld[p] + c]
 * kind 1: dst[g*ld_out + c] (+)= scale * sum over k < rows_per_group of part[0][(g*rows_per_group + k)*ld[0] + c], g < groups
 * C % 4 == 0, all pointers 16-byte aligned; blocks_y >= max over jobs of ceil(C/128) (kind 0) / ceil(groups*C/4096)
 * (kind 1).  Jobs must have distinct destinations.  Replaces the per-tensor bias reductions of a backward pass. */
int st_colsum_batched(const void* jobs, int n_jobs, int blocks_y, void* stream);
/* row softmax of scale*logits: logits fp32 [rows][L] -> p (dtype) */
int st_softmax_fwd(const float* logits, void* p, int dtype, int64_t rows, int L, float scale, void* stream);
/* ds = scale * p * (dp - sum_j dp*p) ; dp fp32, ds dtype */
int st_softmax_bwd(const void* p, const float* dp, void* ds, int dtype, int64_t rows, int L, float scale,
                   void* stream);
/* sinusoidal embedding: out[b] = [sin(labels[b]*freqs), cos(labels[b]*freqs)], freqs[dim/2], out[B][dim] (fp32) */
int st_timestep_embedding(const float* labels, const float* freqs, float* out, int B, int dim, void* stream);
/* [sin(2 pi W x), cos(2 pi W x)] with x = log(sigma): out[B][2*nW] */
int st_fourier_embedding(const float* sigma, const float* W, float* out, int B, int nW, void* stream);
/* NCHW fp32 [n][C][H][W] -> NHWC dtype [n][H][W][Cpad] (channels >= C zero-filled), y = alpha*x + beta;
 * and back: the first C of Cpad channels, optionally times row_scale[n]. */
int st_nchw_to_nhwc(const float* x, void* y, int dtype, int n_img, int C, int H, int W, int Cpad, float alpha,
                    float beta, void* stream);
int st_nhwc_to_nchw(const void* x, int dtype, float* y, int n_img, int C, int H, int W, int Cpad,
                    const float* row_scale, void* stream);
/* bf16 im2col of a small-channel NHWC tensor: out[pixel][Kpad] (zero padded), k = tap*C + c */
int st_im2col_small(const void* x, int dtype, void* out, int n_img, int H, int W, int C, int kh, int kw,
                    int Kpad, void* stream);

/* Strided im2col of an NHWC tensor and its adjoint (the stride-2 convolutions of the input pyramid,
 * models/up_or_down_sampling.py:144-178, models/layerspp.py:142-176):
 * cols[(n,oy,ox)][kh*kw][C], element = x[n][oy*stride + r - pad][ox*stride + q - pad][c] or 0. */
int st_im2col(const void* x, void* cols, int dtype, int n_img, int H, int W, int C, int kh, int kw, int stride,
              int pad, int OH, int OW, void* stream);
int st_col2im(const void* dcols, void* dx, int dtype, int n_img, int H, int W, int C, int kh, int kw, int stride,
              int pad, int OH, int OW, void* stream);

/* ------------------------------------------------------------------ reference native ops */
/* x [major][in_h][in_w][minor] -> y [major][out_h][out_w][minor]; k fp32 [kh][kw]
 * (reference op/upfirdn2d.cpp:12-19; out_h = (in_h*up_y + pad_y0 + pad_y1 - kh + down_y)/down_y). */
int st_upfirdn2d(const void* x, void* y, int dtype, const float* k, int major, int in_h, int in_w, int minor,
                 int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1,
                 int pad_y0, int pad_y1, void* stream);
/* y = act(x + b[(i/step_b) % size_b]) * scale (reference op/fused_bias_act_kernel.cu:18-49);
 * act 1 linear / 3 leaky-relu; grad 0/1/2; b and ref may be NULL. */
int st_fused_bias_act(const void* x, const void* b, const void* ref, void* y, int dtype, int64_t n,
                      int size_b, int step_b, int act, int grad, float alpha, float scale, void* stream);

/* ------------------------------------------------------------------ loss head */
/* x_t[n][i] = mean_coeff[n]*x0[n][i] + std[n]*z[n][i]   (all fp32 NCHW-flat [B][D]) */
int st_dsm_perturb(const float* x0, const float* z, const float* mean_coeff, const float* std, float* xt,
                   int B, int64_t D, void* stream);
/* per-sample loss[n] = w[n] * red( (a[n]*out[n][i] + b[n]*z[n][i])^2 ), red = mean (reduce_mean=1) or
 * 0.5*sum.  When dout != NULL also dout[n][i] = gvec[n] * w[n]*red'*2*(a out + b z)*a, i.e. the
 * gradient w.r.t. `out` for upstream per-sample gradients gvec.  `out` is the raw network output,
 * fp32 [B][D]. */
int st_dsm_loss(const float* out, const float* z, const float* a, const float* b, const float* w, float* loss,
                float* dout, const float* gvec, int B, int64_t D, int reduce_mean, void* stream);

/* ------------------------------------------------------------------ optimizer */
/* acc[0] += sum x^2 (double accumulation across blocks via fp32 partials + one atomic per block) */
int st_sumsq(const float* x, int64_t n, float* acc, void* stream);
/* Fused clip + Adam (+ weight decay) + EMA over flat fp32 buffers.
 *   coef = min(1, clip / (sqrt(*gnorm_sq) + 1e-6)) if clip >= 0 and gnorm_sq != NULL else 1
 *   g = coef*grad (+ wd*p); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2
 *   p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps);  ema -= (1-decay)*(ema - p)   (ema_mask[i]==0 skips)
 *   p16 (optional) receives bf16(p).
 * `dyn` (optional, device): {lr, bc1, bc2, ema_decay} read by the kernel INSTEAD of the by-value arguments, so that a
 * captured CUDA graph of the training step replays with the current step's warm-up / bias-correction / EMA scalars. */
int st_adam_ema(float* p, const float* grad, float* m, float* v, float* ema, const uint8_t* ema_mask,
                void* p16, int64_t n, const float* gnorm_sq, float clip, float lr, float b1, float b2,
                float eps, float wd, float bc1, float bc2, float ema_decay, const float* dyn, void* stream);

/* ------------------------------------------------------------------ sampler */
/* x_mean = ca[n]*x + cb[n]*s ; x_new = x_mean + cc[n]*noise   (fp32 [B][D]); noise may be NULL.
 * Covers Euler-Maruyama, reverse diffusion, Langevin and the denoise step with host/device-computed
 * per-sample coefficients (sampling.py:190-196,205-210,283-290). */
int st_pc_update(const float* x, const float* s, const float* noise, const float* ca, const float* cb,
                 const float* cc, float* x_mean, float* x_new, int B, int64_t D, void* stream);
/* out[0] = mean_n ||a[n]||_2 , out[1] = mean_n ||b[n]||_2  (b may be NULL) */
int st_batch_norms(const float* a, const float* b, float* out, int B, int64_t D, void* stream);
/* Langevin step size from the two norms, on device: step[n] = (snr*out[1]/out[0])^2 * 2 * alpha[n];
 * writes cb[n] = step[n], cc[n] = sqrt(2*step[n]), ca[n] = 1. */
int st_langevin_coeffs(const float* norms, const float* alpha, float snr, float* ca, float* cb, float* cc,
                       int B, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* ST_B200_H_ */
