/* nfk.h — C-ABI of libnfk.so, the B200 (sm_100a) implementation of the normalizing-flow
 * transform hot path of sherryli59/NormalizingFlow.
 *
 * The reference exposes no FFI; its boundary for this path is the Python class protocol
 *   layer.forward(x) -> (z, log_det),  layer.inverse(z) -> (x, log_det)
 * (reference nf/models.py:16-18, :25-27).  Each entry point below replaces the ATen op
 * chain behind one of those methods; the citation on every function names it.
 *
 * Conventions (all functions)
 *   - return 0 on success, non-zero NFK_E* otherwise; never throw/abort across the ABI;
 *     nfk_last_error() returns a thread-local message for the last failure;
 *   - every pointer is a DEVICE pointer on the device that owns `stream`, 16-byte aligned
 *     base, row-major contiguous fp32 unless stated otherwise; the caller allocates and
 *     owns all buffers, the callee never allocates, frees or synchronises;
 *   - launches are asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - `mask` arguments are HOST pointers to small int32 arrays;
 *   - sm_100a only: there is no CPU path and no other-arch path.
 */
#ifndef NFK_H_
#define NFK_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NFK_ABI_VERSION 2

enum {
  NFK_OK = 0,
  NFK_EINVAL = 1,   /* bad shape / argument (Python wrapper raises ValueError, cf. nf/utils.py:64-71) */
  NFK_ECUDA = 2,    /* a CUDA runtime call failed */
  NFK_EUNSUPPORTED = 3
};

/* arithmetic flavour of the spline kernels */
enum {
  NFK_ARITH_EXACT = 0,  /* op-for-op the rounding sequence of the reference's ATen CUDA chain:
                           bins, outputs and per-element log-dets bit-identical to it */
  NFK_ARITH_HYBRID = 1, /* searched-side knot chain EXACT (bins bit-identical), the rest
                           FMA-contracted with MUFU ex2/lg2/rcp (a few ulp) -- the default */
  NFK_ARITH_FAST = 2    /* MUFU ex2/lg2/rcp everywhere (a few ulp).  The stand-alone transform kernels
                           (nfk_rqs_coupling, nfk_unconstrained_rqs, nfk_rqs_elementwise) still return the
                           reference's bin: an input within 1.2e-4*(B/3) of a fast-chain knot has its bin
                           re-decided on the exact chain.  So do the fused layer kernels
                           (nfk_nsf_pairs_fused, nfk_gemm_ws_rqs), on the parameters they compute. */
};

/* 16-bit element format of the tensor-core operand images of the wide conditioner path */
enum {
  NFK_IMG_BF16 = 0, /* bfloat16: 8-bit significand, fp32's range -- gradient images need the range, so the
                       training / log-prob-gradient path keeps every image in bf16 */
  NFK_IMG_F16 = 1   /* IEEE fp16: 11-bit significand (8x less quantisation noise at the same tensor-core
                       rate); values saturate at +-65504.  Inference forward: activations are tanh outputs
                       in (-1, 1), weights O(1) */
};

int nfk_abi_version(void);
const char* nfk_last_error(void);
/* number of kernels launched by this library since load (bench.py "gpu_launches") */
int64_t nfk_launch_count(void);
/* association order of the EXACT cumsum (torch.cumsum on CUDA, nf/utils.py:75):
 * 0 sequential, 1 Sklansky, 2 up-sweep/down-sweep (ATen ScanUtils.cuh).  Process-global. */
int nfk_set_scan_order(int order);
int nfk_get_scan_order(void);
/* override the tile geometry of nfk_rqs_coupling (0 = automatic): rows per tile (multiple of
 * 4), threads per CTA (multiple of 32, <= 640), pipeline stages, CTAs per SM.  For tuning. */
int nfk_set_tuning(int rows_per_tile, int threads, int stages, int ctas_per_sm);
/* test hook: knots [M, K+1] of one side from logits [M, K] (layer_norm != 0: apply the
 * layer's 2B*softmax first).  nf/flows.py:233-234 + nf/utils.py:73-79. */
int nfk_debug_knots(const float* logits, float* knots, int64_t M, int K, float B, int layer_norm,
                    int exact, void* stream);

/* ---- RQS coupling: replaces NSF_CL.forward/inverse after `psi`
 *      (nf/flows.py:232-239, :246-253) + unconstrained_RQS/RQS/searchsorted (nf/utils.py:20-152).
 * x [N, size*dim]; params [N, F_t, 3K-1] raw conditioner output, F_t = size*(dim-n_mask);
 * out [N, size*dim] in the reference's column order (conditioning columns first inside each
 * dim-group); logdet [N] (overwritten, or += when accumulate != 0); bins [N, F_t] int8, may
 * be NULL (-1 marks the identity tails). */
int nfk_rqs_coupling(const float* x, const float* params, float* out, float* logdet,
                     int8_t* bins, int64_t N, int size, int dim, const int32_t* mask,
                     int n_mask, int K, float B, int inverse, int accumulate, int arith,
                     void* stream);

/* backward of the above: given grad_out [N, size*dim] and grad_logdet [N] (may be NULL = 0)
 * writes grad_x [N, size*dim] (w.r.t. ALL input columns, excluding the path through the
 * conditioner) and grad_params [N, F_t, 3K-1].  Replaces autograd through the same chain. */
int nfk_rqs_coupling_bwd(const float* x, const float* params, const float* grad_out,
                         const float* grad_logdet, float* grad_x, float* grad_params,
                         int64_t N, int size, int dim, const int32_t* mask, int n_mask,
                         int K, float B, int inverse, void* stream);

/* ---- free-function spline: replaces unconstrained_RQS (nf/utils.py:27-56) on
 * inputs [M], W,H [M,K], D [M,K-1] (already layer-normalised once) -> out [M], lad [M]. */
int nfk_unconstrained_rqs(const float* inputs, const float* W, const float* H, const float* D,
                          float* out, float* lad, int8_t* bins, int64_t M, int K, float B,
                          int inverse, int arith, void* stream);

/* ---- element-wise spline with the layer-side normalisation: the transform step of the
 * autoregressive spline flow NSF_AR (nf/flows.py:178-190 forward, :196-208 inverse): softmax x 2B on
 * W,H and softplus on D (flows.py:183-185), then unconstrained_RQS (nf/utils.py:27-152).
 * inputs [M], params [M, 3K-1] raw conditioner outputs -> out [M], lad [M], bins [M] (nullable). */
int nfk_rqs_elementwise(const float* inputs, const float* params, float* out, float* lad,
                        int8_t* bins, int64_t M, int K, float B, int inverse, int arith,
                        void* stream);
/* backward: grad_out [M], grad_lad [M] (nullable = 0) -> grad_in [M], grad_params [M, 3K-1] */
int nfk_rqs_elementwise_bwd(const float* inputs, const float* params, const float* grad_out,
                            const float* grad_lad, float* grad_in, float* grad_params, int64_t M,
                            int K, float B, int inverse, void* stream);

/* ---- affine half-coupling: y = t + v*exp(s) (forward) or (v - t)*exp(-s) (inverse),
 * logdet (+)= +-sum_j s.  Replaces nf/flows.py:56, :59, :61-62 and :69, :72, :74-75.
 * v is read from x[:, v_off : v_off+h] (row stride ld_x), s,t are [N,h] contiguous, y is
 * written to out[:, y_off : y_off+h] (row stride ld_out). */
int nfk_affine_halfcoupling(const float* x, int64_t ld_x, int v_off, const float* s,
                            const float* t, float* out, int64_t ld_out, int y_off,
                            float* logdet, int64_t N, int h, int inverse, int accumulate,
                            void* stream);
int nfk_affine_halfcoupling_bwd(const float* x, int64_t ld_x, int v_off, const float* s,
                                const float* t, const float* grad_y, int64_t ld_gy, int gy_off,
                                const float* grad_logdet, float* grad_v, int64_t ld_gv,
                                int gv_off, float* grad_s, float* grad_t, int64_t N, int h,
                                int inverse, void* stream);

/* ---- planar stack: L tanh planar layers applied in sequence in ONE pass over x.
 * Replaces L calls of Planar.forward (nf/flows_1.py:42-60).  w,u [L,d], b [L];
 * nfk_planar_prepare computes the parameter-only terms uhat [L,d] (flows_1.py:51-53) and
 * wuhat [L] = w_l . uhat_l once; nfk_planar_stack then writes out [N,d] and
 * logdet [N] (+)= sum over layers of log(|1 + (1-tanh^2)(w.uhat)| + 1e-4). */
int nfk_planar_prepare(const float* w, const float* u, float* uhat, float* wuhat, int d, int L,
                       void* stream);
int nfk_planar_stack(const float* x, const float* w, const float* uhat, const float* wuhat,
                     const float* b, float* out, float* logdet, int64_t N, int d, int L,
                     int accumulate, void* stream);
/* The same stack at the HBM rate for d = 128, L <= 32 (csrc/planar_mma.cu): the stack is rewritten as two
 * small dense products around a per-row scalar recurrence over the Gram matrix gram[l][m] = w_l . uhat_m
 * ([L, L], from nfk_planar_gram), products on the tensor cores with fp16 hi + lo operand pairs (fp32-class). */
int nfk_planar_gram(const float* w, const float* uhat, float* gram, int d, int L, void* stream);
int nfk_planar_stack_mma(const float* x, const float* w, const float* uhat, const float* gram,
                         const float* b, float* out, float* logdet, int64_t N, int d, int L,
                         int accumulate, void* stream);
/* backward: recomputes the forward layer by layer; grad_w/grad_uhat [L,d], grad_b/grad_wuhat
 * [L] are ACCUMULATED with atomics into zero-initialised buffers (the chain from uhat, wuhat to
 * u, w is parameter-sized and stays in the host wrapper). */
int nfk_planar_stack_bwd(const float* x, const float* w, const float* uhat, const float* wuhat,
                         const float* b, const float* grad_out, const float* grad_logdet,
                         float* grad_x, float* grad_w, float* grad_uhat, float* grad_b,
                         float* grad_wuhat, int64_t N, int d, int L, void* stream);

/* ---- radial layer (nf/flows_1.py:85-97).
 * per_sample == 0: reference behaviour, r = ||x - x0||_F over the WHOLE batch; the caller
 *   provides sumsq (device scalar, sum of squares, already all-reduced across ranks if the
 *   batch is sharded) computed with nfk_radial_sumsq; logdet gets ONE value (logdet[0]).
 * per_sample != 0: r is the per-row norm and logdet is [N]. */
int nfk_radial_sumsq(const float* x, const float* x0, float* sumsq /*zeroed*/, int64_t N, int d,
                     void* stream);
int nfk_radial(const float* x, const float* x0, const float* log_alpha, const float* beta,
               const float* sumsq, float* out, float* logdet, int64_t N, int d, int per_sample,
               int accumulate, void* stream);

/* A run of L PER-SAMPLE radial layers in one pass (csrc/radial_stack.cu): x0 [L,d], log_alpha [L], beta [L];
 * out [N,d], logdet [N] (+)= sum over the layers.  d in {32, 64, 128, 256}. */
int nfk_radial_stack(const float* x, const float* x0, const float* log_alpha, const float* beta,
                     float* out, float* logdet, int64_t N, int d, int L, int accumulate, void* stream);
/* Batch-global mode, one fused pass: when x0 != NULL applies the layer (out = x + beta h (x - x0) with
 * h = 1 / (alpha + sqrt(sumsq[0])), logdet[0] (+)= the layer's single log-det) and, when x0_next != NULL,
 * accumulates sum (value - x0_next)^2 of what it just wrote (of x itself when x0 == NULL) into the zeroed
 * device scalar sumsq_next -- the second pass of layer l is the first pass of layer l+1.  Across ranks the
 * caller all-reduces sumsq_next between launches.  d must be a multiple of 4 that divides 1024. */
int nfk_radial_global(const float* x, const float* x0, const float* log_alpha, const float* beta,
                      const float* sumsq, const float* x0_next, float* sumsq_next, float* out,
                      float* logdet, int64_t N, int d, int accumulate, void* stream);

/* backward of the radial layer: grad_x [N,d]; grad_x0 [d], grad_log_alpha [1], grad_beta [1] are
 * ACCUMULATED with atomics into zero-initialised buffers.  Batch-global mode needs sumsq (as the
 * forward) and dot = sum(grad_out * (x - x0)) from nfk_radial_dot (all-reduced when sharded). */
int nfk_radial_dot(const float* x, const float* x0, const float* grad_out, float* dot /*zeroed*/,
                   int64_t N, int d, void* stream);
int nfk_radial_bwd(const float* x, const float* x0, const float* log_alpha, const float* beta,
                   const float* sumsq, const float* dot, const float* grad_out,
                   const float* grad_logdet, float* grad_x, float* grad_x0, float* grad_log_alpha,
                   float* grad_beta, int64_t N, int d, int per_sample, void* stream);

/* ---- log-prob reduction: out[n] = -0.5*sum_j z[n,j]^2/var - 0.5*d*log(2*pi*var) (+ add[n]).
 * Replaces prior.log_prob + the additions at nf/models.py:19-20, :34, :39. */
int nfk_gauss_logprob(const float* z, const float* add /*nullable*/, float add_sign, float* out,
                      int64_t N, int d, float var, void* stream);

/* ---- conditioner MLP layer: Y = act(X W^T + b).  Replaces one nn.Linear (+Tanh) of FCNN
 * (nf/flows.py:26-35).  X [M, K] with row stride ldx (lets the caller pass a strided column
 * gather), W [Nout, K] row-major (nn.Linear layout), b [Nout], Y [M, Nout] contiguous.
 * act: 0 = identity, 1 = tanh.  fp32 CUDA-core kernel (parity mode). */
int nfk_linear_f32(const float* X, int64_t ldx, const float* W, const float* b, float* Y,
                   int64_t M, int K, int Nout, int act, void* stream);
/* The same layer on the tensor cores with fp32-class accuracy (3xTF32: x = hi + lo, hi*hi + hi*lo + lo*hi
 * accumulated in fp32 in TMEM, tcgen05.mma.kind::tf32): the parity mode's forward GEMM wherever
 * K % 4 == 0 and the rows are 16-byte aligned.  X [M, K] row stride ldx, W [Nout, K] row stride ldw. */
int nfk_linear_tf32x3(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* b, float* Y,
                      int64_t ldy, int64_t M, int K, int Nout, int act, void* stream);
/* generic fp32 GEMM used by the conditioner backward: C[M,N] (+)= op(A) op(B),
 * ta/tb: 0 = as stored row-major [M,K]/[K,N], 1 = stored transposed ([K,M]/[N,K]). */
int nfk_gemm_f32(const float* A, int64_t lda, int ta, const float* Bm, int64_t ldb, int tb,
                 float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate,
                 void* stream);

/* bf16 tensor-core (tcgen05 + TMEM + TMA) layer: X,W bf16, fp32 accumulate, bias fp32,
 * Y bf16 (out_f32 == 0) or fp32 (out_f32 != 0).  K % 16 == 0 (pad), rows/cols arbitrary. */
int nfk_linear_bf16(const void* X, int64_t ldx, const void* W, int64_t ldw, const float* b,
                    void* Y, int64_t ldy, int64_t M, int K, int Nout, int act, int out_f32,
                    void* stream);

/* ---- fused NSF coupling layer: conditioner MLP (tcgen05 bf16 GEMMs, fp32 accumulation in TMEM)
 * with the RQS transform as the epilogue of its last GEMM -- replaces NSF_CL.forward/inverse
 * (nf/flows.py:227-253) in one launch for size = 32, dim = 2, one masked column (mask_col),
 * K = 8, hidden width <= 128.  x, out [N, 64] fp32, N a multiple of
 * nfk_nsf_fused_rows_per_tile(); logdet [N] (+= when accumulate).  w1_img / w2_img / w3_img are
 * the bf16 weights padded to 128 x 64, 128 x 128 and 8 chunks of 96 x 128 (each feature's 23
 * rows padded to 24) in the K-major SWIZZLE_128B shared-memory layout (16-byte chunk j of row r
 * holds source chunk j ^ (r % 8)); b1, b2 [128], b3 [32*24] fp32 padded the same way. */
int nfk_nsf_fused_rows_per_tile(void);
/* In every arithmetic the bin an element uses is the one searchsorted (nf/utils.py:20-25) finds on the
 * spline parameters THIS kernel computed (EXACT / HYBRID search the exact knot chain, FAST re-decides on
 * it next to a knot).  Test hooks (NULL in production): dbg_params [N][32][24] receives those raw
 * parameters (accumulator + b3; the 24th column is padding), dbg_bins [N][32] the bins (-1 = tail).
 * test hook: device buffer of 64 int64 receiving clock64 stamps of CTA 0, third tile (NULL = off) */
int nfk_set_fused_trace(void* dev_buf);
int nfk_nsf_pairs_fused(const float* x, float* out, float* logdet, const void* w1_img,
                        const void* w2_img, const void* w3_img, const float* b1, const float* b2,
                        const float* b3, int64_t N, int mask_col, float B, int inverse,
                        int accumulate, int arith, float* dbg_params /*nullable*/,
                        int8_t* dbg_bins /*nullable*/, void* stream);

/* Second generation of the fused layer kernel (csrc/nsf_fused2.cu), same contract as nfk_nsf_pairs_fused:
 * the layer's phases run on three overlapping sets of warps (spline warps / hidden-layer warps one tile
 * ahead / MMA + TMA warps, registers redistributed with setmaxnreg), W2 and W3 stream through one ring.
 * split == 0: fp16 operands; images as for nfk_nsf_pairs_fused (w2_img = two 128 x 64 K blocks, w3_img = 8
 *   chunks of [2 K blocks][96 x 64]).
 * split != 0: fp32-class conditioner -- every operand is an fp16 pair hi + lo (lo = fp16(v - hi)), every product
 *   three tensor-core MMAs hi*hi + lo*hi + hi*lo with fp32 accumulation: w1_img [128 x 64] carries hi in K
 *   columns 0..31 and lo in 32..63; w2_img = [kb0 hi][kb0 lo][kb1 hi][kb1 lo] (16 KB blocks); w3_img = per
 *   chunk [hi 24 KB][lo 24 KB].  With NFK_ARITH_HYBRID / EXACT this is the strict-parity configuration
 *   (z, log_det to the fp32 gate, bins = the exact search) in ONE launch per layer pass.
 * Replaces NSF_CL.forward/inverse (nf/flows.py:227-253). */
/* consecutive launches of nfk_nsf_pairs_fused2 on one stream overlap their prologue with the previous launch's tail
 * (programmatic dependent launch; the kernel waits for its predecessor before its first global access).
 * on = 0 restores plain stream order.  Process-global; for tuning and tests. */
int nfk_set_fused2_pdl(int on);
int nfk_nsf_pairs_fused2(const float* x, float* out, float* logdet, const void* w1_img,
                         const void* w2_img, const void* w3_img, const float* b1, const float* b2,
                         const float* b3, int64_t N, int mask_col, float B, int inverse,
                         int accumulate, int arith, int split, float* dbg_params /*nullable*/,
                         int8_t* dbg_bins /*nullable*/, void* stream);

/* The same launch as one link of a CHAIN of layer launches over the same rows on one stream (a whole flow, forward or
 * inverse): tile_flags_in / tile_flags_out [N / 128] int32 or NULL make the dependency between consecutive launches
 * per 128-row tile instead of per launch (rows are independent, nf/models.py:16-18).  With tile_flags_out the launch
 * stores tile_flag_epoch into flag t when rows [128 t, 128 t + 128) of out and logdet are complete; with
 * tile_flags_in it does not wait for the previous launch as a whole but takes tile t when flag t >= tile_flag_epoch.
 * Flags only grow: the caller zeroes them once, passes each producer's tile_flags_out as the next launch's
 * tile_flags_in with the same epoch, raises the epoch (from 1) every time the same flag arrays are used again without
 * an intervening stream-wide wait (consecutive evaluations of a trajectory), and passes NULL as tile_flags_in of the
 * first launch (that launch waits for the stream); weight images and biases must be complete before it. */
int nfk_nsf_pairs_fused2_chain(const float* x, float* out, float* logdet, const void* w1_img, const void* w2_img,
                               const void* w3_img, const float* b1, const float* b2, const float* b3, int64_t N,
                               int mask_col, float B, int inverse, int accumulate, int arith, int split,
                               float* dbg_params /*nullable*/, int8_t* dbg_bins /*nullable*/,
                               int32_t* tile_flags_in /*nullable*/, int32_t* tile_flags_out /*nullable*/,
                               int tile_flag_epoch, void* stream);

/* Gradient of one fused layer w.r.t. its input in ONE launch (csrc/nsf_fused_bwd.cu; hidden <= 128, size 32, dim 2,
 * K 8): recomputes the conditioner with the forward kernel's fp16 operands (same parameters, bit for bit), runs the
 * spline adjoint per element in registers, the three dgrad GEMMs (bf16 operands) and the tanh backward on chip.
 * x, grad_out, grad_x [N, 64]; dL/d(out) = grad_out_scale * grad_out in the layer's OUTPUT column order
 * (conditioning, transformed) -- pass z and -1/var to start from an isotropic Gaussian prior; grad_logdet [N] or
 * NULL (grad_logdet_const for every row: 1 for log-prob gradients).  N a multiple of 128.
 * Operand images (all K-major SWIZZLE_128B, built by normalizingflow_b200/_fused.py:packed_bwd): w1_img, w2_img,
 * b1, b2 as for nfk_nsf_pairs_fused2 (split = 0); w3_img and b3 likewise but with every feature's 24 rows in the
 * order (w0, h0, w1, h1, ..., w7, h7, d0 .. d6, 0) so that a width / height logit pair lands in adjacent
 * accumulator columns; transposed bf16 images: w3t_img [4][3][128 x 64] (pair of chunks p, K block: n = hidden
 * unit, k = parameter index within the pair, same order), w2t_img [2][128 x 64] (n = input unit, k = output unit),
 * w1t_img [2][32 x 64] (n = conditioning feature, k = hidden unit).
 * tile_flags_in / tile_flags_out [N / 128] int32 or NULL, tile_flag_epoch >= 1: per-tile dependency between consecutive
 * launches of a chain on one stream (rows are independent), as for nfk_nsf_pairs_fused2_chain: the launch stores the
 * epoch into flag t when rows [128 t, 128 t + 128) of grad_x are complete; with tile_flags_in it does NOT wait for the
 * previous launch as a whole (programmatic dependent launch) but starts tile t when flag t >= epoch.  A tile's x is
 * only touched after its flag, so x may itself be produced inside the chain (the forward launches of the same
 * evaluation).
 * Replaces autograd through NSF_CL.forward / inverse (nf/flows.py:227-253) for dL/dx, as the flow-preconditioned HMC
 * force evaluation needs it (nf/hmc.py:34-41, applications/src/systems.py:308-311). */
int nfk_nsf_pairs_fused_bwd(const float* x, const float* grad_out, float grad_out_scale, const float* grad_logdet,
                            float grad_logdet_const, float* grad_x, const void* w1_img, const void* w2_img,
                            const void* w3_img, const void* w3t_img, const void* w2t_img,
                            const void* w1t_img, const float* b1, const float* b2, const float* b3,
                            int64_t N, int mask_col, float B, int inverse, int32_t* tile_flags_in /*nullable*/,
                            int32_t* tile_flags_out /*nullable*/, int tile_flag_epoch, void* stream);

/* The same launch with the leapfrog update folded in (the launch that completes the force, i.e. the layer nearest the
 * data; reference applications/src/systems.py:331-336 drives kick / drift as separate tensor operations): after
 * grad_x (= the force F when the chain started from the prior's score) of a row is complete,
 *   momentum += kick * F ;  position += drift * momentum   (drift = 0: kick only)
 * for the same row.  momentum, position [N, 64]; position may be x itself (a tile's reads of x are over by then).
 * With tile_flags_out the flag also covers momentum and position, so the first forward launch of the next
 * evaluation can hang on it (nfk_nsf_pairs_fused2_chain with that flag array as tile_flags_in): that launch counts
 * the next evaluation's epoch, which is what tile_flag_out_epoch is for. */
int nfk_nsf_pairs_fused_bwd_leapfrog(const float* x, const float* grad_out, float grad_out_scale,
                                     const float* grad_logdet, float grad_logdet_const, float* grad_x,
                                     const void* w1_img, const void* w2_img, const void* w3_img,
                                     const void* w3t_img, const void* w2t_img, const void* w1t_img,
                                     const float* b1, const float* b2, const float* b3, int64_t N, int mask_col,
                                     float B, int inverse, int32_t* tile_flags_in /*nullable*/,
                                     int32_t* tile_flags_out /*nullable*/, int tile_flag_epoch,
                                     int tile_flag_out_epoch /* stored into tile_flags_out; 0: tile_flag_epoch */,
                                     float* momentum /*nullable*/, float* position /*nullable*/, float kick, float drift,
                                     void* stream);

/* ---- wide conditioner path (hidden width > 128; the class default is 800, nf/flows.py:216):
 * persistent warp-specialised tcgen05 GEMM  Y = act(A W^T + b)  over operands stored in HBM as
 * *shared-memory images*: 128-row x 64-column blocks of 16-bit elements (16 KB; `fmt` = NFK_IMG_BF16 or
 * NFK_IMG_F16, the same for every image of one launch) in the K-major SWIZZLE_128B
 * layout (16-byte chunk j of row r holds source chunk j ^ (r % 8)), moved by 1-D TMA bulk copies.
 *   a_img [ceil(M/128)][KB][128][64] bf16                    activations (rows >= M are zero)
 *   w_img for N tile t of 64*tile_blocks[t] output columns:  [KB][64*tile_blocks[t]][64] bf16,
 *         tiles concatenated; rows/columns beyond the real weight are zero
 *   bias  [64 * sum(tile_blocks)] fp32, zero padded
 *   out   out_f32 == 0: bf16 image [ceil(M/128)][sum(tile_blocks)][128][64] (next layer's a_img)
 *         out_f32 != 0: fp32 rows [M, ldy], first n_out columns written
 * kmma_last = ceil((K - 64*(KB-1)) / 16): tcgen05.mma count in the last K block.
 * act: 0 identity, 1 tanh, 2 tanh BACKWARD: out = (A W^T) * (1 - h^2) with h read from the bf16
 * image `aux` laid out like `out` (the saved forward activation) -- the dgrad step of the
 * conditioner.  `aux` may be NULL otherwise.
 * Replaces one nn.Linear(+Tanh) of FCNN (nf/flows.py:26-35), and autograd through it. */
int nfk_gemm_ws_rows_per_tile(void);
/* CTA-pair mode of the wide-path GEMMs (tcgen05 cta_group::2: two SMs share one 256-row MMA, each
 * holding half of the B tile): -1 automatic (pairs once there are >= 2 M tiles per SM), 0 never,
 * 1 always.  Process-global; for tuning and tests. */
int nfk_set_gemm_ws_pair_mode(int mode);
/* diagnostic: CTA pairs the last pair-mode launch could keep co-resident (cudaOccupancyMaxActiveClusters) */
int nfk_gemm_ws_last_clusters(void);
int nfk_gemm_ws(const void* a_img, const void* w_img, const float* bias, void* out, int64_t M,
                int KB, int kmma_last, const int32_t* tile_blocks /*host*/, int n_tiles, int act,
                int out_f32, int n_out, int64_t ldy, const void* aux, int fmt, void* stream);
/* grouped launch: n_groups independent GEMMs with the same N-tile plan, activation, output kind and
 * row count M run in ONE persistent kernel (work items interleave the groups).  groups_dev is a DEVICE
 * array of n_groups records {a_img, w_img, bias, out (pointers), KB, kmma_last, a_kb, pad (int32)}, each
 * nfk_gemm_ws_group_bytes() bytes.  fp32 row outputs share ldy / n_out (a group's `out` may point at its
 * column offset inside one [M, ldy] tensor).  Used for the dim-1 per-dimension conditioners of NSF_AR
 * (nf/flows.py:167-169, :186). */
int nfk_gemm_ws_group_bytes(void);
int nfk_gemm_ws_grouped(const void* groups_dev, int n_groups, int64_t M, const int32_t* tile_blocks /*host*/,
                        int n_tiles, int act, int out_f32, int n_out, int64_t ldy, int fmt, void* stream);
/* NSF_AR conditioner inputs in one launch (nf/flows.py:172-173, :186): the a_img
 * [ceil(N/128)][ceil(2*dim/64)][128][64] of the interleaved features [cos(pi x_0/B), sin(pi x_0/B),
 * cos(pi x_1/B), ...]; conditioner i reads its first ceil(2i/64) K blocks (group record a_kb =
 * ceil(2*dim/64)) and its weight image carries the columns in the same interleaved order. */
int nfk_nsf_ar_pack(const float* x, void* img, int64_t N, int dim, float B, int fmt, void* stream);
/* last conditioner GEMM with the RQS transform as its epilogue (any size, 2 <= dim <= 4, K = 8,
 * at most 128 transformed features): replaces the third nn.Linear of psi + nf/flows.py:232-239 /
 * :246-253 + nf/utils.py:20-152; the [N, F_t, 23] parameter tensor never reaches HBM.  w_img: each
 * feature's 23 weight rows padded to 24, N tiles of 8 features (192 rows; the last tile zero
 * padded), layout as nfk_gemm_ws; bias [ceil(F_t/8)*8*24] padded the same way.  x, out
 * [M, size*dim] fp32 (out in the reference's column order, quirk Q5); logdet [M] (+= when
 * accumulate); mask is a HOST array of n_mask conditioning columns.  Test hooks (NULL in production):
 * dbg_params [M][8*ceil(F_t/8)][24] / dbg_bins [M][8*ceil(F_t/8)] receive the raw parameters
 * (accumulator + bias) and the bin of every element, as for nfk_nsf_pairs_fused. */
int nfk_gemm_ws_rqs(const void* a_img, const void* w_img, const float* bias, const float* x,
                    float* out, float* logdet, int64_t M, int KB, int kmma_last, int size, int dim,
                    const int32_t* mask, int n_mask, float B, int inverse, int accumulate,
                    int arith, int fmt, float* dbg_params /*nullable*/, int8_t* dbg_bins /*nullable*/,
                    void* stream);
/* BACKWARD twin of nfk_gemm_ws_rqs: recomputes the spline parameters with the same GEMM
 * (a_img = saved last hidden activation, w_img / bias as the forward) and runs the spline's
 * adjoint as the epilogue: grad_x [M, size*dim] receives the direct path (transformed columns
 * through the spline, conditioning columns copied from grad_out in input order) and
 * grad_params_img receives dL/dparams as a bf16 image [ceil(M/128)][3*ceil(F_t/8)][128][64]
 * (per-feature 24-column padding, like w_img's rows) -- the A operand of the dgrad GEMM
 * nfk_gemm_ws(grad_params_img, W3^T image, ..., act = 2, aux = h2 image).  grad_logdet [M] may be
 * NULL: grad_logdet_const is then used for every row (1 for log-prob gradients).
 * Replaces autograd through nf/flows.py:232-239 / :246-253 + nf/utils.py:27-152. */
int nfk_gemm_ws_rqs_bwd(const void* a_img, const void* w_img, const float* bias, const float* x,
                        const float* grad_out, const float* grad_logdet, float grad_logdet_const,
                        float* grad_x, void* grad_params_img, int64_t M, int KB, int kmma_last,
                        int size, int dim, const int32_t* mask, int n_mask, float B, int inverse,
                        void* stream);
/* weight-gradient GEMM: C[p, q] += sum_n A[n, p] * B[n, q] over the batch rows n of two bf16 images
 * (a_img [ceil(M/128)][KBa][128][64], b_img [...][KBb][128][64]); the images are consumed as MN-major
 * tcgen05 operands, split-K over the batch with fp32 atomic accumulation into C [P, ldc] (the caller
 * zero-initialises C).  pad_p != 0: the A column index runs over 24-per-feature padded spline
 * parameters and C has the 23-per-feature rows.  Replaces autograd's weight gradient of one
 * nn.Linear of FCNN (nf/flows.py:26-35). */
int nfk_wgrad_ws(const void* a_img, const void* b_img, float* C, int64_t ldc, int64_t M, int KBa,
                 int KBb, int P, int Q, int pad_p, void* stream);
/* fp32 weight matrix W [n_src_rows, n_src_cols] (row stride ld) -> bf16 w_img for nfk_gemm_ws* with KB
 * K blocks and the given N-tile plan.  transposed != 0 packs W^T (dgrad operands); pad_rows / pad_k
 * != 0: the output-row / K index runs over 24-per-feature padded spline parameters whose source has 23
 * per feature.  Everything outside the source is zero.  One launch per layer and parameter version. */
int nfk_pack_w_img(const float* W, int64_t ld, int n_src_rows, int n_src_cols, void* img, int KB,
                   const int32_t* tile_blocks /*host*/, int n_tiles, int transposed, int pad_rows,
                   int pad_k, int fmt, void* stream);
/* bf16 image [ceil(M/128)][KB][128][64] -> row-major bf16 rows [M, ld], first ncols columns (ncols, ld
 * multiples of 8): hands saved activations / gradient images to the weight-gradient GEMMs */
int nfk_unpack_img_rows(const void* img, void* rows, int64_t M, int KB, int ncols, int64_t ld,
                        void* stream);
/* g[:, s*dim + cols[j]] += dxc[:, s*n_cols + j]: adds the conditioner-input gradient [N, size*n_cols]
 * to the conditioning columns of grad_x (adjoint of the gather at nf/flows.py:230) */
int nfk_scatter_add_cols(float* g, const float* dxc, int64_t N, int size, int dim,
                         const int32_t* cols /*host*/, int n_cols, void* stream);
/* x[:, :, cols].flatten(1) (nf/flows.py:230) -> bf16 a_img [ceil(N/128)][KB][128][64], zero padded */
int nfk_pack_a_img(const float* x, void* img, int64_t N, int size, int dim, const int32_t* cols /*host*/,
                   int n_cols, int KB, int fmt, void* stream);

/* x[:, cols] gather -> dense fp32 or bf16 [N, size*n_cols] (conditioner input, flows.py:230) */
int nfk_gather_cols(const float* x, void* out, int64_t N, int size, int dim,
                    const int32_t* cols, int n_cols, int out_bf16, int64_t ld_out,
                    void* stream);
int nfk_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream);

/* ---- priors / targets evaluated next to the flow (applications/src/systems.py)
 * EinsteinCrystal.log_prob (systems.py:360-366): x [N, natoms*dim], centers [natoms*dim] -> out [N];
 * boxlength <= 0 = no minimum-image wrap; grad_x (nullable) receives d log_prob / dx. */
int nfk_einstein_logprob(const float* x, const float* centers, float* out, float* grad_x, int64_t N,
                         int natoms, int dim, float alpha, float boxlength, void* stream);
/* LJ.potential (systems.py:154-189): pos [N, nparticles, dim] -> out [N]; minimum image when
 * boxlength > 0, cutoff <= 0 = none, shift as in the reference; grad_pos (nullable) receives dU/dpos. */
int nfk_lj_potential(const float* pos, float* out, float* grad_pos, int64_t N, int nparticles, int dim,
                     float boxlength, float epsilon, float sigma, float cutoff, int shift,
                     void* stream);
/* GaussianMixture.log_prob (systems.py:287-292): x [N, npoints, dim], centers [nc, dim], vars [nc]
 * -> out [N] = sum over points of log(mean_c N(x; mu_c, var_c I)). */
int nfk_gmm_logprob(const float* x, const float* centers, const float* vars, float* out, int64_t N,
                    int npoints, int dim, int ncenters, void* stream);

/* ---- free-energy estimators on the device
 * BAR fixed point (applications/src/bar.py:16-67): w_F [T_F], w_R [T_R] work values (fp32, or fp64 when
 * is_f64), fp64 arithmetic, the whole iteration runs in one launch; out[0] = DeltaF, out[1] =
 * iterations used (device doubles). */
int nfk_bar(const void* w_F, const void* w_R, int64_t T_F, int64_t T_R, int is_f64, double DeltaF,
            int max_iter, double rtol, double* out, void* stream);
/* out[c] = logsumexp_r a[r, c] - log(rows): log-mean-exp reweighting (applications/src/test.py:66-68,
 * dynamics.py:36) */
int nfk_log_mean_exp(const float* a, float* out, int64_t rows, int64_t cols, void* stream);

/* ---- leapfrog pieces for flow-preconditioned HMC (hmc.py:34-41 drives
 * simulation.integration_step; the integrator pattern is applications/src/systems.py:331-336).
 * kick-drift: p += 0.5*dt*F ; q += dt*inv_mass*p.   kick: p += 0.5*dt*F. */
int nfk_leapfrog_kick_drift(float* q, float* p, const float* force, int64_t n, float dt,
                            float inv_mass, void* stream);
int nfk_leapfrog_kick(float* p, const float* force, int64_t n, float dt, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NFK_H_ */
