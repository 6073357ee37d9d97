/*
 * tsc_b200.h -- C-ABI of libtsc_b200.so: the sm_100a kernels behind the OS-CNN + feature-level
 * style-transfer training hot path.
 *
 * The reference (BaeHann/feature_level_style_transfer_for_TSC) is pure Python/PyTorch: it has no
 * FFI of its own, its nn.Module classes ARE the boundary.  Each entry point below
 * therefore cites the reference *operation* (file:line under /root/reference) it replaces; the
 * Python host side (feature_level_style_transfer_for_tsc_b200/) binds these with ctypes and keeps
 * the reference's module names, constructor/forward signatures and state_dict keys.
 *
 * Conventions (all functions):
 *   - plain pointers and sizes only; every buffer (incl. workspace) is owned by the caller;
 *     device pointers unless the parameter is documented as HOST;
 *   - no allocation and no host synchronisation: safe to capture in a CUDA graph.  Host-side state
 *     is limited to (a) one-time kernel attribute opt-ins (std::call_once), (b) per-bank geometry /
 *     schedule caches behind a mutex, (c) debug hooks that are inert unless switched on explicitly
 *     (tsc_debug_set_timeline, the TSC_CONV_* environment knobs read once): calls are reentrant
 *     across streams and threads, on ONE device per process (one process per GPU);
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return 0 = OK; < 0 = bad argument / unsupported shape (text via tsc_last_error(), thread
 *     local); > 0 = a cudaError_t from the launch.  Nothing throws, nothing aborts;
 *   - there is no CPU fallback: without a CUDA device every launch returns a cudaError_t.
 *
 * Device data layout "c8": a logical [B, C, L] activation is stored as [B][Cp/8][L][8] with
 * Cp = tsc_pad_channels(C) (multiple of 16); pad channels hold zeros.  One (b, chunk) slab is
 * L rows of 8 channels (16 B in bf16, 32 B in fp32).  It is K-major for the forward/dgrad
 * implicit GEMM (contraction over channels) and MN-major for wgrad / Gram (contraction over
 * positions), and a convolution tap is a pure row offset in it.
 */
#ifndef TSC_B200_H
#define TSC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSC_VERSION 100
#define TSC_MAX_TAPS 96          /* largest supported Kmax (reference caps it at 89, train_and_test.py:40) */
#define TSC_MAX_CHANNELS 256     /* largest padded channel count of one OS layer = one TMEM accumulator tile.  The reference's
                                  * layer recipe (train_and_test.py:38-53) gives 228 at L >= 124 but MORE for short series
                                  * (336 at L = 64, 560 at L = 32): the tcgen05 entry points reject such layers; they are
                                  * served by the fp32 CUDA-core engine up to TSC_MAX_CHANNELS_WIDE */
#define TSC_MAX_CHANNELS_WIDE 2048 /* channel limit of the fp32 CUDA-core engine (TSC_ENGINE_SIMT), of the packing / layout / BatchNorm
                                  * kernels and of tsc_oswgrad on that engine: layers wider than one TMEM tile -- the reference's
                                  * recipe for series shorter than ~80 samples -- run there (the Python modules switch such a
                                  * stack to it) */
#define TSC_MAX_OPT_GROUPS 32    /* parameter groups of one tsc_rmsprop_step call */
#define TSC_MAX_LIST 32          /* tensors of one tsc_multi_l2norm call */
#define TSC_MAX_CLASSES 64       /* classes of the voting kernels */
#define TSC_MAX_VOTERS 8         /* models of one tsc_entropy_vote call */

typedef void* tsc_stream_t;      /* cudaStream_t */

enum tsc_dtype { TSC_F32 = 0, TSC_BF16 = 1 };
/* engine: SIMT = fp32 CUDA-core arithmetic (the "fp32, <=1e-5" mode; also accepts bf16 operands,
 * which makes it the bit-faithful checker of the tensor-core engine); TCGEN05 = bf16 operands,
 * fp32 accumulation in TMEM (the "<=1e-2" mode). */
enum tsc_engine { TSC_ENGINE_SIMT = 0, TSC_ENGINE_TCGEN05 = 1 };
enum tsc_direction { TSC_DIR_FWD = 0, TSC_DIR_DGRAD = 1 };
/* TSC_OUT_POOLED (fused BatchNorm apply only): out[B, C] = mean over L of the activated output -- the
 * AdaptiveAvgPool1d(1) + squeeze of the classifier (OS_CNN.py:93,105) without materialising [B, C, L] */
enum tsc_out_kind { TSC_OUT_C8_F32 = 0, TSC_OUT_C8_BF16 = 1, TSC_OUT_NCL_F32 = 2, TSC_OUT_POOLED = 3 };

int tsc_version(void);
const char* tsc_last_error(void);
int tsc_pad_channels(int C);                       /* round up to 16 */
/* 1 when the current device can run the tcgen05 engine (compute capability 10.x) */
int tsc_device_supports_tcgen05(void);

/* ---- layout conversion at the module boundary (reference tensors are NCL fp32, OS_CNN.py:67) ---- */
int tsc_ncl_to_c8(const float* src_ncl, void* dst_c8, int dst_dtype, int B, int C, int L, tsc_stream_t stream);
int tsc_c8_to_ncl(const float* src_c8, float* dst_ncl, int B, int C, int L, tsc_stream_t stream);

/* ---- masked kernel bank: replaces `weight.data = weight * weight_mask` (OS_CNN.py:68) ----------
 * s_of_tap (HOST, Kmax ints): first out channel whose prime kernel covers tap t; channels >= s(t)
 * are live at t (the bank is nested, OS_CNN.py:9-12,28-42).  s(t) >= Cout marks a dead tap.
 * W is the reference parameter [Cout, Cin, Kmax] fp32.  FWD packs W for Y = conv(X); DGRAD packs
 * the tap-reversed transpose for dX = conv^T(dY).  Only live (channel, tap) pairs are stored.
 * When zero_masked != 0 the masked taps of W itself are zeroed in place, as the reference does. */
size_t tsc_packed_weight_bytes(int direction, int dtype, int Cin, int Cout, int Kmax, const int* s_of_tap);
int tsc_pack_weights(int direction, int dtype, float* W, void* packed, int Cin, int Cout, int Kmax,
                     const int* s_of_tap, int zero_masked, tsc_stream_t stream);
/* Both directions (packed_dgrad may be NULL) and the in-place masking in ONE launch -- what a layer's
 * forward needs every step, since the optimizer changes W between steps. */
int tsc_pack_weights_pair(int dtype, float* W, void* packed_fwd, void* packed_dgrad, int Cin, int Cout, int Kmax,
                          const int* s_of_tap, int zero_masked, tsc_stream_t stream);

/* Every layer of a stack (or of a whole model set) in ONE launch: same outputs as tsc_pack_weights_pair per layer.
 * The batch is a HOST struct passed by value to the kernel (<= TSC_PACK_MAX_LAYERS layers per call). */
#define TSC_PACK_MAX_LAYERS 8
typedef struct tsc_pack_layer {
    float* W;                 /* [Cout, Cin, Kmax] fp32 (masked in place when zero_masked) */
    void* packed_fwd;         /* tsc_packed_weight_bytes(TSC_DIR_FWD, ...) bytes */
    void* packed_dgrad;       /* tsc_packed_weight_bytes(TSC_DIR_DGRAD, ...) bytes, or NULL */
    int Cin, Cout, Kmax, zero_masked;
    short s_of_tap[TSC_MAX_TAPS];
} tsc_pack_layer;
typedef struct tsc_pack_batch {
    int n;
    int pad_;
    tsc_pack_layer layer[TSC_PACK_MAX_LAYERS];
} tsc_pack_batch;
int tsc_pack_weights_multi(int dtype, const tsc_pack_batch* batch, tsc_stream_t stream);

/* ---- multi-kernel-size Conv1d as one implicit GEMM: replaces ConstantPad1d + Conv1d
 * (OS_CNN.py:70-71, 163-164) and their cuDNN/oneDNN dgrad -------------------------------------
 * FWD  : x = X  [B, Cin, L] c8(dtype), y = Y  [B, Cout, L] c8 fp32, bias[Cout] or NULL.
 * DGRAD: x = dY [B, Cout, L] c8(dtype), y = dX [B, Cin, L] c8 fp32, bias ignored.
 * Zero padding (pad_left=(Kmax-1)/2, pad_right=Kmax/2, OS_CNN.py:59) is implicit.
 *
 * plan (tcgen05 engine only, NULL for SIMT): the issue schedule of the bank -- one entry per tensor-core
 * instruction and per weight stage.  It depends on the bank geometry only: build it ONCE on the host with
 * tsc_osconv_plan_build into a HOST buffer of tsc_osconv_plan_bytes bytes, copy it to device memory (16 B
 * aligned) and pass that device pointer to every call.
 *
 * epilogue (tcgen05 engine only, may be NULL): reductions fused into the accumulator read-out, each written as
 * one float2 per (CTA, padded channel), n_cta = B * ceil(L/128), CTA i covering rows [128*(i % ceil(L/128)), +128)
 * of sample i / ceil(L/128):
 *   FWD   stat_partial [n_cta][Cout_p][2] = (mean, M2) of y over the CTA's valid rows  -> BatchNorm statistics
 *         (tsc_bn_apply_fused merges them; replaces a separate pass over y).
 *   DGRAD red_partial  [n_cta][Cin_p][2]  = (S1, S2): the layer below is z = act(BN(mask_y)); the written dX is
 *         already d = dX * [mask_scale*mask_y + mask_shift > 0] (mask_scale NULL = no ReLU), and S1 = sum d,
 *         S2 = sum d * (mask_y - mask_mean) * mask_invstd over the CTA's rows -> tsc_bn_bwd_apply_fused. */
typedef struct tsc_conv_epilogue {
    float* stat_partial;
    const float* mask_y;
    const float* mask_scale;
    const float* mask_shift;
    const float* mask_mean;
    const float* mask_invstd;
    float* red_partial;
    /* FWD, inference (eval-mode BatchNorm folded into the convolution): when affine_out != NULL the kernel writes
     *   z = act(affine_scale[c] * (acc + bias[c]) + affine_shift[c] [+ residual])
     * with (scale, shift) [Cout_p] from tsc_bn_eval_coeffs, or -- affine_scale NULL -- derived in the kernel's prologue from
     * the BatchNorm tensors themselves (bn_gamma, bn_beta, bn_mean = running_mean, bn_var = running_var, [Cout] each, bn_eps):
     * scale = gamma / sqrt(var + eps), shift = beta - mean * scale, i.e. no launch besides the convolution,
     * to affine_out in layout affine_out_kind (TSC_OUT_C8_BF16: the next layer's operand; TSC_OUT_C8_F32; TSC_OUT_NCL_F32:
     * the module boundary; TSC_OUT_POOLED: mean over L, [B, Cout], L <= 128 only) and y_c8 may be NULL (not written).
     * residual: c8 fp32 [B][Cout_p/8][L][8] or NULL (the shortcut branch of Res_OS_layer, OS_CNN.py:176-180). */
    const float* affine_scale;
    const float* affine_shift;
    const float* residual;
    void* affine_out;
    int affine_out_kind;
    int affine_relu;
    const float* bn_gamma;
    const float* bn_beta;
    const float* bn_mean;
    const float* bn_var;
    float bn_eps;
} tsc_conv_epilogue;
size_t tsc_osconv_plan_bytes(int direction, int Cin, int Cout, int Kmax, const int* s_of_tap);
int tsc_osconv_plan_build(int direction, int Cin, int Cout, int Kmax, const int* s_of_tap, void* host_plan);
int tsc_osconv(int engine, int direction, const void* x_c8, int dtype, const void* w_packed, const void* plan,
               const float* bias, float* y_c8, const tsc_conv_epilogue* epilogue, int B, int L, int Cin, int Cout,
               int Kmax, const int* s_of_tap, tsc_stream_t stream);

/* ---- weight gradient on live taps only (masked taps get exact zeros) ---------------
 * dW[co,ci,t] = sum_{b,l} dY[b,co,l] * X[b,ci,l+t-pad_left];  dy/x c8(dtype); dW [Cout,Cin,Kmax] fp32.
 * Deterministic: partial sums per position split, then an ordered reduction (no float atomics).
 * accumulate != 0: dW += result (dW is then typically a slice of the flat gradient bucket that the data-parallel
 * all-reduce and the optimizer read -- the kernel writes the collective's operand in place). */
size_t tsc_oswgrad_workspace_bytes(int engine, int B, int L, int Cin, int Cout, int Kmax);
int tsc_oswgrad(int engine, const void* dy_c8, const void* x_c8, int dtype, float* dW, void* workspace, int accumulate,
                int B, int L, int Cin, int Cout, int Kmax, const int* s_of_tap, tsc_stream_t stream);

/* ---- BatchNorm1d (+ReLU, + shortcut add): replaces OS_CNN.py:72-74, 165, 176-180 --------------
 * tsc_bn_stats: per-channel Welford over (B, L) of y [B,C,L] c8 fp32 -> mean, invstd=1/sqrt(var_b+eps),
 *   scale = gamma*invstd, shift = beta - mean*scale (all [Cp]); running stats (nullable) updated with
 *   momentum and the UNBIASED variance, as torch does.
 * tsc_bn_eval_coeffs: the same coefficients from the running statistics (eval mode with grad,
 *   train_and_test.py:583-586). */
size_t tsc_bn_workspace_bytes(int B, int C, int L);
int tsc_bn_stats(const float* y_c8, const float* gamma, const float* beta, float* workspace,
                 float* mean, float* invstd, float* scale, float* shift,
                 float* running_mean, float* running_var, float momentum, float eps,
                 int B, int C, int L, tsc_stream_t stream);
int tsc_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float eps, float* mean, float* invstd, float* scale, float* shift, int C, tsc_stream_t stream);
/* out = act(scale*y + shift [+ scale2*y2 + shift2]), act = relu when relu != 0. */
int tsc_bn_apply(const float* y_c8, const float* scale, const float* shift,
                 const float* y2_c8, const float* scale2, const float* shift2,
                 int relu, void* out, int out_kind, int B, int C, int L, tsc_stream_t stream);
/* backward, pass 1: d = dz * [act'(.)] ; S1 = sum d ; S2 = sum d * (y-mean)*invstd  (per channel).
 * The activation mask is recomputed from (ym, scale, shift [, ym2, scale2, shift2]); pass ym = NULL
 * for "no ReLU". */
int tsc_bn_bwd_reduce(const float* dz_c8, const float* y_c8, const float* mean, const float* invstd,
                      const float* ym_c8, const float* scale, const float* shift,
                      const float* ym2_c8, const float* scale2, const float* shift2,
                      float* workspace, float* s1, float* s2, int B, int C, int L, tsc_stream_t stream);
/* backward, pass 2: dy = gamma*invstd * (d - S1/N - yhat*S2/N) (training) or gamma*invstd*d (eval),
 * written as c8 (dy_dtype) -- the operand of dgrad and wgrad. */
int tsc_bn_bwd_apply(const float* dz_c8, const float* y_c8, const float* mean, const float* invstd,
                     const float* gamma, const float* s1, const float* s2, int training,
                     const float* ym_c8, const float* scale, const float* shift,
                     const float* ym2_c8, const float* scale2, const float* shift2,
                     void* dy_c8, int dy_dtype, int B, int C, int L, tsc_stream_t stream);

/* ---- fused BatchNorm path of the tcgen05 engine ------------------------------------------------------------
 * The statistics of a training-mode layer arrive as per-CTA partials from the conv epilogue (tsc_osconv) and are
 * merged in the prologue of the apply kernels: forward = conv + tsc_bn_apply_fused (no statistics pass), backward
 * = tsc_bn_bwd_apply_fused + wgrad + dgrad (whose epilogue masks and reduces for the layer below).
 *
 * tsc_bn_branch: one conv+BN branch.  stat_partial != NULL = training mode: [n_part][Cp][2] (mean, M2) per 128-row
 * CTA -> batch statistics (biased variance for normalisation; running statistics, when given and momentum > 0,
 * updated with the unbiased one).  stat_partial == NULL = eval mode with autograd (train_and_test.py:583-586): the
 * running statistics are used.  coef [4][Cp] (mean, invstd, scale, shift) is written for the backward pass. */
typedef struct tsc_bn_branch {
    const float* y_c8;
    const float* stat_partial;
    const float* gamma;
    const float* beta;
    float* running_mean;
    float* running_var;
    float momentum;
    float eps;
    float* coef;
} tsc_bn_branch;
/* out = act(BN_a(y_a) [+ BN_b(y_b)]) : OS_CNN.py:72-74 and, with b, the shortcut add of :176-180.
 * n_part must be B*ceil(L/128) (the conv kernel's CTA count). */
int tsc_bn_apply_fused(const tsc_bn_branch* a, const tsc_bn_branch* b, int n_part, int relu, void* out, int out_kind,
                       int B, int C, int L, tsc_stream_t stream);
/* number of row splits the fused BN kernels use per 8-channel chunk (rows of tsc_bn_bwd_top's red_partial) */
int tsc_bn_fused_splits(int B, int C, int L);

typedef struct tsc_bn_bwd_branch {
    const float* y_c8;        /* pre-BN conv output */
    const float* coef;        /* [4][Cp] from the forward */
    const float* gamma;
    int training;
    float* red_partial;       /* [n_part][Cp][2] (S1, S2) partial sums */
    float* dgamma;            /* outputs of tsc_bn_bwd_apply_fused (nullable) */
    float* dbeta;
    float* dbias;
} tsc_bn_bwd_branch;
/* Top of a stack (gradient arrives NCL fp32): d = dout*[act'] as c8 fp32 and red_partial[tsc_bn_fused_splits][Cp][2]
 * of branch a (and b: same d, its own yhat). */
int tsc_bn_bwd_top(const float* dout_ncl, const tsc_bn_bwd_branch* a, const tsc_bn_bwd_branch* b, int relu, float* d_c8,
                   int B, int C, int L, tsc_stream_t stream);
/* The same when the stack's output was pooled (TSC_OUT_POOLED): dpooled [B, C], dout[b,c,l] = dpooled[b,c] / L. */
int tsc_bn_bwd_top_pooled(const float* dpooled, const tsc_bn_bwd_branch* a, int relu, float* d_c8, int B, int C, int L,
                          tsc_stream_t stream);
/* dy = gamma*invstd*(d - S1/N - yhat*S2/N) (training) or gamma*invstd*d (eval) as c8(dy_dtype); d is already masked;
 * (S1, S2) = sum of the n_part rows of red_partial.  Also dgamma = S2, dbeta = S1, dbias (+= when accumulate). */
int tsc_bn_bwd_apply_fused(const float* d_c8, const tsc_bn_bwd_branch* a, int n_part, int accumulate, void* dy_c8,
                           int dy_dtype, int B, int C, int L, tsc_stream_t stream);

/* ---- feature-level style transfer (inserted at train_and_test.py:552-561; no reference operator,
 * spec = SURVEY 8c).  Rows are the (b, c) rows of an NCL fp32 tensor: R = B*C rows of L floats. ---- */
int tsc_rowstats_welford(const float* x, float* mean, float* var_unbiased, int R, int L, tsc_stream_t stream);
/* stats[R][4] = (mu_c, sigma_c, mu_s, sigma_s), sigma = sqrt(var_unbiased + eps) */
int tsc_adain_fwd(const float* content, const float* style, float* out, float* stats, float eps,
                  int R, int L, tsc_stream_t stream);
int tsc_adain_bwd(const float* dy, const float* content, const float* style, const float* stats,
                  float* dcontent, float* dstyle, int R, int L, tsc_stream_t stream);
/* D[b] = (a_b a_b^T - s_b s_b^T)/(C L)  ([B,C,C] fp32, kept for backward); loss = mean(D^2). */
size_t tsc_gram_workspace_bytes(int B, int C, int L);
int tsc_gram_loss_fwd(int engine, const float* a, const float* s, float* D, float* loss, float* workspace,
                      int B, int C, int L, tsc_stream_t stream);
/* da = g * 4/(B C^3 L) * D a ; ds = -g * 4/(B C^3 L) * D s ; g = *dloss (device scalar). */
int tsc_gram_loss_bwd(int engine, const float* D, const float* a, const float* s, const float* dloss,
                      float* da, float* ds, int B, int C, int L, tsc_stream_t stream);

/* ---- C-DAN consumer of the transferred features (C_DAN.py:49-82; BASELINE configuration 3).  Rows are stacked:
 * m in [0, B) = target half, [B, 2B) = generated (source-to-target) half.  The big random projection
 * y0 = flatten(feature) @ R0 (C_DAN.py:20, a plain dense GEMM) is the caller's (cuBLAS); everything else between the
 * classifier logits / y0 and the critic's input, and everything behind the critic's output, is here.
 *   prob = softmax(logits)                       C_DAN.py:53-54      [2B, K], K <= 32
 *   fusion = (y0 / scale_div) * (prob @ r1)      C_DAN.py:20-25      [2B, D]   (scale_div = D^(1/2) for two views)
 *   u = 1 + exp(-H(prob)), H = -sum p log(p+1e-5) C_DAN.py:32-37,70-71 [2B] */
int tsc_cdan_fuse_fwd(const float* y0, const float* logits, const float* r1, float* fusion, float* prob, float* u,
                      int M, int K, int D, float scale_div, tsc_stream_t stream);
/* Backward of the above including the reference's three gradient reversals (grl_hook, C_DAN.py:38-41):
 * dfusion is the gradient at the critic's input (widgets.py:121-122); coeff (device, 3 floats) = reversal strength of
 * the critic input for the target rows, for the generated rows, and of the entropy (C_DAN.py:68-69).  du may be NULL. */
int tsc_cdan_fuse_bwd(const float* dfusion, const float* y0, const float* prob, const float* r1, const float* u,
                      const float* du, const float* coeff, float* dy0, float* dlogits, int B, int K, int D,
                      float scale_div, tsc_stream_t stream);
/* w = u / sum(u) per half (the sum is a constant, C_DAN.py:72-75); loss = sum_t(w) * sum_t(critic) - the same over the
 * generated half (C_DAN.py:78-81: [B] * [B,1] broadcasts to [B,B] before the sum).  saved: 6 floats for backward. */
int tsc_cdan_distance_fwd(const float* u, const float* critic_out, float* loss, float* saved, int B,
                          tsc_stream_t stream);
int tsc_cdan_distance_bwd(const float* dloss, const float* saved, float* du, float* dcritic_out, int B,
                          tsc_stream_t stream);

/* ---- fused multi-tensor RMSprop over flat fp32 buffers: replaces the 5 torch.optim.RMSprop instances of the
 * path (train_and_test.py:97-101): v = alpha v + (1-alpha) g^2 ; p -= lr g / (sqrt(v)+eps), g = grad_scale*grad.
 * group_end / group_lr (HOST, ngroups entries): exclusive end offset and learning rate of each parameter group. */
int tsc_rmsprop_step(float* params, const float* grads, float* square_avg, long long n,
                     const long long* group_end, const float* group_lr, int ngroups, float alpha, float eps,
                     float grad_scale, tsc_stream_t stream);
/* The same with WGAN weight clipping of chosen groups after their update (the critic's p.data.clamp_(-c, c),
 * train_and_test.py:763-766): group_clamp (HOST, ngroups entries, or NULL) = c, <= 0 for "not clipped". */
int tsc_rmsprop_step_clamped(float* params, const float* grads, float* square_avg, long long n,
                             const long long* group_end, const float* group_lr, const float* group_clamp, int ngroups,
                             float alpha, float eps, float grad_scale, tsc_stream_t stream);

/* ---- callers either side of the path --------------------------------------------------------
 * GradNorm (train_and_test.py:683-690): per loss, sum over the shared block's parameter tensors of
 * torch.norm(w_i * g_p) = |w_i| * ||g_p||.  tsc_multi_l2norm writes norms[i] = ||tensor i||_2 for i < count and
 * norms[count] = their sum (fixed summation order, no float atomics); workspace from the query function. */
typedef struct tsc_tensor_list {
    const float* p[TSC_MAX_LIST];  /* device pointers, fp32, 4-byte aligned (16-byte aligned tensors take the vector path) */
    long long n[TSC_MAX_LIST];     /* elements */
    int count;
    int pad_;
} tsc_tensor_list;
size_t tsc_multi_l2norm_workspace_bytes(int count);
int tsc_multi_l2norm(const tsc_tensor_list* list, float* norms, float* workspace, tsc_stream_t stream);
/* Multi-source voting (multi_source_voting.py:281-311 and the accuracy of utils.py:27-183).
 * logits [N,K] fp32, labels [N] int64 (or NULL) -> pred [N] int32 (first maximum, numpy.argmax; nullable),
 * counts [2K] int32 = (#predicted as k, #predicted as k and labelled k), precision [K] fp64 = correct / predicted,
 * 0 for a class never predicted (nullable). */
int tsc_class_precision(const float* logits, const long long* labels, int* pred, int* counts, double* precision,
                        int N, int K, tsc_stream_t stream);
/* multi_source_voting.py:357-407: logits [M,N,K] of M models, precision [M,K] from tsc_class_precision on the training
 * split.  w_m = precision_m / mean_m(precision) (NaN -> 0); per model p = softmax(logits), H = entropy(p),
 * score += p * (1 + entropy_gain * exp(-H)) * weight_base ^ w_m  (the reference: gain 120, base 9); pred = argmax.
 * score [N,K] fp32, pred [N] int32. */
int tsc_entropy_vote(const float* logits, const double* precision, float* score, int* pred, int M, int N, int K,
                     float entropy_gain, float weight_base, tsc_stream_t stream);

/* ---- classifier head fused with the training loss: replaces `self.hidden(X_f)` (OS_CNN/OS_CNN.py:108-109) followed by
 * nn.CrossEntropyLoss (train_and_test.py:593-603) and their autograd -- one launch forward, one backward.
 *   logits[b,k] = bias[k] + sum_c pooled[b,c] W[k,c];  prob = softmax(logits);  loss = -(1/B) sum_b log prob[b, labels[b]]
 *   backward: g = dlogits (nullable: gradient from other consumers of the logits) + dloss * (prob - onehot) / B
 *             (dloss: device scalar, NULL = 1; labels NULL = no loss term);  dpooled = g W (nullable);  dW = g^T pooled;
 *             dbias = sum_b g;  accumulate != 0 adds into dW / dbias (slices of the flat gradient bucket).
 * pooled [B,C], W [K,C], bias [K], logits / prob [B,K] fp32; labels int64 (NULL: logits only); K <= TSC_MAX_CLASSES.
 * workspace: tsc_head_ce_workspace_bytes(B) bytes, ZERO before its first use (every launch leaves it zeroed). */
size_t tsc_head_ce_workspace_bytes(int B);
int tsc_head_ce_fwd(const float* pooled, const float* W, const float* bias, const long long* labels, float* logits,
                    float* prob, float* loss, void* workspace, int B, int C, int K, tsc_stream_t stream);
int tsc_head_ce_bwd(const float* dloss, const float* dlogits, const float* prob, const long long* labels,
                    const float* pooled, const float* W, float* dpooled, float* dW, float* dbias, int accumulate,
                    int B, int C, int K, tsc_stream_t stream);
/* out = sum_i weights[i] * *terms[i] over n <= 8 device scalars (weights: HOST array): the step's total loss
 * (train_and_test.py:660-672 style weighted sums) in one launch. */
int tsc_weighted_scalar_sum(const float* const* terms, const float* weights, int n, float* out, tsc_stream_t stream);

/* ---- debugging aid: the tcgen05 kernels bound every mbarrier wait; a timed-out wait stores a
 * non-zero code here (device word, read back by the caller when it wants to). */
int tsc_debug_read_and_clear_watchdog(int* host_code);
/* profiling aid: when dev_buf != NULL (device memory, >= 8 x int64), CTA 0 of every following tcgen05 conv launch
 * stores clock64() at its phase boundaries there (entry, setup, tile loaded, first weights, MMAs issued,
 * accumulator ready, epilogue done, exit).  NULL switches it off (the default). */
int tsc_debug_set_timeline(void* dev_buf);

#ifdef __cplusplus
}
#endif
#endif /* TSC_B200_H */
