/* cultionet_b200 -- C ABI of the B200-native TowerUNet hot path.
 *
 * One shared object (cultionet_b200/libcultionet_b200.so), extern "C", plain pointers and sizes.
 * The reference (jgrss/cultionet) has no FFI of its own: it is pure Python over torch/cuDNN/natten.
 * Each entry point below therefore names the reference *Python call site* whose device work it
 * replaces (paths relative to the reference repo root, `src/cultionet/...`).
 *
 * Conventions
 *   - activations are pixel-major ("NHWC"): [P = B*H*W][C], channel contiguous; `dtype` selects the
 *     storage type of activations and packed weights (CNB_F32 parity mode, CNB_BF16 throughput mode);
 *     statistics, parameter gradients, loss terms are always fp32 (loss partial sums fp64);
 *   - every pointer is a device pointer unless stated; outputs and workspaces are caller-allocated;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no entry point synchronises,
 *     allocates or frees, so calls are legal under CUDA-graph capture and from autograd worker threads;
 *   - return value 0 = OK, otherwise a CNB_ERR_* code with a message in cnb_last_error() (thread-local).
 */
#ifndef CULTIONET_B200_H
#define CULTIONET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNB_VERSION 100 /* 0.1.0 */

enum { CNB_OK = 0, CNB_ERR_INVALID = 1, CNB_ERR_CUDA = 2, CNB_ERR_UNSUPPORTED = 3 };
enum { CNB_F32 = 0, CNB_BF16 = 1 };
enum { CNB_MAX_SRC = 6 };

int cnb_version(void);
/* compiled-for architecture (100 for sm_100a); 0 for the CPU test interpreter build */
int cnb_sm_arch(void);
const char* cnb_last_error(void);
/* kernels launched by this library since load (bench.py reports the per-step delta as `gpu_launches`) */
int64_t cnb_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution over a *virtual concatenation* of up to CNB_MAX_SRC pixel-major sources.
 *   y[b,oy,ox,n] = bias[n] + sum_{tap=(ky,kx)} sum_{c} X[b, iy, ix, c] * Wp[tap][n][c]
 *   direct     (transposed=0): iy = oy*stride - pad + ky*dil              (nn.Conv2d, nn.Linear, dgrad of ConvTranspose2d)
 *   transposed (transposed=1): iy = (oy + pad - ky*dil)/stride if exact   (nn.ConvTranspose2d, dgrad of strided Conv2d)
 * X is the channel-wise concatenation of the sources (torch.cat(dim=1) in the reference is never materialised).
 * Replaces: nn.Conv2d in ConvBlock2d (nn/modules/convolution.py:88-116), the 1x1 skips (:311-318), StreamConv2d
 * (nn/modules/unet_parts.py:205-221), nn.ConvTranspose2d (convolution.py:56-62), the qkv/proj nn.Linear of
 * natten.NeighborhoodAttention2D (convolution.py:341-350), torch.cat in TowerUNetBlock.forward (unet_parts.py:733-758)
 * and the autograd dgrad/wgrad of all of these.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    const void* src[CNB_MAX_SRC];   /* source s: element (pixel p, channel c) at src[s] + p*src_stride[s] + c */
    int32_t src_c[CNB_MAX_SRC];     /* channels taken from source s */
    int32_t src_stride[CNB_MAX_SRC];/* elements between consecutive pixels of source s */
    int32_t nsrc;
    int32_t B, Hin, Win, Hout, Wout;
    int32_t KH, KW, stride, pad, dil, transposed;
    const void* w_packed;           /* [KH*KW][N][Ctot] in `dtype`; Ctot = sum(src_c) */
    int64_t w_tap_stride;           /* elements between taps */
    int32_t w_row_stride;           /* elements between rows n */
    int32_t N;                      /* output channels */
    const float* bias;              /* [N] fp32 or NULL */
    void* out;                      /* element (p, n) at out + p*out_stride + n */
    int32_t out_stride;
    float* stats;                   /* optional [2*N] fp32, zeroed by the caller: the tcgen05 kernel adds sum(y) and sum(y^2) per output
                                     * channel over all pixels (y = the stored, rounded value), i.e. BatchNorm's batch statistics come out
                                     * of the convolution epilogue and cnb_bn_stats is skipped.  Only honoured by cnb_conv2d_fwd_tc
                                     * (N <= 1024); the other kernels require NULL. */
    /* Split output (the data gradient of a convolution over a virtual concatenation: one GEMM with N = Ctot whose columns land in
     * the gradient tensors of the individual sources).  nout = 0: the single output above.  nout > 0: `out` is ignored and output
     * channels [sum_{j<i} out_seg_c[j], + out_seg_c[i]) go to out_seg[i] + p*out_seg_stride[i] + c.  Only the tcgen05 kernel
     * implements it (every out_seg_c a multiple of 32, sum = N, no bias); ask cnb_conv2d_tc_eligible() first. */
    int32_t nout;
    void* out_seg[CNB_MAX_SRC];
    int32_t out_seg_c[CNB_MAX_SRC];
    int32_t out_seg_stride[CNB_MAX_SRC];
    /* Fused epilogue for inference (ConvBlock2d in eval mode, convolution.py:88-116: Conv2d -> BatchNorm2d(running statistics) -> [SiLU]):
     *   y[n] = act(acc[n] * ep_scale[n] + ep_shift[n]),  ep_scale = gamma * rsqrt(var + eps), ep_shift = beta - mean * ep_scale
     * (fp32 [N] each, as cnb_bn_finalize produces them), ep_act: a CNB_ACT_* code.  ep_scale == NULL: off.  The normalised
     * activation is written once instead of conv out -> read -> write.  Only the tcgen05 kernel implements it (bias must be NULL,
     * single output); the other kernels return CNB_ERR_UNSUPPORTED. */
    const float* ep_scale;
    const float* ep_shift;
    int32_t ep_act;
} cnb_conv_desc;

/* picks the tiny-channel kernel, else the tcgen05/TMA kernel when cnb_conv2d_tc_eligible(), else the CUDA-core kernel */
int cnb_conv2d_fwd(const cnb_conv_desc* d, int dtype, void* stream);
/* CUDA-core implicit GEMM (any shape, fp32 or bf16 storage, fp32 accumulate): the parity-mode path */
int cnb_conv2d_fwd_generic(const cnb_conv_desc* d, int dtype, void* stream);
/* tcgen05.mma + TMEM + TMA implicit GEMM: bf16, every source with a 16-byte pixel pitch; any channel counts (partial tiles are
 * zero-filled by TMA and masked); stride 1..4 direct or transposed (stride > 1 needs a single source and <= 9 taps).
 * CNB_ERR_UNSUPPORTED otherwise */
int cnb_conv2d_fwd_tc(const cnb_conv_desc* d, int dtype, void* stream);
/* 0: not eligible; 1: eligible; 2: eligible and able to fill `stats` */
int cnb_conv2d_tc_eligible(const cnb_conv_desc* d, int dtype);
/* one thread per output pixel, filter bank in shared memory: N <= 16 output and <= 16 input channels (the Psi-Net heads' 3->1 and
 * 3->3 convolutions, nn/modules/unet_parts.py:215-220, :262-270); cnb_conv2d_fwd picks it first when it applies */
int cnb_conv2d_fwd_tiny(const cnb_conv_desc* d, int dtype, void* stream);

/* Weight gradient of the same convolution for ONE source slice:
 *   dWp[tap][n][k_off + c] += sum_p X_s[gather(p, tap)][c] * dY[p][n]         (fp32, atomically accumulated)
 * `dwp` has the packed geometry [taps][N][Ctot] fp32 and must be zeroed by the caller before the first slice. */
typedef struct {
    const void* src; int32_t src_c, src_stride;
    int32_t k_off, Ctot;
    int32_t B, Hin, Win, Hout, Wout;
    int32_t KH, KW, stride, pad, dil, transposed;
    const void* dy; int32_t dy_stride; int32_t N;
    float* dwp;
} cnb_wgrad_desc;

/* picks the tcgen05 split-K kernel when cnb_conv2d_wgrad_tc_eligible(), else the CUDA-core kernel */
int cnb_conv2d_wgrad(const cnb_wgrad_desc* d, int dtype, void* stream);
int cnb_conv2d_wgrad_generic(const cnb_wgrad_desc* d, int dtype, void* stream);
int cnb_conv2d_wgrad_tc(const cnb_wgrad_desc* d, int dtype, void* stream);
int cnb_conv2d_wgrad_tc_eligible(const cnb_wgrad_desc* d, int dtype);
/* N <= 8 and a source slice of <= 16 channels: per-thread partial sums, one atomic per CTA */
int cnb_conv2d_wgrad_tiny(const cnb_wgrad_desc* d, int dtype, void* stream);
/* dst[p][c] = src[p][c] for c < C, 0 for C <= c < dst_stride: gives a skinny tensor (e.g. the 3-channel gradient of a Psi-Net
 * stream) the 16-byte pixel pitch the TMA-fed kernels need */
int cnb_repitch(const void* src, int src_stride, void* dst, int dst_stride, int64_t P, int C, int dtype, void* stream);
/* Multi-tensor copy / accumulate of small fp32 segments in one launch per 160 segments.  table: HOST array of n_entries records
 * {const float* src; float* dst; int32 n; int32 mode (0: dst = src, 1: dst += src)} (24 bytes each, device pointers inside); it is
 * copied into the kernel arguments before the call returns (no device table, nothing to upload; a captured graph keeps its copy).
 * max_len = the largest n.
 * Used for parameter plumbing that the reference leaves to torch.cat / autograd accumulation: stacking the three Psi-Net stream
 * filters (nn/modules/unet_parts.py:281-309), scattering BatchNorm / LayerNorm / scalar gradients into the flat gradient buffer. */
int cnb_multi_copy(const void* table, int n_entries, int max_len, void* stream);
/* out[b][p][c] = e[b][c] for p < HW: the per-sample GeoEmbeddings vector broadcast over a level's pixels before it joins the
 * full-scale-skip concatenation (nn/modules/unet_parts.py:739-750, geo_encoding.py:5-26) */
int cnb_broadcast_pixels(const void* e, void* out, int B, int64_t HW, int C, int dtype, void* stream);

/* fp32 parameter (any strided [n][k][tap] view) -> packed [tap][N][wp_pitch] in `dtype` (wp_pitch >= K; padding columns zeroed):
 *   wp[tap][n][k] = w[n*s_n + k*s_k + tap*s_tap]
 * nn.Conv2d weight [Cout,Cin,kh,kw]: forward (n=Cout,k=Cin) s_n=Cin*taps,s_k=taps; dgrad (n=Cin,k=Cout) s_n=taps,s_k=Cin*taps.
 * nn.ConvTranspose2d weight [Cin,Cout,kh,kw]: forward (n=Cout,k=Cin) s_n=taps,s_k=Cout*taps; dgrad s_n=Cout*taps,s_k=taps. */
int cnb_pack_weight(const float* w, void* wp, int dtype, int taps, int N, int K, int wp_pitch,
                    int64_t s_n, int64_t s_k, int64_t s_tap, void* stream);
/* Both packed layouts of one parameter from ONE read of it (tiled through shared memory, every access coalesced):
 *   wp[tap][n][k] (pitch_k >= K) as cnb_pack_weight, and, when wd != NULL, the data-gradient layout wd[tap][k][n] (pitch_n >= N);
 * s_n/s_k/s_tap are the parameter's element strides for the FORWARD roles of n (output channel) and k (input channel). */
int cnb_pack_weight2(const float* w, void* wp, void* wd, int dtype, int taps, int N, int K, int pitch_k, int pitch_n,
                     int64_t s_n, int64_t s_k, int64_t s_tap, void* stream);
/* cnb_pack_weight2 for MANY parameters in one launch.  `table` is a DEVICE array of `ndesc` descriptors sorted by tile0; descriptor
 * i owns tiles [tile0, tile0 + tiles_x * tiles_y) of the launch, tiles_x = ceil(max(K, pitch_k) / 32),
 * tiles_y = ceil(max(N, wd ? pitch_n : N) / 32); total_tiles = sum over descriptors; max_taps = largest taps in the table (<= 32). */
typedef struct cnb_pack_desc {
    const float* w;
    void* wp;
    void* wd; /* may be NULL */
    int32_t taps, N, K, pitch_k, pitch_n;
    int32_t tile0, tiles_x;
    int32_t reserved;
    int64_t s_n, s_k, s_tap;
} cnb_pack_desc;
int cnb_pack_weights_batched(const cnb_pack_desc* table, int ndesc, int total_tiles, int max_taps, int dtype, void* stream);
/* inverse scatter of a packed fp32 gradient into the parameter layout: g[...] (+)= dwp[tap][n][k].
 * mode bit 0: accumulate into g (else overwrite); bit 1: clear dwp while reading it (persistent split-K accumulator). */
int cnb_unpack_wgrad(float* dwp, float* g, int taps, int N, int K,
                     int64_t s_n, int64_t s_k, int64_t s_tap, int mode, void* stream);
/* cnb_unpack_wgrad for MANY parameters in one launch (taps <= 32 each).  Same device table layout as cnb_pack_weights_batched with the
 * roles: w = gradient destination g (written), wp = fp32 accumulator dwp[taps][N][K], reserved = mode, tiles_x = ceil(K / 32),
 * tiles_y = ceil(N / 32), descriptors sorted by tile0. */
int cnb_unpack_wgrads_batched(const cnb_pack_desc* table, int ndesc, int total_tiles, int max_taps, void* stream);
/* db[n] (+)= sum_p dy[p][n] (bias gradient) */
int cnb_bias_grad(const void* dy, int dy_stride, int64_t P, int N, float* db, int accumulate, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm (+SiLU) over a [P][L] matrix whose column j belongs to channel (j / ch_div) % C.
 * (L == C, ch_div == 1 for BatchNorm2d on pixel-major data; PreTimeReduction's BatchNorm3d uses ch_div = T'.)
 * Replaces nn.BatchNorm2d/3d + SetActivation("SiLU") in ConvBlock2d (convolution.py:112-116) and Conv3d (models/nunet.py:40-54).
 * ------------------------------------------------------------------------------------------------ */
/* sums[0..C) = sum x, sums[C..2C) = sum x^2 (fp32; zeroed inside) */
int cnb_bn_stats(const void* x, int64_t P, int L, int C, int ch_div, float* sums, int dtype, void* stream);
/* training: mean/var from sums -> save_mean, save_rstd, scale = gamma*rstd, shift = beta - mean*scale; updates running stats
 * (momentum, unbiased variance) when running_mean != NULL.  eval (sums == NULL): uses the running stats. */
int cnb_bn_finalize(const float* sums, int64_t count, int C, const float* gamma, const float* beta, float eps, float momentum,
                    float* running_mean, float* running_var, float* save_mean, float* save_rstd, float* scale, float* shift,
                    void* stream);
/* Activation codes of every `act` / `ep_act` argument (reference nn/modules/activations.py:5-24 builds getattr(torch.nn, name)();
 * SiLU is the reference's default): */
#define CNB_ACT_NONE 0
#define CNB_ACT_SILU 1
#define CNB_ACT_RELU 2
#define CNB_ACT_LEAKY_RELU 3 /* negative_slope 0.01 */
#define CNB_ACT_GELU 4       /* erf form (approximate='none') */
#define CNB_ACT_MISH 5
#define CNB_ACT_ELU 6        /* alpha 1 */
#define CNB_ACT_TANH 7
#define CNB_ACT_SIGMOID 8
#define CNB_ACT_HARDSWISH 9
/* y = act(x*scale[ch] + shift[ch]) (+ residual) ; act: a CNB_ACT_* code */
int cnb_bn_act_fwd(const void* x, const float* scale, const float* shift, const void* residual, void* y,
                   int64_t P, int L, int C, int ch_div, int act, int dtype, void* stream);
/* training-mode cnb_bn_finalize + cnb_bn_act_fwd as one call (one launch on the bulk-copy path: the apply kernel derives scale/shift
 * from the batch sums itself and its first CTA publishes mean/rstd/scale/shift and updates the running statistics) */
int cnb_bn_train_fwd(const void* x, const float* sums, int64_t count, const float* gamma, const float* beta, float eps, float momentum,
                     float* running_mean, float* running_var, float* save_mean, float* save_rstd, float* scale, float* shift,
                     const void* residual, void* y, int64_t P, int L, int C, int ch_div, int act, int dtype, void* stream);
/* backward pass 1: dsums[0..C) = sum dz, dsums[C..2C) = sum dz*xhat, dz = dy * act'(z) */
int cnb_bn_act_bwd_reduce(const void* x, const void* dy, const float* save_mean, const float* save_rstd, const float* gamma,
                          const float* beta, int64_t P, int L, int C, int ch_div, int act, float* dsums, int dtype, void* stream);
/* the same without the clearing memset: dsums must be ZERO on entry (a slice of a pre-cleared arena) -- one memset node less per
 * BatchNorm in a captured step, and the kernel keeps its programmatic dependency on the kernel before it */
int cnb_bn_act_bwd_reduce_acc(const void* x, const void* dy, const float* save_mean, const float* save_rstd, const float* gamma,
                          const float* beta, int64_t P, int L, int C, int ch_div, int act, float* dsums, int dtype, void* stream);
/* backward pass 2: dx = gamma*rstd*(dz - dsum/count - xhat*dsum_xhat/count); also writes dgamma = dsums[C..2C), dbeta = dsums[0..C)
 * when train_stats != 0; with train_stats == 0 (eval BN) dx = dz*gamma*rstd. */
int cnb_bn_act_bwd_apply(const void* x, const void* dy, const float* save_mean, const float* save_rstd, const float* gamma,
                         const float* beta, const float* dsums, int64_t count, void* dx, int64_t P, int L, int C, int ch_div,
                         int act, int train_stats, int dtype, void* stream);

/* out = a + b (+ c) (+ d), elementwise over n elements (ResidualAConv.forward sums, convolution.py:377-395) */
int cnb_add_n(const void* a, const void* b, const void* c, const void* d, void* out, int64_t n, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over the channel axis of [P][C] (nn.LayerNorm in models/nunet.py:86-90 and convolution.py:340,351)
 * ------------------------------------------------------------------------------------------------ */
int cnb_layernorm_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y, float* save_mean, float* save_rstd,
                      int64_t P, int C, int dtype, void* stream);
/* dgamma/dbeta are accumulated atomically: zero them first */
int cnb_layernorm_bwd(const void* x, const void* dy, const float* gamma, const float* save_mean, const float* save_rstd, void* dx,
                      float* dgamma, float* dbeta, int64_t P, int C, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 2-D neighbourhood attention core (natten==0.17.1 NeighborhoodAttention2D between its qkv and proj Linears,
 * as configured at convolution.py:341-350; window rule in oracle/natten_ref.py).
 * qkv: [B,H,W,3*heads*hd] with channel order (3, heads, hd); out: [B,H,W,heads*hd].
 * ------------------------------------------------------------------------------------------------ */
/* Two implementations: a tiled one (8x16 pixel tile of one head per CTA, k/v + halo staged in shared memory, online softmax,
 * gather-form backward without atomics) for head_dim a multiple of the 16-byte vector width, and a warp-per-(pixel, head) kernel
 * for everything else.  cnb_na2d_tiled_eligible() tells the caller which workspaces the backward will need. */
int cnb_na2d_tiled_eligible(int B, int H, int W, int heads, int hd, int ksize, int dilation, int dtype);
/* lse: fp32 [B,H,W,heads] log-sum-exp of the scaled logits (written by the tiled kernel; may be NULL -> warp kernel) */
int cnb_na2d_fwd(const void* qkv, void* out, float* lse, int B, int H, int W, int heads, int hd, int ksize, int dilation, float scale,
                 int dtype, void* stream);
/* tiled path: needs out (the forward result), lse and dvec (fp32 [B,H,W,heads] scratch); dqkv_acc may be NULL.
 * warp path:  needs dqkv_acc, fp32 [B,H,W,3*heads*hd] scratch (zeroed inside); out/lse/dvec may be NULL.
 * dqkv: [B,H,W,3*heads*hd] in `dtype`. */
int cnb_na2d_bwd(const void* qkv, const void* dout, const void* out, const float* lse, float* dvec, float* dqkv_acc, float* pds_ws,
                 void* dqkv, int B, int H, int W, int heads, int hd, int ksize, int dilation, float scale, int dtype, void* stream);
/* Specialised backward (bf16, kernel 3 or 7, dilation 1 or 2, head_dim 32 or 64): the query-side pass stores the attention
 * probabilities and their scaled logit gradients so that the key-side pass is a pure gather.  Returns the number of fp32 elements
 * of that scratch (pds_ws: [B,H,W,heads,k*k] bf16 pairs (p, scale*ds), 4 bytes each) or 0 when the shape takes the generic kernels
 * (pds_ws may then be NULL). */
int64_t cnb_na2d_bwd_workspace_floats(int B, int H, int W, int heads, int hd, int ksize, int dilation, int dtype);

/* ------------------------------------------------------------------------------------------------
 * Bilinear resize, align_corners=True, pixel-major (check_upsample, nn/functional.py:72-81)
 * ------------------------------------------------------------------------------------------------ */
int cnb_resize_bilinear_fwd(const void* x, void* y, int B, int Hin, int Win, int Hout, int Wout, int C, int dtype, void* stream);
int cnb_resize_bilinear_bwd(const void* dy, void* dx, int B, int Hin, int Win, int Hout, int Wout, int C, int dtype, void* stream);
/* the same, also producing colsum[c] (+)= sum over pixels of dx[.., c] in the same pass: the bias gradient of the ConvTranspose2d whose
 * output was resized (convolution.py:45-68 followed by nn/functional.py:72-81), so dx is not read again by cnb_bias_grad.  colsum may be
 * NULL (plain backward); accumulate = 0 overwrites colsum. */
int cnb_resize_bilinear_bwd_colsum(const void* dy, void* dx, int B, int Hin, int Win, int Hout, int Wout, int C, float* colsum, int accumulate,
                                   int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PreTimeReduction first stage (models/nunet.py:31-39): Conv3d(C->C,(k,1,1), bias=False) over x[B,C,T,H,W] fp32.
 *   u[p][c2*T' + t'] = sum_{c,dt} w1[c2][c][dt] * x[b,c,t'+dt,h,w],  T' = T-k+1, p = (b,h,w)
 * ------------------------------------------------------------------------------------------------ */
int cnb_pretime_conv_fwd(const float* x, const float* w1, void* u, int B, int C, int T, int H, int W, int k, int u_pitch, int dtype,
                         void* stream);
/* dw1 (fp32 [C][C][k]) is atomically accumulated: zero it first.  The network input needs no gradient. */
int cnb_pretime_conv_wgrad(const float* x, const void* du, float* dw1, int B, int C, int T, int H, int W, int k, int u_pitch, int dtype,
                           void* stream);

/* The same stage as a GEMM (throughput mode): x[B][C*T][H*W] fp32 -> pixel-major xp[B*H*W][pitch] (columns >= C*T zero), shared
 * by both temporal branches; Wt[rows][C*T] = banded expansion of w1 (rows >= C*T' zero) so that u = xp * Wt^T is cnb_conv2d_fwd
 * with a 1x1 kernel; cnb_toeplitz_fold maps the GEMM's weight gradient dWt[rows][C*T] back onto dw1[C][C][k] (overwritten). */
int cnb_time_to_pixel_major(const float* x, void* xp, int B, int CT, int64_t HW, int pitch, int dtype, void* stream);
int cnb_toeplitz_expand(const float* w1, float* wt, int C, int T, int k, int rows, void* stream);
int cnb_toeplitz_fold(const float* dwt, float* dw1, int C, int T, int k, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Skinny-N unit-stride convolution split into a 1x1 GEMM and a shift-and-add (StreamConv2d's first convolution, C -> 3 per stream,
 * nn/modules/unet_parts.py:205-221): t[p][(tap, n)] = sum_c x[p][c] w[n][c][tap] comes from cnb_conv2d_fwd with a 1x1 kernel;
 *   cnb_tap_shift_add:    out[p][n] = sum_tap t[p + off(tap)][tap*N + n]   (zero padding, off = (ky*dil - pad, kx*dil - pad))
 *   cnb_tap_shift_gather: dt[q][tap*N + n] = dout[q - off(tap)][n]          (its adjoint = the im2col of dout; padding columns zero)
 * ------------------------------------------------------------------------------------------------ */
int cnb_tap_shift_add(const void* t, void* out, int B, int H, int W, int N, int KH, int KW, int pad, int dil, int t_pitch, int out_pitch,
                      int dtype, void* stream);
int cnb_tap_shift_gather(const void* dout, void* dt, int B, int H, int W, int N, int KH, int KW, int pad, int dil, int t_pitch,
                         int out_pitch, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * TowerUNetFinalCombine + SigmoidCrisp (nn/modules/unet_parts.py:86-98, :148-193)
 *   z_t = w_t * (ha[p][t]/g[t][0] + hb[p][t]/g[t][1] + hc[p][t]/g[t][2]) + b_t,  t in {0 distance, 1 edge, 2 crop}
 *   distance = sigmoid(z_0); edge = sigmoid(z_1 / (smooth + sigmoid(crisp_gamma))); crop = sigmoid(z_2)
 * params: fp32[16] = g[3][3], w[3], b[3], crisp_gamma ; flags bit0 edge_activation, bit1 mask_activation
 * ------------------------------------------------------------------------------------------------ */
int cnb_final_combine_fwd(const void* ha, const void* hb, const void* hc, const float* params, float smooth, int flags,
                          float* distance, float* edge, float* crop, int64_t P, int dtype, void* stream);
/* dparams fp32[16] is overwritten; red_ws: fp32[32] scratch */
int cnb_final_combine_bwd(const void* ha, const void* hb, const void* hc, const float* params, float smooth, int flags,
                          const float* d_distance, const float* d_edge, const float* d_crop, void* dha, void* dhb, void* dhc,
                          float* dparams, float* red_ws, int64_t P, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Tanimoto-with-complement loss (losses/losses.py:152-218) incl. LossPreprocessing (:9-59) and the label recoding of
 * LightningModuleMixin.get_true_labels (models/lightning.py:161-207).
 * One "term" = one (prediction, target) pair reduced per sample over (C,H,W); the loss is
 *   sum_terms weight_t * mean_b 0.5*((1 - T(P,S)) + (1 - T(P',S')))
 * target_mode: 0 float targets [B,C,HW] (or [B,1,HW] broadcast when tgt_c==1)
 *              1 int64 labels [B,HW], one-hot against the channel index      (LossPreprocessing one_hot_targets)
 *              2 int64 labels [B,HW], target = (label == edge_class)         (true_edge)
 *              3 int64 labels [B,HW], target = (0 < label < edge_class)      (true_crop)
 * mask_mode:   0 none, 1 float mask [B,HW], 2 int64 mask [B,HW], 3 derived from labels: (label != -1)
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    const float* pred;      /* [B][C][HW] fp32 */
    const void* target;     /* see target_mode */
    const void* mask;       /* see mask_mode (labels for mode 3) */
    float* dpred;           /* backward only: [B][C][HW] */
    int32_t C, tgt_c;
    int32_t target_mode, mask_mode, edge_class;
    float weight;
} cnb_tanimoto_term;

/* sums: fp64 [nterms][B][4] scratch {P,S,sum t,sum p} (zeroed inside); coef: fp32 [nterms][B][4] (d loss / d{P,S,P',S'} already scaled by
 * weight/(2B)); loss: fp32[1 + nterms] = total, then each term's unweighted loss */
/* variant selects the reference's LOSS_DICT entry (models/lightning.py:38-88): 0 TanimotoComplementLoss (losses/losses.py:103-218),
 * 1 TanimotoDistLoss (:221-340; `depth` ignored), 2 CombinedLoss of the two (:62-100) */
int cnb_tanimoto_fwd(const cnb_tanimoto_term* terms, int nterms, int B, int64_t HW, float smooth, int depth, int variant,
                     double* sums, float* coef, float* loss, void* stream);
/* dpred = gscale[0] * dloss/dpred */
int cnb_tanimoto_bwd(const cnb_tanimoto_term* terms, int nterms, int B, int64_t HW, const float* coef, const float* gscale,
                     void* stream);

/* Validation counts of LightningModuleMixin._shared_eval_step (models/lightning.py:374-481) in one pass: out = fp64[12] (zeroed inside)
 * {valid pixels, sum|dist-bdist|, sum (dist-bdist)^2, edge tp/fp/fn/tn, crop tp/fp/fn/tn, unused}; predictions fp32 [n], labels int64 [n]
 * (-1 = unlabelled, excluded), a prediction is positive when > thresh (probas_to_labels, :126-136) */
int cnb_val_counts(const float* dist, const float* edge, const float* crop, const int64_t* y, const float* bdist, int64_t n,
                   int edge_class, float thresh, double* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser step over flat fp32 buffers (LightningModuleMixin.configure_optimizers, models/lightning.py:611-683; gradient
 * clipping = Trainer(gradient_clip_val=1.0), model.py:168-186).
 * ------------------------------------------------------------------------------------------------ */
/* norm_ws: fp32[1] receives sum(g^2) (zeroed inside) */
int cnb_grad_sqnorm(const float* g, int64_t n, float* norm_ws, void* stream);
/* AdamW with decoupled weight decay; clip_norm <= 0 disables clipping; hyper: device fp32[2] = {lr, step (as float, >= 1)} */
int cnb_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float beta1, float beta2, float eps,
                   float weight_decay, float grad_scale, float clip_norm, const float* norm_ws, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optional ResUNet-a block variants (constructor arguments of TowerUNet the reference's own tests use,
 * tests/test_cultionet.py:67-78: attention_weights="spatial_channel", pool_by_max=True, dropout > 0).
 * ------------------------------------------------------------------------------------------------ */
/* F.adaptive_max_pool2d(x, (Hout, Wout)) of PoolResidualConv.forward (nn/modules/convolution.py:499-503), pixel-major.
 * idx: uint8 [B,Hout,Wout,C], position of the (first) maximum inside its window as row*16 + column (windows up to 16 x 16). */
int cnb_adaptive_maxpool_fwd(const void* x, void* y, void* idx, int B, int Hin, int Win, int Hout, int Wout, int C, int dtype, void* stream);
int cnb_adaptive_maxpool_bwd(const void* dy, const void* idx, void* dx, int B, int Hin, int Win, int Hout, int Wout, int C, int dtype,
                             void* stream);

/* SetActivation("SiLU") as a stand-alone operator (the channel MLP of ChannelAttention, nn/modules/attention.py:19-52) */
int cnb_silu_fwd(const void* x, void* y, int64_t n, int dtype, void* stream);
int cnb_silu_bwd(const void* x, const void* dy, void* dx, int64_t n, int dtype, void* stream);
/* y = act(x), dx = dy * act'(x) for any CNB_ACT_* code (SetActivation, nn/modules/activations.py:5-24) */
int cnb_act_fwd(const void* x, void* y, int64_t n, int act, int dtype, void* stream);
int cnb_act_bwd(const void* x, const void* dy, void* dx, int64_t n, int act, int dtype, void* stream);

/* SpatialChannelAttention, pooling side (nn/modules/attention.py:54-63, :78-86) over x[B][HW][C]:
 *   sp[B*HW][2]  = per-pixel (mean, max) over channels; ties[B*HW] = number of channels equal to that max (torch.amax shares the
 *                  gradient between ties);
 *   ch_avg/ch_max[B][C] = per-channel mean / max over the pixels, ch_arg[B][C] = pixel index of the first maximum
 *                  (nn.AdaptiveAvgPool2d(1) / nn.AdaptiveMaxPool2d(1)).
 * ws_sum/ws_max/ws_arg: scratch [B][S][C] with S = cnb_sca_slices(B, HW, C, dtype). */
int cnb_sca_slices(int B, int HW, int C, int dtype);
int cnb_sca_pool_fwd(const void* x, float* sp, float* ties, float* ch_avg, float* ch_max, int32_t* ch_arg, float* ws_sum, float* ws_max,
                     int32_t* ws_arg, int B, int HW, int C, int dtype, void* stream);
/* dx (overwritten) from the gradients of the four pooled tensors */
int cnb_sca_pool_bwd(const void* x, const float* sp, const float* ties, const float* dsp, const float* dch_avg, const float* dch_max,
                     const int32_t* ch_arg, void* dx, int B, int HW, int C, int dtype, void* stream);
/* apply side (attention.py:118-123 + convolution.py:392-393): out = y * (1 + gamma * 0.5 * (sigmoid(cl[b][c]) + sigmoid(sl[b][p])));
 * cl fp32 [B][C] = channel logits (fc1(avg) + fc2(max)), sl fp32 [B][HW] = spatial logits (3x3 conv of sp), gamma fp32[1].
 * Backward: dy overwritten; dcl [B][C], dsl [B][HW], dgamma [1] are zeroed inside and accumulated with fp32 atomics. */
int cnb_sca_apply_fwd(const void* y, const float* cl, const float* sl, const float* gamma, void* out, int B, int HW, int C, int dtype,
                      void* stream);
int cnb_sca_apply_bwd(const void* y, const void* dout, const float* cl, const float* sl, const float* gamma, void* dy, float* dcl,
                      float* dsl, float* dgamma, int B, int HW, int C, int dtype, void* stream);

/* Dropout with a counter-based generator: rng_state = device int64[2] {seed, step counter}; cnb_rng_advance bumps the counter (once
 * per forward, also inside a captured CUDA graph), `site` separates the call sites of one step, the backward is the same call on the
 * gradient (same state and site => same mask).  cnb_dropout = nn.Dropout (natten proj_drop, convolution.py:341-350);
 * cnb_dropout2d = nn.Dropout2d, one draw per (sample, channel) (PoolResidualConv.dropout_layer, convolution.py:487, :509). */
int cnb_rng_advance(void* rng_state, void* stream);
/* natten attn_drop in training mode: neighbourhood attention whose probabilities are dropped (and rescaled by 1/(1-p)) before they
 * weight the values; one-warp-per-(pixel, head) kernels for any shape.  dqkv_acc: fp32 scratch [B,H,W,3*heads*hd] (zeroed inside). */
int cnb_na2d_dropout_fwd(const void* qkv, void* out, int B, int H, int W, int heads, int hd, int ksize, int dilation, float scale,
                         const void* rng_state, int site, float p, int dtype, void* stream);
int cnb_na2d_dropout_bwd(const void* qkv, const void* dout, float* dqkv_acc, void* dqkv, int B, int H, int W, int heads, int hd, int ksize,
                         int dilation, float scale, const void* rng_state, int site, float p, int dtype, void* stream);
int cnb_dropout(const void* x, void* out, int64_t n, const void* rng_state, int site, float p, int dtype, void* stream);
int cnb_dropout2d(const void* x, void* out, int B, int HW, int C, const void* rng_state, int site, float p, int dtype, void* stream);

/* ---- Prediction over a resident tile: the steps either side of predict_step ----
 * cnb_window_load replaces, for prediction, create_predict_dataset's windowing (data/create.py:201-214: chunks of window_size,
 * map_overlap(depth = padding, boundary = 0); data/store.py:68-90: ragged end chunks zero-padded) together with EdgeDataset.get's
 * load arithmetic (data/datasets.py:443: x / 10000 clipped to [1e-9, 1]; utils/normalize.py:78-80: (x - mean_c) / std_c):
 *   out[b][c][t][y][x] = (clip(raw / scale, lo, hi) - mean[c]) / std[c],  raw = tile[t][c][row_off_b - pad + y][col_off_b - pad + x]
 * or 0 outside the tile.  tile: device int16 [T][C][Ht][Wt]; win: device int32 [B][win_stride], row b starts with (row_off, col_off); out: device fp32
 * [B][C][T][window_size + 2 pad][window_size + 2 pad] (the x field of the reference's Data); mean / stdv: device fp32 [C] or NULL. */
int cnb_window_load(const int16_t* tile, int T, int C, int Ht, int Wt, const int32_t* win, int win_stride, int B, int window_size, int pad,
                    float scale, float lo, float hi, const float* mean, const float* stdv, float* out, void* stream);
/* cnb_predict_pack replaces LightningGTiffWriter.write_on_batch_end's per-window work (callbacks.py:176-227): halo slice
 * (get_batch_slice, :136-146), band stack (distance, edge, crop), (v * scale).clip(0, scale) -> uint16 (truncation), placed at the
 * window's offset of the 3-band mosaic (device uint16 [3][Ht][mosaic_pitch >= Wt]).  dist / edge / crop: device fp32, element (b, y, x) at
 * ptr[b * batch_stride + y * Ws + x]; win: device int32 [B][4] = (row_off, col_off, height, width), clipped to the mosaic as
 * :182-185; height 0 = filler window. */
int cnb_predict_pack(const float* dist, const float* edge, const float* crop, int64_t batch_stride, int Hs, int Ws, int pad,
                     const int32_t* win, int B, int window_size, float scale, uint16_t* mosaic, int Ht, int Wt, int mosaic_pitch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CULTIONET_B200_H */
