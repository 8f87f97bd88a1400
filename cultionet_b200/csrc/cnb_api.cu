// extern "C" entry points of libcultionet_b200.so (see include/cultionet_b200.h for the contract).
// Argument validation + launch only: no allocation, no synchronisation, everything on the caller's stream.
#include "cnb_common.cuh"
#include "k_conv_generic.cuh"
#include "k_conv_tiny.cuh"
#include "k_conv_tc.cuh"
#include "k_wgrad_tc.cuh"
#include "k_loss.cuh"
#include "k_misc.cuh"
#include "k_na.cuh"
#include "k_na_fast.cuh"
#include "k_na_tc.cuh"
#include "k_norm.cuh"
#include "k_vec.cuh"
#include "k_stream.cuh"
#include "k_pool_attn.cuh"
#include "k_tile.cuh"

using namespace cnb;

#ifdef CNB_EMU
#define CNB_SET_SMEM(kfn, bytes) ((void)0)
#else
#define CNB_SET_SMEM(kfn, bytes) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#endif

static inline int stream_grid(long n, int per_block = 256, int waves = 8) {
    return cnb_clamp_grid(cnb_div_up(n, per_block), (long)CNB_NUM_SMS * waves);
}

#ifndef CNB_EMU
// opt a streaming kernel into its dynamic shared memory once per process
template <typename K>
static inline bool stream_smem(K kfn, int bytes, bool* done) {
    if (!*done) {
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return false;
        *done = true;
    }
    return true;
}
#endif

static inline int vec_width(int dtype) { return dtype == CNB_BF16 ? 8 : 4; }
// rows that split into whole 16-byte vectors (and per-CTA channel sums that fit in shared memory); column j belongs to channel
// j / ch_div, columns past C*ch_div are row padding
static inline bool bn_vec_ok(int L, int C, int ch_div, int dtype) {
    return L % vec_width(dtype) == 0 && (long)C * ch_div <= L && C <= 4096;
}

static inline long column_stride(long total, int L, int* blocks) {
    // total threads rounded down to a multiple of L (>= L), at most ~8 waves
    long want = (long)CNB_NUM_SMS * 8 * 256;
    if (want > total) want = total;
    long stride = want / L * L;
    if (stride < L) stride = L;
    *blocks = cnb_div_up(stride, 256);
    return stride;
}
extern "C" {

int cnb_version(void) { return CNB_VERSION; }
int cnb_sm_arch(void) {
#ifdef CNB_EMU
    return 0;
#else
    return 100;
#endif
}
const char* cnb_last_error(void) { return cnb_err_buf(); }
int64_t cnb_launch_count(void) { return (int64_t)cnb_launch_counter().load(); }

// ------------------------------------------------------------------------------------------------
static int check_conv_geom(int B, int Hin, int Win, int Hout, int Wout, int KH, int KW, int stride, int pad, int dil, int transposed) {
    CNB_REQUIRE(B > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "conv: empty geometry B=%d in=%dx%d out=%dx%d", B, Hin, Win, Hout, Wout);
    CNB_REQUIRE(KH > 0 && KW > 0 && stride > 0 && dil > 0 && pad >= 0, "conv: bad kernel k=%dx%d stride=%d pad=%d dil=%d", KH, KW, stride, pad, dil);
    if (!transposed) {
        const int eh = (Hin + 2 * pad - dil * (KH - 1) - 1) / stride + 1;
        const int ew = (Win + 2 * pad - dil * (KW - 1) - 1) / stride + 1;
        CNB_REQUIRE(eh == Hout && ew == Wout, "conv: output %dx%d does not match geometry (expected %dx%d)", Hout, Wout, eh, ew);
    } else {
        CNB_REQUIRE(Hout <= (Hin - 1) * stride - 2 * pad + dil * (KH - 1) + 1 + (stride - 1) &&
                        Hout >= (Hin - 1) * stride - 2 * pad + dil * (KH - 1) + 1,
                    "transposed conv: output height %d inconsistent with input %d", Hout, Hin);
        CNB_REQUIRE(Wout <= (Win - 1) * stride - 2 * pad + dil * (KW - 1) + 1 + (stride - 1) &&
                        Wout >= (Win - 1) * stride - 2 * pad + dil * (KW - 1) + 1,
                    "transposed conv: output width %d inconsistent with input %d", Wout, Win);
    }
    return CNB_OK;
}

static int check_conv_desc(const cnb_conv_desc* d) {
    CNB_REQUIRE(d != nullptr, "conv2d_fwd: null descriptor");
    CNB_REQUIRE(d->nsrc >= 1 && d->nsrc <= CNB_MAX_SRC, "conv2d_fwd: nsrc=%d", d->nsrc);
    int rc = check_conv_geom(d->B, d->Hin, d->Win, d->Hout, d->Wout, d->KH, d->KW, d->stride, d->pad, d->dil, d->transposed);
    if (rc) return rc;
    int ctot = 0;
    for (int s = 0; s < d->nsrc; ++s) {
        CNB_REQUIRE(d->src[s] != nullptr && d->src_c[s] > 0 && d->src_stride[s] >= d->src_c[s], "conv2d_fwd: bad source %d", s);
        ctot += d->src_c[s];
    }
    CNB_REQUIRE(d->w_packed && d->N > 0 && d->w_row_stride >= ctot, "conv2d_fwd: bad weight");
    if (d->nout == 0) {
        CNB_REQUIRE(d->out && d->out_stride >= d->N, "conv2d_fwd: bad output");
    } else {
        CNB_REQUIRE(d->nout > 0 && d->nout <= CNB_MAX_SRC && d->bias == nullptr && d->stats == nullptr, "conv2d_fwd: bad split output");
        int n = 0;
        for (int i = 0; i < d->nout; ++i) {
            CNB_REQUIRE(d->out_seg[i] && d->out_seg_c[i] > 0 && d->out_seg_stride[i] >= d->out_seg_c[i], "conv2d_fwd: bad output segment %d", i);
            n += d->out_seg_c[i];
        }
        CNB_REQUIRE(n == d->N, "conv2d_fwd: output segments cover %d of %d channels", n, d->N);
    }
    return CNB_OK;
}

int cnb_conv2d_fwd_generic(const cnb_conv_desc* d, int dtype, void* stream) {
    int rc = check_conv_desc(d);
    if (rc) return rc;
    CNB_REQUIRE(d->stats == nullptr, "conv2d_fwd_generic: fused BatchNorm statistics exist only in the tcgen05 kernel");
    if (d->ep_scale) CNB_FAIL(CNB_ERR_UNSUPPORTED, "conv2d_fwd_generic: the fused BatchNorm/activation epilogue exists only in the tcgen05 kernel");
    CNB_REQUIRE(d->nout == 0, "conv2d_fwd_generic: split outputs exist only in the tcgen05 kernel");
    const long M = (long)d->B * d->Hout * d->Wout;
    dim3 grid(cnb_div_up(M, CG_BM), cnb_div_up(d->N, CG_BN));
    CNB_DISPATCH_DTYPE(dtype, { CNB_LAUNCH((conv_fwd_generic_kernel<T>), grid, dim3(256), 0, (cudaStream_t)stream, *d); });
    CNB_CHECK_LAUNCH("conv_fwd_generic_kernel");
    return CNB_OK;
}

int cnb_conv2d_tc_eligible(const cnb_conv_desc* d, int dtype) {
#ifdef CNB_EMU
    (void)d;
    (void)dtype;
    return 0;
#else
    if (check_conv_desc(d)) return 0;
    if (!tc::eligible(d, dtype)) return 0;
    return tc::stats_ok(d) ? 2 : 1;  // 2: the epilogue can also produce BatchNorm's per-channel sums (cnb_conv_desc::stats)
#endif
}

int cnb_conv2d_fwd_tc(const cnb_conv_desc* d, int dtype, void* stream) {
#ifdef CNB_EMU
    (void)d;
    (void)dtype;
    (void)stream;
    CNB_FAIL(CNB_ERR_UNSUPPORTED, "the tcgen05 kernel exists only in the sm_100a build");
#else
    int rc = check_conv_desc(d);
    if (rc) return rc;
    if (!tc::eligible(d, dtype)) CNB_FAIL(CNB_ERR_UNSUPPORTED, "conv2d_fwd_tc: shape/dtype not eligible for the tcgen05 kernel");
    rc = tc::conv_tc_launch(d, (cudaStream_t)stream);
    if (rc) CNB_FAIL(CNB_ERR_CUDA, "conv2d_fwd_tc: tensor-map encode or launch configuration failed (%d) %s", rc, tc::encode_diag());
    CNB_CHECK_LAUNCH("conv_tc_kernel");
    return CNB_OK;
#endif
}

int cnb_conv2d_fwd_tiny(const cnb_conv_desc* d, int dtype, void* stream) {
    int rc = check_conv_desc(d);
    if (rc) return rc;
    if (!conv_tiny_eligible(d)) CNB_FAIL(CNB_ERR_UNSUPPORTED, "conv2d_fwd_tiny: needs N <= %d and at most %d input channels", TINY_MAX_N, TINY_MAX_C);
    CNB_REQUIRE(d->stats == nullptr, "conv2d_fwd_tiny: fused BatchNorm statistics exist only in the tcgen05 kernel");
    if (d->ep_scale) CNB_FAIL(CNB_ERR_UNSUPPORTED, "conv2d_fwd_tiny: the fused BatchNorm/activation epilogue exists only in the tcgen05 kernel");
    int ctot = 0;
    for (int s = 0; s < d->nsrc; ++s) ctot += d->src_c[s];
    const long M = (long)d->B * d->Hout * d->Wout;
    if (d->nsrc == 1 && ((d->N == 3 && (ctot == 9 || ctot == 3)) || (d->N == 9 && ctot == 3))) {
        // the Psi-Net head shapes: compile-time channel counts (see conv_tiny_fixed_kernel)
        const dim3 grid(stream_grid(M, 256, 8));
        CNB_DISPATCH_DTYPE(dtype, {
            if (d->N == 3 && ctot == 9)
                CNB_LAUNCH((conv_tiny_fixed_kernel<T, 3, 9>), grid, dim3(256), 0, (cudaStream_t)stream, *d);
            else if (d->N == 3)
                CNB_LAUNCH((conv_tiny_fixed_kernel<T, 3, 3>), grid, dim3(256), 0, (cudaStream_t)stream, *d);
            else
                CNB_LAUNCH((conv_tiny_fixed_kernel<T, 9, 3>), grid, dim3(256), 0, (cudaStream_t)stream, *d);
        });
        CNB_CHECK_LAUNCH("conv_tiny_fixed_kernel");
        return CNB_OK;
    }
    CNB_DISPATCH_DTYPE(dtype, { CNB_LAUNCH((conv_tiny_kernel<T>), dim3(stream_grid(M, 256, 4)), dim3(256), 0, (cudaStream_t)stream, *d, ctot); });
    CNB_CHECK_LAUNCH("conv_tiny_kernel");
    return CNB_OK;
}

int cnb_conv2d_fwd(const cnb_conv_desc* d, int dtype, void* stream) {
    if (d && (d->stats || d->ep_scale)) return cnb_conv2d_fwd_tc(d, dtype, stream);
    if (check_conv_desc(d) == CNB_OK && d->nout == 0 && conv_tiny_eligible(d)) return cnb_conv2d_fwd_tiny(d, dtype, stream);
    if (cnb_conv2d_tc_eligible(d, dtype)) return cnb_conv2d_fwd_tc(d, dtype, stream);
    return cnb_conv2d_fwd_generic(d, dtype, stream);
}

static int check_wgrad_desc(const cnb_wgrad_desc* d) {
    CNB_REQUIRE(d != nullptr, "conv2d_wgrad: null descriptor");
    int rc = check_conv_geom(d->B, d->Hin, d->Win, d->Hout, d->Wout, d->KH, d->KW, d->stride, d->pad, d->dil, d->transposed);
    if (rc) return rc;
    CNB_REQUIRE(d->src && d->dy && d->dwp && d->src_c > 0 && d->N > 0 && d->k_off >= 0 && d->k_off + d->src_c <= d->Ctot, "conv2d_wgrad: bad operands");
    return CNB_OK;
}

int cnb_conv2d_wgrad_generic(const cnb_wgrad_desc* d, int dtype, void* stream) {
    int rc = check_wgrad_desc(d);
    if (rc) return rc;
    const long M = (long)d->B * d->Hout * d->Wout;
    const int taps = d->KH * d->KW;
    const int tiles = cnb_div_up(d->src_c, CG_BM) * cnb_div_up(d->N, CG_BN) * taps;
    // enough pixel splits for ~4 waves of CTAs, each split a multiple of the 16-pixel chunk
    int splits = cnb_clamp_grid((4L * CNB_NUM_SMS + tiles - 1) / tiles, cnb_div_up(M, 4 * CG_BK));
    long m_per_split = ((M + splits - 1) / splits + CG_BK - 1) / CG_BK * CG_BK;
    splits = cnb_div_up(M, m_per_split);
    CNB_REQUIRE((long)taps * splits <= 65535, "conv2d_wgrad: grid.z overflow");
    dim3 grid(cnb_div_up(d->src_c, CG_BM), cnb_div_up(d->N, CG_BN), taps * splits);
    CNB_DISPATCH_DTYPE(dtype, { CNB_LAUNCH((conv_wgrad_generic_kernel<T>), grid, dim3(256), 0, (cudaStream_t)stream, *d, splits, m_per_split); });
    CNB_CHECK_LAUNCH("conv_wgrad_generic_kernel");
    return CNB_OK;
}

int cnb_conv2d_wgrad_tc_eligible(const cnb_wgrad_desc* d, int dtype) {
#ifdef CNB_EMU
    (void)d;
    (void)dtype;
    return 0;
#else
    if (check_wgrad_desc(d)) return 0;
    return tc::wgrad_eligible(d, dtype) ? 1 : 0;
#endif
}

int cnb_conv2d_wgrad_tc(const cnb_wgrad_desc* d, int dtype, void* stream) {
#ifdef CNB_EMU
    (void)d;
    (void)dtype;
    (void)stream;
    CNB_FAIL(CNB_ERR_UNSUPPORTED, "the tcgen05 kernel exists only in the sm_100a build");
#else
    int rc = check_wgrad_desc(d);
    if (rc) return rc;
    if (!tc::wgrad_eligible(d, dtype)) CNB_FAIL(CNB_ERR_UNSUPPORTED, "conv2d_wgrad_tc: shape/dtype not eligible for the tcgen05 kernel");
    rc = tc::wgrad_tc_launch(d, (cudaStream_t)stream);
    if (rc) CNB_FAIL(CNB_ERR_CUDA, "conv2d_wgrad_tc: tensor-map encode or launch configuration failed (%d)", rc);
    CNB_CHECK_LAUNCH("wgrad_tc_kernel");
    return CNB_OK;
#endif
}

int cnb_conv2d_wgrad_tiny(const cnb_wgrad_desc* d, int dtype, void* stream) {
    int rc = check_wgrad_desc(d);
    if (rc) return rc;
    if (!wgrad_tiny_eligible(d)) CNB_FAIL(CNB_ERR_UNSUPPORTED, "conv2d_wgrad_tiny: needs N <= %d and a source slice of at most %d channels", TINY_WG_MAX_N, TINY_MAX_C);
    const long M = (long)d->B * d->Hout * d->Wout;
    if (d->N == 3 && (d->src_c == 9 || d->src_c == 3)) {
        // the Psi-Net head shapes: TG taps' products in registers, taps / TG passes over the level (see conv_tiny_wgrad_fixed_kernel)
        const int taps = d->KH * d->KW;
        CNB_DISPATCH_DTYPE(dtype, {
            if (d->src_c == 9)
                CNB_LAUNCH((conv_tiny_wgrad_fixed_kernel<T, 3, 9, 3>), dim3(stream_grid(M, 256, 1), cnb_div_up(taps, 3)), dim3(256), 0,
                           (cudaStream_t)stream, *d);
            else
                CNB_LAUNCH((conv_tiny_wgrad_fixed_kernel<T, 3, 3, 9>), dim3(stream_grid(M, 256, 2), cnb_div_up(taps, 9)), dim3(256), 0,
                           (cudaStream_t)stream, *d);
        });
        CNB_CHECK_LAUNCH("conv_tiny_wgrad_fixed_kernel");
        return CNB_OK;
    }
    dim3 grid(stream_grid(M, 256, 1), d->KH * d->KW, cnb_div_up(d->src_c, TINY_WG_MAX_C));
    CNB_DISPATCH_DTYPE(dtype, { CNB_LAUNCH((conv_tiny_wgrad_kernel<T>), grid, dim3(256), 0, (cudaStream_t)stream, *d); });
    CNB_CHECK_LAUNCH("conv_tiny_wgrad_kernel");
    return CNB_OK;
}

int cnb_conv2d_wgrad(const cnb_wgrad_desc* d, int dtype, void* stream) {
    if (check_wgrad_desc(d) == CNB_OK && wgrad_tiny_eligible(d)) return cnb_conv2d_wgrad_tiny(d, dtype, stream);
    if (cnb_conv2d_wgrad_tc_eligible(d, dtype)) return cnb_conv2d_wgrad_tc(d, dtype, stream);
    return cnb_conv2d_wgrad_generic(d, dtype, stream);
}

int cnb_repitch(const void* src, int src_stride, void* dst, int dst_stride, int64_t P, int C, int dtype, void* stream) {
    CNB_REQUIRE(src && dst && P > 0 && C > 0 && src_stride >= C && dst_stride >= C, "repitch: bad arguments");
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((repitch_kernel<T>), dim3(stream_grid((long)P * dst_stride)), dim3(256), 0, (cudaStream_t)stream, (const T*)src, src_stride,
                   (T*)dst, dst_stride, (long)P, C);
    });
    CNB_CHECK_LAUNCH("repitch_kernel");
    return CNB_OK;
}

int cnb_multi_copy(const void* table, int n_entries, int max_len, void* stream) {
    CNB_REQUIRE(table && n_entries > 0 && max_len > 0, "multi_copy: bad arguments");
    const MultiCopyEntry* e = (const MultiCopyEntry*)table;  // HOST memory: copied into the launch arguments below
    const int bx = cnb_clamp_grid(cnb_div_up(max_len, 256), 64);
    for (int i = 0; i < n_entries; i += MULTI_COPY_BATCH) {
        const int n = n_entries - i < MULTI_COPY_BATCH ? n_entries - i : MULTI_COPY_BATCH;
        MultiCopyBatch batch;
        memcpy(batch.e, e + i, sizeof(MultiCopyEntry) * n);
        CNB_LAUNCH(multi_copy_kernel, dim3(bx, n), dim3(256), 0, (cudaStream_t)stream, batch);
    }
    CNB_CHECK_LAUNCH("multi_copy_kernel");
    return CNB_OK;
}

int cnb_broadcast_pixels(const void* e, void* out, int B, int64_t HW, int C, int dtype, void* stream) {
    CNB_REQUIRE(e && out && B > 0 && HW > 0 && C > 0, "broadcast_pixels: bad arguments");
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((broadcast_pixels_kernel<T>), dim3(stream_grid((long)B * HW * C)), dim3(256), 0, (cudaStream_t)stream, (const T*)e, (T*)out, B,
                   (long)HW, C);
    });
    CNB_CHECK_LAUNCH("broadcast_pixels_kernel");
    return CNB_OK;
}

int cnb_pack_weight(const float* w, void* wp, int dtype, int taps, int N, int K, int wp_pitch, int64_t s_n, int64_t s_k, int64_t s_tap,
                    void* stream) {
    CNB_REQUIRE(w && wp && taps > 0 && N > 0 && K > 0 && wp_pitch >= K, "pack_weight: bad arguments");
    const long total = (long)taps * N * wp_pitch;
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((pack_weight_kernel<T>), dim3(stream_grid(total)), dim3(256), 0, (cudaStream_t)stream, w, (T*)wp, taps, N, K, wp_pitch,
                   (long)s_n, (long)s_k, (long)s_tap);
    });
    CNB_CHECK_LAUNCH("pack_weight_kernel");
    return CNB_OK;
}

int cnb_pack_weight2(const float* w, void* wp, void* wd, int dtype, int taps, int N, int K, int pitch_k, int pitch_n, int64_t s_n,
                     int64_t s_k, int64_t s_tap, void* stream) {
    CNB_REQUIRE(w && wp && taps > 0 && taps <= 32 && N > 0 && K > 0 && pitch_k >= K && (!wd || pitch_n >= N), "pack_weight2: bad arguments");
    const size_t smem = (size_t)taps * PW_T * (PW_T + 1) * sizeof(float);
    const dim3 grid(cnb_div_up(pitch_k > K ? pitch_k : K, PW_T), cnb_div_up(wd && pitch_n > N ? pitch_n : N, PW_T));
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_SET_SMEM((pack_weight_tiled_kernel<T>), smem);
        CNB_LAUNCH((pack_weight_tiled_kernel<T>), grid, dim3(256), smem, (cudaStream_t)stream, w, (T*)wp, (T*)wd, taps, N, K, pitch_k, pitch_n,
                   (long)s_n, (long)s_k, (long)s_tap);
    });
    CNB_CHECK_LAUNCH("pack_weight_tiled_kernel");
    return CNB_OK;
}

int cnb_pack_weights_batched(const cnb_pack_desc* table, int ndesc, int total_tiles, int max_taps, int dtype, void* stream) {
    CNB_REQUIRE(table && ndesc > 0 && total_tiles > 0 && max_taps > 0 && max_taps <= 32, "pack_weights_batched: bad arguments");
    const size_t smem = (size_t)max_taps * PW_T * (PW_T + 1) * sizeof(float);
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_SET_SMEM((pack_weight_batched_kernel<T>), smem);
        CNB_LAUNCH((pack_weight_batched_kernel<T>), dim3(total_tiles), dim3(256), smem, (cudaStream_t)stream, table, ndesc);
    });
    CNB_CHECK_LAUNCH("pack_weight_batched_kernel");
    return CNB_OK;
}

int cnb_unpack_wgrad(float* dwp, float* g, int taps, int N, int K, int64_t s_n, int64_t s_k, int64_t s_tap, int mode, void* stream) {
    CNB_REQUIRE(dwp && g && taps > 0 && N > 0 && K > 0, "unpack_wgrad: bad arguments");
    if (taps <= 32) {
        const size_t smem = (size_t)taps * PW_T * (PW_T + 1) * sizeof(float);
        CNB_SET_SMEM(unpack_wgrad_tiled_kernel, smem);
        CNB_LAUNCH(unpack_wgrad_tiled_kernel, dim3(cnb_div_up(K, PW_T), cnb_div_up(N, PW_T)), dim3(256), smem, (cudaStream_t)stream, dwp, g, taps,
                   N, K, (long)s_n, (long)s_k, (long)s_tap, mode);
        CNB_CHECK_LAUNCH("unpack_wgrad_tiled_kernel");
        return CNB_OK;
    }
    const long total = (long)taps * N * K;
    CNB_LAUNCH(unpack_wgrad_kernel, dim3(stream_grid(total)), dim3(256), 0, (cudaStream_t)stream, dwp, g, taps, N, K, (long)s_n, (long)s_k,
               (long)s_tap, mode);
    CNB_CHECK_LAUNCH("unpack_wgrad_kernel");
    return CNB_OK;
}

int cnb_unpack_wgrads_batched(const cnb_pack_desc* table, int ndesc, int total_tiles, int max_taps, void* stream) {
    CNB_REQUIRE(table && ndesc > 0 && total_tiles > 0 && max_taps > 0 && max_taps <= 32, "unpack_wgrads_batched: bad arguments");
    const size_t smem = (size_t)max_taps * PW_T * (PW_T + 1) * sizeof(float);
    CNB_SET_SMEM(unpack_wgrad_batched_kernel, smem);
    CNB_LAUNCH(unpack_wgrad_batched_kernel, dim3(total_tiles), dim3(256), smem, (cudaStream_t)stream, table, ndesc);
    CNB_CHECK_LAUNCH("unpack_wgrad_batched_kernel");
    return CNB_OK;
}

int cnb_bias_grad(const void* dy, int dy_stride, int64_t P, int N, float* db, int accumulate, int dtype, void* stream) {
    CNB_REQUIRE(dy && db && P > 0 && N > 0 && dy_stride >= N, "bias_grad: bad arguments");
    if (!accumulate) CNB_MEMSET_ASYNC(db, 0, sizeof(float) * N, (cudaStream_t)stream);
    if (dy_stride == N && N % vec_width(dtype) == 0 && N <= 8192 && cnb_aligned16(dy)) {
        const int V = vec_width(dtype), CV = N / V;
        int blocks;
        const long total_v = (long)P * CV;
        const long stride_v = column_stride(total_v, CV, &blocks);
        CNB_DISPATCH_DTYPE(dtype, {
            CNB_LAUNCH((colsum_vec_kernel<T>), dim3(blocks), dim3(256), N * sizeof(float), (cudaStream_t)stream, (const T*)dy, total_v, CV,
                       stride_v, N, db);
        });
        CNB_CHECK_LAUNCH("colsum_vec_kernel");
        return CNB_OK;
    }
    const int cols = cnb_div_up(N, 32);
    const int splits = cnb_clamp_grid((4L * CNB_NUM_SMS + cols - 1) / cols, cnb_div_up(P, 64));
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((bias_grad_kernel<T>), dim3(cols, splits), dim3(256), 0, (cudaStream_t)stream, (const T*)dy, dy_stride, (long)P, N, db);
    });
    CNB_CHECK_LAUNCH("bias_grad_kernel");
    return CNB_OK;
}

// ------------------------------------------------------------------------------------------------
int cnb_bn_stats(const void* x, int64_t P, int L, int C, int ch_div, float* sums, int dtype, void* stream) {
    CNB_REQUIRE(x && sums && P > 0 && L > 0 && C > 0 && ch_div > 0, "bn_stats: bad arguments");
    CNB_MEMSET_ASYNC(sums, 0, sizeof(float) * 2 * C, (cudaStream_t)stream);
    int blocks;
    const long total = (long)P * L;
#ifndef CNB_EMU
    if (st::eligible(L, C, ch_div, dtype, total / 8) && cnb_aligned16(x)) {
        static bool cfg = false;
        const int smem = (st::smem_bytes<1, st::S1>()) + 2 * C * (int)sizeof(float);
        if (stream_smem(st::bn_stats_stream_kernel, (st::smem_bytes<1, st::S1>()) + 2 * 2048 * (int)sizeof(float), &cfg)) {
            CNB_LAUNCH(st::bn_stats_stream_kernel, dim3(st::grid_reduce(total / 8)), dim3(st::THREADS), smem, (cudaStream_t)stream, (const bf16_t*)x,
                       total / 8, L / 8, C, ch_div, sums);
            CNB_CHECK_LAUNCH("bn_stats_stream_kernel");
            return CNB_OK;
        }
    }
#endif
    if (bn_vec_ok(L, C, ch_div, dtype) && cnb_aligned16(x)) {
        const int V = vec_width(dtype), CV = L / V;
        const long total_v = total / V;
        const long stride_v = column_stride(total_v, CV, &blocks);
        CNB_DISPATCH_DTYPE(dtype, {
            CNB_LAUNCH((bn_stats_vec_kernel<T>), dim3(blocks), dim3(256), 2 * C * sizeof(float), (cudaStream_t)stream, (const T*)x, total_v, CV,
                       stride_v, C, ch_div, sums);
        });
        CNB_CHECK_LAUNCH("bn_stats_vec_kernel");
        return CNB_OK;
    }
    const long stride = column_stride(total, L, &blocks);
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((bn_stats_kernel<T>), dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const T*)x, total, L, C, ch_div, stride, sums);
    });
    CNB_CHECK_LAUNCH("bn_stats_kernel");
    return CNB_OK;
}

int cnb_bn_finalize(const float* sums, int64_t count, int C, const float* gamma, const float* beta, float eps, float momentum,
                    float* running_mean, float* running_var, float* save_mean, float* save_rstd, float* scale, float* shift, void* stream) {
    CNB_REQUIRE(C > 0 && scale && shift, "bn_finalize: bad arguments");
    CNB_REQUIRE(sums || (running_mean && running_var), "bn_finalize: eval mode needs running statistics");
    CNB_REQUIRE(!sums || count > 0, "bn_finalize: count must be positive");
    CNB_LAUNCH(bn_finalize_kernel, dim3(cnb_div_up(C, 128)), dim3(128), 0, (cudaStream_t)stream, sums, (long)count, C, gamma, beta, eps,
               momentum, running_mean, running_var, save_mean, save_rstd, scale, shift);
    CNB_CHECK_LAUNCH("bn_finalize_kernel");
    return CNB_OK;
}

int cnb_bn_act_fwd(const void* x, const float* scale, const float* shift, const void* residual, void* y, int64_t P, int L, int C, int ch_div,
                   int act, int dtype, void* stream) {
    CNB_REQUIRE(x && y && scale && shift && P > 0 && L > 0 && C > 0 && ch_div > 0, "bn_act_fwd: bad arguments");
    const long total = (long)P * L;
#ifndef CNB_EMU
    if (st::eligible(L, C, ch_div, dtype, total / 8) && cnb_aligned16(x) && cnb_aligned16(y) && cnb_aligned16(residual)) {
        static bool cfg0 = false, cfg1 = false;
        if (residual) {
            if (stream_smem(st::bn_act_fwd_stream_kernel<true>, (st::smem_bytes<2, st::S2>()), &cfg1)) {
                CNB_LAUNCH(st::bn_act_fwd_stream_kernel<true>, dim3(st::grid(total / 8)), dim3(st::THREADS), (st::smem_bytes<2, st::S2>()),
                           (cudaStream_t)stream, (const bf16_t*)x, scale, shift, (const bf16_t*)residual, (bf16_t*)y, total / 8, L / 8, C, ch_div,
                           act);
                CNB_CHECK_LAUNCH("bn_act_fwd_stream_kernel<res>");
                return CNB_OK;
            }
        } else if (stream_smem(st::bn_act_fwd_stream_kernel<false>, (st::smem_bytes<1, st::S1>()), &cfg0)) {
            CNB_LAUNCH(st::bn_act_fwd_stream_kernel<false>, dim3(st::grid(total / 8)), dim3(st::THREADS), (st::smem_bytes<1, st::S1>()),
                       (cudaStream_t)stream, (const bf16_t*)x, scale, shift, (const bf16_t*)nullptr, (bf16_t*)y, total / 8, L / 8, C, ch_div, act);
            CNB_CHECK_LAUNCH("bn_act_fwd_stream_kernel");
            return CNB_OK;
        }
    }
#endif
    if (bn_vec_ok(L, C, ch_div, dtype) && cnb_aligned16(x) && cnb_aligned16(y) && cnb_aligned16(residual)) {
        const int V = vec_width(dtype), CV = L / V;
        int blocks;
        const long total_v = total / V;
        const long stride_v = column_stride(total_v, CV, &blocks);
        CNB_DISPATCH_DTYPE(dtype, {
            CNB_LAUNCH((bn_act_fwd_vec_kernel<T>), dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const T*)x, scale, shift,
                       (const T*)residual, (T*)y, total_v, CV, stride_v, C, ch_div, act);
        });
        CNB_CHECK_LAUNCH("bn_act_fwd_vec_kernel");
        return CNB_OK;
    }
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((bn_act_fwd_kernel<T>), dim3(stream_grid(total)), dim3(256), 0, (cudaStream_t)stream, (const T*)x, scale, shift,
                   (const T*)residual, (T*)y, total, L, C, ch_div, act);
    });
    CNB_CHECK_LAUNCH("bn_act_fwd_kernel");
    return CNB_OK;
}

int cnb_bn_train_fwd(const void* x, const float* sums, int64_t count, const float* gamma, const float* beta, float eps, float momentum,
                     float* running_mean, float* running_var, float* save_mean, float* save_rstd, float* scale, float* shift,
                     const void* residual, void* y, int64_t P, int L, int C, int ch_div, int act, int dtype, void* stream) {
    CNB_REQUIRE(x && y && sums && save_mean && save_rstd && scale && shift && count > 0 && P > 0 && L > 0 && C > 0 && ch_div > 0,
                "bn_train_fwd: bad arguments");
    CNB_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_train_fwd: running statistics come in pairs");
#ifndef CNB_EMU
    const long total = (long)P * L;
    if (st::eligible(L, C, ch_div, dtype, total / 8) && cnb_aligned16(x) && cnb_aligned16(y) && cnb_aligned16(residual)) {
        static bool cfg0 = false, cfg1 = false;
        if (residual) {
            if (stream_smem(st::bn_train_fwd_stream_kernel<true>, (st::smem_bytes<2, st::S2>()), &cfg1)) {
                CNB_LAUNCH(st::bn_train_fwd_stream_kernel<true>, dim3(st::grid(total / 8)), dim3(st::THREADS), (st::smem_bytes<2, st::S2>()),
                           (cudaStream_t)stream, (const bf16_t*)x, sums, (long)count, gamma, beta, eps, momentum, running_mean, running_var,
                           save_mean, save_rstd, scale, shift, (const bf16_t*)residual, (bf16_t*)y, total / 8, L / 8, C, ch_div, act);
                CNB_CHECK_LAUNCH("bn_train_fwd_stream_kernel<res>");
                return CNB_OK;
            }
        } else if (stream_smem(st::bn_train_fwd_stream_kernel<false>, (st::smem_bytes<1, st::S1>()), &cfg0)) {
            CNB_LAUNCH(st::bn_train_fwd_stream_kernel<false>, dim3(st::grid(total / 8)), dim3(st::THREADS), (st::smem_bytes<1, st::S1>()),
                       (cudaStream_t)stream, (const bf16_t*)x, sums, (long)count, gamma, beta, eps, momentum, running_mean, running_var,
                       save_mean, save_rstd, scale, shift, (const bf16_t*)nullptr, (bf16_t*)y, total / 8, L / 8, C, ch_div, act);
            CNB_CHECK_LAUNCH("bn_train_fwd_stream_kernel");
            return CNB_OK;
        }
    }
#endif
    int rc = cnb_bn_finalize(sums, count, C, gamma, beta, eps, momentum, running_mean, running_var, save_mean, save_rstd, scale, shift, stream);
    if (rc) return rc;
    return cnb_bn_act_fwd(x, scale, shift, residual, y, P, L, C, ch_div, act, dtype, stream);
}

static int bn_act_bwd_reduce_impl(const void* x, const void* dy, const float* save_mean, const float* save_rstd, const float* gamma,
                                  const float* beta, int64_t P, int L, int C, int ch_div, int act, float* dsums, int dtype, void* stream,
                                  bool clear);

int cnb_bn_act_bwd_reduce(const void* x, const void* dy, const float* save_mean, const float* save_rstd, const float* gamma, const float* beta,
                          int64_t P, int L, int C, int ch_div, int act, float* dsums, int dtype, void* stream) {
    return bn_act_bwd_reduce_impl(x, dy, save_mean, save_rstd, gamma, beta, P, L, C, ch_div, act, dsums, dtype, stream, true);
}

int cnb_bn_act_bwd_reduce_acc(const void* x, const void* dy, const float* save_mean, const float* save_rstd, const float* gamma,
                              const float* beta, int64_t P, int L, int C, int ch_div, int act, float* dsums, int dtype, void* stream) {
    return bn_act_bwd_reduce_impl(x, dy, save_mean, save_rstd, gamma, beta, P, L, C, ch_div, act, dsums, dtype, stream, false);
}

static int bn_act_bwd_reduce_impl(const void* x, const void* dy, const float* save_mean, const float* save_rstd, const float* gamma,
                                  const float* beta, int64_t P, int L, int C, int ch_div, int act, float* dsums, int dtype, void* stream,
                                  bool clear) {
    CNB_REQUIRE(x && dy && save_mean && save_rstd && dsums && P > 0 && L > 0 && C > 0 && ch_div > 0, "bn_act_bwd_reduce: bad arguments");
    if (clear) CNB_MEMSET_ASYNC(dsums, 0, sizeof(float) * 2 * C, (cudaStream_t)stream);
    int blocks;
    const long total = (long)P * L;
#ifndef CNB_EMU
    if (st::eligible(L, C, ch_div, dtype, total / 8) && cnb_aligned16(x) && cnb_aligned16(dy)) {
        static bool cfg = false;
        const int smem = (st::smem_bytes<2, st::S2>()) + 2 * C * (int)sizeof(float);
        if (stream_smem(st::bn_act_bwd_reduce_stream_kernel, (st::smem_bytes<2, st::S2>()) + 2 * 2048 * (int)sizeof(float), &cfg)) {
            CNB_LAUNCH(st::bn_act_bwd_reduce_stream_kernel, dim3(st::grid_reduce(total / 8)), dim3(st::THREADS), smem, (cudaStream_t)stream,
                       (const bf16_t*)x, (const bf16_t*)dy, save_mean, save_rstd, gamma, beta, total / 8, L / 8, C, ch_div, act, dsums);
            CNB_CHECK_LAUNCH("bn_act_bwd_reduce_stream_kernel");
            return CNB_OK;
        }
    }
#endif
    if (bn_vec_ok(L, C, ch_div, dtype) && cnb_aligned16(x) && cnb_aligned16(dy)) {
        const int V = vec_width(dtype), CV = L / V;
        const long total_v = total / V;
        const long stride_v = column_stride(total_v, CV, &blocks);
        CNB_DISPATCH_DTYPE(dtype, {
            CNB_LAUNCH((bn_act_bwd_reduce_vec_kernel<T>), dim3(blocks), dim3(256), 2 * C * sizeof(float), (cudaStream_t)stream, (const T*)x,
                       (const T*)dy, save_mean, save_rstd, gamma, beta, total_v, CV, stride_v, C, ch_div, act, dsums);
        });
        CNB_CHECK_LAUNCH("bn_act_bwd_reduce_vec_kernel");
        return CNB_OK;
    }
    const long stride = column_stride(total, L, &blocks);
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((bn_act_bwd_reduce_kernel<T>), dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const T*)x, (const T*)dy, save_mean,
                   save_rstd, gamma, beta, total, L, C, ch_div, act, stride, dsums);
    });
    CNB_CHECK_LAUNCH("bn_act_bwd_reduce_kernel");
    return CNB_OK;
}

int cnb_bn_act_bwd_apply(const void* x, const void* dy, const float* save_mean, const float* save_rstd, const float* gamma, const float* beta,
                         const float* dsums, int64_t count, void* dx, int64_t P, int L, int C, int ch_div, int act, int train_stats, int dtype,
                         void* stream) {
    CNB_REQUIRE(x && dy && dx && save_mean && save_rstd && dsums && count > 0 && P > 0 && L > 0 && C > 0 && ch_div > 0,
                "bn_act_bwd_apply: bad arguments");
    const long total = (long)P * L;
#ifndef CNB_EMU
    if (st::eligible(L, C, ch_div, dtype, total / 8) && cnb_aligned16(x) && cnb_aligned16(dy) && cnb_aligned16(dx)) {
        static bool cfg = false;
        if (stream_smem(st::bn_act_bwd_apply_stream_kernel, (st::smem_bytes<2, st::S2>()), &cfg)) {
            CNB_LAUNCH(st::bn_act_bwd_apply_stream_kernel, dim3(st::grid(total / 8)), dim3(st::THREADS), (st::smem_bytes<2, st::S2>()),
                       (cudaStream_t)stream, (const bf16_t*)x, (const bf16_t*)dy, save_mean, save_rstd, gamma, beta, dsums,
                       1.0f / (float)count, (bf16_t*)dx, total / 8, L / 8, C, ch_div, act, train_stats);
            CNB_CHECK_LAUNCH("bn_act_bwd_apply_stream_kernel");
            return CNB_OK;
        }
    }
#endif
    if (bn_vec_ok(L, C, ch_div, dtype) && cnb_aligned16(x) && cnb_aligned16(dy) && cnb_aligned16(dx)) {
        const int V = vec_width(dtype), CV = L / V;
        int blocks;
        const long total_v = total / V;
        const long stride_v = column_stride(total_v, CV, &blocks);
        CNB_DISPATCH_DTYPE(dtype, {
            CNB_LAUNCH((bn_act_bwd_apply_vec_kernel<T>), dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const T*)x, (const T*)dy,
                       save_mean, save_rstd, gamma, beta, dsums, 1.0f / (float)count, (T*)dx, total_v, CV, stride_v, C, ch_div, act,
                       train_stats);
        });
        CNB_CHECK_LAUNCH("bn_act_bwd_apply_vec_kernel");
        return CNB_OK;
    }
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((bn_act_bwd_apply_kernel<T>), dim3(stream_grid(total)), dim3(256), 0, (cudaStream_t)stream, (const T*)x, (const T*)dy,
                   save_mean, save_rstd, gamma, beta, dsums, 1.0f / (float)count, (T*)dx, total, L, C, ch_div, act, train_stats);
    });
    CNB_CHECK_LAUNCH("bn_act_bwd_apply_kernel");
    return CNB_OK;
}

int cnb_add_n(const void* a, const void* b, const void* c, const void* d, void* out, int64_t n, int dtype, void* stream) {
    CNB_REQUIRE(a && b && out && n > 0, "add_n: bad arguments");
    if (n % vec_width(dtype) == 0 && cnb_aligned16(a) && cnb_aligned16(b) && cnb_aligned16(c) && cnb_aligned16(d) && cnb_aligned16(out)) {
        const long n_v = n / vec_width(dtype);
        CNB_DISPATCH_DTYPE(dtype, {
            CNB_LAUNCH((add_n_vec_kernel<T>), dim3(stream_grid(n_v)), dim3(256), 0, (cudaStream_t)stream, (const T*)a, (const T*)b,
                       (const T*)c, (const T*)d, (T*)out, n_v);
        });
        CNB_CHECK_LAUNCH("add_n_vec_kernel");
        return CNB_OK;
    }
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((add_n_kernel<T>), dim3(stream_grid(n)), dim3(256), 0, (cudaStream_t)stream, (const T*)a, (const T*)b, (const T*)c,
                   (const T*)d, (T*)out, (long)n);
    });
    CNB_CHECK_LAUNCH("add_n_kernel");
    return CNB_OK;
}

// ------------------------------------------------------------------------------------------------
int cnb_layernorm_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y, float* save_mean, float* save_rstd, int64_t P,
                      int C, int dtype, void* stream) {
    CNB_REQUIRE(x && y && gamma && beta && save_mean && save_rstd && P > 0 && C > 0, "layernorm_fwd: bad arguments");
    if (C % vec_width(dtype) == 0 && C <= 128 * vec_width(dtype) && cnb_aligned16(x) && cnb_aligned16(y)) {
        const int K = cnb_div_up(C, 32 * vec_width(dtype));
        const dim3 grid(stream_grid(P, 8 * (K == 1 ? 4 : 2), 3));  // U pixels per warp and trip, 3 CTAs per SM (<= 85 registers)
#define CNB_LN_FWD(KK)                                                                                                              \
    CNB_DISPATCH_DTYPE(dtype, {                                                                                                     \
        CNB_LAUNCH((layernorm_fwd_vec_kernel<T, KK>), grid, dim3(256), 0, (cudaStream_t)stream, (const T*)x, gamma, beta, eps, (T*)y, \
                   save_mean, save_rstd, (long)P, C);                                                                               \
    })
        if (K == 1)
            CNB_LN_FWD(1);
        else if (K == 2)
            CNB_LN_FWD(2);
        else
            CNB_LN_FWD(4);
#undef CNB_LN_FWD
        CNB_CHECK_LAUNCH("layernorm_fwd_vec_kernel");
        return CNB_OK;
    }
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((layernorm_fwd_kernel<T>), dim3(stream_grid(P, 8)), dim3(256), 0, (cudaStream_t)stream, (const T*)x, gamma, beta, eps, (T*)y,
                   save_mean, save_rstd, (long)P, C);
    });
    CNB_CHECK_LAUNCH("layernorm_fwd_kernel");
    return CNB_OK;
}

int cnb_layernorm_bwd(const void* x, const void* dy, const float* gamma, const float* save_mean, const float* save_rstd, void* dx,
                      float* dgamma, float* dbeta, int64_t P, int C, int dtype, void* stream) {
    CNB_REQUIRE(x && dy && dx && gamma && save_mean && save_rstd && dgamma && dbeta && P > 0 && C > 0, "layernorm_bwd: bad arguments");
    CNB_REQUIRE(C <= 32 * LN_MAX_CPL, "layernorm_bwd: C=%d exceeds %d", C, 32 * LN_MAX_CPL);
    if (C % vec_width(dtype) == 0 && C <= 128 * vec_width(dtype) && cnb_aligned16(x) && cnb_aligned16(dy) && cnb_aligned16(dx)) {
        const int K = cnb_div_up(C, 32 * vec_width(dtype));
        const dim3 grid(stream_grid(P, 8 * (K == 1 ? 2 : 1), 3));
#define CNB_LN_BWD(KK)                                                                                                               \
    CNB_DISPATCH_DTYPE(dtype, {                                                                                                      \
        CNB_LAUNCH((layernorm_bwd_vec_kernel<T, KK>), grid, dim3(256), 2 * C * sizeof(float), (cudaStream_t)stream, (const T*)x,     \
                   (const T*)dy, gamma, save_mean, save_rstd, (T*)dx, dgamma, dbeta, (long)P, C);                                    \
    })
        if (K == 1)
            CNB_LN_BWD(1);
        else if (K == 2)
            CNB_LN_BWD(2);
        else
            CNB_LN_BWD(4);
#undef CNB_LN_BWD
        CNB_CHECK_LAUNCH("layernorm_bwd_vec_kernel");
        return CNB_OK;
    }
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((layernorm_bwd_kernel<T>), dim3(stream_grid(P, 8, 2)), dim3(256), 0, (cudaStream_t)stream, (const T*)x, (const T*)dy, gamma,
                   save_mean, save_rstd, (T*)dx, dgamma, dbeta, (long)P, C);
    });
    CNB_CHECK_LAUNCH("layernorm_bwd_kernel");
    return CNB_OK;
}

// ------------------------------------------------------------------------------------------------
static int check_na(int B, int H, int W, int heads, int hd, int ksize, int dilation) {
    CNB_REQUIRE(B > 0 && H > 0 && W > 0 && heads > 0 && hd > 0, "na2d: bad shape");
    CNB_REQUIRE(ksize >= 1 && (ksize & 1) && ksize <= NA_MAX_K, "na2d: kernel_size=%d must be odd and <= %d", ksize, NA_MAX_K);
    CNB_REQUIRE(hd <= 32 * NA_MAX_DPL, "na2d: head_dim=%d exceeds %d", hd, 32 * NA_MAX_DPL);
    CNB_REQUIRE(dilation >= 1 && ksize * dilation <= H && ksize * dilation <= W,
                "na2d: kernel_size*dilation=%d exceeds the %dx%d input (natten raises here too)", ksize * dilation, H, W);
    return CNB_OK;
}

// geometry of the tiled kernels; returns false when the shape must take the warp-per-(pixel, head) kernels instead
static bool na_tile_setup(int B, int H, int W, int heads, int hd, int ksize, int dil, float scale, int dtype, bool with_stats, NaTile* g,
                          int* lph, size_t* smem) {
    const int V = vec_width(dtype);
    const size_t es = dtype == CNB_BF16 ? 2 : 4;
    if (hd % V != 0) return false;
    const int l = hd / V;
    if (l > 32 || (l & (l - 1)) != 0) return false;
    if (dil > NA_TH) return false;  // the halo bound needs TILE >= dilation
    const int span = (ksize - 1) * dil;
    g->B = B, g->H = H, g->W = W, g->heads = heads, g->hd = hd, g->ksize = ksize, g->dil = dil, g->scale = scale;
    g->groups = 1;
    g->RH = H < NA_TH + span ? H : NA_TH + span;
    g->RW = W < NA_TW + span ? W : NA_TW + span;
    g->tiles_y = cnb_div_up(H, NA_TH);
    g->tiles_x = cnb_div_up(W, NA_TW);
    *smem = (size_t)g->RH * g->RW * 2 * hd * es + (with_stats ? (size_t)g->RH * g->RW * 2 * sizeof(float) : 0);
    if (*smem > (size_t)NA_MAX_SMEM) return false;
    if ((long)B * g->tiles_y * g->tiles_x > 2147483647L || heads > 65535) return false;
    *lph = l;
    return true;
}

// Geometry of the specialised kernels (k_na_fast.cuh): dilation d = d*d independent dilation-1 sub-images of at most
// ceil(H/d) x ceil(W/d) pixels; tiles, halo (k/2 sub-image pixels) and the staged region are those of the largest sub-image.
static bool naf_tile_setup(int B, int H, int W, int heads, int hd, int ksize, int dil, float scale, NaTile* g, int* lph, size_t* smem,
                           dim3* grid) {
    if (hd % 8 != 0) return false;
    const int Hs = cnb_div_up(H, dil), Ws = cnb_div_up(W, dil);
    const int span = ksize - 1;
    g->B = B, g->H = H, g->W = W, g->heads = heads, g->hd = hd, g->ksize = ksize, g->dil = 1, g->scale = scale;
    g->groups = dil;
    g->RH = Hs < NA_TH + span ? Hs : NA_TH + span;
    g->RW = Ws < NA_TW + span ? Ws : NA_TW + span;
    g->tiles_y = cnb_div_up(Hs, NA_TH);
    g->tiles_x = cnb_div_up(Ws, NA_TW);
    *smem = (size_t)g->RH * g->RW * 2 * hd * 2;
    if (*smem > (size_t)NA_MAX_SMEM) return false;
    const long ctas = (long)B * dil * dil * g->tiles_y * g->tiles_x;
    if (ctas > 2147483647L || heads > 65535) return false;
    *lph = hd / 8;
    *grid = dim3((unsigned)ctas, heads);
    return true;
}

// which tiling the key-side backward pass of the specialised kernels uses (see k_na_fast.cuh); measured per shape on B200
static bool na_dkv_image_tiles(int ksize, int dilation) {
    static const int forced = [] {
        const char* e = getenv("CNB_NA_DKV");
        return !e ? 0 : (e[0] == 'i' ? 1 : (e[0] == 'g' ? 2 : 0));
    }();
    if (forced) return forced == 1;
    (void)ksize, (void)dilation;
    // B200: image tiles win the key-side pass at every shape measured (k3 d2 128^2: 1.20 vs 1.46 ms bwd; k7 d2 256^2: 12.9 vs 16.1;
    // k3 d1 64^2: 0.36 vs 0.38) -- the sub-image kernel's run-time strides cost it registers (k7: 254 vs 128)
    return true;
}

// launch a `template <typename T, int LPH>` tiled NA kernel
#define CNB_NA_LAUNCH_LPH(KERNEL, LPHV, ...)                                                                       \
    CNB_DISPATCH_DTYPE(dtype, {                                                                                    \
        CNB_SET_SMEM((KERNEL<T, LPHV>), smem);                                                                     \
        CNB_LAUNCH((KERNEL<T, LPHV>), grid, dim3(NA_TILE_THREADS), smem, (cudaStream_t)stream, __VA_ARGS__);       \
    })
#define CNB_NA_LAUNCH(KERNEL, ...)                                   \
    do {                                                             \
        switch (lph) {                                               \
            case 1: CNB_NA_LAUNCH_LPH(KERNEL, 1, __VA_ARGS__); break;   \
            case 2: CNB_NA_LAUNCH_LPH(KERNEL, 2, __VA_ARGS__); break;   \
            case 4: CNB_NA_LAUNCH_LPH(KERNEL, 4, __VA_ARGS__); break;   \
            case 8: CNB_NA_LAUNCH_LPH(KERNEL, 8, __VA_ARGS__); break;   \
            case 16: CNB_NA_LAUNCH_LPH(KERNEL, 16, __VA_ARGS__); break; \
            default: CNB_NA_LAUNCH_LPH(KERNEL, 32, __VA_ARGS__); break; \
        }                                                            \
    } while (0)

// launch a specialised NA kernel (naf::eligible shapes only).  Query-side kernels (`template <int KS, int DIL, int LPH, int VPL>`): head_dim
// 64 runs 4 lanes x 2 vectors, head_dim 32 runs 4 lanes x 1 vector.  Key-side kernels (`template <int KS, int DIL, int LPH>`): hd / 8 lanes.
#define CNB_NAQ_LAUNCH_K(KERNEL, KSV, ...)                                                                               \
    do {                                                                                                                 \
        if (hd == 64) {                                                                                                  \
            CNB_SET_SMEM((KERNEL<KSV, 1, 4, 2>), smem);                                                                  \
            CNB_LAUNCH((KERNEL<KSV, 1, 4, 2>), grid, dim3(NA_TILE_THREADS), smem, (cudaStream_t)stream, __VA_ARGS__);    \
        } else {                                                                                                         \
            CNB_SET_SMEM((KERNEL<KSV, 1, 4, 1>), smem);                                                                  \
            CNB_LAUNCH((KERNEL<KSV, 1, 4, 1>), grid, dim3(NA_TILE_THREADS), smem, (cudaStream_t)stream, __VA_ARGS__);    \
        }                                                                                                                \
    } while (0)
#define CNB_NAQ_LAUNCH(KERNEL, ...)                          \
    do {                                                     \
        if (ksize == 3)                                      \
            CNB_NAQ_LAUNCH_K(KERNEL, 3, __VA_ARGS__);        \
        else                                                 \
            CNB_NAQ_LAUNCH_K(KERNEL, 7, __VA_ARGS__);        \
    } while (0)
#define CNB_NAF_LAUNCH_KDL(KERNEL, KSV, DILV, LPHV, ...)                                                             \
    do {                                                                                                             \
        CNB_SET_SMEM((KERNEL<KSV, DILV, LPHV>), smem);                                                               \
        CNB_LAUNCH((KERNEL<KSV, DILV, LPHV>), grid, dim3(NA_TILE_THREADS), smem, (cudaStream_t)stream, __VA_ARGS__); \
    } while (0)
#define CNB_NAF_LAUNCH_KD(KERNEL, KSV, DILV, ...)                                   \
    do {                                                                            \
        if (lph == 8)                                                               \
            CNB_NAF_LAUNCH_KDL(KERNEL, KSV, DILV, 8, __VA_ARGS__);                  \
        else                                                                        \
            CNB_NAF_LAUNCH_KDL(KERNEL, KSV, DILV, 4, __VA_ARGS__);                  \
    } while (0)
#define CNB_NAF_LAUNCH(KERNEL, ...)                                                 \
    do {                                                                            \
        if (ksize == 3)                                                             \
            CNB_NAF_LAUNCH_KD(KERNEL, 3, 1, __VA_ARGS__);                           \
        else                                                                        \
            CNB_NAF_LAUNCH_KD(KERNEL, 7, 1, __VA_ARGS__);                           \
    } while (0)

int64_t cnb_na2d_bwd_workspace_floats(int B, int H, int W, int heads, int hd, int ksize, int dilation, int dtype) {
    if (check_na(B, H, W, heads, hd, ksize, dilation)) return 0;
    if (natc::usable(B, heads, hd, ksize, dtype)) return 0;  // the tensor-core backward recomputes: no probability records
    NaTile g;
    int lph;
    size_t smem;
    dim3 grid;
    if (!naf::eligible(hd, ksize, dilation, dtype) || !naf_tile_setup(B, H, W, heads, hd, ksize, dilation, 1.f, &g, &lph, &smem, &grid)) return 0;
    return (int64_t)B * H * W * heads * ksize * ksize;  // one bf16 pair (4 bytes) per (pixel, head, neighbour)
}

int cnb_na2d_tiled_eligible(int B, int H, int W, int heads, int hd, int ksize, int dilation, int dtype) {
    if (check_na(B, H, W, heads, hd, ksize, dilation)) return 0;
    if (natc::usable(B, heads, hd, ksize, dtype)) return 1;
    NaTile g;
    int lph;
    size_t smem;
    dim3 grid;
    if (naf::eligible(hd, ksize, dilation, dtype) && naf_tile_setup(B, H, W, heads, hd, ksize, dilation, 1.f, &g, &lph, &smem, &grid)) return 1;
    return na_tile_setup(B, H, W, heads, hd, ksize, dilation, 1.f, dtype, true, &g, &lph, &smem) ? 1 : 0;
}

int cnb_na2d_fwd(const void* qkv, void* out, float* lse, int B, int H, int W, int heads, int hd, int ksize, int dilation, float scale,
                 int dtype, void* stream) {
    int rc = check_na(B, H, W, heads, hd, ksize, dilation);
    if (rc) return rc;
    CNB_REQUIRE(qkv && out, "na2d_fwd: null pointer");
    if (lse && natc::usable(B, heads, hd, ksize, dtype)) return natc::launch_fwd(qkv, out, lse, B, H, W, heads, hd, ksize, dilation, scale, stream);
    NaTile g;
    int lph;
    size_t smem;
    if (lse && cnb_aligned16(qkv) && cnb_aligned16(out) && naf::eligible(hd, ksize, dilation, dtype)) {
        dim3 grid;
        if (naf_tile_setup(B, H, W, heads, hd, ksize, dilation, scale, &g, &lph, &smem, &grid)) {
            CNB_NAQ_LAUNCH(naf::na2d_fwd_fast_kernel, (const bf16_t*)qkv, (bf16_t*)out, lse, g);
            CNB_CHECK_LAUNCH("na2d_fwd_fast_kernel");
            return CNB_OK;
        }
    }
    if (lse && cnb_aligned16(qkv) && cnb_aligned16(out) &&
        na_tile_setup(B, H, W, heads, hd, ksize, dilation, scale, dtype, false, &g, &lph, &smem)) {
        const dim3 grid(B * g.tiles_y * g.tiles_x, heads);
        CNB_NA_LAUNCH(na2d_fwd_tile_kernel, (const T*)qkv, (T*)out, lse, g);
        CNB_CHECK_LAUNCH("na2d_fwd_tile_kernel");
        return CNB_OK;
    }
    const long items = (long)B * H * W * heads;
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((na2d_fwd_kernel<T>), dim3(stream_grid(items, 8)), dim3(256), 0, (cudaStream_t)stream, (const T*)qkv, (T*)out, B, H, W, heads,
                   hd, ksize, dilation, scale, NaDrop{nullptr, 0, 0u, 1.0f});
    });
    CNB_CHECK_LAUNCH("na2d_fwd_kernel");
    return CNB_OK;
}

static inline NaDrop na_drop(const void* rng_state, int site, float p) {
    long t = lroundf(p * 65536.0f);
    t = t < 0 ? 0 : (t > 65536 ? 65536 : t);
    return NaDrop{(const int64_t*)rng_state, site, (uint32_t)t, 1.0f / (1.0f - p)};
}

// training-mode attention dropout (attn_drop > 0): the one-warp-per-(pixel, head) kernels with a keep mask on the probabilities
int cnb_na2d_dropout_fwd(const void* qkv, void* out, int B, int H, int W, int heads, int hd, int ksize, int dilation, float scale,
                         const void* rng_state, int site, float p, int dtype, void* stream) {
    int rc = check_na(B, H, W, heads, hd, ksize, dilation);
    if (rc) return rc;
    CNB_REQUIRE(qkv && out && rng_state && p >= 0.f && p < 1.f, "na2d_dropout_fwd: bad arguments");
    const long items = (long)B * H * W * heads;
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((na2d_fwd_kernel<T>), dim3(stream_grid(items, 8)), dim3(256), 0, (cudaStream_t)stream, (const T*)qkv, (T*)out, B, H, W, heads,
                   hd, ksize, dilation, scale, na_drop(rng_state, site, p));
    });
    CNB_CHECK_LAUNCH("na2d_fwd_kernel");
    return CNB_OK;
}

int cnb_na2d_dropout_bwd(const void* qkv, const void* dout, float* dqkv_acc, void* dqkv, int B, int H, int W, int heads, int hd, int ksize,
                         int dilation, float scale, const void* rng_state, int site, float p, int dtype, void* stream) {
    int rc = check_na(B, H, W, heads, hd, ksize, dilation);
    if (rc) return rc;
    CNB_REQUIRE(qkv && dout && dqkv && dqkv_acc && rng_state && p >= 0.f && p < 1.f, "na2d_dropout_bwd: bad arguments");
    const long items = (long)B * H * W * heads;
    const long n = (long)B * H * W * 3 * heads * hd;
    CNB_MEMSET_ASYNC(dqkv_acc, 0, sizeof(float) * n, (cudaStream_t)stream);
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((na2d_bwd_kernel<T>), dim3(stream_grid(items, 8)), dim3(256), 0, (cudaStream_t)stream, (const T*)qkv, (const T*)dout,
                   dqkv_acc, B, H, W, heads, hd, ksize, dilation, scale, na_drop(rng_state, site, p));
        CNB_LAUNCH((cast_from_f32_kernel<T>), dim3(stream_grid(n)), dim3(256), 0, (cudaStream_t)stream, (const float*)dqkv_acc, (T*)dqkv, n);
    });
    CNB_CHECK_LAUNCH("na2d_bwd_kernel");
    return CNB_OK;
}

int cnb_na2d_bwd(const void* qkv, const void* dout, const void* out, const float* lse, float* dvec, float* dqkv_acc, float* pds_ws, void* dqkv,
                 int B, int H, int W, int heads, int hd, int ksize, int dilation, float scale, int dtype, void* stream) {
    int rc = check_na(B, H, W, heads, hd, ksize, dilation);
    if (rc) return rc;
    CNB_REQUIRE(qkv && dout && dqkv, "na2d_bwd: null pointer");
    if (out && lse && dvec && natc::usable(B, heads, hd, ksize, dtype))
        return natc::launch_bwd(qkv, dout, out, lse, dvec, dqkv, B, H, W, heads, hd, ksize, dilation, scale, stream);
    NaTile g;
    int lph;
    size_t smem;
    if (out && lse && pds_ws && cnb_aligned16(qkv) && cnb_aligned16(dout) && cnb_aligned16(out) && cnb_aligned16(dqkv) &&
        naf::eligible(hd, ksize, dilation, dtype)) {
        dim3 grid;
        if (naf_tile_setup(B, H, W, heads, hd, ksize, dilation, scale, &g, &lph, &smem, &grid)) {
            // the staged rows are k|v (query pass) or q|dout (key pass)
            CNB_NAQ_LAUNCH(naf::na2d_bwd_dq_fast_kernel, (const bf16_t*)qkv, (const bf16_t*)dout, (const bf16_t*)out, lse, (uint32_t*)pds_ws,
                           (bf16_t*)dqkv, g);
            // key side: sub-image tiles, or image-space tiles (the query records are indexed by image pixel: the two passes may tile
            // differently); CNB_NA_DKV=img|group overrides the default
            NaTile gi;
            int lph_i;
            size_t smem_i;
            const bool img_ok = na_tile_setup(B, H, W, heads, hd, ksize, dilation, scale, dtype, false, &gi, &lph_i, &smem_i);
            if (img_ok && na_dkv_image_tiles(ksize, dilation)) {
                const dim3 grid_g = grid;
                (void)grid_g;
                grid = dim3(B * gi.tiles_y * gi.tiles_x, heads);
                smem = smem_i;
// (two vectors per lane were measured for this pass too: k7 d2 256^2 backward 10.5 -> 15.3 ms -- the scattered 4-byte record loads
// are its latency-critical part and halving the lanes halves how many are in flight; it stays at one vector per lane)
#define CNB_NAK_IMG(KSV, DILV)                                                                                                          \
    do {                                                                                                                                \
        if (hd == 64) {                                                                                                                 \
            CNB_SET_SMEM((naf::na2d_bwd_dkv_img_kernel<KSV, DILV, 8, 1>), smem);                                                        \
            CNB_LAUNCH((naf::na2d_bwd_dkv_img_kernel<KSV, DILV, 8, 1>), grid, dim3(NA_TILE_THREADS), smem, (cudaStream_t)stream,        \
                       (const bf16_t*)qkv, (const bf16_t*)dout, (const uint32_t*)pds_ws, (bf16_t*)dqkv, gi);                            \
        } else {                                                                                                                        \
            CNB_SET_SMEM((naf::na2d_bwd_dkv_img_kernel<KSV, DILV, 4, 1>), smem);                                                        \
            CNB_LAUNCH((naf::na2d_bwd_dkv_img_kernel<KSV, DILV, 4, 1>), grid, dim3(NA_TILE_THREADS), smem, (cudaStream_t)stream,        \
                       (const bf16_t*)qkv, (const bf16_t*)dout, (const uint32_t*)pds_ws, (bf16_t*)dqkv, gi);                            \
        }                                                                                                                               \
    } while (0)
                if (dilation == 1) {
                    if (ksize == 3) CNB_NAK_IMG(3, 1);
                    else CNB_NAK_IMG(7, 1);
                } else {
                    if (ksize == 3) CNB_NAK_IMG(3, 2);
                    else CNB_NAK_IMG(7, 2);
                }
#undef CNB_NAK_IMG
            } else {
                CNB_NAF_LAUNCH(naf::na2d_bwd_dkv_fast_kernel, (const bf16_t*)qkv, (const bf16_t*)dout, (const uint32_t*)pds_ws, (bf16_t*)dqkv, g);
            }
            CNB_CHECK_LAUNCH("na2d_bwd_fast_kernels");
            return CNB_OK;
        }
    }
    if (out && lse && dvec && cnb_aligned16(qkv) && cnb_aligned16(dout) && cnb_aligned16(out) && cnb_aligned16(dqkv) &&
        na_tile_setup(B, H, W, heads, hd, ksize, dilation, scale, dtype, true, &g, &lph, &smem)) {
        const dim3 grid(B * g.tiles_y * g.tiles_x, heads);
        CNB_NA_LAUNCH(na2d_bwd_dq_tile_kernel, (const T*)qkv, (const T*)dout, (const T*)out, lse, dvec, (T*)dqkv, g);
        CNB_NA_LAUNCH(na2d_bwd_dkv_tile_kernel, (const T*)qkv, (const T*)dout, lse, (const float*)dvec, (T*)dqkv, g);
        CNB_CHECK_LAUNCH("na2d_bwd_tile_kernels");
        return CNB_OK;
    }
    CNB_REQUIRE(dqkv_acc, "na2d_bwd: this shape takes the scatter kernel and needs the fp32 accumulation workspace");
    const long items = (long)B * H * W * heads;
    const long n = (long)B * H * W * 3 * heads * hd;
    CNB_MEMSET_ASYNC(dqkv_acc, 0, sizeof(float) * n, (cudaStream_t)stream);
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((na2d_bwd_kernel<T>), dim3(stream_grid(items, 8)), dim3(256), 0, (cudaStream_t)stream, (const T*)qkv, (const T*)dout,
                   dqkv_acc, B, H, W, heads, hd, ksize, dilation, scale, NaDrop{nullptr, 0, 0u, 1.0f});
        CNB_LAUNCH((cast_from_f32_kernel<T>), dim3(stream_grid(n)), dim3(256), 0, (cudaStream_t)stream, (const float*)dqkv_acc, (T*)dqkv, n);
    });
    CNB_CHECK_LAUNCH("na2d_bwd_kernel");
    return CNB_OK;
}

// ------------------------------------------------------------------------------------------------
static inline float align_corners_scale(int in_len, int out_len) { return out_len > 1 ? (float)(in_len - 1) / (float)(out_len - 1) : 0.f; }

int cnb_resize_bilinear_fwd(const void* x, void* y, int B, int Hin, int Win, int Hout, int Wout, int C, int dtype, void* stream) {
    CNB_REQUIRE(x && y && B > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && C > 0, "resize_bilinear_fwd: bad arguments");
    const long total = (long)B * Hout * Wout * C;
    if (C % vec_width(dtype) == 0 && cnb_aligned16(x) && cnb_aligned16(y)) {
        CNB_DISPATCH_DTYPE(dtype, {
            CNB_LAUNCH((resize_bilinear_fwd_vec_kernel<T>), dim3(B * Hout), dim3(256), 0, (cudaStream_t)stream,
                       (const T*)x, (T*)y, B, Hin, Win, Hout, Wout, C, align_corners_scale(Hin, Hout), align_corners_scale(Win, Wout));
        });
        CNB_CHECK_LAUNCH("resize_bilinear_fwd_vec_kernel");
        return CNB_OK;
    }
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((resize_bilinear_fwd_kernel<T>), dim3(stream_grid(total)), dim3(256), 0, (cudaStream_t)stream, (const T*)x, (T*)y, B, Hin,
                   Win, Hout, Wout, C, align_corners_scale(Hin, Hout), align_corners_scale(Win, Wout));
    });
    CNB_CHECK_LAUNCH("resize_bilinear_fwd_kernel");
    return CNB_OK;
}

// rows per CTA of the persistent table kernel: the smallest whole number that keeps the grid within 8 CTAs per SM
static inline int resize_bwd_grid(long rows) {
    const long cap = 8L * CNB_NUM_SMS;
    const long per = (rows + cap - 1) / cap;
    return (int)((rows + per - 1) / per);
}

int cnb_resize_bilinear_bwd(const void* dy, void* dx, int B, int Hin, int Win, int Hout, int Wout, int C, int dtype, void* stream) {
    return cnb_resize_bilinear_bwd_colsum(dy, dx, B, Hin, Win, Hout, Wout, C, nullptr, 0, dtype, stream);
}

int cnb_resize_bilinear_bwd_colsum(const void* dy, void* dx, int B, int Hin, int Win, int Hout, int Wout, int C, float* colsum, int accumulate,
                                   int dtype, void* stream) {
    CNB_REQUIRE(dy && dx && B > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && C > 0, "resize_bilinear_bwd: bad arguments");
    const long total = (long)B * Hin * Win * C;
    const float rh_ = align_corners_scale(Hin, Hout), rw_ = align_corners_scale(Win, Wout);
    if (C % vec_width(dtype) == 0 && cnb_aligned16(dy) && cnb_aligned16(dx) && rh_ >= 0.5f && rw_ >= 0.5f && Win <= 4096 && Hin <= 4096) {
        // scale >= 0.5: at most RB_NC = 5 outputs read an input index per axis (support of length 2/scale <= 4); scale >= 0.7 (the
        // ConvTranspose fix-up, scale ~ 1): at most 3
        const bool near1 = rh_ >= 0.7f && rw_ >= 0.7f;
        const int nc = near1 ? RB_NC_NEAR1 : RB_NC;
        const int CV = C / vec_width(dtype);
        // the fused column sum needs a thread to stay on one channel vector (256 % CV == 0) and the sums next to the tables
        const bool fuse_sum = colsum && 256 % CV == 0 && C <= 4096;
        float* cs = fuse_sum ? colsum : nullptr;
        if (fuse_sum && !accumulate) CNB_MEMSET_ASYNC(colsum, 0, sizeof(float) * C, (cudaStream_t)stream);
        static const bool persistent = [] {  // CNB_RESIZE_BWD=persistent: the row-looping variant (A/B)
            const char* e = getenv("CNB_RESIZE_BWD");
            return e && e[0] == 'p';
        }();
        if (!persistent) {
            const size_t smem_row = (size_t)Win * (1 + nc) * 4 + (1 + nc) * 4 + (fuse_sum ? (size_t)C * 4 : 0);
#define CNB_RESIZE_ROW(NCV, CS)                                                                                                              \
    do {                                                                                                                                     \
        CNB_SET_SMEM((resize_bilinear_bwd_row_kernel<T, NCV, CS>), smem_row);                                                                \
        CNB_LAUNCH((resize_bilinear_bwd_row_kernel<T, NCV, CS>), dim3(B * Hin), dim3(256), smem_row, (cudaStream_t)stream, (const T*)dy,     \
                   (T*)dx, B, Hin, Win, Hout, Wout, C, rh_, rw_, cs);                                                                        \
    } while (0)
            CNB_DISPATCH_DTYPE(dtype, {
                if (near1) {
                    if (cs) CNB_RESIZE_ROW(RB_NC_NEAR1, true);
                    else CNB_RESIZE_ROW(RB_NC_NEAR1, false);
                } else {
                    if (cs) CNB_RESIZE_ROW(RB_NC, true);
                    else CNB_RESIZE_ROW(RB_NC, false);
                }
            });
#undef CNB_RESIZE_ROW
            CNB_CHECK_LAUNCH("resize_bilinear_bwd_row_kernel");
            if (colsum && !fuse_sum) return cnb_bias_grad(dx, C, (int64_t)B * Hin * Win, C, colsum, accumulate, dtype, stream);
            return CNB_OK;
        }
        const size_t smem = (size_t)(Win + Hin) * (1 + nc) * 4 + (fuse_sum ? (size_t)C * 4 : 0);
        const dim3 grid(resize_bwd_grid((long)B * Hin));
        CNB_DISPATCH_DTYPE(dtype, {
            if (near1) {
                CNB_SET_SMEM((resize_bilinear_bwd_tab_kernel<T, RB_NC_NEAR1>), smem);
                CNB_LAUNCH((resize_bilinear_bwd_tab_kernel<T, RB_NC_NEAR1>), grid, dim3(256), smem, (cudaStream_t)stream, (const T*)dy, (T*)dx, B,
                           Hin, Win, Hout, Wout, C, rh_, rw_, cs);
            } else {
                CNB_SET_SMEM((resize_bilinear_bwd_tab_kernel<T, RB_NC>), smem);
                CNB_LAUNCH((resize_bilinear_bwd_tab_kernel<T, RB_NC>), grid, dim3(256), smem, (cudaStream_t)stream, (const T*)dy, (T*)dx, B, Hin,
                           Win, Hout, Wout, C, rh_, rw_, cs);
            }
        });
        CNB_CHECK_LAUNCH("resize_bilinear_bwd_tab_kernel");
        if (colsum && !fuse_sum) return cnb_bias_grad(dx, C, (int64_t)B * Hin * Win, C, colsum, accumulate, dtype, stream);
        return CNB_OK;
    }
    if (C % vec_width(dtype) == 0 && cnb_aligned16(dy) && cnb_aligned16(dx)) {
        CNB_DISPATCH_DTYPE(dtype, {
            CNB_LAUNCH((resize_bilinear_bwd_vec_kernel<T>), dim3(B * Hin), dim3(256), 0, (cudaStream_t)stream,
                       (const T*)dy, (T*)dx, B, Hin, Win, Hout, Wout, C, align_corners_scale(Hin, Hout), align_corners_scale(Win, Wout));
        });
        CNB_CHECK_LAUNCH("resize_bilinear_bwd_vec_kernel");
        if (colsum) return cnb_bias_grad(dx, C, (int64_t)B * Hin * Win, C, colsum, accumulate, dtype, stream);
        return CNB_OK;
    }
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((resize_bilinear_bwd_kernel<T>), dim3(stream_grid(total)), dim3(256), 0, (cudaStream_t)stream, (const T*)dy, (T*)dx, B, Hin,
                   Win, Hout, Wout, C, align_corners_scale(Hin, Hout), align_corners_scale(Win, Wout));
    });
    CNB_CHECK_LAUNCH("resize_bilinear_bwd_kernel");
    if (colsum) return cnb_bias_grad(dx, C, (int64_t)B * Hin * Win, C, colsum, accumulate, dtype, stream);
    return CNB_OK;
}

// ------------------------------------------------------------------------------------------------
static size_t pretime_smem(int C, int T_, int k, int up, int dtype, bool wgrad) {
    const size_t es = dtype == CNB_BF16 ? 2 : 4;
    const size_t xs = (size_t)C * T_ * PT_PIX * sizeof(float);
    if (!wgrad) return xs + (size_t)PT_PIX * up * es + (size_t)C * C * k * sizeof(float) + 16;
    return xs + (size_t)PT_PIX * (up + 2) * es + (size_t)C * C * k * sizeof(float) + 16;
}

int cnb_pretime_conv_fwd(const float* x, const float* w1, void* u, int B, int C, int T_, int H, int W, int k, int u_pitch, int dtype,
                         void* stream) {
    CNB_REQUIRE(x && w1 && u && B > 0 && C > 0 && H > 0 && W > 0 && k > 0 && T_ >= k, "pretime_conv_fwd: bad arguments (T=%d, k=%d)", T_, k);
    CNB_REQUIRE(u_pitch >= C * (T_ - k + 1), "pretime_conv_fwd: pitch %d < C*T' = %d", u_pitch, C * (T_ - k + 1));
    const size_t smem = pretime_smem(C, T_, k, u_pitch, dtype, false);
    CNB_REQUIRE(smem <= 200 * 1024, "pretime_conv_fwd: C*T = %d does not fit the shared-memory tile", C * T_);
    const long P = (long)B * H * W;
    dim3 grid(cnb_clamp_grid(cnb_div_up(P, PT_PIX), (long)CNB_NUM_SMS * 4));
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_SET_SMEM((pretime_conv_fwd_kernel<T>), smem);
        CNB_LAUNCH((pretime_conv_fwd_kernel<T>), grid, dim3(PT_THREADS), smem, (cudaStream_t)stream, x, w1, (T*)u, B, C, T_, H, W, k, u_pitch);
    });
    CNB_CHECK_LAUNCH("pretime_conv_fwd_kernel");
    return CNB_OK;
}

int cnb_pretime_conv_wgrad(const float* x, const void* du, float* dw1, int B, int C, int T_, int H, int W, int k, int u_pitch, int dtype,
                           void* stream) {
    CNB_REQUIRE(x && du && dw1 && B > 0 && C > 0 && H > 0 && W > 0 && k > 0 && k <= PT_MAX_K && T_ >= k, "pretime_conv_wgrad: bad arguments");
    CNB_REQUIRE(u_pitch >= C * (T_ - k + 1), "pretime_conv_wgrad: pitch %d < C*T'", u_pitch);
    CNB_REQUIRE(C * C * k <= PT_WG_MAX_TRIPLES * (PT_THREADS / PT_PIX), "pretime_conv_wgrad: C*C*k = %d exceeds %d", C * C * k,
                PT_WG_MAX_TRIPLES * (PT_THREADS / PT_PIX));
    const size_t smem = pretime_smem(C, T_, k, u_pitch, dtype, true);
    CNB_REQUIRE(smem <= 200 * 1024, "pretime_conv_wgrad: C*T = %d does not fit the shared-memory tile", C * T_);
    const long P = (long)B * H * W;
    dim3 grid(cnb_clamp_grid(cnb_div_up(P, PT_PIX), (long)CNB_NUM_SMS * 2));
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_SET_SMEM((pretime_conv_wgrad_kernel<T>), smem);
        CNB_LAUNCH((pretime_conv_wgrad_kernel<T>), grid, dim3(PT_THREADS), smem, (cudaStream_t)stream, x, (const T*)du, dw1, B, C, T_, H, W, k,
                   u_pitch);
    });
    CNB_CHECK_LAUNCH("pretime_conv_wgrad_kernel");
    return CNB_OK;
}

int cnb_time_to_pixel_major(const float* x, void* xp, int B, int CT, int64_t HW, int pitch, int dtype, void* stream) {
    CNB_REQUIRE(x && xp && B > 0 && CT > 0 && HW > 0 && pitch >= CT, "time_to_pixel_major: bad arguments");
    CNB_REQUIRE(pitch % vec_width(dtype) == 0 && cnb_aligned16(xp), "time_to_pixel_major: the pixel pitch must be whole 16-byte vectors");
    const size_t smem = ((size_t)CT * TP_XPITCH + 3) / 4 * 4 * sizeof(float) + (size_t)TP_PIX * pitch * (dtype == CNB_BF16 ? 2 : 4);
    CNB_REQUIRE(smem <= 200 * 1024, "time_to_pixel_major: C*T = %d does not fit the shared-memory tile", CT);
    const long P = (long)B * HW;
    dim3 grid(cnb_clamp_grid(cnb_div_up(P, TP_PIX), (long)CNB_NUM_SMS * 8));
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_SET_SMEM((time_to_pixel_major_kernel<T>), smem);
        CNB_LAUNCH((time_to_pixel_major_kernel<T>), grid, dim3(256), smem, (cudaStream_t)stream, x, (T*)xp, B, CT, (long)HW, pitch);
    });
    CNB_CHECK_LAUNCH("time_to_pixel_major_kernel");
    return CNB_OK;
}

int cnb_toeplitz_expand(const float* w1, float* wt, int C, int T_, int k, int rows, void* stream) {
    CNB_REQUIRE(w1 && wt && C > 0 && k > 0 && T_ >= k && rows >= C * (T_ - k + 1), "toeplitz_expand: bad arguments");
    CNB_LAUNCH(toeplitz_expand_kernel, dim3(stream_grid((long)rows * C * T_)), dim3(256), 0, (cudaStream_t)stream, w1, wt, C, T_, k, rows);
    CNB_CHECK_LAUNCH("toeplitz_expand_kernel");
    return CNB_OK;
}

int cnb_toeplitz_fold(const float* dwt, float* dw1, int C, int T_, int k, void* stream) {
    CNB_REQUIRE(dwt && dw1 && C > 0 && k > 0 && T_ >= k, "toeplitz_fold: bad arguments");
    CNB_LAUNCH(toeplitz_fold_kernel, dim3(stream_grid((long)C * C * k)), dim3(256), 0, (cudaStream_t)stream, dwt, dw1, C, T_, k);
    CNB_CHECK_LAUNCH("toeplitz_fold_kernel");
    return CNB_OK;
}

int cnb_tap_shift_add(const void* t, void* out, int B, int H, int W, int N, int KH, int KW, int pad, int dil, int t_pitch, int out_pitch,
                      int dtype, void* stream) {
    CNB_REQUIRE(t && out && B > 0 && H > 0 && W > 0 && N > 0 && N <= 16 && KH > 0 && KW > 0 && dil > 0 && pad >= 0 && t_pitch >= KH * KW * N &&
                    out_pitch >= N,
                "tap_shift_add: bad arguments (N <= 16)");
    const long P = (long)B * H * W;
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((tap_shift_add_kernel<T>), dim3(stream_grid(P, 256, 16)), dim3(256), 0, (cudaStream_t)stream, (const T*)t, (T*)out, B, H, W, N,
                   KH, KW, pad, dil, t_pitch, out_pitch);
    });
    CNB_CHECK_LAUNCH("tap_shift_add_kernel");
    return CNB_OK;
}

int cnb_tap_shift_gather(const void* dout, void* dt, int B, int H, int W, int N, int KH, int KW, int pad, int dil, int t_pitch, int out_pitch,
                         int dtype, void* stream) {
    CNB_REQUIRE(dout && dt && B > 0 && H > 0 && W > 0 && N > 0 && KH > 0 && KW > 0 && dil > 0 && pad >= 0 && t_pitch >= KH * KW * N &&
                    out_pitch >= N,
                "tap_shift_gather: bad arguments");
    const long total = (long)B * H * W * t_pitch;
    if (N == 9 && KH == 3 && KW == 3 && t_pitch % vec_width(dtype) == 0 && cnb_aligned16(dt)) {  // the three stacked Psi-Net streams
        CNB_DISPATCH_DTYPE(dtype, {
            CNB_LAUNCH((tap_shift_gather_px_kernel<T, 9, 3>), dim3(stream_grid((long)B * H * W, 256, 16)), dim3(256), 0, (cudaStream_t)stream,
                       (const T*)dout, (T*)dt, B, H, W, pad, dil, t_pitch, out_pitch);
        });
        CNB_CHECK_LAUNCH("tap_shift_gather_px_kernel");
        return CNB_OK;
    }
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((tap_shift_gather_kernel<T>), dim3(stream_grid(total, 256, 16)), dim3(256), 0, (cudaStream_t)stream, (const T*)dout, (T*)dt,
                   B, H, W, N, KH, KW, pad, dil, t_pitch, out_pitch);
    });
    CNB_CHECK_LAUNCH("tap_shift_gather_kernel");
    return CNB_OK;
}

// ------------------------------------------------------------------------------------------------
int cnb_final_combine_fwd(const void* ha, const void* hb, const void* hc, const float* params, float smooth, int flags, float* distance,
                          float* edge, float* crop, int64_t P, int dtype, void* stream) {
    CNB_REQUIRE(ha && hb && hc && params && distance && edge && crop && P > 0, "final_combine_fwd: bad arguments");
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((final_combine_fwd_kernel<T>), dim3(stream_grid(P)), dim3(256), 0, (cudaStream_t)stream, (const T*)ha, (const T*)hb,
                   (const T*)hc, params, smooth, flags, distance, edge, crop, (long)P);
    });
    CNB_CHECK_LAUNCH("final_combine_fwd_kernel");
    return CNB_OK;
}

int cnb_final_combine_bwd(const void* ha, const void* hb, const void* hc, const float* params, float smooth, int flags, const float* d_distance,
                          const float* d_edge, const float* d_crop, void* dha, void* dhb, void* dhc, float* dparams, float* red_ws, int64_t P,
                          int dtype, void* stream) {
    CNB_REQUIRE(ha && hb && hc && params && d_distance && d_edge && d_crop && dha && dhb && dhc && dparams && red_ws && P > 0,
                "final_combine_bwd: bad arguments");
    CNB_MEMSET_ASYNC(red_ws, 0, sizeof(float) * 32, (cudaStream_t)stream);
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((final_combine_bwd_kernel<T>), dim3(stream_grid(P, 256, 2)), dim3(256), 0, (cudaStream_t)stream, (const T*)ha, (const T*)hb,
                   (const T*)hc, params, smooth, flags, d_distance, d_edge, d_crop, (T*)dha, (T*)dhb, (T*)dhc, red_ws, (long)P);
    });
    CNB_LAUNCH(final_combine_param_grad_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, params, (const float*)red_ws, smooth, flags, dparams);
    CNB_CHECK_LAUNCH("final_combine_bwd_kernel");
    return CNB_OK;
}

// ------------------------------------------------------------------------------------------------
static int check_terms(const cnb_tanimoto_term* terms, int nterms, int B, int64_t HW, int backward) {
    CNB_REQUIRE(terms && nterms >= 1 && nterms <= TN_MAX_TERMS, "tanimoto: nterms=%d (max %d)", nterms, TN_MAX_TERMS);
    CNB_REQUIRE(B > 0 && B <= 65535 && HW > 0, "tanimoto: bad batch geometry");
    for (int t = 0; t < nterms; ++t) {
        const cnb_tanimoto_term& tm = terms[t];
        CNB_REQUIRE(tm.pred && tm.target && tm.C > 0, "tanimoto: term %d has null operands", t);
        CNB_REQUIRE(tm.target_mode >= 0 && tm.target_mode <= 3 && tm.mask_mode >= 0 && tm.mask_mode <= 3, "tanimoto: term %d bad modes", t);
        CNB_REQUIRE(tm.mask_mode == 0 || tm.mask, "tanimoto: term %d needs a mask pointer", t);
        CNB_REQUIRE(tm.target_mode != 0 || tm.tgt_c == tm.C || tm.tgt_c == 1, "tanimoto: term %d target channels %d vs %d", t, tm.tgt_c, tm.C);
        CNB_REQUIRE(!backward || tm.dpred, "tanimoto: term %d needs dpred", t);
    }
    return CNB_OK;
}

int cnb_tanimoto_fwd(const cnb_tanimoto_term* terms, int nterms, int B, int64_t HW, float smooth, int depth, int variant, double* sums,
                     float* coef, float* loss, void* stream) {
    int rc = check_terms(terms, nterms, B, HW, 0);
    if (rc) return rc;
    CNB_REQUIRE(sums && coef && loss && depth >= 1 && depth <= 32, "tanimoto_fwd: bad arguments");
    CNB_REQUIRE(variant >= 0 && variant <= 2, "tanimoto_fwd: variant %d (0 complement, 1 dist, 2 combined)", variant);
    TanimotoTerms pack;
    memset(&pack, 0, sizeof(pack));
    int cmax = 1;
    for (int t = 0; t < nterms; ++t) {
        pack.t[t] = terms[t];
        if (terms[t].C > cmax) cmax = terms[t].C;
    }
    CNB_MEMSET_ASYNC(sums, 0, sizeof(double) * 4 * nterms * B, (cudaStream_t)stream);
    const void* shared_labels = nullptr;
    if (nterms > 1 && tanimoto_shared_labels(terms, nterms, (long)HW, &shared_labels)) {
        // every term in one pass over the pixels: the label tensor is read once (24 instead of 40 bytes per pixel for the TowerUNet loss)
        const int chunks = cnb_clamp_grid(cnb_div_up((long)HW, 256 * 4 * 2), cnb_div_up(4L * CNB_NUM_SMS, (long)B));
        CNB_LAUNCH(tanimoto_sums_fused_kernel, dim3(chunks, B), dim3(256), 0, (cudaStream_t)stream, pack, nterms, B, (long)HW,
                   (const long long*)shared_labels, sums);
    } else {
        const int chunks = cnb_clamp_grid(cnb_div_up((long)cmax * HW, 256 * 8), cnb_div_up(4L * CNB_NUM_SMS, (long)B * nterms));
        CNB_LAUNCH(tanimoto_sums_kernel, dim3(chunks, B, nterms), dim3(256), 0, (cudaStream_t)stream, pack, B, (long)HW, sums);
    }
    CNB_MEMSET_ASYNC(loss, 0, sizeof(float) * (1 + nterms), (cudaStream_t)stream);
    CNB_LAUNCH(tanimoto_finalize_kernel, dim3(cnb_div_up((long)nterms * B, 256)), dim3(256), 0, (cudaStream_t)stream, pack, nterms, B, (long)HW,
               smooth, depth, variant, (const double*)sums, coef, loss);
    CNB_CHECK_LAUNCH("tanimoto_fwd");
    return CNB_OK;
}

int cnb_tanimoto_bwd(const cnb_tanimoto_term* terms, int nterms, int B, int64_t HW, const float* coef, const float* gscale, void* stream) {
    int rc = check_terms(terms, nterms, B, HW, 1);
    if (rc) return rc;
    CNB_REQUIRE(coef, "tanimoto_bwd: null coef");
    TanimotoTerms pack;
    memset(&pack, 0, sizeof(pack));
    int cmax = 1;
    for (int t = 0; t < nterms; ++t) {
        pack.t[t] = terms[t];
        if (terms[t].C > cmax) cmax = terms[t].C;
    }
    const void* shared_labels = nullptr;
    if (nterms > 1 && tanimoto_shared_labels(terms, nterms, (long)HW, &shared_labels)) {
        const int chunks = cnb_clamp_grid(cnb_div_up((long)HW, 256 * 4), cnb_div_up(8L * CNB_NUM_SMS, (long)B));
        CNB_LAUNCH(tanimoto_bwd_fused_kernel, dim3(chunks, B), dim3(256), 0, (cudaStream_t)stream, pack, nterms, B, (long)HW,
                   (const long long*)shared_labels, coef, gscale);
    } else {
        const int chunks = cnb_clamp_grid(cnb_div_up((long)cmax * HW, 256 * 4), cnb_div_up(8L * CNB_NUM_SMS, (long)B * nterms));
        CNB_LAUNCH(tanimoto_bwd_kernel, dim3(chunks, B, nterms), dim3(256), 0, (cudaStream_t)stream, pack, B, (long)HW, coef, gscale);
    }
    CNB_CHECK_LAUNCH("tanimoto_bwd_kernel");
    return CNB_OK;
}

int cnb_val_counts(const float* dist, const float* edge, const float* crop, const int64_t* y, const float* bdist, int64_t n, int edge_class,
                   float thresh, double* out, void* stream) {
    CNB_REQUIRE(dist && edge && crop && y && bdist && out && n > 0, "val_counts: bad arguments");
    CNB_MEMSET_ASYNC(out, 0, sizeof(double) * 12, (cudaStream_t)stream);
    CNB_LAUNCH(val_counts_kernel, dim3(stream_grid(n, 1024, 4)), dim3(256), 0, (cudaStream_t)stream, dist, edge, crop, (const long long*)y, bdist,
               (long)n, edge_class, thresh, out);
    CNB_CHECK_LAUNCH("val_counts_kernel");
    return CNB_OK;
}

// ------------------------------------------------------------------------------------------------
int cnb_grad_sqnorm(const float* g, int64_t n, float* norm_ws, void* stream) {
    CNB_REQUIRE(g && norm_ws && n > 0, "grad_sqnorm: bad arguments");
    CNB_MEMSET_ASYNC(norm_ws, 0, sizeof(float), (cudaStream_t)stream);
    CNB_LAUNCH(grad_sqnorm_kernel, dim3(stream_grid(n, 1024, 4)), dim3(256), 0, (cudaStream_t)stream, g, (long)n, norm_ws);
    CNB_CHECK_LAUNCH("grad_sqnorm_kernel");
    return CNB_OK;
}

int cnb_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float beta1, float beta2, float eps,
                   float weight_decay, float grad_scale, float clip_norm, const float* norm_ws, void* stream) {
    CNB_REQUIRE(p && g && m && v && hyper && n > 0, "adamw_step: bad arguments");
    CNB_REQUIRE(clip_norm <= 0.f || norm_ws, "adamw_step: clipping needs norm_ws");
    CNB_LAUNCH(adamw_kernel, dim3(stream_grid(n, 1024, 4)), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, (long)n, hyper, beta1, beta2, eps,
               weight_decay, grad_scale, clip_norm, norm_ws);
    CNB_CHECK_LAUNCH("adamw_kernel");
    return CNB_OK;
}

// ------------------------------------------------------------------------------------------------
// optional block variants: adaptive max pooling, spatial-channel attention, stand-alone SiLU, dropout (k_pool_attn.cuh)
// ------------------------------------------------------------------------------------------------
// VEC = whole 16-byte vectors when the channel count and the pointers allow it, else 1
#define CNB_DISPATCH_VEC(dtype, vec_ok, ...)                    \
    do {                                                        \
        if ((dtype) == CNB_F32) {                               \
            typedef float T;                                    \
            if (vec_ok) {                                       \
                constexpr int VEC = 4;                          \
                __VA_ARGS__                                     \
            } else {                                            \
                constexpr int VEC = 1;                          \
                __VA_ARGS__                                     \
            }                                                   \
        } else if ((dtype) == CNB_BF16) {                       \
            typedef bf16_t T;                                   \
            if (vec_ok) {                                       \
                constexpr int VEC = 8;                          \
                __VA_ARGS__                                     \
            } else {                                            \
                constexpr int VEC = 1;                          \
                __VA_ARGS__                                     \
            }                                                   \
        } else {                                                \
            CNB_FAIL(CNB_ERR_INVALID, "unsupported dtype %d", (int)(dtype)); \
        }                                                       \
    } while (0)

int cnb_adaptive_maxpool_fwd(const void* x, void* y, void* idx, int B, int Hin, int Win, int Hout, int Wout, int C, int dtype, void* stream) {
    CNB_REQUIRE(x && y && idx && B > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && C > 0, "adaptive_maxpool_fwd: bad arguments");
    CNB_REQUIRE(Hout <= Hin && Wout <= Win && cnb_div_up(Hin, Hout) + 1 <= 16 && cnb_div_up(Win, Wout) + 1 <= 16,
                "adaptive_maxpool_fwd: windows larger than 16 x 16 are not supported");
    const bool vec_ok = C % vec_width(dtype) == 0 && cnb_aligned16(x) && cnb_aligned16(y);
    CNB_DISPATCH_VEC(dtype, vec_ok, {
        const long total = (long)B * Hout * Wout * (C / VEC);
        CNB_LAUNCH((adaptive_maxpool_fwd_kernel<T, VEC>), dim3(stream_grid(total)), dim3(256), 0, (cudaStream_t)stream, (const T*)x, (T*)y,
                   (uint8_t*)idx, B, Hin, Win, Hout, Wout, C);
    });
    CNB_CHECK_LAUNCH("adaptive_maxpool_fwd_kernel");
    return CNB_OK;
}

int cnb_adaptive_maxpool_bwd(const void* dy, const void* idx, void* dx, int B, int Hin, int Win, int Hout, int Wout, int C, int dtype,
                             void* stream) {
    CNB_REQUIRE(dy && dx && idx && B > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && C > 0, "adaptive_maxpool_bwd: bad arguments");
    CNB_REQUIRE(Hout <= Hin && Wout <= Win, "adaptive_maxpool_bwd: the output must not be larger than the input");
    const bool vec_ok = C % vec_width(dtype) == 0 && cnb_aligned16(dy) && cnb_aligned16(dx);
    CNB_DISPATCH_VEC(dtype, vec_ok, {
        const long total = (long)B * Hin * Win * (C / VEC);
        CNB_LAUNCH((adaptive_maxpool_bwd_kernel<T, VEC>), dim3(stream_grid(total)), dim3(256), 0, (cudaStream_t)stream, (const T*)dy,
                   (const uint8_t*)idx, (T*)dx, B, Hin, Win, Hout, Wout, C);
    });
    CNB_CHECK_LAUNCH("adaptive_maxpool_bwd_kernel");
    return CNB_OK;
}

static bool act_code_ok(int act) { return act >= CNB_ACT_NONE && act <= CNB_ACT_HARDSWISH; }

int cnb_act_fwd(const void* x, void* y, int64_t n, int act, int dtype, void* stream) {
    CNB_REQUIRE(x && y && n > 0 && act_code_ok(act), "act_fwd: bad arguments");
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((act_fwd_kernel<T>), dim3(stream_grid(n)), dim3(256), 0, (cudaStream_t)stream, (const T*)x, (T*)y, (long)n, act);
    });
    CNB_CHECK_LAUNCH("act_fwd_kernel");
    return CNB_OK;
}

int cnb_silu_fwd(const void* x, void* y, int64_t n, int dtype, void* stream) { return cnb_act_fwd(x, y, n, CNB_ACT_SILU, dtype, stream); }

int cnb_act_bwd(const void* x, const void* dy, void* dx, int64_t n, int act, int dtype, void* stream) {
    CNB_REQUIRE(x && dy && dx && n > 0 && act_code_ok(act), "act_bwd: bad arguments");
    CNB_DISPATCH_DTYPE(dtype, {
        CNB_LAUNCH((act_bwd_kernel<T>), dim3(stream_grid(n)), dim3(256), 0, (cudaStream_t)stream, (const T*)x, (const T*)dy, (T*)dx, (long)n, act);
    });
    CNB_CHECK_LAUNCH("act_bwd_kernel");
    return CNB_OK;
}

int cnb_silu_bwd(const void* x, const void* dy, void* dx, int64_t n, int dtype, void* stream) {
    return cnb_act_bwd(x, dy, dx, n, CNB_ACT_SILU, dtype, stream);
}

// column tiles and pixel slices of the (slice, sample, column tile) grids of the attention pooling / apply-backward kernels
static inline void sca_grid(int B, int HW, int C, int dtype, bool vec_ok, int* ztiles, int* slices) {
    const int CV = C / (vec_ok ? vec_width(dtype) : 1);
    const int cols = CV < SCA_THREADS ? CV : SCA_THREADS;
    const int R = SCA_THREADS / cols;
    *ztiles = cnb_div_up(CV, cols);
    const long want = cnb_div_up(4L * CNB_NUM_SMS, (long)B * *ztiles);  // ~4 waves of CTAs
    const long most = cnb_div_up(HW, 4L * R);                           // at least 4 pixels per thread
    *slices = cnb_clamp_grid(want, most);
}

int cnb_sca_slices(int B, int HW, int C, int dtype) {
    if (B <= 0 || HW <= 0 || C <= 0) return 0;
    int zt, S;
    sca_grid(B, HW, C, dtype, C % vec_width(dtype) == 0, &zt, &S);
    int S1;
    sca_grid(B, HW, C, dtype, false, &zt, &S1);  // the scalar path (unaligned pointers) must fit the same workspace
    return S > S1 ? S : S1;
}

int cnb_sca_pool_fwd(const void* x, float* sp, float* ties, float* ch_avg, float* ch_max, int32_t* ch_arg, float* ws_sum, float* ws_max,
                     int32_t* ws_arg, int B, int HW, int C, int dtype, void* stream) {
    CNB_REQUIRE(x && sp && ties && ch_avg && ch_max && ch_arg && ws_sum && ws_max && ws_arg && B > 0 && HW > 0 && C > 0,
                "sca_pool_fwd: bad arguments");
    const bool vec_ok = C % vec_width(dtype) == 0 && cnb_aligned16(x);
    int zt, S;
    sca_grid(B, HW, C, dtype, vec_ok, &zt, &S);
    const long P = (long)B * HW;
    CNB_DISPATCH_VEC(dtype, vec_ok, {
        CNB_LAUNCH((sca_spatial_pool_kernel<T, VEC>), dim3(stream_grid(P * 32)), dim3(256), 0, (cudaStream_t)stream, (const T*)x, sp, ties, P, C);
        CNB_LAUNCH((sca_channel_pool_partial_kernel<T, VEC>), dim3(S, B, zt), dim3(SCA_THREADS), 0, (cudaStream_t)stream, (const T*)x, ws_sum,
                   ws_max, (int*)ws_arg, HW, C, S);
    });
    CNB_CHECK_LAUNCH("sca_pool kernels");
    CNB_LAUNCH(sca_channel_pool_final_kernel, dim3(stream_grid((long)B * C)), dim3(256), 0, (cudaStream_t)stream, (const float*)ws_sum,
               (const float*)ws_max, (const int*)ws_arg, ch_avg, ch_max, (int*)ch_arg, B, C, S, HW);
    CNB_CHECK_LAUNCH("sca_channel_pool_final_kernel");
    return CNB_OK;
}

int cnb_sca_pool_bwd(const void* x, const float* sp, const float* ties, const float* dsp, const float* dch_avg, const float* dch_max,
                     const int32_t* ch_arg, void* dx, int B, int HW, int C, int dtype, void* stream) {
    CNB_REQUIRE(x && sp && ties && dsp && dch_avg && dch_max && ch_arg && dx && B > 0 && HW > 0 && C > 0, "sca_pool_bwd: bad arguments");
    const bool vec_ok = C % vec_width(dtype) == 0 && cnb_aligned16(x) && cnb_aligned16(dx);
    CNB_DISPATCH_VEC(dtype, vec_ok, {
        const long total = (long)B * HW * (C / VEC);
        CNB_LAUNCH((sca_pool_bwd_kernel<T, VEC>), dim3(stream_grid(total)), dim3(256), 0, (cudaStream_t)stream, (const T*)x, sp, ties, dsp,
                   dch_avg, dch_max, (const int*)ch_arg, (T*)dx, B, HW, C);
    });
    CNB_CHECK_LAUNCH("sca_pool_bwd_kernel");
    return CNB_OK;
}

int cnb_sca_apply_fwd(const void* y, const float* cl, const float* sl, const float* gamma, void* out, int B, int HW, int C, int dtype,
                      void* stream) {
    CNB_REQUIRE(y && cl && sl && gamma && out && B > 0 && HW > 0 && C > 0, "sca_apply_fwd: bad arguments");
    const bool vec_ok = C % vec_width(dtype) == 0 && cnb_aligned16(y) && cnb_aligned16(out);
    CNB_DISPATCH_VEC(dtype, vec_ok, {
        const long total = (long)B * HW * (C / VEC);
        CNB_LAUNCH((sca_apply_fwd_kernel<T, VEC>), dim3(stream_grid(total)), dim3(256), 0, (cudaStream_t)stream, (const T*)y, cl, sl, gamma,
                   (T*)out, B, HW, C);
    });
    CNB_CHECK_LAUNCH("sca_apply_fwd_kernel");
    return CNB_OK;
}

int cnb_sca_apply_bwd(const void* y, const void* dout, const float* cl, const float* sl, const float* gamma, void* dy, float* dcl,
                      float* dsl, float* dgamma, int B, int HW, int C, int dtype, void* stream) {
    CNB_REQUIRE(y && dout && cl && sl && gamma && dy && dcl && dsl && dgamma && B > 0 && HW > 0 && C > 0, "sca_apply_bwd: bad arguments");
    const bool vec_ok = C % vec_width(dtype) == 0 && cnb_aligned16(y) && cnb_aligned16(dout) && cnb_aligned16(dy);
    int zt, S;
    sca_grid(B, HW, C, dtype, vec_ok, &zt, &S);
    CNB_MEMSET_ASYNC(dcl, 0, sizeof(float) * (size_t)B * C, (cudaStream_t)stream);
    CNB_MEMSET_ASYNC(dsl, 0, sizeof(float) * (size_t)B * HW, (cudaStream_t)stream);
    CNB_MEMSET_ASYNC(dgamma, 0, sizeof(float), (cudaStream_t)stream);
    CNB_DISPATCH_VEC(dtype, vec_ok, {
        CNB_LAUNCH((sca_apply_bwd_kernel<T, VEC>), dim3(S, B, zt), dim3(SCA_THREADS), 0, (cudaStream_t)stream, (const T*)y, (const T*)dout, cl,
                   sl, gamma, (T*)dy, dcl, dsl, dgamma, HW, C, S);
    });
    CNB_CHECK_LAUNCH("sca_apply_bwd_kernel");
    return CNB_OK;
}

int cnb_rng_advance(void* rng_state, void* stream) {
    CNB_REQUIRE(rng_state, "rng_advance: null state");
    CNB_LAUNCH(rng_advance_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, (int64_t*)rng_state);
    CNB_CHECK_LAUNCH("rng_advance_kernel");
    return CNB_OK;
}

static inline uint32_t dropout_threshold(float p) {
    long t = lroundf(p * 65536.0f);
    return (uint32_t)(t < 0 ? 0 : (t > 65536 ? 65536 : t));
}

int cnb_dropout(const void* x, void* out, int64_t n, const void* rng_state, int site, float p, int dtype, void* stream) {
    CNB_REQUIRE(x && out && rng_state && n > 0 && p >= 0.f && p < 1.f, "dropout: bad arguments");
    const bool vec_ok = n % vec_width(dtype) == 0 && cnb_aligned16(x) && cnb_aligned16(out);
    CNB_DISPATCH_VEC(dtype, vec_ok, {
        const long n_v = n / VEC;
        CNB_LAUNCH((dropout_kernel<T, VEC>), dim3(stream_grid(n_v)), dim3(256), 0, (cudaStream_t)stream, (const T*)x, (T*)out, n_v,
                   (const int64_t*)rng_state, site, dropout_threshold(p), 1.0f / (1.0f - p));
    });
    CNB_CHECK_LAUNCH("dropout_kernel");
    return CNB_OK;
}

int cnb_dropout2d(const void* x, void* out, int B, int HW, int C, const void* rng_state, int site, float p, int dtype, void* stream) {
    CNB_REQUIRE(x && out && rng_state && B > 0 && HW > 0 && C > 0 && p >= 0.f && p < 1.f, "dropout2d: bad arguments");
    const bool vec_ok = C % vec_width(dtype) == 0 && cnb_aligned16(x) && cnb_aligned16(out);
    CNB_DISPATCH_VEC(dtype, vec_ok, {
        const long total = (long)B * HW * (C / VEC);
        CNB_LAUNCH((dropout2d_kernel<T, VEC>), dim3(stream_grid(total)), dim3(256), 0, (cudaStream_t)stream, (const T*)x, (T*)out, B, HW, C,
                   (const int64_t*)rng_state, site, dropout_threshold(p), 1.0f / (1.0f - p));
    });
    CNB_CHECK_LAUNCH("dropout2d_kernel");
    return CNB_OK;
}

int cnb_window_load(const int16_t* tile, int T, int C, int Ht, int Wt, const int32_t* win, int win_stride, int B, int window_size, int pad,
                    float scale, float lo, float hi, const float* mean, const float* stdv, float* out, void* stream) {
    CNB_REQUIRE(tile && win && out && win_stride >= 2 && T > 0 && C > 0 && Ht > 0 && Wt > 0 && B > 0 && window_size > 0 && pad >= 0 && scale != 0.f,
                "window_load: bad arguments");
    const int Hw = window_size + 2 * pad;
    CNB_REQUIRE(Hw % 4 == 0, "window_load: window_size + 2 * padding must be a multiple of 4 (16-byte stores)");
    CNB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "window_load: out must be 16-byte aligned");
    CNB_REQUIRE((long)B * C * T < (1L << 31), "window_load: too many planes for one launch");
    // a quad of four int16 is one aligned 8-byte load when rows and the base pointer are; the kernel adds the per-window column test
    const int vec_ok = Wt % 4 == 0 && (reinterpret_cast<uintptr_t>(tile) & 7) == 0;
    CNB_LAUNCH(window_load_kernel, dim3((unsigned)((long)B * C * T)), dim3(256), 0, (cudaStream_t)stream, tile, T, C, Ht, Wt, win, win_stride, B,
               Hw, Hw, pad, scale, lo, hi, mean, stdv, out, vec_ok);
    CNB_CHECK_LAUNCH("window_load_kernel");
    return CNB_OK;
}

int cnb_predict_pack(const float* dist, const float* edge, const float* crop, int64_t batch_stride, int Hs, int Ws, int pad,
                     const int32_t* win, int B, int window_size, float scale, uint16_t* mosaic, int Ht, int Wt, int mosaic_pitch, void* stream) {
    CNB_REQUIRE(dist && edge && crop && win && mosaic && B > 0 && Hs > 0 && Ws > 0 && pad >= 0 && window_size > 0 && Ht > 0 && Wt > 0 &&
                    mosaic_pitch >= Wt,
                "predict_pack: bad arguments");
    CNB_REQUIRE(batch_stride >= (int64_t)Hs * Ws, "predict_pack: batch stride smaller than one prediction");
    CNB_LAUNCH(predict_pack_kernel, dim3((unsigned)((long)B * 3 * window_size)), dim3(128), 0, (cudaStream_t)stream, dist, edge, crop, (long)batch_stride, Hs, Ws,
               pad, win, B, window_size, scale, mosaic, Ht, Wt, mosaic_pitch);
    CNB_CHECK_LAUNCH("predict_pack_kernel");
    return CNB_OK;
}

}  // extern "C"
