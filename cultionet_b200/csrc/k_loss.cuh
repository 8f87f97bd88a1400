// Tanimoto-with-complement loss and its gradient as warp-shuffle reductions, plus the flat-buffer
// optimiser kernels.  Closed form: per (term, sample) accumulate
//   P = sum t'p', S = sum t'^2 + p'^2, St = sum t', Sp = sum p'   (t' = t*mask, p' = p*mask, N = C*H*W elements)
//   P' = N - St - Sp + P,  S' = 2N - 2St - 2Sp + S                 (the complement pair (1-t', 1-p'))
//   T(P,S) = (P+eps) * (1/D) * sum_d 1/(2^d S - (2^(d+1)-1) P + eps)
//   loss = 0.5*((1-T(P,S)) + (1-T(P',S'))), averaged over the batch.
#pragma once
#include "cnb_common.cuh"

namespace cnb {

constexpr int TN_MAX_TERMS = 4;
struct TanimotoTerms {
    cnb_tanimoto_term t[TN_MAX_TERMS];
};

__device__ __forceinline__ void tanimoto_elem(const cnb_tanimoto_term& tm, long b, long e, long HW, float& tq, float& pq, float& mk) {
    const long c = e / HW, hw = e - c * HW;
    const float p = tm.pred[(b * tm.C + c) * HW + hw];
    float t;
    if (tm.target_mode == 0) {
        const float* tp = reinterpret_cast<const float*>(tm.target);
        t = (tm.tgt_c == tm.C) ? tp[(b * tm.C + c) * HW + hw] : tp[b * HW + hw];
    } else {
        const long long lab = reinterpret_cast<const long long*>(tm.target)[b * HW + hw];
        if (tm.target_mode == 1)
            t = (lab == c) ? 1.f : 0.f;
        else if (tm.target_mode == 2)
            t = (lab == tm.edge_class) ? 1.f : 0.f;
        else
            t = (lab > 0 && lab < tm.edge_class) ? 1.f : 0.f;
    }
    float m = 1.f;
    if (tm.mask_mode == 1)
        m = reinterpret_cast<const float*>(tm.mask)[b * HW + hw];
    else if (tm.mask_mode == 2)
        m = (float)reinterpret_cast<const long long*>(tm.mask)[b * HW + hw];
    else if (tm.mask_mode == 3)
        m = (reinterpret_cast<const long long*>(tm.mask)[b * HW + hw] != -1) ? 1.f : 0.f;
    tq = t * m;
    pq = p * m;
    mk = m;
}

// Four consecutive pixels of a single-channel term (C == 1, HW % 4 == 0, 16-byte aligned operands: every term of the TowerUNet loss):
// one 16-byte load per fp32 operand and two per int64 label quad instead of 4 x (a 64-bit division + scalar loads).  The scalar
// version ran at 1.8 TB/s at a scaled size (B = 512), issue-bound (ncu: 59 % issue slots busy for 24 bytes per pixel).
__host__ __device__ __forceinline__ bool tanimoto_vec_ok(const cnb_tanimoto_term& tm, long HW) {
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    return tm.C == 1 && (HW & 3) == 0 && al(tm.pred) && al(tm.target) && (tm.mask_mode == 0 || al(tm.mask)) &&
           (tm.dpred == nullptr || al(tm.dpred)) && (tm.target_mode != 0 || tm.tgt_c == 1);
}
__device__ __forceinline__ void tanimoto_labels4(const void* base, long i, long long (&lab)[4]) {
    const longlong2* q = reinterpret_cast<const longlong2*>(reinterpret_cast<const long long*>(base) + i);
    const longlong2 u = q[0], v = q[1];
    lab[0] = u.x, lab[1] = u.y, lab[2] = v.x, lab[3] = v.y;
}
__device__ __forceinline__ void tanimoto_elem4(const cnb_tanimoto_term& tm, long b, long hw, long HW, float (&tq)[4], float (&pq)[4],
                                               float (&mk)[4]) {
    const long i = b * HW + hw;
    const float4 p4 = *reinterpret_cast<const float4*>(tm.pred + i);
    const float p[4] = {p4.x, p4.y, p4.z, p4.w};
    float t[4];
    long long lab[4];
    bool have_lab = false;
    if (tm.target_mode == 0) {
        const float4 t4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(tm.target) + i);
        t[0] = t4.x, t[1] = t4.y, t[2] = t4.z, t[3] = t4.w;
    } else {
        tanimoto_labels4(tm.target, i, lab);
        have_lab = true;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            t[j] = tm.target_mode == 1 ? (lab[j] == 0 ? 1.f : 0.f)
                                       : (tm.target_mode == 2 ? (lab[j] == tm.edge_class ? 1.f : 0.f) : ((lab[j] > 0 && lab[j] < tm.edge_class) ? 1.f : 0.f));
    }
    float m[4] = {1.f, 1.f, 1.f, 1.f};
    if (tm.mask_mode == 1) {
        const float4 m4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(tm.mask) + i);
        m[0] = m4.x, m[1] = m4.y, m[2] = m4.z, m[3] = m4.w;
    } else if (tm.mask_mode >= 2) {
        long long ml[4];
        if (have_lab && tm.mask == tm.target) {  // the labels double as the mask (get_true_labels): already in registers
#pragma unroll
            for (int j = 0; j < 4; ++j) ml[j] = lab[j];
        } else {
            tanimoto_labels4(tm.mask, i, ml);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = tm.mask_mode == 2 ? (float)ml[j] : (ml[j] != -1 ? 1.f : 0.f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) tq[j] = t[j] * m[j], pq[j] = p[j] * m[j], mk[j] = m[j];
}

// ---------------------------------------------------------------------------------------------
// All terms in ONE pass (the TowerUNet loss: three single-channel terms whose label-typed operands -- the targets of the edge / crop
// terms and every term's mask -- are the SAME int64 label tensor).  The per-term kernels read 40 bytes per pixel (the 8-byte labels
// three times over: L2 hits at 4 MB, DRAM traffic at a scaled size) for 24 algorithmic ones; here the labels of a pixel quad are loaded
// once and serve every term.
// ---------------------------------------------------------------------------------------------
// the one label tensor the terms share (NULL when no term has a label-typed operand), or false when they do not share one
static inline bool tanimoto_shared_labels(const cnb_tanimoto_term* t, int nterms, long HW, const void** labels) {
    const void* lab = nullptr;
    for (int i = 0; i < nterms; ++i) {
        if (!tanimoto_vec_ok(t[i], HW)) return false;
        const void* cand[2] = {t[i].target_mode != 0 ? (const void*)t[i].target : nullptr, t[i].mask_mode >= 2 ? (const void*)t[i].mask : nullptr};
        for (int k = 0; k < 2; ++k) {
            if (!cand[k]) continue;
            if (lab && cand[k] != lab) return false;
            lab = cand[k];
        }
    }
    *labels = lab;
    return true;
}
// tanimoto_elem4 with the label quad already in registers
__device__ __forceinline__ void tanimoto_elem4_shared(const cnb_tanimoto_term& tm, long i, const long long (&lab)[4], float (&tq)[4],
                                                      float (&pq)[4], float (&mk)[4]) {
    const float4 p4 = *reinterpret_cast<const float4*>(tm.pred + i);
    const float p[4] = {p4.x, p4.y, p4.z, p4.w};
    float t[4];
    if (tm.target_mode == 0) {
        const float4 t4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(tm.target) + i);
        t[0] = t4.x, t[1] = t4.y, t[2] = t4.z, t[3] = t4.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            t[j] = tm.target_mode == 1 ? (lab[j] == 0 ? 1.f : 0.f)
                                       : (tm.target_mode == 2 ? (lab[j] == tm.edge_class ? 1.f : 0.f) : ((lab[j] > 0 && lab[j] < tm.edge_class) ? 1.f : 0.f));
    }
    float m[4] = {1.f, 1.f, 1.f, 1.f};
    if (tm.mask_mode == 1) {
        const float4 m4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(tm.mask) + i);
        m[0] = m4.x, m[1] = m4.y, m[2] = m4.z, m[3] = m4.w;
    } else if (tm.mask_mode >= 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = tm.mask_mode == 2 ? (float)lab[j] : (lab[j] != -1 ? 1.f : 0.f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) tq[j] = t[j] * m[j], pq[j] = p[j] * m[j], mk[j] = m[j];
}

// grid = (chunks, B); the same sums as tanimoto_sums_kernel for every term
__global__ void __launch_bounds__(256) tanimoto_sums_fused_kernel(TanimotoTerms terms, int nterms, int B, long HW, const long long* __restrict__ labels,
                                                                  double* __restrict__ sums) {
    CNB_PDL_SYNC();
    __shared__ double sh[TN_MAX_TERMS * 4];
    if (threadIdx.x < TN_MAX_TERMS * 4) sh[threadIdx.x] = 0.0;
    __syncthreads();
    const long b = blockIdx.y;
    float a[TN_MAX_TERMS][4];
#pragma unroll
    for (int t = 0; t < TN_MAX_TERMS; ++t)
#pragma unroll
        for (int k = 0; k < 4; ++k) a[t][k] = 0.f;
    for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < HW; q += (long)gridDim.x * blockDim.x) {
        const long i = b * HW + q * 4;
        long long lab[4] = {0, 0, 0, 0};
        if (labels) tanimoto_labels4(labels, i, lab);
#pragma unroll
        for (int t = 0; t < TN_MAX_TERMS; ++t) {
            if (t >= nterms) break;
            float tq[4], pq[4], mk[4];
            tanimoto_elem4_shared(terms.t[t], i, lab, tq, pq, mk);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a[t][0] = fmaf(tq[j], pq[j], a[t][0]);
                a[t][1] += tq[j] * tq[j] + pq[j] * pq[j];
                a[t][2] += tq[j];
                a[t][3] += pq[j];
            }
        }
    }
#pragma unroll
    for (int t = 0; t < TN_MAX_TERMS; ++t) {
        if (t >= nterms) break;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double w = cnb_warp_sum((double)a[t][k]);
            if ((threadIdx.x & 31) == 0) atomicAdd(&sh[t * 4 + k], w);
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < nterms * 4) atomicAdd(sums + ((long)(threadIdx.x / 4) * B + b) * 4 + (threadIdx.x & 3), sh[threadIdx.x]);
}

// grid = (chunks, B); tanimoto_bwd_kernel for every term in one pass
__global__ void __launch_bounds__(256) tanimoto_bwd_fused_kernel(TanimotoTerms terms, int nterms, int B, long HW, const long long* __restrict__ labels,
                                                                 const float* __restrict__ coef, const float* __restrict__ gscale) {
    CNB_PDL_SYNC();
    const long b = blockIdx.y;
    const float g = gscale ? gscale[0] : 1.f;
    float c[TN_MAX_TERMS][4];
#pragma unroll
    for (int t = 0; t < TN_MAX_TERMS; ++t)
#pragma unroll
        for (int k = 0; k < 4; ++k) c[t][k] = t < nterms ? coef[((long)t * B + b) * 4 + k] * g : 0.f;
    for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < HW; q += (long)gridDim.x * blockDim.x) {
        const long i = b * HW + q * 4;
        long long lab[4] = {0, 0, 0, 0};
        if (labels) tanimoto_labels4(labels, i, lab);
#pragma unroll
        for (int t = 0; t < TN_MAX_TERMS; ++t) {
            if (t >= nterms) break;
            float tq[4], pq[4], mk[4], d[4];
            tanimoto_elem4_shared(terms.t[t], i, lab, tq, pq, mk);
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = mk[j] * (c[t][0] * tq[j] + c[t][1] * pq[j] + c[t][2] * (1.f - tq[j]) + c[t][3] * (1.f - pq[j]));
            *reinterpret_cast<float4*>(terms.t[t].dpred + i) = make_float4(d[0], d[1], d[2], d[3]);
        }
    }
}

// grid = (chunks, B, nterms); sums[(term*B + b)*4 + {P, S, St, Sp}] in fp64
__global__ void __launch_bounds__(256) tanimoto_sums_kernel(TanimotoTerms terms, int B, long HW, double* __restrict__ sums) {
    CNB_PDL_SYNC();
    __shared__ double sh[4];
    if (threadIdx.x < 4) sh[threadIdx.x] = 0.0;
    __syncthreads();
    const int term = blockIdx.z;
    const long b = blockIdx.y;
    const cnb_tanimoto_term& tm = terms.t[term];
    const long n = (long)tm.C * HW;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (tanimoto_vec_ok(tm, HW)) {  // uniform over the CTA
        for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < HW; q += (long)gridDim.x * blockDim.x) {
            float t[4], p[4], m[4];
            tanimoto_elem4(tm, b, q * 4, HW, t, p, m);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a0 = fmaf(t[j], p[j], a0);
                a1 += t[j] * t[j] + p[j] * p[j];
                a2 += t[j];
                a3 += p[j];
            }
        }
    } else {
        for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
            float t, p, m;
            tanimoto_elem(tm, b, e, HW, t, p, m);
            a0 = fmaf(t, p, a0);
            a1 += t * t + p * p;
            a2 += t;
            a3 += p;
        }
    }
    double d0 = cnb_warp_sum((double)a0), d1 = cnb_warp_sum((double)a1), d2 = cnb_warp_sum((double)a2), d3 = cnb_warp_sum((double)a3);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sh[0], d0);
        atomicAdd(&sh[1], d1);
        atomicAdd(&sh[2], d2);
        atomicAdd(&sh[3], d3);
    }
    __syncthreads();
    if (threadIdx.x < 4) atomicAdd(sums + ((long)term * B + b) * 4 + threadIdx.x, sh[threadIdx.x]);
}

__device__ __forceinline__ void tanimoto_T(double P, double S, double eps, int depth, double& Tv, double& dP, double& dS) {
    double sum_inv = 0.0, sum_b = 0.0, sum_a = 0.0;
    double a = 1.0;
    for (int d = 0; d < depth; ++d) {
        const double bb = -(2.0 * a - 1.0);
        const double den = a * S + bb * P + eps;
        const double inv = 1.0 / den;
        sum_inv += inv;
        sum_b += bb * inv * inv;
        sum_a += a * inv * inv;
        a *= 2.0;
    }
    const double sc = 1.0 / (double)depth;
    Tv = (P + eps) * sum_inv * sc;
    dP = sc * (sum_inv - (P + eps) * sum_b);
    dS = -sc * (P + eps) * sum_a;
}

// coef[(term*B+b)*4] = dloss/d{P, S, P', S'} folded with weight/(2B); loss[0] total, loss[1+term] per term (zeroed by the caller: one
// CTA per 256 (term, sample) pairs adds its share -- a single CTA for B <= 85 with three terms, i.e. a fixed summation order there)
// variant (reference LOSS_DICT, models/lightning.py:38-88): 0 = TanimotoComplementLoss (losses.py:103-218, `depth` terms);
// 1 = TanimotoDistLoss (losses.py:221-340: T = (P+eps)/(S-P+eps), i.e. the depth-1 case of the same closed form);
// 2 = CombinedLoss([TanimotoDistLoss, TanimotoComplementLoss]) = the mean of the two (losses.py:62-100)
__global__ void __launch_bounds__(256) tanimoto_finalize_kernel(TanimotoTerms terms, int nterms, int B, long HW, float smooth, int depth,
                                                               int variant, const double* __restrict__ sums, float* __restrict__ coef,
                                                               float* __restrict__ loss) {
    CNB_PDL_SYNC();
    __shared__ double lsum[TN_MAX_TERMS];
    if (threadIdx.x < TN_MAX_TERMS) lsum[threadIdx.x] = 0.0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nterms * B; i += gridDim.x * blockDim.x) {
        const int term = i / B;
        const double N = (double)terms.t[term].C * (double)HW;
        const double P = sums[i * 4 + 0], S = sums[i * 4 + 1], St = sums[i * 4 + 2], Sp = sums[i * 4 + 3];
        const double Pc = N - St - Sp + P, Sc = 2.0 * N - 2.0 * St - 2.0 * Sp + S;
        double T1, T1p, T1s, T2, T2p, T2s;
        tanimoto_T(P, S, (double)smooth, variant == 1 ? 1 : depth, T1, T1p, T1s);
        tanimoto_T(Pc, Sc, (double)smooth, variant == 1 ? 1 : depth, T2, T2p, T2s);
        if (variant == 2) {  // mean of the depth-1 (TanimotoDistLoss) and depth-D (complement) forms
            double U1, U1p, U1s, U2, U2p, U2s;
            tanimoto_T(P, S, (double)smooth, 1, U1, U1p, U1s);
            tanimoto_T(Pc, Sc, (double)smooth, 1, U2, U2p, U2s);
            T1 = 0.5 * (T1 + U1), T1p = 0.5 * (T1p + U1p), T1s = 0.5 * (T1s + U1s);
            T2 = 0.5 * (T2 + U2), T2p = 0.5 * (T2p + U2p), T2s = 0.5 * (T2s + U2s);
        }
        const double l = 0.5 * ((1.0 - T1) + (1.0 - T2));
        atomicAdd(&lsum[term], l / (double)B);
        const double k = -(double)terms.t[term].weight / (2.0 * (double)B);
        coef[i * 4 + 0] = (float)(k * T1p);
        coef[i * 4 + 1] = (float)(k * 2.0 * T1s);
        coef[i * 4 + 2] = (float)(-k * T2p);
        coef[i * 4 + 3] = (float)(-k * 2.0 * T2s);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int t = 0; t < nterms; ++t) {
            atomicAdd(&loss[1 + t], (float)lsum[t]);
            tot += (double)terms.t[t].weight * lsum[t];
        }
        atomicAdd(&loss[0], (float)tot);
    }
}

// dpred = g * mask * (c0 t' + c1 p' + c2 (1-t') + c3 (1-p'))
__global__ void __launch_bounds__(256) tanimoto_bwd_kernel(TanimotoTerms terms, int B, long HW, const float* __restrict__ coef,
                                                          const float* __restrict__ gscale) {
    CNB_PDL_SYNC();
    const int term = blockIdx.z;
    const long b = blockIdx.y;
    const cnb_tanimoto_term& tm = terms.t[term];
    const long n = (long)tm.C * HW;
    const float g = gscale ? gscale[0] : 1.f;
    const float* cf = coef + ((long)term * B + b) * 4;
    const float c0 = cf[0] * g, c1 = cf[1] * g, c2 = cf[2] * g, c3 = cf[3] * g;
    if (tanimoto_vec_ok(tm, HW)) {
        for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < HW; q += (long)gridDim.x * blockDim.x) {
            float t[4], p[4], m[4], d[4];
            tanimoto_elem4(tm, b, q * 4, HW, t, p, m);
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = m[j] * (c0 * t[j] + c1 * p[j] + c2 * (1.f - t[j]) + c3 * (1.f - p[j]));
            *reinterpret_cast<float4*>(tm.dpred + b * HW + q * 4) = make_float4(d[0], d[1], d[2], d[3]);
        }
        return;
    }
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        float t, p, m;
        tanimoto_elem(tm, b, e, HW, t, p, m);
        tm.dpred[b * n + e] = m * (c0 * t + c1 * p + c2 * (1.f - t) + c3 * (1.f - p));
    }
}

// ---------------------------------------------------------------------------------------------
// validation counts (reference _shared_eval_step, models/lightning.py:374-481): ONE pass over the three predictions and the labels
// gives everything the scorers need.  out[12] (fp64, zeroed by the caller):
//   0 valid pixels, 1 sum |dist - bdist|, 2 sum (dist - bdist)^2,
//   3..6 edge confusion (tp, fp, fn, tn) of (edge > thresh) vs (y == edge_class), 7..10 crop confusion of (crop > thresh) vs (0 < y < edge_class),
//   11 unused.  Pixels with y == -1 are excluded (the reference masks them out only when the batch holds any: same result).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) val_counts_kernel(const float* __restrict__ dist, const float* __restrict__ edge,
                                                        const float* __restrict__ crop, const long long* __restrict__ y,
                                                        const float* __restrict__ bdist, long n, int edge_class, float thresh,
                                                        double* __restrict__ out) {
    CNB_PDL_SYNC();
    __shared__ double sh[11];
    if (threadIdx.x < 11) sh[threadIdx.x] = 0.0;
    __syncthreads();
    float mae = 0.f, mse = 0.f;
    int cnt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // valid, edge tp fp fn tn, crop tp fp fn tn
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long long lab = y[i];
        if (lab == -1) continue;
        cnt[0]++;
        const float d = dist[i] - bdist[i];
        mae += fabsf(d);
        mse = fmaf(d, d, mse);
        const bool te = lab == edge_class, pe = edge[i] > thresh;
        cnt[1 + (pe ? (te ? 0 : 1) : (te ? 2 : 3))]++;
        const bool tc = lab > 0 && lab < edge_class, pc = crop[i] > thresh;
        cnt[5 + (pc ? (tc ? 0 : 1) : (tc ? 2 : 3))]++;
    }
    double v[11];
    v[0] = (double)cnt[0], v[1] = (double)mae, v[2] = (double)mse;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[3 + k] = (double)cnt[1 + k];
#pragma unroll
    for (int k = 0; k < 11; ++k) {
        const double w = cnb_warp_sum(v[k]);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sh[k], w);
    }
    __syncthreads();
    if (threadIdx.x < 11) atomicAdd(out + threadIdx.x, sh[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------
// optimiser
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const float* __restrict__ g, long n, float* __restrict__ out) {
    CNB_PDL_SYNC();
    __shared__ float sh;
    if (threadIdx.x == 0) sh = 0.f;
    __syncthreads();
    float s = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) s = fmaf(g[i], g[i], s);
    s = cnb_warp_sum(s);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sh, s);
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(out, sh);
}

// torch.optim.AdamW semantics (decoupled decay, bias-corrected); optional global-norm clipping as torch.nn.utils.clip_grad_norm_
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long n, const float* __restrict__ hyper, float beta1,
                                                   float beta2, float eps, float wd, float grad_scale, float clip_norm,
                                                   const float* __restrict__ norm_ws) {
    CNB_PDL_SYNC();
    const float lr = hyper[0], step = hyper[1];
    float gs = grad_scale;
    if (clip_norm > 0.f) {
        const float total = sqrtf(norm_ws[0]) * fabsf(grad_scale);
        const float coef = clip_norm / (total + 1e-6f);
        if (coef < 1.f) gs *= coef;
    }
    const float bc1 = 1.f - powf(beta1, step);
    const float bc2 = 1.f - powf(beta2, step);
    const float step_size = lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float gi = g[i] * gs;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        pi -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
        p[i] = pi;
    }
}

}  // namespace cnb
