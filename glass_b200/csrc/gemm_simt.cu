// Label-mixed pair of Linear layers, SIMT fp32 path (any H; the tcgen05 path lives in gemm_tc.cu).
// Reference: impl/models.py:158-162 (trans_fns + activation + z_ratio mix) and :167-173
// (concat + comb_fns + mix).  The two weight sets are evaluated in the same CTA tile so that the
// label select `where(mask, z*p1+(1-z)*p0, z*p0+(1-z)*p1)` happens in registers in the epilogue,
// and the concat of models.py:167 is virtual (two A sources, no materialised cat).
//
// Backward:
//   dP0 = c0(mask) * dOut * act'(p0), dP1 = c1(mask) * dOut * act'(p1)      (built on the fly)
//   dA  = [dP0|dP1] [W0;W1]                       (k_pair_bwd_dx)
//   dW  = [dP0|dP1]^T A, db = colsum([dP0|dP1])   (k_pair_bwd_dw: split over rows, partials reduced
//                                                  in a fixed order by k_pair_bwd_dw_reduce)
#include "common.cuh"

namespace glass {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4, kThreads = 256;  // BK 32 measured 10-40 % slower at K = 17..40
constexpr int PAD = 4;

struct ASrc {
    const float* a1;
    int64_t lda1;
    int k1;
    const float* a2;
    int64_t lda2;
    int k2;
    __device__ __forceinline__ float at(int64_t m, int k) const {
        return k < k1 ? a1[m * lda1 + k] : a2[m * lda2 + (k - k1)];
    }
};

// coefficient of p0 / p1 in the mix for a row (impl/models.py:161-162)
__device__ __forceinline__ void mix_coef(uint8_t labelled, float z, float& c0, float& c1) {
    c1 = labelled ? z : 1.f - z;
    c0 = labelled ? 1.f - z : z;
}

__global__ void __launch_bounds__(kThreads) k_pair_fwd(ASrc A, const float* __restrict__ w0, const float* __restrict__ b0,
                                                       const float* __restrict__ w1, const float* __restrict__ b1,
                                                       const uint8_t* __restrict__ mask, float z, int act,
                                                       float* __restrict__ out, int64_t ldo, float* __restrict__ acts,
                                                       int64_t n, int h) {
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float B0s[BK][BN + PAD];
    __shared__ __align__(16) float B1s[BK][BN + PAD];
    const int K = A.k1 + A.k2;
    const int tid = threadIdx.x, tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int64_t m_blk = (int64_t)blockIdx.x * BM;
    const int n_blk = blockIdx.y * BN;
    float acc0[TM][TN] = {}, acc1[TM][TN] = {};

    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < BM * BK / kThreads; ++i) {
            int e = tid + i * kThreads, mm = e / BK, kk = e % BK;
            int64_t m = m_blk + mm;
            As[kk][mm] = (m < n && k0 + kk < K) ? A.at(m, k0 + kk) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < BN * BK / kThreads; ++i) {
            int e = tid + i * kThreads, nn = e / BK, kk = e % BK;
            bool ok = (n_blk + nn < h) && (k0 + kk < K);
            B0s[kk][nn] = ok ? w0[(int64_t)(n_blk + nn) * K + k0 + kk] : 0.f;
            B1s[kk][nn] = ok ? w1[(int64_t)(n_blk + nn) * K + k0 + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
            float4 p = *reinterpret_cast<const float4*>(&B0s[kk][tx * TN]);
            float4 q = *reinterpret_cast<const float4*>(&B1s[kk][tx * TN]);
            const float av[4] = {a.x, a.y, a.z, a.w}, pv[4] = {p.x, p.y, p.z, p.w}, qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    acc0[i][j] = fmaf(av[i], pv[j], acc0[i][j]);
                    acc1[i][j] = fmaf(av[i], qv[j], acc1[i][j]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int64_t m = m_blk + ty * TM + i;
        if (m >= n) continue;
        float c0, c1;
        mix_coef(mask[m], z, c0, c1);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int c = n_blk + tx * TN + j;
            if (c >= h) continue;
            float p0 = act_fwd(acc0[i][j] + b0[c], act);
            float p1 = act_fwd(acc1[i][j] + b1[c], act);
            if (acts) {
                acts[m * (2 * (int64_t)h) + c] = p0;
                acts[m * (2 * (int64_t)h) + h + c] = p1;
            }
            // z*p1 + (1-z)*p0 for labelled rows, z*p0 + (1-z)*p1 otherwise (two roundings, like the reference)
            out[m * ldo + c] = __fadd_rn(__fmul_rn(c1, p1), __fmul_rn(c0, p0));
        }
    }
}

struct DPSrc {  // dP[m, j], j in [0, 2h): gradient w.r.t. the pre-activation of branch j / h
    const float* dout;
    int64_t lddo;
    const float* acts;  // NULL when act == NONE
    const uint8_t* mask;
    float z;
    int act;
    int h;
    __device__ __forceinline__ float at(int64_t m, int j) const {
        float c0, c1;
        mix_coef(mask[m], z, c0, c1);
        const int c = j < h ? j : j - h;
        float g = dout[m * lddo + c] * (j < h ? c0 : c1);
        if (acts) g *= act_grad_from_out(acts[m * (2 * (int64_t)h) + j], act);
        return g;
    }
};

// dA[n, K] = dP[n, 2h] * Wcat[2h, K]
__global__ void __launch_bounds__(kThreads) k_pair_bwd_dx(DPSrc P, const float* __restrict__ w0,
                                                          const float* __restrict__ w1, float* __restrict__ da1,
                                                          int64_t ldda1, int k1, float* __restrict__ da2, int64_t ldda2,
                                                          int k2, int64_t n) {
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];
    const int K = k1 + k2, J = 2 * P.h;
    const int tid = threadIdx.x, tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int64_t m_blk = (int64_t)blockIdx.x * BM;
    const int n_blk = blockIdx.y * BN;  // over K
    float acc[TM][TN] = {};
    for (int j0 = 0; j0 < J; j0 += BK) {
#pragma unroll
        for (int i = 0; i < BM * BK / kThreads; ++i) {
            int e = tid + i * kThreads, mm = e / BK, jj = e % BK;
            int64_t m = m_blk + mm;
            As[jj][mm] = (m < n && j0 + jj < J) ? P.at(m, j0 + jj) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < BN * BK / kThreads; ++i) {
            int e = tid + i * kThreads, jj = e / BN, kk = e % BN;
            int j = j0 + jj, k = n_blk + kk;
            float v = 0.f;
            if (j < J && k < K) v = (j < P.h ? w0 : w1)[(int64_t)(j < P.h ? j : j - P.h) * K + k];
            Bs[jj][kk] = v;
        }
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < BK; ++jj) {
            float4 a = *reinterpret_cast<const float4*>(&As[jj][ty * TM]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[jj][tx * TN]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int64_t m = m_blk + ty * TM + i;
        if (m >= n) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int k = n_blk + tx * TN + j;
            if (k >= K) continue;
            if (k < k1) {
                if (da1) da1[m * ldda1 + k] = acc[i][j];
            } else if (da2) {
                da2[m * ldda2 + (k - k1)] = acc[i][j];
            }
        }
    }
}

// part[s][j][k] (k < K) = sum over the s-th row chunk of dP[m, j] * A[m, k];  part[s][j][K] = sum dP[m, j]
__global__ void __launch_bounds__(kThreads) k_pair_bwd_dw(DPSrc P, ASrc A, float* __restrict__ part, int64_t n,
                                                          int64_t rows_per_split) {
    __shared__ __align__(16) float Ps[BK][BM + PAD];  // [row m][j]
    __shared__ __align__(16) float Xs[BK][BN + PAD];  // [row m][k]
    const int K = A.k1 + A.k2, J = 2 * P.h;
    const int tid = threadIdx.x, tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int j_blk = blockIdx.x * BM, k_blk = blockIdx.y * BN;
    const int64_t m_lo = (int64_t)blockIdx.z * rows_per_split;
    const int64_t m_hi = m_lo + rows_per_split < n ? m_lo + rows_per_split : n;
    float acc[TM][TN] = {};
    float bacc[TM] = {};
    for (int64_t m0 = m_lo; m0 < m_hi; m0 += BK) {
#pragma unroll
        for (int i = 0; i < BM * BK / kThreads; ++i) {
            int e = tid + i * kThreads, mm = e / BM, jj = e % BM;
            int64_t m = m0 + mm;
            Ps[mm][jj] = (m < m_hi && j_blk + jj < J) ? P.at(m, j_blk + jj) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < BN * BK / kThreads; ++i) {
            int e = tid + i * kThreads, mm = e / BN, kk = e % BN;
            int64_t m = m0 + mm;
            Xs[mm][kk] = (m < m_hi && k_blk + kk < K) ? A.at(m, k_blk + kk) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < BK; ++mm) {
            float4 a = *reinterpret_cast<const float4*>(&Ps[mm][ty * TM]);
            float4 b = *reinterpret_cast<const float4*>(&Xs[mm][tx * TN]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                bacc[i] += av[i];
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
    float* p = part + (int64_t)blockIdx.z * J * (K + 1);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int j = j_blk + ty * TM + i;
        if (j >= J) continue;
#pragma unroll
        for (int jj = 0; jj < TN; ++jj) {
            int k = k_blk + tx * TN + jj;
            if (k < K) p[(int64_t)j * (K + 1) + k] = acc[i][jj];
        }
        if (blockIdx.y == 0 && tx == 0) p[(int64_t)j * (K + 1) + K] = bacc[i];
    }
}

// 512 threads = 32 output elements x 16 split lanes; lane q adds splits q, q+16, ... in order, then the 16
// lane sums are added in lane order (deterministic).  A serial loop over up to 296 splits per output was
// latency bound (21 us at component's H = 17).
constexpr int kRedLanes = 16;
__global__ void __launch_bounds__(32 * kRedLanes) k_pair_bwd_dw_reduce(const float* __restrict__ part, int splits,
                                                                      int h, int K, float* __restrict__ dw0,
                                                                      float* __restrict__ db0, float* __restrict__ dw1,
                                                                      float* __restrict__ db1) {
    __shared__ float sm[kRedLanes][33];
    const int J = 2 * h;
    const int el = threadIdx.x & 31, q = threadIdx.x >> 5;
    const int64_t total = (int64_t)J * (K + 1);
    const int64_t e = blockIdx.x * 32ll + el;
    const bool ok = e < total;
    float s = 0.f;
    if (ok) {
#pragma unroll 4
        for (int sp = q; sp < splits; sp += kRedLanes) s += part[(int64_t)sp * total + e];
    }
    sm[q][el] = s;
    __syncthreads();
    if (q == 0 && ok) {
        float t = sm[0][el];
#pragma unroll
        for (int i = 1; i < kRedLanes; ++i) t += sm[i][el];
        const int j = (int)(e / (K + 1)), k = (int)(e % (K + 1));
        const int jr = j < h ? j : j - h;
        if (k < K) (j < h ? dw0 : dw1)[(int64_t)jr * K + k] = t;
        else (j < h ? db0 : db1)[jr] = t;
    }
}

}  // namespace

// shared with gemm_tc.cu
int dw_splits(int64_t n, int h, int k) {
    int64_t tiles = ceil_div(2 * (int64_t)h, BM) * ceil_div(k, BN);
    int64_t s = ceil_div(2 * 148, tiles);
    int64_t max_s = ceil_div(n, BK);   // small graphs: one BK-row pass per CTA, the whole kernel is one load round
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    return (int)s;
}

int pair_fwd_simt(const float* a1, int64_t lda1, int k1, const float* a2, int64_t lda2, int k2, const float* w0,
                  const float* b0, const float* w1, const float* b1, const uint8_t* mask, float z, int act, float* out,
                  int64_t ldo, float* acts, int64_t n, int h, cudaStream_t st) {
    ASrc A{a1, lda1, k1, a2, lda2, k2};
    dim3 grid((unsigned)ceil_div(n, BM), (unsigned)ceil_div(h, BN));
    k_pair_fwd<<<grid, kThreads, 0, st>>>(A, w0, b0, w1, b1, mask, z, act, out, ldo, acts, n, h);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

int pair_bwd_simt(const float* dout, int64_t lddo, const float* acts, const float* a1, int64_t lda1, int k1,
                  const float* a2, int64_t lda2, int k2, const float* w0, const float* w1, const uint8_t* mask, float z,
                  int act, float* da1, int64_t ldda1, float* da2, int64_t ldda2, float* dw0, float* db0, float* dw1,
                  float* db1, int64_t n, int h, void* workspace, cudaStream_t st) {
    const int K = k1 + k2;
    DPSrc P{dout, lddo, act == GLASS_ACT_NONE ? nullptr : acts, mask, z, act, h};
    ASrc A{a1, lda1, k1, a2, lda2, k2};
    if (da1 || da2) {
        dim3 grid((unsigned)ceil_div(n, BM), (unsigned)ceil_div(K, BN));
        k_pair_bwd_dx<<<grid, kThreads, 0, st>>>(P, w0, w1, da1, ldda1, k1, da2, ldda2, k2, n);
        GLASS_LAUNCH_CHECK();
    }
    const int splits = dw_splits(n, h, K);
    const int64_t rows_per_split = ceil_div(ceil_div(n, splits), BK) * BK;
    float* part = static_cast<float*>(workspace);
    dim3 grid((unsigned)ceil_div(2 * h, BM), (unsigned)ceil_div(K, BN), (unsigned)splits);
    k_pair_bwd_dw<<<grid, kThreads, 0, st>>>(P, A, part, n, rows_per_split);
    GLASS_LAUNCH_CHECK();
    int64_t total = 2 * (int64_t)h * (K + 1);
    k_pair_bwd_dw_reduce<<<(unsigned)ceil_div(total, 32), 32 * kRedLanes, 0, st>>>(part, splits, h, K, dw0, db0, dw1, db1);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

}  // namespace glass
