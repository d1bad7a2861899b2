// Whole-graph GraphNorm (PyG GraphNorm with batch=None; call sites impl/models.py:165, 249, 257, 266)
// fused with the activation / dropout the reference applies right after it (:166, :251, :258-259).
//
//   mu = mean_rows(x); o = x - mean_scale*mu; var = mean_rows(o^2) = E[x^2] - (2a - a^2) mu^2
//   out = keep * pscale * act(weight * o / sqrt(var + eps) + bias)
//
// Forward = column sums of x and x^2 (fp64 accumulators, per-CTA partials reduced in CTA order, so the
// result is run-to-run deterministic) -> per-column constants -> one elementwise pass.
// Backward = column sums S1 = sum u, S2 = sum u*yhat with u = dout*keep*pscale*act'(pre)
//   dweight = S2, dbias = S1, sum_do = rstd*w*(S1 - sum(yhat)*S2/N), dmean_scale = -mu*sum_do
//   dx = rstd*w*u - (rstd*w*S2/N)*yhat - mean_scale*sum_do/N                 (one elementwise pass)
// Bound: HBM/L2 bandwidth; algorithmic bytes fwd = 4*N*C read twice (second pass is L2-resident) +
// 4*N*C written (+ N*C mask bytes).
#include "common.cuh"

namespace glass {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxPartialCtas = 296;  // 2 CTAs per SM on 148 SMs; fixed so the workspace size is device independent

// stats rows
enum { ST_SCALE = 0, ST_AM = 1, ST_MU = 2, ST_RSTD = 3, ST_BIAS = 4 };

// Finalisation: one warp per column; lane l adds partials l, l+32, ... in order (all loads issued up
// front), then a fixed butterfly -> deterministic.  (A "last CTA finalises" variant was measured slower:
// one SM pulling all 296 x 2C partials is latency bound.)
struct Fin {
    const float* weight;
    const float* bias;        // fwd only
    const float* mean_scale;
    float eps;                // fwd only
    float* stats;             // fwd: written; bwd: read
    float* coef;              // bwd
    float* dweight;
    float* dbias;
    float* dmean_scale;
};

__device__ __forceinline__ void reduce_partials(const double* partial, int nblk, int c, int col, double& s, double& q) {
    const int lane = threadIdx.x & 31;
    double vs[(kMaxPartialCtas + 31) / 32], vq[(kMaxPartialCtas + 31) / 32];
#pragma unroll
    for (int i = 0; i < (kMaxPartialCtas + 31) / 32; ++i) {
        const int b = lane + 32 * i;
        vs[i] = b < nblk ? __ldcg(partial + ((int64_t)b * 2 + 0) * c + col) : 0.0;
        vq[i] = b < nblk ? __ldcg(partial + ((int64_t)b * 2 + 1) * c + col) : 0.0;
    }
    s = 0.0;
    q = 0.0;
#pragma unroll
    for (int i = 0; i < (kMaxPartialCtas + 31) / 32; ++i) {
        s += vs[i];
        q += vq[i];
    }
    s = warp_sum(s);
    q = warp_sum(q);
}

__device__ __forceinline__ void finalize_fwd_col(const Fin& f, double s, double q, int64_t n, int c, int col) {
    const double mu = s / (double)n, ex2 = q / (double)n;
    const float muf = (float)mu;
    const float am = __fmul_rn(muf, f.mean_scale[col]);  // mean * mean_scale, rounded like the reference
    // var of (x - am): E[x^2] - 2*am*mu + am^2, evaluated in fp64
    double var = ex2 - 2.0 * (double)am * mu + (double)am * (double)am;
    if (var < 0.0) var = 0.0;
    const float std_ = sqrtf((float)var + f.eps);
    const float rstd = 1.0f / std_;
    f.stats[ST_SCALE * c + col] = f.weight[col] * rstd;
    f.stats[ST_AM * c + col] = am;
    f.stats[ST_MU * c + col] = muf;
    f.stats[ST_RSTD * c + col] = rstd;
    f.stats[ST_BIAS * c + col] = f.bias[col];  // kept with the statistics so that backward can rebuild the pre-activation
}

// coef rows: alpha, beta, gamma
__device__ __forceinline__ void finalize_bwd_col(const Fin& f, double s1, double s2, int64_t n, int c, int col) {
    const double w = f.weight[col], a = f.mean_scale[col];
    const double rstd = f.stats[ST_RSTD * c + col], mu = f.stats[ST_MU * c + col], am = f.stats[ST_AM * c + col];
    const double N = (double)n;
    const double sum_yhat = rstd * N * (mu - am);
    const double sum_do = rstd * w * (s1 - sum_yhat * s2 / N);
    f.dweight[col] = (float)s2;
    f.dbias[col] = (float)s1;
    f.dmean_scale[col] = (float)(-mu * sum_do);
    f.coef[0 * c + col] = (float)(rstd * w);
    f.coef[1 * c + col] = (float)(-rstd * w * s2 / N);
    f.coef[2 * c + col] = (float)(-a * sum_do / N);
}

template <int VEC, bool BWD>
__global__ void __launch_bounds__(kThreads)
k_colsums(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dout, int64_t lddo,
          const float* __restrict__ stats, const float* __restrict__ bias, int act, const uint8_t* __restrict__ keep,
          float pscale, int64_t n, int c, double* __restrict__ partial) {
    // thread -> (column vector cvl, row lane rl).  CVB column vectors are processed per pass.
    const int CV = (c + VEC - 1) / VEC;
    const int CVB = CV < kThreads ? CV : kThreads;
    const int nrl = kThreads / CVB;
    const int cvl = threadIdx.x % CVB, rl = threadIdx.x / CVB;
    const bool active = rl < nrl;
    __shared__ double sm[2][kThreads * VEC];
    for (int cv0 = 0; cv0 < CV; cv0 += CVB) {
        const int cv = cv0 + cvl;
        double s[VEC], q[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) s[k] = q[k] = 0.0;
        if (active && cv < CV) {
            const int col = cv * VEC;
            float sc[VEC], am[VEC], rs[VEC], bs[VEC];
            if (BWD) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    int cc = col + k < c ? col + k : c - 1;
                    sc[k] = stats[ST_SCALE * c + cc];
                    am[k] = stats[ST_AM * c + cc];
                    rs[k] = stats[ST_RSTD * c + cc];
                    bs[k] = bias[cc];
                }
            }
#pragma unroll 4
            for (int64_t r = (int64_t)blockIdx.x * nrl + rl; r < n; r += (int64_t)gridDim.x * nrl) {
                float xv[VEC], gv[VEC];
                if (VEC == 4) {
                    float4 t = ldg_f4(x + r * ldx + col);
                    xv[0] = t.x, xv[1] = t.y, xv[2] = t.z, xv[3] = t.w;
                    if (BWD) {
                        float4 g = ldg_f4(dout + r * lddo + col);
                        gv[0] = g.x, gv[1] = g.y, gv[2] = g.z, gv[3] = g.w;
                    }
                } else {
                    xv[0] = x[r * ldx + col];
                    if (BWD) gv[0] = dout[r * lddo + col];
                }
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    if (!BWD) {
                        s[k] += (double)xv[k];
                        q[k] += (double)xv[k] * (double)xv[k];
                    } else {
                        float o = xv[k] - am[k];
                        float pre = fmaf(sc[k], o, bs[k]);
                        float u = gv[k] * act_grad_from_out(act_fwd(pre, act), act);
                        if (keep) u *= keep[r * (int64_t)c + col + k] ? pscale : 0.f;
                        s[k] += (double)u;
                        q[k] += (double)u * (double)(o * rs[k]);
                    }
                }
            }
        }
        // reduce over row lanes in a fixed order
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            sm[0][threadIdx.x * VEC + k] = s[k];
            sm[1][threadIdx.x * VEC + k] = q[k];
        }
        __syncthreads();
        if (rl == 0 && cv < CV) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                if (cv * VEC + k >= c) continue;
                double ts = 0.0, tq = 0.0;
                for (int j = 0; j < nrl; ++j) {
                    ts += sm[0][(j * CVB + cvl) * VEC + k];
                    tq += sm[1][(j * CVB + cvl) * VEC + k];
                }
                partial[((int64_t)blockIdx.x * 2 + 0) * c + cv * VEC + k] = ts;
                partial[((int64_t)blockIdx.x * 2 + 1) * c + cv * VEC + k] = tq;
            }
        }
        __syncthreads();
    }
}

template <bool BWD>
__global__ void __launch_bounds__(128) k_gn_finalize(const double* partial, int nblk, int64_t n, int c, const Fin fin) {
    const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (col >= c) return;
    double a, b;
    reduce_partials(partial, nblk, c, col, a, b);
    if ((threadIdx.x & 31) != 0) return;
    if (BWD) finalize_bwd_col(fin, a, b, n, c, col);
    else finalize_fwd_col(fin, a, b, n, c, col);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
k_gn_apply(const float* __restrict__ x, int64_t ldx, const float* __restrict__ stats, const float* __restrict__ bias,
           int act, const uint8_t* __restrict__ keep, float pscale, float* __restrict__ out, int64_t ldo, int64_t n,
           int c) {
    const int CV = (c + VEC - 1) / VEC;
    const int64_t total = n * CV, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t r = e / CV;
        const int col = (int)(e % CV) * VEC;
        float xv[VEC], ov[VEC];
        if (VEC == 4) {
            float4 t = ldg_f4(x + r * ldx + col);
            xv[0] = t.x, xv[1] = t.y, xv[2] = t.z, xv[3] = t.w;
        } else {
            xv[0] = x[r * ldx + col];
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            float pre = fmaf(stats[ST_SCALE * c + col + k], xv[k] - stats[ST_AM * c + col + k], bias[col + k]);
            float v = act_fwd(pre, act);
            if (keep) v = keep[r * (int64_t)c + col + k] ? v * pscale : 0.f;
            ov[k] = v;
        }
        if (VEC == 4) *reinterpret_cast<float4*>(out + r * ldo + col) = make_float4(ov[0], ov[1], ov[2], ov[3]);
        else out[r * ldo + col] = ov[0];
    }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
k_gn_bwd_apply(const float* __restrict__ dout, int64_t lddo, const float* __restrict__ x, int64_t ldx,
               const float* __restrict__ stats, const float* __restrict__ bias, const float* __restrict__ coef, int act,
               const uint8_t* __restrict__ keep, float pscale, float* __restrict__ dx, int64_t lddx, int64_t n, int c) {
    const int CV = (c + VEC - 1) / VEC;
    const int64_t total = n * CV, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t r = e / CV;
        const int col = (int)(e % CV) * VEC;
        float xv[VEC], gv[VEC], ov[VEC];
        if (VEC == 4) {
            float4 t = ldg_f4(x + r * ldx + col), g = ldg_f4(dout + r * lddo + col);
            xv[0] = t.x, xv[1] = t.y, xv[2] = t.z, xv[3] = t.w;
            gv[0] = g.x, gv[1] = g.y, gv[2] = g.z, gv[3] = g.w;
        } else {
            xv[0] = x[r * ldx + col];
            gv[0] = dout[r * lddo + col];
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const int cc = col + k;
            float o = xv[k] - stats[ST_AM * c + cc];
            float pre = fmaf(stats[ST_SCALE * c + cc], o, bias[cc]);
            float u = gv[k] * act_grad_from_out(act_fwd(pre, act), act);
            if (keep) u *= keep[r * (int64_t)c + cc] ? pscale : 0.f;
            float yhat = o * stats[ST_RSTD * c + cc];
            ov[k] = fmaf(coef[0 * c + cc], u, fmaf(coef[1 * c + cc], yhat, coef[2 * c + cc]));
        }
        if (VEC == 4) *reinterpret_cast<float4*>(dx + r * lddx + col) = make_float4(ov[0], ov[1], ov[2], ov[3]);
        else dx[r * lddx + col] = ov[0];
    }
}

inline int partial_ctas(int64_t n, int c, int vec) {
    int cv = (c + vec - 1) / vec;
    int cvb = cv < kThreads ? cv : kThreads;
    int nrl = kThreads / cvb;
    int64_t want = ceil_div(n, (int64_t)nrl * 8);  // >= 8 rows per thread
    if (want < 1) want = 1;
    return (int)(want < kMaxPartialCtas ? want : kMaxPartialCtas);
}

inline bool vec_ok(int c, std::initializer_list<int64_t> lds, std::initializer_list<const void*> ptrs) {
    if (c % 4) return false;
    for (int64_t l : lds)
        if (l % 4) return false;
    for (const void* p : ptrs)
        if ((uintptr_t)p % 16) return false;
    return true;
}

}  // namespace
}  // namespace glass

using namespace glass;

extern "C" size_t glass_graphnorm_workspace_bytes(int64_t n, int c) {
    if (n < 0 || c <= 0) return 0;
    return align_up((size_t)kMaxPartialCtas * 2 * (size_t)c * sizeof(double), 256) + align_up(3 * (size_t)c * sizeof(float), 256);
}

extern "C" int glass_graphnorm_fwd(const float* x, int64_t ldx, const float* weight, const float* bias,
                                   const float* mean_scale, float eps, int act, const uint8_t* keep, float pscale,
                                   float* out, int64_t ldo, float* stats, int64_t n, int c, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    GLASS_CHECK_ARG(x && weight && bias && mean_scale && out && stats && n > 0 && c > 0 && ldx >= c && ldo >= c,
                    "graphnorm_fwd: bad arguments");
    if (workspace_bytes < glass_graphnorm_workspace_bytes(n, c) || !workspace) {
        set_error("graphnorm_fwd: workspace too small");
        return GLASS_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    double* partial = static_cast<double*>(workspace);
    const bool vec = vec_ok(c, {ldx, ldo}, {x, out});
    const int nblk = partial_ctas(n, c, vec ? 4 : 1);
    Fin fin{};
    fin.weight = weight, fin.bias = bias, fin.mean_scale = mean_scale, fin.eps = eps, fin.stats = stats;
    if (vec) k_colsums<4, false><<<nblk, kThreads, 0, st>>>(x, ldx, nullptr, 0, nullptr, nullptr, 0, nullptr, 0.f, n, c, partial);
    else k_colsums<1, false><<<nblk, kThreads, 0, st>>>(x, ldx, nullptr, 0, nullptr, nullptr, 0, nullptr, 0.f, n, c, partial);
    k_gn_finalize<false><<<(unsigned)ceil_div(c, 4), 128, 0, st>>>(partial, nblk, n, c, fin);
    const int64_t work = n * (vec ? c / 4 : c);
    unsigned grid = (unsigned)std::min<int64_t>(ceil_div(work, kThreads), (int64_t)sm_count() * 8);
    if (vec) k_gn_apply<4><<<grid, kThreads, 0, st>>>(x, ldx, stats, bias, act, keep, pscale, out, ldo, n, c);
    else k_gn_apply<1><<<grid, kThreads, 0, st>>>(x, ldx, stats, bias, act, keep, pscale, out, ldo, n, c);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

extern "C" int glass_graphnorm_bwd(const float* dout, int64_t lddo, const float* x, int64_t ldx, const float* weight,
                                   const float* mean_scale, const float* stats, int act, const uint8_t* keep,
                                   float pscale, float* dx, int64_t lddx, float* dweight, float* dbias,
                                   float* dmean_scale, int64_t n, int c, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    GLASS_CHECK_ARG(dout && x && weight && mean_scale && stats && dx && dweight && dbias && dmean_scale && n > 0 &&
                        c > 0 && ldx >= c && lddo >= c && lddx >= c,
                    "graphnorm_bwd: bad arguments");
    if (workspace_bytes < glass_graphnorm_workspace_bytes(n, c) || !workspace) {
        set_error("graphnorm_bwd: workspace too small");
        return GLASS_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    double* partial = static_cast<double*>(workspace);
    float* coef = reinterpret_cast<float*>(static_cast<char*>(workspace) +
                                           align_up((size_t)kMaxPartialCtas * 2 * (size_t)c * sizeof(double), 256));
    const float* bias = stats + ST_BIAS * (int64_t)c;  // forward bias saved with the statistics
    const bool vec = vec_ok(c, {ldx, lddo, lddx}, {x, dout, dx});
    const int nblk = partial_ctas(n, c, vec ? 4 : 1);
    Fin fin{};
    fin.weight = weight, fin.mean_scale = mean_scale, fin.stats = const_cast<float*>(stats), fin.coef = coef;
    fin.dweight = dweight, fin.dbias = dbias, fin.dmean_scale = dmean_scale;
    if (vec) k_colsums<4, true><<<nblk, kThreads, 0, st>>>(x, ldx, dout, lddo, stats, bias, act, keep, pscale, n, c, partial);
    else k_colsums<1, true><<<nblk, kThreads, 0, st>>>(x, ldx, dout, lddo, stats, bias, act, keep, pscale, n, c, partial);
    k_gn_finalize<true><<<(unsigned)ceil_div(c, 4), 128, 0, st>>>(partial, nblk, n, c, fin);
    const int64_t work = n * (vec ? c / 4 : c);
    unsigned grid = (unsigned)std::min<int64_t>(ceil_div(work, kThreads), (int64_t)sm_count() * 8);
    if (vec) k_gn_bwd_apply<4><<<grid, kThreads, 0, st>>>(dout, lddo, x, ldx, stats, bias, coef, act, keep, pscale, dx, lddx, n, c);
    else k_gn_bwd_apply<1><<<grid, kThreads, 0, st>>>(dout, lddo, x, ldx, stats, bias, coef, act, keep, pscale, dx, lddx, n, c);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}
