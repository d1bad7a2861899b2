// Whole-graph GraphNorm (PyG GraphNorm with batch=None; call sites impl/models.py:165, 249, 257, 266)
// fused with the activation / dropout the reference applies right after it (:166, :251, :258-259).
//
//   mu = mean_rows(x); o = x - mean_scale*mu; var = mean_rows(o^2) = E[x^2] - (2a - a^2) mu^2
//   out = keep * 1/(1-p) * act(weight * o / sqrt(var + eps) + bias)     (keep: explicit mask or Philox bits)
//
// Forward = column sums of x and x^2 (fp64 accumulators, per-CTA partials reduced in CTA order, so the
// result is run-to-run deterministic) -> per-column constants -> one elementwise pass.
// Backward = column sums S1 = sum u, S2 = sum u*yhat with u = dout*keep*pscale*act'(pre)
//   dweight = S2, dbias = S1, sum_do = rstd*w*(S1 - sum(yhat)*S2/N), dmean_scale = -mu*sum_do
//   dx = rstd*w*u - (rstd*w*S2/N)*yhat - mean_scale*sum_do/N                 (one elementwise pass)
// Bound: HBM/L2 bandwidth; algorithmic bytes fwd = 4*N*C read twice (second pass is L2-resident) +
// 4*N*C written (+ N*C mask bytes).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace glass {
namespace {

namespace cg = cooperative_groups;

constexpr int kThreads = 256;
constexpr int kSumThreads = 256;      // k_colsums; 512 threads per CTA measured slower (em_user bwd 14.6 -> 19.2 us per launch)
constexpr int kMaxPartialCtas = 296;  // 2 CTAs per SM on 148 SMs (4 per SM measured slower); fixed so the workspace size is device independent

// stats rows: ST_* in common.cuh

// Dropout keep decision.  Either an explicit uint8 mask (tests inject one so that a train-mode pass can be
// compared element-wise with the oracle) or a counter-based generator: Philox4x32-10 keyed by a per-device
// seed, with counter = (element index / 4, id of this GraphNorm call).  The call id is drawn from a device
// counter by the finalize kernel and stored with the statistics, so backward regenerates the same bits and a
// CUDA-graph replay draws fresh ones -- no mask tensor is written or read.
struct Drop {
    const uint8_t* keep;                 // explicit mask [n, c] or NULL
    const unsigned long long* rng;       // {seed, call counter, ticket} in device memory or NULL (bits drawn in-kernel)
    const uint32_t* bits;                // packed keep bits (bit r*c + col), written once per call by the finalize kernel
    float pscale;                        // 1 / (1 - p)
    uint32_t thresh;                     // drop iff u32 < thresh  (thresh = p * 2^32)
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

struct DropCtx {   // per-thread constants of the generator
    uint2 key;
    uint32_t call_lo, call_hi;
};
__device__ __forceinline__ DropCtx drop_ctx(const Drop& d, const float* stats, int c) {
    DropCtx x{};
    if (!d.keep && d.rng) {
        const unsigned long long seed = d.rng[0];
        x.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        x.call_lo = __float_as_uint(stats[ST_RNG * c + 0]);
        x.call_hi = c > 1 ? __float_as_uint(stats[ST_RNG * c + 1]) : 0u;
    }
    return x;
}
// multipliers (pscale or 0) of the VEC consecutive elements starting at linear index `lin` (row * c + col)
template <int VEC>
__device__ __forceinline__ void drop_mult(const Drop& d, const DropCtx& x, int64_t lin, float (&m)[VEC]) {
    if (d.keep) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) m[k] = d.keep[lin + k] ? d.pscale : 0.f;
    } else if (d.bits) {
        // VEC == 4: lin % 4 == 0 (c % 4 == 0), so the four bits never straddle a word
        const uint32_t w = __ldg(d.bits + (lin >> 5)) >> (uint32_t)(lin & 31);
#pragma unroll
        for (int k = 0; k < VEC; ++k) m[k] = ((w >> k) & 1u) ? d.pscale : 0.f;
    } else if (d.rng) {
        const uint64_t grp = (uint64_t)lin >> 2;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)grp, (uint32_t)(grp >> 32), x.call_lo, x.call_hi), x.key);
        const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < VEC; ++k) m[k] = u[((int)(lin & 3) + k) & 3] < d.thresh ? 0.f : d.pscale;
    } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) m[k] = 1.f;
    }
}

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

// Partial sums live in a column-major table: partial[(which * c + col) * ldp + blk], which = 0 (sum / S1) or
// 1 (sum of squares / S2), so that the warp that finalises a column reads its partials coalesced whoever
// produced them (k_colsums: <= 296 blocks; the SpMM / pair-GEMM epilogues: one per CTA of their grids).
__device__ __forceinline__ void reduce_partials(const double* partial, int nblk, int ldp, int c, int col, double& s,
                                                double& q) {
    const int lane = threadIdx.x & 31;
    const double* ps = partial + (int64_t)col * ldp;
    const double* pq = partial + ((int64_t)c + col) * ldp;
    double a[4] = {0.0, 0.0, 0.0, 0.0}, b[4] = {0.0, 0.0, 0.0, 0.0};
    int i = lane;
    for (; i + 96 < nblk; i += 128) {             // four independent loads per array in flight
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a[u] += __ldcg(ps + i + 32 * u);
            b[u] += __ldcg(pq + i + 32 * u);
        }
    }
    for (; i < nblk; i += 32) {
        a[0] += __ldcg(ps + i);
        b[0] += __ldcg(pq + i);
    }
    s = warp_sum((a[0] + a[1]) + (a[2] + a[3]));
    q = warp_sum((b[0] + b[1]) + (b[2] + b[3]));
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
__device__ __forceinline__ void finalize_bwd_col(const Fin& f, double s1, double s2, int64_t n, int c, int col,
                                                 bool write_grads = true) {
    const double w = f.weight[col], a = f.mean_scale[col];
    const double rstd = f.stats[ST_RSTD * c + col], mu = f.stats[ST_MU * c + col], am = f.stats[ST_AM * c + col];
    const double N = (double)n;
    const double sum_yhat = rstd * N * (mu - am);
    const double sum_do = rstd * w * (s1 - sum_yhat * s2 / N);
    if (write_grads) {
        f.dweight[col] = (float)s2;
        f.dbias[col] = (float)s1;
        f.dmean_scale[col] = (float)(-mu * sum_do);
    }
    f.coef[0 * c + col] = (float)(rstd * w);
    f.coef[1 * c + col] = (float)(-rstd * w * s2 / N);
    f.coef[2 * c + col] = (float)(-a * sum_do / N);
}

template <int VEC, bool BWD>
__global__ void __launch_bounds__(kSumThreads)
k_colsums(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dout, int64_t lddo,
          const float* __restrict__ stats, const float* __restrict__ bias, int act, const Drop drop, int64_t n, int c,
          double* __restrict__ partial, int ldp) {
    const DropCtx dctx = BWD ? drop_ctx(drop, stats, c) : DropCtx{};
    // thread -> (column vector cvl, row lane rl).  CVB column vectors are processed per pass.
    const int CV = (c + VEC - 1) / VEC;
    const int CVB = CV < kSumThreads ? CV : kSumThreads;
    const int nrl = kSumThreads / CVB;
    const int cvl = threadIdx.x % CVB, rl = threadIdx.x / CVB;
    const bool active = rl < nrl;
    __shared__ double sm[2][kSumThreads * VEC];
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
                float dm[VEC];
                if (BWD) drop_mult<VEC>(drop, dctx, r * (int64_t)c + col, dm);
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    if (!BWD) {
                        s[k] += (double)xv[k];
                        q[k] += (double)xv[k] * (double)xv[k];
                    } else {
                        float o = xv[k] - am[k];
                        float pre = fmaf(sc[k], o, bs[k]);
                        float u = gv[k] * act_grad_from_pre(pre, act) * dm[k];
                        s[k] += (double)u;
                        q[k] += (double)u * (double)(o * rs[k]);
                    }
                }
            }
        }
        // reduce over row lanes by a fixed binary tree (deterministic; a serial loop over up to 256 row lanes
        // on one thread per column dominated the kernel for narrow matrices: 10 us at c = 8)
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            sm[0][threadIdx.x * VEC + k] = s[k];
            sm[1][threadIdx.x * VEC + k] = q[k];
        }
        int span = 1;
        while (span < nrl) span <<= 1;
        for (int st = span >> 1; st > 0; st >>= 1) {
            __syncthreads();
            if (rl < st && rl + st < nrl) {
                const int o = (threadIdx.x + st * CVB) * VEC;
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    sm[0][threadIdx.x * VEC + k] += sm[0][o + k];
                    sm[1][threadIdx.x * VEC + k] += sm[1][o + k];
                }
            }
        }
        __syncthreads();
        if (rl == 0 && cv < CV) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                if (cv * VEC + k >= c) continue;
                partial[(int64_t)(cv * VEC + k) * ldp + blockIdx.x] = sm[0][cvl * VEC + k];
                partial[((int64_t)c + cv * VEC + k) * ldp + blockIdx.x] = sm[1][cvl * VEC + k];
            }
        }
        __syncthreads();
    }
}

// Finalisation (+ dropout bits).  Warp w of the grid finalises column w (when w < c).  With `bits` the same launch
// also draws this call's keep bits: word i holds elements 32 i .. 32 i + 31 (row-major element index), bit set =
// keep, drawn exactly like the in-kernel generator (Philox counter = element index / 4, word = index % 4), so
// every later consumer -- apply, the pair-GEMM operand loaders, backward -- reads one bit instead of running
// Philox again.  The call id is rng[1]; every CTA reads it first and the last CTA to finish advances it
// (ticket in rng[2]), so a CUDA-graph replay draws fresh bits.
template <bool BWD>
__global__ void __launch_bounds__(256) k_gn_finalize(const double* partial, int nblk, int ldp, int64_t n, int c,
                                                     const Fin fin, unsigned long long* rng, uint32_t* bits,
                                                     int64_t n_words, uint32_t thresh) {
    unsigned long long id = 0, seed = 0;
    if (!BWD && rng) {
        seed = rng[0];
        id = *reinterpret_cast<volatile unsigned long long*>(rng + 1);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            fin.stats[ST_RNG * c + 0] = __uint_as_float((uint32_t)id);
            if (c > 1) fin.stats[ST_RNG * c + 1] = __uint_as_float((uint32_t)(id >> 32));
        }
    }
    const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (col < c) {
        double a, b;
        reduce_partials(partial, nblk, ldp, c, col, a, b);
        if ((threadIdx.x & 31) == 0) {
            if (BWD) finalize_bwd_col(fin, a, b, n, c, col);
            else finalize_fwd_col(fin, a, b, n, c, col);
        }
    }
    if (!BWD && rng) {
        if (bits) {
            const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
            const uint32_t call_lo = (uint32_t)id, call_hi = c > 1 ? (uint32_t)(id >> 32) : 0u;   // as drop_ctx() reads it back
            for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
                uint32_t word = 0;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const uint64_t grp = (uint64_t)w * 8 + g;
                    const uint4 r = philox4x32_10(make_uint4((uint32_t)grp, (uint32_t)(grp >> 32), call_lo, call_hi), key);
                    word |= (r.x < thresh ? 0u : 1u) << (4 * g);
                    word |= (r.y < thresh ? 0u : 1u) << (4 * g + 1);
                    word |= (r.z < thresh ? 0u : 1u) << (4 * g + 2);
                    word |= (r.w < thresh ? 0u : 1u) << (4 * g + 3);
                }
                bits[w] = word;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            unsigned* ticket = reinterpret_cast<unsigned*>(rng + 2);
            if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
                rng[1] = id + 1;
                *ticket = 0u;
            }
        }
    }
}

constexpr int kApplyCtasPerSm = 4, kBwdApplyCtasPerSm = 3;   // resident CTAs (registers); grids are one wave

// Element-wise passes.  A thread owns one column vector (VEC consecutive columns) and walks rows, so the
// per-column constants live in registers and there is no index division in the loop (the first version
// mapped a flat element index with a 64-bit div/mod and re-read the constants per element: 300 instructions
// per float4, 70 % issue-slot utilisation -- instruction bound, not memory bound).
template <int VEC>
__global__ void __launch_bounds__(kThreads, kApplyCtasPerSm)
k_gn_apply(const float* __restrict__ x, int64_t ldx, const float* __restrict__ stats, const float* __restrict__ bias,
           int act, const Drop drop, float* __restrict__ out, int64_t ldo, int64_t n, int c) {
    const DropCtx dctx = drop_ctx(drop, stats, c);
    const int CV = (c + VEC - 1) / VEC;
    const int CVB = CV < kThreads ? CV : kThreads;
    const int nrl = kThreads / CVB;
    const int cvl = threadIdx.x % CVB, rl = threadIdx.x / CVB;
    if (rl >= nrl) return;
    const int64_t row0 = (int64_t)blockIdx.x * nrl + rl, row_stride = (int64_t)gridDim.x * nrl;
    for (int cv = cvl; cv < CV; cv += CVB) {
        const int col = cv * VEC;
        float sc[VEC], am[VEC], bs[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const int cc = col + k < c ? col + k : c - 1;
            sc[k] = stats[ST_SCALE * c + cc];
            am[k] = stats[ST_AM * c + cc];
            bs[k] = bias[cc];
        }
#pragma unroll 4
        for (int64_t r = row0; r < n; r += row_stride) {
            float xv[VEC], ov[VEC];
            if (VEC == 4) {
                const float4 t = ldg_f4(x + r * ldx + col);
                xv[0] = t.x, xv[1] = t.y, xv[2] = t.z, xv[3] = t.w;
            } else {
                xv[0] = x[r * ldx + col];
            }
            float dm[VEC];
            drop_mult<VEC>(drop, dctx, r * (int64_t)c + col, dm);
#pragma unroll
            for (int k = 0; k < VEC; ++k) ov[k] = act_fwd(fmaf(sc[k], xv[k] - am[k], bs[k]), act) * dm[k];
            if (VEC == 4) *reinterpret_cast<float4*>(out + r * ldo + col) = make_float4(ov[0], ov[1], ov[2], ov[3]);
            else out[r * ldo + col] = ov[0];
        }
    }
}

// PREMASKED: `dout` already is u = dout * keep/(1-p) * act'(pre) (formed by the producer of the gradient).
template <int VEC, bool PREMASKED>
__global__ void __launch_bounds__(kThreads, kBwdApplyCtasPerSm)
k_gn_bwd_apply(const float* __restrict__ dout, int64_t lddo, const float* __restrict__ x, int64_t ldx,
               const float* __restrict__ stats, const float* __restrict__ bias, const float* __restrict__ coef, int act,
               const Drop drop, float* __restrict__ dx, int64_t lddx, int64_t n, int c) {
    const DropCtx dctx = drop_ctx(drop, stats, c);
    const int CV = (c + VEC - 1) / VEC;
    const int CVB = CV < kThreads ? CV : kThreads;
    const int nrl = kThreads / CVB;
    const int cvl = threadIdx.x % CVB, rl = threadIdx.x / CVB;
    if (rl >= nrl) return;
    const int64_t row0 = (int64_t)blockIdx.x * nrl + rl, row_stride = (int64_t)gridDim.x * nrl;
    for (int cv = cvl; cv < CV; cv += CVB) {
        const int col = cv * VEC;
        float sc[VEC], am[VEC], bs[VEC], rs[VEC], c0[VEC], c1[VEC], c2[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const int cc = col + k < c ? col + k : c - 1;
            sc[k] = stats[ST_SCALE * c + cc];
            am[k] = stats[ST_AM * c + cc];
            rs[k] = stats[ST_RSTD * c + cc];
            bs[k] = bias[cc];
            c0[k] = coef[0 * c + cc];
            c1[k] = coef[1 * c + cc];
            c2[k] = coef[2 * c + cc];
        }
#pragma unroll 4
        for (int64_t r = row0; r < n; r += row_stride) {
            float xv[VEC], gv[VEC], ov[VEC];
            if (VEC == 4) {
                const float4 t = ldg_f4(x + r * ldx + col), g = ldg_f4(dout + r * lddo + col);
                xv[0] = t.x, xv[1] = t.y, xv[2] = t.z, xv[3] = t.w;
                gv[0] = g.x, gv[1] = g.y, gv[2] = g.z, gv[3] = g.w;
            } else {
                xv[0] = x[r * ldx + col];
                gv[0] = dout[r * lddo + col];
            }
            float dm[VEC];
            if (!PREMASKED) drop_mult<VEC>(drop, dctx, r * (int64_t)c + col, dm);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float o = xv[k] - am[k];
                const float u = PREMASKED ? gv[k] : gv[k] * act_grad_from_pre(fmaf(sc[k], o, bs[k]), act) * dm[k];
                ov[k] = fmaf(c0[k], u, fmaf(c1[k], o * rs[k], c2[k]));
            }
            if (VEC == 4) *reinterpret_cast<float4*>(dx + r * lddx + col) = make_float4(ov[0], ov[1], ov[2], ov[3]);
            else dx[r * lddx + col] = ov[0];
        }
    }
}

// grid of the element-wise kernels: at least 4 rows per thread, at most one wave of resident CTAs
inline unsigned apply_ctas(int64_t n, int c, int vec, int ctas_per_sm) {
    const int cv = (c + vec - 1) / vec;
    const int cvb = cv < kThreads ? cv : kThreads;
    const int nrl = kThreads / cvb;
    int64_t want = ceil_div(n, (int64_t)nrl * 4);
    if (want < 1) want = 1;
    return (unsigned)std::min<int64_t>(want, (int64_t)sm_count() * ctas_per_sm);
}

// ---- small matrices: the whole GraphNorm (forward or backward) as ONE cluster launch -----------------------
// On tiny matrices the three-kernel pipeline above is pure launch / round-trip latency (5 + 2 + 2 us on a
// 5,000 x 8 matrix).  Here a cluster of 8 CTAs takes the column sums of its row slices (same fp64
// accumulation and tree as k_colsums), exchanges the 2c partial sums through distributed shared memory
// (added in CTA-rank order -> deterministic), every CTA finalises the per-column constants redundantly
// into its own shared memory, and the element-wise pass re-reads the rows the CTA has just read (L1/L2).
constexpr int kCl = 8, kClThreads = 512, kClMaxC = 256;

template <int VEC, bool BWD>
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kClThreads)
k_gn_cluster(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dout, int64_t lddo, const Fin fin,
             int act, const Drop drop, unsigned long long* rng, float* __restrict__ out, int64_t ldo, int64_t n, int c) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    __shared__ double sm[2][kClThreads * VEC];
    __shared__ float s_stats[6 * kClMaxC];
    __shared__ float s_coef[3 * kClMaxC];
    __shared__ unsigned long long s_id;
    const int tid = threadIdx.x;
    const int CV = (c + VEC - 1) / VEC;
    const int nrl = kClThreads / CV;                  // row lanes per CTA (host guarantees CV <= kClThreads)
    const int cvl = tid % CV, rl = tid / CV;
    const bool active = rl < nrl;
    const int col = cvl * VEC;
    const int64_t row0 = (int64_t)rank * nrl + rl, row_stride = (int64_t)kCl * nrl;

    DropCtx dctx{};
    if (BWD) {
        for (int i = tid; i < 5 * c; i += kClThreads) s_stats[i] = fin.stats[i];
        dctx = drop_ctx(drop, fin.stats, c);
        __syncthreads();
    } else if (rank == 0 && tid == 0 && drop.rng) {   // id of this call for the dropout generator
        const unsigned long long id = rng[1];
        rng[1] = id + 1;
        s_id = id;
    }

    // ---- pass 1: column sums of this CTA's rows
    double s[VEC], q[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) s[k] = q[k] = 0.0;
    float sc[VEC], am[VEC], rs[VEC], bs[VEC];
    if (BWD) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const int cc = col + k < c ? col + k : c - 1;
            sc[k] = s_stats[ST_SCALE * c + cc];
            am[k] = s_stats[ST_AM * c + cc];
            rs[k] = s_stats[ST_RSTD * c + cc];
            bs[k] = s_stats[ST_BIAS * c + cc];
        }
    }
    if (active) {
#pragma unroll 4
        for (int64_t r = row0; r < n; r += row_stride) {
            float xv[VEC], gv[VEC];
            if (VEC == 4) {
                const float4 t = ldg_f4(x + r * ldx + col);
                xv[0] = t.x, xv[1] = t.y, xv[2] = t.z, xv[3] = t.w;
                if (BWD) {
                    const float4 g = ldg_f4(dout + r * lddo + col);
                    gv[0] = g.x, gv[1] = g.y, gv[2] = g.z, gv[3] = g.w;
                }
            } else {
                xv[0] = x[r * ldx + col];
                if (BWD) gv[0] = dout[r * lddo + col];
            }
            float dm[VEC];
            if (BWD) drop_mult<VEC>(drop, dctx, r * (int64_t)c + col, dm);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                if (!BWD) {
                    s[k] += (double)xv[k];
                    q[k] += (double)xv[k] * (double)xv[k];
                } else {
                    const float o = xv[k] - am[k];
                    const float pre = fmaf(sc[k], o, bs[k]);
                    const float u = gv[k] * act_grad_from_pre(pre, act) * dm[k];
                    s[k] += (double)u;
                    q[k] += (double)u * (double)(o * rs[k]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        sm[0][tid * VEC + k] = s[k];
        sm[1][tid * VEC + k] = q[k];
    }
    int span = 1;
    while (span < nrl) span <<= 1;
    for (int st = span >> 1; st > 0; st >>= 1) {
        __syncthreads();
        if (rl < st && rl + st < nrl) {
            const int o = (tid + st * CV) * VEC;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                sm[0][tid * VEC + k] += sm[0][o + k];
                sm[1][tid * VEC + k] += sm[1][o + k];
            }
        }
    }
    cluster.sync();   // sm[.][col] of every CTA now holds its column sums (row lane 0 owns index col)

    // ---- exchange + finalise (every CTA, redundantly; rank 0 publishes)
    if (tid < c) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int rk = 0; rk < kCl; ++rk) {
            const double* remote = cluster.map_shared_rank(&sm[0][0], rk);
            a += remote[tid];
            b += remote[kClThreads * VEC + tid];
        }
        Fin f = fin;
        if (!BWD) {
            f.stats = s_stats;
            finalize_fwd_col(f, a, b, n, c, tid);
        } else {
            f.stats = s_stats;
            f.coef = s_coef;
            finalize_bwd_col(f, a, b, n, c, tid, rank == 0);
        }
    }
    if (!BWD && !drop.keep && drop.rng) {
        const unsigned long long id = *cluster.map_shared_rank(&s_id, 0);
        const unsigned long long seed = drop.rng[0];
        dctx.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        dctx.call_lo = (uint32_t)id;
        dctx.call_hi = c > 1 ? (uint32_t)(id >> 32) : 0u;
        if (rank == 0 && tid == 0) {
            fin.stats[ST_RNG * c + 0] = __uint_as_float((uint32_t)id);
            if (c > 1) fin.stats[ST_RNG * c + 1] = __uint_as_float((uint32_t)(id >> 32));
        }
    }
    cluster.sync();   // remote reads finished; s_stats / s_coef visible to the whole CTA
    if (!BWD && rank == 0)
        for (int i = tid; i < 5 * c; i += kClThreads) fin.stats[i] = s_stats[i];

    // ---- pass 2: element-wise, same (row, column) ownership as pass 1
    if (!active) return;
    float k0[VEC], k1[VEC], k2[VEC], k3[VEC], k4[VEC], k5[VEC], k6[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const int cc = col + k < c ? col + k : c - 1;
        k0[k] = s_stats[ST_SCALE * c + cc];
        k1[k] = s_stats[ST_AM * c + cc];
        k2[k] = s_stats[ST_BIAS * c + cc];
        k3[k] = s_stats[ST_RSTD * c + cc];
        k4[k] = BWD ? s_coef[0 * c + cc] : 0.f;
        k5[k] = BWD ? s_coef[1 * c + cc] : 0.f;
        k6[k] = BWD ? s_coef[2 * c + cc] : 0.f;
    }
#pragma unroll 4
    for (int64_t r = row0; r < n; r += row_stride) {
        float xv[VEC], gv[VEC], ov[VEC];
        if (VEC == 4) {
            const float4 t = ldg_f4(x + r * ldx + col);
            xv[0] = t.x, xv[1] = t.y, xv[2] = t.z, xv[3] = t.w;
            if (BWD) {
                const float4 g = ldg_f4(dout + r * lddo + col);
                gv[0] = g.x, gv[1] = g.y, gv[2] = g.z, gv[3] = g.w;
            }
        } else {
            xv[0] = x[r * ldx + col];
            if (BWD) gv[0] = dout[r * lddo + col];
        }
        float dm[VEC];
        drop_mult<VEC>(drop, dctx, r * (int64_t)c + col, dm);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const float o = xv[k] - k1[k];
            const float pre = fmaf(k0[k], o, k2[k]);
            if (!BWD) {
                ov[k] = act_fwd(pre, act) * dm[k];
            } else {
                const float u = gv[k] * act_grad_from_pre(pre, act) * dm[k];
                ov[k] = fmaf(k4[k], u, fmaf(k5[k], o * k3[k], k6[k]));
            }
        }
        if (VEC == 4) *reinterpret_cast<float4*>(out + r * ldo + col) = make_float4(ov[0], ov[1], ov[2], ov[3]);
        else out[r * ldo + col] = ov[0];
    }
}


// ---- large matrices: the whole GraphNorm (forward or backward) as ONE cooperative launch --------------------------
// The three-kernel pipeline reads its inputs twice (statistics, then the element-wise pass) and pays two extra
// launches.  57,333 x 64 floats are 99 KB per SM -- they fit in shared memory.  One persistent CTA per SM loads its
// contiguous slab of rows ONCE (coalesced float4), keeps it in shared memory while it accumulates the column sums,
// the CTAs exchange 2c fp64 partials through global memory around a grid-wide barrier (added in CTA order ->
// deterministic), every CTA finalises the per-column constants redundantly, and the element-wise pass runs out of
// shared memory.  HBM traffic: inputs once + output once.  Backward caches u = dout*keep/(1-p)*act'(pre) and x.
// Rows that do not fit the cache (wide matrices) are simply read again from L2 in the second phase.
constexpr int kCoopThreads = 1024;
constexpr int kCoopWarps = kCoopThreads / 32;

// reduction scratch (doubles): per-warp partials in phase 1, per-part totals after the grid barrier
__host__ __device__ inline int coop_red_doubles(int c) {
    const int cvn = c / 4, wcv = cvn < 32 ? cvn : 32;
    const int a = kCoopWarps * wcv * 4;
    return a > kCoopThreads ? a : kCoopThreads;
}

struct CoopArgs {
    const float* x;
    int64_t ldx;
    const float* dout;      // BWD
    int64_t lddo;
    Fin fin;
    int act;
    Drop drop;
    unsigned long long* rng;   // fwd, generator dropout: call id source (advanced by CTA 0)
    float* out;
    int64_t ldo;
    int64_t n;
    int c;
    int64_t rows_per_cta;
    int cache_rows;
    double* partial;        // [grid][2][c]
};

template <bool BWD>
__global__ void __launch_bounds__(kCoopThreads, 1) k_gn_coop(const CoopArgs A) {
    extern __shared__ __align__(16) uint8_t coop_smem[];
    cg::grid_group grid = cg::this_grid();
    const int c = A.c, cvn = c >> 2;                       // float4 column vectors per row (kCoopThreads % cvn == 0)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cv = tid % cvn, col = cv * 4;
    const int rstep = kCoopThreads / cvn;                  // rows advanced per pass of the CTA
    const int wcv = cvn < 32 ? cvn : 32;                   // distinct column vectors inside one warp
    // shared memory: [red: kCoopWarps x wcv x 4 doubles][tot: 2c doubles][cst: 8c floats][cache ...]
    double* s_red = reinterpret_cast<double*>(coop_smem);
    double* s_tot = s_red + coop_red_doubles(c);
    float* s_cst = reinterpret_cast<float*>(s_tot + 2 * c);
    float* s_cache = s_cst + 8 * c;                        // fwd: x slab; bwd: u slab then x slab (cache_rows x c each)
    const int64_t r0 = (int64_t)blockIdx.x * A.rows_per_cta;
    const int64_t r1 = min(r0 + A.rows_per_cta, A.n);
    const int cache_rows = A.cache_rows;

    DropCtx dctx{};
    unsigned long long call_id = 0;
    float sc[4], am[4], rs[4], bs[4];
    if (BWD) {
        dctx = drop_ctx(A.drop, A.fin.stats, c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sc[k] = A.fin.stats[ST_SCALE * c + col + k];
            am[k] = A.fin.stats[ST_AM * c + col + k];
            rs[k] = A.fin.stats[ST_RSTD * c + col + k];
            bs[k] = A.fin.stats[ST_BIAS * c + col + k];
        }
    } else if (A.drop.rng) {
        call_id = *reinterpret_cast<volatile unsigned long long*>(A.rng + 1);   // every CTA reads it before the barrier
    }

    // u of one float4 chunk (backward): dout * act'(pre) * keep/(1-p)
    auto make_u = [&](const float4& xv, const float4& gv, int64_t r) {
        float dm[4];
        drop_mult<4>(A.drop, dctx, r * (int64_t)c + col, dm);
        float4 u;
        u.x = gv.x * act_grad_from_pre(fmaf(sc[0], xv.x - am[0], bs[0]), A.act) * dm[0];
        u.y = gv.y * act_grad_from_pre(fmaf(sc[1], xv.y - am[1], bs[1]), A.act) * dm[1];
        u.z = gv.z * act_grad_from_pre(fmaf(sc[2], xv.z - am[2], bs[2]), A.act) * dm[2];
        u.w = gv.w * act_grad_from_pre(fmaf(sc[3], xv.w - am[3], bs[3]), A.act) * dm[3];
        return u;
    };

    // ---- phase 1: load the slab once, cache it, accumulate the column sums
    double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 2
    for (int64_t r = r0 + tid / cvn; r < r1; r += rstep) {
        const float4 xv = ldg_f4(A.x + r * A.ldx + col);
        const int lr = (int)(r - r0);
        if (!BWD) {
            if (lr < cache_rows) *reinterpret_cast<float4*>(s_cache + (int64_t)lr * c + col) = xv;
            s[0] += (double)xv.x, s[1] += (double)xv.y, s[2] += (double)xv.z, s[3] += (double)xv.w;
            q[0] += (double)xv.x * (double)xv.x, q[1] += (double)xv.y * (double)xv.y;
            q[2] += (double)xv.z * (double)xv.z, q[3] += (double)xv.w * (double)xv.w;
        } else {
            const float4 gv = ldg_f4(A.dout + r * A.lddo + col);
            const float4 u = make_u(xv, gv, r);
            if (lr < cache_rows) {
                *reinterpret_cast<float4*>(s_cache + (int64_t)lr * c + col) = u;
                *reinterpret_cast<float4*>(s_cache + ((int64_t)cache_rows + lr) * c + col) = xv;
            }
            s[0] += (double)u.x, s[1] += (double)u.y, s[2] += (double)u.z, s[3] += (double)u.w;
            q[0] += (double)u.x * (double)((xv.x - am[0]) * rs[0]), q[1] += (double)u.y * (double)((xv.y - am[1]) * rs[1]);
            q[2] += (double)u.z * (double)((xv.z - am[2]) * rs[2]), q[3] += (double)u.w * (double)((xv.w - am[3]) * rs[3]);
        }
    }
    // lanes of a warp with the same column vector, then the warps of the CTA in order (two passes: sums, squares)
#pragma unroll
    for (int which = 0; which < 2; ++which) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double t = which ? q[k] : s[k];
            for (int off = cvn; off < 32; off <<= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
            if (lane < wcv) s_red[(warp * wcv + lane) * 4 + k] = t;
        }
        __syncthreads();
        if (tid < c) {
            const int tcv = tid >> 2, k = tid & 3;
            double t = 0.0;
            if (cvn <= 32) {
                for (int w = 0; w < kCoopWarps; ++w) t += s_red[(w * wcv + tcv) * 4 + k];
            } else {            // a warp covers 32 of the cvn column vectors: warp w holds [(32 w) % cvn, +32)
                const int per = cvn >> 5;
                for (int w = (tcv >> 5); w < kCoopWarps; w += per) t += s_red[(w * 32 + (tcv & 31)) * 4 + k];
            }
            A.partial[((int64_t)blockIdx.x * 2 + which) * c + tid] = t;
        }
        __syncthreads();
    }
    __threadfence();
    grid.sync();

    // ---- totals (CTA order), per-column constants (every CTA redundantly; CTA 0 publishes).  All 1024 threads take
    // part: thread (part, j) loads the partials of CTAs part, part + parts, ... (all loads issued before the first
    // add -- one round of L2 latency instead of one per CTA), the parts are then added in order.
    {
        const int parts = min(kCoopThreads / (2 * c), 16);  // >= 4 (c <= 128): at most 37 CTAs per part
        const int j = tid % (2 * c), part = tid / (2 * c);
        const int which = j / c, cc = j - which * c;
        const int nb = (int)gridDim.x;
        double t = 0.0;
        if (part < parts) {
            const int per = (nb + parts - 1) / parts;
            for (int i0 = 0; i0 < per; i0 += 10) {           // ten independent loads per round
                double vals[10];
#pragma unroll
                for (int i = 0; i < 10; ++i) {
                    const int b = part + (i0 + i) * parts;
                    vals[i] = (i0 + i < per && b < nb) ? __ldcg(A.partial + ((int64_t)b * 2 + which) * c + cc) : 0.0;
                }
#pragma unroll
                for (int i = 0; i < 10; ++i) t += vals[i];
            }
        }
        double* s_part = s_red;                               // parts * 2c <= kCoopThreads doubles (coop_red_doubles)
        if (part < parts) s_part[part * 2 * c + j] = t;
        __syncthreads();
        if (tid < 2 * c) {
            double tot = 0.0;
            for (int pp = 0; pp < parts; ++pp) tot += s_part[pp * 2 * c + tid];
            s_tot[tid] = tot;
        }
    }
    __syncthreads();
    if (tid < c) {
        Fin f = A.fin;
        if (!BWD) {
            f.stats = s_cst;                                 // rows ST_SCALE .. ST_BIAS of a private copy
            finalize_fwd_col(f, s_tot[tid], s_tot[c + tid], A.n, c, tid);
            if (blockIdx.x == 0) {
#pragma unroll
                for (int row = 0; row < 5; ++row) A.fin.stats[row * c + tid] = s_cst[row * c + tid];
            }
        } else {
            f.coef = s_cst;
            finalize_bwd_col(f, s_tot[tid], s_tot[c + tid], A.n, c, tid, blockIdx.x == 0);
        }
    }
    if (!BWD && A.drop.rng && !A.drop.keep && !A.drop.bits) {
        const unsigned long long seed = A.drop.rng[0];
        dctx.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        dctx.call_lo = (uint32_t)call_id;
        dctx.call_hi = c > 1 ? (uint32_t)(call_id >> 32) : 0u;
        if (blockIdx.x == 0 && tid == 0) {
            A.rng[1] = call_id + 1;                          // every CTA read the old value before the grid barrier
            A.fin.stats[ST_RNG * c + 0] = __uint_as_float((uint32_t)call_id);
            if (c > 1) A.fin.stats[ST_RNG * c + 1] = __uint_as_float((uint32_t)(call_id >> 32));
        }
    }
    __syncthreads();

    // ---- phase 2: element-wise pass out of shared memory
    float k0[4], k1[4], k2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!BWD) {
            k0[k] = s_cst[ST_SCALE * c + col + k];
            k1[k] = s_cst[ST_AM * c + col + k];
            k2[k] = s_cst[ST_BIAS * c + col + k];
        } else {
            k0[k] = s_cst[0 * c + col + k];
            k1[k] = s_cst[1 * c + col + k];
            k2[k] = s_cst[2 * c + col + k];
        }
    }
#pragma unroll 2
    for (int64_t r = r0 + tid / cvn; r < r1; r += rstep) {
        const int lr = (int)(r - r0);
        float4 o;
        if (!BWD) {
            const float4 xv = lr < cache_rows ? *reinterpret_cast<const float4*>(s_cache + (int64_t)lr * c + col)
                                              : ldg_f4(A.x + r * A.ldx + col);
            float dm[4];
            drop_mult<4>(A.drop, dctx, r * (int64_t)c + col, dm);
            o.x = act_fwd(fmaf(k0[0], xv.x - k1[0], k2[0]), A.act) * dm[0];
            o.y = act_fwd(fmaf(k0[1], xv.y - k1[1], k2[1]), A.act) * dm[1];
            o.z = act_fwd(fmaf(k0[2], xv.z - k1[2], k2[2]), A.act) * dm[2];
            o.w = act_fwd(fmaf(k0[3], xv.w - k1[3], k2[3]), A.act) * dm[3];
        } else {
            float4 u, xv;
            if (lr < cache_rows) {
                u = *reinterpret_cast<const float4*>(s_cache + (int64_t)lr * c + col);
                xv = *reinterpret_cast<const float4*>(s_cache + ((int64_t)cache_rows + lr) * c + col);
            } else {
                xv = ldg_f4(A.x + r * A.ldx + col);
                u = make_u(xv, ldg_f4(A.dout + r * A.lddo + col), r);
            }
            o.x = fmaf(k0[0], u.x, fmaf(k1[0], (xv.x - am[0]) * rs[0], k2[0]));
            o.y = fmaf(k0[1], u.y, fmaf(k1[1], (xv.y - am[1]) * rs[1], k2[1]));
            o.z = fmaf(k0[2], u.z, fmaf(k1[2], (xv.z - am[2]) * rs[2], k2[2]));
            o.w = fmaf(k0[3], u.w, fmaf(k1[3], (xv.w - am[3]) * rs[3], k2[3]));
        }
        *reinterpret_cast<float4*>(A.out + r * A.ldo + col) = o;
    }
}

constexpr size_t kCoopMaxSmem = 232448;   // 227 KB opt-in limit

inline size_t coop_fixed_bytes(int c) {
    return (size_t)coop_red_doubles(c) * sizeof(double) + (size_t)2 * c * sizeof(double) + (size_t)8 * c * sizeof(float);
}

// Largest matrices first: one cooperative launch when the column layout allows it.  GLASS_B200_GN_COOP=0 disables.
inline bool use_coop(int64_t n, int c, bool vec) {
    static const int on = [] {
        const char* e = getenv("GLASS_B200_GN_COOP");
        return e ? atoi(e) : 1;
    }();
    static const int64_t min_elems = [] {
        const char* e = getenv("GLASS_B200_GN_COOP_MIN");
        return e ? (int64_t)atoll(e) : (int64_t)512 * 1024;
    }();
    if (!on || !vec || c > 128 || (kCoopThreads * 4) % c != 0) return false;
    return n * (int64_t)c >= min_elems;
}

template <bool BWD>
int launch_coop(CoopArgs& A, void* workspace, cudaStream_t st) {
    int grid = sm_count();
    const int cvn = A.c / 4;
    const int64_t min_rows = kCoopThreads / cvn;             // at least one pass of the CTA
    if ((int64_t)grid * min_rows > A.n) grid = (int)std::max<int64_t>(1, A.n / min_rows);
    A.rows_per_cta = ceil_div(A.n, grid);
    grid = (int)ceil_div(A.n, A.rows_per_cta);
    const size_t fixed = coop_fixed_bytes(A.c);
    const size_t row_bytes = (size_t)A.c * sizeof(float) * (BWD ? 2 : 1);
    int64_t cache = (int64_t)((kCoopMaxSmem - fixed) / row_bytes);
    if (cache > A.rows_per_cta) cache = A.rows_per_cta;
    A.cache_rows = (int)cache;
    A.partial = static_cast<double*>(workspace);
    const size_t smem = fixed + (size_t)cache * row_bytes;
    static bool attr_done = false;
    if (!attr_done) {
        GLASS_CUDA(cudaFuncSetAttribute(k_gn_coop<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopMaxSmem));
        attr_done = true;
    }
    void* args[] = {&A};
    GLASS_CUDA(cudaLaunchCooperativeKernel((const void*)k_gn_coop<BWD>, dim3(grid), dim3(kCoopThreads), args, smem, st));
    return GLASS_OK;
}

// Largest matrix (elements) the cluster kernel takes; GLASS_B200_GN_FUSED_MAX overrides (0 disables).
// Measured fwd+bwd pair, graph replay (scripts/gn_time.py): 10 K elements 19.9 vs 24.1 us (three kernels),
// 40 K 25.4 vs 26.8, 80 K 29.1 vs 25.2, 160 K 38.7 vs 25.7 -- eight SMs run out of issue slots (Philox, ELU,
// fp64 sums) long before the data is large, so the cut-over is low.
inline bool use_cluster(int64_t n, int c, int vec) {
    static const int64_t max_elems = [] {
        const char* e = getenv("GLASS_B200_GN_FUSED_MAX");
        return e ? (int64_t)atoll(e) : (int64_t)48 * 1024;
    }();
    return c <= kClMaxC && (c + vec - 1) / vec <= kClThreads && n * (int64_t)c <= max_elems;
}

inline Drop make_drop(const uint8_t* keep, const unsigned long long* rng, const uint32_t* bits, float p) {
    Drop d{};
    d.pscale = 1.f;
    if (p > 0.f) {
        d.pscale = 1.f / (1.f - p);
        double t = (double)p * 4294967296.0;
        d.thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
        if (keep) d.keep = keep;
        else if (bits) d.bits = bits;
        else if (rng) d.rng = rng;
    }
    return d;
}

constexpr int kPartialLd = 320;   // leading dimension of the internal partial table (k_colsums: <= 296 blocks)

inline int partial_ctas(int64_t n, int c, int vec) {
    int cv = (c + vec - 1) / vec;
    int cvb = cv < kSumThreads ? cv : kSumThreads;
    int nrl = kSumThreads / cvb;
    int64_t want = ceil_div(n, (int64_t)nrl * 4);  // >= 4 rows per thread
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

inline size_t partial_bytes(int c) { return align_up((size_t)2 * (size_t)c * kPartialLd * sizeof(double), 256); }

// finalize (forward) + this call's dropout bits in one launch
int launch_finalize_fwd(const double* partial, int nblk, int ldp, int64_t n, int c, const Fin& fin, const Drop& drop,
                        unsigned long long* rng, uint32_t* bits, cudaStream_t st) {
    const bool draw = drop.pscale != 1.f && !drop.keep && rng;          // bits (or the call id) come from the generator
    uint32_t* out_bits = (draw && drop.bits) ? bits : nullptr;
    const int64_t n_words = out_bits ? ceil_div(n * (int64_t)c, 32) : 0;
    int64_t grid = ceil_div(c, 8);
    if (out_bits) grid = std::max<int64_t>(grid, std::min<int64_t>(ceil_div(n_words, 256), (int64_t)sm_count() * 4));
    k_gn_finalize<false><<<(unsigned)grid, 256, 0, st>>>(partial, nblk, ldp, n, c, fin, draw ? rng : nullptr, out_bits,
                                                        n_words, drop.thresh);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

int launch_apply(const float* x, int64_t ldx, const float* stats, int act, const Drop& drop, float* out, int64_t ldo,
                 int64_t n, int c, cudaStream_t st) {
    const float* bias = stats + ST_BIAS * (int64_t)c;
    const bool vec = vec_ok(c, {ldx, ldo}, {x, out});
    const unsigned grid = apply_ctas(n, c, vec ? 4 : 1, kApplyCtasPerSm);
    if (vec) k_gn_apply<4><<<grid, kThreads, 0, st>>>(x, ldx, stats, bias, act, drop, out, ldo, n, c);
    else k_gn_apply<1><<<grid, kThreads, 0, st>>>(x, ldx, stats, bias, act, drop, out, ldo, n, c);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

}  // namespace
}  // namespace glass

using namespace glass;

extern "C" size_t glass_graphnorm_workspace_bytes(int64_t n, int c) {
    if (n < 0 || c <= 0) return 0;
    return partial_bytes(c) + align_up(3 * (size_t)c * sizeof(float), 256);
}

extern "C" int glass_graphnorm_launches(int64_t n, int c) {
    if (n <= 0 || c <= 0) return 0;
    if (use_cluster(n, c, c % 4 == 0 ? 4 : 1)) return 1;
    return use_coop(n, c, c % 4 == 0) ? 1 : 3;     // (unaligned operands fall back to three launches at run time)
}

extern "C" size_t glass_dropout_bits_bytes(int64_t n, int c) {
    if (n < 0 || c <= 0) return 0;
    return (size_t)ceil_div(n * (int64_t)c, 32) * sizeof(uint32_t);
}

extern "C" int glass_graphnorm_fwd(const float* x, int64_t ldx, const float* weight, const float* bias,
                                   const float* mean_scale, float eps, int act, const uint8_t* keep, float drop_p,
                                   unsigned long long* rng, uint32_t* bits, float* out, int64_t ldo, float* stats,
                                   int64_t n, int c, void* workspace, size_t workspace_bytes, void* stream) {
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
    GLASS_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "graphnorm_fwd: dropout p must be in [0, 1)");
    if (use_cluster(n, c, vec ? 4 : 1)) {   // tiny matrix: one launch, bits drawn in-kernel (the `bits` buffer stays unused)
        const Drop d = make_drop(keep, rng, nullptr, drop_p);
        if (vec) k_gn_cluster<4, false><<<kCl, kClThreads, 0, st>>>(x, ldx, nullptr, 0, fin, act, d, rng, out, ldo, n, c);
        else k_gn_cluster<1, false><<<kCl, kClThreads, 0, st>>>(x, ldx, nullptr, 0, fin, act, d, rng, out, ldo, n, c);
        GLASS_LAUNCH_CHECK();
        return GLASS_OK;
    }
    if (use_coop(n, c, vec)) {   // one cooperative launch; the bits are drawn in-kernel (the `bits` buffer stays unused)
        CoopArgs A{};
        A.x = x, A.ldx = ldx, A.fin = fin, A.act = act, A.drop = make_drop(keep, rng, nullptr, drop_p), A.rng = rng;
        A.out = out, A.ldo = ldo, A.n = n, A.c = c;
        return launch_coop<false>(A, workspace, st);
    }
    if (vec) k_colsums<4, false><<<nblk, kSumThreads, 0, st>>>(x, ldx, nullptr, 0, nullptr, nullptr, 0, Drop{}, n, c, partial, kPartialLd);
    else k_colsums<1, false><<<nblk, kSumThreads, 0, st>>>(x, ldx, nullptr, 0, nullptr, nullptr, 0, Drop{}, n, c, partial, kPartialLd);
    const Drop drop = make_drop(keep, rng, bits, drop_p);
    int rc = launch_finalize_fwd(partial, nblk, kPartialLd, n, c, fin, drop, rng, bits, st);
    if (rc != GLASS_OK) return rc;
    return launch_apply(x, ldx, stats, act, drop, out, ldo, n, c, st);
}

// Statistics produced elsewhere (the SpMM epilogue, the pair-GEMM epilogue): partial[(which * c + col) * ldp + blk],
// blk < nblk.  Writes `stats` (and draws the dropout bits of this call); the normalisation itself is applied by
// whoever consumes x next (glass_graphnorm_apply, or the operand loader of glass_pair_linear_mix_*_ex).
extern "C" int glass_graphnorm_stats(const double* partial, int nblk, int ldp, const float* weight, const float* bias,
                                     const float* mean_scale, float eps, const uint8_t* keep, float drop_p,
                                     unsigned long long* rng, uint32_t* bits, float* stats, int64_t n, int c,
                                     void* stream) {
    GLASS_CHECK_ARG(partial && nblk > 0 && ldp >= nblk && weight && bias && mean_scale && stats && n > 0 && c > 0,
                    "graphnorm_stats: bad arguments");
    GLASS_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "graphnorm_stats: dropout p must be in [0, 1)");
    GLASS_CHECK_ARG(!(drop_p > 0.f && !keep && rng && !bits), "graphnorm_stats: generator dropout needs a bits buffer");
    Fin fin{};
    fin.weight = weight, fin.bias = bias, fin.mean_scale = mean_scale, fin.eps = eps, fin.stats = stats;
    const Drop drop = make_drop(keep, rng, bits, drop_p);
    return launch_finalize_fwd(partial, nblk, ldp, n, c, fin, drop, rng, bits, as_stream(stream));
}

extern "C" int glass_graphnorm_apply(const float* x, int64_t ldx, const float* stats, int act, const uint8_t* keep,
                                     float drop_p, const uint32_t* bits, float* out, int64_t ldo, int64_t n, int c,
                                     void* stream) {
    GLASS_CHECK_ARG(x && stats && out && n > 0 && c > 0 && ldx >= c && ldo >= c, "graphnorm_apply: bad arguments");
    GLASS_CHECK_ARG(!(drop_p > 0.f && !keep && !bits), "graphnorm_apply: dropout needs a keep mask or packed bits");
    const Drop drop = make_drop(keep, nullptr, bits, drop_p);
    return launch_apply(x, ldx, stats, act, drop, out, ldo, n, c, as_stream(stream));
}

extern "C" int glass_graphnorm_bwd(const float* dout, int64_t lddo, const float* x, int64_t ldx, const float* weight,
                                   const float* mean_scale, const float* stats, int act, const uint8_t* keep,
                                   float drop_p, const unsigned long long* rng, const uint32_t* bits, float* dx,
                                   int64_t lddx, float* dweight, float* dbias, float* dmean_scale, int64_t n, int c,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    GLASS_CHECK_ARG(dout && x && weight && mean_scale && stats && dx && dweight && dbias && dmean_scale && n > 0 &&
                        c > 0 && ldx >= c && lddo >= c && lddx >= c,
                    "graphnorm_bwd: bad arguments");
    if (workspace_bytes < glass_graphnorm_workspace_bytes(n, c) || !workspace) {
        set_error("graphnorm_bwd: workspace too small");
        return GLASS_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    double* partial = static_cast<double*>(workspace);
    float* coef = reinterpret_cast<float*>(static_cast<char*>(workspace) + partial_bytes(c));
    const float* bias = stats + ST_BIAS * (int64_t)c;  // forward bias saved with the statistics
    const bool vec = vec_ok(c, {ldx, lddo, lddx}, {x, dout, dx});
    const int nblk = partial_ctas(n, c, vec ? 4 : 1);
    Fin fin{};
    fin.weight = weight, fin.mean_scale = mean_scale, fin.stats = const_cast<float*>(stats), fin.coef = coef;
    fin.dweight = dweight, fin.dbias = dbias, fin.dmean_scale = dmean_scale;
    if (use_cluster(n, c, vec ? 4 : 1)) {
        const Drop drop = make_drop(keep, rng, nullptr, drop_p);
        if (vec) k_gn_cluster<4, true><<<kCl, kClThreads, 0, st>>>(x, ldx, dout, lddo, fin, act, drop, nullptr, dx, lddx, n, c);
        else k_gn_cluster<1, true><<<kCl, kClThreads, 0, st>>>(x, ldx, dout, lddo, fin, act, drop, nullptr, dx, lddx, n, c);
        GLASS_LAUNCH_CHECK();
        return GLASS_OK;
    }
    if (use_coop(n, c, vec)) {
        CoopArgs A{};
        A.x = x, A.ldx = ldx, A.dout = dout, A.lddo = lddo, A.fin = fin, A.act = act;
        A.drop = make_drop(keep, rng, nullptr, drop_p);
        A.out = dx, A.ldo = lddx, A.n = n, A.c = c;
        return launch_coop<true>(A, workspace, st);
    }
    const Drop drop = make_drop(keep, rng, bits, drop_p);
    if (vec) k_colsums<4, true><<<nblk, kSumThreads, 0, st>>>(x, ldx, dout, lddo, stats, bias, act, drop, n, c, partial, kPartialLd);
    else k_colsums<1, true><<<nblk, kSumThreads, 0, st>>>(x, ldx, dout, lddo, stats, bias, act, drop, n, c, partial, kPartialLd);
    k_gn_finalize<true><<<(unsigned)ceil_div(c, 8), 256, 0, st>>>(partial, nblk, kPartialLd, n, c, fin, nullptr, nullptr, 0, 0u);
    const unsigned grid = apply_ctas(n, c, vec ? 4 : 1, kBwdApplyCtasPerSm);
    if (vec) k_gn_bwd_apply<4, false><<<grid, kThreads, 0, st>>>(dout, lddo, x, ldx, stats, bias, coef, act, drop, dx, lddx, n, c);
    else k_gn_bwd_apply<1, false><<<grid, kThreads, 0, st>>>(dout, lddo, x, ldx, stats, bias, coef, act, drop, dx, lddx, n, c);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

// Backward when the producer of the gradient (the dX epilogue of the pair GEMM) has already formed
// u = dout * keep/(1-p) * act'(pre) and the per-CTA partial sums of S1 = sum u, S2 = sum u * yhat:
// finalise + one element-wise pass dx = alpha u + beta yhat + gamma.  `u` and `dx` may alias.
extern "C" int glass_graphnorm_bwd_from_sums(const double* partial, int nblk, int ldp, const float* u, int64_t ldu,
                                             const float* x, int64_t ldx, const float* weight, const float* mean_scale,
                                             const float* stats, float* dx, int64_t lddx, float* dweight, float* dbias,
                                             float* dmean_scale, int64_t n, int c, void* workspace,
                                             size_t workspace_bytes, void* stream) {
    GLASS_CHECK_ARG(partial && nblk > 0 && ldp >= nblk && u && x && weight && mean_scale && stats && dx && dweight &&
                        dbias && dmean_scale && n > 0 && c > 0 && ldu >= c && ldx >= c && lddx >= c,
                    "graphnorm_bwd_from_sums: bad arguments");
    if (workspace_bytes < align_up(3 * (size_t)c * sizeof(float), 256) || !workspace) {
        set_error("graphnorm_bwd_from_sums: workspace too small");
        return GLASS_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    float* coef = static_cast<float*>(workspace);
    const float* bias = stats + ST_BIAS * (int64_t)c;
    Fin fin{};
    fin.weight = weight, fin.mean_scale = mean_scale, fin.stats = const_cast<float*>(stats), fin.coef = coef;
    fin.dweight = dweight, fin.dbias = dbias, fin.dmean_scale = dmean_scale;
    k_gn_finalize<true><<<(unsigned)ceil_div(c, 8), 256, 0, st>>>(partial, nblk, ldp, n, c, fin, nullptr, nullptr, 0, 0u);
    const bool vec = vec_ok(c, {ldx, ldu, lddx}, {x, u, dx});
    const unsigned grid = apply_ctas(n, c, vec ? 4 : 1, kBwdApplyCtasPerSm);
    const Drop none{};
    if (vec) k_gn_bwd_apply<4, true><<<grid, kThreads, 0, st>>>(u, ldu, x, ldx, stats, bias, coef, GLASS_ACT_NONE, none, dx, lddx, n, c);
    else k_gn_bwd_apply<1, true><<<grid, kThreads, 0, st>>>(u, ldu, x, ldx, stats, bias, coef, GLASS_ACT_NONE, none, dx, lddx, n, c);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

// ---- two-phase forms for row-partitioned graphs ("GraphNorm column stats need an H-float
// all-reduce per norm"): a rank computes the partial column sums of ITS rows, the caller adds them across ranks
// (one 2c-value fp64 all-reduce) and hands the totals back as a one-block partial table with the GLOBAL row count.
//   forward : glass_graphnorm_partials(x)            -> all-reduce -> glass_graphnorm_stats + glass_graphnorm_apply
//   backward: glass_graphnorm_bwd_partials(dout, x)  -> all-reduce -> glass_graphnorm_bwd_finish
// partial layout: partial[(which * c + col) * ldp + blk]; *nblk_host receives the number of blocks written.
extern "C" int glass_graphnorm_partials_ld(void) { return kPartialLd; }

extern "C" int glass_graphnorm_partials(const float* x, int64_t ldx, int64_t n, int c, double* partial, int ldp,
                                        int* nblk_host, void* stream) {
    GLASS_CHECK_ARG(x && partial && nblk_host && n > 0 && c > 0 && ldx >= c && ldp >= kMaxPartialCtas,
                    "graphnorm_partials: bad arguments");
    const bool vec = vec_ok(c, {ldx}, {x});
    const int nblk = partial_ctas(n, c, vec ? 4 : 1);
    cudaStream_t st = as_stream(stream);
    if (vec) k_colsums<4, false><<<nblk, kSumThreads, 0, st>>>(x, ldx, nullptr, 0, nullptr, nullptr, 0, Drop{}, n, c, partial, ldp);
    else k_colsums<1, false><<<nblk, kSumThreads, 0, st>>>(x, ldx, nullptr, 0, nullptr, nullptr, 0, Drop{}, n, c, partial, ldp);
    GLASS_LAUNCH_CHECK();
    *nblk_host = nblk;
    return GLASS_OK;
}

extern "C" int glass_graphnorm_bwd_partials(const float* dout, int64_t lddo, const float* x, int64_t ldx,
                                            const float* stats, int act, const uint8_t* keep, float drop_p,
                                            const uint32_t* bits, int64_t n, int c, double* partial, int ldp,
                                            int* nblk_host, void* stream) {
    GLASS_CHECK_ARG(dout && x && stats && partial && nblk_host && n > 0 && c > 0 && ldx >= c && lddo >= c &&
                        ldp >= kMaxPartialCtas,
                    "graphnorm_bwd_partials: bad arguments");
    GLASS_CHECK_ARG(!(drop_p > 0.f && !keep && !bits), "graphnorm_bwd_partials: dropout needs a keep mask or packed bits");
    const float* bias = stats + ST_BIAS * (int64_t)c;
    const bool vec = vec_ok(c, {ldx, lddo}, {x, dout});
    const int nblk = partial_ctas(n, c, vec ? 4 : 1);
    const Drop drop = make_drop(keep, nullptr, bits, drop_p);
    cudaStream_t st = as_stream(stream);
    if (vec) k_colsums<4, true><<<nblk, kSumThreads, 0, st>>>(x, ldx, dout, lddo, stats, bias, act, drop, n, c, partial, ldp);
    else k_colsums<1, true><<<nblk, kSumThreads, 0, st>>>(x, ldx, dout, lddo, stats, bias, act, drop, n, c, partial, ldp);
    GLASS_LAUNCH_CHECK();
    *nblk_host = nblk;
    return GLASS_OK;
}

// n_total: the GLOBAL row count the totals in `partial` refer to; n: rows of dout / x / dx held here.
extern "C" int glass_graphnorm_bwd_finish(const double* partial, int nblk, int ldp, int64_t n_total, const float* dout,
                                          int64_t lddo, const float* x, int64_t ldx, const float* weight,
                                          const float* mean_scale, const float* stats, int act, const uint8_t* keep,
                                          float drop_p, const uint32_t* bits, float* dx, int64_t lddx, float* dweight,
                                          float* dbias, float* dmean_scale, int64_t n, int c, void* workspace,
                                          size_t workspace_bytes, void* stream) {
    GLASS_CHECK_ARG(partial && nblk > 0 && ldp >= nblk && dout && x && weight && mean_scale && stats && dx && dweight &&
                        dbias && dmean_scale && n > 0 && n_total >= n && c > 0 && ldx >= c && lddo >= c && lddx >= c,
                    "graphnorm_bwd_finish: bad arguments");
    if (workspace_bytes < align_up(3 * (size_t)c * sizeof(float), 256) || !workspace) {
        set_error("graphnorm_bwd_finish: workspace too small");
        return GLASS_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    float* coef = static_cast<float*>(workspace);
    const float* bias = stats + ST_BIAS * (int64_t)c;
    Fin fin{};
    fin.weight = weight, fin.mean_scale = mean_scale, fin.stats = const_cast<float*>(stats), fin.coef = coef;
    fin.dweight = dweight, fin.dbias = dbias, fin.dmean_scale = dmean_scale;
    k_gn_finalize<true><<<(unsigned)ceil_div(c, 8), 256, 0, st>>>(partial, nblk, ldp, n_total, c, fin, nullptr, nullptr, 0, 0u);
    const bool vec = vec_ok(c, {ldx, lddo, lddx}, {x, dout, dx});
    const Drop drop = make_drop(keep, nullptr, bits, drop_p);
    const unsigned grid = apply_ctas(n, c, vec ? 4 : 1, kBwdApplyCtasPerSm);
    if (vec) k_gn_bwd_apply<4, false><<<grid, kThreads, 0, st>>>(dout, lddo, x, ldx, stats, bias, coef, act, drop, dx, lddx, n, c);
    else k_gn_bwd_apply<1, false><<<grid, kThreads, 0, st>>>(dout, lddo, x, ldx, stats, bias, coef, act, drop, dx, lddx, n, c);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}
