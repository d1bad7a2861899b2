// Padded-subgraph pooling: GLASS.Pool (reference impl/models.py:346-350) = pad2batch + emb[pos] gather +
// Add/Mean/Max/Size pool (impl/models.py:295-319, PyG global_*_pool / GraphSizeNorm).
// One CTA per subgraph walks its padded row (skipping -1) so neither the (batch, pos) vectors nor the
// gathered [n_valid, D] matrix are ever materialised.  The CTA is 2-D (columns x row lanes): each row lane
// sums every LQ-th row in pad order, the lane partials are combined in lane order (deterministic).
// Algorithmic bytes: 8*B*Lmax (ids) + 4*D*n_valid (gathered rows) + 4*B*D (output).
#include <float.h>

#include "common.cuh"

namespace glass {
namespace {

// SizePool scales every row by count^-1/2 BEFORE the sum (GraphSizeNorm then add, models.py:318-319).
__device__ __forceinline__ float size_coef(float cnt) { return __fdiv_rn(1.0f, __fsqrt_rn(cnt)); }

// Segment description shared by the padded and the (x, batch) variants.
struct PadSeg {
    const int64_t* pos;
    int64_t lmax;
    int64_t n_node;
    __device__ __forceinline__ int64_t len(int64_t) const { return lmax; }
    __device__ __forceinline__ int64_t row(int64_t b, int64_t l) const {
        int64_t v = pos[b * lmax + l];
        return (v >= 0 && v < n_node) ? v : -1;
    }
};

struct BatchSeg {  // rows already gathered; batch sorted ascending -> segment = [lower_bound(b), lower_bound(b+1))
    const int64_t* batch;
    int64_t m;
    __device__ __forceinline__ int64_t lower(int64_t key) const {
        int64_t lo = 0, hi = m;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (batch[mid] < key) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    }
};

// CTA = CT column threads x LQ row lanes (CT*LQ <= 1024).  Row lane q accumulates rows q, q+LQ, ... of the
// segment (independent loads, ~len/LQ deep instead of len deep), then lane 0 combines the LQ partials in
// lane order.  MAX keeps (value, position) and resolves ties towards the earliest position.
constexpr int kPoolMaxThreads = 1024;

constexpr int kPoolStage = 2048;   // segment entries staged in shared memory per pass

// Optional GraphNorm on gather (the model's LAST GraphNorm feeds nothing but the pooling, impl/models.py:266 -> :348):
// v = fmaf(scale, x - am, bias) exactly as the apply kernel computes it, so the normalised [N, D] matrix is never
// written; ysum receives sum_v (x_v - am) * rstd per subgraph and column (the backward pass needs sum_v yhat_v).
struct PoolNorm {
    const float* stats;   // [6, d] (ST_* rows) or NULL: plain pooling
    float* ysum;          // [B, d]
};

template <class RowFn>
__device__ __forceinline__ void pool_fwd_body(RowFn row_of, int64_t len, const float* __restrict__ emb, int64_t lde,
                                              int mode, float* __restrict__ out, float* __restrict__ cnt_out,
                                              int32_t* __restrict__ argmax, int d, int64_t b, const PoolNorm norm = PoolNorm{}) {
    __shared__ int s_cnt;
    __shared__ float s_val[kPoolMaxThreads];
    __shared__ float s_val2[kPoolMaxThreads];
    __shared__ int s_pos[kPoolMaxThreads];
    __shared__ int s_row[kPoolStage];          // row id (or -1) of every staged segment entry
    const int CT = blockDim.x, LQ = blockDim.y, tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * CT + tx, nthr = CT * LQ;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    {   // valid entries of the whole segment
        int c = 0;
        for (int64_t l = tid; l < len; l += nthr) c += row_of(l) >= 0;
        if (c) atomicAdd(&s_cnt, c);
    }
    __syncthreads();
    const float cnt = (float)s_cnt;
    if (tid == 0 && cnt_out) cnt_out[b] = cnt;
    const float coef = (mode == GLASS_POOL_SIZE && cnt > 0.f) ? size_coef(cnt) : 1.f;
    const int ncol_pass = (d + CT - 1) / CT;
    for (int cp = 0; cp < ncol_pass; ++cp) {
        const int c = cp * CT + tx;
        float acc = (mode == GLASS_POOL_MAX) ? -FLT_MAX : 0.f, acc2 = 0.f;
        int best = -1;
        const bool nrm = norm.stats != nullptr;
        float n_sc = 1.f, n_am = 0.f, n_bs = 0.f, n_rs = 1.f;
        if (nrm && c < d) {
            n_sc = norm.stats[ST_SCALE * d + c], n_am = norm.stats[ST_AM * d + c];
            n_bs = norm.stats[ST_BIAS * d + c], n_rs = norm.stats[ST_RSTD * d + c];
        }
        for (int64_t l0 = 0; l0 < len; l0 += kPoolStage) {
            const int chunk = (int)min((int64_t)kPoolStage, len - l0);
            __syncthreads();
            for (int i = tid; i < chunk; i += nthr) s_row[i] = (int)row_of(l0 + i);   // the id loads are independent
            __syncthreads();
            if (c < d) {
                // row lane ty takes entries ty, ty+LQ, ...; four gathers in flight per thread
                int i = ty;
                for (; i + 3 * LQ < chunk; i += 4 * LQ) {
                    int r[4];
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) r[u] = s_row[i + u * LQ];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = r[u] >= 0 ? emb[(int64_t)r[u] * lde + c] : 0.f;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (r[u] < 0) continue;
                        if (nrm) {
                            acc2 += (v[u] - n_am) * n_rs;
                            v[u] = fmaf(n_sc, v[u] - n_am, n_bs);
                        }
                        if (mode == GLASS_POOL_MAX) {
                            if (best < 0 || v[u] > acc) {
                                acc = v[u];
                                best = (int)(l0 + i + u * LQ);
                            }
                        } else if (mode == GLASS_POOL_SIZE) {
                            acc = __fadd_rn(acc, __fmul_rn(v[u], coef));
                        } else {
                            acc += v[u];
                        }
                    }
                }
                for (; i < chunk; i += LQ) {
                    const int r = s_row[i];
                    if (r < 0) continue;
                    float v = emb[(int64_t)r * lde + c];
                    if (nrm) {
                        acc2 += (v - n_am) * n_rs;
                        v = fmaf(n_sc, v - n_am, n_bs);
                    }
                    if (mode == GLASS_POOL_MAX) {
                        if (best < 0 || v > acc) {
                            acc = v;
                            best = (int)(l0 + i);
                        }
                    } else if (mode == GLASS_POOL_SIZE) {
                        acc = __fadd_rn(acc, __fmul_rn(v, coef));
                    } else {
                        acc += v;
                    }
                }
            }
        }
        s_val[ty * CT + tx] = acc;
        s_val2[ty * CT + tx] = acc2;
        s_pos[ty * CT + tx] = best;
        __syncthreads();
        if (ty == 0 && c < d) {
            for (int q = 1; q < LQ; ++q) {
                const float v = s_val[q * CT + tx];
                const int p = s_pos[q * CT + tx];
                acc2 += s_val2[q * CT + tx];
                if (mode == GLASS_POOL_MAX) {
                    if (p >= 0 && (best < 0 || v > acc || (v == acc && p < best))) {
                        acc = v;
                        best = p;
                    }
                } else {
                    acc += v;
                }
            }
            if (mode == GLASS_POOL_MEAN) acc = acc / fmaxf(cnt, 1.f);
            if (mode == GLASS_POOL_MAX) {
                if (best < 0) acc = 0.f;  // empty segment -> 0 (PyG scatter semantics)
                if (argmax) argmax[b * (int64_t)d + c] = best < 0 ? -1 : (int32_t)row_of(best);
            }
            out[c] = acc;
            if (nrm && norm.ysum) norm.ysum[b * (int64_t)d + c] = acc2;
        }
        __syncthreads();
    }
}

__global__ void k_pool_pad_fwd(const float* __restrict__ emb, int64_t lde, PadSeg seg, int mode,
                               float* __restrict__ out, int64_t ldo, float* __restrict__ cnt, int32_t* __restrict__ argmax,
                               int d, const PoolNorm norm = PoolNorm{}) {
    const int64_t b = blockIdx.x;
    pool_fwd_body([&](int64_t l) { return seg.row(b, l); }, seg.lmax, emb, lde, mode, out + b * ldo, cnt, argmax, d, b, norm);
}

__global__ void k_pool_batch_fwd(const float* __restrict__ x, int64_t ldx, BatchSeg seg, int mode,
                                 float* __restrict__ out, int64_t ldo, float* __restrict__ cnt,
                                 int32_t* __restrict__ argmax, int d) {
    const int64_t b = blockIdx.x;
    const int64_t lo = seg.lower(b), hi = seg.lower(b + 1);
    pool_fwd_body([&](int64_t l) { return lo + l; }, hi - lo, x, ldx, mode, out + b * ldo, cnt, argmax, d, b);
}

__device__ __forceinline__ float bwd_coef(int mode, float cnt) {
    if (mode == GLASS_POOL_MEAN) return 1.f / fmaxf(cnt, 1.f);
    if (mode == GLASS_POOL_SIZE) return cnt > 0.f ? size_coef(cnt) : 0.f;
    return 1.f;
}

__global__ void k_pool_pad_bwd(const float* __restrict__ dout, int64_t lddo, PadSeg seg, int mode,
                               const float* __restrict__ cnt, const int32_t* __restrict__ argmax,
                               float* __restrict__ demb, int64_t ldde, int d) {
    __shared__ int s_row[kPoolStage];
    const int64_t b = blockIdx.x;
    const int CT = blockDim.x, LQ = blockDim.y;
    const int tid = threadIdx.y * CT + threadIdx.x, nthr = CT * LQ;
    if (mode == GLASS_POOL_MAX) {
        if (threadIdx.y == 0)
            for (int c = threadIdx.x; c < d; c += CT) {
                const int32_t a = argmax[b * (int64_t)d + c];
                if (a >= 0) atomicAdd(demb + (int64_t)a * ldde + c, dout[b * lddo + c]);
            }
        return;
    }
    const float coef = bwd_coef(mode, cnt[b]);
    for (int64_t l0 = 0; l0 < seg.lmax; l0 += kPoolStage) {
        const int chunk = (int)min((int64_t)kPoolStage, seg.lmax - l0);
        __syncthreads();
        for (int i = tid; i < chunk; i += nthr) s_row[i] = (int)seg.row(b, l0 + i);
        __syncthreads();
        for (int c = threadIdx.x; c < d; c += CT) {
            const float gc = dout[b * lddo + c] * coef;
            for (int i = threadIdx.y; i < chunk; i += LQ) {
                const int r = s_row[i];
                if (r >= 0) atomicAdd(demb + (int64_t)r * ldde + c, gc);  // nodes may belong to several subgraphs
            }
        }
    }
}

// Deterministic backward of the padded pooling (run-to-run bit-reproducible, SURVEY.md section 5 "Determinism").
// Node-centric: a warp owns kOrdRows consecutive rows of demb and WRITES all of them (zeros for nodes outside the batch --
// the caller's zero-fill disappears).  k_mark_nodes builds, per step, a byte mark per node and one membership BITMAP
// per subgraph (b x ceil(N/32) words); for a marked node the warp reads the node's bit of every subgraph (lane b
// reads bitmap b) and adds coef_b * dout[b, :] for the subgraphs that list it -- in ascending b, one thread per
// column, no atomics on floats.
__global__ void k_mark_nodes(const int64_t* __restrict__ pos, int64_t b_cnt, int64_t lmax, uint8_t* __restrict__ mark,
                             uint32_t* __restrict__ memb, int64_t words, int64_t n_node, int32_t* __restrict__ dup) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= b_cnt * lmax) return;
    const int64_t v = pos[i];
    if (v >= 0 && v < n_node) {
        mark[v] = 1;
        const uint32_t bit = 1u << (v & 31);
        const uint32_t old = atomicOr(memb + (i / lmax) * words + (v >> 5), bit);      // integer OR: order independent
        if (old & bit) dup[i / lmax] = 1;       // this subgraph lists a node more than once (EdgeGNN pairs (u, u))
    }
}

constexpr int kOrdWarps = 8;
constexpr int kOrdRows = 8;     // rows of the gradient a warp owns.  32 rows per warp left 12 warps per SM on the em_user
                                // shape (1,792 warps in all) for a kernel that is a chain of dependent loads per marked row

// GraphNorm backward fused into the ordered pooling backward (NORM): when the pooled matrix was the output of a
// GraphNorm that feeds nothing else (PoolNorm above), the gradient with respect to the norm's INPUT is
//     dx[v, :] = alpha * u[v, :] + beta * yhat[v, :] + gamma,      u[v, :] = sum_{b lists v} coef_b * dout[b, :]
// with u zero outside the ~1 K labelled rows.  The column sums the norm's backward needs are sums over subgraphs:
//     S1 = sum_v u = sum_b cnt_b coef_b dout_b,      S2 = sum_v u * yhat = sum_b coef_b dout_b * ysum_b
// (ysum_b = sum_{v in b} yhat_v from the forward kernel), so no pass over the N x D matrices is needed for them:
// every CTA forms them redundantly in fp64 in ascending b (B x D values) and derives alpha / beta / gamma exactly as
// graphnorm.cu's finalize_bwd_col; CTA 0 also writes the parameter gradients.
struct NormBwd {
    const float* x;          // [n_node, d] input of the norm
    int64_t ldx;
    const float* stats;      // [6, d]
    const float* weight;
    const float* mean_scale;
    const float* ysum;       // [B, d]
    float* dweight;
    float* dbias;
    float* dmean_scale;
};

template <bool NORM>
__global__ void __launch_bounds__(kOrdWarps * 32)
k_pool_pad_bwd_ordered(const float* __restrict__ dout, int64_t lddo, const int64_t* __restrict__ pos, int64_t lmax,
                       int64_t b_cnt, int mode,
                       const float* __restrict__ cnt, const int32_t* __restrict__ argmax,
                       const uint8_t* __restrict__ mark, const uint32_t* __restrict__ memb, int64_t words,
                       const int32_t* __restrict__ dup, float* __restrict__ demb, int64_t ldde, int d, int64_t n_node,
                       const NormBwd nb) {
    __shared__ __align__(16) float s_c[NORM ? 5 * 256 : 1];   // alpha | beta | gamma | am | rstd   (d <= 256)
    float* s_alpha = s_c;
    float* s_beta = s_c + (NORM ? 256 : 0);
    float* s_gamma = s_c + (NORM ? 512 : 0);
    float* s_am = s_c + (NORM ? 768 : 0);
    float* s_rstd = s_c + (NORM ? 1024 : 0);
    if (NORM) {
        for (int c = threadIdx.x; c < d; c += blockDim.x) {
            double s1 = 0.0, s2 = 0.0;
            for (int64_t b = 0; b < b_cnt; ++b) {
                const float g = dout[b * lddo + c] * bwd_coef(mode, cnt[b]);
                s1 += (double)g * (double)cnt[b];
                s2 += (double)g * (double)nb.ysum[b * (int64_t)d + c];
            }
            const double w = nb.weight[c], a = nb.mean_scale[c];
            const double rstd = nb.stats[ST_RSTD * d + c], mu = nb.stats[ST_MU * d + c], am = nb.stats[ST_AM * d + c];
            const double N = (double)n_node;
            const double sum_yhat = rstd * N * (mu - am);
            const double sum_do = rstd * w * (s1 - sum_yhat * s2 / N);
            if (blockIdx.x == 0) {
                nb.dweight[c] = (float)s2;
                nb.dbias[c] = (float)s1;
                nb.dmean_scale[c] = (float)(-mu * sum_do);
            }
            s_alpha[c] = (float)(rstd * w);
            s_beta[c] = (float)(-rstd * w * s2 / N);
            s_gamma[c] = (float)(-a * sum_do / N);
            s_am[c] = (float)am;
            s_rstd[c] = (float)rstd;
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t base = ((int64_t)blockIdx.x * kOrdWarps + (threadIdx.x >> 5)) * kOrdRows;
    if (base >= n_node) return;
    const int64_t mine = base + lane;
    const unsigned labelled = __ballot_sync(0xffffffffu, lane < kOrdRows && mine < n_node && mark[mine] != 0);
    const int rows = (int)min((int64_t)kOrdRows, n_node - base);
    // rows of nodes outside the batch (zeros, or beta * yhat + gamma with NORM) as one coalesced sweep over the warp's
    // row block when possible
    const bool wide = (d % 4 == 0) && (ldde % 4 == 0) && ((uintptr_t)demb % 16 == 0) &&
                      (!NORM || (nb.ldx % 4 == 0 && (uintptr_t)nb.x % 16 == 0));
    if (wide) {
        const int cv = d >> 2;                                 // float4 per row
        constexpr int U = NORM ? 4 : 1;                        // NORM: four row-chunk loads in flight per lane
        for (int q0 = lane; q0 < rows * cv; q0 += 32 * U) {
            float4 xv[U];
            int jj[U], cc[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = q0 + 32 * u;
                jj[u] = -1;
                if (q < rows * cv) {
                    const int j = q / cv;
                    if (!((labelled >> j) & 1u)) {
                        jj[u] = j;
                        cc[u] = (q - j * cv) * 4;
                        if (NORM) xv[u] = __ldg(reinterpret_cast<const float4*>(nb.x + (base + j) * nb.ldx + cc[u]));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (jj[u] < 0) continue;
                const int c4 = cc[u];
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (NORM) {
                    const float4 am = *reinterpret_cast<const float4*>(s_am + c4), rs = *reinterpret_cast<const float4*>(s_rstd + c4);
                    const float4 be = *reinterpret_cast<const float4*>(s_beta + c4), ga = *reinterpret_cast<const float4*>(s_gamma + c4);
                    o.x = fmaf(be.x, (xv[u].x - am.x) * rs.x, ga.x), o.y = fmaf(be.y, (xv[u].y - am.y) * rs.y, ga.y);
                    o.z = fmaf(be.z, (xv[u].z - am.z) * rs.z, ga.z), o.w = fmaf(be.w, (xv[u].w - am.w) * rs.w, ga.w);
                }
                *reinterpret_cast<float4*>(demb + (base + jj[u]) * ldde + c4) = o;
            }
        }
    }
    for (int j = 0; j < rows; ++j) {
        const int64_t v = base + j;
        float* row = demb + v * ldde;
        if (!((labelled >> j) & 1u)) {
            if (!wide)
                for (int c = lane; c < d; c += 32)
                    row[c] = NORM ? fmaf(s_beta[c], (nb.x[v * nb.ldx + c] - s_am[c]) * s_rstd[c], s_gamma[c]) : 0.f;
            continue;
        }
        for (int c0 = 0; c0 < d; c0 += 128) {                 // 4 columns per lane and pass
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int64_t b0 = 0; b0 < b_cnt; b0 += 32) {       // 32 subgraphs per step: lane -> subgraph b0 + lane
                const int64_t bl = b0 + lane;
                const bool in = bl < b_cnt && ((__ldg(memb + bl * words + (v >> 5)) >> (v & 31)) & 1u);
                unsigned hits = __ballot_sync(0xffffffffu, in);
                while (hits) {
                    const int64_t b = b0 + (__ffs(hits) - 1);
                    hits &= hits - 1;
                    if (mode == GLASS_POOL_MAX) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int c = c0 + lane + 32 * k;
                            if (c < d && argmax[b * (int64_t)d + c] == (int32_t)v) acc[k] += dout[b * lddo + c];
                        }
                    } else {
                        // a padded row may list a node more than once (EdgeGNN pairs (u, u)): count the occurrences -- only
                        // in subgraphs that k_mark_nodes flagged (a scan of the padded row per (node, subgraph) otherwise
                        // costs 15 dependent loads at Lmax = 473: 19 us instead of 8 on the em_user batch)
                        int mult = 1;
                        if (dup[b]) {
                            mult = 0;
                            for (int64_t l0 = 0; l0 < lmax; l0 += 32)
                                mult += __popc(__ballot_sync(0xffffffffu, l0 + lane < lmax && __ldg(pos + b * lmax + l0 + lane) == v));
                        }
                        const float coef = bwd_coef(mode, cnt[b]);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int c = c0 + lane + 32 * k;
                            if (c < d) {
                                const float gc = dout[b * lddo + c] * coef;
                                for (int m = 0; m < mult; ++m) acc[k] += gc;      // same value the per-entry sum adds
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = c0 + lane + 32 * k;
                if (c < d) {
                    if (NORM) {
                        const float yhat = (nb.x[v * nb.ldx + c] - s_am[c]) * s_rstd[c];
                        row[c] = fmaf(s_alpha[c], acc[k], fmaf(s_beta[c], yhat, s_gamma[c]));
                    } else {
                        row[c] = acc[k];
                    }
                }
            }
        }
    }
}

// (x, batch) variant: every gathered row belongs to exactly one segment -> plain stores, one thread per element.
__global__ void k_pool_batch_bwd(const float* __restrict__ dout, int64_t lddo, const int64_t* __restrict__ batch,
                                 int64_t m, int mode, const float* __restrict__ cnt, const int32_t* __restrict__ argmax,
                                 float* __restrict__ dx, int64_t lddx, int d) {
    const int64_t total = m * d, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t r = e / d;
        const int c = (int)(e % d);
        const int64_t b = batch[r];
        float g = dout[b * lddo + c];
        if (mode == GLASS_POOL_MAX) g = (argmax[b * (int64_t)d + c] == (int32_t)r) ? g : 0.f;
        else g *= bwd_coef(mode, cnt[b]);
        dx[r * lddx + c] = g;
    }
}

inline dim3 pool_block(int d) {
    int ct = d <= 32 ? 32 : (d >= 128 ? 128 : (d + 31) / 32 * 32);
    return dim3(ct, kPoolMaxThreads / ct >= 1 ? (kPoolMaxThreads / ct > 32 ? 32 : kPoolMaxThreads / ct) : 1);
}
inline bool mode_ok(int mode) { return mode >= GLASS_POOL_SUM && mode <= GLASS_POOL_SIZE; }

}  // namespace
}  // namespace glass

using namespace glass;

extern "C" int glass_segment_pool_fwd(const float* emb, int64_t lde, const int64_t* pos, int64_t b, int64_t lmax,
                                      int mode, float* out, int64_t ldo, float* cnt, int32_t* argmax, int d,
                                      int64_t n_node, void* stream) {
    if (!mode_ok(mode)) {
        set_error("segment_pool: unknown pool mode %d (reference raises NotImplementedError, GLASSTest.py:171)", mode);
        return GLASS_ERR_UNSUPPORTED;
    }
    GLASS_CHECK_ARG(emb && pos && out && cnt && b >= 0 && lmax >= 0 && d > 0 && lde >= d && ldo >= d && n_node > 0,
                    "segment_pool_fwd: bad arguments");
    GLASS_CHECK_ARG(mode != GLASS_POOL_MAX || argmax, "segment_pool_fwd: MAX needs argmax");
    if (b == 0) return GLASS_OK;
    PadSeg seg{pos, lmax, n_node};
    k_pool_pad_fwd<<<(unsigned)b, pool_block(d), 0, as_stream(stream)>>>(emb, lde, seg, mode, out, ldo, cnt, argmax, d);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

// scratch of the ordered backward: b membership bitmaps + one byte mark per node
extern "C" size_t glass_segment_pool_bwd_scratch_bytes(int64_t b, int64_t n_node) {
    if (b < 0 || n_node <= 0) return 0;
    return align_up((size_t)b * (size_t)ceil_div(n_node, 32) * sizeof(uint32_t) + (size_t)b * sizeof(int32_t) + (size_t)n_node, 256);
}

extern "C" int glass_segment_pool_bwd(const float* dout, int64_t lddo, const int64_t* pos, int64_t b, int64_t lmax,
                                      int mode, const float* cnt, const int32_t* argmax, float* demb, int64_t ldde,
                                      int d, int64_t n_node, void* scratch, size_t scratch_bytes, void* stream) {
    if (!mode_ok(mode)) {
        set_error("segment_pool: unknown pool mode %d", mode);
        return GLASS_ERR_UNSUPPORTED;
    }
    GLASS_CHECK_ARG(dout && pos && cnt && demb && b >= 0 && lmax >= 0 && d > 0 && lddo >= d && ldde >= d && n_node > 0,
                    "segment_pool_bwd: bad arguments");
    GLASS_CHECK_ARG(mode != GLASS_POOL_MAX || argmax, "segment_pool_bwd: MAX needs argmax");
    if (scratch) {   // ordered (deterministic) variant: writes EVERY row of demb, no zero-fill needed
        const size_t need = glass_segment_pool_bwd_scratch_bytes(b, n_node);
        if (scratch_bytes < need) {
            set_error("segment_pool_bwd: scratch %zu < required %zu", scratch_bytes, need);
            return GLASS_ERR_WORKSPACE;
        }
        cudaStream_t st = as_stream(stream);
        const int64_t words = ceil_div(n_node, 32);
        uint32_t* memb = static_cast<uint32_t*>(scratch);
        int32_t* dup = reinterpret_cast<int32_t*>(memb + (size_t)b * words);
        uint8_t* mark = reinterpret_cast<uint8_t*>(dup + b);
        GLASS_CUDA(cudaMemsetAsync(scratch, 0, need, st));
        const int64_t n_pos = b * lmax;
        if (n_pos > 0) k_mark_nodes<<<(unsigned)ceil_div(n_pos, 256), 256, 0, st>>>(pos, b, lmax, mark, memb, words, n_node, dup);
        k_pool_pad_bwd_ordered<false><<<(unsigned)ceil_div(n_node, kOrdWarps * kOrdRows), kOrdWarps * 32, 0, st>>>(
            dout, lddo, pos, lmax, b, mode, cnt, argmax, mark, memb, words, dup, demb, ldde, d, n_node, NormBwd{});
        GLASS_LAUNCH_CHECK();
        return GLASS_OK;
    }
    if (b == 0) return GLASS_OK;
    PadSeg seg{pos, lmax, n_node};
    k_pool_pad_bwd<<<(unsigned)b, pool_block(d), 0, as_stream(stream)>>>(dout, lddo, seg, mode, cnt, argmax, demb, ldde, d);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

// ---- last GraphNorm + pooling as one operator (impl/models.py:266 / :272 -> :346-350) ------------------------------
// forward: out[b, :] = pool_v( scale * (x[v, :] - am) + bias ) with the statistics table of glass_graphnorm_stats;
// the normalised [n_node, d] matrix is never written.  ysum [B, d] is saved for backward.  sum / mean / size only.
extern "C" int glass_norm_pool_fwd(const float* x, int64_t ldx, const float* stats, const int64_t* pos, int64_t b,
                                   int64_t lmax, int mode, float* out, int64_t ldo, float* cnt, float* ysum, int d,
                                   int64_t n_node, void* stream) {
    GLASS_CHECK_ARG(mode == GLASS_POOL_SUM || mode == GLASS_POOL_MEAN || mode == GLASS_POOL_SIZE,
                    "norm_pool: pool mode %d not supported (sum / mean / size)", mode);
    GLASS_CHECK_ARG(x && stats && pos && out && cnt && ysum && b >= 0 && lmax >= 0 && d > 0 && d <= 256 && ldx >= d &&
                        ldo >= d && n_node > 0,
                    "norm_pool_fwd: bad arguments");
    if (b == 0) return GLASS_OK;
    PadSeg seg{pos, lmax, n_node};
    k_pool_pad_fwd<<<(unsigned)b, pool_block(d), 0, as_stream(stream)>>>(x, ldx, seg, mode, out, ldo, cnt, nullptr, d,
                                                                        PoolNorm{stats, ysum});
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

// backward: dx (gradient of the norm's input, EVERY row written), dweight / dbias / dmean_scale of the norm.
// scratch as for glass_segment_pool_bwd (glass_segment_pool_bwd_scratch_bytes).
extern "C" int glass_norm_pool_bwd(const float* dout, int64_t lddo, const int64_t* pos, int64_t b, int64_t lmax, int mode,
                                   const float* cnt, const float* ysum, const float* x, int64_t ldx, const float* stats,
                                   const float* weight, const float* mean_scale, float* dx, int64_t lddx, float* dweight,
                                   float* dbias, float* dmean_scale, int d, int64_t n_node, void* scratch,
                                   size_t scratch_bytes, void* stream) {
    GLASS_CHECK_ARG(mode == GLASS_POOL_SUM || mode == GLASS_POOL_MEAN || mode == GLASS_POOL_SIZE,
                    "norm_pool: pool mode %d not supported (sum / mean / size)", mode);
    GLASS_CHECK_ARG(dout && pos && cnt && ysum && x && stats && weight && mean_scale && dx && dweight && dbias && dmean_scale &&
                        b >= 0 && lmax >= 0 && d > 0 && d <= 256 && lddo >= d && ldx >= d && lddx >= d && n_node > 0 && scratch,
                    "norm_pool_bwd: bad arguments");
    const size_t need = glass_segment_pool_bwd_scratch_bytes(b, n_node);
    if (scratch_bytes < need) {
        set_error("norm_pool_bwd: scratch %zu < required %zu", scratch_bytes, need);
        return GLASS_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    const int64_t words = ceil_div(n_node, 32);
    uint32_t* memb = static_cast<uint32_t*>(scratch);
    int32_t* dup = reinterpret_cast<int32_t*>(memb + (size_t)b * words);
    uint8_t* mark = reinterpret_cast<uint8_t*>(dup + b);
    GLASS_CUDA(cudaMemsetAsync(scratch, 0, need, st));
    const int64_t n_pos = b * lmax;
    if (n_pos > 0) k_mark_nodes<<<(unsigned)ceil_div(n_pos, 256), 256, 0, st>>>(pos, b, lmax, mark, memb, words, n_node, dup);
    NormBwd nb{x, ldx, stats, weight, mean_scale, ysum, dweight, dbias, dmean_scale};
    k_pool_pad_bwd_ordered<true><<<(unsigned)ceil_div(n_node, kOrdWarps * kOrdRows), kOrdWarps * 32, 0, st>>>(
        dout, lddo, pos, lmax, b, mode, cnt, nullptr, mark, memb, words, dup, dx, lddx, d, n_node, nb);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

extern "C" int glass_segment_pool_batch_fwd(const float* x, int64_t ldx, const int64_t* batch, int64_t m, int64_t n_seg,
                                            int mode, float* out, int64_t ldo, float* cnt, int32_t* argmax, int d,
                                            void* stream) {
    if (!mode_ok(mode)) {
        set_error("segment_pool: unknown pool mode %d", mode);
        return GLASS_ERR_UNSUPPORTED;
    }
    GLASS_CHECK_ARG(x && batch && out && cnt && m >= 0 && n_seg >= 0 && d > 0 && ldx >= d && ldo >= d,
                    "segment_pool_batch_fwd: bad arguments");
    GLASS_CHECK_ARG(mode != GLASS_POOL_MAX || argmax, "segment_pool_batch_fwd: MAX needs argmax");
    if (n_seg == 0) return GLASS_OK;
    BatchSeg seg{batch, m};
    k_pool_batch_fwd<<<(unsigned)n_seg, pool_block(d), 0, as_stream(stream)>>>(x, ldx, seg, mode, out, ldo, cnt, argmax, d);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

extern "C" int glass_segment_pool_batch_bwd(const float* dout, int64_t lddo, const int64_t* batch, int64_t m,
                                            int64_t n_seg, int mode, const float* cnt, const int32_t* argmax, float* dx,
                                            int64_t lddx, int d, void* stream) {
    if (!mode_ok(mode)) {
        set_error("segment_pool: unknown pool mode %d", mode);
        return GLASS_ERR_UNSUPPORTED;
    }
    GLASS_CHECK_ARG(dout && batch && cnt && dx && m >= 0 && n_seg >= 0 && d > 0 && lddo >= d && lddx >= d,
                    "segment_pool_batch_bwd: bad arguments");
    GLASS_CHECK_ARG(mode != GLASS_POOL_MAX || argmax, "segment_pool_batch_bwd: MAX needs argmax");
    if (m == 0) return GLASS_OK;
    const int64_t work = m * d;
    unsigned grid = (unsigned)std::min<int64_t>(ceil_div(work, 256), (int64_t)sm_count() * 8);
    k_pool_batch_bwd<<<grid, 256, 0, as_stream(stream)>>>(dout, lddo, batch, m, mode, cnt, argmax, dx, lddx, d);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}
