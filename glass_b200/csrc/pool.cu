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

template <class RowFn>
__device__ __forceinline__ void pool_fwd_body(RowFn row_of, int64_t len, const float* __restrict__ emb, int64_t lde,
                                              int mode, float* __restrict__ out, float* __restrict__ cnt_out,
                                              int32_t* __restrict__ argmax, int d, int64_t b) {
    __shared__ int s_cnt;
    __shared__ float s_val[kPoolMaxThreads];
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
        float acc = (mode == GLASS_POOL_MAX) ? -FLT_MAX : 0.f;
        int best = -1;
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
                    const float v = emb[(int64_t)r * lde + c];
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
        s_pos[ty * CT + tx] = best;
        __syncthreads();
        if (ty == 0 && c < d) {
            for (int q = 1; q < LQ; ++q) {
                const float v = s_val[q * CT + tx];
                const int p = s_pos[q * CT + tx];
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
        }
        __syncthreads();
    }
}

__global__ void k_pool_pad_fwd(const float* __restrict__ emb, int64_t lde, PadSeg seg, int mode,
                               float* __restrict__ out, int64_t ldo, float* __restrict__ cnt, int32_t* __restrict__ argmax,
                               int d) {
    const int64_t b = blockIdx.x;
    pool_fwd_body([&](int64_t l) { return seg.row(b, l); }, seg.lmax, emb, lde, mode, out + b * ldo, cnt, argmax, d, b);
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

extern "C" int glass_segment_pool_bwd(const float* dout, int64_t lddo, const int64_t* pos, int64_t b, int64_t lmax,
                                      int mode, const float* cnt, const int32_t* argmax, float* demb, int64_t ldde,
                                      int d, int64_t n_node, void* stream) {
    if (!mode_ok(mode)) {
        set_error("segment_pool: unknown pool mode %d", mode);
        return GLASS_ERR_UNSUPPORTED;
    }
    GLASS_CHECK_ARG(dout && pos && cnt && demb && b >= 0 && lmax >= 0 && d > 0 && lddo >= d && ldde >= d && n_node > 0,
                    "segment_pool_bwd: bad arguments");
    GLASS_CHECK_ARG(mode != GLASS_POOL_MAX || argmax, "segment_pool_bwd: MAX needs argmax");
    if (b == 0) return GLASS_OK;
    PadSeg seg{pos, lmax, n_node};
    k_pool_pad_bwd<<<(unsigned)b, pool_block(d), 0, as_stream(stream)>>>(dout, lddo, seg, mode, cnt, argmax, demb, ldde, d);
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
