// Padded-subgraph pooling: GLASS.Pool (reference impl/models.py:346-350) = pad2batch + emb[pos] gather +
// Add/Mean/Max/Size pool (impl/models.py:295-319, PyG global_*_pool / GraphSizeNorm).
// One CTA per subgraph walks its padded row (skipping -1) so neither the (batch, pos) vectors nor the
// gathered [n_valid, D] matrix are ever materialised.  Rows are accumulated in pad order = the order of
// the reference's index_add, so sums match a sequential CPU loop.
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

template <class RowFn>
__device__ __forceinline__ void pool_fwd_body(RowFn row_of, int64_t len, const float* __restrict__ emb, int64_t lde,
                                              int mode, float* __restrict__ out, float* __restrict__ cnt_out,
                                              int32_t* __restrict__ argmax, int d, int64_t b) {
    __shared__ float s_cnt;
    if (threadIdx.x == 0) {
        int c = 0;
        for (int64_t l = 0; l < len; ++l) c += row_of(l) >= 0;
        s_cnt = (float)c;
        if (cnt_out) cnt_out[b] = (float)c;
    }
    __syncthreads();
    const float cnt = s_cnt;
    const float coef = (mode == GLASS_POOL_SIZE && cnt > 0.f) ? size_coef(cnt) : 1.f;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        float acc = (mode == GLASS_POOL_MAX) ? -FLT_MAX : 0.f;
        int32_t arg = -1;
        for (int64_t l = 0; l < len; ++l) {
            const int64_t r = row_of(l);
            if (r < 0) continue;
            const float v = emb[r * lde + c];
            if (mode == GLASS_POOL_MAX) {
                if (arg < 0 || v > acc) {
                    acc = v;
                    arg = (int32_t)(mode == GLASS_POOL_MAX ? r : 0);
                }
            } else if (mode == GLASS_POOL_SIZE) {
                acc = __fadd_rn(acc, __fmul_rn(v, coef));
            } else {
                acc += v;
            }
        }
        if (mode == GLASS_POOL_MEAN) acc = acc / fmaxf(cnt, 1.f);
        if (mode == GLASS_POOL_MAX) {
            if (arg < 0) acc = 0.f;  // empty segment -> 0 (PyG scatter semantics)
            if (argmax) argmax[b * (int64_t)d + c] = arg;
        }
        out[c] = acc;
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
    const int64_t b = blockIdx.x;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const float g = dout[b * lddo + c];
        if (mode == GLASS_POOL_MAX) {
            const int32_t a = argmax[b * (int64_t)d + c];
            if (a >= 0) atomicAdd(demb + (int64_t)a * ldde + c, g);
            continue;
        }
        const float gc = g * bwd_coef(mode, cnt[b]);
        for (int64_t l = 0; l < seg.lmax; ++l) {
            const int64_t r = seg.row(b, l);
            if (r >= 0) atomicAdd(demb + r * ldde + c, gc);  // nodes may belong to several subgraphs
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

inline int pool_threads(int d) { return d <= 32 ? 32 : (d >= 256 ? 256 : (d + 31) / 32 * 32); }
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
    k_pool_pad_fwd<<<(unsigned)b, pool_threads(d), 0, as_stream(stream)>>>(emb, lde, seg, mode, out, ldo, cnt, argmax, d);
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
    k_pool_pad_bwd<<<(unsigned)b, pool_threads(d), 0, as_stream(stream)>>>(dout, lddo, seg, mode, cnt, argmax, demb, ldde, d);
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
    k_pool_batch_fwd<<<(unsigned)n_seg, pool_threads(d), 0, as_stream(stream)>>>(x, ldx, seg, mode, out, ldo, cnt, argmax, d);
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
