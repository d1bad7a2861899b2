// spmm_csr: y = A x for the normalised adjacency (reference impl/models.py:164, `self.adj @ x`;
// backward runs the same kernel on the transposed CSR).
//
// Mapping: a sub-warp GROUP of G lanes owns one CSR row; lane l of the group owns VEC consecutive
// feature columns (VEC = 4 -> one 16-byte gather per neighbour and lane, a whole 256-byte feature row
// per 16 lanes at H = 64).  The group streams its (col, val) entries G at a time with one coalesced
// load per array, broadcasts them with width-G shuffles, and issues UNROLL independent float4 gathers
// before the dependent FMA chain.  Accumulation is a single fp32 chain per column in CSR order, so the
// result is deterministic and matches a sequential CPU loop over the sorted entries.
//
// Roofline: compulsory HBM bytes = 4(N+1) + 8 nnz + 8 N H.  The gathers
// (4 H nnz bytes) are served by L1/L2: X (14.7 MB at the em_user shape) is L2-resident.
#include <stdlib.h>

#include "common.cuh"

namespace glass {
namespace {

constexpr int kThreads = 256;

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
    using T = float4;
    static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ T load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ void store(float* p, const T& v) { *reinterpret_cast<float4*>(p) = v; }
    static __device__ __forceinline__ void fma(T& a, float s, const T& x) {
        a.x = fmaf(s, x.x, a.x);
        a.y = fmaf(s, x.y, a.y);
        a.z = fmaf(s, x.z, a.z);
        a.w = fmaf(s, x.w, a.w);
    }
};
template <>
struct Vec<1> {
    using T = float;
    static __device__ __forceinline__ T zero() { return 0.f; }
    static __device__ __forceinline__ T load(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void store(float* p, const T& v) { *p = v; }
    static __device__ __forceinline__ void fma(T& a, float s, const T& x) { a = fmaf(s, x, a); }
};

// G lanes per row, VEC floats per lane and chunk, KCH column chunks per lane (h <= G*VEC*KCH).
// EXACT: h == G*VEC*KCH (no column guards).  IDX32: every element offset into x fits 32 bits.
template <int G, int VEC, int KCH, bool EXACT, bool IDX32, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_spmm(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                         const float* __restrict__ val, const float* __restrict__ x,
                                                         int64_t ldx, float* __restrict__ y, int64_t ldy, int64_t n_rows,
                                                         int h) {
    using V = Vec<VEC>;
    // (col, val) staging: one 8-byte slot per lane, double buffered so that one __syncwarp per chunk suffices
    __shared__ int2 s_e[2][kThreads];
    const int lane = threadIdx.x & 31;
    const int l = lane & (G - 1);                      // lane inside the group
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    const int gbase = threadIdx.x & ~(G - 1);
    const int64_t groups_per_grid = (int64_t)gridDim.x * (kThreads / G);
    int64_t row = (int64_t)blockIdx.x * (kThreads / G) + threadIdx.x / G;
    const uint32_t ldx32 = (uint32_t)ldx;

    bool colok[KCH];
    int coff[KCH];
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
        coff[k] = (l + k * G) * VEC;
        colok[k] = EXACT || coff[k] < h;
    }

    for (; row < n_rows; row += groups_per_grid) {
        const int32_t e_begin = __ldg(rowptr + row), e_end = __ldg(rowptr + row + 1);
        typename V::T acc[KCH];
#pragma unroll
        for (int k = 0; k < KCH; ++k) acc[k] = V::zero();
        int buf = 0;
        // software pipeline: the (col, val) pair of the NEXT chunk is in flight while this chunk is gathered
        int nc = 0;
        float nv = 0.f;
        if (e_begin + l < e_end) {
            nc = __ldg(col + e_begin + l);
            nv = __ldg(val + e_begin + l);
        }
        for (int32_t e0 = e_begin; e0 < e_end; e0 += G, buf ^= 1) {
            const int cnt = min(G, e_end - e0);
            s_e[buf][threadIdx.x] = make_int2(nc, __float_as_int(nv));
            if (e0 + G + l < e_end) {
                nc = __ldg(col + e0 + G + l);
                nv = __ldg(val + e0 + G + l);
            }
            __syncwarp(gmask);
            const int2* se = &s_e[buf][gbase];
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const int2 cv = se[j];                                  // broadcast read inside the group
                const float* xr = IDX32 ? x + (uint32_t)cv.x * ldx32 : x + (int64_t)cv.x * ldx;
                const float w = __int_as_float(cv.y);
#pragma unroll
                for (int k = 0; k < KCH; ++k)
                    if (colok[k]) V::fma(acc[k], w, V::load(xr + coff[k]));
            }
        }
        float* yr = y + row * ldy;
#pragma unroll
        for (int k = 0; k < KCH; ++k)
            if (colok[k]) V::store(yr + coff[k], acc[k]);
    }
}

template <int G, int VEC, int KCH, int MINB = 4>
int launch(const int32_t* rowptr, const int32_t* col, const float* val, const float* x, int64_t ldx, float* y,
           int64_t ldy, int64_t n_rows, int64_t n_cols, int h, cudaStream_t st) {
    const int64_t groups_per_block = kThreads / G;
    int64_t blocks = ceil_div(n_rows, groups_per_block);
    const int64_t cap = (int64_t)sm_count() * 8 * 4;  // a few waves of resident CTAs; rows are interleaved
    if (blocks > cap) blocks = cap;
    const bool exact = h == G * VEC * KCH;
    const bool idx32 = n_cols * ldx < (1ll << 31);
    const unsigned grid = (unsigned)blocks;
#define GLASS_SPMM_GO(E, I) k_spmm<G, VEC, KCH, E, I, MINB><<<grid, kThreads, 0, st>>>(rowptr, col, val, x, ldx, y, ldy, n_rows, h)
    if (exact && idx32) GLASS_SPMM_GO(true, true);
    else if (exact) GLASS_SPMM_GO(true, false);
    else if (idx32) GLASS_SPMM_GO(false, true);
    else GLASS_SPMM_GO(false, false);
#undef GLASS_SPMM_GO
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

}  // namespace
}  // namespace glass

using namespace glass;

extern "C" int glass_spmm_csr(const int32_t* rowptr, const int32_t* col, const float* val, const float* x,
                              int64_t ldx, float* y, int64_t ldy, int64_t n_rows, int64_t n_cols, int h, void* stream) {
    GLASS_CHECK_ARG(rowptr && x && y && n_rows >= 0 && n_cols > 0 && h > 0 && ldx >= h && ldy >= h,
                    "spmm_csr: bad arguments");
    GLASS_CHECK_ARG(h <= 256, "spmm_csr: h=%d > 256 not supported", h);
    if (n_rows == 0) return GLASS_OK;
    cudaStream_t st = as_stream(stream);
    const bool vec = (h % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && ((uintptr_t)x % 16 == 0) &&
                     ((uintptr_t)y % 16 == 0);
#define GO(G, V, K) return launch<G, V, K>(rowptr, col, val, x, ldx, y, ldy, n_rows, n_cols, h, st)
    if (vec) {
        const int lanes = h / 4;
        if (lanes <= 2) GO(2, 4, 1);
        if (lanes <= 4) GO(4, 4, 1);
        if (lanes <= 8) GO(8, 4, 1);
        if (lanes <= 16) {
            static const int variant = getenv("GLASS_SPMM_VARIANT") ? atoi(getenv("GLASS_SPMM_VARIANT")) : 0;
            switch (variant) {   // tuning knob: resident CTAs per SM (register cap)
                case 1: return launch<16, 4, 1, 3>(rowptr, col, val, x, ldx, y, ldy, n_rows, n_cols, h, st);
                case 2: return launch<16, 4, 1, 6>(rowptr, col, val, x, ldx, y, ldy, n_rows, n_cols, h, st);
                case 3: return launch<16, 4, 1, 8>(rowptr, col, val, x, ldx, y, ldy, n_rows, n_cols, h, st);
                default: return launch<16, 4, 1, 5>(rowptr, col, val, x, ldx, y, ldy, n_rows, n_cols, h, st);
            }
        }
        if (lanes <= 32) GO(32, 4, 1);
        GO(32, 4, 2);
    } else {
        if (h <= 4) GO(4, 1, 1);
        if (h <= 8) GO(8, 1, 1);
        if (h <= 16) GO(16, 1, 1);
        if (h <= 32) GO(32, 1, 1);
        if (h <= 64) GO(32, 1, 2);
        if (h <= 128) GO(32, 1, 4);
        GO(32, 1, 8);
    }
#undef GO
}
