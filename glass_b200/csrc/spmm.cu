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
template <int G, int VEC, int KCH, int kUnroll, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_spmm(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                   const float* __restrict__ val, const float* __restrict__ x,
                                                   int64_t ldx, float* __restrict__ y, int64_t ldy, int64_t n_rows,
                                                   int h) {
    using V = Vec<VEC>;
    const int lane = threadIdx.x & 31;
    const int l = lane & (G - 1);                      // lane inside the group
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    const int64_t groups_per_grid = (int64_t)gridDim.x * (kThreads / G);
    int64_t row = (int64_t)blockIdx.x * (kThreads / G) + threadIdx.x / G;

    bool colok[KCH];
    int coff[KCH];
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
        coff[k] = (l + k * G) * VEC;
        colok[k] = coff[k] < h;
    }

    for (; row < n_rows; row += groups_per_grid) {
        const int32_t e_begin = rowptr[row], e_end = rowptr[row + 1];
        typename V::T acc[KCH];
#pragma unroll
        for (int k = 0; k < KCH; ++k) acc[k] = V::zero();

        for (int32_t e0 = e_begin; e0 < e_end; e0 += G) {
            const int cnt = min(G, e_end - e0);
            int cj = 0;
            float vj = 0.f;
            if (l < cnt) {
                cj = __ldg(col + e0 + l);
                vj = __ldg(val + e0 + l);
            }
            for (int j = 0; j < cnt; j += kUnroll) {
                int c[kUnroll];
                float v[kUnroll];
                typename V::T xv[kUnroll][KCH];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    c[u] = __shfl_sync(gmask, cj, j + u, G);
                    v[u] = __shfl_sync(gmask, vj, j + u, G);
                }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    if (j + u < cnt) {
                        const float* xr = x + (int64_t)c[u] * ldx;
#pragma unroll
                        for (int k = 0; k < KCH; ++k)
                            if (colok[k]) xv[u][k] = V::load(xr + coff[k]);
                    }
                }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    if (j + u < cnt) {
#pragma unroll
                        for (int k = 0; k < KCH; ++k)
                            if (colok[k]) V::fma(acc[k], v[u], xv[u][k]);
                    }
                }
            }
        }
        float* yr = y + row * ldy;
#pragma unroll
        for (int k = 0; k < KCH; ++k)
            if (colok[k]) V::store(yr + coff[k], acc[k]);
    }
}

template <int G, int VEC, int KCH, int U = 4, int MINB = 2>
int launch(const int32_t* rowptr, const int32_t* col, const float* val, const float* x, int64_t ldx, float* y,
           int64_t ldy, int64_t n_rows, int h, cudaStream_t st) {
    const int64_t groups_per_block = kThreads / G;
    int64_t blocks = ceil_div(n_rows, groups_per_block);
    const int64_t cap = (int64_t)sm_count() * 8 * 4;  // a few waves of resident CTAs; rows are interleaved
    if (blocks > cap) blocks = cap;
    k_spmm<G, VEC, KCH, U, MINB><<<(unsigned)blocks, kThreads, 0, st>>>(rowptr, col, val, x, ldx, y, ldy, n_rows, h);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

}  // namespace
}  // namespace glass

using namespace glass;

extern "C" int glass_spmm_csr(const int32_t* rowptr, const int32_t* col, const float* val, const float* x,
                              int64_t ldx, float* y, int64_t ldy, int64_t n_rows, int h, void* stream) {
    GLASS_CHECK_ARG(rowptr && x && y && n_rows >= 0 && h > 0 && ldx >= h && ldy >= h, "spmm_csr: bad arguments");
    GLASS_CHECK_ARG(h <= 256, "spmm_csr: h=%d > 256 not supported", h);
    if (n_rows == 0) return GLASS_OK;
    cudaStream_t st = as_stream(stream);
    const bool vec = (h % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && ((uintptr_t)x % 16 == 0) &&
                     ((uintptr_t)y % 16 == 0);
#define GO(G, V, K) return launch<G, V, K>(rowptr, col, val, x, ldx, y, ldy, n_rows, h, st)
    if (vec) {
        const int lanes = h / 4;
        if (lanes <= 2) GO(2, 4, 1);
        if (lanes <= 4) GO(4, 4, 1);
        if (lanes <= 8) GO(8, 4, 1);
        if (lanes <= 16) {
            static const int variant = getenv("GLASS_SPMM_VARIANT") ? atoi(getenv("GLASS_SPMM_VARIANT")) : 0;
            switch (variant) {   // tuning knob: gathers in flight per lane x resident CTAs per SM
                case 1: return launch<16, 4, 1, 8, 4>(rowptr, col, val, x, ldx, y, ldy, n_rows, h, st);
                case 2: return launch<16, 4, 1, 8, 3>(rowptr, col, val, x, ldx, y, ldy, n_rows, h, st);
                case 3: return launch<16, 4, 1, 4, 5>(rowptr, col, val, x, ldx, y, ldy, n_rows, h, st);
                case 4: return launch<16, 4, 1, 6, 4>(rowptr, col, val, x, ldx, y, ldy, n_rows, h, st);
                case 5: return launch<16, 4, 1, 2, 8>(rowptr, col, val, x, ldx, y, ldy, n_rows, h, st);
                case 6: return launch<16, 4, 1, 16, 2>(rowptr, col, val, x, ldx, y, ldy, n_rows, h, st);
                default: return launch<16, 4, 1, 4, 4>(rowptr, col, val, x, ldx, y, ldy, n_rows, h, st);
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
