// Measurement aid (no product path calls it): what would a 24-bit gather copy of x buy the L2-bound SpMM?
// glass_l2_gather_probe24 reads `gathers` pseudo-random rows exactly as glass_l2_gather_probe does (same lane layout,
// same index stream, 8 rows in flight per lane group), but every row is split in two planes: hi = the top 16 bits of each
// fp32 (2 bytes per feature) and lo = the next 8 mantissa bits (1 byte per feature), i.e. 3 instead of 4 bytes per
// feature and two loads per neighbour and lane instead of one.  DESIGN.md section 7.
#include "common.cuh"

namespace glass {
namespace {

constexpr int kThreads = 256;

__global__ void k_split24(const float* __restrict__ x, int64_t n_elems, uint16_t* __restrict__ hi, uint8_t* __restrict__ lo) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_elems) return;
    const uint32_t u = __float_as_uint(x[i]) + 0x80u;          // round to 24 bits (carry into the exponent is fine)
    hi[i] = (uint16_t)(u >> 16);
    lo[i] = (uint8_t)(u >> 8);
}

template <int G>
__global__ void __launch_bounds__(kThreads, 5) k_l2_gather_probe24(const uint16_t* __restrict__ hi, const uint8_t* __restrict__ lo,
                                                                 uint32_t h, uint32_t n_rows, int64_t per_group,
                                                                 float* __restrict__ sink) {
    const int lane = threadIdx.x & 31, l = lane & (G - 1);
    const int64_t group = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / G;
    uint32_t s = (uint32_t)group * 2654435761u + 12345u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = 0; i < per_group; i += 8) {
        uint2 vh[8];
        uint32_t vl[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t r = (uint32_t)(((uint64_t)s * n_rows) >> 32);
            vh[u] = __ldg(reinterpret_cast<const uint2*>(hi + (size_t)r * h) + l);
            vl[u] = __ldg(reinterpret_cast<const uint32_t*>(lo + (size_t)r * h) + l);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            acc.x += __uint_as_float((vh[u].x << 16) | ((vl[u] & 0xffu) << 8));
            acc.y += __uint_as_float((vh[u].x & 0xffff0000u) | ((vl[u] & 0xff00u)));
            acc.z += __uint_as_float((vh[u].y << 16) | ((vl[u] >> 8) & 0xff00u));
            acc.w += __uint_as_float((vh[u].y & 0xffff0000u) | ((vl[u] >> 16) & 0xff00u));
        }
    }
    sink[(int64_t)blockIdx.x * kThreads + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

// fp32 rows gathered with 256-bit loads: G lanes x 32 bytes per row (h = 8 G), half the load instructions of the
// float4 layout for the same bytes.
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}

template <int G, int U>
__global__ void __launch_bounds__(kThreads, 4) k_l2_gather_probe256(const float* __restrict__ x, uint32_t ldx, uint32_t n_rows,
                                                                  int64_t per_group, float* __restrict__ sink) {
    const int lane = threadIdx.x & 31, l = lane & (G - 1);
    const int64_t group = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / G;
    uint32_t s = (uint32_t)group * 2654435761u + 12345u;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t i = 0; i < per_group; i += U) {
        float v[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t r = (uint32_t)(((uint64_t)s * n_rows) >> 32);
            ldg256(x + (size_t)r * ldx + 8 * l, v[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += v[u][k];
    }
    sink[(int64_t)blockIdx.x * kThreads + threadIdx.x] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
}

}  // namespace
}  // namespace glass

using namespace glass;

// hi: uint16 [n_rows, h], lo: uint8 [n_rows, h] (filled here from x), sink: >= sm_count * 5 * 256 floats
extern "C" int glass_l2_gather_probe24(const float* x, int64_t n_rows, int h, int64_t gathers, uint16_t* hi, uint8_t* lo,
                                       float* sink, int64_t sink_elems, int convert, void* stream) {
    GLASS_CHECK_ARG(x && hi && lo && sink && n_rows > 0 && gathers > 0 && (h == 32 || h == 64 || h == 128) &&
                        n_rows * (int64_t)h < (1ll << 31) && (uintptr_t)hi % 8 == 0 && (uintptr_t)lo % 4 == 0,
                    "l2_gather_probe24: bad arguments");
    const int g = h / 4;
    const int64_t grid = (int64_t)sm_count() * 5;
    GLASS_CHECK_ARG(sink_elems >= grid * kThreads, "l2_gather_probe24: sink needs %lld floats", (long long)(grid * kThreads));
    cudaStream_t st = as_stream(stream);
    if (convert) {
        const int64_t n = n_rows * h;
        k_split24<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(x, n, hi, lo);
    }
    const int64_t groups = grid * kThreads / g;
    const int64_t per_group = (ceil_div(gathers, groups) + 7) / 8 * 8;
    if (g == 8) k_l2_gather_probe24<8><<<(unsigned)grid, kThreads, 0, st>>>(hi, lo, (uint32_t)h, (uint32_t)n_rows, per_group, sink);
    else if (g == 16) k_l2_gather_probe24<16><<<(unsigned)grid, kThreads, 0, st>>>(hi, lo, (uint32_t)h, (uint32_t)n_rows, per_group, sink);
    else k_l2_gather_probe24<32><<<(unsigned)grid, kThreads, 0, st>>>(hi, lo, (uint32_t)h, (uint32_t)n_rows, per_group, sink);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

// The fp32 probe with 256-bit loads (h in {64, 128, 256}: h / 8 lanes per row, rows 32-byte aligned); `unroll` in {4, 8}.
extern "C" int glass_l2_gather_probe256(const float* x, int64_t ldx, int64_t n_rows, int h, int64_t gathers, float* sink,
                                        int64_t sink_elems, int unroll, void* stream) {
    GLASS_CHECK_ARG(x && sink && n_rows > 0 && gathers > 0 && (h == 64 || h == 128 || h == 256) && ldx >= h && ldx % 8 == 0 &&
                        (uintptr_t)x % 32 == 0 && n_rows * ldx < (1ll << 31) && (unroll == 4 || unroll == 8),
                    "l2_gather_probe256: bad arguments");
    const int g = h / 8;
    const int64_t grid = (int64_t)sm_count() * 4;
    GLASS_CHECK_ARG(sink_elems >= grid * kThreads, "l2_gather_probe256: sink needs %lld floats", (long long)(grid * kThreads));
    const int64_t groups = grid * kThreads / g;
    const int64_t per_group = (ceil_div(gathers, groups) + 7) / 8 * 8;
    cudaStream_t st = as_stream(stream);
#define GO(G, U) k_l2_gather_probe256<G, U><<<(unsigned)grid, kThreads, 0, st>>>(x, (uint32_t)ldx, (uint32_t)n_rows, per_group, sink)
    if (g == 8) { if (unroll == 8) GO(8, 8); else GO(8, 4); }
    else if (g == 16) { if (unroll == 8) GO(16, 8); else GO(16, 4); }
    else { if (unroll == 8) GO(32, 8); else GO(32, 4); }
#undef GO
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}
