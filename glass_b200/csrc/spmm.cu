// spmm_csr: y = A x for the normalised adjacency (reference impl/models.py:164, `self.adj @ x`;
// backward runs the same kernel on the transposed CSR).
//
// Mapping: a sub-warp GROUP of G lanes owns one CSR row; lane l of the group owns VEC consecutive
// feature columns (VEC = 4 -> one 16-byte gather per neighbour and lane, a whole 256-byte feature row
// per 16 lanes at H = 64).  The group streams its (col, val) entries one per lane at a time with one
// coalesced load per array, stages them in a warp-private shared-memory slot, and issues 4 independent
// float4 gathers before the dependent FMA chain.  In the throughput configuration accumulation is a
// single fp32 chain per column in CSR order: deterministic, equal to a sequential CPU loop over the
// sorted entries.  Small graphs use extra lanes across the neighbours of a row (see k_spmm).
//
// Roofline: compulsory HBM bytes = 4(N+1) + 8 nnz + 8 N H.  The gathers
// (4 H nnz bytes) are served by L1/L2: X (14.7 MB at the em_user shape) is L2-resident.  Measured tuning
// (em_user shape): 5 resident CTAs per SM (48 registers) 131 us; 3 / 6 / 8 CTAs 141 / 143 / 150 us; deeper
// unrolling with fewer CTAs 205-225 us; L1::no_allocate gathers 156 us.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

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
    static __device__ __forceinline__ void add_xor(T& a, unsigned mask, int off) {
        a.x += __shfl_xor_sync(mask, a.x, off);
        a.y += __shfl_xor_sync(mask, a.y, off);
        a.z += __shfl_xor_sync(mask, a.z, off);
        a.w += __shfl_xor_sync(mask, a.w, off);
    }
};
template <>
struct Vec<1> {
    using T = float;
    static __device__ __forceinline__ T zero() { return 0.f; }
    static __device__ __forceinline__ T load(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void store(float* p, const T& v) { *p = v; }
    static __device__ __forceinline__ void fma(T& a, float s, const T& x) { a = fmaf(s, x, a); }
    static __device__ __forceinline__ void add_xor(T& a, unsigned mask, int off) { a += __shfl_xor_sync(mask, a, off); }
};

// Work description: either one item per CSR row, or a precomputed plan in which rows longer than
// `max_len` entries are split into several items whose partial sums go to a scratch matrix
// (skewed / power-law graphs: a hub row would otherwise serialise on one lane group).
struct RowWork {
    const int32_t* rowptr;
    float* y;
    int64_t ldy;
    int accumulate;            // y += A x instead of y = A x (column-partitioned phases of one product)
    __device__ __forceinline__ bool get(int64_t item, int32_t& b, int32_t& e, float*& dst) const {
        b = __ldg(rowptr + item);
        e = __ldg(rowptr + item + 1);
        dst = y + item * ldy;
        return true;                   // a complete row of y
    }
    __device__ __forceinline__ int64_t row_of(int64_t item) const { return item; }
};
struct PlanWork {
    const int32_t* item_begin;
    const int32_t* item_end;
    const int32_t* item_dst;   // >= 0: row of y; < 0: scratch row (-1 - dst)
    float* y;
    int64_t ldy;
    float* scratch;            // [n_slots, ld_s]
    int64_t ld_s;
    int accumulate;            // complete rows start from y; split rows get y added by k_spmm_combine (base = y)
    __device__ __forceinline__ bool get(int64_t item, int32_t& b, int32_t& e, float*& dst) const {
        b = __ldg(item_begin + item);
        e = __ldg(item_end + item);
        const int32_t d = __ldg(item_dst + item);
        dst = d >= 0 ? y + (int64_t)d * ldy : scratch + (int64_t)(-1 - d) * ld_s;
        return d >= 0;                 // scratch rows are partial sums of a split row (k_spmm_combine finishes them)
    }
    __device__ __forceinline__ int64_t row_of(int64_t item) const { return __ldg(item_dst + item); }   // only if get() was true
};

// A row group is G x S lanes: lane l of G owns VEC consecutive feature columns per chunk (KCH column
// chunks, h <= G*VEC*KCH), slot s of S takes every S-th stored entry of the row; the S partial sums are
// added by a butterfly at the end of the row (fixed order -> deterministic).  S = 1 is the throughput
// configuration (one sequential fp32 chain per output, exactly the CPU loop order); S = 32/G is used
// for graphs that do not fill the machine, where the longest row's dependent gather chain is the
// whole kernel time (density: 409-entry row on 2 lanes = 65 us; 16 slots -> see DESIGN.md 4.1).
// EXACT: h == G*VEC*KCH (no column guards).  IDX32: every element offset into x fits 32 bits.
// STATS: the epilogue also emits this CTA's fp64 partial column sums of y and y^2 (the GraphNorm statistics of
// impl/models.py:165, so that no separate pass over y is needed): partial[(which * h + c) * ldp + blockIdx.x].
// Every output element is added exactly once, by the thread that stores it, in item order; the lanes / warps
// of a CTA are combined in a fixed order -> deterministic.  UNR: neighbour gathers in flight per lane.
template <int G, int S, int VEC, int KCH, bool EXACT, bool IDX32, int MINB, int UNR, bool STATS, class Work>
__global__ void __launch_bounds__(kThreads, MINB) k_spmm(const Work work, const int32_t* __restrict__ col,
                                                         const float* __restrict__ val, const float* __restrict__ x,
                                                         int64_t ldx, int64_t n_items, int h,
                                                         double* __restrict__ partial, int ldp) {
    using V = Vec<VEC>;
    constexpr int GW = G * S;                          // lanes per row
    static_assert(GW <= 32 && (GW & (GW - 1)) == 0, "row group must be a power-of-two part of a warp");
    // (col, val) staging: one 8-byte slot per lane, double buffered so that one __syncwarp per chunk suffices
    __shared__ int2 s_e[2][kThreads];
    const int lane = threadIdx.x & 31;
    const int lg = lane & (GW - 1);                    // lane inside the row group
    const int l = lg & (G - 1);                        // feature lane
    const int slot = lg / G;                           // neighbour slot
    const unsigned gmask = (GW == 32) ? 0xffffffffu : (((1u << GW) - 1u) << (lane & ~(GW - 1)));
    const int gbase = threadIdx.x & ~(GW - 1);
    const int64_t groups_per_grid = (int64_t)gridDim.x * (kThreads / GW);
    int64_t item = (int64_t)blockIdx.x * (kThreads / GW) + threadIdx.x / GW;
    const uint32_t ldx32 = (uint32_t)ldx;

    bool colok[KCH];
    int coff[KCH];
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
        coff[k] = (l + k * G) * VEC;
        colok[k] = EXACT || coff[k] < h;
    }

    // Statistics: every row group keeps fp32 running sums of the rows it finishes in its OWN shared-memory cells
    // (a group finishes only a handful of rows; keeping the sums in registers instead cost the gather loop its
    // memory-level parallelism: +14..20 us on the em_user shape); groups and CTAs are then added in fp64.
    constexpr int NCOL = G * VEC * KCH;
    __shared__ float s_acc[STATS ? 2 : 1][STATS ? kThreads / GW : 1][STATS ? NCOL : 1];
    const int grp = threadIdx.x / GW;
    if (STATS) {
        for (int i = threadIdx.x; i < 2 * (kThreads / GW) * NCOL; i += kThreads) (&s_acc[0][0][0])[i] = 0.f;
        __syncthreads();
    }

    for (; item < n_items; item += groups_per_grid) {
        int32_t e_begin, e_end;
        float* yr;
        const bool full_row = work.get(item, e_begin, e_end, yr);
        typename V::T acc[KCH];
#pragma unroll
        for (int k = 0; k < KCH; ++k) acc[k] = (work.accumulate && full_row && slot == 0 && colok[k]) ? V::load(yr + coff[k]) : V::zero();
        int buf = 0;
        // software pipeline: the (col, val) pair of the NEXT chunk is in flight while this chunk is gathered
        int nc = 0;
        float nv = 0.f;
        if (e_begin + lg < e_end) {
            nc = __ldg(col + e_begin + lg);
            nv = __ldg(val + e_begin + lg);
        }
        for (int32_t e0 = e_begin; e0 < e_end; e0 += GW, buf ^= 1) {
            const int cnt = min(GW, e_end - e0);
            s_e[buf][threadIdx.x] = make_int2(nc, __float_as_int(nv));
            if (e0 + GW + lg < e_end) {
                nc = __ldg(col + e0 + GW + lg);
                nv = __ldg(val + e0 + GW + lg);
            }
            __syncwarp(gmask);
            const int2* se = &s_e[buf][gbase];
#pragma unroll UNR
            for (int j = slot; j < cnt; j += S) {
                const int2 cv = se[j];                                  // broadcast read inside the group
                const float* xr = IDX32 ? x + (uint32_t)cv.x * ldx32 : x + (int64_t)cv.x * ldx;
                const float w = __int_as_float(cv.y);
#pragma unroll
                for (int k = 0; k < KCH; ++k)
                    if (colok[k]) V::fma(acc[k], w, V::load(xr + coff[k]));
            }
        }
        if (S > 1) {
#pragma unroll
            for (int off = G; off < GW; off <<= 1)
#pragma unroll
                for (int k = 0; k < KCH; ++k) V::add_xor(acc[k], gmask, off);
        }
        if (S == 1 || slot == 0) {
#pragma unroll
            for (int k = 0; k < KCH; ++k)
                if (colok[k]) V::store(yr + coff[k], acc[k]);
            if (STATS && full_row) {
#pragma unroll
                for (int k = 0; k < KCH; ++k) {
                    const float* a = reinterpret_cast<const float*>(&acc[k]);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        s_acc[0][grp][coff[k] + v] += a[v];
                        s_acc[1][grp][coff[k] + v] = fmaf(a[v], a[v], s_acc[1][grp][coff[k] + v]);
                    }
                }
            }
        }
    }
    if (STATS) {
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * h; i += kThreads) {       // groups of the CTA in order -> deterministic
            const int which = i / h, c = i - which * h;
            double t = 0.0;
#pragma unroll 4
            for (int g = 0; g < kThreads / GW; ++g) t += (double)s_acc[which][g][c];
            partial[((int64_t)which * h + c) * ldp + blockIdx.x] = t;
        }
    }
}

// y[long_row[i], :] = sum of its scratch rows, in chunk order (deterministic).  A fixed grid walks the long rows;
// with `partial` the CTAs also emit their column sums of the rows they finish (blocks first_blk .. first_blk+grid-1).
__global__ void k_spmm_combine(const int32_t* __restrict__ long_row, const int32_t* __restrict__ long_slot,
                               const int32_t* __restrict__ long_cnt, int64_t n_long, const float* __restrict__ scratch,
                               int64_t ld_s, float* y, int64_t ldy, int h, double* __restrict__ partial,
                               int ldp, int first_blk, const float* base = nullptr, int64_t ldb = 0) {   // base may be y
    for (int c = threadIdx.x; c < h; c += blockDim.x) {
        double s = 0.0, q = 0.0;
        for (int64_t i = blockIdx.x; i < n_long; i += gridDim.x) {
            const int32_t r = long_row[i], s0 = long_slot[i], cnt = long_cnt[i];
            float acc = base ? base[(int64_t)r * ldb + c] : 0.f;
            for (int k = 0; k < cnt; ++k) acc += scratch[(int64_t)(s0 + k) * ld_s + c];
            y[(int64_t)r * ldy + c] = acc;
            s += (double)acc;
            q = fma((double)acc, (double)acc, q);
        }
        if (partial) {
            partial[(int64_t)c * ldp + first_blk + blockIdx.x] = s;
            partial[((int64_t)h + c) * ldp + first_blk + blockIdx.x] = q;
        }
    }
}

// The same sum for widths h = 4 G (G a power of two <= 32) without statistics: one lane group per long row and a float4
// per lane, up to sixteen scratch rows in flight, added in chunk order (bit-identical to k_spmm_combine).  The column-per-thread
// kernel above walks 16 long rows per CTA one after the other on h threads: 14 us for the 2,345 split rows of the
// power-law em_user-shaped graph.
template <int G>
__global__ void __launch_bounds__(kThreads, 2) k_spmm_combine_vec(const int32_t* __restrict__ long_row,
                                                               const int32_t* __restrict__ long_slot,
                                                               const int32_t* __restrict__ long_cnt, int64_t n_long,
                                                               const float* __restrict__ scratch, int64_t ld_s, float* y,
                                                               int64_t ldy, const float* base, int64_t ldb) {
    const int l = threadIdx.x & (G - 1);
    const int64_t groups = (int64_t)gridDim.x * (kThreads / G);
    for (int64_t i = (int64_t)blockIdx.x * (kThreads / G) + threadIdx.x / G; i < n_long; i += groups) {
        const int32_t r = __ldg(long_row + i), s0 = __ldg(long_slot + i), cnt = __ldg(long_cnt + i);
        float4 acc = base ? *reinterpret_cast<const float4*>(base + (int64_t)r * ldb + 4 * l) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* src = reinterpret_cast<const float4*>(scratch + (int64_t)s0 * ld_s) + l;
        const int64_t step = ld_s >> 2;
        int k = 0;
        // sixteen scratch rows in flight: the longest row of the power-law em_user-shaped graph has 107 chunks, and
        // with four loads per round trip its 27 dependent round trips WERE the kernel (22 us for 2.4 MB)
        for (; k + 16 <= cnt; k += 16) {
            float4 t[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) t[u] = __ldg(src + (int64_t)(k + u) * step);
            asm volatile("" ::: "memory");     // all sixteen loads are issued before the first (ordered) add
#pragma unroll
            for (int u = 0; u < 16; ++u) acc.x += t[u].x, acc.y += t[u].y, acc.z += t[u].z, acc.w += t[u].w;
        }
        for (; k + 4 <= cnt; k += 4) {
            const float4 a = __ldg(src + (int64_t)k * step), b = __ldg(src + (int64_t)(k + 1) * step);
            const float4 c = __ldg(src + (int64_t)(k + 2) * step), d = __ldg(src + (int64_t)(k + 3) * step);
            acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
            acc.x += b.x, acc.y += b.y, acc.z += b.z, acc.w += b.w;
            acc.x += c.x, acc.y += c.y, acc.z += c.z, acc.w += c.w;
            acc.x += d.x, acc.y += d.y, acc.z += d.z, acc.w += d.w;
        }
        for (; k < cnt; ++k) {
            const float4 a = __ldg(src + (int64_t)k * step);
            acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
        }
        *reinterpret_cast<float4*>(y + (int64_t)r * ldy + 4 * l) = acc;
    }
}

// Sparse label correction (reference impl/train.py:20-34 + impl/models.py:161-164).
// For fixed weights the label-mixed features of two label batches differ only on the labelled rows:
//   x_b = U + [labelled] * delta,   so   adj @ x_b = adj @ U + adj[:, labelled] @ delta[labelled]
// with adj @ U (`base`) computed once per evaluation epoch.  This kernel forms, per batch,
//   y[i, :] = base[i, :] + sum_{e in row i, mask[col[e]] != 0} val[e] * delta[col[e], :]      (CSR order)
// It streams the column indices (4 bytes per entry), tests the label byte of every neighbour and gathers only
// for the ~1-2 % labelled ones -- instead of 4*H bytes per entry.  Same statistics epilogue as k_spmm.
template <int G, bool STATS, class Work>
__global__ void __launch_bounds__(kThreads) k_spmm_delta(const Work work, const int32_t* __restrict__ col,
                                                         const float* __restrict__ val, const uint8_t* __restrict__ mask,
                                                         const float* __restrict__ delta, int64_t ldd,
                                                         const float* __restrict__ base, int64_t ldb, int64_t n_items,
                                                         int h, double* __restrict__ partial, int ldp) {
    const int lane = threadIdx.x & 31;
    const int l = lane & (G - 1);
    const int gshift = lane & ~(G - 1);
    const unsigned gbits = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
    const unsigned gmask = gbits << gshift;
    const int64_t groups_per_grid = (int64_t)gridDim.x * (kThreads / G);
    const int coff = l * 4;
    const bool colok = coff < h;
    float st_s[4] = {0.f, 0.f, 0.f, 0.f}, st_q[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t item = (int64_t)blockIdx.x * (kThreads / G) + threadIdx.x / G; item < n_items; item += groups_per_grid) {
        int32_t e_begin, e_end;
        float* yr;
        const bool full_row = work.get(item, e_begin, e_end, yr);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (full_row && colok) acc = ldg_f4(base + work.row_of(item) * ldb + coff);
        // software pipeline: the next chunk's indices / label bytes are in flight while this chunk's hits are gathered
        int32_t nc = -1;
        bool nhit = false;
        if (e_begin + l < e_end) {
            nc = __ldg(col + e_begin + l);
            nhit = __ldg(mask + nc) != 0;
        }
        for (int32_t e0 = e_begin; e0 < e_end; e0 += G) {
            const int32_t c = nc;
            const bool hit = nhit;
            nc = -1;
            nhit = false;
            if (e0 + G + l < e_end) {
                nc = __ldg(col + e0 + G + l);
                nhit = __ldg(mask + nc) != 0;
            }
            unsigned hits = (__ballot_sync(gmask, hit) >> gshift) & gbits;
            const float w = hit ? __ldg(val + e0 + l) : 0.f;
            while (hits) {
                const int j = __ffs(hits) - 1;
                hits &= hits - 1;
                const int32_t cj = __shfl_sync(gmask, c, j, G);
                const float wj = __shfl_sync(gmask, w, j, G);
                if (colok) {
                    const float4 d = ldg_f4(delta + (int64_t)cj * ldd + coff);
                    acc.x = fmaf(wj, d.x, acc.x), acc.y = fmaf(wj, d.y, acc.y);
                    acc.z = fmaf(wj, d.z, acc.z), acc.w = fmaf(wj, d.w, acc.w);
                }
            }
        }
        if (colok) {
            *reinterpret_cast<float4*>(yr + coff) = acc;
            if (STATS && full_row) {
                st_s[0] += acc.x, st_s[1] += acc.y, st_s[2] += acc.z, st_s[3] += acc.w;
                st_q[0] = fmaf(acc.x, acc.x, st_q[0]), st_q[1] = fmaf(acc.y, acc.y, st_q[1]);
                st_q[2] = fmaf(acc.z, acc.z, st_q[2]), st_q[3] = fmaf(acc.w, acc.w, st_q[3]);
            }
        }
    }
    if (STATS) {
        __shared__ double s_red[kThreads / 32][G * 4];
        const int warp = threadIdx.x >> 5;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double t = (double)(which ? st_q[i] : st_s[i]);
#pragma unroll
                for (int off = G; off < 32; off <<= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
                if (lane < G) s_red[warp][coff + i] = t;
            }
            __syncthreads();
            for (int c = threadIdx.x; c < h; c += kThreads) {
                double t = 0.0;
#pragma unroll
                for (int w = 0; w < kThreads / 32; ++w) t += s_red[w][c];
                partial[((int64_t)which * h + c) * ldp + blockIdx.x] = t;
            }
            __syncthreads();
        }
    }
}

struct Plan {   // host view of the arguments of glass_spmm_csr_planned
    const int32_t *item_begin, *item_end, *item_dst;
    int64_t n_items;
    const int32_t *long_row, *long_slot, *long_cnt;
    int64_t n_long;
    float* scratch;
};

struct Stats {   // optional epilogue statistics (GraphNorm column sums)
    double* partial;
    int ldp;
    int* nblk_host;
};

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// tuning knobs (environment at start-up, glass_tune() at run time; measured settings are the defaults)
static int g_variant = env_int("GLASS_B200_SPMM_VARIANT", 0);
static int g_waves = env_int("GLASS_B200_SPMM_WAVES", 0);

// Lanes of one full-occupancy wave (148 SMs x 2048 threads on B200): below about two of them the kernel is
// bound by its longest row, not by gather throughput, and the neighbour-parallel configuration wins.
static bool latency_regime(int64_t n_items) {
    static const int force = env_int("GLASS_B200_SPMM_SLOTS", -1);   // 0: never, 1: always, unset: by size
    if (force >= 0) return force != 0;
    return n_items * 32 <= 2ll * sm_count() * 2048;
}

constexpr int kCombineCtas = 148;

static int64_t grid_cap(bool stats) {
    // plain: a few waves of resident CTAs (items are interleaved).  With statistics every CTA leaves a partial
    // that the finalize kernel has to add, so the grid is kept to about two waves.
    if (g_waves > 0) return (int64_t)sm_count() * 5 * g_waves;
    return stats ? (int64_t)sm_count() * 5 * 2 : (int64_t)sm_count() * 8 * 4;
}

// combine launcher: the lane-group kernel when no statistics are wanted and the rows are float4-addressable
static int launch_combine(const Plan* plan, float* y, int64_t ldy, int h, double* partial, int ldp, int first_blk,
                          const float* base, int64_t ldb, cudaStream_t st) {
    const int g = h / 4;
    const bool vec = !partial && h % 4 == 0 && g >= 1 && g <= 32 && (g & (g - 1)) == 0 && ldy % 4 == 0 && (uintptr_t)y % 16 == 0 &&
                     (uintptr_t)plan->scratch % 16 == 0 && (!base || (ldb % 4 == 0 && (uintptr_t)base % 16 == 0));
    if (vec) {
        const int64_t per = kThreads / g;
        const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(plan->n_long, per), (int64_t)sm_count() * 8);
#define GO(G) k_spmm_combine_vec<G><<<grid, kThreads, 0, st>>>(plan->long_row, plan->long_slot, plan->long_cnt, plan->n_long, \
                                                              plan->scratch, (int64_t)h, y, ldy, base, ldb)
        switch (g) {
            case 1: GO(1); break;
            case 2: GO(2); break;
            case 4: GO(4); break;
            case 8: GO(8); break;
            case 16: GO(16); break;
            default: GO(32); break;
        }
#undef GO
    } else {
        const int n_comb = (int)std::min<int64_t>(plan->n_long, kCombineCtas);
        const int threads = h <= 32 ? 32 : (h >= 256 ? 256 : (h + 31) / 32 * 32);
        k_spmm_combine<<<(unsigned)n_comb, threads, 0, st>>>(plan->long_row, plan->long_slot, plan->long_cnt, plan->n_long,
                                                           plan->scratch, (int64_t)h, y, ldy, h, partial, ldp, first_blk, base, ldb);
    }
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

template <int G, int S, int VEC, int KCH, int MINB = 4, int UNR = 4>
int launch(const int32_t* rowptr, const int32_t* col, const float* val, const float* x, int64_t ldx, float* y,
           int64_t ldy, int64_t n_rows, int64_t n_cols, int h, const Plan* plan, const Stats* stats, cudaStream_t st,
           int accumulate = 0) {
    const int64_t n_items = plan ? plan->n_items : n_rows;
    const int64_t groups_per_block = kThreads / (G * S);
    int64_t blocks = ceil_div(n_items, groups_per_block);
    const int64_t cap = grid_cap(stats != nullptr);
    if (blocks > cap) blocks = cap;
    const bool exact = h == G * VEC * KCH;
    const bool idx32 = n_cols * ldx < (1ll << 31);
    const unsigned grid = (unsigned)blocks;
    const bool comb = plan && plan->n_long > 0;
    const int n_comb = comb ? (int)std::min<int64_t>(plan->n_long, kCombineCtas) : 0;
    double* partial = stats ? stats->partial : nullptr;
    const int ldp = stats ? stats->ldp : 0;
    if (stats) {
        if ((int64_t)grid + n_comb > ldp) {
            set_error("spmm_csr: statistics table too small (ldp %d < %lld blocks)", ldp, (long long)grid + n_comb);
            return GLASS_ERR_WORKSPACE;
        }
        *stats->nblk_host = (int)grid + n_comb;
    }
#define GLASS_SPMM_GO2(E, I, ST)                                                                                      \
    do {                                                                                                              \
        if (plan) {                                                                                                   \
            PlanWork w{plan->item_begin, plan->item_end, plan->item_dst, y, ldy, plan->scratch, (int64_t)h, accumulate}; \
            k_spmm<G, S, VEC, KCH, E, I, MINB, UNR, ST, PlanWork>                                                     \
                <<<grid, kThreads, 0, st>>>(w, col, val, x, ldx, n_items, h, partial, ldp);                           \
        } else {                                                                                                      \
            RowWork w{rowptr, y, ldy, accumulate};                                                                    \
            k_spmm<G, S, VEC, KCH, E, I, MINB, UNR, ST, RowWork>                                                      \
                <<<grid, kThreads, 0, st>>>(w, col, val, x, ldx, n_items, h, partial, ldp);                           \
        }                                                                                                             \
    } while (0)
#define GLASS_SPMM_GO(E, I)                                                                                           \
    do {                                                                                                              \
        if (stats) GLASS_SPMM_GO2(E, I, true);                                                                        \
        else GLASS_SPMM_GO2(E, I, false);                                                                             \
    } while (0)
    if (exact && idx32) GLASS_SPMM_GO(true, true);
    else if (exact) GLASS_SPMM_GO(true, false);
    else if (idx32) GLASS_SPMM_GO(false, true);
    else GLASS_SPMM_GO(false, false);
#undef GLASS_SPMM_GO
#undef GLASS_SPMM_GO2
    GLASS_LAUNCH_CHECK();
    if (comb) return launch_combine(plan, y, ldy, h, partial, ldp, (int)grid, accumulate ? y : nullptr, ldy, st);
    return GLASS_OK;
}

int dispatch(const int32_t* rowptr, const int32_t* col, const float* val, const float* x, int64_t ldx, float* y,
             int64_t ldy, int64_t n_rows, int64_t n_cols, int h, const Plan* plan, const Stats* stats, cudaStream_t st,
             int accumulate = 0) {
    const bool vec = (h % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && ((uintptr_t)x % 16 == 0) &&
                     ((uintptr_t)y % 16 == 0) && (!plan || (uintptr_t)plan->scratch % 16 == 0);
    const int64_t n_items = plan ? plan->n_items : n_rows;
    const bool par = latency_regime(n_items);
#define ARGS rowptr, col, val, x, ldx, y, ldy, n_rows, n_cols, h, plan, stats, st, accumulate
#define GO(G, V, K)                                                                                                   \
    do {                                                                                                              \
        if (par && G < 32) return launch<G, 32 / G, V, K>(ARGS);                                                      \
        return launch<G, 1, V, K>(ARGS);                                                                              \
    } while (0)
    if (vec) {
        const int lanes = h / 4;
        if (lanes <= 2) GO(2, 4, 1);
        if (lanes <= 4) GO(4, 4, 1);
        if (lanes <= 8) GO(8, 4, 1);
        if (lanes <= 16) {
            if (par) return launch<16, 2, 4, 1>(ARGS);
            // throughput configuration of the 64-wide layers.  GLASS_B200_SPMM_VARIANT selects tuning variants
            // (gathers in flight per lane / resident CTAs per SM / lanes per row); measured numbers in DESIGN.md 4.1
            switch (g_variant) {
                case 1: return launch<16, 1, 4, 1, 4, 8>(ARGS);
                case 2: return launch<8, 1, 4, 2, 5, 4>(ARGS);
                case 3: return launch<8, 1, 4, 2, 4, 4>(ARGS);
                case 4: return launch<8, 1, 4, 2, 3, 8>(ARGS);
                case 5: return launch<16, 1, 4, 1, 6, 4>(ARGS);
                case 6: return launch<16, 1, 4, 1, 3, 8>(ARGS);
                case 7: return launch<16, 1, 4, 1, 4, 4>(ARGS);
                default: return launch<16, 1, 4, 1, 5, 4>(ARGS);
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
#undef ARGS
}

}  // namespace
}  // namespace glass

using namespace glass;

// Run-time tuning knobs of the SpMM dispatcher (benchmark scripts): "spmm_variant", "spmm_waves".
extern "C" int glass_tune(const char* name, int value) {
    GLASS_CHECK_ARG(name, "tune: null name");
    if (!strcmp(name, "spmm_variant")) g_variant = value;
    else if (!strcmp(name, "spmm_waves")) g_waves = value;
    else {
        set_error("tune: unknown knob %s", name);
        return GLASS_ERR_BAD_ARG;
    }
    return GLASS_OK;
}

// Upper bound of the statistics blocks a SpMM launch can produce (leading dimension of the partial table).
extern "C" int glass_spmm_stats_ld(void) {
    const int n = sm_count();
    if (n <= 0) return 0;
    return (int)align_up((size_t)sm_count() * 8 * 4 + kCombineCtas, 32);   // >= any grid the dispatcher launches
}

extern "C" int glass_spmm_csr(const int32_t* rowptr, const int32_t* col, const float* val, const float* x,
                              int64_t ldx, float* y, int64_t ldy, int64_t n_rows, int64_t n_cols, int h,
                              double* stats_partial, int stats_ld, int* stats_nblk_host, void* stream) {
    GLASS_CHECK_ARG(rowptr && x && y && n_rows >= 0 && n_cols > 0 && h > 0 && ldx >= h && ldy >= h,
                    "spmm_csr: bad arguments");
    GLASS_CHECK_ARG(h <= 256, "spmm_csr: h=%d > 256 not supported", h);
    GLASS_CHECK_ARG(!stats_partial || (stats_ld > 0 && stats_nblk_host), "spmm_csr: statistics arguments incomplete");
    if (stats_nblk_host) *stats_nblk_host = 0;
    if (n_rows == 0) return GLASS_OK;
    Stats s{stats_partial, stats_ld, stats_nblk_host};
    return dispatch(rowptr, col, val, x, ldx, y, ldy, n_rows, n_cols, h, nullptr, stats_partial ? &s : nullptr,
                    as_stream(stream));
}

// ---- sparse label correction ---------------------------------------------------------------------------------
namespace glass {
namespace {
template <int G>
int launch_delta(const int32_t* rowptr, const int32_t* col, const float* val, const uint8_t* mask, const float* delta,
                 int64_t ldd, const float* base, int64_t ldb, float* y, int64_t ldy, int64_t n_rows, int h,
                 const Plan* plan, const Stats* stats, cudaStream_t st) {
    const int64_t n_items = plan ? plan->n_items : n_rows;
    int64_t blocks = ceil_div(n_items, kThreads / G);
    const int64_t cap = (int64_t)sm_count() * 8 * (stats ? 1 : 4);
    if (blocks > cap) blocks = cap;
    const unsigned grid = (unsigned)blocks;
    const bool comb = plan && plan->n_long > 0;
    const int n_comb = comb ? (int)std::min<int64_t>(plan->n_long, kCombineCtas) : 0;
    double* partial = stats ? stats->partial : nullptr;
    const int ldp = stats ? stats->ldp : 0;
    if (stats) {
        if ((int64_t)grid + n_comb > ldp) {
            set_error("spmm_delta: statistics table too small (ldp %d < %lld blocks)", ldp, (long long)grid + n_comb);
            return GLASS_ERR_WORKSPACE;
        }
        *stats->nblk_host = (int)grid + n_comb;
    }
    if (plan) {
        PlanWork w{plan->item_begin, plan->item_end, plan->item_dst, y, ldy, plan->scratch, (int64_t)h};
        if (stats) k_spmm_delta<G, true, PlanWork><<<grid, kThreads, 0, st>>>(w, col, val, mask, delta, ldd, base, ldb, n_items, h, partial, ldp);
        else k_spmm_delta<G, false, PlanWork><<<grid, kThreads, 0, st>>>(w, col, val, mask, delta, ldd, base, ldb, n_items, h, partial, ldp);
    } else {
        RowWork w{rowptr, y, ldy};
        if (stats) k_spmm_delta<G, true, RowWork><<<grid, kThreads, 0, st>>>(w, col, val, mask, delta, ldd, base, ldb, n_items, h, partial, ldp);
        else k_spmm_delta<G, false, RowWork><<<grid, kThreads, 0, st>>>(w, col, val, mask, delta, ldd, base, ldb, n_items, h, partial, ldp);
    }
    GLASS_LAUNCH_CHECK();
    if (comb) return launch_combine(plan, y, ldy, h, partial, ldp, (int)grid, base, ldb, st);
    return GLASS_OK;
}
}  // namespace
}  // namespace glass

extern "C" int glass_spmm_delta(const int32_t* rowptr, const int32_t* col, const float* val, const uint8_t* mask,
                                const float* delta, int64_t ldd, const float* base, int64_t ldb, float* y, int64_t ldy,
                                int64_t n_rows, int h, const int32_t* item_begin, const int32_t* item_end,
                                const int32_t* item_dst, int64_t n_items, const int32_t* long_row,
                                const int32_t* long_slot, const int32_t* long_cnt, int64_t n_long, float* scratch,
                                double* stats_partial, int stats_ld, int* stats_nblk_host, void* stream) {
    GLASS_CHECK_ARG(col && val && mask && delta && base && y && n_rows >= 0 && h > 0 && ldd >= h && ldb >= h && ldy >= h,
                    "spmm_delta: bad arguments");
    GLASS_CHECK_ARG(h % 4 == 0 && h <= 128 && ldd % 4 == 0 && ldb % 4 == 0 && ldy % 4 == 0 &&
                        ((uintptr_t)delta | (uintptr_t)base | (uintptr_t)y | (uintptr_t)scratch) % 16 == 0,
                    "spmm_delta: needs h %% 4 == 0, h <= 128 and 16-byte aligned rows");
    const bool planned = item_begin != nullptr;
    GLASS_CHECK_ARG(planned ? (item_end && item_dst && n_items >= n_rows && (n_long == 0 || (long_row && long_slot && long_cnt && scratch)))
                            : rowptr != nullptr,
                    "spmm_delta: give either rowptr or a complete row-split plan");
    GLASS_CHECK_ARG(!stats_partial || (stats_ld > 0 && stats_nblk_host), "spmm_delta: statistics arguments incomplete");
    if (stats_nblk_host) *stats_nblk_host = 0;
    if (n_rows == 0) return GLASS_OK;
    Plan p{item_begin, item_end, item_dst, n_items, long_row, long_slot, long_cnt, n_long, scratch};
    Stats s{stats_partial, stats_ld, stats_nblk_host};
    const Plan* pp = planned ? &p : nullptr;
    const Stats* sp = stats_partial ? &s : nullptr;
    cudaStream_t st = as_stream(stream);
    const int lanes = h / 4;
#define GO(G) return launch_delta<G>(rowptr, col, val, mask, delta, ldd, base, ldb, y, ldy, n_rows, h, pp, sp, st)
    if (lanes <= 2) GO(2);
    if (lanes <= 4) GO(4);
    if (lanes <= 8) GO(8);
    if (lanes <= 16) GO(16);
    GO(32);
#undef GO
}

// ---- row-splitting plan (init path; reads rowptr back to the host) ------------------------------------
static int plan_host(const int32_t* rowptr_dev, int64_t n_rows, int max_len, std::vector<int32_t>& rp, cudaStream_t st) {
    rp.resize((size_t)n_rows + 1);
    GLASS_CUDA(cudaMemcpyAsync(rp.data(), rowptr_dev, sizeof(int32_t) * rp.size(), cudaMemcpyDeviceToHost, st));
    GLASS_CUDA(cudaStreamSynchronize(st));
    (void)max_len;
    return GLASS_OK;
}

extern "C" int glass_spmm_plan_size(const int32_t* rowptr, int64_t n_rows, int max_len, int64_t* n_items_host,
                                    int64_t* n_long_host, int64_t* n_slots_host, void* stream) {
    GLASS_CHECK_ARG(rowptr && n_rows >= 0 && max_len >= 32 && n_items_host && n_long_host && n_slots_host,
                    "spmm_plan_size: bad arguments");
    std::vector<int32_t> rp;
    int rc = plan_host(rowptr, n_rows, max_len, rp, as_stream(stream));
    if (rc != GLASS_OK) return rc;
    int64_t items = 0, nlong = 0, slots = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t deg = rp[r + 1] - rp[r];
        if (deg > max_len) {
            const int64_t c = (deg + max_len - 1) / max_len;
            items += c, slots += c, ++nlong;
        } else {
            ++items;
        }
    }
    *n_items_host = items, *n_long_host = nlong, *n_slots_host = slots;
    return GLASS_OK;
}

extern "C" int glass_spmm_plan_build(const int32_t* rowptr, int64_t n_rows, int max_len, int32_t* item_begin,
                                     int32_t* item_end, int32_t* item_dst, int32_t* long_row, int32_t* long_slot,
                                     int32_t* long_cnt, void* stream) {
    GLASS_CHECK_ARG(rowptr && n_rows >= 0 && max_len >= 32 && item_begin && item_end && item_dst,
                    "spmm_plan_build: bad arguments");
    cudaStream_t st = as_stream(stream);
    std::vector<int32_t> rp;
    int rc = plan_host(rowptr, n_rows, max_len, rp, st);
    if (rc != GLASS_OK) return rc;
    std::vector<int32_t> ib, ie, id, lr, ls, lc;
    // heavy items first (chunks of split rows), then one item per ordinary row
    int32_t slot = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t deg = rp[r + 1] - rp[r];
        if (deg <= max_len) continue;
        const int32_t c = (int32_t)((deg + max_len - 1) / max_len);
        lr.push_back((int32_t)r), ls.push_back(slot), lc.push_back(c);
        for (int32_t k = 0; k < c; ++k) {
            const int32_t b = rp[r] + k * max_len;
            ib.push_back(b);
            ie.push_back(std::min<int32_t>(b + max_len, rp[r + 1]));
            id.push_back(-1 - (slot + k));
        }
        slot += c;
    }
    // ... longest first (stable): the lane groups of a warp / CTA work on rows of nearly equal length -- in row order a
    // CTA lives as long as the longest of its 16 random rows while the other groups idle (power-law em_user-shaped
    // graph: achieved occupancy 42 %), and the grid's tail is made of the shortest rows
    std::vector<int32_t> ord;
    ord.reserve((size_t)n_rows);
    for (int64_t r = 0; r < n_rows; ++r)
        if (rp[r + 1] - rp[r] <= max_len) ord.push_back((int32_t)r);
    std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return rp[a + 1] - rp[a] > rp[b + 1] - rp[b]; });
    for (int32_t r : ord) ib.push_back(rp[r]), ie.push_back(rp[r + 1]), id.push_back(r);
    auto up = [&](int32_t* dst, const std::vector<int32_t>& v) -> cudaError_t {
        if (v.empty()) return cudaSuccess;
        return cudaMemcpyAsync(dst, v.data(), sizeof(int32_t) * v.size(), cudaMemcpyHostToDevice, st);
    };
    GLASS_CUDA(up(item_begin, ib));
    GLASS_CUDA(up(item_end, ie));
    GLASS_CUDA(up(item_dst, id));
    if (!lr.empty()) {
        GLASS_CHECK_ARG(long_row && long_slot && long_cnt, "spmm_plan_build: long-row outputs missing");
        GLASS_CUDA(up(long_row, lr));
        GLASS_CUDA(up(long_slot, ls));
        GLASS_CUDA(up(long_cnt, lc));
    }
    GLASS_CUDA(cudaStreamSynchronize(st));   // host vectors go out of scope
    return GLASS_OK;
}

extern "C" int glass_spmm_csr_planned(const int32_t* col, const float* val, const float* x, int64_t ldx, float* y,
                                      int64_t ldy, int64_t n_rows, int64_t n_cols, int h, const int32_t* item_begin,
                                      const int32_t* item_end, const int32_t* item_dst, int64_t n_items,
                                      const int32_t* long_row, const int32_t* long_slot, const int32_t* long_cnt,
                                      int64_t n_long, float* scratch, double* stats_partial, int stats_ld,
                                      int* stats_nblk_host, void* stream) {
    GLASS_CHECK_ARG(col && val && x && y && n_rows >= 0 && n_cols > 0 && h > 0 && ldx >= h && ldy >= h && item_begin &&
                        item_end && item_dst && n_items >= n_rows && (n_long == 0 || (long_row && long_slot && long_cnt && scratch)),
                    "spmm_csr_planned: bad arguments");
    GLASS_CHECK_ARG(h <= 256, "spmm_csr_planned: h=%d > 256 not supported", h);
    GLASS_CHECK_ARG(!stats_partial || (stats_ld > 0 && stats_nblk_host), "spmm_csr_planned: statistics arguments incomplete");
    if (stats_nblk_host) *stats_nblk_host = 0;
    if (n_items == 0) return GLASS_OK;
    Plan p{item_begin, item_end, item_dst, n_items, long_row, long_slot, long_cnt, n_long, scratch};
    Stats s{stats_partial, stats_ld, stats_nblk_host};
    return dispatch(nullptr, col, val, x, ldx, y, ldy, n_rows, n_cols, h, &p, stats_partial ? &s : nullptr,
                    as_stream(stream));
}

// ---- L2 gather probe (measurement aid) -----------------------------------------------------------------------------
namespace glass {
namespace {
template <int G>
__global__ void __launch_bounds__(kThreads, 5) k_l2_gather_probe(const float* __restrict__ x, uint32_t ldx, uint32_t n_rows,
                                                               int64_t per_group, float* __restrict__ sink) {
    const int lane = threadIdx.x & 31, l = lane & (G - 1);
    const int64_t group = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / G;
    // one multiplicative-congruential stream per lane group (the same value in all of its lanes)
    uint32_t s = (uint32_t)group * 2654435761u + 12345u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = 0; i < per_group; i += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t r = (uint32_t)(((uint64_t)s * n_rows) >> 32);
            v[u] = __ldg(reinterpret_cast<const float4*>(x + r * ldx) + l);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc.x += v[u].x, acc.y += v[u].y, acc.z += v[u].z, acc.w += v[u].w;
    }
    sink[(int64_t)blockIdx.x * kThreads + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}
}  // namespace
}  // namespace glass

extern "C" int glass_l2_gather_probe(const float* x, int64_t ldx, int64_t n_rows, int h, int64_t gathers, float* sink,
                                     int64_t sink_elems, void* stream) {
    GLASS_CHECK_ARG(x && sink && n_rows > 0 && gathers > 0 && (h == 32 || h == 64 || h == 128) && ldx >= h && ldx % 4 == 0 &&
                        (uintptr_t)x % 16 == 0 && n_rows * ldx < (1ll << 31),
                    "l2_gather_probe: bad arguments");
    const int g = h / 4;
    const int64_t grid = (int64_t)sm_count() * 5;
    GLASS_CHECK_ARG(sink_elems >= grid * kThreads, "l2_gather_probe: sink needs %lld floats", (long long)(grid * kThreads));
    const int64_t groups = grid * kThreads / g;
    const int64_t per_group = (ceil_div(gathers, groups) + 7) / 8 * 8;
    cudaStream_t st = as_stream(stream);
    if (g == 8) k_l2_gather_probe<8><<<(unsigned)grid, kThreads, 0, st>>>(x, (uint32_t)ldx, (uint32_t)n_rows, per_group, sink);
    else if (g == 16) k_l2_gather_probe<16><<<(unsigned)grid, kThreads, 0, st>>>(x, (uint32_t)ldx, (uint32_t)n_rows, per_group, sink);
    else k_l2_gather_probe<32><<<(unsigned)grid, kThreads, 0, st>>>(x, (uint32_t)ldx, (uint32_t)n_rows, per_group, sink);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

// y (+)= A x: glass_spmm_csr / glass_spmm_csr_planned with an accumulate flag (column-partitioned phases of one product,
// glass_b200/partition.py: the entries whose columns a peer owns are multiplied as that peer's shard arrives).
// item_begin == NULL: plain CSR by rowptr; otherwise a row-split plan.  No statistics epilogue.
extern "C" int glass_spmm_csr_acc(const int32_t* rowptr, const int32_t* col, const float* val, const float* x, int64_t ldx,
                                  float* y, int64_t ldy, int64_t n_rows, int64_t n_cols, int h, const int32_t* item_begin,
                                  const int32_t* item_end, const int32_t* item_dst, int64_t n_items,
                                  const int32_t* long_row, const int32_t* long_slot, const int32_t* long_cnt,
                                  int64_t n_long, float* scratch, int accumulate, void* stream) {
    GLASS_CHECK_ARG(x && y && n_rows >= 0 && n_cols > 0 && h > 0 && h <= 256 && ldx >= h && ldy >= h, "spmm_csr_acc: bad arguments");
    const bool planned = item_begin != nullptr;
    GLASS_CHECK_ARG(planned ? (col && val && item_end && item_dst && n_items >= n_rows &&
                               (n_long == 0 || (long_row && long_slot && long_cnt && scratch)))
                            : rowptr != nullptr,
                    "spmm_csr_acc: give either rowptr or a complete row-split plan");
    if (n_rows == 0) return GLASS_OK;
    Plan p{item_begin, item_end, item_dst, n_items, long_row, long_slot, long_cnt, n_long, scratch};
    return dispatch(rowptr, col, val, x, ldx, y, ldy, n_rows, n_cols, h, planned ? &p : nullptr, nullptr, as_stream(stream),
                    accumulate ? 1 : 0);
}
