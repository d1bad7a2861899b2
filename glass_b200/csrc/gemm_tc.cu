// Label-mixed pair of Linear layers on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
// Reference math: impl/models.py:158-162 / :167-173 (forward) and the dX half of its autograd.
//
//   forward : D[128 x 2H] = [a1|a2][128 x K] * [W0;W1]^T         epilogue: +bias, act, label mix -> out[128 x H]
//   backward: D[128 x K ] = dP[128 x 2H]     * [W0;W1]           dP built on the fly from dOut, acts, mask
//
// fp32 parity on TF32 tensor cores: 3xTF32.  Every operand is split in registers into hi = top 19 bits and
// lo = x - hi (exact); D = Ahi*Bhi + Alo*Bhi + Ahi*Blo accumulates in fp32 in TMEM, error ~2^-21 relative,
// far inside the 1e-4 bar (the dropped Alo*Blo term is ~2^-22).  K <= 256, so the GEMM is HBM-bound even at
// 3x the MMA count.
//
// CTA = 25 warps, persistent over 128-row tiles (grid = #SMs):
//   warps 0-7  epilogue: tcgen05.ld (TMEM lane = row; warp w reads lane quadrant w % 4, warps w and w+4
//              split the columns) -> bias/act/mix -> global
//   warp  8    TMEM allocator; lane 0 issues tcgen05.mma / tcgen05.commit
//   warps 9-24 operand loader: coalesced float4 global loads kept in a register ring 3-4 K-blocks deep
//              -> hi/lo split -> st.shared into the 128B-swizzled K-major UMMA layout ->
//              fence.proxy.async -> mbarrier arrive
// The weight operand (both weight sets, hi and lo) stays resident in shared memory for the CTA's lifetime;
// the row operand streams through a ring of 32-float K-blocks.  Two accumulators (2 x N TMEM columns)
// overlap the epilogue of tile i with the MMAs of tile i+1.
#include <stdlib.h>

#include "common.cuh"

namespace glass {
namespace {

constexpr int BM = 128;                 // rows per tile == UMMA M (TMEM lane == row)
constexpr int KBF = 32;                 // floats per K-block == one 128-byte swizzle row
// 25 warps per CTA: EW epilogue warps, one MMA-issuing warp, 24 - EW operand-loader warps.  EW = 8 (16 loaders) where the
// loader sets the pace (dX: it rebuilds dP from dOut / acts / mask), EW = 16 (8 loaders) where the epilogue does
// (forward with saved activations: 9 K cycles per tile on 8 warps against 1.4 K cycles per K-block of the main loop).
constexpr int kWorkWarps = 24;
constexpr int kThreads = (kWorkWarps + 1) * 32;
constexpr int kMaxEpiWarps = 16;
constexpr int kStageBytes = BM * 128 * 2;   // hi + lo tile of one K-block
constexpr int kMaxSmem = 232448;            // 227 KB opt-in limit per CTA

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LAB_DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "LAB_DONE:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives when every tcgen05 op previously issued by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns: thread i of the warp gets TMEM lane (base_lane + i)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);   // start address, 16-byte units
    d |= (uint64_t)1 << 16;                         // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                         // layout type: SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulate, both operands K-major
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// 256-bit global stores / loads (sm_100: STG.E.ENL2.256): a lane of the epilogue owns a ROW, so every store
// instruction touches 32 different rows; with 128-bit stores each lane writes half a 32-byte sector per instruction
// and the SM's store path (one sector per cycle) runs at half rate -- measured 800-900 cycles for the six stores of
// one epilogue iteration.  One 256-bit store per lane writes whole sectors.
__device__ __forceinline__ void st_global_256(float* p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld_global_256(const float* p, float (&v)[8]) {
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p)
                 : "memory");
}

__device__ __forceinline__ void split_tf32(float4 v, float4& hi, float4& lo) {
    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    lo.x = v.x - hi.x;
    lo.y = v.y - hi.y;
    lo.z = v.z - hi.z;
    lo.w = v.w - hi.w;
}
// 16-byte chunk `ch` of row `row` inside a [rows x 128 B] swizzled block
__device__ __forceinline__ uint32_t swz(int row, int ch) { return (uint32_t)(row * 128 + ((ch ^ (row & 7)) << 4)); }

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
// Row operand normalised while it is loaded: v = keep/(1-p) * act(scale*(x - am) + bias) with the per-column
// constants of a GraphNorm statistics table (rows 0, 1, 4 of stats[6, k]) and packed keep bits -- the GraphNorm
// "apply" + dropout of impl/models.py:165-166 / 249-251 never materialises its output.
struct NormOp {
    const float* stats;
    const uint32_t* bits;
    float pscale;
    int act;
};

struct TcParams {
    // forward: a1/a2 are the row operand; backward: dout/acts build dP
    const float* a1;
    int64_t lda1;
    int k1;
    const float* a2;
    int64_t lda2;
    int k2;
    const float* w0;
    const float* w1;
    const float* b0;
    const float* b1;
    const uint8_t* mask;
    float z;
    int act;
    int h;
    int64_t n;
    // forward outputs
    float* out;
    int64_t ldo;
    float* acts;
    // backward inputs / outputs
    const float* dout;
    int64_t lddo;
    float* da1;
    int64_t ldda1;
    float* da2;
    int64_t ldda2;
    int kdim;     // reduction length of the MMA (forward: k1+k2, backward: 2h)
    int ndim;     // N of the MMA (forward: 2h, backward: k1+k2)
    int stages;
    uint32_t tmem_cols;
    int rows_per_tile;   // <= BM: rows a tile owns (chosen so that every CTA runs the same number of tiles)
    NormOp n1, n2;       // forward: normalisation applied to a1 / a2 on load (stats == NULL: plain operand)
    int acc1, acc2;      // backward: da1 / da2 += instead of =
    int staged;          // epilogue goes through per-warp shared-memory tiles -> 64-byte row segments per store
    int wide;            // epilogue rows are 32-byte aligned: 256-bit stores
    long long* dbg;      // optional timeline of CTA 0 (GLASS_B200_TC_TIMELINE): [0]=start [1]=setup done,
};                       // [16+i] loader consumed K-block i, [80+i] MMA committed K-block i, [144+t] epilogue done tile t, [200]=end

__device__ __forceinline__ void stamp(long long* dbg, int idx) {
    if (dbg && blockIdx.x == 0) dbg[idx] = clock64();
}

// ACT is a template parameter: with a run-time activation every element of the epilogue / the dP loader carried two
// compare-and-branch pairs (the kernels are bound by the instruction count of their CUDA-core warps, not by memory).
template <bool BWD, bool NORM, int ACT, int EW>
__global__ void __launch_bounds__(kThreads, 1) k_pair_tc(const TcParams P) {
    constexpr int kEpiWarps = EW, kLoadWarps = kWorkWarps - EW, kLoadThreads = kLoadWarps * 32;
    constexpr int kCPT = BM * 8 / kLoadThreads;      // 16-byte chunks per loader thread and K-block
    constexpr int PARTS = EW / 4;                    // epilogue warps per TMEM lane quadrant: they split the columns
    static_assert(EW % 4 == 0 && (BM * 8) % kLoadThreads == 0, "warp split");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = P.kdim, N = P.ndim, H = P.h;
    const int nkb = (K + KBF - 1) / KBF;
    const int b_block = N * 128;                              // bytes of one K-block of the weight operand
    uint8_t* b_hi = smem;
    uint8_t* b_lo = smem + (size_t)nkb * b_block;
    uint8_t* a_ring = smem + (size_t)2 * nkb * b_block;       // stages x (hi 16 KB | lo 16 KB), 1024-aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_ring + (size_t)P.stages * kStageBytes);
    uint64_t* full = bars;                     // [stages]   loader -> mma
    uint64_t* empty = bars + P.stages;         // [stages]   mma -> loader
    uint64_t* tfull = bars + 2 * P.stages;     // [2]        mma -> epilogue
    uint64_t* tempty = tfull + 2;              // [2]        epilogue -> mma
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* s_norm = reinterpret_cast<float*>(tmem_slot + 4);   // NORM: [3][K] scale | am | bias, 16-byte aligned
    float* s_stage = s_norm + (NORM ? 3 * K : 0);              // staged epilogue: [8 warps][arrays][32 rows][16 floats]

    if (threadIdx.x == 0) stamp(P.dbg, 0);
    // ---- one-time setup ------------------------------------------------------------------
    if (NORM) {
        for (int k = threadIdx.x; k < K; k += kThreads) {
            const bool first = k < P.k1;
            const NormOp& op = first ? P.n1 : P.n2;
            const int kk = first ? P.k1 : P.k2, c = first ? k : k - P.k1;
            float sc = 1.f, am = 0.f, bs = 0.f;      // identity: fmaf(1, x - 0, 0) == x
            if (op.stats) {
                sc = op.stats[0 * kk + c];
                am = op.stats[1 * kk + c];
                bs = op.stats[4 * kk + c];
            }
            s_norm[k] = sc;
            s_norm[K + k] = am;
            s_norm[2 * K + k] = bs;
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(smem_u32(full + s), kLoadThreads);
            mbar_init(smem_u32(empty + s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(tfull + a), 1);
            mbar_init(smem_u32(tempty + a), kEpiWarps * 32);
        }
        fence_barrier_init();
    }
    if (warp == kEpiWarps) tmem_alloc(smem_u32(tmem_slot), P.tmem_cols);

    // resident weight operand: row r of the MMA "B" matrix holds the K reduction entries of output column r
    {
        const int chunks = N * nkb * 8;                       // 16-byte chunks
        for (int q = threadIdx.x; q < chunks; q += kThreads) {
            const int ch = q & 7, r = (q >> 3) % N, kb = (q >> 3) / N;
            const int k = kb * KBF + ch * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!BWD) {
                // B[r][k] = (r < H ? W0 : W1)[r % H][k]
                if (k < K) v = ldg_f4((r < H ? P.w0 + (int64_t)r * K : P.w1 + (int64_t)(r - H) * K) + k);
            } else {
                // B[r][j] = Wcat[j][r], j = reduction index over the 2H pre-activation columns
                const int Kw = P.k1 + P.k2;
                float t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = k + u;
                    t[u] = (j < K) ? __ldg((j < H ? P.w0 + (int64_t)j * Kw : P.w1 + (int64_t)(j - H) * Kw) + r) : 0.f;
                }
                v = make_float4(t[0], t[1], t[2], t[3]);
            }
            float4 hi, lo;
            split_tf32(v, hi, lo);
            const uint32_t off = (uint32_t)kb * b_block + swz(r, ch);
            *reinterpret_cast<float4*>(b_hi + off) = hi;
            *reinterpret_cast<float4*>(b_lo + off) = lo;
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) stamp(P.dbg, 1);

    const int RT = P.rows_per_tile;
    const int64_t n_tiles = (P.n + RT - 1) / RT;

    if (warp < kEpiWarps) {
        // ================================ epilogue ==========================================
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(smem_u32(tfull + acc), (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const int quad = warp & 3, half = warp >> 2;         // TMEM lane quadrant (== warp % 4), column part
            const int64_t row = tile * RT + quad * 32 + lane;
            const bool row_ok = row < P.n && quad * 32 + lane < RT;
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * N);
            // Staged epilogue (widths that are multiples of 32): a lane owns a ROW in TMEM, so direct stores put 32
            // different rows in every store instruction (16-byte pieces, one memory request each -- measured
            // 10 K cycles per tile, the bottleneck of these kernels).  Instead each warp passes its 32 x 16 block
            // through a private, XOR-swizzled shared-memory tile (conflict-free both ways) and writes 64-byte
            // row segments: 8 rows per store instruction.
            const int arrays = (!BWD && P.acts) ? 3 : 1;
            float* st = s_stage + warp * (arrays * 512);
            const int wr_sw = (lane >> 1) & 3;                          // write side: row == lane
            bool released = false;
            if (!BWD) {
                float c0 = 0.f, c1 = 0.f;
                if (row_ok) {
                    const uint8_t lab = P.mask[row];
                    c1 = lab ? P.z : 1.f - P.z;
                    c0 = lab ? 1.f - P.z : P.z;
                }
                if (P.staged) {
                    const int c_lo = half * (H / PARTS), c_hi = c_lo + (H / PARTS);
                    for (int cb = c_lo; cb < c_hi; cb += 16) {
                        float p0[16], p1[16];
                        tmem_ld16(t_row + (uint32_t)cb, p0);
                        tmem_ld16(t_row + (uint32_t)(H + cb), p1);
                        tmem_ld_wait();
                        if (cb + 16 >= c_hi) {          // last TMEM read of this tile: hand the accumulator back now
                            tc_fence_before();
                            mbar_arrive(smem_u32(tempty + acc));
                            released = true;
                        }
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch) {
                            const float4 ba = ldg_f4(P.b0 + cb + 4 * ch), bb = ldg_f4(P.b1 + cb + 4 * ch);   // 512 B, L1 resident
                            float4 q0, q1, o4;
                            q0.x = act_fwd(p0[4 * ch] + ba.x, ACT), q0.y = act_fwd(p0[4 * ch + 1] + ba.y, ACT);
                            q0.z = act_fwd(p0[4 * ch + 2] + ba.z, ACT), q0.w = act_fwd(p0[4 * ch + 3] + ba.w, ACT);
                            q1.x = act_fwd(p1[4 * ch] + bb.x, ACT), q1.y = act_fwd(p1[4 * ch + 1] + bb.y, ACT);
                            q1.z = act_fwd(p1[4 * ch + 2] + bb.z, ACT), q1.w = act_fwd(p1[4 * ch + 3] + bb.w, ACT);
                            o4.x = __fadd_rn(__fmul_rn(c1, q1.x), __fmul_rn(c0, q0.x));
                            o4.y = __fadd_rn(__fmul_rn(c1, q1.y), __fmul_rn(c0, q0.y));
                            o4.z = __fadd_rn(__fmul_rn(c1, q1.z), __fmul_rn(c0, q0.z));
                            o4.w = __fadd_rn(__fmul_rn(c1, q1.w), __fmul_rn(c0, q0.w));
                            const int o = lane * 16 + ((ch ^ wr_sw) << 2);
                            *reinterpret_cast<float4*>(st + o) = o4;
                            if (arrays == 3) {
                                *reinterpret_cast<float4*>(st + 512 + o) = q0;
                                *reinterpret_cast<float4*>(st + 1024 + o) = q1;
                            }
                        }
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = i * 8 + (lane >> 2), ch = lane & 3;
                            const int o = r * 16 + ((ch ^ ((r >> 1) & 3)) << 2);
                            const int64_t grow = tile * RT + quad * 32 + r;
                            if (grow < P.n && quad * 32 + r < RT) {
                                *reinterpret_cast<float4*>(P.out + grow * P.ldo + cb + 4 * ch) = *reinterpret_cast<const float4*>(st + o);
                                if (arrays == 3) {
                                    float* pa = P.acts + grow * (2 * (int64_t)H) + cb + 4 * ch;
                                    *reinterpret_cast<float4*>(pa) = *reinterpret_cast<const float4*>(st + 512 + o);
                                    *reinterpret_cast<float4*>(pa + H) = *reinterpret_cast<const float4*>(st + 1024 + o);
                                }
                            }
                        }
                        __syncwarp();
                    }
                } else {
                // software pipeline: the TMEM read of the NEXT 8-column block is in flight while this one is computed
                // and stored (a tcgen05.ld round trip measured ~370 cycles per iteration when taken serially)
                float n0[8], n1[8];
                if (half * 8 < H) {
                    tmem_ld8(t_row + (uint32_t)(half * 8), n0);
                    tmem_ld8(t_row + (uint32_t)(H + half * 8), n1);
                }
                for (int c = half * 8; c < H; c += PARTS * 8) {
                    const bool dbg_on = P.dbg && threadIdx.x == 0 && it == 1 && c < 2 * PARTS * 8;     // CTA 0 / warp 0, second tile
                    if (dbg_on) stamp(P.dbg, 210 + (c / (PARTS * 8)) * 4);
                    const float4 ba0 = ldg_f4(P.b0 + c), ba1 = ldg_f4(P.b0 + c + 4);   // biases: 512 B, L1 resident
                    const float4 bb0 = ldg_f4(P.b1 + c), bb1 = ldg_f4(P.b1 + c + 4);
                    const float bA[8] = {ba0.x, ba0.y, ba0.z, ba0.w, ba1.x, ba1.y, ba1.z, ba1.w};
                    const float bB[8] = {bb0.x, bb0.y, bb0.z, bb0.w, bb1.x, bb1.y, bb1.z, bb1.w};
                    tmem_ld_wait();
                    float p0[8], p1[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) p0[u] = n0[u], p1[u] = n1[u];
                    if (c + PARTS * 8 < H) {
                        tmem_ld8(t_row + (uint32_t)(c + PARTS * 8), n0);
                        tmem_ld8(t_row + (uint32_t)(H + c + PARTS * 8), n1);
                    } else {                        // last TMEM read of this tile: hand the accumulator back now
                        tc_fence_before();
                        mbar_arrive(smem_u32(tempty + acc));
                        released = true;
                    }
                    if (dbg_on) stamp(P.dbg, 211 + (c / (PARTS * 8)) * 4);
                    if (row_ok) {
                        float o[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            p0[u] = act_fwd(p0[u] + bA[u], ACT);
                            p1[u] = act_fwd(p1[u] + bB[u], ACT);
                            o[u] = __fadd_rn(__fmul_rn(c1, p1[u]), __fmul_rn(c0, p0[u]));
                        }
                        if (dbg_on) stamp(P.dbg, 212 + (c / (PARTS * 8)) * 4);
                        if (P.wide) {
                            st_global_256(P.out + row * P.ldo + c, o);
                            if (P.acts) {
                                st_global_256(P.acts + row * (2 * (int64_t)H) + c, p0);
                                st_global_256(P.acts + row * (2 * (int64_t)H) + H + c, p1);
                            }
                        } else {
                            float4* po = reinterpret_cast<float4*>(P.out + row * P.ldo + c);
                            po[0] = make_float4(o[0], o[1], o[2], o[3]);
                            po[1] = make_float4(o[4], o[5], o[6], o[7]);
                            if (P.acts) {
                                float4* pa = reinterpret_cast<float4*>(P.acts + row * (2 * (int64_t)H) + c);
                                pa[0] = make_float4(p0[0], p0[1], p0[2], p0[3]);
                                pa[1] = make_float4(p0[4], p0[5], p0[6], p0[7]);
                                float4* pb = reinterpret_cast<float4*>(P.acts + row * (2 * (int64_t)H) + H + c);
                                pb[0] = make_float4(p1[0], p1[1], p1[2], p1[3]);
                                pb[1] = make_float4(p1[4], p1[5], p1[6], p1[7]);
                            }
                        }
                        if (dbg_on) stamp(P.dbg, 213 + (c / (PARTS * 8)) * 4);
                    }
                }
                }
            } else if (P.staged) {
                const int c_lo = half * (N / PARTS), c_hi = c_lo + (N / PARTS);
                for (int cb = c_lo; cb < c_hi; cb += 16) {
                    float d[16];
                    tmem_ld16(t_row + (uint32_t)cb, d);
                    tmem_ld_wait();
                    if (cb + 16 >= c_hi) {
                        tc_fence_before();
                        mbar_arrive(smem_u32(tempty + acc));
                        released = true;
                    }
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch)
                        *reinterpret_cast<float4*>(st + lane * 16 + ((ch ^ wr_sw) << 2)) = make_float4(d[4 * ch], d[4 * ch + 1], d[4 * ch + 2], d[4 * ch + 3]);
                    __syncwarp();
                    const bool first = cb < P.k1;                        // k1 % 16 == 0: a block never straddles a1 | a2
                    float* base = first ? P.da1 : P.da2;
                    const int64_t ldd = first ? P.ldda1 : P.ldda2;
                    const int cc = first ? cb : cb - P.k1;
                    const int accum = first ? P.acc1 : P.acc2;
                    if (base) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = i * 8 + (lane >> 2), ch = lane & 3;
                            const int64_t grow = tile * RT + quad * 32 + r;
                            if (grow < P.n && quad * 32 + r < RT) {
                                float4 v = *reinterpret_cast<const float4*>(st + r * 16 + ((ch ^ ((r >> 1) & 3)) << 2));
                                float4* dst = reinterpret_cast<float4*>(base + grow * ldd + cc + 4 * ch);
                                if (accum) {     // the operand also fed another GEMM: add to its gradient
                                    const float4 o = *dst;
                                    v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
                                }
                                *dst = v;
                            }
                        }
                    }
                    __syncwarp();
                }
            } else {
                float nd[8];
                if (half * 8 < N) tmem_ld8(t_row + (uint32_t)(half * 8), nd);
                for (int c = half * 8; c < N; c += PARTS * 8) {
                    tmem_ld_wait();
                    float d[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) d[u] = nd[u];
                    if (c + PARTS * 8 < N) {
                        tmem_ld8(t_row + (uint32_t)(c + PARTS * 8), nd);
                    } else {
                        tc_fence_before();
                        mbar_arrive(smem_u32(tempty + acc));
                        released = true;
                    }
                    if (row_ok) {
                        float* dst = nullptr;
                        if (c < P.k1) {
                            if (P.da1) dst = P.da1 + row * P.ldda1 + c;
                        } else if (P.da2) {
                            dst = P.da2 + row * P.ldda2 + (c - P.k1);
                        }
                        if (dst && P.wide) {
                            if (c < P.k1 ? P.acc1 : P.acc2) {      // the operand also fed another GEMM: add to its gradient
                                float o[8];
                                ld_global_256(dst, o);
#pragma unroll
                                for (int u = 0; u < 8; ++u) d[u] += o[u];
                            }
                            st_global_256(dst, d);
                        } else if (dst) {
                            if (c < P.k1 ? P.acc1 : P.acc2) {
                                const float4 o0 = reinterpret_cast<const float4*>(dst)[0], o1 = reinterpret_cast<const float4*>(dst)[1];
                                d[0] += o0.x, d[1] += o0.y, d[2] += o0.z, d[3] += o0.w;
                                d[4] += o1.x, d[5] += o1.y, d[6] += o1.z, d[7] += o1.w;
                            }
                            reinterpret_cast<float4*>(dst)[0] = make_float4(d[0], d[1], d[2], d[3]);
                            reinterpret_cast<float4*>(dst)[1] = make_float4(d[4], d[5], d[6], d[7]);
                        }
                    }
                }
            }
            if (!released) {
                tc_fence_before();
                mbar_arrive(smem_u32(tempty + acc));
            }
            if (threadIdx.x == 0 && it < 50) stamp(P.dbg, 144 + it);
        }
    } else if (warp == kEpiWarps) {
        // ================================ MMA issuer ========================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BM, N);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(smem_u32(tempty + acc), (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(smem_u32(full + stage), phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(a_ring + (size_t)stage * kStageBytes);
                    const uint32_t a_lo = a_hi + BM * 128;
                    const uint32_t bh = smem_u32(b_hi + (size_t)kb * b_block);
                    const uint32_t bl = smem_u32(b_lo + (size_t)kb * b_block);
                    const int ksteps = min(KBF, K - kb * KBF) / 8;        // K % 8 == 0
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint64_t dah = make_smem_desc(a_hi + ks * 32), dal = make_smem_desc(a_lo + ks * 32);
                        const uint64_t dbh = make_smem_desc(bh + ks * 32), dbl = make_smem_desc(bl + ks * 32);
                        umma_tf32(d_tmem, dal, dbh, idesc, (kb | ks) ? 1u : 0u);   // small terms first
                        umma_tf32(d_tmem, dah, dbl, idesc, 1u);
                        umma_tf32(d_tmem, dah, dbh, idesc, 1u);
                    }
                    umma_commit(smem_u32(empty + stage));                  // frees the ring slot when the MMAs retire
                    if (it * nkb + kb < 64) stamp(P.dbg, 80 + it * nkb + kb);
                    if (++stage == P.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(smem_u32(tfull + acc));                        // accumulator complete -> epilogue
            }
        }
    } else {
        // ================================ operand loader ====================================
        const int lt = threadIdx.x - (kEpiWarps + 1) * 32;                 // 0 .. kLoadThreads-1
        const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const int64_t total = my_tiles * nkb;                              // K-blocks this CTA streams
        constexpr int D = (BWD || kCPT > 2) ? 2 : 4;                       // register ring depth (K-blocks in flight)
        struct Raw {
            float4 g[kCPT];
            float4 a[BWD ? kCPT : 1];
            uint32_t lab[(BWD || NORM) ? kCPT : 1];   // BWD: raw label byte, NORM: raw keep-bit word; decoded only when
        };                                              // consumed, so that issuing a stage never waits on a load
        Raw ring[D];

        auto issue = [&](int64_t seq, Raw& rw) {
            const int64_t tile = blockIdx.x + (seq / nkb) * gridDim.x;
            const int kb = (int)(seq % nkb);
#pragma unroll
            for (int i = 0; i < kCPT; ++i) {
                const int q = lt + i * kLoadThreads;
                const int r = q >> 3, ch = q & 7;
                const int64_t row = r < RT ? tile * RT + r : P.n;   // rows past the tile's share load nothing
                const int k = kb * KBF + ch * 4;
                rw.g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (BWD) {
                    rw.a[i] = make_float4(1.f, 1.f, 1.f, 1.f);
                    rw.lab[i] = 2u;              // 2 = row / column out of range -> coefficient 0
                }
                if (NORM) rw.lab[i] = 0xffffffffu;
                if (row < P.n && k < K) {
                    if (!BWD) {
                        rw.g[i] = (k < P.k1) ? ldg_f4(P.a1 + row * P.lda1 + k) : ldg_f4(P.a2 + row * P.lda2 + (k - P.k1));
                        if (NORM) {     // keep bits of this chunk's 32-column word (operand widths are multiples of 32)
                            const bool first = k < P.k1;
                            const uint32_t* bits = first ? P.n1.bits : P.n2.bits;
                            const int kk = first ? P.k1 : P.k2, c = first ? k : k - P.k1;
                            if (bits) rw.lab[i] = __ldg(bits + row * (kk >> 5) + (c >> 5));
                        }
                    } else {
                        // dP[row][k..k+3]: k indexes the 2H pre-activation columns (branch 0 | branch 1)
                        const int br = k >= H;
                        rw.g[i] = ldg_f4(P.dout + row * P.lddo + (k - br * H));
                        if (ACT != GLASS_ACT_NONE) rw.a[i] = ldg_f4(P.acts + row * (2 * (int64_t)H) + k);
                        rw.lab[i] = P.mask[row];
                    }
                }
            }
        };
        int stage = 0, nconsumed = 0;
        uint32_t phase = 0;
        auto consume = [&](const Raw& rw, int kb_of_slot) {
            (void)kb_of_slot;
            mbar_wait(smem_u32(empty + stage), phase ^ 1);
            uint8_t* dst = a_ring + (size_t)stage * kStageBytes;
#pragma unroll
            for (int i = 0; i < kCPT; ++i) {
                const int q = lt + i * kLoadThreads;
                const uint32_t off = swz(q >> 3, q & 7);
                float4 v = rw.g[i];
                if (NORM) {
                    const int k = kb_of_slot * KBF + (q & 7) * 4;
                    const NormOp& op = k < P.k1 ? P.n1 : P.n2;
                    if (op.stats && k < K) {
                        const float4 sc = *reinterpret_cast<const float4*>(s_norm + k);
                        const float4 am = *reinterpret_cast<const float4*>(s_norm + K + k);
                        const float4 bs = *reinterpret_cast<const float4*>(s_norm + 2 * K + k);
                        const uint32_t nib = rw.lab[i] >> ((q & 7) * 4);
                        v.x = fmaf(sc.x, v.x - am.x, bs.x), v.y = fmaf(sc.y, v.y - am.y, bs.y);
                        v.z = fmaf(sc.z, v.z - am.z, bs.z), v.w = fmaf(sc.w, v.w - am.w, bs.w);
                        if (op.act != GLASS_ACT_NONE) {      // one uniform branch per chunk, not two per element
                            v.x = act_fwd(v.x, op.act), v.y = act_fwd(v.y, op.act);
                            v.z = act_fwd(v.z, op.act), v.w = act_fwd(v.w, op.act);
                        }
                        v.x *= (nib & 1u) ? op.pscale : 0.f, v.y *= (nib & 2u) ? op.pscale : 0.f;
                        v.z *= (nib & 4u) ? op.pscale : 0.f, v.w *= (nib & 8u) ? op.pscale : 0.f;
                    }
                }
                if (BWD) {
                    // chunk q & 7 of this K-block covers pre-activation columns of ONE branch (H % 4 == 0)
                    const int br = (kb_of_slot * KBF + (q & 7) * 4) >= H;
                    const float c = rw.lab[i] > 1u ? 0.f : (((rw.lab[i] != 0) == (br != 0)) ? P.z : 1.f - P.z);
                    v.x *= c, v.y *= c, v.z *= c, v.w *= c;
                    if (ACT != GLASS_ACT_NONE) {
                        v.x *= act_grad_from_out(rw.a[i].x, ACT);
                        v.y *= act_grad_from_out(rw.a[i].y, ACT);
                        v.z *= act_grad_from_out(rw.a[i].z, ACT);
                        v.w *= act_grad_from_out(rw.a[i].w, ACT);
                    }
                }
                float4 hi, lo;
                split_tf32(v, hi, lo);
                *reinterpret_cast<float4*>(dst + off) = hi;
                *reinterpret_cast<float4*>(dst + BM * 128 + off) = lo;
            }
            fence_proxy_async();
            mbar_arrive(smem_u32(full + stage));
            if (lt == 0 && nconsumed < 64) stamp(P.dbg, 16 + nconsumed);
            ++nconsumed;
            if (++stage == P.stages) {
                stage = 0;
                phase ^= 1;
            }
        };
#pragma unroll
        for (int d = 0; d < D; ++d)
            if (d < total) issue(d, ring[d]);
        for (int64_t seq = 0; seq < total; seq += D) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                if (seq + d < total) {
                    consume(ring[d], (int)((seq + d) % nkb));
                    if (seq + d + D < total) issue(seq + d + D, ring[d]);
                }
            }
        }
    }

    // ---- teardown -----------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kEpiWarps) tmem_dealloc(tmem_base, P.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// dW / db on tensor cores:  dWcat[2H x K] = dP^T[2H x n] * A[n x K],  db = column sums of dP.
// The reduction runs over graph rows, so both operands are MN-major in their natural row-major global
// layout (dP[i][j], A[i][k]: the non-reduction index is contiguous).  MN-major 32-bit operands must use the
// SWIZZLE_128B_BASE32B layout: atoms of 32 elements (128 B) x 4 reduction rows (512 B) whose 32-byte chunks
// are XOR-swizzled with the row index (Swizzle<2,5,2> on the byte address).  Each stage holds 32 rows i:
//   8 row-groups per atom column (SBO = 512), atom columns 4096 B apart (LBO = 4096); one K = 8 MMA
//   consumes two row-groups, i.e. advances the start address by 1024 B
// (canonical layout ((T,8,m),(4,k)):((1,T,LBO),(8T,SBO)) of the UMMA MN-major descriptor).
// CTA s reduces rows [s*R, (s+1)*R) into one TMEM accumulator and writes part[s]; a second kernel sums
// the partials in CTA order (deterministic).  21 warps: 0-3 epilogue, 4 MMA, 5-20 loaders.
// ---------------------------------------------------------------------------------------------
constexpr int kDwLoadWarps = 16, kDwLoadThreads = kDwLoadWarps * 32, kDwEpiWarps = 4;
constexpr int kDwThreads = (kDwLoadWarps + 1) * 32;
constexpr int kDwMmaWarp = kDwLoadWarps;
constexpr int kDwRows = 32;                    // reduction rows per stage

struct DwParams {
    const float* dout;
    int64_t lddo;
    const float* acts;
    const uint8_t* mask;
    float z;
    int act;
    int h;
    const float* a1;
    int64_t lda1;
    int k1;
    const float* a2;
    int64_t lda2;
    int k2;
    int64_t n;
    float* part;          // [splits][2h][part_ld]; column K holds the db partial
    int part_ld;
    int64_t rows_per_cta; // multiple of 32
    int stages;
    uint32_t tmem_cols;
    NormOp n1, n2;        // a1 / a2 normalised on load, exactly as the forward loader does (stats == NULL: plain)
};

__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
    d |= (uint64_t)(4096 >> 4) << 16;               // LBO: next 32-element atom column
    d |= (uint64_t)(512 >> 4) << 32;                // SBO: next group of 4 reduction rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                         // SWIZZLE_128B_BASE32B (the only MN-major layout for tf32)
    return d;
}
// element chunk (4 consecutive MN elements starting at `mn`, reduction row `i` of the stage)
__device__ __forceinline__ uint32_t swz_mn(int mn, int i) {
    const int c16 = (mn & 31) >> 2;                 // 16-byte chunk inside the 128-byte atom row
    return (uint32_t)((mn >> 5) * 4096 + (i >> 2) * 512 + (i & 3) * 128 + ((((c16 >> 1) ^ (i & 3)) << 5) | ((c16 & 1) << 4)));
}

template <bool NORM, int ACT>
__global__ void __launch_bounds__(kDwThreads, 1) k_pair_dw_tc(const DwParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = P.h, K = P.k1 + P.k2;
    const int j0 = blockIdx.y * 128;                          // 128-row slice of the 2H outputs
    const int a_bytes = 128 * 128;                            // 128 j x 32 rows x 4 B
    const int b_bytes = K * 128;
    const int stage_bytes = 2 * a_bytes + 2 * b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)P.stages * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + P.stages;
    uint64_t* tfull = bars + 2 * P.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    float* s_db = reinterpret_cast<float*>(tmem_slot + 4);   // [16][128]
    float* s_norm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s_db + 16 * 128) + 15) & ~uintptr_t(15));   // NORM: [3][K]

    if (NORM) {
        for (int k = threadIdx.x; k < K; k += kDwThreads) {
            const bool first = k < P.k1;
            const NormOp& op = first ? P.n1 : P.n2;
            const int kk = first ? P.k1 : P.k2, c = first ? k : k - P.k1;
            float sc = 1.f, am = 0.f, bs = 0.f;
            if (op.stats) {
                sc = op.stats[0 * kk + c];
                am = op.stats[1 * kk + c];
                bs = op.stats[4 * kk + c];
            }
            s_norm[k] = sc;
            s_norm[K + k] = am;
            s_norm[2 * K + k] = bs;
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(smem_u32(full + s), kDwLoadThreads);
            mbar_init(smem_u32(empty + s), 1);
        }
        mbar_init(smem_u32(tfull), 1);
        fence_barrier_init();
    }
    if (warp == kDwMmaWarp) tmem_alloc(smem_u32(tmem_slot), P.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t m_lo = (int64_t)blockIdx.x * P.rows_per_cta;
    const int64_t m_hi = min(m_lo + P.rows_per_cta, P.n);
    const int64_t total = m_hi > m_lo ? (m_hi - m_lo + kDwRows - 1) / kDwRows : 0;   // stages to stream

    if (warp == kDwMmaWarp) {
        // -------- MMA issuer --------
        if (lane == 0 && total > 0) {
            const uint32_t idesc = make_idesc(128, K) | (1u << 15) | (1u << 16);   // both operands MN-major
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t st = 0; st < total; ++st) {
                mbar_wait(smem_u32(full + stage), phase);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t a_lo = a_hi + a_bytes;
                const uint32_t b_hi = a_lo + a_bytes;
                const uint32_t b_lo = b_hi + b_bytes;
#pragma unroll
                for (int kg = 0; kg < 4; ++kg) {
                    const uint64_t dah = make_smem_desc_mn(a_hi + kg * 1024), dal = make_smem_desc_mn(a_lo + kg * 1024);
                    const uint64_t dbh = make_smem_desc_mn(b_hi + kg * 1024), dbl = make_smem_desc_mn(b_lo + kg * 1024);
                    umma_tf32(tmem_base, dal, dbh, idesc, (st | kg) ? 1u : 0u);
                    umma_tf32(tmem_base, dah, dbl, idesc, 1u);
                    umma_tf32(tmem_base, dah, dbh, idesc, 1u);
                }
                umma_commit(smem_u32(empty + stage));
                if (++stage == P.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(smem_u32(tfull));
        }
    } else {
        // -------- loaders: straight row-major copies into the MN-major tiles --------
        const int lt = threadIdx.x;                                  // 0 .. 511
        const int jc = (lt & 31) * 4;                                // this thread's 4 dP columns (fixed)
        const int j = j0 + jc;
        const int br = j >= H;
        const int kchunks = K >> 2;                                  // float4 chunks per A row
        float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
        struct Raw {
            float4 g[2], a[2], x[2];
            uint32_t lab[2];             // raw label byte (2 = row out of range); coefficient derived when consumed
            uint32_t xw[NORM ? 2 : 1];   // NORM: keep-bit word of the x chunk (0 for rows out of range)
        };
        Raw ring[2];
        auto issue = [&](int64_t st, Raw& rw) {
            const int64_t r0 = m_lo + st * kDwRows;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int i = (lt >> 5) + 16 * t;
                const int64_t row = r0 + i;
                rw.g[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                rw.a[t] = make_float4(1.f, 1.f, 1.f, 1.f);
                rw.lab[t] = 2u;
                if (row < m_hi) {
                    rw.g[t] = ldg_f4(P.dout + row * P.lddo + (j - br * H));
                    if (ACT != GLASS_ACT_NONE) rw.a[t] = ldg_f4(P.acts + row * (2 * (int64_t)H) + j);
                    rw.lab[t] = P.mask[row];
                }
                const int q = lt + t * kDwLoadThreads;
                rw.x[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (NORM) rw.xw[t] = 0u;
                if (q < kDwRows * kchunks) {
                    const int ib = q / kchunks, k = (q % kchunks) * 4;
                    const int64_t rowb = r0 + ib;
                    if (rowb < m_hi) {
                        rw.x[t] = (k < P.k1) ? ldg_f4(P.a1 + rowb * P.lda1 + k) : ldg_f4(P.a2 + rowb * P.lda2 + (k - P.k1));
                        if (NORM) {
                            const bool first = k < P.k1;
                            const uint32_t* bits = first ? P.n1.bits : P.n2.bits;
                            const int kk = first ? P.k1 : P.k2, c = first ? k : k - P.k1;
                            rw.xw[t] = bits ? __ldg(bits + rowb * (kk >> 5) + (c >> 5)) : 0xffffffffu;
                        }
                    }
                }
            }
        };
        int stage = 0;
        uint32_t phase = 0;
        auto consume = [&](const Raw& rw) {
            mbar_wait(smem_u32(empty + stage), phase ^ 1);
            uint8_t* base = smem + (size_t)stage * stage_bytes;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int i = (lt >> 5) + 16 * t;
                float4 v = rw.g[t];
                const float c = rw.lab[t] > 1u ? 0.f : (((rw.lab[t] != 0) == (br != 0)) ? P.z : 1.f - P.z);
                v.x *= c, v.y *= c, v.z *= c, v.w *= c;
                if (ACT != GLASS_ACT_NONE) {
                    v.x *= act_grad_from_out(rw.a[t].x, ACT);
                    v.y *= act_grad_from_out(rw.a[t].y, ACT);
                    v.z *= act_grad_from_out(rw.a[t].z, ACT);
                    v.w *= act_grad_from_out(rw.a[t].w, ACT);
                }
                bsum.x += v.x, bsum.y += v.y, bsum.z += v.z, bsum.w += v.w;
                float4 hi, lo;
                split_tf32(v, hi, lo);
                const uint32_t off = swz_mn(jc, i);
                *reinterpret_cast<float4*>(base + off) = hi;
                *reinterpret_cast<float4*>(base + a_bytes + off) = lo;
                const int q = lt + t * kDwLoadThreads;
                if (q < kDwRows * kchunks) {
                    const int ib = q / kchunks, k = (q % kchunks) * 4;
                    float4 xv = rw.x[t];
                    if (NORM) {
                        const NormOp& op = k < P.k1 ? P.n1 : P.n2;
                        if (op.stats) {
                            const float4 sc = *reinterpret_cast<const float4*>(s_norm + k);
                            const float4 am = *reinterpret_cast<const float4*>(s_norm + K + k);
                            const float4 bs = *reinterpret_cast<const float4*>(s_norm + 2 * K + k);
                            const uint32_t nib = rw.xw[t] >> ((k < P.k1 ? k : k - P.k1) & 31);
                            xv.x = fmaf(sc.x, xv.x - am.x, bs.x), xv.y = fmaf(sc.y, xv.y - am.y, bs.y);
                            xv.z = fmaf(sc.z, xv.z - am.z, bs.z), xv.w = fmaf(sc.w, xv.w - am.w, bs.w);
                            if (op.act != GLASS_ACT_NONE) {
                                xv.x = act_fwd(xv.x, op.act), xv.y = act_fwd(xv.y, op.act);
                                xv.z = act_fwd(xv.z, op.act), xv.w = act_fwd(xv.w, op.act);
                            }
                            xv.x *= (nib & 1u) ? op.pscale : 0.f, xv.y *= (nib & 2u) ? op.pscale : 0.f;
                            xv.z *= (nib & 4u) ? op.pscale : 0.f, xv.w *= (nib & 8u) ? op.pscale : 0.f;
                        }
                    }
                    split_tf32(xv, hi, lo);
                    const uint32_t offb = swz_mn(k, ib);
                    *reinterpret_cast<float4*>(base + 2 * a_bytes + offb) = hi;
                    *reinterpret_cast<float4*>(base + 2 * a_bytes + b_bytes + offb) = lo;
                }
            }
            fence_proxy_async();
            mbar_arrive(smem_u32(full + stage));
            if (++stage == P.stages) {
                stage = 0;
                phase ^= 1;
            }
        };
        if (0 < total) issue(0, ring[0]);
        if (1 < total) issue(1, ring[1]);
        for (int64_t st = 0; st < total; st += 2) {
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                if (st + d < total) {
                    consume(ring[d]);
                    if (st + d + 2 < total) issue(st + d + 2, ring[d]);
                }
            }
        }
        float* sd = s_db + (lt >> 5) * 128 + jc;
        sd[0] = bsum.x, sd[1] = bsum.y, sd[2] = bsum.z, sd[3] = bsum.w;
        if (warp < kDwEpiWarps) {
        // -------- epilogue (after the whole reduction): loader warps 0-3 own TMEM lane quadrants 0-3 --------
        if (total > 0) {
            mbar_wait(smem_u32(tfull), 0);
            tc_fence_after();
        }
        const int j = warp * 32 + lane;
        float* dst = P.part + ((int64_t)blockIdx.x * (2 * H) + j0 + j) * P.part_ld;
        const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < K; c += 8) {
            float d[8];
            if (total > 0) {
                tmem_ld8(t_row + (uint32_t)c, d);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) d[u] = 0.f;
            }
            st_global_256(dst + c, d);              // part_ld % 8 == 0 and the workspace is 256-byte aligned
        }
        tc_fence_before();
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kDwMmaWarp) tmem_dealloc(tmem_base, P.tmem_cols);
    if (threadIdx.x < 128) {   // db partial: fixed-order sum over the 16 loader row lanes
        float s = 0.f;
        for (int t = 0; t < 16; ++t) s += s_db[t * 128 + threadIdx.x];
        P.part[((int64_t)blockIdx.x * (2 * H) + j0 + threadIdx.x) * P.part_ld + K] = s;
    }
}

// 256 threads = 32 output elements x 8 split lanes; lane q sums splits q, q+8, ... in order, the 8 lane
// partials are added in lane order (deterministic).
__global__ void __launch_bounds__(256) k_pair_dw_tc_reduce(const float* __restrict__ part, int splits, int h, int K,
                                                           int part_ld, float* __restrict__ dw0, float* __restrict__ db0,
                                                           float* __restrict__ dw1, float* __restrict__ db1) {
    __shared__ float sm[8][33];
    const int J = 2 * h;
    const int el = threadIdx.x & 31, q = threadIdx.x >> 5;
    const int64_t e = blockIdx.x * 32ll + el;
    const bool ok = e < (int64_t)J * (K + 1);
    const int j = ok ? (int)(e / (K + 1)) : 0, k = ok ? (int)(e % (K + 1)) : 0;
    float s = 0.f;
    if (ok)
        for (int sp = q; sp < splits; sp += 8) s += part[((int64_t)sp * J + j) * part_ld + k];
    sm[q][el] = s;
    __syncthreads();
    if (q == 0 && ok) {
        float t = sm[0][el];
#pragma unroll
        for (int i = 1; i < 8; ++i) t += sm[i][el];
        const int jr = j < h ? j : j - h;
        if (k < K) (j < h ? dw0 : dw1)[(int64_t)jr * K + k] = t;
        else (j < h ? db0 : db1)[jr] = t;
    }
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// shared-memory plan; returns false when the shape does not fit
bool plan(int kdim, int ndim, bool norm, int stage_arrays, int epi_warps, int* stages, size_t* bytes, uint32_t* tmem_cols) {
    if (ndim < 16 || ndim > 256 || ndim % 16 || kdim < 8 || kdim % 8 || kdim > 512) return false;
    const int nkb = (kdim + KBF - 1) / KBF;
    const size_t b = (size_t)2 * nkb * ndim * 128;
    const size_t fixed = b + 1024 /*alignment slack*/ + 256 /*barriers, TMEM slot*/ +
                         (norm ? (size_t)3 * kdim * sizeof(float) : 0) /*operand normalisation constants*/ +
                         (size_t)stage_arrays * epi_warps * 512 * sizeof(float) /*staged epilogue tiles*/;
    if (fixed + 2 * (size_t)kStageBytes > (size_t)kMaxSmem) return false;
    int s = (int)(((size_t)kMaxSmem - fixed) / kStageBytes);
    if (s > 6) s = 6;
    *stages = s;
    *bytes = fixed + (size_t)s * kStageBytes;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * ndim)) cols <<= 1;
    if (cols > 512) return false;
    *tmem_cols = cols;
    return true;
}

template <bool BWD, bool NORM, int ACT, int EW>
int launch_act(TcParams& P, cudaStream_t st);

// Epilogue warps of the forward kernel: 16 when the epilogue sets the pace (saved activations: three output arrays and
// the activation math per tile), else 8.  GLASS_B200_TC_EPI=8|16 overrides (measurements).
inline int fwd_epi_warps(const TcParams& P) {
    static const int force = [] {
        const char* e = getenv("GLASS_B200_TC_EPI");
        return e ? atoi(e) : 0;
    }();
    if (force == 8 || force == 16) return force;
    return (P.h % 32 == 0) ? 16 : 8;
}

template <bool BWD, bool NORM>
int launch(TcParams& P, cudaStream_t st) {
    // backward without saved activations behaves like ACT_NONE (the activation derivative is 1)
    const int act = (BWD && !P.acts) ? GLASS_ACT_NONE : P.act;
    if (!BWD && fwd_epi_warps(P) == 16) {
        if (act == GLASS_ACT_ELU) return launch_act<BWD, NORM, GLASS_ACT_ELU, BWD ? 8 : 16>(P, st);
        if (act == GLASS_ACT_RELU) return launch_act<BWD, NORM, GLASS_ACT_RELU, BWD ? 8 : 16>(P, st);
        return launch_act<BWD, NORM, GLASS_ACT_NONE, BWD ? 8 : 16>(P, st);
    }
    if (act == GLASS_ACT_ELU) return launch_act<BWD, NORM, GLASS_ACT_ELU, 8>(P, st);
    if (act == GLASS_ACT_RELU) return launch_act<BWD, NORM, GLASS_ACT_RELU, 8>(P, st);
    return launch_act<BWD, NORM, GLASS_ACT_NONE, 8>(P, st);
}

template <bool BWD, bool NORM, int ACT, int EW>
int launch_act(TcParams& P, cudaStream_t st) {
    size_t bytes = 0;
    // staged epilogue: output width (forward: h, backward: k1 + k2 with the a1|a2 boundary on a 16-column block)
    const int out_w = BWD ? P.ndim : P.h;
    // (measured: no faster than the direct stores -- the epilogue is not bound by its store pattern -- so it is
    // opt-in, kept for experiments: GLASS_B200_TC_STAGED=1)
    static const bool want_stage = getenv("GLASS_B200_TC_STAGED") != nullptr;
    P.staged = want_stage && (out_w % (16 * (EW / 4)) == 0) && (!BWD || P.k1 % 16 == 0);
    int stage_arrays = P.staged ? ((!BWD && P.acts) ? 3 : 1) : 0;
    if (P.staged && (!plan(P.kdim, P.ndim, NORM, stage_arrays, EW, &P.stages, &bytes, &P.tmem_cols) || P.stages < 2)) {
        P.staged = 0;
        stage_arrays = 0;
    }
    if (!plan(P.kdim, P.ndim, NORM, stage_arrays, EW, &P.stages, &bytes, &P.tmem_cols)) {
        set_error("pair_linear_mix (tcgen05): shape k=%d n=%d does not fit", P.kdim, P.ndim);
        return GLASS_ERR_UNSUPPORTED;
    }
    static bool attr_done = false;
    if (!attr_done) {
        GLASS_CUDA(cudaFuncSetAttribute(k_pair_tc<BWD, NORM, ACT, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        attr_done = true;
    }
    // equal work per CTA: with 128-row tiles 57,333 rows are 448 tiles = 3.03 per SM, i.e. four rounds for a
    // few CTAs; shrink the rows a tile owns so that every CTA runs exactly ceil(tiles / #SMs) tiles
    int grid = sm_count();
    const int64_t full_tiles = ceil_div(P.n, BM);
    if (grid > full_tiles) grid = (int)full_tiles;
    const int64_t rounds = ceil_div(full_tiles, grid);
    int rt = (int)ceil_div(P.n, (int64_t)grid * rounds);
    if (rt > BM) rt = BM;
    P.rows_per_tile = rt;
    const int64_t tiles = ceil_div(P.n, rt);
    if (grid > tiles) grid = (int)tiles;
    static const bool timeline = getenv("GLASS_B200_TC_TIMELINE") != nullptr;
    long long* dbg = nullptr;
    if (timeline) {
        GLASS_CUDA(cudaMalloc(&dbg, 256 * sizeof(long long)));
        GLASS_CUDA(cudaMemset(dbg, 0, 256 * sizeof(long long)));
        P.dbg = dbg;
    }
    k_pair_tc<BWD, NORM, ACT, EW><<<grid, kThreads, bytes, st>>>(P);
    GLASS_LAUNCH_CHECK();
    if (timeline) {   // debugging aid only: synchronises and prints CTA 0's event times in SM cycles
        long long h[256];
        GLASS_CUDA(cudaStreamSynchronize(st));
        GLASS_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(dbg);
        fprintf(stderr, "[tc timeline bwd=%d k=%d n=%d rt=%d stages=%d epi_warps=%d] setup %lld end %lld\n  loader:", (int)BWD, P.kdim,
                P.ndim, P.rows_per_tile, P.stages, EW, h[1] - h[0], h[200] - h[0]);
        for (int i = 0; i < 64 && h[16 + i]; ++i) fprintf(stderr, " %lld", h[16 + i] - h[0]);
        fprintf(stderr, "\n  mma:");
        for (int i = 0; i < 64 && h[80 + i]; ++i) fprintf(stderr, " %lld", h[80 + i] - h[0]);
        fprintf(stderr, "\n  epilogue:");
        for (int i = 0; i < 50 && h[144 + i]; ++i) fprintf(stderr, " %lld", h[144 + i] - h[0]);
        fprintf(stderr, "\n  epilogue detail (tile 1, warp 0; start / tmem ready / math done / stores issued):");
        for (int i = 210; i < 218; ++i) fprintf(stderr, " %lld", h[i] ? h[i] - h[0] : 0);
        fprintf(stderr, "\n");
    }
    return GLASS_OK;
}

}  // namespace

// Shapes the tcgen05 path accepts.  Called with (k1, k2, h) of the forward problem.
bool pair_tc_supported(int k1, int k2, int h, int64_t lda1, int64_t lda2, const void* a1, const void* a2) {
    if (h % 8 || h < 8 || h > 128) return false;                 // N = 2h multiple of 16, <= 256
    if (k1 % 8 || k2 % 8 || k1 < 8) return false;                // float4 chunks never straddle a1|a2; K % 8 == 0
    if (lda1 % 4 || !aligned16(a1)) return false;
    if (k2 && (lda2 % 4 || !aligned16(a2))) return false;
    int s;
    size_t b;
    uint32_t c;
    return plan(k1 + k2, 2 * h, false, 0, 8, &s, &b, &c) && plan(2 * h, k1 + k2, false, 0, 8, &s, &b, &c);
}

static NormOp norm_op(const glass_norm_operand* n) {
    NormOp o{};
    o.pscale = 1.f;
    if (n && n->stats) {
        o.stats = n->stats;
        o.act = n->act;
        if (n->drop_p > 0.f) {
            o.bits = n->bits;
            o.pscale = 1.f / (1.f - n->drop_p);
        }
    }
    return o;
}

// Operands normalised on load additionally need: widths that are multiples of 32 (keep-bit words never straddle
// the a1|a2 boundary or a K-block), the tcgen05 dW kernel (the SIMT kernels read plain operands only), room for
// the constants next to the resident weights.
bool pair_tc_norm_supported(int k1, int k2, int h) {
    if (k1 % 32 || k2 % 32) return false;
    if (!(h == 64 || h == 128)) return false;
    const int K = k1 + k2;
    if (!(K == 32 || K == 64 || K == 96 || K == 128)) return false;
    int s;
    size_t b;
    uint32_t c;
    return plan(K, 2 * h, true, 0, 8, &s, &b, &c) && s >= 2;
}

int pair_fwd_tc(const float* a1, int64_t lda1, int k1, const float* a2, int64_t lda2, int k2, const float* w0,
                const float* b0, const float* w1, const float* b1, const uint8_t* mask, float z, int act, float* out,
                int64_t ldo, float* acts, int64_t n, int h, const glass_norm_operand* n1, const glass_norm_operand* n2,
                cudaStream_t st) {
    if (ldo % 4 || !aligned16(out) || (acts && !aligned16(acts)) || !aligned16(w0) || !aligned16(w1) || !aligned16(b0) ||
        !aligned16(b1)) {
        set_error("pair_linear_mix_fwd (tcgen05): outputs / weights / biases must be 16-byte aligned");
        return GLASS_ERR_UNSUPPORTED;
    }
    TcParams P{};
    P.a1 = a1, P.lda1 = lda1, P.k1 = k1, P.a2 = a2, P.lda2 = lda2, P.k2 = k2;
    P.w0 = w0, P.w1 = w1, P.b0 = b0, P.b1 = b1, P.mask = mask, P.z = z, P.act = act, P.h = h, P.n = n;
    P.out = out, P.ldo = ldo, P.acts = acts;
    P.kdim = k1 + k2, P.ndim = 2 * h;
    P.wide = ldo % 8 == 0 && (uintptr_t)out % 32 == 0 && (!acts || (uintptr_t)acts % 32 == 0) && h % 8 == 0;
    if (getenv("GLASS_B200_TC_NARROW_STORES")) P.wide = 0;                        // A/B switch for measurements
    P.n1 = norm_op(n1), P.n2 = norm_op(k2 ? n2 : nullptr);
    if (P.n1.stats || P.n2.stats) {
        if (!pair_tc_norm_supported(k1, k2, h)) {
            set_error("pair_linear_mix_fwd (tcgen05): normalised operands need k1, k2 %% 32 == 0 and h in {64, 128}");
            return GLASS_ERR_UNSUPPORTED;
        }
        return launch<false, true>(P, st);
    }
    return launch<false, false>(P, st);
}

int pair_bwd_dx_tc(const float* dout, int64_t lddo, const float* acts, const float* w0, const float* w1,
                   const uint8_t* mask, float z, int act, float* da1, int64_t ldda1, int k1, float* da2, int64_t ldda2,
                   int k2, int64_t n, int h, int acc1, int acc2, cudaStream_t st) {
    if (lddo % 4 || !aligned16(dout) || (acts && !aligned16(acts)) || (da1 && (ldda1 % 4 || !aligned16(da1))) ||
        (da2 && (ldda2 % 4 || !aligned16(da2)))) {
        set_error("pair_linear_mix_bwd (tcgen05): operands must be 16-byte aligned");
        return GLASS_ERR_UNSUPPORTED;
    }
    TcParams P{};
    P.k1 = k1, P.k2 = k2, P.w0 = w0, P.w1 = w1, P.mask = mask, P.z = z, P.act = act, P.h = h, P.n = n;
    P.dout = dout, P.lddo = lddo, P.acts = const_cast<float*>(acts);
    P.da1 = da1, P.ldda1 = ldda1, P.da2 = da2, P.ldda2 = ldda2;
    P.kdim = 2 * h, P.ndim = k1 + k2;
    P.acc1 = acc1, P.acc2 = acc2;
    P.wide = k1 % 8 == 0 && k2 % 8 == 0 && (!da1 || (ldda1 % 8 == 0 && (uintptr_t)da1 % 32 == 0)) &&
             (!da2 || (ldda2 % 8 == 0 && (uintptr_t)da2 % 32 == 0));
    static const bool narrow = getenv("GLASS_B200_TC_NARROW_STORES") != nullptr;   // A/B switch for measurements
    if (narrow) P.wide = 0;
    return launch<true, false>(P, st);
}


bool pair_dw_tc_supported(int k1, int k2, int h, int64_t lda1, int64_t lda2, const void* a1, const void* a2) {
    const int K = k1 + k2;
    if (!(h == 64 || h == 128)) return false;                    // 2h = whole 128-row MMA tiles
    if (!(K == 32 || K == 64 || K == 96 || K == 128)) return false;   // whole 32-element atoms, N <= 128
    if (k1 % 4 || k2 % 4 || lda1 % 4 || !aligned16(a1)) return false;
    if (k2 && (lda2 % 4 || !aligned16(a2))) return false;
    return true;
}

static int64_t dw_tc_rows_per_cta(int64_t n) {
    int64_t r = ceil_div(ceil_div(n, 148), kDwRows) * kDwRows;
    return r < kDwRows ? kDwRows : r;
}

size_t pair_dw_tc_workspace_bytes(int64_t n, int h, int k) {
    const int64_t splits = ceil_div(n > 0 ? n : 1, dw_tc_rows_per_cta(n));
    return (size_t)splits * 2 * (size_t)h * ((size_t)k + 8) * sizeof(float);
}

int pair_bwd_dw_tc(const float* dout, int64_t lddo, const float* acts, const float* a1, int64_t lda1, int k1,
                   const float* a2, int64_t lda2, int k2, const uint8_t* mask, float z, int act, float* dw0, float* db0,
                   float* dw1, float* db1, int64_t n, int h, void* workspace, const glass_norm_operand* n1,
                   const glass_norm_operand* n2, cudaStream_t st) {
    if (lddo % 4 || !aligned16(dout) || (acts && !aligned16(acts)) || ((uintptr_t)workspace & 31)) {
        set_error("pair_linear_mix_bwd dW (tcgen05): operands must be 16-byte (workspace: 32-byte) aligned");
        return GLASS_ERR_UNSUPPORTED;
    }
    const int K = k1 + k2;
    DwParams P{};
    P.dout = dout, P.lddo = lddo, P.acts = acts, P.mask = mask, P.z = z, P.act = act, P.h = h;
    P.a1 = a1, P.lda1 = lda1, P.k1 = k1, P.a2 = a2, P.lda2 = lda2, P.k2 = k2, P.n = n;
    P.part = static_cast<float*>(workspace);
    P.part_ld = K + 8;            // column K holds the db partial; rows stay 32-byte aligned for 256-bit stores
    P.rows_per_cta = dw_tc_rows_per_cta(n);
    const int splits = (int)ceil_div(n, P.rows_per_cta);
    const size_t stage_bytes = 2 * (size_t)128 * 128 + 2 * (size_t)K * 128;
    P.n1 = norm_op(n1), P.n2 = norm_op(k2 ? n2 : nullptr);
    const bool norm = P.n1.stats || P.n2.stats;
    if (norm && (k1 % 32 || k2 % 32)) {
        set_error("pair_linear_mix_bwd dW (tcgen05): normalised operands need k1, k2 %% 32 == 0");
        return GLASS_ERR_UNSUPPORTED;
    }
    const size_t fixed = 1024 + 256 + 16 * 128 * sizeof(float) + (norm ? 16 + (size_t)3 * K * sizeof(float) : 0);
    int stages = (int)(((size_t)kMaxSmem - fixed) / stage_bytes);
    if (stages > 4) stages = 4;
    if (stages < 2) {
        set_error("pair_linear_mix_bwd dW (tcgen05): shape does not fit");
        return GLASS_ERR_UNSUPPORTED;
    }
    P.stages = stages;
    uint32_t cols = 32;
    while (cols < (uint32_t)K) cols <<= 1;
    P.tmem_cols = cols;
    dim3 grid((unsigned)splits, (unsigned)(2 * h / 128));
    const size_t smem = fixed + stages * stage_bytes;
    const int kact = acts ? act : GLASS_ACT_NONE;
#define GLASS_DW_GO(NORM_, ACT_)                                                                                   \
    do {                                                                                                           \
        static bool attr_done = false;                                                                             \
        if (!attr_done) {                                                                                          \
            GLASS_CUDA(cudaFuncSetAttribute(k_pair_dw_tc<NORM_, ACT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem)); \
            attr_done = true;                                                                                      \
        }                                                                                                          \
        k_pair_dw_tc<NORM_, ACT_><<<grid, kDwThreads, smem, st>>>(P);                                              \
    } while (0)
    if (norm) {
        if (kact == GLASS_ACT_ELU) GLASS_DW_GO(true, GLASS_ACT_ELU);
        else if (kact == GLASS_ACT_RELU) GLASS_DW_GO(true, GLASS_ACT_RELU);
        else GLASS_DW_GO(true, GLASS_ACT_NONE);
    } else {
        if (kact == GLASS_ACT_ELU) GLASS_DW_GO(false, GLASS_ACT_ELU);
        else if (kact == GLASS_ACT_RELU) GLASS_DW_GO(false, GLASS_ACT_RELU);
        else GLASS_DW_GO(false, GLASS_ACT_NONE);
    }
#undef GLASS_DW_GO
    GLASS_LAUNCH_CHECK();
    const int64_t total = 2 * (int64_t)h * (K + 1);
    k_pair_dw_tc_reduce<<<(unsigned)ceil_div(total, 32), 256, 0, st>>>(P.part, splits, h, K, P.part_ld, dw0, db0, dw1, db1);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

}  // namespace glass
