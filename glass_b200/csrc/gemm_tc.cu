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
// CTA = 13 warps, persistent over 128-row tiles (grid = #SMs):
//   warps 0-3  epilogue: tcgen05.ld (TMEM lane = row) -> bias/act/mix -> global
//   warp  4    TMEM allocator; lane 0 issues tcgen05.mma / tcgen05.commit
//   warps 5-12 operand loader: coalesced float4 global loads -> hi/lo split -> st.shared into the
//              128B-swizzled K-major UMMA layout -> fence.proxy.async -> mbarrier arrive
// The weight operand (both weight sets, hi and lo) stays resident in shared memory for the CTA's lifetime;
// the row operand streams through a ring of 32-float K-blocks.  Two accumulators (2 x N TMEM columns)
// overlap the epilogue of tile i with the MMAs of tile i+1.
#include "common.cuh"

namespace glass {
namespace {

constexpr int BM = 128;                 // rows per tile == UMMA M (TMEM lane == row)
constexpr int KBF = 32;                 // floats per K-block == one 128-byte swizzle row
constexpr int kEpiWarps = 4, kLoadWarps = 8;
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kThreads = (kEpiWarps + 1 + kLoadWarps) * 32;
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
};

template <bool BWD>
__global__ void __launch_bounds__(kThreads, 1) k_pair_tc(const TcParams P) {
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

    // ---- one-time setup ------------------------------------------------------------------
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

    const int64_t n_tiles = (P.n + BM - 1) / BM;

    if (warp < kEpiWarps) {
        // ================================ epilogue ==========================================
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(smem_u32(tfull + acc), (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const int64_t row = tile * BM + warp * 32 + lane;
            const bool row_ok = row < P.n;
            const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * N);
            if (!BWD) {
                float c0 = 0.f, c1 = 0.f;
                if (row_ok) {
                    const uint8_t lab = P.mask[row];
                    c1 = lab ? P.z : 1.f - P.z;
                    c0 = lab ? 1.f - P.z : P.z;
                }
                for (int c = 0; c < H; c += 8) {
                    float p0[8], p1[8];
                    tmem_ld8(t_row + (uint32_t)c, p0);
                    tmem_ld8(t_row + (uint32_t)(H + c), p1);
                    tmem_ld_wait();
                    if (row_ok) {
                        float o[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            p0[u] = act_fwd(p0[u] + __ldg(P.b0 + c + u), P.act);
                            p1[u] = act_fwd(p1[u] + __ldg(P.b1 + c + u), P.act);
                            o[u] = __fadd_rn(__fmul_rn(c1, p1[u]), __fmul_rn(c0, p0[u]));
                        }
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
                }
            } else {
                for (int c = 0; c < N; c += 8) {
                    float d[8];
                    tmem_ld8(t_row + (uint32_t)c, d);
                    tmem_ld_wait();
                    if (row_ok) {
                        float* dst = nullptr;
                        if (c < P.k1) {
                            if (P.da1) dst = P.da1 + row * P.ldda1 + c;
                        } else if (P.da2) {
                            dst = P.da2 + row * P.ldda2 + (c - P.k1);
                        }
                        if (dst) {
                            reinterpret_cast<float4*>(dst)[0] = make_float4(d[0], d[1], d[2], d[3]);
                            reinterpret_cast<float4*>(dst)[1] = make_float4(d[4], d[5], d[6], d[7]);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(tempty + acc));
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
        const int lt = threadIdx.x - (kEpiWarps + 1) * 32;                 // 0 .. 255
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < nkb; ++kb) {
                float4 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int q = lt + i * kLoadThreads;
                    const int r = q >> 3, ch = q & 7;
                    const int64_t row = tile * BM + r;
                    const int k = kb * KBF + ch * 4;
                    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row < P.n && k < K) {
                        if (!BWD) {
                            v[i] = (k < P.k1) ? ldg_f4(P.a1 + row * P.lda1 + k) : ldg_f4(P.a2 + row * P.lda2 + (k - P.k1));
                        } else {
                            // dP[row][k..k+3]: k indexes the 2H pre-activation columns (branch 0 | branch 1)
                            const int br = k >= H;
                            const int c = k - br * H;
                            const uint8_t lab = P.mask[row];
                            const float coef = (lab != 0) == (br != 0) ? P.z : 1.f - P.z;
                            float4 g = ldg_f4(P.dout + row * P.lddo + c);
                            g.x *= coef, g.y *= coef, g.z *= coef, g.w *= coef;
                            if (P.acts) {
                                const float4 a = ldg_f4(P.acts + row * (2 * (int64_t)H) + k);
                                g.x *= act_grad_from_out(a.x, P.act);
                                g.y *= act_grad_from_out(a.y, P.act);
                                g.z *= act_grad_from_out(a.z, P.act);
                                g.w *= act_grad_from_out(a.w, P.act);
                            }
                            v[i] = g;
                        }
                    }
                }
                mbar_wait(smem_u32(empty + stage), phase ^ 1);
                uint8_t* dst = a_ring + (size_t)stage * kStageBytes;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int q = lt + i * kLoadThreads;
                    const uint32_t off = swz(q >> 3, q & 7);
                    float4 hi, lo;
                    split_tf32(v[i], hi, lo);
                    *reinterpret_cast<float4*>(dst + off) = hi;
                    *reinterpret_cast<float4*>(dst + BM * 128 + off) = lo;
                }
                fence_proxy_async();
                mbar_arrive(smem_u32(full + stage));
                if (++stage == P.stages) {
                    stage = 0;
                    phase ^= 1;
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

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// shared-memory plan; returns false when the shape does not fit
bool plan(int kdim, int ndim, int* stages, size_t* bytes, uint32_t* tmem_cols) {
    if (ndim < 16 || ndim > 256 || ndim % 16 || kdim < 8 || kdim % 8 || kdim > 512) return false;
    const int nkb = (kdim + KBF - 1) / KBF;
    const size_t b = (size_t)2 * nkb * ndim * 128;
    const size_t fixed = b + 1024 /*alignment slack*/ + 256 /*barriers*/;
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

template <bool BWD>
int launch(TcParams& P, cudaStream_t st) {
    size_t bytes = 0;
    if (!plan(P.kdim, P.ndim, &P.stages, &bytes, &P.tmem_cols)) {
        set_error("pair_linear_mix (tcgen05): shape k=%d n=%d does not fit", P.kdim, P.ndim);
        return GLASS_ERR_UNSUPPORTED;
    }
    static bool attr_done[2] = {false, false};
    if (!attr_done[BWD]) {
        GLASS_CUDA(cudaFuncSetAttribute(k_pair_tc<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        attr_done[BWD] = true;
    }
    const int64_t tiles = ceil_div(P.n, BM);
    int grid = sm_count();
    if (grid > tiles) grid = (int)tiles;
    k_pair_tc<BWD><<<grid, kThreads, bytes, st>>>(P);
    GLASS_LAUNCH_CHECK();
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
    return plan(k1 + k2, 2 * h, &s, &b, &c) && plan(2 * h, k1 + k2, &s, &b, &c);
}

int pair_fwd_tc(const float* a1, int64_t lda1, int k1, const float* a2, int64_t lda2, int k2, const float* w0,
                const float* b0, const float* w1, const float* b1, const uint8_t* mask, float z, int act, float* out,
                int64_t ldo, float* acts, int64_t n, int h, cudaStream_t st) {
    if (ldo % 4 || !aligned16(out) || (acts && !aligned16(acts)) || !aligned16(w0) || !aligned16(w1)) {
        set_error("pair_linear_mix_fwd (tcgen05): outputs / weights must be 16-byte aligned");
        return GLASS_ERR_UNSUPPORTED;
    }
    TcParams P{};
    P.a1 = a1, P.lda1 = lda1, P.k1 = k1, P.a2 = a2, P.lda2 = lda2, P.k2 = k2;
    P.w0 = w0, P.w1 = w1, P.b0 = b0, P.b1 = b1, P.mask = mask, P.z = z, P.act = act, P.h = h, P.n = n;
    P.out = out, P.ldo = ldo, P.acts = acts;
    P.kdim = k1 + k2, P.ndim = 2 * h;
    return launch<false>(P, st);
}

int pair_bwd_dx_tc(const float* dout, int64_t lddo, const float* acts, const float* w0, const float* w1,
                   const uint8_t* mask, float z, int act, float* da1, int64_t ldda1, int k1, float* da2, int64_t ldda2,
                   int k2, int64_t n, int h, cudaStream_t st) {
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
    return launch<true>(P, st);
}

}  // namespace glass
