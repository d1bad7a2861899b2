// Error reporting, device queries and the small label / index kernels:
// MaxZOZ (impl/utils.py:32-45), the bool mask of impl/models.py:246, pad2batch (impl/utils.py:18-29),
// nn.Embedding lookup (impl/models.py:248) and its gradient.
#include <stdarg.h>

#include "common.cuh"

namespace glass {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached = n;
    return n;
}

namespace {

__global__ void k_scatter_ones(const int64_t* __restrict__ pos, int64_t n_pos, int64_t* __restrict__ z,
                               uint8_t* __restrict__ mask, int64_t n_node) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_pos) return;
    int64_t v = pos[i];
    if (v >= 0 && v < n_node) {  // -1 is padding (impl/utils.py:41-43)
        z[v] = 1;
        if (mask) mask[v] = 1;
    }
}

__global__ void k_label_mask(const int64_t* __restrict__ z, uint8_t* __restrict__ mask, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) mask[i] = z[i] > 0 ? 1 : 0;  // == (z > 0.5) for integer labels, impl/models.py:246
}

// One CTA: stable compaction of the entries >= 0 of a [b, lmax] matrix, row-major.
__global__ void __launch_bounds__(1024) k_pad2batch(const int64_t* __restrict__ pad, int64_t b, int64_t lmax,
                                                    int64_t* __restrict__ batch_out, int64_t* __restrict__ pos_out,
                                                    int64_t* __restrict__ n_valid) {
    __shared__ int64_t s_cnt[1024];
    const int64_t total = b * lmax;
    const int64_t per = (total + blockDim.x - 1) / blockDim.x;
    const int64_t lo = threadIdx.x * per;
    const int64_t hi = lo + per < total ? lo + per : total;
    int64_t cnt = 0;
    for (int64_t i = lo; i < hi; ++i) cnt += pad[i] >= 0;
    s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 counters
    for (int off = 1; off < (int)blockDim.x; off <<= 1) {
        int64_t v = threadIdx.x >= off ? s_cnt[threadIdx.x - off] : 0;
        __syncthreads();
        s_cnt[threadIdx.x] += v;
        __syncthreads();
    }
    int64_t o = s_cnt[threadIdx.x] - cnt;
    for (int64_t i = lo; i < hi; ++i) {
        int64_t v = pad[i];
        if (v >= 0) {
            batch_out[o] = i / lmax;
            pos_out[o] = v;
            ++o;
        }
    }
    if (threadIdx.x == blockDim.x - 1) *n_valid = s_cnt[threadIdx.x];
}

template <int VEC>
__global__ void k_embed_fwd(const float* __restrict__ table, const int64_t* __restrict__ ids,
                            float* __restrict__ out, int64_t ldo, int64_t n, int64_t rows, int h) {
    const int cv = (h + VEC - 1) / VEC;
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; e < n * cv; e += stride) {
        int64_t r = e / cv;
        int c = (int)(e % cv) * VEC;
        int64_t id = ids[r];
        if (id < 0 || id >= rows) id = 0;  // torch would raise; indices are validated on the host side
        if (VEC == 4) {
            *reinterpret_cast<float4*>(out + r * ldo + c) = ldg_f4(table + id * h + c);
        } else {
            out[r * ldo + c] = table[id * h + c];
        }
    }
}

// Gradient of the lookup: each CTA walks RPC consecutive rows, keeps a running sum while the id does not
// change and flushes with one atomicAdd per (run, column).  Constant ids (--use_one) cost one atomic per
// CTA and column; arange ids (--use_nodeid) one uncontended atomic per element.  All RPC (id, value) pairs
// are loaded before the run-length pass, so the kernel is one load round, not RPC dependent ones.
constexpr int kEmbedRowsPerCta = 16;
__global__ void k_embed_bwd(const float* __restrict__ dout, int64_t lddo, const int64_t* __restrict__ ids,
                            float* __restrict__ dtable, int64_t n, int64_t rows, int h) {
    constexpr int RPC = kEmbedRowsPerCta;
    const int64_t r0 = (int64_t)blockIdx.x * RPC;
    for (int c = threadIdx.x; c < h; c += blockDim.x) {
        int64_t id[RPC];
        float v[RPC];
#pragma unroll
        for (int i = 0; i < RPC; ++i) {
            const bool ok = r0 + i < n;
            id[i] = ok ? ids[r0 + i] : -1;
            v[i] = ok ? dout[(r0 + i) * lddo + c] : 0.f;
        }
        float acc = 0.f;
        int64_t cur = -1;
#pragma unroll
        for (int i = 0; i < RPC; ++i) {
            if (id[i] != cur) {
                if (cur >= 0 && cur < rows) atomicAdd(dtable + cur * h + c, acc);
                cur = id[i];
                acc = 0.f;
            }
            acc += v[i];
        }
        if (cur >= 0 && cur < rows) atomicAdd(dtable + cur * h + c, acc);
    }
}

// Deterministic gradient of the lookup (run-to-run bit-reproducible).  The rows are visited in the order of a
// stable sort by id (plan built once per id tensor on the host side: perm, and "runs" = stretches of at most 128
// sorted positions with one id).  Pass 1: one warp per run adds its rows in order (thread per column) into
// run_sum[run, :].  Pass 2: one warp per distinct id adds the runs of that id in order into dtable[id, :].
__global__ void k_embed_bwd_runs(const float* __restrict__ dout, int64_t lddo, const int64_t* __restrict__ perm,
                                 const int32_t* __restrict__ run_begin, const int32_t* __restrict__ run_end,
                                 int64_t n_runs, float* __restrict__ run_sum, int h) {
    const int64_t run = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (run >= n_runs) return;
    const int lane = threadIdx.x & 31;
    const int32_t b = run_begin[run], e = run_end[run];
    for (int c = lane; c < h; c += 32) {
        float acc = 0.f;
        int32_t p = b;
        for (; p + 8 <= e; p += 8) {               // eight independent loads in flight, added in order
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = dout[__ldg(perm + p + u) * lddo + c];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u];
        }
        for (; p < e; ++p) acc += dout[__ldg(perm + p) * lddo + c];
        run_sum[run * h + c] = acc;
    }
}

__global__ void k_embed_bwd_ids(const float* __restrict__ run_sum, const int64_t* __restrict__ uid,
                                const int32_t* __restrict__ uid_first_run, const int32_t* __restrict__ uid_runs,
                                int64_t n_uid, float* __restrict__ dtable, int64_t rows, int h) {
    const int64_t u = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (u >= n_uid) return;
    const int lane = threadIdx.x & 31;
    const int64_t id = uid[u];
    if (id < 0 || id >= rows) return;
    const int32_t r0 = uid_first_run[u], cnt = uid_runs[u];
    for (int c = lane; c < h; c += 32) {
        float acc = 0.f;
        for (int32_t r = 0; r < cnt; ++r) acc += run_sum[(int64_t)(r0 + r) * h + c];
        dtable[id * h + c] = acc;
    }
}

}  // namespace
}  // namespace glass

using namespace glass;

extern "C" int glass_embedding_bwd_ordered(const float* dout, int64_t lddo, const int64_t* perm, const int32_t* run_begin,
                                           const int32_t* run_end, int64_t n_runs, const int64_t* uid,
                                           const int32_t* uid_first_run, const int32_t* uid_runs, int64_t n_uid,
                                           float* run_sum, float* dtable, int64_t rows, int h, void* stream) {
    GLASS_CHECK_ARG(dout && perm && run_begin && run_end && uid && uid_first_run && uid_runs && run_sum && dtable &&
                        n_runs >= 0 && n_uid >= 0 && rows > 0 && h > 0 && lddo >= h,
                    "embedding_bwd_ordered: bad arguments");
    cudaStream_t st = as_stream(stream);
    if (n_runs > 0) k_embed_bwd_runs<<<(unsigned)ceil_div(n_runs, 8), 256, 0, st>>>(dout, lddo, perm, run_begin, run_end, n_runs, run_sum, h);
    if (n_uid > 0) k_embed_bwd_ids<<<(unsigned)ceil_div(n_uid, 8), 256, 0, st>>>(run_sum, uid, uid_first_run, uid_runs, n_uid, dtable, rows, h);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

extern "C" int glass_abi_version(void) { return GLASS_B200_ABI_VERSION; }
extern "C" const char* glass_last_error(void) { return g_err; }
extern "C" int glass_sm_count(void) {
    int n = sm_count();
    if (n <= 0) {
        set_error("no CUDA device");
        return GLASS_ERR_CUDA;
    }
    return n;
}

extern "C" int glass_maxzoz(const int64_t* pos, int64_t n_pos, int64_t* z, uint8_t* mask, int64_t n_node,
                            void* stream) {
    GLASS_CHECK_ARG(z && n_node > 0 && n_pos >= 0 && (pos || n_pos == 0), "maxzoz: bad arguments");
    cudaStream_t st = as_stream(stream);
    GLASS_CUDA(cudaMemsetAsync(z, 0, sizeof(int64_t) * (size_t)n_node, st));
    if (mask) GLASS_CUDA(cudaMemsetAsync(mask, 0, (size_t)n_node, st));
    if (n_pos > 0) {
        k_scatter_ones<<<(unsigned)ceil_div(n_pos, 256), 256, 0, st>>>(pos, n_pos, z, mask, n_node);
        GLASS_LAUNCH_CHECK();
    }
    return GLASS_OK;
}

extern "C" int glass_label_mask(const int64_t* z, uint8_t* mask, int64_t n_node, void* stream) {
    GLASS_CHECK_ARG(z && mask && n_node > 0, "label_mask: bad arguments");
    k_label_mask<<<(unsigned)ceil_div(n_node, 256), 256, 0, as_stream(stream)>>>(z, mask, n_node);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

extern "C" int glass_pad2batch(const int64_t* pad, int64_t b, int64_t lmax, int64_t* batch_out, int64_t* pos_out,
                               int64_t* n_valid, void* stream) {
    GLASS_CHECK_ARG(pad && batch_out && pos_out && n_valid && b >= 0 && lmax >= 0, "pad2batch: bad arguments");
    k_pad2batch<<<1, 1024, 0, as_stream(stream)>>>(pad, b, lmax, batch_out, pos_out, n_valid);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

extern "C" int glass_embedding_fwd(const float* table, const int64_t* ids, float* out, int64_t ldo, int64_t n,
                                   int64_t rows, int h, void* stream) {
    GLASS_CHECK_ARG(table && ids && out && n >= 0 && rows > 0 && h > 0 && ldo >= h, "embedding_fwd: bad arguments");
    if (n == 0) return GLASS_OK;
    cudaStream_t st = as_stream(stream);
    bool vec = (h % 4 == 0) && (ldo % 4 == 0) && ((uintptr_t)table % 16 == 0) && ((uintptr_t)out % 16 == 0);
    int64_t work = n * (vec ? h / 4 : h);
    unsigned grid = (unsigned)std::min<int64_t>(ceil_div(work, 256), (int64_t)sm_count() * 8);
    if (vec) k_embed_fwd<4><<<grid, 256, 0, st>>>(table, ids, out, ldo, n, rows, h);
    else k_embed_fwd<1><<<grid, 256, 0, st>>>(table, ids, out, ldo, n, rows, h);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}

extern "C" int glass_embedding_bwd(const float* dout, int64_t lddo, const int64_t* ids, float* dtable, int64_t n,
                                   int64_t rows, int h, void* stream) {
    GLASS_CHECK_ARG(dout && ids && dtable && n >= 0 && rows > 0 && h > 0 && lddo >= h, "embedding_bwd: bad arguments");
    if (n == 0) return GLASS_OK;
    int threads = h < 32 ? 32 : (h > 256 ? 256 : (h + 31) / 32 * 32);
    k_embed_bwd<<<(unsigned)ceil_div(n, kEmbedRowsPerCta), threads, 0, as_stream(stream)>>>(dout, lddo, ids, dtable, n,
                                                                                          rows, h);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}
