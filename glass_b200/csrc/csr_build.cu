// csr_build: GPU restatement of buildAdj (reference impl/models.py:83-111) producing the
// coalesced, normalised adjacency as CSR plus the CSR of its transpose.
//
// Pipeline (init path, runs once per graph; may synchronise with the host):
//   1. pack (row, col) into 64-bit keys, validate the index range, test "strictly increasing"
//   2. if not already sorted-unique: stable LSD radix sort of (key, entry id)   [cub, device-wide]
//   3. raw row pointers from the sorted rows; degree = sequential fp32 row sum of the RAW weights,
//      deg < 0.5 -> += 1 (models.py:94)
//   4. normalise each raw entry: mean (1/deg)[r]*w (models.py:96-98), sum w (:102),
//      gcn ((deg^-1/2)[r]*w)*(deg^-1/2)[c] (:105-108) -- IEEE rn division / sqrt, no FMA contraction
//   5. merge duplicate (row, col) runs by summing the normalised values in order (== .coalesce())
//   6. row pointers of the merged matrix
//   7. transpose: stable radix sort of the merged entries by column (they are already row-sorted,
//      so the result is (col, row)-sorted), gather rows / values, row pointers of A^T
// HBM traffic: >= 16*nnz (int64 pairs) + 4*nnz (w) read, 2*(8*nnz + 4*(N+1)) written; the sorts add
// 12 B * nnz * 2 * passes when they run.  The device-wide sort and scan are cub primitives
// (header-only, compiled into this library); every other kernel is hand-written.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace glass {
namespace {

constexpr int kThreads = 256;

struct Flags {
    int unsorted;   // some key[i] <= key[i-1]
    int bad_index;  // some index outside [0, n_node)
    int nnz_out;
    int nonunit;    // some weight != 1.0f (unit weights: degree = entry count, exact in any order)
};

__global__ void k_pack_keys(const int64_t* __restrict__ ei, const float* __restrict__ w, int64_t nnz, int64_t n_node,
                            unsigned long long* __restrict__ key, int32_t* __restrict__ idx, Flags* flags) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    if (w[i] != 1.0f) flags->nonunit = 1;
    int64_t r = ei[i], c = ei[nnz + i];
    if (r < 0 || r >= n_node || c < 0 || c >= n_node) {
        flags->bad_index = 1;
        r = 0;
        c = 0;
    }
    unsigned long long k = ((unsigned long long)r << 32) | (unsigned long long)(uint32_t)c;
    key[i] = k;
    idx[i] = (int32_t)i;
    if (i > 0) {
        unsigned long long kp = ((unsigned long long)ei[i - 1] << 32) | (unsigned long long)(uint32_t)ei[nnz + i - 1];
        if (k <= kp) flags->unsorted = 1;
    }
}

__global__ void k_unpack_sorted(const unsigned long long* __restrict__ key, const int32_t* __restrict__ idx,
                                const float* __restrict__ w, int64_t nnz, int32_t* __restrict__ row,
                                int32_t* __restrict__ col, float* __restrict__ ws) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    unsigned long long k = key[i];
    row[i] = (int32_t)(k >> 32);
    col[i] = (int32_t)(k & 0xffffffffu);
    ws[i] = w[idx[i]];
}

// rowptr[r] = first position whose row >= r, for sorted `row`; covers empty rows.
__global__ void k_rowptr(const int32_t* __restrict__ row, int64_t nnz, int64_t n_node, int32_t* __restrict__ rowptr) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i > nnz) return;
    if (nnz == 0) {
        if (i == 0)
            for (int64_t r = 0; r <= n_node; ++r) rowptr[r] = 0;
        return;
    }
    if (i == nnz) {
        for (int64_t r = (int64_t)row[nnz - 1] + 1; r <= n_node; ++r) rowptr[r] = (int32_t)nnz;
        return;
    }
    int64_t cur = row[i];
    int64_t prev = (i == 0) ? -1 : (int64_t)row[i - 1];
    for (int64_t r = prev + 1; r <= cur; ++r) rowptr[r] = (int32_t)i;
}

// Degree = row sum of the raw weights (models.py:90-93).  Unit weights (every dataset the reference produces): the
// sum of k ones is exactly k in fp32 whatever the order (k < 2^24), so the degree is the entry count -- no loop, no
// serial walk over a 660-entry coreness row or a 50 K-entry hub of the stress graph.  Other weights keep the
// sequential fp32 order over the (row, col)-sorted entries that the oracle pins bit for bit.
__global__ void k_degree(const int32_t* __restrict__ rowptr, const float* __restrict__ ws, int64_t n_node,
                         int aggr, float* __restrict__ deg, float* __restrict__ dinv, const Flags* __restrict__ flags,
                         int64_t nnz) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_node) return;
    // (bounds are clamped: during the speculative "input is sorted" pass an unsorted input leaves holes in rowptr)
    const int64_t lo = min(max((int64_t)rowptr[r], (int64_t)0), nnz), hi = min(max((int64_t)rowptr[r + 1], lo), nnz);
    float d = 0.f;
    if (flags && !flags->nonunit) d = (float)(hi - lo);
    else
        for (int64_t e = lo; e < hi; ++e) d = __fadd_rn(d, ws[e]);                    // sequential, sorted order
    if (d < 0.5f) d = __fadd_rn(d, 1.0f);                                          // models.py:94
    deg[r] = d;
    float inv = 1.0f;
    if (aggr == GLASS_AGGR_MEAN) inv = __fdiv_rn(1.0f, d);                          // models.py:96
    else if (aggr == GLASS_AGGR_GCN) inv = __fdiv_rn(1.0f, __fsqrt_rn(d));          // models.py:105 (CPU pow(d,-0.5))
    dinv[r] = inv;
}

__global__ void k_normalise(const int32_t* __restrict__ row, const int32_t* __restrict__ col,
                            const float* __restrict__ ws, const float* __restrict__ dinv, int64_t nnz, int aggr,
                            float* __restrict__ val) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    float w = ws[i];
    float v = w;                                                                     // models.py:102
    if (aggr == GLASS_AGGR_MEAN) v = __fmul_rn(dinv[row[i]], w);                     // models.py:98
    else if (aggr == GLASS_AGGR_GCN) v = __fmul_rn(__fmul_rn(dinv[row[i]], w), dinv[col[i]]);  // models.py:107-108
    val[i] = v;
}

__global__ void k_heads(const int32_t* __restrict__ row, const int32_t* __restrict__ col, int64_t nnz,
                        int32_t* __restrict__ head) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    head[i] = (i == 0 || row[i] != row[i - 1] || col[i] != col[i - 1]) ? 1 : 0;
}

__global__ void k_merge(const int32_t* __restrict__ row, const int32_t* __restrict__ col,
                        const float* __restrict__ val, const int32_t* __restrict__ head,
                        const int32_t* __restrict__ slot, int64_t nnz, int32_t* __restrict__ mrow,
                        int32_t* __restrict__ mcol, float* __restrict__ mval, Flags* flags) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    if (i == nnz - 1) flags->nnz_out = slot[i] + head[i];
    if (!head[i]) return;
    float acc = val[i];
    for (int64_t j = i + 1; j < nnz && !head[j]; ++j) acc = __fadd_rn(acc, val[j]);
    int32_t o = slot[i];
    mrow[o] = row[i];
    mcol[o] = col[i];
    mval[o] = acc;
}

__global__ void k_iota(int32_t* __restrict__ p, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}

__global__ void k_gather_transposed(const int32_t* __restrict__ perm, const int32_t* __restrict__ mrow,
                                    const float* __restrict__ mval, int64_t nnz, int32_t* __restrict__ col_t,
                                    float* __restrict__ val_t) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    int32_t p = perm[i];
    col_t[i] = mrow[p];
    val_t[i] = mval[p];
}


// ---- to_undirected (reference datasets.py:68-71 -> PyG to_undirected: concatenate (row, col) with (col, row), sort by
// (row, col), add the weights of duplicates) on the device ------------------------------------------------------------
__global__ void k_pack_sym(const int64_t* __restrict__ ei, int64_t nnz, int64_t n_node, unsigned long long* __restrict__ key,
                           int32_t* __restrict__ idx, Flags* flags) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    int64_t r = ei[i], c = ei[nnz + i];
    if (r < 0 || r >= n_node || c < 0 || c >= n_node) {
        flags->bad_index = 1;
        r = 0;
        c = 0;
    }
    key[i] = ((unsigned long long)r << 32) | (unsigned long long)(uint32_t)c;
    key[nnz + i] = ((unsigned long long)c << 32) | (unsigned long long)(uint32_t)r;
    idx[i] = (int32_t)i;
    idx[nnz + i] = (int32_t)(nnz + i);  // >= nnz marks the mirrored copy (same weight)
}

__global__ void k_heads64(const unsigned long long* __restrict__ key, int64_t n, int32_t* __restrict__ head) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

// one thread per head: adds the weights of its run in sorted (stable) order; also records whether EVERY run has
// exactly two entries (then the input already was an undirected, duplicate-free edge list)
__global__ void k_merge_sym(const unsigned long long* __restrict__ key, const int32_t* __restrict__ idx,
                            const float* __restrict__ w, const int32_t* __restrict__ head, const int32_t* __restrict__ slot,
                            int64_t n, int64_t* __restrict__ out_row, int64_t* __restrict__ out_col,
                            float* __restrict__ out_w, Flags* flags) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == n - 1) flags->nnz_out = slot[i] + head[i];
    if (!head[i]) return;
    const int32_t half = (int32_t)(n >> 1);
    auto weight = [&](int32_t k) { return w[k >= half ? k - half : k]; };
    float acc = weight(idx[i]);
    int cnt = 1, mirrored = idx[i] >= half ? 1 : 0;
    for (int64_t j = i + 1; j < n && !head[j]; ++j, ++cnt) {
        acc = __fadd_rn(acc, weight(idx[j]));
        mirrored += idx[j] >= half ? 1 : 0;
    }
    // undirected and duplicate-free input <=> every entry occurs once as itself and once as the mirror of its reverse
    if (cnt != 2 || mirrored != 1) flags->unsorted = 1;
    const int32_t o = slot[i];
    out_row[o] = (int64_t)(key[i] >> 32);
    out_col[o] = (int64_t)(key[i] & 0xffffffffu);
    out_w[o] = acc;
}

inline int bits_for(int64_t n) {
    int b = 1;
    while (b < 63 && (1ll << b) < n) ++b;
    return b;
}

struct Plan {
    size_t off_key0, off_key1, off_idx0, off_idx1, off_row, off_col, off_ws, off_val, off_head, off_slot,
        off_mrow, off_dinv, off_flags, off_cub, cub_bytes, total;
};

inline int grid_for(int64_t n) { return (int)ceil_div(n > 0 ? n : 1, kThreads); }

bool make_plan(int64_t nnz, int64_t n_node, Plan* p) {
    size_t cub_a = 0, cub_b = 0, cub_c = 0;
    {
        cub::DoubleBuffer<unsigned long long> k(nullptr, nullptr);
        cub::DoubleBuffer<int32_t> v(nullptr, nullptr);
        if (cub::DeviceRadixSort::SortPairs(nullptr, cub_a, k, v, (int)nnz, 0, 64) != cudaSuccess) return false;
        cub::DoubleBuffer<int32_t> k2(nullptr, nullptr);
        if (cub::DeviceRadixSort::SortPairs(nullptr, cub_b, k2, v, (int)nnz, 0, 32) != cudaSuccess) return false;
        if (cub::DeviceScan::ExclusiveSum(nullptr, cub_c, (int32_t*)nullptr, (int32_t*)nullptr, (int)nnz) != cudaSuccess)
            return false;
    }
    size_t cub_bytes = cub_a > cub_b ? cub_a : cub_b;
    if (cub_c > cub_bytes) cub_bytes = cub_c;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = align_up(o + bytes, 256);
        return at;
    };
    size_t e = (size_t)(nnz > 0 ? nnz : 1);
    p->off_key0 = take(8 * e);
    p->off_key1 = take(8 * e);
    p->off_idx0 = take(4 * e);
    p->off_idx1 = take(4 * e);
    p->off_row = take(4 * e);
    p->off_col = take(4 * e);
    p->off_ws = take(4 * e);
    p->off_val = take(4 * e);
    p->off_head = take(4 * e);
    p->off_slot = take(4 * e);
    p->off_mrow = take(4 * e);
    p->off_dinv = take(4 * (size_t)(n_node > 0 ? n_node : 1));
    p->off_flags = take(sizeof(Flags));
    p->off_cub = take(cub_bytes);
    p->cub_bytes = cub_bytes;
    p->total = o;
    return true;
}

}  // namespace
}  // namespace glass

using namespace glass;

extern "C" size_t glass_csr_build_workspace_bytes(int64_t nnz, int64_t n_node) {
    if (nnz < 0 || n_node < 0 || nnz >= (1ll << 31) || n_node >= (1ll << 31)) {
        set_error("csr_build: nnz=%lld n_node=%lld outside int32 range", (long long)nnz, (long long)n_node);
        return 0;
    }
    Plan p;
    if (!make_plan(nnz, n_node, &p)) {
        set_error("csr_build: cub workspace query failed (no CUDA device?)");
        return 0;
    }
    return p.total;
}

extern "C" int glass_csr_build(const int64_t* edge_index, const float* edge_weight, int64_t nnz, int64_t n_node,
                               int aggr, int32_t* rowptr, int32_t* col, float* val, int32_t* rowptr_t,
                               int32_t* col_t, float* val_t, float* deg, int64_t* nnz_out_host, void* workspace,
                               size_t workspace_bytes, void* stream_) {
    cudaStream_t st = as_stream(stream_);
    if (aggr != GLASS_AGGR_MEAN && aggr != GLASS_AGGR_SUM && aggr != GLASS_AGGR_GCN) {
        set_error("csr_build: unknown aggr %d (reference raises NotImplementedError, models.py:111)", aggr);
        return GLASS_ERR_UNSUPPORTED;
    }
    GLASS_CHECK_ARG(nnz >= 0 && n_node > 0 && nnz < (1ll << 31) && n_node < (1ll << 31),
                    "csr_build: nnz=%lld n_node=%lld outside int32 range", (long long)nnz, (long long)n_node);
    GLASS_CHECK_ARG(rowptr && rowptr_t && deg && nnz_out_host, "csr_build: null output");
    Plan p;
    if (!make_plan(nnz, n_node, &p)) {
        set_error("csr_build: cub workspace query failed");
        return GLASS_ERR_CUDA;
    }
    if (workspace_bytes < p.total || workspace == nullptr) {
        set_error("csr_build: workspace %zu < required %zu", workspace_bytes, p.total);
        return GLASS_ERR_WORKSPACE;
    }
    char* base = static_cast<char*>(workspace);
    auto* key0 = reinterpret_cast<unsigned long long*>(base + p.off_key0);
    auto* key1 = reinterpret_cast<unsigned long long*>(base + p.off_key1);
    auto* idx0 = reinterpret_cast<int32_t*>(base + p.off_idx0);
    auto* idx1 = reinterpret_cast<int32_t*>(base + p.off_idx1);
    auto* row_s = reinterpret_cast<int32_t*>(base + p.off_row);
    auto* col_s = reinterpret_cast<int32_t*>(base + p.off_col);
    auto* w_s = reinterpret_cast<float*>(base + p.off_ws);
    auto* val_s = reinterpret_cast<float*>(base + p.off_val);
    auto* head = reinterpret_cast<int32_t*>(base + p.off_head);
    auto* slot = reinterpret_cast<int32_t*>(base + p.off_slot);
    auto* mrow = reinterpret_cast<int32_t*>(base + p.off_mrow);
    auto* dinv = reinterpret_cast<float*>(base + p.off_dinv);
    auto* flags = reinterpret_cast<Flags*>(base + p.off_flags);
    void* cub_tmp = base + p.off_cub;
    size_t cub_bytes = p.cub_bytes;

    GLASS_CUDA(cudaMemsetAsync(flags, 0, sizeof(Flags), st));
    if (nnz == 0) {
        k_rowptr<<<1, kThreads, 0, st>>>(nullptr, 0, n_node, rowptr);
        k_rowptr<<<1, kThreads, 0, st>>>(nullptr, 0, n_node, rowptr_t);
        k_degree<<<grid_for(n_node), kThreads, 0, st>>>(rowptr, nullptr, n_node, aggr, deg, dinv, nullptr, 0);
        GLASS_LAUNCH_CHECK();
        GLASS_CUDA(cudaStreamSynchronize(st));
        *nnz_out_host = 0;
        return GLASS_OK;
    }
    GLASS_CHECK_ARG(edge_index && edge_weight && col && val && col_t && val_t, "csr_build: null array");

    // 1. keys + validation (flags stay on the device: the common case below needs no host round trip)
    k_pack_keys<<<grid_for(nnz), kThreads, 0, st>>>(edge_index, edge_weight, nnz, n_node, key0, idx0, flags);
    GLASS_LAUNCH_CHECK();

    // Every dataset the reference produces arrives sorted and duplicate-free (datasets.py:68-71), so the pipeline
    // first runs SPECULATIVELY as if that were the case and synchronises once at the end; only when the flags say
    // otherwise (unsorted / duplicate input) is it repeated with the radix sort and the duplicate merge.
    Flags h{};
    int64_t nnz_m = nnz;
    for (int pass = 0; pass < 2; ++pass) {
        const bool sorted_input = pass == 0;
        // 2. sort when needed
        const unsigned long long* key_sorted = key0;
        const int32_t* idx_sorted = idx0;
        if (!sorted_input) {
            cub::DoubleBuffer<unsigned long long> kb(key0, key1);
            cub::DoubleBuffer<int32_t> vb(idx0, idx1);
            GLASS_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, kb, vb, (int)nnz, 0, 32 + bits_for(n_node), st));
            key_sorted = kb.Current();
            idx_sorted = vb.Current();
        }
        k_unpack_sorted<<<grid_for(nnz), kThreads, 0, st>>>(key_sorted, idx_sorted, edge_weight, nnz, row_s, col_s, w_s);
        // 3. raw row pointers (into rowptr, reused if no merge is needed) + degree
        k_rowptr<<<grid_for(nnz + 1), kThreads, 0, st>>>(row_s, nnz, n_node, rowptr);
        k_degree<<<grid_for(n_node), kThreads, 0, st>>>(rowptr, w_s, n_node, aggr, deg, dinv, flags, nnz);
        // 4. normalise
        k_normalise<<<grid_for(nnz), kThreads, 0, st>>>(row_s, col_s, w_s, dinv, nnz, aggr, val_s);
        GLASS_LAUNCH_CHECK();

        // 5./6. merge duplicates (only possible when the input was not strictly increasing)
        nnz_m = nnz;
        const int32_t* mrow_p = row_s;
        if (!sorted_input) {
            k_heads<<<grid_for(nnz), kThreads, 0, st>>>(row_s, col_s, nnz, head);
            GLASS_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, head, slot, (int)nnz, st));
            k_merge<<<grid_for(nnz), kThreads, 0, st>>>(row_s, col_s, val_s, head, slot, nnz, mrow, col, val, flags);
            GLASS_LAUNCH_CHECK();
            GLASS_CUDA(cudaMemcpyAsync(&h, flags, sizeof(Flags), cudaMemcpyDeviceToHost, st));
            GLASS_CUDA(cudaStreamSynchronize(st));
            nnz_m = h.nnz_out;
            mrow_p = mrow;
            k_rowptr<<<grid_for(nnz_m + 1), kThreads, 0, st>>>(mrow, nnz_m, n_node, rowptr);
        } else {
            GLASS_CUDA(cudaMemcpyAsync(col, col_s, 4 * (size_t)nnz, cudaMemcpyDeviceToDevice, st));
            GLASS_CUDA(cudaMemcpyAsync(val, val_s, 4 * (size_t)nnz, cudaMemcpyDeviceToDevice, st));
        }

        // 7. transpose: stable sort by column of the row-sorted merged entries
        {
            int32_t* ck0 = reinterpret_cast<int32_t*>(key1);  // key1 / idx1 are scratch in both passes (key0 / idx0 must
            int32_t* ck1 = ck0 + nnz;                         // survive a speculative first pass for the second one)
            GLASS_CUDA(cudaMemcpyAsync(ck0, col, 4 * (size_t)nnz_m, cudaMemcpyDeviceToDevice, st));
            k_iota<<<grid_for(nnz_m), kThreads, 0, st>>>(idx1, nnz_m);
            cub::DoubleBuffer<int32_t> kb(ck0, ck1);
            cub::DoubleBuffer<int32_t> vb(idx1, reinterpret_cast<int32_t*>(head));
            GLASS_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, kb, vb, (int)nnz_m, 0, bits_for(n_node), st));
            k_gather_transposed<<<grid_for(nnz_m), kThreads, 0, st>>>(vb.Current(), mrow_p, val, nnz_m, col_t, val_t);
            k_rowptr<<<grid_for(nnz_m + 1), kThreads, 0, st>>>(kb.Current(), nnz_m, n_node, rowptr_t);
            GLASS_LAUNCH_CHECK();
        }
        GLASS_CUDA(cudaMemcpyAsync(&h, flags, sizeof(Flags), cudaMemcpyDeviceToHost, st));
        GLASS_CUDA(cudaStreamSynchronize(st));           // the ONE synchronisation of the common (sorted) case
        GLASS_CHECK_ARG(!h.bad_index, "csr_build: edge_index has entries outside [0, %lld)", (long long)n_node);
        if (!sorted_input || !h.unsorted) break;          // speculation held (or this already was the sort pass)
    }
    *nnz_out_host = nnz_m;
    return GLASS_OK;
}

extern "C" size_t glass_to_undirected_workspace_bytes(int64_t nnz) {
    if (nnz < 0 || 2 * nnz >= (1ll << 31)) {
        set_error("to_undirected: nnz=%lld outside the supported range", (long long)nnz);
        return 0;
    }
    Plan p;
    if (!make_plan(2 * nnz, 1, &p)) {
        set_error("to_undirected: cub workspace query failed (no CUDA device?)");
        return 0;
    }
    return p.total;
}

// out_index int64 [2, 2 nnz] (capacity), out_w fp32 [2 nnz]; *nnz_out_host = number of distinct directed entries
// (the rows of out_index are out_index[0 .. m) and out_index[2 nnz .. 2 nnz + m)); *already_host = 1 when the input was
// already undirected and duplicate-free (the reference then leaves it untouched, datasets.py:69).
extern "C" int glass_to_undirected(const int64_t* edge_index, const float* edge_weight, int64_t nnz, int64_t n_node,
                                   int64_t* out_index, float* out_w, int64_t* nnz_out_host, int* already_host,
                                   void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t st = as_stream(stream_);
    GLASS_CHECK_ARG(nnz >= 0 && n_node > 0 && 2 * nnz < (1ll << 31) && n_node < (1ll << 31) && nnz_out_host && already_host,
                    "to_undirected: bad sizes");
    *nnz_out_host = 0;
    *already_host = 0;
    if (nnz == 0) return GLASS_OK;
    GLASS_CHECK_ARG(edge_index && edge_weight && out_index && out_w, "to_undirected: null array");
    const int64_t n2 = 2 * nnz;
    Plan p;
    if (!make_plan(n2, 1, &p) || workspace_bytes < p.total || !workspace) {
        set_error("to_undirected: workspace too small");
        return GLASS_ERR_WORKSPACE;
    }
    char* base = static_cast<char*>(workspace);
    auto* key0 = reinterpret_cast<unsigned long long*>(base + p.off_key0);
    auto* key1 = reinterpret_cast<unsigned long long*>(base + p.off_key1);
    auto* idx0 = reinterpret_cast<int32_t*>(base + p.off_idx0);
    auto* idx1 = reinterpret_cast<int32_t*>(base + p.off_idx1);
    auto* head = reinterpret_cast<int32_t*>(base + p.off_head);
    auto* slot = reinterpret_cast<int32_t*>(base + p.off_slot);
    auto* flags = reinterpret_cast<Flags*>(base + p.off_flags);
    void* cub_tmp = base + p.off_cub;
    size_t cub_bytes = p.cub_bytes;
    GLASS_CUDA(cudaMemsetAsync(flags, 0, sizeof(Flags), st));
    k_pack_sym<<<grid_for(nnz), kThreads, 0, st>>>(edge_index, nnz, n_node, key0, idx0, flags);
    cub::DoubleBuffer<unsigned long long> kb(key0, key1);
    cub::DoubleBuffer<int32_t> vb(idx0, idx1);
    GLASS_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, kb, vb, (int)n2, 0, 32 + bits_for(n_node), st));
    k_heads64<<<grid_for(n2), kThreads, 0, st>>>(kb.Current(), n2, head);
    GLASS_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, head, slot, (int)n2, st));
    k_merge_sym<<<grid_for(n2), kThreads, 0, st>>>(kb.Current(), vb.Current(), edge_weight, head, slot, n2, out_index,
                                                  out_index + n2, out_w, flags);
    GLASS_LAUNCH_CHECK();
    Flags h{};
    GLASS_CUDA(cudaMemcpyAsync(&h, flags, sizeof(Flags), cudaMemcpyDeviceToHost, st));
    GLASS_CUDA(cudaStreamSynchronize(st));
    GLASS_CHECK_ARG(!h.bad_index, "to_undirected: edge_index has entries outside [0, %lld)", (long long)n_node);
    *nnz_out_host = h.nnz_out;
    *already_host = (!h.unsorted && h.nnz_out == nnz) ? 1 : 0;
    return GLASS_OK;
}
