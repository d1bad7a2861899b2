// Label-batch data parallelism without a library collective in the training step.
//
// The reference has no distributed code; here the base graph, its CSR and the parameters are replicated and every
// rank trains on its own label batches, so the only exchange per step is the gradient average.  Its bulk is the
// dense gradient of the N x H embedding table (14.7 MB at the em_user shape); all-reduce -> Adam serialises
// ~100-140 us behind the backward pass (round-1 SCALE: 0.83 efficiency at 8 GPUs).  This kernel fuses
//     reduce-scatter  +  Adam on the owned slice  +  all-gather of the updated parameters
// into ONE launch over NVLink peer memory: every rank keeps [table | table gradient | small gradients | flags] in a
// symmetric (peer-mapped) allocation and
//   1. announces "my gradients are complete" to every peer and waits for theirs         (flag 1, system scope)
//   2. for the table rows it OWNS: g = (sum_r G_r[row]) / P read straight from the peers' buffers in rank order
//      (deterministic; every element is reduced by exactly one rank, so all replicas stay bit-identical), applies
//      Adam to its slice of m / v / p and stores the new parameters into EVERY rank's table
//   3. averages all ranks' small-gradient blocks into a local buffer (same order on every rank -> same values);
//      the ordinary multi-tensor Adam launch then updates the small parameters from it
//   4. announces "my stores are done", the last CTA waits for every peer's announcement  (flag 2): when the
//      kernel ends, the local table is complete and the next step may read it.
// Per rank (P - 1)/P x 14.7 MB cross NVLink in each direction instead of 2 x that for an all-reduce plus 7 full
// passes of Adam over the table.  Flags carry a monotonically increasing step number (no reset races); the spin
// loops give up after ~2 s and raise an error flag instead of hanging the device.
#include <algorithm>

#include "common.cuh"

namespace glass {
namespace {

constexpr int kDpThreads = 512;
constexpr int kDpMaxWorld = 16;

struct DpArgs {
    int world, rank;
    const unsigned long long* peer_base;   // device array [world]: this process' mapping of every rank's block
    unsigned long long off_table, off_grad, off_small, off_flags;   // byte offsets inside a block
    long long table_elems;                 // rows * cols of the table (multiple of 4)
    long long own_begin, own_end;          // element range of the table this rank reduces / updates (multiples of 4)
    long long small_elems;
    float* m;                              // Adam moments of the table (local, full size; only the owned range is used)
    float* v;
    float* small_out;                      // averaged small gradients (local)
    const float* lr;                       // device scalars shared with glass_adam_step
    const float* state;                    // [0] = optimizer step count (advanced by glass_adam_step afterwards)
    float b1, b2, eps, wd;
    unsigned long long* epoch;             // local: number of completed exchanges
    unsigned* ticket;                      // local: CTAs finished
    int* error;                            // local: set to 1 on a spin time-out
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_nc_sys(const float* p) {   // peer memory: do not keep it in L1
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long want, int* error) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < want) {
        if (clock64() - t0 > 4000000000ll) {   // ~2 s at 1.9 GHz: a peer died or was never launched
            *error = 1;
            return false;
        }
        __nanosleep(64);
    }
    return true;
}

__global__ void __launch_bounds__(kDpThreads) k_dp_adam(const DpArgs A) {
    const int P = A.world;
    const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(A.epoch) + 1;
    char* my_block = reinterpret_cast<char*>(A.peer_base[A.rank]);
    unsigned long long* my_flags = reinterpret_cast<unsigned long long*>(my_block + A.off_flags);   // [2][kDpMaxWorld]

    // ---- 1. my gradients are complete (they were written by earlier kernels of this stream) -> tell every peer
    if (blockIdx.x == 0 && threadIdx.x < P) {
        __threadfence_system();
        unsigned long long* f = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(A.peer_base[threadIdx.x]) + A.off_flags);
        st_release_sys(f + A.rank, e);
    }
    if (threadIdx.x < P) spin_until(my_flags + threadIdx.x, e, A.error);
    __syncthreads();

    // ---- 2. owned slice of the table: reduce over ranks, Adam, broadcast
    const float t = A.state[0] + 1.f;
    const float lr = *A.lr;
    const float bc1 = 1.f - powf(A.b1, t), bc2_sqrt = sqrtf(1.f - powf(A.b2, t));
    const float step_size = lr / bc1, inv_p = 1.f / (float)P;
    const long long stride = 4ll * gridDim.x * kDpThreads;
    for (long long i = A.own_begin + 4ll * (blockIdx.x * (long long)kDpThreads + threadIdx.x); i < A.own_end; i += stride) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < P; ++r) {
            const float* src = reinterpret_cast<const float*>(reinterpret_cast<const char*>(A.peer_base[r]) + A.off_grad) + i;
            const float4 x = ld_nc_sys(src);
            g.x += x.x, g.y += x.y, g.z += x.z, g.w += x.w;
        }
        g.x *= inv_p, g.y *= inv_p, g.z *= inv_p, g.w *= inv_p;
        float* pt = reinterpret_cast<float*>(my_block + A.off_table) + i;
        float4 p = *reinterpret_cast<float4*>(pt);
        float4 m = *reinterpret_cast<float4*>(A.m + i), v = *reinterpret_cast<float4*>(A.v + i);
        auto upd = [&](float& pp, float gg, float& mm, float& vv) {
            if (A.wd != 0.f) gg = fmaf(A.wd, pp, gg);
            mm = fmaf(A.b1, mm, (1.f - A.b1) * gg);
            vv = fmaf(A.b2, vv, (1.f - A.b2) * gg * gg);
            pp -= step_size * (mm / (sqrtf(vv) / bc2_sqrt + A.eps));
        };
        upd(p.x, g.x, m.x, v.x);
        upd(p.y, g.y, m.y, v.y);
        upd(p.z, g.z, m.z, v.z);
        upd(p.w, g.w, m.w, v.w);
        *reinterpret_cast<float4*>(A.m + i) = m;
        *reinterpret_cast<float4*>(A.v + i) = v;
        for (int r = 0; r < P; ++r)
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(reinterpret_cast<char*>(A.peer_base[r]) + A.off_table) + i) = p;
    }

    // ---- 3. small gradients: every rank averages all blocks in rank order
    for (long long i = 4ll * (blockIdx.x * (long long)kDpThreads + threadIdx.x); i < A.small_elems; i += stride) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < P; ++r) {
            const float4 x = ld_nc_sys(reinterpret_cast<const float*>(reinterpret_cast<const char*>(A.peer_base[r]) + A.off_small) + i);
            g.x += x.x, g.y += x.y, g.z += x.z, g.w += x.w;
        }
        *reinterpret_cast<float4*>(A.small_out + i) = make_float4(g.x * inv_p, g.y * inv_p, g.z * inv_p, g.w * inv_p);
    }

    // ---- 4. all CTAs done -> announce, the last CTA waits for every peer's announcement
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(A.ticket, 1u) == gridDim.x - 1) {
            for (int r = 0; r < P; ++r) {
                unsigned long long* f = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(A.peer_base[r]) + A.off_flags);
                st_release_sys(f + kDpMaxWorld + A.rank, e);
            }
            for (int r = 0; r < P; ++r)
                if (!spin_until(my_flags + kDpMaxWorld + r, e, A.error)) break;
            *A.epoch = e;
            *A.ticket = 0u;
            __threadfence();
        }
    }
}

}  // namespace
}  // namespace glass

using namespace glass;

extern "C" int glass_dp_flags_bytes(void) { return 2 * kDpMaxWorld * (int)sizeof(unsigned long long); }

extern "C" int glass_dp_adam_step(int world, int rank, const unsigned long long* peer_base, unsigned long long off_table,
                                  unsigned long long off_grad, unsigned long long off_small, unsigned long long off_flags,
                                  int64_t table_elems, int64_t own_begin, int64_t own_end, int64_t small_elems, float* m,
                                  float* v, float* small_out, const float* lr, const float* state, float beta1,
                                  float beta2, float eps, float weight_decay, unsigned long long* epoch,
                                  unsigned* ticket, int* error, void* stream) {
    GLASS_CHECK_ARG(world >= 1 && world <= kDpMaxWorld && rank >= 0 && rank < world && peer_base && lr && state && epoch &&
                        ticket && error,
                    "dp_adam_step: bad arguments");
    GLASS_CHECK_ARG(table_elems >= 0 && table_elems % 4 == 0 && own_begin % 4 == 0 && own_end % 4 == 0 &&
                        own_begin >= 0 && own_begin <= own_end && own_end <= table_elems && small_elems >= 0 &&
                        small_elems % 4 == 0 && (table_elems == 0 || (m && v)) && (small_elems == 0 || small_out),
                    "dp_adam_step: element ranges must be multiples of 4 and inside the table");
    GLASS_CHECK_ARG((off_table | off_grad | off_small | off_flags) % 16 == 0, "dp_adam_step: offsets must be 16-byte aligned");
    DpArgs A{};
    A.world = world, A.rank = rank, A.peer_base = peer_base;
    A.off_table = off_table, A.off_grad = off_grad, A.off_small = off_small, A.off_flags = off_flags;
    A.table_elems = table_elems, A.own_begin = own_begin, A.own_end = own_end, A.small_elems = small_elems;
    A.m = m, A.v = v, A.small_out = small_out, A.lr = lr, A.state = state;
    A.b1 = beta1, A.b2 = beta2, A.eps = eps, A.wd = weight_decay;
    A.epoch = epoch, A.ticket = ticket, A.error = error;
    const int64_t work = std::max<int64_t>((own_end - own_begin) / 4, small_elems / 4);
    int64_t grid = ceil_div(std::max<int64_t>(work, 1), (int64_t)kDpThreads * 2);
    if (grid > 2 * sm_count()) grid = 2 * sm_count();
    k_dp_adam<<<(unsigned)grid, kDpThreads, 0, as_stream(stream)>>>(A);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}
