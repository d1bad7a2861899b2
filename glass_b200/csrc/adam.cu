// Multi-tensor Adam step in one launch (fused optimizer for the captured
// train step).  Same update as torch.optim.Adam (GLASSTest.py:213: betas (0.9, 0.999), eps 1e-8, no
// weight decay by default, amsgrad off):
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The step count t and the learning rate live in device memory, so a captured CUDA graph can be replayed
// (t is advanced by the kernel itself: every CTA reads the old value, the last CTA to finish stores t+1).
// Bound: HBM, 7 passes over the parameters (28 B per fp32 parameter).
#include "common.cuh"

namespace glass {
namespace {

struct AdamTensor {   // one row of the device-side table
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
};

constexpr int kAdamThreads = 256;
constexpr int kAdamChunk = 256 * 16;   // elements per CTA

__global__ void __launch_bounds__(kAdamThreads) k_adam(const AdamTensor* __restrict__ tab, const int* __restrict__ chunk_tensor,
                                                       const long long* __restrict__ chunk_begin, const float* __restrict__ lr_dev,
                                                       float* __restrict__ state /* [0] = step count, [1] = ticket */, float b1, float b2,
                                                       float eps, float wd) {
    const float t = state[0] + 1.f;
    const float lr = *lr_dev;
    const float bc1 = 1.f - powf(b1, t), bc2_sqrt = sqrtf(1.f - powf(b2, t));
    const float step_size = lr / bc1;
    const AdamTensor T = tab[chunk_tensor[blockIdx.x]];
    const long long lo = chunk_begin[blockIdx.x];
    const long long hi = lo + kAdamChunk < T.n ? lo + kAdamChunk : T.n;
    const bool vec = ((lo & 3) == 0) && (((uintptr_t)T.p | (uintptr_t)T.g | (uintptr_t)T.m | (uintptr_t)T.v) & 15) == 0;
    auto upd = [&](float& p, float g, float& m, float& v) {
        if (wd != 0.f) g = fmaf(wd, p, g);
        m = fmaf(b1, m, (1.f - b1) * g);
        v = fmaf(b2, v, (1.f - b2) * g * g);
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        p -= step_size * (m / denom);
    };
    if (hi <= lo) {
        // nothing to do for this chunk (parameter without a gradient this step)
    } else if (vec) {
        long long i = lo + 4ll * threadIdx.x;
        for (; i + 3 < hi; i += 4ll * kAdamThreads) {
            float4 p = *reinterpret_cast<float4*>(T.p + i), m = *reinterpret_cast<float4*>(T.m + i), v = *reinterpret_cast<float4*>(T.v + i);
            const float4 g = __ldg(reinterpret_cast<const float4*>(T.g + i));
            upd(p.x, g.x, m.x, v.x);
            upd(p.y, g.y, m.y, v.y);
            upd(p.z, g.z, m.z, v.z);
            upd(p.w, g.w, m.w, v.w);
            *reinterpret_cast<float4*>(T.p + i) = p;
            *reinterpret_cast<float4*>(T.m + i) = m;
            *reinterpret_cast<float4*>(T.v + i) = v;
        }
        if (threadIdx.x == 0) {   // ragged tail (only the last chunk of a tensor whose length is not a multiple of 4)
            for (long long j = lo + ((hi - lo) & ~3ll); j < hi; ++j) {
                float p = T.p[j], m = T.m[j], v = T.v[j];
                upd(p, T.g[j], m, v);
                T.p[j] = p, T.m[j] = m, T.v[j] = v;
            }
        }
    } else {
        for (long long i = lo + threadIdx.x; i < hi; i += kAdamThreads) {
            float p = T.p[i], m = T.m[i], v = T.v[i];
            upd(p, T.g[i], m, v);
            T.p[i] = p, T.m[i] = m, T.v[i] = v;
        }
    }
    // advance the step count once every CTA has read it
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(state + 1), 1u);
        if (ticket == gridDim.x - 1) {
            state[0] = t;
            *reinterpret_cast<unsigned*>(state + 1) = 0u;
        }
    }
}

}  // namespace
}  // namespace glass

using namespace glass;

// table: n_tensors rows of {p, g, m, v, n} (5 x 8 bytes) in DEVICE memory; chunk_tensor / chunk_begin: one entry per
// CTA (host code splits every tensor into chunks of glass_adam_chunk() elements); state: 2 floats in device memory
// {step count, 0}; lr: 1 float in device memory.
extern "C" int glass_adam_chunk(void) { return kAdamChunk; }

extern "C" int glass_adam_step(const void* table, const int32_t* chunk_tensor, const int64_t* chunk_begin, int64_t n_chunks,
                               const float* lr, float* state, float beta1, float beta2, float eps, float weight_decay,
                               void* stream) {
    GLASS_CHECK_ARG(table && chunk_tensor && chunk_begin && lr && state && n_chunks >= 0, "adam_step: bad arguments");
    if (n_chunks == 0) return GLASS_OK;
    static_assert(sizeof(AdamTensor) == 40, "table row layout");
    k_adam<<<(unsigned)n_chunks, kAdamThreads, 0, as_stream(stream)>>>(static_cast<const AdamTensor*>(table), chunk_tensor,
                                                                      reinterpret_cast<const long long*>(chunk_begin), lr, state,
                                                                      beta1, beta2, eps, weight_decay);
    GLASS_LAUNCH_CHECK();
    return GLASS_OK;
}
