// micro-benchmark: fp64 vs fp32 FMA issue rate on this GPU (is fp64 arithmetic in epilogues affordable?)
#include <cstdio>
#include <cuda_runtime.h>
template <class T> __global__ void k(T* out, int iters) {
    T a = threadIdx.x * (T)1e-3, b = (T)1.0000001, c = (T)1e-7, d = a + 1, e = a + 2, f = a + 3;
    for (int i = 0; i < iters; ++i) { a = a * b + c; d = d * b + c; e = e * b + c; f = f * b + c; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + e + f;
}
template <class T> float run(const char* name) {
    T* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(T));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<T><<<148 * 8, 256>>>(out, 1000); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<T><<<148 * 8, 256>>>(out, 20000); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = 148.0 * 8 * 256 * 20000 * 4;
    printf("%s: %.3f ms, %.2f TFMA/s = %.1f TFLOP/s\n", name, ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
    return ms;
}
int main() { run<float>("fp32"); run<double>("fp64"); return 0; }
