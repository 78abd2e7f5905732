// Non-tensor FMA peak micro-benchmark (scratch tool, not part of the product): SURVEY.md section 8(d) asks for the collision
// kernel to be read against the measured fp32 / fp64 vector-FMA peak, which MEASURED_PEAKS.json does not hold.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/alu_peak scripts/alu_peak.cu && scripts/alu_peak
// Prints one JSON line: {"fp32_tflops": ..., "fp64_tflops": ..., "sm_count": ..., "how": "..."}
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int ILP>
__global__ void __launch_bounds__(256) fma_chain(T* out, int iters, T a, T b) {
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * a + b;   // contracted to one FMA each
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == (T)-1234.5) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // keeps the chain alive, never true
}

template <typename T>
double measure(int sms, int iters) {
    constexpr int ILP = 8;
    T* out;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * sizeof(T));
    const int grid = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fma_chain<T, ILP><<<grid, 256>>>(out, iters / 10, (T)1.000001, (T)1e-7);
    cudaDeviceSynchronize();
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_chain<T, ILP><<<grid, 256>>>(out, iters, (T)1.000001, (T)1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * ILP * (double)iters * grid * 256;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaFree(out);
    return best;
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) {
        printf("{\"error\": \"no CUDA device\"}\n");
        return 1;
    }
    const double f32 = measure<float>(p.multiProcessorCount, 1 << 16);
    const double f64 = measure<double>(p.multiProcessorCount, 1 << 15);
    printf("{\"fp32_tflops\": %.2f, \"fp64_tflops\": %.2f, \"sm_count\": %d, \"gpu\": \"%s\", "
           "\"how\": \"8 independent FMA chains per thread, 256 threads x 8 CTAs per SM, best of 5, CUDA events (2 flop per FMA)\"}\n",
           f32, f64, p.multiProcessorCount, p.name);
    return 0;
}
