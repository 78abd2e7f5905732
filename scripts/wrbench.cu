// Write-pattern micro-benchmark (scratch tool, not part of the product): how fast can B200 absorb the store
// pattern of the evaluate kernel when no arithmetic is attached to it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/wrbench scripts/wrbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void fill_seq(double* p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 1.5;
}

// CTA per row: n_d candidates, each a [fields][npad] block; warp per candidate, lanes over k (like k2 phase B)
template <int FIELDS>
__global__ void __launch_bounds__(256, 3) row_pattern(double* base, int n_d, int npad, int N, double v) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* row = base + (size_t)blockIdx.x * n_d * FIELDS * npad;
    for (int id = warp; id < n_d; id += 8) {
        double* out = row + (size_t)id * FIELDS * npad + lane;
        for (int kb = 0; kb < N; kb += 32) {
            if (kb + lane < N) {
#pragma unroll
                for (int f = 0; f < FIELDS; ++f) out[f * npad + kb] = v + f;
            }
        }
    }
}

// same bytes, but each warp writes its candidate block as one contiguous run (field-interleaving ignored)
template <int FIELDS>
__global__ void __launch_bounds__(256, 3) row_contig(double* base, int n_d, int npad, int N, double v) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* row = base + (size_t)blockIdx.x * n_d * FIELDS * npad;
    for (int id = warp; id < n_d; id += 8) {
        double* out = row + (size_t)id * FIELDS * npad;
        for (int i = lane; i < FIELDS * N; i += 32) out[i] = v;
    }
}

// lane writes 2 consecutive doubles (16-byte stores)
template <int FIELDS>
__global__ void __launch_bounds__(256, 3) row_pattern_v2(double* base, int n_d, int npad, int N, double v) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* row = base + (size_t)blockIdx.x * n_d * FIELDS * npad;
    for (int id = warp; id < n_d; id += 8) {
        double* out = row + (size_t)id * FIELDS * npad + 2 * lane;
        for (int kb = 0; kb < N; kb += 64) {
            if (kb + 2 * lane + 1 < N) {
#pragma unroll
                for (int f = 0; f < FIELDS; ++f) *reinterpret_cast<double2*>(out + f * npad + kb) = make_double2(v + f, v);
            }
        }
    }
}

template <typename F>
float timeit(F f, int reps = 10) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int i = 0; i < reps; ++i) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const int n_d = 80, npad = 100, N = 100;
    const int rows = 512 * 140;
    const size_t n = (size_t)rows * n_d * 6 * npad;
    double* p;
    if (cudaMalloc(&p, n * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    const double gb = n * 8 / 1e9;
    float t;
    t = timeit([&] { fill_seq<<<148 * 8, 256>>>(p, n); });
    printf("fill_seq            %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
    t = timeit([&] { row_pattern<6><<<rows, 256>>>(p, n_d, npad, N, 1.0); });
    printf("row_pattern<6>      %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
    t = timeit([&] { row_pattern<4><<<rows, 256>>>(p, n_d, npad, N, 1.0); });
    printf("row_pattern<4>      %.3f ms  %.0f GB/s\n", t, gb * 4 / 6 / t * 1e3);
    t = timeit([&] { row_contig<6><<<rows, 256>>>(p, n_d, npad, N, 1.0); });
    printf("row_contig<6>       %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
    t = timeit([&] { row_pattern_v2<6><<<rows, 256>>>(p, n_d, npad, N, 1.0); });
    printf("row_pattern_v2<6>   %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
    // ragged N like the real lattice (51..147), padded to 148
    t = timeit([&] { row_pattern<6><<<rows, 256>>>(p, n_d, 100, 99, 1.0); });
    printf("row_pattern<6> N=99 %.3f ms  %.0f GB/s\n", t, gb * 0.99 / t * 1e3);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
