// integrator_f32.cu — the fast instantiation (R = float, reference-order traversal, FMA allowed) and the
// FMA-peak micro-kernels used as roofline denominators.
#include <cstring>

#include "integrator.cuh"

namespace crb {
template int render_impl<float>(const SceneDeviceData&, Workspace&, const CrCamera&, const CrRenderOpts&, void*, void*, int,
                                cudaStream_t, CrStats*, std::string&);
template int trace_batch_impl<float>(const SceneDeviceData&, const double*, size_t, double, double, CrHit*, uint32_t*, uint32_t*, int,
                                      uint32_t*, cudaStream_t, std::string&);

// Register-resident FMA chains: 8 independent accumulators per thread, 2 flops per FMA.
template <typename T>
__global__ void __launch_bounds__(256) k_fma_peak(T* out, int iters, T a, T b) {
    T x0 = a + (T)threadIdx.x, x1 = x0 + (T)1, x2 = x0 + (T)2, x3 = x0 + (T)3, x4 = x0 + (T)4, x5 = x0 + (T)5, x6 = x0 + (T)6,
      x7 = x0 + (T)7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

template <typename T>
static int time_fma(int num_sms, double* tflops, std::string& err) {
    const int blocks = num_sms * 8, threads = 256, iters = 4096;
    T* d = nullptr;
    if (cudaMalloc(&d, sizeof(T) * blocks * threads) != cudaSuccess) {
        err = "measure_fma_peak: cudaMalloc failed";
        return CR_ERR_CUDA;
    }
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        k_fma_peak<T><<<blocks, threads>>>(d, iters, (T)0.999, (T)0.001);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        if (rep > 0 && ms > 0) best = flops / (ms * 1e-3) / 1e12 > best ? flops / (ms * 1e-3) / 1e12 : best;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess) {
        err = "measure_fma_peak: kernel failed";
        return CR_ERR_CUDA;
    }
    *tflops = best;
    return CR_OK;
}

int measure_fma_peak(int num_sms, double* fp64_tflops, double* fp32_tflops, std::string& err) {
    int rc = time_fma<double>(num_sms, fp64_tflops, err);
    if (rc != CR_OK) return rc;
    return time_fma<float>(num_sms, fp32_tflops, err);
}
}  // namespace crb
