// Register-resident pipe microbenchmarks for the roofline denominators that MEASURED_PEAKS.json does not hold (SURVEY.md
// §8(d)): FP64 FMA, FP32 FMA and SFU (MUFU.RSQ, MUFU.RSQ64H) issue rates of this B200, plus a streaming-store bandwidth
// (the fused kernel writes, it hardly reads).  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hrm_b200/lib/peaks tools/peaks.cu && hrm_b200/lib/peaks
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

constexpr int kIters = 4096;
constexpr int kChains = 8;

template <class T>
__global__ void k_fma(T* out, T a, T b) {
    T acc[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc[c] = static_cast<T>(threadIdx.x + c);
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) acc[c] = acc[c] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += acc[c];
    if (s == static_cast<T>(-1.2345)) out[0] = s;  // never true: keeps the chains alive
}

__global__ void k_rsq32(float* out, float a) {
    float acc[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc[c] = 1.0f + threadIdx.x + c;
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(acc[c]));
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += acc[c];
    if (s == -1.2345f) out[0] = s + a;
}

__global__ void k_rsq64(double* out, double a) {
    double acc[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc[c] = 1.0 + threadIdx.x + c;
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(acc[c]));  // MUFU.RSQ64H
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += acc[c];
    if (s == -1.2345) out[0] = s + a;
}

__global__ void k_store(double2* out, size_t n, double v) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = make_double2(v, v);
}

template <class F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) return 1;
    const int sms = prop.multiProcessorCount;
    const int grid = sms * 8, block = 256;
    void* buf;
    const size_t bytes = size_t(4) << 30;  // 4 GiB >> L2
    if (cudaMalloc(&buf, bytes) != cudaSuccess) return 2;
    const double threads = double(grid) * block;
    const double ops = threads * kIters * kChains;
    const double ms64 = time_ms([&] { k_fma<double><<<grid, block>>>(static_cast<double*>(buf), 1.0000001, 1e-9); }, 5);
    const double ms32 = time_ms([&] { k_fma<float><<<grid, block>>>(static_cast<float*>(buf), 1.0000001f, 1e-9f); }, 5);
    const double msr32 = time_ms([&] { k_rsq32<<<grid, block>>>(static_cast<float*>(buf), 1.0f); }, 5);
    const double msr64 = time_ms([&] { k_rsq64<<<grid, block>>>(static_cast<double*>(buf), 1.0); }, 5);
    const double mss = time_ms([&] { k_store<<<sms * 16, 512>>>(static_cast<double2*>(buf), bytes / 16, 1.5); }, 5);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    std::printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_mhz_max\": %d, "
                "\"fp64_fma_tflops\": %.2f, \"fp64_fma_per_clk_per_sm\": %.1f, "
                "\"fp32_fma_tflops\": %.2f, \"fp32_fma_per_clk_per_sm\": %.1f, "
                "\"mufu_rsq32_gops\": %.1f, \"mufu_rsq32_per_clk_per_sm\": %.1f, "
                "\"mufu_rsq64h_gops\": %.1f, \"mufu_rsq64h_per_clk_per_sm\": %.1f, "
                "\"store_only_gbs\": %.1f, "
                "\"method\": \"8 independent register chains per thread x %d iterations, %d CTAs x %d threads, CUDA events; per-clock figures use the maximum SM clock\"}\n",
                prop.name, sms, clk / 1000, 2.0 * ops / (ms64 * 1e-3) / 1e12, ops / (ms64 * 1e-3) / (double(clk) * 1e3) / sms,
                2.0 * ops / (ms32 * 1e-3) / 1e12, ops / (ms32 * 1e-3) / (double(clk) * 1e3) / sms, ops / (msr32 * 1e-3) / 1e9,
                ops / (msr32 * 1e-3) / (double(clk) * 1e3) / sms, ops / (msr64 * 1e-3) / 1e9, ops / (msr64 * 1e-3) / (double(clk) * 1e3) / sms,
                double(bytes) / (mss * 1e-3) / 1e9, kIters, grid, block);
    cudaFree(buf);
    return 0;
}
