// FP64 tensor-core (DMMA m8n8k4) peak next to the DFMA peak, one B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_variants/dmma_peak tools/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a0, double b0)
{
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a0, double b0)
{
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
    const int iters = 20000;
    for (int bps : {1, 2, 4, 8}) {
        float ms = time_ms([&] { dmma_kernel<8><<<sms * bps, 256>>>(out, iters, 1.0, 1e-3); });
        double flops = 2.0 * 256 * 8 * (double)iters * 8 /*warps*/ * sms * bps;
        printf("DMMA m8n8k4  8 acc, %d x 8 warps/SM: %7.2f ms  %6.2f TFLOP/s\n", bps, ms, flops / ms * 1e-9);
    }
    {
        float ms = time_ms([&] { dmma_kernel<36><<<sms, 256>>>(out, iters / 4, 1.0, 1e-3); });
        double flops = 2.0 * 256 * 36 * (double)(iters / 4) * 8 * sms;
        printf("DMMA m8n8k4 36 acc, 8 warps/SM: %7.2f ms  %6.2f TFLOP/s\n", ms, flops / ms * 1e-9);
    }
    for (int bps : {2, 8}) {
        float ms = time_ms([&] { dfma_kernel<<<sms * bps, 256>>>(out, iters, 1.0000001, 1e-3); });
        double flops = 2.0 * 16 * (double)iters * 256 * sms * bps;
        printf("DFMA 16 chains, %d x 8 warps/SM: %7.2f ms  %6.2f TFLOP/s\n", bps, ms, flops / ms * 1e-9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
