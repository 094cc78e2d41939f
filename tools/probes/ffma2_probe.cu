// Throughput probe: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100), 8 independent chains per thread, and the same
// with an integer instruction interleaved (is the gain in the FMA pipe or only in issue slots?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe.bin ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b, unsigned int m)
{
    float2 x[8];
    unsigned int u = threadIdx.x * m + 1u;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, blockIdx.x * 0.002f - i);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 2) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }
            else x[i] = __ffma2_rn(x[i], A, B);
            if (MODE >= 2) u = u * m + 12345u;           // one IMAD per two FMAs
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)u;
}

template <int MODE> float run(float* d, int iters)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 10, 1.0001f, 0.5f, 3u);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.5f, 3u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    const double fmas = 148.0 * 8 * 256 * (double)iters * 16;
    const char* names[4] = {"scalar FFMA", "packed FFMA2", "scalar FFMA + IMAD", "packed FFMA2 + IMAD"};
    float ms[4] = {run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters)};
    for (int i = 0; i < 4; ++i) printf("%-22s %.2f ms  %.1f TFMA/s\n", names[i], ms[i], fmas / ms[i] * 1e-9);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
