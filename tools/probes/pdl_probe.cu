// Launch-latency probe: a chain of N dependent tiny kernels on one stream, plain launches vs programmatic dependent
// launch (griddepcontrol.wait at the top of every kernel, launch_dependents right after).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_probe pdl_probe.cu && ./pdl_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void tiny(int* p, int pdl, int early)
{
    if (pdl && early) asm volatile("griddepcontrol.launch_dependents;");
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1;
    if (pdl && !early) asm volatile("griddepcontrol.launch_dependents;");
}

static float run(int n, int pdl, int early, int grid, int* d, cudaStream_t s)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int rep = 0; rep < 20; ++rep) {
        cudaEventRecord(a, s);
        for (int i = 0; i < n; ++i) {
            cudaLaunchConfig_t lc = {}; lc.gridDim = dim3(grid); lc.blockDim = dim3(128); lc.stream = s;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            lc.attrs = at; lc.numAttrs = pdl ? 1 : 0;
            cudaLaunchKernelEx(&lc, tiny, d, pdl, early);
        }
        cudaEventRecord(b, s); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best * 1000.f / n;
}

static float run_graph(int n, int grid, int* d, cudaStream_t s)
{
    cudaGraph_t g; cudaGraphExec_t ex;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
    for (int i = 0; i < n; ++i) tiny<<<grid, 128, 0, s>>>(d, 0, 0);
    cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ex, g, 0);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int rep = 0; rep < 20; ++rep) {
        cudaEventRecord(a, s);
        for (int k = 0; k < 20; ++k) cudaGraphLaunch(ex, s);
        cudaEventRecord(b, s); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best * 1000.f / (20 * n);
}

int main()
{
    int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
    cudaStream_t s; cudaStreamCreate(&s);
    for (int grid : {1, 148, 1184}) {
        printf("grid %4d: plain %.2f us/kernel | pdl (trigger at end) %.2f | pdl (trigger at start) %.2f\n", grid,
               run(200, 0, 0, grid, d, s), run(200, 1, 0, grid, d, s), run(200, 1, 1, grid, d, s));
        printf("           graph of 10 kernels, 20 launches back to back: %.2f us/kernel\n", run_graph(10, grid, d, s));
    }
    int h; cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost); printf("count %d (err %s)\n", h, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
