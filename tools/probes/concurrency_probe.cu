// Probe: can a persistent low-footprint kernel and a shared-memory-heavy kernel share SMs?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o concurrency_probe concurrency_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void spin_a(long long cycles, int* sink) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (sink && threadIdx.x == 1024) *sink = 1;
}
__global__ void spin_b(long long cycles, int* sink) {
    extern __shared__ int sm[];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (sink && sm[threadIdx.x] == -1) *sink = 1;
}

static float run(int a_carve, int b_carve, int a_dyn, int b_dyn, int b_ctas_per_sm, int sms) {
    cudaStream_t s1, s2;
    cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaFuncSetAttribute(spin_b, cudaFuncAttributeMaxDynamicSharedMemorySize, b_dyn);
    cudaFuncSetAttribute(spin_a, cudaFuncAttributePreferredSharedMemoryCarveout, a_carve);
    cudaFuncSetAttribute(spin_b, cudaFuncAttributePreferredSharedMemoryCarveout, b_carve);
    cudaEvent_t e0, e1, ea, eb;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&ea); cudaEventCreate(&eb);
    float best = 1e9, ta = 0, tb = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0, 0);
        cudaStreamWaitEvent(s1, e0, 0);
        cudaStreamWaitEvent(s2, e0, 0);
        spin_a<<<sms * 2, 256, a_dyn, s1>>>(10000000LL, nullptr);  // ~5 ms at 1.9 GHz
        cudaEventRecord(ea, s1);
        spin_b<<<sms * b_ctas_per_sm, 128, b_dyn, s2>>>(4000000LL, nullptr);  // ~2 ms
        cudaEventRecord(eb, s2);
        cudaStreamWaitEvent(0, ea, 0);
        cudaStreamWaitEvent(0, eb, 0);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float t; cudaEventElapsedTime(&t, e0, e1);
        if (t < best) { best = t; cudaEventElapsedTime(&ta, e0, ea); cudaEventElapsedTime(&tb, e0, eb); }
    }
    printf("a_carve %4d b_carve %4d a_dyn %6d b_dyn %6d b_ctas/sm %d : total %.2f ms (A done %.2f, B done %.2f) %s\n", a_carve, b_carve,
           a_dyn, b_dyn, b_ctas_per_sm, best, ta, tb, cudaGetErrorString(cudaGetLastError()));
    cudaStreamDestroy(s1); cudaStreamDestroy(s2);
    return best;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", sms);
    run(-1, -1, 0, 27 * 1024, 4, sms);
    run(-1, -1, 0, 100 * 1024, 1, sms);
    run(100, 100, 0, 27 * 1024, 4, sms);
    run(100, 100, 0, 100 * 1024, 1, sms);
    run(-1, -1, 27 * 1024, 27 * 1024, 4, sms);
    run(-1, -1, 100 * 1024, 100 * 1024, 1, sms);
    run(100, 100, 0, 27 * 1024, 6, sms);
    return 0;
}
