// Probe: how fast can the copy engines zero HBM (cudaMemsetAsync, device-to-device copies from a small zero buffer),
// and what do they cost a register-resident DFMA kernel that runs at the same time?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ce_fill_probe ce_fill_probe.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); std::exit(1); } } while (0)

__global__ void fill_kernel(double2* o, long long n2) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) o[i] = make_double2(0.0, 0.0);
}

// variants of the fill kernel: 256-bit stores (sm_100 st.global.v4.f64), block-contiguous tiles
__device__ __forceinline__ void st256(double* p) {
    asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(p), "d"(0.0) : "memory");
}
__global__ void fill256_kernel(double* o, long long n4) {  // grid-stride, 32 B per thread
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) st256(o + 4 * i);
}
template <int TILE_KB, bool WIDE>
__global__ void filltile_kernel(double* o, long long nbytes) {  // a CTA writes whole tiles of TILE_KB
    const long long tile = (long long)TILE_KB << 10;
    const long long ntile = nbytes / tile;
    constexpr int W = WIDE ? 32 : 16;
    for (long long t = blockIdx.x; t < ntile; t += gridDim.x) {
        char* base = reinterpret_cast<char*>(o) + t * tile;
#pragma unroll 4
        for (int off = threadIdx.x * W; off < (TILE_KB << 10); off += blockDim.x * W) {
            if (WIDE) st256(reinterpret_cast<double*>(base + off));
            else *reinterpret_cast<double2*>(base + off) = make_double2(0.0, 0.0);
        }
    }
}

// cache-operator variants of the 128-bit grid-stride fill: 0 .cs (streaming), 1 .wt, 2 .cg, 3 L2::evict_first policy,
// 4 L2::evict_last... no: 4 = .cs 256-bit
template <int OP>
__global__ void fillop_kernel(double* o, long long n2) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned long long pol = 0;
    if (OP == 3) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double* p = o + 2 * i;
        if (OP == 0) asm volatile("st.global.cs.v2.f64 [%0], {%1, %1};" ::"l"(p), "d"(0.0) : "memory");
        if (OP == 1) asm volatile("st.global.wt.v2.f64 [%0], {%1, %1};" ::"l"(p), "d"(0.0) : "memory");
        if (OP == 2) asm volatile("st.global.cg.v2.f64 [%0], {%1, %1};" ::"l"(p), "d"(0.0) : "memory");
        if (OP == 3) asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %1}, %2;" ::"l"(p), "d"(0.0), "l"(pol) : "memory");
    }
}
// TMA bulk stores of a zeroed shared-memory tile, one issuing thread per CTA
__global__ void filltma_kernel(char* o, long long nbytes, int tile_bytes, int depth) {
    extern __shared__ __align__(128) char sm[];
    for (int i = threadIdx.x * 16; i < tile_bytes; i += blockDim.x * 16) *reinterpret_cast<int4*>(sm + i) = make_int4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(sm);
        const long long ntile = nbytes / tile_bytes;
        for (long long t = blockIdx.x; t < ntile; t += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(o + t * tile_bytes), "r"(sa), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else if (depth == 4) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 16;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// 8 independent DFMA chains per thread, iters x 8 x 2 flops per thread
__global__ void __launch_bounds__(128) dfma_kernel(double* sink, int iters, double a) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, 1e-9); x1 = fma(x1, a, 1e-9); x2 = fma(x2, a, 1e-9); x3 = fma(x3, a, 1e-9);
        x4 = fma(x4, a, 1e-9); x5 = fma(x5, a, 1e-9); x6 = fma(x6, a, 1e-9); x7 = fma(x7, a, 1e-9);
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678) sink[0] = s;
}

// scattered 8-byte stores (one per thread per step, pseudo-random sectors): the class kernels' store pattern
__global__ void scatter_kernel(double* o, long long n, int per_thread) {
    unsigned long long h = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    for (int i = 0; i < per_thread; ++i) {
        h = h * 6364136223846793005ull + 1442695040888963407ull;
        o[(long long)((h >> 16) % (unsigned long long)n)] = 1.0;
    }
}

struct Fill {
    int kind;  // 0 kernel, 1 memset, 2 copies
    int nstreams;
    size_t zero_bytes;
};

int main(int argc, char** argv) {
    const double gb = argc > 1 ? std::atof(argv[1]) : 16.0;
    const size_t bytes = (size_t)(gb * 1e9) & ~(size_t)0xfffff;
    char* buf = nullptr;
    CK(cudaMalloc(&buf, bytes));
    const size_t zmax = (size_t)1 << 30;
    char* zero = nullptr;
    CK(cudaMalloc(&zero, zmax));
    CK(cudaMemset(zero, 0, zmax));
    double* sink = nullptr;
    CK(cudaMalloc(&sink, 8));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int nce = 0;
    CK(cudaDeviceGetAttribute(&nce, cudaDevAttrAsyncEngineCount, 0));
    std::printf("buffer %.2f GB, %d SMs, asyncEngineCount %d\n", bytes * 1e-9, sms, nce);
    cudaStream_t sf[8], sc;
    for (auto& s : sf) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking));
    cudaEvent_t e0, ef[8], ec, efork;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&ec)); CK(cudaEventCreate(&efork));
    for (auto& e : ef) CK(cudaEventCreate(&e));

    auto issue_fill = [&](const Fill& f) {
        if (f.kind == 0) { fill_kernel<<<sms * f.nstreams, 256, 0, sf[0]>>>((double2*)buf, (long long)(bytes / 16)); return; }
        if (f.kind == 3) { fill256_kernel<<<sms * f.nstreams, 256, 0, sf[0]>>>((double*)buf, (long long)(bytes / 32)); return; }
        if (f.kind == 4) { filltile_kernel<16, false><<<sms * f.nstreams, 256, 0, sf[0]>>>((double*)buf, (long long)bytes); return; }
        if (f.kind == 5) { filltile_kernel<16, true><<<sms * f.nstreams, 256, 0, sf[0]>>>((double*)buf, (long long)bytes); return; }
        if (f.kind == 6) { filltile_kernel<64, true><<<sms * f.nstreams, 512, 0, sf[0]>>>((double*)buf, (long long)bytes); return; }
        if (f.kind == 8) { fillop_kernel<0><<<sms * f.nstreams, 256, 0, sf[0]>>>((double*)buf, (long long)(bytes / 16)); return; }
        if (f.kind == 9) { fillop_kernel<1><<<sms * f.nstreams, 256, 0, sf[0]>>>((double*)buf, (long long)(bytes / 16)); return; }
        if (f.kind == 10) { fillop_kernel<2><<<sms * f.nstreams, 256, 0, sf[0]>>>((double*)buf, (long long)(bytes / 16)); return; }
        if (f.kind == 11) { fillop_kernel<3><<<sms * f.nstreams, 256, 0, sf[0]>>>((double*)buf, (long long)(bytes / 16)); return; }
        if (f.kind >= 12 && f.kind <= 14) {
            const int tile = (int)f.zero_bytes;
            CK(cudaFuncSetAttribute(filltma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 << 10));
            filltma_kernel<<<sms * f.nstreams, 128, tile, sf[0]>>>(buf, (long long)bytes, tile, f.kind == 12 ? 1 : f.kind == 13 ? 4 : 16);
            return;
        }
        if (f.kind == 7) { filltile_kernel<128, true><<<sms * f.nstreams, 1024, 0, sf[0]>>>((double*)buf, (long long)bytes); return; }
        const size_t part = ((bytes + f.nstreams - 1) / f.nstreams + 0xfffff) & ~(size_t)0xfffff;
        for (int i = 0; i < f.nstreams; ++i) {
            const size_t lo = std::min(bytes, part * i), hi = std::min(bytes, part * (i + 1));
            if (f.kind == 1) { if (hi > lo) CK(cudaMemsetAsync(buf + lo, 0, hi - lo, sf[i])); }
            else for (size_t o = lo; o < hi; o += f.zero_bytes)
                CK(cudaMemcpyAsync(buf + o, zero, std::min(f.zero_bytes, hi - o), cudaMemcpyDeviceToDevice, sf[i]));
        }
    };
    // compute kernels: ~8 ms of DFMA on the whole chip; scatter of ~0.5 G stores
    const int dfma_iters = 1 << 20;
    auto issue_compute = [&](int which) {
        if (which == 1) dfma_kernel<<<sms * 4, 128, 0, sc>>>(sink, dfma_iters, 0.999999);
        if (which == 2) scatter_kernel<<<sms * 8, 256, 0, sc>>>((double*)buf, (long long)(bytes / 8), 1024);
    };
    auto run = [&](const Fill* f, int compute, const char* name) {
        float best_f = 1e9f, best_c = 1e9f, best_all = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, 0));
            for (auto& s : sf) CK(cudaStreamWaitEvent(s, e0, 0));
            CK(cudaStreamWaitEvent(sc, e0, 0));
            if (compute) issue_compute(compute);
            CK(cudaEventRecord(ec, sc));
            const int ns = f ? ((f->kind == 0 || f->kind >= 3) ? 1 : f->nstreams) : 0;  // kernels run on one stream
            if (f) issue_fill(*f);
            for (int i = 0; i < ns; ++i) CK(cudaEventRecord(ef[i], sf[i]));
            CK(cudaDeviceSynchronize());
            float tc = 0, tf = 0;
            CK(cudaEventElapsedTime(&tc, e0, ec));
            for (int i = 0; i < ns; ++i) { float t; CK(cudaEventElapsedTime(&t, e0, ef[i])); tf = std::max(tf, t); }
            best_f = std::min(best_f, tf); best_c = std::min(best_c, tc); best_all = std::min(best_all, std::max(tf, tc));
        }
        std::printf("%-44s fill %8.3f ms (%7.1f GB/s)  compute %8.3f ms  both %8.3f ms\n", name, f ? best_f : 0.f,
                    f ? bytes * 1e-6 / best_f : 0.0, compute ? best_c : 0.f, best_all);
        std::fflush(stdout);
    };
    std::vector<std::pair<Fill, const char*>> fills = {
        {{0, 2, 0}, "fill kernel (2 CTAs x 256 / SM, 128-bit)"},
        {{1, 1, 0}, "cudaMemsetAsync, 1 stream"},
        {{1, 2, 0}, "cudaMemsetAsync, 2 streams"},
        {{1, 4, 0}, "cudaMemsetAsync, 4 streams"},
        {{2, 1, (size_t)8 << 20}, "D2D copies from 8 MB zeros, 1 stream"},
        {{2, 1, (size_t)32 << 20}, "D2D copies from 32 MB zeros, 1 stream"},
        {{2, 2, (size_t)32 << 20}, "D2D copies from 32 MB zeros, 2 streams"},
        {{2, 4, (size_t)32 << 20}, "D2D copies from 32 MB zeros, 4 streams"},
        {{2, 8, (size_t)32 << 20}, "D2D copies from 32 MB zeros, 8 streams"},
        {{2, 4, (size_t)8 << 20}, "D2D copies from 8 MB zeros, 4 streams"},
        {{2, 1, (size_t)1 << 30}, "D2D copies from 1 GB zeros, 1 stream"},
        {{2, 4, (size_t)1 << 30}, "D2D copies from 1 GB zeros, 4 streams"},
    };
    std::vector<std::pair<Fill, const char*>> kfills = {
        {{0, 1, 0}, "128-bit grid-stride, 1 CTA/SM"}, {{0, 2, 0}, "128-bit grid-stride, 2 CTA/SM"}, {{0, 4, 0}, "128-bit grid-stride, 4 CTA/SM"},
        {{0, 8, 0}, "128-bit grid-stride, 8 CTA/SM"},
        {{3, 1, 0}, "256-bit grid-stride, 1 CTA/SM"}, {{3, 2, 0}, "256-bit grid-stride, 2 CTA/SM"}, {{3, 4, 0}, "256-bit grid-stride, 4 CTA/SM"},
        {{3, 8, 0}, "256-bit grid-stride, 8 CTA/SM"},
        {{4, 2, 0}, "128-bit 16 KB tiles, 2 CTA/SM"}, {{4, 4, 0}, "128-bit 16 KB tiles, 4 CTA/SM"}, {{4, 8, 0}, "128-bit 16 KB tiles, 8 CTA/SM"},
        {{5, 2, 0}, "256-bit 16 KB tiles, 2 CTA/SM"}, {{5, 4, 0}, "256-bit 16 KB tiles, 4 CTA/SM"}, {{5, 8, 0}, "256-bit 16 KB tiles, 8 CTA/SM"},
        {{6, 1, 0}, "256-bit 64 KB tiles, 512 thr, 1 CTA/SM"}, {{6, 2, 0}, "256-bit 64 KB tiles, 512 thr, 2 CTA/SM"}, {{6, 4, 0}, "256-bit 64 KB tiles, 512 thr, 4 CTA/SM"},
        {{7, 1, 0}, "256-bit 128 KB tiles, 1024 thr, 1 CTA/SM"}, {{7, 2, 0}, "256-bit 128 KB tiles, 1024 thr, 2 CTA/SM"},
    };
    kfills.insert(kfills.end(), {
        {{8, 2, 0}, "128-bit .cs, 2 CTA/SM"}, {{9, 2, 0}, "128-bit .wt, 2 CTA/SM"}, {{10, 2, 0}, "128-bit .cg, 2 CTA/SM"},
        {{11, 2, 0}, "128-bit L2::evict_first, 2 CTA/SM"}, {{11, 4, 0}, "128-bit L2::evict_first, 4 CTA/SM"},
        {{12, 1, 16 << 10}, "TMA bulk 16 KB, depth 1, 1 CTA/SM"}, {{13, 1, 16 << 10}, "TMA bulk 16 KB, depth 4, 1 CTA/SM"},
        {{14, 1, 16 << 10}, "TMA bulk 16 KB, depth 16, 1 CTA/SM"}, {{14, 2, 16 << 10}, "TMA bulk 16 KB, depth 16, 2 CTA/SM"},
        {{13, 1, 64 << 10}, "TMA bulk 64 KB, depth 4, 1 CTA/SM"}, {{14, 1, 64 << 10}, "TMA bulk 64 KB, depth 16, 1 CTA/SM"},
        {{14, 2, 64 << 10}, "TMA bulk 64 KB, depth 16, 2 CTA/SM"}, {{14, 4, 32 << 10}, "TMA bulk 32 KB, depth 16, 4 CTA/SM"},
        {{14, 4, 4 << 10}, "TMA bulk 4 KB, depth 16, 4 CTA/SM"}, {{14, 8, 8 << 10}, "TMA bulk 8 KB, depth 16, 8 CTA/SM"},
        {{1, 1, 0}, "cudaMemsetAsync"},
    });
    std::printf("--- fill kernel variants, alone\n");
    for (auto& f : kfills) run(&f.first, 0, f.second);
    if (argc > 2) return 0;
    std::printf("--- alone\n");
    run(nullptr, 1, "DFMA kernel alone");
    run(nullptr, 2, "scatter kernel alone");
    for (auto& f : fills) run(&f.first, 0, f.second);
    std::printf("--- next to the DFMA kernel\n");
    for (auto& f : fills) run(&f.first, 1, f.second);
    std::printf("--- next to the scatter kernel\n");
    for (auto& f : fills) run(&f.first, 2, f.second);
    return 0;
}
