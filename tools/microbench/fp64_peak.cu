// FP64 pipe microbenchmarks on B200: DFMA (vector) and DMMA (mma.sync f64) peak, both alone and
// concurrently, plus shared-memory LDS.64 bandwidth.  The measured DFMA figure is the P_FP64
// denominator SURVEY §8(d) asks for (MEASURED_PEAKS.json has no FP64 entry).
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void dfmaKernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void dmma884Kernel(double* out, int iters) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

template <int ILP, int K>
__global__ void dmma16Kernel(double* out, int iters) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-4 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (K == 8) dmma1688(c[i], a, b); else dmma16816(c[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

// half the warps DFMA, half DMMA: do the pipes overlap?
__global__ void mixedKernel(double* out, int iters, double a, double b) {
    const int warp = threadIdx.x / 32;
    if (warp & 1) {
        double acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-9 + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
        }
        double s = 0;
        for (int i = 0; i < 8; ++i) s += acc[i];
        if (s == 12345.678) out[0] = s;
    } else {
        double c[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
        const double x = threadIdx.x * 1e-3, y = 1.0 + threadIdx.x * 1e-4;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], x, y);
        }
        double s = 0;
        for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
        if (s == 12345.678) out[0] = s;
    }
}

__global__ void ldsKernel(double* out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double s = 0;
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s += sm[(idx + k * 1024) & 8191];
        idx = (idx + 32) & 8191;
    }
    if (s == 12345.678) out[0] = s;
}

template <class F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        const int threads = warps * 32, blocks = sms * (warps >= 32 ? 2 : 2);
        float ms = timeit([&] { dfmaKernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * 8 * iters * (double)threads * blocks;
        printf("DFMA  ILP8  %2d warps/CTA x2 CTA/SM: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    }
    for (int warps : {4, 8, 16}) {
        const int threads = warps * 32, blocks = sms * 2;
        float ms = timeit([&] { dmma884Kernel<8><<<blocks, threads>>>(out, iters); });
        double fl = 2.0 * 256 * 8 * iters * (double)warps * blocks;
        printf("DMMA m8n8k4   ILP8 %2d warps/CTA: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
        ms = timeit([&] { dmma16Kernel<4, 8><<<blocks, threads>>>(out, iters); });
        fl = 2.0 * 16 * 8 * 8 * 4 * iters * (double)warps * blocks;
        printf("DMMA m16n8k8  ILP4 %2d warps/CTA: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
        ms = timeit([&] { dmma16Kernel<4, 16><<<blocks, threads>>>(out, iters); });
        fl = 2.0 * 16 * 8 * 16 * 4 * iters * (double)warps * blocks;
        printf("DMMA m16n8k16 ILP4 %2d warps/CTA: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    }
    {
        const int threads = 512, blocks = sms * 2;
        float ms = timeit([&] { mixedKernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double flv = 2.0 * 8 * iters * 256.0 * blocks, flt = 2.0 * 256 * 8 * iters * 8.0 * blocks;
        printf("MIXED 8 warps DFMA + 8 warps DMMA per CTA: %8.3f ms  vector %7.2f + tensor %7.2f = %7.2f TFLOP/s\n", ms, flv / ms * 1e-9, flt / ms * 1e-9,
               (flv + flt) / ms * 1e-9);
    }
    {
        const int threads = 1024, blocks = sms;
        CK(cudaFuncSetAttribute(ldsKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        float ms = timeit([&] { ldsKernel<<<blocks, threads, 65536>>>(out, 4000); });
        double bytes = 8.0 * 8 * 4000 * (double)threads * blocks;
        printf("LDS.64 conflict-free: %8.3f ms  %8.1f GB/s  (%.1f B/clk/SM at %d MHz nominal)\n", ms, bytes / ms * 1e-6, bytes / ms * 1e-6 / sms / (p.clockRate * 1e-6), p.clockRate / 1000);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
