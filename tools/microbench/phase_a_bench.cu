// Isolated throughput / latency of phase A (gaussPointCompact) and phase B (elementBlocks) of the sweep kernel
// as a function of the number of resident warps per SM.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../edelweissfe_b200/csrc/ewb_sweep.cuh"
using namespace ewb;

template <int MC>
__global__ void benchA(const double* stateRef, double* stateTemp, int64_t cstride, MatParams mp, int iters, long long* out, int* fail) {
    using R = RecLayout<MC>;
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* rec = sm + (size_t)warp * (4 * R::PER_EL + 108);
    double* stage = rec + 4 * R::PER_EL;
    for (int i = lane; i < 108; i += 32) {
        const int node = i / 6, c = i % 6;
        const int X = node / 9, Y = (node / 3) % 3, Z = node % 3;
        stage[i] = c < 3 ? (c == 0 ? X : (c == 1 ? Y : Z)) * 1.0 + 0.01 * ((i * 37) % 11) : 1e-3 * ((i * 53) % 17 - 8);
    }
    __syncwarp();
    const int ak = lane >> 3, agp = lane & 7;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const int64_t e = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * iters * 4 + it * 4 + ak;
        const int64_t off = e * 8 + agp;
        gaussPointCompact<MC, false>(rec + ak * R::PER_EL + agp * R::RS, stage + ((ak >> 1) * 3 + (ak & 1)) * 6, agp, mp, stateRef + off, stateTemp + off, cstride, true, fail);
        __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
}

template <int MC>
__global__ void benchB(MatParams mp, int iters, long long* out, double* sink) {
    using R = RecLayout<MC>;
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* rec = sm + (size_t)warp * (4 * R::PER_EL + 108);
    for (int i = lane; i < 4 * R::PER_EL; i += 32) rec[i] = 0.01 * ((i * 29) % 23) + ((i % R::RS) % 4 == 0 ? 1.0 : 0.0);
    __syncwarp();
    double dNl[2][3] = {{0.1, -0.2, 0.05}, {0.07, 0.11, -0.13}};
    double acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double K0[9], K1[9], Pr[3];
        elementBlocks<MC>(rec + (it & 3) * R::PER_EL, lane, dNl, mp, true, K0, K1, Pr);
        acc += K0[0] + K1[4] + Pr[1] + K0[8] + K1[2];
    }
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    if (acc == 1.2345) sink[0] = acc;
}

int main() {
    MatParams mp{};
    mp.kind = 0; mp.lambda = 6762.0; mp.G = 8606.0;
    const int iters = 200, nsm = 148;
    long long* out; int* fail; double* sink;
    cudaMalloc(&out, 64 * 1024 * 8); cudaMalloc(&fail, 4); cudaMalloc(&sink, 8);
    for (int warps : {1, 2, 4, 8, 12, 16, 24, 32}) {
        const int64_t nEl = (int64_t)nsm * warps * iters * 4;
        const int64_t cstride = nEl * 8;
        double *sr, *st;
        cudaMalloc(&sr, cstride * 12 * 8); cudaMalloc(&st, cstride * 12 * 8);
        cudaMemset(sr, 0, cstride * 12 * 8);
        const size_t smem = (size_t)warps * (4 * RecLayout<MC_LE>::PER_EL + 108) * 8;
        cudaFuncSetAttribute(benchA<MC_LE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(benchB<MC_LE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        benchA<MC_LE><<<nsm, warps * 32, smem>>>(sr, st, cstride, mp, iters, out, fail);
        benchA<MC_LE><<<nsm, warps * 32, smem>>>(sr, st, cstride, mp, iters, out, fail);
        cudaDeviceSynchronize();
        std::vector<long long> h(nsm * warps);
        cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
        double mA = 0; for (auto v : h) mA += (double)v; mA /= h.size() * iters;
        benchB<MC_LE><<<nsm, warps * 32, smem>>>(mp, iters * 4, out, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
        double mB = 0; for (auto v : h) mB += (double)v; mB /= h.size() * iters * 4;
        printf("%2d warps/SM: phase A %7.0f cycles per warp-task (4 elements) -> %6.1f SM-cycles/element | phase B %6.0f cycles per element-warp -> %6.1f SM-cycles/element   [%s]\n",
               warps, mA, mA / warps / 4, mB, mB / warps, cudaGetErrorString(cudaGetLastError()));
        cudaFree(sr); cudaFree(st);
    }
    return 0;
}
