// Do the tensor-pipe phase (elementBlocks: 18 DMMA + epilogue) and the shared-memory phase (18 accumulator RMWs per lane)
// of the sweep consumers overlap when different warps run them at the same time?
//   mode 0: every warp alternates tile / accumulate, free running
//   mode 1: same, but a CTA barrier after each phase (forced lock step)
//   mode 2: half of the warps only compute tiles (2x iterations), the other half only accumulate (2x iterations)
//   mode 3: tiles only      mode 4: accumulate only
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../edelweissfe_b200/csrc/ewb_sweep.cuh"
using namespace ewb;

__device__ __forceinline__ void accumulate(double* acc, int lane, const double* K0, const double* K1) {
    // same access shape as the sweep kernel: two 3x3 blocks per lane, rows 27 doubles apart
    double* d0 = acc + (lane >> 2) * 129 + (lane & 3) * 20;  // conflict-free: (row + 4q) mod 16 distinct per half-warp
    double* d1 = d0 + 3;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            d0[i * 27 + j] += K0[i * 3 + j];
            d1[i * 27 + j] += K1[i * 3 + j];
        }
}

template <int MODE>
__global__ void bench(MatParams mp, int iters, long long* out, double* sink) {
    using R = RecLayout<MC_LE>;
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    double* rec = sm + (size_t)warp * (4 * R::PER_EL);
    double* acc = sm + (size_t)nw * (4 * R::PER_EL) + (size_t)warp * 8 * 129;
    for (int i = lane; i < 4 * R::PER_EL; i += 32) rec[i] = 0.01 * ((i * 29) % 23) + ((i % R::RS) % 4 == 0 ? 1.0 : 0.0);
    for (int i = lane; i < 8 * 129; i += 32) acc[i] = 0.0;
    __syncthreads();
    double dNl[2][3] = {{0.1, -0.2, 0.05}, {0.07, 0.11, -0.13}};
    double K0[9], K1[9], Pr[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) { K0[i] = 0.001 * i; K1[i] = 0.002 * i; }
    double s = 0;
    const long long t0 = clock64();
    if (MODE == 0 || MODE == 1) {
        for (int it = 0; it < iters; ++it) {
            elementBlocks<MC_LE>(rec + (it & 3) * R::PER_EL, lane, dNl, mp, true, K0, K1, Pr);
            if (MODE == 1) __syncthreads();
            accumulate(acc, lane, K0, K1);
            s += Pr[1];
            if (MODE == 1) __syncthreads();
        }
    } else if (MODE == 2) {
        if (warp & 1) {
            for (int it = 0; it < 2 * iters; ++it) {
                elementBlocks<MC_LE>(rec + (it & 3) * R::PER_EL, lane, dNl, mp, true, K0, K1, Pr);
                s += K0[0] + K1[4] + Pr[1] + K0[8] + K1[2];
            }
        } else {
            for (int it = 0; it < 2 * iters; ++it) { accumulate(acc, lane, K0, K1); asm volatile("" ::: "memory"); }
        }
    } else if (MODE == 5) {  // software pipelined: accumulate the previous element while the next one's tiles are computed
        double N0[9], N1[9];
        elementBlocks<MC_LE>(rec, lane, dNl, mp, true, K0, K1, Pr);
        for (int it = 1; it <= iters; ++it) {
            elementBlocks<MC_LE>(rec + (it & 3) * R::PER_EL, lane, dNl, mp, true, N0, N1, Pr);
            accumulate(acc, lane, K0, K1);
            s += Pr[1];
#pragma unroll
            for (int i = 0; i < 9; ++i) { K0[i] = N0[i]; K1[i] = N1[i]; }
        }
    } else if (MODE == 3) {
        for (int it = 0; it < iters; ++it) {
            elementBlocks<MC_LE>(rec + (it & 3) * R::PER_EL, lane, dNl, mp, true, K0, K1, Pr);
            s += K0[0] + K1[4] + Pr[1] + K0[8] + K1[2];
        }
    } else {
        for (int it = 0; it < iters; ++it) { accumulate(acc, lane, K0, K1); asm volatile("" ::: "memory"); }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (s == 1.2345) sink[0] = s + acc[lane];
}

template <int MODE>
double run(int warps, int iters, MatParams mp, long long* out, double* sink) {
    const int nsm = 148;
    const size_t smem = (size_t)warps * (4 * RecLayout<MC_LE>::PER_EL + 8 * 129) * 8;
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<MODE><<<nsm, warps * 32, smem>>>(mp, iters, out, sink);
    bench<MODE><<<nsm, warps * 32, smem>>>(mp, iters, out, sink);
    cudaDeviceSynchronize();
    std::vector<long long> h(nsm);
    cudaMemcpy(h.data(), out, nsm * 8, cudaMemcpyDeviceToHost);
    double m = 0; for (auto v : h) m += (double)v;
    return m / nsm / ((double)iters * warps);  // SM-cycles per (tile + accumulate) element
}

int main() {
    MatParams mp{};
    mp.kind = 0; mp.lambda = 6762.0; mp.G = 8606.0;
    long long* out; double* sink;
    cudaMalloc(&out, 4096 * 8); cudaMalloc(&sink, 8);
    const int iters = 400;
    for (int warps : {8, 12, 16}) {
        const double a = run<0>(warps, iters, mp, out, sink), b = run<1>(warps, iters, mp, out, sink), c = run<2>(warps, iters, mp, out, sink),
                     d = run<3>(warps, iters, mp, out, sink), e = run<4>(warps, iters, mp, out, sink), f = run<5>(warps, iters, mp, out, sink);
        printf("%2d warps/SM: SM-cycles per element  free-running %6.1f | lock-step %6.1f | role split %6.1f | tiles only %6.1f | accumulate only %6.1f | sw-pipelined %6.1f  [%s]\n",
               warps, a, b, c, d, e, f, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
