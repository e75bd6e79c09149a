// Shared-memory wavefront cost of the gather's LDS.64 patterns (all 12 warps of one CTA run the same pattern; reports SM cycles per LDS).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(long long* out, double* sink) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 6144; i += blockDim.x) sm[i] = 1.0 + i;
    __syncthreads();
    const int s9 = lane / 3, jc = lane % 3, dy = s9 / 3 - 1, dz = s9 % 3 - 1;
    const double* e[4];
    for (int c = 0; c < 4; ++c) {
        const int py = 1 - (c >> 1), pz = 1 - (c & 1), by = py + dy, bz = pz + dz;
        const bool act = lane < 27 && by >= 0 && by <= 1 && bz >= 0 && bz <= 1;
        const double* real = sm + 600 * c + (8 * (4 * py + pz) + (4 * by + bz)) * 9 + jc;
        if (MODE == 0) e[c] = act ? real : sm + 3000;                                  // zero pad (current kernel)
        if (MODE == 1) e[c] = act ? real : sm + 600 * c + (8 * (4 * py + pz)) * 9;       // inactive lanes duplicate an active address (broadcast)
        if (MODE == 2) e[c] = sm + 600 * c + lane;                                     // 32 consecutive doubles
        if (MODE == 3) e[c] = sm + 600 * c + (lane < 16 ? lane : lane - 16);           // 16 consecutive doubles, each read by two lanes
        if (MODE == 4) e[c] = sm + 600 * c + (lane % 12);                              // 12 distinct doubles
        if (MODE == 5) e[c] = sm + 600 * c + 9 * (lane & 15);                          // stride 9 doubles, 16 distinct
    }
    __syncthreads();
    long long t0 = clock64();
    double acc = 0;
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
        double v[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) {
            const int o = 144 * (q / 6) + 18 * ((q / 3) & 1) + 3 * (q % 3);
            v[q] = (e[0][o + rep] + e[1][o + rep]) + (e[2][o + rep] + e[3][o + rep]);
        }
#pragma unroll
        for (int q = 0; q < 12; ++q) acc += v[q];
    }
    __syncthreads();
    long long t1 = clock64();
    sink[threadIdx.x] = acc;
    if (threadIdx.x == 0) out[MODE] = t1 - t0;
}
int main() {
    long long* d; double* s;
    cudaMalloc(&d, 64 * sizeof(long long)); cudaMalloc(&s, 4096 * sizeof(double));
    k<0><<<1, 384, 6144 * 8>>>(d, s); k<1><<<1, 384, 6144 * 8>>>(d, s); k<2><<<1, 384, 6144 * 8>>>(d, s);
    k<3><<<1, 384, 6144 * 8>>>(d, s); k<4><<<1, 384, 6144 * 8>>>(d, s); k<5><<<1, 384, 6144 * 8>>>(d, s);
    cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[6] = {"zero pad", "broadcast dup", "32 consecutive", "16 consecutive x2", "12 distinct", "stride 9 x16"};
    for (int m = 0; m < 6; ++m) printf("mode %d %-18s: %lld cycles for 12 warps x 384 LDS.64 -> %.2f SM cycles per LDS\n", m, names[m], h[m], h[m] / (12.0 * 384));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
