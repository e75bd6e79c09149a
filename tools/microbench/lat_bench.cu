// Latency / issue-rate microbenchmarks (single warp unless stated): DMMA, DFMA/DADD, LDS.64, STS.128 on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lat_bench lat_bench.cu && ./lat_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__global__ void k(long long* out, double* sink, int nwarps_active) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 6144; i += blockDim.x) sm[i] = 1.0 + i;
    __syncthreads();
    if (warp >= nwarps_active) return;
    double a = 1.0 + lane, b = 0.5 * lane;
    long long t[8];
    // 1: dependent DMMA chain
    {
        double c[2] = {0, 0};
        long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < 64; ++i) dmma(c, a, b);
        t[0] = clock64() - t0;
        sink[threadIdx.x] = c[0] + c[1];
    }
    // 2: 9 independent accumulators x 2 k-steps (the element pattern), repeated 4 times
    {
        double c[9][2];
#pragma unroll
        for (int i = 0; i < 9; ++i) c[i][0] = c[i][1] = 0;
        long long t0 = clock64();
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 9; ++i) dmma(c[i], a + r, b + i);
        t[1] = clock64() - t0;
        double s = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) s += c[i][0] + c[i][1];
        sink[threadIdx.x] += s;
    }
    // 3: dependent DFMA chain
    {
        double x = a;
        long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < 64; ++i) x = fma(x, b, a);
        t[2] = clock64() - t0;
        sink[threadIdx.x] += x;
    }
    // 4: 12 independent DFMA chains
    {
        double x[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] = a + i;
        long long t0 = clock64();
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < 12; ++i) x[i] = fma(x[i], b, a);
        t[3] = clock64() - t0;
        double s = 0;
#pragma unroll
        for (int i = 0; i < 12; ++i) s += x[i];
        sink[threadIdx.x] += s;
    }
    // 5: LDS.64 stream, 48 independent loads, conflict-free (lane stride 1)
    {
        double s = 0;
        long long t0 = clock64();
        double v[48];
#pragma unroll
        for (int i = 0; i < 48; ++i) v[i] = sm[lane + 33 * i];
#pragma unroll
        for (int i = 0; i < 48; ++i) s += v[i];
        t[4] = clock64() - t0;
        sink[threadIdx.x] += s;
    }
    // 6: dependent LDS chain (pointer chase) latency
    {
        int idx = lane;
        long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < 32; ++i) idx = (int)sm[idx & 1023] & 1023;
        t[5] = clock64() - t0;
        sink[threadIdx.x] += idx;
    }
    // 7: STS.128 x 9 (lane stride 144 B) repeated 8 times
    {
        double2* p = reinterpret_cast<double2*>(sm + lane * 18);
        long long t0 = clock64();
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 9; ++i) p[i] = make_double2(a + r, b + i);
        __syncwarp();
        t[6] = clock64() - t0;
    }
    // 8: the gather pattern of the row-pipelined kernel: 4 colour pointers, 12 offsets each, tree adds, 9 outputs stored to smem
    {
        const int s9 = lane / 3, jc = lane % 3, dy = s9 / 3 - 1, dz = s9 % 3 - 1;
        const double* e[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int py = 1 - (c >> 1), pz = 1 - (c & 1), by = py + dy, bz = pz + dz;
            const bool act = lane < 27 && by >= 0 && by <= 1 && bz >= 0 && bz <= 1;
            e[c] = act ? sm + 600 * c + (8 * (4 * py + pz) + (4 * by + bz)) * 9 + jc : sm + 3000;
        }
        long long t0 = clock64();
        double acc = 0;
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
            double v[12];
#pragma unroll
            for (int q = 0; q < 12; ++q) {
                const int o = 144 * (q / 6) + 18 * ((q / 3) & 1) + 3 * (q % 3) + 600 * 0;
                v[q] = (e[0][o + rep] + e[1][o + rep]) + (e[2][o + rep] + e[3][o + rep]);
            }
#pragma unroll
            for (int q = 0; q < 12; ++q) sm[3200 + 32 * q + lane + 400 * rep] = v[q];
        }
        __syncwarp();
        t[7] = clock64() - t0;
        sink[threadIdx.x] += acc;
    }
    if (lane == 0)
        for (int i = 0; i < 8; ++i) out[warp * 8 + i] = t[i];
}
int main() {
    long long* d; double* s;
    cudaMalloc(&d, 64 * 8 * sizeof(long long)); cudaMalloc(&s, 4096 * sizeof(double));
    for (int nw : {1, 4, 8, 12}) {
        cudaMemset(d, 0, 64 * 8 * sizeof(long long));
        k<<<1, 32 * 12, 6144 * 8>>>(d, s, nw);
        cudaDeviceSynchronize();
        long long h[64 * 8];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("active warps %2d (1 CTA): dep DMMA %.1f cyc each | 72 DMMA (9 indep x 8) %lld cyc = %.1f each | dep DFMA %.1f | 12-chain DFMA x16: %lld = %.2f/instr | 48 LDS.64+adds %lld | dep LDS %.1f | 72 STS.128 %lld | 4 gather columns %lld\n",
               nw, h[0] / 64.0, h[1], h[1] / 72.0, h[2] / 64.0, h[3], h[3] / 192.0, h[4], h[5] / 32.0, h[6], h[7]);
    }
    printf("error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
