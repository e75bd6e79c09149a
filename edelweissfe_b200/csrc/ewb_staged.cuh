// Staged assembly for 8-node hexahedra: two streaming kernels instead of the fused sweep.
// STATUS (round 1): parity-green alternative, selected with EWB_FLAG_STAGED, NOT the default — measured 7.4 ms per
// 100^3 assembly (K1 2.3 ms, K2 5.1 ms, K2 is instruction bound: its index arithmetic is still evaluated at run time)
// against 2.43 ms for the fused sweep.  Kept because its gather sums every CSR value in the reference's own order.
//
//   K1  elementBlocksKernel   (any connectivity)  every element once (no halo recompute): phase A + tensor-pipe phase B
//                             exactly as in the sweep kernel, then the SYMMETRIC HALF of Ke (36 node-pair blocks of 3x3 =
//                             324 doubles instead of the reference's 576-entry VIJ slice) and Pe are written as
//                             structure-of-arrays  H[(pair*9 + comp)][e],  Pe[3a+i][e]  through a shared-memory
//                             transposition (4 consecutive elements = one 32-byte sector per row);
//   K2  boxGatherKernel       (BoxGen topology)   a CTA owns 32 consecutive nodes of one z-line; lane = node, so every
//                             load of H is a coalesced 256-byte run; every CSR value is the sum of its (at most 8) element
//                             contributions in ASCENDING ELEMENT ORDER — the reference's own updateCSR order
//                             (numerics/csrgenerator.pyx:100-115) — staged in shared memory and written with coalesced
//                             stores.  P and F are gathered the same way (nonlinearimplicitstatic.py:843-844).
//
// Compared with the reference's data flow (VIJ 4608 B/element + int32 slot map 2304 B/element) the intermediate is
// 2592 + 192 B/element and is read through L2 (each block serves two rows).
#pragma once
#include "ewb_sweep.cuh"

namespace ewb {

__host__ __device__ constexpr int pairIndex(int a, int b) { return a * 8 - a * (a - 1) / 2 + (b - a); }  // a <= b, 0..35
constexpr int H_ROWS = 324;   // 36 pairs x 9
constexpr int HP_ROWS = 348;  // + 24 residual entries

template <int MC, bool TL, int NWB>
__global__ void __launch_bounds__(NWB * 32) elementBlocksKernel(int64_t nEl, const int32_t* __restrict__ conn, const double* __restrict__ coords,
                                                              const double* __restrict__ U, const double* __restrict__ dU,
                                                              const double* __restrict__ stateRef, double* __restrict__ stateTemp,
                                                              double* __restrict__ H, double* __restrict__ Pe, MatParams mp, int* failFlag, int wantK) {
    using R = RecLayout<MC>;
    constexpr int TROW = 4;                               // elements per task
    constexpr int PER_WARP = 4 * R::PER_EL + 4 * 48 + HP_ROWS * TROW;
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* rec = smem + (size_t)warp * PER_WARP;
    double* stage = rec + 4 * R::PER_EL;                  // [4 elements][8 nodes][x,y,z,u0,u1,u2]
    double* T = stage + 4 * 48;                           // [HP_ROWS][4]
    const double* __restrict__ uSrc = TL ? U : dU;
    const int64_t cstride = nEl * 8;
    const int64_t nTasks = (nEl + 3) / 4;
    const int ak = lane >> 3, agp = lane & 7;
    // phase-B lane constants
    const int bq = lane & 3;
    const int na = rowNode(lane >> 2);
    double dNl[2][3];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        double xi, eta, zeta, w;
        Gauss<8>::get(4 * ks + bq, xi, eta, zeta, w);
        const double sa = NodeLC<8>::xi(na), sb = NodeLC<8>::eta(na), sc = NodeLC<8>::zeta(na);
        const double fx = 1.0 + sa * xi, fe = 1.0 + sb * eta, fz = 1.0 + sc * zeta;
        dNl[ks][0] = 0.125 * sb * fx * fz;
        dNl[ks][1] = 0.125 * sa * fe * fz;
        dNl[ks][2] = 0.125 * sc * fx * fe;
    }
    int trow[2];  // T row of the (a, b_t) block, -1 if it is the mirrored half
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int nb = rowNode(2 * bq + t);
        trow[t] = na <= nb ? pairIndex(na, nb) * 9 : -1;
    }
    for (int64_t task = (int64_t)blockIdx.x * NWB + warp; task < nTasks; task += (int64_t)gridDim.x * NWB) {
        const int64_t e0 = task * 4;
        // ---- stage the nodal data: lane = (element, local node) ----
        {
            const int64_t e = e0 + ak;
            if (e < nEl) {
                const int64_t n = conn[e * 8 + agp];
                double* d = stage + ak * 48 + agp * 6;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    d[c] = __ldg(coords + 3 * n + c);
                    d[3 + c] = __ldg(uSrc + 3 * n + c);
                }
            }
        }
        __syncwarp();
        // ---- phase A: lane = (element, Gauss point) ----
        {
            const int64_t e = e0 + ak;
            if (e < nEl) {
                const int64_t off = e * 8 + agp;
                gaussPointCompact<MC, TL, true>(rec + ak * R::PER_EL + agp * R::RS, stage + ak * 48, agp, mp, stateRef + off, stateTemp + off, cstride,
                                                true, failFlag);
            }
        }
        __syncwarp();
        // ---- phase B: one element at a time on the tensor pipe; symmetric half + residual into the transposition buffer ----
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            if (e0 + k >= nEl) break;
            double K0[9], K1[9], Pr[3];
            elementBlocks<MC>(rec + k * R::PER_EL, lane, dNl, mp, wantK != 0, K0, K1, Pr);
            if (wantK) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (trow[t] >= 0) {
                        const double* Kt = t ? K1 : K0;
#pragma unroll
                        for (int c = 0; c < 9; ++c) T[(trow[t] + c) * TROW + k] = Kt[c];
                    }
                }
            }
            if (bq == 0) {
#pragma unroll
                for (int i = 0; i < 3; ++i) T[(H_ROWS + na * 3 + i) * TROW + k] = Pr[i];
            }
        }
        __syncwarp();
        // ---- store: 8 rows x 4 elements per instruction = 8 full sectors ----
        const int kk = lane & 3;
        if (e0 + kk < nEl) {
            for (int row = lane >> 2; row < HP_ROWS; row += 8) {
                if (row < H_ROWS) {
                    if (wantK) H[(int64_t)row * nEl + e0 + kk] = T[row * TROW + kk];
                } else {
                    Pe[(int64_t)(row - H_ROWS) * nEl + e0 + kk] = T[row * TROW + kk];
                }
            }
        }
        __syncwarp();
    }
}

// BoxGen local node index from offsets (dx,dy,dz) (generators/boxgen.py:172-185)
__host__ __device__ constexpr int boxLocalNode(int dx, int dy, int dz) {
    // offsets of local nodes 0..7: (0,0,0),(0,0,1),(1,0,1),(1,0,0),(0,1,0),(0,1,1),(1,1,1),(1,1,0)
    return dy * 4 + (dx == 0 ? (dz == 0 ? 0 : 1) : (dz == 1 ? 2 : 3));
}

constexpr int GATHER_NODES = 32;

// One CTA = 32 consecutive nodes (iz0..iz0+31) of the z-line (ix, iy).  8 warps share the 243 outputs per node.
__global__ void __launch_bounds__(256) boxGatherKernel(int nX, int nY, int nZ, const double* __restrict__ H, const double* __restrict__ Pe,
                                                       double* __restrict__ data, double* __restrict__ P, double* __restrict__ F, int accumulatePF,
                                                       int wantK) {
    extern __shared__ double tile[];  // [GATHER_NODES][243]
    const int NY = nY + 1, NZ = nZ + 1, NX = nX + 1;
    const int zChunks = (NZ + GATHER_NODES - 1) / GATHER_NODES;
    int bid = blockIdx.x;
    const int zc = bid % zChunks; bid /= zChunks;
    const int iy = bid % NY;
    const int ix = bid / NY;
    const int iz0 = zc * GATHER_NODES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int iz = iz0 + lane;
    const bool nodeValid = iz < NZ;
    const int64_t nEl = (int64_t)nX * nY * nZ;
    const int64_t strideX = (int64_t)nY * nZ;
    // element whose local node 0 is this node (may be out of range: checked per contribution)
    const int64_t eA = (int64_t)ix * strideX + (int64_t)iy * nZ + iz;

    if (wantK) {
        for (int out = warp; out < 243; out += 8) {
            const int i = out / 81, s27 = (out % 81) / 3, j = out % 3;
            const int dx = s27 / 9 - 1, dy = (s27 / 3) % 3 - 1, dz = s27 % 3 - 1;
            double acc = 0.0;
            // contributing elements e = A - o with o and o + delta in {0,1}^3; ascending e <=> descending o
#pragma unroll
            for (int oi = 7; oi >= 0; --oi) {
                const int ox = oi >> 2, oy = (oi >> 1) & 1, oz = oi & 1;
                const int bx = ox + dx, by = oy + dy, bz = oz + dz;
                if (bx < 0 || bx > 1 || by < 0 || by > 1 || bz < 0 || bz > 1) continue;  // warp uniform
                const int ex = ix - ox, ey = iy - oy;
                if (ex < 0 || ex >= nX || ey < 0 || ey >= nY) continue;                  // warp uniform
                const int a = boxLocalNode(ox, oy, oz), b = boxLocalNode(bx, by, bz);
                const int row = a <= b ? pairIndex(a, b) * 9 + i * 3 + j : pairIndex(b, a) * 9 + j * 3 + i;
                const int ez = iz - oz;
                if (nodeValid && ez >= 0 && ez < nZ) acc += H[(int64_t)row * nEl + (eA - (ox * strideX + oy * nZ + oz))];
            }
            tile[lane * 243 + out] = acc;
        }
    }
    __syncthreads();
    // ---- coalesced write-out with boundary compaction (closed-form CSR row bases, see ewb_sweep.cuh) ----
    auto pre = [](int v) { return v == 0 ? 0 : 3 * v - 1; };
    const int totY = 3 * NY - 2, totZ = 3 * NZ - 2;
    const int cx = (ix > 0) + 1 + (ix < NX - 1), cy = (iy > 0) + 1 + (iy < NY - 1);
    const int nNodes = min(GATHER_NODES, NZ - iz0);
    if (wantK) {
        for (int idx = threadIdx.x; idx < nNodes * 243; idx += 256) {
            const int l = idx / 243, out = idx % 243;
            const int i = out / 81, s27 = (out % 81) / 3, j = out % 3;
            const int dx = s27 / 9 - 1, dy = (s27 / 3) % 3 - 1, dz = s27 % 3 - 1;
            const int z = iz0 + l;
            if (ix + dx < 0 || ix + dx >= NX || iy + dy < 0 || iy + dy >= NY || z + dz < 0 || z + dz >= NZ) continue;
            const int cz = (z > 0) + 1 + (z < NZ - 1);
            const int deg = cx * cy * cz;
            const int64_t base = 9 * ((int64_t)pre(ix) * totY * totZ + (int64_t)cx * ((int64_t)pre(iy) * totZ + (int64_t)cy * pre(z)));
            const int slot = ((dx + (ix > 0 ? 1 : 0)) * cy + (dy + (iy > 0 ? 1 : 0))) * cz + (dz + (z > 0 ? 1 : 0));
            data[base + (int64_t)i * 3 * deg + 3 * slot + j] = tile[idx];
        }
    }
    // ---- residual: P[el] += Pe ; F[el] += |Pe| in ascending element order ----
    if (threadIdx.x < nNodes * 3) {
        const int l = threadIdx.x / 3, c = threadIdx.x % 3;
        const int z = iz0 + l;
        double p = 0.0, f = 0.0;
#pragma unroll
        for (int oi = 7; oi >= 0; --oi) {
            const int ox = oi >> 2, oy = (oi >> 1) & 1, oz = oi & 1;
            const int ex = ix - ox, ey = iy - oy, ez = z - oz;
            if (ex < 0 || ex >= nX || ey < 0 || ey >= nY || ez < 0 || ez >= nZ) continue;
            const int a = boxLocalNode(ox, oy, oz);
            const double v = Pe[(int64_t)(a * 3 + c) * nEl + ((int64_t)ex * strideX + (int64_t)ey * nZ + ez)];
            p += v;
            f += fabs(v);
        }
        const int64_t dof = 3 * (((int64_t)ix * NY + iy) * NZ + z) + c;
        if (accumulatePF) {
            P[dof] += p;
            F[dof] += f;
        } else {
            P[dof] = p;
            F[dof] = f;
        }
    }
}

struct StagedPlan {
    double* H = nullptr;
    double* Pe = nullptr;
    int64_t nElAlloc = 0;
    int nSM = 148;

    void release() {
        cudaFree(H);
        cudaFree(Pe);
        H = Pe = nullptr;
        nElAlloc = 0;
    }

    template <int MC, bool TL>
    int launchK1(int64_t nEl, const int32_t* conn, const MatParams& mp, const ewb_buffers* b, int* failFlag, int wantK, cudaStream_t st) {
        using R = RecLayout<MC>;
        constexpr int NWB = 4;
        auto kern = elementBlocksKernel<MC, TL, NWB>;
        const size_t smem = (size_t)NWB * (4 * R::PER_EL + 4 * 48 + HP_ROWS * 4) * sizeof(double);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return EWB_ERR_CUDA;
        int perSM = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, NWB * 32, smem);
        const int64_t nTasks = (nEl + 3) / 4;
        const int64_t grid = std::min<int64_t>((nTasks + NWB - 1) / NWB, (int64_t)nSM * std::max(perSM, 1) * 4);
        kern<<<(unsigned)grid, NWB * 32, smem, st>>>(nEl, conn, b->coords, b->U, b->dU, b->state_ref, b->state_temp, H, Pe, mp, failFlag, wantK);
        return cudaGetLastError() == cudaSuccess ? EWB_OK : EWB_ERR_CUDA;
    }

    int launch(int elType, int mc, int64_t nEl, int64_t nX, int64_t nY, int64_t nZ, const int32_t* conn, const MatParams& mp, const ewb_buffers* b,
               int* failFlag, int flags, cudaStream_t st, int* launches) {
        if (nEl > nElAlloc) {
            release();
            if (cudaMalloc((void**)&H, (size_t)H_ROWS * nEl * sizeof(double)) != cudaSuccess) return EWB_ERR_CUDA;
            if (cudaMalloc((void**)&Pe, (size_t)24 * nEl * sizeof(double)) != cudaSuccess) return EWB_ERR_CUDA;
            nElAlloc = nEl;
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, dev);
        }
        const int wantK = (flags & EWB_FLAG_NO_STIFFNESS) ? 0 : 1;
        int rc = EWB_ERR_UNSUPPORTED;
        if (elType == EWB_C3D8 && mc == MC_LE) rc = launchK1<MC_LE, false>(nEl, conn, mp, b, failFlag, wantK, st);
        else if (elType == EWB_C3D8 && mc == MC_VM) rc = launchK1<MC_VM, false>(nEl, conn, mp, b, failFlag, wantK, st);
        else if (elType == EWB_C3D8TL && mc == MC_NH) rc = launchK1<MC_NH, true>(nEl, conn, mp, b, failFlag, wantK, st);
        if (rc != EWB_OK) return rc;
        const int NY = (int)nY + 1, NX = (int)nX + 1, NZ = (int)nZ + 1;
        const int64_t grid = (int64_t)NX * NY * ((NZ + GATHER_NODES - 1) / GATHER_NODES);
        const size_t smemG = (size_t)GATHER_NODES * 243 * sizeof(double);
        if (cudaFuncSetAttribute(boxGatherKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemG) != cudaSuccess) return EWB_ERR_CUDA;
        boxGatherKernel<<<(unsigned)grid, 256, smemG, st>>>((int)nX, (int)nY, (int)nZ, H, Pe, b->csr_data, b->P, b->F, (flags & EWB_FLAG_ACCUMULATE_PF) ? 1 : 0, wantK);
        if (cudaGetLastError() != cudaSuccess) return EWB_ERR_CUDA;
        *launches = 2;
        return EWB_OK;
    }
};

}  // namespace ewb
