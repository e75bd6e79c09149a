// Fused BoxGen sweep assembly for 8-node hexahedra: ONE kernel does what the reference does in
// NIST.computeElements + CSRGenerator.updateCSR (solvers/nonlinearimplicitstatic.py:794-849,
// numerics/csrgenerator.pyx:100-115) without ever materialising the VIJ triple:
//
//   * a CTA owns a TY x TZ tile of node columns (y,z) and a chunk of node planes in x; it sweeps the
//     element planes along x, keeping the CSR rows of the two node planes adjacent to the current
//     element plane in shared memory (4 segments of 27 3x3-blocks per node column: plane i {dx=0,
//     dx=+1}, plane i+1 {dx=-1, dx=0}); finished segments are written to the CSR value array with
//     coalesced stores, each CSR value exactly once, no atomics, fixed summation order;
//   * elements on the tile rim are recomputed by the neighbouring CTA (halo recompute), so there is
//     no inter-CTA communication;
//   * phase A (thread = element x Gauss point) is shared with the generic path (ewb_tile.cuh);
//   * phase B runs on the FP64 tensor pipe: per element, M[(i,a),(j,b)] = sum_gp (c_gp g_a,i) g_b,j
//     is a (24x8)x(8x24) product = 3x3 tiles of mma.m8n8k4.f64 with the tile index = component
//     pair (i,j), so that every lane ends up with complete 3x3 node blocks and the isotropic /
//     rank-one / Neo-Hooke tangent assembly is lane-local;
//   * the 2x2 element patch of a warp has 4 different colours; all warps process the same colour
//     between two barriers, so shared-memory accumulation is race free and deterministic.
#pragma once
#include "../../include/edelweiss_b200.h"
#include "ewb_tile.cuh"

namespace ewb {

struct SweepArgs {
    int nX, nY, nZ;  // elements
    int chunkLen, nChunks, tilesY, tilesZ;
    const double* coords;
    const double* U;
    const double* dU;
    const double* stateRef;
    double* stateTemp;
    double* data;
    double* P;
    double* F;
    const int64_t* adjPtr;
    MatParams mp;
    int* failFlag;
    int wantK;
    int accumulatePF;
};

template <int MC>
struct SweepLayout {
    static constexpr bool HASQ = false;
    static constexpr int GST = 28;  // == 4 (mod 16): conflict-free mma fragment loads (lane -> node 3a, gp 28k)
    static constexpr int NCO = (MC == MC_LE) ? 10 : (MC == MC_VM ? 16 : 28);
    static constexpr int OFF_G = 0;
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_CO = 8 * GST;
    static constexpr int RAW = OFF_CO + 8 * NCO;
    static constexpr int PER_EL = RAW + ((2 - RAW % 16) + 16) % 16;  // == 2 (mod 16)
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// local node a of a BoxGen Hexa8: offsets (dx,dy,dz) (generators/boxgen.py:172-185)
__device__ __forceinline__ int ndx(int a) { return (a >> 1) & 1; }
__device__ __forceinline__ int ndy(int a) { return (a >> 2) & 1; }
__device__ __forceinline__ int ndz(int a) { return (a ^ (a >> 1)) & 1; }

template <int MC, bool TL, int TY, int TZ, int NW>
__global__ void __launch_bounds__(NW * 32, 1) sweepKernel(const SweepArgs A) {
    using L = SweepLayout<MC>;
    constexpr int NPY = (TY + 1) / 2, NPZ = (TZ + 1) / 2, NP = NPY * NPZ, NB = (NP + NW - 1) / NW;
    constexpr int NCOL = TY * TZ;
    constexpr int SEG = NCOL * 81;
    constexpr int NT = NW * 32;
    static_assert((TY & 1) == 1 && (TZ & 1) == 1, "tile edge must be odd (2x2 element patches)");

    extern __shared__ double smem[];
    double* seg0a = smem;            // dx=0 segment, ping
    double* seg0b = seg0a + SEG;     // dx=0 segment, pong
    double* segP = seg0b + SEG;      // lower plane, dx=+1
    double* segM = segP + SEG;       // upper plane, dx=-1
    double* pfA = segM + SEG;        // [NCOL][6] P,F of the lower plane
    double* pfB = pfA + NCOL * 6;    // upper plane
    double* tables = pfB + NCOL * 6; // [NW][4][PER_EL]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NX = A.nX + 1, NY = A.nY + 1, NZ = A.nZ + 1;

    // work item -> (chunk, tile)
    int item = blockIdx.x;
    const int tz = item % A.tilesZ; item /= A.tilesZ;
    const int ty = item % A.tilesY; item /= A.tilesY;
    const int chunk = item;
    const int y0 = ty * TY, z0 = tz * TZ;
    const int ny = min(TY, NY - y0), nz = min(TZ, NZ - z0);  // owned node columns
    const int xa = chunk * A.chunkLen, xb = min(xa + A.chunkLen, NX);
    const int exBegin = max(xa - 1, 0), exEnd = min(xb - 1, A.nX - 1);

    for (int i = tid; i < 4 * SEG + 12 * NCOL; i += NT) smem[i] = 0.0;
    __syncthreads();

    double* lo0 = seg0a;
    double* hi0 = seg0b;
    double* pfLo = pfA;
    double* pfHi = pfB;
    const int64_t cstride = (int64_t)A.nX * A.nY * A.nZ * 8;
    const double* __restrict__ uSrc = TL ? A.U : A.dU;

    // CSR row base of node (ix,iy,iz) in closed form: 9 * (sum of the degrees of all preceding nodes);
    // deg = cx*cy*cz with c = 2 on a face, 3 inside (== plan->adjPtr, checked by the parity tests)
    auto pre = [](int i) { return i == 0 ? 0 : 3 * i - 1; };
    const int64_t totY = 3 * NY - 2, totZ = 3 * NZ - 2;
    // lane-fixed part of the flush mapping: lane e < 27 copies entry (s9 = e/3, j = e%3) of a 27-entry sub-row
    const int fl_s9 = lane / 3, fl_j = lane % 3, fl_dy = fl_s9 / 3 - 1, fl_dz = fl_s9 % 3 - 1;
    // flush the finished segments (dx = -1, 0, +1; nullptr = not finished) of every owned node of plane ix, clear them
    auto flushPlane = [&](int ix, double* sM, double* s0, double* sP) {
        const int cx = (ix > 0) + 1 + (ix < NX - 1);
        for (int col = warp; col < NCOL; col += NW) {
            const int ly = col / TZ, lz = col % TZ;
            if (ly >= ny || lz >= nz) continue;
            const int iy = y0 + ly, iz = z0 + lz;
            const int cy = (iy > 0) + 1 + (iy < NY - 1), cz = (iz > 0) + 1 + (iz < NZ - 1);
            const int deg = cx * cy * cz;
            const int64_t base = 9 * ((int64_t)pre(ix) * totY * totZ + (int64_t)cx * (pre(iy) * totZ + (int64_t)cy * pre(iz)));
            const bool nbValid = lane < 27 && iy + fl_dy >= 0 && iy + fl_dy < NY && iz + fl_dz >= 0 && iz + fl_dz < NZ;
            const int slotYZ = (fl_dy + (iy > 0 ? 1 : 0)) * cz + fl_dz + (iz > 0 ? 1 : 0);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                double* seg = d == 0 ? sM : (d == 1 ? s0 : sP);
                if (seg == nullptr) continue;
                const int dx = d - 1;
                const bool dxValid = ix + dx >= 0 && ix + dx < NX;
                const int rx = dx + (ix > 0 ? 1 : 0);
                double* src = seg + col * 81 + lane;
                double* dst = A.data + base + 3 * (rx * cy * cz + slotYZ) + fl_j;
                if (lane < 27) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const double v = src[i * 27];
                        src[i * 27] = 0.0;
                        if (dxValid && nbValid) dst[(int64_t)i * 3 * deg] = v;
                    }
                }
            }
        }
    };
    auto flushPF = [&](double* pf, int ix) {
        for (int t = tid; t < NCOL * 3; t += NT) {
            const int col = t / 3, i = t % 3;
            const int ly = col / TZ, lz = col % TZ;
            const double p = pf[col * 6 + i], f = pf[col * 6 + 3 + i];
            pf[col * 6 + i] = 0.0;
            pf[col * 6 + 3 + i] = 0.0;
            if (ly >= ny || lz >= nz) continue;
            const int64_t dof = 3 * ((((int64_t)ix * NY + (y0 + ly)) * NZ) + (z0 + lz)) + i;
            if (A.accumulatePF) {
                A.P[dof] += p;
                A.F[dof] += f;
            } else {
                A.P[dof] = p;
                A.F[dof] = f;
            }
        }
    };

    for (int ex = exBegin; ex <= exEnd; ++ex) {
        const bool loOwned = ex >= xa, hiOwned = (ex + 1) < xb;
#pragma unroll 1
        for (int bt = 0; bt < NB; ++bt) {
            const int p = bt * NW + warp;
            const int pyq = p / NPZ, pzq = p % NPZ;
            double* wt = tables + (size_t)warp * 4 * L::PER_EL;
            // ---------------- phase A: lane = (element k of the patch, Gauss point) ----------------
            if (p < NP) {
                const int k = lane >> 3, gp = lane & 7;
                const int py = 2 * pyq + (k >> 1), pz = 2 * pzq + (k & 1);
                const int ey = y0 - 1 + py, ez = z0 - 1 + pz;
                const bool valid = ey >= 0 && ey < A.nY && ez >= 0 && ez < A.nZ && py <= ny && pz <= nz;
                if (valid) {
                    double X[24], uu[24];
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        const int64_t n = ((int64_t)(ex + ndx(a)) * NY + (ey + ndy(a))) * NZ + (ez + ndz(a));
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            X[a * 3 + c] = __ldg(A.coords + 3 * n + c);
                            uu[a * 3 + c] = __ldg(uSrc + 3 * n + c);
                        }
                    }
                    const int64_t e = ((int64_t)ex * A.nY + ey) * A.nZ + ez;
                    const int64_t off = e * 8 + gp;
                    const bool writeState = loOwned && py >= 1 && pz >= 1;
                    gaussPointL<L, 8, 8, MC, TL>(wt + k * L::PER_EL, X, uu, gp, A.mp, A.stateRef + off, A.stateTemp + off, cstride, writeState, A.failFlag);
                }
            }
            __syncthreads();
            // ---------------- phase B: 4 colour rounds, one element per warp per round -------------
#pragma unroll 1
            for (int k = 0; k < 4; ++k) {
                const int py = 2 * pyq + (k >> 1), pz = 2 * pzq + (k & 1);
                const int ey = y0 - 1 + py, ez = z0 - 1 + pz;
                const bool valid = p < NP && ey >= 0 && ey < A.nY && ez >= 0 && ez < A.nZ && py <= ny && pz <= nz;
                if (valid) {
                    const double* T = wt + k * L::PER_EL;
                    const int r = lane >> 2, q = lane & 3;
                    double g[2][3], K0[9], K1[9], Pr[3] = {0, 0, 0};
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                        for (int c = 0; c < 3; ++c) g[ks][c] = T[L::OFF_G + (4 * ks + q) * L::GST + r * 3 + c];
                    // residual row partial: (-w detJ S) v_r over this lane's two Gauss points
                    if constexpr (MC == MC_LE) {
                        double c[3][3][2];
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 3; ++j) c[i][j][0] = c[i][j][1] = 0.0;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const double* co = T + L::OFF_CO + (4 * ks + q) * L::NCO;
                            const double w = co[0];
                            const double* S = co + 4;
                            Pr[0] += S[0] * g[ks][0] + S[3] * g[ks][1] + S[4] * g[ks][2];
                            Pr[1] += S[3] * g[ks][0] + S[1] * g[ks][1] + S[5] * g[ks][2];
                            Pr[2] += S[4] * g[ks][0] + S[5] * g[ks][1] + S[2] * g[ks][2];
                            if (A.wantK) {
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    const double ai = w * g[ks][i];
#pragma unroll
                                    for (int j = 0; j < 3; ++j) dmma(c[i][j], ai, g[ks][j]);
                                }
                            }
                        }
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            double* Kt = t ? K1 : K0;
                            const double tr = A.mp.G * (c[0][0][t] + c[1][1][t] + c[2][2][t]);
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int j = 0; j < 3; ++j) Kt[i * 3 + j] = A.mp.lambda * c[i][j][t] + A.mp.G * c[j][i][t] + (i == j ? tr : 0.0);
                        }
                    } else if constexpr (MC == MC_VM) {
                        double c1[3][3][2], c2[3][3][2];
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 3; ++j) c1[i][j][0] = c1[i][j][1] = c2[i][j][0] = c2[i][j][1] = 0.0;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const double* co = T + L::OFF_CO + (4 * ks + q) * L::NCO;
                            const double cl = co[0], cm = co[1], ca = co[2];
                            const double* S = co + 4;
                            const double* n = co + 10;
                            Pr[0] += S[0] * g[ks][0] + S[3] * g[ks][1] + S[4] * g[ks][2];
                            Pr[1] += S[3] * g[ks][0] + S[1] * g[ks][1] + S[5] * g[ks][2];
                            Pr[2] += S[4] * g[ks][0] + S[5] * g[ks][1] + S[2] * g[ks][2];
                            if (A.wantK) {
                                double pv[3];
                                pv[0] = n[0] * g[ks][0] + n[3] * g[ks][1] + n[4] * g[ks][2];
                                pv[1] = n[3] * g[ks][0] + n[1] * g[ks][1] + n[5] * g[ks][2];
                                pv[2] = n[4] * g[ks][0] + n[5] * g[ks][1] + n[2] * g[ks][2];
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    const double li = cl * g[ks][i], mi = cm * g[ks][i], ri = ca * pv[i];
#pragma unroll
                                    for (int j = 0; j < 3; ++j) {
                                        dmma(c1[i][j], li, g[ks][j]);
                                        dmma(c1[i][j], ri, pv[j]);
                                        dmma(c2[i][j], mi, g[ks][j]);
                                    }
                                }
                            }
                        }
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            double* Kt = t ? K1 : K0;
                            const double tr = c2[0][0][t] + c2[1][1][t] + c2[2][2][t];
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int j = 0; j < 3; ++j) Kt[i * 3 + j] = c1[i][j][t] + c2[j][i][t] + (i == j ? tr : 0.0);
                        }
                    } else {
                        double c1[3][3][2], c2[3][3][2], d0[3][2];
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            d0[i][0] = d0[i][1] = 0.0;
#pragma unroll
                            for (int j = 0; j < 3; ++j) c1[i][j][0] = c1[i][j][1] = c2[i][j][0] = c2[i][j][1] = 0.0;
                        }
                        const bool wb = (A.mp.kind == EWB_MAT_NEOHOOKE_WB);
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const double* co = T + L::OFF_CO + (4 * ks + q) * L::NCO;
                            const double k0 = co[0], k1 = co[1], k2 = co[2], k4 = co[3];
                            const double* S = co + 4;
                            const double* F = co + 10;
                            const double* iF = co + 19;
                            double nv[3];
#pragma unroll
                            for (int m = 0; m < 3; ++m) nv[m] = g[ks][0] * iF[m] + g[ks][1] * iF[3 + m] + g[ks][2] * iF[6 + m];
                            Pr[0] += S[0] * nv[0] + S[3] * nv[1] + S[4] * nv[2];
                            Pr[1] += S[3] * nv[0] + S[1] * nv[1] + S[5] * nv[2];
                            Pr[2] += S[4] * nv[0] + S[5] * nv[1] + S[2] * nv[2];
                            if (A.wantK) {
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    const double a1 = k1 * nv[i], a2 = k2 * nv[i];
                                    dmma(d0[i], k0 * g[ks][i], g[ks][i]);
#pragma unroll
                                    for (int j = 0; j < 3; ++j) {
                                        dmma(c1[i][j], a1, nv[j]);
                                        dmma(c2[i][j], a2, nv[j]);
                                    }
                                }
                                if (wb) {  // W_b: + c4 (f_a n_b^T + n_a f_b^T), f = F g
                                    double fv[3];
#pragma unroll
                                    for (int i = 0; i < 3; ++i) fv[i] = F[i * 3] * g[ks][0] + F[i * 3 + 1] * g[ks][1] + F[i * 3 + 2] * g[ks][2];
#pragma unroll
                                    for (int i = 0; i < 3; ++i)
#pragma unroll
                                        for (int j = 0; j < 3; ++j) {
                                            dmma(c1[i][j], k4 * fv[i], nv[j]);
                                            dmma(c1[i][j], k4 * nv[i], fv[j]);
                                        }
                                }
                            }
                        }
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            double* Kt = t ? K1 : K0;
                            const double tr = d0[0][t] + d0[1][t] + d0[2][t];
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int j = 0; j < 3; ++j) Kt[i * 3 + j] = c1[i][j][t] + c2[j][i][t] + (i == j ? tr : 0.0);
                        }
                    }
                    // reduce the residual row over the 4 lanes (Gauss-point pairs) of node r
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        Pr[i] += __shfl_xor_sync(0xffffffffu, Pr[i], 1);
                        Pr[i] += __shfl_xor_sync(0xffffffffu, Pr[i], 2);
                    }
                    // ---- accumulate into the owned rows ----
                    const int a = r;
                    const int ly = py - 1 + ndy(a), lz = pz - 1 + ndz(a);
                    const bool owned = ly >= 0 && ly < ny && lz >= 0 && lz < nz && (ndx(a) ? hiOwned : loOwned);
                    if (owned) {
                        const int col = ly * TZ + lz;
                        if (q == 0) {
                            double* pf = (ndx(a) ? pfHi : pfLo) + col * 6;
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                pf[i] += Pr[i];
                                pf[3 + i] += fabs(Pr[i]);
                            }
                        }
                        if (A.wantK) {
#pragma unroll
                            for (int t = 0; t < 2; ++t) {
                                const int b = 2 * q + t;
                                const int rx = ndx(b) - ndx(a), ry = ndy(b) - ndy(a), rz = ndz(b) - ndz(a);
                                double* seg = ndx(a) ? (rx == 0 ? hi0 : segM) : (rx == 0 ? lo0 : segP);
                                double* dst = seg + col * 81 + ((ry + 1) * 3 + rz + 1) * 3;
                                const double* Kt = t ? K1 : K0;
#pragma unroll
                                for (int i = 0; i < 3; ++i)
#pragma unroll
                                    for (int j = 0; j < 3; ++j) dst[i * 27 + j] += Kt[i * 3 + j];
                            }
                        }
                    }
                }
                __syncthreads();
            }
        }
        // ---------------- flush the finished segments, rotate ----------------
        if (A.wantK) {
            if (loOwned) flushPlane(ex, nullptr, lo0, segP);
            if (hiOwned) flushPlane(ex + 1, segM, nullptr, nullptr);
        }
        if (loOwned) flushPF(pfLo, ex);
        __syncthreads();
        {
            double* t0 = lo0; lo0 = hi0; hi0 = t0;
            double* t1 = pfLo; pfLo = pfHi; pfHi = t1;
        }
    }
    if (xb == NX && exEnd + 1 == NX - 1) {  // the last node plane has no element plane above it
        if (A.wantK) flushPlane(NX - 1, nullptr, lo0, nullptr);
        flushPF(pfLo, NX - 1);
    }
}

struct SweepPlan {
    int64_t nX = 0, nY = 0, nZ = 0;
    const int64_t* adjPtr = nullptr;

    int build(int64_t nx, int64_t ny, int64_t nz) {
        nX = nx; nY = ny; nZ = nz;
        return 0;
    }
    void release() {}

    template <int MC, bool TL, int TY, int TZ, int NW>
    int launchT(const MatParams& mp, const ewb_buffers* b, int* failFlag, int flags, cudaStream_t st) {
        using Lay = SweepLayout<MC>;
        SweepArgs a;
        a.nX = (int)nX; a.nY = (int)nY; a.nZ = (int)nZ;
        a.tilesY = (int)((nY + 1 + TY - 1) / TY);
        a.tilesZ = (int)((nZ + 1 + TZ - 1) / TZ);
        // chunks along x: enough CTAs for >= ~6 waves of 148 SMs, at least ~12 planes per chunk
        const int64_t tiles = (int64_t)a.tilesY * a.tilesZ;
        int nChunks = (int)std::max<int64_t>(1, std::min<int64_t>((nX + 1) / 12, (148 * 6 + tiles - 1) / tiles));
        a.chunkLen = (int)((nX + 1 + nChunks - 1) / nChunks);
        a.nChunks = (int)((nX + 1 + a.chunkLen - 1) / a.chunkLen);
        a.coords = b->coords; a.U = b->U; a.dU = b->dU; a.stateRef = b->state_ref; a.stateTemp = b->state_temp;
        a.data = b->csr_data; a.P = b->P; a.F = b->F; a.adjPtr = adjPtr; a.mp = mp; a.failFlag = failFlag;
        a.wantK = (flags & EWB_FLAG_NO_STIFFNESS) ? 0 : 1;
        a.accumulatePF = (flags & EWB_FLAG_ACCUMULATE_PF) ? 1 : 0;
        auto kern = sweepKernel<MC, TL, TY, TZ, NW>;
        const size_t smem = ((size_t)4 * TY * TZ * 81 + 12 * TY * TZ + (size_t)NW * 4 * Lay::PER_EL) * sizeof(double);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return EWB_ERR_CUDA;
        const int64_t grid = tiles * a.nChunks;
        kern<<<(unsigned)grid, NW * 32, smem, st>>>(a);
        return cudaGetLastError() == cudaSuccess ? EWB_OK : EWB_ERR_CUDA;
    }

    int launch(int elType, int mc, const MatParams& mp, const ewb_buffers* b, int* failFlag, int flags, cudaStream_t st, int* launches) {
        int rc = EWB_ERR_UNSUPPORTED;
        if (elType == EWB_C3D8 && mc == MC_LE) rc = launchT<MC_LE, false, 7, 7, 8>(mp, b, failFlag, flags, st);
        else if (elType == EWB_C3D8 && mc == MC_VM) rc = launchT<MC_VM, false, 7, 7, 8>(mp, b, failFlag, flags, st);
        else if (elType == EWB_C3D8TL && mc == MC_NH) rc = launchT<MC_NH, true, 7, 7, 4>(mp, b, failFlag, flags, st);
        if (rc == EWB_OK) *launches = 1;
        return rc;
    }
};

}  // namespace ewb
