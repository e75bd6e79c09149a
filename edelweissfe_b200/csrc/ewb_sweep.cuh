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
//   * phase A (lane = element x Gauss point, 4 elements of a 2x2 patch per warp): J, J^-1, strain
//     increment / F, constitutive update, state write-back; publishes per Gauss point only J^-1, the
//     tangent coefficients and -w detJ*stress (17..37 doubles) — grad N is rebuilt in phase B from J^-1
//     and lane-constant shape-function derivatives;
//   * phase B (one element per warp at a time) runs on the FP64 tensor pipe: M[(i,a),(j,b)] =
//     sum_gp (c_gp g_a,i) g_b,j is a (24x8)x(8x24) product = 3x3 tiles of mma.m8n8k4.f64 with the tile
//     index = component pair (i,j), so that every lane ends up with complete 3x3 node blocks and the
//     isotropic / rank-one / Neo-Hooke tangent assembly is lane-local;
//   * the 4 elements of a warp's patch have 4 different colours; an element of colour k is accumulated
//     only after the 8 neighbouring patches finished colour k-1 (flags in shared memory, no CTA
//     barrier), so shared-memory accumulation is race free and its order is fixed (deterministic).
#pragma once
#include "../../include/edelweiss_b200.h"
#include "ewb_tile.cuh"

namespace ewb {

struct SweepArgs {
    int nX, nY, nZ;  // elements
    int chunkLen, nChunks, tilesY, tilesZ;
    int tileRows;  // row-pipelined kernel: node rows (y) per tile
    int chunkBase;  // row-pipelined kernel: first x-chunk of this launch (ewb_assemble_chunks: pipelined host I/O)
    const double* coords;
    const double* U;
    const double* dU;
    const double* stateRef;
    double* stateTemp;
    double* data;
    double* P;
    double* F;
    MatParams mp;
    int* failFlag;
    int wantK;
    int accumulatePF;
    int spinNs;  // back-off of the flag polling loops (tuning knob, EWB_SPIN_NS)
    // Slab interface (multi-GPU): when set, the rows / P / F of the LAST node plane (the ghost plane owned by the upper
    // neighbour) are stored straight into that neighbour's receive buffers over NVLink instead of the local tail.
    double* peerData;
    double* peerP;
    double* peerF;
    long long* timing;  // optional [gridDim][NW][12] cycle counters (EWB_TIMING builds only)
};

#ifdef EWB_TIMING
#define EWB_TIC(t) const long long t = clock64()
#define EWB_ACC(slot, t0) tacc[slot] += clock64() - (t0)
#else
#define EWB_TIC(t)
#define EWB_ACC(slot, t0)
#endif

// Per-Gauss-point record published by phase A (doubles):
//   LE: [0..8] J^-1 | [9] w detJ               | [10..15] -w detJ sigma
//   VM: [0..8] J^-1 | [9..11] w(lam', mu', -a) | [12..17] -w detJ sigma | [18..23] n
//   NH: [0..8] J^-1 | [9..12] w(c0,c1,c2,c4)   | [13..18] -w detJ tau   | [19..27] F^-1 | [28..36] F
// odd record stride + element stride == 8 (mod 16): conflict-free phase-A stores.
template <int MC>
struct RecLayout {
    static constexpr int C_CO = 9;
    static constexpr int C_S = (MC == MC_LE) ? 10 : (MC == MC_VM ? 12 : 13);
    static constexpr int C_X = C_S + 6;  // VM: n ; NH: F^-1 then F
    static constexpr int RS = (MC == MC_LE) ? 17 : (MC == MC_VM ? 25 : 37);
    static constexpr int PER_EL = 8 * RS;
};

// Linear-elastic producer/consumer kernel: the producers publish, per element, the scaled shape-function gradients already in
// the mma fragment order, h_a = sqrt(w detJ) grad N_a (so that M = sum_gp h h^T needs one operand only), and S' = -sqrt(w detJ) sigma
// (residual row P_a = sum_gp S' h_a).  A consumer lane then loads 6 + 12 doubles per element instead of 34, and the gradient
// arithmetic moves from the three-consumer scheduler to the producers' schedulers.
//   H[c][ks][row r = rowNode(a)][q]  at (2 c + ks) * KS + 4 r + q     (Gauss point 4 ks + q; one warp load = 32 consecutive doubles)
//   S'[gp][6]                        at OFF_S + 6 gp
// KS == 4, PER_EL == 8 (mod 16): the producer's stores (lane = element x Gauss point) hit 16 different banks per half-warp.
struct RecLayoutH {
    static constexpr int KS = 36;
    static constexpr int OFF_S = 6 * KS;
    static constexpr int PER_EL = OFF_S + 48;  // 264
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Tile-transpose identity.  Every product accumulated here has the form M[(i,a),(j,b)] = sum_gp c u_a,i v_b,j with u == v (or the
// sum of two such terms with the roles swapped), so tile (j,i) is the transpose of tile (i,j) in the node indices:
// M[(j,a),(i,b)] = M[(i,b),(j,a)].  Only the tiles i <= j are computed on the tensor pipe (6 instead of 9 DMMA per k-step); the
// lane (row r, q) obtains its entries of tile (j,i), columns 2q + t, from lane (row 2q + t, q' = r >> 1), slot r & 1, of tile (i,j).
// Measured on B200 (round 2): parity green, but SLOWER in every kernel (LE 598 -> 590, von Mises 334 -> 324, Neo-Hooke 413 -> 395
// Melem/s): the 24 shuffles per accumulator set sit on the warp's critical path, while the FP64 pipe was not the binding unit.
// Off by default; -DEWB_TILE_SYMMETRY=1 enables it.
#ifndef EWB_TILE_SYMMETRY
#define EWB_TILE_SYMMETRY 0
#endif
// Residual row of the fragment-order (linear elastic) records on the tensor pipe; valid in the lanes with q == 0 (all callers use those).
// Measured on B200 (round 2), parity green: row-pipelined kernel 647 -> 614 Melem/s (its tensor warps are FP64-pipe critical: +6 DMMA
// cost more than 18 DFMA + 12 shuffles), first-generation sweep 551 -> 561.  Off by default; -DEWB_RESIDUAL_DMMA=1 enables it.
#ifndef EWB_RESIDUAL_DMMA
#define EWB_RESIDUAL_DMMA 0
#endif
__device__ __forceinline__ void mirrorTiles(double (&c)[3][3][2], int lane) {
    const int r = lane >> 2, q = lane & 3;
    const int src0 = 4 * (2 * q) + (r >> 1), src1 = src0 + 4;
    const bool odd = r & 1;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i + 1; j < 3; ++j) {
            const double a0 = __shfl_sync(0xffffffffu, c[i][j][0], src0), a1 = __shfl_sync(0xffffffffu, c[i][j][1], src0);
            const double b0 = __shfl_sync(0xffffffffu, c[i][j][0], src1), b1 = __shfl_sync(0xffffffffu, c[i][j][1], src1);
            c[j][i][0] = odd ? a1 : a0;
            c[j][i][1] = odd ? b1 : b0;
        }
}

// local node a of a BoxGen Hexa8: offsets (dx,dy,dz) (generators/boxgen.py:172-185)
__device__ __forceinline__ int ndx(int a) { return (a >> 1) & 1; }
__device__ __forceinline__ int ndy(int a) { return (a >> 2) & 1; }
__device__ __forceinline__ int ndz(int a) { return (a ^ (a >> 1)) & 1; }

// mma row/column r  <->  local node ROWPERM[r] = {0,1,3,2,4,5,7,6}[r]  (conflict-free accumulation, see AccLayout)
__device__ __forceinline__ int rowNode(int r) { return (0x67542310u >> (4 * r)) & 7; }

// ---------------------------------------------------------------------------------------------
// phase A: one lane = one Gauss point of one element
// ---------------------------------------------------------------------------------------------
// stg: the warp's staged nodal data [18 patch nodes][x,y,z,u0,u1,u2], already offset to this lane's element
// (patch node of local node a = stg + (9 dx + 3 dy + dz) * 6).
// STG selects the staged image: 0 = 3x3x2 patch [X][Y][Z] (strides 9,3,1), 1 = the element's own 8 nodes [a],
// 2 = strip of four elements along z, [X][Y][Z] with 2 x 2 x 5 nodes (strides 10,5,1; row-pipelined kernel).
template <int MC, bool TL, int STG = 0, bool HREC = false>
__device__ __forceinline__ void gaussPointCompact(double* rec, const double* stg, int gp, const MatParams& mp,
                                                  const double* __restrict__ state_ref, double* __restrict__ state_temp, int64_t cstride,
                                                  bool writeState, int* failFlag, long long* tsub = nullptr, int rowFlip16 = 0) {
#ifdef EWB_TIMING
    long long tq0 = clock64();
#define EWB_SUB(i) do { const long long tq1 = clock64(); if (tsub) tsub[i] += tq1 - tq0; tq0 = tq1; } while (0)
#else
#define EWB_SUB(i)
#endif
    using R = RecLayout<MC>;
    constexpr int NST = 12 + (MC != MC_LE ? 1 : 0);
    double st[13];
#ifdef EWB_EXP_NOSTATE
#pragma unroll
    for (int c = 0; c < NST; ++c) st[c] = 0.0;
#else
#pragma unroll
    for (int c = 0; c < NST; ++c) st[c] = state_ref[c * cstride];
#endif

    double xi, eta, zeta, w;
    Gauss<8>::get(gp, xi, eta, zeta, w);
    // J = sum_a dN_a (x) X_a ; D[i][r] = sum_a u_a[i] dN_a[r]  (local displacement gradient)
    double Jm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double D[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    forNodes<8>([&](auto ic) {
        constexpr int a = decltype(ic)::value;
        double d[3];
        shapeDeriv<8, a>(xi, eta, zeta, d);
        constexpr int so = STG == 1 ? a * 6
                           : (((a >> 1) & 1) * (STG == 2 ? 10 : 9) + ((a >> 2) & 1) * (STG == 2 ? 5 : 3) + ((a ^ (a >> 1)) & 1)) * 6;
        const double x0 = stg[so], x1 = stg[so + 1], x2 = stg[so + 2];
        const double u0 = stg[so + 3], u1 = stg[so + 4], u2 = stg[so + 5];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            Jm[r * 3 + 0] = fma(d[r], x0, Jm[r * 3 + 0]);
            Jm[r * 3 + 1] = fma(d[r], x1, Jm[r * 3 + 1]);
            Jm[r * 3 + 2] = fma(d[r], x2, Jm[r * 3 + 2]);
            D[0 + r] = fma(u0, d[r], D[0 + r]);
            D[3 + r] = fma(u1, d[r], D[3 + r]);
            D[6 + r] = fma(u2, d[r], D[6 + r]);
        }
    });
    EWB_SUB(0);
    const double detJ = det3(Jm);
    double iJ[9];
    inv3(Jm, detJ, iJ);
    const double wd = w * detJ;
    if constexpr (!HREC) {
#pragma unroll
        for (int i = 0; i < 9; ++i) rec[i] = iJ[i];
    }
    EWB_SUB(1);
    // H[i][c] = du_i/dx_c = sum_r D[i][r] iJ[c][r]
    double H[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) H[i * 3 + c] = D[i * 3] * iJ[c * 3] + D[i * 3 + 1] * iJ[c * 3 + 1] + D[i * 3 + 2] * iJ[c * 3 + 2];

    if constexpr (!TL) {
        // Voigt 11,22,33,12,13,23 with engineering shear (_B3D8)
        const double de[6] = {H[0], H[4], H[8], H[1] + H[3], H[2] + H[6], H[5] + H[7]};
        double sg[6] = {st[0], st[1], st[2], st[3], st[4], st[5]};
        if constexpr (MC == MC_LE && HREC) {
            static_assert(!HREC || (MC == MC_LE && !TL), "fragment-order records: linear elastic only");
            hookeAdd(mp, de, sg);
            if (!(wd > 0.0)) atomicOr(failFlag, 8);  // sqrt(w detJ): inverted element
            const double sq = sqrt(wd);
            const int ks = gp >> 2, q = gp & 3;
            // rowFlip16 = 16: the y bit of the row index is inverted (row-pipelined kernel, odd element rows); two bases keep every
            // store offset an immediate
            double* Hq = rec + ks * RecLayoutH::KS + q;
            double* HqLo = Hq + rowFlip16;  // rows 0..3 move up by four rows
            double* HqHi = Hq - rowFlip16;  // rows 4..7 move down
            forNodes<8>([&](auto ic) {
                constexpr int a = decltype(ic)::value;
                constexpr int r = (0x67542310u >> (4 * a)) & 7;  // rowNode is an involution: row of node a
                double d[3];
                shapeDeriv<8, a>(xi, eta, zeta, d);
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    (r < 4 ? HqLo : HqHi)[2 * c * RecLayoutH::KS + 4 * r] = sq * (iJ[c * 3] * d[0] + iJ[c * 3 + 1] * d[1] + iJ[c * 3 + 2] * d[2]);
            });
#pragma unroll
            for (int i = 0; i < 6; ++i) rec[RecLayoutH::OFF_S + 6 * gp + i] = -sq * sg[i];
        } else if constexpr (MC == MC_LE) {
            hookeAdd(mp, de, sg);
            rec[R::C_CO] = wd;
        } else {
            VMResult r;
            double kappa = st[12];
            vonMises(mp, de, sg, kappa, r);
            st[12] = kappa;
            if (r.failed) atomicOr(failFlag, 1);
            rec[R::C_CO] = wd * r.lam;
            rec[R::C_CO + 1] = wd * r.mu;
            rec[R::C_CO + 2] = -wd * r.a;
#pragma unroll
            for (int i = 0; i < 6; ++i) rec[R::C_X + i] = r.n[i];
        }
        if constexpr (!HREC) {
#pragma unroll
            for (int i = 0; i < 6; ++i) rec[R::C_S + i] = -wd * sg[i];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            st[i] = sg[i];
            st[6 + i] += de[i];
        }
    } else {
        double F[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) F[i] = H[i];
        F[0] += 1.0;
        F[4] += 1.0;
        F[8] += 1.0;
        const double Jf = det3(F);
        double iF[9];
        inv3(F, Jf, iF);
        NHResult r;
        neoHooke(mp, F, Jf, r);
        rec[R::C_CO] = wd * r.c0;
        rec[R::C_CO + 1] = wd * r.c1;
        rec[R::C_CO + 2] = wd * r.c2;
        rec[R::C_CO + 3] = wd * r.c4;
#pragma unroll
        for (int i = 0; i < 6; ++i) rec[R::C_S + i] = -wd * r.tau[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) rec[R::C_X + i] = iF[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) rec[R::C_X + 9 + i] = F[i];
        // Kirchhoff stress Voigt 11,22,33,12,23,13 ; Green-Lagrange strain with doubled shear (voigtnotation.py)
        st[0] = r.tau[0];
        st[1] = r.tau[1];
        st[2] = r.tau[2];
        st[3] = r.tau[3];
        st[4] = r.tau[5];
        st[5] = r.tau[4];
        st[6] = H[0] + 0.5 * (H[0] * H[0] + H[3] * H[3] + H[6] * H[6]);
        st[7] = H[4] + 0.5 * (H[1] * H[1] + H[4] * H[4] + H[7] * H[7]);
        st[8] = H[8] + 0.5 * (H[2] * H[2] + H[5] * H[5] + H[8] * H[8]);
        st[9] = H[1] + H[3] + H[0] * H[1] + H[3] * H[4] + H[6] * H[7];
        st[10] = H[5] + H[7] + H[1] * H[2] + H[4] * H[5] + H[7] * H[8];
        st[11] = H[2] + H[6] + H[0] * H[2] + H[3] * H[5] + H[6] * H[8];
        st[12] = r.energy;
    }
    EWB_SUB(2);
#ifdef EWB_EXP_NOSTATE
    if (writeState && st[0] == 123.456) {
#else
    if (writeState) {
#endif
#pragma unroll
        for (int c = 0; c < NST; ++c) state_temp[c * cstride] = st[c];
    }
    EWB_SUB(3);
}

// ---------------------------------------------------------------------------------------------
// phase B for ONE element by one warp.  Lane (row = lane>>2, q = lane&3), a = rowNode(row), b_t =
// rowNode(2q+t): returns K0 = K[a][b_0], K1 = K[a][b_1] and (in all 4 lanes of the row) the residual
// row Pr of node a.  dNl[ks][r]: lane-constant dN_a/d(eta,xi,zeta) at Gauss point 4ks+q.
// ---------------------------------------------------------------------------------------------
// raw tensor-pipe accumulators of one element in one lane (two node blocks: t = 0, 1)
template <int MC> struct TileAcc;
template <> struct TileAcc<MC_LE> { double c[3][3][2]; };
template <> struct TileAcc<MC_VM> { double c1[3][3][2], c2[3][3][2]; bool elastic; };  // elastic: all 8 Gauss points of the element stayed elastic
// NH: c1[i][i] holds the merged diagonal tile sum (c1 + c2) n_i n_i (K_ii needs only their sum), c2[i][i] is unused; d0 = sum_i c0 g_i g_i
template <> struct TileAcc<MC_NH> { double c1[3][3][2], c2[3][3][2], d0[2]; };

// SYM: tile-transpose identity (mirrorTiles): only the tiles i <= j go to the tensor pipe.  Pays off where the tensor warps are the
// critical role and the element needs many products (von Mises in the row-pipelined kernel: 54 -> 36 DMMA, 338 -> 360 Melem/s).
template <int MC, bool SYM = (EWB_TILE_SYMMETRY != 0)>
__device__ __forceinline__ void elementTiles(const double* T, int lane, const double (&dNl)[2][3], const MatParams& mp, bool wantK,
                                             TileAcc<MC>& acc, double (&Pr)[3]) {
    using R = RecLayout<MC>;
    const int q = lane & 3;
    double g[2][3];
    Pr[0] = Pr[1] = Pr[2] = 0.0;
    const double* rec0 = T + q * R::RS;
    const double* rec1 = T + (4 + q) * R::RS;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const double* rec = ks ? rec1 : rec0;
#pragma unroll
        for (int c = 0; c < 3; ++c) g[ks][c] = rec[c * 3] * dNl[ks][0] + rec[c * 3 + 1] * dNl[ks][1] + rec[c * 3 + 2] * dNl[ks][2];
    }
    if constexpr (MC == MC_LE) {
        auto& c = acc.c;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) c[i][j][0] = c[i][j][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const double* rec = ks ? rec1 : rec0;
            const double w = rec[R::C_CO];
            const double* S = rec + R::C_S;
            Pr[0] += S[0] * g[ks][0] + S[3] * g[ks][1] + S[4] * g[ks][2];
            Pr[1] += S[3] * g[ks][0] + S[1] * g[ks][1] + S[5] * g[ks][2];
            Pr[2] += S[4] * g[ks][0] + S[5] * g[ks][1] + S[2] * g[ks][2];
            if (wantK) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double ai = w * g[ks][i];
#pragma unroll
                    for (int j = SYM ? i : 0; j < 3; ++j) dmma(c[i][j], ai, g[ks][j]);
                }
            }
        }
        if (SYM && wantK) mirrorTiles(c, lane);
    } else if constexpr (MC == MC_VM) {
        auto& c1 = acc.c1;
        auto& c2 = acc.c2;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) c1[i][j][0] = c1[i][j][1] = c2[i][j][0] = c2[i][j][1] = 0.0;
        // An element whose 8 Gauss points all stayed elastic has the constant tangent (lam, mu, a = 0): one product
        // M' = sum (w detJ mu grad N_a) grad N_b^T and K_ab = (lam/mu) M' + M'^T + tr(M') I (18 instead of 54 DMMA).
        // The 32 lanes of the warp read all 8 records (q = 0..3, two k-steps), so the vote sees every Gauss point.
        acc.elastic = __all_sync(0xffffffffu, rec0[R::C_CO + 2] == 0.0 && rec1[R::C_CO + 2] == 0.0);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const double* rec = ks ? rec1 : rec0;
            const double cl = rec[R::C_CO], cm = rec[R::C_CO + 1], ca = rec[R::C_CO + 2];
            const double* S = rec + R::C_S;
            const double* n = rec + R::C_X;
            Pr[0] += S[0] * g[ks][0] + S[3] * g[ks][1] + S[4] * g[ks][2];
            Pr[1] += S[3] * g[ks][0] + S[1] * g[ks][1] + S[5] * g[ks][2];
            Pr[2] += S[4] * g[ks][0] + S[5] * g[ks][1] + S[2] * g[ks][2];
            if (wantK && acc.elastic) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double mi = cm * g[ks][i];
#pragma unroll
                    for (int j = SYM ? i : 0; j < 3; ++j) dmma(c2[i][j], mi, g[ks][j]);
                }
            } else if (wantK) {
                // p_a = B_a^T n = N g_a, N = tensor(n)  (n Voigt 11,22,33,12,13,23)
                double pv[3];
                pv[0] = n[0] * g[ks][0] + n[3] * g[ks][1] + n[4] * g[ks][2];
                pv[1] = n[3] * g[ks][0] + n[1] * g[ks][1] + n[5] * g[ks][2];
                pv[2] = n[4] * g[ks][0] + n[5] * g[ks][1] + n[2] * g[ks][2];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double li = cl * g[ks][i], mi = cm * g[ks][i], ri = ca * pv[i];
#pragma unroll
                    for (int j = SYM ? i : 0; j < 3; ++j) {
                        dmma(c1[i][j], li, g[ks][j]);
                        dmma(c1[i][j], ri, pv[j]);
                        dmma(c2[i][j], mi, g[ks][j]);
                    }
                }
            }
        }
        if (SYM && wantK) {
            mirrorTiles(c2, lane);
            if (!acc.elastic) mirrorTiles(c1, lane);  // warp uniform
        }
    } else {
        auto& c1 = acc.c1;
        auto& c2 = acc.c2;
        auto& d0 = acc.d0;
        d0[0] = d0[1] = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) c1[i][j][0] = c1[i][j][1] = c2[i][j][0] = c2[i][j][1] = 0.0;
        }
        const bool wb = (mp.kind == EWB_MAT_NEOHOOKE_WB);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const double* rec = ks ? rec1 : rec0;
            const double k0 = rec[R::C_CO], k1 = rec[R::C_CO + 1], k2 = rec[R::C_CO + 2], k4 = rec[R::C_CO + 3];
            const double* S = rec + R::C_S;
            const double* iF = rec + R::C_X;
            const double* F = rec + R::C_X + 9;
            double nv[3];  // n_a = F^-T grad N_a
#pragma unroll
            for (int m = 0; m < 3; ++m) nv[m] = g[ks][0] * iF[m] + g[ks][1] * iF[3 + m] + g[ks][2] * iF[6 + m];
            Pr[0] += S[0] * nv[0] + S[3] * nv[1] + S[4] * nv[2];
            Pr[1] += S[3] * nv[0] + S[1] * nv[1] + S[5] * nv[2];
            Pr[2] += S[4] * nv[0] + S[5] * nv[1] + S[2] * nv[2];
            if (wantK) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double a1 = k1 * nv[i], a2 = k2 * nv[i];
                    dmma(d0, k0 * g[ks][i], g[ks][i]);
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        if (i == j) {
                            dmma(c1[i][i], a1 + a2, nv[i]);
                        } else if (!SYM || j > i) {
                            dmma(c1[i][j], a1, nv[j]);
                            dmma(c2[i][j], a2, nv[j]);
                        }
                    }
                }
                if (wb) {  // W_b: + c4 (f_a n_b^T + n_a f_b^T), f = F grad N
                    double fv[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) fv[i] = F[i * 3] * g[ks][0] + F[i * 3 + 1] * g[ks][1] + F[i * 3 + 2] * g[ks][2];
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = SYM ? i : 0; j < 3; ++j) {
                            dmma(c1[i][j], k4 * fv[i], nv[j]);
                            dmma(c1[i][j], k4 * nv[i], fv[j]);
                        }
                }
            }
        }
        if (SYM && wantK) {
            mirrorTiles(c1, lane);
            mirrorTiles(c2, lane);
        }
    }
    // reduce the residual row over the 4 lanes (Gauss-point pairs) of the row
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        Pr[i] += __shfl_xor_sync(0xffffffffu, Pr[i], 1);
        Pr[i] += __shfl_xor_sync(0xffffffffu, Pr[i], 2);
    }
}

// phase B from fragment-order records (RecLayoutH): M tiles and the residual row of one linear-elastic element
__device__ __forceinline__ void elementTilesH(const double* T, int lane, bool wantK, TileAcc<MC_LE>& acc, double (&Pr)[3]) {
    using H = RecLayoutH;
    const int q = lane & 3;
    double h[2][3];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int c = 0; c < 3; ++c) h[ks][c] = T[(2 * c + ks) * H::KS + lane];
    auto& c = acc.c;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[i][j][0] = c[i][j][1] = 0.0;
    Pr[0] = Pr[1] = Pr[2] = 0.0;
#if EWB_RESIDUAL_DMMA
    // Residual row on the tensor pipe: P[a][i] = sum_{gp,j} h[a][(gp,j)] S'[(gp,j)][i] is an (8 x 24)(24 x 8) product, six m8n8k4 steps
    // (k = Gauss points 4 ks + q of component j) whose A fragments are the h registers already loaded.  B fragment of this lane:
    // column n = lane >> 2 (= residual component i, columns 3..7 are zero), k = q: the symmetric S'_ij of Gauss point 4 ks + q.
    // Replaces 6 broadcast 128-bit loads (4 wavefronts each), 18 DFMA, 12 shuffles and 6 adds by 6 one-wavefront loads and 6 DMMA.
    {
        const int n = lane >> 2;
        // Voigt index of S'_nj (11,22,33,12,13,23): n = 0: {0,3,4}, n = 1: {3,1,5}, n = 2: {4,5,2}
        const int v0 = n == 0 ? 0 : (n == 1 ? 3 : 4), v1 = n == 0 ? 3 : (n == 1 ? 1 : 5), v2 = n == 0 ? 4 : (n == 1 ? 5 : 2);
        double pacc[2] = {0.0, 0.0};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const double* S = T + H::OFF_S + 6 * (4 * ks + q);
            const double b0 = n < 3 ? S[v0] : 0.0, b1 = n < 3 ? S[v1] : 0.0, b2 = n < 3 ? S[v2] : 0.0;
            dmma(pacc, h[ks][0], b0);
            dmma(pacc, h[ks][1], b1);
            dmma(pacc, h[ks][2], b2);
        }
        // lane (row, q) holds columns 2q, 2q + 1: q = 0 -> P[a][0], P[a][1]; q = 1 -> P[a][2]; hand the third value to the q = 0 lane
        Pr[0] = pacc[0];
        Pr[1] = pacc[1];
        Pr[2] = __shfl_down_sync(0xffffffffu, pacc[0], 1);
    }
#endif
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#if !EWB_RESIDUAL_DMMA
        // S'[gp][6]: 48 bytes per Gauss point, 16-byte aligned (OFF_S and PER_EL are even): three 128-bit loads
        const double2* S2 = reinterpret_cast<const double2*>(T + H::OFF_S + 6 * (4 * ks + q));
        const double2 s01 = S2[0], s23 = S2[1], s45 = S2[2];
        const double S[6] = {s01.x, s01.y, s23.x, s23.y, s45.x, s45.y};
        Pr[0] += S[0] * h[ks][0] + S[3] * h[ks][1] + S[4] * h[ks][2];
        Pr[1] += S[3] * h[ks][0] + S[1] * h[ks][1] + S[5] * h[ks][2];
        Pr[2] += S[4] * h[ks][0] + S[5] * h[ks][1] + S[2] * h[ks][2];
#endif
        if (wantK) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = EWB_TILE_SYMMETRY ? i : 0; j < 3; ++j) dmma(c[i][j], h[ks][i], h[ks][j]);
        }
    }
    if (EWB_TILE_SYMMETRY && wantK) mirrorTiles(c, lane);
#if !EWB_RESIDUAL_DMMA
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        Pr[i] += __shfl_xor_sync(0xffffffffu, Pr[i], 1);
        Pr[i] += __shfl_xor_sync(0xffffffffu, Pr[i], 2);
    }
#endif
}

// tangent assembly of node block t (0/1) of the lane from the raw accumulators: Kt[i*3+j] = Ke[3a+i][3b_t+j]
template <int MC>
__device__ __forceinline__ void finishBlock(const TileAcc<MC>& acc, int t, const MatParams& mp, double (&Kt)[9]) {
    if constexpr (MC == MC_LE) {
        const auto& c = acc.c;
        const double lpm = mp.lambda + mp.G;
        const double tr = mp.G * (c[0][0][t] + c[1][1][t] + c[2][2][t]);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            Kt[i * 3 + i] = fma(lpm, c[i][i][t], tr);
#pragma unroll
            for (int j = i + 1; j < 3; ++j) {
                const double x = c[i][j][t], y = c[j][i][t];
                Kt[i * 3 + j] = fma(mp.lambda, x, mp.G * y);
                Kt[j * 3 + i] = fma(mp.lambda, y, mp.G * x);
            }
        }
    } else if constexpr (MC == MC_VM) {
        const double tr = acc.c2[0][0][t] + acc.c2[1][1][t] + acc.c2[2][2][t];
        if (acc.elastic) {
            const double lom = mp.lambda / mp.G;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) Kt[i * 3 + j] = fma(lom, acc.c2[i][j][t], acc.c2[j][i][t]) + (i == j ? tr : 0.0);
        } else {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) Kt[i * 3 + j] = acc.c1[i][j][t] + acc.c2[j][i][t] + (i == j ? tr : 0.0);
        }
    } else {
        const double tr = acc.d0[t];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Kt[i * 3 + j] = i == j ? acc.c1[i][i][t] + tr : acc.c1[i][j][t] + acc.c2[j][i][t];
    }
}

template <int MC>
__device__ __forceinline__ void elementBlocks(const double* T, int lane, const double (&dNl)[2][3], const MatParams& mp, bool wantK,
                                              double (&K0)[9], double (&K1)[9], double (&Pr)[3]) {
    TileAcc<MC> acc;
    elementTiles<MC>(T, lane, dNl, mp, wantK, acc, Pr);
    finishBlock<MC>(acc, 0, mp, K0);
    finishBlock<MC>(acc, 1, mp, K1);
}

// Accumulator geometry (doubles).  Per node column 81 doubles per segment, laid out [i][s9][j] so that a
// CSR sub-row (27 values) is contiguous.  The four segment bases sit at residues 0,4,12,8 (mod 16 doubles)
// and the mma rows are permuted (rowNode) — found by exhaustive search (tools/bank_search.py): every half-warp
// of the accumulation loads/stores touches 16 different 8-byte banks, in both ping-pong roles of the dx=0
// segments and for TZ = 5 and 7 (1.0 wavefront per half-warp; the naive layout needs 3.0).
__host__ __device__ constexpr int alignRes(int x, int r) { return x + ((r - x % 16) + 16) % 16; }
template <int TY, int TZ>
struct AccLayout {
    static constexpr int NCOL = TY * TZ;
    static constexpr int CS = 81;
    static constexpr int SEGSZ = NCOL * CS;
    static constexpr int OFF_0A = 0;
    static constexpr int OFF_0B = alignRes(OFF_0A + SEGSZ, 4);
    static constexpr int OFF_P = alignRes(OFF_0B + SEGSZ, 12);
    static constexpr int OFF_M = alignRes(OFF_P + SEGSZ, 8);
    static constexpr int ACC_END = OFF_M + SEGSZ;
    static constexpr int PF = ACC_END;               // [2][NCOL][6]
    static constexpr int INFO = PF + 12 * NCOL;      // int32: colPart[NCOL], colCycz[NCOL], laneOff[NCOL][32], done[32], flushed[32]
    static constexpr int INFO_INTS = NCOL * 34 + 64;
    static constexpr int TABLES = alignRes(INFO + (INFO_INTS + 1) / 2, 0);
};

// Ordering of shared-memory data against the shared-memory flags.  A warp's shared-memory accesses are performed in
// program order, so a compiler barrier is enough; __threadfence_block() (MEMBAR.CTA) would also wait for every global
// store in flight (state write-back, CSR flush) on every round.  -DEWB_STRICT_FENCE restores the full fence.
#ifdef EWB_STRICT_FENCE
#define EWB_SMEM_FENCE() __threadfence_block()
#else
#define EWB_SMEM_FENCE() asm volatile("" ::: "memory")
#endif

// Warp-level wait until flag[dep] >= target for up to 9 dependencies (lane i polls dependency i).
// Bounded: a logic error sets bit 2 of the status word instead of hanging the GPU.
__device__ __forceinline__ void waitFlags(const volatile int* flag, int dep, bool has, int target, int* failFlag, int spinNs) {
    int spins = 0;
    while (true) {
        const bool ok = !has || flag[dep] >= target;
        if (__all_sync(0xffffffffu, ok)) break;
        if (++spins > (1 << 22)) {
            atomicOr(failFlag, 4);
            break;
        }
        if (spinNs > 0) __nanosleep(spinNs);
    }
    EWB_SMEM_FENCE();
}

// Lean flush of one plane step for x-interior planes (1 <= ex, ex + 1 <= NX - 2, both planes owned by the chunk, no peer
// buffer involved): the finished segments of the warp's (up to) four node columns go to the CSR value array.  For an interior
// plane the 9 sub-row pieces of a node column (3 rows x {dx = -1, 0, +1}) start at  colPtr + m * (3 cy cz),  m = (dx + 1) + 3 i,
// so one 64-bit base per column and an arithmetic progression replace the per-piece index arithmetic of flushPlane (which
// stays in charge of the first / last planes of the box and of a chunk).  Plane ex: dx = 0 (seg0) and dx = +1 (segP) plus
// P, F; plane ex + 1: dx = -1 (segM).  Every segment value read is cleared.
template <int TY, int TZ>
__device__ __forceinline__ void flushStepInterior(double* smem, const SweepArgs& A, int ex, int pyq, int pzq, int lane, int seg0, int segP,
                                                  int segM, int pf, const int* colPart, const int* colCycz, const int* laneOff,
                                                  int64_t totYZ, int NY, int NZ, int y0, int z0) {
    constexpr int CS = 81;
    const int64_t planeBase = 9 * (int64_t)(3 * ex - 1) * totYZ;
    const int64_t nextPlane = 27 * totYZ;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        const int ly_ = 2 * pyq + (cc >> 1), lz_ = 2 * pzq + (cc & 1);
        if (ly_ >= TY || lz_ >= TZ) continue;  // warp uniform
        const int col = ly_ * TZ + lz_;
        const int lo = laneOff[col * 32 + lane];  // -1: lane >= 27, column outside the box, or neighbour (dy,dz) outside the box
        const int stepE = 3 * colCycz[col];
        const int cp = colPart[col];
        if (lane < 27) {
            double* s0 = smem + (seg0 + col * CS + lane);
            double* sP = smem + (segP + col * CS + lane);
            double* sM = smem + (segM + col * CS + lane);
            double v0[3], vP[3], vM[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                v0[i] = s0[27 * i];
                vP[i] = sP[27 * i];
                vM[i] = sM[27 * i];
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                s0[27 * i] = 0.0;
                sP[27 * i] = 0.0;
                sM[27 * i] = 0.0;
            }
            if (lo >= 0) {
                double* ptr = A.data + (planeBase + 27 * (int64_t)cp + lo);
                double* ptrN = ptr + nextPlane;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    ptrN[(3 * i) * stepE] = vM[i];
                    ptr[(3 * i + 1) * stepE] = v0[i];
                    ptr[(3 * i + 2) * stepE] = vP[i];
                }
            }
        }
    }
    // P, F of plane ex: lane = 3 cc + component
    if (lane < 12) {
        const int cc = lane / 3, c = lane - 3 * cc;
        const int ly_ = 2 * pyq + (cc >> 1), lz_ = 2 * pzq + (cc & 1);
        if (ly_ < TY && lz_ < TZ) {
            const int col = ly_ * TZ + lz_;
            double* pfp = smem + (pf + col * 6 + c);
            const double pv = pfp[0], fv = pfp[3];
            pfp[0] = 0.0;
            pfp[3] = 0.0;
            if (colCycz[col] != 0) {
                const int64_t dof = 3 * ((((int64_t)ex * NY + (y0 + ly_)) * NZ) + (z0 + lz_)) + c;
                A.P[dof] = pv;
                A.F[dof] = fv;
            }
        }
    }
}

// One warp per 2x2 element patch; number of warps == number of patches of the tile.
template <int MC, bool TL, int TY, int TZ>
__global__ void __launch_bounds__(((TY + 1) / 2) * ((TZ + 1) / 2) * 32, 1) sweepKernel(const SweepArgs A) {
    using R = RecLayout<MC>;
    using AL = AccLayout<TY, TZ>;
    constexpr int NPY = (TY + 1) / 2, NPZ = (TZ + 1) / 2, NW = NPY * NPZ;
    constexpr int NCOL = TY * TZ;
    constexpr int CS = AL::CS;
    constexpr int NT = NW * 32;
    static_assert((TY & 1) == 1 && (TZ & 1) == 1, "tile edge must be odd (2x2 element patches)");
    static_assert(NW <= 32, "at most 32 patches per tile");
    constexpr bool COMPUTE_FIRST = NW <= 12;  // >= 168 registers per thread available

    extern __shared__ double smem[];
    // Accumulator segments are addressed by OFFSETS (doubles) into `smem`, never by pointers that get swapped or selected:
    // a selected pointer loses its address space and the accumulation degrades to generic LD/ST with 64-bit address math.
    constexpr int seg0a = AL::OFF_0A;  // dx=0 segment, ping
    constexpr int seg0b = AL::OFF_0B;  // dx=0 segment, pong
    constexpr int segP = AL::OFF_P;    // lower plane, dx=+1
    constexpr int segM = AL::OFF_M;    // upper plane, dx=-1
    constexpr int pfA = AL::PF;        // [NCOL][6] P,F of the lower plane
    constexpr int pfB = pfA + NCOL * 6;  // upper plane
    int* colPart = reinterpret_cast<int*>(smem + AL::INFO);
    int* colCycz = colPart + NCOL;
    int* laneOff = colCycz + NCOL;                // [NCOL][32]
    volatile int* doneCnt = laneOff + NCOL * 32;  // [32] rounds completed per patch
    volatile int* flushedCnt = doneCnt + 32;      // [32] plane steps flushed per patch
    double* tables = smem + AL::TABLES;           // [NW][4][PER_EL]
    double* stageAll = tables + (size_t)NW * 4 * R::PER_EL;  // [NW][108] nodal x,u of the warp's 3x3x2 patch nodes

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

    for (int i = tid; i < AL::INFO; i += NT) smem[i] = 0.0;
    if (tid < 64) doneCnt[tid] = 0;
    // CSR row base of node (ix,iy,iz) in closed form: 9 * (sum of the degrees of all preceding nodes);
    // deg = cx*cy*cz with c = 2 on a face, 3 inside (== the plan's node adjacency, checked by the parity tests).
    // base = 9 * (pre(ix)*totY*totZ + cx * colPart),  colPart = pre(iy)*totZ + cy*pre(iz).
    auto pre = [](int i) { return i == 0 ? 0 : 3 * i - 1; };
    const int totY = 3 * NY - 2, totZ = 3 * NZ - 2;
    for (int t = tid; t < NCOL * 32; t += NT) {
        const int col = t >> 5, e = t & 31;
        const int ly = col / TZ, lz = col % TZ;
        const int iy = y0 + ly, iz = z0 + lz;
        const bool colValid = ly < ny && lz < nz;
        const int cy = (iy > 0) + 1 + (iy < NY - 1), cz = (iz > 0) + 1 + (iz < NZ - 1);
        if (e == 0) {
            colPart[col] = colValid ? pre(iy) * totZ + cy * pre(iz) : 0;
            colCycz[col] = colValid ? cy * cz : 0;
        }
        const int s9 = e / 3, j = e % 3, dy = s9 / 3 - 1, dz = s9 % 3 - 1;
        const bool ok = colValid && e < 27 && iy + dy >= 0 && iy + dy < NY && iz + dz >= 0 && iz + dz < NZ;
        laneOff[t] = ok ? 3 * ((dy + (iy > 0 ? 1 : 0)) * cz + dz + (iz > 0 ? 1 : 0)) + j : -1;
    }
    __syncthreads();  // the only CTA-wide barrier: from here on the warps are ordered by dataflow flags

    int lo0 = seg0a;
    int hi0 = seg0b;
    int pfLo = pfA;
    int pfHi = pfB;
    const int64_t cstride = (int64_t)A.nX * A.nY * A.nZ * 8;
    const double* __restrict__ uSrc = TL ? A.U : A.dU;
    const int64_t totYZ = (int64_t)totY * totZ;

    // this warp's patch and its dependencies
    const int p = warp, pyq = p / NPZ, pzq = p % NPZ;
    // (1) accumulation of round R waits for the 8 neighbouring patches to have finished round R-1
    int depN = 0; bool hasN = false;
    if (lane < 9) {
        const int qy = pyq + lane / 3 - 1, qz = pzq + lane % 3 - 1;
        hasN = lane != 4 && qy >= 0 && qy < NPY && qz >= 0 && qz < NPZ;
        depN = hasN ? qy * NPZ + qz : 0;
    }
    // (2) the first accumulation of a plane step waits for the owners of the columns it touches to have flushed the previous step
    int depF = 0; bool hasF = false;
    if (lane < 4) {
        const int qy = pyq - (lane >> 1), qz = pzq - (lane & 1);
        hasF = qy >= 0 && qz >= 0;
        depF = hasF ? qy * NPZ + qz : 0;
    }
    // (3) the flush of this warp's columns (2pyq+{0,1}, 2pzq+{0,1}) waits for the patches that touch them
    int depD = 0; bool hasD = false;
    if (lane < 4) {
        const int qy = pyq + (lane >> 1), qz = pzq + (lane & 1);
        hasD = qy < NPY && qz < NPZ;
        depD = hasD ? qy * NPZ + qz : 0;
    }

    // ---- lane constants of phase B ----
    const int bRow = lane >> 2, bq = lane & 3;
    const int na = rowNode(bRow);  // node of this lane's mma row
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
    // accumulation targets of the two blocks of this lane: segment selector and offset inside the tile image
    int accOff[2];
    bool accSame[2];  // block stays in the node plane of a (dx = 0 segment) or crosses to the other plane
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int nb = rowNode(2 * bq + t);
        const int ry = ndy(nb) - ndy(na), rz = ndz(nb) - ndz(na);
        accSame[t] = ndx(nb) == ndx(na);
        accOff[t] = (ndy(na) * TZ + ndz(na)) * CS + ((ry + 1) * 3 + rz + 1) * 3;
    }
    const bool aHi = ndx(na) != 0;

    // flush the finished segments (nullptr = not finished) of this warp's owned nodes of plane ix, clear them
    auto flushPlane = [&](int ix, int sM, int s0, int sP, int pf) {  // segment offsets, -1 = not finished
        const int cx = (ix > 0) + 1 + (ix < NX - 1);
        const bool toPeer = A.peerData != nullptr && ix == NX - 1;  // ghost plane -> upper neighbour's receive buffer
        double* xbase = toPeer ? A.peerData : A.data + 9 * (int64_t)pre(ix) * totYZ;
        const int rx0 = ix > 0 ? 1 : 0;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int ly_ = 2 * pyq + (cc >> 1), lz_ = 2 * pzq + (cc & 1);
            if (ly_ >= TY || lz_ >= TZ) continue;
            const int col = ly_ * TZ + lz_;
            const int cycz = colCycz[col];
            if (cycz == 0) continue;
            const int lo = laneOff[col * 32 + lane];
            double* rowBase = xbase + ((int64_t)(9 * cx) * colPart[col] + lo);
            const int rowStride = 3 * cx * cycz;
            if (A.wantK && lane < 27) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const int seg = d == 0 ? sM : (d == 1 ? s0 : sP);
                    if (seg < 0) continue;
                    const int dx = d - 1;
                    const bool ok = ix + dx >= 0 && ix + dx < NX && lo >= 0;
                    double* src = smem + (seg + col * CS + lane);
                    double* dst = rowBase + 3 * ((dx + rx0) * cycz);
                    const double v0 = src[0], v1 = src[27], v2 = src[54];
                    src[0] = 0.0;
                    src[27] = 0.0;
                    src[54] = 0.0;
                    if (ok) {
                        dst[0] = v0;
                        dst[rowStride] = v1;
                        dst[2 * rowStride] = v2;
                    }
                }
            }
            if (pf >= 0 && lane < 3) {
                double* pfp = smem + (pf + col * 6 + lane);
                const double pv = pfp[0], fv = pfp[3];
                pfp[0] = 0.0;
                pfp[3] = 0.0;
                const int ly = 2 * pyq + (cc >> 1), lz = 2 * pzq + (cc & 1);
                const int64_t dof = 3 * ((((int64_t)ix * NY + (y0 + ly)) * NZ) + (z0 + lz)) + lane;
                if (toPeer) {
                    const int64_t pd = dof - 3 * (int64_t)ix * NY * NZ;
                    A.peerP[pd] = pv;
                    A.peerF[pd] = fv;
                } else if (A.accumulatePF) {
                    A.P[dof] += pv;
                    A.F[dof] += fv;
                } else {
                    A.P[dof] = pv;
                    A.F[dof] = fv;
                }
            }
        }
    };

    double* wt = tables + (size_t)warp * 4 * R::PER_EL;
    // phase-A lane constants: element k of the patch, Gauss point
    const int ak = lane >> 3, agp = lane & 7;
    const int apy = 2 * pyq + (ak >> 1), apz = 2 * pzq + (ak & 1);
    const int aey = y0 - 1 + apy, aez = z0 - 1 + apz;
    const bool aValid = aey >= 0 && aey < A.nY && aez >= 0 && aez < A.nZ && apy <= ny && apz <= nz;

    // cooperative, asynchronous staging of the patch's nodal coordinates / displacements for plane step `exs`
    double* stage = stageAll + warp * 108;
    const unsigned stageAddr = (unsigned)__cvta_generic_to_shared(stage);
    int stNode = -1;  // lane < 18 stages patch node `lane` (X = lane/9, Y = (lane/3)%3, Z = lane%3): node offset at plane 0
    if (lane < 18) {
        const int X = lane / 9, Y = (lane / 3) % 3, Z = lane % 3;
        const int iy = y0 - 1 + 2 * pyq + Y, iz = z0 - 1 + 2 * pzq + Z;
        if (iy >= 0 && iy < NY && iz >= 0 && iz < NZ) stNode = 3 * ((X * NY + iy) * NZ + iz);
    }
    const int planeStride3 = 3 * NY * NZ;
    auto stageLoad = [&](int exs) {
        if (stNode >= 0) {
            const int64_t o = (int64_t)exs * planeStride3 + stNode;
            const unsigned dst = stageAddr + 48u * lane;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * c), "l"(A.coords + o + c) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 24u + 8u * c), "l"(uSrc + o + c) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // pull the Gauss-point state of plane step `exs` into L2 ahead of its use
    auto statePrefetch = [&](int exs) {
        if (aValid) {
            const int64_t e = ((int64_t)exs * A.nY + aey) * A.nZ + aez;
            const double* sp = A.stateRef + e * 8 + agp;
#pragma unroll
            for (int c = 0; c < 12 + (MC != MC_LE ? 1 : 0); ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + c * cstride));
        }
    };
    // phase A of plane step `exs` for the 4 elements of this warp's patch (records are warp-private)
#ifdef EWB_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tsubA[4] = {0, 0, 0, 0};
    const long long tstart = clock64();
#endif
    auto phaseA = [&](int exs) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (aValid) {
            const int64_t e = ((int64_t)exs * A.nY + aey) * A.nZ + aez;
            const int64_t off = e * 8 + agp;
            const bool writeState = exs >= xa && apy >= 1 && apz >= 1;
            gaussPointCompact<MC, TL>(wt + ak * R::PER_EL + agp * R::RS, stage + ((ak >> 1) * 3 + (ak & 1)) * 6, agp, A.mp, A.stateRef + off,
                                      A.stateTemp + off, cstride, writeState, A.failFlag
#ifdef EWB_TIMING
                                      , tsubA
#endif
            );
        }
        __syncwarp();
    };
    stageLoad(exBegin);
    phaseA(exBegin);

    // per-round loop invariants: element validity (warp uniform), ownership of this lane's row node, tile-image offset
    unsigned validMask = 0, ownMask = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int py = 2 * pyq + (k >> 1), pz = 2 * pzq + (k & 1);
        const int ey = y0 - 1 + py, ez = z0 - 1 + pz;
        if (ey >= 0 && ey < A.nY && ez >= 0 && ez < A.nZ && py <= ny && pz <= nz) validMask |= 1u << k;
        const int ly = py - 1 + ndy(na), lz = pz - 1 + ndz(na);
        if (ly >= 0 && ly < ny && lz >= 0 && lz < nz) ownMask |= 1u << k;
    }
    const int eOff0 = (2 * pyq - 1) * TZ + (2 * pzq - 1);

    int step = 0;
    for (int ex = exBegin; ex <= exEnd; ++ex, ++step) {
        const bool loOwned = ex >= xa, hiOwned = (ex + 1) < xb;
        if (ex < exEnd) {  // next plane's nodal data and state travel while this plane's stiffness blocks are computed
            stageLoad(ex + 1);
            statePrefetch(ex + 1);
        }
        // per-step accumulation bases of this lane (segments rotate every plane)
        int accBase[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) accBase[t] = (aHi ? (accSame[t] ? hi0 : segM) : (accSame[t] ? lo0 : segP)) + accOff[t];
        const int pfBase = (aHi ? pfHi : pfLo) + (ndy(na) * TZ + ndz(na)) * 6;
        const bool planeOwned = aHi ? hiOwned : loOwned;
        // ------------- phase B: 4 colour rounds, one element per round -------------
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int Rnd = 4 * step + k;
            const bool valid = (validMask >> k) & 1;
            double K0[9], K1[9], Pr[3];
            // ---- ordering: same-colour elements of different patches never share a node.  With registers to spare the
            // blocks are computed BEFORE the wait, so a fast warp's tensor work overlaps its neighbours' accumulation. ----
            EWB_TIC(tB);
            if (COMPUTE_FIRST && valid) elementBlocks<MC>(wt + k * R::PER_EL, lane, dNl, A.mp, A.wantK != 0, K0, K1, Pr);
            EWB_ACC(2, tB);
            EWB_TIC(tW);
            if (k == 0) waitFlags(flushedCnt, depF, hasF, step, A.failFlag, A.spinNs);
            waitFlags(doneCnt, depN, hasN, Rnd, A.failFlag, A.spinNs);
            EWB_ACC(1, tW);
            if (!COMPUTE_FIRST && valid) elementBlocks<MC>(wt + k * R::PER_EL, lane, dNl, A.mp, A.wantK != 0, K0, K1, Pr);
            EWB_TIC(tE);
            if (valid) {
                if (planeOwned && ((ownMask >> k) & 1)) {
                    const int eOff = eOff0 + (k >> 1) * TZ + (k & 1);
                    if (bq == 0) {
                        double* pf = smem + (pfBase + eOff * 6);
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            pf[i] += Pr[i];
                            pf[3 + i] += fabs(Pr[i]);
                        }
                    }
                    if (A.wantK) {
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            double* dst = smem + (accBase[t] + eOff * CS);
                            const double* Kt = t ? K1 : K0;
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int j = 0; j < 3; ++j) dst[i * 27 + j] += Kt[i * 3 + j];
                        }
                    }
                }
            }
            EWB_SMEM_FENCE();
            __syncwarp();
            if (lane == 0) doneCnt[p] = Rnd + 1;
            EWB_ACC(3, tE);
        }
        // ------------- phase A of the NEXT plane: this warp's records are dead, nobody else reads them -------------
        EWB_TIC(tA);
        if (ex < exEnd) phaseA(ex + 1);
        EWB_ACC(0, tA);
        // ---------------- flush this warp's finished columns, rotate ----------------
        EWB_TIC(tWF);
        waitFlags(doneCnt, depD, hasD, 4 * (step + 1), A.failFlag, A.spinNs);
        EWB_ACC(4, tWF);
        EWB_TIC(tF);
        if (loOwned && hiOwned && ex >= 1 && ex + 1 <= NX - 2 && A.wantK && !A.accumulatePF) {
            flushStepInterior<TY, TZ>(smem, A, ex, pyq, pzq, lane, lo0, segP, segM, pfLo, colPart, colCycz, laneOff, totYZ, NY, NZ, y0, z0);
        } else {
            if (loOwned) flushPlane(ex, -1, lo0, segP, pfLo);
            if (hiOwned) flushPlane(ex + 1, segM, -1, -1, -1);
            if (ex == exEnd && xb == NX) flushPlane(NX - 1, -1, hi0, -1, pfHi);  // last node plane: nothing above it
        }
        EWB_SMEM_FENCE();
        __syncwarp();
        if (lane == 0) flushedCnt[p] = step + 1;
        EWB_ACC(5, tF);
        {
            const int t0 = lo0; lo0 = hi0; hi0 = t0;
            const int t1 = pfLo; pfLo = pfHi; pfHi = t1;
        }
    }
#ifdef EWB_TIMING
    if (A.timing != nullptr && lane == 0) {
        tacc[6] = clock64() - tstart;
        tacc[7] = step;
        for (int i = 0; i < 8; ++i) A.timing[((size_t)blockIdx.x * NW + warp) * 12 + i] = tacc[i];
        for (int i = 0; i < 4; ++i) A.timing[((size_t)blockIdx.x * NW + warp) * 12 + 8 + i] = tsubA[i];
    }
#endif
}

// Producer/consumer variant: NP consumer warps (one per 2x2 element patch: tensor-pipe stiffness blocks, shared-memory
// accumulation, flush) + NWP producer warps that run phase A (kinematics, constitutive update, state write-back) one
// plane step ahead into double-buffered Gauss-point records.  The latency-bound phase A thereby overlaps the FP64- and
// shared-memory-bound consumer work instead of preceding it.
template <int MC, bool TL, int TY, int TZ, int NWP>
__global__ void __launch_bounds__((((TY + 1) / 2) * ((TZ + 1) / 2) + NWP) * 32, ((((TY + 1) / 2) * ((TZ + 1) / 2) + NWP) <= 8 ? 2 : 1))
    sweepKernelPC(const SweepArgs A) {
    using R = RecLayout<MC>;
    using AL = AccLayout<TY, TZ>;
    constexpr int NPY = (TY + 1) / 2, NPZ = (TZ + 1) / 2, NW = NPY * NPZ;
    constexpr int NCOL = TY * TZ;
    constexpr int CS = AL::CS;
    constexpr int NT = (NW + NWP) * 32;
    static_assert((TY & 1) == 1 && (TZ & 1) == 1, "tile edge must be odd (2x2 element patches)");
    static_assert(NW <= 32, "at most 32 patches per tile");
    constexpr bool COMPUTE_FIRST = (NW + NWP) <= 12;
    constexpr bool HREC = (MC == MC_LE) && !TL;  // fragment-order records (RecLayoutH)
    constexpr int PEL = HREC ? RecLayoutH::PER_EL : R::PER_EL;

    extern __shared__ double smem[];
    // Accumulator segments are addressed by OFFSETS (doubles) into `smem`, never by pointers that get swapped or selected:
    // a selected pointer loses its address space and the accumulation degrades to generic LD/ST with 64-bit address math.
    constexpr int seg0a = AL::OFF_0A;  // dx=0 segment, ping
    constexpr int seg0b = AL::OFF_0B;  // dx=0 segment, pong
    constexpr int segP = AL::OFF_P;    // lower plane, dx=+1
    constexpr int segM = AL::OFF_M;    // upper plane, dx=-1
    constexpr int pfA = AL::PF;        // [NCOL][6] P,F of the lower plane
    constexpr int pfB = pfA + NCOL * 6;  // upper plane
    int* colPart = reinterpret_cast<int*>(smem + AL::INFO);
    int* colCycz = colPart + NCOL;
    int* laneOff = colCycz + NCOL;                // [NCOL][32]
    volatile int* doneCnt = laneOff + NCOL * 32;  // [32] rounds completed per patch
    volatile int* flushedCnt = doneCnt + 32;      // [32] plane steps flushed per patch
    double* tables = smem + AL::TABLES;           // [2][NW][4][PER_EL] double-buffered Gauss-point records
    double* stageAll = tables + (size_t)2 * NW * 4 * PEL;  // [NWP][2][108] nodal x,u of a patch's 3x3x2 nodes
    volatile int* producedCnt = reinterpret_cast<volatile int*>(stageAll + NWP * 216);  // [32] plane steps whose records are ready
    volatile int* consumedCnt = producedCnt + 32;                                      // [32] plane steps whose records are consumed

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

    for (int i = tid; i < AL::INFO; i += NT) smem[i] = 0.0;
    if (tid < 64) { doneCnt[tid] = 0; producedCnt[tid] = 0; }
    // CSR row base of node (ix,iy,iz) in closed form: 9 * (sum of the degrees of all preceding nodes);
    // deg = cx*cy*cz with c = 2 on a face, 3 inside (== the plan's node adjacency, checked by the parity tests).
    // base = 9 * (pre(ix)*totY*totZ + cx * colPart),  colPart = pre(iy)*totZ + cy*pre(iz).
    auto pre = [](int i) { return i == 0 ? 0 : 3 * i - 1; };
    const int totY = 3 * NY - 2, totZ = 3 * NZ - 2;
    for (int t = tid; t < NCOL * 32; t += NT) {
        const int col = t >> 5, e = t & 31;
        const int ly = col / TZ, lz = col % TZ;
        const int iy = y0 + ly, iz = z0 + lz;
        const bool colValid = ly < ny && lz < nz;
        const int cy = (iy > 0) + 1 + (iy < NY - 1), cz = (iz > 0) + 1 + (iz < NZ - 1);
        if (e == 0) {
            colPart[col] = colValid ? pre(iy) * totZ + cy * pre(iz) : 0;
            colCycz[col] = colValid ? cy * cz : 0;
        }
        const int s9 = e / 3, j = e % 3, dy = s9 / 3 - 1, dz = s9 % 3 - 1;
        const bool ok = colValid && e < 27 && iy + dy >= 0 && iy + dy < NY && iz + dz >= 0 && iz + dz < NZ;
        laneOff[t] = ok ? 3 * ((dy + (iy > 0 ? 1 : 0)) * cz + dz + (iz > 0 ? 1 : 0)) + j : -1;
    }
    __syncthreads();  // the only CTA-wide barrier: from here on the warps are ordered by dataflow flags

    int lo0 = seg0a;
    int hi0 = seg0b;
    int pfLo = pfA;
    int pfHi = pfB;
    const int64_t cstride = (int64_t)A.nX * A.nY * A.nZ * 8;
    const double* __restrict__ uSrc = TL ? A.U : A.dU;
    const int64_t totYZ = (int64_t)totY * totZ;

    if (warp >= NW) {
        // ===================== producer warps: phase A, one plane step ahead =====================
        const int j = warp - NW;
        double* stageBuf = stageAll + j * 216;  // two staging buffers: the next task's nodal data travels during the current one
        const unsigned stageAddr0 = (unsigned)__cvta_generic_to_shared(stageBuf);
        const int planeStride3 = 3 * NY * NZ;
        const int ak = lane >> 3, agp = lane & 7;
        constexpr int TPS = (NW + NWP - 1) / NWP;  // tasks (patches) per plane step of one producer
        const int nSteps = exEnd - exBegin + 1;
        const int nTasks = nSteps * TPS;
        // issue the asynchronous copies (nodal x,u) and the L2 prefetch (Gauss-point state) of task t
        auto issue = [&](int t) {
            const int st = t / TPS, p = j + (t % TPS) * NWP;
            if (t < nTasks && p < NW) {
                const int exs = exBegin + st;
                const int pyq = p / NPZ, pzq = p % NPZ;
                if (lane < 18) {
                    const int X = lane / 9, Y = (lane / 3) % 3, Z = lane % 3;
                    const int iy = y0 - 1 + 2 * pyq + Y, iz = z0 - 1 + 2 * pzq + Z;
                    if (iy >= 0 && iy < NY && iz >= 0 && iz < NZ) {
                        const int64_t o = (int64_t)exs * planeStride3 + 3 * ((X * NY + iy) * NZ + iz);
                        const unsigned dst = stageAddr0 + (unsigned)(t & 1) * 864u + 48u * lane;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * c), "l"(A.coords + o + c) : "memory");
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 24u + 8u * c), "l"(uSrc + o + c) : "memory");
                        }
                    }
                }
                const int apy = 2 * pyq + (ak >> 1), apz = 2 * pzq + (ak & 1);
                const int aey = y0 - 1 + apy, aez = z0 - 1 + apz;
                if (aey >= 0 && aey < A.nY && aez >= 0 && aez < A.nZ && apy <= ny && apz <= nz) {
                    const double* sp = A.stateRef + (((int64_t)exs * A.nY + aey) * A.nZ + aez) * 8 + agp;
#pragma unroll
                    for (int c = 0; c < 12 + (MC != MC_LE ? 1 : 0); ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + c * cstride));
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        issue(0);
#pragma unroll 1
        for (int t = 0; t < nTasks; ++t) {
            issue(t + 1);
            const int step = t / TPS, p = j + (t % TPS) * NWP;
            if (p >= NW) {
                asm volatile("cp.async.wait_group 1;" ::: "memory");
                continue;
            }
            const int ex = exBegin + step;
            const int pyq = p / NPZ, pzq = p % NPZ;
            // the record buffer (step & 1) of this patch must have been consumed (it last held plane step - 2)
            waitFlags(consumedCnt, p, lane == 0, step - 1, A.failFlag, A.spinNs);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
            const double* stage = stageBuf + (t & 1) * 108;
            const int apy = 2 * pyq + (ak >> 1), apz = 2 * pzq + (ak & 1);
            const int aey = y0 - 1 + apy, aez = z0 - 1 + apz;
            const bool aValid = aey >= 0 && aey < A.nY && aez >= 0 && aez < A.nZ && apy <= ny && apz <= nz;
            if (aValid) {
                double* rec = tables + ((size_t)((step & 1) * NW + p) * 4 + ak) * PEL + (HREC ? 0 : agp * R::RS);
                const int64_t e = ((int64_t)ex * A.nY + aey) * A.nZ + aez;
                const int64_t off = e * 8 + agp;
                const bool writeState = ex >= xa && apy >= 1 && apz >= 1;
                gaussPointCompact<MC, TL, 0, HREC>(rec, stage + ((ak >> 1) * 3 + (ak & 1)) * 6, agp, A.mp, A.stateRef + off, A.stateTemp + off,
                                                       cstride, writeState, A.failFlag);
            }
            EWB_SMEM_FENCE();
            __syncwarp();
            if (lane == 0) producedCnt[p] = step + 1;
        }
        return;
    }
    // ===================== consumer warps: one per patch =====================
    // this warp's patch and its dependencies
    const int p = warp, pyq = p / NPZ, pzq = p % NPZ;
    // (1) accumulation of round R waits for the 8 neighbouring patches to have finished round R-1
    int depN = 0; bool hasN = false;
    if (lane < 9) {
        const int qy = pyq + lane / 3 - 1, qz = pzq + lane % 3 - 1;
        hasN = lane != 4 && qy >= 0 && qy < NPY && qz >= 0 && qz < NPZ;
        depN = hasN ? qy * NPZ + qz : 0;
    }
    // (2) the first accumulation of a plane step waits for the owners of the columns it touches to have flushed the previous step
    int depF = 0; bool hasF = false;
    if (lane < 4) {
        const int qy = pyq - (lane >> 1), qz = pzq - (lane & 1);
        hasF = qy >= 0 && qz >= 0;
        depF = hasF ? qy * NPZ + qz : 0;
    }
    // (3) the flush of this warp's columns (2pyq+{0,1}, 2pzq+{0,1}) waits for the patches that touch them
    int depD = 0; bool hasD = false;
    if (lane < 4) {
        const int qy = pyq + (lane >> 1), qz = pzq + (lane & 1);
        hasD = qy < NPY && qz < NPZ;
        depD = hasD ? qy * NPZ + qz : 0;
    }

    // ---- lane constants of phase B ----
    const int bRow = lane >> 2, bq = lane & 3;
    const int na = rowNode(bRow);  // node of this lane's mma row
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
    // accumulation targets of the two blocks of this lane: segment selector and offset inside the tile image
    int accOff[2];
    bool accSame[2];  // block stays in the node plane of a (dx = 0 segment) or crosses to the other plane
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int nb = rowNode(2 * bq + t);
        const int ry = ndy(nb) - ndy(na), rz = ndz(nb) - ndz(na);
        accSame[t] = ndx(nb) == ndx(na);
        accOff[t] = (ndy(na) * TZ + ndz(na)) * CS + ((ry + 1) * 3 + rz + 1) * 3;
    }
    const bool aHi = ndx(na) != 0;

    // flush the finished segments (nullptr = not finished) of this warp's owned nodes of plane ix, clear them
    auto flushPlane = [&](int ix, int sM, int s0, int sP, int pf) {  // segment offsets, -1 = not finished
        const int cx = (ix > 0) + 1 + (ix < NX - 1);
        const bool toPeer = A.peerData != nullptr && ix == NX - 1;  // ghost plane -> upper neighbour's receive buffer
        double* xbase = toPeer ? A.peerData : A.data + 9 * (int64_t)pre(ix) * totYZ;
        const int rx0 = ix > 0 ? 1 : 0;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int ly_ = 2 * pyq + (cc >> 1), lz_ = 2 * pzq + (cc & 1);
            if (ly_ >= TY || lz_ >= TZ) continue;
            const int col = ly_ * TZ + lz_;
            const int cycz = colCycz[col];
            if (cycz == 0) continue;
            const int lo = laneOff[col * 32 + lane];
            double* rowBase = xbase + ((int64_t)(9 * cx) * colPart[col] + lo);
            const int rowStride = 3 * cx * cycz;
            if (A.wantK && lane < 27) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const int seg = d == 0 ? sM : (d == 1 ? s0 : sP);
                    if (seg < 0) continue;
                    const int dx = d - 1;
                    const bool ok = ix + dx >= 0 && ix + dx < NX && lo >= 0;
                    double* src = smem + (seg + col * CS + lane);
                    double* dst = rowBase + 3 * ((dx + rx0) * cycz);
                    const double v0 = src[0], v1 = src[27], v2 = src[54];
                    src[0] = 0.0;
                    src[27] = 0.0;
                    src[54] = 0.0;
                    if (ok) {
                        dst[0] = v0;
                        dst[rowStride] = v1;
                        dst[2 * rowStride] = v2;
                    }
                }
            }
            if (pf >= 0 && lane < 3) {
                double* pfp = smem + (pf + col * 6 + lane);
                const double pv = pfp[0], fv = pfp[3];
                pfp[0] = 0.0;
                pfp[3] = 0.0;
                const int ly = 2 * pyq + (cc >> 1), lz = 2 * pzq + (cc & 1);
                const int64_t dof = 3 * ((((int64_t)ix * NY + (y0 + ly)) * NZ) + (z0 + lz)) + lane;
                if (toPeer) {
                    const int64_t pd = dof - 3 * (int64_t)ix * NY * NZ;
                    A.peerP[pd] = pv;
                    A.peerF[pd] = fv;
                } else if (A.accumulatePF) {
                    A.P[dof] += pv;
                    A.F[dof] += fv;
                } else {
                    A.P[dof] = pv;
                    A.F[dof] = fv;
                }
            }
        }
    };

#ifdef EWB_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tsubA[4] = {0, 0, 0, 0};
    const long long tstart = clock64();
#endif
    // per-round loop invariants: element validity (warp uniform), ownership of this lane's row node, tile-image offset
    unsigned validMask = 0, ownMask = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int py = 2 * pyq + (k >> 1), pz = 2 * pzq + (k & 1);
        const int ey = y0 - 1 + py, ez = z0 - 1 + pz;
        if (ey >= 0 && ey < A.nY && ez >= 0 && ez < A.nZ && py <= ny && pz <= nz) validMask |= 1u << k;
        const int ly = py - 1 + ndy(na), lz = pz - 1 + ndz(na);
        if (ly >= 0 && ly < ny && lz >= 0 && lz < nz) ownMask |= 1u << k;
    }
    const int eOff0 = (2 * pyq - 1) * TZ + (2 * pzq - 1);

    int step = 0;
    for (int ex = exBegin; ex <= exEnd; ++ex, ++step) {
        const bool loOwned = ex >= xa, hiOwned = (ex + 1) < xb;
        // the producers publish this plane step's Gauss-point records
        EWB_TIC(tWP);
        waitFlags(producedCnt, p, lane == 0, step + 1, A.failFlag, A.spinNs);
        EWB_ACC(0, tWP);
        const double* wt = tables + (size_t)((step & 1) * NW + p) * 4 * PEL;
        // per-step accumulation bases of this lane (segments rotate every plane)
        int accBase[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) accBase[t] = (aHi ? (accSame[t] ? hi0 : segM) : (accSame[t] ? lo0 : segP)) + accOff[t];
        const int pfBase = (aHi ? pfHi : pfLo) + (ndy(na) * TZ + ndz(na)) * 6;
        const bool planeOwned = aHi ? hiOwned : loOwned;
        // ------------- phase B: 4 colour rounds, one element per round -------------
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int Rnd = 4 * step + k;
            const bool valid = (validMask >> k) & 1;
            double K0[9], K1[9], Pr[3];
            // ---- ordering: same-colour elements of different patches never share a node.  With registers to spare the
            // blocks are computed BEFORE the wait, so a fast warp's tensor work overlaps its neighbours' accumulation. ----
            EWB_TIC(tB);
            if constexpr (HREC) {
                if (COMPUTE_FIRST && valid) {
                    TileAcc<MC_LE> acc;
                    elementTilesH(wt + k * PEL, lane, A.wantK != 0, acc, Pr);
                    finishBlock<MC_LE>(acc, 0, A.mp, K0);
                    finishBlock<MC_LE>(acc, 1, A.mp, K1);
                }
            } else {
                if (COMPUTE_FIRST && valid) elementBlocks<MC>(wt + k * PEL, lane, dNl, A.mp, A.wantK != 0, K0, K1, Pr);
            }
            if (COMPUTE_FIRST && k == 3) {  // all four elements' records are consumed: the producers may refill this buffer
                __syncwarp();
                if (lane == 0) consumedCnt[p] = step + 1;
            }
            EWB_ACC(2, tB);
            EWB_TIC(tW);
            if (k == 0) waitFlags(flushedCnt, depF, hasF, step, A.failFlag, A.spinNs);
            waitFlags(doneCnt, depN, hasN, Rnd, A.failFlag, A.spinNs);
            EWB_ACC(1, tW);
            EWB_TIC(tB2);
            TileAcc<MC> tacc_;
            if constexpr (HREC) {
                if (!COMPUTE_FIRST && valid) elementTilesH(wt + k * PEL, lane, A.wantK != 0, tacc_, Pr);
            } else {
                if (!COMPUTE_FIRST && valid) elementTiles<MC>(wt + k * PEL, lane, dNl, A.mp, A.wantK != 0, tacc_, Pr);
            }
            if (!COMPUTE_FIRST && k == 3) {  // all four elements' records are consumed: the producers may refill this buffer
                __syncwarp();
                if (lane == 0) consumedCnt[p] = step + 1;
            }
            EWB_ACC(2, tB2);
            EWB_TIC(tE);
            if (valid) {
                if (planeOwned && ((ownMask >> k) & 1)) {
                    const int eOff = eOff0 + (k >> 1) * TZ + (k & 1);
                    if (bq == 0) {
                        double* pf = smem + (pfBase + eOff * 6);
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            pf[i] += Pr[i];
                            pf[3 + i] += fabs(Pr[i]);
                        }
                    }
                    if (A.wantK) {
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            double* dst = smem + (accBase[t] + eOff * CS);
                            double Kf[9];
                            if (!COMPUTE_FIRST) finishBlock<MC>(tacc_, t, A.mp, Kf);
                            const double* Kt = COMPUTE_FIRST ? (t ? K1 : K0) : Kf;
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int j = 0; j < 3; ++j) dst[i * 27 + j] += Kt[i * 3 + j];
                        }
                    }
                }
            }
            EWB_SMEM_FENCE();
            __syncwarp();
            if (lane == 0) doneCnt[p] = Rnd + 1;
            EWB_ACC(3, tE);
        }
        // ---------------- flush this warp's finished columns, rotate ----------------
        EWB_TIC(tWF);
        waitFlags(doneCnt, depD, hasD, 4 * (step + 1), A.failFlag, A.spinNs);
        EWB_ACC(4, tWF);
        EWB_TIC(tF);
        if (loOwned && hiOwned && ex >= 1 && ex + 1 <= NX - 2 && A.wantK && !A.accumulatePF) {
            flushStepInterior<TY, TZ>(smem, A, ex, pyq, pzq, lane, lo0, segP, segM, pfLo, colPart, colCycz, laneOff, totYZ, NY, NZ, y0, z0);
        } else {
            if (loOwned) flushPlane(ex, -1, lo0, segP, pfLo);
            if (hiOwned) flushPlane(ex + 1, segM, -1, -1, -1);
            if (ex == exEnd && xb == NX) flushPlane(NX - 1, -1, hi0, -1, pfHi);  // last node plane: nothing above it
        }
        EWB_SMEM_FENCE();
        __syncwarp();
        if (lane == 0) flushedCnt[p] = step + 1;
        EWB_ACC(5, tF);
        {
            const int t0 = lo0; lo0 = hi0; hi0 = t0;
            const int t1 = pfLo; pfLo = pfHi; pfHi = t1;
        }
    }
#ifdef EWB_TIMING
    if (A.timing != nullptr && lane == 0) {
        tacc[6] = clock64() - tstart;
        tacc[7] = step;
        for (int i = 0; i < 8; ++i) A.timing[((size_t)blockIdx.x * NW + warp) * 12 + i] = tacc[i];
        for (int i = 0; i < 4; ++i) A.timing[((size_t)blockIdx.x * NW + warp) * 12 + 8 + i] = tsubA[i];
    }
#endif
}

struct SweepPlan {
    int64_t nX = 0, nY = 0, nZ = 0;
    int nSM = 148;
    int chunkOverride = 0;  // EWB_CHUNKS (tuning knob, read once at plan creation)
    // row-pipelined kernel, set around a launch by ewb_plan_x_chunks / ewb_assemble_chunks: x-chunk range of the launch
    // (chunkEnd < 0: all chunks) and, when tilingOut is set, "report the tiling instead of launching"
    int chunkBegin = 0, chunkEnd = -1;
    int* tilingOut = nullptr;  // [2]: chunkLen, nChunks
    // x-chunk count of the chunk-by-chunk launches (EWB_PIPE_CHUNKS; 0: the tiling of the single launch; -1: automatic — six chunks
    // for boxes of at least 48 node planes: the first upload and the last download, which nothing overlaps, shrink to a sixth
    // of the vector, 496 -> 531 Melem/s end to end at 100^3; results do not depend on the chunking, bit for bit)
    int pipeChunks = -1;
    int spinNs = 0;         // EWB_SPIN_NS
    long long* timingBuf = nullptr;
    size_t timingCount = 0;

    int build(int64_t nx, int64_t ny, int64_t nz) {
        nX = nx; nY = ny; nZ = nz;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, dev);
        if (nSM <= 0) nSM = 148;
        if (const char* ev = getenv("EWB_CHUNKS")) chunkOverride = std::max(0, atoi(ev));
        if (const char* ev = getenv("EWB_SPIN_NS")) spinNs = atoi(ev);
        if (const char* ev = getenv("EWB_PIPE_CHUNKS")) pipeChunks = std::max(-1, atoi(ev));
        return 0;
    }
    void release() {
        if (timingBuf) cudaFree(timingBuf);
        timingBuf = nullptr;
    }
    double* peerData = nullptr;  // set by ewb_plan_set_peer
    double* peerP = nullptr;
    double* peerF = nullptr;

    void fillCommon(SweepArgs& a, const MatParams& mp, const ewb_buffers* b, int* failFlag, int flags) const {
        a.nX = (int)nX; a.nY = (int)nY; a.nZ = (int)nZ;
        a.spinNs = spinNs;
        a.timing = nullptr;
        a.tileRows = 0;
        a.chunkBase = 0;
        a.coords = b->coords; a.U = b->U; a.dU = b->dU; a.stateRef = b->state_ref; a.stateTemp = b->state_temp;
        a.data = b->csr_data; a.P = b->P; a.F = b->F; a.mp = mp; a.failFlag = failFlag;
        a.wantK = (flags & EWB_FLAG_NO_STIFFNESS) ? 0 : 1;
        a.accumulatePF = (flags & EWB_FLAG_ACCUMULATE_PF) ? 1 : 0;
        a.peerData = peerData; a.peerP = peerP; a.peerF = peerF;
    }

    template <int TY, int TZ>
    int fillArgs(SweepArgs& a, const MatParams& mp, const ewb_buffers* b, int* failFlag, int flags, cudaStream_t st, int NW_) {
        (void)st; (void)NW_;
        fillCommon(a, mp, b, failFlag, flags);
        const int64_t tiles = (int64_t)((nY + 1 + TY - 1) / TY) * ((nZ + 1 + TZ - 1) / TZ);
        a.tilesY = (int)((nY + 1 + TY - 1) / TY);
        a.tilesZ = (int)((nZ + 1 + TZ - 1) / TZ);
        // chunks along x: minimise (number of CTA rounds on nSM SMs) x (planes per chunk incl. the halo plane)
        int best = 1;
        double bestCost = 1e300;
        for (int c = 1; c <= 64 && (nX + 1) / c >= 6; ++c) {
            const int64_t len = (nX + 1 + c - 1) / c, nc = (nX + 1 + len - 1) / len;
            const double rounds = (double)((tiles * nc + nSM - 1) / nSM);
            const double cost = rounds * (double)(len + 1);
            if (cost < bestCost) { bestCost = cost; best = c; }
        }
        if (chunkOverride > 0) best = chunkOverride;
#ifdef EWB_TIMING
        {
            const size_t nT = (size_t)tiles * 64 * NW_ * 12;
            if (timingCount < nT) {
                if (timingBuf) cudaFree(timingBuf);
                cudaMalloc((void**)&timingBuf, nT * sizeof(long long));
                timingCount = nT;
            }
            cudaMemsetAsync(timingBuf, 0, nT * sizeof(long long), st);
            a.timing = timingBuf;
        }
#endif
        a.chunkLen = (int)((nX + 1 + best - 1) / best);
        a.nChunks = (int)((nX + 1 + a.chunkLen - 1) / a.chunkLen);
        return EWB_OK;
    }

    bool indexable() const { return (nX + 1) * (nY + 1) * (nZ + 1) < ((int64_t)1 << 31) / 3; }  // int32 node indexing inside the kernels

    template <int MC, bool TL, int TY, int TZ>
    int launchT(const MatParams& mp, const ewb_buffers* b, int* failFlag, int flags, cudaStream_t st) {
        using Rec = RecLayout<MC>;
        constexpr int NW = ((TY + 1) / 2) * ((TZ + 1) / 2);
        SweepArgs a;
        if (int rc = fillArgs<TY, TZ>(a, mp, b, failFlag, flags, st, NW)) return rc;
        auto kern = sweepKernel<MC, TL, TY, TZ>;
        const size_t smem = ((size_t)AccLayout<TY, TZ>::TABLES + (size_t)NW * 4 * Rec::PER_EL + (size_t)NW * 108) * sizeof(double);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return EWB_ERR_CUDA;
        const int64_t grid = (int64_t)a.tilesY * a.tilesZ * a.nChunks;
        kern<<<(unsigned)grid, NW * 32, smem, st>>>(a);
        return cudaGetLastError() == cudaSuccess ? EWB_OK : EWB_ERR_CUDA;
    }

    template <int MC, bool TL, int TY, int TZ, int NWP>
    int launchPC(const MatParams& mp, const ewb_buffers* b, int* failFlag, int flags, cudaStream_t st) {
        using Rec = RecLayout<MC>;
        constexpr int NW_ = ((TY + 1) / 2) * ((TZ + 1) / 2);
        SweepArgs a;
        if (int rc = fillArgs<TY, TZ>(a, mp, b, failFlag, flags, st, NW_)) return rc;
        auto kern = sweepKernelPC<MC, TL, TY, TZ, NWP>;
        constexpr int PEL = (MC == MC_LE && !TL) ? RecLayoutH::PER_EL : Rec::PER_EL;
        const size_t smem = ((size_t)AccLayout<TY, TZ>::TABLES + (size_t)2 * NW_ * 4 * PEL + (size_t)NWP * 216 + 32) * sizeof(double);
        if (smem > 232448) return EWB_ERR_UNSUPPORTED;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return EWB_ERR_CUDA;
        const int64_t grid = (int64_t)a.tilesY * a.tilesZ * a.nChunks;
        kern<<<(unsigned)grid, (NW_ + NWP) * 32, smem, st>>>(a);
        return cudaGetLastError() == cudaSuccess ? EWB_OK : EWB_ERR_CUDA;
    }

    // first-generation fused sweep (colour-ordered shared-memory accumulation)
    int launchV1(int elType, int mc, const MatParams& mp, const ewb_buffers* b, int* failFlag, int flags, cudaStream_t st) {
        if (elType == EWB_C3D8 && mc == MC_LE) return launchPC<MC_LE, false, 5, 5, 3>(mp, b, failFlag, flags, st);
        if (elType == EWB_C3D8 && mc == MC_VM) return launchT<MC_VM, false, 7, 5>(mp, b, failFlag, flags, st);
        if (elType == EWB_C3D8TL && mc == MC_NH) return launchT<MC_NH, true, 7, 5>(mp, b, failFlag, flags, st);
        return EWB_ERR_UNSUPPORTED;
    }
};

}  // namespace ewb
