// CTA-level element evaluation, shared by every assembly kernel.
//
// Two phases per element, both fed from shared memory (no per-element B cache in HBM — the
// reference keeps a 9.2 kB/element B operator, elements/displacementelement/element.py:194-205;
// here J, grad N are recomputed from the 24/60 nodal coordinates on the fly):
//
//   phase A  thread = (element, Gauss point): J, grad N, strain increment / F, constitutive
//            update, state write-back; publishes grad N (and the second per-node vector:
//            p_a = B_a^T n for von Mises, n_a = F^-T grad N_a for total Lagrange) plus the
//            per-Gauss-point tangent coefficients and -w detJ * stress in shared memory.
//   phase B  thread = (element, node a): the residual row P_a and the 3x3 stiffness blocks
//            K_ab for b = a, a+1, ... (cyclic; every unordered node pair exactly once, the
//            transposed block is emitted from the same registers).  Structured tangents
//             instead of dense B^T C B.
//
// Reference: DisplacementElement.computeYourself (elements/displacementelement/element.py:290-346),
// DisplacementTLElement.computeYourself (elements/displacementtlelement/element.py:346-427).
#pragma once
#include <type_traits>

#include "ewb_device.cuh"

namespace ewb {

template <int NN, int A = 0, class Fn>
__device__ __forceinline__ void forNodes(Fn&& f) {
    if constexpr (A < NN) {
        f(std::integral_constant<int, A>{});
        forNodes<NN, A + 1>(f);
    }
}

// Shared-memory image of one element (in doubles).
template <int NN, int NGP, int MC>
struct TileLayout {
    static constexpr bool HASQ = (MC != MC_LE);
    // Gauss-point stride of the vector tables.  8 nodes: odd (conflict-free phase-A stores, lane = Gauss point).  20 nodes: 60, which
    // makes the mma fragment loads of phaseBDmma20LE (lane = node row x Gauss point: 60 q + 3 r) conflict-free; they outnumber the
    // stores 4 : 1 (61 gave 4-way conflicts on the loads: 2x the ideal wavefront count in the ncu source view).
    static constexpr int GST = NN == 20 ? 60 : NN * 3 + 1;
    // per-Gauss-point scalars: [0..3] tangent coefficients (already * w detJ), [4..9] -w detJ * stress
    // (tensor order xx,yy,zz,xy,xz,yz), [10..18] F (only W_b needs it in phase B)
    // hypo-elastic TL: [10..18] F, [20..25] w detJ mu' (F F^T) (xx,yy,zz,xy,xz,yz); Q holds f_a = F grad N_a, R holds p_a = F N grad N_a
    static constexpr bool HASR = (MC == MC_TLV);
    static constexpr int NCO = isHypoTL(MC) ? 26 : (MC == MC_NH ? 20 : 11);  // LE / VM use slots 0..9; odd stride: conflict-free stores
    static constexpr int OFF_G = 0;
    static constexpr int OFF_Q = OFF_G + NGP * GST;
    static constexpr int OFF_R = OFF_Q + (HASQ ? NGP * GST : 0);
    static constexpr int OFF_CO = OFF_R + (HASR ? NGP * GST : 0);
    static constexpr int OFF_X = OFF_CO + NGP * NCO;  // nodal X, dU, U: [3][NN][3]
    static constexpr int RAW = OFF_X + 9 * NN;
    static constexpr int PER_EL = RAW + ((8 - RAW % 16) + 16) % 16;  // == 8 (mod 16) doubles: see DESIGN.md (banks)
};

// Gather nodal coordinates / U / dU of one element into its shared-memory image.
// Called by T cooperating threads (t = 0..T-1) of the element.
template <int NN, int NGP, int MC, int T>
__device__ __forceinline__ void stageNodes(double* sm, const int32_t* __restrict__ conn_e, const double* __restrict__ coords,
                                           const double* __restrict__ U, const double* __restrict__ dU, int t) {
    using L = TileLayout<NN, NGP, MC>;
    for (int a = t; a < NN; a += T) {
        const int64_t n = conn_e[a];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            sm[L::OFF_X + a * 3 + c] = coords[3 * n + c];
            sm[L::OFF_X + 3 * NN + a * 3 + c] = dU[3 * n + c];
            sm[L::OFF_X + 6 * NN + a * 3 + c] = U[3 * n + c];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// phase A
// ---------------------------------------------------------------------------------------------
// state_ref / state_temp point at component 0 of (element e, this gp); cstride = nEl * NGP.
// L: shared-memory layout (GST, NCO, OFF_G, OFF_Q, OFF_CO, HASQ, and the CO slot indices C_*).
// X: nodal coordinates [NN][3]; uu: nodal dU (small strain) or U (TL) [NN][3] — shared memory or
// registers (all indices are compile-time after unrolling).
template <class L, int NN, int NGP, int MC, bool TL>
__device__ __forceinline__ void gaussPointL(double* sm, const double* X, const double* uu, int gp, const MatParams& mp,
                                            const double* __restrict__ state_ref, double* __restrict__ state_temp, int64_t cstride,
                                            bool writeState, int* failFlag) {
    double* G = sm + L::OFF_G + gp * L::GST;
    double* CO = sm + L::OFF_CO + gp * L::NCO;

    double xi, eta, zeta, w;
    Gauss<NGP>::get(gp, xi, eta, zeta, w);

    double Jm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    forNodes<NN>([&](auto ic) {
        constexpr int a = decltype(ic)::value;
        double d[3];
        shapeDeriv<NN, a>(xi, eta, zeta, d);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) Jm[r * 3 + c] = fma(d[r], X[a * 3 + c], Jm[r * 3 + c]);
    });
    const double detJ = det3(Jm);
    double iJ[9];
    inv3(Jm, detJ, iJ);
    const double wd = w * detJ;

    // grad N_a = J^-1 dN_a ; displacement(-increment) gradient H[i][c] = sum_a u_a[i] g_a[c]
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    forNodes<NN>([&](auto ic) {
        constexpr int a = decltype(ic)::value;
        double d[3], g[3];
        shapeDeriv<NN, a>(xi, eta, zeta, d);
#pragma unroll
        for (int c = 0; c < 3; ++c) g[c] = iJ[c * 3 + 0] * d[0] + iJ[c * 3 + 1] * d[1] + iJ[c * 3 + 2] * d[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) G[a * 3 + c] = g[c];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) H[i * 3 + c] = fma(uu[a * 3 + i], g[c], H[i * 3 + c]);
    });

    double st[13];
#pragma unroll
    for (int c = 0; c < 12 + matStateCount(MC); ++c) st[c] = state_ref[c * cstride];

    if constexpr (!TL) {
        // Voigt 11,22,33,12,13,23 with engineering shear (_B3D8)
        const double de[6] = {H[0], H[4], H[8], H[1] + H[3], H[2] + H[6], H[5] + H[7]};
        double s[6] = {st[0], st[1], st[2], st[3], st[4], st[5]};
        if constexpr (MC == MC_LE) {
            hookeAdd(mp, de, s);
            CO[0] = wd;
        } else {
            VMResult r;
            double kappa = st[12];
            vonMises(mp, de, s, kappa, r);
            st[12] = kappa;
            if (r.failed) atomicOr(failFlag, 1);
            CO[0] = wd * r.lam;
            CO[1] = wd * r.mu;
            CO[2] = -wd * r.a;
            if constexpr (L::NCO >= 16) {
#pragma unroll
                for (int i = 0; i < 6; ++i) CO[10 + i] = r.n[i];
            }
            // p_a = B_a^T n = N g_a, N = tensor(n)  (n Voigt 11,22,33,12,13,23)
            if constexpr (L::HASQ) {
                double* Q = sm + L::OFF_Q + gp * L::GST;
                for (int a = 0; a < NN; ++a) {
                    const double gx = G[a * 3], gy = G[a * 3 + 1], gz = G[a * 3 + 2];
                    Q[a * 3 + 0] = r.n[0] * gx + r.n[3] * gy + r.n[4] * gz;
                    Q[a * 3 + 1] = r.n[3] * gx + r.n[1] * gy + r.n[5] * gz;
                    Q[a * 3 + 2] = r.n[4] * gx + r.n[5] * gy + r.n[2] * gz;
                }
            }
        }
        // -w detJ * stress as a tensor (xx,yy,zz,xy,xz,yz): Voigt 3,4,5 = 12,13,23
#pragma unroll
        for (int i = 0; i < 6; ++i) CO[4 + i] = -wd * s[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            st[i] = s[i];
            st[6 + i] += de[i];
        }
    } else if constexpr (isHypoTL(MC)) {
        // total Lagrange with a small-strain material law on (PK2, Green-Lagrange increment): element.py:415-425
        double F[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) F[i] = H[i];
        F[0] += 1.0;
        F[4] += 1.0;
        F[8] += 1.0;
        // E = (H + H^T + H^T H)/2, Voigt 11,22,33,2*12,2*23,2*13 (voigtnotation.py:32-52); E_old = accepted strain state
        double E[6];
        E[0] = H[0] + 0.5 * (H[0] * H[0] + H[3] * H[3] + H[6] * H[6]);
        E[1] = H[4] + 0.5 * (H[1] * H[1] + H[4] * H[4] + H[7] * H[7]);
        E[2] = H[8] + 0.5 * (H[2] * H[2] + H[5] * H[5] + H[8] * H[8]);
        E[3] = H[1] + H[3] + H[0] * H[1] + H[3] * H[4] + H[6] * H[7];
        E[4] = H[5] + H[7] + H[1] * H[2] + H[4] * H[5] + H[7] * H[8];
        E[5] = H[2] + H[6] + H[0] * H[2] + H[3] * H[5] + H[6] * H[8];
        double de[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) de[i] = E[i] - st[6 + i];
        double s[6] = {st[0], st[1], st[2], st[3], st[4], st[5]};  // PK2, Voigt 11,22,33,12,23,13
        double cm;
        double* Q = sm + L::OFF_Q + gp * L::GST;
        if constexpr (MC == MC_TLE) {
            hookeAdd(mp, de, s);
            CO[0] = wd * mp.lambda;
            cm = wd * mp.G;
            CO[2] = 0.0;
        } else {
            VMResult r;
            double kappa = st[12];
            vonMises(mp, de, s, kappa, r);
            st[12] = kappa;
            if (r.failed) atomicOr(failFlag, 1);
            CO[0] = wd * r.lam;
            cm = wd * r.mu;
            CO[2] = -wd * r.a;
            // p_a = B_a^T n = F N grad N_a, N = tensor(n), n Voigt 11,22,33,12,23,13
            double* R = sm + L::OFF_R + gp * L::GST;
            for (int a = 0; a < NN; ++a) {
                const double gx = G[a * 3], gy = G[a * 3 + 1], gz = G[a * 3 + 2];
                const double tx = r.n[0] * gx + r.n[3] * gy + r.n[5] * gz;
                const double ty = r.n[3] * gx + r.n[1] * gy + r.n[4] * gz;
                const double tz = r.n[5] * gx + r.n[4] * gy + r.n[2] * gz;
#pragma unroll
                for (int k = 0; k < 3; ++k) R[a * 3 + k] = F[k * 3] * tx + F[k * 3 + 1] * ty + F[k * 3 + 2] * tz;
            }
        }
        CO[1] = cm;
        CO[3] = 0.0;
        // -w detJ * S as a tensor (xx,yy,zz,xy,xz,yz): Voigt 3,4,5 = 12,23,13
        CO[4] = -wd * s[0];
        CO[5] = -wd * s[1];
        CO[6] = -wd * s[2];
        CO[7] = -wd * s[3];
        CO[8] = -wd * s[5];
        CO[9] = -wd * s[4];
#pragma unroll
        for (int i = 0; i < 9; ++i) CO[10 + i] = F[i];
        CO[19] = 0.0;
        // w detJ mu' (F F^T)
        CO[20] = cm * (F[0] * F[0] + F[1] * F[1] + F[2] * F[2]);
        CO[21] = cm * (F[3] * F[3] + F[4] * F[4] + F[5] * F[5]);
        CO[22] = cm * (F[6] * F[6] + F[7] * F[7] + F[8] * F[8]);
        CO[23] = cm * (F[0] * F[3] + F[1] * F[4] + F[2] * F[5]);
        CO[24] = cm * (F[0] * F[6] + F[1] * F[7] + F[2] * F[8]);
        CO[25] = cm * (F[3] * F[6] + F[4] * F[7] + F[5] * F[8]);
        // f_a = F grad N_a
        for (int a = 0; a < NN; ++a) {
            const double gx = G[a * 3], gy = G[a * 3 + 1], gz = G[a * 3 + 2];
#pragma unroll
            for (int k = 0; k < 3; ++k) Q[a * 3 + k] = F[k * 3] * gx + F[k * 3 + 1] * gy + F[k * 3 + 2] * gz;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            st[i] = s[i];
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
        CO[0] = wd * r.c0;
        CO[1] = wd * r.c1;
        CO[2] = wd * r.c2;
        CO[3] = wd * r.c4;
#pragma unroll
        for (int i = 0; i < 6; ++i) CO[4 + i] = -wd * r.tau[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) CO[10 + i] = F[i];
        if constexpr (L::NCO >= 28) {
#pragma unroll
            for (int i = 0; i < 9; ++i) CO[19 + i] = iF[i];
        }
        // n_a[m] = sum_j g_a[j] Finv[j][m]
        if constexpr (L::HASQ) {
            double* Q = sm + L::OFF_Q + gp * L::GST;
            for (int a = 0; a < NN; ++a) {
                const double gx = G[a * 3], gy = G[a * 3 + 1], gz = G[a * 3 + 2];
#pragma unroll
                for (int m = 0; m < 3; ++m) Q[a * 3 + m] = gx * iF[m] + gy * iF[3 + m] + gz * iF[6 + m];
            }
        }
        // Green-Lagrange E = (H + H^T + H^T H)/2, Voigt strain 11,22,33,2*12,2*23,2*13 (voigtnotation.py:32-52)
        double E[6];
        E[0] = H[0] + 0.5 * (H[0] * H[0] + H[3] * H[3] + H[6] * H[6]);
        E[1] = H[4] + 0.5 * (H[1] * H[1] + H[4] * H[4] + H[7] * H[7]);
        E[2] = H[8] + 0.5 * (H[2] * H[2] + H[5] * H[5] + H[8] * H[8]);
        const double E01 = 0.5 * (H[1] + H[3] + H[0] * H[1] + H[3] * H[4] + H[6] * H[7]);
        const double E12 = 0.5 * (H[5] + H[7] + H[1] * H[2] + H[4] * H[5] + H[7] * H[8]);
        const double E02 = 0.5 * (H[2] + H[6] + H[0] * H[2] + H[3] * H[5] + H[6] * H[8]);
        E[3] = 2.0 * E01;
        E[4] = 2.0 * E12;
        E[5] = 2.0 * E02;
        // Kirchhoff stress Voigt 11,22,33,12,23,13 (voigtnotation.py:74-93); tau[] is xx,yy,zz,xy,xz,yz
        st[0] = r.tau[0];
        st[1] = r.tau[1];
        st[2] = r.tau[2];
        st[3] = r.tau[3];
        st[4] = r.tau[5];
        st[5] = r.tau[4];
#pragma unroll
        for (int i = 0; i < 6; ++i) st[6 + i] = E[i];
        st[12] = r.energy;
    }
    if (writeState) {
#pragma unroll
        for (int c = 0; c < 12 + matStateCount(MC); ++c) state_temp[c * cstride] = st[c];
    }
}

template <int NN, int NGP, int MC, bool TL>
__device__ __forceinline__ void gaussPoint(double* sm, int gp, const MatParams& mp, const double* __restrict__ state_ref,
                                           double* __restrict__ state_temp, int64_t cstride, bool writeState, int* failFlag) {
    using L = TileLayout<NN, NGP, MC>;
    const double* X = sm + L::OFF_X;
    const double* uu = sm + L::OFF_X + (TL ? 6 * NN : 3 * NN);
    gaussPointL<L, NN, NGP, MC, TL>(sm, X, uu, gp, mp, state_ref, state_temp, cstride, writeState, failFlag);
}

// ---------------------------------------------------------------------------------------------
// phase B
// ---------------------------------------------------------------------------------------------
// Emit must provide:  void residual(int a, const double P[3]);
//                     void block(int a, int b, const double K[9]);   // K[i*3+j] = Ke[3a+i][3b+j]
template <int NN, int NGP, int MC, int BLK, class Emit>
__device__ __forceinline__ void nodeRow(const double* sm, int a, const MatParams& mp, bool wantK, Emit& emit, int passBegin = 0,
                                        int passEnd = 1 << 30, bool doResidual = true) {
    using L = TileLayout<NN, NGP, MC>;
    constexpr int HALF = NN / 2;
    const double* G = sm + L::OFF_G;
    const double* Q = sm + L::OFF_Q;
    const double* CO = sm + L::OFF_CO;
    const int nb = HALF + (a < HALF ? 1 : 0);
    constexpr int NPASS = (HALF + 1 + BLK - 1) / BLK;

    // residual row: P_a = sum_gp (-w detJ S) v_a, v = grad N_a (small strain) or n_a (TL)
    if (doResidual) {
        double P[3] = {0, 0, 0};
        for (int gp = 0; gp < NGP; ++gp) {
            const double* v = (MC == MC_NH ? Q : G) + gp * L::GST + a * 3;
            const double* S = CO + gp * L::NCO + 4;
            const double vx = v[0], vy = v[1], vz = v[2];
            const double t0 = S[0] * vx + S[3] * vy + S[4] * vz;
            const double t1 = S[3] * vx + S[1] * vy + S[5] * vz;
            const double t2 = S[4] * vx + S[5] * vy + S[2] * vz;
            if constexpr (isHypoTL(MC)) {  // P_a = -w detJ B_a^T S = F ((-w detJ S) grad N_a)
                const double* F = CO + gp * L::NCO + 10;
#pragma unroll
                for (int k = 0; k < 3; ++k) P[k] += F[k * 3] * t0 + F[k * 3 + 1] * t1 + F[k * 3 + 2] * t2;
            } else {
                P[0] += t0;
                P[1] += t1;
                P[2] += t2;
            }
        }
        emit.residual(a, P);
    }
    if (!wantK) return;

#pragma unroll 1
    for (int pass = passBegin; pass < min(passEnd, NPASS); ++pass) {
        double acc[BLK][9];
#pragma unroll
        for (int k = 0; k < BLK; ++k)
#pragma unroll
            for (int i = 0; i < 9; ++i) acc[k][i] = 0.0;

#pragma unroll 1
        for (int gp = 0; gp < NGP; ++gp) {
            const double* Gg = G + gp * L::GST;
            const double* Qg = Q + gp * L::GST;
            const double* co = CO + gp * L::NCO;
            const double ga[3] = {Gg[a * 3], Gg[a * 3 + 1], Gg[a * 3 + 2]};
            if constexpr (MC == MC_LE) {
                const double w = co[0];
                const double ha[3] = {w * ga[0], w * ga[1], w * ga[2]};
#pragma unroll
                for (int k = 0; k < BLK; ++k) {
                    const int d = pass * BLK + k;
                    if (d < nb) {
                        int b = a + d;
                        if (b >= NN) b -= NN;
                        const double gb[3] = {Gg[b * 3], Gg[b * 3 + 1], Gg[b * 3 + 2]};
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 3; ++j) acc[k][i * 3 + j] = fma(ha[i], gb[j], acc[k][i * 3 + j]);
                    }
                }
            } else if constexpr (MC == MC_VM) {
                const double cl = co[0], cm = co[1], ca = co[2];
                const double pa[3] = {Qg[a * 3], Qg[a * 3 + 1], Qg[a * 3 + 2]};
                const double la[3] = {cl * ga[0], cl * ga[1], cl * ga[2]};
                const double ma[3] = {cm * ga[0], cm * ga[1], cm * ga[2]};
                const double ra[3] = {ca * pa[0], ca * pa[1], ca * pa[2]};
#pragma unroll
                for (int k = 0; k < BLK; ++k) {
                    const int d = pass * BLK + k;
                    if (d < nb) {
                        int b = a + d;
                        if (b >= NN) b -= NN;
                        const double gb[3] = {Gg[b * 3], Gg[b * 3 + 1], Gg[b * 3 + 2]};
                        const double pb[3] = {Qg[b * 3], Qg[b * 3 + 1], Qg[b * 3 + 2]};
                        const double dot = ma[0] * gb[0] + ma[1] * gb[1] + ma[2] * gb[2];
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                double v = acc[k][i * 3 + j];
                                v = fma(la[i], gb[j], v);
                                v = fma(ma[j], gb[i], v);
                                v = fma(ra[i], pb[j], v);
                                acc[k][i * 3 + j] = v;
                            }
                        acc[k][0] += dot;
                        acc[k][4] += dot;
                        acc[k][8] += dot;
                    }
                }
            } else if constexpr (isHypoTL(MC)) {
                // K_ab = lam' f_a f_b^T + mu' f_b f_a^T + (g_a.g_b) mu' F F^T + (g_a^T S g_b) I - a p_a p_b^T   (all * w detJ)
                const double cl = co[0], cm = co[1], ca = co[2];
                const double* Rg = sm + L::OFF_R + gp * L::GST;
                const double fa[3] = {Qg[a * 3], Qg[a * 3 + 1], Qg[a * 3 + 2]};
                const double la[3] = {cl * fa[0], cl * fa[1], cl * fa[2]};
                const double ma[3] = {cm * fa[0], cm * fa[1], cm * fa[2]};
                double ra[3] = {0, 0, 0};
                if constexpr (L::HASR) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) ra[i] = ca * Rg[a * 3 + i];
                }
                // (w detJ S) g_a  (co[4..9] holds -w detJ S)
                const double sa[3] = {-(co[4] * ga[0] + co[7] * ga[1] + co[8] * ga[2]), -(co[7] * ga[0] + co[5] * ga[1] + co[9] * ga[2]),
                                      -(co[8] * ga[0] + co[9] * ga[1] + co[6] * ga[2])};
                const double* Bm = co + 20;
#pragma unroll
                for (int k = 0; k < BLK; ++k) {
                    const int d = pass * BLK + k;
                    if (d < nb) {
                        int b = a + d;
                        if (b >= NN) b -= NN;
                        const double gb[3] = {Gg[b * 3], Gg[b * 3 + 1], Gg[b * 3 + 2]};
                        const double fb[3] = {Qg[b * 3], Qg[b * 3 + 1], Qg[b * 3 + 2]};
                        const double gg = ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2];
                        const double geo = sa[0] * gb[0] + sa[1] * gb[1] + sa[2] * gb[2];
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                double v = acc[k][i * 3 + j];
                                v = fma(la[i], fb[j], v);
                                v = fma(fb[i], ma[j], v);
                                acc[k][i * 3 + j] = v;
                            }
                        if constexpr (L::HASR) {
                            const double pb[3] = {Rg[b * 3], Rg[b * 3 + 1], Rg[b * 3 + 2]};
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int j = 0; j < 3; ++j) acc[k][i * 3 + j] = fma(ra[i], pb[j], acc[k][i * 3 + j]);
                        }
                        acc[k][0] += gg * Bm[0] + geo;
                        acc[k][4] += gg * Bm[1] + geo;
                        acc[k][8] += gg * Bm[2] + geo;
                        acc[k][1] += gg * Bm[3];
                        acc[k][3] += gg * Bm[3];
                        acc[k][2] += gg * Bm[4];
                        acc[k][6] += gg * Bm[4];
                        acc[k][5] += gg * Bm[5];
                        acc[k][7] += gg * Bm[5];
                    }
                }
            } else {
                const double c0 = co[0], c1 = co[1], c2 = co[2], c4 = co[3];
                const double na[3] = {Qg[a * 3], Qg[a * 3 + 1], Qg[a * 3 + 2]};
                const double g0[3] = {c0 * ga[0], c0 * ga[1], c0 * ga[2]};
                const double n1[3] = {c1 * na[0], c1 * na[1], c1 * na[2]};
                const double n2[3] = {c2 * na[0], c2 * na[1], c2 * na[2]};
                const bool wb = (mp.kind == 3);
                double fa[3] = {0, 0, 0}, n4[3] = {0, 0, 0};
                if (wb) {
                    const double* F = co + 10;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        fa[i] = c4 * (F[i * 3] * ga[0] + F[i * 3 + 1] * ga[1] + F[i * 3 + 2] * ga[2]);
                        n4[i] = c4 * na[i];
                    }
                }
#pragma unroll
                for (int k = 0; k < BLK; ++k) {
                    const int d = pass * BLK + k;
                    if (d < nb) {
                        int b = a + d;
                        if (b >= NN) b -= NN;
                        const double gb[3] = {Gg[b * 3], Gg[b * 3 + 1], Gg[b * 3 + 2]};
                        const double nbv[3] = {Qg[b * 3], Qg[b * 3 + 1], Qg[b * 3 + 2]};
                        const double dot = g0[0] * gb[0] + g0[1] * gb[1] + g0[2] * gb[2];
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                double v = acc[k][i * 3 + j];
                                v = fma(n1[i], nbv[j], v);
                                v = fma(n2[j], nbv[i], v);
                                acc[k][i * 3 + j] = v;
                            }
                        if (wb) {
                            const double* F = co + 10;
                            double fb[3];
#pragma unroll
                            for (int i = 0; i < 3; ++i) fb[i] = F[i * 3] * gb[0] + F[i * 3 + 1] * gb[1] + F[i * 3 + 2] * gb[2];
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int j = 0; j < 3; ++j) {
                                    double v = acc[k][i * 3 + j];
                                    v = fma(fa[i], nbv[j], v);
                                    v = fma(n4[i], fb[j], v);
                                    acc[k][i * 3 + j] = v;
                                }
                        }
                        acc[k][0] += dot;
                        acc[k][4] += dot;
                        acc[k][8] += dot;
                    }
                }
            }
        }

        if constexpr (Emit::HALF) {
            // half-block scratch: the thread's blocks (a, a + d) of this pass are one contiguous run of BLK * 9 doubles
            double out[BLK * 9 + 1];
            out[BLK * 9] = 0.0;
#pragma unroll
            for (int k = 0; k < BLK; ++k) {
                if constexpr (MC == MC_LE) {
                    const double* M = acc[k];
                    const double tr = mp.G * (M[0] + M[4] + M[8]);
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) out[k * 9 + i * 3 + j] = mp.lambda * M[i * 3 + j] + mp.G * M[j * 3 + i] + (i == j ? tr : 0.0);
                } else {
#pragma unroll
                    for (int i = 0; i < 9; ++i) out[k * 9 + i] = acc[k][i];
                }
            }
            emit.template pass<BLK>(a, pass, out);
        } else {
#pragma unroll
        for (int k = 0; k < BLK; ++k) {
            const int d = pass * BLK + k;
            if (d < nb) {
                int b = a + d;
                if (b >= NN) b -= NN;
                double Kb[9];
                if constexpr (MC == MC_LE) {
                    const double* M = acc[k];
                    const double tr = mp.G * (M[0] + M[4] + M[8]);
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) Kb[i * 3 + j] = mp.lambda * M[i * 3 + j] + mp.G * M[j * 3 + i] + (i == j ? tr : 0.0);
                } else {
#pragma unroll
                    for (int i = 0; i < 9; ++i) Kb[i] = acc[k][i];
                }
                emit.block(a, b, Kb);
                if (d != 0) {
                    const double Kt[9] = {Kb[0], Kb[3], Kb[6], Kb[1], Kb[4], Kb[7], Kb[2], Kb[5], Kb[8]};
                    emit.block(b, a, Kt);
                }
            }
        }
        }
    }
}

}  // namespace ewb
