// Device-side building blocks shared by the generic (any connectivity) and the fused
// BoxGen-sweep assembly kernels: shape-function derivatives, Gauss points, 3x3 algebra and
// the constitutive updates.  FP64 throughout.  sm_100a.
//
// Reference behaviour restated here (paths below /root/reference/edelweissfe/):
//   elements/library.py:34-47,212-227,260-275       Gauss points / weights
//   elements/displacementtlelement/_elementcomputationmatrices.py:370-498   dN tables (row order d/d eta, d/d xi, d/d zeta)
//   elements/displacementelement/_elementcomputationmatrices.py:306-366,700-817   J, B (Voigt 11,22,33,12,13,23)
//   materials/linearelastic/linearelastic.py:95-117,185-210
//   materials/vonmises/vonmises.py:186-254
//   materials/neohooke/neohookepencegouformulation{a,b,c}.py:130-145
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ewb {

// material classes (kernel template parameter).  MC_TLE / MC_TLV: linear elastic / von Mises inside the total-Lagrange
// element (non-hyperelastic branch with geometric stiffness, displacementtlelement/element.py:415-425); arbitrary-mesh path only.
enum : int { MC_LE = 0, MC_VM = 1, MC_NH = 2, MC_TLE = 3, MC_TLV = 4 };
__host__ __device__ constexpr int matStateCount(int mc) { return (mc == MC_LE || mc == MC_TLE) ? 0 : 1; }
__host__ __device__ constexpr bool isHypoTL(int mc) { return mc == MC_TLE || mc == MC_TLV; }

// Material parameters, preprocessed on the host and passed by value.
struct MatParams {
    int kind;        // EWB_MAT_*
    double lambda;   // Lame lambda (small strain)
    double G;        // shear modulus
    double fy0, HLin, dfy, delta;  // von Mises hardening
    double mu, K;    // Neo-Hooke
};

// ---------------------------------------------------------------------------------------------
// local node coordinates in the reference's (xi, eta, zeta) naming; eta runs along BoxGen x,
// xi along y, zeta along z.
// ---------------------------------------------------------------------------------------------
template <int NN> struct NodeLC;
template <> struct NodeLC<8> {
    __host__ __device__ static constexpr int xi(int a) { return (a & 4) ? 1 : -1; }
    __host__ __device__ static constexpr int eta(int a) { return (a & 2) ? 1 : -1; }
    __host__ __device__ static constexpr int zeta(int a) { return ((a ^ (a >> 1)) & 1) ? 1 : -1; }
};
template <> struct NodeLC<20> {
    // corners as Hexa8, then mid-edge nodes (generators/boxgen.py:187-299 offsets on the doubled grid)
    __host__ __device__ static constexpr int xi(int a) {
        constexpr int t[20] = {-1, -1, -1, -1, 1, 1, 1, 1, -1, -1, -1, -1, 1, 1, 1, 1, 0, 0, 0, 0};
        return t[a];
    }
    __host__ __device__ static constexpr int eta(int a) {
        constexpr int t[20] = {-1, -1, 1, 1, -1, -1, 1, 1, -1, 0, 1, 0, -1, 0, 1, 0, -1, -1, 1, 1};
        return t[a];
    }
    __host__ __device__ static constexpr int zeta(int a) {
        constexpr int t[20] = {-1, 1, 1, -1, -1, 1, 1, -1, 0, 1, 0, -1, 0, 1, 0, -1, -1, 1, 1, -1};
        return t[a];
    }
};

// Gauss point (xi, eta, zeta, w) of point gp.
template <int NGP> struct Gauss;
template <> struct Gauss<8> {
    __device__ static void get(int gp, double& xi, double& eta, double& zeta, double& w) {
        const double g = 0.57735026918962576451;  // 1/sqrt(3)
        xi = (gp & 4) ? g : -g;
        eta = (gp & 2) ? g : -g;
        zeta = ((gp ^ (gp >> 1)) & 1) ? g : -g;
        w = 1.0;
    }
};
template <> struct Gauss<1> {  // C3D8R: elements/library.py:228-243
    __device__ static void get(int, double& xi, double& eta, double& zeta, double& w) {
        xi = eta = zeta = 0.0;
        w = 8.0;
    }
};
template <> struct Gauss<27> {
    __device__ static void get(int gp, double& xi, double& eta, double& zeta, double& w) {
        const double r = 0.77459666924148337704;  // sqrt(0.6)
        const int ix = gp / 9, ie = (gp % 9) / 3, iz = gp % 3;
        xi = r * (ix - 1);
        eta = r * (ie - 1);
        zeta = r * (iz - 1);
        const double w5 = 5.0 / 9.0, w8 = 8.0 / 9.0;
        w = (ix == 1 ? w8 : w5) * (ie == 1 ? w8 : w5) * (iz == 1 ? w8 : w5);
    }
};

// Shape-function derivatives of node A at (xi,eta,zeta): d[0]=dN/d eta, d[1]=dN/d xi, d[2]=dN/d zeta.
template <int NN, int A>
__device__ __forceinline__ void shapeDeriv(double xi, double eta, double zeta, double d[3]) {
    constexpr int a = NodeLC<NN>::xi(A), b = NodeLC<NN>::eta(A), c = NodeLC<NN>::zeta(A);
    const double fx = 1.0 + a * xi, fe = 1.0 + b * eta, fz = 1.0 + c * zeta;
    if constexpr (NN == 8) {
        d[0] = 0.125 * b * fx * fz;
        d[1] = 0.125 * a * fe * fz;
        d[2] = 0.125 * c * fx * fe;
    } else {
        if constexpr (a == 0) {
            const double q = 0.25 * (1.0 - xi * xi);
            d[1] = -0.5 * xi * fe * fz;
            d[0] = q * b * fz;
            d[2] = q * fe * c;
        } else if constexpr (b == 0) {
            const double q = 0.25 * (1.0 - eta * eta);
            d[1] = q * a * fz;
            d[0] = -0.5 * eta * fx * fz;
            d[2] = q * fx * c;
        } else if constexpr (c == 0) {
            const double q = 0.25 * (1.0 - zeta * zeta);
            d[1] = q * a * fe;
            d[0] = q * fx * b;
            d[2] = -0.5 * zeta * fx * fe;
        } else {
            const double s = a * xi + b * eta + c * zeta - 2.0;
            d[1] = 0.125 * a * fe * fz * (s + fx);
            d[0] = 0.125 * b * fx * fz * (s + fe);
            d[2] = 0.125 * c * fx * fe * (s + fz);
        }
    }
}

// 3x3 helpers (row-major m[r*3+c])
__device__ __forceinline__ double det3(const double* m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
__device__ __forceinline__ void inv3(const double* m, double det, double* r) {
    const double id = 1.0 / det;
    r[0] = (m[4] * m[8] - m[5] * m[7]) * id;
    r[1] = (m[2] * m[7] - m[1] * m[8]) * id;
    r[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    r[3] = (m[5] * m[6] - m[3] * m[8]) * id;
    r[4] = (m[0] * m[8] - m[2] * m[6]) * id;
    r[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    r[6] = (m[3] * m[7] - m[4] * m[6]) * id;
    r[7] = (m[1] * m[6] - m[0] * m[7]) * id;
    r[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// ---------------------------------------------------------------------------------------------
// Constitutive updates.  Voigt order of the small-strain element: 11,22,33,12,13,23,
// engineering shear (displacementelement/_elementcomputationmatrices.py:700-712).
// ---------------------------------------------------------------------------------------------

// sigma += C : de  for isotropic Hooke (linearelastic.py:185-210)
__device__ __forceinline__ void hookeAdd(const MatParams& mp, const double de[6], double s[6]) {
    const double tr = de[0] + de[1] + de[2];
    const double l = mp.lambda * tr, twoG = 2.0 * mp.G;
    s[0] += l + twoG * de[0];
    s[1] += l + twoG * de[1];
    s[2] += l + twoG * de[2];
    s[3] += mp.G * de[3];
    s[4] += mp.G * de[4];
    s[5] += mp.G * de[5];
}

struct VMResult {
    double lam, mu;  // effective isotropic moduli of the tangent: C = lam 1x1 + 2 mu Isym - a n x n
    double a;        // rank-one coefficient (0 on elastic steps)
    double n[6];     // flow direction (Voigt, shear entries NOT doubled)
    bool failed;
};

// J2 plasticity, radial return; scalar Newton on d-kappa: start 0, stop |R| <= 1e-12, at most 15
// updates, else request a cut-back (vonmises.py:186-254).  Tangent in the structured form
// verified in SURVEY §3.3.
__device__ __forceinline__ void vonMises(const MatParams& mp, const double de[6], double s[6], double& kappa, VMResult& r) {
    r.lam = mp.lambda;
    r.mu = mp.G;
    r.a = 0.0;
    r.failed = false;
#pragma unroll
    for (int i = 0; i < 6; ++i) r.n[i] = 0.0;
    double nrm2 = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) nrm2 += de[i] * de[i];
    if (sqrt(nrm2) < 1e-14) return;  // zero increment: stress untouched (vonmises.py:211-213)
    double t[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) t[i] = s[i];
    hookeAdd(mp, de, t);  // elastic predictor
    const double third = 1.0 / 3.0;
    const double pm = (t[0] + t[1] + t[2]) * third;
    double d[6] = {t[0] - pm, t[1] - pm, t[2] - pm, t[3], t[4], t[5]};
    const double devn = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + 2.0 * (d[3] * d[3] + d[4] * d[4] + d[5] * d[5]));
    const double s23 = 0.81649658092772603273;  // sqrt(2/3)
    const double s6 = 2.44948974278317809820;   // sqrt(6)
    const double k0 = kappa;
    const double fy = mp.fy0 + mp.HLin * k0 + mp.dfy * (1.0 - exp(-mp.delta * k0));
    if (devn - s23 * fy > 0.0) {
        double dk = 0.0;
        int counter = 0;
        const double G = mp.G;
        double ex;
        while (true) {
            ex = exp(-mp.delta * (k0 + dk));
            const double R = devn - s6 * G * dk - s23 * (mp.fy0 + mp.HLin * (k0 + dk) + mp.dfy * (1.0 - ex));
            if (!(fabs(R) > 1e-12)) break;
            if (counter == 15) {
                r.failed = true;
                break;
            }
            const double dR = -s6 * G - s23 * (mp.HLin + mp.dfy * mp.delta * ex);
            dk -= R / dR;
            ++counter;
        }
        const double dLambda = 1.22474487139158904909 * dk;  // sqrt(3/2)
        kappa = k0 + dk;
        const double inv = 1.0 / devn;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            r.n[i] = d[i] * inv;
            s[i] = t[i] - 2.0 * G * dLambda * r.n[i];
        }
        const double dfyk = mp.HLin + mp.dfy * mp.delta * ex;  // ex == exp(-delta * kappa): the loop's last evaluation
        const double bcoef = 4.0 * G * G * dLambda * inv;
        r.a = 2.0 * G * (1.0 / (1.0 + dfyk / (3.0 * G)) - 2.0 * G * dLambda * inv);
        r.lam = mp.lambda + bcoef * third;
        r.mu = G - 0.5 * bcoef;
    } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) s[i] = t[i];
    }
}

struct NHResult {
    double tau[6];              // Kirchhoff stress, tensor order xx,yy,zz,xy,xz,yz
    double c0, c1, c2, c4;      // K_ab = c0 (g_a.g_b) I + c1 n_a n_b^T + c2 n_b n_a^T + c4 (f_a n_b^T + n_a f_b^T)
    double energy;
};

// Compressible Neo-Hooke, Pence-Gou W_a / W_b / W_c.  Closed forms of n_a . A . g_b minus the
// geometric term (derivation in DESIGN.md; checked against the dense 3^4 contraction of the oracle).
__device__ __forceinline__ void neoHooke(const MatParams& mp, const double F[9], double J, NHResult& r) {
    // B = F F^T
    double B[6];
    B[0] = F[0] * F[0] + F[1] * F[1] + F[2] * F[2];
    B[1] = F[3] * F[3] + F[4] * F[4] + F[5] * F[5];
    B[2] = F[6] * F[6] + F[7] * F[7] + F[8] * F[8];
    B[3] = F[0] * F[3] + F[1] * F[4] + F[2] * F[5];
    B[4] = F[0] * F[6] + F[1] * F[7] + F[2] * F[8];
    B[5] = F[3] * F[6] + F[4] * F[7] + F[5] * F[8];
    const double mu = mp.mu, K = mp.K;
    double sB, sI;  // tau = sB * B + sI * I
    if (mp.kind == 2) {  // W_a
        const double I1 = B[0] + B[1] + B[2];
        const double lamBar = (K - 2.0 / 3.0 * mu) * (J * J - J) - mu;
        const double muBar = (K - 2.0 / 3.0 * mu) * (2.0 * J * J - J);
        sB = mu;
        sI = lamBar;
        r.c0 = mu;
        r.c1 = muBar;
        r.c2 = -lamBar;
        r.c4 = 0.0;
        r.energy = mu / 2.0 * (I1 - 3.0) + (K / 2.0 - mu / 3.0) * (J - 1.0) * (J - 1.0) - mu * log(J);
    } else if (mp.kind == 3) {  // W_b
        const double I1 = B[0] + B[1] + B[2];
        const double J23 = pow(J, 2.0 / 3.0);
        const double J2 = J * J;
        const double lamHat = K / 2.0 * (J2 + 1.0 / J2);
        const double muBar = mu / (3.0 * J23);
        const double lamBar = K / 4.0 * (J2 - 1.0 / J2) - muBar * I1;
        sB = mu / J23;
        sI = lamBar;
        r.c0 = 3.0 * muBar;
        r.c1 = lamHat + 2.0 / 3.0 * I1 * muBar;
        r.c2 = -lamBar;
        r.c4 = -2.0 * muBar;
        r.energy = mu / 2.0 * (I1 / J23 - 3.0) + K / 8.0 * (J2 + 1.0 / J2 - 2.0);
    } else {  // W_c   (the reference takes I1 = trace(F) here, neohookepencegouformulationc.py:134)
        const double I1 = F[0] + F[4] + F[8];
        const double pw = pow(J, 2.0 / 3.0 - K / mu);
        const double muBar = mu * pw;
        const double lamBar = (K / mu - 2.0 / 3.0) * muBar;
        sB = mu;
        sI = -muBar;
        r.c0 = mu;
        r.c1 = lamBar;
        r.c2 = muBar;
        r.c4 = 0.0;
        r.energy = mu / 2.0 * (I1 - 3.0) + 3.0 * mu * mu / (3.0 * K - 2.0 * mu) * (pw - 1.0);
    }
    r.tau[0] = sB * B[0] + sI;
    r.tau[1] = sB * B[1] + sI;
    r.tau[2] = sB * B[2] + sI;
    r.tau[3] = sB * B[3];
    r.tau[4] = sB * B[4];
    r.tau[5] = sB * B[5];
}

}  // namespace ewb
