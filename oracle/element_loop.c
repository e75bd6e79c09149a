/*
 * CPU ORACLE (C) — TEST / BENCH INFRASTRUCTURE ONLY, never linked into the product.
 *
 * Plain-C restatement of the reference's element loop in its own *dense* formulation
 * (explicit 6 x nDof B operator, Bt C B with a full 6x6 tangent, the 3^4 d tau / d F tensor
 * for the total-Lagrange element), parallelised over elements with OpenMP exactly where the
 * reference's own parallel solver is (prange over elements, serial P/F scatter, serial CSR
 * update).  Paths below /root/reference/edelweissfe/ :
 *
 *   solvers/nonlinearimplicitstaticparallelmk2.pyx:111-198   loop structure (prange :157-160, scatter :187-193)
 *   elements/displacementelement/element.py:290-346          small-strain element
 *   elements/displacementelement/_elementcomputationmatrices.py:306-366, 700-817   J, B
 *   elements/displacementtlelement/element.py:346-427        TL element, hyperelastic branch
 *   materials/linearelastic/linearelastic.py:95-117,185-210
 *   materials/vonmises/vonmises.py:186-254
 *   materials/neohooke/neohookepencegouformulation{a,b,c}.py:130-145
 *   numerics/csrgenerator.pyx:100-115                        updateCSR (sequential, ascending COO index)
 *
 * Parity status: PINNED — tests/test_oracle_golden.py checks this file against tests/golden/*.npz
 * (outputs of the unmodified reference).
 *
 * State layout here is the reference's own: state[e][gp][12+m].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { EL_C3D8 = 0, EL_C3D20 = 1, EL_C3D8TL = 2, EL_C3D8R = 3, EL_C3D8E = 4, EL_C3D20R = 5 }; /* 3..5: integration variants, library.py:228-259, 276-291 */
enum { MAT_LE = 0, MAT_VM = 1, MAT_NHA = 2, MAT_NHB = 3, MAT_NHC = 4 };

static const int OFF8[8][3] = {{0, 0, 0}, {0, 0, 1}, {1, 0, 1}, {1, 0, 0}, {0, 1, 0}, {0, 1, 1}, {1, 1, 1}, {1, 1, 0}};
static const int OFF20[20][3] = {{0, 0, 0}, {0, 0, 2}, {2, 0, 2}, {2, 0, 0}, {0, 2, 0}, {0, 2, 2}, {2, 2, 2}, {2, 2, 0}, {0, 0, 1}, {1, 0, 2},
                                 {2, 0, 1}, {1, 0, 0}, {0, 2, 1}, {1, 2, 2}, {2, 2, 1}, {1, 2, 0}, {0, 1, 0}, {0, 1, 2}, {2, 1, 2}, {2, 1, 0}};

/* local (xi, eta, zeta) of node a: eta runs along BoxGen x, xi along y, zeta along z */
static void node_lc(int nn, int a, double* xi, double* eta, double* zeta) {
    if (nn == 8) {
        *eta = 2.0 * OFF8[a][0] - 1.0; *xi = 2.0 * OFF8[a][1] - 1.0; *zeta = 2.0 * OFF8[a][2] - 1.0;
    } else {
        *eta = OFF20[a][0] - 1.0; *xi = OFF20[a][1] - 1.0; *zeta = OFF20[a][2] - 1.0;
    }
}

/* dN[r][a], rows (d/d eta, d/d xi, d/d zeta) as in the reference tables */
static void shape_derivs(int nn, double X, double E, double Z, double dN[3][20]) {
    for (int n = 0; n < nn; ++n) {
        double a, b, c;
        node_lc(nn, n, &a, &b, &c);
        const double fx = 1 + a * X, fe = 1 + b * E, fz = 1 + c * Z;
        double dxi, deta, dzeta;
        if (nn == 8) {
            dxi = a * fe * fz / 8; deta = b * fx * fz / 8; dzeta = c * fx * fe / 8;
        } else if (a == 0) {
            dxi = -2 * X * fe * fz / 4; deta = (1 - X * X) * b * fz / 4; dzeta = (1 - X * X) * fe * c / 4;
        } else if (b == 0) {
            dxi = a * (1 - E * E) * fz / 4; deta = -2 * E * fx * fz / 4; dzeta = fx * (1 - E * E) * c / 4;
        } else if (c == 0) {
            dxi = a * fe * (1 - Z * Z) / 4; deta = fx * b * (1 - Z * Z) / 4; dzeta = -2 * Z * fx * fe / 4;
        } else {
            const double s = a * X + b * E + c * Z - 2;
            dxi = a * fe * fz * (s + fx) / 8; deta = b * fx * fz * (s + fe) / 8; dzeta = c * fx * fe * (s + fz) / 8;
        }
        dN[0][n] = deta; dN[1][n] = dxi; dN[2][n] = dzeta;
    }
}

static void gauss(int ngp, int g, double* xi, double* eta, double* zeta, double* w) {
    if (ngp == 1) { /* elements/library.py:228-243: reduced integration */
        *xi = *eta = *zeta = 0.0; *w = 8.0;
    } else if (ngp == 8) { /* elements/library.py:38-39, 212-227 */
        const double q = 1.0 / sqrt(3.0);
        static const double s8[4] = {-1, 1, 1, -1}, t8[4] = {-1, -1, 1, 1};
        *xi = (g < 4 ? -q : q); *eta = q * t8[g % 4]; *zeta = q * s8[g % 4]; *w = 1.0;
    } else { /* elements/library.py:40-47, 260-275 */
        const double r = sqrt(0.6);
        const int ix = g / 9, ie = (g % 9) / 3, iz = g % 3;
        *xi = r * (ix - 1); *eta = r * (ie - 1); *zeta = r * (iz - 1);
        const double w5 = 5.0 / 9.0, w8 = 8.0 / 9.0;
        *w = (ix == 1 ? w8 : w5) * (ie == 1 ? w8 : w5) * (iz == 1 ? w8 : w5);
    }
}

static double det3(const double m[9]) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
static void inv3(const double m[9], double r[9]) {
    const double id = 1.0 / det3(m);
    r[0] = (m[4] * m[8] - m[5] * m[7]) * id; r[1] = (m[2] * m[7] - m[1] * m[8]) * id; r[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    r[3] = (m[5] * m[6] - m[3] * m[8]) * id; r[4] = (m[0] * m[8] - m[2] * m[6]) * id; r[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    r[6] = (m[3] * m[7] - m[4] * m[6]) * id; r[7] = (m[1] * m[6] - m[0] * m[7]) * id; r[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

static void elasticity_matrix(double E, double v, double C[36]) {
    memset(C, 0, 36 * sizeof(double));
    const double f = E / ((1 + v) * (1 - 2 * v));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i * 6 + j] = f * (i == j ? (1 - v) : v);
    for (int i = 3; i < 6; ++i) C[i * 6 + i] = f * (1 - 2 * v) / 2;
}

/* vonmises.py:186-254; returns 1 if the reference would raise CutbackRequest */
static int von_mises(const double* props, double s[6], double C[36], const double de[6], double* kappa) {
    const double E = props[0], v = props[1], fy0 = props[2], HLin = props[3], dfy = props[4], delta = props[5];
    const double G = E / (2 * (1.0 + v));
    double Ei[36];
    elasticity_matrix(E, v, Ei);
    memcpy(C, Ei, sizeof(Ei));
    double nrm = 0;
    for (int i = 0; i < 6; ++i) nrm += de[i] * de[i];
    if (sqrt(nrm) < 1e-14) return 0;
    double t[6], d[6];
    for (int i = 0; i < 6; ++i) {
        t[i] = s[i];
        for (int j = 0; j < 6; ++j) t[i] += de[j] * Ei[j * 6 + i];
    }
    const double pm = (t[0] + t[1] + t[2]) / 3.0;
    /* IDev @ trial */
    d[0] = 2. / 3 * t[0] - 1. / 3 * t[1] - 1. / 3 * t[2];
    d[1] = -1. / 3 * t[0] + 2. / 3 * t[1] - 1. / 3 * t[2];
    d[2] = -1. / 3 * t[0] - 1. / 3 * t[1] + 2. / 3 * t[2];
    (void)pm;
    d[3] = t[3]; d[4] = t[4]; d[5] = t[5];
    const double devn = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + 2 * (d[3] * d[3] + d[4] * d[4] + d[5] * d[5]));
    const double k0 = *kappa;
#define FY(k) (fy0 + HLin * (k) + dfy * (1.0 - exp(-delta * (k))))
#define DFY(k) (HLin + dfy * delta * exp(-delta * (k)))
    if (devn - sqrt(2. / 3) * FY(k0) > 0.0) {
        double dk = 0;
        int counter = 0, failed = 0;
        for (;;) {
            const double R = devn - sqrt(6.) * G * dk - sqrt(2. / 3) * FY(k0 + dk);
            if (!(fabs(R) > 1e-12)) break;
            if (counter == 15) { failed = 1; break; }
            const double dR = -sqrt(6.) * G - sqrt(2. / 3) * DFY(k0 + dk);
            dk -= R / dR;
            ++counter;
        }
        const double dLambda = sqrt(3. / 2) * dk;
        *kappa = k0 + dk;
        double n[6];
        for (int i = 0; i < 6; ++i) { n[i] = d[i] / devn; s[i] = t[i] - 2.0 * G * dLambda * n[i]; }
        const double fn = 2.0 * G * (1.0 / (1.0 + DFY(*kappa) / (3.0 * G)) - 2.0 * G * dLambda / devn);
        const double fd = 4.0 * G * G * dLambda / devn;
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
                double idh = 0;
                if (i < 3 && j < 3) idh = (i == j ? 2. / 3 : -1. / 3);
                else if (i == j) idh = 0.5;
                C[i * 6 + j] = Ei[i * 6 + j] - fn * n[i] * n[j] - fd * idh;
            }
        return failed;
    }
    for (int i = 0; i < 6; ++i) s[i] = t[i];
    return 0;
#undef FY
#undef DFY
}

/* neohookepencegouformulation{a,b,c}.py:130-145: tau (3x3), A = d tau/d F (3^4), energy */
static void neo_hooke(int kind, const double* props, const double F[9], double tau[9], double A[81], double* energy) {
    const double mu = props[0], K = props[1];
    double iF[9], B[9];
    inv3(F, iF);
    const double J = det3(F);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            B[i * 3 + j] = 0;
            for (int k = 0; k < 3; ++k) B[i * 3 + j] += F[i * 3 + k] * F[j * 3 + k];
        }
    const double trB = B[0] + B[4] + B[8];
#define IDX(i, j, k, l) ((((i)*3 + (j)) * 3 + (k)) * 3 + (l))
#define DL(i, j) ((i) == (j) ? 1.0 : 0.0)
    if (kind == MAT_NHA) {
        const double lamBar = (K - 2. / 3 * mu) * (J * J - J) - mu, muBar = (K - 2. / 3 * mu) * (2 * J * J - J);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) tau[i * 3 + j] = mu * B[i * 3 + j] + lamBar * DL(i, j);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l)
            A[IDX(i, j, k, l)] = mu * (DL(i, k) * F[j * 3 + l] + F[i * 3 + l] * DL(j, k)) + muBar * DL(i, j) * iF[l * 3 + k];
        *energy = mu / 2 * (trB - 3) + (K / 2 - mu / 3) * (J - 1) * (J - 1) - mu * log(J);
    } else if (kind == MAT_NHB) {
        const double J23 = pow(J, 2. / 3), lamHat = K / 2 * (J * J + 1 / (J * J)), muBar = mu / (3 * J23);
        const double lamBar = K / 4 * (J * J - 1 / (J * J)) - muBar * trB;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) tau[i * 3 + j] = mu / J23 * B[i * 3 + j] + lamBar * DL(i, j);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l)
            A[IDX(i, j, k, l)] = 3 * muBar * (DL(i, k) * F[j * 3 + l] + F[i * 3 + l] * DL(j, k)) - 2 * muBar * B[i * 3 + j] * iF[l * 3 + k] +
                                 (lamHat + 2. / 3 * trB * muBar) * DL(i, j) * iF[l * 3 + k] - 2 * muBar * DL(i, j) * F[k * 3 + l];
        *energy = mu / 2 * (trB / J23 - 3) + K / 8 * (J * J + 1 / (J * J) - 2);
    } else {
        const double I1 = F[0] + F[4] + F[8]; /* sic: trace(F), ...c.py:134 */
        const double pw = pow(J, 2. / 3 - K / mu), muBar = mu * pw, lamBar = (K / mu - 2. / 3) * muBar;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) tau[i * 3 + j] = mu * B[i * 3 + j] - muBar * DL(i, j);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l)
            A[IDX(i, j, k, l)] = mu * (DL(i, k) * F[j * 3 + l] + F[i * 3 + l] * DL(j, k)) + lamBar * DL(i, j) * iF[l * 3 + k];
        *energy = mu / 2 * (I1 - 3) + 3 * mu * mu / (3 * K - 2 * mu) * (pw - 1);
    }
}

/* one element: Ke (row-major nd x nd, accumulated into a zeroed slice), Pe, stateTemp; returns failure flag */
static int compute_element(int eltype, int material, const double* props, int nn, int ngp, int nstate, const double* X /*[nn][3]*/,
                           const double* U, const double* dU, const double* sref, double* stemp, double* Ke, double* Pe) {
    const int nd = 3 * nn;
    int failed = 0;
    memset(Ke, 0, (size_t)nd * nd * sizeof(double));
    memset(Pe, 0, (size_t)nd * sizeof(double));
    memcpy(stemp, sref, (size_t)ngp * nstate * sizeof(double)); /* element.py:320 */
    double dN[3][20], gN[3][20], B[6][60], CB[6][60];
    for (int g = 0; g < ngp; ++g) {
        double xi, eta, zeta, w;
        gauss(ngp, g, &xi, &eta, &zeta, &w);
        shape_derivs(nn, xi, eta, zeta, dN);
        double Jm[9] = {0}, iJ[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                for (int a = 0; a < nn; ++a) Jm[r * 3 + c] += dN[r][a] * X[a * 3 + c];
        const double detJ = det3(Jm);
        inv3(Jm, iJ);
        for (int c = 0; c < 3; ++c)
            for (int a = 0; a < nn; ++a) gN[c][a] = iJ[c * 3] * dN[0][a] + iJ[c * 3 + 1] * dN[1][a] + iJ[c * 3 + 2] * dN[2][a];
        double* st = stemp + (size_t)g * nstate;
        const double scale = detJ * w;
        if (eltype != EL_C3D8TL) {
            memset(B, 0, sizeof(B));
            for (int a = 0; a < nn; ++a) { /* Voigt rows 11,22,33,12,13,23 */
                B[0][3 * a] = gN[0][a]; B[1][3 * a + 1] = gN[1][a]; B[2][3 * a + 2] = gN[2][a];
                B[3][3 * a] = gN[1][a]; B[3][3 * a + 1] = gN[0][a];
                B[4][3 * a] = gN[2][a]; B[4][3 * a + 2] = gN[0][a];
                B[5][3 * a + 1] = gN[2][a]; B[5][3 * a + 2] = gN[1][a];
            }
            double de[6] = {0}, C[36];
            for (int v = 0; v < 6; ++v)
                for (int q = 0; q < nd; ++q) de[v] += B[v][q] * dU[q];
            if (material == MAT_LE) {
                elasticity_matrix(props[0], props[1], C);
                for (int i = 0; i < 6; ++i)
                    for (int j = 0; j < 6; ++j) st[i] += C[i * 6 + j] * de[j];
            } else {
                failed |= von_mises(props, st, C, de, st + 12);
            }
            for (int v = 0; v < 6; ++v)
                for (int q = 0; q < nd; ++q) {
                    double acc = 0;
                    for (int u = 0; u < 6; ++u) acc += C[v * 6 + u] * B[u][q];
                    CB[v][q] = acc;
                }
            for (int p = 0; p < nd; ++p)
                for (int q = 0; q < nd; ++q) {
                    double acc = 0;
                    for (int v = 0; v < 6; ++v) acc += B[v][p] * CB[v][q];
                    Ke[p * nd + q] += acc * scale;
                }
            for (int p = 0; p < nd; ++p) {
                double acc = 0;
                for (int v = 0; v < 6; ++v) acc += B[v][p] * st[v];
                Pe[p] -= acc * scale;
            }
            for (int v = 0; v < 6; ++v) st[6 + v] += de[v];
        } else if (material == MAT_LE || material == MAT_VM) {
            /* total Lagrange, non-hyperelastic branch: B^T C B + Hgeo (displacementtlelement/element.py:415-425, :48-73);
             * E_old is the strain part of the accepted state (Voigt, doubled shear), see oracle/port.py */
            double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            for (int a = 0; a < nn; ++a)
                for (int i = 0; i < 3; ++i)
                    for (int c = 0; c < 3; ++c) F[i * 3 + c] += U[3 * a + i] * gN[c][a];
            double H[9], Eg[9];
            for (int i = 0; i < 9; ++i) H[i] = F[i] - (i % 4 == 0 ? 1.0 : 0.0);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    double hh = 0;
                    for (int k = 0; k < 3; ++k) hh += H[k * 3 + i] * H[k * 3 + j];
                    Eg[i * 3 + j] = 0.5 * (H[i * 3 + j] + H[j * 3 + i] + hh);
                }
            const double Ev[6] = {Eg[0], Eg[4], Eg[8], 2 * Eg[1], 2 * Eg[5], 2 * Eg[2]}; /* 11,22,33,12,23,13 */
            double de[6], C[36];
            for (int v = 0; v < 6; ++v) de[v] = Ev[v] - st[6 + v];
            if (material == MAT_LE) {
                elasticity_matrix(props[0], props[1], C);
                for (int i = 0; i < 6; ++i)
                    for (int j = 0; j < 6; ++j) st[i] += C[i * 6 + j] * de[j];
            } else {
                failed |= von_mises(props, st, C, de, st + 12);
            }
            for (int a = 0; a < nn; ++a) /* _B03D, _elementcomputationmatrices.py:867-916 */
                for (int k = 0; k < 3; ++k) {
                    const double fx = F[k * 3], fy = F[k * 3 + 1], fz = F[k * 3 + 2];
                    B[0][3 * a + k] = gN[0][a] * fx;
                    B[1][3 * a + k] = gN[1][a] * fy;
                    B[2][3 * a + k] = gN[2][a] * fz;
                    B[3][3 * a + k] = gN[0][a] * fy + gN[1][a] * fx;
                    B[4][3 * a + k] = gN[1][a] * fz + gN[2][a] * fy;
                    B[5][3 * a + k] = gN[0][a] * fz + gN[2][a] * fx;
                }
            for (int v = 0; v < 6; ++v)
                for (int q = 0; q < nd; ++q) {
                    double acc = 0;
                    for (int u = 0; u < 6; ++u) acc += C[v * 6 + u] * B[u][q];
                    CB[v][q] = acc;
                }
            const double S[9] = {st[0], st[3], st[5], st[3], st[1], st[4], st[5], st[4], st[2]};
            for (int p = 0; p < nd; ++p)
                for (int q = 0; q < nd; ++q) {
                    double acc = 0;
                    for (int v = 0; v < 6; ++v) acc += B[v][p] * CB[v][q];
                    if (p % 3 == q % 3) { /* Hgeo: (grad N_a . S . grad N_b) I */
                        const int a = p / 3, b = q / 3;
                        for (int i = 0; i < 3; ++i)
                            for (int j = 0; j < 3; ++j) acc += gN[i][a] * S[i * 3 + j] * gN[j][b];
                    }
                    Ke[p * nd + q] += acc * scale;
                }
            for (int p = 0; p < nd; ++p) {
                double acc = 0;
                for (int v = 0; v < 6; ++v) acc += B[v][p] * st[v];
                Pe[p] -= acc * scale;
            }
            for (int v = 0; v < 6; ++v) st[6 + v] += de[v];
        } else {
            double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, iF[9], tau[9], A[81], energy;
            for (int a = 0; a < nn; ++a)
                for (int i = 0; i < 3; ++i)
                    for (int c = 0; c < 3; ++c) F[i * 3 + c] += U[3 * a + i] * gN[c][a];
            double H[9], Eg[9];
            for (int i = 0; i < 9; ++i) H[i] = F[i] - (i % 4 == 0 ? 1.0 : 0.0);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    double hh = 0;
                    for (int k = 0; k < 3; ++k) hh += H[k * 3 + i] * H[k * 3 + j];
                    Eg[i * 3 + j] = 0.5 * (H[i * 3 + j] + H[j * 3 + i] + hh);
                }
            inv3(F, iF);
            neo_hooke(material, props, F, tau, A, &energy);
            double NAi[20][3], PK1[9];
            for (int a = 0; a < nn; ++a)
                for (int m = 0; m < 3; ++m) NAi[a][m] = gN[0][a] * iF[m] + gN[1][a] * iF[3 + m] + gN[2][a] * iF[6 + m];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    PK1[i * 3 + j] = 0;
                    for (int k = 0; k < 3; ++k) PK1[i * 3 + j] += iF[i * 3 + k] * tau[k * 3 + j];
                }
            /* Hk[a j b k] = NAi[a,i] A[i,j,k,l] gN[l,b] - NAi[a,k] NAi[b,i] T[i,j]  (element.py:408-412) */
            for (int a = 0; a < nn; ++a)
                for (int j = 0; j < 3; ++j) {
                    double tmp[3][3]; /* tmp[k][l] = sum_i NAi[a,i] A[i,j,k,l] */
                    for (int k = 0; k < 3; ++k)
                        for (int l = 0; l < 3; ++l) tmp[k][l] = NAi[a][0] * A[IDX(0, j, k, l)] + NAi[a][1] * A[IDX(1, j, k, l)] + NAi[a][2] * A[IDX(2, j, k, l)];
                    for (int b = 0; b < nn; ++b) {
                        const double tn = NAi[b][0] * tau[0 * 3 + j] + NAi[b][1] * tau[1 * 3 + j] + NAi[b][2] * tau[2 * 3 + j];
                        for (int k = 0; k < 3; ++k) {
                            const double m1 = tmp[k][0] * gN[0][b] + tmp[k][1] * gN[1][b] + tmp[k][2] * gN[2][b];
                            Ke[(3 * a + j) * nd + 3 * b + k] += (m1 - NAi[a][k] * tn) * scale;
                        }
                    }
                }
            for (int a = 0; a < nn; ++a)
                for (int k = 0; k < 3; ++k) Pe[3 * a + k] -= (gN[0][a] * PK1[k] + gN[1][a] * PK1[3 + k] + gN[2][a] * PK1[6 + k]) * scale;
            st[0] = tau[0]; st[1] = tau[4]; st[2] = tau[8]; st[3] = tau[1]; st[4] = tau[5]; st[5] = tau[2];
            st[6] = Eg[0]; st[7] = Eg[4]; st[8] = Eg[8]; st[9] = 2 * Eg[1]; st[10] = 2 * Eg[5]; st[11] = 2 * Eg[2];
            st[12] = energy;
        }
    }
    return failed;
}
#undef IDX
#undef DL

int ewo_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* NISTParallel.computeElements loop body (mk2.pyx:157-184): every element writes its own VIJ slice and Pe slab */
int ewo_compute_elements(int eltype, int material, const double* props, int64_t nEl, const int32_t* conn, const double* coords, const double* U,
                         const double* dU, const double* stateRef, double* stateTemp, double* V, double* Pe, int nthreads) {
    const int nn = (eltype == EL_C3D20 || eltype == EL_C3D20R) ? 20 : 8;
    const int ngp = (eltype == EL_C3D20 || eltype == EL_C3D8E) ? 27 : (eltype == EL_C3D8R ? 1 : 8);
    const int nstate = 12 + (material == MAT_LE ? 0 : 1), nd = 3 * nn;
    int failedAny = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64) reduction(| : failedAny)
    for (int64_t e = 0; e < nEl; ++e) {
        double X[60], Ue[60], dUe[60];
        for (int a = 0; a < nn; ++a) {
            const int64_t n = conn[e * nn + a];
            for (int c = 0; c < 3; ++c) {
                X[3 * a + c] = coords[3 * n + c];
                Ue[3 * a + c] = U[3 * n + c];
                dUe[3 * a + c] = dU[3 * n + c];
            }
        }
        failedAny |= compute_element(eltype, material, props, nn, ngp, nstate, X, Ue, dUe, stateRef + (size_t)e * ngp * nstate,
                                     stateTemp + (size_t)e * ngp * nstate, V + (size_t)e * nd * nd, Pe + (size_t)e * nd);
    }
    return failedAny;
}

/* P[el] += Pe ; F[el] += |Pe|, serial in element order (mk2.pyx:187-193) */
void ewo_scatter_pf(int64_t nEl, int nn, const int32_t* conn, const double* Pe, double* P, double* F) {
    const int nd = 3 * nn;
    for (int64_t e = 0; e < nEl; ++e)
        for (int a = 0; a < nn; ++a)
            for (int c = 0; c < 3; ++c) {
                const double v = Pe[e * nd + 3 * a + c];
                const int64_t dof = 3 * (int64_t)conn[e * nn + a] + c;
                P[dof] += v;
                F[dof] += fabs(v);
            }
}

/* CSRGenerator.updateCSR (csrgenerator.pyx:100-115): data[:] = 0; data[x[p]] += V[p], sequential */
void ewo_update_csr(int64_t nCoo, const int32_t* x, const double* V, double* data, int64_t nnz) {
    memset(data, 0, (size_t)nnz * sizeof(double));
    for (int64_t p = 0; p < nCoo; ++p) data[x[p]] += V[p];
}
