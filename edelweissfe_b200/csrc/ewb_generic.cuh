// Generic (arbitrary connectivity) two-phase path — the device restatement of the reference's own
// data flow:   NIST.computeElements  -> VIJ values + per-element Pe     (nonlinearimplicitstatic.py:794-849)
//              CSRGenerator.updateCSR -> CSR data, summed in ascending COO order (csrgenerator.pyx:100-115)
//              P[el] += Pe ; F[el] += |Pe| -> gathered per node in ascending element order (:843-844)
// Deterministic (no atomics).  The intermediate is either the reference's V array (ewb_compute_elements_vij / ewb_update_csr,
// bit-for-bit the reference's data flow) or, inside ewb_assemble, the half-block scratch below; either way the path moves
// 2-3x the algorithmic bytes, so the fused BoxGen sweep (ewb_sweep.cuh) is the fast path for structured meshes.
#pragma once
#include "ewb_tile.cuh"

namespace ewb {

// Internal scratch of the arbitrary-mesh path ("half-block" layout): only the blocks the element loop computes are stored,
// K[a][a + d] for d = 0 .. NN/2 (node pairs in circulant order, Ke is symmetric), 9 doubles each, contiguous per node a:
//   S[e][a][d][3][3]  at  e * SE + a * SA + 10 d   (9 values + 1 pad: 16-byte aligned blocks),   SA = 10 (NN/2 + 1).
// 61 % of the VIJ bytes (reference layout: 9 NN^2 doubles per element), written as 16-byte stores of runs that are contiguous
// per thread; rowGatherHalfKernel reads block (a, b) either directly (d = b - a < nb(a)) or as the transpose of (b, a).
template <int NN>
struct HalfLayout {
    static constexpr int HALF = NN / 2, NB = HALF + 1;
    static constexpr int BS = 10;  // block stride: 9 values + 1 pad, so that every block is 16-byte aligned (5 x 128-bit accesses)
    static constexpr int SA = NB * BS;
    static constexpr int SE = NN * SA;
};

template <int NN>
struct HalfEmit {
    static constexpr bool HALF = true;
    double* S;   // element slice [SE]
    double* Pe;  // element slice [3 NN]
    __device__ __forceinline__ void residual(int a, const double P[3]) const {
#pragma unroll
        for (int i = 0; i < 3; ++i) Pe[3 * a + i] = P[i];
    }
    __device__ __forceinline__ void block(int, int, const double*) const {}
    template <int BLK>
    __device__ __forceinline__ void pass(int a, int ps, const double (&out)[BLK * 9 + 1]) const {
        using HL = HalfLayout<NN>;
        double2* dst = reinterpret_cast<double2*>(S + a * HL::SA + ps * (BLK * HL::BS));
        const int nBlk = min(BLK, HL::NB - ps * BLK);  // blocks d = ps*BLK .. ps*BLK + nBlk - 1
#pragma unroll
        for (int k = 0; k < BLK; ++k)
            if (k < nBlk) {
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[k * 5 + q] = make_double2(out[k * 9 + 2 * q], out[k * 9 + 2 * q + 1]);
                dst[k * 5 + 4] = make_double2(out[k * 9 + 8], 0.0);
            }
    }
};

template <int NN>
struct VijEmit {
    static constexpr bool HALF = false;
    template <int BLK>
    __device__ __forceinline__ void pass(int, int, const double (&)[BLK * 9 + 1]) const {}
    static constexpr int ND = 3 * NN;
    double* V;   // element slice [ND*ND] in the reference's VIJ layout, Ke row-major (element.py:318), or nullptr
    double* Pe;  // element slice [ND]
    __device__ __forceinline__ void residual(int a, const double P[3]) const {
#pragma unroll
        for (int i = 0; i < 3; ++i) Pe[3 * a + i] = P[i];
    }
    __device__ __forceinline__ void block(int a, int b, const double K[9]) const {
        if (V == nullptr) return;
        double* d = V + (3 * a) * ND + 3 * b;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) d[i * ND + j] = K[i * 3 + j];
    }
};

// Phase B of a linear-elastic 20-node element on the FP64 tensor pipe (one warp per element, half-block scratch).
// ncu of the scalar form (nodeRow): the LSU data pipe is 85 % busy — every 9 FMA of a block need 3 shared-memory loads of
// grad N_b, 62 % of all shared-memory wavefronts — while the FP64 pipe runs at 38 %.  As a (60 x 27) x (27 x 60) product on
// 8x8x4 tiles the operands are reused across the 9 component pairs: 7 loads per 9 DMMA (2304 FMA), 6 x fewer wavefronts.
// Node tiles ta <= tb of 8 rows / columns (20 nodes padded to 24, 27 Gauss points to 28); lane (r, q) ends with the two
// complete 3x3 blocks (a = 8 ta + r, b = 8 tb + 2 q + t) and stores those with a <= b into the circulant half layout
// (directly as K[a][a+d] if d = b - a < nb(a), else transposed under b).
__device__ __forceinline__ void dmmaG(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int NGP>
__device__ __forceinline__ void phaseBDmma20LE(const double* sm, int lane, const MatParams& mp, double* Se) {
    constexpr int NN = 20;
    using L = TileLayout<NN, NGP, MC_LE>;
    using HL = HalfLayout<NN>;
    constexpr int KSTEPS = (NGP + 3) / 4;
    const double* G = sm + L::OFF_G;
    const double* CO = sm + L::OFF_CO;
    const int r = lane >> 2, q = lane & 3;
#pragma unroll 1
    for (int ta = 0; ta < 3; ++ta) {
        const int a = 8 * ta + r;
        const bool aok = a < NN;
#pragma unroll 1
        for (int tb = ta; tb < 3; ++tb) {
            const int bn = 8 * tb + r;  // node of this lane's B-fragment column
            const bool bok = bn < NN;
            double c[3][3][2];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) c[i][j][0] = c[i][j][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const int gp = 4 * ks + q;
                const bool gok = gp < NGP;
                const double* Gg = G + (gok ? gp : 0) * L::GST;
                const double w = gok ? CO[gp * L::NCO] : 0.0;
                double A[3], B[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    A[i] = aok ? w * Gg[a * 3 + i] : 0.0;
                    B[i] = (bok && gok) ? Gg[bn * 3 + i] : 0.0;
                }
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) dmmaG(c[i][j], A[i], B[j]);
            }
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int b = 8 * tb + 2 * q + t;
                if (aok && b < NN && a <= b) {
                    double K[9];
                    const double tr = mp.G * (c[0][0][t] + c[1][1][t] + c[2][2][t]);
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) K[i * 3 + j] = mp.lambda * c[i][j][t] + mp.G * c[j][i][t] + (i == j ? tr : 0.0);
                    const int d = b - a;
                    const bool direct = d < HL::HALF + (a < HL::HALF ? 1 : 0);
                    double2* dst = reinterpret_cast<double2*>(Se + (direct ? a * HL::SA + d * HL::BS : b * HL::SA + (NN - d) * HL::BS));
                    double o[10];
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) o[i * 3 + j] = direct ? K[i * 3 + j] : K[j * 3 + i];
                    o[9] = 0.0;
#pragma unroll
                    for (int qq = 0; qq < 5; ++qq) dst[qq] = make_double2(o[2 * qq], o[2 * qq + 1]);
                }
            }
        }
    }
}

// T threads cooperate on one element (T >= NGP and T >= NN), E elements per CTA.
template <int NN, int NGP, int MC, bool TL, int T, int E, int BLK>
__global__ void __launch_bounds__(T* E, (NN == 20 ? 3 : 1)) computeElementsVijKernel(int64_t nEl, const int32_t* __restrict__ conn, const double* __restrict__ coords,
                                                                const double* __restrict__ U, const double* __restrict__ dU,
                                                                const double* __restrict__ stateRef, double* __restrict__ stateTemp,
                                                                double* __restrict__ V, double* __restrict__ Pe, MatParams mp, int* failFlag,
                                                                int halfScratch) {  // halfScratch: V is the internal half-block scratch (HalfLayout), else the reference's VIJ array
    using L = TileLayout<NN, NGP, MC>;
    extern __shared__ double smem[];
    const int el = threadIdx.x / T, t = threadIdx.x % T;
    const int64_t e = (int64_t)blockIdx.x * E + el;
    double* sm = smem + el * L::PER_EL;
    const bool active = e < nEl;
    if (active && t < NGP) {  // the Gauss-point state is needed after the first barrier: pull it into L2 now
        const double* sp = stateRef + e * NGP + t;
#pragma unroll
        for (int c = 0; c < 12 + matStateCount(MC); ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + c * (nEl * NGP)));
    }
    if (active) stageNodes<NN, NGP, MC, T>(sm, conn + e * NN, coords, U, dU, t);
    __syncthreads();
    if (active && t < NGP) {
        const int64_t cstride = nEl * NGP;
        const int64_t off = e * NGP + t;
        gaussPoint<NN, NGP, MC, TL>(sm, t, mp, stateRef + off, stateTemp + off, cstride, true, failFlag);
    }
    __syncthreads();
    // phase B work items: (node row a, pass of BLK circulant blocks), spread over the T threads of the element
    constexpr int NPASS = (NN / 2 + 1 + BLK - 1) / BLK;
    if constexpr (NN == 20 && MC == MC_LE && T == 32) {
        if (active && halfScratch && V != nullptr) {  // warp-uniform: one warp per element
            if (t < NN) {
                HalfEmit<NN> emit{V + e * (int64_t)HalfLayout<NN>::SE, Pe + e * (3 * NN)};
                nodeRow<NN, NGP, MC, BLK>(sm, t, mp, false, emit);  // residual row only
            }
            __syncwarp();
            phaseBDmma20LE<NGP>(sm, t, mp, V + e * (int64_t)HalfLayout<NN>::SE);
            return;
        }
    }
    if (active) {
#pragma unroll 1
        for (int it = t; it < NN * NPASS; it += T) {
            const int a = it % NN, ps = it / NN;
            if (halfScratch) {  // internal half-block scratch
                HalfEmit<NN> emit{V + e * (int64_t)HalfLayout<NN>::SE, Pe + e * (3 * NN)};
                nodeRow<NN, NGP, MC, BLK>(sm, a, mp, V != nullptr, emit, ps, ps + 1, ps == 0);
            } else {
                VijEmit<NN> emit{V ? V + e * (int64_t)(9 * NN * NN) : nullptr, Pe + e * (3 * NN)};
                nodeRow<NN, NGP, MC, BLK>(sm, a, mp, V != nullptr, emit, ps, ps + 1, ps == 0);
            }
        }
    }
}

// CSRGenerator.updateCSR as a gather: one thread per (node A, neighbour slot s) sums the 3x3 block
// over the elements incident to A in ascending element order == ascending COO index.
// Row of dof 3A+i starts at 9*adjPtr[A] + i*3*deg(A).
template <int NN>
__global__ void updateCsrKernel(int64_t nNode, const int64_t* __restrict__ adjPtr, const int32_t* __restrict__ adj,
                                const int64_t* __restrict__ incPtr, const int32_t* __restrict__ inc, const int32_t* __restrict__ conn,
                                const double* __restrict__ V, double* __restrict__ data, int64_t nSlots) {
    constexpr int ND = 3 * NN;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nSlots) return;
    // binary search the node owning global slot idx
    int64_t lo = 0, hi = nNode;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (adjPtr[mid] <= idx) lo = mid; else hi = mid;
    }
    const int64_t A = lo;
    const int64_t s = idx - adjPtr[A];
    const int64_t deg = adjPtr[A + 1] - adjPtr[A];
    const int32_t B = adj[idx];
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t k = incPtr[A]; k < incPtr[A + 1]; ++k) {
        const int32_t ea = inc[k];
        const int64_t e = ea / NN;
        const int a = ea % NN;
        int b = -1;
        for (int q = 0; q < NN; ++q)
            if (conn[e * NN + q] == B) { b = q; break; }
        if (b < 0) continue;
        // row dof = dof[p % n] = 3a+i, col dof = dof[p / n] = 3b+j  ->  p = (3b+j)*n + 3a+i   (dofmanager.py:552-553)
        const double* Ve = V + e * (int64_t)(ND * ND);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[i * 3 + j] += Ve[(3 * b + j) * ND + 3 * a + i];
    }
    const int64_t base = 9 * adjPtr[A];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) data[base + i * 3 * deg + 3 * s + j] = acc[i * 3 + j];
}

// Precomputed element-to-CSR-slot map of the row gather: for incidence k = (node A, element e, local node a) and every local
// node b of e, the position of conn[e][b] in A's sorted neighbour list (node degree <= 255: one byte).  Built once per plan.
template <int NN>
__global__ void gatherSlotKernel(int64_t nNode, const int64_t* __restrict__ adjPtr, const int32_t* __restrict__ adj,
                                 const int64_t* __restrict__ incPtr, const int32_t* __restrict__ inc, const int32_t* __restrict__ conn,
                                 unsigned char* __restrict__ slotTab) {
    const int64_t A = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (A >= nNode || lane >= NN) return;
    const int64_t s0 = adjPtr[A];
    const int deg = (int)(adjPtr[A + 1] - s0);
    const int32_t* nb = adj + s0;
    for (int64_t k = incPtr[A]; k < incPtr[A + 1]; ++k) {
        const int64_t e = inc[k] / NN;
        const int32_t B = conn[e * NN + lane];
        int lo = 0, hi = deg - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (nb[mid] < B) lo = mid + 1; else hi = mid;
        }
        slotTab[k * NN + lane] = (unsigned char)lo;
    }
}

// CSRGenerator.updateCSR on the half-block scratch (HalfLayout): one warp per node A, lane b < NN owns the block (a, b) of the
// incident element (e, a) — stored as K[a][a+d] if d = b - a (mod NN) < nb(a), else as the transpose of K[b][b+d'] — and adds it
// to A's three CSR rows in shared memory.  Elements are visited in ascending order (== ascending COO index, the reference's
// summation order); the next element's block travels while the current one is added.
template <int NN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 4) rowGatherHalfKernel(int64_t nNode, const int64_t* __restrict__ adjPtr, const int32_t* __restrict__ adj,
                                                                  const int64_t* __restrict__ incPtr, const int32_t* __restrict__ inc,
                                                                  const int32_t* __restrict__ conn, const double* __restrict__ S, double* __restrict__ data,
                                                                  int maxDeg, const int32_t* __restrict__ order,
                                                                  const unsigned char* __restrict__ slotTab) {
    using HL = HalfLayout<NN>;
    static_assert(NN <= 32, "one lane per element node");
    extern __shared__ double rowBufAll[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t slotA = (int64_t)blockIdx.x * WARPS + warp;
    if (slotA >= nNode) return;
    const int64_t A = order ? order[slotA] : slotA;  // visiting order only (locality of the scratch reads)
    double* buf = rowBufAll + (size_t)warp * 9 * maxDeg;
    const int64_t s0 = adjPtr[A];
    const int deg = (int)(adjPtr[A + 1] - s0);
    const int rowLen = 3 * deg;
    for (int i = lane; i < 3 * rowLen; i += 32) buf[i] = 0.0;
    const int32_t* nb = adj + s0;
    const int64_t k0 = incPtr[A], k1 = incPtr[A + 1];
    constexpr int G = NN <= 8 ? 32 / NN : 1;  // incident elements per batch
    const int g = lane / NN, lb = lane - g * NN;
    const bool laneOn = lane < G * NN;
    // fetch: this lane's block of incident element k (registers) and its slot in A's sorted neighbour list
    auto fetch = [&](int64_t k, double2 (&v)[5], int& slot, bool& direct) {
        const int32_t ea = inc[k];
        const int64_t e = ea / NN;
        const int a = ea % NN;
        const int b = lb;
        int d = b - a;
        if (d < 0) d += NN;
        direct = d < HL::HALF + (a < HL::HALF ? 1 : 0);
        const double2* src = reinterpret_cast<const double2*>(S + e * (int64_t)HL::SE + (direct ? a * HL::SA + d * HL::BS : b * HL::SA + (NN - d) * HL::BS));
#pragma unroll
        for (int i = 0; i < 5; ++i) v[i] = src[i];
        if (slotTab) {
            slot = slotTab[k * NN + b];
        } else {
            const int32_t B = conn[e * NN + b];
            int lo = 0, hi = deg - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (nb[mid] < B) lo = mid + 1; else hi = mid;
            }
            slot = lo;
        }
    };
    double2 v[5], w[5];
    int slot = 0, slotN = 0;
    bool direct = true, directN = true;
    // 8-node elements: four incident elements per batch (lane group g = lane / NN takes element kb + g), so that all 32 lanes
    // fetch; the groups then add one after the other (ascending element order, and two elements may hit the same CSR entry).
    bool have = laneOn && k0 + g < k1;
    if (have) fetch(k0 + g, v, slot, direct);
    __syncwarp();
    for (int64_t kb = k0; kb < k1; kb += G) {
        const bool haveN = laneOn && kb + G + g < k1;
        if (haveN) fetch(kb + G + g, w, slotN, directN);
#pragma unroll
        for (int gg = 0; gg < G; ++gg) {
            if (have && g == gg) {
                double* dst = buf + 3 * slot;
                const double x[9] = {v[0].x, v[0].y, v[1].x, v[1].y, v[2].x, v[2].y, v[3].x, v[3].y, v[4].x};
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) dst[i * rowLen + j] += direct ? x[i * 3 + j] : x[j * 3 + i];
            }
            __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) v[i] = w[i];
        slot = slotN;
        direct = directN;
        have = haveN;
    }
    double* out = data + 9 * s0;
    for (int i = lane; i < 3 * rowLen; i += 32) out[i] = buf[i];
}

// P[el] += Pe ; F[el] += |Pe| : per node, ascending element order.
template <int NN>
__global__ void gatherResidualKernel(int64_t nNode, const int64_t* __restrict__ incPtr, const int32_t* __restrict__ inc,
                                     const double* __restrict__ Pe, double* __restrict__ P, double* __restrict__ F, int accumulate) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 3 * nNode) return;
    const int64_t A = idx / 3;
    const int c = idx % 3;
    double p = accumulate ? P[idx] : 0.0, f = accumulate ? F[idx] : 0.0;
    for (int64_t k = incPtr[A]; k < incPtr[A + 1]; ++k) {
        const int32_t ea = inc[k];
        const double v = Pe[(int64_t)(ea / NN) * (3 * NN) + 3 * (ea % NN) + c];
        p += v;
        f += fabs(v);
    }
    P[idx] = p;
    F[idx] = f;
}

// Shape function of node A as the reference's computeNOperator evaluates it: its node table has xi and eta
// swapped relative to the derivative tables (displacementelement/_elementcomputationmatrices.py:164-211).
template <int NN, int A>
__device__ __forceinline__ double shapeFnBodyForce(double xi, double eta, double zeta) {
    constexpr int a = NodeLC<NN>::eta(A), b = NodeLC<NN>::xi(A), c = NodeLC<NN>::zeta(A);  // swapped on purpose
    const double fx = 1.0 + a * xi, fe = 1.0 + b * eta, fz = 1.0 + c * zeta;
    if constexpr (NN == 8) return 0.125 * fx * fe * fz;
    else if constexpr (a == 0) return 0.25 * (1.0 - xi * xi) * fe * fz;
    else if constexpr (b == 0) return 0.25 * fx * (1.0 - eta * eta) * fz;
    else if constexpr (c == 0) return 0.25 * fx * fe * (1.0 - zeta * zeta);
    else return 0.125 * fx * fe * fz * (a * xi + b * eta + c * zeta - 2.0);
}

// computeBodyForce for every element (elements/displacementelement/element.py:348-371):
// Pe[3a+i] = sum_gp N_a load_i detJ w.  One thread per element; Pe -> per-element slab, gathered per node afterwards.
template <int NN, int NGP>
__global__ void bodyForceKernel(int64_t nEl, const int32_t* __restrict__ conn, const double* __restrict__ coords, double lx, double ly, double lz,
                                double* __restrict__ Pe) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nEl) return;
    double X[NN * 3];
#pragma unroll
    for (int a = 0; a < NN; ++a) {
        const int64_t n = conn[e * NN + a];
#pragma unroll
        for (int c = 0; c < 3; ++c) X[a * 3 + c] = coords[3 * n + c];
    }
    double w8[NN];
#pragma unroll
    for (int a = 0; a < NN; ++a) w8[a] = 0.0;
    for (int gp = 0; gp < NGP; ++gp) {
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
        const double wd = w * det3(Jm);
        forNodes<NN>([&](auto ic) {
            constexpr int a = decltype(ic)::value;
            w8[a] = fma(shapeFnBodyForce<NN, a>(xi, eta, zeta), wd, w8[a]);
        });
    }
#pragma unroll
    for (int a = 0; a < NN; ++a) {
        Pe[e * (3 * NN) + 3 * a + 0] = w8[a] * lx;
        Pe[e * (3 * NN) + 3 * a + 1] = w8[a] * ly;
        Pe[e * (3 * NN) + 3 * a + 2] = w8[a] * lz;
    }
}

// PExt[el] += Pe per node in ascending element order (nonlinearimplicitstatic.py:545-553)
template <int NN>
__global__ void gatherLoadKernel(int64_t nNode, const int64_t* __restrict__ incPtr, const int32_t* __restrict__ inc, const double* __restrict__ Pe,
                                 double* __restrict__ PExt) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 3 * nNode) return;
    const int64_t A = idx / 3;
    const int c = idx % 3;
    double p = PExt[idx];
    for (int64_t k = incPtr[A]; k < incPtr[A + 1]; ++k) {
        const int32_t ea = inc[k];
        p += Pe[(int64_t)(ea / NN) * (3 * NN) + 3 * (ea % NN) + c];
    }
    PExt[idx] = p;
}

}  // namespace ewb
