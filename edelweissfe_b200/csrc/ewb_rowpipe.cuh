// Row-pipelined gather sweep for BoxGen 8-node hexahedra — the second-generation fused kernel.
//
// Same contract as sweepKernel (one launch = NIST.computeElements + CSRGenerator.updateCSR,
// solvers/nonlinearimplicitstatic.py:794-849, numerics/csrgenerator.pyx:100-115; every CSR value, P, F and state value
// written exactly once, no atomics, fixed summation order), different dataflow:
//
//   * a CTA owns a (rows x TZ) tile of node columns and a chunk of node planes in x.  It sweeps the element planes along
//     x and, inside a plane, the element ROWS along y (TZ + 1 elements per row, incl. the halo elements);
//   * three warp roles form a software pipeline over element rows, linked by mbarrier full/empty rings:
//       P (producer) warps : phase A — kinematics, constitutive update, state write-back; publish the Gauss-point
//                            records of a row (ring of REC_STAGES rows),
//       T (tensor) warps   : phase B — stiffness blocks of one element per warp on the FP64 tensor pipe (DMMA), stored
//                            with plain 128-bit stores into the element's slot [lane][2 blocks][3x3] (ring of 3 rows),
//       G (gather) warps   : one node column at a time — sum the (up to) four elements of the row pair that touch the
//                            column, in fixed order, in registers; add the carry of the previous element plane; store
//                            the finished CSR sub-rows / P / F straight to global memory (coalesced 216-byte pieces);
//   * no shared-memory read-modify-write accumulation, hence no colour ordering and no polled flags: FP64-bound (P, T)
//     and LSU-bound (G) work of different rows overlap instead of alternating in lock step;
//   * the only state carried between element planes is the dx = 0 part of the upper node plane's rows (81 doubles per
//     node column, shared memory, owned by the column's gather warp).
//
// Slot geometry: mma row/column r <-> node with (dx,dy,dz) = ((r>>1)&1, r>>2, r&1)  (rowNode permutation), so block
// (row r, col c) sits at (8 r + c) * 9 and the upper node plane is +2 in r and c (+144 / +18 doubles).
#pragma once
#include "ewb_sweep.cuh"

namespace ewb {

__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ bool mbarTryWait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smemAddr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a logic error sets bit 2 of the status word and the CTA-wide abort flag (all later waits return at
// once) instead of hanging the GPU.
__device__ __forceinline__ void mbarWait(uint64_t* bar, unsigned parity, volatile int* abortFlag, int* failFlag) {
    if (mbarTryWait(bar, parity)) return;
    int spins = 0;
    while (!mbarTryWait(bar, parity)) {
        if (*abortFlag) return;
        if (++spins > (1 << 22)) {
            *abortFlag = 1;
            atomicOr(failFlag, 4);
            return;
        }
    }
}

#ifdef EWB_TIMING
#define RP_T0() long long rpT = clock64()
#define RP_LAP(slot) do { const long long rpN = clock64(); rpAcc[slot] += rpN - rpT; rpT = rpN; } while (0)
#define RP_DECL() long long rpAcc[6] = {0, 0, 0, 0, 0, 0}; const long long rpStart = clock64()
#define RP_FLUSH(cnt) do { if (A.timing != nullptr && lane == 0) { long long* o = A.timing + ((size_t)blockIdx.x * (NT / 32) + warp) * 8; \
    for (int i = 0; i < 6; ++i) o[i] = rpAcc[i]; o[6] = clock64() - rpStart; o[7] = (cnt); } } while (0)
#else
#define RP_T0()
#define RP_LAP(slot)
#define RP_DECL()
#define RP_FLUSH(cnt)
#endif

template <int MC, bool TL, int TZ, int NPW, int RECST = 2, bool USEH = true>
struct RowPipeLayout {
    static constexpr int NE = TZ + 1;  // elements per row (incl. halo)
    static constexpr bool HREC = (MC == MC_LE) && !TL && USEH;  // fragment-order records (the producers compute the scaled gradients)
    static constexpr int PEL = HREC ? RecLayoutH::PER_EL : RecLayout<MC>::PER_EL;
    static constexpr int SLOT_EL = 600;  // 576 stiffness-block doubles [lane][18] + 24 residual doubles [row][3]
    static constexpr int SLOT_STAGES = 3;
    static constexpr int REC_STAGES = RECST;
    static constexpr int STAGE_EL = 120;  // 2 x 2 x 5 nodes x (x,y,z,u0,u1,u2)
    static constexpr int OFF_SLOTS = 0;
    static constexpr int OFF_REC = OFF_SLOTS + SLOT_STAGES * NE * SLOT_EL;
    static constexpr int OFF_STAGE = OFF_REC + REC_STAGES * NE * PEL;
    static constexpr int OFF_BAR = OFF_STAGE + NPW * 2 * STAGE_EL;  // 2*REC_STAGES + 2*SLOT_STAGES mbarriers, abort flag
    static constexpr int OFF_ZERO = OFF_BAR + 16;  // zeros: what a gather lane reads for a colour whose element does not hold its neighbour
    static constexpr int ZERO_PAD = 176;
    static constexpr int OFF_CARRY = OFF_ZERO + ZERO_PAD;
    static constexpr int CARRY_COL = 81 + 6;  // per node column: dx=0 rows of the upper plane [i][27] + P[3], F[3]
    static constexpr int fixedDoubles() { return OFF_CARRY; }
    static constexpr int carryDoubles(int rows) { return rows * TZ * CARRY_COL; }
    static_assert(NE % 4 == 0, "a row is processed in strips of four elements");
};

// RP/RT/RG > 0: per-role register budgets (setmaxnreg, warp groups of four warps: NPW, NTW, NGW must be multiples of 4)
template <int MC, bool TL, int TZ, int NPW, int NTW, int NGW, int RP = 0, int RT = 0, int RG = 0, int RECST = 2, bool USECHAIN = true, bool USEH = true>
__global__ void __launch_bounds__((NPW + NTW + NGW) * 32, 1) rowPipeKernel(const SweepArgs A) {
    static_assert(RP == 0 || (NPW % 4 == 0 && NTW % 4 == 0 && NGW % 4 == 0), "setmaxnreg works on warp groups");
    // A P warp's consecutive tasks are NPW / HALVES rows apart and wait on the parity of a record stage only: the wait is
    // unambiguous as long as the warp cannot be two phases ahead of the T warps, i.e. row stride <= number of record stages.
    static_assert((NPW + (TZ + 1) / 4 - 1) / ((TZ + 1) / 4) <= RECST, "row stride of the producer warps must not exceed the record ring depth");
    constexpr int REG0 = (65536 / ((NPW + NTW + NGW) * 32)) / 8 * 8;  // registers per thread at launch
    using L = RowPipeLayout<MC, TL, TZ, NPW, RECST, USEH>;
    using R = RecLayout<MC>;
    constexpr int NE = L::NE, PEL = L::PEL, SLOT_EL = L::SLOT_EL;
    constexpr bool HREC = L::HREC;
    constexpr int NT = (NPW + NTW + NGW) * 32;
    constexpr int HALVES = NE / 4;
    // y-chaining: with one T warp per element position, the warp adds the previous element row's blocks of the shared y-face
    // to this row's (registers), so that the gather reads every (column, neighbour) pair from ONE element row.  Odd element
    // rows use the row index with the y bit inverted (flip), which puts the shared face at the same fragment positions.
    constexpr bool CHAIN = (NTW == NE) && USECHAIN;

    extern __shared__ double smem[];
    double* slots = smem + L::OFF_SLOTS;
    double* records = smem + L::OFF_REC;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* recFull = bars;                       // [REC_STAGES]
    uint64_t* recEmpty = bars + L::REC_STAGES;      // [REC_STAGES]
    uint64_t* slotFull = bars + 2 * L::REC_STAGES;  // [SLOT_STAGES]
    uint64_t* slotEmpty = slotFull + L::SLOT_STAGES;
    volatile int* abortFlag = reinterpret_cast<volatile int*>(bars + 2 * L::REC_STAGES + 2 * L::SLOT_STAGES);
    double* carry = smem + L::OFF_CARRY;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NX = A.nX + 1, NY = A.nY + 1, NZ = A.nZ + 1;

    // work item -> (chunk, tile)
    int item = blockIdx.x;
    const int tz = item % A.tilesZ; item /= A.tilesZ;
    const int ty = item % A.tilesY; item /= A.tilesY;
    const int chunk = item + A.chunkBase;
    const int y0 = ty * A.tileRows, z0 = tz * TZ;
    const int ny = min(A.tileRows, NY - y0), nz = min(TZ, NZ - z0);  // owned node rows / columns
    const int xa = chunk * A.chunkLen, xb = min(xa + A.chunkLen, NX);
    const int exBegin = max(xa - 1, 0), exEnd = min(xb - 1, A.nX - 1);
    const int nSteps = exEnd - exBegin + 1;
    const int rowsPerStep = ny + 1;  // element rows j = -1 .. ny-1
    const int nRows = nSteps * rowsPerStep;

    for (int i = tid; i < L::ZERO_PAD + L::carryDoubles(A.tileRows); i += NT) smem[L::OFF_ZERO + i] = 0.0;
    if (tid == 0) {
        for (int i = 0; i < L::REC_STAGES; ++i) {
            mbarInit(recFull + i, HALVES * 32);
            mbarInit(recEmpty + i, NTW * 32);
        }
        for (int i = 0; i < L::SLOT_STAGES; ++i) {
            mbarInit(slotFull + i, NTW * 32);
            mbarInit(slotEmpty + i, NGW * 32);
        }
        *abortFlag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // the only CTA-wide barrier

    const int64_t cstride = (int64_t)A.nX * A.nY * A.nZ * 8;

    if (warp < NPW) {
        if constexpr (RP > 0) {
            if constexpr (RP >= REG0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RP));
            else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RP));
        }
        // ===================== P warps: phase A, strips of four elements =====================
        const double* __restrict__ uSrc = TL ? A.U : A.dU;
        double* stageBuf = smem + L::OFF_STAGE + warp * 2 * L::STAGE_EL;
        const unsigned stageAddr0 = smemAddr(stageBuf);
        const int ak = lane >> 3, agp = lane & 7;
        const int nTasks = nRows * HALVES;
        constexpr int NST = 12 + (MC != MC_LE ? 1 : 0);
        // task t = (row n, strip h); this warp's tasks are t = warp, warp + NPW, ...  The decoded position of the task being
        // computed (c*) and of the one whose nodal data is being fetched (f*) advance incrementally: no divisions in the loop.
        struct Pos { int n, h, s, jj; };
        auto advance = [&](Pos& p) {
            const int hn = p.h + NPW;
            const int dn = hn / HALVES;
            p.h = hn % HALVES;
            p.n += dn;
            p.jj += dn;
            while (p.jj >= rowsPerStep) { p.jj -= rowsPerStep; ++p.s; }
        };
        auto issue = [&](const Pos& p, int par) {
            if (p.n < nRows) {
                const int exs = exBegin + p.s, ey = y0 - 1 + p.jj;
                if (ey >= 0 && ey < A.nY) {
                    if (lane < 20) {
                        const int X = lane / 10, Y = (lane / 5) & 1, Z = lane % 5;
                        const int iy = ey + Y, iz = z0 - 1 + 4 * p.h + Z;
                        if (iz >= 0 && iz < NZ) {
                            const int64_t o = 3 * (((int64_t)(exs + X) * NY + iy) * NZ + iz);
                            const unsigned dst = stageAddr0 + (unsigned)par * (L::STAGE_EL * 8u) + 48u * lane;
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * c), "l"(A.coords + o + c) : "memory");
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 24u + 8u * c), "l"(uSrc + o + c) : "memory");
                            }
                        }
                    }
#ifndef EWB_NO_STATE_PREFETCH  // L2 prefetch of the next task's Gauss-point state (622 -> 647 Melem/s once the producers are the critical role)
                    const int k = 4 * p.h + ak, ez = z0 - 1 + k;
                    if (ez >= 0 && ez < A.nZ && k <= nz) {
                        const double* sp = A.stateRef + (((int64_t)exs * A.nY + ey) * A.nZ + ez) * 8 + agp;
#pragma unroll
                        for (int c = 0; c < NST; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + c * cstride));
                    }
#endif
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        int par = 0;
        RP_DECL();
        Pos cur{warp / HALVES, warp % HALVES, 0, warp / HALVES};
        while (cur.jj >= rowsPerStep) { cur.jj -= rowsPerStep; ++cur.s; }
        Pos nxt = cur;
        issue(cur, par);
#pragma unroll 1
        for (; cur.n < nRows; par ^= 1) {
            RP_T0();
            advance(nxt);
            issue(nxt, par ^ 1);
            const int n = cur.n, h = cur.h, jj = cur.jj;
            const int ex = exBegin + cur.s, ey = y0 - 1 + jj;
            const int rs = n % L::REC_STAGES;
            RP_LAP(2);
            mbarWait(recEmpty + rs, ((n / L::REC_STAGES) & 1) ^ 1, abortFlag, A.failFlag);
            RP_LAP(0);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
            RP_LAP(1);
            const int k = 4 * h + ak, ez = z0 - 1 + k;
            if (ey >= 0 && ey < A.nY && ez >= 0 && ez < A.nZ && k <= nz) {
                double* rec = records + (size_t)(rs * NE + k) * PEL + (HREC ? 0 : agp * R::RS);
                const int64_t off = (((int64_t)ex * A.nY + ey) * A.nZ + ez) * 8 + agp;
                const bool writeState = ex >= xa && jj >= 1 && k >= 1;
                gaussPointCompact<MC, TL, 2, HREC>(rec, stageBuf + par * L::STAGE_EL + ak * 6, agp, A.mp, A.stateRef + off, A.stateTemp + off, cstride,
                                                   writeState, A.failFlag, nullptr, CHAIN ? 16 * (jj & 1) : 0);
            }
            __syncwarp();
            mbarArrive(recFull + rs);
            RP_LAP(3);
            cur = nxt;
        }
        RP_FLUSH(nTasks);
        return;
    }

    if (warp < NPW + NTW) {
        if constexpr (RT > 0) {
            if constexpr (RT >= REG0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RT));
            else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RT));
        }
        // ===================== T warps: stiffness blocks of one element per warp, into the row's slot ring =====================
        const int tw = warp - NPW;
        const int bRow = lane >> 2, bq = lane & 3;
        double dNl[2][3];
        auto setShapeDerivs = [&](int flip) {  // lane-constant dN of the lane's row node (rows of flipped element rows have the y bit inverted)
            const int na = rowNode(bRow ^ (flip << 2));
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
        };
        if constexpr (!HREC) setShapeDerivs(0);
        const bool wantK = A.wantK != 0;
        RP_DECL();
        // One element row: tiles -> blocks (Kc), chained with the previous row's blocks (Kq) of the same element position.  The two
        // register sets swap roles every row (the loop below is unrolled by two), so the chaining costs no register moves.
        int jj = 0;  // element row inside the plane step, advanced incrementally (no division per row)
        auto row = [&](int n, double (&Kc0)[9], double (&Kc1)[9], const double (&Kq0)[9], const double (&Kq1)[9]) {
            const int ey = y0 - 1 + jj;
            const int rs = n % L::REC_STAGES, ss = n % L::SLOT_STAGES;
            RP_T0();
            mbarWait(recFull + rs, (n / L::REC_STAGES) & 1, abortFlag, A.failFlag);
            RP_LAP(0);
            bool slotReady = false;
            const int flip = CHAIN ? (jj & 1) : 0;
            if constexpr (CHAIN && !HREC) setShapeDerivs(flip);
#pragma unroll
            for (int k = tw; k < NE; k += NTW) {
                const int ez = z0 - 1 + k;
                const bool valid = ey >= 0 && ey < A.nY && ez >= 0 && ez < A.nZ && k <= nz;  // warp uniform
                double Pr[3];
                if (valid) {
                    const double* T = records + (size_t)(rs * NE + k) * PEL;
                    TileAcc<MC> acc;
                    if constexpr (HREC) elementTilesH(T, lane, wantK, acc, Pr);
                    else elementTiles<MC, MC == MC_VM>(T, lane, dNl, A.mp, wantK, acc, Pr);
                    finishBlock<MC>(acc, 0, A.mp, Kc0);
                    finishBlock<MC>(acc, 1, A.mp, Kc1);
                    if constexpr (CHAIN) {
                        if (jj != 0 && (bRow >> 2) == flip && (bq >> 1) == flip) {  // both nodes on the face shared with the previous row
#pragma unroll
                            for (int i = 0; i < 9; ++i) { Kc0[i] += Kq0[i]; Kc1[i] += Kq1[i]; }
                        }
                    }
                } else if constexpr (CHAIN) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) Kc0[i] = Kc1[i] = 0.0;  // nothing to chain into the next row
                }
                if (!slotReady) {
                    RP_LAP(2);
                    mbarWait(slotEmpty + ss, ((n / L::SLOT_STAGES) & 1) ^ 1, abortFlag, A.failFlag);
                    slotReady = true;
                    RP_LAP(1);
                }
                if (valid) {
                    double* slot = slots + (size_t)(ss * NE + k) * SLOT_EL;
                    if (wantK) {
                        double2* dst = reinterpret_cast<double2*>(slot + lane * 18);
                        dst[0] = make_double2(Kc0[0], Kc0[1]);
                        dst[1] = make_double2(Kc0[2], Kc0[3]);
                        dst[2] = make_double2(Kc0[4], Kc0[5]);
                        dst[3] = make_double2(Kc0[6], Kc0[7]);
                        dst[4] = make_double2(Kc0[8], Kc1[0]);
                        dst[5] = make_double2(Kc1[1], Kc1[2]);
                        dst[6] = make_double2(Kc1[3], Kc1[4]);
                        dst[7] = make_double2(Kc1[5], Kc1[6]);
                        dst[8] = make_double2(Kc1[7], Kc1[8]);
                    }
                    if (bq == 0) {
                        slot[576 + 3 * bRow] = Pr[0];
                        slot[576 + 3 * bRow + 1] = Pr[1];
                        slot[576 + 3 * bRow + 2] = Pr[2];
                    }
                }
            }
            if (!slotReady) mbarWait(slotEmpty + ss, ((n / L::SLOT_STAGES) & 1) ^ 1, abortFlag, A.failFlag);  // never arrive ahead of the ring
            __syncwarp();
            mbarArrive(recEmpty + rs);
            mbarArrive(slotFull + ss);
            RP_LAP(3);
            if (++jj == rowsPerStep) jj = 0;
        };
        if constexpr (CHAIN) {
            double Ka0[9], Ka1[9], Kb0[9], Kb1[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Ka0[i] = Ka1[i] = Kb0[i] = Kb1[i] = 0.0;
#pragma unroll 1
            for (int n = 0; n < nRows; n += 2) {
                row(n, Ka0, Ka1, Kb0, Kb1);
                if (n + 1 < nRows) row(n + 1, Kb0, Kb1, Ka0, Ka1);
            }
        } else {
            double Ka0[9], Ka1[9];  // no chaining: one block set
#pragma unroll 1
            for (int n = 0; n < nRows; ++n) row(n, Ka0, Ka1, Ka0, Ka1);
        }
        RP_FLUSH(nRows);
        return;
    }

    // ===================== G warps: per node column, gather the row pair's elements and store the finished rows =====================
    if constexpr (RG > 0) {
        if constexpr (RG >= REG0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RG));
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RG));
    }
    const int gw = warp - NPW - NTW;
    auto pre = [](int i) { return i == 0 ? 0 : 3 * i - 1; };
    const int totY = 3 * NY - 2, totZ = 3 * NZ - 2;
    const int64_t totYZ = (int64_t)totY * totZ;
    // lane constants: lane < 27 owns output (neighbour offset s9 = (dy,dz), column component j) of every CSR sub-row
    const int s9 = lane / 3, jc = lane - 3 * s9;
    const int dy = s9 / 3 - 1, dz = s9 % 3 - 1;
    int cbase[4];  // per colour (cy,cz): offset of block (a_lo, b_lo)[0][jc] inside the element slot, -1 = the colour's element does not hold this neighbour
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int cy = c >> 1, cz = c & 1;
        const int py = 1 - cy, pz = 1 - cz;  // in-plane position of the column's node inside the element
        const int by = py + dy, bz = pz + dz;
        const bool act = lane < 27 && by >= 0 && by <= 1 && bz >= 0 && bz <= 1;
        cbase[c] = act ? (8 * (4 * py + pz) + (4 * by + bz)) * 9 + jc : -1;
    }
    // CHAIN: the element row that holds the lane's neighbour is fixed by dy (dy = -1: row below, else the row above, which carries the
    // chained sum for dy = 0); per flip parity f of the upper row and per z-colour: offset of block (a_lo, b_lo)[0][jc], -1 = not held
    int cbz[2][2];
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int cz = 0; cz < 2; ++cz) {
            const int pz = 1 - cz, bz = pz + dz;
            const bool act = lane < 27 && bz >= 0 && bz <= 1;
            const int rowA = 4 * f + pz, rowB = (dy == 0 ? 4 * f : 4 * (1 - f)) + bz;
            cbz[f][cz] = act ? (8 * rowA + rowB) * 9 + jc : -1;
        }
    // residual pass (one warp per row): lane < 3 TZ owns (column lane / 3, component lane % 3)
    const int pfCol = lane / 3, pfi = lane - 3 * pfCol;
    const bool wantK = A.wantK != 0;
    const double* zeroPad = smem + L::OFF_ZERO;

    RP_DECL();
    int jjNext = 0, exNext = exBegin;  // advanced incrementally (no division per row)
#pragma unroll 1
    for (int n = 0; n < nRows; ++n) {
        const int jj = jjNext, ex = exNext;
        if (++jjNext == rowsPerStep) { jjNext = 0; ++exNext; }
        const int ss = n % L::SLOT_STAGES;
        RP_T0();
        mbarWait(slotFull + ss, (n / L::SLOT_STAGES) & 1, abortFlag, A.failFlag);
        RP_LAP(0);
        if (jj == 0) continue;  // halo row below the tile: nothing finished yet
        const int ssPrev = (n - 1) % L::SLOT_STAGES;
        const int ly = jj - 1, iy = y0 + ly;
        const bool loOwned = ex >= xa, hiOwned = (ex + 1) < xb;
        const bool lastPlane = (ex == exEnd) && (xb == NX);  // node plane NX-1: nothing above it, its dx=0 rows are final now
        const int cyN = (iy > 0) + 1 + (iy < NY - 1);
        const bool vy0 = iy - 1 >= 0, vy1 = iy < A.nY;
        const double* rowLo = slots + (size_t)(ssPrev * NE) * SLOT_EL;  // element row iy - 1 (colours cy = 0)
        const double* rowHi = slots + (size_t)(ss * NE) * SLOT_EL;      // element row iy     (colours cy = 1)
        // x- and y-interior row of an interior plane step: every column with an interior z takes the lean path
        const bool leanRow = wantK && ex >= 1 && ex + 1 <= NX - 2 && iy >= 1 && iy <= NY - 2;
        const int flip = jj & 1;  // CHAIN: flip parity of the upper element row (the lower one has the opposite parity)
        double* leanBase = A.data + (9 * (int64_t)(3 * ex - 1) * totYZ + 27 * ((int64_t)(3 * iy - 1) * totZ - 3) + lane);
        const int64_t nextPlane = 27 * totYZ;
        // columns of this warp: (ly * TZ + lzz) % NGW == gw
        for (int lzz = (((gw - ly * TZ) % NGW) + NGW) % NGW; lzz < nz; lzz += NGW) {
            const int cidx = ly * TZ + lzz;
            const int iz = z0 + lzz;
            double* cc = carry + (size_t)cidx * L::CARRY_COL;
            RP_LAP(3);
            if (leanRow && iz >= 1 && iz <= NZ - 2) {
                // ---- lean path: all four elements exist, the 9 sub-row pieces of the column are 27-double runs at base + m * 27.
                // Branch-free: a lane whose neighbour is not in colour c's element reads zeros, so all 48 loads are independent. ----
                if (lane < 27) {
                    double v[2][2][3];
                    if constexpr (CHAIN) {
                        const double* rowSel = (dy == -1 ? rowLo : rowHi) + lzz * SLOT_EL;
                        const int c0 = flip ? cbz[1][0] : cbz[0][0], c1 = flip ? cbz[1][1] : cbz[0][1];
                        const double* e0 = c0 >= 0 ? rowSel + c0 : zeroPad;
                        const double* e1 = c1 >= 0 ? rowSel + SLOT_EL + c1 : zeroPad;
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int b = 0; b < 2; ++b)
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    const int o = 144 * a + 18 * b + 3 * i;
                                    v[a][b][i] = e0[o] + e1[o];
                                }
                    } else {
                        const double* e0 = cbase[0] >= 0 ? rowLo + lzz * SLOT_EL + cbase[0] : zeroPad;
                        const double* e1 = cbase[1] >= 0 ? rowLo + (lzz + 1) * SLOT_EL + cbase[1] : zeroPad;
                        const double* e2 = cbase[2] >= 0 ? rowHi + lzz * SLOT_EL + cbase[2] : zeroPad;
                        const double* e3 = cbase[3] >= 0 ? rowHi + (lzz + 1) * SLOT_EL + cbase[3] : zeroPad;
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int b = 0; b < 2; ++b)
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    const int o = 144 * a + 18 * b + 3 * i;
                                    v[a][b][i] = (e0[o] + e1[o]) + (e2[o] + e3[o]);
                                }
                    }
                    RP_LAP(1);
                    double* ptr = leanBase + 243 * (int64_t)iz;
                    double* ptrN = ptr + nextPlane;
                    if (loOwned) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            ptr[(3 * i + 1) * 27] = v[0][0][i] + cc[27 * i + lane];
                            ptr[(3 * i + 2) * 27] = v[0][1][i];
                        }
                    }
                    if (hiOwned) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            ptrN[(3 * i) * 27] = v[1][0][i];
                            cc[27 * i + lane] = v[1][1][i];
                        }
                    }
                    RP_LAP(2);
                }
                continue;
            }
            // ---- general path: box faces / edges, first and last planes of the box and of a chunk, peer planes ----
            const bool vz0 = iz - 1 >= 0, vz1 = iz < A.nZ;
            double v[2][2][3];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) v[a][b][0] = v[a][b][1] = v[a][b][2] = 0.0;
            if constexpr (CHAIN) {
                // the lane's source row: below for dy = -1, above otherwise — except on the top face of the box (no element row
                // above), where the dy = 0 blocks are the lower row's own y-face blocks (same offsets, see cbz)
                const bool useLo = dy == -1 || !vy1;
                const bool rowOk = dy == -1 ? vy0 : (dy == 1 ? vy1 : true);
                const double* rowSel = (useLo ? rowLo : rowHi) + lzz * SLOT_EL;
#pragma unroll
                for (int cz = 0; cz < 2; ++cz) {
                    const int cb = flip ? cbz[1][cz] : cbz[0][cz];
                    if (wantK && rowOk && (cz ? vz1 : vz0) && cb >= 0) {
                        const double* eb = rowSel + cz * SLOT_EL + cb;
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int b = 0; b < 2; ++b)
#pragma unroll
                                for (int i = 0; i < 3; ++i) v[a][b][i] += eb[144 * a + 18 * b + 3 * i];
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int cy = c >> 1, cz = c & 1;
                    const bool cv = (cy ? vy1 : vy0) && (cz ? vz1 : vz0);  // warp uniform
                    if (!cv) continue;
                    if (wantK && cbase[c] >= 0) {
                        const double* eb = (cy ? rowHi : rowLo) + (lzz + cz) * SLOT_EL + cbase[c];
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int b = 0; b < 2; ++b)
#pragma unroll
                                for (int i = 0; i < 3; ++i) v[a][b][i] += eb[144 * a + 18 * b + 3 * i];
                    }
                }
            }
            // outputs of this column: node plane ex (rows dx = 0, +1) and node plane ex + 1 (dx = -1; carry)
            const int czN = (iz > 0) + 1 + (iz < NZ - 1);
            const int cycz = cyN * czN;
            const int colPart = pre(iy) * totZ + cyN * pre(iz);
            int lo = -1;
            if (lane < 27 && iy + dy >= 0 && iy + dy < NY && iz + dz >= 0 && iz + dz < NZ) lo = 3 * ((dy + (iy > 0 ? 1 : 0)) * czN + dz + (iz > 0 ? 1 : 0)) + jc;
            // store one piece (three sub-rows) of node plane ix: dxs = dx + 1 in {0,1,2}
            auto storePiece = [&](int ix, int dxs, const double (&val)[3]) {
                const int cx = (ix > 0) + 1 + (ix < NX - 1);
                const bool toPeer = A.peerData != nullptr && ix == NX - 1;
                double* xbase = toPeer ? A.peerData : A.data + 9 * (int64_t)pre(ix) * totYZ;
                const int rx0 = ix > 0 ? 1 : 0;
                double* dst = xbase + ((int64_t)(9 * cx) * colPart + lo) + 3 * ((dxs - 1 + rx0) * cycz);
                const int rowStride = 3 * cx * cycz;
                dst[0] = val[0];
                dst[rowStride] = val[1];
                dst[2 * rowStride] = val[2];
            };
            if (wantK && lane < 27) {
                if (loOwned) {
                    double v0[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) v0[i] = cc[27 * i + lane] + v[0][0][i];
                    if (lo >= 0) {
                        storePiece(ex, 1, v0);
                        storePiece(ex, 2, v[0][1]);
                    }
                }
                if (hiOwned) {
                    if (lo >= 0) storePiece(ex + 1, 0, v[1][0]);
                    if (lastPlane) {
                        if (lo >= 0) storePiece(ex + 1, 1, v[1][1]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 3; ++i) cc[27 * i + lane] = v[1][1][i];
                    }
                }
            }
        }
        RP_LAP(3);
        // ---- residual of the row's node columns: P, F of node plane ex; carry of node plane ex + 1 (one warp per row) ----
        if (n % NGW == gw && pfCol < nz) {
            const int iz = z0 + pfCol;
            const bool vz0 = iz - 1 >= 0, vz1 = iz < A.nZ;
            double pl = 0.0, ph = 0.0, fl = 0.0, fh = 0.0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int cy = c >> 1, cz = c & 1;
                if ((cy ? vy1 : vy0) && (cz ? vz1 : vz0)) {
                    // fragment row of the column's node in element (cy, cz); CHAIN: rows of flipped element rows have the y bit inverted
                    const int fr = CHAIN ? 4 * flip + (1 - cz) : 4 * (1 - cy) + (1 - cz);
                    const double* e = (cy ? rowHi : rowLo) + (pfCol + cz) * SLOT_EL + 576 + 3 * fr + pfi;
                    const double plo = e[0], phi = e[6];
                    pl += plo; fl += fabs(plo);
                    ph += phi; fh += fabs(phi);
                }
            }
            double* cc = carry + (size_t)(ly * TZ + pfCol) * L::CARRY_COL + 81;
            auto storePF = [&](int ix, double pv, double fv) {
                const int64_t dof = 3 * ((((int64_t)ix * NY + iy) * NZ) + iz) + pfi;
                if (A.peerData != nullptr && ix == NX - 1) {
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
            };
            if (loOwned) storePF(ex, cc[pfi] + pl, cc[3 + pfi] + fl);
            if (hiOwned) {
                if (lastPlane) {
                    storePF(ex + 1, ph, fh);
                } else {
                    cc[pfi] = ph;
                    cc[3 + pfi] = fh;
                }
            }
        }
        RP_LAP(4);
        __syncwarp();
        mbarArrive(slotEmpty + ssPrev);
        if (jj == rowsPerStep - 1) mbarArrive(slotEmpty + ss);
        RP_LAP(5);
    }
    RP_FLUSH(nRows);
}

// ---- host side --------------------------------------------------------------------------------------------------
// Tile decomposition: z in tiles of TZ node columns, y in `tilesY` tiles of `tileRows` node rows (bounded by the carry
// buffer that has to fit next to the rings in shared memory), x in chunks.  The search minimises
// (CTA rounds on nSM SMs) x (element rows per plane incl. the halo row) x (element planes per chunk incl. the halo plane).
struct RowPipeTiling {
    int tilesY, tilesZ, tileRows, chunkLen, nChunks;
};

template <int TZ>
inline RowPipeTiling rowPipeTiling(int64_t nX, int64_t nY, int64_t nZ, int nSM, int rowsMax, int chunkOverride) {
    const int NX = (int)nX + 1, NY = (int)nY + 1, NZ = (int)nZ + 1;
    RowPipeTiling best{1, 1, 1, NX, 1};
    double bestCost = 1e300;
    const int tilesZ = (NZ + TZ - 1) / TZ;
    for (int ty = 1; ty <= NY; ++ty) {
        const int rows = (NY + ty - 1) / ty;
        if (rows > rowsMax) continue;
        const int tilesY = (NY + rows - 1) / rows;
        for (int c = 1; c <= 64; ++c) {
            if (chunkOverride > 0 && c != chunkOverride) continue;
            const int len = (NX + c - 1) / c;
            if (c > 1 && len < 4) break;
            const int nc = (NX + len - 1) / len;
            const double rounds = (double)(((int64_t)tilesZ * tilesY * nc + nSM - 1) / nSM);
            const double cost = rounds * (rows + 1) * (nc > 1 ? len + 1 : len - 1);  // element planes swept per CTA
            if (cost < bestCost) { bestCost = cost; best = RowPipeTiling{tilesY, tilesZ, rows, len, nc}; }
        }
        if (rows == 1) break;
    }
    // an override that no tile shape accepts (chunks shorter than four planes) is ignored rather than leaving the placeholder tiling
    if (bestCost == 1e300 && chunkOverride > 0) return rowPipeTiling<TZ>(nX, nY, nZ, nSM, rowsMax, 0);
    return best;
}

template <int MC, bool TL, int TZ, int NPW, int NTW, int NGW, int RP = 0, int RT = 0, int RG = 0, int RECST = 2, bool USECHAIN = true, bool USEH = true>
int launchRowPipe(SweepPlan& sp, const MatParams& mp, const ewb_buffers* b, int* failFlag, int flags, cudaStream_t st) {
    using L = RowPipeLayout<MC, TL, TZ, NPW, RECST, USEH>;
    constexpr int SMEM_MAX = 232448;
    const int rowsMax = (SMEM_MAX / 8 - L::fixedDoubles()) / (TZ * L::CARRY_COL);
    if (rowsMax < 1) return EWB_ERR_UNSUPPORTED;
    SweepArgs a;
    sp.fillCommon(a, mp, b, failFlag, flags);
    const bool piecewise = sp.tilingOut != nullptr || sp.chunkEnd >= 0;  // ewb_plan_x_chunks / ewb_assemble_chunks
    const int pipeChunks = sp.pipeChunks >= 0 ? sp.pipeChunks : (sp.chunkOverride == 0 && sp.nX + 1 >= 48 ? 6 : 0);
    const RowPipeTiling t = rowPipeTiling<TZ>(sp.nX, sp.nY, sp.nZ, sp.nSM, rowsMax, piecewise && pipeChunks > 0 ? pipeChunks : sp.chunkOverride);
    a.tilesY = t.tilesY; a.tilesZ = t.tilesZ; a.tileRows = t.tileRows; a.chunkLen = t.chunkLen; a.nChunks = t.nChunks;
    if (sp.tilingOut != nullptr) {  // query only (ewb_plan_x_chunks)
        sp.tilingOut[0] = t.chunkLen;
        sp.tilingOut[1] = t.nChunks;
        return EWB_OK;
    }
    const int c0 = sp.chunkEnd < 0 ? 0 : sp.chunkBegin, c1 = sp.chunkEnd < 0 ? t.nChunks : sp.chunkEnd;
    if (c0 < 0 || c1 > t.nChunks || c0 >= c1) return EWB_ERR_ARG;
    a.chunkBase = c0;
    const int64_t grid = (int64_t)a.tilesY * a.tilesZ * (c1 - c0);
#ifdef EWB_TIMING
    {
        const size_t nT = (size_t)grid * (NPW + NTW + NGW) * 8;
        if (sp.timingCount < nT) {
            if (sp.timingBuf) cudaFree(sp.timingBuf);
            cudaMalloc((void**)&sp.timingBuf, nT * sizeof(long long));
            sp.timingCount = nT;
        }
        cudaMemsetAsync(sp.timingBuf, 0, nT * sizeof(long long), st);
        a.timing = sp.timingBuf;
    }
#endif
    auto kern = rowPipeKernel<MC, TL, TZ, NPW, NTW, NGW, RP, RT, RG, RECST, USECHAIN, USEH>;
    const size_t smem = ((size_t)L::fixedDoubles() + (size_t)L::carryDoubles(t.tileRows)) * sizeof(double);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return EWB_ERR_CUDA;
    kern<<<(unsigned)grid, (NPW + NTW + NGW) * 32, smem, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? EWB_OK : EWB_ERR_CUDA;
}

// warps: variant code 10000 P + 100 T + G (e.g. 40404 = 4 producer, 4 tensor, 4 gather warps)
template <int MC, bool TL>
int launchRowPipeVariant(SweepPlan& sp, int variant, const MatParams& mp, const ewb_buffers* b, int* failFlag, int flags, cudaStream_t st) {
    switch (variant) {
#ifdef EWB_VARIANTS
        case 40804: return launchRowPipe<MC, TL, 7, 4, 8, 4>(sp, mp, b, failFlag, flags, st);
        case 30405: return launchRowPipe<MC, TL, 7, 3, 4, 5>(sp, mp, b, failFlag, flags, st);
        case 1040804: return launchRowPipe<MC, TL, 7, 4, 8, 4, 168, 112, 120>(sp, mp, b, failFlag, flags, st);
        case 2040404: return launchRowPipe<MC, TL, 7, 4, 4, 4, 0, 0, 0, 3>(sp, mp, b, failFlag, flags, st);
        case 10040804: return launchRowPipe<MC, TL, 7, 4, 8, 4, 152, 128, 104, 3, true, false>(sp, mp, b, failFlag, flags, st);
        case 11040804: return launchRowPipe<MC, TL, 7, 4, 8, 4, 144, 136, 96, 4, true, false>(sp, mp, b, failFlag, flags, st);
        case 5080404: return launchRowPipe<MC, TL, 7, 8, 4, 4, 152, 128, 80, 4>(sp, mp, b, failFlag, flags, st);
        case 7080404: return launchRowPipe<MC, TL, 7, 8, 4, 4, 144, 144, 80, 4>(sp, mp, b, failFlag, flags, st);
#endif
        case 40404: return launchRowPipe<MC, TL, 7, 4, 4, 4>(sp, mp, b, failFlag, flags, st);
        case 8040804: return launchRowPipe<MC, TL, 7, 4, 8, 4, 144, 136, 96, 3, false>(sp, mp, b, failFlag, flags, st);  // no chaining (von Mises / Neo-Hooke)
        case 9040804: return launchRowPipe<MC, TL, 7, 4, 8, 4, 152, 136, 88, 3, false>(sp, mp, b, failFlag, flags, st);
#ifdef EWB_VARIANTS
        case 2040804: return launchRowPipe<MC, TL, 7, 4, 8, 4, 168, 112, 120, 3>(sp, mp, b, failFlag, flags, st);
        case 3040804: return launchRowPipe<MC, TL, 7, 4, 8, 4, 168, 120, 104, 3>(sp, mp, b, failFlag, flags, st);
#endif
        case 4040804:
        default: return launchRowPipe<MC, TL, 7, 4, 8, 4, 152, 128, 104, 3>(sp, mp, b, failFlag, flags, st);
    }
}

inline int launchRowPipeAny(SweepPlan& sp, int variant, int elType, int mc, const MatParams& mp, const ewb_buffers* b, int* failFlag, int flags,
                            cudaStream_t st) {
    if (elType == EWB_C3D8 && mc == MC_LE) return launchRowPipeVariant<MC_LE, false>(sp, variant, mp, b, failFlag, flags, st);
    if (elType == EWB_C3D8 && mc == MC_VM) return launchRowPipeVariant<MC_VM, false>(sp, variant, mp, b, failFlag, flags, st);
    if (elType == EWB_C3D8TL && mc == MC_NH) return launchRowPipeVariant<MC_NH, true>(sp, variant, mp, b, failFlag, flags, st);
    return EWB_ERR_UNSUPPORTED;
}

}  // namespace ewb
