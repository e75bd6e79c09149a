// EXPERIMENT (round 2), not part of the product build — kept for the record, see DESIGN.md §6.
// To try it: copy next to ewb_rowpipe.cuh, add to SweepArgs { const int4* items; const int* itemPtr; double* carryG; int64_t carryStride; }
// and to SweepPlan { int4* streamItems; int* streamItemPtr; double* streamCarry; int streamCtas, streamRows, streamTZ; int64_t
// streamCarryStride; }, include it from ewb_api.cu and dispatch launchRowStream<MC, TL, 7, 4, 8, 4, 152, 128, 104, 3>.
// Result on B200, 100^3 linear elastic: parity green (34 / 34 fused-kernel tests), 26-row tiles, 148 CTAs of equal work (8208..9216
// element-row units), work overhead 1.21 instead of 1.45 — but 387-417 Melem/s against 647 for the uniform decomposition: the global
// carry adds 6 global accesses of 216 bytes per node column and plane step to an LSU data pipe that is already 70 % busy (gather
// role +35 % per row even with the carry prefetched a row ahead), and the producers' issue phase doubles.
//
// Row-pipelined gather sweep with STREAMED work items — the row-pipelined kernel of ewb_rowpipe.cuh (same warp roles, rings, y-chaining
// and gather; see there) with two changes that remove most of its redundant halo work:
//   * the plane carry (dx = 0 rows of the upper node plane, 87 doubles per node column) lives in a per-CTA slice of a global buffer
//     that stays in L2 instead of in shared memory, so a tile may have as many node rows as the host likes (y-halo 1 / rows);
//   * the x-ranges of the (few, large) tiles are cut into pieces of equal work and dealt to exactly one CTA per SM ("stream-K" along
//     x): no CTA-count quantisation; the roles walk the CTA's item list, the mbarrier rings simply continue across items.
#pragma once
#include <cmath>
#include <cstdio>
#include <vector>

#include "ewb_rowpipe.cuh"

namespace ewb {

template <int MC, bool TL, int TZ, int NPW, int RECST = 2, bool USEH = true, bool STREAM = false>
struct RowStreamLayout {
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
    static constexpr int carryDoubles(int rows) { return STREAM ? 0 : rows * TZ * CARRY_COL; }  // STREAM: the carry lives in global memory
    static_assert(NE % 4 == 0, "a row is processed in strips of four elements");
};

// RP/RT/RG > 0: per-role register budgets (setmaxnreg, warp groups of four warps: NPW, NTW, NGW must be multiples of 4)
template <int MC, bool TL, int TZ, int NPW, int NTW, int NGW, int RP = 0, int RT = 0, int RG = 0, int RECST = 2, bool USECHAIN = true, bool USEH = true, bool STREAM = false>
__global__ void __launch_bounds__((NPW + NTW + NGW) * 32, 1) rowStreamKernel(const SweepArgs A) {
    static_assert(RP == 0 || (NPW % 4 == 0 && NTW % 4 == 0 && NGW % 4 == 0), "setmaxnreg works on warp groups");
    // A P warp's consecutive tasks are NPW / HALVES rows apart and wait on the parity of a record stage only: the wait is
    // unambiguous as long as the warp cannot be two phases ahead of the T warps, i.e. row stride <= number of record stages.
    static_assert((NPW + (TZ + 1) / 4 - 1) / ((TZ + 1) / 4) <= RECST, "row stride of the producer warps must not exceed the record ring depth");
    constexpr int REG0 = (65536 / ((NPW + NTW + NGW) * 32)) / 8 * 8;  // registers per thread at launch
    using L = RowStreamLayout<MC, TL, TZ, NPW, RECST, USEH, STREAM>;
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
    // plane carry: shared memory, or (STREAM) this CTA's slice of a global buffer that stays in L2 — read and written by the same
    // thread in consecutive plane steps
    double* carry = STREAM ? A.carryG + (size_t)blockIdx.x * A.carryStride : smem + L::OFF_CARRY;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NX = A.nX + 1, NY = A.nY + 1, NZ = A.nZ + 1;

    // Work items of this CTA.  Uniform decomposition: one item (tile, x-chunk) from the block index.  STREAM: a list of
    // (tile, node-plane range) pieces balanced by the host over all CTAs; the warp roles walk the same list, the mbarrier
    // rings and their phase counters simply continue across items (rowBase = rows of the previous items).
    const int itBegin = STREAM ? A.itemPtr[blockIdx.x] : 0, itEnd = STREAM ? A.itemPtr[blockIdx.x + 1] : 1;
    int y0 = 0, z0 = 0, ny = 0, nz = 0, xa = 0, xb = 0, exBegin = 0, exEnd = 0, rowsPerStep = 1, nRows = 0;
    auto loadItem = [&](int it) {
        int ty, tz;
        if constexpr (STREAM) {
            const int4 v = A.items[it];
            ty = v.x; tz = v.y; xa = v.z; xb = v.w;
        } else {
            int item = blockIdx.x;
            tz = item % A.tilesZ; item /= A.tilesZ;
            ty = item % A.tilesY; item /= A.tilesY;
            xa = item * A.chunkLen; xb = min(xa + A.chunkLen, NX);
        }
        y0 = ty * A.tileRows; z0 = tz * TZ;
        ny = min(A.tileRows, NY - y0); nz = min(TZ, NZ - z0);  // owned node rows / columns
        exBegin = max(xa - 1, 0); exEnd = min(xb - 1, A.nX - 1);
        rowsPerStep = ny + 1;  // element rows j = -1 .. ny-1
        nRows = (exEnd - exBegin + 1) * rowsPerStep;
    };

    for (int i = tid; i < L::ZERO_PAD + L::carryDoubles(A.tileRows); i += NT) smem[L::OFF_ZERO + i] = 0.0;  // zero pad (+ shared-memory carry)
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
        constexpr int NST = 12 + (MC != MC_LE ? 1 : 0);
        int rowBase = 0;  // element rows of the previous items: ring stages / phases run on the global row counter
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
        int nTasksAll = 0;
#pragma unroll 1
        for (int it = itBegin; it < itEnd; ++it) {
            loadItem(it);
            // tasks are numbered globally (rowBase * HALVES + local index) and dealt round-robin: the first local task of this warp
            const int t0 = (((warp - rowBase * HALVES) % NPW) + NPW) % NPW;
            Pos cur{t0 / HALVES, t0 % HALVES, 0, t0 / HALVES};
            while (cur.jj >= rowsPerStep) { cur.jj -= rowsPerStep; ++cur.s; }
            Pos nxt = cur;
            issue(cur, par);
#pragma unroll 1
            for (; cur.n < nRows; par ^= 1) {
                RP_T0();
                advance(nxt);
                issue(nxt, par ^ 1);
                const int n = rowBase + cur.n, h = cur.h, jj = cur.jj;
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
            asm volatile("cp.async.wait_group 0;" ::: "memory");  // the trailing (empty) prefetch group of this item
            rowBase += nRows;
            nTasksAll += nRows * HALVES;
        }
        RP_FLUSH(nTasksAll);
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
                    else elementTiles<MC>(T, lane, dNl, A.mp, wantK, acc, Pr);
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
        double Ka0[9], Ka1[9], Kb0[9], Kb1[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Ka0[i] = Ka1[i] = Kb0[i] = Kb1[i] = 0.0;
        int rowBase = 0;
#pragma unroll 1
        for (int it = itBegin; it < itEnd; ++it) {
            loadItem(it);
            jj = 0;
#pragma unroll 1
            for (int n = 0; n < nRows; n += 2) {
                row(rowBase + n, Ka0, Ka1, Kb0, Kb1);
                if (n + 1 < nRows) row(rowBase + n + 1, Kb0, Kb1, Ka0, Ka1);
            }
            rowBase += nRows;
        }
        RP_FLUSH(rowBase);
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
    int rowBase = 0;
#pragma unroll 1
    for (int it = itBegin; it < itEnd; ++it) {
    loadItem(it);
    int jjNext = 0, exNext = exBegin;  // advanced incrementally (no division per row)
#pragma unroll 1
    for (int nl = 0; nl < nRows; ++nl) {
        const int n = rowBase + nl;  // global row: ring stage and phase
        const int jj = jjNext, ex = exNext;
        if (++jjNext == rowsPerStep) { jjNext = 0; ++exNext; }
        const int ss = n % L::SLOT_STAGES;
        RP_T0();
        // STREAM: the carry written one plane step ago sits in L2 (several hundred cycles away): fetch this warp's (at most two)
        // columns of the row before waiting for the row's slots
        double cin[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
        if constexpr (STREAM) {
            static_assert(!STREAM || (TZ + NGW - 1) / NGW <= 2, "carry preload holds two columns per warp and row");
            if (jj != 0 && ex >= xa && ex > 0 && lane < 27) {
                const int lyp = jj - 1;
                const int first = (((gw - lyp * TZ) % NGW) + NGW) % NGW;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int lzp = first + k * NGW;
                    if (lzp < nz) {
                        const double* cp = carry + (size_t)(lyp * TZ + lzp) * L::CARRY_COL + lane;
#pragma unroll
                        for (int i = 0; i < 3; ++i) cin[k][i] = __ldcg(cp + 27 * i);
                    }
                }
            }
        }
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
        int kcol = -1;
        for (int lzz = (((gw - ly * TZ) % NGW) + NGW) % NGW; lzz < nz; lzz += NGW) {
            ++kcol;
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
                            ptr[(3 * i + 1) * 27] = v[0][0][i] + (STREAM ? (kcol == 0 ? cin[0][i] : cin[1][i]) : cc[27 * i + lane]);  // (lean rows have ex >= 1)
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
                    for (int i = 0; i < 3; ++i) v0[i] = (STREAM ? (kcol == 0 ? cin[0][i] : cin[1][i]) : (ex > 0 ? cc[27 * i + lane] : 0.0)) + v[0][0][i];  // no plane below the first one
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
        if (ly % NGW == gw && pfCol < nz) {  // the same warp (thread) owns a node row's P, F carry in every plane step
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
            if (loOwned) storePF(ex, (ex > 0 ? cc[pfi] : 0.0) + pl, (ex > 0 ? cc[3 + pfi] : 0.0) + fl);
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
    rowBase += nRows;
    }
    RP_FLUSH(rowBase);
}

// ---- streamed work items ("stream-K" along x) ----------------------------------------------------------------------
// With the plane carry in global memory a tile may have as many node rows as it likes, so the y-halo shrinks to 1 / rows; the
// price of few, large tiles — CTA counts that do not divide by the number of SMs — is removed by cutting the tiles' x-ranges into
// pieces of equal WORK: the tiles are laid end to end (tile-major, then x), and CTA b takes the b-th equal share of that line, i.e.
// at most the tail of one tile, some whole tiles and the head of another.  Every piece with xa > 0 pays one halo plane.
struct StreamSchedule {
    std::vector<int4> items;
    std::vector<int> itemPtr;
    int rows = 1, nCtas = 1;
};

template <int TZ>
inline StreamSchedule rowStreamSchedule(int64_t nX, int64_t nY, int64_t nZ, int nSM) {
    const int NX = (int)nX + 1, NY = (int)nY + 1, NZ = (int)nZ + 1;
    StreamSchedule sc;
    // tile rows: per CTA about V = NX NY NZ / nSM node columns x planes, TZ columns wide -> rows x len = V / TZ, halo minimal for rows == len
    const double V = (double)NX * NY * NZ / nSM / TZ;
    int rows = (int)std::lround(std::sqrt(std::max(1.0, V)));
    rows = std::max(1, std::min(rows, NY));
    const int tilesY = (NY + rows - 1) / rows;
    rows = (NY + tilesY - 1) / tilesY;
    const int tilesZ = (NZ + TZ - 1) / TZ;
    sc.rows = rows;
    // work of one element-plane step of tile (ty, tz): (rows + 1) x (elements per row)
    auto stepWork = [&](int ty, int tz) {
        const int ny = std::min(rows, NY - ty * rows), nz = std::min(TZ, NZ - tz * TZ);
        return (double)(ny + 1) * (nz + 1);
    };
    double total = 0.0;
    for (int ty = 0; ty < tilesY; ++ty)
        for (int tz = 0; tz < tilesZ; ++tz) total += stepWork(ty, tz) * (NX - 1);
    const int minLen = 3;  // node planes per piece, at least
    int nCtas = (int)std::min<int64_t>(nSM, std::max<int64_t>(1, (int64_t)tilesY * tilesZ * std::max(1, NX / (2 * minLen))));
    sc.nCtas = nCtas;
    sc.itemPtr.assign(1, 0);
    double load = 0.0;       // work already given to the current CTA
    double remaining = total;  // work not yet assigned (without the halo planes of future cuts)
    int cta = 0;
    for (int ty = 0; ty < tilesY; ++ty)
        for (int tz = 0; tz < tilesZ; ++tz) {
            const double w = stepWork(ty, tz);
            int x = 0;  // next unassigned node plane of this tile
            while (x < NX) {
                const bool last = cta == nCtas - 1;
                // the current CTA's fair share of what is left (re-evaluated at every cut: rounding never piles up on the last CTA)
                const double target = (load + remaining) / (nCtas - cta);
                // node planes [x, xb): element planes max(x-1,0) .. min(xb-1, NX-2)
                int take = last ? NX - x : (int)std::floor((target - load) / w - (x > 0 ? 1.0 : 0.0) + 0.5);
                if (take < minLen) take = (load > 0.0 && !last) ? 0 : minLen;  // too little room: close this CTA (unless it is empty)
                if (take > NX - x || NX - x - take < minLen) take = NX - x;     // no tiny remainder
                if (take > 0) {
                    const int xb = x + take;
                    const int steps = std::min(xb - 1, NX - 2) - std::max(x - 1, 0) + 1;
                    sc.items.push_back(make_int4(ty, tz, x, xb));
                    load += w * steps;
                    remaining -= w * (std::min(xb - 1, NX - 2) - std::max(x, 1) + 1 + (x == 0 ? 1 : 0));  // own planes only
                    x = xb;
                }
                if (!last && (take == 0 || load >= target - 0.5 * w)) {  // CTA full: next one
                    sc.itemPtr.push_back((int)sc.items.size());
                    ++cta;
                    load = 0.0;
                }
            }
        }
    while ((int)sc.itemPtr.size() < nCtas + 1) sc.itemPtr.push_back((int)sc.items.size());
    // self-check: the pieces of every tile tile its node planes [0, NX) exactly once, in ascending order
    {
        std::vector<int> next((size_t)tilesY * tilesZ, 0);
        bool ok = (int)sc.itemPtr.size() == nCtas + 1 && sc.itemPtr.back() == (int)sc.items.size();
        for (const int4& it : sc.items) {
            int& x = next[(size_t)it.x * tilesZ + it.y];
            ok = ok && it.z == x && it.w > it.z && it.w <= NX;
            x = it.w;
        }
        for (int v : next) ok = ok && v == NX;
        if (!ok) sc.nCtas = 0;  // reported by the caller
    }
    return sc;
}

template <int MC, bool TL, int TZ, int NPW, int NTW, int NGW, int RP, int RT, int RG, int RECST>
int launchRowStream(SweepPlan& sp, const MatParams& mp, const ewb_buffers* b, int* failFlag, int flags, cudaStream_t st) {
    using L = RowStreamLayout<MC, TL, TZ, NPW, RECST, true, true>;
    SweepArgs a;
    sp.fillCommon(a, mp, b, failFlag, flags);
    if (sp.streamCtas == 0 || sp.streamTZ != TZ) {
        const StreamSchedule sc = rowStreamSchedule<TZ>(sp.nX, sp.nY, sp.nZ, sp.nSM);
        if (sc.nCtas <= 0) return EWB_ERR_ARG;  // scheduler self-check failed
        cudaFree(sp.streamItems); cudaFree(sp.streamItemPtr); cudaFree(sp.streamCarry);
        sp.streamItems = nullptr; sp.streamItemPtr = nullptr; sp.streamCarry = nullptr;
        sp.streamCarryStride = (int64_t)sc.rows * TZ * L::CARRY_COL;
        if (cudaMalloc((void**)&sp.streamItems, std::max<size_t>(1, sc.items.size()) * sizeof(int4)) != cudaSuccess ||
            cudaMalloc((void**)&sp.streamItemPtr, sc.itemPtr.size() * sizeof(int)) != cudaSuccess ||
            cudaMalloc((void**)&sp.streamCarry, (size_t)sc.nCtas * sp.streamCarryStride * sizeof(double)) != cudaSuccess)
            return EWB_ERR_CUDA;
        cudaMemcpy(sp.streamItems, sc.items.data(), sc.items.size() * sizeof(int4), cudaMemcpyHostToDevice);
        cudaMemcpy(sp.streamItemPtr, sc.itemPtr.data(), sc.itemPtr.size() * sizeof(int), cudaMemcpyHostToDevice);
        sp.streamCtas = sc.nCtas; sp.streamRows = sc.rows; sp.streamTZ = TZ;
        if (getenv("EWB_DEBUG_SCHEDULE")) {
            double wmin = 1e300, wmax = 0.0;
            for (int c = 0; c < sc.nCtas; ++c) {
                double wl = 0.0;
                for (int i = sc.itemPtr[c]; i < sc.itemPtr[c + 1]; ++i) {
                    const int4 v = sc.items[i];
                    const int NXn = (int)sp.nX + 1, NYn = (int)sp.nY + 1, NZn = (int)sp.nZ + 1;
                    const int nyv = std::min(sc.rows, NYn - v.x * sc.rows), nzv = std::min(TZ, NZn - v.y * TZ);
                    wl += (double)(nyv + 1) * (nzv + 1) * (std::min(v.w - 1, NXn - 2) - std::max(v.z - 1, 0) + 1);
                }
                wmin = std::min(wmin, wl); wmax = std::max(wmax, wl);
            }
            fprintf(stderr, "row-stream schedule: rows %d, %d CTAs, %zu items, work per CTA min %.0f max %.0f\n", sc.rows, sc.nCtas, sc.items.size(), wmin, wmax);
            for (int c = 0; c < sc.nCtas && c < 6; ++c)
                for (int i = sc.itemPtr[c]; i < sc.itemPtr[c + 1]; ++i)
                    fprintf(stderr, "  cta %d: tile (%d,%d) planes [%d,%d)\n", c, sc.items[i].x, sc.items[i].y, sc.items[i].z, sc.items[i].w);
        }
    }
    a.tileRows = sp.streamRows;
    a.tilesY = 0; a.tilesZ = 0; a.chunkLen = 0; a.nChunks = 0;
    a.items = sp.streamItems; a.itemPtr = sp.streamItemPtr; a.carryG = sp.streamCarry; a.carryStride = sp.streamCarryStride;
    auto kern = rowStreamKernel<MC, TL, TZ, NPW, NTW, NGW, RP, RT, RG, RECST, true, true, true>;
    const size_t smem = (size_t)L::fixedDoubles() * sizeof(double);
    if (smem > 232448) return EWB_ERR_UNSUPPORTED;
#ifdef EWB_TIMING
    {
        const size_t nT = (size_t)sp.streamCtas * (NPW + NTW + NGW) * 8;
        if (sp.timingCount < nT) {
            if (sp.timingBuf) cudaFree(sp.timingBuf);
            cudaMalloc((void**)&sp.timingBuf, nT * sizeof(long long));
            sp.timingCount = nT;
        }
        cudaMemsetAsync(sp.timingBuf, 0, nT * sizeof(long long), st);
        a.timing = sp.timingBuf;
    }
#endif
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return EWB_ERR_CUDA;
    kern<<<(unsigned)sp.streamCtas, (NPW + NTW + NGW) * 32, smem, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? EWB_OK : EWB_ERR_CUDA;
}

}  // namespace ewb
