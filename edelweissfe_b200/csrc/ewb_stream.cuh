// Task-stream kernel of the arbitrary-mesh path: NIST.computeElements (nonlinearimplicitstatic.py:794-849) and
// CSRGenerator.updateCSR (csrgenerator.pyx:100-115) as ONE persistent launch for 20-node hexahedra.  OPT-IN (EWB_STREAM=1): on B200 it
// is slower than the two-phase path it was meant to replace — measurements and the reason in DESIGN.md §4.5.
//
// Idea: the two-phase path (ewb_generic.cuh) writes the element matrices to a scratch, ends the kernel and reads the scratch back
// in a second kernel: 37.9 GB of DRAM traffic per step against 11.4 GB algorithmic at 100 x 100 x 50 C3D20, and the FP64-bound
// element kernel and the LSU-bound row gather run one after the other.  Here both are *warp tasks* of one kernel:
//   * element task: one warp = one element (phase A, residual row, phase B on the FP64 tensor pipe), Ke -> row scratch;
//   * gather task : one warp = one node, sums the rows of its incident elements in ascending element order (the reference's COO order)
//                   into its three CSR rows, P and F;
// handed out in a fixed order by a ticket counter.  The order (built once per plan on the host, streamSchedule in ewb_api.cu)
// interleaves the gather of the nodes whose last incident element lies in element chunk c with the element tasks of chunk c + D.
// A row has exactly one reader (the scratch holds the full Ke, both triangles), which discards its L2 lines after the read
// (`discard.global.L2`): a row that is still in L2 when it is consumed never costs DRAM traffic.  What was measured: the discard works,
// but the L2 keeps < 1000 elements' rows while 1776 elements are in flight (one warp each, 12 warps per SM), so most rows take the
// round trip through DRAM anyway, and the gather inherits the element task's low occupancy.
// Dependencies: a gather task waits (bounded) on per-chunk completion counters; it only ever depends on tasks with a lower ticket,
// which are running on resident warps that never wait themselves — no deadlock for any grid size (tests/test_stream_schedule.py).
// Results are bitwise those of the two-phase path (same blocks, same summation order), for any element processing order.
#pragma once
#include "ewb_generic.cuh"

namespace ewb {

// Row scratch: S[e][a][i][RS] — row i of the three dof rows of local node a: columns 3 b + j (Ke[3a+i][3b+j]), column 3 NN = P_a[i]
// (the element's residual entry travels with its row), padded to a multiple of 128 bytes so that rows never share an L2 line.
template <int NN>
struct RowLayout {
    static constexpr int RS = (3 * NN + 1 + 31) / 32 * 32;  // 20 nodes: 64 doubles
    static constexpr int ROW = 3 * RS;                      // 1536 bytes
    static constexpr int SE = NN * ROW;
    static constexpr int PCOL = 3 * NN;
    static_assert((ROW * 8) % 128 == 0, "rows must be whole L2 lines");
};

template <int NN>
struct RowEmit {
    static constexpr bool HALF = false;
    template <int BLK>
    __device__ __forceinline__ void pass(int, int, const double (&)[BLK * 9 + 1]) const {}
    double* S;  // element slice [SE]
    __device__ __forceinline__ void residual(int a, const double P[3]) const {
        using RL = RowLayout<NN>;
#pragma unroll
        for (int i = 0; i < 3; ++i) S[a * RL::ROW + i * RL::RS + RL::PCOL] = P[i];
    }
    __device__ __forceinline__ void block(int a, int b, const double K[9]) const {
        using RL = RowLayout<NN>;
        double* d = S + a * RL::ROW + 3 * b;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) d[i * RL::RS + j] = K[i * 3 + j];
    }
};

// Phase B of a linear-elastic 20-node element on the FP64 tensor pipe, row scratch.  Same products, same tiles and the same
// a <= b rule as phaseBDmma20LE (the block (b, a) is the exact transpose of (a, b)); only the destination differs.
template <int NGP>
__device__ __forceinline__ void phaseBDmma20Rows(const double* sm, int lane, const MatParams& mp, double* Se) {
    constexpr int NN = 20;
    using L = TileLayout<NN, NGP, MC_LE>;
    using RL = RowLayout<NN>;
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
            const int bn = 8 * tb + r;
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
            const int b0 = 8 * tb + 2 * q;
            double K[2][9];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const double tr = mp.G * (c[0][0][t] + c[1][1][t] + c[2][2][t]);
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) K[t][i * 3 + j] = mp.lambda * c[i][j][t] + mp.G * c[j][i][t] + (i == j ? tr : 0.0);
            }
            if (aok && b0 < NN) {  // NN is even: b0 + 1 < NN as well
                double* row = Se + a * RL::ROW + 3 * b0;  // 3 b0 = 6 (4 tb + q): 16-byte aligned
                if (a <= b0) {  // both blocks of the lane: six contiguous doubles per dof row
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        double2* d = reinterpret_cast<double2*>(row + i * RL::RS);
                        d[0] = make_double2(K[0][i * 3], K[0][i * 3 + 1]);
                        d[1] = make_double2(K[0][i * 3 + 2], K[1][i * 3]);
                        d[2] = make_double2(K[1][i * 3 + 1], K[1][i * 3 + 2]);
                    }
                } else if (a == b0 + 1) {  // diagonal tile: only the second block is on or above the diagonal
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) row[i * RL::RS + 3 + j] = K[1][i * 3 + j];
                }
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int b = b0 + t;
                    if (a < b) {  // transpose: row b, columns 3 a + j
                        double* d = Se + b * RL::ROW + 3 * a;
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 3; ++j) d[i * RL::RS + j] = K[t][j * 3 + i];
                    }
                }
            }
        }
    }
}

struct StreamArgs {
    int64_t nEl, nNode;
    const int32_t* conn;
    const double* coords;
    const double* U;
    const double* dU;
    const double* stateRef;
    double* stateTemp;
    double* rows;  // row scratch [nEl][RowLayout::SE]; written and read inside the kernel: never through the read-only path
    const int64_t* adjPtr;
    const int64_t* incPtr;
    const int32_t* inc;
    const unsigned char* slotTab;
    double* data;
    double* P;
    double* F;
    const int2* tasks;  // x = key << 8 | count << 1 | kind (0 element task, key = its chunk; 1 gather task, key = last chunk it needs), y = first position
    int nTasks;
    const int32_t* elOrder;      // position -> element (nullptr: identity)
    const int32_t* gatherNodes;  // position -> node
    const int32_t* chunkTarget;  // element tasks per chunk
    int nChunks;
    int* sync;  // [0] ticket, [1] frontier (all chunks below it are complete), [2 + c] finished element tasks of chunk c; zeroed before the launch
    int* failFlag;
    int accumulate;
    int discard;
    int wantK;
};

__device__ __forceinline__ int ldAcquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void redReleaseAdd(int* p, int v) { asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void discardLine(const void* p) { asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory"); }

// All chunks <= need complete?  Advances the shared frontier on the way.  Bounded: false after ~2 s (ordering bug -> status bit 4).
__device__ __forceinline__ bool streamWait(const StreamArgs& A, int need, int lane) {
    int f = ldAcquire(A.sync + 1);
    int spins = 0;
    while (f <= need) {
        const int c = f + lane;
        const bool ok = c < A.nChunks && ldAcquire(A.sync + 2 + c) >= A.chunkTarget[c];
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        const int adv = (m == 0xffffffffu) ? 32 : __ffs(~m) - 1;
        if (adv > 0) {
            f += adv;
            __threadfence();
            if (lane == 0) atomicMax(A.sync + 1, f);
        } else {
            // give up after ~0.5 s, or at once when another warp already did (the launch is lost anyway: drain quickly)
            if (++spins > (1 << 20) || ((spins & 255) == 0 && (ldAcquire(A.failFlag) & 4))) return false;
            __nanosleep(256);
        }
    }
    return true;
}

template <int NN, int NGP, int MC>
__device__ __forceinline__ void streamElement(const StreamArgs& A, const MatParams& mp, double* sm, int64_t e, int lane) {
    using RL = RowLayout<NN>;
    if (lane < NGP) {
        const double* sp = A.stateRef + e * NGP + lane;
#pragma unroll
        for (int c = 0; c < 12 + matStateCount(MC); ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + c * (A.nEl * NGP)));
    }
    stageNodes<NN, NGP, MC, 32>(sm, A.conn + e * NN, A.coords, A.U, A.dU, lane);
    __syncwarp();
    if (lane < NGP) {
        const int64_t off = e * NGP + lane;
        gaussPoint<NN, NGP, MC, false>(sm, lane, mp, A.stateRef + off, A.stateTemp + off, A.nEl * NGP, true, A.failFlag);
    }
    __syncwarp();
    double* Se = A.rows + e * (int64_t)RL::SE;
    RowEmit<NN> emit{Se};
    if constexpr (MC == MC_LE) {
        if (lane < NN) nodeRow<NN, NGP, MC, 4>(sm, lane, mp, false, emit);  // residual row only
        if (A.wantK) phaseBDmma20Rows<NGP>(sm, lane, mp, Se);
    } else {
        constexpr int BLK = 4, NPASS = (NN / 2 + 1 + BLK - 1) / BLK;
#pragma unroll 1
        for (int it = lane; it < NN * NPASS; it += 32) {
            const int a = it % NN, ps = it / NN;
            nodeRow<NN, NGP, MC, BLK>(sm, a, mp, A.wantK != 0, emit, ps, ps + 1, ps == 0);
        }
    }
    __syncwarp();
}

// Gather task: `cnt` nodes (<= 32), one after the other: CSR rows (9 deg doubles, contiguous), P, F of each.
// buf: 9 maxDeg doubles of the warp's shared memory.  A warp has few co-resident warps to hide latency behind (the register
// and shared-memory budget is the element task's), so the task keeps its own loads in flight: the node records of the whole task
// are fetched at once, the incidence list of the next node travels while the current node is summed, and the rows of up to
// eight incident elements (24 x 128-bit loads per lane) are requested before the first one is added.
template <int NN, int B = 8>  // B: incident elements per batch (rows in flight per lane: 3 B 128-bit loads)
__device__ __forceinline__ void streamGatherTask(const StreamArgs& A, double* buf, int first, int cnt, int lane) {
    using RL = RowLayout<NN>;
    static_assert(RL::RS == 64, "lane = two columns of a row");
    const int p0 = 2 * lane, p1 = 2 * lane + 1;
    const int b0 = p0 / 3, j0 = p0 - 3 * b0, b1 = p1 / 3, j1 = p1 - 3 * b1;
    const bool isK = lane < (3 * NN) / 2 && A.wantK, isP = lane == (3 * NN) / 2;  // lane 30: column 60 = residual
    // node records of the task, one per lane
    // (32-bit: 9 * slots = nnz and the incidence count both fit an int, checked at plan creation)
    int nodeL = 0, s0L = 0, s1L = 0, k0L = 0, k1L = 0;
    if (lane < cnt) {
        nodeL = A.gatherNodes ? A.gatherNodes[first + lane] : first + lane;
        s0L = (int)A.adjPtr[nodeL];
        s1L = (int)A.adjPtr[nodeL + 1];
        k0L = (int)A.incPtr[nodeL];
        k1L = (int)A.incPtr[nodeL + 1];
    }
    int32_t eaNext = -1;  // lane u < B: incidence k0 + u of the next node
    {
        const int k0 = __shfl_sync(0xffffffffu, k0L, 0), k1 = __shfl_sync(0xffffffffu, k1L, 0);
        if (lane < B && k0 + lane < k1) eaNext = A.inc[k0 + lane];
    }
    for (int q = 0; q < cnt; ++q) {
        const int node = __shfl_sync(0xffffffffu, nodeL, q);
        const int s0 = __shfl_sync(0xffffffffu, s0L, q), s1 = __shfl_sync(0xffffffffu, s1L, q);
        const int k0 = __shfl_sync(0xffffffffu, k0L, q), k1 = __shfl_sync(0xffffffffu, k1L, q);
        const int rowLen = 3 * (s1 - s0);
        if (A.wantK)
            for (int i = lane; i < 3 * rowLen; i += 32) buf[i] = 0.0;
        double p[3] = {0, 0, 0}, f[3] = {0, 0, 0};
        if (isP && A.accumulate) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                p[i] = A.P[3 * (int64_t)node + i];
                f[i] = A.F[3 * (int64_t)node + i];
            }
        }
        __syncwarp();
        for (int kb = k0; kb < k1; kb += B) {
            int32_t eaL = eaNext;
            if (kb != k0) eaL = (lane < B && kb + lane < k1) ? A.inc[kb + lane] : -1;
            if (kb + B >= k1 && q + 1 < cnt) {  // last batch of this node: request the first batch of the next node
                const int n0 = __shfl_sync(0xffffffffu, k0L, q + 1), n1 = __shfl_sync(0xffffffffu, k1L, q + 1);
                eaNext = (lane < B && n0 + lane < n1) ? A.inc[n0 + lane] : -1;
            }
            double2 v[B][3];
            // slots of the lane's two values in the node's sorted neighbour list: RAW bytes here, arithmetic only after all loads of
            // the batch are issued (a use inside the issue loop makes every element's loads wait for the previous element's slots)
            unsigned char sb0[B], sb1[B];
            auto rowOf = [&](int u) -> const double* {
                const int32_t ea = __shfl_sync(0xffffffffu, eaL, u);
                if (ea < 0) return nullptr;
                const int64_t e = ea / NN;
                const int a = ea - (int32_t)e * NN;
                return A.rows + e * (int64_t)RL::SE + a * RL::ROW;
            };
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const double* rp = rowOf(u);
                if (rp != nullptr) {
                    const double2* src = reinterpret_cast<const double2*>(rp) + lane;
                    if (isK || isP) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) v[u][i] = __ldcg(src + i * (RL::RS / 2));
                    }
                    if (isK) {
                        sb0[u] = A.slotTab[(int64_t)(kb + u) * NN + b0];
                        sb1[u] = A.slotTab[(int64_t)(kb + u) * NN + b1];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const double* rp = rowOf(u);
                if (rp != nullptr) {  // warp-uniform
                    if (isK) {
                        const int d0 = 3 * sb0[u] + j0, d1 = 3 * sb1[u] + j1;
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            buf[i * rowLen + d0] += v[u][i].x;
                            buf[i * rowLen + d1] += v[u][i].y;
                        }
                    }
                    if (isP) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            p[i] += v[u][i].x;
                            f[i] += fabs(v[u][i].x);
                        }
                    }
                    __syncwarp();  // the next element may hit the same CSR entries from other lanes
                    if (A.discard && lane < RL::ROW * 8 / 128) discardLine(rp + lane * 16);  // the only reader of this row is done with it
                }
            }
        }
        if (A.wantK) {
            double* out = A.data + 9 * (int64_t)s0;
            for (int i = lane; i < 3 * rowLen; i += 32) __stcs(out + i, buf[i]);
        }
        if (isP) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                A.P[3 * (int64_t)node + i] = p[i];
                A.F[3 * (int64_t)node + i] = f[i];
            }
        }
        __syncwarp();
    }
}

template <int NN, int NGP, int MC, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) streamKernel(const __grid_constant__ StreamArgs A, const MatParams mp, int warpStride) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sm = smem + (size_t)warp * warpStride;  // warpStride >= max(TileLayout::PER_EL, 9 maxDeg)
    int t = 0;
    if (lane == 0) t = atomicAdd(A.sync, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    while (t < A.nTasks) {
        int tn = 0;
        if (lane == 0) tn = atomicAdd(A.sync, 1);  // the next ticket travels while this task runs
        const int2 task = A.tasks[t];
        const int cnt = (task.x >> 1) & 127, key = task.x >> 8;
        if (!(task.x & 1)) {
            for (int qq = 0; qq < cnt; ++qq) {
                const int64_t pos = (int64_t)task.y + qq;
                const int64_t e = A.elOrder ? (int64_t)A.elOrder[pos] : pos;
                streamElement<NN, NGP, MC>(A, mp, sm, e, lane);
            }
            __threadfence();
            __syncwarp();
            if (lane == 0) redReleaseAdd(A.sync + 2 + key, 1);
        } else {
            if (streamWait(A, key, lane)) {
                streamGatherTask<NN>(A, sm, task.y, cnt, lane);
            } else if (lane == 0) {
                atomicOr(A.failFlag, 4);
            }
        }
        t = __shfl_sync(0xffffffffu, tn, 0);
    }
}

// ---- the same two task bodies as two ordinary kernels ("row" two-phase path, EWB_C3D20_ROWS=1): every warp of the first launch
// computes one element into the row scratch, every warp of the second gathers `npt` consecutive nodes.  Against the half-block
// two-phase path of ewb_generic.cuh: the scratch is read once instead of twice and with whole-line loads (the gather is no longer
// LSU bound), at the price of writing both triangles (15.4 instead of 8.8 GB at 100 x 100 x 50).  Same results, bitwise.
template <int NN, int NGP, int MC, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) rowElementsKernel(const __grid_constant__ StreamArgs A, const MatParams mp, int warpStride) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t e = (int64_t)blockIdx.x * WARPS + warp;
    if (e >= A.nEl) return;
    streamElement<NN, NGP, MC>(A, mp, smem + (size_t)warp * warpStride, e, lane);
}

template <int NN, int WARPS, int B, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) rowGatherKernel(const __grid_constant__ StreamArgs A, int maxDeg, int npt) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t first = ((int64_t)blockIdx.x * WARPS + warp) * npt;
    if (first >= A.nNode) return;
    const int cnt = (int)min((int64_t)npt, A.nNode - first);
    streamGatherTask<NN, B>(A, smem + (size_t)warp * 9 * maxDeg, (int)first, cnt, lane);
}

}  // namespace ewb
