// C ABI of libedelweiss_b200.so (see include/edelweiss_b200.h).  Host-side plan construction
// (topology -> node adjacency -> CSR pattern, replacing DofManager._initializeVIJPattern and
// CSRGenerator.__init__ of the reference) and kernel dispatch.  No torch types, no CPU fallback:
// every compute entry point launches CUDA kernels or fails.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/edelweiss_b200.h"
#include "ewb_generic.cuh"
#include "ewb_sweep.cuh"
#include "ewb_rowpipe.cuh"
#include "ewb_solver.cuh"
#include "ewb_stream.cuh"

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};
inline void countLaunch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) return fail(EWB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define LAUNCH_CHECK()                                                                    \
    do {                                                                                  \
        countLaunch();                                                                    \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) return fail(EWB_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
    } while (0)

// Makes `device` current for the duration of an entry point and restores the caller's device afterwards (the host layer —
// torch — must not find its current device changed by a call into this library).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if (err == cudaSuccess) prev = -1;  // nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define WITH_DEVICE(dev)                                                                                      \
    DeviceGuard _guard(dev);                                                                                  \
    if (_guard.err != cudaSuccess) return fail(EWB_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(_guard.err))

// device that owns a device pointer (entry points without a plan)
int deviceOf(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        int d = 0;
        cudaGetDevice(&d);
        return d;
    }
    return a.device;
}

template <class T>
int upload(T** dst, const std::vector<T>& v) {
    CUDA_TRY(cudaMalloc((void**)dst, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) CUDA_TRY(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return EWB_OK;
}

}  // namespace

struct ewb_plan {
    int device = 0;
    int elType = 0;
    int nn = 8, ngp = 8;
    int64_t nEl = 0, nNode = 0, nnz = 0, nSlots = 0;
    // device tables
    int32_t* conn = nullptr;     // [nEl][nn]
    int64_t* adjPtr = nullptr;   // [nNode+1]  node adjacency (sorted unique neighbour nodes, incl. self)
    int32_t* adj = nullptr;      // [nSlots]
    int64_t* incPtr = nullptr;   // [nNode+1]  node -> incident (element*nn + localNode), ascending element
    int32_t* inc = nullptr;
    int maxDeg = 0;              // largest node degree (row-gather shared-memory size)
    unsigned char* gatherSlots = nullptr;  // [nInc][nn] element-to-CSR-slot map of the row gather (built at first use, maxDeg <= 255)
    int32_t* gatherOrder = nullptr;  // optional [nNode] visiting order of the row gather (ewb_plan_set_gather_order)
    int* failFlag = nullptr;     // device status word
    int* failHost = nullptr;     // pinned mirror
    double* peScratch = nullptr; // [nEl][3nn] per-element residual (generic path)
    double* vijScratch = nullptr;
    // device consumer (ewb_pcg_solve): r, z, p, q, Minv, mask [nDof] each; partial sums; scalars; pinned mirror of the scalars
    double* pcgWork = nullptr;
    double* pcgPartial = nullptr;
    double* pcgScal = nullptr;
    double* pcgScalHost = nullptr;
    int64_t pcgBlocks = 0;
    // structured (BoxGen) description
    bool isBox = false;
    int64_t nX = 0, nY = 0, nZ = 0;
    std::vector<int32_t> connHost;
    ewb::SweepPlan sweep;
    // task-stream kernel of the arbitrary-mesh path (ewb_stream.cuh): schedule built at first use
    std::vector<int32_t> elOrderHost;  // optional element processing order (ewb_plan_set_element_order)
    bool streamBuilt = false;
    int2* streamTasks = nullptr;
    int streamNTasks = 0, streamNChunks = 0;
    int32_t* streamElOrder = nullptr;
    int32_t* streamGatherNodes = nullptr;
    int32_t* streamChunkTarget = nullptr;
    int* streamSync = nullptr;
    double* streamRows = nullptr;
    int streamDiscard = 1;   // EWB_STREAM_DISCARD, read once at plan creation
    int streamEnabled = 0;   // EWB_STREAM=1 selects the task-stream kernel for 20-node hexahedra (read once at plan creation)
    int streamChunk = 256, streamDelay = 12, streamElPerTask = 2, streamNodesPerTask = 8;  // EWB_STREAM_CHUNK / _DELAY / _EPT / _NPT
    int rowsVariant = 0;      // EWB_ROWS_VARIANT
    int rowsTwoPhase = 0;     // EWB_C3D20_ROWS=1: element kernel + row gather over the row scratch as two ordinary launches
    int streamCtasPerSm = 0;  // EWB_STREAM_CTAS: cap of resident CTAs per SM (0 = occupancy limit)
    int fusedVariant = 0;  // 0 = automatic (measured best per material); 1 = first-generation sweep; else row-pipelined kernel variant (EWB_KERNEL, read once at plan creation)
};

namespace {

int elementInfo(int elType, int* nn, int* ngp) {
    switch (elType) {
        case EWB_C3D8: *nn = 8; *ngp = 8; return EWB_OK;
        case EWB_C3D8TL: *nn = 8; *ngp = 8; return EWB_OK;
        case EWB_C3D20: *nn = 20; *ngp = 27; return EWB_OK;
        case EWB_C3D8R: *nn = 8; *ngp = 1; return EWB_OK;
        case EWB_C3D8E: *nn = 8; *ngp = 27; return EWB_OK;
        case EWB_C3D20R: *nn = 20; *ngp = 8; return EWB_OK;
    }
    return fail(EWB_ERR_UNSUPPORTED, "unknown element type");
}

int materialClass(int elType, int material, const double* props, int nProps, ewb::MatParams* mp, int* mc, int* nMatState) {
    std::memset(mp, 0, sizeof(*mp));
    mp->kind = material;
    const bool tl = (elType == EWB_C3D8TL);
    if (material == EWB_MAT_LINEARELASTIC || material == EWB_MAT_VONMISES) {
        if (nProps < (material == EWB_MAT_VONMISES ? 6 : 2)) return fail(EWB_ERR_ARG, "too few material properties");
        const double E = props[0], v = props[1];
        mp->lambda = E * v / ((1.0 + v) * (1.0 - 2.0 * v));
        mp->G = E / (2.0 * (1.0 + v));
        if (material == EWB_MAT_VONMISES) {
            mp->fy0 = props[2]; mp->HLin = props[3]; mp->dfy = props[4]; mp->delta = props[5];
            *mc = tl ? ewb::MC_TLV : ewb::MC_VM; *nMatState = 1;
        } else {
            *mc = tl ? ewb::MC_TLE : ewb::MC_LE; *nMatState = 0;
        }
        return EWB_OK;
    }
    if (material >= EWB_MAT_NEOHOOKE_WA && material <= EWB_MAT_NEOHOOKE_WC) {
        if (!tl) return fail(EWB_ERR_UNSUPPORTED, "hyperelastic materials need the total-Lagrange element (element.py:333-334)");
        if (nProps < 2) return fail(EWB_ERR_ARG, "too few material properties");
        mp->mu = props[0]; mp->K = props[1];
        *mc = ewb::MC_NH; *nMatState = 1;
        return EWB_OK;
    }
    return fail(EWB_ERR_UNSUPPORTED, "unknown material");
}

template <int NN, int NGP, int MC, bool TL, int T, int E, int BLK>
int launchVij(ewb_plan* p, const ewb::MatParams& mp, const ewb_buffers* b, double* V, double* Pe, cudaStream_t st, int halfScratch) {
    using L = ewb::TileLayout<NN, NGP, MC>;
    auto kern = ewb::computeElementsVijKernel<NN, NGP, MC, TL, T, E, BLK>;
    const size_t smem = (size_t)E * L::PER_EL * sizeof(double);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = (p->nEl + E - 1) / E;
    kern<<<(unsigned)grid, T * E, smem, st>>>(p->nEl, p->conn, b->coords, b->U, b->dU, b->state_ref, b->state_temp, V, Pe, mp, p->failFlag, halfScratch);
    LAUNCH_CHECK();
    return EWB_OK;
}

int dispatchVij(ewb_plan* p, int mc, const ewb::MatParams& mp, const ewb_buffers* b, double* V, double* Pe, cudaStream_t st, int halfScratch = 0) {
    if (p->elType == EWB_C3D8) {
        if (mc == ewb::MC_LE) return launchVij<8, 8, ewb::MC_LE, false, 8, 16, 5>(p, mp, b, V, Pe, st, halfScratch);
        if (mc == ewb::MC_VM) return launchVij<8, 8, ewb::MC_VM, false, 8, 16, 5>(p, mp, b, V, Pe, st, halfScratch);
    } else if (p->elType == EWB_C3D8TL) {
        if (mc == ewb::MC_NH) return launchVij<8, 8, ewb::MC_NH, true, 8, 16, 5>(p, mp, b, V, Pe, st, halfScratch);
        if (mc == ewb::MC_TLE) return launchVij<8, 8, ewb::MC_TLE, true, 8, 16, 5>(p, mp, b, V, Pe, st, halfScratch);  // element.py:415-425
        if (mc == ewb::MC_TLV) return launchVij<8, 8, ewb::MC_TLV, true, 8, 16, 5>(p, mp, b, V, Pe, st, halfScratch);
    } else if (p->elType == EWB_C3D20) {
        if (mc == ewb::MC_LE) return launchVij<20, 27, ewb::MC_LE, false, 32, 4, 4>(p, mp, b, V, Pe, st, halfScratch);
        if (mc == ewb::MC_VM) return launchVij<20, 27, ewb::MC_VM, false, 32, 4, 4>(p, mp, b, V, Pe, st, halfScratch);
    } else if (p->elType == EWB_C3D8R) {
        if (mc == ewb::MC_LE) return launchVij<8, 1, ewb::MC_LE, false, 8, 16, 5>(p, mp, b, V, Pe, st, halfScratch);
        if (mc == ewb::MC_VM) return launchVij<8, 1, ewb::MC_VM, false, 8, 16, 5>(p, mp, b, V, Pe, st, halfScratch);
    } else if (p->elType == EWB_C3D8E) {
        if (mc == ewb::MC_LE) return launchVij<8, 27, ewb::MC_LE, false, 32, 4, 5>(p, mp, b, V, Pe, st, halfScratch);
        if (mc == ewb::MC_VM) return launchVij<8, 27, ewb::MC_VM, false, 32, 4, 5>(p, mp, b, V, Pe, st, halfScratch);
    } else if (p->elType == EWB_C3D20R) {
        if (mc == ewb::MC_LE) return launchVij<20, 8, ewb::MC_LE, false, 32, 4, 4>(p, mp, b, V, Pe, st, halfScratch);
        if (mc == ewb::MC_VM) return launchVij<20, 8, ewb::MC_VM, false, 32, 4, 4>(p, mp, b, V, Pe, st, halfScratch);
    }
    return fail(EWB_ERR_UNSUPPORTED, "element/material combination not implemented");
}

// ---- task-stream kernel (ewb_stream.cuh): schedule ------------------------------------------------------------------------
// Element positions (the processing order) are cut into chunks of CE; a node is ready when the last chunk holding one of its
// incident elements is complete.  Ticket order: element tasks of chunk c, then the gather tasks of the nodes that became ready
// with chunk c - D.  Every gather task therefore depends only on tasks with a lower ticket.
void streamRelease(ewb_plan* p) {
    cudaFree(p->streamTasks); cudaFree(p->streamElOrder); cudaFree(p->streamGatherNodes); cudaFree(p->streamChunkTarget); cudaFree(p->streamSync);
    p->streamTasks = nullptr; p->streamElOrder = nullptr; p->streamGatherNodes = nullptr; p->streamChunkTarget = nullptr; p->streamSync = nullptr;
    p->streamBuilt = false;
}

// Pure host part (no CUDA; also behind the test hook ewb_debug_stream_schedule): ticket list, gather order, element tasks per chunk.
int streamSchedule(int nn, int64_t nEl, int64_t nNode, const int32_t* conn, const int32_t* order, int64_t CE, int EPT, int NPT, int D,
                   std::vector<int2>& tasks, std::vector<int32_t>& gatherNodes, std::vector<int32_t>& target) {
    const int64_t nChunks = (nEl + CE - 1) / CE;
    if (nChunks >= (1 << 22)) return fail(EWB_ERR_UNSUPPORTED, "stream schedule: too many chunks");
    // chunk of every element
    std::vector<int32_t> chunkOf((size_t)nEl);
    for (int64_t pos = 0; pos < nEl; ++pos) chunkOf[order ? order[pos] : pos] = (int32_t)(pos / CE);
    // ready chunk of every node (nodes without elements: chunk 0)
    std::vector<int32_t> ready((size_t)nNode, 0);
    for (int64_t e = 0; e < nEl; ++e)
        for (int a = 0; a < nn; ++a) {
            int32_t& r = ready[conn[e * nn + a]];
            r = std::max(r, chunkOf[e]);
        }
    // nodes by ready chunk, ascending node inside a chunk (counting sort)
    std::vector<int64_t> gPtr((size_t)nChunks + 1, 0);
    for (int64_t n = 0; n < nNode; ++n) gPtr[ready[n] + 1]++;
    for (int64_t c = 0; c < nChunks; ++c) gPtr[c + 1] += gPtr[c];
    gatherNodes.assign((size_t)nNode, 0);
    {
        std::vector<int64_t> cur(gPtr.begin(), gPtr.end() - 1);
        for (int64_t n = 0; n < nNode; ++n) gatherNodes[cur[ready[n]]++] = (int32_t)n;
    }
    tasks.clear();
    tasks.reserve((size_t)(nEl / EPT + nNode / NPT + 2 * nChunks + 16));
    target.assign((size_t)nChunks, 0);
    auto pushGather = [&](int64_t c) {
        for (int64_t g = gPtr[c]; g < gPtr[c + 1]; g += NPT) {
            const int cnt = (int)std::min<int64_t>(NPT, gPtr[c + 1] - g);
            tasks.push_back(make_int2((int)((c << 8) | (cnt << 1) | 1), (int)g));
        }
    };
    for (int64_t c = 0; c < nChunks; ++c) {
        const int64_t p0 = c * CE, p1 = std::min(nEl, p0 + CE);
        for (int64_t q = p0; q < p1; q += EPT) {
            const int cnt = (int)std::min<int64_t>(EPT, p1 - q);
            tasks.push_back(make_int2((int)((c << 8) | (cnt << 1)), (int)q));
            target[c]++;
        }
        if (c - D >= 0) pushGather(c - D);
    }
    for (int64_t c = std::max<int64_t>(0, nChunks - D); c < nChunks; ++c) pushGather(c);
    if (tasks.size() >= ((size_t)1 << 30)) return fail(EWB_ERR_UNSUPPORTED, "stream schedule: too many tasks");
    return EWB_OK;
}

int streamBuild(ewb_plan* p) {
    streamRelease(p);
    const bool ordered = !p->elOrderHost.empty();
    std::vector<int2> tasks;
    std::vector<int32_t> gatherNodes, target;
    if (int rc = streamSchedule(p->nn, p->nEl, p->nNode, p->connHost.data(), ordered ? p->elOrderHost.data() : nullptr, p->streamChunk,
                                p->streamElPerTask, p->streamNodesPerTask, p->streamDelay, tasks, gatherNodes, target))
        return rc;
    const int64_t nChunks = (int64_t)target.size();
    p->streamNTasks = (int)tasks.size();
    p->streamNChunks = (int)nChunks;
    int rc = EWB_OK;
    if ((rc = upload(&p->streamTasks, tasks)) || (rc = upload(&p->streamGatherNodes, gatherNodes)) || (rc = upload(&p->streamChunkTarget, target))) return rc;
    if (ordered && (rc = upload(&p->streamElOrder, p->elOrderHost))) return rc;
    CUDA_TRY(cudaMalloc((void**)&p->streamSync, (size_t)(nChunks + 2) * sizeof(int)));
    p->streamBuilt = true;
    return EWB_OK;
}

template <int NGP, int MC, int MINB>
int streamLaunchT(ewb_plan* p, const ewb::StreamArgs& a, const ewb::MatParams& mp, cudaStream_t st) {
    constexpr int WARPS = 4;
    using L = ewb::TileLayout<20, NGP, MC>;
    int stride = std::max<int>(L::PER_EL, 9 * p->maxDeg);
    stride += ((8 - stride % 16) + 16) % 16;  // == 8 (mod 16) doubles, like TileLayout::PER_EL (bank spread of the warps' images)
    const size_t smem = (size_t)WARPS * stride * sizeof(double);
    auto kern = ewb::streamKernel<20, NGP, MC, WARPS, MINB>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int perSm = 0, nSm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, WARPS * 32, smem));
    CUDA_TRY(cudaDeviceGetAttribute(&nSm, cudaDevAttrMultiProcessorCount, p->device));
    if (perSm < 1) return fail(EWB_ERR_CUDA, "stream kernel does not fit on an SM");
    const int64_t want = ((int64_t)a.nTasks + WARPS - 1) / WARPS;
    if (p->streamCtasPerSm > 0) perSm = std::min(perSm, p->streamCtasPerSm);
    const unsigned grid = (unsigned)std::min<int64_t>((int64_t)perSm * nSm, std::max<int64_t>(want, 1));
    CUDA_TRY(cudaMemsetAsync(p->streamSync, 0, (size_t)(p->streamNChunks + 2) * sizeof(int), st));
    kern<<<grid, WARPS * 32, smem, st>>>(a, mp, stride);
    LAUNCH_CHECK();
    return EWB_OK;
}

// the two task bodies of the stream kernel as two ordinary launches over the row scratch (EWB_C3D20_ROWS=1)
template <int NGP, int MC, int MINB>
int rowsLaunchT(ewb_plan* p, const ewb::StreamArgs& a, const ewb::MatParams& mp, cudaStream_t st) {
    constexpr int WARPS = 4, GW = 8, NPT = 4;
    using L = ewb::TileLayout<20, NGP, MC>;
    const int stride = L::PER_EL;
    const size_t smemE = (size_t)WARPS * stride * sizeof(double);
    auto kernE = ewb::rowElementsKernel<20, NGP, MC, WARPS, MINB>;
    CUDA_TRY(cudaFuncSetAttribute(kernE, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemE));
    kernE<<<(unsigned)((p->nEl + WARPS - 1) / WARPS), WARPS * 32, smemE, st>>>(a, mp, stride);
    LAUNCH_CHECK();
    const int64_t warpsNeeded = (p->nNode + NPT - 1) / NPT;
    auto launchG = [&](auto kernG, int gw) -> int {
        const size_t smemG = (size_t)gw * 9 * p->maxDeg * sizeof(double);
        CUDA_TRY(cudaFuncSetAttribute(kernG, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemG));
        kernG<<<(unsigned)((warpsNeeded + gw - 1) / gw), gw * 32, smemG, st>>>(a, p->maxDeg, NPT);
        LAUNCH_CHECK();
        return EWB_OK;
    };
    // rows in flight per lane x occupancy (EWB_ROWS_VARIANT, tuning knob): 0: 4 rows, 24 warps per SM; 1: 8 rows, 8 warps; 2: 8 rows, 12 warps
    if (p->rowsVariant == 1) return launchG(ewb::rowGatherKernel<20, GW, 8, 1>, GW);
    if (p->rowsVariant == 2) return launchG(ewb::rowGatherKernel<20, 4, 8, 3>, 4);
    return launchG(ewb::rowGatherKernel<20, GW, 4, 3>, GW);
}

// ewb_assemble for 20-node hexahedra on the row scratch: task-stream kernel (EWB_STREAM=1) or its two task bodies as two launches
// (EWB_C3D20_ROWS=1); EWB_ERR_UNSUPPORTED = use the half-block two-phase path
int streamAssemble(ewb_plan* p, int mc, const ewb::MatParams& mp, const ewb_buffers* b, int flags, cudaStream_t st) {
    if (p->nn != 20 || !(p->streamEnabled || p->rowsTwoPhase) || p->maxDeg > 255) return EWB_ERR_UNSUPPORTED;
    if (mc != ewb::MC_LE && mc != ewb::MC_VM) return EWB_ERR_UNSUPPORTED;
    if (p->streamEnabled && !p->streamBuilt)
        if (int rc = streamBuild(p)) return rc;
    using RL = ewb::RowLayout<20>;
    if (!p->streamRows) CUDA_TRY(cudaMalloc((void**)&p->streamRows, (size_t)p->nEl * RL::SE * sizeof(double)));
    if (!p->gatherSlots) {
        const int64_t nInc = p->nEl * p->nn;
        CUDA_TRY(cudaMalloc((void**)&p->gatherSlots, (size_t)nInc * p->nn));
        const unsigned g = (unsigned)((p->nNode + 7) / 8);
        ewb::gatherSlotKernel<20><<<g, 256, 0, st>>>(p->nNode, p->adjPtr, p->adj, p->incPtr, p->inc, p->conn, p->gatherSlots);
        LAUNCH_CHECK();
    }
    ewb::StreamArgs a;
    a.nEl = p->nEl; a.nNode = p->nNode; a.conn = p->conn; a.coords = b->coords; a.U = b->U; a.dU = b->dU;
    a.stateRef = b->state_ref; a.stateTemp = b->state_temp; a.rows = p->streamRows;
    a.adjPtr = p->adjPtr; a.incPtr = p->incPtr; a.inc = p->inc; a.slotTab = p->gatherSlots;
    a.data = b->csr_data; a.P = b->P; a.F = b->F;
    a.tasks = p->streamTasks; a.nTasks = p->streamNTasks; a.elOrder = p->streamElOrder; a.gatherNodes = p->streamGatherNodes;
    a.chunkTarget = p->streamChunkTarget; a.nChunks = p->streamNChunks; a.sync = p->streamSync; a.failFlag = p->failFlag;
    a.accumulate = (flags & EWB_FLAG_ACCUMULATE_PF) ? 1 : 0;
    a.discard = p->streamDiscard;
    a.wantK = (flags & EWB_FLAG_NO_STIFFNESS) ? 0 : 1;
    if (!p->streamEnabled) {  // two ordinary launches
        a.discard = 0;
        a.tasks = nullptr; a.nTasks = 0; a.elOrder = nullptr; a.gatherNodes = nullptr; a.chunkTarget = nullptr; a.nChunks = 0; a.sync = nullptr;
        if (p->ngp == 27) return mc == ewb::MC_LE ? rowsLaunchT<27, ewb::MC_LE, 3>(p, a, mp, st) : rowsLaunchT<27, ewb::MC_VM, 1>(p, a, mp, st);
        if (p->ngp == 8) return mc == ewb::MC_LE ? rowsLaunchT<8, ewb::MC_LE, 3>(p, a, mp, st) : rowsLaunchT<8, ewb::MC_VM, 2>(p, a, mp, st);
        return EWB_ERR_UNSUPPORTED;
    }
    if (p->ngp == 27) {
        if (mc == ewb::MC_LE) return streamLaunchT<27, ewb::MC_LE, 3>(p, a, mp, st);
        return streamLaunchT<27, ewb::MC_VM, 1>(p, a, mp, st);
    }
    if (p->ngp == 8) {
        if (mc == ewb::MC_LE) return streamLaunchT<8, ewb::MC_LE, 3>(p, a, mp, st);
        return streamLaunchT<8, ewb::MC_VM, 2>(p, a, mp, st);
    }
    return EWB_ERR_UNSUPPORTED;
}

int envInt(const char* name, int dflt, int lo, int hi) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return std::min(hi, std::max(lo, atoi(v)));
}

}  // namespace

extern "C" {

const char* ewb_last_error(void) { return g_err.c_str(); }
int ewb_version(void) { return 100; }
int64_t ewb_launch_count(void) { return g_launches.load(); }

int ewb_plan_create(ewb_plan** out, int el_type, int64_t n_el, int64_t n_node, const int32_t* conn_host, int device) {
    if (!out || !conn_host || n_el <= 0 || n_node <= 0) return fail(EWB_ERR_ARG, "ewb_plan_create: bad arguments");
    int nn, ngp;
    if (int rc = elementInfo(el_type, &nn, &ngp)) return rc;
    if (n_el * nn >= (int64_t)1 << 31) return fail(EWB_ERR_UNSUPPORTED, "too many element nodes for int32 incidence");
    WITH_DEVICE(device);
    auto* p = new ewb_plan();
    p->device = device; p->elType = el_type; p->nn = nn; p->ngp = ngp; p->nEl = n_el; p->nNode = n_node;
    p->connHost.assign(conn_host, conn_host + n_el * nn);
    for (int64_t i = 0; i < n_el * nn; ++i)
        if (conn_host[i] < 0 || conn_host[i] >= n_node) { delete p; return fail(EWB_ERR_ARG, "connectivity index out of range"); }
    // collapsed elements (a node listed twice) would make two local nodes share one CSR slot in the row gather: not supported
    for (int64_t e = 0; e < n_el; ++e)
        for (int a = 1; a < nn; ++a)
            for (int b = 0; b < a; ++b)
                if (conn_host[e * nn + a] == conn_host[e * nn + b]) {
                    delete p;
                    return fail(EWB_ERR_UNSUPPORTED, "element " + std::to_string(e) + " lists a node twice (collapsed elements are not supported)");
                }

    // node -> incident (element, local node), ascending element (counting sort)
    std::vector<int64_t> incPtr(n_node + 1, 0);
    for (int64_t i = 0; i < n_el * nn; ++i) incPtr[conn_host[i] + 1]++;
    for (int64_t n = 0; n < n_node; ++n) incPtr[n + 1] += incPtr[n];
    std::vector<int32_t> inc(incPtr[n_node]);
    {
        std::vector<int64_t> cur(incPtr.begin(), incPtr.end() - 1);
        for (int64_t e = 0; e < n_el; ++e)
            for (int a = 0; a < nn; ++a) inc[cur[conn_host[e * nn + a]]++] = (int32_t)(e * nn + a);
    }
    // node adjacency: sorted unique union of the nodes of all incident elements
    std::vector<int64_t> adjPtr(n_node + 1, 0);
    std::vector<std::vector<int32_t>> chunks;
    const int64_t CH = 1 << 16;
    const int64_t nChunks = (n_node + CH - 1) / CH;
    chunks.resize(nChunks);
#pragma omp parallel for schedule(dynamic)
    for (int64_t c = 0; c < nChunks; ++c) {
        std::vector<int32_t> tmp;
        auto& outv = chunks[c];
        for (int64_t n = c * CH; n < std::min(n_node, (c + 1) * CH); ++n) {
            tmp.clear();
            for (int64_t k = incPtr[n]; k < incPtr[n + 1]; ++k) {
                const int64_t e = inc[k] / nn;
                for (int a = 0; a < nn; ++a) tmp.push_back(conn_host[e * nn + a]);
            }
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            adjPtr[n + 1] = (int64_t)tmp.size();
            outv.insert(outv.end(), tmp.begin(), tmp.end());
        }
    }
    for (int64_t n = 0; n < n_node; ++n) adjPtr[n + 1] += adjPtr[n];
    std::vector<int32_t> adj;
    adj.reserve(adjPtr[n_node]);
    for (auto& c : chunks) { adj.insert(adj.end(), c.begin(), c.end()); std::vector<int32_t>().swap(c); }
    p->nSlots = adjPtr[n_node];
    for (int64_t n = 0; n < n_node; ++n) p->maxDeg = std::max<int>(p->maxDeg, (int)(adjPtr[n + 1] - adjPtr[n]));
    p->nnz = 9 * p->nSlots;
    if (p->nnz >= (int64_t)1 << 31) { delete p; return fail(EWB_ERR_UNSUPPORTED, "nnz exceeds int32 CSR indices (csrgenerator.pyx uses C int)"); }

    int rc = EWB_OK;
    if ((rc = upload(&p->conn, p->connHost)) || (rc = upload(&p->adjPtr, adjPtr)) || (rc = upload(&p->adj, adj)) ||
        (rc = upload(&p->incPtr, incPtr)) || (rc = upload(&p->inc, inc))) {
        ewb_plan_destroy(p);
        return rc;
    }
    {
        cudaError_t e = cudaMalloc((void**)&p->failFlag, sizeof(int));
        if (e == cudaSuccess) e = cudaMemset(p->failFlag, 0, sizeof(int));
        if (e == cudaSuccess) e = cudaMallocHost((void**)&p->failHost, sizeof(int));
        if (e != cudaSuccess) {
            ewb_plan_destroy(p);
            return fail(EWB_ERR_CUDA, std::string("ewb_plan_create: ") + cudaGetErrorString(e));
        }
    }
    *p->failHost = 0;
    // tuning / A-B knobs of the task-stream kernel, read once
    p->streamEnabled = envInt("EWB_STREAM", 0, 0, 1);  // measured slower than the two-phase path on B200: opt-in
    p->streamDiscard = envInt("EWB_STREAM_DISCARD", 1, 0, 1);
    p->streamChunk = envInt("EWB_STREAM_CHUNK", 256, 1, 1 << 20);
    p->streamDelay = envInt("EWB_STREAM_DELAY", 12, 0, 1 << 10);
    p->streamElPerTask = envInt("EWB_STREAM_EPT", 2, 1, 127);
    p->streamNodesPerTask = envInt("EWB_STREAM_NPT", 8, 1, 32);
    p->streamCtasPerSm = envInt("EWB_STREAM_CTAS", 0, 0, 32);
    p->rowsTwoPhase = envInt("EWB_C3D20_ROWS", 0, 0, 1);
    p->rowsVariant = envInt("EWB_ROWS_VARIANT", 0, 0, 2);
    *out = p;
    return EWB_OK;
}

void ewb_plan_destroy(ewb_plan* p) {
    if (!p) return;
    DeviceGuard guard(p->device);
    cudaFree(p->conn); cudaFree(p->adjPtr); cudaFree(p->adj); cudaFree(p->incPtr); cudaFree(p->inc);
    cudaFree(p->failFlag); cudaFree(p->peScratch); cudaFree(p->vijScratch); cudaFree(p->gatherOrder); cudaFree(p->gatherSlots);
    cudaFree(p->pcgWork); cudaFree(p->pcgPartial); cudaFree(p->pcgScal);
    if (p->pcgScalHost) cudaFreeHost(p->pcgScalHost);
    if (p->failHost) cudaFreeHost(p->failHost);
    p->sweep.release();
    streamRelease(p);
    cudaFree(p->streamRows);
    delete p;
}

int64_t ewb_plan_nnz(const ewb_plan* p) { return p ? p->nnz : -1; }
int64_t ewb_plan_ndof(const ewb_plan* p) { return p ? 3 * p->nNode : -1; }
int ewb_plan_n_gauss(const ewb_plan* p) { return p ? p->ngp : -1; }
int ewb_plan_n_el_dof(const ewb_plan* p) { return p ? 3 * p->nn : -1; }
int ewb_plan_is_box(const ewb_plan* p) { return p && p->isBox ? 1 : 0; }

}  // extern "C"

namespace {

__global__ void csrPatternKernel(int64_t nNode, const int64_t* __restrict__ adjPtr, const int32_t* __restrict__ adj, int32_t* __restrict__ indptr,
                                 int32_t* __restrict__ indices) {
    const int64_t A = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (A >= nNode) return;
    const int64_t s0 = adjPtr[A], deg = adjPtr[A + 1] - s0, base = 9 * s0;
    for (int i = 0; i < 3; ++i) {
        indptr[3 * A + i] = (int32_t)(base + i * 3 * deg);
        for (int64_t s = 0; s < deg; ++s) {
            const int32_t B = adj[s0 + s];
            for (int j = 0; j < 3; ++j) indices[base + i * 3 * deg + 3 * s + j] = 3 * B + j;
        }
    }
    if (A == nNode - 1) indptr[3 * nNode] = (int32_t)(9 * adjPtr[nNode]);
}

// x[p] for COO pair p of element e: row = dof[p % n], col = dof[p / n]  (dofmanager.py:552-553, csrgenerator.pyx:82-98)
__global__ void slotMapKernel(int nn, int64_t e0, int64_t e1, const int32_t* __restrict__ conn, const int64_t* __restrict__ adjPtr,
                              const int32_t* __restrict__ adj, int32_t* __restrict__ x) {
    const int nd = 3 * nn;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (e1 - e0) * nd * nd;
    if (idx >= total) return;
    const int64_t e = e0 + idx / (nd * nd);
    const int p = (int)(idx % (nd * nd));
    const int r = p % nd, c = p / nd;
    const int32_t A = conn[e * nn + r / 3], B = conn[e * nn + c / 3];
    const int64_t s0 = adjPtr[A], deg = adjPtr[A + 1] - s0;
    int64_t lo = 0, hi = deg - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (adj[s0 + mid] < B) lo = mid + 1; else hi = mid;
    }
    x[idx] = (int32_t)(9 * s0 + (r % 3) * 3 * deg + 3 * lo + (c % 3));
}

__global__ void stateTransposeKernel(const double* __restrict__ src, double* __restrict__ dst, int64_t nEl, int nGp, int nState, int toSoa) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = nEl * nGp * nState;
    if (idx >= total) return;
    // idx enumerates the destination
    if (toSoa) {  // dst[c][e][gp] = src[e][gp][c]
        const int64_t c = idx / (nEl * nGp), r = idx % (nEl * nGp);
        dst[idx] = src[r * nState + c];
    } else {  // dst[e][gp][c] = src[c][e][gp]
        const int64_t r = idx / nState, c = idx % nState;
        dst[idx] = src[c * (nEl * nGp) + r];
    }
}

// one warp per CSR row of the receiver's first node plane
__global__ void interfaceAddKernel(const int32_t* __restrict__ indptr, int64_t nRows, double* __restrict__ data, const double* __restrict__ recv,
                                   double* __restrict__ P, double* __restrict__ F, const double* __restrict__ rP, const double* __restrict__ rF) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nRows) return;
    const int64_t r0 = indptr[row], len = indptr[row + 1] - r0, half = len >> 1;
    for (int64_t t = lane; t < half; t += 32) data[r0 + t] += recv[r0 + half + t];
    if (lane == 0 && P != nullptr) {
        P[row] += rP[row];
        F[row] += rF[row];
    }
}

__global__ void dirichletKernel(const int32_t* __restrict__ indptr3, const int64_t* __restrict__ adjPtr, const int32_t* __restrict__ adj,
                                double* __restrict__ data, const int32_t* __restrict__ dofs, int64_t n) {
    (void)indptr3;
    const int64_t k = blockIdx.x;
    if (k >= n) return;
    const int32_t dof = dofs[k];
    const int64_t A = dof / 3, i = dof % 3;
    const int64_t s0 = adjPtr[A], deg = adjPtr[A + 1] - s0;
    const int64_t row0 = 9 * s0 + i * 3 * deg;
    for (int64_t q = threadIdx.x; q < 3 * deg; q += blockDim.x) {
        const int32_t col = 3 * adj[s0 + q / 3] + (int32_t)(q % 3);
        data[row0 + q] = (col == dof) ? 1.0 : 0.0;
    }
}

}  // namespace

extern "C" {

int ewb_plan_csr_pattern(const ewb_plan* p, int32_t* indptr_dev, int32_t* indices_dev, void* stream) {
    if (!p || !indptr_dev || !indices_dev) return fail(EWB_ERR_ARG, "ewb_plan_csr_pattern: bad arguments");
    WITH_DEVICE(p->device);
    const int B = 128;
    csrPatternKernel<<<(unsigned)((p->nNode + B - 1) / B), B, 0, (cudaStream_t)stream>>>(p->nNode, p->adjPtr, p->adj, indptr_dev, indices_dev);
    LAUNCH_CHECK();
    return EWB_OK;
}

int ewb_plan_slot_map(const ewb_plan* p, int64_t e0, int64_t e1, int32_t* x_dev, void* stream) {
    if (!p || !x_dev || e0 < 0 || e1 > p->nEl || e0 >= e1) return fail(EWB_ERR_ARG, "ewb_plan_slot_map: bad arguments");
    WITH_DEVICE(p->device);
    const int nd = 3 * p->nn;
    const int64_t total = (e1 - e0) * nd * nd;
    const int B = 256;
    slotMapKernel<<<(unsigned)((total + B - 1) / B), B, 0, (cudaStream_t)stream>>>(p->nn, e0, e1, p->conn, p->adjPtr, p->adj, x_dev);
    LAUNCH_CHECK();
    return EWB_OK;
}

int ewb_plan_set_box(ewb_plan* p, int64_t nX, int64_t nY, int64_t nZ) {
    if (!p) return fail(EWB_ERR_ARG, "null plan");
    if (p->nn != 8) return fail(EWB_ERR_UNSUPPORTED, "structured sweep implemented for 8-node hexahedra only");
    if (nX * nY * nZ != p->nEl || (nX + 1) * (nY + 1) * (nZ + 1) != p->nNode) return fail(EWB_ERR_NOT_BOX, "box dimensions do not match the mesh");
    // verify against the generator's closed form (generators/boxgen.py:133-141,168-185)
    static const int off[8][3] = {{0, 0, 0}, {0, 0, 1}, {1, 0, 1}, {1, 0, 0}, {0, 1, 0}, {0, 1, 1}, {1, 1, 1}, {1, 1, 0}};
    const int64_t NY = nY + 1, NZ = nZ + 1;
    bool ok = true;
#pragma omp parallel for reduction(&& : ok)
    for (int64_t e = 0; e < p->nEl; ++e) {
        const int64_t ix = e / (nY * nZ), iy = (e / nZ) % nY, iz = e % nZ;
        for (int a = 0; a < 8; ++a) {
            const int64_t n = (ix + off[a][0]) * NY * NZ + (iy + off[a][1]) * NZ + (iz + off[a][2]);
            ok = ok && (p->connHost[e * 8 + a] == (int32_t)n);
        }
    }
    if (!ok) return fail(EWB_ERR_NOT_BOX, "connectivity is not BoxGen-ordered");
    p->isBox = true; p->nX = nX; p->nY = nY; p->nZ = nZ;
    if (int rc = p->sweep.build(nX, nY, nZ)) return fail(rc, "sweep plan build failed");
    if (const char* ev = getenv("EWB_KERNEL")) {  // tuning / A-B knob: "v1" or "rp<P><T><G>"
        const std::string k(ev);
        if (k == "v1") p->fusedVariant = 1;
        else if (k.rfind("rp", 0) == 0) {  // rp<P>_<T>_<G>
            int P = 4, T = 4, G = 4;
            if (sscanf(k.c_str(), "rph%d_%d_%d", &P, &T, &G) == 3) p->fusedVariant = 10000000 + 10000 * P + 100 * T + G;  // plain (J^-1) records
            else if (sscanf(k.c_str(), "rpi%d_%d_%d", &P, &T, &G) == 3) p->fusedVariant = 11000000 + 10000 * P + 100 * T + G;
            else if (sscanf(k.c_str(), "rpn%d_%d_%d", &P, &T, &G) == 3) p->fusedVariant = 8000000 + 10000 * P + 100 * T + G;  // no chaining
            else if (sscanf(k.c_str(), "rpm%d_%d_%d", &P, &T, &G) == 3) p->fusedVariant = 9000000 + 10000 * P + 100 * T + G;
            else if (sscanf(k.c_str(), "rpc%d_%d_%d", &P, &T, &G) == 3) p->fusedVariant = 5000000 + 10000 * P + 100 * T + G;
            else if (sscanf(k.c_str(), "rpd%d_%d_%d", &P, &T, &G) == 3) p->fusedVariant = 6000000 + 10000 * P + 100 * T + G;
            else if (sscanf(k.c_str(), "rpe%d_%d_%d", &P, &T, &G) == 3) p->fusedVariant = 7000000 + 10000 * P + 100 * T + G;
            else if (sscanf(k.c_str(), "rpa%d_%d_%d", &P, &T, &G) == 3) p->fusedVariant = 3000000 + 10000 * P + 100 * T + G;  // register split a
            else if (sscanf(k.c_str(), "rpb%d_%d_%d", &P, &T, &G) == 3) p->fusedVariant = 4000000 + 10000 * P + 100 * T + G;  // register split b
            else if (sscanf(k.c_str(), "rpr%d_%d_%d", &P, &T, &G) == 3) p->fusedVariant = 2000000 + 10000 * P + 100 * T + G;  // 3 record stages
            else if (sscanf(k.c_str(), "rps%d_%d_%d", &P, &T, &G) == 3) p->fusedVariant = 1000000 + 10000 * P + 100 * T + G;  // with per-role register budgets
            else if (sscanf(k.c_str(), "rp%d_%d_%d", &P, &T, &G) == 3) p->fusedVariant = 10000 * P + 100 * T + G;
        }
    }
    return EWB_OK;
}

int ewb_plan_set_gather_order(ewb_plan* p, const int32_t* order_host) {
    if (!p) return fail(EWB_ERR_ARG, "null plan");
    WITH_DEVICE(p->device);
    if (!order_host) {
        cudaFree(p->gatherOrder);
        p->gatherOrder = nullptr;
        return EWB_OK;
    }
    std::vector<char> seen((size_t)p->nNode, 0);
    for (int64_t i = 0; i < p->nNode; ++i) {
        const int32_t v = order_host[i];
        if (v < 0 || v >= p->nNode || seen[v]) return fail(EWB_ERR_ARG, "ewb_plan_set_gather_order: not a permutation of the nodes");
        seen[v] = 1;
    }
    if (!p->gatherOrder) CUDA_TRY(cudaMalloc((void**)&p->gatherOrder, (size_t)p->nNode * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpy(p->gatherOrder, order_host, (size_t)p->nNode * sizeof(int32_t), cudaMemcpyHostToDevice));
    return EWB_OK;
}

int ewb_plan_set_element_order(ewb_plan* p, const int32_t* order_host) {
    if (!p) return fail(EWB_ERR_ARG, "null plan");
    WITH_DEVICE(p->device);
    if (order_host) {
        std::vector<char> seen((size_t)p->nEl, 0);
        for (int64_t i = 0; i < p->nEl; ++i) {
            const int32_t v = order_host[i];
            if (v < 0 || v >= p->nEl || seen[v]) return fail(EWB_ERR_ARG, "ewb_plan_set_element_order: not a permutation of the elements");
            seen[v] = 1;
        }
        p->elOrderHost.assign(order_host, order_host + p->nEl);
    } else {
        p->elOrderHost.clear();
    }
    streamRelease(p);  // rebuilt at the next assembly
    return EWB_OK;
}

int64_t ewb_debug_stream_schedule(int nn, int64_t n_el, int64_t n_node, const int32_t* conn_host, const int32_t* order_host, int chunk, int delay,
                                  int el_per_task, int nodes_per_task, int32_t* tasks_out, int64_t max_tasks, int32_t* gather_nodes_out,
                                  int32_t* chunk_target_out, int64_t max_chunks) {
    if (!conn_host || !tasks_out || !gather_nodes_out || !chunk_target_out || nn < 1 || n_el < 1 || n_node < 1 || chunk < 1 || delay < 0 ||
        el_per_task < 1 || el_per_task > 127 || nodes_per_task < 1 || nodes_per_task > 32)
        return fail(EWB_ERR_ARG, "ewb_debug_stream_schedule: bad arguments");
    std::vector<int2> tasks;
    std::vector<int32_t> gatherNodes, target;
    if (int rc = streamSchedule(nn, n_el, n_node, conn_host, order_host, chunk, el_per_task, nodes_per_task, delay, tasks, gatherNodes, target)) return rc;
    if ((int64_t)tasks.size() > max_tasks || (int64_t)target.size() > max_chunks) return fail(EWB_ERR_ARG, "ewb_debug_stream_schedule: output too small");
    for (size_t t = 0; t < tasks.size(); ++t) {
        tasks_out[2 * t] = tasks[t].x;
        tasks_out[2 * t + 1] = tasks[t].y;
    }
    std::copy(gatherNodes.begin(), gatherNodes.end(), gather_nodes_out);
    std::copy(target.begin(), target.end(), chunk_target_out);
    return (int64_t)tasks.size();
}

int ewb_compute_elements_vij(ewb_plan* p, int material, const double* props, int n_props, const ewb_buffers* b, double* pe_dev, int flags,
                             void* stream) {
    if (!p || !b || !props || !pe_dev) return fail(EWB_ERR_ARG, "ewb_compute_elements_vij: bad arguments");
    WITH_DEVICE(p->device);
    ewb::MatParams mp; int mc, nms;
    if (int rc = materialClass(p->elType, material, props, n_props, &mp, &mc, &nms)) return rc;
    double* V = (flags & EWB_FLAG_NO_STIFFNESS) ? nullptr : b->vij;
    return dispatchVij(p, mc, mp, b, V, pe_dev, (cudaStream_t)stream);
}

int ewb_update_csr(ewb_plan* p, const double* vij_dev, double* csr_data_dev, void* stream) {
    if (!p || !vij_dev || !csr_data_dev) return fail(EWB_ERR_ARG, "ewb_update_csr: bad arguments");
    WITH_DEVICE(p->device);
    const int B = 128;
    const unsigned grid = (unsigned)((p->nSlots + B - 1) / B);
    if (p->nn == 8)
        ewb::updateCsrKernel<8><<<grid, B, 0, (cudaStream_t)stream>>>(p->nNode, p->adjPtr, p->adj, p->incPtr, p->inc, p->conn, vij_dev, csr_data_dev, p->nSlots);
    else
        ewb::updateCsrKernel<20><<<grid, B, 0, (cudaStream_t)stream>>>(p->nNode, p->adjPtr, p->adj, p->incPtr, p->inc, p->conn, vij_dev, csr_data_dev, p->nSlots);
    LAUNCH_CHECK();
    return EWB_OK;
}

int ewb_assemble(ewb_plan* p, int material, const double* props, int n_props, const ewb_buffers* b, const double time[2], double dT, int flags,
                 void* stream) {
    (void)time; (void)dT;  // the scoped materials are rate independent (element.py:332 passes them through)
    if (!p || !b || !props) return fail(EWB_ERR_ARG, "ewb_assemble: bad arguments");
    if (!b->coords || !b->U || !b->dU || !b->state_ref || !b->state_temp || !b->P || !b->F) return fail(EWB_ERR_ARG, "ewb_assemble: null buffer");
    const bool wantK = !(flags & EWB_FLAG_NO_STIFFNESS);
    if (wantK && !b->csr_data) return fail(EWB_ERR_ARG, "ewb_assemble: csr_data is null");
    WITH_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    ewb::MatParams mp; int mc, nms;
    if (int rc = materialClass(p->elType, material, props, n_props, &mp, &mc, &nms)) return rc;

    if (p->isBox && !(flags & EWB_FLAG_FORCE_GENERIC) && b->vij == nullptr) {
        if (!p->sweep.indexable()) return fail(EWB_ERR_UNSUPPORTED, "ewb_assemble: the fused BoxGen kernels index nodes with int32 (3 * nodes < 2^31); use EWB_FLAG_FORCE_GENERIC");
        // automatic choice (B200, round 2, Melem/s at the BASELINE sizes): the row-pipelined kernel for every material — linear elasticity with
        // y-chaining (648 vs 551 for the first-generation sweep), von Mises without chaining and with the tile-transpose identity (360 vs 333),
        // Neo-Hooke without chaining (442 vs 413); the role register budgets differ (P / T / G: 152/128/104, 152/136/88, 144/136/96)
        const int variant = p->fusedVariant != 0 ? p->fusedVariant : (mc == ewb::MC_LE ? 4040804 : (mc == ewb::MC_VM ? 9040804 : 8040804));
        const bool v1 = (flags & EWB_FLAG_SWEEP_V1) || variant == 1;
        const int rc = v1 ? p->sweep.launchV1(p->elType, mc, mp, b, p->failFlag, flags, st)
                          : ewb::launchRowPipeAny(p->sweep, variant, p->elType, mc, mp, b, p->failFlag, flags, st);
        if (rc == EWB_OK) {
            countLaunch();
            return EWB_OK;
        }
        if (rc != EWB_ERR_UNSUPPORTED) return fail(rc, std::string("fused sweep launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    }

    if (p->sweep.peerData) return fail(EWB_ERR_UNSUPPORTED, "ewb_assemble: peer interface buffers need the fused sweep path");
    // 20-node hexahedra: element loop and row gather as warp tasks of one persistent kernel (ewb_stream.cuh)
    if (b->vij == nullptr && !(flags & EWB_FLAG_TWO_PHASE)) {
        const int rc = streamAssemble(p, mc, mp, b, flags, st);
        if (rc != EWB_ERR_UNSUPPORTED) return rc;
    }
    // generic two-phase path
    const int nd = 3 * p->nn;
    if (!p->peScratch) CUDA_TRY(cudaMalloc((void**)&p->peScratch, (size_t)p->nEl * nd * sizeof(double)));
    double* V = nullptr;
    if (wantK) {
        V = b->vij;
        if (!V) {
            // internal scratch: half-block layout (ewb_generic.cuh HalfLayout), 9 (NN/2 + 1) doubles (+ pad) per element node
            const size_t se = (size_t)(p->nn == 8 ? ewb::HalfLayout<8>::SE : ewb::HalfLayout<20>::SE);
            if (!p->vijScratch) CUDA_TRY(cudaMalloc((void**)&p->vijScratch, (size_t)p->nEl * se * sizeof(double)));
            V = p->vijScratch;
        }
    }
    const bool internalV = wantK && V == p->vijScratch;  // nobody else reads it: half-block scratch (HalfLayout)
    if (int rc = dispatchVij(p, mc, mp, b, V, p->peScratch, st, internalV ? 1 : 0)) return rc;
    {
        const int B = 256;
        const unsigned grid = (unsigned)((3 * p->nNode + B - 1) / B);
        const int acc = (flags & EWB_FLAG_ACCUMULATE_PF) ? 1 : 0;
        if (p->nn == 8) ewb::gatherResidualKernel<8><<<grid, B, 0, st>>>(p->nNode, p->incPtr, p->inc, p->peScratch, b->P, b->F, acc);
        else ewb::gatherResidualKernel<20><<<grid, B, 0, st>>>(p->nNode, p->incPtr, p->inc, p->peScratch, b->P, b->F, acc);
        LAUNCH_CHECK();
    }
    if (wantK && internalV) {
        constexpr int W = 8;
        if (!p->gatherSlots && p->maxDeg <= 255) {
            const int64_t nInc = p->nEl * p->nn;
            CUDA_TRY(cudaMalloc((void**)&p->gatherSlots, (size_t)nInc * p->nn));
            const unsigned g = (unsigned)((p->nNode + 7) / 8);
            if (p->nn == 8) ewb::gatherSlotKernel<8><<<g, 256, 0, st>>>(p->nNode, p->adjPtr, p->adj, p->incPtr, p->inc, p->conn, p->gatherSlots);
            else ewb::gatherSlotKernel<20><<<g, 256, 0, st>>>(p->nNode, p->adjPtr, p->adj, p->incPtr, p->inc, p->conn, p->gatherSlots);
            LAUNCH_CHECK();
        }
        const size_t smem = (size_t)W * 9 * p->maxDeg * sizeof(double);
        const unsigned grid = (unsigned)((p->nNode + W - 1) / W);
        if (p->nn == 8) {
            CUDA_TRY(cudaFuncSetAttribute(ewb::rowGatherHalfKernel<8, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ewb::rowGatherHalfKernel<8, W><<<grid, W * 32, smem, st>>>(p->nNode, p->adjPtr, p->adj, p->incPtr, p->inc, p->conn, V, b->csr_data, p->maxDeg, p->gatherOrder, p->gatherSlots);
        } else {
            CUDA_TRY(cudaFuncSetAttribute(ewb::rowGatherHalfKernel<20, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ewb::rowGatherHalfKernel<20, W><<<grid, W * 32, smem, st>>>(p->nNode, p->adjPtr, p->adj, p->incPtr, p->inc, p->conn, V, b->csr_data, p->maxDeg, p->gatherOrder, p->gatherSlots);
        }
        LAUNCH_CHECK();
        return EWB_OK;
    }
    if (wantK) return ewb_update_csr(p, V, b->csr_data, stream);
    return EWB_OK;
}

// ---- pipelined host I/O: the x-chunks of the row-pipelined kernel as separate launches -------------------------------------
namespace {
// variant of the automatic choice in ewb_assemble; 0 when the plan / material does not run the row-pipelined kernel
int rowPipeVariantOf(const ewb_plan* p, int mc, int flags) {
    if (!p->isBox || (flags & (EWB_FLAG_FORCE_GENERIC | EWB_FLAG_SWEEP_V1)) || !p->sweep.indexable()) return 0;
    const int variant = p->fusedVariant != 0 ? p->fusedVariant : (mc == ewb::MC_LE ? 4040804 : (mc == ewb::MC_VM ? 9040804 : 8040804));
    return variant == 1 ? 0 : variant;
}
}  // namespace

int ewb_plan_x_chunks(ewb_plan* p, int material, const double* props, int n_props, int flags, int32_t* bounds_out, int max_bounds) {
    if (!p || !props || !bounds_out || max_bounds < 2) return fail(EWB_ERR_ARG, "ewb_plan_x_chunks: bad arguments");
    WITH_DEVICE(p->device);
    ewb::MatParams mp; int mc, nms;
    if (int rc = materialClass(p->elType, material, props, n_props, &mp, &mc, &nms)) return rc;
    const int variant = rowPipeVariantOf(p, mc, flags);
    if (!variant) return 0;
    int tiling[2] = {0, 0};
    ewb_buffers dummy;
    std::memset(&dummy, 0, sizeof(dummy));
    p->sweep.tilingOut = tiling;
    const int rc = ewb::launchRowPipeAny(p->sweep, variant, p->elType, mc, mp, &dummy, p->failFlag, flags, nullptr);
    p->sweep.tilingOut = nullptr;
    if (rc == EWB_ERR_UNSUPPORTED) return 0;
    if (rc != EWB_OK) return fail(rc, "ewb_plan_x_chunks: tiling query failed");
    const int n = tiling[1];
    if (n + 1 > max_bounds) return fail(EWB_ERR_ARG, "ewb_plan_x_chunks: bounds_out too small");
    for (int c = 0; c <= n; ++c) bounds_out[c] = (int32_t)std::min<int64_t>((int64_t)c * tiling[0], p->nX + 1);
    return n;
}

int ewb_assemble_chunks(ewb_plan* p, int material, const double* props, int n_props, const ewb_buffers* b, int flags, int chunk_begin, int chunk_end,
                        void* stream) {
    if (!p || !b || !props) return fail(EWB_ERR_ARG, "ewb_assemble_chunks: bad arguments");
    if (!b->coords || !b->U || !b->dU || !b->state_ref || !b->state_temp || !b->P || !b->F) return fail(EWB_ERR_ARG, "ewb_assemble_chunks: null buffer");
    if (!(flags & EWB_FLAG_NO_STIFFNESS) && !b->csr_data) return fail(EWB_ERR_ARG, "ewb_assemble_chunks: csr_data is null");
    WITH_DEVICE(p->device);
    ewb::MatParams mp; int mc, nms;
    if (int rc = materialClass(p->elType, material, props, n_props, &mp, &mc, &nms)) return rc;
    const int variant = rowPipeVariantOf(p, mc, flags);
    if (!variant || b->vij != nullptr) return fail(EWB_ERR_UNSUPPORTED, "ewb_assemble_chunks: needs a BoxGen plan on the row-pipelined kernel (ewb_plan_x_chunks > 0)");
    p->sweep.chunkBegin = chunk_begin;
    p->sweep.chunkEnd = chunk_end;
    const int rc = ewb::launchRowPipeAny(p->sweep, variant, p->elType, mc, mp, b, p->failFlag, flags, (cudaStream_t)stream);
    p->sweep.chunkBegin = 0;
    p->sweep.chunkEnd = -1;
    if (rc == EWB_ERR_ARG) return fail(rc, "ewb_assemble_chunks: chunk range out of bounds");
    if (rc != EWB_OK) return fail(rc, std::string("ewb_assemble_chunks: launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    countLaunch();
    return EWB_OK;
}

// Pin a caller-owned host array in place (the solver's dU lives for a whole step): its upload then needs no staging copy.
int ewb_host_register(void* ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return fail(EWB_ERR_ARG, "ewb_host_register: bad arguments");
    const cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();  // not sticky, but it must not show up in the next launch check
        return fail(EWB_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
    }
    return EWB_OK;
}

int ewb_host_unregister(void* ptr) {
    if (!ptr) return fail(EWB_ERR_ARG, "ewb_host_unregister: null pointer");
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(EWB_ERR_CUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(e));
    }
    return EWB_OK;
}

int ewb_body_force(ewb_plan* p, const double* coords_dev, const double load_host[3], double* pext_dev, void* stream) {
    if (!p || !coords_dev || !load_host || !pext_dev) return fail(EWB_ERR_ARG, "ewb_body_force: bad arguments");
    WITH_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int nd = 3 * p->nn;
    if (!p->peScratch) CUDA_TRY(cudaMalloc((void**)&p->peScratch, (size_t)p->nEl * nd * sizeof(double)));
    const int B = 128;
    const unsigned grid = (unsigned)((p->nEl + B - 1) / B);
#define EWB_BF(NN_, NGP_) ewb::bodyForceKernel<NN_, NGP_><<<grid, B, 0, st>>>(p->nEl, p->conn, coords_dev, load_host[0], load_host[1], load_host[2], p->peScratch)
    if (p->nn == 8 && p->ngp == 8) EWB_BF(8, 8);
    else if (p->nn == 8 && p->ngp == 1) EWB_BF(8, 1);
    else if (p->nn == 8 && p->ngp == 27) EWB_BF(8, 27);
    else if (p->nn == 20 && p->ngp == 27) EWB_BF(20, 27);
    else if (p->nn == 20 && p->ngp == 8) EWB_BF(20, 8);
    else return fail(EWB_ERR_UNSUPPORTED, "ewb_body_force: element type not implemented");
#undef EWB_BF
    LAUNCH_CHECK();
    const unsigned g2 = (unsigned)((3 * p->nNode + 255) / 256);
    if (p->nn == 8) ewb::gatherLoadKernel<8><<<g2, 256, 0, st>>>(p->nNode, p->incPtr, p->inc, p->peScratch, pext_dev);
    else ewb::gatherLoadKernel<20><<<g2, 256, 0, st>>>(p->nNode, p->incPtr, p->inc, p->peScratch, pext_dev);
    LAUNCH_CHECK();
    return EWB_OK;
}

int ewb_poll_status(ewb_plan* p, void* stream, double* pNewDT) {
    if (!p) return fail(EWB_ERR_ARG, "null plan");
    WITH_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(p->failHost, p->failFlag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemsetAsync(p->failFlag, 0, sizeof(int), st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (*p->failHost & 4) return fail(EWB_ERR_CUDA, "internal error: sweep kernel flag wait timed out (ordering bug)");
    if (*p->failHost & 8) return fail(EWB_ERR_UNSUPPORTED, "element with non-positive Jacobian determinant (fused sweep needs w detJ > 0)");
    if (*p->failHost & 1) {
        if (pNewDT) *pNewDT = 0.5;
        g_err = "Von Mises Newton failed.";
        return EWB_CUTBACK;
    }
    if (pNewDT) *pNewDT = 1.0;
    return EWB_OK;
}

#ifdef EWB_TIMING
/* debug builds only: per-warp phase cycle counters of the last sweep launch, [nCTA][NW][8] */
int64_t ewb_debug_timing(ewb_plan* p, long long* out_host, int64_t n) {
    if (!p || !p->sweep.timingBuf) return -1;
    const int64_t m = std::min<int64_t>(n, (int64_t)p->sweep.timingCount);
    cudaDeviceSynchronize();
    cudaMemcpy(out_host, p->sweep.timingBuf, m * sizeof(long long), cudaMemcpyDeviceToHost);
    return m;
}
#endif

int ewb_state_to_soa(const double* aos, double* soa, int64_t nEl, int nGp, int nState, void* stream) {
    if (!aos || !soa) return fail(EWB_ERR_ARG, "null state buffer");
    WITH_DEVICE(deviceOf(soa));
    const int64_t total = nEl * nGp * nState;
    stateTransposeKernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(aos, soa, nEl, nGp, nState, 1);
    LAUNCH_CHECK();
    return EWB_OK;
}

int ewb_state_to_aos(const double* soa, double* aos, int64_t nEl, int nGp, int nState, void* stream) {
    if (!aos || !soa) return fail(EWB_ERR_ARG, "null state buffer");
    WITH_DEVICE(deviceOf(soa));
    const int64_t total = nEl * nGp * nState;
    stateTransposeKernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(soa, aos, nEl, nGp, nState, 0);
    LAUNCH_CHECK();
    return EWB_OK;
}

int ewb_plan_set_peer(ewb_plan* p, double* peer_rows, double* peer_P, double* peer_F) {
    if (!p) return fail(EWB_ERR_ARG, "ewb_plan_set_peer: null plan");
    const bool any = peer_rows || peer_P || peer_F;
    if (any && !(peer_rows && peer_P && peer_F)) return fail(EWB_ERR_ARG, "ewb_plan_set_peer: all three peer buffers or none");
    if (any && !p->isBox) return fail(EWB_ERR_UNSUPPORTED, "ewb_plan_set_peer: structured (BoxGen) plans only");
    p->sweep.peerData = peer_rows; p->sweep.peerP = peer_P; p->sweep.peerF = peer_F;
    return EWB_OK;
}

int ewb_plan_status_ptr(ewb_plan* p, void** out) {
    if (!p || !out) return fail(EWB_ERR_ARG, "ewb_plan_status_ptr: bad arguments");
    *out = p->failFlag;
    return EWB_OK;
}

int ewb_peer_alloc(int64_t bytes, void** ptr_out, unsigned char handle_out[EWB_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == EWB_IPC_HANDLE_BYTES, "IPC handle size");
    if (bytes <= 0 || !ptr_out || !handle_out) return fail(EWB_ERR_ARG, "ewb_peer_alloc: bad arguments");
    void* ptr = nullptr;
    CUDA_TRY(cudaMalloc(&ptr, (size_t)bytes));
    cudaError_t e = cudaMemset(ptr, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) {
        cudaFree(ptr);
        return fail(EWB_ERR_CUDA, std::string("ewb_peer_alloc: ") + cudaGetErrorString(e));
    }
    std::memcpy(handle_out, &h, sizeof(h));
    *ptr_out = ptr;
    return EWB_OK;
}

int ewb_peer_free(void* ptr) {
    if (ptr) CUDA_TRY(cudaFree(ptr));
    return EWB_OK;
}

int ewb_peer_open(const unsigned char handle[EWB_IPC_HANDLE_BYTES], void** ptr_out) {
    if (!handle || !ptr_out) return fail(EWB_ERR_ARG, "ewb_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    CUDA_TRY(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return EWB_OK;
}

int ewb_peer_close(void* ptr) {
    if (ptr) CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return EWB_OK;
}

int ewb_interface_add(const int32_t* indptr_dev, int64_t n_rows, double* data, const double* recv, double* P, double* F, const double* rP,
                      const double* rF, void* stream) {
    if (!indptr_dev || !data || !recv || n_rows <= 0) return fail(EWB_ERR_ARG, "ewb_interface_add: bad arguments");
    WITH_DEVICE(deviceOf(data));
    const int B = 256;
    const int64_t threads = n_rows * 32;
    interfaceAddKernel<<<(unsigned)((threads + B - 1) / B), B, 0, (cudaStream_t)stream>>>(indptr_dev, n_rows, data, recv, P, F, rP, rF);
    LAUNCH_CHECK();
    return EWB_OK;
}

int ewb_apply_dirichlet_k(const ewb_plan* p, double* data, const int32_t* dofs_dev, int64_t n, void* stream) {
    if (!p || !data || (!dofs_dev && n > 0)) return fail(EWB_ERR_ARG, "ewb_apply_dirichlet_k: bad arguments");
    if (n == 0) return EWB_OK;
    WITH_DEVICE(p->device);
    dirichletKernel<<<(unsigned)n, 96, 0, (cudaStream_t)stream>>>(nullptr, p->adjPtr, p->adj, data, dofs_dev, n);
    LAUNCH_CHECK();
    return EWB_OK;
}


int ewb_apply_dirichlet_r(double* r_dev, const int32_t* dofs_dev, const double* values_dev, int64_t n, void* stream) {
    if (!r_dev || (!dofs_dev && n > 0)) return fail(EWB_ERR_ARG, "ewb_apply_dirichlet_r: bad arguments");
    if (n == 0) return EWB_OK;
    WITH_DEVICE(deviceOf(r_dev));
    ewb::dirichletRKernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(r_dev, dofs_dev, values_dev, n);
    LAUNCH_CHECK();
    return EWB_OK;
}

int ewb_spmv(const ewb_plan* p, const double* csr_data_dev, const double* x_dev, double* y_dev, void* stream) {
    if (!p || !csr_data_dev || !x_dev || !y_dev) return fail(EWB_ERR_ARG, "ewb_spmv: bad arguments");
    WITH_DEVICE(p->device);
    ewb::spmvNodeKernel<false, false><<<(unsigned)((p->nNode + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p->nNode, p->adjPtr, p->adj, csr_data_dev, x_dev, y_dev,
                                                                                                      nullptr, nullptr, nullptr);
    LAUNCH_CHECK();
    return EWB_OK;
}

int ewb_pcg_solve(ewb_plan* p, const double* A, const double* b, double* x, const int32_t* dirichlet_dofs_dev, int64_t n_dirichlet, double rel_tol,
                  int max_iter, int* iters_out, double* relres_out, void* stream) {
    if (!p || !A || !b || !x || (n_dirichlet > 0 && !dirichlet_dofs_dev) || max_iter < 0) return fail(EWB_ERR_ARG, "ewb_pcg_solve: bad arguments");
    WITH_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nDof = 3 * p->nNode;
    const int64_t nbV = (nDof + 255) / 256, nbS = (p->nNode + 7) / 8;
    const int64_t stride = std::max(nbV, nbS);
    if (!p->pcgWork) {
        CUDA_TRY(cudaMalloc((void**)&p->pcgWork, (size_t)6 * nDof * sizeof(double)));
        CUDA_TRY(cudaMalloc((void**)&p->pcgPartial, (size_t)2 * stride * sizeof(double)));
        CUDA_TRY(cudaMalloc((void**)&p->pcgScal, 8 * sizeof(double)));
        CUDA_TRY(cudaMallocHost((void**)&p->pcgScalHost, 8 * sizeof(double)));
        p->pcgBlocks = stride;
    }
    double *r = p->pcgWork, *z = r + nDof, *pp = z + nDof, *q = pp + nDof, *minv = q + nDof, *mask = minv + nDof;
    double* partial = p->pcgPartial;
    double* scal = p->pcgScal;
    const unsigned gV = (unsigned)nbV, gS = (unsigned)nbS;
    CUDA_TRY(cudaMemsetAsync(scal, 0, 8 * sizeof(double), st));
    ewb::fillKernel<<<gV, 256, 0, st>>>(mask, nDof, 1.0);
    LAUNCH_CHECK();
    if (n_dirichlet > 0) {
        ewb::maskKernel<<<(unsigned)((n_dirichlet + 255) / 256), 256, 0, st>>>(mask, nDof, dirichlet_dofs_dev, n_dirichlet);
        LAUNCH_CHECK();
    }
    ewb::pcgSetupKernel<<<gV, 256, 0, st>>>(nDof, p->adjPtr, p->adj, A, b, mask, minv, x);
    LAUNCH_CHECK();
    // q = A x0;  r = m (b - q), z = Minv r, p = z
    ewb::spmvNodeKernel<false, false><<<gS, 256, 0, st>>>(p->nNode, p->adjPtr, p->adj, A, x, q, nullptr, nullptr, nullptr);
    LAUNCH_CHECK();
    ewb::pcgInitKernel<<<gV, 256, 0, st>>>(nDof, b, q, mask, minv, r, z, pp, partial, stride);
    LAUNCH_CHECK();
    ewb::reduceKernel<<<1, 1024, 0, st>>>(partial, nbV, scal + 2, 2, stride);  // rzNew, rr
    LAUNCH_CHECK();
    ewb::pcgShiftKernel<<<1, 1, 0, st>>>(scal, 1);  // rz = rzNew, rr0 = rr
    LAUNCH_CHECK();
    int it = 0;
    double relres = 0.0;
    const int CHECK = 16;
    while (true) {
        if (it % CHECK == 0 || it >= max_iter) {
            CUDA_TRY(cudaMemcpyAsync(p->pcgScalHost, scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            const double rr = p->pcgScalHost[3], rr0 = p->pcgScalHost[4];
            relres = rr0 > 0.0 ? std::sqrt(rr / rr0) : 0.0;
            if (!(rr == rr)) return fail(EWB_ERR_CUDA, "ewb_pcg_solve: NaN residual (singular or indefinite matrix)");
            if (relres <= rel_tol || it >= max_iter) break;
        }
        ewb::spmvNodeKernel<true, true><<<gS, 256, 0, st>>>(p->nNode, p->adjPtr, p->adj, A, pp, q, mask, pp, partial);
        LAUNCH_CHECK();
        ewb::reduceKernel<<<1, 1024, 0, st>>>(partial, nbS, scal + 1, 1, stride);  // pq
        LAUNCH_CHECK();
        ewb::pcgUpdateKernel<<<gV, 256, 0, st>>>(nDof, scal, pp, q, minv, x, r, z, partial, stride);
        LAUNCH_CHECK();
        ewb::reduceKernel<<<1, 1024, 0, st>>>(partial, nbV, scal + 2, 2, stride);  // rzNew, rr
        LAUNCH_CHECK();
        ewb::pcgDirectionKernel<<<gV, 256, 0, st>>>(nDof, scal, z, pp);
        LAUNCH_CHECK();
        ewb::pcgShiftKernel<<<1, 1, 0, st>>>(scal, 0);
        LAUNCH_CHECK();
        ++it;
    }
    if (iters_out) *iters_out = it;
    if (relres_out) *relres_out = relres;
    return EWB_OK;
}

int ewb_surface_pressure(ewb_plan* p, const double* coords_dev, int64_t n_faces, const int32_t* elem_host, const int32_t* face_host, double pressure,
                         double* pext_dev, void* stream) {
    if (!p || !coords_dev || !pext_dev || n_faces < 0 || (n_faces > 0 && (!elem_host || !face_host))) return fail(EWB_ERR_ARG, "ewb_surface_pressure: bad arguments");
    if (p->nn != 8) return fail(EWB_ERR_UNSUPPORTED, "ewb_surface_pressure: 4-node faces of 8-node hexahedra only");
    if (n_faces == 0) return EWB_OK;
    WITH_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    static const int faceNodes[6][4] = {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}};
    // loaded nodes and their (face, local node) incidence in ascending order: fixed summation order
    std::vector<std::pair<int32_t, int32_t>> pairs;  // (node, 4 f + a)
    pairs.reserve((size_t)4 * n_faces);
    for (int64_t f = 0; f < n_faces; ++f) {
        if (elem_host[f] < 0 || elem_host[f] >= p->nEl || face_host[f] < 1 || face_host[f] > 6) return fail(EWB_ERR_ARG, "ewb_surface_pressure: element / face id out of range");
        for (int a = 0; a < 4; ++a) pairs.emplace_back(p->connHost[(size_t)elem_host[f] * 8 + faceNodes[face_host[f] - 1][a]], (int32_t)(4 * f + a));
    }
    std::sort(pairs.begin(), pairs.end());
    std::vector<int32_t> nodes, inc(pairs.size());
    std::vector<int64_t> incPtr;
    for (size_t k = 0; k < pairs.size(); ++k) {
        if (k == 0 || pairs[k].first != pairs[k - 1].first) { nodes.push_back(pairs[k].first); incPtr.push_back((int64_t)k); }
        inc[k] = pairs[k].second;
    }
    incPtr.push_back((int64_t)pairs.size());
    int32_t *dElem = nullptr, *dFace = nullptr, *dNodes = nullptr, *dInc = nullptr;
    int64_t* dIncPtr = nullptr;
    double* scratch = nullptr;
    std::vector<int32_t> eh(elem_host, elem_host + n_faces), fh(face_host, face_host + n_faces);
    int rc = EWB_OK;
    if ((rc = upload(&dElem, eh)) || (rc = upload(&dFace, fh)) || (rc = upload(&dNodes, nodes)) || (rc = upload(&dInc, inc)) || (rc = upload(&dIncPtr, incPtr))) {
        cudaFree(dElem); cudaFree(dFace); cudaFree(dNodes); cudaFree(dInc); cudaFree(dIncPtr);
        return rc;
    }
    cudaError_t e = cudaMalloc((void**)&scratch, (size_t)12 * n_faces * sizeof(double));
    if (e == cudaSuccess) {
        ewb::facePressureKernel<<<(unsigned)((n_faces + 127) / 128), 128, 0, st>>>(n_faces, dElem, dFace, p->conn, coords_dev, pressure, scratch);
        countLaunch();
        const int64_t nL = (int64_t)nodes.size();
        ewb::faceGatherKernel<<<(unsigned)((3 * nL + 127) / 128), 128, 0, st>>>(nL, dNodes, dIncPtr, dInc, scratch, pext_dev);
        countLaunch();
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the temporary tables are freed below
    }
    cudaFree(dElem); cudaFree(dFace); cudaFree(dNodes); cudaFree(dInc); cudaFree(dIncPtr); cudaFree(scratch);
    if (e != cudaSuccess) return fail(EWB_ERR_CUDA, std::string("ewb_surface_pressure: ") + cudaGetErrorString(e));
    return EWB_OK;
}

}  // extern "C"
