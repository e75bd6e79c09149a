/*
 * edelweiss_b200.h — C ABI of the B200-native element-loop / CSR-assembly path.
 *
 * This is the drop-in boundary for EdelweissFE's hot path.  Every entry
 * point names the reference interface it replaces (paths relative to the reference tree,
 * edelweissfe/...).  Plain pointers and sizes only; device pointers are raw CUDA device
 * addresses (e.g. torch.Tensor.data_ptr()), owned by the caller for the duration of a call.
 *
 * Conventions shared with the reference:
 *   - dof(node i, component c) = 3*i + c, i = position of the node in model.nodes order
 *     (numerics/dofmanager.py:280-292, 445-471);
 *   - element dof list is node-major (element.py:116-121 dofIndicesPermutation = identity);
 *   - CSR pattern = SciPy-canonical union pattern, int32 indptr/indices, explicit zeros kept
 *     (numerics/csrgenerator.pyx:68-77);
 *   - Ke is written row-major into the element's VIJ slice whose (I,J) are column-major, i.e.
 *     K_global[dof[j], dof[i]] += Ke[i][j] (numerics/dofmanager.py:543-555, element.py:318);
 *   - P -= Bt sigma detJ w (element.py:344); F accumulates |Pe| (nonlinearimplicitstatic.py:844).
 *
 * Gauss-point state on the device is component-major SoA:  state[c][e][gp]
 * (c < 12 + nMaterialState), the transposition of the reference's per-element
 * _stateVarsRef[gp][c] (element.py:225-236).  ewb_state_to_soa / ewb_state_to_aos convert.
 */
#ifndef EDELWEISS_B200_H
#define EDELWEISS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element formulations (provider "edelweiss"; elements/library.py:212-227, 260-275, 453-468) */
enum {
    EWB_C3D8 = 0, EWB_C3D20 = 1,
    EWB_C3D8TL = 2, /* total Lagrange: Neo-Hooke (hyperelastic branch, element.py:391-414) on every path; linear elastic /
                       von Mises (B^T C B + geometric stiffness, element.py:415-425) on the arbitrary-mesh path */
    /* integration variants of the small-strain element (elements/library.py:228-259, 276-291): arbitrary-mesh path */
    EWB_C3D8R = 3,  /* 8 nodes, 1 Gauss point (weight 8)  */
    EWB_C3D8E = 4,  /* 8 nodes, 3x3x3 Gauss points        */
    EWB_C3D20R = 5  /* 20 nodes, 2x2x2 Gauss points       */
};

/* materials (materials/linearelastic, materials/vonmises, materials/neohooke) */
enum {
    EWB_MAT_LINEARELASTIC = 0, /* props: E, nu                               (linearelastic.py:185-210) */
    EWB_MAT_VONMISES = 1,      /* props: E, nu, fy0, HLin, dfy, delta        (vonmises.py:186-254)      */
    EWB_MAT_NEOHOOKE_WA = 2,   /* props: mu, K                 (neohookepencegouformulationa.py:107-141) */
    EWB_MAT_NEOHOOKE_WB = 3,   /*                              (neohookepencegouformulationb.py:130-145) */
    EWB_MAT_NEOHOOKE_WC = 4    /*                              (neohookepencegouformulationc.py:130-143) */
};

/* status codes */
enum {
    EWB_OK = 0,
    EWB_CUTBACK = 1,       /* a material asked for a smaller increment; *pNewDT holds the factor.
                              Mirrors CutbackRequest("Von Mises Newton failed.", 0.5), vonmises.py:230-231,
                              and the pNewDT convention of marmotelement/element.pxd:97-103. */
    EWB_ERR_ARG = -1,
    EWB_ERR_CUDA = -2,
    EWB_ERR_UNSUPPORTED = -3,
    EWB_ERR_NOT_BOX = -4
};

/* assembly flags */
enum {
    EWB_FLAG_ACCUMULATE_PF = 1, /* P += , F += (reference semantics on a caller-zeroed vector); default overwrites */
    EWB_FLAG_FORCE_GENERIC = 2, /* use the two-phase VIJ path even when a structured (BoxGen) plan exists */
    EWB_FLAG_NO_STIFFNESS = 4,  /* residual / state only (P, F, stateTemp) */
    EWB_FLAG_SWEEP_V1 = 8,      /* BoxGen plans: the first-generation fused sweep (colour-ordered shared-memory accumulation) instead of the row-pipelined gather sweep */
    EWB_FLAG_TWO_PHASE = 16     /* arbitrary-mesh path, 20-node hexahedra: separate element and row-gather kernels instead of the task-stream kernel (same results, bitwise) */
};

typedef struct ewb_plan ewb_plan; /* opaque: mesh topology + CSR slot tables on one device */

/* Device buffers of one assembly call.  All pointers are device addresses. */
typedef struct ewb_buffers {
    const double* coords;    /* [nNode][3]                     Node.coordinates                          */
    const double* U;         /* [nDof]   U_np                  nonlinearimplicitstatic.py:416             */
    const double* dU;        /* [nDof]   dU                    nonlinearimplicitstatic.py:417             */
    const double* state_ref; /* [nState][nEl][nGp]  last accepted state (element.py:320 copies it)        */
    double* state_temp;      /* [nState][nEl][nGp]  updated state (element.py:346), committed by the caller */
    double* csr_data;        /* [nnz]    overwritten: CSRGenerator.updateCSR (csrgenerator.pyx:100-115)    */
    double* P;               /* [nDof]   reaction vector                                                   */
    double* F;               /* [nDof]   accumulated |flux|                                                */
    double* vij;             /* optional [nEl*nDofEl^2]: VIJSystemMatrix values, reference layout, or NULL */
} ewb_buffers;

const char* ewb_last_error(void);
int ewb_version(void);

/* ---- topology / pattern -------------------------------------------------------------------
 * Replaces DofManager._locateNodeCouplingEntitiesInDofVector + _initializeVIJPattern
 * (numerics/dofmanager.py:445-471, 522-557) and CSRGenerator.__init__ (csrgenerator.pyx:47-98).
 * conn_host: HOST int32 [nEl][nNodesPerElement], 0-based node positions. */
int ewb_plan_create(ewb_plan** plan, int el_type, int64_t n_el, int64_t n_node, const int32_t* conn_host, int device);
void ewb_plan_destroy(ewb_plan* plan);
int64_t ewb_plan_nnz(const ewb_plan* plan);
int64_t ewb_plan_ndof(const ewb_plan* plan);
int ewb_plan_n_gauss(const ewb_plan* plan);
int ewb_plan_n_el_dof(const ewb_plan* plan);
/* CSR pattern into caller-owned device buffers: indptr int32[nDof+1], indices int32[nnz]. */
int ewb_plan_csr_pattern(const ewb_plan* plan, int32_t* indptr_dev, int32_t* indices_dev, void* stream);
/* COO->CSR slot map x (csrgenerator.pyx:82-98) for elements [e0,e1): int32[(e1-e0)*nDofEl^2], device. */
int ewb_plan_slot_map(const ewb_plan* plan, int64_t e0, int64_t e1, int32_t* x_dev, void* stream);
/* Declare the mesh a BoxGen Hexa8 box (generators/boxgen.py:124-185) of nX x nY x nZ elements.
 * The connectivity is verified against the generator's closed form; enables the fused sweep kernel. */
int ewb_plan_set_box(ewb_plan* plan, int64_t nX, int64_t nY, int64_t nZ);
int ewb_plan_is_box(const ewb_plan* plan);
/* Optional locality hint for the arbitrary-mesh path: the order (a permutation of 0..n_node-1, host array) in which the row
 * gather visits the nodes.  Every block of the element scratch is read by two nodes; visiting spatially close nodes together
 * (the host layer passes a Morton order of the coordinates) lets the second read hit L2.  Results do not depend on it.
 * NULL restores the node order.  No counterpart in the reference (csrgenerator.pyx:100-115 is a serial scatter). */
int ewb_plan_set_gather_order(ewb_plan* plan, const int32_t* order_host);
/* Optional locality hint for the task-stream kernel of the arbitrary-mesh path (20-node hexahedra): the order (a permutation of
 * 0..n_el-1, host array) in which the elements are processed.  A node's CSR rows are gathered as soon as its last incident
 * element is done; processing spatially close elements together (the host layer passes a Morton order of the element
 * centroids) keeps the element matrices in L2 between the two.  Results do not depend on it (the summation order per node is
 * always ascending element index = ascending COO index, csrgenerator.pyx:100-115).  NULL restores the element order. */
int ewb_plan_set_element_order(ewb_plan* plan, const int32_t* order_host);
/* Test hook, host only (no CUDA call): the ticket list of the task-stream kernel for a connectivity.  tasks_out[2 t] =
 * key << 8 | count << 1 | kind (kind 0: element task, key = its chunk, tasks_out[2 t + 1] = first position in the processing order;
 * kind 1: gather task, key = last chunk it waits for, tasks_out[2 t + 1] = first position in gather_nodes_out), chunk_target_out[c] =
 * element tasks of chunk c.  Returns the number of tasks.  Invariant checked by tests/test_stream_schedule.py: a gather task's nodes
 * only touch elements of chunks <= key, and every element task of those chunks has a lower ticket (no deadlock for any grid size). */
int64_t ewb_debug_stream_schedule(int nn, int64_t n_el, int64_t n_node, const int32_t* conn_host, const int32_t* order_host, int chunk, int delay,
                                  int el_per_task, int nodes_per_task, int32_t* tasks_out, int64_t max_tasks, int32_t* gather_nodes_out,
                                  int32_t* chunk_target_out, int64_t max_chunks);

/* ---- the hot path ---------------------------------------------------------------------------
 * One NIST.computeElements pass + CSRGenerator.updateCSR on the device
 * (solvers/nonlinearimplicitstatic.py:794-849, 753-769).  Asynchronous on `stream`. */
int ewb_assemble(ewb_plan* plan, int material, const double* props_host, int n_props, const ewb_buffers* buf,
                 const double time[2], double dT, int flags, void* stream);
/* Pipelined host I/O for BoxGen plans.  The fused kernel cuts the box into x-chunks of node planes whose CTAs are independent
 * (a chunk re-computes the element plane below its first node plane): chunk c writes the CSR rows, P and F of the node planes
 * [bounds[c], bounds[c+1]) and reads U / dU of the node planes [bounds[c] - 1, bounds[c+1]].  Launched chunk by chunk on its own
 * streams, the upload of the next chunk's dU and the download of the previous chunk's P overlap the kernel (BoxGen numbers nodes
 * x-major, generators/boxgen.py:133-136, so a chunk's dofs are one contiguous range).  Same results as ewb_assemble, bitwise.
 * ewb_plan_x_chunks: writes the n + 1 node-plane boundaries, returns n, or 0 when this plan / material / flags combination does
 * not run the chunked kernel (use ewb_assemble).  No counterpart in the reference (its element loop is serial). */
int ewb_plan_x_chunks(ewb_plan* plan, int material, const double* props_host, int n_props, int flags, int32_t* bounds_out, int max_bounds);
int ewb_assemble_chunks(ewb_plan* plan, int material, const double* props_host, int n_props, const ewb_buffers* buf, int flags,
                        int chunk_begin, int chunk_end, void* stream);
/* Pin / unpin a caller-owned host array in place (cudaHostRegister): the solver's dU vector lives for a whole step
 * (nonlinearimplicitstatic.py:163), so the host layer registers it once and uploads from it directly instead of staging it
 * through a pinned copy.  A failure (e.g. locked-memory limit) leaves no CUDA error state behind; the caller falls back to staging. */
int ewb_host_register(void* host_ptr, int64_t bytes);
int ewb_host_unregister(void* host_ptr);
/* Synchronises `stream`, returns EWB_OK or EWB_CUTBACK (then *pNewDT = 0.5). */
int ewb_poll_status(ewb_plan* plan, void* stream, double* pNewDT);

/* ---- pieces of the reference-faithful two-phase path (exposed for parity tests) -------------- */
/* computeElements only: V (VIJ values), Pe (per-element residual [nEl][nDofEl]), stateTemp. */
int ewb_compute_elements_vij(ewb_plan* plan, int material, const double* props_host, int n_props,
                             const ewb_buffers* buf, double* pe_dev, int flags, void* stream);
/* CSRGenerator.updateCSR: data[x[p]] += V[p], summed in ascending p (deterministic gather). */
int ewb_update_csr(ewb_plan* plan, const double* vij_dev, double* csr_data_dev, void* stream);

/* ---- state layout helpers (element.py:225-236 <-> device SoA) ------------------------------- */
int ewb_state_to_soa(const double* aos_dev, double* soa_dev, int64_t n_el, int n_gp, int n_state, void* stream);
int ewb_state_to_aos(const double* soa_dev, double* aos_dev, int64_t n_el, int n_gp, int n_state, void* stream);

/* ---- NISTParallel.applyDirichletK (nonlinearimplicitstaticparallelmk2.pyx:66-109):
 * zero the CSR rows of the given dofs and put 1 on their diagonal (pattern unchanged). */
int ewb_apply_dirichlet_k(const ewb_plan* plan, double* csr_data_dev, const int32_t* dofs_dev, int64_t n, void* stream);

/* ---- device-side consumers of the assembled system --------------------------------------
 * NIST.applyDirichlet on the residual (solvers/nonlinearimplicitstatic.py:595-623): R[dofs[k]] = values[k];
 * values_dev == NULL writes zeros (the later iterations' R[dirichlet] = 0, :432-433). */
int ewb_apply_dirichlet_r(double* r_dev, const int32_t* dofs_dev, const double* values_dev, int64_t n, void* stream);
/* y = K x on the plan's CSR pattern (FP64, one warp per node).  What a device linear solver builds on; also lets tests
 * check K x against the reference's scipy matrix (csr_matrix.dot, used by linearSolve's callers). */
int ewb_spmv(const ewb_plan* plan, const double* csr_data_dev, const double* x_dev, double* y_dev, void* stream);
/* NIST.linearSolve (solvers/nonlinearimplicitstatic.py:727-751) for the system after applyDirichletK (:559-593), on the
 * device: Jacobi-preconditioned conjugate gradients on the free dofs, x[dirichlet] = b[dirichlet] (the prescribed rows are
 * identity rows).  csr_data_dev is NOT modified (rows of the Dirichlet dofs are skipped through the mask, so calling
 * ewb_apply_dirichlet_k first is allowed but not required).  Stops at |r| <= rel_tol |r0| or max_iter; synchronises `stream`.
 * Reductions have a fixed order: the result is bitwise reproducible. */
int ewb_pcg_solve(ewb_plan* plan, const double* csr_data_dev, const double* b_dev, double* x_dev, const int32_t* dirichlet_dofs_dev,
                  int64_t n_dirichlet, double rel_tol, int max_iter, int* iters_out, double* relres_out, void* stream);

/* ---- distributed surface load (NIST.computeDistributedLoads, solvers/nonlinearimplicitstatic.py:460-508) for
 * `type=pressure` on 4-node faces of 8-node hexahedra: PExt += -p int N n dA over the listed (element, Abaqus face id 1..6)
 * pairs, dead load on the reference geometry (what BASELINE config 1, testfiles/LinearElasticIsotropic/test.inp, applies
 * through the Marmot element; the reference's own Python element raises for distributed loads, element.py:255-288).
 * elem_host / face_host: HOST int32 arrays (surface definitions are set-up data).  Synchronises `stream`. */
int ewb_surface_pressure(ewb_plan* plan, const double* coords_dev, int64_t n_faces, const int32_t* elem_host, const int32_t* face_host,
                         double pressure, double* pext_dev, void* stream);

/* ---- body force element loop: NIST.computeBodyForces (solvers/nonlinearimplicitstatic.py:516-557) with
 * computeBodyForce of every element of the plan (elements/displacementelement/element.py:348-371, same in the TL element):
 * PExt[el] += sum_gp outer(N[gp], load) detJ w.  The reference's N operator node ordering (xi/eta swapped relative to the
 * derivative tables) is reproduced.  load_host: the current force vector (bodyforce.getCurrentLoad). */
int ewb_body_force(ewb_plan* plan, const double* coords_dev, const double load_host[3], double* pext_dev, void* stream);

/* ---- multi-GPU slab interface --------------------------------------------------------
 * Rank g owns the node planes [a_g, a_g+1) of a BoxGen box split along x; its local mesh also holds
 * the ghost plane a_g+1, whose partial CSR rows / P / F (contiguous tail of the local arrays) are sent
 * to rank g+1.  On the receiver the first node plane has rows of the same lengths, laid out
 * [dx=0 | dx=+1] where the sender has [dx=-1 | dx=0]: this adds the neighbour's dx=0 half onto the
 * local rows (own contribution first, neighbour second: fixed order).  The dx=-1 half stays in `recv`
 * and is the rank's lower halo block (columns = nodes of the neighbour's last owned plane).
 * indptr_dev: local CSR indptr; n_rows = 3 * nodes per plane. */
int ewb_interface_add(const int32_t* indptr_dev, int64_t n_rows, double* csr_data_dev, const double* recv_rows_dev, double* P_dev,
                      double* F_dev, const double* recv_P_dev, const double* recv_F_dev, void* stream);

/* Fused interface transfer: with peer buffers set, ewb_assemble's sweep kernel stores the ghost plane's rows / P / F
 * straight into the upper neighbour's receive buffers (peer memory over NVLink) while the rest of the slab is still being
 * computed, instead of into the local tail; no send afterwards.  The receiver orders its ewb_interface_add after the
 * sender's kernel with one small collective (the all-reduce of the status words, below).  Pass NULLs to switch back to
 * the local tail.  Structured (BoxGen) plans only. */
int ewb_plan_set_peer(ewb_plan* plan, double* peer_rows_dev, double* peer_P_dev, double* peer_F_dev);

/* Device address of the plan's status word (int32; bit 0 = cutback requested, see ewb_poll_status): a distributed caller
 * max-reduces it across ranks so that every rank takes the same cut-back decision (nonlinearimplicitstatic.py:253-262). */
int ewb_plan_status_ptr(ewb_plan* plan, void** status_dev_out);

/* Receive buffers reachable from another process' GPU (CUDA IPC).  ewb_peer_alloc: device allocation + 64-byte handle to
 * pass to the neighbour process; ewb_peer_open maps a neighbour's handle into this process (peer access enabled). */
#define EWB_IPC_HANDLE_BYTES 64
int ewb_peer_alloc(int64_t bytes, void** ptr_out, unsigned char handle_out[EWB_IPC_HANDLE_BYTES]);
int ewb_peer_free(void* ptr);
int ewb_peer_open(const unsigned char handle[EWB_IPC_HANDLE_BYTES], void** ptr_out);
int ewb_peer_close(void* ptr);

/* number of this library's kernel launches since load (bench.py's gpu_launches) */
int64_t ewb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* EDELWEISS_B200_H */
