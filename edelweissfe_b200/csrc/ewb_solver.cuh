// Device-side consumers of the assembled system: what follows computeElements in the reference's Newton
// loop (solvers/nonlinearimplicitstatic.py:419-456) without the CSR values ever leaving the GPU.
//
//   dirichletRKernel       applyDirichlet on the residual (:595-623) and the later R[dirichlet] = 0 (:432-433)
//   spmvNodeKernel         y = A x on the plan's node-block CSR (one warp per node: its three rows share the column blocks,
//                          so the pattern is read as 27 node ids instead of 243 column indices)
//   pcg* kernels           Jacobi-preconditioned conjugate gradients on the Dirichlet-modified matrix: the prescribed rows are
//                          identity rows (applyDirichletK, :559-593), so with x_D = b_D fixed the iteration runs on the free
//                          dofs only, operator v -> m (A (m v)), m = free-dof mask — the symmetric positive definite K_FF.
//                          All reductions are two-stage with a fixed order (bitwise reproducible).
//   facePressureKernel     Abaqus-style pressure on 4-node hexahedron faces (config 1 of BASELINE.json,
//                          testfiles/LinearElasticIsotropic/test.inp: `distributedload, type=pressure`); dead load on the
//                          reference geometry, gathered per node in ascending face order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ewb {

__global__ void dirichletRKernel(double* __restrict__ R, const int32_t* __restrict__ dofs, const double* __restrict__ values, int64_t n) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) R[dofs[k]] = values ? values[k] : 0.0;
}

// y[3A+i] = sum_B sum_j data[row(A,i)][3 s(B) + j] x[3B+j];  optional: y *= mask, partial[blockIdx] = sum over the block's rows of w .* y
template <bool MASK, bool DOT>
__global__ void __launch_bounds__(256) spmvNodeKernel(int64_t nNode, const int64_t* __restrict__ adjPtr, const int32_t* __restrict__ adj,
                                                      const double* __restrict__ data, const double* __restrict__ x, double* __restrict__ y,
                                                      const double* __restrict__ mask, const double* __restrict__ w, double* __restrict__ partial) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t A = (int64_t)blockIdx.x * 8 + warp;
    double dot = 0.0;
    if (A < nNode) {
        const int64_t s0 = adjPtr[A];
        const int deg = (int)(adjPtr[A + 1] - s0);
        const double* row0 = data + 9 * s0;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int s = lane; s < deg; s += 32) {
            const int64_t B = adj[s0 + s];
            const double x0 = x[3 * B], x1 = x[3 * B + 1], x2 = x[3 * B + 2];
            const double* d0 = row0 + 3 * s;
            const double* d1 = d0 + 3 * deg;
            const double* d2 = d1 + 3 * deg;
            a0 += d0[0] * x0 + d0[1] * x1 + d0[2] * x2;
            a1 += d1[0] * x0 + d1[1] * x1 + d1[2] * x2;
            a2 += d2[0] * x0 + d2[1] * x1 + d2[2] * x2;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (lane < 3) {
            double v = lane == 0 ? a0 : (lane == 1 ? a1 : a2);
            if (MASK) v *= mask[3 * A + lane];
            y[3 * A + lane] = v;
            if (DOT) dot = v * w[3 * A + lane];
        }
    }
    if (DOT) {
        // fixed-order block reduction: lanes 0..2 of every warp hold a term
        __shared__ double sh[24];
        if (lane < 3) sh[warp * 3 + lane] = dot;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int i = 0; i < 24; ++i) s += sh[i];
            partial[blockIdx.x] = s;
        }
    }
}

// out[0] = sum(partial[0..n)) in a fixed order (single block)
__global__ void __launch_bounds__(1024) reduceKernel(const double* __restrict__ partial, int64_t n, double* __restrict__ out, int nOut, int64_t stride) {
    __shared__ double sh[1024];
    for (int o = 0; o < nOut; ++o) {
        double s = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += 1024) s += partial[o * stride + i];
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int w = 512; w > 0; w >>= 1) {
            if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[o] = sh[0];
        __syncthreads();
    }
}

// scal layout (device doubles): [0] rz  [1] pq  [2] rzNew  [3] rr  [4] rr0
// set-up: mask, Minv, x = (1 - m) b
__global__ void pcgSetupKernel(int64_t nDof, const int64_t* __restrict__ adjPtr, const int32_t* __restrict__ adj, const double* __restrict__ data,
                               const double* __restrict__ b, const double* __restrict__ mask, double* __restrict__ minv, double* __restrict__ x) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nDof) return;
    const int64_t A = r / 3;
    const int i = (int)(r - 3 * A);
    const int64_t s0 = adjPtr[A];
    const int deg = (int)(adjPtr[A + 1] - s0);
    // diagonal entry: column block of A itself (adjacency is sorted: binary search)
    int lo = 0, hi = deg - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (adj[s0 + mid] < A) lo = mid + 1; else hi = mid;
    }
    const double d = data[9 * s0 + (int64_t)i * 3 * deg + 3 * lo + i];
    const double m = mask[r];
    minv[r] = (m != 0.0 && d != 0.0) ? 1.0 / d : 0.0;
    x[r] = (1.0 - m) * b[r];
}

// r = m (b - q), z = Minv r, p = z; partials of r.z and r.r      (q = A x0 computed before)
__global__ void __launch_bounds__(256) pcgInitKernel(int64_t nDof, const double* __restrict__ b, const double* __restrict__ q, const double* __restrict__ mask,
                                                     const double* __restrict__ minv, double* __restrict__ r, double* __restrict__ z, double* __restrict__ p,
                                                     double* __restrict__ partial, int64_t stride) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double rz = 0.0, rr = 0.0;
    if (i < nDof) {
        const double ri = mask[i] * (b[i] - q[i]);
        const double zi = minv[i] * ri;
        r[i] = ri; z[i] = zi; p[i] = zi;
        rz = ri * zi; rr = ri * ri;
    }
    __shared__ double s0[256], s1[256];
    s0[threadIdx.x] = rz; s1[threadIdx.x] = rr;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) { s0[threadIdx.x] += s0[threadIdx.x + w]; s1[threadIdx.x] += s1[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[blockIdx.x] = s0[0]; partial[stride + blockIdx.x] = s1[0]; }
}

// alpha = rz / pq; x += alpha p; r -= alpha q; z = Minv r; partials of r.z and r.r
__global__ void __launch_bounds__(256) pcgUpdateKernel(int64_t nDof, const double* __restrict__ scal, const double* __restrict__ p, const double* __restrict__ q,
                                                       const double* __restrict__ minv, double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                                                       double* __restrict__ partial, int64_t stride) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const double pq = scal[1];
    const double alpha = pq != 0.0 ? scal[0] / pq : 0.0;
    double rz = 0.0, rr = 0.0;
    if (i < nDof) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        const double zi = minv[i] * ri;
        r[i] = ri; z[i] = zi;
        rz = ri * zi; rr = ri * ri;
    }
    __shared__ double s0[256], s1[256];
    s0[threadIdx.x] = rz; s1[threadIdx.x] = rr;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) { s0[threadIdx.x] += s0[threadIdx.x + w]; s1[threadIdx.x] += s1[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[blockIdx.x] = s0[0]; partial[stride + blockIdx.x] = s1[0]; }
}

// beta = rzNew / rz; p = z + beta p; rz <- rzNew (done by thread 0 of block 0 AFTER every block has read it: rz is copied by the host-side sequence)
__global__ void __launch_bounds__(256) pcgDirectionKernel(int64_t nDof, const double* __restrict__ scal, const double* __restrict__ z, double* __restrict__ p) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const double rz = scal[0];
    const double beta = rz != 0.0 ? scal[2] / rz : 0.0;
    if (i < nDof) p[i] = z[i] + beta * p[i];
}

__global__ void pcgShiftKernel(double* scal, int first) {  // after the direction update: rz <- rzNew; keep rr0 of the first iteration
    scal[0] = scal[2];
    if (first) scal[4] = scal[3];
}

__global__ void maskKernel(double* __restrict__ mask, int64_t nDof, const int32_t* __restrict__ dofs, int64_t n) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) mask[dofs[k]] = 0.0;
}

__global__ void fillKernel(double* __restrict__ v, int64_t n, double val) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) v[k] = val;
}

// ---- face pressure (4-node faces of 8-node hexahedra, Abaqus face numbering 1..6) ---------------------------------
// local nodes of face f ordered so that (x1 - x0) x (x3 - x0) is the OUTWARD normal
__constant__ int kHexFaceNodes[6][4] = {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}};

// one thread per face: nodal forces f_a = -p int N_a n dA (2x2 Gauss on the bilinear face), scratch[face][4][3]
__global__ void facePressureKernel(int64_t nFaces, const int32_t* __restrict__ elem, const int32_t* __restrict__ face, const int32_t* __restrict__ conn,
                                   const double* __restrict__ coords, double pressure, double* __restrict__ scratch) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nFaces) return;
    const int fid = face[f] - 1;
    double X[4][3];
    for (int a = 0; a < 4; ++a) {
        const int64_t n = conn[(int64_t)elem[f] * 8 + kHexFaceNodes[fid][a]];
        for (int c = 0; c < 3; ++c) X[a][c] = coords[3 * n + c];
    }
    double out[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const double g = 0.57735026918962576451;
    const double rs[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
    for (int gp = 0; gp < 4; ++gp) {
        const double r = g * rs[gp][0], s = g * rs[gp][1];
        double N[4], dr[4], ds[4];
        for (int a = 0; a < 4; ++a) {
            N[a] = 0.25 * (1 + rs[a][0] * r) * (1 + rs[a][1] * s);
            dr[a] = 0.25 * rs[a][0] * (1 + rs[a][1] * s);
            ds[a] = 0.25 * rs[a][1] * (1 + rs[a][0] * r);
        }
        double tr[3] = {0, 0, 0}, ts[3] = {0, 0, 0};
        for (int a = 0; a < 4; ++a)
            for (int c = 0; c < 3; ++c) { tr[c] += dr[a] * X[a][c]; ts[c] += ds[a] * X[a][c]; }
        const double nA[3] = {tr[1] * ts[2] - tr[2] * ts[1], tr[2] * ts[0] - tr[0] * ts[2], tr[0] * ts[1] - tr[1] * ts[0]};  // outward normal x dA
        for (int a = 0; a < 4; ++a)
            for (int c = 0; c < 3; ++c) out[a][c] -= pressure * N[a] * nA[c];
    }
    for (int a = 0; a < 4; ++a)
        for (int c = 0; c < 3; ++c) scratch[(f * 4 + a) * 3 + c] = out[a][c];
}

// one thread per (loaded node, component): sum its face contributions in ascending (face, local node) order
__global__ void faceGatherKernel(int64_t nLoaded, const int32_t* __restrict__ nodes, const int64_t* __restrict__ incPtr, const int32_t* __restrict__ inc,
                                 const double* __restrict__ scratch, double* __restrict__ pext) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * nLoaded) return;
    const int64_t k = t / 3;
    const int c = (int)(t - 3 * k);
    double s = 0.0;
    for (int64_t q = incPtr[k]; q < incPtr[k + 1]; ++q) s += scratch[(int64_t)inc[q] * 3 + c];
    pext[3 * (int64_t)nodes[k] + c] += s;
}

}  // namespace ewb
