// ek_slab_poisson.cu -- the distributed fast_Poisson() of the x-slab path
// (replaces poisson.cu:75-103 on a domain split along x).
//
// The linear system is the one of ek_poisson.cu path 0: Fourier in the periodic
// x and y, second differences in z between the Dirichlet walls.  x is split over
// the ranks, so the x-transform sits between two slab transposes:
//
//   c+ - c-  [y][z][x_local]   (the LBM kernel writes it in this layout in slab mode: EkConst::dq_sy/dq_sz)
//     cuFFT D2Z along y  -> S  [ky][z][x_local]           (ONE strided batch per chunk, no packing pass:
//                                                          rows of S are already grouped by the
//                                                          destination rank of the transpose)
//     transpose 1: all-to-all over the ky blocks  (host: NCCL; or direct peer writes)
//     k_copy_rows        -> X  [ky_local][z][x_global]    (hand-written)
//     cuFFT Z2Z along x, in place
//     k_zsolve           tridiagonal z-system of every (ky,kx) column, in place
//     cuFFT Z2Z inverse along x, in place
//     k_copy_rows        -> S  [rank][ky_local][z][x_local]
//     transpose 2: all-to-all back                -> R = [ky][z][x_local]
//     cuFFT Z2D along y  -> A  [y][z][x_local]
//     k_rows_to_planes   -> phi [z][y][x_local]           (hand-written)
//
// Everything is chunked along z (the chunks are groups of the LBM kernel's
// z-blocks), so that the host can start the forward half of chunk c as soon as
// the LBM launch that produces its planes has finished, and so that the
// all-to-all of one chunk travels while the next one is transformed.
// cuFFT does the transforms only; packing, the eigen-solve and unpacking are
// the kernels of this file and k_zsolve (ek_poisson.cu).
#include <stdlib.h>
#include <string.h>

#include "ek_handle.h"

namespace {

// phi[z*plane + y*PX + x] = A[(y*NZ + z)*NXl + x]   for the planes z = za+1 .. of a chunk
__global__ void __launch_bounds__(256) k_rows_to_planes(int NXl, int NZ, int PX, long long plane, int za,
                                                        const double *__restrict__ A, double *__restrict__ phi)
{
    // one block per row of NXl doubles as pairs of columns (NXl and PX are even); four loads in flight per thread
    const int y = blockIdx.x, z = za + 1 + blockIdx.y, n2 = NXl / 2;
    const double2 *__restrict__ s = reinterpret_cast<const double2 *>(A + ((size_t)y * NZ + z) * NXl);
    double2 *__restrict__ d = reinterpret_cast<double2 *>(phi + (size_t)z * plane + (size_t)y * PX);
    int x = threadIdx.x;
    for (; x + 3 * 256 < n2; x += 4 * 256) {
        const double2 a = s[x], b = s[x + 256], c = s[x + 512], e = s[x + 768];
        d[x] = a; d[x + 256] = b; d[x + 512] = c; d[x + 768] = e;
    }
    for (; x < n2; x += 256) d[x] = s[x];
}

// The two re-blockings around the transposes are copies of rows of NXl complex
// numbers between a [rank][ky_local][z][x_local] buffer and the full-x pencils
// X[ky_local][z][x_global]: row (i, ky, zi) <-> X[ky][z0+zi][i*NXl ...].  src/dst
// are per-rank base pointers, so the same kernel serves the NCCL transport
// (local buffers) and direct peer-memory writes.
struct RowPtrs {
    const double2 *src[EK_MAX_RANKS];
    double2 *dst[EK_MAX_RANKS];
};

__global__ void __launch_bounds__(256) k_copy_rows(RowPtrs p, int rowlen, int nz, long long src_ky, long long src_z,
                                                   long long dst_ky, long long dst_z)
{
    // one block per row; four independent 16-byte loads in flight per thread (one per thread ran at 2.6 TB/s)
    const int ky = blockIdx.x / nz, zi = blockIdx.x % nz, i = blockIdx.y;
    const double2 *__restrict__ s = p.src[i] + (size_t)ky * src_ky + (size_t)zi * src_z;
    double2 *__restrict__ d = p.dst[i] + (size_t)ky * dst_ky + (size_t)zi * dst_z;
    int x = threadIdx.x;
    for (; x + 3 * 256 < rowlen; x += 4 * 256) {
        const double2 a = s[x], b = s[x + 256], c = s[x + 512], e = s[x + 768];
        d[x] = a; d[x + 256] = b; d[x + 512] = c; d[x + 768] = e;
    }
    for (; x < rowlen; x += 256) d[x] = s[x];
}

// Re-blocking for the way back WITH the ghost columns of phi: row (i, ky, zi) of rank i's block gets its NXl
// columns of the pencil plus column (i+1)*NXl and column i*NXl - 1 (periodic in x) -- what rank i needs as the
// right / left ghost column of phi, so that no separate phi halo exchange is needed after the inverse y-transform.
__global__ void __launch_bounds__(256) k_scatter_rows_ghost(const double2 *__restrict__ X, double2 *__restrict__ Sd, int NXl,
                                                            int NXg, int nz, int kyl, long long x_ky, long long x_z)
{
    const int ky = blockIdx.x / nz, zi = blockIdx.x % nz, i = blockIdx.y;
    const double2 *__restrict__ s = X + (size_t)ky * x_ky + (size_t)zi * x_z;
    const int W = NXl + 2;
    double2 *__restrict__ d = Sd + ((size_t)i * kyl * nz + (size_t)ky * nz + zi) * W;
    const int x0 = i * NXl;
    for (int x = threadIdx.x; x < W; x += 256) {
        int col = x0 + x;
        if (x == NXl) col = (x0 + NXl) % NXg;
        else if (x == NXl + 1) col = (x0 - 1 + NXg) % NXg;
        d[x] = s[col];
    }
}

ek_status plan_for(ek_handle *h, std::map<int, cufftHandle> &plans, int nzc, bool forward)
{
    if (plans.count(nzc)) return EK_OK;
    EkSlabPoisson &S = h->sp;
    cufftHandle plan;
    int n[1] = {S.NY};
    int emb[1] = {S.NY};
    const int NZ = S.M + 2;
    if (forward) {
        // in: dq + (za+1)*NXl, element (y; b) at y*(NZ*NXl) + b;  out: chunk buffer, (ky; b) at ky*(nzc*NXl) + b
        EK_CUFFT(h, cufftPlanMany(&plan, 1, n, emb, NZ * S.NXl, 1, emb, nzc * S.NXl, 1, CUFFT_D2Z, nzc * S.NXl));
    } else {
        EK_CUFFT(h, cufftPlanMany(&plan, 1, n, emb, nzc * S.NXl, 1, emb, NZ * S.NXl, 1, CUFFT_Z2D, nzc * S.NXl));
    }
    plans[nzc] = plan;
    return EK_OK;
}

}  // namespace

void ek_slab_poisson_destroy(ek_handle *h)
{
    EkSlabPoisson &S = h->sp;
    for (auto &kv : S.plan_yf) cufftDestroy(kv.second);
    for (auto &kv : S.plan_yb) cufftDestroy(kv.second);
    for (auto &kv : S.plan_ybg) cufftDestroy(kv.second);
    S.plan_yf.clear();
    S.plan_yb.clear();
    S.plan_ybg.clear();
    cudaFree(S.Sg); cudaFree(S.Rg);
    S.Sg = S.Rg = nullptr;
    S.ghosts = false;
    if (S.plan_x_ok) cufftDestroy(S.plan_x);
    S.plan_x_ok = false;
    for (int i = 0; i < EK_MAX_RANKS; ++i) {
        if (S.peer_ipc[i]) { cudaIpcCloseMemHandle(S.peerX[i]); cudaIpcCloseMemHandle(S.peerR[i]); }
        S.peer_ipc[i] = false;
        S.peerX[i] = S.peerR[i] = nullptr;
    }
    cudaFree(S.A); cudaFree(S.S); cudaFree(S.R); cudaFree(S.X); cudaFree(S.cp);
    S.A = nullptr; S.S = S.R = S.X = nullptr; S.cp = nullptr;
    S.ready = false;
}

extern "C" {

// Chunks of the pipelined stage in LBM z-blocks: fills bounds[0..K] (bounds[0] = 0, bounds[K] = nblocks, strictly
// increasing) and returns K <= EK_MAX_CHUNKS.  nchunks >= 1: that many equal chunks (at most one per z-block).
// nchunks <= 0: automatic -- seven chunks of sizes 1:2:3:4:3:2:1 from 16 z-blocks on (measured best at 2 x 134 M
// cells, 46.0 ms per step against 46.4 for four equal chunks: the LAST chunk's forward half and the FIRST chunks'
// way back are the parts of the stage that nothing overlaps, so they are the small ones), up to four equal chunks
// below.  sizes_csv (the EK_POISSON_CHUNK_BLOCKS environment variable, e.g. "2,3,4,4,2,1") overrides both when its
// entries are positive and add up to nblocks.  Pure host arithmetic (CPU-tested).
int ek_slab_poisson_plan_chunks(int nblocks, int nchunks, const char *sizes_csv, int *bounds)
{
    if (nblocks < 1 || !bounds) return 0;
    int K = nchunks < 1 ? 1 : (nchunks > nblocks ? nblocks : nchunks);
    if (K > EK_MAX_CHUNKS) K = EK_MAX_CHUNKS;
    if (nchunks <= 0 && nblocks >= 16) {
        static const int w[7] = {1, 2, 3, 4, 3, 2, 1};
        K = 7;
        bounds[0] = 0;
        int acc = 0;
        for (int k = 0; k < 7; ++k) {
            acc += w[k];
            bounds[k + 1] = (int)((long long)nblocks * acc / 16);
            if (bounds[k + 1] <= bounds[k]) bounds[k + 1] = bounds[k] + 1;
        }
        bounds[7] = nblocks;
    } else {
        if (nchunks <= 0) K = nblocks < 4 ? nblocks : 4;
        for (int k = 0; k <= K; ++k) bounds[k] = (int)((long long)nblocks * k / K);
    }
    if (sizes_csv) {
        int sizes[EK_MAX_CHUNKS], n = 0, sum = 0;
        bool ok = true;
        for (const char *q = sizes_csv; *q;) {
            if (n == EK_MAX_CHUNKS) { ok = false; break; }
            sizes[n] = atoi(q);
            if (sizes[n] < 1) { ok = false; break; }
            sum += sizes[n++];
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
        if (ok && n >= 1 && sum == nblocks) {
            K = n;
            bounds[0] = 0;
            for (int k = 0; k < n; ++k) bounds[k + 1] = bounds[k] + sizes[k];
        }
    }
    return K;
}

// nchunks groups of the LBM z-blocks ("zchunk" planes each) define the chunks.
ek_status ek_slab_poisson_setup(ek_handle *h, int nchunks)
{
    if (!h || !h->slab) return EK_ERR_STATE;
    if (h->nranks > EK_MAX_RANKS) { ek_set_error(h, "too many ranks"); return EK_ERR_INVALID; }
    DeviceGuard g(h->device);
    ek_status st = ek_alloc_state(h);
    if (st != EK_OK) return st;
    ek_slab_poisson_destroy(h);
    EkSlabPoisson &S = h->sp;
    const EkConst &c = h->c;
    if (c.NX % 2) { ek_set_error(h, "the slab width must be even"); return EK_ERR_INVALID; }
    S.P = h->nranks; S.r = h->rank;
    S.NXl = c.NX; S.NXg = h->NXg; S.NY = c.NY; S.NYH = c.NY / 2 + 1; S.M = c.NZ - 2;
    S.kyl = (S.NYH + S.P - 1) / S.P;
    const int nblocks = (c.NZ + h->zchunk - 1) / h->zchunk;
    int bounds[EK_MAX_CHUNKS + 1];
    S.K = ek_slab_poisson_plan_chunks(nblocks, nchunks, getenv("EK_POISSON_CHUNK_BLOCKS"), bounds);
    for (int k = 0; k <= S.K; ++k) {
        const int b = bounds[k];     // first LBM z-block of chunk k
        int z = b * h->zchunk;                                   // first plane
        if (z > c.NZ) z = c.NZ;
        S.block0[k] = b;
        // interior-plane index zi = z - 1, clipped to [0, M]
        int zi = z - 1;
        if (zi < 0) zi = 0;
        if (zi > S.M) zi = S.M;
        if (k == S.K) zi = S.M;
        S.z0[k] = zi;
    }
    const size_t nreal = (size_t)S.NY * c.NZ * (S.NXl + 2);   // rows of NXl (+ 2 ghost columns, ek_slab_poisson_enable_ghosts)
    const size_t nspec = (size_t)S.P * S.kyl * S.M * S.NXl;
    const size_t npen = (size_t)S.kyl * S.M * S.NXg;
    EK_CUDA(h, cudaMalloc((void **)&S.A, nreal * sizeof(double)));
    EK_CUDA(h, cudaMalloc((void **)&S.S, nspec * sizeof(cufftDoubleComplex)));
    EK_CUDA(h, cudaMalloc((void **)&S.R, nspec * sizeof(cufftDoubleComplex)));
    EK_CUDA(h, cudaMalloc((void **)&S.X, npen * sizeof(cufftDoubleComplex)));
    // rows ky >= NY/2+1 of the last rank's block are padding: keep them zero
    EK_CUDA(h, cudaMemsetAsync(S.S, 0, nspec * sizeof(cufftDoubleComplex), h->stream));
    EK_CUDA(h, cudaMemsetAsync(S.R, 0, nspec * sizeof(cufftDoubleComplex), h->stream));
    EK_CUDA(h, cudaMemsetAsync(S.X, 0, npen * sizeof(cufftDoubleComplex), h->stream));
    for (int k = 0; k < S.K; ++k) {
        const int nzc = S.z0[k + 1] - S.z0[k];
        if (nzc <= 0) continue;
        st = plan_for(h, S.plan_yf, nzc, true);
        if (st != EK_OK) return st;
        st = plan_for(h, S.plan_yb, nzc, false);
        if (st != EK_OK) return st;
    }
    {
        int n[1] = {S.NXg};
        EK_CUFFT(h, cufftPlanMany(&S.plan_x, 1, n, nullptr, 1, S.NXg, nullptr, 1, S.NXg, CUFFT_Z2Z, S.kyl * S.M));
        S.plan_x_ok = true;
    }
    // LU factors of the z-operator of my (ky, kx) columns
    const int ncols = S.kyl * S.NXg;
    EK_CUDA(h, cudaMalloc((void **)&S.cp, (size_t)S.M * ncols * sizeof(double)));
    ek_launch_zfactor_cols(ncols, S.NXg, c.NY, S.r * S.kyl, S.M, h->p.Lx, h->p.Ly, c.dz, S.cp, h->stream);
    EK_CUDA(h, cudaGetLastError());
    for (int i = 0; i < EK_MAX_RANKS; ++i) { S.peerX[i] = nullptr; S.peerR[i] = nullptr; }
    S.peerX[S.r] = S.X;
    S.peerR[S.r] = S.R;
    // from now on c+ - c- is kept as rows for the y-transform: [y][z][x] without ghost columns
    // (the allocation of NZ*NY*PX doubles is large enough); whoever wrote it before must write it again
    h->c.dq_sy = (long long)c.NZ * S.NXl;
    h->c.dq_sz = S.NXl;
    S.ready = true;
    return EK_OK;
}

int ek_slab_poisson_chunks(ek_handle *h) { return (h && h->sp.ready) ? h->sp.K : 0; }

// chunk k: the LBM z-blocks [block0, block1) produce its planes; `send`/`recv`
// are its transpose buffers, `count` complex numbers each (nranks equal parts)
ek_status ek_slab_poisson_chunk(ek_handle *h, int k, int *block0, int *block1, void **send, void **recv,
                                long long *count)
{
    if (!h || !h->sp.ready || k < 0 || k >= h->sp.K) return EK_ERR_INVALID;
    EkSlabPoisson &S = h->sp;
    const size_t off = (size_t)S.P * S.kyl * S.z0[k] * S.NXl;
    if (block0) *block0 = S.block0[k];
    if (block1) *block1 = S.block0[k + 1];
    if (send) *send = S.S + off;
    if (recv) *recv = S.R + off;
    if (count) *count = (long long)S.P * S.kyl * (S.z0[k + 1] - S.z0[k]) * S.NXl;
    return EK_OK;
}

// c+ - c- of the planes of chunk k -> y-spectrum in the chunk's send buffer
ek_status ek_slab_poisson_forward(ek_handle *h, int k)
{
    if (!h || !h->sp.ready || k < 0 || k >= h->sp.K) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    EkSlabPoisson &S = h->sp;
    const EkConst &c = h->c;
    const int za = S.z0[k], nzc = S.z0[k + 1] - za;
    if (nzc <= 0) return EK_OK;
    cufftHandle plan = S.plan_yf[nzc];
    cufftDoubleComplex *out = S.S + (size_t)S.P * S.kyl * za * S.NXl;
    // rows ky >= NY/2+1 (padding of the last rank's block) are not written by the transform
    if (S.P * S.kyl > S.NYH)
        EK_CUDA(h, cudaMemsetAsync(out + (size_t)S.NYH * nzc * S.NXl, 0,
                                   (size_t)(S.P * S.kyl - S.NYH) * nzc * S.NXl * sizeof(cufftDoubleComplex), h->stream));
    EK_CUFFT(h, cufftSetStream(plan, h->stream));
    EK_CUFFT(h, cufftExecD2Z(plan, h->dq + (size_t)(za + 1) * S.NXl, out));
    h->poisson_launches += 1;
    return EK_OK;
}

// after transpose 1 of chunk k: received ky blocks -> full-x pencils
ek_status ek_slab_poisson_gather_x(ek_handle *h, int k)
{
    if (!h || !h->sp.ready || k < 0 || k >= h->sp.K) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    EkSlabPoisson &S = h->sp;
    const int za = S.z0[k], nzc = S.z0[k + 1] - za;
    if (nzc <= 0) return EK_OK;
    RowPtrs p;
    const double2 *R = reinterpret_cast<const double2 *>(S.R) + (size_t)S.P * S.kyl * za * S.NXl;
    double2 *X = reinterpret_cast<double2 *>(S.X);
    for (int i = 0; i < S.P; ++i) {
        p.src[i] = R + (size_t)i * S.kyl * nzc * S.NXl;                // what rank i sent: [kyl][nzc][NXl]
        p.dst[i] = X + (size_t)za * S.NXg + (size_t)i * S.NXl;         // its columns of my pencils
    }
    dim3 b(256), gr(S.kyl * nzc, S.P);
    k_copy_rows<<<gr, b, 0, h->stream>>>(p, S.NXl, nzc, (long long)nzc * S.NXl, S.NXl, (long long)S.M * S.NXg, S.NXg);
    EK_CUDA(h, cudaGetLastError());
    h->poisson_launches += 1;
    return EK_OK;
}

// x-transform, z-solve, inverse x-transform of my ky rows (all chunks present)
ek_status ek_slab_poisson_solve(ek_handle *h)
{
    if (!h || !h->sp.ready) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    EkSlabPoisson &S = h->sp;
    const EkConst &c = h->c;
    EK_CUFFT(h, cufftSetStream(S.plan_x, h->stream));
    EK_CUFFT(h, cufftExecZ2Z(S.plan_x, S.X, S.X, CUFFT_FORWARD));
    const double nxy = (double)S.NXg * (double)c.NY;
    const double size = (double)((unsigned int)S.NXg * (unsigned int)c.NY * (unsigned int)(2 * (c.NZ - 1)));
    const double off = h->dc_mode == EK_DC_PRESCRIBED ? -h->dc_ghat0 / size : 0.0;
    ek_launch_zsolve_rows(S.kyl, S.NXg, S.M, reinterpret_cast<double *>(S.X), S.cp, -(c.CtoC / c.eps) * c.dz * c.dz,
                          -c.voltage * nxy, -c.voltage2 * nxy, 1.0 / nxy, off, S.r == 0, h->stream);
    EK_CUDA(h, cudaGetLastError());
    EK_CUFFT(h, cufftExecZ2Z(S.plan_x, S.X, S.X, CUFFT_INVERSE));
    h->poisson_launches += 1;
    return EK_OK;
}

// before transpose 2 of chunk k: pencils -> per-rank x blocks in the send buffer
ek_status ek_slab_poisson_scatter_x(ek_handle *h, int k)
{
    if (!h || !h->sp.ready || k < 0 || k >= h->sp.K) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    EkSlabPoisson &S = h->sp;
    const int za = S.z0[k], nzc = S.z0[k + 1] - za;
    if (nzc <= 0) return EK_OK;
    RowPtrs p;
    double2 *Sd = reinterpret_cast<double2 *>(S.S) + (size_t)S.P * S.kyl * za * S.NXl;
    const double2 *X = reinterpret_cast<const double2 *>(S.X);
    for (int i = 0; i < S.P; ++i) {
        p.src[i] = X + (size_t)za * S.NXg + (size_t)i * S.NXl;
        p.dst[i] = Sd + (size_t)i * S.kyl * nzc * S.NXl;
    }
    dim3 b(256), gr(S.kyl * nzc, S.P);
    k_copy_rows<<<gr, b, 0, h->stream>>>(p, S.NXl, nzc, (long long)S.M * S.NXg, S.NXg, (long long)nzc * S.NXl, S.NXl);
    EK_CUDA(h, cudaGetLastError());
    h->poisson_launches += 1;
    return EK_OK;
}

// ---- direct peer-memory transport of the two transposes ------------------------------
// Every rank maps the pencil buffer X and the receive buffer R of every other rank (CUDA
// IPC between the processes of one node; plain pointers when the slabs share a process).
// The re-blocking kernel then writes its rows straight into the peers' buffers over
// NVLink: push_x(k) = all-to-all 1 + gather_x(k), push_back(k) = scatter_x(k) + all-to-all 2,
// without the intermediate send/receive copies.  The host provides the cross-rank
// barriers: one after the push_x of all chunks, one after the push_back of all chunks.
ek_status ek_slab_poisson_ipc_export(ek_handle *h, void *handles)
{
    if (!h || !h->sp.ready || !handles) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    cudaIpcMemHandle_t hx, hr;
    EK_CUDA(h, cudaIpcGetMemHandle(&hx, h->sp.X));
    EK_CUDA(h, cudaIpcGetMemHandle(&hr, h->sp.R));
    memcpy(handles, &hx, sizeof(hx));
    memcpy((char *)handles + sizeof(hx), &hr, sizeof(hr));
    return EK_OK;
}

int ek_slab_poisson_ipc_bytes(void) { return (int)(2 * sizeof(cudaIpcMemHandle_t)); }

ek_status ek_slab_poisson_ipc_import(ek_handle *h, int rank, const void *handles)
{
    if (!h || !h->sp.ready || !handles || rank < 0 || rank >= h->sp.P) return EK_ERR_INVALID;
    if (rank == h->sp.r) return EK_OK;
    DeviceGuard g(h->device);
    cudaIpcMemHandle_t hx, hr;
    memcpy(&hx, handles, sizeof(hx));
    memcpy(&hr, (const char *)handles + sizeof(hx), sizeof(hr));
    void *px = nullptr, *pr = nullptr;
    EK_CUDA(h, cudaIpcOpenMemHandle(&px, hx, cudaIpcMemLazyEnablePeerAccess));
    EK_CUDA(h, cudaIpcOpenMemHandle(&pr, hr, cudaIpcMemLazyEnablePeerAccess));
    h->sp.peerX[rank] = (cufftDoubleComplex *)px;
    h->sp.peerR[rank] = (cufftDoubleComplex *)pr;
    h->sp.peer_ipc[rank] = true;
    return EK_OK;
}

// peers living in this process (single-GPU tests): their buffers as plain pointers
ek_status ek_slab_poisson_set_peer(ek_handle *h, int rank, void *X, void *R)
{
    if (!h || !h->sp.ready || rank < 0 || rank >= h->sp.P) return EK_ERR_INVALID;
    h->sp.peerX[rank] = (cufftDoubleComplex *)X;
    h->sp.peerR[rank] = (cufftDoubleComplex *)R;
    return EK_OK;
}

ek_status ek_slab_poisson_my_buffers(ek_handle *h, void **X, void **R)
{
    if (!h || !h->sp.ready || !X || !R) return EK_ERR_INVALID;
    *X = h->sp.X;
    *R = h->sp.R;
    return EK_OK;
}

// rows (ky, zi) of NXl complex numbers, [kyl][nzc] of them, between two pitched layouts: one strided
// 3-D copy on the copy engines (NVLink DMA for a peer destination) instead of SM loads/stores
static cudaError_t copy_rows_dma(const void *src, size_t src_row_bytes, size_t src_rows_per_ky, void *dst,
                                 size_t dst_row_bytes, size_t dst_rows_per_ky, size_t width_bytes, int nzc, int kyl,
                                 cudaStream_t st)
{
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    p.srcPtr = make_cudaPitchedPtr(const_cast<void *>(src), src_row_bytes, width_bytes, src_rows_per_ky);
    p.dstPtr = make_cudaPitchedPtr(dst, dst_row_bytes, width_bytes, dst_rows_per_ky);
    p.extent = make_cudaExtent(width_bytes, (size_t)nzc, (size_t)kyl);
    p.kind = cudaMemcpyDefault;
    return cudaMemcpy3DAsync(&p, st);
}

static bool peers_ready(const EkSlabPoisson &S)
{
    for (int i = 0; i < S.P; ++i)
        if (!S.peerX[i] || !S.peerR[i]) return false;
    return true;
}

// y-spectrum of chunk k (my send buffer) -> every rank's pencils
ek_status ek_slab_poisson_push_x(ek_handle *h, int k)
{
    if (!h || !h->sp.ready || k < 0 || k >= h->sp.K) return EK_ERR_INVALID;
    EkSlabPoisson &S = h->sp;
    if (!peers_ready(S)) { ek_set_error(h, "peer buffers not mapped"); return EK_ERR_STATE; }
    DeviceGuard g(h->device);
    const int za = S.z0[k], nzc = S.z0[k + 1] - za;
    if (nzc <= 0) return EK_OK;
    RowPtrs p;
    const double2 *Ss = reinterpret_cast<const double2 *>(S.S) + (size_t)S.P * S.kyl * za * S.NXl;
    for (int i = 0; i < S.P; ++i) {
        p.src[i] = Ss + (size_t)i * S.kyl * nzc * S.NXl;                                      // rank i's ky rows
        p.dst[i] = reinterpret_cast<double2 *>(S.peerX[i]) + (size_t)za * S.NXg + (size_t)S.r * S.NXl;  // my columns there
    }
    if (S.dma) {
        // start with my right-hand neighbour so that the ranks do not all write to the same peer at once
        for (int j = 0; j < S.P; ++j) {
            const int i = (S.r + 1 + j) % S.P;
            EK_CUDA(h, copy_rows_dma(p.src[i], (size_t)S.NXl * 16, nzc, p.dst[i], (size_t)S.NXg * 16, S.M,
                                     (size_t)S.NXl * 16, nzc, S.kyl, h->stream));
        }
        return EK_OK;
    }
    dim3 b(256), gr(S.kyl * nzc, S.P);
    k_copy_rows<<<gr, b, 0, h->stream>>>(p, S.NXl, nzc, (long long)nzc * S.NXl, S.NXl, (long long)S.M * S.NXg, S.NXg);
    EK_CUDA(h, cudaGetLastError());
    h->poisson_launches += 1;
    return EK_OK;
}

// my pencils of chunk k -> every rank's receive buffer (its x block of my ky rows)
ek_status ek_slab_poisson_push_back(ek_handle *h, int k)
{
    if (!h || !h->sp.ready || k < 0 || k >= h->sp.K) return EK_ERR_INVALID;
    EkSlabPoisson &S = h->sp;
    if (!peers_ready(S)) { ek_set_error(h, "peer buffers not mapped"); return EK_ERR_STATE; }
    DeviceGuard g(h->device);
    const int za = S.z0[k], nzc = S.z0[k + 1] - za;
    if (nzc <= 0) return EK_OK;
    RowPtrs p;
    const double2 *X = reinterpret_cast<const double2 *>(S.X);
    for (int i = 0; i < S.P; ++i) {
        p.src[i] = X + (size_t)za * S.NXg + (size_t)i * S.NXl;
        p.dst[i] = reinterpret_cast<double2 *>(S.peerR[i]) + (size_t)S.P * S.kyl * za * S.NXl
                   + (size_t)S.r * S.kyl * nzc * S.NXl;
    }
    if (S.dma) {
        for (int j = 0; j < S.P; ++j) {
            const int i = (S.r + 1 + j) % S.P;
            EK_CUDA(h, copy_rows_dma(p.src[i], (size_t)S.NXg * 16, S.M, p.dst[i], (size_t)S.NXl * 16, nzc,
                                     (size_t)S.NXl * 16, nzc, S.kyl, h->stream));
        }
        return EK_OK;
    }
    dim3 b(256), gr(S.kyl * nzc, S.P);
    k_copy_rows<<<gr, b, 0, h->stream>>>(p, S.NXl, nzc, (long long)S.M * S.NXg, S.NXg, (long long)nzc * S.NXl, S.NXl);
    EK_CUDA(h, cudaGetLastError());
    h->poisson_launches += 1;
    return EK_OK;
}

// 1: the peer-memory pushes use the copy engines (strided 3-D copies), 0: the re-blocking kernel
ek_status ek_slab_poisson_set_dma(ek_handle *h, int on)
{
    if (!h) return EK_ERR_INVALID;
    h->sp.dma = on != 0;
    return EK_OK;
}

// after transpose 2 of chunk k: inverse y-transform -> interior planes of phi
ek_status ek_slab_poisson_backward(ek_handle *h, int k)
{
    if (!h || !h->sp.ready || k < 0 || k >= h->sp.K) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    EkSlabPoisson &S = h->sp;
    const EkConst &c = h->c;
    const int za = S.z0[k], nzc = S.z0[k + 1] - za;
    if (nzc <= 0) return EK_OK;
    cufftHandle plan = S.plan_yb[nzc];
    EK_CUFFT(h, cufftSetStream(plan, h->stream));
    EK_CUFFT(h, cufftExecZ2D(plan, S.R + (size_t)S.P * S.kyl * za * S.NXl, S.A + (size_t)(za + 1) * S.NXl));
    dim3 b(256), gr(S.NY, nzc);
    k_rows_to_planes<<<gr, b, 0, h->stream>>>(S.NXl, c.NZ, c.PX, c.plane, za, S.A, h->fld[EK_PHI]);
    EK_CUDA(h, cudaGetLastError());
    h->poisson_launches += 1;
    return EK_OK;
}

// ---- way back with the ghost columns of phi inside transpose 2 --------------------------------
// After enable_ghosts() the host runs, per chunk, scatter_xg(k); all-to-all(chunk_back buffers);
// backward_g(k) instead of scatter_x / all-to-all / backward / phi halo exchange: the rows that travel
// are two columns wider (0.2 % more data at 1024 columns per slab) and the separate NCCL exchange of the phi
// ghost columns per chunk -- pure latency -- disappears.
ek_status ek_slab_poisson_enable_ghosts(ek_handle *h)
{
    if (!h || !h->sp.ready) return EK_ERR_STATE;
    EkSlabPoisson &S = h->sp;
    if (S.ghosts) return EK_OK;
    DeviceGuard g(h->device);
    const size_t nspec = (size_t)S.P * S.kyl * S.M * (S.NXl + 2);
    EK_CUDA(h, cudaMalloc((void **)&S.Sg, nspec * sizeof(cufftDoubleComplex)));
    EK_CUDA(h, cudaMalloc((void **)&S.Rg, nspec * sizeof(cufftDoubleComplex)));
    EK_CUDA(h, cudaMemsetAsync(S.Sg, 0, nspec * sizeof(cufftDoubleComplex), h->stream));
    EK_CUDA(h, cudaMemsetAsync(S.Rg, 0, nspec * sizeof(cufftDoubleComplex), h->stream));
    const int W = S.NXl + 2, NZ = S.M + 2;
    for (int k = 0; k < S.K; ++k) {
        const int nzc = S.z0[k + 1] - S.z0[k];
        if (nzc <= 0 || S.plan_ybg.count(nzc)) continue;
        cufftHandle plan;
        int n[1] = {S.NY}, emb[1] = {S.NY};
        EK_CUFFT(h, cufftPlanMany(&plan, 1, n, emb, nzc * W, 1, emb, NZ * W, 1, CUFFT_Z2D, nzc * W));
        S.plan_ybg[nzc] = plan;
    }
    S.ghosts = true;
    return EK_OK;
}

ek_status ek_slab_poisson_chunk_back(ek_handle *h, int k, void **send, void **recv, long long *count)
{
    if (!h || !h->sp.ready || !h->sp.ghosts || k < 0 || k >= h->sp.K) return EK_ERR_INVALID;
    EkSlabPoisson &S = h->sp;
    const size_t off = (size_t)S.P * S.kyl * S.z0[k] * (S.NXl + 2);
    if (send) *send = S.Sg + off;
    if (recv) *recv = S.Rg + off;
    if (count) *count = (long long)S.P * S.kyl * (S.z0[k + 1] - S.z0[k]) * (S.NXl + 2);
    return EK_OK;
}

ek_status ek_slab_poisson_scatter_xg(ek_handle *h, int k)
{
    if (!h || !h->sp.ready || !h->sp.ghosts || k < 0 || k >= h->sp.K) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    EkSlabPoisson &S = h->sp;
    const int za = S.z0[k], nzc = S.z0[k + 1] - za;
    if (nzc <= 0) return EK_OK;
    double2 *Sd = reinterpret_cast<double2 *>(S.Sg) + (size_t)S.P * S.kyl * za * (S.NXl + 2);
    const double2 *X = reinterpret_cast<const double2 *>(S.X) + (size_t)za * S.NXg;
    dim3 b(256), gr(S.kyl * nzc, S.P);
    k_scatter_rows_ghost<<<gr, b, 0, h->stream>>>(X, Sd, S.NXl, S.NXg, nzc, S.kyl, (long long)S.M * S.NXg, S.NXg);
    EK_CUDA(h, cudaGetLastError());
    h->poisson_launches += 1;
    return EK_OK;
}

// after transpose 2 of chunk k: inverse y-transform -> interior planes of phi, ghost columns included
ek_status ek_slab_poisson_backward_g(ek_handle *h, int k)
{
    if (!h || !h->sp.ready || !h->sp.ghosts || k < 0 || k >= h->sp.K) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    EkSlabPoisson &S = h->sp;
    const EkConst &c = h->c;
    const int za = S.z0[k], nzc = S.z0[k + 1] - za, W = S.NXl + 2;
    if (nzc <= 0) return EK_OK;
    if (c.PX != W) { ek_set_error(h, "row pitch of phi is not NX + 2"); return EK_ERR_STATE; }
    cufftHandle plan = S.plan_ybg[nzc];
    EK_CUFFT(h, cufftSetStream(plan, h->stream));
    EK_CUFFT(h, cufftExecZ2D(plan, S.Rg + (size_t)S.P * S.kyl * za * W, S.A + (size_t)(za + 1) * W));
    dim3 b(256), gr(S.NY, nzc);
    k_rows_to_planes<<<gr, b, 0, h->stream>>>(W, c.NZ, c.PX, c.plane, za, S.A, h->fld[EK_PHI]);
    EK_CUDA(h, cudaGetLastError());
    h->poisson_launches += 1;
    return EK_OK;
}

}  // extern "C"
