// ek_slab.cu -- x-slab decomposition: the device-side pieces of the multi-GPU
// path.  One handle = one slab = one process/GPU; the host
// (ek-pnp-3d_b200/slab.py, torch.distributed over NCCL) moves the buffers.
//
//  * populations: two ghost columns per slab (x = NX: right neighbour's first
//    column, x = NX+1: left neighbour's last column).  With the A-A pattern only
//    the 9 populations that cross a face travel, and only once per step:
//      phase A (after an even step): boundary columns -> neighbours' ghosts,
//               so that the odd step can pull through the face;
//      phase B (after an odd step):  ghost columns (now holding what the odd
//               step pushed through the face) -> neighbours' boundary columns.
//    9*NY*NZ doubles per set, face and direction (18.9 MB for 4 sets at 256^2).
//  * potential: one ghost column of phi per face for the fused E = -grad(phi).
//  * Poisson: the z-solve kernel on an arbitrary block of (ky, kx) columns of
//    the full-x spectrum, between the host's transposes.
#include <string.h>

#include "ek_handle.h"

namespace {

__constant__ int kPlus[9] = {1, 7, 9, 13, 15, 19, 21, 23, 26};    // c_x = +1
__constant__ int kMinus[9] = {2, 8, 10, 14, 16, 20, 22, 24, 25};  // c_x = -1

struct Lat4 { double *p[4]; };

// buf[((s*9 + k)*NZ + z)*NY + y]  <->  slot list[k] of column `col` of set s.  One launch serves both faces of
// the slab (blockIdx.z = face*36 + s*9 + k): half the launches of a halo exchange, which at 8 M cells per GPU is
// latency, not bandwidth
struct HaloFaces {
    int col[2], plus[2];
    double *buf[2];
};

template <bool PACK>
__global__ void k_halo_column(EkConst c, Lat4 lat, HaloFaces f)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= c.NY) return;
    const int z = blockIdx.y;
    const int face = blockIdx.z / 36, sk = blockIdx.z % 36;
    const int s = sk / 9, k = sk % 9;
    const int d = f.plus[face] ? kPlus[k] : kMinus[k];
    double *q = lat.p[s] + (size_t)z * c.lplane + (size_t)y * c.lrow + ek_lat_col(f.col[face]) + (size_t)d * EK_TILE;
    double *b = f.buf[face] + (((size_t)(s * 9 + k) * c.NZ + z) * c.NY + y);
    if (PACK) *b = *q; else *q = *b;
}

struct PhiFaces {
    int col[2];
    double *buf[2];
};

template <bool PACK>
__global__ void k_phi_column(EkConst c, double *phi, PhiFaces f, int z0)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= c.NY) return;
    const int z = z0 + blockIdx.y;
    const int face = blockIdx.z;
    double *q = phi + (size_t)z * c.plane + (size_t)y * c.PX + f.col[face];
    double *b = f.buf[face] + (size_t)z * c.NY + y;
    if (PACK) *b = *q; else *q = *b;
}

Lat4 current_lattice(ek_handle *h)
{
    Lat4 l;
    for (int s = 0; s < 4; ++s) l.p[s] = h->lat[0][s];
    return l;
}

}  // namespace

extern "C" {

ek_status ek_create_slab(const ek_params *global, int device, int rank, int nranks, ek_handle **out)
{
    if (!global || !out || nranks < 1 || rank < 0 || rank >= nranks) return EK_ERR_INVALID;
    if (global->NX % nranks != 0) return EK_ERR_INVALID;
    ek_params local = *global;
    local.NX = global->NX / nranks;      // Lx stays the GLOBAL length (wavenumbers)
    ek_handle *h = nullptr;
    ek_status st = ek_create(&local, device, &h);
    if (st != EK_OK) return st;
    h->slab = true;
    h->rank = rank;
    h->nranks = nranks;
    h->NXg = global->NX;
    h->stream_mode = EK_STREAM_AA;
    ek_compute_consts(local, h->c, true);
    h->zchunk = ek_auto_zchunk(h->c);
    *out = h;
    return EK_OK;
}

// Run on a caller-provided stream (e.g. torch's current stream, so that NCCL
// collectives issued by the host are ordered with the kernels).
ek_status ek_set_stream(ek_handle *h, void *stream)
{
    if (!h) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    EK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)stream;
    h->own_stream = false;
    h->epoch += 1;
    if (h->poisson.plans) { cufftSetStream(h->poisson.plan_fwd, h->stream); cufftSetStream(h->poisson.plan_inv, h->stream); }
    if (h->poisson.plans2) { cufftSetStream(h->poisson.plan2_fwd, h->stream); cufftSetStream(h->poisson.plan2_inv, h->stream); }
    return EK_OK;
}

ek_status ek_switch_stream(ek_handle *h, void *stream)
{
    if (!h || h->own_stream) return EK_ERR_STATE;
    h->stream = (cudaStream_t)stream;
    return EK_OK;
}

long long ek_halo_doubles(ek_handle *h) { return h ? (long long)4 * 9 * h->c.NY * h->c.NZ : 0; }
int ek_lbm_parity(ek_handle *h) { return h ? h->parity : -1; }

// phase 0 = A (call after an even step), phase 1 = B (after an odd step)
ek_status ek_halo_pack(ek_handle *h, int phase, double *to_left, double *to_right)
{
    if (!h || !h->slab || !h->pops_ready) return EK_ERR_STATE;
    DeviceGuard g(h->device);
    const EkConst &c = h->c;
    dim3 b(128), gr((c.NY + 127) / 128, c.NZ, 72);
    Lat4 l = current_lattice(h);
    HaloFaces f;
    f.buf[0] = to_right; f.buf[1] = to_left;
    if (phase == 0) {
        f.col[0] = c.NX - 1; f.plus[0] = 0;   // last column, c_x = -1 slots
        f.col[1] = 0;        f.plus[1] = 1;   // first column, c_x = +1 slots
    } else {
        f.col[0] = c.xhi;    f.plus[0] = 1;   // right ghost, c_x = +1 slots
        f.col[1] = c.xlo;    f.plus[1] = 0;   // left ghost, c_x = -1 slots
    }
    k_halo_column<true><<<gr, b, 0, h->stream>>>(c, l, f);
    EK_CUDA(h, cudaGetLastError());
    return EK_OK;
}

ek_status ek_halo_unpack(ek_handle *h, int phase, const double *from_left, const double *from_right)
{
    if (!h || !h->slab || !h->pops_ready) return EK_ERR_STATE;
    DeviceGuard g(h->device);
    const EkConst &c = h->c;
    dim3 b(128), gr((c.NY + 127) / 128, c.NZ, 72);
    Lat4 l = current_lattice(h);
    HaloFaces f;
    f.buf[0] = const_cast<double *>(from_left); f.buf[1] = const_cast<double *>(from_right);
    if (phase == 0) {
        f.col[0] = c.xlo; f.plus[0] = 0;
        f.col[1] = c.xhi; f.plus[1] = 1;
    } else {
        f.col[0] = 0;        f.plus[0] = 1;
        f.col[1] = c.NX - 1; f.plus[1] = 0;
    }
    k_halo_column<false><<<gr, b, 0, h->stream>>>(c, l, f);
    EK_CUDA(h, cudaGetLastError());
    return EK_OK;
}

// phi: my first column goes to the left neighbour's right ghost, my last column
// to the right neighbour's left ghost (NY*NZ doubles each)
ek_status ek_phi_halo_pack_range(ek_handle *h, int z0, int z1, double *to_left, double *to_right)
{
    if (!h || !h->slab || !h->allocated) return EK_ERR_STATE;
    if (z0 < 0 || z1 > h->c.NZ || z1 <= z0) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    const EkConst &c = h->c;
    dim3 b(128), gr((c.NY + 127) / 128, z1 - z0, 2);
    PhiFaces f;
    f.col[0] = 0; f.buf[0] = to_left;
    f.col[1] = c.NX - 1; f.buf[1] = to_right;
    k_phi_column<true><<<gr, b, 0, h->stream>>>(c, h->fld[EK_PHI], f, z0);
    EK_CUDA(h, cudaGetLastError());
    return EK_OK;
}

ek_status ek_phi_halo_unpack_range(ek_handle *h, int z0, int z1, const double *from_left, const double *from_right)
{
    if (!h || !h->slab || !h->allocated) return EK_ERR_STATE;
    if (z0 < 0 || z1 > h->c.NZ || z1 <= z0) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    const EkConst &c = h->c;
    dim3 b(128), gr((c.NY + 127) / 128, z1 - z0, 2);
    PhiFaces f;
    f.col[0] = c.xlo; f.buf[0] = const_cast<double *>(from_left);
    f.col[1] = c.xhi; f.buf[1] = const_cast<double *>(from_right);
    k_phi_column<false><<<gr, b, 0, h->stream>>>(c, h->fld[EK_PHI], f, z0);
    EK_CUDA(h, cudaGetLastError());
    return EK_OK;
}

ek_status ek_phi_halo_pack(ek_handle *h, double *to_left, double *to_right)
{
    return h ? ek_phi_halo_pack_range(h, 0, h->c.NZ, to_left, to_right) : EK_ERR_INVALID;
}

ek_status ek_phi_halo_unpack(ek_handle *h, const double *from_left, const double *from_right)
{
    return h ? ek_phi_halo_unpack_range(h, 0, h->c.NZ, from_left, from_right) : EK_ERR_INVALID;
}

ek_status ek_dq_ptr(ek_handle *h, double **p)
{
    if (!h || !p) return EK_ERR_INVALID;
    if (!h->allocated) return EK_ERR_STATE;
    *p = h->dq;
    return EK_OK;
}

int ek_row_pitch(ek_handle *h) { return h ? h->c.PX : 0; }

ek_status ek_ensure_allocated(ek_handle *h)
{
    if (!h) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    return ek_alloc_state(h);
}

// z-solve of the distributed Poisson stage.  spec: [NZ-2][kyl][NXg] complex
// (interleaved), the full complex-to-complex x-spectrum of the ky rows
// ky0 .. ky0+kyl-1 of the y-half-spectrum; in place.  Result scaled by
// 1/(NXg*NY) for the unnormalised inverse transforms.
ek_status ek_zsolve_columns(ek_handle *h, double *spec, int ky0, int kyl)
{
    if (!h || !spec || kyl < 0) return EK_ERR_INVALID;
    if (kyl == 0) return EK_OK;
    DeviceGuard g(h->device);
    const EkConst &c = h->c;
    const int M = c.NZ - 2, NXg = h->NXg ? h->NXg : c.NX;
    const int ncols = kyl * NXg;
    if (h->cp_ky0 != ky0 || h->cp_kyl != kyl) {
        cudaFree(h->cp_cols);
        h->cp_cols = nullptr;
        EK_CUDA(h, cudaMalloc((void **)&h->cp_cols, (size_t)M * ncols * sizeof(double)));
        ek_launch_zfactor_cols(ncols, NXg, c.NY, ky0, M, h->p.Lx, h->p.Ly, c.dz, h->cp_cols, h->stream);
        h->cp_ky0 = ky0;
        h->cp_kyl = kyl;
    }
    const double nxy = (double)NXg * (double)c.NY;
    const double size = (double)((unsigned int)NXg * (unsigned int)c.NY * (unsigned int)(2 * (c.NZ - 1)));
    const double off = h->dc_mode == EK_DC_PRESCRIBED ? -h->dc_ghat0 / size : 0.0;
    ek_launch_zsolve(2 * ncols, ncols, M, spec, h->cp_cols, -(c.CtoC / c.eps) * c.dz * c.dz, -c.voltage * nxy,
                     -c.voltage2 * nxy, 1.0 / nxy, off, ky0 == 0 ? 0 : -1, h->stream);
    EK_CUDA(h, cudaGetLastError());
    h->poisson_launches += 1;
    return EK_OK;
}

// After the host wrote the interior planes of phi: wall planes, flags, E arrays.
ek_status ek_poisson_finish(ek_handle *h, int set_walls)
{
    if (!h) return EK_ERR_INVALID;
    if (!h->allocated) return EK_ERR_STATE;
    DeviceGuard g(h->device);
    if (set_walls || h->phi_walls_dirty) {
        // only needed when something other than the solver touched phi (start-up
        // relaxation, uploads): the solver itself never writes the wall planes
        h->phi_walls_dirty = false;
        ek_launch_set_walls(h->c, h->fld[EK_PHI], h->stream);
        h->poisson_launches += 1;
        EK_CUDA(h, cudaGetLastError());
    }
    h->e_from_arrays = false;
    h->efield_stale = true;
    return EK_OK;
}

// E arrays from phi (ghost columns of phi must be current in slab mode)
ek_status ek_compute_efield(ek_handle *h)
{
    if (!h) return EK_ERR_INVALID;
    if (!h->allocated) return EK_ERR_STATE;
    DeviceGuard g(h->device);
    ek_launch_efield(h->c, h->fld[EK_PHI], h->fld[EK_EX], h->fld[EK_EY], h->fld[EK_EZ], h->stream);
    h->efield_stale = false;
    h->poisson_launches += 1;
    EK_CUDA(h, cudaGetLastError());
    return EK_OK;
}

// the pieces of initialization() (LBM.cu:68-146) for a host-driven PB loop
ek_status ek_init_uniform(ek_handle *h)
{
    if (!h) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    ek_status st = ek_alloc_state(h);
    if (st != EK_OK) return st;
    ek_launch_initialization(h->c, h->p, h->fld, h->stream);
    h->phi_walls_dirty = true;
    EK_CUDA(h, cudaMemcpyAsync(h->phi_old, h->fld[EK_PHI], (size_t)h->c.N * sizeof(double), cudaMemcpyDeviceToDevice,
                               h->stream));
    return EK_OK;
}

ek_status ek_pbe(ek_handle *h)
{
    if (!h || !h->allocated) return EK_ERR_STATE;
    DeviceGuard g(h->device);
    ek_launch_pbe(h->c, h->p, h->fld[EK_PHI], h->fld[EK_CHARGE], h->fld[EK_CHARGEN], h->dq, h->stream);
    EK_CUDA(h, cudaGetLastError());
    return EK_OK;
}

ek_status ek_pbe_relax(ek_handle *h)
{
    if (!h || !h->allocated) return EK_ERR_STATE;
    DeviceGuard g(h->device);
    ek_launch_pbe_relax(h->c, h->p.PB_omega, h->fld[EK_PHI], h->phi_old, h->stream);
    h->phi_walls_dirty = true;
    EK_CUDA(h, cudaGetLastError());
    return EK_OK;
}

}  // extern "C"
