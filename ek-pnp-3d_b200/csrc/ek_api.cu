// ek_api.cu -- the C ABI of include/ek_b200.h: handle life cycle, start-up,
// the coupled step loop (main.cu:189-200 of the reference) and accessors.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include <vector>

#include "ek_handle.h"

void ek_set_error(ek_handle *h, const std::string &msg)
{
    if (h) h->err = msg;
}

namespace {

__global__ void k_dq_from_fields(EkConst c, const double *ch, const double *chn, double *dq)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= c.NX) return;
    const size_t i = (size_t)blockIdx.z * c.plane + blockIdx.y * c.PX + x;
    dq[(size_t)blockIdx.z * c.dq_sz + (size_t)blockIdx.y * c.dq_sy + x] = ch[i] - chn[i];
}

}  // namespace

// derived constants, each evaluated in the form the reference uses
void ek_compute_consts(const ek_params &p, EkConst &c, bool slab)
{
    memset(&c, 0, sizeof(c));
    c.NX = p.NX; c.NY = p.NY; c.NZ = p.NZ;
    if (!slab) {
        c.PX = (p.NX + 1) & ~1;  // even row pitch: every z-plane of a field array stays 16 B aligned for cuFFT
        c.xlo = p.NX - 1;        // periodic wrap inside the array
        c.xhi = 0;
    } else {
        // x-slab of a larger domain: two ghost columns after the NX owned ones
        c.PX = (p.NX + 2 + 1) & ~1;
        c.xhi = p.NX;            // column holding the right neighbour's x = 0
        c.xlo = p.NX + 1;        // column holding the left neighbour's last x
    }
    c.plane = (long long)c.NY * c.PX;
    c.N = (long long)c.NZ * c.plane;
    c.dq_sy = c.PX;
    c.dq_sz = c.plane;
    c.NXT = (c.PX + EK_TILE - 1) / EK_TILE;
    c.lrow = (unsigned)c.NXT * EK_TILE_ELEMS;
    c.lplane = (unsigned)c.NY * c.lrow;
    c.Nlat = (unsigned long long)c.NZ * c.NY * c.NXT * EK_TILE_ELEMS;
    const double dt = p.dt, cs2 = p.cs_square;
    c.cflinv = 1.0 / p.CFL;
    c.cflinv2 = c.cflinv * c.cflinv / cs2;
    c.inv_cs2 = 1.0 / cs2;
    c.cs_square = cs2;
    c.CFL = p.CFL;
    c.tfac = 1.0 / cs2 / p.CFL;
    c.dt = dt;
    c.CtoC = p.convertCtoCharge; c.Ext = p.Ext; c.exf = p.exf; c.eps = p.eps;
    c.rho0 = p.rho0; c.Ra = p.Ra; c.nu = p.nu; c.D = p.D;
    c.K = p.K; c.Kn = p.Kn;
    c.w[0] = p.w0; c.w[1] = p.ws; c.w[2] = p.wa; c.w[3] = p.wd;
    for (int k = 0; k < 4; ++k) c.coe[k] = c.w[k] / cs2;
    // LBM.cu:488-495
    const double omega_plus = 1.0 / (p.nu / cs2 / dt + 1.0 / 2.0) / dt;
    const double omega_minus = 1.0 / (p.V / (p.nu / cs2 / dt) + 1.0 / 2.0) / dt;
    const double omega_c_minus = 1.0 / (p.diffu / cs2 / dt + 1.0 / 2.0) / dt;
    const double omega_c_plus = 1.0 / (p.VC / (p.diffu / cs2 / dt) + 1.0 / 2.0) / dt;
    const double omega_cn_minus = 1.0 / (p.diffun / cs2 / dt + 1.0 / 2.0) / dt;
    const double omega_cn_plus = 1.0 / (p.VCn / (p.diffun / cs2 / dt) + 1.0 / 2.0) / dt;
    const double omega_T_minus = 1.0 / (p.D / cs2 / dt + 1.0 / 2.0) / dt;
    const double omega_T_plus = 1.0 / (p.VT / (p.D / cs2 / dt) + 1.0 / 2.0) / dt;
    // LBM.cu:1700-1707
    c.wp[0] = omega_plus * dt;    c.wm[0] = omega_minus * dt;
    c.wp[1] = omega_c_plus * dt;  c.wm[1] = omega_c_minus * dt;
    c.wp[2] = omega_cn_plus * dt; c.wm[2] = omega_cn_minus * dt;
    c.wp[3] = omega_T_plus * dt;  c.wm[3] = omega_T_minus * dt;
    // LBM.cu:1660-1661
    c.sp = 1.0 - 0.5 * dt * omega_plus;
    c.sm = 1.0 - 0.5 * dt * omega_minus;
    // LBM.cu:1896-1898
    c.multi[0] = 0.0;
    c.multi[1] = 2.0 * p.rho0 * p.uw / cs2 * p.ws / p.CFL;
    c.multi[2] = 2.0 * p.rho0 * p.uw / cs2 * p.wa / p.CFL;
    c.multi[3] = 2.0 * p.rho0 * p.uw / cs2 * p.wd / p.CFL;
    // LBM.cu:2226-2229
    for (int k = 0; k < 4; ++k) c.twoTw[k] = 2.0 * p.TH * c.w[k];
    c.dx = p.dx; c.dy = p.dy; c.dz = p.dz;
    c.voltage = p.voltage; c.voltage2 = p.voltage2;
}

// z-planes a CTA walks.  Large grids: 8 (round 2, with the prefetching odd kernel at 256^3: 5.09 ms per step
// against 5.13 at 16, 5.20 at 32, 5.31 at 64; no difference on 1024-wide rows; round 1's kernel had its
// optimum at 16).  Small grids need more CTAs than (x-tiles * rows) to fill 148 SMs x 4 CTAs:
// 50x8x51 runs its LBM pass in 16 us with 2 planes per CTA against 87 us with 32.  At least 2,
// so that the owner of the z = 0 node also owns z = 1 (LBM.cu:663-801).
int ek_auto_zchunk(const EkConst &c)
{
    const long long cols = (long long)((c.NX + 31) / 32) * c.NY;
    long long z = cols * c.NZ / (148LL * 4 * 8);
    if (z < 2) z = 2;
    if (z > 8) z = 8;
    return (int)z;
}

namespace {

void free_state(ek_handle *h)
{
    for (int l = 0; l < 2; ++l)
        for (int s = 0; s < 4; ++s) { cudaFree(h->lat[l][s]); h->lat[l][s] = nullptr; }
    cudaFree(h->wall); h->wall = nullptr;
    for (int k = 0; k < EK_NFIELDS; ++k) {
        if (!h->fld_external[k]) cudaFree(h->fld[k]);
        h->fld[k] = nullptr;
        h->fld_external[k] = false;
    }
    cudaFree(h->dq); h->dq = nullptr;
    cudaFree(h->phi_old); h->phi_old = nullptr;
    cudaFree(h->cp_cols); h->cp_cols = nullptr; h->cp_ky0 = -1; h->cp_kyl = 0;
    ek_poisson_destroy(h->poisson);
    ek_slab_poisson_destroy(h);
    h->allocated = false;
}

void collect_events(ek_handle *h)
{
    for (size_t i = 0; i < h->ev_lbm.size(); ++i) {
        auto &e = h->ev_lbm[i];
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.first, e.second);
        h->lbm_ms += ms;
        h->lbm_ms_mode[h->ev_lbm_mode[i]] += ms;
        cudaEventDestroy(e.first); cudaEventDestroy(e.second);
    }
    h->ev_lbm.clear();
    h->ev_lbm_mode.clear();
    for (auto &e : h->ev_poi) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.first, e.second);
        h->poisson_ms += ms;
        cudaEventDestroy(e.first); cudaEventDestroy(e.second);
    }
    h->ev_poi.clear();
}

}  // namespace

ek_status ek_alloc_state(ek_handle *h)
{
    if (h->allocated) return EK_OK;
    const size_t N = (size_t)h->c.N;
    const size_t Nlat = (size_t)h->c.Nlat;
    const int nlat = h->stream_mode == EK_STREAM_PUSH ? 2 : 1;
    for (int l = 0; l < nlat; ++l)
        for (int s = 0; s < 4; ++s) EK_CUDA(h, cudaMalloc((void **)&h->lat[l][s], Nlat * sizeof(double)));
    EK_CUDA(h, cudaMalloc((void **)&h->wall, (size_t)3 * 2 * 27 * h->c.plane * sizeof(double)));
    for (int k = 0; k < EK_NFIELDS; ++k) {
        EK_CUDA(h, cudaMalloc((void **)&h->fld[k], N * sizeof(double)));
        EK_CUDA(h, cudaMemsetAsync(h->fld[k], 0, N * sizeof(double), h->stream));
    }
    EK_CUDA(h, cudaMalloc((void **)&h->dq, N * sizeof(double)));
    EK_CUDA(h, cudaMalloc((void **)&h->phi_old, N * sizeof(double)));
    ek_status st = ek_poisson_create(h, h->poisson, h->p, h->c.PX, h->stream);
    if (st != EK_OK) return st;
    h->allocated = true;
    return EK_OK;
}

StepArgs ek_step_args(ek_handle *h)
{
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.c = h->c;
    const bool push = h->stream_mode == EK_STREAM_PUSH;
    for (int s = 0; s < 4; ++s) {
        a.in[s] = h->lat[push ? h->cur : 0][s];
        a.out[s] = h->lat[push ? (h->cur ^ 1) : 0][s];
    }
    a.wall = h->wall;
    a.phi = h->fld[EK_PHI];
    a.E[0] = h->fld[EK_EX]; a.E[1] = h->fld[EK_EY]; a.E[2] = h->fld[EK_EZ];
    a.dq = h->dq;
    a.fld[0] = h->fld[EK_RHO]; a.fld[1] = h->fld[EK_UX]; a.fld[2] = h->fld[EK_UY]; a.fld[3] = h->fld[EK_UZ];
    a.fld[4] = h->fld[EK_CHARGE]; a.fld[5] = h->fld[EK_CHARGEN]; a.fld[6] = h->fld[EK_T];
    a.zchunk = h->zchunk;
    a.row_imm = h->kernel == 0 ? 1 : 0;   // kernel 4: the generic lean kernel for the odd step too (A/B partner)
    return a;
}

extern "C" {

int ek_abi_version(void) { return EK_B200_ABI_VERSION; }

// 1: this is the cross-check build (extra kernel variants, Poisson path 1), 0: the product library
int ek_is_xcheck_build(void)
{
#ifdef EK_XCHECK
    return 1;
#else
    return 0;
#endif
}

int ek_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void ek_default_params(ek_params *p)
{
    // LBM.h:29-125 as shipped
    p->NX = 50; p->NY = 8; p->NZ = 51;
    p->Lx = 0.5e-6; p->Ly = 0.08e-6; p->Lz = 0.5e-6;
    p->dx = 1.0e-6 / 100.0; p->dy = 1.0e-6 / 100.0; p->dz = 1.0e-6 / 100.0;
    p->uw = 0.0; p->exf = 0.0;
    p->CFL = 0.01;
    p->dt = 0.01 * 1.0e-6 / 100.0;
    p->cs_square = 1.0 / 3.0 / (0.01 * 0.01);
    p->rho0 = 1000.0;
    p->chargeinf = 0.01;
    p->voltage = -5.2574e-3; p->voltage2 = -5.2574e-3;
    p->Ext = 1.0e4; p->eps = 6.95e-10;
    p->diffu = 1.0e-8; p->nu = 0.889e-6; p->K = 4.245e-7;
    p->diffun = 1.0e-8; p->Kn = -4.245e-7;
    p->kB = 1.38e-23; p->electron = 1.6e-19; p->roomT = 273.0;
    p->convertCtoCharge = 9.64e4; p->PB_omega = 0.05;
    p->D = 0.889e-6; p->Ra = 1; p->TH = 1;
    p->w0 = 8.0 / 27.0; p->ws = 2.0 / 27.0; p->wa = 1.0 / 54.0; p->wd = 1.0 / 216.0;
    p->V = 1.0 / 12.0; p->VC = 1.0e-6; p->VCn = 1.0e-6; p->VT = 1.0 / 12.0;
    p->pb_iters = 501;
}

ek_status ek_create(const ek_params *p, int device, ek_handle **out)
{
    if (!p || !out) return EK_ERR_INVALID;
    *out = nullptr;
    if (p->NX < 2 || p->NY < 1 || p->NZ < 5) return EK_ERR_INVALID;
    if (p->NY > 65535 || p->NZ > 65535) return EK_ERR_INVALID;
    if ((long long)p->NX * p->NY * p->NZ >= (1LL << 31)) return EK_ERR_INVALID;  // field index is 32-bit
    // lattice element offsets are unsigned 32-bit (159 M cells per device)
    if ((unsigned long long)p->NZ * p->NY * ((p->NX + 2 + 31) / 32) * EK_TILE_ELEMS >= (1ULL << 32)) return EK_ERR_INVALID;
    if (!(p->dt > 0) || !(p->CFL > 0) || !(p->cs_square > 0) || !(p->dz > 0)) return EK_ERR_INVALID;
    ek_handle *h = new (std::nothrow) ek_handle();
    if (!h) return EK_ERR_NOMEM;
    h->p = *p;
    ek_compute_consts(*p, h->c, false);
    h->zchunk = ek_auto_zchunk(h->c);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        // no silent CPU path: the product requires a CUDA device
        delete h;
        return EK_ERR_CUDA;
    }
    if (device < 0) cudaGetDevice(&h->device); else h->device = device;
    if (h->device >= ndev) { delete h; return EK_ERR_INVALID; }
    DeviceGuard g(h->device);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return EK_ERR_CUDA; }
    *out = h;
    return EK_OK;
}

ek_status ek_destroy(ek_handle *h)
{
    if (!h) return EK_OK;
    DeviceGuard g(h->device);
    cudaStreamSynchronize(h->stream);
    collect_events(h);
    if (h->pair_graph) cudaGraphExecDestroy(h->pair_graph);
    for (cudaEvent_t e : h->job_events) cudaEventDestroy(e);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    free_state(h);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return EK_OK;
}

const char *ek_last_error(ek_handle *h) { return h ? h->err.c_str() : "null handle"; }
void *ek_stream(ek_handle *h) { return h ? (void *)h->stream : nullptr; }

ek_status ek_set_option(ek_handle *h, const char *key, long long value)
{
    if (!h || !key) return EK_ERR_INVALID;
    h->epoch += 1;   // whatever changes may be baked into the step graph
    if (!strcmp(key, "graph")) {
        if (value < -1 || value > 1) return EK_ERR_INVALID;
        h->graph_opt = (int)value;
        return EK_OK;
    }
    if (!strcmp(key, "stream_mode")) {
        if (h->allocated) { ek_set_error(h, "stream_mode must be set before the first init/set_fields"); return EK_ERR_STATE; }
        if (value != EK_STREAM_AA && value != EK_STREAM_PUSH) return EK_ERR_INVALID;
        if (h->slab && value != EK_STREAM_AA) { ek_set_error(h, "x-slabs use the A-A scheme (halo phases, ek_slab.cu)"); return EK_ERR_STATE; }
        h->stream_mode = (int)value;
        return EK_OK;
    }
    if (!strcmp(key, "zchunk")) {
        if (value != 0 && value < 2) return EK_ERR_INVALID;  // the owner of z = 0 must own z = 1
        if (h->sp.ready) {   // the chunks of the distributed Poisson stage are groups of z-blocks of this size
            ek_set_error(h, "zchunk cannot change after ek_slab_poisson_setup");
            return EK_ERR_STATE;
        }
        h->zchunk = value == 0 ? ek_auto_zchunk(h->c) : (int)value;   // 0: automatic
        return EK_OK;
    }
    if (!strcmp(key, "profile")) { h->profile = value != 0; return EK_OK; }
    if (!strcmp(key, "kernel")) {
        // 0: four warps, lean deep-interior path (default); 1: eight warps; 2: five warps;
        // 3: four warps, general path everywhere (cross-check of the lean path)
        // 5: x-marching rows with sector-aligned stores for the odd A-A step (measured slower than the
        // z-walking default; kept selectable for the A/B profile)
        if (value < 0 || value > 6) return EK_ERR_INVALID;   // 4: generic lean kernel (no row-stride immediates); 5/6: marching
#ifndef EK_XCHECK
        if (value == 1 || value == 2 || value >= 5) { ek_set_error(h, "kernel variants 1, 2, 5, 6 are only in the cross-check build libek_b200_xcheck.so"); return EK_ERR_INVALID; }
#endif
        h->kernel = (int)value;
        return EK_OK;
    }
    if (!strcmp(key, "poisson_path")) {
        if (value != 0 && value != 1) return EK_ERR_INVALID;
#ifndef EK_XCHECK
        if (value == 1) { ek_set_error(h, "Poisson path 1 is only in the cross-check build libek_b200_xcheck.so"); return EK_ERR_INVALID; }
#endif
        h->poisson_path = (int)value;
        return EK_OK;
    }

    return EK_ERR_INVALID;
}

ek_status ek_set_poisson_dc(ek_handle *h, int mode, double ghat0)
{
    if (!h || mode < EK_DC_ZERO || mode > EK_DC_PRESCRIBED) return EK_ERR_INVALID;
#ifndef EK_XCHECK
    if (mode == EK_DC_LITERAL) { ek_set_error(h, "EK_DC_LITERAL needs the odd-extension transform of the cross-check build libek_b200_xcheck.so"); return EK_ERR_INVALID; }
#endif
    h->dc_mode = mode;
    h->dc_ghat0 = ghat0;
    h->epoch += 1;
    return EK_OK;
}

ek_status ek_sync(ek_handle *h)
{
    if (!h) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    EK_CUDA(h, cudaStreamSynchronize(h->stream));
    return EK_OK;
}

ek_status ek_get_counter(ek_handle *h, const char *key, double *value)
{
    if (!h || !key || !value) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    if (!strcmp(key, "steps")) { *value = (double)h->steps; return EK_OK; }
    if (!strcmp(key, "graph_replays")) { *value = (double)h->graph_replays; return EK_OK; }
    if (!strcmp(key, "zchunk")) { *value = (double)h->zchunk; return EK_OK; }
    if (!strcmp(key, "lbm_launches")) { *value = (double)h->lbm_launches; return EK_OK; }
    if (!strcmp(key, "poisson_launches")) { *value = (double)h->poisson_launches; return EK_OK; }
    if (!strcmp(key, "kernel_launches")) { *value = (double)(h->lbm_launches + h->poisson_launches); return EK_OK; }
    if (!strncmp(key, "lbm_ms", 6) || !strcmp(key, "poisson_ms")) {
        EK_CUDA(h, cudaStreamSynchronize(h->stream));
        collect_events(h);
        if (!strcmp(key, "lbm_ms")) *value = h->lbm_ms;
        else if (!strcmp(key, "lbm_ms_even")) *value = h->lbm_ms_mode[EK_MODE_AA_EVEN];
        else if (!strcmp(key, "lbm_ms_odd")) *value = h->lbm_ms_mode[EK_MODE_AA_ODD];
        else if (!strcmp(key, "lbm_ms_push")) *value = h->lbm_ms_mode[EK_MODE_PUSH];
        else if (!strcmp(key, "poisson_ms")) *value = h->poisson_ms;
        else return EK_ERR_INVALID;
        return EK_OK;
    }
    return EK_ERR_INVALID;
}

ek_status ek_reset_counters(ek_handle *h)
{
    if (!h) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    EK_CUDA(h, cudaStreamSynchronize(h->stream));
    collect_events(h);
    h->steps = h->lbm_launches = h->poisson_launches = 0;
    h->lbm_ms = h->poisson_ms = 0.0;
    h->lbm_ms_mode[0] = h->lbm_ms_mode[1] = h->lbm_ms_mode[2] = 0.0;
    return EK_OK;
}

// single-domain entry points on a slab of a multi-rank domain would solve a Poisson problem that is
// periodic over the LOCAL width: refuse instead of computing something silently wrong
static bool refuse_on_slab(ek_handle *h, const char *what)
{
    if (h->nranks <= 1) return false;
    ek_set_error(h, std::string(what) + ": x-slab of a multi-rank domain -- use ek_rank_* / ek_multi_* (distributed Poisson stage)");
    return true;
}

ek_status ek_init_fields(ek_handle *h)
{
    if (!h) return EK_ERR_INVALID;
    if (refuse_on_slab(h, "ek_init_fields")) return EK_ERR_STATE;
    DeviceGuard g(h->device);
    ek_status st = ek_alloc_state(h);
    if (st != EK_OK) return st;
    const EkConst &c = h->c;
    const size_t bytes = (size_t)c.N * sizeof(double);
    ek_launch_initialization(c, h->p, h->fld, h->stream);                       // LBM.cu:76
    h->phi_walls_dirty = true;
    EK_CUDA(h, cudaMemcpyAsync(h->phi_old, h->fld[EK_PHI], bytes, cudaMemcpyDeviceToDevice, h->stream));  // LBM.cu:82-86
    for (int i = 0; i < h->p.pb_iters; ++i) {                                     // LBM.cu:89
        ek_launch_pbe(c, h->p, h->fld[EK_PHI], h->fld[EK_CHARGE], h->fld[EK_CHARGEN], h->dq, h->stream);
        // E is only observable after the last solve (it is taken from the un-relaxed phi)
        const bool last = (i == h->p.pb_iters - 1);
        int n = 0;
        st = ek_poisson_solve(h, h->poisson, h->p, c, h->dq, h->fld[EK_PHI], last ? h->fld[EK_EX] : nullptr,
                              h->fld[EK_EY], h->fld[EK_EZ], h->poisson_path, h->dc_mode, h->dc_ghat0, h->stream, &n);  // LBM.cu:96
        if (st != EK_OK) return st;
        ek_launch_pbe_relax(c, h->p.PB_omega, h->fld[EK_PHI], h->phi_old, h->stream);  // LBM.cu:98-104
        h->phi_walls_dirty = true;   // the relaxation mixes the wall planes too: the next solve re-imposes them
    }
    EK_CUDA(h, cudaGetLastError());
    h->fields_ready = true;
    h->pops_ready = false;
    h->e_from_arrays = true;
    return EK_OK;
}

ek_status ek_set_fields(ek_handle *h, const double *const fields[EK_NFIELDS], int src_on_device)
{
    if (!h || !fields) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    ek_status st = ek_alloc_state(h);
    if (st != EK_OK) return st;
    const EkConst &c = h->c;
    for (int k = 0; k < EK_NFIELDS; ++k) {
        if (!fields[k]) continue;
        EK_CUDA(h, cudaMemcpy2DAsync(h->fld[k], (size_t)c.PX * sizeof(double), fields[k], (size_t)c.NX * sizeof(double),
                                     (size_t)c.NX * sizeof(double), (size_t)c.NY * c.NZ,
                                     src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    }
    EK_CUDA(h, cudaStreamSynchronize(h->stream));
    h->fields_ready = true;
    h->pops_ready = false;
    h->e_from_arrays = true;
    if (fields[EK_PHI]) h->phi_walls_dirty = true;
    return EK_OK;
}

ek_status ek_init_equilibrium(ek_handle *h)
{
    if (!h) return EK_ERR_INVALID;
    if (!h->fields_ready) { ek_set_error(h, "ek_init_equilibrium before ek_init_fields/ek_set_fields"); return EK_ERR_STATE; }
    DeviceGuard g(h->device);
    h->cur = 0;
    h->parity = 0;
    StepArgs a = ek_step_args(h);
    EK_CUDA(h, ek_launch_init_equilibrium(a, h->fld, h->stream));
    dim3 b(128), gr((h->c.NX + 127) / 128, h->c.NY, h->c.NZ);
    k_dq_from_fields<<<gr, b, 0, h->stream>>>(h->c, h->fld[EK_CHARGE], h->fld[EK_CHARGEN], h->dq);
    EK_CUDA(h, cudaGetLastError());
    h->pops_ready = true;
    return EK_OK;
}

ek_status ek_init(ek_handle *h)
{
    if (!h) return EK_ERR_INVALID;
    if (refuse_on_slab(h, "ek_init")) return EK_ERR_STATE;
    ek_status st = ek_init_fields(h);
    if (st != EK_OK) return st;
    return ek_init_equilibrium(h);
}

ek_status ek_stream_collide_save(ek_handle *h, int write_fields)
{
    return ek_stream_collide_save_range(h, write_fields, 0, 0, 1);
}

// One LBM pass restricted to the z-chunks [zblock0, zblock1) (chunks of "zchunk"
// planes; 0,0 = everything).  The launches of one pass may be issued in any
// order and on different streams (the A-A access sets of different nodes are
// disjoint); `last` != 0 on the final launch of the pass advances the parity.
ek_status ek_stream_collide_save_range(ek_handle *h, int write_fields, int zblock0, int zblock1, int last)
{
    return ek_stream_collide_save_part(h, write_fields, zblock0, zblock1, 0, last);
}

// The same restricted in x as well: xtiles = 0 every 32-column tile of a row, 1 the two boundary tiles (first
// and last), 2 the interior tiles.  The slab pipeline launches the boundary tiles of the whole pass first and
// sends the population halos under the launches of the interior tiles (default kernel only).
ek_status ek_stream_collide_save_part(ek_handle *h, int write_fields, int zblock0, int zblock1, int xtiles, int last)
{
    if (!h) return EK_ERR_INVALID;
    if (!h->pops_ready) { ek_set_error(h, "ek_stream_collide_save before ek_init_equilibrium"); return EK_ERR_STATE; }
    const int nblocks = (h->c.NZ + h->zchunk - 1) / h->zchunk;
    if (zblock0 < 0 || zblock1 > nblocks || (zblock1 != 0 && zblock1 <= zblock0)) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    StepArgs a = ek_step_args(h);
    a.zblock0 = zblock0;
    a.nzblocks = zblock1 > 0 ? zblock1 - zblock0 : 0;
    if (xtiles < 0 || xtiles > 2 || (xtiles != 0 && h->kernel != 0 && h->kernel != 3 && h->kernel != 4)) return EK_ERR_INVALID;
    a.xt_mode = xtiles;
    const int mode = h->stream_mode == EK_STREAM_PUSH ? EK_MODE_PUSH : (h->parity ? EK_MODE_AA_ODD : EK_MODE_AA_EVEN);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->profile) {
        EK_CUDA(h, cudaEventCreate(&e0)); EK_CUDA(h, cudaEventCreate(&e1));
        EK_CUDA(h, cudaEventRecord(e0, h->stream));
    }
#ifdef EK_XCHECK
    if (h->kernel == 1) EK_CUDA(h, ek_launch_step8(a, mode, write_fields != 0, h->e_from_arrays, h->stream));
    else if (h->kernel == 2) EK_CUDA(h, ek_launch_step5(a, mode, write_fields != 0, h->e_from_arrays, h->stream));
    else
#endif
#ifdef EK_XCHECK
    if (mode == EK_MODE_AA_ODD && (h->kernel == 5 || h->kernel == 6) && !h->e_from_arrays && ek_march_applicable(h->c)) {
        // odd step: x-marching rows with sector-aligned stores (ek_lbm.cu)
        const int z0 = zblock0 * h->zchunk;
        const int z1 = zblock1 > 0 ? (zblock1 * h->zchunk < h->c.NZ ? zblock1 * h->zchunk : h->c.NZ) : h->c.NZ;
        EK_CUDA(h, ek_launch_march(a, write_fields != 0, z0, z1, h->kernel == 6 ? 2 : 1, h->stream));
    } else
#endif
    EK_CUDA(h, ek_launch_step(a, mode, write_fields != 0, h->e_from_arrays, h->kernel != 3, h->stream));
    if (h->profile) {
        EK_CUDA(h, cudaEventRecord(e1, h->stream));
        h->ev_lbm.emplace_back(e0, e1);
        h->ev_lbm_mode.push_back(mode);
    }
    h->lbm_launches += 1;
    if (last) {
        if (h->stream_mode == EK_STREAM_PUSH) h->cur ^= 1; else h->parity ^= 1;
    }
    return EK_OK;
}

ek_status ek_fast_poisson(ek_handle *h, int write_efield)
{
    if (!h) return EK_ERR_INVALID;
    if (!h->allocated) { ek_set_error(h, "ek_fast_poisson before initialisation"); return EK_ERR_STATE; }
    if (h->nranks > 1) {
        ek_set_error(h, "x-slab of a multi-rank domain: the distributed solve is driven by the host (slab.py)");
        return EK_ERR_STATE;
    }
    DeviceGuard g(h->device);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->profile) {
        EK_CUDA(h, cudaEventCreate(&e0)); EK_CUDA(h, cudaEventCreate(&e1));
        EK_CUDA(h, cudaEventRecord(e0, h->stream));
    }
    int n = 0;
    ek_status st = ek_poisson_solve(h, h->poisson, h->p, h->c, h->dq, h->fld[EK_PHI], write_efield ? h->fld[EK_EX] : nullptr,
                                    h->fld[EK_EY], h->fld[EK_EZ], h->poisson_path, h->dc_mode, h->dc_ghat0, h->stream, &n);
    if (st != EK_OK) return st;
    if (h->profile) {
        EK_CUDA(h, cudaEventRecord(e1, h->stream));
        h->ev_poi.emplace_back(e0, e1);
    }
    h->poisson_launches += n;      // hand-written kernels only (the two cuFFT executions are library launches)
    h->e_from_arrays = false;      // from now on E = -grad(phi) of the fresh potential
    h->efield_stale = !write_efield;
    return EK_OK;
}

static ek_status one_step(ek_handle *h, int full)
{
    ek_status st = ek_stream_collide_save(h, full);
    if (st != EK_OK) return st;
    st = ek_fast_poisson(h, full);
    if (st != EK_OK) return st;
    h->steps += 1;
    return EK_OK;
}

// may the next two (non-final) steps run as one graph launch?
static bool graph_usable(ek_handle *h)
{
    const bool want = h->graph_opt == 1 || (h->graph_opt == -1 && (long long)h->c.NX * h->c.NY * h->c.NZ < (1LL << 22));
    // (plans and scratch of the Poisson path in use exist: nothing may allocate during the capture;
    // the legacy default stream of the shim cannot be captured)
    const bool path1 = h->poisson_path == 1 || h->dc_mode == EK_DC_LITERAL;
    return want && !h->profile && h->dc_mode != EK_DC_PRESCRIBED && !h->e_from_arrays && !h->phi_walls_dirty &&
           !h->slab && (path1 ? h->poisson.plans : h->poisson.plans2) && h->stream != nullptr &&
           // the pair was captured from the natural layout (A-A parity 0 / push lattice 0): replay from there only
           h->parity == 0 && h->cur == 0;
}

// capture two steps on the handle's stream (nothing executes), instantiate
static ek_status build_pair_graph(ek_handle *h)
{
    if (h->pair_graph) { cudaGraphExecDestroy(h->pair_graph); h->pair_graph = nullptr; }
    const int parity = h->parity, cur = h->cur;
    const long long steps = h->steps, l0 = h->lbm_launches, p0 = h->poisson_launches;
    EK_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
    ek_status st = one_step(h, 0);
    if (st == EK_OK) st = one_step(h, 0);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    // the capture only recorded: host-side state goes back to where it was
    h->pair_lbm_launches = h->lbm_launches - l0;
    h->pair_poisson_launches = h->poisson_launches - p0;
    h->parity = parity; h->cur = cur; h->steps = steps; h->lbm_launches = l0; h->poisson_launches = p0;
    if (st != EK_OK) { if (g) cudaGraphDestroy(g); return st; }
    EK_CUDA(h, e);
    e = cudaGraphInstantiate(&h->pair_graph, g, 0);
    cudaGraphDestroy(g);
    EK_CUDA(h, e);
    h->graph_epoch = h->epoch;
    return EK_OK;
}

ek_status ek_step(ek_handle *h, int nsteps)
{
    if (!h || nsteps < 0) return EK_ERR_INVALID;
    if (refuse_on_slab(h, "ek_step")) return EK_ERR_STATE;
    DeviceGuard g(h->device);
    int i = 0;
    while (i < nsteps) {
        // the last step of the call writes the macroscopic arrays; pairs of the others replay the graph
        if (nsteps - 1 - i >= 2 && h->pops_ready && graph_usable(h)) {
            if (!h->pair_graph || h->graph_epoch != h->epoch) {
                ek_status st = build_pair_graph(h);
                if (st != EK_OK) return st;
            }
            EK_CUDA(h, cudaGraphLaunch(h->pair_graph, h->stream));
            h->steps += 2;
            h->lbm_launches += h->pair_lbm_launches;
            h->poisson_launches += h->pair_poisson_launches;
            h->graph_replays += 1;
            i += 2;
            continue;
        }
        ek_status st = one_step(h, i == nsteps - 1);
        if (st != EK_OK) return st;
        ++i;
    }
    return EK_OK;
}

// ---------------------------------------------------------------------------
// One whole job on HOST arrays: what a caller with its fields in host memory does with
// ek_set_fields + ek_init_equilibrium + ek_step(n) + 11 x ek_get_field, as ONE call whose PCIe
// copies overlap the device work instead of bracketing it:
//   * the upload runs plane group by plane group on a copy stream; as soon as a group has landed its
//     populations are initialised and the FIRST LBM pass runs on it (the first pass is the even A-A
//     step, node-local, and takes E from the uploaded arrays, main.cu:174,192): the start-up kernels
//     and one of the n LBM passes hide under the upload;
//   * the LAST LBM pass is launched group by group and the rho, u, c+, c-, T planes it writes are
//     downloaded behind it; phi and E follow after the last solve.
// Same results as the plain sequence, bit for bit (tests).  Host buffers should be pinned.
// ---------------------------------------------------------------------------
static ek_status run_from_host_pipelined(ek_handle *h, const double *const in[EK_NFIELDS], int nsteps,
                                         double *const out[EK_NFIELDS], int nblocks);

ek_status ek_run_from_host(ek_handle *h, const double *const in[EK_NFIELDS], int nsteps, double *const out[EK_NFIELDS])
{
    if (!h || !in || !out || nsteps < 1) return EK_ERR_INVALID;
    for (int k = 0; k < EK_NFIELDS; ++k)
        if (!in[k] || !out[k]) return EK_ERR_INVALID;
    if (refuse_on_slab(h, "ek_run_from_host")) return EK_ERR_STATE;
    DeviceGuard g(h->device);
    ek_status st = ek_alloc_state(h);
    if (st != EK_OK) return st;
    const EkConst &c = h->c;
    const int nblocks = (c.NZ + h->zchunk - 1) / h->zchunk;
    // plain sequence when there is nothing to pipeline (two-lattice scheme, adopted arrays, one step, tiny grids)
    bool external = false;
    for (int k = 0; k < EK_NFIELDS; ++k) external = external || h->fld_external[k];
    if (h->stream_mode != EK_STREAM_AA || external || nsteps < 2 || nblocks < 4 || h->profile || h->stream == nullptr) {
        st = ek_set_fields(h, in, 0);
        if (st == EK_OK) st = ek_init_equilibrium(h);
        if (st == EK_OK) st = ek_step(h, nsteps);
        for (int k = 0; k < EK_NFIELDS && st == EK_OK; ++k) st = ek_get_field(h, k, out[k], 0);
        return st;
    }
    st = run_from_host_pipelined(h, in, nsteps, out, nblocks);
    if (st != EK_OK) {
        // a job that stopped half way leaves nothing to continue from: drain both streams, back to "not initialised"
        if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
        cudaStreamSynchronize(h->stream);
        h->fields_ready = false;
        h->pops_ready = false;
    }
    return st;
}

static ek_status run_from_host_pipelined(ek_handle *h, const double *const in[EK_NFIELDS], int nsteps,
                                         double *const out[EK_NFIELDS], int nblocks)
{
    const EkConst &c = h->c;
    ek_status st = EK_OK;
    if (!h->copy_stream) EK_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    const int G = nblocks < 8 ? nblocks : 8;      // plane groups (multiples of the LBM kernel's z-blocks)
    while ((int)h->job_events.size() < 2 * G + 2) {
        cudaEvent_t e;
        EK_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->job_events.push_back(e);
    }
    auto block_of = [&](int gi) { return (int)((long long)nblocks * gi / G); };
    auto plane_of = [&](int b) { return b * h->zchunk < c.NZ ? b * h->zchunk : c.NZ; };
    auto copy_planes = [&](int id, int z0, int z1, bool up) -> cudaError_t {
        const size_t rows = (size_t)c.NY * (z1 - z0);
        double *dev = h->fld[id] + (size_t)z0 * c.plane;
        if (up)
            return cudaMemcpy2DAsync(dev, (size_t)c.PX * sizeof(double), in[id] + (size_t)z0 * c.NY * c.NX,
                                     (size_t)c.NX * sizeof(double), (size_t)c.NX * sizeof(double), rows,
                                     cudaMemcpyHostToDevice, h->copy_stream);
        return cudaMemcpy2DAsync(out[id] + (size_t)z0 * c.NY * c.NX, (size_t)c.NX * sizeof(double), dev,
                                 (size_t)c.PX * sizeof(double), (size_t)c.NX * sizeof(double), rows, cudaMemcpyDeviceToHost,
                                 h->copy_stream);
    };
    // EK_JOB_DEBUG=1: synchronise after every phase and print its wall time (development aid)
    const bool dbg = getenv("EK_JOB_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_prev = now();
    auto phase = [&](const char *name) {
        if (!dbg) return;
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamSynchronize(h->stream);
        const double t = now();
        fprintf(stderr, "ek_run_from_host: %-28s %8.3f ms\n", name, t - t_prev);
        t_prev = t;
    };
    // the copy stream starts after whatever the handle's stream was doing with these arrays
    EK_CUDA(h, cudaEventRecord(h->job_events[2 * G], h->stream));
    EK_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->job_events[2 * G], 0));
    // ---- upload | init_equilibrium | first (even, node-local) LBM pass, group by group
    h->fields_ready = true;
    h->pops_ready = true;
    h->e_from_arrays = true;
    h->phi_walls_dirty = true;
    h->efield_stale = false;
    h->cur = 0;
    h->parity = 0;
    h->epoch += 1;
    StepArgs a = ek_step_args(h);
    for (int gi = 0; gi < G; ++gi) {
        const int b0 = block_of(gi), b1 = block_of(gi + 1);
        const int z0 = plane_of(b0), z1 = plane_of(b1);
        for (int k = 0; k < EK_NFIELDS; ++k) EK_CUDA(h, copy_planes(k, z0, z1, true));
        EK_CUDA(h, cudaEventRecord(h->job_events[gi], h->copy_stream));
        EK_CUDA(h, cudaStreamWaitEvent(h->stream, h->job_events[gi], 0));
        EK_CUDA(h, ek_launch_init_equilibrium_range(a, h->fld, z0, z1, h->stream));
        st = ek_stream_collide_save_range(h, 0, b0, b1, gi == G - 1);
        if (st != EK_OK) return st;
    }
    phase("upload+init+first LBM pass");
    st = ek_fast_poisson(h, 0);
    if (st != EK_OK) return st;
    h->steps += 1;
    phase("first Poisson solve");
    // ---- the steps in between
    if (nsteps > 2) {
        // (ek_step writes the macroscopic arrays on its last step; harmless here, the final pass rewrites them)
        st = ek_step(h, nsteps - 2);
        if (st != EK_OK) return st;
    }
    phase("steps in between");
    // ---- last LBM pass group by group, rho u c+ c- T downloaded behind it
    const int lbm_fields[7] = {EK_RHO, EK_UX, EK_UY, EK_UZ, EK_CHARGE, EK_CHARGEN, EK_T};
    for (int gi = 0; gi < G; ++gi) {
        const int b0 = block_of(gi), b1 = block_of(gi + 1);
        st = ek_stream_collide_save_range(h, 1, b0, b1, gi == G - 1);
        if (st != EK_OK) return st;
        EK_CUDA(h, cudaEventRecord(h->job_events[G + gi], h->stream));
        EK_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->job_events[G + gi], 0));
        for (int k = 0; k < 7; ++k) EK_CUDA(h, copy_planes(lbm_fields[k], plane_of(b0), plane_of(b1), false));
    }
    phase("last LBM pass + 7 downloads");
    st = ek_fast_poisson(h, 1);
    if (st != EK_OK) return st;
    h->steps += 1;
    EK_CUDA(h, cudaEventRecord(h->job_events[2 * G + 1], h->stream));
    EK_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->job_events[2 * G + 1], 0));
    const int poisson_fields[4] = {EK_PHI, EK_EX, EK_EY, EK_EZ};
    for (int k = 0; k < 4; ++k) EK_CUDA(h, copy_planes(poisson_fields[k], 0, c.NZ, false));
    EK_CUDA(h, cudaStreamSynchronize(h->copy_stream));
    EK_CUDA(h, cudaStreamSynchronize(h->stream));
    phase("last solve + 4 downloads");
    return EK_OK;
}

ek_status ek_step_timed(ek_handle *h, int nsteps, float *ms)
{
    if (!h || !ms) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    cudaEvent_t e0, e1;
    EK_CUDA(h, cudaEventCreate(&e0));
    EK_CUDA(h, cudaEventCreate(&e1));
    EK_CUDA(h, cudaStreamSynchronize(h->stream));
    EK_CUDA(h, cudaEventRecord(e0, h->stream));
    ek_status st = ek_step(h, nsteps);
    cudaError_t e = cudaEventRecord(e1, h->stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (st != EK_OK) return st;
    EK_CUDA(h, e);
    return EK_OK;
}

ek_status ek_adopt_field(ek_handle *h, int id, double *dev_ptr)
{
    if (!h || !dev_ptr || id < 0 || id >= EK_NFIELDS) return EK_ERR_INVALID;
    if (h->c.PX != h->c.NX) { ek_set_error(h, "ek_adopt_field needs an even NX (dense rows)"); return EK_ERR_INVALID; }
    if (h->allocated && h->fld[id] == dev_ptr) return EK_OK;
    DeviceGuard g(h->device);
    ek_status st = ek_alloc_state(h);
    if (st != EK_OK) return st;
    EK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (!h->fld_external[id]) cudaFree(h->fld[id]);
    h->fld[id] = dev_ptr;
    h->fld_external[id] = true;
    h->epoch += 1;
    if (id == EK_PHI) h->phi_walls_dirty = true;
    return EK_OK;
}

ek_status ek_refresh_charge_difference(ek_handle *h)
{
    if (!h) return EK_ERR_INVALID;
    if (!h->allocated) return EK_ERR_STATE;
    DeviceGuard g(h->device);
    dim3 b(128), gr((h->c.NX + 127) / 128, h->c.NY, h->c.NZ);
    k_dq_from_fields<<<gr, b, 0, h->stream>>>(h->c, h->fld[EK_CHARGE], h->fld[EK_CHARGEN], h->dq);
    EK_CUDA(h, cudaGetLastError());
    return EK_OK;
}

ek_status ek_mark_fields_ready(ek_handle *h)
{
    if (!h) return EK_ERR_INVALID;
    if (!h->allocated) return EK_ERR_STATE;
    h->fields_ready = true;
    h->pops_ready = false;
    h->e_from_arrays = true;
    h->efield_stale = false;
    return EK_OK;
}

ek_status ek_field_ptr(ek_handle *h, int id, double **dev_ptr)
{
    if (!h || !dev_ptr || id < 0 || id >= EK_NFIELDS) return EK_ERR_INVALID;
    if (!h->allocated) return EK_ERR_STATE;
    *dev_ptr = h->fld[id];
    return EK_OK;
}

ek_status ek_get_field(ek_handle *h, int id, double *dst, int dst_on_device)
{
    if (!h || !dst || id < 0 || id >= EK_NFIELDS) return EK_ERR_INVALID;
    if (!h->allocated) { ek_set_error(h, "ek_get_field before initialisation"); return EK_ERR_STATE; }
    DeviceGuard g(h->device);
    const EkConst &c = h->c;
    if (id >= EK_EX && h->efield_stale) {
        ek_launch_efield(c, h->fld[EK_PHI], h->fld[EK_EX], h->fld[EK_EY], h->fld[EK_EZ], h->stream);
        h->efield_stale = false;
    }
    EK_CUDA(h, cudaMemcpy2DAsync(dst, (size_t)c.NX * sizeof(double), h->fld[id], (size_t)c.PX * sizeof(double),
                                 (size_t)c.NX * sizeof(double), (size_t)c.NY * c.NZ,
                                 dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    EK_CUDA(h, cudaStreamSynchronize(h->stream));
    return EK_OK;
}

ek_status ek_get_populations(ek_handle *h, int set, double *dst, int dst_on_device)
{
    if (!h || !dst || set < 0 || set >= EK_NSETS) return EK_ERR_INVALID;
    if (!h->pops_ready) { ek_set_error(h, "ek_get_populations before ek_init_equilibrium"); return EK_ERR_STATE; }
    DeviceGuard g(h->device);
    const size_t cells = (size_t)h->c.NX * h->c.NY * h->c.NZ;
    StepArgs a = ek_step_args(h);
    const int mode = (h->stream_mode == EK_STREAM_AA && h->parity) ? EK_MODE_AA_ODD : EK_MODE_AA_EVEN;
    double *buf = dst;
    if (!dst_on_device) EK_CUDA(h, cudaMalloc((void **)&buf, 27 * cells * sizeof(double)));
    cudaError_t e = ek_launch_export(a, mode, set, buf, h->stream);
    if (e == cudaSuccess && !dst_on_device)
        e = cudaMemcpyAsync(dst, buf, 27 * cells * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (!dst_on_device) cudaFree(buf);
    EK_CUDA(h, e);
    return EK_OK;
}

// Restore one population set from host memory: (27, NZ, NY, NX) pre-collision values in the
// reference's order, as ek_get_populations() returns them.  The lattice goes back to the
// natural layout (A-A parity 0); call ek_populations_restored() after the four sets.
ek_status ek_set_populations(ek_handle *h, int set, const double *src)
{
    if (!h || !src || set < 0 || set >= EK_NSETS) return EK_ERR_INVALID;
    DeviceGuard g(h->device);
    ek_status st = ek_alloc_state(h);
    if (st != EK_OK) return st;
    const size_t cells = (size_t)h->c.NX * h->c.NY * h->c.NZ;
    h->cur = 0;
    h->parity = 0;
    StepArgs a = ek_step_args(h);
    double *buf = nullptr;
    EK_CUDA(h, cudaMalloc((void **)&buf, 27 * cells * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(buf, src, 27 * cells * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = ek_launch_import(a, set, buf, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(buf);
    EK_CUDA(h, e);
    return EK_OK;
}

// the state machine after ek_set_fields() + 4 x ek_set_populations(): a run that continues
ek_status ek_populations_restored(ek_handle *h)
{
    if (!h) return EK_ERR_INVALID;
    if (!h->allocated || !h->fields_ready) { ek_set_error(h, "ek_populations_restored before the fields were set"); return EK_ERR_STATE; }
    h->cur = 0;
    h->parity = 0;
    h->pops_ready = true;
    h->e_from_arrays = false;   // E = -grad(phi) of the restored potential, as in the run that was saved
    h->efield_stale = false;
    return EK_OK;
}

}  // extern "C"
