// ek_multi.cu -- native (C++) driver of the x-slab path: ONE host process, one slab handle
// per GPU, no Python and no NCCL.  It is the multi-GPU form of ek_create/ek_init/ek_step for a
// C++ caller such as the reference's main() (the reference is single-GPU, main.cu:58; SURVEY.md
// 8b asked for an `ek_create(params, ndev, devs)`).
//
// The per-slab work is the C ABI of ek_slab.cu / ek_slab_poisson.cu; what this file adds is the
// orchestration that ek-pnp-3d_b200/slab.py does with torch.distributed, expressed with CUDA
// peer access inside one process:
//   * population and phi halos: pack kernel -> cudaMemcpyAsync between devices -> unpack kernel;
//   * Poisson transposes: the re-blocking pushes of ek_slab_poisson.cu straight into the peers'
//     buffers (copy engines or kernel), no intermediate send/receive copies;
//   * cross-device ordering: CUDA events (every stream waits for every other stream's event)
//     where the multi-process path uses a one-element all-reduce.
// The slabs of one ek_multi may also share a device (used by the single-GPU tests).
#include <string.h>

#include <vector>

#include "ek_handle.h"

struct ek_multi {
    ek_params global;
    int P = 0;
    std::vector<ek_handle *> h;
    std::vector<int> dev;
    std::vector<cudaEvent_t> ev;          // one per slab, re-recorded at every barrier
    std::vector<cudaEvent_t> ev2;         // second set (side / halo streams)
    std::vector<cudaStream_t> side, halo; // per slab: Poisson forward half behind the LBM launches; halos
    bool pipeline = true;
    // halo buffers on each slab's device
    std::vector<double *> to_l, to_r, from_l, from_r;       // populations: ek_halo_doubles() each
    std::vector<double *> pto_l, pto_r, pfrom_l, pfrom_r;   // phi: NY*NZ each
    int K = 1;                            // Poisson chunks
    bool pops = false;
    std::string err;
};

namespace {

ek_status fail(ek_multi *m, ek_handle *h, const char *what, ek_status st)
{
    m->err = std::string(what) + ": " + (h ? ek_last_error(h) : "");
    return st;
}

#define MK(m, hh, call)                                            \
    do {                                                           \
        ek_status _s = (call);                                     \
        if (_s != EK_OK) return fail((m), (hh), #call, _s);        \
    } while (0)

#define MCUDA(m, call)                                                             \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e != cudaSuccess) {                                                   \
            (m)->err = std::string(#call) + ": " + cudaGetErrorString(_e);         \
            return EK_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)

// RAII: the calls of the C ABI made in this scope run on another stream of the slab's device
struct OnStream {
    ek_handle *h;
    cudaStream_t saved;
    OnStream(ek_handle *hh, cudaStream_t st) : h(hh), saved(hh->stream) { h->stream = st; }
    ~OnStream() { h->stream = saved; }
};

// stream `a` waits for everything issued so far on stream `b` (same or another device)
ek_status wait_for(ek_multi *m, int dev_a, cudaStream_t a, int dev_b, cudaStream_t b, cudaEvent_t ev)
{
    {
        DeviceGuard g(dev_b);
        MCUDA(m, cudaEventRecord(ev, b));
    }
    DeviceGuard g(dev_a);
    MCUDA(m, cudaStreamWaitEvent(a, ev, 0));
    return EK_OK;
}

// every slab's stream waits until every slab's stream has reached this point
ek_status barrier_all(ek_multi *m)
{
    for (int s = 0; s < m->P; ++s) {
        DeviceGuard g(m->dev[s]);
        MCUDA(m, cudaEventRecord(m->ev[s], m->h[s]->stream));
    }
    for (int s = 0; s < m->P; ++s) {
        DeviceGuard g(m->dev[s]);
        for (int t = 0; t < m->P; ++t)
            if (t != s) MCUDA(m, cudaStreamWaitEvent(m->h[s]->stream, m->ev[t], 0));
    }
    return EK_OK;
}

// ring exchange of per-slab buffers: from_l[s] <- to_r[s-1], from_r[s] <- to_l[s+1]
ek_status ring_exchange(ek_multi *m, std::vector<double *> &to_l, std::vector<double *> &to_r,
                        std::vector<double *> &from_l, std::vector<double *> &from_r, size_t offset, size_t count)
{
    MK(m, nullptr, barrier_all(m));   // the neighbours' pack kernels are done
    for (int s = 0; s < m->P; ++s) {
        const int l = (s + m->P - 1) % m->P, r = (s + 1) % m->P;
        DeviceGuard g(m->dev[s]);
        MCUDA(m, cudaMemcpyAsync(from_l[s] + offset, to_r[l] + offset, count * sizeof(double), cudaMemcpyDefault,
                                 m->h[s]->stream));
        MCUDA(m, cudaMemcpyAsync(from_r[s] + offset, to_l[r] + offset, count * sizeof(double), cudaMemcpyDefault,
                                 m->h[s]->stream));
    }
    // (the senders re-pack these buffers only in the next step, several barriers later)
    return EK_OK;
}

// population halos on the halo streams, next to whatever the main streams do meanwhile
ek_status halo_exchange_async(ek_multi *m, int phase)
{
    const size_t n = (size_t)ek_halo_doubles(m->h[0]);
    for (int s = 0; s < m->P; ++s) {   // after this slab's LBM pass
        MK(m, nullptr, wait_for(m, m->dev[s], m->halo[s], m->dev[s], m->h[s]->stream, m->ev2[s]));
        OnStream on(m->h[s], m->halo[s]);
        MK(m, m->h[s], ek_halo_pack(m->h[s], phase, m->to_l[s], m->to_r[s]));
    }
    for (int s = 0; s < m->P; ++s) {
        DeviceGuard g(m->dev[s]);
        MCUDA(m, cudaEventRecord(m->ev2[s], m->halo[s]));
    }
    for (int s = 0; s < m->P; ++s) {
        const int l = (s + m->P - 1) % m->P, r = (s + 1) % m->P;
        DeviceGuard g(m->dev[s]);
        MCUDA(m, cudaStreamWaitEvent(m->halo[s], m->ev2[l], 0));
        MCUDA(m, cudaStreamWaitEvent(m->halo[s], m->ev2[r], 0));
        MCUDA(m, cudaMemcpyAsync(m->from_l[s], m->to_r[l], n * sizeof(double), cudaMemcpyDefault, m->halo[s]));
        MCUDA(m, cudaMemcpyAsync(m->from_r[s], m->to_l[r], n * sizeof(double), cudaMemcpyDefault, m->halo[s]));
        OnStream on(m->h[s], m->halo[s]);
        MK(m, m->h[s], ek_halo_unpack(m->h[s], phase, m->from_l[s], m->from_r[s]));
    }
    return EK_OK;
}

ek_status halo_join(ek_multi *m)
{
    for (int s = 0; s < m->P; ++s)
        MK(m, nullptr, wait_for(m, m->dev[s], m->h[s]->stream, m->dev[s], m->halo[s], m->ev2[s]));
    return EK_OK;
}

ek_status halo_exchange(ek_multi *m, int phase)
{
    for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_halo_pack(m->h[s], phase, m->to_l[s], m->to_r[s]));
    const size_t n = (size_t)ek_halo_doubles(m->h[0]);
    MK(m, nullptr, ring_exchange(m, m->to_l, m->to_r, m->from_l, m->from_r, 0, n));
    for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_halo_unpack(m->h[s], phase, m->from_l[s], m->from_r[s]));
    return EK_OK;
}

ek_status phi_halo_exchange(ek_multi *m)
{
    for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_phi_halo_pack(m->h[s], m->pto_l[s], m->pto_r[s]));
    const size_t n = (size_t)m->global.NY * m->global.NZ;
    MK(m, nullptr, ring_exchange(m, m->pto_l, m->pto_r, m->pfrom_l, m->pfrom_r, 0, n));
    for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_phi_halo_unpack(m->h[s], m->pfrom_l[s], m->pfrom_r[s]));
    return EK_OK;
}

ek_status poisson_rest(ek_multi *m, bool join_halo);

// the distributed fast_Poisson(): c+ - c- -> phi, ghost columns included
ek_status poisson(ek_multi *m)
{
    for (int k = 0; k < m->K; ++k)
        for (int s = 0; s < m->P; ++s) {
            MK(m, m->h[s], ek_slab_poisson_forward(m->h[s], k));
            MK(m, m->h[s], ek_slab_poisson_push_x(m->h[s], k));
        }
    return poisson_rest(m, false);
}

// One LBM pass launched chunk by chunk; chunk k's y-transform and its pushes into the peers' pencils
// run on the side streams behind the launches of the later chunks (as slab.py's lbm_and_forward)
ek_status lbm_and_forward(ek_multi *m, int full)
{
    for (int k = 0; k < m->K; ++k) {
        for (int s = 0; s < m->P; ++s) {
            const EkSlabPoisson &S = m->h[s]->sp;
            MK(m, m->h[s], ek_stream_collide_save_range(m->h[s], full, S.block0[k], S.block0[k + 1], k == m->K - 1));
        }
        for (int s = 0; s < m->P; ++s) {
            MK(m, nullptr, wait_for(m, m->dev[s], m->side[s], m->dev[s], m->h[s]->stream, m->ev[s]));
            OnStream on(m->h[s], m->side[s]);
            MK(m, m->h[s], ek_slab_poisson_forward(m->h[s], k));
            MK(m, m->h[s], ek_slab_poisson_push_x(m->h[s], k));
        }
    }
    return EK_OK;
}

ek_status poisson_rest(ek_multi *m, bool join_halo)
{
    if (m->pipeline)   // the main streams join their side streams before the barrier
        for (int s = 0; s < m->P; ++s)
            MK(m, nullptr, wait_for(m, m->dev[s], m->h[s]->stream, m->dev[s], m->side[s], m->ev[s]));
    MK(m, nullptr, barrier_all(m));   // every slab's rows have landed in everybody's pencils
    for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_slab_poisson_solve(m->h[s]));
    for (int k = 0; k < m->K; ++k)
        for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_slab_poisson_push_back(m->h[s], k));
    MK(m, nullptr, barrier_all(m));
    for (int k = 0; k < m->K; ++k)
        for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_slab_poisson_backward(m->h[s], k));
    for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_poisson_finish(m->h[s], 0));
    // the population halos have had the whole stage to travel: the main streams join them here, so that
    // the barrier of the phi exchange also orders every neighbour's copy before anybody's next pack
    if (join_halo) MK(m, nullptr, halo_join(m));
    return phi_halo_exchange(m);
}

// after a checkpoint was loaded into the natural layout (A-A parity 0): the next, even step is
// node-local and refreshes the population ghosts itself; only phi's ghost columns are needed
ek_status restore_ghosts(ek_multi *m) { return phi_halo_exchange(m); }

void release(ek_multi *m)
{
    for (int s = 0; s < (int)m->h.size(); ++s) {
        DeviceGuard g(m->dev[s]);
        if (s < (int)m->ev.size() && m->ev[s]) cudaEventDestroy(m->ev[s]);
        if (s < (int)m->ev2.size() && m->ev2[s]) cudaEventDestroy(m->ev2[s]);
        if (s < (int)m->side.size() && m->side[s]) cudaStreamDestroy(m->side[s]);
        if (s < (int)m->halo.size() && m->halo[s]) cudaStreamDestroy(m->halo[s]);
        auto fr = [&](std::vector<double *> &v) { if (s < (int)v.size()) cudaFree(v[s]); };
        fr(m->to_l); fr(m->to_r); fr(m->from_l); fr(m->from_r);
        fr(m->pto_l); fr(m->pto_r); fr(m->pfrom_l); fr(m->pfrom_r);
        if (m->h[s]) ek_destroy(m->h[s]);
    }
    delete m;
}

}  // namespace

extern "C" {

const char *ek_multi_last_error(ek_multi *m) { return m ? m->err.c_str() : ""; }

// `global` describes the whole domain; slab s (columns [s*NX/nslabs, (s+1)*NX/nslabs)) lives on
// CUDA device devices[s] (devices may repeat).  poisson_chunks <= 0: default (4).
ek_status ek_multi_create(const ek_params *global, int nslabs, const int *devices, int poisson_chunks, ek_multi **out)
{
    if (!global || !devices || !out || nslabs < 1 || nslabs > EK_MAX_RANKS) return EK_ERR_INVALID;
    *out = nullptr;
    ek_multi *m = new (std::nothrow) ek_multi();
    if (!m) return EK_ERR_NOMEM;
    m->global = *global;
    m->P = nslabs;
    // peer access between all pairs of distinct devices (one process: no IPC needed)
    for (int a = 0; a < nslabs; ++a)
        for (int b = 0; b < nslabs; ++b) {
            if (devices[a] == devices[b]) continue;
            DeviceGuard g(devices[a]);
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { cudaGetLastError(); release(m); return EK_ERR_CUDA; }
        }
    for (int s = 0; s < nslabs; ++s) {
        ek_handle *h = nullptr;
        ek_status st = ek_create_slab(global, devices[s], s, nslabs, &h);
        if (st != EK_OK) { release(m); return st; }
        m->h.push_back(h);
        m->dev.push_back(devices[s]);
    }
    ek_status st = EK_OK;
    for (int s = 0; s < nslabs && st == EK_OK; ++s) {
        DeviceGuard g(m->dev[s]);
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { st = EK_ERR_CUDA; break; }
        m->ev.push_back(e);
        cudaEvent_t e2 = nullptr;
        cudaStream_t s1 = nullptr, s2 = nullptr;
        if (cudaEventCreateWithFlags(&e2, cudaEventDisableTiming) != cudaSuccess ||
            cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking) != cudaSuccess) { st = EK_ERR_CUDA; break; }
        m->ev2.push_back(e2);
        m->side.push_back(s1);
        m->halo.push_back(s2);
        st = ek_slab_poisson_setup(m->h[s], poisson_chunks > 0 ? poisson_chunks : 4);
        if (st != EK_OK) break;
        const size_t nh = (size_t)ek_halo_doubles(m->h[s]) * sizeof(double);
        const size_t np = (size_t)global->NY * global->NZ * sizeof(double);
        auto al = [&](std::vector<double *> &v, size_t bytes) {
            double *p = nullptr;
            if (cudaMalloc((void **)&p, bytes) != cudaSuccess) st = EK_ERR_NOMEM;
            v.push_back(p);
        };
        al(m->to_l, nh); al(m->to_r, nh); al(m->from_l, nh); al(m->from_r, nh);
        al(m->pto_l, np); al(m->pto_r, np); al(m->pfrom_l, np); al(m->pfrom_r, np);
    }
    if (st != EK_OK) { release(m); return st; }
    m->K = ek_slab_poisson_chunks(m->h[0]);
    // everybody's pencil / receive buffers as peer pointers; pushes on the copy engines
    for (int s = 0; s < nslabs; ++s) {
        void *X = nullptr, *R = nullptr;
        ek_slab_poisson_my_buffers(m->h[s], &X, &R);
        for (int t = 0; t < nslabs; ++t) ek_slab_poisson_set_peer(m->h[t], s, X, R);
    }
    for (int s = 0; s < nslabs; ++s) ek_slab_poisson_set_dma(m->h[s], 1);
    *out = m;
    return EK_OK;
}

ek_status ek_multi_destroy(ek_multi *m)
{
    if (!m) return EK_ERR_INVALID;
    for (int s = 0; s < m->P; ++s) ek_sync(m->h[s]);
    release(m);
    return EK_OK;
}

// 1 (default): forward half of the Poisson stage behind the LBM launches, halos next to the Poisson
// stage; 0: everything in sequence on one stream per slab
ek_status ek_multi_set_pipeline(ek_multi *m, int on)
{
    if (!m) return EK_ERR_INVALID;
    m->pipeline = on != 0;
    return EK_OK;
}

int ek_multi_slabs(ek_multi *m) { return m ? m->P : 0; }
ek_handle *ek_multi_slab(ek_multi *m, int s) { return (m && s >= 0 && s < m->P) ? m->h[s] : nullptr; }

ek_status ek_multi_sync(ek_multi *m)
{
    if (!m) return EK_ERR_INVALID;
    for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_sync(m->h[s]));
    return EK_OK;
}

// initialization() of the reference (LBM.cu:68-146) on the decomposed domain
ek_status ek_multi_init_fields(ek_multi *m)
{
    if (!m) return EK_ERR_INVALID;
    for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_init_uniform(m->h[s]));
    for (int it = 0; it < m->global.pb_iters; ++it) {
        for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_pbe(m->h[s]));
        MK(m, nullptr, poisson(m));
        if (it == m->global.pb_iters - 1)   // E is taken from the un-relaxed phi of the last solve (LBM.cu:96-104)
            for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_compute_efield(m->h[s]));
        for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_pbe_relax(m->h[s]));
    }
    for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_mark_fields_ready(m->h[s]));
    m->pops = false;
    return EK_OK;
}

ek_status ek_multi_init_equilibrium(ek_multi *m)
{
    if (!m) return EK_ERR_INVALID;
    for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_init_equilibrium(m->h[s]));
    m->pops = true;
    return EK_OK;
}

ek_status ek_multi_init(ek_multi *m)
{
    ek_status st = ek_multi_init_fields(m);
    return st != EK_OK ? st : ek_multi_init_equilibrium(m);
}

// nsteps iterations of main.cu:189-200 on all slabs
ek_status ek_multi_step(ek_multi *m, int nsteps)
{
    if (!m || nsteps < 0) return EK_ERR_INVALID;
    if (!m->pops) { m->err = "ek_multi_step before ek_multi_init_equilibrium"; return EK_ERR_STATE; }
    for (int i = 0; i < nsteps; ++i) {
        const int full = (i == nsteps - 1);
        const int phase = ek_lbm_parity(m->h[0]) == 0 ? 0 : 1;
        if (m->pipeline && m->K > 1) {
            MK(m, nullptr, lbm_and_forward(m, full));
            MK(m, nullptr, halo_exchange_async(m, phase));   // next to the Poisson stage
            MK(m, nullptr, poisson_rest(m, true));
        } else {
            for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_stream_collide_save(m->h[s], full));
            MK(m, nullptr, halo_exchange(m, phase));
            MK(m, nullptr, poisson(m));
        }
        if (full)
            for (int s = 0; s < m->P; ++s) MK(m, m->h[s], ek_compute_efield(m->h[s]));
        for (int s = 0; s < m->P; ++s) m->h[s]->steps += 1;   // what ek_multi_checkpoint_save writes
    }
    return EK_OK;
}

ek_status ek_multi_step_timed(ek_multi *m, int nsteps, float *ms)
{
    if (!m || !ms) return EK_ERR_INVALID;
    MK(m, nullptr, ek_multi_sync(m));
    DeviceGuard g(m->dev[0]);
    cudaEvent_t e0, e1;
    MCUDA(m, cudaEventCreate(&e0));
    MCUDA(m, cudaEventCreate(&e1));
    MCUDA(m, cudaEventRecord(e0, m->h[0]->stream));
    ek_status st = ek_multi_step(m, nsteps);
    if (st == EK_OK) st = barrier_all(m);
    cudaError_t e = cudaEventRecord(e1, m->h[0]->stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (st != EK_OK) return st;
    MCUDA(m, e);
    return ek_multi_sync(m);
}

// global arrays in the reference's layout NXg*(NY*z+y)+x (host memory)
ek_status ek_multi_get_field(ek_multi *m, int id, double *host_global)
{
    if (!m || !host_global || id < 0 || id >= EK_NFIELDS) return EK_ERR_INVALID;
    const int NXg = m->global.NX, NXl = NXg / m->P, NY = m->global.NY, NZ = m->global.NZ;
    std::vector<double> buf((size_t)NXl * NY * NZ);
    for (int s = 0; s < m->P; ++s) {
        MK(m, m->h[s], ek_get_field(m->h[s], id, buf.data(), 0));
        for (size_t r = 0; r < (size_t)NY * NZ; ++r)
            memcpy(host_global + r * NXg + (size_t)s * NXl, buf.data() + r * NXl, (size_t)NXl * sizeof(double));
    }
    return EK_OK;
}

ek_status ek_multi_set_fields(ek_multi *m, const double *const host_global[EK_NFIELDS])
{
    if (!m || !host_global) return EK_ERR_INVALID;
    const int NXg = m->global.NX, NXl = NXg / m->P, NY = m->global.NY, NZ = m->global.NZ;
    std::vector<double> buf[EK_NFIELDS];
    for (int s = 0; s < m->P; ++s) {
        const double *ptr[EK_NFIELDS];
        for (int k = 0; k < EK_NFIELDS; ++k) {
            ptr[k] = nullptr;
            if (!host_global[k]) continue;
            buf[k].resize((size_t)NXl * NY * NZ);
            for (size_t r = 0; r < (size_t)NY * NZ; ++r)
                memcpy(buf[k].data() + r * NXl, host_global[k] + r * NXg + (size_t)s * NXl, (size_t)NXl * sizeof(double));
            ptr[k] = buf[k].data();
        }
        MK(m, m->h[s], ek_set_fields(m->h[s], ptr, 0));
    }
    m->pops = false;
    return EK_OK;
}

// ---- diagnostics and dumps of the whole domain (LBM.cu:2492-2753) ----------------------------
ek_status ek_multi_wall_current(ek_multi *m, double *current)
{
    if (!m || !current) return EK_ERR_INVALID;
    double I = 0.0;
    for (int s = 0; s < m->P; ++s) {
        double Is = 0.0;
        MK(m, m->h[s], ek_wall_current(m->h[s], &Is));
        I += Is;
    }
    *current = I;
    return EK_OK;
}

ek_status ek_multi_max_uz(ek_multi *m, double *umax)
{
    if (!m || !umax) return EK_ERR_INVALID;
    double v = 0.0;
    for (int s = 0; s < m->P; ++s) {
        double vs = 0.0;
        MK(m, m->h[s], ek_max_uz(m->h[s], &vs));
        v = vs > v ? vs : v;
    }
    *umax = v;
    return EK_OK;
}

static ek_status gather_all(ek_multi *m, EkHostFields &H)
{
    const size_t cells = (size_t)m->global.NX * m->global.NY * m->global.NZ;
    for (int k = 0; k < EK_NFIELDS; ++k) {
        H.f[k].resize(cells);
        ek_status st = ek_multi_get_field(m, k, H.f[k].data());
        if (st != EK_OK) return st;
    }
    ek_io_extrapolate_walls(H, m->global.NX, m->global.NY, m->global.NZ);
    return EK_OK;
}

ek_status ek_multi_save_data_tecplot(ek_multi *m, const char *path, double time, int append, int first)
{
    if (!m || !path) return EK_ERR_INVALID;
    EkHostFields H;
    ek_status st = gather_all(m, H);
    if (st != EK_OK) return st;
    const EkDumpGrid g = {m->global.NX, m->global.NY, m->global.NZ, m->global.dx, m->global.dy, m->global.dz};
    if (!ek_io_write_tecplot(path, g, H, time, append, first)) { m->err = std::string("cannot open ") + path; return EK_ERR_INVALID; }
    return EK_OK;
}

ek_status ek_multi_save_data_end(ek_multi *m, const char *path, double time)
{
    if (!m || !path) return EK_ERR_INVALID;
    EkHostFields H;
    ek_status st = gather_all(m, H);
    if (st != EK_OK) return st;
    const EkDumpGrid g = {m->global.NX, m->global.NY, m->global.NZ, m->global.dx, m->global.dy, m->global.dz};
    if (!ek_io_write_end(path, g, H, time)) { m->err = std::string("cannot open ") + path; return EK_ERR_INVALID; }
    return EK_OK;
}

// ---- restart of the whole domain: the reference's text file and the exact checkpoint, in the
// formats of ek_read_data / ek_checkpoint_* (a single-GPU file restarts a multi-GPU run and back)
ek_status ek_multi_read_data(ek_multi *m, const char *path, double *time)
{
    if (!m || !path) return EK_ERR_INVALID;
    EkHostFields H;
    if (!ek_io_read_end(path, (size_t)m->global.NX * m->global.NY * m->global.NZ, H, time, m->err)) return EK_ERR_INVALID;
    const double *ptr[EK_NFIELDS];
    for (int k = 0; k < EK_NFIELDS; ++k) ptr[k] = H.f[k].data();
    return ek_multi_set_fields(m, ptr);
}

ek_status ek_multi_checkpoint_save(ek_multi *m, const char *path, double time)
{
    if (!m || !path) return EK_ERR_INVALID;
    if (!m->pops) { m->err = "ek_multi_checkpoint_save before the populations exist"; return EK_ERR_STATE; }
    FILE *f = fopen(path, "wb");
    if (!f) { m->err = std::string("cannot open ") + path; return EK_ERR_INVALID; }
    const int NXg = m->global.NX, NXl = NXg / m->P, NY = m->global.NY, NZ = m->global.NZ;
    const size_t cells = (size_t)NXg * NY * NZ, lcells = (size_t)NXl * NY * NZ, rows = (size_t)NY * NZ;
    EkCkptHeader hd;
    memcpy(hd.magic, "EKB200C1", 8);
    hd.NX = NXg; hd.NY = NY; hd.NZ = NZ; hd.nfields = EK_NFIELDS;
    hd.steps = m->h[0]->steps; hd.time = time;
    bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1;
    ek_status st = EK_OK;
    std::vector<double> g(cells), l(27 * lcells);
    for (int k = 0; k < EK_NFIELDS && ok && st == EK_OK; ++k) {
        st = ek_multi_get_field(m, k, g.data());
        ok = st == EK_OK && fwrite(g.data(), sizeof(double), cells, f) == cells;
    }
    // populations: one direction of the whole domain at a time would need 27 passes over the slabs;
    // instead one set at a time (27 x cells doubles on the host)
    std::vector<double> gp;
    for (int set = 0; set < EK_NSETS && ok && st == EK_OK; ++set) {
        gp.resize(27 * cells);
        for (int s = 0; s < m->P && st == EK_OK; ++s) {
            st = ek_get_populations(m->h[s], set, l.data(), 0);
            if (st != EK_OK) { fail(m, m->h[s], "ek_get_populations", st); break; }
            for (int d = 0; d < 27; ++d)
                for (size_t r = 0; r < rows; ++r)
                    memcpy(gp.data() + (size_t)d * cells + r * NXg + (size_t)s * NXl, l.data() + (size_t)d * lcells + r * NXl,
                           (size_t)NXl * sizeof(double));
        }
        ok = st == EK_OK && fwrite(gp.data(), sizeof(double), 27 * cells, f) == 27 * cells;
    }
    fclose(f);
    if (st != EK_OK) return st;
    if (!ok) { m->err = std::string("short write to ") + path; return EK_ERR_INVALID; }
    return EK_OK;
}

ek_status ek_multi_checkpoint_load(ek_multi *m, const char *path, double *time)
{
    if (!m || !path) return EK_ERR_INVALID;
    FILE *f = fopen(path, "rb");
    if (!f) { m->err = std::string("cannot open ") + path; return EK_ERR_INVALID; }
    const int NXg = m->global.NX, NXl = NXg / m->P, NY = m->global.NY, NZ = m->global.NZ;
    const size_t cells = (size_t)NXg * NY * NZ, lcells = (size_t)NXl * NY * NZ, rows = (size_t)NY * NZ;
    EkCkptHeader hd;
    if (fread(&hd, sizeof(hd), 1, f) != 1 || memcmp(hd.magic, "EKB200C1", 8) != 0 || hd.NX != NXg || hd.NY != NY ||
        hd.NZ != NZ || hd.nfields != EK_NFIELDS) {
        fclose(f);
        m->err = std::string(path) + ": not a checkpoint of this grid";
        return EK_ERR_INVALID;
    }
    EkHostFields H;
    bool ok = true;
    for (int k = 0; k < EK_NFIELDS && ok; ++k) {
        H.f[k].resize(cells);
        ok = fread(H.f[k].data(), sizeof(double), cells, f) == cells;
    }
    ek_status st = EK_OK;
    if (ok) {
        const double *ptr[EK_NFIELDS];
        for (int k = 0; k < EK_NFIELDS; ++k) ptr[k] = H.f[k].data();
        st = ek_multi_set_fields(m, ptr);
    }
    std::vector<double> gp(27 * cells), l(27 * lcells);
    for (int set = 0; set < EK_NSETS && ok && st == EK_OK; ++set) {
        ok = fread(gp.data(), sizeof(double), 27 * cells, f) == 27 * cells;
        for (int s = 0; s < m->P && ok && st == EK_OK; ++s) {
            for (int d = 0; d < 27; ++d)
                for (size_t r = 0; r < rows; ++r)
                    memcpy(l.data() + (size_t)d * lcells + r * NXl, gp.data() + (size_t)d * cells + r * NXg + (size_t)s * NXl,
                           (size_t)NXl * sizeof(double));
            st = ek_set_populations(m->h[s], set, l.data());
            if (st != EK_OK) fail(m, m->h[s], "ek_set_populations", st);
        }
    }
    fclose(f);
    if (st != EK_OK) return st;
    if (!ok) { m->err = std::string(path) + ": truncated checkpoint"; return EK_ERR_INVALID; }
    for (int s = 0; s < m->P; ++s) {
        MK(m, m->h[s], ek_populations_restored(m->h[s]));
        m->h[s]->steps = hd.steps;
    }
    // the ghost columns are not part of the file: the face populations of the natural layout and phi
    MK(m, nullptr, restore_ghosts(m));
    m->pops = true;
    if (time) *time = hd.time;
    return EK_OK;
}

}  // extern "C"
