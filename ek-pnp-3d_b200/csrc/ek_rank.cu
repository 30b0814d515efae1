// ek_rank.cu -- native driver of the x-slab path with ONE PROCESS PER GPU (the torchrun / mpirun
// layout): this rank's slab handle plus everything that has to happen between the ranks, in C++:
//
//   * population halos: pack kernel -> ncclSend/ncclRecv to the two ring neighbours in one
//     ncclGroup -> unpack kernel, on their own stream next to the Poisson stage;
//   * Poisson transposes: grouped ncclSend/ncclRecv with all P ranks (an all-to-all) per z-chunk;
//     chunk k's forward half (y-transform + transpose 1) runs on a side stream behind the LBM launches
//     of the chunks k+1.., the transposes back travel while the previous chunk is transformed;
//   * phi halos per chunk on the way back; the next step's LBM launches only wait for the chunks whose
//     planes they read.
//
// No Python between the launches of a step: the host enqueues ~60 asynchronous operations per coupled
// step and runs ahead of the device.  The process launcher only has to hand every rank the same NCCL
// unique ids (ek_rank_nccl_unique_id on rank 0, then broadcast by whatever the launcher offers:
// torch.distributed in bench.py, MPI_Bcast or a shared file elsewhere).
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"): the single-GPU product path has no NCCL
// dependency, and under torchrun the copy that PyTorch already loaded is the one used.
// The reference is single-GPU (main.cu:58); the per-step call sequence replaces main.cu:189-200.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "ek_handle.h"

namespace {

struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};

NcclApi g_nccl;

bool load_nccl(std::string &err)
{
    if (g_nccl.lib) return true;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define EK_SYM(name)                                                                     \
    g_nccl.name = (decltype(g_nccl.name))dlsym(lib, "nccl" #name);                       \
    if (!g_nccl.name) { err = "libnccl: symbol nccl" #name " missing"; return false; }
    EK_SYM(GetUniqueId) EK_SYM(CommInitRank) EK_SYM(CommDestroy) EK_SYM(GroupStart) EK_SYM(GroupEnd)
    EK_SYM(Send) EK_SYM(Recv) EK_SYM(GetErrorString) EK_SYM(GetVersion)
#undef EK_SYM
    g_nccl.lib = lib;
    return true;
}

enum { COMM_X = 0, COMM_H = 1, COMM_P = 2, NCOMM = 3 };   // transposes, population halos, phi halos

}  // namespace

struct ek_rank {
    ek_handle *h = nullptr;
    int rank = 0, P = 1, device = 0, K = 1;
    ncclComm_t comm[NCOMM] = {nullptr, nullptr, nullptr};
    cudaStream_t side = nullptr, halo = nullptr, copy = nullptr, back = nullptr;
    std::vector<cudaEvent_t> ev_lbm, ev_sc, ev_landed, ev_phi;
    cudaEvent_t ev_side = nullptr, ev_main = nullptr, ev_halo = nullptr, ev_back = nullptr, ev_bnd = nullptr;
    bool have_phi_ready = false;
    bool overlap = true, overlap_back = true;
    bool ghosts = true;           // phi ghost columns inside transpose 2 (ek_slab_poisson_enable_ghosts)
    // LBM pass: boundary x-tiles first, population halos under the interior launches.  Measured at 2 x 134 M cells:
    // the split pass costs 1.4 ms more than it hides (halos are hidden behind the Poisson stage anyway): off by default
    bool boundary_first = false;
    double *to_l = nullptr, *to_r = nullptr, *from_l = nullptr, *from_r = nullptr;      // populations
    double *pto_l = nullptr, *pto_r = nullptr, *pfrom_l = nullptr, *pfrom_r = nullptr;  // phi
    bool pops = false;
    long long nccl_groups = 0;
    // phase profile (ek_rank_profile): everything in sequence on the main stream, timed events between the phases
    bool profile = false;
    std::vector<std::pair<const char *, cudaEvent_t>> marks;
    std::string err;
};

namespace {

#define RK(r, call)                                                         \
    do {                                                                    \
        ek_status _s = (call);                                              \
        if (_s != EK_OK) {                                                  \
            (r)->err = std::string(#call) + ": " + ek_last_error((r)->h);   \
            return _s;                                                      \
        }                                                                   \
    } while (0)

#define RCUDA(r, call)                                                      \
    do {                                                                    \
        cudaError_t _e = (call);                                            \
        if (_e != cudaSuccess) {                                            \
            (r)->err = std::string(#call) + ": " + cudaGetErrorString(_e);  \
            return EK_ERR_CUDA;                                             \
        }                                                                   \
    } while (0)

#define RNCCL(r, call)                                                               \
    do {                                                                             \
        ncclResult_t _n = (call);                                                    \
        if (_n != ncclSuccess) {                                                     \
            (r)->err = std::string(#call) + ": NCCL: " + g_nccl.GetErrorString(_n);  \
            return EK_ERR_CUDA;                                                      \
        }                                                                            \
    } while (0)

// the calls of the slab's C ABI made in this scope run on another stream of the device
struct OnStream {
    ek_handle *h;
    cudaStream_t saved;
    OnStream(ek_handle *hh, cudaStream_t st) : h(hh), saved(hh->stream) { h->stream = st; }
    ~OnStream() { h->stream = saved; }
};

// stream `a` waits for everything issued so far on stream `b`
ek_status wait_for(ek_rank *r, cudaStream_t a, cudaStream_t b, cudaEvent_t ev)
{
    RCUDA(r, cudaEventRecord(ev, b));
    RCUDA(r, cudaStreamWaitEvent(a, ev, 0));
    return EK_OK;
}

// profile mode: a timed event on the main stream closes the phase `name`
ek_status mark(ek_rank *r, const char *name)
{
    if (!r->profile) return EK_OK;
    cudaEvent_t e;
    RCUDA(r, cudaEventCreate(&e));
    RCUDA(r, cudaEventRecord(e, r->h->stream));
    r->marks.emplace_back(name, e);
    return EK_OK;
}

// ring exchange: my to_right goes to the right neighbour's from_left, my to_left to the left neighbour's from_right
ek_status ring_exchange(ek_rank *r, int which, const double *to_l, const double *to_r, double *from_l, double *from_r,
                        size_t count, cudaStream_t st)
{
    const int left = (r->rank + r->P - 1) % r->P, right = (r->rank + 1) % r->P;
    if (r->P == 1) {
        RCUDA(r, cudaMemcpyAsync(from_l, to_r, count * sizeof(double), cudaMemcpyDeviceToDevice, st));
        RCUDA(r, cudaMemcpyAsync(from_r, to_l, count * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return EK_OK;
    }
    // order matters when left == right (P = 2): the peer's first send (its to_right) is my from_left
    RNCCL(r, g_nccl.GroupStart());
    RNCCL(r, g_nccl.Send(to_r, count, ncclDouble, right, r->comm[which], st));
    RNCCL(r, g_nccl.Send(to_l, count, ncclDouble, left, r->comm[which], st));
    RNCCL(r, g_nccl.Recv(from_l, count, ncclDouble, left, r->comm[which], st));
    RNCCL(r, g_nccl.Recv(from_r, count, ncclDouble, right, r->comm[which], st));
    RNCCL(r, g_nccl.GroupEnd());
    r->nccl_groups += 1;
    return EK_OK;
}

// all-to-all of the transpose buffers of chunk k: part i of `send` travels to rank i
ek_status all_to_all(ek_rank *r, int k, cudaStream_t st, bool back = false)
{
    void *send = nullptr, *recv = nullptr;
    long long count = 0;
    if (back && r->ghosts) RK(r, ek_slab_poisson_chunk_back(r->h, k, &send, &recv, &count));
    else RK(r, ek_slab_poisson_chunk(r->h, k, nullptr, nullptr, &send, &recv, &count));
    if (count == 0) return EK_OK;
    const size_t n = (size_t)(count / r->P) * 2;   // doubles per part (complex)
    if (r->P == 1) {
        RCUDA(r, cudaMemcpyAsync(recv, send, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return EK_OK;
    }
    RNCCL(r, g_nccl.GroupStart());
    for (int j = 0; j < r->P; ++j) {
        const int i = (r->rank + j) % r->P;
        RNCCL(r, g_nccl.Send((const double *)send + (size_t)i * n, n, ncclDouble, i, r->comm[COMM_X], st));
        RNCCL(r, g_nccl.Recv((double *)recv + (size_t)i * n, n, ncclDouble, i, r->comm[COMM_X], st));
    }
    RNCCL(r, g_nccl.GroupEnd());
    r->nccl_groups += 1;
    return EK_OK;
}

void chunk_planes(ek_rank *r, int k, int *z0, int *z1)
{
    const EkSlabPoisson &S = r->h->sp;
    const int zc = r->h->zchunk, NZ = r->h->c.NZ;
    *z0 = S.block0[k] * zc;
    *z1 = S.block0[k + 1] * zc < NZ ? S.block0[k + 1] * zc : NZ;
}

ek_status halo_start(ek_rank *r, int phase, cudaEvent_t after)
{
    ek_handle *h = r->h;
    RCUDA(r, cudaStreamWaitEvent(r->halo, after, 0));   // after the launches that touch the boundary / ghost columns
    OnStream on(h, r->halo);
    RK(r, ek_halo_pack(h, phase, r->to_l, r->to_r));
    RK(r, ring_exchange(r, COMM_H, r->to_l, r->to_r, r->from_l, r->from_r, (size_t)ek_halo_doubles(h), r->halo));
    RK(r, ek_halo_unpack(h, phase, r->from_l, r->from_r));
    RCUDA(r, cudaEventRecord(r->ev_halo, r->halo));
    return EK_OK;
}

// way back, chunk by chunk on the `back` stream: wait until chunk k has landed, inverse y-transform into
// phi, ghost columns of its planes, then the event that the next LBM launches wait for
ek_status poisson_tail(ek_rank *r)
{
    ek_handle *h = r->h;
    RK(r, wait_for(r, r->back, h->stream, r->ev_back));
    OnStream on(h, r->back);
    RK(r, ek_poisson_finish(h, 0));   // wall planes first (only when something other than the solver wrote phi)
    for (int k = 0; k < r->K; ++k) {
        RCUDA(r, cudaStreamWaitEvent(r->back, r->ev_landed[k], 0));
        RK(r, r->ghosts ? ek_slab_poisson_backward_g(h, k) : ek_slab_poisson_backward(h, k));
        int z0, z1;
        chunk_planes(r, k, &z0, &z1);
        if (z1 > z0 && !r->ghosts) {
            RK(r, ek_phi_halo_pack_range(h, z0, z1, r->pto_l, r->pto_r));
            const size_t a = (size_t)z0 * h->c.NY, n = (size_t)(z1 - z0) * h->c.NY;
            RK(r, ring_exchange(r, COMM_P, r->pto_l + a, r->pto_r + a, r->pfrom_l + a, r->pfrom_r + a, n, r->back));
            RK(r, ek_phi_halo_unpack_range(h, z0, z1, r->pfrom_l, r->pfrom_r));
        }
        RCUDA(r, cudaEventRecord(r->ev_phi[k], r->back));
    }
    r->have_phi_ready = true;
    return EK_OK;
}

ek_status join_back(ek_rank *r)
{
    RK(r, wait_for(r, r->h->stream, r->back, r->ev_back));
    r->have_phi_ready = false;
    return EK_OK;
}

// everything after the forward halves (y-transform + transpose 1 of every chunk) were enqueued on `fwd`
ek_status poisson_rest(ek_rank *r, cudaStream_t fwd)
{
    ek_handle *h = r->h;
    if (fwd != h->stream) RK(r, wait_for(r, h->stream, fwd, r->ev_side));
    RK(r, mark(r, "wait_for_forward_half"));
    RK(r, ek_slab_poisson_solve(h));
    RK(r, mark(r, "x_fft_zsolve_x_ifft"));
    for (int k = 0; k < r->K; ++k) {
        RK(r, r->ghosts ? ek_slab_poisson_scatter_xg(h, k) : ek_slab_poisson_scatter_x(h, k));
        RCUDA(r, cudaEventRecord(r->ev_sc[k], h->stream));
        RCUDA(r, cudaStreamWaitEvent(r->copy, r->ev_sc[k], 0));
        RK(r, all_to_all(r, k, r->copy, true));    // chunk k travels while chunk k+1 is re-blocked
        RCUDA(r, cudaEventRecord(r->ev_landed[k], r->copy));
    }
    RK(r, mark(r, "scatter_x"));
    return poisson_tail(r);
}

// forward half of chunk k on the handle's current stream: y-transform of my columns, transpose 1,
// received ky blocks into my full-x pencils
ek_status forward_chunk(ek_rank *r, int k)
{
    ek_handle *h = r->h;
    RK(r, ek_slab_poisson_forward(h, k));
    RK(r, all_to_all(r, k, h->stream));
    RK(r, ek_slab_poisson_gather_x(h, k));
    return EK_OK;
}

// the distributed fast_Poisson() in sequence (start-up loop): c+ - c- -> phi, ghost columns included
ek_status poisson(ek_rank *r)
{
    for (int k = 0; k < r->K; ++k) RK(r, forward_chunk(r, k));
    RK(r, poisson_rest(r, r->h->stream));
    return join_back(r);
}

// One LBM pass launched chunk by chunk; chunk k's forward half runs on the side stream behind the
// launches of the later chunks.  With boundary_first the two boundary x-tiles of every row are launched
// before the interior ones and the population halos (pack -> NCCL -> unpack, which only touch the boundary
// and ghost columns) travel under the interior launches.
ek_status lbm_and_forward(ek_rank *r, int full, int phase)
{
    ek_handle *h = r->h;
    const EkSlabPoisson &S = h->sp;
    const bool split = r->boundary_first && (h->c.NX + 31) / 32 >= 3 && (h->kernel == 0 || h->kernel == 3);
    // the planes of chunk k take grad(phi) from the chunks k-1 .. k+1 of the previous solve, whose way
    // back may still be running on the back stream
    auto wait_phi = [&](int k) -> ek_status {
        if (r->have_phi_ready) RCUDA(r, cudaStreamWaitEvent(h->stream, r->ev_phi[k + 1 < r->K ? k + 1 : r->K - 1], 0));
        return EK_OK;
    };
    if (split) {
        for (int k = 0; k < r->K; ++k) {
            RK(r, wait_phi(k));
            RK(r, ek_stream_collide_save_part(h, full, S.block0[k], S.block0[k + 1], 1, 0));
        }
        RCUDA(r, cudaEventRecord(r->ev_bnd, h->stream));
        RK(r, halo_start(r, phase, r->ev_bnd));
    }
    for (int k = 0; k < r->K; ++k) {
        if (!split) RK(r, wait_phi(k));
        RK(r, ek_stream_collide_save_part(h, full, S.block0[k], S.block0[k + 1], split ? 2 : 0, k == r->K - 1));
        RCUDA(r, cudaEventRecord(r->ev_lbm[k], h->stream));
        if (r->overlap) {
            RCUDA(r, cudaStreamWaitEvent(r->side, r->ev_lbm[k], 0));
            OnStream on(h, r->side);
            RK(r, forward_chunk(r, k));
        }
    }
    if (!split) RK(r, halo_start(r, phase, r->ev_lbm[r->K - 1]));
    if (!r->overlap) {
        RK(r, mark(r, "lbm"));
        for (int k = 0; k < r->K; ++k) RK(r, forward_chunk(r, k));
    }
    return EK_OK;
}

void release(ek_rank *r)
{
    if (!r) return;
    DeviceGuard g(r->device);
    if (r->h) ek_sync(r->h);
    for (cudaStream_t s : {r->side, r->halo, r->copy, r->back})
        if (s) { cudaStreamSynchronize(s); }
    for (int i = 0; i < NCOMM; ++i)
        if (r->comm[i]) g_nccl.CommDestroy(r->comm[i]);
    for (cudaStream_t s : {r->side, r->halo, r->copy, r->back})
        if (s) cudaStreamDestroy(s);
    for (auto *v : {&r->ev_lbm, &r->ev_sc, &r->ev_landed, &r->ev_phi})
        for (cudaEvent_t e : *v) cudaEventDestroy(e);
    for (cudaEvent_t e : {r->ev_side, r->ev_main, r->ev_halo, r->ev_back, r->ev_bnd})
        if (e) cudaEventDestroy(e);
    for (auto &m : r->marks) cudaEventDestroy(m.second);
    for (double *p : {r->to_l, r->to_r, r->from_l, r->from_r, r->pto_l, r->pto_r, r->pfrom_l, r->pfrom_r}) cudaFree(p);
    if (r->h) ek_destroy(r->h);
    delete r;
}

}  // namespace

extern "C" {

const char *ek_rank_last_error(ek_rank *r) { return r ? r->err.c_str() : "null rank handle"; }

// bytes of the opaque id block that rank 0 creates and every rank passes to ek_rank_create
int ek_rank_nccl_id_bytes(void) { return (int)(NCOMM * sizeof(ncclUniqueId)); }

ek_status ek_rank_nccl_unique_id(void *id)
{
    std::string err;
    if (!id) return EK_ERR_INVALID;
    if (!load_nccl(err)) { fprintf(stderr, "ek_b200: %s\n", err.c_str()); return EK_ERR_STATE; }
    for (int i = 0; i < NCOMM; ++i)
        if (g_nccl.GetUniqueId((ncclUniqueId *)id + i) != ncclSuccess) return EK_ERR_CUDA;
    return EK_OK;
}

int ek_rank_nccl_version(void)
{
    std::string err;
    int v = 0;
    if (!load_nccl(err) || g_nccl.GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

// `global` describes the whole domain; this process owns the x-slab `rank` of `nranks` on CUDA device `device`.
// nccl_id: ek_rank_nccl_id_bytes() bytes from rank 0's ek_rank_nccl_unique_id() (may be NULL when nranks == 1).
// Collective: every rank must call it.
ek_status ek_rank_create(const ek_params *global, int device, int rank, int nranks, const void *nccl_id,
                         int poisson_chunks, ek_rank **out)
{
    if (!global || !out || nranks < 1 || rank < 0 || rank >= nranks || nranks > EK_MAX_RANKS) return EK_ERR_INVALID;
    if (nranks > 1 && !nccl_id) return EK_ERR_INVALID;
    *out = nullptr;
    ek_rank *r = new (std::nothrow) ek_rank();
    if (!r) return EK_ERR_NOMEM;
    r->rank = rank; r->P = nranks; r->device = device;
    ek_status st = ek_create_slab(global, device, rank, nranks, &r->h);
    if (st != EK_OK) { delete r; return st; }
    r->device = r->h->device;
    DeviceGuard g(r->device);
    auto fail = [&](ek_status s, const std::string &msg) {
        fprintf(stderr, "ek_rank_create (rank %d): %s\n", rank, msg.c_str());
        release(r);
        return s;
    };
    if (nranks > 1) {
        std::string err;
        if (!load_nccl(err)) return fail(EK_ERR_STATE, err);
        for (int i = 0; i < NCOMM; ++i) {
            ncclUniqueId id;
            memcpy(&id, (const ncclUniqueId *)nccl_id + i, sizeof(id));
            ncclResult_t n = g_nccl.CommInitRank(&r->comm[i], nranks, id, rank);
            if (n != ncclSuccess) return fail(EK_ERR_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(n));
        }
    }
    st = ek_slab_poisson_setup(r->h, poisson_chunks);   // <= 0: automatic chunk sizes
    if (st != EK_OK) return fail(st, ek_last_error(r->h));
    r->K = ek_slab_poisson_chunks(r->h);
    if (const char *gh = getenv("EK_RANK_GHOSTS")) r->ghosts = atoi(gh) != 0;
    if (r->ghosts) {
        st = ek_slab_poisson_enable_ghosts(r->h);
        if (st != EK_OK) return fail(st, ek_last_error(r->h));
    }
    bool ok = true;
    // The side streams run at the HIGHEST priority: an LBM launch keeps every SM full for milliseconds, and at
    // equal priority the transforms / NCCL kernels queued next to it only get thread-block slots when the LBM
    // grid drains, i.e. the "overlapped" forward half piles up behind the last LBM chunk (measured: 2.2 ms of
    // waiting per step at 134 M cells per GPU).  With priority their blocks are scheduled as slots free up,
    // and the NVLink transfers really travel under the LBM kernel.
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    const char *prio_env = getenv("EK_RANK_PRIORITY");
    const int prio = (prio_env && !atoi(prio_env)) ? least : greatest;
    for (cudaStream_t *s : {&r->side, &r->halo, &r->copy, &r->back})
        ok = ok && cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, prio) == cudaSuccess;
    auto mkev = [&](cudaEvent_t *e) { ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess; };
    for (auto *v : {&r->ev_lbm, &r->ev_sc, &r->ev_landed, &r->ev_phi}) {
        v->assign(r->K, nullptr);
        for (int k = 0; k < r->K; ++k) mkev(&(*v)[k]);
    }
    for (cudaEvent_t *e : {&r->ev_side, &r->ev_main, &r->ev_halo, &r->ev_back, &r->ev_bnd}) mkev(e);
    const char *bf = getenv("EK_RANK_BOUNDARY_FIRST");
    if (bf) r->boundary_first = atoi(bf) != 0;
    const size_t nh = (size_t)ek_halo_doubles(r->h) * sizeof(double);
    const size_t np = (size_t)global->NY * global->NZ * sizeof(double);
    for (double **p : {&r->to_l, &r->to_r, &r->from_l, &r->from_r}) ok = ok && cudaMalloc((void **)p, nh) == cudaSuccess;
    for (double **p : {&r->pto_l, &r->pto_r, &r->pfrom_l, &r->pfrom_r}) ok = ok && cudaMalloc((void **)p, np) == cudaSuccess;
    if (!ok) return fail(EK_ERR_CUDA, std::string("streams / events / halo buffers: ") + cudaGetErrorString(cudaGetLastError()));
    *out = r;
    return EK_OK;
}

ek_status ek_rank_destroy(ek_rank *r)
{
    if (!r) return EK_ERR_INVALID;
    release(r);
    return EK_OK;
}

ek_handle *ek_rank_slab(ek_rank *r) { return r ? r->h : nullptr; }
int ek_rank_chunks(ek_rank *r) { return r ? r->K : 0; }

// 1 (default): forward half of the Poisson stage behind the LBM launches; back: its way back behind the
// next step's first LBM launches
ek_status ek_rank_set_pipeline(ek_rank *r, int overlap, int overlap_back)
{
    if (!r) return EK_ERR_INVALID;
    r->overlap = overlap != 0;
    r->overlap_back = overlap_back != 0;
    return EK_OK;
}

ek_status ek_rank_set_boundary_first(ek_rank *r, int on)
{
    if (!r) return EK_ERR_INVALID;
    r->boundary_first = on != 0;
    return EK_OK;
}

ek_status ek_rank_sync(ek_rank *r)
{
    if (!r) return EK_ERR_INVALID;
    DeviceGuard g(r->device);
    for (cudaStream_t s : {r->h->stream, r->side, r->halo, r->copy, r->back}) RCUDA(r, cudaStreamSynchronize(s));
    return EK_OK;
}

// initialization() of the reference (LBM.cu:68-146) on the decomposed domain.  Collective.
ek_status ek_rank_init_fields(ek_rank *r)
{
    if (!r) return EK_ERR_INVALID;
    DeviceGuard g(r->device);
    ek_handle *h = r->h;
    RK(r, ek_init_uniform(h));
    const int iters = h->p.pb_iters;
    for (int it = 0; it < iters; ++it) {
        RK(r, ek_pbe(h));
        RK(r, poisson(r));
        if (it == iters - 1) RK(r, ek_compute_efield(h));   // E of the un-relaxed phi of the last solve (LBM.cu:96-104)
        RK(r, ek_pbe_relax(h));
    }
    RK(r, ek_mark_fields_ready(h));
    r->pops = false;
    return EK_OK;
}

// after the caller wrote the slab's macroscopic arrays (ek_set_fields on ek_rank_slab()) or after
// ek_rank_init_fields: init_equilibrium() of the reference (LBM.cu:150-463)
ek_status ek_rank_init_equilibrium(ek_rank *r)
{
    if (!r) return EK_ERR_INVALID;
    DeviceGuard g(r->device);
    RK(r, ek_init_equilibrium(r->h));
    r->pops = true;
    r->have_phi_ready = false;
    return EK_OK;
}

ek_status ek_rank_init(ek_rank *r)
{
    ek_status st = ek_rank_init_fields(r);
    return st != EK_OK ? st : ek_rank_init_equilibrium(r);
}

// nsteps iterations of main.cu:189-200 on this rank's slab.  Collective; asynchronous.
ek_status ek_rank_step(ek_rank *r, int nsteps)
{
    if (!r || nsteps < 0) return EK_ERR_INVALID;
    if (!r->pops) { r->err = "ek_rank_step before ek_rank_init_equilibrium"; return EK_ERR_STATE; }
    DeviceGuard g(r->device);
    ek_handle *h = r->h;
    for (int i = 0; i < nsteps; ++i) {
        const int full = (i == nsteps - 1);
        const int phase = ek_lbm_parity(h) == 0 ? 0 : 1;
        RK(r, mark(r, "begin"));
        // (the population halos are started in there: under the interior launches, or after the pass)
        RK(r, lbm_and_forward(r, full, phase));
        RK(r, mark(r, r->overlap ? "lbm_launches" : "y_fft_transpose_1_gather_x"));
        if (r->profile && !r->overlap) {   // sequential profile: the halo exchange as its own phase
            RCUDA(r, cudaStreamWaitEvent(h->stream, r->ev_halo, 0));
            RK(r, mark(r, "population_halos"));
        }
        RK(r, poisson_rest(r, r->overlap ? r->side : h->stream));
        if (full || !r->overlap_back) RK(r, join_back(r));
        // else: the way back of the last chunks runs behind the next step's first LBM launches
        RK(r, mark(r, "transpose_2_y_ifft_phi_halos"));
        RCUDA(r, cudaStreamWaitEvent(h->stream, r->ev_halo, 0));
        RK(r, mark(r, "wait_for_population_halos"));
        if (full) RK(r, ek_compute_efield(h));
        h->steps += 1;
    }
    return EK_OK;
}

// ek_rank_step bracketed by CUDA events on this rank's main stream; blocks until this rank is done and
// returns its device time.  The caller provides the barrier before and the max over the ranks after.
ek_status ek_rank_step_timed(ek_rank *r, int nsteps, float *ms)
{
    if (!r || !ms) return EK_ERR_INVALID;
    DeviceGuard g(r->device);
    RK(r, ek_rank_sync(r));
    cudaEvent_t e0, e1;
    RCUDA(r, cudaEventCreate(&e0));
    RCUDA(r, cudaEventCreate(&e1));
    RCUDA(r, cudaEventRecord(e0, r->h->stream));
    ek_status st = ek_rank_step(r, nsteps);
    cudaError_t e = cudaEventRecord(e1, r->h->stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (st != EK_OK) return st;
    RCUDA(r, e);
    return ek_rank_sync(r);
}

// Phase split of nsteps steps as JSON {"phase": ms per step, ...} (development / DESIGN.md evidence).
// sequential != 0: no stream overlap at all, so that every phase shows its own cost; 0: the production
// pipeline, where a phase is what the MAIN stream waits for (overlapped work shows up as waits).
ek_status ek_rank_profile(ek_rank *r, int nsteps, int sequential, char *json, int cap)
{
    if (!r || !json || cap < 64 || nsteps < 1) return EK_ERR_INVALID;
    DeviceGuard g(r->device);
    const bool ov = r->overlap, ovb = r->overlap_back;
    if (sequential) { r->overlap = false; r->overlap_back = false; }
    RK(r, ek_rank_sync(r));
    r->profile = true;
    ek_status st = ek_rank_step(r, nsteps);
    r->profile = false;
    r->overlap = ov; r->overlap_back = ovb;
    auto drop_marks = [&]() {
        for (auto &m : r->marks) cudaEventDestroy(m.second);
        r->marks.clear();
    };
    if (st == EK_OK) st = ek_rank_sync(r);
    if (st != EK_OK) { drop_marks(); return st; }
    std::vector<std::pair<std::string, double>> acc;
    for (size_t i = 1; i < r->marks.size(); ++i) {
        if (!strcmp(r->marks[i].first, "begin")) continue;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r->marks[i - 1].second, r->marks[i].second);
        bool found = false;
        for (auto &a : acc)
            if (a.first == r->marks[i].first) { a.second += ms; found = true; }
        if (!found) acc.emplace_back(r->marks[i].first, ms);
    }
    drop_marks();
    std::string out = "{";
    for (size_t i = 0; i < acc.size(); ++i) {
        char buf[160];
        snprintf(buf, sizeof(buf), "%s\"%s\": %.4f", i ? ", " : "", acc[i].first.c_str(), acc[i].second / nsteps);
        out += buf;
    }
    out += "}";
    if ((int)out.size() + 1 > cap) return EK_ERR_INVALID;
    memcpy(json, out.c_str(), out.size() + 1);
    return EK_OK;
}

// counters: "nccl_groups" (grouped send/recv calls issued), "kernel_launches", "steps"
ek_status ek_rank_get_counter(ek_rank *r, const char *key, double *value)
{
    if (!r || !key || !value) return EK_ERR_INVALID;
    if (!strcmp(key, "nccl_groups")) { *value = (double)r->nccl_groups; return EK_OK; }
    return ek_get_counter(r->h, key, value);
}

}  // extern "C"
