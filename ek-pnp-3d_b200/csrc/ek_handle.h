// ek_handle.h -- the simulation handle behind the C ABI (one per device).
#pragma once

#include <new>
#include <utility>
#include <vector>

#include "ek_internal.cuh"

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (dev != prev) cudaSetDevice(dev); }
    ~DeviceGuard() { int cur; cudaGetDevice(&cur); if (cur != prev && prev >= 0) cudaSetDevice(prev); }
};

struct ek_handle {
    ek_params p;          // local parameters: NX is the slab width, Lx the GLOBAL length
    EkConst c;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    // x-slab decomposition (ek_slab.cu): rank r owns global columns [r*NX, (r+1)*NX)
    int rank = 0, nranks = 1, NXg = 0;
    bool slab = false;
    double *cp_cols = nullptr;     // LU factor for the distributed z-solve (cached)
    int cp_ky0 = -1, cp_kyl = 0;
    std::string err;

    // storage
    int stream_mode = EK_STREAM_AA;
    bool allocated = false;
    double *lat[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    int cur = 0;            // PUSH: lattice holding the current state
    int parity = 0;         // AA: 0 natural layout, 1 after an even step
    double *wall = nullptr; // scalar-set wall state
    double *fld[EK_NFIELDS] = {};
    bool fld_external[EK_NFIELDS] = {};  // adopted caller arrays (ek_adopt_field): not freed by the handle
    double *dq = nullptr;
    double *phi_old = nullptr;
    EkPoisson poisson;
    EkSlabPoisson sp;     // distributed solve of the x-slab path

    // state machine
    bool fields_ready = false;     // macroscopic arrays hold an initial state
    bool pops_ready = false;       // populations initialised
    bool e_from_arrays = false;    // next LBM pass takes E from the arrays, not from grad(phi)
    bool efield_stale = false;     // Ex/Ey/Ez arrays are older than phi (recomputed on demand)
    bool phi_walls_dirty = true;   // something other than the solver wrote phi: its wall planes must be re-imposed
    int zchunk = 16;               // z-planes per CTA (ek_auto_zchunk at creation; option "zchunk")
    int kernel = 0;                // 0 = four warps + lean interior path, 1 = eight warps, 2 = five warps, 3 = four warps, general path only
    int dc_mode = EK_DC_ZERO;
    int poisson_path = 0;          // 0: xy-FFT + tridiagonal z-solve, 1: odd-extension 3-D FFT
    double dc_ghat0 = 0.0;

    // CUDA graph of a PAIR of coupled steps (even + odd A-A step, or two push steps): small grids are
    // launch-latency bound (the shipped 50x8x51 case: ~25 us of kernels in a 52 us step), and after a
    // pair the parity is back where it was, so one instantiated graph replays for the whole run
    int graph_opt = -1;                    // option "graph": 0 off, 1 on, -1 automatic (grids below 4 M cells)
    cudaGraphExec_t pair_graph = nullptr;
    long long epoch = 0, graph_epoch = -1; // anything that changes a baked-in argument bumps `epoch`
    long long pair_lbm_launches = 0, pair_poisson_launches = 0;
    long long graph_replays = 0;

    // ek_run_from_host: copy stream and events of the pipelined upload / download
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> job_events;

    // counters / profiling
    bool profile = false;
    long long steps = 0, lbm_launches = 0, poisson_launches = 0;
    double lbm_ms = 0.0, poisson_ms = 0.0;
    double lbm_ms_mode[3] = {0.0, 0.0, 0.0};   // per launch mode (A-A even, A-A odd, push)
    std::vector<int> ev_lbm_mode;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_lbm, ev_poi;
};

void ek_compute_consts(const ek_params &p, EkConst &c, bool slab);
int ek_auto_zchunk(const EkConst &c);
ek_status ek_alloc_state(ek_handle *h);
StepArgs ek_step_args(ek_handle *h);

// dumps from host arrays of a whole domain (ek_io.cu; shared with the multi-GPU driver)
struct EkHostFields {
    std::vector<double> f[EK_NFIELDS];
};
struct EkDumpGrid {
    int NX, NY, NZ;
    double dx, dy, dz;
};
void ek_io_extrapolate_walls(EkHostFields &H, int NX, int NY, int NZ);
bool ek_io_write_tecplot(const char *path, const EkDumpGrid &g, const EkHostFields &H, double time, int append, int first);
bool ek_io_write_end(const char *path, const EkDumpGrid &g, const EkHostFields &H, double time);
bool ek_io_read_end(const char *path, size_t cells, EkHostFields &H, double *time, std::string &err);
// header of the binary checkpoint (then 11 fields, then 4 sets x 27 x cells pre-collision populations,
// all in the reference's natural order over the WHOLE domain: single- and multi-GPU runs share the format)
struct EkCkptHeader {
    char magic[8];        // "EKB200C1"
    int NX, NY, NZ, nfields;
    long long steps;
    double time;
};
