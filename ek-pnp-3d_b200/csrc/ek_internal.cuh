// ek_internal.cuh -- shared definitions of the B200-native coupled step.
//
// Reference semantics followed here are cited as file:line relative to the
// reference repository (gyf135/EK-PNP-3D); SURVEY.md App. A is the compact
// numerical specification.
#pragma once

#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <map>
#include <string>

#include "../../include/ek_b200.h"

// ---------------------------------------------------------------------------
// D3Q27 velocity set (LBM.cu:872-1103, :639-644, :1983-2092).
// Odd d and d+1 are opposite directions.
// ---------------------------------------------------------------------------
__host__ __device__ constexpr int ek_cx(int d)
{
    constexpr int t[27] = {0, 1,-1, 0, 0, 0, 0, 1,-1, 1,-1, 0, 0, 1,-1, 1,-1, 0, 0, 1,-1, 1,-1, 1,-1,-1, 1};
    return t[d];
}
__host__ __device__ constexpr int ek_cy(int d)
{
    constexpr int t[27] = {0, 0, 0, 1,-1, 0, 0, 1,-1, 0, 0, 1,-1,-1, 1, 0, 0, 1,-1, 1,-1, 1,-1,-1, 1, 1,-1};
    return t[d];
}
__host__ __device__ constexpr int ek_cz(int d)
{
    constexpr int t[27] = {0, 0, 0, 0, 0, 1,-1, 0, 0, 1,-1, 1,-1, 0, 0,-1, 1,-1, 1, 1,-1,-1, 1, 1,-1, 1,-1};
    return t[d];
}
__host__ __device__ constexpr int ek_opp(int d) { return d == 0 ? 0 : ((d & 1) ? d + 1 : d - 1); }
// weight class: 0 rest, 1 axis (ws), 2 face diagonal (wa), 3 cube diagonal (wd)
__host__ __device__ constexpr int ek_wclass(int d) { return d == 0 ? 0 : (d <= 6 ? 1 : (d <= 18 ? 2 : 3)); }
// moving-wall sign table of gpu_boundary (LBM.cu:1902-1927), asymmetries included
__host__ __device__ constexpr int ek_uwsign(int d)
{
    constexpr int t[27] = {0, +1,-1, +1, 0, 0, 0, +1,-1, +1,-1, 0, 0, +1,-1, +1,-1, 0, 0, +1,-1, +1,-1, +1,-1, -1,+1};
    return t[d];
}

enum { EK_MODE_AA_EVEN = 0, EK_MODE_AA_ODD = 1, EK_MODE_PUSH = 2 };

// ---------------------------------------------------------------------------
// Population storage, per set: [z][y][x-tile][27 slots][32 lanes] doubles.
// One x-tile of one row is a contiguous 6912 B block; slot d of a node is
// 256*d bytes from slot 0 (an immediate in every load/store), and the 32 lanes
// of a warp read one aligned 256 B segment per slot.
// ---------------------------------------------------------------------------
#define EK_TILE 32
#define EK_TILE_ELEMS (27 * EK_TILE)
__host__ __device__ inline unsigned ek_lat_col(int col)
{
    return (unsigned)(col >> 5) * (unsigned)EK_TILE_ELEMS + (unsigned)(col & 31);
}

// ---------------------------------------------------------------------------
// Constants handed to every kernel by value.
// ---------------------------------------------------------------------------
struct EkConst {
    int NX, NY, NZ;      // local grid
    int PX;              // row pitch of every array (>= NX; ghost columns live in [NX, PX))
    int xlo, xhi;        // column index of the x-1 neighbour of x=0 and of the x+1 neighbour of x=NX-1
    long long plane;     // NY*PX
    long long N;         // NZ*NY*PX : elements per field array
    long long dq_sy, dq_sz;  // c+ - c- lives at dq[z*dq_sz + y*dq_sy + x]: (PX, plane) like the fields, or
                         // (NZ*NX, NX) = rows for the y-transform of the distributed Poisson stage
    int NXT;             // x-tiles per row of the population lattice: ceil(PX/32)
    unsigned lrow;       // lattice elements per (z,y) row:  NXT*27*32
    unsigned lplane;     // lattice elements per z plane:    NY*lrow
    unsigned long long Nlat;  // lattice elements per set:   NZ*lplane  (< 2^32)
    double cflinv;       // 1/CFL                       (LBM.cu:1112)
    double cflinv2;      // cflinv*cflinv/cs_square      (LBM.cu:1115)
    double inv_cs2;      // 1/cs_square
    double cs_square, CFL;
    double tfac;         // 1/cs_square/CFL              (LBM.cu:854)
    double dt;
    double CtoC, Ext, exf, eps;
    double rho0, Ra, nu, D;
    double K, Kn;
    double w[4];         // w0, ws, wa, wd              (LBM.h:109-112)
    double coe[4];       // w/cs_square                  (LBM.cu:1107-1110)
    double wp[4], wm[4]; // omega_plus*dt, omega_minus*dt per set (LBM.cu:488-495,1700-1707)
    double sp, sm;       // 1 - dt*omega/2 for the fluid (LBM.cu:1660-1661)
    double multi[4];     // moving-wall terms per class  (LBM.cu:1896-1898)
    double twoTw[4];     // 2*TH*w per class             (LBM.cu:2226-2229)
    double dx, dy, dz;
    double voltage, voltage2;
};

struct StepArgs {
    EkConst c;
    double *in[4];        // population lattices (tiled layout above), Nlat doubles per set
    double *out[4];       // == in for the A-A scheme
    double *wall;         // scalar-set wall state: [3 sets][2 planes][27][plane]
    const double *phi;    // potential (E = -grad phi fused into the step)
    const double *E[3];   // optional explicit field arrays (first step after init / shim)
    double *dq;           // c+ - c- for the Poisson stage
    double *fld[7];       // rho ux uy uz charge chargen T (only when fields are written)
    int zchunk;
    int zblock0, nzblocks;  // sub-range of z-chunks for this launch (nzblocks = 0: all)
    int row_imm;            // 1: odd lean launches may take the kernel with the row stride as an immediate
    int xt_mode;            // x-tiles of this launch: 0 all, 1 the two boundary tiles of the row (0 and last), 2 the interior ones
    // x-marching launch of the odd A-A step (ek_march_kernel): deep-interior planes [march_z0, march_z0 +
    // march_planes) walk their x-rows; the wall-adjacent plane ranges take the general node path
    int march_z0, march_planes;
    int wall_n, wall_z0[2], wall_z1[2];
};

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
struct ek_handle;
void ek_set_error(ek_handle *h, const std::string &msg);

#define EK_CUDA(h, call)                                                        \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) {                                                \
            ek_set_error((h), std::string(#call) + ": " + cudaGetErrorString(_e)); \
            return EK_ERR_CUDA;                                                 \
        }                                                                       \
    } while (0)

#define EK_CUFFT(h, call)                                                       \
    do {                                                                        \
        cufftResult _r = (call);                                                \
        if (_r != CUFFT_SUCCESS) {                                              \
            ek_set_error((h), std::string(#call) + ": cufft error " + std::to_string((int)_r)); \
            return EK_ERR_CUFFT;                                                \
        }                                                                       \
    } while (0)

// ---------------------------------------------------------------------------
// Poisson stage (ek_poisson.cu)
// ---------------------------------------------------------------------------
struct EkPoisson {
    int NX = 0, NY = 0, NZ = 0, NE = 0, NXH = 0, PX = 0;
    cufftHandle plan_fwd = 0, plan_inv = 0;
    bool plans = false;
    double *real_ext = nullptr;          // [NE][NY][NX]
    cufftDoubleComplex *spec = nullptr;  // [NE][NY][NXH]
    double *kx2 = nullptr, *ky2 = nullptr, *kz_term = nullptr;  // device tables
    // path 0: batched 2-D transforms + tridiagonal z-solve
    cufftHandle plan2_fwd = 0, plan2_inv = 0;
    bool plans2 = false;
    cufftDoubleComplex *spec2 = nullptr;  // [NZ-2][NY][NXH]
    double *cp = nullptr;                 // LU factor c'_j of every column, [NZ-2][NY*NXH]
};

ek_status ek_poisson_create(ek_handle *h, EkPoisson &P, const ek_params &p, int PX, cudaStream_t st);
void ek_poisson_destroy(EkPoisson &P);
// dq = c+ - c-  ->  phi (and Ex,Ey,Ez when E != nullptr)
ek_status ek_poisson_solve(ek_handle *h, EkPoisson &P, const ek_params &p, const EkConst &c, const double *dq,
                           double *phi, double *Ex, double *Ey, double *Ez, int path, int dc_mode, double dc_ghat0,
                           cudaStream_t st, int *launches);
void ek_launch_efield(const EkConst &c, const double *phi, double *Ex, double *Ey, double *Ez, cudaStream_t st);
void ek_launch_zfactor_cols(int ncols, int NXg, int NY, int ky0, int M, double Lx, double Ly, double dz, double *cp,
                            cudaStream_t st);
void ek_launch_zsolve(int nreal, int ncols, int M, double *x, const double *cp, double scale_dz2, double lift0,
                      double lift1, double norm, double dc_offset, int lift_r, cudaStream_t st);
void ek_launch_zsolve_rows(int rows, int NXg, int M, double *x, const double *cp, double scale_dz2, double lift0,
                           double lift1, double norm, double dc_offset, bool has_dc, cudaStream_t st);
void ek_launch_set_walls(const EkConst &c, double *phi, cudaStream_t st);

// distributed Poisson stage of the x-slab path (ek_slab_poisson.cu)
#define EK_MAX_RANKS 16
#define EK_MAX_CHUNKS 16
struct EkSlabPoisson {
    bool ready = false;
    int P = 1, r = 0;                 // ranks, my rank
    int NXl = 0, NXg = 0, NY = 0, NYH = 0, kyl = 0, M = 0;
    int K = 1;                        // z-chunks
    int z0[EK_MAX_CHUNKS + 1] = {};   // chunk bounds (interior-plane index 0..M)
    int block0[EK_MAX_CHUNKS + 1] = {};  // first LBM z-block of each chunk
    double *A = nullptr;              // [NY][NZ][NXl] real rows for the inverse y-transforms (the forward
                                      // ones read c+ - c-, which the LBM kernel writes in this layout)
    cufftDoubleComplex *S = nullptr;  // send buffers, per chunk [P*kyl][nzc][NXl]
    cufftDoubleComplex *R = nullptr;  // receive buffers, same shape
    cufftDoubleComplex *X = nullptr;  // full-x pencils [kyl][M][NXg]
    double *cp = nullptr;             // LU factors of my columns [M][kyl*NXg]
    std::map<int, cufftHandle> plan_yf, plan_yb;  // y-transforms per chunk height
    cufftHandle plan_x = 0;
    bool plan_x_ok = false;
    cufftDoubleComplex *peerX[EK_MAX_RANKS] = {};  // direct peer-memory transport (optional)
    cufftDoubleComplex *peerR[EK_MAX_RANKS] = {};
    bool peer_ipc[EK_MAX_RANKS] = {};              // mapped through CUDA IPC (to be closed)
    bool dma = false;                              // pushes by strided copies on the copy engines
    // way back with the two ghost columns of phi travelling inside transpose 2 (ek_slab_poisson_enable_ghosts):
    // rows of NXl + 2 columns [my columns, right neighbour's first, left neighbour's last]
    bool ghosts = false;
    cufftDoubleComplex *Sg = nullptr, *Rg = nullptr;   // [P*kyl][M][NXl+2] per chunk, as S / R
    std::map<int, cufftHandle> plan_ybg;               // inverse y-transforms of NXl+2 columns per chunk height
};
void ek_slab_poisson_destroy(ek_handle *h);

// ---------------------------------------------------------------------------
// LBM stage (ek_lbm.cu) and start-up kernels (ek_init.cu)
// ---------------------------------------------------------------------------
cudaError_t ek_launch_step(const StepArgs &a, int mode, bool write_fields, bool e_from_arrays, bool lean,
                           cudaStream_t st);
// odd A-A step with sector-aligned stores (x-marching warps); planes [z0, z1) of the launch
// variant 1: aligned stores; 2: aligned loads and stores, row pointers
cudaError_t ek_launch_march(StepArgs a, bool write_fields, int z0, int z1, int variant, cudaStream_t st);
bool ek_march_applicable(const EkConst &c);
cudaError_t ek_launch_step5(const StepArgs &a, int mode, bool write_fields, bool e_from_arrays, cudaStream_t st);
cudaError_t ek_launch_step8(const StepArgs &a, int mode, bool write_fields, bool e_from_arrays, cudaStream_t st);
cudaError_t ek_launch_export(const StepArgs &a, int mode, int set, double *dst, cudaStream_t st);
cudaError_t ek_launch_import(const StepArgs &a, int set, const double *src, cudaStream_t st);
cudaError_t ek_launch_init_equilibrium(const StepArgs &a, const double *const fld[EK_NFIELDS], cudaStream_t st);
cudaError_t ek_launch_init_equilibrium_range(const StepArgs &a, const double *const fld[EK_NFIELDS], int z0, int z1,
                                             cudaStream_t st);
void ek_launch_initialization(const EkConst &c, const ek_params &p, double *const fld[EK_NFIELDS], cudaStream_t st);
void ek_launch_pbe(const EkConst &c, const ek_params &p, const double *phi, double *charge, double *chargen,
                   double *dq, cudaStream_t st);
void ek_launch_pbe_relax(const EkConst &c, double omega, double *phi, double *phi_old, cudaStream_t st);
