// ek_init.cu -- start-up kernels: the reference's initial state and the
// equilibrium initialisation of the four population sets.
//
// Replaces gpu_initialization (LBM.cu:111-128), gpu_PBE (LBM.cu:139-146),
// gpu_PBE_phi (LBM.cu:131-137) and gpu_init_equilibrium (LBM.cu:162-463).
// The host round trips of the reference's Poisson-Boltzmann loop
// (LBM.cu:101-104) are gone: phi_old is updated on the device.
#include "ek_internal.cuh"

namespace {

__device__ __forceinline__ bool cell_of_thread(const EkConst &c, int &x, int &y, int &z, int &i)
{
    x = blockIdx.x * blockDim.x + threadIdx.x;
    y = blockIdx.y;
    z = blockIdx.z;
    i = (int)(z * c.plane) + y * c.PX + x;
    return x < c.NX;
}

__global__ void k_initialization(EkConst c, double rho0, double voltage, double TH, double Lz, double dz,
                                 double *r, double *u, double *v, double *w, double *ch, double *chn, double *fi,
                                 double *T, double *ex, double *ey, double *ez)
{
    int x, y, z, i;
    if (!cell_of_thread(c, x, y, z, i)) return;
    r[i] = rho0;
    ch[i] = 0.0;
    chn[i] = 0.0;
    fi[i] = voltage;
    u[i] = 0.0; v[i] = 0.0; w[i] = 0.0;
    ex[i] = 0.0; ey[i] = 0.0; ez[i] = 0.0;
    T[i] = TH * (Lz - dz * z) / Lz;  // LBM.cu:127
}

// c+- = c_inf * exp(-+ e*phi/kB/T0)  (LBM.cu:144-145); also c+ - c- for the solver
__global__ void k_pbe(EkConst c, double chargeinf, double electron, double kB, double roomT, const double *fi,
                      double *ch, double *chn, double *dq)
{
    int x, y, z, i;
    if (!cell_of_thread(c, x, y, z, i)) return;
    const double a = chargeinf * exp(-electron * fi[i] / kB / roomT);
    const double b = chargeinf * exp(electron * fi[i] / kB / roomT);
    ch[i] = a;
    chn[i] = b;
    dq[(size_t)z * c.dq_sz + (size_t)y * c.dq_sy + x] = a - b;
}

// phi <- omega*phi + (1-omega)*phi_old ; phi_old <- phi  (LBM.cu:136, 101-104)
__global__ void k_pbe_relax(EkConst c, double omega, double *fi, double *fi_old)
{
    int x, y, z, i;
    if (!cell_of_thread(c, x, y, z, i)) return;
    const double v = omega * fi[i] + (1.0 - omega) * fi_old[i];
    fi[i] = v;
    fi_old[i] = v;
}

// literal direction terms of LBM.cu:230-462
__device__ __forceinline__ double cidot_literal(int d, double tx, double ty, double tz)
{
    switch (d) {
    case 1: return tx;            case 2: return -tx;
    case 3: return ty;            case 4: return -ty;
    case 5: return tz;            case 6: return -tz;
    case 7: return tx + ty;       case 8: return -ty - tx;
    case 9: return tx + tz;       case 10: return -tx - tz;
    case 11: return tz + ty;      case 12: return -ty - tz;
    case 13: return tx - ty;      case 14: return ty - tx;
    case 15: return tx - tz;      case 16: return tz - tx;
    case 17: return ty - tz;      case 18: return tz - ty;
    case 19: return tx + ty + tz; case 20: return -ty - tx - tz;
    case 21: return tx + ty - tz; case 22: return tz - tx - ty;
    case 23: return tx + tz - ty; case 24: return ty - tx - tz;
    case 25: return ty + tz - tx; case 26: return tx - ty - tz;
    default: return 0.0;
    }
}

// Equilibrium populations in the natural layout (slot d at the node holds the
// population that arrives there along d), plus the wall-node side buffer of
// the three scalar sets.
__global__ void k_init_equilibrium(StepArgs a, const double *r, const double *u, const double *v, const double *w,
                                   const double *ch, const double *chn, const double *T, const double *ex,
                                   const double *ey, const double *ez, double cs_square, double CFL, int zfirst)
{
    const EkConst &c = a.c;
    int x, y, z, i;
    if (!cell_of_thread(c, x, y, z, i)) return;
    z += zfirst;                       // launch over a range of planes [zfirst, zfirst + gridDim.z)
    i += (int)(zfirst * c.plane);
    const double ux = u[i], uy = v[i], uz = w[i];
    const double Ex = ex[i], Ey = ey[i], Ez = ez[i];
    const bool wall = (z == 0 || z == c.NZ - 1);
    const size_t li = (size_t)z * c.lplane + (size_t)y * c.lrow + ek_lat_col(x);
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        const double m = s == 0 ? r[i] : (s == 1 ? ch[i] : (s == 2 ? chn[i] : T[i]));
        const double Ks = s == 1 ? c.K : (s == 2 ? c.Kn : 0.0);
        double vx = ux, vy = uy, vz = uz;
        if (s == 1 || s == 2) { vx = ux + Ks * Ex; vy = uy + Ks * Ey; vz = uz + Ks * Ez; }
        const double omusq = 1.0 - 0.5 * (vx * vx + vy * vy + vz * vz) / cs_square;
        const double tx = vx / cs_square / CFL, ty = vy / cs_square / CFL, tz = vz / cs_square / CFL;
        double *lat = a.in[s];
        double *Wn = nullptr;
        if (s > 0 && wall)
            Wn = a.wall + (size_t)(s - 1) * 2 * 27 * c.plane + (size_t)(z == 0 ? 0 : 27) * c.plane + y * c.PX + x;
        for (int d = 0; d < 27; ++d) {
            const double wm = c.w[ek_wclass(d)] * m;
            const double cd = cidot_literal(d, tx, ty, tz);
            const double e = d == 0 ? wm * (omusq) : wm * (omusq + cd * (1.0 + 0.5 * cd));
            lat[li + (size_t)d * EK_TILE] = e;
            if (Wn) Wn[(size_t)d * c.plane] = e;
        }
    }
}

dim3 cell_grid(const EkConst &c, dim3 &block)
{
    block = dim3(128, 1, 1);
    return dim3((c.NX + 127) / 128, c.NY, c.NZ);
}

}  // namespace

void ek_launch_initialization(const EkConst &c, const ek_params &p, double *const f[EK_NFIELDS], cudaStream_t st)
{
    dim3 b, g = cell_grid(c, b);
    k_initialization<<<g, b, 0, st>>>(c, p.rho0, p.voltage, p.TH, p.Lz, p.dz, f[EK_RHO], f[EK_UX], f[EK_UY], f[EK_UZ],
                                      f[EK_CHARGE], f[EK_CHARGEN], f[EK_PHI], f[EK_T], f[EK_EX], f[EK_EY], f[EK_EZ]);
}

void ek_launch_pbe(const EkConst &c, const ek_params &p, const double *phi, double *charge, double *chargen,
                   double *dq, cudaStream_t st)
{
    dim3 b, g = cell_grid(c, b);
    k_pbe<<<g, b, 0, st>>>(c, p.chargeinf, p.electron, p.kB, p.roomT, phi, charge, chargen, dq);
}

void ek_launch_pbe_relax(const EkConst &c, double omega, double *phi, double *phi_old, cudaStream_t st)
{
    dim3 b, g = cell_grid(c, b);
    k_pbe_relax<<<g, b, 0, st>>>(c, omega, phi, phi_old);
}

cudaError_t ek_launch_init_equilibrium(const StepArgs &a, const double *const f[EK_NFIELDS], cudaStream_t st)
{
    return ek_launch_init_equilibrium_range(a, f, 0, a.c.NZ, st);
}

// the planes [z0, z1) only (ek_run_from_host pipelines the upload of the arrays against this kernel)
cudaError_t ek_launch_init_equilibrium_range(const StepArgs &a, const double *const f[EK_NFIELDS], int z0, int z1,
                                             cudaStream_t st)
{
    dim3 b, g = cell_grid(a.c, b);
    g.z = z1 - z0;
    k_init_equilibrium<<<g, b, 0, st>>>(a, f[EK_RHO], f[EK_UX], f[EK_UY], f[EK_UZ], f[EK_CHARGE], f[EK_CHARGEN],
                                        f[EK_T], f[EK_EX], f[EK_EY], f[EK_EZ], a.c.cs_square, a.c.CFL, z0);
    return cudaGetLastError();
}
