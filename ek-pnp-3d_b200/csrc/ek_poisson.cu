// ek_poisson.cu -- spectral Poisson solve for the electric potential.
//
// Replaces fast_Poisson (poisson.cu:75-103) and its helpers odd_extension,
// gpu_derivative, odd_extract, gpu_efield, gpu_bc.  Same discrete operator and
// boundary treatment: second differences in z with Dirichlet
// walls realised as an odd extension of length NE = 2(NZ-1), spectral in the
// periodic x and y, eigenvalue mu = kx^2 + ky^2 + (4/dz^2) sin^2(kz dz/2).
//
// What changed: the extended array is REAL, so the transforms are cuFFT D2Z /
// Z2D (half the data of the reference's Z2Z); scratch is persistent (the
// reference cudaMallocs/cudaFrees 3 x 8x-oversized buffers per call,
// poisson.cu:77-79,100-102); the wavenumber terms come from small tables
// instead of one sin() per element per step (poisson.cu:174); pack, eigenvalue
// division and unpack are hand-written kernels around the two cuFFT calls.
#include "ek_internal.cuh"

#include <math.h>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {

// odd_extension (poisson.cu:114-158) on a real array; dq = c+ - c-
__global__ void k_pack_odd(EkConst c, int NE, const double *__restrict__ dq, double *__restrict__ ext, double eps)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= c.NX) return;
    const int y = blockIdx.y, z = blockIdx.z, NZ = c.NZ;
    const double dz = c.dz;
    double v;
    if (z == 0) v = 0.0;
    else if (z == 1) v = -c.CtoC * dq[(size_t)z * c.plane + y * c.PX + x] / eps - c.voltage / dz / dz;
    else if (z > 1 && z < NZ - 2) v = -c.CtoC * dq[(size_t)z * c.plane + y * c.PX + x] / eps;
    else if (z == NZ - 2) v = -c.CtoC * dq[(size_t)z * c.plane + y * c.PX + x] / eps - c.voltage2 / dz / dz;
    else if (z == NZ - 1) v = 0.0;
    else if (z == NZ) v = c.CtoC * dq[(size_t)(NE - z) * c.plane + y * c.PX + x] / eps + c.voltage2 / dz / dz;
    else if (z > NZ && z < NE - 1) v = c.CtoC * dq[(size_t)(NE - z) * c.plane + y * c.PX + x] / eps;
    else v = c.CtoC * dq[(size_t)1 * c.plane + y * c.PX + x] / eps + c.voltage / dz / dz;
    ext[((size_t)z * c.NY + y) * c.NX + x] = v;
}

// gpu_derivative (poisson.cu:169-180) on the half spectrum
__global__ void k_divide(int NXH, int NY, const double *__restrict__ kx, const double *__restrict__ ky,
                         const double *__restrict__ kzterm, cufftDoubleComplex *spec, int dc_mode, double dc_ghat0)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= NXH) return;
    const int y = blockIdx.y, z = blockIdx.z;
    const double I = kx[x], J = ky[y];
    double mu = kzterm[z] + I * I + J * J;
    if (y == 0 && x == 0 && z == 0) mu = 1.0;
    const size_t i = ((size_t)z * NY + y) * NXH + x;
    cufftDoubleComplex v = spec[i];
    if (y == 0 && x == 0 && z == 0) {
        // zero by oddness; see ek_set_poisson_dc() in ek_b200.h
        if (dc_mode == EK_DC_ZERO) { v.x = 0.0; v.y = 0.0; }
        else if (dc_mode == EK_DC_PRESCRIBED) { v.x = dc_ghat0; v.y = 0.0; }
    }
    v.x = -v.x / mu;
    v.y = -v.y / mu;
    spec[i] = v;
}

// odd_extract (poisson.cu:191-204)
__global__ void k_unpack(EkConst c, const double *__restrict__ ext, double size, double *__restrict__ phi)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= c.NX) return;
    const int y = blockIdx.y, z = blockIdx.z;
    double v;
    if (z == 0) v = c.voltage;
    else if (z == c.NZ - 1) v = c.voltage2;
    else v = ext[((size_t)z * c.NY + y) * c.NX + x] / size;
    phi[(size_t)z * c.plane + y * c.PX + x] = v;
}

// gpu_efield + gpu_bc (poisson.cu:40-69) in one pass
__global__ void k_efield(EkConst c, const double *__restrict__ phi, double *ex, double *ey, double *ez)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= c.NX) return;
    const int y = blockIdx.y, z = blockIdx.z;
    const int xm = x == 0 ? c.xlo : x - 1, xp = x == c.NX - 1 ? c.xhi : x + 1;
    const int ym = y == 0 ? c.NY - 1 : y - 1, yp = y == c.NY - 1 ? 0 : y + 1;
    const int zc = z < 1 ? 1 : (z > c.NZ - 2 ? c.NZ - 2 : z);
    const size_t zb = (size_t)z * c.plane;
    const size_t i = zb + y * c.PX + x;
    ex[i] = 0.5 * (phi[zb + y * c.PX + xm] - phi[zb + y * c.PX + xp]) / c.dx;
    ey[i] = 0.5 * (phi[zb + ym * c.PX + x] - phi[zb + yp * c.PX + x]) / c.dy;
    ez[i] = 0.5 * (phi[(size_t)(zc - 1) * c.plane + y * c.PX + x] - phi[(size_t)(zc + 1) * c.plane + y * c.PX + x]) / c.dz;
}

}  // namespace

void ek_launch_efield(const EkConst &c, const double *phi, double *Ex, double *Ey, double *Ez, cudaStream_t st)
{
    dim3 b(128), g((c.NX + 127) / 128, c.NY, c.NZ);
    k_efield<<<g, b, 0, st>>>(c, phi, Ex, Ey, Ez);
}

ek_status ek_poisson_create(ek_handle *h, EkPoisson &P, const ek_params &p, int PX, cudaStream_t st)
{
    P.NX = p.NX; P.NY = p.NY; P.NZ = p.NZ; P.PX = PX;
    P.NE = 2 * (p.NZ - 1);          // LBM.h:37
    P.NXH = p.NX / 2 + 1;
    const size_t nreal = (size_t)P.NE * P.NY * P.NX;
    const size_t nspec = (size_t)P.NE * P.NY * P.NXH;
    EK_CUDA(h, cudaMalloc((void **)&P.real_ext, nreal * sizeof(double)));
    EK_CUDA(h, cudaMalloc((void **)&P.spec, nspec * sizeof(cufftDoubleComplex)));
    // wavenumber tables in FFT order (main.cu:119-145)
    std::vector<double> kx(P.NXH), ky(P.NY), kzt(P.NE);
    for (int i = 0; i < P.NXH; ++i) kx[i] = (double)i * 2.0 * M_PI / p.Lx;
    for (int i = 0; i < P.NY; ++i)
        ky[i] = (i <= P.NY / 2 ? (double)i : (double)i - P.NY) * 2.0 * M_PI / p.Ly;
    for (int i = 0; i < P.NE; ++i) {
        const double K = (i <= P.NE / 2 ? (double)i : (double)i - P.NE) * 2.0 * M_PI / (P.NE * p.dz);
        kzt[i] = (4.0 / p.dz / p.dz) * (sin(K * p.dz * 0.5) * sin(K * p.dz * 0.5));  // poisson.cu:174
    }
    EK_CUDA(h, cudaMalloc((void **)&P.kx2, P.NXH * sizeof(double)));
    EK_CUDA(h, cudaMalloc((void **)&P.ky2, P.NY * sizeof(double)));
    EK_CUDA(h, cudaMalloc((void **)&P.kz_term, P.NE * sizeof(double)));
    EK_CUDA(h, cudaMemcpy(P.kx2, kx.data(), P.NXH * sizeof(double), cudaMemcpyHostToDevice));
    EK_CUDA(h, cudaMemcpy(P.ky2, ky.data(), P.NY * sizeof(double), cudaMemcpyHostToDevice));
    EK_CUDA(h, cudaMemcpy(P.kz_term, kzt.data(), P.NE * sizeof(double), cudaMemcpyHostToDevice));
    // same transform shape as the reference's plan (main.cu:112), real instead of complex
    EK_CUFFT(h, cufftPlan3d(&P.plan_fwd, P.NE, P.NY, P.NX, CUFFT_D2Z));
    EK_CUFFT(h, cufftPlan3d(&P.plan_inv, P.NE, P.NY, P.NX, CUFFT_Z2D));
    P.plans = true;
    EK_CUFFT(h, cufftSetStream(P.plan_fwd, st));
    EK_CUFFT(h, cufftSetStream(P.plan_inv, st));
    return EK_OK;
}

void ek_poisson_destroy(EkPoisson &P)
{
    if (P.plans) { cufftDestroy(P.plan_fwd); cufftDestroy(P.plan_inv); P.plans = false; }
    cudaFree(P.real_ext); cudaFree(P.spec); cudaFree(P.kx2); cudaFree(P.ky2); cudaFree(P.kz_term);
    P.real_ext = nullptr; P.spec = nullptr; P.kx2 = P.ky2 = P.kz_term = nullptr;
}

ek_status ek_poisson_solve(ek_handle *h, EkPoisson &P, const EkConst &c, const double *dq, double *phi,
                           double *Ex, double *Ey, double *Ez, int dc_mode, double dc_ghat0, cudaStream_t st,
                           int *launches)
{
    dim3 b(128);
    dim3 ge((c.NX + 127) / 128, c.NY, P.NE), gs((P.NXH + 127) / 128, c.NY, P.NE), gz((c.NX + 127) / 128, c.NY, c.NZ);
    k_pack_odd<<<ge, b, 0, st>>>(c, P.NE, dq, P.real_ext, c.eps);
    EK_CUFFT(h, cufftExecD2Z(P.plan_fwd, P.real_ext, P.spec));
    k_divide<<<gs, b, 0, st>>>(P.NXH, c.NY, P.kx2, P.ky2, P.kz_term, P.spec, dc_mode, dc_ghat0);
    EK_CUFFT(h, cufftExecZ2D(P.plan_inv, P.spec, P.real_ext));
    const double size = (double)((unsigned int)c.NX * (unsigned int)c.NY * (unsigned int)P.NE);  // LBM.h:38
    k_unpack<<<gz, b, 0, st>>>(c, P.real_ext, size, phi);
    int n = 3;
    if (Ex) { ek_launch_efield(c, phi, Ex, Ey, Ez, st); ++n; }
    if (launches) *launches += n;
    EK_CUDA(h, cudaGetLastError());
    return EK_OK;
}
