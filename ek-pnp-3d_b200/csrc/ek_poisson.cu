// ek_poisson.cu -- spectral Poisson solve for the electric potential.
//
// Replaces fast_Poisson (poisson.cu:75-103) and its helpers odd_extension,
// gpu_derivative, odd_extract, gpu_efield, gpu_bc.  Same discrete operator and
// boundary treatment: spectral in the periodic x and y, second
// differences in z between Dirichlet walls,
//     -(kx^2+ky^2) phi^_j + (phi^_{j-1} - 2 phi^_j + phi^_{j+1})/dz^2 = g^_j ,
//     g = -F (c+ - c-)/eps ,  phi_0 = voltage, phi_{NZ-1} = voltage2 .
//
// Two realisations of that one linear system:
//
//  path 0 (default) "xy-FFT + z-solve": batched 2-D cuFFT D2Z over the NZ-2
//    interior planes of c+ - c- (no packing pass: cuFFT reads the field array
//    in place), one hand-written kernel that scales, applies the wall lift and
//    solves the tridiagonal z-system of every (kx,ky) column with a
//    pre-factorised LU (Thomas) sweep, batched 2-D Z2D straight into phi.
//    ~100 B of DRAM traffic per cell instead of ~560 B, and O(NZ) work per
//    column for any NZ (the reference's NE = 510 costs cuFFT four passes per
//    direction).  The z-system is diagonally dominant; its measured error
//    against the sine-transform solution is <= 1e-13 of max|phi| at NZ = 256
//    (tests/test_parity_gpu.py::test_poisson_paths_agree).
//
//  path 1 "odd extension" (cross-check build only, -DEK_XCHECK -> libek_b200_xcheck.so):
//    the reference's own algorithm on a REAL extended
//    array of length NE = 2(NZ-1): cuFFT D2Z / Z2D 3-D, eigenvalue
//    mu = kx^2 + ky^2 + (4/dz^2) sin^2(kz dz/2) (poisson.cu:174), persistent
//    scratch instead of 3 cudaMalloc + 3 cudaFree per call (poisson.cu:77-102).
//    Kept as the literal cross-check and for EK_DC_LITERAL.
#include "ek_handle.h"

#include <math.h>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {

#ifdef EK_XCHECK
// ------------------------------ path 1 kernels ------------------------------
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
#endif  // EK_XCHECK

// ------------------------------ path 0 kernels ------------------------------
// LU factor of the z-operator of every (kx,ky) column, once per handle:
//   b = -(2 + (kx^2+ky^2) dz^2),  c'_1 = 1/b,  c'_j = 1/(b - c'_{j-1})
__global__ void k_zfactor(int ncols, int NXH, int M, const double *__restrict__ kx, const double *__restrict__ ky,
                          double dz, double *__restrict__ cp)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    const double I = kx[col % NXH], J = ky[col / NXH];
    const double b = -(2.0 + (I * I + J * J) * dz * dz);
    double cprev = 0.0;
    for (int j = 0; j < M; ++j) {
        cprev = 1.0 / (b - cprev);
        cp[(size_t)j * ncols + col] = cprev;
    }
}

// One thread per REAL stream of the half spectrum (re and im parts of a column
// are independent real systems): forward elimination then back substitution
// along z, in place.  x[j][r], j = 0..M-1 <-> planes z = 1..NZ-2.
//   d_j  = dz^2 * ( -(F/eps) * dq^_j  -  [j=0] V0/dz^2 * NXY  -  [j=M-1] V1/dz^2 * NXY )   (lift: (0,0) column only)
//   d'_j = (d_j - d'_{j-1}) c'_j ;   phi^_j = d'_j - c'_j phi^_{j+1}
// The result is scaled by 1/(NX*NY) for cuFFT's unnormalised inverse.
// planes fetched ahead of the recurrence per thread: 16 measured best at 256^3 (Poisson stage
// 0.354 ms against 0.383 with 8 and 0.390 with 32; 64-thread blocks make no difference)
#ifndef EK_ZSOLVE_UNROLL
#define EK_ZSOLVE_UNROLL 16
#endif
#ifndef EK_ZSOLVE_BLOCK
#define EK_ZSOLVE_BLOCK 128
#endif
// Layout: x[blockIdx.y * x_outer + j * xj + r], cp[j * ncols + blockIdx.y * cp_outer + (r >> 1)]: blockIdx.y = 0
// and xj = nreal for the single-GPU half spectrum [j][ky][kx]; blockIdx.y = local ky row and xj = 2*NXg for
// the distributed solve's pencils [ky][j][kx] (ek_slab_poisson.cu).
template <int UNROLL>
__global__ void __launch_bounds__(EK_ZSOLVE_BLOCK) k_zsolve(int nreal, int ncols, int M, double *__restrict__ x,
                                                const double *__restrict__ cp, double scale_dz2, double lift0,
                                                double lift1, double norm, double dc_offset, int lift_r,
                                                long long xj, long long x_outer, int cp_outer)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreal) return;
    const int col = blockIdx.y * cp_outer + (r >> 1);
    x += (size_t)blockIdx.y * x_outer;
    // real part of the (kx,ky) = (0,0) column (lift_r = -1: not in this set of columns)
    const bool lifted = (r == lift_r && blockIdx.y == 0);
    double prev = 0.0;
    int j = 0;
    // forward elimination; loads are independent of the recurrence, so a block
    // of UNROLL planes is fetched before the dependent chain runs
    for (; j + UNROLL <= M; j += UNROLL) {
        double g[UNROLL], c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            g[u] = x[(size_t)(j + u) * xj + r];
            c[u] = cp[(size_t)(j + u) * ncols + col];
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            double d = scale_dz2 * g[u];
            if (lifted) {
                if (j + u == 0) d += lift0;
                if (j + u == M - 1) d += lift1;
            }
            prev = (d - prev) * c[u];
            x[(size_t)(j + u) * xj + r] = prev;
        }
    }
    for (; j < M; ++j) {
        double d = scale_dz2 * x[(size_t)j * xj + r];
        if (lifted) {
            if (j == 0) d += lift0;
            if (j == M - 1) d += lift1;
        }
        prev = (d - prev) * cp[(size_t)j * ncols + col];
        x[(size_t)j * xj + r] = prev;
    }
    // back substitution
    const double off = lifted ? dc_offset : 0.0;
    double phi = prev;
    x[(size_t)(M - 1) * xj + r] = phi * norm + off;
    j = M - 2;
    for (; j - UNROLL + 1 >= 0; j -= UNROLL) {
        double g[UNROLL], c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            g[u] = x[(size_t)(j - u) * xj + r];
            c[u] = cp[(size_t)(j - u) * ncols + col];
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            phi = g[u] - c[u] * phi;
            x[(size_t)(j - u) * xj + r] = phi * norm + off;
        }
    }
    for (; j >= 0; --j) {
        phi = x[(size_t)j * xj + r] - cp[(size_t)j * ncols + col] * phi;
        x[(size_t)j * xj + r] = phi * norm + off;
    }
}

__global__ void k_set_walls(EkConst c, double *__restrict__ phi)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    // every column of the row, ghost columns of a slab included (the neighbours' wall values are the same constants)
    if (x >= c.PX) return;
    const int y = blockIdx.y;
    const size_t top = (size_t)(c.NZ - 1) * c.plane;
    phi[(size_t)y * c.PX + x] = c.voltage;          // poisson.cu:195-197
    phi[top + (size_t)y * c.PX + x] = c.voltage2;   // poisson.cu:199-201
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

#ifdef EK_XCHECK
ek_status create_path1(ek_handle *h, EkPoisson &P, cudaStream_t st)
{
    if (P.plans) return EK_OK;
    const size_t nreal = (size_t)P.NE * P.NY * P.NX;
    const size_t nspec = (size_t)P.NE * P.NY * P.NXH;
    EK_CUDA(h, cudaMalloc((void **)&P.real_ext, nreal * sizeof(double)));
    EK_CUDA(h, cudaMalloc((void **)&P.spec, nspec * sizeof(cufftDoubleComplex)));
    // same transform shape as the reference's plan (main.cu:112), real instead of complex
    EK_CUFFT(h, cufftPlan3d(&P.plan_fwd, P.NE, P.NY, P.NX, CUFFT_D2Z));
    EK_CUFFT(h, cufftPlan3d(&P.plan_inv, P.NE, P.NY, P.NX, CUFFT_Z2D));
    P.plans = true;
    EK_CUFFT(h, cufftSetStream(P.plan_fwd, st));
    EK_CUFFT(h, cufftSetStream(P.plan_inv, st));
    return EK_OK;
}
#endif  // EK_XCHECK

ek_status create_path0(ek_handle *h, EkPoisson &P, const ek_params &p, cudaStream_t st)
{
    if (P.plans2) return EK_OK;
    const int M = P.NZ - 2;
    const int ncols = P.NY * P.NXH;
    EK_CUDA(h, cudaMalloc((void **)&P.spec2, (size_t)M * ncols * sizeof(cufftDoubleComplex)));
    EK_CUDA(h, cudaMalloc((void **)&P.cp, (size_t)M * ncols * sizeof(double)));
    k_zfactor<<<(ncols + 127) / 128, 128, 0, st>>>(ncols, P.NXH, M, P.kx2, P.ky2, p.dz, P.cp);
    EK_CUDA(h, cudaGetLastError());
    // batched 2-D transforms of the interior planes, read from / written to the
    // field arrays in place (row pitch PX, plane pitch NY*PX)
    int n[2] = {P.NY, P.NX};
    int rembed[2] = {P.NY, P.PX};
    int cembed[2] = {P.NY, P.NXH};
    EK_CUFFT(h, cufftPlanMany(&P.plan2_fwd, 2, n, rembed, 1, P.NY * P.PX, cembed, 1, P.NY * P.NXH, CUFFT_D2Z, M));
    EK_CUFFT(h, cufftPlanMany(&P.plan2_inv, 2, n, cembed, 1, P.NY * P.NXH, rembed, 1, P.NY * P.PX, CUFFT_Z2D, M));
    P.plans2 = true;
    EK_CUFFT(h, cufftSetStream(P.plan2_fwd, st));
    EK_CUFFT(h, cufftSetStream(P.plan2_inv, st));
    return EK_OK;
}

}  // namespace

// LU factor for an arbitrary block of columns of the FULL (complex-to-complex
// in x) spectrum: column = iky*NXg + ikx, ky index ky0+iky, used by the
// distributed solve (ek_slab.cu)
__global__ void k_zfactor_cols(int ncols, int NXg, int NY, int ky0, int M, double Lx, double Ly, double dz,
                               double *__restrict__ cp)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    const int ikx = col % NXg, iky = ky0 + col / NXg;
    // wavenumbers in FFT order (main.cu:119-145)
    const double I = (ikx <= NXg / 2 ? (double)ikx : (double)ikx - NXg) * 2.0 * M_PI / Lx;
    const double J = (iky <= NY / 2 ? (double)iky : (double)iky - NY) * 2.0 * M_PI / Ly;
    const double b = -(2.0 + (I * I + J * J) * dz * dz);
    double cprev = 0.0;
    for (int j = 0; j < M; ++j) {
        cprev = 1.0 / (b - cprev);
        cp[(size_t)j * ncols + col] = cprev;
    }
}

void ek_launch_zfactor_cols(int ncols, int NXg, int NY, int ky0, int M, double Lx, double Ly, double dz, double *cp,
                            cudaStream_t st)
{
    k_zfactor_cols<<<(ncols + 127) / 128, 128, 0, st>>>(ncols, NXg, NY, ky0, M, Lx, Ly, dz, cp);
}

void ek_launch_zsolve(int nreal, int ncols, int M, double *x, const double *cp, double scale_dz2, double lift0,
                      double lift1, double norm, double dc_offset, int lift_r, cudaStream_t st)
{
    k_zsolve<EK_ZSOLVE_UNROLL><<<(nreal + EK_ZSOLVE_BLOCK - 1) / EK_ZSOLVE_BLOCK, EK_ZSOLVE_BLOCK, 0, st>>>(nreal, ncols, M, x, cp, scale_dz2, lift0, lift1, norm, dc_offset,
                                                       lift_r, nreal, 0, 0);
}

// pencils x[ky][j][kx] complex (rows local ky rows, NXg complex per row); cp[j][rows*NXg]
void ek_launch_zsolve_rows(int rows, int NXg, int M, double *x, const double *cp, double scale_dz2, double lift0,
                           double lift1, double norm, double dc_offset, bool has_dc, cudaStream_t st)
{
    const int nreal = 2 * NXg;
    dim3 grid((nreal + EK_ZSOLVE_BLOCK - 1) / EK_ZSOLVE_BLOCK, rows);
    k_zsolve<EK_ZSOLVE_UNROLL><<<grid, EK_ZSOLVE_BLOCK, 0, st>>>(nreal, rows * NXg, M, x, cp, scale_dz2, lift0, lift1, norm, dc_offset,
                                        has_dc ? 0 : -1, (long long)nreal, (long long)M * nreal, NXg);
}

void ek_launch_set_walls(const EkConst &c, double *phi, cudaStream_t st)
{
    k_set_walls<<<dim3((c.PX + 127) / 128, c.NY), 128, 0, st>>>(c, phi);
}

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
    (void)st;
    return EK_OK;
}

void ek_poisson_destroy(EkPoisson &P)
{
    if (P.plans) { cufftDestroy(P.plan_fwd); cufftDestroy(P.plan_inv); P.plans = false; }
    if (P.plans2) { cufftDestroy(P.plan2_fwd); cufftDestroy(P.plan2_inv); P.plans2 = false; }
    cudaFree(P.real_ext); cudaFree(P.spec); cudaFree(P.kx2); cudaFree(P.ky2); cudaFree(P.kz_term);
    cudaFree(P.spec2); cudaFree(P.cp);
    P.real_ext = nullptr; P.spec = nullptr; P.kx2 = P.ky2 = P.kz_term = nullptr;
    P.spec2 = nullptr; P.cp = nullptr;
}

ek_status ek_poisson_solve(ek_handle *h, EkPoisson &P, const ek_params &p, const EkConst &c, const double *dq,
                           double *phi, double *Ex, double *Ey, double *Ez, int path, int dc_mode, double dc_ghat0,
                           cudaStream_t st, int *launches)
{
    dim3 b(128);
    int n = 0;
    // EK_DC_LITERAL is only defined for the odd-extension transform (there is no kz = 0 mode in the z-solve)
    if (path == 0 && dc_mode != EK_DC_LITERAL) {
        ek_status s0 = create_path0(h, P, p, st);
        if (s0 != EK_OK) return s0;
        const int M = c.NZ - 2, ncols = c.NY * P.NXH, nreal = 2 * ncols;
        EK_CUFFT(h, cufftExecD2Z(P.plan2_fwd, const_cast<double *>(dq) + c.plane, P.spec2));
        const double nxy = (double)c.NX * (double)c.NY;
        const double size = (double)((unsigned int)c.NX * (unsigned int)c.NY * (unsigned int)P.NE);
        const double off = dc_mode == EK_DC_PRESCRIBED ? -dc_ghat0 / size : 0.0;
        k_zsolve<EK_ZSOLVE_UNROLL><<<(nreal + EK_ZSOLVE_BLOCK - 1) / EK_ZSOLVE_BLOCK, EK_ZSOLVE_BLOCK, 0, st>>>(nreal, ncols, M, reinterpret_cast<double *>(P.spec2), P.cp,
                                                           -(c.CtoC / c.eps) * c.dz * c.dz, -c.voltage * nxy,
                                                           -c.voltage2 * nxy, 1.0 / nxy, off, 0, nreal, 0, 0);
        EK_CUFFT(h, cufftExecZ2D(P.plan2_inv, P.spec2, phi + c.plane));
        n = 1;
        // the transforms never touch the wall planes: re-impose them only when something else wrote phi
        // (start-up relaxation, uploads; poisson.cu:195-201 does it on every call)
        if (h->phi_walls_dirty || h->fld_external[EK_PHI]) {   // an adopted array may be written by its owner
            k_set_walls<<<dim3((c.PX + 127) / 128, c.NY), b, 0, st>>>(c, phi);
            h->phi_walls_dirty = false;
            n = 2;
        }
    } else {
#ifndef EK_XCHECK
        ek_set_error(h, "Poisson path 1 / EK_DC_LITERAL (the reference's odd-extension transform) is only in the "
                        "cross-check build libek_b200_xcheck.so");
        return EK_ERR_INVALID;
#else
        ek_status s1 = create_path1(h, P, st);
        if (s1 != EK_OK) return s1;
        dim3 ge((c.NX + 127) / 128, c.NY, P.NE), gs((P.NXH + 127) / 128, c.NY, P.NE), gz((c.NX + 127) / 128, c.NY, c.NZ);
        k_pack_odd<<<ge, b, 0, st>>>(c, P.NE, dq, P.real_ext, c.eps);
        EK_CUFFT(h, cufftExecD2Z(P.plan_fwd, P.real_ext, P.spec));
        k_divide<<<gs, b, 0, st>>>(P.NXH, c.NY, P.kx2, P.ky2, P.kz_term, P.spec, dc_mode, dc_ghat0);
        EK_CUFFT(h, cufftExecZ2D(P.plan_inv, P.spec, P.real_ext));
        const double size = (double)((unsigned int)c.NX * (unsigned int)c.NY * (unsigned int)P.NE);  // LBM.h:38
        k_unpack<<<gz, b, 0, st>>>(c, P.real_ext, size, phi);
        h->phi_walls_dirty = false;
        n = 3;
#endif
    }
    if (Ex) { ek_launch_efield(c, phi, Ex, Ey, Ez, st); ++n; }
    if (launches) *launches += n;
    EK_CUDA(h, cudaGetLastError());
    return EK_OK;
}
