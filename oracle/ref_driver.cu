/*
 * ref_driver.cu -- host driver around the UNMODIFIED reference kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is ours; the reference's LBM.cu and
 * poisson.cu are #included from where they lie under /root/reference (the
 * reference itself is one translation unit, main.cu:12-15) and only LBM.h is
 * regenerated per case by oracle/build_ref.py, because every input parameter
 * of the reference is a compile-time constant there.  The
 * built binary goes to oracle/_ref/ (git-ignored, shipped to the GPU box).
 *
 * It performs the same call sequence as the reference's main() -- setup
 * (main.cu:23-35, 78-152), initialization() or a raw initial state, optional
 * deterministic perturbation, init_equilibrium() (main.cu:174), then N times
 * stream_collide_save() + fast_Poisson() (main.cu:192-198) -- and adds what
 * the reference lacks: raw fp64 dumps and step-only CUDA-event timing.
 *
 * usage: ek_ref [--steps N] [--warmup W] [--perturb AMP] [--load-init FILE]
 *               [--dump-init FILE] [--dump-final FILE] [--dump-pops FILE]
 *               [--dump-dc FILE] [--split] [--quiet]
 *
 * --dump-dc records, per step, the (0,0,0) coefficient of the reference's own
 * forward transform: the driver repeats extension() + cufftExecZ2Z with the
 * reference's plan on a scratch buffer right before each fast_Poisson() call
 * (same plan, same input: the same bits fast_Poisson sees).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "LBM.h"      /* generated copy with the case's constants (first on -I) */
#include "LBM.cu"     /* /root/reference/LBM.cu, unmodified */
#include "poisson.cu" /* /root/reference/poisson.cu, unmodified */
#include <cuda_runtime.h>
#include <cufft.h>

static double *field_ptr(int id)
{
    switch (id) {
    case 0: return rho_gpu;  case 1: return ux_gpu;  case 2: return uy_gpu;  case 3: return uz_gpu;
    case 4: return charge_gpu; case 5: return chargen_gpu; case 6: return phi_gpu; case 7: return T_gpu;
    case 8: return Ex_gpu;   case 9: return Ey_gpu;  default: return Ez_gpu;
    }
}

static void dump_fields(const char *path)
{
    const size_t N = (size_t)NX * NY * NZ;
    double *h = (double *)malloc(N * sizeof(double));
    FILE *f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    for (int id = 0; id < 11; ++id) {
        checkCudaErrors(cudaMemcpy(h, field_ptr(id), N * sizeof(double), cudaMemcpyDeviceToHost));
        fwrite(h, sizeof(double), N, f);
    }
    fclose(f);
    free(h);
}

static void load_fields(const char *path)
{
    const size_t N = (size_t)NX * NY * NZ;
    double *h = (double *)malloc(N * sizeof(double));
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    for (int id = 0; id < 11; ++id) {
        if (fread(h, sizeof(double), N, f) != N) { fprintf(stderr, "short read %s\n", path); exit(2); }
        checkCudaErrors(cudaMemcpy(field_ptr(id), h, N * sizeof(double), cudaMemcpyHostToDevice));
    }
    fclose(f);
    free(h);
}

static void dump_pops(const char *path)
{
    const size_t N = (size_t)NX * NY * NZ;
    double *h = (double *)malloc(26 * N * sizeof(double));
    FILE *f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    double *rest[4] = {f0_gpu, h0_gpu, hn0_gpu, temp0_gpu};
    double *mov[4] = {f1_gpu, h1_gpu, hn1_gpu, temp1_gpu};
    for (int s = 0; s < 4; ++s) {
        checkCudaErrors(cudaMemcpy(h, rest[s], N * sizeof(double), cudaMemcpyDeviceToHost));
        fwrite(h, sizeof(double), N, f);
        checkCudaErrors(cudaMemcpy(h, mov[s], 26 * N * sizeof(double), cudaMemcpyDeviceToHost));
        fwrite(h, sizeof(double), 26 * N, f);
    }
    fclose(f);
    free(h);
}

/* c+, c-, T *= 1 + amp*sin(2 pi x/NX)*cos(2 pi y/NY)*sin(pi z/(NZ-1)) */
static void perturb_fields(double amp)
{
    const size_t N = (size_t)NX * NY * NZ;
    double *h = (double *)malloc(N * sizeof(double));
    double *targets[3] = {charge_gpu, chargen_gpu, T_gpu};
    for (int k = 0; k < 3; ++k) {
        checkCudaErrors(cudaMemcpy(h, targets[k], N * sizeof(double), cudaMemcpyDeviceToHost));
        for (unsigned z = 0; z < NZ; ++z)
            for (unsigned y = 0; y < NY; ++y)
                for (unsigned x = 0; x < NX; ++x) {
                    const double g = 1.0 + amp * sin(2.0 * M_PI * x / NX) * cos(2.0 * M_PI * y / NY)
                                               * sin(M_PI * z / (NZ - 1));
                    h[scalar_index(x, y, z)] *= g;
                }
        checkCudaErrors(cudaMemcpy(targets[k], h, N * sizeof(double), cudaMemcpyHostToDevice));
    }
    free(h);
}

int main(int argc, char **argv)
{
    int steps = 0, warmup = 0, split = 0, quiet_run = 0;
    double amp = 0.0;
    const char *load_init = NULL, *dump_init = NULL, *dump_final = NULL, *dump_p = NULL, *dump_dc = NULL;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) warmup = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--perturb") && i + 1 < argc) amp = atof(argv[++i]);
        else if (!strcmp(argv[i], "--load-init") && i + 1 < argc) load_init = argv[++i];
        else if (!strcmp(argv[i], "--dump-init") && i + 1 < argc) dump_init = argv[++i];
        else if (!strcmp(argv[i], "--dump-final") && i + 1 < argc) dump_final = argv[++i];
        else if (!strcmp(argv[i], "--dump-pops") && i + 1 < argc) dump_p = argv[++i];
        else if (!strcmp(argv[i], "--dump-dc") && i + 1 < argc) dump_dc = argv[++i];
        else if (!strcmp(argv[i], "--split")) split = 1;
        else if (!strcmp(argv[i], "--quiet")) quiet_run = 1;
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }

    checkCudaErrors(cudaSetDevice(0));
    /* host mirrors of the device constants, device copies of the host-side
     * physics (the reference does this at main.cu:23-35) */
    cudaMemcpyFromSymbol(&dt_host, dt, sizeof(double));
    cudaMemcpyFromSymbol(&Lx_host, Lx, sizeof(double));
    cudaMemcpyFromSymbol(&Ly_host, Ly, sizeof(double));
    cudaMemcpyFromSymbol(&dy_host, dy, sizeof(double));
    cudaMemcpyFromSymbol(&Lz_host, Lz, sizeof(double));
    cudaMemcpyFromSymbol(&dz_host, dz, sizeof(double));
    cudaMemcpyToSymbol(nu, &nu_host, sizeof(double));
    cudaMemcpyToSymbol(uw, &uw_host, sizeof(double));
    cudaMemcpyToSymbol(exf, &exf_host, sizeof(double));
    cudaMemcpyToSymbol(K, &K_host, sizeof(double));
    cudaMemcpyToSymbol(Kn, &Kn_host, sizeof(double));
    cudaMemcpyToSymbol(epsn, &epsn_host, sizeof(double));

    /* the reference's allocations (main.cu:78-109) */
    checkCudaErrors(cudaMalloc((void **)&f0bc, sizeof(double) * NX * NY * 2));
    double **rest[4] = {&f0_gpu, &h0_gpu, &hn0_gpu, &temp0_gpu};
    double **m1[4] = {&f1_gpu, &h1_gpu, &hn1_gpu, &temp1_gpu};
    double **m2[4] = {&f2_gpu, &h2_gpu, &hn2_gpu, &temp2_gpu};
    for (int s = 0; s < 4; ++s) {
        checkCudaErrors(cudaMalloc((void **)rest[s], mem_size_0dir));
        checkCudaErrors(cudaMalloc((void **)m1[s], mem_size_n0dir));
        checkCudaErrors(cudaMalloc((void **)m2[s], mem_size_n0dir));
    }
    double **scal[11] = {&rho_gpu, &ux_gpu, &uy_gpu, &uz_gpu, &charge_gpu, &chargen_gpu,
                         &phi_gpu, &T_gpu, &Ex_gpu, &Ey_gpu, &Ez_gpu};
    for (int k = 0; k < 11; ++k) checkCudaErrors(cudaMalloc((void **)scal[k], mem_size_scalar));
    checkCudaErrors(cudaMalloc((void **)&kx, sizeof(double) * NX));
    checkCudaErrors(cudaMalloc((void **)&ky, sizeof(double) * NY));
    checkCudaErrors(cudaMalloc((void **)&kz, sizeof(double) * NE));
    CHECK_CUFFT(cufftPlan3d(&plan, NE, NY, NX, CUFFT_Z2Z)); /* main.cu:112 */

    /* wavenumbers in FFT order (main.cu:119-152) */
    for (unsigned i = 0; i < NX; ++i)
        kx_host[i] = (i <= NX / 2 ? (double)i : (double)i - NX) * 2.0 * M_PI / Lx_host;
    for (unsigned i = 0; i < NY; ++i)
        ky_host[i] = (i <= NY / 2 ? (double)i : (double)i - NY) * 2.0 * M_PI / Ly_host;
    for (unsigned i = 0; i < NE; ++i)
        kz_host[i] = (i <= NE / 2 ? (double)i : (double)i - NE) * 2.0 * M_PI / (NE * dz_host);
    CHECK(cudaMemcpy(kx, kx_host, sizeof(double) * NX, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(ky, ky_host, sizeof(double) * NY, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(kz, kz_host, sizeof(double) * NE, cudaMemcpyHostToDevice));

    cudaEvent_t e0, e1;
    checkCudaErrors(cudaEventCreate(&e0));
    checkCudaErrors(cudaEventCreate(&e1));

    float init_ms = 0.f;
    if (load_init) {
        load_fields(load_init);
    } else {
        checkCudaErrors(cudaEventRecord(e0, 0));
        initialization(rho_gpu, charge_gpu, chargen_gpu, phi_gpu, ux_gpu, uy_gpu, uz_gpu,
                       Ex_gpu, Ey_gpu, Ez_gpu, T_gpu);
        checkCudaErrors(cudaEventRecord(e1, 0));
        checkCudaErrors(cudaEventSynchronize(e1));
        checkCudaErrors(cudaEventElapsedTime(&init_ms, e0, e1));
    }
    if (amp != 0.0) perturb_fields(amp);
    if (dump_init) dump_fields(dump_init);

    init_equilibrium(f0_gpu, f1_gpu, h0_gpu, h1_gpu, hn0_gpu, hn1_gpu, temp0_gpu, temp1_gpu,
                     rho_gpu, charge_gpu, chargen_gpu, ux_gpu, uy_gpu, uz_gpu, Ex_gpu, Ey_gpu, Ez_gpu, T_gpu);
    t = 0;

    cufftDoubleComplex *dc_in = NULL, *dc_out = NULL;
    double *dc_host = NULL;
    if (dump_dc) {
        checkCudaErrors(cudaMalloc((void **)&dc_in, sizeof(cufftDoubleComplex) * NX * NY * NE));
        checkCudaErrors(cudaMalloc((void **)&dc_out, sizeof(cufftDoubleComplex) * NX * NY * NE));
        dc_host = (double *)calloc((size_t)(steps > 0 ? steps : 1), sizeof(double));
    }
    double lbm_ms = 0.0, poi_ms = 0.0;
    float loop_ms = 0.f;
    for (int phase = 0; phase < 2; ++phase) {
        const int n = phase == 0 ? warmup : steps;
        if (phase == 1) { checkCudaErrors(cudaDeviceSynchronize()); checkCudaErrors(cudaEventRecord(e0, 0)); }
        for (int i = 0; i < n; ++i) {
            cudaEvent_t a, b, c;
            if (split && phase == 1) {
                cudaEventCreate(&a); cudaEventCreate(&b); cudaEventCreate(&c);
                cudaEventRecord(a, 0);
            }
            stream_collide_save(f0_gpu, f1_gpu, f2_gpu, h0_gpu, h1_gpu, h2_gpu, hn0_gpu, hn1_gpu, hn2_gpu,
                                temp0_gpu, temp1_gpu, temp2_gpu, rho_gpu, charge_gpu, chargen_gpu,
                                ux_gpu, uy_gpu, uz_gpu, Ex_gpu, Ey_gpu, Ez_gpu, T_gpu, t, f0bc);
            if (split && phase == 1) cudaEventRecord(b, 0);
            if (dump_dc && phase == 1) {
                cufftDoubleComplex first;
                extension(charge_gpu, chargen_gpu, dc_in);                       /* poisson.cu:83 */
                CHECK_CUFFT(cufftExecZ2Z(plan, dc_in, dc_out, CUFFT_FORWARD));   /* poisson.cu:86 */
                checkCudaErrors(cudaMemcpy(&first, dc_out, sizeof(first), cudaMemcpyDeviceToHost));
                dc_host[i] = first.x;
            }
            fast_Poisson(charge_gpu, chargen_gpu, kx, ky, kz, plan);
            if (split && phase == 1) {
                cudaEventRecord(c, 0);
                cudaEventSynchronize(c);
                float x1 = 0, x2 = 0;
                cudaEventElapsedTime(&x1, a, b); cudaEventElapsedTime(&x2, b, c);
                lbm_ms += x1; poi_ms += x2;
                cudaEventDestroy(a); cudaEventDestroy(b); cudaEventDestroy(c);
            }
            t = t + dt_host;
        }
        if (phase == 1) {
            checkCudaErrors(cudaEventRecord(e1, 0));
            checkCudaErrors(cudaEventSynchronize(e1));
            checkCudaErrors(cudaEventElapsedTime(&loop_ms, e0, e1));
        }
    }

    if (dump_dc) {
        FILE *f = fopen(dump_dc, "wb");
        fwrite(dc_host, sizeof(double), steps, f);
        fclose(f);
    }
    if (dump_final) dump_fields(dump_final);
    if (dump_p) dump_pops(dump_p);

    const double cells = (double)NX * NY * NZ;
    const double mlups = steps > 0 ? cells * steps / (1e3 * loop_ms) : 0.0;
    if (!quiet_run || 1)
        printf("{\"impl\": \"reference-cuda\", \"NX\": %u, \"NY\": %u, \"NZ\": %u, \"nThreads\": %d, "
               "\"steps\": %d, \"warmup\": %d, \"loop_ms\": %.6f, \"ms_per_step\": %.6f, \"mlups\": %.4f, "
               "\"lbm_ms_per_step\": %.6f, \"poisson_ms_per_step\": %.6f, \"init_ms\": %.3f}\n",
               NX, NY, NZ, nThreads, steps, warmup, loop_ms, steps > 0 ? loop_ms / steps : 0.0, mlups,
               (split && steps > 0) ? lbm_ms / steps : 0.0, (split && steps > 0) ? poi_ms / steps : 0.0, init_ms);
    return 0;
}
