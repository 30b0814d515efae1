/*
 * ek_oracle.c -- CPU restatement of EK-PNP-3D's coupled time step (fp64).
 *
 * TEST INFRASTRUCTURE ONLY (see ek_oracle.h).  It follows the reference's
 * four-kernel, two-lattice sequence literally -- collide, boundary, stream,
 * ion/temperature wall pass, then the odd-extension FFT Poisson solve -- with
 * the reference's evaluation order for every sum whose rounding matters.
 * Build with -ffp-contract=off so that no FMA contraction is introduced.
 *
 * All file:line citations are relative to /root/reference.
 */
#include "ek_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* D3Q27 velocity set, machine-checked in SURVEY.md A.2 against
 * LBM.cu:872-1103 (equilibria), :639-644 (momentum), :1983-2092 (stream). */
static const int CX[27] = {0, 1,-1, 0, 0, 0, 0, 1,-1, 1,-1, 0, 0, 1,-1, 1,-1, 0, 0, 1,-1, 1,-1, 1,-1,-1, 1};
static const int CY[27] = {0, 0, 0, 1,-1, 0, 0, 1,-1, 0, 0, 1,-1,-1, 1, 0, 0, 1,-1, 1,-1, 1,-1,-1, 1, 1,-1};
static const int CZ[27] = {0, 0, 0, 0, 0, 1,-1, 0, 0, 1,-1, 1,-1, 0, 0,-1, 1,-1, 1, 1,-1,-1, 1, 1,-1, 1,-1};

struct eko_state {
    eko_params p;
    size_t N;          /* NX*NY*NZ */
    int NE;            /* 2*(NZ-1), LBM.h:37 */
    double *x0[4];     /* rest populations f0,h0,hn0,temp0   [N]   */
    double *x1[4];     /* f1,h1,hn1,temp1                    [26N] */
    double *x2[4];     /* f2,h2,hn2,temp2                    [26N] */
    double *fld[EKO_NFIELDS];
    double *f0bc;      /* [2][NY][NX], LBM.cu:502-504 */
    double *kx, *ky, *kz;
    double *ext;       /* complex scratch, interleaved, NX*NY*NE */
    double *phi_old;
    int dc_mode;       /* 0 zero, 1 literal (reference, default), 2 prescribed */
    double dc_ghat0;
    double last_dc;    /* forward DC coefficient seen by the last solve */
};

/* ---- index helpers: LBM.cu:17-30 ---- */
static inline size_t sidx(const eko_params *p, int x, int y, int z)
{
    return (size_t)p->NX * ((size_t)p->NY * z + y) + x;
}
static inline size_t nidx(const eko_params *p, int x, int y, int z, int d)
{
    return (size_t)p->NX * ((size_t)p->NY * ((size_t)p->NZ * (d - 1) + z) + y) + x;
}

int eko_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* LBM.h:29-125 as shipped */
void eko_default_params(eko_params *p)
{
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

/* main.cu:78-152: allocations and wavenumber tables */
eko_state *eko_create(const eko_params *p)
{
    eko_state *s = (eko_state *)calloc(1, sizeof(*s));
    s->p = *p;
    s->N = (size_t)p->NX * p->NY * p->NZ;
    s->NE = 2 * (p->NZ - 1);
    for (int k = 0; k < 4; ++k) {
        s->x0[k] = (double *)calloc(s->N, sizeof(double));
        s->x1[k] = (double *)calloc(s->N * 26, sizeof(double));
        s->x2[k] = (double *)calloc(s->N * 26, sizeof(double));
    }
    for (int k = 0; k < EKO_NFIELDS; ++k) s->fld[k] = (double *)calloc(s->N, sizeof(double));
    s->f0bc = (double *)calloc((size_t)2 * p->NX * p->NY, sizeof(double));
    s->kx = (double *)malloc(sizeof(double) * p->NX);
    s->ky = (double *)malloc(sizeof(double) * p->NY);
    s->kz = (double *)malloc(sizeof(double) * s->NE);
    s->ext = (double *)malloc(sizeof(double) * 2 * (size_t)p->NX * p->NY * s->NE);
    s->phi_old = (double *)calloc(s->N, sizeof(double));
    s->dc_mode = 1;
    /* main.cu:119-145 (unsigned loop indices; "(double)i - NX") */
    const int NX = p->NX, NY = p->NY, NE = s->NE;
    for (int i = 0; i <= NX / 2; i++) s->kx[i] = (double)i * 2.0 * M_PI / p->Lx;
    for (int i = NX / 2 + 1; i < NX; i++) s->kx[i] = ((double)i - NX) * 2.0 * M_PI / p->Lx;
    for (int i = 0; i <= NY / 2; i++) s->ky[i] = (double)i * 2.0 * M_PI / p->Ly;
    for (int i = NY / 2 + 1; i < NY; i++) s->ky[i] = ((double)i - NY) * 2.0 * M_PI / p->Ly;
    for (int i = 0; i <= NE / 2; i++) s->kz[i] = (double)i * 2.0 * M_PI / (NE * p->dz);
    for (int i = NE / 2 + 1; i < NE; i++) s->kz[i] = ((double)i - NE) * 2.0 * M_PI / (NE * p->dz);
    return s;
}

void eko_destroy(eko_state *s)
{
    if (!s) return;
    for (int k = 0; k < 4; ++k) { free(s->x0[k]); free(s->x1[k]); free(s->x2[k]); }
    for (int k = 0; k < EKO_NFIELDS; ++k) free(s->fld[k]);
    free(s->f0bc); free(s->kx); free(s->ky); free(s->kz); free(s->ext); free(s->phi_old);
    free(s);
}

double *eko_field(eko_state *s, int id) { return s->fld[id]; }

void eko_set_poisson_dc(eko_state *s, int mode, double ghat0) { s->dc_mode = mode; s->dc_ghat0 = ghat0; }
double eko_last_dc(eko_state *s) { return s->last_dc; }

void eko_get_populations(eko_state *s, int set, double *out)
{
    memcpy(out, s->x0[set], s->N * sizeof(double));
    memcpy(out + s->N, s->x1[set], s->N * 26 * sizeof(double));
}

/* ------------------------------------------------------------------ */
/* equilibrium direction term "cidot3u", literal forms of LBM.cu:872-1103
 * (identical at LBM.cu:230-462).  Two-term forms are rounding-equivalent
 * under commutation/negation; the three-term forms are not, so each is
 * spelled out as in the reference. */
static inline double cidot(int d, double tx, double ty, double tz)
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

static inline double wclass(const eko_params *p, int d)
{
    if (d == 0) return p->w0;
    if (d <= 6) return p->ws;
    if (d <= 18) return p->wa;
    return p->wd;
}

/* equilibrium set: eq[d] = w_d*m*(omusq + s*(1+0.5*s)), LBM.cu:830-1103 */
static void equilibrium(const eko_params *p, double m, double vx, double vy, double vz, double eq[27])
{
    const double w0m = p->w0 * m, wsm = p->ws * m, wam = p->wa * m, wdm = p->wd * m;
    const double omusq = 1.0 - 0.5 * (vx * vx + vy * vy + vz * vz) / p->cs_square;
    const double tx = vx / p->cs_square / p->CFL;
    const double ty = vy / p->cs_square / p->CFL;
    const double tz = vz / p->cs_square / p->CFL;
    eq[0] = w0m * (omusq);
    for (int d = 1; d < 27; ++d) {
        const double c = cidot(d, tx, ty, tz);
        const double wm = d <= 6 ? wsm : (d <= 18 ? wam : wdm);
        eq[d] = wm * (omusq + c * (1.0 + 0.5 * c));
    }
}

/* sums of LBM.cu:621-630 (left to right) */
static inline double sum27(const double *v)
{
    double a = v[0];
    for (int d = 1; d < 27; ++d) a = a + v[d];
    return a;
}

/* momentum brackets of LBM.cu:639-644 (and :795-800) */
static inline void momentum(const double *f, double m[3])
{
    m[0] = (f[1] + f[7] + f[9] + f[13] + f[15] + f[19] + f[21] + f[23] + f[26]
          - (f[2] + f[8] + f[10] + f[14] + f[16] + f[20] + f[22] + f[24] + f[25]));
    m[1] = (f[3] + f[7] + f[11] + f[14] + f[17] + f[19] + f[21] + f[24] + f[25]
          - (f[4] + f[8] + f[12] + f[13] + f[18] + f[20] + f[22] + f[23] + f[26]));
    m[2] = (f[5] + f[9] + f[11] + f[16] + f[18] + f[19] + f[22] + f[23] + f[25]
          - (f[6] + f[10] + f[12] + f[15] + f[17] + f[20] + f[21] + f[24] + f[26]));
}

static void load_node(const eko_state *s, int set, int x, int y, int z, double v[27])
{
    const eko_params *p = &s->p;
    v[0] = s->x0[set][sidx(p, x, y, z)];
    for (int d = 1; d < 27; ++d) v[d] = s->x1[set][nidx(p, x, y, z, d)];
}

/* Guo force populations, LBM.cu:1107-1145.  For an axis k with c_k != 0 the
 * term is ((c_k*cflinv - u_k) + (c_k*(c.u))*cflinv2)*F_k, for c_k == 0 it is
 * -u_k*F_k; q = c_a*(c.u) is evaluated left to right in x,y,z order starting
 * from the first non-zero axis (the other axes use +q or -q, exact), and the
 * terms are added in the reference's order (see comments per class). */
static void guo_force(const eko_params *p, const double u[3], const double F[3], double fpop[27])
{
    const double coe0 = p->w0 / p->cs_square, coes = p->ws / p->cs_square;
    const double coea = p->wa / p->cs_square, coed = p->wd / p->cs_square;
    const double cflinv = 1.0 / p->CFL;
    const double cflinv2 = cflinv * cflinv / p->cs_square;
    fpop[0] = -coe0 * (u[0] * F[0] + u[1] * F[1] + u[2] * F[2]);
    for (int d = 1; d < 27; ++d) {
        const int c[3] = {CX[d], CY[d], CZ[d]};
        int nz[3], nnz = 0, zr[3], nzr = 0;
        for (int k = 0; k < 3; ++k) { if (c[k]) nz[nnz++] = k; else zr[nzr++] = k; }
        if (nnz == 1) {
            /* LBM.cu:1117-1122: coes*(-u_b*F_b - u_c*F_c + ((+-cflinv - u_a) + (cflinv2*u_a))*F_a) */
            const int a = nz[0], b = zr[0], cc = zr[1];
            const double Ta = ((c[a] * cflinv - u[a]) + (cflinv2 * u[a])) * F[a];
            fpop[d] = coes * (-u[b] * F[b] - u[cc] * F[cc] + Ta);
        } else if (nnz == 2) {
            /* LBM.cu:1124-1136: coea*(T_a + T_b - u_c*F_c) */
            const int a = nz[0], b = nz[1], cc = zr[0];
            const double q = (c[a] * c[b] > 0) ? (u[a] + u[b]) : (u[a] - u[b]);
            const double qb = (c[a] * c[b] > 0) ? q : -q;
            const double Ta = ((c[a] * cflinv - u[a]) + q * cflinv2) * F[a];
            const double Tb = ((c[b] * cflinv - u[b]) + qb * cflinv2) * F[b];
            fpop[d] = coea * (Ta + Tb - u[cc] * F[cc]);
        } else {
            /* LBM.cu:1138-1145: coed*(T_x + T_y + T_z) */
            double q = u[0];
            q = (c[0] * c[1] > 0) ? q + u[1] : q - u[1];
            q = (c[0] * c[2] > 0) ? q + u[2] : q - u[2];
            const double qy = (c[0] * c[1] > 0) ? q : -q;
            const double qz = (c[0] * c[2] > 0) ? q : -q;
            const double Tx = ((c[0] * cflinv - u[0]) + q * cflinv2) * F[0];
            const double Ty = ((c[1] * cflinv - u[1]) + qy * cflinv2) * F[1];
            const double Tz = ((c[2] * cflinv - u[2]) + qz * cflinv2) * F[2];
            fpop[d] = coed * (Tx + Ty + Tz);
        }
    }
}

/* TRT relaxation of one set at one node, LBM.cu:1148-1845.
 * post[d] = X_d - (wp*(X+_d - E+_d) + wm*(X-_d - E-_d)) [+ dt*source_d] */
static void trt(const double X[27], const double E[27], double wp, double wm,
                const double *source, double dt, double post[27])
{
    {
        const double xp = X[0], xm = 0.0, ep = E[0], em = 0.0;
        post[0] = X[0] - (wp * (xp - ep) + wm * (xm - em));
        if (source) post[0] = post[0] + dt * source[0];
    }
    for (int d = 1; d < 27; d += 2) {
        const double xp = 0.5 * (X[d] + X[d + 1]);
        const double xm = 0.5 * (X[d] - X[d + 1]);
        const double ep = 0.5 * (E[d] + E[d + 1]);
        const double em = 0.5 * (E[d] - E[d + 1]);
        post[d]     = X[d]     - (wp * (xp - ep) + wm * (xm - em));
        post[d + 1] = X[d + 1] - (wp * (xp - ep) + wm * ((-xm) - (-em)));
        if (source) {
            post[d]     = post[d]     + dt * source[d];
            post[d + 1] = post[d + 1] + dt * source[d + 1];
        }
    }
}

/* gpu_collide_save, LBM.cu:483-1846 */
static void collide_save(eko_state *s)
{
    const eko_params *p = &s->p;
    const int NX = p->NX, NY = p->NY, NZ = p->NZ;
    /* LBM.cu:488-495 */
    const double dt = p->dt, cs2 = p->cs_square;
    const double omega_plus = 1.0 / (p->nu / cs2 / dt + 1.0 / 2.0) / dt;
    const double omega_minus = 1.0 / (p->V / (p->nu / cs2 / dt) + 1.0 / 2.0) / dt;
    const double omega_c_minus = 1.0 / (p->diffu / cs2 / dt + 1.0 / 2.0) / dt;
    const double omega_c_plus = 1.0 / (p->VC / (p->diffu / cs2 / dt) + 1.0 / 2.0) / dt;
    const double omega_cn_minus = 1.0 / (p->diffun / cs2 / dt + 1.0 / 2.0) / dt;
    const double omega_cn_plus = 1.0 / (p->VCn / (p->diffun / cs2 / dt) + 1.0 / 2.0) / dt;
    const double omega_T_minus = 1.0 / (p->D / cs2 / dt + 1.0 / 2.0) / dt;
    const double omega_T_plus = 1.0 / (p->VT / (p->D / cs2 / dt) + 1.0 / 2.0) / dt;
    /* LBM.cu:1700-1707 */
    const double tw0rp = omega_plus * dt, tw0rm = omega_minus * dt;
    const double tw0cp = omega_c_plus * dt, tw0cm = omega_c_minus * dt;
    const double tw0cnp = omega_cn_plus * dt, tw0cnm = omega_cn_minus * dt;
    const double tw0Tp = omega_T_plus * dt, tw0Tm = omega_T_minus * dt;
    /* LBM.cu:1660-1661 */
    const double sp = 1.0 - 0.5 * dt * omega_plus;
    const double sm = 1.0 - 0.5 * dt * omega_minus;

    double *r = s->fld[EKO_RHO], *u = s->fld[EKO_UX], *v = s->fld[EKO_UY], *w = s->fld[EKO_UZ];
    double *c = s->fld[EKO_CHARGE], *cn = s->fld[EKO_CHARGEN], *Temperature = s->fld[EKO_T];
    const double *ex = s->fld[EKO_EX], *ey = s->fld[EKO_EY], *ez = s->fld[EKO_EZ];

    /* The z = 0 nodes read the rest populations of the z = 1 nodes, which the
     * z = 1 nodes overwrite in place at the end of the same kernel
     * (LBM.cu:664-667 vs :1711-1714).  The de-facto semantics (blocks are
     * dispatched z = 0 first) is "pre-collision values"; snapshot them. */
    const size_t plane = (size_t)NX * NY;
    double *rest1[4];
    for (int k = 0; k < 4; ++k) {
        rest1[k] = (double *)malloc(plane * sizeof(double));
        memcpy(rest1[k], s->x0[k] + plane, plane * sizeof(double));
    }

#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < NZ; ++z)
    for (int y = 0; y < NY; ++y)
    for (int x = 0; x < NX; ++x) {
        const size_t si = sidx(p, x, y, z);
        /* LBM.cu:502-504 */
        if (z == 0) s->f0bc[(size_t)NX * (NY * 0 + y) + x] = s->x0[0][si];
        if (z == NZ - 1) s->f0bc[(size_t)NX * (NY * 1 + y) + x] = s->x0[0][si];

        double ft[27], ht[27], hnt[27], tt[27];
        load_node(s, 0, x, y, z, ft);
        load_node(s, 1, x, y, z, ht);
        load_node(s, 2, x, y, z, hnt);
        load_node(s, 3, x, y, z, tt);

        /* LBM.cu:621-637 */
        const double rho = sum27(ft);
        const double rhoinv = 1.0 / rho;
        const double charge = sum27(ht);
        const double chargen = sum27(hnt);
        const double temp = sum27(tt);
        const double Ex = ex[si], Ey = ey[si], Ez = ez[si];
        double F[3];
        F[0] = p->convertCtoCharge * (charge - chargen) * (Ex + p->Ext) + p->exf;
        F[1] = p->convertCtoCharge * (charge - chargen) * Ey;
        F[2] = p->convertCtoCharge * (charge - chargen) * Ez + p->rho0 * temp * p->Ra * p->nu * p->D;

        /* LBM.cu:639-644 */
        double m[3], uu[3];
        momentum(ft, m);
        uu[0] = rhoinv * (m[0] / p->CFL + F[0] * dt * 0.5);
        uu[1] = rhoinv * (m[1] / p->CFL + F[1] * dt * 0.5);
        uu[2] = rhoinv * (m[2] / p->CFL + F[2] * dt * 0.5);

        /* LBM.cu:663-801: bottom wall takes minus the z = 1 momentum, divided
         * by the wall node's own density (rhoinvm = 1.0/rho, :780). */
        if (z == 0) {
            double fm_[27], hm_[27], hnm_[27], tm_[27];
            load_node(s, 0, x, y, 1, fm_);
            load_node(s, 1, x, y, 1, hm_);
            load_node(s, 2, x, y, 1, hnm_);
            load_node(s, 3, x, y, 1, tm_);
            const size_t pi = (size_t)NX * y + x;
            fm_[0] = rest1[0][pi]; hm_[0] = rest1[1][pi]; hnm_[0] = rest1[2][pi]; tm_[0] = rest1[3][pi];
            const double rhoinvm = 1.0 / rho;
            const double chargem = sum27(hm_);
            const double chargenm = sum27(hnm_);
            const double tempm = sum27(tm_);
            const size_t s1 = sidx(p, x, y, 1);
            const double Exm = ex[s1], Eym = ey[s1], Ezm = ez[s1];
            const double Fxm = p->convertCtoCharge * (chargem - chargenm) * (Exm + p->Ext) + p->exf;
            const double Fym = p->convertCtoCharge * (chargem - chargenm) * Eym;
            const double Fzm = p->convertCtoCharge * (chargem - chargenm) * Ezm + p->rho0 * tempm * p->Ra * p->nu * p->D;
            double mm[3];
            momentum(fm_, mm);
            uu[0] = -rhoinvm * (mm[0] / p->CFL + Fxm * dt * 0.5);
            uu[1] = -rhoinvm * (mm[1] / p->CFL + Fym * dt * 0.5);
            uu[2] = -rhoinvm * (mm[2] / p->CFL + Fzm * dt * 0.5);
        }

        /* LBM.cu:807-813 */
        r[si] = rho; u[si] = uu[0]; v[si] = uu[1]; w[si] = uu[2];
        c[si] = charge; cn[si] = chargen; Temperature[si] = temp;

        /* LBM.cu:850-1103 */
        double fe[27], he[27], hne[27], te[27];
        equilibrium(p, rho, uu[0], uu[1], uu[2], fe);
        equilibrium(p, charge, uu[0] + p->K * Ex, uu[1] + p->K * Ey, uu[2] + p->K * Ez, he);
        equilibrium(p, chargen, uu[0] + p->Kn * Ex, uu[1] + p->Kn * Ey, uu[2] + p->Kn * Ez, hne);
        equilibrium(p, temp, uu[0], uu[1], uu[2], te);

        /* LBM.cu:1107-1145, 1608-1689 */
        double fpop[27], source[27];
        guo_force(p, uu, F, fpop);
        source[0] = sp * fpop[0];
        for (int d = 1; d < 27; d += 2) {
            const double fp_ = 0.5 * (fpop[d] + fpop[d + 1]);
            const double fm = 0.5 * (fpop[d] - fpop[d + 1]);
            source[d] = sp * fp_ + sm * fm;
            source[d + 1] = sp * fp_ + sm * (-fm);
        }

        /* LBM.cu:1711-1845 */
        double fo[27], ho[27], hno[27], to[27];
        trt(ft, fe, tw0rp, tw0rm, source, dt, fo);
        trt(ht, he, tw0cp, tw0cm, NULL, dt, ho);
        trt(hnt, hne, tw0cnp, tw0cnm, NULL, dt, hno);
        trt(tt, te, tw0Tp, tw0Tm, NULL, dt, to);
        s->x0[0][si] = fo[0]; s->x0[1][si] = ho[0]; s->x0[2][si] = hno[0]; s->x0[3][si] = to[0];
        for (int d = 1; d < 27; ++d) {
            const size_t ni = nidx(p, x, y, z, d);
            s->x2[0][ni] = fo[d]; s->x2[1][ni] = ho[d]; s->x2[2][ni] = hno[d]; s->x2[3][ni] = to[d];
        }
    }
    for (int k = 0; k < 4; ++k) free(rest1[k]);
}

/* gpu_boundary, LBM.cu:1848-1961 */
static void boundary(eko_state *s)
{
    const eko_params *p = &s->p;
    const int NX = p->NX, NY = p->NY, NZ = p->NZ;
    double *f0 = s->x0[0], *f1 = s->x1[0], *f2 = s->x2[0];
    const double multis = 2.0 * p->rho0 * p->uw / p->cs_square * p->ws / p->CFL;
    const double multia = 2.0 * p->rho0 * p->uw / p->cs_square * p->wa / p->CFL;
    const double multid = 2.0 * p->rho0 * p->uw / p->cs_square * p->wd / p->CFL;
    /* sign table of LBM.cu:1902-1927: +1, -1 or 0 times the class value */
    static const int sgn[27] = {0, +1, -1, +1, 0, 0, 0, +1, -1, +1, -1, 0, 0, +1, -1, +1, -1, 0, 0,
                                +1, -1, +1, -1, +1, -1, -1, +1};
#pragma omp parallel for schedule(static)
    for (int y = 0; y < NY; ++y)
    for (int x = 0; x < NX; ++x) {
        f0[sidx(p, x, y, 0)] = s->f0bc[(size_t)NX * (NY * 0 + y) + x];
        for (int d = 1; d < 27; ++d) {
            const int o = (d & 1) ? d + 1 : d - 1;
            f2[nidx(p, x, y, 0, d)] = f1[nidx(p, x, y, 0, o)];
        }
        f0[sidx(p, x, y, NZ - 1)] = s->f0bc[(size_t)NX * (NY * 1 + y) + x];
        for (int d = 1; d < 27; ++d) {
            const int o = (d & 1) ? d + 1 : d - 1;
            const double mu = d <= 6 ? multis : (d <= 18 ? multia : multid);
            double val = f1[nidx(p, x, y, NZ - 1, o)];
            if (sgn[d] > 0) val = val + mu;
            else if (sgn[d] < 0) val = val - mu;
            f2[nidx(p, x, y, NZ - 1, d)] = val;
        }
    }
}

/* gpu_stream, LBM.cu:1963-2093: X1[x,d] = X2[x - c_d (mod NX,NY,NZ), d] */
static void stream(eko_state *s)
{
    const eko_params *p = &s->p;
    const int NX = p->NX, NY = p->NY, NZ = p->NZ;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < 4; ++k)
    for (int d = 1; d < 27; ++d) {
        const double *src = s->x2[k];
        double *dst = s->x1[k];
        for (int z = 0; z < NZ; ++z) {
            const int zs = (z - CZ[d] + NZ) % NZ;
            for (int y = 0; y < NY; ++y) {
                const int ys = (y - CY[d] + NY) % NY;
                for (int x = 0; x < NX; ++x) {
                    const int xs = (x - CX[d] + NX) % NX;
                    dst[nidx(p, x, y, z, d)] = src[nidx(p, xs, ys, zs, d)];
                }
            }
        }
    }
}

/* gpu_bc_charge, LBM.cu:2095-2416 */
static void bc_charge(eko_state *s)
{
    const eko_params *p = &s->p;
    const int NX = p->NX, NY = p->NY, NZ = p->NZ;
    const double multi0T = 2.0 * p->TH * p->w0, multisT = 2.0 * p->TH * p->ws;
    const double multiaT = 2.0 * p->TH * p->wa, multidT = 2.0 * p->TH * p->wd;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < NY; ++y)
    for (int x = 0; x < NX; ++x) {
        for (int wz = 0; wz < 2; ++wz) {
            const int z = wz ? NZ - 1 : 0;
            /* ions: LBM.cu:2102-2218 */
            for (int k = 1; k <= 2; ++k)
                for (int d = 1; d < 27; ++d) {
                    const int o = (d & 1) ? d + 1 : d - 1;
                    s->x1[k][nidx(p, x, y, z, d)] = s->x2[k][nidx(p, x, y, z, o)];
                }
            /* temperature: bottom LBM.cu:2220-2350, top LBM.cu:2354-2413 */
            const size_t si = sidx(p, x, y, z);
            if (z == 0) {
                s->x0[3][si] = -s->x0[3][si] + multi0T;
                for (int d = 1; d < 27; ++d) {
                    const int o = (d & 1) ? d + 1 : d - 1;
                    const double mu = d <= 6 ? multisT : (d <= 18 ? multiaT : multidT);
                    s->x1[3][nidx(p, x, y, z, d)] = -s->x2[3][nidx(p, x, y, z, o)] + mu;
                }
            } else {
                s->x0[3][si] = -s->x0[3][si];
                for (int d = 1; d < 27; ++d) {
                    const int o = (d & 1) ? d + 1 : d - 1;
                    s->x1[3][nidx(p, x, y, z, d)] = -s->x2[3][nidx(p, x, y, z, o)];
                }
            }
        }
    }
}

/* stream_collide_save, LBM.cu:465-481 */
void eko_stream_collide_save(eko_state *s)
{
    collide_save(s);
    boundary(s);
    stream(s);
    bc_charge(s);
}

/* ------------------------------------------------------------------ */
/* Complex DFT of arbitrary length (what cufftExecZ2Z computes,
 * poisson.cu:86,92: unnormalised, sign -1 forward / +1 inverse).
 * Recursive mixed-radix decimation in time; prime factors by direct sums. */
typedef struct { double re, im; } cplx;

static void fft_rec(const cplx *in, cplx *out, int n, int istride,
                    const cplx *tw, int ntop)
{
    if (n == 1) { out[0] = in[0]; return; }
    int p = 2;
    while (n % p) ++p;
    const int m = n / p;
    for (int r = 0; r < p; ++r) fft_rec(in + (size_t)r * istride, out + (size_t)r * m, m, istride * p, tw, ntop);
    const int tws = ntop / n;
    cplx tmp[64];
    cplx *t = p <= 64 ? tmp : (cplx *)malloc(sizeof(cplx) * p);
    for (int k = 0; k < m; ++k) {
        for (int r = 0; r < p; ++r) t[r] = out[(size_t)r * m + k];
        for (int q = 0; q < p; ++q) {
            const int kk = k + q * m;
            double are = t[0].re, aim = t[0].im;
            for (int r = 1; r < p; ++r) {
                const cplx wv = tw[(size_t)(((long long)r * kk) % n) * tws];
                are += t[r].re * wv.re - t[r].im * wv.im;
                aim += t[r].re * wv.im + t[r].im * wv.re;
            }
            out[kk].re = are; out[kk].im = aim;
        }
    }
    if (t != tmp) free(t);
}

static cplx *make_twiddles(int n, int sign)
{
    cplx *tw = (cplx *)malloc(sizeof(cplx) * n);
    for (int k = 0; k < n; ++k) {
        const double a = 2.0 * M_PI * (double)k / (double)n;
        tw[k].re = cos(a);
        tw[k].im = sign * sin(a);
    }
    return tw;
}

void eko_fft1d(double *data, int n, int stride, int sign)
{
    cplx *tw = make_twiddles(n, sign);
    cplx *in = (cplx *)malloc(sizeof(cplx) * n), *out = (cplx *)malloc(sizeof(cplx) * n);
    for (int i = 0; i < n; ++i) { in[i].re = data[2 * (size_t)i * stride]; in[i].im = data[2 * (size_t)i * stride + 1]; }
    fft_rec(in, out, n, 1, tw, n);
    for (int i = 0; i < n; ++i) { data[2 * (size_t)i * stride] = out[i].re; data[2 * (size_t)i * stride + 1] = out[i].im; }
    free(in); free(out); free(tw);
}

/* 3-D transform of an [n2][n1][n0] interleaved-complex array (n0 fastest) */
static void fft3d(double *a, int n0, int n1, int n2, int sign)
{
    const int dims[3] = {n0, n1, n2};
    const size_t strides[3] = {1, (size_t)n0, (size_t)n0 * n1};
    for (int ax = 0; ax < 3; ++ax) {
        const int n = dims[ax];
        const size_t st = strides[ax];
        cplx *tw = make_twiddles(n, sign);
        const int oa = (ax + 1) % 3, ob = (ax + 2) % 3;
        const long long nlines = (long long)dims[oa] * dims[ob];
#pragma omp parallel
        {
            cplx *in = (cplx *)malloc(sizeof(cplx) * n), *out = (cplx *)malloc(sizeof(cplx) * n);
#pragma omp for schedule(static)
            for (long long l = 0; l < nlines; ++l) {
                const size_t ia = (size_t)(l % dims[oa]), ib = (size_t)(l / dims[oa]);
                cplx *base = (cplx *)a + ia * strides[oa] + ib * strides[ob];
                for (int i = 0; i < n; ++i) in[i] = base[(size_t)i * st];
                fft_rec(in, out, n, 1, tw, n);
                for (int i = 0; i < n; ++i) base[(size_t)i * st] = out[i];
            }
            free(in); free(out);
        }
        free(tw);
    }
}

/* fast_Poisson, poisson.cu:75-103 (scratch is persistent here; the
 * per-call cudaMalloc/cudaFree has no numerical effect) */
void eko_fast_poisson(eko_state *s)
{
    const eko_params *p = &s->p;
    const int NX = p->NX, NY = p->NY, NZ = p->NZ, NE = s->NE;
    const double *charge = s->fld[EKO_CHARGE], *chargen = s->fld[EKO_CHARGEN];
    cplx *ext = (cplx *)s->ext;
    const double eps = p->eps, CtoC = p->convertCtoCharge, dz = p->dz;

    /* odd_extension, poisson.cu:114-158 */
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < NE; ++z)
    for (int y = 0; y < NY; ++y)
    for (int x = 0; x < NX; ++x) {
        const size_t i = sidx(p, x, y, z);
        double v = 0.0;
        if (z == 0) v = 0.0;
        else if (z == 1) v = -CtoC * (charge[sidx(p, x, y, z)] - chargen[sidx(p, x, y, z)]) / eps - p->voltage / dz / dz;
        else if (z > 1 && z < NZ - 2) v = -CtoC * (charge[sidx(p, x, y, z)] - chargen[sidx(p, x, y, z)]) / eps;
        else if (z == NZ - 2) v = -CtoC * (charge[sidx(p, x, y, z)] - chargen[sidx(p, x, y, z)]) / eps - p->voltage2 / dz / dz;
        else if (z == NZ - 1) v = 0.0;
        else if (z == NZ) v = CtoC * (charge[sidx(p, x, y, NE - z)] - chargen[sidx(p, x, y, NE - z)]) / eps + p->voltage2 / dz / dz;
        else if (z > NZ && z < NE - 1) v = CtoC * (charge[sidx(p, x, y, NE - z)] - chargen[sidx(p, x, y, NE - z)]) / eps;
        else if (z == NE - 1) v = CtoC * (charge[sidx(p, x, y, 1)] - chargen[sidx(p, x, y, 1)]) / eps + p->voltage / dz / dz;
        ext[i].re = v; ext[i].im = 0.0;
    }

    /* cufftExecZ2Z forward, poisson.cu:86; plan main.cu:112 (NE x NY x NX) */
    fft3d(s->ext, NX, NY, NE, -1);

    /* The (0,0,0) coefficient is zero by oddness in exact arithmetic; what is
     * left is the transform's rounding residue, which the reference divides
     * by mu := 1 (poisson.cu:176-177).  dc_mode 1 keeps that literal
     * behaviour (with THIS transform's residue), 0 enforces zero, 2 replays a
     * residue recorded from the reference's cuFFT run. */
    s->last_dc = ext[0].re;
    if (s->dc_mode == 0) { ext[0].re = 0.0; ext[0].im = 0.0; }
    else if (s->dc_mode == 2) { ext[0].re = s->dc_ghat0; ext[0].im = 0.0; }

    /* gpu_derivative, poisson.cu:169-180 */
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < NE; ++z)
    for (int y = 0; y < NY; ++y)
    for (int x = 0; x < NX; ++x) {
        const double I = s->kx[x], J = s->ky[y], Kz = s->kz[z];
        double mu = (4.0 / dz / dz) * (sin(Kz * dz * 0.5) * sin(Kz * dz * 0.5)) + I * I + J * J;
        if (y == 0 && x == 0 && z == 0) mu = 1.0;
        const size_t i = sidx(p, x, y, z);
        ext[i].re = -ext[i].re / mu;
        ext[i].im = -ext[i].im / mu;
    }

    /* cufftExecZ2Z inverse, poisson.cu:92 */
    fft3d(s->ext, NX, NY, NE, +1);

    /* odd_extract, poisson.cu:191-204; size = NX*NY*NE (LBM.h:38) */
    double *phi = s->fld[EKO_PHI];
    const double size = (double)((unsigned int)NX * (unsigned int)NY * (unsigned int)NE);
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < NZ; ++z)
    for (int y = 0; y < NY; ++y)
    for (int x = 0; x < NX; ++x) {
        const size_t i = sidx(p, x, y, z);
        if (z == 0) phi[i] = p->voltage;
        else if (z == NZ - 1) phi[i] = p->voltage2;
        else phi[i] = ext[i].re / size;
    }

    /* gpu_efield, poisson.cu:40-56 (periodic wrap in all three axes) */
    double *Ex = s->fld[EKO_EX], *Ey = s->fld[EKO_EY], *Ez = s->fld[EKO_EZ];
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < NZ; ++z)
    for (int y = 0; y < NY; ++y)
    for (int x = 0; x < NX; ++x) {
        const int xp1 = (x + 1) % NX, yp1 = (y + 1) % NY, zp1 = (z + 1) % NZ;
        const int xm1 = (NX + x - 1) % NX, ym1 = (NY + y - 1) % NY, zm1 = (NZ + z - 1) % NZ;
        const size_t i = sidx(p, x, y, z);
        Ex[i] = 0.5 * (phi[sidx(p, xm1, y, z)] - phi[sidx(p, xp1, y, z)]) / p->dx;
        Ey[i] = 0.5 * (phi[sidx(p, x, ym1, z)] - phi[sidx(p, x, yp1, z)]) / p->dy;
        Ez[i] = 0.5 * (phi[sidx(p, x, y, zm1)] - phi[sidx(p, x, y, zp1)]) / p->dz;
    }
    /* gpu_bc, poisson.cu:57-69 */
    for (int y = 0; y < NY; ++y)
    for (int x = 0; x < NX; ++x) {
        Ez[sidx(p, x, y, 0)] = Ez[sidx(p, x, y, 1)];
        Ez[sidx(p, x, y, NZ - 1)] = Ez[sidx(p, x, y, NZ - 2)];
    }
}

/* initialization, LBM.cu:68-146 */
void eko_initialization(eko_state *s)
{
    const eko_params *p = &s->p;
    const int NX = p->NX, NY = p->NY, NZ = p->NZ;
    /* gpu_initialization, LBM.cu:111-128 */
    for (int z = 0; z < NZ; ++z)
    for (int y = 0; y < NY; ++y)
    for (int x = 0; x < NX; ++x) {
        const size_t i = sidx(p, x, y, z);
        s->fld[EKO_RHO][i] = p->rho0;
        s->fld[EKO_CHARGE][i] = 0.0;
        s->fld[EKO_CHARGEN][i] = 0.0;
        s->fld[EKO_PHI][i] = p->voltage;
        s->fld[EKO_UX][i] = 0.0; s->fld[EKO_UY][i] = 0.0; s->fld[EKO_UZ][i] = 0.0;
        s->fld[EKO_EX][i] = 0.0; s->fld[EKO_EY][i] = 0.0; s->fld[EKO_EZ][i] = 0.0;
        s->fld[EKO_T][i] = p->TH * (p->Lz - p->dz * z) / p->Lz;
    }
    memcpy(s->phi_old, s->fld[EKO_PHI], s->N * sizeof(double));  /* LBM.cu:82-86 */
    double *c = s->fld[EKO_CHARGE], *cn = s->fld[EKO_CHARGEN], *fi = s->fld[EKO_PHI];
    for (int it = 0; it < p->pb_iters; ++it) {                    /* LBM.cu:89 */
        /* gpu_PBE, LBM.cu:139-146 */
        for (size_t i = 0; i < s->N; ++i) {
            c[i] = p->chargeinf * exp(-p->electron * fi[i] / p->kB / p->roomT);
            cn[i] = p->chargeinf * exp(p->electron * fi[i] / p->kB / p->roomT);
        }
        eko_fast_poisson(s);                                      /* LBM.cu:96 */
        /* gpu_PBE_phi, LBM.cu:131-137 */
        for (size_t i = 0; i < s->N; ++i)
            fi[i] = p->PB_omega * fi[i] + (1.0 - p->PB_omega) * s->phi_old[i];
        memcpy(s->phi_old, fi, s->N * sizeof(double));            /* LBM.cu:101-104 */
    }
}

/* gpu_init_equilibrium, LBM.cu:162-463 */
void eko_init_equilibrium(eko_state *s)
{
    const eko_params *p = &s->p;
    const int NX = p->NX, NY = p->NY, NZ = p->NZ;
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < NZ; ++z)
    for (int y = 0; y < NY; ++y)
    for (int x = 0; x < NX; ++x) {
        const size_t si = sidx(p, x, y, z);
        const double rho = s->fld[EKO_RHO][si];
        const double ux = s->fld[EKO_UX][si], uy = s->fld[EKO_UY][si], uz = s->fld[EKO_UZ][si];
        const double charge = s->fld[EKO_CHARGE][si], chargen = s->fld[EKO_CHARGEN][si];
        const double Ex = s->fld[EKO_EX][si], Ey = s->fld[EKO_EY][si], Ez = s->fld[EKO_EZ][si];
        const double Temp = s->fld[EKO_T][si];
        double eq[4][27];
        equilibrium(p, rho, ux, uy, uz, eq[0]);
        equilibrium(p, charge, ux + p->K * Ex, uy + p->K * Ey, uz + p->K * Ez, eq[1]);
        equilibrium(p, chargen, ux + p->Kn * Ex, uy + p->Kn * Ey, uz + p->Kn * Ez, eq[2]);
        equilibrium(p, Temp, ux, uy, uz, eq[3]);
        for (int k = 0; k < 4; ++k) {
            s->x0[k][si] = eq[k][0];
            for (int d = 1; d < 27; ++d) s->x1[k][nidx(p, x, y, z, d)] = eq[k][d];
        }
    }
}

/* one iteration of main.cu:189-200 per step */
void eko_step(eko_state *s, int nsteps)
{
    for (int i = 0; i < nsteps; ++i) {
        eko_stream_collide_save(s);
        eko_fast_poisson(s);
    }
}
