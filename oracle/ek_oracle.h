/*
 * ek_oracle.h -- CPU restatement of EK-PNP-3D's coupled time step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 *
 * Parity status: the reference ships no tests, golden vectors or fixtures
 * (SURVEY.md section 8c).  This restatement is pinned against raw fp64 dumps
 * of the reference's own CUDA build (oracle/_ref, built by oracle/build_ref.py
 * from the sources under /root/reference and run on a B200); the committed
 * fixtures are under tests/golden/ together with the script that made them.
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference).
 */
#ifndef EK_ORACLE_H
#define EK_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Input parameters: the compile-time constants of LBM.h:29-125. */
typedef struct eko_params {
    int NX, NY, NZ;                 /* LBM.h:32-35 */
    double Lx, Ly, Lz;              /* LBM.h:40-42 */
    double dx, dy, dz;              /* LBM.h:43-45 */
    double uw, exf;                 /* LBM.h:47-50 */
    double CFL, dt, cs_square, rho0;/* LBM.h:51-54 */
    double chargeinf;               /* LBM.h:56 */
    double voltage, voltage2;       /* LBM.h:60,62 */
    double Ext, eps;                /* LBM.h:64-65 */
    double diffu, nu, K;            /* LBM.h:66-70 */
    double diffun, Kn;              /* LBM.h:73-76 */
    double kB, electron, roomT, convertCtoCharge, PB_omega; /* LBM.h:87-91 */
    double D, Ra, TH;               /* LBM.h:95-98 */
    double w0, ws, wa, wd;          /* LBM.h:109-112 */
    double V, VC, VCn, VT;          /* LBM.h:115-118 */
    int pb_iters;                   /* LBM.cu:89 (501 as shipped) */
} eko_params;

typedef struct eko_state eko_state;

/* field ids, in the column order of the reference's data_end.dat (LBM.cu:2613) */
enum {
    EKO_RHO = 0, EKO_UX, EKO_UY, EKO_UZ, EKO_CHARGE, EKO_CHARGEN,
    EKO_PHI, EKO_T, EKO_EX, EKO_EY, EKO_EZ, EKO_NFIELDS
};

void eko_default_params(eko_params *p);            /* LBM.h as shipped */
eko_state *eko_create(const eko_params *p);        /* main.cu:78-152 */
void eko_destroy(eko_state *s);
void eko_initialization(eko_state *s);             /* LBM.cu:68-146 */
void eko_init_equilibrium(eko_state *s);           /* LBM.cu:150-463 */
void eko_step(eko_state *s, int nsteps);           /* main.cu:189-200 */
void eko_stream_collide_save(eko_state *s);        /* LBM.cu:465-481 */
void eko_fast_poisson(eko_state *s);               /* poisson.cu:75-103 */
double *eko_field(eko_state *s, int id);           /* N doubles, x fastest */
/* DC mode of the extended right-hand side: 0 zero, 1 literal (poisson.cu:176-177,
 * default), 2 prescribed forward coefficient ghat0 */
void eko_set_poisson_dc(eko_state *s, int mode, double ghat0);
double eko_last_dc(eko_state *s);
/* populations of set 0..3 (fluid, cation, anion, temperature):
 * 27*N doubles, [d][z][y][x]; pre-collision ("X1") state */
void eko_get_populations(eko_state *s, int set, double *out);
int eko_num_threads(void);

/* 1-D complex DFT of arbitrary length, exposed for unit tests */
void eko_fft1d(double *re_im_interleaved, int n, int stride, int sign);

#ifdef __cplusplus
}
#endif
#endif
