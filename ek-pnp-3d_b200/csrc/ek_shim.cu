// ek_shim.cu -- link-compatible stand-ins for the reference's hot-path entry
// points, so that main.cu's loop (main.cu:189-224) runs on this library
// unchanged.  Built into libek_b200_shim.so; see INTEGRATION.md.
//
// The functions below have the reference's exact C++ signatures
// (LBM.h:159-176):
//   init_equilibrium(18 ptrs)      LBM.cu:150   -> ek_init_equilibrium
//   stream_collide_save(24 args)   LBM.cu:465   -> ek_stream_collide_save
//   fast_Poisson(6 args)           poisson.cu:75 -> ek_fast_poisson
// The caller's macroscopic arrays are ADOPTED (zero copy): the kernels write
// rho,u,c+,c-,T and the solver writes phi,Ex,Ey,Ez straight into them, which
// is what save_data_tecplot/current/record_umax read back.  The caller's
// population arrays (f0/f1/f2, ...) are not touched: populations live in the
// handle's own in-place lattice (their layout is private in the reference,
// SURVEY.md 8b).
#include <cufft.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/ek_b200.h"

namespace {
ek_handle *g_h = nullptr;
ek_params g_p;
bool g_configured = false;
double *g_phi = nullptr, *g_E[3] = {nullptr, nullptr, nullptr};

void die(const char *what, ek_status st)
{
    // the reference's error contract: message on stderr, then exit (LBM.cu:35-53)
    fprintf(stderr, "ek_b200 shim: %s failed: status %d: %s\n", what, (int)st, g_h ? ek_last_error(g_h) : "");
    exit(-1);
}

void ensure_handle()
{
    if (g_h) return;
    if (!g_configured) ek_default_params(&g_p);
    ek_status st = ek_create(&g_p, -1, &g_h);
    if (st != EK_OK) die("ek_create", st);
}

void adopt(int id, double *p)
{
    ek_status st = ek_adopt_field(g_h, id, p);
    if (st != EK_OK) die("ek_adopt_field", st);
}
}  // namespace

// The reference has no runtime parameters (everything is a constant of LBM.h);
// an adapted main() passes them once.  Optional: LBM.h as shipped is the default.
extern "C" void ek_shim_configure(const ek_params *p)
{
    g_p = *p;
    g_configured = true;
}

// phi_gpu, Ex_gpu, Ey_gpu, Ez_gpu are globals of the reference (LBM.h:139-141)
// that fast_Poisson writes without receiving them (poisson.cu:95,98).
extern "C" void ek_shim_bind_potential(double *phi, double *ex, double *ey, double *ez)
{
    g_phi = phi; g_E[0] = ex; g_E[1] = ey; g_E[2] = ez;
}

extern "C" ek_handle *ek_shim_handle(void) { return g_h; }

void init_equilibrium(double *f0, double *f1, double *h0, double *h1, double *hn0, double *hn1, double *temp0,
                      double *temp1, double *r, double *c, double *cn, double *u, double *v, double *w, double *ex,
                      double *ey, double *ez, double *temp)
{
    (void)f0; (void)f1; (void)h0; (void)h1; (void)hn0; (void)hn1; (void)temp0; (void)temp1;
    ensure_handle();
    adopt(EK_RHO, r); adopt(EK_CHARGE, c); adopt(EK_CHARGEN, cn);
    adopt(EK_UX, u); adopt(EK_UY, v); adopt(EK_UZ, w);
    adopt(EK_EX, ex); adopt(EK_EY, ey); adopt(EK_EZ, ez); adopt(EK_T, temp);
    if (g_phi) adopt(EK_PHI, g_phi);
    ek_status st = ek_mark_fields_ready(g_h);
    if (st == EK_OK) st = ek_init_equilibrium(g_h);
    if (st != EK_OK) die("init_equilibrium", st);
}

void stream_collide_save(double *f0, double *f1, double *f2, double *h0, double *h1, double *h2, double *hn0,
                         double *hn1, double *hn2, double *temp0, double *temp1, double *temp2, double *r, double *c,
                         double *cn, double *u, double *v, double *w, double *ex, double *ey, double *ez,
                         double *Temp, double t, double *f0bc)
{
    (void)f0; (void)f1; (void)f2; (void)h0; (void)h1; (void)h2; (void)hn0; (void)hn1; (void)hn2;
    (void)temp0; (void)temp1; (void)temp2; (void)r; (void)c; (void)cn; (void)u; (void)v; (void)w;
    (void)ex; (void)ey; (void)ez; (void)Temp; (void)t; (void)f0bc;
    if (!g_h) { fprintf(stderr, "ek_b200 shim: stream_collide_save before init_equilibrium\n"); exit(-1); }
    ek_status st = ek_stream_collide_save(g_h, 1);
    if (st != EK_OK) die("stream_collide_save", st);
}

void fast_Poisson(double *charge, double *chargen, double *kx, double *ky, double *kz, cufftHandle plan)
{
    (void)kx; (void)ky; (void)kz; (void)plan;
    // also called by the reference's own initialization() (LBM.cu:96), i.e.
    // before init_equilibrium: adopt what we are given and solve
    ensure_handle();
    adopt(EK_CHARGE, charge);
    adopt(EK_CHARGEN, chargen);
    if (g_phi) { adopt(EK_PHI, g_phi); adopt(EK_EX, g_E[0]); adopt(EK_EY, g_E[1]); adopt(EK_EZ, g_E[2]); }
    ek_status st = ek_refresh_charge_difference(g_h);
    if (st == EK_OK) st = ek_fast_poisson(g_h, 1);
    if (st == EK_OK) st = ek_sync(g_h);   // the reference's cudaFree calls synchronise here (poisson.cu:100-102)
    if (st != EK_OK) die("fast_Poisson", st);
}
