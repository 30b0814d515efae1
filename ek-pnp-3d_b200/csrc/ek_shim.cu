// ek_shim.cu -- link-compatible stand-ins for the reference's hot-path entry
// points, so that main.cu's loop (main.cu:189-224) runs on this library
// unchanged.  Built into libek_b200_shim.so; see INTEGRATION.md.
//
// The functions below have the reference's exact C++ signatures
// (LBM.h:159-176):
//   initialization(11 ptrs)        LBM.cu:68    -> ek_init_fields (device-resident PB loop)
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
#include <string.h>

#include <string>

#include "../../include/ek_b200.h"

// The reference's globals that fast_Poisson writes without receiving them (LBM.h:139-141,
// poisson.cu:95,98).  When the reference's main.cu is linked against this library the dynamic
// linker resolves them to main's definitions; otherwise (ctypes, tests) they stay null and
// ek_shim_bind_potential() supplies the arrays.
extern double *phi_gpu __attribute__((weak));
extern double *Ex_gpu __attribute__((weak));
extern double *Ey_gpu __attribute__((weak));
extern double *Ez_gpu __attribute__((weak));

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

// EK_SHIM_PARAMS="NX=128,NY=64,NZ=64,TH=0": the constants a maintainer changed in LBM.h (the
// reference has no runtime parameters and its constants have internal linkage, so an unmodified
// main.cu cannot hand them over); Lx, Ly, Lz follow the grid unless given.
void params_from_env(ek_params &p)
{
    const char *env = getenv("EK_SHIM_PARAMS");
    if (!env || !*env) return;
    struct { const char *name; int *i; double *d; } tab[] = {
        {"NX", &p.NX, nullptr}, {"NY", &p.NY, nullptr}, {"NZ", &p.NZ, nullptr}, {"pb_iters", &p.pb_iters, nullptr},
        {"Lx", nullptr, &p.Lx}, {"Ly", nullptr, &p.Ly}, {"Lz", nullptr, &p.Lz},
        {"dx", nullptr, &p.dx}, {"dy", nullptr, &p.dy}, {"dz", nullptr, &p.dz},
        {"uw", nullptr, &p.uw}, {"exf", nullptr, &p.exf}, {"uw_host", nullptr, &p.uw}, {"exf_host", nullptr, &p.exf},
        {"chargeinf", nullptr, &p.chargeinf}, {"voltage", nullptr, &p.voltage}, {"voltage2", nullptr, &p.voltage2},
        {"Ext", nullptr, &p.Ext}, {"eps", nullptr, &p.eps}, {"diffu", nullptr, &p.diffu}, {"diffun", nullptr, &p.diffun},
        {"nu", nullptr, &p.nu}, {"K", nullptr, &p.K}, {"Kn", nullptr, &p.Kn}, {"D", nullptr, &p.D}, {"Ra", nullptr, &p.Ra},
        {"TH", nullptr, &p.TH}, {"PB_omega", nullptr, &p.PB_omega}, {"rho0", nullptr, &p.rho0},
    };
    bool lx = false, ly = false, lz = false;
    std::string all(env);
    size_t pos = 0;
    while (pos < all.size()) {
        size_t end = all.find(',', pos);
        if (end == std::string::npos) end = all.size();
        const std::string item = all.substr(pos, end - pos);
        pos = end + 1;
        const size_t eq = item.find('=');
        if (eq == std::string::npos) continue;
        const std::string key = item.substr(0, eq), val = item.substr(eq + 1);
        bool found = false;
        for (auto &t : tab)
            if (key == t.name) {
                if (t.i) *t.i = atoi(val.c_str()); else *t.d = atof(val.c_str());
                found = true;
            }
        if (!found) { fprintf(stderr, "ek_b200 shim: EK_SHIM_PARAMS: unknown parameter '%s'\n", key.c_str()); exit(-1); }
        lx |= key == "Lx"; ly |= key == "Ly"; lz |= key == "Lz";
    }
    if (!lx) p.Lx = p.NX * p.dx;
    if (!ly) p.Ly = p.NY * p.dy;
    if (!lz) p.Lz = (p.NZ - 1) * p.dz;
}

void ensure_handle()
{
    if (g_h) return;
    if (!g_configured) { ek_default_params(&g_p); params_from_env(g_p); }
    ek_status st = ek_create(&g_p, -1, &g_h);
    if (st != EK_OK) die("ek_create", st);
    // the reference launches everything on the legacy default stream and reads its arrays back with
    // blocking cudaMemcpy (LBM.cu:2511-2521): run there too, so the caller's ordering assumptions hold
    st = ek_set_stream(g_h, nullptr);
    if (st != EK_OK) die("ek_set_stream", st);
    const char *dc = getenv("EK_SHIM_DC");   // "literal": mu(0,0,0) := 1 with this library's own rounding residue
    if (dc && !strcmp(dc, "literal")) ek_set_poisson_dc(g_h, EK_DC_LITERAL, 0.0);
}

// where fast_Poisson writes: bound explicitly, or the reference's globals when main.cu is linked in
bool potential_arrays(double **phi, double *E[3])
{
    if (g_phi) { *phi = g_phi; E[0] = g_E[0]; E[1] = g_E[1]; E[2] = g_E[2]; return true; }
    if (&phi_gpu && &Ex_gpu && &Ey_gpu && &Ez_gpu && phi_gpu) {
        *phi = phi_gpu; E[0] = Ex_gpu; E[1] = Ey_gpu; E[2] = Ez_gpu;
        return true;
    }
    return false;
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

// LBM.cu:68-109: uniform state + 501 under-relaxed Poisson-Boltzmann iterations, written into the
// caller's arrays (here entirely on the device: no host round trips of phi, LBM.cu:101-104)
void initialization(double *r, double *c, double *cn, double *fi, double *u, double *v, double *w, double *ex, double *ey,
                    double *ez, double *temp)
{
    ensure_handle();
    adopt(EK_RHO, r); adopt(EK_CHARGE, c); adopt(EK_CHARGEN, cn); adopt(EK_PHI, fi);
    adopt(EK_UX, u); adopt(EK_UY, v); adopt(EK_UZ, w);
    adopt(EK_EX, ex); adopt(EK_EY, ey); adopt(EK_EZ, ez); adopt(EK_T, temp);
    g_phi = fi; g_E[0] = ex; g_E[1] = ey; g_E[2] = ez;
    ek_status st = ek_init_fields(g_h);
    if (st == EK_OK) st = ek_sync(g_h);
    if (st != EK_OK) die("initialization", st);
}

void init_equilibrium(double *f0, double *f1, double *h0, double *h1, double *hn0, double *hn1, double *temp0,
                      double *temp1, double *r, double *c, double *cn, double *u, double *v, double *w, double *ex,
                      double *ey, double *ez, double *temp)
{
    (void)f0; (void)f1; (void)h0; (void)h1; (void)hn0; (void)hn1; (void)temp0; (void)temp1;
    ensure_handle();
    adopt(EK_RHO, r); adopt(EK_CHARGE, c); adopt(EK_CHARGEN, cn);
    adopt(EK_UX, u); adopt(EK_UY, v); adopt(EK_UZ, w);
    adopt(EK_EX, ex); adopt(EK_EY, ey); adopt(EK_EZ, ez); adopt(EK_T, temp);
    double *phi = nullptr, *E[3];
    if (potential_arrays(&phi, E)) adopt(EK_PHI, phi);
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
    double *phi = nullptr, *E[3];
    if (!potential_arrays(&phi, E)) {
        // the reference's fast_Poisson writes phi_gpu/Ex_gpu/Ey_gpu/Ez_gpu implicitly (poisson.cu:95,98):
        // solving into private arrays would leave the caller iterating on a stale potential
        fprintf(stderr, "ek_b200 shim: fast_Poisson: the potential arrays are unknown -- link main.cu against this "
                        "library (phi_gpu, Ex_gpu, Ey_gpu, Ez_gpu are then resolved by symbol) or call "
                        "ek_shim_bind_potential(phi, Ex, Ey, Ez) first\n");
        exit(-1);
    }
    adopt(EK_PHI, phi); adopt(EK_EX, E[0]); adopt(EK_EY, E[1]); adopt(EK_EZ, E[2]);
    ek_status st = ek_refresh_charge_difference(g_h);
    if (st == EK_OK) st = ek_fast_poisson(g_h, 1);
    if (st == EK_OK) st = ek_sync(g_h);   // the reference's cudaFree calls synchronise here (poisson.cu:100-102)
    if (st != EK_OK) die("fast_Poisson", st);
}
