// ek_main.cpp -- the reference's main() (main.cu:19-295) over the C ABI of include/ek_b200.h.
//
// Same program flow, same console messages and the same output files (data.dat in the
// Tecplot format of save_data_tecplot, umax.dat, data_end.dat) as gyf135/EK-PNP-3D's main.cu,
// with the simulation loop main.cu:189-224 running on libek_b200.  Plain C++ (no CUDA in this
// translation unit): a maintainer of the reference can build it with
//     g++ -O2 -Iinclude examples/ek_main.cpp -Lek-pnp-3d_b200 -lek_b200 -Wl,-rpath,'$ORIGIN' -o ek_main
// The constants of LBM.h:29-125 are the defaults of ek_default_params(); the command line
// overrides the ones that LBM.h makes people edit (grid, step counts, a few physical inputs):
//     ek_main [--nx N] [--ny N] [--nz N] [--nsteps N] [--nsave N] [--print-current N]
//             [--ext V/m] [--exf N/m3] [--uw m/s] [--th K] [--cinf mol] [--pb-iters N]
//             [--restart 0|1] [--checkpoint file] [--resume file] [--device d] [--gpus N]
// Without --restart the program asks on stdin like main.cu:158-159.  --gpus N splits the domain into
// N x-slabs on devices 0..N-1 of this process (ek_multi_*, NX must be divisible by 2N).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include "ek_b200.h"

static ek_multi *g_multi = nullptr;

static void die(ek_handle *h, const char *what, ek_status st)
{
    // the reference's contract: message on stderr, exit(-1) (LBM.cu:35-53)
    fprintf(stderr, "%s failed: status %d: %s\n", what, (int)st,
            g_multi ? ek_multi_last_error(g_multi) : (h ? ek_last_error(h) : ""));
    exit(-1);
}
#define CK(call) do { ek_status _s = (call); if (_s != EK_OK) die(h, #call, _s); } while (0)

// the same calls on one handle or on the slabs of an ek_multi
struct Sim {
    ek_handle *h = nullptr;
    ek_multi *m = nullptr;
    ek_status init_fields() { return m ? ek_multi_init_fields(m) : ek_init_fields(h); }
    ek_status init_equilibrium() { return m ? ek_multi_init_equilibrium(m) : ek_init_equilibrium(h); }
    ek_status step_timed(int n, float *ms) { return m ? ek_multi_step_timed(m, n, ms) : ek_step_timed(h, n, ms); }
    ek_status tecplot(const char *p, double t, int append, int first)
    {
        return m ? ek_multi_save_data_tecplot(m, p, t, append, first) : ek_save_data_tecplot(h, p, t, append, first);
    }
    ek_status save_end(const char *p, double t) { return m ? ek_multi_save_data_end(m, p, t) : ek_save_data_end(h, p, t); }
    ek_status wall_current(double *I) { return m ? ek_multi_wall_current(m, I) : ek_wall_current(h, I); }
    ek_status max_uz(double *u) { return m ? ek_multi_max_uz(m, u) : ek_max_uz(h, u); }
    ek_status sync() { return m ? ek_multi_sync(m) : ek_sync(h); }
};

int main(int argc, char **argv)
{
    ek_params P;
    ek_default_params(&P);                                   // LBM.h:29-125 as shipped
    unsigned NSTEPS = 1000, NSAVE = 0, printCurrent = 50;     // LBM.h:122-125
    int restart = -1, device = 0, gpus = 1;
    const char *ckpt_out = nullptr, *ckpt_in = nullptr;
    bool grid_changed = false;
    for (int i = 1; i + 1 < argc; i += 2) {
        const char *k = argv[i], *v = argv[i + 1];
        if (!strcmp(k, "--nx")) { P.NX = atoi(v); grid_changed = true; }
        else if (!strcmp(k, "--ny")) { P.NY = atoi(v); grid_changed = true; }
        else if (!strcmp(k, "--nz")) { P.NZ = atoi(v); grid_changed = true; }
        else if (!strcmp(k, "--nsteps")) NSTEPS = (unsigned)atoi(v);
        else if (!strcmp(k, "--nsave")) NSAVE = (unsigned)atoi(v);
        else if (!strcmp(k, "--print-current")) printCurrent = (unsigned)atoi(v);
        else if (!strcmp(k, "--ext")) P.Ext = atof(v);
        else if (!strcmp(k, "--exf")) P.exf = atof(v);
        else if (!strcmp(k, "--uw")) P.uw = atof(v);
        else if (!strcmp(k, "--th")) P.TH = atof(v);
        else if (!strcmp(k, "--cinf")) P.chargeinf = atof(v);
        else if (!strcmp(k, "--pb-iters")) P.pb_iters = atoi(v);
        else if (!strcmp(k, "--restart")) restart = atoi(v);
        else if (!strcmp(k, "--checkpoint")) ckpt_out = v;
        else if (!strcmp(k, "--resume")) ckpt_in = v;
        else if (!strcmp(k, "--device")) device = atoi(v);
        else if (!strcmp(k, "--gpus")) gpus = atoi(v);
        else { fprintf(stderr, "unknown option %s\n", k); return 2; }
    }
    if (grid_changed) {   // LBM.h:40-42: the box follows the grid at the shipped spacing
        P.Lx = P.NX * P.dx; P.Ly = P.NY * P.dy; P.Lz = (P.NZ - 1) * P.dz;
    }
    if (NSAVE == 0) NSAVE = NSTEPS / 2 ? NSTEPS / 2 : 1;      // LBM.h:123
    if (printCurrent == 0) printCurrent = 1;

    ek_handle *h = nullptr;
    Sim sim;
    if (gpus > 1) {
        int devs[16];
        const int ndev = ek_device_count() > 0 ? ek_device_count() : 1;
        if (gpus > 16) gpus = 16;
        for (int d = 0; d < gpus; ++d) devs[d] = d % ndev;   // fewer devices than slabs: slabs share devices
        ek_status st = ek_multi_create(&P, gpus, devs, 0, &sim.m);
        if (st != EK_OK) die(nullptr, "ek_multi_create (NX divisible by 2*gpus, <= 16 CUDA devices of one node)", st);
        g_multi = sim.m;
    } else {
        ek_status st = ek_create(&P, device, &h);
        if (st != EK_OK) die(nullptr, "ek_create (a CUDA device is required, there is no CPU path)", st);
        sim.h = h;
    }

    printf("Simulating 3D electrokinetic flow (EK-PNP) on a %d x %d x %d grid\n", P.NX, P.NY, P.NZ);
    double t = 0.0;
    if (ckpt_in) {
        printf("Resuming from checkpoint %s...\n", ckpt_in);
        CK(sim.m ? ek_multi_checkpoint_load(sim.m, ckpt_in, &t) : ek_checkpoint_load(h, ckpt_in, &t));
    } else {
        if (restart < 0) {
            printf("Read previous data: Press 1. Start a new simulation: Press 0.\n ");   // main.cu:158
            if (scanf("%d", &restart) != 1) restart = 0;
        }
        if (restart == 1) {
            printf("Reading previous data...\n");                                          // main.cu:162
            CK(sim.m ? ek_multi_read_data(sim.m, "data_end.dat", &t) : ek_read_data(h, "data_end.dat", &t));
        } else {
            printf("Initializing...\n");                                                    // main.cu:166
            CK(sim.init_fields());                                                          // main.cu:169
            t = 0;
        }
        CK(sim.init_equilibrium());                                                         // main.cu:174
    }
    CK(sim.tecplot("data.dat", t, 0, 1));                                                   // main.cu:178-179
    FILE *fumax = fopen("umax.dat", "wb+");                                                // main.cu:180
    if (!fumax) { fprintf(stderr, "cannot open umax.dat\n"); return 1; }

    const auto begin = std::chrono::steady_clock::now();
    float gpu_ms = 0.0f;
    // main simulation loop (main.cu:189-224).  The reference looks at the step counter after
    // every step; here the steps between two look-ups run as one ek_step_timed() call.
    unsigned i = 0;
    while (i < NSTEPS) {
        unsigned next = NSTEPS - 1;   // last index
        for (unsigned j = i; j < NSTEPS; ++j)
            if (j % NSAVE == 1 || j % printCurrent == 1) { next = j; break; }
        const int n = (int)(next - i + 1);
        float ms = 0.0f;
        CK(sim.step_timed(n, &ms));
        gpu_ms += ms;
        t += n * P.dt;
        i = next;
        if (i % NSAVE == 1) {
            CK(sim.tecplot("data.dat", t, 1, 1));                                           // main.cu:206-209
            printf("Iteration: %u, physical time: %g.\n", i, t);
        }
        if (i % printCurrent == 1) {
            double I = 0.0, umax = 0.0;
            CK(sim.wall_current(&I));                                                       // main.cu:211-216
            printf("Iteration: %u, physical time: %g, Current = %g\n", i, t, I);
            CK(sim.max_uz(&umax));                                                          // main.cu:221
            fprintf(fumax, "%10.6f %10.6f\n", t, umax);                                     // LBM.cu:2747
        }
        ++i;
    }
    CK(sim.sync());
    const double runtime = std::chrono::duration<double>(std::chrono::steady_clock::now() - begin).count();
    const double nodes_updated = (double)NSTEPS * (double)P.NX * P.NY * P.NZ;              // main.cu:239
    printf(" ----- performance information -----\n");                                       // main.cu:247-251
    printf("               timesteps: %u\n", NSTEPS);
    printf("           clock runtime: %.3f (s)\n", runtime);
    printf("             gpu runtime: %.3f (s)\n", 0.001 * gpu_ms);
    printf("                   speed: %.2f (Mlups)\n", nodes_updated / (1e6 * runtime));

    CK(sim.tecplot("data.dat", t, 1, 1));                                                   // main.cu:253
    fclose(fumax);
    CK(sim.save_end("data_end.dat", t));                                                    // main.cu:256-257
    if (ckpt_out) CK(sim.m ? ek_multi_checkpoint_save(sim.m, ckpt_out, t) : ek_checkpoint_save(h, ckpt_out, t));
    if (sim.m) ek_multi_destroy(sim.m); else ek_destroy(h);
    return 0;
}
