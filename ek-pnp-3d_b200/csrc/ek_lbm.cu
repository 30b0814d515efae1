// ek_lbm.cu -- fused stream-and-collide step for the four D3Q27 TRT
// population sets (fluid, cation, anion, temperature) of EK-PNP-3D.
//
// One launch per time step replaces the reference's gpu_collide_save,
// gpu_boundary, gpu_stream and gpu_bc_charge (LBM.cu:465-2416) and the
// separate gpu_efield/gpu_bc field kernels (poisson.cu:40-69):
//
//  * one warp ROLE per population set.  A CTA is four warps working on the same
//    32 cells (one x-row segment): warp 0 = fluid, 1 = cation, 2 = anion,
//    3 = temperature (+ E = -grad phi).  Each thread holds the 27 populations of
//    ONE set in registers, so nothing spills (the reference keeps all 108 per
//    thread: 255 registers + 10 KB of local-memory traffic per cell);
//  * the cross-set dependency inside a cell (u needs c+, c-, T; the scalar
//    equilibria need u) is resolved through 2.3 KB of shared memory and two
//    named barriers per cell row, not through DRAM;
//  * streaming is fused: in-place A-A pattern (even step: read slot d, write
//    slot opp(d) at the node; odd step: read slot opp(d) at x-c_d, write slot
//    d at x+c_d) or a two-lattice collide-and-push.  Every population is read
//    once and written once per step: 4*27*16 B = 1728 B per cell update;
//  * walls are folded in: full-way bounce-back of the
//    pre-collision fluid populations with the frozen wall rest population,
//    z-periodic "ghost" streaming through the walls for the fluid, post-collision
//    swap for the ions, anti-bounce-back Dirichlet for the temperature.  The
//    scalar sets keep their wall-node state in a small side buffer so that the
//    in-place scheme stays race free;
//  * a CTA walks a chunk of z so that the thread that owns the z = 0 node also
//    owns z = 1: the bottom wall takes minus the z = 1 momentum (LBM.cu:663-801)
//    and in an in-place scheme only the owner may read that node.
#include "ek_internal.cuh"

namespace {

struct Sh {
    double cp[32], cn[32], T[32];
    double E[3][32];
    double u[3][32];
};

// Offsets of the 3x3x3 neighbourhood of a node.  Populations live in a tiled
// layout per set, [z][y][x-tile][27 slots][32 lanes] (ek_internal.cuh): the slot
// stride is a compile-time 256 B, so the 27 accesses of a node differ only by an
// immediate and one 64-bit address per neighbour position is all the integer
// work a gather or scatter needs.  Macroscopic fields stay in the reference's
// [z][y][x] order (LBM.cu:22-25).
struct Nbr {
    unsigned lx[3], ly[3], lz[3];  // lattice element offsets of x-1,x,x+1 / y-1,y,y+1 / z-1,z,z+1 (z periodic)
    int fx[3], fy[3], fz[3];       // the same for the field arrays
    __device__ __forceinline__ unsigned at(int ax, int ay, int az) const { return lz[az + 1] + ly[ay + 1] + lx[ax + 1]; }
    __device__ __forceinline__ unsigned lc() const { return lz[1] + ly[1] + lx[1]; }
    __device__ __forceinline__ int fc() const { return fz[1] + fy[1] + fx[1]; }
};

__device__ __forceinline__ void set_xy(Nbr &nb, const EkConst &c, int x, int y)
{
    nb.fx[0] = x == 0 ? c.xlo : x - 1;
    nb.fx[1] = x;
    nb.fx[2] = x == c.NX - 1 ? c.xhi : x + 1;
    const int ym = y == 0 ? c.NY - 1 : y - 1, yp = y == c.NY - 1 ? 0 : y + 1;
    nb.fy[0] = ym * c.PX; nb.fy[1] = y * c.PX; nb.fy[2] = yp * c.PX;
#pragma unroll
    for (int k = 0; k < 3; ++k) nb.lx[k] = ek_lat_col(nb.fx[k]);
    nb.ly[0] = (unsigned)ym * c.lrow; nb.ly[1] = (unsigned)y * c.lrow; nb.ly[2] = (unsigned)yp * c.lrow;
}

__device__ __forceinline__ void set_z(Nbr &nb, const EkConst &c, int z)
{
    const int zm = z == 0 ? c.NZ - 1 : z - 1;
    const int zp = z == c.NZ - 1 ? 0 : z + 1;
    nb.fz[0] = (int)(zm * c.plane); nb.fz[1] = (int)(z * c.plane); nb.fz[2] = (int)(zp * c.plane);
    nb.lz[0] = (unsigned)zm * c.lplane; nb.lz[1] = (unsigned)z * c.lplane; nb.lz[2] = (unsigned)zp * c.lplane;
}

__device__ __forceinline__ void bar_moments() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void bar_velocity() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

// pre-collision populations of a node ("pull" restatement)
template <int MODE>
__device__ __forceinline__ void gather27(const double *lat, const Nbr &nb, double S[27])
{
    if (MODE == EK_MODE_AA_ODD) {
#pragma unroll
        for (int d = 0; d < 27; ++d) {
            const double *q = lat + nb.at(-ek_cx(d), -ek_cy(d), -ek_cz(d));
            S[d] = q[ek_opp(d) * EK_TILE];
        }
    } else {
        const double *q = lat + nb.lc();
#pragma unroll
        for (int d = 0; d < 27; ++d) S[d] = q[d * EK_TILE];
    }
}

// where the post-collision population of direction d goes
template <int MODE, int d>
__device__ __forceinline__ void put(double *lat, const Nbr &nb, double v)
{
    if (MODE == EK_MODE_AA_EVEN) {
        double *q = lat + nb.lc();
        q[ek_opp(d) * EK_TILE] = v;
    } else {
        double *q = lat + nb.at(ek_cx(d), ek_cy(d), ek_cz(d));
        q[d * EK_TILE] = v;
    }
}

// LBM.cu:621-630: left-to-right sum in index order
__device__ __forceinline__ double sum27(const double S[27])
{
    double a = S[0];
#pragma unroll
    for (int d = 1; d < 27; ++d) a = a + S[d];
    return a;
}

// LBM.cu:639-644: the three momentum brackets, grouped as in the reference
__device__ __forceinline__ void momentum(const double f[27], double m[3])
{
    m[0] = (f[1] + f[7] + f[9] + f[13] + f[15] + f[19] + f[21] + f[23] + f[26]
          - (f[2] + f[8] + f[10] + f[14] + f[16] + f[20] + f[22] + f[24] + f[25]));
    m[1] = (f[3] + f[7] + f[11] + f[14] + f[17] + f[19] + f[21] + f[24] + f[25]
          - (f[4] + f[8] + f[12] + f[13] + f[18] + f[20] + f[22] + f[23] + f[26]));
    m[2] = (f[5] + f[9] + f[11] + f[16] + f[18] + f[19] + f[22] + f[23] + f[25]
          - (f[6] + f[10] + f[12] + f[15] + f[17] + f[20] + f[21] + f[24] + f[26]));
}

template <int d>
__device__ __forceinline__ double cdot(double ax, double ay, double az)
{
    double s = 0.0;
    bool first = true;
    if (ek_cx(d) != 0) { s = ek_cx(d) > 0 ? ax : -ax; first = false; }
    if (ek_cy(d) != 0) { s = first ? (ek_cy(d) > 0 ? ay : -ay) : (ek_cy(d) > 0 ? s + ay : s - ay); first = false; }
    if (ek_cz(d) != 0) { s = first ? (ek_cz(d) > 0 ? az : -az) : (ek_cz(d) > 0 ? s + az : s - az); }
    return s;
}

// E = -grad phi with the reference's wall treatment (poisson.cu:40-69):
// central differences, periodic x and y, Ez of the wall planes copied from
// the first interior plane.
template <bool EARR>
__device__ __forceinline__ void efield_at(const StepArgs &a, const Nbr &nb, int z, double E[3])
{
    const EkConst &c = a.c;
    if (EARR) {
        const int i = nb.fc();
        E[0] = a.E[0][i]; E[1] = a.E[1][i]; E[2] = a.E[2][i];
    } else {
        const double *phi = a.phi;
        const int zb = nb.fz[1];
        E[0] = 0.5 * (phi[zb + nb.fy[1] + nb.fx[0]] - phi[zb + nb.fy[1] + nb.fx[2]]) / c.dx;
        E[1] = 0.5 * (phi[zb + nb.fy[0] + nb.fx[1]] - phi[zb + nb.fy[2] + nb.fx[1]]) / c.dy;
        const int zc = z < 1 ? 1 : (z > c.NZ - 2 ? c.NZ - 2 : z);
        const int col = nb.fy[1] + nb.fx[1];
        E[2] = 0.5 * (phi[(size_t)(zc - 1) * c.plane + col] - phi[(size_t)(zc + 1) * c.plane + col]) / c.dz;
    }
}

// ---------------------------------------------------------------------------
// scalar roles: cation (s = 1), anion (s = 2), temperature (s = 3)
// ---------------------------------------------------------------------------
template <int MODE, int p>
struct ScalarPairs {
    // TRT relaxation of the opposite pair (d, d+1), d = 2p+1 (LBM.cu:1148-1845),
    // then delivery of both results.
    static __device__ __forceinline__ void run(const double S[27], double wcm[4], double omusq, double vtx, double vty,
                                               double vtz, double wp, double wmn, bool wall, bool bottom, bool is_temp,
                                               const EkConst &c, double *lout, const Nbr &nb, int z, double *Wn,
                                               bool act)
    {
        constexpr int d = 2 * p + 1, o = d + 1, cls = ek_wclass(d);
        const double s_ = cdot<d>(vtx, vty, vtz);
        const double wm_ = wcm[cls];
        const double ep = wm_ * (omusq + 0.5 * s_ * s_);
        const double em = wm_ * s_;
        const double a = S[d], b = S[o];
        const double np_ = wp * (0.5 * (a + b) - ep);
        const double nm_ = wmn * (0.5 * (a - b) - em);
        const double Oa = a - (np_ + nm_);
        const double Ob = b - (np_ - nm_);
        if (act) {
            if (!wall) {
                if (MODE == EK_MODE_AA_EVEN) {
                    put<MODE, d>(lout, nb, Oa);
                    put<MODE, o>(lout, nb, Ob);
                } else {
                    // inflow into wall nodes is discarded (LBM.cu:2102-2218 overwrite it)
                    const int za = z + ek_cz(d), zb = z - ek_cz(d);
                    if (ek_cz(d) == 0 || !(za == 0 || za == c.NZ - 1)) put<MODE, d>(lout, nb, Oa);
                    if (ek_cz(d) == 0 || !(zb == 0 || zb == c.NZ - 1)) put<MODE, o>(lout, nb, Ob);
                }
            } else {
                // the wall feeds the first interior plane ...
                if (ek_cz(d) != 0) {
                    const bool a_inward = bottom ? (ek_cz(d) > 0) : (ek_cz(d) < 0);
                    if (a_inward) put<MODE, d>(lout, nb, Oa);
                    else put<MODE, o>(lout, nb, Ob);
                }
                // ... and itself: ions swap post-collision populations
                // (LBM.cu:2132-2218), temperature anti-bounce-back (LBM.cu:2226-2413)
                if (!is_temp) {
                    Wn[(size_t)d * c.plane] = Ob;
                    Wn[(size_t)o * c.plane] = Oa;
                } else if (bottom) {
                    Wn[(size_t)d * c.plane] = -Ob + c.twoTw[cls];
                    Wn[(size_t)o * c.plane] = -Oa + c.twoTw[cls];
                } else {
                    Wn[(size_t)d * c.plane] = -Ob;
                    Wn[(size_t)o * c.plane] = -Oa;
                }
            }
        }
        ScalarPairs<MODE, p + 1>::run(S, wcm, omusq, vtx, vty, vtz, wp, wmn, wall, bottom, is_temp, c, lout, nb, z, Wn,
                                      act);
    }
};
template <int MODE>
struct ScalarPairs<MODE, 13> {
    static __device__ __forceinline__ void run(const double *, double *, double, double, double, double, double, double,
                                               bool, bool, bool, const EkConst &, double *, const Nbr &, int, double *,
                                               bool) {}
};

template <int MODE, bool FULL, bool EARR>
__device__ __forceinline__ void scalar_role(const StepArgs &a, Sh &sh, const int s, const int lane, const bool act,
                                            Nbr &nb, const int pi, const int z0, const int z1)
{
    const EkConst &c = a.c;
    double *lin = a.in[s];
    double *lout = a.out[s];
    double *W = a.wall + (size_t)(s - 1) * 2 * 27 * c.plane + pi;
    const double wp = c.wp[s], wmn = c.wm[s];
    const bool is_temp = (s == 3);
    const double Ks = s == 1 ? c.K : c.Kn;
    double *mom_sh = s == 1 ? sh.cp : (s == 2 ? sh.cn : sh.T);
    double S[27];

    if (z0 == 0) {
        // the bottom wall needs the moments of the z = 1 node first (LBM.cu:663-801)
        set_z(nb, c, 1);
        gather27<MODE>(lin, nb, S);
        mom_sh[lane] = sum27(S);
        if (is_temp) {
            double E[3];
            efield_at<EARR>(a, nb, 1, E);
            sh.E[0][lane] = E[0]; sh.E[1][lane] = E[1]; sh.E[2][lane] = E[2];
        }
        bar_moments();
        bar_velocity();
    }

    for (int z = z0; z < z1; ++z) {
        set_z(nb, c, z);
        const bool bottom = (z == 0);
        const bool wall = bottom || (z == c.NZ - 1);
        double *Wn = W + (size_t)(bottom ? 0 : 27) * c.plane;
        if (wall) {
#pragma unroll
            for (int d = 0; d < 27; ++d) S[d] = Wn[(size_t)d * c.plane];
        } else {
            gather27<MODE>(lin, nb, S);
        }
        const double m = sum27(S);
        double E[3] = {0.0, 0.0, 0.0};
        mom_sh[lane] = m;
        if (is_temp) {
            efield_at<EARR>(a, nb, z, E);
            sh.E[0][lane] = E[0]; sh.E[1][lane] = E[1]; sh.E[2][lane] = E[2];
        }
        if (FULL && act) a.fld[3 + s][nb.fc()] = m;  // charge, chargen, T (LBM.cu:811-813)
        bar_moments();
        // E is read between the two barriers: the temperature warp may only
        // overwrite it after every warp has passed bar_velocity()
        if (!is_temp) { E[0] = sh.E[0][lane]; E[1] = sh.E[1][lane]; E[2] = sh.E[2][lane]; }
        bar_velocity();
        double vx = sh.u[0][lane], vy = sh.u[1][lane], vz = sh.u[2][lane];
        if (!is_temp) {
            // ion drift u + K*E; Ext does not enter here (LBM.cu:851-862)
            vx = vx + Ks * E[0];
            vy = vy + Ks * E[1];
            vz = vz + Ks * E[2];
        }
        double wcm[4] = {c.w[0] * m, c.w[1] * m, c.w[2] * m, c.w[3] * m};
        const double omusq = 1.0 - 0.5 * (vx * vx + vy * vy + vz * vz) * c.inv_cs2;
        const double vtx = vx * c.tfac, vty = vy * c.tfac, vtz = vz * c.tfac;
        // rest population: relaxed in place (LBM.cu:1712-1714); wall rule LBM.cu:2131,2231,2385
        const double O0 = S[0] - wp * (S[0] - wcm[0] * omusq);
        if (act) {
            if (!wall) lout[nb.lc()] = O0;
            else if (!is_temp) Wn[0] = O0;
            else if (bottom) Wn[0] = -O0 + c.twoTw[0];
            else Wn[0] = -O0;
        }
        ScalarPairs<MODE, 0>::run(S, wcm, omusq, vtx, vty, vtz, wp, wmn, wall, bottom, is_temp, c, lout, nb, z, Wn, act);
    }
}

// ---------------------------------------------------------------------------
// fluid role
// ---------------------------------------------------------------------------
template <int MODE, int p>
struct FluidPairs {
    static __device__ __forceinline__ void run(const double S[27], double wcr[4], double omusq, const double u[3],
                                               const double F[3], double uF, bool wall, bool top, const EkConst &c,
                                               double *lout, const Nbr &nb, bool act)
    {
        constexpr int d = 2 * p + 1, o = d + 1, cls = ek_wclass(d);
        double Oa, Ob;
        if (!wall) {
            const double cu = cdot<d>(u[0], u[1], u[2]);
            const double cF = cdot<d>(F[0], F[1], F[2]);
            const double s_ = cu * c.tfac;
            const double wr = wcr[cls];
            const double ep = wr * (omusq + 0.5 * s_ * s_);
            const double em = wr * s_;
            const double a = S[d], b = S[o];
            const double np_ = c.wp[0] * (0.5 * (a + b) - ep);
            const double nm_ = c.wm[0] * (0.5 * (a - b) - em);
            // Guo forcing split into its symmetric / antisymmetric parts
            // (LBM.cu:1107-1145, 1608-1689): F+ = coe*((c.u)(c.F)*cflinv2 - u.F), F- = coe*cflinv*(c.F)
            const double Fp = c.sp * (c.coe[cls] * (cu * cF * c.cflinv2 - uF));
            const double Fm = c.sm * (c.coe[cls] * c.cflinv * cF);
            Oa = a - (np_ + nm_) + c.dt * (Fp + Fm);
            Ob = b - (np_ - nm_) + c.dt * (Fp - Fm);
        } else {
            // full-way bounce-back from the PRE-collision populations
            // (LBM.cu:1862-1887), moving-wall terms at the top (LBM.cu:1902-1927)
            Oa = S[o];
            Ob = S[d];
            if (top) {
                if (ek_uwsign(d) > 0) Oa = Oa + c.multi[cls]; else if (ek_uwsign(d) < 0) Oa = Oa - c.multi[cls];
                if (ek_uwsign(o) > 0) Ob = Ob + c.multi[cls]; else if (ek_uwsign(o) < 0) Ob = Ob - c.multi[cls];
            }
        }
        if (act) {
            put<MODE, d>(lout, nb, Oa);
            put<MODE, o>(lout, nb, Ob);
        }
        FluidPairs<MODE, p + 1>::run(S, wcr, omusq, u, F, uF, wall, top, c, lout, nb, act);
    }
};
template <int MODE>
struct FluidPairs<MODE, 13> {
    static __device__ __forceinline__ void run(const double *, double *, double, const double *, const double *, double,
                                               bool, bool, const EkConst &, double *, const Nbr &, bool) {}
};

// momentum expression of LBM.cu:639-644 without the 1/rho factor
__device__ __forceinline__ void node_force_and_momentum(const EkConst &c, const double m[3], double dq, double T,
                                                        const double E[3], double F[3], double expr[3])
{
    // LBM.cu:635-637 (Ext enters only here)
    F[0] = c.CtoC * dq * (E[0] + c.Ext) + c.exf;
    F[1] = c.CtoC * dq * E[1];
    F[2] = c.CtoC * dq * E[2] + c.rho0 * T * c.Ra * c.nu * c.D;
    expr[0] = m[0] * c.cflinv + F[0] * c.dt * 0.5;
    expr[1] = m[1] * c.cflinv + F[1] * c.dt * 0.5;
    expr[2] = m[2] * c.cflinv + F[2] * c.dt * 0.5;
}

template <int MODE, bool FULL>
__device__ __forceinline__ void fluid_role(const StepArgs &a, Sh &sh, const int lane, const bool act, Nbr &nb,
                                           const int z0, const int z1)
{
    const EkConst &c = a.c;
    double *lin = a.in[0];
    double *lout = a.out[0];
    double S[27];
    double expr1[3] = {0.0, 0.0, 0.0};

    if (z0 == 0) {
        set_z(nb, c, 1);
        gather27<MODE>(lin, nb, S);
        double m[3];
        momentum(S, m);
        bar_moments();
        const double E[3] = {sh.E[0][lane], sh.E[1][lane], sh.E[2][lane]};
        double F[3];
        node_force_and_momentum(c, m, sh.cp[lane] - sh.cn[lane], sh.T[lane], E, F, expr1);
        bar_velocity();
    }

    for (int z = z0; z < z1; ++z) {
        set_z(nb, c, z);
        const bool top = (z == c.NZ - 1);
        const bool wall = (z == 0) || top;
        gather27<MODE>(lin, nb, S);
        const double rho = sum27(S);
        double m[3];
        momentum(S, m);
        bar_moments();
        const double E[3] = {sh.E[0][lane], sh.E[1][lane], sh.E[2][lane]};
        const double dq = sh.cp[lane] - sh.cn[lane];
        double F[3], ex_[3], u[3];
        node_force_and_momentum(c, m, dq, sh.T[lane], E, F, ex_);
        const double rhoinv = 1.0 / rho;
        if (z == 0) {
            // u(z=0) = -(momentum expression of z=1) / rho(z=0)   (LBM.cu:778-800)
            u[0] = -rhoinv * expr1[0]; u[1] = -rhoinv * expr1[1]; u[2] = -rhoinv * expr1[2];
        } else {
            u[0] = rhoinv * ex_[0]; u[1] = rhoinv * ex_[1]; u[2] = rhoinv * ex_[2];
        }
        sh.u[0][lane] = u[0]; sh.u[1][lane] = u[1]; sh.u[2][lane] = u[2];
        bar_velocity();
        if (act) {
            const int i = nb.fc();
            a.dq[i] = dq;
            if (FULL) {  // LBM.cu:807-810
                a.fld[0][i] = rho; a.fld[1][i] = u[0]; a.fld[2][i] = u[1]; a.fld[3][i] = u[2];
            }
        }
        double wcr[4] = {c.w[0] * rho, c.w[1] * rho, c.w[2] * rho, c.w[3] * rho};
        const double omusq = 1.0 - 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) * c.inv_cs2;
        const double uF = u[0] * F[0] + u[1] * F[1] + u[2] * F[2];
        // rest population: TRT + source in the interior, frozen on the walls
        // (LBM.cu:502-504,1711,1861,1901)
        double O0 = S[0];
        if (!wall) O0 = S[0] - c.wp[0] * (S[0] - wcr[0] * omusq) + c.dt * (c.sp * (-c.coe[0] * uF));
        if (act) {
            if (MODE == EK_MODE_PUSH) lout[nb.lc()] = O0;
            else if (!wall) lout[nb.lc()] = O0;
        }
        FluidPairs<MODE, 0>::run(S, wcr, omusq, u, F, uF, wall, top, c, lout, nb, act);
    }
}

template <int MODE, bool FULL, bool EARR>
__global__ void __launch_bounds__(128, 4) ek_step_kernel(const __grid_constant__ StepArgs a)
{
    __shared__ Sh sh;
    const EkConst &c = a.c;
    const int lane = threadIdx.x & 31;
    const int role = threadIdx.x >> 5;
    int x = blockIdx.x * 32 + lane;
    const bool act = x < c.NX;
    if (!act) x = c.NX - 1;  // clamped duplicate: loads stay in bounds, stores are masked
    const int y = blockIdx.y;
    const int z0 = blockIdx.z * a.zchunk;
    const int z1 = min(z0 + a.zchunk, c.NZ);
    Nbr nb;
    set_xy(nb, c, x, y);
    const int pi = y * c.PX + x;
    if (role == 0) fluid_role<MODE, FULL>(a, sh, lane, act, nb, z0, z1);
    else scalar_role<MODE, FULL, EARR>(a, sh, role, lane, act, nb, pi, z0, z1);
}

// natural-layout export of the pre-collision state (tests, checkpoints)
template <int MODE>
__global__ void ek_export_kernel(const StepArgs a, int s, double *dst)
{
    const EkConst &c = a.c;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= c.NX) return;
    const int y = blockIdx.y, z = blockIdx.z;
    Nbr nb;
    set_xy(nb, c, x, y);
    set_z(nb, c, z);
    double S[27];
    const bool wall = (z == 0 || z == c.NZ - 1);
    if (s > 0 && wall) {
        const double *Wn = a.wall + (size_t)(s - 1) * 2 * 27 * c.plane + (size_t)(z == 0 ? 0 : 27) * c.plane + y * c.PX + x;
#pragma unroll
        for (int d = 0; d < 27; ++d) S[d] = Wn[(size_t)d * c.plane];
    } else {
        gather27<MODE>(a.in[s], nb, S);
    }
    const size_t cells = (size_t)c.NX * c.NY * c.NZ;
    const size_t o = (size_t)c.NX * ((size_t)c.NY * z + y) + x;
#pragma unroll
    for (int d = 0; d < 27; ++d) dst[(size_t)d * cells + o] = S[d];
}

template <int MODE>
cudaError_t launch_mode(const StepArgs &a, bool full, bool earr, dim3 grid, cudaStream_t st)
{
    if (full) {
        if (earr) ek_step_kernel<MODE, true, true><<<grid, 128, 0, st>>>(a);
        else ek_step_kernel<MODE, true, false><<<grid, 128, 0, st>>>(a);
    } else {
        if (earr) ek_step_kernel<MODE, false, true><<<grid, 128, 0, st>>>(a);
        else ek_step_kernel<MODE, false, false><<<grid, 128, 0, st>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t ek_launch_step(const StepArgs &a, int mode, bool write_fields, bool e_from_arrays, cudaStream_t st)
{
    const EkConst &c = a.c;
    dim3 grid((c.NX + 31) / 32, c.NY, (c.NZ + a.zchunk - 1) / a.zchunk);
    switch (mode) {
    case EK_MODE_AA_EVEN: return launch_mode<EK_MODE_AA_EVEN>(a, write_fields, e_from_arrays, grid, st);
    case EK_MODE_AA_ODD: return launch_mode<EK_MODE_AA_ODD>(a, write_fields, e_from_arrays, grid, st);
    default: return launch_mode<EK_MODE_PUSH>(a, write_fields, e_from_arrays, grid, st);
    }
}

cudaError_t ek_launch_export(const StepArgs &a, int mode, int set, double *dst, cudaStream_t st)
{
    const EkConst &c = a.c;
    dim3 block(64), grid((c.NX + 63) / 64, c.NY, c.NZ);
    if (mode == EK_MODE_AA_ODD) ek_export_kernel<EK_MODE_AA_ODD><<<grid, block, 0, st>>>(a, set, dst);
    else ek_export_kernel<EK_MODE_AA_EVEN><<<grid, block, 0, st>>>(a, set, dst);
    return cudaGetLastError();
}
