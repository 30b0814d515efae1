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
#include <stdio.h>
#include <stdlib.h>

#include "ek_lbm_common.cuh"

namespace {

struct Sh {
    double cp[32], cn[32], T[32];
    double E[3][32];
    double u[3][32];
#ifdef EK_XCHECK
    // five-warp variant: partial moments of fluid half A, and rho / F for it
    double rhoA[32], mpA[3][32], mnA[3][32];
    double rho[32], F[3][32];
#endif
};

template <int NT> __device__ __forceinline__ void bar_moments() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
template <int NT> __device__ __forceinline__ void bar_velocity() { asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory"); }

// ---------------------------------------------------------------------------
// scalar roles: cation (s = 1), anion (s = 2), temperature (s = 3)
// ---------------------------------------------------------------------------
// TRT relaxation of the opposite pair (d, d+1) of a scalar set (LBM.cu:1148-1845): shared by every
// node path, so that all of them produce the same bits
template <int d>
__device__ __forceinline__ void scalar_pair_trt(const double S[27], const double wcm[4], double omusq, double vtx,
                                                double vty, double vtz, double wp, double wmn, double &Oa, double &Ob)
{
    constexpr int o = d + 1, cls = ek_wclass(d);
    const double s_ = cdot<d>(vtx, vty, vtz);
    const double wm_ = wcm[cls];
    const double ep = wm_ * (omusq + 0.5 * s_ * s_);
    const double em = wm_ * s_;
    const double a = S[d], b = S[o];
    const double np_ = wp * (0.5 * (a + b) - ep);
    const double nm_ = wmn * (0.5 * (a - b) - em);
    Oa = a - (np_ + nm_);
    Ob = b - (np_ - nm_);
}

template <int MODE, bool LEAN, int p, int LROW = 0>
struct ScalarPairs {
    // TRT relaxation of the opposite pair (d, d+1), d = 2p+1 (LBM.cu:1148-1845),
    // then delivery of both results.
    static __device__ __forceinline__ void run(const double S[27], double wcm[4], double omusq, double vtx, double vty,
                                               double vtz, double wp, double wmn, bool wall, bool bottom, bool is_temp,
                                               const EkConst &c, double *lout, const Nbr &nb, const LeanAddr &la, int z,
                                               double *Wn, bool act)
    {
        constexpr int d = 2 * p + 1, o = d + 1, cls = ek_wclass(d);
        double Oa, Ob;
        scalar_pair_trt<d>(S, wcm, omusq, vtx, vty, vtz, wp, wmn, Oa, Ob);
        if (act) {
            if (LEAN) {
                putx<MODE, d, true, LROW>(lout, nb, la, Oa);
                putx<MODE, o, true, LROW>(lout, nb, la, Ob);
            } else if (!wall) {
                if (MODE == EK_MODE_AA_EVEN) {
                    put<MODE, d>(lout, nb, Oa);
                    put<MODE, o>(lout, nb, Ob);
                } else {
                    // inflow into wall nodes is discarded (LBM.cu:2102-2218 overwrite it)
                    // z is interior here: the target z +- 1 is a wall plane only next to a wall
                    const bool near_bottom = (z == 1), near_top = (z == c.NZ - 2);
                    const bool drop_a = ek_cz(d) > 0 ? near_top : (ek_cz(d) < 0 ? near_bottom : false);
                    const bool drop_b = ek_cz(d) > 0 ? near_bottom : (ek_cz(d) < 0 ? near_top : false);
                    if (!drop_a) put<MODE, d>(lout, nb, Oa);
                    if (!drop_b) put<MODE, o>(lout, nb, Ob);
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
        ScalarPairs<MODE, LEAN, p + 1, LROW>::run(S, wcm, omusq, vtx, vty, vtz, wp, wmn, wall, bottom, is_temp, c, lout, nb,
                                                  la, z, Wn, act);
    }
};
template <int MODE, bool LEAN, int LROW>
struct ScalarPairs<MODE, LEAN, 13, LROW> {
    static __device__ __forceinline__ void run(const double *, double *, double, double, double, double, double, double,
                                               bool, bool, bool, const EkConst &, double *, const Nbr &,
                                               const LeanAddr &, int, double *, bool) {}
};

// one node of a scalar set.  LEAN: deep interior (2 <= z <= NZ-3), see LeanAddr.
template <int MODE, bool FULL, bool EARR, int NT, bool LEAN, int LROW = 0>
__device__ __forceinline__ void scalar_node(const StepArgs &a, Sh &sh, const int s, const int lane, const bool act,
                                            const int x, const int y, LeanAddr &la, double *W, double *mom_sh,
                                            const int z)
{
    const EkConst &c = a.c;
    Nbr nb;
    if (!LEAN) set_xy(nb, c, x, y);
    double *lin = a.in[s];
    double *lout = a.out[s];
    const double wp = c.wp[s], wmn = c.wm[s];
    const bool is_temp = (s == 3);
    const double Ks = s == 1 ? c.K : c.Kn;
    double S[27];
    if (LEAN) { lean_set_z(la, lin, c, z); if (LROW > 0) lean_set_pim(la); } else set_z(nb, c, z);
    const bool bottom = LEAN ? false : (z == 0);
    const bool wall = LEAN ? false : (bottom || (z == c.NZ - 1));
    double *Wn = W + (size_t)(bottom ? 0 : 27) * c.plane;
    if (LEAN) {
        gather27_lean<MODE, LROW>(la, S);
#ifdef EK_ODD_PREFETCH
        if (MODE == EK_MODE_AA_ODD && z + 1 < la.zend) {
            if (LROW > 0) prefetch27_lean_odd_imm<LROW>(la, c.lplane); else prefetch27_lean_odd(la, c.lplane);
        }
#ifndef EK_NO_EVEN_PREFETCH   // even launch 4.52 -> 4.43 ms at 256^3
        if (MODE == EK_MODE_AA_EVEN && z + 1 < la.zend) prefetch27_lean_even(la, c.lplane);
#endif
#endif
    } else if (wall) {
#pragma unroll
        for (int d = 0; d < 27; ++d) S[d] = Wn[(size_t)d * c.plane];
    } else {
        gather27<MODE>(lin, nb, S);
    }
    const double m = sum27(S);
    double E[3] = {0.0, 0.0, 0.0};
    mom_sh[lane] = m;
    if (is_temp) {
        if (LEAN) efield_lean(a, la, z, E); else efield_at<EARR>(a, nb, z, E);
        sh.E[0][lane] = E[0]; sh.E[1][lane] = E[1]; sh.E[2][lane] = E[2];
    }
    if (FULL && act) a.fld[3 + s][LEAN ? (size_t)z * c.plane + la.fc : (size_t)nb.fc()] = m;  // charge, chargen, T (LBM.cu:811-813)
    bar_moments<NT>();
    // E is read between the two barriers: the temperature warp may only
    // overwrite it after every warp has passed bar_velocity()
    if (!is_temp) { E[0] = sh.E[0][lane]; E[1] = sh.E[1][lane]; E[2] = sh.E[2][lane]; }
    bar_velocity<NT>();
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
        if (LEAN && LROW > 0) la.pim[1][1][0] = O0;
        else if (LEAN) lean_ptr(la.b[1], la.oxy[1][1])[0] = O0;
        else if (!wall) lout[nb.lc()] = O0;
        else if (!is_temp) Wn[0] = O0;
        else if (bottom) Wn[0] = -O0 + c.twoTw[0];
        else Wn[0] = -O0;
    }
    ScalarPairs<MODE, LEAN, 0, LROW>::run(S, wcm, omusq, vtx, vty, vtz, wp, wmn, wall, bottom, is_temp, c, lout, nb, la, z,
                                          Wn, act);
}

template <int MODE, bool FULL, bool EARR, int NT, bool LEANOK = false, int LROW = 0>
__device__ __forceinline__ void scalar_role(const StepArgs &a, Sh &sh, const int s, const int lane, const bool act,
                                            const int x, const int y, const int pi, const int z0, const int z1)
{
    const EkConst &c = a.c;
    double *W = a.wall + (size_t)(s - 1) * 2 * 27 * c.plane + pi;
    const bool is_temp = (s == 3);
    double *mom_sh = s == 1 ? sh.cp : (s == 2 ? sh.cn : sh.T);
    LeanAddr la;
    if (LEANOK) {
        Nbr nb;
        set_xy(nb, c, x, y);
        lean_init(la, nb);
        la.zend = z1 < c.NZ - 2 ? z1 : c.NZ - 2;
        la.pf_m = !(x == 0 && c.xlo >= c.NX);       // x-1 / x+1 is a ghost column of an x-slab
        la.pf_p = !(x == c.NX - 1 && c.xhi >= c.NX);
    }

    if (z0 == 0) {
        // the bottom wall needs the moments of the z = 1 node first (LBM.cu:663-801)
        double S[27];
        Nbr nb;
        set_xy(nb, c, x, y);
        set_z(nb, c, 1);
        gather27<MODE>(a.in[s], nb, S);
        mom_sh[lane] = sum27(S);
        if (is_temp) {
            double E[3];
            efield_at<EARR>(a, nb, 1, E);
            sh.E[0][lane] = E[0]; sh.E[1][lane] = E[1]; sh.E[2][lane] = E[2];
        }
        bar_moments<NT>();
        bar_velocity<NT>();
    }

    // rows y-1..y+1 without periodic wrap: the row stride is an immediate (LROW).  Two copies of the loop, so
    // that the register allocation of each is its own (the immediate form needs 3 of the 9 column offsets)
    if (LROW > 0 && y > 0 && y < c.NY - 1) {
        for (int z = z0; z < z1; ++z) {
            if (LEANOK && z >= 2 && z < c.NZ - 2) scalar_node<MODE, FULL, false, NT, true, LROW>(a, sh, s, lane, act, x, y, la, W, mom_sh, z);
            else scalar_node<MODE, FULL, EARR, NT, false>(a, sh, s, lane, act, x, y, la, W, mom_sh, z);
        }
    } else {
        for (int z = z0; z < z1; ++z) {
            if (LEANOK && z >= 2 && z < c.NZ - 2) scalar_node<MODE, FULL, false, NT, true>(a, sh, s, lane, act, x, y, la, W, mom_sh, z);
            else scalar_node<MODE, FULL, EARR, NT, false>(a, sh, s, lane, act, x, y, la, W, mom_sh, z);
        }
    }
}

// ---------------------------------------------------------------------------
// fluid role
// ---------------------------------------------------------------------------
// TRT + Guo forcing of the opposite pair (d, d+1) of the fluid set, interior nodes
template <int d>
__device__ __forceinline__ void fluid_pair_trt(const double S[27], const double wcr[4], double omusq, const double u[3],
                                               const double F[3], double uF, const EkConst &c, double &Oa, double &Ob)
{
    constexpr int o = d + 1, cls = ek_wclass(d);
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
}

template <int MODE, bool LEAN, int p, int LROW = 0>
struct FluidPairs {
    static __device__ __forceinline__ void run(const double S[27], double wcr[4], double omusq, const double u[3],
                                               const double F[3], double uF, bool wall, bool top, const EkConst &c,
                                               double *lout, const Nbr &nb, const LeanAddr &la, bool act)
    {
        constexpr int d = 2 * p + 1, o = d + 1, cls = ek_wclass(d);
        double Oa, Ob;
        if (LEAN || !wall) {
            fluid_pair_trt<d>(S, wcr, omusq, u, F, uF, c, Oa, Ob);
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
            putx<MODE, d, LEAN, LROW>(lout, nb, la, Oa);
            putx<MODE, o, LEAN, LROW>(lout, nb, la, Ob);
        }
        FluidPairs<MODE, LEAN, p + 1, LROW>::run(S, wcr, omusq, u, F, uF, wall, top, c, lout, nb, la, act);
    }
};
template <int MODE, bool LEAN, int LROW>
struct FluidPairs<MODE, LEAN, 13, LROW> {
    static __device__ __forceinline__ void run(const double *, double *, double, const double *, const double *, double,
                                               bool, bool, const EkConst &, double *, const Nbr &, const LeanAddr &,
                                               bool) {}
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

// one node of the fluid set.  LEAN: deep interior (2 <= z <= NZ-3), see LeanAddr.
template <int MODE, bool FULL, int NT, bool LEAN, int LROW = 0>
__device__ __forceinline__ void fluid_node(const StepArgs &a, Sh &sh, const int lane, const bool act, const int x,
                                           const int y, LeanAddr &la, const double expr1[3], const int z)
{
    const EkConst &c = a.c;
    Nbr nb;
    if (!LEAN) set_xy(nb, c, x, y);
    double *lin = a.in[0];
    double *lout = a.out[0];
    double S[27];
    const bool top = LEAN ? false : (z == c.NZ - 1);
    const bool wall = LEAN ? false : ((z == 0) || top);
    if (LEAN) {
        lean_set_z(la, lin, c, z);
        if (LROW > 0) lean_set_pim(la);
        gather27_lean<MODE, LROW>(la, S);
#ifdef EK_ODD_PREFETCH
        if (MODE == EK_MODE_AA_ODD && z + 1 < la.zend) {
            if (LROW > 0) prefetch27_lean_odd_imm<LROW>(la, c.lplane); else prefetch27_lean_odd(la, c.lplane);
        }
#ifndef EK_NO_EVEN_PREFETCH   // even launch 4.52 -> 4.43 ms at 256^3
        if (MODE == EK_MODE_AA_EVEN && z + 1 < la.zend) prefetch27_lean_even(la, c.lplane);
#endif
#endif
    } else {
        set_z(nb, c, z);
        gather27<MODE>(lin, nb, S);
    }
    const double rho = sum27(S);
    double m[3];
    momentum(S, m);
    // 1/rho does not depend on the other sets: its latency overlaps the momentum sums and the wait
    // at the barrier instead of sitting between the two barriers where three warps wait for it
    const double rhoinv = 1.0 / rho;
    bar_moments<NT>();
    const double E[3] = {sh.E[0][lane], sh.E[1][lane], sh.E[2][lane]};
    const double dq = sh.cp[lane] - sh.cn[lane];
    double F[3], ex_[3], u[3];
    node_force_and_momentum(c, m, dq, sh.T[lane], E, F, ex_);
    if (!LEAN && z == 0) {
        // u(z=0) = -(momentum expression of z=1) / rho(z=0)   (LBM.cu:778-800)
        u[0] = -rhoinv * expr1[0]; u[1] = -rhoinv * expr1[1]; u[2] = -rhoinv * expr1[2];
    } else {
        u[0] = rhoinv * ex_[0]; u[1] = rhoinv * ex_[1]; u[2] = rhoinv * ex_[2];
    }
    sh.u[0][lane] = u[0]; sh.u[1][lane] = u[1]; sh.u[2][lane] = u[2];
    bar_velocity<NT>();
    if (act) {
        if (LEAN) {
            const size_t i = (size_t)z * c.plane + la.fc;
            a.dq[(size_t)z * c.dq_sz + la.fdq] = dq;
            if (FULL) {  // LBM.cu:807-810
                a.fld[0][i] = rho; a.fld[1][i] = u[0]; a.fld[2][i] = u[1]; a.fld[3][i] = u[2];
            }
        } else {
            const int i = nb.fc();
            a.dq[dq_at(c, nb, z)] = dq;
            if (FULL) {  // LBM.cu:807-810
                a.fld[0][i] = rho; a.fld[1][i] = u[0]; a.fld[2][i] = u[1]; a.fld[3][i] = u[2];
            }
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
        if (LEAN && LROW > 0) la.pim[1][1][0] = O0;
        else if (LEAN) lean_ptr(la.b[1], la.oxy[1][1])[0] = O0;
        else if (MODE == EK_MODE_PUSH) lout[nb.lc()] = O0;
        else if (!wall) lout[nb.lc()] = O0;
    }
    FluidPairs<MODE, LEAN, 0, LROW>::run(S, wcr, omusq, u, F, uF, wall, top, c, lout, nb, la, act);
}

template <int MODE, bool FULL, int NT, bool LEANOK = false, int LROW = 0>
__device__ __forceinline__ void fluid_role(const StepArgs &a, Sh &sh, const int lane, const bool act, const int x,
                                           const int y, const int z0, const int z1)
{
    const EkConst &c = a.c;
    double expr1[3] = {0.0, 0.0, 0.0};
    LeanAddr la;
    if (LEANOK) {
        Nbr nb;
        set_xy(nb, c, x, y);
        lean_init(la, nb);
        la.fdq = (long long)y * c.dq_sy + x;
        la.zend = z1 < c.NZ - 2 ? z1 : c.NZ - 2;
        la.pf_m = !(x == 0 && c.xlo >= c.NX);       // x-1 / x+1 is a ghost column of an x-slab
        la.pf_p = !(x == c.NX - 1 && c.xhi >= c.NX);
    }

    if (z0 == 0) {
        double S[27];
        Nbr nb;
        set_xy(nb, c, x, y);
        set_z(nb, c, 1);
        gather27<MODE>(a.in[0], nb, S);
        double m[3];
        momentum(S, m);
        bar_moments<NT>();
        const double E[3] = {sh.E[0][lane], sh.E[1][lane], sh.E[2][lane]};
        double F[3];
        node_force_and_momentum(c, m, sh.cp[lane] - sh.cn[lane], sh.T[lane], E, F, expr1);
        bar_velocity<NT>();
    }

    if (LROW > 0 && y > 0 && y < c.NY - 1) {
        for (int z = z0; z < z1; ++z) {
            if (LEANOK && z >= 2 && z < c.NZ - 2) fluid_node<MODE, FULL, NT, true, LROW>(a, sh, lane, act, x, y, la, expr1, z);
            else fluid_node<MODE, FULL, NT, false>(a, sh, lane, act, x, y, la, expr1, z);
        }
    } else {
        for (int z = z0; z < z1; ++z) {
            if (LEANOK && z >= 2 && z < c.NZ - 2) fluid_node<MODE, FULL, NT, true>(a, sh, lane, act, x, y, la, expr1, z);
            else fluid_node<MODE, FULL, NT, false>(a, sh, lane, act, x, y, la, expr1, z);
        }
    }
}

// LEAN: deep-interior planes take the lean node path (A-A modes only)
#ifndef EK_MIN_CTAS
#define EK_MIN_CTAS 4   // 128 registers; 3 (168 registers, no rematerialisation) measured slower, DESIGN.md 3.6
#endif
// (a variant without the store predicate for NX % 32 == 0 was measured: the odd kernel then spills 128 B
// and the pass takes 5.01 instead of 4.92 ms at 256^3)
// LROW > 0: the row stride of the lattice (NXT * 27 * 32 elements) as a compile-time constant, for the odd A-A
// step of the deep-interior planes (see LeanAddr::pim)
template <int MODE, bool FULL, bool EARR, bool LEAN, int LROW = 0>
__global__ void __launch_bounds__(128, EK_MIN_CTAS) ek_step_kernel(const __grid_constant__ StepArgs a)
{
    __shared__ Sh sh;
    const EkConst &c = a.c;
    const int lane = threadIdx.x & 31;
    // broadcast through a shuffle: tells the compiler that the role is warp-uniform, so that the role's
    // loop counters, plane bases and branches can live in the uniform datapath
    const int role = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    // x-tile of this CTA: all tiles, or -- slab pipeline -- the two boundary tiles of the row first, so that the
    // halo exchange travels under the launch of the interior tiles
    const int xtile = a.xt_mode == 0 ? (int)blockIdx.x
                    : (a.xt_mode == 1 ? (blockIdx.x == 0 ? 0 : (c.NX + 31) / 32 - 1) : (int)blockIdx.x + 1);
    int x = xtile * 32 + lane;
    const bool act = x < c.NX;
    if (!act) x = c.NX - 1;  // clamped duplicate: loads stay in bounds, stores are masked
    const int y = blockIdx.y;
    const int z0 = (blockIdx.z + a.zblock0) * a.zchunk;
    const int z1 = min(z0 + a.zchunk, c.NZ);
    const int pi = y * c.PX + x;
    if (role == 0) fluid_role<MODE, FULL, 128, LEAN, LROW>(a, sh, lane, act, x, y, z0, z1);
    else scalar_role<MODE, FULL, EARR, 128, LEAN, LROW>(a, sh, role, lane, act, x, y, pi, z0, z1);
}

#ifdef EK_XCHECK
// ---------------------------------------------------------------------------
// Five warps per 32 cells: the fluid set is split over two warps (half A: rest +
// pairs 1..6, half B: pairs 7..13) because the fluid warp of the four-warp kernel
// carries ~1.75x the instructions of a scalar warp and everybody waits for it at
// the barriers (ncu: 43 % of the warp stalls).  The sums are CHAINED from A to B
// in the reference's order (LBM.cu:621-644), so the results are bit-identical to
// the four-warp kernel.
// ---------------------------------------------------------------------------
template <int MODE, bool FULL, int NT>
__device__ __forceinline__ void fluid_half_a(const StepArgs &a, Sh &sh, const int lane, const bool act, Nbr &nb,
                                             const int z0, const int z1)
{
    const EkConst &c = a.c;
    double *lin = a.in[0];
    double *lout = a.out[0];
    double S[13];
    for (int zz = (z0 == 0 ? -1 : z0); zz < z1; ++zz) {
        const bool pre = zz < 0;
        const int z = pre ? 1 : zz;
        set_z(nb, c, z);
        const bool top = (z == c.NZ - 1);
        const bool wall = (z == 0) || top;
        gather_half<MODE, 0>(lin, nb, S);
        sh.rhoA[lane] = sum_half<13>(S);
        sh.mpA[0][lane] = S[1] + S[7] + S[9];   sh.mnA[0][lane] = S[2] + S[8] + S[10];
        sh.mpA[1][lane] = S[3] + S[7] + S[11];  sh.mnA[1][lane] = S[4] + S[8] + S[12];
        sh.mpA[2][lane] = S[5] + S[9] + S[11];  sh.mnA[2][lane] = S[6] + S[10] + S[12];
        bar_moments<NT>();
        bar_velocity<NT>();
        if (pre) continue;
        const double rho = sh.rho[lane];
        const double u[3] = {sh.u[0][lane], sh.u[1][lane], sh.u[2][lane]};
        const double F[3] = {sh.F[0][lane], sh.F[1][lane], sh.F[2][lane]};
        if (FULL && act) {  // LBM.cu:807-810
            const int i = nb.fc();
            a.fld[0][i] = rho; a.fld[1][i] = u[0]; a.fld[2][i] = u[1]; a.fld[3][i] = u[2];
        }
        const double wcr[4] = {c.w[0] * rho, c.w[1] * rho, c.w[2] * rho, c.w[3] * rho};
        const double omusq = 1.0 - 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) * c.inv_cs2;
        const double uF = u[0] * F[0] + u[1] * F[1] + u[2] * F[2];
        double O0 = S[0];
        if (!wall) O0 = S[0] - c.wp[0] * (S[0] - wcr[0] * omusq) + c.dt * (c.sp * (-c.coe[0] * uF));
        if (act) {
            if (MODE == EK_MODE_PUSH) lout[nb.lc()] = O0;
            else if (!wall) lout[nb.lc()] = O0;
        }
        FluidPairs8<MODE, 0, 0>::run(S, wcr, omusq, u, F, uF, wall, top, c, lout, nb, act);
    }
}

template <int MODE, bool FULL, int NT>
__device__ __forceinline__ void fluid_half_b(const StepArgs &a, Sh &sh, const int lane, const bool act, Nbr &nb,
                                             const int z0, const int z1)
{
    const EkConst &c = a.c;
    double *lin = a.in[0];
    double *lout = a.out[0];
    double S[14];
    double expr1[3] = {0.0, 0.0, 0.0};
#define F_(d) S[(d) - 13]
    for (int zz = (z0 == 0 ? -1 : z0); zz < z1; ++zz) {
        const bool pre = zz < 0;
        const int z = pre ? 1 : zz;
        set_z(nb, c, z);
        const bool top = (z == c.NZ - 1);
        const bool wall = (z == 0) || top;
        gather_half<MODE, 1>(lin, nb, S);
        bar_moments<NT>();
        // continue half A's chains in the reference's order (LBM.cu:621-644)
        double rho = sh.rhoA[lane];
#pragma unroll
        for (int i = 0; i < 14; ++i) rho = rho + S[i];
        double m[3];
        m[0] = (sh.mpA[0][lane] + F_(13) + F_(15) + F_(19) + F_(21) + F_(23) + F_(26)
              - (sh.mnA[0][lane] + F_(14) + F_(16) + F_(20) + F_(22) + F_(24) + F_(25)));
        m[1] = (sh.mpA[1][lane] + F_(14) + F_(17) + F_(19) + F_(21) + F_(24) + F_(25)
              - (sh.mnA[1][lane] + F_(13) + F_(18) + F_(20) + F_(22) + F_(23) + F_(26)));
        m[2] = (sh.mpA[2][lane] + F_(16) + F_(18) + F_(19) + F_(22) + F_(23) + F_(25)
              - (sh.mnA[2][lane] + F_(15) + F_(17) + F_(20) + F_(21) + F_(24) + F_(26)));
        const double E[3] = {sh.E[0][lane], sh.E[1][lane], sh.E[2][lane]};
        const double dq = sh.cp[lane] - sh.cn[lane];
        double F[3], ex_[3], u[3];
        node_force_and_momentum(c, m, dq, sh.T[lane], E, F, ex_);
        if (pre) {
            expr1[0] = ex_[0]; expr1[1] = ex_[1]; expr1[2] = ex_[2];
            bar_velocity<NT>();
            continue;
        }
        const double rhoinv = 1.0 / rho;
        if (z == 0) {
            u[0] = -rhoinv * expr1[0]; u[1] = -rhoinv * expr1[1]; u[2] = -rhoinv * expr1[2];
        } else {
            u[0] = rhoinv * ex_[0]; u[1] = rhoinv * ex_[1]; u[2] = rhoinv * ex_[2];
        }
        sh.u[0][lane] = u[0]; sh.u[1][lane] = u[1]; sh.u[2][lane] = u[2];
        sh.rho[lane] = rho;
        sh.F[0][lane] = F[0]; sh.F[1][lane] = F[1]; sh.F[2][lane] = F[2];
        bar_velocity<NT>();
        if (act) a.dq[dq_at(c, nb, z)] = dq;
        const double wcr[4] = {c.w[0] * rho, c.w[1] * rho, c.w[2] * rho, c.w[3] * rho};
        const double omusq = 1.0 - 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) * c.inv_cs2;
        const double uF = u[0] * F[0] + u[1] * F[1] + u[2] * F[2];
        FluidPairs8<MODE, 1, 6>::run(S, wcr, omusq, u, F, uF, wall, top, c, lout, nb, act);
    }
#undef F_
}

template <int MODE, bool FULL, bool EARR>
__global__ void __launch_bounds__(160, 3) ek_step5_kernel(const __grid_constant__ StepArgs a)
{
    __shared__ Sh sh;
    const EkConst &c = a.c;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int x = blockIdx.x * 32 + lane;
    const bool act = x < c.NX;
    if (!act) x = c.NX - 1;
    const int y = blockIdx.y;
    const int z0 = (blockIdx.z + a.zblock0) * a.zchunk;
    const int z1 = min(z0 + a.zchunk, c.NZ);
    Nbr nb;
    set_xy(nb, c, x, y);
    const int pi = y * c.PX + x;
    if (warp == 0) fluid_half_a<MODE, FULL, 160>(a, sh, lane, act, nb, z0, z1);
    else if (warp == 1) fluid_half_b<MODE, FULL, 160>(a, sh, lane, act, nb, z0, z1);
    else scalar_role<MODE, FULL, EARR, 160>(a, sh, warp - 1, lane, act, x, y, pi, z0, z1);
}

#endif  // EK_XCHECK

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

// the inverse of the export for the natural layout (A-A parity 0 / current push lattice):
// restores a checkpointed pre-collision state, wall-node side buffers included
__global__ void ek_import_kernel(const StepArgs a, int s, const double *src)
{
    const EkConst &c = a.c;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= c.NX) return;
    const int y = blockIdx.y, z = blockIdx.z;
    const size_t cells = (size_t)c.NX * c.NY * c.NZ;
    const size_t o = (size_t)c.NX * ((size_t)c.NY * z + y) + x;
    const bool wall = (z == 0 || z == c.NZ - 1);
    double *lat = a.in[s] + (size_t)z * c.lplane + (size_t)y * c.lrow + ek_lat_col(x);
    double *Wn = nullptr;
    if (s > 0 && wall)
        Wn = a.wall + (size_t)(s - 1) * 2 * 27 * c.plane + (size_t)(z == 0 ? 0 : 27) * c.plane + y * c.PX + x;
#pragma unroll
    for (int d = 0; d < 27; ++d) {
        const double v = src[(size_t)d * cells + o];
        lat[(size_t)d * EK_TILE] = v;
        if (Wn) Wn[(size_t)d * c.plane] = v;
    }
}

#ifdef EK_XCHECK
template <int MODE>
cudaError_t launch_mode5(const StepArgs &a, bool full, bool earr, dim3 grid, cudaStream_t st)
{
    if (full) {
        if (earr) ek_step5_kernel<MODE, true, true><<<grid, 160, 0, st>>>(a);
        else ek_step5_kernel<MODE, true, false><<<grid, 160, 0, st>>>(a);
    } else {
        if (earr) ek_step5_kernel<MODE, false, true><<<grid, 160, 0, st>>>(a);
        else ek_step5_kernel<MODE, false, false><<<grid, 160, 0, st>>>(a);
    }
    return cudaGetLastError();
}
#endif  // EK_XCHECK

#ifdef EK_XCHECK
// ---------------------------------------------------------------------------
// x-marching variant of the ODD A-A step for the deep-interior planes.
//
// In the odd step a node writes slot d at x + c_d: for the 18 directions with c_x = +-1 a warp's
// 256-byte store starts 8 bytes off a sector boundary and its last element lands in the NEXT x-tile
// (6912 bytes away): 9 sectors, two of them partial, per request.  Measured at 256^3 (DESIGN.md
// 3.6): 0.31 ms of the 5.28 ms launch, the largest single loss against the even step.
//
// (Cross-check build only: both marching kernels are measured SLOWER than the z-walking default, DESIGN.md 3.7;
// they are kept, bit-identical and tested, as the record of that experiment.)
// Here a CTA (still one warp per population set) owns one (y, z) ROW and walks its x-tiles:
//   * c_x = +1 outputs: lane i's value belongs to column i+1 -> one shuffle up; lane 0 takes what
//     lane 31 produced for the previous tile (9 doubles per warp carried through shared memory);
//     the store is then the full aligned 256-byte segment of tile T.
//   * c_x = -1 outputs: lane i's value belongs to column i-1 and column 31 of tile T comes from lane 0
//     of tile T+1 -> the warp parks the 9 values in a two-tile ring in shared memory and writes
//     tile T-1 as full aligned segments one iteration later.
// Only the two ends of a row (periodic wrap, or the ghost columns of an x-slab) remain single-element
// stores: 2 per row and direction instead of 2 per tile.  The loads keep the neighbour gather of the
// lean path.  Same arithmetic (scalar_pair_trt / fluid_pair_trt), same bits; requires NX % 32 == 0.
// ---------------------------------------------------------------------------
struct MarchSm {                 // per warp
    double ring[9][64];          // c_x = -1 outputs of the tiles T-1 / T (halves alternate)
    double carry[2][9];          // lane 31's c_x = +1 outputs of tile T (index T & 1)
};

// index of direction d among the nine directions with the same c_x sign
__host__ __device__ constexpr int ek_xrank(int d)
{
    int k = 0;
    for (int e = 1; e < d; ++e)
        if (ek_cx(e) == ek_cx(d)) ++k;
    return k;
}

struct MarchCtx {
    LeanAddr la;
    unsigned ly[3];              // lattice offsets of the rows y-1, y, y+1 within a plane
    int lane, T, NT;
    MarchSm *ms;
};

__device__ __forceinline__ void march_tile(MarchCtx &m, const EkConst &c, int y)
{
    // neighbour columns of x = 32 T + lane; the row ends wrap (or reach the ghost columns of a slab)
    const int x = m.T * 32 + m.lane;
    const unsigned ox1 = (unsigned)m.T * EK_TILE_ELEMS + (unsigned)m.lane;
    const unsigned ox0 = m.lane > 0 ? ox1 - 1u : (m.T > 0 ? ox1 - (unsigned)(EK_TILE_ELEMS - 31) : ek_lat_col(c.xlo));
    const unsigned ox2 = m.lane < 31 ? ox1 + 1u : (m.T < m.NT - 1 ? ox1 + (unsigned)(EK_TILE_ELEMS - 31) : ek_lat_col(c.xhi));
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        m.la.oxy[j][0] = m.ly[j] + ox0;
        m.la.oxy[j][1] = m.ly[j] + ox1;
        m.la.oxy[j][2] = m.ly[j] + ox2;
    }
    const int ym = y == 0 ? c.NY - 1 : y - 1, yp = y == c.NY - 1 ? 0 : y + 1;
    m.la.fc = y * c.PX + x;
    m.la.fxm = y * c.PX + (x == 0 ? c.xlo : x - 1);
    m.la.fxp = y * c.PX + (x == c.NX - 1 ? c.xhi : x + 1);
    m.la.fym = ym * c.PX + x;
    m.la.fyp = yp * c.PX + x;
    m.la.fdq = (long long)y * c.dq_sy + x;
}

// deliver the post-collision value of direction d (odd A-A step: slot d of the node at x + c_d)
template <int d>
__device__ __forceinline__ void march_put(const MarchCtx &m, double v)
{
    double *row = m.la.b[1 + ek_cz(d)];
    if (ek_cx(d) == 0) {
        EK_ST(row + m.la.oxy[1 + ek_cy(d)][1] + d * EK_TILE, v);
    } else if (ek_cx(d) > 0) {
        constexpr int k = ek_xrank(d);
        double w = __shfl_up_sync(0xffffffffu, v, 1);
        if (m.lane == 31) {
            m.ms->carry[m.T & 1][k] = v;
            if (m.T == m.NT - 1) EK_ST(row + m.la.oxy[1 + ek_cy(d)][2] + d * EK_TILE, v);   // end of the row: column xhi
        }
        if (m.lane == 0 && m.T > 0) w = m.ms->carry[(m.T & 1) ^ 1][k];
        if (m.lane > 0 || m.T > 0) EK_ST(row + m.la.oxy[1 + ek_cy(d)][1] + d * EK_TILE, w);
    } else {
        constexpr int k = ek_xrank(d);
        m.ms->ring[k][(m.T & 1) * 32 + m.lane] = v;
        if (m.lane == 0 && m.T == 0) EK_ST(row + m.la.oxy[1 + ek_cy(d)][0] + d * EK_TILE, v);   // start of the row: column xlo
    }
}

// write the parked c_x = -1 outputs of tile `T - 1` (all 32 columns), or -- last = true, after the
// loop -- of the final tile (columns 0..30; column 31 was written by the first iteration / is a ghost)
template <int d>
struct MarchDrain {
    static __device__ __forceinline__ void run(const MarchCtx &m, bool last)
    {
        if (ek_cx(d) < 0) {
            constexpr int k = ek_xrank(d);
            const int half = last ? (m.T & 1) : ((m.T & 1) ^ 1);
            const double w = m.ms->ring[k][(half * 32 + m.lane + 1) & 63];
            double *q = m.la.b[1 + ek_cz(d)] + m.la.oxy[1 + ek_cy(d)][1] + d * EK_TILE;
            if (!last) EK_ST(q - EK_TILE_ELEMS, w);
            else if (m.lane < 31) EK_ST(q, w);
        }
        MarchDrain<d + 1>::run(m, last);
    }
};
template <>
struct MarchDrain<27> {
    static __device__ __forceinline__ void run(const MarchCtx &, bool) {}
};

template <int p>
struct MarchScalarPairs {
    static __device__ __forceinline__ void run(const double S[27], const double wcm[4], double omusq, double vtx, double vty,
                                               double vtz, double wp, double wmn, const MarchCtx &m)
    {
        constexpr int d = 2 * p + 1, o = d + 1;
        double Oa, Ob;
        scalar_pair_trt<d>(S, wcm, omusq, vtx, vty, vtz, wp, wmn, Oa, Ob);
        march_put<d>(m, Oa);
        march_put<o>(m, Ob);
        MarchScalarPairs<p + 1>::run(S, wcm, omusq, vtx, vty, vtz, wp, wmn, m);
    }
};
template <>
struct MarchScalarPairs<13> {
    static __device__ __forceinline__ void run(const double *, const double *, double, double, double, double, double,
                                               double, const MarchCtx &) {}
};

template <int p>
struct MarchFluidPairs {
    static __device__ __forceinline__ void run(const double S[27], const double wcr[4], double omusq, const double u[3],
                                               const double F[3], double uF, const EkConst &c, const MarchCtx &m)
    {
        constexpr int d = 2 * p + 1, o = d + 1;
        double Oa, Ob;
        fluid_pair_trt<d>(S, wcr, omusq, u, F, uF, c, Oa, Ob);
        march_put<d>(m, Oa);
        march_put<o>(m, Ob);
        MarchFluidPairs<p + 1>::run(S, wcr, omusq, u, F, uF, c, m);
    }
};
template <>
struct MarchFluidPairs<13> {
    static __device__ __forceinline__ void run(const double *, const double *, double, const double *, const double *,
                                               double, const EkConst &, const MarchCtx &) {}
};

__device__ __forceinline__ void march_begin(MarchCtx &m, const EkConst &c, double *lat, MarchSm *ms, int lane, int y, int z)
{
    m.lane = lane;
    m.NT = c.NX >> 5;
    m.ms = ms;
    const int ym = y == 0 ? c.NY - 1 : y - 1, yp = y == c.NY - 1 ? 0 : y + 1;
    m.ly[0] = (unsigned)ym * c.lrow; m.ly[1] = (unsigned)y * c.lrow; m.ly[2] = (unsigned)yp * c.lrow;
    lean_set_z(m.la, lat, c, z);
}

// the x-row (y, z) of a scalar set; the arithmetic is scalar_node<.., LEAN = true>
template <bool FULL, int NT_>
__device__ __forceinline__ void march_scalar_row(const StepArgs &a, Sh &sh, MarchSm *ms, const int s, const int lane,
                                                 const int y, const int z)
{
    const EkConst &c = a.c;
    const double wp = c.wp[s], wmn = c.wm[s];
    const bool is_temp = (s == 3);
    const double Ks = s == 1 ? c.K : c.Kn;
    double *mom_sh = s == 1 ? sh.cp : (s == 2 ? sh.cn : sh.T);
    MarchCtx m;
    march_begin(m, c, a.in[s], ms, lane, y, z);
    for (m.T = 0; m.T < m.NT; ++m.T) {
        march_tile(m, c, y);
        double S[27];
        gather27_lean<EK_MODE_AA_ODD>(m.la, S);
        const double mm = sum27(S);
        double E[3] = {0.0, 0.0, 0.0};
        mom_sh[lane] = mm;
        if (is_temp) {
            efield_lean(a, m.la, z, E);
            sh.E[0][lane] = E[0]; sh.E[1][lane] = E[1]; sh.E[2][lane] = E[2];
        }
        if (FULL) a.fld[3 + s][(size_t)z * c.plane + m.la.fc] = mm;   // charge, chargen, T (LBM.cu:811-813)
        bar_moments<NT_>();
        if (!is_temp) { E[0] = sh.E[0][lane]; E[1] = sh.E[1][lane]; E[2] = sh.E[2][lane]; }
        bar_velocity<NT_>();
        double vx = sh.u[0][lane], vy = sh.u[1][lane], vz = sh.u[2][lane];
        if (!is_temp) {
            vx = vx + Ks * E[0];
            vy = vy + Ks * E[1];
            vz = vz + Ks * E[2];
        }
        double wcm[4] = {c.w[0] * mm, c.w[1] * mm, c.w[2] * mm, c.w[3] * mm};
        const double omusq = 1.0 - 0.5 * (vx * vx + vy * vy + vz * vz) * c.inv_cs2;
        const double vtx = vx * c.tfac, vty = vy * c.tfac, vtz = vz * c.tfac;
        const double O0 = S[0] - wp * (S[0] - wcm[0] * omusq);
        lean_ptr(m.la.b[1], m.la.oxy[1][1])[0] = O0;
        MarchScalarPairs<0>::run(S, wcm, omusq, vtx, vty, vtz, wp, wmn, m);
        __syncwarp();
        if (m.T > 0) MarchDrain<1>::run(m, false);
    }
    m.T = m.NT - 1;
    MarchDrain<1>::run(m, true);
}

// the x-row (y, z) of the fluid set; the arithmetic is fluid_node<.., LEAN = true>
template <bool FULL, int NT_>
__device__ __forceinline__ void march_fluid_row(const StepArgs &a, Sh &sh, MarchSm *ms, const int lane, const int y,
                                                const int z)
{
    const EkConst &c = a.c;
    MarchCtx m;
    march_begin(m, c, a.in[0], ms, lane, y, z);
    for (m.T = 0; m.T < m.NT; ++m.T) {
        march_tile(m, c, y);
        double S[27];
        gather27_lean<EK_MODE_AA_ODD>(m.la, S);
        const double rho = sum27(S);
        double mo[3];
        momentum(S, mo);
        const double rhoinv = 1.0 / rho;
        bar_moments<NT_>();
        const double E[3] = {sh.E[0][lane], sh.E[1][lane], sh.E[2][lane]};
        const double dq = sh.cp[lane] - sh.cn[lane];
        double F[3], ex_[3], u[3];
        node_force_and_momentum(c, mo, dq, sh.T[lane], E, F, ex_);
        u[0] = rhoinv * ex_[0]; u[1] = rhoinv * ex_[1]; u[2] = rhoinv * ex_[2];
        sh.u[0][lane] = u[0]; sh.u[1][lane] = u[1]; sh.u[2][lane] = u[2];
        bar_velocity<NT_>();
        {
            const size_t i = (size_t)z * c.plane + m.la.fc;
            a.dq[(size_t)z * c.dq_sz + m.la.fdq] = dq;
            if (FULL) {  // LBM.cu:807-810
                a.fld[0][i] = rho; a.fld[1][i] = u[0]; a.fld[2][i] = u[1]; a.fld[3][i] = u[2];
            }
        }
        double wcr[4] = {c.w[0] * rho, c.w[1] * rho, c.w[2] * rho, c.w[3] * rho};
        const double omusq = 1.0 - 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) * c.inv_cs2;
        const double uF = u[0] * F[0] + u[1] * F[1] + u[2] * F[2];
        const double O0 = S[0] - c.wp[0] * (S[0] - wcr[0] * omusq) + c.dt * (c.sp * (-c.coe[0] * uF));
        lean_ptr(m.la.b[1], m.la.oxy[1][1])[0] = O0;
        MarchFluidPairs<0>::run(S, wcr, omusq, u, F, uF, c, m);
        __syncwarp();
        if (m.T > 0) MarchDrain<1>::run(m, false);
    }
    m.T = m.NT - 1;
    MarchDrain<1>::run(m, true);
}

template <bool FULL>
__global__ void __launch_bounds__(128, EK_MIN_CTAS) ek_march_kernel(const __grid_constant__ StepArgs a)
{
    __shared__ Sh sh;
    __shared__ MarchSm msm[4];
    const EkConst &c = a.c;
    const int lane = threadIdx.x & 31;
    const int role = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int y = blockIdx.x;
    if ((int)blockIdx.y < a.march_planes) {
        const int z = a.march_z0 + blockIdx.y;
        if (role == 0) march_fluid_row<FULL, 128>(a, sh, &msm[0], lane, y, z);
        else march_scalar_row<FULL, 128>(a, sh, &msm[role], role, lane, y, z);
    } else {
        // wall-adjacent planes (z = 0, 1 and NZ-2, NZ-1): the general node path, one x-tile per CTA
        const int w = blockIdx.y - a.march_planes, NT = c.NX >> 5;
        const int side = w / NT, tile = w - side * NT;
        const int x = tile * 32 + lane;
        const int z0 = a.wall_z0[side], z1 = a.wall_z1[side];
        const int pi = y * c.PX + x;
        if (role == 0) fluid_role<EK_MODE_AA_ODD, FULL, 128, false>(a, sh, lane, true, x, y, z0, z1);
        else scalar_role<EK_MODE_AA_ODD, FULL, false, 128, false>(a, sh, role, lane, true, x, y, pi, z0, z1);
    }
}

// ---------------------------------------------------------------------------
// Marching kernel, second form: the LOADS are aligned as well.  Every access of a tile is then
// "row pointer + immediate": nine row pointers (y-1..y+1 x z-1..z+1, tile T, column `lane`) advanced by
// one tile per iteration replace the 54 neighbour addresses that the z-walking kernel re-forms for every
// node (~135 integer instructions per thread and node, ek_lbm.cu SASS).  A population that comes from
// x-1 is loaded from column `lane` and shuffled up one lane, lane 0 taking what lane 31 loaded for the
// previous tile (carried in shared memory); one that comes from x+1 is shuffled down, lane 31 taking
// column 0 of the next tile (one extra sector per request: the only unaligned access left).
// ---------------------------------------------------------------------------
struct March2Sm {                // per warp
    double ring[9][64];          // c_x = -1 outputs of the tiles T-1 / T
    double carry_st[2][9];       // lane 31's c_x = +1 outputs of tile T
    double carry_ld[2][9];       // lane 31's loads of tile T for the directions that come from x-1
};

struct March2 {
    double *rp[3][3];            // [kz][ky]: rows z-1..z+1, y-1..y+1 of this set: tile T, column lane, slot 0
    LeanAddr la;                 // field offsets only (fc, fxm, fxp, fym, fyp, fdq)
    int lane, T, NT;
    int row0;                    // T*864 + lane: rp[..][..] - row0 is the start of the row
    unsigned col_lo, col_hi;     // lattice offsets of the columns xlo / xhi within a row
    March2Sm *ms;
};

__device__ __forceinline__ void march2_begin(March2 &m, const EkConst &c, double *lat, March2Sm *ms, int lane, int y, int z)
{
    m.lane = lane;
    m.NT = c.NX >> 5;
    m.T = 0;
    m.ms = ms;
    m.row0 = lane;
    m.col_lo = ek_lat_col(c.xlo);
    m.col_hi = ek_lat_col(c.xhi);
    const int yy[3] = {y == 0 ? c.NY - 1 : y - 1, y, y == c.NY - 1 ? 0 : y + 1};
#pragma unroll
    for (int kz = 0; kz < 3; ++kz)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
            m.rp[kz][ky] = lat + (size_t)(z - 1 + kz) * c.lplane + (size_t)yy[ky] * c.lrow + lane;
}

__device__ __forceinline__ void march2_fields(March2 &m, const EkConst &c, int y)
{
    const int x = m.T * 32 + m.lane;
    const int ym = y == 0 ? c.NY - 1 : y - 1, yp = y == c.NY - 1 ? 0 : y + 1;
    m.la.fc = y * c.PX + x;
    m.la.fxm = y * c.PX + (x == 0 ? c.xlo : x - 1);
    m.la.fxp = y * c.PX + (x == c.NX - 1 ? c.xhi : x + 1);
    m.la.fym = ym * c.PX + x;
    m.la.fyp = yp * c.PX + x;
    m.la.fdq = (long long)y * c.dq_sy + x;
}

__device__ __forceinline__ void march2_next(March2 &m)
{
#pragma unroll
    for (int kz = 0; kz < 3; ++kz)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) m.rp[kz][ky] += EK_TILE_ELEMS;
    m.row0 += EK_TILE_ELEMS;
}

// pre-collision populations of the node (32 T + lane, y, z): odd A-A step, S[d] = slot opp(d) at x - c_d.
// First every load is issued (27 aligned ones and lane 31's nine next-tile elements), then the shifts run.
template <int d>
struct March2Load {
    static __device__ __forceinline__ void run(const March2 &m, double S[27], double edge[9])
    {
        const double *q = m.rp[1 - ek_cz(d)][1 - ek_cy(d)] + ek_opp(d) * EK_TILE;
        S[d] = EK_LD(q);
        if (ek_cx(d) < 0) {
            // comes from x+1: lane 31 needs column 0 of the next tile (or the column xhi at the end of the row)
            constexpr int k = ek_xrank(d);
            edge[k] = 0.0;
            if (m.lane == 31) edge[k] = m.T < m.NT - 1 ? EK_LD(q + (EK_TILE_ELEMS - 31)) : EK_LD(q - m.row0 + m.col_hi);
        }
        March2Load<d + 1>::run(m, S, edge);
    }
};
template <>
struct March2Load<27> {
    static __device__ __forceinline__ void run(const March2 &, double *, double *) {}
};

template <int d>
struct March2Shift {
    static __device__ __forceinline__ void run(const March2 &m, double S[27], const double edge[9])
    {
        if (ek_cx(d) > 0) {
            constexpr int k = ek_xrank(d);
            const double a = S[d];
            double w = __shfl_up_sync(0xffffffffu, a, 1);
            if (m.lane == 31) m.ms->carry_ld[m.T & 1][k] = a;
            if (m.lane == 0) {
                if (m.T > 0) w = m.ms->carry_ld[(m.T & 1) ^ 1][k];
                else w = EK_LD(m.rp[1 - ek_cz(d)][1 - ek_cy(d)] + ek_opp(d) * EK_TILE - m.row0 + m.col_lo);
            }
            S[d] = w;
        } else if (ek_cx(d) < 0) {
            constexpr int k = ek_xrank(d);
            double w = __shfl_down_sync(0xffffffffu, S[d], 1);
            if (m.lane == 31) w = edge[k];
            S[d] = w;
        }
        March2Shift<d + 1>::run(m, S, edge);
    }
};
template <>
struct March2Shift<27> {
    static __device__ __forceinline__ void run(const March2 &, double *, const double *) {}
};

__device__ __forceinline__ void march2_gather(const March2 &m, double S[27])
{
    double edge[9];
    March2Load<0>::run(m, S, edge);
    March2Shift<1>::run(m, S, edge);
}

template <int d>
__device__ __forceinline__ void march2_put(const March2 &m, double v)
{
    double *q = m.rp[1 + ek_cz(d)][1 + ek_cy(d)] + d * EK_TILE;
    if (ek_cx(d) == 0) {
        EK_ST(q, v);
    } else if (ek_cx(d) > 0) {
        constexpr int k = ek_xrank(d);
        double w = __shfl_up_sync(0xffffffffu, v, 1);
        if (m.lane == 31) {
            m.ms->carry_st[m.T & 1][k] = v;
            if (m.T == m.NT - 1) EK_ST(q - m.row0 + m.col_hi, v);   // end of the row: column xhi
        }
        if (m.lane == 0 && m.T > 0) w = m.ms->carry_st[(m.T & 1) ^ 1][k];
        if (m.lane > 0 || m.T > 0) EK_ST(q, w);
    } else {
        constexpr int k = ek_xrank(d);
        m.ms->ring[k][(m.T & 1) * 32 + m.lane] = v;
        if (m.lane == 0 && m.T == 0) EK_ST(q - m.row0 + m.col_lo, v);   // start of the row: column xlo
    }
}

template <int d>
struct March2Drain {
    static __device__ __forceinline__ void run(const March2 &m, bool last)
    {
        if (ek_cx(d) < 0) {
            constexpr int k = ek_xrank(d);
            const int half = last ? (m.T & 1) : ((m.T & 1) ^ 1);
            const double w = m.ms->ring[k][(half * 32 + m.lane + 1) & 63];
            double *q = m.rp[1 + ek_cz(d)][1 + ek_cy(d)] + d * EK_TILE;
            if (!last) EK_ST(q - EK_TILE_ELEMS, w);
            else if (m.lane < 31) EK_ST(q, w);
        }
        March2Drain<d + 1>::run(m, last);
    }
};
template <>
struct March2Drain<27> {
    static __device__ __forceinline__ void run(const March2 &, bool) {}
};

template <int p>
struct March2ScalarPairs {
    static __device__ __forceinline__ void run(const double S[27], const double wcm[4], double omusq, double vtx, double vty,
                                               double vtz, double wp, double wmn, const March2 &m)
    {
        constexpr int d = 2 * p + 1, o = d + 1;
        double Oa, Ob;
        scalar_pair_trt<d>(S, wcm, omusq, vtx, vty, vtz, wp, wmn, Oa, Ob);
        march2_put<d>(m, Oa);
        march2_put<o>(m, Ob);
        March2ScalarPairs<p + 1>::run(S, wcm, omusq, vtx, vty, vtz, wp, wmn, m);
    }
};
template <>
struct March2ScalarPairs<13> {
    static __device__ __forceinline__ void run(const double *, const double *, double, double, double, double, double,
                                               double, const March2 &) {}
};

template <int p>
struct March2FluidPairs {
    static __device__ __forceinline__ void run(const double S[27], const double wcr[4], double omusq, const double u[3],
                                               const double F[3], double uF, const EkConst &c, const March2 &m)
    {
        constexpr int d = 2 * p + 1, o = d + 1;
        double Oa, Ob;
        fluid_pair_trt<d>(S, wcr, omusq, u, F, uF, c, Oa, Ob);
        march2_put<d>(m, Oa);
        march2_put<o>(m, Ob);
        March2FluidPairs<p + 1>::run(S, wcr, omusq, u, F, uF, c, m);
    }
};
template <>
struct March2FluidPairs<13> {
    static __device__ __forceinline__ void run(const double *, const double *, double, const double *, const double *,
                                               double, const EkConst &, const March2 &) {}
};

template <bool FULL, int NT_>
__device__ __forceinline__ void march2_scalar_row(const StepArgs &a, Sh &sh, March2Sm *ms, const int s, const int lane,
                                                  const int y, const int z)
{
    const EkConst &c = a.c;
    const double wp = c.wp[s], wmn = c.wm[s];
    const bool is_temp = (s == 3);
    const double Ks = s == 1 ? c.K : c.Kn;
    double *mom_sh = s == 1 ? sh.cp : (s == 2 ? sh.cn : sh.T);
    March2 m;
    march2_begin(m, c, a.in[s], ms, lane, y, z);
    for (; m.T < m.NT; ++m.T) {
        march2_fields(m, c, y);
        double S[27];
        march2_gather(m, S);
        const double mm = sum27(S);
        double E[3] = {0.0, 0.0, 0.0};
        mom_sh[lane] = mm;
        if (is_temp) {
            efield_lean(a, m.la, z, E);
            sh.E[0][lane] = E[0]; sh.E[1][lane] = E[1]; sh.E[2][lane] = E[2];
        }
        if (FULL) a.fld[3 + s][(size_t)z * c.plane + m.la.fc] = mm;   // charge, chargen, T (LBM.cu:811-813)
        bar_moments<NT_>();
        if (!is_temp) { E[0] = sh.E[0][lane]; E[1] = sh.E[1][lane]; E[2] = sh.E[2][lane]; }
        bar_velocity<NT_>();
        double vx = sh.u[0][lane], vy = sh.u[1][lane], vz = sh.u[2][lane];
        if (!is_temp) {
            vx = vx + Ks * E[0];
            vy = vy + Ks * E[1];
            vz = vz + Ks * E[2];
        }
        double wcm[4] = {c.w[0] * mm, c.w[1] * mm, c.w[2] * mm, c.w[3] * mm};
        const double omusq = 1.0 - 0.5 * (vx * vx + vy * vy + vz * vz) * c.inv_cs2;
        const double vtx = vx * c.tfac, vty = vy * c.tfac, vtz = vz * c.tfac;
        const double O0 = S[0] - wp * (S[0] - wcm[0] * omusq);
        m.rp[1][1][0] = O0;
        March2ScalarPairs<0>::run(S, wcm, omusq, vtx, vty, vtz, wp, wmn, m);
        __syncwarp();
        if (m.T > 0) March2Drain<1>::run(m, false);
        if (m.T < m.NT - 1) march2_next(m);
    }
    m.T = m.NT - 1;
    March2Drain<1>::run(m, true);
}

template <bool FULL, int NT_>
__device__ __forceinline__ void march2_fluid_row(const StepArgs &a, Sh &sh, March2Sm *ms, const int lane, const int y,
                                                 const int z)
{
    const EkConst &c = a.c;
    March2 m;
    march2_begin(m, c, a.in[0], ms, lane, y, z);
    for (; m.T < m.NT; ++m.T) {
        march2_fields(m, c, y);
        double S[27];
        march2_gather(m, S);
        const double rho = sum27(S);
        double mo[3];
        momentum(S, mo);
        const double rhoinv = 1.0 / rho;
        bar_moments<NT_>();
        const double E[3] = {sh.E[0][lane], sh.E[1][lane], sh.E[2][lane]};
        const double dq = sh.cp[lane] - sh.cn[lane];
        double F[3], ex_[3], u[3];
        node_force_and_momentum(c, mo, dq, sh.T[lane], E, F, ex_);
        u[0] = rhoinv * ex_[0]; u[1] = rhoinv * ex_[1]; u[2] = rhoinv * ex_[2];
        sh.u[0][lane] = u[0]; sh.u[1][lane] = u[1]; sh.u[2][lane] = u[2];
        bar_velocity<NT_>();
        {
            const size_t i = (size_t)z * c.plane + m.la.fc;
            a.dq[(size_t)z * c.dq_sz + m.la.fdq] = dq;
            if (FULL) {  // LBM.cu:807-810
                a.fld[0][i] = rho; a.fld[1][i] = u[0]; a.fld[2][i] = u[1]; a.fld[3][i] = u[2];
            }
        }
        double wcr[4] = {c.w[0] * rho, c.w[1] * rho, c.w[2] * rho, c.w[3] * rho};
        const double omusq = 1.0 - 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) * c.inv_cs2;
        const double uF = u[0] * F[0] + u[1] * F[1] + u[2] * F[2];
        const double O0 = S[0] - c.wp[0] * (S[0] - wcr[0] * omusq) + c.dt * (c.sp * (-c.coe[0] * uF));
        m.rp[1][1][0] = O0;
        March2FluidPairs<0>::run(S, wcr, omusq, u, F, uF, c, m);
        __syncwarp();
        if (m.T > 0) March2Drain<1>::run(m, false);
        if (m.T < m.NT - 1) march2_next(m);
    }
    m.T = m.NT - 1;
    March2Drain<1>::run(m, true);
}

template <bool FULL>
__global__ void __launch_bounds__(128, EK_MIN_CTAS) ek_march2_kernel(const __grid_constant__ StepArgs a)
{
    __shared__ Sh sh;
    __shared__ March2Sm msm[4];
    const EkConst &c = a.c;
    const int lane = threadIdx.x & 31;
    const int role = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int y = blockIdx.x;
    if ((int)blockIdx.y < a.march_planes) {
        const int z = a.march_z0 + blockIdx.y;
        if (role == 0) march2_fluid_row<FULL, 128>(a, sh, &msm[0], lane, y, z);
        else march2_scalar_row<FULL, 128>(a, sh, &msm[role], role, lane, y, z);
    } else {
        const int w = blockIdx.y - a.march_planes, NT = c.NX >> 5;
        const int side = w / NT, tile = w - side * NT;
        const int x = tile * 32 + lane;
        const int z0 = a.wall_z0[side], z1 = a.wall_z1[side];
        const int pi = y * c.PX + x;
        if (role == 0) fluid_role<EK_MODE_AA_ODD, FULL, 128, false>(a, sh, lane, true, x, y, z0, z1);
        else scalar_role<EK_MODE_AA_ODD, FULL, false, 128, false>(a, sh, role, lane, true, x, y, pi, z0, z1);
    }
}

#endif  // EK_XCHECK (marching kernels)

// odd A-A step, lean path, row stride as an immediate: instantiated for the x-tile counts of the named configs
// (C2 128 -> 4, C3 256 -> 8, 512 -> 16, C4 1024 -> 32; x-slabs carry a ghost tile: 130 -> 5, 258 -> 9, 514 -> 17,
// 1026 -> 33); any other row length takes the generic lean kernel
template <bool FULL, int NXT>
bool launch_odd_imm(const StepArgs &a, dim3 grid, cudaStream_t st)
{
    if (a.c.NXT != NXT) return false;
    ek_step_kernel<EK_MODE_AA_ODD, FULL, false, true, NXT * EK_TILE_ELEMS><<<grid, 128, 0, st>>>(a);
    return true;
}
template <bool FULL>
bool launch_odd_imm_any(const StepArgs &a, dim3 grid, cudaStream_t st)
{
#ifdef EK_NO_ROW_IMMEDIATE
    return false;
#else
    return launch_odd_imm<FULL, 4>(a, grid, st) || launch_odd_imm<FULL, 5>(a, grid, st) ||
           launch_odd_imm<FULL, 8>(a, grid, st) || launch_odd_imm<FULL, 9>(a, grid, st) ||
           launch_odd_imm<FULL, 16>(a, grid, st) || launch_odd_imm<FULL, 17>(a, grid, st) ||
           launch_odd_imm<FULL, 32>(a, grid, st) || launch_odd_imm<FULL, 33>(a, grid, st);
#endif
}

template <int MODE>
cudaError_t launch_mode(const StepArgs &a, bool full, bool earr, bool lean, dim3 grid, cudaStream_t st)
{
    constexpr bool AA = (MODE != EK_MODE_PUSH);
    if (MODE == EK_MODE_AA_ODD && lean && !earr && a.row_imm) {
        if (full ? launch_odd_imm_any<true>(a, grid, st) : launch_odd_imm_any<false>(a, grid, st)) return cudaGetLastError();
    }
    if (full) {
        if (earr) ek_step_kernel<MODE, true, true, false><<<grid, 128, 0, st>>>(a);
        else if (lean && AA) ek_step_kernel<MODE, true, false, AA><<<grid, 128, 0, st>>>(a);
        else ek_step_kernel<MODE, true, false, false><<<grid, 128, 0, st>>>(a);
    } else {
        if (earr) ek_step_kernel<MODE, false, true, false><<<grid, 128, 0, st>>>(a);
        else if (lean && AA) ek_step_kernel<MODE, false, false, AA><<<grid, 128, 0, st>>>(a);
        else ek_step_kernel<MODE, false, false, false><<<grid, 128, 0, st>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t ek_launch_step(const StepArgs &a, int mode, bool write_fields, bool e_from_arrays, bool lean,
                           cudaStream_t st)
{
    const EkConst &c = a.c;
    const int NT = (c.NX + 31) / 32;
    const int nxt = a.xt_mode == 0 ? NT : (a.xt_mode == 1 ? (NT < 2 ? NT : 2) : NT - 2);
    if (nxt <= 0) return cudaSuccess;
    dim3 grid(nxt, c.NY, a.nzblocks > 0 ? a.nzblocks : (c.NZ + a.zchunk - 1) / a.zchunk);
    switch (mode) {
    case EK_MODE_AA_EVEN: return launch_mode<EK_MODE_AA_EVEN>(a, write_fields, e_from_arrays, lean, grid, st);
    case EK_MODE_AA_ODD: return launch_mode<EK_MODE_AA_ODD>(a, write_fields, e_from_arrays, lean, grid, st);
    default: return launch_mode<EK_MODE_PUSH>(a, write_fields, e_from_arrays, false, grid, st);
    }
}

#ifdef EK_XCHECK
bool ek_march_applicable(const EkConst &c) { return (c.NX % 32) == 0 && c.NZ >= 6; }

// odd A-A step of the planes [z0, z1): x-marching rows for the deep interior (2 <= z <= NZ-3), the general
// node path for the wall-adjacent planes, in ONE launch
cudaError_t ek_launch_march(StepArgs a, bool write_fields, int z0, int z1, int variant, cudaStream_t st)
{
    const EkConst &c = a.c;
    const int m0 = z0 > 2 ? z0 : 2, m1 = z1 < c.NZ - 2 ? z1 : c.NZ - 2;
    a.march_z0 = m0;
    a.march_planes = m1 > m0 ? m1 - m0 : 0;
    a.wall_n = 0;
    if (z0 < 2) { a.wall_z0[a.wall_n] = z0; a.wall_z1[a.wall_n] = z1 < 2 ? z1 : 2; ++a.wall_n; }
    if (z1 > c.NZ - 2) { a.wall_z0[a.wall_n] = z0 > c.NZ - 2 ? z0 : c.NZ - 2; a.wall_z1[a.wall_n] = z1; ++a.wall_n; }
    dim3 grid(c.NY, a.march_planes + a.wall_n * (c.NX / 32));
    if (grid.y == 0) return cudaSuccess;
    // 21 KB of static shared memory per CTA: ask for a carve-out that keeps four CTAs per SM resident
    // (the default heuristic may pick a smaller one and the kernel loses a quarter of its warps)
    static bool configured = false;
    if (!configured) {
        const char *env = getenv("EK_MARCH_CARVEOUT");
        const int pct = env ? atoi(env) : 50;
        if (pct >= 0) {
            cudaFuncSetAttribute(ek_march_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            cudaFuncSetAttribute(ek_march_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            cudaFuncSetAttribute(ek_march2_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            cudaFuncSetAttribute(ek_march2_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
        if (getenv("EK_DEBUG")) {
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ek_march_kernel<false>, 128, 0);
            fprintf(stderr, "ek_march_kernel: %d CTAs per SM (carve-out %d %%)\n", nb, pct);
        }
        configured = true;
    }
    if (variant == 2) {
        if (write_fields) ek_march2_kernel<true><<<grid, 128, 0, st>>>(a);
        else ek_march2_kernel<false><<<grid, 128, 0, st>>>(a);
    } else {
        if (write_fields) ek_march_kernel<true><<<grid, 128, 0, st>>>(a);
        else ek_march_kernel<false><<<grid, 128, 0, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t ek_launch_step5(const StepArgs &a, int mode, bool write_fields, bool e_from_arrays, cudaStream_t st)
{
    const EkConst &c = a.c;
    dim3 grid((c.NX + 31) / 32, c.NY, a.nzblocks > 0 ? a.nzblocks : (c.NZ + a.zchunk - 1) / a.zchunk);
    switch (mode) {
    case EK_MODE_AA_EVEN: return launch_mode5<EK_MODE_AA_EVEN>(a, write_fields, e_from_arrays, grid, st);
    case EK_MODE_AA_ODD: return launch_mode5<EK_MODE_AA_ODD>(a, write_fields, e_from_arrays, grid, st);
    default: return launch_mode5<EK_MODE_PUSH>(a, write_fields, e_from_arrays, grid, st);
    }
}
#endif  // EK_XCHECK

cudaError_t ek_launch_export(const StepArgs &a, int mode, int set, double *dst, cudaStream_t st)
{
    const EkConst &c = a.c;
    dim3 block(64), grid((c.NX + 63) / 64, c.NY, c.NZ);
    if (mode == EK_MODE_AA_ODD) ek_export_kernel<EK_MODE_AA_ODD><<<grid, block, 0, st>>>(a, set, dst);
    else ek_export_kernel<EK_MODE_AA_EVEN><<<grid, block, 0, st>>>(a, set, dst);
    return cudaGetLastError();
}

cudaError_t ek_launch_import(const StepArgs &a, int set, const double *src, cudaStream_t st)
{
    const EkConst &c = a.c;
    dim3 block(64), grid((c.NX + 63) / 64, c.NY, c.NZ);
    ek_import_kernel<<<grid, block, 0, st>>>(a, set, src);
    return cudaGetLastError();
}
