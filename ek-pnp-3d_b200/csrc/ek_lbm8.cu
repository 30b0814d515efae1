// ek_lbm8.cu -- the fused stream-and-collide step with EIGHT warps per 32 cells.
//
// Same arithmetic, walls and streaming as ek_lbm.cu (see there for the citations
// of the reference), different work split.  ncu showed the four-warp kernel
// limited by the serial latency of its fluid warp (~1100 dependent-ish
// instructions per row at ~15 cycles each, 16 warps/SM), not by DRAM.  Here each
// population set is handled by TWO warps -- half A: rest + pairs 1..6 (slots
// 0..12), half B: pairs 7..13 (slots 13..26) -- so a thread holds 13/14
// populations (<= 80 registers, 3 CTAs = 24 warps per SM) and the per-row
// critical path shrinks ~3x:
//
//   1. every warp gathers its slots and publishes PARTIAL moments (zeroth moment
//      of its half; for the fluid also the six momentum brackets of
//      LBM.cu:639-644 restricted to its half; the temperature A-warp also E);
//   2. ONE barrier (shared memory is double-buffered by row parity);
//   3. every warp forms rho, c+, c-, T, F, u itself from the 23 partials
//      (~80 redundant flops per warp instead of a second barrier and a serial
//      fluid warp), then relaxes and scatters its own pairs.
//
// The zeroth moments are therefore (S0+..+S12) + (S13+..+S26) instead of the
// reference's single left-to-right chain: a 1-ulp-level difference, covered by
// the parity tolerances.
#include "ek_lbm_common.cuh"

namespace {

struct Sh8 {
    double rho[2][32];     // fluid: sum of slots 0..12 / 13..26
    double mp[2][3][32];   // fluid: positive momentum brackets per half
    double mn[2][3][32];   // fluid: negative momentum brackets per half
    double sc[3][2][32];   // cation, anion, temperature: zeroth moment per half
    double E[3][32];
};

__device__ __forceinline__ void bar_row() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// momentum brackets of LBM.cu:639-644 restricted to one half, same order
template <int HALF>
__device__ __forceinline__ void momentum_half(const double *S, double mp[3], double mn[3])
{
    if (HALF == 0) {
        // S[i] = slot i
        mp[0] = S[1] + S[7] + S[9];   mn[0] = S[2] + S[8] + S[10];
        mp[1] = S[3] + S[7] + S[11];  mn[1] = S[4] + S[8] + S[12];
        mp[2] = S[5] + S[9] + S[11];  mn[2] = S[6] + S[10] + S[12];
    } else {
        // S[i] = slot 13 + i
#define F_(d) S[(d) - 13]
        mp[0] = F_(13) + F_(15) + F_(19) + F_(21) + F_(23) + F_(26);
        mn[0] = F_(14) + F_(16) + F_(20) + F_(22) + F_(24) + F_(25);
        mp[1] = F_(14) + F_(17) + F_(19) + F_(21) + F_(24) + F_(25);
        mn[1] = F_(13) + F_(18) + F_(20) + F_(22) + F_(23) + F_(26);
        mp[2] = F_(16) + F_(18) + F_(19) + F_(22) + F_(23) + F_(25);
        mn[2] = F_(15) + F_(17) + F_(20) + F_(21) + F_(24) + F_(26);
#undef F_
    }
}

// what every warp derives from the published partials of one cell
struct Macro {
    double rho, cp, cn, T, dq;
    double E[3], F[3], ex[3];
};

__device__ __forceinline__ void read_macro(const EkConst &c, const Sh8 &sh, int lane, Macro &m)
{
    m.rho = sh.rho[0][lane] + sh.rho[1][lane];
    m.cp = sh.sc[0][0][lane] + sh.sc[0][1][lane];
    m.cn = sh.sc[1][0][lane] + sh.sc[1][1][lane];
    m.T = sh.sc[2][0][lane] + sh.sc[2][1][lane];
    m.dq = m.cp - m.cn;
#pragma unroll
    for (int k = 0; k < 3; ++k) m.E[k] = sh.E[k][lane];
    double mom[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        mom[k] = (sh.mp[0][k][lane] + sh.mp[1][k][lane]) - (sh.mn[0][k][lane] + sh.mn[1][k][lane]);
    // LBM.cu:635-644 (Ext enters only the force)
    m.F[0] = c.CtoC * m.dq * (m.E[0] + c.Ext) + c.exf;
    m.F[1] = c.CtoC * m.dq * m.E[1];
    m.F[2] = c.CtoC * m.dq * m.E[2] + c.rho0 * m.T * c.Ra * c.nu * c.D;
#pragma unroll
    for (int k = 0; k < 3; ++k) m.ex[k] = mom[k] * c.cflinv + m.F[k] * c.dt * 0.5;
}

// ------------------------------------------------------------------ scalar sets
template <int MODE, int HALF, int p>
struct ScalarPairs8 {
    static __device__ __forceinline__ void run(const double *S, const double wcm[4], double omusq, double vtx,
                                               double vty, double vtz, double wp, double wmn, bool wall, bool bottom,
                                               bool is_temp, const EkConst &c, double *lout, const Nbr &nb, int z,
                                               double *Wn, bool act)
    {
        constexpr int D0 = Half<HALF>::D0;
        constexpr int d = 2 * p + 1, o = d + 1, cls = ek_wclass(d);
        const double s_ = cdot<d>(vtx, vty, vtz);
        const double wm_ = wcm[cls];
        const double ep = wm_ * (omusq + 0.5 * s_ * s_);
        const double em = wm_ * s_;
        const double a = S[d - D0], b = S[o - D0];
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
                    // z is interior here: the target z +- 1 is a wall plane only next to a wall
                    const bool near_bottom = (z == 1), near_top = (z == c.NZ - 2);
                    const bool drop_a = ek_cz(d) > 0 ? near_top : (ek_cz(d) < 0 ? near_bottom : false);
                    const bool drop_b = ek_cz(d) > 0 ? near_bottom : (ek_cz(d) < 0 ? near_top : false);
                    if (!drop_a) put<MODE, d>(lout, nb, Oa);
                    if (!drop_b) put<MODE, o>(lout, nb, Ob);
                }
            } else {
                if (ek_cz(d) != 0) {
                    const bool a_inward = bottom ? (ek_cz(d) > 0) : (ek_cz(d) < 0);
                    if (a_inward) put<MODE, d>(lout, nb, Oa);
                    else put<MODE, o>(lout, nb, Ob);
                }
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
        ScalarPairs8<MODE, HALF, p + 1>::run(S, wcm, omusq, vtx, vty, vtz, wp, wmn, wall, bottom, is_temp, c, lout, nb,
                                             z, Wn, act);
    }
};
template <int MODE>
struct ScalarPairs8<MODE, 0, 6> {
    static __device__ __forceinline__ void run(const double *, const double *, double, double, double, double, double,
                                               double, bool, bool, bool, const EkConst &, double *, const Nbr &, int,
                                               double *, bool) {}
};
template <int MODE>
struct ScalarPairs8<MODE, 1, 13> {
    static __device__ __forceinline__ void run(const double *, const double *, double, double, double, double, double,
                                               double, bool, bool, bool, const EkConst &, double *, const Nbr &, int,
                                               double *, bool) {}
};

// One warp = (population set s, half HALF) of 32 cells, walking z0..z1-1.
template <int MODE, bool FULL, bool EARR, int HALF>
__device__ __forceinline__ void warp_role(const StepArgs &a, Sh8 *sh, const int s, const int lane, const bool act,
                                          Nbr &nb, const int pi, const int z0, const int z1)
{
    constexpr int D0 = Half<HALF>::D0, ND = Half<HALF>::ND, P0 = Half<HALF>::P0;
    const EkConst &c = a.c;
    double *lin = a.in[s];
    double *lout = a.out[s];
    const bool fluid = (s == 0), is_temp = (s == 3);
    double *W = fluid ? nullptr : a.wall + (size_t)(s - 1) * 2 * 27 * c.plane + pi;
    const double wp = c.wp[s], wmn = c.wm[s];
    const double Ks = s == 1 ? c.K : c.Kn;
    double S[ND];
    double expr1[3] = {0.0, 0.0, 0.0};
    int buf = 0;

    // z = -1 stands for the pre-pass on node z = 1 that the bottom wall needs (LBM.cu:663-801)
    for (int zz = (z0 == 0 ? -1 : z0); zz < z1; ++zz) {
        const bool pre = zz < 0;
        const int z = pre ? 1 : zz;
        set_z(nb, c, z);
        const bool bottom = (z == 0), top = (z == c.NZ - 1);
        const bool wall = bottom || top;
        double *Wn = fluid ? nullptr : W + (size_t)(bottom ? 0 : 27) * c.plane;
        if (!fluid && wall) {
#pragma unroll
            for (int i = 0; i < ND; ++i) S[i] = Wn[(size_t)(D0 + i) * c.plane];
        } else {
            gather_half<MODE, HALF>(lin, nb, S);
        }
        Sh8 &b = sh[buf];
        buf ^= 1;
        // ---- publish this half's partial moments
        const double part = sum_half<ND>(S);
        if (fluid) {
            double mp[3], mn[3];
            momentum_half<HALF>(S, mp, mn);
            b.rho[HALF][lane] = part;
#pragma unroll
            for (int k = 0; k < 3; ++k) { b.mp[HALF][k][lane] = mp[k]; b.mn[HALF][k][lane] = mn[k]; }
        } else {
            b.sc[s - 1][HALF][lane] = part;
            if (is_temp && HALF == 0) {
                double E[3];
                efield_at<EARR>(a, nb, z, E);
                b.E[0][lane] = E[0]; b.E[1][lane] = E[1]; b.E[2][lane] = E[2];
            }
        }
        bar_row();
        // ---- every warp forms the moments, the force and the velocity itself
        Macro m;
        read_macro(c, b, lane, m);
        if (pre) {
            expr1[0] = m.ex[0]; expr1[1] = m.ex[1]; expr1[2] = m.ex[2];
            continue;
        }
        const double rhoinv = 1.0 / m.rho;
        double u[3];
        if (bottom) {
            // u(z=0) = -(momentum expression of z=1) / rho(z=0)   (LBM.cu:778-800)
            u[0] = -rhoinv * expr1[0]; u[1] = -rhoinv * expr1[1]; u[2] = -rhoinv * expr1[2];
        } else {
            u[0] = rhoinv * m.ex[0]; u[1] = rhoinv * m.ex[1]; u[2] = rhoinv * m.ex[2];
        }
        if (fluid) {
            if (act) {
                const int i = nb.fc();
                if (HALF == 1) a.dq[dq_at(c, nb, z)] = m.dq;
                if (FULL && HALF == 0) {  // LBM.cu:807-810
                    a.fld[0][i] = m.rho; a.fld[1][i] = u[0]; a.fld[2][i] = u[1]; a.fld[3][i] = u[2];
                }
            }
            const double wcr[4] = {c.w[0] * m.rho, c.w[1] * m.rho, c.w[2] * m.rho, c.w[3] * m.rho};
            const double omusq = 1.0 - 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) * c.inv_cs2;
            const double uF = u[0] * m.F[0] + u[1] * m.F[1] + u[2] * m.F[2];
            if (HALF == 0) {
                // rest population: TRT + source in the interior, frozen on the walls
                double O0 = S[0];
                if (!wall) O0 = S[0] - c.wp[0] * (S[0] - wcr[0] * omusq) + c.dt * (c.sp * (-c.coe[0] * uF));
                if (act) {
                    if (MODE == EK_MODE_PUSH) lout[nb.lc()] = O0;
                    else if (!wall) lout[nb.lc()] = O0;
                }
            }
            FluidPairs8<MODE, HALF, P0>::run(S, wcr, omusq, u, m.F, uF, wall, top, c, lout, nb, act);
        } else {
            const double mom = s == 1 ? m.cp : (s == 2 ? m.cn : m.T);
            if (FULL && HALF == 0 && act) a.fld[3 + s][nb.fc()] = mom;  // charge, chargen, T (LBM.cu:811-813)
            double vx = u[0], vy = u[1], vz = u[2];
            if (!is_temp) {
                // ion drift u + K*E; Ext does not enter here (LBM.cu:851-862)
                vx = vx + Ks * m.E[0]; vy = vy + Ks * m.E[1]; vz = vz + Ks * m.E[2];
            }
            const double wcm[4] = {c.w[0] * mom, c.w[1] * mom, c.w[2] * mom, c.w[3] * mom};
            const double omusq = 1.0 - 0.5 * (vx * vx + vy * vy + vz * vz) * c.inv_cs2;
            const double vtx = vx * c.tfac, vty = vy * c.tfac, vtz = vz * c.tfac;
            if (HALF == 0) {
                const double O0 = S[0] - wp * (S[0] - wcm[0] * omusq);
                if (act) {
                    if (!wall) lout[nb.lc()] = O0;
                    else if (!is_temp) Wn[0] = O0;
                    else if (bottom) Wn[0] = -O0 + c.twoTw[0];
                    else Wn[0] = -O0;
                }
            }
            ScalarPairs8<MODE, HALF, P0>::run(S, wcm, omusq, vtx, vty, vtz, wp, wmn, wall, bottom, is_temp, c, lout,
                                              nb, z, Wn, act);
        }
    }
}

template <int MODE, bool FULL, bool EARR>
__global__ void __launch_bounds__(256, 3) ek_step8_kernel(const __grid_constant__ StepArgs a)
{
    __shared__ Sh8 sh[2];
    const EkConst &c = a.c;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int s = warp >> 1, half = warp & 1;
    int x = blockIdx.x * 32 + lane;
    const bool act = x < c.NX;
    if (!act) x = c.NX - 1;  // clamped duplicate: loads stay in bounds, stores are masked
    const int y = blockIdx.y;
    const int z0 = (blockIdx.z + a.zblock0) * a.zchunk;
    const int z1 = min(z0 + a.zchunk, c.NZ);
    Nbr nb;
    set_xy(nb, c, x, y);
    const int pi = y * c.PX + x;
    if (half == 0) warp_role<MODE, FULL, EARR, 0>(a, sh, s, lane, act, nb, pi, z0, z1);
    else warp_role<MODE, FULL, EARR, 1>(a, sh, s, lane, act, nb, pi, z0, z1);
}

template <int MODE>
cudaError_t launch_mode8(const StepArgs &a, bool full, bool earr, dim3 grid, cudaStream_t st)
{
    if (full) {
        if (earr) ek_step8_kernel<MODE, true, true><<<grid, 256, 0, st>>>(a);
        else ek_step8_kernel<MODE, true, false><<<grid, 256, 0, st>>>(a);
    } else {
        if (earr) ek_step8_kernel<MODE, false, true><<<grid, 256, 0, st>>>(a);
        else ek_step8_kernel<MODE, false, false><<<grid, 256, 0, st>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t ek_launch_step8(const StepArgs &a, int mode, bool write_fields, bool e_from_arrays, cudaStream_t st)
{
    const EkConst &c = a.c;
    dim3 grid((c.NX + 31) / 32, c.NY, a.nzblocks > 0 ? a.nzblocks : (c.NZ + a.zchunk - 1) / a.zchunk);
    switch (mode) {
    case EK_MODE_AA_EVEN: return launch_mode8<EK_MODE_AA_EVEN>(a, write_fields, e_from_arrays, grid, st);
    case EK_MODE_AA_ODD: return launch_mode8<EK_MODE_AA_ODD>(a, write_fields, e_from_arrays, grid, st);
    default: return launch_mode8<EK_MODE_PUSH>(a, write_fields, e_from_arrays, grid, st);
    }
}
