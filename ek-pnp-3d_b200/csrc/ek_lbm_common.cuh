// ek_lbm_common.cuh -- device helpers shared by the step kernels (ek_lbm.cu:
// four warps per 32 cells; ek_lbm8.cu: eight warps per 32 cells).
#pragma once

#include "ek_internal.cuh"

namespace {

// Offsets of the 3x3x3 neighbourhood of a node.  Populations live in a tiled
// layout per set, [z][y][x-tile][27 slots][32 lanes] (ek_internal.cuh): the slot
// stride is a compile-time 256 B, so the 27 accesses of a node differ only by an
// immediate and one 64-bit address per neighbour position is all the integer
// work a gather or scatter needs.  Macroscopic fields stay in the reference's
// [z][y][x] order (LBM.cu:22-25).
struct Nbr {
    unsigned lx[3], ly[3], lz[3];  // lattice element offsets of x-1,x,x+1 / y-1,y,y+1 / z-1,z,z+1 (z periodic)
    int fx[3], fy[3], fz[3];       // the same for the field arrays
    int yy;                        // row index (for the c+ - c- array, which has its own strides)
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
    nb.yy = y;
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

// index of a node in the c+ - c- array
__device__ __forceinline__ size_t dq_at(const EkConst &c, const Nbr &nb, int z)
{
    return (size_t)z * c.dq_sz + (size_t)nb.yy * c.dq_sy + nb.fx[1];
}

// population loads/stores.  The lattice is streamed once per step (no reuse inside a launch):
// EK_LD/EK_ST select the cache operator (default: plain; -DEK_CACHE_CS: evict-first streaming,
// -DEK_CACHE_CG: L2 only) -- measured variants are recorded in DESIGN.md 3.6.
#if defined(EK_CACHE_CS)
#define EK_LD(p) __ldcs(p)
#define EK_ST(p, v) __stcs((p), (v))
#elif defined(EK_CACHE_CG)
#define EK_LD(p) __ldcg(p)
#define EK_ST(p, v) __stcg((p), (v))
#elif defined(EK_CACHE_LDCG)
#define EK_LD(p) __ldcg(p)
#define EK_ST(p, v) (*(p) = (v))
#elif defined(EK_CACHE_STCG)
#define EK_LD(p) (*(p))
#define EK_ST(p, v) __stcg((p), (v))
#elif defined(EK_CACHE_LDLU)
#define EK_LD(p) __ldlu(p)
#define EK_ST(p, v) (*(p) = (v))
#else
#define EK_LD(p) (*(p))
#define EK_ST(p, v) (*(p) = (v))
#endif

// pre-collision populations of a node ("pull" restatement)
template <int MODE>
__device__ __forceinline__ void gather27(const double *lat, const Nbr &nb, double S[27])
{
    if (MODE == EK_MODE_AA_ODD) {
#pragma unroll
        for (int d = 0; d < 27; ++d) {
            const double *q = lat + nb.at(-ek_cx(d), -ek_cy(d), -ek_cz(d));
            S[d] = EK_LD(q + ek_opp(d) * EK_TILE);
        }
    } else {
        const double *q = lat + nb.lc();
#pragma unroll
        for (int d = 0; d < 27; ++d) S[d] = EK_LD(q + d * EK_TILE);
    }
}

// where the post-collision population of direction d goes
template <int MODE, int d>
__device__ __forceinline__ void put(double *lat, const Nbr &nb, double v)
{
    if (MODE == EK_MODE_AA_EVEN) {
        double *q = lat + nb.lc();
        EK_ST(q + ek_opp(d) * EK_TILE, v);
    } else {
        double *q = lat + nb.at(ek_cx(d), ek_cy(d), ek_cz(d));
        EK_ST(q + d * EK_TILE, v);
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
// Deep-interior nodes (2 <= z <= NZ-3): no wall rule applies to the node or to
// any of its 26 neighbours and z never wraps, so the per-access address is ONE
// IMAD.WIDE (plane base + 32-bit column offset) and the wall predicates vanish.
// 254 of 256 planes of the benchmark grids take this path; the arithmetic is
// the same as in the general path (results are bit-identical, tested).
// ---------------------------------------------------------------------------
struct LeanAddr {
    unsigned oxy[3][3];  // [cy+1][cx+1]: lattice element offset of column (x+cx, y+cy) within a plane
    double *b[3];        // lattice base of the planes z-1, z, z+1
    // LROW > 0 (row stride known at compile time, rows y-1..y+1 do not wrap): pim[k][i] = plane z-1+k, row y,
    // column x-1+i, slot 0 -- every access of the node is then "one of nine pointers + an immediate"
    // ((cy * LROW + slot * 32) * 8 bytes) instead of a 64-bit address formed per access
    double *pim[3][3];
    int zend;            // prefetch of the next z-iteration only inside this CTA's chunk: z + 1 < zend
    bool pf_m, pf_p;     // prefetch the x-1 / x+1 column too?  Not when it is a ghost column of a slab: nobody else
                         // needs the rest of a ghost tile's line
    int fc, fxm, fxp, fym, fyp;  // field offsets within a plane: centre, x-1, x+1, y-1, y+1
    long long fdq;               // y*dq_sy + x of the c+ - c- array
};

// (forcing the address to ONE mad.wide.u32 with the offsets pinned in registers
// was measured: 686 instead of 810 instructions per fluid node in the odd step,
// but 25 local-memory spill accesses, and 5.43 ms instead of 4.95 ms at 256^3)
__device__ __forceinline__ double *lean_ptr(double *base, unsigned off) { return base + off; }

__device__ __forceinline__ void lean_init(LeanAddr &la, const Nbr &nb)
{
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            la.oxy[j][i] = nb.ly[j] + nb.lx[i];
        }
    la.fc = nb.fy[1] + nb.fx[1];
    la.fxm = nb.fy[1] + nb.fx[0]; la.fxp = nb.fy[1] + nb.fx[2];
    la.fym = nb.fy[0] + nb.fx[1]; la.fyp = nb.fy[2] + nb.fx[1];
}

__device__ __forceinline__ void lean_set_z(LeanAddr &la, double *lat, const EkConst &c, int z)
{
    la.b[1] = lat + (size_t)z * c.lplane;
    la.b[0] = la.b[1] - c.lplane;
    la.b[2] = la.b[1] + c.lplane;
}

__device__ __forceinline__ void lean_set_pim(LeanAddr &la)
{
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int i = 0; i < 3; ++i) la.pim[k][i] = la.b[k] + la.oxy[1][i];
}

template <int MODE, int LROW = 0>
__device__ __forceinline__ void gather27_lean(const LeanAddr &la, double S[27])
{
    if (MODE == EK_MODE_AA_ODD && LROW > 0) {
#pragma unroll
        for (int d = 0; d < 27; ++d) {
            const double *q = la.pim[1 - ek_cz(d)][1 - ek_cx(d)] + (-ek_cy(d) * LROW + ek_opp(d) * EK_TILE);
            S[d] = EK_LD(q);
        }
    } else if (MODE == EK_MODE_AA_ODD) {
#pragma unroll
        for (int d = 0; d < 27; ++d) {
            const double *q = lean_ptr(la.b[1 - ek_cz(d)], la.oxy[1 - ek_cy(d)][1 - ek_cx(d)]);
            S[d] = EK_LD(q + ek_opp(d) * EK_TILE);
        }
    } else {
        const double *q = lean_ptr(la.b[1], la.oxy[1][1]);
#pragma unroll
        for (int d = 0; d < 27; ++d) S[d] = EK_LD(q + d * EK_TILE);
    }
}

// Prefetch of the 27 lines the NEXT z-iteration of this thread will gather (odd A-A step): the kernel is
// latency bound at 16 warps per SM, its loads in flight are limited by registers, a prefetch is not.
// Next iteration's planes z, z+1 are this iteration's b[1], b[2] (addresses already formed, other slots);
// only plane z+2 needs new addresses.
// (measured at 256^3 / 1024x512x64, odd launch: 5.214 -> 5.179 ms / 10.58 -> 10.39 ms with the L1 form on top of the
// row-stride immediates, the L2 form within 0.1 % of it; -DEK_NO_ODD_PREFETCH builds without)
#if !defined(EK_NO_ODD_PREFETCH) && !defined(EK_ODD_PREFETCH)
#define EK_ODD_PREFETCH 1
#endif
#ifndef EK_PF_INSTR
#define EK_PF_INSTR "prefetch.global.L1"
#endif
__device__ __forceinline__ void prefetch27_lean_odd(const LeanAddr &la, unsigned lplane)
{
#pragma unroll
    for (int d = 0; d < 27; ++d) {
        const int k = 1 - ek_cz(d);
        double *base = k == 2 ? la.b[2] + lplane : la.b[k + 1];
        const double *q = lean_ptr(base, la.oxy[1 - ek_cy(d)][1 - ek_cx(d)]) + ek_opp(d) * EK_TILE;
        if (ek_cx(d) > 0 && !la.pf_m) continue;
        if (ek_cx(d) < 0 && !la.pf_p) continue;
        asm volatile(EK_PF_INSTR " [%0];" ::"l"(q));
    }
}

// even A-A step: the node's own 27 slots of the next plane (one pointer + immediates)
__device__ __forceinline__ void prefetch27_lean_even(const LeanAddr &la, unsigned lplane)
{
    const double *q = lean_ptr(la.b[2], la.oxy[1][1]);
#pragma unroll
    for (int d = 0; d < 27; ++d) asm volatile(EK_PF_INSTR " [%0];" ::"l"(q + d * EK_TILE));
}

// the same with the row stride as an immediate: the next iteration's planes z, z+1 are pim[1], pim[2] (pointers
// already formed); plane z+2 costs three 64-bit additions
template <int LROW>
__device__ __forceinline__ void prefetch27_lean_odd_imm(const LeanAddr &la, unsigned lplane)
{
    const double *p3[3] = {la.pim[2][0] + lplane, la.pim[2][1] + lplane, la.pim[2][2] + lplane};
#pragma unroll
    for (int d = 0; d < 27; ++d) {
        const int k = 1 - ek_cz(d), i = 1 - ek_cx(d);
        const double *q = (k == 2 ? p3[i] : la.pim[k + 1][i]) + (-ek_cy(d) * LROW + ek_opp(d) * EK_TILE);
        if (i == 0 && !la.pf_m) continue;
        if (i == 2 && !la.pf_p) continue;
        asm volatile(EK_PF_INSTR " [%0];" ::"l"(q));
    }
}

template <int MODE, int d, bool LEAN, int LROW = 0>
__device__ __forceinline__ void putx(double *lat, const Nbr &nb, const LeanAddr &la, double v)
{
    if (LEAN && MODE == EK_MODE_AA_ODD && LROW > 0) {
        double *q = la.pim[1 + ek_cz(d)][1 + ek_cx(d)] + (ek_cy(d) * LROW + d * EK_TILE);
        EK_ST(q, v);
    } else if (LEAN) {
        if (MODE == EK_MODE_AA_EVEN) {
            double *q = lean_ptr(la.b[1], la.oxy[1][1]);
            EK_ST(q + ek_opp(d) * EK_TILE, v);
        } else {
            double *q = lean_ptr(la.b[1 + ek_cz(d)], la.oxy[1 + ek_cy(d)][1 + ek_cx(d)]);
            EK_ST(q + d * EK_TILE, v);
        }
    } else {
        put<MODE, d>(lat, nb, v);
    }
}

// efield_at<false>() for an interior plane without the z clamp
__device__ __forceinline__ void efield_lean(const StepArgs &a, const LeanAddr &la, int z, double E[3])
{
    const EkConst &c = a.c;
    const double *pc = a.phi + (size_t)z * c.plane;
    E[0] = 0.5 * (pc[la.fxm] - pc[la.fxp]) / c.dx;
    E[1] = 0.5 * (pc[la.fym] - pc[la.fyp]) / c.dy;
    const double *pcc = pc + la.fc;
    E[2] = 0.5 * (pcc[-c.plane] - pcc[c.plane]) / c.dz;
}


#ifdef EK_XCHECK
// ---- a population set split over two warps: half A = rest + pairs 1..6 (slots
// 0..12), half B = pairs 7..13 (slots 13..26)
template <int HALF> struct Half;
template <> struct Half<0> { static constexpr int D0 = 0, ND = 13, P0 = 0, P1 = 6; };
template <> struct Half<1> { static constexpr int D0 = 13, ND = 14, P0 = 6, P1 = 13; };

template <int MODE, int HALF>
__device__ __forceinline__ void gather_half(const double *lat, const Nbr &nb, double *S)
{
    constexpr int D0 = Half<HALF>::D0, ND = Half<HALF>::ND;
    if (MODE == EK_MODE_AA_ODD) {
#pragma unroll
        for (int i = 0; i < ND; ++i) {
            const int d = D0 + i;
            const double *q = lat + nb.at(-ek_cx(d), -ek_cy(d), -ek_cz(d));
            S[i] = q[ek_opp(d) * EK_TILE];
        }
    } else {
        const double *q = lat + nb.lc();
#pragma unroll
        for (int i = 0; i < ND; ++i) S[i] = q[(D0 + i) * EK_TILE];
    }
}

template <int ND>
__device__ __forceinline__ double sum_half(const double *S)
{
    double a = S[0];
#pragma unroll
    for (int i = 1; i < ND; ++i) a = a + S[i];
    return a;
}


// ------------------------------------------------------------------ fluid
template <int MODE, int HALF, int p>
struct FluidPairs8 {
    static __device__ __forceinline__ void run(const double *S, const double wcr[4], double omusq, const double u[3],
                                               const double F[3], double uF, bool wall, bool top, const EkConst &c,
                                               double *lout, const Nbr &nb, bool act)
    {
        constexpr int D0 = Half<HALF>::D0;
        constexpr int d = 2 * p + 1, o = d + 1, cls = ek_wclass(d);
        double Oa, Ob;
        if (!wall) {
            const double cu = cdot<d>(u[0], u[1], u[2]);
            const double cF = cdot<d>(F[0], F[1], F[2]);
            const double s_ = cu * c.tfac;
            const double wr = wcr[cls];
            const double ep = wr * (omusq + 0.5 * s_ * s_);
            const double em = wr * s_;
            const double a = S[d - D0], b = S[o - D0];
            const double np_ = c.wp[0] * (0.5 * (a + b) - ep);
            const double nm_ = c.wm[0] * (0.5 * (a - b) - em);
            const double Fp = c.sp * (c.coe[cls] * (cu * cF * c.cflinv2 - uF));
            const double Fm = c.sm * (c.coe[cls] * c.cflinv * cF);
            Oa = a - (np_ + nm_) + c.dt * (Fp + Fm);
            Ob = b - (np_ - nm_) + c.dt * (Fp - Fm);
        } else {
            Oa = S[o - D0];
            Ob = S[d - D0];
            if (top) {
                if (ek_uwsign(d) > 0) Oa = Oa + c.multi[cls]; else if (ek_uwsign(d) < 0) Oa = Oa - c.multi[cls];
                if (ek_uwsign(o) > 0) Ob = Ob + c.multi[cls]; else if (ek_uwsign(o) < 0) Ob = Ob - c.multi[cls];
            }
        }
        if (act) {
            put<MODE, d>(lout, nb, Oa);
            put<MODE, o>(lout, nb, Ob);
        }
        FluidPairs8<MODE, HALF, p + 1>::run(S, wcr, omusq, u, F, uF, wall, top, c, lout, nb, act);
    }
};
template <int MODE>
struct FluidPairs8<MODE, 0, 6> {
    static __device__ __forceinline__ void run(const double *, const double *, double, const double *, const double *,
                                               double, bool, bool, const EkConst &, double *, const Nbr &, bool) {}
};
template <int MODE>
struct FluidPairs8<MODE, 1, 13> {
    static __device__ __forceinline__ void run(const double *, const double *, double, const double *, const double *,
                                               double, bool, bool, const EkConst &, double *, const Nbr &, bool) {}
};

#endif  // EK_XCHECK

}  // namespace
