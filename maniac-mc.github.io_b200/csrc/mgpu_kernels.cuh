// mgpu_kernels.cuh -- sm_100a kernels of the MANIAC per-trial-move energy path.
//
// K1  pair_sums         fused real-space LJ + erfc-Coulomb of a trial molecule (old and
//                       new geometry in one pass over the targets) vs the host framework
//                       and the walker's guests            (pairwise_energy_utils.f90:21-179,
//                                                           geometry_utils.f90:210-284)
// K2  fill_trial_tables 1-D phase tables of the trial's charged atoms (charge folded in), dS(k),
//     kspace_entries    S_trial = S + dS, sum_k ffW |S_trial|^2     (ewald_phase.f90:205-312,
//                                                                   ewald_energy.f90:64-164)
// K3  k_widom_batch     one warp per test insertion, K1 + K2 read-only
// K4  k_total_energy    full recompute (energy_utils.f90:22-134, ewald_energy.f90:20-58)
//     k_sweep           device-resident move drivers + Metropolis (translation.f90,
//                       rotation.f90, creation.f90, deletion.f90, widom.f90,
//                       monte_carlo.f90:50-99), ONE WARP PER WALKER (or a team of two / four
//                       warps when a launch has few walkers), the walkers of a CTA in phase
//     k_trial           host-driven trials (compute_old/new_energy of monte_carlo_utils.f90),
//                       one warp per task for large batches, one CTA per task otherwise
//
// Every energy routine is written once for a cooperating "group" of NT threads
// (NT = 32: a warp, synchronised with __syncwarp; NT = 64 / 128: a team of warps with its own
// named barrier; NT = MGPU_BLOCK: a CTA, __syncthreads).
//
// All arithmetic is IEEE binary64.  No tensor cores: nothing here is a dense contraction.
// The per-pair erfc(alpha r)/r comes from a piecewise degree-6 polynomial in r^2 staged in
// shared memory, replicated per bank group so the gather is conflict free (error < 4e-15 of
// 1/r, see mgpu_internal.h); 1/r^2 for the LJ term comes from MUFU.RCP64H + two Newton steps.
// Out-of-range pairs use the exact forms.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "mgpu_internal.h"

__constant__ DevSys c_sys;

// software prefetch (register rotation) switches: framework pass with U > 1 targets per thread, k-space loop.
// Measured on B200 in round 1 (4736 walkers x 256 steps): none 18.22, k-space only 19.10, guest pass only 18.56,
// framework-U only 18.86, all three 18.36 M moves/s -- at 128 registers per thread the rotations compete for
// registers, so only k-space is on (the guest-pass rotation was removed).
// (round 2, on the build without the guest-pass rotation: framework-U prefetch on 24.69 -> 25.11 M moves/s, r02i)
#ifndef MGPU_PF_HOSTU
#define MGPU_PF_HOSTU 1
#endif
// (a cp.async staging of the framework atoms two iterations ahead was measured in round 1: 19.45 M moves/s against
// 20.5 M with the register rotation -- two more LDS.128 + two LDGSTS per iteration on an LSU that already serves the
// table gather; removed.)
// framework atoms per thread and iteration in the 3-probe-atom passes (1: three pair chains per thread; 2: six)
#ifndef MGPU_HOST_U3
#define MGPU_HOST_U3 1
#endif
#ifndef MGPU_ALIGN_GROUPS
#define MGPU_ALIGN_GROUPS 1               // phase-alignment groups per CTA (warp id mod this): 1 = every warp of the CTA meets at the barriers (measured best, r03c); 4 = the warps of one SM sub-partition (round 1); 8 = pairs
#endif
// (A hand-pipelined Coulomb-only framework pass -- geometry of the next atom and table part of the current one as two streams of
// one loop body -- was measured in round 2 (r03a): ptxas serialises the three table chains on eight reused registers, no gain;
// removed.)
#ifndef MGPU_SCREEN_NOTHING
#define MGPU_SCREEN_NOTHING 1
#endif
#ifndef MGPU_ACC_PER_ATOM
#define MGPU_ACC_PER_ATOM 0
#endif
// (Round 2 measured stored absolute atom positions for the guest passes -- three loads per target instead of six, with
// and without a prefetch of the next molecule: 22.9 / 23.0 M moves/s against 24.2 M for com + offset.  Half of the
// kernel's long-scoreboard stalls sit on those six loads, but the extra rows and registers cost more than they save.)
// Triclinic cells (configs[4]): every probe atom goes through the SAME pass instantiation (LJ + Coulomb with zero
// coefficients where a term is absent) instead of one pass per interaction class.  The r02k capture of k_sweep<true> put
// 29 % of the warp-stall samples on instruction fetch: with two guest species and four move kinds the four walkers of a
// quartet rarely run the same instantiation, and the classes multiply the hot code.  Costs the absent terms' arithmetic
// (N2's centre site: an LJ term with A = B = 0), removes the "nothing"-list screen (those pairs are simply evaluated).
#ifndef MGPU_TRI_MERGED
#define MGPU_TRI_MERGED 1
#endif
#ifndef MGPU_BLOCK
#define MGPU_BLOCK 256                    // CTA-per-task kernels (NT = MGPU_BLOCK)
#endif
#define MGPU_WARPS (MGPU_BLOCK / 32)
#ifndef MGPU_WBLOCK
#define MGPU_WBLOCK 512                   // warp-per-task kernels (NT = 32): one CTA per SM
#endif
#define MGPU_WGROUPS (MGPU_WBLOCK / 32)

// ------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------
struct cplx { double re, im; };
__device__ __forceinline__ cplx c_mul(cplx a, cplx b) { return { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Thread-group abstraction: NT = 32 (one warp), MGPU_TEAM / MGPU_TEAM2 (a team of four / two warps of one walker, on
// different SM sub-partitions, meeting at the team's own named barrier) or MGPU_BLOCK (the whole CTA).
#define MGPU_TEAM 128
#define MGPU_TEAM2 64
#define MGPU_BAR_PHASE 15                 // named barrier of the CTA-wide phase alignment of the team sweep
template <int NT> struct Grp;
template <> struct Grp<32> {
    static __device__ __forceinline__ int tid() { return threadIdx.x & 31; }
    static __device__ __forceinline__ int id() { return threadIdx.x >> 5; }          // group index inside the CTA
    static __device__ __forceinline__ void sync() { __syncwarp(); }
    // sum of NV values over the group; result valid in every thread
    template <int NV> static __device__ __forceinline__ void sum(double (&v)[NV], double *)
    {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    }
};
template <int NT> struct GrpTeam {        // NT = 64 or 128: at most 8 teams per CTA, named barriers 1 .. 8
    static __device__ __forceinline__ int tid() { return threadIdx.x & (NT - 1); }
    static __device__ __forceinline__ int id() { return threadIdx.x / NT; }
    static __device__ __forceinline__ void sync() { asm volatile("bar.sync %0, %1;" :: "r"(1 + (int)(threadIdx.x / NT)), "n"(NT) : "memory"); }
    template <int NV> static __device__ __forceinline__ void sum(double (&v)[NV], double *red)
    {
        const int lane = threadIdx.x & 31, wid = (threadIdx.x & (NT - 1)) >> 5;
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
        sync();                                // protect red from a previous use
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) red[i * MGPU_WARPS + wid] = v[i];
        }
        sync();
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NT / 32; ++w) s += red[i * MGPU_WARPS + w];   // fixed order
            v[i] = s;
        }
    }
};
template <> struct Grp<MGPU_TEAM> : GrpTeam<MGPU_TEAM> {};
template <> struct Grp<MGPU_TEAM2> : GrpTeam<MGPU_TEAM2> {};
template <> struct Grp<MGPU_BLOCK> {
    static __device__ __forceinline__ int tid() { return threadIdx.x; }
    static __device__ __forceinline__ int id() { return 0; }
    static __device__ __forceinline__ void sync() { __syncthreads(); }
    template <int NV> static __device__ __forceinline__ void sum(double (&v)[NV], double *red)
    {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
        __syncthreads();                       // protect red from a previous use
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) red[i * MGPU_WARPS + wid] = v[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < MGPU_WARPS; ++w) s += red[i * MGPU_WARPS + w];   // fixed order
            v[i] = s;
        }
    }
};

// gfortran MODULO(a,p) for reals (fmod + sign fix)
__device__ __forceinline__ double f_modulo(double a, double p)
{
    double r = fmod(a, p);
    if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p;
    return r;
}

// apply_PBC, geometry_utils.f90:45-97
__device__ __forceinline__ void apply_PBC(double pos[3])
{
    if (!c_sys.triclinic) {
#pragma unroll
        for (int d = 0; d < 3; ++d) pos[d] = c_sys.lo[d] + f_modulo(pos[d] - c_sys.lo[d], c_sys.H[d * 3 + d]);
    } else {
        double rel[3], f[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) rel[d] = pos[d] - c_sys.lo[d];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            f[i] = __dadd_rn(__dadd_rn(__dmul_rn(c_sys.Hinv[i * 3 + 0], rel[0]), __dmul_rn(c_sys.Hinv[i * 3 + 1], rel[1])),
                             __dmul_rn(c_sys.Hinv[i * 3 + 2], rel[2]));
            f[i] = f_modulo(f[i], 1.0);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
            pos[i] = c_sys.lo[i] + __dadd_rn(__dadd_rn(__dmul_rn(c_sys.H[i * 3 + 0], f[0]), __dmul_rn(c_sys.H[i * 3 + 1], f[1])),
                                             __dmul_rn(c_sys.H[i * 3 + 2], f[2]));
    }
}

// ------------------------------------------------------------------------------------
// minimum image (geometry_utils.f90:210-284) -> squared distance
// ------------------------------------------------------------------------------------
#define MGPU_RINT_MAGIC 6755399441055744.0     // 1.5 * 2^52: (x + M) - M = rint(x) for |x| < 2^51
// The reference's triclinic branch, literally: minimum over the 27 shifts i c1 + j c2 + k c3 of the RAW
// difference vector, c = COLUMNS of matrix (geometry_utils.f90:263-280).
__device__ __noinline__ double min_image_27(double dx, double dy, double dz)
{
    double best = 1.7976931348623157e308;
#pragma unroll 1
    for (int sx = -1; sx <= 1; ++sx)
#pragma unroll 1
        for (int sy = -1; sy <= 1; ++sy)
#pragma unroll
            for (int sz = -1; sz <= 1; ++sz) {
                double tx = dx + sx * c_sys.H[0] + sy * c_sys.H[1] + sz * c_sys.H[2];
                double ty = dy + sx * c_sys.H[3] + sy * c_sys.H[4] + sz * c_sys.H[5];
                double tz = dz + sx * c_sys.H[6] + sy * c_sys.H[7] + sz * c_sys.H[8];
                double d2 = tx * tx + ty * ty + tz * tz;
                best = fmin(best, d2);
            }
    return best;
}

// Triclinic minimum image of a Cartesian difference (guest passes, intramolecular pairs, setup kernels).  Deliberately NOT
// inlined: with this body inlined into every pair chain of every pass, k_sweep<true> is 34 k instructions; the framework
// passes, where the time goes, use min_image_frac below and the rest can afford a call.
__device__ __noinline__ double min_image_tri(double dx, double dy, double dz)
{
    // Same minimum as the 27-image search, found in 1 + tri_nrel candidates: round the fractional
    // coordinates (image n), then try the few lattice vectors +-C m that can shorten ANY vector with
    // fractional coordinates in [-1/2, 1/2]^3 (m G m < sum_d |(G m)_d|, G = C^T C; listed at init).
    // That is the global minimum over the lattice; it is the reference's answer whenever its shift
    // n - m lies in {-1,0,1}^3, which is checked -- otherwise (atoms far outside the cell, extreme
    // skew) the literal 27-image search decides.
    if (c_sys.tri_nrel < 0) return min_image_27(dx, dy, dz);
    // fractional coordinates: c_sys.Hinv holds the reference's "reciprocal" = TRANSPOSE of inverse(matrix)
    const double f0 = fma(c_sys.Hinv[0], dx, fma(c_sys.Hinv[3], dy, c_sys.Hinv[6] * dz));
    const double f1 = fma(c_sys.Hinv[1], dx, fma(c_sys.Hinv[4], dy, c_sys.Hinv[7] * dz));
    const double f2 = fma(c_sys.Hinv[2], dx, fma(c_sys.Hinv[5], dy, c_sys.Hinv[8] * dz));
    const double n0 = (f0 + MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC, n1 = (f1 + MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC,
                 n2 = (f2 + MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC;
    const double tx = dx - fma(c_sys.H[0], n0, fma(c_sys.H[1], n1, c_sys.H[2] * n2));
    const double ty = dy - fma(c_sys.H[3], n0, fma(c_sys.H[4], n1, c_sys.H[5] * n2));
    const double tz = dz - fma(c_sys.H[6], n0, fma(c_sys.H[7], n1, c_sys.H[8] * n2));
    // the list holds one vector of every +-m pair; of the two, only the one pointing against t can
    // shorten it: |t -+ C m|^2 - |t|^2 = |C m|^2 - 2 |t . C m|
    double gain = 0.0, bsign = 0.0;
    int bk = -1;
    // (framework atoms are Morton-sorted: the lanes of a warp look at one small region, so this is usually unanimous)
    const double t2 = fma(tx, tx, fma(ty, ty, tz * tz));
    const int nrel = __any_sync(__activemask(), t2 > c_sys.tri_safe2) ? c_sys.tri_nrel : 0;
    for (int k = 0; k < nrel; ++k) {
        const double dot = fma(tx, c_sys.tri_rel[k][0], fma(ty, c_sys.tri_rel[k][1], tz * c_sys.tri_rel[k][2]));
        const double gk = fma(-2.0, fabs(dot), c_sys.tri_len2[k]);
        if (gk < gain) { gain = gk; bk = k; bsign = dot > 0.0 ? -1.0 : 1.0; }
    }
    double cx = tx, cy = ty, cz = tz, o0 = n0, o1 = n1, o2 = n2;   // o = shift of the winner in the reference's convention, -(n - s m)
    if (bk >= 0) {
        cx = fma(bsign, c_sys.tri_rel[bk][0], tx); cy = fma(bsign, c_sys.tri_rel[bk][1], ty); cz = fma(bsign, c_sys.tri_rel[bk][2], tz);
        o0 = fma(-bsign, c_sys.tri_m[bk][0], n0); o1 = fma(-bsign, c_sys.tri_m[bk][1], n1); o2 = fma(-bsign, c_sys.tri_m[bk][2], n2);
    }
    if (fmax(fabs(o0), fmax(fabs(o1), fabs(o2))) > 1.0) return min_image_27(dx, dy, dz);
    return fma(cx, cx, fma(cy, cy, cz * cz));
}

template <bool TRI>
__device__ __forceinline__ double min_image_r2(double dx, double dy, double dz)
{
    if (!TRI) {
        // delta_d = modulo(delta_d + L/2, L) - L/2 restated as delta - L*rint(delta/L): the same
        // image except on the exact tie |delta| = L/2, where both images have the same length.
        const double nx = fma(dx, c_sys.invL[0], MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC;
        const double ny = fma(dy, c_sys.invL[1], MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC;
        const double nz = fma(dz, c_sys.invL[2], MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC;
        dx = fma(-c_sys.L[0], nx, dx);
        dy = fma(-c_sys.L[1], ny, dy);
        dz = fma(-c_sys.L[2], nz, dz);
        return fma(dx, dx, fma(dy, dy, dz * dz));
    } else {
        return min_image_tri(dx, dy, dz);
    }
}

// The same minimum image from the FRACTIONAL coordinate difference g = f(target) - f(probe) (framework passes of
// triclinic cells: the framework is stored in fractional coordinates, the probe atoms are converted once per pass).
// Rounding g itself wraps it (no projection of a Cartesian difference: 9 FMAs less per pair), and t = C (g - n) costs 6
// FMAs for a lower-triangular cell matrix.  The listed lattice vectors are only tried when the rounded vector sits
// within tri_eps of a face of the fractional cube (mgpu_init: elsewhere no listed vector can shorten it); the reference's
// 27-image search is the fallback whenever the winning shift leaves {-1,0,1}^3 (atoms far outside the cell), exactly as
// in min_image_tri.  The rare part is a separate function; the "rare" tests are integer compares on high words.
__device__ __noinline__ double min_image_frac_slow(double g0, double g1, double g2, unsigned cand)
{
    const double n0 = (g0 + MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC, n1 = (g1 + MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC,
                 n2 = (g2 + MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC;
    const double f0 = g0 - n0, f1 = g1 - n1, f2 = g2 - n2;
    const double tx = fma(c_sys.H[0], f0, fma(c_sys.H[1], f1, c_sys.H[2] * f2));
    const double ty = fma(c_sys.H[3], f0, fma(c_sys.H[4], f1, c_sys.H[5] * f2));
    const double tz = fma(c_sys.H[6], f0, fma(c_sys.H[7], f1, c_sys.H[8] * f2));
    // cand (warp-uniform, c_sys.tri_lut): the listed vectors that can matter near the faces this warp's lanes are near.
    // A vector that cannot shorten t has gk >= 0 and loses against gain = 0, so trying it is harmless.
    double cx = tx, cy = ty, cz = tz, o0 = n0, o1 = n1, o2 = n2;
    const unsigned cv = cand & 0x7fffffffu;
    if ((cv & (cv - 1u)) == 0u && cv != 0u) {
        // exactly one vector (the usual case in a mildly tilted cell): its data loaded once, applied by selects
        const int k = __ffs((int)cv) - 1;
        const double r0 = c_sys.tri_rel[k][0], r1 = c_sys.tri_rel[k][1], r2 = c_sys.tri_rel[k][2];
        const double m0 = c_sys.tri_m[k][0], m1 = c_sys.tri_m[k][1], m2 = c_sys.tri_m[k][2];
        const double dot = fma(tx, r0, fma(ty, r1, tz * r2));
        const double gk = fma(-2.0, fabs(dot), c_sys.tri_len2[k]);
        const double sg = (gk < 0.0) ? (dot > 0.0 ? -1.0 : 1.0) : 0.0;       // 0: the vector does not help this lane
        cx = fma(sg, r0, tx); cy = fma(sg, r1, ty); cz = fma(sg, r2, tz);
        o0 = fma(-sg, m0, n0); o1 = fma(-sg, m1, n1); o2 = fma(-sg, m2, n2);
    } else {
        double gain = 0.0, bsign = 0.0;
        int bk = -1;
        for (unsigned c = cv; c; c &= c - 1u) {
            const int k = __ffs((int)c) - 1;
            const double dot = fma(tx, c_sys.tri_rel[k][0], fma(ty, c_sys.tri_rel[k][1], tz * c_sys.tri_rel[k][2]));
            const double gk = fma(-2.0, fabs(dot), c_sys.tri_len2[k]);
            if (gk < gain) { gain = gk; bk = k; bsign = dot > 0.0 ? -1.0 : 1.0; }
        }
        if (bk >= 0) {
            cx = fma(bsign, c_sys.tri_rel[bk][0], tx); cy = fma(bsign, c_sys.tri_rel[bk][1], ty); cz = fma(bsign, c_sys.tri_rel[bk][2], tz);
            o0 = fma(-bsign, c_sys.tri_m[bk][0], n0); o1 = fma(-bsign, c_sys.tri_m[bk][1], n1); o2 = fma(-bsign, c_sys.tri_m[bk][2], n2);
        }
    }
    if (c_sys.tri_nrel < 0 || fmax(fabs(o0), fmax(fabs(o1), fabs(o2))) > 1.0) {      // the literal search, on the raw Cartesian difference
        const double dx = fma(c_sys.H[0], g0, fma(c_sys.H[1], g1, c_sys.H[2] * g2));
        const double dy = fma(c_sys.H[3], g0, fma(c_sys.H[4], g1, c_sys.H[5] * g2));
        const double dz = fma(c_sys.H[6], g0, fma(c_sys.H[7], g1, c_sys.H[8] * g2));
        return min_image_27(dx, dy, dz);
    }
    return fma(cx, cx, fma(cy, cy, cz * cz));
}
// Branch-free part: the rounded image's squared length, and in `near` whether the rounded vector is near ANY face of the
// fractional cube that matters to some listed vector (high word of max |f_d| against the smallest threshold: a superset, five
// integer instructions).  The caller votes ONCE for all the pair chains of an iteration (HostPass::block); only warps with
// a lane near a face work out which faces (frac_faces) and send the chains concerned through min_image_frac_slow with the
// listed vectors tri_lut selects: a vector only matters near EVERY face of its tri_req set, so in a mildly tilted cell a
// warp near one face tries one vector instead of the whole list (mgpu_init).  No branch in here -- not even on the cell's
// triangular shape (three FMAs with zero coefficients are cheaper than a basic-block boundary in every chain): with a vote
// and a call inside each chain the three chains of an iteration ran one after the other (r03b capture: 193 instructions
// per pair, a third of the samples fixed-latency waits).  Vectors beyond the reference's 27 images (|g_d| >= 1.5) cannot
// occur for probe atoms inside (-1/2, 3/2) of the cell, which HostPass::run checks once per pass.
__device__ __forceinline__ double min_image_frac_fast(double g0, double g1, double g2, bool &near)
{
    const double n0 = (g0 + MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC, n1 = (g1 + MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC,
                 n2 = (g2 + MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC;
    const double f0 = g0 - n0, f1 = g1 - n1, f2 = g2 - n2;                 // in [-1/2, 1/2]
    const double tx = fma(c_sys.H[0], f0, fma(c_sys.H[1], f1, c_sys.H[2] * f2));
    const double ty = fma(c_sys.H[3], f0, fma(c_sys.H[4], f1, c_sys.H[5] * f2));
    const double tz = fma(c_sys.H[6], f0, fma(c_sys.H[7], f1, c_sys.H[8] * f2));
    const int a0 = __double2hiint(f0) & 0x7fffffff, a1 = __double2hiint(f1) & 0x7fffffff, a2 = __double2hiint(f2) & 0x7fffffff;
    near = max(a0, max(a1, a2)) >= c_sys.tri_thr_min;
    return fma(tx, tx, fma(ty, ty, tz * tz));
}
// the faces (bits 0-2) the rounded vector is near, bit 3: |g_d| >= 1.5 for some d (index of tri_lut)
__device__ __forceinline__ unsigned frac_faces(double g0, double g1, double g2)
{
    const double n0 = (g0 + MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC, n1 = (g1 + MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC,
                 n2 = (g2 + MGPU_RINT_MAGIC) - MGPU_RINT_MAGIC;
    const double f0 = g0 - n0, f1 = g1 - n1, f2 = g2 - n2;
    const int a0 = __double2hiint(f0) & 0x7fffffff, a1 = __double2hiint(f1) & 0x7fffffff, a2 = __double2hiint(f2) & 0x7fffffff;
    const int b0 = __double2hiint(g0) & 0x7fffffff, b1 = __double2hiint(g1) & 0x7fffffff, b2 = __double2hiint(g2) & 0x7fffffff;
    return (a0 >= c_sys.tri_thr_hi[0] ? 1u : 0u) | (a1 >= c_sys.tri_thr_hi[1] ? 2u : 0u) | (a2 >= c_sys.tri_thr_hi[2] ? 4u : 0u) |
           (max(b0, max(b1, b2)) >= 0x3ff80000 ? 8u : 0u);
}

// Work counters for the roofline accounting: pairs evaluated, LJ terms inside the
// cutoff, erfc-Coulomb terms.  Integer adds on the otherwise idle ALU pipe.
struct PairCount { unsigned geom, lj, coul, scr; };      // scr: pairs of "nothing" lists settled by the per-molecule screen

// 1/s: MUFU.RCP64H seed + two Newton steps (s is a normal positive double here)
__device__ __forceinline__ double rcp_fast(double s)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    double e = fma(-s, y, 1.0);
    y = fma(y, e, y);
    e = fma(-s, y, 1.0);
    y = fma(y, fma(e, e, e), y);          // third-order step: error e^3
    return y;
}

// The reference's pair formulas evaluated literally (pairwise_lj_energy :95-138,
// pairwise_coulomb_energy :143-179); used off the hot path (host-host, intramolecular,
// pairs outside the table range).  A = 4 eps sigma^12, B = 4 eps sigma^6.
__device__ __noinline__ double2 pair_exact(double s, double A, double B, double qq, bool doC)
{
    double e_lj = 0.0, e_c = 0.0;
    const double r = sqrt(s);
    if (r < MGPU_ERR_TOL) {
        if (r < c_sys.rc) e_lj = c_sys.overlap;
        if (doC) e_c = c_sys.overlap;
        return make_double2(e_lj, e_c);
    }
    if (r < c_sys.rc) {
        const double y = 1.0 / s, y3 = y * y * y;
        e_lj = (A * y3 - B) * y3;
    }
    if (doC) e_c = qq * erfc(c_sys.alpha * r) / r;
    return make_double2(e_lj, e_c);
}

// One atom pair on the hot path.  tab = this lane's replica of the Coulomb table in shared
// memory (double2 units, see mgpu_internal.h).  qq = q_i q_j with either factor already zeroed
// when |q| < 1e-10 (:157).  AB = {4 eps sigma^12, 4 eps sigma^6}.
template <int REP>
__device__ __forceinline__ void pair_terms(double s, double2 AB, double qq, const double2 *__restrict__ tab,
                                           double &e_lj, double &e_c, PairCount &pc)
{
    pc.geom += 1u;
    const bool doL = (AB.x != 0.0) || (AB.y != 0.0), doC = (qq != 0.0);
    const int hi = __double2hiint(s);
    const int idx = (hi >> (20 - MGPU_TAB_K)) - c_sys.tab_ibase;
    if ((unsigned)idx >= (unsigned)c_sys.tab_nint) {           // r < table start (incl. overlap) or beyond its end: rare
        if (s >= c_sys.s_zero && s >= c_sys.rc2) { pc.coul += doC; return; }    // erfc(alpha r)/r < 1e-24: below the sum's rounding
        const double2 e = pair_exact(s, AB.x, AB.y, qq, doC);
        e_lj += e.x; e_c += e.y;
        pc.lj += (doL && s < c_sys.rc2); pc.coul += doC;
        return;
    }
    if (doL) {
        const double y = rcp_fast(s), y3 = y * y * y;
        const double e = (AB.x * y3 - AB.y) * y3;
        const bool in = s < c_sys.rc2;
        e_lj += in ? e : 0.0;
        pc.lj += in;
    }
    if (doC) {
        const int chi = (hi & ~((1 << (20 - MGPU_TAB_K)) - 1)) | (1 << (19 - MGPU_TAB_K));
        const double u = s - __hiloint2double(chi, 0);          // exact: same binade
        const double2 *t = tab + idx * (3 * REP);
        const double2 c01 = t[0], c23 = t[REP], c45 = t[2 * REP];
        const float uf = (float)u;
        const float pf = fmaf(__int_as_float(__double2hiint(c45.y)), uf, __int_as_float(__double2loint(c45.y)));
        double p = (double)pf;
        p = fma(p, u, c45.x);
        p = fma(p, u, c23.y);
        p = fma(p, u, c23.x);
        p = fma(p, u, c01.y);
        p = fma(p, u, c01.x);
        e_c = fma(qq, p, e_c);
        pc.coul += 1u;
    }
}

// ------------------------------------------------------------------------------------
// Probe: the trial molecule, staged in shared memory (one per group)
// ------------------------------------------------------------------------------------
struct Probe {
    int32_t kind, res, mol, na;
    int32_t has_old, has_new;
    int32_t excl_res, excl_mol;              // target slot to skip (the molecule itself)
    int32_t order_res, order_mol;            // >= 0: only targets with (res,mol) > this (ordering check)
    int32_t host_old, host_new;              // which geometries need a framework pass (old: 0 when the cache serves it)
    double q[MGPU_MAX_SITES];                // 0 when |q| < 1e-10
    int32_t type[MGPU_MAX_SITES];
    double (*po)[3];                         // old atom positions (com + offset), [natom_max][3] behind the workspace
    double (*pn)[3];                         // new atom positions
};

// Proposal / decision record of the device-resident drivers (one per group)
struct SweepShared {
    int32_t valid, move, kind, res, mol, accept, traced_accept, res2;   // res2: the new residue type of a swap
    double com[3];
    double off[MGPU_MAX_SITES][3];
    double prob;
    double e_old[6], e_new[6];
};

// Per-walker accumulators kept in shared memory during a sweep launch
struct WalkerLocal {
    unsigned long long rng[4];
    long long cnt[12];
    double avgN[MGPU_MAX_RES], avgN2[MGPU_MAX_RES];
    double avgE;
    long long n_samples;
    unsigned long long pc[3];
};

// Fixed part of a group's workspace; the two phase tables follow it.  `red` (cross-warp reduction
// scratch) is only needed when the group is a whole CTA: warp groups end their workspace before it, which
// keeps the 16-walker CTA under 196 KB of shared memory, i.e. one carve-out step lower and 32 KB more L1
// for the framework atoms every quartet streams.
struct GroupWS {
    Probe probe;
    int32_t count[MGPU_MAX_RES];
    int32_t sync_q, sync_k;                 // k_sweep: threads of this warp's alignment group if the group re-aligns before the guest pass / before k-space, else 0
    int32_t sync_top, sync_eval;            // ... at the top of the MC step / before the energy evaluation
    SweepShared sh;
    WalkerLocal loc;
    double red[8 * MGPU_WARPS];
};
__host__ __device__ inline size_t smem_ws_bytes(bool with_red)
{
    return ((with_red ? sizeof(GroupWS) : offsetof(GroupWS, red)) + 15) & ~size_t(15);
}

// The dynamic shared memory of every kernel here starts with the replicated Coulomb table,
// followed by the LJ {A,B} pairs.  Going through this accessor (not through a pointer stored
// in a struct) keeps the address space visible to the compiler: LDS.128, not generic LD.
extern __shared__ __align__(16) unsigned char mgpu_smem[];
// REP = number of replicas: MGPU_TAB_REP in the warp-per-task kernels (one CTA per SM, filled once
// per launch), 1 in the CTA-per-task kernels (latency path: a 15 KB fill per task, not 123 KB).
template <int NT> struct TabRep { static constexpr int v = (NT == 32 || NT == MGPU_TEAM || NT == MGPU_TEAM2) ? MGPU_TAB_REP : 1; };
template <int REP> __device__ __forceinline__ const double2 *smem_ctab() { return reinterpret_cast<const double2 *>(mgpu_smem) + (threadIdx.x & (REP - 1)); }
template <int REP> __device__ __forceinline__ const double2 *smem_ljAB() { return reinterpret_cast<const double2 *>(mgpu_smem) + (size_t)(c_sys.tab_nint + 1) * 3 * REP; }

// Shared-memory image of a CTA: replicated Coulomb table, LJ {A,B} pairs, then `groups` workspaces.
struct Smem {
    GroupWS *ws;
    double2 *tab_old, *tab_new;
};
// with_red: the group spans more than one warp and needs the cross-warp reduction scratch
__host__ __device__ inline size_t smem_group_bytes(int kmax_max, int natom_max, bool with_red)
{
    size_t b = smem_ws_bytes(with_red);
    b += (sizeof(double) * 2 * 3 * (size_t)natom_max + 15) & ~size_t(15);          // probe positions, old and new
    b += sizeof(double2) * 2 * (size_t)natom_max * 3 * (kmax_max + 1);
    return b;
}
__host__ __device__ inline size_t smem_common_bytes(int ntypes, int tab_nint, int rep)
{
    return sizeof(double2) * ((size_t)(tab_nint + 1) * 3 * rep + (size_t)ntypes * ntypes);      // + the all-zero row that closes the table
}
__host__ __device__ inline size_t smem_bytes(int ntypes, int tab_nint, int kmax_max, int natom_max, int groups, int rep, bool with_red)
{
    return smem_common_bytes(ntypes, tab_nint, rep) + groups * smem_group_bytes(kmax_max, natom_max, with_red) + 16;
}
// Carve the CTA's dynamic shared memory and (cooperatively, whole CTA) load the common part.
// Every thread of the CTA must call this; it ends with __syncthreads().
template <int REP, int NT>
__device__ __forceinline__ Smem smem_setup(unsigned char *base, int natom_max)
{
    Smem s;
    const int group = Grp<NT>::id();
    constexpr bool with_red = (NT != 32);
    double2 *d = reinterpret_cast<double2 *>(base);
    const int nt2 = c_sys.ntypes * c_sys.ntypes;
    const int nchunk = (c_sys.tab_nint + 1) * 3;              // ctab holds tab_nint rows + one all-zero row
    const double2 *src = reinterpret_cast<const double2 *>(c_sys.ctab);
    for (int i = threadIdx.x; i < nchunk * REP; i += blockDim.x) d[i] = src[i / REP];
    double2 *lj = d + (size_t)nchunk * REP;
    for (int i = threadIdx.x; i < nt2; i += blockDim.x) lj[i] = make_double2(c_sys.ljA[i], c_sys.ljB[i]);
    unsigned char *g = base + smem_common_bytes(c_sys.ntypes, c_sys.tab_nint, REP) + (size_t)group * smem_group_bytes(c_sys.kmax_max, natom_max, with_red);
    s.ws = reinterpret_cast<GroupWS *>(g);
    double (*ppos)[3] = reinterpret_cast<double (*)[3]>(g + smem_ws_bytes(with_red));
    if (Grp<NT>::tid() == 0) { s.ws->probe.po = ppos; s.ws->probe.pn = ppos + natom_max; s.ws->sync_q = 0; s.ws->sync_k = 0; }
    s.tab_old = reinterpret_cast<double2 *>(g + smem_ws_bytes(with_red) + ((sizeof(double) * 2 * 3 * (size_t)natom_max + 15) & ~size_t(15)));
    s.tab_new = s.tab_old + (size_t)natom_max * 3 * (c_sys.kmax_max + 1);
    __syncthreads();
    return s;
}

// ---- host-framework passes --------------------------------------------------------------
// The probe's atoms are sorted once per trial into lists by what they can feel from the
// framework: LJ only (MODE 1), Coulomb only (MODE 2), both (MODE 3), nothing (MODE 0, geometry
// only: the reference's overlap sentinel does not depend on eps or q).  Each list is swept in
// chunks of N <= 4 atoms held in registers against U framework atoms per thread and
// iteration; the body is branch free (N*U independent pair chains for the scheduler), pairs
// that fall outside the Coulomb table are flagged and redone exactly afterwards.
template <bool TRI, int MODE, int N, int U, int REP>
struct HostPass {
    double px[N], py[N], pz[N], q[N];
    int trow[N];

    __device__ __forceinline__ void load(const Probe &P, const double (*pos)[3], const int8_t *list)
    {
        const int nt = c_sys.ntypes;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int a = list[i];
            px[i] = pos[a][0]; py[i] = pos[a][1]; pz[i] = pos[a][2];
            q[i] = P.q[a]; trow[i] = P.type[a] * nt;
        }
    }

    // probe atoms of this pass that carry a charge (all of them in a Coulomb list; fewer in the merged passes of triclinic cells)
    __device__ __forceinline__ int n_charged() const
    {
        int c = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) c += (q[i] != 0.0) ? 1 : 0;
        return c;
    }

    template <int UU> struct Atoms { double2 xy[UU], zq[UU]; int tt[UU]; };

    template <int UU>
    __device__ __forceinline__ void fetch(Atoms<UU> &A, int j, int stride) const
    {
        // triclinic cells: the framework in fractional coordinates (min_image_frac)
        const double2 *__restrict__ hxy = TRI ? c_sys.host_fxy : c_sys.host_xy;
        const double2 *__restrict__ hzq = TRI ? c_sys.host_fzq : c_sys.host_zq;
        const int32_t *__restrict__ ht = c_sys.host_type;
#pragma unroll
        for (int u = 0; u < UU; ++u) {
            A.xy[u] = __ldg(hxy + j + u * stride); A.zq[u] = __ldg(hzq + j + u * stride);
            A.tt[u] = (MODE & 1) ? __ldg(ht + j + u * stride) : 0;
        }
    }

    // vmask: bit u set = target u is real (guest passes mask the tail / the excluded molecule;
    // framework passes hand in a constant all-ones mask and the tests fold away).
    // Coulomb sums are kept per probe atom WITHOUT its charge (acc[i] += q_j g(s)); the caller multiplies by
    // q_i once per pass.  Pairs below the table start (r < 1 A, incl. the overlap sentinel) read the all-zero
    // row that closes the table (the unsigned clamp sends both ends there), are left out of the LJ sum, and
    // are redone with the exact formulas when the block's smallest r^2 says there was one (e_x: their sum).
    // (bx, by, bz) = the probe atoms in the coordinates the targets are given in: Cartesian, or fractional (FRAC)
    template <int UU, bool FRAC>
    __device__ __forceinline__ void block(const Atoms<UU> &A, const unsigned vmask, const bool far_pass, const double (&bx)[N], const double (&by)[N], const double (&bz)[N],
                                          double &e_lj, double (&acc)[N], double2 &e_x, PairCount &pc) const
    {
        const double2 (&txy)[UU] = A.xy; const double2 (&tzq)[UU] = A.zq; const int (&tt)[UU] = A.tt;
        const double2 *ctab = smem_ctab<REP>(), *ljAB = smem_ljAB<REP>();
        const bool all = (vmask == (1u << UU) - 1u);
        int hmin = 0x7fffffff;
        double sv[UU][N];
        if (FRAC) {
            // fractional framework coordinates: every chain's rounded image first (one basic block, the chains interleave),
            // then ONE vote on the union of their faces; only warps near a relevant face look closer, chain by chain
            bool near_any = far_pass;
#pragma unroll
            for (int u = 0; u < UU; ++u)
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    bool nr;
                    sv[u][i] = min_image_frac_fast(txy[u].x - bx[i], txy[u].y - by[i], tzq[u].x - bz[i], nr);
                    near_any |= nr;
                }
            if (__any_sync(__activemask(), near_any)) {
#pragma unroll
                for (int u = 0; u < UU; ++u)
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        const double g0 = txy[u].x - bx[i], g1 = txy[u].y - by[i], g2 = tzq[u].x - bz[i];
                        const unsigned cand = c_sys.tri_lut[__reduce_or_sync(__activemask(), frac_faces(g0, g1, g2))];
                        if (cand) sv[u][i] = min_image_frac_slow(g0, g1, g2, cand);
                    }
            }
        } else if (TRI) {
#pragma unroll
            for (int u = 0; u < UU; ++u)
#pragma unroll
                for (int i = 0; i < N; ++i) sv[u][i] = min_image_r2<true>(txy[u].x - bx[i], txy[u].y - by[i], tzq[u].x - bz[i]);
        }
        // triclinic (large cells): whole warps are beyond the LJ cutoff -- one vote for all the chains of the iteration
        bool lj_pass = (MODE & 1) != 0;
        if ((MODE & 1) && TRI) {
            double smin = 1.0e300;
#pragma unroll
            for (int u = 0; u < UU; ++u)
#pragma unroll
                for (int i = 0; i < N; ++i) smin = fmin(smin, (all || ((vmask >> u) & 1u)) ? sv[u][i] : 1.0e300);
            lj_pass = __any_sync(__activemask(), smin < c_sys.rc2);
        }
#pragma unroll
        for (int u = 0; u < UU; ++u) {
            const bool val = (vmask >> u) & 1u;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                // (orthorhombic cells: the minimum image stays inside the chain, exactly as the loop was tuned -- r03)
                double s = TRI ? sv[u][i] : min_image_r2<false>(txy[u].x - bx[i], txy[u].y - by[i], tzq[u].x - bz[i]);
                if (!all) s = val ? s : 1.0e30;                                 // not a target: beyond every range (and (float)(s - centre) stays finite)
                sv[u][i] = s;
                const int hi = __double2hiint(s);
                hmin = min(hmin, hi);
                if ((MODE & 1) && lj_pass) {
                    const double2 AB = ljAB[trow[i] + tt[u]];
                    const double y = rcp_fast(s), y3 = y * y * y;
                    const double e = (AB.x * y3 - AB.y) * y3;
                    const bool in = (s < c_sys.rc2) && (hi >= c_sys.tab_hi_lo);
                    e_lj += in ? e : 0.0;
                    pc.lj += (in && (AB.x != 0.0 || AB.y != 0.0));
                }
                if (MODE & 2) {
                    const unsigned idx = min((unsigned)((hi >> (20 - MGPU_TAB_K)) - c_sys.tab_ibase), (unsigned)c_sys.tab_nint);
                    const int chi = (hi & ~((1 << (20 - MGPU_TAB_K)) - 1)) | (1 << (19 - MGPU_TAB_K));
                    const double uu = s - __hiloint2double(chi, 0);         // exact: same binade
                    const double2 *t = ctab + idx * (3 * REP);
                    const double2 c01 = t[0], c23 = t[REP], c45 = t[2 * REP];
                    const float uf = (float)uu;
                    const float pf = fmaf(__int_as_float(__double2hiint(c45.y)), uf, __int_as_float(__double2loint(c45.y)));
                    double p = (double)pf;
                    p = fma(p, uu, c45.x);
                    p = fma(p, uu, c23.y);
                    p = fma(p, uu, c23.x);
                    p = fma(p, uu, c01.y);
                    p = fma(p, uu, c01.x);
                    if (MGPU_ACC_PER_ATOM) acc[i] = fma(tzq[u].y, p, acc[i]);
                    else acc[0] = fma(q[i] * tzq[u].y, p, acc[0]);
                }
            }
        }
        if (hmin < c_sys.tab_hi_lo) {                               // rare: some r < 1 A (incl. overlap)
#pragma unroll
            for (int u = 0; u < UU; ++u)
#pragma unroll
                for (int i = 0; i < N; ++i)
                    if (__double2hiint(sv[u][i]) < c_sys.tab_hi_lo) {
                        const double s = sv[u][i];
                        double2 AB = make_double2(0.0, 0.0);
                        if (MODE & 1) AB = ljAB[trow[i] + tt[u]];
                        const double qq = (MODE & 2) ? q[i] * tzq[u].y : 0.0;
                        const double2 e = pair_exact(s, AB.x, AB.y, qq, qq != 0.0);
                        e_x.x += e.x; e_x.y += e.y;
                        if (MODE & 1) pc.lj += ((AB.x != 0.0 || AB.y != 0.0) && s < c_sys.rc2);
                    }
        }
    }

    // software-pipelined: the next block's atoms are in flight (L1 / L2 latency) while this one is evaluated
    // (plain register rotation; the last fetch reloads the current block).  Accumulators are taken and
    // returned by value so they stay in registers.
    __device__ __forceinline__ void run(int t0, int stride, double &e_lj_io, double &e_c_io, PairCount &pc_io) const
    {
        double e_lj = e_lj_io;
        double acc[N];
#pragma unroll
        for (int i = 0; i < N; ++i) acc[i] = 0.0;
        double2 e_x = make_double2(0.0, 0.0);
        PairCount pc = pc_io;
        const int n = c_sys.n_host;
        // the probe atoms in the coordinates the framework is stored in: fractional for triclinic cells
        // (c_sys.Hinv = transposed inverse of the cell matrix), Cartesian otherwise
        double bx[N], by[N], bz[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (TRI) {
                bx[i] = fma(c_sys.Hinv[0], px[i], fma(c_sys.Hinv[3], py[i], c_sys.Hinv[6] * pz[i]));
                by[i] = fma(c_sys.Hinv[1], px[i], fma(c_sys.Hinv[4], py[i], c_sys.Hinv[7] * pz[i]));
                bz[i] = fma(c_sys.Hinv[2], px[i], fma(c_sys.Hinv[5], py[i], c_sys.Hinv[8] * pz[i]));
            } else { bx[i] = px[i]; by[i] = py[i]; bz[i] = pz[i]; }
        }
        // triclinic: a probe atom farther than half a cell outside the cell (host-provided geometries only) can put the rounded
        // image beyond the reference's 27 -- then every pair of the pass takes the complete search
        bool far_pass = false;
        if (TRI) {
#pragma unroll
            for (int i = 0; i < N; ++i) far_pass |= !(fabs(bx[i] - 0.5) < 1.0 && fabs(by[i] - 0.5) < 1.0 && fabs(bz[i] - 0.5) < 1.0);
        }
        const int step = U * stride, reach = (U - 1) * stride;
        int j = t0;
        if (!MGPU_PF_HOSTU && U > 1) {
            for (; j + reach < n; j += step) { Atoms<U> a; fetch<U>(a, j, stride); block<U, TRI>(a, (1u << U) - 1u, far_pass, bx, by, bz, e_lj, acc, e_x, pc); }
        } else if (j + reach < n) {
            Atoms<U> cur;
            fetch<U>(cur, j, stride);
            for (;;) {
                const int jn = j + step;
                const bool more = jn + reach < n;
                Atoms<U> nxt;
                fetch<U>(nxt, more ? jn : j, stride);
                block<U, TRI>(cur, (1u << U) - 1u, far_pass, bx, by, bz, e_lj, acc, e_x, pc);
                j = jn;
                if (!more) break;
                cur = nxt;
            }
        }
        for (; j < n; j += stride) { Atoms<1> a1; fetch<1>(a1, j, stride); block<1, TRI>(a1, 1u, far_pass, bx, by, bz, e_lj, acc, e_x, pc); }
        double e_c = e_x.y;
        if (MGPU_ACC_PER_ATOM) {
#pragma unroll
            for (int i = 0; i < N; ++i) e_c = fma(q[i], acc[i], e_c);
        } else e_c += acc[0];
        if (t0 == 0) {                                           // work counters of the whole pass, once (SURVEY 8d accounting)
            pc.geom += (unsigned)(N * n);
            if (MODE & 2) pc.coul += (unsigned)(n_charged() * c_sys.n_host_charged);
        }
        e_lj_io = e_lj + e_x.x; e_c_io = e_c_io + e_c; pc_io = pc;
    }

    // The same body against ONE atom (index b) of every molecule of a guest residue type of the
    // walker: targets are com[m] + off_b[m], m = t0, t0 + stride, ... < n (molecule index fastest
    // in memory, so the lanes' loads coalesce); tq / ttype are the target atom's charge (0 if tiny)
    // and type.  Molecule m_skip is left out (the probe itself), and so is every m <= m_order
    // (ordering check of pairwise_energy_for_molecule, :60-62; -1 = none).
    __device__ __forceinline__ void run_guest(const double *__restrict__ com, const double *__restrict__ offb, int cap, int n,
                                              int t0, int stride, int m_skip, int m_order, double tq, int ttype,
                                              double &e_lj_io, double &e_c_io, PairCount &pc_io) const
    {
        double e_lj = e_lj_io;
        double acc[N];
#pragma unroll
        for (int i = 0; i < N; ++i) acc[i] = 0.0;
        double2 e_x = make_double2(0.0, 0.0);
        PairCount pc = pc_io;
        for (int m = t0; m < n; m += U * stride) {
            Atoms<U> A;
            unsigned vm = 0u;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int mm = m + u * stride;
                const bool ok = (mm < n) && (mm != m_skip) && (mm > m_order);
                const int mc = (mm < n) ? mm : m;
                A.xy[u] = make_double2(com[mc] + offb[mc], com[cap + mc] + offb[cap + mc]);
                A.zq[u] = make_double2(com[2 * cap + mc] + offb[2 * cap + mc], tq);
                A.tt[u] = ttype;
                vm |= ok ? (1u << u) : 0u;
            }
            block<U, false>(A, vm, false, px, py, pz, e_lj, acc, e_x, pc);
        }

        double e_c = e_x.y;
        if (MGPU_ACC_PER_ATOM) {
#pragma unroll
            for (int i = 0; i < N; ++i) e_c = fma(q[i], acc[i], e_c);
        } else e_c += acc[0];
        if (t0 == 0) {                                           // work counters of the whole list, once
            const int first = m_order + 1;                       // molecules first .. n-1 except m_skip
            const int nv = (n - first) - ((m_skip >= first && m_skip < n) ? 1 : 0);
            pc.geom += (unsigned)(N * nv);
            if ((MODE & 2) && tq != 0.0) pc.coul += (unsigned)(n_charged() * nv);
        }
        e_lj_io = e_lj + e_x.x; e_c_io = e_c_io + e_c; pc_io = pc;
    }
};

template <bool TRI, int MODE, int REP>
__device__ __forceinline__ void host_list(const Smem &S, const Probe &P, const double (*pos)[3], const int8_t *list, int n,
                                          int t0, int stride, double &e_lj, double &e_c, PairCount &pc)
{
    // chunks of at most 3 probe atoms: ~3 independent pair chains per thread fit the 128-register budget
    for (int base = 0; base < n; base += 3) {
        const int m = min(3, n - base);
        if (m == 3) { HostPass<TRI, MODE, 3, MGPU_HOST_U3, REP> hp; hp.load(P, pos, list + base); hp.run(t0, stride, e_lj, e_c, pc); }
        else if (m == 2) { HostPass<TRI, MODE, 2, 1, REP> hp; hp.load(P, pos, list + base); hp.run(t0, stride, e_lj, e_c, pc); }
        else { HostPass<TRI, MODE, 1, 3, REP> hp; hp.load(P, pos, list + base); hp.run(t0, stride, e_lj, e_c, pc); }
    }
}

// One target atom of a guest molecule against every probe atom (tq already zeroed if tiny).
template <bool TRI, int REP>
__device__ __forceinline__ void probe_vs_guest_atom(const Probe &P, const double (*pos)[3], const Smem &S,
                                                    double tx, double ty, double tz, double tq, int ttype,
                                                    double &e_lj, double &e_c, PairCount &pc)
{
    const int nt = c_sys.ntypes;
    const int na = P.na;
    for (int a = 0; a < na; ++a) {
        const double2 AB = smem_ljAB<REP>()[P.type[a] * nt + ttype];
        pair_terms<REP>(min_image_r2<TRI>(tx - pos[a][0], ty - pos[a][1], tz - pos[a][2]), AB, P.q[a] * tq, smem_ctab<REP>(), e_lj, e_c, pc);
    }
}

// The two target loops for ONE geometry of the probe: host framework (passes above) and the
// walker's guests (one thread per molecule).  Threads t0, t0 + stride, ... of the group take part.
template <bool TRI, int REP>
__device__ __forceinline__ void host_loops(const Probe &P, const double (*pos)[3], const Smem &S, int t0, int stride,
                                           double &e_lj, double &e_c, PairCount &pc)
{
    if (TRI && MGPU_TRI_MERGED) {
        // triclinic cells: ONE instantiation for every probe atom (LJ + Coulomb, zero coefficients where a term is absent)
        // instead of a pass per interaction class -- see MGPU_TRI_MERGED
        if (c_sys.n_host > 0) host_list<TRI, 3, REP>(S, P, pos, c_sys.iota, P.na, t0, stride, e_lj, e_c, pc);
    } else if (c_sys.n_host > 0) {
        const int r = P.res;
        host_list<TRI, 1, REP>(S, P, pos, c_sys.hl_list[r][1], c_sys.hl_n[r][1], t0, stride, e_lj, e_c, pc);
        host_list<TRI, 2, REP>(S, P, pos, c_sys.hl_list[r][2], c_sys.hl_n[r][2], t0, stride, e_lj, e_c, pc);
        host_list<TRI, 3, REP>(S, P, pos, c_sys.hl_list[r][3], c_sys.hl_n[r][3], t0, stride, e_lj, e_c, pc);
        host_list<TRI, 0, REP>(S, P, pos, c_sys.hl_list[r][0], c_sys.hl_n[r][0], t0, stride, e_lj, e_c, pc);
    }
}
// index of the probe-atom list of (probe residue ri) x (target residue g, target atom b) x mode
__device__ __forceinline__ int gl_index(int ri, int g, int b, int mode) { return ((ri * MGPU_MAX_RES + g) * MGPU_MAX_SITES + b) * 4 + mode; }

template <bool TRI, int MODE, int REP>
__device__ __forceinline__ void guest_list(const Probe &P, const double (*pos)[3], const int8_t *list, int nl,
                                           const double *com, const double *offb, int cap, int n, int t0, int stride,
                                           int m_skip, int m_order, double tq, int ttype, double &e_lj, double &e_c, PairCount &pc)
{
    for (int base = 0; base < nl; base += 3) {
        const int k = min(3, nl - base);
        if (k == 3) { HostPass<TRI, MODE, 3, 1, REP> hp; hp.load(P, pos, list + base); hp.run_guest(com, offb, cap, n, t0, stride, m_skip, m_order, tq, ttype, e_lj, e_c, pc); }
        else if (k == 2) { HostPass<TRI, MODE, 2, 1, REP> hp; hp.load(P, pos, list + base); hp.run_guest(com, offb, cap, n, t0, stride, m_skip, m_order, tq, ttype, e_lj, e_c, pc); }
        else { HostPass<TRI, MODE, 1, 3, REP> hp; hp.load(P, pos, list + base); hp.run_guest(com, offb, cap, n, t0, stride, m_skip, m_order, tq, ttype, e_lj, e_c, pc); }
    }
}

// The walker's guests: for every (target residue g, target atom b) the probe's atoms were sorted at
// init by what they exchange with that target atom (nothing / LJ / Coulomb / both: it only depends on
// the two atom types and charges), and each list is swept with the branch-free body of the
// framework passes, lanes over the molecules of g.
template <bool TRI, int REP>
__device__ __forceinline__ void guest_loops(const Probe &P, const double (*pos)[3], const Smem &S, int w, int t0, int stride,
                                            double &e_lj, double &e_c, PairCount &pc)
{
    const double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride;
    // Screen for the "nothing" lists (atom pairs with neither LJ nor Coulomb, e.g. TIP4P O x M/H: 6 of the 16 pairs of two
    // waters).  Their only possible contribution is the reference's overlap sentinel at r < 1e-10 (pairwise_lj_energy :116-119),
    // which needs two atoms on top of each other -- impossible for a molecule whose centre is farther from the probe's
    // centre than the two molecular radii.  Each thread checks that once per molecule it is responsible for (one
    // minimum image per molecule instead of one per atom pair) and only sweeps the "nothing" lists when it fails.
    double pcx = 0.0, pcy = 0.0, pcz = 0.0, prad2 = 0.0;
    if (MGPU_SCREEN_NOTHING) {
        const int na = P.na;
        for (int a = 0; a < na; ++a) { pcx += pos[a][0]; pcy += pos[a][1]; pcz += pos[a][2]; }
        const double inv = 1.0 / (double)na;
        pcx *= inv; pcy *= inv; pcz *= inv;
        for (int a = 0; a < na; ++a) {
            const double ax = pos[a][0] - pcx, ay = pos[a][1] - pcy, az = pos[a][2] - pcz;
            prad2 = fmax(prad2, ax * ax + ay * ay + az * az);
        }
    }
    for (int g = 0; g < c_sys.nres; ++g) {
        if (!c_sys.active[g]) continue;
        const int n = S.ws->count[g];
        if (n == 0) continue;
        const int cap = c_sys.cap[g], na_g = c_sys.natom[g];
        const double *com = wc + c_sys.goff[g];
        const double *off = com + 3 * (int64_t)cap;
        const int m_skip = (g == P.excl_res) ? P.excl_mol : -1;
        int m_order = -1;                                        // molecules m <= m_order are skipped
        if (P.order_res >= 0) m_order = (g < P.order_res) ? n : (g == P.order_res ? P.order_mol : -1);
        if (m_order >= n - 1 || (n == 1 && m_skip == 0)) continue;
        bool nothing_lists = true;
        if (MGPU_SCREEN_NOTHING && !(TRI && MGPU_TRI_MERGED)) {
            const double rr = sqrt(prad2) + sqrt(c_sys.rmax2[g] * (1.0 + 1.0e-6)) + 1.0e-9;
            const double thr2 = rr * rr;
            nothing_lists = false;
            for (int m = t0; m < n; m += stride) {
                const double s = min_image_r2<TRI>(com[m] - pcx, com[cap + m] - pcy, com[2 * cap + m] - pcz);
                nothing_lists |= (s < thr2) && (m != m_skip);
            }
        }
        for (int b = 0; b < na_g; ++b) {
            const double *offb = off + (int64_t)b * 3 * cap;
            double tq = c_sys.charge[g][b];
            if (fabs(tq) < MGPU_ERR_TOL) tq = 0.0;
            const int ttype = c_sys.type[g][b];
            if (TRI && MGPU_TRI_MERGED) {
                guest_list<TRI, 3, REP>(P, pos, c_sys.iota, P.na, com, offb, cap, n, t0, stride, m_skip, m_order, tq, ttype, e_lj, e_c, pc);
                continue;
            }
            const int gi = gl_index(P.res, g, b, 0);
            const int8_t *L = c_sys.gl_list + (int64_t)gi * MGPU_MAX_SITES;
            const int8_t *N = c_sys.gl_n + gi;
            guest_list<TRI, 1, REP>(P, pos, L + 1 * MGPU_MAX_SITES, N[1], com, offb, cap, n, t0, stride, m_skip, m_order, tq, ttype, e_lj, e_c, pc);
            guest_list<TRI, 2, REP>(P, pos, L + 2 * MGPU_MAX_SITES, N[2], com, offb, cap, n, t0, stride, m_skip, m_order, tq, ttype, e_lj, e_c, pc);
            guest_list<TRI, 3, REP>(P, pos, L + 3 * MGPU_MAX_SITES, N[3], com, offb, cap, n, t0, stride, m_skip, m_order, tq, ttype, e_lj, e_c, pc);
            if (nothing_lists)
                guest_list<TRI, 0, REP>(P, pos, L + 0 * MGPU_MAX_SITES, N[0], com, offb, cap, n, t0, stride, m_skip, m_order, tq, ttype, e_lj, e_c, pc);
            else if (t0 == 0) {                                  // this thread's share is not known per list; credit the whole list once
                const int first = m_order + 1;
                pc.scr += (unsigned)(N[0] * ((n - first) - ((m_skip >= first && m_skip < n) ? 1 : 0)));
            }
        }
    }
}

// named barrier of an alignment group of warps (see k_sweep): the warps with the same (warp id mod MGPU_ALIGN_GROUPS);
// MGPU_ALIGN_GROUPS = 1: every warp of the CTA.  The group is a compile-time constant on purpose: the same kernel with the
// mask in a register ran 6 % slower (r03g: 26.1 against 27.7 M moves/s).
__device__ __forceinline__ void quartet_sync(int nthreads_in_group)
{
    if (MGPU_ALIGN_GROUPS == 1) asm volatile("bar.sync 1, %0;" :: "r"(nthreads_in_group) : "memory");
    else asm volatile("bar.sync %0, %1;" :: "r"(1 + (int)((threadIdx.x >> 5) & (MGPU_ALIGN_GROUPS - 1))), "r"(nthreads_in_group) : "memory");
}
// Phase alignment of the sweep kernels.  One warp per walker (NT = 32): the group barrier above, n = threads of the
// group.  One TEAM per walker (NT = MGPU_TEAM / MGPU_TEAM2; every team has a warp on every / every other sub-partition):
// all teams of the CTA meet, n = threads of the CTA.
template <int NT> __device__ __forceinline__ void phase_barrier(int n)
{
    if (NT == 32) quartet_sync(n);
    else asm volatile("bar.sync %0, %1;" :: "n"(MGPU_BAR_PHASE), "r"(n) : "memory");
}

// K1: group-cooperative pair sums of the probe against host atoms + the walker's guests.
// When two geometries need the same kind of pass (old AND new of a move) the group splits in
// two halves, one per geometry, so every thread works on a single geometry; otherwise the
// whole group works on the one geometry there is (creation / deletion / Widom; and the
// framework pass of a move whose old-geometry framework sum comes from the per-molecule cache).
// out[0..3] = framework part {lj_old, coul_old(e^2/A), lj_new, coul_new}; out[4..7] = guest part.
// Valid in every thread of the group.
template <bool TRI, int NT>
__device__ __noinline__ void pair_sums(const Smem &S, int w, double (&out)[8], PairCount &pc)
{
    const Probe &P = S.ws->probe;
    const int gt = Grp<NT>::tid();
    PairCount pcl = pc;                      // by value: keeps the counters in registers inside the loops
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0;
    if (P.host_old || P.host_new) {
        const bool both = P.host_old && P.host_new;
        const bool new_set = both ? (gt >= NT / 2) : (P.host_new != 0);
        const int stride = both ? NT / 2 : NT;
        const int t0 = both ? (gt & (NT / 2 - 1)) : gt;
        double e_lj = 0.0, e_c = 0.0;
        host_loops<TRI, TabRep<NT>::v>(P, new_set ? P.pn : P.po, S, t0, stride, e_lj, e_c, pcl);
        acc[0] = new_set ? 0.0 : e_lj; acc[1] = new_set ? 0.0 : e_c;
        acc[2] = new_set ? e_lj : 0.0; acc[3] = new_set ? e_c : 0.0;
    }
    if (NT != MGPU_BLOCK && S.ws->sync_q) phase_barrier<NT>(S.ws->sync_q);     // k_sweep, MGPU_OPT_PHASE_SYNC bit 2: re-align before the guest pass
    {
        const bool both = P.has_old && P.has_new;
        const bool new_set = both ? (gt >= NT / 2) : (P.has_new != 0);
        const int stride = both ? NT / 2 : NT;
        const int t0 = both ? (gt & (NT / 2 - 1)) : gt;
        double e_lj = 0.0, e_c = 0.0;
        guest_loops<TRI, TabRep<NT>::v>(P, new_set ? P.pn : P.po, S, w, t0, stride, e_lj, e_c, pcl);
        acc[4] = new_set ? 0.0 : e_lj; acc[5] = new_set ? 0.0 : e_c;
        acc[6] = new_set ? e_lj : 0.0; acc[7] = new_set ? e_c : 0.0;
    }
    pc = pcl;
    Grp<NT>::template sum<8>(acc, S.ws->red);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = acc[i];
}

// per-molecule framework-energy cache of walker w, residue res: rows {lj, coulomb (e^2/A)} x cap,
// stored right behind the offset rows of the residue block
__device__ __forceinline__ double *hcache_rows(int w, int res)
{
    return c_sys.coords + (int64_t)w * c_sys.coord_stride + c_sys.goff[res] + (int64_t)(3 + 3 * c_sys.natom[res]) * c_sys.cap[res];
}

// intra_res_real_coulomb_energy, ewald_energy.f90:212-252 (one thread, tiny)
template <bool TRI>
__device__ double intra_energy(int res, const double (*pos)[3])
{
    const int na = c_sys.natom[res];
    double u = 0.0;
    for (int a1 = 0; a1 < na - 1; ++a1) {
        const double c1 = c_sys.charge[res][a1];
        for (int a2 = a1 + 1; a2 < na; ++a2) {
            const double c2 = c_sys.charge[res][a2];
            const double r2 = min_image_r2<TRI>(pos[a2][0] - pos[a1][0], pos[a2][1] - pos[a1][1], pos[a2][2] - pos[a1][2]);
            const double d = sqrt(r2);
            if (d < MGPU_ERR_TOL) continue;
            u = u + c1 * c2 * (erfc(c_sys.alpha * d) - 1.0) / d;
        }
    }
    return u * c_sys.eps0_inv_real;
}

// ------------------------------------------------------------------------------------
// K2: reciprocal space
// ------------------------------------------------------------------------------------
// Fill the 1-D phase tables e^{i k theta_d}, k = 0..kmax_d, for n atoms at positions pos:
// compute_atom_phase (ewald_phase.f90:255-280) + compute_phase_factor (:286-312).  One
// sincos per (atom, dim); higher k by repeated complex multiplication (the reference calls
// cos/sin per k; the two agree to ~k ulp).  tab layout: [atom][dim][k], stride KW = kmax_max+1.
// qskip (optional): per-atom charges; atoms whose charge is exactly 0 add exactly nothing to S(k) in the reference
// (q * phase), so their tables are neither filled nor read (TIP4P oxygen: a quarter of the molecule's k-space work).
__device__ __forceinline__ void fill_phase_tables(double2 *tab, const double (*pos)[3], int n, int tid0, int nthreads, const double *qskip = nullptr)
{
    const int KW = c_sys.kmax_max + 1;
    for (int e = tid0; e < n * 3; e += nthreads) {
        const int d = e % 3, a = e / 3;
        if (qskip && qskip[a] == 0.0) continue;
        double ph = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) ph = __dadd_rn(ph, __dmul_rn(c_sys.Hinv[j * 3 + d], pos[a][j]));
        ph = c_sys.twopi * ph;
        double sn, cs;
        sincos(ph, &sn, &cs);
        double2 *t = tab + (int64_t)e * KW;
        double cr = 1.0, ci = 0.0;
        t[0] = make_double2(1.0, 0.0);
        const int km = c_sys.kmax[d];
        for (int k = 1; k <= km; ++k) {
            const double nr = cr * cs - ci * sn, ni = cr * sn + ci * cs;
            cr = nr; ci = ni;
            // re-anchor every 8 steps so the recurrence error stays ~1e-16 for large kmax
            if ((k & 7) == 0) sincos((double)k * ph, &ci, &cr);
            t[k] = make_double2(cr, ci);
        }
    }
}

__device__ __forceinline__ cplx phase_product(const double2 *tab, int a, int kx, int ky, int kz)
{
    const int KW = c_sys.kmax_max + 1;
    const double2 *t = tab + (int64_t)a * 3 * KW;
    const double2 f1 = t[kx];                                   // kx >= 0 always
    const double2 f2 = t[KW + (ky < 0 ? -ky : ky)];
    const double2 f3 = t[2 * KW + (kz < 0 ? -kz : kz)];
    cplx a1 = { f1.x, f1.y };
    cplx a2 = { f2.x, ky < 0 ? -f2.y : f2.y };
    cplx a3 = { f3.x, kz < 0 ? -f3.y : f3.y };
    return c_mul(c_mul(a1, a2), a3);
}

// The phase tables of ONE trial, as a list of ENTRIES: every atom of the probe whose charge is not exactly 0 (c_sys.qlist; the
// others add exactly nothing to S(k) in the reference, q * phase), first in the new geometry, then in the old one.  The charge
// is folded into the entry's x-table, with a minus sign for the old geometry, so that dS(k) = sum over entries of
// tx[kx] ty[ky] tz[kz] with no per-entry branch, charge load or sign.  Old and new geometry are filled in one round (twice
// the threads busy during the sincos latency).  Layout [entry][dim][k], stride KW = kmax_max + 1, in the space of the
// group's two tables (S.tab_old onwards).  Returns the number of entries.
template <int NT>
__device__ __forceinline__ int fill_trial_tables(const Smem &S, const Probe &P)
{
    const int KW = c_sys.kmax_max + 1;
    const int nq = c_sys.nq[P.res];
    const int n_new = P.has_new ? nq : 0, ne = n_new + (P.has_old ? nq : 0);
    for (int idx = Grp<NT>::tid(); idx < ne * 3; idx += NT) {
        const int e = idx / 3, d = idx - 3 * e;
        const bool is_new = e < n_new;
        const int a = c_sys.qlist[P.res][is_new ? e : e - n_new];
        const double (*pos)[3] = is_new ? P.pn : P.po;
        double ph = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) ph = __dadd_rn(ph, __dmul_rn(c_sys.Hinv[j * 3 + d], pos[a][j]));
        ph = c_sys.twopi * ph;
        double sn, cs;
        sincos(ph, &sn, &cs);
        double sc = 1.0;
        if (d == 0) { const double q = c_sys.charge[P.res][a]; sc = is_new ? q : -q; }
        double2 *t = S.tab_old + (int64_t)idx * KW;
        double cr = 1.0, ci = 0.0;
        t[0] = make_double2(sc, 0.0);
        const int km = c_sys.kmax[d];
        for (int k = 1; k <= km; ++k) {
            const double nr = cr * cs - ci * sn, ni = cr * sn + ci * cs;
            cr = nr; ci = ni;
            if ((k & 7) == 0) sincos((double)k * ph, &ci, &cr);   // re-anchor, as in fill_phase_tables
            t[k] = make_double2(sc * cr, sc * ci);
        }
    }
    return ne;
}

// One entry's contribution to dS(k): tx[kx] * ty[|ky|]^(*) * tz[|kz|]^(*) added to (sr, si); 8 FP64 instructions.
__device__ __forceinline__ void entry_term(const double2 *__restrict__ t, int o1, int o2, int o3, int sy, int sz, double &sr, double &si)
{
    const double2 f1 = t[o1], f2 = t[o2], f3 = t[o3];
    // conjugate for negative ky / kz: flip the sign bit of the imaginary part (sy, sz = 0 or 0x80000000)
    const double f2i = __hiloint2double(__double2hiint(f2.y) ^ sy, __double2loint(f2.y));
    const double f3i = __hiloint2double(__double2hiint(f3.y) ^ sz, __double2loint(f3.y));
    const double pr = fma(-f1.y, f2i, f1.x * f2.x), pi = fma(f1.y, f2.x, f1.x * f2i);
    sr = fma(-pi, f3i, fma(pr, f3.x, sr));
    si = fma(pi, f3.x, fma(pr, f3i, si));
}

// S_trial(k) = S(k) + dS(k) and E = sum_k ffW |S_trial|^2 * EPS0_INV_real * TWOPI / V over the entries of fill_trial_tables.
// Every thread takes TWO k-vectors per iteration (i and i + NT): two independent dependency chains through the table
// look-ups and complex products, where one chain per thread left the group latency-bound (r02k capture: 41 % of this
// function's samples were fixed-latency waits).  The next pair's index triples, weights and S(k) are in flight meanwhile.
template <int NT>
__device__ double kspace_entries(const Smem &S, int ne, const double *__restrict__ S_in, double *__restrict__ S_out)
{
    const int nk = c_sys.nk, KW = c_sys.kmax_max + 1;
    const double2 *__restrict__ tab = S.tab_old;
    const int32_t *__restrict__ gkx = c_sys.kx, *__restrict__ gky = c_sys.ky, *__restrict__ gkz = c_sys.kz;
    double part[1] = { 0.0 };
    // S(k) of the next pair is in flight (per-walker data: L2 latency) while this pair is evaluated; the k-vector
    // triples and weights are shared by every walker of the SM and come from L1
    int i = Grp<NT>::tid();
    double reA = 0.0, imA = 0.0, reB = 0.0, imB = 0.0;
    if (i < nk) { reA = S_in[i]; imA = S_in[nk + i]; const int ib = min(i + NT, nk - 1); reB = S_in[ib]; imB = S_in[nk + ib]; }
    while (i < nk) {
        const int ib = min(i + NT, nk - 1);
        const bool hasB = i + NT < nk;
        const int kxA = __ldg(gkx + i), kyA = __ldg(gky + i), kzA = __ldg(gkz + i);
        const int kxB = __ldg(gkx + ib), kyB = __ldg(gky + ib), kzB = __ldg(gkz + ib);
        const int in = i + 2 * NT;
        double nreA = 0.0, nimA = 0.0, nreB = 0.0, nimB = 0.0;
        if (in < nk) { nreA = S_in[in]; nimA = S_in[nk + in]; const int inb = min(in + NT, nk - 1); nreB = S_in[inb]; nimB = S_in[nk + inb]; }
        const int o2A = KW + abs(kyA), o3A = 2 * KW + abs(kzA), syA = kyA & 0x80000000, szA = kzA & 0x80000000;
        const int o2B = KW + abs(kyB), o3B = 2 * KW + abs(kzB), syB = kyB & 0x80000000, szB = kzB & 0x80000000;
        double srA = 0.0, siA = 0.0, srB = 0.0, siB = 0.0;
        const double2 *t = tab;
#pragma unroll 2
        for (int e = 0; e < ne; ++e, t += 3 * KW) {
            entry_term(t, kxA, o2A, o3A, syA, szA, srA, siA);
            entry_term(t, kxB, o2B, o3B, syB, szB, srB, siB);
        }
        reA += srA; imA += siA; reB += srB; imB += siB;
        if (S_out) {
            S_out[i] = reA; S_out[nk + i] = imA;
            if (hasB) { S_out[ib] = reB; S_out[nk + ib] = imB; }
        }
        part[0] += __ldg(c_sys.ffW + i) * (reA * reA + imA * imA);
        if (hasB) part[0] += __ldg(c_sys.ffW + ib) * (reB * reB + imB * imB);
        i = in; reA = nreA; imA = nimA; reB = nreB; imB = nimB;
    }
    Grp<NT>::template sum<1>(part, S.ws->red);
    return part[0] * c_sys.eps0_inv_real * c_sys.twopi / c_sys.volume;
}

// reciprocal_ewald_energy (ewald_energy.f90:139-164) of a stored S(k)
template <int NT>
__device__ double recip_energy(const double *Sk, double *red)
{
    const int nk = c_sys.nk;
    double part[1] = { 0.0 };
    for (int i = Grp<NT>::tid(); i < nk; i += NT) {
        const double re = Sk[i], im = Sk[nk + i];
        part[0] += c_sys.ffW[i] * (re * re + im * im);
    }
    Grp<NT>::template sum<1>(part, red);
    return part[0] * c_sys.eps0_inv_real * c_sys.twopi / c_sys.volume;
}

// ------------------------------------------------------------------------------------
// trial evaluation
// ------------------------------------------------------------------------------------
template <int NT> __device__ __forceinline__ void stage_counts(const Smem &S, int w)
{
    if (Grp<NT>::tid() < MGPU_MAX_RES) S.ws->count[Grp<NT>::tid()] = c_sys.count[(int64_t)w * MGPU_MAX_RES + Grp<NT>::tid()];
}

// load the committed geometry of (res, mol) of walker w as absolute atom positions
__device__ __forceinline__ void load_positions(int w, int res, int mol, double (*pos)[3], int a)
{
    const double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride + c_sys.goff[res];
    const int cap = c_sys.cap[res];
    const double *off = wc + 3 * (int64_t)cap + (int64_t)a * 3 * cap;
#pragma unroll
    for (int d = 0; d < 3; ++d) pos[a][d] = wc[d * cap + mol] + off[d * cap + mol];
}

// Stage the probe for a trial of `kind` on (res, mol) with new geometry (com, off).
// (xres, xmol) >= 0: the target slot to leave out instead of (res, mol) -- the creation half of a
// swap, whose new molecule of type `res` must not see the molecule it replaces.
template <int NT>
__device__ __forceinline__ void stage_probe(const Smem &S, int w, int kind, int res, int mol,
                                            const double *com, const double (*off)[3], int xres = -1, int xmol = -1)
{
    Probe &P = S.ws->probe;
    const int na = c_sys.natom[res];
    const int t = Grp<NT>::tid();
    if (t == 0) {
        P.kind = kind; P.res = res; P.mol = mol; P.na = na;
        P.has_old = (kind != MGPU_KIND_CREATE);
        P.has_new = (kind != MGPU_KIND_DELETE);
        P.excl_res = (xres >= 0) ? xres : res; P.excl_mol = (xres >= 0) ? xmol : mol;
        P.order_res = -1; P.order_mol = -1;
        P.host_old = P.has_old && !c_sys.use_hcache;
        P.host_new = P.has_new;
    }
    if (t < na) {
        const double q = c_sys.charge[res][t];
        P.q[t] = (fabs(q) < MGPU_ERR_TOL) ? 0.0 : q;
        P.type[t] = c_sys.type[res][t];
        if (kind != MGPU_KIND_CREATE) load_positions(w, res, mol, P.po, t);
        if (kind != MGPU_KIND_DELETE) {
#pragma unroll
            for (int d = 0; d < 3; ++d) P.pn[t][d] = com[d] + off[t][d];
        }
    }
}

// Evaluate one trial held in the group's probe (positions staged): fills e_old / e_new
// (6 each, valid in every thread) and writes S_trial into the walker's non-committed buffer
// (store_S) -- compute_old_energy / compute_new_energy, monte_carlo_utils.f90:300-423.
// hc_new = framework part of the new geometry {lj, coulomb (e^2/A)}: the cache entry if the trial is committed.
template <bool TRI, int NT>
__device__ void evaluate_trial(const Smem &S, int w, bool store_S, double e_old[6], double e_new[6], double hc_new[2], PairCount &pc,
                               bool do_kspace = true)
{
    const Probe &P = S.ws->probe;
    double ps[8];
    pair_sums<TRI, NT>(S, w, ps, pc);
    if (P.has_old && !P.host_old) {          // framework sum of the committed geometry: cached at its own commit / rebuild
        const double *hc = hcache_rows(w, P.res);
        ps[0] = hc[P.mol]; ps[1] = hc[c_sys.cap[P.res] + P.mol];
    }
    hc_new[0] = ps[2]; hc_new[1] = ps[3];
    const double recip_cur = c_sys.energy[(int64_t)w * 6 + MGPU_E_RECIP];
    double recip_new = recip_cur;
    if (do_kspace) {
        if (NT != MGPU_BLOCK && S.ws->sync_k) phase_barrier<NT>(S.ws->sync_k);      // k_sweep, MGPU_OPT_PHASE_SYNC bit 3: re-align before k-space
        const int cur = c_sys.cur[w];
        const double *S_in = c_sys.S + ((int64_t)w * 2 + cur) * 2 * c_sys.nk;
        double *S_out = store_S ? c_sys.S + ((int64_t)w * 2 + (cur ^ 1)) * 2 * c_sys.nk : nullptr;
        const int ne = fill_trial_tables<NT>(S, P);
        Grp<NT>::sync();
        recip_new = kspace_entries<NT>(S, ne, S_in, S_out);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) { e_old[i] = 0.0; e_new[i] = 0.0; }
    e_old[MGPU_E_RECIP] = recip_cur;
    e_new[MGPU_E_RECIP] = recip_new;
    if (P.has_old) {
        e_old[MGPU_E_NON_COULOMB] = ps[0] + ps[4];
        e_old[MGPU_E_COULOMB] = (ps[1] + ps[5]) * c_sys.eps0_inv_real;
    }
    if (P.has_new) {
        e_new[MGPU_E_NON_COULOMB] = ps[2] + ps[6];
        e_new[MGPU_E_COULOMB] = (ps[3] + ps[7]) * c_sys.eps0_inv_real;
    }
    if (P.kind != MGPU_KIND_MOVE) {
        // rigid molecule: one thread evaluates the 1/2 na (na-1) intramolecular pairs, then broadcast
        double intra = 0.0;
        if (Grp<NT>::tid() == 0) intra = intra_energy<TRI>(P.res, P.kind == MGPU_KIND_CREATE ? P.pn : P.po);
        if (NT == 32) intra = __shfl_sync(0xffffffffu, intra, 0);
        else {
            Grp<NT>::sync();
            if (Grp<NT>::tid() == 0) S.ws->red[0] = intra;
            Grp<NT>::sync();
            intra = S.ws->red[0];
        }
        if (P.kind == MGPU_KIND_CREATE) { e_new[MGPU_E_SELF] = c_sys.e_self[P.res]; e_new[MGPU_E_INTRA] = intra; }
        else { e_old[MGPU_E_SELF] = c_sys.e_self[P.res]; e_old[MGPU_E_INTRA] = intra; }
    }
    e_old[MGPU_E_TOTAL] = e_old[0] + e_old[1] + e_old[2] + e_old[3] + e_old[4];
    e_new[MGPU_E_TOTAL] = e_new[0] + e_new[1] + e_new[2] + e_new[3] + e_new[4];
}

// A molecule whose geometry came from the host (mgpu_new_energy / mgpu_trial_batch) may be larger than anything seen so
// far: keep the per-residue bound of |offset|^2 (screen of the "nothing" lists) valid.  Device-generated geometries are
// rotations of stored ones and never grow it beyond its 1e-6 margin.
__device__ __forceinline__ void grow_rmax2(int res, const double (*off)[3], int na, int tid, int nthreads)
{
    for (int a = tid; a < na; a += nthreads) {
        const double r2 = off[a][0] * off[a][0] + off[a][1] * off[a][1] + off[a][2] * off[a][2];
        if (r2 > c_sys.rmax2[res]) atomicMax(reinterpret_cast<unsigned long long *>(c_sys.rmax2 + res), (unsigned long long)__double_as_longlong(r2));
    }
}

// Apply an accepted trial to the walker (group-cooperative).
// accept_molecule_move / accept_creation_move / accept_deletion_move + remove_molecule +
// update_counts (monte_carlo_utils.f90:429-442,642-672; creation.f90:82-116; deletion.f90:83-122).
// n = the walker's count of `res` BEFORE the commit, read by the caller while no thread of the group can have
// written it yet (thread 0 overwrites c_sys.count below; with more than one warp a late reader would otherwise
// pick the wrong "last" slot).
__device__ void commit_trial(int w, int kind, int res, int mol, int n, const double *com, const double (*off)[3],
                             const double e_old[6], const double e_new[6], const double hc_new[2], int tid, int nthreads)
{
    double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride + c_sys.goff[res];
    const int cap = c_sys.cap[res], na = c_sys.natom[res];
    double *offs = wc + 3 * (int64_t)cap;
    if (kind == MGPU_KIND_DELETE) {
        const int last = n - 1;
        if (mol != last) {                  // the last molecule moves into the hole, with its cache rows (na*3, na*3+1)
            if (tid < 3) wc[tid * cap + mol] = wc[tid * cap + last];
            for (int e = tid; e < na * 3 + 2; e += nthreads) offs[(int64_t)e * cap + mol] = offs[(int64_t)e * cap + last];
        }
    } else {
        if (tid < 3) wc[tid * cap + mol] = com[tid];
        for (int e = tid; e < na * 3; e += nthreads) offs[(int64_t)e * cap + mol] = off[e / 3][e % 3];
        if (tid < 2) offs[(int64_t)(na * 3 + tid) * cap + mol] = hc_new[tid];
        grow_rmax2(res, off, na, tid, nthreads);
    }
    if (tid == 0) {
        if (kind == MGPU_KIND_CREATE) c_sys.count[(int64_t)w * MGPU_MAX_RES + res] = n + 1;
        if (kind == MGPU_KIND_DELETE) c_sys.count[(int64_t)w * MGPU_MAX_RES + res] = n - 1;
        c_sys.cur[w] ^= 1;
        double *E = c_sys.energy + (int64_t)w * 6;
        E[MGPU_E_RECIP] = e_new[MGPU_E_RECIP];
        E[MGPU_E_NON_COULOMB] = E[MGPU_E_NON_COULOMB] + e_new[MGPU_E_NON_COULOMB] - e_old[MGPU_E_NON_COULOMB];
        E[MGPU_E_COULOMB] = E[MGPU_E_COULOMB] + e_new[MGPU_E_COULOMB] - e_old[MGPU_E_COULOMB];
        if (kind != MGPU_KIND_MOVE) {
            E[MGPU_E_SELF] = E[MGPU_E_SELF] + e_new[MGPU_E_SELF] - e_old[MGPU_E_SELF];
            E[MGPU_E_INTRA] = E[MGPU_E_INTRA] + e_new[MGPU_E_INTRA] - e_old[MGPU_E_INTRA];
        }
        E[MGPU_E_TOTAL] = E[MGPU_E_TOTAL] + e_new[MGPU_E_TOTAL] - e_old[MGPU_E_TOTAL];
    }
}

// attempt_swap_move, swapping.f90:34-105: old = deletion-style energies of molecule (resA, molA)
// (compute_old_energy(..., is_deletion), :62); new = creation-style energies of a molecule of type
// resB with geometry (com, off) in slot count(resB), evaluated with (resA, molA) gone from the
// coordinates (:69-88) but on top of an S(k) that STILL contains it -- the reference never applies
// the deletion dS of the swapped-out molecule; S_trial = S + dS(B).
// The probe is left staged for the creation half.  count[] = the walker's counts before the swap.
template <bool TRI, int NT>
__device__ void evaluate_swap(const Smem &S, int w, bool store_S, int resA, int molA, int resB, int molB,
                              const double *com, const double (*off)[3],
                              double e_old[6], double e_new[6], double hc_new[2], PairCount &pc)
{
    double e_tmp[6], hc_tmp[2];
    stage_probe<NT>(S, w, MGPU_KIND_DELETE, resA, molA, nullptr, nullptr);
    const int sq = S.ws->sync_q;                     // one quartet barrier per MC step inside pair_sums: the creation half takes it
    Grp<NT>::sync();
    if (Grp<NT>::tid() == 0) S.ws->sync_q = 0;
    Grp<NT>::sync();
    evaluate_trial<TRI, NT>(S, w, false, e_old, e_tmp, hc_tmp, pc, false);
    Grp<NT>::sync();
    if (Grp<NT>::tid() == 0) S.ws->sync_q = sq;
    stage_probe<NT>(S, w, MGPU_KIND_CREATE, resB, molB, com, off, resA, molA);
    Grp<NT>::sync();
    evaluate_trial<TRI, NT>(S, w, store_S, e_tmp, e_new, hc_new, pc, true);
}

// accept_swap_move (swapping.f90:143-155) + the coordinate / count changes the reference made
// before the accept test (:66-81): (resA, molA) removed by swap-with-last, the new molecule of type
// resB appended, S flipped, all five components updated.
// nA, nB = the walker's counts of the two types BEFORE the commit (read by the caller, see commit_trial).
__device__ void commit_swap(int w, int resA, int molA, int nA, int resB, int nB, const double *com, const double (*off)[3],
                            const double e_old[6], const double e_new[6], const double hc_new[2], int tid, int nthreads)
{
    {
        double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride + c_sys.goff[resA];
        const int cap = c_sys.cap[resA], na = c_sys.natom[resA];
        double *offs = wc + 3 * (int64_t)cap;
        const int last = nA - 1;
        if (molA != last) {
            if (tid < 3) wc[tid * cap + molA] = wc[tid * cap + last];
            for (int e = tid; e < na * 3 + 2; e += nthreads) offs[(int64_t)e * cap + molA] = offs[(int64_t)e * cap + last];
        }
    }
    {
        double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride + c_sys.goff[resB];
        const int cap = c_sys.cap[resB], na = c_sys.natom[resB];
        double *offs = wc + 3 * (int64_t)cap;
        if (tid < 3) wc[tid * cap + nB] = com[tid];
        for (int e = tid; e < na * 3; e += nthreads) offs[(int64_t)e * cap + nB] = off[e / 3][e % 3];
        if (tid < 2) offs[(int64_t)(na * 3 + tid) * cap + nB] = hc_new[tid];
        grow_rmax2(resB, off, na, tid, nthreads);
    }
    if (tid == 0) {
        c_sys.count[(int64_t)w * MGPU_MAX_RES + resA] = nA - 1;
        c_sys.count[(int64_t)w * MGPU_MAX_RES + resB] = nB + 1;
        c_sys.cur[w] ^= 1;
        double *E = c_sys.energy + (int64_t)w * 6;
        E[MGPU_E_RECIP] = e_new[MGPU_E_RECIP];
        E[MGPU_E_NON_COULOMB] = E[MGPU_E_NON_COULOMB] + e_new[MGPU_E_NON_COULOMB] - e_old[MGPU_E_NON_COULOMB];
        E[MGPU_E_COULOMB] = E[MGPU_E_COULOMB] + e_new[MGPU_E_COULOMB] - e_old[MGPU_E_COULOMB];
        E[MGPU_E_SELF] = E[MGPU_E_SELF] + e_new[MGPU_E_SELF] - e_old[MGPU_E_SELF];
        E[MGPU_E_INTRA] = E[MGPU_E_INTRA] + e_new[MGPU_E_INTRA] - e_old[MGPU_E_INTRA];
        E[MGPU_E_TOTAL] = E[MGPU_E_TOTAL] + e_new[MGPU_E_TOTAL] - e_old[MGPU_E_TOTAL];
    }
}

// Coordinates of a slot only (no counts / energies / S).  The reference writes a created
// molecule into slot N+1 before the accept test and does not undo it on reject
// (reject_creation_move, monte_carlo_utils.f90:579-591); that is observable when N = 0,
// because slot 1 is the geometry template of the next insertion (:563-564).
__device__ __forceinline__ void write_slot(int w, int res, int mol, const double *com, const double (*off)[3], int tid, int nthreads)
{
    double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride + c_sys.goff[res];
    const int cap = c_sys.cap[res], na = c_sys.natom[res];
    double *offs = wc + 3 * (int64_t)cap;
    if (tid < 3) wc[tid * cap + mol] = com[tid];
    for (int e = tid; e < na * 3; e += nthreads) offs[(int64_t)e * cap + mol] = off[e / 3][e % 3];
    grow_rmax2(res, off, na, tid, nthreads);
}

__device__ __forceinline__ void flush_pair_count(const PairCount &pc)
{
    const unsigned g = __reduce_add_sync(0xffffffffu, pc.geom), l = __reduce_add_sync(0xffffffffu, pc.lj),
                   c = __reduce_add_sync(0xffffffffu, pc.coul), x = __reduce_add_sync(0xffffffffu, pc.scr);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(c_sys.pair_count + 0, (unsigned long long)g);
        atomicAdd(c_sys.pair_count + 1, (unsigned long long)l);
        atomicAdd(c_sys.pair_count + 2, (unsigned long long)c);
        atomicAdd(c_sys.pair_count + 3, (unsigned long long)x);
    }
}

// ------------------------------------------------------------------------------------
// host-driven trial batch: one group per task (NT = 32: MGPU_WARPS tasks per CTA)
// ------------------------------------------------------------------------------------
struct TaskArrays {
    const int4 *meta;      // [n] {walker, res, mol, kind}
    const double *com;     // [n][3]
    const double *off;     // [n][MGPU_MAX_SITES][3]
    double *out;           // [n][12] e_old, e_new
};

template <bool TRI, int NT>
__global__ void __launch_bounds__(NT == 32 ? MGPU_WBLOCK : MGPU_BLOCK, 1) k_trial(TaskArrays T, int n_tasks, int natom_max)
{
    const Smem S = smem_setup<TabRep<NT>::v, NT>(mgpu_smem, natom_max);
    const int t = blockIdx.x * (NT == 32 ? (int)(blockDim.x >> 5) : 1) + Grp<NT>::id();
    if (t >= n_tasks) return;                       // whole groups leave together (no later CTA barrier for NT = 32)
    const int4 meta = T.meta[t];
    const int w = meta.x, res = meta.y, mol = meta.z, kind = meta.w & 0xff, res2 = meta.w >> 8;   // res2: new type of a swap
    const int gt = Grp<NT>::tid();
    stage_counts<NT>(S, w);
    const double *com = T.com + (int64_t)t * 3;
    const double(*off)[3] = reinterpret_cast<const double(*)[3]>(T.off + (int64_t)t * MGPU_MAX_SITES * 3);
    double e_old[6], e_new[6], hc_new[2];
    PairCount pc = { 0u, 0u, 0u, 0u };
    if (kind == MGPU_KIND_SWAP) {
        Grp<NT>::sync();
        evaluate_swap<TRI, NT>(S, w, true, res, mol, res2, S.ws->count[res2], com, off, e_old, e_new, hc_new, pc);
    } else {
        stage_probe<NT>(S, w, kind, res, mol, com, off);
        Grp<NT>::sync();
        evaluate_trial<TRI, NT>(S, w, true, e_old, e_new, hc_new, pc);
    }
    flush_pair_count(pc);
    // record the pending trial
    MgpuTrial *tr = c_sys.trial + w;
    if (gt == 0) {
        tr->active = 1; tr->kind = kind; tr->res = res; tr->mol = mol; tr->res2 = res2;
        for (int d = 0; d < 3; ++d) tr->com[d] = (kind == MGPU_KIND_DELETE) ? 0.0 : com[d];
        tr->hc_new[0] = hc_new[0]; tr->hc_new[1] = hc_new[1];
        for (int i = 0; i < 6; ++i) { tr->e_old[i] = e_old[i]; tr->e_new[i] = e_new[i]; T.out[(int64_t)t * 12 + i] = e_old[i]; T.out[(int64_t)t * 12 + 6 + i] = e_new[i]; }
    }
    if (kind != MGPU_KIND_DELETE) {
        const int na_new = c_sys.natom[kind == MGPU_KIND_SWAP ? res2 : res];
        for (int e = gt; e < na_new * 3; e += NT) tr->off[e / 3][e % 3] = off[e / 3][e % 3];
    }
}

__global__ void __launch_bounds__(128) k_commit(const int32_t *walker, const int32_t *accept, int32_t *err)
{
    const int t = blockIdx.x;
    const int w = walker[t];
    MgpuTrial *tr = c_sys.trial + w;
    if (!tr->active) { if (threadIdx.x == 0) atomicExch(err, 1); return; }
    // every thread reads the counts, THEN the CTA meets, and only then may thread 0 overwrite them: molecules of
    // 11+ sites spread the row copies over several warps (na * 3 + 2 > 32)
    const int nA = c_sys.count[(int64_t)w * MGPU_MAX_RES + tr->res];
    const int nB = (tr->kind == MGPU_KIND_SWAP) ? c_sys.count[(int64_t)w * MGPU_MAX_RES + tr->res2] : 0;
    __syncthreads();
    if (accept[t]) {
        if (tr->kind == MGPU_KIND_SWAP) commit_swap(w, tr->res, tr->mol, nA, tr->res2, nB, tr->com, tr->off, tr->e_old, tr->e_new, tr->hc_new, threadIdx.x, blockDim.x);
        else commit_trial(w, tr->kind, tr->res, tr->mol, nA, tr->com, tr->off, tr->e_old, tr->e_new, tr->hc_new, threadIdx.x, blockDim.x);
    } else if (tr->kind == MGPU_KIND_CREATE && tr->mol == 0) write_slot(w, tr->res, 0, tr->com, tr->off, threadIdx.x, blockDim.x);
    __syncthreads();
    if (threadIdx.x == 0) tr->active = 0;
}

// pair energy of one guest molecule with optional explicit geometry and ordering check
template <bool TRI>
__global__ void __launch_bounds__(MGPU_BLOCK) k_pair_molecule(int w, int res, int mol, int skip_ordering,
                                                              const double *geom /* com[3] + off[na][3] or NULL */,
                                                              double *out2, int natom_max)
{
    const Smem S = smem_setup<1, MGPU_BLOCK>(mgpu_smem, natom_max);
    stage_counts<MGPU_BLOCK>(S, w);
    Probe &P = S.ws->probe;
    const int na = c_sys.natom[res];
    if (threadIdx.x == 0) {
        P.kind = MGPU_KIND_DELETE; P.res = res; P.mol = mol; P.na = na; P.has_old = 1; P.has_new = 0;
        P.host_old = 1; P.host_new = 0;
        P.excl_res = res; P.excl_mol = mol;
        P.order_res = skip_ordering ? -1 : res; P.order_mol = skip_ordering ? -1 : mol;
    }
    if (threadIdx.x < na) {
        const int a = threadIdx.x;
        const double q = c_sys.charge[res][a];
        P.q[a] = (fabs(q) < MGPU_ERR_TOL) ? 0.0 : q; P.type[a] = c_sys.type[res][a];
        if (geom) { for (int d = 0; d < 3; ++d) P.po[a][d] = geom[d] + geom[3 + a * 3 + d]; }
        else load_positions(w, res, mol, P.po, a);
    }
    __syncthreads();
    double ps[8];
    PairCount pc = { 0u, 0u, 0u, 0u };
    pair_sums<TRI, MGPU_BLOCK>(S, w, ps, pc);
    if (threadIdx.x == 0) { out2[0] = ps[0] + ps[4]; out2[1] = (ps[1] + ps[5]) * c_sys.eps0_inv_real; }
}

template <bool TRI>
__global__ void k_intra(int w, int res, int mol, const double *geom, double *out)
{
    __shared__ double pos[MGPU_MAX_SITES][3];
    const int na = c_sys.natom[res];
    if (threadIdx.x < na) {
        const int a = threadIdx.x;
        if (geom) { for (int d = 0; d < 3; ++d) pos[a][d] = geom[d] + geom[3 + a * 3 + d]; }
        else load_positions(w, res, mol, pos, a);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[0] = intra_energy<TRI>(res, pos);
}

// ------------------------------------------------------------------------------------
// S(k) construction (compute_all_recip_amplitude, ewald_energy.f90:261-328)
// grid = (k blocks, walkers or 1).  mode 0: host atoms -> S_host.  mode 1: S[w] = S_host + guests.
// ------------------------------------------------------------------------------------
#define MGPU_STILE MGPU_MAX_SITES
__global__ void __launch_bounds__(MGPU_BLOCK) k_build_S(int mode, double *S_host_out, int first_walker)
{
    double2 *tab = reinterpret_cast<double2 *>(mgpu_smem);
    __shared__ double pos[MGPU_STILE][3];
    __shared__ double qs[MGPU_STILE];
    const int nk = c_sys.nk;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < nk;
    int kx = 0, ky = 0, kz = 0;
    if (live) { kx = c_sys.kx[i]; ky = c_sys.ky[i]; kz = c_sys.kz[i]; }
    double re = 0.0, im = 0.0;
    if (mode == 0) {
        for (int base = 0; base < c_sys.n_host; base += MGPU_STILE) {
            const int n = min(MGPU_STILE, c_sys.n_host - base);
            __syncthreads();
            if (threadIdx.x < n) {
                const double4 t = c_sys.host_xyzq[base + threadIdx.x];
                pos[threadIdx.x][0] = t.x; pos[threadIdx.x][1] = t.y; pos[threadIdx.x][2] = t.z; qs[threadIdx.x] = c_sys.host_qraw[base + threadIdx.x];
            }
            __syncthreads();
            fill_phase_tables(tab, pos, n, threadIdx.x, blockDim.x);
            __syncthreads();
            if (live)
                for (int a = 0; a < n; ++a) {
                    const cplx p = phase_product(tab, a, kx, ky, kz);
                    re += qs[a] * p.re; im += qs[a] * p.im;
                }
        }
        if (live) { S_host_out[i] = re; S_host_out[nk + i] = im; }
        return;
    }
    const int w = first_walker + blockIdx.y;
    if (live) { re = c_sys.S_host[i]; im = c_sys.S_host[nk + i]; }
    for (int g = 0; g < c_sys.nres; ++g) {
        if (!c_sys.active[g]) continue;
        const int n = c_sys.count[(int64_t)w * MGPU_MAX_RES + g], na = c_sys.natom[g];
        for (int m = 0; m < n; ++m) {
            __syncthreads();
            if (threadIdx.x < na) { load_positions(w, g, m, pos, threadIdx.x); qs[threadIdx.x] = c_sys.charge[g][threadIdx.x]; }
            __syncthreads();
            fill_phase_tables(tab, pos, na, threadIdx.x, blockDim.x);
            __syncthreads();
            if (live)
                for (int a = 0; a < na; ++a) {
                    const cplx p = phase_product(tab, a, kx, ky, kz);
                    re += qs[a] * p.re; im += qs[a] * p.im;
                }
        }
    }
    if (live) {
        double *Sk = c_sys.S + ((int64_t)w * 2 + c_sys.cur[w]) * 2 * nk;
        Sk[i] = re; Sk[nk + i] = im;
    }
}

// static host-host pair energy (pairs of host atoms that belong to different molecules),
// exact formulas, once at init
template <bool TRI>
__global__ void __launch_bounds__(MGPU_BLOCK) k_host_host(const int32_t *host_mol, double *partial /* [grid][2] */)
{
    __shared__ double red[2 * MGPU_WARPS];
    double acc[2] = { 0.0, 0.0 };
    const int n = c_sys.n_host, nt = c_sys.ntypes;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double4 a = c_sys.host_xyzq[i];
        const int ta = c_sys.host_type[i], ma = host_mol[i];
        for (int j = i + 1; j < n; ++j) {
            if (host_mol[j] == ma) continue;
            const double4 b = c_sys.host_xyzq[j];
            const int ti = ta * nt + c_sys.host_type[j];
            const double qq = a.w * b.w;
            const double r2 = min_image_r2<TRI>(b.x - a.x, b.y - a.y, b.z - a.z);
            const double2 e = pair_exact(r2, c_sys.ljA[ti], c_sys.ljB[ti], qq, qq != 0.0);
            acc[0] += e.x; acc[1] += e.y;
        }
    }
    Grp<MGPU_BLOCK>::sum<2>(acc, red);
    if (threadIdx.x == 0) { partial[blockIdx.x * 2] = acc[0]; partial[blockIdx.x * 2 + 1] = acc[1]; }
}

// ------------------------------------------------------------------------------------
// K4: full energy of a walker (update_system_energy, energy_utils.f90:22-39).
// S(k) of the walker must have been rebuilt by k_build_S.  One CTA per walker.
// ------------------------------------------------------------------------------------
template <bool TRI>
__global__ void __launch_bounds__(MGPU_BLOCK) k_total_energy(int first_walker, int natom_max)
{
    const Smem S = smem_setup<1, MGPU_BLOCK>(mgpu_smem, natom_max);
    const int w = first_walker + blockIdx.x;
    stage_counts<MGPU_BLOCK>(S, w);
    __syncthreads();
    Probe &P = S.ws->probe;
    double e_lj = c_sys.hh_lj, e_c = 0.0, e_intra = 0.0, e_self = c_sys.self_host_total;
    PairCount pc = { 0u, 0u, 0u, 0u };
    for (int g = 0; g < c_sys.nres; ++g) {
        if (!c_sys.active[g]) continue;
        const int n = S.ws->count[g], na = c_sys.natom[g];
        for (int m = 0; m < n; ++m) {
            __syncthreads();
            if (threadIdx.x == 0) {
                P.kind = MGPU_KIND_DELETE; P.res = g; P.mol = m; P.na = na; P.has_old = 1; P.has_new = 0;
                P.host_old = 1; P.host_new = 0;
                P.excl_res = g; P.excl_mol = m; P.order_res = g; P.order_mol = m;
            }
            if (threadIdx.x < na) {
                const double q = c_sys.charge[g][threadIdx.x];
                P.q[threadIdx.x] = (fabs(q) < MGPU_ERR_TOL) ? 0.0 : q; P.type[threadIdx.x] = c_sys.type[g][threadIdx.x];
                load_positions(w, g, m, P.po, threadIdx.x);
            }
            __syncthreads();
            double ps[8];
            pair_sums<TRI, MGPU_BLOCK>(S, w, ps, pc);
            e_lj += ps[0] + ps[4]; e_c += ps[1] + ps[5];
            if (threadIdx.x == 0) {
                e_intra += intra_energy<TRI>(g, P.po);
                double *hc = hcache_rows(w, g);              // (re)fill the molecule's framework-energy cache
                hc[m] = ps[0]; hc[c_sys.cap[g] + m] = ps[1];
            }
        }
        e_self += c_sys.e_self[g] * n;
    }
    const double *Sk = c_sys.S + ((int64_t)w * 2 + c_sys.cur[w]) * 2 * c_sys.nk;
    const double recip = recip_energy<MGPU_BLOCK>(Sk, S.ws->red);
    if (threadIdx.x == 0) {
        double *E = c_sys.energy + (int64_t)w * 6;
        E[MGPU_E_NON_COULOMB] = e_lj;
        E[MGPU_E_COULOMB] = e_c * c_sys.eps0_inv_real + c_sys.hh_coul;
        E[MGPU_E_RECIP] = recip;
        E[MGPU_E_SELF] = e_self;
        E[MGPU_E_INTRA] = e_intra;
        E[MGPU_E_TOTAL] = E[MGPU_E_RECIP] + E[MGPU_E_NON_COULOMB] + E[MGPU_E_COULOMB] + E[MGPU_E_SELF] + E[MGPU_E_INTRA];
    }
}

// ------------------------------------------------------------------------------------
// RNG contract: xoshiro256** (per walker), uniform = top 53 bits * 2^-53
// ------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64_mix(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
#define MGPU_GOLDEN 0x9E3779B97F4A7C15ULL
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
// state lives in shared memory (WalkerLocal::rng); only the group's thread 0 draws
__device__ __forceinline__ double rng_uniform(unsigned long long *s)
{
    const uint64_t s0 = s[0], s1 = s[1];
    uint64_t s2 = s[2], s3 = s[3];
    const uint64_t r = rotl64(s1 * 5u, 7) * 9u;
    const uint64_t t = s1 << 17;
    s2 ^= s0; s3 ^= s1;
    s[1] = s1 ^ s2; s[0] = s0 ^ s3;
    s2 ^= t; s3 = rotl64(s3, 45);
    s[2] = s2; s[3] = s3;
    return (double)(r >> 11) * 0x1.0p-53;
}

// return_rotation_matrix (helper_utils.f90:30-75) applied to offsets (matmul, j = 1..3 in order)
__device__ __forceinline__ void rotate_offsets(int axis, double theta, double (*off)[3], int na)
{
    double sn, c;
    sincos(theta, &sn, &c);
    double M[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
    if (axis == 1) { M[1][1] = c; M[1][2] = -sn; M[2][1] = sn; M[2][2] = c; }
    else if (axis == 2) { M[0][0] = c; M[0][2] = sn; M[2][0] = -sn; M[2][2] = c; }
    else { M[0][0] = c; M[0][1] = -sn; M[1][0] = sn; M[1][1] = c; }
    for (int a = 0; a < na; ++a) {
        const double v0 = off[a][0], v1 = off[a][1], v2 = off[a][2];
#pragma unroll
        for (int i = 0; i < 3; ++i)
            off[a][i] = __dadd_rn(__dadd_rn(__dmul_rn(M[i][0], v0), __dmul_rn(M[i][1], v1)), __dmul_rn(M[i][2], v2));
    }
}

// ------------------------------------------------------------------------------------
// device-resident Monte Carlo: ONE WARP PER WALKER, n_steps iterations of the body of
// monte_carlo_loop (monte_carlo.f90:50-99).  Draw order per step:
// residue pick, molecule pick (if count > 0), move select, [create/delete coin],
// proposal draws, accept draw.  Lane 0 plays the Fortran driver (RNG, proposal,
// Metropolis); all 32 lanes evaluate the energies.  Warps never meet at a CTA barrier
// after set-up, so a walker's serial section only idles its own warp.
// ------------------------------------------------------------------------------------
// lane 0: draw and propose one step into sh (move drivers' proposal halves)
__device__ __noinline__ void propose_step(int w, GroupWS &ws, int32_t *err)
{
    SweepShared &sh = ws.sh;
    unsigned long long *rng = ws.loc.rng;
    const double cumul_translation = c_sys.p_trans;
    const double cumul_rotation = cumul_translation + c_sys.p_rot;
    const double cumul_swap = cumul_rotation + c_sys.p_swap;
    sh.valid = 0; sh.move = MGPU_MV_NONE; sh.accept = 0; sh.traced_accept = 0; sh.prob = 0.0;
    // pick_random_residue_type / pick_random_molecule_index, monte_carlo_utils.f90:140-201
    const int res = c_sys.active_list[(int)(rng_uniform(rng) * c_sys.nactive)];
    const int n = ws.count[res];
    int mol = -1;
    if (n > 0) { mol = (int)(rng_uniform(rng) * n) + 1; if (mol > n) mol = n; mol -= 1; }
    const double draw = rng_uniform(rng);
    const int na = c_sys.natom[res];
    sh.res = res; sh.mol = mol;
    const double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride + c_sys.goff[res];
    const int cap = c_sys.cap[res];
    const double *offs = wc + 3 * (int64_t)cap;
    if (draw <= cumul_translation) {
        if (mol >= 0) {                                           // translation.f90:21-97
            sh.valid = 1; sh.move = MGPU_MV_TRANSLATE; sh.kind = MGPU_KIND_MOVE;
            double tp[3];
            for (int d = 0; d < 3; ++d) tp[d] = rng_uniform(rng);
            for (int d = 0; d < 3; ++d) sh.com[d] = wc[d * cap + mol] + (tp[d] - 0.5) * c_sys.step[(int64_t)w * 2];
            apply_PBC(sh.com);
            for (int e = 0; e < na * 3; ++e) sh.off[e / 3][e % 3] = offs[(int64_t)e * cap + mol];
        }
    } else if (draw <= cumul_rotation) {
        if (na != 1 && mol >= 0) {                                // rotation.f90:22-60
            sh.valid = 1; sh.move = MGPU_MV_ROTATE; sh.kind = MGPU_KIND_MOVE;
            for (int d = 0; d < 3; ++d) sh.com[d] = wc[d * cap + mol];
            for (int e = 0; e < na * 3; ++e) sh.off[e / 3][e % 3] = offs[(int64_t)e * cap + mol];
            const double theta = (rng_uniform(rng) - 0.5) * c_sys.step[(int64_t)w * 2 + 1];
            const int axis = (int)(rng_uniform(rng) * 3.0) + 1;
            rotate_offsets(axis, theta, sh.off, na);
        }
    } else if (draw <= cumul_swap) {                              // swapping.f90:34-105
        int res_bis = -1;                                         // pick_different_residue_type :165-197
        for (int attempt = 0; attempt < 10; ++attempt) {
            const int nt = c_sys.active_list[(int)(rng_uniform(rng) * c_sys.nactive)];
            if (nt != res) { res_bis = nt; break; }
        }
        // no molecule of the old type: the reference indexes molecule 0 (undefined); like the oracle, nothing is tried
        if (res_bis >= 0 && ws.count[res_bis] > 0 && mol >= 0) {
            const int nb = ws.count[res_bis], capb = c_sys.cap[res_bis], nab = c_sys.natom[res_bis];
            if (nb >= capb) atomicExch(err, 2);
            else {
                sh.valid = 1; sh.move = MGPU_MV_SWAP; sh.kind = MGPU_KIND_SWAP; sh.res2 = res_bis;
                for (int d = 0; d < 3; ++d) sh.com[d] = wc[d * cap + mol];            // CoM of the swapped-out molecule (:80)
                const double *offb = c_sys.coords + (int64_t)w * c_sys.coord_stride + c_sys.goff[res_bis] + 3 * (int64_t)capb;
                for (int e = 0; e < nab * 3; ++e) sh.off[e / 3][e % 3] = offb[(int64_t)e * capb + 0];   // geometry of molecule 1 of the new type
                if (nab != 1) {
                    const double theta = rng_uniform(rng) * c_sys.twopi;
                    const int axis = (int)(rng_uniform(rng) * 3.0) + 1;
                    rotate_offsets(axis, theta, sh.off, nab);
                }
            }
        }
    } else {
        bool create = false, widom = false, del = false;
        if (c_sys.p_insdel > 0) { if (rng_uniform(rng) <= 0.5) create = true; else del = true; }
        else if (c_sys.p_widom > 0) widom = true;
        if (create || widom) {                                    // creation.f90:30-76, widom.f90:30-68
            if (n >= cap) atomicExch(err, 2);
            else {
                sh.valid = 1; sh.move = widom ? MGPU_MV_WIDOM : MGPU_MV_CREATE; sh.kind = MGPU_KIND_CREATE;
                sh.mol = n;
                double t3[3];
                for (int d = 0; d < 3; ++d) t3[d] = rng_uniform(rng);
                for (int i = 0; i < 3; ++i)
                    sh.com[i] = c_sys.lo[i] + __dadd_rn(__dadd_rn(__dmul_rn(c_sys.H[i * 3 + 0], t3[0]), __dmul_rn(c_sys.H[i * 3 + 1], t3[1])),
                                                        __dmul_rn(c_sys.H[i * 3 + 2], t3[2]));
                for (int e = 0; e < na * 3; ++e) sh.off[e / 3][e % 3] = offs[(int64_t)e * cap + 0];   // geometry of molecule 1
                if (na != 1) {
                    const double theta = rng_uniform(rng) * c_sys.twopi;
                    const int axis = (int)(rng_uniform(rng) * 3.0) + 1;
                    rotate_offsets(axis, theta, sh.off, na);
                }
            }
        } else if (del) {
            if (n > 0) { sh.valid = 1; sh.move = MGPU_MV_DELETE; sh.kind = MGPU_KIND_DELETE; }   // deletion.f90:29-77
        }
    }
}

// lane 0: acceptance rule + counters (compute_acceptance_probability, monte_carlo_utils.f90:204-255;
// accumulate_widom_weight, widom.f90:74-92)
__device__ __noinline__ void decide_step(int w, GroupWS &ws, const double e_old[6], const double e_new[6])
{
    SweepShared &sh = ws.sh;
    long long *cnt = ws.loc.cnt;
    const int res = sh.res;
    const double dU = e_new[MGPU_E_TOTAL] - e_old[MGPU_E_TOTAL];
    const double mu = c_sys.mu[(int64_t)w * MGPU_MAX_RES + res];
    const double lam = c_sys.lambda[res];
    double p;
    int acc;
    if (sh.move == MGPU_MV_WIDOM) {
        p = exp(-dU * c_sys.beta);
        acc = 0;
        cnt[10] += 1;
        const int ok = p > MGPU_ERR_TOL;
        if (ok) { cnt[11] += 1; c_sys.widom_w[(int64_t)w * MGPU_MAX_RES + res] += p; }
        c_sys.widom_n[(int64_t)w * MGPU_MAX_RES + res] += 1;
        sh.traced_accept = ok;
    } else {
        if (sh.move == MGPU_MV_CREATE) {
            const double N = (double)(ws.count[res] + 1);        // count already incremented in the reference
            p = fmin(1.0, c_sys.volume / N / (lam * lam * lam) * exp(-c_sys.beta * (dU - mu)));
        } else if (sh.move == MGPU_MV_DELETE) {
            const double Np1 = (double)(ws.count[res] - 1) + 1.0;
            p = fmin(1.0, Np1 * (lam * lam * lam) / c_sys.volume * exp(-c_sys.beta * (dU + mu)));
        } else if (sh.move == MGPU_MV_SWAP) {
            // swap_acceptance_probability, monte_carlo_utils.f90:264-294, with the counts the reference has at
            // that point: old type already decremented, new type already incremented (swapping.f90:73-77)
            const double N_old = (double)(ws.count[res] - 1), Nplus1 = (double)(ws.count[sh.res2] + 1) + 1.0;
            const double mu_new = c_sys.mu[(int64_t)w * MGPU_MAX_RES + sh.res2];
            p = fmin(1.0, (N_old / Nplus1) * exp(-c_sys.beta * (dU + mu_new - mu)));
        } else p = fmin(1.0, exp(-c_sys.beta * dU));
        acc = rng_uniform(ws.loc.rng) <= p;
        const int ci = (sh.move == MGPU_MV_TRANSLATE) ? 0 : (sh.move == MGPU_MV_ROTATE) ? 1 : (sh.move == MGPU_MV_CREATE) ? 2
                     : (sh.move == MGPU_MV_DELETE) ? 3 : 4;
        cnt[2 * ci] += 1;
        if (acc) { cnt[2 * ci + 1] += 1; if (ci >= 2) cnt[2 * ci] += 1; }   // creations/deletions/swaps bump both slots
        sh.traced_accept = acc;
    }
    sh.prob = p;
    sh.accept = acc;
    for (int i = 0; i < 6; ++i) { sh.e_old[i] = e_old[i]; sh.e_new[i] = e_new[i]; }
}

// Phase alignment.  The four warps with the same (warp id mod 4) share an SM sub-partition, i.e. its
// issue port and its ~6 KB L0 instruction cache.  Left alone, the walkers of a CTA drift apart in the
// MC step (the r01k profile: 46 % of all stall samples were "no instruction", in every loop), because
// four warps in four different loops evict each other's code.  Named barriers per quartet keep those
// four warps in the same loop at the same time: at the top of every MC step, before the guest pass
// (in pair_sums) and before k-space (in evaluate_trial); a barrier before the energy evaluation is
// available too.  Measured (moves/s): none 10.9 M, top only 20.3 M, top + evaluation 20.6 M,
// + guest pass 21.8 M, + k-space 22.7 M, top + guest pass + k-space 22.8 M (default).  Quartets stay
// independent of each other; a step that evaluates nothing takes the same barriers, and a swap (two
// evaluations) takes them in its creation half, so every warp of a quartet arrives equally often.
// NT = 32: one warp per walker (the throughput shape: 16 walkers per CTA).  NT = MGPU_TEAM: a team of four warps per
// walker, used when a launch has fewer walkers than half the GPU's warp slots (strong scaling of a FIXED isotherm over
// more GPUs): the energy loops split over 128 threads, the driver part (RNG, proposal, Metropolis) still
// runs on one thread.  Same arithmetic per pair; sums are reduced in a different order (parity 1e-10, like the
// CTA-per-trial kernels that share these templates).
template <bool TRI, int NT>
__global__ void __launch_bounds__(MGPU_WBLOCK, 1) k_sweep(int first_walker, int n_walkers, long long n_steps, int natom_max,
                                                      int trace_walker, mgpu_step_trace *trace, int32_t *err, int phase_sync)
{
    const Smem S = smem_setup<MGPU_TAB_REP, NT>(mgpu_smem, natom_max);
    const int ngroups = (int)blockDim.x / NT;
    const int wl = blockIdx.x * ngroups + Grp<NT>::id();
    const bool live = wl < n_walkers;
    const int w = first_walker + (live ? wl : 0);
    const int nwarps = (int)(blockDim.x >> 5);
    // Phase alignment is for loaded pores.  When the CTA's walkers are nearly empty most steps evaluate nothing, and a CTA-wide
    // barrier would make every step as long as that of the one walker in sixteen that inserts (r03e: 51 against 100 M moves/s at
    // zero loading): a CTA where fewer than half of the walkers hold four molecules runs free (bit 16 of phase_sync asks for this
    // test; the counts are read once per launch).
    bool free_run = false;
    if (NT == 32 && (phase_sync & 16)) {
        int loaded = 0;
        if (live && (threadIdx.x & 31) == 0) {
            int c = 0;
            for (int r = 0; r < c_sys.nres; ++r) if (c_sys.active[r]) c += c_sys.count[(int64_t)w * MGPU_MAX_RES + r];
            loaded = c >= 4;
        }
        const int n_loaded = __syncthreads_count(loaded);
        const int n_live = min(ngroups, n_walkers - (int)blockIdx.x * ngroups);
        free_run = 2 * n_loaded < n_live;
    }
    if (!live && (free_run || !(phase_sync & 15))) return;
    {
        // threads that meet at a phase barrier: the warps of this warp's alignment group (NT = 32) or the whole CTA (teams).
        // The barrier set lives in the group's workspace (0 = not taken), so that the step loop carries no extra register:
        // with phase_sync itself rewritten in a register the kernel ran 6 % slower (r03i: 26.0 against 27.7 M moves/s).
        const int qthreads = free_run ? 0 : ((NT == 32) ? 32 * ((nwarps - (int)((threadIdx.x >> 5) & (MGPU_ALIGN_GROUPS - 1)) + MGPU_ALIGN_GROUPS - 1) / MGPU_ALIGN_GROUPS) : (int)blockDim.x);
        if (Grp<NT>::tid() == 0) {
            S.ws->sync_top = (phase_sync & 1) ? qthreads : 0; S.ws->sync_eval = (phase_sync & 2) ? qthreads : 0;
            S.ws->sync_q = (phase_sync & 4) ? qthreads : 0; S.ws->sync_k = (phase_sync & 8) ? qthreads : 0;
        }
    }
    const int lane = Grp<NT>::tid();                    // index inside the walker's group; 0 plays the Fortran driver
    GroupWS &ws = *S.ws;
    if (live) {
        stage_counts<NT>(S, w);
        if (lane < 4) ws.loc.rng[lane] = c_sys.rng[(int64_t)w * 4 + lane];
        if (lane < 12) ws.loc.cnt[lane] = 0;
        if (lane < MGPU_MAX_RES) { ws.loc.avgN[lane] = 0.0; ws.loc.avgN2[lane] = 0.0; }
        if (lane == 0) { ws.loc.avgE = 0.0; ws.loc.n_samples = 0; }
    }
    if (lane == 0) ws.sh.valid = 0;
    PairCount pc = { 0u, 0u, 0u, 0u };
    Grp<NT>::sync();

    for (long long step = 0; step < n_steps; ++step) {
        if (ws.sync_top) phase_barrier<NT>(ws.sync_top);
        if (!live) { if (ws.sync_eval) phase_barrier<NT>(ws.sync_eval); if (ws.sync_q) phase_barrier<NT>(ws.sync_q); if (ws.sync_k) phase_barrier<NT>(ws.sync_k); continue; }
        if (lane == 0) propose_step(w, ws, err);
        Grp<NT>::sync();
        if (ws.sync_eval) phase_barrier<NT>(ws.sync_eval);
        const SweepShared &sh = ws.sh;
        if (!sh.valid && ws.sync_q) phase_barrier<NT>(ws.sync_q);       // the barrier pair_sums would have taken
        if (!sh.valid && ws.sync_k) phase_barrier<NT>(ws.sync_k);       // ... and the one before k-space
        if (sh.valid) {
            double e_old[6], e_new[6], hc_new[2];
            if (sh.kind == MGPU_KIND_SWAP) {
                evaluate_swap<TRI, NT>(S, w, true, sh.res, sh.mol, sh.res2, ws.count[sh.res2], sh.com, sh.off, e_old, e_new, hc_new, pc);
            } else {
                stage_probe<NT>(S, w, sh.kind, sh.res, sh.mol, sh.com, sh.off);
                Grp<NT>::sync();
                evaluate_trial<TRI, NT>(S, w, sh.move != MGPU_MV_WIDOM, e_old, e_new, hc_new, pc);
            }
            if (lane == 0) decide_step(w, ws, e_old, e_new);
            Grp<NT>::sync();
            if (sh.accept) {
                const int nA = ws.count[sh.res], nB = (sh.kind == MGPU_KIND_SWAP) ? ws.count[sh.res2] : 0;
                Grp<NT>::sync();
                if (sh.kind == MGPU_KIND_SWAP) commit_swap(w, sh.res, sh.mol, nA, sh.res2, nB, sh.com, sh.off, sh.e_old, sh.e_new, hc_new, lane, NT);
                else commit_trial(w, sh.kind, sh.res, sh.mol, nA, sh.com, sh.off, sh.e_old, sh.e_new, hc_new, lane, NT);
                if (lane == 0) {
                    if (sh.kind == MGPU_KIND_CREATE) ws.count[sh.res] += 1;
                    if (sh.kind == MGPU_KIND_DELETE) ws.count[sh.res] -= 1;
                    if (sh.kind == MGPU_KIND_SWAP) { ws.count[sh.res] -= 1; ws.count[sh.res2] += 1; }
                }
            } else if (sh.kind == MGPU_KIND_CREATE && sh.mol == 0) {
                write_slot(w, sh.res, 0, sh.com, sh.off, lane, NT);      // rejected / Widom insertion into an empty walker
            }
        }
        Grp<NT>::sync();
        if (lane == 0) {
            if (trace && w == trace_walker) {
                mgpu_step_trace *t = trace + step;
                t->move = sh.move; t->res = sh.res; t->mol = sh.mol; t->accepted = sh.traced_accept;
                t->prob = sh.prob;
                if (sh.valid) {
                    t->dE = sh.e_new[MGPU_E_TOTAL] - sh.e_old[MGPU_E_TOTAL];
                    for (int i = 0; i < 6; ++i) { t->e_old[i] = sh.e_old[i]; t->e_new[i] = sh.e_new[i]; }
                } else {
                    t->dE = 0.0;
                    for (int i = 0; i < 6; ++i) { t->e_old[i] = 0.0; t->e_new[i] = 0.0; }
                }
            }
            for (int g = 0; g < c_sys.nres; ++g) {
                if (!c_sys.active[g]) continue;
                const double N = (double)ws.count[g];
                ws.loc.avgN[g] += N; ws.loc.avgN2[g] += N * N;
            }
            ws.loc.avgE += c_sys.energy[(int64_t)w * 6 + MGPU_E_TOTAL];
            ws.loc.n_samples += 1;
        }
        __threadfence_block();
        Grp<NT>::sync();     // commit visible (coordinates, S flip, counts) to every thread of the group before the next step
    }
    if (!live) return;
    flush_pair_count(pc);
    if (lane < 4) c_sys.rng[(int64_t)w * 4 + lane] = ws.loc.rng[lane];
    if (lane < 12) c_sys.counters[(int64_t)w * 12 + lane] += ws.loc.cnt[lane];
    if (lane < c_sys.nres) {
        double *A = c_sys.avg + ((int64_t)w * MGPU_MAX_RES + lane) * 4;
        A[0] += ws.loc.avgN[lane]; A[1] += ws.loc.avgN2[lane]; A[2] += ws.loc.avgE; A[3] += (double)ws.loc.n_samples;
    }
}

// adjust_move_step_sizes (monte_carlo_utils.f90:98-134): Robbins-Monro update of every walker's
// translation / rotation step from its own cumulative counters, at the end of a block
__global__ void k_adjust_steps(int first, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int w = first + i;
    const double gamma = 0.10, TARGET_ACCEPTANCE = 0.40;
    const long long *c = c_sys.counters + (int64_t)w * 12;
    double *st = c_sys.step + (int64_t)w * 2;
    if (c[0] > 500) {                                           // MIN_TRIALS_FOR_RECALIBRATION, parameters.f90:24
        const double acc = (double)c[1] / (double)c[0];
        double v = st[0] * exp(gamma * (acc - TARGET_ACCEPTANCE));
        st[0] = fmax(1.0e-3, fmin(v, 3.0));                      // MIN/MAX_TRANSLATION_STEP
    }
    if (c[2] > 500) {
        const double acc = (double)c[3] / (double)c[2];
        double v = st[1] * exp(gamma * (acc - TARGET_ACCEPTANCE));
        st[1] = fmax(1.0e-3, fmin(v, 0.78));                     // MIN/MAX_ROTATION_ANGLE
    }
}

// ------------------------------------------------------------------------------------
// K3: Widom batch, one warp per insertion (persistent warps, grid-stride over ids).
// u(id,k) = top53(mix(seed + GOLDEN*(8*id + k + 1))), k = 0..4: COM x3, angle, axis.
// Per-warp partial sums are written in warp order and added by the host in that order,
// so the result does not depend on scheduling.
// ------------------------------------------------------------------------------------
template <bool TRI>
__global__ void __launch_bounds__(MGPU_WBLOCK, 1) k_widom_batch(int w, int res, long long first_id, long long n,
                                                            unsigned long long seed, double *dE_out,
                                                            double *warp_sum_w, long long *warp_n_ok, int natom_max)
{
    const Smem S = smem_setup<MGPU_TAB_REP, 32>(mgpu_smem, natom_max);
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + Grp<32>::id();
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    GroupWS &ws = *S.ws;
    stage_counts<32>(S, w);
    __syncwarp();
    const int na = c_sys.natom[res];
    const int slot = ws.count[res];                  // the test molecule would occupy slot N+1
    const double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride;
    const int cap = c_sys.cap[res];
    const double *offs = wc + c_sys.goff[res] + 3 * (int64_t)cap;
    const double recip_cur = c_sys.energy[(int64_t)w * 6 + MGPU_E_RECIP];
    double my_w = 0.0; long long my_ok = 0;
    PairCount pc = { 0u, 0u, 0u, 0u };
    for (long long i = gw; i < n; i += nwarps) {
        const unsigned long long id = (unsigned long long)(first_id + i);
        if (lane == 0) {
            SweepShared &sh = ws.sh;
            double u[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) u[k] = (double)(splitmix64_mix(seed + MGPU_GOLDEN * (8ull * id + (unsigned long long)k + 1ull)) >> 11) * 0x1.0p-53;
            for (int d = 0; d < 3; ++d)
                sh.com[d] = c_sys.lo[d] + __dadd_rn(__dadd_rn(__dmul_rn(c_sys.H[d * 3 + 0], u[0]), __dmul_rn(c_sys.H[d * 3 + 1], u[1])),
                                                    __dmul_rn(c_sys.H[d * 3 + 2], u[2]));
            // geometry of molecule 1 of this residue type (insert_and_orient_molecule, :563-568)
            for (int e = 0; e < na * 3; ++e) sh.off[e / 3][e % 3] = offs[(int64_t)e * cap + 0];
            if (na != 1) {
                const double theta = u[3] * c_sys.twopi;
                const int axis = (int)(u[4] * 3.0) + 1;
                rotate_offsets(axis, theta, sh.off, na);
            }
        }
        __syncwarp();
        stage_probe<32>(S, w, MGPU_KIND_CREATE, res, slot, ws.sh.com, ws.sh.off);
        __syncwarp();
        double e_old[6], e_new[6], hc_new[2];
        evaluate_trial<TRI, 32>(S, w, false, e_old, e_new, hc_new, pc);
        if (lane == 0) {
            const double dU = e_new[MGPU_E_TOTAL] - recip_cur;
            if (dE_out) dE_out[i] = dU;
            const double wgt = exp(-dU * c_sys.beta);
            if (wgt > MGPU_ERR_TOL) { my_w += wgt; my_ok += 1; }
        }
        __syncwarp();
    }
    flush_pair_count(pc);
    if (lane == 0) { warp_sum_w[gw] = my_w; warp_n_ok[gw] = my_ok; }
}

// ------------------------------------------------------------------------------------
// numerical self-test of the fast pair math against the exact forms (one launch, tiny):
// out[0] = max relative error of rcp_fast, out[1] = max relative error of the Coulomb table
// ------------------------------------------------------------------------------------
__global__ void k_selftest(double s_lo, double s_hi, int n, double *out)
{
    double e_r = 0.0, e_g = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double s = s_lo * exp(log(s_hi / s_lo) * ((double)i + 0.5) / (double)n);
        e_r = fmax(e_r, fabs(rcp_fast(s) * s - 1.0));
        const int hi = __double2hiint(s);
        const int idx = (hi >> (20 - MGPU_TAB_K)) - c_sys.tab_ibase;
        if ((unsigned)idx < (unsigned)c_sys.tab_nint) {
            const int chi = (hi & ~((1 << (20 - MGPU_TAB_K)) - 1)) | (1 << (19 - MGPU_TAB_K));
            const double u = s - __hiloint2double(chi, 0);
            const double *row = c_sys.ctab + (size_t)idx * MGPU_TAB_ROW;
            const float uf = (float)u;
            double p = (double)fmaf(__int_as_float(__double2hiint(row[5])), uf, __int_as_float(__double2loint(row[5])));
            for (int k = 4; k >= 0; --k) p = fma(p, u, row[k]);
            const double r = sqrt(s), ref = erfc(c_sys.alpha * r) / r;
            e_g = fmax(e_g, fabs(p - ref) * r);     // error relative to the pair's Coulomb scale (g <= 1/r)
        }
    }
    // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long *>(out + 0), (unsigned long long)__double_as_longlong(e_r));
    atomicMax(reinterpret_cast<unsigned long long *>(out + 1), (unsigned long long)__double_as_longlong(e_g));
}

// ------------------------------------------------------------------------------------
// L2 read-bandwidth microbenchmark (roofline denominator of K2): grid-stride 16-byte loads over an L2-resident buffer
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_l2_read(const double2 *__restrict__ buf, size_t n16, int passes, double *out)
{
    double ax = 0.0, ay = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n16; i += 4 * stride) {          // four independent loads in flight per thread
            double2 v0, v1, v2, v3;
            asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v0.x), "=d"(v0.y) : "l"(buf + i));
            asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v1.x), "=d"(v1.y) : "l"(buf + i + stride));
            asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v2.x), "=d"(v2.y) : "l"(buf + i + 2 * stride));
            asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v3.x), "=d"(v3.y) : "l"(buf + i + 3 * stride));
            ax += v0.x + v1.x + v2.x + v3.x; ay += v0.y + v1.y + v2.y + v3.y;
        }
        for (; i < n16; i += stride) { const double2 v = buf[i]; ax += v.x; ay += v.y; }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ax + ay;
}

// ------------------------------------------------------------------------------------
// FP64 FMA peak microbenchmark (roofline denominator of K1/K3)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-9, x2 = x0 + 2e-9, x3 = x0 + 3e-9;
    double x4 = x0 + 4e-9, x5 = x0 + 5e-9, x6 = x0 + 6e-9, x7 = x0 + 7e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
