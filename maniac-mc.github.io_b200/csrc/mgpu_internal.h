// mgpu_internal.h -- device-side data model of the MANIAC energy engine (sm_100a).
//
// Layout in HBM (see DESIGN.md "Data layout"):
//   * static, shared by all walkers: host-framework atoms as SoA-of-double4 {x,y,z,q} +
//     type id, the per-type-pair LJ table (A = 4 eps sigma^12, B = 4 eps sigma^6), the
//     piecewise-polynomial table of erfc(alpha r)/r in r^2, the k-vector list
//     {kx,ky,kz} + ffW(k), S_host(k);
//   * per walker: guest coordinates as SoA with the molecule index fastest
//     (com[3][cap], offset[natom][3][cap], hcache[2][cap] per guest residue type; hcache =
//     the molecule's LJ / erfc-Coulomb sum against the static framework, refreshed on every
//     commit, so the OLD geometry of a move or deletion needs no framework pass), two S(k) buffers
//     (committed / trial, flipped on accept), running energies, counts, RNG state,
//     counters and accumulators, one pending-trial record.
#pragma once
#include <stdint.h>
#include "../../include/maniac_gpu.h"

#define MGPU_ERR_TOL 1.0e-10      // "error", src/parameters.f90:57
#define MGPU_NB_MAX_MOLECULE 5000 // src/parameters.f90:10

// Real-space Coulomb kernel g(s) = erfc(alpha sqrt(s)) / sqrt(s), s = r^2, tabulated as one
// degree-6 polynomial per interval; intervals are the top MGPU_TAB_K mantissa bits of s inside
// each binary octave, so the interval index is two integer instructions on the high word of s
// and no sqrt / rsqrt / erfc is evaluated per pair.  A table row is 48 bytes:
//   { c0, c1 | c2, c3 | c4, (float c5, float c6) }      (three 16-byte chunks)
// c5, c6 only need single precision (their terms are < 2^-30 of g).  In shared memory the
// table is replicated 8 times, replica r occupying 16-byte bank group r, and lane l reads
// replica l & 7: each LDS.128 of a quarter-warp touches 8 distinct bank groups, i.e. the
// gather is bank-conflict free whatever intervals the lanes need.  Error of the fit relative
// to the pair's Coulomb scale 1/r is < 4e-15 (tests/test_host_logic.py::
// test_coulomb_table_accuracy); pairs outside the tabulated range take the exact erfc path,
// and g = 0 beyond alpha r > MGPU_TAB_XCUT where erfc(x)/r < 1e-24.
#define MGPU_TRI_MAXREL 14
#define MGPU_TAB_K 5
#define MGPU_TAB_ROW 6            // doubles per row
#define MGPU_TAB_REP 8            // shared-memory replicas
#define MGPU_TAB_MAXOCT 14
#define MGPU_TAB_XCUT 7.0
// first tabulated distance: closer pairs take the exact erfc path (a divergent call).  1 A, not more: starting at 2 A
// would save 24.6 KB of shared memory, but M...H contacts of hydrogen-bonded waters (1.8 A) then hit the exact path
// often enough to cost 10 % (measured: 20.7 -> 18.6 M moves/s).
#ifndef MGPU_TAB_RLO
#define MGPU_TAB_RLO 1.0
#endif

struct MgpuTrial {
    int32_t active;               // 1 while a trial is pending on the walker
    int32_t kind, res, mol;
    int32_t res2, pad_;           // res2: the new residue type of a swap trial
    double  com[3];
    double  off[MGPU_MAX_SITES][3];
    double  e_old[6], e_new[6];
    double  hc_new[2];            // framework part {lj, coulomb (e^2/A)} of the trial geometry (-> per-molecule cache on commit)
};

// Constant-memory image of everything the kernels need.  Pointers are device pointers.
struct DevSys {
    // box (type_cell, src/simulation_state.f90:103-112)
    double H[9], Hinv[9], lo[3], L[3], invL[3], volume;
    int32_t triclinic;
    // triclinic minimum image without the 27-image loop (see min_image_r2<true>): lattice vectors
    // C m (tri_rel) that can beat the fractionally rounded image, and their integer coefficients m
    // (tri_m); tri_nrel < 0 = too many of them, keep the reference's 27-image search.
    int32_t tri_nrel;
    double tri_safe2;             // a listed vector can only shorten t when |t|^2 > tri_safe2 = min |C m|^2 / 4
    double tri_rel[MGPU_TRI_MAXREL][3], tri_m[MGPU_TRI_MAXREL][3], tri_len2[MGPU_TRI_MAXREL];   // one of every +-m pair
    // framework passes of triclinic cells work on FRACTIONAL coordinates (min_image_frac): wrapping is a rounding of the
    // coordinate difference itself, and the projection back costs 6 FMAs when the cell matrix is lower triangular
    int32_t tri_lower;            // matrix(0,1) = matrix(0,2) = matrix(1,2) = 0 (LAMMPS-style cell in the reference's column convention)
    int32_t tri_thr_hi[3];        // the listed vectors are tried when hi(|f_d|) >= tri_thr_hi[d] for some axis d (|f_d| within tri_eps of 1/2)
    double tri_eps[3];
    uint32_t tri_lut[16];         // listed vectors to try, by the set of faces the lanes of a warp are near (mgpu_init)
    const double2 *host_fxy, *host_fzq;   // framework atoms {f0, f1}, {f2, q}: fractional coordinates wrapped into [0, 1)  (triclinic cells only)
    // ewald / constants
    double rc, rc2, alpha, eps0_inv_real, twopi, beta, overlap;
    int32_t kmax[3], kmax_max, nk;
    // Coulomb table
    int32_t tab_ibase;            // (1023 + emin) << MGPU_TAB_K
    int32_t tab_nint;             // number of intervals (octaves << MGPU_TAB_K); row tab_nint is all zeros
    int32_t tab_hi_lo;            // high word of the table's first s: pairs below it (r < 1 A) take the exact path
    int32_t n_host_charged;       // framework atoms with |q| >= 1e-10 (work counters)
    const double *ctab;           // [tab_nint][MGPU_TAB_ROW]
    double s_zero;                // g == 0 for s >= s_zero (alpha r > MGPU_TAB_XCUT)
    // residues
    int32_t nres, ntypes, nactive;
    int32_t natom[MGPU_MAX_RES], active[MGPU_MAX_RES], cap[MGPU_MAX_RES];
    int32_t active_list[MGPU_MAX_RES];
    int32_t host_count[MGPU_MAX_RES];         // molecule count of inactive residues (static)
    int64_t goff[MGPU_MAX_RES];               // offset (doubles) of the residue block inside a walker's coordinates
    int32_t use_hcache;                       // 1: old-geometry framework sums come from the per-molecule cache
    double  charge[MGPU_MAX_RES][MGPU_MAX_SITES];
    int32_t type[MGPU_MAX_RES][MGPU_MAX_SITES];
    int32_t nq[MGPU_MAX_RES];                 // atoms of the residue type whose charge is not exactly 0 ...
    int8_t  qlist[MGPU_MAX_RES][MGPU_MAX_SITES];   // ... and their indices (k-space entries of a trial, fill_trial_tables)
    double  e_self[MGPU_MAX_RES];             // ewald_self_energy_single_mol(res)
    double  lambda[MGPU_MAX_RES];             // res%lambda
    double  self_host_total;                  // sum over inactive residues of e_self*count
    double  hh_lj, hh_coul;                   // static host-host pair energy
    // MC inputs
    double p_trans, p_rot, p_swap, p_insdel, p_widom, tstep, rstep;
    // static arrays
    const double4 *host_xyzq; const int32_t *host_type; int32_t n_host;   // q stored as 0 when |q| < 1e-10
    const double *host_qraw;                  // unthresholded charges (S_host build)
    const double2 *host_xy, *host_zq;         // the same atoms as two coalesced 16-byte streams {x,y}, {z,q}
    // guest atoms of each residue type by what they feel from the framework: 0 nothing, 1 LJ, 2 Coulomb, 3 both
    int32_t hl_n[MGPU_MAX_RES][4];
    int8_t  hl_list[MGPU_MAX_RES][4][MGPU_MAX_SITES];
    int8_t  iota[MGPU_MAX_SITES];             // 0, 1, 2, ...: "every atom of the probe" as a list (merged passes of triclinic cells)
    // the same classification of a probe residue's atoms against atom b of guest residue g:
    // gl_n[((ri*MAX_RES + g)*MAX_SITES + b)*4 + mode], gl_list[that index][MAX_SITES]   (global memory, warp-uniform reads)
    const int8_t *gl_n, *gl_list;
    const double *ljA, *ljB;                  // [ntypes*ntypes] 4 eps sigma^12, 4 eps sigma^6
    const int32_t *kx, *ky, *kz; const double *ffW;
    const double *S_host;                     // [2][nk] re, im
    // per-walker arrays
    int32_t n_walkers;
    int64_t coord_stride;                     // doubles per walker
    double  *coords;
    int32_t *count;                           // [W][MGPU_MAX_RES]
    double  *S;                               // [W][2 buffers][2][nk]
    int32_t *cur;                             // [W] index of the committed S buffer
    double  *energy;                          // [W][6]
    double  *mu;                              // [W][MGPU_MAX_RES]
    uint64_t *rng;                            // [W][4]
    double  *rmax2;                           // [MGPU_MAX_RES] upper bound of |offset|^2 over every molecule of the residue type, all walkers
    double  *step;                            // [W][2] translation_step, rotation_step_angle of the walker (adjust_move_step_sizes)
    long long *counters;                      // [W][12]
    double  *widom_w; long long *widom_n;     // [W][MGPU_MAX_RES]
    double  *avg;                             // [W][MGPU_MAX_RES][4]
    MgpuTrial *trial;                         // [W]
    unsigned long long *pair_count;           // [3] pairs evaluated, LJ terms, Coulomb terms (all walkers)
    // (appended, so that the offsets of everything above -- and with them the constant-bank loads the orthorhombic kernels
    // were tuned with -- stay what they were)
    int32_t tri_thr_min, tri_pad_;            // min of tri_thr_hi: "near some face that matters" in one compare (min_image_frac_fast)
};
