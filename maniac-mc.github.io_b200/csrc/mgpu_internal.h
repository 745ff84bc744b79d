// mgpu_internal.h -- device-side data model of the MANIAC energy engine (sm_100a).
//
// Layout in HBM (see DESIGN.md "Data layout"):
//   * static, shared by all walkers: host-framework atoms as SoA-of-double4 {x,y,z,q} +
//     type id, the per-type-pair LJ table, the k-vector list {kx,ky,kz} + ffW(k),
//     S_host(k);
//   * per walker: guest coordinates as SoA with the molecule index fastest
//     (com[3][cap], offset[natom][3][cap] per guest residue type), two S(k) buffers
//     (committed / trial, flipped on accept), running energies, counts, RNG state,
//     counters and accumulators, one pending-trial record.
#pragma once
#include <stdint.h>
#include "../../include/maniac_gpu.h"

#define MGPU_ERR_TOL 1.0e-10      // "error", src/parameters.f90:57
#define MGPU_NB_MAX_MOLECULE 5000 // src/parameters.f90:10

struct MgpuTrial {
    int32_t active;               // 1 while a trial is pending on the walker
    int32_t kind, res, mol;
    double  com[3];
    double  off[MGPU_MAX_SITES][3];
    double  e_old[6], e_new[6];
};

// Constant-memory image of everything the kernels need.  Pointers are device pointers.
struct DevSys {
    // box (type_cell, src/simulation_state.f90:103-112)
    double H[9], Hinv[9], lo[3], L[3], invL[3], volume;
    int32_t triclinic;
    // ewald / constants
    double rc, alpha, eps0_inv_real, twopi, beta, overlap;
    int32_t kmax[3], kmax_max, nk;
    // residues
    int32_t nres, ntypes, nactive;
    int32_t natom[MGPU_MAX_RES], active[MGPU_MAX_RES], cap[MGPU_MAX_RES];
    int32_t active_list[MGPU_MAX_RES];
    int32_t host_count[MGPU_MAX_RES];         // molecule count of inactive residues (static)
    int64_t goff[MGPU_MAX_RES];               // offset (doubles) of the residue block inside a walker's coordinates
    double  charge[MGPU_MAX_RES][MGPU_MAX_SITES];
    int32_t type[MGPU_MAX_RES][MGPU_MAX_SITES];
    double  e_self[MGPU_MAX_RES];             // ewald_self_energy_single_mol(res)
    double  lambda[MGPU_MAX_RES];             // res%lambda
    double  self_host_total;                  // sum over inactive residues of e_self*count
    double  hh_lj, hh_coul;                   // static host-host pair energy
    // MC inputs
    double p_trans, p_rot, p_swap, p_insdel, p_widom, tstep, rstep;
    // static arrays
    const double4 *host_xyzq; const int32_t *host_type; int32_t n_host;
    const double *eps, *sig;                  // [ntypes*ntypes]
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
    long long *counters;                      // [W][12]
    double  *widom_w; long long *widom_n;     // [W][MGPU_MAX_RES]
    double  *avg;                             // [W][MGPU_MAX_RES][4]
    MgpuTrial *trial;                         // [W]
    unsigned long long *pair_count;           // [3] pairs evaluated, LJ terms, Coulomb terms (all walkers)
};
