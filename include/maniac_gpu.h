/*
 * maniac_gpu.h -- C ABI of the B200 (sm_100a) energy engine for MANIAC-MC.
 *
 * This is the drop-in boundary between MANIAC's Fortran move drivers and the
 * CUDA implementation of the per-trial-move energy change.  The reference has
 * no FFI of its own: the seam is the set of energy routines that
 * src/monte_carlo_utils.f90 (compute_old_energy :367-423, compute_new_energy
 * :300-361, accept/reject :429-505, remove_molecule :642-657) and
 * src/energy_utils.f90 (update_system_energy :22-39) call.  Each export below
 * names the reference routine(s) it replaces; INTEGRATION.md shows the
 * ISO_C_BINDING interface block and the Makefile lines a maintainer adds.
 *
 * Conventions
 *   - plain C types only; every array is caller-owned and copied during the
 *     call (the library keeps no host pointer after returning);
 *   - indices are 0-based here; the Fortran shim subtracts 1;
 *   - 3x3 matrices are passed as 9 doubles, element (i,j) at [i*3+j] with the
 *     reference's own (i,j) meaning, i.e. box%cell%matrix(i+1,j+1);
 *   - a molecule's geometry is com[3] + offset[natom][3] (offset(:,atom) is
 *     contiguous, which is what the Fortran section guest%offset(:,res,mol,1:n)
 *     becomes as a contiguous temporary);
 *   - every function returns 0 on success, non-zero on error, with the text
 *     available from mgpu_last_error(); the Fortran wrapper turns non-zero
 *     into `call abort_run(msg, code)` (src/output_utils.f90:581-605);
 *   - one global context per process, synchronous, not re-entrant (the
 *     reference is single-threaded with module-level state); multi-GPU = one
 *     process per device;
 *   - there is NO CPU fallback: without a CUDA device mgpu_init fails.
 *
 * State model: the device always holds the last COMMITTED configuration of
 * every walker.  A trial (mgpu_new_energy / mgpu_trial_batch) never mutates
 * it; mgpu_commit applies the pending trial (coordinates, S(k), counts,
 * running energies), mgpu_rollback drops it.  This replaces the reference's
 * "mutate Ak in place, copy Ak_old back on reject" (src/ewald_phase.f90:17-125).
 */
#ifndef MANIAC_GPU_H
#define MANIAC_GPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGPU_MAX_RES    8    /* residue types                                      */
#define MGPU_MAX_SITES  16   /* atoms per guest molecule                           */

/* energy_type, src/simulation_state.f90:61-69, same field order */
enum { MGPU_E_NON_COULOMB = 0, MGPU_E_COULOMB = 1, MGPU_E_RECIP = 2,
       MGPU_E_SELF = 3, MGPU_E_INTRA = 4, MGPU_E_TOTAL = 5 };

/* is_creation / is_deletion flags of compute_old/new_energy */
enum { MGPU_KIND_MOVE = 0, MGPU_KIND_CREATE = 1, MGPU_KIND_DELETE = 2,
       /* identity swap (src/swapping.f90:34-105): deletion-style old energy of (res, mol) +
        * creation-style new energy of a molecule of another type at the same CoM.  In
        * mgpu_trial_batch the new type rides in the upper bits: kind = MGPU_KIND_SWAP | (res_new << 8) */
       MGPU_KIND_SWAP = 3 };

/* move codes reported by mgpu_sweep traces (same numbering as the oracle) */
enum { MGPU_MV_NONE = 0, MGPU_MV_TRANSLATE = 1, MGPU_MV_ROTATE = 2, MGPU_MV_CREATE = 3,
       MGPU_MV_DELETE = 4, MGPU_MV_SWAP = 5, MGPU_MV_WIDOM = 6 };

/* One residue type: res%atom, res%role / thermo%is_active, primary%atoms%charges,
 * primary%atoms%types (src/simulation_state.f90:134-141,215-231). */
typedef struct mgpu_residue {
    int32_t natom;            /* res%atom(res)                                    */
    int32_t is_active;        /* thermo%is_active(res): 1 guest, 0 host           */
    int32_t nmol;             /* primary%num%residues(res) at init                */
    int32_t capacity;         /* max molecules per walker (<= NB_MAX_MOLECULE)    */
    const double  *charges;   /* [natom]                                          */
    const int32_t *types;     /* [natom] atom type id, 0-based                    */
    const double  *com;       /* [nmol][3]  coord%com(:,res,mol)                  */
    const double  *offset;    /* [nmol][natom][3] coord%offset(:,res,mol,atom)    */
    double mass;              /* res%mass(res) (g/mol), for the de Broglie length */
    double fugacity;          /* thermo%fugacity(res); < 0 = use chemical_potential */
    double chemical_potential;/* thermo%chemical_potential(res) (kcal/mol)        */
} mgpu_residue;

/* Everything setup_simulation_parameters (src/prepare_utils.f90:20-43) has
 * produced by the time the MC loop starts. */
typedef struct mgpu_system {
    double matrix[9];         /* primary%cell%matrix                              */
    double lo[3];             /* primary%cell%bounds(:,1)                         */
    int32_t nres;
    const mgpu_residue *residues;
    int32_t ntypes;           /* atom types                                       */
    const double *epsilon;    /* [ntypes*ntypes] per type pair, AFTER Lorentz-Berthelot */
    const double *sigma;      /* (coeff%epsilon/sigma(res_i,res_j,atom_i,atom_j) collapsed on types, A16) */
    double temperature;       /* thermo%temperature                               */
    double ewald_tolerance;   /* ewald%param%tolerance (input)                    */
    double real_space_cutoff; /* mc_input%real_space_cutoff (input)               */
    double translation_step;  /* mc_input%translation_step                        */
    double rotation_step_angle;
    double p_translation, p_rotation, p_swap, p_insertion_deletion, p_widom; /* proba% */
    int32_t n_walkers;        /* independent copies of the system on this device  */
    int32_t device;           /* CUDA device ordinal                              */
} mgpu_system;

/* ---- lifecycle ---------------------------------------------------------------------
 * mgpu_init replaces setup_ewald + allocate_array + precompute_valid_reciprocal_vectors
 * + prepare_monte_carlo (src/prepare_utils.f90:20-43,141-259; src/ewald_kvectors.f90:24-65)
 * and uploads the SoA image of host/guest coordinates, charges and the LJ table.
 * Every walker starts as a copy of the configuration in `sys`. */
int  mgpu_init(const mgpu_system *sys);
void mgpu_finalize(void);
const char *mgpu_last_error(void);
int  mgpu_device_info(char *name, int name_len, int *sm_count, double *mem_gb);

/* Ewald / box parameters the engine derived (to cross-check against the host's own) */
int mgpu_get_ewald(double *alpha, int32_t kmax[3], int32_t *nkvec, double *rc_used);
int mgpu_get_kvectors(int32_t *kx, int32_t *ky, int32_t *kz, double *k_squared_mag, double *form_factor_times_weight);
int mgpu_get_box(double matrix[9], double reciprocal[9], double *volume, int32_t *is_triclinic);
/* launch shape of the warp-per-walker kernels: walkers (warps) per CTA, dynamic shared memory per CTA, SMs */
int mgpu_get_launch_info(int32_t *walkers_per_cta, int64_t *smem_bytes_per_cta, int32_t *sm_count);
/* triclinic cells: number of extra lattice vectors min_image examines per pair after rounding the
 * fractional coordinates (0 for orthorhombic cells, and for large triclinic cells where a candidate could only
 * matter beyond the LJ cutoff at distances whose erfc-Coulomb terms sum to < 1e-12 kcal/mol per trial -- bound
 * evaluated in mgpu_init); -1 = the literal 27-image search of src/geometry_utils.f90:263-280 is used for every
 * pair (very skewed cell) */
int mgpu_get_triclinic_candidates(int32_t *n);
/* launch shape mgpu_sweep / mgpu_block would use for n_walkers walkers in flight: threads per walker (32, 64 or 128) and
 * walkers per CTA (one CTA per SM); see MGPU_OPT_SWEEP_TEAM */
int mgpu_get_sweep_shape(int32_t n_walkers, int32_t *threads_per_walker, int32_t *walkers_per_cta);
/* The two pieces of host-side planning behind the above, as pure functions that need no device (unit tests; neither has a
 * counterpart in the reference):
 *  - the launch shape for n_walkers on a GPU with sm_count SMs and warps_per_cta warps per sweep CTA (forced = the
 *    MGPU_OPT_SWEEP_TEAM value);
 *  - the triclinic minimum-image plan for a cell matrix (row major, cell vectors = COLUMNS as in geometry_utils.f90:263-280):
 *    the lattice vectors C m that can beat the fractionally rounded image (n_vectors of them, -1 = too skewed, the literal
 *    27-image search is kept; vectors[k][3] = C m, coefficients[k][3] = m, one of every +-m pair), faces[k] = the faces of the
 *    fractional cube (bit d = axis d) a rounded vector has to be near for vector k to matter, thr_hi[3] = high word of the
 *    |f_d| from which face d counts as near, lut[16] = vectors to try by the set of near faces (bit 3 = far outside the
 *    cell, bit 31 of an entry = complete search).  Arrays hold up to 14 vectors. */
int mgpu_plan_sweep_shape(int32_t n_walkers, int32_t sm_count, int32_t warps_per_cta, int32_t forced,
                          int32_t *threads_per_walker, int32_t *walkers_per_cta);
int mgpu_plan_triclinic(const double *matrix, int32_t *n_vectors, double *vectors, int32_t *coefficients, int32_t *faces,
                        int32_t *thr_hi, uint32_t *lut);
/*  - the order in which mgpu_init stores the static framework atoms (xyz[n_atoms][3] Cartesian, lo = cell origin): order[k] =
 *    index of the atom stored k-th; columns[3] = bins per axis when the cell's plan leaves an axis free (1 = not binned). */
int mgpu_plan_framework_order(const double *matrix, const double *lo, int32_t n_atoms, const double *xyz, int32_t *order, int32_t *columns);
int mgpu_get_thermo(int32_t res, double *beta, double *lambda, double *mu_walker0);

/* ---- per-walker state --------------------------------------------------------------- */
/* coordinates of one molecule (mirror of guest%com / guest%offset); invalidates S(k) and
 * the running energies of that walker until mgpu_total_energy is called */
int mgpu_set_molecule(int32_t walker, int32_t res, int32_t mol, const double com[3], const double *offset);
int mgpu_get_molecule(int32_t walker, int32_t res, int32_t mol, double com[3], double *offset);
/* update_counts, src/monte_carlo_utils.f90:662-672 (absolute value) */
int mgpu_set_count(int32_t walker, int32_t res, int32_t n);
int mgpu_get_count(int32_t walker, int32_t res, int32_t *n);
/* thermo%chemical_potential(res) / fugacity of one walker (isotherm points) */
int mgpu_set_chemical_potential(int32_t walker, int32_t res, double mu);
int mgpu_set_fugacity(int32_t walker, int32_t res, double fugacity);
/* the same for a run of walkers in one copy (mu[n_walkers], kcal/mol), and their molecule counts of one type */
int mgpu_set_chemical_potentials(int32_t first_walker, int32_t n_walkers, int32_t res, const double *mu);
int mgpu_get_counts(int32_t first_walker, int32_t n_walkers, int32_t res, int32_t *counts);
/* S(k) of a walker (ewald%Ak), interleaved re,im */
int mgpu_get_Ak(int32_t walker, double *re_im);
/* running totals `energy` of a walker */
int mgpu_get_energy(int32_t walker, double out[6]);

/* Engine options.  MGPU_OPT_HOST_CACHE (default 1): the framework is static, so a guest
 * molecule's LJ + erfc-Coulomb sum against it only changes when that molecule moves; the
 * engine keeps it per molecule (written at every commit and every full recompute) and the
 * OLD geometry of a move / deletion reads it instead of sweeping the framework again.
 * 0 = recompute it every trial, like pairwise_energy_for_molecule on the old geometry
 * (src/monte_carlo_utils.f90:367-423) does.  Same energies to rounding either way. */
enum { MGPU_OPT_HOST_CACHE = 1,
       /* MGPU_OPT_PHASE_SYNC (default 1 = the measured best set for the launch shape): in mgpu_sweep the walkers of a
        * CTA meet at named barriers inside every MC step, so they run the same loop at the same time and share its
        * instructions and the framework atoms it streams in the caches.  0 = free-running; other values select the
        * barriers bit by bit (1 top of the step, 2 before the energy evaluation, 4 before the guest pass, 8 before
        * k-space, 16 = a CTA whose walkers are nearly empty -- fewer than 4 molecules each -- runs free).
        * 1 means 1|4|16 when one warp owns a walker and 1|4|8 for the team shapes.  Results are identical either way
        * (walkers never exchange data). */
       MGPU_OPT_PHASE_SYNC = 2,
       /* MGPU_OPT_BLOCK_SLICES (default 8, at most 16): mgpu_block cuts its walkers into this many slices, each on its
        * own stream, so that the records of slice i+1 travel host -> device and those of slice i-1 device -> host while
        * slice i is being swept.  1 = no pipelining (load, sweep, save one after the other).  Same results either way. */
       MGPU_OPT_BLOCK_SLICES = 3,
       /* MGPU_OPT_SWEEP_TEAM (default -1 = automatic): shape of mgpu_sweep / mgpu_block launches.  0 = one warp per
        * walker (16 walkers per CTA: the throughput shape, fills the GPU from 2368 walkers up).  1 = a team of four
        * warps per walker (4 walkers per CTA), 2 = a team of two warps (8 per CTA): the energy loops of a trial are
        * split over 128 / 64 threads, for launches with few walkers per GPU (a fixed isotherm spread over more GPUs).
        * -1 picks the widest team with which one wave of CTAs still holds every walker, and the fewest walkers per CTA
        * that need no extra round of CTAs (mgpu_get_sweep_shape reports the choice).  Same trajectories either way
        * (sums are reduced in another order). */
       MGPU_OPT_SWEEP_TEAM = 4 };
int mgpu_set_option(int32_t option, int32_t value);

/* ---- energy routines (single walker, drop-in) -------------------------------------- */
/* update_system_energy, src/energy_utils.f90:22-39: full recompute of the 5 components,
 * rebuilds S(k) (compute_total_reciprocal_energy, src/ewald_energy.f90:20-58) */
int mgpu_total_energy(int32_t walker, double out[6]);
/* pairwise_energy_for_molecule, src/pairwise_energy_utils.f90:21-90, for a guest molecule.
 * com/offset may be NULL = use the committed coordinates. */
int mgpu_pairwise_energy_for_molecule(int32_t walker, int32_t res, int32_t mol, int32_t skip_ordering_check,
                                      const double *com, const double *offset,
                                      double *e_non_coulomb, double *e_coulomb);
/* ewald_self_energy_single_mol :177-205 and intra_res_real_coulomb_energy :212-252 of
 * src/ewald_energy.f90 */
int mgpu_ewald_self_energy_single_mol(int32_t res, double *e);
int mgpu_intra_res_real_coulomb_energy(int32_t walker, int32_t res, int32_t mol,
                                       const double *com, const double *offset, double *e);
/* reciprocal_ewald_energy :139-164 on the committed S(k) */
int mgpu_reciprocal_ewald_energy(int32_t walker, double *e);
/* compute_old_energy, src/monte_carlo_utils.f90:367-423 (committed coordinates) */
int mgpu_old_energy(int32_t walker, int32_t res, int32_t mol, int32_t kind, double out[6]);
/* compute_new_energy :300-361.  MOVE / CREATE: (com, offset) is the trial geometry the
 * host proposed; DELETE: ignored (may be NULL).  Leaves a pending trial on the walker. */
int mgpu_new_energy(int32_t walker, int32_t res, int32_t mol, int32_t kind,
                    const double *com, const double *offset, double out[6]);
/* attempt_swap_move, src/swapping.f90:59-88: compute_old_energy(res_old, mol_old, is_deletion) and,
 * with that molecule removed, compute_new_energy(res_new, N(res_new)+1, is_creation) for the geometry
 * (com, offset[natom(res_new)][3]) the host built with insert_and_orient_molecule.  Like the
 * reference, S(k) of the trial keeps the swapped-out molecule's contribution.  Leaves a pending
 * trial: mgpu_commit = accept_swap_move (:143-155), mgpu_rollback = reject_swap_move (:113-137). */
int mgpu_swap_energy(int32_t walker, int32_t res_old, int32_t mol_old, int32_t res_new,
                     const double *com, const double *offset, double e_old[6], double e_new[6]);
/* accept_molecule_move :429-442 / accept_creation_move (creation.f90:82-116) /
 * accept_deletion_move (deletion.f90:83-122) incl. remove_molecule + update_counts */
int mgpu_commit(int32_t walker);
/* reject_molecule_move :480-505 / reject_creation_move :579-591 / reject_deletion_move */
int mgpu_rollback(int32_t walker);

/* ---- batched host-driven trials (many walkers per call) ----------------------------- */
/* One trial per listed walker: old and new energies in ONE launch.  com[n][3],
 * offset[n][MGPU_MAX_SITES][3] (only the first natom rows are read).  e_old/e_new [n][6]. */
int mgpu_trial_batch(int32_t n, const int32_t *walker, const int32_t *res, const int32_t *mol,
                     const int32_t *kind, const double *com, const double *offset,
                     double *e_old, double *e_new);
int mgpu_commit_batch(int32_t n, const int32_t *walker, const int32_t *accept);

/* ---- device-resident Monte Carlo (new capability; drivers of src/translation.f90,
 * rotation.f90, creation.f90, deletion.f90, widom.f90 and the loop body of
 * src/monte_carlo.f90:50-99 run on the GPU, one warp per walker; incl. swapping.f90 when p_swap > 0) ---- */
typedef struct mgpu_step_trace {
    int32_t move, res, mol, accepted;
    double  dE, prob;
    double  e_old[6], e_new[6];
} mgpu_step_trace;

/* per-walker generator: xoshiro256** seeded by splitmix64(seed + 104729*walker) */
int mgpu_seed(uint64_t seed);
int mgpu_get_rng_state(int32_t walker, uint64_t st[4]);
/* n_steps MC steps for every walker in [first_walker, first_walker+n_walkers).
 * trace (optional) receives the steps of walker `trace_walker` only ([n_steps]). */
int mgpu_sweep(int32_t first_walker, int32_t n_walkers, int64_t n_steps,
               int32_t trace_walker, mgpu_step_trace *trace);
/* ---- block-level entry: HOST state in, HOST state out -------------------------------------
 * One block of monte_carlo_loop (src/monte_carlo.f90:40-118: nb_step MC steps, then the per-block
 * outputs energy.dat / number_*.dat / moves.dat / trajectory, src/write_utils.f90:16-34) for the
 * walkers [first, first+n), from and to flat HOST arrays.  A walker's record (doubles) is
 *   [0] record length | [1..8] primary%num%residues | [9..14] energy (energy_type order) |
 *   [15..18] RNG state (raw bits) | [19..30] counters%... (trials, successes) x 6 | [31] 0 |
 *   [32..63] block averages [res][sum N, sum N^2, sum E, samples] | [64] translation_step [65] rotation_step_angle
 *   [66..71] 0 | [72..72+2 nk) ewald%Ak (re[nk], im[nk]) |
 *   then for every active residue type, for mol = 1..max(count,1): guest%com(:,res,mol), guest%offset(:,res,mol,1:natom),
 *   and the molecule's cached framework energy {lj, coulomb (e^2/A)}.  (Slot 1 is stored even for an empty type: it is
 *   the geometry template of the next insertion, src/monte_carlo_utils.f90:563-564.)
 * Record i starts at blob[offsets[i]]; offsets[n] is the total length.  mgpu_save_walkers fills
 * blob and offsets (capacity in doubles; mgpu_record_doubles_max() bounds one record);
 * mgpu_load_walkers restores walkers from records (no recompute: the record is the full state);
 * mgpu_block = load (if blob_in) + mgpu_sweep(n_steps) + save (if blob_out).  Buffers from
 * mgpu_host_alloc are page-locked, so the copies run at full PCIe rate; any host pointer works. */
void *mgpu_host_alloc(size_t bytes);
void  mgpu_host_free(void *p);
int64_t mgpu_record_doubles_max(void);
int mgpu_save_walkers(int32_t first_walker, int32_t n_walkers, double *blob, int64_t capacity, int64_t *offsets);
int mgpu_load_walkers(int32_t first_walker, int32_t n_walkers, const double *blob, const int64_t *offsets);
int mgpu_block(int32_t first_walker, int32_t n_walkers, int64_t n_steps,
               const double *blob_in, const int64_t *offsets_in,
               double *blob_out, int64_t capacity_out, int64_t *offsets_out);
/* bytes moved host->device / device->host by the three calls above since the last reset */
int mgpu_get_traffic(int64_t *h2d_bytes, int64_t *d2h_bytes, int32_t reset);
/* adjust_move_step_sizes, src/monte_carlo_utils.f90:98-134: Robbins-Monro update of mc_input%translation_step /
 * rotation_step_angle from the cumulative counters, called at the end of a block when recalibrate_moves is set.
 * Every walker carries its own pair of step sizes (each starts from the input values). */
int mgpu_adjust_move_step_sizes(int32_t first_walker, int32_t n_walkers);
int mgpu_get_step_sizes(int32_t walker, double out[2]);
int mgpu_set_step_sizes(int32_t walker, double translation_step, double rotation_step_angle);
/* counters%{translations,rotations,creations,deletions,swaps,widom}(1:2) of a walker */
int mgpu_get_counters(int32_t walker, int64_t out[12]);
/* statistic%weight / statistic%sample (src/widom.f90:74-92) accumulated by mgpu_sweep */
int mgpu_get_widom(int32_t walker, int32_t res, double *sum_weight, int64_t *samples);
/* block averages accumulated by mgpu_sweep: sum N, sum N^2, sum E_total, samples */
int mgpu_get_averages(int32_t walker, int32_t res, double out[4]);
int mgpu_reset_averages(void);

/* Widom test-particle batch: n independent insertions of residue `res` into the
 * committed configuration of `walker` (body of widom_trial, src/widom.f90:30-68, with
 * insertion id -> 5 uniforms from a counter-based stream; state is never changed).
 * dE_out may be NULL.  sum_w = sum of exp(-beta dU) over insertions with weight > 1e-10,
 * n_ok = their number (accumulate_widom_weight :74-92). */
int mgpu_widom_batch(int32_t walker, int32_t res, int64_t first_id, int64_t n, uint64_t seed,
                     double *dE_out, double *sum_w, int64_t *n_ok);

/* ---- multi-GPU: shard walkers over ranks, reduce averages at the end ---------------- */
/* NCCL is bootstrapped by the caller: rank 0 obtains an id, every rank receives it
 * (torch.distributed / MPI / a file) and calls mgpu_nccl_init.  mgpu_reduce_averages
 * is an in-place sum over ranks of `count` doubles held in a HOST buffer (per
 * isotherm point: sum N, sum N^2, sum E, samples, sum w, n_w).  nranks == 1: no-op. */
int mgpu_nccl_unique_id(char id[128]);
int mgpu_nccl_init(const char id[128], int32_t nranks, int32_t rank);
int mgpu_reduce_averages(double *buf, int32_t count);
void mgpu_nccl_finalize(void);
const char *mgpu_nccl_last_error(void);

/* ---- measurement helpers (bench.py) ------------------------------------------------- */
/* average device time (ms) per launch of the named kernel class since the last reset,
 * measured with CUDA events on the engine's stream; and the number of launches */
int mgpu_timing_reset(void);
int mgpu_timing_get(const char *kernel, double *total_ms, int64_t *launches);
/* work counters since the last reset, summed over all walkers: atom pairs evaluated,
 * LJ terms inside the cutoff, erfc-Coulomb terms (the algorithmic-work figures of the
 * roofline) */
int mgpu_get_pair_counts(int64_t out[3]);
/* atom pairs with neither LJ nor Coulomb interaction (only the r < 1e-10 overlap sentinel could make them count) that the
 * per-molecule screen settled without evaluating them; the reference visits them (they belong to its operation count) */
int mgpu_get_screened_pairs(int64_t *n);
int mgpu_reset_pair_counts(void);
/* achieved FP64 FMA throughput of a register-resident DFMA loop (TFLOP/s) and the SM
 * clock it ran at: the roofline denominator for the FP64-bound kernels */
int mgpu_measure_fp64_peak(double *tflops, double *seconds);
/* achieved read bandwidth (GB/s) of a 48 MB buffer that stays resident in L2, all SMs streaming it with 16-byte loads:
 * the roofline denominator of the k-space traffic (S(k) of the resident walkers lives in L2) */
int mgpu_measure_l2_peak(double *gbytes_per_s);
/* accuracy of the two fast per-pair primitives on the device, measured against the exact
 * forms over the whole tabulated range: max relative error of 1/r^2 (MUFU seed + Newton) and
 * max error of the erfc(alpha r)/r table relative to the pair's Coulomb scale */
int mgpu_selftest_math(double *max_rel_rcp, double *max_err_table);
/* host-only (no GPU needed): builds the erfc(alpha r)/r table exactly as mgpu_init does for
 * [r_lo, r_hi] and compares n log-spaced samples with long-double erfc.  Returns the number
 * of samples inside the tabulated range. */
int mgpu_coulomb_table_check(double alpha, double r_lo, double r_hi, int32_t n,
                             double *max_rel_err, double *max_abs_err_times_r);

#ifdef __cplusplus
}
#endif
#endif /* MANIAC_GPU_H */
