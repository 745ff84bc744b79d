/*
 * maniac_oracle.h -- CPU ORACLE for the MANIAC-MC per-trial-move energy path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * Fortran algorithm (file:line citations are relative to /root/reference/).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product path (libmaniac_gpu.so) never does.
 *
 * Parity status: the static energies are PINNED against the reference's own
 * known-answer values (LAMMPS logs / analytic values used by the reference's
 * integration tests, see tests/test_oracle_golden.py).  Per-move dE, accept /
 * reject sequences, uptake and Widom mu_ex are "parity unpinned": the reference
 * holds no fixture for them and no Fortran compiler exists in the build
 * container, so they are pinned oracle-vs-GPU only.
 *
 * Indices in this API are 0-based (Fortran ids minus one).
 */
#ifndef MANIAC_ORACLE_H
#define MANIAC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_system orc_system;

/* energy_type, src/simulation_state.f90:61-69 (same field order) */
enum { ORC_E_NON_COULOMB = 0, ORC_E_COULOMB = 1, ORC_E_RECIP = 2,
       ORC_E_SELF = 3, ORC_E_INTRA = 4, ORC_E_TOTAL = 5 };

/* move kinds for compute_old/new_energy (monte_carlo_utils.f90:300-423) */
enum { ORC_KIND_MOVE = 0, ORC_KIND_CREATE = 1, ORC_KIND_DELETE = 2 };

/* trace codes of one MC step */
enum { ORC_MV_NONE = 0, ORC_MV_TRANSLATE = 1, ORC_MV_ROTATE = 2, ORC_MV_CREATE = 3,
       ORC_MV_DELETE = 4, ORC_MV_SWAP = 5, ORC_MV_WIDOM = 6 };

typedef struct orc_step_trace {
    int32_t move;        /* ORC_MV_* (NONE when the driver returned early)     */
    int32_t res;         /* residue type picked                                  */
    int32_t mol;         /* molecule index picked (0-based, -1 = none)           */
    int32_t accepted;    /* 1 accepted, 0 rejected                               */
    double  dE;          /* new%total - old%total                                */
    double  prob;        /* acceptance probability                               */
    double  e_old[6];
    double  e_new[6];
} orc_step_trace;

/* ---- lifecycle ------------------------------------------------------------ */
orc_system *orc_create(void);
void        orc_destroy(orc_system *s);
const char *orc_last_error(const orc_system *s);

/* ---- constants (src/constants.f90:8-21, src/parameters.f90:37) ------------ */
double orc_const_PI(void);
double orc_const_TWOPI(void);
double orc_const_SQRTPI(void);
double orc_const_EPS0_INV_real(void);
double orc_const_KB_kcalmol(void);

/* ---- system definition ---------------------------------------------------- */
/* matrix[i*3+j] = box%cell%matrix(i+1,j+1) (rows = a,b,c as the reader fills it,
 * readers_utils.f90:256-258); lo = bounds(:,1).  Runs prepare_simulation_box. */
int orc_set_box(orc_system *s, const double matrix[9], const double lo[3]);
/* returns residue id, or -1 */
int orc_add_residue(orc_system *s, int natom, int is_active, const double *charges,
                    const int *types, double mass, int capacity);
/* LJ table semantics of parameters_parser.f90:19-178, at atom-type level */
int orc_lj_begin(orc_system *s, int ntypes);
int orc_lj_pair_coeff(orc_system *s, int ti, int tj, double eps, double sig);
int orc_lj_finalize(orc_system *s);   /* applies Lorentz-Berthelot fill */
int orc_lj_get(const orc_system *s, double *eps, double *sig); /* ntypes*ntypes each */

int orc_set_molecule(orc_system *s, int res, int mol, const double com[3], const double *offset);
int orc_get_molecule(const orc_system *s, int res, int mol, double com[3], double *offset);
int orc_set_count(orc_system *s, int res, int n);
int orc_get_count(const orc_system *s, int res);

/* setup_ewald + allocate_array + precompute_valid_reciprocal_vectors
 * (prepare_utils.f90:20-43,110-226; ewald_kvectors.f90:24-65) */
int orc_setup_ewald(orc_system *s, double tolerance, double real_space_cutoff);
int orc_get_ewald(const orc_system *s, double *alpha, int kmax[3], int *nk, double *rc);
int orc_get_kvectors(const orc_system *s, int *kx, int *ky, int *kz, double *k2, double *ff);
int orc_get_box(const orc_system *s, double matrix[9], double reciprocal[9], double *volume, int *shape);

/* prepare_monte_carlo (prepare_utils.f90:231-259): beta, mu, lambda */
int orc_set_thermo(orc_system *s, double temperature);
int orc_set_fugacity(orc_system *s, int res, double fugacity);        /* mu = ln(f)/beta */
int orc_set_chemical_potential(orc_system *s, int res, double mu);
int orc_set_mc_input(orc_system *s, double translation_step, double rotation_step_angle,
                     double p_translation, double p_rotation, double p_swap,
                     double p_insertion_deletion, double p_widom);
double orc_get_beta(const orc_system *s);
double orc_get_lambda(const orc_system *s, int res);
double orc_get_mu(const orc_system *s, int res);

/* ---- energy routines (the hot path) -------------------------------------- */
double orc_minimum_image_distance(const orc_system *s, int r1, int m1, int a1, int r2, int m2, int a2);
int orc_pairwise_energy_for_molecule(const orc_system *s, int res, int mol, int skip_ordering_check,
                                     double *e_non_coulomb, double *e_coulomb);
int orc_update_system_energy(orc_system *s, double out[6]);   /* energy_utils.f90:22-39 */
int orc_get_energy(const orc_system *s, double out[6]);
double orc_ewald_self_energy_single_mol(const orc_system *s, int res);
double orc_intra_res_real_coulomb_energy(const orc_system *s, int res, int mol);
double orc_reciprocal_ewald_energy(const orc_system *s);
int orc_get_Ak(const orc_system *s, double *re_im /* 2*nk */);
int orc_compute_ewald_phase_factors(orc_system *s, int res, int mol);
int orc_save_single_mol_fourier_terms(orc_system *s, int res, int mol);
int orc_restore_single_mol_fourier(orc_system *s, int res, int mol);
int orc_update_reciprocal_amplitude_single_mol(orc_system *s, int res, int mol, int kind);
int orc_compute_old_energy(orc_system *s, int res, int mol, int kind, double out[6]);
int orc_compute_new_energy(orc_system *s, int res, int mol, int kind, double out[6]);

/* geometry helpers pinned by the reference's unit tests */
void orc_apply_PBC(const orc_system *s, double pos[3]);
void orc_wrap_into_box(const orc_system *s, double pos[3]);

/* ---- RNG contract (replaces libgfortran random_number) ----- */
void   orc_seed_rng(orc_system *s, uint64_t seed);
double orc_rand_uniform(orc_system *s);
void   orc_get_rng_state(const orc_system *s, uint64_t st[4]);
void   orc_set_rng_state(orc_system *s, const uint64_t st[4]);
/* optional externally supplied uniform stream (consumed in draw order) */
int    orc_set_uniform_stream(orc_system *s, const double *u, int64_t n);
int64_t orc_uniform_stream_used(const orc_system *s);

/* ---- move drivers -------------------------------------------------------- */
int orc_attempt_translation_move(orc_system *s, int res, int mol, orc_step_trace *t);
int orc_attempt_rotation_move(orc_system *s, int res, int mol, orc_step_trace *t);
int orc_attempt_creation_move(orc_system *s, int res, int mol, orc_step_trace *t);
int orc_attempt_deletion_move(orc_system *s, int res, int mol, orc_step_trace *t);
int orc_attempt_swap_move(orc_system *s, int res, int mol, orc_step_trace *t);
int orc_widom_trial(orc_system *s, int res, int mol, orc_step_trace *t);
/* n iterations of the body of monte_carlo_loop (monte_carlo.f90:50-120), no
 * block bookkeeping / step-size recalibration / file output */
int orc_monte_carlo_steps(orc_system *s, int64_t nsteps, orc_step_trace *trace /* nsteps or NULL */);
int orc_adjust_move_step_sizes(orc_system *s);           /* monte_carlo_utils.f90:98-134 */
int orc_get_step_sizes(const orc_system *s, double out[2]);
int orc_get_counters(const orc_system *s, int64_t out[12]); /* trans,rot,create,delete,swap,widom x (trial,success) */
int orc_get_widom(const orc_system *s, int res, double *sum_weight, int64_t *samples);
int orc_reset_widom(orc_system *s);

/* Widom batch with per-insertion counter-based draws: the contract the GPU
 * batch kernel follows (insertion id -> 5 uniforms), see DESIGN.md */
int orc_widom_batch(orc_system *s, int res, int64_t first_id, int64_t n, uint64_t seed,
                    double *dE_out /* n or NULL */, double *sum_w, int64_t *n_ok);

#ifdef __cplusplus
}
#endif
#endif
