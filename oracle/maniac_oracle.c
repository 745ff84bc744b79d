/*
 * maniac_oracle.c -- CPU ORACLE (test infrastructure, see maniac_oracle.h).
 *
 * Line-by-line restatement, in plain C, of the MANIAC-MC (v0.4.0) energy path
 * and of the move drivers that consume it.  Citations "file:line" are relative
 * to /root/reference/.  Loop orders, accumulation orders, constants and the
 * known quirks of the reference are preserved on purpose; nothing here is
 * tuned.  Build: gcc -O2 -ffp-contract=off (no FMA contraction, like the
 * reference's x86-64 gfortran build).
 *
 * Pinning.  PINNED on every golden value the reference's own tests hold for this
 * path (tests/golden/kat.json, tests/test_oracle_golden.py): total energies vs
 * LAMMPS (LJ-gas, ZIF8-H2O, H2O-gas), the three analytic cases, the methanol
 * create / delete values, the 150-move drift gate, the minimum-image / PBC unit
 * values -- inside the reference's own tolerances.  PARITY UNPINNED (the
 * reference holds no test for them and cannot be built here: Fortran only, no
 * Fortran compiler in the image): per-move dE, accept / reject sequences, swap
 * moves, triclinic energies, step-size adaptation, Widom weights.  For those the
 * oracle is a careful reading of the Fortran, nothing more.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this.
 */
#include "maniac_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* constants: src/constants.f90:8-21, src/parameters.f90:10-57                */
/* ------------------------------------------------------------------------- */
static const double PI = 3.1415926536;                 /* constants.f90:8 (truncated on purpose) */
#define TWOPI (2.0 * PI)                               /* constants.f90:9 */
static const double H_PLANCK = 6.62607015e-34;         /* constants.f90:11 */
static const double EPS0 = 8.854187817e-12;            /* constants.f90:12 */
static const double KB = 1.380658e-23;                 /* constants.f90:14 */
static const double E_CHARGE = 1.602176634e-19;        /* constants.f90:15 */
static const double G_TO_KG = 1.0e-3;                  /* parameters.f90:34 */
static const double M_TO_A = 1.0e10;                   /* parameters.f90:31 */
static const double ERR = 1.0e-10;                     /* parameters.f90:57 "error" */
static const double PROB_CREATE_DELETE = 0.5;          /* parameters.f90:22 */
#define NB_MAX_MOLECULE 5000                           /* parameters.f90:10 */
#define TYPE_HOST 1
#define TYPE_GUEST 2
#define ORTHORHOMBIC 1
#define TRICLINIC 2
#define MAXRES 16

/* NA and J_to_kcal are single-precision literals in the reference
 * (constants.f90:13 "6.02214076e23", parameters.f90:37 "0.000239005736"):
 * the value is rounded to binary32 first, then widened. */
static double NA_(void) { return (double)6.02214076e23f; }
static double J_to_kcal_(void) { return (double)0.000239005736f; }
/* overlap sentinel "1.0e20" is also a single-precision literal
 * (pairwise_energy_utils.f90:117,166) */
static double OVERLAP_(void) { return (double)1.0e20f; }

static double SQRTPI_(void) { return sqrt(PI); }       /* constants.f90:10 */
static double EPS0_INV_real_(void)                     /* constants.f90:17-18, left-to-right */
{
    double eps0_inv = E_CHARGE * E_CHARGE / (4.0 * PI * EPS0);
    return eps0_inv * J_to_kcal_() * M_TO_A * NA_();
}
static double KB_kcalmol_(void) { return KB * NA_() * J_to_kcal_(); } /* constants.f90:21 */

double orc_const_PI(void) { return PI; }
double orc_const_TWOPI(void) { return TWOPI; }
double orc_const_SQRTPI(void) { return SQRTPI_(); }
double orc_const_EPS0_INV_real(void) { return EPS0_INV_real_(); }
double orc_const_KB_kcalmol(void) { return KB_kcalmol_(); }

/* gfortran MODULO for reals without -ffast-math: fmod + sign fix */
static double f_modulo(double a, double p)
{
    double r = fmod(a, p);
    if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p;
    return r;
}
/* gfortran NINT: round half away from zero */
static int f_nint(double x) { return (int)lround(x); }

typedef struct { double re, im; } cplx;
static cplx c_mul(cplx a, cplx b) { cplx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static cplx c_add(cplx a, cplx b) { cplx r = { a.re + b.re, a.im + b.im }; return r; }
static cplx c_sub(cplx a, cplx b) { cplx r = { a.re - b.re, a.im - b.im }; return r; }
static cplx c_scale(double q, cplx a) { cplx r = { q * a.re, q * a.im }; return r; }

/* ------------------------------------------------------------------------- */
/* state: src/simulation_state.f90                                            */
/* ------------------------------------------------------------------------- */
typedef struct {
    int natom, role, active, nmol, cap;
    double mass, lambda, mu, fugacity;
    double *charge;   /* primary%atoms%charges(res, :)            */
    int    *type;     /* primary%atoms%types(res, :), 0-based      */
    double *com;      /* coord%com(:, res, mol)      [cap][3]      */
    double *off;      /* coord%offset(:, res, mol, a) [cap][natom][3] */
    cplx   *factor;   /* ewald%phase%factor_{host,guest}(dim,res,mol,atom,k) -> [cap][natom][3][2K+1] */
} residue_t;

struct orc_system {
    int nres;
    residue_t res[MAXRES];
    /* type_cell, simulation_state.f90:103-112 */
    double matrix[3][3], reciprocal[3][3], lo[3], metrics[9], volume, determinant;
    int shape;
    /* type_coeff at atom-type level (A16) */
    int ntypes;
    double *eps, *sig;
    /* type_ewald, simulation_state.f90:283-293 */
    double tolerance, alpha, screen, fprecision, rc;
    int kmax[3], kmax_max, nk;
    int *kx, *ky, *kz;
    double *k2norm, *k2mag, *kweights, *form_factor;
    cplx *Ak, *Ak_old;
    cplx *factor_old;     /* [natom_max][3][2K+1] */
    int natom_max;
    /* energies */
    double energy[6], e_old[6], e_new[6];
    /* thermo / mc_input / proba */
    double temperature, beta;
    double translation_step, rotation_step_angle;
    double p_translation, p_rotation, p_swap, p_insdel, p_widom;
    /* saved coordinate, simulation_state.f90:95-98 */
    double saved_com[3];
    double *saved_offset;
    /* counters: trans, rot, create, delete, swap, widom x {trial, success} */
    int64_t counter[6][2];
    double *widom_weight;
    int64_t *widom_sample;
    /* RNG */
    uint64_t rng[4];
    const double *ustream;
    int64_t ustream_n, ustream_pos;
    char err[256];
};

static int fail(orc_system *s, const char *msg)
{
    snprintf(s->err, sizeof s->err, "%s", msg);
    return 1;
}
const char *orc_last_error(const orc_system *s) { return s->err; }

orc_system *orc_create(void)
{
    orc_system *s = (orc_system *)calloc(1, sizeof *s);
    if (!s) return NULL;
    s->translation_step = 1.0;
    s->rotation_step_angle = 0.685;
    orc_seed_rng(s, 12345u);
    return s;
}

static void free_ewald(orc_system *s)
{
    free(s->kx); free(s->ky); free(s->kz); free(s->k2norm); free(s->k2mag);
    free(s->kweights); free(s->form_factor); free(s->Ak); free(s->Ak_old); free(s->factor_old);
    s->kx = s->ky = s->kz = NULL; s->k2norm = s->k2mag = s->kweights = s->form_factor = NULL;
    s->Ak = s->Ak_old = s->factor_old = NULL;
    for (int r = 0; r < s->nres; ++r) { free(s->res[r].factor); s->res[r].factor = NULL; }
}

void orc_destroy(orc_system *s)
{
    if (!s) return;
    free_ewald(s);
    for (int r = 0; r < s->nres; ++r) {
        free(s->res[r].charge); free(s->res[r].type); free(s->res[r].com); free(s->res[r].off);
    }
    free(s->eps); free(s->sig); free(s->saved_offset); free(s->widom_weight); free(s->widom_sample);
    free(s);
}

/* ------------------------------------------------------------------------- */
/* RNG contract.  The reference calls libgfortran's random_number
 * (random_utils.f90:20,35), whose stream is unpinned.  The engine
 * and this oracle share an explicit generator instead: xoshiro256** seeded by
 * splitmix64, real64 = top 53 bits * 2^-53 in [0,1).                          */
/* ------------------------------------------------------------------------- */
static uint64_t splitmix64_mix(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
#define GOLDEN 0x9E3779B97F4A7C15ULL
static uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

void orc_seed_rng(orc_system *s, uint64_t seed)
{
    uint64_t z = seed;
    for (int i = 0; i < 4; ++i) { z += GOLDEN; s->rng[i] = splitmix64_mix(z); }
}
void orc_get_rng_state(const orc_system *s, uint64_t st[4]) { memcpy(st, s->rng, sizeof s->rng); }
void orc_set_rng_state(orc_system *s, const uint64_t st[4]) { memcpy(s->rng, st, sizeof s->rng); }
int orc_set_uniform_stream(orc_system *s, const double *u, int64_t n)
{
    s->ustream = u; s->ustream_n = n; s->ustream_pos = 0; return 0;
}
int64_t orc_uniform_stream_used(const orc_system *s) { return s->ustream_pos; }

/* rand_uniform, random_utils.f90:15-22 */
double orc_rand_uniform(orc_system *s)
{
    if (s->ustream) {
        if (s->ustream_pos < s->ustream_n) return s->ustream[s->ustream_pos++];
        fail(s, "uniform stream exhausted");
        return 0.5;
    }
    uint64_t *st = s->rng;
    uint64_t result = rotl64(st[1] * 5u, 7) * 9u;
    uint64_t t = st[1] << 17;
    st[2] ^= st[0]; st[3] ^= st[1]; st[1] ^= st[2]; st[0] ^= st[3];
    st[2] ^= t; st[3] = rotl64(st[3], 45);
    return (double)(result >> 11) * 0x1.0p-53;
}

/* ------------------------------------------------------------------------- */
/* box: geometry_utils.f90:19-35,148-200,293-361                               */
/* ------------------------------------------------------------------------- */
static void cross(const double a[3], const double b[3], double c[3])  /* helper_utils.f90:120-131 */
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static double vnorm(const double v[3]) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); } /* :136-143 */
static void col(const orc_system *s, int j, double v[3]) { for (int i = 0; i < 3; ++i) v[i] = s->matrix[i][j]; }

int orc_set_box(orc_system *s, const double matrix[9], const double lo[3])
{
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) s->matrix[i][j] = matrix[i * 3 + j];
        s->lo[i] = lo[i];
    }
    /* determine_box_symmetry, geometry_utils.f90:341-361 */
    double off_diag[6] = { s->matrix[0][1], s->matrix[0][2], s->matrix[1][0],
                           s->matrix[1][2], s->matrix[2][0], s->matrix[2][1] };
    double mx = 0.0;
    for (int i = 0; i < 6; ++i) if (fabs(off_diag[i]) > mx) mx = fabs(off_diag[i]);
    s->shape = (mx > ERR) ? TRICLINIC : ORTHORHOMBIC;

    /* compute_cell_properties, geometry_utils.f90:293-336 */
    double a[3], b[3], c[3], axb[3], bxc[3], cxa[3];
    col(s, 0, a); col(s, 1, b); col(s, 2, c);
    for (int j = 0; j < 3; ++j) {
        double v[3]; col(s, j, v);
        s->metrics[j] = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    }
    double la = vnorm(a), lb = vnorm(b), lc = vnorm(c);
    s->metrics[3] = (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) / (la * lb);
    s->metrics[4] = (a[0] * c[0] + a[1] * c[1] + a[2] * c[2]) / (la * lc);
    s->metrics[5] = (b[0] * c[0] + b[1] * c[1] + b[2] * c[2]) / (lb * lc);
    cross(a, b, axb); cross(b, c, bxc); cross(c, a, cxa);
    s->volume = fabs(a[0] * bxc[0] + a[1] * bxc[1] + a[2] * bxc[2]);
    s->metrics[6] = s->volume / vnorm(bxc);
    s->metrics[7] = s->volume / vnorm(cxa);
    s->metrics[8] = s->volume / vnorm(axb);

    /* compute_box_determinant_and_inverse, geometry_utils.f90:148-200 */
    double adj[3][3], v1[3], v2[3], cr[3];
    col(s, 1, v1); col(s, 2, v2); cross(v1, v2, cr); for (int i = 0; i < 3; ++i) adj[i][0] = cr[i];
    col(s, 2, v1); col(s, 0, v2); cross(v1, v2, cr); for (int i = 0; i < 3; ++i) adj[i][1] = cr[i];
    col(s, 0, v1); col(s, 1, v2); cross(v1, v2, cr); for (int i = 0; i < 3; ++i) adj[i][2] = cr[i];
    col(s, 0, v1);
    s->determinant = v1[0] * adj[0][0] + v1[1] * adj[1][0] + v1[2] * adj[2][0];
    if (fabs(s->determinant) < 1.0) return fail(s, "Error: Determinant fell into denormal/underflow range");
    double rcp = 1.0 / s->determinant;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) s->reciprocal[i][j] = rcp * adj[i][j];
    return 0;
}

int orc_get_box(const orc_system *s, double matrix[9], double reciprocal[9], double *volume, int *shape)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { matrix[i * 3 + j] = s->matrix[i][j]; reciprocal[i * 3 + j] = s->reciprocal[i][j]; }
    *volume = s->volume; *shape = s->shape;
    return 0;
}

/* apply_PBC, geometry_utils.f90:45-97 */
void orc_apply_PBC(const orc_system *s, double pos[3])
{
    if (s->shape == ORTHORHOMBIC) {
        for (int d = 0; d < 3; ++d) {
            double L = s->matrix[d][d], lo = s->lo[d];
            pos[d] = lo + f_modulo(pos[d] - lo, L);
        }
    } else {
        double rel[3], frac[3];
        for (int d = 0; d < 3; ++d) rel[d] = pos[d] - s->lo[d];
        for (int i = 0; i < 3; ++i) {   /* matmul(reciprocal, rel) */
            frac[i] = s->reciprocal[i][0] * rel[0] + s->reciprocal[i][1] * rel[1] + s->reciprocal[i][2] * rel[2];
            frac[i] = f_modulo(frac[i], 1.0);
        }
        for (int i = 0; i < 3; ++i)     /* lo + matmul(matrix, frac) */
            pos[i] = s->lo[i] + (s->matrix[i][0] * frac[0] + s->matrix[i][1] * frac[1] + s->matrix[i][2] * frac[2]);
    }
}

/* wrap_into_box, geometry_utils.f90:105-140 */
void orc_wrap_into_box(const orc_system *s, double pos[3])
{
    if (s->shape == ORTHORHOMBIC) {
        for (int d = 0; d < 3; ++d) pos[d] = pos[d] - s->matrix[d][d] * f_nint(pos[d] / s->matrix[d][d]);
    } else {
        double f[3];
        for (int i = 0; i < 3; ++i) {
            f[i] = s->reciprocal[i][0] * pos[0] + s->reciprocal[i][1] * pos[1] + s->reciprocal[i][2] * pos[2];
        }
        for (int i = 0; i < 3; ++i) f[i] = f[i] - f_nint(f[i]);
        double p[3];
        for (int i = 0; i < 3; ++i) p[i] = s->matrix[i][0] * f[0] + s->matrix[i][1] * f[1] + s->matrix[i][2] * f[2];
        for (int i = 0; i < 3; ++i) pos[i] = p[i];
    }
}

/* ------------------------------------------------------------------------- */
/* residues, LJ table                                                          */
/* ------------------------------------------------------------------------- */
int orc_add_residue(orc_system *s, int natom, int is_active, const double *charges,
                    const int *types, double mass, int capacity)
{
    if (s->nres >= MAXRES) { fail(s, "too many residue types"); return -1; }
    if (capacity > NB_MAX_MOLECULE) { fail(s, "capacity exceeds NB_MAX_MOLECULE"); return -1; }
    residue_t *r = &s->res[s->nres];
    memset(r, 0, sizeof *r);
    r->natom = natom; r->active = is_active;
    r->role = is_active ? TYPE_GUEST : TYPE_HOST;       /* data_parser.f90:1411-1415 */
    r->cap = capacity; r->mass = mass; r->nmol = 0; r->fugacity = -1.0;
    r->charge = (double *)malloc(sizeof(double) * natom);
    r->type = (int *)malloc(sizeof(int) * natom);
    r->com = (double *)calloc((size_t)capacity * 3, sizeof(double));
    r->off = (double *)calloc((size_t)capacity * natom * 3, sizeof(double));
    memcpy(r->charge, charges, sizeof(double) * natom);
    memcpy(r->type, types, sizeof(int) * natom);
    if (natom > s->natom_max) {
        s->natom_max = natom;
        s->saved_offset = (double *)realloc(s->saved_offset, sizeof(double) * 3 * natom);
    }
    s->widom_weight = (double *)realloc(s->widom_weight, sizeof(double) * (s->nres + 1));
    s->widom_sample = (int64_t *)realloc(s->widom_sample, sizeof(int64_t) * (s->nres + 1));
    s->widom_weight[s->nres] = 0.0; s->widom_sample[s->nres] = 0;
    return s->nres++;
}

int orc_lj_begin(orc_system *s, int ntypes)
{
    free(s->eps); free(s->sig);
    s->ntypes = ntypes;
    s->eps = (double *)calloc((size_t)ntypes * ntypes, sizeof(double));   /* parameters_parser.f90:54-55 */
    s->sig = (double *)calloc((size_t)ntypes * ntypes, sizeof(double));
    return 0;
}
/* one "pair_coeff i j eps sigma" line, parameters_parser.f90:81-99 (set + symmetric) */
int orc_lj_pair_coeff(orc_system *s, int ti, int tj, double eps, double sig)
{
    if (ti < 0 || tj < 0 || ti >= s->ntypes || tj >= s->ntypes) return fail(s, "pair_coeff type out of range");
    s->sig[ti * s->ntypes + tj] = sig; s->eps[ti * s->ntypes + tj] = eps;
    s->sig[tj * s->ntypes + ti] = sig; s->eps[tj * s->ntypes + ti] = eps;
    return 0;
}
/* apply_lorentz_berthelot, parameters_parser.f90:112-178.  The reference loops
 * over (res_i, atom_k, res_j, atom_l); entries depend only on the two atom
 * types, so the loop is restated over the types that occur, in first-occurrence
 * order of that nest. */
int orc_lj_finalize(orc_system *s)
{
    int n = s->ntypes;
    for (int i = 0; i < s->nres; ++i)
        for (int k = 0; k < s->res[i].natom; ++k)
            for (int j = 0; j < s->nres; ++j)
                for (int l = 0; l < s->res[j].natom; ++l) {
                    int ti = s->res[i].type[k], tj = s->res[j].type[l];
                    double *e = &s->eps[ti * n + tj], *g = &s->sig[ti * n + tj];
                    if (fabs(*e) < ERR && fabs(*g) < ERR) {
                        double sigma = (s->sig[ti * n + ti] + s->sig[tj * n + tj]) / 2;
                        double epsilon = sqrt(s->eps[ti * n + ti] * s->eps[tj * n + tj]);
                        if (sigma > ERR && epsilon > ERR) { *g = sigma; *e = epsilon; }
                    }
                }
    return 0;
}
int orc_lj_get(const orc_system *s, double *eps, double *sig)
{
    memcpy(eps, s->eps, sizeof(double) * s->ntypes * s->ntypes);
    memcpy(sig, s->sig, sizeof(double) * s->ntypes * s->ntypes);
    return 0;
}

static int chk(const orc_system *s, int res, int mol)
{
    return (res >= 0 && res < s->nres && mol >= 0 && mol < s->res[res].cap);
}
int orc_set_molecule(orc_system *s, int res, int mol, const double com[3], const double *offset)
{
    if (!chk(s, res, mol)) return fail(s, "set_molecule: index out of range");
    residue_t *r = &s->res[res];
    memcpy(&r->com[(size_t)mol * 3], com, 3 * sizeof(double));
    memcpy(&r->off[(size_t)mol * r->natom * 3], offset, (size_t)r->natom * 3 * sizeof(double));
    return 0;
}
int orc_get_molecule(const orc_system *s, int res, int mol, double com[3], double *offset)
{
    if (!chk(s, res, mol)) return 1;
    const residue_t *r = &s->res[res];
    memcpy(com, &r->com[(size_t)mol * 3], 3 * sizeof(double));
    memcpy(offset, &r->off[(size_t)mol * r->natom * 3], (size_t)r->natom * 3 * sizeof(double));
    return 0;
}
int orc_set_count(orc_system *s, int res, int n)
{
    if (res < 0 || res >= s->nres || n < 0 || n > s->res[res].cap) return fail(s, "set_count: out of range");
    s->res[res].nmol = n; return 0;
}
int orc_get_count(const orc_system *s, int res) { return s->res[res].nmol; }

static void atom_pos(const orc_system *s, int res, int mol, int a, double p[3])
{
    const residue_t *r = &s->res[res];
    const double *c = &r->com[(size_t)mol * 3];
    const double *o = &r->off[((size_t)mol * r->natom + a) * 3];
    p[0] = c[0] + o[0]; p[1] = c[1] + o[1]; p[2] = c[2] + o[2];   /* geometry_utils.f90:235-241 */
}

/* ------------------------------------------------------------------------- */
/* Ewald setup: prepare_utils.f90:110-226, ewald_kvectors.f90:24-175           */
/* ------------------------------------------------------------------------- */
static double normalized_K_squared(int kx, int ky, int kz, const int kmax[3])   /* ewald_kvectors.f90:71-85 */
{
    double a = (double)kx / (double)kmax[0], b = (double)ky / (double)kmax[1], c = (double)kz / (double)kmax[2];
    return a * a + b * b + c * c;
}
static int check_valid_reciprocal_vector(double k2) { return (fabs(k2) >= ERR) && (k2 <= 1.0); } /* :111-123 */

static size_t fac_idx(const orc_system *s, const residue_t *r, int mol, int a, int d, int k)
{
    int W = 2 * s->kmax_max + 1;
    return ((((size_t)mol * r->natom + a) * 3 + d) * W) + (size_t)(k + s->kmax_max);
}
static size_t old_idx(const orc_system *s, int a, int d, int k)
{
    int W = 2 * s->kmax_max + 1;
    return (((size_t)a * 3 + d) * W) + (size_t)(k + s->kmax_max);
}

int orc_setup_ewald(orc_system *s, double tolerance, double real_space_cutoff)
{
    free_ewald(s);
    s->rc = real_space_cutoff;
    /* adjust_real_space_cutoff, prepare_utils.f90:110-135 */
    if (s->rc > s->metrics[0] || s->rc > s->metrics[1] || s->rc > s->metrics[2]) {
        double m = s->metrics[0];
        if (s->metrics[1] < m) m = s->metrics[1];
        if (s->metrics[2] < m) m = s->metrics[2];
        s->rc = m / 2.0;
    }
    /* clamp_tolerance :202-207 */
    s->tolerance = fmin(fabs(tolerance), 0.5);
    /* compute_ewald_parameters :213-226 */
    s->screen = sqrt(fabs(log(s->tolerance * s->rc)));
    s->alpha = sqrt(fabs(log(s->tolerance * s->rc * s->screen))) / s->rc;
    {
        double t = 2.0 * s->screen * s->alpha;
        s->fprecision = sqrt(-log(s->tolerance * s->rc * (t * t)));
    }
    /* compute_fourier_indices :141-169 */
    for (int d = 0; d < 3; ++d) s->kmax[d] = f_nint(0.25 + s->metrics[d] * s->alpha * s->fprecision / PI);
    int count = 0;
    for (int kx = 0; kx <= s->kmax[0]; ++kx)
        for (int ky = -s->kmax[1]; ky <= s->kmax[1]; ++ky)
            for (int kz = -s->kmax[2]; kz <= s->kmax[2]; ++kz) {
                if (kx == 0 && ky == 0 && kz == 0) continue;
                if (check_valid_reciprocal_vector(normalized_K_squared(kx, ky, kz, s->kmax))) ++count;
            }
    s->nk = count;
    s->kmax_max = s->kmax[0];
    if (s->kmax[1] > s->kmax_max) s->kmax_max = s->kmax[1];
    if (s->kmax[2] > s->kmax_max) s->kmax_max = s->kmax[2];

    /* allocate_array :48-103 */
    size_t nk = (size_t)(count > 0 ? count : 1);
    s->kx = (int *)calloc(nk, sizeof(int)); s->ky = (int *)calloc(nk, sizeof(int)); s->kz = (int *)calloc(nk, sizeof(int));
    s->k2norm = (double *)calloc(nk, sizeof(double)); s->k2mag = (double *)calloc(nk, sizeof(double));
    s->kweights = (double *)calloc(nk, sizeof(double)); s->form_factor = (double *)calloc(nk, sizeof(double));
    s->Ak = (cplx *)calloc(nk, sizeof(cplx)); s->Ak_old = (cplx *)calloc(nk, sizeof(cplx));
    int W = 2 * s->kmax_max + 1;
    s->factor_old = (cplx *)calloc((size_t)(s->natom_max > 0 ? s->natom_max : 1) * 3 * W, sizeof(cplx));
    for (int r = 0; r < s->nres; ++r)
        s->res[r].factor = (cplx *)calloc((size_t)s->res[r].cap * s->res[r].natom * 3 * W, sizeof(cplx));

    /* precompute_valid_reciprocal_vectors, ewald_kvectors.f90:24-65 */
    double kvec_matrix[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) kvec_matrix[i][j] = TWOPI * s->reciprocal[i][j];
    count = 0;
    for (int kx = 0; kx <= s->kmax[0]; ++kx)
        for (int ky = -s->kmax[1]; ky <= s->kmax[1]; ++ky)
            for (int kz = -s->kmax[2]; kz <= s->kmax[2]; ++kz) {
                if (kx == 0 && ky == 0 && kz == 0) continue;
                double k2 = normalized_K_squared(kx, ky, kz, s->kmax);
                if (!check_valid_reciprocal_vector(k2)) continue;
                /* compute_cartesian_k_squared :155-175 (columns of kvec_matrix) */
                double kv[3];
                for (int i = 0; i < 3; ++i)
                    kv[i] = (double)kx * kvec_matrix[i][0] + (double)ky * kvec_matrix[i][1] + (double)kz * kvec_matrix[i][2];
                double k2mag = kv[0] * kv[0] + kv[1] * kv[1] + kv[2] * kv[2];
                s->kx[count] = kx; s->ky[count] = ky; s->kz[count] = kz;
                s->k2norm[count] = k2; s->k2mag[count] = k2mag;
                s->form_factor[count] = (kx == 0) ? 1.0 : 2.0;     /* :91-105 */
                ++count;
            }
    return 0;
}

int orc_get_ewald(const orc_system *s, double *alpha, int kmax[3], int *nk, double *rc)
{
    *alpha = s->alpha; *nk = s->nk; *rc = s->rc;
    for (int d = 0; d < 3; ++d) kmax[d] = s->kmax[d];
    return 0;
}
int orc_get_kvectors(const orc_system *s, int *kx, int *ky, int *kz, double *k2, double *ff)
{
    for (int i = 0; i < s->nk; ++i) { kx[i] = s->kx[i]; ky[i] = s->ky[i]; kz[i] = s->kz[i]; k2[i] = s->k2mag[i]; ff[i] = s->form_factor[i]; }
    return 0;
}

/* compute_reciprocal_weights, ewald_kvectors.f90:130-149 */
static void compute_reciprocal_weights(orc_system *s)
{
    double alpha_squared = s->alpha * s->alpha;
    for (int i = 0; i < s->nk; ++i) {
        double k2 = s->k2mag[i];
        s->kweights[i] = exp(-k2 / (4.0 * alpha_squared)) / k2;
    }
}

/* ------------------------------------------------------------------------- */
/* thermo: prepare_monte_carlo, prepare_utils.f90:231-259                      */
/* ------------------------------------------------------------------------- */
int orc_set_thermo(orc_system *s, double temperature)
{
    s->temperature = temperature;
    s->beta = 1 / (KB_kcalmol_() * temperature);
    for (int r = 0; r < s->nres; ++r) {
        residue_t *R = &s->res[r];
        if (!R->active) continue;
        if (R->fugacity >= 0.0) R->mu = log(R->fugacity) / s->beta;
        double mass = R->mass * G_TO_KG / NA_();
        R->lambda = H_PLANCK / sqrt(TWOPI * mass * KB * temperature);
        R->lambda = R->lambda * M_TO_A;
    }
    return 0;
}
int orc_set_fugacity(orc_system *s, int res, double f)
{
    s->res[res].fugacity = f;
    if (f >= 0.0 && s->beta != 0.0) s->res[res].mu = log(f) / s->beta;
    return 0;
}
int orc_set_chemical_potential(orc_system *s, int res, double mu) { s->res[res].mu = mu; return 0; }
int orc_set_mc_input(orc_system *s, double tstep, double rstep, double pt, double pr, double ps, double pid, double pw)
{
    s->translation_step = tstep; s->rotation_step_angle = rstep;
    s->p_translation = pt; s->p_rotation = pr; s->p_swap = ps; s->p_insdel = pid; s->p_widom = pw;
    return 0;
}
/* adjust_move_step_sizes, monte_carlo_utils.f90:98-134 (Robbins-Monro, called at the end of a block when
 * mc_input%recalibrate_moves is set; constants of parameters.f90:17-24) */
int orc_adjust_move_step_sizes(orc_system *s)
{
    const double gamma = 0.10, TARGET_ACCEPTANCE = 0.40;
    const double MIN_TRANSLATION_STEP = 1.0e-3, MAX_TRANSLATION_STEP = 3.0, MIN_ROTATION_ANGLE = 1.0e-3, MAX_ROTATION_ANGLE = 0.78;
    const int64_t MIN_TRIALS_FOR_RECALIBRATION = 500;
    if (s->counter[0][0] > MIN_TRIALS_FOR_RECALIBRATION) {
        double acc_trans = (double)s->counter[0][1] / (double)s->counter[0][0];
        s->translation_step = s->translation_step * exp(gamma * (acc_trans - TARGET_ACCEPTANCE));
        s->translation_step = fmax(MIN_TRANSLATION_STEP, fmin(s->translation_step, MAX_TRANSLATION_STEP));
    }
    if (s->counter[1][0] > MIN_TRIALS_FOR_RECALIBRATION) {
        double acc_rot = (double)s->counter[1][1] / (double)s->counter[1][0];
        s->rotation_step_angle = s->rotation_step_angle * exp(gamma * (acc_rot - TARGET_ACCEPTANCE));
        s->rotation_step_angle = fmax(MIN_ROTATION_ANGLE, fmin(s->rotation_step_angle, MAX_ROTATION_ANGLE));
    }
    return 0;
}
int orc_get_step_sizes(const orc_system *s, double out[2]) { out[0] = s->translation_step; out[1] = s->rotation_step_angle; return 0; }
double orc_get_beta(const orc_system *s) { return s->beta; }
double orc_get_lambda(const orc_system *s, int res) { return s->res[res].lambda; }
double orc_get_mu(const orc_system *s, int res) { return s->res[res].mu; }

/* ------------------------------------------------------------------------- */
/* minimum_image_distance, geometry_utils.f90:210-284                          */
/* ------------------------------------------------------------------------- */
double orc_minimum_image_distance(const orc_system *s, int r1, int m1, int a1, int r2, int m2, int a2)
{
    double pos1[3], pos2[3], delta[3];
    atom_pos(s, r1, m1, a1, pos1);
    atom_pos(s, r2, m2, a2, pos2);
    for (int d = 0; d < 3; ++d) delta[d] = pos2[d] - pos1[d];
    if (s->shape == ORTHORHOMBIC) {
        for (int d = 0; d < 3; ++d) {
            double L = s->matrix[d][d];
            delta[d] = f_modulo(delta[d] + 0.5 * L, L) - 0.5 * L;
        }
        return vnorm(delta);
    }
    double min_dist2 = 1.7976931348623157e308;   /* huge(one) */
    for (int sx = -1; sx <= 1; ++sx)
        for (int sy = -1; sy <= 1; ++sy)
            for (int sz = -1; sz <= 1; ++sz) {
                double t[3];
                for (int i = 0; i < 3; ++i)      /* columns of matrix used as cell vectors (quirk) */
                    t[i] = delta[i] + sx * s->matrix[i][0] + sy * s->matrix[i][1] + sz * s->matrix[i][2];
                double d2 = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
                if (d2 < min_dist2) min_dist2 = d2;
            }
    return sqrt(min_dist2);
}

/* pairwise_lj_energy, pairwise_energy_utils.f90:95-138 (use_table = .false.) */
static double pairwise_lj_energy(const orc_system *s, double r, double sigma, double epsilon)
{
    if (r >= s->rc) return 0.0;
    if (r < ERR) return OVERLAP_();
    double x = sigma / r;
    double x2 = x * x, x4 = x2 * x2;
    double r6 = x2 * x4;            /* (sigma/r)**6 via __powidf2 at -O0 */
    double r12 = r6 * r6;
    return 4.0 * epsilon * (r12 - r6);
}
/* pairwise_coulomb_energy, pairwise_energy_utils.f90:143-179 */
static double pairwise_coulomb_energy(const orc_system *s, double r, double q_i, double q_j)
{
    if (fabs(q_i) < ERR || fabs(q_j) < ERR) return 0.0;
    if (r < ERR) return OVERLAP_();
    return q_i * q_j * erfc(s->alpha * r) / r;
}

/* pairwise_energy_for_molecule, pairwise_energy_utils.f90:21-90 */
int orc_pairwise_energy_for_molecule(const orc_system *s, int res_i, int mol_i, int skip_ordering_check,
                                     double *e_non_coulomb, double *e_coulomb)
{
    double enc = 0.0, ec = 0.0;
    const residue_t *Ri = &s->res[res_i];
    for (int atom_i = 0; atom_i < Ri->natom; ++atom_i)
        for (int res_j = 0; res_j < s->nres; ++res_j) {
            const residue_t *Rj = &s->res[res_j];
            for (int mol_j = 0; mol_j < Rj->nmol; ++mol_j) {
                if (mol_i == mol_j && res_i == res_j) continue;
                if (!skip_ordering_check)
                    if (res_j < res_i || (res_j == res_i && mol_j <= mol_i)) continue;
                for (int atom_j = 0; atom_j < Rj->natom; ++atom_j) {
                    int ti = Ri->type[atom_i], tj = Rj->type[atom_j];
                    double sigma = s->sig[ti * s->ntypes + tj];
                    double epsilon = s->eps[ti * s->ntypes + tj];
                    double q_i = Ri->charge[atom_i], q_j = Rj->charge[atom_j];
                    double r = orc_minimum_image_distance(s, res_i, mol_i, atom_i, res_j, mol_j, atom_j);
                    enc = enc + pairwise_lj_energy(s, r, sigma, epsilon);
                    ec = ec + pairwise_coulomb_energy(s, r, q_i, q_j);
                }
            }
        }
    *e_non_coulomb = enc;
    *e_coulomb = ec * EPS0_INV_real_();
    return 0;
}

/* ------------------------------------------------------------------------- */
/* phase factors: ewald_phase.f90                                              */
/* ------------------------------------------------------------------------- */
/* compute_ewald_phase_factors :205-250, compute_atom_phase :255-280,
 * compute_phase_factor :286-312 */
int orc_compute_ewald_phase_factors(orc_system *s, int res, int mol)
{
    residue_t *R = &s->res[res];
    for (int a = 0; a < R->natom; ++a) {
        double atom[3], phase[3];
        atom_pos(s, res, mol, a, atom);
        for (int i = 0; i < 3; ++i) {
            phase[i] = 0.0;
            for (int j = 0; j < 3; ++j) phase[i] = phase[i] + s->reciprocal[j][i] * atom[j];
            phase[i] = TWOPI * phase[i];
        }
        for (int d = 0; d < 3; ++d)
            for (int k = 0; k <= s->kmax[d]; ++k) {
                cplx e = { cos(k * phase[d]), sin(k * phase[d]) };
                R->factor[fac_idx(s, R, mol, a, d, k)] = e;
                if (k != 0) { cplx c = { e.re, -e.im }; R->factor[fac_idx(s, R, mol, a, d, -k)] = c; }
            }
    }
    return 0;
}
static void compute_all_ewald_phase_factors(orc_system *s)   /* :180-199 */
{
    for (int r = 0; r < s->nres; ++r)
        for (int m = 0; m < s->res[r].nmol; ++m) orc_compute_ewald_phase_factors(s, r, m);
}

/* save_single_mol_fourier_terms :17-68 */
int orc_save_single_mol_fourier_terms(orc_system *s, int res, int mol)
{
    residue_t *R = &s->res[res];
    if (R->role == TYPE_HOST) return fail(s, "Inconsistence in fourier routine");
    for (int a = 0; a < R->natom; ++a)
        for (int d = 0; d < 3; ++d)
            for (int k = 0; k <= s->kmax[d]; ++k) {
                s->factor_old[old_idx(s, a, d, k)] = R->factor[fac_idx(s, R, mol, a, d, k)];
                if (k != 0 && d != 0) s->factor_old[old_idx(s, a, d, -k)] = R->factor[fac_idx(s, R, mol, a, d, -k)];
            }
    for (int i = 0; i < s->nk; ++i) s->Ak_old[i] = s->Ak[i];
    return 0;
}
/* restore_single_mol_fourier :74-125 */
int orc_restore_single_mol_fourier(orc_system *s, int res, int mol)
{
    residue_t *R = &s->res[res];
    if (R->role == TYPE_HOST) return fail(s, "Inconsistence in fourier routine");
    for (int a = 0; a < R->natom; ++a)
        for (int d = 0; d < 3; ++d)
            for (int k = 0; k <= s->kmax[d]; ++k) {
                R->factor[fac_idx(s, R, mol, a, d, k)] = s->factor_old[old_idx(s, a, d, k)];
                if (k != 0 && d != 0) R->factor[fac_idx(s, R, mol, a, d, -k)] = s->factor_old[old_idx(s, a, d, -k)];
            }
    for (int i = 0; i < s->nk; ++i) s->Ak[i] = s->Ak_old[i];
    return 0;
}
/* replace_fourier_terms_single_mol :131-175 */
static void replace_fourier_terms_single_mol(orc_system *s, int res, int i1, int i2)
{
    residue_t *R = &s->res[res];
    for (int a = 0; a < R->natom; ++a)
        for (int d = 0; d < 3; ++d)
            for (int k = 0; k <= s->kmax[d]; ++k) {
                R->factor[fac_idx(s, R, i1, a, d, k)] = R->factor[fac_idx(s, R, i2, a, d, k)];
                if (k != 0 && d != 0) R->factor[fac_idx(s, R, i1, a, d, -k)] = R->factor[fac_idx(s, R, i2, a, d, -k)];
            }
}

/* ------------------------------------------------------------------------- */
/* ewald_energy.f90                                                            */
/* ------------------------------------------------------------------------- */
static double amplitude_squared(cplx z)    /* helper_utils.f90:97-106: real(z*conjg(z)) */
{
    cplx c = { z.re, -z.im };
    return c_mul(z, c).re;
}

/* compute_all_recip_amplitude :261-328 */
static cplx compute_all_recip_amplitude(const orc_system *s, int kx, int ky, int kz)
{
    cplx Ak = { 0.0, 0.0 };
    for (int r = 0; r < s->nres; ++r) {
        const residue_t *R = &s->res[r];
        for (int m = 0; m < R->nmol; ++m)
            for (int a = 0; a < R->natom; ++a) {
                double q = R->charge[a];
                cplx p = c_mul(c_mul(R->factor[fac_idx(s, R, m, a, 0, kx)], R->factor[fac_idx(s, R, m, a, 1, ky)]),
                               R->factor[fac_idx(s, R, m, a, 2, kz)]);
                Ak = c_add(Ak, c_scale(q, p));
            }
    }
    return Ak;
}

/* compute_total_reciprocal_energy :20-58 */
static void compute_total_reciprocal_energy(orc_system *s)
{
    double Ek = 0.0;
    for (int i = 0; i < s->nk; ++i) {
        s->Ak[i] = compute_all_recip_amplitude(s, s->kx[i], s->ky[i], s->kz[i]);
        double Ak_square = amplitude_squared(s->Ak[i]);
        Ek = Ek + s->form_factor[i] * s->kweights[i] * Ak_square;
    }
    s->energy[ORC_E_RECIP] = Ek * EPS0_INV_real_() * TWOPI / s->volume;
}

/* update_reciprocal_amplitude_single_mol :64-134 */
int orc_update_reciprocal_amplitude_single_mol(orc_system *s, int res, int mol, int kind)
{
    residue_t *R = &s->res[res];
    if (R->role == TYPE_HOST) return fail(s, "Inconsistence in fourier routine");
    int natoms = R->natom;
    for (int i = 0; i < s->nk; ++i) {
        int kx = s->kx[i], ky = s->ky[i], kz = s->kz[i];
        cplx sum = { 0.0, 0.0 };
        for (int a = 0; a < natoms; ++a) {
            cplx pn = c_mul(c_mul(R->factor[fac_idx(s, R, mol, a, 0, kx)], R->factor[fac_idx(s, R, mol, a, 1, ky)]),
                            R->factor[fac_idx(s, R, mol, a, 2, kz)]);
            cplx po = c_mul(c_mul(s->factor_old[old_idx(s, a, 0, kx)], s->factor_old[old_idx(s, a, 1, ky)]),
                            s->factor_old[old_idx(s, a, 2, kz)]);
            double q = R->charge[a];
            if (kind == ORC_KIND_CREATE) sum = c_add(sum, c_scale(q, pn));
            else if (kind == ORC_KIND_DELETE) sum = c_add(sum, c_scale(q, po));
            else sum = c_add(sum, c_scale(q, c_sub(pn, po)));
        }
        if (kind == ORC_KIND_DELETE) s->Ak[i] = c_sub(s->Ak[i], sum);
        else s->Ak[i] = c_add(s->Ak[i], sum);
    }
    return 0;
}

/* reciprocal_ewald_energy :139-164 */
double orc_reciprocal_ewald_energy(const orc_system *s)
{
    double e = 0.0;
    for (int i = 0; i < s->nk; ++i) {
        double a2 = amplitude_squared(s->Ak[i]);
        e = e + s->form_factor[i] * s->kweights[i] * a2;
    }
    return e * EPS0_INV_real_() * TWOPI / s->volume;
}
int orc_get_Ak(const orc_system *s, double *re_im)
{
    for (int i = 0; i < s->nk; ++i) { re_im[2 * i] = s->Ak[i].re; re_im[2 * i + 1] = s->Ak[i].im; }
    return 0;
}

/* ewald_self_energy_single_mol :177-205 (== single_mol_ewald_self, self_energy_utils.f90:56-87) */
double orc_ewald_self_energy_single_mol(const orc_system *s, int res)
{
    const residue_t *R = &s->res[res];
    double self_energy = 0.0;
    for (int a = 0; a < R->natom; ++a) {
        double charge = R->charge[a];
        if (fabs(charge) < ERR) continue;
        self_energy = self_energy - s->alpha / SQRTPI_() * (charge * charge);
    }
    return self_energy * EPS0_INV_real_();
}

/* intra_res_real_coulomb_energy :212-252 */
double orc_intra_res_real_coulomb_energy(const orc_system *s, int res, int mol)
{
    const residue_t *R = &s->res[res];
    double u = 0;
    for (int a1 = 0; a1 < R->natom - 1; ++a1) {
        double c1 = R->charge[a1];
        for (int a2 = a1 + 1; a2 < R->natom; ++a2) {
            double c2 = R->charge[a2];
            double distance = orc_minimum_image_distance(s, res, mol, a1, res, mol, a2);
            if (distance < ERR) continue;
            u = u + c1 * c2 * (erfc(s->alpha * distance) - 1.0) / distance;
        }
    }
    return u * EPS0_INV_real_();
}

/* ------------------------------------------------------------------------- */
/* energy_utils.f90                                                            */
/* ------------------------------------------------------------------------- */
int orc_update_system_energy(orc_system *s, double out[6])
{
    /* evaluate_pairwise_energy :77-108 */
    s->energy[ORC_E_NON_COULOMB] = 0.0; s->energy[ORC_E_COULOMB] = 0.0;
    for (int r = 0; r < s->nres; ++r)
        for (int m = 0; m < s->res[r].nmol; ++m) {
            double enc, ec;
            orc_pairwise_energy_for_molecule(s, r, m, 0, &enc, &ec);
            s->energy[ORC_E_NON_COULOMB] += enc;
            s->energy[ORC_E_COULOMB] += ec;
        }
    /* evaluate_ewald_self_energy, self_energy_utils.f90:24-47 */
    s->energy[ORC_E_SELF] = 0.0;
    for (int r = 0; r < s->nres; ++r) {
        double e = orc_ewald_self_energy_single_mol(s, r);
        e = e * s->res[r].nmol;
        s->energy[ORC_E_SELF] = s->energy[ORC_E_SELF] + e;
    }
    /* evaluate_ewald_recip_energy :118-132 */
    compute_reciprocal_weights(s);
    compute_all_ewald_phase_factors(s);
    compute_total_reciprocal_energy(s);
    /* evaluate_intra_real_coulomb_energy :46-75 (active residues only) */
    s->energy[ORC_E_INTRA] = 0.0;
    for (int r = 0; r < s->nres; ++r) {
        if (!s->res[r].active) continue;
        for (int m = 0; m < s->res[r].nmol; ++m)
            s->energy[ORC_E_INTRA] = s->energy[ORC_E_INTRA] + orc_intra_res_real_coulomb_energy(s, r, m);
    }
    /* :34-35 */
    s->energy[ORC_E_TOTAL] = s->energy[ORC_E_RECIP] + s->energy[ORC_E_NON_COULOMB] + s->energy[ORC_E_COULOMB]
                           + s->energy[ORC_E_SELF] + s->energy[ORC_E_INTRA];
    if (out) memcpy(out, s->energy, sizeof s->energy);
    return 0;
}
int orc_get_energy(const orc_system *s, double out[6]) { memcpy(out, s->energy, sizeof s->energy); return 0; }

/* ------------------------------------------------------------------------- */
/* monte_carlo_utils.f90: compute_old_energy :367-423, compute_new_energy :300-361 */
/* ------------------------------------------------------------------------- */
static void total_of(double e[6])
{
    e[ORC_E_TOTAL] = e[ORC_E_NON_COULOMB] + e[ORC_E_COULOMB] + e[ORC_E_RECIP] + e[ORC_E_SELF] + e[ORC_E_INTRA];
}

int orc_compute_old_energy(orc_system *s, int res, int mol, int kind, double out[6])
{
    double *o = s->e_old;
    if (kind == ORC_KIND_CREATE) {
        o[ORC_E_NON_COULOMB] = 0.0; o[ORC_E_COULOMB] = 0.0; o[ORC_E_SELF] = 0.0; o[ORC_E_INTRA] = 0.0;
        o[ORC_E_RECIP] = s->energy[ORC_E_RECIP];
    } else if (kind == ORC_KIND_DELETE) {
        o[ORC_E_SELF] = orc_ewald_self_energy_single_mol(s, res);
        o[ORC_E_INTRA] = orc_intra_res_real_coulomb_energy(s, res, mol);
        orc_pairwise_energy_for_molecule(s, res, mol, 1, &o[ORC_E_NON_COULOMB], &o[ORC_E_COULOMB]);
        o[ORC_E_RECIP] = s->energy[ORC_E_RECIP];
    } else {
        o[ORC_E_SELF] = 0.0; o[ORC_E_INTRA] = 0.0;
        o[ORC_E_RECIP] = s->energy[ORC_E_RECIP];
        orc_pairwise_energy_for_molecule(s, res, mol, 1, &o[ORC_E_NON_COULOMB], &o[ORC_E_COULOMB]);
    }
    total_of(o);
    if (out) memcpy(out, o, 6 * sizeof(double));
    return 0;
}

int orc_compute_new_energy(orc_system *s, int res, int mol, int kind, double out[6])
{
    double *n = s->e_new;
    if (kind == ORC_KIND_CREATE) {
        orc_compute_ewald_phase_factors(s, res, mol);
        orc_update_reciprocal_amplitude_single_mol(s, res, mol, ORC_KIND_CREATE);
        n[ORC_E_RECIP] = orc_reciprocal_ewald_energy(s);
        orc_pairwise_energy_for_molecule(s, res, mol, 1, &n[ORC_E_NON_COULOMB], &n[ORC_E_COULOMB]);
        n[ORC_E_SELF] = orc_ewald_self_energy_single_mol(s, res);
        n[ORC_E_INTRA] = orc_intra_res_real_coulomb_energy(s, res, mol);
    } else if (kind == ORC_KIND_DELETE) {
        n[ORC_E_NON_COULOMB] = 0; n[ORC_E_COULOMB] = 0; n[ORC_E_SELF] = 0; n[ORC_E_INTRA] = 0;
        orc_update_reciprocal_amplitude_single_mol(s, res, mol, ORC_KIND_DELETE);
        n[ORC_E_RECIP] = orc_reciprocal_ewald_energy(s);
    } else {
        n[ORC_E_SELF] = 0; n[ORC_E_INTRA] = 0;
        orc_compute_ewald_phase_factors(s, res, mol);
        orc_update_reciprocal_amplitude_single_mol(s, res, mol, ORC_KIND_MOVE);
        n[ORC_E_RECIP] = orc_reciprocal_ewald_energy(s);
        orc_pairwise_energy_for_molecule(s, res, mol, 1, &n[ORC_E_NON_COULOMB], &n[ORC_E_COULOMB]);
    }
    total_of(n);
    if (out) memcpy(out, n, 6 * sizeof(double));
    return 0;
}

/* ------------------------------------------------------------------------- */
/* acceptance rules: monte_carlo_utils.f90:204-294                             */
/* ------------------------------------------------------------------------- */
enum { TYPE_CREATION = 1, TYPE_DELETION = 2, TYPE_TRANSLATION = 3, TYPE_ROTATION = 4 };

static double compute_acceptance_probability(const orc_system *s, int res, int move_type)
{
    const residue_t *R = &s->res[res];
    double N = (double)R->nmol, Nplus1 = N + 1.0;
    double deltaU = s->e_new[ORC_E_TOTAL] - s->e_old[ORC_E_TOTAL];
    double mu = R->mu, lambda = R->lambda, prefactor;
    switch (move_type) {
    case TYPE_CREATION:
        prefactor = s->volume / N / (lambda * lambda * lambda);
        return fmin(1.0, prefactor * exp(-s->beta * (deltaU - mu)));
    case TYPE_DELETION:
        prefactor = Nplus1 * (lambda * lambda * lambda) / s->volume;
        return fmin(1.0, prefactor * exp(-s->beta * (deltaU + mu)));
    default:
        return fmin(1.0, exp(-s->beta * deltaU));
    }
}
static double swap_acceptance_probability(const orc_system *s, int type_old, int type_new)
{
    double N_new = (double)s->res[type_new].nmol, N_old = (double)s->res[type_old].nmol;
    double Nplus1 = N_new + 1.0;
    double deltaU = s->e_new[ORC_E_TOTAL] - s->e_old[ORC_E_TOTAL];
    return fmin(1.0, (N_old / Nplus1) * exp(-s->beta * (deltaU + s->res[type_new].mu - s->res[type_old].mu)));
}

/* ------------------------------------------------------------------------- */
/* helpers of the move drivers                                                 */
/* ------------------------------------------------------------------------- */
/* return_rotation_matrix, helper_utils.f90:30-75 */
static void return_rotation_matrix(int axis, double theta, double R[3][3])
{
    double c = cos(theta), sn = sin(theta);
    memset(R, 0, 9 * sizeof(double));
    R[0][0] = R[1][1] = R[2][2] = 1.0;
    if (axis == 1) { R[1][1] = c; R[1][2] = -sn; R[2][1] = sn; R[2][2] = c; }
    else if (axis == 2) { R[0][0] = c; R[0][2] = sn; R[2][0] = -sn; R[2][2] = c; }
    else { R[0][0] = c; R[0][1] = -sn; R[1][0] = sn; R[1][1] = c; }
}
/* choose_rotation_angle, monte_carlo_utils.f90:70-92 */
static double choose_rotation_angle(orc_system *s, int full)
{
    if (!full) return (orc_rand_uniform(s) - 0.5) * s->rotation_step_angle;
    return orc_rand_uniform(s) * TWOPI;
}
/* apply_random_rotation, monte_carlo_utils.f90:25-64 */
static void apply_random_rotation(orc_system *s, int res, int mol, int full)
{
    residue_t *R = &s->res[res];
    if (R->natom == 1) return;
    double theta = choose_rotation_angle(s, full);
    int axis = (int)(orc_rand_uniform(s) * 3.0) + 1;
    double M[3][3];
    return_rotation_matrix(axis, theta, M);
    for (int a = 0; a < R->natom; ++a) {
        double *o = &R->off[((size_t)mol * R->natom + a) * 3];
        double v[3] = { o[0], o[1], o[2] };
        for (int i = 0; i < 3; ++i)      /* matmul: sum over j in order */
            o[i] = M[i][0] * v[0] + M[i][1] * v[1] + M[i][2] * v[2];
    }
}
/* update_counts, monte_carlo_utils.f90:662-672 */
static void update_counts(orc_system *s, int res, int sign) { s->res[res].nmol += sign; }

/* save_molecule_state :449-474 */
static void save_molecule_state(orc_system *s, int res, int mol, int want_com, int want_off)
{
    residue_t *R = &s->res[res];
    orc_save_single_mol_fourier_terms(s, res, mol);
    if (want_com) memcpy(s->saved_com, &R->com[(size_t)mol * 3], 3 * sizeof(double));
    if (want_off) memcpy(s->saved_offset, &R->off[(size_t)mol * R->natom * 3], (size_t)R->natom * 3 * sizeof(double));
}
/* reject_molecule_move :480-505 */
static void reject_molecule_move(orc_system *s, int res, int mol, int have_com, int have_off)
{
    residue_t *R = &s->res[res];
    if (have_com) memcpy(&R->com[(size_t)mol * 3], s->saved_com, 3 * sizeof(double));
    if (have_off) memcpy(&R->off[(size_t)mol * R->natom * 3], s->saved_offset, (size_t)R->natom * 3 * sizeof(double));
    orc_restore_single_mol_fourier(s, res, mol);
}
/* accept_molecule_move :429-442 */
static void accept_molecule_move(orc_system *s, int64_t counter_var[2])
{
    double *e = s->energy, *o = s->e_old, *n = s->e_new;
    e[ORC_E_RECIP] = n[ORC_E_RECIP];
    e[ORC_E_NON_COULOMB] = e[ORC_E_NON_COULOMB] + n[ORC_E_NON_COULOMB] - o[ORC_E_NON_COULOMB];
    e[ORC_E_COULOMB] = e[ORC_E_COULOMB] + n[ORC_E_COULOMB] - o[ORC_E_COULOMB];
    e[ORC_E_TOTAL] = e[ORC_E_TOTAL] + n[ORC_E_TOTAL] - o[ORC_E_TOTAL];
    counter_var[1] += 1;
}
/* energy bookkeeping shared by accept_creation/deletion/swap (creation.f90:93-97,
 * deletion.f90:93-98, swapping.f90:145-150) */
static void accept_all_components(orc_system *s)
{
    double *e = s->energy, *o = s->e_old, *n = s->e_new;
    e[ORC_E_RECIP] = n[ORC_E_RECIP];
    e[ORC_E_NON_COULOMB] = e[ORC_E_NON_COULOMB] + n[ORC_E_NON_COULOMB] - o[ORC_E_NON_COULOMB];
    e[ORC_E_COULOMB] = e[ORC_E_COULOMB] + n[ORC_E_COULOMB] - o[ORC_E_COULOMB];
    e[ORC_E_SELF] = e[ORC_E_SELF] + n[ORC_E_SELF] - o[ORC_E_SELF];
    e[ORC_E_INTRA] = e[ORC_E_INTRA] + n[ORC_E_INTRA] - o[ORC_E_INTRA];
    e[ORC_E_TOTAL] = e[ORC_E_TOTAL] + n[ORC_E_TOTAL] - o[ORC_E_TOTAL];
}

/* insert_and_orient_molecule :512-573, no-reservoir branch */
static void insert_and_orient_molecule(orc_system *s, int res, int mol, int place_random_com)
{
    residue_t *R = &s->res[res];
    if (place_random_com) {
        double t[3];
        for (int i = 0; i < 3; ++i) t[i] = orc_rand_uniform(s);     /* call random_number(trial_pos) */
        for (int i = 0; i < 3; ++i)
            R->com[(size_t)mol * 3 + i] = s->lo[i] + (s->matrix[i][0] * t[0] + s->matrix[i][1] * t[1] + s->matrix[i][2] * t[2]);
    }
    /* copy site offsets from the first molecule (:563-564) */
    if (mol != 0) memmove(&R->off[(size_t)mol * R->natom * 3], &R->off[0], (size_t)R->natom * 3 * sizeof(double));
    apply_random_rotation(s, res, mol, 1);
}
/* remove_molecule :642-657 */
static void remove_molecule(orc_system *s, int res, int mol, int last)
{
    residue_t *R = &s->res[res];
    if (mol != last) {
        memcpy(&R->com[(size_t)mol * 3], &R->com[(size_t)last * 3], 3 * sizeof(double));
        memcpy(&R->off[(size_t)mol * R->natom * 3], &R->off[(size_t)last * R->natom * 3], (size_t)R->natom * 3 * sizeof(double));
    }
    replace_fourier_terms_single_mol(s, res, mol, last);
}
/* reject_creation_move :579-591 */
static void reject_creation_move(orc_system *s, int res, int mol)
{
    update_counts(s, res, -1);
    orc_restore_single_mol_fourier(s, res, mol);
}

static void trace_fill(orc_system *s, orc_step_trace *t, int move, int res, int mol, int accepted, double prob)
{
    if (!t) return;
    t->move = move; t->res = res; t->mol = mol; t->accepted = accepted; t->prob = prob;
    t->dE = s->e_new[ORC_E_TOTAL] - s->e_old[ORC_E_TOTAL];
    memcpy(t->e_old, s->e_old, sizeof t->e_old);
    memcpy(t->e_new, s->e_new, sizeof t->e_new);
}
static void trace_none(orc_step_trace *t, int res, int mol)
{
    if (!t) return;
    memset(t, 0, sizeof *t);
    t->move = ORC_MV_NONE; t->res = res; t->mol = mol;
}

/* ------------------------------------------------------------------------- */
/* move drivers                                                                */
/* ------------------------------------------------------------------------- */
/* attempt_translation_move, translation.f90:21-66; propose :75-97 */
int orc_attempt_translation_move(orc_system *s, int res, int mol, orc_step_trace *t)
{
    trace_none(t, res, mol);
    if (mol < 0) return 0;                                   /* mol_index == 0 */
    residue_t *R = &s->res[res];
    s->counter[0][0] += 1;
    save_molecule_state(s, res, mol, 1, 0);
    orc_compute_old_energy(s, res, mol, ORC_KIND_MOVE, NULL);
    {
        double tp[3];
        for (int i = 0; i < 3; ++i) tp[i] = orc_rand_uniform(s);          /* rand_symmetric(3) */
        for (int i = 0; i < 3; ++i) tp[i] = (tp[i] - 0.5) * s->translation_step;
        double *c = &R->com[(size_t)mol * 3];
        for (int i = 0; i < 3; ++i) c[i] = c[i] + tp[i];
        orc_apply_PBC(s, c);
    }
    orc_compute_new_energy(s, res, mol, ORC_KIND_MOVE, NULL);
    double p = compute_acceptance_probability(s, res, TYPE_TRANSLATION);
    int acc = orc_rand_uniform(s) <= p;
    if (acc) accept_molecule_move(s, s->counter[0]);
    else reject_molecule_move(s, res, mol, 1, 0);
    trace_fill(s, t, ORC_MV_TRANSLATE, res, mol, acc, p);
    return 0;
}

/* attempt_rotation_move, rotation.f90:22-60 */
int orc_attempt_rotation_move(orc_system *s, int res, int mol, orc_step_trace *t)
{
    trace_none(t, res, mol);
    if (s->res[res].natom == 1 || mol < 0) return 0;
    s->counter[1][0] += 1;
    save_molecule_state(s, res, mol, 0, 1);
    orc_compute_old_energy(s, res, mol, ORC_KIND_MOVE, NULL);
    apply_random_rotation(s, res, mol, 0);
    orc_compute_new_energy(s, res, mol, ORC_KIND_MOVE, NULL);
    double p = compute_acceptance_probability(s, res, TYPE_ROTATION);
    int acc = orc_rand_uniform(s) <= p;
    if (acc) accept_molecule_move(s, s->counter[1]);
    else reject_molecule_move(s, res, mol, 0, 1);
    trace_fill(s, t, ORC_MV_ROTATE, res, mol, acc, p);
    return 0;
}

/* attempt_creation_move, creation.f90:30-76; accept :82-116 (no reservoir) */
int orc_attempt_creation_move(orc_system *s, int res, int mol, orc_step_trace *t)
{
    trace_none(t, res, mol);
    if (mol + 1 > NB_MAX_MOLECULE || mol >= s->res[res].cap)             /* check_molecule_index */
        return fail(s, "Trying to insert a molecule beyond the maximum allowed number of molecules");
    s->counter[2][0] += 1;
    orc_compute_old_energy(s, res, mol, ORC_KIND_CREATE, NULL);
    update_counts(s, res, +1);
    orc_save_single_mol_fourier_terms(s, res, mol);
    insert_and_orient_molecule(s, res, mol, 1);
    orc_compute_new_energy(s, res, mol, ORC_KIND_CREATE, NULL);
    double p = compute_acceptance_probability(s, res, TYPE_CREATION);
    int acc = orc_rand_uniform(s) <= p;
    if (acc) {
        accept_all_components(s);
        s->counter[2][0] += 1; s->counter[2][1] += 1;                    /* "counter%creations = counter%creations + 1": both slots */
    } else {
        reject_creation_move(s, res, mol);
    }
    trace_fill(s, t, ORC_MV_CREATE, res, mol, acc, p);
    return 0;
}

/* attempt_deletion_move, deletion.f90:29-77; accept :83-122; reject :128-149 */
int orc_attempt_deletion_move(orc_system *s, int res, int mol, orc_step_trace *t)
{
    trace_none(t, res, mol);
    residue_t *R = &s->res[res];
    if (R->nmol == 0) return 0;
    s->counter[3][0] += 1;
    orc_compute_old_energy(s, res, mol, ORC_KIND_DELETE, NULL);
    save_molecule_state(s, res, mol, 1, 1);
    int last = R->nmol - 1;
    remove_molecule(s, res, mol, last);
    update_counts(s, res, -1);
    orc_compute_new_energy(s, res, mol, ORC_KIND_DELETE, NULL);
    double p = compute_acceptance_probability(s, res, TYPE_DELETION);
    int acc = orc_rand_uniform(s) <= p;
    if (acc) {
        accept_all_components(s);
        s->counter[3][0] += 1; s->counter[3][1] += 1;                    /* both slots, deletion.f90:101 */
    } else {
        update_counts(s, res, +1);
        memcpy(&R->com[(size_t)mol * 3], s->saved_com, 3 * sizeof(double));
        memcpy(&R->off[(size_t)mol * R->natom * 3], s->saved_offset, (size_t)R->natom * 3 * sizeof(double));
        orc_restore_single_mol_fourier(s, res, mol);
    }
    trace_fill(s, t, ORC_MV_DELETE, res, mol, acc, p);
    return 0;
}

/* pick_random_residue_type, monte_carlo_utils.f90:140-174 */
static int pick_random_residue_type(orc_system *s)
{
    int n_active = 0, idx[MAXRES];
    for (int r = 0; r < s->nres; ++r) if (s->res[r].active) idx[n_active++] = r;
    if (n_active == 0) return -1;
    return idx[(int)(orc_rand_uniform(s) * n_active)];
}
/* pick_random_molecule_index :180-201 */
static int pick_random_molecule_index(orc_system *s, int count)
{
    if (count == 0) return -1;
    int m = (int)(orc_rand_uniform(s) * count) + 1;
    if (m > count) m = count;
    return m - 1;
}

/* attempt_swap_move, swapping.f90:34-105.  Reference behaviour kept: old energy
 * is deletion-style for type A, new energy is creation-style for type B on top
 * of an Ak that still contains A.  mol < 0 (no molecule of type A)
 * indexes molecule 0 in the reference (undefined); the oracle returns early. */
int orc_attempt_swap_move(orc_system *s, int res, int mol, orc_step_trace *t)
{
    trace_none(t, res, mol);
    /* pick_different_residue_type :165-197 */
    int res_bis = -1;
    for (int attempt = 0; attempt < 10; ++attempt) {
        int nt = pick_random_residue_type(s);
        if (nt != res) { res_bis = nt; break; }
    }
    if (res_bis == -1) return 0;
    if (s->res[res_bis].nmol == 0) return 0;
    if (mol < 0) return 0;
    residue_t *RA = &s->res[res], *RB = &s->res[res_bis];
    if (RB->nmol >= RB->cap) return fail(s, "swap: capacity exceeded");
    s->counter[4][0] += 1;
    int mol_bis = RB->nmol;
    orc_compute_old_energy(s, res, mol, ORC_KIND_DELETE, NULL);
    save_molecule_state(s, res, mol, 1, 1);
    int last = RA->nmol - 1;
    remove_molecule(s, res, mol, last);
    RA->nmol -= 1;
    RB->nmol += 1;
    memcpy(&RB->com[(size_t)mol_bis * 3], s->saved_com, 3 * sizeof(double));
    insert_and_orient_molecule(s, res_bis, mol_bis, 0);
    orc_compute_new_energy(s, res_bis, mol_bis, ORC_KIND_CREATE, NULL);
    double p = swap_acceptance_probability(s, res, res_bis);
    int acc = orc_rand_uniform(s) <= p;
    if (acc) {
        accept_all_components(s);
        s->counter[4][0] += 1; s->counter[4][1] += 1;
    } else {
        /* reject_swap_move :113-137 */
        RA->nmol += 1; RB->nmol -= 1;
        memcpy(&RA->com[(size_t)mol * 3], s->saved_com, 3 * sizeof(double));
        memcpy(&RA->off[(size_t)mol * RA->natom * 3], s->saved_offset, (size_t)RA->natom * 3 * sizeof(double));
        orc_restore_single_mol_fourier(s, res, mol);
    }
    trace_fill(s, t, ORC_MV_SWAP, res, mol, acc, p);
    return 0;
}

/* widom_trial, widom.f90:30-68; accumulate_widom_weight :74-92 */
int orc_widom_trial(orc_system *s, int res, int mol, orc_step_trace *t)
{
    trace_none(t, res, mol);
    if (mol + 1 > NB_MAX_MOLECULE || mol >= s->res[res].cap)
        return fail(s, "Trying to insert a molecule beyond the maximum allowed number of molecules");
    s->counter[5][0] += 1;
    orc_compute_old_energy(s, res, mol, ORC_KIND_CREATE, NULL);
    update_counts(s, res, +1);
    orc_save_single_mol_fourier_terms(s, res, mol);
    insert_and_orient_molecule(s, res, mol, 1);
    orc_compute_new_energy(s, res, mol, ORC_KIND_CREATE, NULL);
    reject_creation_move(s, res, mol);
    double deltaU = s->e_new[ORC_E_TOTAL] - s->e_old[ORC_E_TOTAL];
    double weight = exp(-deltaU * s->beta);
    int ok = weight > ERR;
    if (ok) { s->counter[5][1] += 1; s->widom_weight[res] += weight; }
    s->widom_sample[res] += 1;
    trace_fill(s, t, ORC_MV_WIDOM, res, mol, ok, weight);
    return 0;
}

/* body of monte_carlo_loop, monte_carlo.f90:50-99 */
int orc_monte_carlo_steps(orc_system *s, int64_t nsteps, orc_step_trace *trace)
{
    double cumul_translation = s->p_translation;
    double cumul_rotation = cumul_translation + s->p_rotation;
    double cumul_swap = cumul_rotation + s->p_swap;
    for (int64_t it = 0; it < nsteps; ++it) {
        orc_step_trace *t = trace ? &trace[it] : NULL;
        int res = pick_random_residue_type(s);
        if (res < 0) return fail(s, "no active residue");
        int mol = pick_random_molecule_index(s, s->res[res].nmol);
        double random_draw = orc_rand_uniform(s);
        int rc = 0;
        if (random_draw <= cumul_translation) rc = orc_attempt_translation_move(s, res, mol, t);
        else if (random_draw <= cumul_rotation) rc = orc_attempt_rotation_move(s, res, mol, t);
        else if (random_draw <= cumul_swap) rc = orc_attempt_swap_move(s, res, mol, t);
        else {
            if (s->p_insdel > 0) {
                if (orc_rand_uniform(s) <= PROB_CREATE_DELETE) rc = orc_attempt_creation_move(s, res, s->res[res].nmol, t);
                else rc = orc_attempt_deletion_move(s, res, mol, t);
            } else if (s->p_widom > 0) {
                rc = orc_widom_trial(s, res, s->res[res].nmol, t);
            } else trace_none(t, res, mol);
        }
        if (rc) return rc;
    }
    return 0;
}

int orc_get_counters(const orc_system *s, int64_t out[12])
{
    for (int i = 0; i < 6; ++i) { out[2 * i] = s->counter[i][0]; out[2 * i + 1] = s->counter[i][1]; }
    return 0;
}
int orc_get_widom(const orc_system *s, int res, double *sum_weight, int64_t *samples)
{
    *sum_weight = s->widom_weight[res]; *samples = s->widom_sample[res]; return 0;
}
int orc_reset_widom(orc_system *s)
{
    for (int r = 0; r < s->nres; ++r) { s->widom_weight[r] = 0.0; s->widom_sample[r] = 0; }
    return 0;
}

/* Widom batch: the body of widom_trial for insertion ids first_id..first_id+n-1,
 * each drawing its 5 uniforms (COM x3, angle, axis) from a counter-based
 * splitmix64 stream:  u(id,k) = top53(mix(seed + GOLDEN*(8*id + k + 1))). */
int orc_widom_batch(orc_system *s, int res, int64_t first_id, int64_t n, uint64_t seed,
                    double *dE_out, double *sum_w, int64_t *n_ok)
{
    double u[8];
    double acc = 0.0; int64_t ok = 0;
    const double *saved_stream = s->ustream; int64_t saved_n = s->ustream_n, saved_pos = s->ustream_pos;
    residue_t *R = &s->res[res];
    int slot = R->nmol;
    if (slot >= R->cap) return fail(s, "widom batch: capacity exceeded");
    /* Every insertion of a batch starts from the same template geometry (the
     * offsets of molecule 1 as they are when the batch starts).  The serial
     * reference rotates slot 1 in place when N = 0 (monte_carlo_utils.f90:563-568
     * with molecule_index = 1), which chains the orientations of successive
     * trials; a batch of independent insertions cannot, so the slot is put back. */
    double slot_com[3];
    double *slot_off = (double *)malloc(sizeof(double) * 3 * R->natom);
    memcpy(slot_com, &R->com[(size_t)slot * 3], sizeof slot_com);
    memcpy(slot_off, &R->off[(size_t)slot * R->natom * 3], sizeof(double) * 3 * R->natom);
    for (int64_t i = 0; i < n; ++i) {
        memcpy(&R->com[(size_t)slot * 3], slot_com, sizeof slot_com);
        memcpy(&R->off[(size_t)slot * R->natom * 3], slot_off, sizeof(double) * 3 * R->natom);
        uint64_t id = (uint64_t)(first_id + i);
        for (int k = 0; k < 8; ++k)
            u[k] = (double)(splitmix64_mix(seed + GOLDEN * (8u * id + (uint64_t)k + 1u)) >> 11) * 0x1.0p-53;
        orc_set_uniform_stream(s, u, 8);
        double w0 = s->widom_weight[res]; int64_t s0 = s->widom_sample[res];
        int64_t c0 = s->counter[5][0], c1 = s->counter[5][1];
        int rc = orc_widom_trial(s, res, s->res[res].nmol, NULL);
        if (rc) { s->ustream = saved_stream; s->ustream_n = saved_n; s->ustream_pos = saved_pos; free(slot_off); return rc; }
        double dE = s->e_new[ORC_E_TOTAL] - s->e_old[ORC_E_TOTAL];
        if (dE_out) dE_out[i] = dE;
        double weight = exp(-dE * s->beta);
        if (weight > ERR) { acc += weight; ok += 1; }
        /* batch statistics are returned, not accumulated in the walker */
        s->widom_weight[res] = w0; s->widom_sample[res] = s0; s->counter[5][0] = c0; s->counter[5][1] = c1;
    }
    memcpy(&R->com[(size_t)slot * 3], slot_com, sizeof slot_com);
    memcpy(&R->off[(size_t)slot * R->natom * 3], slot_off, sizeof(double) * 3 * R->natom);
    free(slot_off);
    s->ustream = saved_stream; s->ustream_n = saved_n; s->ustream_pos = saved_pos;
    *sum_w = acc; *n_ok = ok;
    return 0;
}
