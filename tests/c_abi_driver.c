/*
 * c_abi_driver.c -- a plain-C host that calls libmaniac_gpu exactly the way the Fortran shim does
 * (integration/fortran/maniac_gpu_iface.f90 + energy_glue_gpu.f90): `mgpu_system` / `mgpu_residue` built in
 * caller memory and passed BY POINTER, scalars BY VALUE (the interface block's `value` attribute), residue and
 * molecule ids kept 1-based on this side and converted (-1) at the call like the shim does, the trial geometry
 * handed over as com(3) + off(3, MGPU_MAX_SITES) in column-major order (= offset[MGPU_MAX_SITES][3] here),
 * energies received as real(c_double) :: buf(6).  No Fortran compiler exists in the build image, so this
 * program is what exercises the argument conventions of the drop-in boundary; tests/test_c_abi_driver.py builds
 * it with gcc, feeds it a system + a list of trial moves and checks every number against the CPU oracle.
 *
 * The sequence per trial is the one the reference's move drivers run
 * (src/translation.f90:21-97, creation.f90:30-116, deletion.f90:29-122 through
 * src/monte_carlo_utils.f90:300-423):
 *     compute_old_energy(res, mol, is_creation, is_deletion)   -> mgpu_old_energy
 *     compute_new_energy(res, mol, is_creation, is_deletion)   -> mgpu_new_energy (leaves a pending trial)
 *     accept_* / reject_*                                      -> mgpu_commit / mgpu_rollback
 * and, at the start and the end, update_system_energy (src/energy_utils.f90:22-39) -> mgpu_total_energy.
 *
 * Input (binary, native endianness) -- written by the test:
 *   int32 magic 0x4D475055 | double matrix[9], lo[3] | int32 nres, ntypes |
 *   per residue: int32 natom, is_active, nmol, capacity; double mass, fugacity, chemical_potential;
 *                double charges[natom]; int32 types[natom]; double com[nmol][3]; double offset[nmol][natom][3] |
 *   double epsilon[ntypes^2], sigma[ntypes^2] |
 *   double temperature, ewald_tolerance, real_space_cutoff, translation_step, rotation_step_angle,
 *          p_translation, p_rotation, p_swap, p_insertion_deletion, p_widom |
 *   int32 ntrials | per trial: int32 kind, res (1-based), mol (1-based), accept; double com[3], off[MGPU_MAX_SITES][3]
 * Output: text lines on stdout (%.17g), see main().
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "maniac_gpu.h"

#define WALKER0 0 /* the Fortran host drives one walker (energy_glue_gpu.f90) */

static FILE *in;

static void rd(void *p, size_t sz, size_t n)
{
    if (fread(p, sz, n, in) != n) { fprintf(stderr, "c_abi_driver: short read\n"); exit(2); }
}

/* mgpu_check of the shim: non-zero -> abort_run(msg, code) (src/output_utils.f90:581-605) */
static void mgpu_check(int rc, const char *what)
{
    if (rc == 0) return;
    printf("ABORT %s: %s\n", what, mgpu_last_error());
    fflush(stdout);
    exit(1);
}

static void print6(const char *tag, int i, const double e[6])
{
    printf("%s %d %.17g %.17g %.17g %.17g %.17g %.17g\n", tag, i, e[0], e[1], e[2], e[3], e[4], e[5]);
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: c_abi_driver input.bin\n"); return 2; }
    in = fopen(argv[1], "rb");
    if (!in) { perror(argv[1]); return 2; }
    int32_t magic = 0;
    rd(&magic, 4, 1);
    if (magic != 0x4D475055) { fprintf(stderr, "c_abi_driver: bad magic\n"); return 2; }

    mgpu_system sys;
    memset(&sys, 0, sizeof sys);
    rd(sys.matrix, 8, 9);
    rd(sys.lo, 8, 3);
    rd(&sys.nres, 4, 1);
    rd(&sys.ntypes, 4, 1);
    mgpu_residue *res = (mgpu_residue *)calloc((size_t)sys.nres, sizeof *res);
    for (int r = 0; r < sys.nres; ++r) {
        int32_t h[4];
        double d[3];
        rd(h, 4, 4);
        rd(d, 8, 3);
        res[r].natom = h[0]; res[r].is_active = h[1]; res[r].nmol = h[2]; res[r].capacity = h[3];
        res[r].mass = d[0]; res[r].fugacity = d[1]; res[r].chemical_potential = d[2];
        double *q = (double *)malloc(sizeof(double) * (size_t)h[0]);
        int32_t *t = (int32_t *)malloc(sizeof(int32_t) * (size_t)h[0]);
        double *com = (double *)malloc(sizeof(double) * 3 * (size_t)(h[2] ? h[2] : 1));
        double *off = (double *)malloc(sizeof(double) * 3 * (size_t)h[0] * (size_t)(h[2] ? h[2] : 1));
        rd(q, 8, (size_t)h[0]);
        rd(t, 4, (size_t)h[0]);
        rd(com, 8, 3 * (size_t)h[2]);
        rd(off, 8, 3 * (size_t)h[0] * (size_t)h[2]);
        res[r].charges = q; res[r].types = t; res[r].com = com; res[r].offset = off;
    }
    sys.residues = res;
    const size_t nt2 = (size_t)sys.ntypes * (size_t)sys.ntypes;
    double *eps = (double *)malloc(8 * nt2), *sig = (double *)malloc(8 * nt2);
    rd(eps, 8, nt2);
    rd(sig, 8, nt2);
    sys.epsilon = eps; sys.sigma = sig;
    double sc[10];
    rd(sc, 8, 10);
    sys.temperature = sc[0]; sys.ewald_tolerance = sc[1]; sys.real_space_cutoff = sc[2];
    sys.translation_step = sc[3]; sys.rotation_step_angle = sc[4];
    sys.p_translation = sc[5]; sys.p_rotation = sc[6]; sys.p_swap = sc[7]; sys.p_insertion_deletion = sc[8]; sys.p_widom = sc[9];
    sys.n_walkers = 1;
    sys.device = 0;

    mgpu_check(mgpu_init(&sys), "mgpu_init");                       /* struct by pointer */

    double alpha, rc;
    int32_t kmax[3], nk;
    mgpu_check(mgpu_get_ewald(&alpha, kmax, &nk, &rc), "mgpu_get_ewald");
    printf("EWALD %.17g %d %d %d %d %.17g\n", alpha, kmax[0], kmax[1], kmax[2], nk, rc);

    double buf[6];
    mgpu_check(mgpu_total_energy(WALKER0, buf), "update_system_energy");   /* scalar by value, buf(6) by reference */
    print6("TOTAL", 0, buf);

    int32_t ntrials = 0;
    rd(&ntrials, 4, 1);
    for (int i = 0; i < ntrials; ++i) {
        int32_t h[4];
        double com[3], off[MGPU_MAX_SITES][3];                      /* = real(c_double) :: off(3, MGPU_MAX_SITES) */
        rd(h, 4, 4);
        rd(com, 8, 3);
        rd(off, 8, 3 * MGPU_MAX_SITES);
        const int32_t kind = h[0], res_type = h[1], mol_index = h[2], accept = h[3];   /* 1-based, as the drivers hold them */
        double e_old[6], e_new[6];
        mgpu_check(mgpu_old_energy(WALKER0, res_type - 1, mol_index - 1, kind, e_old), "compute_old_energy");
        mgpu_check(mgpu_new_energy(WALKER0, res_type - 1, mol_index - 1, kind,
                                   kind == MGPU_KIND_DELETE ? NULL : com, kind == MGPU_KIND_DELETE ? NULL : &off[0][0], e_new),
                   "compute_new_energy");
        print6("OLD", i, e_old);
        print6("NEW", i, e_new);
        if (accept) mgpu_check(mgpu_commit(WALKER0), "accept");
        else mgpu_check(mgpu_rollback(WALKER0), "reject");
        mgpu_check(mgpu_get_energy(WALKER0, buf), "energy");
        print6("RUN", i, buf);
        int32_t cnt = -1;
        mgpu_check(mgpu_get_count(WALKER0, res_type - 1, &cnt), "count");
        printf("COUNT %d %d\n", i, cnt);
    }
    mgpu_check(mgpu_total_energy(WALKER0, buf), "update_system_energy");
    print6("TOTAL", 1, buf);

    /* error convention: an out-of-range request must come back non-zero with a message, not crash */
    const int rc_bad = mgpu_old_energy(WALKER0, sys.nres + 3, 0, MGPU_KIND_MOVE, buf);
    printf("BADRES %d %s\n", rc_bad != 0, rc_bad ? mgpu_last_error() : "");

    mgpu_finalize();
    printf("OK\n");
    return 0;
}
