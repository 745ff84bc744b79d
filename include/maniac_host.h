/*
 * maniac_host.h -- host-side Monte Carlo drivers on top of the C ABI (maniac_gpu.h).
 *
 * These play the part of MANIAC's Fortran move drivers when no Fortran compiler is
 * available (this build image has none): the loop body of monte_carlo_loop
 * (src/monte_carlo.f90:50-99) and attempt_{translation,rotation,creation,deletion}_move /
 * widom_trial (src/translation.f90, rotation.f90, creation.f90, deletion.f90, widom.f90)
 * with the acceptance rules of src/monte_carlo_utils.f90:204-255, written in C++ and
 * calling ONLY the exported mgpu_* energy entry points -- exactly what the Fortran
 * drivers do through the ISO_C_BINDING shim.  The host owns the
 * coordinates (AoS like coord%com / coord%offset), draws the random numbers, proposes,
 * and decides; the GPU computes the energies.  All walkers advance in lock step, one
 * mgpu_trial_batch + one mgpu_commit_batch per MC step.
 *
 * RNG contract = the engine's (xoshiro256** per walker, splitmix64(seed + 104729*w)),
 * so a host-driven run, a device-resident mgpu_sweep and the CPU oracle produce the
 * same trajectory.
 */
#ifndef MANIAC_HOST_H
#define MANIAC_HOST_H

#include "maniac_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mhost_sim mhost_sim;

/* Builds host mirrors for every walker from the same description mgpu_init received
 * (call mgpu_init first). */
mhost_sim *mhost_create(const mgpu_system *sys, uint64_t seed);
void mhost_destroy(mhost_sim *sim);
const char *mhost_last_error(void);

/* n_steps MC steps for all walkers.  trace (optional, [n_steps]) records walker
 * `trace_walker`.  Returns 0 or an error (message in mhost_last_error). */
int mhost_run(mhost_sim *sim, int64_t n_steps, int32_t trace_walker, mgpu_step_trace *trace);

/* adjust_move_step_sizes (src/monte_carlo_utils.f90:98-134) for every walker, from its own counters */
int mhost_adjust_move_step_sizes(mhost_sim *sim);
int mhost_get_step_sizes(const mhost_sim *sim, int32_t walker, double out[2]);
/* thermo%chemical_potential(res) of one walker (isotherm points) */
int mhost_set_chemical_potential(mhost_sim *sim, int32_t walker, int32_t res, double mu);
int mhost_get_count(const mhost_sim *sim, int32_t walker, int32_t res);
int mhost_get_energy(const mhost_sim *sim, int32_t walker, double out[6]);
int mhost_get_counters(const mhost_sim *sim, int32_t walker, int64_t out[12]);
int mhost_get_molecule(const mhost_sim *sim, int32_t walker, int32_t res, int32_t mol, double com[3], double *offset);
/* bytes moved host->device / device->host by the energy calls since creation, and trials issued */
int mhost_get_traffic(const mhost_sim *sim, int64_t *h2d_bytes, int64_t *d2h_bytes, int64_t *trials);

#ifdef __cplusplus
}
#endif
#endif
