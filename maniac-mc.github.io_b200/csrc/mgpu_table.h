// mgpu_table.h -- host-side builder of the real-space Coulomb table (see mgpu_internal.h).
#pragma once
#include <vector>
#include <cuda_runtime.h>
#include "mgpu_internal.h"

// Choose the octave range [2^emin, 2^(emin+noct)) of s = r^2 that covers r_lo..r_hi
// (at most MGPU_TAB_MAXOCT octaves) and fit one degree-6 polynomial per interval.
// tab is [interval][MGPU_TAB_ROW] (see mgpu_internal.h for the row format).
void mgpu_build_coulomb_table(double alpha, double r_lo, double r_hi, int *emin, int *noct, std::vector<double> &tab);

// Evaluate the table exactly as the device does (same integer indexing, same Horner order).
// Returns false when s is outside the tabulated range.
bool mgpu_eval_coulomb_table(const std::vector<double> &tab, int emin, int noct, double s, double *g);
