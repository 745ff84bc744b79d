// mgpu_kernels.cuh -- sm_100a kernels of the MANIAC per-trial-move energy path.
//
// K1  pair_sums_cta     fused real-space LJ + erfc-Coulomb of a trial molecule (old and
//                       new geometry in one pass over the targets) vs the host framework
//                       and the walker's guests            (pairwise_energy_utils.f90:21-179,
//                                                           geometry_utils.f90:210-284)
// K2  kspace_cta        1-D phase tables, dS(k), S_trial = S + dS, sum_k ffW |S_trial|^2
//                                                          (ewald_phase.f90:205-312,
//                                                           ewald_energy.f90:64-164)
// K3  k_widom_batch     thread-per-insertion real space + warp-per-insertion k space
// K4  k_total_energy    full recompute (energy_utils.f90:22-134, ewald_energy.f90:20-58)
//     k_sweep           device-resident move drivers + Metropolis (translation.f90,
//                       rotation.f90, creation.f90, deletion.f90, widom.f90,
//                       monte_carlo.f90:50-99), one CTA per walker
//
// All arithmetic is IEEE binary64.  No tensor cores: nothing here is a dense contraction.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "mgpu_internal.h"

__constant__ DevSys c_sys;

#define MGPU_BLOCK 256
#define MGPU_WARPS (MGPU_BLOCK / 32)

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

// Block-wide sum of NV values per thread; result valid in every thread's out[].
// red: shared scratch of NV*MGPU_WARPS doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *red)
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
template <bool TRI>
__device__ __forceinline__ double min_image_r2(double dx, double dy, double dz)
{
    if (!TRI) {
        // delta_d = modulo(delta_d + L/2, L) - L/2 restated as delta - L*rint(delta/L): the same
        // image except on the exact tie |delta| = L/2, where both images have the same length.
        dx = fma(-c_sys.L[0], rint(dx * c_sys.invL[0]), dx);
        dy = fma(-c_sys.L[1], rint(dy * c_sys.invL[1]), dy);
        dz = fma(-c_sys.L[2], rint(dz * c_sys.invL[2]), dz);
        return fma(dx, dx, fma(dy, dy, dz * dz));
    } else {
        // 27-image search with the COLUMNS of matrix as cell vectors, exactly like the reference
        double best = 1.7976931348623157e308;
#pragma unroll
        for (int sx = -1; sx <= 1; ++sx)
#pragma unroll
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
}

// One atom pair: LJ (pairwise_lj_energy, pairwise_energy_utils.f90:95-138) and erfc
// Coulomb (pairwise_coulomb_energy :143-179, NO cutoff on the Coulomb term).
// eps4 = 4*epsilon; qq = q_i*q_j or 0 when either |q| < 1e-10.
// Work counters for the roofline accounting: pairs evaluated, LJ terms inside the
// cutoff, erfc-Coulomb terms.  Integer adds on the otherwise idle ALU pipe.
struct PairCount { unsigned geom, lj, coul; };

__device__ __forceinline__ void pair_terms(double r2, double eps4, double sig, double qq, bool charged,
                                           double &e_lj, double &e_coul, PairCount &pc)
{
    pc.geom += 1u;
    const double rinv = rsqrt(r2);
    const double r = r2 * rinv;
    if (r < MGPU_ERR_TOL || !(r2 > 0.0)) {          // overlap sentinel (r = 0 gives rinv = inf)
        if (0.0 < c_sys.rc) e_lj += c_sys.overlap;
        if (charged) e_coul += c_sys.overlap;
        return;
    }
    if (r < c_sys.rc) {
        const double s = sig * rinv;
        const double s2 = s * s;
        const double s6 = s2 * s2 * s2;
        e_lj += eps4 * (s6 * s6 - s6);
        pc.lj += (eps4 != 0.0);
    }
    if (charged) { e_coul += qq * erfc(c_sys.alpha * r) * rinv; pc.coul += 1u; }
}
__device__ __forceinline__ void flush_pair_count(const PairCount &pc)
{
    const unsigned g = __reduce_add_sync(0xffffffffu, pc.geom), l = __reduce_add_sync(0xffffffffu, pc.lj),
                   c = __reduce_add_sync(0xffffffffu, pc.coul);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(c_sys.pair_count + 0, (unsigned long long)g);
        atomicAdd(c_sys.pair_count + 1, (unsigned long long)l);
        atomicAdd(c_sys.pair_count + 2, (unsigned long long)c);
    }
}

// ------------------------------------------------------------------------------------
// Probe: the trial molecule, staged in shared memory
// ------------------------------------------------------------------------------------
struct Probe {
    int32_t kind, res, mol, na;
    int32_t has_old, has_new;
    int32_t excl_res, excl_mol;              // target slot to skip (the molecule itself)
    int32_t order_res, order_mol;            // >= 0: only targets with (res,mol) > this (ordering check)
    double q[MGPU_MAX_SITES];
    int32_t type[MGPU_MAX_SITES];
    double po[MGPU_MAX_SITES][3];            // old atom positions (com + offset)
    double pn[MGPU_MAX_SITES][3];            // new atom positions
};

// Accumulate the probe's pair energies against one target atom.
template <bool TRI>
__device__ __forceinline__ void probe_vs_atom(const Probe &P, const double *s_eps4, const double *s_sig,
                                              double tx, double ty, double tz, double tq, int ttype,
                                              double (&acc)[4], PairCount &pc)
{
    const bool tcharged = fabs(tq) >= MGPU_ERR_TOL;
    const int nt = c_sys.ntypes;
    for (int a = 0; a < P.na; ++a) {
        const double qa = P.q[a];
        const bool charged = tcharged && (fabs(qa) >= MGPU_ERR_TOL);
        const double qq = qa * tq;
        const int ti = P.type[a] * nt + ttype;
        const double eps4 = s_eps4[ti], sig = s_sig[ti];
        if (eps4 == 0.0 && !charged) continue;
        if (P.has_old) {
            double r2 = min_image_r2<TRI>(tx - P.po[a][0], ty - P.po[a][1], tz - P.po[a][2]);
            pair_terms(r2, eps4, sig, qq, charged, acc[0], acc[1], pc);
        }
        if (P.has_new) {
            double r2 = min_image_r2<TRI>(tx - P.pn[a][0], ty - P.pn[a][1], tz - P.pn[a][2]);
            pair_terms(r2, eps4, sig, qq, charged, acc[2], acc[3], pc);
        }
    }
}

// K1: block-cooperative pair sums of the probe against host atoms + the walker's guests.
// Returns {lj_old, coul_old(e^2/A), lj_new, coul_new(e^2/A)} in every thread.
template <bool TRI>
__device__ void pair_sums_cta(const Probe &P, int w, const int32_t *s_count, const double *s_eps4,
                              const double *s_sig, double *red, double (&out)[4])
{
    double acc[4] = { 0.0, 0.0, 0.0, 0.0 };
    PairCount pc = { 0u, 0u, 0u };
    // host framework: one lane per host atom, coalesced 32-byte loads
    if (P.order_res < 0 || true) {
        const double4 *__restrict__ hx = c_sys.host_xyzq;
        const int32_t *__restrict__ ht = c_sys.host_type;
        // host residues have lower or higher res id than the probe: the ordering check of
        // pairwise_energy_for_molecule (:60-62) only matters in the full-energy pass, where
        // host atoms are handled by host_order_ok below.
        for (int j = threadIdx.x; j < c_sys.n_host; j += blockDim.x) {
            const double4 t = hx[j];
            probe_vs_atom<TRI>(P, s_eps4, s_sig, t.x, t.y, t.z, t.w, ht[j], acc, pc);
        }
    }
    // guests of this walker: one lane per molecule
    const double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride;
    for (int g = 0; g < c_sys.nres; ++g) {
        if (!c_sys.active[g]) continue;
        const int cap = c_sys.cap[g], na_g = c_sys.natom[g], n = s_count[g];
        const double *com = wc + c_sys.goff[g];
        const double *off = com + 3 * (int64_t)cap;
        for (int m = threadIdx.x; m < n; m += blockDim.x) {
            if (g == P.excl_res && m == P.excl_mol) continue;
            if (P.order_res >= 0 && (g < P.order_res || (g == P.order_res && m <= P.order_mol))) continue;
            const double cx = com[m], cy = com[cap + m], cz = com[2 * cap + m];
            for (int b = 0; b < na_g; ++b) {
                const double *ob = off + (int64_t)b * 3 * cap;
                const double tx = cx + ob[m], ty = cy + ob[cap + m], tz = cz + ob[2 * cap + m];
                probe_vs_atom<TRI>(P, s_eps4, s_sig, tx, ty, tz, c_sys.charge[g][b], c_sys.type[g][b], acc, pc);
            }
        }
    }
    flush_pair_count(pc);
    block_sum<4>(acc, red);
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = acc[i];
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
// compute_atom_phase (ewald_phase.f90:255-280) + compute_phase_factor (:286-312).
// tab layout: [atom][dim][k] (stride KW = kmax_max+1), as double2 {cos, sin}.
__device__ __forceinline__ void fill_phase_tables(double2 *tab, const double (*pos)[3], int n, int tid0, int nthreads)
{
    const int KW = c_sys.kmax_max + 1;
    const int total = n * 3 * KW;
    for (int e = tid0; e < total; e += nthreads) {
        const int k = e % KW, d = (e / KW) % 3, a = e / (3 * KW);
        if (k > c_sys.kmax[d]) continue;
        double ph = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) ph = __dadd_rn(ph, __dmul_rn(c_sys.Hinv[j * 3 + d], pos[a][j]));
        ph = c_sys.twopi * ph;
        double sn, cs;
        sincos((double)k * ph, &sn, &cs);
        tab[e] = make_double2(cs, sn);
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

// S_trial(k) = S(k) + dS(k) and E = sum_k ffW |S_trial|^2 * EPS0_INV_real * TWOPI / V.
// tab_old / tab_new: phase tables of the probe's old / new geometry (shared memory).
// S_out may be NULL (Widom: nothing is stored).
__device__ double kspace_cta(const Probe &P, const double2 *tab_old, const double2 *tab_new,
                             const double *__restrict__ S_in, double *__restrict__ S_out, double *red)
{
    const int nk = c_sys.nk;
    double part[1] = { 0.0 };
    for (int i = threadIdx.x; i < nk; i += blockDim.x) {
        const int kx = c_sys.kx[i], ky = c_sys.ky[i], kz = c_sys.kz[i];
        double sr = 0.0, si = 0.0;
        for (int a = 0; a < P.na; ++a) {
            const double q = P.q[a];
            cplx pn = { 0.0, 0.0 }, po = { 0.0, 0.0 };
            if (P.has_new) pn = phase_product(tab_new, a, kx, ky, kz);
            if (P.has_old) po = phase_product(tab_old, a, kx, ky, kz);
            if (P.kind == MGPU_KIND_CREATE) { sr += q * pn.re; si += q * pn.im; }
            else if (P.kind == MGPU_KIND_DELETE) { sr += q * po.re; si += q * po.im; }
            else { sr += q * (pn.re - po.re); si += q * (pn.im - po.im); }
        }
        double re = S_in[i], im = S_in[nk + i];
        if (P.kind == MGPU_KIND_DELETE) { re -= sr; im -= si; } else { re += sr; im += si; }
        if (S_out) { S_out[i] = re; S_out[nk + i] = im; }
        part[0] += c_sys.ffW[i] * (re * re + im * im);
    }
    block_sum<1>(part, red);
    return part[0] * c_sys.eps0_inv_real * c_sys.twopi / c_sys.volume;
}

// reciprocal_ewald_energy (ewald_energy.f90:139-164) of a stored S(k)
__device__ double recip_energy_cta(const double *__restrict__ S, double *red)
{
    const int nk = c_sys.nk;
    double part[1] = { 0.0 };
    for (int i = threadIdx.x; i < nk; i += blockDim.x) {
        const double re = S[i], im = S[nk + i];
        part[0] += c_sys.ffW[i] * (re * re + im * im);
    }
    block_sum<1>(part, red);
    return part[0] * c_sys.eps0_inv_real * c_sys.twopi / c_sys.volume;
}

// ------------------------------------------------------------------------------------
// shared-memory carve-up used by the trial / sweep / total kernels
// ------------------------------------------------------------------------------------
struct SmemLayout {
    Probe   *probe;
    double  *red;        // 4*MGPU_WARPS
    int32_t *count;      // MGPU_MAX_RES
    double  *eps4, *sig; // ntypes^2 each
    double2 *tab_old, *tab_new;
};
__host__ __device__ inline size_t smem_bytes(int ntypes, int kmax_max, int natom_max)
{
    size_t b = 0;
    b += (sizeof(Probe) + 15) & ~size_t(15);
    b += sizeof(double) * 4 * MGPU_WARPS;
    b += sizeof(double) * 8;                                  // count (int32 x MGPU_MAX_RES) + pad
    b += sizeof(double) * 2 * (size_t)ntypes * ntypes;
    b += sizeof(double2) * 2 * (size_t)natom_max * 3 * (kmax_max + 1);
    return b + 64;
}
__device__ __forceinline__ SmemLayout carve(unsigned char *base, int natom_max)
{
    SmemLayout s;
    size_t o = 0;
    s.probe = reinterpret_cast<Probe *>(base + o); o += (sizeof(Probe) + 15) & ~size_t(15);
    s.red = reinterpret_cast<double *>(base + o); o += sizeof(double) * 4 * MGPU_WARPS;
    s.count = reinterpret_cast<int32_t *>(base + o); o += sizeof(double) * 8;
    const int nt2 = c_sys.ntypes * c_sys.ntypes;
    s.eps4 = reinterpret_cast<double *>(base + o); o += sizeof(double) * nt2;
    s.sig = reinterpret_cast<double *>(base + o); o += sizeof(double) * nt2;
    o = (o + 15) & ~size_t(15);
    const size_t tab = (size_t)natom_max * 3 * (c_sys.kmax_max + 1);
    s.tab_old = reinterpret_cast<double2 *>(base + o); o += sizeof(double2) * tab;
    s.tab_new = reinterpret_cast<double2 *>(base + o);
    return s;
}
__device__ __forceinline__ void stage_common(const SmemLayout &s, int w)
{
    const int nt2 = c_sys.ntypes * c_sys.ntypes;
    for (int i = threadIdx.x; i < nt2; i += blockDim.x) { s.eps4[i] = 4.0 * c_sys.eps[i]; s.sig[i] = c_sys.sig[i]; }
    if (threadIdx.x < MGPU_MAX_RES) s.count[threadIdx.x] = c_sys.count[(int64_t)w * MGPU_MAX_RES + threadIdx.x];
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

// Evaluate one trial held in s.probe (positions staged): fills e_old / e_new (6 each) in
// thread 0's registers and writes S_trial into the walker's non-committed buffer.
// compute_old_energy / compute_new_energy, monte_carlo_utils.f90:300-423.
template <bool TRI>
__device__ void evaluate_trial(const SmemLayout &s, int w, int natom_max, double e_old[6], double e_new[6])
{
    const Probe &P = *s.probe;
    double ps[4];
    pair_sums_cta<TRI>(P, w, s.count, s.eps4, s.sig, s.red, ps);
    // phase tables
    if (P.has_old) fill_phase_tables(s.tab_old, P.po, P.na, threadIdx.x, blockDim.x);
    if (P.has_new) fill_phase_tables(s.tab_new, P.pn, P.na, threadIdx.x, blockDim.x);
    __syncthreads();
    const int cur = c_sys.cur[w];
    const double *S_in = c_sys.S + ((int64_t)w * 2 + cur) * 2 * c_sys.nk;
    double *S_out = c_sys.S + ((int64_t)w * 2 + (cur ^ 1)) * 2 * c_sys.nk;
    const double recip_new = kspace_cta(P, s.tab_old, s.tab_new, S_in, S_out, s.red);
    const double recip_cur = c_sys.energy[(int64_t)w * 6 + MGPU_E_RECIP];
#pragma unroll
    for (int i = 0; i < 6; ++i) { e_old[i] = 0.0; e_new[i] = 0.0; }
    e_old[MGPU_E_RECIP] = recip_cur;
    e_new[MGPU_E_RECIP] = recip_new;
    if (P.has_old) {
        e_old[MGPU_E_NON_COULOMB] = ps[0];
        e_old[MGPU_E_COULOMB] = ps[1] * c_sys.eps0_inv_real;
    }
    if (P.has_new) {
        e_new[MGPU_E_NON_COULOMB] = ps[2];
        e_new[MGPU_E_COULOMB] = ps[3] * c_sys.eps0_inv_real;
    }
    if (P.kind == MGPU_KIND_CREATE) {
        e_new[MGPU_E_SELF] = c_sys.e_self[P.res];
        e_new[MGPU_E_INTRA] = intra_energy<TRI>(P.res, P.pn);
    } else if (P.kind == MGPU_KIND_DELETE) {
        e_old[MGPU_E_SELF] = c_sys.e_self[P.res];
        e_old[MGPU_E_INTRA] = intra_energy<TRI>(P.res, P.po);
    }
    e_old[MGPU_E_TOTAL] = e_old[0] + e_old[1] + e_old[2] + e_old[3] + e_old[4];
    e_new[MGPU_E_TOTAL] = e_new[0] + e_new[1] + e_new[2] + e_new[3] + e_new[4];
}

// Stage the probe for a trial of `kind` on (res, mol) with new geometry (com, off).
__device__ __forceinline__ void stage_probe(const SmemLayout &s, int w, int kind, int res, int mol,
                                            const double *com, const double (*off)[3])
{
    Probe &P = *s.probe;
    const int na = c_sys.natom[res];
    if (threadIdx.x == 0) {
        P.kind = kind; P.res = res; P.mol = mol; P.na = na;
        P.has_old = (kind != MGPU_KIND_CREATE);
        P.has_new = (kind != MGPU_KIND_DELETE);
        P.excl_res = res; P.excl_mol = mol;
        P.order_res = -1; P.order_mol = -1;
    }
    if (threadIdx.x < na) {
        const int a = threadIdx.x;
        P.q[a] = c_sys.charge[res][a];
        P.type[a] = c_sys.type[res][a];
        if (kind != MGPU_KIND_CREATE) load_positions(w, res, mol, P.po, a);
        if (kind != MGPU_KIND_DELETE) {
#pragma unroll
            for (int d = 0; d < 3; ++d) P.pn[a][d] = com[d] + off[a][d];
        }
    }
}

// Apply an accepted trial to the walker (block-cooperative).
// accept_molecule_move / accept_creation_move / accept_deletion_move + remove_molecule +
// update_counts (monte_carlo_utils.f90:429-442,642-672; creation.f90:82-116; deletion.f90:83-122).
__device__ void commit_trial(int w, int kind, int res, int mol, const double *com, const double (*off)[3],
                             const double e_old[6], const double e_new[6])
{
    double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride + c_sys.goff[res];
    const int cap = c_sys.cap[res], na = c_sys.natom[res];
    double *offs = wc + 3 * (int64_t)cap;
    const int n = c_sys.count[(int64_t)w * MGPU_MAX_RES + res];
    const int tid = threadIdx.x;
    if (kind == MGPU_KIND_DELETE) {
        const int last = n - 1;
        if (mol != last) {
            if (tid < 3) wc[tid * cap + mol] = wc[tid * cap + last];
            for (int e = tid; e < na * 3; e += blockDim.x) offs[(int64_t)e * cap + mol] = offs[(int64_t)e * cap + last];
        }
    } else {
        if (tid < 3) wc[tid * cap + mol] = com[tid];
        for (int e = tid; e < na * 3; e += blockDim.x) offs[(int64_t)e * cap + mol] = off[e / 3][e % 3];
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

// Coordinates of a slot only (no counts / energies / S).  The reference writes a created
// molecule into slot N+1 before the accept test and does not undo it on reject
// (reject_creation_move, monte_carlo_utils.f90:579-591); that is observable when N = 0,
// because slot 1 is the geometry template of the next insertion (:563-564).
__device__ __forceinline__ void write_slot(int w, int res, int mol, const double *com, const double (*off)[3])
{
    double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride + c_sys.goff[res];
    const int cap = c_sys.cap[res], na = c_sys.natom[res];
    double *offs = wc + 3 * (int64_t)cap;
    if (threadIdx.x < 3) wc[threadIdx.x * cap + mol] = com[threadIdx.x];
    for (int e = threadIdx.x; e < na * 3; e += blockDim.x) offs[(int64_t)e * cap + mol] = off[e / 3][e % 3];
}

// ------------------------------------------------------------------------------------
// host-driven trial batch: one CTA per task
// ------------------------------------------------------------------------------------
struct TaskArrays {
    const int4 *meta;      // [n] {walker, res, mol, kind}
    const double *com;     // [n][3]
    const double *off;     // [n][MGPU_MAX_SITES][3]
    double *out;           // [n][12] e_old, e_new
};

template <bool TRI>
__global__ void __launch_bounds__(MGPU_BLOCK) k_trial(TaskArrays T, int natom_max)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout s = carve(smem, natom_max);
    const int t = blockIdx.x;
    const int4 meta = T.meta[t];
    const int w = meta.x, res = meta.y, mol = meta.z, kind = meta.w;
    stage_common(s, w);
    const double *com = T.com + (int64_t)t * 3;
    const double(*off)[3] = reinterpret_cast<const double(*)[3]>(T.off + (int64_t)t * MGPU_MAX_SITES * 3);
    stage_probe(s, w, kind, res, mol, com, off);
    __syncthreads();
    double e_old[6], e_new[6];
    evaluate_trial<TRI>(s, w, natom_max, e_old, e_new);
    // record the pending trial
    MgpuTrial *tr = c_sys.trial + w;
    if (threadIdx.x == 0) {
        tr->active = 1; tr->kind = kind; tr->res = res; tr->mol = mol;
        for (int d = 0; d < 3; ++d) tr->com[d] = (kind == MGPU_KIND_DELETE) ? 0.0 : com[d];
        for (int i = 0; i < 6; ++i) { tr->e_old[i] = e_old[i]; tr->e_new[i] = e_new[i]; T.out[(int64_t)t * 12 + i] = e_old[i]; T.out[(int64_t)t * 12 + 6 + i] = e_new[i]; }
    }
    if (kind != MGPU_KIND_DELETE)
        for (int e = threadIdx.x; e < c_sys.natom[res] * 3; e += blockDim.x) tr->off[e / 3][e % 3] = off[e / 3][e % 3];
}

__global__ void __launch_bounds__(128) k_commit(const int32_t *walker, const int32_t *accept, int32_t *err)
{
    const int t = blockIdx.x;
    const int w = walker[t];
    MgpuTrial *tr = c_sys.trial + w;
    if (!tr->active) { if (threadIdx.x == 0) atomicExch(err, 1); return; }
    if (accept[t]) commit_trial(w, tr->kind, tr->res, tr->mol, tr->com, tr->off, tr->e_old, tr->e_new);
    else if (tr->kind == MGPU_KIND_CREATE && tr->mol == 0) write_slot(w, tr->res, 0, tr->com, tr->off);
    __syncthreads();
    if (threadIdx.x == 0) tr->active = 0;
}

// pair energy of one guest molecule with optional explicit geometry and ordering check
template <bool TRI>
__global__ void __launch_bounds__(MGPU_BLOCK) k_pair_molecule(int w, int res, int mol, int skip_ordering,
                                                              const double *geom /* com[3] + off[na][3] or NULL */,
                                                              double *out2, int natom_max)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout s = carve(smem, natom_max);
    stage_common(s, w);
    Probe &P = *s.probe;
    const int na = c_sys.natom[res];
    if (threadIdx.x == 0) {
        P.kind = MGPU_KIND_DELETE; P.res = res; P.mol = mol; P.na = na; P.has_old = 1; P.has_new = 0;
        P.excl_res = res; P.excl_mol = mol;
        P.order_res = skip_ordering ? -1 : res; P.order_mol = skip_ordering ? -1 : mol;
    }
    if (threadIdx.x < na) {
        const int a = threadIdx.x;
        P.q[a] = c_sys.charge[res][a]; P.type[a] = c_sys.type[res][a];
        if (geom) { for (int d = 0; d < 3; ++d) P.po[a][d] = geom[d] + geom[3 + a * 3 + d]; }
        else load_positions(w, res, mol, P.po, a);
    }
    __syncthreads();
    double ps[4];
    pair_sums_cta<TRI>(P, w, s.count, s.eps4, s.sig, s.red, ps);
    if (threadIdx.x == 0) { out2[0] = ps[0]; out2[1] = ps[1] * c_sys.eps0_inv_real; }
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
    extern __shared__ __align__(16) unsigned char smem[];
    double2 *tab = reinterpret_cast<double2 *>(smem);
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
                pos[threadIdx.x][0] = t.x; pos[threadIdx.x][1] = t.y; pos[threadIdx.x][2] = t.z; qs[threadIdx.x] = t.w;
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
        double *S = c_sys.S + ((int64_t)w * 2 + c_sys.cur[w]) * 2 * nk;
        S[i] = re; S[nk + i] = im;
    }
}

// static host-host pair energy (pairs of host atoms that belong to different molecules)
template <bool TRI>
__global__ void __launch_bounds__(MGPU_BLOCK) k_host_host(const int32_t *host_mol, double *partial /* [grid][2] */)
{
    __shared__ double red[2 * MGPU_WARPS];
    double acc[2] = { 0.0, 0.0 };
    PairCount pc = { 0u, 0u, 0u };
    const int n = c_sys.n_host, nt = c_sys.ntypes;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double4 a = c_sys.host_xyzq[i];
        const int ta = c_sys.host_type[i], ma = host_mol[i];
        for (int j = i + 1; j < n; ++j) {
            if (host_mol[j] == ma) continue;
            const double4 b = c_sys.host_xyzq[j];
            const int ti = ta * nt + c_sys.host_type[j];
            const bool charged = fabs(a.w) >= MGPU_ERR_TOL && fabs(b.w) >= MGPU_ERR_TOL;
            const double r2 = min_image_r2<TRI>(b.x - a.x, b.y - a.y, b.z - a.z);
            pair_terms(r2, 4.0 * c_sys.eps[ti], c_sys.sig[ti], a.w * b.w, charged, acc[0], acc[1], pc);
        }
    }
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) { partial[blockIdx.x * 2] = acc[0]; partial[blockIdx.x * 2 + 1] = acc[1]; }
}

// ------------------------------------------------------------------------------------
// K4: full energy of a walker (update_system_energy, energy_utils.f90:22-39).
// S(k) of the walker must have been rebuilt by k_build_S.  One CTA per walker.
// ------------------------------------------------------------------------------------
template <bool TRI>
__global__ void __launch_bounds__(MGPU_BLOCK) k_total_energy(int first_walker, int natom_max)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout s = carve(smem, natom_max);
    const int w = first_walker + blockIdx.x;
    stage_common(s, w);
    __syncthreads();
    Probe &P = *s.probe;
    double e_lj = c_sys.hh_lj, e_c = 0.0, e_intra = 0.0, e_self = c_sys.self_host_total;
    for (int g = 0; g < c_sys.nres; ++g) {
        if (!c_sys.active[g]) continue;
        const int n = s.count[g], na = c_sys.natom[g];
        for (int m = 0; m < n; ++m) {
            __syncthreads();
            if (threadIdx.x == 0) {
                P.kind = MGPU_KIND_DELETE; P.res = g; P.mol = m; P.na = na; P.has_old = 1; P.has_new = 0;
                P.excl_res = g; P.excl_mol = m; P.order_res = g; P.order_mol = m;
            }
            if (threadIdx.x < na) {
                P.q[threadIdx.x] = c_sys.charge[g][threadIdx.x]; P.type[threadIdx.x] = c_sys.type[g][threadIdx.x];
                load_positions(w, g, m, P.po, threadIdx.x);
            }
            __syncthreads();
            double ps[4];
            pair_sums_cta<TRI>(P, w, s.count, s.eps4, s.sig, s.red, ps);
            e_lj += ps[0]; e_c += ps[1];
            if (threadIdx.x == 0) e_intra += intra_energy<TRI>(g, P.po);
        }
        e_self += c_sys.e_self[g] * n;
    }
    const double *S = c_sys.S + ((int64_t)w * 2 + c_sys.cur[w]) * 2 * c_sys.nk;
    const double recip = recip_energy_cta(S, s.red);
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
struct Rng {
    uint64_t s[4];
    __device__ __forceinline__ double uniform()
    {
        const uint64_t r = rotl(s[1] * 5u, 7) * 9u;
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t; s[3] = rotl(s[3], 45);
        return (double)(r >> 11) * 0x1.0p-53;
    }
    static __device__ __forceinline__ uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
};

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
// device-resident Monte Carlo: one CTA per walker, n_steps iterations of the body of
// monte_carlo_loop (monte_carlo.f90:50-99).  Draw order per step:
// residue pick, molecule pick (if count > 0), move select, [create/delete coin],
// proposal draws, accept draw.
// ------------------------------------------------------------------------------------
struct SweepShared {
    int32_t valid, move, kind, res, mol, accept;
    double com[3];
    double off[MGPU_MAX_SITES][3];
    double prob;
};

template <bool TRI>
__global__ void __launch_bounds__(MGPU_BLOCK) k_sweep(int first_walker, long long n_steps, int natom_max,
                                                      int trace_walker, mgpu_step_trace *trace, int32_t *err)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout s = carve(smem, natom_max);
    __shared__ SweepShared sh;
    const int w = first_walker + blockIdx.x;
    const int tid = threadIdx.x;
    stage_common(s, w);
    Rng rng;
    if (tid == 0) for (int i = 0; i < 4; ++i) rng.s[i] = c_sys.rng[(int64_t)w * 4 + i];
    long long cnt[12];
    double acc_avg[MGPU_MAX_RES][3];
    long long n_samples = 0;
    if (tid == 0) {
        for (int i = 0; i < 12; ++i) cnt[i] = 0;
        for (int g = 0; g < MGPU_MAX_RES; ++g) acc_avg[g][0] = acc_avg[g][1] = acc_avg[g][2] = 0.0;
    }
    __syncthreads();
    const double cumul_translation = c_sys.p_trans;
    const double cumul_rotation = cumul_translation + c_sys.p_rot;
    const double cumul_swap = cumul_rotation + c_sys.p_swap;

    for (long long step = 0; step < n_steps; ++step) {
        if (tid == 0) {
            sh.valid = 0; sh.move = MGPU_MV_NONE; sh.accept = 0; sh.prob = 0.0;
            // pick_random_residue_type / pick_random_molecule_index, monte_carlo_utils.f90:140-201
            const int res = c_sys.active_list[(int)(rng.uniform() * c_sys.nactive)];
            const int n = s.count[res];
            int mol = -1;
            if (n > 0) { mol = (int)(rng.uniform() * n) + 1; if (mol > n) mol = n; mol -= 1; }
            const double draw = rng.uniform();
            const int na = c_sys.natom[res];
            sh.res = res; sh.mol = mol;
            const double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride + c_sys.goff[res];
            const int cap = c_sys.cap[res];
            const double *offs = wc + 3 * (int64_t)cap;
            if (draw <= cumul_translation) {
                if (mol >= 0) {                                           // translation.f90:21-97
                    sh.valid = 1; sh.move = MGPU_MV_TRANSLATE; sh.kind = MGPU_KIND_MOVE;
                    double tp[3];
                    for (int d = 0; d < 3; ++d) tp[d] = rng.uniform();
                    for (int d = 0; d < 3; ++d) sh.com[d] = wc[d * cap + mol] + (tp[d] - 0.5) * c_sys.tstep;
                    apply_PBC(sh.com);
                    for (int e = 0; e < na * 3; ++e) sh.off[e / 3][e % 3] = offs[(int64_t)e * cap + mol];
                }
            } else if (draw <= cumul_rotation) {
                if (na != 1 && mol >= 0) {                                // rotation.f90:22-60
                    sh.valid = 1; sh.move = MGPU_MV_ROTATE; sh.kind = MGPU_KIND_MOVE;
                    for (int d = 0; d < 3; ++d) sh.com[d] = wc[d * cap + mol];
                    for (int e = 0; e < na * 3; ++e) sh.off[e / 3][e % 3] = offs[(int64_t)e * cap + mol];
                    const double theta = (rng.uniform() - 0.5) * c_sys.rstep;
                    const int axis = (int)(rng.uniform() * 3.0) + 1;
                    rotate_offsets(axis, theta, sh.off, na);
                }
            } else if (draw <= cumul_swap) {
                atomicExch(err, 3);                                       // swap moves: host-driven path only
            } else {
                bool create = false, widom = false, del = false;
                if (c_sys.p_insdel > 0) { if (rng.uniform() <= 0.5) create = true; else del = true; }
                else if (c_sys.p_widom > 0) widom = true;
                if (create || widom) {                                    // creation.f90:30-76, widom.f90:30-68
                    if (n >= cap) atomicExch(err, 2);
                    else {
                        sh.valid = 1; sh.move = widom ? MGPU_MV_WIDOM : MGPU_MV_CREATE; sh.kind = MGPU_KIND_CREATE;
                        sh.mol = n;
                        double t3[3];
                        for (int d = 0; d < 3; ++d) t3[d] = rng.uniform();
                        for (int i = 0; i < 3; ++i)
                            sh.com[i] = c_sys.lo[i] + __dadd_rn(__dadd_rn(__dmul_rn(c_sys.H[i * 3 + 0], t3[0]), __dmul_rn(c_sys.H[i * 3 + 1], t3[1])),
                                                                __dmul_rn(c_sys.H[i * 3 + 2], t3[2]));
                        for (int e = 0; e < na * 3; ++e) sh.off[e / 3][e % 3] = offs[(int64_t)e * cap + 0];   // geometry of molecule 1
                        if (na != 1) {
                            const double theta = rng.uniform() * c_sys.twopi;
                            const int axis = (int)(rng.uniform() * 3.0) + 1;
                            rotate_offsets(axis, theta, sh.off, na);
                        }
                    }
                } else if (del) {
                    if (n > 0) { sh.valid = 1; sh.move = MGPU_MV_DELETE; sh.kind = MGPU_KIND_DELETE; }   // deletion.f90:29-77
                }
            }
        }
        __syncthreads();
        double e_old[6], e_new[6];
        if (sh.valid) {
            stage_probe(s, w, sh.kind, sh.res, sh.mol, sh.com, sh.off);
            __syncthreads();
            evaluate_trial<TRI>(s, w, natom_max, e_old, e_new);
            if (tid == 0) {
                const int res = sh.res;
                const double dU = e_new[MGPU_E_TOTAL] - e_old[MGPU_E_TOTAL];
                const double mu = c_sys.mu[(int64_t)w * MGPU_MAX_RES + res];
                const double lam = c_sys.lambda[res];
                double p;
                int acc;
                if (sh.move == MGPU_MV_WIDOM) {                           // widom.f90:74-92
                    p = exp(-dU * c_sys.beta);
                    acc = 0;
                    cnt[10] += 1;
                    if (p > MGPU_ERR_TOL) { cnt[11] += 1; c_sys.widom_w[(int64_t)w * MGPU_MAX_RES + res] += p; }
                    c_sys.widom_n[(int64_t)w * MGPU_MAX_RES + res] += 1;
                    sh.prob = p;
                } else {
                    // compute_acceptance_probability, monte_carlo_utils.f90:204-255
                    if (sh.move == MGPU_MV_CREATE) {
                        const double N = (double)(s.count[res] + 1);     // count already incremented in the reference
                        p = fmin(1.0, c_sys.volume / N / (lam * lam * lam) * exp(-c_sys.beta * (dU - mu)));
                    } else if (sh.move == MGPU_MV_DELETE) {
                        const double Np1 = (double)(s.count[res] - 1) + 1.0;
                        p = fmin(1.0, Np1 * (lam * lam * lam) / c_sys.volume * exp(-c_sys.beta * (dU + mu)));
                    } else p = fmin(1.0, exp(-c_sys.beta * dU));
                    acc = rng.uniform() <= p;
                    const int ci = (sh.move == MGPU_MV_TRANSLATE) ? 0 : (sh.move == MGPU_MV_ROTATE) ? 1 : (sh.move == MGPU_MV_CREATE) ? 2 : 3;
                    cnt[2 * ci] += 1;
                    if (acc) { cnt[2 * ci + 1] += 1; if (ci >= 2) cnt[2 * ci] += 1; }   // creations/deletions bump both slots
                    sh.prob = p;
                }
                sh.accept = acc;
            }
            __syncthreads();
            if (sh.accept) {
                commit_trial(w, sh.kind, sh.res, sh.mol, sh.com, sh.off, e_old, e_new);
                if (tid == 0) {
                    if (sh.kind == MGPU_KIND_CREATE) s.count[sh.res] += 1;
                    if (sh.kind == MGPU_KIND_DELETE) s.count[sh.res] -= 1;
                }
            } else if (sh.kind == MGPU_KIND_CREATE && sh.mol == 0) {
                write_slot(w, sh.res, 0, sh.com, sh.off);      // rejected / Widom insertion into an empty walker
            }
        }
        if (tid == 0) {
            if (trace && w == trace_walker) {
                mgpu_step_trace *t = trace + step;
                t->move = sh.move; t->res = sh.res; t->mol = sh.mol; t->accepted = sh.accept;
                t->prob = sh.prob;
                if (sh.valid) {
                    t->dE = e_new[MGPU_E_TOTAL] - e_old[MGPU_E_TOTAL];
                    for (int i = 0; i < 6; ++i) { t->e_old[i] = e_old[i]; t->e_new[i] = e_new[i]; }
                } else {
                    t->dE = 0.0;
                    for (int i = 0; i < 6; ++i) { t->e_old[i] = 0.0; t->e_new[i] = 0.0; }
                }
            }
            for (int g = 0; g < c_sys.nres; ++g) {
                if (!c_sys.active[g]) continue;
                const double N = (double)s.count[g];
                acc_avg[g][0] += N; acc_avg[g][1] += N * N;
            }
            acc_avg[0][2] += c_sys.energy[(int64_t)w * 6 + MGPU_E_TOTAL];
            n_samples += 1;
        }
        __syncthreads();     // commit visible (coordinates, S flip, counts) before the next step
    }
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) c_sys.rng[(int64_t)w * 4 + i] = rng.s[i];
        for (int i = 0; i < 12; ++i) c_sys.counters[(int64_t)w * 12 + i] += cnt[i];
        for (int g = 0; g < c_sys.nres; ++g) {
            double *A = c_sys.avg + ((int64_t)w * MGPU_MAX_RES + g) * 4;
            A[0] += acc_avg[g][0]; A[1] += acc_avg[g][1]; A[2] += acc_avg[0][2]; A[3] += (double)n_samples;
        }
    }
}

// ------------------------------------------------------------------------------------
// K3: Widom batch.  Real space: one thread per insertion, every lane reads the same
// target (broadcast from L1).  k space: one warp per insertion, lanes over k-vectors.
// u(id,k) = top53(mix(seed + GOLDEN*(8*id + k + 1))), k = 0..4: COM x3, angle, axis.
// ------------------------------------------------------------------------------------
#define MGPU_WIDOM_BLOCK 128
template <bool TRI>
__global__ void __launch_bounds__(MGPU_WIDOM_BLOCK) k_widom_batch(int w, int res, long long first_id, long long n,
                                                                  unsigned long long seed, double *dE_out,
                                                                  double *block_sum_w, long long *block_n_ok, int natom_max)
{
    extern __shared__ __align__(16) unsigned char smem[];
    // layout: eps4[nt2], sig[nt2], per-warp phase table [warps][natom_max*3*KW] double2,
    //         per-thread positions [block][natom_max][3], e_real[block]
    const int nt2 = c_sys.ntypes * c_sys.ntypes;
    const int KW = c_sys.kmax_max + 1;
    double *s_eps4 = reinterpret_cast<double *>(smem);
    double *s_sig = s_eps4 + nt2;
    size_t o = (sizeof(double) * 2 * nt2 + 15) & ~size_t(15);
    double2 *s_tab = reinterpret_cast<double2 *>(smem + o); o += sizeof(double2) * (MGPU_WIDOM_BLOCK / 32) * (size_t)natom_max * 3 * KW;
    double(*s_pos)[3] = reinterpret_cast<double(*)[3]>(smem + o); o += sizeof(double) * 3 * (size_t)natom_max * MGPU_WIDOM_BLOCK;
    double *s_e = reinterpret_cast<double *>(smem + o);
    __shared__ double red_w[MGPU_WIDOM_BLOCK / 32];
    __shared__ long long red_n[MGPU_WIDOM_BLOCK / 32];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < nt2; i += blockDim.x) { s_eps4[i] = 4.0 * c_sys.eps[i]; s_sig[i] = c_sys.sig[i]; }
    __syncthreads();
    const int na = c_sys.natom[res];
    const long long gid = (long long)blockIdx.x * blockDim.x + tid;
    const bool live = gid < n;
    const unsigned long long id = (unsigned long long)(first_id + gid);
    double(*mypos)[3] = s_pos + (size_t)tid * natom_max;
    double e_lj = 0.0, e_c = 0.0, e_intra = 0.0;
    PairCount pc = { 0u, 0u, 0u };
    const double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride;
    if (live) {
        double u[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) u[k] = (double)(splitmix64_mix(seed + MGPU_GOLDEN * (8ull * id + (unsigned long long)k + 1ull)) >> 11) * 0x1.0p-53;
        double com[3];
        for (int i = 0; i < 3; ++i)
            com[i] = c_sys.lo[i] + __dadd_rn(__dadd_rn(__dmul_rn(c_sys.H[i * 3 + 0], u[0]), __dmul_rn(c_sys.H[i * 3 + 1], u[1])),
                                             __dmul_rn(c_sys.H[i * 3 + 2], u[2]));
        // geometry of molecule 1 of this residue type (insert_and_orient_molecule, :563-568)
        const int cap = c_sys.cap[res];
        const double *offs = wc + c_sys.goff[res] + 3 * (int64_t)cap;
        for (int e = 0; e < na * 3; ++e) mypos[e / 3][e % 3] = offs[(int64_t)e * cap + 0];
        if (na != 1) {
            const double theta = u[3] * c_sys.twopi;
            const int axis = (int)(u[4] * 3.0) + 1;
            rotate_offsets(axis, theta, mypos, na);
        }
        for (int a = 0; a < na; ++a)
            for (int d = 0; d < 3; ++d) mypos[a][d] = com[d] + mypos[a][d];
        e_intra = intra_energy<TRI>(res, mypos);
    }
    // real space: all lanes walk the same target list
    if (live) {
        const int nt = c_sys.ntypes;
        for (int a = 0; a < na; ++a) {
            const double px = mypos[a][0], py = mypos[a][1], pz = mypos[a][2];
            const double qa = c_sys.charge[res][a];
            const bool acharged = fabs(qa) >= MGPU_ERR_TOL;
            const int trow = c_sys.type[res][a] * nt;
            for (int j = 0; j < c_sys.n_host; ++j) {
                const double4 t = c_sys.host_xyzq[j];
                const int ti = trow + __ldg(&c_sys.host_type[j]);
                const bool charged = acharged && fabs(t.w) >= MGPU_ERR_TOL;
                const double eps4 = s_eps4[ti];
                if (eps4 == 0.0 && !charged) continue;
                const double r2 = min_image_r2<TRI>(t.x - px, t.y - py, t.z - pz);
                pair_terms(r2, eps4, s_sig[ti], qa * t.w, charged, e_lj, e_c, pc);
            }
            for (int g = 0; g < c_sys.nres; ++g) {
                if (!c_sys.active[g]) continue;
                const int capg = c_sys.cap[g], nag = c_sys.natom[g], ng = c_sys.count[(int64_t)w * MGPU_MAX_RES + g];
                const double *comg = wc + c_sys.goff[g];
                const double *offg = comg + 3 * (int64_t)capg;
                for (int m = 0; m < ng; ++m)
                    for (int b = 0; b < nag; ++b) {
                        const double *ob = offg + (int64_t)b * 3 * capg;
                        const double tx = comg[m] + ob[m], ty = comg[capg + m] + ob[capg + m], tz = comg[2 * capg + m] + ob[2 * capg + m];
                        const double tq = c_sys.charge[g][b];
                        const int ti = trow + c_sys.type[g][b];
                        const bool charged = acharged && fabs(tq) >= MGPU_ERR_TOL;
                        const double eps4 = s_eps4[ti];
                        if (eps4 == 0.0 && !charged) continue;
                        const double r2 = min_image_r2<TRI>(tx - px, ty - py, tz - pz);
                        pair_terms(r2, eps4, s_sig[ti], qa * tq, charged, e_lj, e_c, pc);
                    }
            }
        }
    }
    flush_pair_count(pc);
    s_e[tid] = e_lj + e_c * c_sys.eps0_inv_real + c_sys.e_self[res] + e_intra;   // non-recip part of new%total
    __syncthreads();
    // k space: warp `wid` handles the 32 insertions of its own lanes, one after the other
    const int nk = c_sys.nk;
    const double *S_in = c_sys.S + ((int64_t)w * 2 + c_sys.cur[w]) * 2 * nk;
    const double recip_cur = c_sys.energy[(int64_t)w * 6 + MGPU_E_RECIP];
    double2 *tab = s_tab + (size_t)wid * natom_max * 3 * KW;
    double my_w = 0.0; long long my_ok = 0;
    for (int l = 0; l < 32; ++l) {
        const long long g2 = (long long)blockIdx.x * blockDim.x + wid * 32 + l;
        if (g2 >= n) break;                                   // uniform across the warp
        const double(*pp)[3] = s_pos + (size_t)(wid * 32 + l) * natom_max;
        __syncwarp();
        fill_phase_tables(tab, pp, na, lane, 32);
        __syncwarp();
        double part = 0.0;
        for (int i = lane; i < nk; i += 32) {
            const int kx = c_sys.kx[i], ky = c_sys.ky[i], kz = c_sys.kz[i];
            double sr = 0.0, si = 0.0;
            for (int a = 0; a < na; ++a) {
                const cplx p = phase_product(tab, a, kx, ky, kz);
                const double q = c_sys.charge[res][a];
                sr += q * p.re; si += q * p.im;
            }
            const double re = S_in[i] + sr, im = S_in[nk + i] + si;
            part += c_sys.ffW[i] * (re * re + im * im);
        }
        part = warp_sum(part);
        if (lane == l) {
            const double recip_new = part * c_sys.eps0_inv_real * c_sys.twopi / c_sys.volume;
            const double e_new_total = s_e[tid] + recip_new;     // same association as new%total up to ordering
            const double dU = e_new_total - recip_cur;
            if (dE_out) dE_out[gid] = dU;
            const double wgt = exp(-dU * c_sys.beta);
            if (wgt > MGPU_ERR_TOL) { my_w = wgt; my_ok = 1; }
        }
    }
    // block reduction of weights (fixed order)
    my_w = warp_sum(my_w);
    unsigned okmask = __ballot_sync(0xffffffffu, my_ok != 0);
    if (lane == 0) { red_w[wid] = my_w; red_n[wid] = __popc(okmask); }
    __syncthreads();
    if (tid == 0) {
        double sw = 0.0; long long sn = 0;
        for (int i = 0; i < MGPU_WIDOM_BLOCK / 32; ++i) { sw += red_w[i]; sn += red_n[i]; }
        block_sum_w[blockIdx.x] = sw; block_n_ok[blockIdx.x] = sn;
    }
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
