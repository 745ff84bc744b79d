// maniac_gpu.cu -- C ABI (include/maniac_gpu.h) and context management of the sm_100a
// energy engine.  Host-side setup restates the reference's *setup* arithmetic
// (src/prepare_utils.f90:110-259, src/ewald_kvectors.f90:24-175, src/constants.f90) so the
// Fortran host and the engine agree on alpha / kmax / the k-vector list bit for bit.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <map>

#ifndef MGPU_TRI_COLUMNS
#define MGPU_TRI_COLUMNS 1        // triclinic cells: framework atoms in columns along the axis whose faces no single listed vector needs (see mgpu_init)
#endif
#ifndef MGPU_HILBERT
#define MGPU_HILBERT 1            // framework atoms along a Hilbert curve (0: Z curve, round 1)
#endif
#include "mgpu_kernels.cuh"
#include "mgpu_records.cuh"
#include "mgpu_table.h"

// ---- constants (src/constants.f90:8-21, src/parameters.f90:31-37) --------------------
namespace {
const double PI = 3.1415926536;                    // truncated on purpose
const double TWOPI = 2.0 * PI;
const double H_PLANCK = 6.62607015e-34;
const double EPS0 = 8.854187817e-12;
const double KB = 1.380658e-23;
const double E_CHARGE = 1.602176634e-19;
const double G_TO_KG = 1.0e-3;
const double M_TO_A = 1.0e10;
inline double NA() { return (double)6.02214076e23f; }          // single-precision literal
inline double J_to_kcal() { return (double)0.000239005736f; }  // single-precision literal
inline double OVERLAP() { return (double)1.0e20f; }            // "1.0e20" literal
inline double SQRTPI() { return std::sqrt(PI); }
inline double EPS0_INV_real()
{
    const double eps0_inv = E_CHARGE * E_CHARGE / (4.0 * PI * EPS0);
    return eps0_inv * J_to_kcal() * M_TO_A * NA();
}
inline double KB_kcalmol() { return KB * NA() * J_to_kcal(); }
inline int f_nint(double x) { return (int)std::lround(x); }

struct TimingSlot { double ms = 0.0; long long launches = 0; };

struct Context {
    bool ready = false;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DevSys h{};                       // host copy of the constant block
    int natom_max = 1;
    size_t smem1 = 0, smem8 = 0, smem_buildS = 0;   // dynamic shared memory: 1 group (CTA) / MGPU_WARPS groups (warps) per CTA
    int sm_count = 0, ctas_per_sm = 1;
    int phase_sync = -1;               // MGPU_OPT_PHASE_SYNC: -1 = the measured best set per launch shape (warps: 1|4 + adaptive groups, teams: 1|4|8); else bits: 1 top of the MC step, 2 before the energy evaluation, 4 before the guest pass, 8 before k-space, 16 nearly empty CTAs run free
    int tri_listed = 0;                // triclinic: candidates listed but provably irrelevant (see mgpu_init)
    int tri_req[MGPU_TRI_MAXREL] = {};  // faces (bit d = axis d) the rounded vector must be near for listed vector k to matter
    int wgroups = MGPU_WGROUPS;        // walkers (warps) per CTA of the warp-per-task kernels
    size_t smem_team = 0;              // dynamic shared memory of the team sweep (wgroups * 32 / MGPU_TEAM walkers per CTA)
    size_t smem_team2 = 0;             // ... of the two-warp team sweep (wgroups * 32 / MGPU_TEAM2 walkers per CTA)
    int sweep_team = -1;               // MGPU_OPT_SWEEP_TEAM: -1 auto (see sweep_shape), 0 one warp per walker, 1 four warps, 2 two warps
    int tab_emin = 0, tab_noct = 0;
    std::vector<void *> allocs;
    // host-side mirrors needed by the API
    std::vector<int> natom, active, cap;
    std::vector<double> fug;
    // staging
    int task_cap = 0;
    int32_t *d_task_i = nullptr;      // [task_cap] int4 {walker,res,mol,kind}
    double *d_task_com = nullptr, *d_task_off = nullptr, *d_task_out = nullptr;
    int32_t *d_accept = nullptr, *d_err = nullptr;
    int32_t *h_task_i = nullptr; double *h_task_com = nullptr, *h_task_off = nullptr, *h_task_out = nullptr;
    int32_t *h_accept = nullptr;
    double *d_scratch = nullptr;      // small scalar results
    double *h_scratch = nullptr;
    double *d_geom = nullptr;         // com + off of one molecule
    std::map<std::string, TimingSlot> timing;
    std::vector<char> dirty;          // per walker: coordinates changed outside commit
    std::vector<double> rmax2;        // host mirror of DevSys::rmax2 (grown by mgpu_set_molecule)
    std::vector<char> pending;        // per walker: a trial is pending (host mirror of MgpuTrial::active)
    std::vector<int32_t> pend_kind, pend_res, pend_res2;   // ... and what it would change (host mirror of the counts below)
    std::vector<int32_t> counts;      // host mirror of DevSys::count [W][MGPU_MAX_RES]; refreshed lazily (counts_valid)
    bool counts_valid = false;
    std::vector<uint32_t> batch_stamp; uint32_t batch_id = 0;   // duplicate walkers inside one mgpu_trial_batch
    // pipelined mgpu_block: slices of walkers on several streams (copy of slice i+1 / i-1 under the sweep of slice i)
    static const int NPIPE = 16;
    cudaStream_t pstream[NPIPE] = {};
    cudaEvent_t pevent[NPIPE] = {};
    double *p_in[NPIPE] = {}, *p_out[NPIPE] = {};                 // device staging of the slice's records, in / out
    size_t p_in_cap[NPIPE] = {}, p_out_cap[NPIPE] = {};
    long long *p_off_in[NPIPE] = {}, *p_off_out[NPIPE] = {};      // device: record offsets local to the slice
    long long *ph_off_in[NPIPE] = {}, *ph_off_out[NPIPE] = {};    // pinned host mirrors of the two
    size_t p_off_cap[NPIPE] = {};
    int block_slices = 8;             // MGPU_OPT_BLOCK_SLICES (0 / 1 = the unpipelined load -> sweep -> save)
    // latency path (few trials per call, e.g. one unchanged Fortran process): the kernels read the task and write
    // the energies straight from / to page-locked host memory (UVA), so a call is one launch + one synchronisation
    int32_t *hc_walker = nullptr, *hc_accept = nullptr;    // pinned, commit arguments
    bool commit_inflight = false;
    // walker records (mgpu_save_walkers / mgpu_load_walkers / mgpu_block)
    double *d_blob = nullptr; size_t blob_cap = 0;
    long long *d_rec_off = nullptr; size_t rec_cap = 0;
    long long h2d_bytes = 0, d2h_bytes = 0;
};
Context g;
std::string g_err;

int fail(const std::string &m) { g_err = m; return 1; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
// every export starts here; a commit queued by the latency path (no synchronisation of its own) is waited for first,
// because the getters copy on the legacy stream, which does not order against the engine's non-blocking stream
#define NEED_READY_NOSYNC() do { if (!g.ready) return fail("mgpu: engine not initialised (mgpu_init)"); } while (0)
#define NEED_READY() do { NEED_READY_NOSYNC(); if (g.commit_inflight) { cudaStreamSynchronize(g.stream); g.commit_inflight = false; } } while (0)

template <typename T> int dalloc(T **p, size_t n)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, (n ? n : 1) * sizeof(T));
    if (e != cudaSuccess) return fail(std::string("cudaMalloc: ") + cudaGetErrorString(e));
    g.allocs.push_back(q);
    *p = reinterpret_cast<T *>(q);
    return 0;
}
int upload_sys() { CK(cudaMemcpyToSymbolAsync(c_sys, &g.h, sizeof(DevSys), 0, cudaMemcpyHostToDevice, g.stream)); return 0; }

struct Timer {
    const char *name;
    explicit Timer(const char *n) : name(n) { cudaEventRecord(g.ev0, g.stream); }
    void stop(long long launches = 1)
    {
        cudaEventRecord(g.ev1, g.stream);
        cudaEventSynchronize(g.ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g.ev0, g.ev1);
        TimingSlot &s = g.timing[name];
        s.ms += ms; s.launches += launches;
    }
};

int check_walker(int w) { if (w < 0 || w >= g.h.n_walkers) return fail("mgpu: walker index out of range"); return 0; }
int check_guest(int res) { if (res < 0 || res >= g.h.nres || !g.h.active[res]) return fail("mgpu: residue index out of range or not an active (guest) residue"); return 0; }

int ensure_task_cap(int n)
{
    if (n <= g.task_cap) return 0;
    int cap = 1; while (cap < n) cap <<= 1;
    // (old buffers stay in allocs and are released in finalize)
    if (dalloc(&g.d_task_i, (size_t)4 * cap)) return 1;
    if (dalloc(&g.d_task_com, (size_t)3 * cap)) return 1;
    if (dalloc(&g.d_task_off, (size_t)3 * MGPU_MAX_SITES * cap)) return 1;
    if (dalloc(&g.d_task_out, (size_t)12 * cap)) return 1;
    if (dalloc(&g.d_accept, (size_t)cap)) return 1;
    if (g.h_task_i) { cudaStreamSynchronize(g.stream); cudaFreeHost(g.h_task_i); cudaFreeHost(g.h_task_com); cudaFreeHost(g.h_task_off); cudaFreeHost(g.h_task_out); cudaFreeHost(g.h_accept); cudaFreeHost(g.hc_walker); cudaFreeHost(g.hc_accept); }
    CK(cudaMallocHost(&g.h_task_i, sizeof(int32_t) * 4 * cap));
    CK(cudaMallocHost(&g.h_task_com, sizeof(double) * 3 * cap));
    CK(cudaMallocHost(&g.h_task_off, sizeof(double) * 3 * MGPU_MAX_SITES * cap));
    CK(cudaMallocHost(&g.h_task_out, sizeof(double) * 12 * cap));
    CK(cudaMallocHost(&g.h_accept, sizeof(int32_t) * cap));
    CK(cudaMallocHost(&g.hc_walker, sizeof(int32_t) * cap));
    CK(cudaMallocHost(&g.hc_accept, sizeof(int32_t) * cap));
    g.task_cap = cap;
    return 0;
}

int check_err_flag(const char *where)
{
    int32_t e = 0;
    CK(cudaMemcpyAsync(&e, g.d_err, sizeof e, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    if (e) {
        int32_t z = 0;
        cudaMemcpyAsync(g.d_err, &z, sizeof z, cudaMemcpyHostToDevice, g.stream);
        cudaStreamSynchronize(g.stream);
        if (e == 1) return fail(std::string(where) + ": commit without a pending trial");
        if (e == 4) return fail(std::string(where) + ": malformed walker record (length / counts / capacity)");
        if (e == 2) return fail(std::string(where) + ": Trying to insert a molecule beyond the walker's capacity (NB_MAX_MOLECULE analogue)");
        return fail(std::string(where) + ": device error flag");
    }
    return 0;
}

// rebuild S(k) and the running energies of walkers [first, first+n)
int rebuild(int first, int n)
{
    const int nk = g.h.nk;
    dim3 grid((nk + MGPU_BLOCK - 1) / MGPU_BLOCK, n);
    if (nk > 0) k_build_S<<<grid, MGPU_BLOCK, g.smem_buildS, g.stream>>>(1, nullptr, first);
    if (g.h.triclinic) k_total_energy<true><<<n, MGPU_BLOCK, g.smem1, g.stream>>>(first, g.natom_max);
    else k_total_energy<false><<<n, MGPU_BLOCK, g.smem1, g.stream>>>(first, g.natom_max);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(g.stream));
    for (int w = first; w < first + n; ++w) g.dirty[w] = 0;
    return 0;
}
int ensure_clean(int w) { if (g.dirty[w]) return rebuild(w, 1); return 0; }
// ---- triclinic minimum image: which lattice vectors can beat the fractionally rounded image, and where (host-only math, also
// behind mgpu_plan_triclinic so that it can be tested without a device) ----
struct TriPlan {
    int nrel = 0;                                  // listed vectors (one of every +-m pair); -1 = too skewed: literal 27-image search
    double rel[MGPU_TRI_MAXREL][3], m[MGPU_TRI_MAXREL][3], len2[MGPU_TRI_MAXREL];
    int req[MGPU_TRI_MAXREL];                      // faces (bit d = axis d) the rounded vector must be near for vector k to matter
    double eps[3] = { 0.0, 0.0, 0.0 };             // largest such face distance per axis (fractional units)
    double safe2 = 1e300;                          // a listed vector can only shorten t when |t|^2 > safe2 = min |C m|^2 / 4
};
static void plan_triclinic(const double M[3][3], TriPlan &tp)
{
    // lattice vectors that can beat the fractionally rounded image (min_image_r2<true>): m is relevant iff
    // min over f in [-1/2,1/2]^3 of |C(f+m)|^2 - |C f|^2 = m.G.m - sum_d |(G m)_d| < 0, G = C^T C (C = columns of matrix)
    double G[3][3];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { G[a][b] = 0.0; for (int i = 0; i < 3; ++i) G[a][b] += M[i][a] * M[i][b]; }
    int n = 0;
    for (int a = -3; a <= 3 && n >= 0; ++a) for (int b = -3; b <= 3 && n >= 0; ++b) for (int c = -3; c <= 3; ++c) {
        if (!a && !b && !c) continue;
        if (a < 0 || (a == 0 && (b < 0 || (b == 0 && c < 0)))) continue;       // one of every +-m pair
        const double m[3] = { (double)a, (double)b, (double)c };
        double Gm[3], mGm = 0.0, s1 = 0.0;
        for (int d = 0; d < 3; ++d) { Gm[d] = G[d][0] * m[0] + G[d][1] * m[1] + G[d][2] * m[2]; mGm += m[d] * Gm[d]; s1 += std::fabs(Gm[d]); }
        if (!(mGm < s1 * (1.0 - 1e-13))) continue;
        if (n == MGPU_TRI_MAXREL || std::abs(a) == 3 || std::abs(b) == 3 || std::abs(c) == 3) { n = -1; break; }   // very skewed cell: literal search
        for (int i = 0; i < 3; ++i) { tp.rel[n][i] = M[i][0] * m[0] + M[i][1] * m[1] + M[i][2] * m[2]; tp.m[n][i] = m[i]; }
        tp.len2[n] = tp.rel[n][0] * tp.rel[n][0] + tp.rel[n][1] * tp.rel[n][1] + tp.rel[n][2] * tp.rel[n][2];
        ++n;
    }
    tp.nrel = n;
    tp.safe2 = 1e300;
    for (int k = 0; k < n; ++k) tp.safe2 = std::fmin(tp.safe2, 0.25 * tp.len2[k]);
    // Gate of the candidate search in fractional space.  With f_d = +-(1/2 - u_d), u_d in [0, 1/2]:
    //   |t|^2 - |t -+ C m|^2 = 2 |f . G m| - m.G.m  <=  s1 - m.G.m - 2 sum_d u_d |(G m)_d|,   s1 = sum_d |(G m)_d|,
    // so m can only help when sum_d u_d |(G m)_d| < (s1 - m.G.m) / 2 =: D_m, hence u_d < D_m / |(G m)_d| for EVERY axis d
    // with (G m)_d != 0.  Axes where that bound is below 1/2 are the faces the rounded vector has to be near for m to
    // matter (req[k], a 3-bit set); eps[d] = the largest such bound of axis d over the listed vectors.
    tp.eps[0] = tp.eps[1] = tp.eps[2] = 0.0;
    for (int k = 0; k < n; ++k) {
        double Gm[3], mGm = 0.0, s1 = 0.0;
        for (int d = 0; d < 3; ++d) { Gm[d] = G[d][0] * tp.m[k][0] + G[d][1] * tp.m[k][1] + G[d][2] * tp.m[k][2]; mGm += tp.m[k][d] * Gm[d]; s1 += std::fabs(Gm[d]); }
        tp.req[k] = 0;
        for (int d = 0; d < 3; ++d) {
            if (std::fabs(Gm[d]) == 0.0) continue;
            const double e = 0.5 * (s1 - mGm) / std::fabs(Gm[d]) * (1.0 + 1e-9) + 1e-12;
            if (e >= 0.5) continue;                                            // no constraint from this axis
            tp.req[k] |= 1 << d;
            tp.eps[d] = std::fmax(tp.eps[d], e);
        }
    }
}
// thr_hi[d]: high word of the |f_d| from which a listed vector can matter near face d; lut[faces]: which listed vectors have to
// be tried when the lanes of a warp are near the faces in `faces` (bits 0-2 = axes, bit 3 = some |g_d| >= 1.5, i.e. an atom far
// outside the cell).  Bit k = vector k; bit 31 = the complete search (every vector, then the reference's 27 images if the winner
// leaves {-1,0,1}^3).  0 = the rounded image is the answer.
static void plan_triclinic_gate(bool triclinic, int nrel, const int *req, const double *eps, int32_t thr_hi[3], uint32_t lut[16])
{
    for (int d = 0; d < 3; ++d) {
        thr_hi[d] = (triclinic && nrel < 0) ? 0 : 0x7ff00000;      // very skewed cell: always the literal search; else never ...
        if (triclinic && nrel > 0 && eps[d] > 0.0) {                // ... unless a listed vector can matter near this face
            const double thr = std::fmax(0.0, 0.5 - eps[d]);
            uint64_t bits; std::memcpy(&bits, &thr, 8);
            thr_hi[d] = (int32_t)(bits >> 32);                      // hi(|f|) >= hi(thr) is implied by |f| >= thr: a superset
        }
    }
    for (int f = 0; f < 16; ++f) {
        uint32_t m = 0;
        if (triclinic) {
            if (nrel < 0 || (f & 8)) m = 0x80000000u | ((nrel > 0) ? ((1u << nrel) - 1u) : 0u);
            else for (int k = 0; k < nrel; ++k) if ((req[k] & f) == req[k]) m |= 1u << k;
        }
        lut[f] = m;
    }
}
// prepare_simulation_box (geometry_utils.f90:148-200,293-361): cell edge lengths (cell vectors = COLUMNS of the matrix), volume
// and the reference's "reciprocal" = transposed inverse, by the adjugate as the reference computes it.  Returns 1 when the
// determinant is degenerate (the reference's own check).
static int box_geometry(const double M[3][3], double metrics[3], double hinv[9], double *volume)
{
    auto colv = [&](int j, double v[3]) { for (int i = 0; i < 3; ++i) v[i] = M[i][j]; };
    auto cross = [](const double a[3], const double b[3], double c[3]) {
        c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0]; };
    double a[3], b[3], c[3], bxc[3];
    colv(0, a); colv(1, b); colv(2, c);
    for (int j = 0; j < 3; ++j) { double v[3]; colv(j, v); metrics[j] = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    cross(b, c, bxc);
    *volume = std::fabs(a[0] * bxc[0] + a[1] * bxc[1] + a[2] * bxc[2]);
    double adj[3][3], v1[3], v2[3], cr[3];
    colv(1, v1); colv(2, v2); cross(v1, v2, cr); for (int i = 0; i < 3; ++i) adj[i][0] = cr[i];
    colv(2, v1); colv(0, v2); cross(v1, v2, cr); for (int i = 0; i < 3; ++i) adj[i][1] = cr[i];
    colv(0, v1); colv(1, v2); cross(v1, v2, cr); for (int i = 0; i < 3; ++i) adj[i][2] = cr[i];
    colv(0, v1);
    const double det = v1[0] * adj[0][0] + v1[1] * adj[1][0] + v1[2] * adj[2][0];
    if (std::fabs(det) < 1.0) return 1;
    const double rcp = 1.0 / det;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) hinv[i * 3 + j] = rcp * adj[i][j];
    return 0;
}
// Order of the static framework atoms (host-only; also behind mgpu_plan_framework_order for tests without a device).
// Default: along a space-filling curve of the fractional coordinates, so that the 32 consecutive atoms a warp takes per
// iteration sit in one small region and "beyond the LJ cutoff" / "no lattice vector can help" become warp-uniform facts.
// hinv = the reference's "reciprocal" (transposed inverse: f_d = sum_j hinv[j*3+d] r_j); metrics = cell edge lengths;
// req / nrel = the triclinic plan (columns only when `columns` is set); cols_out = bins per axis (1 = not binned).
static void order_framework(int n_host, const double *xyz, const double lo[3], const double hinv[9], const double metrics[3],
                            bool columns, const int *req, int nrel, std::vector<int> &order, int cols_out[3])
{
#if MGPU_HILBERT
    // Hilbert curve (Skilling's transpose form, 10 bits per axis): unlike the Z curve it has no jumps, so EVERY run of 32
    // consecutive atoms is one compact cluster -- with Z order a run that straddles a jump of the curve is two clusters
    auto curve_key = [](uint32_t x, uint32_t y, uint32_t z) {
        uint32_t X[3] = { x & 1023u, y & 1023u, z & 1023u };
        const uint32_t Mb = 1u << 9;
        for (uint32_t Q = Mb; Q > 1; Q >>= 1) {
            const uint32_t P = Q - 1;
            for (int i = 0; i < 3; ++i) {
                if (X[i] & Q) X[0] ^= P;
                else { const uint32_t t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
            }
        }
        for (int i = 1; i < 3; ++i) X[i] ^= X[i - 1];
        uint32_t t = 0;
        for (uint32_t Q = Mb; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
        for (int i = 0; i < 3; ++i) X[i] ^= t;
        uint32_t k = 0;
        for (int b = 9; b >= 0; --b) for (int i = 0; i < 3; ++i) k = (k << 1) | ((X[i] >> b) & 1u);
        return k;
    };
#else
    auto spread = [](uint32_t v) { v &= 1023u; v = (v | (v << 16)) & 0x030000FFu; v = (v | (v << 8)) & 0x0300F00Fu; v = (v | (v << 4)) & 0x030C30C3u; v = (v | (v << 2)) & 0x09249249u; return v; };
    auto curve_key = [&](uint32_t x, uint32_t y, uint32_t z) { return spread(x) | (spread(y) << 1) | (spread(z) << 2); };
#endif
    // Triclinic cells whose listed lattice vectors leave one axis free of single-face vectors (e.g. an xy tilt: the vectors
    // that matter near ONE face are +-c1 and +-c2, none for the z faces): COLUMNS instead of compact clusters -- thin
    // (about 4 A) along the axes whose faces matter, long along the free axis.  A warp iteration takes the candidate search
    // of the triclinic minimum image when any of its 32 atoms is within ~3 A of such a face: simulated on configs[4], 32 % of
    // the iterations for Hilbert runs, 20 % for 16 x 16 columns (and 17 % instead of 13 % of them reach into the LJ cutoff
    // sphere, which costs less than it saves).
    int thin[3] = { 0, 0, 0 }, n_thin = 0, cols[3] = { 1, 1, 1 };
    if (columns && nrel > 0 && MGPU_TRI_COLUMNS) {
        for (int k2 = 0; k2 < nrel; ++k2) for (int d = 0; d < 3; ++d) if (req[k2] == (1 << d)) thin[d] = 1;
        n_thin = thin[0] + thin[1] + thin[2];
        if (n_thin == 1 || n_thin == 2) {
            long long ncol = 1;
            for (int d = 0; d < 3; ++d) if (thin[d]) { cols[d] = std::max(1, (int)std::lround(metrics[d] / 4.25)); ncol *= cols[d]; }
            while (ncol * 32 > n_host) {                       // keep at least one warp iteration per column
                ncol = 1;
                for (int d = 0; d < 3; ++d) if (thin[d]) { cols[d] = std::max(1, cols[d] - 1); ncol *= cols[d]; }
                if (ncol == 1) break;
            }
        } else n_thin = 0;
    }
    std::vector<uint64_t> key64(n_host);
    for (int k = 0; k < n_host; ++k) {
        const double r[3] = { xyz[(size_t)k * 3] - lo[0], xyz[(size_t)k * 3 + 1] - lo[1], xyz[(size_t)k * 3 + 2] - lo[2] };
        uint32_t q[3];
        uint64_t col = 0;
        for (int d = 0; d < 3; ++d) {
            double f = hinv[0 * 3 + d] * r[0] + hinv[1 * 3 + d] * r[1] + hinv[2 * 3 + d] * r[2];
            f -= std::floor(f);
            q[d] = (uint32_t)std::fmin(1023.0, f * 1024.0);
            if (n_thin && thin[d]) {
                int b = std::min(cols[d] - 1, (int)(f * cols[d]));
                if (col & 1) b = cols[d] - 1 - b;              // serpentine: neighbouring columns follow each other
                col = col * cols[d] + b;
                q[d] = 0;                                      // the curve only runs along the free axes inside a column
            }
        }
        uint32_t ck = curve_key(q[0], q[1], q[2]);
        if (n_thin) {                                          // Z order of the free coordinates (monotone along a single free axis)
            auto spread3 = [](uint32_t v) { v &= 1023u; v = (v | (v << 16)) & 0x030000FFu; v = (v | (v << 8)) & 0x0300F00Fu; v = (v | (v << 4)) & 0x030C30C3u; v = (v | (v << 2)) & 0x09249249u; return v; };
            ck = spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2);
        }
        key64[k] = (col << 32) | ck;
    }
    order.resize(n_host);
    for (int k = 0; k < n_host; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key64[a] < key64[b]; });
    for (int d = 0; d < 3; ++d) cols_out[d] = n_thin ? cols[d] : 1;
}
// One launch of the device-resident drivers for walkers [first, first + n).  Shape = threads per walker x walkers per CTA
// (one CTA per SM), chosen from the number of walkers in flight so that ONE wave covers every SM when the walkers allow it:
//   * a team of four warps per walker (up to wgroups / 4 walkers per CTA) while the walkers fit such a wave,
//   * a team of two warps (up to wgroups / 2 per CTA) up to twice as many,
//   * one warp per walker (up to wgroups per CTA: the throughput shape) beyond that;
// and the walkers per CTA are the smallest number that still needs no extra round of CTAs -- e.g. 4096 walkers on 148 SMs
// run as 2 rounds of 14-walker CTAs (all SMs busy in both) rather than 1.73 rounds of 16-walker CTAs.  Fixed isotherms
// spread over more GPUs are the case this is for.  MGPU_OPT_SWEEP_TEAM forces the threads per walker
// (0: 32, 1: 128, 2: 64) with full CTAs.
struct SweepShape { int nt, per_cta; };
// pure: slots = warps of a full sweep CTA, forced = MGPU_OPT_SWEEP_TEAM (also behind mgpu_plan_sweep_shape, for tests without a device)
static SweepShape plan_sweep_shape(int n_total, int sm_count, int slots, int forced)
{
    int nt = 32;
    const bool can4 = slots * 32 >= MGPU_TEAM, can2 = slots * 32 >= MGPU_TEAM2;
    if (forced == 1 && can4) nt = MGPU_TEAM;
    else if (forced == 2 && can2) nt = MGPU_TEAM2;
    else if (forced < 0) {
        if (can4 && (long long)n_total <= (long long)sm_count * (slots / 4)) nt = MGPU_TEAM;
        else if (can2 && (long long)n_total <= (long long)sm_count * (slots / 2)) nt = MGPU_TEAM2;
    }
    const int max_per = std::max(1, slots * 32 / nt);
    int per = max_per;
    if (forced < 0) {
        const long long wave = (long long)max_per * sm_count;
        const long long rounds = (n_total + wave - 1) / wave;
        per = (int)((n_total + rounds * sm_count - 1) / (rounds * sm_count));
        per = std::max(1, std::min(per, max_per));
    }
    return { nt, per };
}
SweepShape sweep_shape(int n_total) { return plan_sweep_shape(n_total, g.sm_count, g.wgroups, g.sweep_team); }
// n_total: the walkers in flight together (the call's, when it is cut into slices on several streams)
void launch_sweep(cudaStream_t st, int first, int n, long long n_steps, int trace_walker, mgpu_step_trace *d_trace, int n_total)
{
    const SweepShape sh = sweep_shape(n_total);
    const int threads = sh.per_cta * sh.nt, nb = (n + sh.per_cta - 1) / sh.per_cta;
    // phase alignment (see k_sweep): measured best sets -- top of the step + before the guest pass (r03e: 27.65 M moves/s against
    // 26.9 M with the k-space barrier as well; teams r03f: 14.2 against 13.5 M at 512 walkers); none in the large triclinic cells
    // of configs[4], where a pass over 17 664 framework atoms dwarfs everything a barrier could align (2.19 against 2.05 M)
    const int ps = g.phase_sync >= 0 ? g.phase_sync : (g.h.triclinic ? 0 : (sh.nt == 32 ? (1 | 4 | 16) : (1 | 4)));
#define SWEEP(TRI, NT, SM) k_sweep<TRI, NT><<<nb, threads, SM, st>>>(first, n, n_steps, g.natom_max, trace_walker, d_trace, g.d_err, ps)
    if (sh.nt == MGPU_TEAM) { if (g.h.triclinic) SWEEP(true, MGPU_TEAM, g.smem_team); else SWEEP(false, MGPU_TEAM, g.smem_team); }
    else if (sh.nt == MGPU_TEAM2) { if (g.h.triclinic) SWEEP(true, MGPU_TEAM2, g.smem_team2); else SWEEP(false, MGPU_TEAM2, g.smem_team2); }
    else { if (g.h.triclinic) SWEEP(true, 32, g.smem8); else SWEEP(false, 32, g.smem8); }
#undef SWEEP
}
// host mirror of the molecule counts (argument checks of the host-driven trials): kept current by commits and
// mgpu_set_count, re-read from the device after anything that changes counts there (sweeps, record loads)
int ensure_counts()
{
    if (g.counts_valid) return 0;
    g.counts.resize((size_t)g.h.n_walkers * MGPU_MAX_RES);
    CK(cudaMemcpyAsync(g.counts.data(), g.h.count, sizeof(int32_t) * g.counts.size(), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    g.counts_valid = true;
    return 0;
}
} // namespace

// =====================================================================================
extern "C" {

const char *mgpu_last_error(void) { return g_err.c_str(); }

void mgpu_finalize(void)
{
    if (g.d_blob) { cudaFree(g.d_blob); g.d_blob = nullptr; g.blob_cap = 0; }
    if (g.d_rec_off) { cudaFree(g.d_rec_off); g.d_rec_off = nullptr; g.rec_cap = 0; }
    for (int i = 0; i < Context::NPIPE; ++i) {
        if (g.pstream[i]) { cudaStreamSynchronize(g.pstream[i]); cudaStreamDestroy(g.pstream[i]); cudaEventDestroy(g.pevent[i]); }
        if (g.p_in[i]) cudaFree(g.p_in[i]);
        if (g.p_out[i]) cudaFree(g.p_out[i]);
        if (g.p_off_in[i]) { cudaFree(g.p_off_in[i]); cudaFree(g.p_off_out[i]); cudaFreeHost(g.ph_off_in[i]); cudaFreeHost(g.ph_off_out[i]); }
    }
    if (g.stream) cudaStreamSynchronize(g.stream);
    for (void *p : g.allocs) cudaFree(p);
    g.allocs.clear();
    if (g.h_task_i) { cudaStreamSynchronize(g.stream); cudaFreeHost(g.h_task_i); cudaFreeHost(g.h_task_com); cudaFreeHost(g.h_task_off); cudaFreeHost(g.h_task_out); cudaFreeHost(g.h_accept); cudaFreeHost(g.hc_walker); cudaFreeHost(g.hc_accept); }
    if (g.h_scratch) cudaFreeHost(g.h_scratch);
    if (g.ev0) cudaEventDestroy(g.ev0);
    if (g.ev1) cudaEventDestroy(g.ev1);
    if (g.stream) cudaStreamDestroy(g.stream);
    g = Context{};
}

int mgpu_device_info(char *name, int name_len, int *sm_count, double *mem_gb)
{
    int dev = 0;
    CK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev));
    if (name && name_len > 0) { std::strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (mem_gb) *mem_gb = (double)p.totalGlobalMem / 1e9;
    return 0;
}

int mgpu_init(const mgpu_system *sys)
{
    if (g.ready) mgpu_finalize();
    if (!sys) return fail("mgpu_init: null system");
    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev == 0)
        return fail(std::string("mgpu_init: no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e0));
    if (sys->device < 0 || sys->device >= ndev) return fail("mgpu_init: device ordinal out of range");
    if (sys->nres < 1 || sys->nres > MGPU_MAX_RES) return fail("mgpu_init: number of residue types out of range");
    if (sys->n_walkers < 1) return fail("mgpu_init: n_walkers must be >= 1");
    CK(cudaSetDevice(sys->device));
    g.device = sys->device;
    CK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&g.ev0));
    CK(cudaEventCreate(&g.ev1));
    DevSys &h = g.h;
    std::memset(&h, 0, sizeof h);

    // ---- box: prepare_simulation_box, geometry_utils.f90:19-35,148-200,293-361 ----
    double M[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { M[i][j] = sys->matrix[i * 3 + j]; h.H[i * 3 + j] = M[i][j]; }
    for (int d = 0; d < 3; ++d) h.lo[d] = sys->lo[d];
    {
        const double od[6] = { M[0][1], M[0][2], M[1][0], M[1][2], M[2][0], M[2][1] };
        double mx = 0.0; for (double v : od) mx = std::fmax(mx, std::fabs(v));
        h.triclinic = mx > MGPU_ERR_TOL;
    }
    double metrics[3];
    if (box_geometry(M, metrics, h.Hinv, &h.volume)) return fail("Error: Determinant fell into denormal/underflow range");
    for (int d = 0; d < 3; ++d) { h.L[d] = M[d][d]; h.invL[d] = 1.0 / M[d][d]; }
    h.tri_nrel = 0; g.tri_listed = 0;
    if (h.triclinic) {
        TriPlan tp;
        plan_triclinic(M, tp);
        h.tri_nrel = tp.nrel; h.tri_safe2 = tp.safe2;
        for (int k = 0; k < std::max(tp.nrel, 0); ++k) {
            for (int i = 0; i < 3; ++i) { h.tri_rel[k][i] = tp.rel[k][i]; h.tri_m[k][i] = tp.m[k][i]; }
            h.tri_len2[k] = tp.len2[k]; g.tri_req[k] = tp.req[k];
        }
        for (int d = 0; d < 3; ++d) h.tri_eps[d] = tp.eps[d];
    }
    h.tri_lower = (M[0][1] == 0.0 && M[0][2] == 0.0 && M[1][2] == 0.0) ? 1 : 0;
    // ---- Ewald: setup_ewald, prepare_utils.f90:110-226 ----
    double rc = sys->real_space_cutoff;
    if (rc > metrics[0] || rc > metrics[1] || rc > metrics[2]) rc = std::fmin(metrics[0], std::fmin(metrics[1], metrics[2])) / 2.0;
    const double tol = std::fmin(std::fabs(sys->ewald_tolerance), 0.5);
    const double screen = std::sqrt(std::fabs(std::log(tol * rc)));
    const double alpha = std::sqrt(std::fabs(std::log(tol * rc * screen))) / rc;
    const double t2 = 2.0 * screen * alpha;
    const double fprec = std::sqrt(-std::log(tol * rc * (t2 * t2)));
    h.rc = rc; h.rc2 = rc * rc; h.alpha = alpha;
    if (h.triclinic && h.tri_nrel > 0) {
        // A listed lattice vector C m can only shorten a rounded vector t when |t| > |C m| / 2.  If every pair that far
        // apart is beyond the LJ cutoff and its erfc-Coulomb term is so small that ALL such pairs of a trial together
        // stay below 1e-12 kcal/mol (bound: targets x max|q|^2 x erfc(alpha r)/r at r = min|C m|/2), the candidates
        // cannot change any energy at the 1e-9 parity tolerance and are not evaluated at all: the rounded image is used.
        double r_safe = 1e300;
        for (int k = 0; k < h.tri_nrel; ++k) r_safe = std::fmin(r_safe, 0.5 * std::sqrt(h.tri_len2[k]));
        double qmax = 0.0; long long targets = 0;
        for (int r = 0; r < sys->nres; ++r) {
            const mgpu_residue &R = sys->residues[r];
            for (int a = 0; a < R.natom; ++a) qmax = std::fmax(qmax, std::fabs(R.charges[a]));
            targets += (long long)R.natom * (R.is_active ? R.capacity : R.nmol);
        }
        const double bound = (double)targets * MGPU_MAX_SITES * qmax * qmax * std::erfc(alpha * r_safe) / r_safe * EPS0_INV_real();
        if (rc <= r_safe && bound < 1.0e-12) { g.tri_listed = h.tri_nrel; h.tri_nrel = 0; }
    }
    plan_triclinic_gate(h.triclinic != 0, h.tri_nrel, g.tri_req, h.tri_eps, h.tri_thr_hi, h.tri_lut);
    h.tri_thr_min = std::min(h.tri_thr_hi[0], std::min(h.tri_thr_hi[1], h.tri_thr_hi[2]));
    for (int d = 0; d < 3; ++d) h.kmax[d] = f_nint(0.25 + metrics[d] * alpha * fprec / PI);
    h.kmax_max = std::max(h.kmax[0], std::max(h.kmax[1], h.kmax[2]));
    h.eps0_inv_real = EPS0_INV_real(); h.twopi = TWOPI; h.overlap = OVERLAP();
    h.beta = 1 / (KB_kcalmol() * sys->temperature);

    // ---- k-vectors: ewald_kvectors.f90:24-65,130-149 ----
    std::vector<int32_t> kx, ky, kz; std::vector<double> ffW, k2m;
    {
        double kvm[3][3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) kvm[i][j] = TWOPI * h.Hinv[i * 3 + j];
        const double alpha_squared = alpha * alpha;
        for (int ix = 0; ix <= h.kmax[0]; ++ix)
            for (int iy = -h.kmax[1]; iy <= h.kmax[1]; ++iy)
                for (int iz = -h.kmax[2]; iz <= h.kmax[2]; ++iz) {
                    if (ix == 0 && iy == 0 && iz == 0) continue;
                    const double a = (double)ix / (double)h.kmax[0], b = (double)iy / (double)h.kmax[1], c = (double)iz / (double)h.kmax[2];
                    const double k2 = a * a + b * b + c * c;
                    if (!((std::fabs(k2) >= MGPU_ERR_TOL) && (k2 <= 1.0))) continue;
                    double kv[3];
                    for (int i = 0; i < 3; ++i) kv[i] = (double)ix * kvm[i][0] + (double)iy * kvm[i][1] + (double)iz * kvm[i][2];
                    const double k2mag = kv[0] * kv[0] + kv[1] * kv[1] + kv[2] * kv[2];
                    const double W = std::exp(-k2mag / (4.0 * alpha_squared)) / k2mag;
                    const double ff = (ix == 0) ? 1.0 : 2.0;
                    kx.push_back(ix); ky.push_back(iy); kz.push_back(iz); k2m.push_back(k2mag); ffW.push_back(ff * W);
                }
    }
    h.nk = (int)kx.size();

    // ---- residues ----
    h.nres = sys->nres; h.ntypes = sys->ntypes; h.nactive = 0;
    g.natom.clear(); g.active.clear(); g.cap.clear(); g.fug.clear();
    g.natom_max = 1;
    int64_t stride = 0;
    int n_host = 0;
    double self_host = 0.0;
    for (int r = 0; r < sys->nres; ++r) {
        const mgpu_residue &R = sys->residues[r];
        if (R.natom < 1) return fail("mgpu_init: residue with no atoms");
        if (R.is_active && R.natom > MGPU_MAX_SITES) return fail("mgpu_init: guest residue has more than MGPU_MAX_SITES atoms");
        if (R.is_active && (R.capacity < R.nmol || R.capacity < 1 || R.capacity > MGPU_NB_MAX_MOLECULE))
            return fail("mgpu_init: capacity must be in [max(nmol,1), NB_MAX_MOLECULE]");
        h.natom[r] = R.natom; h.active[r] = R.is_active ? 1 : 0; h.cap[r] = R.is_active ? R.capacity : 0;
        g.natom.push_back(R.natom); g.active.push_back(R.is_active); g.cap.push_back(R.capacity); g.fug.push_back(R.fugacity);
        double es = 0.0;                                      // ewald_self_energy_single_mol, ewald_energy.f90:177-205
        for (int a = 0; a < R.natom; ++a) {
            const double q = R.charges[a];
            if (R.types[a] < 0 || R.types[a] >= sys->ntypes) return fail("mgpu_init: atom type id out of range");
            if (std::fabs(q) < MGPU_ERR_TOL) continue;
            es = es - alpha / SQRTPI() * (q * q);
        }
        es = es * EPS0_INV_real();
        h.e_self[r] = es;
        if (R.is_active) {
            h.active_list[h.nactive++] = r;
            if (R.natom > g.natom_max) g.natom_max = R.natom;
            for (int a = 0; a < R.natom; ++a) { h.charge[r][a] = R.charges[a]; h.type[r][a] = R.types[a]; }
            for (int a = 0; a < R.natom; ++a) if (R.charges[a] != 0.0) h.qlist[r][h.nq[r]++] = (int8_t)a;
            h.goff[r] = stride;
            stride += (int64_t)(3 + 3 * R.natom + 2) * R.capacity;      // com, offsets, framework-energy cache rows
            // prepare_monte_carlo, prepare_utils.f90:231-259
            const double mass = R.mass * G_TO_KG / NA();
            double lam = H_PLANCK / std::sqrt(TWOPI * mass * KB * sys->temperature);
            h.lambda[r] = lam * M_TO_A;
        } else {
            h.host_count[r] = R.nmol;
            n_host += R.nmol * R.natom;
            const double e = es * R.nmol;                     // evaluate_ewald_self_energy, self_energy_utils.f90:24-47
            self_host = self_host + e;
        }
    }
    if (h.nactive == 0) return fail("mgpu_init: no active (guest) residue");
    h.self_host_total = self_host;
    h.coord_stride = stride; h.n_host = n_host; h.n_walkers = sys->n_walkers;
    h.p_trans = sys->p_translation; h.p_rot = sys->p_rotation; h.p_swap = sys->p_swap;
    h.p_insdel = sys->p_insertion_deletion; h.p_widom = sys->p_widom;
    h.tstep = sys->translation_step; h.rstep = sys->rotation_step_angle;
    h.use_hcache = 1;
    for (int a = 0; a < MGPU_MAX_SITES; ++a) h.iota[a] = (int8_t)a;

    // ---- static arrays ----
    std::vector<double4> hx(n_host ? n_host : 1); std::vector<int32_t> ht(n_host ? n_host : 1), hm(n_host ? n_host : 1); std::vector<double> hq(n_host ? n_host : 1);
    {
        int k = 0, molid = 0;
        for (int r = 0; r < sys->nres; ++r) {
            const mgpu_residue &R = sys->residues[r];
            if (R.is_active) continue;
            for (int m = 0; m < R.nmol; ++m, ++molid)
                for (int a = 0; a < R.natom; ++a, ++k) {
                    const double *c = R.com + (size_t)m * 3, *o = R.offset + ((size_t)m * R.natom + a) * 3;
                    const double q = R.charges[a];
                    hx[k] = make_double4(c[0] + o[0], c[1] + o[1], c[2] + o[2], std::fabs(q) < MGPU_ERR_TOL ? 0.0 : q);   // geometry_utils.f90:235-241; :157
                    hq[k] = q; ht[k] = R.types[a]; hm[k] = molid;
                    if (std::fabs(q) >= MGPU_ERR_TOL) ++h.n_host_charged;
                }
        }
    }
    if (n_host > 64) {
        // Static framework atoms in an order that makes the 32 consecutive atoms of a warp iteration one small region (see
        // order_framework): sums are order-independent up to rounding; the host-host constant and S_host use the same
        // permuted arrays.
        std::vector<double> xyz((size_t)n_host * 3);
        for (int k = 0; k < n_host; ++k) { xyz[(size_t)k * 3] = hx[k].x; xyz[(size_t)k * 3 + 1] = hx[k].y; xyz[(size_t)k * 3 + 2] = hx[k].z; }
        std::vector<int> order;
        int cols[3];
        order_framework(n_host, xyz.data(), h.lo, h.Hinv, metrics, h.triclinic && h.tri_nrel > 0, g.tri_req, h.tri_nrel, order, cols);
        std::vector<double4> hx2(n_host); std::vector<int32_t> ht2(n_host), hm2(n_host); std::vector<double> hq2(n_host);
        for (int k = 0; k < n_host; ++k) { const int o = order[k]; hx2[k] = hx[o]; ht2[k] = ht[o]; hm2[k] = hm[o]; hq2[k] = hq[o]; }
        hx.swap(hx2); ht.swap(ht2); hm.swap(hm2); hq.swap(hq2);
    }
    double4 *d_hx; int32_t *d_ht, *d_hm; double *d_eps, *d_sig, *d_ffW, *d_Shost, *d_hq, *d_ctab; int32_t *d_kx, *d_ky, *d_kz;
    const size_t nk1 = h.nk ? h.nk : 1, nt2 = (size_t)sys->ntypes * sys->ntypes;
    if (dalloc(&d_hx, hx.size()) || dalloc(&d_ht, ht.size()) || dalloc(&d_hm, hm.size()) || dalloc(&d_eps, nt2) || dalloc(&d_sig, nt2) ||
        dalloc(&d_hq, hq.size()) || dalloc(&d_ctab, ((size_t)MGPU_TAB_MAXOCT * (1 << MGPU_TAB_K) + 1) * MGPU_TAB_ROW) ||
        dalloc(&d_ffW, nk1) || dalloc(&d_kx, nk1) || dalloc(&d_ky, nk1) || dalloc(&d_kz, nk1) || dalloc(&d_Shost, 2 * nk1)) return 1;
    CK(cudaMemcpy(d_hx, hx.data(), sizeof(double4) * hx.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ht, ht.data(), sizeof(int32_t) * ht.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_hm, hm.data(), sizeof(int32_t) * hm.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_hq, hq.data(), sizeof(double) * hq.size(), cudaMemcpyHostToDevice));
    {
        std::vector<double2> hxy(hx.size()), hzq(hx.size());
        for (size_t i = 0; i < hx.size(); ++i) { hxy[i] = make_double2(hx[i].x, hx[i].y); hzq[i] = make_double2(hx[i].z, hx[i].w); }
        double2 *d_xy, *d_zq;
        if (dalloc(&d_xy, hxy.size()) || dalloc(&d_zq, hzq.size())) return 1;
        CK(cudaMemcpy(d_xy, hxy.data(), sizeof(double2) * hxy.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_zq, hzq.data(), sizeof(double2) * hzq.size(), cudaMemcpyHostToDevice));
        h.host_xy = d_xy; h.host_zq = d_zq;
        h.host_fxy = d_xy; h.host_fzq = d_zq;
        if (h.triclinic && n_host > 0) {
            // framework in fractional coordinates wrapped into [0, 1) (Hinv = transposed inverse: f_d = sum_j Hinv[j][d] x_j)
            std::vector<double2> fxy(hx.size()), fzq(hx.size());
            for (size_t i = 0; i < (size_t)n_host; ++i) {
                const double x[3] = { hx[i].x, hx[i].y, hx[i].z };
                double f[3];
                for (int d = 0; d < 3; ++d) { f[d] = h.Hinv[0 * 3 + d] * x[0] + h.Hinv[1 * 3 + d] * x[1] + h.Hinv[2 * 3 + d] * x[2]; f[d] -= std::floor(f[d]); }
                fxy[i] = make_double2(f[0], f[1]); fzq[i] = make_double2(f[2], hx[i].w);
            }
            double2 *d_fxy, *d_fzq;
            if (dalloc(&d_fxy, fxy.size()) || dalloc(&d_fzq, fzq.size())) return 1;
            CK(cudaMemcpy(d_fxy, fxy.data(), sizeof(double2) * fxy.size(), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(d_fzq, fzq.data(), sizeof(double2) * fzq.size(), cudaMemcpyHostToDevice));
            h.host_fxy = d_fxy; h.host_fzq = d_fzq;
        }
        // classes of guest atoms with respect to the framework
        std::vector<char> type_present(sys->ntypes, 0);
        bool host_charged = false;
        for (int i = 0; i < n_host; ++i) { type_present[ht[i]] = 1; host_charged = host_charged || hx[i].w != 0.0; }
        for (int r = 0; r < sys->nres; ++r) {
            const mgpu_residue &R = sys->residues[r];
            if (!R.is_active) continue;
            for (int a = 0; a < R.natom; ++a) {
                bool lj = false;
                for (int t = 0; t < sys->ntypes; ++t)
                    if (type_present[t] && sys->epsilon[(size_t)R.types[a] * sys->ntypes + t] != 0.0 && sys->sigma[(size_t)R.types[a] * sys->ntypes + t] != 0.0) lj = true;
                const bool co = host_charged && std::fabs(R.charges[a]) >= MGPU_ERR_TOL;
                const int mode = (lj ? 1 : 0) | (co ? 2 : 0);
                h.hl_list[r][mode][h.hl_n[r][mode]++] = (int8_t)a;
            }
        }
    }
    {
        // classes of a probe residue's atoms against atom b of guest residue g (guest passes of K1)
        const size_t nl = (size_t)MGPU_MAX_RES * MGPU_MAX_RES * MGPU_MAX_SITES * 4;
        std::vector<int8_t> gn(nl, 0), gl(nl * MGPU_MAX_SITES, 0);
        for (int ri = 0; ri < sys->nres; ++ri) {
            const mgpu_residue &Ri = sys->residues[ri];
            if (!Ri.is_active) continue;
            for (int gg = 0; gg < sys->nres; ++gg) {
                const mgpu_residue &Rg = sys->residues[gg];
                if (!Rg.is_active) continue;
                for (int b = 0; b < Rg.natom; ++b)
                    for (int a = 0; a < Ri.natom; ++a) {
                        const size_t ti = (size_t)Ri.types[a] * sys->ntypes + Rg.types[b];
                        const bool lj = sys->epsilon[ti] != 0.0 && sys->sigma[ti] != 0.0;
                        const bool co = std::fabs(Ri.charges[a]) >= MGPU_ERR_TOL && std::fabs(Rg.charges[b]) >= MGPU_ERR_TOL;
                        const int mode = (lj ? 1 : 0) | (co ? 2 : 0);
                        const size_t gi = (((size_t)ri * MGPU_MAX_RES + gg) * MGPU_MAX_SITES + b) * 4 + mode;
                        gl[gi * MGPU_MAX_SITES + gn[gi]++] = (int8_t)a;
                    }
            }
        }
        int8_t *d_gn, *d_gl;
        if (dalloc(&d_gn, gn.size()) || dalloc(&d_gl, gl.size())) return 1;
        CK(cudaMemcpy(d_gn, gn.data(), gn.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_gl, gl.data(), gl.size(), cudaMemcpyHostToDevice));
        h.gl_n = d_gn; h.gl_list = d_gl;
    }
    {
        // pairwise_lj_energy (pairwise_energy_utils.f90:131-136): 4 eps ((sigma/r)^12 - (sigma/r)^6) = A / s^6 - B / s^3
        std::vector<double> A(nt2), B(nt2);
        for (size_t i = 0; i < nt2; ++i) {
            const double sg = sys->sigma[i], s2 = sg * sg, s6 = s2 * s2 * s2;
            B[i] = 4.0 * sys->epsilon[i] * s6;
            A[i] = B[i] * s6;
        }
        CK(cudaMemcpy(d_eps, A.data(), sizeof(double) * nt2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_sig, B.data(), sizeof(double) * nt2, cudaMemcpyHostToDevice));
    }
    {
        // Coulomb table over every minimum-image distance the box allows (orthorhombic: half the body
        // diagonal; triclinic: the 27-image search is bounded by the same sum of cell-vector lengths)
        double r_hi = h.triclinic ? 0.5 * (metrics[0] + metrics[1] + metrics[2]) + 1.0
                                  : 0.5 * std::sqrt(metrics[0] * metrics[0] + metrics[1] * metrics[1] + metrics[2] * metrics[2]) + 1.0;
        const double r_zero = MGPU_TAB_XCUT / alpha;         // erfc(7)/r < 1e-24: below the rounding of any sum it enters
        h.s_zero = r_zero * r_zero;
        if (r_hi > r_zero) r_hi = r_zero;
        std::vector<double> tab;
        mgpu_build_coulomb_table(alpha, MGPU_TAB_RLO, r_hi, &g.tab_emin, &g.tab_noct, tab);
        // the hot loops send everything beyond the last interval to an all-zero row, so the table has to reach
        // the largest distance that still matters (the box limit or alpha r = 7, whichever is smaller)
        if (std::ldexp(1.0, g.tab_emin + g.tab_noct) < r_hi * r_hi)
            return fail("mgpu_init: real-space Coulomb table would need more than MGPU_TAB_MAXOCT octaves (very small alpha / huge cutoff)");
        tab.resize(tab.size() + MGPU_TAB_ROW, 0.0);           // the all-zero closing row
        CK(cudaMemcpy(d_ctab, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice));
        h.tab_ibase = (1023 + g.tab_emin) << MGPU_TAB_K;
        h.tab_nint = g.tab_noct << MGPU_TAB_K;
        h.tab_hi_lo = (1023 + g.tab_emin) << 20;
        h.ctab = d_ctab;
    }
    if (h.nk) {
        CK(cudaMemcpy(d_ffW, ffW.data(), sizeof(double) * h.nk, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_kx, kx.data(), sizeof(int32_t) * h.nk, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_ky, ky.data(), sizeof(int32_t) * h.nk, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_kz, kz.data(), sizeof(int32_t) * h.nk, cudaMemcpyHostToDevice));
    }
    h.host_xyzq = d_hx; h.host_type = d_ht; h.host_qraw = d_hq; h.ljA = d_eps; h.ljB = d_sig; h.ffW = d_ffW; h.kx = d_kx; h.ky = d_ky; h.kz = d_kz; h.S_host = d_Shost;

    // ---- per-walker arrays ----
    const size_t W = sys->n_walkers;
    if (dalloc(&h.coords, W * (size_t)stride) || dalloc(&h.count, W * MGPU_MAX_RES) || dalloc(&h.S, W * 4 * nk1) || dalloc(&h.cur, W) ||
        dalloc(&h.energy, W * 6) || dalloc(&h.mu, W * MGPU_MAX_RES) || dalloc(&h.rng, W * 4) || dalloc(&h.step, W * 2) || dalloc(&h.rmax2, MGPU_MAX_RES) || dalloc(&h.counters, W * 12) ||
        dalloc(&h.widom_w, W * MGPU_MAX_RES) || dalloc(&h.widom_n, W * MGPU_MAX_RES) || dalloc(&h.avg, W * MGPU_MAX_RES * 4) ||
        dalloc(&h.trial, W) || dalloc(&h.pair_count, 4) || dalloc(&g.d_err, 1) || dalloc(&g.d_scratch, 64) || dalloc(&g.d_geom, 3 + 3 * MGPU_MAX_SITES)) return 1;
    CK(cudaMallocHost(&g.h_scratch, sizeof(double) * 64));
    CK(cudaMemset(h.cur, 0, sizeof(int32_t) * W));
    CK(cudaMemset(h.S, 0, sizeof(double) * W * 4 * nk1));
    CK(cudaMemset(h.energy, 0, sizeof(double) * W * 6));
    CK(cudaMemset(h.counters, 0, sizeof(long long) * W * 12));
    CK(cudaMemset(h.widom_w, 0, sizeof(double) * W * MGPU_MAX_RES));
    CK(cudaMemset(h.widom_n, 0, sizeof(long long) * W * MGPU_MAX_RES));
    CK(cudaMemset(h.avg, 0, sizeof(double) * W * MGPU_MAX_RES * 4));
    CK(cudaMemset(h.trial, 0, sizeof(MgpuTrial) * W));
    CK(cudaMemset(g.d_err, 0, sizeof(int32_t)));
    CK(cudaMemset(h.pair_count, 0, sizeof(unsigned long long) * 4));
    {
        // one walker image, replicated
        std::vector<double> img(stride ? stride : 1, 0.0); std::vector<int32_t> cnt(MGPU_MAX_RES, 0); std::vector<double> mu(MGPU_MAX_RES, 0.0);
        for (int r = 0; r < sys->nres; ++r) {
            const mgpu_residue &R = sys->residues[r];
            if (!R.is_active) continue;
            cnt[r] = R.nmol;
            mu[r] = (R.fugacity >= 0.0) ? std::log(R.fugacity) / h.beta : R.chemical_potential;   // prepare_utils.f90:245-247
            double *com = img.data() + h.goff[r], *off = com + 3 * (size_t)R.capacity;
            for (int m = 0; m < R.nmol; ++m) {
                for (int d = 0; d < 3; ++d) com[(size_t)d * R.capacity + m] = R.com[(size_t)m * 3 + d];
                for (int a = 0; a < R.natom; ++a)
                    for (int d = 0; d < 3; ++d) off[((size_t)a * 3 + d) * R.capacity + m] = R.offset[((size_t)m * R.natom + a) * 3 + d];
            }
        }
        std::vector<double> all(W * (size_t)(stride ? stride : 1)); std::vector<int32_t> allc(W * MGPU_MAX_RES); std::vector<double> allmu(W * MGPU_MAX_RES);
        for (size_t w = 0; w < W; ++w) {
            if (stride) std::memcpy(all.data() + w * stride, img.data(), sizeof(double) * stride);
            std::memcpy(allc.data() + w * MGPU_MAX_RES, cnt.data(), sizeof(int32_t) * MGPU_MAX_RES);
            std::memcpy(allmu.data() + w * MGPU_MAX_RES, mu.data(), sizeof(double) * MGPU_MAX_RES);
        }
        if (stride) CK(cudaMemcpy(h.coords, all.data(), sizeof(double) * W * stride, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h.count, allc.data(), sizeof(int32_t) * allc.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h.mu, allmu.data(), sizeof(double) * allmu.size(), cudaMemcpyHostToDevice));
    }
    {
        std::vector<double> st(W * 2);
        for (size_t w = 0; w < W; ++w) { st[2 * w] = sys->translation_step; st[2 * w + 1] = sys->rotation_step_angle; }
        CK(cudaMemcpy(h.step, st.data(), sizeof(double) * st.size(), cudaMemcpyHostToDevice));
    }
    {
        // bound of |offset|^2 per residue type (screen of the "nothing" lists, guest_loops)
        g.rmax2.assign(MGPU_MAX_RES, 0.0);
        for (int r = 0; r < sys->nres; ++r) {
            const mgpu_residue &R = sys->residues[r];
            if (!R.is_active) continue;
            for (int m = 0; m < R.nmol; ++m)
                for (int a = 0; a < R.natom; ++a) {
                    const double *o = R.offset + ((size_t)m * R.natom + a) * 3;
                    g.rmax2[r] = std::fmax(g.rmax2[r], o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
                }
        }
        CK(cudaMemcpy(h.rmax2, g.rmax2.data(), sizeof(double) * MGPU_MAX_RES, cudaMemcpyHostToDevice));
    }
    g.dirty.assign(W, 0);
    g.pending.assign(W, 0);
    g.pend_kind.assign(W, 0); g.pend_res.assign(W, 0); g.pend_res2.assign(W, 0);
    g.batch_stamp.assign(W, 0); g.batch_id = 0;
    g.counts_valid = false;
    g.commit_inflight = false;

    // ---- shared-memory budgets ----
    g.smem1 = smem_bytes(h.ntypes, h.tab_nint, h.kmax_max, g.natom_max, 1, 1, true);
    g.wgroups = MGPU_WGROUPS;
    while (g.wgroups > 4 && smem_bytes(h.ntypes, h.tab_nint, h.kmax_max, g.natom_max, g.wgroups, MGPU_TAB_REP, false) > 227 * 1024) g.wgroups -= 4;   // whole quartets / teams
    g.smem8 = smem_bytes(h.ntypes, h.tab_nint, h.kmax_max, g.natom_max, g.wgroups, MGPU_TAB_REP, false);
    g.smem_buildS = sizeof(double2) * (size_t)MGPU_STILE * 3 * (h.kmax_max + 1);
    const size_t smem_max = std::max(g.smem8, g.smem_buildS);
    if (smem_max > 227 * 1024) return fail("mgpu_init: kmax / ntypes / molecule size need more than 227 KB of shared memory per CTA");
    CK(cudaDeviceGetAttribute(&g.sm_count, cudaDevAttrMultiProcessorCount, g.device));
#define SET_SMEM(k, b) CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(b)))
    SET_SMEM((k_trial<false, 32>), g.smem8); SET_SMEM((k_trial<true, 32>), g.smem8);
    SET_SMEM((k_trial<false, MGPU_BLOCK>), g.smem1); SET_SMEM((k_trial<true, MGPU_BLOCK>), g.smem1);
    g.smem_team = smem_bytes(h.ntypes, h.tab_nint, h.kmax_max, g.natom_max, std::max(1, g.wgroups * 32 / MGPU_TEAM), MGPU_TAB_REP, true);
    SET_SMEM((k_sweep<false, 32>), g.smem8); SET_SMEM((k_sweep<true, 32>), g.smem8);
    SET_SMEM((k_sweep<false, MGPU_TEAM>), g.smem_team); SET_SMEM((k_sweep<true, MGPU_TEAM>), g.smem_team);
    g.smem_team2 = smem_bytes(h.ntypes, h.tab_nint, h.kmax_max, g.natom_max, std::max(1, g.wgroups * 32 / MGPU_TEAM2), MGPU_TAB_REP, true);
    SET_SMEM((k_sweep<false, MGPU_TEAM2>), g.smem_team2); SET_SMEM((k_sweep<true, MGPU_TEAM2>), g.smem_team2);
    SET_SMEM(k_total_energy<false>, g.smem1); SET_SMEM(k_total_energy<true>, g.smem1);
    SET_SMEM(k_pair_molecule<false>, g.smem1); SET_SMEM(k_pair_molecule<true>, g.smem1);
    SET_SMEM(k_widom_batch<false>, g.smem8); SET_SMEM(k_widom_batch<true>, g.smem8);
    SET_SMEM(k_build_S, g.smem_buildS);
#undef SET_SMEM
    {
        // shared-memory carve-out of the warp-per-walker kernels: the smallest supported configuration that holds one
        // CTA (dynamic + 1 KB reserved), so that the rest of the 256 KB stays L1 for the framework atoms.
        // MGPU_CARVEOUT (percent of the maximum) overrides it for experiments.
        int pct = (int)((g.smem8 + 1024 + 2048) * 100 / (228 * 1024)) + 1;
        if (pct > 100) pct = 100;
        if (const char *e = std::getenv("MGPU_CARVEOUT")) pct = std::atoi(e);
        cudaFuncSetAttribute((k_sweep<false, 32>), cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute((k_sweep<true, 32>), cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(k_widom_batch<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(k_widom_batch<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute((k_trial<false, 32>), cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute((k_trial<true, 32>), cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    {
        int nb = 1;
        if (h.triclinic) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_widom_batch<true>, 32 * g.wgroups, g.smem8));
        else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_widom_batch<false>, 32 * g.wgroups, g.smem8));
        g.ctas_per_sm = nb > 0 ? nb : 1;
    }

    if (upload_sys()) return 1;
    CK(cudaStreamSynchronize(g.stream));

    // ---- static host terms: S_host(k) and host-host pair energy ----
    if (h.nk) {
        k_build_S<<<(h.nk + MGPU_BLOCK - 1) / MGPU_BLOCK, MGPU_BLOCK, g.smem_buildS, g.stream>>>(0, d_Shost, 0);
        CK(cudaGetLastError());
    }
    {
        const int nb = 64;
        double *d_part; if (dalloc(&d_part, 2 * nb)) return 1;
        if (h.triclinic) k_host_host<true><<<nb, MGPU_BLOCK, 0, g.stream>>>(d_hm, d_part);
        else k_host_host<false><<<nb, MGPU_BLOCK, 0, g.stream>>>(d_hm, d_part);
        CK(cudaGetLastError());
        std::vector<double> part(2 * nb);
        CK(cudaMemcpyAsync(part.data(), d_part, sizeof(double) * 2 * nb, cudaMemcpyDeviceToHost, g.stream));
        CK(cudaStreamSynchronize(g.stream));
        double lj = 0.0, co = 0.0;
        for (int b = 0; b < nb; ++b) { lj += part[2 * b]; co += part[2 * b + 1]; }
        h.hh_lj = lj; h.hh_coul = co * EPS0_INV_real();
    }
    if (upload_sys()) return 1;
    CK(cudaStreamSynchronize(g.stream));
    g.ready = true;
    if (mgpu_seed(12345u)) return 1;
    return rebuild(0, (int)W);
}

// ---- queries --------------------------------------------------------------------------
int mgpu_get_ewald(double *alpha, int32_t kmax[3], int32_t *nkvec, double *rc_used)
{
    NEED_READY();
    if (alpha) *alpha = g.h.alpha;
    if (kmax) for (int d = 0; d < 3; ++d) kmax[d] = g.h.kmax[d];
    if (nkvec) *nkvec = g.h.nk;
    if (rc_used) *rc_used = g.h.rc;
    return 0;
}
int mgpu_get_kvectors(int32_t *kx, int32_t *ky, int32_t *kz, double *k2, double *ffw)
{
    NEED_READY();
    const size_t n = g.h.nk;
    if (kx) CK(cudaMemcpy(kx, g.h.kx, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
    if (ky) CK(cudaMemcpy(ky, g.h.ky, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
    if (kz) CK(cudaMemcpy(kz, g.h.kz, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
    if (ffw) CK(cudaMemcpy(ffw, g.h.ffW, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (k2) {
        std::vector<int32_t> x(n), y(n), z(n);
        CK(cudaMemcpy(x.data(), g.h.kx, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(y.data(), g.h.ky, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(z.data(), g.h.kz, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; ++i) {
            double kv[3];
            for (int r = 0; r < 3; ++r)
                kv[r] = (double)x[i] * (TWOPI * g.h.Hinv[r * 3 + 0]) + (double)y[i] * (TWOPI * g.h.Hinv[r * 3 + 1]) + (double)z[i] * (TWOPI * g.h.Hinv[r * 3 + 2]);
            k2[i] = kv[0] * kv[0] + kv[1] * kv[1] + kv[2] * kv[2];
        }
    }
    return 0;
}
int mgpu_get_box(double matrix[9], double reciprocal[9], double *volume, int32_t *is_triclinic)
{
    NEED_READY();
    if (matrix) std::memcpy(matrix, g.h.H, sizeof g.h.H);
    if (reciprocal) std::memcpy(reciprocal, g.h.Hinv, sizeof g.h.Hinv);
    if (volume) *volume = g.h.volume;
    if (is_triclinic) *is_triclinic = g.h.triclinic;
    return 0;
}
int mgpu_get_launch_info(int32_t *walkers_per_cta, int64_t *smem_bytes_per_cta, int32_t *sm_count)
{
    NEED_READY();
    if (walkers_per_cta) *walkers_per_cta = g.wgroups;
    if (smem_bytes_per_cta) *smem_bytes_per_cta = (int64_t)g.smem8;
    if (sm_count) *sm_count = g.sm_count;
    return 0;
}
int mgpu_get_triclinic_candidates(int32_t *n) { NEED_READY(); *n = g.h.tri_nrel; return 0; }
int mgpu_plan_sweep_shape(int32_t n_walkers, int32_t sm_count, int32_t warps_per_cta, int32_t forced,
                          int32_t *threads_per_walker, int32_t *walkers_per_cta)
{
    if (n_walkers < 1 || sm_count < 1 || warps_per_cta < 1) return fail("mgpu_plan_sweep_shape: arguments must be positive");
    const SweepShape sh = plan_sweep_shape(n_walkers, sm_count, warps_per_cta, forced);
    if (threads_per_walker) *threads_per_walker = sh.nt;
    if (walkers_per_cta) *walkers_per_cta = sh.per_cta;
    return 0;
}
int mgpu_plan_framework_order(const double *matrix, const double *lo, int32_t n_atoms, const double *xyz, int32_t *order, int32_t *columns)
{
    if (!matrix || !lo || !xyz || !order || n_atoms < 1) return fail("mgpu_plan_framework_order: bad argument");
    double M[3][3], metrics[3], hinv[9], vol;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[i][j] = matrix[i * 3 + j];
    if (box_geometry(M, metrics, hinv, &vol)) return fail("Error: Determinant fell into denormal/underflow range");
    const double od[6] = { M[0][1], M[0][2], M[1][0], M[1][2], M[2][0], M[2][1] };
    double mx = 0.0; for (double v : od) mx = std::fmax(mx, std::fabs(v));
    TriPlan tp;
    if (mx > MGPU_ERR_TOL) plan_triclinic(M, tp);
    std::vector<int> ord;
    int cols[3];
    order_framework(n_atoms, xyz, lo, hinv, metrics, mx > MGPU_ERR_TOL && tp.nrel > 0, tp.req, tp.nrel, ord, cols);
    for (int k = 0; k < n_atoms; ++k) order[k] = ord[k];
    if (columns) for (int d = 0; d < 3; ++d) columns[d] = cols[d];
    return 0;
}
int mgpu_plan_triclinic(const double *matrix, int32_t *n_vectors, double *vectors, int32_t *coefficients, int32_t *faces,
                        int32_t *thr_hi, uint32_t *lut)
{
    if (!matrix || !n_vectors) return fail("mgpu_plan_triclinic: null argument");
    double M[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[i][j] = matrix[i * 3 + j];
    TriPlan tp;
    plan_triclinic(M, tp);
    *n_vectors = tp.nrel;
    for (int k = 0; k < std::max(tp.nrel, 0); ++k) {
        for (int i = 0; i < 3; ++i) { if (vectors) vectors[k * 3 + i] = tp.rel[k][i]; if (coefficients) coefficients[k * 3 + i] = (int32_t)tp.m[k][i]; }
        if (faces) faces[k] = tp.req[k];
    }
    int32_t th[3]; uint32_t lt[16];
    plan_triclinic_gate(true, tp.nrel, tp.req, tp.eps, th, lt);
    for (int d = 0; d < 3; ++d) if (thr_hi) thr_hi[d] = th[d];
    for (int f = 0; f < 16; ++f) if (lut) lut[f] = lt[f];
    return 0;
}
int mgpu_get_sweep_shape(int32_t n_walkers, int32_t *threads_per_walker, int32_t *walkers_per_cta)
{
    NEED_READY();
    const SweepShape sh = sweep_shape(n_walkers < 1 ? 1 : n_walkers);
    if (threads_per_walker) *threads_per_walker = sh.nt;
    if (walkers_per_cta) *walkers_per_cta = sh.per_cta;
    return 0;
}
int mgpu_get_thermo(int32_t res, double *beta, double *lambda, double *mu_walker0)
{
    NEED_READY();
    if (res < 0 || res >= g.h.nres) return fail("mgpu_get_thermo: residue out of range");
    if (beta) *beta = g.h.beta;
    if (lambda) *lambda = g.h.lambda[res];
    if (mu_walker0) CK(cudaMemcpy(mu_walker0, g.h.mu + res, sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

// ---- per-walker state -----------------------------------------------------------------
int mgpu_set_molecule(int32_t w, int32_t res, int32_t mol, const double com[3], const double *offset)
{
    NEED_READY();
    if (check_walker(w) || check_guest(res)) return 1;
    if (mol < 0 || mol >= g.h.cap[res]) return fail("mgpu_set_molecule: molecule index out of range");
    const int cap = g.h.cap[res], na = g.h.natom[res];
    double *base = g.h.coords + (int64_t)w * g.h.coord_stride + g.h.goff[res];
    // strided scatter through a 2-D copy: element e of {com[3], off[na*3]} goes to base[e*cap + mol]
    std::vector<double> tmp(3 + 3 * na);
    for (int d = 0; d < 3; ++d) tmp[d] = com[d];
    for (int e = 0; e < 3 * na; ++e) tmp[3 + e] = offset[e];
    CK(cudaMemcpy2DAsync(base + mol, sizeof(double) * cap, tmp.data(), sizeof(double), sizeof(double), 3 + 3 * na, cudaMemcpyHostToDevice, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    double r2 = 0.0;
    for (int a = 0; a < na; ++a) r2 = std::fmax(r2, offset[3 * a] * offset[3 * a] + offset[3 * a + 1] * offset[3 * a + 1] + offset[3 * a + 2] * offset[3 * a + 2]);
    if (r2 > g.rmax2[res]) {                                 // keep the |offset|^2 bound of the residue type valid
        CK(cudaMemcpy(g.rmax2.data(), g.h.rmax2, sizeof(double) * MGPU_MAX_RES, cudaMemcpyDeviceToHost));   // device side may have grown too
        if (r2 > g.rmax2[res]) { g.rmax2[res] = r2; CK(cudaMemcpy(g.h.rmax2 + res, &r2, sizeof(double), cudaMemcpyHostToDevice)); }
    }
    g.dirty[w] = 1;
    return 0;
}
int mgpu_get_molecule(int32_t w, int32_t res, int32_t mol, double com[3], double *offset)
{
    NEED_READY();
    if (check_walker(w) || check_guest(res)) return 1;
    if (mol < 0 || mol >= g.h.cap[res]) return fail("mgpu_get_molecule: molecule index out of range");
    const int cap = g.h.cap[res], na = g.h.natom[res];
    const double *base = g.h.coords + (int64_t)w * g.h.coord_stride + g.h.goff[res];
    std::vector<double> tmp(3 + 3 * na);
    CK(cudaMemcpy2DAsync(tmp.data(), sizeof(double), base + mol, sizeof(double) * cap, sizeof(double), 3 + 3 * na, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    for (int d = 0; d < 3; ++d) com[d] = tmp[d];
    for (int e = 0; e < 3 * na; ++e) offset[e] = tmp[3 + e];
    return 0;
}
int mgpu_set_count(int32_t w, int32_t res, int32_t n)
{
    NEED_READY();
    if (check_walker(w) || check_guest(res)) return 1;
    if (n < 0 || n > g.h.cap[res]) return fail("mgpu_set_count: count out of range");
    CK(cudaMemcpy(g.h.count + (int64_t)w * MGPU_MAX_RES + res, &n, sizeof n, cudaMemcpyHostToDevice));
    if (g.counts_valid) g.counts[(size_t)w * MGPU_MAX_RES + res] = n;
    g.dirty[w] = 1;
    return 0;
}
int mgpu_get_count(int32_t w, int32_t res, int32_t *n)
{
    NEED_READY();
    if (check_walker(w)) return 1;
    if (res < 0 || res >= g.h.nres) return fail("mgpu_get_count: residue out of range");
    if (!g.h.active[res]) { *n = g.h.host_count[res]; return 0; }
    CK(cudaMemcpy(n, g.h.count + (int64_t)w * MGPU_MAX_RES + res, sizeof *n, cudaMemcpyDeviceToHost));
    return 0;
}
int mgpu_set_chemical_potential(int32_t w, int32_t res, double mu)
{
    NEED_READY();
    if (check_walker(w) || check_guest(res)) return 1;
    CK(cudaMemcpy(g.h.mu + (int64_t)w * MGPU_MAX_RES + res, &mu, sizeof mu, cudaMemcpyHostToDevice));
    return 0;
}
int mgpu_set_chemical_potentials(int32_t first, int32_t n, int32_t res, const double *mu)
{
    NEED_READY();
    if (n <= 0) return 0;
    if (check_walker(first) || check_walker(first + n - 1) || check_guest(res)) return 1;
    if (!mu) return fail("mgpu_set_chemical_potentials: null array");
    // element (w, res) of the [W][MGPU_MAX_RES] array: a strided scatter in one copy
    CK(cudaMemcpy2D(g.h.mu + (int64_t)first * MGPU_MAX_RES + res, sizeof(double) * MGPU_MAX_RES, mu, sizeof(double), sizeof(double), n, cudaMemcpyHostToDevice));
    return 0;
}
int mgpu_get_counts(int32_t first, int32_t n, int32_t res, int32_t *counts)
{
    NEED_READY();
    if (n <= 0) return 0;
    if (check_walker(first) || check_walker(first + n - 1) || check_guest(res)) return 1;
    CK(cudaMemcpy2D(counts, sizeof(int32_t), g.h.count + (int64_t)first * MGPU_MAX_RES + res, sizeof(int32_t) * MGPU_MAX_RES, sizeof(int32_t), n, cudaMemcpyDeviceToHost));
    return 0;
}
int mgpu_set_fugacity(int32_t w, int32_t res, double f)
{
    if (!(f >= 0.0)) return fail("mgpu_set_fugacity: fugacity must be >= 0");
    return mgpu_set_chemical_potential(w, res, std::log(f) / g.h.beta);
}
int mgpu_get_Ak(int32_t w, double *re_im)
{
    NEED_READY();
    if (check_walker(w) || ensure_clean(w)) return 1;
    int32_t cur = 0;
    CK(cudaMemcpy(&cur, g.h.cur + w, sizeof cur, cudaMemcpyDeviceToHost));
    const size_t nk = g.h.nk;
    std::vector<double> tmp(2 * nk);
    CK(cudaMemcpy(tmp.data(), g.h.S + ((int64_t)w * 2 + cur) * 2 * nk, sizeof(double) * 2 * nk, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < nk; ++i) { re_im[2 * i] = tmp[i]; re_im[2 * i + 1] = tmp[nk + i]; }
    return 0;
}
int mgpu_get_energy(int32_t w, double out[6])
{
    NEED_READY();
    if (check_walker(w) || ensure_clean(w)) return 1;
    CK(cudaMemcpy(out, g.h.energy + (int64_t)w * 6, sizeof(double) * 6, cudaMemcpyDeviceToHost));
    return 0;
}

int mgpu_set_option(int32_t option, int32_t value)
{
    NEED_READY();
    if (option == MGPU_OPT_PHASE_SYNC) { g.phase_sync = (value == 1 || value < 0) ? -1 : (value & 31); return 0; }   // bits: 1 top of step, 2 before the evaluation, 4 before the guest pass, 8 before k-space
    if (option == MGPU_OPT_SWEEP_TEAM) { g.sweep_team = value < 0 ? -1 : (value > 2 ? 1 : value); return 0; }
    if (option == MGPU_OPT_BLOCK_SLICES) { g.block_slices = value < 1 ? 1 : (value > Context::NPIPE ? (int)Context::NPIPE : value); return 0; }
    if (option == MGPU_OPT_HOST_CACHE) {
        g.h.use_hcache = value ? 1 : 0;
        if (upload_sys()) return 1;
        CK(cudaStreamSynchronize(g.stream));
        return 0;
    }
    return fail("mgpu_set_option: unknown option");
}

// ---- energies -------------------------------------------------------------------------
int mgpu_total_energy(int32_t w, double out[6])
{
    NEED_READY();
    if (check_walker(w)) return 1;
    if (rebuild(w, 1)) return 1;
    if (out) CK(cudaMemcpy(out, g.h.energy + (int64_t)w * 6, sizeof(double) * 6, cudaMemcpyDeviceToHost));
    return 0;
}

static int upload_geom(int res, const double *com, const double *offset, const double **d_out)
{
    *d_out = nullptr;
    if (!com || !offset) return 0;
    const int na = g.h.natom[res];
    std::vector<double> tmp(3 + 3 * na);
    for (int d = 0; d < 3; ++d) tmp[d] = com[d];
    for (int e = 0; e < 3 * na; ++e) tmp[3 + e] = offset[e];
    CK(cudaMemcpyAsync(g.d_geom, tmp.data(), sizeof(double) * tmp.size(), cudaMemcpyHostToDevice, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    *d_out = g.d_geom;
    return 0;
}

int mgpu_pairwise_energy_for_molecule(int32_t w, int32_t res, int32_t mol, int32_t skip, const double *com, const double *offset,
                                      double *e_nc, double *e_c)
{
    NEED_READY();
    if (check_walker(w) || check_guest(res)) return 1;
    if (mol < 0 || mol >= g.h.cap[res]) return fail("pairwise_energy_for_molecule: molecule index out of range");
    const double *dg;
    if (upload_geom(res, com, offset, &dg)) return 1;
    if (g.h.triclinic) k_pair_molecule<true><<<1, MGPU_BLOCK, g.smem1, g.stream>>>(w, res, mol, skip, dg, g.d_scratch, g.natom_max);
    else k_pair_molecule<false><<<1, MGPU_BLOCK, g.smem1, g.stream>>>(w, res, mol, skip, dg, g.d_scratch, g.natom_max);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(g.h_scratch, g.d_scratch, sizeof(double) * 2, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    *e_nc = g.h_scratch[0]; *e_c = g.h_scratch[1];
    return 0;
}
int mgpu_ewald_self_energy_single_mol(int32_t res, double *e)
{
    NEED_READY();
    if (res < 0 || res >= g.h.nres) return fail("ewald_self_energy_single_mol: residue out of range");
    *e = g.h.e_self[res];
    return 0;
}
int mgpu_intra_res_real_coulomb_energy(int32_t w, int32_t res, int32_t mol, const double *com, const double *offset, double *e)
{
    NEED_READY();
    if (check_walker(w) || check_guest(res)) return 1;
    const double *dg;
    if (upload_geom(res, com, offset, &dg)) return 1;
    if (g.h.triclinic) k_intra<true><<<1, 32, 0, g.stream>>>(w, res, mol, dg, g.d_scratch);
    else k_intra<false><<<1, 32, 0, g.stream>>>(w, res, mol, dg, g.d_scratch);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(g.h_scratch, g.d_scratch, sizeof(double), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    *e = g.h_scratch[0];
    return 0;
}
int mgpu_reciprocal_ewald_energy(int32_t w, double *e)
{
    double E[6];
    if (mgpu_get_energy(w, E)) return 1;
    *e = E[MGPU_E_RECIP];
    return 0;
}

int mgpu_trial_batch(int32_t n, const int32_t *walker, const int32_t *res, const int32_t *mol, const int32_t *kind,
                     const double *com, const double *offset, double *e_old, double *e_new)
{
    NEED_READY_NOSYNC();
    if (n <= 0) return 0;
    if (ensure_task_cap(n) || ensure_counts()) return 1;
    // one trial per walker: a second task on a walker would race with the first on the walker's spare S(k) buffer and
    // on its pending-trial record, and a trial on top of a pending one would be committed onto flipped buffers
    ++g.batch_id;
    if (g.batch_id == 0) { std::fill(g.batch_stamp.begin(), g.batch_stamp.end(), 0u); g.batch_id = 1; }
    for (int t = 0; t < n; ++t) {
        if (check_walker(walker[t]) || check_guest(res[t])) return 1;
        if (g.pending[walker[t]]) return fail("mgpu_trial_batch: walker already has a pending trial (commit or roll it back first)");
        if (g.batch_stamp[walker[t]] == g.batch_id) return fail("mgpu_trial_batch: a walker is listed twice in one batch");
        g.batch_stamp[walker[t]] = g.batch_id;
    }
    for (int t = 0; t < n; ++t) {
        const int k = kind[t] & 0xff, res2 = kind[t] >> 8;
        if (k < MGPU_KIND_MOVE || k > MGPU_KIND_SWAP || (k != MGPU_KIND_SWAP && res2 != 0)) return fail("mgpu_trial_batch: unknown kind");
        if (mol[t] < 0 || mol[t] >= g.h.cap[res[t]]) return fail("Trying to insert / move a molecule with an index beyond the walker's capacity");
        const int cnt = g.counts[(size_t)walker[t] * MGPU_MAX_RES + res[t]];
        if (k == MGPU_KIND_CREATE ? (mol[t] != cnt) : (mol[t] >= cnt))
            return fail(k == MGPU_KIND_CREATE ? "mgpu_trial_batch: a creation trial goes into slot N + 1 (mol must equal the current count)"
                                              : "mgpu_trial_batch: molecule index beyond the walker's current count");
        if (k == MGPU_KIND_SWAP) {
            if (check_guest(res2)) return 1;
            if (res2 == res[t]) return fail("mgpu_trial_batch: swap needs two different residue types");
            if (g.counts[(size_t)walker[t] * MGPU_MAX_RES + res2] >= g.h.cap[res2]) return fail("Trying to insert a molecule beyond the walker's capacity (NB_MAX_MOLECULE analogue)");
        }
        if (ensure_clean(walker[t])) return 1;
        g.h_task_i[4 * t] = walker[t]; g.h_task_i[4 * t + 1] = res[t]; g.h_task_i[4 * t + 2] = mol[t]; g.h_task_i[4 * t + 3] = kind[t];
    }
    if (com) std::memcpy(g.h_task_com, com, sizeof(double) * 3 * n); else std::memset(g.h_task_com, 0, sizeof(double) * 3 * n);
    if (offset) std::memcpy(g.h_task_off, offset, sizeof(double) * 3 * MGPU_MAX_SITES * n); else std::memset(g.h_task_off, 0, sizeof(double) * 3 * MGPU_MAX_SITES * n);
    const bool latency = n < g.sm_count;
    if (latency) {
        // few trials: no staging copies at all -- the kernel dereferences the page-locked host buffers (zero copy)
        TaskArrays T{ reinterpret_cast<const int4 *>(g.h_task_i), g.h_task_com, g.h_task_off, g.h_task_out };
        Timer tm("trial");
        if (g.h.triclinic) k_trial<true, MGPU_BLOCK><<<n, MGPU_BLOCK, g.smem1, g.stream>>>(T, n, g.natom_max);
        else k_trial<false, MGPU_BLOCK><<<n, MGPU_BLOCK, g.smem1, g.stream>>>(T, n, g.natom_max);
        tm.stop();                                   // synchronises on the kernel's end (and on any commit queued before it)
        g.commit_inflight = false;
        CK(cudaGetLastError());
    } else {
        CK(cudaMemcpyAsync(g.d_task_i, g.h_task_i, sizeof(int32_t) * 4 * n, cudaMemcpyHostToDevice, g.stream));
        CK(cudaMemcpyAsync(g.d_task_com, g.h_task_com, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, g.stream));
        CK(cudaMemcpyAsync(g.d_task_off, g.h_task_off, sizeof(double) * 3 * MGPU_MAX_SITES * n, cudaMemcpyHostToDevice, g.stream));
        TaskArrays T{ reinterpret_cast<const int4 *>(g.d_task_i), g.d_task_com, g.d_task_off, g.d_task_out };
        Timer tm("trial");
        // one warp per task, spread over every SM: ceil(n / SMs) warps per CTA, at most wgroups
        int grp = (n + g.sm_count - 1) / g.sm_count;
        if (grp > g.wgroups) grp = g.wgroups;
        const int nb = (n + grp - 1) / grp;
        const size_t sm = smem_bytes(g.h.ntypes, g.h.tab_nint, g.h.kmax_max, g.natom_max, grp, MGPU_TAB_REP, false);
        if (g.h.triclinic) k_trial<true, 32><<<nb, 32 * grp, sm, g.stream>>>(T, n, g.natom_max);
        else k_trial<false, 32><<<nb, 32 * grp, sm, g.stream>>>(T, n, g.natom_max);
        tm.stop();
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(g.h_task_out, g.d_task_out, sizeof(double) * 12 * n, cudaMemcpyDeviceToHost, g.stream));
        CK(cudaStreamSynchronize(g.stream));
        g.commit_inflight = false;
    }
    for (int t = 0; t < n; ++t) {
        g.pending[walker[t]] = 1;
        g.pend_kind[walker[t]] = kind[t] & 0xff; g.pend_res[walker[t]] = res[t]; g.pend_res2[walker[t]] = kind[t] >> 8;
        if (e_old) std::memcpy(e_old + 6 * (size_t)t, g.h_task_out + 12 * (size_t)t, sizeof(double) * 6);
        if (e_new) std::memcpy(e_new + 6 * (size_t)t, g.h_task_out + 12 * (size_t)t + 6, sizeof(double) * 6);
    }
    return 0;
}

int mgpu_commit_batch(int32_t n, const int32_t *walker, const int32_t *accept)
{
    NEED_READY_NOSYNC();
    if (n <= 0) return 0;
    if (ensure_task_cap(n)) return 1;
    for (int t = 0; t < n; ++t) {
        if (check_walker(walker[t])) return 1;
        if (!g.pending[walker[t]]) return fail("mgpu_commit: commit without a pending trial");
    }
    if (g.commit_inflight) { CK(cudaStreamSynchronize(g.stream)); g.commit_inflight = false; }   // the argument buffers are being read
    for (int t = 0; t < n; ++t) {
        const int w = walker[t];
        g.hc_walker[t] = w; g.hc_accept[t] = accept[t] ? 1 : 0; g.pending[w] = 0;
        if (accept[t] && g.counts_valid) {                    // update_counts, mirrored on the host
            int32_t *c = g.counts.data() + (size_t)w * MGPU_MAX_RES;
            if (g.pend_kind[w] == MGPU_KIND_CREATE) c[g.pend_res[w]] += 1;
            else if (g.pend_kind[w] == MGPU_KIND_DELETE) c[g.pend_res[w]] -= 1;
            else if (g.pend_kind[w] == MGPU_KIND_SWAP) { c[g.pend_res[w]] -= 1; c[g.pend_res2[w]] += 1; }
        }
    }
    if (n < g.sm_count) {
        // latency path: arguments read from page-locked host memory, no synchronisation -- the next energy call (or any
        // getter, all of which synchronise the stream) orders after it
        k_commit<<<n, 128, 0, g.stream>>>(g.hc_walker, g.hc_accept, g.d_err);
        CK(cudaGetLastError());
        g.commit_inflight = true;
        return 0;
    }
    CK(cudaMemcpyAsync(g.d_task_i, g.hc_walker, sizeof(int32_t) * n, cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemcpyAsync(g.d_accept, g.hc_accept, sizeof(int32_t) * n, cudaMemcpyHostToDevice, g.stream));
    k_commit<<<n, 128, 0, g.stream>>>(g.d_task_i, g.d_accept, g.d_err);
    CK(cudaGetLastError());
    g.commit_inflight = true;                    // queued like the latency path: the host mirror above already ruled out the one device-side error
    return 0;
}

int mgpu_old_energy(int32_t w, int32_t res, int32_t mol, int32_t kind, double out[6])
{
    NEED_READY();
    if (check_walker(w) || check_guest(res) || ensure_clean(w)) return 1;
    for (int i = 0; i < 6; ++i) out[i] = 0.0;
    double E[6];
    CK(cudaMemcpy(E, g.h.energy + (int64_t)w * 6, sizeof E, cudaMemcpyDeviceToHost));
    out[MGPU_E_RECIP] = E[MGPU_E_RECIP];
    if (kind != MGPU_KIND_CREATE) {
        if (mgpu_pairwise_energy_for_molecule(w, res, mol, 1, nullptr, nullptr, &out[MGPU_E_NON_COULOMB], &out[MGPU_E_COULOMB])) return 1;
        if (kind == MGPU_KIND_DELETE) {
            out[MGPU_E_SELF] = g.h.e_self[res];
            if (mgpu_intra_res_real_coulomb_energy(w, res, mol, nullptr, nullptr, &out[MGPU_E_INTRA])) return 1;
        }
    }
    out[MGPU_E_TOTAL] = out[0] + out[1] + out[2] + out[3] + out[4];
    return 0;
}
int mgpu_new_energy(int32_t w, int32_t res, int32_t mol, int32_t kind, const double *com, const double *offset, double out[6])
{
    NEED_READY();
    if (check_walker(w) || check_guest(res)) return 1;
    if (kind != MGPU_KIND_DELETE && (!com || !offset)) return fail("mgpu_new_energy: trial geometry required for MOVE / CREATE");
    double off16[MGPU_MAX_SITES * 3] = { 0 };
    if (offset) std::memcpy(off16, offset, sizeof(double) * 3 * g.h.natom[res]);
    double c3[3] = { 0, 0, 0 };
    if (com) std::memcpy(c3, com, sizeof c3);
    return mgpu_trial_batch(1, &w, &res, &mol, &kind, c3, off16, nullptr, out);
}
int mgpu_swap_energy(int32_t w, int32_t res_old, int32_t mol_old, int32_t res_new, const double *com, const double *offset,
                     double e_old[6], double e_new[6])
{
    NEED_READY();
    if (check_walker(w) || check_guest(res_old) || check_guest(res_new)) return 1;
    if (!com || !offset) return fail("mgpu_swap_energy: geometry of the new molecule required");
    double off16[MGPU_MAX_SITES * 3] = { 0 };
    std::memcpy(off16, offset, sizeof(double) * 3 * g.h.natom[res_new]);
    int32_t kind = MGPU_KIND_SWAP | (res_new << 8);
    return mgpu_trial_batch(1, &w, &res_old, &mol_old, &kind, com, off16, e_old, e_new);
}
int mgpu_commit(int32_t w) { int32_t one = 1; return mgpu_commit_batch(1, &w, &one); }
int mgpu_rollback(int32_t w) { int32_t zero = 0; return mgpu_commit_batch(1, &w, &zero); }

// ---- device-resident MC ---------------------------------------------------------------
int mgpu_seed(uint64_t seed)
{
    NEED_READY();
    const size_t W = g.h.n_walkers;
    std::vector<uint64_t> st(W * 4);
    for (size_t w = 0; w < W; ++w) {
        uint64_t z = seed + 104729ull * (uint64_t)w;       // seed + 104729*(i-1), like seed_rng's stride (random_utils.f90:66)
        for (int i = 0; i < 4; ++i) { z += MGPU_GOLDEN; st[w * 4 + i] = splitmix64_mix(z); }
    }
    CK(cudaMemcpy(g.h.rng, st.data(), sizeof(uint64_t) * st.size(), cudaMemcpyHostToDevice));
    return 0;
}
int mgpu_get_rng_state(int32_t w, uint64_t st[4])
{
    NEED_READY();
    if (check_walker(w)) return 1;
    CK(cudaMemcpy(st, g.h.rng + (int64_t)w * 4, sizeof(uint64_t) * 4, cudaMemcpyDeviceToHost));
    return 0;
}
int mgpu_sweep(int32_t first, int32_t n, int64_t n_steps, int32_t trace_walker, mgpu_step_trace *trace)
{
    NEED_READY();
    if (n <= 0 || n_steps <= 0) return 0;
    if (check_walker(first) || check_walker(first + n - 1)) return 1;
    for (int w = first; w < first + n; ++w) {
        if (g.pending[w]) return fail("mgpu_sweep: a walker has a pending trial (commit or roll it back first)");
        if (ensure_clean(w)) return 1;
    }
    g.counts_valid = false;
    mgpu_step_trace *d_trace = nullptr;
    if (trace) CK(cudaMalloc(&d_trace, sizeof(mgpu_step_trace) * n_steps));
    Timer tm("sweep");
    launch_sweep(g.stream, first, n, n_steps, trace_walker, d_trace, n);
    tm.stop();
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) { if (d_trace) cudaFree(d_trace); return fail(std::string("k_sweep: ") + cudaGetErrorString(le)); }
    if (trace) {
        cudaMemcpyAsync(trace, d_trace, sizeof(mgpu_step_trace) * n_steps, cudaMemcpyDeviceToHost, g.stream);
        cudaStreamSynchronize(g.stream);
        cudaFree(d_trace);
    }
    return check_err_flag("mgpu_sweep");
}
// ---- walker records: host state <-> device state ----------------------------------------------
static int ensure_rec_cap(size_t n_walkers, size_t n_doubles)
{
    if (n_walkers + 1 > g.rec_cap) {
        if (g.d_rec_off) cudaFree(g.d_rec_off);
        g.rec_cap = n_walkers + 1;
        CK(cudaMalloc(&g.d_rec_off, sizeof(long long) * g.rec_cap));
    }
    if (n_doubles > g.blob_cap) {
        if (g.d_blob) cudaFree(g.d_blob);
        g.blob_cap = n_doubles + n_doubles / 8;
        CK(cudaMalloc(&g.d_blob, sizeof(double) * g.blob_cap));
    }
    return 0;
}
// Validation of a batch of records on the host, BEFORE anything is uploaded or written: offsets start at 0, every
// record is at least as long as its fixed part, its counts are integers inside the capacity, and its length field,
// its offsets and its counts agree.  A malformed batch leaves the engine untouched.
static int validate_records(int n, const double *blob, const int64_t *offsets)
{
    if (offsets[0] != 0) return fail("mgpu_load_walkers: offsets[0] must be 0");
    const int64_t lmin = MGPU_REC_HDR + 2 * (int64_t)g.h.nk;
    for (int i = 0; i < n; ++i) {
        const int64_t len = offsets[i + 1] - offsets[i];
        if (len < lmin) return fail("mgpu_load_walkers: malformed walker record (length / counts / capacity)");
        const double *rec = blob + offsets[i];
        int64_t L = lmin;
        for (int r = 0; r < g.h.nres; ++r) {
            if (!g.h.active[r]) continue;
            const double c = rec[1 + r];
            if (!(c >= 0.0 && c <= (double)g.h.cap[r]) || c != std::floor(c)) return fail("mgpu_load_walkers: malformed walker record (length / counts / capacity)");
            const int64_t stored = c > 0.0 ? (int64_t)c : 1;
            L += stored * (3 + 3 * g.h.natom[r] + 2);
        }
        if (rec[0] != (double)L || len != L) return fail("mgpu_load_walkers: malformed walker record (length / counts / capacity)");
    }
    return 0;
}
void *mgpu_host_alloc(size_t bytes) { void *p = nullptr; return cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess ? p : nullptr; }
void mgpu_host_free(void *p) { if (p) cudaFreeHost(p); }
int64_t mgpu_record_doubles_max(void)
{
    if (!g.ready) return 0;
    int64_t L = MGPU_REC_HDR + 2 * (int64_t)g.h.nk;
    for (int r = 0; r < g.h.nres; ++r) if (g.h.active[r]) L += (int64_t)g.h.cap[r] * (3 + 3 * g.h.natom[r] + 2);
    return L;
}
int mgpu_save_walkers(int32_t first, int32_t n, double *blob, int64_t capacity, int64_t *offsets)
{
    NEED_READY();
    if (n <= 0) return 0;
    if (!blob || !offsets) return fail("mgpu_save_walkers: null buffer");
    if (check_walker(first) || check_walker(first + n - 1)) return 1;
    for (int w = first; w < first + n; ++w) if (ensure_clean(w)) return 1;
    if (ensure_rec_cap(n, 0)) return 1;
    k_record_len<<<(n + 255) / 256, 256, 0, g.stream>>>(first, n, g.d_rec_off);
    std::vector<long long> len(n);
    CK(cudaMemcpyAsync(len.data(), g.d_rec_off, sizeof(long long) * n, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    offsets[0] = 0;
    for (int i = 0; i < n; ++i) offsets[i + 1] = offsets[i] + len[i];
    if (offsets[n] > capacity) return fail("mgpu_save_walkers: buffer too small, need " + std::to_string((long long)offsets[n]) + " doubles");
    if (ensure_rec_cap(n, (size_t)offsets[n])) return 1;
    static_assert(sizeof(long long) == sizeof(int64_t), "int64");
    CK(cudaMemcpyAsync(g.d_rec_off, offsets, sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, g.stream));
    Timer tm("pack");
    k_pack<<<(n + 7) / 8, 256, 0, g.stream>>>(first, n, g.d_rec_off, g.d_blob);
    tm.stop();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(blob, g.d_blob, sizeof(double) * offsets[n], cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    g.d2h_bytes += (long long)sizeof(double) * offsets[n] + (long long)sizeof(long long) * n;
    g.h2d_bytes += (long long)sizeof(long long) * (n + 1);
    return 0;
}
int mgpu_load_walkers(int32_t first, int32_t n, const double *blob, const int64_t *offsets)
{
    NEED_READY();
    if (n <= 0) return 0;
    if (!blob || !offsets) return fail("mgpu_load_walkers: null buffer");
    if (check_walker(first) || check_walker(first + n - 1)) return 1;
    if (validate_records(n, blob, offsets)) return 1;
    if (ensure_rec_cap(n, (size_t)offsets[n])) return 1;
    CK(cudaMemcpyAsync(g.d_rec_off, offsets, sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemcpyAsync(g.d_blob, blob, sizeof(double) * offsets[n], cudaMemcpyHostToDevice, g.stream));
    g.h2d_bytes += (long long)sizeof(double) * offsets[n] + (long long)sizeof(long long) * (n + 1);
    Timer tm("unpack");
    k_unpack<<<(n + 7) / 8, 256, 0, g.stream>>>(first, n, g.d_rec_off, g.d_blob);
    tm.stop();
    CK(cudaGetLastError());
    for (int w = first; w < first + n; ++w) { g.dirty[w] = 0; g.pending[w] = 0; }   // the record carries S(k), energies and the cache rows
    g.counts_valid = false;
    return 0;
}
// mgpu_block, pipelined: the walkers are cut into slices of whole sweep CTAs, every slice gets its own stream and
// staging buffers and runs  records H2D -> k_unpack -> k_sweep -> record offsets (device scan) -> k_pack -> records D2H
// in stream order.  All slices are in flight at once: the copy engines move slice i+1 in and slice i-1 out while the
// SMs sweep slice i (a slice's CTAs start as soon as SMs are free, whichever stream they come from).  The only host
// round trip per slice is its record lengths (a few KB), needed to size the D2H copy and to place the slice in the
// caller's blob.  Walkers are independent, so the results equal the unpipelined load -> sweep -> save bit for bit.
static int block_pipelined(int32_t first, int32_t n, int64_t n_steps, const double *blob_in, const int64_t *offsets_in,
                           double *blob_out, int64_t capacity_out, int64_t *offsets_out)
{
    const int gran = g.wgroups;
    int S = std::min(g.block_slices, (int)Context::NPIPE);
    int per = ((n + S - 1) / S + gran - 1) / gran * gran;
    S = (n + per - 1) / per;
    const int64_t rec_max = mgpu_record_doubles_max();
    int64_t molmax = 0;
    for (int r = 0; r < g.h.nres; ++r) if (g.h.active[r]) molmax = std::max<int64_t>(molmax, 3 + 3 * g.h.natom[r] + 2);
    for (int i = 0; i < S; ++i) {
        if (!g.pstream[i]) { CK(cudaStreamCreateWithFlags(&g.pstream[i], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&g.pevent[i], cudaEventDisableTiming)); }
        const int m = std::min(per, n - i * per);
        if ((size_t)m + 1 > g.p_off_cap[i]) {
            if (g.p_off_in[i]) { cudaFree(g.p_off_in[i]); cudaFree(g.p_off_out[i]); cudaFreeHost(g.ph_off_in[i]); cudaFreeHost(g.ph_off_out[i]); }
            g.p_off_cap[i] = (size_t)per + 1;
            CK(cudaMalloc(&g.p_off_in[i], sizeof(long long) * g.p_off_cap[i]));
            CK(cudaMalloc(&g.p_off_out[i], sizeof(long long) * g.p_off_cap[i]));
            CK(cudaMallocHost(&g.ph_off_in[i], sizeof(long long) * g.p_off_cap[i]));
            CK(cudaMallocHost(&g.ph_off_out[i], sizeof(long long) * g.p_off_cap[i]));
        }
        const int64_t in_len = offsets_in[i * per + m] - offsets_in[i * per];
        if ((size_t)in_len > g.p_in_cap[i]) {
            if (g.p_in[i]) cudaFree(g.p_in[i]);
            g.p_in_cap[i] = (size_t)(in_len + in_len / 8);
            CK(cudaMalloc(&g.p_in[i], sizeof(double) * g.p_in_cap[i]));
        }
        // a walker gains at most one molecule per MC step
        int64_t out_bound = 0;
        for (int k = 0; k < m; ++k) out_bound += std::min<int64_t>(rec_max, (offsets_in[i * per + k + 1] - offsets_in[i * per + k]) + n_steps * molmax);
        if ((size_t)out_bound > g.p_out_cap[i]) {
            if (g.p_out[i]) cudaFree(g.p_out[i]);
            g.p_out_cap[i] = (size_t)out_bound;
            CK(cudaMalloc(&g.p_out[i], sizeof(double) * g.p_out_cap[i]));
        }
    }
    CK(cudaStreamSynchronize(g.stream));
    for (int i = 0; i < S; ++i) {                      // stage 1 of every slice, back to back
        cudaStream_t st = g.pstream[i];
        const int w0 = first + i * per, m = std::min(per, n - i * per);
        const int64_t b0 = offsets_in[i * per], in_len = offsets_in[i * per + m] - b0;
        for (int k = 0; k <= m; ++k) g.ph_off_in[i][k] = offsets_in[i * per + k] - b0;
        CK(cudaMemcpyAsync(g.p_off_in[i], g.ph_off_in[i], sizeof(long long) * (m + 1), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(g.p_in[i], blob_in + b0, sizeof(double) * in_len, cudaMemcpyHostToDevice, st));
        g.h2d_bytes += (long long)sizeof(double) * in_len + (long long)sizeof(long long) * (m + 1);
        k_unpack<<<(m + 7) / 8, 256, 0, st>>>(w0, m, g.p_off_in[i], g.p_in[i]);
        launch_sweep(st, w0, m, n_steps, -1, nullptr, n);
        k_record_offsets<<<1, 256, 0, st>>>(w0, m, g.p_off_out[i]);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(g.ph_off_out[i], g.p_off_out[i], sizeof(long long) * (m + 1), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(g.pevent[i], st));
    }
    int rc = 0;
    int64_t base = 0;
    offsets_out[0] = 0;
    for (int i = 0; i < S; ++i) {                      // stage 2 as the slices finish, in order
        cudaStream_t st = g.pstream[i];
        const int w0 = first + i * per, m = std::min(per, n - i * per);
        CK(cudaEventSynchronize(g.pevent[i]));
        const int64_t total = g.ph_off_out[i][m];
        if (!rc && (total > (int64_t)g.p_out_cap[i] || base + total > capacity_out))
            rc = fail("mgpu_block: output buffer too small, need more than " + std::to_string((long long)(base + total)) + " doubles");
        for (int k = 1; k <= m; ++k) offsets_out[i * per + k] = base + g.ph_off_out[i][k];
        if (!rc) {
            k_pack<<<(m + 7) / 8, 256, 0, st>>>(w0, m, g.p_off_out[i], g.p_out[i]);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(blob_out + base, g.p_out[i], sizeof(double) * total, cudaMemcpyDeviceToHost, st));
            g.d2h_bytes += (long long)sizeof(double) * total + (long long)sizeof(long long) * (m + 1);
        }
        base += total;
    }
    for (int i = 0; i < S; ++i) CK(cudaStreamSynchronize(g.pstream[i]));
    for (int w = first; w < first + n; ++w) { g.dirty[w] = 0; g.pending[w] = 0; }
    g.counts_valid = false;
    if (check_err_flag("mgpu_block")) return 1;
    return rc;
}

int mgpu_block(int32_t first, int32_t n, int64_t n_steps, const double *blob_in, const int64_t *offsets_in,
               double *blob_out, int64_t capacity_out, int64_t *offsets_out)
{
    NEED_READY();
    if (n <= 0) return 0;
    if (check_walker(first) || check_walker(first + n - 1)) return 1;
    if (blob_in && blob_out && offsets_in && offsets_out && n_steps > 0 && g.block_slices > 1 && n >= 2 * g.wgroups) {
        if (validate_records(n, blob_in, offsets_in)) return 1;
        return block_pipelined(first, n, n_steps, blob_in, offsets_in, blob_out, capacity_out, offsets_out);
    }
    if (blob_in && mgpu_load_walkers(first, n, blob_in, offsets_in)) return 1;
    if (mgpu_sweep(first, n, n_steps, -1, nullptr)) return 1;
    if (blob_out && mgpu_save_walkers(first, n, blob_out, capacity_out, offsets_out)) return 1;
    return 0;
}
int mgpu_get_traffic(int64_t *h2d_bytes, int64_t *d2h_bytes, int32_t reset)
{
    if (h2d_bytes) *h2d_bytes = g.h2d_bytes;
    if (d2h_bytes) *d2h_bytes = g.d2h_bytes;
    if (reset) { g.h2d_bytes = 0; g.d2h_bytes = 0; }
    return 0;
}
int mgpu_adjust_move_step_sizes(int32_t first, int32_t n)
{
    NEED_READY();
    if (n <= 0) return 0;
    if (check_walker(first) || check_walker(first + n - 1)) return 1;
    k_adjust_steps<<<(n + 127) / 128, 128, 0, g.stream>>>(first, n);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}
int mgpu_get_step_sizes(int32_t w, double out[2])
{
    NEED_READY();
    if (check_walker(w)) return 1;
    CK(cudaMemcpy(out, g.h.step + (int64_t)w * 2, sizeof(double) * 2, cudaMemcpyDeviceToHost));
    return 0;
}
int mgpu_set_step_sizes(int32_t w, double translation_step, double rotation_step_angle)
{
    NEED_READY();
    if (check_walker(w)) return 1;
    const double v[2] = { translation_step, rotation_step_angle };
    CK(cudaMemcpy(g.h.step + (int64_t)w * 2, v, sizeof v, cudaMemcpyHostToDevice));
    return 0;
}
int mgpu_get_counters(int32_t w, int64_t out[12])
{
    NEED_READY();
    if (check_walker(w)) return 1;
    CK(cudaMemcpy(out, g.h.counters + (int64_t)w * 12, sizeof(int64_t) * 12, cudaMemcpyDeviceToHost));
    return 0;
}
int mgpu_get_widom(int32_t w, int32_t res, double *sw, int64_t *ns)
{
    NEED_READY();
    if (check_walker(w) || check_guest(res)) return 1;
    long long n = 0;
    CK(cudaMemcpy(sw, g.h.widom_w + (int64_t)w * MGPU_MAX_RES + res, sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&n, g.h.widom_n + (int64_t)w * MGPU_MAX_RES + res, sizeof n, cudaMemcpyDeviceToHost));
    *ns = n;
    return 0;
}
int mgpu_get_averages(int32_t w, int32_t res, double out[4])
{
    NEED_READY();
    if (check_walker(w) || check_guest(res)) return 1;
    CK(cudaMemcpy(out, g.h.avg + ((int64_t)w * MGPU_MAX_RES + res) * 4, sizeof(double) * 4, cudaMemcpyDeviceToHost));
    return 0;
}
int mgpu_reset_averages(void)
{
    NEED_READY();
    const size_t W = g.h.n_walkers;
    CK(cudaMemset(g.h.avg, 0, sizeof(double) * W * MGPU_MAX_RES * 4));
    CK(cudaMemset(g.h.widom_w, 0, sizeof(double) * W * MGPU_MAX_RES));
    CK(cudaMemset(g.h.widom_n, 0, sizeof(long long) * W * MGPU_MAX_RES));
    CK(cudaMemset(g.h.counters, 0, sizeof(long long) * W * 12));
    return 0;
}

int mgpu_widom_batch(int32_t w, int32_t res, int64_t first_id, int64_t n, uint64_t seed, double *dE_out, double *sum_w, int64_t *n_ok)
{
    NEED_READY();
    if (check_walker(w) || check_guest(res) || ensure_clean(w)) return 1;
    if (n <= 0) { *sum_w = 0.0; *n_ok = 0; return 0; }
    // persistent warps: one insertion per warp at a time, grid = resident CTAs (or fewer for small n)
    long long nb = (n + g.wgroups - 1) / g.wgroups;
    const long long nb_max = (long long)g.sm_count * g.ctas_per_sm;
    if (nb > nb_max) nb = nb_max;
    const long long nw = nb * g.wgroups;
    double *d_dE = nullptr, *d_bw = nullptr; long long *d_bn = nullptr;
    if (dE_out) CK(cudaMalloc(&d_dE, sizeof(double) * n));
    CK(cudaMalloc(&d_bw, sizeof(double) * nw));
    CK(cudaMalloc(&d_bn, sizeof(long long) * nw));
    Timer tm("widom");
    if (g.h.triclinic) k_widom_batch<true><<<(unsigned)nb, 32 * g.wgroups, g.smem8, g.stream>>>(w, res, first_id, n, seed, d_dE, d_bw, d_bn, g.natom_max);
    else k_widom_batch<false><<<(unsigned)nb, 32 * g.wgroups, g.smem8, g.stream>>>(w, res, first_id, n, seed, d_dE, d_bw, d_bn, g.natom_max);
    tm.stop();
    cudaError_t le = cudaGetLastError();
    std::vector<double> bw(nw); std::vector<long long> bn(nw);
    if (le == cudaSuccess) le = cudaMemcpyAsync(bw.data(), d_bw, sizeof(double) * nw, cudaMemcpyDeviceToHost, g.stream);
    if (le == cudaSuccess) le = cudaMemcpyAsync(bn.data(), d_bn, sizeof(long long) * nw, cudaMemcpyDeviceToHost, g.stream);
    if (le == cudaSuccess && dE_out) le = cudaMemcpyAsync(dE_out, d_dE, sizeof(double) * n, cudaMemcpyDeviceToHost, g.stream);
    if (le == cudaSuccess) le = cudaStreamSynchronize(g.stream);
    if (d_dE) cudaFree(d_dE);
    cudaFree(d_bw); cudaFree(d_bn);
    if (le != cudaSuccess) return fail(std::string("mgpu_widom_batch: ") + cudaGetErrorString(le));
    double sw = 0.0; long long ok = 0;
    for (long long b = 0; b < nw; ++b) { sw += bw[b]; ok += bn[b]; }
    *sum_w = sw; *n_ok = ok;
    return 0;
}

// ---- measurement ----------------------------------------------------------------------
int mgpu_get_pair_counts(int64_t out[3])
{
    NEED_READY();
    CK(cudaMemcpy(out, g.h.pair_count, sizeof(int64_t) * 3, cudaMemcpyDeviceToHost));
    return 0;
}
int mgpu_get_screened_pairs(int64_t *n)
{
    NEED_READY();
    CK(cudaMemcpy(n, g.h.pair_count + 3, sizeof(int64_t), cudaMemcpyDeviceToHost));
    return 0;
}
int mgpu_reset_pair_counts(void)
{
    NEED_READY();
    CK(cudaMemset(g.h.pair_count, 0, sizeof(unsigned long long) * 4));
    return 0;
}
int mgpu_timing_reset(void) { g.timing.clear(); return 0; }
int mgpu_timing_get(const char *kernel, double *total_ms, int64_t *launches)
{
    auto it = g.timing.find(kernel ? kernel : "");
    if (it == g.timing.end()) { if (total_ms) *total_ms = 0.0; if (launches) *launches = 0; return 0; }
    if (total_ms) *total_ms = it->second.ms;
    if (launches) *launches = it->second.launches;
    return 0;
}
int mgpu_selftest_math(double *max_rel_rcp, double *max_err_table)
{
    NEED_READY();
    double *d_out;
    CK(cudaMalloc(&d_out, sizeof(double) * 2));
    CK(cudaMemsetAsync(d_out, 0, sizeof(double) * 2, g.stream));
    const double s_lo = std::ldexp(1.0, g.tab_emin) * 1.0000001, s_hi = std::ldexp(1.0, g.tab_emin + g.tab_noct) * 0.9999999;
    k_selftest<<<g.sm_count * 4, 256, 0, g.stream>>>(s_lo, s_hi, 1 << 22, d_out);
    CK(cudaGetLastError());
    double out[2] = { 0.0, 0.0 };
    CK(cudaMemcpyAsync(out, d_out, sizeof out, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    cudaFree(d_out);
    if (max_rel_rcp) *max_rel_rcp = out[0];
    if (max_err_table) *max_err_table = out[1];
    return 0;
}
// L2 read bandwidth (roofline denominator of K2): every SM streams a 48 MB buffer that stays resident in
// the 126 MB L2, 16-byte loads, several passes; the first pass (DRAM) is not timed.
int mgpu_measure_l2_peak(double *gbytes_per_s)
{
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t st = g.stream;
    bool own = false;
    if (!st) { CK(cudaStreamCreate(&st)); own = true; }
    const size_t n16 = (size_t)48 * 1024 * 1024 / 16;
    double2 *buf; double *out;
    CK(cudaMalloc(&buf, n16 * 16));
    CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
    CK(cudaMemsetAsync(buf, 0, n16 * 16, st));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int passes = 20;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(a, st));
        k_l2_read<<<sms * 8, 256, 0, st>>>(buf, n16, rep == 0 ? 1 : passes, out);
        CK(cudaEventRecord(b, st));
        CK(cudaEventSynchronize(b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0) best = std::fmax(best, (double)n16 * 16.0 * passes / (ms * 1e-3) / 1e9);
    }
    CK(cudaGetLastError());
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(buf); cudaFree(out);
    if (own) cudaStreamDestroy(st);
    if (gbytes_per_s) *gbytes_per_s = best;
    return 0;
}
int mgpu_measure_fp64_peak(double *tflops, double *seconds)
{
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t st = g.stream;
    bool own = false;
    if (!st) { CK(cudaStreamCreate(&st)); own = true; }
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double *d_out;
    CK(cudaMalloc(&d_out, sizeof(double) * blocks * threads));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    double best = 0.0, best_s = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(a, st));
        k_dfma_peak<<<blocks, threads, 0, st>>>(d_out, iters, 0.999999, 1e-9);
        CK(cudaEventRecord(b, st));
        CK(cudaEventSynchronize(b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) { best = tf; best_s = ms * 1e-3; }
    }
    CK(cudaGetLastError());
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d_out);
    if (own) cudaStreamDestroy(st);
    if (tflops) *tflops = best;
    if (seconds) *seconds = best_s;
    return 0;
}

} // extern "C"
