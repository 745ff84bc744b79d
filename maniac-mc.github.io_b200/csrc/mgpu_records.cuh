// mgpu_records.cuh -- walker records: the whole state of a walker as one flat run of doubles in
// HOST memory, so that a block of the MC loop (monte_carlo.f90:40-118: nb_step steps, then the
// per-block outputs) can go host state -> device -> host state in one call (mgpu_block).
//
// Record of one walker (doubles; integers are stored as exactly representable doubles, the RNG
// state as raw bit patterns):
//   [0]            record length
//   [1 .. 8]       primary%num%residues(res), res = 0..MGPU_MAX_RES-1
//   [9 .. 14]      energy (energy_type order)
//   [15 .. 18]     xoshiro256** state
//   [19 .. 30]     counters (translations, rotations, creations, deletions, swaps, widom) x (trials, successes)
//   [32 .. 63]     block averages [res][sum N, sum N^2, sum E, samples]
//   [64], [65]     mc_input%translation_step, mc_input%rotation_step_angle of the walker;  [66 .. 71] reserved (0)
//   [72 .. 72+2nk) S(k) = ewald%Ak: re[nk], im[nk]
//   then, for every ACTIVE residue type in index order, max(count(res), 1) molecules of
//   { com[3], offset[natom][3], framework-energy cache {lj, coulomb} }   (guest%com(:,res,mol), guest%offset(:,res,mol,1:natom))
//   -- slot 1 travels even when the walker holds no molecule of the type: it is the geometry template of the next
//   insertion (insert_and_orient_molecule copies offset(:,res,1,:), monte_carlo_utils.f90:563-564) and a rejected
//   creation leaves its trial geometry there (:579-591), so an empty walker's future depends on it.
#pragma once
#include "mgpu_kernels.cuh"

#define MGPU_REC_HDR 72

__device__ __forceinline__ int rec_molsize(int res) { return 3 + 3 * c_sys.natom[res] + 2; }
__device__ __forceinline__ int rec_stored(int count) { return count > 0 ? count : 1; }      // molecules of a type stored in a record

__global__ void k_record_len(int first, int n, long long *len)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int w = first + i;
    long long L = MGPU_REC_HDR + 2 * (long long)c_sys.nk;
    for (int r = 0; r < c_sys.nres; ++r)
        if (c_sys.active[r]) L += (long long)rec_stored(c_sys.count[(int64_t)w * MGPU_MAX_RES + r]) * rec_molsize(r);
    len[i] = L;
}

// record lengths of walkers [first, first + n) and their exclusive prefix sum, in one launch of ONE CTA (pipelined
// mgpu_block: no host round trip between the sweep and the pack).  off[0] = 0, off[n] = total.
__global__ void __launch_bounds__(256) k_record_offsets(int first, int n, long long *off)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int w = first + i;
        long long L = MGPU_REC_HDR + 2 * (long long)c_sys.nk;
        for (int r = 0; r < c_sys.nres; ++r)
            if (c_sys.active[r]) L += (long long)rec_stored(c_sys.count[(int64_t)w * MGPU_MAX_RES + r]) * rec_molsize(r);
        off[i + 1] = L;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        off[0] = 0;
        for (int i = 0; i < n; ++i) { run += off[i + 1]; off[i + 1] = run; }
    }
}

// one warp per walker
__global__ void __launch_bounds__(256) k_pack(int first, int n, const long long *off, double *blob)
{
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const int w = first + i;
    double *rec = blob + off[i];
    const int nk = c_sys.nk;
    if (lane == 0) rec[0] = (double)(off[i + 1] - off[i]);
    if (lane < MGPU_MAX_RES) rec[1 + lane] = (double)c_sys.count[(int64_t)w * MGPU_MAX_RES + lane];
    if (lane < 6) rec[9 + lane] = c_sys.energy[(int64_t)w * 6 + lane];
    if (lane < 4) rec[15 + lane] = __longlong_as_double((long long)c_sys.rng[(int64_t)w * 4 + lane]);
    if (lane < 12) rec[19 + lane] = (double)c_sys.counters[(int64_t)w * 12 + lane];
    if (lane == 0) rec[31] = 0.0;
    rec[32 + lane] = c_sys.avg[(int64_t)w * MGPU_MAX_RES * 4 + lane];
    if (lane < 8) rec[64 + lane] = lane < 2 ? c_sys.step[(int64_t)w * 2 + lane] : 0.0;
    const double *Sk = c_sys.S + ((int64_t)w * 2 + c_sys.cur[w]) * 2 * nk;
    for (int k = lane; k < 2 * nk; k += 32) rec[MGPU_REC_HDR + k] = Sk[k];
    double *p = rec + MGPU_REC_HDR + 2 * nk;
    const double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride;
    for (int r = 0; r < c_sys.nres; ++r) {
        if (!c_sys.active[r]) continue;
        const int cnt = rec_stored(c_sys.count[(int64_t)w * MGPU_MAX_RES + r]), ms = rec_molsize(r), cap = c_sys.cap[r];
        const double *src = wc + c_sys.goff[r];         // rows: com x,y,z | offset rows (atom, dim) | cache lj, coulomb; each cap long
        for (int t = lane; t < cnt * ms; t += 32) {
            const int m = t / ms, e = t - m * ms;
            p[t] = src[(int64_t)e * cap + m];
        }
        p += (long long)cnt * ms;
    }
}

// (Records are validated on the host before anything is uploaded: validate_records in maniac_gpu.cu.)
__global__ void __launch_bounds__(256) k_unpack(int first, int n, const long long *off, const double *blob)
{
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const int w = first + i;
    const double *rec = blob + off[i];
    const int nk = c_sys.nk;
    if (lane < MGPU_MAX_RES && c_sys.active[lane]) c_sys.count[(int64_t)w * MGPU_MAX_RES + lane] = (int)rec[1 + lane];
    if (lane < 6) c_sys.energy[(int64_t)w * 6 + lane] = rec[9 + lane];
    if (lane < 4) c_sys.rng[(int64_t)w * 4 + lane] = (uint64_t)__double_as_longlong(rec[15 + lane]);
    if (lane < 12) c_sys.counters[(int64_t)w * 12 + lane] = (long long)rec[19 + lane];
    c_sys.avg[(int64_t)w * MGPU_MAX_RES * 4 + lane] = rec[32 + lane];
    if (lane < 2) c_sys.step[(int64_t)w * 2 + lane] = rec[64 + lane];
    if (lane == 0) { c_sys.cur[w] = 0; c_sys.trial[w].active = 0; }
    double *Sk = c_sys.S + ((int64_t)w * 2 + 0) * 2 * nk;
    for (int k = lane; k < 2 * nk; k += 32) Sk[k] = rec[MGPU_REC_HDR + k];
    const double *p = rec + MGPU_REC_HDR + 2 * nk;
    double *wc = c_sys.coords + (int64_t)w * c_sys.coord_stride;
    for (int r = 0; r < c_sys.nres; ++r) {
        if (!c_sys.active[r]) continue;
        const int cnt = rec_stored((int)rec[1 + r]), ms = rec_molsize(r), cap = c_sys.cap[r];
        double *dst = wc + c_sys.goff[r];
        for (int t = lane; t < cnt * ms; t += 32) {
            const int m = t / ms, e = t - m * ms;
            dst[(int64_t)e * cap + m] = p[t];
        }
        // keep the bound on |offset|^2 of this residue type valid for whatever the record brings in
        double r2 = 0.0;
        const int na = c_sys.natom[r];
        for (int t = lane; t < cnt * na; t += 32) {
            const int m = t / na, a = t - m * na;
            const double *o = p + (long long)m * ms + 3 + 3 * a;
            r2 = fmax(r2, o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, o));
        if (lane == 0 && r2 > c_sys.rmax2[r])
            atomicMax(reinterpret_cast<unsigned long long *>(c_sys.rmax2 + r), (unsigned long long)__double_as_longlong(r2));
        p += (long long)cnt * ms;
    }
}
