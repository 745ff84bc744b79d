// maniac_host.cpp -- host-side MC drivers over the C ABI (see include/maniac_host.h).
// The role of src/monte_carlo.f90:50-99 + translation/rotation/creation/deletion/widom.f90:
// propose on the host, ask the GPU for old/new energies, apply the acceptance rule of
// src/monte_carlo_utils.f90:204-255, commit or roll back.  Only mgpu_* exports are used.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/maniac_host.h"

namespace {
std::string h_err;
const double ERR = 1.0e-10;
const double PI_ = 3.1415926536, TWOPI_ = 2.0 * PI_;

inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
struct Rng {
    uint64_t s[4];
    void seed(uint64_t z) { for (int i = 0; i < 4; ++i) { z += 0x9E3779B97F4A7C15ULL; s[i] = mix64(z); } }
    double uniform()
    {
        const uint64_t r = rotl(s[1] * 5u, 7) * 9u, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return (double)(r >> 11) * 0x1.0p-53;
    }
};
inline double f_modulo(double a, double p)
{
    double r = std::fmod(a, p);
    if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p;
    return r;
}

struct ResInfo { int natom = 0, active = 0, cap = 0; double lambda = 0.0; };
struct Walker {
    std::vector<int> count;                       // primary%num%residues
    std::vector<std::vector<double>> com, off;    // guest%com(:,res,mol), guest%offset(:,res,mol,atom)
    std::vector<double> mu;
    double energy[6];
    Rng rng;
    long long counter[6][2];
    std::vector<double> widom_w; std::vector<long long> widom_n;
    double tstep = 0.0, rstep = 0.0;              // mc_input%translation_step / rotation_step_angle (adjust_move_step_sizes)
};
struct Pending { int valid, move, kind, res, mol, res2; double com[3]; double off[MGPU_MAX_SITES][3]; };
} // namespace

struct mhost_sim {
    int nres = 0, nw = 0;
    std::vector<ResInfo> res;
    std::vector<int> active_list;
    double H[9], Hinv[9], lo[3], volume = 0.0, beta = 0.0;
    int triclinic = 0;
    double p_trans, p_rot, p_swap, p_insdel, p_widom, tstep, rstep;
    std::vector<Walker> w;
    std::vector<Pending> pend;
    // batch buffers
    std::vector<int32_t> b_walker, b_res, b_mol, b_kind, b_accept;
    std::vector<double> b_com, b_off, b_eold, b_enew;
    long long h2d = 0, d2h = 0, trials = 0;
};

namespace {
int hfail(const std::string &m) { h_err = m; return 1; }

// The host loops of a step (propose, fill the batch, decide) touch one walker each and walkers share nothing, so they
// are split over a small team of threads.  The team only lives inside mhost_run and spins between the three loops of a
// step (a condition-variable wake-up costs as much as the loop it would start); results do not depend on the split.
class Team {
public:
    explicit Team(int n) : nt_(n < 1 ? 1 : n)
    {
        for (int t = 1; t < nt_; ++t) th_.emplace_back([this, t] { work(t); });
    }
    ~Team()
    {
        stop_.store(true, std::memory_order_release);
        gen_.fetch_add(1, std::memory_order_acq_rel);
        for (auto &t : th_) t.join();
    }
    void run(int n, const std::function<void(int)> &fn)
    {
        if (nt_ == 1 || n < 64) { for (int i = 0; i < n; ++i) fn(i); return; }
        fn_ = &fn; n_ = n;
        done_.store(0, std::memory_order_relaxed);
        gen_.fetch_add(1, std::memory_order_acq_rel);
        chunk(0);
        while (done_.load(std::memory_order_acquire) != nt_ - 1) { }
    }
private:
    void chunk(int t) const { const int lo = (int)((long long)n_ * t / nt_), hi = (int)((long long)n_ * (t + 1) / nt_); for (int i = lo; i < hi; ++i) (*fn_)(i); }
    void work(int t)
    {
        unsigned seen = 0;
        for (;;) {
            unsigned g;
            while ((g = gen_.load(std::memory_order_acquire)) == seen) { }
            seen = g;
            if (stop_.load(std::memory_order_acquire)) return;
            chunk(t);
            done_.fetch_add(1, std::memory_order_acq_rel);
        }
    }
    int nt_, n_ = 0;
    const std::function<void(int)> *fn_ = nullptr;
    std::vector<std::thread> th_;
    std::atomic<unsigned> gen_{0};
    std::atomic<int> done_{0};
    std::atomic<bool> stop_{false};
};
int team_size()
{
    if (const char *e = std::getenv("MHOST_THREADS")) return std::atoi(e);
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)(hw >= 16 ? 8 : hw >= 4 ? hw / 2 : 1);
}

void apply_PBC(const mhost_sim *S, double pos[3])      // geometry_utils.f90:45-97
{
    if (!S->triclinic) {
        for (int d = 0; d < 3; ++d) pos[d] = S->lo[d] + f_modulo(pos[d] - S->lo[d], S->H[d * 3 + d]);
    } else {
        double rel[3], f[3];
        for (int d = 0; d < 3; ++d) rel[d] = pos[d] - S->lo[d];
        for (int i = 0; i < 3; ++i) {
            f[i] = S->Hinv[i * 3 + 0] * rel[0] + S->Hinv[i * 3 + 1] * rel[1] + S->Hinv[i * 3 + 2] * rel[2];
            f[i] = f_modulo(f[i], 1.0);
        }
        for (int i = 0; i < 3; ++i) pos[i] = S->lo[i] + (S->H[i * 3 + 0] * f[0] + S->H[i * 3 + 1] * f[1] + S->H[i * 3 + 2] * f[2]);
    }
}
void rotate_offsets(int axis, double theta, double (*off)[3], int na)   // helper_utils.f90:30-75 + matmul
{
    const double c = std::cos(theta), sn = std::sin(theta);
    double M[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
    if (axis == 1) { M[1][1] = c; M[1][2] = -sn; M[2][1] = sn; M[2][2] = c; }
    else if (axis == 2) { M[0][0] = c; M[0][2] = sn; M[2][0] = -sn; M[2][2] = c; }
    else { M[0][0] = c; M[0][1] = -sn; M[1][0] = sn; M[1][1] = c; }
    for (int a = 0; a < na; ++a) {
        const double v[3] = { off[a][0], off[a][1], off[a][2] };
        for (int i = 0; i < 3; ++i) off[a][i] = M[i][0] * v[0] + M[i][1] * v[1] + M[i][2] * v[2];
    }
}
} // namespace

extern "C" {

const char *mhost_last_error(void) { return h_err.c_str(); }

mhost_sim *mhost_create(const mgpu_system *sys, uint64_t seed)
{
    mhost_sim *S = new mhost_sim();
    S->nres = sys->nres; S->nw = sys->n_walkers;
    int32_t tri = 0;
    if (mgpu_get_box(S->H, S->Hinv, &S->volume, &tri)) { h_err = mgpu_last_error(); delete S; return nullptr; }
    S->triclinic = tri;
    for (int d = 0; d < 3; ++d) S->lo[d] = sys->lo[d];
    S->p_trans = sys->p_translation; S->p_rot = sys->p_rotation; S->p_swap = sys->p_swap;
    S->p_insdel = sys->p_insertion_deletion; S->p_widom = sys->p_widom;
    S->tstep = sys->translation_step; S->rstep = sys->rotation_step_angle;
    S->res.resize(S->nres);
    for (int r = 0; r < S->nres; ++r) {
        const mgpu_residue &R = sys->residues[r];
        S->res[r].natom = R.natom; S->res[r].active = R.is_active; S->res[r].cap = R.capacity;
        if (R.is_active) {
            S->active_list.push_back(r);
            double mu0;
            if (mgpu_get_thermo(r, &S->beta, &S->res[r].lambda, &mu0)) { h_err = mgpu_last_error(); delete S; return nullptr; }
        }
    }
    S->w.resize(S->nw);
    S->pend.resize(S->nw);
    for (int w = 0; w < S->nw; ++w) {
        Walker &W = S->w[w];
        W.count.assign(S->nres, 0); W.com.resize(S->nres); W.off.resize(S->nres); W.mu.assign(S->nres, 0.0);
        W.widom_w.assign(S->nres, 0.0); W.widom_n.assign(S->nres, 0);
        std::memset(W.counter, 0, sizeof W.counter);
        for (int r = 0; r < S->nres; ++r) {
            const mgpu_residue &R = sys->residues[r];
            if (!R.is_active) continue;
            W.count[r] = R.nmol;
            W.com[r].assign((size_t)3 * R.capacity, 0.0);
            W.off[r].assign((size_t)3 * R.natom * R.capacity, 0.0);
            std::memcpy(W.com[r].data(), R.com, sizeof(double) * 3 * R.nmol);
            std::memcpy(W.off[r].data(), R.offset, sizeof(double) * 3 * R.natom * R.nmol);
            double b, l;
            mgpu_get_thermo(r, &b, &l, nullptr);
            // per-walker chemical potential as the engine holds it
            W.mu[r] = (R.fugacity >= 0.0) ? std::log(R.fugacity) / S->beta : R.chemical_potential;
        }
        if (mgpu_get_energy(w, W.energy)) { h_err = mgpu_last_error(); delete S; return nullptr; }
        W.rng.seed(seed + 104729ull * (uint64_t)w);
        W.tstep = S->tstep; W.rstep = S->rstep;
    }
    return S;
}

void mhost_destroy(mhost_sim *S) { delete S; }

int mhost_run(mhost_sim *S, int64_t n_steps, int32_t trace_walker, mgpu_step_trace *trace)
{
    const int nw = S->nw;
    Team team(nw >= 64 ? team_size() : 1);
    std::atomic<int> overflow{0};
    std::vector<int> slot(nw, 0);
    S->b_walker.resize(nw); S->b_res.resize(nw); S->b_mol.resize(nw); S->b_kind.resize(nw); S->b_accept.resize(nw);
    S->b_com.resize((size_t)3 * nw); S->b_off.resize((size_t)3 * MGPU_MAX_SITES * nw);
    S->b_eold.resize((size_t)6 * nw); S->b_enew.resize((size_t)6 * nw);
    const double cumul_translation = S->p_trans, cumul_rotation = cumul_translation + S->p_rot, cumul_swap = cumul_rotation + S->p_swap;
    for (int64_t step = 0; step < n_steps; ++step) {
        int nb = 0;
        // ---- propose (host), monte_carlo.f90:53-99 ----
        const std::function<void(int)> propose_one = [&](int w) {
            Walker &W = S->w[w];
            Pending &P = S->pend[w];
            P.valid = 0; P.move = MGPU_MV_NONE;
            const int res = S->active_list[(int)(W.rng.uniform() * (double)S->active_list.size())];
            const int n = W.count[res], na = S->res[res].natom, cap = S->res[res].cap;
            int mol = -1;
            if (n > 0) { mol = (int)(W.rng.uniform() * n) + 1; if (mol > n) mol = n; mol -= 1; }
            const double draw = W.rng.uniform();
            P.res = res; P.mol = mol;
            double *com = W.com[res].data(), *off = W.off[res].data();
            if (draw <= cumul_translation) {
                if (mol >= 0) {
                    P.valid = 1; P.move = MGPU_MV_TRANSLATE; P.kind = MGPU_KIND_MOVE;
                    double tp[3];
                    for (int d = 0; d < 3; ++d) tp[d] = W.rng.uniform();
                    for (int d = 0; d < 3; ++d) P.com[d] = com[3 * mol + d] + (tp[d] - 0.5) * W.tstep;
                    apply_PBC(S, P.com);
                    std::memcpy(P.off, off + (size_t)3 * na * mol, sizeof(double) * 3 * na);
                }
            } else if (draw <= cumul_rotation) {
                if (na != 1 && mol >= 0) {
                    P.valid = 1; P.move = MGPU_MV_ROTATE; P.kind = MGPU_KIND_MOVE;
                    std::memcpy(P.com, com + 3 * mol, sizeof(double) * 3);
                    std::memcpy(P.off, off + (size_t)3 * na * mol, sizeof(double) * 3 * na);
                    const double theta = (W.rng.uniform() - 0.5) * W.rstep;
                    const int axis = (int)(W.rng.uniform() * 3.0) + 1;
                    rotate_offsets(axis, theta, P.off, na);
                }
            } else if (draw <= cumul_swap) {                              // swapping.f90:34-105
                int res_bis = -1;
                for (int attempt = 0; attempt < 10; ++attempt) {          // pick_different_residue_type :165-197
                    const int nt = S->active_list[(int)(W.rng.uniform() * (double)S->active_list.size())];
                    if (nt != res) { res_bis = nt; break; }
                }
                if (res_bis >= 0 && W.count[res_bis] > 0 && mol >= 0) {
                    const int nb2 = W.count[res_bis], nab = S->res[res_bis].natom;
                    if (nb2 >= S->res[res_bis].cap) { overflow.store(1); return; }
                    P.valid = 1; P.move = MGPU_MV_SWAP; P.kind = MGPU_KIND_SWAP; P.res2 = res_bis;
                    std::memcpy(P.com, com + 3 * mol, sizeof(double) * 3);                        // same CoM (:80)
                    std::memcpy(P.off, W.off[res_bis].data(), sizeof(double) * 3 * nab);          // geometry of molecule 1 of the new type
                    if (nab != 1) {
                        const double theta = W.rng.uniform() * TWOPI_;
                        const int axis = (int)(W.rng.uniform() * 3.0) + 1;
                        rotate_offsets(axis, theta, P.off, nab);
                    }
                }
            } else {
                bool create = false, widom = false, del = false;
                if (S->p_insdel > 0) { if (W.rng.uniform() <= 0.5) create = true; else del = true; }
                else if (S->p_widom > 0) widom = true;
                if (create || widom) {
                    if (n >= cap) { overflow.store(1); return; }
                    P.valid = 1; P.move = widom ? MGPU_MV_WIDOM : MGPU_MV_CREATE; P.kind = MGPU_KIND_CREATE; P.mol = n;
                    double t3[3];
                    for (int d = 0; d < 3; ++d) t3[d] = W.rng.uniform();
                    for (int i = 0; i < 3; ++i) P.com[i] = S->lo[i] + (S->H[i * 3 + 0] * t3[0] + S->H[i * 3 + 1] * t3[1] + S->H[i * 3 + 2] * t3[2]);
                    std::memcpy(P.off, off, sizeof(double) * 3 * na);           // geometry of molecule 1
                    if (na != 1) {
                        const double theta = W.rng.uniform() * TWOPI_;
                        const int axis = (int)(W.rng.uniform() * 3.0) + 1;
                        rotate_offsets(axis, theta, P.off, na);
                    }
                    // the reference writes the trial molecule into slot N+1 before the accept test
                    std::memcpy(com + 3 * (size_t)n, P.com, sizeof(double) * 3);
                    std::memcpy(off + (size_t)3 * na * n, P.off, sizeof(double) * 3 * na);
                } else if (del) {
                    if (n > 0) { P.valid = 1; P.move = MGPU_MV_DELETE; P.kind = MGPU_KIND_DELETE; }
                }
            }
        };
        team.run(nw, propose_one);
        if (overflow.load()) return hfail("Trying to insert a molecule beyond the walker's capacity");
        // batch slots in walker order, then the copies in parallel
        for (int w = 0; w < nw; ++w) if (S->pend[w].valid) slot[w] = nb++;
        const std::function<void(int)> fill_one = [&](int w) {
            const Pending &P = S->pend[w];
            if (!P.valid) return;
            const int b = slot[w], na = S->res[P.res].natom;
            S->b_walker[b] = w; S->b_res[b] = P.res; S->b_mol[b] = P.mol;
            S->b_kind[b] = (P.kind == MGPU_KIND_SWAP) ? (MGPU_KIND_SWAP | (P.res2 << 8)) : P.kind;
            std::memcpy(&S->b_com[(size_t)3 * b], P.com, sizeof(double) * 3);
            std::memset(&S->b_off[(size_t)3 * MGPU_MAX_SITES * b], 0, sizeof(double) * 3 * MGPU_MAX_SITES);
            const int na_new = (P.kind == MGPU_KIND_SWAP) ? S->res[P.res2].natom : na;
            if (P.kind != MGPU_KIND_DELETE) std::memcpy(&S->b_off[(size_t)3 * MGPU_MAX_SITES * b], P.off, sizeof(double) * 3 * na_new);
        };
        team.run(nw, fill_one);
        // ---- energies (GPU): compute_old_energy + compute_new_energy for every walker ----
        if (nb) {
            if (mgpu_trial_batch(nb, S->b_walker.data(), S->b_res.data(), S->b_mol.data(), S->b_kind.data(), S->b_com.data(),
                                 S->b_off.data(), S->b_eold.data(), S->b_enew.data())) return hfail(mgpu_last_error());
            S->h2d += (long long)nb * (16 + 24 + 8 * 3 * MGPU_MAX_SITES);
            S->d2h += (long long)nb * 96;
            S->trials += nb;
        }
        // ---- Metropolis (host), monte_carlo_utils.f90:204-255 ----
        const std::function<void(int)> decide_one = [&](int b) {
            const int w = S->b_walker[b];
            Walker &W = S->w[w];
            Pending &P = S->pend[w];
            const double *eo = &S->b_eold[(size_t)6 * b], *en = &S->b_enew[(size_t)6 * b];
            const int res = P.res, na = S->res[res].natom;
            const double dU = en[5] - eo[5], lam = S->res[res].lambda, mu = W.mu[res];
            double p; int acc = 0;
            if (P.move == MGPU_MV_WIDOM) {
                p = std::exp(-dU * S->beta);
                W.counter[5][0] += 1;
                if (p > ERR) { W.counter[5][1] += 1; W.widom_w[res] += p; }
                W.widom_n[res] += 1;
            } else {
                if (P.move == MGPU_MV_CREATE) {
                    const double N = (double)(W.count[res] + 1);
                    p = std::fmin(1.0, S->volume / N / (lam * lam * lam) * std::exp(-S->beta * (dU - mu)));
                } else if (P.move == MGPU_MV_DELETE) {
                    const double Np1 = (double)(W.count[res] - 1) + 1.0;
                    p = std::fmin(1.0, Np1 * (lam * lam * lam) / S->volume * std::exp(-S->beta * (dU + mu)));
                } else if (P.move == MGPU_MV_SWAP) {                       // swap_acceptance_probability :264-294, counts as the reference has them there
                    const double N_old = (double)(W.count[res] - 1), Nplus1 = (double)(W.count[P.res2] + 1) + 1.0;
                    p = std::fmin(1.0, (N_old / Nplus1) * std::exp(-S->beta * (dU + W.mu[P.res2] - mu)));
                } else p = std::fmin(1.0, std::exp(-S->beta * dU));
                acc = W.rng.uniform() <= p;
                const int ci = (P.move == MGPU_MV_TRANSLATE) ? 0 : (P.move == MGPU_MV_ROTATE) ? 1 : (P.move == MGPU_MV_CREATE) ? 2
                             : (P.move == MGPU_MV_DELETE) ? 3 : 4;
                W.counter[ci][0] += 1;
                if (acc) { W.counter[ci][1] += 1; if (ci >= 2) W.counter[ci][0] += 1; }
            }
            S->b_accept[b] = acc;
            if (acc) {
                double *com = W.com[res].data(), *off = W.off[res].data();
                if (P.kind == MGPU_KIND_DELETE || P.kind == MGPU_KIND_SWAP) { // remove_molecule + update_counts
                    const int last = W.count[res] - 1;
                    if (P.mol != last) {
                        std::memcpy(com + 3 * (size_t)P.mol, com + 3 * (size_t)last, sizeof(double) * 3);
                        std::memcpy(off + (size_t)3 * na * P.mol, off + (size_t)3 * na * last, sizeof(double) * 3 * na);
                    }
                    W.count[res] -= 1;
                    if (P.kind == MGPU_KIND_SWAP) {                       // the new molecule goes to slot N+1 of its type
                        const int r2 = P.res2, n2 = W.count[r2], na2 = S->res[r2].natom;
                        std::memcpy(W.com[r2].data() + 3 * (size_t)n2, P.com, sizeof(double) * 3);
                        std::memcpy(W.off[r2].data() + (size_t)3 * na2 * n2, P.off, sizeof(double) * 3 * na2);
                        W.count[r2] += 1;
                    }
                } else {
                    std::memcpy(com + 3 * (size_t)P.mol, P.com, sizeof(double) * 3);
                    std::memcpy(off + (size_t)3 * na * P.mol, P.off, sizeof(double) * 3 * na);
                    if (P.kind == MGPU_KIND_CREATE) W.count[res] += 1;
                }
                double *E = W.energy;                                     // accept_* bookkeeping
                E[MGPU_E_RECIP] = en[MGPU_E_RECIP];
                E[MGPU_E_NON_COULOMB] = E[MGPU_E_NON_COULOMB] + en[MGPU_E_NON_COULOMB] - eo[MGPU_E_NON_COULOMB];
                E[MGPU_E_COULOMB] = E[MGPU_E_COULOMB] + en[MGPU_E_COULOMB] - eo[MGPU_E_COULOMB];
                if (P.kind != MGPU_KIND_MOVE) {
                    E[MGPU_E_SELF] = E[MGPU_E_SELF] + en[MGPU_E_SELF] - eo[MGPU_E_SELF];
                    E[MGPU_E_INTRA] = E[MGPU_E_INTRA] + en[MGPU_E_INTRA] - eo[MGPU_E_INTRA];
                }
                E[MGPU_E_TOTAL] = E[MGPU_E_TOTAL] + en[MGPU_E_TOTAL] - eo[MGPU_E_TOTAL];
            }
            if (trace && w == trace_walker) {
                mgpu_step_trace &t = trace[step];
                t.move = P.move; t.res = P.res; t.mol = P.mol; t.accepted = acc; t.dE = dU; t.prob = p;
                std::memcpy(t.e_old, eo, sizeof t.e_old); std::memcpy(t.e_new, en, sizeof t.e_new);
            }
        };
        team.run(nb, decide_one);
        if (trace && trace_walker >= 0 && trace_walker < nw && !S->pend[trace_walker].valid) {
            mgpu_step_trace &t = trace[step];
            std::memset(&t, 0, sizeof t);
            t.move = MGPU_MV_NONE; t.res = S->pend[trace_walker].res; t.mol = S->pend[trace_walker].mol;
        }
        if (nb) {
            if (mgpu_commit_batch(nb, S->b_walker.data(), S->b_accept.data())) return hfail(mgpu_last_error());
            S->h2d += (long long)nb * 8;
        }
    }
    return 0;
}

// adjust_move_step_sizes, monte_carlo_utils.f90:98-134 (host side, the walker's own counters)
int mhost_adjust_move_step_sizes(mhost_sim *S)
{
    const double gamma = 0.10, TARGET_ACCEPTANCE = 0.40;
    for (Walker &W : S->w) {
        if (W.counter[0][0] > 500) {
            const double acc = (double)W.counter[0][1] / (double)W.counter[0][0];
            W.tstep = std::fmax(1.0e-3, std::fmin(W.tstep * std::exp(gamma * (acc - TARGET_ACCEPTANCE)), 3.0));
        }
        if (W.counter[1][0] > 500) {
            const double acc = (double)W.counter[1][1] / (double)W.counter[1][0];
            W.rstep = std::fmax(1.0e-3, std::fmin(W.rstep * std::exp(gamma * (acc - TARGET_ACCEPTANCE)), 0.78));
        }
    }
    return 0;
}
int mhost_get_step_sizes(const mhost_sim *S, int32_t w, double out[2]) { out[0] = S->w[w].tstep; out[1] = S->w[w].rstep; return 0; }
int mhost_set_chemical_potential(mhost_sim *S, int32_t w, int32_t res, double mu)
{
    if (w < 0 || w >= S->nw || res < 0 || res >= S->nres) return hfail("mhost_set_chemical_potential: index out of range");
    S->w[w].mu[res] = mu;
    return 0;
}
int mhost_get_count(const mhost_sim *S, int32_t w, int32_t res) { return S->w[w].count[res]; }
int mhost_get_energy(const mhost_sim *S, int32_t w, double out[6]) { std::memcpy(out, S->w[w].energy, sizeof(double) * 6); return 0; }
int mhost_get_counters(const mhost_sim *S, int32_t w, int64_t out[12])
{
    for (int i = 0; i < 6; ++i) { out[2 * i] = S->w[w].counter[i][0]; out[2 * i + 1] = S->w[w].counter[i][1]; }
    return 0;
}
int mhost_get_molecule(const mhost_sim *S, int32_t w, int32_t res, int32_t mol, double com[3], double *offset)
{
    const int na = S->res[res].natom;
    std::memcpy(com, S->w[w].com[res].data() + 3 * (size_t)mol, sizeof(double) * 3);
    std::memcpy(offset, S->w[w].off[res].data() + (size_t)3 * na * mol, sizeof(double) * 3 * na);
    return 0;
}
int mhost_get_traffic(const mhost_sim *S, int64_t *h2d, int64_t *d2h, int64_t *trials)
{
    if (h2d) *h2d = S->h2d;
    if (d2h) *d2h = S->d2h;
    if (trials) *trials = S->trials;
    return 0;
}

} // extern "C"
