// hb_pt.cu -- kernels of the device-resident parallel-tempering sampler (see hb_pt.cuh for the
// algorithm, its reference citations and the positions on the reference's sampler bugs).
//
// Layout: E ensembles (independent PT ladders) x T rungs.  Chain state x[E*T][21], logL[E*T] is
// indexed by chain SLOT; index[E][T] maps rung -> slot inside the ensemble (the reference swaps
// this permutation, not the states, mcmc_wrapper2.c:812-816); the DE history ring is per RUNG
// (history[j], mcmc_wrapper2.c:424,543-546).  One iteration = k_pt_propose -> k_prologue ->
// k_chain_eval (the likelihood of all E*T proposals, one batch) -> k_pt_accept -> k_pt_swap.
#include <cstdio>

#include "hb_kernels.h"
#include "hb_select.cuh"
#include "hb_device.cuh"
#include "hb_pt.cuh"
#include "hb_pt_run.h"

namespace hb {

__global__ void k_pt_init_random(const PtConfig* __restrict__ cfgp, double* __restrict__ x, int W)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const PtConfig& cfg = *cfgp;
    PtRng g;
    g.init(cfg.seed, (uint32_t)(w + cfg.ens_offset * cfg.n_temps), 0xFFFFFFFFu, 3u);
    double* xw = x + (size_t)w * kPtNpars;
    // uniform in the prior box, period pinned, T0 folded (mcmc_wrapper2.c:236-251)
    for (int i = 0; i < kPtNpars; i++) {
        const double u = g.next();
        xw[i] = cfg.lo[i] + u * (cfg.hi[i] - cfg.lo[i]);
    }
    xw[2] = cfg.log_lc_period;
    xw[6] = fmod(xw[6], cfg.lc_period);
}

// ---- warp-per-walker step kernels --------------------------------------------------------
// At the reference's own size (50 rungs, a few hundred samples) the step is latency-bound, so the
// per-walker work is spread over a warp: lane n owns parameter n (21 of 32 lanes), every lane
// evaluates its own Philox draws by index (pt_draw) and its own bound / prior term; sums that the
// sequential formulation takes in parameter order are taken in that order by lane 0, so the values
// are those of pt_propose / pt_accept in hb_pt.cuh bit for bit.
__device__ __forceinline__ double warp_ordered_sum(double term, int count)
{
    double acc = 0.;
    for (int n = 0; n < count; n++) acc += __shfl_sync(0xffffffffu, term, n);
    return acc;  // identical on every lane
}

// The random part of one walker's proposal at iteration `iter`: everything that depends on the rung's Philox stream
// alone, not on the walker's state -- so a sampler that keeps its walkers resident (k_pt_run) draws iteration i + 1
// on an idle warp while iteration i is being evaluated.  Lane n holds the normals of parameter n; the scalars are
// the same on every lane.  mcmc_wrapper2.c:390-436, :1062-1140.
struct PtDraws {
    double jscale;   // 10^(-6 + 6 alpha)
    double z_gauss;  // this lane's normal of the Gaussian jump (meaningful when !de)
    double z_de;     // this lane's normal of the scaled DE jump (meaningful when de && scaled)
    int de, a, b, scaled;
    uint32_t d_fallback;  // first draw of the Gaussian jump a degenerate DE proposal falls back to
};

__device__ __forceinline__ PtDraws pt_propose_draws(const PtConfig& cfg, unsigned iter, int r, int lane)
{
    const int n = lane < kPtNpars ? lane : kPtNpars - 1;  // idle lanes shadow the last parameter
    const unsigned long long seed = cfg.seed;
    const uint32_t id = (uint32_t)(r + cfg.ens_offset * cfg.n_temps);  // global rung id
    PtDraws w;
    const double alpha = pt_draw(seed, id, iter, 0u, 0u);
    w.jscale = pow(10., -6. + 6. * alpha);
    w.de = (pt_draw(seed, id, iter, 0u, 1u) < 0.5) && ((long long)iter > (long long)cfg.npast);
    w.a = 0; w.b = 0; w.scaled = 0; w.d_fallback = 0u;
    w.z_gauss = 0.; w.z_de = 0.;
    if (!w.de) {
        w.z_gauss = pt_normal(seed, id, iter, 0u, 2u, n);
    } else {
        uint32_t d = 2u;
        if (!cfg.quirks) w.a = (int)(pt_draw(seed, id, iter, 0u, d++) * cfg.npast);
        do { w.b = (int)(pt_draw(seed, id, iter, 0u, d++) * cfg.npast); } while (w.b == w.a);
        w.scaled = pt_draw(seed, id, iter, 0u, d++) < 0.9;
        if (w.scaled) w.z_de = pt_normal(seed, id, iter, 0u, d, n);
        w.d_fallback = d + (w.scaled ? 22u : 0u);
    }
    return w;
}

// The state-dependent part: rung r at temperature rung j, this lane's component xn of the current state, the rung's
// history ring, the draws of this iteration.  Returns this lane's component of y; lane 0 also gets the prior term and
// the jump type.  mcmc_wrapper2.c:390-481.
__device__ __forceinline__ double pt_propose_apply(const PtConfig& cfg, unsigned iter, int r, int j, int lane, const PtDraws& w,
                                                   const double xn, const double* __restrict__ hist, double& logP_out,
                                                   int& jump_out)
{
    const int n = lane < kPtNpars ? lane : kPtNpars - 1;
    const double sqtemp = sqrt(cfg.temp[j]);
    double yn;
    int jump_type = 1;
    if (!w.de) {
        yn = xn + w.z_gauss * cfg.sigma[n] * sqtemp * w.jscale;
    } else {
        const double eps_fac = cfg.quirks ? (pt_gaussian(0., 0., 1.e-4) - 0.5) : 0.0;
        HB_CHK(w.a, cfg.npast, 30);
        HB_CHK(w.b, cfg.npast, 30);
        double dx = hist[w.b * kPtNpars + n] - hist[w.a * kPtNpars + n];
        const double eps = dx * eps_fac;
        if (w.scaled) dx *= w.z_de * cfg.gamma;
        dx += eps;
        yn = xn + dx;
        const double dx_mag = warp_ordered_sum((xn - yn) * (xn - yn), kPtNpars);
        jump_type = 2;
        if (dx_mag < 1e-6) {  // mcmc_wrapper2.c:432-436; the Gaussian draws continue the stream
            const uint32_t id = (uint32_t)(r + cfg.ens_offset * cfg.n_temps);
            yn = xn + pt_normal(cfg.seed, id, iter, 0u, w.d_fallback, n) * cfg.sigma[n] * sqtemp * w.jscale;
            jump_type = 1;
        }
    }
    yn = pt_bound_one(yn, n, cfg);
    // mass ordering (quirk Q5), pinned period, T0 mod P (mcmc_wrapper2.c:470-481)
    const double y0 = __shfl_sync(0xffffffffu, yn, 0), y1 = __shfl_sync(0xffffffffu, yn, 1);
    if (y1 > y0) {
        if (cfg.quirks) { if (lane == 1) yn = y0; }
        else { if (lane == 0) yn = y1; else if (lane == 1) yn = y0; }
    }
    if (lane == 2) yn = cfg.log_lc_period;
    if (lane == 6) yn = fmod(yn, cfg.lc_period);
    // Gaussian priors (mcmc_wrapper2.c:703-765), summed in parameter order
    double mean, sig;
    pt_prior_of(n, mean, sig);
    const double term = (cfg.gauss[n] == 1) ? log(pt_gaussian(yn, mean, sig)) : 0.0;
    logP_out = warp_ordered_sum(term, kPtNpars);
    jump_out = jump_type;
    return yn;
}

// One walker's proposal by its warp (lane n owns parameter n): draws, then their application to the state.
__device__ __forceinline__ double pt_propose_warp(const PtConfig& cfg, unsigned iter, int r, int j, int lane,
                                                  const double xn, const double* __restrict__ hist,
                                                  double& logP_out, int& jump_out)
{
    const PtDraws w = pt_propose_draws(cfg, iter, r, lane);
    return pt_propose_apply(cfg, iter, r, j, lane, w, xn, hist, logP_out, jump_out);
}

__global__ void __launch_bounds__(128) k_pt_propose(const PtConfig* __restrict__ cfgp, const unsigned* __restrict__ iter_ptr,
                                                    const double* __restrict__ x, const int* __restrict__ index,
                                                    const double* __restrict__ history, double* __restrict__ y,
                                                    double* __restrict__ logPy, int* __restrict__ jump, int W)
{
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // local rung id = ens * T + j
    const int lane = threadIdx.x & 31;
    if (r >= W) return;
    const unsigned iter = *iter_ptr;
    const PtConfig& cfg = *cfgp;
    const int T = cfg.n_temps, ens = r / T, j = r - ens * T;
    const int c = ens * T + index[r];
    double lp;
    int jump_type;
    const double xn = x[(size_t)c * kPtNpars + (lane < kPtNpars ? lane : kPtNpars - 1)];
    const double yn = pt_propose_warp(cfg, iter, r, j, lane, xn, history + (size_t)r * cfg.npast * kPtNpars, lp, jump_type);
    if (lane < kPtNpars) y[(size_t)c * kPtNpars + lane] = yn;
    if (lane == 0) {
        logPy[c] = lp;
        jump[c] = jump_type;
    }
}

// counters per ensemble: 0 acc (chain slot 0 accepted, the reference's `acc`), 1 DE trials of slot 0,
// 2 DE accepted of slot 0, 3 accepted over all rungs, 4 proposals over all rungs, 5 swaps accepted,
// 6 swaps proposed, 7 iterations
// Metropolis-Hastings decision for the walker at local rung r (chain slot state xn per lane); mcmc_wrapper2.c:492-505.
// log prior of a state by its warp (lane n holds component n); get_logP, mcmc_wrapper2.c:703-765
__device__ __forceinline__ double pt_prior_warp(const PtConfig& cfg, int lane, double xn)
{
    const int n = lane < kPtNpars ? lane : kPtNpars - 1;
    double mean, sig;
    pt_prior_of(n, mean, sig);
    const double term = (cfg.gauss[n] == 1) ? log(pt_gaussian(xn, mean, sig)) : 0.0;
    return warp_ordered_sum(term, kPtNpars);
}

__device__ __forceinline__ bool pt_accept_warp(const PtConfig& cfg, unsigned iter, int r, int j, int lane, double xn,
                                               double logLx, double logLy, double logPy)
{
    const double logPx = pt_prior_warp(cfg, lane, xn);
    return pt_accept(cfg, (uint32_t)(r + cfg.ens_offset * cfg.n_temps), iter, cfg.temp[j], logLx, logLy, logPx, logPy);
}

// counters of one walker's step (lane 0 calls): see the list above
__device__ __forceinline__ void pt_count_step(unsigned long long* cnt, int slot, int jt, bool acc)
{
    if (slot == 0 && jt == 2) atomicAdd(&cnt[1], 1ull);
    atomicAdd(&cnt[4], 1ull);
    if (acc) {
        atomicAdd(&cnt[3], 1ull);
        if (slot == 0) {
            atomicAdd(&cnt[0], 1ull);
            if (jt == 2) atomicAdd(&cnt[2], 1ull);
        }
    }
}

__global__ void __launch_bounds__(128) k_pt_accept(const PtConfig* __restrict__ cfgp, const unsigned* __restrict__ iter_ptr, double* __restrict__ x,
                                                   const double* __restrict__ y, double* __restrict__ logLx,
                                                   const double* __restrict__ logLy, const double* __restrict__ logPy,
                                                   const int* __restrict__ jump, const int* __restrict__ index,
                                                   double* __restrict__ history, unsigned long long* __restrict__ counters,
                                                   int W)
{
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= W) return;
    const unsigned iter = *iter_ptr;
    const PtConfig& cfg = *cfgp;
    const int T = cfg.n_temps, ens = r / T, j = r - ens * T;
    const int slot = index[r], c = ens * T + slot;
    const int n = lane < kPtNpars ? lane : kPtNpars - 1;
    double xn = x[(size_t)c * kPtNpars + n];
    const bool acc = pt_accept_warp(cfg, iter, r, j, lane, xn, logLx[c], logLy[c], logPy[c]);
    const int jt = jump[c];
    if (acc) {
        xn = y[(size_t)c * kPtNpars + n];
        if (lane < kPtNpars) x[(size_t)c * kPtNpars + lane] = xn;
    }
    if (lane == 0) {
        pt_count_step(counters + (size_t)ens * 8, slot, jt, acc);
        if (acc) logLx[c] = logLy[c];
    }
    // history[j][iter % NPAST] = x[chain_id]  (mcmc_wrapper2.c:381,543-546)
    if (lane < kPtNpars) history[((size_t)r * cfg.npast + (iter % (unsigned)cfg.npast)) * kPtNpars + lane] = xn;
}

// The n_temps swap proposals of one ensemble by one warp (mcmc_wrapper2.c:554-563, ptmcmc :768-817): the lanes draw
// the (pair, beta) of all proposals in parallel (swap s consumes exactly block s of the ensemble's Philox stream,
// pt_swap_draws), lane 0 then applies them in order (pt_swap_apply).  s_idx (rung -> slot), s_logL (by slot) and
// s_dbeta (pt_stage_dbeta) are staged by the caller; on return s_idx holds the new permutation (every lane may read it
// after the trailing __syncwarp) and the accepted count is returned to lane 0.
// pt_stage_dbeta: (heat_b - heat_{b+1}) / (heat_b heat_{b+1}) of every adjacent pair, staged so that the serial loop
// touches shared memory only (the ladder lives in global memory: a dependent load per swap otherwise)
__device__ __forceinline__ void pt_stage_dbeta(const PtConfig& cfg, int lane, int nlanes, double* s_dbeta)
{
    for (int s = lane; s + 1 < cfg.n_temps; s += nlanes) {
        const double heat1 = cfg.temp[s + 1], heat2 = cfg.temp[s];
        s_dbeta[s] = (heat2 - heat1) / (heat2 * heat1);
    }
}

// The draws of the n_temps swap proposals of one ensemble at iteration `iter` (state-independent: a resident sampler
// takes them on an idle warp ahead of time): pair b and log of the acceptance draw per proposal, and the BATCHES of the
// proposal sequence -- maximal runs of consecutive proposals whose pairs {b, b + 1} are disjoint.  Proposals of a run
// commute, so pt_swap_apply decides a whole run at once, one lane per proposal (~5 proposals per run at 50 rungs).
// s_batch[i] .. s_batch[i + 1] - 1 are the proposals of run i; returns the number of runs (uniform).
__device__ __forceinline__ int pt_swap_draws(const PtConfig& cfg, unsigned iter, int ens_local, int lane, int* __restrict__ s_b,
                                             double* __restrict__ s_beta, int* __restrict__ s_batch)
{
    const int T = cfg.n_temps;
    for (int s = lane; s < T; s += 32) {
        U4 c; c.x = 0x80000000u | (uint32_t)(ens_local + cfg.ens_offset); c.y = iter; c.z = 2u; c.w = (uint32_t)s;
        const U4 r = philox4x32_10(c, (uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32));
        const double u0 = ((double)(((uint64_t)r.x << 21) | (r.y >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
        const double u1 = ((double)(((uint64_t)r.z << 21) | (r.w >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
        int b = (int)(u0 * (double)(T - 1));
        if (b > T - 2) b = T - 2;
        s_b[s] = b;
        s_beta[s] = log(u1);  // exp(x) >= beta  <=>  x >= log(beta): the log is taken here, in parallel
    }
    __syncwarp();
    int nb = 0;
    if (lane == 0 && T > 1) {
        unsigned long long occ[2] = {0ull, 0ull};  // rung positions taken by the current run (kPtMaxTemps = 128 bits)
        int len = 0;
        for (int s = 0; s < T; s++) {
            const int b = s_b[s];
            const unsigned long long m0 = (b < 64) ? (3ull << b) : 0ull;                                // bits b, b + 1
            const unsigned long long m1 = (b >= 64) ? (3ull << (b - 64)) : ((b == 63) ? 1ull : 0ull);
            if (len == 32 || ((occ[0] & m0) | (occ[1] & m1)) != 0ull || s == 0) {
                s_batch[nb++] = s;
                occ[0] = occ[1] = 0ull;
                len = 0;
            }
            occ[0] |= m0;
            occ[1] |= m1;
            len++;
        }
        s_batch[nb] = T;
    }
    return __shfl_sync(0xffffffffu, nb, 0);
}

// The proposals applied run by run (a lane per proposal of the run), the runs in order -- each decision depends on the
// permutation left by the previous runs.  s_pl is work space: the log-likelihood BY RUNG POSITION, carried along with
// the permutation, so that a step is one level of independent shared-memory loads, three FP64 instructions and the
// stores.  Same operands, same operations as logL[idx[b]] - logL[idx[a]] taken one proposal at a time: same decisions.
__device__ __forceinline__ int pt_swap_apply(const PtConfig& cfg, int lane, int n_batch, const int* __restrict__ s_batch,
                                             const int* __restrict__ s_b, const double* __restrict__ s_beta,
                                             const double* __restrict__ s_dbeta, int* __restrict__ s_idx,
                                             const double* __restrict__ s_logL, double* __restrict__ s_pl)
{
    const int T = cfg.n_temps;
    for (int s = lane; s < T; s += 32) s_pl[s] = s_logL[s_idx[s]];
    __syncwarp();
    int nacc = 0;
    int s0 = n_batch > 0 ? s_batch[0] : 0;
    for (int i = 0; i < n_batch; i++) {
        const int s1 = s_batch[i + 1];
        const int s = s0 + lane;
        bool acc = false;
        if (s < s1) {
            const int b = s_b[s], a = b + 1;
            const double Lb = s_pl[b], La = s_pl[a];
            const int ib = s_idx[b], ia = s_idx[a];
            const double lalpha = (Lb - La) * s_dbeta[b];
            acc = lalpha >= s_beta[s];
            if (acc) {
                s_pl[b] = La;
                s_pl[a] = Lb;
                s_idx[b] = ia;
                s_idx[a] = ib;
            }
        }
        nacc += __popc(__ballot_sync(0xffffffffu, acc));
        __syncwarp();
        s0 = s1;
    }
    return nacc;
}

__device__ __forceinline__ int pt_swap_warp(const PtConfig& cfg, unsigned iter, int ens_local, int lane, int* s_b, double* s_beta,
                                            const double* s_dbeta, int* s_idx, const double* s_logL, double* s_pl, int* s_batch)
{
    const int nb = pt_swap_draws(cfg, iter, ens_local, lane, s_b, s_beta, s_batch);
    __syncwarp();
    return pt_swap_apply(cfg, lane, nb, s_batch, s_b, s_beta, s_dbeta, s_idx, s_logL, s_pl);
}

// One warp per ensemble: the lanes draw the (pair, beta) of all n_temps swap proposals in parallel
// (swap s consumes exactly block s of the ensemble's Philox stream), lane 0 then applies them in
// order -- each decision depends on the permutation left by the previous one (mcmc_wrapper2.c:554-563).
__global__ void __launch_bounds__(32) k_pt_swap(const PtConfig* __restrict__ cfgp, unsigned* __restrict__ iter_ptr,
                                                int* __restrict__ index, const double* __restrict__ logLx,
                                                const double* __restrict__ x, unsigned long long* __restrict__ counters,
                                                double* __restrict__ xmap, double* __restrict__ logLmap, int E)
{
    const int ens = blockIdx.x, lane = threadIdx.x;
    if (ens >= E) return;
    const unsigned iter = iter_ptr[0];
    const PtConfig& cfg = *cfgp;
    const int T = cfg.n_temps;
    __shared__ int s_b[kPtMaxTemps];
    __shared__ double s_beta[kPtMaxTemps];  // log of the acceptance draw
    __shared__ int s_idx[kPtMaxTemps];
    __shared__ double s_logL[kPtMaxTemps];
    __shared__ double s_dbeta[kPtMaxTemps];
    __shared__ double s_pl[kPtMaxTemps];
    __shared__ int s_batch[kPtMaxTemps + 1];
    pt_stage_dbeta(cfg, lane, 32, s_dbeta);
    for (int s = lane; s < T; s += 32) {
        s_idx[s] = index[(size_t)ens * T + s];
        s_logL[s] = logLx[(size_t)ens * T + s];
    }
    __syncwarp();
    const int nacc = pt_swap_warp(cfg, iter, ens, lane, s_b, s_beta, s_dbeta, s_idx, s_logL, s_pl, s_batch);
    if (lane == 0) {
        unsigned long long* cnt = counters + (size_t)ens * 8;
        cnt[5] += (unsigned long long)nacc;
        cnt[6] += (unsigned long long)T;
        cnt[7] += 1ull;
    }
    for (int s = lane; s < T; s += 32) index[(size_t)ens * T + s] = s_idx[s];
    // MAP of the cold rung (mcmc_wrapper2.c:565-572)
    const int c0 = ens * T + s_idx[0];
    const bool better = logLx[c0] > logLmap[ens];
    __syncwarp();
    if (better) {
        if (lane < kPtNpars) xmap[(size_t)ens * kPtNpars + lane] = x[(size_t)c0 * kPtNpars + lane];
        if (lane == 0) logLmap[ens] = logLx[c0];
    }
    // The iteration counter lives on the device so that a captured CUDA graph of one step can be
    // replayed: the last ensemble to finish advances it (iter_ptr[1] is the arrival ticket).
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(&iter_ptr[1], 1u) == (unsigned)E - 1u) {
            iter_ptr[1] = 0u;
            iter_ptr[0] = iter + 1u;
        }
    }
}

// gather the cold-rung state of every ensemble: out_x[E][21], out_logL[E]
__global__ void k_pt_gather_cold(const PtConfig* __restrict__ cfgp, const int* __restrict__ index,
                                 const double* __restrict__ x, const double* __restrict__ logLx,
                                 double* __restrict__ out_x, double* __restrict__ out_logL, int E)
{
    const int ens = blockIdx.x * blockDim.x + threadIdx.x;
    if (ens >= E) return;
    const int T = cfgp->n_temps;
    const int c0 = ens * T + index[(size_t)ens * T];
    out_logL[ens] = logLx[c0];
    for (int i = 0; i < kPtNpars; i++) out_x[(size_t)ens * kPtNpars + i] = x[(size_t)c0 * kPtNpars + i];
}

// logL by rung (what logL.*.dat prints, mcmc_wrapper2.c:608-611)
__global__ void k_pt_logL_by_rung(const PtConfig* __restrict__ cfgp, const int* __restrict__ index,
                                  const double* __restrict__ logLx, double* __restrict__ out, int W)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= W) return;
    const int T = cfgp->n_temps, ens = r / T;
    out[r] = logLx[ens * T + index[r]];
}


// ---------------------------------------------------------------------------------------------------------------
// The whole step loop in ONE launch, for ladders whose light curves are short (the reference's real, folded light
// curves have 163-763 points): at that size an iteration is a few microseconds of arithmetic and the five launches
// of the stream-ordered path (propose, prologue, likelihood, accept, swap) are all latency.  Here every walker
// (ensemble, rung) owns a CTA for the whole run: warp 0 proposes (lane = parameter) and folds the proposal's
// constants (chain prologue, lanes = libm calls), the CTA evaluates the light curve and its exact median from
// shared memory, warp 0 accepts; ONE grid-wide barrier per iteration publishes the new log-likelihoods, after which
// every CTA of a ladder replays the ladder's n_temps swap proposals for itself (the permutation lives in each CTA's
// shared memory; rung 0's CTA keeps the counters and the MAP).  Same device functions, same order of operations as
// the stream-ordered kernels: the chains are identical bit for bit (tests/test_gpu_pt.py).  mcmc_wrapper2.c:378-572.
struct PtRunShared {
    ChainConst cc;
    SelectCtl<kEvalThreads> ctl;
    double red[32];
    double y[kPtNpars + 3];
    double logPy, logLy;
    int jump;
    int s_b[kPtMaxTemps];
    int s_batch[kPtMaxTemps + 1];  // runs of commuting swap proposals (pt_swap_draws)
    int n_batch;
    int s_idx[kPtMaxTemps];
    double s_beta[kPtMaxTemps], s_dbeta[kPtMaxTemps], s_logL[kPtMaxTemps], s_pl[kPtMaxTemps];
    PtDraws draws[2][32];  // the proposal's random part per lane: [iteration parity] (drawn one iteration ahead)
    double2 sctab[kSinTabN];
    uint64_t keys[kPtRunMaxPoints];  // template keys of the proposal's light curve
    uint64_t bufA[kPtRunMaxPoints], bufB[kPtRunMaxPoints];  // survivor buffers of the select
};

size_t pt_run_smem_bytes() { return sizeof(PtRunShared); }

// likelihood of the chain whose constants sit in sm.cc (the small-light-curve path of k_chain_eval: template stored,
// exact order statistic, chi^2 summed in the reference's own form, likelihood3.c:681-685,809-873)
__device__ __forceinline__ double pt_run_loglike(PtRunShared& sm, const PtRunArgs& a, int tid, long long* prof = nullptr)
{
    [[maybe_unused]] long long q0 = clock64();
    const ChainConst& cc = sm.cc;
    const int lane = tid & 31, wid = tid >> 5;
    const int N = a.N;
    const int flag = (int)cc.flag;
    const bool roche = flag & 1, nan_model = flag & 2;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    if (nan_model || roche || N <= 0) return roche ? -0.5 * kBig : (nan_model ? qnan : -0.5 * cc.chi2_extra);  // (uniform; Q13)
    if (tid == 0 && a.evaluated != nullptr) atomicAdd(a.evaluated, 1ull);
    int nanflag = 0;
    const int n_tiles = (N + kEvalThreads - 1) / kEvalThreads;
    int krank = (N % 2 == 0) ? N / 2 : N / 2 + 1;  // likelihood3.c:97-101 (quirk Q3)
    if (krank > N - 1) krank = N - 1;
    const double blend = cc.blend, ft = cc.ft;
    double S0 = 0.;
    if (n_tiles <= 2) {
        // the reference's folded light curves (163-763 points; here up to 512): two samples per thread, template
        // and data in registers, the whole template sorted in one go (block_sort2)
        double2 v[2];
        uint64_t key[2];
#pragma unroll
        for (int tile = 0; tile < 2; tile++) {
            const int i = tile * kEvalThreads + tid;
            key[tile] = ~0ull;  // padding sorts above every number
            v[tile] = make_double2(0., 0.);
            if (tile < n_tiles) {  // (uniform)
                const double ts[1] = {a.tsec[i]};  // (padded to whole tiles)
                v[tile] = a.fw[i];                 // (likewise, weight 0) -- asked for ahead of the sort that hides it
                double u[1];
                raw_flux<1, true, true, false>(cc, nullptr, sm.sctab, ts, u);
                const bool valid = i < N;
                nanflag |= valid & (u[0] != u[0]);
                if (valid) key[tile] = dkey(u[0]);
            }
        }
        if (__syncthreads_or(nanflag)) return qnan;
        if (prof) { long long q1 = clock64(); prof[0] += q1 - q0; q0 = q1; }
        const KeyPair srt = block_sort2<kEvalThreads>(key[0], key[1], sm.bufA);
        if (tid == (krank & (kEvalThreads - 1))) sm.ctl.result = krank < kEvalThreads ? srt.a : srt.b;
        __syncthreads();
        const double med = dunkey(sm.ctl.result);
        if (prof) { long long q1 = clock64(); prof[1] += q1 - q0; q0 = q1; }
#pragma unroll
        for (int tile = 0; tile < 2; tile++) {  // in the order of the strided loop below
            const int i = tile * kEvalThreads + tid;
            if (i < N) {
                const double r = (finish_template(dunkey(key[tile]), med, blend, ft) - v[tile].x) * v[tile].y;
                S0 = fma(r, r, S0);
            }
        }
    } else {
        for (int tile = 0; tile < n_tiles; tile++) {
            const int i = tile * kEvalThreads + tid;
            const double ts[1] = {a.tsec[i]};  // (padded to whole tiles)
            double u[1];
            raw_flux<1, true, true, false>(cc, nullptr, sm.sctab, ts, u);
            const bool valid = i < N;
            nanflag |= valid & (u[0] != u[0]);
            if (valid) sm.keys[i] = dkey(u[0]);
        }
        if (__syncthreads_or(nanflag)) return qnan;
        if (prof) { long long q1 = clock64(); prof[0] += q1 - q0; q0 = q1; }
        const SelectBuf bufs[2] = {{sm.bufA, kPtRunMaxPoints}, {sm.bufB, kPtRunMaxPoints}};
        const double med = dunkey(block_select_key<kEvalThreads>(sm.keys, N, krank, sm.ctl, bufs, 2, (uint32_t)cc.seed ^ 0x9e3779b9u));
        if (prof) { long long q1 = clock64(); prof[1] += q1 - q0; q0 = q1; }
        for (int i = tid; i < N; i += kEvalThreads) {
            const double2 v = a.fw[i];
            const double r = (finish_template(dunkey(sm.keys[i]), med, blend, ft) - v.x) * v.y;
            S0 = fma(r, r, S0);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) S0 += __shfl_xor_sync(0xffffffffu, S0, o);
    if (lane == 0) sm.red[wid] = S0;
    __syncthreads();
    double t0 = 0.;
    for (int i = 0; i < kEvalThreads / 32; i++) t0 += sm.red[i];
    __syncthreads();
    if (prof) { long long q1 = clock64(); prof[2] += q1 - q0; prof[3] += 1; }
    return -0.5 * (t0 + cc.chi2_extra);
}

#ifndef HB_PT_RUN_MIN_BLOCKS
#define HB_PT_RUN_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(kEvalThreads, HB_PT_RUN_MIN_BLOCKS) k_pt_run(const __grid_constant__ PtRunArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PtRunShared& sm = *reinterpret_cast<PtRunShared*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const PtConfig& cfg = *a.cfg;
    const int T = cfg.n_temps;
    const int r = blockIdx.x, ens = r / T, j = r - ens * T;  // this CTA's walker: local rung id, ladder, rung
    const int n = lane < kPtNpars ? lane : kPtNpars - 1;
    for (int i = tid; i < kSinTabN; i += kEvalThreads) sm.sctab[i] = a.sctab[i];
    for (int s = tid; s < T; s += kEvalThreads) sm.s_idx[s] = a.index[(size_t)ens * T + s];
    pt_stage_dbeta(cfg, tid, kEvalThreads, sm.s_dbeta);  // once per launch
    __syncthreads();
    if (wid == 0) {  // prior of the state this walker starts from (every slot is some rung's: all get theirs)
        const int c = ens * T + sm.s_idx[j];
        const double lp = pt_prior_warp(cfg, lane, a.x[(size_t)c * kPtNpars + n]);
        if (lane == 0) a.logPx[c] = lp;
    }
    unsigned iter = a.d_iter[0];
    const double* hist = a.history + (size_t)r * cfg.npast * kPtNpars;
#ifdef HB_PT_PROF
    long long tp[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tq[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = clock64(), t1;
    int nbig[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long lprof[4] = {0, 0, 0, 0};
#define PT_MARK(k) do { t1 = clock64(); tp[k] += t1 - t0; if (t1 - t0 > tq[k]) tq[k] = t1 - t0; if (k == 2 && t1 - t0 > 12000) nbig[2]++; if (k == 0 && t1 - t0 > 10000) nbig[0]++; if (k == 6 && t1 - t0 > 15000) nbig[6]++; t0 = t1; } while (0)
#else
#define PT_MARK(k) ((void)0)
#endif
    // The random numbers of an iteration do not depend on the walkers' states: while warp 0 works through the
    // dependent chain of iteration i (proposal -> constants), warp 1 draws the swap proposals of iteration i and warp 2
    // the proposal of iteration i + 1 (the barriers of the iteration publish both).
    if (wid == 2) sm.draws[iter & 1u][lane] = pt_propose_draws(cfg, iter, r, lane);
    __syncthreads();
    for (long it = 0; it < a.n_iters; it++, iter++) {
        const int slot = sm.s_idx[j], c = ens * T + slot;
        double xn = 0., logLx = 0., logPx = 0.;  // (warp 0) the walker's current state: read once, ahead of their use
        if (wid == 1) {
            const int nb = pt_swap_draws(cfg, iter, ens, lane, sm.s_b, sm.s_beta, sm.s_batch);
            if (lane == 0) sm.n_batch = nb;
        }
        if (wid == 2) sm.draws[(iter + 1u) & 1u][lane] = pt_propose_draws(cfg, iter + 1u, r, lane);
        if (wid == 0) {  // proposal
            xn = __ldcg(&a.x[(size_t)c * kPtNpars + n]);  // (written by another CTA when the slot changed hands)
            logLx = __ldcg(&a.logLx[c]);
            // the prior of the current state is the prior its proposal had when it was accepted: kept per slot
            // (same function of the same numbers as the stream-ordered kernel evaluates afresh: same bits)
            logPx = __ldcg(&a.logPx[c]);
            double lp;
            int jt;
            const double yn = pt_propose_apply(cfg, iter, r, j, lane, sm.draws[iter & 1u][lane], xn, hist, lp, jt);
            if (lane < kPtNpars) {
                sm.y[lane] = yn;
                a.y[(size_t)c * kPtNpars + lane] = yn;
            }
            if (lane == 0) {
                sm.logPy = lp;
                sm.jump = jt;
                a.logPy[c] = lp;
                a.jump[c] = jt;
            }
        }
        __syncthreads();
        PT_MARK(0);
        // the proposal's folded constants (chain prologue): four warps each take the libm levels (lanes = calls; the
        // same calls on every warp, side by side on the four sub-partitions) and then one section of the assembly
        if (wid < 4) {
            PrologueT P;
            prologue_trans_warp(sm.y, a.ms, P, lane);
            PT_MARK(6);
            if (lane == 0) {
                if (wid == 0) prologue_assemble_sections<kAsmStarA>(sm.y, a.ms, P, sm.cc);
                else if (wid == 1) prologue_assemble_sections<kAsmStarB>(sm.y, a.ms, P, sm.cc);
                else if (wid == 2) prologue_assemble_sections<kAsmOrbit>(sm.y, a.ms, P, sm.cc);
                else prologue_assemble_sections<kAsmAux>(sm.y, a.ms, P, sm.cc);
            }
        }
        __syncthreads();
        PT_MARK(1);
#ifdef HB_PT_PROF
        const double logLy = pt_run_loglike(sm, a, tid, lprof);
#else
        const double logLy = pt_run_loglike(sm, a, tid);
#endif
        PT_MARK(2);
        if (wid == 0) {  // accept / reject, history ring
            const bool acc = pt_accept(cfg, (uint32_t)(r + cfg.ens_offset * T), iter, cfg.temp[j], logLx, logLy, logPx, sm.logPy);
            if (acc) {
                xn = sm.y[n];
                if (lane < kPtNpars) a.x[(size_t)c * kPtNpars + lane] = xn;
            }
            if (lane == 0) {
                a.logLy[c] = logLy;
                pt_count_step(a.counters + (size_t)ens * 8, slot, sm.jump, acc);
                if (acc) {
                    a.logLx[c] = logLy;
                    a.logPx[c] = sm.logPy;
                }
            }
            if (lane < kPtNpars) a.history[((size_t)r * cfg.npast + (iter % (unsigned)cfg.npast)) * kPtNpars + lane] = xn;
        }
        PT_MARK(3);
        // ---- every walker's new state and log-likelihood are published: one grid-wide barrier per iteration ----
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            atomicAdd(a.barrier, 1u);
            const unsigned want = (unsigned)(it + 1) * gridDim.x;
            while (*((volatile unsigned*)a.barrier) < want) {}
            __threadfence();
        }
        __syncthreads();
        PT_MARK(4);
        if (wid == 0) {  // the ladder's swap proposals, replayed identically by each of its CTAs
            for (int s = lane; s < T; s += 32) sm.s_logL[s] = __ldcg(&a.logLx[(size_t)ens * T + s]);
            __syncwarp();
            const int nacc = pt_swap_apply(cfg, lane, sm.n_batch, sm.s_batch, sm.s_b, sm.s_beta, sm.s_dbeta, sm.s_idx, sm.s_logL, sm.s_pl);
            if (j == 0) {  // rung 0's CTA keeps the ladder's books: counters and the MAP (mcmc_wrapper2.c:565-572)
                if (lane == 0) {
                    unsigned long long* cnt = a.counters + (size_t)ens * 8;
                    cnt[5] += (unsigned long long)nacc;
                    cnt[6] += (unsigned long long)T;
                    cnt[7] += 1ull;
                }
                const int s0 = sm.s_idx[0], c0 = ens * T + s0;
                const bool better = sm.s_logL[s0] > a.logLmap[ens];
                __syncwarp();
                if (better) {
                    if (lane < kPtNpars) a.xmap[(size_t)ens * kPtNpars + lane] = __ldcg(&a.x[(size_t)c0 * kPtNpars + lane]);
                    if (lane == 0) a.logLmap[ens] = sm.s_logL[s0];
                }
            }
        }
        __syncthreads();
        PT_MARK(5);
    }
#ifdef HB_PT_PROF
    if (tid == 0 && a.n_iters >= 1000)
        printf("k_pt_run block %d: cycles per iteration  propose %lld  prologue: libm levels %lld + assembly %lld  loglike %lld  accept %lld  barrier %lld  swap %lld | max: propose %lld libm %lld asm %lld loglike %lld accept %lld | share of iterations: propose > 10k %.3f  libm > 15k %.3f  loglike > 12k %.3f\n",
               (int)blockIdx.x, tp[0] / a.n_iters, tp[6] / a.n_iters, tp[1] / a.n_iters, tp[2] / a.n_iters, tp[3] / a.n_iters, tp[4] / a.n_iters, tp[5] / a.n_iters,
               tq[0], tq[6], tq[1], tq[2], tq[3], (double)nbig[0] / a.n_iters, (double)nbig[6] / a.n_iters, (double)nbig[2] / a.n_iters);
    if (tid == 0 && a.n_iters >= 1000 && lprof[3] > 0)
        printf("   block %d evaluated %lld times: model %lld  select %lld  chi2 %lld cycles\n", (int)blockIdx.x, lprof[3], lprof[0] / lprof[3], lprof[1] / lprof[3], lprof[2] / lprof[3]);
#endif
    if (j == 0)
        for (int s = tid; s < T; s += kEvalThreads) a.index[(size_t)ens * T + s] = sm.s_idx[s];
    if (blockIdx.x == 0 && tid == 0) a.d_iter[0] = iter;
}

cudaError_t configure_pt_run()
{
    return cudaFuncSetAttribute(k_pt_run, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PtRunShared));
}

// the most walkers one launch can hold (every CTA must be resident for the grid barrier)
cudaError_t pt_run_max_walkers(int sm_count, int* out)
{
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pt_run, kEvalThreads, sizeof(PtRunShared));
    *out = per_sm * sm_count;
    return e;
}

cudaError_t launch_pt_run(const PtRunArgs& a, int W, cudaStream_t s)
{
    void* params[1] = {const_cast<PtRunArgs*>(&a)};
    return cudaLaunchCooperativeKernel((const void*)k_pt_run, dim3(W), dim3(kEvalThreads), params, sizeof(PtRunShared), s);
}

#define LAUNCH1D(kern, n, s, ...)                                         \
    do {                                                                  \
        if ((n) > 0) kern<<<((n) + 127) / 128, 128, 0, s>>>(__VA_ARGS__); \
        return cudaGetLastError();                                        \
    } while (0)

cudaError_t launch_pt_init_random(const PtConfig* cfg, double* x, int W, cudaStream_t s) { LAUNCH1D(k_pt_init_random, W, s, cfg, x, W); }
cudaError_t launch_pt_propose(const PtConfig* cfg, const unsigned* iter, const double* x, const int* index, const double* history,
                              double* y, double* logPy, int* jump, int W, cudaStream_t s)
{
    if (W > 0) k_pt_propose<<<(W + 3) / 4, 128, 0, s>>>(cfg, iter, x, index, history, y, logPy, jump, W);
    return cudaGetLastError();
}
cudaError_t launch_pt_accept(const PtConfig* cfg, const unsigned* iter, double* x, const double* y, double* logLx,
                             const double* logLy, const double* logPy, const int* jump, const int* index, double* history,
                             unsigned long long* counters, int W, cudaStream_t s)
{
    if (W > 0) k_pt_accept<<<(W + 3) / 4, 128, 0, s>>>(cfg, iter, x, y, logLx, logLy, logPy, jump, index, history, counters, W);
    return cudaGetLastError();
}
cudaError_t launch_pt_swap(const PtConfig* cfg, unsigned* iter, int* index, const double* logLx, const double* x,
                           unsigned long long* counters, double* xmap, double* logLmap, int E, cudaStream_t s)
{
    if (E > 0) k_pt_swap<<<E, 32, 0, s>>>(cfg, iter, index, logLx, x, counters, xmap, logLmap, E);
    return cudaGetLastError();
}
cudaError_t launch_pt_gather_cold(const PtConfig* cfg, const int* index, const double* x, const double* logLx,
                                  double* out_x, double* out_logL, int E, cudaStream_t s)
{
    LAUNCH1D(k_pt_gather_cold, E, s, cfg, index, x, logLx, out_x, out_logL, E);
}
cudaError_t launch_pt_logL_by_rung(const PtConfig* cfg, const int* index, const double* logLx, double* out, int W,
                                   cudaStream_t s)
{
    LAUNCH1D(k_pt_logL_by_rung, W, s, cfg, index, logLx, out, W);
}

}  // namespace hb
