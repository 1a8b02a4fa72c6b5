// hb_pt.cu -- kernels of the device-resident parallel-tempering sampler (see hb_pt.cuh for the
// algorithm, its reference citations and the positions on the reference's sampler bugs).
//
// Layout: E ensembles (independent PT ladders) x T rungs.  Chain state x[E*T][21], logL[E*T] is
// indexed by chain SLOT; index[E][T] maps rung -> slot inside the ensemble (the reference swaps
// this permutation, not the states, mcmc_wrapper2.c:812-816); the DE history ring is per RUNG
// (history[j], mcmc_wrapper2.c:424,543-546).  One iteration = k_pt_propose -> k_prologue ->
// k_chain_eval (the likelihood of all E*T proposals, one batch) -> k_pt_accept -> k_pt_swap.
#include "hb_kernels.h"
#include "hb_select.cuh"
#include "hb_pt.cuh"

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

__global__ void __launch_bounds__(128) k_pt_propose(const PtConfig* __restrict__ cfgp, const unsigned* __restrict__ iter_ptr,
                                                    const double* __restrict__ x, const int* __restrict__ index,
                                                    const double* __restrict__ history, double* __restrict__ y,
                                                    double* __restrict__ logPy, int* __restrict__ jump, int W)
{
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // global rung id = ens * T + j
    const int lane = threadIdx.x & 31;
    if (r >= W) return;
    const unsigned iter = *iter_ptr;
    const PtConfig& cfg = *cfgp;
    const int T = cfg.n_temps, ens = r / T, j = r - ens * T;
    const int c = ens * T + index[r];
    const int n = lane < kPtNpars ? lane : kPtNpars - 1;  // idle lanes shadow the last parameter
    const unsigned long long seed = cfg.seed;
    const uint32_t id = (uint32_t)(r + cfg.ens_offset * T);  // global rung id
    const double temp = cfg.temp[j];
    const double xn = x[(size_t)c * kPtNpars + n];
    const double* hist = history + (size_t)r * cfg.npast * kPtNpars;

    const double alpha = pt_draw(seed, id, iter, 0u, 0u);
    const double jscale = pow(10., -6. + 6. * alpha);
    const bool de = (pt_draw(seed, id, iter, 0u, 1u) < 0.5) && ((long long)iter > (long long)cfg.npast);
    const double sqtemp = sqrt(temp);
    double yn;
    int jump_type = 1;
    if (!de) {
        yn = xn + pt_normal(seed, id, iter, 0u, 2u, n) * cfg.sigma[n] * sqtemp * jscale;
    } else {
        uint32_t d = 2u;
        int a = 0, b;
        if (!cfg.quirks) a = (int)(pt_draw(seed, id, iter, 0u, d++) * cfg.npast);
        do { b = (int)(pt_draw(seed, id, iter, 0u, d++) * cfg.npast); } while (b == a);
        const bool scaled = pt_draw(seed, id, iter, 0u, d++) < 0.9;
        const double eps_fac = cfg.quirks ? (pt_gaussian(0., 0., 1.e-4) - 0.5) : 0.0;
        HB_CHK(a, cfg.npast, 30);
        HB_CHK(b, cfg.npast, 30);
        double dx = hist[b * kPtNpars + n] - hist[a * kPtNpars + n];
        const double eps = dx * eps_fac;
        if (scaled) dx *= pt_normal(seed, id, iter, 0u, d, n) * cfg.gamma;
        dx += eps;
        yn = xn + dx;
        const double dx_mag = warp_ordered_sum((xn - yn) * (xn - yn), kPtNpars);
        jump_type = 2;
        if (dx_mag < 1e-6) {  // mcmc_wrapper2.c:432-436; the Gaussian draws continue the stream
            const uint32_t d2 = d + (scaled ? 22u : 0u);
            yn = xn + pt_normal(seed, id, iter, 0u, d2, n) * cfg.sigma[n] * sqtemp * jscale;
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
    const double lp = warp_ordered_sum(term, kPtNpars);
    if (lane < kPtNpars) y[(size_t)c * kPtNpars + lane] = yn;
    if (lane == 0) {
        logPy[c] = lp;
        jump[c] = jump_type;
    }
}

// counters per ensemble: 0 acc (chain slot 0 accepted, the reference's `acc`), 1 DE trials of slot 0,
// 2 DE accepted of slot 0, 3 accepted over all rungs, 4 proposals over all rungs, 5 swaps accepted,
// 6 swaps proposed, 7 iterations
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
    double mean, sig;
    pt_prior_of(n, mean, sig);
    const double term = (cfg.gauss[n] == 1) ? log(pt_gaussian(xn, mean, sig)) : 0.0;
    const double logPx = warp_ordered_sum(term, kPtNpars);
    const bool acc = pt_accept(cfg, (uint32_t)(r + cfg.ens_offset * T), iter, cfg.temp[j], logLx[c], logLy[c], logPx, logPy[c]);
    const int jt = jump[c];
    if (acc) {
        xn = y[(size_t)c * kPtNpars + n];
        if (lane < kPtNpars) x[(size_t)c * kPtNpars + lane] = xn;
    }
    if (lane == 0) {
        unsigned long long* cnt = counters + (size_t)ens * 8;
        if (slot == 0 && jt == 2) atomicAdd(&cnt[1], 1ull);
        atomicAdd(&cnt[4], 1ull);
        if (acc) {
            logLx[c] = logLy[c];
            atomicAdd(&cnt[3], 1ull);
            if (slot == 0) {
                atomicAdd(&cnt[0], 1ull);
                if (jt == 2) atomicAdd(&cnt[2], 1ull);
            }
        }
    }
    // history[j][iter % NPAST] = x[chain_id]  (mcmc_wrapper2.c:381,543-546)
    if (lane < kPtNpars) history[((size_t)r * cfg.npast + (iter % (unsigned)cfg.npast)) * kPtNpars + lane] = xn;
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
    // (heat_b - heat_{b+1}) / (heat_b heat_{b+1}) of every adjacent pair, staged so that the serial loop below
    // touches shared memory only (the ladder lives in global memory: a dependent load per swap otherwise)
    __shared__ double s_dbeta[kPtMaxTemps];
    for (int s = lane; s + 1 < T; s += 32) {
        const double heat1 = cfg.temp[s + 1], heat2 = cfg.temp[s];
        s_dbeta[s] = (heat2 - heat1) / (heat2 * heat1);
    }
    for (int s = lane; s < T; s += 32) {
        U4 c; c.x = 0x80000000u | (uint32_t)(ens + cfg.ens_offset); c.y = iter; c.z = 2u; c.w = (uint32_t)s;
        const U4 r = philox4x32_10(c, (uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32));
        const double u0 = ((double)(((uint64_t)r.x << 21) | (r.y >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
        const double u1 = ((double)(((uint64_t)r.z << 21) | (r.w >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
        int b = (int)(u0 * (double)(T - 1));
        if (b > T - 2) b = T - 2;
        s_b[s] = b;
        s_beta[s] = log(u1);  // exp(x) >= beta  <=>  x >= log(beta): the log is taken here, in parallel
        s_idx[s] = index[(size_t)ens * T + s];
        s_logL[s] = logLx[(size_t)ens * T + s];
    }
    __syncwarp();
    if (lane == 0) {
        int nacc = 0;
        for (int s = 0; s < T && T > 1; s++) {
            const int b = s_b[s], a = b + 1;
            const int olda = s_idx[a], oldb = s_idx[b];
            const double lalpha = (s_logL[oldb] - s_logL[olda]) * s_dbeta[b];
            if (lalpha >= s_beta[s]) {
                s_idx[a] = oldb;
                s_idx[b] = olda;
                nacc++;
            }
        }
        unsigned long long* cnt = counters + (size_t)ens * 8;
        cnt[5] += (unsigned long long)nacc;
        cnt[6] += (unsigned long long)T;
        cnt[7] += 1ull;
    }
    __syncwarp();
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
