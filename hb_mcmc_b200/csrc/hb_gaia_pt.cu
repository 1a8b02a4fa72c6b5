// hb_gaia_pt.cu -- the Gaia-colour sampler of GAIA_mcmc.c as ONE persistent kernel.
//
// The reference (GAIA_mcmc.c:663-780) walks 20 rungs x 6 parameters against a 4-point likelihood
// (G, B-V, V-G, G-T; get_mags :198-250, model_likelihood :255-269): a step costs a few dozen libm
// calls, so a kernel launch per step would be all latency.  Here a whole run lives on the device:
// one WARP per ensemble (independent ladder: another star, or another seed of the same star),
// one LANE per rung (n_temps <= 32), the ladder's state in shared memory, the loop over iterations
// inside the kernel, the thinned chain written straight to the output buffers.  Ensembles never
// communicate, so the grid scales over SMs and over GPUs without any collective.
//
// Per iteration and lane (= rung j), following run_chain (:518-590):
//   propose   jump scale 10^(-6+6u); DE (:1004-1026) with probability 1/2 once iter > NPAST, else
//             Gaussian (:987-1002); a DE jump shorter than 1e-3 falls back to Gaussian (:541-542)
//   bounds    single reflection / wrap per side (:546-559 -- `if`, not the driver's `while`)
//   prior     log10 of a Gaussian WITHOUT the 1/2 (:168-191), centred on the box, sigma = half box
//   accept    alpha <= exp(dlogL / T_j) * 10^(dlogP)  (:569-583)
// then, for k = 0..T-1 (run_mcmc :741-748): one swap proposal of a random adjacent pair (ptmcmc,
// :893-938), followed by the fill of rung k's history ring from the slot index[k] holds right then.
//
// Differences kept deliberately (documented in DESIGN.md): Philox streams (id, iter, stage) replace
// GSL ranlxs1 / libc rand(); the current-state likelihood is cached instead of re-evaluated
// (:565 -- same value); logLmap starts at the first cold likelihood (the reference passes it by
// value to init_chain and reads it uninitialised afterwards, :480-487,751).
#include "hb_device.cuh"
#include "hb_gaia_pt.cuh"
#include "hb_kernels.h"
#include "hb_pt.cuh"

namespace hb {

// GAIA_mcmc.c:168-172
__device__ __forceinline__ double gaia_gaussian(double x, double mean, double sigma)
{
    const double r = (x - mean) / sigma;  // pow(r, 2.) of the reference is the correctly rounded r * r
    return (1 / sigma / 2.5066282746) * exp(-(r * r));
}

// GAIA_mcmc.c:175-191
__device__ double gaia_log_prior(const double* p, const GaiaPtConfig& cfg)
{
    double logP = 0.;
    for (int i = 0; i < kGaiaNpars; i++) {
        if (cfg.gauss[i] == 1) {
            const double mean = 0.5 * (cfg.lo[i] + cfg.hi[i]);
            const double sigma = (cfg.hi[i] - cfg.lo[i]) / 2.;
            logP += log10(gaia_gaussian(p[i], mean, sigma));
        }
    }
    return logP;
}

// GAIA_mcmc.c:255-269
__device__ double gaia_logL(const double* p, double D, const double* data, const double* err)
{
    double m[4];
    gaia_mags(p, D, m);
    double chi2 = 0.;
    for (int i = 0; i < 4; i++) {
        const double r = (data[i] - m[i]) / err[i];
        chi2 += r * r;
    }
    return (-chi2 / 2.0);
}

// three Box-Muller pairs in component order (the stream order of the oracle's gaia_gaussian_jump)
__device__ __forceinline__ void gaia_normals(PtRng& g, double z[kGaiaNpars])
{
    for (int n = 0; n < kGaiaNpars; n += 2) g.normal2(z[n], z[n + 1]);
}

__global__ void k_gaia_pt_init(const GaiaPtConfig* __restrict__ cfgp, double* __restrict__ x, int W)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const GaiaPtConfig& cfg = *cfgp;
    PtRng g;
    g.init(cfg.seed, (uint32_t)w, 0xFFFFFFFFu, 3u);
    // uniform in the prior box (init_chain, GAIA_mcmc.c:463-473)
    for (int i = 0; i < kGaiaNpars; i++) x[(size_t)w * kGaiaNpars + i] = cfg.lo[i] + g.next() * (cfg.hi[i] - cfg.lo[i]);
}

// logL of every chain slot + per-ensemble MAP reset to the cold rung
__global__ void k_gaia_pt_eval(const GaiaPtConfig* __restrict__ cfgp, const double* __restrict__ x,
                               const double* __restrict__ D, const double* __restrict__ data,
                               const double* __restrict__ err, const int* __restrict__ index, double* __restrict__ logL,
                               double* __restrict__ xmap, double* __restrict__ logLmap, int W)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const int T = cfgp->n_temps, ens = w / T;
    const double l = gaia_logL(x + (size_t)w * kGaiaNpars, D[ens], data + 4 * ens, err + 4 * ens);
    logL[w] = l;
    if (w - ens * T == index[(size_t)ens * T]) {
        logLmap[ens] = l;
        for (int i = 0; i < kGaiaNpars; i++) xmap[(size_t)ens * kGaiaNpars + i] = x[(size_t)w * kGaiaNpars + i];
    }
}

struct GaiaWarpState {
    double x[32][kGaiaNpars];  // by chain slot
    double logL[32];           // by chain slot
    double logP[32];           // log prior by chain slot (the reference recomputes it every step, :563)
    double beta[32];           // log of swap k's acceptance draw
    double dbeta[32];          // (heat_b - heat_{b+1}) / (heat_b heat_{b+1}) of the adjacent pair (b, b+1)
    int idx[32];               // rung -> slot
    int b[32];                 // swap k's lower rung
    int fill[32];              // slot whose state fills rung k's history ring
};

__global__ void __launch_bounds__(kGaiaWarpsPerBlock * 32)
k_gaia_pt_run(const GaiaPtConfig* __restrict__ cfgp, GaiaPtArrays a, unsigned iter0, unsigned n_iters, int thin,
              long rec_cap)
{
    __shared__ GaiaWarpState s_all[kGaiaWarpsPerBlock];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ens = blockIdx.x * kGaiaWarpsPerBlock + warp;
    const GaiaPtConfig& cfg = *cfgp;
    if (ens >= cfg.n_ens) return;
    GaiaWarpState& s = s_all[warp];
    const int T = cfg.n_temps, npast = cfg.npast;
    const bool live = lane < T;
    const int j = live ? lane : 0;  // lanes beyond the ladder idle through the walker part of the loop
    const uint32_t rid = (uint32_t)(ens * T + j);
    const double temp = cfg.temp[j], sqtemp = sqrt(temp);
    const double D = a.D[ens];
    double data[4], err[4];
    for (int i = 0; i < 4; i++) {
        data[i] = a.data[4 * ens + i];
        err[i] = a.err[4 * ens + i];
    }
    if (live) {
        for (int i = 0; i < kGaiaNpars; i++) s.x[lane][i] = a.x[((size_t)ens * T + lane) * kGaiaNpars + i];
        s.logL[lane] = a.logL[(size_t)ens * T + lane];
        s.idx[lane] = a.index[(size_t)ens * T + lane];
        s.logP[lane] = gaia_log_prior(s.x[lane], cfg);
        if (lane + 1 < T) {
            const double heat1 = cfg.temp[lane + 1], heat2 = cfg.temp[lane];
            s.dbeta[lane] = (heat2 - heat1) / (heat2 * heat1);
        }
    }
    double* hist = a.history + (size_t)rid * npast * kGaiaNpars;  // this rung's ring
    double logLmap = a.logLmap[ens];
    unsigned long long n_acc = 0, n_de = 0, n_de_acc = 0, n_acc0 = 0, n_swaps = 0;
    __syncwarp();

    for (unsigned it = iter0; it != iter0 + n_iters; it++) {
      if (live) {
        const int slot = s.idx[j];
        double x[kGaiaNpars], y[kGaiaNpars];
        for (int i = 0; i < kGaiaNpars; i++) x[i] = s.x[slot][i];
        const double logLx = s.logL[slot];

        // ---- propose (run_chain, GAIA_mcmc.c:525-543)
        PtRng g;
        g.init(cfg.seed, rid, it, 0u);
        const double alpha = g.next();
        const double jscale = pow(10., -6. + 6. * alpha);
        const bool de = (g.next() < 0.5) && ((long long)it > (long long)npast);
        int jt = 1;
        bool gauss_jump = !de;
        if (de) {
            const int ha = (int)(g.next() * npast);
            int hb_ = ha;
            while (hb_ == ha) hb_ = (int)(g.next() * npast);
            double dx[kGaiaNpars];
            for (int i = 0; i < kGaiaNpars; i++) dx[i] = hist[hb_ * kGaiaNpars + i] - hist[ha * kGaiaNpars + i];
            if (g.next() < 0.9) {
                double z[kGaiaNpars];
                gaia_normals(g, z);
                for (int i = 0; i < kGaiaNpars; i++) dx[i] *= z[i] * cfg.gamma;
            }
            double mag = 0.;
            for (int i = 0; i < kGaiaNpars; i++) {
                y[i] = x[i] + dx[i];
                mag += (x[i] - y[i]) * (x[i] - y[i]);
            }
            jt = 2;
            if (mag < 1e-6) {
                gauss_jump = true;
                jt = 1;
            }
        }
        if (gauss_jump) {
            double z[kGaiaNpars];
            gaia_normals(g, z);
            for (int i = 0; i < kGaiaNpars; i++) y[i] = x[i] + z[i] * cfg.sigma[i] * sqtemp * jscale;
        }
        // ---- boundary conditions (:546-559)
        for (int i = 0; i < kGaiaNpars; i++) {
            if ((cfg.mode_lo[i] == 1) && (y[i] < cfg.lo[i])) y[i] = 2.0 * cfg.lo[i] - y[i];
            if ((cfg.mode_hi[i] == 1) && (y[i] > cfg.hi[i])) y[i] = 2.0 * cfg.hi[i] - y[i];
            if ((cfg.mode_lo[i] == 2) && (y[i] < cfg.lo[i])) y[i] = cfg.hi[i] + (y[i] - cfg.lo[i]);
            if ((cfg.mode_hi[i] == 2) && (y[i] > cfg.hi[i])) y[i] = cfg.lo[i] + (y[i] - cfg.hi[i]);
        }
        // ---- priors, likelihood, Metropolis-Hastings (:561-583)
        const double logPx = s.logP[slot], logPy = gaia_log_prior(y, cfg);
        const double logLy = gaia_logL(y, D, data, err);
        const double H = exp((logLy - logLx) / temp) * pow(10., logPy - logPx);
        const double u_acc = pt_draw(cfg.seed, rid, it, 1u, 0u);
        const bool acc = u_acc <= H;
        {
            if (acc) {
                for (int i = 0; i < kGaiaNpars; i++) s.x[slot][i] = y[i];
                s.logL[slot] = logLy;
                s.logP[slot] = logPy;
                n_acc++;
                if (slot == 0) n_acc0++;
                if (slot == 0 && jt == 2) n_de_acc++;
            }
            if (slot == 0 && jt == 2) n_de++;
            if (it + 1u == iter0 + n_iters && a.last_y != nullptr) {  // the last proposals, for the tests
                for (int i = 0; i < kGaiaNpars; i++) a.last_y[(size_t)rid * kGaiaNpars + i] = y[i];
                a.last_logLy[rid] = logLy;
                a.last_logPy[rid] = logPy;
                a.last_jump[rid] = jt;
            }
            // ---- this iteration's swap draws, one per lane (consumed in order by lane 0 below)
            U4 c;
            c.x = 0x80000000u | (uint32_t)ens; c.y = it; c.z = 2u; c.w = (uint32_t)lane;
            const U4 r = philox4x32_10(c, (uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32));
            const double u0 = ((double)(((uint64_t)r.x << 21) | (r.y >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
            const double u1 = ((double)(((uint64_t)r.z << 21) | (r.w >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
            int b = (int)(u0 * (double)(T - 1));
            if (b > T - 2) b = T - 2;
            s.b[lane] = b;
            s.beta[lane] = log(u1);  // exp(x) >= beta  <=>  x >= log(beta)
        }
      }
        __syncwarp();
        // ---- swaps interleaved with the history fill (run_mcmc :741-748)
        if (lane == 0) {
            for (int k = 0; k < T; k++) {
                if (T > 1) {
                    const int b = s.b[k], a2 = b + 1;
                    const int olda = s.idx[a2], oldb = s.idx[b];
                    const double lalpha = (s.logL[oldb] - s.logL[olda]) * s.dbeta[b];
                    if (lalpha >= s.beta[k]) {
                        s.idx[a2] = oldb;
                        s.idx[b] = olda;
                        n_swaps++;
                    }
                }
                s.fill[k] = s.idx[k];
            }
        }
        __syncwarp();
        if (live) {
            const int fs = s.fill[lane];
            double* h = hist + (size_t)(it % (unsigned)npast) * kGaiaNpars;
            for (int i = 0; i < kGaiaNpars; i++) h[i] = s.x[fs][i];
        }
        // ---- MAP of the cold rung (:751-757) and the thinned log (:760-764, log_data :595-636)
        const int c0 = s.idx[0];
        const double l0 = s.logL[c0];
        if (l0 > logLmap) {
            logLmap = l0;
            if (lane < kGaiaNpars) a.xmap[(size_t)ens * kGaiaNpars + lane] = s.x[c0][lane];
        }
        if (thin > 0 && it % (unsigned)thin == 0u) {
            const long rec = (long)(it / (unsigned)thin) - (long)((iter0 + (unsigned)thin - 1u) / (unsigned)thin);
            if (rec >= 0 && rec < rec_cap) {
                if (a.rec_chain != nullptr) {
                    double* rc = a.rec_chain + ((size_t)ens * rec_cap + rec) * (kGaiaNpars + 1);
                    if (lane == 0) rc[0] = l0;
                    if (lane < kGaiaNpars) rc[1 + lane] = s.x[c0][lane];
                }
                if (a.rec_logL != nullptr && live) a.rec_logL[((size_t)ens * rec_cap + rec) * T + lane] = s.logL[s.idx[lane]];
            }
        }
        __syncwarp();
    }

    if (live) {
        for (int i = 0; i < kGaiaNpars; i++) a.x[((size_t)ens * T + lane) * kGaiaNpars + i] = s.x[lane][i];
        a.logL[(size_t)ens * T + lane] = s.logL[lane];
        a.index[(size_t)ens * T + lane] = s.idx[lane];
    }
    // counters per ensemble (same meaning as hb_pt_get_counters)
    const unsigned full = 0xffffffffu;
    for (int o = 16; o > 0; o >>= 1) {
        n_acc += __shfl_xor_sync(full, n_acc, o);
        n_de += __shfl_xor_sync(full, n_de, o);
        n_de_acc += __shfl_xor_sync(full, n_de_acc, o);
        n_acc0 += __shfl_xor_sync(full, n_acc0, o);
    }
    if (lane == 0) {
        unsigned long long* cnt = a.counters + (size_t)ens * 8;
        cnt[0] += n_acc0;
        cnt[1] += n_de;
        cnt[2] += n_de_acc;
        cnt[3] += n_acc;
        cnt[4] += (unsigned long long)n_iters * (unsigned long long)T;
        cnt[5] += n_swaps;
        cnt[6] += (unsigned long long)n_iters * (unsigned long long)T;
        cnt[7] += n_iters;
        a.logLmap[ens] = logLmap;
    }
}

cudaError_t launch_gaia_pt_init(const GaiaPtConfig* cfg, double* x, int W, cudaStream_t s)
{
    if (W > 0) k_gaia_pt_init<<<(W + 127) / 128, 128, 0, s>>>(cfg, x, W);
    return cudaGetLastError();
}

cudaError_t launch_gaia_pt_eval(const GaiaPtConfig* cfg, const GaiaPtArrays& a, int W, cudaStream_t s)
{
    if (W > 0) k_gaia_pt_eval<<<(W + 127) / 128, 128, 0, s>>>(cfg, a.x, a.D, a.data, a.err, a.index, a.logL, a.xmap, a.logLmap, W);
    return cudaGetLastError();
}

cudaError_t launch_gaia_pt_run(const GaiaPtConfig* cfg, const GaiaPtArrays& a, int n_ens, unsigned iter0, unsigned n_iters,
                               int thin, long rec_cap, cudaStream_t s)
{
    if (n_ens > 0 && n_iters > 0)
        k_gaia_pt_run<<<(n_ens + kGaiaWarpsPerBlock - 1) / kGaiaWarpsPerBlock, kGaiaWarpsPerBlock * 32, 0, s>>>(cfg, a, iter0,
                                                                                                                n_iters, thin, rec_cap);
    return cudaGetLastError();
}

}  // namespace hb
