// hb_pt.cuh -- building blocks of the parallel-tempering step of mcmc_wrapper2.c as device code.
// The kernels (hb_pt.cu) give every walker a warp, one lane per parameter:
//   propose  mcmc_wrapper2.c:390-481  jump scale, Gaussian (:1062-1088) or differential-evolution
//            (:1091-1140) proposal, reflect / periodic bounds (:440-467), mass ordering, pinned
//            period, T0 mod P, Gaussian priors get_logP (:703-765)
//   accept   mcmc_wrapper2.c:492-546  Metropolis-Hastings with the rung temperature, history ring
//   swap     mcmc_wrapper2.c:554-563 + ptmcmc :768-817  NCHAINS random adjacent-pair swaps
//
// RNG: Philox4x32-10, counter = (stream id, iteration, stage, draw block), key = seed.  It replaces
// ran2/gasdev2 (:833-974, quirk Q7: the reference's per-rung streams are identical for every seed
// > 0 because the shuffle table is never initialised) and libc rand() in the swaps; streams are
// NOT bit-compatible with the reference by design -- parity for this layer is (a) deterministic
// agreement with the oracle's restatement fed the same Philox stream and (b) statistical.
//
// Positions on the reference's sampler bugs, selected by PtConfig::quirks:
//   quirks = 1 (default, "as compiled"): Q5 the mass-ordering block copies instead of swapping
//     (y[1] = y[0] when y[1] > y[0]); Q6 the DE proposal always uses history sample a = 0 and adds
//     epsilon = dx (gaussian(c,0,1e-4) - 0.5) with the uninitialised int c, which the reference
//     built with gcc -O3 evaluates with c == 0, i.e. epsilon = dx (3989.42 - 0.5) -- measured on
//     oracle/_ref/libref_mcmc.so; Q8 swaps ignore the prior (likelihood only).
//   quirks = 0 ("as intended"): masses are swapped, a is uniform on the history, epsilon = 0.
// Always: the current-state likelihood is cached (the reference re-evaluates it every step).
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef HB_HOST_EMUL
#include <cuda_runtime.h>
#define HB_HD __host__ __device__ __forceinline__
#else
#define HB_HD static inline
#endif

namespace hb {

constexpr int kPtNpars = 21;
constexpr int kPtMaxTemps = 128;

struct PtConfig {
    int n_temps, n_ens, npast, quirks;
    int ens_offset;        // global id of this sampler's first ensemble (ensembles sharded over GPUs): the Philox
    int pad_;              // streams are keyed on GLOBAL rung / ensemble ids, so a chain does not depend on the split
    unsigned long long seed;
    double dtemp;          // temperature ladder ratio, 1.4 (mcmc_wrapper2.c:331)
    double temp[kPtMaxTemps];  // temp[0] = 1, temp[i] = temp[i-1] * dtemp (mcmc_wrapper2.c:332-338)
    double log_lc_period;  // argv[3]
    double lc_period;      // 10^log_lc_period
    double gamma;          // 2.388 / sqrt(2 NPARS)  (mcmc_wrapper2.h:13)
    double lo[kPtNpars], hi[kPtNpars], mode_lo[kPtNpars], mode_hi[kPtNpars], sigma[kPtNpars];
    int gauss[kPtNpars];
};

// ---- Philox4x32-10 ----------------------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };

HB_HD void philox_round(U4& c, uint32_t k0, uint32_t k1)
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
    U4 r;
    r.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
    r.y = (uint32_t)p1;
    r.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
    r.w = (uint32_t)p0;
    c = r;
}

HB_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1)
{
    for (int i = 0; i < 10; i++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// Sequential uniform doubles in (0,1) from the stream (id, iter, stage): block n of the stream is
// philox(ctr = {id, iter, stage, n}) and yields two doubles.
struct PtRng {
    uint32_t id, iter, stage, n, k0, k1;
    double spare;
    int have;
    HB_HD void init(unsigned long long seed, uint32_t id_, uint32_t iter_, uint32_t stage_)
    {
        id = id_; iter = iter_; stage = stage_; n = 0; have = 0; spare = 0.;
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
    }
    HB_HD double next()
    {
        if (have) { have = 0; return spare; }
        U4 c; c.x = id; c.y = iter; c.z = stage; c.w = n++;
        const U4 r = philox4x32_10(c, k0, k1);
        const uint64_t a = ((uint64_t)r.x << 21) | (r.y >> 11);
        const uint64_t b = ((uint64_t)r.z << 21) | (r.w >> 11);
        spare = ((double)b + 0.5) * (1.0 / 9007199254740992.0);
        have = 1;
        return ((double)a + 0.5) * (1.0 / 9007199254740992.0);
    }
    // two standard normals (Box-Muller)
    HB_HD void normal2(double& z0, double& z1)
    {
        const double u1 = next(), u2 = next();
        const double r = sqrt(-2.0 * log(u1));
        const double a = 6.283185307179586 * u2;
        z0 = r * cos(a);
        z1 = r * sin(a);
    }
};

// Draw number d (0-based) of stream (id, iter, stage) without walking the stream: draws 2n and
// 2n+1 are the two halves of Philox block n.  PtRng::next() returns exactly this sequence, so a
// warp can evaluate all draws of one walker in parallel.
HB_HD double pt_draw(unsigned long long seed, uint32_t id, uint32_t iter, uint32_t stage, uint32_t d)
{
    U4 c; c.x = id; c.y = iter; c.z = stage; c.w = d >> 1;
    const U4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t v = (d & 1u) ? (((uint64_t)r.z << 21) | (r.w >> 11)) : (((uint64_t)r.x << 21) | (r.y >> 11));
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

// normal number n of a run of Box-Muller pairs whose first uniform is draw d0 (PtRng::normal2 order)
HB_HD double pt_normal(unsigned long long seed, uint32_t id, uint32_t iter, uint32_t stage, uint32_t d0, int n)
{
    const uint32_t d = d0 + 2u * (uint32_t)(n >> 1);
    const double u1 = pt_draw(seed, id, iter, stage, d), u2 = pt_draw(seed, id, iter, stage, d + 1u);
    const double r = sqrt(-2.0 * log(u1));
    const double a = 6.283185307179586 * u2;
    return (n & 1) ? r * sin(a) : r * cos(a);
}

// prior mean / sigma of parameter i (mcmc_wrapper2.c:710-757)
HB_HD void pt_prior_of(int i, double& mean, double& sig)
{
    mean = 0.; sig = 1e15;
    if (i == 7 || i == 8 || i == 17 || i == 18) sig = 1.;
    else if (i == 9 || i == 11) { mean = 0.16; sig = 0.04; }
    else if (i == 10 || i == 12) { mean = 0.34; sig = 0.04; }
    else if (i == 13 || i == 14) { mean = 1.; sig = 0.2; }
    else if (i == 15 || i == 16) sig = 0.1;
}

// mcmc_wrapper2.c:1175-1178 with SQRT_2PI of mcmc_wrapper2.h:10
HB_HD double pt_gaussian(double x, double mean, double sigma)
{
    const double z = (x - mean) / sigma;  // (z * z is the correctly rounded pow(z, 2.) of the reference's expression)
    return (1 / sigma / 2.5066282746) * exp(-(z * z) / 2.);
}

// Boundary handling of parameter i (mcmc_wrapper2.c:440-467): reflect at sides with mode 1, wrap
// when both sides have mode 2.  Hot rungs (sqrt(T) up to 4e4) and the "as compiled" DE jumps
// overshoot the box by thousands of widths; the reference bounces them back one reflection at a
// time.  Two reflections are a translation by 2 (hi - lo) and one wrap a translation by (hi - lo),
// so whole multiples are removed in one step and the loops only finish the last bounce.
HB_HD double pt_bound_one(double y, int i, const PtConfig& cfg)
{
    const double lo = cfg.lo[i], hi = cfg.hi[i];
    const bool rl = cfg.mode_lo[i] == 1, rh = cfg.mode_hi[i] == 1;
    const bool pl = cfg.mode_lo[i] == 2, ph = cfg.mode_hi[i] == 2;
    const double period = (rl && rh) ? 2.0 * (hi - lo) : ((pl && ph) ? (hi - lo) : 0.0);
    if (period > 0.0 && fabs(y) < 1e300) {
        if (y > hi + period) y -= period * floor((y - hi) / period);
        else if (y < lo - period) y += period * floor((lo - y) / period);
    }
    int guard = 0;
    while (((rl && (y < lo)) || (rh && (y > hi))) && guard < 200000) {
        y = (y < lo) ? 2.0 * lo - y : 2.0 * hi - y;
        if (!(fabs(y) < 1e300)) return NAN;  // +-inf would bounce forever; NaN rejects the proposal
        guard++;
    }
    if (guard >= 200000) return NAN;  // the reference would still be looping
    guard = 0;
    while (pl && (y < lo) && guard++ < 200000) y = hi + (y - lo);
    while (ph && (y > hi) && guard++ < 200000) y = lo + (y - hi);
    return y;
}

// Metropolis-Hastings decision of mcmc_wrapper2.c:492-505 (NaN H rejects)
HB_HD bool pt_accept(const PtConfig& cfg, uint32_t rung_id, uint32_t iter, double temp, double logLx, double logLy,
                     double logPx, double logPy)
{
    PtRng g;
    g.init(cfg.seed, rung_id, iter, 1u);
    const double alpha = g.next();
    const double H = exp((logLy - logLx) / temp + (logPy - logPx));
    return alpha <= H;
}

}  // namespace hb
