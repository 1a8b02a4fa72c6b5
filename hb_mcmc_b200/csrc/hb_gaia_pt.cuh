// hb_gaia_pt.cuh -- configuration and device arrays of the Gaia-colour sampler (hb_gaia_pt.cu).
#pragma once
#include <cuda_runtime.h>

namespace hb {

constexpr int kGaiaNpars = 6;          // MAGPARS, GAIA_mcmc.c:26
constexpr int kGaiaMaxTemps = 32;      // one lane per rung (the reference: NCHAINS = 20, :23)
constexpr int kGaiaWarpsPerBlock = 4;  // ensembles per CTA

struct GaiaPtConfig {
    int n_temps, n_ens, npast, pad;
    unsigned long long seed;
    double gamma;  // 2.388 / sqrt(2 MAGPARS), GAIA_mcmc.c:27
    double temp[kGaiaMaxTemps];  // temp[i] = 1.2^i (:476-484)
    double lo[kGaiaNpars], hi[kGaiaNpars], mode_lo[kGaiaNpars], mode_hi[kGaiaNpars], sigma[kGaiaNpars];
    int gauss[kGaiaNpars];
};

// device pointers; E = n_ens, T = n_temps
struct GaiaPtArrays {
    double *x, *logL;          // [E*T][6], [E*T] by chain slot
    int* index;                // [E][T] rung -> slot
    double* history;           // [E*T][npast][6] by rung
    double *xmap, *logLmap;    // [E][6], [E]
    unsigned long long* counters;  // [E][8]
    const double *D, *data, *err;  // [E], [E][4], [E][4]: distance (pc), {G, B-V, V-G, G-T} and their errors
    double *rec_chain, *rec_logL;  // thinned log: [E][rec_cap][7] = {logL, x} of the cold rung; [E][rec_cap][T]
    double *last_y, *last_logLy, *last_logPy;  // last proposals by rung (tests)
    int* last_jump;
};

cudaError_t launch_gaia_pt_init(const GaiaPtConfig* cfg, double* x, int W, cudaStream_t s);
cudaError_t launch_gaia_pt_eval(const GaiaPtConfig* cfg, const GaiaPtArrays& a, int W, cudaStream_t s);
cudaError_t launch_gaia_pt_run(const GaiaPtConfig* cfg, const GaiaPtArrays& a, int n_ens, unsigned iter0, unsigned n_iters,
                               int thin, long rec_cap, cudaStream_t s);

}  // namespace hb
