/*
 * hb_b200.h -- C ABI of libhb_b200.so: the B200 (sm_100a) replacement for the hot path of
 * sidruns30/HB_MCMC -- the light-curve model + chi^2 log-likelihood of likelihood3.c evaluated
 * for every chain / temperature rung of the parallel-tempering driver mcmc_wrapper2.c.
 *
 * Plain C, caller-owned buffers, `int` status codes (0 = HB_OK).  No torch / C++ types cross
 * this boundary.  File:line citations are into the reference tree (/root/reference/src).
 *
 * Parameter vector (21 doubles per chain, likelihood3.c:533-578):
 *   0 logM1  1 logM2  2 logP[d]  3 e  4 inc  5 omega0  6 T0[d]  7 rr1  8 rr2  9 mu1  10 tau1
 *   11 mu2  12 tau2  13 alpha_ref1  14 alpha_ref2  15 ln xbeam1  16 ln xbeam2  17 aTeff1
 *   18 aTeff2  19 blending  20 flux_tune
 *
 * There is no CPU fallback: every entry point that computes runs on the device and fails
 * with HB_ERR_CUDA when no sm_100 device / driver is usable.
 */
#ifndef HB_B200_H
#define HB_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define HB_NPARS 21 /* likelihood3.h:20 */

enum {
    HB_OK = 0,
    HB_ERR_ARG = 1,   /* bad argument (NULL, negative size, data not set) */
    HB_ERR_CUDA = 2,  /* CUDA runtime error; text in hb_last_error() */
    HB_ERR_STATE = 3  /* call order (e.g. likelihood before hb_set_data) */
};

typedef struct hb_ctx hb_ctx;

/* ---- context -------------------------------------------------------------------------- */
/* One context per device / host thread group.  Calls on one context are serialised by an
 * internal mutex, so the reference's 25 OpenMP threads (mcmc_wrapper2.c:78-83,383) may share it. */
int hb_create(hb_ctx** out, int device);
void hb_destroy(hb_ctx* ctx);
const char* hb_last_error(const hb_ctx* ctx); /* valid until the next call on ctx */
const char* hb_global_error(void);            /* error text when hb_create itself failed */
int hb_device_info(hb_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, long* global_mem_mb);
/* Use an existing CUDA stream (e.g. torch's current stream) for all work of this context;
 * NULL restores the context's own stream. */
int hb_set_stream(hb_ctx* ctx, void* cuda_stream);
int hb_sync(hb_ctx* ctx);

/* ---- data ------------------------------------------------------------------------------ */
/* Upload the observed light curve once (mcmc_wrapper2.c:257-298: t_data, a_data, e_data).
 * err[] is clamped to >= 1e-5 in the DEVICE copy (likelihood3.c:824-827, quirk Q2); the host
 * array is left untouched -- the shim below reproduces the in-place side effect. */
int hb_set_data(hb_ctx* ctx, const double* t, const double* flux, const double* err, long n);
/* mag_data[5] = {D, G, B-V, V-G, G-T}, magerr[4] (mcmc_wrapper2.c:302-328); use_gmag/use_color
 * are likelihood3.h:11-12 made runtime.  Defaults: {1000,1,1,1,1}, {1e15 x4}, 1, 0. */
int hb_set_mags(hb_ctx* ctx, const double* mag_data, const double* magerr, int use_gmag, int use_color);

/* ---- the hot path ---------------------------------------------------------------------- */
/* logL[c] = loglikelihood(t, flux, err, N, params[c], mag_data, magerr)   (likelihood3.c:809-873)
 * for c < n_chains; params row-major [n_chains][21].  Host buffers; returns after the result
 * is in logL.  Roche-overflow chains give exactly -5e14 (likelihood3.c:863-869).  When params and logL are both
 * page-locked (cudaHostAlloc / cudaHostRegister) the kernels read and write them in place over the bus -- no copy
 * (HB_ZERO_COPY=0 in the environment: cudaMemcpyAsync both ways); pageable ones are staged in chunks. */
int hb_loglikelihood_batch(hb_ctx* ctx, const double* params, long n_chains, double* logL);
/* Same with DEVICE buffers, asynchronous on the context's stream (no copies, no sync). */
int hb_loglikelihood_batch_dev(hb_ctx* ctx, const double* d_params, long n_chains, double* d_logL);
/* templates[c][i] = calc_light_curve(t, N, params[c])[i]   (likelihood3.c:530-686) on the
 * uploaded time grid.  Host buffers, [n_chains][N] row-major. */
int hb_light_curve_batch(hb_ctx* ctx, const double* params, long n_chains, double* templates);
/* calc_light_curve on an arbitrary time array (the pyHB / driver call: pyHB.pyx:66,
 * mcmc_wrapper2.c:632,659).  Does not disturb the uploaded data set. */
int hb_calc_light_curve(hb_ctx* ctx, const double* times, long nt, const double* pars, double* tmpl);

/* ---- per-chain helpers ----------------------------------------------------------------- */
/* out[c][9] = {R1, R2, Teff1, Teff2, G, B-V, V-G, G-T, RocheOverflow}: calc_radii_and_Teffs
 * (likelihood3.c:693-717), calc_mags at distance D (:725-795), RocheOverflow (:953-974). */
int hb_chain_info_batch(hb_ctx* ctx, const double* params, long n_chains, double D, double* out);
/* traj() of likelihood3.c:125-185; traj_pars[7] = {M1,M2 [g], P [s], e, inc, omega0, T0 [s]} */
int hb_traj(hb_ctx* ctx, const double* times, long nt, const double* traj_pars, double* d_arr, double* Z1_arr,
            double* Z2_arr, double* rr_arr, double* ff_arr);
/* Exact k-th smallest (0-based) of x[0..n): the order statistic remove_median() obtains by
 * sorting a copy (likelihood3.c:86-105).  NaN in x gives NaN. */
int hb_order_statistic(hb_ctx* ctx, const double* x, long n, long k, double* out);
/* remove_median(arr, 0, n) of likelihood3.c:86-105 in place: subtracts the reference's "median"
 * order statistic (even n: sorted[n/2], odd n: sorted[n/2+1], quirk Q3). */
int hb_remove_median(hb_ctx* ctx, double* arr, long n);
/* Scalar model functions evaluated on the device.  op / args:
 *   0 _getT(logM)  1 _getR(logM)  2 envelope_Temp(logM)  3 envelope_Radius(logM)
 *   4 get_alpha_beam(logT)  5 eclipse_area(R1,R2,d)  6 beaming(8 args)  7 ellipsoidal(11 args)
 *   8 reflection(9 args)          (likelihood3.c:194-209,224-389,396-507) */
int hb_scalar(hb_ctx* ctx, int op, const double* args, int nargs, double* out);
/* Gaia flavour (GAIA_mcmc.c:198-269): p6[n][6] = {logM1,logM2,rr1,rr2,aT1,aT2}; mags[n][4] and/or
 * logL[n] (either may be NULL); data[4], err[4] needed when logL != NULL. */
int hb_gaia_batch(hb_ctx* ctx, const double* p6, long n, double D, const double* data, const double* err,
                  double* mags, double* logL);

/* ---- device-resident parallel tempering ------------------------------------------------ */
/* The step/swap loop of mcmc_wrapper2.c:378-563 with all chain state on the device: n_ens
 * independent ladders of n_temps rungs (the reference: 1 x NCHAINS = 50, mcmc_wrapper2.h:11),
 * temperatures dtemp^j (1.4, :331-339), DE history of npast samples per rung (500, :12).
 * Philox4x32-10 keyed on (seed, rung, iteration) replaces ran2/gasdev2/rand().  quirks != 0 keeps
 * the reference's sampler bugs as compiled (Q5, Q6, Q8 of SURVEY Appendix B; see hb_pt.cuh).
 * Needs hb_set_data first; limits and proposal sigmas are those of set_limits /
 * initialize_proposals (likelihood3.c:986-1179) for the context's use_gmag / use_color. */
typedef struct hb_pt hb_pt;
int hb_pt_create(hb_ctx* ctx, hb_pt** out, int n_temps, int n_ens, double log_lc_period, unsigned long long seed,
                 double dtemp, int npast, int quirks);
void hb_pt_destroy(hb_pt* pt);
/* uniform draws in the prior box, period pinned, T0 mod P (mcmc_wrapper2.c:236-251) + first logL */
int hb_pt_init_random(hb_pt* pt);
/* x[n_ens*n_temps][21] by chain slot (slot = ens*n_temps + chain id) + first logL */
int hb_pt_set_state(hb_pt* pt, const double* x);
/* n_iters full iterations: every rung proposes, is evaluated (ONE likelihood per rung: the
 * current-state value is cached, the reference re-evaluates it, :488), accepts/rejects; then
 * n_temps swap proposals per ensemble (:554-563) and the MAP update (:565-572) */
int hb_pt_step(hb_pt* pt, long n_iters);
/* Light curves of at most 1024 points (the reference's real, folded ones have 163-763) with every walker resident at
 * once: hb_pt_step runs the whole n_iters loop in ONE launch (a CTA per walker, one grid-wide barrier per iteration)
 * instead of five stream-ordered kernels per iteration -- same chains bit for bit.  enable = 0 turns that off. */
int hb_pt_set_one_launch(hb_pt* pt, int enable);
long hb_pt_iteration(const hb_pt* pt);
int hb_pt_get_state(hb_pt* pt, double* x, double* logL, int* index);       /* any may be NULL */
int hb_pt_get_proposal(hb_pt* pt, double* y, double* logLy, double* logPy); /* last proposals, by slot */
int hb_pt_get_cold(hb_pt* pt, double* x_cold, double* logL_cold);          /* rung 0: [n_ens][21], [n_ens] */
int hb_pt_get_logL_by_rung(hb_pt* pt, double* out);                        /* [n_ens][n_temps] */
int hb_pt_get_map(hb_pt* pt, double* xmap, double* logLmap);               /* [n_ens][21], [n_ens] */
/* per ensemble 8 counters: acc of slot 0, DE trials / DE acc of slot 0 (the reference's acc, DEtrial,
 * DEacc), accepted / proposed over all rungs, swaps accepted / proposed, iterations */
int hb_pt_get_counters(hb_pt* pt, unsigned long long* out);
/* device pointer of logL[n_ens*n_temps] (for an NCCL all-gather by the caller) */
void* hb_pt_device_logL(hb_pt* pt);
/* cold-rung logL of every ensemble into a DEVICE buffer d_out[n_ens], asynchronous on the context's
 * stream: the send buffer of the per-step NCCL all-gather */
int hb_pt_cold_logL_dev(hb_pt* pt, double* d_out);

/* ---- the sampler on several GPUs ---------------------------------------------------------- */
/* Two ways to use more than one GPU; in both, every random number is keyed on GLOBAL rung / ensemble ids, so the
 * chains are those of the one-GPU run bit for bit, whatever the split.
 *  (1) whole ensembles per GPU (n_ens >= GPUs): hb_pt_create_sharded(ens_offset = global id of the first local
 *      ensemble).  Ladders never talk to each other: no exchange at all (mcmc_wrapper2.c:383 treats rungs alike).
 *  (2) fewer ensembles than GPUs -- the reference's own case is ONE ladder (mcmc_wrapper2.h:11): every rank holds
 *      the whole sampler state and proposes, accepts and swaps redundantly (counter-based RNG), but evaluates the
 *      likelihood -- all of the cost -- only for its contiguous shard of the walkers; the per-step log-likelihood
 *      vector is then all-gathered (8 bytes per walker: the only exchange) and the swaps (mcmc_wrapper2.c:554-563,
 *      ptmcmc :768-817) are decided identically everywhere.  hb_pt_set_eval_shard + either a communicator
 *      (hb_pt_set_comm: ncclAllGather on the context's stream, inside the captured step) or, within one process,
 *      hb_pt_step_begin / hb_pt_exchange_local / hb_pt_step_end (peer copies driven by the host). */
int hb_pt_create_sharded(hb_ctx* ctx, hb_pt** out, int n_temps, int n_ens, int ens_offset, double log_lc_period,
                         unsigned long long seed, double dtemp, int npast, int quirks);
int hb_pt_set_eval_shard(hb_pt* pt, int rank, int world);                        /* world <= 64 */
int hb_pt_get_eval_shard(const hb_pt* pt, long* first, long* count, long* chunk); /* walkers evaluated here */
typedef struct hb_comm hb_comm;
#define HB_COMM_ID_BYTES 128
/* NCCL is bound at run time (dlopen of libnccl.so.2, or $HB_NCCL_LIB).  One process per GPU: rank 0 makes an id,
 * hands the 128 bytes to the others by any means (torch.distributed, a file, MPI ...), every rank creates its
 * communicator on its device.  One process, several GPUs: hb_comm_create_all (ncclCommInitAll). */
int hb_comm_unique_id(unsigned char id[HB_COMM_ID_BYTES]);
int hb_comm_create(hb_comm** out, int device, const unsigned char id[HB_COMM_ID_BYTES], int rank, int world);
int hb_comm_create_all(hb_comm** out, const int* devices, int n);
void hb_comm_destroy(hb_comm* comm);
int hb_comm_rank(const hb_comm* comm);
int hb_comm_world(const hb_comm* comm);
int hb_comm_nccl_version(int* version);
const char* hb_comm_last_error(void);
/* in-place all-gather of doubles on a CUDA stream: rank r's count_per_rank values sit at d_buf + r * count_per_rank */
int hb_comm_allgather_f64(hb_comm* comm, double* d_buf, long count_per_rank, void* cuda_stream);
int hb_pt_set_comm(hb_pt* pt, hb_comm* comm);   /* not owned; NULL unbinds */
int hb_pt_step_begin(hb_pt* pt);                 /* propose + likelihood of this rank's shard (asynchronous) */
int hb_pt_exchange_local(hb_pt** pts, int n);    /* the n samplers of one process swap their shards' logL */
int hb_pt_step_end(hb_pt* pt);                   /* accept + swaps + MAP */

/* ---- Gaia-colour sampler ---------------------------------------------------------------- */
/* The stand-alone sampler of GAIA_mcmc.c:663-780 (run_mcmc) on the device: n_ens independent
 * ladders (one warp each, one lane per rung; n_temps <= 32; the reference: NCHAINS = 20 rungs with
 * ratio dtemp = 1.2, NPAST = 100, GAIA_mcmc.c:23-24,476) walk 6 parameters {logM1, logM2, rr1, rr2,
 * aT1, aT2} against the 4-point likelihood of hb_gaia_batch.  The whole run is ONE kernel launch;
 * Philox replaces GSL ranlxs1 / rand().  Limits and priors: set_limits (:346-390); proposal sigmas:
 * init_proposals (:449-458) -- {1e-2, 1e-2, 0, 0, 0, 0}, see hb_gaia_pt_set_sigma. */
typedef struct hb_gaia_pt hb_gaia_pt;
int hb_gaia_pt_create(hb_ctx* ctx, hb_gaia_pt** out, int n_temps, int n_ens, unsigned long long seed, double dtemp,
                      int npast);
void hb_gaia_pt_destroy(hb_gaia_pt* pt);
/* per ensemble: distance D[n_ens] (pc), data[n_ens][4] = {G, B-V, V-G, G-T} and err[n_ens][4]
 * (read_mag_data, GAIA_mcmc.c:314-343) */
int hb_gaia_pt_set_data(hb_gaia_pt* pt, const double* D, const double* data, const double* err);
/* the reference leaves sigma[2..5] unset (fresh heap memory, 0 in practice); override here */
int hb_gaia_pt_set_sigma(hb_gaia_pt* pt, const double* sigma6);
int hb_gaia_pt_init_random(hb_gaia_pt* pt);                 /* init_chain, :463-473, + first logL */
int hb_gaia_pt_set_state(hb_gaia_pt* pt, const double* x);  /* x[n_ens*n_temps][6] by chain slot */
/* number of log records n_iters further iterations produce: iterations it with it % thin == 0 (:760) */
long hb_gaia_pt_records(const hb_gaia_pt* pt, long n_iters, int thin);
/* n_iters iterations in one launch.  chain[n_ens][records][7] = {logL, x[6]} of the cold rung (the
 * chain file, log_data :599-605) and logL_by_rung[n_ens][records][n_temps] (the logL file, :617-622);
 * either may be NULL (thin is then ignored). */
int hb_gaia_pt_run(hb_gaia_pt* pt, long n_iters, int thin, double* chain, double* logL_by_rung);
long hb_gaia_pt_iteration(const hb_gaia_pt* pt);
int hb_gaia_pt_get_state(hb_gaia_pt* pt, double* x, double* logL, int* index);  /* any may be NULL */
/* proposals of the LAST iteration by rung: y[n_ens*n_temps][6], logLy, logPy, jump type (1 Gaussian, 2 DE) */
int hb_gaia_pt_get_proposal(hb_gaia_pt* pt, double* y, double* logLy, double* logPy, int* jump);
int hb_gaia_pt_get_history(hb_gaia_pt* pt, double* history);  /* [n_ens*n_temps][npast][6] by rung */
int hb_gaia_pt_get_map(hb_gaia_pt* pt, double* xmap, double* logLmap);  /* [n_ens][6], [n_ens] */
int hb_gaia_pt_get_counters(hb_gaia_pt* pt, unsigned long long* out);  /* as hb_pt_get_counters */

/* ---- measurement ----------------------------------------------------------------------- */
/* DFMA throughput of the device in TFLOP/s (2 flop per FMA), the FP64 roofline denominator. */
int hb_fp64_peak(hb_ctx* ctx, double seconds_target, double* tflops);
/* Tuning / test knob of the fused median: half-width (in binomial standard deviations, default 2.5) of
 * the rank bracket the 256-sample pre-sample puts around the reference's median rank (remove_median,
 * likelihood3.c:86-105).  A chain whose bracket misses is evaluated a second time with its template stored
 * and selected exactly, so the result never depends on this value -- only the cost does (0 forces every
 * chain down the miss path; the tests use that). */
int hb_set_bracket_sigma(hb_ctx* ctx, double sigma);
/* Batches smaller than the grid: k_chain_eval spreads every light curve over up to max_parts CTAs (a power of two
 * between 1 and 64; default 64; 1 = one CTA per chain always).  The chi^2 is summed per time segment and in segment
 * order whatever the spread is, so a chain's logL does not depend on this value or on the size of the batch it
 * arrives in -- only the latency of small batches does (likelihood3.c:147,649,822 loop serially over the samples). */
int hb_set_max_parts(hb_ctx* ctx, int max_parts);
/* Chains whose model was really evaluated by the likelihood kernel since the last reset (Roche-overflow and e >= 1
 * chains return early without touching the light curve, quirk Q13, and are NOT counted). */
int hb_evaluated_chains(hb_ctx* ctx, unsigned long long* count, int reset);
/* Second test knob.  The logL-only pass of k_chain_eval does not range-check its table sincos per sample: it
 * records the largest |E| any Newton iterate of the chain reached and, when that exceeds max_abs (default and
 * maximum 1024, the validity range of the table sincos), evaluates the chain again with the per-sample check
 * and the libm fallback.  Inside the prior box the Roche test keeps e below ~0.9985, where such iterates do not
 * occur; lowering max_abs forces the re-evaluation so that the tests can hold it to the normal path's bits. */
int hb_set_sincos_range(hb_ctx* ctx, double max_abs);
/* When enabled, every likelihood / light-curve call records CUDA events around its k_chain_eval
 * launch on the launching stream; hb_last_eval_kernel_ms waits for and returns that duration. */
int hb_time_kernels(hb_ctx* ctx, int enable);
int hb_last_eval_kernel_ms(hb_ctx* ctx, double* ms);
/* Number of kernel launches issued by this context since creation. */
long hb_launch_count(const hb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* HB_B200_H */
