/*
 * hb_mcmc.c -- the parallel-tempering driver of the reference (mcmc_wrapper2.c: main) on top of
 * the B200 library.  Same command line and the same files:
 *
 *     ./hb_mcmc NITER TIC_ID log10_period run_id            (mcmc_wrapper2.c:70-73)
 *
 *   reads   <prefix>/lightcurves/folded_lightcurves/<TIC>_new.txt   "Npts" then "t\tflux\terr" rows
 *           <prefix>/magnitudes/<TIC>.txt  (optional)               dist; G e; B-V e; V-G e; G-T e
 *           <prefix>/pars/par.<suffix>.dat (only when HB_USE_RAND_PARS=0, as USE_RAND_PARS does)
 *   writes  <prefix>/chains/chain.<suffix>.dat      every 100 its: "iter/10 logL p0..p20" (%.12g)
 *           <prefix>/logL/logL.<suffix>.dat         every 100 its: logL of every rung
 *           <prefix>/lightcurves/mcmc_lightcurves/<suffix>.out   MAP model: "t data model" (%12.5e)
 *           <prefix>/subpars/subpar.<suffix>.dat, <prefix>/pars/par.<suffix>.dat, <prefix>/log/log.<suffix>.dat
 *   with <suffix> = TIC[_gmag][_color]_B200_<run> (the reference appends _OMP for its OpenMP build).
 *
 * What runs where: file I/O, argument parsing and logging are host C (this file); proposals,
 * priors, the batched likelihood, accept/reject, swaps and the MAP update are device kernels
 * behind hb_pt_* (include/hb_b200.h).  The reference's per-rung debug dumps
 * (/scratch/.../debug/temp_<j>_log.txt, mcmc_wrapper2.c:360-374) are not produced.
 *
 * Environment: HB_DATA_PREFIX (default /scratch/ssolanski/HB_MCMC/data, the reference's hard-coded
 * prefix, mcmc_wrapper2.c:110), HB_DEVICE, HB_NTEMPS (50 = NCHAINS), HB_NENS (1), HB_SEED,
 * HB_USE_GMAG (1), HB_USE_COLOR_INFO (0), HB_USE_RAND_PARS (1), HB_QUIRKS (1).
 *
 * Several GPUs from this one process: HB_DEVICES=0,1,...,7 -- one context and one host thread per device.
 *   HB_NENS >= devices: whole ladders per device (hb_pt_create_sharded), nothing is exchanged;
 *   HB_NENS <  devices (the reference's ONE ladder): every device holds the ladder and evaluates its shard of the
 *   rungs' likelihoods; the per-step logL vector is all-gathered over NCCL from inside the captured step
 *   (hb_comm_create_all; HB_EXCHANGE=peer uses host-driven peer copies instead).
 * The chains and files are those of the one-device run, bit for bit (random streams are keyed on global ids).
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "hb_b200.h"

#define NPARS HB_NPARS
#define BIG_NUM 1.e15

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

static void die(hb_ctx *ctx, const char *what)
{
    fprintf(stderr, "hb_mcmc: %s: %s\n", what, ctx ? hb_last_error(ctx) : hb_global_error());
    exit(2);
}

#define CK(call) do { if ((call) != HB_OK) die(ctx, #call); } while (0)

#define MAX_DEV 64
/* the schedule of hb_pt_step calls every device follows (so that the collectives inside the steps pair up):
 * run up to and including the next iteration whose index is a multiple of 100 (a log point) */
static long next_chunk(long done, long NITER)
{
    const long next_log = ((done + 99) / 100) * 100;
    long n = next_log + 1 - done;
    if (done + n > NITER) n = NITER - done;
    return n;
}

struct worker {
    hb_ctx *ctx;
    hb_pt *pt;
    long NITER;
    int peer_mode, rank, ndev;
    hb_pt **all;                 /* peer mode: every device's sampler */
    pthread_barrier_t *bar;
};

/* one iteration with the host-driven exchange: all devices propose + evaluate, device 0 copies, all accept */
static void peer_steps(struct worker *w, long n)
{
    hb_ctx *ctx = w->ctx;
    for (long k = 0; k < n; k++) {
        CK(hb_pt_step_begin(w->pt));
        pthread_barrier_wait(w->bar);
        if (w->rank == 0 && hb_pt_exchange_local(w->all, w->ndev) != HB_OK) die(ctx, "hb_pt_exchange_local");
        pthread_barrier_wait(w->bar);
        CK(hb_pt_step_end(w->pt));
    }
}

static void *worker_main(void *arg)
{
    struct worker *w = arg;
    hb_ctx *ctx = w->ctx;
    long done = 0;
    while (done < w->NITER) {
        const long n = next_chunk(done, w->NITER);
        if (w->peer_mode) peer_steps(w, n);
        else CK(hb_pt_step(w->pt, n));
        done += n;
    }
    CK(hb_sync(ctx));
    return NULL;
}

int main(int argc, char *argv[])
{
    if (argc < 5) {
        fprintf(stderr, "usage: %s NITER TIC_ID log10_period run_id\n", argv[0]);
        return 1;
    }
    const long NITER = atol(argv[1]);
    const char *RUN_ID = argv[2];
    const double log_LC_PERIOD = atof(argv[3]);
    const int run = atoi(argv[4]);
    const double LC_PERIOD = pow(10., log_LC_PERIOD);

    const char *prefix = getenv("HB_DATA_PREFIX") ? getenv("HB_DATA_PREFIX") : "/scratch/ssolanski/HB_MCMC/data";
    const int use_gmag = env_int("HB_USE_GMAG", 1), use_color = env_int("HB_USE_COLOR_INFO", 0);
    const int n_temps = env_int("HB_NTEMPS", 50);
    int n_ens = env_int("HB_NENS", 1);
    const int n_ens_total = n_ens;
    const int use_rand_pars = env_int("HB_USE_RAND_PARS", 1), quirks = env_int("HB_QUIRKS", 1);
    const unsigned long long seed = (unsigned long long)env_int("HB_SEED", 0) + (unsigned long long)run;

    char suffix[256], parname[512], subparname[512], chainname[512], logLname[512], logname[512], outname[512],
        mag_name[512], dfname[512];
    snprintf(suffix, sizeof suffix, "%s%s%s_B200_%d", RUN_ID, use_gmag ? "_gmag" : "", use_color ? "_color" : "", run);
    snprintf(subparname, sizeof subparname, "%s/subpars/subpar.%s.dat", prefix, suffix);
    snprintf(parname, sizeof parname, "%s/pars/par.%s.dat", prefix, suffix);
    snprintf(chainname, sizeof chainname, "%s/chains/chain.%s.dat", prefix, suffix);
    snprintf(logLname, sizeof logLname, "%s/logL/logL.%s.dat", prefix, suffix);
    snprintf(logname, sizeof logname, "%s/log/log.%s.dat", prefix, suffix);
    snprintf(outname, sizeof outname, "%s/lightcurves/mcmc_lightcurves/%s.out", prefix, suffix);
    snprintf(mag_name, sizeof mag_name, "%s/magnitudes/%s.txt", prefix, RUN_ID);
    snprintf(dfname, sizeof dfname, "%s/lightcurves/folded_lightcurves/%s_new.txt", prefix, RUN_ID);
    printf("Parfile: %s\nSubparfile: %s\nChainfile: %s\nlogLfile: %s\noutfile: %s\nlogfile: %s\n", parname, subparname,
           chainname, logLname, outname, logname);

    /* folded light curve (mcmc_wrapper2.c:257-298) */
    printf("Opening folded lc data file %s \n", dfname);
    FILE *f = fopen(dfname, "r");
    if (!f) {
        printf("Lightcurve datafile not found; terminating program \n");
        return 0; /* the reference exits 0 here (mcmc_wrapper2.c:279-283) */
    }
    long Nt = 0;
    if (fscanf(f, "%ld\n", &Nt) != 1 || Nt <= 0) {
        fprintf(stderr, "hb_mcmc: bad point count in %s\n", dfname);
        return 2;
    }
    double *t_data = malloc(sizeof(double) * Nt), *a_data = malloc(sizeof(double) * Nt), *e_data = malloc(sizeof(double) * Nt),
           *a_model = malloc(sizeof(double) * Nt);
    for (long i = 0; i < Nt; i++)
        if (fscanf(f, "%lf\t%lf\t%lf\n", &t_data[i], &a_data[i], &e_data[i]) != 3) {
            fprintf(stderr, "hb_mcmc: %s: row %ld unreadable\n", dfname, i);
            return 2;
        }
    fclose(f);

    /* magnitudes (mcmc_wrapper2.c:302-328) */
    double mag_data[5] = {1000., 1., 1., 1., 1.}, mag_err[4] = {BIG_NUM, BIG_NUM, BIG_NUM, BIG_NUM};
    if ((use_color || use_gmag) && access(mag_name, R_OK) == 0) {
        printf("Using color / GMAG information \n");
        f = fopen(mag_name, "r");
        int ok = fscanf(f, "%lf\n", &mag_data[0]) == 1;
        for (int i = 0; ok && i < 4; i++) ok = fscanf(f, "%lf\t%lf\n", &mag_data[i + 1], &mag_err[i]) == 2;
        fclose(f);
        if (!ok) {
            fprintf(stderr, "hb_mcmc: %s unreadable\n", mag_name);
            return 2;
        }
    } else {
        printf("Magnitude file not found/used; assigning infinite error to mag data \n");
    }

    /* devices */
    int devs[MAX_DEV], ndev = 0;
    {
        const char *dl = getenv("HB_DEVICES");
        if (dl && *dl) {
            char *copy = strdup(dl), *save = NULL;
            for (char *tok = strtok_r(copy, ",", &save); tok && ndev < MAX_DEV; tok = strtok_r(NULL, ",", &save)) devs[ndev++] = atoi(tok);
            free(copy);
        }
        if (ndev == 0) devs[ndev++] = env_int("HB_DEVICE", 0);
    }
    const int rung_split = n_ens < ndev; /* fewer ladders than devices: split the likelihood evaluation of every ladder */
    const char *xch = getenv("HB_EXCHANGE");
    const int peer_mode = rung_split && xch && strcmp(xch, "peer") == 0;
    hb_ctx *ctxs[MAX_DEV];
    hb_pt *pts[MAX_DEV];
    hb_comm *comms[MAX_DEV];
    int ens_first[MAX_DEV], ens_count[MAX_DEV];
    for (int d = 0; d < ndev; d++) {
        hb_ctx *ctx = NULL;
        if (hb_create(&ctx, devs[d]) != HB_OK) die(NULL, "hb_create");
        ctxs[d] = ctx;
        comms[d] = NULL;
        CK(hb_set_mags(ctx, mag_data, mag_err, use_gmag, use_color));
        CK(hb_set_data(ctx, t_data, a_data, e_data, Nt));
        if (rung_split) {
            ens_first[d] = 0;
            ens_count[d] = n_ens;
        } else { /* contiguous blocks of whole ladders */
            const int base = n_ens / ndev, rem = n_ens % ndev;
            ens_count[d] = base + (d < rem ? 1 : 0);
            ens_first[d] = d * base + (d < rem ? d : rem);
        }
        CK(hb_pt_create_sharded(ctx, &pts[d], n_temps, ens_count[d], ens_first[d], log_LC_PERIOD, seed, 1.4, 500, quirks));
        if (rung_split) CK(hb_pt_set_eval_shard(pts[d], d, ndev));
    }
    if (rung_split && !peer_mode) {
        if (hb_comm_create_all(comms, devs, ndev) != HB_OK) {
            fprintf(stderr, "hb_mcmc: NCCL communicator: %s\n", hb_comm_last_error());
            return 2;
        }
        for (int d = 0; d < ndev; d++) {
            hb_ctx *ctx = ctxs[d];
            CK(hb_pt_set_comm(pts[d], comms[d]));
        }
    }
    hb_ctx *ctx = ctxs[0]; /* device 0 holds ladder 0 -- the ladder the files follow -- in either split */
    hb_pt *pt = pts[0];
    n_ens = ens_count[0];
    const int W = n_temps * n_ens;
    if (ndev > 1)
        printf("%d devices: %s\n", ndev, rung_split ? (peer_mode ? "rungs split, peer-copy exchange" : "rungs split, NCCL all-gather of logL per step")
                                                      : "whole ladders per device, no exchange");

    /* initial state (mcmc_wrapper2.c:203-252) */
    if (!use_rand_pars && access(parname, R_OK) == 0) {
        printf("Reading contents from parameter file \n");
        double p[NPARS], *x = malloc(sizeof(double) * W * NPARS);
        f = fopen(parname, "r");
        for (int i = 0; i < NPARS; i++) {
            if (fscanf(f, "%lf", &p[i]) != 1) p[i] = 0.;
            if (i == 2) p[i] = log_LC_PERIOD;
            if (i == 6) p[i] = fmod(p[i], LC_PERIOD);
        }
        fclose(f);
        for (int w = 0; w < W; w++) memcpy(x + (size_t)w * NPARS, p, sizeof p);
        for (int d = 0; d < ndev; d++) { /* (device 0 holds at least as many ladders as any other) */
            hb_ctx *ctx = ctxs[d];
            CK(hb_pt_set_state(pts[d], x));
        }
        free(x);
    } else {
        printf("Parameter file not found, assigning random pars \n");
        for (int d = 0; d < ndev; d++) {
            hb_ctx *ctx = ctxs[d];
            CK(hb_pt_init_random(pts[d]));
        }
    }
    double *cold_x = malloc(sizeof(double) * n_ens * NPARS), *cold_L = malloc(sizeof(double) * n_ens);
    double *rung_L = malloc(sizeof(double) * W), *xmap = malloc(sizeof(double) * n_ens * NPARS),
           *Lmap = malloc(sizeof(double) * n_ens);
    unsigned long long *cnt = malloc(sizeof(unsigned long long) * n_ens * 8), last_acc = 0, last_de = 0, last_det = 0,
                       last_it = 0;
    CK(hb_pt_get_cold(pt, cold_x, cold_L));
    printf("initial chi2 and likelihood %lf \t %lf\n", -2 * cold_L[0], cold_L[0]);

    printf("Creating chain and log files %s and %s \n", chainname, logLname);
    FILE *chain_file = fopen(chainname, "w"), *logL_file = fopen(logLname, "w"), *logfile = fopen(logname, "w");
    if (!chain_file || !logL_file || !logfile) {
        fprintf(stderr, "hb_mcmc: cannot create output files under %s (chains/ logL/ log/ must exist)\n", prefix);
        return 2;
    }
    fprintf(logfile, "hb_mcmc (B200): NITER %ld TIC %s log10P %.10g run %d rungs %d ensembles %d points %ld quirks %d\n", NITER,
            RUN_ID, log_LC_PERIOD, run, n_temps, n_ens, Nt, quirks);

    printf("Begining main mcmc loop \n");
    const clock_t c0 = clock();
    struct timespec ts0, ts1;
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    pthread_t threads[MAX_DEV];
    pthread_barrier_t bar;
    struct worker wk[MAX_DEV];
    pthread_barrier_init(&bar, NULL, (unsigned)ndev);
    for (int d = 0; d < ndev; d++) {
        wk[d].ctx = ctxs[d]; wk[d].pt = pts[d]; wk[d].NITER = NITER; wk[d].peer_mode = peer_mode; wk[d].rank = d;
        wk[d].ndev = ndev; wk[d].all = pts; wk[d].bar = &bar;
        if (d > 0 && pthread_create(&threads[d], NULL, worker_main, &wk[d]) != 0) {
            fprintf(stderr, "hb_mcmc: cannot start the host thread of device %d\n", devs[d]);
            return 2;
        }
    }
    long done = 0; /* iterations completed */
    while (done < NITER) {
        /* run up to and including the next iteration whose index is a multiple of 100 */
        const long n = next_chunk(done, NITER);
        if (peer_mode) peer_steps(&wk[0], n);
        else CK(hb_pt_step(pt, n));
        done += n;
        const long iter = done - 1;
        if (iter % 100 != 0) break; /* ran out of iterations before the next log point */

        CK(hb_pt_get_cold(pt, cold_x, cold_L));
        CK(hb_pt_get_logL_by_rung(pt, rung_L));
        CK(hb_pt_get_map(pt, xmap, Lmap));
        if (iter % 1000 == 0) { /* progress (mcmc_wrapper2.c:575-589) */
            CK(hb_pt_get_counters(pt, cnt));
            const double dacc = (double)(cnt[0] - last_acc), dit = (double)(cnt[7] - last_it);
            const double dde = (double)(cnt[2] - last_de), ddet = (double)(cnt[1] - last_det);
            printf("%ld/%ld logL=%.10g acc=%.3g DEacc=%.3g\n", iter, NITER, cold_L[0], dit > 0 ? dacc / dit : 0.,
                   ddet > 0 ? dde / ddet : 0.);
            last_acc = cnt[0]; last_it = cnt[7]; last_de = cnt[2]; last_det = cnt[1];
        }
        /* chain + logL lines (mcmc_wrapper2.c:593-620); ensemble 0 is the reference's single ladder */
        fprintf(chain_file, "%ld %.12g ", iter / 10, cold_L[0]);
        for (int i = 0; i < NPARS; i++) fprintf(chain_file, "%.12g ", cold_x[i]);
        fprintf(chain_file, "\n");
        fprintf(logL_file, "%ld ", iter / 10);
        for (int i = 0; i < n_temps; i++) fprintf(logL_file, "%.12g ", rung_L[i]);
        fprintf(logL_file, "\n");
        /* MAP model light curve and sub-parameters (mcmc_wrapper2.c:631-648) */
        CK(hb_calc_light_curve(ctx, t_data, Nt, xmap, a_model));
        f = fopen(outname, "w");
        if (f) {
            fprintf(f, "%ld\n", Nt);
            for (long i = 0; i < Nt; i++) fprintf(f, "%12.5e %12.5e %12.5e\n", t_data[i], a_data[i], a_model[i]);
            fclose(f);
        }
        f = fopen(subparname, "w");
        if (f) {
            for (int z = 0; z < NPARS; z++) fprintf(f, "%12.5e ", cold_x[z]);
            fprintf(f, "\n");
            fclose(f);
        }
    }
    CK(hb_sync(ctx));
    for (int d = 1; d < ndev; d++) pthread_join(threads[d], NULL);
    clock_gettime(CLOCK_MONOTONIC, &ts1);
    const double wall = (ts1.tv_sec - ts0.tv_sec) + 1e-9 * (ts1.tv_nsec - ts0.tv_nsec);

    /* final outputs (mcmc_wrapper2.c:655-681) */
    CK(hb_pt_get_cold(pt, cold_x, cold_L));
    CK(hb_pt_get_map(pt, xmap, Lmap));
    CK(hb_calc_light_curve(ctx, t_data, Nt, xmap, a_model));
    f = fopen(outname, "w");
    if (f) {
        fprintf(f, "%ld\n", Nt);
        for (long i = 0; i < Nt; i++) fprintf(f, "%12.5e %12.5e %12.5e\n", t_data[i], a_data[i], a_model[i]);
        fclose(f);
    }
    f = fopen(parname, "w");
    if (f) {
        for (int z = 0; z < NPARS; z++) fprintf(f, "%12.5e ", cold_x[z]);
        fprintf(f, "\n");
        fclose(f);
    }
    CK(hb_pt_get_counters(pt, cnt));
    fprintf(logfile, "iterations %ld wall_s %.3f steps_per_s %.2f cpu_s %.3f accepted %llu proposed %llu swaps %llu/%llu MAP logL %.12g\n",
            done, wall, wall > 0 ? done / wall : 0., (double)(clock() - c0) / CLOCKS_PER_SEC, cnt[3], cnt[4], cnt[5], cnt[6], Lmap[0]);
    printf("done: %ld iterations in %.3f s (%.1f PT steps/s, %d rungs x %d ensembles, %ld points), MAP logL %.10g\n", done, wall,
           wall > 0 ? done / wall : 0., n_temps, n_ens, Nt, Lmap[0]);
    if (ndev > 1 && !rung_split) { /* every ladder's MAP logL, by global ensemble id */
        fprintf(logfile, "MAP logL by ensemble:");
        double *L = malloc(sizeof(double) * n_ens_total);
        for (int d = 0; d < ndev; d++) {
            hb_ctx *ctx = ctxs[d];
            CK(hb_pt_get_map(pts[d], NULL, L + ens_first[d]));
        }
        for (int e = 0; e < n_ens_total; e++) fprintf(logfile, " %.12g", L[e]);
        fprintf(logfile, "\n");
        free(L);
    }
    fprintf(logfile, "devices %d split %s ensembles %d\n", ndev, ndev == 1 ? "none" : (rung_split ? "rungs" : "ensembles"), n_ens_total);
    fclose(logfile); fclose(chain_file); fclose(logL_file);
    for (int d = 0; d < ndev; d++) {
        hb_pt_destroy(pts[d]);
        if (comms[d]) hb_comm_destroy(comms[d]);
        hb_destroy(ctxs[d]);
    }
    free(t_data); free(a_data); free(e_data); free(a_model); free(cold_x); free(cold_L); free(rung_L); free(xmap);
    free(Lmap); free(cnt);
    return 0;
}
