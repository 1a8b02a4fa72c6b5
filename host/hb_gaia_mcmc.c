/*
 * hb_gaia_mcmc.c -- the stand-alone Gaia-colour sampler of the reference (GAIA_mcmc.c: main,
 * run_mcmc) on top of the B200 library.  Same command line and the same files:
 *
 *     ./hb_gaia_mcmc NITER TIC_ID NTHREADS                  (GAIA_mcmc.c:783-808)
 *
 *   reads   <prefix>/magnitudes/<TIC>.txt        distance; then "value<TAB>error" for G, B-V, V-G, G-T
 *   writes  <prefix>/chains/<TIC>_GAIA_run.txt   every 10 its: "logL p0..p5" of the cold rung (%.10g, tab separated)
 *           <prefix>/logL/<TIC>_GAIA_run.txt     every 10 its: logL of every rung
 *           <prefix>/subpars/<TIC>_GAIA_run.txt  the latest cold-rung parameters (rewritten in place)
 *           <prefix>/GAIA_runs/<TIC>_GAIA_run.txt the latest cold-rung model magnitudes
 *
 * The reference hard-codes <prefix> = "../data" relative to its working directory (:316,400-421);
 * HB_DATA_PREFIX overrides.  NTHREADS is accepted and ignored: the ladder is one warp on the GPU.
 * The reference seeds its generators with NITER (:676); so does this driver (HB_SEED overrides).
 * HB_NENS > 1 runs that many independent ladders side by side (other seeds of the same star); the
 * files then describe ensemble 0 and <TIC>_GAIA_run.ens<k>.txt the others' chains.
 *
 * What runs where: file I/O is host C (this file); the whole sampling loop is ONE kernel launch per
 * block of HB_GAIA_BLOCK iterations (default 100000) behind hb_gaia_pt_run (include/hb_b200.h).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hb_b200.h"

#define MAGPARS 6
#define SUBN 4
#define NCHAINS 20 /* GAIA_mcmc.c:23 */
#define NPAST 100  /* :24 */
#define THIN 10    /* :760 */

static long env_long(const char *name, long dflt)
{
    const char *v = getenv(name);
    return v ? atol(v) : dflt;
}

static void die(hb_ctx *ctx, const char *what)
{
    fprintf(stderr, "hb_gaia_mcmc: %s: %s\n", what, ctx ? hb_last_error(ctx) : hb_global_error());
    exit(2);
}

#define CK(call) do { if ((call) != HB_OK) die(ctx, #call); } while (0)

int main(int argc, char *argv[])
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s NITER TIC_ID [NTHREADS]\n", argv[0]);
        return 1;
    }
    const long NITER = atol(argv[1]);
    const char *TIC = argv[2];
    const char *prefix = getenv("HB_DATA_PREFIX") ? getenv("HB_DATA_PREFIX") : "../data";
    const int n_ens = (int)env_long("HB_NENS", 1), n_temps = (int)env_long("HB_NTEMPS", NCHAINS);
    const long block = env_long("HB_GAIA_BLOCK", 100000);
    const unsigned long long seed = (unsigned long long)env_long("HB_SEED", NITER);
    if (argc > 3) printf("number of threads = %d (ignored: one warp per ladder on the GPU)\n", atoi(argv[3]));

    /* read_mag_data, GAIA_mcmc.c:314-343 */
    char fname[600];
    snprintf(fname, sizeof fname, "%s/magnitudes/%s.txt", prefix, TIC);
    printf("Opening magnitude file %s \n", fname);
    FILE *mf = fopen(fname, "r");
    if (!mf) {
        printf("Could not open magnitude file %s \n", fname);
        return 1;
    }
    double distance, ydata[SUBN], yerr[SUBN];
    int nread = fscanf(mf, "%lf", &distance);
    for (int i = 0; i < SUBN; i++) nread += fscanf(mf, "%lf %lf", &ydata[i], &yerr[i]);
    fclose(mf);
    if (nread != 1 + 2 * SUBN) {
        fprintf(stderr, "hb_gaia_mcmc: %s: expected 9 numbers, read %d\n", fname, nread);
        return 1;
    }

    hb_ctx *ctx = NULL;
    if (hb_create(&ctx, (int)env_long("HB_DEVICE", 0)) != HB_OK) die(NULL, "hb_create");
    hb_gaia_pt *pt = NULL;
    CK(hb_gaia_pt_create(ctx, &pt, n_temps, n_ens, seed, 1.2, NPAST));
    double *D = malloc(sizeof(double) * n_ens), *dat = malloc(sizeof(double) * 4 * n_ens), *er = malloc(sizeof(double) * 4 * n_ens);
    for (int e = 0; e < n_ens; e++) {
        D[e] = distance;
        memcpy(dat + 4 * e, ydata, sizeof ydata);
        memcpy(er + 4 * e, yerr, sizeof yerr);
    }
    CK(hb_gaia_pt_set_data(pt, D, dat, er));
    CK(hb_gaia_pt_init_random(pt));

    /* create_log_files, :397-426 */
    char chain_fname[600], par_fname[600], logL_fname[600], out_fname[600];
    snprintf(chain_fname, sizeof chain_fname, "%s/chains/%s_GAIA_run.txt", prefix, TIC);
    snprintf(par_fname, sizeof par_fname, "%s/subpars/%s_GAIA_run.txt", prefix, TIC);
    snprintf(logL_fname, sizeof logL_fname, "%s/logL/%s_GAIA_run.txt", prefix, TIC);
    snprintf(out_fname, sizeof out_fname, "%s/GAIA_runs/%s_GAIA_run.txt", prefix, TIC);
    FILE *chain_file = fopen(chain_fname, "w"), *par_file = fopen(par_fname, "w"), *logL_file = fopen(logL_fname, "w"),
         *out_file = fopen(out_fname, "w");
    if (!chain_file || !par_file || !logL_file || !out_file) {
        fprintf(stderr, "hb_gaia_mcmc: cannot open the output files under %s\n", prefix);
        return 1;
    }
    FILE **ens_files = calloc((size_t)n_ens, sizeof(FILE *));
    for (int e = 1; e < n_ens; e++) {
        char nm[640];
        snprintf(nm, sizeof nm, "%s/chains/%s_GAIA_run.ens%d.txt", prefix, TIC, e);
        ens_files[e] = fopen(nm, "w");
    }
    {
        double *l0 = malloc(sizeof(double) * n_ens * n_temps);
        CK(hb_gaia_pt_get_state(pt, NULL, l0, NULL));
        printf("initial chi2 %g\n", -2 * l0[0]); /* init_chain, :487 */
        free(l0);
    }

    long done = 0;
    const long max_rec = block / THIN + 1;
    double *chain = malloc(sizeof(double) * (size_t)n_ens * max_rec * (MAGPARS + 1));
    double *rung = malloc(sizeof(double) * (size_t)n_ens * max_rec * n_temps);
    double last[MAGPARS + 1];
    int have_last = 0;
    while (done < NITER) {
        const long n = NITER - done < block ? NITER - done : block;
        const long nrec = hb_gaia_pt_records(pt, n, THIN);
        CK(hb_gaia_pt_run(pt, n, THIN, chain, rung));
        for (long r = 0; r < nrec; r++) { /* log_data, :595-636 */
            const double *c = chain + r * (MAGPARS + 1);
            for (int i = 0; i <= MAGPARS; i++) fprintf(chain_file, "%.10g\t", c[i]);
            fprintf(chain_file, "\n");
            for (int j = 0; j < n_temps; j++) fprintf(logL_file, "%.10g\t", rung[r * n_temps + j]);
            fprintf(logL_file, "\n");
            const long it = (done + THIN - 1) / THIN * THIN + r * THIN;
            if (it % 10000 == 0) printf("Iter: %ld \t Best likelihood %f \n", it, c[0]); /* :767-771 */
            memcpy(last, c, sizeof last);
            have_last = 1;
        }
        for (int e = 1; e < n_ens; e++) {
            if (!ens_files[e]) continue;
            for (long r = 0; r < nrec; r++) {
                const double *c = chain + ((size_t)e * nrec + r) * (MAGPARS + 1);
                for (int i = 0; i <= MAGPARS; i++) fprintf(ens_files[e], "%.10g\t", c[i]);
                fprintf(ens_files[e], "\n");
            }
        }
        done += n;
    }
    if (have_last) { /* the files the reference rewinds and rewrites at every log step hold the last record */
        for (int i = 0; i < MAGPARS; i++) fprintf(par_file, "%.10g\t", last[1 + i]);
        fprintf(par_file, "\n");
        double mags[SUBN];
        CK(hb_gaia_batch(ctx, last + 1, 1, distance, ydata, yerr, mags, NULL));
        for (int i = 0; i < SUBN; i++) fprintf(out_file, "%.10g\n", mags[i]);
    }
    unsigned long long *cnt = malloc(sizeof(unsigned long long) * 8 * n_ens);
    CK(hb_gaia_pt_get_counters(pt, cnt));
    printf("acceptance (all rungs) %.4f  swaps %.4f  cold-rung acc %.4f\n", cnt[4] ? (double)cnt[3] / cnt[4] : 0.,
           cnt[6] ? (double)cnt[5] / cnt[6] : 0., cnt[7] ? (double)cnt[0] / cnt[7] : 0.);
    fclose(chain_file); fclose(par_file); fclose(logL_file); fclose(out_file);
    for (int e = 1; e < n_ens; e++) if (ens_files[e]) fclose(ens_files[e]);
    hb_gaia_pt_destroy(pt);
    hb_destroy(ctx);
    return 0;
}
