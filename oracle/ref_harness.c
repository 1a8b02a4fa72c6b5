/*
 * ref_harness.c -- thin batch/timing harness AROUND the unmodified reference (TEST
 * INFRASTRUCTURE ONLY).  Linked together with /root/reference/src/likelihood3.c into
 * oracle/_ref/libref_lik3.so by oracle/Makefile; nothing here restates the model.
 *
 * Why threads with explicit stacks: the reference keeps 9 x Nt doubles in stack VLAs
 * (likelihood3.c:89,632-643,812; quirk Q12) and segfaults at Nt = 200k on an 8 MB stack.
 * The harness runs every call on pthreads whose stacks are sized from Nt, which is what
 * `ulimit -s unlimited` + OMP_STACKSIZE would do for the reference's own OpenMP loop
 * (mcmc_wrapper2.c:383).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* the reference entry points (likelihood3.h:81-87) */
double loglikelihood(double time[], double lightcurve[], double noise[], long N, double params[],
                     double mag_data[], double magerr[]);
void calc_light_curve(double *times, long Nt, double *pars, double *template_);

#define REF_NPARS 21

typedef struct {
    double *time, *flux, *noise;
    long N;
    const double *params;
    long n;
    double *mag_data, *magerr;
    double *logL;
    long *next; /* shared work counter */
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *job = (batch_job *)arg;
    for (;;) {
        long c = __atomic_fetch_add(job->next, 1, __ATOMIC_RELAXED);
        if (c >= job->n) break;
        double p[REF_NPARS];
        memcpy(p, job->params + c * REF_NPARS, sizeof(p));
        job->logL[c] = loglikelihood(job->time, job->flux, job->noise, job->N, p, job->mag_data, job->magerr);
    }
    return NULL;
}

static size_t stack_bytes(long N) { return (size_t)(64u << 20) + (size_t)N * 8u * 12u; }

/* n chains x N points through the reference loglikelihood on `nthreads` host threads
 * (dynamic distribution, like schedule(dynamic)).  Returns 0 on success. */
int ref_loglikelihood_batch(double *time, double *flux, double *noise, long N, const double *params, long n,
                            double *mag_data, double *magerr, int nthreads, double *logL)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    long next = 0;
    batch_job job = {time, flux, noise, N, params, n, mag_data, magerr, logL, &next};
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, stack_bytes(N));
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    int started = 0;
    for (int i = 0; i < nthreads; i++) {
        if (pthread_create(&th[i], &attr, batch_worker, &job) != 0) break;
        started++;
    }
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    pthread_attr_destroy(&attr);
    free(th);
    return started > 0 ? 0 : -1;
}

typedef struct {
    double *times;
    long Nt;
    double *pars;
    double *out;
} lc_job;

static void *lc_worker(void *arg)
{
    lc_job *j = (lc_job *)arg;
    calc_light_curve(j->times, j->Nt, j->pars, j->out);
    return NULL;
}

/* reference calc_light_curve on a thread whose stack fits the VLAs */
int ref_calc_light_curve(double *times, long Nt, double *pars, double *out)
{
    lc_job j = {times, Nt, pars, out};
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, stack_bytes(Nt));
    pthread_t th;
    if (pthread_create(&th, &attr, lc_worker, &j) != 0) return -1;
    pthread_join(th, NULL);
    pthread_attr_destroy(&attr);
    return 0;
}
