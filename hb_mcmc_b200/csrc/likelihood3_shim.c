/*
 * likelihood3_shim.c -- libhb_likelihood3.so: the entry points of the reference's likelihood3.h
 * (likelihood3.h:68-89, plus the four helpers likelihood3.pxd:10-13 reaches into) with the SAME
 * names, argument lists and side effects, computed on the B200 through the C ABI of hb_b200.h.
 *
 * Link the unmodified reference driver / binding against this library INSTEAD of compiling
 * likelihood3.c:
 *     gcc -O3 -std=c99 -fopenmp mcmc_wrapper2.c -L<dir> -lhb_likelihood3 -lhb_b200 -lm
 *
 * Every function that computes runs on the device.  There is no CPU fallback: if the context
 * cannot be created (no B200, no driver) the first call prints the reason and abort()s, because
 * the reference API has no error channel (SURVEY.md 8b "Error convention: none").
 * Only set_limits / initialize_proposals (constant tables, likelihood3.c:986-1211) and the
 * generic array helpers partition / quickSort (likelihood3.c:48-83, not on the hot path: the
 * device model never sorts) are plain host C.
 *
 * Threading: the reference calls loglikelihood from up to 25 OpenMP threads on shared arrays
 * (mcmc_wrapper2.c:383,488-489).  Concurrent calls are COMBINED into one batched device call (see
 * loglikelihood below); a call that repeats a recent one bit for bit is answered from a memo; everything else
 * funnels into one context whose C ABI serialises it.
 */
#define _GNU_SOURCE /* clock_gettime, sched_yield, syscall under -std=c99 */
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <time.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <unistd.h>
#include <sys/syscall.h>
#include <linux/futex.h>

#include "hb_b200.h"

/* the three structs of likelihood3.h:39-66, redeclared (binary layout is the contract) */
struct bounds { double lo; double hi; };
struct gauss_bounds { int flag; };
typedef struct bounds bounds;
typedef struct gauss_bounds gauss_bounds;

#define NPARS HB_NPARS
#define PI 3.14159265358979323846

static hb_ctx *g_ctx = NULL;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static int g_use_gmag = 1, g_use_color = 0; /* likelihood3.h:11-12 */

/* cached copy of the last data set handed to loglikelihood() */
static double *g_t = NULL, *g_f = NULL, *g_e = NULL;
static long g_n = -1;
static long g_gen = 0;                                        /* bumped whenever the cached data set is replaced */
static pthread_rwlock_t g_data_rw = PTHREAD_RWLOCK_INITIALIZER; /* g_t/g_f/g_e/g_n/g_gen: readers = memo look-ups */

static void die(const char *what, const char *why)
{
    fprintf(stderr, "libhb_likelihood3: %s failed: %s\n(no CPU fallback: a B200 and libhb_b200.so are required)\n", what, why);
    abort();
}

/* HB_USE_GMAG / HB_USE_COLOR_INFO stand in for the compile-time macros of likelihood3.h:11-12.  They are read
 * once, without creating a device context: the reference driver calls set_limits / initialize_proposals
 * (mcmc_wrapper2.c:195-198) before its first loglikelihood (:342), and the sigma table depends on them. */
static pthread_once_t g_env_once = PTHREAD_ONCE_INIT;
static void read_env_flags(void)
{
    const char *env;
    if ((env = getenv("HB_USE_GMAG"))) g_use_gmag = atoi(env);
    if ((env = getenv("HB_USE_COLOR_INFO"))) g_use_color = atoi(env);
}

static hb_ctx *ctx(void)
{
    pthread_once(&g_env_once, read_env_flags);
    pthread_mutex_lock(&g_mu);
    if (!g_ctx) {
        const char *dev = getenv("HB_DEVICE");
        if (hb_create(&g_ctx, dev ? atoi(dev) : 0) != HB_OK) die("hb_create", hb_global_error());
    }
    pthread_mutex_unlock(&g_mu);
    return g_ctx;
}

#define CK(call)                                                   \
    do {                                                           \
        if ((call) != HB_OK) die(#call, hb_last_error(g_ctx));     \
    } while (0)

/* runtime switch for the compile-time macros USE_GMAG / USE_COLOR_INFO (likelihood3.h:11-12) */
void hb_shim_set_flags(int use_gmag, int use_color)
{
    pthread_once(&g_env_once, read_env_flags); /* an explicit call wins over the environment */
    pthread_rwlock_wrlock(&g_data_rw);
    g_use_gmag = use_gmag;
    g_use_color = use_color;
    g_gen++; /* values remembered under the old flags are not answers any more */
    pthread_rwlock_unlock(&g_data_rw);
}

void hb_shim_shutdown(void)
{
    pthread_mutex_lock(&g_mu);
    if (g_ctx) hb_destroy(g_ctx);
    g_ctx = NULL;
    pthread_rwlock_wrlock(&g_data_rw);
    free(g_t); free(g_f); free(g_e);
    g_t = g_f = g_e = NULL;
    g_n = -1;
    g_gen++;
    pthread_rwlock_unlock(&g_data_rw);
    pthread_mutex_unlock(&g_mu);
}

/* ---- likelihood3.h:68-70 ------------------------------------------------------------------ */
double partition(double arr[], int low, int high)
{
    double pivot = arr[high];
    int i = low - 1;
    for (int j = low; j < high; j++)
        if (arr[j] < pivot) {
            i++;
            double t = arr[i]; arr[i] = arr[j]; arr[j] = t;
        }
    double t = arr[i + 1]; arr[i + 1] = arr[high]; arr[high] = t;
    return (i + 1);
}

void quickSort(double arr[], int low, int high)
{
    while (low < high) { /* recurse on the smaller side: bounded stack */
        int p = (int)partition(arr, low, high);
        if (p - low < high - p) { quickSort(arr, low, p - 1); low = p + 1; }
        else { quickSort(arr, p + 1, high); high = p - 1; }
    }
}

void remove_median(double *arr, long begin, long end)
{
    if (end > begin) CK(hb_remove_median(ctx(), arr + begin, end - begin));
}

/* ---- likelihood3.h:71-80 ------------------------------------------------------------------ */
void traj(double *times, double *traj_pars, double *d_arr, double *Z1_arr, double *Z2_arr, double *rr_arr,
          double *ff_arr, int Nt)
{
    CK(hb_traj(ctx(), times, Nt, traj_pars, d_arr, Z1_arr, Z2_arr, rr_arr, ff_arr));
}

static double scalar(int op, const double *a, int n)
{
    double out = 0.;
    CK(hb_scalar(ctx(), op, a, n, &out));
    return out;
}

double _getT(double logM) { return scalar(0, &logM, 1); }
double _getR(double logM) { return scalar(1, &logM, 1); }
double envelope_Temp(double logM) { return scalar(2, &logM, 1); }
double envelope_Radius(double logM) { return scalar(3, &logM, 1); }
double get_alpha_beam(double logT) { return scalar(4, &logT, 1); }

double eclipse_area(double R1, double R2, double d)
{
    double a[3] = {R1, R2, d};
    return scalar(5, a, 3);
}

double beaming(double P, double M1, double M2, double e, double inc, double omega0, double nu, double alpha_beam)
{
    double a[8] = {P, M1, M2, e, inc, omega0, nu, alpha_beam};
    return scalar(6, a, 8);
}

double ellipsoidal(double P, double M1, double M2, double e, double inc, double omega0, double nu, double R1,
                   double a_, double mu, double tau)
{
    double a[11] = {P, M1, M2, e, inc, omega0, nu, R1, a_, mu, tau};
    return scalar(7, a, 11);
}

double reflection(double P, double M1, double M2, double e, double inc, double omega0, double nu, double R2,
                  double alpha_ref1)
{
    double a[9] = {P, M1, M2, e, inc, omega0, nu, R2, alpha_ref1};
    return scalar(8, a, 9);
}

/* ---- likelihood3.h:81-87 ------------------------------------------------------------------ */
void calc_mags(double params[], double D, double *Gmg, double *BminusV, double *VminusG, double *GminusT)
{
    double o[9];
    CK(hb_chain_info_batch(ctx(), params, 1, D, o));
    *Gmg = o[4]; *BminusV = o[5]; *VminusG = o[6]; *GminusT = o[7];
}

void calc_light_curve(double *times, long Nt, double *pars, double *template_)
{
    CK(hb_calc_light_curve(ctx(), times, Nt, pars, template_));
}

void calc_radii_and_Teffs(double params[], double *R1, double *R2, double *Teff1, double *Teff2)
{
    double o[9];
    CK(hb_chain_info_batch(ctx(), params, 1, 1000., o));
    *R1 = o[0]; *R2 = o[1]; *Teff1 = o[2]; *Teff2 = o[3];
}

int RocheOverflow(double *pars)
{
    double o[9];
    CK(hb_chain_info_batch(ctx(), pars, 1, 1000., o));
    return (int)o[8];
}

/*
 * loglikelihood() with combining.  The reference's rung loop calls it from many OpenMP threads at once
 * (mcmc_wrapper2.c:383,488-489), one chain per call; one device launch per call would be all latency.
 * Concurrent callers therefore queue their request; the first one to find no leader becomes the leader,
 * waits a few tens of microseconds for the other threads of the team to arrive (adaptively: until the
 * queue reaches the size of the previous batch, or stops growing), evaluates the whole queue with ONE
 * hb_loglikelihood_batch call and hands every caller its value.  A single-threaded caller pays one short
 * wait at most (the expected batch size decays to 1).  Requests that differ in data arrays or magnitudes
 * are evaluated group by group.
 */
#define SHIM_QMAX 256
/* patience of the collecting leader: a woken OpenMP team takes some tens of microseconds to come back with its
 * next call, and one device call costs ~50 us whatever its size -- waiting for the team pays */
#define SHIM_IDLE_US 50.0
#define SHIM_MAX_US 300.0
#define SHIM_DECAY_BATCHES 16
/* states of shim_req.done, the word a waiting caller sleeps on */
enum { REQ_WAITING = 0, REQ_DONE = 1, REQ_LEAD = 2 };
static void futex_wait(int *addr, int val) { syscall(SYS_futex, addr, FUTEX_WAIT_PRIVATE, val, NULL, NULL, 0); }
static void futex_wake(int *addr) { syscall(SYS_futex, addr, FUTEX_WAKE_PRIVATE, 1, NULL, NULL, 0); }
typedef struct {
    const double *time, *flux, *noise, *params, *mag_data, *magerr;
    long N;
    double out;
    int done;
    long gen; /* data-set generation the value belongs to (memo) */
} shim_req;

static pthread_mutex_t q_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t q_cv = PTHREAD_COND_INITIALIZER;
static shim_req *q_items[SHIM_QMAX];
static int q_len = 0, q_leader = 0, q_expect = 1, q_decay = 0;
static double idle_us = SHIM_IDLE_US;
/* HB_SHIM_STATS=1: batches, requests and the time spent collecting / evaluating are printed at exit */
static long st_batches = 0, st_reqs = 0, st_memo_hits = 0;
static double st_collect_us = 0., st_eval_us = 0., st_first_us = 0., st_last_us = 0.; /* end of the first / last batch */
static int st_on = -1;
static void st_print(void)
{
    if (st_batches > 0)
        fprintf(stderr, "libhb_likelihood3: %ld loglikelihood calls in %ld batches (%.1f per batch), %.1f us collecting and %.1f us evaluating per batch; %ld more calls answered from the memo\n",
                st_reqs, st_batches, (double)st_reqs / st_batches, st_collect_us / st_batches, st_eval_us / st_batches, st_memo_hits);
    if (st_batches > 1)
        fprintf(stderr, "libhb_likelihood3: span %.6f s from the end of the first batch to the end of the last (start-up excluded)\n",
                (st_last_us - st_first_us) * 1e-6);
}

static double now_us(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

static int same_group(const shim_req *a, const shim_req *b)
{
    return a->time == b->time && a->flux == b->flux && a->noise == b->noise && a->N == b->N &&
           memcmp(a->mag_data, b->mag_data, 5 * sizeof(double)) == 0 && memcmp(a->magerr, b->magerr, 4 * sizeof(double)) == 0;
}

/* evaluate items[0..n) (all of one group) in one device call; g_mu serialises the context's data set */
static void eval_group(hb_ctx *c, shim_req **items, int n)
{
    static double *pbuf = NULL, *obuf = NULL;
    static int cap = 0;
    pthread_mutex_lock(&g_mu);
    if (n > cap) {
        cap = n + 64;
        pbuf = (double *)realloc(pbuf, (size_t)cap * NPARS * sizeof(double));
        obuf = (double *)realloc(obuf, (size_t)cap * sizeof(double));
    }
    const shim_req *r0 = items[0];
    const long N = r0->N;
    size_t bytes = (size_t)(N > 0 ? N : 0) * sizeof(double);
    if (N != g_n || memcmp(r0->time, g_t, bytes) || memcmp(r0->flux, g_f, bytes) || memcmp(r0->noise, g_e, bytes)) {
        pthread_rwlock_wrlock(&g_data_rw);
        g_t = (double *)realloc(g_t, bytes + 8);
        g_f = (double *)realloc(g_f, bytes + 8);
        g_e = (double *)realloc(g_e, bytes + 8);
        memcpy(g_t, r0->time, bytes); memcpy(g_f, r0->flux, bytes); memcpy(g_e, r0->noise, bytes);
        g_n = N;
        g_gen++;
        pthread_rwlock_unlock(&g_data_rw);
        CK(hb_set_data(c, r0->time, r0->flux, r0->noise, N));
    }
    const long gen = g_gen; /* taken BEFORE the flags are read: a concurrent hb_shim_set_flags voids this batch's memo entries */
    CK(hb_set_mags(c, r0->mag_data, r0->magerr, g_use_gmag, g_use_color));
    for (int i = 0; i < n; i++) memcpy(pbuf + (size_t)i * NPARS, items[i]->params, NPARS * sizeof(double));
    CK(hb_loglikelihood_batch(c, pbuf, n, obuf));
    for (int i = 0; i < n; i++) {
        items[i]->out = obuf[i];
        items[i]->gen = gen;
    }
    pthread_mutex_unlock(&g_mu);
}

/*
 * Memo of recent values.  The reference's rung loop evaluates the CURRENT state of every rung again at every
 * step (mcmc_wrapper2.c:488) although its value is the previous step's logLx or logLy: half of the driver's
 * calls repeat a (data, params, magnitudes) triple seen one step earlier.  loglikelihood is a pure function of
 * those inputs, so such a call is answered from a small direct-mapped table -- after checking EVERY input bit
 * for bit (the 21 parameters, the 9 magnitude numbers, and the caller's three data arrays against the copy the
 * device data set was made from), never on pointers alone.  HB_SHIM_MEMO=0 switches it off.
 */
#define MEMO_SETS 2048 /* x 4 ways: a rung's current state must survive the ~100 stores of one driver step */
#define MEMO_WAYS 4
typedef struct {
    double params[NPARS], mags[9], out;
    long gen, N;
    int valid;
    unsigned long stamp; /* last store or hit: the least recently used way of a set is replaced */
} memo_ent;
static memo_ent memo[MEMO_SETS][MEMO_WAYS];
static unsigned long memo_clock = 0;
static pthread_mutex_t memo_mu = PTHREAD_MUTEX_INITIALIZER;
static int memo_on = -1;

/* management call: 0 = every call is evaluated on the device, 1 = memo on (the default; env HB_SHIM_MEMO) */
void hb_shim_set_memo(int on) { memo_on = on != 0; }
long hb_shim_memo_hits(void) { return st_memo_hits; }

static unsigned memo_set(const double *params)
{
    unsigned long long h = 1469598103934665603ULL, w;
    for (int i = 0; i < NPARS; i++) {
        memcpy(&w, &params[i], sizeof w);
        h = (h ^ w) * 1099511628211ULL;
        h ^= h >> 29;
    }
    return (unsigned)(h % MEMO_SETS);
}

static int memo_same_key(const memo_ent *e, const shim_req *r)
{
    return e->valid && e->N == r->N && !memcmp(e->params, r->params, sizeof e->params) &&
           !memcmp(e->mags, r->mag_data, 5 * sizeof(double)) && !memcmp(e->mags + 5, r->magerr, 4 * sizeof(double));
}

static int memo_lookup(const shim_req *r, double *out)
{
    memo_ent e;
    int found = 0;
    const unsigned set = memo_set(r->params);
    pthread_mutex_lock(&memo_mu);
    for (int w = 0; w < MEMO_WAYS && !found; w++)
        if (memo_same_key(&memo[set][w], r)) {
            memo[set][w].stamp = ++memo_clock; /* a state that keeps being asked for stays */
            e = memo[set][w];
            found = 1;
        }
    pthread_mutex_unlock(&memo_mu);
    if (!found) return 0;
    int hit = 0;
    const size_t bytes = (size_t)(r->N > 0 ? r->N : 0) * sizeof(double);
    pthread_rwlock_rdlock(&g_data_rw);
    if (e.gen == g_gen && r->N == g_n && !memcmp(r->time, g_t, bytes) && !memcmp(r->flux, g_f, bytes) &&
        !memcmp(r->noise, g_e, bytes))
        hit = 1;
    pthread_rwlock_unlock(&g_data_rw);
    if (hit) *out = e.out;
    return hit;
}

static void memo_store(const shim_req *r)
{
    memo_ent e;
    memcpy(e.params, r->params, sizeof e.params);
    memcpy(e.mags, r->mag_data, 5 * sizeof(double));
    memcpy(e.mags + 5, r->magerr, 4 * sizeof(double));
    e.out = r->out;
    e.gen = r->gen;
    e.N = r->N;
    e.valid = 1;
    const unsigned set = memo_set(r->params);
    pthread_mutex_lock(&memo_mu);
    int w = -1;
    for (int k = 0; k < MEMO_WAYS && w < 0; k++)
        if (memo_same_key(&memo[set][k], r) || !memo[set][k].valid) w = k; /* same key (older data set) or free */
    if (w < 0) {
        w = 0;
        for (int k = 1; k < MEMO_WAYS; k++)
            if (memo[set][k].stamp < memo[set][w].stamp) w = k;
    }
    e.stamp = ++memo_clock;
    memo[set][w] = e;
    pthread_mutex_unlock(&memo_mu);
}

double loglikelihood(double time[], double lightcurve[], double noise[], long N, double params[], double mag_data[],
                     double magerr[])
{
    hb_ctx *c = ctx();
    /* side effect of likelihood3.c:824-827: the caller's noise[] is clamped in place (quirk Q2) */
    for (long i = 0; i < N; i++)
        if (noise[i] < 1.e-5) noise[i] = 1.e-5;
    shim_req me = {time, lightcurve, noise, params, mag_data, magerr, N, 0., 0, 0};
    if (memo_on < 0) {
        const char *env = getenv("HB_SHIM_MEMO");
        memo_on = env ? atoi(env) != 0 : 1;
    }
    if (memo_on && memo_lookup(&me, &me.out)) {
        __sync_fetch_and_add(&st_memo_hits, 1);
        return me.out;
    }

    pthread_mutex_lock(&q_mu);
    while (q_len >= SHIM_QMAX) pthread_cond_wait(&q_cv, &q_mu);
    q_items[q_len++] = &me;
    int lead = !q_leader; /* nobody collecting or evaluating: lead ONE batch (my own request is queued, so it is in it) */
    if (lead) q_leader = 1;
    pthread_mutex_unlock(&q_mu);
    if (!lead) {
        /* wait on my OWN word (futex), not on a shared condition variable: 25 sleepers woken by one broadcast come
         * back one after the other through q_mu, and that queue was most of the next batch's collecting time */
        int st;
        while ((st = __atomic_load_n(&me.done, __ATOMIC_ACQUIRE)) == REQ_WAITING) futex_wait(&me.done, REQ_WAITING);
        if (st == REQ_DONE) goto finished; /* the leader does not touch `me` after publishing REQ_DONE */
        /* REQ_LEAD: the previous leader handed the leadership to me (q_leader is still 1) */
        __atomic_store_n(&me.done, REQ_WAITING, __ATOMIC_RELAXED);
    }
    {
        pthread_mutex_lock(&q_mu);
        if (st_on < 0) {
            const char *env = getenv("HB_SHIM_IDLE_US"); /* patience of the collecting leader (default SHIM_IDLE_US) */
            if (env && atof(env) > 0.) idle_us = atof(env);
            st_on = getenv("HB_SHIM_STATS") != NULL;
            if (st_on) atexit(st_print);
        }
        const double st_t0 = st_on ? now_us() : 0.;
        /* collect: wait until the queue holds as many requests as the previous batch did, or stops growing */
        if (q_expect > 1 || q_len > 1) {
            const double t0 = now_us();
            int last = q_len;
            double t_last = t0;
            while (q_len < q_expect) {
                pthread_mutex_unlock(&q_mu);
                sched_yield();
                pthread_mutex_lock(&q_mu);
                const double t = now_us();
                if (q_len != last) { last = q_len; t_last = t; }
                if (t - t_last > idle_us || t - t0 > SHIM_MAX_US) break;
            }
        }
        shim_req *batch[SHIM_QMAX];
        const int n = q_len;
        memcpy(batch, q_items, (size_t)n * sizeof(batch[0]));
        q_len = 0;
        /* follows the team size up at once, down slowly: one step per SHIM_DECAY_BATCHES smaller batches.  (A
         * decrement at every smaller batch settles on splitting a 25-thread team into 15 + 10 once the memo
         * answers half of the calls and the arrivals spread out: the late third comes a few us after the
         * leader has left.) */
        if (n >= q_expect) { q_expect = n; q_decay = 0; }
        else if (++q_decay >= SHIM_DECAY_BATCHES) { q_expect--; q_decay = 0; }
        pthread_cond_broadcast(&q_cv);                          /* room in the queue again (callers blocked on a full queue) */
        pthread_mutex_unlock(&q_mu);
        const double st_t1 = st_on ? now_us() : 0.;
        /* evaluate group by group (normally one group) */
        int used[SHIM_QMAX] = {0};
        for (int i = 0; i < n; i++) {
            if (used[i]) continue;
            shim_req *grp[SHIM_QMAX];
            int m = 0;
            for (int j = i; j < n; j++)
                if (!used[j] && same_group(batch[i], batch[j])) { grp[m++] = batch[j]; used[j] = 1; }
            eval_group(c, grp, m);
        }
        pthread_mutex_lock(&q_mu);
        if (st_on) {
            st_batches++;
            st_reqs += n;
            st_collect_us += st_t1 - st_t0;
            st_last_us = now_us();
            st_eval_us += st_last_us - st_t1;
            if (st_batches == 1) st_first_us = st_last_us;
        }
        /* whoever queued up meanwhile leads the next batch: the leadership passes to the first of them directly */
        shim_req *next = q_len > 0 ? q_items[0] : NULL;
        if (!next) q_leader = 0;
        pthread_mutex_unlock(&q_mu);
        for (int i = 0; i < n; i++)
            if (batch[i] != &me) { /* results are in: publish, wake, and never touch batch[i] again */
                int *w = &batch[i]->done;
                __atomic_store_n(w, REQ_DONE, __ATOMIC_RELEASE);
                futex_wake(w);
            }
        if (next) {
            __atomic_store_n(&next->done, REQ_LEAD, __ATOMIC_RELEASE);
            futex_wake(&next->done);
        }
    }
finished:
    if (memo_on) memo_store(&me);
    return me.out;
}

/* ---- likelihood3.h:88-89: constant tables (likelihood3.c:986-1211) ------------------------ */
void set_limits(bounds limited[], bounds limits[], gauss_bounds gauss_pars[], double LC_PERIOD)
{
    static const double lo[NPARS] = {-1.5, -1.5, -2.0, 0.0, 0.0, -PI, 0.0, -5., -5., 0.12, 0.3, 0.12, 0.3, 0.5, 0.5,
                                     -0.3, -0.3, -5., -5., 0., 0.99};
    static const double hi[NPARS] = {2.0, 2.0, 3.0, 1.0, PI, PI, 0.0, 5., 5., 0.20, 0.38, 0.20, 0.38, 1.5, 1.5,
                                     0.3, 0.3, 5., 5., 1., 1.01};
    for (int i = 0; i < NPARS; i++) {
        limits[i].lo = lo[i];
        limits[i].hi = hi[i];
        limited[i].lo = 1;  /* 1 = reflecting */
        limited[i].hi = 1;
        gauss_pars[i].flag = (i >= 7 && i <= 18) ? 1 : 0;
    }
    limited[3].hi = 0.99;  /* quirk Q4: mode and limit of e swapped in the reference; kept */
    limited[5].lo = 2;     /* 2 = periodic */
    limited[5].hi = 2;
    limits[6].hi = LC_PERIOD;
}

void initialize_proposals(double *sigma, double ***history)
{
    static const double base[NPARS] = {1e-2, 1e-2, 1e-8, 1e-2, 1e-3, 1e-3, 1e-3, 1e-1, 1e-1, 1e-2, 1e-2,
                                       1e-2, 1e-2, 1e-2, 1e-2, 1e-2, 1e-2, 1e-1, 1e-1, 1e-3, 1e-5};
    (void)history;
    pthread_once(&g_env_once, read_env_flags);
    memcpy(sigma, base, sizeof(base));
    if ((!g_use_color) || (!g_use_gmag)) { /* the enlarged set, likelihood3.c:1158-1179 */
        sigma[0] = sigma[1] = 1e-1;
        sigma[4] = sigma[5] = 1e-2;
        sigma[6] = 1e-3;
        for (int i = 9; i <= 18; i++) sigma[i] = 1e-1;
    }
}
