/* oracle/gsl_stub -- TEST INFRASTRUCTURE ONLY.
 * GSL is not installed in this image and GAIA_mcmc.c (reference, compiled where it lies) uses it
 * for random numbers only (GAIA_mcmc.c:14-16,700-703,866-887).  This is the minimal API surface
 * that file touches, backed by a small generator of our own (gsl_stub.c); the stream is NOT
 * ranlxs1, so the stub serves (a) statistical comparisons against the unmodified sampler and
 * (b) deterministic tests that FEED the draws (gsl_stub_feed) so that every deterministic
 * function of GAIA_mcmc.c can be pinned bit for bit. */
#ifndef HB_GSL_STUB_RNG_H
#define HB_GSL_STUB_RNG_H
typedef struct gsl_rng_type_s { const char *name; } gsl_rng_type;
typedef struct gsl_rng_s { unsigned long long s; double spare; int have; } gsl_rng;
extern const gsl_rng_type *gsl_rng_ranlxs1;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(gsl_rng *r, unsigned long seed);
unsigned long gsl_rng_get(gsl_rng *r);
unsigned long gsl_rng_max(const gsl_rng *r);
double gsl_rng_uniform(gsl_rng *r);
void gsl_rng_free(gsl_rng *r);
/* test hook: while a feed is pending, gsl_rng_uniform / gsl_ran_gaussian pop from it */
void gsl_stub_feed(const double *uniforms, int n_uniforms, const double *normals, int n_normals);
int gsl_stub_feed_left(int which);
#endif
