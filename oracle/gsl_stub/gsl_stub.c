/* oracle/gsl_stub/gsl_stub.c -- TEST INFRASTRUCTURE ONLY; see gsl/gsl_rng.h. */
#include <math.h>
#include <stdlib.h>
#include <gsl/gsl_randist.h>

static const gsl_rng_type k_type = {"stub-splitmix64"};
const gsl_rng_type *gsl_rng_ranlxs1 = &k_type;

static const double *g_u = 0, *g_n = 0;
static int g_nu = 0, g_nn = 0;

void gsl_stub_feed(const double *uniforms, int n_uniforms, const double *normals, int n_normals)
{
    g_u = uniforms; g_nu = n_uniforms;
    g_n = normals; g_nn = n_normals;
}

int gsl_stub_feed_left(int which) { return which ? g_nn : g_nu; }

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
    (void)T;
    gsl_rng *r = (gsl_rng *)calloc(1, sizeof(gsl_rng));
    r->s = 0x9E3779B97F4A7C15ull;
    return r;
}

void gsl_rng_set(gsl_rng *r, unsigned long seed)
{
    r->s = 0x9E3779B97F4A7C15ull ^ ((unsigned long long)seed * 0xD1342543DE82EF95ull);
    r->have = 0;
}

static unsigned long long next64(gsl_rng *r)
{
    unsigned long long z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

unsigned long gsl_rng_get(gsl_rng *r) { return (unsigned long)(next64(r) >> 40); }
unsigned long gsl_rng_max(const gsl_rng *r) { (void)r; return 0xFFFFFFul; }

double gsl_rng_uniform(gsl_rng *r)
{
    if (g_nu > 0) { g_nu--; return *g_u++; }
    return (double)(next64(r) >> 11) * (1.0 / 9007199254740992.0);
}

double gsl_ran_gaussian(gsl_rng *r, double sigma)
{
    if (g_nn > 0) { g_nn--; return sigma * *g_n++; }
    if (r->have) { r->have = 0; return sigma * r->spare; }
    double v1, v2, s;
    do {
        v1 = 2.0 * gsl_rng_uniform(r) - 1.0;
        v2 = 2.0 * gsl_rng_uniform(r) - 1.0;
        s = v1 * v1 + v2 * v2;
    } while (s >= 1.0 || s == 0.0);
    const double f = sqrt(-2.0 * log(s) / s);
    r->spare = v1 * f;
    r->have = 1;
    return sigma * v2 * f;
}

void gsl_rng_free(gsl_rng *r) { free(r); }
