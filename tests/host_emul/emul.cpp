// Host build of hb_device.cuh (TEST ONLY): lets the CPU test-suite exercise the exact source
// of the device model (prologue folding, Kepler solve, flux polynomial, eclipse) against the
// oracle without a GPU.  CUDA intrinsics are mapped to their IEEE host equivalents; the
// product never uses this file.
#include <cfenv>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __drcp_rn(double a) { volatile double r = 1.0 / a; return r; }
static inline double __fma_rd(double a, double b, double c)
{
    volatile double va = a, vb = b, vc = c;
    const int mode = fegetround();
    fesetround(FE_DOWNWARD);
    volatile double r = fma(va, vb, vc);
    fesetround(mode);
    return r;
}
static inline void hb_sincos(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
#define sincos hb_sincos
static inline double __longlong_as_double(long long x) { double d; memcpy(&d, &x, 8); return d; }
static inline long long __double_as_longlong(double x) { long long d; memcpy(&d, &x, 8); return d; }
static inline int __double2hiint(double x) { return (int)(__double_as_longlong(x) >> 32); }
static inline int __double2loint(double x) { return (int)(__double_as_longlong(x) & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo) { return __longlong_as_double(((long long)hi << 32) | (unsigned int)lo); }
static inline double __int2double_rn(int x) { return (double)x; }
static inline int __double2int_rn(double x) { return (int)nearbyint(x); }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int __double2int_rd(double x) { return (int)floor(x); }
#define HB_HOST_EMUL 1
#define __constant__ static const
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r = {x, y}; return r; }
#include "hb_device_host.cuh"
#include "../../hb_mcmc_b200/csrc/hb_sincos_tab.h"
using namespace hb;

// the sin/cos table of sincos_tab, filled exactly as hb_create does
static const double2* host_sctab()
{
    static double2 tab[kSinTabN];
    static bool ready = false;
    if (!ready) {
        fill_sincos_table(reinterpret_cast<double*>(tab));
        ready = true;
    }
    return tab;
}

static MagSetup default_mags(const double* md, const double* me, int g, int c)
{
    MagSetup ms;
    for (int i = 0; i < 5; i++) ms.mag_data[i] = md[i];
    for (int i = 0; i < 4; i++) ms.magerr[i] = me[i];
    ms.use_gmag = g;
    ms.use_color = c;
    return ms;
}

extern "C" int emul_const_size(void) { return (int)(sizeof(ChainConst) / sizeof(double)); }
// positions (in doubles) of the fields the tests look at
extern "C" int emul_const_flag_index(void) { return (int)(offsetof(ChainConst, flag) / sizeof(double)); }
extern "C" int emul_const_info_index(void) { return (int)(offsetof(ChainConst, info) / sizeof(double)); }

extern "C" void emul_prologue(const double* p, const double* md, const double* me, int g, int c, double* out)
{
    ChainConst cc;
    chain_prologue(p, default_mags(md, me, g, c), cc);
    memcpy(out, &cc, sizeof(cc));
}

// raw (un-normalised) template, likelihood3.c:673
// use_table: bit 0 = the kernel's E(M) starter table, bit 1 = the table sincos (both on in the kernel)
extern "C" void emul_raw(const double* p, const double* t, long n, double* out, int use_table)
{
    const double2* sctab = (use_table & 2) ? host_sctab() : nullptr;
    const double md[5] = {1000, 1, 1, 1, 1}, me[4] = {1e15, 1e15, 1e15, 1e15};
    ChainConst cc;
    chain_prologue(p, default_mags(md, me, 1, 0), cc);
    // the kernel's starter table for chains with 0 <= e <= kTableMaxE (use_table != 0)
    static double2 tab[2 * kTableNodes];
    const double2* ktab = nullptr;
    if ((use_table & 1) && cc.e >= 0.0 && cc.e <= kTableMaxE) {
        for (int k = 0; k < kTableSolved; k++) {
            double2 a, b, am, bm;
            kepler_table_node(k, cc.e, a, b);
            kepler_table_mirror(a, b, am, bm);
            tab[2 * (kTableN - k)] = am;
            tab[2 * (kTableN - k) + 1] = bm;
            tab[2 * k] = a;
            tab[2 * k + 1] = b;
        }
        ktab = tab;
    }
    long i = 0;
    for (; i + 2 <= n; i += 2) {  // the two-wide path the kernel uses
        const double ts[2] = {__dmul_rn(t[i], kSecDay), __dmul_rn(t[i + 1], kSecDay)};
        double u[2];
        if (sctab) raw_flux<2, true, true>(cc, ktab, sctab, ts, u);
        else raw_flux<2, true, false>(cc, ktab, sctab, ts, u);
        out[i] = u[0];
        out[i + 1] = u[1];
    }
    for (; i < n; i++)
        out[i] = sctab ? raw_flux1<false, true>(cc, ktab, sctab, __dmul_rn(t[i], kSecDay))
                       : raw_flux1<false, false>(cc, ktab, sctab, __dmul_rn(t[i], kSecDay));
}

// the E(M) starter of the hot loop: the chain's Taylor table built as emul_raw builds it, evaluated at m[0..n)
extern "C" void emul_table_guess(double e, const double* m, long n, double* out)
{
    static double2 tab[2 * kTableNodes];
    for (int k = 0; k < kTableSolved; k++) {
        double2 a, b, am, bm;
        kepler_table_node(k, e, a, b);
        kepler_table_mirror(a, b, am, bm);
        tab[2 * (kTableN - k)] = am;
        tab[2 * (kTableN - k) + 1] = bm;
        tab[2 * k] = a;
        tab[2 * k + 1] = b;
    }
    for (long i = 0; i < n; i++) out[i] = kepler_table_guess(tab, m[i]);
}

extern "C" void emul_finish(const double* u, long n, double med, double blend, double ft, double* out)
{
    for (long i = 0; i < n; i++) out[i] = finish_template(u[i], med, blend, ft);
}

extern "C" double emul_fmod_twopi(double M) { return fmod_twopi(M); }

extern "C" void emul_sincos(const double* x, long n, double* s, double* c)
{
    for (long i = 0; i < n; i++) {
        const double xv[1] = {x[i]};
        double sv[1], cv[1];
        int hm = 0;
        sincos_lean<1>(xv, sv, cv, hm);
        if (hm > kSincosHiLimit) { sv[0] = sin(x[i]); cv[0] = cos(x[i]); }
        s[i] = sv[0];
        c[i] = cv[0];
    }
}
extern "C" void emul_sincos_tab(const double* x, long n, double* s, double* c)
{
    for (long i = 0; i < n; i++) {
        const double xv[1] = {x[i]};
        double sv[1], cv[1];
        int hm = 0;
        sincos_tab<1>(xv, host_sctab(), sv, cv, hm);
        if (hm > kSincosTabHiLimit) { sv[0] = sin(x[i]); cv[0] = cos(x[i]); }
        s[i] = sv[0];
        c[i] = cv[0];
    }
}
extern "C" void emul_sincos_table(double* out) { memcpy(out, host_sctab(), sizeof(double2) * kSinTabN); }
extern "C" void emul_div(const double* a, const double* b, long n, double* q, double* r)
{
    for (long i = 0; i < n; i++) { q[i] = div_fast(a[i], b[i]); r[i] = rcp_fast(b[i]); }
}
// the Markstein phase division of kepler_point: must equal IEEE x / P
extern "C" void emul_phase_div(const double* x, const double* P, long n, double* out)
{
    for (long i = 0; i < n; i++) {
        const double rP = __drcp_rn(P[i]);
        const double q0 = __dmul_rn(x[i], rP);
        out[i] = fma(fma(-P[i], q0, x[i]), rP, q0);
    }
}
