// hb_device.cuh -- device-side model of the heartbeat-star light curve (sm_100a, FP64).
//
// What is computed follows the reference (paths relative to /root/reference/src):
//   likelihood3.c:125-185  traj           -> kepler_point()  (phase, fmod, starter, 5 Newton steps)
//   likelihood3.c:224-337  beaming/ellipsoidal/reflection -> folded into 12 per-chain
//                          coefficients by chain_prologue(), evaluated by raw_flux()
//   likelihood3.c:353-389  eclipse_area   -> eclipse_area_dev()
//   likelihood3.c:396-507,693-717 stellar relations -> dev_getT/_getR/...
//   likelihood3.c:725-795  calc_mags, GAIA_mcmc.c:198-250 get_mags -> two_bb_mags()
//   likelihood3.c:945-974  RocheOverflow  -> roche_overflow_dev()
// HOW it is computed is not the reference's: everything that does not depend on the time
// sample is hoisted into ChainConst once per chain, the true anomaly is never formed
// (cos/sin nu come algebraically from cos/sin E), the harmonics of omega+nu come from
// Chebyshev recurrences, and star 2 (omega+pi) reuses star 1's harmonics with signs.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#ifndef HB_NEWTON_UNROLL
#define HB_NEWTON_UNROLL 5
#endif

#ifndef HB_CHK  // (defined by hb_select.cuh in the bounds-asserting debug build; the host emulation never checks)
#define HB_CHK(i, cap, site) ((void)0)
#endif

namespace hb {

constexpr int NPARS = 21;  // likelihood3.h:20
constexpr int kNewtonUnroll = HB_NEWTON_UNROLL;  // 1 = rolled Newton loop (small I-cache footprint)

// Which samples may start Newton from the per-chain E(M) table (see the table section below).
// Chains up to kTableAllE use the table at every sample.  Between kTableAllE and kTableMaxE the reference's
// five steps still converge to rounding EXCEPT within |M| <= 0.0493 of periastron (largest at e = 0.91;
// tests/tools/kepler_convergence_scan.c, M on a log + uniform grid, e = 0.80 .. 0.99): such chains take the
// table starter only for samples at least kTableMinM = 0.1 rad of mean anomaly away from periastron and
// the reference's own starter -- hence its exact, un-converged iterates -- inside that window.
constexpr double kTableAllE = 0.8;
constexpr double kTableMaxE = 0.99;
constexpr double kTableMinM = 0.1;

// physical constants, likelihood3.h:4-10,31
constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 2.0 * 3.14159265358979323846;  // == fl(2*PI) used by fmod(M, 2*PI)
constexpr double kTwoPiExact = 6.283185307179586476925;   // nearest double to 2 pi (the same number)
constexpr double kG = 6.6743e-8;
constexpr double kC = 2.998e10;
constexpr double kMsun = 1.9885e33;
constexpr double kRsun = 6.955e10;
constexpr double kSecDay = 86400.0;
constexpr double kBig = 1.e15;

// Per-chain folded constants (written by chain_prologue, read by the point kernels).
struct ChainConst {
    // orbit
    double e;          // eccentricity
    double T0s, Ps;    // T0*86400, P[s]  (exactly the reference's T0_cgs / P_cgs)
    double rPs;        // RN(1/Ps): seed of the exact phase division
    double cw, sw;     // cos/sin omega0
    double cwq, swq;   // cos/sin omega0 times sqrt(1 - e^2)
    double tab_min_m;  // 0: the E(M) table serves every sample; > 0: only samples this far (rad) from periastron
    double ci, si;     // cos/sin inc
    double ar;         // a / RSUN
    // raw flux u = K0 + K1 c + b^2 (a0 + a1 s + a2 c2 + b (b0 + b1 c2 + b (c1 s + c3 s3 + b (d0 + d2 c2 + d4 c4))))
    // with the two highest harmonics rewritten in c2 = cos 2x:  sin 3x = s (1 + 2 c2),  cos 4x = 2 c2^2 - 1
    //   c1 s + c3 s3 = s (q1 + q3 c2),          q1 = c1 + c3,  q3 = 2 c3
    //   d0 + d2 c2 + d4 c4 = r0 + c2 (d2 + r4 c2),  r0 = d0 - d4,  r4 = 2 d4
    // STORED for h = c2 / 2 = c^2 - 1/2, which the sample loop forms in one operation (cos 2x takes two): a2, b1, d2
    // and q3 hold twice, r4 four times the coefficient named above -- exact scalings, so every FMA of the polynomial
    // rounds the very number it would round with c2.
    double K0, K1, a0, a1, a2, b0, b1, q1, q3, r0, d2, r4;
    // eclipse pre-test on squared quantities: (1 - e cos E)^2 (cos^2 i + sin^2 i c^2) < thr with
    // c^2 = (1 + c2)/2 = 1/2 + h: si2 = sin^2 i (multiplies h), ci2 = cos^2 i + sin^2 i / 2, thr = ((R1 + R2)/ar)^2 (1 + 1e-9);
    // tested as cos^2 i + sin^2 i c^2 < thr beta^2 (beta = 1/(1 - e cos E), whose square the flux needs anyway)
    double si2, ci2, thr;
    // eclipse (radii in Rsun, sorted big/small as eclipse_area does)
    double Rb, Rs, ecl1, ecl2;  // ecl_k = Norm_k / (pi R_k^2)
    // normalisation: model = A (u - median) + ft,  A = ft (1 - blending)
    double blend, ft;
    // per-chain additive chi^2 (Gaia magnitude / colour terms) and flags
    double chi2_extra;
    double flag;  // bit 0: Roche overflow, bit 1: model is NaN by construction (e >= 1 or NaN e),
                  // bit 2: the projected separation never drops below R1 + R2 (no sample can be in eclipse)
    double seed;  // 32-bit hash of the parameter bits: sampling jitter of the select (a chain's
                  // result does not depend on where it sits in the batch)
    // diagnostics (hb_chain_info): R1 R2 T1 T2 G B-V V-G G-T
    double info[8];
};

struct MagSetup {
    double mag_data[5];  // D, G, B-V, V-G, G-T   (mcmc_wrapper2.c:302-317)
    double magerr[4];
    int use_gmag, use_color;  // likelihood3.h:11-12 made runtime
};

__device__ __forceinline__ double sq(double x) { return x * x; }

// ---------------------------------------------------------------------------
// stellar relations (likelihood3.c:396-507)
// ---------------------------------------------------------------------------
__device__ inline double interp_nodes(double m, const double* mn, const double* yn, int n)
{
    if (m <= mn[0]) return yn[0];
    if (m >= mn[n - 1]) return yn[n - 1];
    double r = yn[n - 1];
    for (int j = 1; j < n; j++) {
        if (m < mn[j]) {
            r = yn[j - 1] + (m - mn[j - 1]) * (yn[j] - yn[j - 1]) / (mn[j] - mn[j - 1]);
            break;
        }
    }
    return r;
}

// m = 10^logM is passed in by the prologue (one exp10 per star instead of three pow calls)
__device__ inline double dev_getT_m(double m)
{
    const double mn[16] = {0.1, 0.26, 0.47, 0.59, 0.69, 0.87, 0.98, 1.085, 1.4, 1.65, 2.0, 2.5, 3.0, 4.4, 15., 40.};
    const double tn[16] = {3.491, 3.531, 3.547, 3.584, 3.644, 3.712, 3.745, 3.774,
                           3.823, 3.863, 3.913, 3.991, 4.057, 4.182, 4.477, 4.623};
    return interp_nodes(m, mn, tn, 16);
}

__device__ inline double dev_getR_m(double m)
{
    const double mn[10] = {0.07, 0.2, 0.356, 0.655, 0.784, 0.787, 1.377, 4.4, 15., 40.};
    const double rn[10] = {-0.953, -0.627, -0.423, -0.154, -0.082, -0.087, 0.295, 0.477, 0.792, 1.041};
    return interp_nodes(m, mn, rn, 10);
}

__device__ inline double dev_envelope_radius_m(double m)
{
    const double n = 4.22, slope = 15.68, floor_ = 0.01, corner = 1.055, ceil_ = 0.17;
    return 1 / (1 / ceil_ + 1 / (slope * pow((pow(m, n) + pow(corner, n)), (1 / n)) - (slope * corner - floor_)));
}

__device__ inline double dev_getT(double logM)
{
    const double mn[16] = {0.1, 0.26, 0.47, 0.59, 0.69, 0.87, 0.98, 1.085, 1.4, 1.65, 2.0, 2.5, 3.0, 4.4, 15., 40.};
    const double tn[16] = {3.491, 3.531, 3.547, 3.584, 3.644, 3.712, 3.745, 3.774,
                           3.823, 3.863, 3.913, 3.991, 4.057, 4.182, 4.477, 4.623};
    return interp_nodes(pow(10., logM), mn, tn, 16);
}

__device__ inline double dev_getR(double logM)
{
    const double mn[10] = {0.07, 0.2, 0.356, 0.655, 0.784, 0.787, 1.377, 4.4, 15., 40.};
    const double rn[10] = {-0.953, -0.627, -0.423, -0.154, -0.082, -0.087, 0.295, 0.477, 0.792, 1.041};
    return interp_nodes(pow(10., logM), mn, rn, 10);
}

__device__ inline double dev_envelope_temp(double) { return 0.0224; }

__device__ inline double dev_envelope_radius(double logM)
{
    const double n = 4.22, slope = 15.68, floor_ = 0.01, corner = 1.055, ceil_ = 0.17;
    double m = pow(10., logM);
    return 1 / (1 / ceil_ + 1 / (slope * pow((pow(m, n) + pow(corner, n)), (1 / n)) - (slope * corner - floor_)));
}

__device__ inline double dev_alpha_beam(double logT)  // likelihood3.c:194-209
{
    const double al[4] = {6.5, 4.0, 2.5, 1.2};
    const double lt[4] = {3.5, 3.7, 3.9, 4.5};
    if (logT >= lt[3]) return 1.2 / 4;
    if (logT < lt[0]) return 6.5 / 4;
    int j = 3;
    while (logT < lt[j]) j--;
    return ((al[j + 1] + (al[j + 1] - al[j]) / (lt[j + 1] - lt[j]) * (logT - lt[j + 1])) / 4);
}

// likelihood3.c:760-789 (gaia == 0: exp()-1 and /(1-blending)) and GAIA_mcmc.c:214-249
// (gaia == 1: expm1, no blending).  out = {G, B-V, V-G, G-T}
__device__ inline void two_bb_mags(double R1, double R2, double T1, double T2, double D, double blending, int gaia,
                                   double out[4])
{
    const double lam[4] = {442, 540, 673, 750};
    const double h = 6.626e-27, k = 1.38e-16, pc = 3.086e18;
    double mag[4];
    R1 *= kRsun;
    R2 *= kRsun;
    for (int j = 0; j < 4; j++) {
        double nu = kC / (lam[j] * 1e-7);
        double pl = 2. * h * (nu * nu * nu) / sq(kC);
        double f;
        if (!gaia) {
            f = kPi * (R1 * R1 * (pl / (exp(h * nu / (k * T1)) - 1.)) + R2 * R2 * (pl / (exp(h * nu / (k * T2)) - 1.))) /
                (sq(D) * sq(pc));
            f = f / (1 - blending);
        } else {
            f = kPi * (sq(R1) * (pl / expm1(h * nu / (k * T1))) + sq(R2) * (pl / expm1(h * nu / (k * T2)))) /
                (sq(D) * sq(pc));
        }
        mag[j] = -2.5 * log10(f) - 48.6;
    }
    out[0] = mag[2];
    out[1] = mag[0] - mag[1];
    out[2] = mag[1] - mag[2];
    out[3] = mag[2] - mag[3];
}

// GAIA_mcmc.c:198-250 get_mags: {G, B-V, V-G, G-T} of the 6-parameter layout {logM1, logM2, rr1, rr2, aT1, aT2}
__device__ inline void gaia_mags(const double* p, double D, double m[4])
{
    // _getR, _getT and envelope_Radius each start with the same pow(10., logM): evaluate it once per star
    const double m1 = pow(10., p[0]), m2 = pow(10., p[1]);
    const double R1 = pow(10., dev_getR_m(m1) + p[2] * dev_envelope_radius_m(m1));
    const double R2 = pow(10., dev_getR_m(m2) + p[3] * dev_envelope_radius_m(m2));
    const double T1 = pow(10., dev_getT_m(m1) + p[4] * dev_envelope_temp(p[0]));
    const double T2 = pow(10., dev_getT_m(m2) + p[5] * dev_envelope_temp(p[1]));
    two_bb_mags(R1, R2, T1, T2, D, 0., 1, m);
}

__device__ inline double eggleton_dev(double q)  // likelihood3.c:945-948
{
    const double q13 = cbrt(q), q23 = q13 * q13;
    return 0.49 * q23 / (0.6 * q23 + log(1 + q13));
}

// ---------------------------------------------------------------------------
// chain prologue: everything that does not depend on the time sample
// ---------------------------------------------------------------------------
// Every transcendental value the prologue needs.  The sequential path (one thread per chain) and
// the warp path (one warp per chain, all lanes calling the same libm routine on different arguments)
// fill this struct with the same calls on the same arguments; prologue_assemble() is pure arithmetic.
//
// Everything that feeds the orbit phase or the eclipse geometry (P, M, R, a) keeps the reference's
// pow() calls: CUDA's pow and glibc's agree to the last bit almost always, and an ulp there is
// amplified ~1e5 x by the ill-conditioned asin(h/R) of eclipse_area (measured: worst logL
// deviation 1.1e-11 with pow, 8.5e-11 with exp10/cbrt).  10^logM is formed once per star, and
// the flux COEFFICIENTS (smooth, O(1e-3) terms) use cbrt instead of pow(x, k/3).
struct PrologueT {
    double M[2], Pd, xb[2], si, ci, sw, cw;  // 10^logM, 10^logP, exp(p15|16), sin/cos inc, sin/cos omega0
    double sq1me2, sq1me;                    // sqrt(1 - e^2), sqrt(1 - e)
    double pm[2], cn, envroot[2];            // M^4.22, 1.055^4.22, (pm + cn)^(1/4.22)   (envelope_Radius)
    double Te[2], R[2], lTe[2];              // Teff, radius [Rsun], log10 Teff
    double cM[2], cq[2], cP;                 // cbrt(M_k), cbrt(1 + q_k), cbrt(Pd)
    double a, sep, eq[2], elog[2];           // a [cm] (traj), Roche separation, cbrt(q), cbrt(1/q), log(1 + .)
    double ex[2][4], lf[4];                  // exp(h nu_j / k T_s), log10 f_nu_j
};

constexpr double kEnvN = 4.22, kEnvCorner = 1.055;  // envelope_Radius, likelihood3.c:497-501
__device__ __forceinline__ double env_from_root(double root)
{
    const double slope = 15.68, floor_ = 0.01, ceil_ = 0.17;
    return 1 / (1 / ceil_ + 1 / (slope * root - (slope * kEnvCorner - floor_)));
}
__device__ __forceinline__ double te_arg(const double* p, int k, double m) { return dev_getT_m(m) + p[17 + k] * 0.0224; }
__device__ __forceinline__ double r_arg(const double* p, int k, double m, double root)
{
    return dev_getR_m(m) + p[7 + k] * env_from_root(root);
}
__device__ __forceinline__ double a_arg(const double* M, double Ps)  // likelihood3.c:141-142
{
    const double Mtot = M[0] * kMsun + M[1] * kMsun;
    return kG * Mtot * sq(Ps) / sq(2 * kPi);
}
__device__ __forceinline__ double sep_arg(const double* M, double Ps)  // likelihood3.c:962
{
    return kG * (M[0] * kMsun + M[1] * kMsun) * sq(Ps) / (4.0 * kPi * kPi);
}
// blackbody pieces of calc_mags (likelihood3.c:760-789)
__device__ __forceinline__ double bb_nu(int j)
{
    const double lam[4] = {442, 540, 673, 750};
    return kC / (lam[j] * 1e-7);
}
__device__ __forceinline__ double bb_exp_arg(int j, double T) { return 6.626e-27 * bb_nu(j) / (1.38e-16 * T); }
__device__ __forceinline__ double bb_flux(int j, double R1, double R2, double e1, double e2, double D, double blending)
{
    const double h = 6.626e-27, pc = 3.086e18, nu = bb_nu(j);
    const double pl = 2. * h * (nu * nu * nu) / sq(kC);
    R1 *= kRsun;
    R2 *= kRsun;
    const double f = kPi * (R1 * R1 * (pl / (e1 - 1.)) + R2 * R2 * (pl / (e2 - 1.))) / (sq(D) * sq(pc));
    return f / (1 - blending);
}

// sequential fill: one thread evaluates the ~50 calls one after the other
__device__ inline void prologue_trans_seq(const double* __restrict__ p, const MagSetup& ms, PrologueT& T)
{
    const double e = p[3];
    T.M[0] = pow(10., p[0]);
    T.M[1] = pow(10., p[1]);
    T.Pd = pow(10., p[2]);
    T.xb[0] = exp(p[15]);
    T.xb[1] = exp(p[16]);
    sincos(p[4], &T.si, &T.ci);
    sincos(p[5], &T.sw, &T.cw);
    T.sq1me2 = sqrt(1 - sq(e));
    T.sq1me = sqrt(1 - e);
    T.cn = pow(kEnvCorner, kEnvN);
    const double Ps = T.Pd * kSecDay;
    for (int k = 0; k < 2; k++) {
        const double q = T.M[1 - k] / T.M[k];
        T.pm[k] = pow(T.M[k], kEnvN);
        T.envroot[k] = pow(T.pm[k] + T.cn, 1 / kEnvN);
        T.Te[k] = pow(10., te_arg(p, k, T.M[k]));
        T.R[k] = pow(10., r_arg(p, k, T.M[k], T.envroot[k]));
        T.lTe[k] = log10(T.Te[k]);
        T.cM[k] = cbrt(T.M[k]);
        T.cq[k] = cbrt(1 + q);
        for (int j = 0; j < 4; j++) T.ex[k][j] = exp(bb_exp_arg(j, T.Te[k]));
    }
    T.cP = cbrt(T.Pd);
    T.a = pow(a_arg(T.M, Ps), 1. / 3.);
    T.sep = cbrt(sep_arg(T.M, Ps));
    const double qr = (T.M[0] * kMsun) / (T.M[1] * kMsun);
    T.eq[0] = cbrt(qr);
    T.eq[1] = cbrt(1 / qr);
    T.elog[0] = log(1 + T.eq[0]);
    T.elog[1] = log(1 + T.eq[1]);
    for (int j = 0; j < 4; j++)
        T.lf[j] = log10(bb_flux(j, T.R[0], T.R[1], T.ex[0][j], T.ex[1][j], ms.mag_data[0], p[19]));
}

#ifndef HB_HOST_EMUL
// warp fill: the lanes of one warp call each libm routine ONCE, on different arguments, and
// exchange the results by shuffle -- five dependent levels instead of ~50 sequential calls.
__device__ __forceinline__ void prologue_trans_warp(const double* __restrict__ p, const MagSetup& ms, PrologueT& T, int lane)
{
    const unsigned full = 0xffffffffu;
    const double e = p[3];
    auto bc = [&](double v, int src) { return __shfl_sync(full, v, src); };
    // level 0: pow lanes 0-3, exp lanes 0-1, sincos lanes 0-1, sqrt lanes 0-1
    {
        const double pa = (lane == 3) ? kEnvCorner : 10., pb = (lane < 3) ? p[lane] : ((lane == 3) ? kEnvN : 0.);
        const double pw = pow(pa, pb);
        const double ex = exp(p[15 + (lane & 1)]);
        double sn, cs;
        sincos(p[4 + (lane & 1)], &sn, &cs);
        const double sr = sqrt((lane & 1) ? (1 - e) : (1 - sq(e)));
        T.M[0] = bc(pw, 0); T.M[1] = bc(pw, 1); T.Pd = bc(pw, 2); T.cn = bc(pw, 3);
        T.xb[0] = bc(ex, 0); T.xb[1] = bc(ex, 1);
        T.si = bc(sn, 0); T.ci = bc(cs, 0); T.sw = bc(sn, 1); T.cw = bc(cs, 1);
        T.sq1me2 = bc(sr, 0); T.sq1me = bc(sr, 1);
    }
    const double Ps = T.Pd * kSecDay;
    const double qr = (T.M[0] * kMsun) / (T.M[1] * kMsun);
    // level 1: pow lanes 0-1 (M^4.22), 2-3 (Teff), 4 (a); cbrt lanes 0-1 (M), 2-3 (1+q), 4 (Pd), 5 (sep), 6-7 (q, 1/q)
    {
        const int k = lane & 1;
        double pa = 1., pb = 1.;
        if (lane < 2) { pa = T.M[k]; pb = kEnvN; }
        else if (lane < 4) { pa = 10.; pb = te_arg(p, k, T.M[k]); }
        else if (lane == 4) { pa = a_arg(T.M, Ps); pb = 1. / 3.; }
        const double pw = pow(pa, pb);
        double ca = 1.;
        if (lane < 2) ca = T.M[k];
        else if (lane < 4) ca = 1 + T.M[1 - k] / T.M[k];
        else if (lane == 4) ca = T.Pd;
        else if (lane == 5) ca = sep_arg(T.M, Ps);
        else if (lane == 6) ca = qr;
        else if (lane == 7) ca = 1 / qr;
        const double cb = cbrt(ca);
        T.pm[0] = bc(pw, 0); T.pm[1] = bc(pw, 1); T.Te[0] = bc(pw, 2); T.Te[1] = bc(pw, 3); T.a = bc(pw, 4);
        T.cM[0] = bc(cb, 0); T.cM[1] = bc(cb, 1); T.cq[0] = bc(cb, 2); T.cq[1] = bc(cb, 3);
        T.cP = bc(cb, 4); T.sep = bc(cb, 5); T.eq[0] = bc(cb, 6); T.eq[1] = bc(cb, 7);
    }
    // level 2: pow lanes 0-1 (envelope root); log10 lanes 0-1 (Teff); exp lanes 0-7 (blackbody); log lanes 0-1
    {
        const int k = lane & 1;
        const double pw = pow(T.pm[k] + T.cn, 1 / kEnvN);
        const double lg = log10(T.Te[k]);
        const double ex = exp(bb_exp_arg((lane >> 1) & 3, T.Te[k]));  // lane = 2 j + k
        const double ln = log(1 + T.eq[k]);
        T.envroot[0] = bc(pw, 0); T.envroot[1] = bc(pw, 1);
        T.lTe[0] = bc(lg, 0); T.lTe[1] = bc(lg, 1);
        for (int j = 0; j < 4; j++) { T.ex[0][j] = bc(ex, 2 * j); T.ex[1][j] = bc(ex, 2 * j + 1); }
        T.elog[0] = bc(ln, 0); T.elog[1] = bc(ln, 1);
    }
    // level 3: radii
    {
        const int k = lane & 1;
        const double pw = pow(10., r_arg(p, k, T.M[k], T.envroot[k]));
        T.R[0] = bc(pw, 0); T.R[1] = bc(pw, 1);
    }
    // level 4: magnitudes
    {
        const int j = lane & 3;
        const double lg = log10(bb_flux(j, T.R[0], T.R[1], T.ex[0][j], T.ex[1][j], ms.mag_data[0], p[19]));
        for (int i = 0; i < 4; i++) T.lf[i] = bc(lg, i);
    }
}
#endif

// pure arithmetic: PrologueT -> ChainConst, in four independent sections (each writes its own fields of cc from p and T
// alone), so that a caller with idle warps can run them side by side (k_pt_run); prologue_assemble runs all four.
// The same expressions in the same order whichever way they are called: the same bits.
constexpr unsigned kAsmStarA = 1u;  // K0 K1 a0 a1 a2 b0 b1: beaming, reflection, the R^3 ellipsoidal terms
constexpr unsigned kAsmStarB = 2u;  // q1 q3 r0 d2 r4: the R^4 and R^5 ellipsoidal terms
constexpr unsigned kAsmOrbit = 4u;  // orbit, geometry, eclipse scale, normalisation, flags
constexpr unsigned kAsmAux = 8u;    // Gaia chi^2 term, seed, diagnostics
constexpr unsigned kAsmAll = 15u;

template <unsigned kSections>
__device__ __forceinline__ void prologue_assemble_sections(const double* __restrict__ p, const MagSetup& ms, const PrologueT& T,
                                                           ChainConst& cc)
{
    const double Pd = T.Pd;
    const double e = p[3], T0 = p[6];
    const double mu[2] = {p[9], p[11]}, tau[2] = {p[10], p[12]};
    const double aref[2] = {p[13], p[14]};
    const double blending = p[19], ft = p[20];
    const double* M = T.M;
    const double* R = T.R;
    const double* Te = T.Te;
    // luminosity fractions (likelihood3.c:612-614)
    const double L1 = sq(R[0]) * sq(sq(Te[0])), L2 = sq(R[1]) * sq(sq(Te[1]));
    const double Nrm[2] = {L1 / (L1 + L2), L2 / (L1 + L2)};
    const double si = T.si, ci = T.ci;
    const double si2 = si * si, si3 = si2 * si, si4 = si2 * si2;
    const double ppm = 1.e-6;

    if constexpr ((kSections & (kAsmStarA | kAsmStarB)) != 0) {
        // beaming alphas (likelihood3.c:617-624)
        const double ab[2] = {dev_alpha_beam(T.lTe[0]) * T.xb[0], dev_alpha_beam(T.lTe[1]) * T.xb[1]};
        const double Pm13 = 1.0 / T.cP, Pm43 = Pm13 / Pd, Pm83 = Pm43 * Pm43, Pm103 = Pm83 * Pm13 * Pm13;
        const double Prot = Pd * ((1 - e) * T.sq1me);
        double K0 = 0, K1 = 0, a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, c1 = 0, c3 = 0, d0 = 0, d2 = 0, d4 = 0;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double Ma = M[k], Mb = M[1 - k];
            const double Rs = R[k], Ro = R[1 - k];  // own radius (ellipsoidal), companion radius (reflection)
            const double sgn = (k == 0) ? 1.0 : -1.0;  // star 2 sees omega0 + pi: odd harmonics flip
            const double N = Nrm[k];
            const double q = Mb / Ma;
            const double Ma13 = T.cM[k], q13 = T.cq[k];                          // Ma^(1/3), (1+q)^(1/3)
            const double iMa23 = 1.0 / (Ma13 * Ma13), iq23 = 1.0 / (q13 * q13);  // ^(-2/3)
            // ellipsoidal coefficient set, likelihood3.c:258-264
            const double al11 = 15 * mu[k] * (2 + tau[k]) / (32 * (3 - mu[k]));
            const double al21 = 3 * (15 + mu[k]) * (1 + tau[k]) / (20 * (3 - mu[k]));
            const double al2b1 = 15 * (1 - mu[k]) * (3 + tau[k]) / (64 * (3 - mu[k]));
            const double al01 = al21 / 9, al0b1 = 3 * al2b1 / 20, al31 = 5 * al11 / 3, al41 = 7 * al2b1 / 4;
            const double R3 = Rs * Rs * Rs, R4 = R3 * Rs, R5 = R4 * Rs;
            if constexpr ((kSections & kAsmStarA) != 0) {
                // beaming, likelihood3.c:224-236 (pow(1+q, 2/3) == 1, quirk Q1)
                const double B = -2830. * ab[k] * q * Ma13 * Pm13 * si / T.sq1me2 * ppm;
                const double qq = q / (1 + q);
                const double AM1 = 13435. * 2 * al01 * (2 - 3 * si2) / Ma / sq(Prot) * R3 * ppm;
                const double AM2 = 13435. * 3 * al01 * (2 - 3 * si2) / Ma * qq / sq(Pd) * R3 * ppm;       // x beta^3
                const double C21 = 13435. * al21 * si2 / Ma * qq / sq(Pd) * R3 * ppm;                     // x beta^3 cos2x
                // reflection, likelihood3.c:322-337
                const double Rf = 56514. * aref[k] * iq23 * iMa23 * Pm43 * sq(Ro) * ppm;  // x beta^2
                K0 += N * (1 + AM1);
                K1 += sgn * N * B;
                a0 += N * Rf * (0.64 + 0.18 * si2);
                a1 += -sgn * N * Rf * si;
                a2 += -N * Rf * 0.18 * si2;
                b0 += N * AM2;
                b1 += N * C21;
            }
            if constexpr ((kSections & kAsmStarB) != 0) {
                const double Mm53 = (iMa23 / Ma) * q * (iq23 / (1 + q)) * Pm103;        // Ma^-5/3 q (1+q)^-5/3 P^-10/3
                const double Mm43 = (iMa23 * iMa23) * q * (iq23 * iq23) * Pm83;         // Ma^-4/3 q (1+q)^-4/3 P^-8/3
                const double AM3 = 759. * al0b1 * (8 - 40 * si2 + 35 * si4) * Mm53 * R5 * ppm;            // x beta^5
                const double S1 = 3194. * al11 * (4 * si - 5 * si3) * Mm43 * R4 * ppm;                    // x beta^4 sin x
                const double C22 = 759. * al2b1 * (6 * si2 - 7 * si4) * Mm53 * R5 * ppm;                  // x beta^5 cos2x
                const double S3 = 3194. * al31 * si3 * Mm43 * R4 * ppm;                                   // x beta^4 sin3x
                const double C4 = 759. * al41 * si4 * Mm53 * R5 * ppm;                                    // x beta^5 cos4x
                c1 += sgn * N * S1;
                c3 += sgn * N * S3;
                d0 += N * AM3;
                d2 += N * C22;
                d4 += N * C4;
            }
        }
        if constexpr ((kSections & kAsmStarA) != 0) {
            // (the sample loop works with h = cos 2x / 2 = c^2 - 1/2, one operation less than cos 2x: the coefficients of
            // cos 2x carry the factor 2 -- exact, so every FMA of the polynomial rounds the same number as before)
            cc.K0 = K0; cc.K1 = K1; cc.a0 = a0; cc.a1 = a1; cc.a2 = 2.0 * a2; cc.b0 = b0; cc.b1 = 2.0 * b1;
        }
        if constexpr ((kSections & kAsmStarB) != 0) {
            cc.q1 = c1 + c3; cc.q3 = 4.0 * c3; cc.r0 = d0 - d4; cc.d2 = 2.0 * d2; cc.r4 = 8.0 * d4;  // (x 2 per power of h)
        }
    }

    if constexpr ((kSections & kAsmOrbit) != 0) {
        cc.e = e;
        cc.T0s = T0 * kSecDay;
        cc.Ps = Pd * kSecDay;
        cc.rPs = __drcp_rn(cc.Ps);
        cc.cw = T.cw;
        cc.sw = T.sw;
        cc.tab_min_m = (e > kTableAllE) ? kTableMinM : 0.0;
        cc.cwq = T.cw * T.sq1me2;
        cc.swq = T.sw * T.sq1me2;
        cc.ci = ci;
        cc.si = si;
        const double ar = T.a / kRsun;  // semi-major axis as traj() forms it (likelihood3.c:141-142)
        cc.ar = ar;
        const double Rb = fmax(R[0], R[1]), Rsm = fmin(R[0], R[1]);
        cc.Rb = Rb;
        cc.Rs = Rsm;
        cc.si2 = 2.0 * (0.5 * si * si);  // (multiplies h = cos 2x / 2)
        cc.ci2 = ci * ci + 0.5 * si * si;
        {
            const double lim = (Rb + Rsm) / ar;
            cc.thr = lim * lim * (1.0 + 1e-9);
        }
        cc.ecl1 = Nrm[0] / (kPi * sq(R[0]));
        cc.ecl2 = Nrm[1] / (kPi * sq(R[1]));
        cc.blend = blending;
        cc.ft = ft;
        // Roche overflow (likelihood3.c:945-974)
        int roche;
        {
            const double r1 = R[0] * kRsun / (T.sep * (1 - e));
            const double r2 = R[1] * kRsun / (T.sep * (1 - e));
            const double q23a = T.eq[0] * T.eq[0], q23b = T.eq[1] * T.eq[1];
            const double RL1 = 0.49 * q23a / (0.6 * q23a + T.elog[0]);
            const double RL2 = 0.49 * q23b / (0.6 * q23b + T.elog[1]);
            roche = ((RL1 < r1) || (RL2 < r2)) ? 1 : 0;
        }
        // e >= 1 (reachable, quirk Q4) or NaN e: the reference's template is NaN at every sample
        const int nan_model = !(e < 1.0) ? 1 : 0;
        // d >= r_min |cos i| with r_min = a (1 - |e|): chains that can never eclipse skip the per-sample test
        const int no_eclipse = (ar * (1 - fabs(e)) * fabs(ci) >= (Rb + Rsm) * (1.0 + 1e-9)) ? 1 : 0;
        cc.flag = (double)(roche | (nan_model << 1) | (no_eclipse << 2));
    }

    if constexpr ((kSections & kAsmAux) != 0) {
        // Gaia magnitude / colour chi^2 terms (likelihood3.c:780-789, 834-860)
        const double Bm = -2.5 * T.lf[0] - 48.6, Vm = -2.5 * T.lf[1] - 48.6, Gm = -2.5 * T.lf[2] - 48.6, Tm = -2.5 * T.lf[3] - 48.6;
        const double mags[4] = {Gm, Bm - Vm, Vm - Gm, Gm - Tm};
        double extra = 0.;
        if (ms.use_gmag) {
            double r = (mags[0] - ms.mag_data[1]) / ms.magerr[0];
            extra += r * r;
        }
        if (ms.use_color) {
            for (int i = 1; i < 4; i++) {
                double r = (mags[i] - ms.mag_data[i + 1]) / ms.magerr[i];
                extra += r * r;
            }
        }
        cc.chi2_extra = extra;
        {
            uint32_t h = 0x811c9dc5u;
            for (int i = 0; i < NPARS; i++) {
                const unsigned long long b = (unsigned long long)__double_as_longlong(p[i]);
                h = (h ^ (uint32_t)b) * 0x01000193u;
                h = (h ^ (uint32_t)(b >> 32)) * 0x01000193u;
            }
            cc.seed = (double)h;
        }
        cc.info[0] = R[0]; cc.info[1] = R[1]; cc.info[2] = Te[0]; cc.info[3] = Te[1];
        cc.info[4] = mags[0]; cc.info[5] = mags[1]; cc.info[6] = mags[2]; cc.info[7] = mags[3];
    }
}

__device__ inline void prologue_assemble(const double* __restrict__ p, const MagSetup& ms, const PrologueT& T, ChainConst& cc)
{
    prologue_assemble_sections<kAsmAll>(p, ms, T, cc);
}

// one thread per chain (host emulation, small helpers)
__device__ inline void chain_prologue(const double* __restrict__ p, const MagSetup& ms, ChainConst& cc)
{
    PrologueT T;
    prologue_trans_seq(p, ms, T);
    prologue_assemble(p, ms, T, cc);
}

// ---------------------------------------------------------------------------
// per-sample pieces
// ---------------------------------------------------------------------------

// ---- lean FP64 primitives --------------------------------------------------------------
// The CUDA library sincos()/division cost ~3 non-FP64 instructions per FP64 one (range and
// special-case handling, 64-bit immediates moved through uniform registers); measured with
// ncu that made the model pass issue-bound at 34 % FP64-pipe utilisation.  The versions below
// keep the FP64 pipe as the only busy pipe: coefficients sit in the constant bank (operands of
// DFMA, no UMOV), quadrant logic is 9 integer ops, divisions are a MUFU seed + Newton steps.

#ifndef HB_HOST_EMUL
// true when the predicate holds on every lane that is executing this call together
template <bool kFullWarp>
__device__ __forceinline__ bool warp_all(bool p)
{
    return kFullWarp ? __all_sync(0xffffffffu, p) : __all_sync(__activemask(), p);
}
#else
template <bool kFullWarp>
static inline bool warp_all(bool p) { return p; }
#endif

#ifndef HB_HOST_EMUL
__device__ __forceinline__ double rcp_seed(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));  // MUFU.RCP64H, ~2^-20 relative
    return y;
}
#else
static inline double rcp_seed(double x) { return (double)(float)(1.0 / x); }
#endif

// 1/x to ~1 ulp for normal, finite x (two Newton steps on the MUFU seed)
__device__ __forceinline__ double rcp_fast(double x)
{
    double y = rcp_seed(x);
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// a/b for the Newton step of Kepler's equation: one Newton step on the MUFU seed gives y ~ 1/b to
// 2^-46 and the quotient a*y to 1.5e-14 relative.  That is below the rounding noise of the
// numerator itself (E - e sin E - M cancels to ~1e-16 absolute, i.e. >= 1e-14 relative to a step
// of 0.01) -- measured: logL moves by 1e-15 relative against the 1-ulp division (HB_DIV_EXACT).
// y is handed back: it seeds 1/den of the next, nearly identical, denominator.
__device__ __forceinline__ double div_fast(double a, double b, double& y_out)
{
    double y = rcp_seed(b);
    const double e = fma(-b, y, 1.0);
    y = fma(y, e, y);
    y_out = y;
    const double q = a * y;
#ifndef HB_DIV_EXACT
    return q;
#else
    const double r = fma(-b, q, a);
    return fma(r, y, q);
#endif
}
__device__ __forceinline__ double div_fast(double a, double b)
{
    double y;
    return div_fast(a, b, y);
}

// fdlibm __kernel_sin / __kernel_cos minimax coefficients on [-pi/4, pi/4]
__constant__ double kSinC[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                                2.75573137070700676789e-06,  -2.50507602534068634195e-08, 1.58969099521155010221e-10};
__constant__ double kCosC[6] = {4.16666666666666019037e-02,  -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                                -2.75573143513906633035e-07, 2.08757232129817482790e-09,  -1.13596475577881948265e-11};
// Scalar constants of the per-sample code as constant-bank operands (a literal double with a non-zero
// low word costs two moves per use): 0 fl(2 pi)  1 1/fl(2 pi)  4 E(M)-table nodes per radian
// 7 fl(pi)  (2, 3, 5, 6: unused since the magic number became an instruction immediate)
__constant__ double kMisc[8] = {6.283185307179586476925, 0.15915494309189534561, 0.0, 0.0,
                                122.23099629457562, 0.0, 0.0, 3.14159265358979323846};
// pi/2 split in three (Cody-Waite), 2/pi, and the round-to-integer magic number 1.5 * 2^52
__constant__ double kRed[5] = {1.57079632679489655800e+00, 6.12323399573676603587e-17, -1.49738490485916983880e-33,
                               6.36619772367581382433e-01, 6755399441055744.0};

// All per-sample routines below are written V samples wide (V = 1 or 2 per thread): the V
// dependency chains are independent, so the FP64 pipe (8-cycle DFMA latency, one warp
// instruction per 2 cycles per SM sub-partition, measured) always has a second chain to issue
// from, and the per-sample bookkeeping (constant loads, loop control) is shared.

static __device__ __noinline__ void sincos_library(double x, double* s, double* c) { sincos(x, s, c); }

// sin and cos, ~1 ulp, for |x| <= 1e5 (Kepler iterates are O(10)).  No range check here: the
// caller accumulates the largest exponent word it passed in (`hi_max`) and, in the rare case it
// exceeds kSincosHiLimit (wild Newton iterates at e -> 1, inf, NaN), redoes the sample with the
// library functions (kepler_point_careful).
constexpr int kSincosHiLimit = 0x40f86a00;  // high word of 1e5
template <int V>
__device__ __forceinline__ void sincos_lean(const double (&x)[V], double (&s_out)[V], double (&c_out)[V], int& hi_max)
{
#pragma unroll
    for (int j = 0; j < V; j++) {
        hi_max = max(hi_max, __double2hiint(x[j]) & 0x7fffffff);
        const double t = fma(x[j], kRed[3], kRed[4]);
        const int k = __double2loint(t);  // nearest integer to x * 2/pi sits in the low word
        const double kd = t - kRed[4];
        double r = fma(-kd, kRed[0], x[j]);
        r = fma(-kd, kRed[1], r);
        r = fma(-kd, kRed[2], r);
        const double z = r * r;
        double ps = fma(z, kSinC[5], kSinC[4]);
        double pc = fma(z, kCosC[5], kCosC[4]);
        ps = fma(z, ps, kSinC[3]);
        pc = fma(z, pc, kCosC[3]);
        ps = fma(z, ps, kSinC[2]);
        pc = fma(z, pc, kCosC[2]);
        ps = fma(z, ps, kSinC[1]);
        pc = fma(z, pc, kCosC[1]);
        ps = fma(z, ps, kSinC[0]);
        pc = fma(z, pc, kCosC[0]);
        const double sr = fma(z * r, ps, r);
        const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
        // quadrant: k odd swaps, bit 1 of k (sin) / of k+1 (cos) flips the sign
        const bool odd = k & 1;
        const double sa = odd ? cr : sr;
        const double ca = odd ? sr : cr;
        const int ssign = (k & 2) << 30;
        const int csign = ((k + 1) & 2) << 30;
        s_out[j] = __hiloint2double(__double2hiint(sa) ^ ssign, __double2loint(sa));
        c_out[j] = __hiloint2double(__double2hiint(ca) ^ csign, __double2loint(ca));
    }
}

// Table sincos for the hot loop: x = k h + r with h = 2 pi / 1024 (the Cody-Waite split of pi/2
// scaled by 2^-8, exact), |r| <= h/2 = 3.1e-3, so sin r and cos r need two Taylor terms each
// (next terms: r^6/5040 = 2e-19, r^6/720 = 1e-18 relative) and sin/cos(k h) come from a 1024-entry
// table of correctly rounded values (16 KB in global memory, L1-resident; filled once per context
// on the host in extended precision, hb_sincos_tab.h).  14 FP64 instructions and one 16-byte load
// against 22 + 11 quadrant instructions + 5 constant loads of sincos_lean; same ~1 ulp accuracy
// (sin is exact-relative near its zeros, which are table nodes).  Same contract as sincos_lean.
constexpr int kSinTabN = 1024;
// Doubles whose low 32 bits are zero are encoded as immediates in FP64 SASS instructions (no constant
// load, no register).  Used where 20 mantissa bits are enough:
//   kMagic   1.5 * 2^52, exact
//   kT5, kT4 1/120 and 1/24 to 20 bits: the Taylor terms they scale are below 1e-11 and 4e-12
// Validity range |x| <= 1024 (checked by the caller, who falls back to the library path beyond it,
// like sincos_lean does at 1e5): Kepler iterates are O(10), and the two-term reduction is exact to
// 1e-19 there.
constexpr int kSincosTabHiLimit = 0x40900000;  // high word of 1024.0
constexpr double kMagic = 6755399441055744.0;
constexpr double kT5 = 0x1.11111p-7;
constexpr double kT4 = 0x1.55555p-5;
// 0 nodes per radian  1, 2 h = pi/512 as (pi/2 hi, lo of kRed) / 256  3 -1/2  4 -1/6
// (an FMA takes one immediate at most: where two constants meet, one comes from here)
__constant__ double kTabC[6] = {162.97466172610082624, 1.57079632679489655800e+00 / 256.0, 6.12323399573676603587e-17 / 256.0,
                                -0.5, -1.0 / 6.0, 0.0};
template <int V, bool kTrack = true>
__device__ __forceinline__ void sincos_tab(const double (&x)[V], const double2* __restrict__ tab, double (&s_out)[V],
                                           double (&c_out)[V], int& hi_max)
{
#pragma unroll
    for (int j = 0; j < V; j++) {
        if (kTrack) hi_max = max(hi_max, __double2hiint(x[j]) & 0x7fffffff);
        const double t = fma(x[j], kTabC[0], kMagic);
        const int k = __double2loint(t) & (kSinTabN - 1);  // nearest node, periodic (two's complement for x < 0)
        const double kd = t - kMagic;
        double r = fma(-kd, kTabC[1], x[j]);
        r = fma(-kd, kTabC[2], r);
        HB_CHK(k, kSinTabN, 1);
        const double2 sc = tab[k];  // {sin, cos}(k h)
        const double z = r * r;
        const double sd = fma(r * z, fma(z, kT5, kTabC[4]), r);
        const double cd = fma(z, fma(z, kT4, kTabC[3]), 1.0);
        s_out[j] = fma(sc.y, sd, sc.x * cd);
        c_out[j] = fma(-sc.x, sd, sc.y * cd);
    }
}

// rarest tail of fmod_twopi: huge / inf / NaN input
static __device__ __noinline__ double fmod_twopi_lib(double am) { return fmod(am, kTwoPi); }

// Exact fmod(M, fl(2 pi)) keeping the dividend's sign (likelihood3.c:153): a rounded quotient
// from the magic-number trick and one exact FMA remainder; one range test catches both the
// off-by-one quotient and non-finite input.
// kNoCall (the hot pass): the huge / inf / NaN tail raises *hi_acc beyond every limit instead of going to the
// library -- the caller's deferred range check then has the chain evaluated again by the general pass.  A CALL on the
// sample's path, even one that is never taken, makes ptxas rebuild the shared-memory window address (uniform
// registers do not survive calls) for everything that follows it in the iteration.
template <bool kNoCall = false>
__device__ __forceinline__ double fmod_twopi(double M, int* hi_acc = nullptr)
{
    const double am = fabs(M);
    // floor(am / y): the magic-number trick with the FMA rounding toward -inf (at 1.5 * 2^52 one ulp is 1).
    // Adding `magic - 0.5` in round-to-nearest does NOT work: that constant is not representable.
    double q = __fma_rd(am, kMisc[1], kMagic) - kMagic;
    double r = fma(-q, kMisc[0], am);
    // (one integer test on the sample's path instead of two FP64 compares: the high word of r is at or above that of
    // 2 pi for r < 0, NaN and the last 2e-6 below 2 pi -- the exact test follows only then)
    if ((unsigned)__double2hiint(r) >= 0x401921fbu)
    if (!(r >= 0.0 && r < kMisc[0])) {
        // rare: quotient off by one (|M| within rounding of a multiple of 2 pi) -- repaired in line, so that the
        // sample loop does not marshal registers around a call at every sample -- or huge / inf / NaN input
        if (am < 1.0e15) {
            q += (r < 0.0) ? -1.0 : 1.0;
            r = fma(-q, kMisc[0], am);  // exact: 0 <= r < y is representable
        } else if (kNoCall) {
            *hi_acc = 0x7fffffff;
            r = 0.0;
        } else {
            r = fmod_twopi_lib(am);
        }
    }
    return copysign(r, M);
}

// Mean anomaly of likelihood3.c:149-153, bit-identical with the reference's 2 pi (t - T0) / P:
// the two products are explicitly rounded and the division is the correctly rounded Markstein
// sequence on rP = RN(1/P).  tsec = t * 86400 exactly as the reference forms it.
template <bool kNoCall = false>
__device__ __forceinline__ double mean_anomaly(double tsec, double T0s, double Ps, double rPs, int* hi_acc = nullptr)
{
    const double x = __dmul_rn(kMisc[0], __dsub_rn(tsec, T0s));
    const double q0 = __dmul_rn(x, rPs);
    return fmod_twopi<kNoCall>(fma(fma(-Ps, q0, x), rPs, q0), hi_acc);
}

// Starter of likelihood3.c:154-157: E0 = M + 0.85 e sign(sin M) (E0 = M when sin M == 0).
// For |M| < fl(2 pi), sin M > 0 on (0, fl(pi)] and < 0 above (sin(fl(pi)) = 1.2e-16 > 0), so the
// sign is read off M: sign(M) flipped when |M| > fl(pi), and no offset when M == 0.
__device__ __forceinline__ double kepler_starter(double m, double e)
{
    const int hi = __double2hiint(m), lo = __double2loint(m);
    const int ahi = hi & 0x7fffffff;
    const bool above_pi = (ahi > 0x400921fb) | ((ahi == 0x400921fb) & ((unsigned)lo > 0x54442d18u));
    const bool zero = (ahi | lo) == 0;
    const double off = 0.85 * e;
    int ohi = __double2hiint(off) ^ (hi & 0x80000000) ^ (above_pi ? 0x80000000 : 0);
    const double so = zero ? 0.0 : __hiloint2double(ohi, __double2loint(off));
    return m + so;
}

// Library-function version of the solve for one sample (cold path; exactly five steps).
static __device__ __noinline__ void kepler_point_careful(double m, double e, double* cE, double* sE)
{
    const double am = fabs(m);
    double sg = (am <= kPi) ? 1.0 : -1.0;
    sg = (m < 0.0) ? -sg : sg;
    sg = (am == 0.0) ? 0.0 : sg;
    double E = fma(0.85 * e, sg, m), s, c;
    for (int k = 0; k < 5; k++) {
        sincos(E, &s, &c);
        E -= (fma(-e, s, E) - m) / fma(-e, c, 1.0);
    }
    sincos(E, &s, &c);
    *cE = c;
    *sE = s;
}

// ---- per-chain E(M) table ----------------------------------------------------------------
// For 0 <= e <= kTableMaxE the reference's five Newton steps reach the root of Kepler's equation
// to rounding for every M (scanned on the CPU: max |E5 - E*| = 2.6e-15 at e = 0.8, 3.4e-15 at
// 0.85; the un-converged tail only starts between 0.85 and 0.90).  The converged root does not
// depend on the starter, so such chains may start Newton from anything convergent: the cubic
// Taylor polynomial of E(M) about the NEAREST of kTableN+1 uniform nodes of M on [0, 2 pi], whose
// four coefficients -- E, E1 h, E2 h^2/2, E3 h^3/6 with the derivatives E1 = beta = 1/(1 - e cos E),
// E2 = -e sin E beta^3, E3 = beta^4 (3 e^2 sin^2 E beta - e cos E) -- sit in 32 bytes of shared
// memory per node (two 16-byte loads, three FMAs; negative M uses E(-M) = -E(M), the upper half of
// the table is the mirror image E(2 pi - M) = 2 pi - E(M) of the solved lower half).  Truncation error
// (h/2)^4/24 E4 = 1.9e-10 E4: a third of what cubic Lagrange interpolation on 1024 intervals gave
// (rounds 1-2) at half its arithmetic, so the warp-uniform exit of the Newton loop fires after ONE step
// (two near periastron at the high end) instead of three to four from the reference starter.  Measured
// (tests/test_host_emul.py): |E0 - E| <= 1.8e-9 for every M at e <= 0.6 -- below the 2^-27 = 7.45e-9 of the
// exit test, i.e. one step everywhere --, 1.1e-8 at e = 0.7 (0.15 % of M take a second step), 1.2e-7 at
// e = 0.8 (1.2 %), 1.4e-7 at e = 0.95 outside the periastron window.
constexpr int kTableN = 768;                   // intervals on [0, 2 pi]  (h = 2 pi / 768)
constexpr int kTableNodes = kTableN + 1;       // nodes 0 .. kTableN
constexpr int kTableSolved = kTableN / 2 + 1;  // nodes 0 .. kTableN/2 are solved, the rest mirrored
constexpr int kTableMinN = 4108;               // light curves shorter than this do not pay for a table
// (kTableAllE, kTableMaxE, kTableMinM: see the top of this file)

// The two 16-byte halves {E, c1}, {c2, c3} of a node from the root E and sin/cos E there.
__device__ __forceinline__ void kepler_table_coeffs(double e, double E, double sE, double cE, double2& a, double2& b)
{
    constexpr double h = kTwoPi / (double)kTableN;
    const double beta = rcp_fast(fma(-e, cE, 1.0));
    const double es = e * sE, ec = e * cE, b2 = beta * beta;
    a.x = E;
    a.y = h * beta;
    b.x = (-0.5 * h * h) * (es * (b2 * beta));
    b.y = (h * h * h / 6.0) * ((b2 * b2) * fma(3.0 * es * es, beta, -ec));
}
// node n of the table from node kTableN - n
__device__ __forceinline__ void kepler_table_mirror(const double2& a, const double2& b, double2& am, double2& bm)
{
    am.x = kTwoPi - a.x;
    am.y = a.y;
    bm.x = -b.x;
    bm.y = b.y;
}

// starter from the table: tab[2 k], tab[2 k + 1] = node k (M = 2 pi k / kTableN), |m| < 2 pi.
// kClamp: m may be NaN (the general pass; the hot pass hands over finite |m| < 2 pi only).
template <bool kClamp = true>
__device__ __forceinline__ double kepler_table_guess(const double2* __restrict__ tab, double m)
{
    const double am = fabs(m);
    const double t = fma(am, kMisc[4], kMagic);  // the nearest node's number in the low word (at 1.5 * 2^52 one ulp is 1)
    int k = __double2loint(t);
    if (kClamp) k = (int)min((unsigned)k, (unsigned)kTableN);
    const double d = fma(am, kMisc[4], kMagic - t);  // distance from the node in node spacings, |d| <= 1/2
    HB_CHK(k, kTableNodes, 2);
    const double2 a = tab[2 * k], b = tab[2 * k + 1];
    const double E = fma(d, fma(d, fma(d, b.y, b.x), a.y), a.x);
    return copysign(E, m);
}

// One node for the host emulation (library-free, node by node).
__device__ __forceinline__ void kepler_table_node(int k, double e, double2& a, double2& b)
{
    const double m = (double)k * (kTwoPi / (double)kTableN);
    double E[1] = {kepler_starter(m, e)}, s[1], c[1];
    int hi = 0;
    for (int it = 0; it < 6; it++) {  // quadratic convergence: a starter only needs ~1e-10
        sincos_lean<1>(E, s, c, hi);
        E[0] -= div_fast(fma(-e, s[0], E[0]) - m, fma(-e, c[0], 1.0));
    }
    sincos_lean<1>(E, s, c, hi);
    kepler_table_coeffs(e, E[0], s[0], c[0], a, b);
}

// The whole table of one chain, built by the CTA (blockDim.x = kThreads threads, all of them call).  Nodes
// 0 .. N/2 are solved, two per thread as ONE interleaved latency chain (nodes tid and kThreads + tid); every
// thread also writes the mirror images of its nodes.
// Out of line: it runs once per chain and must not weigh on the register allocation of the sample loop.
#ifndef HB_HOST_EMUL  // (the host emulation fills its table node by node with kepler_table_node)
template <int kThreads>
static __device__ __noinline__ void build_kepler_table(double2* __restrict__ ktab, double e, const double2* __restrict__ sctab)
{
    static_assert(2 * kThreads >= kTableSolved, "two table nodes per thread must cover the solved half");
    const int tid = threadIdx.x;
    const int j[2] = {tid, min(kThreads + tid, kTableSolved - 1)};
    const double m[2] = {(double)j[0] * (kTwoPi / (double)kTableN), (double)j[1] * (kTwoPi / (double)kTableN)};
    double E[2] = {kepler_starter(m[0], e), kepler_starter(m[1], e)}, s[2], c[2], step[2] = {0.0, 0.0};
    int hi = 0;
    for (int k = 0; k < 6; k++) {  // quadratic convergence: a starter only needs ~1e-10
        // (0 <= M <= pi and e <= 0.99: Newton from the reference starter converges from above, every iterate stays
        // below 2 pi -- far inside the table sincos' range; the largest argument seen is not tracked)
        sincos_tab<2>(E, sctab, s, c, hi);
        bool small = true;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            step[i] = div_fast(fma(-e, s[i], E[i]) - m[i], fma(-e, c[i], 1.0));
            E[i] -= step[i];
            small &= fabs(step[i]) < 1e-6;  // the NEXT step is then ~C step^2 < 1e-11: this iterate is a starter already
        }
        // (warp-uniform: low eccentricities are there after three steps; every CTA of a chain builds the same table)
        if (__all_sync(0xffffffffu, small)) break;
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        // sin/cos of the last iterate to first order in the last step (1e-12 where the loop left early; nodes that
        // did not get there in six steps lie inside the periastron window of e -> 1, where the table is not used)
        const double cE = fma(s[i], step[i], c[i]), sE = fma(-c[i], step[i], s[i]);
        double2 a, b, am, bm;
        kepler_table_coeffs(e, E[i], sE, cE, a, b);
        kepler_table_mirror(a, b, am, bm);
        HB_CHK(j[i], kTableNodes, 3);
        HB_CHK(kTableN - j[i], kTableNodes, 3);
        ktab[2 * (kTableN - j[i])] = am;  // (node N/2 is its own image: the node itself is written last)
        ktab[2 * (kTableN - j[i]) + 1] = bm;
        ktab[2 * j[i]] = a;
        ktab[2 * j[i] + 1] = b;
    }
    __syncthreads();
}
#endif

// likelihood3.c:149-160 for V samples: outputs cos E, sin E, den = 1 - e cos E and beta = 1/den.
// kFullWarp: all 32 lanes of the warp execute this call together (true in the model pass).
// ktab: the chain's E(M) table (nullptr: reference starter, always valid).
// kSinTab: use sincos_tab with the table sctab (else the polynomial sincos_lean; sctab is ignored).
// kDeferRange: do not test the fast sincos' argument range per sample; the largest exponent word seen is
// accumulated into *hi_acc and the CALLER checks it once (and redoes its work without kDeferRange if it
// is out of range -- wild Newton iterates at e -> 1 are rare).
// kLowE: the caller knows that the chain takes the table starter at EVERY sample (ktab != nullptr, tab_min_m == 0,
// i.e. e <= kTableAllE): no starter choice per sample, and no tracking of the sincos arguments -- from the table
// starter every Newton iterate stays within a step of [-2 pi, 2 pi], far inside the table sincos' range.
template <int V, bool kFullWarp, bool kSinTab = false, bool kDeferRange = false, bool kLowE = false>
__device__ __forceinline__ void kepler_points(const double (&tsec)[V], const double e, const double T0s, const double Ps,
                                              const double rPs, const double2* __restrict__ ktab, const double tab_min_m,
                                              const double2* __restrict__ sctab, double (&cE)[V], double (&sE)[V],
                                              double (&den)[V], double (&beta)[V], int* hi_acc = nullptr,
                                              int flag_known = -1)
{
    double M[V], E[V], dE[V], yr[V];
    // (kDeferRange: bit 3 of the caller's register copy of the chain's flags says tab_min_m > 0 -- no load, no FP64
    // compare per sample for it)
    const bool window = kDeferRange ? ((flag_known & 8) != 0) : (tab_min_m > 0.0);
#pragma unroll
    for (int j = 0; j < V; j++) {
        M[j] = mean_anomaly<kDeferRange>(tsec[j], T0s, Ps, rPs, hi_acc);
        if (kLowE) {
            E[j] = kepler_table_guess<!kDeferRange>(ktab, M[j]);
        } else if (ktab == nullptr) {
            E[j] = kepler_starter(M[j], e);
        } else if (window) {  // eccentric chain: the reference's own path inside the periastron window
            const double am = fabs(M[j]);
            const bool far = fmin(am, kTwoPi - am) >= tab_min_m;
            E[j] = far ? kepler_table_guess<!kDeferRange>(ktab, M[j]) : kepler_starter(M[j], e);
        } else {
            E[j] = kepler_table_guess<!kDeferRange>(ktab, M[j]);
        }
    }
    // The reference always takes five Newton steps.  Once a step is below 2^-27 the next iterate
    // is the root to rounding (quadratic convergence: the following step is ~C step^2 < 1e-16) and
    // the reference's remaining steps only move E by rounding noise of f(E) -- the same noise any
    // other sin/cos implementation produces.  So the warp leaves the loop as soon as ALL its lanes
    // have such a step (warp-uniform, no divergence); lanes that never converge (the e -> 1 tail
    // near periastron) run all five steps exactly like the reference.
    bool tiny = false;
    int hi_max = 0;
#pragma unroll kNewtonUnroll
    for (int k = 0; k < 5; k++) {
        if (kSinTab) sincos_tab<V, !kLowE>(E, sctab, sE, cE, hi_max);
        else sincos_lean<V>(E, sE, cE, hi_max);
        tiny = true;
#pragma unroll
        for (int j = 0; j < V; j++) {
            const double num = fma(-e, sE[j], E[j]) - M[j];
            const double dn = fma(-e, cE[j], 1.0);
            dE[j] = div_fast(num, dn, yr[j]);  // the step: the next iterate is E - dE
            tiny &= (__double2hiint(dE[j]) & 0x7fffffff) < 0x3e400000;  // |dE| < 2^-27
        }
        tiny = warp_all<kFullWarp>(tiny);
        if (tiny) break;  // (E itself is not needed any more: sin/cos E follow from the step, below)
#pragma unroll
        for (int j = 0; j < V; j++) E[j] -= dE[j];
    }
    if (tiny) {
        // sin/cos(E_prev - dE) to first order (the neglected dE^2/2 < 3e-17 is relative), and
        // 1/den from the last reciprocal: den moved by < 1e-8 relative, one Newton step restores 1 ulp
#pragma unroll
        for (int j = 0; j < V; j++) {
            const double c0 = cE[j], s0 = sE[j];
            cE[j] = fma(s0, dE[j], c0);
            sE[j] = fma(-c0, dE[j], s0);
            den[j] = fma(-e, cE[j], 1.0);
            beta[j] = fma(yr[j], fma(-den[j], yr[j], 1.0), yr[j]);
        }
    } else {
        if (kSinTab) sincos_tab<V, !kLowE>(E, sctab, sE, cE, hi_max);
        else sincos_lean<V>(E, sE, cE, hi_max);
#pragma unroll
        for (int j = 0; j < V; j++) {
            den[j] = fma(-e, cE[j], 1.0);
            beta[j] = rcp_fast(den[j]);
        }
    }
    if (kLowE) {
    } else if (kDeferRange) {
        *hi_acc = max(*hi_acc, hi_max);
    } else if (hi_max > (kSinTab ? kSincosTabHiLimit : kSincosHiLimit)) {  // an iterate left the fast sincos' range: library
#pragma unroll
        for (int j = 0; j < V; j++) {
            kepler_point_careful(M[j], e, &cE[j], &sE[j]);
            den[j] = fma(-e, cE[j], 1.0);
            beta[j] = 1.0 / den[j];
        }
    }
}

// likelihood3.c:353-389 with R1 >= R2 already sorted and d already in Rsun.  Written with
// explicitly rounded (never FMA-contracted) operations in the reference's order: at the
// contact points (d == dc, d == R1 +- R2) h/R reaches 1 and a differently rounded h_sq
// would turn asin() into NaN where the reference is finite (quirk Q10).  Only in-eclipse
// samples come here, so the few extra instructions are free.
// kGuard (the model pass): next to d == dc = sqrt(R1^2 - R2^2), h^2 comes out of a cancelling difference of numbers
// ~4 d^2 R1^2 and its rounding noise can push h/R2 above 1 or R2^2 - h^2 below 0 -- by 8e-12 at R1/R2 = 300,
// anywhere within 1.4e-8 dc of the contact.  Whether the reference returns NaN there is decided by the last bit of ITS
// separation (libm tan/atan/sin/cos against the algebraic form here), i.e. it cannot be followed: a seeded scan met one
// chain in 21 000 where this kernel had the NaN and the reference had not.  With the guard the model pass takes
// the formula's limit (asin 1, sqrt 0) wherever only rounding left its domain; genuine NaN input still propagates
// (every test is false on NaN), and the stand-alone entry point (hb_scalar op 5) keeps the reference's bits.
template <bool kGuard = false>
static __device__ __noinline__ double eclipse_area_dev(double R1, double R2, double d)
{
    const double R1s = __dmul_rn(R1, R1), R2s = __dmul_rn(R2, R2);
    double area = 0.;
    const double dc = sqrt(__dsub_rn(R1s, R2s));
    const double sum = __dadd_rn(R1, R2), dif = __dsub_rn(R1, R2);
    const double full = __dmul_rn(__dmul_rn(kPi, R2), R2);
    if (d < dif) area = full;
    const bool partial_out = (d > dc) & (d < sum);
    const bool partial_in = (d <= dc) & (d >= dif);
    if (partial_out | partial_in) {
        const double dd = __dmul_rn(d, d);
        const double four_dd = __dmul_rn(__dmul_rn(4., d), d);
        const double a = __dmul_rn(__dmul_rn(four_dd, R1), R1);                 // 4 d d R1 R1
        const double b = __dadd_rn(__dsub_rn(dd, R2s), R1s);                    // d d - R2 R2 + R1 R1
        double h_sq = __ddiv_rn(__dsub_rn(a, __dmul_rn(b, b)), four_dd);
        if (kGuard && h_sq < 0.0) h_sq = 0.0;
        const double h = sqrt(h_sq);
        const double hh = __dmul_rn(h, h);
        double q1 = __ddiv_rn(h, R1), q2 = __ddiv_rn(h, R2), r1 = __dsub_rn(R1s, hh), r2 = __dsub_rn(R2s, hh);
        if (kGuard) {
            q1 = q1 > 1.0 ? 1.0 : q1;
            q2 = q2 > 1.0 ? 1.0 : q2;
            r1 = r1 < 0.0 ? 0.0 : r1;
            r2 = r2 < 0.0 ? 0.0 : r2;
        }
        const double A1 = __dsub_rn(__dmul_rn(R1s, asin(q1)), __dmul_rn(h, sqrt(r1)));
        const double A2 = __dsub_rn(__dmul_rn(R2s, asin(q2)), __dmul_rn(h, sqrt(r2)));
        area = partial_out ? __dadd_rn(A1, A2) : __dsub_rn(full, __dadd_rn(-A1, A2));
    }
    return area;
}

// Raw (un-normalised) template values Amag1 + Amag2 of likelihood3.c:649-675 at V samples
// (tsec = t * 86400, formed once per data set).
template <int V, bool kFullWarp, bool kSinTab = false, bool kDeferRange = false, bool kLowE = false>
__device__ __forceinline__ void raw_flux(const ChainConst& cc, const double2* __restrict__ ktab,
                                         const double2* __restrict__ sctab, const double (&tsec)[V], double (&u)[V],
                                         int* hi_acc = nullptr, int flag_known = -1)
{
    // (flag_known in the kDeferRange pass: the caller's register copy of cc.flag, bit 3 = tab_min_m > 0 added -- a
    // loop that stores to shared memory would otherwise re-read and re-convert the flag at every sample)
    const bool may_eclipse = ((kDeferRange ? flag_known : (int)cc.flag) & 4) == 0;
    double cE[V], sE[V], den[V], bet[V];
    kepler_points<V, kFullWarp, kSinTab, kDeferRange, kLowE>(tsec, cc.e, cc.T0s, cc.Ps, cc.rPs, ktab, cc.tab_min_m, sctab, cE, sE, den, bet, hi_acc, flag_known);
#pragma unroll
    for (int j = 0; j < V; j++) {
        const double beta = bet[j];  // (1 + e cos nu)/(1 - e^2) == 1/(1 - e cos E)
        // cos nu = (cos E - e) beta, sin nu = sqrt(1 - e^2) sin E beta; rotate by omega0, then scale
        const double X = cE[j] - cc.e;
        const double c = fma(-cc.swq, sE[j], cc.cw * X) * beta;  // cos(omega0 + nu)
        const double s = fma(cc.cwq, sE[j], cc.sw * X) * beta;   // sin(omega0 + nu)
        const double c2 = fma(c, c, -0.5);                       // h = cos 2x / 2 (the factor 2 sits in the coefficients)

        const double P5 = fma(fma(cc.r4, c2, cc.d2), c2, cc.r0);
        const double P4 = s * fma(cc.q3, c2, cc.q1);
        const double P3 = fma(cc.b1, c2, cc.b0);
        const double P2 = fma(cc.a2, c2, fma(cc.a1, s, cc.a0));
        const double poly = fma(beta, fma(beta, fma(beta, P5, P4), P3), P2);
        const double bb = beta * beta;
        double uj = fma(bb, poly, fma(cc.K1, c, cc.K0));

        // eclipse: the squared projected separation over a^2 against the chain's threshold (s^2 = 1 - c^2 to
        // rounding; the 1e-9 slack of thr absorbs that); the rare in-eclipse sample then forms the
        // separation exactly as the reference does (likelihood3.c:173-176, 365)
        // (den^2 g < thr as g < thr beta^2: beta^2 is there already, and den is not kept alive for the test -- the 1e-9
        // slack of thr is seven orders of magnitude above what the two forms differ by)
        if (may_eclipse && fma(c2, cc.si2, cc.ci2) < cc.thr * bb) {
            const double sc = s * cc.ci;
            const double proj2 = fma(c, c, sc * sc);
            const double rr = cc.ar * den[j];
            const double lim = cc.Rb + cc.Rs;
            const double d = fabs(rr * sqrt(proj2));
            if (!(d >= lim)) {
                const double area = eclipse_area_dev<true>(cc.Rb, cc.Rs, d);
                const double zz = s * cc.si;  // sign of ZZ (likelihood3.c:173,669-670)
                if (zz < 0.0) uj -= area * cc.ecl2;
                else if (zz > 0.0) uj -= area * cc.ecl1;
            }
        }
        u[j] = uj;
    }
}

template <bool kFullWarp, bool kSinTab = false>
__device__ __forceinline__ double raw_flux1(const ChainConst& cc, const double2* __restrict__ ktab,
                                            const double2* __restrict__ sctab, double tsec)
{
    const double t[1] = {tsec};
    double u[1];
    raw_flux<1, kFullWarp, kSinTab>(cc, ktab, sctab, t, u);
    return u[0];
}

// likelihood3.c:681-685 in the reference's order: ((u - med) + 1) blended, times flux_tune
__device__ __forceinline__ double finish_template(double u, double med, double blend, double ft)
{
    double v = __dadd_rn(__dsub_rn(u, med), 1.0);
    return __dmul_rn(__dadd_rn(blend, __dmul_rn(v, __dsub_rn(1.0, blend))), ft);
}

}  // namespace hb
