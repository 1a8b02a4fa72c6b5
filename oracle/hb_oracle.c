/*
 * hb_oracle.c -- CPU restatement of the HB_MCMC hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker for the CUDA path in hb_mcmc_b200/csrc.  It is
 * imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  The product (libhb_b200.so, the likelihood3 shim, the PT
 * driver) never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED.  In the build container the functions below are compared
 * bit-for-bit with the unmodified reference compiled from /root/reference/src into
 * oracle/_ref/ (tests/test_oracle_vs_ref.py), and on any box against the golden
 * vectors in tests/golden/ that were generated from that compiled reference
 * (tests/golden/make_golden.py).
 *
 * The arithmetic follows the reference's operation ORDER exactly (ISO C, no FMA
 * contraction: build with -std=c99 -O3, no -ffast-math / -march=native), so that the
 * results are bit-identical with the reference on the same libm.  Every function
 * cites the reference location (paths relative to /root/reference/src) it restates.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "hb_oracle.h"

/* physical constants, likelihood3.h:4-10,31 */
static const double kPi = 3.14159265358979323846;
static const double kG = 6.6743e-8;
static const double kC = 2.998e10;
static const double kMsun = 1.9885e33;
static const double kRsun = 6.955e10;
static const double kSecDay = 86400.0;
static const double kBig = 1.e15;

static inline double sq(double x) { return x * x; }
static inline double cube(double x) { return x * x * x; }
static inline double quad(double x) { return x * x * x * x; }

/* ------------------------------------------------------------------ */
/* stellar relations                                                   */
/* ------------------------------------------------------------------ */

/* piecewise-linear table in LINEAR mass, clamped at both ends.
 * likelihood3.c:396-438 (_getT) and :445-476 (_getR) share this shape. */
static double interp_mass_table(double logM, const double *m_nodes, const double *y_nodes, int n)
{
    double m = pow(10., logM);
    if (m <= m_nodes[0]) return y_nodes[0];
    if (m >= m_nodes[n - 1]) return y_nodes[n - 1];
    for (int j = 1; j < n; j++) {
        if (m < m_nodes[j]) {
            return y_nodes[j - 1] +
                   (m - m_nodes[j - 1]) * (y_nodes[j] - y_nodes[j - 1]) / (m_nodes[j] - m_nodes[j - 1]);
        }
    }
    return y_nodes[n - 1]; /* unreachable */
}

/* likelihood3.c:396-438 */
double orc_getT(double logM)
{
    static const double m_nodes[16] = {0.1, 0.26, 0.47, 0.59, 0.69, 0.87, 0.98, 1.085,
                                       1.4, 1.65, 2.0, 2.5, 3.0, 4.4, 15., 40.};
    static const double t_nodes[16] = {3.491, 3.531, 3.547, 3.584, 3.644, 3.712, 3.745, 3.774,
                                       3.823, 3.863, 3.913, 3.991, 4.057, 4.182, 4.477, 4.623};
    return interp_mass_table(logM, m_nodes, t_nodes, 16);
}

/* likelihood3.c:445-476 (node 0.784 -> 0.787 is non-monotone on purpose) */
double orc_getR(double logM)
{
    static const double m_nodes[10] = {0.07, 0.2, 0.356, 0.655, 0.784, 0.787, 1.377, 4.4, 15., 40.};
    static const double r_nodes[10] = {-0.953, -0.627, -0.423, -0.154, -0.082,
                                       -0.087, 0.295,  0.477,  0.792,  1.041};
    return interp_mass_table(logM, m_nodes, r_nodes, 10);
}

/* likelihood3.c:483-493 */
double orc_envelope_temp(double logM)
{
    (void)logM;
    return 0.0224;
}

/* likelihood3.c:495-507 */
double orc_envelope_radius(double logM)
{
    const double n = 4.22, slope = 15.68, floor_ = 0.01, corner = 1.055, ceil_ = 0.17;
    double m = pow(10., logM);
    return 1 / (1 / ceil_ + 1 / (slope * pow((pow(m, n) + pow(corner, n)), (1 / n)) - (slope * corner - floor_)));
}

/* likelihood3.c:693-717 */
void orc_radii_teffs(const double *p, double *R1, double *R2, double *T1, double *T2)
{
    *R1 = pow(10., orc_getR(p[0]) + p[7] * orc_envelope_radius(p[0]));
    *R2 = pow(10., orc_getR(p[1]) + p[8] * orc_envelope_radius(p[1]));
    *T1 = pow(10., orc_getT(p[0]) + p[17] * orc_envelope_temp(p[0]));
    *T2 = pow(10., orc_getT(p[1]) + p[18] * orc_envelope_temp(p[1]));
}

/* likelihood3.c:194-209 */
double orc_alpha_beam(double logT)
{
    static const double alphas[4] = {6.5, 4.0, 2.5, 1.2};
    static const double logTs[4] = {3.5, 3.7, 3.9, 4.5};
    if (logT >= logTs[3]) return 1.2 / 4;
    if (logT < logTs[0]) return 6.5 / 4;
    int j = 3;
    while (logT < logTs[j]) j--;
    return ((alphas[j + 1] + (alphas[j + 1] - alphas[j]) / (logTs[j + 1] - logTs[j]) * (logT - logTs[j + 1])) / 4);
}

/* ------------------------------------------------------------------ */
/* orbit                                                               */
/* ------------------------------------------------------------------ */

/* likelihood3.c:125-185.  tp = {M1,M2 [g], P [s], e, inc, omega0, T0 [s]} */
void orc_traj(const double *times, const double *tp, double *d_arr, double *Z1_arr, double *Z2_arr,
              double *rr_arr, double *ff_arr, long Nt)
{
    double Ma = tp[0], Mb = tp[1];
    const double P = tp[2], e = tp[3], inc = tp[4], omega0 = tp[5], T0 = tp[6];
    if (Mb > Ma) { double s = Ma; Ma = Mb; Mb = s; }
    const double Mtot = Ma + Mb;
    const double a = pow(kG * Mtot * sq(P) / sq(2 * kPi), 1. / 3.);

    for (long i = 0; i < Nt; i++) {
        double tsec = times[i] * kSecDay;
        double M = 2. * kPi * (tsec - T0) / P;
        M = fmod(M, 2 * kPi);
        double E = M;
        double sM = sin(M);
        if (sM != 0.0) E = M + 0.85 * e * sM / fabs(sM);
        /* exactly five Newton steps, likelihood3.c:160 */
        for (int k = 0; k < 5; k++) E = E - (E - e * sin(E) - M) / (1 - e * cos(E));

        rr_arr[i] = a * (1 - e * cos(E));
        ff_arr[i] = 2. * atan(sqrt((1. + e) / (1. - e)) * tan(E / 2.));

        double cwf = cos(omega0 + ff_arr[i]);
        double swf = sin(omega0 + ff_arr[i]);
        double ci = cos(inc);
        double si = sin(inc);

        double ZZ = rr_arr[i] * swf * si;
        double dfac = sqrt(sq(cwf) + sq(swf * ci));
        d_arr[i] = rr_arr[i] * dfac;
        Z1_arr[i] = ZZ * (Mb / Mtot);
        Z2_arr[i] = -ZZ * (Ma / Mtot);
    }
}

/* ------------------------------------------------------------------ */
/* flux terms                                                          */
/* ------------------------------------------------------------------ */

/* likelihood3.c:224-236.  Quirk Q1: the exponent 2/3 is integer division = 0. */
double orc_beaming(double P, double M1, double M2, double e, double inc, double omega0, double nu,
                   double alpha_beam)
{
    double q = M2 / M1;
    double f1 = q / pow(1 + q, 2 / 3);
    double f2 = pow(M1, 1. / 3);
    double f3 = pow(P, -1. / 3);
    double f4 = sin(inc) * cos(omega0 + nu) / sqrt(1 - sq(e));
    return -2830. * alpha_beam * f1 * f2 * f3 * f4 * 1.e-6;
}

/* likelihood3.c:255-307 */
double orc_ellipsoidal(double P, double M1, double M2, double e, double inc, double omega0, double nu,
                       double R1, double a, double mu, double tau)
{
    (void)a;
    double al11 = 15 * mu * (2 + tau) / (32 * (3 - mu));
    double al21 = 3 * (15 + mu) * (1 + tau) / (20 * (3 - mu));
    double al2b1 = 15 * (1 - mu) * (3 + tau) / (64 * (3 - mu));
    double al01 = al21 / 9;
    double al0b1 = 3 * al2b1 / 20;
    double al31 = 5 * al11 / 3;
    double al41 = 7 * al2b1 / 4;

    double beta = (1 + e * cos(nu)) / (1 - sq(e));
    double q = M2 / M1;
    double Prot = P * pow(1 - e, 3. / 2);
    const double ppm = 1.e-6;
    double tot = 0.;

    double AM1 = 13435 * 2 * al01 * (2 - 3 * sq(sin(inc))) * (1 / M1) * (1 / sq(Prot)) * cube(R1);
    double AM2 = 13435 * 3 * al01 * (2 - 3 * sq(sin(inc))) * (1 / M1) * q / (1 + q) * (1 / sq(P)) * cube(beta * R1);
    double C2_1 = 13435 * al21 * sq(sin(inc)) * (1 / M1) * q / (1 + q) * (1 / sq(P)) * cube(beta * R1) *
                  cos(2 * (omega0 + nu));
    tot += (AM1 + AM2 + C2_1) * ppm;

    double AM3 = 759 * al0b1 * (8 - 40 * sq(sin(inc)) + 35 * quad(sin(inc))) * pow(M1, -5. / 3) * q /
                 pow(1 + q, 5. / 3) * pow(P, -10. / 3) * pow(beta * R1, 5);
    double S1 = 3194 * al11 * (4 * sin(inc) - 5 * cube(sin(inc))) * pow(M1, -4. / 3) * q / pow(1 + q, 4. / 3) *
                pow(P, -8. / 3) * quad(beta * R1) * sin(omega0 + nu);
    double C2_2 = 759 * al2b1 * (6 * sq(sin(inc)) - 7 * quad(sin(inc))) * pow(M1, -5. / 3) * q /
                  pow(1 + q, 5. / 3) * pow(P, -10. / 3) * (beta * R1) * quad(beta * R1) * cos(2 * (omega0 + nu));
    double S3 = 3194 * al31 * cube(sin(inc)) * pow(M1, -4. / 3) * q / pow(1 + q, 4. / 3) * pow(P, -8. / 3) *
                quad(beta * R1) * sin(3 * (omega0 + nu));
    double C4 = 759 * al41 * quad(sin(inc)) * pow(M1, -5. / 3) * q / pow(1 + q, 5. / 3) * pow(P, -10. / 3) *
                (beta * R1) * quad(beta * R1) * cos(4 * (omega0 + nu));
    tot += (AM3 + S1 + C2_2 + S3 + C4) * ppm;
    return tot;
}

/* likelihood3.c:322-337 */
double orc_reflection(double P, double M1, double M2, double e, double inc, double omega0, double nu,
                      double R2, double alpha_ref)
{
    double q = M2 / M1;
    double beta = (1 + e * cos(nu)) / (1 - sq(e));
    double f1 = pow(1 + q, -2. / 3);
    double f2 = pow(M1, -2. / 3);
    double f3 = pow(P, -4. / 3);
    double f4 = sq(beta * R2);
    double f5 = 0.64 - sin(inc) * sin(omega0 + nu) + 0.18 * sq(sin(inc)) * (1 - cos(2 * (omega0 + nu)));
    return 56514 * alpha_ref * f1 * f2 * f3 * f4 * f5 * 1.e-6;
}

/* likelihood3.c:353-389.  Radii in Rsun, d in cm. */
double orc_eclipse_area(double R1, double R2, double d)
{
    if (R2 > R1) { double s = R1; R1 = R2; R2 = s; }
    double area = 0.;
    d = fabs(d) / kRsun;
    double dc = sqrt(R1 * R1 - R2 * R2);
    if (d >= (R1 + R2)) area = 0.;
    if (d < (R1 - R2)) area = kPi * R2 * R2;
    if ((d > dc) & (d < (R1 + R2))) {
        double h_sq = (4. * d * d * R1 * R1 - sq(d * d - R2 * R2 + R1 * R1)) / (4. * d * d);
        double h = sqrt(h_sq);
        double A1 = R1 * R1 * asin(h / R1) - h * sqrt(R1 * R1 - h * h);
        double A2 = R2 * R2 * asin(h / R2) - h * sqrt(R2 * R2 - h * h);
        area = A1 + A2;
    }
    if ((d <= dc) & (d >= (R1 - R2))) {
        double h_sq = (4. * d * d * R1 * R1 - sq(d * d - R2 * R2 + R1 * R1)) / (4. * d * d);
        double h = sqrt(h_sq);
        double A1 = R1 * R1 * asin(h / R1) - h * sqrt(R1 * R1 - h * h);
        double A2 = R2 * R2 * asin(h / R2) - h * sqrt(R2 * R2 - h * h);
        area = kPi * R2 * R2 - (-A1 + A2);
    }
    return area;
}

/* ------------------------------------------------------------------ */
/* median                                                              */
/* ------------------------------------------------------------------ */

static int cmp_dbl(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* Index rule of likelihood3.c:97-101 (quirk Q3): even N -> N/2, odd N -> N/2+1. */
long orc_median_rank(long N) { return (N % 2 == 0) ? N / 2 : N / 2 + 1; }

/* likelihood3.c:86-105.  The reference sorts a copy with a Lomuto quicksort; for
 * NaN-free input every correct sort gives the same order statistic, so libc qsort is
 * used here (heap copy instead of a stack VLA: no segfault at Nt = 200k, quirk Q12). */
void orc_remove_median(double *arr, long begin, long end)
{
    long N = end - begin;
    if (N <= 0) return;
    double *tmp = (double *)malloc((size_t)N * sizeof(double));
    memcpy(tmp, arr + begin, (size_t)N * sizeof(double));
    qsort(tmp, (size_t)N, sizeof(double), cmp_dbl);
    long mid = orc_median_rank(N);
    /* the reference reads sorted[N/2+1] for odd N, which is one past the end for N == 1 */
    double med = (mid < N) ? tmp[mid] : tmp[N - 1];
    free(tmp);
    for (long i = 0; i < N; i++) arr[begin + i] -= med;
}

/* ------------------------------------------------------------------ */
/* light curve                                                         */
/* ------------------------------------------------------------------ */

/* likelihood3.c:530-686.  `raw` (optional, may be NULL) receives the un-normalised
 * template Amag1+Amag2 of :673, which the GPU tests use to check the median stage. */
void orc_calc_light_curve_ex(const double *times, long Nt, const double *pars, double *tmpl, double *raw)
{
    const double logM1 = pars[0], logM2 = pars[1];
    const double P = pow(10., pars[2]) * kSecDay;
    const double Pdays = pow(10., pars[2]);
    const double e = pars[3], inc = pars[4], omega0 = pars[5], T0 = pars[6];
    const double mu1 = pars[9], tau1 = pars[10], mu2 = pars[11], tau2 = pars[12];
    const double aref1 = pars[13], aref2 = pars[14];
    const double xbeam1 = exp(pars[15]), xbeam2 = exp(pars[16]);
    const double blending = pars[19], flux_tune = pars[20];

    const double M1 = pow(10., logM1), M2 = pow(10., logM2);
    double tp[7] = {M1 * kMsun, M2 * kMsun, P, e, inc, omega0, T0 * kSecDay};

    double R1 = 0., R2 = 0., Te1 = 0., Te2 = 0.;
    orc_radii_teffs(pars, &R1, &R2, &Te1, &Te2);

    double Norm1 = sq(R1) * quad(Te1) / (sq(R1) * quad(Te1) + sq(R2) * quad(Te2));
    double Norm2 = sq(R2) * quad(Te2) / (sq(R1) * quad(Te1) + sq(R2) * quad(Te2));

    double ab1 = orc_alpha_beam(log10(Te1));
    double ab2 = orc_alpha_beam(log10(Te2));
    ab1 *= xbeam1;
    ab2 *= xbeam2;

    double Mtot = (M1 + M2) * kMsun;
    double a = pow(kG * Mtot * P * P / (4.0 * kPi * kPi), 1. / 3.);
    double ar = a / kRsun;

    double *buf = (double *)malloc((size_t)(5 * (Nt > 0 ? Nt : 1)) * sizeof(double));
    double *d_arr = buf, *Z1 = buf + Nt, *Z2 = buf + 2 * Nt, *r_arr = buf + 3 * Nt, *nu = buf + 4 * Nt;
    orc_traj(times, tp, d_arr, Z1, Z2, r_arr, nu, Nt);

    for (long i = 0; i < Nt; i++) {
        double b1 = orc_beaming(Pdays, M1, M2, e, inc, omega0, nu[i], ab1);
        double e1 = orc_ellipsoidal(Pdays, M1, M2, e, inc, omega0, nu[i], R1, ar, mu1, tau1);
        double r1 = orc_reflection(Pdays, M1, M2, e, inc, omega0, nu[i], R2, aref1);
        double A1 = Norm1 * (1 + b1 + e1 + r1);

        double b2 = orc_beaming(Pdays, M2, M1, e, inc, (omega0 + kPi), nu[i], ab2);
        double e2 = orc_ellipsoidal(Pdays, M2, M1, e, inc, (omega0 + kPi), nu[i], R2, ar, mu2, tau2);
        double r2 = orc_reflection(Pdays, M2, M1, e, inc, (omega0 + kPi), nu[i], R1, aref2);
        double A2 = Norm2 * (1 + b2 + e2 + r2);

        double area = orc_eclipse_area(R1, R2, d_arr[i]);
        if (Z2[i] > Z1[i]) A2 -= area * Norm2 / (kPi * sq(R2));
        else if (Z2[i] < Z1[i]) A1 -= area * Norm1 / (kPi * sq(R1));

        tmpl[i] = (A1 + A2);
    }
    free(buf);
    if (raw) memcpy(raw, tmpl, (size_t)Nt * sizeof(double));

    orc_remove_median(tmpl, 0, Nt);
    for (long i = 0; i < Nt; i++) {
        tmpl[i] += 1;
        tmpl[i] = (1 * blending + tmpl[i] * (1 - blending)) * flux_tune;
    }
}

void orc_calc_light_curve(const double *times, long Nt, const double *pars, double *tmpl)
{
    orc_calc_light_curve_ex(times, Nt, pars, tmpl, NULL);
}

/* ------------------------------------------------------------------ */
/* magnitudes, Roche, likelihood                                       */
/* ------------------------------------------------------------------ */

/* Two-blackbody AB magnitudes shared by likelihood3.c:725-795 (flavour 0: exp()-1 and
 * the /(1-blending) correction) and GAIA_mcmc.c:198-250 (flavour 1: expm1, no blending). */
static void two_bb_mags(double R1, double R2, double T1, double T2, double D, double blending, int gaia,
                        double out[4])
{
    static const double lam[4] = {442, 540, 673, 750};
    const double h = 6.626e-27, k = 1.38e-16, pc = 3.086e18;
    double f[4];
    R1 *= kRsun;
    R2 *= kRsun;
    for (int j = 0; j < 4; j++) {
        double nu = kC / (lam[j] * 1e-7);
        if (!gaia) {
            f[j] = kPi * (R1 * R1 * (2. * h * cube(nu) / sq(kC) / (exp(h * nu / (k * T1)) - 1.)) +
                          R2 * R2 * (2. * h * cube(nu) / sq(kC) / (exp(h * nu / (k * T2)) - 1.))) /
                   (sq(D) * sq(pc));
            f[j] = f[j] / (1 - blending);
        } else {
            f[j] = kPi * (sq(R1) * (2. * h * cube(nu) / sq(kC) / expm1(h * nu / (k * T1))) +
                          sq(R2) * (2. * h * cube(nu) / sq(kC) / expm1(h * nu / (k * T2)))) /
                   (sq(D) * sq(pc));
        }
    }
    double B = -2.5 * log10(f[0]) - 48.6;
    double V = -2.5 * log10(f[1]) - 48.6;
    double Gm = -2.5 * log10(f[2]) - 48.6;
    double T = -2.5 * log10(f[3]) - 48.6;
    out[0] = Gm;
    out[1] = B - V;
    out[2] = V - Gm;
    out[3] = Gm - T;
}

/* likelihood3.c:725-795; out = {G, B-V, V-G, G-T} */
void orc_calc_mags(const double *p, double D, double out[4])
{
    double R1, R2, T1, T2;
    orc_radii_teffs(p, &R1, &R2, &T1, &T2);
    two_bb_mags(R1, R2, T1, T2, D, p[19], 0, out);
}

/* GAIA_mcmc.c:198-250; 6-parameter layout {logM1, logM2, rr1, rr2, aT1, aT2} */
void orc_gaia_get_mags(const double *p6, double D, double out[4])
{
    double R1 = pow(10., orc_getR(p6[0]) + p6[2] * orc_envelope_radius(p6[0]));
    double R2 = pow(10., orc_getR(p6[1]) + p6[3] * orc_envelope_radius(p6[1]));
    double T1 = pow(10., orc_getT(p6[0]) + p6[4] * orc_envelope_temp(p6[0]));
    double T2 = pow(10., orc_getT(p6[1]) + p6[5] * orc_envelope_temp(p6[1]));
    two_bb_mags(R1, R2, T1, T2, D, 0., 1, out);
}

/* GAIA_mcmc.c:255-269; data/err/model have 4 entries */
double orc_gaia_model_likelihood(const double *data, const double *err, const double *p6, double D)
{
    double model[4];
    orc_gaia_get_mags(p6, D, model);
    double chi2 = 0.;
    for (int i = 0; i < 4; i++) {
        double r = (data[i] - model[i]) / err[i];
        chi2 += r * r;
    }
    return (-chi2 / 2.0);
}

/* likelihood3.c:945-948 */
static double eggleton(double q)
{
    return 0.49 * pow(q, 2. / 3) / (0.6 * pow(q, 2. / 3) + log(1 + pow(q, 1. / 3)));
}

/* likelihood3.c:953-974 */
int orc_roche_overflow(const double *p)
{
    double M1 = pow(10., p[0]) * kMsun;
    double M2 = pow(10., p[1]) * kMsun;
    double q = M1 / M2;
    double period = pow(10., p[2]) * kSecDay;
    double ecc = p[3];
    double R1 = pow(10., orc_getR(p[0]) + p[7] * orc_envelope_radius(p[0])) * kRsun;
    double R2 = pow(10., orc_getR(p[1]) + p[8] * orc_envelope_radius(p[1])) * kRsun;
    double sep = pow(kG * (M1 + M2) * sq(period) / (4.0 * kPi * kPi), 1. / 3.);
    double RL1 = eggleton(q);
    double RL2 = eggleton(1 / q);
    double r1 = R1 / (sep * (1 - ecc));
    double r2 = R2 / (sep * (1 - ecc));
    return ((RL1 < r1) || (RL2 < r2)) ? 1 : 0;
}

/* likelihood3.c:809-873 with USE_GMAG / USE_COLOR_INFO (likelihood3.h:11-12) made
 * runtime flags.  Like the reference it clamps noise[] IN PLACE (quirk Q2). */
double orc_loglikelihood(const double *time, const double *flux, double *noise, long N, const double *params,
                         const double *mag_data, const double *magerr, int use_gmag, int use_color)
{
    double *tmpl = (double *)malloc((size_t)(N > 0 ? N : 1) * sizeof(double));
    orc_calc_light_curve(time, N, params, tmpl);

    double chi2 = 0.;
    for (long i = 0; i < N; i++) {
        if (noise[i] < 1.e-5) noise[i] = 1.e-5;
        double r = (tmpl[i] - flux[i]) / noise[i];
        chi2 += r * r;
    }
    free(tmpl);

    if (use_color || use_gmag) {
        double m[4];
        orc_calc_mags(params, mag_data[0], m);
        if (use_gmag) {
            double r = (m[0] - mag_data[1]) / magerr[0];
            chi2 += r * r;
        }
        if (use_color) {
            for (int i = 1; i < 4; i++) {
                double r = (m[i] - mag_data[i + 1]) / magerr[i];
                chi2 += r * r;
            }
        }
    }
    if (orc_roche_overflow(params)) chi2 = kBig;
    return (-chi2 / 2.0);
}

/* ------------------------------------------------------------------ */
/* sampler pieces (deterministic parts only)                           */
/* ------------------------------------------------------------------ */

/* likelihood3.c:986-1121.  mode: 1 = reflect, 2 = periodic, anything else = unbounded
 * (quirk Q4: the upper mode of e is 0.99 and its upper limit 1).  Arrays of ORC_NPARS. */
void orc_set_limits(double *lo, double *hi, double *mode_lo, double *mode_hi, int *gauss, double lc_period)
{
    static const double L[ORC_NPARS][2] = {
        {-1.5, 2.0}, {-1.5, 2.0}, {-2.0, 3.0}, {0.0, 1.0},   {0.0, 0.0},  {0.0, 0.0},   {0.0, 0.0},
        {-5., 5.},   {-5., 5.},   {0.12, 0.20}, {0.3, 0.38}, {0.12, 0.20}, {0.3, 0.38}, {0.5, 1.5},
        {0.5, 1.5},  {-0.3, 0.3}, {-0.3, 0.3},  {-5., 5.},   {-5., 5.},    {0., 1.},    {0.99, 1.01}};
    for (int i = 0; i < ORC_NPARS; i++) {
        lo[i] = L[i][0];
        hi[i] = L[i][1];
        mode_lo[i] = 1;
        mode_hi[i] = 1;
        gauss[i] = (i >= 7 && i <= 18) ? 1 : 0;
    }
    mode_hi[3] = 0.99;
    lo[4] = 0;
    hi[4] = kPi;
    lo[5] = -kPi;
    hi[5] = kPi;
    mode_lo[5] = 2;
    mode_hi[5] = 2;
    lo[6] = 0.;
    hi[6] = lc_period;
}

/* likelihood3.c:1135-1179 (the enlarged set is active because USE_COLOR_INFO == 0) */
void orc_proposal_sigmas(double *sigma, int use_gmag, int use_color)
{
    static const double base[ORC_NPARS] = {1e-2, 1e-2, 1e-8, 1e-2, 1e-3, 1e-3, 1e-3, 1e-1, 1e-1, 1e-2, 1e-2,
                                           1e-2, 1e-2, 1e-2, 1e-2, 1e-2, 1e-2, 1e-1, 1e-1, 1e-3, 1e-5};
    memcpy(sigma, base, sizeof(base));
    if ((!use_color) || (!use_gmag)) {
        sigma[0] = sigma[1] = 1e-1;
        sigma[4] = sigma[5] = 1e-2;
        sigma[6] = 1e-3;
        for (int i = 9; i <= 18; i++) sigma[i] = 1e-1;
    }
}

/* mcmc_wrapper2.c:1175-1178 with SQRT_2PI of mcmc_wrapper2.h:10 */
static double gaussian_pdf(double x, double mean, double sigma)
{
    return (1 / sigma / 2.5066282746) * exp(-pow((x - mean) / sigma, 2.) / 2.);
}

/* mcmc_wrapper2.c:703-765 */
double orc_get_logP(const double *pars, const int *gauss)
{
    double logP = 0.;
    for (int i = 0; i < ORC_NPARS; i++) {
        double mean, sigma;
        if (i == 7 || i == 8) { mean = 0.; sigma = 1.; }
        else if (i == 9 || i == 11) { mean = 0.16; sigma = 0.04; }
        else if (i == 10 || i == 12) { mean = 0.34; sigma = 0.04; }
        else if (i == 13 || i == 14) { mean = 1.; sigma = 0.2; }
        else if (i == 15 || i == 16) { mean = 0.; sigma = 0.1; }
        else if (i == 17 || i == 18) { mean = 0.; sigma = 1.; }
        else { mean = 0.; sigma = kBig; }
        if (gauss[i] == 1) logP += log(gaussian_pdf(pars[i], mean, sigma));
    }
    return logP;
}

/* Boundary handling + the post-proposal fix-ups of mcmc_wrapper2.c:440-481, applied in
 * place to a proposal y.  Keeps quirk Q5: the "order the masses" block assigns y[1] = y[0] and
 * then y[0] = y[1] (its tmp is never used), so both end up equal to the old y[0]. */
void orc_enforce_bounds(double *y, const double *lo, const double *hi, const double *mode_lo,
                        const double *mode_hi, double log_lc_period, double lc_period)
{
    for (int i = 0; i < ORC_NPARS; i++) {
        while (((mode_lo[i] == 1) && (y[i] < lo[i])) || ((mode_hi[i] == 1) && (y[i] > hi[i]))) {
            if (y[i] < lo[i]) y[i] = 2.0 * lo[i] - y[i];
            else y[i] = 2.0 * hi[i] - y[i];
        }
        while ((mode_lo[i] == 2) && (y[i] < lo[i])) y[i] = hi[i] + (y[i] - lo[i]);
        while ((mode_hi[i] == 2) && (y[i] > hi[i])) y[i] = lo[i] + (y[i] - hi[i]);
    }
    if (y[1] > y[0]) {
        y[1] = y[0];
        y[0] = y[1];
    }
    y[2] = log_lc_period;
    y[6] = fmod(y[6], lc_period);
}

/* Swap rule of mcmc_wrapper2.c:768-817 for the rung pair (b, b+1) given a uniform draw
 * beta in [0,1].  Returns 1 and swaps index[b], index[b+1] when accepted.  (dlogP is
 * computed and ignored by the reference, quirk Q8.) */
int orc_pt_swap_pair(int *index, const double *temp, const double *logL, int b, double beta)
{
    int a = b + 1;
    int olda = index[a], oldb = index[b];
    double heat1 = temp[a], heat2 = temp[b];
    double dlogL = logL[oldb] - logL[olda];
    double H = (heat2 - heat1) / (heat2 * heat1);
    double alpha = exp(dlogL * H);
    if (alpha >= beta) {
        index[a] = oldb;
        index[b] = olda;
        return 1;
    }
    return 0;
}

/* Metropolis-Hastings ratio of mcmc_wrapper2.c:495 */
double orc_hastings(double logLx, double logLy, double logPx, double logPy, double temp)
{
    return exp((logLy - logLx) / temp + (logPy - logPx));
}

/* ------------------------------------------------------------------ */
/* batch helper used as the "port" CPU baseline                        */
/* ------------------------------------------------------------------ */

/* n chains, params row-major [n][21]; OpenMP over chains when built with -fopenmp,
 * mirroring the rung-parallel loop of mcmc_wrapper2.c:383. */
void orc_loglikelihood_batch(const double *time, const double *flux, const double *noise, long N,
                             const double *params, long n, const double *mag_data, const double *magerr,
                             int use_gmag, int use_color, double *logL)
{
    double *clamped = (double *)malloc((size_t)(N > 0 ? N : 1) * sizeof(double));
    for (long i = 0; i < N; i++) clamped[i] = noise[i] < 1.e-5 ? 1.e-5 : noise[i];
#pragma omp parallel for schedule(dynamic)
    for (long c = 0; c < n; c++) {
        logL[c] = orc_loglikelihood(time, flux, clamped, N, params + c * ORC_NPARS, mag_data, magerr, use_gmag,
                                    use_color);
    }
    free(clamped);
}

/* ------------------------------------------------------------------ */
/* parallel-tempering step with the Philox stream of the device path   */
/* ------------------------------------------------------------------ */
/* The reference's RNG (ran2/gasdev2/rand, mcmc_wrapper2.c:796-974) is replaced on the device by
 * Philox4x32-10 (Salmon et al. 2011).  To check the device step deterministically the oracle
 * draws from the same counter-based stream: block n of stream (id, iter, stage) is
 * philox(counter = {id, iter, stage, n}, key = seed) and yields two uniforms in (0,1). */
#include <stdint.h>

static void philox4x32_10(const uint32_t ctr[4], uint32_t k0, uint32_t k1, uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct {
    uint32_t id, iter, stage, block, k0, k1;
    double buf[2];
    int left;
} orc_stream;

static void stream_open(orc_stream *s, unsigned long long seed, uint32_t id, uint32_t iter, uint32_t stage)
{
    s->id = id; s->iter = iter; s->stage = stage; s->block = 0; s->left = 0;
    s->k0 = (uint32_t)seed; s->k1 = (uint32_t)(seed >> 32);
}

static double stream_next(orc_stream *s)
{
    if (s->left == 0) {
        uint32_t ctr[4] = {s->id, s->iter, s->stage, s->block++}, r[4];
        philox4x32_10(ctr, s->k0, s->k1, r);
        uint64_t a = ((uint64_t)r[0] << 21) | (r[1] >> 11);
        uint64_t b = ((uint64_t)r[2] << 21) | (r[3] >> 11);
        s->buf[0] = ((double)a + 0.5) / 9007199254740992.0;
        s->buf[1] = ((double)b + 0.5) / 9007199254740992.0;
        s->left = 2;
    }
    return s->buf[2 - s->left--];
}

void orc_pt_uniforms(unsigned long long seed, unsigned id, unsigned iter, unsigned stage, int n, double *out)
{
    orc_stream s;
    stream_open(&s, seed, id, iter, stage);
    for (int i = 0; i < n; i++) out[i] = stream_next(&s);
}

/* known-answer test vector of Random123 (kat_vectors: philox4x32 10 rounds) */
void orc_philox_raw(const unsigned *ctr, const unsigned *key, unsigned *out)
{
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, r[4];
    philox4x32_10(c, key[0], key[1], r);
    for (int i = 0; i < 4; i++) out[i] = r[i];
}

static void stream_normal_pair(orc_stream *s, double *z0, double *z1)
{
    double u1 = stream_next(s), u2 = stream_next(s);
    double r = sqrt(-2.0 * log(u1)), a = 6.283185307179586 * u2;
    *z0 = r * cos(a);
    *z1 = r * sin(a);
}

/* Gaussian jump, mcmc_wrapper2.c:1062-1088 (normals consumed in pairs, component order) */
static void gaussian_jump(orc_stream *s, const double *x, const double *sigma, double scale, double temp, double *y)
{
    double sq = sqrt(temp), z[ORC_NPARS + 1];
    for (int n = 0; n < ORC_NPARS; n += 2) stream_normal_pair(s, &z[n], &z[n + 1]);
    for (int n = 0; n < ORC_NPARS; n++) y[n] = x[n] + z[n] * sigma[n] * sq * scale;
}

/* One proposal (mcmc_wrapper2.c:390-485) for stream `id` at iteration `iter`; history is this
 * rung's ring [npast][21]; quirks as in hb_pt.cuh.  Returns the jump type (1 Gaussian, 2 DE). */
int orc_pt_propose(unsigned long long seed, unsigned id, unsigned iter, double temp, int npast, int quirks,
                   const double *x, const double *history, const double *lo, const double *hi,
                   const double *mode_lo, const double *mode_hi, const int *gauss, const double *sigma,
                   double log_lc_period, double *y, double *logPy)
{
    orc_stream s;
    stream_open(&s, seed, id, iter, 0u);
    double alpha = stream_next(&s);
    double jscale = pow(10., -6. + 6. * alpha);
    int type = 1;
    int de = (stream_next(&s) < 0.5) && ((long)iter > (long)npast);
    if (!de) {
        gaussian_jump(&s, x, sigma, jscale, temp, y);
    } else {
        const double gamma = 2.388 / sqrt(2. * ORC_NPARS); /* mcmc_wrapper2.h:13 */
        int a = 0, b;
        if (!quirks) a = (int)(stream_next(&s) * npast);
        do { b = (int)(stream_next(&s) * npast); } while (b == a);
        int scaled = stream_next(&s) < 0.9;
        double eps_fac = quirks ? (gaussian_pdf(0., 0., 1.e-4) - 0.5) : 0.0; /* Q6 as compiled (c == 0) */
        double z[ORC_NPARS + 1], mag = 0.;
        if (scaled)
            for (int n = 0; n < ORC_NPARS; n += 2) stream_normal_pair(&s, &z[n], &z[n + 1]);
        for (int n = 0; n < ORC_NPARS; n++) {
            double dx = history[b * ORC_NPARS + n] - history[a * ORC_NPARS + n];
            double eps = dx * eps_fac;
            if (scaled) dx *= z[n] * gamma;
            dx += eps;
            y[n] = x[n] + dx;
            mag += (x[n] - y[n]) * (x[n] - y[n]);
        }
        type = 2;
        if (mag < 1e-6) {
            gaussian_jump(&s, x, sigma, jscale, temp, y);
            type = 1;
        }
    }
    /* bounds + fix-ups; the "as intended" variant swaps the masses instead of copying */
    if (quirks) {
        orc_enforce_bounds(y, lo, hi, mode_lo, mode_hi, log_lc_period, pow(10., log_lc_period));
    } else {
        /* same boundary loops as orc_enforce_bounds, then a real swap of the masses */
        for (int i = 0; i < ORC_NPARS; i++) {
            while (((mode_lo[i] == 1) && (y[i] < lo[i])) || ((mode_hi[i] == 1) && (y[i] > hi[i]))) {
                if (y[i] < lo[i]) y[i] = 2.0 * lo[i] - y[i];
                else y[i] = 2.0 * hi[i] - y[i];
            }
            while ((mode_lo[i] == 2) && (y[i] < lo[i])) y[i] = hi[i] + (y[i] - lo[i]);
            while ((mode_hi[i] == 2) && (y[i] > hi[i])) y[i] = lo[i] + (y[i] - hi[i]);
        }
        if (y[1] > y[0]) { double t = y[1]; y[1] = y[0]; y[0] = t; }
        y[2] = log_lc_period;
        y[6] = fmod(y[6], pow(10., log_lc_period));
    }
    *logPy = orc_get_logP(y, gauss);
    return type;
}

/* MH decision (mcmc_wrapper2.c:492-505) with the accept stream (stage 1) */
int orc_pt_accept(unsigned long long seed, unsigned id, unsigned iter, double temp, double logLx, double logLy,
                  double logPx, double logPy)
{
    orc_stream s;
    stream_open(&s, seed, id, iter, 1u);
    double alpha = stream_next(&s);
    return alpha <= orc_hastings(logLx, logLy, logPx, logPy, temp);
}

/* n_temps swap proposals (mcmc_wrapper2.c:554-563) with the swap stream (stage 2, id = 2^31 | ens) */
int orc_pt_swap_ensemble(unsigned long long seed, unsigned ens, unsigned iter, int n_temps, const double *temp,
                         int *index, const double *logL)
{
    orc_stream s;
    stream_open(&s, seed, 0x80000000u | ens, iter, 2u);
    int acc = 0;
    for (int k = 0; k < n_temps && n_temps > 1; k++) {
        int b = (int)(stream_next(&s) * (double)(n_temps - 1));
        if (b > n_temps - 2) b = n_temps - 2;
        double beta = stream_next(&s);
        acc += orc_pt_swap_pair(index, temp, logL, b, beta);
    }
    return acc;
}

/* ------------------------------------------------------------------ */
/* Gaia-colour sampler (GAIA_mcmc.c), SURVEY 8f row 4                  */
/* ------------------------------------------------------------------ */
/* 6 parameters {logM1, logM2, rr1, rr2, aT1, aT2}, NCHAINS = 20 rungs with ratio 1.2, NPAST = 100
 * (GAIA_mcmc.c:24-29,476).  The flavour differs from mcmc_wrapper2.c in several places (SURVEY
 * quirk Q11), all restated as written:
 *   - `gaussian` has no 1/2 in the exponent (:168-172);
 *   - the prior is log10 of that pdf with mean = box centre, sigma = half the box (:175-191) and
 *     enters the Hastings ratio as pow(10, dlogP) (:569);
 *   - the boundary conditions are single `if`s, not loops (:546-559);
 *   - the DE jump draws both history samples (a != b) and scales every COMPONENT by its own normal
 *     (:1007-1021);
 *   - the swaps run before the history ring is filled, interleaved rung by rung (:741-748).
 * GSL's ranlxs1 / libc rand() are replaced by the Philox streams of the device path. */
#define ORC_GPARS 6

/* GAIA_mcmc.c:168-172 */
double orc_gaia_gaussian(double x, double mean, double sigma)
{
    return (1 / sigma / 2.5066282746) * exp(-pow((x - mean) / sigma, 2.));
}

/* GAIA_mcmc.c:175-191 */
double orc_gaia_get_logP(const double *pars, const double *lo, const double *hi, const int *gauss)
{
    double logP = 0.;
    for (int i = 0; i < ORC_GPARS; i++) {
        if (gauss[i] == 1) {
            double mean = 0.5 * (lo[i] + hi[i]);
            double sigma = (hi[i] - lo[i]) / 2.;
            logP += log10(orc_gaia_gaussian(pars[i], mean, sigma));
        }
    }
    return logP;
}

/* GAIA_mcmc.c:346-390: all six parameters reflect at both ends; the four shape parameters carry
 * the Gaussian prior */
void orc_gaia_set_limits(double *lo, double *hi, double *mode_lo, double *mode_hi, int *gauss)
{
    for (int i = 0; i < ORC_GPARS; i++) {
        mode_lo[i] = 1; mode_hi[i] = 1;
        lo[i] = (i < 2) ? -1.5 : -3.;
        hi[i] = (i < 2) ? 2.0 : 3.;
        gauss[i] = (i < 2) ? 0 : 1;
    }
}

/* GAIA_mcmc.c:546-559 */
void orc_gaia_enforce_bounds(double *y, const double *lo, const double *hi, const double *mode_lo, const double *mode_hi)
{
    for (int i = 0; i < ORC_GPARS; i++) {
        if ((mode_lo[i] == 1) && (y[i] < lo[i])) y[i] = 2.0 * lo[i] - y[i];
        if ((mode_hi[i] == 1) && (y[i] > hi[i])) y[i] = 2.0 * hi[i] - y[i];
        if ((mode_lo[i] == 2) && (y[i] < lo[i])) y[i] = hi[i] + (y[i] - lo[i]);
        if ((mode_hi[i] == 2) && (y[i] > hi[i])) y[i] = lo[i] + (y[i] - hi[i]);
    }
}

/* GAIA_mcmc.c:987-1002 (normals consumed in pairs, component order) */
static void gaia_gaussian_jump(orc_stream *s, const double *x, const double *sigma, double scale, double temp, double *y)
{
    double sq = sqrt(temp), z[ORC_GPARS];
    for (int n = 0; n < ORC_GPARS; n += 2) stream_normal_pair(s, &z[n], &z[n + 1]);
    for (int n = 0; n < ORC_GPARS; n++) y[n] = x[n] + z[n] * sigma[n] * sq * scale;
}

/* One proposal of run_chain (GAIA_mcmc.c:525-563) for stream `id` at iteration `iter`; history is
 * this rung's ring [npast][6].  Returns the jump type (1 Gaussian, 2 DE). */
int orc_gaia_propose(unsigned long long seed, unsigned id, unsigned iter, double temp, int npast, const double *x,
                     const double *history, const double *lo, const double *hi, const double *mode_lo,
                     const double *mode_hi, const int *gauss, const double *sigma, double *y, double *logPy)
{
    orc_stream s;
    stream_open(&s, seed, id, iter, 0u);
    double alpha = stream_next(&s);
    double jscale = pow(10., -6. + 6. * alpha);
    int type = 1;
    int de = (stream_next(&s) < 0.5) && ((long)iter > (long)npast);
    if (!de) {
        gaia_gaussian_jump(&s, x, sigma, jscale, temp, y);
    } else {
        const double gamma = 2.388 / sqrt(2. * ORC_GPARS); /* GAIA_mcmc.c:27 */
        int a = (int)(stream_next(&s) * npast), b = a;
        while (b == a) b = (int)(stream_next(&s) * npast);
        double dx[ORC_GPARS], mag = 0.;
        for (int n = 0; n < ORC_GPARS; n++) dx[n] = history[b * ORC_GPARS + n] - history[a * ORC_GPARS + n];
        if (stream_next(&s) < 0.9) {
            double z[ORC_GPARS];
            for (int n = 0; n < ORC_GPARS; n += 2) stream_normal_pair(&s, &z[n], &z[n + 1]);
            for (int n = 0; n < ORC_GPARS; n++) dx[n] *= z[n] * gamma;
        }
        for (int n = 0; n < ORC_GPARS; n++) {
            y[n] = x[n] + dx[n];
            mag += (x[n] - y[n]) * (x[n] - y[n]);
        }
        type = 2;
        if (mag < 1e-6) {
            gaia_gaussian_jump(&s, x, sigma, jscale, temp, y);
            type = 1;
        }
    }
    orc_gaia_enforce_bounds(y, lo, hi, mode_lo, mode_hi);
    *logPy = orc_gaia_get_logP(y, lo, hi, gauss);
    return type;
}

/* GAIA_mcmc.c:569-574 with the accept stream (stage 1) */
int orc_gaia_accept(unsigned long long seed, unsigned id, unsigned iter, double temp, double logLx, double logLy,
                    double logPx, double logPy)
{
    orc_stream s;
    stream_open(&s, seed, id, iter, 1u);
    double alpha = stream_next(&s);
    double H = exp((logLy - logLx) / temp) * pow(10., logPy - logPx);
    return alpha <= H;
}

/* The post-step loop of run_mcmc (GAIA_mcmc.c:741-748): for every rung k, one swap proposal
 * (ptmcmc, :893-938, same rule as orc_pt_swap_pair) and THEN the history fill of rung k from the
 * slot index[k] holds at that moment; fill_slot[k] reports that slot.  Returns accepted swaps. */
int orc_gaia_swap_ensemble(unsigned long long seed, unsigned ens, unsigned iter, int n_temps, const double *temp,
                           int *index, const double *logL, int *fill_slot)
{
    orc_stream s;
    stream_open(&s, seed, 0x80000000u | ens, iter, 2u);
    int acc = 0;
    for (int k = 0; k < n_temps; k++) {
        if (n_temps > 1) {
            int b = (int)(stream_next(&s) * (double)(n_temps - 1));
            if (b > n_temps - 2) b = n_temps - 2;
            double beta = stream_next(&s);
            acc += orc_pt_swap_pair(index, temp, logL, b, beta);
        }
        fill_slot[k] = index[k];
    }
    return acc;
}
