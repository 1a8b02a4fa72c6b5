/* hb_oracle.h -- prototypes of the CPU restatement (TEST INFRASTRUCTURE ONLY, see hb_oracle.c). */
#ifndef HB_ORACLE_H
#define HB_ORACLE_H

#define ORC_NPARS 21 /* likelihood3.h:20 */

#ifdef __cplusplus
extern "C" {
#endif

double orc_getT(double logM);
double orc_getR(double logM);
double orc_envelope_temp(double logM);
double orc_envelope_radius(double logM);
void orc_radii_teffs(const double *p, double *R1, double *R2, double *T1, double *T2);
double orc_alpha_beam(double logT);

void orc_traj(const double *times, const double *tp, double *d_arr, double *Z1_arr, double *Z2_arr,
              double *rr_arr, double *ff_arr, long Nt);
double orc_beaming(double P, double M1, double M2, double e, double inc, double omega0, double nu,
                   double alpha_beam);
double orc_ellipsoidal(double P, double M1, double M2, double e, double inc, double omega0, double nu,
                       double R1, double a, double mu, double tau);
double orc_reflection(double P, double M1, double M2, double e, double inc, double omega0, double nu,
                      double R2, double alpha_ref);
double orc_eclipse_area(double R1, double R2, double d);

long orc_median_rank(long N);
void orc_remove_median(double *arr, long begin, long end);
void orc_calc_light_curve(const double *times, long Nt, const double *pars, double *tmpl);
void orc_calc_light_curve_ex(const double *times, long Nt, const double *pars, double *tmpl, double *raw);

void orc_calc_mags(const double *p, double D, double out[4]);
void orc_gaia_get_mags(const double *p6, double D, double out[4]);
double orc_gaia_model_likelihood(const double *data, const double *err, const double *p6, double D);
int orc_roche_overflow(const double *p);
double orc_loglikelihood(const double *time, const double *flux, double *noise, long N, const double *params,
                         const double *mag_data, const double *magerr, int use_gmag, int use_color);
void orc_loglikelihood_batch(const double *time, const double *flux, const double *noise, long N,
                             const double *params, long n, const double *mag_data, const double *magerr,
                             int use_gmag, int use_color, double *logL);

void orc_set_limits(double *lo, double *hi, double *mode_lo, double *mode_hi, int *gauss, double lc_period);
void orc_proposal_sigmas(double *sigma, int use_gmag, int use_color);
double orc_get_logP(const double *pars, const int *gauss);
void orc_enforce_bounds(double *y, const double *lo, const double *hi, const double *mode_lo,
                        const double *mode_hi, double log_lc_period, double lc_period);
int orc_pt_swap_pair(int *index, const double *temp, const double *logL, int b, double beta);
double orc_hastings(double logLx, double logLy, double logPx, double logPy, double temp);

#ifdef __cplusplus
}
#endif
#endif

/* parallel-tempering step on the device's Philox stream (declared late: see hb_oracle.c) */
#ifdef __cplusplus
extern "C" {
#endif
void orc_pt_uniforms(unsigned long long seed, unsigned id, unsigned iter, unsigned stage, int n, double *out);
void orc_philox_raw(const unsigned *ctr, const unsigned *key, unsigned *out);
int orc_pt_propose(unsigned long long seed, unsigned id, unsigned iter, double temp, int npast, int quirks,
                   const double *x, const double *history, const double *lo, const double *hi,
                   const double *mode_lo, const double *mode_hi, const int *gauss, const double *sigma,
                   double log_lc_period, double *y, double *logPy);
int orc_pt_accept(unsigned long long seed, unsigned id, unsigned iter, double temp, double logLx, double logLy,
                  double logPx, double logPy);
int orc_pt_swap_ensemble(unsigned long long seed, unsigned ens, unsigned iter, int n_temps, const double *temp,
                         int *index, const double *logL);
/* Gaia-colour sampler, GAIA_mcmc.c (6 parameters) */
double orc_gaia_gaussian(double x, double mean, double sigma);
double orc_gaia_get_logP(const double *pars, const double *lo, const double *hi, const int *gauss);
void orc_gaia_set_limits(double *lo, double *hi, double *mode_lo, double *mode_hi, int *gauss);
void orc_gaia_enforce_bounds(double *y, const double *lo, const double *hi, const double *mode_lo, const double *mode_hi);
int orc_gaia_propose(unsigned long long seed, unsigned id, unsigned iter, double temp, int npast, const double *x,
                     const double *history, const double *lo, const double *hi, const double *mode_lo,
                     const double *mode_hi, const int *gauss, const double *sigma, double *y, double *logPy);
int orc_gaia_accept(unsigned long long seed, unsigned id, unsigned iter, double temp, double logLx, double logLy,
                    double logPx, double logPy);
int orc_gaia_swap_ensemble(unsigned long long seed, unsigned ens, unsigned iter, int n_temps, const double *temp,
                           int *index, const double *logL, int *fill_slot);

#ifdef __cplusplus
}
#endif
