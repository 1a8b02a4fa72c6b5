/*
 * hb_likelihood3_abi.h -- symbols exported by libhb_likelihood3.so, the link-level drop-in for
 * the reference's likelihood3.c.  The prototypes are the reference's own (likelihood3.h:68-89;
 * the last four are the un-headered helpers its Cython binding uses, likelihood3.pxd:10-13), so
 * a caller keeps including ITS likelihood3.h and only changes what it links against.  This header
 * exists so the test-suite can check the export list; it declares nothing new except the
 * hb_shim_* management calls.
 *
 *   symbol                 replaces (reference file:line)        computed
 *   partition, quickSort   likelihood3.c:48-83                   host (generic array helpers)
 *   remove_median          likelihood3.c:86-105                  device (exact order statistic)
 *   traj                   likelihood3.c:125-185                 device
 *   get_alpha_beam         likelihood3.c:194-209                 device
 *   beaming                likelihood3.c:224-236                 device
 *   ellipsoidal            likelihood3.c:255-307                 device
 *   reflection             likelihood3.c:322-337                 device
 *   eclipse_area           likelihood3.c:353-389                 device
 *   calc_mags              likelihood3.c:725-795                 device
 *   calc_light_curve       likelihood3.c:530-686                 device
 *   calc_radii_and_Teffs   likelihood3.c:693-717                 device
 *   RocheOverflow          likelihood3.c:953-974                 device
 *   loglikelihood          likelihood3.c:809-873                 device (clamps noise[] in place, Q2)
 *   set_limits             likelihood3.c:986-1121                host (constant table)
 *   initialize_proposals   likelihood3.c:1123-1211               host (constant table)
 *   _getT, _getR           likelihood3.c:396-476                 device
 *   envelope_Temp/Radius   likelihood3.c:483-507                 device
 */
#ifndef HB_LIKELIHOOD3_ABI_H
#define HB_LIKELIHOOD3_ABI_H
#ifdef __cplusplus
extern "C" {
#endif

struct bounds;
struct gauss_bounds;

double partition(double arr[], int low, int high);
void quickSort(double arr[], int low, int high);
void remove_median(double *arr, long begin, long end);
void traj(double *times, double *traj_pars, double *d_arr, double *Z1_arr, double *Z2_arr, double *rr_arr,
          double *ff_arr, int Nt);
double get_alpha_beam(double logT);
double beaming(double P, double M1, double M2, double e, double inc, double omega0, double nu, double alpha_beam);
double ellipsoidal(double P, double M1, double M2, double e, double inc, double omega0, double nu, double R1, double a,
                   double mu, double tau);
double reflection(double P, double M1, double M2, double e, double inc, double omega0, double nu, double R2,
                  double alpha_ref1);
double eclipse_area(double R1, double R2, double d);
void calc_mags(double params[], double D, double *Gmg, double *BminusV, double *VminusG, double *GminusT);
void calc_light_curve(double *times, long Nt, double *pars, double *template_);
void calc_radii_and_Teffs(double params[], double *R1, double *R2, double *Teff1, double *Teff2);
int RocheOverflow(double *pars);
double loglikelihood(double time[], double lightcurve[], double noise[], long N, double params[], double mag_data[],
                     double magerr[]);
void set_limits(struct bounds *limited, struct bounds *limits, struct gauss_bounds *gauss_pars, double LC_PERIOD);
void initialize_proposals(double *sigma, double ***history);
double _getT(double logM);
double _getR(double logM);
double envelope_Temp(double logM);
double envelope_Radius(double logM);

/* USE_GMAG / USE_COLOR_INFO (likelihood3.h:11-12) at run time; defaults 1 / 0 */
void hb_shim_set_flags(int use_gmag, int use_color);
void hb_shim_shutdown(void);
/* loglikelihood() answers a call whose inputs (parameters, magnitudes, data arrays -- compared bit for bit)
 * repeat a recent one from a memo: the reference driver re-evaluates every rung's current state at every step
 * (mcmc_wrapper2.c:488).  on = 0 evaluates every call on the device (also: env HB_SHIM_MEMO=0). */
void hb_shim_set_memo(int on);
long hb_shim_memo_hits(void);

#ifdef __cplusplus
}
#endif
#endif
