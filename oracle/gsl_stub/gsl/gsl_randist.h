/* oracle/gsl_stub -- see gsl_rng.h */
#ifndef HB_GSL_STUB_RANDIST_H
#define HB_GSL_STUB_RANDIST_H
#include <gsl/gsl_rng.h>
double gsl_ran_gaussian(gsl_rng *r, double sigma);
#endif
