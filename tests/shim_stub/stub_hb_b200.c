/* TEST-ONLY CPU stand-in for libhb_b200.so (never shipped, never linked by the product): lets the host logic of
 * likelihood3_shim.c -- call combining, leader election, memo -- be stressed without a GPU.  "logL" is a
 * checkable function of the inputs (sum of the flux + weighted sum of the parameters) after a ~45 us busy wait
 * that stands for the device round trip. */
#define _POSIX_C_SOURCE 200809L
#include <time.h>
#include <string.h>
#include "hb_b200.h"
struct hb_ctx { int dummy; };
static struct hb_ctx g; static long gN; static double gsum;
static void busy_us(double us){ struct timespec a,b; clock_gettime(CLOCK_MONOTONIC,&a); do{clock_gettime(CLOCK_MONOTONIC,&b);}while((b.tv_sec-a.tv_sec)*1e6+(b.tv_nsec-a.tv_nsec)*1e-3<us); }
int hb_create(hb_ctx** out,int d){(void)d;*out=&g;return 0;}
void hb_destroy(hb_ctx*c){(void)c;}
const char* hb_last_error(const hb_ctx*c){(void)c;return "";}
const char* hb_global_error(void){return "";}
int hb_set_data(hb_ctx*c,const double*t,const double*f,const double*e,long n){(void)c;(void)t;(void)e;gN=n;gsum=0;for(long i=0;i<n;i++)gsum+=f[i];return 0;}
int hb_set_mags(hb_ctx*c,const double*m,const double*e,int a,int b){(void)c;(void)m;(void)e;(void)a;(void)b;return 0;}
int hb_loglikelihood_batch(hb_ctx*c,const double*p,long n,double*o){(void)c;busy_us(45.0);for(long i=0;i<n;i++){double s=gsum;for(int k=0;k<HB_NPARS;k++)s+=p[i*HB_NPARS+k]*(k+1);o[i]=s;}return 0;}
int hb_calc_light_curve(hb_ctx*c,const double*t,long n,const double*p,double*o){(void)c;(void)t;(void)p;memset(o,0,n*8);return 0;}
int hb_chain_info_batch(hb_ctx*c,const double*p,long n,double D,double*o){(void)c;(void)p;(void)D;memset(o,0,n*9*8);return 0;}
int hb_traj(hb_ctx*c,const double*t,long n,const double*tp,double*a,double*b,double*d,double*e,double*f){(void)c;(void)t;(void)n;(void)tp;(void)a;(void)b;(void)d;(void)e;(void)f;return 0;}
int hb_remove_median(hb_ctx*c,double*a,long n){(void)c;(void)a;(void)n;return 0;}
int hb_scalar(hb_ctx*c,int op,const double*a,int n,double*o){(void)c;(void)op;(void)a;(void)n;*o=0;return 0;}
