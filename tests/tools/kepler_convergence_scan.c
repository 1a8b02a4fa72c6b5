// Where do the reference's five Newton steps (likelihood3.c:153-160) reach the root of Kepler's equation to rounding?
// gcc -O2 -std=c99 kepler_convergence_scan.c -lm && ./a.out   (4 minutes)
// Result (glibc 2.39): converged for every M up to e = 0.86; for 0.87 <= e <= 0.99 the un-converged M lie within
// |M| <= 0.0493 of periastron (largest at e = 0.91).  hb_device.cuh uses the E(M) table outside |M| >= 0.1 only.
#include <math.h>
#include <stdio.h>
// reference's solve (likelihood3.c:153-160): starter + 5 Newton steps, double, libm sin/cos
static double ref5(double M, double e){ double E=M; double s=sin(M); if(s!=0) E = M + 0.85*e*(s>0?1:-1); for(int k=0;k<5;k++) E -= (E - e*sin(E) - M)/(1 - e*cos(E)); return E; }
static long double root(long double M, long double e){ // robust: bisection on [M, M+e] then Newton in long double
  long double lo=M, hi=M+e; for(int i=0;i<200;i++){ long double mid=(lo+hi)/2; if(mid - e*sinl(mid) - M > 0) hi=mid; else lo=mid;} long double E=(lo+hi)/2; for(int k=0;k<5;k++) E -= (E - e*sinl(E) - M)/(1 - e*cosl(E)); return E; }
int main(){
  for(double e=0.80; e<=0.9901; e+=0.01){
    double worstM=0, worstErr=0; long nbad=0;
    // M in (0, pi]: log grid near zero + uniform
    for(int i=0;i<400000;i++){
      double M = (i<200000)? pow(10., -8 + 8.5*i/200000.) : 3.14159265358979*(i-200000+1)/200000.;
      if(M>3.14159265358979) continue;
      double E5=ref5(M,e); long double Es=root(M,e);
      double err=fabs((double)(E5-Es));
      double tol=4e-15*fmax(1.0,fabs((double)Es));
      if(err>tol){ nbad++; if(M>worstM) worstM=M; if(err>worstErr) worstErr=err; }
    }
    printf("e=%.2f  nbad=%ld  largest bad M=%.6g  (1-e)=%.3g  ratio M/(1-e)=%.4g  worst err=%.3g\n", e, nbad, worstM, 1-e, worstM/(1-e), worstErr);
  }
  return 0; }
