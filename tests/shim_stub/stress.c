/* Mimics the reference's rung loop (mcmc_wrapper2.c:383,488-489): T OpenMP threads over 50 rungs, per rung and step
 * one call on the current state (a repeat of an earlier call) and one on a proposal (new); every value is checked
 * against the stub's formula and against the value remembered for the state.  Exit code 1 on any mismatch. */
#include <stdio.h>
#include <stdlib.h>
#include <omp.h>
#include <math.h>
double loglikelihood(double t[], double f[], double e[], long N, double p[], double md[], double me[]);
#define NP 21
int main(int argc,char**argv){
  int steps=argc>1?atoi(argv[1]):2000, T=argc>2?atoi(argv[2]):25, R=50; long N=375;
  static double t[375],f[375],e[375]; for(int i=0;i<N;i++){t[i]=i;f[i]=1+1e-3*i;e[i]=3e-4;}
  double md[5]={1000,1,1,1,1},me[4]={1e15,1e15,1e15,1e15};
  static double x[50][NP], lx[50]; unsigned s=1; long bad=0;
  for(int j=0;j<R;j++){for(int k=0;k<NP;k++)x[j][k]=j+0.01*k; lx[j]=loglikelihood(t,f,e,N,x[j],md,me);}
  omp_set_num_threads(T);
  double t0=omp_get_wtime();
  for(int it=0;it<steps;it++){
    #pragma omp parallel for schedule(static) reduction(+:bad)
    for(int j=0;j<R;j++){
      double y[NP]; for(int k=0;k<NP;k++)y[k]=x[j][k]+1e-3*((it*31+j*7+k)%13);
      double a=loglikelihood(t,f,e,N,x[j],md,me), b=loglikelihood(t,f,e,N,y,md,me);
      double wa=0,wb=0; for(int i=0;i<N;i++){wa+=f[i];} wb=wa; for(int k=0;k<NP;k++){wa+=x[j][k]*(k+1);wb+=y[k]*(k+1);}
      if(a!=wa||b!=wb||a!=lx[j])bad++;
      if((it+j)%3==0){for(int k=0;k<NP;k++)x[j][k]=y[k]; lx[j]=b;}
    }
  }
  double dt=omp_get_wtime()-t0; (void)s;
  printf("%d steps, %d threads: %.1f steps/s, %.1f us/step, bad=%ld\n",steps,T,steps/dt,dt/steps*1e6,bad);
  return bad!=0;
}
