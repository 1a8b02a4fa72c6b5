// hb_pt_run.h -- the one-launch step loop of the sampler (k_pt_run, hb_pt.cu): arguments and launcher.
#pragma once
#include <cuda_runtime.h>

#include "hb_device.cuh"
#include "hb_pt.cuh"

namespace hb {

constexpr int kPtRunMaxPoints = 1024;  // longest light curve the one-launch step loop takes (k_pt_run)
// arguments of k_pt_run (all device pointers; the sampler's arrays as in hb_capi.cu)
struct PtRunArgs {
    const PtConfig* cfg;
    unsigned* d_iter;
    double *x, *y, *logLx, *logLy, *logPy;
    double* logPx;                   // [W] log prior of the current state, by slot (filled at launch, kept by the loop)
    int *jump, *index;
    double* history;
    unsigned long long* counters;
    double *xmap, *logLmap;
    const double* tsec;
    const double2* fw;
    int N;
    const double2* sctab;
    unsigned* barrier;               // zero at launch
    unsigned long long* evaluated;   // += 1 per chain whose model is evaluated
    long n_iters;
    MagSetup ms;
};
size_t pt_run_smem_bytes();
cudaError_t configure_pt_run();
cudaError_t pt_run_max_walkers(int sm_count, int* out);
cudaError_t launch_pt_run(const PtRunArgs& a, int W, cudaStream_t s);

}  // namespace hb
