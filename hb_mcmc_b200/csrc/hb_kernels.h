// hb_kernels.h -- launcher prototypes shared by hb_kernels.cu and hb_capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace hb {

struct ChainConst;
struct MagSetup;

constexpr int kEvalThreads = 256;   // threads per CTA of k_chain_eval
constexpr int kEvalCtasPerSm = 2;   // resident CTAs per SM (register budget: 128/thread)
constexpr int kCandA = 8192;        // shared-memory survivor buffers of the select (keys)
constexpr int kCandB = 2048;

size_t eval_smem_bytes();
cudaError_t configure_eval();
cudaError_t launch_prologue(const double* params, int n, const MagSetup& ms, ChainConst* out, cudaStream_t s);
cudaError_t launch_chain_eval(const ChainConst* cc, int n_chains, const double* t, const double* flux, const double* w,
                              int N, uint64_t* scratch, size_t scratch_stride, int grid, double* logL, double* lc_out,
                              int* counter, cudaStream_t s);
cudaError_t launch_order_stat(const double* x, int n, int k, uint64_t* scratch, size_t stride, double* out,
                              cudaStream_t s);
cudaError_t launch_traj(const double* times, int Nt, const double* tp, double* d, double* Z1, double* Z2, double* rr,
                        double* ff, cudaStream_t s);
cudaError_t launch_scalar(int op, const double* args, double* out, cudaStream_t s);
cudaError_t launch_gaia(const double* p6, int n, double D, const double* data, const double* err, double* mags,
                        double* logL, cudaStream_t s);
cudaError_t launch_chain_info(const ChainConst* cc, int n, double* out, cudaStream_t s);
cudaError_t launch_to_seconds(const double* t, int n, double* tsec, cudaStream_t s);
cudaError_t launch_fp64_peak(double* out, int blocks, int iters, cudaStream_t s);

}  // namespace hb
