// hb_kernels.h -- launcher prototypes shared by hb_kernels.cu and hb_capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace hb {

struct ChainConst;
struct MagSetup;

// tuning knobs (overridable with -D for sweeps; the defaults are the measured best on B200)
#ifndef HB_EVAL_THREADS
#define HB_EVAL_THREADS 256
#endif
#ifndef HB_EVAL_CTAS
#define HB_EVAL_CTAS 4
#endif
#ifndef HB_CAND_A
#define HB_CAND_A 2048
#endif
#ifndef HB_CAND_B
#define HB_CAND_B 1024
#endif
#ifndef HB_POINTS_PER_THREAD
#define HB_POINTS_PER_THREAD 1
#endif
constexpr int kPointsPerThread = HB_POINTS_PER_THREAD;  // independent samples in flight per thread
constexpr int kEvalThreads = HB_EVAL_THREADS;  // threads per CTA of k_chain_eval
constexpr int kEvalCtasPerSm = HB_EVAL_CTAS;   // resident CTAs per SM (sets the register budget)
constexpr int kCandA = HB_CAND_A;              // shared-memory survivor buffers of the select (keys)
constexpr int kCandB = HB_CAND_B;

size_t eval_smem_bytes();
int eval_tile();  // samples per TMA tile: device time / flux / weight arrays are padded to a multiple of it
cudaError_t configure_eval();
cudaError_t launch_prologue(const double* params, int n, const MagSetup& ms, ChainConst* out, int* eval_counter,
                            int eval_grid, cudaStream_t s);
cudaError_t launch_chain_eval(const ChainConst* cc, int n_chains, const double* t, const double2* fw,
                              int N, uint64_t* scratch, size_t scratch_stride, int grid, double* logL, double* lc_out,
                              int* counter, float bracket_sigma, const double2* sctab, int hot_hi_limit, cudaStream_t s);
cudaError_t launch_order_stat(const double* x, int n, int k, uint64_t* scratch, size_t stride, double* out,
                              cudaStream_t s);
cudaError_t launch_subtract(double* arr, int n, const double* value, cudaStream_t s);
cudaError_t launch_traj(const double* times, int Nt, const double* tp, double* d, double* Z1, double* Z2, double* rr,
                        double* ff, cudaStream_t s);
cudaError_t launch_scalar(int op, const double* args, double* out, cudaStream_t s);
cudaError_t launch_gaia(const double* p6, int n, double D, const double* data, const double* err, double* mags,
                        double* logL, cudaStream_t s);
cudaError_t launch_chain_info(const ChainConst* cc, int n, double* out, cudaStream_t s);
cudaError_t launch_to_seconds(const double* t, int n, double* tsec, cudaStream_t s);
cudaError_t launch_fp64_peak(double* out, int blocks, int iters, cudaStream_t s);

// parallel tempering (hb_pt.cu)
struct PtConfig;
cudaError_t launch_pt_init_random(const PtConfig* cfg, double* x, int W, cudaStream_t s);
cudaError_t launch_pt_propose(const PtConfig* cfg, const unsigned* iter, const double* x, const int* index, const double* history,
                              double* y, double* logPy, int* jump, int W, cudaStream_t s);
cudaError_t launch_pt_accept(const PtConfig* cfg, const unsigned* iter, double* x, const double* y, double* logLx,
                             const double* logLy, const double* logPy, const int* jump, const int* index, double* history,
                             unsigned long long* counters, int W, cudaStream_t s);
cudaError_t launch_pt_swap(const PtConfig* cfg, unsigned* iter, int* index, const double* logLx, const double* x,
                           unsigned long long* counters, double* xmap, double* logLmap, int E, cudaStream_t s);
cudaError_t launch_pt_gather_cold(const PtConfig* cfg, const int* index, const double* x, const double* logLx,
                                  double* out_x, double* out_logL, int E, cudaStream_t s);
cudaError_t launch_pt_logL_by_rung(const PtConfig* cfg, const int* index, const double* logLx, double* out, int W,
                                   cudaStream_t s);

}  // namespace hb
