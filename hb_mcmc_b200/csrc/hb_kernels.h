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
constexpr int kCandA = HB_CAND_A;              // shared-memory candidate list of short light curves / survivor buffer of the select (keys)
constexpr int kCandB = HB_CAND_B;              // work area of the select (keys); running chi^2 sums during the model pass

constexpr int kMaxSegments = 64;                // most segments (hence CTAs) one light curve is cut into

// hand-over words of a chain whose light curve is shared by several CTAs (zero between launches)
struct ChainSync {
    int cnt;    // candidate keys appended to the chain's flat list
    int c_lt;   // samples below the bracket
    int done;   // parts that have finished
    int flags;  // bit 0: a part saw a Newton iterate outside the table sincos' range
};

// arguments of k_chain_eval
struct EvalArgs {
    const ChainConst* cc;  // [n_chains], from k_prologue
    int n_chains;
    int N;                 // samples per light curve
    const double* tsec;    // t * 86400, padded with one whole tile beyond the last (partial) one
    const double2* fw;     // {flux, 1 / max(sigma, 1e-5)}, padded likewise with weight 0; nullptr: model only
    double sum_w2;         // sum of the squared weights (S2 of the chi^2 expansion), formed once per data set
    uint64_t* scratch;     // per region: template keys, two key buffers (key_stride each), nseg x 2 x threads partial sums
    size_t region_stride, key_stride;
    double* logL;          // [n_chains] or nullptr
    double* lc_out;        // [n_chains][N] or nullptr
    int* counter;          // work-item scheduler (armed by k_prologue)
    unsigned long long* evaluated;  // += 1 per chain whose model is really evaluated (not Roche / NaN early-outs)
    ChainSync* sync;       // [n_chains] when nparts > 1
    const double2* sctab;
    float bracket_sigma;
    int hot_hi_limit;
    int nparts;            // CTAs per chain (each takes a contiguous run of segments); > 1: the kernel raises it to what
                           // the chains that are really evaluated leave room for
    int max_parts;         // ... up to this many
    int nseg;              // eval_segments(N)
    int seg_shift;         // eval_seg_shift(N): a segment is 2^seg_shift tiles
};

size_t eval_smem_bytes();
int prologue_small_max();  // batches up to this many chains take k_prologue_small (and, from host buffers, no DMA copies)
int eval_tile();  // samples per tile: device time / flux / weight arrays are padded to whole tiles plus one
int eval_segments(long n_points);
int eval_seg_shift(long n_points);
cudaError_t configure_eval();
// eval_grid: the grid of the k_chain_eval launch that follows (its work-item counter starts there)
cudaError_t launch_prologue(const double* params, int n, const MagSetup& ms, ChainConst* out, int* eval_counter,
                            int eval_grid, cudaStream_t s);
cudaError_t launch_chain_eval(const EvalArgs& a, int grid, cudaStream_t s);
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
#ifdef HB_DEBUG_BOUNDS
cudaError_t launch_bounds_selftest(int i, int* out, cudaStream_t s);
#endif

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
