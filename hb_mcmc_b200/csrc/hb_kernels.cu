// hb_kernels.cu -- sm_100a kernels of the batched heartbeat-star likelihood.
//
//   k_prologue    one warp per chain: parameters -> ChainConst (hb_device.cuh)
//   k_chain_eval  persistent CTAs, one chain at a time per CTA:
//                   pre-sample  256 model values -> bracket of the median rank + chi^2 pivot
//                   model pass  u at every sample; candidates in the bracket, chi^2 partial sums
//                   select      exact order statistic (reference median rule, quirk Q3) -> logL
//   k_order_stat / k_traj / k_scalar / k_gaia / k_chain_info   the small likelihood3.h entry points
//   k_fp64_peak   DFMA throughput probe (the roofline denominator, measured on the box)
//
// Replaces likelihood3.c:809-873 (loglikelihood) and :530-686 (calc_light_curve).
#include <algorithm>
#include <type_traits>

#include "hb_kernels.h"
#include "hb_select.cuh"
#include "hb_device.cuh"

// -DHB_PHASE_PROF (tools/phase_time.py, never in the product build): thread 0 of every CTA adds the clock64 cycles of
// each per-chain phase of k_chain_eval to g_phase[] -- 0 table, 1 pre-sample, 2 model pass, 3 sums / hand-over, 4 select.
#ifdef HB_PHASE_PROF
__device__ unsigned long long g_phase[8];
#define HB_T0() long long t_prev_ = clock64()
#define HB_T(k) do { if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&g_phase[k], (unsigned long long)(now_ - t_prev_)); t_prev_ = now_; } } while (0)
extern "C" void hb_phase_read(unsigned long long* out)
{
    cudaMemcpyFromSymbol(out, g_phase, sizeof(unsigned long long) * 8);
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g_phase, z, sizeof(z));
}
#else
#define HB_T0() ((void)0)
#define HB_T(k) ((void)0)
#endif
namespace hb {

// ---------------------------------------------------------------------------
// One warp per chain: the lanes share out the ~50 libm calls (prologue_trans_warp), lane 0 assembles
// and the warp stores the 47 doubles of ChainConst together.
#ifndef HB_PROLOGUE_SMALL_MAX
#define HB_PROLOGUE_SMALL_MAX 148  // chains up to which a batch gets four warps per chain (k_prologue_small; measured: a gain up to ~150 chains, a loss at 500)
#endif
constexpr int kPrologueSmallMax = HB_PROLOGUE_SMALL_MAX;
#ifndef HB_PROLOGUE_BLOCKS
#define HB_PROLOGUE_BLOCKS 4  // 128 registers: 16 warps per SM; fewer (184 registers) leaves 4096 chains waiting in 3.5 waves
#endif
__global__ void __launch_bounds__(128, HB_PROLOGUE_BLOCKS) k_prologue(const double* __restrict__ params, int n, MagSetup ms,
                                                  ChainConst* __restrict__ out, int* __restrict__ eval_counter, int eval_grid)
{
    // arm the chain scheduler of the k_chain_eval launch that follows on the stream (saves a memset node):
    // its CTAs start on chains 0 .. grid-1 and fetch the next ones from here
    if (eval_counter != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *eval_counter = eval_grid;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= n) return;
    __shared__ ChainConst s_cc[4];
    ChainConst& cc = s_cc[threadIdx.x >> 5];
    // the chain's parameters: ONE coalesced read into shared memory -- a caller's page-locked buffer is read in place
    // (mapped host memory, hb_loglikelihood_batch), where every further access would cross the bus again
    __shared__ double s_p[4][NPARS];
    double* const sp = s_p[threadIdx.x >> 5];
    if (lane < NPARS) sp[lane] = params[(size_t)c * NPARS + lane];
    __syncwarp();
    const double* p = sp;
    PrologueT T;
    prologue_trans_warp(p, ms, T, lane);
    if (lane == 0) prologue_assemble(p, ms, T, cc);
    __syncwarp();
    const double* src = reinterpret_cast<const double*>(&cc);
    double* dst = reinterpret_cast<double*>(out + c);
    for (int i = lane; i < (int)(sizeof(ChainConst) / sizeof(double)); i += 32) dst[i] = src[i];
}

// The same for a batch that leaves the machine idle anyway (a ladder, the calls of one OpenMP team, C1): a CTA of
// four warps per chain.  Every warp takes the libm levels for itself -- the same calls side by side on the four
// sub-partitions -- and then one of the four sections of the assembly (~45 IEEE divisions in all), so the chain's
// assembly takes a third of the time it takes one lane (measured: a 50-rung ladder's call 1-3 us shorter).  Same
// expressions in the same order: same bits.
__global__ void __launch_bounds__(128) k_prologue_small(const double* __restrict__ params, int n, MagSetup ms,
                                                        ChainConst* __restrict__ out, int* __restrict__ eval_counter, int eval_grid)
{
    if (eval_counter != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *eval_counter = eval_grid;
    const int c = blockIdx.x, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (c >= n) return;
    __shared__ ChainConst cc;
    __shared__ double sp[NPARS];  // the chain's parameters: one coalesced read (they may sit in mapped host memory)
    if (threadIdx.x < NPARS) sp[threadIdx.x] = params[(size_t)c * NPARS + threadIdx.x];
    __syncthreads();
    const double* p = sp;
    PrologueT T;
    prologue_trans_warp(p, ms, T, lane);
    if (lane == 0) {
        if (wid == 0) prologue_assemble_sections<kAsmStarA>(p, ms, T, cc);
        else if (wid == 1) prologue_assemble_sections<kAsmStarB>(p, ms, T, cc);
        else if (wid == 2) prologue_assemble_sections<kAsmOrbit>(p, ms, T, cc);
        else prologue_assemble_sections<kAsmAux>(p, ms, T, cc);
    }
    __syncthreads();
    const double* src = reinterpret_cast<const double*>(&cc);
    double* dst = reinterpret_cast<double*>(out + c);
    for (int i = threadIdx.x; i < (int)(sizeof(ChainConst) / sizeof(double)); i += 128) dst[i] = src[i];
}

// ---- staging of the data stream ---------------------------------------------------------------
// The observed light curve (t, flux, 1/sigma: 24 B per sample) is read by every chain and stays
// L2-resident.  Two ways to bring it to the math were built and measured on B200 (C2, 4096 x 20k):
//   HB_TMA_STAGING = 0 (default)  coalesced LDG (8 B of t one iteration ahead, 16 B of {flux, 1/sigma}),
//                                 software-pipelined in registers: 0.786 ms per call
//   HB_TMA_STAGING = 1            TMA bulk copies (cp.async.bulk -> SASS UBLKCP) of whole 256-sample
//                                 tiles into a 2-stage shared-memory ring, completion by mbarrier
//                                 expect_tx/complete_tx, stage reuse by a second mbarrier the warps
//                                 arrive on: 0.971 ms (the 12 KB ring costs the fourth CTA per SM and
//                                 couples the warps to within one iteration of each other); same bits.
// At ~1000 cycles of FP64 work per 24 bytes the stream is far from any bandwidth limit, so the
// variant with the least synchronisation and the smallest footprint wins; the TMA path stays selectable.
#ifndef HB_TMA_STAGING
#define HB_TMA_STAGING 0
#endif
constexpr int kTile = kPointsPerThread * kEvalThreads;  // samples per loop iteration of a CTA
[[maybe_unused]] constexpr int kStages = 2;

struct TileStage {
    double ts[kTile];
    double2 fw[kTile];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// doubles pairs of the E(M) table that do not fit into candA
constexpr int kKtabHead = 2 * kTableNodes - kCandA / 2 > 0 ? 2 * kTableNodes - kCandA / 2 : 1;
struct EvalShared {
#if HB_TMA_STAGING
    alignas(16) TileStage stage[kStages];
    uint64_t full_bar[kStages], empty_bar[kStages];
#endif
    ChainConst cc;
    SelectCtl<kEvalThreads> ctl;
    double red[2 * 32];
    double pivot;  // chi^2 expansion point u0 (a template value near the median)
    int bcast[4];  // block-wide broadcasts of the CTA that finishes a shared chain
    double2 sctab[kSinTabN];  // {sin, cos}(2 pi k / 1024) for sincos_tab, copied once per CTA
    // The chain's E(M) starter table (hb_device.cuh: kTableNodes x 32 bytes) begins here and runs on THROUGH candA: a
    // chain that takes the table keeps its candidates in global scratch, and the select -- the other user of candA --
    // only starts when the last model pass of the chain is over (the next chain builds its own table).
    alignas(16) double2 ktab_head[kKtabHead];
    uint64_t candA[kCandA];
    uint64_t candB[kCandB];   // work area of the select; during the model pass: the running chi^2 sums [2][threads]
};
static_assert(kCandB * sizeof(uint64_t) >= 2 * kEvalThreads * sizeof(double), "candB holds the running sums of the pass");
static_assert(offsetof(EvalShared, candA) == offsetof(EvalShared, ktab_head) + sizeof(double2) * kKtabHead,
              "the E(M) table runs from ktab_head into candA without a gap");
// four CTAs per SM: 228 KB of shared memory less 1 KB reserved per CTA
static_assert(HB_TMA_STAGING || kEvalCtasPerSm != 4 || sizeof(EvalShared) <= 56 * 1024, "EvalShared no longer fits four CTAs per SM");

size_t eval_smem_bytes() { return sizeof(EvalShared); }
int prologue_small_max() { return kPrologueSmallMax; }
int eval_tile() { return kTile; }

// Median rank of likelihood3.c:97-101 (quirk Q3): even N -> N/2, odd N -> N/2 + 1.  N == 1
// would read one past the end in the reference; the only element is used instead.
__device__ __forceinline__ int median_rank(int N)
{
    int r = (N % 2 == 0) ? N / 2 : N / 2 + 1;
    return r < N ? r : N - 1;
}

// One chain at a time per CTA (persistent CTAs; first work item = block index, further ones from an atomic
// counter).  A work item is (chain, part): the time axis of a light curve is cut into SEGMENTS of 2^seg_shift
// tiles -- a function of N alone (eval_seg_shift) -- and a launch hands every chain to `nparts` CTAs (1 for
// batches that fill the grid), each taking a contiguous run of segments.  The chi^2 sums are formed per thread
// and per segment and added up in segment order, then across the block: the same operations in the same order
// whatever `nparts` is, so a chain's logL does not depend on the batch it arrives in, while a batch smaller than
// the grid (one chain, a 50-rung ladder of long light curves) spreads every light curve over many SMs.  The
// reference loops serially over the samples (likelihood3.c:147,649,822).
// Per work item:
//   table       E(M) starter table of the chain in shared memory (e <= 0.99, build_kepler_table)
//   pre-sample  every thread evaluates the model at one jittered-stride sample of the WHOLE light curve; the
//               block sorts the kThreads values; sample order statistics give a bracket [lo, hi] around the
//               reference's median rank and the expansion point u0 of the chi^2 (every part of a chain repeats
//               this bit for bit, so all agree on the bracket without talking to each other).
//   model pass  u_i at every sample of the part (the FP64-bound part): #(u < lo) counted, u in [lo, hi] appended
//               to the chain's candidate list (warp-aggregated atomic; shared memory, or global scratch for long
//               or shared light curves, which stays in L2), and per segment the two sums of
//                   chi^2(m) = S0 + 2 d S1 + d^2 S2,   d = -A (m - u0),
//               S0 = sum ((A (u_i - u0) + ft - f_i) w_i)^2,  S1 = sum (A (u_i - u0) + ft - f_i) w_i^2,
//               S2 = sum w_i^2 (per data set, from the host)  (model_i = A (u_i - m) + ft, A = ft (1 - blending),
//               likelihood3.c:681-685).  Compile-time variants: the logL-only pass keeps everything on chip (one
//               instance for a CTA that owns its chain, one for a part of a shared chain); the general one also
//               stores the template keys (light-curve output, small N, re-runs) and checks per sample what the
//               hot one checks per chain (sincos range, NaN).
//   hand-over   (nparts > 1) counts and flags go to the chain's ChainSync words; the CTA that arrives last
//               finishes the chain, the others move on.
//   select      exact order statistic among the candidates: one histogram pass (block_select_hist), else
//               sampling rounds (block_select_key); a chain whose bracket missed or whose Newton iterates
//               left the table sincos' range is evaluated once more, whole, with the general pass.
// Between chains a CTA sits at barriers while its latency-bound neighbours on the SM cannot speed up, so
// the per-chain phases count in full: clock64 gives table 3 k, pre-sample 6-8 k, pass 120-220 k, select
// 8-13 k, final 2 k cycles at N = 20 000.
// kShared: the launch hands every chain to several CTAs (a batch smaller than the grid); compiled apart so that the
// pass of a batch that fills the grid carries none of the hand-over code.
template <int kThreads, bool kShared>
__global__ void __launch_bounds__(kThreads, kEvalCtasPerSm)
k_chain_eval(const ChainConst* __restrict__ cc_all, int n_chains, const double* __restrict__ tsec,
             const double2* __restrict__ fw, int N, uint64_t* __restrict__ scratch,
             size_t region_stride, size_t key_stride, double* __restrict__ logL, double* __restrict__ lc_out,
             int* __restrict__ counter, float bracket_sigma, const double2* __restrict__ sctab_g, int hot_hi_limit,
             double sum_w2, unsigned long long* __restrict__ evaluated, ChainSync* __restrict__ sync_all, int nparts,
             int nseg, int seg_shift, int max_parts, int seg_mask)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EvalShared& sm = *reinterpret_cast<EvalShared*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    __shared__ int s_work;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const int krank = median_rank(N);
    const int n_tiles = (N + kTile - 1) / kTile;
    int n_work = n_chains * nparts;
    static_assert((kTile & (kTile - 1)) == 0, "segment ends are read off the bits of the sample index");
    // (seg_mask = 2^seg_shift - 1, the tile-in-segment bits of a tile index, comes as a parameter: a constant-bank operand)
    for (int i = tid; i < kSinTabN; i += kThreads) sm.sctab[i] = sctab_g[i];  // published by the first barrier below
    const double2* sctab = sm.sctab;
#if HB_TMA_STAGING
    if (tid == 0) {
        for (int st = 0; st < kStages; st++) {
            mbar_init(&sm.full_bar[st], 1);
            mbar_init(&sm.empty_bar[st], kThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t git = 0;  // tiles consumed so far by this CTA (all chains): stage = git % 2, phase = git / 2
    const uint32_t tile_bytes = kTile * sizeof(double);
#endif

    // ---- shared batches: only the chains that will be evaluated get CTAs ----
    // A batch smaller than the grid (nparts > 1) is a ladder of a sampler more often than not, and the hot rungs of a
    // ladder propose Roche-overflowing or unphysical states about every other step (quirk Q13: no evaluation).  Every
    // CTA reads the flags of the batch, numbers the chains that need the model (in chain order: no atomics, nothing to
    // reset) and spreads the grid over THOSE: each gets gridDim.x / n_eval CTAs instead of gridDim.x / n_chains.  The
    // skipped chains' results are written here.  A chain's value does not depend on its number of parts (see above).
    int shared_chain = 0;
    if (kShared && n_chains > 1) {
        static_assert(kThreads * sizeof(uint64_t) >= 2 * kThreads * sizeof(int), "the sort's exchange buffer holds the list");
        int* const s_list = reinterpret_cast<int*>(sm.ctl.xch);  // [2 kThreads] >= n_chains (the host side sees to that)
        int* const s_wcnt = sm.ctl.ired;                          // [2][kThreads / 32]
        bool ev[2];
        unsigned below[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int c = h * kThreads + tid;
            ev[h] = false;
            if (c < n_chains) {
                const int f = (int)cc_all[c].flag;
                ev[h] = (f & 3) == 0;  // neither Roche overflow nor a model that is NaN by construction
                if (!ev[h] && logL != nullptr && c % (int)gridDim.x == (int)blockIdx.x)
                    logL[c] = (f & 1) ? -0.5 * kBig : qnan;
            }
            const unsigned m = __ballot_sync(0xffffffffu, ev[h]);
            below[h] = __popc(m & ((1u << lane) - 1u));
            if (lane == 0) s_wcnt[h * (kThreads / 32) + (tid >> 5)] = __popc(m);
        }
        __syncthreads();
        int n_eval = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            int base = n_eval;  // chains of the lower half come first: chain order
            for (int w = 0; w < kThreads / 32; w++) {
                const int cw = s_wcnt[h * (kThreads / 32) + w];
                if (w < (tid >> 5)) base += cw;
                n_eval += cw;
            }
            if (ev[h]) s_list[base + (int)below[h]] = h * kThreads + tid;
        }
        __syncthreads();
        if (n_eval == 0) return;
        const int p_room = min(min(nseg, max_parts), (int)gridDim.x / n_eval);  // (>= the host's nparts: n_eval <= n_chains)
        const int spp = (nseg + p_room - 1) / p_room;
        nparts = (nseg + spp - 1) / spp;  // parts that get at least one segment
        n_work = n_eval * nparts;
        if ((int)blockIdx.x >= n_work) return;
        shared_chain = s_list[(int)blockIdx.x / nparts];
        __syncthreads();  // (the list lives in the sort's exchange buffer)
    }

    for (int round = 0;; round++) {
        // dynamic scheduler: chains differ in cost (eclipse fraction, Roche early-out).  The first work item of
        // a CTA is its own index (the counter starts at gridDim.x), so a batch that fits one wave -- a PT step
        // at the reference's size, every shared batch -- runs without a single atomic round trip.
        if (round > 0) {
            if (n_work <= (int)gridDim.x) break;
            if (tid == 0) s_work = atomicAdd(counter, 1);
            __syncthreads();
        }
        const int work = round == 0 ? (int)blockIdx.x : s_work;
        if (work >= n_work) break;
        // (a shared chain's work item is always the block index: chains x parts never exceeds the grid)
        const int chain = !kShared ? work : (n_chains > 1 ? shared_chain : 0);
        const int part = !kShared ? 0 : (int)blockIdx.x % nparts;
        // scratch region: the CTA's own, or -- when several CTAs share a chain -- the chain's
        uint64_t* tmpl = scratch + (size_t)(!kShared ? (int)blockIdx.x : chain) * region_stride;
        uint64_t* gbufB = tmpl + key_stride;
        uint64_t* gbufC = tmpl + 2 * key_stride;
        double* partials = reinterpret_cast<double*>(tmpl + 3 * key_stride);  // [nseg][2][kThreads]
        const SelectBuf bufs[4] = {{sm.candB, kCandB}, {sm.candA, kCandA}, {gbufB, (int)key_stride}, {gbufC, (int)key_stride}};
        {
            const double* src = reinterpret_cast<const double*>(cc_all + chain);
            double* dst = reinterpret_cast<double*>(&sm.cc);
            for (int i = tid; i < (int)(sizeof(ChainConst) / sizeof(double)); i += kThreads) dst[i] = src[i];
        }
        __syncthreads();
        const ChainConst& cc = sm.cc;
        const int flag = (int)cc.flag;
        const bool roche = flag & 1, nan_model = flag & 2;

        if (nan_model || (roche && lc_out == nullptr) || N <= 0) {
            // quirk Q13: the reference evaluates the model and then discards it on Roche overflow
            if (part == 0) {
                if (lc_out != nullptr)
                    for (int i = tid; i < N; i += kThreads) lc_out[(size_t)chain * N + i] = qnan;
                if (tid == 0 && logL != nullptr)
                    logL[chain] = roche ? -0.5 * kBig : (nan_model ? qnan : -0.5 * cc.chi2_extra);
            }
            __syncthreads();
            continue;
        }
        if (tid == 0 && part == 0 && evaluated != nullptr) atomicAdd(evaluated, 1ull);

        HB_T0();
        // ---- E(M) starter table for chains whose solve is path-independent where the table is used ----
        const bool use_table = (cc.e >= 0.0) && (cc.e <= kTableMaxE) && (N >= kTableMinN);
        const double2* ktab = use_table ? sm.ktab_head : nullptr;
        if (use_table) {
            build_kepler_table<kThreads>(sm.ktab_head, cc.e, sctab);
        }

        HB_T(0);
        // ---- pre-sample: bracket of the median rank + expansion point ----
        // small light curves fit the shared-memory candidate list whole: no bracket, no pre-sample
        const bool bracketed = N > kCandA / 2;
        uint64_t* cand = sm.candA;  // first-round survivors
        int cand_cap = kCandA;
        bool cand_small = true;  // the candidate list is the shared-memory one
        double lo = -INFINITY, hi = INFINITY;
        if (bracketed) {
            const uint32_t seed = (uint32_t)cc.seed;
            const double us = raw_flux1<true, true>(cc, ktab, sctab, tsec[sample_index(tid, kThreads, N, seed)]);
            // a NaN sample sorts above every number; the model pass flags NaN and aborts the chain
            const uint64_t sorted = block_sort<kThreads>(dkey(us), sm.ctl.xch);
            int r_lo, r_hi, r_mid;
            bracket_ranks(kThreads, N, krank, bracket_sigma, r_lo, r_hi, r_mid);
            const float frac = fminf(1.0f, (float)(r_hi - r_lo + 1) / (float)kThreads);
            // survivors go to global scratch (always when the E(M) table occupies candA)
            if (kShared || use_table || (int)(frac * (float)N * 1.5f) + 64 > kCandA) {
                cand = gbufB;
                cand_cap = (int)key_stride;
                cand_small = false;
            }
            if (tid == 0) {
                sm.ctl.cnt = 0;
                // a bracket rank outside the sample leaves that side open (the keys of -inf / +inf decode to
                // themselves; the integer sentinels 0 / ~0 would decode to NaN and reject every sample)
                if (r_lo < 0) sm.ctl.lo = dkey(-INFINITY);
                if (r_hi > kThreads - 1) sm.ctl.hi = dkey(INFINITY);
            }
            if (tid == r_lo) sm.ctl.lo = sorted;
            if (tid == r_hi) sm.ctl.hi = sorted;
            if (tid == r_mid) sm.pivot = dunkey(sorted);
            __syncthreads();
            lo = dunkey(sm.ctl.lo);
            hi = dunkey(sm.ctl.hi);
        } else {
            if (tid == 0) {
                sm.ctl.cnt = 0;
                sm.pivot = cc.K0;  // template values are K0 + O(1e-2)
            }
            __syncthreads();
        }
        HB_T(1);
        const double u0 = sm.pivot;
        const double A = cc.ft * (1.0 - cc.blend), ft = cc.ft;

        // ---- model pass: one sample per thread and iteration ----
        // The template keys go to global scratch only when something will read them back: light-curve
        // output, the small-N direct chi^2, or the re-run after a missed bracket (below).  The common
        // logL-only pass keeps everything on chip: no store, no address arithmetic for it.
        int nanflag = 0, c_lt = 0;
        double S0 = 0., S1 = 0.;
        constexpr int V = kPointsPerThread;
        static_assert(V == 1, "the segment bookkeeping of the pass is written for one sample per thread and iteration");
        int hi_acc = 0;  // largest sincos argument exponent of the hot pass (checked once, below)
        const int flag_pass = flag | (cc.tab_min_m > 0.0 ? 8 : 0);  // what the sample loop asks per sample, in a register
        ChainSync* const sync = sync_all + (!kShared ? 0 : chain);
        auto model_pass = [&](auto store_tag, auto data_tag, auto part_tag, auto lowe_tag) {
        constexpr bool kLowE = decltype(lowe_tag)::value;  // table starter at every sample (see kepler_points)
        constexpr bool kStore = decltype(store_tag)::value;
        constexpr bool kHot = !kStore;  // the logL-only pass defers the sincos range check to the end of the chain
        constexpr bool kData = decltype(data_tag)::value;  // fw != nullptr, known at compile time in the hot variant
        constexpr bool kPart = decltype(part_tag)::value;  // this CTA evaluates one part of a chain shared by nparts CTAs
        static_assert(!(kStore && kPart), "parts are logL-only");
        // the tiles of this pass: the whole light curve, or the part's run of segments
        int tile_first = 0, tile_last = n_tiles;
        if (kPart) {
            const int spp = (nseg + nparts - 1) / nparts;  // segments per part (the last parts may get fewer, or none)
            tile_first = min((part * spp) << seg_shift, n_tiles);
            tile_last = min(((part + 1) * spp) << seg_shift, n_tiles);
        }
        // running sums of a CTA that owns its chain: in shared memory (the select's work area is idle now),
        // so that the per-segment sums can restart without costing registers
        double* const tsum = reinterpret_cast<double*>(sm.candB);  // [2][kThreads]
        if (kData && !kPart) tsum[tid] = tsum[kThreads + tid] = 0.0;
        nanflag = 0;
        c_lt = 0;
        S0 = S1 = 0.;
        if (tile_first >= tile_last) return;
#if HB_TMA_STAGING
        // the data arrays are padded to whole tiles by the host side, so every copy is a full tile
        auto issue_tile = [&](int tile, uint32_t g) {
            const int st = g % kStages;
            const size_t off = (size_t)tile * kTile;
            mbar_expect_tx(&sm.full_bar[st], fw != nullptr ? 3 * tile_bytes : tile_bytes);
            bulk_g2s(sm.stage[st].ts, tsec + off, tile_bytes, &sm.full_bar[st]);
            if (fw != nullptr) bulk_g2s(sm.stage[st].fw, fw + off, 2 * tile_bytes, &sm.full_bar[st]);
        };
        if (tid == 0) issue_tile(tile_first, git);  // every stage is free here: the chain-level barriers drained the pipe
        for (int tile = tile_first; tile < tile_last; tile++, git++) {
            const int base = tile * kTile;
            const int st = git % kStages;
            if (tid == 0 && tile + 1 < tile_last) {
                // next tile into the other stage, once every warp has taken the previous tile out of it
                if (tile > tile_first) mbar_wait(&sm.empty_bar[(git + 1) % kStages], ((git - 1) / kStages) & 1);
                issue_tile(tile + 1, git + 1);
            }
            mbar_wait(&sm.full_bar[st], (git / kStages) & 1);
            int idx[V];
            double ts[V], u[V], fl[V], wv[V];
#pragma unroll
            for (int j = 0; j < V; j++) {
                idx[j] = base + j * kThreads + tid;
                ts[j] = sm.stage[st].ts[j * kThreads + tid];
                fl[j] = sm.stage[st].fw[j * kThreads + tid].x;
                wv[j] = sm.stage[st].fw[j * kThreads + tid].y;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty_bar[st]);
#else
        // software pipeline: the next iteration's time samples and this iteration's (flux, weight)
        // are requested from L2 before the ~2000-cycle model evaluation that hides their latency
        // (the arrays are padded to whole tiles, so no index clamp is needed)
        double ts_next[V];
#pragma unroll
        for (int j = 0; j < V; j++) ts_next[j] = tsec[tile_first * kTile + j * kThreads + tid];
        // this thread's first sample of the tile, carried in a register the compiler cannot re-derive
        // (it would otherwise rebuild it from the tile counter and SR_TID twice per iteration)
        int i0 = tile_first * kTile + tid;
        double u_last = 0.0;  // this thread's sample of the last tile of the pass
        for (int tile = tile_first; tile < tile_last; tile++, i0 += kTile) {
            asm volatile("" : "+r"(i0));
            int idx[V];
            double ts[V], u[V], fl[V], wv[V];
#pragma unroll
            for (int j = 0; j < V; j++) {
                idx[j] = i0 + j * kThreads;
                ts[j] = ts_next[j];
                HB_CHK(idx[j] + kTile, (size_t)(n_tiles + 1) * kTile, 6);
                ts_next[j] = tsec[idx[j] + kTile];  // (the array is padded with a whole tile beyond the last one)
                fl[j] = wv[j] = 0.0;
                if (kData) {  // one 16-byte load: {flux, 1/sigma} are interleaved
                    const double2 v = fw[idx[j]];
                    fl[j] = v.x;
                    wv[j] = v.y;
                }
            }
#endif
            raw_flux<V, true, true, kHot, kLowE>(cc, ktab, sctab, ts, u, &hi_acc, flag_pass);
#pragma unroll
            for (int j = 0; j < V; j++) {
                const int i = idx[j];
                const bool valid = i < N;
                const double uj = u[j];
                // with data, a NaN sample poisons S0 and is detected once after the loop
                if (!kData) nanflag |= valid & (uj != uj);  // the padding of a staged time grid may hold anything
                if (kStore && valid) {
                    HB_CHK(i, key_stride, 4);
                    tmpl[i] = dkey(uj);
                }
                // #(u < lo): one compare of the index, one FP64 compare taking it as a predicate input, one predicated
                // add (the compiler turns `if (valid & (uj < lo)) c_lt++` into a select and two moves)
                // The hot pass does not even ask whether the sample exists: the padded samples of the light curve's last
                // tile are taken out of the count behind the loop (two instructions per sample fewer).
#if !HB_TMA_STAGING
                if (kHot && kData) {
                    asm("{\n .reg .pred pl;\n setp.lt.f64 pl, %1, %2;\n @pl add.s32 %0, %0, 1;\n}" : "+r"(c_lt) : "d"(uj), "d"(lo));
                    u_last = uj;
                } else
#endif
                {
                    asm("{\n .reg .pred pv, pl;\n setp.lt.s32 pv, %3, %4;\n setp.lt.and.f64 pl, %1, %2, pv;\n @pl add.s32 %0, %0, 1;\n}"
                        : "+r"(c_lt) : "d"(uj), "d"(lo), "r"(i), "r"(N));
                }
                const bool inr = valid & (uj >= lo) & (uj <= hi);
                if (__any_sync(0xffffffffu, inr)) {  // (the vote goes straight to a predicate; the mask only where it is used)
                    const unsigned mask = __ballot_sync(0xffffffffu, inr);
                    const int leader = __ffs(mask) - 1;
                    int basepos = 0;
                    if (lane == leader) basepos = kPart ? atomicAdd(&sync->cnt, __popc(mask)) : atomicAdd(&sm.ctl.cnt, __popc(mask));
                    basepos = __shfl_sync(0xffffffffu, basepos, leader);
                    if (inr) {
                        const int pos = basepos + __popc(mask & ((1u << lane) - 1u));
                        // (the list in global scratch holds a whole light curve: only the shared-memory one can overflow)
                        if (!cand_small || pos < kCandA) {
                            HB_CHK(pos, cand_cap, 5);
                            cand[pos] = dkey(uj);
                        }
                    }
                }
                if (kData) {  // padded samples carry weight 0 and a finite model: they add exactly 0
                    const double wi = wv[j];
                    const double a = fma(A, uj - u0, ft - fl[j]);
                    const double r = a * wi;
                    S0 = fma(r, r, S0);
                    S1 = fma(r, wi, S1);
                }
            }
            // a segment (or the pass) ends with this tile: this thread's sums of the segment join the running
            // sums in segment order -- here, or (a part) through the chain's scratch at the hands of the finisher
            // (the segment's end is read off the tile counter against seg_mask, a constant-bank operand)
            if (kData && (((tile + 1) & seg_mask) == 0)) {  // (a last, partial segment is handed over behind the loop)
                if (kPart) {
                    const int seg = tile >> seg_shift;
                    HB_CHK(seg, nseg, 8);
                    partials[(size_t)(2 * seg) * kThreads + tid] = S0;
                    partials[(size_t)(2 * seg + 1) * kThreads + tid] = S1;
                } else {
                    tsum[tid] += S0;
                    tsum[kThreads + tid] += S1;
                }
                S0 = S1 = 0.;
            }
        }
#if !HB_TMA_STAGING
        if (kHot && kData && i0 - kTile >= N && u_last < lo) c_lt--;  // a padded sample that was counted (see the loop)
#endif
        if (kData && (tile_last & seg_mask) != 0) {  // the light curve's last segment is a partial one
            if (kPart) {
                const int seg = (tile_last - 1) >> seg_shift;
                HB_CHK(seg, nseg, 8);
                partials[(size_t)(2 * seg) * kThreads + tid] = S0;
                partials[(size_t)(2 * seg + 1) * kThreads + tid] = S1;
            } else {
                tsum[tid] += S0;
                tsum[kThreads + tid] += S1;
            }
            S0 = S1 = 0.;
        }
        };
        // instantiations: the hot logL-only pass (whole chain / one part), and a general one (template stored) for
        // light-curve output, small N and the re-run after a missed bracket
        const bool store_template = (lc_out != nullptr) || !bracketed || (fw == nullptr);
        auto general_pass = [&]() {  // the whole chain, by this CTA alone
            if (tid == 0) sm.ctl.cnt = 0;
            __syncthreads();
            if (fw != nullptr) model_pass(std::true_type{}, std::true_type{}, std::false_type{}, std::false_type{});
            else model_pass(std::true_type{}, std::false_type{}, std::false_type{}, std::false_type{});
        };
        auto load_sums = [&](bool from_parts) {  // this thread's chi^2 sums: its segment sums added in segment order
            if (fw == nullptr) return;
            if (!from_parts) {
                const double* tsum = reinterpret_cast<const double*>(sm.candB);
                S0 = tsum[tid];
                S1 = tsum[kThreads + tid];
            } else {  // written by the other CTAs of the chain: read from L2, eight loads in flight
                S0 = S1 = 0.;
                for (int seg = 0; seg < nseg; seg += 4) {
                    double p0[4], p1[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int sg = seg + k < nseg ? seg + k : nseg - 1;
                        p0[k] = __ldcg(&partials[(size_t)(2 * sg) * kThreads + tid]);
                        p1[k] = __ldcg(&partials[(size_t)(2 * sg + 1) * kThreads + tid]);
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (seg + k < nseg) {
                            S0 += p0[k];
                            S1 += p1[k];
                        }
                }
            }
            nanflag = (S0 != S0);
        };
        bool have_template = store_template;
        int c_in;
        if (store_template) {
            general_pass();
            c_lt = block_sum_int<kThreads>(c_lt, sm.ctl.ired);
            c_in = sm.ctl.cnt;
            load_sums(false);
        } else {
            // (a second copy of the sample loop for chains that take the table starter everywhere -- 96 % of prior
            // draws -- without the starter choice and the argument tracking: 2 % on light curves of 20-50 k points,
            // 1 % on 200 k points)
            if constexpr (!kShared) {
                if (use_table && !(cc.tab_min_m > 0.0))
                    model_pass(std::false_type{}, std::true_type{}, std::false_type{}, std::true_type{});
                else
                    model_pass(std::false_type{}, std::true_type{}, std::false_type{}, std::false_type{});
            } else {
                model_pass(std::false_type{}, std::true_type{}, std::true_type{}, std::false_type{});
            }
            // a Newton iterate left the table sincos' range somewhere in this chain (e -> 1 only): the sums are
            // not trustworthy; the chain is evaluated again with the per-sample check and the library fallback
            HB_T(2);
            int redo = __syncthreads_or(hi_acc > hot_hi_limit);
            c_lt = block_sum_int<kThreads>(c_lt, sm.ctl.ired);
            if constexpr (kShared) {
                // hand-over: this part's counts and flags to the chain's sync words; the last part to arrive
                // finishes the chain.  Every thread's stores (segment sums, candidate keys) are fenced before the
                // barrier in front of thread 0's ticket.
                __threadfence();
                __syncthreads();
                if (tid == 0) {
                    atomicAdd(&sync->c_lt, c_lt);
                    if (redo) atomicOr(&sync->flags, 1);
                    __threadfence();
                    const int ticket = atomicAdd(&sync->done, 1);
                    sm.bcast[0] = (ticket == nparts - 1);
                    if (ticket == nparts - 1) {
                        __threadfence();
                        sm.bcast[1] = atomicAdd(&sync->c_lt, 0);
                        sm.bcast[2] = atomicAdd(&sync->cnt, 0);
                        sm.bcast[3] = atomicAdd(&sync->flags, 0);
                        sync->cnt = 0; sync->c_lt = 0; sync->done = 0; sync->flags = 0;  // ready for the next launch
                    }
                }
                __syncthreads();
                if (!sm.bcast[0]) continue;  // not the last: on to the next work item
                __threadfence();
                c_lt = sm.bcast[1];
                c_in = sm.bcast[2];
                redo = sm.bcast[3];
            } else {
                c_in = sm.ctl.cnt;
            }
            if (redo) {
                general_pass();
                have_template = true;
                c_lt = block_sum_int<kThreads>(c_lt, sm.ctl.ired);
                c_in = sm.ctl.cnt;
                load_sums(false);
            } else {
                load_sums(kShared);
            }
        }
        const int any_nan = __syncthreads_or(nanflag);  // (also the barrier between reading cnt and the select's reset of it)
        if (any_nan) {
            if (lc_out != nullptr)
                for (int i = tid; i < N; i += kThreads) lc_out[(size_t)chain * N + i] = qnan;
            if (tid == 0 && logL != nullptr) logL[chain] = roche ? -0.5 * kBig : qnan;
            __syncthreads();
            continue;
        }

        HB_T(3);
        // ---- exact order statistic ----
        const uint32_t seed2 = (uint32_t)cc.seed ^ 0x9e3779b9u;
        uint64_t mkey;
        if (krank >= c_lt && krank < c_lt + c_in && c_in <= cand_cap) {
            // the candidates lie between two order statistics of the pre-sample: one histogram pass finds the
            // bin of the median (hb_select.cuh); sampling rounds only when that fails (ties) or there is no bracket
            if (!(bracketed && c_in > kThreads &&
                  block_select_hist<kThreads>(cand, c_in, krank - c_lt, lo, hi, sm.ctl, sm.candB, &mkey))) {
                mkey = block_select_key<kThreads>(cand, c_in, krank - c_lt, sm.ctl, bufs, 4, seed2);  // (skips the buffer it reads)
            }
        }
        else {
            // the bracket missed (or overflowed; ~1 % of chains by construction of the 2.5 sigma bracket):
            // evaluate the chain once more, this time storing the template, and select on that
            if (!have_template) {
                general_pass();
                __syncthreads();
                load_sums(false);
            }
            mkey = block_select_key<kThreads>(tmpl, N, krank, sm.ctl, bufs, 4, seed2);
        }
        HB_T(4);
        const double med = dunkey(mkey);

        // ---- results ----
        if (lc_out != nullptr) {
            const double blend = cc.blend;
            for (int i = tid; i < N; i += kThreads)
                lc_out[(size_t)chain * N + i] = finish_template(dunkey(tmpl[i]), med, blend, ft);
        }
        if (logL != nullptr) {
            if (!bracketed) {
                // small light curve: no sample-based pivot exists, so chi^2 is summed directly from the
                // stored template in the reference's own form (likelihood3.c:681-685, 829-830)
                S0 = 0.;
                const double blend = cc.blend;
                if (fw != nullptr)
                    for (int i = tid; i < N; i += kThreads) {
                        const double2 v = fw[i];
                        const double r = (finish_template(dunkey(tmpl[i]), med, blend, ft) - v.x) * v.y;
                        S0 = fma(r, r, S0);
                    }
                S1 = 0.;
            }
            // two block sums in one go
            const int wid = tid >> 5;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                S0 += __shfl_xor_sync(0xffffffffu, S0, o);
                S1 += __shfl_xor_sync(0xffffffffu, S1, o);
            }
            if (lane == 0) {
                sm.red[wid] = S0;
                sm.red[32 + wid] = S1;
            }
            __syncthreads();
            if (tid == 0) {
                double t0 = 0., t1 = 0.;
                for (int i = 0; i < kThreads / 32; i++) {
                    t0 += sm.red[i];
                    t1 += sm.red[32 + i];
                }
                const double d = -A * (med - u0);
                const double chi2 = bracketed ? t0 + d * (2.0 * t1 + d * sum_w2) : t0;
                logL[chain] = roche ? -0.5 * kBig : -0.5 * (chi2 + cc.chi2_extra);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// small entry points
// ---------------------------------------------------------------------------

// k-th order statistic of x[0..n) with one CTA (remove_median's sort, likelihood3.c:86-105,
// as a stand-alone call).  out[0] = value, out[1] = 1 when x holds a NaN (value = NaN then).
template <int kThreads>
__global__ void __launch_bounds__(kThreads, kEvalCtasPerSm)
k_order_stat(const double* __restrict__ x, int n, int k, uint64_t* __restrict__ scratch, size_t stride,
             double* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EvalShared& sm = *reinterpret_cast<EvalShared*>(smem_raw);
    SelectBuf bufs[4] = {{sm.candB, kCandB}, {sm.candA, kCandA}, {scratch + stride, n}, {scratch + 2 * stride, n}};
    int nanflag = 0;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const double v = x[i];
        nanflag |= (v != v);
        scratch[i] = dkey(v);
    }
    const int any_nan = __syncthreads_or(nanflag);
    if (any_nan) {
        if (threadIdx.x == 0) { out[0] = __longlong_as_double(0x7ff8000000000000LL); out[1] = 1.0; }
        return;
    }
    const double v = dunkey(block_select_key<kThreads>(scratch, n, k, sm.ctl, bufs, 4, 0x1234567u));
    if (threadIdx.x == 0) { out[0] = v; out[1] = 0.0; }
}

// arr[i] -= *value (remove_median's subtraction, likelihood3.c:103); value lives on the device
__global__ void k_subtract(double* __restrict__ arr, int n, const double* __restrict__ value)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) arr[i] -= value[0];
}

// traj() of likelihood3.c:125-185 for one parameter set: d, Z1, Z2, r [cm], nu [rad].
// tp = {M1, M2 [g], P [s], e, inc, omega0, T0 [s]}.  nu is reported through atan2 of the
// algebraic sin/cos nu, which equals 2 atan(sqrt((1+e)/(1-e)) tan(E/2)) on (-pi, pi).
__global__ void k_traj(const double* __restrict__ times, int Nt, const double* __restrict__ tp,
                       double* __restrict__ d_arr, double* __restrict__ Z1, double* __restrict__ Z2,
                       double* __restrict__ rr, double* __restrict__ ff)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nt) return;
    double Ma = tp[0], Mb = tp[1];
    if (Mb > Ma) { double s = Ma; Ma = Mb; Mb = s; }
    const double P = tp[2], e = tp[3], inc = tp[4], w0 = tp[5], T0 = tp[6];
    const double Mtot = Ma + Mb;
    const double a = pow(kG * Mtot * sq(P) / sq(2 * kPi), 1. / 3.);
    const double ts[1] = {__dmul_rn(times[i], kSecDay)};
    double cE[1], sE[1], den[1], bet[1];
    kepler_points<1, false>(ts, e, T0, P, __drcp_rn(P), nullptr, 0.0, nullptr, cE, sE, den, bet);
    const double r = a * den[0];
    const double sq1 = sqrt(1 - e * e);
    const double nu = atan2(sq1 * sE[0], cE[0] - e);
    double sw, cw, si, ci;
    sincos(w0 + nu, &sw, &cw);
    sincos(inc, &si, &ci);
    const double ZZ = r * sw * si;
    rr[i] = r;
    ff[i] = nu;
    d_arr[i] = r * sqrt(cw * cw + sq(sw * ci));
    Z1[i] = ZZ * (Mb / Mtot);
    Z2[i] = -ZZ * (Ma / Mtot);
}

// op codes of hb_scalar (hb_b200.h)
__global__ void k_scalar(int op, const double* __restrict__ a, double* __restrict__ out)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double r = 0.;
    switch (op) {
        case 0: r = dev_getT(a[0]); break;
        case 1: r = dev_getR(a[0]); break;
        case 2: r = dev_envelope_temp(a[0]); break;
        case 3: r = dev_envelope_radius(a[0]); break;
        case 4: r = dev_alpha_beam(a[0]); break;
        case 5: {  // eclipse_area(R1, R2, d[cm])
            double R1 = a[0], R2 = a[1];
            if (R2 > R1) { double s = R1; R1 = R2; R2 = s; }
            const double d = fabs(a[2]) / kRsun;
            r = (d >= R1 + R2) ? 0. : eclipse_area_dev(R1, R2, d);
            break;
        }
        case 6: {  // beaming(P, M1, M2, e, inc, omega0, nu, alpha_beam), likelihood3.c:224-236
            const double q = a[2] / a[1];
            r = -2830. * a[7] * q * pow(a[1], 1. / 3) * pow(a[0], -1. / 3) * (sin(a[4]) * cos(a[5] + a[6]) / sqrt(1 - sq(a[3]))) * 1.e-6;
            break;
        }
        case 7: {  // ellipsoidal(P, M1, M2, e, inc, omega0, nu, R1, a, mu, tau), likelihood3.c:255-307
            const double P = a[0], M1 = a[1], M2 = a[2], e = a[3], inc = a[4], x = a[5] + a[6], nu = a[6], R1 = a[7];
            const double mu = a[9], tau = a[10];
            const double al11 = 15 * mu * (2 + tau) / (32 * (3 - mu));
            const double al21 = 3 * (15 + mu) * (1 + tau) / (20 * (3 - mu));
            const double al2b1 = 15 * (1 - mu) * (3 + tau) / (64 * (3 - mu));
            const double al01 = al21 / 9, al0b1 = 3 * al2b1 / 20, al31 = 5 * al11 / 3, al41 = 7 * al2b1 / 4;
            const double beta = (1 + e * cos(nu)) / (1 - sq(e));
            const double q = M2 / M1, Prot = P * pow(1 - e, 3. / 2);
            const double si = sin(inc), si2 = si * si, bR = beta * R1;
            const double o3 = 13435. / M1 * q / (1 + q) / sq(P) * bR * bR * bR;
            const double o5 = 759. * pow(M1, -5. / 3) * q / pow(1 + q, 5. / 3) * pow(P, -10. / 3) * pow(bR, 5);
            const double o4 = 3194. * pow(M1, -4. / 3) * q / pow(1 + q, 4. / 3) * pow(P, -8. / 3) * sq(sq(bR));
            r = (13435. * 2 * al01 * (2 - 3 * si2) / M1 / sq(Prot) * R1 * R1 * R1 + o3 * 3 * al01 * (2 - 3 * si2) +
                 o3 * al21 * si2 * cos(2 * x) + o5 * al0b1 * (8 - 40 * si2 + 35 * si2 * si2) +
                 o4 * al11 * (4 * si - 5 * si2 * si) * sin(x) + o5 * al2b1 * (6 * si2 - 7 * si2 * si2) * cos(2 * x) +
                 o4 * al31 * si2 * si * sin(3 * x) + o5 * al41 * si2 * si2 * cos(4 * x)) * 1.e-6;
            break;
        }
        case 8: {  // reflection(P, M1, M2, e, inc, omega0, nu, R2, alpha_ref), likelihood3.c:322-337
            const double q = a[2] / a[1], x = a[5] + a[6], si = sin(a[4]);
            const double beta = (1 + a[3] * cos(a[6])) / (1 - sq(a[3]));
            r = 56514. * a[8] * pow(1 + q, -2. / 3) * pow(a[1], -2. / 3) * pow(a[0], -4. / 3) * sq(beta * a[7]) *
                (0.64 - si * sin(x) + 0.18 * si * si * (1 - cos(2 * x))) * 1.e-6;
            break;
        }
        default: r = __longlong_as_double(0x7ff8000000000000LL);
    }
    out[0] = r;
}

// GAIA_mcmc.c:198-269: magnitudes + 4-term chi^2 for the 6-parameter layout, one thread per row.
__global__ void k_gaia(const double* __restrict__ p6, int n, double D, const double* __restrict__ data,
                       const double* __restrict__ err, double* __restrict__ mags_out, double* __restrict__ logL)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const double* p = p6 + (size_t)c * 6;
    double m[4];
    gaia_mags(p, D, m);
    if (mags_out != nullptr)
        for (int i = 0; i < 4; i++) mags_out[(size_t)c * 4 + i] = m[i];
    if (logL != nullptr) {
        double chi2 = 0.;
        for (int i = 0; i < 4; i++) {
            const double r = (data[i] - m[i]) / err[i];
            chi2 += r * r;
        }
        logL[c] = -chi2 / 2.0;
    }
}

// chain diagnostics: out[c][0..8] = R1 R2 T1 T2 G B-V V-G G-T roche
__global__ void k_chain_info(const ChainConst* __restrict__ cc, int n, double* __restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    for (int i = 0; i < 8; i++) out[(size_t)c * 9 + i] = cc[c].info[i];
    out[(size_t)c * 9 + 8] = (double)(((int)cc[c].flag) & 1);
}

// tsec[i] = t[i] * 86400 (likelihood3.c:149), once per uploaded time grid
__global__ void k_to_seconds(const double* __restrict__ t, int n, double* __restrict__ tsec)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tsec[i] = __dmul_rn(t[i], kSecDay);
}

// DFMA throughput probe: 8 independent accumulators per thread, `iters` x 8 x 4 DFMA each.
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b)
{
    double x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
            x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;  // keeps the chain live
}

// ---------------------------------------------------------------------------
// launchers (plain C++ so that hb_capi.cu stays free of <<< >>>)
// ---------------------------------------------------------------------------
cudaError_t launch_prologue(const double* params, int n, const MagSetup& ms, ChainConst* out, int* eval_counter,
                            int eval_grid, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    if (n <= kPrologueSmallMax) k_prologue_small<<<n, 128, 0, s>>>(params, n, ms, out, eval_counter, eval_grid);
    else k_prologue<<<(n + 3) / 4, 128, 0, s>>>(params, n, ms, out, eval_counter, eval_grid);
    return cudaGetLastError();
}

cudaError_t configure_eval()
{
    cudaError_t e = cudaFuncSetAttribute(k_chain_eval<kEvalThreads, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(EvalShared));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_chain_eval<kEvalThreads, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(EvalShared));
}

// Segments of the time axis: runs of 2^shift tiles (at least 4 tiles; doubled until at most kMaxSegments segments
// cover the light curve) -- a function of the light curve's length alone; the summation order of the chi^2 follows it.
int eval_seg_shift(long n_points)
{
    const long n_tiles = (n_points + kTile - 1) / kTile;
    int shift = 2;
    while (((n_tiles + (1L << shift) - 1) >> shift) > kMaxSegments) shift++;
    return shift;
}
int eval_segments(long n_points)
{
    const long n_tiles = std::max(1L, (n_points + kTile - 1) / kTile);
    const int shift = eval_seg_shift(n_points);
    return (int)((n_tiles + (1L << shift) - 1) >> shift);
}

cudaError_t launch_chain_eval(const EvalArgs& a, int grid, cudaStream_t s)
{
    if (a.n_chains <= 0) return cudaSuccess;
    // *counter was set to this grid by the k_prologue launch in front of this one
    if (a.nparts > 1)
        k_chain_eval<kEvalThreads, true><<<grid, kEvalThreads, sizeof(EvalShared), s>>>(
            a.cc, a.n_chains, a.tsec, a.fw, a.N, a.scratch, a.region_stride, a.key_stride, a.logL, a.lc_out, a.counter, a.bracket_sigma,
            a.sctab, a.hot_hi_limit, a.sum_w2, a.evaluated, a.sync, a.nparts, a.nseg, a.seg_shift, a.max_parts, (1 << a.seg_shift) - 1);
    else
        k_chain_eval<kEvalThreads, false><<<grid, kEvalThreads, sizeof(EvalShared), s>>>(
            a.cc, a.n_chains, a.tsec, a.fw, a.N, a.scratch, a.region_stride, a.key_stride, a.logL, a.lc_out, a.counter, a.bracket_sigma,
            a.sctab, a.hot_hi_limit, a.sum_w2, a.evaluated, a.sync, 1, a.nseg, a.seg_shift, a.max_parts, (1 << a.seg_shift) - 1);
    return cudaGetLastError();
}

cudaError_t launch_order_stat(const double* x, int n, int k, uint64_t* scratch, size_t stride, double* out, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(k_order_stat<kEvalThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(EvalShared));
    if (e != cudaSuccess) return e;
    k_order_stat<kEvalThreads><<<1, kEvalThreads, sizeof(EvalShared), s>>>(x, n, k, scratch, stride, out);
    return cudaGetLastError();
}

cudaError_t launch_subtract(double* arr, int n, const double* value, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_subtract<<<(n + 255) / 256, 256, 0, s>>>(arr, n, value);
    return cudaGetLastError();
}

cudaError_t launch_traj(const double* times, int Nt, const double* tp, double* d, double* Z1, double* Z2, double* rr,
                        double* ff, cudaStream_t s)
{
    if (Nt <= 0) return cudaSuccess;
    k_traj<<<(Nt + 127) / 128, 128, 0, s>>>(times, Nt, tp, d, Z1, Z2, rr, ff);
    return cudaGetLastError();
}

cudaError_t launch_scalar(int op, const double* args, double* out, cudaStream_t s)
{
    k_scalar<<<1, 32, 0, s>>>(op, args, out);
    return cudaGetLastError();
}

cudaError_t launch_gaia(const double* p6, int n, double D, const double* data, const double* err, double* mags,
                        double* logL, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_gaia<<<(n + 127) / 128, 128, 0, s>>>(p6, n, D, data, err, mags, logL);
    return cudaGetLastError();
}

cudaError_t launch_chain_info(const ChainConst* cc, int n, double* out, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_chain_info<<<(n + 127) / 128, 128, 0, s>>>(cc, n, out);
    return cudaGetLastError();
}

cudaError_t launch_to_seconds(const double* t, int n, double* tsec, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_to_seconds<<<(n + 255) / 256, 256, 0, s>>>(t, n, tsec);
    return cudaGetLastError();
}

#ifdef HB_DEBUG_BOUNDS
// self-test of the bounds-asserting build: index `i` against capacity 4 (traps when i >= 4)
__global__ void k_bounds_selftest(int i, int* out)
{
    HB_CHK(i, 4, 99);
    out[0] = i;
}
cudaError_t launch_bounds_selftest(int i, int* out, cudaStream_t s)
{
    k_bounds_selftest<<<1, 1, 0, s>>>(i, out);
    return cudaGetLastError();
}
#endif

cudaError_t launch_fp64_peak(double* out, int blocks, int iters, cudaStream_t s)
{
    k_fp64_peak<<<blocks, 256, 0, s>>>(out, iters, 1.0000001, 0.9999999);
    return cudaGetLastError();
}

}  // namespace hb
