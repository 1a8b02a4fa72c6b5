// hb_select.cuh -- block-wide EXACT k-th order statistic of n doubles (sm_100a).
//
// Replaces the reference's copy + Lomuto quicksort + sorted[mid] of remove_median()
// (likelihood3.c:36-105).  The reference needs one order statistic, not a sorted array, so
// the device version is a sampling select (Floyd-Rivest style) on order-preserving integer keys:
//   round:  S = blockDim jittered-stride samples (one per thread) -> bitonic sort across the
//           block (shuffles below stride 32, shared memory above) -> two sample order statistics
//           lo/hi bracket the target rank -> one pass counts keys < lo and compacts keys in
//           [lo, hi] into a smaller buffer (warp-aggregated append).
//   finish: <= blockDim survivors are sorted directly; sorted[k] wins.
//   ties / bad luck: a round that does not shrink the problem is retried with a new jitter and a
//           wider bracket; after kMaxFails misses a 64-step bisection on the key bits (tie-proof).
// k_chain_eval runs the FIRST round inside its model pass (the bracket comes from a pre-sample of
// the model itself), so the common case never re-reads the template.
// Nothing here touches the FP64 pipe that the model pass saturates.  NaN-free input is a
// precondition (callers short-circuit NaN templates).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Bounds-asserting debug build (-DHB_DEBUG_BOUNDS, libhb_b200_dbg.so): every indexed access of the likelihood path
// -- candidate lists, template keys, the E(M) and sin/cos tables, select buffers, histogram bins, partial sums,
// the padded data arrays, the sampler's history rings -- is checked against its capacity and TRAPS with the site
// number; tests/test_gpu_debug_bounds.py runs the edge-size and select stress sweeps under it (compute-sanitizer
// is not available on the pool).  In the normal build the macro is empty.
#ifdef HB_DEBUG_BOUNDS
#include <cstdio>
static __device__ __noinline__ void hb_bounds_trap(long long i, long long cap, int site)
{
    printf("hb_b200 bounds violation: site %d index %lld capacity %lld (block %d thread %d)\n", site, i, cap, (int)blockIdx.x,
           (int)threadIdx.x);
    __trap();
}
#define HB_CHK(i, cap, site)                                                                               \
    do {                                                                                                     \
        if ((unsigned long long)(long long)(i) >= (unsigned long long)(long long)(cap)) hb_bounds_trap((long long)(i), (long long)(cap), site); \
    } while (0)
#else
#define HB_CHK(i, cap, site) ((void)0)
#endif

namespace hb {

constexpr int kMaxFails = 4;  // unsuccessful rounds before the bisection fallback

__device__ __forceinline__ uint64_t dkey(double x)
{
    uint64_t b = (uint64_t)__double_as_longlong(x);
    return b ^ ((uint64_t)((int64_t)b >> 63) | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(uint64_t k)
{
    uint64_t b = (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// index of sample j of S in [0, n): stratified with a hashed jitter (defeats aliasing of the
// stride with the orbital period)
__device__ __forceinline__ int sample_index(int j, int S, int n, uint32_t seed)
{
    const uint32_t h = mix32(seed ^ mix32((uint32_t)j + 0x9e3779b9u));
    const float jit = (float)(h >> 8) * (1.0f / 16777216.0f);
    long long idx = (long long)(((double)j + (double)jit) * (double)n / (double)S);
    return (int)(idx > n - 1 ? n - 1 : idx);
}

struct SelectBuf {
    uint64_t* ptr;
    int cap;
};

// Shared-memory control block of the select (the key buffers are passed separately).
template <int kThreads>
struct SelectCtl {
    uint64_t xch[kThreads];  // exchange buffer of the sort / sorted sample
    uint64_t lo, hi, result;
    int cnt;
    int ired[32];
};

template <int kThreads>
__device__ __forceinline__ int block_sum_int(int v, int* red)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();  // protects red[] against the previous use
    if (lane == 0) red[wid] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; i++) t += red[i];
    return t;
}

// Bitonic sort of one key per thread across the block: afterwards thread t holds sorted[t].
// Strides below 32 exchange through shuffles, the others through xch[] (kThreads entries).
template <int kThreads>
__device__ __noinline__ uint64_t block_sort(uint64_t key, uint64_t* xch)
{
    static_assert((kThreads & (kThreads - 1)) == 0, "block size must be a power of two");
    const int tid = threadIdx.x;
    // fully unrolled: k and j are compile-time constants in each of the log2(n)(log2(n)+1)/2 stages,
    // so a stage is two shuffles (or one shared-memory exchange), a 64-bit compare and two selects
#pragma unroll
    for (int k = 2; k <= kThreads; k <<= 1) {
        const bool descending = (tid & k) != 0;
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            uint64_t other;
            if (j >= 32) {
                xch[tid] = key;
                __syncthreads();
                HB_CHK(tid ^ j, kThreads, 20);
                other = xch[tid ^ j];
                __syncthreads();
            } else {
                const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)key, j);
                const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(key >> 32), j);
                other = ((uint64_t)hi << 32) | lo;
            }
            // the lower partner of an ascending pair keeps the minimum
            const bool take_min = descending == ((tid & j) != 0);
            const bool less = other < key;
            key = (less == take_min) ? other : key;
        }
    }
    return key;
}

// Bitonic sort of TWO keys per thread (elements tid and kThreads + tid of a 2 kThreads array): afterwards thread t
// gets sorted[t] and sorted[kThreads + t] back.  Both halves go through the network of block_sort side by
// side (the second half descending at the last level), then one in-thread exchange and a merge: a light curve of up
// to 2 kThreads points is sorted whole in ~3 k cycles, where sampling rounds take two full sorts and a compaction.
// xch2 holds 2 kThreads keys.
struct KeyPair { uint64_t a, b; };
template <int kThreads>
__device__ __noinline__ KeyPair block_sort2(uint64_t a, uint64_t b, uint64_t* xch2)
{
    static_assert((kThreads & (kThreads - 1)) == 0, "block size must be a power of two");
    const int tid = threadIdx.x;
    auto exchange = [&](uint64_t key, int j, uint64_t* xch) -> uint64_t {  // the key of thread tid ^ j (j < 32: shuffles)
        const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)key, j);
        const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(key >> 32), j);
        return ((uint64_t)hi << 32) | lo;
    };
#pragma unroll
    for (int k = 2; k <= 2 * kThreads; k <<= 1) {
        // element index of a is tid, of b kThreads + tid: the direction bit of b differs from a's at k == kThreads only
        const bool desc_a = (k < 2 * kThreads) && ((tid & k) != 0);
        const bool desc_b = (k < kThreads) ? desc_a : (k == kThreads);
        if (k == 2 * kThreads) {  // partner distance kThreads: in the thread (ascending)
            const uint64_t mn = a < b ? a : b, mx = a < b ? b : a;
            a = mn;
            b = mx;
        }
#pragma unroll
        for (int j = (k == 2 * kThreads ? kThreads : k) >> 1; j > 0; j >>= 1) {
            uint64_t oa, ob;
            if (j >= 32) {
                xch2[tid] = a;
                xch2[kThreads + tid] = b;
                __syncthreads();
                HB_CHK(tid ^ j, kThreads, 21);
                oa = xch2[tid ^ j];
                ob = xch2[kThreads + (tid ^ j)];
                __syncthreads();
            } else {
                oa = exchange(a, j, xch2);
                ob = exchange(b, j, xch2);
            }
            const bool upper = (tid & j) != 0;
            // the lower partner of an ascending pair keeps the minimum
            a = ((oa < a) == (desc_a == upper)) ? oa : a;
            b = ((ob < b) == (desc_b == upper)) ? ob : b;
        }
    }
    return KeyPair{a, b};
}

// Bracket ranks (in a sorted sample of S) around the target quantile q = (k + 0.5) / n.
__device__ __forceinline__ void bracket_ranks(int S, int n, int k, float z, int& r_lo, int& r_hi, int& r_mid)
{
    const float q = ((float)k + 0.5f) / (float)n;
    const float ks = q * (float)S;
    int m = (int)ceilf(z * sqrtf((float)S * q * (1.0f - q))) + 1;
    if (m < 2) m = 2;
    r_lo = (int)floorf(ks) - m;
    r_hi = (int)ceilf(ks) + m;
    r_mid = (int)ks;
    if (r_mid > S - 1) r_mid = S - 1;
}

// Tie-proof fallback: smallest key K with #(keys <= K) >= k+1, by bisection on the bits.
template <int kThreads>
__device__ __noinline__ uint64_t select_bisect(const uint64_t* keys, int n, int k, int* ired)
{
    uint64_t lo = 0, hi = ~0ull;
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        int c = 0;
        for (int i = threadIdx.x; i < n; i += kThreads) c += (keys[i] <= mid);
        c = block_sum_int<kThreads>(c, ired);
        if (c >= k + 1) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// One-shot select for keys known to lie in [lo, hi] with a roughly uniform density -- the candidates the
// model pass of k_chain_eval collects between two order statistics of its pre-sample.  Instead of two
// more sample / sort / compact rounds: histogram the keys over kThreads equal bins of [lo, hi] (pass A),
// find the bin holding rank k by a block scan, gather that bin (pass B) and sort it.  Both passes read
// the keys eight loads at a time.  Returns false -- nothing decided, the caller falls back to
// block_select_key -- when the bin holds more than kThreads keys (ties, a plateau) or [lo, hi] is degenerate.
// `work` is scratch shared memory of at least 2 * kThreads 64-bit words.
template <int kThreads>
__device__ __noinline__ bool block_select_hist(const uint64_t* keys, int n, int k, double lo, double hi,
                                               SelectCtl<kThreads>& ctl, uint64_t* work, uint64_t* result)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int* hist = reinterpret_cast<int*>(work);  // kThreads bins
    uint64_t* surv = work + kThreads;          // kThreads keys
    const double scale = (double)kThreads / (hi - lo);
    if (!(scale > 0.0) || !(scale < 1e300)) return false;  // hi == lo, or a non-finite bound (uniform decision)
    hist[tid] = 0;
    if (tid == 0) ctl.cnt = 0;
    __syncthreads();
    constexpr int kBatch = 16;  // one batch covers 16 * kThreads keys: the usual candidate count, read ONCE
    auto bin_of = [&](uint64_t x) {
        const int b = (int)((dunkey(x) - lo) * scale);  // monotone in x; the same value in both passes
        return min(max(b, 0), kThreads - 1);
    };
    const bool one_batch = n <= kBatch * kThreads;
    uint64_t xs[kBatch];
    for (int base = 0; base < n; base += kBatch * kThreads) {  // pass A: histogram
#pragma unroll
        for (int u = 0; u < kBatch; u++) {
            const int i = base + u * kThreads + tid;
            xs[u] = (i < n) ? keys[i] : 0ull;
        }
#pragma unroll
        for (int u = 0; u < kBatch; u++)
            if (base + u * kThreads + tid < n) {
                HB_CHK(bin_of(xs[u]), kThreads, 21);
                atomicAdd(&hist[bin_of(xs[u])], 1);
            }
    }
    __syncthreads();
    // block-wide inclusive scan of the bins (one per thread)
    const int h = hist[tid];
    int incl = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) ctl.ired[wid] = incl;
    __syncthreads();
    int offset = 0;
    for (int w = 0; w < wid; w++) offset += ctl.ired[w];
    incl += offset;
    const int excl = incl - h;
    if (excl <= k && k < incl) {  // exactly one thread: its bin holds rank k
        ctl.result = (uint64_t)tid;
        ctl.lo = (uint64_t)excl;
        ctl.hi = (uint64_t)h;
    }
    __syncthreads();
    const int bin = (int)ctl.result, below = (int)ctl.lo, count = (int)ctl.hi;
    if (count > kThreads) return false;
    // pass B: gather the bin (from the registers of pass A when one batch held everything)
    for (int base = 0; base < n; base += kBatch * kThreads) {
        if (!one_batch) {
#pragma unroll
            for (int u = 0; u < kBatch; u++) {
                const int i = base + u * kThreads + tid;
                xs[u] = (i < n) ? keys[i] : 0ull;
            }
        }
#pragma unroll
        for (int u = 0; u < kBatch; u++)
            if (base + u * kThreads + tid < n && bin_of(xs[u]) == bin) {
                const int slot = atomicAdd(&ctl.cnt, 1);
                HB_CHK(slot, kThreads, 22);
                surv[slot] = xs[u];
            }
    }
    __syncthreads();
    if (count <= 32) {
        // a handful of keys (the usual case): warp 0 sorts them with shuffles alone
        if (wid == 0) {
            uint64_t key = (lane < count) ? surv[lane] : ~0ull;
#pragma unroll
            for (int kk = 2; kk <= 32; kk <<= 1) {
                const bool descending = (lane & kk) != 0;
#pragma unroll
                for (int j = kk >> 1; j > 0; j >>= 1) {
                    const uint32_t olo = __shfl_xor_sync(0xffffffffu, (uint32_t)key, j);
                    const uint32_t ohi = __shfl_xor_sync(0xffffffffu, (uint32_t)(key >> 32), j);
                    const uint64_t other = ((uint64_t)ohi << 32) | olo;
                    const bool take_min = descending == ((lane & j) != 0);
                    key = ((other < key) == take_min) ? other : key;
                }
            }
            if (lane == k - below) ctl.result = key;
        }
    } else {
        const uint64_t mine = (tid < count) ? surv[tid] : ~0ull;
        const uint64_t srt = block_sort<kThreads>(mine, ctl.xch);
        if (tid == k - below) ctl.result = srt;
    }
    __syncthreads();
    *result = ctl.result;
    return true;
}

// Exact k-th smallest (0-based) of keys[0..n).  `keys` may live in global or shared memory and
// is not modified.  `bufs` are nb scratch key buffers (any mix of shared / global) that must not
// alias `keys`; at least two, and the largest must hold n keys.  Returns the key to every thread.
template <int kThreads>
__device__ __noinline__ uint64_t block_select_key(const uint64_t* keys, int n, int k, SelectCtl<kThreads>& ctl,
                                                  const SelectBuf* bufs, int nb, uint32_t seed)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const uint64_t* cur = keys;
    int cur_n = n, cur_k = k, fails = 0;
    for (int round = 0;; ++round) {
        if (cur_n <= kThreads) {
            // finish: sort the survivors (padded with the largest key) and read sorted[k]
            const uint64_t mine = (tid < cur_n) ? cur[tid] : ~0ull;
            __syncthreads();
            const uint64_t s = block_sort<kThreads>(mine, ctl.xch);
            if (tid == cur_k) ctl.result = s;
            __syncthreads();
            return ctl.result;
        }
        if (fails >= kMaxFails) return select_bisect<kThreads>(cur, cur_n, cur_k, ctl.ired);

        // ---- sample + sort ----
        HB_CHK(sample_index(tid, kThreads, cur_n, seed + 0x632be5abu * (uint32_t)(round + 1)), cur_n, 23);
        const uint64_t smp = cur[sample_index(tid, kThreads, cur_n, seed + 0x632be5abu * (uint32_t)(round + 1))];
        __syncthreads();
        const uint64_t sorted = block_sort<kThreads>(smp, ctl.xch);
        int r_lo, r_hi, r_mid;
        bracket_ranks(kThreads, cur_n, cur_k, 2.5f + 1.0f * (float)fails, r_lo, r_hi, r_mid);
        if (tid == 0) {
            ctl.cnt = 0;
            if (r_lo < 0) ctl.lo = 0ull;               // below every key
            if (r_hi > kThreads - 1) ctl.hi = ~0ull;   // above every key
        }
        if (tid == r_lo) ctl.lo = sorted;
        if (tid == r_hi) ctl.hi = sorted;
        __syncthreads();
        const uint64_t lo = ctl.lo, hi = ctl.hi;

        // smallest scratch buffer (not the current one) that should hold the survivors
        const float frac = fminf(1.0f, (float)(r_hi - r_lo + 1) / (float)kThreads);
        const int expect = (int)(frac * (float)cur_n * 1.5f) + 64;
        int pick = -1, big = -1;
        for (int b = 0; b < nb; b++) {
            if (bufs[b].ptr == cur) continue;
            if (big < 0 || bufs[b].cap > bufs[big].cap) big = b;
            if (bufs[b].cap >= expect && (pick < 0 || bufs[b].cap < bufs[pick].cap)) pick = b;
        }
        if (pick < 0) pick = big;
        uint64_t* out = bufs[pick].ptr;
        const int cap = bufs[pick].cap;

        // ---- count + compact pass ----
        // The keys usually sit in the CTA's global scratch (L2): a thread's loads are issued eight at a time,
        // so the pass costs a few memory latencies instead of one per key.
        int c_lt = 0;
        constexpr int kBatch = 8;
        for (int base = 0; base < cur_n; base += kBatch * kThreads) {
            uint64_t xs[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; u++) {
                const int i = base + u * kThreads + tid;
                xs[u] = (i < cur_n) ? cur[i] : 0ull;
            }
#pragma unroll
            for (int u = 0; u < kBatch; u++) {
                const int i = base + u * kThreads + tid;
                if (base + u * kThreads >= cur_n) break;  // uniform: no lane of the block has work left
                const bool valid = i < cur_n;
                const uint64_t x = xs[u];
                c_lt += (valid & (x < lo));
                const bool inr = valid & (x >= lo) & (x <= hi);
                const unsigned mask = __ballot_sync(0xffffffffu, inr);
                if (mask) {
                    const int leader = __ffs(mask) - 1;
                    int basepos = 0;
                    if (lane == leader) basepos = atomicAdd(&ctl.cnt, __popc(mask));
                    basepos = __shfl_sync(0xffffffffu, basepos, leader);
                    if (inr) {
                        const int pos = basepos + __popc(mask & ((1u << lane) - 1u));
                        if (pos < cap) {
                            HB_CHK(pos, cap, 24);
                            out[pos] = x;
                        }
                    }
                }
            }
        }
        c_lt = block_sum_int<kThreads>(c_lt, ctl.ired);  // contains the barriers that publish cnt / out[]
        const int c_in = ctl.cnt;
        __syncthreads();  // everyone has read cnt before the next round resets it
        if (cur_k >= c_lt && cur_k < c_lt + c_in && c_in <= cap && c_in < cur_n) {
            cur = out;
            cur_n = c_in;
            cur_k -= c_lt;
            fails = 0;
        } else {
            fails++;
        }
    }
}

}  // namespace hb
