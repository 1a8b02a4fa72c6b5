// hb_select.cuh -- block-wide EXACT k-th order statistic of n doubles (sm_100a).
//
// Replaces the reference's copy + Lomuto quicksort + sorted[mid] of remove_median()
// (likelihood3.c:36-105).  The reference needs one order statistic, not a sorted array, so
// the device version is a sampling select (Floyd-Rivest style) done with integer keys:
//   round:  draw S jittered-stride samples -> rank them by counting -> take two sample
//           order statistics lo/hi that bracket the target rank -> one pass that counts
//           keys < lo and compacts keys in [lo, hi] into a smaller buffer.
//   finish: <= kDirect survivors are ranked by counting; the key with the target rank wins.
//   ties / bad luck: a round that does not shrink the problem is retried with a new jitter;
//           after kMaxFails misses a 64-step bisection on the key bits (tie-proof) finishes.
// All comparisons are on order-preserving 64-bit integer keys, so nothing here touches the
// FP64 pipe that the model pass saturates.  NaN-free input is a precondition (the caller
// short-circuits NaN templates, see k_chain_eval).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

constexpr int kDirect = 512;    // survivors ranked directly
constexpr int kMaxFails = 4;    // unsuccessful rounds before the bisection fallback
constexpr int kSampMax = 512;   // largest sample

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

struct SelectBuf {
    uint64_t* ptr;
    int cap;
};

// Shared-memory control block of the select (the key buffers are passed separately).
struct SelectCtl {
    uint64_t samp[kSampMax];
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

// Rank `n` keys by counting; returns (to every thread) the key whose rank is k.
template <int kThreads>
__device__ uint64_t select_direct(const uint64_t* keys, int n, int k, SelectCtl& ctl)
{
    for (int j = threadIdx.x; j < n; j += kThreads) {
        const uint64_t x = keys[j];
        int r = 0;
        for (int i = 0; i < n; i++) {
            const uint64_t y = keys[i];
            r += (y < x) | ((y == x) & (i < j));
        }
        if (r == k) ctl.result = x;
    }
    __syncthreads();
    return ctl.result;
}

// Tie-proof fallback: smallest key K with #(keys <= K) >= k+1, by bisection on the bits.
template <int kThreads>
__device__ uint64_t select_bisect(const uint64_t* keys, int n, int k, SelectCtl& ctl)
{
    uint64_t lo = 0, hi = ~0ull;
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        int c = 0;
        for (int i = threadIdx.x; i < n; i += kThreads) c += (keys[i] <= mid);
        c = block_sum_int<kThreads>(c, ctl.ired);
        if (c >= k + 1) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// Exact k-th smallest (0-based) of keys[0..n).  `keys` may live in global or shared memory
// and is not modified.  `bufs` are nb scratch key buffers (any mix of shared / global) that
// must not alias `keys`; at least two, and the largest must hold n keys.
template <int kThreads>
__device__ uint64_t block_select_key(const uint64_t* keys, int n, int k, SelectCtl& ctl, const SelectBuf* bufs,
                                     int nb, uint32_t seed)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const uint64_t* cur = keys;
    int cur_n = n, cur_k = k, fails = 0;
    for (int round = 0;; ++round) {
        if (cur_n <= kDirect) return select_direct<kThreads>(cur, cur_n, cur_k, ctl);
        if (fails >= kMaxFails) return select_bisect<kThreads>(cur, cur_n, cur_k, ctl);

        // ---- sample ----
        const int S = (cur_n >= 16384) ? kSampMax : 256;
        const float q = ((float)cur_k + 0.5f) / (float)cur_n;
        const float ks = q * (float)S;
        // bracket half-width in sample ranks: z sigma of the binomial rank error, widened after a miss
        const float z = 2.5f + 1.0f * (float)fails;
        int m = (int)ceilf(z * sqrtf((float)S * q * (1.0f - q))) + 1;
        if (m < 2) m = 2;
        const int r_lo = (int)floorf(ks) - m, r_hi = (int)ceilf(ks) + m;
        const float frac = fminf(1.0f, (float)(r_hi - r_lo + 1) / (float)S);
        const int expect = (int)(frac * (float)cur_n * 1.5f) + 64;
        // smallest scratch buffer (not the current one) that should hold the survivors
        int pick = -1, big = -1;
        for (int b = 0; b < nb; b++) {
            if (bufs[b].ptr == cur) continue;
            if (big < 0 || bufs[b].cap > bufs[big].cap) big = b;
            if (bufs[b].cap >= expect && (pick < 0 || bufs[b].cap < bufs[pick].cap)) pick = b;
        }
        if (pick < 0) pick = big;
        uint64_t* out = bufs[pick].ptr;
        const int cap = bufs[pick].cap;

        for (int j = tid; j < S; j += kThreads) {
            const uint32_t h = mix32(seed ^ mix32((uint32_t)(round * 4099 + j) + 0x9e3779b9u));
            const float jit = (float)(h >> 8) * (1.0f / 16777216.0f);
            long long idx = (long long)(((double)j + (double)jit) * (double)cur_n / (double)S);
            if (idx > cur_n - 1) idx = cur_n - 1;
            ctl.samp[j] = cur[idx];
        }
        if (tid == 0) {
            ctl.lo = 0ull;       // below every key  (-> nothing is "less than lo")
            ctl.hi = ~0ull;      // above every key
            ctl.cnt = 0;
        }
        __syncthreads();
        for (int j = tid; j < S; j += kThreads) {
            const uint64_t x = ctl.samp[j];
            int r = 0;
            for (int i = 0; i < S; i++) {
                const uint64_t y = ctl.samp[i];
                r += (y < x) | ((y == x) & (i < j));
            }
            if (r == r_lo) ctl.lo = x;
            if (r == r_hi) ctl.hi = x;
        }
        __syncthreads();
        const uint64_t lo = ctl.lo, hi = ctl.hi;

        // ---- count + compact pass ----
        int c_lt = 0;
        for (int base = 0; base < cur_n; base += kThreads) {
            const int i = base + tid;
            const bool valid = i < cur_n;
            const uint64_t x = valid ? cur[i] : 0ull;
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
                    if (pos < cap) out[pos] = x;
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
