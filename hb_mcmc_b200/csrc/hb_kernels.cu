// hb_kernels.cu -- sm_100a kernels of the batched heartbeat-star likelihood.
//
//   k_prologue    one thread per chain: parameters -> ChainConst (hb_device.cuh)
//   k_chain_eval  persistent CTAs, one chain at a time per CTA:
//                   pass B  model at every time sample  -> template keys in scratch
//                   select  exact order statistic (reference median rule, quirk Q3)
//                   pass D  normalise + chi^2 against (flux, 1/sigma) -> logL
//   k_traj / k_scalar / k_mags   device versions of the small likelihood3.h entry points
//   k_fp64_peak   DFMA throughput probe (the roofline denominator, measured on the box)
//
// Replaces likelihood3.c:809-873 (loglikelihood) and :530-686 (calc_light_curve).
#include "hb_kernels.h"
#include "hb_device.cuh"
#include "hb_select.cuh"

namespace hb {

// ---------------------------------------------------------------------------
__global__ void k_prologue(const double* __restrict__ params, int n, MagSetup ms, ChainConst* __restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    ChainConst cc;
    chain_prologue(params + (size_t)c * NPARS, ms, cc);
    out[c] = cc;
}

// ---------------------------------------------------------------------------
template <int kThreads>
__device__ __forceinline__ double block_sum_double(double v, double* red)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double t = 0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; i++) t += red[i];
    return t;
}

struct EvalShared {
    ChainConst cc;
    SelectCtl ctl;
    double red[32];
    uint64_t candA[kCandA];
    uint64_t candB[kCandB];
};

size_t eval_smem_bytes() { return sizeof(EvalShared); }

// Median rank of likelihood3.c:97-101 (quirk Q3): even N -> N/2, odd N -> N/2 + 1.  N == 1
// would read one past the end in the reference; the only element is used instead.
__device__ __forceinline__ int median_rank(int N)
{
    int r = (N % 2 == 0) ? N / 2 : N / 2 + 1;
    return r < N ? r : N - 1;
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads, kEvalCtasPerSm)
k_chain_eval(const ChainConst* __restrict__ cc_all, int n_chains, const double* __restrict__ t,
             const double* __restrict__ flux, const double* __restrict__ w, int N, uint64_t* __restrict__ scratch,
             size_t scratch_stride, double* __restrict__ logL, double* __restrict__ lc_out, int* __restrict__ counter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EvalShared& sm = *reinterpret_cast<EvalShared*>(smem_raw);
    const int tid = threadIdx.x;
    uint64_t* tmpl = scratch + (size_t)blockIdx.x * 3 * scratch_stride;
    SelectBuf bufs[4] = {{sm.candB, kCandB}, {sm.candA, kCandA}, {tmpl + scratch_stride, N}, {tmpl + 2 * scratch_stride, N}};
    __shared__ int s_chain;

    for (;;) {
        // dynamic chain scheduler: chains differ in cost (eclipse fraction, Roche early-out)
        if (tid == 0) s_chain = atomicAdd(counter, 1);
        __syncthreads();
        const int chain = s_chain;
        if (chain >= n_chains) break;
        {
            const double* src = reinterpret_cast<const double*>(cc_all + chain);
            double* dst = reinterpret_cast<double*>(&sm.cc);
            for (int i = tid; i < (int)(sizeof(ChainConst) / sizeof(double)); i += kThreads) dst[i] = src[i];
        }
        __syncthreads();
        const int flag = (int)sm.cc.flag;
        const bool roche = flag & 1, nan_model = flag & 2;
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);

        if (nan_model || (roche && lc_out == nullptr) || N <= 0) {
            // quirk Q13: the reference evaluates the model and then discards it on Roche overflow
            if (lc_out != nullptr)
                for (int i = tid; i < N; i += kThreads) lc_out[(size_t)chain * N + i] = qnan;
            if (tid == 0 && logL != nullptr)
                logL[chain] = roche ? -0.5 * kBig : (nan_model ? qnan : -0.5 * sm.cc.chi2_extra);
            __syncthreads();
            continue;
        }

        // ---- pass B: model ----
        int nanflag = 0;
        {
            const ChainConst& cc = sm.cc;
            for (int i = tid; i < N; i += kThreads) {
                const double u = raw_flux(cc, t[i]);
                nanflag |= (u != u);
                tmpl[i] = dkey(u);
            }
        }
        const int any_nan = __syncthreads_or(nanflag);
        if (any_nan) {
            if (lc_out != nullptr)
                for (int i = tid; i < N; i += kThreads) lc_out[(size_t)chain * N + i] = qnan;
            if (tid == 0 && logL != nullptr) logL[chain] = roche ? -0.5 * kBig : qnan;
            __syncthreads();
            continue;
        }

        // ---- exact order statistic ----
        const double med = dunkey(block_select_key<kThreads>(tmpl, N, median_rank(N), sm.ctl, bufs, 4,
                                                            0x5bd1e995u * (uint32_t)(chain + 1)));

        // ---- pass D: normalise, chi^2 ----
        const double blend = sm.cc.blend, ft = sm.cc.ft;
        double acc = 0.;
        for (int i = tid; i < N; i += kThreads) {
            const double model = finish_template(dunkey(tmpl[i]), med, blend, ft);
            if (lc_out != nullptr) lc_out[(size_t)chain * N + i] = model;
            if (flux != nullptr) {
                const double r = (model - flux[i]) * w[i];
                acc = fma(r, r, acc);
            }
        }
        if (logL != nullptr) {
            const double chi2 = block_sum_double<kThreads>(acc, sm.red);
            if (tid == 0) logL[chain] = roche ? -0.5 * kBig : -0.5 * (chi2 + sm.cc.chi2_extra);
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
    const OrbitPoint o = kepler_point(times[i], e, T0, P);
    const double r = a * o.den;
    const double sq1 = sqrt(1 - e * e);
    const double nu = atan2(sq1 * o.sE, o.cE - e);
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
    const double R1 = pow(10., dev_getR(p[0]) + p[2] * dev_envelope_radius(p[0]));
    const double R2 = pow(10., dev_getR(p[1]) + p[3] * dev_envelope_radius(p[1]));
    const double T1 = pow(10., dev_getT(p[0]) + p[4] * dev_envelope_temp(p[0]));
    const double T2 = pow(10., dev_getT(p[1]) + p[5] * dev_envelope_temp(p[1]));
    double m[4];
    two_bb_mags(R1, R2, T1, T2, D, 0., 1, m);
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
cudaError_t launch_prologue(const double* params, int n, const MagSetup& ms, ChainConst* out, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_prologue<<<(n + 127) / 128, 128, 0, s>>>(params, n, ms, out);
    return cudaGetLastError();
}

cudaError_t configure_eval()
{
    return cudaFuncSetAttribute(k_chain_eval<kEvalThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(EvalShared));
}

cudaError_t launch_chain_eval(const ChainConst* cc, int n_chains, const double* t, const double* flux, const double* w,
                              int N, uint64_t* scratch, size_t scratch_stride, int grid, double* logL, double* lc_out,
                              int* counter, cudaStream_t s)
{
    if (n_chains <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    if (grid > n_chains) grid = n_chains;
    k_chain_eval<kEvalThreads><<<grid, kEvalThreads, sizeof(EvalShared), s>>>(cc, n_chains, t, flux, w, N, scratch,
                                                                               scratch_stride, logL, lc_out, counter);
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

cudaError_t launch_fp64_peak(double* out, int blocks, int iters, cudaStream_t s)
{
    k_fp64_peak<<<blocks, 256, 0, s>>>(out, iters, 1.0000001, 0.9999999);
    return cudaGetLastError();
}

}  // namespace hb
