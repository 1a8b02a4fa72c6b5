// hb_capi.cu -- the C ABI of libhb_b200.so (include/hb_b200.h): context, device buffers,
// host<->device staging and kernel launches.  No model arithmetic lives here.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hb_b200.h"
#include "hb_device.cuh"
#include "hb_kernels.h"
#include "hb_comm.h"
#include "hb_gaia_pt.cuh"
#include "hb_pt.cuh"
#include "hb_pt_run.h"
#include "hb_sincos_tab.h"

using namespace hb;

static std::string g_error;
static std::mutex g_error_mu;

struct hb_ctx {
    int device = 0;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t global_mem = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::mutex mu;
    std::string err;
    long launches = 0;

    // observed light curve (device)
    double* d_t = nullptr;    // seconds (t * 86400)
    double2* d_sctab = nullptr;  // {sin, cos}(2 pi k / 1024) for sincos_tab, filled at hb_create
    double2* d_fw = nullptr;  // {flux, 1 / max(sigma, 1e-5)} interleaved: one 16-byte load per sample
    long N = 0;
    bool has_data = false;
    MagSetup ms;

    // per-batch work buffers, grown on demand
    double* d_params = nullptr;
    ChainConst* d_cc = nullptr;
    double* d_logL = nullptr;
    long cap_chains = 0;
    uint64_t* d_scratch = nullptr;   // per region: 3 key buffers of key_stride + the per-segment partial sums
    size_t key_stride = 0, region_stride = 0;
    int scratch_segments = 0;        // segments the partial-sum area of a region holds
    int grid = 0;
    int* d_counter = nullptr;        // [0] work-item scheduler of k_chain_eval
    unsigned long long* d_evaluated = nullptr;  // chains whose model was really evaluated (hb_evaluated_chains)
    ChainSync* d_sync = nullptr;     // [grid] hand-over words of chains shared by several CTAs
    double sum_w2 = 0.;              // sum of the squared weights of the uploaded data set
    bool zero_copy = true;           // page-locked caller buffers are read / written in place (HB_ZERO_COPY=0: DMA copies)
    int max_parts = kMaxSegments;    // most CTAs one light curve may be spread over (hb_set_max_parts)
    double* d_lc = nullptr;
    size_t cap_lc = 0;
    double* d_times2 = nullptr;  // hb_calc_light_curve / hb_traj time grid
    long cap_times2 = 0;
    double* d_small = nullptr;   // 64 doubles of scalar args / results
    double* d_aux = nullptr;     // generic output buffer
    size_t cap_aux = 0;

    // pinned host staging
    double* h_pin = nullptr;
    size_t cap_pin = 0;

    // events bracketing the last k_chain_eval launch on its stream (hb_last_eval_kernel_ms)
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;
    bool ev_valid = false;
    bool time_kernels = false;

    // bumped whenever a device pointer or a by-value kernel argument of the step changes
    // (hb_set_data, hb_set_mags, buffer growth): captured graphs of older generations are stale
    unsigned long generation = 0;

    // half-width, in binomial sigmas, of the median bracket taken from the pre-sample (hb_set_bracket_sigma)
    float bracket_sigma = 2.5f;
    // largest exponent word of a Newton iterate the logL-only pass accepts before it re-evaluates the chain
    // with the per-sample library fallback (hb_set_sincos_range)
    int hot_hi_limit = kSincosTabHiLimit;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int fail_cuda(hb_ctx* c, cudaError_t e, const char* what)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    c->err = buf;
    return HB_ERR_CUDA;
}
int fail_arg(hb_ctx* c, const char* what)
{
    c->err = what;
    return HB_ERR_ARG;
}

#define CK(call)                                                      \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return fail_cuda(ctx, e__, #call);    \
    } while (0)

template <typename T>
cudaError_t grow(T*& p, size_t& cap, size_t need)
{
    if (need <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t n = std::max(need, (size_t)64);
    cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
}

cudaError_t grow_pin(hb_ctx* c, size_t doubles)
{
    if (doubles <= c->cap_pin) return cudaSuccess;
    if (c->h_pin) cudaFreeHost(c->h_pin);
    c->h_pin = nullptr;
    c->cap_pin = 0;
    cudaError_t e = cudaMallocHost((void**)&c->h_pin, doubles * sizeof(double));
    if (e == cudaSuccess) c->cap_pin = doubles;
    return e;
}

int ensure_chains(hb_ctx* ctx, long n)
{
    if (n <= ctx->cap_chains) return HB_OK;
    long cap = std::max(n, 2 * ctx->cap_chains);
    if (ctx->d_params) cudaFree(ctx->d_params);
    if (ctx->d_cc) cudaFree(ctx->d_cc);
    if (ctx->d_logL) cudaFree(ctx->d_logL);
    ctx->d_params = nullptr; ctx->d_cc = nullptr; ctx->d_logL = nullptr; ctx->cap_chains = 0;
    CK(cudaMalloc((void**)&ctx->d_params, (size_t)cap * NPARS * sizeof(double)));
    CK(cudaMalloc((void**)&ctx->d_cc, (size_t)cap * sizeof(ChainConst)));
    CK(cudaMalloc((void**)&ctx->d_logL, (size_t)cap * sizeof(double)));
    ctx->cap_chains = cap;
    ctx->generation++;
    return HB_OK;
}

// scratch: per region (one per resident CTA, or per chain when chains are shared by CTAs) three key arrays of
// `key_stride` entries (template + two select buffers) and 2 x threads partial sums for every segment
int ensure_scratch(hb_ctx* ctx, long n_points)
{
    const size_t tile = (size_t)eval_tile();
    const size_t stride = ((size_t)std::max(n_points, 1L) + tile - 1) / tile * tile;
    const int nseg = eval_segments(n_points);
    if (ctx->d_scratch && stride <= ctx->key_stride && nseg <= ctx->scratch_segments) return HB_OK;
    const size_t ks = std::max(stride, ctx->key_stride);
    const int sg = std::max(nseg, ctx->scratch_segments);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    ctx->d_scratch = nullptr;
    ctx->key_stride = ctx->region_stride = 0;
    ctx->scratch_segments = 0;
    const size_t region = 3 * ks + (size_t)sg * 2 * tile;
    CK(cudaMalloc((void**)&ctx->d_scratch, (size_t)ctx->grid * region * sizeof(uint64_t)));
    ctx->key_stride = ks;
    ctx->region_stride = region;
    ctx->scratch_segments = sg;
    ctx->generation++;
    return HB_OK;
}

// CTAs per chain: 1 when the batch fills more than half the grid (or the pass has to store the template), else as
// many as keep chains x parts within the grid, do not exceed the light curve's segment count and respect
// hb_set_max_parts (the kernel raises it further for the chains of the batch it skips).  The result never depends
// on it (see k_chain_eval).  (Sharing only the chains of a large batch's last, partial wave was built and measured:
// the per-item decode costs the main path more than the shorter tail gives back.)
int choose_parts(const hb_ctx* ctx, long n_chains, long N, bool hot)
{
    if (!hot || N <= kCandA / 2 || n_chains > 2 * kEvalThreads) return 1;  // (the kernel lists a shared batch's chains in 2 x threads slots)
    const int nseg = eval_segments(N);
    const int room = (int)std::min<long>(std::min(nseg, ctx->max_parts), ctx->grid / n_chains);
    if (room < 2) return 1;
    const int spp = (nseg + room - 1) / room;  // segments per part
    return (nseg + spp - 1) / spp;             // parts that get at least one segment (need not divide nseg)
}

// The device's address of page-locked host memory it can read and write in place (cudaHostAlloc / cudaHostRegister
// under unified addressing: torch pin_memory, the shim's own buffers), or nullptr.
void* mapped_host(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

// host -> device through the pinned staging buffer, on the context stream
int upload(hb_ctx* ctx, double* dst, const double* src, size_t n)
{
    if (n == 0) return HB_OK;
    CK(grow_pin(ctx, n));
    std::memcpy(ctx->h_pin, src, n * sizeof(double));
    CK(cudaMemcpyAsync(dst, ctx->h_pin, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // h_pin is reused
    return HB_OK;
}
int download(hb_ctx* ctx, double* dst, const double* src, size_t n)
{
    if (n == 0) return HB_OK;
    CK(grow_pin(ctx, n));
    CK(cudaMemcpyAsync(ctx->h_pin, src, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::memcpy(dst, ctx->h_pin, n * sizeof(double));
    return HB_OK;
}

int run_eval(hb_ctx* ctx, const double* d_params, long n, const double* d_t, const double2* d_fw,
             long N, double* d_logL, double* d_lc)
{
    EvalArgs a;
    a.cc = ctx->d_cc;
    a.n_chains = (int)n;
    a.N = (int)N;
    a.tsec = d_t;
    a.fw = d_fw;
    a.sum_w2 = d_fw ? ctx->sum_w2 : 0.0;
    a.scratch = ctx->d_scratch;
    a.region_stride = ctx->region_stride;
    a.key_stride = ctx->key_stride;
    a.logL = d_logL;
    a.lc_out = d_lc;
    a.counter = ctx->d_counter;
    a.evaluated = ctx->d_evaluated;
    a.sync = ctx->d_sync;
    a.sctab = ctx->d_sctab;
    a.bracket_sigma = ctx->bracket_sigma;
    a.hot_hi_limit = ctx->hot_hi_limit;
    a.nseg = eval_segments(N);
    a.seg_shift = eval_seg_shift(N);
    a.nparts = choose_parts(ctx, n, N, d_fw != nullptr && d_lc == nullptr);
    a.max_parts = ctx->max_parts;
    const long n_work = n * a.nparts;
    // (a shared batch of several chains gets the whole grid: the kernel hands the CTAs of chains it skips to the others)
    const int grid = (a.nparts > 1 && n > 1) ? ctx->grid : (int)std::min<long>(ctx->grid, n_work);
    CK(launch_prologue(d_params, (int)n, ctx->ms, ctx->d_cc, ctx->d_counter, grid, ctx->stream));
    if (ctx->time_kernels) CK(cudaEventRecord(ctx->ev_k0, ctx->stream));
    CK(launch_chain_eval(a, grid, ctx->stream));
    if (ctx->time_kernels) {
        CK(cudaEventRecord(ctx->ev_k1, ctx->stream));
        ctx->ev_valid = true;
    }
    ctx->launches += 2;
    return HB_OK;
}

}  // namespace

extern "C" {

const char* hb_global_error(void) { return g_error.c_str(); }

int hb_create(hb_ctx** out, int device)
{
    if (!out) return HB_ERR_ARG;
    *out = nullptr;
    auto set_global = [](const std::string& s) {
        std::lock_guard<std::mutex> g(g_error_mu);
        g_error = s;
    };
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_global(std::string("hb_create: no CUDA device: ") + cudaGetErrorString(e));
        return HB_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        set_global("hb_create: device index out of range");
        return HB_ERR_ARG;
    }
    hb_ctx* ctx = new hb_ctx();
    ctx->device = device;
    DeviceGuard g(device);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        set_global(std::string("hb_create: ") + cudaGetErrorString(e));
        delete ctx;
        return HB_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->global_mem = prop.totalGlobalMem;
    if (prop.major != 10) {
        // the library carries sm_100a code only: fail loudly instead of hitting "no kernel image"
        char buf[256];
        snprintf(buf, sizeof(buf), "hb_create: device %d is sm_%d%d; libhb_b200 is built for sm_100a (B200) only", device,
                 prop.major, prop.minor);
        set_global(buf);
        delete ctx;
        return HB_ERR_CUDA;
    }
    ctx->grid = ctx->sm_count * kEvalCtasPerSm;
    if (const char* z = std::getenv("HB_ZERO_COPY")) ctx->zero_copy = !(z[0] == '0');  // (tools: the DMA-copy path for comparison)
    ctx->ms.mag_data[0] = 1000.;  // mcmc_wrapper2.c:322-327
    for (int i = 0; i < 4; i++) {
        ctx->ms.mag_data[i + 1] = 1.;
        ctx->ms.magerr[i] = kBig;
    }
    ctx->ms.use_gmag = 1;  // likelihood3.h:11-12
    ctx->ms.use_color = 0;
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc((void**)&ctx->d_counter, 2 * sizeof(int)) == cudaSuccess &&
              cudaMemset(ctx->d_counter, 0, 2 * sizeof(int)) == cudaSuccess &&
              cudaMalloc((void**)&ctx->d_evaluated, sizeof(unsigned long long)) == cudaSuccess &&
              cudaMemset(ctx->d_evaluated, 0, sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc((void**)&ctx->d_sync, (size_t)ctx->grid * sizeof(ChainSync)) == cudaSuccess &&
              cudaMemset(ctx->d_sync, 0, (size_t)ctx->grid * sizeof(ChainSync)) == cudaSuccess &&
              cudaMalloc((void**)&ctx->d_small, 64 * sizeof(double)) == cudaSuccess &&
              cudaMalloc((void**)&ctx->d_sctab, kSinTabN * sizeof(double2)) == cudaSuccess &&
              cudaEventCreate(&ctx->ev_k0) == cudaSuccess && cudaEventCreate(&ctx->ev_k1) == cudaSuccess &&
              configure_eval() == cudaSuccess;
    if (ok) {
        std::vector<double> tab(2 * kSinTabN);
        fill_sincos_table(tab.data());
        ok = cudaMemcpy(ctx->d_sctab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (!ok) {
        set_global(std::string("hb_create: ") + cudaGetErrorString(cudaGetLastError()));
        hb_destroy(ctx);
        return HB_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return HB_OK;
}

void hb_destroy(hb_ctx* ctx)
{
    if (!ctx) return;
    {
        DeviceGuard g(ctx->device);
        cudaDeviceSynchronize();
        cudaFree(ctx->d_t); cudaFree(ctx->d_fw);
        cudaFree(ctx->d_params); cudaFree(ctx->d_cc); cudaFree(ctx->d_logL);
        cudaFree(ctx->d_scratch); cudaFree(ctx->d_counter); cudaFree(ctx->d_lc); cudaFree(ctx->d_evaluated); cudaFree(ctx->d_sync);
        cudaFree(ctx->d_times2); cudaFree(ctx->d_small); cudaFree(ctx->d_aux); cudaFree(ctx->d_sctab);
        if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
        if (ctx->ev_k0) cudaEventDestroy(ctx->ev_k0);
        if (ctx->ev_k1) cudaEventDestroy(ctx->ev_k1);
        if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    }
    delete ctx;
}

const char* hb_last_error(const hb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

long hb_launch_count(const hb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int hb_device_info(hb_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, long* global_mem_mb)
{
    if (!ctx) return HB_ERR_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (global_mem_mb) *global_mem_mb = (long)(ctx->global_mem >> 20);
    return HB_OK;
}

int hb_set_stream(hb_ctx* ctx, void* cuda_stream)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return HB_OK;
}

int hb_sync(hb_ctx* ctx)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

int hb_set_data(hb_ctx* ctx, const double* t, const double* flux, const double* err, long n)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (n < 0 || (n > 0 && (!t || !flux || !err))) return fail_arg(ctx, "hb_set_data: null array or negative n");
    if (n > 0x7fffff00L) return fail_arg(ctx, "hb_set_data: n too large");
    DeviceGuard g(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_t); cudaFree(ctx->d_fw);
    ctx->d_t = nullptr;
    ctx->d_fw = nullptr;
    ctx->has_data = false;
    // padded to whole tiles plus one (k_chain_eval reads full tiles and requests the next tile's time samples one
    // iteration ahead without a bounds test); the padding is finite and adds nothing to any result
    const size_t tile = (size_t)eval_tile();
    const size_t padded = (((size_t)std::max(n, 1L) + tile - 1) / tile + 1) * tile;
    size_t alloc = padded * sizeof(double);
    CK(cudaMalloc((void**)&ctx->d_t, alloc));
    CK(cudaMalloc((void**)&ctx->d_fw, 2 * alloc));
    CK(cudaMemsetAsync(ctx->d_t, 0, alloc, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_fw, 0, 2 * alloc, ctx->stream));  // padding: flux 0, weight 0 (adds nothing to chi^2)
    // weights 1/max(sigma, 1e-5): the clamp of likelihood3.c:824-827 applied once at upload
    std::vector<double> fw(2 * (size_t)n);
    long double sw2 = 0.0L;  // S2 = sum w^2 of the chi^2 expansion does not depend on the chain: formed once, here
    for (long i = 0; i < n; i++) {
        fw[2 * i] = flux[i];
        const double w = 1.0 / (err[i] < 1.e-5 ? 1.e-5 : err[i]);
        fw[2 * i + 1] = w;
        sw2 += (long double)w * (long double)w;
    }
    ctx->sum_w2 = (double)sw2;
    int rc;
    if ((rc = upload(ctx, ctx->d_t, t, (size_t)n)) != HB_OK) return rc;
    CK(launch_to_seconds(ctx->d_t, (int)n, ctx->d_t, ctx->stream));  // d_t holds t * 86400 from here on
    ctx->launches += (n > 0);
    if ((rc = upload(ctx, reinterpret_cast<double*>(ctx->d_fw), fw.data(), 2 * (size_t)n)) != HB_OK) return rc;
    ctx->N = n;
    if ((rc = ensure_scratch(ctx, n)) != HB_OK) return rc;
    ctx->has_data = true;
    ctx->generation++;
    return HB_OK;
}

int hb_set_mags(hb_ctx* ctx, const double* mag_data, const double* magerr, int use_gmag, int use_color)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!mag_data || !magerr) return fail_arg(ctx, "hb_set_mags: null array");
    for (int i = 0; i < 5; i++) ctx->ms.mag_data[i] = mag_data[i];
    for (int i = 0; i < 4; i++) ctx->ms.magerr[i] = magerr[i];
    ctx->ms.use_gmag = use_gmag ? 1 : 0;
    ctx->ms.use_color = use_color ? 1 : 0;
    ctx->generation++;
    return HB_OK;
}

int hb_loglikelihood_batch_dev(hb_ctx* ctx, const double* d_params, long n_chains, double* d_logL)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->has_data) {
        ctx->err = "hb_loglikelihood_batch_dev: hb_set_data has not been called";
        return HB_ERR_STATE;
    }
    if (n_chains < 0 || (n_chains > 0 && (!d_params || !d_logL))) return fail_arg(ctx, "hb_loglikelihood_batch_dev: bad argument");
    if (n_chains == 0) return HB_OK;
    DeviceGuard g(ctx->device);
    int rc;
    if ((rc = ensure_chains(ctx, n_chains)) != HB_OK) return rc;
    return run_eval(ctx, d_params, n_chains, ctx->d_t, ctx->d_fw, ctx->N, d_logL, nullptr);
}

int hb_loglikelihood_batch(hb_ctx* ctx, const double* params, long n_chains, double* logL)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->has_data) {
        ctx->err = "hb_loglikelihood_batch: hb_set_data has not been called";
        return HB_ERR_STATE;
    }
    if (n_chains < 0 || (n_chains > 0 && (!params || !logL))) return fail_arg(ctx, "hb_loglikelihood_batch: bad argument");
    if (n_chains == 0) return HB_OK;
    DeviceGuard g(ctx->device);
    int rc;
    if ((rc = ensure_chains(ctx, n_chains)) != HB_OK) return rc;
    const size_t np = (size_t)n_chains * NPARS;
    if (n_chains <= prologue_small_max()) {
        // A small batch (the calls of one OpenMP team through the shim, a ladder, C1) is latency, and the two DMA
        // copies are a third of it: the kernels read the parameters from the pinned staging buffer and write the
        // results into it directly instead (page-locked memory is mapped into the device's address space under
        // unified addressing; k_prologue_small fetches a chain's 21 parameters in one coalesced read).
        CK(grow_pin(ctx, np + (size_t)n_chains));
        std::memcpy(ctx->h_pin, params, np * sizeof(double));
        double* h_out = ctx->h_pin + np;
        if ((rc = run_eval(ctx, ctx->h_pin, n_chains, ctx->d_t, ctx->d_fw, ctx->N, h_out, nullptr)) != HB_OK) return rc;
        CK(cudaStreamSynchronize(ctx->stream));
        std::memcpy(logL, h_out, (size_t)n_chains * sizeof(double));
        return HB_OK;
    }
    // Page-locked caller buffers (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) are DMA'd in place;
    // pageable ones go through the context's pinned staging buffer in chunks, so that the host copy of
    // chunk k+1 overlaps the transfer of chunk k.
    // Page-locked on both sides: no copy at all.  k_prologue reads each chain's 21 parameters from the caller's buffer
    // in one coalesced read while the other warps compute (the 688 KB of C2 cross the bus behind the prologue's libm
    // work instead of in front of it), and k_chain_eval writes logL into the caller's array: 25 us of a 0.76 ms call.
    double* const m_in = static_cast<double*>(mapped_host(params));
    double* const m_out = static_cast<double*>(mapped_host(logL));
    if (m_in != nullptr && m_out != nullptr && ctx->zero_copy) {
        if ((rc = run_eval(ctx, m_in, n_chains, ctx->d_t, ctx->d_fw, ctx->N, m_out, nullptr)) != HB_OK) return rc;
        CK(cudaStreamSynchronize(ctx->stream));
        return HB_OK;
    }
    const bool in_pinned = m_in != nullptr, out_pinned = m_out != nullptr;
    if (in_pinned) {
        CK(cudaMemcpyAsync(ctx->d_params, params, np * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    } else {
        CK(grow_pin(ctx, np));
        const long n_chunks = n_chains >= 1024 ? 4 : 1;
        for (long k = 0; k < n_chunks; k++) {
            const size_t a = (size_t)(n_chains * k / n_chunks) * NPARS, b = (size_t)(n_chains * (k + 1) / n_chunks) * NPARS;
            std::memcpy(ctx->h_pin + a, params + a, (b - a) * sizeof(double));
            CK(cudaMemcpyAsync(ctx->d_params + a, ctx->h_pin + a, (b - a) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    if ((rc = run_eval(ctx, ctx->d_params, n_chains, ctx->d_t, ctx->d_fw, ctx->N, ctx->d_logL, nullptr)) != HB_OK)
        return rc;
    if (out_pinned) {
        CK(cudaMemcpyAsync(logL, ctx->d_logL, (size_t)n_chains * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } else {
        // the D2H lands in the head of the pinned buffer; stream order keeps it after the H2D reads
        CK(grow_pin(ctx, (size_t)n_chains));
        CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_logL, (size_t)n_chains * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        std::memcpy(logL, ctx->h_pin, (size_t)n_chains * sizeof(double));
    }
    return HB_OK;
}

int hb_light_curve_batch(hb_ctx* ctx, const double* params, long n_chains, double* templates)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->has_data) {
        ctx->err = "hb_light_curve_batch: hb_set_data has not been called";
        return HB_ERR_STATE;
    }
    if (n_chains < 0 || (n_chains > 0 && (!params || !templates))) return fail_arg(ctx, "hb_light_curve_batch: bad argument");
    if (n_chains == 0 || ctx->N == 0) return HB_OK;
    DeviceGuard g(ctx->device);
    int rc;
    if ((rc = ensure_chains(ctx, n_chains)) != HB_OK) return rc;
    CK(grow(ctx->d_lc, ctx->cap_lc, (size_t)n_chains * (size_t)ctx->N));
    if ((rc = upload(ctx, ctx->d_params, params, (size_t)n_chains * NPARS)) != HB_OK) return rc;
    if ((rc = run_eval(ctx, ctx->d_params, n_chains, ctx->d_t, nullptr, ctx->N, nullptr, ctx->d_lc)) != HB_OK) return rc;
    return download(ctx, templates, ctx->d_lc, (size_t)n_chains * (size_t)ctx->N);
}

static int stage_times(hb_ctx* ctx, const double* times, long nt)
{
    if (nt > ctx->cap_times2) {
        if (ctx->d_times2) cudaFree(ctx->d_times2);
        ctx->d_times2 = nullptr;
        ctx->cap_times2 = 0;
        const size_t tile = (size_t)eval_tile();
        const size_t padded = (((size_t)nt + tile - 1) / tile + 1) * tile;  // whole tiles plus one (see hb_set_data)
        CK(cudaMalloc((void**)&ctx->d_times2, padded * sizeof(double)));
        CK(cudaMemsetAsync(ctx->d_times2, 0, padded * sizeof(double), ctx->stream));
        ctx->cap_times2 = (long)padded;
    }
    return upload(ctx, ctx->d_times2, times, (size_t)nt);
}

int hb_calc_light_curve(hb_ctx* ctx, const double* times, long nt, const double* pars, double* tmpl)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (nt < 0 || !pars || (nt > 0 && (!times || !tmpl))) return fail_arg(ctx, "hb_calc_light_curve: bad argument");
    if (nt == 0) return HB_OK;
    if (nt > 0x7fffff00L) return fail_arg(ctx, "hb_calc_light_curve: nt too large");
    DeviceGuard g(ctx->device);
    int rc;
    if ((rc = ensure_chains(ctx, 1)) != HB_OK) return rc;
    if ((rc = ensure_scratch(ctx, nt)) != HB_OK) return rc;
    if ((rc = stage_times(ctx, times, nt)) != HB_OK) return rc;
    CK(launch_to_seconds(ctx->d_times2, (int)nt, ctx->d_times2, ctx->stream));
    ctx->launches += 1;
    CK(grow(ctx->d_lc, ctx->cap_lc, (size_t)nt));
    if ((rc = upload(ctx, ctx->d_params, pars, NPARS)) != HB_OK) return rc;
    if ((rc = run_eval(ctx, ctx->d_params, 1, ctx->d_times2, nullptr, nt, nullptr, ctx->d_lc)) != HB_OK) return rc;
    return download(ctx, tmpl, ctx->d_lc, (size_t)nt);
}

int hb_chain_info_batch(hb_ctx* ctx, const double* params, long n_chains, double D, double* out)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (n_chains < 0 || (n_chains > 0 && (!params || !out))) return fail_arg(ctx, "hb_chain_info_batch: bad argument");
    if (n_chains == 0) return HB_OK;
    DeviceGuard g(ctx->device);
    int rc;
    if ((rc = ensure_chains(ctx, n_chains)) != HB_OK) return rc;
    CK(grow(ctx->d_aux, ctx->cap_aux, (size_t)n_chains * 9));
    if ((rc = upload(ctx, ctx->d_params, params, (size_t)n_chains * NPARS)) != HB_OK) return rc;
    MagSetup ms = ctx->ms;
    ms.mag_data[0] = D;
    CK(launch_prologue(ctx->d_params, (int)n_chains, ms, ctx->d_cc, nullptr, 0, ctx->stream));
    CK(launch_chain_info(ctx->d_cc, (int)n_chains, ctx->d_aux, ctx->stream));
    ctx->launches += 2;
    return download(ctx, out, ctx->d_aux, (size_t)n_chains * 9);
}

int hb_traj(hb_ctx* ctx, const double* times, long nt, const double* traj_pars, double* d_arr, double* Z1_arr,
            double* Z2_arr, double* rr_arr, double* ff_arr)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (nt < 0 || !traj_pars || (nt > 0 && (!times || !d_arr || !Z1_arr || !Z2_arr || !rr_arr || !ff_arr)))
        return fail_arg(ctx, "hb_traj: bad argument");
    if (nt == 0) return HB_OK;
    DeviceGuard g(ctx->device);
    int rc;
    if ((rc = stage_times(ctx, times, nt)) != HB_OK) return rc;
    CK(grow(ctx->d_aux, ctx->cap_aux, (size_t)nt * 5));
    if ((rc = upload(ctx, ctx->d_small, traj_pars, 7)) != HB_OK) return rc;
    double* o = ctx->d_aux;
    CK(launch_traj(ctx->d_times2, (int)nt, ctx->d_small, o, o + nt, o + 2 * nt, o + 3 * nt, o + 4 * nt, ctx->stream));
    ctx->launches += 1;
    std::vector<double> h((size_t)nt * 5);
    if ((rc = download(ctx, h.data(), o, (size_t)nt * 5)) != HB_OK) return rc;
    std::memcpy(d_arr, h.data(), nt * sizeof(double));
    std::memcpy(Z1_arr, h.data() + nt, nt * sizeof(double));
    std::memcpy(Z2_arr, h.data() + 2 * nt, nt * sizeof(double));
    std::memcpy(rr_arr, h.data() + 3 * nt, nt * sizeof(double));
    std::memcpy(ff_arr, h.data() + 4 * nt, nt * sizeof(double));
    return HB_OK;
}

int hb_order_statistic(hb_ctx* ctx, const double* x, long n, long k, double* out)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!x || !out || n <= 0 || k < 0 || k >= n || n > 0x7fffff00L) return fail_arg(ctx, "hb_order_statistic: bad argument");
    DeviceGuard g(ctx->device);
    int rc;
    if ((rc = ensure_scratch(ctx, n)) != HB_OK) return rc;
    if ((rc = stage_times(ctx, x, n)) != HB_OK) return rc;
    CK(launch_order_stat(ctx->d_times2, (int)n, (int)k, ctx->d_scratch, ctx->key_stride, ctx->d_small + 48, ctx->stream));
    ctx->launches += 1;
    double r[2];
    if ((rc = download(ctx, r, ctx->d_small + 48, 2)) != HB_OK) return rc;
    *out = r[0];
    return HB_OK;
}

int hb_remove_median(hb_ctx* ctx, double* arr, long n)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (n < 0 || (n > 0 && !arr) || n > 0x7fffff00L) return fail_arg(ctx, "hb_remove_median: bad argument");
    if (n == 0) return HB_OK;
    DeviceGuard g(ctx->device);
    int rc;
    if ((rc = ensure_scratch(ctx, n)) != HB_OK) return rc;
    if ((rc = stage_times(ctx, arr, n)) != HB_OK) return rc;
    // index rule of likelihood3.c:97-101 (quirk Q3); n == 1 uses the only element
    long k = (n % 2 == 0) ? n / 2 : n / 2 + 1;
    if (k > n - 1) k = n - 1;
    CK(launch_order_stat(ctx->d_times2, (int)n, (int)k, ctx->d_scratch, ctx->key_stride, ctx->d_small + 48, ctx->stream));
    CK(launch_subtract(ctx->d_times2, (int)n, ctx->d_small + 48, ctx->stream));
    ctx->launches += 2;
    return download(ctx, arr, ctx->d_times2, (size_t)n);
}

int hb_scalar(hb_ctx* ctx, int op, const double* args, int nargs, double* out)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    static const int need[9] = {1, 1, 1, 1, 1, 3, 8, 11, 9};
    if (op < 0 || op > 8 || !args || !out || nargs != need[op]) return fail_arg(ctx, "hb_scalar: bad op / argument count");
    DeviceGuard g(ctx->device);
    int rc;
    if ((rc = upload(ctx, ctx->d_small, args, (size_t)nargs)) != HB_OK) return rc;
    CK(launch_scalar(op, ctx->d_small, ctx->d_small + 32, ctx->stream));
    ctx->launches += 1;
    return download(ctx, out, ctx->d_small + 32, 1);
}

int hb_gaia_batch(hb_ctx* ctx, const double* p6, long n, double D, const double* data, const double* err, double* mags,
                  double* logL)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (n < 0 || (n > 0 && !p6) || (logL && (!data || !err))) return fail_arg(ctx, "hb_gaia_batch: bad argument");
    if (n == 0) return HB_OK;
    DeviceGuard g(ctx->device);
    int rc;
    CK(grow(ctx->d_aux, ctx->cap_aux, (size_t)n * 11));
    double* d_p = ctx->d_aux;
    double* d_m = d_p + (size_t)n * 6;
    double* d_l = d_m + (size_t)n * 4;
    if ((rc = upload(ctx, d_p, p6, (size_t)n * 6)) != HB_OK) return rc;
    if (logL) {
        double de[8];
        for (int i = 0; i < 4; i++) { de[i] = data[i]; de[4 + i] = err[i]; }
        if ((rc = upload(ctx, ctx->d_small, de, 8)) != HB_OK) return rc;
    }
    CK(launch_gaia(d_p, (int)n, D, ctx->d_small, ctx->d_small + 4, mags ? d_m : nullptr, logL ? d_l : nullptr, ctx->stream));
    ctx->launches += 1;
    if (mags && (rc = download(ctx, mags, d_m, (size_t)n * 4)) != HB_OK) return rc;
    if (logL && (rc = download(ctx, logL, d_l, (size_t)n)) != HB_OK) return rc;
    return HB_OK;
}

int hb_set_bracket_sigma(hb_ctx* ctx, double sigma)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!(sigma >= 0.0) || sigma > 100.0) return fail_arg(ctx, "hb_set_bracket_sigma: need 0 <= sigma <= 100");
    ctx->bracket_sigma = (float)sigma;
    ctx->generation++;  // a by-value argument of the captured step changed
    return HB_OK;
}

int hb_set_max_parts(hb_ctx* ctx, int max_parts)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (max_parts < 1 || max_parts > kMaxSegments || (max_parts & (max_parts - 1)))
        return fail_arg(ctx, "hb_set_max_parts: need a power of two between 1 and 64");
    ctx->max_parts = max_parts;
    ctx->generation++;
    return HB_OK;
}

int hb_evaluated_chains(hb_ctx* ctx, unsigned long long* count, int reset)
{
    if (!ctx || !count) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    CK(cudaMemcpyAsync(count, ctx->d_evaluated, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    if (reset) CK(cudaMemsetAsync(ctx->d_evaluated, 0, sizeof(unsigned long long), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

int hb_set_sincos_range(hb_ctx* ctx, double max_abs)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!(max_abs > 0.0) || max_abs > 1024.0) return fail_arg(ctx, "hb_set_sincos_range: need 0 < max_abs <= 1024");
    long long bits;
    std::memcpy(&bits, &max_abs, sizeof bits);
    ctx->hot_hi_limit = (int)(bits >> 32);
    ctx->generation++;
    return HB_OK;
}

int hb_time_kernels(hb_ctx* ctx, int enable)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->time_kernels = enable != 0;
    ctx->ev_valid = false;
    return HB_OK;
}

int hb_last_eval_kernel_ms(hb_ctx* ctx, double* ms)
{
    if (!ctx || !ms) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->ev_valid) {
        ctx->err = "hb_last_eval_kernel_ms: no timed launch (call hb_time_kernels(ctx, 1) first)";
        return HB_ERR_STATE;
    }
    DeviceGuard g(ctx->device);
    CK(cudaEventSynchronize(ctx->ev_k1));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, ctx->ev_k0, ctx->ev_k1));
    *ms = (double)f;
    return HB_OK;
}

#ifdef HB_DEBUG_BOUNDS
/* libhb_b200_dbg.so only: proves that the bounds assertions fire (index against capacity 4; the trap poisons the context) */
int hb_debug_bounds_selftest(hb_ctx* ctx, int index)
{
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    CK(launch_bounds_selftest(index, ctx->d_counter + 1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}
#endif

int hb_fp64_peak(hb_ctx* ctx, double seconds_target, double* tflops)
{
    if (!ctx || !tflops) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const int blocks = ctx->sm_count * 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int iters = 2000;
    double best = 0.;
    double total_s = 0.;
    if (seconds_target <= 0) seconds_target = 0.2;
    for (int rep = 0; rep < 64 && total_s < seconds_target; rep++) {
        CK(cudaEventRecord(e0, ctx->stream));
        CK(launch_fp64_peak(ctx->d_small + 40, blocks, iters, ctx->stream));
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaEventSynchronize(e1));
        ctx->launches += 1;
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = (double)blocks * 256.0 * (double)iters * 32.0 * 2.0;
        const double tf = flops / (ms * 1e-3) * 1e-12;
        if (rep > 0) best = std::max(best, tf);  // first launch is warm-up
        total_s += ms * 1e-3;
        if (ms < 20.f) iters *= 2;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return HB_OK;
}


/* ---- parallel tempering ------------------------------------------------------------------ */
}  // extern "C"

struct hb_pt {
    hb_ctx* ctx = nullptr;
    PtConfig cfg;
    PtConfig* d_cfg = nullptr;
    int W = 0;
    long iter = 0;
    double *x = nullptr, *y = nullptr, *logLx = nullptr, *logLy = nullptr, *logPy = nullptr;
    double *history = nullptr, *xmap = nullptr, *logLmap = nullptr, *tmp = nullptr;
    int *index = nullptr, *jump = nullptr;
    unsigned long long* counters = nullptr;
    unsigned* d_iter = nullptr;  // [0] iteration, [1] arrival ticket of k_pt_swap
    // evaluation shard (hb_pt_set_eval_shard): every rank holds the whole sampler state and proposes, accepts and
    // swaps redundantly; the likelihood -- all of the cost -- is evaluated for the walkers [eval_first,
    // eval_first + eval_count) only and the logL vector is all-gathered (chunk doubles per rank, in place)
    int shard_rank = 0, shard_world = 1;
    long eval_first = 0, eval_count = 0, chunk = 0;
    hb_comm* comm = nullptr;  // NCCL communicator of the all-gather (not owned); nullptr: host-driven exchange
    // one-launch step loop (k_pt_run) for short light curves: on by default where it applies
    bool allow_run = true;
    unsigned* d_barrier = nullptr;
    int run_max_walkers = -1;
    // one iteration captured as a CUDA graph (the step is latency-bound at the reference's sizes)
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    // kPtGraphIters iterations in one graph: one graph launch (and one inter-launch gap) per 8 steps
    cudaGraph_t graph_k = nullptr;
    cudaGraphExec_t graph_exec_k = nullptr;
    unsigned long graph_generation = 0;
};
constexpr long kPtGraphIters = 8;
constexpr int kPtMaxShards = 64;  // most ranks the evaluation of one sampler is sharded over

namespace {

void fill_limits(PtConfig& c, int use_gmag, int use_color)
{
    // set_limits (likelihood3.c:986-1121) and initialize_proposals (:1123-1179)
    static const double lo[21] = {-1.5, -1.5, -2.0, 0.0, 0.0, -kPi, 0.0, -5., -5., 0.12, 0.3, 0.12, 0.3, 0.5, 0.5,
                                  -0.3, -0.3, -5., -5., 0., 0.99};
    static const double hi[21] = {2.0, 2.0, 3.0, 1.0, kPi, kPi, 0.0, 5., 5., 0.20, 0.38, 0.20, 0.38, 1.5, 1.5,
                                  0.3, 0.3, 5., 5., 1., 1.01};
    static const double sg[21] = {1e-2, 1e-2, 1e-8, 1e-2, 1e-3, 1e-3, 1e-3, 1e-1, 1e-1, 1e-2, 1e-2,
                                  1e-2, 1e-2, 1e-2, 1e-2, 1e-2, 1e-2, 1e-1, 1e-1, 1e-3, 1e-5};
    for (int i = 0; i < 21; i++) {
        c.lo[i] = lo[i];
        c.hi[i] = hi[i];
        c.mode_lo[i] = 1;
        c.mode_hi[i] = 1;
        c.gauss[i] = (i >= 7 && i <= 18) ? 1 : 0;
        c.sigma[i] = sg[i];
    }
    c.mode_hi[3] = 0.99;  // quirk Q4
    c.mode_lo[5] = c.mode_hi[5] = 2;
    c.hi[6] = c.lc_period;
    if ((!use_color) || (!use_gmag)) {
        c.sigma[0] = c.sigma[1] = 1e-1;
        c.sigma[4] = c.sigma[5] = 1e-2;
        c.sigma[6] = 1e-3;
        for (int i = 9; i <= 18; i++) c.sigma[i] = 1e-1;
    }
}

int pt_eval_current(hb_pt* pt)
{
    hb_ctx* ctx = pt->ctx;
    int rc;
    if ((rc = ensure_chains(ctx, pt->W)) != HB_OK) return rc;
    return run_eval(ctx, pt->x, pt->W, ctx->d_t, ctx->d_fw, ctx->N, pt->logLx, nullptr);
}

}  // namespace

extern "C" {

int hb_pt_create(hb_ctx* ctx, hb_pt** out, int n_temps, int n_ens, double log_lc_period, unsigned long long seed,
                 double dtemp, int npast, int quirks)
{
    return hb_pt_create_sharded(ctx, out, n_temps, n_ens, 0, log_lc_period, seed, dtemp, npast, quirks);
}

int hb_pt_create_sharded(hb_ctx* ctx, hb_pt** out, int n_temps, int n_ens, int ens_offset, double log_lc_period,
                         unsigned long long seed, double dtemp, int npast, int quirks)
{
    if (!ctx || !out) return HB_ERR_ARG;
    *out = nullptr;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (n_temps < 1 || n_temps > kPtMaxTemps || n_ens < 1 || npast < 2 || !(dtemp > 1.0) || ens_offset < 0 ||
        ((long)ens_offset + n_ens) * n_temps > 0x7fffffffL)
        return fail_arg(ctx, "hb_pt_create: need 1 <= n_temps <= 128, n_ens >= 1, ens_offset >= 0, npast >= 2, dtemp > 1");
    if (!ctx->has_data) {
        ctx->err = "hb_pt_create: hb_set_data has not been called";
        return HB_ERR_STATE;
    }
    DeviceGuard g(ctx->device);
    hb_pt* pt = new hb_pt();
    pt->ctx = ctx;
    PtConfig& c = pt->cfg;
    std::memset(&c, 0, sizeof(c));
    c.n_temps = n_temps; c.n_ens = n_ens; c.npast = npast; c.quirks = quirks ? 1 : 0;
    c.ens_offset = ens_offset;
    c.seed = seed; c.dtemp = dtemp;
    c.temp[0] = 1.0;
    for (int i = 1; i < n_temps; i++) c.temp[i] = c.temp[i - 1] * dtemp;
    c.log_lc_period = log_lc_period;
    c.lc_period = pow(10., log_lc_period);
    c.gamma = 2.388 / sqrt(2. * kPtNpars);
    fill_limits(c, ctx->ms.use_gmag, ctx->ms.use_color);
    const int W = pt->W = n_temps * n_ens;
    pt->eval_count = pt->chunk = W;
    const size_t wd = (size_t)W * sizeof(double);
    bool ok = cudaMalloc((void**)&pt->d_cfg, sizeof(PtConfig)) == cudaSuccess &&
              cudaMalloc((void**)&pt->x, wd * kPtNpars) == cudaSuccess && cudaMalloc((void**)&pt->y, wd * kPtNpars) == cudaSuccess &&
              cudaMalloc((void**)&pt->logLx, wd) == cudaSuccess &&
              cudaMalloc((void**)&pt->logLy, wd + kPtMaxShards * sizeof(double)) == cudaSuccess &&  // padded to whole chunks
              cudaMalloc((void**)&pt->logPy, wd) == cudaSuccess && cudaMalloc((void**)&pt->tmp, wd * (kPtNpars + 1)) == cudaSuccess &&
              cudaMalloc((void**)&pt->history, wd * kPtNpars * (size_t)npast) == cudaSuccess &&
              cudaMalloc((void**)&pt->xmap, (size_t)n_ens * kPtNpars * sizeof(double)) == cudaSuccess &&
              cudaMalloc((void**)&pt->logLmap, (size_t)n_ens * sizeof(double)) == cudaSuccess &&
              cudaMalloc((void**)&pt->index, (size_t)W * sizeof(int)) == cudaSuccess &&
              cudaMalloc((void**)&pt->jump, (size_t)W * sizeof(int)) == cudaSuccess &&
              cudaMalloc((void**)&pt->counters, (size_t)n_ens * 8 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc((void**)&pt->d_iter, 2 * sizeof(unsigned)) == cudaSuccess &&
              cudaMalloc((void**)&pt->d_barrier, sizeof(unsigned)) == cudaSuccess;
    if (!ok) {
        fail_cuda(ctx, cudaGetLastError(), "hb_pt_create: cudaMalloc");
        hb_pt_destroy(pt);
        return HB_ERR_CUDA;
    }
    CK(cudaMemcpyAsync(pt->d_cfg, &c, sizeof(PtConfig), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(pt->history, 0, wd * kPtNpars * (size_t)npast, ctx->stream));
    CK(cudaMemsetAsync(pt->counters, 0, (size_t)n_ens * 8 * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(pt->xmap, 0, (size_t)n_ens * kPtNpars * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(pt->d_iter, 0, 2 * sizeof(unsigned), ctx->stream));
    std::vector<int> idx((size_t)W);
    for (int i = 0; i < W; i++) idx[i] = i % n_temps;  // index[i] = i (mcmc_wrapper2.c:333-338)
    CK(cudaMemcpyAsync(pt->index, idx.data(), (size_t)W * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<double> ninf((size_t)n_ens, -INFINITY);
    CK(cudaMemcpyAsync(pt->logLmap, ninf.data(), (size_t)n_ens * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *out = pt;
    return HB_OK;
}

void hb_pt_destroy(hb_pt* pt)
{
    if (!pt) return;
    {
        DeviceGuard g(pt->ctx->device);
        cudaStreamSynchronize(pt->ctx->stream);
        cudaFree(pt->d_cfg); cudaFree(pt->x); cudaFree(pt->y); cudaFree(pt->logLx); cudaFree(pt->logLy);
        cudaFree(pt->logPy); cudaFree(pt->tmp); cudaFree(pt->history); cudaFree(pt->xmap); cudaFree(pt->logLmap);
        cudaFree(pt->index); cudaFree(pt->jump); cudaFree(pt->counters); cudaFree(pt->d_iter); cudaFree(pt->d_barrier);
        if (pt->graph_exec) cudaGraphExecDestroy(pt->graph_exec);
        if (pt->graph) cudaGraphDestroy(pt->graph);
        if (pt->graph_exec_k) cudaGraphExecDestroy(pt->graph_exec_k);
        if (pt->graph_k) cudaGraphDestroy(pt->graph_k);
    }
    delete pt;
}

int hb_pt_init_random(hb_pt* pt)
{
    if (!pt) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    CK(launch_pt_init_random(pt->d_cfg, pt->x, pt->W, ctx->stream));
    ctx->launches += 1;
    pt->iter = 0;
    CK(cudaMemsetAsync(pt->d_iter, 0, 2 * sizeof(unsigned), ctx->stream));
    return pt_eval_current(pt);
}

int hb_pt_set_state(hb_pt* pt, const double* x)
{
    if (!pt || !x) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    int rc;
    if ((rc = upload(ctx, pt->x, x, (size_t)pt->W * kPtNpars)) != HB_OK) return rc;
    return pt_eval_current(pt);
}

// enqueue the kernels of ONE iteration on the context's stream.  phases: 1 = propose + likelihood of this rank's
// shard (+ the NCCL all-gather when a communicator is bound), 2 = accept + swaps, 3 = both
static int pt_enqueue_step(hb_pt* pt, int phases = 3)
{
    hb_ctx* ctx = pt->ctx;
    int rc;
    if (phases & 1) {
        CK(launch_pt_propose(pt->d_cfg, pt->d_iter, pt->x, pt->index, pt->history, pt->y, pt->logPy, pt->jump, pt->W, ctx->stream));
        ctx->launches += 1;
        if (pt->eval_count > 0 &&
            (rc = run_eval(ctx, pt->y + (size_t)pt->eval_first * kPtNpars, pt->eval_count, ctx->d_t, ctx->d_fw, ctx->N,
                           pt->logLy + pt->eval_first, nullptr)) != HB_OK)
            return rc;
        if (pt->shard_world > 1 && pt->comm != nullptr) {
            if (hb::comm_allgather_inplace(pt->comm, pt->logLy, (size_t)pt->chunk, ctx->stream) != HB_OK) {
                ctx->err = std::string("hb_pt_step: ") + hb_comm_last_error();
                return HB_ERR_CUDA;
            }
        }
    }
    if (phases & 2) {
        CK(launch_pt_accept(pt->d_cfg, pt->d_iter, pt->x, pt->y, pt->logLx, pt->logLy, pt->logPy, pt->jump, pt->index, pt->history,
                            pt->counters, pt->W, ctx->stream));
        CK(launch_pt_swap(pt->d_cfg, pt->d_iter, pt->index, pt->logLx, pt->x, pt->counters, pt->xmap, pt->logLmap, pt->cfg.n_ens,
                          ctx->stream));
        ctx->launches += 2;
    }
    return HB_OK;
}

int hb_pt_step(hb_pt* pt, long n_iters)
{
    if (!pt || n_iters < 0) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    int rc;
    if (pt->shard_world > 1 && pt->comm == nullptr) {
        ctx->err = "hb_pt_step: the evaluation is sharded but no communicator is bound (hb_pt_set_comm), use hb_pt_step_begin / "
                   "hb_pt_exchange_local / hb_pt_step_end";
        return HB_ERR_STATE;
    }
    if ((rc = ensure_chains(ctx, pt->W)) != HB_OK) return rc;
    // Short light curves, every walker resident at once, nothing sharded: the whole loop in ONE launch (k_pt_run)
    if (pt->allow_run && pt->shard_world == 1 && n_iters > 0 && ctx->N > 0 && ctx->N <= kPtRunMaxPoints && !ctx->time_kernels) {
        if (pt->run_max_walkers < 0) {
            CK(configure_pt_run());
            CK(pt_run_max_walkers(ctx->sm_count, &pt->run_max_walkers));
        }
        if (pt->W <= pt->run_max_walkers) {
            PtRunArgs a;
            a.cfg = pt->d_cfg; a.d_iter = pt->d_iter; a.x = pt->x; a.y = pt->y; a.logLx = pt->logLx; a.logLy = pt->logLy;
            a.logPy = pt->logPy; a.logPx = pt->tmp; a.jump = pt->jump; a.index = pt->index; a.history = pt->history; a.counters = pt->counters;
            a.xmap = pt->xmap; a.logLmap = pt->logLmap; a.tsec = ctx->d_t; a.fw = ctx->d_fw; a.N = (int)ctx->N;
            a.sctab = ctx->d_sctab; a.barrier = pt->d_barrier; a.evaluated = ctx->d_evaluated; a.n_iters = n_iters; a.ms = ctx->ms;
            CK(cudaMemsetAsync(pt->d_barrier, 0, sizeof(unsigned), ctx->stream));
            CK(launch_pt_run(a, pt->W, ctx->stream));
            ctx->launches += 1;
            pt->iter += n_iters;
            return HB_OK;
        }
    }
    // Several iterations in one call: replay a captured graph of one iteration (6 nodes) instead of
    // issuing 6 launches per iteration.  The graph is re-captured when the context's buffers or
    // by-value kernel arguments changed (generation), never while kernel timing is on.
    const bool use_graph = n_iters >= 4 && !ctx->time_kernels;
    if (use_graph && (pt->graph_exec == nullptr || pt->graph_generation != ctx->generation)) {
        if (pt->graph_exec) { cudaGraphExecDestroy(pt->graph_exec); pt->graph_exec = nullptr; }
        if (pt->graph) { cudaGraphDestroy(pt->graph); pt->graph = nullptr; }
        const long launches_before = ctx->launches;
        CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        rc = pt_enqueue_step(pt);
        cudaError_t ce = cudaStreamEndCapture(ctx->stream, &pt->graph);
        ctx->launches = launches_before;  // capturing launches nothing
        if (rc != HB_OK) return rc;
        if (ce != cudaSuccess) return fail_cuda(ctx, ce, "cudaStreamEndCapture");
        CK(cudaGraphInstantiate(&pt->graph_exec, pt->graph, 0));
        // the same iteration kPtGraphIters times over (the iteration counter lives on the device)
        if (pt->graph_exec_k) { cudaGraphExecDestroy(pt->graph_exec_k); pt->graph_exec_k = nullptr; }
        if (pt->graph_k) { cudaGraphDestroy(pt->graph_k); pt->graph_k = nullptr; }
        CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        for (long k = 0; k < kPtGraphIters && rc == HB_OK; k++) rc = pt_enqueue_step(pt);
        ce = cudaStreamEndCapture(ctx->stream, &pt->graph_k);
        ctx->launches = launches_before;
        if (rc != HB_OK) return rc;
        if (ce != cudaSuccess) return fail_cuda(ctx, ce, "cudaStreamEndCapture");
        CK(cudaGraphInstantiate(&pt->graph_exec_k, pt->graph_k, 0));
        pt->graph_generation = ctx->generation;
    }
    long k = 0;
    if (use_graph)
        for (; k + kPtGraphIters <= n_iters; k += kPtGraphIters) {
            CK(cudaGraphLaunch(pt->graph_exec_k, ctx->stream));
            ctx->launches += 5 * kPtGraphIters;
            pt->iter += kPtGraphIters;
        }
    for (; k < n_iters; k++) {
        if (use_graph) {
            CK(cudaGraphLaunch(pt->graph_exec, ctx->stream));
            ctx->launches += 5;
        } else if ((rc = pt_enqueue_step(pt)) != HB_OK) {
            return rc;
        }
        pt->iter++;
    }
    return HB_OK;
}

int hb_pt_set_one_launch(hb_pt* pt, int enable)
{
    if (!pt) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(pt->ctx->mu);
    pt->allow_run = enable != 0;
    return HB_OK;
}

int hb_pt_set_eval_shard(hb_pt* pt, int rank, int world)
{
    if (!pt) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (world < 1 || world > kPtMaxShards || rank < 0 || rank >= world)
        return fail_arg(ctx, "hb_pt_set_eval_shard: need 0 <= rank < world <= 64");
    pt->shard_rank = rank;
    pt->shard_world = world;
    pt->chunk = (pt->W + world - 1) / world;  // walkers per rank; the logL buffer is padded to world whole chunks
    pt->eval_first = std::min<long>((long)rank * pt->chunk, pt->W);
    pt->eval_count = std::min<long>(pt->chunk, pt->W - pt->eval_first);
    ctx->generation++;  // captured steps are stale
    return HB_OK;
}

int hb_pt_set_comm(hb_pt* pt, hb_comm* comm)
{
    if (!pt) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (comm && (hb::comm_rank(comm) != pt->shard_rank || hb::comm_world(comm) != pt->shard_world))
        return fail_arg(ctx, "hb_pt_set_comm: the communicator's rank / size differ from the evaluation shard's");
    pt->comm = comm;
    ctx->generation++;
    return HB_OK;
}

int hb_pt_get_eval_shard(const hb_pt* pt, long* first, long* count, long* chunk)
{
    if (!pt) return HB_ERR_ARG;
    if (first) *first = pt->eval_first;
    if (count) *count = pt->eval_count;
    if (chunk) *chunk = pt->chunk;
    return HB_OK;
}

int hb_pt_step_begin(hb_pt* pt)
{
    if (!pt) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    int rc;
    if ((rc = ensure_chains(ctx, pt->W)) != HB_OK) return rc;
    return pt_enqueue_step(pt, 1);
}

int hb_pt_step_end(hb_pt* pt)
{
    if (!pt) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    int rc = pt_enqueue_step(pt, 2);
    if (rc == HB_OK) pt->iter++;
    return rc;
}

int hb_pt_exchange_local(hb_pt** pts, int n)
{
    if (!pts || n < 1) return HB_ERR_ARG;
    // every sampler's shard of the proposal logL into every other sampler's vector (same process: device-to-device
    // or peer copies), then all streams are drained: the host-driven stand-in for the all-gather
    for (int i = 0; i < n; i++) {
        if (!pts[i] || pts[i]->shard_world != n || pts[i]->shard_rank != i || pts[i]->W != pts[0]->W) return HB_ERR_ARG;
    }
    for (int i = 0; i < n; i++) {
        hb_ctx* ctx = pts[i]->ctx;
        DeviceGuard g(ctx->device);
        CK(cudaStreamSynchronize(ctx->stream));
    }
    for (int i = 0; i < n; i++) {
        hb_pt* src = pts[i];
        if (src->eval_count == 0) continue;
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            hb_ctx* ctx = pts[j]->ctx;
            DeviceGuard g(ctx->device);
            CK(cudaMemcpyPeerAsync(pts[j]->logLy + src->eval_first, ctx->device, src->logLy + src->eval_first, src->ctx->device,
                                   (size_t)src->eval_count * sizeof(double), ctx->stream));
        }
    }
    for (int i = 0; i < n; i++) {
        hb_ctx* ctx = pts[i]->ctx;
        DeviceGuard g(ctx->device);
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return HB_OK;
}

long hb_pt_iteration(const hb_pt* pt) { return pt ? pt->iter : -1; }

int hb_pt_get_state(hb_pt* pt, double* x, double* logL, int* index)
{
    if (!pt) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    int rc;
    if (x && (rc = download(ctx, x, pt->x, (size_t)pt->W * kPtNpars)) != HB_OK) return rc;
    if (logL && (rc = download(ctx, logL, pt->logLx, (size_t)pt->W)) != HB_OK) return rc;
    if (index) {
        CK(cudaMemcpyAsync(index, pt->index, (size_t)pt->W * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return HB_OK;
}

int hb_pt_get_proposal(hb_pt* pt, double* y, double* logLy, double* logPy)
{
    if (!pt) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    int rc;
    if (y && (rc = download(ctx, y, pt->y, (size_t)pt->W * kPtNpars)) != HB_OK) return rc;
    if (logLy && (rc = download(ctx, logLy, pt->logLy, (size_t)pt->W)) != HB_OK) return rc;
    if (logPy && (rc = download(ctx, logPy, pt->logPy, (size_t)pt->W)) != HB_OK) return rc;
    return HB_OK;
}

int hb_pt_get_cold(hb_pt* pt, double* x_cold, double* logL_cold)
{
    if (!pt) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const int E = pt->cfg.n_ens;
    CK(launch_pt_gather_cold(pt->d_cfg, pt->index, pt->x, pt->logLx, pt->tmp, pt->tmp + (size_t)E * kPtNpars, E, ctx->stream));
    ctx->launches += 1;
    int rc;
    if (x_cold && (rc = download(ctx, x_cold, pt->tmp, (size_t)E * kPtNpars)) != HB_OK) return rc;
    if (logL_cold && (rc = download(ctx, logL_cold, pt->tmp + (size_t)E * kPtNpars, (size_t)E)) != HB_OK) return rc;
    return HB_OK;
}

int hb_pt_get_logL_by_rung(hb_pt* pt, double* out)
{
    if (!pt || !out) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    CK(launch_pt_logL_by_rung(pt->d_cfg, pt->index, pt->logLx, pt->tmp, pt->W, ctx->stream));
    ctx->launches += 1;
    return download(ctx, out, pt->tmp, (size_t)pt->W);
}

int hb_pt_get_map(hb_pt* pt, double* xmap, double* logLmap)
{
    if (!pt) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    int rc;
    if (xmap && (rc = download(ctx, xmap, pt->xmap, (size_t)pt->cfg.n_ens * kPtNpars)) != HB_OK) return rc;
    if (logLmap && (rc = download(ctx, logLmap, pt->logLmap, (size_t)pt->cfg.n_ens)) != HB_OK) return rc;
    return HB_OK;
}

int hb_pt_get_counters(hb_pt* pt, unsigned long long* out)
{
    if (!pt || !out) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    CK(cudaMemcpyAsync(out, pt->counters, (size_t)pt->cfg.n_ens * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

void* hb_pt_device_logL(hb_pt* pt) { return pt ? (void*)pt->logLx : nullptr; }

int hb_pt_cold_logL_dev(hb_pt* pt, double* d_out)
{
    if (!pt || !d_out) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const int E = pt->cfg.n_ens;
    CK(launch_pt_gather_cold(pt->d_cfg, pt->index, pt->x, pt->logLx, pt->tmp, d_out, E, ctx->stream));
    ctx->launches += 1;
    return HB_OK;
}


/* ---- Gaia-colour sampler (GAIA_mcmc.c) ------------------------------------------------------ */
}  // extern "C"

struct hb_gaia_pt {
    hb_ctx* ctx = nullptr;
    GaiaPtConfig cfg;
    GaiaPtConfig* d_cfg = nullptr;
    GaiaPtArrays a;
    double* d_obs = nullptr;  // D[E], data[E][4], err[E][4]
    int W = 0;
    long iter = 0;
    bool has_data = false, has_state = false;
    long rec_cap = 0;  // capacity (records per ensemble) of the device log buffers
};

namespace {

int gaia_grow_rec(hb_gaia_pt* pt, long need)
{
    hb_ctx* ctx = pt->ctx;
    if (need <= pt->rec_cap) return HB_OK;
    cudaFree(pt->a.rec_chain);
    cudaFree(pt->a.rec_logL);
    pt->a.rec_chain = pt->a.rec_logL = nullptr;
    pt->rec_cap = 0;
    const size_t E = (size_t)pt->cfg.n_ens;
    CK(cudaMalloc((void**)&pt->a.rec_chain, E * (size_t)need * (kGaiaNpars + 1) * sizeof(double)));
    CK(cudaMalloc((void**)&pt->a.rec_logL, E * (size_t)need * (size_t)pt->cfg.n_temps * sizeof(double)));
    pt->rec_cap = need;
    return HB_OK;
}

}  // namespace

extern "C" {

int hb_gaia_pt_create(hb_ctx* ctx, hb_gaia_pt** out, int n_temps, int n_ens, unsigned long long seed, double dtemp,
                      int npast)
{
    if (!ctx || !out) return HB_ERR_ARG;
    *out = nullptr;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (n_temps < 1 || n_temps > kGaiaMaxTemps || n_ens < 1 || npast < 2 || !(dtemp > 1.0))
        return fail_arg(ctx, "hb_gaia_pt_create: need 1 <= n_temps <= 32, n_ens >= 1, npast >= 2, dtemp > 1");
    DeviceGuard g(ctx->device);
    hb_gaia_pt* pt = new hb_gaia_pt();
    pt->ctx = ctx;
    std::memset(&pt->a, 0, sizeof(pt->a));
    GaiaPtConfig& c = pt->cfg;
    std::memset(&c, 0, sizeof(c));
    c.n_temps = n_temps; c.n_ens = n_ens; c.npast = npast; c.seed = seed;
    c.gamma = 2.388 / sqrt(2. * kGaiaNpars);
    c.temp[0] = 1.0;
    for (int i = 1; i < n_temps; i++) c.temp[i] = c.temp[i - 1] * dtemp;
    // set_limits (GAIA_mcmc.c:346-390) and init_proposals (:449-458).  The reference sets sigma[0]
    // and sigma[1] only; sigma[2..5] are read from a fresh malloc block, i.e. 0 in practice: the shape
    // parameters move through DE jumps alone.  hb_gaia_pt_set_sigma overrides.
    for (int i = 0; i < kGaiaNpars; i++) {
        c.lo[i] = (i < 2) ? -1.5 : -3.;
        c.hi[i] = (i < 2) ? 2.0 : 3.;
        c.mode_lo[i] = c.mode_hi[i] = 1;
        c.gauss[i] = (i < 2) ? 0 : 1;
        c.sigma[i] = (i < 2) ? 1.e-2 : 0.;
    }
    const int W = pt->W = n_temps * n_ens;
    const size_t wd = (size_t)W * sizeof(double), E = (size_t)n_ens;
    GaiaPtArrays& a = pt->a;
    bool ok = cudaMalloc((void**)&pt->d_cfg, sizeof(GaiaPtConfig)) == cudaSuccess &&
              cudaMalloc((void**)&a.x, wd * kGaiaNpars) == cudaSuccess && cudaMalloc((void**)&a.logL, wd) == cudaSuccess &&
              cudaMalloc((void**)&a.index, (size_t)W * sizeof(int)) == cudaSuccess &&
              cudaMalloc((void**)&a.history, wd * kGaiaNpars * (size_t)npast) == cudaSuccess &&
              cudaMalloc((void**)&a.xmap, E * kGaiaNpars * sizeof(double)) == cudaSuccess &&
              cudaMalloc((void**)&a.logLmap, E * sizeof(double)) == cudaSuccess &&
              cudaMalloc((void**)&a.counters, E * 8 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc((void**)&pt->d_obs, E * 9 * sizeof(double)) == cudaSuccess &&
              cudaMalloc((void**)&a.last_y, wd * kGaiaNpars) == cudaSuccess && cudaMalloc((void**)&a.last_logLy, wd) == cudaSuccess &&
              cudaMalloc((void**)&a.last_logPy, wd) == cudaSuccess && cudaMalloc((void**)&a.last_jump, (size_t)W * sizeof(int)) == cudaSuccess;
    if (!ok) {
        fail_cuda(ctx, cudaGetLastError(), "hb_gaia_pt_create: cudaMalloc");
        hb_gaia_pt_destroy(pt);
        return HB_ERR_CUDA;
    }
    a.D = pt->d_obs;
    a.data = pt->d_obs + E;
    a.err = pt->d_obs + 5 * E;
    CK(cudaMemcpyAsync(pt->d_cfg, &c, sizeof(c), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(a.history, 0, wd * kGaiaNpars * (size_t)npast, ctx->stream));
    CK(cudaMemsetAsync(a.counters, 0, E * 8 * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(a.last_y, 0, wd * kGaiaNpars, ctx->stream));
    CK(cudaMemsetAsync(a.last_logLy, 0, wd, ctx->stream));
    CK(cudaMemsetAsync(a.last_logPy, 0, wd, ctx->stream));
    CK(cudaMemsetAsync(a.last_jump, 0, (size_t)W * sizeof(int), ctx->stream));
    std::vector<int> idx((size_t)W);
    for (int i = 0; i < W; i++) idx[i] = i % n_temps;  // index[i] = i (GAIA_mcmc.c:478-484)
    CK(cudaMemcpyAsync(a.index, idx.data(), (size_t)W * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *out = pt;
    return HB_OK;
}

void hb_gaia_pt_destroy(hb_gaia_pt* pt)
{
    if (!pt) return;
    {
        DeviceGuard g(pt->ctx->device);
        cudaStreamSynchronize(pt->ctx->stream);
        GaiaPtArrays& a = pt->a;
        cudaFree(pt->d_cfg); cudaFree(a.x); cudaFree(a.logL); cudaFree(a.index); cudaFree(a.history); cudaFree(a.xmap);
        cudaFree(a.logLmap); cudaFree(a.counters); cudaFree(pt->d_obs); cudaFree(a.last_y); cudaFree(a.last_logLy);
        cudaFree(a.last_logPy); cudaFree(a.last_jump); cudaFree(a.rec_chain); cudaFree(a.rec_logL);
    }
    delete pt;
}

int hb_gaia_pt_set_data(hb_gaia_pt* pt, const double* D, const double* data, const double* err)
{
    if (!pt || !D || !data || !err) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const size_t E = (size_t)pt->cfg.n_ens;
    std::vector<double> obs(E * 9);
    std::memcpy(obs.data(), D, E * sizeof(double));
    std::memcpy(obs.data() + E, data, 4 * E * sizeof(double));
    std::memcpy(obs.data() + 5 * E, err, 4 * E * sizeof(double));
    int rc;
    if ((rc = upload(ctx, pt->d_obs, obs.data(), E * 9)) != HB_OK) return rc;
    pt->has_data = true;
    if (pt->has_state) {  // new data: the cached likelihoods are stale
        CK(launch_gaia_pt_eval(pt->d_cfg, pt->a, pt->W, ctx->stream));
        ctx->launches += 1;
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return HB_OK;
}

int hb_gaia_pt_set_sigma(hb_gaia_pt* pt, const double* sigma)
{
    if (!pt || !sigma) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    for (int i = 0; i < kGaiaNpars; i++) pt->cfg.sigma[i] = sigma[i];
    CK(cudaMemcpyAsync(pt->d_cfg, &pt->cfg, sizeof(pt->cfg), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

static int gaia_after_state(hb_gaia_pt* pt)
{
    hb_ctx* ctx = pt->ctx;
    std::vector<int> idx((size_t)pt->W);
    for (int i = 0; i < pt->W; i++) idx[i] = i % pt->cfg.n_temps;
    CK(cudaMemcpyAsync(pt->a.index, idx.data(), (size_t)pt->W * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_gaia_pt_eval(pt->d_cfg, pt->a, pt->W, ctx->stream));
    ctx->launches += 1;
    CK(cudaStreamSynchronize(ctx->stream));
    pt->iter = 0;
    pt->has_state = true;
    return HB_OK;
}

int hb_gaia_pt_init_random(hb_gaia_pt* pt)
{
    if (!pt) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!pt->has_data) {
        ctx->err = "hb_gaia_pt_init_random: hb_gaia_pt_set_data has not been called";
        return HB_ERR_STATE;
    }
    DeviceGuard g(ctx->device);
    CK(launch_gaia_pt_init(pt->d_cfg, pt->a.x, pt->W, ctx->stream));
    ctx->launches += 1;
    return gaia_after_state(pt);
}

int hb_gaia_pt_set_state(hb_gaia_pt* pt, const double* x)
{
    if (!pt || !x) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!pt->has_data) {
        ctx->err = "hb_gaia_pt_set_state: hb_gaia_pt_set_data has not been called";
        return HB_ERR_STATE;
    }
    DeviceGuard g(ctx->device);
    int rc;
    if ((rc = upload(ctx, pt->a.x, x, (size_t)pt->W * kGaiaNpars)) != HB_OK) return rc;
    return gaia_after_state(pt);
}

long hb_gaia_pt_records(const hb_gaia_pt* pt, long n_iters, int thin)
{
    if (!pt || n_iters < 0 || thin <= 0) return 0;
    const long i0 = pt->iter, i1 = pt->iter + n_iters;  // records at iterations i0 <= it < i1 with it % thin == 0
    return (i1 + thin - 1) / thin - (i0 + thin - 1) / thin;
}

int hb_gaia_pt_run(hb_gaia_pt* pt, long n_iters, int thin, double* chain, double* logL_by_rung)
{
    if (!pt || n_iters < 0) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!pt->has_state) {
        ctx->err = "hb_gaia_pt_run: no state (hb_gaia_pt_init_random / hb_gaia_pt_set_state)";
        return HB_ERR_STATE;
    }
    if ((chain || logL_by_rung) && thin <= 0) return fail_arg(ctx, "hb_gaia_pt_run: thin must be positive when a log is requested");
    if (pt->iter + n_iters > 0xFFFFFFF0L) return fail_arg(ctx, "hb_gaia_pt_run: iteration counter would exceed 32 bits");
    DeviceGuard g(ctx->device);
    const bool want = chain || logL_by_rung;
    const long nrec = want ? hb_gaia_pt_records(pt, n_iters, thin) : 0;
    int rc;
    if (nrec > 0 && (rc = gaia_grow_rec(pt, nrec)) != HB_OK) return rc;
    CK(launch_gaia_pt_run(pt->d_cfg, pt->a, pt->cfg.n_ens, (unsigned)pt->iter, (unsigned)n_iters, want ? thin : 0,
                          pt->rec_cap, ctx->stream));
    if (n_iters > 0) ctx->launches += 1;
    pt->iter += n_iters;
    const size_t E = (size_t)pt->cfg.n_ens, T = (size_t)pt->cfg.n_temps;
    if (nrec > 0) {
        // device layout [E][rec_cap][..] -> caller layout [E][nrec][..]
        for (size_t e = 0; e < E; e++) {
            if (chain) CK(cudaMemcpyAsync(chain + e * nrec * (kGaiaNpars + 1), pt->a.rec_chain + e * pt->rec_cap * (kGaiaNpars + 1),
                                          (size_t)nrec * (kGaiaNpars + 1) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            if (logL_by_rung) CK(cudaMemcpyAsync(logL_by_rung + e * nrec * T, pt->a.rec_logL + e * pt->rec_cap * T,
                                                 (size_t)nrec * T * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

long hb_gaia_pt_iteration(const hb_gaia_pt* pt) { return pt ? pt->iter : -1; }

int hb_gaia_pt_get_state(hb_gaia_pt* pt, double* x, double* logL, int* index)
{
    if (!pt) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    int rc;
    if (x && (rc = download(ctx, x, pt->a.x, (size_t)pt->W * kGaiaNpars)) != HB_OK) return rc;
    if (logL && (rc = download(ctx, logL, pt->a.logL, (size_t)pt->W)) != HB_OK) return rc;
    if (index) {
        CK(cudaMemcpyAsync(index, pt->a.index, (size_t)pt->W * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return HB_OK;
}

int hb_gaia_pt_get_proposal(hb_gaia_pt* pt, double* y, double* logLy, double* logPy, int* jump)
{
    if (!pt) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    int rc;
    if (y && (rc = download(ctx, y, pt->a.last_y, (size_t)pt->W * kGaiaNpars)) != HB_OK) return rc;
    if (logLy && (rc = download(ctx, logLy, pt->a.last_logLy, (size_t)pt->W)) != HB_OK) return rc;
    if (logPy && (rc = download(ctx, logPy, pt->a.last_logPy, (size_t)pt->W)) != HB_OK) return rc;
    if (jump) {
        CK(cudaMemcpyAsync(jump, pt->a.last_jump, (size_t)pt->W * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return HB_OK;
}

int hb_gaia_pt_get_history(hb_gaia_pt* pt, double* history)
{
    if (!pt || !history) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    return download(ctx, history, pt->a.history, (size_t)pt->W * (size_t)pt->cfg.npast * kGaiaNpars);
}

int hb_gaia_pt_get_map(hb_gaia_pt* pt, double* xmap, double* logLmap)
{
    if (!pt) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    int rc;
    if (xmap && (rc = download(ctx, xmap, pt->a.xmap, (size_t)pt->cfg.n_ens * kGaiaNpars)) != HB_OK) return rc;
    if (logLmap && (rc = download(ctx, logLmap, pt->a.logLmap, (size_t)pt->cfg.n_ens)) != HB_OK) return rc;
    return HB_OK;
}

int hb_gaia_pt_get_counters(hb_gaia_pt* pt, unsigned long long* out)
{
    if (!pt || !out) return HB_ERR_ARG;
    hb_ctx* ctx = pt->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    CK(cudaMemcpyAsync(out, pt->a.counters, (size_t)pt->cfg.n_ens * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

}  // extern "C"
