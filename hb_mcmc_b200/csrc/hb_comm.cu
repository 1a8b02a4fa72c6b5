// hb_comm.cu -- the one exchange of the multi-GPU sampler: an all-gather of the per-step log-likelihood vector
// (8 bytes per walker) over NCCL / NVLink, issued from C on the context's stream so that it can be captured in
// the step's CUDA graph.  NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy a hosting PyTorch process
// has already loaded, else the system one), so the library has no NCCL dependency unless a communicator is made.
// The handful of entry points used are declared here as NCCL's public header declares them (nccl.h, 2.x ABI).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/hb_b200.h"
#include "hb_comm.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;  // ncclSuccess == 0
constexpr int kNcclFloat64 = 8;  // ncclDataType_t: ncclFloat64 / ncclDouble

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string error;
};

NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("HB_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = std::string("libnccl.so.2 could not be loaded (set HB_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
            return;
        }
        auto sym = [&](const char* s) {
            void* p = dlsym(api.handle, s);
            if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + s;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    });
    return api;
}

std::string g_comm_error;
std::mutex g_comm_mu;

int fail(const std::string& s)
{
    std::lock_guard<std::mutex> g(g_comm_mu);
    g_comm_error = s;
    return HB_ERR_CUDA;
}

int check(ncclResult_t r, const char* what)
{
    if (r == 0) return HB_OK;
    NcclApi& a = nccl();
    return fail(std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(r) : "NCCL error"));
}

}  // namespace

struct hb_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
};

namespace hb {

int comm_rank(const hb_comm* c) { return c ? c->rank : 0; }
int comm_world(const hb_comm* c) { return c ? c->world : 1; }

// in-place all-gather: every rank's `count` doubles at buf + rank * count
int comm_allgather_inplace(hb_comm* c, double* d_buf, size_t count, cudaStream_t stream)
{
    if (!c || c->world == 1) return HB_OK;
    NcclApi& a = nccl();
    return check(a.AllGather(d_buf + (size_t)c->rank * count, d_buf, count, kNcclFloat64, c->comm, stream), "ncclAllGather");
}

}  // namespace hb

extern "C" {

const char* hb_comm_last_error(void) { return g_comm_error.c_str(); }

int hb_comm_nccl_version(int* version)
{
    NcclApi& a = nccl();
    if (!a.error.empty() || !a.GetVersion) return fail(a.error.empty() ? "ncclGetVersion missing" : a.error);
    return check(a.GetVersion(version), "ncclGetVersion");
}

int hb_comm_unique_id(unsigned char id[HB_COMM_ID_BYTES])
{
    if (!id) return HB_ERR_ARG;
    NcclApi& a = nccl();
    if (!a.error.empty()) return fail(a.error);
    ncclUniqueId u;
    int rc = check(a.GetUniqueId(&u), "ncclGetUniqueId");
    if (rc == HB_OK) std::memcpy(id, u.internal, HB_COMM_ID_BYTES);
    return rc;
}

int hb_comm_create(hb_comm** out, int device, const unsigned char id[HB_COMM_ID_BYTES], int rank, int world)
{
    if (!out || !id || world < 1 || rank < 0 || rank >= world) return HB_ERR_ARG;
    *out = nullptr;
    NcclApi& a = nccl();
    if (!a.error.empty()) return fail(a.error);
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return fail("hb_comm_create: cudaSetDevice failed");
    ncclUniqueId u;
    std::memcpy(u.internal, id, HB_COMM_ID_BYTES);
    hb_comm* c = new hb_comm();
    c->rank = rank; c->world = world; c->device = device;
    int rc = check(a.CommInitRank(&c->comm, world, u, rank), "ncclCommInitRank");
    if (prev >= 0) cudaSetDevice(prev);
    if (rc != HB_OK) { delete c; return rc; }
    *out = c;
    return HB_OK;
}

int hb_comm_create_all(hb_comm** out, const int* devices, int n)
{
    if (!out || !devices || n < 1) return HB_ERR_ARG;
    NcclApi& a = nccl();
    if (!a.error.empty()) return fail(a.error);
    ncclComm_t* comms = new ncclComm_t[n];
    int rc = check(a.CommInitAll(comms, n, devices), "ncclCommInitAll");
    if (rc == HB_OK)
        for (int i = 0; i < n; i++) {
            out[i] = new hb_comm();
            out[i]->comm = comms[i]; out[i]->rank = i; out[i]->world = n; out[i]->device = devices[i];
        }
    delete[] comms;
    return rc;
}

void hb_comm_destroy(hb_comm* c)
{
    if (!c) return;
    NcclApi& a = nccl();
    if (c->comm && a.CommDestroy) a.CommDestroy(c->comm);
    delete c;
}

int hb_comm_rank(const hb_comm* c) { return c ? c->rank : -1; }
int hb_comm_world(const hb_comm* c) { return c ? c->world : -1; }

int hb_comm_allgather_f64(hb_comm* c, double* d_buf, long count_per_rank, void* cuda_stream)
{
    if (!c || !d_buf || count_per_rank < 0) return HB_ERR_ARG;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != c->device) cudaSetDevice(c->device);
    int rc = hb::comm_allgather_inplace(c, d_buf, (size_t)count_per_rank, (cudaStream_t)cuda_stream);
    if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
    return rc;
}

}  // extern "C"
