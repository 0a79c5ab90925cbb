// Result gather over NCCL inside the C ABI (bgp_gather_results).
//
// Work items are independent, so the only exchange of a multi-GPU sweep is the final collection of the per-item results
// (ELBO, hyper-gradients, status) -- a few KB per rank over NVLink.  The Python layer does it with torch.distributed
// (branchedgp_b200/sharding.py); these entry points let a host program that has only this C library do the same: every rank
// creates a handle on its GPU, rank 0 obtains a 128-byte NCCL id and ships it to the others by whatever means it has (file, MPI,
// environment), every rank calls bgp_comm_init, and bgp_gather_results all-gathers fixed-size row blocks.
// libnccl is opened at run time (dlopen), so the library has no link-time dependency on it and loads on boxes without NCCL.
#include <dlfcn.h>

#include "bgp_internal.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };   // ncclDataType_t value of double (nccl.h: ncclInt8 0 ... ncclFloat16 6, ncclFloat32 7, ncclFloat64 8)

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* nccl_api(bgp_handle* h) {
    static NcclApi api;
    if (api.lib) return &api;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) { fail(h, BGP_ERR_ARG, "libnccl.so.2 not found: %s", dlerror()); return nullptr; }
    api.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.lib, "ncclCommDestroy");
    api.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(api.lib, "ncclAllGather");
    api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather) {
        fail(h, BGP_ERR_ARG, "libnccl lacks an expected symbol");
        dlclose(api.lib);
        api.lib = nullptr;
        return nullptr;
    }
    return &api;
}

int nccl_fail(bgp_handle* h, NcclApi* a, const char* what, int rc) {
    return fail(h, BGP_ERR_CUDA, "%s: %s", what, a->GetErrorString ? a->GetErrorString(rc) : "NCCL error");
}

}  // namespace

extern "C" {

int bgp_comm_unique_id(bgp_handle* h, void* id128) {
    if (!h || !id128) return BGP_ERR_ARG;
    NcclApi* a = nccl_api(h);
    if (!a) return BGP_ERR_ARG;
    ncclUniqueId id;
    const int rc = a->GetUniqueId(&id);
    if (rc != ncclSuccess) return nccl_fail(h, a, "ncclGetUniqueId", rc);
    memcpy(id128, &id, sizeof(id));
    return BGP_OK;
}

int bgp_comm_init(bgp_handle* h, int nranks, int rank, const void* id128) {
    if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return BGP_ERR_ARG;
    NcclApi* a = nccl_api(h);
    if (!a) return BGP_ERR_ARG;
    CU(cudaSetDevice(h->device));
    if (h->comm) { a->CommDestroy((ncclComm_t)h->comm); h->comm = nullptr; }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    const int rc = a->CommInitRank(&comm, nranks, id, rank);
    if (rc != ncclSuccess) return nccl_fail(h, a, "ncclCommInitRank", rc);
    h->comm = comm; h->comm_ranks = nranks; h->comm_rank = rank;
    return BGP_OK;
}

int bgp_comm_destroy(bgp_handle* h) {
    if (!h) return BGP_ERR_ARG;
    if (h->comm) {
        NcclApi* a = nccl_api(h);
        if (a) a->CommDestroy((ncclComm_t)h->comm);
        h->comm = nullptr;
    }
    return BGP_OK;
}

// all[r][0 .. rows_max) = rank r's `rows_local` rows of `k` doubles (rows beyond rows_local are zero).  Host pointers.
int bgp_gather_results(bgp_handle* h, const double* local, int rows_local, int k, int rows_max, double* all) {
    if (!h || !all || k < 1 || rows_max < 1 || rows_local < 0 || rows_local > rows_max || (rows_local > 0 && !local)) return BGP_ERR_ARG;
    if (!h->comm) return fail(h, BGP_ERR_ARG, "bgp_comm_init has not been called on this handle");
    NcclApi* a = nccl_api(h);
    if (!a) return BGP_ERR_ARG;
    CU(cudaSetDevice(h->device));
    const size_t blk = (size_t)rows_max * k;
    const size_t need = (1 + (size_t)h->comm_ranks) * blk * 8;
    if (h->comm_buf_bytes < need) {
        if (h->comm_buf) cudaFree(h->comm_buf);
        h->comm_buf = nullptr; h->comm_buf_bytes = 0;
        CU(cudaMalloc(&h->comm_buf, need));
        h->comm_buf_bytes = need;
    }
    double* d_local = (double*)h->comm_buf;
    double* d_all = d_local + blk;
    cudaStream_t st = 0;
    CU(cudaMemsetAsync(d_local, 0, blk * 8, st));
    if (rows_local > 0) CU(cudaMemcpyAsync(d_local, local, (size_t)rows_local * k * 8, cudaMemcpyHostToDevice, st));
    const int rc = a->AllGather(d_local, d_all, blk, ncclFloat64, (ncclComm_t)h->comm, st);
    if (rc != ncclSuccess) return nccl_fail(h, a, "ncclAllGather", rc);
    CU(cudaMemcpyAsync(all, d_all, (size_t)h->comm_ranks * blk * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return BGP_OK;
}

}  // extern "C"
