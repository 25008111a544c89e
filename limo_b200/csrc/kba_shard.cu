// kba_shard.cu -- NCCL plumbing of the sharded window solve (include/kba_b200.h, "ONE large window sharded ...").
// libnccl is opened at run time (dlopen) so that single-GPU users need no NCCL, and so that inside a PyTorch process
// the library torch already loaded is the one used.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <string>

#include <cuda_runtime.h>
#include <nccl.h>  // types only: no symbol of libnccl is linked

#include "kba_b200.h"
#include "kba_kernels.h"

extern "C" {
int kba_internal_stream(kba_handle* h, cudaStream_t* s, int* device);
int kba_internal_fail(int code, const char* msg);
}

namespace {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) get_unique_id = nullptr;
    decltype(&ncclCommInitRank) comm_init_rank = nullptr;
    decltype(&ncclAllReduce) all_reduce = nullptr;
    decltype(&ncclCommDestroy) comm_destroy = nullptr;
    decltype(&ncclGetErrorString) error_string = nullptr;
    bool ok() const { return lib && get_unique_id && comm_init_rank && all_reduce && comm_destroy && error_string; }
};

NcclApi& api() {
    static NcclApi a;
    if (a.lib) return a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) return a;
    a.get_unique_id = (decltype(a.get_unique_id))dlsym(a.lib, "ncclGetUniqueId");
    a.comm_init_rank = (decltype(a.comm_init_rank))dlsym(a.lib, "ncclCommInitRank");
    a.all_reduce = (decltype(a.all_reduce))dlsym(a.lib, "ncclAllReduce");
    a.comm_destroy = (decltype(a.comm_destroy))dlsym(a.lib, "ncclCommDestroy");
    a.error_string = (decltype(a.error_string))dlsym(a.lib, "ncclGetErrorString");
    return a;
}

int nccl_fail(const char* what, ncclResult_t r) {
    std::string m = std::string(what) + ": " + (api().error_string ? api().error_string(r) : "nccl error");
    return kba_internal_fail(KBA_ERR_NCCL, m.c_str());
}

}  // namespace

struct kba_shard_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
};

static int shard_allreduce(void* user, const double* send, double* recv, long long count, int op, cudaStream_t s) {
    kba_shard_comm* c = (kba_shard_comm*)user;
    const ncclResult_t r = api().all_reduce(send, recv, (size_t)count, ncclDouble, op == 1 ? ncclMax : ncclSum, c->comm, s);
    return r == ncclSuccess ? 0 : nccl_fail("ncclAllReduce", r);
}

// used by kba_api.cu
kba::Exchange kba_shard_exchange(kba_shard_comm* c) {
    kba::Exchange x;
    x.allreduce = &shard_allreduce;
    x.user = c;
    x.rank = c->rank; x.world = c->world;
    return x;
}

extern "C" {

int kba_shard_unique_id(void* id_out) {
    static_assert(sizeof(ncclUniqueId) == KBA_SHARD_ID_BYTES, "NCCL unique id size");
    if (!id_out) return kba_internal_fail(KBA_ERR_BAD_ARG, "null id buffer");
    if (!api().ok()) return kba_internal_fail(KBA_ERR_NCCL, "libnccl.so.2 not found");
    ncclUniqueId id;
    const ncclResult_t r = api().get_unique_id(&id);
    if (r != ncclSuccess) return nccl_fail("ncclGetUniqueId", r);
    memcpy(id_out, &id, sizeof id);
    return KBA_OK;
}

int kba_shard_comm_create(kba_handle* h, int32_t rank, int32_t world, const void* id, kba_shard_comm** out) {
    if (!h || !id || !out || world < 1 || rank < 0 || rank >= world) return kba_internal_fail(KBA_ERR_BAD_ARG, "bad argument to kba_shard_comm_create");
    if (!api().ok()) return kba_internal_fail(KBA_ERR_NCCL, "libnccl.so.2 not found");
    cudaStream_t s;
    int device = 0;
    if (int rc = kba_internal_stream(h, &s, &device)) return rc;
    if (cudaSetDevice(device) != cudaSuccess) return kba_internal_fail(KBA_ERR_CUDA, "cudaSetDevice failed");
    kba_shard_comm* c = new kba_shard_comm;
    c->rank = rank; c->world = world; c->device = device;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    const ncclResult_t r = api().comm_init_rank(&c->comm, world, uid, rank);
    if (r != ncclSuccess) { delete c; return nccl_fail("ncclCommInitRank", r); }
    *out = c;
    return KBA_OK;
}

void kba_shard_comm_destroy(kba_shard_comm* c) {
    if (!c) return;
    if (c->comm && api().ok()) api().comm_destroy(c->comm);
    delete c;
}

}  // extern "C"
