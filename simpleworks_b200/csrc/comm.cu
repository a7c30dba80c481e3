// Multi-GPU exchange inside the library: one NCCL communicator per context (one process per GPU) and the only
// collective the Marlin hot path needs -- an all-gather of the ranks' 144-byte partial MSM results on the
// context's stream, summed on the host by every rank ("per-GPU partial sums combined through a tiny
// NCCL all-gather over NVLink").  A Rust host linking libswb200.a gets it through swb_comm_*; nothing here goes
// through Python.  NCCL is resolved at run time (dlopen of libnccl.so.2: in a torch process that is the copy
// torch already loaded), so the library itself has no link-time dependency on it and loads on boxes without it.
#include <dlfcn.h>

#include <vector>

#include "ctx.hpp"

namespace swb {

struct NcclUniqueId { char internal[128]; };      // == ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* NcclComm;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int /*ncclDataType_t*/, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok() const { return GetUniqueId && CommInitRank && AllGather && CommDestroy; }
};
constexpr size_t COMM_STAGING_BYTES = 1 << 20;     // [mine | all ranks'] x 144 bytes x batch

static NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (a.lib) {
            a.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(a.lib, "ncclGetUniqueId");
            a.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(a.lib, "ncclCommInitRank");
            a.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(a.lib, "ncclAllGather");
            a.CommDestroy = (int (*)(NcclComm))dlsym(a.lib, "ncclCommDestroy");
            a.GetErrorString = (const char* (*)(int))dlsym(a.lib, "ncclGetErrorString");
        }
        return a;
    }();
    return api;
}
static int nccl_fail(swb_ctx* c, int rc, const char* what) {
    const char* txt = nccl().GetErrorString ? nccl().GetErrorString(rc) : "?";
    return set_err(c, SWB_EINTERNAL, "NCCL: %s failed: %s", what, txt);
}

int comm_sum_g1(swb_ctx* c, const swb_g1_jacobian* mine, size_t count, swb_g1_jacobian* out) {
    SWB_REQUIRE(c, c->comm != nullptr, "comm: swb_comm_init has not been called on this context");
    SWB_REQUIRE(c, count > 0 && count <= 256 && mine && out, "comm_sum_g1: bad arguments");
    SWB_CUDA(c, cudaSetDevice(c->device));
    const size_t bytes = count * sizeof(swb_g1_jacobian), world = (size_t)c->comm_world;
    // staging: [mine | everybody's] on the device, the same in pinned host memory
    if (!c->comm_dev) {
        SWB_CUDA(c, cudaMalloc(&c->comm_dev, COMM_STAGING_BYTES));
        SWB_CUDA(c, cudaMallocHost(&c->comm_host, COMM_STAGING_BYTES));
    }
    SWB_REQUIRE(c, (world + 1) * bytes <= COMM_STAGING_BYTES, "comm_sum_g1: batch too large for this world size");
    uint8_t* dev = static_cast<uint8_t*>(c->comm_dev);
    uint8_t* host = static_cast<uint8_t*>(c->comm_host);
    memcpy(host, mine, bytes);
    SWB_CUDA(c, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
    const int rc = nccl().AllGather(dev, dev + bytes, bytes, 0 /* ncclInt8 */, c->comm, c->stream);
    if (rc != 0) return nccl_fail(c, rc, "ncclAllGather");
    SWB_CUDA(c, cudaMemcpyAsync(host + bytes, dev + bytes, world * bytes, cudaMemcpyDeviceToHost, c->stream));
    SWB_CUDA(c, cudaStreamSynchronize(c->stream));
    const swb_g1_jacobian* all = reinterpret_cast<const swb_g1_jacobian*>(host + bytes);
    std::vector<swb_g1_jacobian> col(world);
    for (size_t k = 0; k < count; k++) {
        for (size_t r = 0; r < world; r++) col[r] = all[r * count + k];
        const int rs = swb_g1_sum_jacobian(c, col.data(), world, &out[k]);
        if (rs != SWB_OK) return rs;
    }
    return SWB_OK;
}

}  // namespace swb

using namespace swb;

extern "C" {

int swb_comm_unique_id(uint8_t id[128]) {
    if (!id || !nccl().ok()) return SWB_EINTERNAL;
    NcclUniqueId u;
    if (nccl().GetUniqueId(&u) != 0) return SWB_EINTERNAL;
    memcpy(id, u.internal, 128);
    return SWB_OK;
}

int swb_comm_init(swb_ctx* c, const uint8_t id[128], int rank, int world) {
    if (!c) return SWB_EARG;
    SWB_REQUIRE(c, id != nullptr && world >= 1 && rank >= 0 && rank < world, "comm_init: bad arguments");
    SWB_REQUIRE(c, c->comm == nullptr, "comm_init: this context already has a communicator");
    if (!nccl().ok()) return set_err(c, SWB_EINTERNAL, "%s", "comm_init: libnccl.so.2 could not be loaded");
    SWB_CUDA(c, cudaSetDevice(c->device));
    NcclUniqueId u;
    memcpy(u.internal, id, 128);
    NcclComm comm = nullptr;
    const int rc = nccl().CommInitRank(&comm, world, u, rank);
    if (rc != 0) return nccl_fail(c, rc, "ncclCommInitRank");
    c->comm = comm;
    c->comm_rank = rank;
    c->comm_world = world;
    return SWB_OK;
}

int swb_comm_destroy(swb_ctx* c) {
    if (!c) return SWB_EARG;
    if (c->comm) {
        cudaSetDevice(c->device);
        sync_all_streams(c);
        nccl().CommDestroy(c->comm);
        c->comm = nullptr;
    }
    if (c->comm_dev) { cudaFree(c->comm_dev); c->comm_dev = nullptr; }
    if (c->comm_host) { cudaFreeHost(c->comm_host); c->comm_host = nullptr; }
    c->comm_rank = 0;
    c->comm_world = 1;
    return SWB_OK;
}

int swb_comm_info(const swb_ctx* c, int* rank, int* world) {
    if (!c) return SWB_EARG;
    if (rank) *rank = c->comm ? c->comm_rank : 0;
    if (world) *world = c->comm ? c->comm_world : 1;
    return SWB_OK;
}

int swb_comm_sum_g1(swb_ctx* c, const swb_g1_jacobian* mine, size_t count, swb_g1_jacobian* out) {
    if (!c) return SWB_EARG;
    return comm_sum_g1(c, mine, count, out);
}

}  // extern "C"
