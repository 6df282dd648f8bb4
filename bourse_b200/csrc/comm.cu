// Multi-GPU side of the C ABI (include/bourse_b200.h, "multi-GPU"): books are independent, so a job shards its envs over
// GPUs with no per-step collective; the ONE exchange is the end-of-run all-gather of each shard's statistics block
// (SURVEY.md 8e), done here with ncclAllGather over NVLink / NVSwitch.  Torch-free: a Rust / C host gets the whole
// multi-GPU path from this library.  NCCL is resolved at run time (dlopen "libnccl.so.2": the system library, or the one a
// host process has already loaded), so a single-GPU user of the library needs no NCCL at all.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/bourse_b200.h"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi* nccl() {
    static NcclApi api;
    if (api.lib || !api.err.empty()) return &api;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) {
        api.err = std::string("NCCL is not available (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
        return &api;
    }
#define RESOLVE(field, sym)                                                   \
    *(void**)(&api.field) = dlsym(api.lib, sym);                              \
    if (!api.field && api.err.empty()) api.err = std::string("NCCL symbol missing: ") + sym;
    RESOLVE(GetUniqueId, "ncclGetUniqueId") RESOLVE(CommInitRank, "ncclCommInitRank") RESOLVE(CommInitAll, "ncclCommInitAll")
    RESOLVE(CommDestroy, "ncclCommDestroy") RESOLVE(AllGather, "ncclAllGather") RESOLVE(GroupStart, "ncclGroupStart")
    RESOLVE(GroupEnd, "ncclGroupEnd") RESOLVE(GetErrorString, "ncclGetErrorString")
#undef RESOLVE
    return &api;
}

thread_local std::string g_comm_error;

int cfail(int code, const std::string& msg) {
    g_comm_error = msg;
    return code;
}

constexpr int BLOCK_WORDS = 10;  // 8 x bb_stats_t fields, elapsed time (double bits), reserved

}  // namespace

struct bb_comm {
    int n_ranks = 0;
    struct Local {
        int rank = 0, device = 0;
        ncclComm_t comm = nullptr;
        cudaStream_t stream = nullptr;
        unsigned long long *d_send = nullptr, *d_recv = nullptr;
    };
    std::vector<Local> local;  // the ranks this process drives (one per process under torchrun, all of them with init_all)
};

namespace {

int alloc_local(bb_comm* c, bb_comm::Local& l) {
    if (cudaSetDevice(l.device) != cudaSuccess) return cfail(BB_ECUDA, "cudaSetDevice failed");
    if (cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) != cudaSuccess) return cfail(BB_ECUDA, "stream creation failed");
    if (cudaMalloc(&l.d_send, BLOCK_WORDS * 8) != cudaSuccess || cudaMalloc(&l.d_recv, (size_t)c->n_ranks * BLOCK_WORDS * 8) != cudaSuccess)
        return cfail(BB_ECUDA, "cudaMalloc failed");
    return BB_OK;
}

}  // namespace

extern "C" {

const char* bb_comm_last_error(void) { return g_comm_error.c_str(); }

int bb_comm_unique_id(unsigned char* out /* BB_COMM_ID_BYTES */) {
    NcclApi* n = nccl();
    if (!n->err.empty()) return cfail(BB_ECUDA, n->err);
    if (!out) return cfail(BB_EINVAL, "null argument");
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == BB_COMM_ID_BYTES, "ncclUniqueId size");
    const ncclResult_t r = n->GetUniqueId(&id);
    if (r != ncclSuccess) return cfail(BB_ECUDA, std::string("ncclGetUniqueId: ") + n->GetErrorString(r));
    memcpy(out, &id, sizeof id);
    return BB_OK;
}

int bb_comm_init_rank(const unsigned char* id_bytes, int n_ranks, int rank, int device, bb_comm** out) {
    NcclApi* n = nccl();
    if (!n->err.empty()) return cfail(BB_ECUDA, n->err);
    if (!id_bytes || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return cfail(BB_EINVAL, "bad argument");
    bb_comm* c = new bb_comm();
    c->n_ranks = n_ranks;
    c->local.resize(1);
    c->local[0].rank = rank;
    c->local[0].device = device;
    int rc = alloc_local(c, c->local[0]);
    if (rc) { bb_comm_destroy(c); return rc; }
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    const ncclResult_t r = n->CommInitRank(&c->local[0].comm, n_ranks, id, rank);
    if (r != ncclSuccess) {
        bb_comm_destroy(c);
        return cfail(BB_ECUDA, std::string("ncclCommInitRank: ") + n->GetErrorString(r));
    }
    *out = c;
    return BB_OK;
}

int bb_comm_init_all(int n_dev, const int* devices, bb_comm** out) {
    NcclApi* n = nccl();
    if (!n->err.empty()) return cfail(BB_ECUDA, n->err);
    if (!out || n_dev < 1) return cfail(BB_EINVAL, "bad argument");
    bb_comm* c = new bb_comm();
    c->n_ranks = n_dev;
    c->local.resize(n_dev);
    std::vector<int> devs(n_dev);
    for (int i = 0; i < n_dev; ++i) {
        devs[i] = devices ? devices[i] : i;
        c->local[i].rank = i;
        c->local[i].device = devs[i];
        int rc = alloc_local(c, c->local[i]);
        if (rc) { bb_comm_destroy(c); return rc; }
    }
    std::vector<ncclComm_t> comms(n_dev);
    const ncclResult_t r = n->CommInitAll(comms.data(), n_dev, devs.data());
    if (r != ncclSuccess) {
        bb_comm_destroy(c);
        return cfail(BB_ECUDA, std::string("ncclCommInitAll: ") + n->GetErrorString(r));
    }
    for (int i = 0; i < n_dev; ++i) c->local[i].comm = comms[i];
    *out = c;
    return BB_OK;
}

int bb_comm_n_ranks(const bb_comm* c) { return c ? c->n_ranks : 0; }

int bb_comm_destroy(bb_comm* c) {
    if (!c) return BB_OK;
    NcclApi* n = nccl();
    for (auto& l : c->local) {
        cudaSetDevice(l.device);
        if (l.comm && n->CommDestroy) n->CommDestroy(l.comm);
        cudaFree(l.d_send);
        cudaFree(l.d_recv);
        if (l.stream) cudaStreamDestroy(l.stream);
    }
    delete c;
    return BB_OK;
}

int bb_gather_stats(bb_comm* c, bb_handle* const* handles, uint32_t n_local, const double* elapsed_ms, bb_stats_t* out_stats,
                    double* out_elapsed_ms) {
    NcclApi* n = nccl();
    if (!n->err.empty()) return cfail(BB_ECUDA, n->err);
    if (!c || !handles || !out_stats) return cfail(BB_EINVAL, "null argument");
    if (n_local != c->local.size()) return cfail(BB_EINVAL, "one handle per local rank of the communicator");
    // each shard's statistics block (bb_stats: counters reduced on its own device), staged in device memory
    for (uint32_t i = 0; i < n_local; ++i) {
        bb_stats_t st;
        const int rc = bb_stats(handles[i], &st);
        if (rc) return cfail(rc, std::string("bb_stats: ") + bb_last_error(handles[i]));
        unsigned long long block[BLOCK_WORDS] = {st.instructions, st.orders_created, st.trades, st.traded_volume, st.env_steps,
                                                 st.transitions, st.error_envs, st.l1_checksum, 0, 0};
        const double ms = elapsed_ms ? elapsed_ms[i] : 0.0;
        memcpy(&block[8], &ms, 8);
        auto& l = c->local[i];
        if (cudaSetDevice(l.device) != cudaSuccess ||
            cudaMemcpyAsync(l.d_send, block, sizeof block, cudaMemcpyHostToDevice, l.stream) != cudaSuccess ||
            cudaStreamSynchronize(l.stream) != cudaSuccess)
            return cfail(BB_ECUDA, "staging the statistics block failed");
    }
    // the run's only collective
    ncclResult_t r = n->GroupStart();
    for (uint32_t i = 0; i < n_local && r == ncclSuccess; ++i) {
        auto& l = c->local[i];
        cudaSetDevice(l.device);
        r = n->AllGather(l.d_send, l.d_recv, BLOCK_WORDS, ncclUint64, l.comm, l.stream);
    }
    const ncclResult_t r2 = n->GroupEnd();
    if (r != ncclSuccess || r2 != ncclSuccess) return cfail(BB_ECUDA, std::string("ncclAllGather: ") + n->GetErrorString(r != ncclSuccess ? r : r2));
    for (auto& l : c->local) {
        cudaSetDevice(l.device);
        if (cudaStreamSynchronize(l.stream) != cudaSuccess) return cfail(BB_ECUDA, "all-gather failed");
    }
    std::vector<unsigned long long> all((size_t)c->n_ranks * BLOCK_WORDS);
    auto& l0 = c->local[0];
    cudaSetDevice(l0.device);
    if (cudaMemcpy(all.data(), l0.d_recv, all.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return cfail(BB_ECUDA, "read-back failed");
    for (int k = 0; k < c->n_ranks; ++k) {
        const unsigned long long* b = &all[(size_t)k * BLOCK_WORDS];
        bb_stats_t& s = out_stats[k];
        s.instructions = b[0]; s.orders_created = b[1]; s.trades = b[2]; s.traded_volume = b[3]; s.env_steps = b[4];
        s.transitions = b[5]; s.error_envs = b[6]; s.l1_checksum = b[7];
        if (out_elapsed_ms) memcpy(&out_elapsed_ms[k], &b[8], 8);
    }
    return BB_OK;
}

}  // extern "C"
