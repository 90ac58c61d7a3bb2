// comm.cu — multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
// NCCL is dlopen()ed on first use so the single-GPU library has no link-time dependency on it.
#include <dlfcn.h>

#include "comm.cuh"

namespace nnlm {

namespace {

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

// resolved exactly once, also under concurrent first use (the worker threads of a multi-GPU nnlm_nnmf call): a
// function-local static is initialised under the C++11 guard, and a failed load throws out of the initialiser so the
// next caller retries
NcclApi load_api()
{
    NcclApi a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) throw Error(NNLM_E_NCCL, std::string("cannot load libnccl: ") + dlerror());
    auto sym = [&](const char* s) {
        void* p = dlsym(a.handle, s);
        if (!p) throw Error(NNLM_E_NCCL, std::string("libnccl lacks symbol ") + s);
        return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    return a;
}

NcclApi& api()
{
    static NcclApi a = load_api();
    return a;
}

void check(int rc, const char* what)
{
    if (rc != 0) throw Error(NNLM_E_NCCL, std::string(what) + ": " + api().GetErrorString(rc));
}

constexpr int kNcclFloat64 = 8;   // ncclDouble
constexpr int kNcclUint64 = 5;    // ncclUint64
constexpr int kNcclSum = 0;       // ncclSum
constexpr int kNcclMax = 2;       // ncclMax

}  // namespace

void Comm::unique_id(NcclId* id) { check(api().GetUniqueId(id), "ncclGetUniqueId"); }

Comm::Comm(const NcclId& id, int rank, int nranks, int device) : rank_(rank), nranks_(nranks), device_(device)
{
    NNLM_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / world size");
    if (device_ >= 0) NNLM_CUDA_CHECK(cudaSetDevice(device_));
    check(api().CommInitRank(&comm_, nranks, id, rank), "ncclCommInitRank");
}

Comm::~Comm() { if (comm_) api().CommDestroy(comm_); }

void Comm::allreduce_sum_f64(double* buf, size_t count, cudaStream_t st)
{
    check(api().AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, comm_, st), "ncclAllReduce");
}

void Comm::allreduce_sum_u64(unsigned long long* buf, size_t count, cudaStream_t st)
{
    check(api().AllReduce(buf, buf, count, kNcclUint64, kNcclSum, comm_, st), "ncclAllReduce");
}

void Comm::allreduce_max_u64(unsigned long long* buf, size_t count, cudaStream_t st)
{
    check(api().AllReduce(buf, buf, count, kNcclUint64, kNcclMax, comm_, st), "ncclAllReduce");
}

void Comm::allgather_f64(const double* send, double* recv, size_t count_per_rank, cudaStream_t st)
{
    check(api().AllGather(send, recv, count_per_rank, kNcclFloat64, comm_, st), "ncclAllGather");
}

void Comm::broadcast_f64(double* buf, size_t count, int root, cudaStream_t st)
{
    check(api().Broadcast(buf, buf, count, kNcclFloat64, root, comm_, st), "ncclBroadcast");
}

}  // namespace nnlm

using namespace nnlm;

struct nnlm_comm { Comm* c; };

extern "C" {
#pragma GCC visibility push(default)

int nnlm_comm_unique_id(unsigned char id[NNLM_COMM_ID_BYTES], char* err, size_t errlen)
{
    try {
        NcclId nid;
        Comm::unique_id(&nid);
        std::memcpy(id, nid.bytes, NNLM_COMM_ID_BYTES);
        return NNLM_OK;
    } catch (const Error& e) {
        if (err && errlen) std::snprintf(err, errlen, "%s", e.what());
        return e.code;
    }
}

int nnlm_comm_init(nnlm_comm** out, const unsigned char id[NNLM_COMM_ID_BYTES], int32_t rank, int32_t nranks,
                   int32_t device, char* err, size_t errlen)
{
    try {
        NcclId nid;
        std::memcpy(nid.bytes, id, NNLM_COMM_ID_BYTES);
        nnlm_comm* c = new nnlm_comm{new Comm(nid, rank, nranks, device)};
        *out = c;
        return NNLM_OK;
    } catch (const Error& e) {
        if (err && errlen) std::snprintf(err, errlen, "%s", e.what());
        return e.code;
    }
}

void nnlm_comm_destroy(nnlm_comm* c)
{
    if (!c) return;
    delete c->c;
    delete c;
}

#pragma GCC visibility pop
}  // extern "C"

nnlm::Comm* nnlm_comm_get(nnlm_comm* c) { return c ? c->c : nullptr; }
