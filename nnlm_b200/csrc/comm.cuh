// comm.cuh — thin RAII wrapper over the NCCL communicator of one rank (one process per GPU).
#pragma once
#include "common.cuh"

namespace nnlm {

struct NcclId { char bytes[NNLM_COMM_ID_BYTES]; };   // ncclUniqueId is 128 opaque bytes

class Comm {
public:
    static void unique_id(NcclId* id);
    Comm(const NcclId& id, int rank, int nranks, int device);
    ~Comm();
    Comm(const Comm&) = delete;
    Comm& operator=(const Comm&) = delete;
    int rank() const { return rank_; }
    int nranks() const { return nranks_; }
    void allreduce_sum_f64(double* buf, size_t count, cudaStream_t st);
    void allreduce_sum_u64(unsigned long long* buf, size_t count, cudaStream_t st);
    void allreduce_max_u64(unsigned long long* buf, size_t count, cudaStream_t st);
    void allgather_f64(const double* send, double* recv, size_t count_per_rank, cudaStream_t st);
    void broadcast_f64(double* buf, size_t count, int root, cudaStream_t st);
private:
    void* comm_ = nullptr;
    int rank_, nranks_, device_;
};

}  // namespace nnlm

struct nnlm_comm;
nnlm::Comm* nnlm_comm_get(nnlm_comm* c);
