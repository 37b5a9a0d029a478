// Multi-GPU plumbing for the slab-partitioned solvers: one process per GPU, NCCL for the halo planes and the
// two scalar reductions of a CG iteration.  NCCL is resolved at run time with dlopen("libnccl.so.2") so that the
// library shares the copy torch.distributed already loaded and has no link-time dependency on it.
#pragma once
#include "fs_common.cuh"

struct fs_comm {
    void* nccl_comm = nullptr;
    int rank = 0;
    int nranks = 1;
};

namespace fs {

enum CommType { COMM_F32 = 0, COMM_F64 = 1, COMM_U8 = 2 };

int comm_allreduce_sum_f64(fs_comm* c, double* dev_inout, int count, cudaStream_t s);
// exchange `count`-element planes with the low (rank-1) and high (rank+1) neighbours in one NCCL group:
// for k < nplanes: send send_lo[k] -> rank-1, recv recv_lo[k] <- rank-1 (if has_lo); same with *_hi and rank+1.
int comm_halo_exchange(fs_comm* c, int has_lo, int has_hi, int nplanes, const void* const* send_lo, void* const* recv_lo,
                       const void* const* send_hi, void* const* recv_hi, size_t count, int type, cudaStream_t s);

}  // namespace fs
