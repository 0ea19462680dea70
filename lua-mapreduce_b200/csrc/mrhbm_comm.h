// mrhbm_comm.h -- NCCL plumbing for the partition exchange (internal, C++).
// NCCL is bound at run time with dlopen so that the library shares the NCCL already loaded
// into the process (torch's) instead of linking a second copy.
#pragma once
#include <cstdint>
#include <string>
#include <cuda_runtime.h>

namespace mrhbm {
struct Comm;
int comm_unique_id(void* id128, std::string* err);
int comm_create(Comm** out, const void* id128, int rank, int world, int dev, std::string* err);
void comm_destroy(Comm*);
int comm_rank(const Comm*);
int comm_world(const Comm*);
// all ranks: gather `count` u32 from every rank into recv[world*count]
int comm_allgather_u32(Comm*, const uint32_t* send, uint32_t* recv, size_t count, cudaStream_t s, std::string* err);
// all-to-all-v in bytes: send_off/recv_off/… are per-peer byte offsets and counts (host arrays)
int comm_alltoallv(Comm*, const void* send, const uint64_t* send_off, const uint64_t* send_cnt, void* recv,
                   const uint64_t* recv_off, const uint64_t* recv_cnt, cudaStream_t s, std::string* err);
}  // namespace mrhbm
