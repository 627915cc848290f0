// NCCL access for the partitioned (multi-GPU) solve: the library is dlopen'ed at first use, so that single-GPU
// users need neither NCCL nor torch.  One process per GPU (torchrun); the communicator is created from a unique
// id that the Python layer distributes over torch.distributed.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace lsa {

struct Comm {
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
};

void nccl_load(const char* path);                     // optional explicit path of libnccl.so.2
void nccl_unique_id(void* out128);                    // 128 bytes
void comm_init(Comm& c, const void* id128, int rank, int world);
void comm_destroy(Comm& c);
// sum of `count` doubles over all ranks, in place
void comm_allreduce_sum(const Comm& c, double* buf, size_t count, cudaStream_t st);
// grouped broadcasts: begin, any number of bcast (each from its own root), end
void comm_group_begin();
void comm_bcast(const Comm& c, double* buf, size_t count, int root, cudaStream_t st);
void comm_group_end();

}  // namespace lsa
