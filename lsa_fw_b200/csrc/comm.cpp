// See comm.h.  Only the handful of NCCL entry points the partitioned solve needs, resolved with dlsym.
#include "comm.h"

#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>

namespace lsa {
namespace {

// mirror of the public NCCL types used here (nccl.h 2.x: stable ABI)
struct UniqueId {
  char internal[128];
};
typedef int Result;
enum { kSum = 0, kDouble = 8 };

struct Api {
  void* lib = nullptr;
  Result (*GetUniqueId)(UniqueId*) = nullptr;
  Result (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
  Result (*CommDestroy)(void*) = nullptr;
  Result (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  Result (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  Result (*GroupStart)() = nullptr;
  Result (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(Result) = nullptr;
};
Api g_api;
std::mutex g_mu;

template <class F>
void sym(F& f, const char* name) {
  f = reinterpret_cast<F>(dlsym(g_api.lib, name));
  if (!f) throw std::runtime_error(std::string("NCCL symbol missing: ") + name);
}

Api& api(const char* path = nullptr) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_api.lib) return g_api;
  const char* cands[] = {path, "libnccl.so.2", "libnccl.so"};
  for (const char* c : cands) {
    if (!c) continue;
    g_api.lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (g_api.lib) break;
  }
  if (!g_api.lib)
    throw std::runtime_error("libnccl.so.2 not found (import torch first, or pass its path to lsa_nccl_load): the "
                             "partitioned solve needs NCCL");
  sym(g_api.GetUniqueId, "ncclGetUniqueId");
  sym(g_api.CommInitRank, "ncclCommInitRank");
  sym(g_api.CommDestroy, "ncclCommDestroy");
  sym(g_api.AllReduce, "ncclAllReduce");
  sym(g_api.Broadcast, "ncclBroadcast");
  sym(g_api.GroupStart, "ncclGroupStart");
  sym(g_api.GroupEnd, "ncclGroupEnd");
  sym(g_api.GetErrorString, "ncclGetErrorString");
  return g_api;
}

void check(Result r, const char* what) {
  if (r != 0) throw std::runtime_error(std::string(what) + ": " + (g_api.GetErrorString ? g_api.GetErrorString(r) : "NCCL error"));
}

}  // namespace

void nccl_load(const char* path) { api(path); }

void nccl_unique_id(void* out128) {
  UniqueId id;
  check(api().GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out128, id.internal, 128);
}

void comm_init(Comm& c, const void* id128, int rank, int world) {
  UniqueId id;
  std::memcpy(id.internal, id128, 128);
  check(api().CommInitRank(&c.nccl_comm, world, id, rank), "ncclCommInitRank");
  c.rank = rank;
  c.world = world;
}

void comm_destroy(Comm& c) {
  if (c.nccl_comm) api().CommDestroy(c.nccl_comm);
  c.nccl_comm = nullptr;
}

void comm_allreduce_sum(const Comm& c, double* buf, size_t count, cudaStream_t st) {
  if (c.world <= 1 || count == 0) return;
  if (!c.nccl_comm) throw std::runtime_error("partitioned handle without a communicator (lsa_set_comm)");
  check(api().AllReduce(buf, buf, count, kDouble, kSum, c.nccl_comm, st), "ncclAllReduce");
}

void comm_group_begin() { check(api().GroupStart(), "ncclGroupStart"); }
void comm_bcast(const Comm& c, double* buf, size_t count, int root, cudaStream_t st) {
  if (!c.nccl_comm) throw std::runtime_error("partitioned handle without a communicator (lsa_set_comm)");
  check(api().Broadcast(buf, buf, count, kDouble, root, c.nccl_comm, st), "ncclBroadcast");
}
void comm_group_end() { check(api().GroupEnd(), "ncclGroupEnd"); }

}  // namespace lsa
