// mrhbm_comm.cu -- dlopen-bound NCCL: unique id, communicator, all-gather, all-to-all-v.
// The exchange replaces the reference's store-and-forward through GridFS / scp
// (mapreduce/job.lua:217-221,255-260; mapreduce/fs.lua:143-160).
#include "mrhbm_comm.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstring>

#include "../../include/mrhbm.h"

namespace mrhbm {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclUint8 = 1, ncclUint32 = 3 };

struct Nccl {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static Nccl g_nccl;

static int nccl_load(std::string* err) {
  if (g_nccl.h) return 0;
  // prefer a copy that is already mapped into the process (torch's bundled NCCL)
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    const char* env = getenv("MRHBM_NCCL_LIB");
    if (env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  }
  for (const char* n : names) {
    if (h) break;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!h) {
    if (err) *err = std::string("cannot load NCCL: ") + dlerror();
    return MRHBM_E_NCCL;
  }
#define SYM(field, name)                                            \
  *(void**)(&g_nccl.field) = dlsym(h, name);                        \
  if (!g_nccl.field) {                                              \
    if (err) *err = std::string("NCCL symbol missing: ") + name;    \
    return MRHBM_E_NCCL;                                            \
  }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllGather, "ncclAllGather")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.h = h;
  return 0;
}

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

#define NC(expr)                                                                         \
  do {                                                                                   \
    int r_ = (expr);                                                                     \
    if (r_ != ncclSuccess) {                                                             \
      if (err) *err = std::string("NCCL: ") + g_nccl.GetErrorString(r_) + " (" #expr ")"; \
      return MRHBM_E_NCCL;                                                               \
    }                                                                                    \
  } while (0)

int comm_unique_id(void* id128, std::string* err) {
  int rc = nccl_load(err);
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == MRHBM_UNIQUE_ID_BYTES, "unique id size");
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return 0;
}

int comm_create(Comm** out, const void* id128, int rank, int world, int dev, std::string* err) {
  int rc = nccl_load(err);
  if (rc) return rc;
  if (cudaSetDevice(dev) != cudaSuccess) {
    if (err) *err = "cudaSetDevice failed";
    return MRHBM_E_CUDA;
  }
  Comm* c = new Comm();
  c->rank = rank;
  c->world = world;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  int r = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    if (err) *err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r);
    delete c;
    return MRHBM_E_NCCL;
  }
  *out = c;
  return 0;
}

void comm_destroy(Comm* c) {
  if (!c) return;
  if (c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
}
int comm_rank(const Comm* c) { return c ? c->rank : 0; }
int comm_world(const Comm* c) { return c ? c->world : 1; }

int comm_allgather_u32(Comm* c, const uint32_t* send, uint32_t* recv, size_t count, cudaStream_t s, std::string* err) {
  NC(g_nccl.AllGather(send, recv, count, ncclUint32, c->comm, s));
  return 0;
}

int comm_alltoallv(Comm* c, const void* send, const uint64_t* send_off, const uint64_t* send_cnt, void* recv,
                   const uint64_t* recv_off, const uint64_t* recv_cnt, cudaStream_t s, std::string* err) {
  NC(g_nccl.GroupStart());
  for (int p = 0; p < c->world; p++) {
    if (send_cnt[p]) NC(g_nccl.Send((const char*)send + send_off[p], send_cnt[p], ncclUint8, p, c->comm, s));
    if (recv_cnt[p]) NC(g_nccl.Recv((char*)recv + recv_off[p], recv_cnt[p], ncclUint8, p, c->comm, s));
  }
  NC(g_nccl.GroupEnd());
  return 0;
}

}  // namespace mrhbm
