// comm.cu -- NCCL plumbing for the row-sharded KKT solve (one process per GPU).
//
// The reference has no distributed path at all ("no NCCL, MPI, Gloo");
// this is new work for SURVEY section 8(e).  NCCL is loaded with dlopen at run time
// (libnccl.so.2 -- inside a torch process that is torch's bundled NCCL), so the
// single-GPU library has no link-time dependency on it.  The ncclUniqueId is created
// by rank 0 (scs_b200_comm_unique_id) and distributed by the launcher (bench.py /
// the tests use torch.distributed for that); every collective is enqueued on the
// library stream, so it is ordered with the kernels like any other launch.
#include "../common.cuh"
#include "../dev_api.h"
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSum_ = 0, ncclMax_ = 2 };
enum { ncclFloat64_ = 8 };  // ncclDouble

static struct {
  void *h;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)(void);
  ncclResult_t (*GroupEnd)(void);
  const char *(*GetErrorString)(ncclResult_t);
} N;
static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_nranks = 1;

static int load_nccl(void) {
  if (N.h) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !N.h; ++i) N.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!N.h) {
    fprintf(stderr, "scs_b200: cannot dlopen libnccl.so.2: %s\n", dlerror());
    return -1;
  }
#define SYM(field, name)                                           \
  *(void **)(&N.field) = dlsym(N.h, name);                         \
  if (!N.field) { fprintf(stderr, "scs_b200: NCCL symbol %s missing\n", name); return -1; }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(Broadcast, "ncclBroadcast")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return 0;
}
#define NCCL_OK(call)                                                                   \
  do {                                                                                  \
    ncclResult_t r__ = (call);                                                          \
    if (r__ != 0) {                                                                     \
      fprintf(stderr, "scs_b200: %s failed: %s\n", #call, N.GetErrorString(r__));       \
      return -1;                                                                        \
    }                                                                                   \
  } while (0)

extern "C" int scs_b200_comm_unique_id(char *out128) {
  if (load_nccl() != 0) return -1;
  ncclUniqueId id;
  NCCL_OK(N.GetUniqueId(&id));
  memcpy(out128, id.internal, 128);
  return 0;
}
extern "C" int scs_b200_comm_init(int rank, int nranks, const char *id128) {
  if (nranks <= 1) { g_rank = 0; g_nranks = 1; return 0; }
  if (b200_runtime_init() != 0) return -1;
  if (load_nccl() != 0) return -1;
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  NCCL_OK(N.CommInitRank(&g_comm, nranks, id, rank));
  g_rank = rank;
  g_nranks = nranks;
  return 0;
}
extern "C" int scs_b200_comm_finalize(void) {
  if (g_comm) {
    b200_sync();
    N.CommDestroy(g_comm);
    g_comm = nullptr;
  }
  g_rank = 0;
  g_nranks = 1;
  return 0;
}
extern "C" int b200_comm_rank(void) { return g_rank; }
extern "C" int b200_comm_nranks(void) { return g_nranks; }

extern "C" int b200_allreduce_sum(double *d_buf, size_t count) {
  if (g_nranks <= 1) return 0;
  NCCL_OK(N.AllReduce(d_buf, d_buf, count, ncclFloat64_, ncclSum_, g_comm, (cudaStream_t)b200_stream()));
  return 0;
}
// in-place all-gather of contiguous, possibly unequal blocks: rank r owns d_buf[offsets[r] .. offsets[r+1])
extern "C" int b200_allgatherv(double *d_buf, const int *offsets) {
  if (g_nranks <= 1) return 0;
  NCCL_OK(N.GroupStart());
  for (int r = 0; r < g_nranks; ++r) {
    const size_t cnt = (size_t)(offsets[r + 1] - offsets[r]);
    if (cnt == 0) continue;
    NCCL_OK(N.Broadcast(d_buf + offsets[r], d_buf + offsets[r], cnt, ncclFloat64_, r, g_comm,
                        (cudaStream_t)b200_stream()));
  }
  NCCL_OK(N.GroupEnd());
  return 0;
}
