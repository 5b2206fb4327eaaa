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
#include <stdlib.h>
#include <string.h>
#include <vector>

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
static void p2p_teardown(void);
extern "C" int scs_b200_comm_finalize(void) {
  p2p_teardown();
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

// ---------------------------------------------------------------------------------------------
// Peer-memory exchange buffers (NVLink P2P through CUDA IPC) for the fused "sum the partial
// A_g' z over the ranks + R_x p + p'Gp" kernel of the row-sharded CG (kernels/cg.cu). Each rank
// owns ONE allocation: red[2][n] doubles (double-buffered partials), rs[2][n] doubles (its slice of
// the reduced vector, two-phase mode), p[n] and inbox[G][ceil(n/G)] (sharded-x push mode, kernels/cg.cu) and a line
// of 64 flags (slots 0..7 "partial ready" per peer, 8..15 "slice ready" per peer, 16.. sharded-x) plus 64 scalar
// slots; every rank maps the allocation of every other rank. The handles travel in one
// NCCL all-gather at setup; after that the CG inner loop does not call NCCL at all.
#define P2P_MAX_RANKS 8
typedef struct {
  int ok, n;
  double *base[P2P_MAX_RANKS];               // base[r]: rank r's allocation as mapped in THIS process
  unsigned long long *flags[P2P_MAX_RANKS];  // flags[r] = (unsigned long long*)(base[r] + P2P_OFF_FLAGS(n))
} B200P2p;
static B200P2p g_p2p;
static ncclResult_t (*N_AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;

static unsigned long long g_p2p_seq = 0;
extern "C" int b200_p2p_ok(int n) { return g_p2p.ok && g_p2p.n >= n; }
extern "C" int b200_p2p_stride(void) { return g_p2p.n; }
extern "C" double *b200_p2p_base(int r) { return g_p2p.base[r]; }
extern "C" unsigned long long *b200_p2p_flags(int r) { return g_p2p.flags[r]; }
extern "C" unsigned long long b200_p2p_next_seq(void) { return ++g_p2p_seq; }

// offsets (in doubles) inside a rank's exchange allocation of stride n
#define P2P_OFF_P(n) ((size_t)4 * (n))          /* p vector of the sharded-x mode */
#define P2P_OFF_INBOX(n) ((size_t)5 * (n))      /* inbox [G][ceil(n/G)] of 16-byte self-validating elements */
#define P2P_OFF_PBOX(n) ((size_t)7 * (n) + 32)  /* the peers' slices of p as 16-byte self-validating elements, [n][2] */
#define P2P_OFF_FLAGS(n) ((size_t)9 * (n) + 64) /* 64 flags, then 64 scalar-message words */
extern "C" double *b200_p2p_pvec(int r) { return g_p2p.base[r] + P2P_OFF_P(g_p2p.n); }
extern "C" double *b200_p2p_inbox(int r) { return g_p2p.base[r] + P2P_OFF_INBOX(g_p2p.n); }
extern "C" double *b200_p2p_pbox(int r) { return g_p2p.base[r] + P2P_OFF_PBOX(g_p2p.n); }

// the exchange p vector can serve ONE workspace at a time (sharded-x mode): claim / release
static int g_p2p_p_claimed = 0;
extern "C" int b200_p2p_claim_pvec(void) {
  if (!g_p2p.ok || g_p2p_p_claimed) return -1;
  g_p2p_p_claimed = 1;
  return 0;
}
extern "C" void b200_p2p_release_pvec(void) { g_p2p_p_claimed = 0; }

extern "C" int b200_p2p_setup(int n_req) {
  if (g_nranks <= 1 || g_nranks > P2P_MAX_RANKS) return -1;
  if (g_p2p.ok && g_p2p.n >= n_req) return 0;
  if (g_p2p.ok) return -1;  // one allocation per process, sized generously below; larger systems use NCCL
  // stride of the allocation: at least 2^23 doubles, so that the workspaces of different sizes a process creates one
  // after the other (tests, bench.py) share it; 9 x 64 MB = 600 MB of the 180 GB
  const int n = n_req > (1 << 23) ? n_req : (1 << 23);
  // Every rank reaches the two collectives below whatever happens locally (an early return on one
  // rank -- SCS_B200_P2P=0 set for it alone, a failed allocation -- would leave the others blocked in NCCL):
  // local failures only clear `can_try`, and the final agreement makes all ranks fall back together.
  int can_try = 1;
  const char *e = getenv("SCS_B200_P2P");
  if (e && atoi(e) == 0) can_try = 0;
  if (!N_AllGather) *(void **)(&N_AllGather) = dlsym(N.h, "ncclAllGather");
  if (!N_AllGather) return -1;  // same library on every rank: a uniform outcome
  cudaStream_t st = (cudaStream_t)b200_stream();
  const size_t bytes = (P2P_OFF_FLAGS(n) + 64 + 64) * 8;
  // straight from the driver, not from the caching allocator: the block is exported to the peers through CUDA IPC and
  // must not be recycled for anything else while they may have it mapped
  double *mine = nullptr;
  if (can_try && cudaMalloc((void **)&mine, bytes) != cudaSuccess) { cudaGetLastError(); mine = nullptr; }
  cudaIpcMemHandle_t h;
  memset(&h, 0, sizeof(h));
  if (!mine || b200_memset0(mine, bytes) != 0) can_try = 0;
  if (can_try && cudaIpcGetMemHandle(&h, mine) != cudaSuccess) { cudaGetLastError(); can_try = 0; }
  // all-gather the 64-byte handles through NCCL (device buffers); a rank that cannot try sends zeros
  char *d_all = (char *)b200_malloc((size_t)g_nranks * sizeof(h));
  std::vector<cudaIpcMemHandle_t> all(g_nranks);
  int rc = -1;
  std::vector<void *> opened;
  if (d_all && cudaMemcpyAsync(d_all + (size_t)g_rank * sizeof(h), &h, sizeof(h), cudaMemcpyHostToDevice, st) == cudaSuccess &&
      N_AllGather(d_all + (size_t)g_rank * sizeof(h), d_all, sizeof(h), /*ncclChar*/ 0, g_comm, st) == 0 &&
      cudaMemcpyAsync(all.data(), d_all, (size_t)g_nranks * sizeof(h), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
      cudaStreamSynchronize(st) == cudaSuccess && can_try) {
    rc = 0;
    for (int r = 0; r < g_nranks && rc == 0; ++r) {
      void *p = mine;
      if (r != g_rank) {
        if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          rc = -1;
          break;
        }
        opened.push_back(p);
      }
      g_p2p.base[r] = (double *)p;
      g_p2p.flags[r] = (unsigned long long *)((double *)p + P2P_OFF_FLAGS(n));
    }
  }
  b200_free(d_all);
  // every rank must agree that the mapping worked (otherwise all fall back to NCCL)
  {
    double *d_flag = (double *)b200_malloc(8);
    double hv = (rc == 0) ? 0.0 : 1.0, out = 1.0;
    if (d_flag && cudaMemcpyAsync(d_flag, &hv, 8, cudaMemcpyHostToDevice, st) == cudaSuccess &&
        b200_allreduce_sum(d_flag, 1) == 0 &&
        cudaMemcpyAsync(&out, d_flag, 8, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
        cudaStreamSynchronize(st) == cudaSuccess && out == 0.0) {
      g_p2p.ok = 1;
      g_p2p.n = n;
    }
    b200_free(d_flag);
  }
  if (!g_p2p.ok) {
    for (void *p : opened) cudaIpcCloseMemHandle(p);
    if (mine) cudaFree(mine);
    memset(&g_p2p, 0, sizeof(g_p2p));
    fprintf(stderr, "scs_b200: peer-memory (CUDA IPC) mapping unavailable, using NCCL all-reduce\n");
  }
  return g_p2p.ok ? 0 : -1;
}

// close the peer mappings and free this rank's exchange allocation (scs_b200_comm_finalize)
static void p2p_teardown(void) {
  if (!g_p2p.ok) return;
  b200_sync();
  for (int r = 0; r < g_nranks && r < P2P_MAX_RANKS; ++r) {
    if (!g_p2p.base[r]) continue;
    if (r == g_rank) cudaFree(g_p2p.base[r]);
    else cudaIpcCloseMemHandle(g_p2p.base[r]);
  }
  memset(&g_p2p, 0, sizeof(g_p2p));
}
