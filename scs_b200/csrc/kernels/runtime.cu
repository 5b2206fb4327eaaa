// runtime.cu -- device selection, the library stream, memory and error plumbing.
// One process drives one GPU (LOCAL_RANK, overridable with SCS_B200_DEVICE).
#include "../common.cuh"
#include "../dev_api.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int g_init = 0;  // 0 = not tried, 1 = ok, -1 = failed
static int g_dev = 0;
static int g_num_sms = B200_NUM_SMS_FALLBACK;
static cudaStream_t g_stream = nullptr;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static char g_err[512] = "";
static long long g_launches = 0;
static void *g_stage = nullptr;  // pinned staging buffer for pageable host copies
static size_t g_stage_bytes = 0;

extern "C" void b200_set_error(const char *what, cudaError_t e, const char *file, int line) {
  snprintf(g_err, sizeof(g_err), "%s:%d: %s -> %s", file, line, what, cudaGetErrorString(e));
  if (getenv("SCS_B200_VERBOSE")) fprintf(stderr, "scs_b200: %s\n", g_err);
}
extern "C" void b200_count_launch(int n) { g_launches += n; }
extern "C" long long b200_launches(void) { return g_launches; }
extern "C" const char *b200_last_error(void) { return g_err; }

extern "C" int b200_runtime_init(void) {
  if (g_init == 1) return 0;
  if (g_init == -1) return -1;
  g_init = -1;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    snprintf(g_err, sizeof(g_err), "no CUDA device visible");
    return -1;
  }
  const char *e = getenv("SCS_B200_DEVICE");
  if (!e) e = getenv("LOCAL_RANK");
  g_dev = e ? atoi(e) % ndev : 0;
  CUDA_OK(cudaSetDevice(g_dev));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, g_dev));
  if (prop.major != 10) {
    snprintf(g_err, sizeof(g_err), "device %d is sm_%d%d; this library is built for sm_100a only",
             g_dev, prop.major, prop.minor);
    return -1;
  }
  g_num_sms = prop.multiProcessorCount;
  CUDA_OK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreate(&g_ev0));
  CUDA_OK(cudaEventCreate(&g_ev1));
  g_stage_bytes = (size_t)64 << 20;
  CUDA_OK(cudaMallocHost(&g_stage, g_stage_bytes));
  g_init = 1;
  return 0;
}
extern "C" int b200_device_ok(void) { return b200_runtime_init() == 0; }
extern "C" int b200_num_sms(void) { return g_num_sms; }
extern "C" void *b200_stream(void) { return (void *)g_stream; }

extern "C" void *b200_malloc(size_t bytes) {
  if (b200_runtime_init() != 0) return nullptr;
  void *p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    b200_set_error("cudaMalloc", e, __FILE__, __LINE__);
    return nullptr;
  }
  return p;
}
extern "C" void b200_free(void *p) {
  if (p) cudaFree(p);
}
extern "C" void *b200_host_alloc(size_t bytes) {
  if (b200_runtime_init() != 0) return nullptr;
  void *p = nullptr;
  if (bytes == 0) bytes = 16;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
extern "C" void b200_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

// Host<->device copies. The caller's buffer may be pageable: stage through the
// pinned bounce buffer in chunks so that the copy is a true async DMA and the
// caller's memory can be reused as soon as the call returns.
extern "C" int b200_h2d(void *d_dst, const void *src, size_t bytes) {
  if (b200_runtime_init() != 0) return -1;
  cudaPointerAttributes at;
  bool pinned = (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost);
  cudaGetLastError();
  if (pinned) {
    CUDA_OK(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, g_stream));
    return 0;
  }
  size_t off = 0;
  while (off < bytes) {
    size_t c = bytes - off < g_stage_bytes ? bytes - off : g_stage_bytes;
    CUDA_OK(cudaStreamSynchronize(g_stream));  // staging buffer free
    memcpy(g_stage, (const char *)src + off, c);
    CUDA_OK(cudaMemcpyAsync((char *)d_dst + off, g_stage, c, cudaMemcpyHostToDevice, g_stream));
    off += c;
  }
  CUDA_OK(cudaStreamSynchronize(g_stream));
  return 0;
}
extern "C" int b200_d2h(void *dst, const void *d_src, size_t bytes) {
  if (b200_runtime_init() != 0) return -1;
  cudaPointerAttributes at;
  bool pinned = (cudaPointerGetAttributes(&at, dst) == cudaSuccess && at.type == cudaMemoryTypeHost);
  cudaGetLastError();
  if (pinned) {
    CUDA_OK(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, g_stream));
    return 0;
  }
  size_t off = 0;
  while (off < bytes) {
    size_t c = bytes - off < g_stage_bytes ? bytes - off : g_stage_bytes;
    CUDA_OK(cudaMemcpyAsync(g_stage, (const char *)d_src + off, c, cudaMemcpyDeviceToHost, g_stream));
    CUDA_OK(cudaStreamSynchronize(g_stream));
    memcpy((char *)dst + off, g_stage, c);
    off += c;
  }
  return 0;
}
extern "C" int b200_d2d(void *d_dst, const void *d_src, size_t bytes) {
  CUDA_OK(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, g_stream));
  return 0;
}
extern "C" int b200_memset0(void *d_dst, size_t bytes) {
  if (b200_runtime_init() != 0) return -1;
  CUDA_OK(cudaMemsetAsync(d_dst, 0, bytes, g_stream));
  return 0;
}
extern "C" int b200_sync(void) {
  if (b200_runtime_init() != 0) return -1;
  CUDA_OK(cudaStreamSynchronize(g_stream));
  return 0;
}
extern "C" int b200_timer_start(void) {
  CUDA_OK(cudaEventRecord(g_ev0, g_stream));
  return 0;
}
extern "C" double b200_timer_stop_ms(void) {
  if (cudaEventRecord(g_ev1, g_stream) != cudaSuccess) return -1.0;
  if (cudaEventSynchronize(g_ev1) != cudaSuccess) return -1.0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, g_ev0, g_ev1) != cudaSuccess) return -1.0;
  return (double)ms;
}
