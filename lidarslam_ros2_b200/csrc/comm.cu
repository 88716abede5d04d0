// extern "C" implementation of include/b200comm.h: the result-row all-gather of the sharded loop-closure sweep through
// NCCL, bound with dlopen (no link-time dependency; a host application's already-loaded libnccl.so.2 is reused).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>

#include "../../include/b200comm.h"
#include "../../include/b200reg.h"
#include "engine.hpp"

namespace {

// the slice of nccl.h this file needs (ABI of NCCL 2.x)
struct NcclUniqueId {
  char internal[128];
};
typedef void* NcclComm;
enum { kNcclSuccess = 0, kNcclFloat32 = 7 };
using fn_get_unique_id = int (*)(NcclUniqueId*);
using fn_comm_init_rank = int (*)(NcclComm*, int, NcclUniqueId, int);
using fn_comm_destroy = int (*)(NcclComm);
using fn_all_gather = int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
using fn_error_string = const char* (*)(int);

struct Nccl {
  void* lib = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_gather all_gather = nullptr;
  fn_error_string error_string = nullptr;
};

std::mutex g_mu;
Nccl g_nccl;
thread_local std::string g_err;

bool load_nccl() {
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_nccl.lib) return true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    g_err = std::string("dlopen(libnccl.so.2): ") + dlerror();
    return false;
  }
  Nccl n;
  n.lib = lib;
  n.get_unique_id = (fn_get_unique_id)dlsym(lib, "ncclGetUniqueId");
  n.comm_init_rank = (fn_comm_init_rank)dlsym(lib, "ncclCommInitRank");
  n.comm_destroy = (fn_comm_destroy)dlsym(lib, "ncclCommDestroy");
  n.all_gather = (fn_all_gather)dlsym(lib, "ncclAllGather");
  n.error_string = (fn_error_string)dlsym(lib, "ncclGetErrorString");
  if (!n.get_unique_id || !n.comm_init_rank || !n.comm_destroy || !n.all_gather) {
    g_err = "libnccl.so.2 lacks ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclAllGather";
    return false;
  }
  g_nccl = n;
  return true;
}

int nccl_fail(int rc, const char* what) {
  g_err = std::string(what) + ": " + (g_nccl.error_string ? g_nccl.error_string(rc) : "NCCL error") + " (" + std::to_string(rc) + ")";
  return B200REG_ERR_CUDA;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  cudaGetLastError();
  return B200REG_ERR_CUDA;
}

}  // namespace

struct b200comm {
  NcclComm comm = nullptr;
  int rank = 0, world = 1, device = 0;
  cudaStream_t stream = nullptr;
  float *d_send = nullptr, *d_recv = nullptr, *h_pinned = nullptr;
  size_t cap_floats = 0;  // capacity of d_send; d_recv / h_pinned hold world times that
};

extern "C" {

const char* b200comm_last_error(void) { return g_err.c_str(); }

int b200comm_unique_id(unsigned char out[B200COMM_UNIQUE_ID_BYTES]) {
  if (!out) return B200REG_ERR_ARG;
  if (!load_nccl()) return B200REG_ERR_CUDA;
  NcclUniqueId id;
  const int rc = g_nccl.get_unique_id(&id);
  if (rc != kNcclSuccess) return nccl_fail(rc, "ncclGetUniqueId");
  std::memcpy(out, id.internal, B200COMM_UNIQUE_ID_BYTES);
  return B200REG_OK;
}

int b200comm_create(const unsigned char id_bytes[B200COMM_UNIQUE_ID_BYTES], int rank, int world, int device, b200comm_t* out) {
  if (!id_bytes || !out || world < 1 || rank < 0 || rank >= world) return B200REG_ERR_ARG;
  *out = nullptr;
  if (!load_nccl()) return B200REG_ERR_CUDA;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  b200comm* c = new b200comm();
  c->rank = rank;
  c->world = world;
  c->device = device;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete c;
    return cuda_fail(e, "cudaStreamCreate");
  }
  NcclUniqueId id;
  std::memcpy(id.internal, id_bytes, B200COMM_UNIQUE_ID_BYTES);
  const int rc = g_nccl.comm_init_rank(&c->comm, world, id, rank);
  if (rc != kNcclSuccess) {
    cudaStreamDestroy(c->stream);
    delete c;
    return nccl_fail(rc, "ncclCommInitRank");
  }
  *out = c;
  return B200REG_OK;
}

int b200comm_destroy(b200comm_t c) {
  if (!c) return B200REG_ERR_ARG;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->comm) g_nccl.comm_destroy(c->comm);
  if (c->d_send) cudaFree(c->d_send);
  if (c->d_recv) cudaFree(c->d_recv);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return B200REG_OK;
}

int b200comm_rank(b200comm_t c, int* rank, int* world) {
  if (!c) return B200REG_ERR_ARG;
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  return B200REG_OK;
}

int b200comm_all_gather_rows(b200comm_t c, const float* rows_local, int rows_per_rank, int row_floats, float* rows_all) {
  if (!c || !rows_local || !rows_all || rows_per_rank < 0 || row_floats <= 0) return B200REG_ERR_ARG;
  const size_t n = (size_t)rows_per_rank * (size_t)row_floats;
  if (n == 0) return B200REG_OK;
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  if (n > c->cap_floats) {
    if (c->d_send) cudaFree(c->d_send);
    if (c->d_recv) cudaFree(c->d_recv);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    c->d_send = c->d_recv = c->h_pinned = nullptr;
    c->cap_floats = 0;
    const size_t cap = n + n / 4 + 64;
    if ((e = cudaMalloc(&c->d_send, cap * sizeof(float))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&c->d_recv, cap * c->world * sizeof(float))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMallocHost(&c->h_pinned, cap * c->world * sizeof(float))) != cudaSuccess) return cuda_fail(e, "cudaMallocHost");
    c->cap_floats = cap;
  }
  std::memcpy(c->h_pinned, rows_local, n * sizeof(float));
  if ((e = cudaMemcpyAsync(c->d_send, c->h_pinned, n * sizeof(float), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess)
    return cuda_fail(e, "cudaMemcpyAsync H2D");
  const int rc = g_nccl.all_gather(c->d_send, c->d_recv, n, kNcclFloat32, c->comm, c->stream);  // the one collective
  if (rc != kNcclSuccess) return nccl_fail(rc, "ncclAllGather");
  if ((e = cudaMemcpyAsync(c->h_pinned, c->d_recv, n * c->world * sizeof(float), cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess)
    return cuda_fail(e, "cudaMemcpyAsync D2H");
  if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
  std::memcpy(rows_all, c->h_pinned, n * c->world * sizeof(float));
  return B200REG_OK;
}

// ---- pose board (engine.hpp): one buffer per rank, mapped by every rank through CUDA IPC ---------------------------
int b200comm_board_create(b200comm_t c, int max_rows, b200comm_board_t* out) {
  if (!c || !out || max_rows < 1 || max_rows > (1 << 16)) return B200REG_ERR_ARG;
  *out = nullptr;
  if (c->world > b200::POSE_BOARD_MAX_PEERS) {
    g_err = "pose board: at most 8 ranks (one NVSwitch domain)";
    return B200REG_ERR_ARG;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle travels as one row of 16 floats");
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  b200comm_board* b = new b200comm_board();
  b->device = c->device;
  b->view.world = c->world;
  b->view.rank = c->rank;
  b->view.rows = max_rows;
  b->view.tag = 0;
  const size_t bytes = b200::pose_board_words(c->world, max_rows) * sizeof(unsigned long long);
  int rc = B200REG_OK;
  float handles[b200::POSE_BOARD_MAX_PEERS * 16];
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  // a failure on one rank must not leave the others inside the collective: every rank always performs both all-gathers
  if ((e = cudaMalloc(&b->own, bytes)) != cudaSuccess) rc = cuda_fail(e, "cudaMalloc(pose board)");
  if (rc == B200REG_OK && (e = cudaMemset(b->own, 0, bytes)) != cudaSuccess) rc = cuda_fail(e, "cudaMemset(pose board)");
  if (rc == B200REG_OK && (e = cudaDeviceSynchronize()) != cudaSuccess) rc = cuda_fail(e, "cudaDeviceSynchronize");
  if (rc == B200REG_OK && (e = cudaMallocHost(&b->h_rows, (size_t)c->world * max_rows * 16 * sizeof(float))) != cudaSuccess)
    rc = cuda_fail(e, "cudaMallocHost");
  if (rc == B200REG_OK && (e = cudaMallocHost(&b->h_counts, (size_t)(c->world + 1) * sizeof(int))) != cudaSuccess)
    rc = cuda_fail(e, "cudaMallocHost");
  if (rc == B200REG_OK && c->world > 1 && (e = cudaIpcGetMemHandle(&mine, b->own)) != cudaSuccess) rc = cuda_fail(e, "cudaIpcGetMemHandle");
  float row[16];
  std::memcpy(row, &mine, 64);  // bytes only: no float arithmetic touches them on the way
  if (c->world > 1) {
    const int g = b200comm_all_gather_rows(c, row, 1, 16, handles);
    if (rc == B200REG_OK && g != B200REG_OK) rc = g;
  }
  b->view.peer[c->rank] = b->own;
  for (int p = 0; p < c->world && rc == B200REG_OK; p++) {
    if (p == c->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + 16 * p, 64);
    bool zero = true;
    for (size_t i = 0; i < sizeof(h); i++) zero = zero && reinterpret_cast<const unsigned char*>(&h)[i] == 0;
    if (zero) {
      g_err = "pose board: rank " + std::to_string(p) + " could not export its board";
      rc = B200REG_ERR_CUDA;
      break;
    }
    if ((e = cudaIpcOpenMemHandle(&b->opened[p], h, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess) {
      rc = cuda_fail(e, "cudaIpcOpenMemHandle");
      break;
    }
    b->view.peer[p] = static_cast<unsigned long long*>(b->opened[p]);
  }
  // second round: did every rank map every board? (a board is usable only if all of them did)
  float ok_row[16] = {rc == B200REG_OK ? 1.0f : 0.0f};
  float ok_all[b200::POSE_BOARD_MAX_PEERS * 16];
  if (c->world > 1) {
    const std::string keep = rc == B200REG_OK ? std::string() : g_err;
    const int g = b200comm_all_gather_rows(c, ok_row, 1, 16, ok_all);
    if (!keep.empty()) g_err = keep;
    if (rc == B200REG_OK && g != B200REG_OK) rc = g;
    for (int p = 0; p < c->world && rc == B200REG_OK; p++)
      if (ok_all[16 * p] != 1.0f) {
        g_err = "pose board: rank " + std::to_string(p) + " could not map the peers' boards";
        rc = B200REG_ERR_CUDA;
      }
  }
  if (rc != B200REG_OK) {
    const std::string keep = g_err;
    b200comm_board_destroy(b);
    g_err = keep;
    return rc;
  }
  *out = b;
  return B200REG_OK;
}

int b200comm_board_destroy(b200comm_board_t b) {
  if (!b) return B200REG_ERR_ARG;
  cudaSetDevice(b->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < b200::POSE_BOARD_MAX_PEERS; p++)
    if (b->opened[p]) cudaIpcCloseMemHandle(b->opened[p]);
  if (b->own) cudaFree(b->own);
  if (b->h_rows) cudaFreeHost(b->h_rows);
  if (b->h_counts) cudaFreeHost(b->h_counts);
  delete b;
  return B200REG_OK;
}

int b200comm_board_info(b200comm_board_t b, int* rank, int* world, int* max_rows) {
  if (!b) return B200REG_ERR_ARG;
  if (rank) *rank = b->view.rank;
  if (world) *world = b->view.world;
  if (max_rows) *max_rows = b->view.rows;
  return B200REG_OK;
}

}  // extern "C"
