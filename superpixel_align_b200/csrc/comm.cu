// Peer-memory communicator for the dataset-wide k-means (BASELINE configs[4]): one process per
// GPU, every rank allocates one exchange buffer with cudaMalloc, exports it as a CUDA IPC handle
// and maps the buffers of all other ranks, so that the last CTA of a k-means iteration can
// store its reduced centroid sums straight into every peer's HBM over NVLink and wait for the
// peers' flags (kmeans_sweep_kernel, fused finish).  No NCCL call and no host round trip sits
// inside an iteration; the handles travel once through the caller's control plane
// (torch.distributed all_gather in dist_kmeans.py).
//
// Buffer layout (bytes): [0, 256): exchange counter; [256, 256 + 2*world*8): flags [2][world];
// then, 256-byte aligned, the inbox [2][world][pv_cap] doubles.
#include <string.h>

#include "common.cuh"

struct spalign_comm {
  int world, rank, device;
  long long pv_cap;
  size_t bytes;
  char* base[spalign::KM_MAX_WORLD];  // base[rank] = own allocation, others = IPC mappings
  bool connected;
};

namespace spalign {

static size_t flags_off() { return 256; }
static size_t inbox_off(int world) { return align_up(256 + (size_t)2 * world * 8, 256); }

int comm_fill_peer(spalign_comm_t* c, long long pv, PeerComm* out) {
  SPALIGN_REQUIRE(c && c->connected, "comm: not connected");
  SPALIGN_REQUIRE(pv <= c->pv_cap, "comm: vector of %lld doubles exceeds the capacity %lld", pv,
                  c->pv_cap);
  memset(out, 0, sizeof(*out));
  out->world = c->world;
  out->rank = c->rank;
  out->pv_cap = c->pv_cap;
  for (int r = 0; r < c->world; ++r) {
    out->flags[r] = reinterpret_cast<unsigned long long*>(c->base[r] + flags_off());
    out->inbox[r] = reinterpret_cast<double*>(c->base[r] + inbox_off(c->world));
  }
  out->xcount = reinterpret_cast<unsigned long long*>(c->base[c->rank]);
  return SPALIGN_OK;
}

}  // namespace spalign

using namespace spalign;

extern "C" int spalign_comm_create(int world, int rank, int64_t pv_cap, spalign_comm_t** out) {
  SPALIGN_REQUIRE(out != nullptr, "comm_create: NULL out");
  SPALIGN_REQUIRE(world >= 1 && world <= KM_MAX_WORLD && rank >= 0 && rank < world && pv_cap > 0,
                  "comm_create: need 1 <= world <= %d, 0 <= rank < world, pv_cap > 0",
                  KM_MAX_WORLD);
  spalign_comm* c = new spalign_comm();
  memset(c, 0, sizeof(*c));
  c->world = world;
  c->rank = rank;
  c->pv_cap = pv_cap;
  c->bytes = inbox_off(world) + (size_t)2 * world * pv_cap * sizeof(double);
  cudaError_t e = cudaGetDevice(&c->device);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&c->base[rank]), c->bytes);
  if (e == cudaSuccess) e = cudaMemset(c->base[rank], 0, c->bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("comm_create: %s", cudaGetErrorString(e));
    if (c->base[rank]) cudaFree(c->base[rank]);
    delete c;
    return SPALIGN_E_CUDA;
  }
  c->connected = world == 1;
  *out = c;
  return SPALIGN_OK;
}

extern "C" int spalign_comm_handle(spalign_comm_t* c, void* handle_out) {
  SPALIGN_REQUIRE(c && handle_out, "comm_handle: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == SPALIGN_COMM_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  SPALIGN_CUDA(cudaIpcGetMemHandle(&h, c->base[c->rank]));
  memcpy(handle_out, &h, sizeof(h));
  return SPALIGN_OK;
}

extern "C" int spalign_comm_connect(spalign_comm_t* c, const void* handles) {
  SPALIGN_REQUIRE(c && handles, "comm_connect: NULL argument");
  SPALIGN_REQUIRE(!c->connected || c->world == 1, "comm_connect: already connected");
  const char* hb = static_cast<const char*>(handles);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hb + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    SPALIGN_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->base[r] = static_cast<char*>(p);
  }
  c->connected = true;
  return SPALIGN_OK;
}

extern "C" int spalign_comm_destroy(spalign_comm_t* c) {
  if (c == nullptr) return SPALIGN_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r) {
    if (c->base[r] == nullptr) continue;
    if (r == c->rank) cudaFree(c->base[r]);
    else cudaIpcCloseMemHandle(c->base[r]);
  }
  delete c;
  return SPALIGN_OK;
}
