// Shared helpers for libspalign_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "spalign.h"

namespace spalign {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: %s", what, cudaGetErrorString(e));
    return SPALIGN_E_CUDA;
  }
  return SPALIGN_OK;
}

#define SPALIGN_REQUIRE(cond, ...)      \
  do {                                  \
    if (!(cond)) {                      \
      ::spalign::set_error(__VA_ARGS__); \
      return SPALIGN_E_INVALID;         \
    }                                   \
  } while (0)

#define SPALIGN_CUDA(call)                                                   \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) {                                                \
      ::spalign::set_error("%s: %s", #call, cudaGetErrorString(e__));        \
      return SPALIGN_E_CUDA;                                                 \
    }                                                                        \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// carve a workspace pointer
struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* p) : base(static_cast<char*>(p)), off(0) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
  size_t used() const { return align_up(off, 256); }
};

constexpr int kNumSMs = 148;  // B200

// Multi-GPU exchange state as the kernels see it (spalign_comm_* in comm.cu): every rank owns
// an inbox [2][world][pv_cap] doubles and flags [2][world] in its own HBM, mapped into every
// peer process (CUDA IPC, loads/stores travel over NVLink).
constexpr int KM_MAX_WORLD = 8;
struct PeerComm {
  int world, rank;
  long long pv_cap;
  double* inbox[KM_MAX_WORLD];              // rank r's inbox as seen from this process
  unsigned long long* flags[KM_MAX_WORLD];  // rank r's flags as seen from this process
  unsigned long long* xcount;               // local: number of exchanges completed so far
};
// fills `out` for vectors of `pv` doubles (comm.cu)
int comm_fill_peer(spalign_comm_t* comm, long long pv, PeerComm* out);

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// streaming 128-bit loads that do not allocate in L1 (data touched once)
__device__ __forceinline__ int4 ld_stream_int4(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream_float4(const void* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

}  // namespace spalign
