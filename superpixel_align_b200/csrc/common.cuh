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

// K4 fused into the k-means finish kernel: work description + the tile painter both kernels use.
// One tile = PAINT_TILE pixels of one image; tiles of an image are handed out through an atomic
// counter, so any CTA with nothing better to do can paint any finished image.
constexpr int PAINT_TILE = 8192;
struct PaintJob {
  const int32_t* labels;   // [n_img][n_pix] int32 (NULL: no painting)
  const int64_t* sp_off;   // [n_img+1]
  const int32_t* table;    // assignment per row (read through L2: written by other SMs)
  uint8_t* cluster_map;    // [n_img][n_pix] or NULL
  uint8_t* road_mask;      // [n_img][n_pix] or NULL
  int32_t* next_tile;      // [n_img] tile counters, zero on entry
  long long n_pix;
  int n_img;
  int road_value;
};

// Paints tiles of image `img` until its counter runs out.  Whole block (256 threads), uniform
// control flow; returns the number of tiles this block painted.
__device__ __forceinline__ int paint_image_tiles(const PaintJob& pj, int img, int* s_tile) {
  const long long n_pix = pj.n_pix;
  const int n_tiles = (int)((n_pix + PAINT_TILE - 1) / PAINT_TILE);
  const int64_t row0 = pj.sp_off[img];
  const int n_sp = (int)(pj.sp_off[img + 1] - row0);
  const int32_t* lab = pj.labels + (size_t)img * n_pix;
  const int32_t* tab = pj.table + row0;
  int painted = 0;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) *s_tile = atomicAdd(&pj.next_tile[img], 1);
    __syncthreads();
    const int tile = *s_tile;
    if (tile >= n_tiles) break;
    ++painted;
    const long long p0 = (long long)tile * PAINT_TILE;
#pragma unroll 2
    for (int i = threadIdx.x * 4; i < PAINT_TILE; i += blockDim.x * 4) {
      const long long p = p0 + i;
      if (p + 3 < n_pix && (n_pix & 3) == 0) {
        const int4 q = ld_stream_int4(lab + p);
        int c[4];
        c[0] = (unsigned)q.x < (unsigned)n_sp ? __ldcg(tab + q.x) : 0;
        c[1] = (unsigned)q.y < (unsigned)n_sp ? __ldcg(tab + q.y) : 0;
        c[2] = (unsigned)q.z < (unsigned)n_sp ? __ldcg(tab + q.z) : 0;
        c[3] = (unsigned)q.w < (unsigned)n_sp ? __ldcg(tab + q.w) : 0;
        if (pj.cluster_map)
          *reinterpret_cast<uchar4*>(pj.cluster_map + (size_t)img * n_pix + p) =
              make_uchar4((unsigned char)c[0], (unsigned char)c[1], (unsigned char)c[2], (unsigned char)c[3]);
        if (pj.road_mask)
          *reinterpret_cast<uchar4*>(pj.road_mask + (size_t)img * n_pix + p) =
              make_uchar4(c[0] == pj.road_value, c[1] == pj.road_value, c[2] == pj.road_value,
                          c[3] == pj.road_value);
      } else {
        for (int j = 0; j < 4 && p + j < n_pix; ++j) {
          const int q = lab[p + j];
          const int c = (unsigned)q < (unsigned)n_sp ? __ldcg(tab + q) : 0;
          if (pj.cluster_map) pj.cluster_map[(size_t)img * n_pix + p + j] = (unsigned char)c;
          if (pj.road_mask) pj.road_mask[(size_t)img * n_pix + p + j] = c == pj.road_value;
        }
      }
    }
  }
  return painted;
}

}  // namespace spalign
