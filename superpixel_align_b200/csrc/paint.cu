// K4 paint-back, K5 overlap refine, and the 2-class confusion count.
#include <stdarg.h>

#include "common.cuh"

namespace spalign {

// thread-local last error (declared in common.cuh)
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

template <typename T>
struct Pack4;  // four consecutive elements
template <>
struct Pack4<int32_t> {
  int4 v;
  __device__ void load(const int32_t* p) { v = ld_stream_int4(p); }
  __device__ long long get(int i) const { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
};
template <>
struct Pack4<int64_t> {
  longlong2 a, b;
  __device__ void load(const int64_t* p) {
    a = __ldg(reinterpret_cast<const longlong2*>(p));
    b = __ldg(reinterpret_cast<const longlong2*>(p) + 1);
  }
  __device__ long long get(int i) const { return i == 0 ? a.x : i == 1 ? a.y : i == 2 ? b.x : b.y; }
};

template <typename OutT>
__device__ __forceinline__ void store4(OutT* p, const int* c);
template <>
__device__ __forceinline__ void store4<uint8_t>(uint8_t* p, const int* c) {
  *reinterpret_cast<uchar4*>(p) = make_uchar4((unsigned char)c[0], (unsigned char)c[1],
                                              (unsigned char)c[2], (unsigned char)c[3]);
}
template <>
__device__ __forceinline__ void store4<int32_t>(int32_t* p, const int* c) {
  *reinterpret_cast<int4*>(p) = make_int4(c[0], c[1], c[2], c[3]);
}
template <>
__device__ __forceinline__ void store4<int64_t>(int64_t* p, const int* c) {
  reinterpret_cast<longlong2*>(p)[0] = make_longlong2(c[0], c[1]);
  reinterpret_cast<longlong2*>(p)[1] = make_longlong2(c[2], c[3]);
}

// out[p] = table[sp_off[img] + label[p]]; 4 pixels per thread (n_pix % 4 == 0 fast path)
template <typename LabelT, typename OutT>
__global__ void __launch_bounds__(256)
paint_kernel(const LabelT* __restrict__ labels, int64_t n_pix, const int64_t* __restrict__ sp_off,
             const int32_t* __restrict__ table, OutT* cluster_map, uint8_t* road_mask,
             int road_value) {
  const int img = blockIdx.y;
  const int64_t row0 = sp_off[img];
  const long long n_sp = sp_off[img + 1] - row0;
  const int32_t* tb = table + row0;
  const LabelT* lab = labels + (size_t)img * n_pix;
  OutT* cm = cluster_map ? cluster_map + (size_t)img * n_pix : nullptr;
  uint8_t* rm = road_mask ? road_mask + (size_t)img * n_pix : nullptr;
  const int64_t n4 = n_pix >> 2;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    Pack4<LabelT> pk;
    pk.load(lab + 4 * i);
    int c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long l = pk.get(k);
      c[k] = (l >= 0 && l < n_sp) ? __ldg(tb + l) : 0;
    }
    if (cm) store4<OutT>(cm + 4 * i, c);
    if (rm)
      *reinterpret_cast<uchar4*>(rm + 4 * i) =
          make_uchar4(c[0] == road_value, c[1] == road_value, c[2] == road_value,
                      c[3] == road_value);
  }
  // tail (n_pix not a multiple of 4)
  if (blockIdx.x == 0) {
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n_pix; i += 256) {
      const long long l = (long long)lab[i];
      const int c = (l >= 0 && l < n_sp) ? tb[l] : 0;
      if (cm) cm[i] = (OutT)c;
      if (rm) rm[i] = c == road_value;
    }
  }
}

// overlap[r] = sum_j counts[j] * road_cell[img, indices[j]]  (one warp per row)
__global__ void __launch_bounds__(256)
refine_overlap_kernel(const int64_t* __restrict__ sp_off, int ncell,
                      const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                      const int32_t* __restrict__ counts, const uint8_t* __restrict__ road_cell,
                      int64_t* overlap, int64_t* road_px) {
  const int img = blockIdx.y;
  const int64_t row0 = sp_off[img];
  const int n_sp = (int)(sp_off[img + 1] - row0);
  const int s = blockIdx.x * 8 + warp_id();
  if (s >= n_sp) return;
  const int64_t r = row0 + s;
  const uint8_t* road = road_cell + (size_t)img * ncell;
  long long ov = 0;
  for (int j = indptr[r] + lane_id(); j < indptr[r + 1]; j += 32)
    ov += road[indices[j]] ? counts[j] : 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ov += __shfl_xor_sync(0xffffffffu, ov, d);
  if (lane_id() == 0) {
    overlap[r] = ov;
    if (ov) atomicAdd(reinterpret_cast<unsigned long long*>(&road_px[img]), (unsigned long long)ov);
  }
}

__global__ void __launch_bounds__(256)
refine_keep_kernel(const int64_t* __restrict__ sp_off, const int64_t* __restrict__ overlap,
                   const int64_t* __restrict__ road_px, double thr, int32_t* keep) {
  const int img = blockIdx.y;
  const int64_t row0 = sp_off[img];
  const int n_sp = (int)(sp_off[img + 1] - row0);
  const int s = blockIdx.x * 256 + threadIdx.x;
  if (s >= n_sp) return;
  const long long rp = road_px[img];
  // superpixel_overlaps.py:368: n_pred_road_pixels > 0 and overlap / n_pred_road_pixels > thr
  keep[row0 + s] = (rp > 0 && ((double)overlap[row0 + s] / (double)rp) > thr) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
confusion2_kernel(const uint8_t* __restrict__ pred, const int32_t* __restrict__ gt, int64_t n_pix,
                  int64_t* conf) {
  const int img = blockIdx.y;
  const uint8_t* p = pred + (size_t)img * n_pix;
  const int32_t* g = gt + (size_t)img * n_pix;
  int c[4] = {0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_pix; i += (int64_t)gridDim.x * 256) {
    const int gv = g[i];
    if (gv >= 0) {
      const int idx = 2 * (gv ? 1 : 0) + (p[i] ? 1 : 0);
      c[0] += idx == 0; c[1] += idx == 1; c[2] += idx == 2; c[3] += idx == 3;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c[k] += __shfl_xor_sync(0xffffffffu, c[k], d);
  }
  if (lane_id() == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (c[k]) atomicAdd(reinterpret_cast<unsigned long long*>(&conf[(size_t)img * 4 + k]),
                          (unsigned long long)c[k]);
  }
}

}  // namespace
}  // namespace spalign

using namespace spalign;

extern "C" int spalign_abi_version(void) { return SPALIGN_ABI_VERSION; }
extern "C" const char* spalign_last_error(void) { return g_err; }

extern "C" int spalign_paint(const void* labels, int label_dtype, int n_img, int H, int W,
                             const int64_t* sp_off, const int32_t* table, void* cluster_map,
                             int out_dtype, uint8_t* road_mask, int road_value,
                             spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(labels && sp_off && table && (cluster_map || road_mask), "paint: NULL argument");
  SPALIGN_REQUIRE(n_img > 0 && n_img <= 65535 && H > 0 && W > 0, "paint: bad shape");
  SPALIGN_REQUIRE(label_dtype == SPALIGN_I32 || label_dtype == SPALIGN_I64, "paint: bad label_dtype");
  SPALIGN_REQUIRE(out_dtype == SPALIGN_U8 || out_dtype == SPALIGN_I32 || out_dtype == SPALIGN_I64,
                  "paint: bad out_dtype");
  const int64_t n_pix = (int64_t)H * W;
  SPALIGN_REQUIRE(reinterpret_cast<size_t>(labels) % 16 == 0 &&
                      (!cluster_map || reinterpret_cast<size_t>(cluster_map) % 16 == 0) &&
                      (!road_mask || reinterpret_cast<size_t>(road_mask) % 4 == 0),
                  "paint: buffers must be 16-byte aligned");
  SPALIGN_REQUIRE(n_img == 1 || n_pix % 4 == 0, "paint: H*W must be a multiple of 4 for batches");
  int gx = (int)((n_pix / 4 + 256 * 4 - 1) / (256 * 4));
  gx = gx < 1 ? 1 : (gx > 8 * kNumSMs ? 8 * kNumSMs : gx);
  dim3 grid(gx, n_img);
#define PAINT(LT, OT)                                                                          \
  paint_kernel<LT, OT><<<grid, 256, 0, stream>>>((const LT*)labels, n_pix, sp_off, table,      \
                                                 (OT*)cluster_map, road_mask, road_value)
  if (label_dtype == SPALIGN_I32) {
    if (out_dtype == SPALIGN_U8) PAINT(int32_t, uint8_t);
    else if (out_dtype == SPALIGN_I32) PAINT(int32_t, int32_t);
    else PAINT(int32_t, int64_t);
  } else {
    if (out_dtype == SPALIGN_U8) PAINT(int64_t, uint8_t);
    else if (out_dtype == SPALIGN_I32) PAINT(int64_t, int32_t);
    else PAINT(int64_t, int64_t);
  }
#undef PAINT
  return check_launch("paint");
}

extern "C" int spalign_refine(const int64_t* sp_off, int n_img, int64_t n_rows, int ncell,
                              int max_rows_per_image, const int32_t* indptr,
                              const int32_t* indices, const int32_t* counts,
                              const uint8_t* road_cell, double thr, int64_t* overlap,
                              int64_t* road_px, int32_t* keep, spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(sp_off && indptr && indices && counts && road_cell && overlap && road_px && keep,
                  "refine: NULL argument");
  SPALIGN_REQUIRE(n_img > 0 && n_img <= 65535 && n_rows > 0 && ncell > 0 && max_rows_per_image > 0,
                  "refine: bad shape");
  SPALIGN_CUDA(cudaMemsetAsync(road_px, 0, sizeof(int64_t) * n_img, stream));
  refine_overlap_kernel<<<dim3((max_rows_per_image + 7) / 8, n_img), 256, 0, stream>>>(
      sp_off, ncell, indptr, indices, counts, road_cell, overlap, road_px);
  refine_keep_kernel<<<dim3((max_rows_per_image + 255) / 256, n_img), 256, 0, stream>>>(
      sp_off, overlap, road_px, thr, keep);
  return check_launch("refine");
}

// cv2.resize(..., interpolation=cv2.INTER_NEAREST): source index = min(floor(dst * ifx), src - 1)
// with ifx = 1.0 / (dst_size / src_size) in double, as OpenCV computes it
__global__ void __launch_bounds__(256)
resize_nearest_u8_kernel(const uint8_t* __restrict__ src, int h, int w, uint8_t* __restrict__ dst,
                         int H, int W, double sy, double sx) {
  const int img = blockIdx.z;
  const int x = blockIdx.x * 64 + (threadIdx.x & 63);
  const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x >= W || y >= H) return;
  const int yy = min((int)floor(__dmul_rn((double)y, sy)), h - 1);
  const int xx = min((int)floor(__dmul_rn((double)x, sx)), w - 1);
  dst[((size_t)img * H + y) * W + x] = src[((size_t)img * h + yy) * w + xx];
}

extern "C" int spalign_resize_nearest_u8(const uint8_t* src, int n_img, int h, int w, uint8_t* dst,
                                         int H, int W, spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(src && dst && n_img > 0 && n_img <= 65535 && h > 0 && w > 0 && H > 0 && W > 0,
                  "resize_nearest_u8: bad arguments");
  const dim3 grid((W + 63) / 64, (H + 3) / 4, n_img);
  // OpenCV: inv_scale = dst / src (double), ifx = 1.0 / inv_scale, sx = cvFloor(x * ifx)
  const double ify = 1.0 / ((double)H / (double)h), ifx = 1.0 / ((double)W / (double)w);
  resize_nearest_u8_kernel<<<grid, 256, 0, stream>>>(src, h, w, dst, H, W, ify, ifx);
  return check_launch("resize_nearest_u8");
}

extern "C" int spalign_confusion2(const uint8_t* pred, const int32_t* gt, int n_img,
                                  int64_t n_pix, int64_t* conf, spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(pred && gt && conf && n_img > 0 && n_img <= 65535 && n_pix > 0,
                  "confusion2: bad arguments");
  SPALIGN_CUDA(cudaMemsetAsync(conf, 0, sizeof(int64_t) * 4 * n_img, stream));
  int gx = (int)((n_pix + 256 * 16 - 1) / (256 * 16));
  gx = gx < 1 ? 1 : (gx > 4 * kNumSMs ? 4 * kNumSMs : gx);
  confusion2_kernel<<<dim3(gx, n_img), 256, 0, stream>>>(pred, gt, n_pix, conf);
  return check_launch("confusion2");
}
