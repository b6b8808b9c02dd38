// K2: superpixel-align pooling = CSR SpMM of the overlap matrix with the cell-major
// feature map (replaces superpixel_align(), batch_spalign_kmeans.py:210-276).
//
// Memory-bound: every 2 KB feature row (512 fp32 channels of one stride-8 cell) is read
// ~1.4 times (cells on superpixel borders belong to 2-4 rows; the repeats hit L2), each
// output row is written once.  One CTA of C/4 threads owns one superpixel: thread t holds
// channels 4t..4t+3 in registers, walks the row's (cell, count) list in ascending cell order
// and issues one coalesced 128-bit load per entry, PF entries in flight.
#include "common.cuh"

namespace spalign {
namespace {

constexpr int PF = 8;  // feature rows in flight per thread

// layout helper: [n, C, ncell] -> [n, ncell, C], 32x32 tiles through shared memory
__global__ void __launch_bounds__(256)
nchw_to_cellmajor_kernel(const float* __restrict__ src, float* __restrict__ dst, int C,
                         int ncell) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const float* s = src + (size_t)img * C * ncell;
  float* d = dst + (size_t)img * C * ncell;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    int c = c0 + ty + k, p = p0 + tx;
    if (c < C && p < ncell) tile[ty + k][tx] = s[(size_t)c * ncell + p];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    int p = p0 + ty + k, c = c0 + tx;
    if (c < C && p < ncell) d[(size_t)p * C + c] = tile[tx][ty + k];
  }
}

#ifndef POOL_ALTERNATE
#define POOL_ALTERNATE 1
#endif
// NACC float4 accumulators per thread: C = 4 * blockDim.x * NACC
template <int NACC>
__global__ void __launch_bounds__(256)
pool_rows_kernel(const float* __restrict__ feat, int C, int ncell, int pool_fw,
                 const int64_t* __restrict__ sp_off, const int32_t* __restrict__ indptr,
                 const int32_t* __restrict__ indices, const int32_t* __restrict__ counts,
                 const double* __restrict__ wvals, const int32_t* __restrict__ area, const int64_t* __restrict__ sum_y,
                 const int64_t* __restrict__ sum_x, int append_pos, float* __restrict__ out,
                 int64_t ld_out) {
  const int img = blockIdx.y;
  const int64_t row0 = sp_off[img];
  const int n_sp = (int)(sp_off[img + 1] - row0);
  if ((int)blockIdx.x >= n_sp) return;
  const int64_t r = row0 + blockIdx.x;
  const int base = indptr[r];
  const int L = indptr[r + 1] - base;
  const int C4 = C >> 2;
  const float4* f = reinterpret_cast<const float4*>(feat + (size_t)img * ncell * C);
  const int t = threadIdx.x;

  __shared__ int s_idx[256];
  __shared__ float s_cnt[256];

  float4 acc[NACC];
#pragma unroll
  for (int a = 0; a < NACC; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);

  // Cells on the border between two superpixels are read by both.  Superpixels in alternate
  // bands of the image walk their cells in opposite directions (band = the row's middle cell
  // row / the typical superpixel height), so vertical neighbours meet at their shared border
  // at the same time -- both at the start or both at the end of their lives -- and the second
  // read hits L2 instead of DRAM.  The order is a fixed function of the matrix: deterministic.
  bool reverse = false;
#if POOL_ALTERNATE
  if (L > 0 && counts != nullptr) {
    const int fw_ = pool_fw;
    const float mid = 0.5f * (float)(indices[base] / fw_ + indices[base + L - 1] / fw_);
    const float hc = sqrtf((float)ncell / (float)max(n_sp, 1));
    reverse = ((int)(mid / fmaxf(hc, 1.f)) & 1) != 0;
  }
#endif

  for (int e0 = 0; e0 < L; e0 += 256) {
    const int n = min(256, L - e0);
    __syncthreads();
    for (int e = t; e < n; e += blockDim.x) {
      const int src = reverse ? base + L - 1 - (e0 + e) : base + e0 + e;
      s_idx[e] = indices[src];
      s_cnt[e] = counts != nullptr ? (float)counts[src] : (float)wvals[src];
    }
    __syncthreads();
    int e = 0;
    for (; e + PF <= n; e += PF) {
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
        float4 v[PF];
        const int ch = t + a * blockDim.x;
#pragma unroll
        for (int k = 0; k < PF; ++k) v[k] = __ldg(f + (size_t)s_idx[e + k] * C4 + ch);
#pragma unroll
        for (int k = 0; k < PF; ++k) {
          const float w = s_cnt[e + k];
          acc[a].x = fmaf(w, v[k].x, acc[a].x);
          acc[a].y = fmaf(w, v[k].y, acc[a].y);
          acc[a].z = fmaf(w, v[k].z, acc[a].z);
          acc[a].w = fmaf(w, v[k].w, acc[a].w);
        }
      }
    }
    for (; e < n; ++e) {
      const float w = s_cnt[e];
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
        const float4 v = __ldg(f + (size_t)s_idx[e] * C4 + t + a * blockDim.x);
        acc[a].x = fmaf(w, v.x, acc[a].x);
        acc[a].y = fmaf(w, v.y, acc[a].y);
        acc[a].z = fmaf(w, v.z, acc[a].z);
        acc[a].w = fmaf(w, v.w, acc[a].w);
      }
    }
  }
  const float ar = (float)area[r];
  float4* o = reinterpret_cast<float4*>(out + (size_t)r * ld_out);
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    float4 q = acc[a];
    q.x /= ar; q.y /= ar; q.z /= ar; q.w /= ar;
    o[t + a * blockDim.x] = q;
  }
  if (t == 0) {
    float* tail = out + (size_t)r * ld_out + C;
    int k = 0;
    if (append_pos) {
      const double a = (double)area[r];
      tail[0] = (float)((double)sum_y[r] / a);
      tail[1] = (float)((double)sum_x[r] / a);
      k = 2;
    }
    for (; C + k < ld_out; ++k) tail[k] = 0.f;
  }
}

}  // namespace
}  // namespace spalign

using namespace spalign;

extern "C" int spalign_nchw_to_cellmajor(const float* src, float* dst, int n_img, int C,
                                         int ncell, spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(src && dst && n_img > 0 && C > 0 && ncell > 0, "nchw_to_cellmajor: bad args");
  SPALIGN_REQUIRE(n_img <= 65535 && (C + 31) / 32 <= 65535, "nchw_to_cellmajor: grid too large");
  dim3 grid((ncell + 31) / 32, (C + 31) / 32, n_img);
  nchw_to_cellmajor_kernel<<<grid, 256, 0, stream>>>(src, dst, C, ncell);
  return check_launch("nchw_to_cellmajor");
}

static int pool_impl(const float* feat, int n_img, int C, int fh, int fw, const int64_t* sp_off,
                     int64_t n_rows, int max_rows_per_image, const int32_t* indptr,
                     const int32_t* indices, const int32_t* counts, const double* wvals,
                     const int32_t* area, const int64_t* sum_y, const int64_t* sum_x,
                     int append_pos, float* out, int64_t ld_out, spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(feat && sp_off && indptr && indices && (counts || wvals) && area && out,
                  "pool: NULL argument");
  SPALIGN_REQUIRE(!append_pos || (sum_y && sum_x), "pool: append_pos needs sum_y/sum_x");
  SPALIGN_REQUIRE(n_img > 0 && n_img <= 65535 && fh > 0 && fw > 0 && n_rows > 0 &&
                      max_rows_per_image > 0,
                  "pool: bad shape");
  SPALIGN_REQUIRE(C >= 4 && C % 4 == 0 && C <= 4096, "pool: C must be a multiple of 4, <= 4096");
  SPALIGN_REQUIRE(ld_out >= C + (append_pos ? 2 : 0) && ld_out % 4 == 0,
                  "pool: ld_out must be >= C+2*append_pos and a multiple of 4");
  SPALIGN_REQUIRE(reinterpret_cast<size_t>(feat) % 16 == 0 &&
                      reinterpret_cast<size_t>(out) % 16 == 0,
                  "pool: feat/out must be 16-byte aligned");
  const int ncell = fh * fw;
  const int C4 = C / 4;
  dim3 grid(max_rows_per_image, n_img);
  // threads * NACC == C/4
  int nacc = 1, threads = C4;
  while (threads > 256 || (threads % 32 != 0 && threads > 32)) {
    // prefer a whole number of warps; fall back to splitting over accumulators
    if (threads % 2 == 0 && nacc < 4) {
      threads /= 2;
      nacc *= 2;
    } else {
      break;
    }
  }
  SPALIGN_REQUIRE(threads <= 256 && threads * nacc == C4,
                  "pool: unsupported channel count %d", C);
#define LAUNCH(N)                                                                              \
  pool_rows_kernel<N><<<grid, threads, 0, stream>>>(feat, C, ncell, fw, sp_off, indptr, indices, \
                                                    counts, wvals, area, sum_y, sum_x,         \
                                                    append_pos, out, ld_out)
  if (nacc == 1) LAUNCH(1);
  else if (nacc == 2) LAUNCH(2);
  else LAUNCH(4);
#undef LAUNCH
  return check_launch("pool");
}

extern "C" int spalign_pool(const float* feat, int n_img, int C, int fh, int fw,
                            const int64_t* sp_off, int64_t n_rows, int max_rows_per_image,
                            const int32_t* indptr, const int32_t* indices,
                            const int32_t* counts, const int32_t* area, const int64_t* sum_y,
                            const int64_t* sum_x, int append_pos, float* out, int64_t ld_out,
                            spalign_stream_t stream) {
  SPALIGN_REQUIRE(counts != nullptr, "pool: NULL counts");
  return pool_impl(feat, n_img, C, fh, fw, sp_off, n_rows, max_rows_per_image, indptr, indices,
                   counts, nullptr, area, sum_y, sum_x, append_pos, out, ld_out, stream);
}

extern "C" int spalign_pool_weighted(const float* feat, int n_img, int C, int fh, int fw,
                                     const int64_t* sp_off, int64_t n_rows,
                                     int max_rows_per_image, const int32_t* indptr,
                                     const int32_t* indices, const double* wvals,
                                     const int32_t* area, const int64_t* sum_y,
                                     const int64_t* sum_x, int append_pos, float* out,
                                     int64_t ld_out, spalign_stream_t stream) {
  SPALIGN_REQUIRE(wvals != nullptr, "pool_weighted: NULL wvals");
  return pool_impl(feat, n_img, C, fh, fw, sp_off, n_rows, max_rows_per_image, indptr, indices,
                   nullptr, wvals, area, sum_y, sum_x, append_pos, out, ld_out, stream);
}
