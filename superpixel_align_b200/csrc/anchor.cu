// f2: the reference's own pooling -- anchor-sampled superpixel align
// (superpixel_align, batch_spalign_kmeans.py:226-274): n_select member pixels per superpixel,
// each sampled bilinearly from the 4 nearest feature-cell centres, averaged.
//
//   sample   one warp per superpixel draws n_select distinct member pixels (counter-based
//            hash of (seed, row, draw); uniform over the members, found through the CSR overlap
//            matrix: rank -> cell by the running counts, -> pixel by a raster scan of the cell).
//            The reference shuffles the member list with Python's global `random` (:232); that
//            stream cannot be reproduced here, so WHICH pixels are drawn is a statistical
//            equivalent, not a replay -- callers that need the reference's anchors pass them in.
//   weights  one thread per anchor: feature coordinates p = pixel * (fh / H) + 0.5 clipped to
//            [0, f - 0.5] (the height ratio on both axes, :215, :235-240), the 4 nearest cell
//            centres among the 4 x 4 block around p (ties: lower flat cell index -- the
//            reference's argsort is unstable there), their bounding box, the four corner cells and
//            bilinear weights (:247-266) divided by the number of anchors.  Output is a CSR with
//            4 * n_select entries per row that spalign_pool_weighted consumes.
#include "common.cuh"

namespace spalign {
namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

template <typename LabelT>
__global__ void __launch_bounds__(256)
sample_anchors_kernel(const LabelT* __restrict__ labels, int n_img, int H, int W, int fh, int fw,
                      const int64_t* __restrict__ sp_off, int64_t R,
                      const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                      const int32_t* __restrict__ counts, const int32_t* __restrict__ area,
                      int n_select, unsigned long long seed, int32_t* anchors, int32_t* n_valid) {
  const int64_t r = (int64_t)blockIdx.x * 8 + warp_id();
  if (r >= R) return;
  const int lane = lane_id();
  // image of this row: sp_off is short, binary search
  int lo = 0, hi = n_img;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (sp_off[mid] <= r) lo = mid; else hi = mid;
  }
  const int img = lo;
  const long long value = r - sp_off[img];
  const int A = area[r];
  const int n = min(n_select, A);
  if (lane == 0) n_valid[r] = n;
  // distinct ranks in [0, A): lane l settles after the lanes below it
  int rk = -1;
  for (int l = 0; l < n; ++l) {
    int cand = 0;
    if (lane == l) cand = (int)(mix64(seed ^ mix64((unsigned long long)r * 64ull + l)) % (unsigned long long)A);
    cand = __shfl_sync(0xffffffffu, cand, l);
    bool again = true;
    while (again) {   // linear probing over the settled ranks (uniform control flow)
      const bool clash = lane < l && rk == cand;
      again = __any_sync(0xffffffffu, clash);
      if (again) cand = cand + 1 == A ? 0 : cand + 1;
    }
    if (lane == l) rk = cand;
  }
  int32_t* out = anchors + ((size_t)r * n_select + lane) * 2;
  if (lane < n_select && lane >= n) {
    out[0] = -1;
    out[1] = -1;
  }
  if (lane >= n) return;
  // rank -> cell through the running counts of the CSR row
  const int b = indptr[r], e = indptr[r + 1];
  int cum = 0, cell = -1;
  for (int j = b; j < e; ++j) {
    const int c = counts[j];
    if (rk < cum + c) {
      cell = indices[j];
      break;
    }
    cum += c;
  }
  int ay = -1, ax = -1;
  if (cell >= 0) {
    int want = rk - cum;
    const int cy = cell / fw, cx = cell - cy * fw;
    const int y0 = (int)(((long long)cy * H + fh - 1) / fh), y1 = (int)(((long long)(cy + 1) * H + fh - 1) / fh);
    const int x0 = (int)(((long long)cx * W + fw - 1) / fw), x1 = (int)(((long long)(cx + 1) * W + fw - 1) / fw);
    const LabelT* base = labels + (size_t)img * H * W;
    for (int y = y0; y < y1 && ay < 0; ++y)
      for (int x = x0; x < x1; ++x)
        if ((long long)base[(size_t)y * W + x] == value) {
          if (want == 0) {
            ay = y;
            ax = x;
            break;
          }
          --want;
        }
  }
  out[0] = ay;
  out[1] = ax;
}

__global__ void __launch_bounds__(256)
anchor_weights_kernel(const int32_t* __restrict__ anchors, const int32_t* __restrict__ n_valid,
                      int64_t R, int n_select, int H, int fh, int fw, int32_t* indptr,
                      int32_t* indices, double* wvals) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i <= R) indptr[i] = (int32_t)(i * 4 * n_select);
  if (i >= R * n_select) return;
  const int64_t r = i / n_select;
  const int a = (int)(i - r * n_select);
  int32_t* ci = indices + i * 4;
  double* cw = wvals + i * 4;
  const int nv = n_valid[r];
  const int ay = anchors[i * 2], ax = anchors[i * 2 + 1];
  if (a >= nv || ay < 0) {
    for (int k = 0; k < 4; ++k) {
      ci[k] = 0;
      cw[k] = 0.0;
    }
    return;
  }
  const double ratio = (double)fh / (double)H;                       // (:215)
  double py = __dadd_rn(__dmul_rn((double)ay, ratio), 0.5);
  double px = __dadd_rn(__dmul_rn((double)ax, ratio), 0.5);
  py = fmin(fmax(py, 0.0), (double)(fh - 1) + 0.5);                  // (:237-240)
  px = fmin(fmax(px, 0.0), (double)(fw - 1) + 0.5);
  // the 4 nearest centres lie in the 4 x 4 block around p; keep the 4 smallest (d2, cell)
  const int i0 = (int)floor(py - 0.5), j0 = (int)floor(px - 0.5);
  double bd[4];
  int bc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    bd[k] = __longlong_as_double(0x7ff0000000000000LL);
    bc[k] = 0x7fffffff;
  }
  for (int ii = i0 - 1; ii <= i0 + 2; ++ii) {
    if (ii < 0 || ii >= fh) continue;
    for (int jj = j0 - 1; jj <= j0 + 2; ++jj) {
      if (jj < 0 || jj >= fw) continue;
      const double dy = __dadd_rn((double)ii + 0.5, -py), dx = __dadd_rn((double)jj + 0.5, -px);
      double d = __dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx));
      int c = ii * fw + jj;
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // insertion into the sorted top 4
        if (d < bd[k] || (d == bd[k] && c < bc[k])) {
          const double td = bd[k];
          const int tc = bc[k];
          bd[k] = d;
          bc[k] = c;
          d = td;
          c = tc;
        }
      }
    }
  }
  int mny = 0x7fffffff, mxy = -1, mnx = 0x7fffffff, mxx = -1;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (bc[k] == 0x7fffffff) continue;
    const int yy = bc[k] / fw, xx = bc[k] - yy * fw;
    mny = min(mny, yy); mxy = max(mxy, yy);
    mnx = min(mnx, xx); mxx = max(mxx, xx);
  }
  const double min_y = mny + 0.5, max_y = mxy + 0.5, min_x = mnx + 0.5, max_x = mxx + 0.5;
  const double box = __dmul_rn(__dadd_rn(max_x, -min_x), __dadd_rn(max_y, -min_y));
  const double inv = 1.0 / ((double)nv);
  const double wx0 = __dadd_rn(max_x, -px), wx1 = __dadd_rn(px, -min_x);
  const double wy0 = __dadd_rn(max_y, -py), wy1 = __dadd_rn(py, -min_y);
  ci[0] = mny * fw + mnx; cw[0] = __dmul_rn(__ddiv_rn(__dmul_rn(wx0, wy0), box), inv);  // f11
  ci[1] = mxy * fw + mnx; cw[1] = __dmul_rn(__ddiv_rn(__dmul_rn(wx0, wy1), box), inv);  // f12
  ci[2] = mny * fw + mxx; cw[2] = __dmul_rn(__ddiv_rn(__dmul_rn(wx1, wy0), box), inv);  // f21
  ci[3] = mxy * fw + mxx; cw[3] = __dmul_rn(__ddiv_rn(__dmul_rn(wx1, wy1), box), inv);  // f22
}

}  // namespace
}  // namespace spalign

using namespace spalign;

extern "C" int spalign_sample_anchors(const void* labels, int label_dtype, int n_img, int H, int W,
                                      int fh, int fw, const int64_t* sp_off, int64_t n_rows,
                                      const int32_t* indptr, const int32_t* indices,
                                      const int32_t* counts, const int32_t* area, int n_select,
                                      uint64_t seed, int32_t* anchors, int32_t* n_valid,
                                      spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(labels && sp_off && indptr && indices && counts && area && anchors && n_valid,
                  "sample_anchors: NULL argument");
  SPALIGN_REQUIRE(n_img > 0 && H > 0 && W > 0 && fh > 0 && fw > 0 && n_rows > 0 && n_select > 0 &&
                      n_select <= 32,
                  "sample_anchors: bad arguments (1 <= n_select <= 32)");
  SPALIGN_REQUIRE(label_dtype == SPALIGN_I32 || label_dtype == SPALIGN_I64,
                  "sample_anchors: label_dtype must be I32 or I64");
  const unsigned grid = (unsigned)((n_rows + 7) / 8);
  if (label_dtype == SPALIGN_I32)
    sample_anchors_kernel<int32_t><<<grid, 256, 0, stream>>>(
        (const int32_t*)labels, n_img, H, W, fh, fw, sp_off, n_rows, indptr, indices, counts, area,
        n_select, (unsigned long long)seed, anchors, n_valid);
  else
    sample_anchors_kernel<int64_t><<<grid, 256, 0, stream>>>(
        (const int64_t*)labels, n_img, H, W, fh, fw, sp_off, n_rows, indptr, indices, counts, area,
        n_select, (unsigned long long)seed, anchors, n_valid);
  return check_launch("sample_anchors");
}

extern "C" int spalign_anchor_weights(const int32_t* anchors, const int32_t* n_valid, int64_t n_rows,
                                      int n_select, int H, int fh, int fw, int32_t* indptr,
                                      int32_t* indices, double* wvals, spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(anchors && n_valid && indptr && indices && wvals, "anchor_weights: NULL argument");
  SPALIGN_REQUIRE(n_rows > 0 && n_select > 0 && H > 0 && fh > 0 && fw > 0 &&
                      n_rows * 4 * n_select < 0x7fffffffLL,
                  "anchor_weights: bad arguments");
  const int64_t total = n_rows * n_select > n_rows + 1 ? n_rows * n_select : n_rows + 1;
  anchor_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      anchors, n_valid, n_rows, n_select, H, fh, fw, indptr, indices, wvals);
  return check_launch("anchor_weights");
}
