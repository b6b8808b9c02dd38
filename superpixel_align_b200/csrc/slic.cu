// f3: SLIC superpixels on the device (replaces batch_superpixel, batch_spalign_kmeans.py:299-313:
// skimage.segmentation.slic(img.transpose(1, 2, 0), n_segments) per image on the CPU).
//
// scikit-image 0.13.1 is neither in the reference tree nor installed: PARITY UNPINNED.  The
// contract is the published algorithm of that version, restated in oracle/spalign_oracle.py
// (slic, slic_enforce_connectivity), with two choices that make the result bit-reproducible
// and identical between this file and the NumPy restatement:
//   * Lab colours / compactness are quantised to multiples of 2^-12 once; the centre sums are
//     64-bit integer atomics (exact, order independent); every floating-point expression below
//     is written with explicit round-to-nearest operations in the oracle's order (no FMA);
//   * connectivity: 4-connected components by union-find (root = first pixel in raster order);
//     a component below min_size joins the segment LEFT of its first pixel, else the one ABOVE;
//     final ids are numbered in raster order of first pixels (as skimage numbers them).
//
// Launch sequence per batch (no host synchronisation; the iteration loop stops on a device flag):
//   quantise -> seed centres -> max_iter x [assign -> accumulate -> update] -> union-find
//   (init, link, flatten) -> sizes -> merge targets (+ pointer jumping) -> compaction scan -> ids
#include "common.cuh"

namespace spalign {
namespace {

constexpr double SLIC_Q = 4096.0;
constexpr int SCAN_PER_BLOCK = 2048;  // pixels per block of the compaction scan

struct SlicGrid {
  int y0, sy, x0, sx, ny, nx;
};

struct SlicWs {
  int4* q;             // [n][P] quantised colours (x, y, z, unused)
  int* nearest;        // [n][P]
  int* parent;         // [n][P] union-find, then per-pixel final root
  int* size;           // [n][P] component sizes at roots
  int* target;         // [n][P] merge target of each root
  int* block_count;    // [n][n_scan_blocks]
  double* centre;      // [n][n_seg][5] cy, cx, c0, c1, c2
  unsigned long long* sums;  // [n][n_seg][6] count, y, x, q0, q1, q2
  int* flags;          // [n][4]: changed, stop, -, -
};

size_t slic_carve(SlicWs& ws, void* base, int n, int64_t P, int n_seg) {
  Carver c(base);
  const int64_t nb = (P + SCAN_PER_BLOCK - 1) / SCAN_PER_BLOCK;
  ws.q = c.take<int4>((size_t)n * P);
  ws.nearest = c.take<int>((size_t)n * P);
  ws.parent = c.take<int>((size_t)n * P);
  ws.size = c.take<int>((size_t)n * P);
  ws.target = c.take<int>((size_t)n * P);
  ws.block_count = c.take<int>((size_t)n * nb);
  ws.centre = c.take<double>((size_t)n * n_seg * 5);
  ws.sums = c.take<unsigned long long>((size_t)n * n_seg * 6);
  ws.flags = c.take<int>((size_t)n * 4);
  return c.used();
}

// skimage.color.rgb2lab on the values as given (D65, 2 degree observer)
__device__ __forceinline__ void rgb2lab_f64(double r, double g, double b, double* out) {
  double lin[3] = {r, g, b};
#pragma unroll
  for (int i = 0; i < 3; ++i)
    lin[i] = lin[i] > 0.04045 ? pow(__ddiv_rn(__dadd_rn(lin[i], 0.055), 1.055), 2.4)
                              : __ddiv_rn(lin[i], 12.92);
  const double m[3][3] = {{0.412453, 0.357580, 0.180423},
                          {0.212671, 0.715160, 0.072169},
                          {0.019334, 0.119193, 0.950227}};
  const double white[3] = {0.95047, 1.0, 1.08883};
  double f[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double v = __dmul_rn(lin[0], m[i][0]);
    v = __dadd_rn(v, __dmul_rn(lin[1], m[i][1]));
    v = __dadd_rn(v, __dmul_rn(lin[2], m[i][2]));
    v = __ddiv_rn(v, white[i]);
    f[i] = v > 0.008856 ? cbrt(v) : __dadd_rn(__dmul_rn(7.787, v), 16.0 / 116.0);
  }
  out[0] = __dadd_rn(__dmul_rn(116.0, f[1]), -16.0);
  out[1] = __dmul_rn(500.0, __dadd_rn(f[0], -f[1]));
  out[2] = __dmul_rn(200.0, __dadd_rn(f[1], -f[2]));
}

__global__ void __launch_bounds__(256)
slic_quantise_kernel(const float* __restrict__ img, int64_t P, double inv_compactness,
                     int convert2lab, SlicWs ws) {
  const int im = blockIdx.y;
  const float* src = img + (size_t)im * 3 * P;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < P; p += (int64_t)gridDim.x * 256) {
    double c[3] = {(double)src[p], (double)src[P + p], (double)src[2 * P + p]};
    if (convert2lab) rgb2lab_f64(c[0], c[1], c[2], c);
    int4 o;
    o.x = (int)rint(__dmul_rn(__dmul_rn(c[0], inv_compactness), SLIC_Q));
    o.y = (int)rint(__dmul_rn(__dmul_rn(c[1], inv_compactness), SLIC_Q));
    o.z = (int)rint(__dmul_rn(__dmul_rn(c[2], inv_compactness), SLIC_Q));
    o.w = 0;
    ws.q[(size_t)im * P + p] = o;
  }
}

__global__ void __launch_bounds__(256)
slic_seed_kernel(SlicGrid g, int W, int64_t P, int n_seg, SlicWs ws) {
  const int im = blockIdx.y;
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k == 0) {
    ws.flags[im * 4 + 0] = 0;
    ws.flags[im * 4 + 1] = 0;
  }
  if (k >= n_seg) return;
  const int y = g.y0 + (k / g.nx) * g.sy, x = g.x0 + (k % g.nx) * g.sx;
  const int4 c = ws.q[(size_t)im * P + (size_t)y * W + x];
  double* ce = ws.centre + ((size_t)im * n_seg + k) * 5;
  ce[0] = (double)y;
  ce[1] = (double)x;
  ce[2] = __ddiv_rn((double)c.x, SLIC_Q);
  ce[3] = __ddiv_rn((double)c.y, SLIC_Q);
  ce[4] = __ddiv_rn((double)c.z, SLIC_Q);
}

__global__ void __launch_bounds__(256)
slic_zero_nearest_kernel(int64_t total, SlicWs ws) {
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < total; p += (int64_t)gridDim.x * 256)
    ws.nearest[p] = 0;
}

// one thread per pixel: candidates are the segments seeded within 3 grid nodes of the pixel's
// nearest node, visited in ascending id; the exact 2*step window test and strict '<' of the
// reference loop decide (the lower id wins ties)
__global__ void __launch_bounds__(256)
slic_assign_kernel(SlicGrid g, int H, int W, int n_seg, double wsp, SlicWs ws) {
  const int im = blockIdx.z;
  if (ws.flags[im * 4 + 1]) return;  // converged earlier
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const int64_t P = (int64_t)H * W;
  const size_t p = (size_t)im * P + (size_t)y * W + x;
  const int4 qc = ws.q[p];
  const double c0 = __ddiv_rn((double)qc.x, SLIC_Q), c1 = __ddiv_rn((double)qc.y, SLIC_Q),
               c2 = __ddiv_rn((double)qc.z, SLIC_Q);
  int iy = (y - g.y0 + g.sy / 2) / g.sy, ix = (x - g.x0 + g.sx / 2) / g.sx;
  iy = min(max(iy, 0), g.ny - 1);
  ix = min(max(ix, 0), g.nx - 1);
  const double* cen = ws.centre + (size_t)im * n_seg * 5;
  const int old = ws.nearest[p];
  int best = old;
  double dist = __longlong_as_double(0x7ff0000000000000LL);
  const double wy = 2.0 * g.sy, wx = 2.0 * g.sx;
  for (int jy = max(iy - 3, 0); jy <= min(iy + 3, g.ny - 1); ++jy) {
    for (int jx = max(ix - 3, 0); jx <= min(ix + 3, g.nx - 1); ++jx) {
      const int k = jy * g.nx + jx;
      const double* ce = cen + (size_t)k * 5;
      const double cy = ce[0], cx = ce[1];
      const int ya = (int)fmax(__dadd_rn(cy, -wy), 0.0), yb = (int)fmin(__dadd_rn(__dadd_rn(cy, wy), 1.0), (double)H);
      const int xa = (int)fmax(__dadd_rn(cx, -wx), 0.0), xb = (int)fmin(__dadd_rn(__dadd_rn(cx, wx), 1.0), (double)W);
      if (y < ya || y >= yb || x < xa || x >= xb) continue;
      const double dy = __dadd_rn(cy, -(double)y), dx = __dadd_rn(cx, -(double)x);
      double d = __dmul_rn(__dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx)), wsp);
      const double e0 = __dadd_rn(c0, -ce[2]), e1 = __dadd_rn(c1, -ce[3]), e2 = __dadd_rn(c2, -ce[4]);
      d = __dadd_rn(d, __dadd_rn(__dadd_rn(__dmul_rn(e0, e0), __dmul_rn(e1, e1)), __dmul_rn(e2, e2)));
      if (d < dist) {
        dist = d;
        best = k;
      }
    }
  }
  if (best != old) {
    ws.nearest[p] = best;
    ws.flags[im * 4 + 0] = 1;
  }
}

__global__ void __launch_bounds__(256)
slic_zero_sums_kernel(int n_seg, SlicWs ws) {
  const int im = blockIdx.y;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n_seg * 6; i += gridDim.x * 256)
    ws.sums[(size_t)im * n_seg * 6 + i] = 0ull;
}

__global__ void __launch_bounds__(256)
slic_accumulate_kernel(int H, int W, int n_seg, SlicWs ws) {
  const int im = blockIdx.z;
  if (ws.flags[im * 4 + 1] || !ws.flags[im * 4 + 0]) return;  // stopped, or nothing changed
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const size_t p = (size_t)im * H * W + (size_t)y * W + x;
  const int k = ws.nearest[p];
  const int4 qc = ws.q[p];
  unsigned long long* s = ws.sums + ((size_t)im * n_seg + k) * 6;
  atomicAdd(s + 0, 1ull);
  atomicAdd(s + 1, (unsigned long long)y);
  atomicAdd(s + 2, (unsigned long long)x);
  atomicAdd(s + 3, (unsigned long long)(long long)qc.x);
  atomicAdd(s + 4, (unsigned long long)(long long)qc.y);
  atomicAdd(s + 5, (unsigned long long)(long long)qc.z);
}

__global__ void __launch_bounds__(256)
slic_update_kernel(int n_seg, SlicWs ws) {
  const int im = blockIdx.y;
  __shared__ int s_changed;
  if (threadIdx.x == 0) s_changed = ws.flags[im * 4 + 0];
  __syncthreads();
  if (ws.flags[im * 4 + 1]) return;
  if (!s_changed) {  // assignment unchanged: the reference loop breaks before the centre update
    if (blockIdx.x == 0 && threadIdx.x == 0) ws.flags[im * 4 + 1] = 1;
    return;
  }
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k < n_seg) {
    const unsigned long long* s = ws.sums + ((size_t)im * n_seg + k) * 6;
    const long long cnt = (long long)s[0];
    if (cnt > 0) {  // an empty segment keeps its centre
      double* ce = ws.centre + ((size_t)im * n_seg + k) * 5;
      const double dc = (double)cnt;
      ce[0] = __ddiv_rn((double)(long long)s[1], dc);
      ce[1] = __ddiv_rn((double)(long long)s[2], dc);
      ce[2] = __ddiv_rn(__ddiv_rn((double)(long long)s[3], SLIC_Q), dc);
      ce[3] = __ddiv_rn(__ddiv_rn((double)(long long)s[4], SLIC_Q), dc);
      ce[4] = __ddiv_rn(__ddiv_rn((double)(long long)s[5], SLIC_Q), dc);
    }
  }
}

// the update kernels of all blocks read the flag before anyone clears it: cleared here, one
// launch later
__global__ void slic_clear_changed_kernel(SlicWs ws) {
  const int im = blockIdx.x;
  if (threadIdx.x == 0) ws.flags[im * 4 + 0] = 0;
}

// ---- connectivity -----------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int* parent, int x) {
  int p = parent[x];
  while (p != x) {
    x = p;
    p = parent[x];
  }
  return x;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a > b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(&parent[b], a);  // hang the larger root under the smaller
    if (old == b) return;
    b = old;
  }
}

__global__ void __launch_bounds__(256)
slic_uf_init_kernel(int64_t P, SlicWs ws) {
  const int im = blockIdx.y;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < P; p += (int64_t)gridDim.x * 256) {
    ws.parent[(size_t)im * P + p] = (int)p;
    ws.size[(size_t)im * P + p] = 0;
  }
}

__global__ void __launch_bounds__(256)
slic_uf_link_kernel(int H, int W, SlicWs ws) {
  const int im = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const size_t base = (size_t)im * H * W;
  const int p = y * W + x;
  const int k = ws.nearest[base + p];
  int* parent = ws.parent + base;
  if (x + 1 < W && ws.nearest[base + p + 1] == k) uf_union(parent, p, p + 1);
  if (y + 1 < H && ws.nearest[base + p + W] == k) uf_union(parent, p, p + W);
}

__global__ void __launch_bounds__(256)
slic_uf_flatten_kernel(int64_t P, SlicWs ws) {
  const int im = blockIdx.y;
  int* parent = ws.parent + (size_t)im * P;
  int* size = ws.size + (size_t)im * P;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < P; p += (int64_t)gridDim.x * 256) {
    const int r = uf_find(parent, (int)p);
    ws.target[(size_t)im * P + p] = r;  // scratch: root of every pixel
    atomicAdd(&size[r], 1);
  }
}

// after this kernel parent[p] = root of p (flattened), target[r] = merge target of root r
__global__ void __launch_bounds__(256)
slic_merge_target_kernel(int W, int64_t P, int min_size, SlicWs ws) {
  const int im = blockIdx.y;
  int* root = ws.target + (size_t)im * P;   // per-pixel roots written by flatten
  int* parent = ws.parent + (size_t)im * P;
  const int* size = ws.size + (size_t)im * P;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < P; p += (int64_t)gridDim.x * 256) {
    const int r = root[p];
    int t = r;
    if (r == (int)p && size[r] < min_size) {
      const int x = (int)(p % W);
      if (x > 0) t = root[p - 1];
      else if (p >= W) t = root[p - W];
    }
    // parent is dead as a union-find now: reuse it as "target of the root stored at p"
    parent[p] = (r == (int)p) ? t : -1 - r;   // non-roots remember their root (negative code)
  }
}

// pointer jumping over the roots: target chains end at a root that stays
__global__ void __launch_bounds__(256)
slic_jump_kernel(int64_t P, SlicWs ws) {
  const int im = blockIdx.y;
  int* parent = ws.parent + (size_t)im * P;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < P; p += (int64_t)gridDim.x * 256) {
    const int t = parent[p];
    if (t >= 0 && t != (int)p) {
      const int tt = parent[t];
      if (tt >= 0 && tt != t) parent[p] = tt;
    }
  }
}

// final root per pixel into target[]; count kept roots per scan block
__global__ void __launch_bounds__(256)
slic_final_root_kernel(int64_t P, int64_t n_blocks, SlicWs ws) {
  const int im = blockIdx.y;
  const int* parent = ws.parent + (size_t)im * P;
  int* fin = ws.target + (size_t)im * P;
  __shared__ int s_cnt;
  for (int64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < SCAN_PER_BLOCK; i += 256) {
      const int64_t p = b * SCAN_PER_BLOCK + i;
      if (p >= P) break;
      const int t = parent[p];
      const int r = t >= 0 ? (int)p : -1 - t;   // root of p
      fin[p] = parent[r];                       // its (jumped) target
      if (t == (int)p) ++mine;                  // a root that stays: one final id
    }
    atomicAdd(&s_cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0) ws.block_count[(size_t)im * n_blocks + b] = s_cnt;
    __syncthreads();
  }
}

// exclusive scan of the per-block counts (one block per image), total -> n_labels
__global__ void __launch_bounds__(1024)
slic_scan_blocks_kernel(int64_t n_blocks, SlicWs ws, int32_t* n_labels) {
  const int im = blockIdx.x;
  int* bc = ws.block_count + (size_t)im * n_blocks;
  __shared__ int s_part[1024];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < n_blocks; b0 += 1024) {
    const int64_t b = b0 + threadIdx.x;
    const int v = b < n_blocks ? bc[b] : 0;
    s_part[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      const int t = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0;
      __syncthreads();
      s_part[threadIdx.x] += t;
      __syncthreads();
    }
    if (b < n_blocks) bc[b] = s_carry + s_part[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += s_part[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) n_labels[im] = s_carry;
}

// id of a kept root = number of kept roots before it (raster order); stored in size[]
__global__ void __launch_bounds__(256)
slic_number_roots_kernel(int64_t P, int64_t n_blocks, SlicWs ws) {
  const int im = blockIdx.y;
  const int* parent = ws.parent + (size_t)im * P;
  int* id = ws.size + (size_t)im * P;
  __shared__ int s_warp[8];
  __shared__ int s_run;
  for (int64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
    if (threadIdx.x == 0) s_run = ws.block_count[(size_t)im * n_blocks + b];
    __syncthreads();
    for (int i0 = 0; i0 < SCAN_PER_BLOCK; i0 += 256) {
      const int64_t p = b * SCAN_PER_BLOCK + i0 + threadIdx.x;
      const bool keep = p < P && parent[p] == (int)p;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (lane_id() == 0) s_warp[warp_id()] = __popc(m);
      __syncthreads();
      int before = s_run;
      for (int w = 0; w < warp_id(); ++w) before += s_warp[w];
      if (keep) id[p] = before + __popc(m & ((1u << lane_id()) - 1u));
      __syncthreads();
      if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) tot += s_warp[w];
        s_run += tot;
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256)
slic_write_labels_kernel(int64_t P, SlicWs ws, int32_t* labels) {
  const int im = blockIdx.y;
  const int* fin = ws.target + (size_t)im * P;
  const int* id = ws.size + (size_t)im * P;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < P; p += (int64_t)gridDim.x * 256)
    labels[(size_t)im * P + p] = id[fin[p]];
}

// no connectivity pass: make every cluster id "its own root" so the same numbering code runs
__global__ void __launch_bounds__(256)
slic_raw_roots_kernel(int64_t P, int n_seg, SlicWs ws) {
  const int im = blockIdx.y;
  int* parent = ws.parent + (size_t)im * P;
  int* fin = ws.target + (size_t)im * P;
  // parent[p] = p marks "kept id p" for p < n_seg that occur; others are non-roots
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < P; p += (int64_t)gridDim.x * 256) {
    parent[p] = -1;
    fin[p] = ws.nearest[(size_t)im * P + p];
  }
}
__global__ void __launch_bounds__(256)
slic_raw_mark_kernel(int64_t P, SlicWs ws) {
  const int im = blockIdx.y;
  int* parent = ws.parent + (size_t)im * P;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < P; p += (int64_t)gridDim.x * 256) {
    const int k = ws.nearest[(size_t)im * P + p];
    parent[k] = k;   // benign race: every writer stores the same value
  }
}
__global__ void __launch_bounds__(256)
slic_raw_count_kernel(int64_t P, int64_t n_blocks, SlicWs ws) {
  const int im = blockIdx.y;
  const int* parent = ws.parent + (size_t)im * P;
  __shared__ int s_cnt;
  for (int64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < SCAN_PER_BLOCK; i += 256) {
      const int64_t p = b * SCAN_PER_BLOCK + i;
      if (p < P && parent[p] == (int)p) ++mine;
    }
    atomicAdd(&s_cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0) ws.block_count[(size_t)im * n_blocks + b] = s_cnt;
    __syncthreads();
  }
}

SlicGrid make_grid(int H, int W, int n_segments) {
  // skimage.util.regular_grid for a (1, H, W) volume (oracle: slic_grid)
  const double area = (double)H * (double)W / (double)n_segments;
  double sy, sx;
  const double step = sqrt(area);
  if (H < step) { sy = (double)H; sx = area / H; }
  else if (W < step) { sy = area / W; sx = (double)W; }
  else { sy = sx = step; }
  SlicGrid g;
  g.y0 = (int)floor(sy / 2.0);
  g.x0 = (int)floor(sx / 2.0);
  g.sy = (int)nearbyint(sy);
  g.sx = (int)nearbyint(sx);
  if (g.sy < 1) g.sy = 1;
  if (g.sx < 1) g.sx = 1;
  g.ny = (H - g.y0 + g.sy - 1) / g.sy;
  g.nx = (W - g.x0 + g.sx - 1) / g.sx;
  if (g.ny < 1) g.ny = 1;
  if (g.nx < 1) g.nx = 1;
  return g;
}

}  // namespace
}  // namespace spalign

using namespace spalign;

extern "C" int spalign_slic_segments(int H, int W, int n_segments) {
  if (H <= 0 || W <= 0 || n_segments <= 0) return 0;
  const SlicGrid g = make_grid(H, W, n_segments);
  return g.ny * g.nx;
}

extern "C" size_t spalign_slic_workspace_bytes(int n_img, int H, int W, int n_segments) {
  SlicWs ws;
  const SlicGrid g = make_grid(H, W, n_segments);
  return slic_carve(ws, nullptr, n_img, (int64_t)H * W, g.ny * g.nx) + 256;
}

extern "C" int spalign_slic(const float* images, int n_img, int H, int W, int n_segments,
                            double compactness, int max_iter, int convert2lab,
                            int enforce_connectivity, double min_size_factor, int32_t* labels,
                            int32_t* n_labels, void* workspace, size_t ws_bytes,
                            spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(images && labels && n_labels && workspace, "slic: NULL argument");
  SPALIGN_REQUIRE(n_img > 0 && n_img <= 65535 && H > 0 && W > 0 && n_segments > 0 && max_iter >= 0 &&
                      compactness > 0.0 && (int64_t)H * W < 0x7fffffffLL,
                  "slic: bad arguments");
  const int64_t P = (int64_t)H * W;
  const SlicGrid g = make_grid(H, W, n_segments);
  const int n_seg = g.ny * g.nx;
  size_t need = spalign_slic_workspace_bytes(n_img, H, W, n_segments);
  if (ws_bytes < need) {
    set_error("slic: workspace %zu < %zu bytes", ws_bytes, need);
    return SPALIGN_E_WORKSPACE;
  }
  SlicWs ws;
  void* aligned = reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  slic_carve(ws, aligned, n_img, P, n_seg);
  const int gx1 = (int)((P + 256 * 8 - 1) / (256 * 8)) < 4 * kNumSMs ? (int)((P + 256 * 8 - 1) / (256 * 8))
                                                                   : 4 * kNumSMs;
  const dim3 g1(gx1 < 1 ? 1 : gx1, n_img);
  const dim3 g2((W + 31) / 32, (H + 7) / 8, n_img);
  const dim3 gs((n_seg + 255) / 256, n_img);
  const double step = (double)(g.sy > g.sx ? g.sy : g.sx);
  const double wsp = 1.0 / (step * step);
  slic_quantise_kernel<<<g1, 256, 0, stream>>>(images, P, 1.0 / compactness, convert2lab, ws);
  slic_seed_kernel<<<gs, 256, 0, stream>>>(g, W, P, n_seg, ws);
  slic_zero_nearest_kernel<<<4 * kNumSMs, 256, 0, stream>>>((int64_t)n_img * P, ws);
  for (int it = 0; it < max_iter; ++it) {
    slic_clear_changed_kernel<<<n_img, 32, 0, stream>>>(ws);
    slic_assign_kernel<<<g2, 256, 0, stream>>>(g, H, W, n_seg, wsp, ws);
    slic_zero_sums_kernel<<<gs, 256, 0, stream>>>(n_seg, ws);
    slic_accumulate_kernel<<<g2, 256, 0, stream>>>(H, W, n_seg, ws);
    slic_update_kernel<<<gs, 256, 0, stream>>>(n_seg, ws);
  }
  const int64_t n_blocks = (P + SCAN_PER_BLOCK - 1) / SCAN_PER_BLOCK;
  const dim3 gb((unsigned)(n_blocks < 4 * kNumSMs ? n_blocks : 4 * kNumSMs), n_img);
  if (enforce_connectivity) {
    const int min_size = (int)(min_size_factor * (double)P / (double)n_segments);
    slic_uf_init_kernel<<<g1, 256, 0, stream>>>(P, ws);
    slic_uf_link_kernel<<<g2, 256, 0, stream>>>(H, W, ws);
    slic_uf_flatten_kernel<<<g1, 256, 0, stream>>>(P, ws);
    slic_merge_target_kernel<<<g1, 256, 0, stream>>>(W, P, min_size, ws);
    for (int j = 0; j < 24; ++j) slic_jump_kernel<<<g1, 256, 0, stream>>>(P, ws);
    slic_final_root_kernel<<<gb, 256, 0, stream>>>(P, n_blocks, ws);
  } else {
    slic_raw_roots_kernel<<<g1, 256, 0, stream>>>(P, n_seg, ws);
    slic_raw_mark_kernel<<<g1, 256, 0, stream>>>(P, ws);
    slic_raw_count_kernel<<<gb, 256, 0, stream>>>(P, n_blocks, ws);
  }
  slic_scan_blocks_kernel<<<n_img, 1024, 0, stream>>>(n_blocks, ws, n_labels);
  slic_number_roots_kernel<<<gb, 256, 0, stream>>>(P, n_blocks, ws);
  slic_write_labels_kernel<<<g1, 256, 0, stream>>>(P, ws, labels);
  return check_launch("slic");
}
