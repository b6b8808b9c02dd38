// K0/K1: SLIC label map -> CSR overlap (pixel-count) matrix + per-superpixel statistics.
//
// Replaces the S full-image boolean masks of the reference (batch_spalign_kmeans.py:226-233,
// create_prior :111-129, superpixel_overlaps.py:365-369) with one integer pass over the
// label map.  Pipeline (all on one stream, no host sync):
//
//   init      zero counters / accumulators
//   emit      one thread per feature cell: distinct labels of the cell -> (row, cell, count,
//             sum_y, sum_x, prior) pairs, staged in shared memory, flushed with one global
//             allocation per block; integer atomics for row lengths and centroid sums
//   scan      row lengths -> indptr (two-level), heavy-row list
//   scatter   pairs -> their row segment (arbitrary order inside the row)
//   rowsort   per row: rank-sort by cell id (a warp per row; rows longer than 256 cells go
//             through a dense scatter/compact path), area and prior summed in column order
//
// The only floating-point reduction (prior) is summed in sorted column order with a fixed
// tree, so the whole result is bit-reproducible.
#include <stdlib.h>

#include "common.cuh"

namespace spalign {
namespace {

constexpr int EMIT_THREADS = 128;
#ifndef EMIT_MIN_BLOCKS
#define EMIT_MIN_BLOCKS 4  // 128 registers, no spills: four resident blocks per SM
#endif
constexpr int STAGE_CAP = 768;      // pairs staged per block
constexpr int WARP_TIER_MAX = 256;  // rows up to this many cells are sorted by one warp
constexpr int SCAN_TILE = 2048;
constexpr int HEAVY_SLOTS = 32;
constexpr int HEAVY_THREADS = 256;

struct OverlapWs {
  int* zero_begin;   // start of the region that init zeroes
  size_t zero_ints;  // its length in ints
  int* row_nnz;      // [R]
  int* cursor;       // [R]
  int* pair_count;   // [n_img]
  int* heavy_count;  // [1]
  int* tile_sum;     // [n_tiles]
  int* heavy_rows;   // [heavy_cap]
  int heavy_cap;
  int* t_row;        // [cap] temp pairs (per-image regions of cap_img)
  int* t_col;
  int* t_cnt;
  double* t_prior;   // reused as sorted-prior scratch by rowsort
  int* u_col;        // [cap] row-segmented, unsorted
  int* u_cnt;
  double* u_prior;
  int* d_cnt;        // [HEAVY_SLOTS * ncell] dense scratch of the heavy tier
  double* d_prior;
};

size_t carve(OverlapWs& ws, void* base, int n_img, int ncell, int64_t R, int64_t cap) {
  Carver c(base);
  int n_tiles = (int)((R + SCAN_TILE - 1) / SCAN_TILE);
  ws.heavy_cap = (int)(cap / (WARP_TIER_MAX + 1) + 1);
  ws.row_nnz = c.take<int>(R);
  ws.zero_begin = ws.row_nnz;
  ws.cursor = c.take<int>(R);
  ws.pair_count = c.take<int>(n_img);
  ws.heavy_count = c.take<int>(1);
  size_t zero_end = c.off;
  ws.zero_ints = (zero_end - 0) / sizeof(int);
  ws.tile_sum = c.take<int>(n_tiles);
  ws.heavy_rows = c.take<int>(ws.heavy_cap);
  ws.t_row = c.take<int>(cap);
  ws.t_col = c.take<int>(cap);
  ws.t_cnt = c.take<int>(cap);
  ws.t_prior = c.take<double>(cap);
  ws.u_col = c.take<int>(cap);
  ws.u_cnt = c.take<int>(cap);
  ws.u_prior = c.take<double>(cap);
  ws.d_cnt = c.take<int>((size_t)HEAVY_SLOTS * ncell);
  ws.d_prior = c.take<double>((size_t)HEAVY_SLOTS * ncell);
  return c.used();
}

// ------------------------------------------------------------------------------------------
__global__ void init_kernel(int* zero_begin, size_t zero_ints, int64_t* sum_y, int64_t* sum_x,
                            int64_t R, int64_t* nnz_flags) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = i; k < zero_ints; k += stride) zero_begin[k] = 0;
  for (size_t k = i; k < (size_t)R; k += stride) {
    sum_y[k] = 0;
    sum_x[k] = 0;
  }
  if (i < 4) nnz_flags[i] = 0;
}

// ------------------------------------------------------------------------------------------
struct PairStage {
  int row[STAGE_CAP];
  int col[STAGE_CAP];
  int cnt[STAGE_CAP];
  long long sy[STAGE_CAP];
  long long sx[STAGE_CAP];
  double prior[STAGE_CAP];
  int count;
  int base;
};

__device__ __forceinline__ void publish_pair(const OverlapWs& ws, int img, int cap_img, int dst,
                                             int row, int col, int cnt, long long sy,
                                             long long sx, double prior, int64_t* sum_y,
                                             int64_t* sum_x, int64_t* nnz_flags) {
  atomicAdd(&ws.row_nnz[row], 1);
  atomicAdd(reinterpret_cast<unsigned long long*>(&sum_y[row]), (unsigned long long)sy);
  atomicAdd(reinterpret_cast<unsigned long long*>(&sum_x[row]), (unsigned long long)sx);
  if (dst < cap_img) {
    size_t t = (size_t)img * cap_img + dst;
    ws.t_row[t] = row;
    ws.t_col[t] = col;
    ws.t_cnt[t] = cnt;
    ws.t_prior[t] = prior;
  } else {
    atomicOr(reinterpret_cast<unsigned long long*>(&nnz_flags[1]),
             (unsigned long long)SPALIGN_F_NNZ_OVERFLOW);
  }
}

__device__ __forceinline__ void stage_pair(PairStage& st, const OverlapWs& ws, int img,
                                           int cap_img, int row, int col, int cnt, long long sy,
                                           long long sx, double prior, int64_t* sum_y,
                                           int64_t* sum_x, int64_t* nnz_flags) {
  int pos = atomicAdd(&st.count, 1);
  if (pos < STAGE_CAP) {
    st.row[pos] = row;
    st.col[pos] = col;
    st.cnt[pos] = cnt;
    st.sy[pos] = sy;
    st.sx[pos] = sx;
    st.prior[pos] = prior;
  } else {  // stage full (pathological label maps): go straight to global memory
    int dst = atomicAdd(&ws.pair_count[img], 1);
    publish_pair(ws, img, cap_img, dst, row, col, cnt, sy, sx, prior, sum_y, sum_x, nnz_flags);
  }
}

__device__ __forceinline__ void flush_stage(PairStage& st, const OverlapWs& ws, int img,
                                            int cap_img, int64_t* sum_y, int64_t* sum_x,
                                            int64_t* nnz_flags) {
  __syncthreads();
  int n = min(st.count, STAGE_CAP);
  if (threadIdx.x == 0) {
    st.base = n ? atomicAdd(&ws.pair_count[img], n) : 0;
  }
  __syncthreads();
  int base = st.base;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    publish_pair(ws, img, cap_img, base + i, st.row[i], st.col[i], st.cnt[i], st.sy[i],
                 st.sx[i], st.prior[i], sum_y, sum_x, nnz_flags);
}

// Fast path: 8x8-pixel cells (DRN stride 8), one thread owns one cell in registers.
template <typename LabelT>
__global__ void __launch_bounds__(EMIT_THREADS, EMIT_MIN_BLOCKS)
emit_s8_kernel(const LabelT* __restrict__ labels, int H, int W, int fh, int fw,
               const int64_t* __restrict__ sp_off, const double* __restrict__ gy,
               const double* __restrict__ gx, int cap_img, OverlapWs ws, int64_t* sum_y,
               int64_t* sum_x, int64_t* nnz_flags) {
  __shared__ PairStage st;
  const int img = blockIdx.y;
  const int ncell = fh * fw;
  const int c = blockIdx.x * EMIT_THREADS + threadIdx.x;
  if (threadIdx.x == 0) st.count = 0;
  __syncthreads();
  const int64_t row0 = sp_off[img];
  const int n_sp = (int)(sp_off[img + 1] - row0);
  if (c < ncell) {
    const int cy = c / fw, cx = c - cy * fw;
    const LabelT* p = labels + ((size_t)img * H + (size_t)cy * 8) * W + (size_t)cx * 8;
    int v[64];
    bool bad = false;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (sizeof(LabelT) == 4) {
        int4 a = ld_stream_int4(p + (size_t)r * W);
        int4 b = ld_stream_int4(p + (size_t)r * W + 4);
        v[r * 8 + 0] = a.x; v[r * 8 + 1] = a.y; v[r * 8 + 2] = a.z; v[r * 8 + 3] = a.w;
        v[r * 8 + 4] = b.x; v[r * 8 + 5] = b.y; v[r * 8 + 6] = b.z; v[r * 8 + 7] = b.w;
      } else {
        const longlong2* q = reinterpret_cast<const longlong2*>(p + (size_t)r * W);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          longlong2 a = __ldg(q + h);
          v[r * 8 + 2 * h] = (a.x >= 0 && a.x < n_sp) ? (int)a.x : -1;
          v[r * 8 + 2 * h + 1] = (a.y >= 0 && a.y < n_sp) ? (int)a.y : -1;
        }
      }
    }
    unsigned long long remaining = ~0ull;  // label validity is checked once per distinct label
    double gyl[8], gxl[8];
    const bool have_prior = gy != nullptr;
    double cell_prior = 0.0;  // prior of the whole cell: sum_r gy[r] * (sum_j gx[j])
    if (have_prior) {
      double gx8 = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        gyl[k] = gy[cy * 8 + k];
        gxl[k] = gx[cx * 8 + k];
        gx8 = __dadd_rn(gx8, gxl[k]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) cell_prior = __fma_rn(gyl[k], gx8, cell_prior);
    }
    double emitted_prior = 0.0;  // prior already attributed to earlier labels of this cell
    bool can_complement = true;
    bool first = true;
    while (remaining) {
      // next label: the first still-unprocessed pixel among a few fixed positions (corners,
      // edge and centre pixels -- static register indices), else the first unprocessed pixel
      int L = 0;
      bool found = false;
#define SPALIGN_CAND(p)                                  \
  if (!found && ((remaining >> (p)) & 1ull)) {            \
    L = v[p];                                             \
    found = true;                                         \
  }
      SPALIGN_CAND(0) SPALIGN_CAND(7) SPALIGN_CAND(56) SPALIGN_CAND(63) SPALIGN_CAND(3)
      SPALIGN_CAND(60) SPALIGN_CAND(24) SPALIGN_CAND(31) SPALIGN_CAND(32) SPALIGN_CAND(39)
      SPALIGN_CAND(27) SPALIGN_CAND(36)
#undef SPALIGN_CAND
      if (!found) {
        const int i = __ffsll((long long)remaining) - 1;
        L = v[0];
#pragma unroll
        for (int j = 1; j < 64; ++j) L = (i == j) ? v[j] : L;
      }
      unsigned lo = 0, hi = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        lo |= (v[j] == L) ? (1u << j) : 0u;
        hi |= (v[32 + j] == L) ? (1u << j) : 0u;
      }
      const unsigned long long m = (unsigned long long)lo | ((unsigned long long)hi << 32);
      remaining &= ~m;
      if ((unsigned)L >= (unsigned)n_sp) {  // label outside [0, n_sp): flag it, emit nothing
        bad = true;
        can_complement = false;
        first = false;
        continue;
      }
      const int cnt = __popc(lo) + __popc(hi);
      // sum of the row (column) indices of the set bits: bit i sits in row i >> 3, column
      // i & 7, so each index bit contributes 2^b * popc(m & {bits whose index has bit b set})
      const int sxl = __popcll(m & 0xaaaaaaaaaaaaaaaaull) + 2 * __popcll(m & 0xccccccccccccccccull) +
                      4 * __popcll(m & 0xf0f0f0f0f0f0f0f0ull);
      const int syl = __popcll(m & 0xff00ff00ff00ff00ull) + 2 * __popcll(m & 0xffff0000ffff0000ull) +
                      4 * __popcll(m & 0xffffffff00000000ull);
      double pr = 0.0;
      if (have_prior) {
        if (remaining == 0ull && can_complement) {
          // last label of the cell: whole-cell prior minus what the other labels took
          // (a cell with a single label gets cell_prior itself)
          pr = first ? cell_prior : __dadd_rn(cell_prior, -emitted_prior);
        } else {
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const unsigned b = (unsigned)(m >> (8 * r)) & 0xffu;
            double rs = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) rs = __dadd_rn(rs, ((b >> j) & 1u) ? gxl[j] : 0.0);
            pr = __fma_rn(gyl[r], rs, pr);
          }
          emitted_prior = __dadd_rn(emitted_prior, pr);
        }
      }
      first = false;
      stage_pair(st, ws, img, cap_img, (int)(row0 + L), c, cnt,
                 (long long)cnt * (cy * 8) + syl, (long long)cnt * (cx * 8) + sxl, pr, sum_y,
                 sum_x, nnz_flags);
    }
    if (bad)
      atomicOr(reinterpret_cast<unsigned long long*>(&nnz_flags[1]),
               (unsigned long long)SPALIGN_F_LABEL_RANGE);
  }
  flush_stage(st, ws, img, cap_img, sum_y, sum_x, nnz_flags);
}

// General path: any H, W, fh, fw.  One thread per cell walks the cell's pixel rectangle once
// per distinct label (ascending label order).
template <typename LabelT>
__global__ void __launch_bounds__(EMIT_THREADS)
emit_generic_kernel(const LabelT* __restrict__ labels, int H, int W, int fh, int fw,
                    const int64_t* __restrict__ sp_off, const double* __restrict__ gy,
                    const double* __restrict__ gx, int cap_img, OverlapWs ws, int64_t* sum_y,
                    int64_t* sum_x, int64_t* nnz_flags) {
  __shared__ PairStage st;
  const int img = blockIdx.y;
  const int ncell = fh * fw;
  const int c = blockIdx.x * EMIT_THREADS + threadIdx.x;
  if (threadIdx.x == 0) st.count = 0;
  __syncthreads();
  const int64_t row0 = sp_off[img];
  const long long n_sp = sp_off[img + 1] - row0;
  if (c < ncell) {
    const int cy = c / fw, cx = c - cy * fw;
    // rows y with floor(y*fh/H) == cy  <=>  y in [ceil(cy*H/fh), ceil((cy+1)*H/fh))
    const int y0 = (int)(((long long)cy * H + fh - 1) / fh);
    const int y1 = (int)(((long long)(cy + 1) * H + fh - 1) / fh);
    const int x0 = (int)(((long long)cx * W + fw - 1) / fw);
    const int x1 = (int)(((long long)(cx + 1) * W + fw - 1) / fw);
    const LabelT* base = labels + (size_t)img * H * W;
    const bool have_prior = gy != nullptr;
    long long lo = -1;
    bool bad = false;
    while (true) {
      long long cur = 0x7fffffffffffffffLL;
      int cnt = 0;
      long long sy = 0, sx = 0;
      double pr = 0.0;
      for (int y = y0; y < y1; ++y) {
        const LabelT* rowp = base + (size_t)y * W;
        for (int x = x0; x < x1; ++x) {
          const long long q = (long long)rowp[x];
          if (q < 0 || q >= n_sp) {
            bad = true;
            continue;
          }
          if (q > lo) {
            if (q < cur) {
              cur = q;
              cnt = 0;
              sy = 0;
              sx = 0;
              pr = 0.0;
            }
            if (q == cur) {
              ++cnt;
              sy += y;
              sx += x;
              if (have_prior) pr = __dadd_rn(pr, __dmul_rn(gy[y], gx[x]));
            }
          }
        }
      }
      if (cur == 0x7fffffffffffffffLL) break;
      stage_pair(st, ws, img, cap_img, (int)(row0 + cur), c, cnt, sy, sx, pr, sum_y, sum_x,
                 nnz_flags);
      lo = cur;
    }
    if (bad)
      atomicOr(reinterpret_cast<unsigned long long*>(&nnz_flags[1]),
               (unsigned long long)SPALIGN_F_LABEL_RANGE);
  }
  flush_stage(st, ws, img, cap_img, sum_y, sum_x, nnz_flags);
}

// Bilinear-weight pairs (dense pooling of notebooks/Superpixel_Align.ipynb cell 4):
// W[s, c] = sum over the pixels p of superpixel s of the bilinear weight of cell c at p, with
// the corner-aligned sampling of chainer.functions.resize_images.  One thread per cell walks
// the cell's support window (the pixel rows/columns whose lower neighbour is cy-1 or cy) once
// to find the next label and once to accumulate it, labels in ascending order; sums are
// float64 in fixed row-major order.  `cnt` carries the number of contributing pixels.
struct BilinearAxes {
  const int* iy0;      // [H] lower neighbour row of pixel row y (0..fh-2)
  const double* wy0;   // [H] weight of iy0[y]
  const double* wy1;   // [H] weight of iy0[y] + 1
  const int* ystart;   // [fh+1] first y with iy0[y] >= cy
  const int* ix0;      // [W]
  const double* wx0;
  const double* wx1;
  const int* xstart;   // [fw+1]
};

template <typename LabelT>
__global__ void __launch_bounds__(EMIT_THREADS)
emit_bilinear_kernel(const LabelT* __restrict__ labels, int H, int W, int fh, int fw,
                     const int64_t* __restrict__ sp_off, BilinearAxes ax, int cap_img,
                     OverlapWs ws, int64_t* sum_y, int64_t* sum_x, int64_t* nnz_flags) {
  __shared__ PairStage st;
  const int img = blockIdx.y;
  const int ncell = fh * fw;
  const int c = blockIdx.x * EMIT_THREADS + threadIdx.x;
  if (threadIdx.x == 0) st.count = 0;
  __syncthreads();
  const int64_t row0 = sp_off[img];
  const long long n_sp = sp_off[img + 1] - row0;
  if (c < ncell) {
    const int cy = c / fw, cx = c - cy * fw;
    const int y0 = ax.ystart[cy > 0 ? cy - 1 : 0], y1 = ax.ystart[cy + 1];
    const int x0 = ax.xstart[cx > 0 ? cx - 1 : 0], x1 = ax.xstart[cx + 1];
    const LabelT* base = labels + (size_t)img * H * W;
    long long lo = -1;
    bool bad = false;
    while (true) {
      long long cur = 0x7fffffffffffffffLL;
      for (int y = y0; y < y1; ++y) {
        const LabelT* rowp = base + (size_t)y * W;
        for (int x = x0; x < x1; ++x) {
          const long long q = (long long)rowp[x];
          if (q < 0 || q >= n_sp) {
            bad = true;
          } else if (q > lo && q < cur) {
            cur = q;
          }
        }
      }
      if (cur == 0x7fffffffffffffffLL) break;
      int cnt = 0;
      double acc = 0.0;
      for (int y = y0; y < y1; ++y) {
        const LabelT* rowp = base + (size_t)y * W;
        const double hy = (ax.iy0[y] == cy) ? ax.wy0[y] : ax.wy1[y];
        double ra = 0.0;
        for (int x = x0; x < x1; ++x) {
          if ((long long)rowp[x] == cur) {
            ra = __dadd_rn(ra, (ax.ix0[x] == cx) ? ax.wx0[x] : ax.wx1[x]);
            ++cnt;
          }
        }
        acc = __fma_rn(hy, ra, acc);
      }
      stage_pair(st, ws, img, cap_img, (int)(row0 + cur), c, cnt, 0, 0, acc, sum_y, sum_x,
                 nnz_flags);
      lo = cur;
    }
    if (bad)
      atomicOr(reinterpret_cast<unsigned long long*>(&nnz_flags[1]),
               (unsigned long long)SPALIGN_F_LABEL_RANGE);
  }
  flush_stage(st, ws, img, cap_img, sum_y, sum_x, nnz_flags);
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane_id() >= d) v += t;
  }
  return v;
}

// exclusive scan of one int per thread across a 256-thread block; returns exclusive prefix,
// *total receives the block sum
__device__ __forceinline__ int block_excl_scan_256(int v, int* total) {
  __shared__ int wsum[8];
  __shared__ int wtot;
  int inc = warp_incl_scan(v);
  if (lane_id() == 31) wsum[warp_id()] = inc;
  __syncthreads();
  if (warp_id() == 0) {
    int w = lane_id() < 8 ? wsum[lane_id()] : 0;
    int winc = warp_incl_scan(w);
    if (lane_id() < 8) wsum[lane_id()] = winc - w;
    if (lane_id() == 7) wtot = winc;
  }
  __syncthreads();
  int excl = inc - v + wsum[warp_id()];
  *total = wtot;
  __syncthreads();
  return excl;
}

__global__ void __launch_bounds__(256)
scan_tile_sums_kernel(const int* __restrict__ row_nnz, int64_t R, int* tile_sum) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_TILE / 256; ++k) {
    int64_t idx = base + k * 256 + threadIdx.x;
    if (idx < R) s += row_nnz[idx];
  }
  int total;
  block_excl_scan_256(s, &total);
  if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256)
scan_finish_kernel(const int* __restrict__ row_nnz, int64_t R, const int* __restrict__ tile_sum,
                   int n_tiles, int* indptr, int64_t* nnz_flags, int64_t nnz_cap,
                   int* heavy_rows, int* heavy_count, int heavy_cap,
                   const int* __restrict__ pair_count, int n_img, int heavy_thr) {
  if (blockIdx.x == 0) {  // high-water mark of pairs per image (sizes nnz_cap after an overflow)
    int mx = 0;
    for (int i = threadIdx.x; i < n_img; i += 256) mx = max(mx, pair_count[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if (lane_id() == 0 && mx > 0) atomicMax(reinterpret_cast<long long*>(&nnz_flags[2]), (long long)mx);
  }
  __shared__ long long s_prefix;
  // prefix of the tiles before this one
  long long pre = 0;
  for (int i = threadIdx.x; i < (int)blockIdx.x; i += 256) pre += tile_sum[i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, d);
  __shared__ long long wpre[8];
  if (lane_id() == 0) wpre[warp_id()] = pre;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int i = 0; i < 8; ++i) t += wpre[i];
    s_prefix = t;
  }
  __syncthreads();
  const long long prefix = s_prefix;
  constexpr int PER = SCAN_TILE / 256;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * PER;
  int vals[PER];
  int tsum = 0;
  bool empty = false;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    int64_t idx = base + k;
    vals[k] = idx < R ? row_nnz[idx] : 0;
    if (idx < R && vals[k] == 0) empty = true;
    tsum += vals[k];
  }
  int total;
  int excl = block_excl_scan_256(tsum, &total);
  long long run = prefix + excl;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    int64_t idx = base + k;
    if (idx < R) {
      indptr[idx] = (int)run;
      if (vals[k] > heavy_thr) {
        int h = atomicAdd(heavy_count, 1);
        if (h < heavy_cap) heavy_rows[h] = (int)idx;
      }
    }
    run += vals[k];
  }
  if (empty)
    atomicOr(reinterpret_cast<unsigned long long*>(&nnz_flags[1]),
             (unsigned long long)SPALIGN_F_EMPTY_ROW);
  if ((int)blockIdx.x == n_tiles - 1 && threadIdx.x == 0) {
    long long nnz = prefix + total;
    nnz_flags[0] = nnz;
    indptr[R] = (int)(nnz > 0x7fffffffLL ? 0x7fffffffLL : nnz);
    if (nnz > nnz_cap)
      atomicOr(reinterpret_cast<unsigned long long*>(&nnz_flags[1]),
               (unsigned long long)SPALIGN_F_NNZ_OVERFLOW);
  }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scatter_kernel(OverlapWs ws, int cap_img, const int* __restrict__ indptr, int64_t nnz_cap,
               const int64_t* __restrict__ nnz_flags) {
  if (nnz_flags[1] & SPALIGN_F_NNZ_OVERFLOW) return;
  const int img = blockIdx.y;
  const int n = min(ws.pair_count[img], cap_img);
  for (int t = blockIdx.x * 256 + threadIdx.x; t < n; t += gridDim.x * 256) {
    const size_t src = (size_t)img * cap_img + t;
    const int r = ws.t_row[src];
    const int pos = indptr[r] + atomicAdd(&ws.cursor[r], 1);
    if (pos < nnz_cap) {
      ws.u_col[pos] = ws.t_col[src];
      ws.u_cnt[pos] = ws.t_cnt[src];
      ws.u_prior[pos] = ws.t_prior[src];
    }
  }
}

// one warp per row with <= WARP_TIER_MAX cells: rank sort by cell id
__global__ void __launch_bounds__(256)
rowsort_warp_kernel(OverlapWs ws, int* indptr, int64_t R, int* indices,
                    int* counts, int* area, double* sum_prior, double* wvals,
                    const int64_t* __restrict__ nnz_flags, int ncell_hint) {
  const int64_t r = (int64_t)blockIdx.x * 8 + warp_id();
  if (r >= R) return;
  const int lane = lane_id();
  if (nnz_flags[1] & SPALIGN_F_NNZ_OVERFLOW) {
    // capacity exceeded: leave an EMPTY matrix behind (no row reaches into unwritten storage),
    // the flag tells the caller
    if (lane == 0) {
      indptr[r] = 0;
      if (r == R - 1) indptr[R] = 0;
      area[r] = 0;
      if (sum_prior != nullptr) sum_prior[r] = 0.0;
    }
    return;
  }
  const int base = indptr[r];
  const int L = indptr[r + 1] - base;
  if (L > WARP_TIER_MAX) return;
  int a_sum = 0;
  double* sorted_prior = ws.t_prior;  // pairs are dead after scatter
  if (L <= 64 && ncell_hint <= (1 << 24)) {
    // the common case: bitonic sort of (cell << 7 | position) keys, two per lane (elements
    // lane and lane + 32), 21 compare-exchange stages; the payload is gathered afterwards
    const unsigned PAD = 0xffffffffu;
    unsigned k0 = lane < L ? ((unsigned)ws.u_col[base + lane] << 7) | (unsigned)lane : PAD;
    unsigned k1 = lane + 32 < L
                      ? ((unsigned)ws.u_col[base + lane + 32] << 7) | (unsigned)(lane + 32)
                      : PAD;
    if (L > 32) {
#pragma unroll
      for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
          if (j == 32) {  // partner in the same lane (only in the last merge: ascending)
            const unsigned lo = min(k0, k1), hi = max(k0, k1);
            k0 = lo;
            k1 = hi;
          } else {
            const unsigned p0 = __shfl_xor_sync(0xffffffffu, k0, j);
            const unsigned p1 = __shfl_xor_sync(0xffffffffu, k1, j);
            const bool lower = (lane & j) == 0;
            // element indices lane and lane + 32: direction bit from (index & k)
            const bool up0 = (lane & k) == 0, up1 = ((lane + 32) & k) == 0;
            k0 = (lower == up0) ? min(k0, p0) : max(k0, p0);
            k1 = (lower == up1) ? min(k1, p1) : max(k1, p1);
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
          const unsigned p0 = __shfl_xor_sync(0xffffffffu, k0, j);
          const bool lower = (lane & j) == 0, up0 = (lane & k) == 0 || k == 32;
          k0 = (lower == up0) ? min(k0, p0) : max(k0, p0);
        }
      }
    }
    double pv0 = 0.0, pv1 = 0.0;
    if (lane < L) {
      const int src = (int)(k0 & 127u);
      const int cn = ws.u_cnt[base + src];
      pv0 = ws.u_prior[base + src];
      indices[base + lane] = (int)(k0 >> 7);
      counts[base + lane] = cn;
      sorted_prior[base + lane] = pv0;
      if (wvals != nullptr) wvals[base + lane] = pv0;
      a_sum += cn;
    }
    if (lane + 32 < L) {
      const int src = (int)(k1 & 127u);
      const int cn = ws.u_cnt[base + src];
      pv1 = ws.u_prior[base + src];
      indices[base + lane + 32] = (int)(k1 >> 7);
      counts[base + lane + 32] = cn;
      sorted_prior[base + lane + 32] = pv1;
      if (wvals != nullptr) wvals[base + lane + 32] = pv1;
      a_sum += cn;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a_sum += __shfl_xor_sync(0xffffffffu, a_sum, d);
    if (lane == 0) area[r] = a_sum;
    if (sum_prior != nullptr) {  // same order as the general path: lane, lane + 32, then the tree
      double ps = __dadd_rn(0.0, pv0);
      if (lane + 32 < L) ps = __dadd_rn(ps, pv1);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) ps = __dadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, d));
      if (lane == 0) sum_prior[r] = ps;
    }
    return;
  }
  for (int a = 0; a < L; a += 32) {
    const int e = a + lane;
    const bool valid = e < L;
    const int myc = valid ? ws.u_col[base + e] : 0x7fffffff;
    int rank = 0;
    for (int b = 0; b < L; b += 32) {
      const int oc = (b + lane < L) ? ws.u_col[base + b + lane] : 0x7fffffff;
      const int nb = min(32, L - b);
      for (int j = 0; j < nb; ++j) {
        const int o = __shfl_sync(0xffffffffu, oc, j);
        rank += (o < myc) ? 1 : 0;
      }
    }
    if (valid) {
      const int cn = ws.u_cnt[base + e];
      const double pv = ws.u_prior[base + e];
      indices[base + rank] = myc;
      counts[base + rank] = cn;
      sorted_prior[base + rank] = pv;
      if (wvals != nullptr) wvals[base + rank] = pv;
      a_sum += cn;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) a_sum += __shfl_xor_sync(0xffffffffu, a_sum, d);
  if (lane == 0) area[r] = a_sum;
  if (sum_prior != nullptr) {
    __syncwarp();
    double ps = 0.0;
    for (int e = lane; e < L; e += 32) ps = __dadd_rn(ps, sorted_prior[base + e]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ps = __dadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, d));
    if (lane == 0) sum_prior[r] = ps;
  }
}

// rows longer than WARP_TIER_MAX: scatter into a dense per-cell array, compact in order
__global__ void __launch_bounds__(HEAVY_THREADS)
rowsort_heavy_kernel(OverlapWs ws, const int* __restrict__ indptr, int ncell, int* indices,
                     int* counts, int* area, double* sum_prior, double* wvals,
                     const int64_t* __restrict__ nnz_flags) {
  if (nnz_flags[1] & SPALIGN_F_NNZ_OVERFLOW) return;
  const int n_heavy = min(*ws.heavy_count, ws.heavy_cap);
  int* d_cnt = ws.d_cnt + (size_t)blockIdx.x * ncell;
  double* d_prior = ws.d_prior + (size_t)blockIdx.x * ncell;
  __shared__ double red_p[HEAVY_THREADS];
  __shared__ int red_a[HEAVY_THREADS];
  for (int h = blockIdx.x; h < n_heavy; h += gridDim.x) {
    const int r = ws.heavy_rows[h];
    const int base = indptr[r];
    const int L = indptr[r + 1] - base;
    for (int c = threadIdx.x; c < ncell; c += HEAVY_THREADS) d_cnt[c] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < L; e += HEAVY_THREADS) {
      const int c = ws.u_col[base + e];
      d_cnt[c] = ws.u_cnt[base + e];
      d_prior[c] = ws.u_prior[base + e];
    }
    __syncthreads();
    int running = 0;
    int a_sum = 0;
    double p_sum = 0.0;
    for (int c0 = 0; c0 < ncell; c0 += HEAVY_THREADS) {
      const int c = c0 + threadIdx.x;
      const int cn = c < ncell ? d_cnt[c] : 0;
      const int flag = cn > 0 ? 1 : 0;
      int total;
      const int excl = block_excl_scan_256(flag, &total);
      if (flag) {
        indices[base + running + excl] = c;
        counts[base + running + excl] = cn;
        if (wvals != nullptr) wvals[base + running + excl] = d_prior[c];
        a_sum += cn;
        p_sum = __dadd_rn(p_sum, d_prior[c]);
      }
      running += total;
    }
    red_p[threadIdx.x] = p_sum;
    red_a[threadIdx.x] = a_sum;
    __syncthreads();
    for (int s = HEAVY_THREADS / 2; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) {
        red_p[threadIdx.x] = __dadd_rn(red_p[threadIdx.x], red_p[threadIdx.x + s]);
        red_a[threadIdx.x] += red_a[threadIdx.x + s];
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      area[r] = red_a[0];
      if (sum_prior != nullptr) sum_prior[r] = red_p[0];
    }
    __syncthreads();
  }
}

// ==========================================================================================
// Stride-8 pipeline (DRN stride 8: 8x8-pixel cells), the path of every BASELINE config:
//
//   init_s8     zero the row cursors; prior lookup tables (sums of gx over every 4-bit column
//               subset of each half cell, per cell column; sums of gy / gx per cell)
//   emit_s8     one thread per cell, blocks of 16 x 8 cells.  The 64 labels sit in registers;
//               per distinct label one pass of 64 compares builds the pixel mask (the first
//               pass doubles as the uniformity test: ~70 % of the cells stop there with closed
//               forms).  count, sum of row / column offsets (popcounts) are PACKED into one
//               word, the prior comes from the nibble tables (8 rows x 2 lookups) or, for the
//               last label of a cell, as the complement of the whole-cell prior.  The pair goes
//               straight into the row's BUCKET (128 slots of 16 bytes; slot = one returning
//               atomic on the row cursor, which is also the row length); rows longer than a
//               bucket spill into a global list.  No other atomics: area and the centroid sums
//               are rebuilt exactly from the packed words by rowsort.
//   scan        cursors -> indptr (shared with the generic pipeline), rows > 128 cells listed
//   spill       spilled pairs -> the tail of their row segment (only rows > 128 cells)
//   rowsort     one warp per row: bitonic sort by cell id straight out of the bucket, unpack,
//               integer sums (area, sum_y, sum_x) and the float64 prior in sorted order
//   heavy       rows > 128 cells: dense scatter + ordered compaction
// ==========================================================================================
// Where the float64 prior of a pair is evaluated.  1: in emit, L - 1 masked sums per cell of L
// labels (the last label takes the complement of the whole-cell prior), bucket entries carry the
// prior.  0: in rowsort from the pixel masks the entries carry (every partial pair pays a
// masked sum: rowsort 0.15 -> 0.38 ms per 300 images, K1 1.12 ms against 0.95 ms).
#ifndef K1_PRIOR_IN_EMIT
#define K1_PRIOR_IN_EMIT 1
#endif
constexpr int BUCKET_CAP = 128;
constexpr int TILE_W = 16, TILE_H = 8;  // cells per emit block
#ifndef EMIT2_MIN_BLOCKS
#define EMIT2_MIN_BLOCKS 4
#endif

struct S8Ws {
  int* zero_begin;
  size_t zero_ints;
  int* cursor;       // [R] pairs of the row so far (= row length after emit)
  int* cursor2;      // [R] spilled pairs placed so far
  int* heavy_count;  // [1]
  int* spill_count;  // [1]
  int* tile_sum;
  int* heavy_rows;
  int heavy_cap;
  int4* bucket;      // [R][BUCKET_CAP]: (cell, packed, prior lo, prior hi)
  int* t_row;        // spill list [cap]
  int* t_col;
  int* t_cnt;
  double* t_prior;
  int* u_col;        // [cap] tails of heavy rows at their CSR position; u_prior doubles as the
  int* u_cnt;        //       sorted-prior scratch of 65..128-cell rows
  double* u_prior;
  int* d_cnt;        // [HEAVY_SLOTS * ncell]
  double* d_prior;
  double* gxT;       // [fw][32]
  double* gx8;       // [fw]
  double* gy8;       // [fh]
  double* gyT;       // [8][fh]: gy of pixel row r of cell row cy at gyT[r * fh + cy]
};

size_t carve_s8(S8Ws& ws, void* base, int fh, int fw, int64_t R, int64_t cap) {
  Carver c(base);
  const int ncell = fh * fw;
  int n_tiles = (int)((R + SCAN_TILE - 1) / SCAN_TILE);
  ws.heavy_cap = (int)(cap / (BUCKET_CAP + 1) + 1);
  ws.cursor = c.take<int>(R);
  ws.zero_begin = ws.cursor;
  ws.cursor2 = c.take<int>(R);
  ws.heavy_count = c.take<int>(1);
  ws.spill_count = c.take<int>(1);
  ws.zero_ints = c.off / sizeof(int);
  ws.tile_sum = c.take<int>(n_tiles);
  ws.heavy_rows = c.take<int>(ws.heavy_cap);
  ws.bucket = c.take<int4>((size_t)R * BUCKET_CAP);
  ws.t_row = c.take<int>(cap);
  ws.t_col = c.take<int>(cap);
  ws.t_cnt = c.take<int>(cap);
  ws.t_prior = c.take<double>(cap);
  ws.u_col = c.take<int>(cap);
  ws.u_cnt = c.take<int>(cap);
  ws.u_prior = c.take<double>(cap);
  ws.d_cnt = c.take<int>((size_t)HEAVY_SLOTS * ncell);
  ws.d_prior = c.take<double>((size_t)HEAVY_SLOTS * ncell);
  ws.gxT = c.take<double>((size_t)fw * 32);
  ws.gx8 = c.take<double>(fw);
  ws.gy8 = c.take<double>(fh);
  ws.gyT = c.take<double>((size_t)8 * fh);
  return c.used();
}

// packed pair word: count (7 bits, 1..64) | sum of row offsets (8 bits, <= 224) << 7 |
// sum of column offsets (8 bits) << 15
constexpr int PACKED_FULL = 64 | (224 << 7) | (224 << 15);
__device__ __forceinline__ int packed_cnt(int p) { return p & 127; }
__device__ __forceinline__ int packed_sy(int p) { return (p >> 7) & 255; }
__device__ __forceinline__ int packed_sx(int p) { return (p >> 15) & 255; }

__global__ void init_s8_kernel(S8Ws ws, int fh, int fw, const double* __restrict__ gy,
                               const double* __restrict__ gx, int64_t* nnz_flags) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = i; k < ws.zero_ints; k += stride) ws.zero_begin[k] = 0;
  if (i < 4) nnz_flags[i] = 0;
  if (gy == nullptr) return;
  for (size_t k = i; k < (size_t)fw * 32; k += stride) {
    const int c = (int)(k >> 5), n = (int)(k & 15), half = (int)((k >> 4) & 1);
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((n >> j) & 1) sum = __dadd_rn(sum, gx[c * 8 + half * 4 + j]);
    ws.gxT[k] = sum;
  }
  for (size_t k = i; k < (size_t)fw; k += stride) {
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) sum = __dadd_rn(sum, gx[k * 8 + j]);
    ws.gx8[k] = sum;
  }
  for (size_t k = i; k < (size_t)fh; k += stride) {
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sum = __dadd_rn(sum, gy[k * 8 + j]);
      ws.gyT[(size_t)j * fh + k] = gy[k * 8 + j];
    }
    ws.gy8[k] = sum;
  }
}

__device__ __forceinline__ void store_pair(const S8Ws& ws, int row, int pos, const int4& e,
                                           int64_t spill_cap, int64_t* nnz_flags) {
  if (pos < BUCKET_CAP) {
    ws.bucket[(size_t)row * BUCKET_CAP + pos] = e;
  } else {
    const int sp = atomicAdd(ws.spill_count, 1);
    if (sp < spill_cap) {
      ws.t_row[sp] = row;
      ws.t_col[sp] = e.x;
      ws.t_cnt[sp] = e.y;
      ws.t_prior[sp] = __hiloint2double(e.w, e.z);  // the pixel mask, bit pattern only
    } else {
      atomicOr(reinterpret_cast<unsigned long long*>(&nnz_flags[1]),
               (unsigned long long)SPALIGN_F_NNZ_OVERFLOW);
    }
  }
}
__device__ __forceinline__ void place_pair(const S8Ws& ws, int row, const int4& e,
                                           int64_t spill_cap, int64_t* nnz_flags) {
  store_pair(ws, row, atomicAdd(&ws.cursor[row], 1), e, spill_cap, nnz_flags);
}

// Prior mass of one pair from its pixel mask (lo = pixel rows 0..3, hi = rows 4..7, one byte per
// row): sum_r gy[r] * (sum of gx over the row's set columns), the inner sums looked up in the
// nibble tables of the cell column.  A whole cell is gy8 * gx8.  Fixed evaluation order.
struct PriorTabs {
  const double* gyT;  // [8][fh] (lanes of a warp hold neighbouring cells: one line per load)
  int fh;
  const double* gxT;  // [fw][32]
  const double* gx8;  // [fw]
  const double* gy8;  // [fh]
};
[[maybe_unused]] __device__ __forceinline__ double pair_prior(const PriorTabs& pt, int cyy, int cxx,
                                                              int packed, unsigned lo,
                                                              unsigned hi) {  // (K1_PRIOR_IN_EMIT 0)
  if (packed_cnt(packed) == 64) return __dmul_rn(__ldg(pt.gy8 + cyy), __ldg(pt.gx8 + cxx));
  const double* T = pt.gxT + (size_t)cxx * 32;
  const double* g = pt.gyT + cyy;
  double pr = 0.0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const unsigned b = (r < 4 ? lo >> (8 * r) : hi >> (8 * (r - 4))) & 0xffu;
    const double rs = __dadd_rn(__ldg(T + (b & 15u)), __ldg(T + 16 + (b >> 4)));
    pr = __fma_rn(__ldg(g + (size_t)r * pt.fh), rs, pr);
  }
  return pr;
}

// prior of a bucket entry (c, packed, e2, e3)
__device__ __forceinline__ double entry_prior(const PriorTabs& pt, int cyy, int cxx, int packed,
                                              int e2, int e3) {
#if K1_PRIOR_IN_EMIT
  return __hiloint2double(e3, e2);
#else
  return pair_prior(pt, cyy, cxx, packed, (unsigned)e2, (unsigned)e3);
#endif
}

// bit j of the result = (v[j] == L): one compare and one predicated OR per pixel
__device__ __forceinline__ void mask_or_eq(unsigned& acc, int v, int L, unsigned bit) {
  asm("{ .reg .pred q; setp.eq.s32 q, %1, %2; @q or.b32 %0, %0, %3; }"
      : "+r"(acc)
      : "r"(v), "r"(L), "r"(bit));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
#if K1_LABELS_L1
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
#else
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

#ifndef EMIT_CELLS_PER_THREAD
#define EMIT_CELLS_PER_THREAD 8
#endif
// A/B switches (profiles/README.md, round 2; K1 per 300 DISTINCT label maps on a B200):
//   K1_LABELS_L1 1: cp.async.ca -- labels also allocate in L1, the next-label re-read hits L1:
//                   0.93-0.98 ms against 1.09-1.11 ms with cp.async.cg + L2 re-read
//   K1_SYNCWARP / K1_CANDIDATES: reconverging the warp before every cell / looking for the next
//                   label in fixed register positions first: no measurable difference -> off
#ifndef K1_LABELS_L1
#define K1_LABELS_L1 1
#endif
#ifndef K1_SYNCWARP
#define K1_SYNCWARP 0
#endif
#ifndef K1_CANDIDATES
#define K1_CANDIDATES 0
#endif
constexpr int EMIT_CELLS = EMIT_CELLS_PER_THREAD;  // cells per thread (16-cell tiles along x)


template <typename LabelT>
__global__ void __launch_bounds__(TILE_W * TILE_H, EMIT2_MIN_BLOCKS)
emit_s8v2_kernel(const LabelT* __restrict__ labels, int H, int W, int fh, int fw,
                 const int64_t* __restrict__ sp_off, S8Ws ws, int64_t spill_cap,
                 int64_t* nnz_flags, PriorTabs pt, bool have_prior) {
  // Software pipeline per thread, no block barrier anywhere: the labels of the NEXT cell travel
  // global -> shared memory with cp.async while the current cell is processed out of registers;
  // the pairs of a cell wait in shared memory until all of them are known, their slot atomics
  // are issued back to back, and the returned slots are only consumed after the first pass
  // over the next cell.
  constexpr int PEND = 4, NT = TILE_W * TILE_H;
  constexpr int CH = 64 * (int)sizeof(LabelT) / 16;  // 16-byte chunks per cell
  extern __shared__ int4 s_dyn[];
  int4 (*s_lab)[NT] = reinterpret_cast<int4 (*)[NT]>(s_dyn);  // [chunk][thread]: conflict-free
  int4 (*s_pend)[NT] = reinterpret_cast<int4 (*)[NT]>(s_dyn + CH * NT);
  int (*s_prow)[NT] = reinterpret_cast<int (*)[NT]>(s_dyn + (CH + PEND) * NT);
  // prior tables of the block's 16 cell columns and TILE_H * EMIT_CELLS cell rows: the block walks
  // DOWN the image, so the column tables (33 doubles per column: 2 x 16 nibble sums + the sum of
  // all 8) are loaded once; through L1 they competed with the label stream (L2 read traffic 3x
  // the label bytes, long-scoreboard stalls on every lookup)
  double (*s_T)[33] = reinterpret_cast<double (*)[33]>(s_dyn + (CH + PEND) * NT + PEND * NT / 4);
  double* s_gy = reinterpret_cast<double*>(s_T + TILE_W);      // [EMIT_CELLS * TILE_H * 8]
  double* s_gy8 = s_gy + EMIT_CELLS * TILE_H * 8;              // [EMIT_CELLS * TILE_H]
  const int t = threadIdx.x, tx = t & (TILE_W - 1), ty = t / TILE_W;
  const int img = blockIdx.z;
  const int cx = blockIdx.x * TILE_W + tx;
  const int cy0 = blockIdx.y * (TILE_H * EMIT_CELLS) + ty;
  constexpr int PER_ROW = 8 * (int)sizeof(LabelT) / 16;  // chunks per pixel row of a cell
  const LabelT* pcol = labels + (size_t)img * H * W + (size_t)cx * 8;
  const bool in_x = cx < fw;
  auto prefetch = [&](int cy) {
    const char* src = reinterpret_cast<const char*>(pcol + (size_t)cy * 8 * W);
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int h = 0; h < PER_ROW; ++h)
        cp_async16(&s_lab[r * PER_ROW + h][t], src + ((size_t)r * W) * sizeof(LabelT) + h * 16);
    cp_async_commit();
  };
  if (in_x && cy0 < fh) prefetch(cy0);
#if K1_PRIOR_IN_EMIT
  if (have_prior) {
    for (int e = t; e < TILE_W * 33; e += NT) {
      const int c = e / 33, k = e - c * 33, gc = blockIdx.x * TILE_W + c;
      double val = 0.0;
      if (gc < fw) val = k < 32 ? pt.gxT[(size_t)gc * 32 + k] : pt.gx8[gc];
      s_T[c][k] = val;
    }
    for (int e = t; e < EMIT_CELLS * TILE_H * 8; e += NT) {
      const int cyy = blockIdx.y * (TILE_H * EMIT_CELLS) + (e >> 3);
      s_gy[e] = cyy < fh ? pt.gyT[(size_t)(e & 7) * pt.fh + cyy] : 0.0;
    }
    for (int e = t; e < EMIT_CELLS * TILE_H; e += NT) {
      const int cyy = blockIdx.y * (TILE_H * EMIT_CELLS) + e;
      s_gy8[e] = cyy < fh ? pt.gy8[cyy] : 0.0;
    }
  }
  __syncthreads();   // the only block barrier: tables ready (the first label tile is in flight)
#endif
  if (!in_x || cy0 >= fh) return;
  const int64_t row0 = sp_off[img];
  const int n_sp = (int)(sp_off[img + 1] - row0);
  int npend = 0, pos[PEND];
  bool bad = false;
#pragma unroll 1
  for (int it = 0; it < EMIT_CELLS; ++it) {
    const int cy = cy0 + it * TILE_H;
    // lanes leave the label loop below at different times and the warp stays split into
    // sub-warps for the following cells; forcing it back together did not pay (see above)
#if K1_SYNCWARP
    __syncwarp();
#endif
    if (cy >= fh) break;
    const int c = cy * fw + cx;
    const LabelT* p = pcol + (size_t)cy * 8 * W;
    int v[64];
    cp_async_wait_all();
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int4 q = s_lab[k][t];
      if (sizeof(LabelT) == 4) {
        v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
      } else {
        const long long a = ((long long)q.y << 32) | (unsigned)q.x;
        const long long b2 = ((long long)q.w << 32) | (unsigned)q.z;
        v[2 * k] = (a >= 0 && a < n_sp) ? (int)a : -1;
        v[2 * k + 1] = (b2 >= 0 && b2 < n_sp) ? (int)b2 : -1;
      }
    }
    if (it + 1 < EMIT_CELLS && cy + TILE_H < fh) prefetch(cy + TILE_H);
    unsigned long long remaining = ~0ull;
#if K1_PRIOR_IN_EMIT
    const int ly = it * TILE_H + ty;     // cell row inside the block
    const double cell_prior = have_prior ? __dmul_rn(s_gy8[ly], s_T[tx][32]) : 0.0;
    double emitted_prior = 0.0;
    bool can_complement = true;
#endif
    int L = v[0];
    int npend_prev = npend;  // pairs of the previous cell whose slot atomics are in flight
    npend = 0;
    while (true) {
      unsigned lo0 = 0, lo1 = 0, hi0 = 0, hi1 = 0;
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        mask_or_eq(lo0, v[j], L, 1u << j);
        mask_or_eq(lo1, v[j + 1], L, 2u << j);
        mask_or_eq(hi0, v[32 + j], L, 1u << j);
        mask_or_eq(hi1, v[33 + j], L, 2u << j);
      }
      const unsigned lo = lo0 | lo1, hi = hi0 | hi1;
      const unsigned long long m = (unsigned long long)lo | ((unsigned long long)hi << 32);
      remaining &= ~m;
      // next label: an uncovered pixel among a few fixed positions (corners, edge and centre
      // pixels: static register indices), else the first uncovered pixel re-read from L2
      int nextl = 0;
      bool found = false;
#define SPALIGN_CAND(q)                          \
  if (!found && ((remaining >> (q)) & 1ull)) {   \
    nextl = v[q];                                \
    found = true;                                \
  }
#if K1_CANDIDATES
      SPALIGN_CAND(63) SPALIGN_CAND(7) SPALIGN_CAND(56) SPALIGN_CAND(36) SPALIGN_CAND(27)
      SPALIGN_CAND(60) SPALIGN_CAND(3) SPALIGN_CAND(39) SPALIGN_CAND(24)
#endif
#undef SPALIGN_CAND
      if (remaining != 0ull && !found) {
        const int i = __ffsll((long long)remaining) - 1;
#if K1_LABELS_L1
        const LabelT nextq = __ldg(p + (size_t)(i >> 3) * W + (i & 7));
#else
        const LabelT nextq = __ldcg(p + (size_t)(i >> 3) * W + (i & 7));
#endif
        nextl = (sizeof(LabelT) == 4) ? (int)nextq
                                      : ((nextq >= 0 && nextq < (LabelT)n_sp) ? (int)nextq : -1);
      }
      if (npend_prev) {
        // the previous cell's slots have had a whole pass to come back: store its pairs
#pragma unroll
        for (int k = 0; k < PEND; ++k)
          if (k < npend_prev) store_pair(ws, s_prow[k][t], pos[k], s_pend[k][t], spill_cap, nnz_flags);
        npend_prev = 0;
      }
      if ((unsigned)L < (unsigned)n_sp) {
        const int cnt = __popc(lo) + __popc(hi);
        int packed = PACKED_FULL;
        if (cnt != 64) {
          // bit i = pixel row i >> 3, column i & 7: sums of the offsets of the set bits
          const int sxl = __popcll(m & 0xaaaaaaaaaaaaaaaaull) + 2 * __popcll(m & 0xccccccccccccccccull) +
                          4 * __popcll(m & 0xf0f0f0f0f0f0f0f0ull);
          const int syl = __popcll(m & 0xff00ff00ff00ff00ull) + 2 * __popcll(m & 0xffff0000ffff0000ull) +
                          4 * __popcll(m & 0xffffffff00000000ull);
          packed = cnt | (syl << 7) | (sxl << 15);
        }
#if K1_PRIOR_IN_EMIT
        double pr = cell_prior;
        if (cnt != 64 && have_prior) {
          if (remaining == 0ull && can_complement) {
            // last label of the cell: whole-cell prior minus what the other labels took
            pr = __dadd_rn(cell_prior, -emitted_prior);
          } else {
            pr = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const unsigned b = (r < 4 ? lo >> (8 * r) : hi >> (8 * (r - 4))) & 0xffu;
              const double rs = __dadd_rn(s_T[tx][b & 15u], s_T[tx][16 + (b >> 4)]);
              pr = __fma_rn(s_gy[ly * 8 + r], rs, pr);
            }
            emitted_prior = __dadd_rn(emitted_prior, pr);
          }
        }
        const int e2 = __double2loint(pr), e3 = __double2hiint(pr);
#else
        const int e2 = (int)lo, e3 = (int)hi;
#endif
        if (npend == PEND) {  // more than PEND labels in one cell: place the oldest now
          const int row = s_prow[0][t];
          const int4 e = s_pend[0][t];
#pragma unroll
          for (int k = 0; k + 1 < PEND; ++k) {
            s_prow[k][t] = s_prow[k + 1][t];
            s_pend[k][t] = s_pend[k + 1][t];
          }
          --npend;
          place_pair(ws, row, e, spill_cap, nnz_flags);
        }
        s_prow[npend][t] = (int)(row0 + L);
        s_pend[npend][t] = make_int4(c, packed, e2, e3);
        ++npend;
      } else {  // label outside [0, n_sp): flag it, emit nothing
        bad = true;
#if K1_PRIOR_IN_EMIT
        can_complement = false;
#endif
      }
      if (remaining == 0ull) break;
      L = nextl;
    }
    // slot atomics of all pairs of the cell, back to back; consumed during the next cell
#pragma unroll
    for (int k = 0; k < PEND; ++k)
      if (k < npend) pos[k] = atomicAdd(&ws.cursor[s_prow[k][t]], 1);
  }
#pragma unroll
  for (int k = 0; k < PEND; ++k)
    if (k < npend) store_pair(ws, s_prow[k][t], pos[k], s_pend[k][t], spill_cap, nnz_flags);
  if (bad)
    atomicOr(reinterpret_cast<unsigned long long*>(&nnz_flags[1]),
             (unsigned long long)SPALIGN_F_LABEL_RANGE);
}

// ------------------------------------------------------------------------------------------
// emit v3: the same bucket contract as v2, but the label passes run at full SIMD width.
//
// In v2 every lane walks the distinct labels of its own cell; 68 % of the cells hold one label,
// so most lanes of a warp idle while a few make their second, third, fourth pass (≈1100
// warp-instructions per 32 cells).  Here a warp splits the work:
//   phase A  (every cell, converged): the 64 labels of the lane's cell are loaded straight into
//            registers (16 x 128-bit loads in flight per lane) and folded into one word, OR of
//            (label ^ first label): zero = uniform cell -> one pair with closed-form sums.  Mixed
//            cells are pushed onto the warp's queue in shared memory (ballot + prefix popcount).
//   phase B  (whenever the queue holds 32 cells, and once at the end for the rest): every lane
//            takes one queued cell, re-reads its labels (L1 / L2 hits: DRAM traffic stays at
//            1.07x the label bytes) and runs the label passes, all lanes busy for at least two.
// No block barrier in the loop, no cross-warp traffic, no staging buffer.  Slot atomics never
// stall the lane that issued them: a uniform pair's atomic is consumed after the NEXT cell's
// loads have come back; the pairs of a mixed cell wait in shared memory, their atomics are issued
// back to back when the cell is done and consumed when the lane starts its next mixed cell.
// Measured alternatives (profiles/README.md): cp.async staging of the next cell, prefetch.L2 one
// to three iterations ahead (evicted before use: DRAM traffic 1.5x), a queue of 16-bit label pairs
// in shared memory with fp16x2 compares instead of the re-read -- all slower.
#ifndef EMIT3_CELLS_PER_THREAD
#define EMIT3_CELLS_PER_THREAD 16
#endif
#ifndef EMIT3_MIN_BLOCKS
#define EMIT3_MIN_BLOCKS 4
#endif
#ifndef EMIT3_PICK_TREE
#define EMIT3_PICK_TREE 1
#endif
#ifndef EMIT3_TILE_W
#define EMIT3_TILE_W 32
#endif
// cells per block: E3_W along x (one warp = 32 neighbouring cells = 1 KB of every pixel row) x E3_H
constexpr int E3_W = EMIT3_TILE_W, E3_H = 128 / E3_W;
constexpr int EMIT3_CELLS = EMIT3_CELLS_PER_THREAD;
constexpr int EMIT3_ROWS = E3_H * EMIT3_CELLS;  // cell rows per block
constexpr int EMIT3_PEND = 4;

constexpr size_t emit3_smem_bytes() {
  return (size_t)EMIT3_PEND * (E3_W * E3_H) * 16                          // parked pairs
         + (size_t)(E3_W * E3_H / 32) * 64 * sizeof(int)                  // warp queues
         + (size_t)(E3_W * 33 + EMIT3_ROWS * 9) * sizeof(double);           // prior tables
}

// v[i] for a run-time i out of 64 registers: a six-level select tree (63 selects) instead of a
// trip to memory for the label of the first uncovered pixel
__device__ __forceinline__ int pick64(const int (&v)[64], int i) {
  int a[32], b[16], c[8], d[4];
  const bool b0 = i & 1, b1 = i & 2, b2 = i & 4, b3 = i & 8, b4 = i & 16, b5 = i & 32;
#pragma unroll
  for (int j = 0; j < 32; ++j) a[j] = b0 ? v[2 * j + 1] : v[2 * j];
#pragma unroll
  for (int j = 0; j < 16; ++j) b[j] = b1 ? a[2 * j + 1] : a[2 * j];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j] = b2 ? b[2 * j + 1] : b[2 * j];
#pragma unroll
  for (int j = 0; j < 4; ++j) d[j] = b3 ? c[2 * j + 1] : c[2 * j];
  const int e0 = b4 ? d[1] : d[0], e1 = b4 ? d[3] : d[2];
  return b5 ? e1 : e0;
}

// 32 bytes of labels (one pixel row of a cell of 32-bit labels, half a row of 64-bit ones):
// one 256-bit load where the label map is 32-byte aligned (sm_100 LDG.256: half the load
// instructions and L1 tag look-ups of two 128-bit loads), else two 128-bit loads
template <bool V8>
__device__ __forceinline__ void ld_labels32(const void* p, int4& a, int4& b) {
  if (V8) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z),
                   "=r"(b.w)
                 : "l"(p));
  } else {
    a = __ldg(reinterpret_cast<const int4*>(p));
    b = __ldg(reinterpret_cast<const int4*>(p) + 1);
  }
}

template <typename LabelT, bool V8>
__global__ void __launch_bounds__(E3_W * E3_H, EMIT3_MIN_BLOCKS)
emit_s8v3_kernel(const LabelT* __restrict__ labels, int H, int W, int fh, int fw,
                 const int64_t* __restrict__ sp_off, S8Ws ws, int64_t spill_cap,
                 int64_t* nnz_flags, PriorTabs pt, const double* __restrict__ gy,
                 bool have_prior) {
  constexpr int NT = E3_W * E3_H, NWARP = NT / 32;
  constexpr int PER_ROW = 8 * (int)sizeof(LabelT) / 16;  // 16-byte chunks per pixel row of a cell
  extern __shared__ int4 s_dyn[];
  int4 (*s_pend)[NT] = reinterpret_cast<int4 (*)[NT]>(s_dyn);  // (row, packed, prior lo, hi)
  int (*s_q)[64] = reinterpret_cast<int (*)[64]>(s_dyn + EMIT3_PEND * NT);
  double (*s_T)[33] = reinterpret_cast<double (*)[33]>(s_dyn + EMIT3_PEND * NT + NWARP * 64 / 4);
  double* s_gy = reinterpret_cast<double*>(s_T + E3_W);  // [EMIT3_ROWS * 8] = gy of the rows
  double* s_gy8 = s_gy + EMIT3_ROWS * 8;                   // [EMIT3_ROWS]
  const int t = threadIdx.x, tx = t & (E3_W - 1), ty = t / E3_W;
  const int lane = t & 31, wid = t >> 5;
  const int img = blockIdx.z;
  const int bx0 = blockIdx.x * E3_W, by0 = blockIdx.y * EMIT3_ROWS;
  const int cx = bx0 + tx;
  const bool in_x = cx < fw;
  const LabelT* limg = labels + (size_t)img * H * W;
  const LabelT* pcol = limg + (size_t)cx * 8;
  if (have_prior) {
    // all loads first, then the stores: one round trip instead of one per loop iteration
    constexpr int NTAB = (E3_W * 33 + NT - 1) / NT, NGY = EMIT3_ROWS * 8 / NT;
    double tab[NTAB], g[NGY], g8 = 0.0;
#pragma unroll
    for (int i = 0; i < NTAB; ++i) {
      const int e = t + i * NT, c = e / 33, k = e - c * 33, gc = bx0 + c;
      tab[i] = 0.0;
      if (e < E3_W * 33 && gc < fw) tab[i] = k < 32 ? pt.gxT[(size_t)gc * 32 + k] : pt.gx8[gc];
    }
#pragma unroll
    for (int i = 0; i < NGY; ++i) {
      const int e = t + i * NT;
      g[i] = (by0 * 8 + e < H) ? gy[by0 * 8 + e] : 0.0;
    }
    if (t < EMIT3_ROWS && by0 + t < fh) g8 = pt.gy8[by0 + t];
#pragma unroll
    for (int i = 0; i < NTAB; ++i) {
      const int e = t + i * NT;
      if (e < E3_W * 33) s_T[e / 33][e % 33] = tab[i];
    }
#pragma unroll
    for (int i = 0; i < NGY; ++i) s_gy[t + i * NT] = g[i];
    if (t < EMIT3_ROWS) s_gy8[t] = g8;
  }
  __syncthreads();  // the only block barrier: tables ready
  const int64_t row0 = sp_off[img];
  const int n_sp = (int)(sp_off[img + 1] - row0);
  bool bad = false;
  int npend = 0, pend_c = 0, ppos[EMIT3_PEND];  // parked pairs of the lane's last mixed cell

  auto place_parked = [&]() {
#pragma unroll
    for (int k = 0; k < EMIT3_PEND; ++k)
      if (k < npend) {
        const int4 e = s_pend[k][t];
        store_pair(ws, e.x, ppos[k], make_int4(pend_c, e.y, e.z, e.w), spill_cap, nnz_flags);
      }
    npend = 0;
  };

  // phase B: all label passes of one mixed cell (code = local cell row * 16 + local column)
  auto mixed_cell = [&](int code) {
    place_parked();
    const int ly = code / E3_W, lx = code % E3_W;
    const int cyy = by0 + ly, cxx = bx0 + lx;
    const int c = cyy * fw + cxx;
    const LabelT* p = limg + (size_t)cyy * 8 * W + (size_t)cxx * 8;
    int v[64];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int h2 = 0; h2 < PER_ROW; h2 += 2) {
        int4 qq[2];
        ld_labels32<V8>(reinterpret_cast<const int4*>(p + (size_t)r * W) + h2, qq[0], qq[1]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int4 q = qq[j];
          const int k = r * PER_ROW + h2 + j;
          if (sizeof(LabelT) == 4) {
            v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
          } else {
            const long long a = ((long long)q.y << 32) | (unsigned)q.x;
            const long long b2 = ((long long)q.w << 32) | (unsigned)q.z;
            v[2 * k] = (a >= 0 && a < n_sp) ? (int)a : -1;
            v[2 * k + 1] = (b2 >= 0 && b2 < n_sp) ? (int)b2 : -1;
          }
        }
      }
    const double cell_prior = have_prior ? __dmul_rn(s_gy8[ly], s_T[lx][32]) : 0.0;
    double emitted_prior = 0.0;
    bool can_complement = true;
    unsigned long long remaining = ~0ull;
    int L = v[0];
    while (true) {
      unsigned lo0 = 0, lo1 = 0, hi0 = 0, hi1 = 0;
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        mask_or_eq(lo0, v[j], L, 1u << j);
        mask_or_eq(lo1, v[j + 1], L, 2u << j);
        mask_or_eq(hi0, v[32 + j], L, 1u << j);
        mask_or_eq(hi1, v[33 + j], L, 2u << j);
      }
      const unsigned lo = lo0 | lo1, hi = hi0 | hi1;
      const unsigned long long m = (unsigned long long)lo | ((unsigned long long)hi << 32);
      remaining &= ~m;
      int nextl = 0;
      if (remaining != 0ull) {
#if EMIT3_PICK_TREE
        nextl = pick64(v, __ffsll((long long)remaining) - 1);
#else
        const int i = __ffsll((long long)remaining) - 1;
        const LabelT nextq = __ldg(p + (size_t)(i >> 3) * W + (i & 7));
        nextl = (sizeof(LabelT) == 4) ? (int)nextq
                                      : ((nextq >= 0 && nextq < (LabelT)n_sp) ? (int)nextq : -1);
#endif
      }
      if ((unsigned)L < (unsigned)n_sp) {
        const int cnt = __popc(lo) + __popc(hi);
        // bit i = pixel row i >> 3, column i & 7: sums of the offsets of the set bits
        const int sxl = __popcll(m & 0xaaaaaaaaaaaaaaaaull) + 2 * __popcll(m & 0xccccccccccccccccull) +
                        4 * __popcll(m & 0xf0f0f0f0f0f0f0f0ull);
        const int syl = __popcll(m & 0xff00ff00ff00ff00ull) + 2 * __popcll(m & 0xffff0000ffff0000ull) +
                        4 * __popcll(m & 0xffffffff00000000ull);
        const int packed = cnt | (syl << 7) | (sxl << 15);
        double pr = 0.0;
        if (have_prior) {
          if (remaining == 0ull && can_complement) {
            // last label of the cell: whole-cell prior minus what the other labels took
            pr = __dadd_rn(cell_prior, -emitted_prior);
          } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const unsigned b = (r < 4 ? lo >> (8 * r) : hi >> (8 * (r - 4))) & 0xffu;
              const double rs = __dadd_rn(s_T[lx][b & 15u], s_T[lx][16 + (b >> 4)]);
              pr = __fma_rn(s_gy[ly * 8 + r], rs, pr);
            }
            emitted_prior = __dadd_rn(emitted_prior, pr);
          }
        }
        if (npend == EMIT3_PEND) {  // more than EMIT3_PEND labels in one cell: place the oldest
          const int4 e = s_pend[0][t];
#pragma unroll
          for (int k = 0; k + 1 < EMIT3_PEND; ++k) s_pend[k][t] = s_pend[k + 1][t];
          --npend;
          place_pair(ws, e.x, make_int4(c, e.y, e.z, e.w), spill_cap, nnz_flags);
        }
        s_pend[npend][t] = make_int4((int)(row0 + L), packed, __double2loint(pr),
                                     __double2hiint(pr));
        ++npend;
      } else {  // label outside [0, n_sp): flag it, emit nothing
        bad = true;
        can_complement = false;
      }
      if (remaining == 0ull) break;
      L = nextl;
    }
    pend_c = c;
    // slot atomics of all pairs of the cell, back to back; consumed at the lane's next mixed cell
#pragma unroll
    for (int k = 0; k < EMIT3_PEND; ++k)
      if (k < npend) ppos[k] = atomicAdd(&ws.cursor[s_pend[k][t].x], 1);
  };

  int qn = 0;         // cells in this warp's queue (warp-uniform)
  bool upend = false;  // a uniform pair whose slot atomic is in flight
  int upend_row = 0, upend_c = 0, upend_pos = 0;
  double upend_prior = 0.0;
#pragma unroll 1
  for (int it = 0; it < EMIT3_CELLS; ++it) {
    const int ly = it * E3_H + ty;
    const int cy = by0 + ly;
    if (by0 + it * E3_H + wid * 32 / E3_W >= fh) break;  // warp-uniform: first cell row of the warp
    const bool act = in_x && cy < fh;
    int v0 = 0;
    bool uniform = false, mixed = false;
    if (act) {
      const LabelT* p = pcol + (size_t)cy * 8 * W;
      if (sizeof(LabelT) == 4) {
        int4 q[16];
#pragma unroll
        for (int r = 0; r < 8; ++r) ld_labels32<V8>(p + (size_t)r * W, q[2 * r], q[2 * r + 1]);
        v0 = q[0].x;
        unsigned diff = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          diff |= (unsigned)(q[k].x ^ v0) | (unsigned)(q[k].y ^ v0);
          diff |= (unsigned)(q[k].z ^ v0) | (unsigned)(q[k].w ^ v0);
        }
        mixed = diff != 0;
        uniform = !mixed && (unsigned)v0 < (unsigned)n_sp;
      } else {
        // 64-bit labels, two halves of the cell: a uniform cell is checked on the full value,
        // cells with different out-of-range labels count as mixed (phase B flags them)
        unsigned dlo = 0, dhi = 0;
        int v0h = 0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          int4 q[16];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int h = 0; h < 4; h += 2)
              ld_labels32<V8>(reinterpret_cast<const int4*>(p + (size_t)(4 * half + r) * W) + h,
                              q[4 * r + h], q[4 * r + h + 1]);
          if (half == 0) { v0 = q[0].x; v0h = q[0].y; }
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            dlo |= (unsigned)(q[k].x ^ v0) | (unsigned)(q[k].z ^ v0);
            dhi |= (unsigned)(q[k].y ^ v0h) | (unsigned)(q[k].w ^ v0h);
          }
        }
        mixed = (dlo | dhi) != 0;
        uniform = !mixed && v0h == 0 && (unsigned)v0 < (unsigned)n_sp;
      }
      if (!mixed && !uniform) bad = true;
    }
    if (upend) {  // last iteration's uniform pair: its slot came back while the loads were out
      store_pair(ws, upend_row, upend_pos,
                 make_int4(upend_c, PACKED_FULL, __double2loint(upend_prior),
                           __double2hiint(upend_prior)),
                 spill_cap, nnz_flags);
      upend = false;
    }
    const unsigned mb = __ballot_sync(0xffffffffu, mixed);
    if (mixed) s_q[wid][qn + __popc(mb & ((1u << lane) - 1u))] = ly * E3_W + tx;
    qn += __popc(mb);
    __syncwarp();
    if (qn >= 32) {
      qn -= 32;
      const int code = s_q[wid][qn + lane];
      __syncwarp();
      mixed_cell(code);
      __syncwarp();
    }
    if (uniform) {
      upend_row = (int)(row0 + v0);
      upend_c = cy * fw + cx;
      upend_prior = have_prior ? __dmul_rn(s_gy8[ly], s_T[tx][32]) : 0.0;
      upend_pos = atomicAdd(&ws.cursor[upend_row], 1);
      upend = true;
    }
  }
  if (upend)
    store_pair(ws, upend_row, upend_pos,
               make_int4(upend_c, PACKED_FULL, __double2loint(upend_prior),
                         __double2hiint(upend_prior)),
               spill_cap, nnz_flags);
  if (lane < qn) mixed_cell(s_q[wid][lane]);
  place_parked();
  if (bad)
    atomicOr(reinterpret_cast<unsigned long long*>(&nnz_flags[1]),
             (unsigned long long)SPALIGN_F_LABEL_RANGE);
}

// spilled pairs (rows longer than a bucket) -> tail of their row segment
__global__ void __launch_bounds__(256)
spill_scatter_kernel(S8Ws ws, const int* __restrict__ indptr, int64_t nnz_cap,
                     const int64_t* __restrict__ nnz_flags) {
  if (nnz_flags[1] & SPALIGN_F_NNZ_OVERFLOW) return;
  const int n = *ws.spill_count;
  for (int t = blockIdx.x * 256 + threadIdx.x; t < n; t += gridDim.x * 256) {
    const int r = ws.t_row[t];
    const int64_t pos = (int64_t)indptr[r] + BUCKET_CAP + atomicAdd(&ws.cursor2[r], 1);
    if (pos < nnz_cap) {
      ws.u_col[pos] = ws.t_col[t];
      ws.u_cnt[pos] = ws.t_cnt[t];
      ws.u_prior[pos] = ws.t_prior[t];
    }
  }
}

// c / fw for 0 <= c < 2^24 without the ~20-instruction integer division: float quotient, then
// one exact correction step
__device__ __forceinline__ int div_cells(int c, int fw, float inv_fw, int& rem) {
  int q = __float2int_rz(__int2float_rn(c) * inv_fw);
  int r = c - q * fw;
  if (r < 0) { --q; r += fw; }
  if (r >= fw) { ++q; r -= fw; }
  rem = r;
  return q;
}

// one warp per row of <= BUCKET_CAP cells, straight out of the bucket
__global__ void __launch_bounds__(256)
rowsort_bucket_kernel(S8Ws ws, int* indptr, int64_t R, int fw, int* indices, int* counts,
                      int* area, int64_t* sum_y, int64_t* sum_x, double* sum_prior,
                      const int64_t* __restrict__ nnz_flags, int ncell, PriorTabs pt) {
  __shared__ unsigned s_keys[8][64];
  const int64_t r = (int64_t)blockIdx.x * 8 + warp_id();
  if (r >= R) return;
  const int lane = lane_id();
  if (nnz_flags[1] & SPALIGN_F_NNZ_OVERFLOW) {
    // capacity exceeded: leave an EMPTY matrix behind (no row reaches into unwritten storage),
    // the flag tells the caller
    if (lane == 0) {
      indptr[r] = 0;
      if (r == R - 1) indptr[R] = 0;
      area[r] = 0;
      sum_y[r] = 0;
      sum_x[r] = 0;
      if (sum_prior != nullptr) sum_prior[r] = 0.0;
    }
    return;
  }
  const int L = ws.cursor[r];
  if (L > BUCKET_CAP) return;
  const float inv_fw = 1.0f / (float)fw;
  const int base = indptr[r];
  const int4* bk = ws.bucket + (size_t)r * BUCKET_CAP;
  int a_sum = 0;
  long long y_sum = 0, x_sum = 0;
  double ps = 0.0;
  if (L <= 64 && ncell <= (1 << 24)) {
    // keys (cell << 7 | slot), two per lane (elements lane and lane + 32), to be sorted by cell
    const unsigned PAD = 0xffffffffu;
    const int c0 = lane < L ? bk[lane].x : -1, c1 = lane + 32 < L ? bk[lane + 32].x : -1;
    unsigned k0 = c0 >= 0 ? ((unsigned)c0 << 7) | (unsigned)lane : PAD;
    unsigned k1 = c1 >= 0 ? ((unsigned)c1 << 7) | (unsigned)(lane + 32) : PAD;
    // A superpixel is compact: its cells sit in a small bounding box.  Box of <= 256 cells (every
    // row of a SLIC-like map): one bit per box cell, OR-reduced over the warp; the rank of a
    // cell is the number of set bits below its own -- ~70 warp-instructions against the 435 of
    // the 21-stage network below.
    int x0, x1;
    const int y0 = div_cells(max(c0, 0), fw, inv_fw, x0), y1 = div_cells(max(c1, 0), fw, inv_fw, x1);
    const int big = 0x7fffffff;
    const int ymin = __reduce_min_sync(0xffffffffu, min(c0 >= 0 ? y0 : big, c1 >= 0 ? y1 : big));
    const int ymax = __reduce_max_sync(0xffffffffu, max(c0 >= 0 ? y0 : -1, c1 >= 0 ? y1 : -1));
    const int xmin = __reduce_min_sync(0xffffffffu, min(c0 >= 0 ? x0 : big, c1 >= 0 ? x1 : big));
    const int xmax = __reduce_max_sync(0xffffffffu, max(c0 >= 0 ? x0 : -1, c1 >= 0 ? x1 : -1));
    const int wbox = xmax - xmin + 1, nbits = wbox * (ymax - ymin + 1);
    if (nbits <= 256) {
      const int b0 = c0 >= 0 ? (y0 - ymin) * wbox + (x0 - xmin) : -1;
      const int b1 = c1 >= 0 ? (y1 - ymin) * wbox + (x1 - xmin) : -1;
      int r0 = 0, r1 = 0;
#pragma unroll 1
      for (int wd = 0; wd * 32 < nbits; ++wd) {  // warp-uniform trip count (2 for SLIC-like maps)
        {
          const unsigned mine = ((b0 >> 5) == wd ? 1u << (b0 & 31) : 0u) |
                                ((b1 >> 5) == wd ? 1u << (b1 & 31) : 0u);
          const unsigned word = __reduce_or_sync(0xffffffffu, mine);
          const int full = __popc(word);
          r0 += (b0 >> 5) > wd ? full : ((b0 >> 5) == wd ? __popc(word & ((1u << (b0 & 31)) - 1u)) : 0);
          r1 += (b1 >> 5) > wd ? full : ((b1 >> 5) == wd ? __popc(word & ((1u << (b1 & 31)) - 1u)) : 0);
        }
      }
      // hand the keys to the lanes that own their sorted positions
      unsigned* s_key = s_keys[warp_id()];
      if (c0 >= 0) s_key[r0] = k0;
      if (c1 >= 0) s_key[r1] = k1;
      __syncwarp();
      k0 = lane < L ? s_key[lane] : PAD;
      k1 = lane + 32 < L ? s_key[lane + 32] : PAD;
      __syncwarp();
    } else if (L > 32) {
#pragma unroll
      for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
          if (j == 32) {
            const unsigned mn = min(k0, k1), mx = max(k0, k1);
            k0 = mn;
            k1 = mx;
          } else {
            const unsigned p0 = __shfl_xor_sync(0xffffffffu, k0, j);
            const unsigned p1 = __shfl_xor_sync(0xffffffffu, k1, j);
            const bool lower = (lane & j) == 0;
            const bool up0 = (lane & k) == 0, up1 = ((lane + 32) & k) == 0;
            k0 = (lower == up0) ? min(k0, p0) : max(k0, p0);
            k1 = (lower == up1) ? min(k1, p1) : max(k1, p1);
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
          const unsigned p0 = __shfl_xor_sync(0xffffffffu, k0, j);
          const bool lower = (lane & j) == 0, up0 = (lane & k) == 0 || k == 32;
          k0 = (lower == up0) ? min(k0, p0) : max(k0, p0);
        }
      }
    }
    double pv1 = 0.0;
    if (lane < L) {
      const int4 e = bk[k0 & 127u];
      int cxx;
      const int cn = packed_cnt(e.y), cyy = div_cells(e.x, fw, inv_fw, cxx);
      indices[base + lane] = e.x;
      counts[base + lane] = cn;
      a_sum = cn;
      y_sum = (long long)cn * (cyy * 8) + packed_sy(e.y);
      x_sum = (long long)cn * (cxx * 8) + packed_sx(e.y);
      if (sum_prior != nullptr)
        ps = __dadd_rn(0.0, entry_prior(pt, cyy, cxx, e.y, e.z, e.w));
    }
    if (lane + 32 < L) {
      const int4 e = bk[k1 & 127u];
      int cxx;
      const int cn = packed_cnt(e.y), cyy = div_cells(e.x, fw, inv_fw, cxx);
      indices[base + lane + 32] = e.x;
      counts[base + lane + 32] = cn;
      a_sum += cn;
      y_sum += (long long)cn * (cyy * 8) + packed_sy(e.y);
      x_sum += (long long)cn * (cxx * 8) + packed_sx(e.y);
      if (sum_prior != nullptr) {
        pv1 = entry_prior(pt, cyy, cxx, e.y, e.z, e.w);
        ps = __dadd_rn(ps, pv1);
      }
    }
  } else {
    double* sorted_prior = ws.u_prior;  // this row's CSR range is unused by the heavy tier
    for (int a = 0; a < L; a += 32) {
      const int e = a + lane;
      const bool valid = e < L;
      const int4 me = valid ? bk[e] : make_int4(0x7fffffff, 0, 0, 0);
      int rank = 0;
      for (int b = 0; b < L; b += 32) {
        const int oc = (b + lane < L) ? bk[b + lane].x : 0x7fffffff;
        const int nb = min(32, L - b);
        for (int j = 0; j < nb; ++j) {
          const int o = __shfl_sync(0xffffffffu, oc, j);
          rank += (o < me.x) ? 1 : 0;
        }
      }
      if (valid) {
        const int cn = packed_cnt(me.y), cyy = me.x / fw, cxx = me.x - cyy * fw;
        indices[base + rank] = me.x;
        counts[base + rank] = cn;
        if (sum_prior != nullptr)
          sorted_prior[base + rank] = entry_prior(pt, cyy, cxx, me.y, me.z, me.w);
        a_sum += cn;
        y_sum += (long long)cn * (cyy * 8) + packed_sy(me.y);
        x_sum += (long long)cn * (cxx * 8) + packed_sx(me.y);
      }
    }
    if (sum_prior != nullptr) {
      __syncwarp();
      for (int e = lane; e < L; e += 32) ps = __dadd_rn(ps, sorted_prior[base + e]);
    }
  }
  if (L <= 64 && ncell <= (1 << 24)) {
    // <= 4096 pixels per row and coordinates below 2^15 (ncell <= 2^24 cells of 8x8 pixels would
    // allow 2^27 along one axis only for degenerate shapes; those take the general branch):
    // the integer sums fit 32 bits, one REDUX each
    const bool small = (long long)ncell / fw * 8 <= (1 << 19) && (long long)fw * 8 <= (1 << 19);
    if (small) {
      a_sum = (int)__reduce_add_sync(0xffffffffu, (unsigned)a_sum);
      y_sum = (long long)__reduce_add_sync(0xffffffffu, (unsigned)y_sum);
      x_sum = (long long)__reduce_add_sync(0xffffffffu, (unsigned)x_sum);
    } else {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        a_sum += __shfl_xor_sync(0xffffffffu, a_sum, d);
        y_sum += __shfl_xor_sync(0xffffffffu, y_sum, d);
        x_sum += __shfl_xor_sync(0xffffffffu, x_sum, d);
      }
    }
  } else {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      a_sum += __shfl_xor_sync(0xffffffffu, a_sum, d);
      y_sum += __shfl_xor_sync(0xffffffffu, y_sum, d);
      x_sum += __shfl_xor_sync(0xffffffffu, x_sum, d);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ps = __dadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, d));
  if (lane == 0) {
    area[r] = a_sum;
    sum_y[r] = y_sum;
    sum_x[r] = x_sum;
    if (sum_prior != nullptr) sum_prior[r] = ps;
  }
}

// rows longer than a bucket: bucket + spilled tail -> dense per-cell array -> ordered compaction
__global__ void __launch_bounds__(HEAVY_THREADS)
rowsort_heavy_s8_kernel(S8Ws ws, const int* __restrict__ indptr, int ncell, int fw, int* indices,
                        int* counts, int* area, int64_t* sum_y, int64_t* sum_x,
                        double* sum_prior, const int64_t* __restrict__ nnz_flags, PriorTabs pt) {
  if (nnz_flags[1] & SPALIGN_F_NNZ_OVERFLOW) return;
  const int n_heavy = min(*ws.heavy_count, ws.heavy_cap);
  int* d_cnt = ws.d_cnt + (size_t)blockIdx.x * ncell;
  double* d_prior = ws.d_prior + (size_t)blockIdx.x * ncell;
  __shared__ double red_p[HEAVY_THREADS];
  __shared__ int red_a[HEAVY_THREADS];
  __shared__ long long red_y[HEAVY_THREADS];
  __shared__ long long red_x[HEAVY_THREADS];
  for (int h = blockIdx.x; h < n_heavy; h += gridDim.x) {
    const int r = ws.heavy_rows[h];
    const int base = indptr[r];
    const int L = indptr[r + 1] - base;
    const int4* bk = ws.bucket + (size_t)r * BUCKET_CAP;
    for (int c = threadIdx.x; c < ncell; c += HEAVY_THREADS) d_cnt[c] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < L; e += HEAVY_THREADS) {
      int c, pk;
      double pv;
      if (e < BUCKET_CAP) {
        const int4 q = bk[e];
        c = q.x; pk = q.y; pv = __hiloint2double(q.w, q.z);  // the mask bits
      } else {
        c = ws.u_col[base + e]; pk = ws.u_cnt[base + e]; pv = ws.u_prior[base + e];
      }
      d_cnt[c] = pk;
      d_prior[c] = pv;
    }
    __syncthreads();
    int running = 0;
    int a_sum = 0;
    long long y_sum = 0, x_sum = 0;
    double p_sum = 0.0;
    for (int c0 = 0; c0 < ncell; c0 += HEAVY_THREADS) {
      const int c = c0 + threadIdx.x;
      const int pk = c < ncell ? d_cnt[c] : 0;
      const int flag = pk != 0 ? 1 : 0;
      int total;
      const int excl = block_excl_scan_256(flag, &total);
      if (flag) {
        const int cn = packed_cnt(pk), cyy = c / fw, cxx = c - cyy * fw;
        indices[base + running + excl] = c;
        counts[base + running + excl] = cn;
        a_sum += cn;
        y_sum += (long long)cn * (cyy * 8) + packed_sy(pk);
        x_sum += (long long)cn * (cxx * 8) + packed_sx(pk);
        if (sum_prior != nullptr) {
          const double mb = d_prior[c];  // prior, or the mask bits
          p_sum = __dadd_rn(p_sum, entry_prior(pt, cyy, cxx, pk, __double2loint(mb),
                                               __double2hiint(mb)));
        }
      }
      running += total;
    }
    red_p[threadIdx.x] = p_sum;
    red_a[threadIdx.x] = a_sum;
    red_y[threadIdx.x] = y_sum;
    red_x[threadIdx.x] = x_sum;
    __syncthreads();
    for (int s = HEAVY_THREADS / 2; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) {
        red_p[threadIdx.x] = __dadd_rn(red_p[threadIdx.x], red_p[threadIdx.x + s]);
        red_a[threadIdx.x] += red_a[threadIdx.x + s];
        red_y[threadIdx.x] += red_y[threadIdx.x + s];
        red_x[threadIdx.x] += red_x[threadIdx.x + s];
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      area[r] = red_a[0];
      sum_y[r] = red_y[0];
      sum_x[r] = red_x[0];
      if (sum_prior != nullptr) sum_prior[r] = red_p[0];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
template <typename LabelT>
__global__ void __launch_bounds__(256)
label_max_kernel(const LabelT* __restrict__ labels, int64_t n_pix, int32_t* max_out) {
  const int img = blockIdx.y;
  const LabelT* p = labels + (size_t)img * n_pix;
  long long m = -1;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_pix;
       i += (int64_t)gridDim.x * 256) {
    long long q = (long long)p[i];
    m = q > m ? q : m;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    long long o = __shfl_xor_sync(0xffffffffu, m, d);
    m = o > m ? o : m;
  }
  if (lane_id() == 0) {
    if (m > 0x7ffffffeLL) m = 0x7ffffffeLL;
    atomicMax(&max_out[img], (int)m);
  }
}

}  // namespace
}  // namespace spalign

using namespace spalign;

extern "C" int spalign_label_max(const void* labels, int label_dtype, int n_img, int H, int W,
                                 int32_t* max_out, spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(labels && max_out && n_img > 0 && H > 0 && W > 0, "label_max: bad arguments");
  SPALIGN_REQUIRE(label_dtype == SPALIGN_I32 || label_dtype == SPALIGN_I64,
                  "label_max: label_dtype must be I32 or I64");
  SPALIGN_CUDA(cudaMemsetAsync(max_out, 0xff, sizeof(int32_t) * n_img, stream));
  const int64_t n_pix = (int64_t)H * W;
  int gx = (int)((n_pix + 256 * 16 - 1) / (256 * 16));
  gx = gx < 1 ? 1 : (gx > 4 * kNumSMs ? 4 * kNumSMs : gx);
  dim3 grid(gx, n_img);
  if (label_dtype == SPALIGN_I32)
    label_max_kernel<int32_t><<<grid, 256, 0, stream>>>((const int32_t*)labels, n_pix, max_out);
  else
    label_max_kernel<int64_t><<<grid, 256, 0, stream>>>((const int64_t*)labels, n_pix, max_out);
  return check_launch("label_max");
}

extern "C" size_t spalign_overlap_workspace_bytes(int n_img, int H, int W, int fh, int fw,
                                                  int64_t n_rows, int64_t nnz_cap) {
  OverlapWs ws;
  size_t need = carve(ws, nullptr, n_img, fh * fw, n_rows, nnz_cap) + 256;
  if (H == 8 * fh && W == 8 * fw) {
    S8Ws w8;
    const size_t need8 = carve_s8(w8, nullptr, fh, fw, n_rows, nnz_cap) + 256;
    if (need8 > need) need = need8;
  }
  return need;
}

extern "C" int spalign_overlap_csr(const void* labels, int label_dtype, int n_img, int H, int W,
                                   int fh, int fw, const int64_t* sp_off, int64_t n_rows,
                                   const double* gy, const double* gx, int64_t nnz_cap,
                                   int32_t* indptr, int32_t* indices, int32_t* counts,
                                   int32_t* area, int64_t* sum_y, int64_t* sum_x,
                                   double* sum_prior, int64_t* nnz_flags, void* workspace,
                                   size_t ws_bytes, spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(labels && sp_off && indptr && indices && counts && area && sum_y && sum_x &&
                      nnz_flags && workspace,
                  "overlap_csr: NULL argument");
  SPALIGN_REQUIRE(n_img > 0 && H > 0 && W > 0 && fh > 0 && fw > 0 && n_rows > 0,
                  "overlap_csr: bad shape");
  SPALIGN_REQUIRE(label_dtype == SPALIGN_I32 || label_dtype == SPALIGN_I64,
                  "overlap_csr: label_dtype must be I32 or I64");
  SPALIGN_REQUIRE((gy == nullptr) == (gx == nullptr), "overlap_csr: gy and gx go together");
  SPALIGN_REQUIRE(sum_prior == nullptr || gy != nullptr, "overlap_csr: sum_prior needs gy/gx");
  SPALIGN_REQUIRE(n_rows < 0x7fffffffLL && nnz_cap > 0 && nnz_cap < 0x7fffffffLL,
                  "overlap_csr: n_rows / nnz_cap must fit int32");
  SPALIGN_REQUIRE((int64_t)fh * fw < 0x7fffffffLL, "overlap_csr: too many cells");
  const int cap_img = (int)(nnz_cap / n_img);
  SPALIGN_REQUIRE(cap_img > 0, "overlap_csr: nnz_cap smaller than n_img");
  const int ncell = fh * fw;
  OverlapWs ws;
  size_t need = spalign_overlap_workspace_bytes(n_img, H, W, fh, fw, n_rows, nnz_cap);
  if (ws_bytes < need) {
    set_error("overlap_csr: workspace %zu < %zu bytes", ws_bytes, need);
    return SPALIGN_E_WORKSPACE;
  }
  void* aligned = reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  const bool s8 = (H == 8 * fh) && (W == 8 * fw) &&
                  (reinterpret_cast<size_t>(labels) % 16 == 0);
  if (s8 && getenv("SPALIGN_K1_LEGACY") == nullptr) {
    S8Ws w8;
    carve_s8(w8, aligned, fh, fw, n_rows, nnz_cap);
    init_s8_kernel<<<2 * kNumSMs, 256, 0, stream>>>(w8, fh, fw, gy, gx, nnz_flags);
    constexpr int NT = TILE_W * TILE_H;
    const PriorTabs pt{w8.gyT, fh, w8.gxT, w8.gx8, w8.gy8};
    const char* emit_sel = getenv("SPALIGN_K1_EMIT");  // "2": round 2's one-lane-per-cell passes
    if (emit_sel != nullptr && emit_sel[0] == '2') {
      dim3 egrid((fw + TILE_W - 1) / TILE_W,
                 (fh + TILE_H * EMIT_CELLS - 1) / (TILE_H * EMIT_CELLS), n_img);
      const size_t tab = (size_t)(TILE_W * 33 + EMIT_CELLS * TILE_H * 9) * sizeof(double);
      if (label_dtype == SPALIGN_I32) {
        const size_t smem = (size_t)NT * ((16 + 4) * 16 + 4 * 4) + tab;
        SPALIGN_CUDA(cudaFuncSetAttribute(emit_s8v2_kernel<int32_t>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        emit_s8v2_kernel<int32_t><<<egrid, NT, smem, stream>>>(
            (const int32_t*)labels, H, W, fh, fw, sp_off, w8, nnz_cap, nnz_flags, pt, gy != nullptr);
      } else {
        const size_t smem = (size_t)NT * ((32 + 4) * 16 + 4 * 4) + tab;
        SPALIGN_CUDA(cudaFuncSetAttribute(emit_s8v2_kernel<int64_t>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        emit_s8v2_kernel<int64_t><<<egrid, NT, smem, stream>>>(
            (const int64_t*)labels, H, W, fh, fw, sp_off, w8, nnz_cap, nnz_flags, pt, gy != nullptr);
      }
    } else {
      dim3 egrid((fw + E3_W - 1) / E3_W, (fh + EMIT3_ROWS - 1) / EMIT3_ROWS, n_img);
      constexpr size_t smem = emit3_smem_bytes();
      const bool v8 = reinterpret_cast<size_t>(labels) % 32 == 0 &&  // 256-bit label loads
                      getenv("SPALIGN_K1_LD128") == nullptr;
#define SPALIGN_EMIT3(LT, V8)                                                                  \
  do {                                                                                         \
    SPALIGN_CUDA(cudaFuncSetAttribute(emit_s8v3_kernel<LT, V8>,                                \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    emit_s8v3_kernel<LT, V8><<<egrid, NT, smem, stream>>>(                                     \
        (const LT*)labels, H, W, fh, fw, sp_off, w8, nnz_cap, nnz_flags, pt, gy, gy != nullptr); \
  } while (0)
      if (label_dtype == SPALIGN_I32) {
        if (v8) SPALIGN_EMIT3(int32_t, true); else SPALIGN_EMIT3(int32_t, false);
      } else {
        if (v8) SPALIGN_EMIT3(int64_t, true); else SPALIGN_EMIT3(int64_t, false);
      }
#undef SPALIGN_EMIT3
    }
    const int n_tiles = (int)((n_rows + SCAN_TILE - 1) / SCAN_TILE);
    scan_tile_sums_kernel<<<n_tiles, 256, 0, stream>>>(w8.cursor, n_rows, w8.tile_sum);
    scan_finish_kernel<<<n_tiles, 256, 0, stream>>>(w8.cursor, n_rows, w8.tile_sum, n_tiles, indptr,
                                                    nnz_flags, nnz_cap, w8.heavy_rows,
                                                    w8.heavy_count, w8.heavy_cap, nullptr, 0,
                                                    BUCKET_CAP);
    spill_scatter_kernel<<<2 * kNumSMs, 256, 0, stream>>>(w8, indptr, nnz_cap, nnz_flags);
    rowsort_bucket_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, stream>>>(
        w8, indptr, n_rows, fw, indices, counts, area, sum_y, sum_x, sum_prior, nnz_flags, ncell,
        pt);
    rowsort_heavy_s8_kernel<<<HEAVY_SLOTS, HEAVY_THREADS, 0, stream>>>(
        w8, indptr, ncell, fw, indices, counts, area, sum_y, sum_x, sum_prior, nnz_flags, pt);
    return check_launch("overlap_csr");
  }
  carve(ws, aligned, n_img, ncell, n_rows, nnz_cap);

  init_kernel<<<2 * kNumSMs, 256, 0, stream>>>(ws.zero_begin, ws.zero_ints, sum_y, sum_x, n_rows,
                                               nnz_flags);
  dim3 egrid((ncell + EMIT_THREADS - 1) / EMIT_THREADS, n_img);
  if (label_dtype == SPALIGN_I32) {
    if (s8)
      emit_s8_kernel<int32_t><<<egrid, EMIT_THREADS, 0, stream>>>(
          (const int32_t*)labels, H, W, fh, fw, sp_off, gy, gx, cap_img, ws, sum_y, sum_x,
          nnz_flags);
    else
      emit_generic_kernel<int32_t><<<egrid, EMIT_THREADS, 0, stream>>>(
          (const int32_t*)labels, H, W, fh, fw, sp_off, gy, gx, cap_img, ws, sum_y, sum_x,
          nnz_flags);
  } else {
    if (s8)
      emit_s8_kernel<int64_t><<<egrid, EMIT_THREADS, 0, stream>>>(
          (const int64_t*)labels, H, W, fh, fw, sp_off, gy, gx, cap_img, ws, sum_y, sum_x,
          nnz_flags);
    else
      emit_generic_kernel<int64_t><<<egrid, EMIT_THREADS, 0, stream>>>(
          (const int64_t*)labels, H, W, fh, fw, sp_off, gy, gx, cap_img, ws, sum_y, sum_x,
          nnz_flags);
  }
  const int n_tiles = (int)((n_rows + SCAN_TILE - 1) / SCAN_TILE);
  scan_tile_sums_kernel<<<n_tiles, 256, 0, stream>>>(ws.row_nnz, n_rows, ws.tile_sum);
  scan_finish_kernel<<<n_tiles, 256, 0, stream>>>(ws.row_nnz, n_rows, ws.tile_sum, n_tiles,
                                                  indptr, nnz_flags, nnz_cap, ws.heavy_rows,
                                                  ws.heavy_count, ws.heavy_cap, ws.pair_count,
                                                  n_img, WARP_TIER_MAX);
  int sgx = (cap_img + 256 * 4 - 1) / (256 * 4);
  sgx = sgx < 1 ? 1 : (sgx > 64 ? 64 : sgx);
  scatter_kernel<<<dim3(sgx, n_img), 256, 0, stream>>>(ws, cap_img, indptr, nnz_cap, nnz_flags);
  rowsort_warp_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, stream>>>(
      ws, indptr, n_rows, indices, counts, area, sum_prior, nullptr, nnz_flags, ncell);
  rowsort_heavy_kernel<<<HEAVY_SLOTS, HEAVY_THREADS, 0, stream>>>(ws, indptr, ncell, indices,
                                                                 counts, area, sum_prior,
                                                                 nullptr, nnz_flags);
  return check_launch("overlap_csr");
}

extern "C" size_t spalign_overlap_bilinear_workspace_bytes(int n_img, int H, int W, int fh, int fw,
                                                           int64_t n_rows, int64_t nnz_cap) {
  (void)H;
  (void)W;
  OverlapWs ws;
  // + counts[cap], area[R], sum_y[R], sum_x[R], sum_prior[R] scratch that the caller does not see
  return carve(ws, nullptr, n_img, fh * fw, n_rows, nnz_cap) + (size_t)nnz_cap * 4 +
         (size_t)n_rows * (4 + 8 + 8 + 8) + 6 * 256;
}

extern "C" int spalign_overlap_bilinear_csr(
    const void* labels, int label_dtype, int n_img, int H, int W, int fh, int fw,
    const int64_t* sp_off, int64_t n_rows, const int32_t* iy0, const double* wy0,
    const double* wy1, const int32_t* ystart, const int32_t* ix0, const double* wx0,
    const double* wx1, const int32_t* xstart, int64_t nnz_cap, int32_t* indptr, int32_t* indices,
    double* wvals, double* row_weight, int64_t* nnz_flags, void* workspace, size_t ws_bytes,
    spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(labels && sp_off && iy0 && wy0 && wy1 && ystart && ix0 && wx0 && wx1 && xstart &&
                      indptr && indices && wvals && nnz_flags && workspace,
                  "overlap_bilinear_csr: NULL argument");
  SPALIGN_REQUIRE(n_img > 0 && H > 0 && W > 0 && fh >= 2 && fw >= 2 && n_rows > 0,
                  "overlap_bilinear_csr: bad shape (fh, fw must be >= 2)");
  SPALIGN_REQUIRE(label_dtype == SPALIGN_I32 || label_dtype == SPALIGN_I64,
                  "overlap_bilinear_csr: label_dtype must be I32 or I64");
  SPALIGN_REQUIRE(n_rows < 0x7fffffffLL && nnz_cap > 0 && nnz_cap < 0x7fffffffLL,
                  "overlap_bilinear_csr: n_rows / nnz_cap must fit int32");
  const int cap_img = (int)(nnz_cap / n_img);
  SPALIGN_REQUIRE(cap_img > 0, "overlap_bilinear_csr: nnz_cap smaller than n_img");
  const int ncell = fh * fw;
  size_t need = spalign_overlap_bilinear_workspace_bytes(n_img, H, W, fh, fw, n_rows, nnz_cap);
  if (ws_bytes < need) {
    set_error("overlap_bilinear_csr: workspace %zu < %zu bytes", ws_bytes, need);
    return SPALIGN_E_WORKSPACE;
  }
  void* aligned = reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  OverlapWs ws;
  size_t used = carve(ws, aligned, n_img, ncell, n_rows, nnz_cap);
  Carver extra(static_cast<char*>(aligned) + used);
  int* counts = extra.take<int>(nnz_cap);
  int* area = extra.take<int>(n_rows);
  int64_t* sum_y = extra.take<int64_t>(n_rows);
  int64_t* sum_x = extra.take<int64_t>(n_rows);
  double* sum_prior = row_weight ? row_weight : extra.take<double>(n_rows);

  init_kernel<<<2 * kNumSMs, 256, 0, stream>>>(ws.zero_begin, ws.zero_ints, sum_y, sum_x, n_rows,
                                               nnz_flags);
  BilinearAxes ax{iy0, wy0, wy1, ystart, ix0, wx0, wx1, xstart};
  dim3 egrid((ncell + EMIT_THREADS - 1) / EMIT_THREADS, n_img);
  if (label_dtype == SPALIGN_I32)
    emit_bilinear_kernel<int32_t><<<egrid, EMIT_THREADS, 0, stream>>>(
        (const int32_t*)labels, H, W, fh, fw, sp_off, ax, cap_img, ws, sum_y, sum_x, nnz_flags);
  else
    emit_bilinear_kernel<int64_t><<<egrid, EMIT_THREADS, 0, stream>>>(
        (const int64_t*)labels, H, W, fh, fw, sp_off, ax, cap_img, ws, sum_y, sum_x, nnz_flags);
  const int n_tiles = (int)((n_rows + SCAN_TILE - 1) / SCAN_TILE);
  scan_tile_sums_kernel<<<n_tiles, 256, 0, stream>>>(ws.row_nnz, n_rows, ws.tile_sum);
  scan_finish_kernel<<<n_tiles, 256, 0, stream>>>(ws.row_nnz, n_rows, ws.tile_sum, n_tiles,
                                                  indptr, nnz_flags, nnz_cap, ws.heavy_rows,
                                                  ws.heavy_count, ws.heavy_cap, ws.pair_count,
                                                  n_img, WARP_TIER_MAX);
  int sgx = (cap_img + 256 * 4 - 1) / (256 * 4);
  sgx = sgx < 1 ? 1 : (sgx > 64 ? 64 : sgx);
  scatter_kernel<<<dim3(sgx, n_img), 256, 0, stream>>>(ws, cap_img, indptr, nnz_cap, nnz_flags);
  rowsort_warp_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, stream>>>(
      ws, indptr, n_rows, indices, counts, area, sum_prior, wvals, nnz_flags, ncell);
  rowsort_heavy_kernel<<<HEAVY_SLOTS, HEAVY_THREADS, 0, stream>>>(ws, indptr, ncell, indices,
                                                                 counts, area, sum_prior, wvals,
                                                                 nnz_flags);
  return check_launch("overlap_bilinear_csr");
}
