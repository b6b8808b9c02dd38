// f3 (second branch of batch_superpixel): Felzenszwalb-Huttenlocher graph segmentation on the
// device, the reference's default --superpixel_method (batch_spalign_kmeans.py:301-307:
// skimage.segmentation.felzenszwalb(img / 255., scale=300, sigma=0.8, min_size=20), one image at a
// time on the CPU).  scikit-image is not in the reference tree: the contract is the algorithm of
// scikit-image 0.13 (_felzenszwalb_cy.pyx) as restated in oracle/spalign_oracle.py:felzenszwalb;
// PARITY UNPINNED by the reference.  Every floating-point step uses explicit round-to-nearest
// float64 operations in the oracle's order, so the labels are bit-identical to the restatement.
//
//   blur      separable Gaussian (scipy.ndimage.gaussian_filter, mode 'reflect'): axis 0, axis 1
//   costs     8-connectivity edge weights (right, down, down-right, up-right), sqrt of the sum of
//             squared channel differences -> sortable 64-bit keys
//   sort      stable radix sort of (cost, edge index) per image (cub::DeviceRadixSort)
//   merge     the greedy pass over the sorted edges is sequential by definition (every decision
//             depends on the component sizes all earlier merges left): one warp per image walks
//             the edges -- 32 at a time are fetched and decoded by the lanes, lane 0 runs the
//             union-find (parents and set sizes in shared memory when the image fits: 224 x 224
//             does) -- and images run side by side, one CTA each.  Then the min_size pass.
//   relabel   roots -> 0..S-1 in ascending root order (np.unique), block-wide scan
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace spalign {
namespace {

constexpr int FELZ_MAX_RADIUS = 16;
struct FelzWeights {
  int radius;
  double w[FELZ_MAX_RADIUS + 1];
};

__device__ __forceinline__ int reflect_index(int i, int n) {
  while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
  return i;
}

// one pass of the separable filter over [n_img][3][H][W] float64 planes along axis (0: y, 1: x)
template <typename InT>
__global__ void __launch_bounds__(256)
felz_blur_kernel(const InT* __restrict__ in, double* __restrict__ out, int H, int W, int axis,
                 FelzWeights fw, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const InT* plane = in + (i - (int64_t)y * W - x);
  const int n = axis == 0 ? H : W, pos = axis == 0 ? y : x;
  const int64_t stride = axis == 0 ? W : 1;
  const InT* line = plane + (axis == 0 ? x : (int64_t)y * W);
  double acc = __dmul_rn((double)line[(int64_t)pos * stride], fw.w[0]);
  for (int d = fw.radius; d >= 1; --d) {
    const double a = (double)line[(int64_t)reflect_index(pos - d, n) * stride];
    const double b = (double)line[(int64_t)reflect_index(pos + d, n) * stride];
    acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(a, b), fw.w[d]));
  }
  out[i] = acc;
}

struct FelzGeom {
  int H, W;
  int64_t n_right, n_down, n_diag;  // edges per class (down-right and up-right: n_diag each)
  int64_t E;
};
__host__ __device__ inline FelzGeom felz_geom(int H, int W) {
  FelzGeom g;
  g.H = H; g.W = W;
  g.n_right = (int64_t)H * (W - 1);
  g.n_down = (int64_t)(H - 1) * W;
  g.n_diag = (int64_t)(H - 1) * (W - 1);
  g.E = g.n_right + g.n_down + 2 * g.n_diag;
  return g;
}
// endpoints (a, b) of edge e in skimage's order
__device__ __forceinline__ void felz_edge(const FelzGeom& g, int64_t e, int& a, int& b) {
  const int W = g.W;
  if (e < g.n_right) {
    const int r = (int)(e / (W - 1)), c = (int)(e % (W - 1));
    a = r * W + c + 1; b = r * W + c;
  } else if ((e -= g.n_right) < g.n_down) {
    const int r = (int)(e / W), c = (int)(e % W);
    a = (r + 1) * W + c; b = r * W + c;
  } else if ((e -= g.n_down) < g.n_diag) {
    const int r = (int)(e / (W - 1)), c = (int)(e % (W - 1));
    a = (r + 1) * W + c + 1; b = r * W + c;
  } else {
    e -= g.n_diag;
    const int r = (int)(e / (W - 1)), c = (int)(e % (W - 1));
    a = r * W + c + 1; b = (r + 1) * W + c;
  }
}

// keys[img][e] = bits of the (non-negative) float64 cost, vals[img][e] = e
__global__ void __launch_bounds__(256)
felz_cost_kernel(const double* __restrict__ sm, FelzGeom g, unsigned long long* keys,
                 unsigned* vals) {
  const int img = blockIdx.y;
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= g.E) return;
  int a, b;
  felz_edge(g, e, a, b);
  // the up-right class pairs edge (r, c+1)-(r+1, c) with the difference image[r+1, c] - image[r, c+1]
  const bool up = e >= g.n_right + g.n_down + g.n_diag;
  const int p = up ? b : a, q = up ? a : b;
  const int64_t hw = (int64_t)g.H * g.W;
  const double* base = sm + (size_t)img * 3 * hw;
  const double d0 = __dadd_rn(base[p], -base[q]);
  const double d1 = __dadd_rn(base[hw + p], -base[hw + q]);
  const double d2 = __dadd_rn(base[2 * hw + p], -base[2 * hw + q]);
  const double s = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
  keys[(size_t)img * g.E + e] = (unsigned long long)__double_as_longlong(__dsqrt_rn(s));
  vals[(size_t)img * g.E + e] = (unsigned)e;
}

// union-find in ONE int per pixel: parent[i] >= 0 is the parent, parent[i] < 0 marks a root and
// holds minus the size of its set (so the sizes live next to the parents, in shared memory)
__device__ __forceinline__ int felz_find(int* parent, int i) {
  int p = parent[i];
  while (p >= 0) {
    const int gp = parent[p];
    if (gp < 0) return p;
    parent[i] = gp;  // path halving: never changes a root
    i = gp;
    p = parent[i];
  }
  return i;
}

// greedy merge + min_size pass of one image per CTA (one warp)
__global__ void __launch_bounds__(32)
felz_merge_kernel(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ order,
                  FelzGeom g, double scale, int min_size, int* parent_g, double* cint_g,
                  int use_smem) {
  extern __shared__ int s_parent[];
  const int img = blockIdx.x, lane = threadIdx.x;
  const int n = g.H * g.W;
  int* parent = use_smem ? s_parent : parent_g + (size_t)img * n;
  double* cint = cint_g + (size_t)img * n;
  for (int i = lane; i < n; i += 32) {
    parent[i] = -1;
    cint[i] = 0.0;
  }
  __syncwarp();
  const unsigned long long* kk = keys + (size_t)img * g.E;
  const unsigned* oo = order + (size_t)img * g.E;
  for (int pass = 0; pass < 2; ++pass) {
    for (int64_t e0 = 0; e0 < g.E; e0 += 32) {
      int a = 0, b = 0;
      double c = 0.0;
      if (e0 + lane < g.E) {
        felz_edge(g, (int64_t)oo[e0 + lane], a, b);
        c = __longlong_as_double((long long)kk[e0 + lane]);
      }
      const int cnt = (int)min((int64_t)32, g.E - e0);
      for (int j = 0; j < cnt; ++j) {
        const int ea = __shfl_sync(0xffffffffu, a, j), eb = __shfl_sync(0xffffffffu, b, j);
        const double ec = __shfl_sync(0xffffffffu, c, j);
        if (lane == 0) {
          const int ra = felz_find(parent, ea), rb = felz_find(parent, eb);
          if (ra != rb) {
            const int sa = -parent[ra], sb = -parent[rb];
            bool join;
            if (pass == 0) {
              const double ia = __dadd_rn(cint[ra], __ddiv_rn(scale, (double)sa));
              const double ib = __dadd_rn(cint[rb], __ddiv_rn(scale, (double)sb));
              join = ec < fmin(ia, ib);
            } else {
              join = sa < min_size || sb < min_size;
            }
            if (join) {
              const int r = min(ra, rb), o = max(ra, rb);
              parent[o] = r;
              parent[r] = -(sa + sb);
              if (pass == 0) cint[r] = ec;
            }
          }
        }
      }
    }
    __syncwarp();
  }
  if (use_smem) {
    int* pg = parent_g + (size_t)img * n;
    for (int i = lane; i < n; i += 32) pg[i] = s_parent[i];
  }
}

// labels = rank of the root among the sorted roots (np.unique(...)[1]); one CTA per image
__global__ void __launch_bounds__(1024)
felz_relabel_kernel(int* parent_g, int n, int* rank_scratch, int32_t* labels, int32_t* n_labels) {
  const int img = blockIdx.x, t = threadIdx.x;
  int* parent = parent_g + (size_t)img * n;
  int* rank = rank_scratch + (size_t)img * n;
  __shared__ int s_warp[32];
  __shared__ int s_base;
  if (t == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + t;
    const int flag = (i < n && parent[i] < 0) ? 1 : 0;
    int incl = flag;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if ((t & 31) >= d) incl += v;
    }
    if ((t & 31) == 31) s_warp[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
      int v = s_warp[t];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, d);
        if (t >= d) v += u;
      }
      s_warp[t] = v;
    }
    __syncthreads();
    const int before = s_base + ((t >> 5) ? s_warp[(t >> 5) - 1] : 0) + incl - flag;
    if (flag) rank[i] = before;
    __syncthreads();
    if (t == 0) s_base += s_warp[31];
    __syncthreads();
  }
  if (t == 0) n_labels[img] = s_base;
  for (int i = t; i < n; i += 1024) {
    int r = i;
    while (parent[r] >= 0) r = parent[r];
    labels[(size_t)img * n + i] = rank[r];
  }
}

struct FelzWs {
  double* plane_a;   // [n_img][3][H][W]
  double* plane_b;
  unsigned long long* keys;      // [n_img][E] and the radix sort's alternate buffer
  unsigned long long* keys_alt;
  unsigned* vals;
  unsigned* vals_alt;
  int* parent;       // [n_img][H*W] parent, or minus the set size at a root
  double* cint;      // (reused as rank scratch by the relabel pass)
  void* sort_tmp;
  size_t sort_tmp_bytes;
};
constexpr size_t FELZ_SORT_TMP = 8u << 20;

size_t felz_carve(FelzWs& ws, void* base, int n_img, int H, int W) {
  Carver c(base);
  const FelzGeom g = felz_geom(H, W);
  const size_t hw = (size_t)H * W;
  ws.plane_a = c.take<double>((size_t)n_img * 3 * hw);
  ws.plane_b = c.take<double>((size_t)n_img * 3 * hw);
  ws.keys = c.take<unsigned long long>((size_t)n_img * g.E);
  ws.keys_alt = c.take<unsigned long long>((size_t)n_img * g.E);
  ws.vals = c.take<unsigned>((size_t)n_img * g.E);
  ws.vals_alt = c.take<unsigned>((size_t)n_img * g.E);
  ws.parent = c.take<int>((size_t)n_img * hw);
  ws.cint = c.take<double>((size_t)n_img * hw);
  ws.sort_tmp = c.take<char>(FELZ_SORT_TMP);
  ws.sort_tmp_bytes = FELZ_SORT_TMP;
  return c.used();
}

}  // namespace
}  // namespace spalign

using namespace spalign;

extern "C" size_t spalign_felzenszwalb_workspace_bytes(int n_img, int H, int W) {
  if (n_img <= 0 || H < 2 || W < 2) return 0;
  FelzWs ws;
  return felz_carve(ws, nullptr, n_img, H, W) + 256;
}

extern "C" int spalign_felzenszwalb(const float* images, int n_img, int H, int W, double scale,
                                    double sigma, int min_size, int32_t* labels,
                                    int32_t* n_labels, void* workspace, size_t ws_bytes,
                                    spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(images && labels && n_labels && workspace, "felzenszwalb: NULL argument");
  SPALIGN_REQUIRE(n_img > 0 && H >= 2 && W >= 2 && (int64_t)H * W < (1ll << 29) && sigma >= 0.0 &&
                      scale >= 0.0 && min_size >= 0,
                  "felzenszwalb: bad arguments");
  const size_t need = spalign_felzenszwalb_workspace_bytes(n_img, H, W);
  if (ws_bytes < need) {
    set_error("felzenszwalb: workspace %zu < %zu bytes", ws_bytes, need);
    return SPALIGN_E_WORKSPACE;
  }
  FelzWs ws;
  felz_carve(ws, reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(workspace), 256)), n_img,
             H, W);
  const FelzGeom g = felz_geom(H, W);
  const int64_t total = (int64_t)n_img * 3 * H * W;
  const unsigned blur_blocks = (unsigned)((total + 255) / 256);
  // scipy.ndimage._gaussian_kernel1d (order 0), weights by distance from the centre
  FelzWeights fw;
  fw.radius = (int)(4.0 * sigma + 0.5);
  SPALIGN_REQUIRE(fw.radius <= FELZ_MAX_RADIUS, "felzenszwalb: sigma too large");
  {
    const double sd = sigma * sigma;
    double tot = 1.0;
    fw.w[0] = 1.0;
    for (int ii = 1; ii <= fw.radius; ++ii) {
      fw.w[ii] = exp(-0.5 * (double)(ii * ii) / sd);
      tot += 2.0 * fw.w[ii];
    }
    for (int ii = 0; ii <= fw.radius; ++ii) fw.w[ii] /= tot;
  }
  const double* smooth;
  if (fw.radius > 0) {
    felz_blur_kernel<float><<<blur_blocks, 256, 0, stream>>>(images, ws.plane_a, H, W, 0, fw, total);
    felz_blur_kernel<double><<<blur_blocks, 256, 0, stream>>>(ws.plane_a, ws.plane_b, H, W, 1, fw,
                                                               total);
    smooth = ws.plane_b;
  } else {  // sigma = 0: the image itself, as float64 (a radius-0 pass multiplies by 1)
    fw.w[0] = 1.0;
    felz_blur_kernel<float><<<blur_blocks, 256, 0, stream>>>(images, ws.plane_a, H, W, 0, fw, total);
    smooth = ws.plane_a;
  }
  felz_cost_kernel<<<dim3((unsigned)((g.E + 255) / 256), n_img), 256, 0, stream>>>(smooth, g, ws.keys,
                                                                                    ws.vals);
  // stable LSD radix sort per image: equal costs keep their edge order
  const unsigned long long* skeys = ws.keys;
  const unsigned* svals = ws.vals;
  for (int i = 0; i < n_img; ++i) {
    cub::DoubleBuffer<unsigned long long> dk(ws.keys + (size_t)i * g.E, ws.keys_alt + (size_t)i * g.E);
    cub::DoubleBuffer<unsigned> dv(ws.vals + (size_t)i * g.E, ws.vals_alt + (size_t)i * g.E);
    size_t tmp = 0;
    SPALIGN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int)g.E, 0, 64, stream));
    if (tmp > ws.sort_tmp_bytes) {
      set_error("felzenszwalb: radix sort needs %zu bytes of temporary storage", tmp);
      return SPALIGN_E_WORKSPACE;
    }
    SPALIGN_CUDA(cub::DeviceRadixSort::SortPairs(ws.sort_tmp, tmp, dk, dv, (int)g.E, 0, 64, stream));
    // 8 passes of 8 bits would end in the original buffer; CUB picks its own digit width, so
    // move the result where the merge kernel reads it if it ended in the alternate
    if (dk.Current() != ws.keys + (size_t)i * g.E) {
      SPALIGN_CUDA(cudaMemcpyAsync(ws.keys + (size_t)i * g.E, dk.Current(),
                                   sizeof(unsigned long long) * g.E, cudaMemcpyDeviceToDevice,
                                   stream));
      SPALIGN_CUDA(cudaMemcpyAsync(ws.vals + (size_t)i * g.E, dv.Current(), sizeof(unsigned) * g.E,
                                   cudaMemcpyDeviceToDevice, stream));
    }
  }
  const size_t par_bytes = (size_t)H * W * sizeof(int);
  const int use_smem = par_bytes <= 200 * 1024 ? 1 : 0;
  if (use_smem)
    SPALIGN_CUDA(cudaFuncSetAttribute(felz_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)par_bytes));
  felz_merge_kernel<<<n_img, 32, use_smem ? par_bytes : 0, stream>>>(
      skeys, svals, g, scale / 255.0, min_size, ws.parent, ws.cint, use_smem);
  felz_relabel_kernel<<<n_img, 1024, 0, stream>>>(ws.parent, H * W, reinterpret_cast<int*>(ws.cint),
                                                  labels, n_labels);
  return check_launch("felzenszwalb");
}
