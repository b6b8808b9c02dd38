// K3: prior-weighted k-means (replaces kmeans(), batch_spalign_kmeans.py:136-183).
//
// One "sweep" streams the rows of X through shared memory in tiles of TR rows
// (cp.async, double buffered) and does, per tile:
//   phase 1  distances: thread (row, part) accumulates sum_d (x-c_k)^2 in float64 over its
//            slice of the columns for every cluster k; centres are read from shared memory
//            as warp-wide broadcasts
//   combine  warp 0: fixed-order sum of the column-slice partials, sqrt, NumPy-style argmin
//            (first minimum, first NaN wins), new assignment, omega = w or 1-w, stable
//            grouping of the tile's rows by cluster
//   phase 2  centroid sums: thread t owns columns t, t+256, ... and adds omega*x for the
//            tile's rows cluster by cluster, in row order -> every accumulator is a plain
//            sequential float64 sum, bit-reproducible
// Two drivers share the sweep: `kmeans_groups_kernel` (one persistent CTA per independent
// problem, whole iteration loop on the device, no host sync) and `kmeans_sweep_kernel`
// (one CTA per row chunk of a large problem; partials are summed in fixed chunk order by
// `kmeans_reduce_kernel`, optionally all-reduced across GPUs, then `kmeans_update_kernel`).
#include "common.cuh"

namespace spalign {
namespace {

constexpr int KM_THREADS = 256;
constexpr int KMAX = 8;
constexpr int NS = 8;  // column slots per thread -> D + 2 <= NS * KM_THREADS

struct KmArgs {
  const void* X;
  int64_t ldx;        // row stride in elements
  int pos_mode;       // 1: two virtual columns (x, y) cell indices
  int pos_w;
  int64_t pos_period;
  int64_t pos_row0;   // global index of row 0 (multi-GPU shards)
  const double* w;
  int D;              // columns incl. virtual ones
  int Dr;             // stored columns
  int Dc;             // centre row stride (doubles)
  int K;
  int srow;           // shared-memory row stride in bytes
  int copy16;         // 16-byte chunks copied per row
};

struct KmSmem {
  char* buf[2];
  double* cen;     // [K][Dc]
  double* part;    // [KM_THREADS][KMAX]
  double* om;      // [TR] omega by sorted position
  int* order;      // [TR]
  int* start;      // [KMAX+1]
  double* wsum;    // [KMAX]
  double* cnt;     // [KMAX]
  double* red;     // [8][KMAX] block reduction scratch of the exact pass
  float* cen32;    // [K][Dc] centres rounded to fp32 (screening pass)
  float* cnorm;    // [KMAX] upper bound of ||c_k||
  int* anew;       // [TR] new assignment per tile row (-1: undecided)
  int* amb;        // [TR] tile rows that need the exact float64 pass
  int* namb;       // [1]
  int* changed;    // [1]
};

__host__ __device__ inline size_t km_smem_bytes(int TR, int srow, int K, int Dc) {
  size_t b = 0;
  b += (size_t)2 * TR * srow + 32;
  b += (size_t)K * Dc * sizeof(double);
  b += (size_t)KM_THREADS * KMAX * sizeof(double);
  b += (size_t)TR * sizeof(double);
  b += (size_t)TR * sizeof(int);
  b += (KMAX + 1) * sizeof(int) + 2 * KMAX * sizeof(double) + 64;
  b += (size_t)8 * KMAX * sizeof(double) + (size_t)K * Dc * sizeof(float) + KMAX * sizeof(float);
  b += (size_t)2 * TR * sizeof(int) + 64;
  return b + 128;
}

__device__ inline void km_carve(KmSmem& s, char* base, int TR, int srow, int K, int Dc) {
  size_t o = 0;
  s.buf[0] = base + o; o += (size_t)TR * srow;
  s.buf[1] = base + o; o += (size_t)TR * srow + 32;
  o = (o + 15) & ~(size_t)15;
  s.cen = reinterpret_cast<double*>(base + o); o += (size_t)K * Dc * sizeof(double);
  s.part = reinterpret_cast<double*>(base + o); o += (size_t)KM_THREADS * KMAX * sizeof(double);
  s.om = reinterpret_cast<double*>(base + o); o += (size_t)TR * sizeof(double);
  s.wsum = reinterpret_cast<double*>(base + o); o += KMAX * sizeof(double);
  s.cnt = reinterpret_cast<double*>(base + o); o += KMAX * sizeof(double);
  s.red = reinterpret_cast<double*>(base + o); o += (size_t)8 * KMAX * sizeof(double);
  s.cen32 = reinterpret_cast<float*>(base + o); o += (size_t)K * Dc * sizeof(float);
  s.cnorm = reinterpret_cast<float*>(base + o); o += KMAX * sizeof(float);
  s.anew = reinterpret_cast<int*>(base + o); o += (size_t)TR * sizeof(int);
  s.amb = reinterpret_cast<int*>(base + o); o += (size_t)TR * sizeof(int);
  s.namb = reinterpret_cast<int*>(base + o); o += 16;
  s.order = reinterpret_cast<int*>(base + o); o += (size_t)TR * sizeof(int);
  s.start = reinterpret_cast<int*>(base + o); o += (KMAX + 1) * sizeof(int);
  s.changed = reinterpret_cast<int*>(base + o);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename XT>
__device__ __forceinline__ void issue_tile(const KmArgs& a, char* buf, int64_t row0, int nvalid) {
  const char* src = reinterpret_cast<const char*>(a.X);
  const int total = nvalid * a.copy16;
  for (int i = threadIdx.x; i < total; i += KM_THREADS) {
    const int r = i / a.copy16, c = i - r * a.copy16;
    cp_async16(buf + (size_t)r * a.srow + (size_t)c * 16,
               src + ((size_t)(row0 + r) * a.ldx) * sizeof(XT) + (size_t)c * 16);
  }
}

template <typename XT, int VE>
__device__ __forceinline__ void load_chunk(const char* p, double* xv);
template <>
__device__ __forceinline__ void load_chunk<float, 4>(const char* p, double* xv) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  xv[0] = (double)v.x; xv[1] = (double)v.y; xv[2] = (double)v.z; xv[3] = (double)v.w;
}
template <>
__device__ __forceinline__ void load_chunk<double, 2>(const char* p, double* xv) {
  const double2 v = *reinterpret_cast<const double2*>(p);
  xv[0] = v.x; xv[1] = v.y;
}

// virtual position columns of global row n (direct_clustering.py:300-303): (x, y) cell index
__device__ __forceinline__ void virtual_pos(const KmArgs& a, int64_t row, double* px, double* py) {
  const int64_t n = (a.pos_row0 + row) % a.pos_period;
  *px = (double)(n % a.pos_w);
  *py = (double)(n / a.pos_w);
}

// NumPy argmin over K doubles: first minimum; a NaN is minimal and the first NaN wins
__device__ __forceinline__ int np_argmin(const double* d, int K) {
  double best = d[0];
  int idx = 0;
  if (best != best) return 0;
#pragma unroll
  for (int k = 1; k < KMAX; ++k) {
    if (k >= K) break;
    if (!(d[k] >= best)) {
      best = d[k];
      idx = k;
      if (best != best) break;
    }
  }
  return idx;
}

// One pass over rows [row_begin, row_end).  mode 0: keep `assign`, omega = 1 (init means).
// mode 1: reassign against s.cen, omega = w / 1-w.  acc[k][slot] accumulates column
// slot*256+t of cluster k; columns D and D+1 are sum(omega) and the member count.
template <typename XT, int TR>
__device__ __forceinline__ void km_sweep(const KmArgs& a, KmSmem& s, int64_t row_begin, int64_t row_end, int mode,
                         int32_t* __restrict__ assign, double (&acc)[KMAX][NS]) {
  constexpr int NPART = KM_THREADS / TR;
  constexpr int VE = 16 / (int)sizeof(XT);
  const int t = threadIdx.x;
  const int K = a.K, Dr = a.Dr, D = a.D;
  const int64_t N = row_end - row_begin;
  const int ntiles = (int)((N + TR - 1) / TR);
  if (ntiles == 0) return;
  const int row = t % TR, part = t / TR;
  const int nch = (Dr + VE - 1) / VE;
  const int ch0 = (int)((long long)part * nch / NPART);
  const int ch1 = (int)((long long)(part + 1) * nch / NPART);

  issue_tile<XT>(a, s.buf[0], row_begin, (int)min((int64_t)TR, N));
  cp_async_commit();
  for (int ti = 0; ti < ntiles; ++ti) {
    const int64_t trow0 = row_begin + (int64_t)ti * TR;
    const int nvalid = (int)min((int64_t)TR, row_end - trow0);
    if (ti + 1 < ntiles) {
      issue_tile<XT>(a, s.buf[(ti + 1) & 1], trow0 + TR,
                     (int)min((int64_t)TR, row_end - (trow0 + TR)));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const char* tile = s.buf[ti & 1];

    if (mode == 1) {
      if (sizeof(XT) == 4) {
        // ---- phase 1 (fp32 screening): partial squared distances over this thread's slice ----
        float dist[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) dist[k] = 0.f;
        if (row < nvalid) {
          const char* xr = tile + (size_t)row * a.srow;
          for (int ch = ch0; ch < ch1; ++ch) {
            const float4 xv = *reinterpret_cast<const float4*>(xr + (size_t)ch * 16);
            const int d = ch * 4;
            const bool full = d + 4 <= Dr;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
              if (k >= K) break;
              const float4 cv = *reinterpret_cast<const float4*>(s.cen32 + (size_t)k * a.Dc + d);
              if (full) {
                float df = xv.x - cv.x; dist[k] = fmaf(df, df, dist[k]);
                df = xv.y - cv.y; dist[k] = fmaf(df, df, dist[k]);
                df = xv.z - cv.z; dist[k] = fmaf(df, df, dist[k]);
                df = xv.w - cv.w; dist[k] = fmaf(df, df, dist[k]);
              } else {
                if (d + 0 < Dr) { const float df = xv.x - cv.x; dist[k] = fmaf(df, df, dist[k]); }
                if (d + 1 < Dr) { const float df = xv.y - cv.y; dist[k] = fmaf(df, df, dist[k]); }
                if (d + 2 < Dr) { const float df = xv.z - cv.z; dist[k] = fmaf(df, df, dist[k]); }
              }
            }
          }
        }
        float* partf = reinterpret_cast<float*>(s.part);
#pragma unroll
        for (int k = 0; k < KMAX; ++k) partf[(size_t)t * KMAX + k] = dist[k];
        __syncthreads();

        // ---- combine A: fp32 argmin with a rigorous error bound; undecided rows -> exact ----
        // |F_k - D_k| <= 2u*sqrt(D_k)*||c_k|| + (2u + gamma)*D_k + (2u||c_k||)^2, u = 2^-24
        // (centre rounding + subtraction rounding, then fp32 accumulation); constants below
        // are 2x / 256u generous.  A row is decided only if every other cluster stays
        // strictly farther after both bounds are applied.
        if (t < 32) {
          const int lane = t;
          const bool valid = lane < nvalid && lane < TR;
          bool undecided = false;
          if (valid) {
            float F[KMAX];
            float px = 0.f, py = 0.f;
            if (a.pos_mode) {
              double dpx, dpy;
              virtual_pos(a, trow0 + lane, &dpx, &dpy);
              px = (float)dpx;
              py = (float)dpy;
            }
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
              if (k >= K) break;
              float sum = 0.f;
              for (int p = 0; p < NPART; ++p) sum += partf[(size_t)(p * TR + lane) * KMAX + k];
              if (a.pos_mode) {
                const float dx = px - s.cen32[(size_t)k * a.Dc + Dr];
                const float dy = py - s.cen32[(size_t)k * a.Dc + Dr + 1];
                sum = fmaf(dx, dx, sum);
                sum = fmaf(dy, dy, sum);
              }
              F[k] = sum;
            }
            int j = 0;
            float best = F[0];
#pragma unroll
            for (int k = 1; k < KMAX; ++k) {
              if (k >= K) break;
              if (F[k] < best) { best = F[k]; j = k; }
            }
            const float c1 = 2.4e-7f, c2 = 1.6e-5f;
            bool certain = true;
            float lim = 0.f;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
              if (k >= K) break;
              const float cn = c1 * s.cnorm[k];
              const float B = cn * sqrtf(F[k]) + c2 * F[k] + cn * cn;
              if (k == j) lim = F[k] + B;
            }
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
              if (k >= K) break;
              const float cn = c1 * s.cnorm[k];
              const float B = cn * sqrtf(F[k]) + c2 * F[k] + cn * cn;
              if (k != j) certain = certain && (F[k] - B > lim);
            }
            certain = certain && (lim == lim) && (lim < 3.0e38f);
            s.anew[lane] = certain ? j : -1;
            undecided = !certain;
          }
          const unsigned um = __ballot_sync(0xffffffffu, undecided);
          if (undecided) s.amb[__popc(um & ((1u << lane) - 1u))] = lane;
          if (lane == 0) *s.namb = __popc(um);
        }
        __syncthreads();

        // ---- exact pass: float64 distances for the undecided rows, whole block per row ----
        const int namb = *s.namb;
        for (int ai = 0; ai < namb; ++ai) {
          const int r = s.amb[ai];
          const XT* xr = reinterpret_cast<const XT*>(tile + (size_t)r * a.srow);
          double p[KMAX];
#pragma unroll
          for (int k = 0; k < KMAX; ++k) p[k] = 0.0;
          for (int d = t; d < Dr; d += KM_THREADS) {
            const double xv = (double)xr[d];
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
              if (k >= K) break;
              const double df = xv - s.cen[(size_t)k * a.Dc + d];
              p[k] = fma(df, df, p[k]);
            }
          }
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            if (k >= K) break;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) p[k] += __shfl_xor_sync(0xffffffffu, p[k], o);
          }
          if ((t & 31) == 0) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) s.red[(size_t)(t >> 5) * KMAX + k] = p[k];
          }
          __syncthreads();
          if (t == 0) {
            double dd[KMAX];
            double px = 0.0, py = 0.0;
            if (a.pos_mode) virtual_pos(a, trow0 + r, &px, &py);
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
              if (k >= K) break;
              double sum = 0.0;
              for (int wq = 0; wq < KM_THREADS / 32; ++wq) sum += s.red[(size_t)wq * KMAX + k];
              if (a.pos_mode) {
                const double dx = px - s.cen[(size_t)k * a.Dc + Dr];
                const double dy = py - s.cen[(size_t)k * a.Dc + Dr + 1];
                sum = fma(dx, dx, sum);
                sum = fma(dy, dy, sum);
              }
              dd[k] = sqrt(sum);
            }
            s.anew[r] = np_argmin(dd, K);
          }
          __syncthreads();
        }
      } else {
        // ---- float64 rows: phase 1 entirely in float64 (no screening) ----
        double dist[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) dist[k] = 0.0;
        if (row < nvalid) {
          const char* xr = tile + (size_t)row * a.srow;
          for (int ch = ch0; ch < ch1; ++ch) {
            double xv[VE];
            load_chunk<XT, VE>(xr + (size_t)ch * 16, xv);
            const int d = ch * VE;
            const bool full = d + VE <= Dr;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
              if (k >= K) break;
              const double* ck = s.cen + (size_t)k * a.Dc + d;
              double cv[VE];
#pragma unroll
              for (int e = 0; e < VE; e += 2) {
                const double2 c2 = *reinterpret_cast<const double2*>(ck + e);
                cv[e] = c2.x;
                cv[e + 1] = c2.y;
              }
#pragma unroll
              for (int e = 0; e < VE; ++e) {
                if (full || d + e < Dr) {
                  const double df = xv[e] - cv[e];
                  dist[k] = fma(df, df, dist[k]);
                }
              }
            }
          }
        }
#pragma unroll
        for (int k = 0; k < KMAX; ++k) s.part[(size_t)t * KMAX + k] = dist[k];
        __syncthreads();
        if (t < 32 && t < nvalid && t < TR) {
          double d[KMAX];
          double px = 0.0, py = 0.0;
          if (a.pos_mode) virtual_pos(a, trow0 + t, &px, &py);
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            if (k >= K) break;
            double sum = 0.0;
            for (int p = 0; p < NPART; ++p) sum += s.part[(size_t)(p * TR + t) * KMAX + k];
            if (a.pos_mode) {
              const double dx = px - s.cen[(size_t)k * a.Dc + Dr];
              const double dy = py - s.cen[(size_t)k * a.Dc + Dr + 1];
              sum = fma(dx, dx, sum);
              sum = fma(dy, dy, sum);
            }
            d[k] = sqrt(sum);
          }
          s.anew[t] = np_argmin(d, K);
        }
        __syncthreads();
      }
    }

    // ---- combine B: adopt the new assignment, omega, stable grouping by cluster ----
    if (t < 32) {
      const int lane = t;
      const bool valid = lane < nvalid && lane < TR;
      int a_new = -1;
      double om = 0.0;
      int chg = 0;
      if (valid) {
        const int64_t grow = trow0 + lane;
        const int a_old = assign[grow];
        if (mode == 1) {
          a_new = s.anew[lane];
          chg = a_new != a_old;
          assign[grow] = a_new;
          const double wv = a.w[grow];
          om = a_new == 0 ? wv : 1.0 - wv;
        } else {
          a_new = a_old;
          om = 1.0;
        }
      }
      int pos = 0, base = 0;
      const unsigned lt = (1u << lane) - 1u;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k >= K) break;
        const unsigned m = __ballot_sync(0xffffffffu, valid && a_new == k);
        if (valid && a_new == k) pos = base + __popc(m & lt);
        if (lane == 0) s.start[k] = base;
        base += __popc(m);
      }
      if (lane == 0) s.start[K] = base;
      if (valid && a_new >= 0 && a_new < K) {
        s.order[pos] = lane;
        s.om[pos] = om;
      }
      const unsigned cm = __ballot_sync(0xffffffffu, chg != 0);
      if (lane == 0 && cm) *s.changed += __popc(cm);
    }
    __syncthreads();

    // ---- phase 2: centroid sums, cluster by cluster, rows in order ----
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k >= K) break;
      const int i1 = s.start[k + 1];
      for (int i = s.start[k]; i < i1; ++i) {
        const int r = s.order[i];
        const double om = s.om[i];
        const XT* xr = reinterpret_cast<const XT*>(tile + (size_t)r * a.srow);
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
          if (sl * KM_THREADS >= D + 2) break;
          const int d = sl * KM_THREADS + t;
          if (d < Dr) {
            acc[k][sl] = fma(om, (double)xr[d], acc[k][sl]);
          } else if (d < D) {  // virtual position columns
            double px, py;
            virtual_pos(a, trow0 + r, &px, &py);
            acc[k][sl] = fma(om, (d == Dr) ? px : py, acc[k][sl]);
          } else if (d == D) {
            acc[k][sl] += om;
          } else if (d == D + 1) {
            acc[k][sl] += 1.0;
          }
        }
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void zero_acc(double (&acc)[KMAX][NS]) {
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) acc[k][sl] = 0.0;
}

struct GroupArgs {
  KmArgs a;
  const int64_t* group_off;
  int32_t* assign;
  double* centers;  // may be null
  int32_t* iters;
  int32_t* status;
  int n_iter;
};

// centres = sums / sum(omega); returns true when some cluster has no member
__device__ __forceinline__ bool finalize_centers(const KmArgs& a, KmSmem& s, double (&acc)[KMAX][NS]) {
  const int t = threadIdx.x;
  const int D = a.D, K = a.K;
#pragma unroll
  for (int sl = 0; sl < NS; ++sl) {
    const int d = sl * KM_THREADS + t;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (d == D) s.wsum[k] = acc[k][sl];
      if (d == D + 1) s.cnt[k] = acc[k][sl];
    }
  }
  __syncthreads();
#pragma unroll
  for (int sl = 0; sl < NS; ++sl) {
    const int d = sl * KM_THREADS + t;
    if (d < D) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) s.cen[(size_t)k * a.Dc + d] = acc[k][sl] / s.wsum[k];
    }
  }
  bool empty = false;
  for (int k = 0; k < K; ++k) empty |= (s.cnt[k] == 0.0);
  __syncthreads();
  return empty;
}

// fp32 copies of the centres and ||c_k|| upper bounds for the screening pass
__device__ __forceinline__ void prepare_screen(const KmArgs& a, KmSmem& s) {
  const int t = threadIdx.x;
  for (int i = t; i < a.K * a.Dc; i += KM_THREADS) {
    const int d = i % a.Dc;
    s.cen32[i] = d < a.D ? (float)s.cen[i] : 0.f;
  }
  const int wq = t >> 5, lane = t & 31;
  if (wq < a.K) {
    double sum = 0.0;
    for (int d = lane; d < a.D; d += 32) {
      const double c = s.cen[(size_t)wq * a.Dc + d];
      sum = fma(c, c, sum);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s.cnorm[wq] = (float)sqrt(sum) * 1.000001f;
  }
  __syncthreads();
}

template <typename XT, int TR>
__global__ void __launch_bounds__(KM_THREADS, 1) kmeans_groups_kernel(GroupArgs g) {
  extern __shared__ __align__(128) char smem_raw[];
  KmSmem s;
  km_carve(s, smem_raw, TR, g.a.srow, g.a.K, g.a.Dc);
  const int grp = blockIdx.x;
  const int64_t r0 = g.group_off[grp], r1 = g.group_off[grp + 1];
  const int t = threadIdx.x;
  if (r1 <= r0) {
    if (t == 0) {
      g.iters[grp] = 0;
      g.status[grp] = SPALIGN_KM_CONVERGED;
    }
    return;
  }
  double acc[KMAX][NS];
  zero_acc(acc);
  if (t == 0) *s.changed = 0;
  __syncthreads();
  km_sweep<XT, TR>(g.a, s, r0, r1, 0, g.assign, acc);
  finalize_centers(g.a, s, acc);
  int it = 0, status = SPALIGN_KM_ITER_CAP;
  while (it < g.n_iter) {
    ++it;
    zero_acc(acc);
    if (t == 0) *s.changed = 0;
    if (sizeof(XT) == 4) prepare_screen(g.a, s);
    __syncthreads();
    km_sweep<XT, TR>(g.a, s, r0, r1, 1, g.assign, acc);
    const int changed = *s.changed;  // km_sweep ends with __syncthreads
    if (changed == 0) {
      status = SPALIGN_KM_CONVERGED;
      break;
    }
    if (finalize_centers(g.a, s, acc)) {
      status = SPALIGN_KM_EMPTY_CLUSTER;
      break;
    }
  }
  if (t == 0) {
    g.iters[grp] = it;
    g.status[grp] = status;
  }
  if (g.centers != nullptr) {
    double* out = g.centers + (size_t)grp * g.a.K * g.a.D;
    for (int i = t; i < g.a.K * g.a.D; i += KM_THREADS) {
      const int k = i / g.a.D, d = i - k * g.a.D;
      out[i] = s.cen[(size_t)k * g.a.Dc + d];
    }
  }
}

struct SweepArgs {
  KmArgs a;
  const int64_t* chunks;  // [n_chunks][3]
  const double* centers;  // [G][K][D]
  int mode;
  int32_t* assign;
  const int32_t* status;
  double* partials;       // [n_chunks][K*(D+2)+1]
};

template <typename XT, int TR>
__global__ void __launch_bounds__(KM_THREADS, 1) kmeans_sweep_kernel(SweepArgs g) {
  extern __shared__ __align__(128) char smem_raw[];
  KmSmem s;
  km_carve(s, smem_raw, TR, g.a.srow, g.a.K, g.a.Dc);
  const int ck = blockIdx.x;
  const int grp = (int)g.chunks[(size_t)ck * 3];
  const int64_t rb = g.chunks[(size_t)ck * 3 + 1], re = g.chunks[(size_t)ck * 3 + 2];
  if (g.status[grp] != SPALIGN_KM_RUNNING) return;
  const int t = threadIdx.x;
  const int K = g.a.K, D = g.a.D;
  if (g.mode == 1) {
    const double* c = g.centers + (size_t)grp * K * D;
    for (int i = t; i < K * D; i += KM_THREADS) {
      const int k = i / D, d = i - k * D;
      s.cen[(size_t)k * g.a.Dc + d] = c[i];
    }
  }
  double acc[KMAX][NS];
  zero_acc(acc);
  if (t == 0) *s.changed = 0;
  __syncthreads();
  if (g.mode == 1 && sizeof(XT) == 4) prepare_screen(g.a, s);
  km_sweep<XT, TR>(g.a, s, rb, re, g.mode, g.assign, acc);
  const size_t pv = (size_t)K * (D + 2) + 1;
  double* out = g.partials + (size_t)ck * pv;
#pragma unroll
  for (int sl = 0; sl < NS; ++sl) {
    const int d = sl * KM_THREADS + t;
    if (d < D + 2) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) out[(size_t)k * (D + 2) + d] = acc[k][sl];
    }
  }
  if (t == 0) out[pv - 1] = (double)*s.changed;
}

__global__ void __launch_bounds__(256)
kmeans_reduce_kernel(const double* __restrict__ partials, const int32_t* __restrict__ gco,
                     int pv, double* __restrict__ totals) {
  const int grp = blockIdx.y;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= pv) return;
  double sum = 0.0;
  for (int c = gco[grp]; c < gco[grp + 1]; ++c) sum += partials[(size_t)c * pv + j];
  totals[(size_t)grp * pv + j] = sum;
}

__global__ void __launch_bounds__(256)
kmeans_update_kernel(const double* __restrict__ totals, int D, int K, int mode, int n_iter,
                     double* centers, int32_t* iters, int32_t* status) {
  const int grp = blockIdx.x;
  if (status[grp] != SPALIGN_KM_RUNNING) return;
  const int pv = K * (D + 2) + 1;
  const double* tt = totals + (size_t)grp * pv;
  const int t = threadIdx.x;
  __shared__ int s_stop;
  if (t == 0) s_stop = (mode == 1 && tt[pv - 1] == 0.0) ? 1 : 0;
  __syncthreads();
  if (s_stop) {  // assignment unchanged: centres stay (batch_spalign_kmeans.py:158-159)
    if (t == 0) {
      iters[grp] += 1;
      status[grp] = SPALIGN_KM_CONVERGED;
    }
    return;
  }
  double* c = centers + (size_t)grp * K * D;
  for (int i = t; i < K * D; i += 256) {
    const int k = i / D, d = i - k * D;
    c[i] = tt[(size_t)k * (D + 2) + d] / tt[(size_t)k * (D + 2) + D];
  }
  if (t == 0 && mode == 1) {
    const int it = iters[grp] + 1;
    iters[grp] = it;
    bool empty = false;
    for (int k = 0; k < K; ++k) empty |= (tt[(size_t)k * (D + 2) + D + 1] == 0.0);
    if (empty) status[grp] = SPALIGN_KM_EMPTY_CLUSTER;
    else if (it >= n_iter) status[grp] = SPALIGN_KM_ITER_CAP;
  }
}

// seeded init for small groups: upper median by bitonic sort in shared memory
constexpr int INIT_MAX = 4096;
__global__ void __launch_bounds__(256)
kmeans_init_kernel(const double* __restrict__ w, const int64_t* __restrict__ group_off,
                   const int32_t* __restrict__ shuffled, const int64_t* __restrict__ shuf_off,
                   int32_t* assign, int32_t* m_out) {
  __shared__ double key[INIT_MAX];
  const int grp = blockIdx.x;
  const int64_t r0 = group_off[grp];
  const int n = (int)(group_off[grp + 1] - r0);
  const int t = threadIdx.x;
  if (n <= 0) {
    if (t == 0) m_out[grp] = 0;
    return;
  }
  int p2 = 1;
  while (p2 < n) p2 <<= 1;
  if (p2 > INIT_MAX) {  // too large for the shared-memory sort: the host initialises instead
    if (t == 0) m_out[grp] = -1;
    return;
  }
  for (int i = t; i < p2; i += 256) key[i] = i < n ? w[r0 + i] : __longlong_as_double(0x7ff0000000000000LL);
  __syncthreads();
  for (int k = 2; k <= p2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < p2; i += 256) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const double a = key[i], b = key[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            key[i] = b;
            key[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  const double thr = key[n / 2];
  __syncthreads();
  const int64_t s0 = shuf_off[grp];
  const int m_exp = (int)(shuf_off[grp + 1] - s0);
  // ordered rank of rows with w <= thr (reference: assign[cond] = idx)
  __shared__ int wcount[8];
  __shared__ int running;
  if (t == 0) running = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 256) {
    const int i = i0 + t;
    const bool cond = i < n && w[r0 + i] <= thr;
    const unsigned m = __ballot_sync(0xffffffffu, cond);
    if (lane_id() == 0) wcount[warp_id()] = __popc(m);
    __syncthreads();
    int before = running;
    for (int q = 0; q < warp_id(); ++q) before += wcount[q];
    const int rank = before + __popc(m & ((1u << lane_id()) - 1u));
    if (i < n) assign[r0 + i] = cond ? (rank < m_exp ? shuffled[s0 + rank] : 1) : 0;
    __syncthreads();
    if (t == 0) {
      int tot = 0;
      for (int q = 0; q < 8; ++q) tot += wcount[q];
      running += tot;
    }
    __syncthreads();
  }
  if (t == 0) m_out[grp] = running;
}

// ------------------------------------------------------------------------------------------
struct Plan {
  int TR;
  int srow;
  int copy16;
  int Dc;
  size_t smem;
};

bool make_plan(int x_dtype, int D, int Dr, int K, Plan* p) {
  const int es = x_dtype == SPALIGN_F32 ? 4 : 8;
  const int row_bytes = (int)align_up((size_t)Dr * es, 16);
  int srow = row_bytes;
  // rows 16 bytes apart modulo 128 -> conflict-free 16-byte reads with one row per lane
  while (srow % 128 != 16) srow += 16;
  p->srow = srow;
  p->copy16 = row_bytes / 16;
  p->Dc = (int)align_up((size_t)D, 4);
  const int trs[3] = {32, 16, 8};
  for (int i = 0; i < 3; ++i) {
    size_t b = km_smem_bytes(trs[i], srow, K, p->Dc);
    if (b <= 200 * 1024) {
      p->TR = trs[i];
      p->smem = b;
      return true;
    }
  }
  return false;
}

int fill_args(KmArgs* a, const Plan& p, const void* X, int x_dtype, int64_t ldx, int pos_mode,
              int pos_w, int64_t pos_period, int64_t pos_row0, const double* w, int D, int K) {
  SPALIGN_REQUIRE(X && w, "kmeans: NULL argument");
  SPALIGN_REQUIRE(x_dtype == SPALIGN_F32 || x_dtype == SPALIGN_F64, "kmeans: bad x_dtype");
  SPALIGN_REQUIRE(K >= 2 && K <= KMAX, "kmeans: K must be in [2, %d]", KMAX);
  SPALIGN_REQUIRE(pos_mode == 0 || pos_mode == 1, "kmeans: bad pos_mode");
  const int Dr = D - (pos_mode ? 2 : 0);
  SPALIGN_REQUIRE(Dr >= 1 && D + 2 <= NS * KM_THREADS, "kmeans: D out of range (max %d)",
                  NS * KM_THREADS - 2);
  const int es = x_dtype == SPALIGN_F32 ? 4 : 8;
  SPALIGN_REQUIRE((ldx * es) % 16 == 0 && ldx * es >= (int64_t)align_up((size_t)Dr * es, 16),
                  "kmeans: row stride must be a multiple of 16 bytes covering the padded row");
  SPALIGN_REQUIRE(reinterpret_cast<size_t>(X) % 16 == 0, "kmeans: X must be 16-byte aligned");
  SPALIGN_REQUIRE(!pos_mode || (pos_w > 0 && pos_period > 0), "kmeans: bad pos_w/pos_period");
  a->X = X; a->ldx = ldx; a->pos_mode = pos_mode; a->pos_w = pos_w ? pos_w : 1;
  a->pos_period = pos_period ? pos_period : 1; a->pos_row0 = pos_row0; a->w = w;
  a->D = D; a->Dr = Dr; a->Dc = p.Dc; a->K = K; a->srow = p.srow; a->copy16 = p.copy16;
  return SPALIGN_OK;
}

template <typename KernelT>
int set_smem(KernelT kernel, size_t bytes) {
  SPALIGN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)bytes));
  return SPALIGN_OK;
}

}  // namespace
}  // namespace spalign

using namespace spalign;

extern "C" size_t spalign_kmeans_groups_workspace_bytes(int D, int K, int G) {
  (void)D; (void)K; (void)G;
  return 256;  // the persistent kernel keeps all state on chip
}

#define KM_DISPATCH(KERNEL, ARGS, GRID)                                                        \
  do {                                                                                         \
    int rc__ = SPALIGN_OK;                                                                     \
    if (x_dtype == SPALIGN_F32) {                                                              \
      if (plan.TR == 32) { rc__ = set_smem(KERNEL<float, 32>, plan.smem); if (rc__) return rc__; \
        KERNEL<float, 32><<<GRID, KM_THREADS, plan.smem, stream>>>(ARGS); }                     \
      else if (plan.TR == 16) { rc__ = set_smem(KERNEL<float, 16>, plan.smem); if (rc__) return rc__; \
        KERNEL<float, 16><<<GRID, KM_THREADS, plan.smem, stream>>>(ARGS); }                     \
      else { rc__ = set_smem(KERNEL<float, 8>, plan.smem); if (rc__) return rc__;              \
        KERNEL<float, 8><<<GRID, KM_THREADS, plan.smem, stream>>>(ARGS); }                      \
    } else {                                                                                   \
      if (plan.TR == 32) { rc__ = set_smem(KERNEL<double, 32>, plan.smem); if (rc__) return rc__; \
        KERNEL<double, 32><<<GRID, KM_THREADS, plan.smem, stream>>>(ARGS); }                    \
      else if (plan.TR == 16) { rc__ = set_smem(KERNEL<double, 16>, plan.smem); if (rc__) return rc__; \
        KERNEL<double, 16><<<GRID, KM_THREADS, plan.smem, stream>>>(ARGS); }                    \
      else { rc__ = set_smem(KERNEL<double, 8>, plan.smem); if (rc__) return rc__;             \
        KERNEL<double, 8><<<GRID, KM_THREADS, plan.smem, stream>>>(ARGS); }                     \
    }                                                                                          \
  } while (0)

extern "C" int spalign_kmeans_groups(const void* X, int x_dtype, int64_t ldx, int pos_mode,
                                     int pos_w, int64_t pos_period, const double* w, int D,
                                     int K, int n_iter, const int64_t* group_off, int G,
                                     int32_t* assign, double* centers, int32_t* iters,
                                     int32_t* status, void* workspace, size_t ws_bytes,
                                     spalign_stream_t stream_) {
  (void)workspace; (void)ws_bytes;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(group_off && assign && iters && status && G > 0 && n_iter >= 0,
                  "kmeans_groups: bad arguments");
  Plan plan;
  const int Dr = D - (pos_mode ? 2 : 0);
  if (!make_plan(x_dtype, D, Dr, K, &plan)) {
    set_error("kmeans_groups: D=%d does not fit shared memory", D);
    return SPALIGN_E_UNSUPPORTED;
  }
  GroupArgs g;
  int rc = fill_args(&g.a, plan, X, x_dtype, ldx, pos_mode, pos_w, pos_period, 0, w, D, K);
  if (rc) return rc;
  g.group_off = group_off; g.assign = assign; g.centers = centers; g.iters = iters;
  g.status = status; g.n_iter = n_iter;
  KM_DISPATCH(kmeans_groups_kernel, g, G);
  return check_launch("kmeans_groups");
}

extern "C" int spalign_kmeans_sweep(const void* X, int x_dtype, int64_t ldx, int pos_mode,
                                    int pos_w, int64_t pos_period, int64_t pos_row0,
                                    const double* w, int D, int K, const int64_t* chunks,
                                    int n_chunks, const double* centers, int mode,
                                    int32_t* assign, const int32_t* status, double* partials,
                                    spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(chunks && assign && status && partials && n_chunks > 0 &&
                      (mode == 0 || (mode == 1 && centers)),
                  "kmeans_sweep: bad arguments");
  Plan plan;
  const int Dr = D - (pos_mode ? 2 : 0);
  if (!make_plan(x_dtype, D, Dr, K, &plan)) {
    set_error("kmeans_sweep: D=%d does not fit shared memory", D);
    return SPALIGN_E_UNSUPPORTED;
  }
  SweepArgs g;
  int rc = fill_args(&g.a, plan, X, x_dtype, ldx, pos_mode, pos_w, pos_period, pos_row0, w, D, K);
  if (rc) return rc;
  g.chunks = chunks; g.centers = centers; g.mode = mode; g.assign = assign; g.status = status;
  g.partials = partials;
  KM_DISPATCH(kmeans_sweep_kernel, g, n_chunks);
  return check_launch("kmeans_sweep");
}

extern "C" int spalign_kmeans_reduce(const double* partials, const int32_t* group_chunk_off,
                                     int G, int D, int K, double* totals,
                                     spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(partials && group_chunk_off && totals && G > 0 && G <= 65535 && D > 0 &&
                      K >= 2 && K <= KMAX,
                  "kmeans_reduce: bad arguments");
  const int pv = K * (D + 2) + 1;
  kmeans_reduce_kernel<<<dim3((pv + 255) / 256, G), 256, 0, stream>>>(partials, group_chunk_off,
                                                                      pv, totals);
  return check_launch("kmeans_reduce");
}

extern "C" int spalign_kmeans_update(const double* totals, int G, int D, int K, int mode,
                                     int n_iter, double* centers, int32_t* iters,
                                     int32_t* status, spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(totals && centers && iters && status && G > 0 && D > 0 && K >= 2 && K <= KMAX,
                  "kmeans_update: bad arguments");
  kmeans_update_kernel<<<G, 256, 0, stream>>>(totals, D, K, mode, n_iter, centers, iters, status);
  return check_launch("kmeans_update");
}

extern "C" int spalign_kmeans_init(const double* w, const int64_t* group_off, int G,
                                   const int32_t* shuffled, const int64_t* shuf_off,
                                   int32_t* assign, int32_t* m_out, spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(w && group_off && shuffled && shuf_off && assign && m_out && G > 0,
                  "kmeans_init: bad arguments");
  kmeans_init_kernel<<<G, 256, 0, stream>>>(w, group_off, shuffled, shuf_off, assign, m_out);
  return check_launch("kmeans_init");
}
