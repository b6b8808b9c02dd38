// K3: prior-weighted k-means (replaces kmeans(), batch_spalign_kmeans.py:136-183).
//
// Building blocks (DESIGN.md section 4 has the measurements):
//   screening   fp32, on differences of squared distances to cluster 0 (one subtraction and K
//               FFMA per element); a row is decided only when a rigorous rounding-error bound
//               separates the winner (km_screen_partial / km_screen_decide: fp32 decision
//               with explicit error terms, then float64, then ...)
//   exact pass  ... float64 distances with sqrt and NumPy's argmin rule (first minimum, first
//               NaN wins) for the rows no bound can separate: exact ties, NaN centres
//   sums        float64 centroid sums, thread t owns columns 2t, 2t+1 (+512, ...), rows added
//               cluster by cluster in row order -> bit-reproducible; after the first full
//               iteration only rows that changed cluster are moved (running sums)
//   bounds      Hamerly upper/lower distance bounds per row, shifted by the centre drift; a
//               sweep only gathers and screens the rows the bounds cannot prove stable
// float64 is kept to the places that decide the result.  (DFMA runs at full rate on B200; it
// is the F2F.F64.F32 conversions that are slow -- tools/fp64_probe.cu -- and a float64
// distance pass is 2K flops per element on half the lanes of the fp32 screening.)
//
// Sweeps: km_sweep streams the rows of a chunk through shared memory in tiles (bulk copies on
// mbarriers, double buffered; modes 0/1: init sums / full iteration; mode 3: move the listed
// changed rows); km_sweep_sparse runs warp-independent pipelines over the rows a bounds pass
// left active.  Kernels: kmeans_sweep_kernel (one CTA per row chunk; the chunk of a group
// that arrives last reduces and updates -> one launch per iteration), kmeans_tail_kernel (one
// persistent CTA per group runs all remaining iterations on-chip), kmeans_groups_kernel (one
// persistent CTA per group, full sweeps; small problems), and the reduce / update kernels of
// the three-step form that leaves room for an all-reduce.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace spalign {
namespace {

constexpr int KM_THREADS = 256;
constexpr int KMAX = 8;
constexpr int KM_NBAR = 2 + 2 * (KM_THREADS / 32);
// columns outside the fp32 screening loop: <= 3 stored (KM_NTAIL) + 2 virtual
constexpr int KM_NT = 5;
// fz layout: gt[KMAX][KM_NT] | c0t[KM_NT] | ek[KMAX] | ak[KMAX] | c0tn
constexpr int FZ_GT = 0, FZ_C0T = KMAX * KM_NT, FZ_EK = FZ_C0T + KM_NT, FZ_AK = FZ_EK + KMAX,
              FZ_C0TN = FZ_AK + KMAX, KM_FZ = FZ_C0TN + 3;
constexpr int ACT_MAX = 1024;  // rows per chunk up to which stable rows are compacted away
// entries of the active-row list: chunk-relative row index plus, once the row has been screened,
// its new / old cluster and flags
constexpr int ACT_ROW = 0xffff;
constexpr int ACT_NEW_SHIFT = 16, ACT_OLD_SHIFT = 20;
constexpr int ACT_CHG = 1 << 24;   // the row changed cluster
constexpr int ACT_AMB = 1 << 25;   // the row needs the exact float64 pass

// diagnostics: [0] rows screened, [1] rows sent to the exact float64 pass
__device__ unsigned long long g_km_stats[2];
#ifdef KM_PROFILE
__device__ unsigned long long g_km_prof[8];
__device__ unsigned long long g_km_prof2[8];
__device__ unsigned long long g_km_prof4[16];
__device__ unsigned long long g_km_prof3[8];
__device__ unsigned long long g_km_slow;
__device__ unsigned long long g_km_trace[3 * 1024];  // per group: start ns, end ns, iterations
#define KM_TICK3(i) do { if (threadIdx.x == 0) { long long now__ = clock64(); prof3__[i] += now__ - last3__; last3__ = now__; } } while (0)
#define KM_TICK2(i) do { if (threadIdx.x == 0) { long long now__ = clock64(); prof2__[i] += now__ - last2__; last2__ = now__; } } while (0)
#define KM_TICK(i) do { if (threadIdx.x == 0) { long long now__ = clock64(); prof__[i] += now__ - last__; last__ = now__; } } while (0)
#else
#define KM_TICK(i) do {} while (0)
#define KM_TICK2(i) do {} while (0)
#define KM_TICK3(i) do {} while (0)
#endif

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
  int Dc;             // centre row stride (elements)
  int Dm;             // leading stored columns screened in fp32 (multiple of 4)
  int K;
  int srow;           // shared-memory row stride in bytes
  int copy16;         // 16-byte chunks copied per row
  int TR;             // rows per tile (power of two, <= 32)
  int logTR;
  float* ub;          // [N] Hamerly upper bound: distance to the assigned centre (may be NULL)
  float* lb;          // [N] Hamerly lower bound: distance to the closest other centre
  int Kc;             // clusters the kernel variant is unrolled for (>= K)
  int part_bytes;     // size of the phase-1 partial-sum scratch
  int buf_rows;       // rows per half of the tile / row-buffer region (>= TR)
  int xcols;          // last stored columns (<= 2) summed with the per-cluster scalars, not in phase 2
};

// shared-memory carve-up (all offsets from the dynamic smem base)
struct KmSmem {
  char* buf0;       // 2 tiles of TR*srow bytes
  int tile_bytes;
  double* cen;      // [K][Dc]
  float* cen32;     // [K][Dc] centres rounded to fp32 (screening pass)
  char* part;       // [KM_THREADS][KMAX] partial distances (float or double)
  double* red;      // [8][KMAX] block reduction scratch of the exact pass
  double* om;       // [8 warps][32] omega by sorted position (per-warp copy)
  double* extra;    // [KMAX][4] sum(omega), count, sum(omega*px), sum(omega*py)
  float* cnorm;     // [KMAX] upper bound of ||g_k|| (fp32 screening)
  double* hk;       // [KMAX] screening constants (hk[0]: magnitude scale)
  unsigned short* order;  // [8 warps][32] tile rows grouped by cluster (per-warp copy)
  int* anew;        // [TR] new assignment per tile row (-1: undecided)
  int* amb;         // [TR] tile rows that need the exact float64 pass
  int* namb;        // [1]
  int* changed;     // [1]
  unsigned long long* bar;  // [2] mbarrier per tile buffer, then [8 warps][2] per-warp row buffers
  int* act;         // [ACT_MAX] chunk-relative indices of the rows that must be re-examined
  float* dk;        // [KMAX] centre drift of the last update (rounded up)
  float* dexcl;     // [KMAX] largest drift among the OTHER centres
  int* nact;        // [1]
  float* fz;        // [KM_FZ] fp32 constants of the fast decision path (see prepare_screen)
};

__host__ __device__ inline size_t km_carve(KmSmem* s, char* base, int TR, int srow, int K,
                                           int Dc, int Kc, int part_bytes, int buf_rows = 0) {
  size_t o = 0;
  const size_t tile = (size_t)TR * srow;
  if (s) { s->buf0 = base; s->tile_bytes = (int)tile; }
  // two tiles of TR rows; the sparse sweep's per-warp row buffers (8 warps x 2 x R rows) share
  // the region and may need more
  o += 2 * (buf_rows > TR ? (size_t)buf_rows * srow : tile) + 32;
  o = (o + 15) & ~(size_t)15;
  if (s) s->cen = reinterpret_cast<double*>(base + o);
  o += (size_t)K * Dc * sizeof(double);
  if (s) s->cen32 = reinterpret_cast<float*>(base + o);
  o += (size_t)Kc * Dc * sizeof(float);
  o = (o + 15) & ~(size_t)15;
  if (s) s->part = base + o;
  o += (size_t)part_bytes;
  o = (o + 15) & ~(size_t)15;
  if (s) s->red = reinterpret_cast<double*>(base + o);
  o += (size_t)8 * KMAX * sizeof(double);
  if (s) s->om = reinterpret_cast<double*>(base + o);
  o += (size_t)8 * 32 * sizeof(double);
  if (s) s->extra = reinterpret_cast<double*>(base + o);
  o += (size_t)KMAX * 4 * sizeof(double);
  if (s) s->hk = reinterpret_cast<double*>(base + o);
  o += KMAX * sizeof(double);
  if (s) s->cnorm = reinterpret_cast<float*>(base + o);
  o += KMAX * sizeof(float);
  if (s) s->order = reinterpret_cast<unsigned short*>(base + o);
  o += 8 * 32 * sizeof(unsigned short);
  if (s) s->anew = reinterpret_cast<int*>(base + o);
  o += 32 * sizeof(int);
  if (s) s->amb = reinterpret_cast<int*>(base + o);
  o += 32 * sizeof(int);
  if (s) s->namb = reinterpret_cast<int*>(base + o);
  o += sizeof(int);
  if (s) s->changed = reinterpret_cast<int*>(base + o);
  o += sizeof(int);
  o = (o + 15) & ~(size_t)15;
  if (s) s->bar = reinterpret_cast<unsigned long long*>(base + o);
  o += KM_NBAR * sizeof(unsigned long long);
  if (s) s->act = reinterpret_cast<int*>(base + o);
  o += ACT_MAX * sizeof(int);
  if (s) s->dk = reinterpret_cast<float*>(base + o);
  o += KMAX * sizeof(float);
  if (s) s->dexcl = reinterpret_cast<float*>(base + o);
  o += KMAX * sizeof(float);
  if (s) s->nact = reinterpret_cast<int*>(base + o);
  o += 4 * sizeof(int);
  if (s) s->fz = reinterpret_cast<float*>(base + o);
  o += KM_FZ * sizeof(float);
  return o + 64;
}

// ---- TMA-style bulk copies (cp.async.bulk, SASS UBLKCP) with an mbarrier per tile buffer ----
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes,
                                         unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  // bounded spin: a byte-count mismatch must fault, never hang the GPU
  for (unsigned spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 28)) __trap();
}

__device__ __forceinline__ void km_init_barriers(const KmSmem& s) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < KM_NBAR; ++i) mbar_init(s.bar + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
}

// warp 0 issues the tile: lane r copies row r with one bulk copy into the padded shared-memory
// layout; lane 0 first posts the expected byte count of the whole tile
template <typename XT>
__device__ __forceinline__ void issue_tile(const KmArgs& a, char* buf, unsigned long long* bar,
                                           int64_t row0, int nvalid, const int* rows = nullptr) {
  const int lane = threadIdx.x & 31;
  const char* src = reinterpret_cast<const char*>(a.X);
  if (rows != nullptr) {  // gathered tile: row r of the tile is chunk row rows[r]
    const unsigned row_bytes = (unsigned)a.copy16 * 16u;
    if (lane == 0) mbar_expect_tx(bar, row_bytes * (unsigned)nvalid);
    __syncwarp();
    for (int r = lane; r < nvalid; r += 32)
      bulk_g2s(buf + (size_t)r * a.srow,
               src + ((size_t)(row0 + (rows[r] & ACT_ROW)) * a.ldx) * sizeof(XT), row_bytes, bar);
    return;
  }
  if ((size_t)a.ldx * sizeof(XT) == (size_t)a.srow) {
    // the global row stride equals the padded shared-memory stride (e.g. 516-float descriptor
    // rows): the whole tile is one contiguous block -> a single bulk copy
    if (lane == 0) {
      const unsigned bytes = (unsigned)a.srow * (unsigned)nvalid;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(buf, src + (size_t)row0 * a.srow, bytes, bar);
    }
    return;
  }
  const unsigned row_bytes = (unsigned)a.copy16 * 16u;
  if (lane == 0) mbar_expect_tx(bar, row_bytes * (unsigned)nvalid);
  __syncwarp();
  for (int r = lane; r < nvalid; r += 32)
    bulk_g2s(buf + (size_t)r * a.srow, src + ((size_t)(row0 + r) * a.ldx) * sizeof(XT), row_bytes,
             bar);
}

// virtual position columns of global row n (direct_clustering.py:300-303): (x, y) cell index
__device__ __forceinline__ void virtual_pos(const KmArgs& a, int64_t row, double* px, double* py) {
  const int64_t n = (a.pos_row0 + row) % a.pos_period;
  *px = (double)(n % a.pos_w);
  *py = (double)(n / a.pos_w);
}

// NumPy argmin over K doubles: first minimum; a NaN is minimal and the first NaN wins
template <int KT>
__device__ __forceinline__ int np_argmin(const double (&d)[KT], int K) {
  double best = d[0];
  int idx = 0;
  bool done = best != best;
#pragma unroll
  for (int k = 1; k < KT; ++k) {
    if (k < K && !done && !(d[k] >= best)) {
      best = d[k];
      idx = k;
      done = best != best;
    }
  }
  return idx;
}

// fp32 screening, part 1: the warp's R shared-memory rows xw, xw + srow, ... against all
// clusters.  Works on DIFFERENCES of squared distances to cluster 0.  With df = x - c_0:
//   d_0^2 = sum df^2,   d_k^2 - d_0^2 = sum_d g_k[d]*df[d] + e_k,
//   g_k = 2(c_0 - c_k),  e_k = ||c_0 - c_k||^2,
// so a row costs one subtraction and K FFMA per element instead of 2K operations, and the
// argmin over {0, delta_1, ...} is the argmin over the distances.  The first a.Dm columns go
// through fp32; the few remaining stored columns (for the 514-column descriptors: the centroid
// coordinates, the large-magnitude ones) and the virtual columns are added in float64 by
// km_screen_decide.  Returns, in lane v < R*KT, the sum for row v / KT and cluster v % KT
// (cluster 0: d_0^2, cluster k: the dot product with g_k), summed in a fixed order.
template <int KT, int R>
__device__ __forceinline__ float km_screen_partial(const KmArgs& a, const KmSmem s,
                                                   const char* xw) {
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  constexpr int NV = R * KT;      // partial sums per lane
  constexpr int LPV = 32 / NV;    // lanes that share the final sum of one value
  static_assert(NV <= 32 && 32 % NV == 0, "R*KT must divide 32");
  float a1[R][KT];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < KT; ++k) a1[r][k] = 0.f;
  const int main_d = a.Dm;
#pragma unroll(R == 4 ? 2 : 1)
  for (int d = lane * 4; d < main_d; d += 128) {
    const float4 c0v = *reinterpret_cast<const float4*>(s.cen32 + d);
    float4 df[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 xv = *reinterpret_cast<const float4*>(xw + (size_t)r * a.srow + (size_t)d * 4);
      df[r].x = xv.x - c0v.x; df[r].y = xv.y - c0v.y;
      df[r].z = xv.z - c0v.z; df[r].w = xv.w - c0v.w;
      a1[r][0] = fmaf(df[r].x, df[r].x, a1[r][0]);
      a1[r][0] = fmaf(df[r].y, df[r].y, a1[r][0]);
      a1[r][0] = fmaf(df[r].z, df[r].z, a1[r][0]);
      a1[r][0] = fmaf(df[r].w, df[r].w, a1[r][0]);
    }
#pragma unroll
    for (int k = 1; k < KT; ++k) {
      const float4 gv = *reinterpret_cast<const float4*>(s.cen32 + (size_t)k * a.Dc + d);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        a1[r][k] = fmaf(df[r].x, gv.x, a1[r][k]);
        a1[r][k] = fmaf(df[r].y, gv.y, a1[r][k]);
        a1[r][k] = fmaf(df[r].z, gv.z, a1[r][k]);
        a1[r][k] = fmaf(df[r].w, gv.w, a1[r][k]);
      }
    }
  }
  // warp-level sum of the NV values through a padded shared-memory transpose (fixed order)
  float* pw = reinterpret_cast<float*>(s.part) + (size_t)wq * (NV * 33);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < KT; ++k) pw[(r * KT + k) * 33 + lane] = a1[r][k];
  __syncwarp();
  float tot = 0.f;
  {
    const int v = lane % NV, seg = lane / NV;
    const float* src = pw + v * 33 + seg * (32 / LPV);
#pragma unroll
    for (int j = 0; j < 32 / LPV; ++j) tot += src[j];
#pragma unroll
    for (int o = NV; o < 32; o <<= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  }
  return tot;
}

// fp32 screening, part 2: one lane decides one row from its KT screening sums F, the row's
// stored columns past a.Dm (xt, at most KM_NTAIL) and its global row index gr.  Returns the
// proven nearest cluster, or -1 when the rounding-error bound cannot separate the candidates
// (the row then needs the exact float64 pass).  Decided rows refresh their Hamerly bounds.
// Rounding of the fp32 part (u = 2^-24): c_0 and g are rounded once, df once, and the
// accumulation depth is below 100 (4 per 128 columns per lane, 16 in the transpose, 2
// shuffles), so with d0 = ||df|| over the fp32 columns
//   |dot_k - exact| <= (103u*d0 + u*||c_0||) * ||g_k||,
//   |F_0  - exact| <= 102u*d0^2 + 2u*d0*||c_0|| + (u*||c_0||)^2.
// We use eta = 2^-17 = 128u for all of them, norms rounded up, plus a floor of 2^-30 of the
// magnitude scale so that float64-rounding-sized gaps stay undecided.  A row is decided only
// when every other cluster stays strictly farther after both bounds; NaN/inf never decide.
constexpr int KM_NTAIL = 3;
template <int KT>
__device__ __forceinline__ int km_screen_decide(const KmArgs& a, const KmSmem s,
                                                const float (&F)[KT],
                                                const float (&xt)[KM_NTAIL], int64_t gr,
                                                bool bounds) {
  const int K = a.K, Dr = a.Dr;
  {
    // ---- fast path, all fp32, every rounding accounted for ----
    // With x'_c = x_c - c0_c over the tail columns, delta_k = d_k^2 - d_0^2 is
    //   F[k] + e_k + sum_c gt[k][c] x'_c,          tail share of d_0^2: sum_c x'_c^2.
    // Evaluated in fp32 the deviation from the float64 evaluation below is at most
    //   eps_k = 2^-20 M_k + 2^-23 ak[k],   M_k = |F[k]| + |e_k| + sum_c |gt[k][c] x'_c|
    // (roundings of c0_c, gt, e_k and of the <= 7 additions), so a gap that exceeds the
    // screening bounds by eps_k + eps_j as well is also a gap of the float64 test, with the
    // same winner.  Rows that fail here take the float64 path; NaN/inf fail every comparison.
    float xp[KM_NT];
    float tail0 = 0.f;
#pragma unroll
    for (int c = 0; c < KM_NT; ++c) xp[c] = 0.f;
#pragma unroll
    for (int c = 0; c < KM_NTAIL; ++c)
      if (a.Dm + c < Dr) xp[c] = xt[c] - s.fz[FZ_C0T + c];
    if (a.pos_mode) {
      double px, py;
      virtual_pos(a, gr, &px, &py);
      const int c = Dr - a.Dm;
#pragma unroll
      for (int q = 0; q < KM_NTAIL + 1; ++q) {
        if (q == c) {
          xp[q] = (float)px - s.fz[FZ_C0T + q];
          xp[q + 1] = (float)py - s.fz[FZ_C0T + q + 1];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < KM_NT; ++c) tail0 = fmaf(xp[c], xp[c], tail0);
    float dl[KT], ep[KT], Bf[KT];
    dl[0] = 0.f; ep[0] = 0.f; Bf[0] = 0.f;
    const float d0 = sqrtf(F[0]) * 1.0001f;
    const float c0n = s.cnorm[0];
#pragma unroll
    for (int k = 1; k < KT; ++k) {
      const float ek = s.fz[FZ_EK + k];
      float v = F[k] + ek, m = fabsf(F[k]) + fabsf(ek);
#pragma unroll
      for (int c = 0; c < KM_NT; ++c) {
        const float p = s.fz[FZ_GT + k * KM_NT + c] * xp[c];
        v += p;
        m += fabsf(p);
      }
      dl[k] = k < K ? v : 3.0e38f;
      ep[k] = k < K ? 9.6e-7f * m + 1.2e-7f * s.fz[FZ_AK + k] : 0.f;
      Bf[k] = 7.64e-6f * ((d0 + c0n) * s.cnorm[k]);
    }
    int j = 0;
    float best = 0.f;
#pragma unroll
    for (int k = 1; k < KT; ++k)
      if (dl[k] < best) { best = dl[k]; j = k; }
    float epj = 0.f, Bj = 0.f;
#pragma unroll
    for (int k = 1; k < KT; ++k)
      if (k == j) { epj = ep[k]; Bj = Bf[k]; }
    const float D0 = F[0] + tail0;
    const float E0 = 7.64e-6f * (F[0] + 2.f * (d0 * c0n)) + (c0n * 6.0e-8f) * (c0n * 6.0e-8f);
    const float floor_ = 9.5e-10f * (D0 + __double2float_ru(s.hk[0]));
    // deviation of D0 from d_0^2 beyond E0: roundings of x'_c (c0_c) and of the sums
    const float epD = 6.0e-7f * D0 + 1.2e-7f * sqrtf(tail0) * s.fz[FZ_C0TN] +
                      1.0e-12f * s.fz[FZ_C0TN] * s.fz[FZ_C0TN];
    bool ok = (F[0] < 3.0e38f) && (j < K);
    float second = 3.0e38f;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      if (k != j) {
        const float rhs = Bf[k] + Bj + floor_ + ep[k] + epj;
        ok = ok && ((dl[k] - best) * 0.999999f > rhs * 1.00001f);
        second = fminf(second, dl[k] - ep[k] - Bf[k]);
      }
    }
    if (ok) {
      if (bounds) {
        const float slack = E0 + epD;
        const float u2 = (D0 + best) + (slack + epj + Bj);
        const float l2 = (D0 + second) - slack;
        const float pad = 1.0e-6f * (D0 + fabsf(best) + fabsf(second));
        a.ub[gr] = __fmul_ru(__fsqrt_ru(fmaxf(u2 + pad, 0.f)), 1.000001f);
        a.lb[gr] = __fmul_rd(__fsqrt_rd(fmaxf(l2 - pad, 0.f)), 0.999999f);
      }
      return j;
    }
  }
#ifdef KM_PROFILE
  atomicAdd(&g_km_slow, 1ULL);
#endif
  double delta[KT];
  delta[0] = 0.0;
#pragma unroll
  for (int k = 1; k < KT; ++k) delta[k] = (double)F[k] - s.hk[k];
  // stored columns past a.Dm and the virtual columns, exactly, in float64
  double tail0 = 0.0;   // their share of d_0^2
  auto add_col = [&](double xv, int d) {
    const double c0 = s.cen[d];
    tail0 = fma(xv - c0, xv - c0, tail0);
#pragma unroll
    for (int k = 1; k < KT; ++k) {
      if (k < K) {
        const double ck = s.cen[(size_t)k * a.Dc + d];
        delta[k] = fma(c0 - ck, 2.0 * xv - c0 - ck, delta[k]);
      }
    }
  };
#pragma unroll
  for (int c = 0; c < KM_NTAIL; ++c)
    if (a.Dm + c < Dr) add_col((double)xt[c], a.Dm + c);
  if (a.pos_mode) {
    double px, py;
    virtual_pos(a, gr, &px, &py);
    add_col(px, Dr);
    add_col(py, Dr + 1);
  }
  int j = 0;
  double best = 0.0;
#pragma unroll
  for (int k = 1; k < KT; ++k)
    if (delta[k] < best) { best = delta[k]; j = k; }
  const float d0 = sqrtf(F[0]) * 1.0001f;
  const float c0n = s.cnorm[0];
  const double D0 = (double)F[0] + tail0;                       // d_0^2, all columns
  const double E0 = 7.63e-6 * ((double)F[0] + 2.0 * (double)(d0 * c0n)) +
                    (double)(c0n * 6.0e-8f) * (double)(c0n * 6.0e-8f);
  const double floor_ = 9.4e-10 * (D0 + s.hk[0]);               // 2^-30 * scale
  double B[KT];
  B[0] = 0.0;
#pragma unroll
  for (int k = 1; k < KT; ++k) B[k] = 7.63e-6 * (double)((d0 + c0n) * s.cnorm[k]);
  double Bj = 0.0;
#pragma unroll
  for (int k = 1; k < KT; ++k)
    if (k == j) Bj = B[k];
  bool certain = (best == best) && (F[0] < 3.0e38f) && (j < K);
  double second = 1.0e300;   // smallest provable squared distance to another centre
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    if (k != j) {
      certain = certain && (delta[k] - best > B[k] + Bj + floor_);
      second = fmin(second, D0 + delta[k] - E0 - B[k]);
    }
  }
  if (!certain) return -1;
  if (bounds) {
    // bounds on the distances: squared bounds rounded outwards to fp32, then directed sqrt
    const float u2 = __double2float_ru(fmax(D0 + best + E0 + Bj, 0.0));
    const float l2 = __double2float_rd(fmax(second, 0.0));
    a.ub[gr] = __fmul_ru(__fsqrt_ru(u2), 1.000001f);
    a.lb[gr] = __fmul_rd(__fsqrt_rd(l2), 0.999999f);
  }
  return j;
}

// stage the warp's screening sums (lane v < R*KT) and its rows' tail columns for the lanes
// that will decide the rows: stF[(slot0 + r)*KT + k], stX[(slot0 + r)*KM_NTAIL + c]
template <int KT, int R>
__device__ __forceinline__ void km_screen_stage(const KmArgs& a, const char* xw, float tot,
                                                float* stF, float* stX, int slot0) {
  const int lane = threadIdx.x & 31;
  if (lane < R * KT) stF[slot0 * KT + lane] = tot;
  const int ntail = a.Dr - a.Dm;
  if (lane < R * ntail) {
    const int r = lane / ntail, c = lane - r * ntail;
    stX[(slot0 + r) * KM_NTAIL + c] =
        reinterpret_cast<const float*>(xw + (size_t)r * a.srow)[a.Dm + c];
  }
}

// Hamerly bounds pass over the rows [row_begin, row_begin + N) of a chunk (mode 2): shift every
// row's bounds by the centre drift of the last update, keep the rows whose bounds no longer
// prove a stable assignment, compacted in row order into s.act.  Returns their number.
template <int KT>
__device__ __forceinline__ int km_bounds_pass(const KmArgs& a, const KmSmem s, int64_t row_begin,
                                              int N, const int32_t* __restrict__ assign,
                                              const double* cdelta) {
  const int t = threadIdx.x, lane = t & 31, wq = t >> 5;
  const int K = a.K;
  constexpr int NB = ACT_MAX / KM_THREADS;  // rows per thread: b * KM_THREADS + t
  // all global loads first (one round trip): the centre drifts and the rows' state
  float dk[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) dk[k] = k < K ? __double2float_ru(cdelta[k]) : 0.f;
  int ai[NB];
  float ub[NB], lb[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int i = b * KM_THREADS + t;
    ai[b] = -1; ub[b] = 0.f; lb[b] = 0.f;
    if (i < N) {
      const int64_t gr = row_begin + i;
      ai[b] = assign[gr];
      ub[b] = a.ub[gr];
      lb[b] = a.lb[gr];
    }
  }
  // largest and second largest drift: the drift of "every other centre" is the largest one,
  // or the second largest for the rows of the centre that moved most
  float m1 = 0.f, m2 = 0.f;
  int am = -1;
  bool bad = false;  // NaN drift (NaN centres) must keep every row active
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    if (k < K) {
      bad |= !(dk[k] == dk[k]);
      if (dk[k] > m1) { m2 = m1; m1 = dk[k]; am = k; }
      else if (dk[k] > m2) m2 = dk[k];
    }
  }
  unsigned msk[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int i = b * KM_THREADS + t;
    bool active = false;
    if (i < N) {
      const int64_t gr = row_begin + i;
      const bool okc = ai[b] >= 0 && ai[b] < K;
      float own = 0.f;
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (k == ai[b]) own = dk[k];
      const float oth = bad ? __int_as_float(0x7f800000) : (ai[b] == am ? m2 : m1);
      const float u2 = __fadd_ru(ub[b], okc ? own : 0.f);
      const float l2 = __fsub_rd(lb[b], okc ? oth : 0.f);
      a.ub[gr] = u2;
      a.lb[gr] = l2;
      active = !(okc && u2 < l2);
    }
    msk[b] = __ballot_sync(0xffffffffu, active);
  }
  __shared__ int wcnt[NB][KM_THREADS / 32];
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < NB; ++b) wcnt[b][wq] = __popc(msk[b]);
  }
  __syncthreads();
  int run = 0;  // active rows before (b, warp) in row order
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    int before = run;
#pragma unroll
    for (int q = 0; q < KM_THREADS / 32; ++q) {
      const int c = wcnt[b][q];
      if (q < wq) before += c;
      run += c;
    }
    if (msk[b] & (1u << lane))
      s.act[before + __popc(msk[b] & ((1u << lane) - 1u))] = b * KM_THREADS + t;
  }
  __syncthreads();
  return run;
}

// One pass over rows [row_begin, row_end).  mode 0: keep `assign`, omega = 1 (init means).
// mode 1: reassign against s.cen, omega = w / 1-w.  acc[k][sl][j] accumulates stored column
// 2*(sl*256+t)+j of cluster k; s.extra[k] receives sum(omega), count and the virtual columns.
template <typename XT, int KT, int NS2, int R>
__device__ __forceinline__ void km_sweep(const KmArgs& a, const KmSmem s, int64_t row_begin,
                                         int64_t row_end, int mode, int32_t* __restrict__ assign,
                                         double (&acc)[KT][NS2][2], unsigned& tile_base,
                                         int nrows_in = -1) {
  constexpr int VE = 16 / (int)sizeof(XT);
  constexpr bool kF32 = sizeof(XT) == 4;
  const int t = threadIdx.x;
  const int K = a.K, Dr = a.Dr, TR = a.TR;
  const int NPART = KM_THREADS >> a.logTR;
  const int64_t N = row_end - row_begin;
  const int row = t & (TR - 1), part = t >> a.logTR;
  const int nch = (Dr + VE - 1) / VE;
  const int ch0 = (int)((long long)part * nch / NPART);
  const int ch1 = (int)((long long)(part + 1) * nch / NPART);
  // Hamerly bounds (mode 2, fp32 rows): a row whose upper bound on the distance to its own
  // centre stays below its lower bound on the distance to every other centre, after both are
  // shifted by how far the centres moved, keeps its assignment -- it is neither loaded nor
  // screened.  The remaining rows of the chunk are compacted (in row order) and gathered.
  const bool bounds = kF32 && a.ub != nullptr && a.lb != nullptr;
  const bool compact = nrows_in >= 0;   // the caller ran km_bounds_pass: rows listed in s.act
  const int nrows = compact ? nrows_in : (int)N;
  const int ntiles = (nrows + TR - 1) / TR;
  // lane k of warp 0: running sum(omega), count, sum(omega*px), sum(omega*py) of cluster k
  double e_w = 0.0, e_n = 0.0, e_x = 0.0, e_y = 0.0;

  // tile `ti` lands in buffer ti&1 and completes phase (ti>>1)&1 of that buffer's mbarrier,
  // counted from `tile_base` (tiles this CTA has already streamed through the barriers)
  if (ntiles > 0 && t < 32)
    issue_tile<XT>(a, s.buf0 + (size_t)(tile_base & 1) * s.tile_bytes, s.bar + (tile_base & 1),
                   row_begin, min(TR, nrows), compact ? s.act : nullptr);
#define KM_ROW(r) (compact ? row_begin + (int64_t)(s.act[tb + (r)] & ACT_ROW) : row_begin + (int64_t)(tb + (r)))
#ifdef KM_PROFILE
  long long prof__[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long last__ = clock64();
#endif
  const int lane_ = t & 31, wq_ = t >> 5;
  unsigned short* my_order = s.order + wq_ * 32;      // per-warp copy of the tile's row grouping
  double* my_om = s.om + wq_ * 32;
  for (int ti = 0; ti < ntiles; ++ti) {
    const int tb = ti * TR;                          // index of the tile's first row in the
    const int nvalid = min(TR, nrows - tb);          // (possibly compacted) row list
    const unsigned gt = tile_base + (unsigned)ti;   // tile counter across sweeps
    // every warp keeps the old assignment / prior weight of tile row `lane` (L2 hits)
    // (mode 2 uses two lanes per row: lane r adds row r to its new cluster, lane TR + r
    // removes it from the old one)
    int pre_a = -1;
    double pre_w = 0.0;
    const int prow = lane_ & (TR - 1);
    int pre_n = -1;  // mode 3: new cluster, decided by the screening pass
    if (lane_ < 2 * TR && lane_ < 32 && prow < nvalid) {
      if (mode == 3) {
        const int e = s.act[tb + prow];
        pre_a = (e >> ACT_OLD_SHIFT) & 15;
        pre_n = (e >> ACT_NEW_SHIFT) & 15;
      } else {
        pre_a = assign[KM_ROW(prow)];
      }
      if (mode != 0) pre_w = a.w[KM_ROW(prow)];
    }
    __syncthreads();  // (A) all warps are past phase 2 of tile ti-1: its buffer is free
    if (t == 0) *s.namb = 0;  // (read by every warp before this barrier, filled after the next)
    KM_TICK(0);
    if (ti + 1 < ntiles && t < 32)  // prefetch the next tile into the buffer just released
      issue_tile<XT>(a, s.buf0 + (size_t)((gt + 1) & 1) * s.tile_bytes, s.bar + ((gt + 1) & 1),
                     compact ? row_begin : row_begin + tb + TR, min(TR, nrows - (tb + TR)),
                     compact ? s.act + tb + TR : nullptr);
    KM_TICK(5);
    mbar_wait(s.bar + (gt & 1), (gt >> 1) & 1);     // tile ti has landed
    const char* tile = s.buf0 + (size_t)(gt & 1) * s.tile_bytes;
    KM_TICK(1);

    if (mode == 1 || mode == 2) {
      if (kF32) {
        // ---- phase 1 (fp32 screening): one warp owns R rows, lanes stride the columns; the
        // sums are staged and warp 0 decides all rows of the tile at once ----
        __shared__ float stF[32 * KT];
        __shared__ float stX[32 * KM_NTAIL];
        {
          const int rbase = (t >> 5) * R;
          const char* xw = tile + (size_t)rbase * a.srow;
          const float tot = km_screen_partial<KT, R>(a, s, xw);
          km_screen_stage<KT, R>(a, xw, tot, stF, stX, rbase);
        }
        __syncthreads();
        if (t < nvalid) {
          float F[KT], xt[KM_NTAIL];
#pragma unroll
          for (int k = 0; k < KT; ++k) F[k] = stF[t * KT + k];
#pragma unroll
          for (int c = 0; c < KM_NTAIL; ++c) xt[c] = stX[t * KM_NTAIL + c];
          const int j = km_screen_decide<KT>(a, s, F, xt, KM_ROW(t), bounds);
          s.anew[t] = j;
          if (j < 0) s.amb[atomicAdd(s.namb, 1)] = t;
        }
        __syncthreads();

        // ---- exact pass: float64 distances for the undecided rows, whole block per row ----
        const int namb = *s.namb;
        if (t == 0) {
          atomicAdd(&g_km_stats[0], (unsigned long long)nvalid);
          if (namb) atomicAdd(&g_km_stats[1], (unsigned long long)namb);
        }
        for (int ai = 0; ai < namb; ++ai) {
          const int r = s.amb[ai];
          const XT* xr = reinterpret_cast<const XT*>(tile + (size_t)r * a.srow);
          double p[KT];
#pragma unroll
          for (int k = 0; k < KT; ++k) p[k] = 0.0;
          for (int d = t; d < Dr; d += KM_THREADS) {
            const double xv = (double)xr[d];
#pragma unroll
            for (int k = 0; k < KT; ++k) {
              if (k < K) {
                const double df = xv - s.cen[(size_t)k * a.Dc + d];
                p[k] = fma(df, df, p[k]);
              }
            }
          }
#pragma unroll
          for (int k = 0; k < KT; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) p[k] += __shfl_xor_sync(0xffffffffu, p[k], o);
          }
          if ((t & 31) == 0) {
#pragma unroll
            for (int k = 0; k < KT; ++k) s.red[(size_t)(t >> 5) * KMAX + k] = p[k];
          }
          __syncthreads();
          if (t == 0) {
            double dd[KT];
            double px = 0.0, py = 0.0;
            if (a.pos_mode) virtual_pos(a, KM_ROW(r), &px, &py);
#pragma unroll
            for (int k = 0; k < KT; ++k) {
              double sum = 0.0;
              if (k < K) {
                for (int wq = 0; wq < KM_THREADS / 32; ++wq) sum += s.red[(size_t)wq * KMAX + k];
                if (a.pos_mode) {
                  const double dx = px - s.cen[(size_t)k * a.Dc + Dr];
                  const double dy = py - s.cen[(size_t)k * a.Dc + Dr + 1];
                  sum = fma(dx, dx, sum);
                  sum = fma(dy, dy, sum);
                }
              }
              dd[k] = sqrt(sum);
            }
            const int jj = np_argmin<KT>(dd, K);
            s.anew[r] = jj;
            if (bounds) {  // exact distances: bounds with a 1e-6 relative cushion
              double sec = 1.0e300;
#pragma unroll
              for (int k = 0; k < KT; ++k)
                if (k < K && k != jj) sec = fmin(sec, dd[k]);
              const int64_t gr = KM_ROW(r);
              const bool okd = dd[jj] == dd[jj] && sec == sec;
              a.ub[gr] = okd ? __fmul_ru(__double2float_ru(dd[jj]), 1.000001f)
                             : __int_as_float(0x7f800000);
              a.lb[gr] = okd ? __fmul_rd(__double2float_rd(sec), 0.999999f) : 0.f;
            }
          }
          __syncthreads();
        }
      } else {
        // ---- float64 rows: phase 1 entirely in float64 (no screening) ----
        double dist[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) dist[k] = 0.0;
        if (row < nvalid) {
          const char* xr = tile + (size_t)row * a.srow;
          for (int ch = ch0; ch < ch1; ++ch) {
            const double2 xv = *reinterpret_cast<const double2*>(xr + (size_t)ch * 16);
            const int d = ch * 2;
#pragma unroll
            for (int k = 0; k < KT; ++k) {
              if (k < K) {
                const double2 cv = *reinterpret_cast<const double2*>(s.cen + (size_t)k * a.Dc + d);
                double df = xv.x - cv.x;
                dist[k] = fma(df, df, dist[k]);
                if (d + 1 < Dr) {
                  df = xv.y - cv.y;
                  dist[k] = fma(df, df, dist[k]);
                }
              }
            }
          }
        }
        double* partd = reinterpret_cast<double*>(s.part);
#pragma unroll
        for (int k = 0; k < KT; ++k) partd[(size_t)t * KT + k] = dist[k];
        __syncthreads();
        if (t < 32 && t < nvalid && t < TR) {
          double dd[KT];
          double px = 0.0, py = 0.0;
          if (a.pos_mode) virtual_pos(a, KM_ROW(t), &px, &py);
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            double sum = 0.0;
            if (k < K) {
              for (int p = 0; p < NPART; ++p) sum += partd[(size_t)(p * TR + t) * KT + k];
              if (a.pos_mode) {
                const double dx = px - s.cen[(size_t)k * a.Dc + Dr];
                const double dy = py - s.cen[(size_t)k * a.Dc + Dr + 1];
                sum = fma(dx, dx, sum);
                sum = fma(dy, dy, sum);
              }
            }
            dd[k] = sqrt(sum);
          }
          s.anew[t] = np_argmin<KT>(dd, K);
        }
        __syncthreads();
      }
    }

    KM_TICK(2);
    // ---- grouping (every warp, no block barrier): new assignment, omega, stable grouping of
    // the tile's rows by cluster into per-warp shared-memory lists ----
    int startk[KT + 1];
    {
      const int lane = lane_;
      const int half = lane >> a.logTR;          // 0: "add" entry, 1: "remove" entry (mode 2)
      const bool in_tile = half < 2 && prow < nvalid;
      bool valid = false;                        // this lane contributes a list entry
      int a_new = -1;                            // cluster of the entry
      double om = 0.0;                           // signed weight of the entry
      int chg = 0;
      if (in_tile) {
        if (mode == 0) {
          valid = half == 0;
          a_new = pre_a;
          om = 1.0;
        } else {
          const int an = mode == 3 ? pre_n : s.anew[prow];
          chg = an != pre_a;
          if (mode == 1) {                       // full sums: every row, new cluster
            valid = half == 0;
            a_new = an;
            om = an == 0 ? pre_w : 1.0 - pre_w;
          } else if (chg) {                      // incremental: only rows that moved
            valid = true;
            a_new = half == 0 ? an : pre_a;
            const double o = a_new == 0 ? pre_w : 1.0 - pre_w;
            om = half == 0 ? o : -o;
          }
        }
      }
      const int an_row = (mode != 0 && mode != 3 && in_tile) ? s.anew[prow] : pre_a;
      int pos = 0, base = 0;
      const unsigned lt = (1u << lane) - 1u;
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        startk[k] = base;
        if (k < K) {
          const unsigned m = __ballot_sync(0xffffffffu, valid && a_new == k);
          if (valid && a_new == k) pos = base + __popc(m & lt);
          base += __popc(m);
        }
      }
      startk[KT] = base;
      if (valid && a_new >= 0 && a_new < K) {
        my_order[pos] = (unsigned short)(prow | (half << 8));  // bit 8: the entry removes the row
        my_om[pos] = om;
      }
      const unsigned cm = __ballot_sync(0xffffffffu, chg != 0 && half == 0);
      __syncwarp();
      if (wq_ == 0) {  // warp 0 owns the assignment write-back
        if (chg && half == 0 && mode != 3) assign[KM_ROW(prow)] = an_row;
        if (lane == 0 && cm) *s.changed += __popc(cm);
      }
      if (wq_ == 1) {  // warp 1 owns the per-cluster scalars
        if (lane < K) {  // lane k: sum(omega), count, virtual columns; rows in order
          int i0 = 0, i1 = 0;
#pragma unroll
          for (int k = 0; k < KT; ++k)
            if (k == lane) { i0 = startk[k]; i1 = startk[k + 1]; }
          for (int i = i0; i < i1; ++i) {
            const double o = my_om[i];
            const int ent = my_order[i];
            e_w += o;
            e_n += (ent & 0x100) ? -1.0 : 1.0;
            if (a.pos_mode) {
              double px, py;
              virtual_pos(a, KM_ROW(ent & 0xff), &px, &py);
              e_x = fma(o, px, e_x);
              e_y = fma(o, py, e_y);
            } else if (a.xcols) {  // the last stored columns (see fill_args)
              const XT* xr = reinterpret_cast<const XT*>(tile + (size_t)(ent & 0xff) * a.srow);
              e_x = fma(o, (double)xr[Dr - a.xcols], e_x);
              if (a.xcols == 2) e_y = fma(o, (double)xr[Dr - 1], e_y);
            }
          }
        }
      }
    }

    KM_TICK(3);
    // ---- phase 2: centroid sums, cluster by cluster, rows in order ----
    const int Dp = Dr - a.xcols;  // columns summed here
#pragma unroll
    for (int sl = 0; sl < NS2; ++sl) {
      const int c0 = 2 * (sl * KM_THREADS + t);
      if (c0 + 1 < Dp) {
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          if (k < K) {
            const int i1 = startk[k + 1];
#pragma unroll 4
            for (int i = startk[k]; i < i1; ++i) {
              const double om = my_om[i];
              const XT* xr =
                  reinterpret_cast<const XT*>(tile + (size_t)(my_order[i] & 0xff) * a.srow);
              double x0, x1;
              if (kF32) {
                const float2 v = *reinterpret_cast<const float2*>(xr + c0);
                x0 = (double)v.x;
                x1 = (double)v.y;
              } else {
                const double2 v = *reinterpret_cast<const double2*>(xr + c0);
                x0 = v.x;
                x1 = v.y;
              }
              acc[k][sl][0] = fma(om, x0, acc[k][sl][0]);
              acc[k][sl][1] = fma(om, x1, acc[k][sl][1]);
            }
          }
        }
      } else if (c0 < Dp) {
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          if (k < K) {
            const int i1 = startk[k + 1];
            for (int i = startk[k]; i < i1; ++i) {
              const XT* xr =
                  reinterpret_cast<const XT*>(tile + (size_t)(my_order[i] & 0xff) * a.srow);
              acc[k][sl][0] = fma(my_om[i], (double)xr[c0], acc[k][sl][0]);
            }
          }
        }
      }
    }
    KM_TICK(4);
  }
  __syncthreads();
#ifdef KM_PROFILE
  if (threadIdx.x == 0) {  // [0]: compacted (mode 3) tiles, [1]: mode 0, [2]: mode 1
    unsigned long long* dst = compact ? g_km_prof : (mode == 0 ? g_km_prof4 : g_km_prof4 + 8);
    for (int i = 0; i < 6; ++i) atomicAdd(&dst[i], (unsigned long long)prof__[i]);
    atomicAdd(&dst[6], (unsigned long long)ntiles);
    atomicAdd(&dst[7], 1ULL);
  }
#endif
#undef KM_ROW
  tile_base += (unsigned)ntiles;
  if (wq_ == 1 && lane_ < K) {
    s.extra[lane_ * 4 + 0] = e_w;
    s.extra[lane_ * 4 + 1] = e_n;
    s.extra[lane_ * 4 + 2] = e_x;
    s.extra[lane_ * 4 + 3] = e_y;
  }
  __syncthreads();
}

// Mode-2 sweep over the `na` rows listed in s.act (the rows a bounds pass left active).
//   1. screening, warp-independent: every warp gathers its own R rows at a time into a private
//      double buffer (bulk copies on its own mbarriers), screens them and records the outcome
//      in the row's list entry -- no block barrier, so the eight warps overlap each other's
//      copy, FFMA and finalisation latency;
//   2. exact float64 pass, whole block, for the (rare) rows the screening could not prove;
//   3. the rows that changed cluster are compacted (in row order) and handed to km_sweep in
//      mode 3, which moves them between the running sums exactly as mode 2 does.
// Same decisions, same summation order as the tiled mode-2 sweep.
template <int KT, int NS2, int R>
__device__ __forceinline__ void km_sweep_sparse(const KmArgs& a, const KmSmem s,
                                                int64_t row_begin, int na,
                                                int32_t* __restrict__ assign,
                                                double (&acc)[KT][NS2][2], unsigned& tile_base,
                                                unsigned& wtile) {
  const int t = threadIdx.x, lane = t & 31, wq = t >> 5;
  const int K = a.K, Dr = a.Dr;
  const char* Xb = reinterpret_cast<const char*>(a.X);
  char* wbuf = s.buf0 + (size_t)wq * 2 * R * a.srow;          // this warp's two R-row buffers
  const int nsets = (na + R - 1) / R;                          // sets of R list entries
  constexpr int BR = R == 4 ? 32 : 16;                         // rows per decision batch
  constexpr int SPB = BR / R;
  __shared__ float stF[KM_THREADS / 32][BR * KT];
  __shared__ float stX[KM_THREADS / 32][BR * KM_NTAIL];
  const int my_n = nsets > wq ? (nsets - wq + 7) / 8 : 0;      // sets wq, wq + 8, ...
  const int ntail = Dr - a.Dm;
  const int st_r = ntail > 0 ? lane / ntail : 0, st_c = ntail > 0 ? lane - st_r * ntail : 0;
  const unsigned row_bytes = (unsigned)a.copy16 * 16u;
  unsigned long long* wbar = s.bar + 2 + wq * 2;
  auto issue = [&](int set, unsigned gt) {  // bulk copies of the set's rows on the warp's mbarrier
    const int base = set * R;
    const int nv = min(R, na - base);
    char* dst = wbuf + (size_t)(gt & 1) * R * a.srow;
    if (lane == 0) mbar_expect_tx(wbar + (gt & 1), row_bytes * (unsigned)nv);
    __syncwarp();
    if (lane < nv)
      bulk_g2s(dst + (size_t)lane * a.srow,
               Xb + ((size_t)(row_begin + (s.act[base + lane] & ACT_ROW)) * a.ldx) * sizeof(float),
               row_bytes, wbar + (gt & 1));
  };
#ifdef KM_PROFILE
  long long prof3__[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long last3__ = clock64();
#endif
  if (my_n > 0) issue(wq, wtile);
  int old_b = 0;  // lane q: assignment of the q-th row of the current batch before the sweep
  for (int i = 0; i < my_n; ++i) {
    const int set = wq + 8 * i;
    const int sb = i % SPB;  // set within the batch
    __syncwarp();  // every lane is done reading the buffer the next copy overwrites
    const unsigned gt = wtile + (unsigned)i;
    if (i + 1 < my_n) issue(set + 8, gt + 1);
    if (sb == 0) {
      const int pos = (wq + 8 * (i + lane / R)) * R + lane % R;
      if (lane < BR && pos < na) old_b = assign[row_begin + (s.act[pos] & ACT_ROW)];
    }
    KM_TICK3(0);
    mbar_wait(wbar + (gt & 1), (gt >> 1) & 1);
    KM_TICK3(1);
    const char* xw = wbuf + (size_t)(gt & 1) * R * a.srow;
    const float tot = km_screen_partial<KT, R>(a, s, xw);
    KM_TICK3(2);
    if (lane < R * KT) stF[wq][sb * R * KT + lane] = tot;
    if (lane < R * ntail)
      stX[wq][(sb * R + st_r) * KM_NTAIL + st_c] =
          reinterpret_cast<const float*>(xw + (size_t)st_r * a.srow)[a.Dm + st_c];
    KM_TICK3(3);
    if (sb == SPB - 1 || i == my_n - 1) {
      // decide the batch: lane q takes the q-th staged row
      __syncwarp();
      const int i0 = i - sb;
      const int pos = (wq + 8 * (i0 + lane / R)) * R + lane % R;   // list position of the row
      if (lane < (sb + 1) * R && pos < na) {
        const int64_t gr = row_begin + (s.act[pos] & ACT_ROW);
        float F[KT], xt[KM_NTAIL];
#pragma unroll
        for (int k = 0; k < KT; ++k) F[k] = stF[wq][lane * KT + k];
#pragma unroll
        for (int c = 0; c < KM_NTAIL; ++c) xt[c] = stX[wq][lane * KM_NTAIL + c];
        const int j = km_screen_decide<KT>(a, s, F, xt, gr, true);
        int e = s.act[pos] & ACT_ROW;
        if (j < 0) {
          e |= ACT_AMB;
        } else if (j != old_b) {
          e |= ACT_CHG | (j << ACT_NEW_SHIFT) | ((old_b & 15) << ACT_OLD_SHIFT);
          assign[gr] = j;
        }
        s.act[pos] = e;
      }
      __syncwarp();
#ifdef KM_PROFILE
      prof3__[6] += 1;
#endif
    }
    KM_TICK3(5);
  }
#ifdef KM_PROFILE
  prof3__[7] += my_n;
#endif
  KM_TICK3(0);
  wtile += (unsigned)my_n;
  if (t == 0) atomicAdd(&g_km_stats[0], (unsigned long long)na);
  // ---- exact pass ----
  int any = 0;
  __syncthreads();
  for (int p = t; p < na; p += KM_THREADS) any |= (s.act[p] & ACT_AMB) != 0;
  any = __syncthreads_or(any);
  if (any) {
    for (int p = 0; p < na; ++p) {
      const int e = s.act[p];
      if (!(e & ACT_AMB)) continue;  // uniform: every thread reads the same entry
      const int64_t gr = row_begin + (e & ACT_ROW);
      const float* xr = reinterpret_cast<const float*>(Xb + (size_t)gr * a.ldx * sizeof(float));
      double pd[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) pd[k] = 0.0;
      for (int d = t; d < Dr; d += KM_THREADS) {
        const double xv = (double)xr[d];
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          if (k < K) {
            const double df = xv - s.cen[(size_t)k * a.Dc + d];
            pd[k] = fma(df, df, pd[k]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < KT; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pd[k] += __shfl_xor_sync(0xffffffffu, pd[k], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < KT; ++k) s.red[(size_t)wq * KMAX + k] = pd[k];
      }
      __syncthreads();
      if (t == 0) {
        double dd[KT];
        double px = 0.0, py = 0.0;
        if (a.pos_mode) virtual_pos(a, gr, &px, &py);
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          double sum = 0.0;
          if (k < K) {
            for (int q = 0; q < KM_THREADS / 32; ++q) sum += s.red[(size_t)q * KMAX + k];
            if (a.pos_mode) {
              const double dx = px - s.cen[(size_t)k * a.Dc + Dr];
              const double dy = py - s.cen[(size_t)k * a.Dc + Dr + 1];
              sum = fma(dx, dx, sum);
              sum = fma(dy, dy, sum);
            }
          }
          dd[k] = sqrt(sum);
        }
        const int jj = np_argmin<KT>(dd, K);
        double sec = 1.0e300;
#pragma unroll
        for (int k = 0; k < KT; ++k)
          if (k < K && k != jj) sec = fmin(sec, dd[k]);
        const bool okd = dd[jj] == dd[jj] && sec == sec;
        a.ub[gr] = okd ? __fmul_ru(__double2float_ru(dd[jj]), 1.000001f)
                       : __int_as_float(0x7f800000);
        a.lb[gr] = okd ? __fmul_rd(__double2float_rd(sec), 0.999999f) : 0.f;
        const int old = assign[gr];
        int ne = e & ACT_ROW;
        if (jj != old) {
          ne |= ACT_CHG | (jj << ACT_NEW_SHIFT) | ((old & 15) << ACT_OLD_SHIFT);
          assign[gr] = jj;
        }
        s.act[p] = ne;
        atomicAdd(&g_km_stats[1], 1ULL);
      }
      __syncthreads();
    }
  }
  // ---- changed rows, compacted in place in list order ----
  __shared__ int wcnt2[KM_THREADS / 32];
  __shared__ int nchg_s;
  if (t == 0) nchg_s = 0;
  __syncthreads();
  for (int i0 = 0; i0 < na; i0 += KM_THREADS) {
    const int i = i0 + t;
    const int e = i < na ? s.act[i] : 0;
    const bool c = (e & ACT_CHG) != 0;
    const unsigned m = __ballot_sync(0xffffffffu, c);
    if (lane == 0) wcnt2[wq] = __popc(m);
    __syncthreads();
    int before = nchg_s;
    for (int q = 0; q < wq; ++q) before += wcnt2[q];
    if (c) s.act[before + __popc(m & ((1u << lane) - 1u))] = e;
    __syncthreads();
    if (t == 0) {
      int tot = 0;
      for (int q = 0; q < KM_THREADS / 32; ++q) tot += wcnt2[q];
      nchg_s += tot;
    }
    __syncthreads();
  }
  const int nchg = nchg_s;
  KM_TICK3(4);
#ifdef KM_PROFILE
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) atomicAdd(&g_km_prof3[i], (unsigned long long)prof3__[i]);
  }
#endif
  km_sweep<float, KT, NS2, R>(a, s, row_begin, row_begin + ACT_MAX, 3, assign, acc, tile_base,
                              nchg);
}

template <int KT, int NS2>
__device__ __forceinline__ void zero_acc(double (&acc)[KT][NS2][2]) {
#pragma unroll
  for (int k = 0; k < KT; ++k)
#pragma unroll
    for (int sl = 0; sl < NS2; ++sl) acc[k][sl][0] = acc[k][sl][1] = 0.0;
}

// centres = sums / sum(omega); returns true when some cluster has no member
template <int KT, int NS2>
__device__ __forceinline__ bool finalize_centers(const KmArgs& a, const KmSmem s,
                                                 double (&acc)[KT][NS2][2]) {
  const int t = threadIdx.x;
  const int Dr = a.Dr, K = a.K;
#pragma unroll
  for (int sl = 0; sl < NS2; ++sl) {
    const int c0 = 2 * (sl * KM_THREADS + t);
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      if (k < K) {
        const double ws = s.extra[k * 4 + 0];
        if (c0 < Dr - a.xcols) s.cen[(size_t)k * a.Dc + c0] = acc[k][sl][0] / ws;
        if (c0 + 1 < Dr - a.xcols) s.cen[(size_t)k * a.Dc + c0 + 1] = acc[k][sl][1] / ws;
      }
    }
  }
  if (a.pos_mode && t < K) {
    s.cen[(size_t)t * a.Dc + Dr] = s.extra[t * 4 + 2] / s.extra[t * 4 + 0];
    s.cen[(size_t)t * a.Dc + Dr + 1] = s.extra[t * 4 + 3] / s.extra[t * 4 + 0];
  }
  if (a.xcols && t < K) {
    for (int i = 0; i < a.xcols; ++i)
      s.cen[(size_t)t * a.Dc + Dr - a.xcols + i] = s.extra[t * 4 + 2 + i] / s.extra[t * 4 + 0];
  }
  bool empty = false;
  for (int k = 0; k < K; ++k) empty |= (s.extra[k * 4 + 1] == 0.0);
  __syncthreads();
  return empty;
}

// fp32 copies of the centres and ||c_k|| upper bounds for the screening pass.  Clusters
// K..KT-1 of the unrolled kernel get a far-away dummy centre: they never win and never make a
// row undecided, so the hot loops need no per-cluster guards.
template <int KT>
__device__ __forceinline__ void prepare_screen(const KmArgs& a, const KmSmem s) {
  // cen32 row 0 = c_0 rounded to fp32, rows 1..KT-1 = g_k = 2(c_0 - c_k), all over the fp32
  // columns [0, main_d); hk[k] = -e_k = -||c_0 - c_k||^2 over those columns (delta = dot - hk);
  // cnorm[0] >= ||c_0||, cnorm[k] >= ||g_k||; hk[0] = largest ||c_k||^2 (magnitude scale of the
  // float64 floor).  Clusters K..KT-1 of the unrolled kernel get g = 0 and a huge delta: they
  // never win and never make a row undecided.
  const int t = threadIdx.x;
  const int main_d = a.Dm;
  for (int i = t; i < KT * a.Dc; i += KM_THREADS) {
    const int k = i / a.Dc, d = i - k * a.Dc;
    float g = 0.f;
    if (d < main_d) {
      if (k == 0) g = (float)s.cen[d];
      else if (k < a.K) g = (float)(2.0 * (s.cen[d] - s.cen[i]));
    }
    s.cen32[i] = g;
  }
  const int wq = t >> 5, lane = t & 31;
  if (wq < KT) {
    double e = 0.0, g2 = 0.0, c2 = 0.0, c2m = 0.0;
    if (wq < a.K) {
      for (int d = lane; d < a.D; d += 32) {
        const double c0 = s.cen[d], ck = s.cen[(size_t)wq * a.Dc + d];
        c2 = fma(ck, ck, c2);
        if (d < main_d) {
          c2m = fma(ck, ck, c2m);
          e = fma(c0 - ck, c0 - ck, e);
        }
      }
      g2 = 4.0 * e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      e += __shfl_xor_sync(0xffffffffu, e, o);
      g2 += __shfl_xor_sync(0xffffffffu, g2, o);
      c2 += __shfl_xor_sync(0xffffffffu, c2, o);
      c2m += __shfl_xor_sync(0xffffffffu, c2m, o);
    }
    if (lane == 0) {
      if (wq == 0) s.cnorm[0] = (float)sqrt(c2m) * 1.0001f;
      else s.cnorm[wq] = wq < a.K ? (float)sqrt(g2) * 1.0001f : 0.f;
      s.red[wq] = c2;                                   // ||c_k||^2, gathered below
      if (wq >= 1) s.hk[wq] = wq < a.K ? -e : -1.0e300;  // delta = dot - hk: dummies far away
    }
  }
  // fp32 constants of the fast decision path, over the nt columns outside the screening loop
  // (tail column c is column Dm + c, stored or virtual):
  //   gt[k][c] = 2(c0_c - ck_c), c0t[c] = c0_c, ek[k] = ||c0 - ck||^2 over ALL columns,
  //   ak[k] >= sum_c |gt[k][c]| |c0_c|, c0tn >= ||c0 tail||
  if (t < KT) {
    const int nt = a.D - main_d;
    double ek = 0.0, ak = 0.0;
    for (int c = 0; c < KM_NT; ++c) {
      double g = 0.0;
      if (c < nt && t < a.K) {
        const double c0 = s.cen[main_d + c], ck = s.cen[(size_t)t * a.Dc + main_d + c];
        g = 2.0 * (c0 - ck);
        ek = fma(c0 - ck, c0 - ck, ek);
        ak += fabs(g) * fabs(c0);
      }
      s.fz[FZ_GT + t * KM_NT + c] = (float)g;
    }
    s.fz[FZ_AK + t] = __double2float_ru(ak) * 1.001f;
    s.red[KMAX + t] = ek;  // tail share of e_k; the main share is -hk[k], added below
  }
  if (t == KT) {
    const int nt = a.D - main_d;
    double n2 = 0.0;
    for (int c = 0; c < KM_NT; ++c) {
      const double c0 = c < nt ? s.cen[main_d + c] : 0.0;
      s.fz[FZ_C0T + c] = (float)c0;
      n2 = fma(c0, c0, n2);
    }
    s.fz[FZ_C0TN] = __double2float_ru(sqrt(n2)) * 1.001f;
  }
  __syncthreads();
  if (t == 0) {
    double m = 0.0;
    for (int k = 0; k < a.K; ++k) m = fmax(m, s.red[k]);   // NaN centres: fmax skips them,
    s.hk[0] = m;                                            // the row test catches NaN deltas
  }
  if (t >= 1 && t < KT) s.fz[FZ_EK + t] = (float)(s.red[KMAX + t] - s.hk[t]);
  __syncthreads();
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

template <typename XT, int KT, int NS2, int R, int MINB>
__global__ void __launch_bounds__(KM_THREADS, MINB) kmeans_groups_kernel(GroupArgs g) {
  extern __shared__ __align__(128) char smem_raw[];
  KmSmem s;
  km_carve(&s, smem_raw, g.a.TR, g.a.srow, g.a.K, g.a.Dc, g.a.Kc, g.a.part_bytes, g.a.buf_rows);
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
  double acc[KT][NS2][2];
  zero_acc<KT, NS2>(acc);
  if (t == 0) *s.changed = 0;
  unsigned tile_base = 0;
  km_init_barriers(s);
  __syncthreads();
  km_sweep<XT, KT, NS2, R>(g.a, s, r0, r1, 0, g.assign, acc, tile_base);
  finalize_centers<KT, NS2>(g.a, s, acc);
  int it = 0, status = SPALIGN_KM_ITER_CAP;
  while (it < g.n_iter) {
    ++it;
    zero_acc<KT, NS2>(acc);
    if (t == 0) *s.changed = 0;
    if (sizeof(XT) == 4) prepare_screen<KT>(g.a, s);
    __syncthreads();
    km_sweep<XT, KT, NS2, R>(g.a, s, r0, r1, 1, g.assign, acc, tile_base);
    const int changed = *s.changed;  // km_sweep ends with __syncthreads
    if (changed == 0) {
      status = SPALIGN_KM_CONVERGED;
      break;
    }
    if (finalize_centers<KT, NS2>(g.a, s, acc)) {
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

// Reduction tree of the chunk partials of one group (the same in kmeans_reduce_kernel and in the
// fused finish, so both give the same bits): runs of KM_SUPER consecutive slots are summed in
// slot order, then the run sums are summed in run order.  In the fused finish the runs are
// summed by different CTAs (the chunk of a run that finishes last), so the serial part of a
// group of n chunks is n/KM_SUPER + KM_SUPER vector reads instead of n.
constexpr int KM_SUPER = 32;

// sum of base[i * stride], i = 0..n-1 ascending, U loads in flight (L2 loads: the values were
// written by other CTAs of this launch)
template <int U>
__device__ __forceinline__ double ordered_sum_cg(const double* base, size_t stride, int n) {
  double sum = 0.0;
  int c = 0;
  for (; c + U <= n; c += U) {
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldcg(base + (size_t)(c + u) * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) sum += v[u];
  }
  for (; c < n; ++c) sum += __ldcg(base + (size_t)c * stride);
  return sum;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

struct SweepArgs {
  KmArgs a;
  const int64_t* chunks;  // [n_chunks][4]: group, row_begin, row_end, slot
  const double* centers;  // [G][K][D]
  int mode;
  int32_t* assign;
  const int32_t* status;
  double* partials;       // [n_chunks][K*(D+2)+1]
  // fused finish (optional): the last chunk of a group to finish reduces and updates
  const int32_t* gco;     // [G+1] chunk offsets per group
  int32_t* counters;      // [n_chunks + G] arrival tickets (first-level reducers by slot, then
                          // groups), zero between launches; NULL = no fused finish
  double* totals;         // [G][K*(D+2)+1]
  double* centers_rw;     // [G][K][D]
  int32_t* iters;
  int32_t* status_rw;
  int n_iter;
  double* cdelta;         // [G][K] centre drift of the last update (Hamerly bounds); may be NULL
  PeerComm peer;          // multi-GPU exchange of the reduced vector (world <= 1: none)
};

// centres / stop flags of one group from its reduced totals (shared by kmeans_update_kernel
// and the fused finish of kmeans_sweep_kernel); whole block, uniform control flow
__device__ __forceinline__ void km_update_group(const double* tt, int D, int K, int mode,
                                                int n_iter, double* c, int32_t* iters,
                                                int32_t* status, int grp,
                                                double* cdelta = nullptr) {
  const int pv = K * (D + 2) + 1;
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
  // new centres; cdelta[k] = ||c_k(new) - c_k(old)|| for the Hamerly bounds (fixed-order sums)
  __shared__ double s_d2[KMAX][8];
  for (int k = 0; k < K; ++k) {
    double d2 = 0.0;
    for (int d = t; d < D; d += blockDim.x) {
      const double nv = tt[(size_t)k * (D + 2) + d] / tt[(size_t)k * (D + 2) + D];
      const double df = nv - c[(size_t)k * D + d];
      d2 = fma(df, df, d2);
      c[(size_t)k * D + d] = nv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    if ((t & 31) == 0) s_d2[k][t >> 5] = d2;
  }
  __syncthreads();
  if (cdelta != nullptr && t < K) {
    double d2 = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) d2 += s_d2[t][q];
    cdelta[(size_t)grp * K + t] = sqrt(d2) * (1.0 + 1e-12);
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

template <typename XT, int KT, int NS2, int R, int MINB>
__global__ void __launch_bounds__(KM_THREADS, MINB) kmeans_sweep_kernel(SweepArgs g) {
  extern __shared__ __align__(128) char smem_raw[];
  KmSmem s;
  km_carve(&s, smem_raw, g.a.TR, g.a.srow, g.a.K, g.a.Dc, g.a.Kc, g.a.part_bytes, g.a.buf_rows);
  // chunks are listed in launch order (large ones first); column 3 is the chunk's slot in the
  // group-contiguous, row-ordered partials array that the reduction walks
  const int grp = (int)g.chunks[(size_t)blockIdx.x * 4];
  const int64_t rb = g.chunks[(size_t)blockIdx.x * 4 + 1], re = g.chunks[(size_t)blockIdx.x * 4 + 2];
  const int ck = (int)g.chunks[(size_t)blockIdx.x * 4 + 3];
  if (g.status[grp] != SPALIGN_KM_RUNNING) return;
  const int t = threadIdx.x;
  const int K = g.a.K, D = g.a.D, Dr = g.a.Dr;
  double acc[KT][NS2][2];
  zero_acc<KT, NS2>(acc);
  if (t == 0) *s.changed = 0;
  unsigned tile_base = 0;
  km_init_barriers(s);
  __syncthreads();
  // mode 2 with Hamerly bounds: find the rows that must be re-examined before anything else;
  // a chunk without such rows only reports zeros
  int nrows_in = -1;
  if (sizeof(XT) == 4 && g.mode == 2 && g.a.ub != nullptr && g.cdelta != nullptr &&
      re - rb <= ACT_MAX)
    nrows_in = km_bounds_pass<KT>(g.a, s, rb, (int)(re - rb), g.assign,
                                  g.cdelta + (size_t)grp * K);
  if (nrows_in != 0) {
    if (g.mode != 0) {
      const double* c = g.centers + (size_t)grp * K * D;
      for (int i = t; i < K * D; i += KM_THREADS) {
        const int k = i / D, d = i - k * D;
        s.cen[(size_t)k * g.a.Dc + d] = c[i];
      }
      __syncthreads();
      if (sizeof(XT) == 4) prepare_screen<KT>(g.a, s);
    }
    if constexpr (sizeof(XT) == 4) {
      if (nrows_in > 0) {
        unsigned wtile = 0;
        km_sweep_sparse<KT, NS2, R>(g.a, s, rb, nrows_in, g.assign, acc, tile_base, wtile);
      } else {
        km_sweep<XT, KT, NS2, R>(g.a, s, rb, re, g.mode, g.assign, acc, tile_base);
      }
    } else {
      km_sweep<XT, KT, NS2, R>(g.a, s, rb, re, g.mode, g.assign, acc, tile_base);
    }
  } else if (t < K) {
    s.extra[t * 4 + 0] = s.extra[t * 4 + 1] = s.extra[t * 4 + 2] = s.extra[t * 4 + 3] = 0.0;
  }
  __syncthreads();
  const size_t pv = (size_t)K * (D + 2) + 1;
  double* out = g.partials + (size_t)ck * pv;
#pragma unroll
  for (int sl = 0; sl < NS2; ++sl) {
    const int c0 = 2 * (sl * KM_THREADS + t);
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      if (k < K) {
        if (c0 < Dr - g.a.xcols) out[(size_t)k * (D + 2) + c0] = acc[k][sl][0];
        if (c0 + 1 < Dr - g.a.xcols) out[(size_t)k * (D + 2) + c0 + 1] = acc[k][sl][1];
      }
    }
  }
  if (t < K) {
    if (g.a.pos_mode) {
      out[(size_t)t * (D + 2) + Dr] = s.extra[t * 4 + 2];
      out[(size_t)t * (D + 2) + Dr + 1] = s.extra[t * 4 + 3];
    }
    for (int i = 0; i < g.a.xcols; ++i)
      out[(size_t)t * (D + 2) + Dr - g.a.xcols + i] = s.extra[t * 4 + 2 + i];
    out[(size_t)t * (D + 2) + D] = s.extra[t * 4 + 0];
    out[(size_t)t * (D + 2) + D + 1] = s.extra[t * 4 + 1];
  }
  if (t == 0) out[pv - 1] = (double)*s.changed;
  if (g.counters != nullptr) {
    // fused finish: the chunks that arrive last sum the group's partials along the fixed tree
    // (same as kmeans_reduce_kernel -> same bits) and apply the update
    __shared__ int s_last;
    const int c0 = g.gco[grp], c1 = g.gco[grp + 1];
    const int nch = c1 - c0;
    const int nsup = (nch + KM_SUPER - 1) / KM_SUPER;
    int32_t* gcount = g.counters + gridDim.x + grp;
    __threadfence();
    __syncthreads();
    if (nsup > 1) {
      // first level: the last chunk of this run of KM_SUPER slots sums the run into its first slot
      const int sidx = (ck - c0) / KM_SUPER;
      const int sb = c0 + sidx * KM_SUPER, se = min(sb + KM_SUPER, c1);
      if (t == 0) s_last = atomicAdd(&g.counters[sb], 1) == se - sb - 1;
      __syncthreads();
      if (!s_last) return;
      __threadfence();
      double* run = g.partials + (size_t)sb * pv;
      for (int j = t; j < (int)pv; j += 2 * KM_THREADS) {
        const int j2 = j + KM_THREADS;
        const double v0 = ordered_sum_cg<16>(run + j, pv, se - sb);
        const double v1 = j2 < (int)pv ? ordered_sum_cg<16>(run + j2, pv, se - sb) : 0.0;
        __stcg(run + j, v0);
        if (j2 < (int)pv) __stcg(run + j2, v1);
      }
      if (t == 0) g.counters[sb] = 0;
      __threadfence();
      __syncthreads();
    }
    if (t == 0) s_last = atomicAdd(gcount, 1) == (nsup > 1 ? nsup : nch) - 1;
    __syncthreads();
    if (s_last) {
      __threadfence();
      double* tt = g.totals + (size_t)grp * pv;
      const double* first = g.partials + (size_t)c0 * pv;
      const size_t stride = nsup > 1 ? (size_t)KM_SUPER * pv : pv;
      const int terms = nsup > 1 ? nsup : nch;
      const PeerComm& pc = g.peer;
      unsigned long long seq = 0;
      int par = 0;
      if (pc.world > 1) {
        seq = *pc.xcount + 1;
        par = (int)(seq & 1ull);
      }
      for (int j = t; j < (int)pv; j += KM_THREADS) {
        const double sum = ordered_sum_cg<16>(first + j, stride, terms);
        if (pc.world > 1) {
          // publish this rank's vector into every rank's inbox (own one included)
          const size_t slot = ((size_t)par * pc.world + pc.rank) * pc.pv_cap + j;
          for (int r = 0; r < pc.world; ++r) pc.inbox[r][slot] = sum;
        } else {
          // mode 2: the partials are deltas of the running sums (rows that changed cluster)
          tt[j] = (g.mode == 2 && j != (int)pv - 1) ? tt[j] + sum : sum;
        }
      }
      if (pc.world > 1) {
        __shared__ int s_timeout;
        if (t == 0) s_timeout = 0;
        __threadfence_system();
        __syncthreads();
        if (t < pc.world) {
          st_release_sys(pc.flags[t] + (size_t)par * pc.world + pc.rank, seq);
          const unsigned long long* mine = pc.flags[pc.rank] + (size_t)par * pc.world + t;
          const long long t0 = clock64();
          while (ld_acquire_sys(mine) < seq) {
            if (clock64() - t0 > 6000000000LL) {  // ~3 s: a peer is gone; give up cleanly
              s_timeout = 1;
              break;
            }
          }
        }
        __syncthreads();
        if (s_timeout) {
          if (t == 0) {
            g.status_rw[grp] = SPALIGN_KM_COMM_TIMEOUT;
            *gcount = 0;
          }
          return;
        }
        // every rank adds the world's vectors in rank order -> bit-identical totals everywhere
        const double* in = pc.inbox[pc.rank] + (size_t)par * pc.world * pc.pv_cap;
        for (int j = t; j < (int)pv; j += KM_THREADS) {
          double sum = 0.0;
          for (int r = 0; r < pc.world; ++r) sum += ld_volatile_f64(in + (size_t)r * pc.pv_cap + j);
          tt[j] = (g.mode == 2 && j != (int)pv - 1) ? tt[j] + sum : sum;
        }
        if (t == 0) *pc.xcount = seq;
      }
      __syncthreads();
      km_update_group(tt, D, K, g.mode == 2 ? 1 : g.mode, g.n_iter,
                      g.centers_rw + (size_t)grp * K * D, g.iters, g.status_rw, grp, g.cdelta);
      if (t == 0) *gcount = 0;
    }
  }
}

struct TailArgs {
  KmArgs a;
  const int64_t* group_off;  // [G+1]
  int32_t* assign;
  double* totals;            // [G][K*(D+2)+1] running sums left by the full iteration
  double* centers;           // [G][K][D]
  int32_t* iters;
  int32_t* status;
  double* cdelta;            // [G][K]
  int n_iter;
  int slice;                 // iterations to run in this launch (0: until the group stops)
};

// Remaining iterations of one group by one persistent CTA (fp32 rows, after the first full
// iteration of kmeans_sweep_kernel has left running sums, centres, centre drift and Hamerly
// bounds).  Every iteration is a mode-2 sweep: bounds pass, gather + screen only the rows the
// bounds cannot prove stable, move the rows that changed cluster between the running sums,
// new centres and drift -- all without leaving the SM, so an iteration costs a few
// microseconds instead of a launch, and groups proceed independently of each other.
template <typename XT, int KT, int NS2, int R, int MINB>
__global__ void __launch_bounds__(KM_THREADS, MINB) kmeans_tail_kernel(TailArgs g) {
  extern __shared__ __align__(128) char smem_raw[];
  if (sizeof(XT) != 4) return;  // host side refuses float64 rows
  KmSmem s;
  km_carve(&s, smem_raw, g.a.TR, g.a.srow, g.a.K, g.a.Dc, g.a.Kc, g.a.part_bytes, g.a.buf_rows);
  const int grp = blockIdx.x;
  if (g.status[grp] != SPALIGN_KM_RUNNING) return;
  const int64_t r0 = g.group_off[grp], r1 = g.group_off[grp + 1];
  const int t = threadIdx.x;
  const int K = g.a.K, D = g.a.D, Dr = g.a.Dr, Dc = g.a.Dc;
  const size_t pv = (size_t)K * (D + 2) + 1;
  double* tt = g.totals + (size_t)grp * pv;
  double* cg = g.centers + (size_t)grp * K * D;
  double* cd = g.cdelta + (size_t)grp * K;
  __shared__ double xs[KMAX * 4];
  km_init_barriers(s);
  for (int i = t; i < K * D; i += KM_THREADS) {
    const int k = i / D, d = i - k * D;
    s.cen[(size_t)k * Dc + d] = cg[i];
  }
  __syncthreads();
  int it = g.iters[grp];
  int status = SPALIGN_KM_ITER_CAP;
#ifdef KM_PROFILE
  if (t == 0 && grp < 1024) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    g_km_trace[grp * 3] = ns;
  }
#endif
  unsigned tile_base = 0, wtile = 0;
  bool first = true;
  int done_here = 0;
  double acc[KT][NS2][2];
#ifdef KM_PROFILE
  long long prof2__[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long last2__ = clock64();
#endif
  while (it < g.n_iter) {
    zero_acc<KT, NS2>(acc);
    if (t == 0) *s.changed = 0;
    if (t < KMAX * 4) xs[t] = 0.0;
    if (first) prepare_screen<KT>(g.a, s);  // later iterations: done by the centre update
    first = false;
    __syncthreads();
    KM_TICK2(0);
    for (int64_t rb = r0; rb < r1; rb += ACT_MAX) {
      const int n = (int)min((int64_t)ACT_MAX, r1 - rb);
      const int na = km_bounds_pass<KT>(g.a, s, rb, n, g.assign, cd);
      KM_TICK2(1);
      if (na > 0) {
        if constexpr (sizeof(XT) == 4)
          km_sweep_sparse<KT, NS2, R>(g.a, s, rb, na, g.assign, acc, tile_base, wtile);
        if (t < K * 4) xs[t] += s.extra[t];
        __syncthreads();
      }
      KM_TICK2(2);
    }
    __syncthreads();
    ++it;
#ifdef KM_PROFILE
    prof2__[4] += 1;
#endif
    if (*s.changed == 0) {  // assignment unchanged: centres stay
      status = SPALIGN_KM_CONVERGED;
      break;
    }
    // running sums += this iteration's moves (thread-owned columns; issued first so that this
    // global round trip overlaps the one of the per-cluster scalars below)
#pragma unroll
    for (int sl = 0; sl < NS2; ++sl) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = 2 * (sl * KM_THREADS + t) + j;
        if (c < Dr - g.a.xcols) {
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            if (k < K) {
              const size_t idx = (size_t)k * (D + 2) + c;
              acc[k][sl][j] += tt[idx];   // acc now holds the new running sum
              tt[idx] = acc[k][sl][j];
            }
          }
        }
      }
    }
    // per-cluster scalars, new centres and their drift
    if (t < K) {
      double* e = tt + (size_t)t * (D + 2);
      const double ws = e[D] + xs[t * 4 + 0];
      const double cnt = e[D + 1] + xs[t * 4 + 1];
      e[D] = ws;
      e[D + 1] = cnt;
      double dp = 0.0;
      if (g.a.pos_mode) {
        const double sx = e[Dr] + xs[t * 4 + 2], sy = e[Dr + 1] + xs[t * 4 + 3];
        e[Dr] = sx;
        e[Dr + 1] = sy;
        const double nx = sx / ws, ny = sy / ws;
        const double dx = nx - s.cen[(size_t)t * Dc + Dr], dy = ny - s.cen[(size_t)t * Dc + Dr + 1];
        dp = fma(dx, dx, dy * dy);
        s.cen[(size_t)t * Dc + Dr] = nx;
        s.cen[(size_t)t * Dc + Dr + 1] = ny;
      }
      for (int i = 0; i < g.a.xcols; ++i) {  // the last stored columns travel with the scalars
        const int c = Dr - g.a.xcols + i;
        const double v = e[c] + xs[t * 4 + 2 + i];
        e[c] = v;
        const double nv = v / ws;
        const double df = nv - s.cen[(size_t)t * Dc + c];
        dp = fma(df, df, dp);
        s.cen[(size_t)t * Dc + c] = nv;
      }
      s.extra[t * 4 + 0] = ws;
      s.extra[t * 4 + 1] = cnt;
      s.extra[t * 4 + 2] = dp;
    }
    __syncthreads();
    // every thread owns its columns of all K centres: new centres, their drift, and -- in the
    // same pass -- everything the next iteration's screening needs (what prepare_screen
    // computes from scratch): cen32 = c_0 | g_k, e_k = ||c_0 - c_k||^2, ||c_k||^2
    double q[KT][3];   // per cluster: drift^2, e_k over the fp32 columns, ||c_k||^2 (stored)
    double c2m = 0.0;  // ||c_0||^2 over the fp32 columns
#pragma unroll
    for (int k = 0; k < KT; ++k) q[k][0] = q[k][1] = q[k][2] = 0.0;
#pragma unroll
    for (int sl = 0; sl < NS2; ++sl) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = 2 * (sl * KM_THREADS + t) + j;
        if (c < Dr - g.a.xcols) {
          const bool mainc = c < g.a.Dm;
          double cn0 = 0.0;
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            if (k < K) {
              const double nv = acc[k][sl][j] / s.extra[k * 4 + 0];
              const double df = nv - s.cen[(size_t)k * Dc + c];
              q[k][0] = fma(df, df, q[k][0]);
              s.cen[(size_t)k * Dc + c] = nv;
              q[k][2] = fma(nv, nv, q[k][2]);
              if (k == 0) {
                cn0 = nv;
                if (mainc) c2m = fma(nv, nv, c2m);
                s.cen32[c] = mainc ? (float)nv : 0.f;
              } else {
                const double dd = cn0 - nv;
                if (mainc) q[k][1] = fma(dd, dd, q[k][1]);
                s.cen32[(size_t)k * Dc + c] = mainc ? (float)(2.0 * dd) : 0.f;
              }
            }
          }
        }
      }
    }
    // block sums in a fixed order: lanes (xor tree), then warps
    __shared__ double redq[KM_THREADS / 32][3 * KMAX + 1];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
#pragma unroll
      for (int u = 0; u < 3; ++u) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q[k][u] += __shfl_xor_sync(0xffffffffu, q[k][u], o);
        if ((t & 31) == 0) redq[t >> 5][k * 3 + u] = q[k][u];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c2m += __shfl_xor_sync(0xffffffffu, c2m, o);
    if ((t & 31) == 0) redq[t >> 5][3 * KMAX] = c2m;
    __syncthreads();
    if (t < KT) {
      const bool in = t < K;
      double d2s = 0.0, es = 0.0, c2s = 0.0, c2ms = 0.0;
      for (int w8 = 0; w8 < KM_THREADS / 32; ++w8) {
        d2s += redq[w8][t * 3 + 0];
        es += redq[w8][t * 3 + 1];
        c2s += redq[w8][t * 3 + 2];
        c2ms += redq[w8][3 * KMAX];
      }
      if (in) {
        cd[t] = sqrt(d2s + s.extra[t * 4 + 2]) * (1.0 + 1e-12);
        if (g.a.pos_mode) {
          const double px = s.cen[(size_t)t * Dc + Dr], py = s.cen[(size_t)t * Dc + Dr + 1];
          c2s = fma(px, px, fma(py, py, c2s));
        }
        for (int i = 0; i < g.a.xcols; ++i) {
          const double cv = s.cen[(size_t)t * Dc + Dr - g.a.xcols + i];
          c2s = fma(cv, cv, c2s);
        }
      }
      if (t == 0) s.cnorm[0] = (float)sqrt(c2ms) * 1.0001f;
      else s.cnorm[t] = in ? (float)sqrt(4.0 * es) * 1.0001f : 0.f;
      s.red[t] = in ? c2s : 0.0;
      if (t >= 1) s.hk[t] = in ? -es : -1.0e300;
      // fp32 constants of the fast decision path (see prepare_screen)
      const int nt = D - g.a.Dm;
      double ek = 0.0, ak = 0.0;
      for (int c = 0; c < KM_NT; ++c) {
        double gg = 0.0;
        if (c < nt && in) {
          const double c0 = s.cen[g.a.Dm + c], ck = s.cen[(size_t)t * Dc + g.a.Dm + c];
          gg = 2.0 * (c0 - ck);
          ek = fma(c0 - ck, c0 - ck, ek);
          ak += fabs(gg) * fabs(c0);
        }
        s.fz[FZ_GT + t * KM_NT + c] = (float)gg;
      }
      s.fz[FZ_AK + t] = __double2float_ru(ak) * 1.001f;
      if (t >= 1) s.fz[FZ_EK + t] = (float)(ek + (in ? es : 1.0e300));
    }
    if (t == KT) {
      const int nt = D - g.a.Dm;
      double n2 = 0.0;
      for (int c = 0; c < KM_NT; ++c) {
        const double c0 = c < nt ? s.cen[g.a.Dm + c] : 0.0;
        s.fz[FZ_C0T + c] = (float)c0;
        n2 = fma(c0, c0, n2);
      }
      s.fz[FZ_C0TN] = __double2float_ru(sqrt(n2)) * 1.001f;
    }
    bool empty = false;
    for (int k = 0; k < K; ++k) empty |= (s.extra[k * 4 + 1] == 0.0);
    if (empty) {
      status = SPALIGN_KM_EMPTY_CLUSTER;
      break;
    }
    __syncthreads();
    if (t == 0) {  // magnitude scale of the float64 floor (read after the next block barriers)
      double m = 0.0;
      for (int k = 0; k < K; ++k) m = fmax(m, s.red[k]);
      s.hk[0] = m;
    }
    KM_TICK2(3);
    if (g.slice > 0 && ++done_here >= g.slice && it < g.n_iter) {
      status = SPALIGN_KM_RUNNING;  // paused: a later launch continues from the global state
      break;
    }
  }
  __syncthreads();
#ifdef KM_PROFILE
  if (threadIdx.x == 0)
    for (int i = 0; i < 5; ++i) atomicAdd(&g_km_prof2[i], (unsigned long long)prof2__[i]);
#endif
  for (int i = t; i < K * D; i += KM_THREADS) {
    const int k = i / D, d = i - k * D;
    cg[i] = s.cen[(size_t)k * Dc + d];
  }
  if (t == 0) {
    g.iters[grp] = it;
    g.status[grp] = status;
#ifdef KM_PROFILE
    if (grp < 1024) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
      g_km_trace[grp * 3 + 1] = ns;
      g_km_trace[grp * 3 + 2] = (unsigned long long)it;
    }
#endif
  }
}

__global__ void __launch_bounds__(256)
kmeans_reduce_kernel(const double* __restrict__ partials, const int32_t* __restrict__ gco,
                     int pv, double* __restrict__ totals) {
  const int grp = blockIdx.y;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= pv) return;
  // the tree of the fused finish: runs of KM_SUPER slots in order, then the run sums in order
  const int c0 = gco[grp], c1 = gco[grp + 1];
  const int nsup = (c1 - c0 + KM_SUPER - 1) / KM_SUPER;
  double sum = 0.0;
  if (nsup <= 1) {
    sum = ordered_sum_cg<16>(partials + (size_t)c0 * pv + j, pv, c1 - c0);
  } else {
    for (int s = 0; s < nsup; ++s) {
      const int sb = c0 + s * KM_SUPER, se = min(sb + KM_SUPER, c1);
      sum += ordered_sum_cg<16>(partials + (size_t)sb * pv + j, pv, se - sb);
    }
  }
  totals[(size_t)grp * pv + j] = sum;
}

__global__ void __launch_bounds__(256)
kmeans_update_kernel(const double* __restrict__ totals, int D, int K, int mode, int n_iter,
                     double* centers, int32_t* iters, int32_t* status) {
  const int grp = blockIdx.x;
  if (status[grp] != SPALIGN_KM_RUNNING) return;
  const int pv = K * (D + 2) + 1;
  km_update_group(totals + (size_t)grp * pv, D, K, mode, n_iter, centers + (size_t)grp * K * D,
                  iters, status, grp);
}

// seeded init: upper median of the group's weights -- bitonic sort in shared memory up to
// INIT_MAX rows, an 8-pass radix select over the rows in global memory beyond that (joint
// clustering of 30 images = 30 000 rows, direct cell clustering) -- then the rows at or below it
// take the host-shuffled cluster ids in row order
constexpr int INIT_MAX = 4096;
__device__ __forceinline__ unsigned long long f64_order_key(double d) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);  // ascending keys == ascending doubles
}
__global__ void __launch_bounds__(256)
kmeans_init_kernel(const double* __restrict__ w, const int64_t* __restrict__ group_off,
                   const int32_t* __restrict__ shuffled, const int64_t* __restrict__ shuf_off,
                   int32_t* assign, int32_t* m_out) {
  __shared__ double key[INIT_MAX];
  const int grp = blockIdx.x;
  const int64_t r0 = group_off[grp];
  const int64_t n64 = group_off[grp + 1] - r0;
  const int t = threadIdx.x;
  if (n64 <= 0) {
    if (t == 0) m_out[grp] = 0;
    return;
  }
  if (n64 > 0x7fffffffLL) {  // not a size this path is meant for
    if (t == 0) m_out[grp] = -1;
    return;
  }
  const int n = (int)n64;
  int p2 = 1;
  while (p2 < n) p2 <<= 1;
  double thr;
  if (p2 <= INIT_MAX) {
    for (int i = t; i < p2; i += 256)
      key[i] = i < n ? w[r0 + i] : __longlong_as_double(0x7ff0000000000000LL);
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
    thr = key[n / 2];
    __syncthreads();
  } else {
    // radix select of the element of rank n/2 (ascending), most significant byte first
    int* hist = reinterpret_cast<int*>(key);
    __shared__ unsigned long long s_prefix;
    __shared__ int s_rank;
    if (t == 0) {
      s_prefix = 0ull;
      s_rank = n / 2;
    }
    for (int pass = 7; pass >= 0; --pass) {
      hist[t] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      for (int i = t; i < n; i += 256) {
        const unsigned long long k = f64_order_key(w[r0 + i]);
        if (pass == 7 || (k >> (8 * (pass + 1))) == prefix)
          atomicAdd(&hist[(int)((k >> (8 * pass)) & 255ull)], 1);
      }
      __syncthreads();
      if (t == 0) {
        int rank = s_rank, b = 0;
        while (b < 255 && rank >= hist[b]) rank -= hist[b++];
        s_rank = rank;
        s_prefix = (prefix << 8) | (unsigned long long)b;
      }
      __syncthreads();
    }
    const unsigned long long k = s_prefix;
    thr = __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
    __syncthreads();
  }
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
  int variant;  // 0: <float,K<=4,Dr<=1024>   1: <float,K<=8,Dr<=2048>   2: <double,K<=8,Dr<=2048>
  int R;        // rows per warp in the fp32 screening pass (TR = 8*R); double path: TR free
  int TR, logTR;
  int srow;
  int copy16;
  int Dc;
  int Kc;
  int part_bytes;
  int buf_rows;
  size_t smem;
};

bool make_plan(int x_dtype, int D, int Dr, int K, Plan* p, int max_tr = 32, int sparse_r = 0) {
  const int es = x_dtype == SPALIGN_F32 ? 4 : 8;
  const int row_bytes = (int)align_up((size_t)Dr * es, 16);
  int srow = row_bytes;
  // rows 16 bytes apart modulo 128 -> conflict-free 16-byte reads with one row per lane
  while (srow % 128 != 16) srow += 16;
  p->srow = srow;
  p->copy16 = row_bytes / 16;
  p->Dc = (int)align_up((size_t)D, 4);
  const bool small = x_dtype == SPALIGN_F32 && K <= 4 && Dr <= 1024;
  p->variant = x_dtype == SPALIGN_F64 ? 2 : (small ? 0 : 1);
  const size_t limit = (size_t)200 * 1024;
  p->Kc = small ? 4 : 8;
  p->buf_rows = 0;
  if (sparse_r == 4 && small) {
    // finish kernel, latency over occupancy: the warps of the sparse sweep take 4 rows per set
    // (one CTA per SM); the tiles that move changed rows keep 16 rows
    const int pb = (KM_THREADS / 32) * (4 * p->Kc) * 33 * (int)sizeof(float);
    const size_t bytes = km_carve(nullptr, nullptr, 16, srow, K, p->Dc, p->Kc, pb, 32);
    if (bytes <= limit) {
      p->R = 4; p->TR = 16; p->logTR = 4; p->part_bytes = pb; p->buf_rows = 32; p->smem = bytes;
      return true;
    }
  }
  if (p->variant == 2) {
    for (int tr = 32, lg = 5; tr >= 2; tr >>= 1, --lg) {
      const int pb = KM_THREADS * p->Kc * (int)sizeof(double);
      size_t bytes = km_carve(nullptr, nullptr, tr, srow, K, p->Dc, p->Kc, pb);
      if (bytes <= limit) {
        p->TR = tr; p->logTR = lg; p->R = 1; p->part_bytes = pb; p->smem = bytes;
        return true;
      }
    }
    return false;
  }
  // fp32: TR = 8*R.  Variant 0 prefers R=2 (two CTAs per SM, so the fp32 and fp64 phases of the
  // two CTAs overlap) unless SPALIGN_KM_R=4 asks for the bigger tile; variant 1: R=2 then 1.
  int want = 2;
  if (const char* e = getenv("SPALIGN_KM_R")) {
    int v = atoi(e);
    if (small && (v == 2 || v == 4)) want = v;
  }
  const int cands[3] = {want, 2, 1};
  for (int i = 0; i < 3; ++i) {
    int r = cands[i];
    if (!small && r == 4) continue;
    if (small && r == 1) continue;
    if (8 * r > max_tr) continue;
    const int pb = (KM_THREADS / 32) * (r * p->Kc) * 33 * (int)sizeof(float);
    size_t bytes = km_carve(nullptr, nullptr, 8 * r, srow, K, p->Dc, p->Kc, pb);
    if (bytes <= limit) {
      p->R = r; p->TR = 8 * r; p->logTR = r == 4 ? 5 : (r == 2 ? 4 : 3);
      p->part_bytes = pb; p->smem = bytes;
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
  SPALIGN_REQUIRE(Dr >= 1 && Dr <= 2048, "kmeans: stored columns out of range (max 2048)");
  const int es = x_dtype == SPALIGN_F32 ? 4 : 8;
  SPALIGN_REQUIRE((ldx * es) % 16 == 0 && ldx * es >= (int64_t)align_up((size_t)Dr * es, 16),
                  "kmeans: row stride must be a multiple of 16 bytes covering the padded row");
  SPALIGN_REQUIRE(reinterpret_cast<size_t>(X) % 16 == 0, "kmeans: X must be 16-byte aligned");
  SPALIGN_REQUIRE(!pos_mode || (pos_w > 0 && pos_period > 0), "kmeans: bad pos_w/pos_period");
  a->X = X; a->ldx = ldx; a->pos_mode = pos_mode; a->pos_w = pos_w ? pos_w : 1;
  a->pos_period = pos_period ? pos_period : 1; a->pos_row0 = pos_row0; a->w = w;
  // fp32-screened columns: all but the last <= 3 stored ones; a row length of 4n + 2 is taken
  // as n*4 features plus the two centroid coordinates (large magnitudes, kept in float64)
  a->Dm = (Dr % 4 == 2) ? Dr - 2 : (Dr & ~3);
  a->D = D; a->Dr = Dr; a->Dc = p.Dc; a->K = K; a->srow = p.srow; a->copy16 = p.copy16;
  a->TR = p.TR; a->logTR = p.logTR; a->Kc = p.Kc; a->part_bytes = p.part_bytes;
  a->buf_rows = p.buf_rows;
  a->xcols = 0;
  a->ub = nullptr; a->lb = nullptr;
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

#define KM_LAUNCH(KERNEL, XT, KT, NS2, R, MINB, ARGS, GRID)                               \
  do {                                                                                    \
    int rc__ = set_smem(KERNEL<XT, KT, NS2, R, MINB>, plan.smem);                         \
    if (rc__) return rc__;                                                                \
    KERNEL<XT, KT, NS2, R, MINB><<<GRID, KM_THREADS, plan.smem, stream>>>(ARGS);          \
  } while (0)

#define KM_DISPATCH(KERNEL, ARGS, GRID)                                                   \
  do {                                                                                    \
    if (plan.variant == 0 && plan.R == 2) KM_LAUNCH(KERNEL, float, 4, 2, 2, 2, ARGS, GRID); \
    else if (plan.variant == 0) KM_LAUNCH(KERNEL, float, 4, 2, 4, 1, ARGS, GRID);         \
    else if (plan.variant == 1 && plan.R == 2) KM_LAUNCH(KERNEL, float, 8, 4, 2, 1, ARGS, GRID); \
    else if (plan.variant == 1) KM_LAUNCH(KERNEL, float, 8, 4, 1, 1, ARGS, GRID);         \
    else KM_LAUNCH(KERNEL, double, 8, 4, 1, 1, ARGS, GRID);                               \
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
  g.gco = nullptr; g.counters = nullptr; g.totals = nullptr; g.centers_rw = nullptr;
  g.iters = nullptr; g.status_rw = nullptr; g.n_iter = 0; g.cdelta = nullptr;
  memset(&g.peer, 0, sizeof(g.peer));
  KM_DISPATCH(kmeans_sweep_kernel, g, n_chunks);
  return check_launch("kmeans_sweep");
}


static int kmeans_iterate_impl(const void* X, int x_dtype, int64_t ldx, int pos_mode, int pos_w,
                               int64_t pos_period, int64_t pos_row0, const double* w, int D, int K,
                               const int64_t* chunks, int n_chunks,
                               const int32_t* group_chunk_off, int mode, int n_iter,
                               int32_t* assign, double* partials, double* totals, double* centers,
                               int32_t* iters, int32_t* status, int32_t* counters, float* ub,
                               float* lb, double* cdelta, void* comm, cudaStream_t stream) {
  SPALIGN_REQUIRE(chunks && group_chunk_off && assign && partials && totals && centers && iters &&
                      status && counters && n_chunks > 0 && mode >= 0 && mode <= 2,
                  "kmeans_iterate: bad arguments");
  Plan plan;
  const int Dr = D - (pos_mode ? 2 : 0);
  if (!make_plan(x_dtype, D, Dr, K, &plan)) {
    set_error("kmeans_iterate: D=%d does not fit shared memory", D);
    return SPALIGN_E_UNSUPPORTED;
  }
  if (mode == 2 && plan.TR > 16) mode = 1;  // two lanes per row need TR <= 16: recompute instead
  SweepArgs g;
  int rc = fill_args(&g.a, plan, X, x_dtype, ldx, pos_mode, pos_w, pos_period, pos_row0, w, D, K);
  if (rc) return rc;
  g.chunks = chunks; g.centers = centers; g.mode = mode; g.assign = assign; g.status = status;
  g.partials = partials;
  g.gco = group_chunk_off; g.counters = counters; g.totals = totals; g.centers_rw = centers;
  g.iters = iters; g.status_rw = status; g.n_iter = n_iter;
  g.cdelta = (ub && lb) ? cdelta : nullptr;
  g.a.ub = cdelta ? ub : nullptr; g.a.lb = cdelta ? lb : nullptr;
  memset(&g.peer, 0, sizeof(g.peer));
  if (comm != nullptr) {
    rc = comm_fill_peer(static_cast<spalign_comm_t*>(comm), (long long)K * (D + 2) + 1, &g.peer);
    if (rc) return rc;
  }
  KM_DISPATCH(kmeans_sweep_kernel, g, n_chunks);
  return check_launch("kmeans_iterate");
}

extern "C" int spalign_kmeans_iterate(const void* X, int x_dtype, int64_t ldx, int pos_mode,
                                      int pos_w, int64_t pos_period, int64_t pos_row0,
                                      const double* w, int D, int K, const int64_t* chunks,
                                      int n_chunks, const int32_t* group_chunk_off, int mode,
                                      int n_iter, int32_t* assign, double* partials,
                                      double* totals, double* centers, int32_t* iters,
                                      int32_t* status, int32_t* counters,
                                      float* ub, float* lb, double* cdelta,
                                      spalign_stream_t stream_) {
  return kmeans_iterate_impl(X, x_dtype, ldx, pos_mode, pos_w, pos_period, pos_row0, w, D, K,
                             chunks, n_chunks, group_chunk_off, mode, n_iter, assign, partials,
                             totals, centers, iters, status, counters, ub, lb, cdelta, nullptr,
                             static_cast<cudaStream_t>(stream_));
}

extern "C" int spalign_kmeans_iterate_dist(const void* X, int x_dtype, int64_t ldx, int pos_mode,
                                           int pos_w, int64_t pos_period, int64_t pos_row0,
                                           const double* w, int D, int K, const int64_t* chunks,
                                           int n_chunks, const int32_t* group_chunk_off, int mode,
                                           int n_iter, int32_t* assign, double* partials,
                                           double* totals, double* centers, int32_t* iters,
                                           int32_t* status, int32_t* counters, float* ub,
                                           float* lb, double* cdelta, spalign_comm_t* comm,
                                           spalign_stream_t stream_) {
  SPALIGN_REQUIRE(comm != nullptr, "kmeans_iterate_dist: NULL communicator");
  return kmeans_iterate_impl(X, x_dtype, ldx, pos_mode, pos_w, pos_period, pos_row0, w, D, K,
                             chunks, n_chunks, group_chunk_off, mode, n_iter, assign, partials,
                             totals, centers, iters, status, counters, ub, lb, cdelta, comm,
                             static_cast<cudaStream_t>(stream_));
}

extern "C" int spalign_kmeans_finish(const void* X, int x_dtype, int64_t ldx, int pos_mode,
                                     int pos_w, int64_t pos_period, int64_t pos_row0,
                                     const double* w, int D, int K, const int64_t* group_off,
                                     int G, int n_iter, int32_t* assign, double* totals,
                                     double* centers, int32_t* iters, int32_t* status, float* ub,
                                     float* lb, double* cdelta, int slice_iters,
                                     int rows_per_set, spalign_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPALIGN_REQUIRE(group_off && assign && totals && centers && iters && status && ub && lb &&
                      cdelta && G > 0 && n_iter >= 0 && slice_iters >= 0 &&
                      (rows_per_set == 0 || rows_per_set == 2 || rows_per_set == 4),
                  "kmeans_finish: bad arguments");
  SPALIGN_REQUIRE(x_dtype == SPALIGN_F32, "kmeans_finish: fp32 rows only");
  Plan plan;
  const int Dr = D - (pos_mode ? 2 : 0);
  const int sparse_r = rows_per_set == 0 ? 4 : rows_per_set;
  if (!make_plan(x_dtype, D, Dr, K, &plan, 16, sparse_r)) {
    set_error("kmeans_finish: D=%d does not fit shared memory", D);
    return SPALIGN_E_UNSUPPORTED;
  }
  TailArgs g;
  int rc = fill_args(&g.a, plan, X, x_dtype, ldx, pos_mode, pos_w, pos_period, pos_row0, w, D, K);
  if (rc) return rc;
  g.a.ub = ub; g.a.lb = lb;
  // Thread t of phase 2 owns columns 2t, 2t+1 (+512, ...).  A row of 512n + 1 or 512n + 2 stored
  // columns (the 514-column descriptors) leaves one thread -- and with it its whole warp -- a
  // second pass over every list entry for those last columns alone.  When only the short
  // changed-row tiles of the finish kernel run, they are summed by the lanes that keep the
  // per-cluster scalars instead (same accumulation order, same bits); in the full sweeps that
  // serial loop would cost more than the second pass (measured), so only here.
  g.a.xcols = (!pos_mode && Dr > 2 * KM_THREADS && Dr % (2 * KM_THREADS) <= 2)
                  ? Dr % (2 * KM_THREADS) : 0;
  g.group_off = group_off; g.assign = assign; g.totals = totals; g.centers = centers;
  g.iters = iters; g.status = status; g.cdelta = cdelta; g.n_iter = n_iter; g.slice = slice_iters;
  KM_DISPATCH(kmeans_tail_kernel, g, G);
  return check_launch("kmeans_finish");
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

// Diagnostics (synchronises the device): out[0] = rows that went through the fp32 screening
// pass since the last reset, out[1] = rows that needed the exact float64 pass.
extern "C" int spalign_kmeans_debug_stats(int64_t* out_host, int reset) {
#ifdef KM_PROFILE
  {
    unsigned long long p[8];
    cudaMemcpyFromSymbol(p, g_km_prof, sizeof(p));
    if (p[6]) {
      fprintf(stderr, "[km_profile] tiles=%llu  cycles/tile (thread 0): barrierA %.0f | issue %.0f | tile wait %.0f | phase1+barrierB %.0f | grouping %.0f | phase2 %.0f\n",
              p[6], (double)p[0] / p[6], (double)p[5] / p[6], (double)p[1] / p[6], (double)p[2] / p[6], (double)p[3] / p[6], (double)p[4] / p[6]);
    }
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_km_prof, z, sizeof(z));
    {
      unsigned long long p4[16], z4[16] = {0};
      cudaMemcpyFromSymbol(p4, g_km_prof4, sizeof(p4));
      for (int m = 0; m < 2; ++m) {
        const unsigned long long* q = p4 + 8 * m;
        if (q[6])
          fprintf(stderr, "[km_profile] full sweep mode %d: CTAs=%llu tiles=%llu  cycles/tile (thread 0): barrierA %.0f | issue %.0f | tile wait %.0f | phase1+decide %.0f | grouping %.0f | phase2 %.0f\n",
                  m, q[7], q[6], (double)q[0] / q[6], (double)q[5] / q[6], (double)q[1] / q[6], (double)q[2] / q[6], (double)q[3] / q[6], (double)q[4] / q[6]);
      }
      cudaMemcpyToSymbol(g_km_prof4, z4, sizeof(z4));
    }
    cudaMemcpyFromSymbol(p, g_km_prof2, sizeof(p));
    if (p[4])
      fprintf(stderr, "[km_profile] finish kernel: iterations=%llu  cycles/iteration (thread 0): prepare %.0f | bounds %.0f | sweep %.0f | update %.0f\n",
              p[4], (double)p[0] / p[4], (double)p[1] / p[4], (double)p[2] / p[4], (double)p[3] / p[4]);
    cudaMemcpyToSymbol(g_km_prof2, z, sizeof(z));
    {
      unsigned long long slow = 0, zero = 0;
      cudaMemcpyFromSymbol(&slow, g_km_slow, sizeof(slow));
      cudaMemcpyToSymbol(g_km_slow, &zero, sizeof(zero));
      fprintf(stderr, "[km_profile] rows that took the float64 decision path: %llu\n", slow);
    }
    {
      static unsigned long long tr[3 * 1024];
      cudaMemcpyFromSymbol(tr, g_km_trace, sizeof(tr));
      unsigned long long t0 = ~0ull, t1 = 0;
      int n = 0;
      for (int g = 0; g < 1024; ++g)
        if (tr[g * 3 + 1]) { ++n; if (tr[g * 3] < t0) t0 = tr[g * 3]; if (tr[g * 3 + 1] > t1) t1 = tr[g * 3 + 1]; }
      if (n) {
        fprintf(stderr, "[km_profile] finish kernel trace: %d groups, span %.1f us; per group (start us, run us, iterations), sorted by end:\n", n, (t1 - t0) / 1e3);
        // print the 12 groups that end last and a few aggregate numbers
        double sum_run = 0; int late = 0;
        for (int g = 0; g < 1024; ++g) if (tr[g * 3 + 1]) { sum_run += (tr[g * 3 + 1] - tr[g * 3]) / 1e3; if (tr[g * 3] - t0 > 20000) ++late; }
        fprintf(stderr, "[km_profile]   mean run %.1f us, groups that started > 20 us late: %d\n", sum_run / n, late);
        for (int k = 0; k < 12; ++k) {
          int best = -1;
          for (int g = 0; g < 1024; ++g) if (tr[g * 3 + 1] && (best < 0 || tr[g * 3 + 1] > tr[best * 3 + 1])) best = g;
          if (best < 0) break;
          fprintf(stderr, "[km_profile]   group %d: start %.1f run %.1f iters %llu (%.1f us/iter)\n", best, (tr[best * 3] - t0) / 1e3, (tr[best * 3 + 1] - tr[best * 3]) / 1e3, tr[best * 3 + 2], (tr[best * 3 + 1] - tr[best * 3]) / 1e3 / (tr[best * 3 + 2] > 1 ? tr[best * 3 + 2] - 1 : 1));
          tr[best * 3 + 1] = 0;
        }
      }
      static unsigned long long zz[3 * 1024];
      cudaMemcpyToSymbol(g_km_trace, zz, sizeof(zz));
    }
    cudaMemcpyFromSymbol(p, g_km_prof3, sizeof(p));
    if (p[7])
      fprintf(stderr, "[km_profile] sparse sweep, warp 0: sets=%llu  cycles/set: issue %.0f | wait %.0f | partial %.0f | stage %.0f ; exact+compaction per set %.0f ; decide calls %llu, cycles each %.0f\n",
              p[7], (double)p[0] / p[7], (double)p[1] / p[7], (double)p[2] / p[7], (double)p[3] / p[7], (double)p[4] / p[7], p[6], (double)p[5] / (p[6] ? p[6] : 1));
    cudaMemcpyToSymbol(g_km_prof3, z, sizeof(z));
    {
      Plan plan;
      if (make_plan(SPALIGN_F32, 514, 514, 4, &plan, 16)) {
        int nb_tail = 0, nb_sweep = 0;
        set_smem(kmeans_tail_kernel<float, 4, 2, 2, 2>, plan.smem);
        set_smem(kmeans_sweep_kernel<float, 4, 2, 2, 2>, plan.smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_tail, kmeans_tail_kernel<float, 4, 2, 2, 2>,
                                                      KM_THREADS, plan.smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_sweep, kmeans_sweep_kernel<float, 4, 2, 2, 2>,
                                                      KM_THREADS, plan.smem);
        cudaFuncAttributes fa, fb;
        cudaFuncGetAttributes(&fa, kmeans_tail_kernel<float, 4, 2, 2, 2>);
        cudaFuncGetAttributes(&fb, kmeans_sweep_kernel<float, 4, 2, 2, 2>);
        fprintf(stderr, "[km_profile] D=514 K=4: dynamic smem %zu B; CTAs/SM: finish kernel %d (static %zu B, %d regs), sweep kernel %d (static %zu B, %d regs)\n",
                plan.smem, nb_tail, fa.sharedSizeBytes, fa.numRegs, nb_sweep, fb.sharedSizeBytes, fb.numRegs);
      }
    }
  }
#endif
  unsigned long long h[2] = {0, 0};
  SPALIGN_CUDA(cudaMemcpyFromSymbol(h, g_km_stats, sizeof(h)));
  if (out_host) {
    out_host[0] = (int64_t)h[0];
    out_host[1] = (int64_t)h[1];
  }
  if (reset) {
    unsigned long long z[2] = {0, 0};
    SPALIGN_CUDA(cudaMemcpyToSymbol(g_km_stats, z, sizeof(z)));
  }
  return SPALIGN_OK;
}
