// Hot-path kernels (sm_100a).  Round-1 layout: CSR of the owned full-matrix rows, one
// warp reduces one chunk (<= kChunk stored cells of one row), a second tiny kernel adds
// the chunk partials of each row in a fixed order -> deterministic, atomic-free.
//
//   K1  k_chunk_reduce<F64>      s_i = sum_j v_ij * x_j          (normalize_hic.py:136-143,154-162)
//   K2  k_ice_reduce/_update     S = x.s, meanS/min/max/dev, B *= S/meanS, x = 1/B (:146-151,217-222)
//   K3  K1 + k_dot               sum v/(b_i b_j)                 (hic_data.py:350-375)
//   K4  k_chunk_reduce<I64*>     row sums / masked column sums   (hic_filtering.py:36-39,141-145,185-200)
//   K5  k_diag_hist              per-distance sums               (normalize_hic.py:271-283)
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace tb {

namespace {

enum { M_F64 = 0, M_I64 = 1, M_I64_MASKED = 2 };

__device__ __forceinline__ int32_t ld_stream(const int32_t *p) {
  int32_t r;   // streamed once: keep it out of L1 so the gathered vector stays resident
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(256)
k_chunk_reduce(const int32_t *__restrict__ col, const int32_t *__restrict__ val,
               const int64_t *__restrict__ chunk_beg, const int32_t *__restrict__ chunk_row,
               const int64_t *__restrict__ rowptr, int64_t nchunks,
               const double *__restrict__ x, const uint8_t *__restrict__ bad,
               double *__restrict__ pf, long long *__restrict__ pi,
               const IceState *__restrict__ state) {
  if (state && state->stopped) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < nchunks; c += nwarp) {
    const int64_t beg = chunk_beg[c], end = chunk_beg[c + 1];
    double fa = 0.0, fb = 0.0;
    long long ia = 0;
    int64_t k = beg + lane;
    for (; k + 96 < end; k += 128) {
      int32_t c0 = ld_stream(col + k), c1 = ld_stream(col + k + 32);
      int32_t c2 = ld_stream(col + k + 64), c3 = ld_stream(col + k + 96);
      int32_t v0 = ld_stream(val + k), v1 = ld_stream(val + k + 32);
      int32_t v2 = ld_stream(val + k + 64), v3 = ld_stream(val + k + 96);
      if (MODE == M_F64) {
        double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
        fa = fma((double)v0, x0, fa);
        fb = fma((double)v1, x1, fb);
        fa = fma((double)v2, x2, fa);
        fb = fma((double)v3, x3, fb);
      } else if (MODE == M_I64) {
        ia += (long long)v0 + v1 + v2 + v3;
      } else {
        ia += (bad[c0] ? 0 : v0) + (long long)(bad[c1] ? 0 : v1) + (bad[c2] ? 0 : v2) +
              (long long)(bad[c3] ? 0 : v3);
      }
    }
    for (; k < end; k += 32) {
      int32_t c0 = ld_stream(col + k), v0 = ld_stream(val + k);
      if (MODE == M_F64) {
        double x0 = __ldg(x + c0);
        fa = fma((double)v0, x0, fa);
      } else if (MODE == M_I64) {
        ia += v0;
      } else {
        ia += bad[c0] ? 0 : v0;
      }
    }
    if (MODE == M_F64) {
      fa += fb;
      for (int o = 16; o; o >>= 1) fa += __shfl_xor_sync(0xffffffffu, fa, o);
      if (lane == 0) pf[c] = fa;
    } else {
      for (int o = 16; o; o >>= 1) ia += __shfl_xor_sync(0xffffffffu, ia, o);
      if (lane == 0) pi[c] = ia;
    }
  }
}

// sum over the work units of the row's block of the sparse-tile partials (tiles.cu), unit order
struct TileSums {
  const double *partial;        // null when the store has no sparse tiles
  const int32_t *unit_of_rb;
};
__device__ __forceinline__ double tile_sum(const TileSums &ts, int64_t r) {
  if (!ts.partial) return 0.0;
  const int64_t rb = r / kTileRows;
  double a = 0.0;
  for (int32_t u = ts.unit_of_rb[rb]; u < ts.unit_of_rb[rb + 1]; ++u)
    a += ts.partial[(int64_t)u * kTileRows + (r % kTileRows)];
  return a;
}

// out[rb + r] = sum of the row's chunk partials in chunk order + its sparse-tile sums
__global__ void k_combine_f64(const double *__restrict__ pf, const int64_t *__restrict__ row_chunk,
                              int64_t nrows, int64_t rb, double *__restrict__ out,
                              const IceState *__restrict__ state, TileSums ts) {
  if (state && state->stopped) return;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x) {
    double a = 0.0;
    for (int64_t c = row_chunk[r]; c < row_chunk[r + 1]; ++c) a += pf[c];
    a += tile_sum(ts, r);
    out[rb + r] = a;
  }
}

__global__ void k_combine_i64(const long long *__restrict__ pi, const int64_t *__restrict__ row_chunk,
                              int64_t nrows, int64_t rb, long long *__restrict__ out) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x) {
    long long a = 0;
    for (int64_t c = row_chunk[r]; c < row_chunk[r + 1]; ++c) a += pi[c];
    out[rb + r] = a;
  }
}

__global__ void k_row_nnz(const int64_t *__restrict__ rowptr, int64_t nrows, int64_t rb,
                          long long *__restrict__ out) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x)
    out[rb + r] = rowptr[r + 1] - rowptr[r];
}

// ---------------------------------------------------------------------------------
// ICE vector kernels
// ---------------------------------------------------------------------------------
// "Present" rows (copy_matrix, normalize_hic.py:165-177): non-bad rows that keep at least one
// stored cell whose partner is not bad.  live_i = stored cells of row i - cells of row i whose
// partner is bad; by symmetry the second term is collected by walking only the BAD rows (a few
// per cent of the matrix).  Counts are exact small integers carried as doubles so that they can
// ride in the same all-reduce as the row sums.
__global__ void k_live_init(const int64_t *__restrict__ rowptr, int64_t nrows, int64_t rb,
                            double *__restrict__ live) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x)
    live[rb + r] = (double)(rowptr[r + 1] - rowptr[r]);
}

__global__ void __launch_bounds__(256)
k_bad_partners(const int32_t *__restrict__ col, const int64_t *__restrict__ chunk_beg,
               const int32_t *__restrict__ chunk_row, int64_t nchunks, int64_t rb,
               const uint8_t *__restrict__ bad, double *__restrict__ live) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < nchunks; c += nwarp) {
    if (!bad[rb + chunk_row[c]]) continue;
    for (int64_t k = chunk_beg[c] + lane; k < chunk_beg[c + 1]; k += 32)
      atomicAdd(&live[col[k]], -1.0);       // integers: the order of the additions is irrelevant
  }
}

__global__ void k_ice_init(const uint8_t *__restrict__ bad, int64_t n, double *x, double *b,
                           IceState *state) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    x[i] = bad[i] ? 0.0 : 1.0;
    b[i] = 1.0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    state->mean_s = state->s_min = state->s_max = state->dev = 0.0;
    state->n_present = 0;
    state->passes_done = 0;
    state->stopped = 0;
    state->all_bad = 0;
    state->blocks_arrived = 0;
  }
}

constexpr int kRedThreads = 256;

// S_i = x_i * s_i over present rows -> sum / min / max; the last block to arrive folds the
// per-block partials in block order and takes the reference's loop decision.
// FUSED (single GPU): the row's chunk partials are added here instead of in k_combine_f64
// (one kernel and one pass over s less); s_full is still written for k_ice_update.
template <bool FUSED>
__global__ void __launch_bounds__(kRedThreads)
k_ice_reduce(double *__restrict__ s_full, const double *__restrict__ x,
             const uint8_t *__restrict__ bad, uint8_t *__restrict__ present, int64_t n,
             int pass, int iterations, double max_dev, double *__restrict__ block_red,
             double *__restrict__ trace, IceState *state, const double *__restrict__ pf,
             const int64_t *__restrict__ row_chunk, TileSums ts) {
  if (state->stopped) return;
  __shared__ double sh_sum[kRedThreads / 32], sh_min[kRedThreads / 32], sh_max[kRedThreads / 32];
  __shared__ long long sh_cnt[kRedThreads / 32];
  __shared__ bool is_last;
  double sum = 0.0, mn = INFINITY, mx = -INFINITY;
  long long cnt = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    bool p;
    double si;
    if (FUSED) {
      double a = 0.0;
      for (int64_t c = row_chunk[i]; c < row_chunk[i + 1]; ++c) a += pf[c];
      a += tile_sum(ts, i);
      s_full[i] = si = a;
      if (pass == 0) present[i] = p = !bad[i] && s_full[n + i] > 0.0;
      else p = present[i];
    } else {
      si = s_full[i];
      if (pass == 0) {
        p = !bad[i] && s_full[n + i] > 0.0;    // row kept by copy_matrix (normalize_hic.py:165-177)
        present[i] = p;
      } else {
        p = present[i];
      }
    }
    if (p) {
      double S = x[i] * si;
      sum += S;
      mn = fmin(mn, S);
      mx = fmax(mx, S);
      ++cnt;
    }
  }
  for (int o = 16; o; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh_sum[w] = sum; sh_min[w] = mn; sh_max[w] = mx; sh_cnt[w] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < kRedThreads / 32; ++k) {
      sum += sh_sum[k]; mn = fmin(mn, sh_min[k]); mx = fmax(mx, sh_max[k]); cnt += sh_cnt[k];
    }
    block_red[4 * blockIdx.x + 0] = sum;
    block_red[4 * blockIdx.x + 1] = mn;
    block_red[4 * blockIdx.x + 2] = mx;
    block_red[4 * blockIdx.x + 3] = (double)cnt;
    __threadfence();
    unsigned int t = atomicAdd(&state->blocks_arrived, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last || threadIdx.x >= 32) return;
  __threadfence();
  // last block: one warp folds the per-block partials, lane l takes blocks l, l+32, ... and the
  // lanes are then combined by a fixed butterfly -> same bits every run
  sum = 0.0; mn = INFINITY; mx = -INFINITY;
  double dcnt = 0.0;
  for (unsigned int bl = threadIdx.x; bl < gridDim.x; bl += 32) {
    const volatile double *br = block_red + 4 * bl;
    sum += br[0]; mn = fmin(mn, br[1]); mx = fmax(mx, br[2]); dcnt += br[3];
  }
  for (int o = 16; o; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    dcnt += __shfl_xor_sync(0xffffffffu, dcnt, o);
  }
  if (threadIdx.x != 0) return;
  state->blocks_arrived = 0;
  if (pass == 0) state->n_present = (long long)dcnt;
  const long long np = state->n_present;
  if (np == 0) {
    state->all_bad = 1;
    state->stopped = 1;
    return;
  }
  const double mean = sum / (double)np;
  state->mean_s = mean;
  state->s_min = mn;
  state->s_max = mx;
  if (iterations == 0) {            // single pass, no deviation test (normalize_hic.py:214-215)
    state->passes_done = pass + 1;
    state->stopped = 1;
    return;
  }
  const double dev = fmax(fabs(mn / mean - 1.0), fabs(mx / mean - 1.0));
  state->dev = dev;
  trace[4 * pass + 0] = mn;
  trace[4 * pass + 1] = mean;
  trace[4 * pass + 2] = mx;
  trace[4 * pass + 3] = dev;
  state->passes_done = pass + 1;
  if (dev < max_dev) state->stopped = 1;
}

// B_i *= S_i / meanS ; x_i = 1 / B_i.  Runs for every pass whose statistics were finalised,
// including the one that stops the loop (the reference updates B before it tests dev).
__global__ void k_ice_update(const double *__restrict__ s_full, double *__restrict__ x,
                             double *__restrict__ b, const uint8_t *__restrict__ present,
                             int64_t n, int pass, const IceState *__restrict__ state) {
  if (state->passes_done != pass + 1 || state->all_bad) return;
  const double mean = state->mean_s;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (!present[i]) {
      x[i] = 0.0;
      continue;
    }
    const double S = x[i] * s_full[i];
    const double bi = b[i] * (S / mean);
    b[i] = bi;
    x[i] = (bi != 0.0) ? 1.0 / bi : 0.0;     // B == 0: row frozen by the ZeroDivisionError skip
  }
}

__global__ void k_ice_finalize(const double *__restrict__ b, const uint8_t *__restrict__ present,
                               int64_t n, const IceState *__restrict__ state,
                               double *__restrict__ bias) {
  const double f = sqrt(state->mean_s);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double bi = b[i];
    bias[i] = (present[i] && bi != 0.0) ? bi * f : 1.0;   // normalize_hic.py:223-230
  }
}

__global__ void k_inverse_bias(const double *__restrict__ bias, const uint8_t *__restrict__ bad,
                               int64_t n, double *__restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    y[i] = bad[i] ? 0.0 : 1.0 / bias[i];
}

// sum over owned rows of y_i * s_i, single block (N fits: one pass, fixed order)
__global__ void __launch_bounds__(1024)
k_dot_rows(const double *__restrict__ y, const double *__restrict__ s, int64_t rb, int64_t re,
           double *__restrict__ out) {
  __shared__ double sh[32];
  double a = 0.0;
  for (int64_t i = rb + threadIdx.x; i < re; i += blockDim.x) a += y[i] * s[i];
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) a += sh[k];
    out[0] = a;
  }
}

__global__ void __launch_bounds__(1024)
k_sum_good_rows(const long long *__restrict__ v, const uint8_t *__restrict__ bad, int64_t rb,
                int64_t re, long long *__restrict__ out) {
  __shared__ long long sh[32];
  long long a = 0;
  for (int64_t i = rb + threadIdx.x; i < re; i += blockDim.x) a += bad[i] ? 0 : v[i];
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) a += sh[k];
    out[0] = a;
  }
}

// ---------------------------------------------------------------------------------
// K5: per-distance sums.  Cells of one row have distinct distances, and a warp walks
// consecutive cells of one row, so the shared-memory histogram sees no intra-warp
// collisions; distances beyond the shared window go straight to L2 atomics.
// ---------------------------------------------------------------------------------
constexpr int kDiagShared = 4096;

__global__ void __launch_bounds__(256)
k_diag_hist(const int32_t *__restrict__ col, const int32_t *__restrict__ val,
            const int64_t *__restrict__ chunk_beg, const int32_t *__restrict__ chunk_row,
            const int64_t *__restrict__ rowptr, int64_t nchunks, int64_t rb,
            const uint8_t *__restrict__ bad, const int32_t *__restrict__ chrom_of, int64_t ndiag,
            unsigned long long *__restrict__ hist) {
  __shared__ unsigned long long sh[kDiagShared];
  for (int k = threadIdx.x; k < kDiagShared; k += blockDim.x) sh[k] = 0ull;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < nchunks; c += nwarp) {
    const int64_t r = chunk_row[c];
    const int64_t g = rb + r;
    if (bad[g]) continue;
    const int64_t beg = chunk_beg[c], end = chunk_beg[c + 1];
    if (col[end - 1] < g) continue;            // chunk entirely below the diagonal
    const int32_t cg = chrom_of[g];
    for (int64_t k = beg + lane; k < end; k += 32) {
      const int64_t j = ld_stream(col + k);
      const int64_t d = j - g;
      if (d < 0 || d >= ndiag || chrom_of[j] != cg) continue;
      const unsigned long long v = (unsigned long long)(long long)ld_stream(val + k);
      if (d < kDiagShared) atomicAdd(&sh[d], v);
      else atomicAdd(&hist[d], v);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kDiagShared && k < ndiag; k += blockDim.x)
    if (sh[k]) atomicAdd(&hist[k], sh[k]);
}

inline int grid_1d(int64_t n, int threads, int sm_count, int per_sm = 8) {
  int64_t g = (n + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

inline int chunk_grid(const tb_store *s) {
  // 8 warps per CTA, 8 CTAs per SM resident; grid = multiple of the SM count
  int64_t need = (s->nchunks + 7) / 8;
  int64_t cap = (int64_t)s->sm_count * 8;
  int64_t g = need < cap ? need : cap;
  if (g < 1) g = 1;
  return (int)g;
}

void upload_bad(tb_store *s, const uint8_t *h_bad) {
  if (h_bad)
    TB_CUDA(cudaMemcpyAsync(s->bad.p, h_bad, s->nbins, cudaMemcpyHostToDevice, s->stream));
  else
    TB_CUDA(cudaMemsetAsync(s->bad.p, 0, s->nbins, s->stream));
}

template <int MODE>
void launch_chunks(tb_store *s, const double *x, const uint8_t *bad, const IceState *state) {
  if (s->nchunks == 0) return;
  k_chunk_reduce<MODE><<<chunk_grid(s), 256, 0, s->stream>>>(
      s->col.p, s->val.p, s->chunk_beg.p, s->chunk_row.p, s->rowptr.p, s->nchunks, x, bad,
      s->partial_f.p, s->partial_i.p, state);
}

// ---------------------------------------------------------------------------------
// Batch ICE: every section of the store is an independent matrix (block-diagonal store made
// by tb_store_create_batch).  K1 is unchanged -- rows are rows -- but the statistics, the
// stop decision and the bias update are per section, and a section that has converged is
// frozen while the others go on (BASELINE configs[4]: many experiments normalised at once).
// ---------------------------------------------------------------------------------
struct SegState {
  double mean_s, dev;
  long long n_present;
  int passes_done;      // passes whose statistics were finalised for this section
  int stopped, all_bad, pad;
};

__global__ void k_seg_init(SegState *seg, int nseg, unsigned int *n_stopped) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nseg; k += gridDim.x * blockDim.x) {
    seg[k].mean_s = seg[k].dev = 0.0;
    seg[k].n_present = 0;
    seg[k].passes_done = seg[k].stopped = seg[k].all_bad = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_stopped = 0;
}

// one CTA per section: combine the chunk partials of its rows, S = x*s, sum/min/max, decision
__global__ void __launch_bounds__(1024)
k_seg_reduce(double *__restrict__ s_full, const double *__restrict__ x,
             const uint8_t *__restrict__ bad, uint8_t *__restrict__ present,
             const int64_t *__restrict__ seg_off, int nseg, int64_t nbins, int pass, int iterations,
             double max_dev,
             double *__restrict__ trace, int npass_max, SegState *seg, unsigned int *n_stopped,
             IceState *state, const double *__restrict__ pf,
             const int64_t *__restrict__ row_chunk, TileSums ts) {
  const int k = blockIdx.x;
  if (seg[k].stopped) return;
  __shared__ double sh_sum[32], sh_min[32], sh_max[32];
  __shared__ long long sh_cnt[32];
  double sum = 0.0, mn = INFINITY, mx = -INFINITY;
  long long cnt = 0;
  for (int64_t i = seg_off[k] + threadIdx.x; i < seg_off[k + 1]; i += blockDim.x) {
    double a = 0.0;
    for (int64_t c = row_chunk[i]; c < row_chunk[i + 1]; ++c) a += pf[c];
    a += tile_sum(ts, i);
    s_full[i] = a;
    bool p;
    if (pass == 0) present[i] = p = !bad[i] && s_full[nbins + i] > 0.0;
    else p = present[i];
    if (p) {
      const double S = x[i] * a;
      sum += S; mn = fmin(mn, S); mx = fmax(mx, S); ++cnt;
    }
  }
  for (int o = 16; o; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh_sum[w] = sum; sh_min[w] = mn; sh_max[w] = mx; sh_cnt[w] = cnt; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int q = 1; q < (int)(blockDim.x >> 5); ++q) {
    sum += sh_sum[q]; mn = fmin(mn, sh_min[q]); mx = fmax(mx, sh_max[q]); cnt += sh_cnt[q];
  }
  SegState &g = seg[k];
  if (pass == 0) g.n_present = cnt;
  bool stop = false;
  if (g.n_present == 0) {
    g.all_bad = 1;
    stop = true;
  } else {
    const double mean = sum / (double)g.n_present;
    g.mean_s = mean;
    g.passes_done = pass + 1;
    if (iterations == 0) {
      stop = true;
    } else {
      const double dev = fmax(fabs(mn / mean - 1.0), fabs(mx / mean - 1.0));
      g.dev = dev;
      double *t = trace + ((int64_t)k * npass_max + pass) * 4;
      t[0] = mn; t[1] = mean; t[2] = mx; t[3] = dev;
      stop = dev < max_dev;
    }
  }
  if (stop) {
    g.stopped = 1;
    if (atomicAdd(n_stopped, 1u) + 1u == (unsigned)nseg) state->stopped = 1;
  }
}

__global__ void k_seg_update(const double *__restrict__ s_full, double *__restrict__ x,
                             double *__restrict__ b, const uint8_t *__restrict__ present,
                             const int32_t *__restrict__ seg_of, int64_t n, int pass,
                             const SegState *__restrict__ seg) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const SegState g = seg[seg_of[i]];
    if (g.passes_done != pass + 1 || g.all_bad) continue;     // frozen or failed section
    if (!present[i]) { x[i] = 0.0; continue; }
    const double bi = b[i] * ((x[i] * s_full[i]) / g.mean_s);
    b[i] = bi;
    x[i] = (bi != 0.0) ? 1.0 / bi : 0.0;
  }
}

__global__ void k_seg_finalize(const double *__restrict__ b, const uint8_t *__restrict__ present,
                               const int32_t *__restrict__ seg_of, int64_t n,
                               const SegState *__restrict__ seg, double *__restrict__ bias) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const SegState g = seg[seg_of[i]];
    const double bi = b[i];
    bias[i] = (!g.all_bad && present[i] && bi != 0.0) ? bi * sqrt(g.mean_s) : 1.0;
  }
}

// A/B switch: walk the uncompressed CSR in every pass instead of the compact stream.  Set per
// store with tb_store_set_option("rowsum", "csr"); the environment TB_ROWSUM=csr sets the default.
bool use_csr_rowsum(const tb_store *s) {
  if (s->rowsum_mode >= 0) return s->rowsum_mode == 1;
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("TB_ROWSUM");
    v = (e && !strcmp(e, "csr")) ? 1 : 0;
  }
  return v == 1;
}

}  // namespace

// ---------------------------------------------------------------------------------
void row_stats(tb_store *s, int64_t *h_nnz_row, int64_t *h_row_sum) {
  cudaStream_t st = s->stream;
  const int64_t N = s->nbins;
  long long *dst = s->world > 1 ? s->i_part.p : s->i_full.p;
  if (s->world == 1) TB_CUDA(cudaMemsetAsync(dst, 0, 2 * N * sizeof(long long), st));
  launch_chunks<M_I64>(s, nullptr, nullptr, nullptr);
  if (s->nrows) {
    k_row_nnz<<<grid_1d(s->nrows, 256, s->sm_count), 256, 0, st>>>(s->rowptr.p, s->nrows,
                                                                   s->row_begin, dst);
    k_combine_i64<<<grid_1d(s->nrows, 256, s->sm_count), 256, 0, st>>>(
        s->partial_i.p, s->row_chunk.p, s->nrows, s->row_begin, dst + N);
  }
  if (s->world > 1) allreduce_i64(s, s->i_part.p, s->i_full.p, 2 * N);
  TB_CUDA(cudaMemcpyAsync(h_nnz_row, s->i_full.p, N * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(h_row_sum, s->i_full.p + N, N * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
}

void col_sums_masked(tb_store *s, const uint8_t *h_bad, int64_t *h_col_sum) {
  cudaStream_t st = s->stream;
  const int64_t N = s->nbins;
  upload_bad(s, h_bad);
  long long *dst = s->world > 1 ? s->i_part.p : s->i_full.p;
  if (s->world == 1) TB_CUDA(cudaMemsetAsync(dst, 0, N * sizeof(long long), st));
  launch_chunks<M_I64_MASKED>(s, nullptr, s->bad.p, nullptr);
  if (s->nrows)
    k_combine_i64<<<grid_1d(s->nrows, 256, s->sm_count), 256, 0, st>>>(
        s->partial_i.p, s->row_chunk.p, s->nrows, s->row_begin, dst);
  if (s->world > 1) allreduce_i64(s, s->i_part.p, s->i_full.p, N);
  TB_CUDA(cudaMemcpyAsync(h_col_sum, s->i_full.p, N * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
}

// live[i] for the owned rows (+ the bad-partner corrections of any row), zeros elsewhere
void prepare_live(tb_store *s, double *live) {
  cudaStream_t st = s->stream;
  TB_CUDA(cudaMemsetAsync(live, 0, s->nbins * sizeof(double), st));
  if (!s->nrows) return;
  k_live_init<<<grid_1d(s->nrows, 256, s->sm_count), 256, 0, st>>>(s->rowptr.p, s->nrows,
                                                                   s->row_begin, live);
  if (s->nchunks)
    k_bad_partners<<<chunk_grid(s), 256, 0, st>>>(s->col.p, s->chunk_beg.p, s->chunk_row.p, s->nchunks,
                                                  s->row_begin, s->bad.p, live);
}

int ice(tb_store *s, const uint8_t *h_bad, int32_t iterations, double max_dev, double *h_bias,
        double *h_trace, int32_t *n_passes) {
  cudaStream_t st = s->stream;
  const int64_t N = s->nbins;
  TB_REQUIRE(iterations >= 0, "ice: iterations must be >= 0");
  TB_REQUIRE(N > 0, "ice: empty matrix");
  TB_REQUIRE(s->world > 1 || s->nrows == N,
             "ice: the store owns a row block only; connect the shards first (tb_comm_init)");
  const int npass_max = iterations + 1;
  if (s->trace.n < (size_t)npass_max * 4) s->trace.alloc((size_t)npass_max * 4);
  upload_bad(s, h_bad);
  const int vg = grid_1d(N, 256, s->sm_count);
  int rg = grid_1d(N, kRedThreads, s->sm_count, 4);
  if (rg > 768) rg = 768;
  k_ice_init<<<vg, 256, 0, st>>>(s->bad.p, N, s->x.p, s->b.p, s->state.p);

  std::vector<cudaEvent_t> ev(2 * (size_t)npass_max + 2);
  for (auto &e : ev) TB_CUDA(cudaEventCreate(&e));
  // TB_TIMING=1 on several GPUs: where the time between two row sums goes (combine, all-reduce,
  // statistics + update), printed per rank after the call
  const char *tenv = getenv("TB_TIMING");
  std::vector<cudaEvent_t> evx;
  if (s->world > 1 && tenv && tenv[0] == '1') {
    evx.resize(3 * (size_t)npass_max);
    for (auto &e : evx) TB_CUDA(cudaEventCreate(&e));
  }
  double *dst = s->world > 1 ? s->s_part.p : s->s_full.p;
  const bool csr_walk = use_csr_rowsum(s);
  const TileSums ts = {(!csr_walk && s->t_nunits) ? s->t_partial.p : nullptr, s->t_unit_of_rb.p};
  prepare_live(s, dst + N);
  if (ts.partial) TB_CUDA(cudaMemsetAsync(s->partial_f.p, 0, s->partial_f.bytes(), st));
  // The host only needs to know WHEN the loop stopped: passes queued after the stop are no-ops
  // on the device.  ICE converges geometrically, so the next look is scheduled where the
  // deviation is predicted to cross max_dev (at most 16 passes ahead).
  int next_poll = iterations == 0 ? 1 : 4;
  double prev_dev = -1.0;
  int prev_done = 0;
  int launched = 0;
  bool stopped = false;
  TB_CUDA(cudaEventRecord(ev[2 * npass_max], st));
  for (int p = 0; p < npass_max && !stopped; ++p) {
    TB_CUDA(cudaEventRecord(ev[2 * p], st));
    if (csr_walk) launch_chunks<M_F64>(s, s->x.p, nullptr, s->state.p);
    else launch_rowsum(s, s->x.p, s->state.p);
    TB_CUDA(cudaEventRecord(ev[2 * p + 1], st));
    if (s->world > 1) {
      if (s->nrows)
        k_combine_f64<<<grid_1d(s->nrows, 256, s->sm_count), 256, 0, st>>>(
            s->partial_f.p, s->row_chunk.p, s->nrows, s->row_begin, dst, s->state.p, ts);
      if (!evx.empty()) TB_CUDA(cudaEventRecord(evx[3 * p], st));
      allreduce_f64(s, s->s_part.p, s->s_full.p, p == 0 ? 2 * N : N);
      if (!evx.empty()) TB_CUDA(cudaEventRecord(evx[3 * p + 1], st));
      k_ice_reduce<false><<<rg, kRedThreads, 0, st>>>(
          s->s_full.p, s->x.p, s->bad.p, s->present.p, N, p, iterations, max_dev, s->block_red.p,
          s->trace.p, s->state.p, nullptr, nullptr, ts);
    } else {
      k_ice_reduce<true><<<rg, kRedThreads, 0, st>>>(
          s->s_full.p, s->x.p, s->bad.p, s->present.p, N, p, iterations, max_dev, s->block_red.p,
          s->trace.p, s->state.p, s->partial_f.p, s->row_chunk.p, ts);
    }
    k_ice_update<<<vg, 256, 0, st>>>(s->s_full.p, s->x.p, s->b.p, s->present.p, N, p, s->state.p);
    if (!evx.empty()) TB_CUDA(cudaEventRecord(evx[3 * p + 2], st));
    ++launched;
    if (p + 1 >= next_poll || p == npass_max - 1) {
      TB_CUDA(cudaMemcpyAsync(s->h_state.p, s->state.p, sizeof(IceState), cudaMemcpyDeviceToHost, st));
      TB_CUDA(cudaStreamSynchronize(st));
      const IceState &h = *s->h_state.p;
      stopped = h.stopped != 0;
      int step = 4;
      if (!stopped && prev_dev > 0.0 && h.dev > 0.0 && h.dev < prev_dev && h.passes_done > prev_done) {
        const double rate = log(prev_dev / h.dev) / (double)(h.passes_done - prev_done);
        const double remain = log(h.dev / max_dev) / rate;
        step = remain < 1.0 ? 1 : (remain > 16.0 ? 16 : (int)ceil(remain));
      }
      prev_dev = h.dev;
      prev_done = h.passes_done;
      next_poll = p + 1 + step;
    }
  }
  TB_CUDA(cudaEventRecord(ev[2 * npass_max + 1], st));
  TB_CUDA(cudaMemcpyAsync(s->h_state.p, s->state.p, sizeof(IceState), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  const IceState hs = *s->h_state.p;
  int rc = TB_OK;
  if (hs.all_bad) {
    set_error("ERROR: normalization failed, all bad columns");
    rc = TB_ERR_ALL_BAD;
  } else {
    // bias into s_part's tail is avoided: reuse partial-free buffer b -> finalize into x
    k_ice_finalize<<<vg, 256, 0, st>>>(s->b.p, s->present.p, N, s->state.p, s->x.p);
    TB_CUDA(cudaMemcpyAsync(h_bias, s->x.p, N * sizeof(double), cudaMemcpyDeviceToHost, st));
    const int done = hs.passes_done;
    if (h_trace && iterations > 0 && done > 0)
      TB_CUDA(cudaMemcpyAsync(h_trace, s->trace.p, (size_t)done * 4 * sizeof(double),
                              cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    if (n_passes) *n_passes = done;
  }
  // timing (only passes that did work)
  float ms = 0.f;
  TB_CUDA(cudaEventElapsedTime(&ms, ev[2 * npass_max], ev[2 * npass_max + 1]));
  s->loop_ms = ms;
  s->rowsum_ms = 0;
  s->rowsum_launches = 0;
  const int real = hs.all_bad ? 0 : hs.passes_done;
  for (int p = 0; p < real && p < launched; ++p) {
    TB_CUDA(cudaEventElapsedTime(&ms, ev[2 * p], ev[2 * p + 1]));
    s->rowsum_ms += ms;
    ++s->rowsum_launches;
  }
  {   // kernels of this library launched by the call (NCCL's own kernels are not counted)
    const int k1 = csr_walk ? 1 : (s->t_nunits > 0) + (s->n_groups > 0) + (s->n_sparse > 0);
    const int k2 = 2 + (s->world > 1 && s->nrows ? 1 : 0);       // (combine) + reduce + update
    s->total_launches = 4 /* init, live, bad partners, finalize */ + launched * (k1 + k2);
  }
  if (!evx.empty()) {
    double t_comb = 0, t_red = 0, t_upd = 0;
    int cnt = 0;
    for (int p = 1; p < real && p < launched; ++p) {       // pass 0 carries 2N in the all-reduce
      float a = 0.f, b = 0.f, c = 0.f;
      cudaEventElapsedTime(&a, ev[2 * p + 1], evx[3 * p]);
      cudaEventElapsedTime(&b, evx[3 * p], evx[3 * p + 1]);
      cudaEventElapsedTime(&c, evx[3 * p + 1], evx[3 * p + 2]);
      t_comb += a; t_red += b; t_upd += c;
      ++cnt;
    }
    if (cnt)
      fprintf(stderr, "[tadbit_b200] rank %d: per pass row sum %.1f us, combine %.1f us, all-reduce %.1f us, "
                      "statistics + update %.1f us (%d passes)\n", s->rank,
              1e3 * s->rowsum_ms / (s->rowsum_launches ? s->rowsum_launches : 1), 1e3 * t_comb / cnt,
              1e3 * t_red / cnt, 1e3 * t_upd / cnt, cnt);
    for (auto &e : evx) cudaEventDestroy(e);
  }
  for (auto &e : ev) cudaEventDestroy(e);
  return rc;
}

void norm_sum(tb_store *s, const double *h_bias, const uint8_t *h_bad, double *out) {
  cudaStream_t st = s->stream;
  const int64_t N = s->nbins;
  TB_REQUIRE(out, "norm_sum: null output");
  upload_bad(s, h_bad);
  if (!h_bias) {        // raw sum: exact integers
    launch_chunks<M_I64_MASKED>(s, nullptr, s->bad.p, nullptr);
    if (s->nrows)
      k_combine_i64<<<grid_1d(s->nrows, 256, s->sm_count), 256, 0, st>>>(
          s->partial_i.p, s->row_chunk.p, s->nrows, s->row_begin, s->i_full.p);
    k_sum_good_rows<<<1, 1024, 0, st>>>(s->i_full.p, s->bad.p, s->row_begin, s->row_end,
                                        s->scal_i.p);
    if (s->world > 1) allreduce_i64(s, s->scal_i.p, s->scal_i.p + 8, 1);
    long long v = 0;
    TB_CUDA(cudaMemcpyAsync(&v, s->world > 1 ? s->scal_i.p + 8 : s->scal_i.p, sizeof(long long),
                            cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    *out = (double)v;
    return;
  }
  // y = 1/bias on good bins, 0 on bad ones; sum = y . (W y)
  TB_CUDA(cudaMemcpyAsync(s->b.p, h_bias, N * sizeof(double), cudaMemcpyHostToDevice, st));
  k_inverse_bias<<<grid_1d(N, 256, s->sm_count), 256, 0, st>>>(s->b.p, s->bad.p, N, s->x.p);
  if (use_csr_rowsum(s)) {
    launch_chunks<M_F64>(s, s->x.p, nullptr, nullptr);
  } else {
    // chunks whose cells live in the sparse tiles are not written by the compact kernels: make
    // sure no partial of an earlier CSR walk (pass 0, or the "csr" option) is left in them
    if (s->t_nunits) TB_CUDA(cudaMemsetAsync(s->partial_f.p, 0, s->partial_f.bytes(), st));
    launch_rowsum(s, s->x.p, nullptr);
  }
  if (s->nrows)
    k_combine_f64<<<grid_1d(s->nrows, 256, s->sm_count), 256, 0, st>>>(
        s->partial_f.p, s->row_chunk.p, s->nrows, s->row_begin, s->s_full.p,
        nullptr, TileSums{(!use_csr_rowsum(s) && s->t_nunits) ? s->t_partial.p : nullptr,
                          s->t_unit_of_rb.p});
  k_dot_rows<<<1, 1024, 0, st>>>(s->x.p, s->s_full.p, s->row_begin, s->row_end, s->scal_f.p);
  if (s->world > 1) allreduce_f64(s, s->scal_f.p, s->scal_f.p + 8, 1);
  TB_CUDA(cudaMemcpyAsync(out, s->world > 1 ? s->scal_f.p + 8 : s->scal_f.p, sizeof(double),
                          cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
}

void diag_sums(tb_store *s, const uint8_t *h_bad, int64_t ndiag, int64_t *h_sum_d) {
  cudaStream_t st = s->stream;
  TB_REQUIRE(ndiag >= 0 && h_sum_d, "diag_sums: bad arguments");
  if (ndiag == 0) return;
  upload_bad(s, h_bad);
  DevBuf<unsigned long long> hist, red;
  hist.alloc(ndiag);
  TB_CUDA(cudaMemsetAsync(hist.p, 0, hist.bytes(), st));
  if (s->nchunks) {
    int64_t need = (s->nchunks + 7) / 8, cap = (int64_t)s->sm_count * 4;
    int g = (int)(need < cap ? need : cap);
    k_diag_hist<<<g, 256, 0, st>>>(s->col.p, s->val.p, s->chunk_beg.p, s->chunk_row.p, s->rowptr.p,
                                   s->nchunks, s->row_begin, s->bad.p, s->chrom_of.p, ndiag, hist.p);
  }
  unsigned long long *src = hist.p;
  if (s->world > 1) {
    red.alloc(ndiag);
    allreduce_i64(s, (const long long *)hist.p, (long long *)red.p, ndiag);
    src = red.p;
  }
  TB_CUDA(cudaMemcpyAsync(h_sum_d, src, ndiag * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
}

void ice_batch(tb_store *s, const uint8_t *h_bad, int32_t iterations, double max_dev, double *h_bias,
               double *h_trace, int32_t *n_passes) {
  cudaStream_t st = s->stream;
  const int64_t N = s->nbins;
  TB_REQUIRE(iterations >= 0, "ice_batch: iterations must be >= 0");
  TB_REQUIRE(s->world == 1, "ice_batch: batches are not sharded (one replica per GPU)");
  const int nseg = (int)s->chrom_offsets.size() - 1;
  const int npass_max = iterations + 1;
  DevBuf<SegState> seg;
  DevBuf<unsigned int> n_stopped;
  DevBuf<int64_t> seg_off;
  DevBuf<double> trace;
  seg.alloc(nseg);
  n_stopped.alloc(1);
  seg_off.alloc(nseg + 1);
  trace.alloc((size_t)nseg * npass_max * 4);
  TB_CUDA(cudaMemcpyAsync(seg_off.p, s->chrom_offsets.data(), (nseg + 1) * sizeof(int64_t),
                          cudaMemcpyHostToDevice, st));
  TB_CUDA(cudaMemsetAsync(trace.p, 0, trace.bytes(), st));
  upload_bad(s, h_bad);
  const int vg = grid_1d(N, 256, s->sm_count);
  k_ice_init<<<vg, 256, 0, st>>>(s->bad.p, N, s->x.p, s->b.p, s->state.p);
  k_seg_init<<<(nseg + 255) / 256, 256, 0, st>>>(seg.p, nseg, n_stopped.p);
  const bool csr_walk = use_csr_rowsum(s);
  const TileSums ts = {(!csr_walk && s->t_nunits) ? s->t_partial.p : nullptr, s->t_unit_of_rb.p};
  prepare_live(s, s->s_full.p + N);
  if (ts.partial) TB_CUDA(cudaMemsetAsync(s->partial_f.p, 0, s->partial_f.bytes(), st));
  cudaEvent_t e0, e1;
  TB_CUDA(cudaEventCreate(&e0));
  TB_CUDA(cudaEventCreate(&e1));
  TB_CUDA(cudaEventRecord(e0, st));
  bool stopped = false;
  int launched = 0;
  for (int p = 0; p < npass_max && !stopped; ++p) {
    if (csr_walk) launch_chunks<M_F64>(s, s->x.p, nullptr, s->state.p);
    else launch_rowsum(s, s->x.p, s->state.p);
    k_seg_reduce<<<nseg, 1024, 0, st>>>(s->s_full.p, s->x.p, s->bad.p, s->present.p, seg_off.p, nseg,
                                        N, p, iterations, max_dev, trace.p, npass_max, seg.p,
                                        n_stopped.p, s->state.p, s->partial_f.p, s->row_chunk.p, ts);
    k_seg_update<<<vg, 256, 0, st>>>(s->s_full.p, s->x.p, s->b.p, s->present.p, s->chrom_of.p, N, p,
                                     seg.p);
    ++launched;
    if ((p + 1) % 4 == 0 || p == npass_max - 1 || iterations == 0) {
      TB_CUDA(cudaMemcpyAsync(s->h_state.p, s->state.p, sizeof(IceState), cudaMemcpyDeviceToHost, st));
      TB_CUDA(cudaStreamSynchronize(st));
      stopped = s->h_state.p->stopped != 0;
    }
  }
  TB_CUDA(cudaEventRecord(e1, st));
  k_seg_finalize<<<vg, 256, 0, st>>>(s->b.p, s->present.p, s->chrom_of.p, N, seg.p, s->x.p);
  std::vector<SegState> hseg(nseg);
  TB_CUDA(cudaMemcpyAsync(h_bias, s->x.p, N * sizeof(double), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(hseg.data(), seg.p, nseg * sizeof(SegState), cudaMemcpyDeviceToHost, st));
  if (h_trace && iterations > 0)
    TB_CUDA(cudaMemcpyAsync(h_trace, trace.p, trace.bytes(), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < nseg; ++k) n_passes[k] = hseg[k].all_bad ? -1 : hseg[k].passes_done;
  float ms = 0.f;
  TB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  s->loop_ms = ms;
  s->rowsum_ms = 0;
  s->rowsum_launches = 0;
  {
    const int k1 = csr_walk ? 1 : (s->t_nunits > 0) + (s->n_groups > 0) + (s->n_sparse > 0);
    s->total_launches = 3 + (launched ? 3 : 0) + (launched > 1 ? (launched - 1) * (k1 + 2) : 0);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

// Work of every owned row in the two K1 kernels: slots of its dense chunks, cells in sparse tiles.
namespace {
__global__ void k_row_work(const int4 *__restrict__ desc, const uint16_t *__restrict__ tovf,
                           const int64_t *__restrict__ row_chunk,
                           int64_t nrows, long long *__restrict__ dense_slots,
                           long long *__restrict__ tile_cells) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x) {
    long long d = 0, t = 0;
    for (int64_t c = row_chunk[r]; c < row_chunk[r + 1]; ++c) {
      const int4 e = desc[c];
      const int enc = e.z & 0xff;
      if (enc >= ENC_D32 && enc <= ENC_D8) d += e.y * (enc == ENC_D32 ? 4 : (enc == ENC_D16 ? 2 : 1));
      else t += e.y;
      t += tovf[c];
    }
    dense_slots[r] = d;      // weighted by bytes per slot
    tile_cells[r] = t;
  }
}
}  // namespace

void row_work(tb_store *s, int64_t *h_dense, int64_t *h_tile) {
  cudaStream_t st = s->stream;
  if (!s->nrows) return;
  DevBuf<long long> d, t;
  d.alloc(s->nrows);
  t.alloc(s->nrows);
  k_row_work<<<grid_1d(s->nrows, 256, s->sm_count), 256, 0, st>>>(s->enc_desc.p, s->enc_tovf.p, s->row_chunk.p,
                                                                  s->nrows, d.p, t.p);
  TB_CUDA(cudaMemcpyAsync(h_dense, d.p, s->nrows * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(h_tile, t.p, s->nrows * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
}

// Mean device time of one K1 evaluation (x = 1) of this store alone -- total, sparse-tile kernel,
// dense (+ per-cell sparse) kernels -- for balancing shards before they are connected.
void probe_rowsum(tb_store *s, int32_t reps, double ms[3]) {
  cudaStream_t st = s->stream;
  TB_REQUIRE(reps > 0 && ms, "probe_rowsum: bad arguments");
  TB_CUDA(cudaMemsetAsync(s->bad.p, 0, s->nbins, st));
  k_ice_init<<<grid_1d(s->nbins, 256, s->sm_count), 256, 0, st>>>(s->bad.p, s->nbins, s->x.p, s->b.p,
                                                                  s->state.p);
  launch_rowsum(s, s->x.p, nullptr);                 // warm-up
  cudaEvent_t e[3];
  for (auto &x : e) TB_CUDA(cudaEventCreate(&x));
  float tiles = 0.f, rest = 0.f, f = 0.f;
  for (int r = 0; r < reps; ++r) {
    TB_CUDA(cudaEventRecord(e[0], st));
    launch_rowsum_tiles(s, s->x.p, nullptr);
    TB_CUDA(cudaEventRecord(e[1], st));
    launch_rowsum_chunks(s, s->x.p, nullptr);
    TB_CUDA(cudaEventRecord(e[2], st));
    TB_CUDA(cudaStreamSynchronize(st));
    TB_CUDA(cudaEventElapsedTime(&f, e[0], e[1]));
    tiles += f;
    TB_CUDA(cudaEventElapsedTime(&f, e[1], e[2]));
    rest += f;
  }
  for (auto &x : e) cudaEventDestroy(x);
  ms[1] = tiles / reps;
  ms[2] = rest / reps;
  ms[0] = ms[1] + ms[2];
}

}  // namespace tb
