// Hot-path kernels (sm_100a) around K1 (rowsum.cu, tiles.cu):
//
//   K2  k_ice_exchange           own row sums -> every GPU's buffer over NVLink, flag barrier,
//                                S = x.s, meanS/min/max/dev, B *= S/meanS, x = 1/B in ONE kernel
//                                (normalize_hic.py:146-151,217-222); k_ice_reduce/_update + NCCL
//                                all-reduce = the same pass as separate steps (TB_EXCHANGE=nccl)
//   K3  K1 + k_dot               sum v/(b_i b_j)                 (hic_data.py:350-375)
//   K4  K1 with x = good mask    masked column sums, exact       (hic_filtering.py:36-39,141-145);
//       row_nnz / row_sum        kept from the build             (hic_filtering.py:185-200)
//   K5  DiagVisitor              per-distance sums               (normalize_hic.py:271-283)
//
// Stores that hold zeros or negative values keep their CSR; the k_chunk_reduce kernels walk it
// (also the "rowsum=csr" test switch: a second, independent K1).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "visit.cuh"

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
// (loads issued four at a time, additions in unit order: the sum does not depend on the unrolling)
__device__ __forceinline__ double tile_sum(const TileSums &ts, int64_t r) {
  if (!ts.partial) return 0.0;
  const int64_t rb = r / kTileRows;
  const double *p = ts.partial + (r % kTileRows);
  int32_t u = ts.unit_of_rb[rb];
  const int32_t u1 = ts.unit_of_rb[rb + 1];
  double a = 0.0;
  for (; u + 4 <= u1; u += 4) {
    const double v0 = p[(int64_t)u * kTileRows], v1 = p[(int64_t)(u + 1) * kTileRows];
    const double v2 = p[(int64_t)(u + 2) * kTileRows], v3 = p[(int64_t)(u + 3) * kTileRows];
    a += v0; a += v1; a += v2; a += v3;
  }
  for (; u < u1; ++u) a += p[(int64_t)u * kTileRows];
  return a;
}

// row sum of owned row r = its chunk partials in chunk order + its sparse-tile sums.  The two
// chains of dependent loads (row -> chunks -> partials, row block -> units -> partials) are started
// together: two round trips to L2 instead of four or five.
__device__ __forceinline__ double row_total(const double *__restrict__ pf, const int64_t *__restrict__ row_chunk,
                                            const TileSums &ts, int64_t r) {
  int64_t c = row_chunk[r];
  const int64_t c1 = row_chunk[r + 1];
  TB_DCHECK(c >= 0 && c <= c1 && c1 <= c_lim.nchunks && r < c_lim.nrows);
  int32_t u = 0, u1 = 0;
  const double *p = nullptr;
  if (ts.partial) {
    const int64_t rb = r / kTileRows;
    u = ts.unit_of_rb[rb];
    u1 = ts.unit_of_rb[rb + 1];
    p = ts.partial + (r % kTileRows);
    TB_DCHECK(u >= 0 && u <= u1 && u1 <= c_lim.t_nunits);
  }
  // first batch of both lists at once
  double v[4], w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = c + k < c1 ? pf[c + k] : 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = u + k < u1 ? p[(int64_t)(u + k) * kTileRows] : 0.0;
  double a = 0.0, t = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) if (c + k < c1) a += v[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) if (u + k < u1) t += w[k];
  for (c += 4; c + 4 <= c1; c += 4) {
    const double v0 = pf[c], v1 = pf[c + 1], v2 = pf[c + 2], v3 = pf[c + 3];
    a += v0; a += v1; a += v2; a += v3;
  }
  for (; c < c1; ++c) a += pf[c];
  for (u += 4; u + 4 <= u1; u += 4) {
    const double v0 = p[(int64_t)u * kTileRows], v1 = p[(int64_t)(u + 1) * kTileRows];
    const double v2 = p[(int64_t)(u + 2) * kTileRows], v3 = p[(int64_t)(u + 3) * kTileRows];
    t += v0; t += v1; t += v2; t += v3;
  }
  for (; u < u1; ++u) t += p[(int64_t)u * kTileRows];
  return a + t;
}

// Rows kept by copy_matrix (normalize_hic.py:165-177): not bad, and at least one stored cell whose
// partner is not bad.  The OWNER of a row decides, in pass 0: a store of positive counts has no
// CSR, the row keeps a cell iff its first row sum (x = the good mask) is > 0; a store that kept
// its CSR (zeros, negative values) counts the cells with a good partner (`live`, per owned row).
// The decision travels with the row sum of pass 0: a row that is not present is published as
// NaN (its sum is never used), so no second exchange is needed and ranks need not agree on
// which rule they apply.
__device__ __forceinline__ double mark_present(const uint8_t *bad, const double *live, int64_t g, int64_t r,
                                               double a) {
  const bool p = !bad[g] && (live ? live[r] > 0.0 : a > 0.0);
  return p ? a : __longlong_as_double(0x7ff8000000000000ll);
}

// out[rb + r] = sum of the row's chunk partials in chunk order + its sparse-tile sums.
// mark: pass 0 of an ICE call on several GPUs -- rows that are not present are written as NaN.
__global__ void k_combine_f64(const double *__restrict__ pf, const int64_t *__restrict__ row_chunk,
                              int64_t nrows, int64_t rb, double *__restrict__ out,
                              const IceState *__restrict__ state, TileSums ts, bool mark,
                              const uint8_t *__restrict__ bad, const double *__restrict__ live) {
  if (state && state->stopped) return;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x) {
    double a = row_total(pf, row_chunk, ts, r);
    out[rb + r] = mark ? mark_present(bad, live, rb + r, r, a) : a;
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

// out[rb + r] = v[r]: the owned rows' entries of a full-length vector (zeros elsewhere)
__global__ void k_place_i64(const long long *__restrict__ v, int64_t nrows, int64_t rb,
                            long long *__restrict__ out) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x)
    out[rb + r] = v[r];
}

// exact integer row sums computed in fp64 (every partial sum is an integer < 2^53) -> int64
__global__ void k_f64_to_i64(const double *__restrict__ v, int64_t rb, int64_t re,
                             long long *__restrict__ out) {
  for (int64_t i = rb + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < re;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (long long)v[i];
}

__global__ void k_good_mask(const uint8_t *__restrict__ bad, int64_t n, double *__restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    x[i] = bad[i] ? 0.0 : 1.0;
}

// ---------------------------------------------------------------------------------
// ICE vector kernels
// ---------------------------------------------------------------------------------
// "Present" rows (copy_matrix, normalize_hic.py:165-177) of a store that kept its CSR:
// live[r] = stored cells of owned row r whose partner is not bad.  Small integers carried as
// doubles: the order of the additions is irrelevant.
__global__ void __launch_bounds__(256)
k_live_rows(const int32_t *__restrict__ col, const int64_t *__restrict__ chunk_beg,
            const int32_t *__restrict__ chunk_row, int64_t nchunks,
            const uint8_t *__restrict__ bad, double *__restrict__ live) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < nchunks; c += nwarp) {
    int cnt = 0;
    for (int64_t k = chunk_beg[c] + lane; k < chunk_beg[c + 1]; k += 32) cnt += !bad[col[k]];
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0 && cnt) atomicAdd(&live[chunk_row[c]], (double)cnt);
  }
}

__global__ void k_ice_init(const uint8_t *__restrict__ bad, int64_t n, double *x, double *b,
                           IceState *state, unsigned long long epoch_base) {
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
    state->zero_div = 0;
    state->fault = 0;
    state->arrive1 = state->arrive2 = 0;
    state->fold_epoch = epoch_base;
    state->epoch_base = epoch_base;
  }
}

constexpr int kRedThreads = 256;

// The reference's loop decision for one pass, taken by ONE thread once sum / min / max of S over
// the present rows are known (normalize_hic.py:211-222).
__device__ void ice_decide(double sum, double mn, double mx, double dcnt, int pass, int iterations,
                           double max_dev, double *trace, IceState *state) {
  if (pass == 0) state->n_present = (long long)dcnt;
  const long long np = state->n_present;
  if (np == 0) {
    state->all_bad = 1;
    state->stopped = 1;
    return;
  }
  const double mean = sum / (double)np;
  if (mean == 0.0) {                // float(S) / meanS raises in the reference (_updateDB)
    state->zero_div = 1;
    state->stopped = 1;
    return;
  }
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

// per-thread sum / min / max / count -> per-block entry of block_red (thread 0 writes it)
__device__ __forceinline__ void block_stats(double &sum, double &mn, double &mx, long long &cnt,
                                            double *sh, int nwarps) {
  for (int o = 16; o; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[4 * w] = sum; sh[4 * w + 1] = mn; sh[4 * w + 2] = mx; sh[4 * w + 3] = (double)cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double c = (double)cnt;
    for (int k = 1; k < nwarps; ++k) {
      sum += sh[4 * k]; mn = fmin(mn, sh[4 * k + 1]); mx = fmax(mx, sh[4 * k + 2]); c += sh[4 * k + 3];
    }
    sh[3] = c;
  }
}

// one warp folds the per-block entries, lane l takes blocks l, l+32, ... and the lanes are then
// combined by a fixed butterfly -> same bits every run and on every GPU
__device__ __forceinline__ void fold_blocks(const double *block_red, unsigned int nblocks, double &sum,
                                            double &mn, double &mx, double &dcnt) {
  sum = 0.0; mn = INFINITY; mx = -INFINITY; dcnt = 0.0;
  // the entries of 8 blocks per lane are requested together (independent loads, one L2 round trip),
  // then added in block order
  for (unsigned int b0 = threadIdx.x; b0 < nblocks; b0 += 32 * 8) {
    double v[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const unsigned int bl = b0 + 32 * k;
      const double2 *br = reinterpret_cast<const double2 *>(block_red + 4 * (bl < nblocks ? bl : b0));
      const double2 a = __ldcg(br), c = __ldcg(br + 1);     // L2: written by other CTAs of this launch
      v[k][0] = a.x; v[k][1] = a.y; v[k][2] = c.x; v[k][3] = c.y;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (b0 + 32 * k < nblocks) {
        sum += v[k][0]; mn = fmin(mn, v[k][1]); mx = fmax(mx, v[k][2]); dcnt += v[k][3];
      }
  }
  for (int o = 16; o; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    dcnt += __shfl_xor_sync(0xffffffffu, dcnt, o);
  }
}

// S_i = x_i * s_i over present rows -> sum / min / max; the last block to arrive folds the
// per-block partials in block order and takes the reference's loop decision.
// FUSED (single GPU): the row's chunk partials are added here instead of in k_combine_f64
// (one kernel and one pass over s less); s_full is still written for k_ice_update.
template <bool FUSED>
__global__ void __launch_bounds__(kRedThreads)
k_ice_reduce(double *__restrict__ s_full, const double *__restrict__ x,
             const uint8_t *__restrict__ bad, uint8_t *__restrict__ present,
             const double *__restrict__ live, int64_t n,
             int pass, int iterations, double max_dev, double *__restrict__ block_red,
             double *__restrict__ trace, IceState *state, const double *__restrict__ pf,
             const int64_t *__restrict__ row_chunk, TileSums ts) {
  if (state->stopped) return;
  __shared__ double sh[4 * (kRedThreads / 32)];
  __shared__ bool is_last;
  double sum = 0.0, mn = INFINITY, mx = -INFINITY;
  long long cnt = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double si;
    if (FUSED) {                       // one GPU: row i is local row i
      double a = row_total(pf, row_chunk, ts, i);
      if (pass == 0) a = mark_present(bad, live, i, i, a);
      s_full[i] = si = a;
    } else {
      si = s_full[i];
    }
    bool p;
    if (pass == 0) present[i] = p = !isnan(si);
    else p = present[i];
    if (p) {
      double S = x[i] * si;
      sum += S;
      mn = fmin(mn, S);
      mx = fmax(mx, S);
      ++cnt;
    }
  }
  block_stats(sum, mn, mx, cnt, sh, kRedThreads / 32);
  if (threadIdx.x == 0) {
    block_red[4 * blockIdx.x + 0] = sum;
    block_red[4 * blockIdx.x + 1] = mn;
    block_red[4 * blockIdx.x + 2] = mx;
    block_red[4 * blockIdx.x + 3] = sh[3];
    __threadfence();
    unsigned int t = atomicAdd(&state->blocks_arrived, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last || threadIdx.x >= 32) return;
  __threadfence();
  double dcnt;
  fold_blocks(block_red, gridDim.x, sum, mn, mx, dcnt);
  if (threadIdx.x != 0) return;
  state->blocks_arrived = 0;
  ice_decide(sum, mn, mx, dcnt, pass, iterations, max_dev, trace, state);
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

// ---------------------------------------------------------------------------------
// K2 fused with the exchange between GPUs.  One launch per pass does everything that follows
// the row sums:
//   1. every rank adds up the partials of the rows it owns and stores the sums straight into
//      the buffer of EVERY rank (peer-mapped memory, NVLink stores; buffers alternate by pass so
//      a fast rank never overwrites what a slow one still reads);
//   2. flag barrier: when all CTAs of a rank have stored, one thread publishes the exchange
//      number in every peer's flag array (st.release.sys), and every CTA waits until all peers'
//      numbers have arrived in its own array (ld.acquire.sys);
//   3. statistics over the full vector (each rank computes them redundantly from the same bits
//      in the same order, so all ranks take the same stop decision without a second exchange),
//      grid-wide fold by the last CTA, loop decision;
//   4. bias update B *= S / meanS, x = 1 / B.
// The pass number is read from the device state, so the launch arguments never change.
// The grid must be co-resident: CTAs wait for each other.  It is sized from the occupancy the
// runtime reports (exchange_grid), two CTAs of 512 threads per SM; the row phase is a chain of
// dependent loads per row, so every thread works on two rows at a time to keep loads in flight.
// ---------------------------------------------------------------------------------
constexpr int kXThreads = 512;
constexpr unsigned long long kPeerTimeoutNs = 4000000000ull;    // 4 s: a peer that never arrives

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// LL exchange: a row sum travels as two 8-byte words, each = (pass tag << 32) | half of the double.
// An aligned 8-byte store is single-copy atomic, so a reader that finds the tag of the current
// pass in both words has the value: no fence, no flag, no barrier between the GPUs.
__device__ __forceinline__ void st_ll(unsigned long long *p, double v, unsigned long long tag) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long w0 = tag | (bits & 0xffffffffull), w1 = tag | (bits >> 32);
  asm volatile("st.relaxed.sys.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool ld_ll(const unsigned long long *p, unsigned long long tag, double &v) {
  unsigned long long w0, w1;
  asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
  v = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
  return (w0 & 0xffffffff00000000ull) == tag && (w1 & 0xffffffff00000000ull) == tag;
}

__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <bool LL>
__global__ void __launch_bounds__(kXThreads, 2)
k_ice_exchange(PeerView pv, const double *__restrict__ pf, const int64_t *__restrict__ row_chunk,
               TileSums ts, int64_t nrows, int64_t rb, int64_t n, double *__restrict__ x,
               double *__restrict__ b, const uint8_t *__restrict__ bad, uint8_t *__restrict__ present,
               const double *__restrict__ live, int iterations, double max_dev,
               double *__restrict__ block_red, double *__restrict__ trace, IceState *state,
               unsigned long long *__restrict__ prof) {
  volatile IceState *vs = state;
  if (vs->stopped) return;
  // TB_TIMING=1: CTA 0 adds the time it spends in each phase (ns, %globaltimer) to prof[0..5]
  unsigned long long tprev = 0;
  auto lap = [&](int k) {
    if (prof && blockIdx.x == 0 && threadIdx.x == 0) {
      const unsigned long long t = now_ns();
      if (k >= 0) prof[k] += t - tprev;
      tprev = t;
    }
  };
  lap(-1);
  __shared__ double sh[4 * (kXThreads / 32)];
  __shared__ int s_last, s_fault;
  const int pass = vs->passes_done;
  const unsigned long long epoch = vs->epoch_base + (unsigned long long)pass + 1ull;
  const int buf = pass & 1;
  const unsigned long long tag = (epoch & 0xffffffffull) << 32;
  if (threadIdx.x == 0) s_fault = 0;
  // 1. the rows of this rank -> every rank (two rows per thread and step)
  {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows; r += 2 * stride) {
      const int64_t r2 = r + stride;
      const bool two = r2 < nrows;
      double a = row_total(pf, row_chunk, ts, r);
      double a2 = two ? row_total(pf, row_chunk, ts, r2) : 0.0;
      if (pass == 0) {
        a = mark_present(bad, live, rb + r, r, a);
        if (two) a2 = mark_present(bad, live, rb + r2, r2, a2);
      }
#pragma unroll
      for (int q = 0; q < kMaxPeers; ++q)          // constant indices: the parameter stays in the constant bank
        if (q < pv.world) {
          double *dst = buf ? pv.s[q][1] : pv.s[q][0];
          if (LL) {
            unsigned long long *d = reinterpret_cast<unsigned long long *>(dst);
            st_ll(d + 2 * (rb + r), a, tag);
            if (two) st_ll(d + 2 * (rb + r2), a2, tag);
          } else {
            dst[rb + r] = a;
            if (two) dst[rb + r2] = a2;
          }
        }
    }
  }
  lap(0);
  // CTA barrier, then ONE fence per CTA (as a cooperative-groups grid sync does): the barrier orders the
  // stores of the CTA's threads before thread 0's fence, whose cumulativity carries them to system
  // scope before the arrival is counted and the flag published
  __syncthreads();
  lap(1);
  // 2. (flag protocol only) publish once every CTA has stored, then wait for every peer
  if (!LL && threadIdx.x == 0) {
    if (pv.world > 1) __threadfence_system();
    else __threadfence();
    const unsigned int t = atomicAdd(&state->arrive1, 1u);
    if (t == gridDim.x - 1) {
      state->arrive1 = 0;
      if (pv.world > 1) __threadfence_system();
      else __threadfence();
#pragma unroll
      for (int q = 0; q < kMaxPeers; ++q)
        if (q < pv.world) st_release_sys(pv.flag[q] + pv.rank, epoch);
    }
  }
  if (!LL && threadIdx.x < pv.world) {
    const unsigned long long t0 = now_ns();
    while (ld_acquire_sys(pv.my_flags + threadIdx.x) < epoch) {
      if (now_ns() - t0 > kPeerTimeoutNs) { s_fault = 1; break; }
    }
  }
  if (!LL) {
    __syncthreads();
    if (s_fault) {                     // give up: the host reports it, nothing is updated
      if (threadIdx.x == 0) { vs->fault = 1; vs->stopped = 1; }
      return;
    }
  }
  lap(2);
  // 3. statistics of the pass (LL: every thread waits for the words of its own entries)
  const double *sfull = buf ? pv.mine[1] : pv.mine[0];
  const unsigned long long *sll = reinterpret_cast<const unsigned long long *>(sfull);
  double sum = 0.0, mn = INFINITY, mx = -INFINITY;
  long long cnt = 0;
  bool late = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double si;
    if (LL) {
      if (!ld_ll(sll + 2 * i, tag, si)) {
        const unsigned long long t0 = now_ns();
        unsigned int spins = 0;
        while (!ld_ll(sll + 2 * i, tag, si))
          if ((++spins & 1023u) == 0u && now_ns() - t0 > kPeerTimeoutNs) { late = true; break; }
      }
      if (late) break;
    } else {
      si = __ldcg(sfull + i);
    }
    bool p;
    if (pass == 0) present[i] = p = !isnan(si);
    else p = present[i];
    if (p) {
      const double S = x[i] * si;
      sum += S;
      mn = fmin(mn, S);
      mx = fmax(mx, S);
      ++cnt;
    }
  }
  if (LL) {
    if (late) s_fault = 1;
    __syncthreads();
    if (s_fault) {                     // a peer never delivered: the host reports it, nothing is updated
      if (threadIdx.x == 0) { vs->fault = 1; vs->stopped = 1; }
      return;
    }
  }
  block_stats(sum, mn, mx, cnt, sh, kXThreads / 32);
  if (threadIdx.x == 0) {
    block_red[4 * blockIdx.x + 0] = sum;
    block_red[4 * blockIdx.x + 1] = mn;
    block_red[4 * blockIdx.x + 2] = mx;
    block_red[4 * blockIdx.x + 3] = sh[3];
    __threadfence();
    const unsigned int t = atomicAdd(&state->arrive2, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence();
    double dcnt;
    fold_blocks(block_red, gridDim.x, sum, mn, mx, dcnt);
    if (threadIdx.x == 0) {
      state->arrive2 = 0;
      ice_decide(sum, mn, mx, dcnt, pass, iterations, max_dev, trace, state);
      __threadfence();
      st_release_gpu(&state->fold_epoch, epoch);
    }
  }
  lap(3);
  // 4. every CTA waits for the decision, then updates its part of B and x
  if (threadIdx.x == 0) {
    const unsigned long long t0 = now_ns();
    while (ld_acquire_gpu(&state->fold_epoch) < epoch) {
      if (now_ns() - t0 > kPeerTimeoutNs) { s_fault = 1; break; }
    }
  }
  __syncthreads();
  if (s_fault) {
    if (threadIdx.x == 0) { vs->fault = 1; vs->stopped = 1; }
    return;
  }
  lap(4);
  if (vs->passes_done != pass + 1 || vs->all_bad || vs->zero_div) return;
  const double mean = vs->mean_s;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (!present[i]) {
      x[i] = 0.0;
      continue;
    }
    double si;
    if (LL) ld_ll(sll + 2 * i, tag, si);
    else si = __ldcg(sfull + i);
    const double S = x[i] * si;
    const double bi = b[i] * (S / mean);
    b[i] = bi;
    x[i] = (bi != 0.0) ? 1.0 / bi : 0.0;     // B == 0: row frozen by the ZeroDivisionError skip
  }
  lap(5);
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
                               int64_t n, double *__restrict__ y, int *__restrict__ zero_bias) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (!bad[i] && bias[i] == 0.0) *zero_bias = 1;      // the reference divides by it and raises
    y[i] = bad[i] ? 0.0 : 1.0 / bias[i];
  }
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
// K5: per-distance sums of the cells (i, i + d) inside one section whose ROW is not bad
// (normalize_hic.py:271-283: only i is tested).  Visitor over the stored upper-triangle cells:
// shared-memory histogram for d < 4096, L2 atomics beyond.  Integer sums: any order is exact.
// ---------------------------------------------------------------------------------
#ifndef TB_DIAG_SHARED
#define TB_DIAG_SHARED 24576      // distances kept in shared memory (8 B each: 192 KB, one CTA per SM)
#endif
constexpr int kDiagShared = TB_DIAG_SHARED;

struct DiagVisitor {
  static constexpr int kShared = kDiagShared * 8;
  static constexpr int kThreads = kDiagShared > 4096 ? 1024 : 256;
  const uint8_t *bad;
  const int32_t *chrom_of;
  const int64_t *chrom_end;       // [sections] one past the last bin of each section
  int64_t ndiag;
  unsigned long long *hist;
  // shared histogram as two 32-bit halves per distance: a 64-bit atomic add on shared memory is a
  // compare-and-swap loop (ncu, first version: 6 wavefronts per add), a 32-bit one is native; the
  // carry out of the low half (rare) and the sign extension of a negative count go to the high half
  unsigned int *lo, *hi;
  int64_t jend;                   // per thread: cells of the current row count up to this column
  __device__ void begin(uint8_t *smem) {
    lo = reinterpret_cast<unsigned int *>(smem);
    hi = lo + kDiagShared;
    for (int k = threadIdx.x; k < 2 * kDiagShared; k += blockDim.x) lo[k] = 0u;
  }
  // rows [i0, i1) reach at most to the end of the section of row i1 - 1
  __device__ bool block(int64_t i0, int64_t i1, int64_t j0, int64_t j1) {
    return j0 < chrom_end[chrom_of[i1 - 1 < ndiag_rows ? i1 - 1 : ndiag_rows - 1]];
  }
  __device__ bool row(int64_t i) {
    if (bad[i]) return false;
    jend = chrom_end[chrom_of[i]];
    if (jend > i + ndiag) jend = i + ndiag;
    return true;
  }
  __device__ void cell(int64_t i, int64_t j, int32_t v) {
    if (j >= jend) return;
    const int64_t d = j - i;
    if (d < kDiagShared) {
      const unsigned int add = (unsigned int)v;
      const unsigned int old = atomicAdd(&lo[d], add);
      const unsigned int up = (v < 0 ? 0xffffffffu : 0u) + ((unsigned int)(old + add) < old ? 1u : 0u);
      if (up) atomicAdd(&hi[d], up);
    } else {
      atomicAdd(&hist[d], (unsigned long long)(long long)v);
    }
  }
  __device__ void end(uint8_t *smem) {
    for (int k = threadIdx.x; k < kDiagShared && k < ndiag; k += blockDim.x) {
      const unsigned long long t = ((unsigned long long)hi[k] << 32) + lo[k];
      if (t) atomicAdd(&hist[k], t);
    }
  }
  int64_t ndiag_rows;             // number of bins (rows of a partial last tile lie beyond it)
};

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
             const int64_t *__restrict__ row_chunk, TileSums ts, const double *__restrict__ live) {
  const int k = blockIdx.x;
  if (seg[k].stopped) return;
  __shared__ double sh_sum[32], sh_min[32], sh_max[32];
  __shared__ long long sh_cnt[32];
  double sum = 0.0, mn = INFINITY, mx = -INFINITY;
  long long cnt = 0;
  for (int64_t i = seg_off[k] + threadIdx.x; i < seg_off[k + 1]; i += blockDim.x) {
    double a = row_total(pf, row_chunk, ts, i);
    s_full[i] = a;
    bool p;
    if (pass == 0) present[i] = p = !isnan(mark_present(bad, live, i, i, a));
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
  } else if (sum / (double)g.n_present == 0.0) {   // meanS == 0: ZeroDivisionError in the reference
    g.all_bad = 2;
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
// Only stores that kept their CSR can (TB_KEEP_CSR=1, or matrices with zeros / negative values).
bool use_csr_rowsum(const tb_store *s) {
  if (!s->csr_resident) return false;
  if (s->rowsum_mode >= 0) return s->rowsum_mode == 1;
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("TB_ROWSUM");
    v = (e && !strcmp(e, "csr")) ? 1 : 0;
  }
  return v == 1;
}

// TB_EXCHANGE=nccl: the pass is combine -> ncclAllReduce -> statistics -> update as separate
// steps (the parity reference of the fused exchange); also taken when the peers' buffers
// could not be mapped.
bool use_fused_exchange(const tb_store *s) {
  int want = s->exchange_mode;           // tb_store_set_option("exchange", ...)
  if (want < 0) {
    static int v = -1;
    if (v < 0) {
      const char *e = getenv("TB_EXCHANGE");
      v = (e && !strcmp(e, "nccl")) ? 0 : 1;
    }
    want = v;
  }
  return want == 1 && s->px.ready && (s->world == 1 || s->px.mapped);
}

TileSums tile_sums(const tb_store *s) {
  return TileSums{(!use_csr_rowsum(s) && s->t_nunits) ? s->t_partial.p : nullptr, s->t_unit_of_rb.p};
}

// One K1 evaluation s = W x into partial_f / t_partial (combined by the caller).
void rowsum_once(tb_store *s, const double *x, const IceState *state) {
  if (use_csr_rowsum(s)) {
    launch_chunks<M_F64>(s, x, nullptr, state);
  } else {
    launch_rowsum(s, x, state);
  }
}

// Chunks whose cells live in the sparse tiles are not written by the compact kernels: after a
// CSR walk (which writes every chunk's partial) they must be cleared again.
void clear_tile_chunk_partials(tb_store *s) {
  if (s->csr_resident && s->t_nunits)
    TB_CUDA(cudaMemsetAsync(s->partial_f.p, 0, s->partial_f.bytes(), s->stream));
}

// Integer row sums of the owned rows over the good columns, exact, into dst[row_begin + r].
void masked_row_sums(tb_store *s, long long *dst) {
  cudaStream_t st = s->stream;
  const int64_t N = s->nbins;
  if (!s->nrows) return;
  set_check_limits(s);
  const int rgrid = grid_1d(s->nrows, 256, s->sm_count);
  if (s->positive && !use_csr_rowsum(s)) {
    // K1 with x = the good mask: every term and every partial sum is an integer below 2^53
    k_good_mask<<<grid_1d(N, 256, s->sm_count), 256, 0, st>>>(s->bad.p, N, s->x.p);
    clear_tile_chunk_partials(s);
    launch_rowsum(s, s->x.p, nullptr);
    k_combine_f64<<<rgrid, 256, 0, st>>>(s->partial_f.p, s->row_chunk.p, s->nrows, s->row_begin,
                                         s->s_full.p, nullptr, tile_sums(s), false, nullptr, nullptr);
    k_f64_to_i64<<<rgrid, 256, 0, st>>>(s->s_full.p, s->row_begin, s->row_end, dst);
  } else {
    TB_REQUIRE(s->csr_resident, "internal: integer row sums need the CSR for this matrix");
    launch_chunks<M_I64_MASKED>(s, nullptr, s->bad.p, nullptr);
    k_combine_i64<<<rgrid, 256, 0, st>>>(s->partial_i.p, s->row_chunk.p, s->nrows, s->row_begin, dst);
  }
}

// TB_EXCHANGE=flags: data stores + one flag per peer behind a system-scope fence (the first fused
// protocol); default: LL words that carry their own pass tag
bool exchange_ll() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("TB_EXCHANGE");
    v = (e && !strcmp(e, "flags")) ? 0 : 1;
  }
  return v == 1;
}

// CTAs of k_ice_exchange that are resident at the same time on this device
int exchange_grid(const tb_store *s) {
  static int per_sm[64] = {0};
  int &v = per_sm[s->device & 63];
  if (!v) {
    TB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_ice_exchange<true>, kXThreads, 0));
    int v2 = v;
    TB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v2, k_ice_exchange<false>, kXThreads, 0));
    if (v2 < v) v = v2;
    if (v < 1) v = 1;
    if (v > 2) v = 2;
  }
  return v * s->sm_count;
}

cudaEvent_t ice_event(tb_store *s, size_t k) {
  while (s->events.size() <= k) {
    cudaEvent_t e;
    TB_CUDA(cudaEventCreate(&e));
    s->events.push_back(e);
  }
  return s->events[k];
}

}  // namespace

// ---------------------------------------------------------------------------------
void row_stats(tb_store *s, int64_t *h_nnz_row, int64_t *h_row_sum) {
  cudaStream_t st = s->stream;
  const int64_t N = s->nbins;
  long long *dst = s->world > 1 ? s->i_part.p : s->i_full.p;
  if (s->world == 1) TB_CUDA(cudaMemsetAsync(dst, 0, 2 * N * sizeof(long long), st));
  if (s->nrows) {      // both vectors are kept from the build (finish_layout)
    const int g = grid_1d(s->nrows, 256, s->sm_count);
    k_place_i64<<<g, 256, 0, st>>>(s->row_nnz.p, s->nrows, s->row_begin, dst);
    k_place_i64<<<g, 256, 0, st>>>(s->row_sum.p, s->nrows, s->row_begin, dst + N);
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
  cudaEvent_t e0 = ice_event(s, 0), e1 = ice_event(s, 1);
  TB_CUDA(cudaEventRecord(e0, st));
  masked_row_sums(s, dst);       // column sums over good rows = row sums over good columns
  TB_CUDA(cudaEventRecord(e1, st));
  if (s->world > 1) allreduce_i64(s, s->i_part.p, s->i_full.p, N);
  TB_CUDA(cudaMemcpyAsync(h_col_sum, s->i_full.p, N * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  float ms = 0.f;
  TB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  s->last_kernel_ms = ms;
}

// Stores that kept their CSR because they hold zeros or negative values: cells with a good
// partner per owned row (mark_present).  Positive stores need none of this.
const double *live_counts(tb_store *s) {
  if (s->positive) return nullptr;          // with or without the CSR: the same rule, the same result
  TB_REQUIRE(s->csr_resident, "internal: live counts need the CSR");
  cudaStream_t st = s->stream;
  if (s->live.n < (size_t)s->nrows + 1) s->live.alloc(s->nrows + 1);
  TB_CUDA(cudaMemsetAsync(s->live.p, 0, s->live.bytes(), st));
  if (s->nchunks)
    k_live_rows<<<chunk_grid(s), 256, 0, st>>>(s->col.p, s->chunk_beg.p, s->chunk_row.p, s->nchunks,
                                               s->bad.p, s->live.p);
  return s->live.p;
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
  if (!s->px.ready) exchange_setup(s);
  const bool fused = use_fused_exchange(s);
  set_check_limits(s);
  upload_bad(s, h_bad);
  const int vg = grid_1d(N, 256, s->sm_count);
  int rg = grid_1d(N, kRedThreads, s->sm_count, 4);
  if (rg > 768) rg = 768;
  // every call takes npass_max exchange numbers, used or not: all ranks stay in step
  const unsigned long long epoch_base = s->px.epoch;
  s->px.epoch += (unsigned long long)npass_max + 1ull;
  k_ice_init<<<vg, 256, 0, st>>>(s->bad.p, N, s->x.p, s->b.p, s->state.p, epoch_base);

  // TB_TIMING=1: where the time between two row sums goes, printed per rank after the call
  const char *tenv = getenv("TB_TIMING");
  const bool split = tenv && tenv[0] == '1';
  const size_t ev_pass0 = 2, ev_per_pass = split && !fused && s->world > 1 ? 5 : 3;
  auto ev = [&](int p, int k) { return ice_event(s, ev_pass0 + (size_t)p * ev_per_pass + k); };
  const bool csr_walk = use_csr_rowsum(s);
  const TileSums ts = tile_sums(s);
  const double *live = live_counts(s);
  if (!csr_walk) clear_tile_chunk_partials(s);
  double *dst = s->world > 1 ? s->s_part.p : s->s_full.p;      // separate-step path only
  DevBuf<unsigned long long> prof;
  if (split && fused) {
    prof.alloc(8);
    TB_CUDA(cudaMemsetAsync(prof.p, 0, prof.bytes(), st));
  }
  // The host only needs to know WHEN the loop stopped: passes queued after the stop are no-ops
  // on the device.  ICE converges geometrically, so the next look is scheduled where the
  // deviation is predicted to cross max_dev (at most 16 passes ahead).
  int next_poll = iterations == 0 ? 1 : 4;
  double prev_dev = -1.0;
  int prev_done = 0;
  int launched = 0;
  bool stopped = false;
  TB_CUDA(cudaEventRecord(ice_event(s, 0), st));
  for (int p = 0; p < npass_max && !stopped; ++p) {
    TB_CUDA(cudaEventRecord(ev(p, 0), st));
    rowsum_once(s, s->x.p, s->state.p);
    TB_CUDA(cudaEventRecord(ev(p, 1), st));
    if (fused) {
      (exchange_ll() ? k_ice_exchange<true> : k_ice_exchange<false>)<<<exchange_grid(s), kXThreads, 0, st>>>(
          s->px.view, s->partial_f.p, s->row_chunk.p, ts, s->nrows, s->row_begin, N, s->x.p, s->b.p,
          s->bad.p, s->present.p, live, iterations, max_dev, s->block_red.p, s->trace.p, s->state.p, prof.p);
    } else if (s->world > 1) {
      if (s->nrows)
        k_combine_f64<<<grid_1d(s->nrows, 256, s->sm_count), 256, 0, st>>>(
            s->partial_f.p, s->row_chunk.p, s->nrows, s->row_begin, dst, s->state.p, ts, p == 0,
            s->bad.p, live);
      if (ev_per_pass == 5) TB_CUDA(cudaEventRecord(ev(p, 3), st));
      allreduce_f64(s, s->s_part.p, s->s_full.p, N);
      if (ev_per_pass == 5) TB_CUDA(cudaEventRecord(ev(p, 4), st));
      k_ice_reduce<false><<<rg, kRedThreads, 0, st>>>(
          s->s_full.p, s->x.p, s->bad.p, s->present.p, live, N, p, iterations, max_dev,
          s->block_red.p, s->trace.p, s->state.p, nullptr, nullptr, ts);
      k_ice_update<<<vg, 256, 0, st>>>(s->s_full.p, s->x.p, s->b.p, s->present.p, N, p, s->state.p);
    } else {
      k_ice_reduce<true><<<rg, kRedThreads, 0, st>>>(
          s->s_full.p, s->x.p, s->bad.p, s->present.p, live, N, p, iterations, max_dev,
          s->block_red.p, s->trace.p, s->state.p, s->partial_f.p, s->row_chunk.p, ts);
      k_ice_update<<<vg, 256, 0, st>>>(s->s_full.p, s->x.p, s->b.p, s->present.p, N, p, s->state.p);
    }
    TB_CUDA(cudaEventRecord(ev(p, 2), st));
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
  TB_CUDA(cudaEventRecord(ice_event(s, 1), st));
  TB_CUDA(cudaMemcpyAsync(s->h_state.p, s->state.p, sizeof(IceState), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  const IceState hs = *s->h_state.p;
  int rc = TB_OK;
  if (hs.fault) {
    set_error("ice: a GPU of the group did not reach the row-sum exchange within %.0f s",
              (double)kPeerTimeoutNs * 1e-9);
    rc = TB_ERR_STATE;
  } else if (hs.all_bad) {
    set_error("ERROR: normalization failed, all bad columns");
    rc = TB_ERR_ALL_BAD;
  } else if (hs.zero_div) {
    set_error("float division by zero");
    rc = TB_ERR_ZERO_DIV;
  } else {
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
  TB_CUDA(cudaEventElapsedTime(&ms, ice_event(s, 0), ice_event(s, 1)));
  s->loop_ms = ms;
  s->rowsum_ms = 0;
  s->exchange_ms = 0;
  s->rowsum_launches = 0;
  const int real = rc == TB_OK ? hs.passes_done : 0;
  double t_comb = 0, t_red = 0, t_upd = 0;
  for (int p = 0; p < real && p < launched; ++p) {
    TB_CUDA(cudaEventElapsedTime(&ms, ev(p, 0), ev(p, 1)));
    s->rowsum_ms += ms;
    TB_CUDA(cudaEventElapsedTime(&ms, ev(p, 1), ev(p, 2)));
    s->exchange_ms += ms;
    ++s->rowsum_launches;
    if (ev_per_pass == 5) {
      float a = 0.f, b = 0.f, c = 0.f;
      cudaEventElapsedTime(&a, ev(p, 1), ev(p, 3));
      cudaEventElapsedTime(&b, ev(p, 3), ev(p, 4));
      cudaEventElapsedTime(&c, ev(p, 4), ev(p, 2));
      t_comb += a; t_red += b; t_upd += c;
    }
  }
  {   // kernels of this library launched by the call (NCCL's own kernels are not counted)
    const int k1 = csr_walk ? 1 : (s->t_nunits > 0) + (s->n_groups > 0) + (s->n_sparse > 0 && !sparse_in_dense_tail(s));
    const int k2 = fused ? 1 : 2 + (s->world > 1 && s->nrows ? 1 : 0);   // exchange | (combine) + reduce + update
    s->total_launches = 2 /* init, finalize */ + (live ? 1 : 0) + launched * (k1 + k2);
  }
  if (split && real > 0) {
    const int cnt = real < launched ? real : launched;
    if (ev_per_pass == 5)
      fprintf(stderr, "[tadbit_b200] rank %d: per pass row sum %.1f us, combine %.1f us, all-reduce %.1f us, "
                      "statistics + update %.1f us (%d passes)\n", s->rank, 1e3 * s->rowsum_ms / cnt,
              1e3 * t_comb / cnt, 1e3 * t_red / cnt, 1e3 * t_upd / cnt, cnt);
    else
      fprintf(stderr, "[tadbit_b200] rank %d: per pass row sum %.1f us, %s %.1f us, whole pass %.1f us (%d passes)\n",
              s->rank, 1e3 * s->rowsum_ms / cnt, fused ? "fused exchange + statistics + update" : "statistics + update",
              1e3 * s->exchange_ms / cnt, 1e3 * s->loop_ms / cnt, cnt);
    if (prof.p) {
      unsigned long long h[8];
      TB_CUDA(cudaMemcpy(h, prof.p, sizeof(h), cudaMemcpyDeviceToHost));
      fprintf(stderr, "[tadbit_b200] rank %d: k_ice_exchange CTA 0 per pass: row sums + peer stores %.1f us, fence + "
                      "CTA barrier %.1f us, arrive + publish + wait for peers %.1f us, statistics + arrive %.1f us, "
                      "fold + wait for decision %.1f us, update %.1f us\n", s->rank, 1e-3 * h[0] / cnt,
              1e-3 * h[1] / cnt, 1e-3 * h[2] / cnt, 1e-3 * h[3] / cnt, 1e-3 * h[4] / cnt, 1e-3 * h[5] / cnt);
    }
  }
  return rc;
}

void norm_sum(tb_store *s, const double *h_bias, const uint8_t *h_bad, double *out) {
  cudaStream_t st = s->stream;
  const int64_t N = s->nbins;
  TB_REQUIRE(out, "norm_sum: null output");
  set_check_limits(s);
  upload_bad(s, h_bad);
  if (!h_bias) {        // raw sum: exact integers
    masked_row_sums(s, s->i_full.p);
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
  TB_CUDA(cudaMemsetAsync(s->scal_i.p, 0, sizeof(long long), st));
  k_inverse_bias<<<grid_1d(N, 256, s->sm_count), 256, 0, st>>>(s->b.p, s->bad.p, N, s->x.p,
                                                               reinterpret_cast<int *>(s->scal_i.p));
  if (!use_csr_rowsum(s)) clear_tile_chunk_partials(s);
  rowsum_once(s, s->x.p, nullptr);
  if (s->nrows)
    k_combine_f64<<<grid_1d(s->nrows, 256, s->sm_count), 256, 0, st>>>(
        s->partial_f.p, s->row_chunk.p, s->nrows, s->row_begin, s->s_full.p, nullptr, tile_sums(s),
        false, nullptr, nullptr);
  k_dot_rows<<<1, 1024, 0, st>>>(s->x.p, s->s_full.p, s->row_begin, s->row_end, s->scal_f.p);
  if (s->world > 1) allreduce_f64(s, s->scal_f.p, s->scal_f.p + 8, 1);
  int zero_bias = 0;
  TB_CUDA(cudaMemcpyAsync(&zero_bias, s->scal_i.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(out, s->world > 1 ? s->scal_f.p + 8 : s->scal_f.p, sizeof(double),
                          cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  if (zero_bias) {      // v / bias[i] / bias[j] raises in the reference (hic_data.py:371-374)
    set_error("float division by zero");
    throw Fail{TB_ERR_ZERO_DIV};
  }
}

void diag_sums(tb_store *s, const uint8_t *h_bad, int64_t ndiag, int64_t *h_sum_d) {
  cudaStream_t st = s->stream;
  TB_REQUIRE(ndiag >= 0 && h_sum_d, "diag_sums: bad arguments");
  if (ndiag == 0) return;
  upload_bad(s, h_bad);
  DevBuf<unsigned long long> hist, red;
  hist.alloc(ndiag);
  TB_CUDA(cudaMemsetAsync(hist.p, 0, hist.bytes(), st));
  cudaEvent_t e0 = ice_event(s, 0), e1 = ice_event(s, 1);
  TB_CUDA(cudaEventRecord(e0, st));
  DiagVisitor vis;
  vis.bad = s->bad.p;
  vis.chrom_of = s->chrom_of.p;
  vis.chrom_end = s->chrom_end.p;
  vis.ndiag_rows = s->nbins;
  vis.jend = 0;
  vis.ndiag = ndiag;
  vis.hist = hist.p;
  vis.lo = vis.hi = nullptr;
  visit_upper(s, vis);
  TB_CUDA(cudaEventRecord(e1, st));
  unsigned long long *src = hist.p;
  if (s->world > 1) {
    red.alloc(ndiag);
    allreduce_i64(s, (const long long *)hist.p, (long long *)red.p, ndiag);
    src = red.p;
  }
  TB_CUDA(cudaMemcpyAsync(h_sum_d, src, ndiag * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  float ms = 0.f;
  TB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  s->last_kernel_ms = ms;
}

void ice_batch(tb_store *s, const uint8_t *h_bad, int32_t iterations, double max_dev, double *h_bias,
               double *h_trace, int32_t *n_passes) {
  cudaStream_t st = s->stream;
  const int64_t N = s->nbins;
  TB_REQUIRE(iterations >= 0, "ice_batch: iterations must be >= 0");
  TB_REQUIRE(s->world == 1, "ice_batch: batches are not sharded (one replica per GPU)");
  set_check_limits(s);
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
  k_ice_init<<<vg, 256, 0, st>>>(s->bad.p, N, s->x.p, s->b.p, s->state.p, 0ull);
  k_seg_init<<<(nseg + 255) / 256, 256, 0, st>>>(seg.p, nseg, n_stopped.p);
  const bool csr_walk = use_csr_rowsum(s);
  const TileSums ts = tile_sums(s);
  const double *live = live_counts(s);
  if (!csr_walk) clear_tile_chunk_partials(s);
  cudaEvent_t e0 = ice_event(s, 0), e1 = ice_event(s, 1);
  TB_CUDA(cudaEventRecord(e0, st));
  bool stopped = false;
  int launched = 0;
  for (int p = 0; p < npass_max && !stopped; ++p) {
    rowsum_once(s, s->x.p, s->state.p);
    k_seg_reduce<<<nseg, 1024, 0, st>>>(s->s_full.p, s->x.p, s->bad.p, s->present.p, seg_off.p, nseg,
                                        N, p, iterations, max_dev, trace.p, npass_max, seg.p,
                                        n_stopped.p, s->state.p, s->partial_f.p, s->row_chunk.p, ts, live);
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
  // -1: no row left in the matrix (all bad columns); -2: its mean row sum is 0 (float division by zero)
  for (int k = 0; k < nseg; ++k) n_passes[k] = hseg[k].all_bad ? -hseg[k].all_bad : hseg[k].passes_done;
  float ms = 0.f;
  TB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  s->loop_ms = ms;
  s->rowsum_ms = 0;
  s->exchange_ms = 0;
  s->rowsum_launches = 0;
  {
    const int k1 = csr_walk ? 1 : (s->t_nunits > 0) + (s->n_groups > 0) + (s->n_sparse > 0 && !sparse_in_dense_tail(s));
    s->total_launches = 3 + (live ? 1 : 0) + launched * (k1 + 2);
  }
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
                                                                  s->state.p, s->px.epoch);
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
  const char *tenv = getenv("TB_TIMING");
  if (tenv && tenv[0] == '1' && s->t_nunits) {     // when do the CTAs of the tile kernel start and finish?
    const int grid = (int)(s->t_nunits < s->sm_count ? s->t_nunits : s->sm_count);
    DevBuf<unsigned long long> prof;
    prof.alloc(4 * (size_t)grid);
    launch_rowsum_tiles(s, s->x.p, nullptr, prof.p);
    std::vector<unsigned long long> h(4 * (size_t)grid);
    TB_CUDA(cudaMemcpyAsync(h.data(), prof.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    unsigned long long t0 = ~0ull, t1 = 0, umin = ~0ull, umax = 0, utot = 0, ttot = 0;
    double end_sum = 0.0, end_min = 1e30;
    for (int b = 0; b < grid; ++b) t0 = h[4 * b] < t0 ? h[4 * b] : t0;
    for (int b = 0; b < grid; ++b) {
      const double end = (double)(h[4 * b + 1] - t0);
      end_sum += end;
      end_min = end < end_min ? end : end_min;
      t1 = h[4 * b + 1] > t1 ? h[4 * b + 1] : t1;
      umin = h[4 * b + 2] < umin ? h[4 * b + 2] : umin;
      umax = h[4 * b + 2] > umax ? h[4 * b + 2] : umax;
      utot += h[4 * b + 2];
      ttot += h[4 * b + 3];
    }
    fprintf(stderr, "[tadbit_b200] k_rowsum_tiles: %d CTAs, %llu units (%llu..%llu per CTA), %llu tile visits; "
                    "CTAs finish %.1f us (first) / %.1f us (mean) / %.1f us (last) after the first one starts\n",
            grid, utot, umin, umax, ttot, 1e-3 * end_min, 1e-3 * end_sum / grid, 1e-3 * (double)(t1 - t0));
  }
}

// Mean device time of one K1 evaluation as the ICE loop issues it: ms[0] with its two kernels
// back to back on one stream, ms[1] with the dense kernel on a second stream next to the tile
// kernel (rowsum.cu:launch_rowsum; option k1_streams).  Also on an unconnected shard.
void probe_k1(tb_store *s, int32_t reps, double ms[2]) {
  cudaStream_t st = s->stream;
  TB_REQUIRE(reps > 0 && ms, "probe_k1: bad arguments");
  TB_CUDA(cudaMemsetAsync(s->bad.p, 0, s->nbins, st));
  k_ice_init<<<grid_1d(s->nbins, 256, s->sm_count), 256, 0, st>>>(s->bad.p, s->nbins, s->x.p, s->b.p,
                                                                  s->state.p, s->px.epoch);
  const int keep = s->k1_streams_mode;
  cudaEvent_t e0 = ice_event(s, 0), e1 = ice_event(s, 1);
  for (int mode = 1; mode <= 2; ++mode) {
    s->k1_streams_mode = mode;
    launch_rowsum(s, s->x.p, nullptr);               // warm-up (creates the second stream)
    TB_CUDA(cudaEventRecord(e0, st));
    for (int r = 0; r < reps; ++r) launch_rowsum(s, s->x.p, nullptr);
    TB_CUDA(cudaEventRecord(e1, st));
    TB_CUDA(cudaStreamSynchronize(st));
    float f = 0.f;
    TB_CUDA(cudaEventElapsedTime(&f, e0, e1));
    ms[mode - 1] = f / reps;
  }
  s->k1_streams_mode = keep;
}

}  // namespace tb
