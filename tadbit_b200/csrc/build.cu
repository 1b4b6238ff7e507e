// Contact-store construction: upper-triangle COO -> device-resident CSR of the owned
// rows of the FULL symmetric matrix (both triangles, diagonal once), columns ascending.
//
// Why both triangles: every reduction of the hot path (ICE row sums, filter statistics,
// masked column sums) is a sum over a full matrix row (SURVEY.md 8a' trap 1).  Holding the
// mirrored cells turns all of them into gather-only, atomic-free, deterministic row
// reductions of one stream; the reference's dict holds both triangles too
// (_pytadbit/hic_data.py:54-118).
//
// The CSR is scaffolding: finish_layout derives the compact row-sum stream (encode.cu, tiles.cu)
// and the per-row statistics from it and then RELEASES it when every stored value is > 0 (what
// the parsers produce); every later kernel walks the compact stream (visit.cuh for the ones that
// need the cells).  A matrix with stored zeros or negative values keeps its CSR: an explicit zero
// cannot be told from an absent cell in a dense panel.  TB_KEEP_CSR=1 keeps it always (tests).
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <cub/cub.cuh>

#include "common.cuh"
#include "visit.cuh"

namespace tb {

namespace {

__global__ void k_validate_count(const int32_t *__restrict__ row, const int32_t *__restrict__ col,
                                 int64_t nnz, int64_t nbins, int64_t rb, int64_t re,
                                 unsigned long long *n_out, int *bad_input) {
  // bad_input[0] invalid cell, [1] duplicate (set later), [2] cells not in strict row-major order
  unsigned long long mine = 0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz;
       k += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = row[k], j = col[k];
    if (i < 0 || j < i || j >= nbins) {
      *bad_input = 1;
      continue;
    }
    if (k + 1 < nnz) {
      const int64_t i2 = row[k + 1], j2 = col[k + 1];
      if (i2 < i || (i2 == i && j2 <= j)) bad_input[2] = 1;
    }
    mine += (i >= rb && i < re);
    mine += (i != j && j >= rb && j < re);
  }
  for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_out, mine);
}

// Emit the directed cells (row in block) as sortable keys (local_row << 32 | col).
__global__ void k_emit(const int32_t *__restrict__ row, const int32_t *__restrict__ col,
                       const int32_t *__restrict__ val, int64_t nnz, int64_t rb, int64_t re,
                       unsigned long long *cursor, uint64_t *keys, int32_t *vals) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t nround = (nnz + stride - 1) / stride;
  for (int64_t r = 0; r < nround; ++r) {
    int64_t k = r * stride + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t i = 0, j = 0;
    int32_t v = 0;
    int c1 = 0, c2 = 0;
    if (k < nnz) {
      i = row[k];
      j = col[k];
      v = val[k];
      c1 = (i >= rb && i < re);
      c2 = (i != j && j >= rb && j < re);
    }
    int cnt = c1 + c2, incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 31 && total) base = atomicAdd(cursor, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    unsigned long long pos = base + (incl - cnt);
    if (c1) {
      keys[pos] = ((uint64_t)(i - rb) << 32) | (uint32_t)j;
      vals[pos] = v;
      ++pos;
    }
    if (c2) {
      keys[pos] = ((uint64_t)(j - rb) << 32) | (uint32_t)i;
      vals[pos] = v;
    }
  }
}

// keys sorted: split into col[] and rowptr[], flag duplicate cells.
__global__ void k_split_keys(const uint64_t *__restrict__ keys, int64_t m, int64_t nrows,
                             int32_t *col, int64_t *rowptr, int *dup) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < m;
       k += (int64_t)gridDim.x * blockDim.x) {
    uint64_t key = keys[k];
    int64_t r = (int64_t)(key >> 32);
    col[k] = (int32_t)(key & 0xffffffffu);
    int64_t rprev = -1;
    if (k) {
      uint64_t kp = keys[k - 1];
      if (kp == key) *dup = 1;
      rprev = (int64_t)(kp >> 32);
    }
    for (int64_t rr = rprev + 1; rr <= r; ++rr) rowptr[rr] = k;
    if (k == m - 1)
      for (int64_t rr = r + 1; rr <= nrows; ++rr) rowptr[rr] = m;
  }
}

// ---- fast path for row-major sorted input (whole matrix) ------------------------------------
// rows sorted: upper row pointers from the row boundaries; column counts by atomics
__global__ void k_sorted_counts(const int32_t *__restrict__ row, const int32_t *__restrict__ col,
                                int64_t nnz, int64_t nbins, int64_t *__restrict__ up_ptr,
                                unsigned int *__restrict__ col_cnt) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = row[k], rprev = k ? row[k - 1] : -1;
    for (int64_t rr = rprev + 1; rr <= r; ++rr) up_ptr[rr] = k;
    if (k == nnz - 1)
      for (int64_t rr = r + 1; rr <= nbins; ++rr) up_ptr[rr] = nnz;
    atomicAdd(&col_cnt[col[k]], 1u);
  }
}

// full-row length = cells above the diagonal in this column + cells of the upper row
__global__ void k_full_counts(const int64_t *__restrict__ up_ptr, const unsigned int *__restrict__ col_cnt,
                              const int32_t *__restrict__ col, int64_t nbins, int64_t *__restrict__ full_cnt,
                              int64_t *__restrict__ col_cnt64) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nbins;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t up = up_ptr[r + 1] - up_ptr[r];
    const int has_diag = up > 0 && col[up_ptr[r]] == r;      // first cell of the upper row
    col_cnt64[r] = col_cnt[r];
    full_cnt[r] = (int64_t)col_cnt[r] - has_diag + up;
  }
}

__global__ void k_fill_upper(const int32_t *__restrict__ row, const int32_t *__restrict__ col,
                             const int32_t *__restrict__ val, int64_t nnz,
                             const int64_t *__restrict__ up_ptr, const int64_t *__restrict__ rowptr,
                             int32_t *__restrict__ ocol, int32_t *__restrict__ oval) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = row[k];
    const int64_t up = up_ptr[r + 1] - up_ptr[r];
    const int64_t dest = rowptr[r + 1] - up + (k - up_ptr[r]);   // upper cells close the row
    ocol[dest] = col[k];
    oval[dest] = val[k];
  }
}

// cells stably sorted by column: position p of column j's run is the p-th smallest row
__global__ void k_fill_lower(const uint32_t *__restrict__ sorted_col, const uint32_t *__restrict__ idx,
                             const int32_t *__restrict__ row, const int32_t *__restrict__ val, int64_t nnz,
                             const int64_t *__restrict__ cptr, const int64_t *__restrict__ rowptr,
                             int32_t *__restrict__ ocol, int32_t *__restrict__ oval) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz;
       p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = sorted_col[p];
    const uint32_t k = idx[p];
    const int32_t i = row[k];
    if (i == j) continue;                                       // the diagonal is stored once
    const int64_t dest = rowptr[j] + (p - cptr[j]);
    ocol[dest] = i;
    oval[dest] = val[k];
  }
}

__global__ void k_iota_u32(uint32_t *p, int64_t n) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
       k += (int64_t)gridDim.x * blockDim.x)
    p[k] = (uint32_t)k;
}

__global__ void k_fill_i64(int64_t *p, int64_t n, int64_t v) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
       k += (int64_t)gridDim.x * blockDim.x)
    p[k] = v;
}

__device__ __forceinline__ int64_t lower_bound_col(const int32_t *__restrict__ col, int64_t lo,
                                                   int64_t hi, int64_t key) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (col[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Split every row into chunks (the warp work items).  Columns are cut at multiples of kPanel:
// a panel holding >= 1/dense_fill of its columns becomes one chunk on its own (a candidate for the
// dense encodings, and aligned with the same panel of every other row so that several rows
// can share one pass over x); runs of emptier panels are merged up to kChunk cells.
// FILL == false counts the chunks of each row, FILL == true writes them.
template <bool FILL>
__global__ void k_plan_chunks(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                              int64_t nrows, int64_t nbins, int dense_fill,
                              int64_t *__restrict__ nchunk,
                              const int64_t *__restrict__ row_chunk, int64_t *__restrict__ chunk_beg,
                              int32_t *__restrict__ chunk_row) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = rowptr[r], e = rowptr[r + 1];
    int64_t n = 0, c = FILL ? row_chunk[r] : 0;
    int64_t pos = b;
    while (pos < e) {
      const int64_t p = col[pos] / kPanel;
      const int64_t pend = lower_bound_col(col, pos, e, (p + 1) * kPanel);
      int64_t width = nbins - p * kPanel;
      if (width > kPanel) width = kPanel;
      int64_t stop = pend;
      if ((pend - pos) * dense_fill < width) {  // sparse: absorb the following sparse panels
        while (stop < e) {
          const int64_t p2 = col[stop] / kPanel;
          const int64_t q2 = lower_bound_col(col, stop, e, (p2 + 1) * kPanel);
          int64_t w2 = nbins - p2 * kPanel;
          if (w2 > kPanel) w2 = kPanel;
          if ((q2 - stop) * dense_fill >= w2 || q2 - pos > kChunk) break;
          stop = q2;
        }
      }
      if (FILL) {
        chunk_beg[c] = pos;
        chunk_row[c] = (int32_t)r;
        ++c;
      }
      ++n;
      pos = stop;
    }
    if (!FILL) nchunk[r] = n;
  }
}

// per owned row: stored cells and their sum; over the block: zeros, negative values, sum of |v|
__global__ void __launch_bounds__(256)
k_build_row_stats(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ val, int64_t nrows,
                  long long *__restrict__ row_nnz, long long *__restrict__ row_sum,
                  unsigned long long *__restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long zeros = 0, negs = 0, mag = 0;
  for (int64_t r = warp; r < nrows; r += nwarp) {
    const int64_t b = rowptr[r], e = rowptr[r + 1];
    long long a = 0;
    for (int64_t k = b + lane; k < e; k += 32) {
      const int32_t v = val[k];
      a += v;
      zeros += v == 0;
      negs += v < 0;
      mag += (unsigned long long)(v < 0 ? -(long long)v : (long long)v);
    }
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) { row_nnz[r] = e - b; row_sum[r] = a; }
  }
  for (int o = 16; o; o >>= 1) {
    zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    negs += __shfl_xor_sync(0xffffffffu, negs, o);
    mag += __shfl_xor_sync(0xffffffffu, mag, o);
  }
  if (lane == 0) {
    if (zeros) atomicAdd(stats, zeros);
    if (negs) atomicAdd(stats + 1, negs);
    if (mag) atomicAdd(stats + 2, mag);
  }
}

// ---- export of the stored upper-triangle cells from the compact stream (visit.cuh) ----------
struct CountUpperVisitor {
  static constexpr int kShared = 0;
  static constexpr int kThreads = 256;
  int64_t rb;
  unsigned long long *cnt;            // [nrows]
  __device__ void begin(uint8_t *) {}
  __device__ bool block(int64_t, int64_t, int64_t, int64_t) { return true; }
  __device__ bool row(int64_t) { return true; }
  __device__ void cell(int64_t i, int64_t, int32_t) { atomicAdd(&cnt[i - rb], 1ull); }
  __device__ void end(uint8_t *) {}
};
struct PlaceUpperVisitor {
  static constexpr int kShared = 0;
  static constexpr int kThreads = 256;
  int64_t rb;
  const int64_t *off;                 // [nrows] first output slot of each row
  unsigned long long *cursor;         // [nrows]
  uint64_t *key;                      // (local row << 32) | column
  int32_t *val;
  __device__ void begin(uint8_t *) {}
  __device__ bool block(int64_t, int64_t, int64_t, int64_t) { return true; }
  __device__ bool row(int64_t) { return true; }
  __device__ void cell(int64_t i, int64_t j, int32_t v) {
    const int64_t r = i - rb;
    const int64_t pos = off[r] + (int64_t)atomicAdd(&cursor[r], 1ull);
    key[pos] = ((uint64_t)r << 32) | (uint32_t)j;
    val[pos] = v;
  }
  __device__ void end(uint8_t *) {}
};
__global__ void k_unpack_keys(const uint64_t *__restrict__ key, int64_t n, int64_t rb, int32_t row_off,
                              int32_t col_off, int32_t *__restrict__ orow, int32_t *__restrict__ ocol) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
       k += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t e = key[k];
    if (orow) orow[k] = (int32_t)(rb + (int64_t)(e >> 32)) + row_off;
    ocol[k] = (int32_t)(e & 0xffffffffu) + col_off;
  }
}

// first stored cell of each owned row with col >= its own (global) row index
__global__ void k_upper_start(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                              int64_t nrows, int64_t rb, bool whole_row, int64_t *ustart, int64_t *ucount) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = rowptr[r], hi = rowptr[r + 1], end = hi;
    int32_t g = whole_row ? 0 : (int32_t)(rb + r);
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (col[mid] < g) lo = mid + 1; else hi = mid;
    }
    ustart[r] = lo;
    ucount[r] = end - lo;
  }
}

__global__ void k_copy_upper(const int32_t *__restrict__ col, const int32_t *__restrict__ val,
                             const int64_t *__restrict__ ustart, const int64_t *__restrict__ ucount,
                             const int64_t *__restrict__ uoff, int64_t nrows, int64_t rb, int32_t bin_off,
                             int32_t *orow, int32_t *ocol, int32_t *oval) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < nrows; r += nwarp) {
    int64_t src = ustart[r], n = ucount[r], dst = uoff[r];
    for (int64_t k = lane; k < n; k += 32) {
      if (orow) orow[dst + k] = (int32_t)(rb + r) + bin_off;
      ocol[dst + k] = col[src + k] + bin_off;
      oval[dst + k] = val[src + k];
    }
  }
}

inline int grid_for(int64_t n, int threads, int sm_count) {
  int64_t g = (n + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

void exclusive_scan_i64(cudaStream_t st, const int64_t *in, int64_t *out, int64_t n) {
  size_t tmp_bytes = 0;
  TB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, n, st));
  DevBuf<uint8_t> tmp;
  tmp.alloc(tmp_bytes ? tmp_bytes : 1);
  TB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, n, st));
  TB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace

namespace {
struct StageTimer {            // TB_TIMING=1: wall time of the build stages on stderr
  bool on;
  cudaStream_t st;
  std::chrono::steady_clock::time_point t0;
  explicit StageTimer(cudaStream_t s) : st(s) {
    const char *e = getenv("TB_TIMING");
    on = e && e[0] == '1';
    t0 = std::chrono::steady_clock::now();
  }
  void lap(const char *what) {
    if (!on) return;
    cudaStreamSynchronize(st);
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[tadbit_b200] %-28s %8.1f ms\n", what,
            std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};
}  // namespace

void build_from_coo(tb_store *s, int64_t nnz, const int32_t *row, const int32_t *col,
                    const int32_t *val, int mem) {
  cudaStream_t st = s->stream;
  StageTimer tm(st);
  DevBuf<int32_t> drow, dcol, dval;
  const int32_t *r = row, *c = col, *v = val;
  if (mem == TB_MEM_HOST && nnz) {
    drow.alloc(nnz); dcol.alloc(nnz); dval.alloc(nnz);
    TB_CUDA(cudaMemcpyAsync(drow.p, row, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    TB_CUDA(cudaMemcpyAsync(dcol.p, col, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    TB_CUDA(cudaMemcpyAsync(dval.p, val, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    r = drow.p; c = dcol.p; v = dval.p;
  }
  tm.lap("H2D cells");
  DevBuf<unsigned long long> counters;   // [0] directed cells, [1] emit cursor
  DevBuf<int> flags;                     // [0] bad input, [1] duplicate, [2] not row-major sorted
  counters.alloc(2);
  flags.alloc(4);
  TB_CUDA(cudaMemsetAsync(counters.p, 0, counters.bytes(), st));
  TB_CUDA(cudaMemsetAsync(flags.p, 0, flags.bytes(), st));
  const int threads = 256;
  if (nnz)
    k_validate_count<<<grid_for(nnz, threads, s->sm_count), threads, 0, st>>>(
        r, c, nnz, s->nbins, s->row_begin, s->row_end, counters.p, flags.p);
  unsigned long long h_cnt[2] = {0, 0};
  int h_flags[4] = {0, 0, 0, 0};
  TB_CUDA(cudaMemcpyAsync(h_cnt, counters.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(h_flags, flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  TB_REQUIRE(!h_flags[0], "contact store: a cell has row < 0, row > col or col >= nbins");
  const int64_t m = (int64_t)h_cnt[0];
  s->nnz_full = m;
  s->rowptr.alloc(s->nrows + 1);
  s->col.alloc(m ? m : 1);
  s->val.alloc(m ? m : 1);
  if (m == 0) {
    k_fill_i64<<<grid_for(s->nrows + 1, threads, s->sm_count), threads, 0, st>>>(
        s->rowptr.p, s->nrows + 1, 0);
    TB_CUDA(cudaStreamSynchronize(st));
    finish_layout(s);
    return;
  }
  const char *no_fast = getenv("TB_NO_SORTED_PATH");
  if (!h_flags[2] && s->row_begin == 0 && s->row_end == s->nbins && nnz < 4294967295LL &&
      !(no_fast && no_fast[0] == '1')) {
    // Row-major sorted cells of the whole matrix (what HiC_data and tb_store_export_upper
    // produce): the upper rows are already in place, and the mirrored half only needs the cells
    // stably sorted by COLUMN -- a 32-bit key sort of nnz items instead of a 64-bit key sort of
    // 2 nnz.  Strictly increasing (row, col) also rules out duplicates.
    const int64_t N = s->nbins;
    DevBuf<int64_t> up_ptr, full_cnt, col_cnt64, cptr;
    DevBuf<unsigned int> col_cnt;
    up_ptr.alloc(N + 1); full_cnt.alloc(N + 1); col_cnt64.alloc(N + 1); cptr.alloc(N + 1);
    col_cnt.alloc(N + 1);
    TB_CUDA(cudaMemsetAsync(col_cnt.p, 0, col_cnt.bytes(), st));
    TB_CUDA(cudaMemsetAsync(full_cnt.p, 0, full_cnt.bytes(), st));
    TB_CUDA(cudaMemsetAsync(col_cnt64.p, 0, col_cnt64.bytes(), st));
    k_sorted_counts<<<grid_for(nnz, threads, s->sm_count), threads, 0, st>>>(r, c, nnz, N, up_ptr.p,
                                                                             col_cnt.p);
    k_full_counts<<<grid_for(N, threads, s->sm_count), threads, 0, st>>>(up_ptr.p, col_cnt.p, c, N,
                                                                         full_cnt.p, col_cnt64.p);
    exclusive_scan_i64(st, full_cnt.p, s->rowptr.p, N + 1);
    exclusive_scan_i64(st, col_cnt64.p, cptr.p, N + 1);
    k_fill_upper<<<grid_for(nnz, threads, s->sm_count), threads, 0, st>>>(r, c, v, nnz, up_ptr.p,
                                                                          s->rowptr.p, s->col.p, s->val.p);
    {
      DevBuf<uint32_t> key_out, idx_in, idx_out;
      key_out.alloc(nnz); idx_in.alloc(nnz); idx_out.alloc(nnz);
      k_iota_u32<<<grid_for(nnz, threads, s->sm_count), threads, 0, st>>>(idx_in.p, nnz);
      int bits = 1;
      while (((int64_t)1 << bits) < N) ++bits;
      size_t tmp_bytes = 0;
      const uint32_t *key_in = reinterpret_cast<const uint32_t *>(c);
      TB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_in, key_out.p, idx_in.p,
                                              idx_out.p, nnz, 0, bits, st));
      DevBuf<uint8_t> tmp;
      tmp.alloc(tmp_bytes ? tmp_bytes : 1);
      TB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key_in, key_out.p, idx_in.p,
                                              idx_out.p, nnz, 0, bits, st));
      k_fill_lower<<<grid_for(nnz, threads, s->sm_count), threads, 0, st>>>(
          key_out.p, idx_out.p, r, v, nnz, cptr.p, s->rowptr.p, s->col.p, s->val.p);
      TB_CUDA(cudaStreamSynchronize(st));
    }
    tm.lap("mirror by column sort");
    drow.release(); dcol.release(); dval.release(); counters.release(); flags.release();
    tm.lap("free input copies");
    finish_layout(s);
    return;
  }
  {
    DevBuf<uint64_t> keys_in, keys_out;
    DevBuf<int32_t> vals_in;
    keys_in.alloc(m); keys_out.alloc(m); vals_in.alloc(m);
    k_emit<<<grid_for(nnz, threads, s->sm_count), threads, 0, st>>>(
        r, c, v, nnz, s->row_begin, s->row_end, counters.p + 1, keys_in.p, vals_in.p);
    int row_bits = 1;
    while (((int64_t)1 << row_bits) < s->nrows) ++row_bits;
    const int end_bit = 32 + row_bits;
    size_t tmp_bytes = 0;
    TB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in.p, keys_out.p, vals_in.p,
                                            s->val.p, m, 0, end_bit, st));
    DevBuf<uint8_t> tmp;
    tmp.alloc(tmp_bytes ? tmp_bytes : 1);
    TB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_in.p, keys_out.p, vals_in.p,
                                            s->val.p, m, 0, end_bit, st));
    k_split_keys<<<grid_for(m, threads, s->sm_count), threads, 0, st>>>(
        keys_out.p, m, s->nrows, s->col.p, s->rowptr.p, flags.p + 1);
    TB_CUDA(cudaMemcpyAsync(h_flags, flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    TB_REQUIRE(!h_flags[1], "contact store: duplicate cells (each upper-triangle cell must appear once)");
  }
  tm.lap("mirror + sort + split");
  drow.release(); dcol.release(); dval.release(); counters.release(); flags.release();
  tm.lap("free input copies");
  finish_layout(s);
}

namespace {
__global__ void k_expand_rows(const int64_t *__restrict__ ptr, int64_t nrows, int32_t *__restrict__ row) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < nrows; r += nwarp)
    for (int64_t k = ptr[r] + lane; k < ptr[r + 1]; k += 32) row[k] = (int32_t)r;
}
// whole rows handed over as they are: columns in range and strictly ascending inside each row
__global__ void k_check_rows(const int64_t *__restrict__ ptr, const int32_t *__restrict__ col, int64_t nrows,
                             int64_t nbins, int64_t nnz, int *bad_input) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = ptr[r], e = ptr[r + 1];
    if (b > e || b < 0 || e > nnz || (r == 0 && b != 0) || (r == nrows - 1 && e != nnz)) { *bad_input = 1; continue; }
    for (int64_t k = b; k < e; ++k)
      if (col[k] < 0 || col[k] >= nbins || (k > b && col[k] <= col[k - 1])) { *bad_input = 1; break; }
  }
}
}  // namespace

// Cells in CSR form.  TB_CELLS_UPPER: the upper-triangle rows of the WHOLE matrix (ptr has nbins + 1
// entries) -- 8 bytes per cell over PCIe instead of the 12 of the coordinate form; the row index
// is expanded on the device and the cells then take the path of build_from_coo.
// TB_CELLS_FULL_ROWS: every stored cell of the owned rows [row_begin, row_end) (both triangles,
// ptr has nrows + 1 entries, columns ascending): the pre-partitioned input of a sharded matrix.
// The arrays ARE the store's CSR, nothing is mirrored or sorted.
void build_from_csr(tb_store *s, int form, const int64_t *ptr, const int32_t *col, const int32_t *val, int mem) {
  cudaStream_t st = s->stream;
  StageTimer tm(st);
  const int threads = 256;
  const int64_t nptr = (form == TB_CELLS_UPPER ? s->nbins : s->nrows) + 1;
  DevBuf<int64_t> dptr;
  const int64_t *p = ptr;
  int64_t nnz = 0;
  if (mem == TB_MEM_HOST) {
    nnz = ptr[nptr - 1];
    dptr.alloc(nptr);
    TB_CUDA(cudaMemcpyAsync(dptr.p, ptr, nptr * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    p = dptr.p;
  } else {
    TB_CUDA(cudaMemcpy(&nnz, ptr + nptr - 1, sizeof(int64_t), cudaMemcpyDeviceToHost));
  }
  TB_REQUIRE(nnz >= 0 && (nnz == 0 || (col && val)), "contact store: bad row pointers or null cell arrays");
  if (form == TB_CELLS_UPPER) {
    TB_REQUIRE(s->row_begin == 0 && s->row_end == s->nbins,
               "TB_CELLS_UPPER describes the whole matrix; shards take TB_CELLS_FULL_ROWS");
    DevBuf<int32_t> drow, dcol, dval;
    drow.alloc(nnz ? nnz : 1);
    const int32_t *c = col, *v = val;
    if (mem == TB_MEM_HOST && nnz) {
      dcol.alloc(nnz); dval.alloc(nnz);
      TB_CUDA(cudaMemcpyAsync(dcol.p, col, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      TB_CUDA(cudaMemcpyAsync(dval.p, val, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      c = dcol.p; v = dval.p;
    }
    if (nnz)
      k_expand_rows<<<grid_for(s->nbins * 32, threads, s->sm_count), threads, 0, st>>>(p, s->nbins, drow.p);
    tm.lap("H2D cells (CSR form)");
    build_from_coo(s, nnz, drow.p, c, v, TB_MEM_DEVICE);
    return;
  }
  s->nnz_full = nnz;
  s->rowptr.alloc(s->nrows + 1);
  s->col.alloc(nnz ? nnz : 1);
  s->val.alloc(nnz ? nnz : 1);
  TB_CUDA(cudaMemcpyAsync(s->rowptr.p, p, (s->nrows + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  const cudaMemcpyKind kind = mem == TB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  if (nnz) {
    TB_CUDA(cudaMemcpyAsync(s->col.p, col, nnz * sizeof(int32_t), kind, st));
    TB_CUDA(cudaMemcpyAsync(s->val.p, val, nnz * sizeof(int32_t), kind, st));
  }
  DevBuf<int> flag;
  flag.alloc(1);
  TB_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
  if (s->nrows)
    k_check_rows<<<grid_for(s->nrows, threads, s->sm_count), threads, 0, st>>>(s->rowptr.p, s->col.p, s->nrows,
                                                                               s->nbins, nnz, flag.p);
  int h_flag = 0;
  TB_CUDA(cudaMemcpyAsync(&h_flag, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  TB_REQUIRE(!h_flag, "contact store: row pointers not ascending, or a row's columns out of range / not "
                      "strictly ascending");
  tm.lap("H2D rows (CSR form)");
  dptr.release();
  flag.release();
  finish_layout(s);
}

// A panel is encoded densely (explicit zeros) when at least 1/F of its columns hold a cell.  By
// bytes alone the break-even against 4-byte tile cells is F = 4; measured per stored cell the
// dense kernel is faster than the tile kernel, but explicit zeros at low fill cost more than
// that buys: F = 8 and 16 were slower than 4 on the genome-wide 5 kb matrix (TB_DENSE_FILL to try).
int dense_fill_factor() {
  static int f = 0;
  if (!f) {
    const char *e = getenv("TB_DENSE_FILL");
    f = e ? atoi(e) : 4;
    if (f < 1) f = 1;
    if (f > 64) f = 64;
  }
  return f;
}

void finish_layout(tb_store *s) {
  cudaStream_t st = s->stream;
  StageTimer tm(st);
  const int threads = 256;
  const int64_t nrows = s->nrows, N = s->nbins;
  // per-row statistics that outlive the CSR, and what kind of values the block holds
  s->row_nnz.alloc(nrows + 1);
  s->row_sum.alloc(nrows + 1);
  {
    DevBuf<unsigned long long> stats;
    stats.alloc(4);
    TB_CUDA(cudaMemsetAsync(stats.p, 0, stats.bytes(), st));
    if (nrows)
      k_build_row_stats<<<grid_for(nrows * 32, threads, s->sm_count), threads, 0, st>>>(
          s->rowptr.p, s->val.p, nrows, s->row_nnz.p, s->row_sum.p, stats.p);
    unsigned long long h[4] = {0, 0, 0, 0};
    TB_CUDA(cudaMemcpyAsync(h, stats.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    s->n_zero_cells = (int64_t)h[0];
    // exact integer sums in fp64 need every partial sum below 2^53
    s->positive = h[0] == 0 && h[1] == 0 && h[2] < (1ull << 53);
  }
  // chunks
  s->row_chunk.alloc(nrows + 1);
  {
    DevBuf<int64_t> cnt;
    cnt.alloc(nrows + 1);
    TB_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), st));
    if (nrows)
      k_plan_chunks<false><<<grid_for(nrows, threads, s->sm_count), threads, 0, st>>>(
          s->rowptr.p, s->col.p, nrows, N, dense_fill_factor(), cnt.p, nullptr, nullptr, nullptr);
    exclusive_scan_i64(st, cnt.p, s->row_chunk.p, nrows + 1);
  }
  int64_t nchunks = 0;
  TB_CUDA(cudaMemcpy(&nchunks, s->row_chunk.p + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost));
  s->nchunks = nchunks;
  s->chunk_beg.alloc(nchunks + 1);
  s->chunk_row.alloc(nchunks + 1);
  if (nrows)
    k_plan_chunks<true><<<grid_for(nrows, threads, s->sm_count), threads, 0, st>>>(
        s->rowptr.p, s->col.p, nrows, N, dense_fill_factor(), nullptr, s->row_chunk.p, s->chunk_beg.p,
        s->chunk_row.p);
  // chunk c covers cells [chunk_beg[c], chunk_beg[c+1]): close the last one
  TB_CUDA(cudaMemcpyAsync(s->chunk_beg.p + nchunks, &s->nnz_full, sizeof(int64_t),
                          cudaMemcpyHostToDevice, st));
  // upper-triangle count of the block
  {
    DevBuf<int64_t> us, uc, total;
    us.alloc(nrows + 1); uc.alloc(nrows + 1); total.alloc(1);
    TB_CUDA(cudaMemsetAsync(uc.p, 0, uc.bytes(), st));
    if (nrows)
      k_upper_start<<<grid_for(nrows, threads, s->sm_count), threads, 0, st>>>(
          s->rowptr.p, s->col.p, nrows, s->row_begin, false, us.p, uc.p);
    size_t tmp_bytes = 0;
    TB_CUDA(cub::DeviceReduce::Sum(nullptr, tmp_bytes, uc.p, total.p, nrows + 1, st));
    DevBuf<uint8_t> tmp;
    tmp.alloc(tmp_bytes ? tmp_bytes : 1);
    TB_CUDA(cub::DeviceReduce::Sum(tmp.p, tmp_bytes, uc.p, total.p, nrows + 1, st));
    TB_CUDA(cudaMemcpyAsync(&s->nnz_upper, total.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
  }
  // scratch
  const size_t nc = nchunks ? nchunks : 1, nn = N ? N : 1;
  s->partial_f.alloc(nc);
  TB_CUDA(cudaMemsetAsync(s->partial_f.p, 0, s->partial_f.bytes(), st));   // tile chunks never write theirs
  s->partial_i.alloc(nc);
  s->x.alloc(nn + kXPad); s->b.alloc(nn);
  TB_CUDA(cudaMemsetAsync(s->x.p, 0, s->x.bytes(), st));
  s->s_part.alloc(2 * nn); s->s_full.alloc(2 * nn);
  s->i_part.alloc(2 * nn); s->i_full.alloc(2 * nn);
  s->bad.alloc(nn); s->present.alloc(nn);
  s->block_red.alloc(4 * 1024);
  s->scal_f.alloc(16);
  s->scal_i.alloc(16);
  s->state.alloc(1);
  s->h_state.alloc(1);
  TB_CUDA(cudaMemsetAsync(s->s_part.p, 0, s->s_part.bytes(), st));
  TB_CUDA(cudaMemsetAsync(s->i_part.p, 0, s->i_part.bytes(), st));
  TB_CUDA(cudaStreamSynchronize(st));
  tm.lap("chunks + scratch");
  build_encoded(s);
  tm.lap("encode dense stream");
  build_tiles(s);
  tm.lap("sparse tiles");
  {
    const char *keep = getenv("TB_KEEP_CSR");
    s->csr_resident = true;
    if (s->positive && !(keep && keep[0] == '1')) {
      TB_CUDA(cudaStreamSynchronize(st));
      s->rowptr.release();
      s->col.release();
      s->val.release();
      s->chunk_beg.release();
      s->partial_i.release();
      s->csr_resident = false;
    }
  }
  s->device_bytes = s->t_cells.bytes() + s->t_perm.bytes() + s->t_slice.bytes() + s->t_off.bytes() +
                    s->t_desc.bytes() + s->t_desc_off.bytes() +
                    s->t_partial.bytes() + s->enc_data.bytes() + s->enc_off.bytes() + s->enc_desc.bytes() + s->rowptr.bytes() + s->col.bytes() + s->val.bytes() + s->chunk_beg.bytes() +
                    s->chunk_row.bytes() + s->row_chunk.bytes() + s->partial_f.bytes() +
                    s->partial_i.bytes() + s->x.bytes() + s->b.bytes() +
                    s->s_part.bytes() + s->s_full.bytes() + s->i_part.bytes() + s->i_full.bytes() +
                    s->bad.bytes() + s->present.bytes() + s->chrom_of.bytes() + s->row_nnz.bytes() +
                    s->row_sum.bytes() + s->t_list.bytes() + s->t_unit.bytes() + s->groups.bytes() +
                    s->work_ids.bytes() + s->enc_tovf.bytes();
}

// Block-diagonal concatenation of whole (unsharded) stores on the same device: member k
// becomes section k.  The members' upper-triangle cells are decoded on the device (row-major
// sorted, bins shifted by the member's offset) and the batch store is built from them like any
// other matrix.  No host traffic.
void build_batch(tb_store *s, int32_t n, tb_store *const *members) {
  int64_t cells = 0;
  for (int32_t k = 0; k < n; ++k) cells += members[k]->nnz_upper;
  TB_REQUIRE(cells < 4294967295LL, "batch too large");
  DevBuf<int32_t> row, col, val;
  row.alloc(cells ? cells : 1);
  col.alloc(cells ? cells : 1);
  val.alloc(cells ? cells : 1);
  int64_t bin_off = 0, cell_off = 0;
  for (int32_t k = 0; k < n; ++k) {
    const tb_store *m = members[k];
    export_upper_device(m, row.p + cell_off, col.p + cell_off, val.p + cell_off, (int32_t)bin_off);
    bin_off += m->nbins;
    cell_off += m->nnz_upper;
  }
  build_from_coo(s, cells, row.p, col.p, val.p, TB_MEM_DEVICE);
}

void setup_sections(tb_store *s) {
  const int64_t N = s->nbins;
  std::vector<int32_t> h(N ? N : 1);
  for (int32_t c = 0; c + 1 < (int32_t)s->chrom_offsets.size(); ++c)
    for (int64_t i = s->chrom_offsets[c]; i < s->chrom_offsets[c + 1]; ++i) h[i] = c;
  s->chrom_of.alloc(N ? N : 1);
  if (N) TB_CUDA(cudaMemcpy(s->chrom_of.p, h.data(), N * sizeof(int32_t), cudaMemcpyHostToDevice));
  const size_t nsec = s->chrom_offsets.size() - 1;
  s->chrom_end.alloc(nsec);
  TB_CUDA(cudaMemcpy(s->chrom_end.p, s->chrom_offsets.data() + 1, nsec * sizeof(int64_t), cudaMemcpyHostToDevice));
}

// Cells of the owned rows, row-major sorted, into device arrays: the upper triangle (col >= row,
// nnz_upper entries) or, full_rows, every stored cell of the rows (nnz_full entries).  d_row may
// be null; d_ptr (nrows + 1, may be null) receives the first cell of every row; bin_off is added
// to rows and columns.  Runs on the store's stream and waits for it.
void export_cells_device(const tb_store *s, bool full_rows, int32_t *d_row, int32_t *d_col, int32_t *d_val,
                         int64_t *d_ptr, int32_t bin_off) {
  cudaStream_t st = s->stream;
  const int threads = 256;
  const int64_t nrows = s->nrows, m = full_rows ? s->nnz_full : s->nnz_upper;
  if (m == 0) {
    if (d_ptr) TB_CUDA(cudaMemsetAsync(d_ptr, 0, (nrows + 1) * sizeof(int64_t), st));
    TB_CUDA(cudaStreamSynchronize(st));
    return;
  }
  if (s->csr_resident) {
    DevBuf<int64_t> us, uc, uo;
    us.alloc(nrows + 1); uc.alloc(nrows + 1); uo.alloc(nrows + 1);
    TB_CUDA(cudaMemsetAsync(uc.p, 0, uc.bytes(), st));
    k_upper_start<<<grid_for(nrows, threads, s->sm_count), threads, 0, st>>>(
        s->rowptr.p, s->col.p, nrows, s->row_begin, full_rows, us.p, uc.p);
    exclusive_scan_i64(st, uc.p, uo.p, nrows + 1);
    k_copy_upper<<<grid_for(nrows * 32, threads, s->sm_count), threads, 0, st>>>(
        s->col.p, s->val.p, us.p, uc.p, uo.p, nrows, s->row_begin, bin_off, d_row, d_col, d_val);
    if (d_ptr) TB_CUDA(cudaMemcpyAsync(d_ptr, uo.p, (nrows + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    TB_CUDA(cudaStreamSynchronize(st));
    return;
  }
  // decode the compact stream: count per row, place (any order inside a row), sort by (row, column)
  DevBuf<unsigned long long> cnt;
  DevBuf<int64_t> off;
  cnt.alloc(nrows + 1);
  off.alloc(nrows + 1);
  TB_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), st));
  CountUpperVisitor cv;
  cv.rb = s->row_begin;
  cv.cnt = cnt.p;
  visit_upper(s, cv, full_rows);
  exclusive_scan_i64(st, reinterpret_cast<const int64_t *>(cnt.p), off.p, nrows + 1);
  int64_t total = 0;
  TB_CUDA(cudaMemcpy(&total, off.p + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost));
  TB_REQUIRE(total == m, "internal: compact stream holds %lld cells, expected %lld",
             (long long)total, (long long)m);
  DevBuf<uint64_t> key_in, key_out;
  DevBuf<int32_t> val_in;
  key_in.alloc(m); key_out.alloc(m); val_in.alloc(m);
  TB_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), st));
  PlaceUpperVisitor pv;
  pv.rb = s->row_begin;
  pv.off = off.p;
  pv.cursor = cnt.p;
  pv.key = key_in.p;
  pv.val = val_in.p;
  visit_upper(s, pv, full_rows);
  int row_bits = 1;
  while (((int64_t)1 << row_bits) < nrows) ++row_bits;
  size_t tmp_bytes = 0;
  TB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_in.p, key_out.p, val_in.p, d_val, m, 0,
                                          32 + row_bits, st));
  DevBuf<uint8_t> tmp;
  tmp.alloc(tmp_bytes ? tmp_bytes : 1);
  TB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key_in.p, key_out.p, val_in.p, d_val, m, 0,
                                          32 + row_bits, st));
  k_unpack_keys<<<grid_for(m, threads, s->sm_count), threads, 0, st>>>(key_out.p, m, s->row_begin, bin_off,
                                                                       bin_off, d_row, d_col);
  if (d_ptr) TB_CUDA(cudaMemcpyAsync(d_ptr, off.p, (nrows + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  TB_CUDA(cudaStreamSynchronize(st));
}

void export_upper_device(const tb_store *s, int32_t *d_row, int32_t *d_col, int32_t *d_val, int32_t bin_off) {
  export_cells_device(s, false, d_row, d_col, d_val, nullptr, bin_off);
}

// CSR form on the host: h_ptr[nrows + 1], h_col / h_val [*nnz]; pass null arrays to get *nnz only.
void export_csr(const tb_store *s, int form, int64_t *nnz, int64_t *h_ptr, int32_t *h_col, int32_t *h_val) {
  const bool full = form == TB_CELLS_FULL_ROWS;
  const int64_t m = full ? s->nnz_full : s->nnz_upper;
  *nnz = m;
  if (!h_ptr) return;
  TB_REQUIRE(h_col && h_val, "export: null output buffer");
  cudaStream_t st = s->stream;
  DevBuf<int32_t> ocol, oval;
  DevBuf<int64_t> optr;
  ocol.alloc(m ? m : 1); oval.alloc(m ? m : 1); optr.alloc(s->nrows + 1);
  export_cells_device(s, full, nullptr, ocol.p, oval.p, optr.p, 0);
  TB_CUDA(cudaMemcpyAsync(h_ptr, optr.p, (s->nrows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  if (m) {
    TB_CUDA(cudaMemcpyAsync(h_col, ocol.p, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaMemcpyAsync(h_val, oval.p, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  }
  TB_CUDA(cudaStreamSynchronize(st));
}

void export_upper(const tb_store *s, int64_t *nnz, int32_t *h_row, int32_t *h_col, int32_t *h_val) {
  *nnz = s->nnz_upper;
  if (!h_row) return;
  TB_REQUIRE(h_col && h_val, "export: null output buffer");
  if (s->nnz_upper == 0) return;
  cudaStream_t st = s->stream;
  DevBuf<int32_t> orow, ocol, oval;
  orow.alloc(s->nnz_upper); ocol.alloc(s->nnz_upper); oval.alloc(s->nnz_upper);
  export_upper_device(s, orow.p, ocol.p, oval.p, 0);
  TB_CUDA(cudaMemcpyAsync(h_row, orow.p, orow.bytes(), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(h_col, ocol.p, ocol.bytes(), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(h_val, oval.p, oval.bytes(), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace tb
