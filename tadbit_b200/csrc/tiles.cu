// Sparse tiles: the part of K1 (s_i = sum_j v_ij x_j) for cells outside the dense panels.
//
// Gathering x_j straight from L2 costs a 32-byte sector per 8-byte value and saturates the L2
// (measured on the genome-wide 5 kb matrix: 13 TB/s of sector traffic, 11 % L1 hit rate).
// Here the sparse cells are tiled kTileRows x kTileCols: a CTA owns a row block, walks its
// non-empty tiles and keeps the tile's x window (kTileCols = 16384 doubles, 128 KB) in shared
// memory -- fetched by cp.async.bulk while the threads already load their tile metadata and
// first cells -- so every gather is a shared-memory read.  (A wide single buffer beats two
// narrow ones here: far from the diagonal a tile holds only a few cells per row, and the fixed
// cost per tile is what matters.)
//
// Inside a tile the rows are sorted by their number of cells (per-tile permutation) and stored
// as sliced ELL: 32 consecutive sorted rows form a slice, cell k of lane l sits at
// slice_base + 32 k + l (one coalesced 128-byte load per warp and k), padded to the slice's
// longest row; sorting keeps the padding small even where a row has 0-10 cells per tile.
// A tile has 64 slices for 32 warps: warp w takes slice w and slice 63-w (longest with
// shortest), which evens out the work between the warps that meet at the tile's barrier.
// A cell is 32 bits: column offset in the window (14) | count (18).  A thread accumulates its
// row of the tile in a register and adds it to the row block's accumulators in shared memory
// (conflict-free: the permutation maps threads to distinct rows).  Row blocks are cut into work
// units of roughly equal cell count; a unit writes kTileRows partial sums, added per row in
// unit order by the combine step -> deterministic, no atomics in the pass loop.
#include <algorithm>
#include <cub/cub.cuh>

#include "common.cuh"

namespace tb {

namespace {

constexpr int kTileThreads = kTileRows / 2;
constexpr int kSlices = kTileRows / 32;
constexpr int kSliceTab = kSlices + 1;
constexpr int kRowBits = 11;                               // kTileRows = 2048
constexpr int kWinBytes = kTileCols * 8;
constexpr int kTileSmem = kWinBytes + kTileRows * 8 + 64;
constexpr uint32_t kColMask = (1u << kTileColBits) - 1u;
static_assert(kTileRows == 1 << kRowBits, "kRowBits");

__device__ __forceinline__ double u2d(uint32_t v) {
  return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const void *p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// ---- build ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_tile_count(const int32_t *__restrict__ col, const int64_t *__restrict__ chunk_beg,
             const int32_t *__restrict__ chunk_row, const int4 *__restrict__ desc, int64_t nchunks,
             int64_t nwin, int32_t *__restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < nchunks; c += nwarp) {
    if ((desc[c].z & 0xff) != ENC_T) continue;
    const int64_t beg = chunk_beg[c], end = chunk_beg[c + 1], row = chunk_row[c];
    for (int64_t k = beg + lane; k < end; k += 32)
      atomicAdd(&cnt[row * nwin + col[k] / kTileCols], 1);
  }
}

// one CTA per tile: sort the tile's rows by cell count (descending, ties by row), write the
// permutation, its inverse, the slice offsets and the padded size of the tile
__global__ void __launch_bounds__(kTileThreads)
k_tile_sort(const int32_t *__restrict__ cnt, int64_t nwin, uint16_t *__restrict__ perm,
            uint16_t *__restrict__ inv, uint32_t *__restrict__ slice, int64_t *__restrict__ tile_size,
            unsigned long long *__restrict__ totals) {
  __shared__ uint32_t key[kTileRows];
  const int t = threadIdx.x;
  const int64_t tile = blockIdx.x, rb = tile / nwin, w = tile % nwin;
  unsigned long long real = 0;
  for (int e = t; e < kTileRows; e += kTileThreads) {
    const uint32_t mine = (uint32_t)cnt[(rb * kTileRows + e) * nwin + w];
    key[e] = (mine << kRowBits) | (uint32_t)(kTileRows - 1 - e);
    real += mine;
  }
  __syncthreads();
  for (int k = 2; k <= kTileRows; k <<= 1)            // bitonic sort, descending
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int e = t; e < kTileRows; e += kTileThreads) {
        const int o = e ^ j;
        if (o > e) {
          const uint32_t a = key[e], b = key[o];
          const bool desc_dir = (e & k) == 0;
          if (desc_dir ? a < b : a > b) { key[e] = b; key[o] = a; }
        }
      }
      __syncthreads();
    }
  for (int e = t; e < kTileRows; e += kTileThreads) {
    const int row = kTileRows - 1 - (int)(key[e] & (uint32_t)(kTileRows - 1));
    perm[tile * kTileRows + e] = (uint16_t)row;
    inv[(rb * kTileRows + row) * nwin + w] = (uint16_t)e;
  }
  if (t == 0) {                                        // slice s is as long as its first row
    uint32_t run = 0;
    for (int s = 0; s < kSlices; ++s) {
      slice[tile * kSliceTab + s] = run;
      run += (key[s * 32] >> kRowBits) * 32u;
    }
    slice[tile * kSliceTab + kSlices] = run;
    tile_size[tile] = run;
  }
  for (int o = 16; o; o >>= 1) real += __shfl_xor_sync(0xffffffffu, real, o);
  if ((t & 31) == 0 && real) atomicAdd(totals, real);  // unpadded cells, for the layout report
}

// one thread per row: for every tile chunk, the number of tile cells of the same row that sit
// in the chunk's FIRST window but in earlier chunks (a window spans several panels, so dense
// chunks can split a row's sparse cells of one window over several chunks)
__global__ void k_tile_rank_base(const int32_t *__restrict__ col, const int64_t *__restrict__ chunk_beg,
                                 const int64_t *__restrict__ row_chunk, const int4 *__restrict__ desc,
                                 int64_t nrows, int32_t *__restrict__ base) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x) {
    int64_t cur_w = -1, cur_n = 0;
    for (int64_t c = row_chunk[r]; c < row_chunk[r + 1]; ++c) {
      if ((desc[c].z & 0xff) != ENC_T) continue;
      const int64_t beg = chunk_beg[c], end = chunk_beg[c + 1];
      const int64_t wf = col[beg] / kTileCols, wl = col[end - 1] / kTileCols;
      base[c] = wf == cur_w ? (int32_t)cur_n : 0;
      int64_t lo = beg, hi = end;                      // first cell of the chunk in window wl
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (col[mid] < wl * kTileCols) lo = mid + 1; else hi = mid;
      }
      cur_n = (wl == cur_w ? cur_n : 0) + (end - lo);
      cur_w = wl;
    }
  }
}

__global__ void __launch_bounds__(256)
k_tile_fill(const int32_t *__restrict__ col, const int32_t *__restrict__ val,
            const int64_t *__restrict__ chunk_beg, const int32_t *__restrict__ chunk_row,
            const int4 *__restrict__ desc, const int32_t *__restrict__ rank_base, int64_t nchunks,
            int64_t nwin, const uint16_t *__restrict__ inv, const uint32_t *__restrict__ slice,
            const int64_t *__restrict__ t_off, uint32_t *__restrict__ cells) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < nchunks; c += nwarp) {
    if ((desc[c].z & 0xff) != ENC_T) continue;
    const int64_t beg = chunk_beg[c], end = chunk_beg[c + 1], row = chunk_row[c];
    const int64_t rb = row / kTileRows;
    const int64_t wfirst = col[beg] / kTileCols;
    const int32_t base = rank_base[c];
    for (int64_t k = beg + lane; k < end; k += 32) {
      const int32_t cc = col[k];
      const int64_t w = cc / kTileCols;
      int64_t lo = beg, hi = k;                        // first cell of this chunk in window w
      const int32_t wstart = (int32_t)(w * kTileCols);
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (col[mid] < wstart) lo = mid + 1; else hi = mid;
      }
      const int64_t rank = k - lo + (w == wfirst ? base : 0), tile = rb * nwin + w;
      const int t = inv[row * nwin + w];
      cells[t_off[tile] + slice[tile * kSliceTab + (t >> 5)] + rank * 32 + (t & 31)] =
          (uint32_t)(cc - wstart) | ((uint32_t)val[k] << kTileColBits);
    }
  }
}

// ---- the pass kernel --------------------------------------------------------------------------
__device__ __forceinline__ double cell_term(uint32_t c, const double *xw) {
  return u2d(c >> kTileColBits) * xw[c & kColMask];
}

// A warp's cells of one unit form a stream that does not depend on the x windows: slice a of
// tile 0, slice b of tile 0, slice a of tile 1, ...  The warp pulls that stream through a private
// ring in shared memory with cp.async.bulk (pieces of up to kPiece k-steps = 1 KB), the producer
// cursor running kRing pieces ahead of the consumer cursor across slice and tile boundaries, so
// HBM latency is covered by bytes in flight in the copy engine instead of registers.
constexpr int kPiece = 8;                                  // k-steps (128 B each) per ring stage
constexpr int kRing = 2;                                   // stages per warp
constexpr int kRingBytes = kRing * kPiece * 128;           // 2 KB per warp
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kTileSmemRing = kWinBytes + kTileRows * 8 + kTileWarps * kRingBytes + kTileWarps * kRing * 8 + 64;

struct SliceCur {            // a warp's position in its cell stream
  int j, which, k, len;
  const uint32_t *p;
};

__device__ __forceinline__ void cur_load(SliceCur &c, const int32_t *__restrict__ list, int i0, int n,
                                         const uint32_t *__restrict__ slice,
                                         const int64_t *__restrict__ t_off,
                                         const uint32_t *__restrict__ cells, int sa, int sb, int part,
                                         int parts) {
  // move to the next non-empty slice at or after (j, which); len = 0 and j = n when exhausted
  while (c.j < n) {
    const int64_t tile = list[i0 + c.j];
    const int s = c.which ? sb : sa;
    const uint32_t s0 = slice[tile * kSliceTab + s], s1 = slice[tile * kSliceTab + s + 1];
    const int full = (int)((s1 - s0) >> 5);
    const int kb = full * part / parts;
    c.len = full * (part + 1) / parts - kb;
    c.k = 0;
    c.p = cells + t_off[tile] + s0 + kb * 32;
    if (c.len > 0) return;
    if (c.which) { c.which = 0; ++c.j; } else c.which = 1;
  }
  c.len = 0;
}

__device__ __forceinline__ void cur_advance(SliceCur &c, int steps, const int32_t *__restrict__ list,
                                            int i0, int n, const uint32_t *__restrict__ slice,
                                            const int64_t *__restrict__ t_off,
                                            const uint32_t *__restrict__ cells, int sa, int sb, int part,
                                            int parts) {
  c.k += steps;
  if (c.k < c.len) return;
  if (c.which) { c.which = 0; ++c.j; } else c.which = 1;
  cur_load(c, list, i0, n, slice, t_off, cells, sa, sb, part, parts);
}

__global__ void __launch_bounds__(kTileThreads, 1)
k_rowsum_tiles(const uint32_t *__restrict__ cells, const int64_t *__restrict__ t_off,
               const uint16_t *__restrict__ perm, const uint32_t *__restrict__ slice,
               const int32_t *__restrict__ list, const int4 *__restrict__ unit, int64_t nunits,
               int64_t nwin, const double *__restrict__ x, double *__restrict__ partial,
               const IceState *__restrict__ state) {
  if (state && state->stopped) return;
  extern __shared__ __align__(128) uint8_t smem[];
  double *xw = reinterpret_cast<double *>(smem);                       // [kTileCols]
  double *acc = reinterpret_cast<double *>(smem + kWinBytes);          // [kTileRows]
  const int t = threadIdx.x, lane = t & 31, wic = t >> 5;
  uint8_t *ring = smem + kWinBytes + kTileRows * 8 + wic * kRingBytes;
  const uint32_t ring_u32 = smem_u32(ring);
  const uint32_t bars = smem_u32(smem + kWinBytes + kTileRows * 8 + kTileWarps * kRingBytes);
  const uint32_t wbar = bars + kTileWarps * kRing * 8;                 // window barrier
  const uint32_t rbar = bars + wic * kRing * 8;                        // this warp's ring barriers
  const int sa = wic, sb = kSlices - 1 - wic;          // longest slices paired with shortest
  if (t == 0) mbar_init(wbar, 1);
  if (lane == 0) {
    for (int q = 0; q < kRing; ++q) mbar_init(rbar + 8 * q, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint32_t wparity = 0, rparity = 0;                    // rparity bit q = parity of ring stage q
  for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x) {
    const int4 un = unit[u];
    const int i0 = un.y, n = un.z - un.y;
    const int part = un.w >> 8, parts = un.w & 0xff;   // a heavy tile is split over several units
    acc[t] = 0.0;
    acc[t + kTileThreads] = 0.0;
    // producer: runs ahead over this warp's stream of the whole unit
    SliceCur pc;
    pc.j = 0; pc.which = 0;
    cur_load(pc, list, i0, n, slice, t_off, cells, sa, sb, part, parts);
    auto issue = [&](int stage) {
      if (pc.len == 0) return;                          // stream exhausted
      const int steps = min(kPiece, pc.len - pc.k);
      if (lane == 0) {
        mbar_expect_tx(rbar + 8 * stage, steps * 128);
        bulk_g2s(ring_u32 + stage * kPiece * 128, pc.p + pc.k * 32, steps * 128, rbar + 8 * stage);
      }
      cur_advance(pc, steps, list, i0, n, slice, t_off, cells, sa, sb, part, parts);
    };
#pragma unroll
    for (int q = 0; q < kRing; ++q) issue(q);
    int stage = 0;
    for (int j = 0; j < n; ++j) {
      const int64_t tile = list[i0 + j];
      __syncthreads();                       // the window of the previous tile is consumed
      if (t == 0) {                          // 128 KB window, four 32 KB bulk copies
        const double *src = x + (tile % nwin) * kTileCols;
        mbar_expect_tx(wbar, kWinBytes);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          bulk_g2s(smem_u32(xw) + q * (kWinBytes / 4), src + q * (kTileCols / 4), kWinBytes / 4, wbar);
      }
      const uint32_t *st = slice + tile * kSliceTab;
      int len[2], row[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int s = h ? sb : sa;
        const int full = (int)((st[s + 1] - st[s]) >> 5);
        len[h] = full * (part + 1) / parts - full * part / parts;
        row[h] = perm[tile * kTileRows + s * 32 + lane];
      }
      mbar_wait(wbar, wparity);
      wparity ^= 1u;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double a0 = 0.0, a1 = 0.0;
        for (int k = 0; k < len[h]; k += kPiece) {
          const int steps = min(kPiece, len[h] - k);
          mbar_wait(rbar + 8 * stage, (rparity >> stage) & 1u);
          rparity ^= 1u << stage;
          const uint32_t *rp = reinterpret_cast<const uint32_t *>(ring + stage * kPiece * 128) + lane;
          if (steps == kPiece) {
#pragma unroll
            for (int q = 0; q < kPiece; q += 2) {
              a0 += cell_term(rp[q * 32], xw);
              a1 += cell_term(rp[(q + 1) * 32], xw);
            }
          } else {
            for (int q = 0; q < steps; ++q) a0 += cell_term(rp[q * 32], xw);
          }
          __syncwarp();                      // every lane has read this stage
          issue(stage);
          stage = (stage + 1 == kRing) ? 0 : stage + 1;
        }
        if (len[h]) acc[row[h]] += a0 + a1;  // distinct rows per thread: no conflict
      }
    }
    __syncthreads();
    partial[u * kTileRows + t] = acc[t];
    partial[u * kTileRows + t + kTileThreads] = acc[t + kTileThreads];
    __syncthreads();                          // acc is re-zeroed at the top of the next unit
  }
}

inline int grid_1d(int64_t n, int threads, int sm_count, int per_sm) {
  int64_t g = (n + threads - 1) / threads, cap = (int64_t)sm_count * per_sm;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace

void build_tiles(tb_store *s) {
  cudaStream_t st = s->stream;
  s->t_nrb = s->t_nwin = s->t_ncells = s->t_npadded = s->t_nunits = 0;
  if (s->enc_chunks[ENC_T] == 0) return;
  const int64_t nrb = (s->nrows + kTileRows - 1) / kTileRows;
  const int64_t nwin = (s->nbins + kTileCols - 1) / kTileCols;
  const int64_t ntiles = nrb * nwin;
  TB_REQUIRE(ntiles < 2147483647LL, "too many sparse tiles");
  s->t_nrb = nrb;
  s->t_nwin = nwin;
  DevBuf<int32_t> cnt;
  DevBuf<uint16_t> inv;
  DevBuf<int64_t> tile_size;
  DevBuf<unsigned long long> totals;
  cnt.alloc(nrb * kTileRows * nwin);
  inv.alloc(nrb * kTileRows * nwin);
  tile_size.alloc(ntiles + 1);
  totals.alloc(1);
  TB_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), st));
  TB_CUDA(cudaMemsetAsync(tile_size.p, 0, tile_size.bytes(), st));
  TB_CUDA(cudaMemsetAsync(totals.p, 0, totals.bytes(), st));
  const int cg = grid_1d(s->nchunks * 32, 256, s->sm_count, 8);
  k_tile_count<<<cg, 256, 0, st>>>(s->col.p, s->chunk_beg.p, s->chunk_row.p, s->enc_desc.p,
                                   s->nchunks, nwin, cnt.p);
  s->t_perm.alloc(ntiles * kTileRows);
  s->t_slice.alloc(ntiles * kSliceTab);
  s->t_off.alloc(ntiles + 1);
  k_tile_sort<<<(unsigned)ntiles, kTileThreads, 0, st>>>(cnt.p, nwin, s->t_perm.p, inv.p,
                                                        s->t_slice.p, tile_size.p, totals.p);
  {
    size_t tmp_bytes = 0;
    TB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, tile_size.p, s->t_off.p, ntiles + 1, st));
    DevBuf<uint8_t> tmp;
    tmp.alloc(tmp_bytes ? tmp_bytes : 1);
    TB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, tile_size.p, s->t_off.p, ntiles + 1, st));
    TB_CUDA(cudaStreamSynchronize(st));
  }
  std::vector<int64_t> h_size(ntiles);
  unsigned long long h_total = 0;
  TB_CUDA(cudaMemcpy(h_size.data(), tile_size.p, ntiles * sizeof(int64_t), cudaMemcpyDeviceToHost));
  TB_CUDA(cudaMemcpy(&h_total, totals.p, sizeof(h_total), cudaMemcpyDeviceToHost));
  TB_CUDA(cudaMemcpy(&s->t_npadded, s->t_off.p + ntiles, sizeof(int64_t), cudaMemcpyDeviceToHost));
  s->t_ncells = (int64_t)h_total;
  cnt.release();
  s->t_cells.alloc(s->t_npadded ? s->t_npadded : 1);
  TB_CUDA(cudaMemsetAsync(s->t_cells.p, 0, s->t_cells.bytes(), st));       // padding = (col 0, count 0)
  DevBuf<int32_t> rank_base;
  rank_base.alloc(s->nchunks + 1);
  TB_CUDA(cudaMemsetAsync(rank_base.p, 0, rank_base.bytes(), st));
  k_tile_rank_base<<<grid_1d(s->nrows, 256, s->sm_count, 16), 256, 0, st>>>(
      s->col.p, s->chunk_beg.p, s->row_chunk.p, s->enc_desc.p, s->nrows, rank_base.p);
  k_tile_fill<<<cg, 256, 0, st>>>(s->col.p, s->val.p, s->chunk_beg.p, s->chunk_row.p, s->enc_desc.p,
                                  rank_base.p, s->nchunks, nwin, inv.p, s->t_slice.p, s->t_off.p,
                                  s->t_cells.p);
  // Work units: each row block's non-empty tiles, cut into runs of about equal stored cells so
  // that there are ~16 units per SM whatever the number of row blocks (1 GPU or a 1/8 shard).
  // A tile heavier than that (the ones on the diagonal) is split: `parts` units each take a
  // share of every slice of the tile.  unit = (row block, list begin, list end, part<<8|parts).
  const double target = (double)s->t_npadded / (16.0 * s->sm_count) + 1.0;
  std::vector<int32_t> list, units, unit_of_rb(nrb + 1, 0);
  for (int64_t rb = 0; rb < nrb; ++rb) {
    unit_of_rb[rb] = (int32_t)(units.size() / 4);
    int64_t run = 0;
    int32_t begin = (int32_t)list.size();
    auto flush = [&]() {
      if ((int32_t)list.size() > begin)
        units.insert(units.end(), {(int32_t)rb, begin, (int32_t)list.size(), 1});
      begin = (int32_t)list.size();
      run = 0;
    };
    for (int64_t w = 0; w < nwin; ++w) {
      const int64_t sz = h_size[rb * nwin + w];
      if (!sz) continue;
      if ((double)sz > 1.5 * target) {
        flush();
        int parts = (int)((double)sz / target + 0.999);
        if (parts > 64) parts = 64;
        list.push_back((int32_t)(rb * nwin + w));
        for (int p = 0; p < parts; ++p)
          units.insert(units.end(), {(int32_t)rb, begin, begin + 1, (p << 8) | parts});
        begin = (int32_t)list.size();
        continue;
      }
      list.push_back((int32_t)(rb * nwin + w));
      run += sz;
      if ((double)run >= target) flush();
    }
    flush();
  }
  unit_of_rb[nrb] = (int32_t)(units.size() / 4);
  s->t_nunits = (int64_t)units.size() / 4;
  s->t_list.alloc(list.size() ? list.size() : 1);
  s->t_unit.alloc(units.size() ? units.size() : 4);
  s->t_unit_of_rb.alloc(nrb + 1);
  s->t_partial.alloc(s->t_nunits ? s->t_nunits * kTileRows : 1);
  if (!list.empty())
    TB_CUDA(cudaMemcpy(s->t_list.p, list.data(), list.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (!units.empty())
    TB_CUDA(cudaMemcpy(s->t_unit.p, units.data(), units.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  TB_CUDA(cudaMemcpy(s->t_unit_of_rb.p, unit_of_rb.data(), (nrb + 1) * sizeof(int32_t),
                     cudaMemcpyHostToDevice));
  TB_CUDA(cudaMemsetAsync(s->t_partial.p, 0, s->t_partial.bytes(), st));
  TB_CUDA(cudaStreamSynchronize(st));
}

void launch_rowsum_tiles(tb_store *s, const double *x, const IceState *state) {
  if (s->t_nunits == 0) return;
  static bool attr_set[64] = {false};            // per device: the attribute is per context
  if (!attr_set[s->device & 63]) {
    TB_CUDA(cudaFuncSetAttribute(k_rowsum_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemRing));
    attr_set[s->device & 63] = true;
  }
  const int grid = (int)(s->t_nunits < s->sm_count ? s->t_nunits : s->sm_count);
  k_rowsum_tiles<<<grid, kTileThreads, kTileSmemRing, s->stream>>>(
      s->t_cells.p, s->t_off.p, s->t_perm.p, s->t_slice.p, s->t_list.p,
      reinterpret_cast<const int4 *>(s->t_unit.p), s->t_nunits, s->t_nwin, x, s->t_partial.p, state);
}

}  // namespace tb
