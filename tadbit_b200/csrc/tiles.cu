// Sparse tiles: the part of K1 (s_i = sum_j v_ij x_j) for cells outside the dense panels.
//
// Gathering x_j straight from L2 costs a 32-byte sector per 8-byte value and saturates the L2
// (measured on the genome-wide 5 kb matrix: 13 TB/s of sector traffic, 11 % L1 hit rate).
// Here the sparse cells are tiled kTileRows x kTileCols: a CTA owns a row block, walks its
// non-empty tiles and keeps the tile's x window (kTileCols = 16384 doubles, 128 KB) in shared
// memory, fetched by cp.async.bulk, so every gather is a shared-memory read.
//
// Layout.  Inside a tile the rows are sorted by length (per-tile permutation) and stored as
// sliced ELL: 32 consecutive sorted rows form a slice.  At high resolution most cells outside
// the dense panels hold exactly ONE contact (genome-wide 5 kb model: 96 % of the tile cells),
// and a cell worth 1 needs no count: every slice therefore has two sections,
//
//   ones     16 bits per cell = column offset in the window; the term is x_j itself (one DADD
//            instead of unpack + convert + DFMA),
//   general  32 bits per cell = (column offset * 4) << 16 | count, count < 2^15: `cell >> 15`
//            is the byte offset of x_j in the window and `cell & 0xffff` the count,
//
// both stored in 512-byte blocks (one 16-byte shared-memory read per lane: 8 ones or 4 general
// cells of that lane), each section padded to the longest row of the slice.  Padding slots are
// written for every lane and step by k_tile_init: a ones slot points at one of 16 zeros kept
// behind the x window (index 16384 + r), a general slot has count 0; either way at the bank
// the slot prefers, r = (lane + step) mod 16.  Rows are sorted by their number of ones.
//
// Bank conflicts.  ncu showed the first version of this kernel waiting on the shared-memory
// pipe (mio_throttle): a random 8-byte gather by 32 lanes costs ~6 wavefronts instead of 2.
// The order of a row's cells inside a tile is free, so the fill kernel deals them out by bank:
// at step k lane l prefers a cell whose column is congruent to l + k mod 16 (the 16 lanes of a
// half-warp then hit 16 different 8-byte bank pairs); cells that find no such step take the
// steps left over.
//
// Streaming.  A warp's cells of one work unit form a stream that does not depend on the x
// windows.  It is cut at build time into pieces of up to 4 blocks (2 KB), one 8-byte copy
// descriptor each (where the piece starts, which rows it belongs to, end-of-slice / end-of-tile
// marks); the warp pulls the pieces through a private two-stage ring in shared memory with
// cp.async.bulk, two pieces ahead of what it consumes, across slice and tile boundaries.  A CTA
// has 16 warps; warp w takes slices w, 31 - w, 32 + w and 63 - w of every tile (long with
// short).
//
// Scheduling.  Row blocks are cut into work units of roughly equal cost; CTAs (one per SM)
// fetch them from an atomic counter, most expensive first.  A unit writes kTileRows partial
// sums, added per row in unit order by the combine step, so results do not depend on which CTA
// ran which unit: deterministic, no atomics on the data path.
#include "common.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <numeric>
#include <type_traits>
#include <cub/cub.cuh>

#include "common.cuh"

namespace tb {

namespace {

#ifndef TB_TILE_WARPS
#define TB_TILE_WARPS 16
#endif
#ifndef TB_FILL_HIST
#define TB_FILL_HIST 1      // k_tile_fill counts residues in a shared-memory histogram per warp (match_any +
#endif                      // one add per group); 0 = 32 ballots per 32 cells (measured on the genome-wide
                            // 5 kb matrix: tile build 103 ms vs 165 ms, same layout)
constexpr int kSortThreads = kTileRows / 2 < 1024 ? kTileRows / 2 : 1024;   // k_tile_sort
constexpr int kTileWarps = TB_TILE_WARPS;                  // warps of a k_rowsum_tiles CTA
constexpr int kTileThreads = kTileWarps * 32;
constexpr int kPieceBlocks = 4;                            // 512-byte blocks per ring stage
constexpr int kSlices = kTileRows / 32;
constexpr int kSliceTab = 2 * kSlices + 1;                 // (ones start, general start) per slice + end
constexpr int kRowBits = kTileRows == 4096 ? 12 : 11;      // kTileRows = 2048 | 4096
constexpr int kWinBytes = kTileCols * 8;
constexpr int kZeroTail = 128;                             // 16 zeros behind the window (ones padding)
constexpr int kWarpSlices = kSlices / kTileWarps;          // slices of a tile one warp streams
static_assert(kTileRows == 1 << kRowBits, "kRowBits");
static_assert(kSlices % kTileWarps == 0, "tile kernel shape");
static_assert(kTileCols * 4 <= 65536 && kTileValBits == 15 && kTileCols + 16 <= 32768, "cell packing");

// copy descriptor: x = first 512-byte block of the piece in the cell array,
// y = row group (tile * 64 + slice, 25 bits) | general section << 25 | blocks << 26 |
//     end-of-slice << 30 | end-of-tile << 31
constexpr uint32_t kDescGroupMask = (1u << 25) - 1u;
constexpr uint32_t kDescGeneral = 1u << 25;
constexpr uint32_t kDescEndSlice = 1u << 30, kDescEndTile = 1u << 31;
constexpr int kPieceBytes = kPieceBlocks * 512;
constexpr int kRing = 2;
constexpr int kAccOff = kWinBytes + kZeroTail;
constexpr int kRingOff = kAccOff + kTileRows * 8;
constexpr int kTailOff = kRingOff + kTileWarps * kRing * kPieceBytes;
constexpr int kTileSmem = kTailOff + (kTileWarps * kRing + 1) * 8 + 64;
// fixed cost of a tile in a unit (window load, barrier, descriptor chain), in general cells
constexpr int64_t kTileFixedCost = 24000;
constexpr double kBlockCost = 128.0;                       // a block, in general cells

__device__ __forceinline__ double u2d(uint32_t v) {
  return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}
__device__ __forceinline__ uint32_t pack_cell(int col_off, int v) {
  return ((uint32_t)col_off << 18) | (uint32_t)v;
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
// h-th slice of warp w: blocks of kTileWarps slices, walked back and forth (long with short)
__host__ __device__ __forceinline__ int warp_slice(int w, int h) {
  return h * kTileWarps + ((h & 1) ? kTileWarps - 1 - w : w);
}

// The tiles hold every cell of the ENC_T chunks, plus the cells of D8 chunks that are too large
// for a byte but fit a tile cell (encode.cu counts them per chunk in tovf).
__device__ __forceinline__ bool tile_overflow(int32_t v) { return v > 255 && v < (1 << kTileValBits); }

// cnt[row][window]: ones in the low half, general cells in the high half
__global__ void __launch_bounds__(256)
k_tile_count(const int32_t *__restrict__ col, const int32_t *__restrict__ val,
             const int64_t *__restrict__ chunk_beg, const int32_t *__restrict__ chunk_row,
             const int4 *__restrict__ desc, const uint16_t *__restrict__ tovf, int64_t nchunks,
             int64_t nwin, int32_t *__restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < nchunks; c += nwarp) {
    const bool all = (desc[c].z & 0xff) == ENC_T;
    if (!all && tovf[c] == 0) continue;
    const int64_t beg = chunk_beg[c], end = chunk_beg[c + 1], row = chunk_row[c];
    for (int64_t k = beg + lane; k < end; k += 32) {
      const int32_t v = val[k];
      if (all || tile_overflow(v))
        atomicAdd(&cnt[row * nwin + col[k] / kTileCols], (all && v == 1) ? 1 : (1 << 16));
    }
  }
}

// one CTA per tile: sort the tile's rows by their number of ones (descending, ties by row), write
// the permutation, its inverse, the block offsets of the slice sections and the size of the tile
__global__ void __launch_bounds__(kSortThreads)
k_tile_sort(const int32_t *__restrict__ cnt, int64_t nwin, uint16_t *__restrict__ perm,
            uint16_t *__restrict__ inv, uint32_t *__restrict__ slice, int64_t *__restrict__ tile_size,
            int32_t *__restrict__ tile_crit, unsigned long long *__restrict__ totals) {
  __shared__ uint32_t key[kTileRows];
  __shared__ uint16_t gen[kTileRows];                  // general cells of each row
  __shared__ uint32_t b1[kSlices], b2[kSlices];        // blocks of the two sections of each slice
  const int t = threadIdx.x;
  const int64_t tile = blockIdx.x, rb = tile / nwin, w = tile % nwin;
  unsigned long long real = 0;
  for (int e = t; e < kTileRows; e += kSortThreads) {
    const uint32_t both = (uint32_t)cnt[(rb * kTileRows + e) * nwin + w];
    const uint32_t ones = both & 0xffffu;
    gen[e] = (uint16_t)(both >> 16);
    key[e] = (ones << kRowBits) | (uint32_t)(kTileRows - 1 - e);
    real += ones + (both >> 16);
  }
  __syncthreads();
  for (int k = 2; k <= kTileRows; k <<= 1)            // bitonic sort, descending
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int e = t; e < kTileRows; e += kSortThreads) {
        const int o = e ^ j;
        if (o > e) {
          const uint32_t a = key[e], b = key[o];
          const bool desc_dir = (e & k) == 0;
          if (desc_dir ? a < b : a > b) { key[e] = b; key[o] = a; }
        }
      }
      __syncthreads();
    }
  for (int e = t; e < kTileRows; e += kSortThreads) {
    const int row = kTileRows - 1 - (int)(key[e] & (uint32_t)(kTileRows - 1));
    perm[tile * kTileRows + e] = (uint16_t)row;
    inv[(rb * kTileRows + row) * nwin + w] = (uint16_t)e;
  }
  if (t < kSlices) {          // sections are as long as the longest row of the slice
    uint32_t gmax = 0;
    for (int i = 0; i < 32; ++i) {
      const int row = kTileRows - 1 - (int)(key[t * 32 + i] & (uint32_t)(kTileRows - 1));
      gmax = max(gmax, (uint32_t)gen[row]);
    }
    b1[t] = ((key[t * 32] >> kRowBits) + 7u) >> 3;     // 8 ones per lane and block
    b2[t] = (gmax + 3u) >> 2;                          // 4 general cells per lane and block
  }
  __syncthreads();
  if (t == 0) {
    uint32_t run = 0, n1 = 0, n2 = 0;
    for (int s = 0; s < kSlices; ++s) {
      slice[tile * kSliceTab + 2 * s] = run;
      run += b1[s];
      slice[tile * kSliceTab + 2 * s + 1] = run;
      run += b2[s];
      n1 += b1[s];
      n2 += b2[s];
    }
    slice[tile * kSliceTab + 2 * kSlices] = run;
    tile_size[tile] = run;
    uint32_t crit = 0;                                 // blocks of the busiest warp
    for (int wq = 0; wq < kTileWarps; ++wq) {
      uint32_t mine = 0;
      for (int h = 0; h < kWarpSlices; ++h) mine += b1[warp_slice(wq, h)] + b2[warp_slice(wq, h)];
      crit = max(crit, mine);
    }
    tile_crit[tile] = (int32_t)crit;
    if (n1) atomicAdd(totals + 1, (unsigned long long)n1);
    if (n2) atomicAdd(totals + 2, (unsigned long long)n2);
  }
  for (int o = 16; o; o >>= 1) real += __shfl_xor_sync(0xffffffffu, real, o);
  if ((t & 31) == 0 && real) atomicAdd(totals, real);  // unpadded cells, for the layout report
}

// one CTA per tile: every slot starts as padding at its preferred bank, r = (lane + step) mod 16
__global__ void __launch_bounds__(256)
k_tile_init(const uint32_t *__restrict__ slice, const int64_t *__restrict__ t_off,
            uint32_t *__restrict__ cells) {
  const int64_t tile = blockIdx.x;
  const uint32_t *st = slice + tile * kSliceTab;
  uint32_t *base = cells + t_off[tile] * 128;          // 128 words per block
  const uint32_t nblocks = st[2 * kSlices];
  for (int s = 0; s < kSlices; ++s) {
    const uint32_t o1 = st[2 * s], o2 = st[2 * s + 1], o3 = st[2 * s + 2];
    // ones section: word i of a block holds steps 2 (i & 3), 2 (i & 3) + 1 of lane i >> 2
    for (uint32_t i = threadIdx.x; i < (o2 - o1) * 128; i += blockDim.x) {
      const uint32_t blk = i >> 7, wi = i & 127, lane = wi >> 2, k = blk * 8 + (wi & 3) * 2;
      const uint32_t lo = 16384u + ((lane + k) & 15u), hi = 16384u + ((lane + k + 1) & 15u);
      base[(size_t)o1 * 128 + i] = lo | (hi << 16);
    }
    // general section: word i of a block holds step i & 3 of lane i >> 2
    for (uint32_t i = threadIdx.x; i < (o3 - o2) * 128; i += blockDim.x) {
      const uint32_t blk = i >> 7, wi = i & 127, lane = wi >> 2, k = blk * 4 + (wi & 3);
      base[(size_t)o2 * 128 + i] = pack_cell((int)((lane + k) & 15u), 0);
    }
  }
  (void)nblocks;
}

// A warp's position in the tile cells of one row: the row's chunks that hold tile cells, in
// column order.  `filter`: the current chunk is a D8 chunk, only its tile_overflow cells count.
struct RowCursor {
  int64_t c, cend, k, kend;
  bool filter;
  const int64_t *chunk_beg;
  const int4 *desc;
  const uint16_t *tovf;
  __device__ __forceinline__ void seek() {       // to the next chunk with tile cells at or after c
    while (c < cend && (desc[c].z & 0xff) != ENC_T && tovf[c] == 0) ++c;
    if (c < cend) { k = chunk_beg[c]; kend = chunk_beg[c + 1]; filter = (desc[c].z & 0xff) != ENC_T; }
  }
  // after a block of `na` (<= 32) cells of the current window was taken: true = the window goes on
  __device__ __forceinline__ bool advance(int na, const int32_t *__restrict__ col, int32_t whi) {
    k += na;
    if (k < kend) return na == 32;               // fewer than 32: the next cell is in a later window
    ++c;
    seek();
    return c < cend && col[k] < whi;
  }
};

// per-residue slot tables of one (row, tile) segment: lane r holds residue r of the ones
// section, lane 16 + r residue r of the general section
struct SlotTables {
  int served, freebase, fre;
};
// slot of the t-th cell of a section (base lane 0 or 16) that found no step of its own residue
__device__ __forceinline__ void leftover_slot(const SlotTables &tb_, int base, int t, bool want, int &r2,
                                              int &m2) {
#if TB_FILL_HIST
  if (!__any_sync(0xffffffffu, want)) return;          // every cell of the block found its own bank
#endif
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int fb = __shfl_sync(0xffffffffu, tb_.freebase, base + q);
    const int fr = __shfl_sync(0xffffffffu, tb_.fre, base + q);
    const int sv = __shfl_sync(0xffffffffu, tb_.served, base + q);
    if (want && t >= fb && t < fb + fr) { r2 = q; m2 = sv + t - fb; }
  }
}

// one warp per row: deal the row's cells of every tile out over the steps of its lane
__global__ void __launch_bounds__(256)
k_tile_fill(const int32_t *__restrict__ col, const int32_t *__restrict__ val,
            const int64_t *__restrict__ chunk_beg, const int64_t *__restrict__ row_chunk,
            const int4 *__restrict__ desc, const uint16_t *__restrict__ tovf, int64_t nrows,
            int64_t nwin, const uint16_t *__restrict__ inv, const uint32_t *__restrict__ slice,
            const int64_t *__restrict__ t_off, uint32_t *__restrict__ cells) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
#if TB_FILL_HIST
  __shared__ int s_hist[8][32];                        // per warp: cells per key / cells placed per key
  int *hist = s_hist[threadIdx.x >> 5];
#endif
  for (int64_t r = warp; r < nrows; r += nwarp) {
    const int64_t rb = r / kTileRows;
    RowCursor cur;
    cur.c = row_chunk[r]; cur.cend = row_chunk[r + 1]; cur.k = cur.kend = 0;
    cur.filter = false;
    cur.chunk_beg = chunk_beg; cur.desc = desc; cur.tovf = tovf;
    cur.seek();
    while (cur.c < cur.cend) {
      const int32_t w = col[cur.k] / kTileCols, wlo = w * kTileCols, whi = wlo + kTileCols;
      const int64_t tile = rb * nwin + w;
      const int p = inv[r * nwin + w], l = p & 31, sl = p >> 5;
      const uint32_t *st = slice + tile * kSliceTab + 2 * sl;
      const uint32_t o1 = st[0], o2 = st[1], o3 = st[2];
      const int L1 = (int)(o2 - o1) * 8, L2 = (int)(o3 - o2) * 4;
      uint32_t *tbase = cells + t_off[tile] * 128;
      uint16_t *out1 = reinterpret_cast<uint16_t *>(tbase + (size_t)o1 * 128) + l * 8;
      uint32_t *out2 = tbase + (size_t)o2 * 128 + l * 4;
      const RowCursor start = cur;
      // sweep 1: cells per (section, residue): lane q counts key q = 16 * general + residue
      int b = 0;
#if TB_FILL_HIST
      hist[lane] = 0;
      __syncwarp();
#endif
      for (;;) {
        const int64_t idx = cur.k + lane;
        const int32_t cc = idx < cur.kend ? col[idx] : 0x7fffffff;
        const bool inwin = cc < whi;
        const int32_t vv = inwin ? val[idx] : 0;
        const bool act = inwin && (!cur.filter || tile_overflow(vv));
        const int key = (cc & 15) + ((cur.filter || vv != 1) ? 16 : 0);
        const int na = __popc(__ballot_sync(0xffffffffu, inwin));
#if TB_FILL_HIST
        // lanes with the same key form a group; its first lane adds the group to the histogram
        const unsigned grp = __match_any_sync(0xffffffffu, act ? key : 32 + lane);
        if (act && (grp & lt) == 0u) hist[key] += __popc(grp);
        __syncwarp();
#else
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const unsigned m = __ballot_sync(0xffffffffu, act && key == q);
          if (lane == q) b += __popc(m);
        }
#endif
        if (!cur.advance(na, col, whi)) break;
      }
#if TB_FILL_HIST
      b = hist[lane];
      __syncwarp();
      hist[lane] = 0;                        // reused as the cells placed per key in sweep 2
      __syncwarp();
#endif
      // steps of residue q for this lane: k = ((q - l) & 15) + 16 m, m < req
      const int L = lane < 16 ? L1 : L2;
      const int d = ((lane & 15) - l) & 15;
      const int req = d < L ? (L - 1 - d) / 16 + 1 : 0;
      SlotTables tabs;
      tabs.served = min(b, req);
      const int left = b - tabs.served;
      tabs.fre = req - tabs.served;
      int li = left, fi = tabs.fre;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, li, o), f = __shfl_up_sync(0xffffffffu, fi, o);
        if (lane >= o) { li += a; fi += f; }
      }
      // the scans run over both sections: take the first section's totals off the second
      const int l15 = __shfl_sync(0xffffffffu, li, 15), f15 = __shfl_sync(0xffffffffu, fi, 15);
      if (lane >= 16) { li -= l15; fi -= f15; }
      const int leftbase = li - left;
      tabs.freebase = fi - tabs.fre;
      // sweep 2: place the cells
      cur = start;
#if !TB_FILL_HIST
      int run = 0;
#endif
      for (;;) {
        const int64_t idx = cur.k + lane;
        const int32_t cc = idx < cur.kend ? col[idx] : 0x7fffffff;
        const bool inwin = cc < whi;
        const int32_t vv = inwin ? val[idx] : 0;
        const bool act = inwin && (!cur.filter || tile_overflow(vv));
        const int general = (cur.filter || vv != 1) ? 16 : 0;
        const int key = (cc & 15) + general;
        const int na = __popc(__ballot_sync(0xffffffffu, inwin));
        const unsigned same = __match_any_sync(0xffffffffu, act ? key : 32 + lane);
#if TB_FILL_HIST
        const int m = hist[key] + __popc(same & lt);
        __syncwarp();                        // every lane has read the counts of this block
        if (act && (same & lt) == 0u) hist[key] += __popc(same);
        __syncwarp();
#else
        const int m = __shfl_sync(0xffffffffu, run, key) + __popc(same & lt);
#endif
        const int sv = __shfl_sync(0xffffffffu, tabs.served, key);
        const int lb = __shfl_sync(0xffffffffu, leftbase, key);
        int r2 = key & 15, m2 = m;
        leftover_slot(tabs, general, lb + m - sv, act && m >= sv, r2, m2);
        if (act) {
          const int ks = ((r2 - l) & 15) + 16 * m2;
          if (general) out2[(ks >> 2) * 128 + (ks & 3)] = pack_cell(cc - wlo, vv);
          else out1[(ks >> 3) * 256 + (ks & 7)] = (uint16_t)(cc - wlo);
        }
#if !TB_FILL_HIST
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const unsigned mq = __ballot_sync(0xffffffffu, act && key == q);
          if (lane == q) run += __popc(mq);
        }
#endif
        if (!cur.advance(na, col, whi)) break;
      }
    }
  }
}

// blocks [kb, kb + len) of a section that belong to part `part` of `parts`
__device__ __forceinline__ int part_range(uint32_t nblocks, int part, int parts, int &kb) {
  kb = (int)((int64_t)nblocks * part / parts);
  return (int)((int64_t)nblocks * (part + 1) / parts) - kb;
}

// one thread per (unit, warp): count / write the copy descriptors of the warp's cell stream
template <bool FILL>
__global__ void __launch_bounds__(256)
k_tile_desc(const int4 *__restrict__ unit, const int32_t *__restrict__ list,
            const uint32_t *__restrict__ slice, const int64_t *__restrict__ t_off, int64_t nunits,
            int64_t *__restrict__ count, const int64_t *__restrict__ desc_off, uint2 *__restrict__ desc) {
  const int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (id >= nunits * kTileWarps) return;
  const int w = (int)(id % kTileWarps);
  const int4 un = unit[id / kTileWarps];
  const int part = un.w >> 8, parts = un.w & 0xff;
  uint2 *o = FILL ? desc + desc_off[id] : nullptr;
  int64_t n_out = 0;
  for (int j = un.y; j < un.z; ++j) {
    const int64_t tile = list[j];
    const uint32_t *st = slice + tile * kSliceTab;
    // the warp's sections of this tile, in stream order: ones then general of each of its slices
    int kb[2 * kWarpSlices], len[2 * kWarpSlices], last = -1;
#pragma unroll
    for (int h = 0; h < 2 * kWarpSlices; ++h) {
      const int sl = warp_slice(w, h >> 1);
      len[h] = part_range(st[2 * sl + (h & 1) + 1] - st[2 * sl + (h & 1)], part, parts, kb[h]);
      if (len[h]) last = h;
    }
    if (last < 0) {
      if (FILL) o[n_out] = make_uint2(0u, (uint32_t)(tile * kSlices + w) | kDescEndTile);
      ++n_out;
      continue;
    }
#pragma unroll
    for (int h = 0; h < 2 * kWarpSlices; ++h) {
      const int sl = warp_slice(w, h >> 1);
      const int pieces = (len[h] + kPieceBlocks - 1) / kPieceBlocks;
      if (FILL) {
        const int64_t first = t_off[tile] + st[2 * sl + (h & 1)] + kb[h];
        // the slice ends with this section unless its general section (h odd) still has blocks
        const bool ends_slice = (h & 1) || len[h + 1] == 0;
        for (int i = 0; i < pieces; ++i) {
          const int nb = min(kPieceBlocks, len[h] - kPieceBlocks * i);
          uint32_t y = (uint32_t)(tile * kSlices + sl) | ((h & 1) ? kDescGeneral : 0u) | ((uint32_t)nb << 26);
          if (i == pieces - 1) {
            if (ends_slice) y |= kDescEndSlice;
            if (h == last) y |= kDescEndTile;
          }
          o[n_out + i] = make_uint2((uint32_t)(first + kPieceBlocks * i), y);
        }
      }
      n_out += pieces;
    }
  }
  if (!FILL) count[id] = n_out;
}

// ---- the pass kernel --------------------------------------------------------------------------
__device__ __forceinline__ double cell_term(uint32_t c, const uint8_t *xw) {
  TB_DCHECK((c >> 15) + 8u <= (uint32_t)kWinBytes && ((c >> 15) & 7u) == 0u);
  return u2d(c & 0xffffu) * *reinterpret_cast<const double *>(xw + (c >> 15));
}
__device__ __forceinline__ double x_at(uint32_t byte_off, const uint8_t *xw) {
  TB_DCHECK(byte_off + 8u <= (uint32_t)(kWinBytes + kZeroTail));
  return *reinterpret_cast<const double *>(xw + byte_off);
}

// general section: 4 cells per lane and block
template <int NB>
__device__ __forceinline__ void consume_general(const uint8_t *stage, int lane, const uint8_t *xw, double &a0,
                                                double &a1) {
  const uint4 *rp = reinterpret_cast<const uint4 *>(stage) + lane;
  uint4 q[NB];
#pragma unroll
  for (int g = 0; g < NB; ++g) q[g] = rp[g * 32];
#pragma unroll
  for (int g = 0; g < NB; ++g) {
    a0 += cell_term(q[g].x, xw);
    a1 += cell_term(q[g].y, xw);
    a0 += cell_term(q[g].z, xw);
    a1 += cell_term(q[g].w, xw);
  }
}

// ones section: 8 cells per lane and block, the term is x itself
template <int NB>
__device__ __forceinline__ void consume_ones(const uint8_t *stage, int lane, const uint8_t *xw, double &a0,
                                             double &a1) {
  const uint4 *rp = reinterpret_cast<const uint4 *>(stage) + lane;
  uint4 q[NB];
#pragma unroll
  for (int g = 0; g < NB; ++g) q[g] = rp[g * 32];
#pragma unroll
  for (int g = 0; g < NB; ++g) {
    const uint32_t wv[4] = {q[g].x, q[g].y, q[g].z, q[g].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a0 += x_at((wv[i] << 3) & 0x7fff8u, xw);
      a1 += x_at((wv[i] >> 13) & 0x7fff8u, xw);
    }
  }
}

template <bool GENERAL>
__device__ __forceinline__ void consume_piece(uint32_t nb, const uint8_t *rs, int lane, const uint8_t *xw,
                                              double &a0, double &a1) {
  if (nb == kPieceBlocks) {
    if (GENERAL) consume_general<kPieceBlocks>(rs, lane, xw, a0, a1);
    else consume_ones<kPieceBlocks>(rs, lane, xw, a0, a1);
  } else {                                             // a short last piece: pairs, then one
    uint32_t g = 0;
    for (; g + 2 <= nb; g += 2) {
      if (GENERAL) consume_general<2>(rs + g * 512, lane, xw, a0, a1);
      else consume_ones<2>(rs + g * 512, lane, xw, a0, a1);
    }
    if (g < nb) {
      if (GENERAL) consume_general<1>(rs + g * 512, lane, xw, a0, a1);
      else consume_ones<1>(rs + g * 512, lane, xw, a0, a1);
    }
  }
}

__global__ void __launch_bounds__(kTileThreads, 1)
k_rowsum_tiles(const uint32_t *__restrict__ cells, const uint2 *__restrict__ desc,
               const int64_t *__restrict__ desc_off, const uint16_t *__restrict__ perm,
               const int32_t *__restrict__ list, const int4 *__restrict__ unit,
               const int32_t *__restrict__ order, unsigned int nunits, int64_t nwin,
               const double *__restrict__ x, double *__restrict__ partial,
               unsigned int *__restrict__ counter, int cur, const IceState *__restrict__ state,
               unsigned long long *__restrict__ prof) {
  // the next launch fetches from the other counter
  if (blockIdx.x == 0 && threadIdx.x == 0) counter[cur ^ 1] = 0u;
  if (state && state->stopped) return;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *xw = smem;                                   // [kTileCols] doubles + 16 zeros
  double *acc = reinterpret_cast<double *>(smem + kAccOff);            // [kTileRows]
  const int t = threadIdx.x, lane = t & 31, wic = t >> 5;
  uint8_t *ring = smem + kRingOff + wic * (kRing * kPieceBytes);
  const uint32_t ring_u32 = smem_u32(ring);
  uint8_t *tail = smem + kTailOff;
  const uint32_t bars = smem_u32(tail);
  const uint32_t wbar = bars + kTileWarps * kRing * 8;                 // window barrier
  const uint32_t rbar = bars + wic * kRing * 8;                        // this warp's ring barriers
  volatile unsigned int *s_next = reinterpret_cast<volatile unsigned int *>(tail + (kTileWarps * kRing + 1) * 8);
  if (t < kZeroTail / 8) reinterpret_cast<double *>(smem + kWinBytes)[t] = 0.0;
  if (t == 0) mbar_init(wbar, 1);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < kRing; ++q) mbar_init(rbar + 8 * q, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  uint32_t wparity = 0, rparity = 0;                    // rparity bit q = parity of ring stage q
  // TB_TIMING probe (tb_probe_rowsum): per CTA first / last timestamp, units and tiles it processed
  unsigned long long p_units = 0, p_tiles = 0;
  if (prof && t == 0) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    prof[4 * blockIdx.x] = now;
  }
  for (;;) {
    if (t == 0) *s_next = atomicAdd(&counter[cur], 1u);
    __syncthreads();
    const unsigned int ui = *s_next;
    if (ui >= nunits) break;
    ++p_units;
    const int u = order[ui];
    TB_DCHECK(u >= 0 && u < c_lim.t_nunits);
    const int4 un = unit[u];
    TB_DCHECK(un.y >= 0 && un.y <= un.z && un.z <= c_lim.t_nlist);
    const int i0 = un.y, n = un.z - un.y;
#pragma unroll
    for (int e = t; e < kTileRows; e += kTileThreads) acc[e] = 0.0;
    const int64_t dbeg = desc_off[(int64_t)u * kTileWarps + wic];
    const int nd = (int)(desc_off[(int64_t)u * kTileWarps + wic + 1] - dbeg);
    const uint2 *dp = desc + dbeg;
    TB_DCHECK(nd >= 1 && dbeg >= 0 && dbeg + nd <= c_lim.t_ndesc);
    // start the copy of piece `e` into `stage`; fetch the rows of its slice if it ends one
    auto issue = [&](int stage, const uint2 &e, uint32_t &row) {
      const uint32_t nb = (e.y >> 26) & 15u;
      TB_DCHECK(nb <= (uint32_t)kPieceBlocks && ((long long)e.x + nb) * 128 <= c_lim.t_cell_words);
      TB_DCHECK((long long)(e.y & kDescGroupMask) < c_lim.t_ntiles * kSlices);
      if (nb && lane == 0) {
        mbar_expect_tx(rbar + 8 * stage, nb * 512u);
        bulk_g2s(ring_u32 + stage * kPieceBytes, cells + (size_t)e.x * 128, nb * 512u, rbar + 8 * stage);
      }
      if (e.y & kDescEndSlice) row = perm[(size_t)(e.y & kDescGroupMask) * 32 + lane];
    };
    uint2 e0 = dp[0], e1 = make_uint2(0u, 0u);
    uint32_t row0 = 0, row1 = 0;
    issue(0, e0, row0);
    if (nd > 1) { e1 = dp[1]; issue(1, e1, row1); }
    uint2 pf = nd > 2 ? dp[2] : make_uint2(0u, 0u);     // descriptor two ahead of the consumer
    int ci = 0, stage = 0;
    p_tiles += n;
    for (int j = 0; j < n; ++j) {
      __syncthreads();                       // the window of the previous tile is consumed
      if (t == 0) {                          // 128 KB window, four 32 KB bulk copies
        TB_DCHECK(list[i0 + j] >= 0 && list[i0 + j] < c_lim.t_ntiles);
        const double *src = x + (list[i0 + j] % nwin) * kTileCols;
        TB_DCHECK((list[i0 + j] % nwin) * (long long)kTileCols + kTileCols <= c_lim.nbins + kXPad);
        mbar_expect_tx(wbar, kWinBytes);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          bulk_g2s(smem_u32(xw) + q * (kWinBytes / 4), src + q * (kTileCols / 4), kWinBytes / 4, wbar);
      }
      mbar_wait(wbar, wparity);
      wparity ^= 1u;
      double a0 = 0.0, a1 = 0.0;
      // one piece of the stream; the ring stage is a compile-time constant (two copies of the code)
      auto step = [&](auto stage_c) -> bool {
        constexpr int ST = decltype(stage_c)::value;
        uint2 &e = ST ? e1 : e0;
        uint32_t &erow = ST ? row1 : row0;
        const uint32_t meta = e.y, row = erow;
        const uint32_t nb = (meta >> 26) & 15u;
        if (nb) {
          mbar_wait(rbar + 8 * ST, (rparity >> ST) & 1u);
          rparity ^= 1u << ST;
          const uint8_t *rs = ring + ST * kPieceBytes;
          if (meta & kDescGeneral) consume_piece<true>(nb, rs, lane, xw, a0, a1);
          else consume_piece<false>(nb, rs, lane, xw, a0, a1);
        }
        __syncwarp();                        // every lane has read this stage
        if (meta & kDescEndSlice) {          // a row belongs to one slice: no other warp adds to it
          TB_DCHECK(row < (uint32_t)kTileRows);
          acc[row] += a0 + a1;
          a0 = a1 = 0.0;
        }
        if (ci + 2 < nd) {
          e = pf;
          issue(ST, e, erow);
          if (ci + 3 < nd) pf = dp[ci + 3];
        }
        ++ci;
        return (meta & kDescEndTile) != 0u;
      };
      bool done = false;
      if (stage) { done = step(std::integral_constant<int, 1>()); stage = 0; }
      while (!done) {
        done = step(std::integral_constant<int, 0>());
        stage = 1;
        if (done) break;
        done = step(std::integral_constant<int, 1>());
        stage = 0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = t; e < kTileRows; e += kTileThreads) partial[(int64_t)u * kTileRows + e] = acc[e];
    __syncthreads();                          // acc is re-zeroed, s_next rewritten for the next unit
  }
  if (prof && t == 0) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    prof[4 * blockIdx.x + 1] = now;
    prof[4 * blockIdx.x + 2] = p_units;
    prof[4 * blockIdx.x + 3] = p_tiles;
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
  s->t_launches = 0;
  if (s->enc_chunks[ENC_T] == 0 && s->n_overflow_tiled == 0) return;
  const int64_t nrb = (s->nrows + kTileRows - 1) / kTileRows;
  const int64_t nwin = (s->nbins + kTileCols - 1) / kTileCols;
  const int64_t ntiles = nrb * nwin;
  TB_REQUIRE(ntiles < (1LL << 19), "too many sparse tiles");
  s->t_nrb = nrb;
  s->t_nwin = nwin;
  DevBuf<int32_t> cnt;
  DevBuf<uint16_t> inv;
  DevBuf<int64_t> tile_size;
  DevBuf<int32_t> tile_crit;
  DevBuf<unsigned long long> totals;       // [0] real cells, [1] ones blocks, [2] general blocks
  cnt.alloc(nrb * kTileRows * nwin);
  inv.alloc(nrb * kTileRows * nwin);
  tile_size.alloc(ntiles + 1);
  tile_crit.alloc(ntiles + 1);
  totals.alloc(4);
  TB_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), st));
  TB_CUDA(cudaMemsetAsync(tile_size.p, 0, tile_size.bytes(), st));
  TB_CUDA(cudaMemsetAsync(totals.p, 0, totals.bytes(), st));
  const int cg = grid_1d(s->nchunks * 32, 256, s->sm_count, 8);
  k_tile_count<<<cg, 256, 0, st>>>(s->col.p, s->val.p, s->chunk_beg.p, s->chunk_row.p, s->enc_desc.p,
                                   s->enc_tovf.p, s->nchunks, nwin, cnt.p);
  s->t_perm.alloc(ntiles * kTileRows);
  s->t_slice.alloc(ntiles * kSliceTab);
  s->t_off.alloc(ntiles + 1);
  k_tile_sort<<<(unsigned)ntiles, kSortThreads, 0, st>>>(cnt.p, nwin, s->t_perm.p, inv.p,
                                                        s->t_slice.p, tile_size.p, tile_crit.p, totals.p);
  {
    size_t tmp_bytes = 0;
    TB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, tile_size.p, s->t_off.p, ntiles + 1, st));
    DevBuf<uint8_t> tmp;
    tmp.alloc(tmp_bytes ? tmp_bytes : 1);
    TB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, tile_size.p, s->t_off.p, ntiles + 1, st));
    TB_CUDA(cudaStreamSynchronize(st));
  }
  std::vector<int64_t> h_size(ntiles);     // blocks per tile
  std::vector<int32_t> h_crit(ntiles);
  unsigned long long h_totals[4] = {0, 0, 0, 0};
  int64_t nblocks = 0;
  TB_CUDA(cudaMemcpy(h_size.data(), tile_size.p, ntiles * sizeof(int64_t), cudaMemcpyDeviceToHost));
  TB_CUDA(cudaMemcpy(h_crit.data(), tile_crit.p, ntiles * sizeof(int32_t), cudaMemcpyDeviceToHost));
  TB_CUDA(cudaMemcpy(h_totals, totals.p, sizeof(h_totals), cudaMemcpyDeviceToHost));
  TB_CUDA(cudaMemcpy(&nblocks, s->t_off.p + ntiles, sizeof(int64_t), cudaMemcpyDeviceToHost));
  s->t_ncells = (int64_t)h_totals[0];
  s->t_npadded = nblocks * 128;            // stored stream in 4-byte units (bytes / 4)
  cnt.release();
  s->t_cells.alloc(nblocks ? nblocks * 128 : 1);
  k_tile_init<<<(unsigned)ntiles, 256, 0, st>>>(s->t_slice.p, s->t_off.p, s->t_cells.p);
  k_tile_fill<<<grid_1d(s->nrows * 32, 256, s->sm_count, 8), 256, 0, st>>>(
      s->col.p, s->val.p, s->chunk_beg.p, s->row_chunk.p, s->enc_desc.p, s->enc_tovf.p, s->nrows, nwin, inv.p,
      s->t_slice.p, s->t_off.p, s->t_cells.p);
  // Work units: each row block's non-empty tiles, cut into runs of about equal cost so that there
  // are ~16 units per SM whatever the number of row blocks (1 GPU or a 1/8 shard).  A tile heavier
  // than that (the ones on the diagonal) is split: `parts` units each take a share of every
  // section of the tile.  unit = (row block, list begin, list end, part<<8|parts).
  // The cost of a tile is its blocks, or what its busiest warp streams if the rows are so
  // uneven that this warp, not the SM, sets the time: a lone warp moves about 1/(0.7 * 32) of
  // what the SM does (two pieces in flight).
  std::vector<double> h_eff(ntiles);
  int64_t nonempty = 0;
  double total_cost = 0.0;
  for (int64_t q = 0; q < ntiles; ++q) {
    h_eff[q] = kBlockCost * std::max((double)h_size[q], 0.7 * kTileWarps * 2.0 * (double)h_crit[q]);
    if (h_size[q]) { ++nonempty; total_cost += h_eff[q] + (double)kTileFixedCost; }
  }
  // ... but no smaller than a few times the fixed cost of a unit, unless that would leave SMs
  // without work (small or mostly dense matrices: one unit per SM then)
  // Units per SM: 16 on a whole genome-wide matrix (measured best for the row-sum kernel).  A row
  // block's units each leave a partial per row that the pass kernel (k_ice_exchange) has to add up,
  // so a small shard -- few row blocks -- gets fewer units: the row-sum kernel does not care (a
  // 1/8 shard measured the same with 4, 6, 10 and 16 units per SM), the per-row chain of the
  // exchange kernel shrinks from ~60 to ~16 additions.
  double units_per_sm = std::min(16.0, std::max(4.0, 8.0 * (double)nrb / s->sm_count));
  if (const char *e = getenv("TB_TILE_UNITS_PER_SM")) units_per_sm = std::max(1.0, atof(e));
  const double target = std::max(total_cost / (units_per_sm * s->sm_count),
                                 std::min(3.0 * (double)kTileFixedCost, total_cost / s->sm_count)) + 1.0;
  std::vector<int32_t> list, units, unit_of_rb(nrb + 1, 0);
  std::vector<double> cost;
  // Units are fetched most expensive first; the work of the last row blocks is cut finer so that the
  // CTAs finish together (TB_TILE_TAIL="share,divisor": the last `share` of the cost in units of
  // target / divisor)
  double tail_share = 0.0, tail_div = 1.0, seen_cost = 0.0;
  if (const char *e = getenv("TB_TILE_TAIL")) sscanf(e, "%lf,%lf", &tail_share, &tail_div);
  if (tail_div < 1.0) tail_div = 1.0;
  const double target_coarse = target;
  for (int64_t rb = 0; rb < nrb; ++rb) {
    unit_of_rb[rb] = (int32_t)(units.size() / 4);
    double run = 0;
    int32_t begin = (int32_t)list.size();
    auto flush = [&]() {
      if ((int32_t)list.size() > begin) {
        units.insert(units.end(), {(int32_t)rb, begin, (int32_t)list.size(), 1});
        cost.push_back(run);
      }
      begin = (int32_t)list.size();
      run = 0;
    };
    for (int64_t w = 0; w < nwin; ++w) {
      if (!h_size[rb * nwin + w]) continue;
      const double sz = h_eff[rb * nwin + w];
      const double target = seen_cost >= (1.0 - tail_share) * total_cost ? target_coarse / tail_div : target_coarse;
      seen_cost += sz + (double)kTileFixedCost;
      if (sz > 1.5 * target) {
        flush();
        int parts = (int)(sz / target + 0.999);
        if (parts > 128) parts = 128;
        list.push_back((int32_t)(rb * nwin + w));
        for (int p = 0; p < parts; ++p) {
          units.insert(units.end(), {(int32_t)rb, begin, begin + 1, (p << 8) | parts});
          cost.push_back(sz / parts + (double)kTileFixedCost);
        }
        begin = (int32_t)list.size();
        continue;
      }
      list.push_back((int32_t)(rb * nwin + w));
      run += sz + (double)kTileFixedCost;
      if (run >= target) flush();
    }
    flush();
  }
  unit_of_rb[nrb] = (int32_t)(units.size() / 4);
  s->t_nunits = (int64_t)units.size() / 4;
  std::vector<int32_t> order(s->t_nunits);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return cost[a] > cost[b]; });
  s->t_list.alloc(list.size() ? list.size() : 1);
  s->t_unit.alloc(units.size() ? units.size() : 4);
  s->t_order.alloc(s->t_nunits ? s->t_nunits : 1);
  s->t_unit_of_rb.alloc(nrb + 1);
  s->t_partial.alloc(s->t_nunits ? s->t_nunits * kTileRows : 1);
  s->t_counter.alloc(2);
  if (!list.empty())
    TB_CUDA(cudaMemcpy(s->t_list.p, list.data(), list.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (!units.empty()) {
    TB_CUDA(cudaMemcpy(s->t_unit.p, units.data(), units.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    TB_CUDA(cudaMemcpy(s->t_order.p, order.data(), order.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  TB_CUDA(cudaMemcpy(s->t_unit_of_rb.p, unit_of_rb.data(), (nrb + 1) * sizeof(int32_t),
                     cudaMemcpyHostToDevice));
  TB_CUDA(cudaMemsetAsync(s->t_partial.p, 0, s->t_partial.bytes(), st));
  TB_CUDA(cudaMemsetAsync(s->t_counter.p, 0, s->t_counter.bytes(), st));
  // copy descriptors of every (unit, warp) stream
  const int64_t nstreams = s->t_nunits * kTileWarps;
  s->t_desc_off.alloc(nstreams + 1);
  if (nstreams) {
    DevBuf<int64_t> dcount;
    dcount.alloc(nstreams + 1);
    TB_CUDA(cudaMemsetAsync(dcount.p, 0, dcount.bytes(), st));
    const unsigned dg = (unsigned)((nstreams + 255) / 256);
    const int4 *d_unit = reinterpret_cast<const int4 *>(s->t_unit.p);
    k_tile_desc<false><<<dg, 256, 0, st>>>(d_unit, s->t_list.p, s->t_slice.p, s->t_off.p, s->t_nunits,
                                           dcount.p, nullptr, nullptr);
    size_t tmp_bytes = 0;
    TB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, dcount.p, s->t_desc_off.p, nstreams + 1, st));
    DevBuf<uint8_t> tmp;
    tmp.alloc(tmp_bytes ? tmp_bytes : 1);
    TB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, dcount.p, s->t_desc_off.p, nstreams + 1, st));
    int64_t ndesc = 0;
    TB_CUDA(cudaMemcpyAsync(&ndesc, s->t_desc_off.p + nstreams, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    s->t_desc.alloc(ndesc ? ndesc : 1);
    k_tile_desc<true><<<dg, 256, 0, st>>>(d_unit, s->t_list.p, s->t_slice.p, s->t_off.p, s->t_nunits,
                                          nullptr, s->t_desc_off.p, s->t_desc.p);
  }
  TB_CUDA(cudaStreamSynchronize(st));
}

void launch_rowsum_tiles(tb_store *s, const double *x, const IceState *state, unsigned long long *prof) {
  if (s->t_nunits == 0) return;
  static bool attr_set[64] = {false};            // per device: the attribute is per context
  if (!attr_set[s->device & 63]) {
    TB_CUDA(cudaFuncSetAttribute(k_rowsum_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmem));
    attr_set[s->device & 63] = true;
  }
  const int grid = (int)(s->t_nunits < s->sm_count ? s->t_nunits : s->sm_count);
  const int cur = s->t_launches & 1;
  ++s->t_launches;
  set_check_limits(s);
  k_rowsum_tiles<<<grid, kTileThreads, kTileSmem, s->stream>>>(
      s->t_cells.p, s->t_desc.p, s->t_desc_off.p, s->t_perm.p, s->t_list.p,
      reinterpret_cast<const int4 *>(s->t_unit.p), s->t_order.p, (unsigned int)s->t_nunits, s->t_nwin, x,
      s->t_partial.p, s->t_counter.p, cur, state, prof);
}

}  // namespace tb
