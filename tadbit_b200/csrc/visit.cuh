// Cell visitors: run a functor over every stored UPPER-triangle cell (i <= j) of the rows a
// store owns.  The one-pass reductions that need (i, j, count) per cell -- per-distance sums
// (K5), the normalised-value statistics and decay sums (extras.cu), the export of the cells --
// are written once as a Visitor and run over whichever copy of the matrix the store holds:
//
//   compact stream   dense chunks D8/D16/D32 (+ their overflow lists), per-cell chunks S16/S32
//                    (k_visit_chunks) and the two-section sparse tiles (k_visit_tiles).  This is
//                    the only copy a store keeps when every stored value is > 0 (the CSR is
//                    released after the build, build.cu:finish_layout);
//   CSR              k_visit_csr: stores that hold zeros or negative values keep their CSR (an
//                    explicit zero cannot be told from an absent cell in a dense panel).
//
// The stream holds full rows (both triangles); cells left of the diagonal are skipped, chunks
// and tiles that lie entirely left of it are skipped as a whole.
//
// A Visitor is copied to every thread and provides
//     static constexpr int kShared, kThreads;       // shared memory per CTA, threads per CTA
//     __device__ void begin(uint8_t *smem);         // all threads, before the walk
//     __device__ bool block(int64_t i0, int64_t i1, int64_t j0, int64_t j1);
//                                                   // can a cell of rows [i0, i1) x columns [j0, j1)
//                                                   // matter?  false skips a whole panel / tile
//     __device__ bool row(int64_t i);               // once per thread and row before its cells:
//                                                   // per-row state; false skips the row
//     __device__ void cell(int64_t i, int64_t j, int32_t v);
//     __device__ void end(uint8_t *smem);           // all threads, after the walk
#pragma once
#include "common.cuh"

namespace tb {

constexpr int kVisitSlices = kTileRows / 32;
constexpr int kVisitSliceTab = 2 * kVisitSlices + 1;      // tiles.cu: (ones, general) start per slice + end

struct VisitSrc {
  // chunks of the compact stream
  const uint8_t *enc_data;
  const int64_t *enc_off;
  const int4 *enc_desc;
  const int32_t *chunk_row;
  int64_t nchunks;
  // sparse tiles
  const uint32_t *t_cells;
  const int64_t *t_off;
  const uint32_t *t_slice;
  const uint16_t *t_perm;
  int64_t t_nrb, t_nwin;
  // CSR (only when the store kept it)
  const int64_t *chunk_beg;
  const int32_t *col, *val;
  int64_t row_begin, nrows, nbins;
  bool full_rows;                      // also visit the cells left of the diagonal (export of whole rows)
};

inline VisitSrc visit_src(const tb_store *s) {
  VisitSrc v;
  v.enc_data = s->enc_data.p;
  v.enc_off = s->enc_off.p;
  v.enc_desc = s->enc_desc.p;
  v.chunk_row = s->chunk_row.p;
  v.nchunks = s->nchunks;
  v.t_cells = s->t_cells.p;
  v.t_off = s->t_off.p;
  v.t_slice = s->t_slice.p;
  v.t_perm = s->t_perm.p;
  v.t_nrb = s->t_nunits ? s->t_nrb : 0;
  v.t_nwin = s->t_nwin;
  v.chunk_beg = s->chunk_beg.p;
  v.col = s->col.p;
  v.val = s->val.p;
  v.row_begin = s->row_begin;
  v.nrows = s->nrows;
  v.nbins = s->nbins;
  v.full_rows = false;
  return v;
}

// column of the value at element position p of a dense payload (inverse of encode.cu:dense_pos)
__device__ __forceinline__ int32_t dense_slot(int32_t p) {
  const int32_t q = p & 127, ln = q >> 2, k = q & 3;
  return (p & ~127) + (k >> 1) * 64 + ln * 2 + (k & 1);
}

template <typename V>
__global__ void __launch_bounds__(1024) k_visit_chunks(VisitSrc src, V vis) {
  extern __shared__ __align__(16) uint8_t vsm[];
  vis.begin(vsm);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < src.nchunks; c += nwarp) {
    const int4 d = src.enc_desc[c];
    const int enc = d.z & 0xff;
    if (enc == ENC_T) continue;                         // its cells are in the tiles
    const int64_t gi = src.row_begin + src.chunk_row[c];
    const int64_t g = src.full_rows ? -1 : gi;          // cells with column >= g are visited
    const uint8_t *p = src.enc_data + src.enc_off[c];
    if (enc >= ENC_D32 && ((int64_t)d.x + d.y <= g || !vis.block(gi, gi + 1, d.x, (int64_t)d.x + d.y)))
      continue;                                         // panel left of the diagonal, or of no interest
    if (!vis.row(gi)) continue;
    if (enc == ENC_S32) {
      const int2 *w = reinterpret_cast<const int2 *>(p);
      for (int k = lane; k < d.y; k += 32) {
        const int2 e = w[k];
        if (e.x >= g) vis.cell(gi, e.x, e.y);
      }
    } else if (enc == ENC_S16) {
      const uint32_t *w = reinterpret_cast<const uint32_t *>(p);
      for (int k = lane; k < d.y; k += 32) {
        const uint32_t e = w[k];
        const int64_t j = d.x + (int64_t)(e & 0xffffu);
        if (j >= g) vis.cell(gi, j, (int32_t)(e >> 16));
      }
    } else {
      const int c0 = d.x, nslots = d.y, novf = d.z >> 8;
      const int width = enc == ENC_D32 ? 4 : (enc == ENC_D16 ? 2 : 1);
      const int per = 16 / width;                       // values per 16-byte word
      const int nwords = nslots / per;
      const uint4 *w = reinterpret_cast<const uint4 *>(p);
      for (int wi = lane; wi < nwords; wi += 32) {
        const uint4 q = w[wi];
        const uint32_t u[4] = {q.x, q.y, q.z, q.w};
        if ((q.x | q.y | q.z | q.w) == 0u) continue;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          if (e >= per) break;
          int32_t v;
          if (width == 1) v = (int32_t)((u[e >> 2] >> (8 * (e & 3))) & 0xffu);
          else if (width == 2) v = (int32_t)((u[e >> 1] >> (16 * (e & 1))) & 0xffffu);
          else v = (int32_t)u[e];
          if (v == 0) continue;
          const int64_t j = (int64_t)c0 + dense_slot(wi * per + e);
          TB_DCHECK(j < src.nbins && gi < src.nbins);
          if (j >= g) vis.cell(gi, j, v);
        }
      }
      const int2 *ovf = reinterpret_cast<const int2 *>(p + (int64_t)nslots * width);
      for (int k = lane; k < novf; k += 32) {
        const int2 e = ovf[k];
        if (e.x >= g) vis.cell(gi, e.x, e.y);
      }
    }
  }
  __syncthreads();
  vis.end(vsm);
}

// one warp per (tile, slice): lane l walks the cells of sorted row 32 * slice + l
template <typename V>
__global__ void __launch_bounds__(1024) k_visit_tiles(VisitSrc src, V vis) {
  extern __shared__ __align__(16) uint8_t vsm[];
  vis.begin(vsm);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t nitems = src.t_nrb * src.t_nwin * kVisitSlices;
  for (int64_t it = warp; it < nitems; it += nwarp) {
    const int64_t tile = it / kVisitSlices;
    const int sl = (int)(it % kVisitSlices);
    const int64_t rb = tile / src.t_nwin, w = tile % src.t_nwin;
    const int64_t wlo = w * kTileCols, g0 = src.row_begin + rb * kTileRows;
    if (!src.full_rows && wlo + kTileCols <= g0) continue;     // tile left of the diagonal
    if (!vis.block(g0, g0 + kTileRows, wlo, wlo + kTileCols)) continue;
    const uint32_t *st = src.t_slice + tile * kVisitSliceTab + 2 * sl;
    const uint32_t o1 = st[0], o2 = st[1], o3 = st[2];
    if (o1 == o3) continue;
    const int64_t gi = g0 + src.t_perm[tile * kTileRows + sl * 32 + lane];
    const int64_t g = src.full_rows ? -1 : gi;
    const bool mine = gi < src.nbins && vis.row(gi);    // (rows of a partial last tile hold no cells)
    const uint4 *base = reinterpret_cast<const uint4 *>(src.t_cells + src.t_off[tile] * 128);
    TB_DCHECK(o1 <= o2 && o2 <= o3 && (src.t_off[tile] + o3) * 128 <= c_lim.t_cell_words);
    for (uint32_t blk = o1; blk < o2; ++blk) {          // ones: 8 column offsets per lane and block
      if (!mine) break;
      const uint4 q = base[(size_t)blk * 32 + lane];
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t off = (u[e >> 1] >> (16 * (e & 1))) & 0xffffu;
        if (off >= (uint32_t)kTileCols) continue;       // padding points behind the window
        const int64_t j = wlo + off;
        TB_DCHECK(j < src.nbins && gi < src.nbins);
        if (j >= g) vis.cell(gi, j, 1);
      }
    }
    for (uint32_t blk = o2; blk < o3; ++blk) {          // general: 4 cells per lane and block
      if (!mine) break;
      const uint4 q = base[(size_t)blk * 32 + lane];
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int32_t v = (int32_t)(u[e] & 0xffffu);
        if (v == 0) continue;                           // padding
        const int64_t j = wlo + (u[e] >> 18);
        TB_DCHECK(j < src.nbins && gi < src.nbins);
        if (j >= g) vis.cell(gi, j, v);
      }
    }
  }
  __syncthreads();
  vis.end(vsm);
}

template <typename V>
__global__ void __launch_bounds__(1024) k_visit_csr(VisitSrc src, V vis) {
  extern __shared__ __align__(16) uint8_t vsm[];
  vis.begin(vsm);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < src.nchunks; c += nwarp) {
    const int64_t gi = src.row_begin + src.chunk_row[c];
    const int64_t g = src.full_rows ? -1 : gi;
    const int64_t beg = src.chunk_beg[c], end = src.chunk_beg[c + 1];
    if (src.col[end - 1] < g) continue;                 // chunk entirely left of the diagonal
    if (!vis.row(gi)) continue;
    for (int64_t k = beg + lane; k < end; k += 32) {
      const int64_t j = src.col[k];
      if (j >= g) vis.cell(gi, j, src.val[k]);
    }
  }
  __syncthreads();
  vis.end(vsm);
}

// Run `vis` over every stored upper-triangle cell of the store's rows (stream s->stream).
// V::kThreads threads per CTA; as many CTAs as fit an SM with V::kShared bytes of shared memory.
template <typename V, typename K>
int visit_grid(const tb_store *s, K kernel, int64_t warps_wanted) {
  if (V::kShared > 48 * 1024)
    TB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V::kShared));
  int per_sm = 1;
  TB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, V::kThreads, V::kShared));
  if (per_sm < 1) per_sm = 1;
  const int64_t need = (warps_wanted * 32 + V::kThreads - 1) / V::kThreads;
  const int64_t cap = (int64_t)s->sm_count * per_sm;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

template <typename V>
void visit_upper(const tb_store *s, const V &vis, bool full_rows = false) {
  VisitSrc src = visit_src(s);
  src.full_rows = full_rows;
  set_check_limits(s);
  if (s->nchunks == 0) return;
  if (s->csr_resident) {
    k_visit_csr<V><<<visit_grid<V>(s, k_visit_csr<V>, s->nchunks), V::kThreads, V::kShared, s->stream>>>(src, vis);
    return;
  }
  k_visit_chunks<V><<<visit_grid<V>(s, k_visit_chunks<V>, s->nchunks), V::kThreads, V::kShared, s->stream>>>(src, vis);
  if (src.t_nrb) {
    const int64_t items = src.t_nrb * src.t_nwin * kVisitSlices;
    k_visit_tiles<V><<<visit_grid<V>(s, k_visit_tiles<V>, items), V::kThreads, V::kShared, s->stream>>>(src, vis);
  }
}

}  // namespace tb
