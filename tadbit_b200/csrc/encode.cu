// Compact row-sum stream.
//
// The ICE pass re-reads the matrix every iteration, so every chunk (build.cu:
// k_plan_chunks) is re-encoded into the smallest of these layouts:
//
//   S32  (col:int32, val:int32) per cell                      8 B / cell
//   S16  (col - c0 : u16, val : u16) per cell                 4 B / cell   span, values < 65536
//   D32  one int32 per COLUMN of the chunk's panel             4 B / slot   zeros explicit
//   D16  one u16 per column  + overflow cells                  2 B / slot   (+ 8 B per cell >= 65536)
//   D8   one u8 per column   + overflow cells                  1 B / slot   (+ 8 B per cell >= 256)
//
// Dense chunks cover their whole kPanel-aligned column panel (padded to 1024 slots), so the
// same panel of different rows has the same shape: the row-sum kernel walks kRows such chunks
// at once and reads x a single time for all of them.  A cell too large for the chunk's width
// is written as 0 in the dense part and appended to the chunk's overflow list (col, val); a
// handful of huge cells next to the diagonal then no longer force a whole panel to 32 bits.
// Inside a 128-slot group the slots are permuted so that lane l owns {2l, 2l+1, 64+2l,
// 64+2l+1}: its four values are one 4/8/16-byte word and the matching x entries are two
// aligned 16-byte loads, contiguous across the warp.
//
// Work order = payload order: sparse chunks first (row order), then dense chunks sorted by
// (encoding, panel) with rows ascending inside -- each kernel streams its payload front to
// back.  The CSR stays resident for the one-pass kernels (filter statistics, diagonal sums,
// export).
#include <stdlib.h>

#include <algorithm>
#include <cub/cub.cuh>

#include "common.cuh"

namespace tb {

namespace {

__device__ __forceinline__ int64_t align16(int64_t b) { return (b + 15) & ~(int64_t)15; }

// position (in elements) of slot s inside the permuted dense layout
__device__ __forceinline__ int64_t dense_pos(int64_t s) {
  const int64_t g = s >> 7;
  const int r = (int)(s & 127);
  const int half = r >> 6, rr = r & 63;
  return g * 128 + (rr >> 1) * 4 + half * 2 + (rr & 1);
}

// one warp per chunk: choose the encoding, report payload size and the work-order key
// desc = (first column, slots or cells, encoding | overflow cells << 8, payload bytes)
__global__ void __launch_bounds__(256)
k_plan(const int32_t *__restrict__ col, const int32_t *__restrict__ val,
       const int64_t *__restrict__ chunk_beg, int64_t nchunks, int64_t nbins, bool use_tiles, int dense_fill,
       int4 *__restrict__ desc, uint16_t *__restrict__ tovf, uint64_t *__restrict__ key,
       int32_t *__restrict__ ids) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < nchunks; c += nwarp) {
    const int64_t beg = chunk_beg[c], end = chunk_beg[c + 1];
    int32_t vmin = 2147483647, vmax = -2147483647 - 1;
    int n8 = 0, n16 = 0;                       // cells that do not fit u8 / u16
    int nbig = 0;                              // ... and do not fit a sparse-tile cell either
    for (int64_t k = beg + lane; k < end; k += 32) {
      const int32_t v = val[k];
      vmin = min(vmin, v);
      vmax = max(vmax, v);
      n8 += (v < 0 || v > 255);
      n16 += (v < 0 || v > 65535);
      nbig += (v < 0 || v >= (1 << kTileValBits));
    }
    for (int o = 16; o; o >>= 1) {
      vmin = min(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
      vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      n8 += __shfl_xor_sync(0xffffffffu, n8, o);
      n16 += __shfl_xor_sync(0xffffffffu, n16, o);
      nbig += __shfl_xor_sync(0xffffffffu, nbig, o);
    }
    if (lane == 0) {
      const int64_t count = end - beg;
      const int32_t cfirst = col[beg], clast = col[end - 1];
      int enc = ENC_S32, novf = 0;
      int64_t best = 8 * count;
      if (n16 == 0 && (int64_t)clast - cfirst < 65536 && 4 * count < best) {
        enc = ENC_S16; best = 4 * count;
      }
      // small non-negative counts: the cells go to the sparse tiles (tiles.cu), 4 B each there
      // (tile cells are weighted dense_fill bytes, not 4: see dense_fill_factor)
      if (use_tiles && vmin >= 0 && vmax < (1 << kTileValBits)) { enc = ENC_T; best = (int64_t)dense_fill * count; }
      int32_t c0 = cfirst;
      int64_t n = count;
      if (cfirst / kPanel == clast / kPanel) {          // inside one panel: dense is possible
        const int32_t p0 = (cfirst / kPanel) * kPanel;
        int64_t width = nbins - p0;
        if (width > kPanel) width = kPanel;
        const int64_t nslots = (width + kSub - 1) / kSub * kSub;     // whole sub-blocks (rowsum.cu)
        int denc = -1, dovf = 0;
        int64_t dbest = best + 1;
        if (4 * nslots <= best) { denc = ENC_D32; dbest = 4 * nslots; dovf = 0; }
        if (2 * nslots + 8 * (int64_t)n16 <= (denc >= 0 ? dbest : best)) {
          denc = ENC_D16; dbest = 2 * nslots + 8 * (int64_t)n16; dovf = n16;
        }
        // D8: cells above 255 go to the sparse tiles (4 B each) when they fit there, else to
        // the chunk's overflow list (8 B each)
        const int n8_list = use_tiles ? nbig : n8;
        const int64_t d8 = nslots + 8 * (int64_t)n8_list + 4 * (int64_t)(n8 - n8_list);
        if (d8 <= (denc >= 0 ? dbest : best)) {
          denc = ENC_D8; dbest = nslots + 8 * (int64_t)n8_list; dovf = n8_list;
        }
        if (denc >= 0) { enc = denc; best = dbest; c0 = p0; n = nslots; novf = dovf; }
      }
      if (enc == ENC_T) best = 0;                       // no payload in the chunk stream
      tovf[c] = (uint16_t)((enc == ENC_D8 && use_tiles) ? n8 - nbig : 0);
      desc[c] = make_int4(c0, (int)n, enc | (novf << 8), (int)align16(best));
      // work order: sparse chunks in row order (key 0), dense chunks by (panel, encoding),
      // tile chunks last (they are not walked chunk by chunk)
      key[c] = enc == ENC_T ? ((uint64_t)1 << 41)
               : (enc >= ENC_D32 ? (((uint64_t)1 << 40) | ((uint64_t)(uint32_t)c0 << 3) | (uint64_t)enc) : 0);
      ids[c] = (int32_t)c;
    }
  }
}

__global__ void k_sorted_bytes(const int32_t *__restrict__ ids, const int4 *__restrict__ desc,
                               int64_t nchunks, int64_t *__restrict__ bytes) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nchunks;
       k += (int64_t)gridDim.x * blockDim.x)
    bytes[k] = desc[ids[k]].w;
}

__global__ void k_scatter_off(const int32_t *__restrict__ ids, const int64_t *__restrict__ off_sorted,
                              int64_t nchunks, int64_t *__restrict__ off) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nchunks;
       k += (int64_t)gridDim.x * blockDim.x)
    off[ids[k]] = off_sorted[k];
}

__global__ void __launch_bounds__(256)
k_encode(const int32_t *__restrict__ col, const int32_t *__restrict__ val,
         const int64_t *__restrict__ chunk_beg, int64_t nchunks, const int4 *__restrict__ desc,
         const int64_t *__restrict__ off, bool use_tiles, uint8_t *__restrict__ data) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < nchunks; c += nwarp) {
    const int64_t beg = chunk_beg[c], end = chunk_beg[c + 1];
    const int4 d = desc[c];
    uint8_t *out = data + off[c];
    const int c0 = d.x, enc = d.z & 0xff;
    if (enc == ENC_S32) {
      int2 *o = reinterpret_cast<int2 *>(out);
      for (int64_t k = beg + lane; k < end; k += 32) o[k - beg] = make_int2(col[k], val[k]);
    } else if (enc == ENC_T) {
      // cells are written by build_tiles
    } else if (enc == ENC_S16) {
      uint32_t *o = reinterpret_cast<uint32_t *>(out);
      for (int64_t k = beg + lane; k < end; k += 32)
        o[k - beg] = (uint32_t)(col[k] - c0) | ((uint32_t)val[k] << 16);
    } else {
      const int width = enc == ENC_D32 ? 4 : (enc == ENC_D16 ? 2 : 1);
      const int32_t vcap = enc == ENC_D32 ? 2147483647 : (enc == ENC_D16 ? 65535 : 255);
      const int64_t nbytes = (int64_t)d.y * width;           // multiple of kSub
      uint4 *z = reinterpret_cast<uint4 *>(out);
      for (int64_t k = lane; k < nbytes / 16; k += 32) z[k] = make_uint4(0, 0, 0, 0);
      __syncwarp();
      int2 *ovf = reinterpret_cast<int2 *>(out + nbytes);
      int novf = 0;
      for (int64_t k0 = beg; k0 < end; k0 += 32) {
        const int64_t k = k0 + lane;
        int32_t v = 0, cc = 0;
        bool fits = true;
        if (k < end) {
          v = val[k];
          cc = col[k];
          fits = enc == ENC_D32 || (v >= 0 && v <= vcap);
          if (fits) {
            const int64_t p = dense_pos(cc - c0);
            if (enc == ENC_D32) reinterpret_cast<int32_t *>(out)[p] = v;
            else if (enc == ENC_D16) reinterpret_cast<uint16_t *>(out)[p] = (uint16_t)v;
            else out[p] = (uint8_t)v;
          }
        }
        // a D8 cell that fits a sparse-tile cell is stored there (tiles.cu), not in the list
        const bool listed = !fits && !(use_tiles && enc == ENC_D8 && v > 0 && v < (1 << kTileValBits));
        const unsigned m = __ballot_sync(0xffffffffu, listed);
        if (listed) ovf[novf + __popc(m & ((1u << lane) - 1u))] = make_int2(cc, v);
        novf += __popc(m);
      }
    }
  }
}

__global__ void k_count_enc(const int4 *__restrict__ desc, const uint16_t *__restrict__ tovf,
                            int64_t nchunks, unsigned long long *counts) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nchunks;
       c += (int64_t)gridDim.x * blockDim.x) {
    atomicAdd(&counts[desc[c].z & 0xff], 1ull);
    if (desc[c].z >> 8) atomicAdd(&counts[6], (unsigned long long)(desc[c].z >> 8));
    if (tovf[c]) atomicAdd(&counts[7], (unsigned long long)tovf[c]);
  }
}

}  // namespace

void build_encoded(tb_store *s) {
  cudaStream_t st = s->stream;
  const int64_t nc = s->nchunks;
  s->enc_bytes = 0;
  s->n_sparse = s->n_dense = s->n_groups = 0;
  s->n_overflow = s->n_overflow_tiled = 0;
  for (auto &c : s->enc_chunks) c = 0;
  s->enc_tovf.alloc(nc + 1);
  s->enc_off.alloc(nc + 1);
  s->enc_desc.alloc(nc + 1);
  s->work_ids.alloc(nc + 1);
  if (nc == 0) {
    s->enc_data.alloc(16);
    s->groups.alloc(1);
    return;
  }
  int64_t need = (nc + 7) / 8, cap = (int64_t)s->sm_count * 8;
  const int grid = (int)(need < cap ? need : cap);
  int64_t need1 = (nc + 255) / 256, cap1 = (int64_t)s->sm_count * 16;
  const int grid1 = (int)(need1 < cap1 ? need1 : cap1);
  const char *env = getenv("TB_TILES");
  const bool use_tiles = !(env && env[0] == '0');        // TB_TILES=0: A/B switch, L2-gather path only
  {
    DevBuf<uint64_t> key_in, key_out;
    DevBuf<int32_t> ids_in;
    DevBuf<int64_t> bytes, off_sorted;
    key_in.alloc(nc); key_out.alloc(nc); ids_in.alloc(nc);
    bytes.alloc(nc + 1); off_sorted.alloc(nc + 1);
    k_plan<<<grid, 256, 0, st>>>(s->col.p, s->val.p, s->chunk_beg.p, nc, s->nbins, use_tiles, dense_fill_factor(),
                                 s->enc_desc.p, s->enc_tovf.p, key_in.p, ids_in.p);
    size_t tmp_bytes = 0, tmp2 = 0;
    TB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_in.p, key_out.p, ids_in.p,
                                            s->work_ids.p, nc, 0, 42, st));
    TB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, bytes.p, off_sorted.p, nc + 1, st));
    DevBuf<uint8_t> tmp;
    tmp.alloc((tmp_bytes > tmp2 ? tmp_bytes : tmp2) + 16);
    TB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key_in.p, key_out.p, ids_in.p,
                                            s->work_ids.p, nc, 0, 42, st));
    TB_CUDA(cudaMemsetAsync(bytes.p + nc, 0, sizeof(int64_t), st));
    k_sorted_bytes<<<grid1, 256, 0, st>>>(s->work_ids.p, s->enc_desc.p, nc, bytes.p);
    TB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp2, bytes.p, off_sorted.p, nc + 1, st));
    k_scatter_off<<<grid1, 256, 0, st>>>(s->work_ids.p, off_sorted.p, nc, s->enc_off.p);
    TB_CUDA(cudaMemcpyAsync(&s->enc_bytes, off_sorted.p + nc, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    // Group records: runs of equal sort keys = dense chunks of one (encoding, panel), rows
    // ascending; each run is cut into groups of kRows seats, the last one padded with empty
    // seats.  One entry per chunk, so this runs on the host.
    std::vector<uint64_t> hk(nc);
    std::vector<int32_t> hid(nc);
    std::vector<int4> hdesc(nc);
    std::vector<int64_t> hoff(nc);
    TB_CUDA(cudaMemcpy(hk.data(), key_out.p, nc * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    TB_CUDA(cudaMemcpy(hid.data(), s->work_ids.p, nc * sizeof(int32_t), cudaMemcpyDeviceToHost));
    TB_CUDA(cudaMemcpy(hdesc.data(), s->enc_desc.p, nc * sizeof(int4), cudaMemcpyDeviceToHost));
    TB_CUDA(cudaMemcpy(hoff.data(), s->enc_off.p, nc * sizeof(int64_t), cudaMemcpyDeviceToHost));
    int64_t first_dense = 0;
    while (first_dense < nc && hk[first_dense] == 0) ++first_dense;
    std::vector<GroupRec> recs;
    recs.reserve((size_t)(nc - first_dense) / kRows + 64);
    for (int64_t i = first_dense; i < nc && hk[i] < ((uint64_t)1 << 41);) {
      int64_t j = i;
      while (j < nc && hk[j] == hk[i]) ++j;
      for (int64_t k = i; k < j; k += kRows) {
        GroupRec g;
        const int4 d0 = hdesc[hid[k]];
        const int enc = d0.z & 0xff;
        const int width = enc == ENC_D32 ? 4 : (enc == ENC_D16 ? 2 : 1);
        g.c0 = d0.x;
        g.nsub = d0.y * width / kSub;
        g.enc = enc;
        g.seg_end = 0;
        for (int r = 0; r < kRows; ++r) {
          if (k + r < j) {
            const int32_t id = hid[k + r];
            g.id[r] = id;
            g.off16[r] = (uint32_t)(hoff[id] >> 4);
            g.novf[r] = hdesc[id].z >> 8;
            if (g.novf[r] != 0) g.enc |= 1 << 8;
          } else {
            g.id[r] = -1;
            g.off16[r] = 0;
            g.novf[r] = 0;
          }
        }
        recs.push_back(g);
      }
      i = j;
    }
    TB_REQUIRE(s->enc_bytes < ((int64_t)1 << 36), "compact stream larger than 64 GB per GPU");
    // Panel segments (the x panel lives in shared memory per CTA) and the contiguous group range
    // of each CTA, cut at equal cost: slots (wider encodings weigh more, they are HBM-bound where
    // D8 is issue-bound); a seat with an overflow list stalls its warp on a chain of loads.
    std::vector<double> cost(recs.size() + 1, 0.0);
    for (size_t i = 0; i < recs.size();) {
      size_t j = i;
      while (j < recs.size() && recs[j].c0 == recs[i].c0) ++j;
      for (size_t k = i; k < j; ++k) recs[k].seg_end = (int32_t)j;
      i = j;
    }
    for (size_t i = 0; i < recs.size(); ++i) {
      const int enc = recs[i].enc & 0xff;
      const int width = enc == ENC_D32 ? 4 : (enc == ENC_D16 ? 2 : 1);
      double c = 0.0;
      for (int r = 0; r < kRows; ++r)
        if (recs[i].id[r] >= 0)
          c += (double)recs[i].nsub * kSub / width * (width == 1 ? 1.0 : 1.4 * (width / 2)) +
               (recs[i].novf[r] ? 600.0 + 8.0 * (double)recs[i].novf[r] : 0.0);
      cost[i + 1] = cost[i] + c;
    }
    int64_t nctas = ((int64_t)recs.size() + kDenseWarps - 1) / kDenseWarps;
    if (nctas > s->sm_count) nctas = s->sm_count;
    if (nctas < 1) nctas = 1;
    std::vector<int32_t> range(nctas + 1, 0);
    for (int64_t b = 1; b < nctas; ++b) {
      const double want = cost[recs.size()] * (double)b / (double)nctas;
      range[b] = (int32_t)(std::lower_bound(cost.begin(), cost.end(), want) - cost.begin());
      if (range[b] < range[b - 1]) range[b] = range[b - 1];
    }
    range[nctas] = (int32_t)recs.size();
    s->n_dense_ctas = (int)nctas;
    s->group_range.alloc(nctas + 1);
    TB_CUDA(cudaMemcpy(s->group_range.p, range.data(), (nctas + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
    s->n_groups = (int64_t)recs.size();
    s->groups.alloc(recs.size() ? recs.size() : 1);
    if (!recs.empty())
      TB_CUDA(cudaMemcpy(s->groups.p, recs.data(), recs.size() * sizeof(GroupRec),
                         cudaMemcpyHostToDevice));
  }
  s->enc_data.alloc(s->enc_bytes + 16);
  k_encode<<<grid, 256, 0, st>>>(s->col.p, s->val.p, s->chunk_beg.p, nc, s->enc_desc.p, s->enc_off.p,
                                 use_tiles, s->enc_data.p);
  DevBuf<unsigned long long> counts;
  counts.alloc(8);
  TB_CUDA(cudaMemsetAsync(counts.p, 0, counts.bytes(), st));
  k_count_enc<<<grid1, 256, 0, st>>>(s->enc_desc.p, s->enc_tovf.p, nc, counts.p);
  unsigned long long h[8];
  TB_CUDA(cudaMemcpyAsync(h, counts.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < 6; ++k) s->enc_chunks[k] = (int64_t)h[k];
  s->n_overflow = (int64_t)h[6];
  s->n_overflow_tiled = (int64_t)h[7];
  s->n_sparse = s->enc_chunks[ENC_S32] + s->enc_chunks[ENC_S16];
  s->n_dense = s->enc_chunks[ENC_D32] + s->enc_chunks[ENC_D16] + s->enc_chunks[ENC_D8];
}

}  // namespace tb
