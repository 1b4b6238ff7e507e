// Synthetic Hi-C contact matrices generated on the device (bench / scale tests only).
//
// Model: per-bin visibility b_i ~ LogNormal(0, sigma), a fraction of bins
// scaled down (low coverage) or emptied; cis mean a*b_i*b_j*(|i-j|+1)^-alpha, trans mean
// t*b_i*b_j; count ~ Poisson(mean).  Every CELL draws from a counter-based hash of
// (seed, min(i,j), max(i,j)), so cell (i,j) == cell (j,i) and the matrix is the same
// however its rows are sharded over GPUs.  A warp walks one full row; stored cells are
// compacted in column order (ballot + popc), so the CSR comes out sorted with no sort.
#include <cub/cub.cuh>
#include <math.h>

#include "common.cuh"

namespace tb {

namespace {

struct SynthDev {
  uint32_t seed_lo, seed_hi;
  float trans_mean, sigma, low_frac, low_scale, empty_frac;
};

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x7FEB352Du;
  h ^= h >> 15; h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}
__device__ __forceinline__ float unit(uint32_t h) {      // (0,1), 24 bits
  return ((float)(h >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

__global__ void k_visibility(SynthDev p, int64_t n, float *vis) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t h1 = mix32((uint32_t)i ^ p.seed_lo);
    uint32_t h2 = mix32(h1 + p.seed_hi + 0x9E3779B9u);
    uint32_t h3 = mix32(h2 ^ 0x85EBCA6Bu);
    uint32_t h4 = mix32(h3 + 0xC2B2AE35u);
    float z = sqrtf(-2.0f * logf(unit(h1))) * cospif(2.0f * unit(h2));
    float b = expf(p.sigma * z);
    if (unit(h3) < p.low_frac) b *= p.low_scale;
    if (unit(h4) < p.empty_frac) b = 0.0f;
    vis[i] = b;
  }
}

__global__ void k_distance_table(double a, double alpha, int64_t n, float *tab) {
  for (int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; d < n;
       d += (int64_t)gridDim.x * blockDim.x)
    tab[d] = (float)(a * pow((double)d + 1.0, -alpha));
}

__device__ __forceinline__ int32_t cell_count(const SynthDev &p, int64_t i, int64_t j,
                                              const float *__restrict__ vis,
                                              const int32_t *__restrict__ chrom_of,
                                              const float *__restrict__ tab) {
  const int64_t lo = i < j ? i : j, hi = i < j ? j : i;
  const float bb = __ldg(vis + lo) * __ldg(vis + hi);
  const float lam = (__ldg(chrom_of + lo) == __ldg(chrom_of + hi)) ? __ldg(tab + (hi - lo)) * bb
                                                                  : p.trans_mean * bb;
  const uint32_t h = mix32(mix32((uint32_t)lo ^ p.seed_lo) + (uint32_t)hi * 0x9E3779B1u + p.seed_hi);
  const float u = unit(h);
  if (!(lam > 0.0f) || u >= lam) return 0;         // 1 - exp(-lam) <= lam
  if (lam < 32.0f) {
    const float w = 1.0f - u;
    float pk = expf(-lam), cdf = pk;
    int k = 0;
    while (cdf < w && k < 255) {
      ++k;
      pk *= lam / (float)k;
      cdf += pk;
    }
    return k;
  }
  const uint32_t h2 = mix32(h ^ 0x68E31DA4u);
  const float z = sqrtf(-2.0f * logf(u)) * cospif(2.0f * unit(h2));
  const float v = rintf(lam + sqrtf(lam) * z);
  return v > 0.0f ? (v < 2.0e9f ? (int32_t)v : 2000000000) : 0;
}

template <bool FILL>
__global__ void __launch_bounds__(256)
k_rows(SynthDev p, int64_t n, int64_t rb, int64_t nrows, const float *__restrict__ vis,
       const int32_t *__restrict__ chrom_of, const float *__restrict__ tab,
       int64_t *__restrict__ counts, const int64_t *__restrict__ rowptr,
       int32_t *__restrict__ col, int32_t *__restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < nrows; r += nwarp) {
    const int64_t i = rb + r;
    int64_t base = FILL ? rowptr[r] : 0;
    if (__ldg(vis + i) == 0.0f) {          // empty bin: nothing stored in this row
      if (!FILL && lane == 0) counts[r] = 0;
      continue;
    }
    for (int64_t j0 = 0; j0 < n; j0 += 32) {
      const int64_t j = j0 + lane;
      int32_t k = 0;
      if (j < n) k = cell_count(p, i, j, vis, chrom_of, tab);
      const unsigned m = __ballot_sync(0xffffffffu, k > 0);
      if (FILL && k > 0) {
        const int64_t pos = base + __popc(m & ((1u << lane) - 1u));
        col[pos] = (int32_t)j;
        val[pos] = k;
      }
      base += __popc(m);
    }
    if (!FILL && lane == 0) counts[r] = base;
  }
}

struct Tables {
  DevBuf<float> vis, tab;
  DevBuf<int32_t> chrom_of;
  SynthDev dev;
};

void make_tables(Tables &t, int64_t n, const std::vector<int64_t> &offs, const tb_synth_params *p,
                 const int32_t *d_chrom_of, cudaStream_t st, int sm_count) {
  t.dev.seed_lo = (uint32_t)(p->seed & 0xffffffffu);
  t.dev.seed_hi = (uint32_t)(p->seed >> 32);
  t.dev.trans_mean = (float)p->trans_mean;
  t.dev.sigma = (float)p->vis_sigma;
  t.dev.low_frac = (float)p->low_frac;
  t.dev.low_scale = (float)p->low_scale;
  t.dev.empty_frac = (float)p->empty_frac;
  int64_t maxlen = 1;
  for (size_t c = 0; c + 1 < offs.size(); ++c)
    if (offs[c + 1] - offs[c] > maxlen) maxlen = offs[c + 1] - offs[c];
  t.vis.alloc(n);
  t.tab.alloc(maxlen);
  const int g = sm_count * 8;
  k_visibility<<<g, 256, 0, st>>>(t.dev, n, t.vis.p);
  k_distance_table<<<g, 256, 0, st>>>(p->cis_scale, p->cis_alpha, maxlen, t.tab.p);
  if (!d_chrom_of) {
    std::vector<int32_t> h(n);
    for (int32_t c = 0; c + 1 < (int32_t)offs.size(); ++c)
      for (int64_t i = offs[c]; i < offs[c + 1]; ++i) h[i] = c;
    t.chrom_of.alloc(n);
    TB_CUDA(cudaMemcpy(t.chrom_of.p, h.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
}

}  // namespace

void build_synthetic(tb_store *s, const tb_synth_params *p) {
  cudaStream_t st = s->stream;
  const int64_t n = s->nbins, nrows = s->nrows;
  Tables t;
  make_tables(t, n, s->chrom_offsets, p, s->chrom_of.p, st, s->sm_count);
  const int g = s->sm_count * 8;
  DevBuf<int64_t> counts;
  counts.alloc(nrows + 1);
  TB_CUDA(cudaMemsetAsync(counts.p, 0, counts.bytes(), st));
  s->rowptr.alloc(nrows + 1);
  if (nrows)
    k_rows<false><<<g, 256, 0, st>>>(t.dev, n, s->row_begin, nrows, t.vis.p, s->chrom_of.p, t.tab.p,
                                     counts.p, nullptr, nullptr, nullptr);
  {
    size_t tmp_bytes = 0;
    TB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts.p, s->rowptr.p, nrows + 1, st));
    DevBuf<uint8_t> tmp;
    tmp.alloc(tmp_bytes ? tmp_bytes : 1);
    TB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, counts.p, s->rowptr.p, nrows + 1, st));
    TB_CUDA(cudaMemcpyAsync(&s->nnz_full, s->rowptr.p + nrows, sizeof(int64_t),
                            cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
  }
  s->col.alloc(s->nnz_full ? s->nnz_full : 1);
  s->val.alloc(s->nnz_full ? s->nnz_full : 1);
  if (nrows && s->nnz_full)
    k_rows<true><<<g, 256, 0, st>>>(t.dev, n, s->row_begin, nrows, t.vis.p, s->chrom_of.p, t.tab.p,
                                    nullptr, s->rowptr.p, s->col.p, s->val.p);
  TB_CUDA(cudaStreamSynchronize(st));
  finish_layout(s);
}

void synthetic_row_counts(int64_t nbins, int32_t n_chrom, const int64_t *chrom_offsets,
                          const tb_synth_params *p, int64_t rb, int64_t re, int device,
                          int64_t *h_counts) {
  TB_REQUIRE(nbins > 0 && rb >= 0 && re >= rb && re <= nbins && p && h_counts,
             "synthetic_row_counts: bad arguments");
  TB_CUDA(cudaSetDevice(device));
  struct { int multiProcessorCount; } prop;
  TB_CUDA(cudaDeviceGetAttribute(&prop.multiProcessorCount, cudaDevAttrMultiProcessorCount, device));
  std::vector<int64_t> offs;
  if (n_chrom > 0) offs.assign(chrom_offsets, chrom_offsets + n_chrom + 1);
  else offs = {0, nbins};
  Tables t;
  make_tables(t, nbins, offs, p, nullptr, nullptr, prop.multiProcessorCount);
  const int64_t nrows = re - rb;
  if (!nrows) return;
  DevBuf<int64_t> counts;
  counts.alloc(nrows);
  k_rows<false><<<prop.multiProcessorCount * 8, 256>>>(t.dev, nbins, rb, nrows, t.vis.p,
                                                       t.chrom_of.p, t.tab.p, counts.p, nullptr,
                                                       nullptr, nullptr);
  TB_CUDA(cudaMemcpy(h_counts, counts.p, nrows * sizeof(int64_t), cudaMemcpyDeviceToHost));
}

}  // namespace tb
