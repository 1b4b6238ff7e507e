// The consumers of the biases that stream the matrix once more, as
// cell visitors (visit.cuh) over the compact stream:
//
//   N1  normalised values and z-scores of the interaction pairs
//       (_pytadbit/experiment.py:737-743 Experiment.normalize_hic builds hic[i,j] / bias[i] / bias[j]
//        for all N^2 pairs; :751-799 get_hic_zscores + utils/tadmaths.py:94-172 take the log10 and
//        the z-score of the pairs i < j between columns that are not filtered out)
//         pair_stats   count, smallest value, sum of log10 and sum of squared deviations of the
//                      non-zero pairs (everything the z-score needs; pairs with no contact are
//                      accounted for in closed form by the host)
//         pair_values  the pairs themselves (row, column, raw | normalised | z-score), row-major
//   N2  what `tadbit normalize` derives from the biases (_pytadbit/tools/tadbit_normalize.py:963-1107,
//       helpers :1110-1159): cis / total normalised interactions (get_cis_perc) and, per chromosome
//       and genomic distance, the sums of normalised and raw counts (sum_dec_matrix)
//         decay_sums
//
// Reductions of doubles go through per-CTA partials that are added in CTA order on the host side of
// the call (same grid, same static work split -> same bits every run); the per-distance sums are
// atomics on doubles (their order is not fixed: last-bit noise, far below the 1e-10 of north_star).
#include <math.h>

#include <cub/cub.cuh>

#include "common.cuh"
#include "visit.cuh"

namespace tb {

namespace {

inline int grid_1d(int64_t n, int threads, int sm_count, int per_sm = 8) {
  int64_t g = (n + threads - 1) / threads, cap = (int64_t)sm_count * per_sm;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

// 1 / bias on good bins (NaN biases count as bad: tadbit normalize marks bad columns that way)
__global__ void k_pair_weights(const double *__restrict__ bias, const uint8_t *__restrict__ bad, int64_t n,
                               double *__restrict__ w, uint8_t *__restrict__ skip) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const bool b = bad[i] || (bias && isnan(bias[i]));
    skip[i] = b;
    w[i] = bias ? bias[i] : 1.0;
  }
}

constexpr int kStatSlots = 4;        // per CTA: count, min, sum log10, sum squared deviation

// value of the pair as the reference computes it: float(hic[i, j]) / bias[i] / bias[j]
__device__ __forceinline__ double pair_value(int32_t v, const double *w, int64_t i, int64_t j, bool normalised) {
  return normalised ? (double)v / w[i] / w[j] : (double)v;
}

struct PairStatsVisitor {
  static constexpr int kShared = 32 * kStatSlots * 8;
  static constexpr int kThreads = 256;
  const uint8_t *skip;
  const double *w;
  bool normalised;
  double center;
  double *cta_out;                   // [gridDim.x * kStatSlots] of the kernel this visitor runs in
  double cnt, mn, sum, dev2;
  __device__ void begin(uint8_t *) { cnt = 0.0; mn = INFINITY; sum = 0.0; dev2 = 0.0; }
  __device__ bool block(int64_t, int64_t, int64_t, int64_t) { return true; }
  __device__ bool row(int64_t i) { return !skip[i]; }
  __device__ void cell(int64_t i, int64_t j, int32_t v) {
    if (j == i || v == 0 || skip[j]) return;
    const double x = pair_value(v, w, i, j, normalised);
    const double l = log10(x);
    cnt += 1.0;
    mn = fmin(mn, x);
    sum += l;
    dev2 += (l - center) * (l - center);
  }
  __device__ void end(uint8_t *smem) {
    double *sh = reinterpret_cast<double *>(smem);
    for (int o = 16; o; o >>= 1) {
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      dev2 += __shfl_xor_sync(0xffffffffu, dev2, o);
    }
    const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) {
      sh[wid * kStatSlots] = cnt; sh[wid * kStatSlots + 1] = mn;
      sh[wid * kStatSlots + 2] = sum; sh[wid * kStatSlots + 3] = dev2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < nw; ++k) {
        cnt += sh[k * kStatSlots]; mn = fmin(mn, sh[k * kStatSlots + 1]);
        sum += sh[k * kStatSlots + 2]; dev2 += sh[k * kStatSlots + 3];
      }
      // the chunk and the tile kernel both run this visitor: each adds to its CTA's slots
      double *o = cta_out + (size_t)blockIdx.x * kStatSlots;
      o[0] += cnt; o[1] = fmin(o[1], mn); o[2] += sum; o[3] += dev2;
    }
  }
};

__global__ void k_stats_init(double *o, int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    o[k] = (k % kStatSlots) == 1 ? INFINITY : 0.0;
}

struct PairCountVisitor {
  static constexpr int kShared = 0;
  static constexpr int kThreads = 256;
  const uint8_t *skip;
  int64_t rb;
  unsigned long long *cnt;            // [nrows]
  __device__ void begin(uint8_t *) {}
  __device__ bool block(int64_t, int64_t, int64_t, int64_t) { return true; }
  __device__ bool row(int64_t i) { return !skip[i]; }
  __device__ void cell(int64_t i, int64_t j, int32_t v) {
    if (j == i || v == 0 || skip[j]) return;
    atomicAdd(&cnt[i - rb], 1ull);
  }
  __device__ void end(uint8_t *) {}
};

struct PairPlaceVisitor {
  static constexpr int kShared = 0;
  static constexpr int kThreads = 256;
  const uint8_t *skip;
  const double *w;
  bool normalised, zscored;
  double mean, inv_std;
  int64_t rb;
  const int64_t *off;
  unsigned long long *cursor;
  uint64_t *key;
  double *val;
  __device__ void begin(uint8_t *) {}
  __device__ bool block(int64_t, int64_t, int64_t, int64_t) { return true; }
  __device__ bool row(int64_t i) { return !skip[i]; }
  __device__ void cell(int64_t i, int64_t j, int32_t v) {
    if (j == i || v == 0 || skip[j]) return;
    double x = pair_value(v, w, i, j, normalised);
    if (zscored) x = (log10(x) - mean) * inv_std;
    const int64_t r = i - rb;
    const int64_t pos = off[r] + (int64_t)atomicAdd(&cursor[r], 1ull);
    key[pos] = ((uint64_t)r << 32) | (uint32_t)j;
    val[pos] = x;
  }
  __device__ void end(uint8_t *) {}
};

__global__ void k_unpack_pairs(const uint64_t *__restrict__ key, int64_t n, int64_t rb,
                               int32_t *__restrict__ orow, int32_t *__restrict__ ocol) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
       k += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t e = key[k];
    orow[k] = (int32_t)(rb + (int64_t)(e >> 32));
    ocol[k] = (int32_t)(e & 0xffffffffu);
  }
}

// N2: cells i <= j between good bins.  Off-diagonal: cis / total normalised sums (get_cis_perc takes
// i > j of a dict that holds both triangles: every pair once, diagonal excluded).  Same section:
// per-distance sums at index section start + distance, diagonal included (sum_dec_matrix, i >= j).
struct DecayVisitor {
  static constexpr int kShared = 32 * 2 * 8;
  static constexpr int kThreads = 256;
  const uint8_t *skip;
  const double *w;
  const int32_t *chrom_of;
  const int64_t *chrom_end;
  const int64_t *chrom_beg;
  double *nrm;                        // [nbins]
  unsigned long long *raw;            // [nbins]
  double *cta_out;                    // [gridDim.x * 2] cis, total
  double cis, total;
  int64_t jend, base;
  __device__ void begin(uint8_t *) { cis = 0.0; total = 0.0; }
  __device__ bool block(int64_t, int64_t, int64_t, int64_t) { return true; }
  __device__ bool row(int64_t i) {
    if (skip[i]) return false;
    const int32_t c = chrom_of[i];
    jend = chrom_end[c];
    base = chrom_beg[c];
    return true;
  }
  __device__ void cell(int64_t i, int64_t j, int32_t v) {
    if (skip[j]) return;
    const double x = (double)v / w[i] / w[j];
    const bool same = j < jend;
    if (j != i) {
      total += x;
      if (same) cis += x;
    }
    if (same) {
      atomicAdd(&nrm[base + (j - i)], x);
      atomicAdd(&raw[base + (j - i)], (unsigned long long)(long long)v);
    }
  }
  __device__ void end(uint8_t *smem) {
    double *sh = reinterpret_cast<double *>(smem);
    for (int o = 16; o; o >>= 1) {
      cis += __shfl_xor_sync(0xffffffffu, cis, o);
      total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) { sh[2 * wid] = cis; sh[2 * wid + 1] = total; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < nw; ++k) { cis += sh[2 * k]; total += sh[2 * k + 1]; }
      cta_out[2 * (size_t)blockIdx.x] += cis;
      cta_out[2 * (size_t)blockIdx.x + 1] += total;
    }
  }
};

// N3: a dense window of the (normalised) matrix, rows [r0, r1) x columns [c0, c1)
struct WindowVisitor {
  static constexpr int kShared = 0;
  static constexpr int kThreads = 256;
  int64_t r0, r1, c0, c1;
  const double *w;                    // biases, or null for raw counts
  double *out;                        // [(r1 - r0) * (c1 - c0)], zero-initialised
  __device__ void begin(uint8_t *) {}
  __device__ bool block(int64_t i0, int64_t i1, int64_t j0, int64_t j1) {
    return i0 < r1 && i1 > r0 && j0 < c1 && j1 > c0;
  }
  __device__ bool row(int64_t i) { return i >= r0 && i < r1; }
  __device__ void cell(int64_t i, int64_t j, int32_t v) {
    if (j < c0 || j >= c1) return;
    out[(i - r0) * (c1 - c0) + (j - c0)] = w ? (double)v / w[i] / w[j] : (double)v;
  }
  __device__ void end(uint8_t *) {}
};

// N3 (extraction side): the stored cells inside rows [r0, r1) x columns [c0, c1) of the full
// symmetric matrix between bins that are not bad (get_matrix / write_matrix of
// _pytadbit/parsers/hic_bam_parser.py:755-872, :892-1242: `if i not in bads1 and j not in bads2`).
// half: only the cells whose position inside the row range is not behind their position inside
// the column range (what _read_half_bam_frag keeps, :447-458: `if pos1 > pos2: continue` on the
// region-local indices).
struct RegionVisitorBase {
  static constexpr int kShared = 0;
  static constexpr int kThreads = 256;
  int64_t r0, r1, c0, c1, rb;
  bool half;
  const uint8_t *skip;
  __device__ void begin(uint8_t *) {}
  __device__ bool block(int64_t i0, int64_t i1, int64_t j0, int64_t j1) {
    return i0 < r1 && i1 > r0 && j0 < c1 && j1 > c0;
  }
  __device__ bool row(int64_t i) { return i >= r0 && i < r1 && !skip[i]; }
  __device__ bool wanted(int64_t i, int64_t j, int32_t v) const {
    if (j < c0 || j >= c1 || v == 0 || skip[j]) return false;
    return !half || (i - r0) <= (j - c0);
  }
  __device__ void end(uint8_t *) {}
};

struct RegionCountVisitor : RegionVisitorBase {
  unsigned long long *cnt;            // [nrows]
  __device__ void cell(int64_t i, int64_t j, int32_t v) {
    if (wanted(i, j, v)) atomicAdd(&cnt[i - rb], 1ull);
  }
};

struct RegionPlaceVisitor : RegionVisitorBase {
  const int64_t *off;
  unsigned long long *cursor;
  uint64_t *key;
  int32_t *raw;
  __device__ void cell(int64_t i, int64_t j, int32_t v) {
    if (!wanted(i, j, v)) return;
    const int64_t r = i - rb;
    const int64_t pos = off[r] + (int64_t)atomicAdd(&cursor[r], 1ull);
    key[pos] = ((uint64_t)r << 32) | (uint32_t)j;
    raw[pos] = v;
  }
};

// sorted keys -> (row, column) and the transformed values: transform_value_norm / _decay of
// hic_bam_parser.py:828-833 (v / bias1[a] / bias2[b] / decay[c][|a - b|], the same order of
// divisions; between two sections there is no decay: NaN)
__global__ void k_unpack_region(const uint64_t *__restrict__ key, const int32_t *__restrict__ raw, int64_t n,
                                int64_t rb, const double *__restrict__ bias, const double *__restrict__ decay,
                                const int32_t *__restrict__ chrom_of, const int64_t *__restrict__ chrom_beg,
                                int32_t *__restrict__ orow, int32_t *__restrict__ ocol,
                                double *__restrict__ onrm, double *__restrict__ odec) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
       k += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t e = key[k];
    const int64_t i = rb + (int64_t)(e >> 32), j = (int64_t)(e & 0xffffffffu);
    orow[k] = (int32_t)i;
    ocol[k] = (int32_t)j;
    if (!bias) continue;
    const double x = (double)raw[k] / bias[i] / bias[j];
    onrm[k] = x;
    if (!decay) continue;
    const int32_t c = chrom_of[i];
    const int64_t d = i < j ? j - i : i - j;
    odec[k] = chrom_of[j] == c ? x / decay[chrom_beg[c] + d] : nan("");
  }
}

struct Weights {
  DevBuf<double> w;
  DevBuf<uint8_t> skip;
};

void make_weights(tb_store *s, const double *h_bias, const uint8_t *h_bad, Weights &wt) {
  cudaStream_t st = s->stream;
  const int64_t N = s->nbins;
  wt.w.alloc(N);
  wt.skip.alloc(N);
  if (h_bad) TB_CUDA(cudaMemcpyAsync(s->bad.p, h_bad, N, cudaMemcpyHostToDevice, st));
  else TB_CUDA(cudaMemsetAsync(s->bad.p, 0, N, st));
  const double *bias = nullptr;
  if (h_bias) {
    TB_CUDA(cudaMemcpyAsync(s->b.p, h_bias, N * sizeof(double), cudaMemcpyHostToDevice, st));
    bias = s->b.p;
  }
  k_pair_weights<<<grid_1d(N, 256, s->sm_count), 256, 0, st>>>(bias, s->bad.p, N, wt.w.p, wt.skip.p);
}

constexpr int kMaxCtas = 148 * 16 * 4;

}  // namespace

// out[0] pairs with a non-zero value, [1] the smallest such value, [2] sum of log10(value),
// [3] sum of (log10(value) - center)^2
void pair_stats(tb_store *s, const double *h_bias, const uint8_t *h_bad, double center, double out[4]) {
  cudaStream_t st = s->stream;
  Weights wt;
  make_weights(s, h_bias, h_bad, wt);
  DevBuf<double> cta;
  cta.alloc((size_t)kMaxCtas * kStatSlots);
  k_stats_init<<<64, 256, 0, st>>>(cta.p, kMaxCtas * kStatSlots);
  PairStatsVisitor vis;
  vis.skip = wt.skip.p;
  vis.w = wt.w.p;
  vis.normalised = h_bias != nullptr;
  vis.center = center;
  vis.cta_out = cta.p;
  vis.cnt = vis.sum = vis.dev2 = 0.0;
  vis.mn = 0.0;
  visit_upper(s, vis);
  std::vector<double> h((size_t)kMaxCtas * kStatSlots);
  TB_CUDA(cudaMemcpyAsync(h.data(), cta.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  double cnt = 0.0, mn = INFINITY, sum = 0.0, dev2 = 0.0;
  for (int k = 0; k < kMaxCtas; ++k) {            // CTA order: the same bits every run
    cnt += h[(size_t)k * kStatSlots];
    mn = fmin(mn, h[(size_t)k * kStatSlots + 1]);
    sum += h[(size_t)k * kStatSlots + 2];
    dev2 += h[(size_t)k * kStatSlots + 3];
  }
  if (s->world > 1) {                             // sums over the shards; min as -max
    double part[4] = {cnt, sum, dev2, 0.0};
    TB_CUDA(cudaMemcpyAsync(s->scal_f.p, part, sizeof(part), cudaMemcpyHostToDevice, st));
    allreduce_f64(s, s->scal_f.p, s->scal_f.p + 8, 4);
    TB_CUDA(cudaMemcpyAsync(part, s->scal_f.p + 8, sizeof(part), cudaMemcpyDeviceToHost, st));
    // the minimum: every rank contributes its own in its own slot, the sum carries them all
    std::vector<double> slots(s->world, 0.0), all(s->world, 0.0);
    slots[s->rank] = isinf(mn) ? 0.0 : mn;
    DevBuf<double> dm;
    dm.alloc(2 * (size_t)s->world);
    TB_CUDA(cudaMemcpyAsync(dm.p, slots.data(), s->world * sizeof(double), cudaMemcpyHostToDevice, st));
    allreduce_f64(s, dm.p, dm.p + s->world, s->world);
    TB_CUDA(cudaMemcpyAsync(all.data(), dm.p + s->world, s->world * sizeof(double), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    cnt = part[0]; sum = part[1]; dev2 = part[2];
    mn = INFINITY;
    for (double v : all) if (v > 0.0) mn = fmin(mn, v);
  }
  out[0] = cnt; out[1] = mn; out[2] = sum; out[3] = dev2;
}

// The non-zero pairs i < j between good bins of the owned rows, row-major: value = raw count,
// normalised count (h_bias given) or z-score of its log10 (zscored).  Call with null arrays to get *n.
void pair_values(tb_store *s, const double *h_bias, const uint8_t *h_bad, int zscored, double mean, double std,
                 int64_t *n, int32_t *h_row, int32_t *h_col, double *h_val) {
  cudaStream_t st = s->stream;
  const int64_t nrows = s->nrows;
  Weights wt;
  make_weights(s, h_bias, h_bad, wt);
  DevBuf<unsigned long long> cnt;
  DevBuf<int64_t> off;
  cnt.alloc(nrows + 1);
  off.alloc(nrows + 1);
  TB_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), st));
  PairCountVisitor cv;
  cv.skip = wt.skip.p;
  cv.rb = s->row_begin;
  cv.cnt = cnt.p;
  visit_upper(s, cv);
  {
    size_t tmp_bytes = 0;
    TB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, reinterpret_cast<const int64_t *>(cnt.p), off.p,
                                          nrows + 1, st));
    DevBuf<uint8_t> tmp;
    tmp.alloc(tmp_bytes ? tmp_bytes : 1);
    TB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, reinterpret_cast<const int64_t *>(cnt.p), off.p,
                                          nrows + 1, st));
    TB_CUDA(cudaStreamSynchronize(st));
  }
  int64_t m = 0;
  TB_CUDA(cudaMemcpy(&m, off.p + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost));
  *n = m;
  if (!h_row || m == 0) return;
  TB_REQUIRE(h_col && h_val, "pair_values: null output buffer");
  DevBuf<uint64_t> key_in, key_out;
  DevBuf<double> val_in, val_out;
  DevBuf<int32_t> orow, ocol;
  key_in.alloc(m); key_out.alloc(m); val_in.alloc(m); val_out.alloc(m); orow.alloc(m); ocol.alloc(m);
  TB_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), st));
  PairPlaceVisitor pv;
  pv.skip = wt.skip.p;
  pv.w = wt.w.p;
  pv.normalised = h_bias != nullptr;
  pv.zscored = zscored != 0;
  pv.mean = mean;
  pv.inv_std = zscored ? 1.0 / std : 1.0;
  pv.rb = s->row_begin;
  pv.off = off.p;
  pv.cursor = cnt.p;
  pv.key = key_in.p;
  pv.val = val_in.p;
  visit_upper(s, pv);
  int row_bits = 1;
  while (((int64_t)1 << row_bits) < nrows) ++row_bits;
  size_t tmp_bytes = 0;
  TB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_in.p, key_out.p, val_in.p, val_out.p, m, 0,
                                          32 + row_bits, st));
  DevBuf<uint8_t> tmp;
  tmp.alloc(tmp_bytes ? tmp_bytes : 1);
  TB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key_in.p, key_out.p, val_in.p, val_out.p, m, 0,
                                          32 + row_bits, st));
  k_unpack_pairs<<<grid_1d(m, 256, s->sm_count), 256, 0, st>>>(key_out.p, m, s->row_begin, orow.p, ocol.p);
  TB_CUDA(cudaMemcpyAsync(h_row, orow.p, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(h_col, ocol.p, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(h_val, val_out.p, m * sizeof(double), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
}

// h_nrm / h_raw [nbins]: entry (first bin of section c) + d = sum over the cells (i, i + d) of section c
// between good bins of count / bias_i / bias_j and of count.  out[0] cis, out[1] total normalised
// interactions of the off-diagonal pairs between good bins.
void decay_sums(tb_store *s, const double *h_bias, const uint8_t *h_bad, double *h_nrm, int64_t *h_raw,
                double out[2]) {
  cudaStream_t st = s->stream;
  const int64_t N = s->nbins;
  TB_REQUIRE(h_bias && h_nrm && h_raw && out, "decay_sums: null argument");
  Weights wt;
  make_weights(s, h_bias, h_bad, wt);
  DevBuf<double> nrm, cta;
  DevBuf<unsigned long long> raw;
  DevBuf<int64_t> beg;
  nrm.alloc(2 * N); raw.alloc(2 * N); cta.alloc((size_t)kMaxCtas * 2);
  const size_t nsec = s->chrom_offsets.size() - 1;
  beg.alloc(nsec);
  TB_CUDA(cudaMemcpyAsync(beg.p, s->chrom_offsets.data(), nsec * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  TB_CUDA(cudaMemsetAsync(nrm.p, 0, nrm.bytes(), st));
  TB_CUDA(cudaMemsetAsync(raw.p, 0, raw.bytes(), st));
  TB_CUDA(cudaMemsetAsync(cta.p, 0, cta.bytes(), st));
  DecayVisitor vis;
  vis.skip = wt.skip.p;
  vis.w = wt.w.p;
  vis.chrom_of = s->chrom_of.p;
  vis.chrom_end = s->chrom_end.p;
  vis.chrom_beg = beg.p;
  vis.nrm = nrm.p;
  vis.raw = raw.p;
  vis.cta_out = cta.p;
  vis.cis = vis.total = 0.0;
  vis.jend = vis.base = 0;
  visit_upper(s, vis);
  std::vector<double> h((size_t)kMaxCtas * 2);
  TB_CUDA(cudaMemcpyAsync(h.data(), cta.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  double cis = 0.0, total = 0.0;
  for (int k = 0; k < kMaxCtas; ++k) { cis += h[2 * (size_t)k]; total += h[2 * (size_t)k + 1]; }
  const double *src_n = nrm.p;
  const unsigned long long *src_r = raw.p;
  if (s->world > 1) {
    double part[2] = {cis, total};
    TB_CUDA(cudaMemcpyAsync(s->scal_f.p, part, sizeof(part), cudaMemcpyHostToDevice, st));
    allreduce_f64(s, s->scal_f.p, s->scal_f.p + 8, 2);
    TB_CUDA(cudaMemcpyAsync(part, s->scal_f.p + 8, sizeof(part), cudaMemcpyDeviceToHost, st));
    allreduce_f64(s, nrm.p, nrm.p + N, N);
    allreduce_i64(s, reinterpret_cast<const long long *>(raw.p), reinterpret_cast<long long *>(raw.p + N), N);
    TB_CUDA(cudaStreamSynchronize(st));
    cis = part[0]; total = part[1];
    src_n = nrm.p + N;
    src_r = raw.p + N;
  }
  TB_CUDA(cudaMemcpyAsync(h_nrm, src_n, N * sizeof(double), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(h_raw, src_r, N * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  out[0] = cis;
  out[1] = total;
}

// Dense window rows [r0, r1) x columns [c0, c1) of the full symmetric matrix: counts, or
// count / bias[i] / bias[j] when h_bias is given (HiC_data.get_matrix / yield_matrix,
// _pytadbit/hic_data.py:632-686, :1514-1582).  A row-sharded store fills the rows it owns.
void window(tb_store *s, int64_t r0, int64_t r1, int64_t c0, int64_t c1, const double *h_bias, double *h_out) {
  cudaStream_t st = s->stream;
  const int64_t N = s->nbins;
  TB_REQUIRE(r0 >= 0 && r0 <= r1 && r1 <= N && c0 >= 0 && c0 <= c1 && c1 <= N && h_out, "window: bad arguments");
  const int64_t cells = (r1 - r0) * (c1 - c0);
  if (cells == 0) return;
  TB_REQUIRE(cells < (1LL << 33), "window: more than 2^33 cells requested");
  DevBuf<double> out;
  out.alloc(cells);
  TB_CUDA(cudaMemsetAsync(out.p, 0, out.bytes(), st));
  const double *w = nullptr;
  if (h_bias) {
    TB_CUDA(cudaMemcpyAsync(s->b.p, h_bias, N * sizeof(double), cudaMemcpyHostToDevice, st));
    w = s->b.p;
  }
  WindowVisitor vis;
  vis.r0 = r0; vis.r1 = r1; vis.c0 = c0; vis.c1 = c1;
  vis.w = w;
  vis.out = out.p;
  visit_upper(s, vis, true);            // whole rows: the window may lie below the diagonal
  TB_CUDA(cudaMemcpyAsync(h_out, out.p, out.bytes(), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
}

// The stored cells of rows [r0, r1) x columns [c0, c1) of the full symmetric matrix, row-major, bad bins
// dropped: global (row, column), count, count / bias_i / bias_j (h_bias given) and that over
// h_decay[first bin of the section + |i - j|] (h_decay given; NaN between sections).  Call with null
// arrays to get *n.  A row-sharded store returns the cells of its own rows.
void region_cells(tb_store *s, int64_t r0, int64_t r1, int64_t c0, int64_t c1, const uint8_t *h_bad,
                  const double *h_bias, const double *h_decay, int half, int64_t *n, int32_t *h_row,
                  int32_t *h_col, int32_t *h_raw, double *h_nrm, double *h_dec) {
  cudaStream_t st = s->stream;
  const int64_t N = s->nbins, nrows = s->nrows;
  TB_REQUIRE(r0 >= 0 && r0 <= r1 && r1 <= N && c0 >= 0 && c0 <= c1 && c1 <= N && n, "region_cells: bad arguments");
  TB_REQUIRE(!h_decay || h_bias, "region_cells: the decay needs the biases");
  *n = 0;
  if (r0 == r1 || c0 == c1 || nrows == 0) return;
  if (h_bad) TB_CUDA(cudaMemcpyAsync(s->bad.p, h_bad, N, cudaMemcpyHostToDevice, st));
  else TB_CUDA(cudaMemsetAsync(s->bad.p, 0, N, st));
  DevBuf<unsigned long long> cnt;
  DevBuf<int64_t> off;
  cnt.alloc(nrows + 1);
  off.alloc(nrows + 1);
  TB_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), st));
  RegionCountVisitor cv;
  cv.r0 = r0; cv.r1 = r1; cv.c0 = c0; cv.c1 = c1; cv.rb = s->row_begin;
  cv.half = half != 0;
  cv.skip = s->bad.p;
  cv.cnt = cnt.p;
  visit_upper(s, cv, true);             // whole rows: the region may lie below the diagonal
  {
    size_t tmp_bytes = 0;
    TB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, reinterpret_cast<const int64_t *>(cnt.p), off.p,
                                          nrows + 1, st));
    DevBuf<uint8_t> tmp;
    tmp.alloc(tmp_bytes ? tmp_bytes : 1);
    TB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, reinterpret_cast<const int64_t *>(cnt.p), off.p,
                                          nrows + 1, st));
    TB_CUDA(cudaStreamSynchronize(st));
  }
  int64_t m = 0;
  TB_CUDA(cudaMemcpy(&m, off.p + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost));
  *n = m;
  if (!h_row || m == 0) return;
  TB_REQUIRE(h_col && h_raw && (!h_bias || h_nrm) && (!h_decay || h_dec), "region_cells: null output buffer");
  DevBuf<uint64_t> key_in, key_out;
  DevBuf<int32_t> raw_in, raw_out, orow, ocol;
  DevBuf<double> bias, decay, onrm, odec;
  DevBuf<int64_t> beg;
  key_in.alloc(m); key_out.alloc(m); raw_in.alloc(m); raw_out.alloc(m); orow.alloc(m); ocol.alloc(m);
  TB_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), st));
  RegionPlaceVisitor pv;
  pv.r0 = r0; pv.r1 = r1; pv.c0 = c0; pv.c1 = c1; pv.rb = s->row_begin;
  pv.half = half != 0;
  pv.skip = s->bad.p;
  pv.off = off.p;
  pv.cursor = cnt.p;
  pv.key = key_in.p;
  pv.raw = raw_in.p;
  visit_upper(s, pv, true);
  int row_bits = 1;
  while (((int64_t)1 << row_bits) < nrows) ++row_bits;
  size_t tmp_bytes = 0;
  TB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_in.p, key_out.p, raw_in.p, raw_out.p, m, 0,
                                          32 + row_bits, st));
  DevBuf<uint8_t> tmp;
  tmp.alloc(tmp_bytes ? tmp_bytes : 1);
  TB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key_in.p, key_out.p, raw_in.p, raw_out.p, m, 0,
                                          32 + row_bits, st));
  if (h_bias) {
    bias.alloc(N);
    onrm.alloc(m);
    TB_CUDA(cudaMemcpyAsync(bias.p, h_bias, N * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  if (h_decay) {
    decay.alloc(N);
    odec.alloc(m);
    const size_t nsec = s->chrom_offsets.size() - 1;
    beg.alloc(nsec);
    TB_CUDA(cudaMemcpyAsync(decay.p, h_decay, N * sizeof(double), cudaMemcpyHostToDevice, st));
    TB_CUDA(cudaMemcpyAsync(beg.p, s->chrom_offsets.data(), nsec * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  }
  k_unpack_region<<<grid_1d(m, 256, s->sm_count), 256, 0, st>>>(key_out.p, raw_out.p, m, s->row_begin, bias.p,
                                                               decay.p, s->chrom_of.p, beg.p, orow.p, ocol.p,
                                                               onrm.p, odec.p);
  TB_CUDA(cudaGetLastError());
  TB_CUDA(cudaMemcpyAsync(h_row, orow.p, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(h_col, ocol.p, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(h_raw, raw_out.p, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (h_bias) TB_CUDA(cudaMemcpyAsync(h_nrm, onrm.p, m * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (h_decay) TB_CUDA(cudaMemcpyAsync(h_dec, odec.p, m * sizeof(double), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace tb
