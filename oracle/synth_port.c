/* TEST INFRASTRUCTURE -- CPU generator of the synthetic Hi-C model used by bench.py's CPU legs.
 *
 * Same model as the device generator of the product: per-bin visibility
 * b_i ~ LogNormal(0, sigma), a fraction of bins scaled down or emptied; cis mean
 * a*b_i*b_j*(|i-j|+1)^-alpha, trans mean t*b_i*b_j; count ~ Poisson(mean); every cell drawn from
 * a counter-based hash of (seed, i, j).  Written independently in plain C so that the CPU baseline
 * (bench.py --impl reference, cpu_baseline) needs no GPU and none of the product's code; libm's
 * single-precision functions round differently from the device's, so individual cells may differ
 * from the device matrix -- the workload (size, density, value distribution) is the same.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  uint64_t seed;
  double cis_scale, cis_alpha, trans_mean, vis_sigma, low_frac, low_scale, empty_frac;
} synth_params;

static inline uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x7FEB352Du;
  h ^= h >> 15; h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}
static inline float unit(uint32_t h) { return ((float)(h >> 8) + 0.5f) * (1.0f / 16777216.0f); }

static const float kTwoPi = 6.28318530717958647692f;

static void visibility(const synth_params *p, int64_t n, float *vis) {
  const uint32_t lo = (uint32_t)(p->seed & 0xffffffffu), hi = (uint32_t)(p->seed >> 32);
  for (int64_t i = 0; i < n; ++i) {
    uint32_t h1 = mix32((uint32_t)i ^ lo);
    uint32_t h2 = mix32(h1 + hi + 0x9E3779B9u);
    uint32_t h3 = mix32(h2 ^ 0x85EBCA6Bu);
    uint32_t h4 = mix32(h3 + 0xC2B2AE35u);
    float z = sqrtf(-2.0f * logf(unit(h1))) * cosf(kTwoPi * unit(h2));
    float b = expf((float)p->vis_sigma * z);
    if (unit(h3) < (float)p->low_frac) b *= (float)p->low_scale;
    if (unit(h4) < (float)p->empty_frac) b = 0.0f;
    vis[i] = b;
  }
}

static inline int32_t cell_count(uint32_t seed_lo, uint32_t seed_hi, float trans_mean, int64_t lo, int64_t hi,
                                 const float *vis, const int32_t *chrom_of, const float *tab) {
  const float bb = vis[lo] * vis[hi];
  const float lam = chrom_of[lo] == chrom_of[hi] ? tab[hi - lo] * bb : trans_mean * bb;
  const uint32_t h = mix32(mix32((uint32_t)lo ^ seed_lo) + (uint32_t)hi * 0x9E3779B1u + seed_hi);
  const float u = unit(h);
  if (!(lam > 0.0f) || u >= lam) return 0;
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
  const float z = sqrtf(-2.0f * logf(u)) * cosf(kTwoPi * unit(h2));
  const float v = rintf(lam + sqrtf(lam) * z);
  return v > 0.0f ? (v < 2.0e9f ? (int32_t)v : 2000000000) : 0;
}

/* Upper-triangle cells (row <= col, row-major sorted) of the leading nbins x nbins block.
 * First call with rows == NULL returns the number of cells; the second call fills the arrays. */
int64_t synth_block(int64_t nbins, int32_t n_chrom, const int64_t *chrom_offsets, const synth_params *p,
                    int32_t *rows, int32_t *cols, int32_t *vals) {
  float *vis = (float *)malloc((size_t)nbins * sizeof(float));
  float *tab = (float *)malloc((size_t)nbins * sizeof(float));
  int32_t *chrom_of = (int32_t *)calloc((size_t)nbins, sizeof(int32_t));
  int64_t *count = (int64_t *)calloc((size_t)nbins + 1, sizeof(int64_t));
  if (!vis || !tab || !chrom_of || !count) { free(vis); free(tab); free(chrom_of); free(count); return -1; }
  visibility(p, nbins, vis);
  for (int64_t d = 0; d < nbins; ++d) tab[d] = (float)(p->cis_scale * pow((double)d + 1.0, -p->cis_alpha));
  for (int32_t c = 0; c < n_chrom; ++c)
    for (int64_t i = chrom_offsets[c]; i < chrom_offsets[c + 1] && i < nbins; ++i) chrom_of[i] = c;
  const uint32_t lo = (uint32_t)(p->seed & 0xffffffffu), hi = (uint32_t)(p->seed >> 32);
  const float tm = (float)p->trans_mean;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < nbins; ++i) {
    int64_t n = 0;
    if (vis[i] != 0.0f)
      for (int64_t j = i; j < nbins; ++j) n += cell_count(lo, hi, tm, i, j, vis, chrom_of, tab) > 0;
    count[i + 1] = n;
  }
  for (int64_t i = 0; i < nbins; ++i) count[i + 1] += count[i];
  const int64_t total = count[nbins];
  if (rows) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < nbins; ++i) {
      int64_t k = count[i];
      if (vis[i] != 0.0f)
        for (int64_t j = i; j < nbins; ++j) {
          const int32_t v = cell_count(lo, hi, tm, i, j, vis, chrom_of, tab);
          if (v > 0) { rows[k] = (int32_t)i; cols[k] = (int32_t)j; vals[k] = v; ++k; }
        }
    }
  }
  free(vis); free(tab); free(chrom_of); free(count);
  return total;
}
