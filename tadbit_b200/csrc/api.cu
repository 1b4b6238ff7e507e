// C-ABI of include/tadbit_b200.h: argument checks, exception -> status mapping.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "common.cuh"
#include "lanczos.hpp"

namespace tb {

static thread_local char g_error[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

namespace {
struct PoolState {
  bool ready = false, enabled = true;
  cudaStream_t stream = nullptr;
};
PoolState g_pool[64];

PoolState &pool_for_current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  PoolState &ps = g_pool[dev & 63];
  if (!ps.ready) {
    const char *e = getenv("TB_POOL");
    ps.enabled = !(e && e[0] == '0');
    if (ps.enabled) {
      cudaMemPool_t pool;
      uint64_t keep = UINT64_MAX;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess ||
          cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess ||
          cudaStreamCreateWithFlags(&ps.stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        ps.enabled = false;
      }
    }
    ps.ready = true;
  }
  return ps;
}
}  // namespace

void *dev_malloc(size_t bytes) {
  PoolState &ps = pool_for_current_device();
  void *p = nullptr;
  if (ps.enabled) {
    TB_CUDA(cudaMallocAsync(&p, bytes, ps.stream));
    TB_CUDA(cudaStreamSynchronize(ps.stream));       // usable from any stream afterwards
  } else {
    TB_CUDA(cudaMalloc(&p, bytes));
  }
  return p;
}

void dev_free(void *p) {
  PoolState &ps = pool_for_current_device();
  if (ps.enabled) {
    cudaDeviceSynchronize();                         // same guarantee as cudaFree: no user left
    cudaFreeAsync(p, ps.stream);
  } else {
    cudaFree(p);
  }
}

namespace {

template <typename F>
int guarded(F &&f) {
  try {
    return f();
  } catch (const Fail &e) {
    return e.code;
  } catch (const std::bad_alloc &) {
    set_error("host allocation failed");
    return TB_ERR_ARG;
  } catch (...) {
    set_error("unexpected C++ exception");
    return TB_ERR_ARG;
  }
}

tb_store *new_store(int64_t nbins, int32_t n_chrom, const int64_t *chrom_offsets,
                    int64_t row_begin, int64_t row_end, int device) {
  TB_REQUIRE(nbins > 0 && nbins < 2147483647LL, "nbins must be in [1, 2^31)");
  TB_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= nbins,
             "row block [%lld,%lld) not inside [0,%lld)", (long long)row_begin,
             (long long)row_end, (long long)nbins);
  TB_REQUIRE(n_chrom >= 0, "n_chrom must be >= 0");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("no CUDA device available (%s); tadbit_b200 has no CPU fallback",
              cudaGetErrorString(e));
    throw Fail{TB_ERR_CUDA};
  }
  TB_REQUIRE(device >= 0 && device < count, "device %d out of range (%d devices)", device, count);
  TB_CUDA(cudaSetDevice(device));
  tb_store *s = new tb_store();
  try {
    s->device = device;
    TB_CUDA(cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device));
    TB_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    s->nbins = nbins;
    s->row_begin = row_begin;
    s->row_end = row_end;
    s->nrows = row_end - row_begin;
    s->n_chrom = n_chrom;
    if (n_chrom > 0) {
      TB_REQUIRE(chrom_offsets, "chrom_offsets is null");
      s->chrom_offsets.assign(chrom_offsets, chrom_offsets + n_chrom + 1);
      TB_REQUIRE(s->chrom_offsets.front() == 0 && s->chrom_offsets.back() == nbins,
                 "chrom_offsets must start at 0 and end at nbins");
      for (int32_t c = 0; c < n_chrom; ++c)
        TB_REQUIRE(s->chrom_offsets[c] <= s->chrom_offsets[c + 1], "chrom_offsets not ascending");
    } else {
      s->chrom_offsets = {0, nbins};
    }
    setup_sections(s);
  } catch (...) {
    tb_store_destroy(s);
    throw;
  }
  return s;
}

struct DeviceGuard {          // every call runs on the store's device, then restores
  int prev = 0;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    cudaSetDevice(dev);
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

}  // namespace
}  // namespace tb

using namespace tb;

extern "C" {

int tb_version(void) { return 200 + (TB_CHECK ? 1000 : 0); }     /* +1000: the checked build */

const char *tb_last_error(void) { return g_error; }

int tb_trim_memory(int device) {
  int prev = 0;
  cudaGetDevice(&prev);
  if (cudaSetDevice(device) != cudaSuccess) { set_error("bad device"); return TB_ERR_ARG; }
  cudaDeviceSynchronize();
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
  cudaSetDevice(prev);
  return TB_OK;
}

int tb_device_count(int *count) {
  if (!count) return TB_ERR_ARG;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return TB_ERR_CUDA;
  }
  return TB_OK;
}

int tb_store_create_coo(int64_t nbins, int64_t nnz, const int32_t *row, const int32_t *col,
                        const int32_t *val, int mem, int32_t n_chrom,
                        const int64_t *chrom_offsets, int64_t row_begin, int64_t row_end,
                        int device, tb_store **out) {
  return guarded([&]() -> int {
    TB_REQUIRE(out, "out is null");
    *out = nullptr;
    TB_REQUIRE(nnz >= 0 && (nnz == 0 || (row && col && val)), "null cell arrays");
    TB_REQUIRE(mem == TB_MEM_HOST || mem == TB_MEM_DEVICE, "mem must be TB_MEM_HOST/TB_MEM_DEVICE");
    int prev = 0;
    cudaGetDevice(&prev);
    tb_store *s = new_store(nbins, n_chrom, chrom_offsets, row_begin, row_end, device);
    try {
      build_from_coo(s, nnz, row, col, val, mem);
    } catch (...) {
      tb_store_destroy(s);
      cudaSetDevice(prev);
      throw;
    }
    cudaSetDevice(prev);
    *out = s;
    return TB_OK;
  });
}

int tb_store_create_csr(int64_t nbins, int form, const int64_t *ptr, const int32_t *col,
                        const int32_t *val, int mem, int32_t n_chrom, const int64_t *chrom_offsets,
                        int64_t row_begin, int64_t row_end, int device, tb_store **out) {
  return guarded([&]() -> int {
    TB_REQUIRE(out && ptr, "null argument");
    *out = nullptr;
    TB_REQUIRE(form == TB_CELLS_UPPER || form == TB_CELLS_FULL_ROWS, "form must be TB_CELLS_UPPER/TB_CELLS_FULL_ROWS");
    TB_REQUIRE(mem == TB_MEM_HOST || mem == TB_MEM_DEVICE, "mem must be TB_MEM_HOST/TB_MEM_DEVICE");
    int prev = 0;
    cudaGetDevice(&prev);
    tb_store *s = new_store(nbins, n_chrom, chrom_offsets, row_begin, row_end, device);
    try {
      build_from_csr(s, form, ptr, col, val, mem);
    } catch (...) {
      tb_store_destroy(s);
      cudaSetDevice(prev);
      throw;
    }
    cudaSetDevice(prev);
    *out = s;
    return TB_OK;
  });
}

int tb_store_create_synthetic(int64_t nbins, int32_t n_chrom, const int64_t *chrom_offsets,
                              const tb_synth_params *params, int64_t row_begin,
                              int64_t row_end, int device, tb_store **out) {
  return guarded([&]() -> int {
    TB_REQUIRE(out && params, "null argument");
    *out = nullptr;
    int prev = 0;
    cudaGetDevice(&prev);
    tb_store *s = new_store(nbins, n_chrom, chrom_offsets, row_begin, row_end, device);
    try {
      build_synthetic(s, params);
    } catch (...) {
      tb_store_destroy(s);
      cudaSetDevice(prev);
      throw;
    }
    cudaSetDevice(prev);
    *out = s;
    return TB_OK;
  });
}

int tb_synthetic_row_counts(int64_t nbins, int32_t n_chrom, const int64_t *chrom_offsets,
                            const tb_synth_params *params, int64_t row_begin,
                            int64_t row_end, int device, int64_t *h_counts) {
  return guarded([&]() -> int {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
      set_error("no CUDA device available; tadbit_b200 has no CPU fallback");
      return TB_ERR_CUDA;
    }
    DeviceGuard g(device);
    synthetic_row_counts(nbins, n_chrom, chrom_offsets, params, row_begin, row_end, device,
                         h_counts);
    return TB_OK;
  });
}

int tb_store_create_batch(int32_t n, tb_store *const *members, int device, tb_store **out) {
  return guarded([&]() -> int {
    TB_REQUIRE(out && members && n > 0, "bad arguments");
    *out = nullptr;
    std::vector<int64_t> offs(1, 0);
    for (int32_t k = 0; k < n; ++k) {
      const tb_store *m = members[k];
      TB_REQUIRE(m, "null member store");
      TB_REQUIRE(m->device == device, "batch members must live on the batch's device");
      TB_REQUIRE(m->row_begin == 0 && m->row_end == m->nbins && m->world == 1,
                 "batch members must be whole (unsharded) matrices");
      offs.push_back(offs.back() + m->nbins);
    }
    int prev = 0;
    cudaGetDevice(&prev);
    tb_store *s = new_store(offs.back(), n, offs.data(), 0, offs.back(), device);
    try {
      build_batch(s, n, members);
    } catch (...) {
      tb_store_destroy(s);
      cudaSetDevice(prev);
      throw;
    }
    cudaSetDevice(prev);
    *out = s;
    return TB_OK;
  });
}

int tb_ice_batch(tb_store *s, const uint8_t *h_bad, int32_t iterations, double max_dev,
                 double *h_bias, double *h_trace, int32_t *n_passes) {
  return guarded([&]() -> int {
    TB_REQUIRE(s && h_bias && n_passes, "null argument");
    DeviceGuard g(s->device);
    ice_batch(s, h_bad, iterations, max_dev, h_bias, h_trace, n_passes);
    return TB_OK;
  });
}

int tb_store_destroy(tb_store *s) {
  if (!s) return TB_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(s->device);
  try {
    comm_destroy(s);
  } catch (...) {
  }
  if (s->stream2) {
    cudaStreamSynchronize(s->stream2);
    cudaStreamDestroy(s->stream2);
  }
  if (s->k1_fork) cudaEventDestroy(s->k1_fork);
  if (s->k1_join) cudaEventDestroy(s->k1_join);
  if (s->stream) {
    cudaStreamSynchronize(s->stream);
    cudaStreamDestroy(s->stream);
  }
  for (cudaEvent_t e : s->events) cudaEventDestroy(e);
  delete s;
  cudaSetDevice(prev);
  return TB_OK;
}

int tb_store_info(const tb_store *s, int64_t *nbins, int64_t *row_begin, int64_t *row_end,
                  int64_t *nnz_upper_owned, int64_t *nnz_full_owned, int64_t *device_bytes) {
  if (!s) { set_error("null store"); return TB_ERR_ARG; }
  if (nbins) *nbins = s->nbins;
  if (row_begin) *row_begin = s->row_begin;
  if (row_end) *row_end = s->row_end;
  if (nnz_upper_owned) *nnz_upper_owned = s->nnz_upper;
  if (nnz_full_owned) *nnz_full_owned = s->nnz_full;
  if (device_bytes) *device_bytes = s->device_bytes;
  return TB_OK;
}

int tb_store_set_option(tb_store *s, const char *name, const char *value) {
  if (!s || !name || !value) { set_error("null argument"); return TB_ERR_ARG; }
  if (!strcmp(name, "rowsum")) {
    if (!strcmp(value, "csr")) s->rowsum_mode = 1;
    else if (!strcmp(value, "compact")) s->rowsum_mode = 0;
    else { set_error("rowsum must be 'compact' or 'csr'"); return TB_ERR_ARG; }
    if (s->rowsum_mode == 1 && !s->csr_resident) {
      s->rowsum_mode = -1;
      set_error("rowsum=csr: this store released its CSR after the build (create it with TB_KEEP_CSR=1)");
      return TB_ERR_STATE;
    }
    return TB_OK;
  }
  if (!strcmp(name, "exchange")) {       // every rank of a group must choose the same
    if (!strcmp(value, "nccl")) s->exchange_mode = 0;
    else if (!strcmp(value, "fused")) s->exchange_mode = 1;
    else { set_error("exchange must be 'fused' or 'nccl'"); return TB_ERR_ARG; }
    return TB_OK;
  }
  if (!strcmp(name, "k1_streams")) {     // K1's two kernels back to back (1) or on two streams (2)
    if (!strcmp(value, "1")) s->k1_streams_mode = 1;
    else if (!strcmp(value, "2")) s->k1_streams_mode = 2;
    else { set_error("k1_streams must be '1' or '2'"); return TB_ERR_ARG; }
    return TB_OK;
  }
  if (!strcmp(name, "corr_gemm")) {      // X X^T of tb_corr_build: fp64 tensor tiles, or the plain parity kernel
    if (!strcmp(value, "dmma")) s->corr_gemm_mode = 0;
    else if (!strcmp(value, "simple")) s->corr_gemm_mode = 1;
    else { set_error("corr_gemm must be 'dmma' or 'simple'"); return TB_ERR_ARG; }
    return TB_OK;
  }
  set_error("unknown option %s", name);
  return TB_ERR_ARG;
}

int tb_store_layout(const tb_store *s, int64_t out[16]) {
  if (!s || !out) { set_error("null argument"); return TB_ERR_ARG; }
  for (int k = 0; k < 16; ++k) out[k] = 0;
  out[0] = s->enc_bytes;
  for (int k = 0; k < 6; ++k) out[1 + k] = s->enc_chunks[k];
  out[7] = s->nchunks;
  out[8] = (int64_t)(s->col.bytes() + s->val.bytes() + s->rowptr.bytes());
  out[9] = s->t_ncells;
  out[10] = s->t_npadded;
  out[11] = s->t_nunits;
  out[12] = s->n_overflow;
  out[14] = s->n_overflow_tiled;
  out[13] = s->t_nunits ? (int64_t)(s->t_npadded * 4 + s->t_perm.bytes() + s->t_desc.bytes() + s->t_desc_off.bytes()) : 0;
  return TB_OK;
}

int tb_store_export_upper(const tb_store *s, int64_t *nnz, int32_t *h_row, int32_t *h_col,
                          int32_t *h_val) {
  return guarded([&]() -> int {
    TB_REQUIRE(s && nnz, "null argument");
    DeviceGuard g(s->device);
    export_upper(s, nnz, h_row, h_col, h_val);
    return TB_OK;
  });
}

int tb_store_export_csr(const tb_store *s, int form, int64_t *nnz, int64_t *h_ptr, int32_t *h_col,
                        int32_t *h_val) {
  return guarded([&]() -> int {
    TB_REQUIRE(s && nnz, "null argument");
    TB_REQUIRE(form == TB_CELLS_UPPER || form == TB_CELLS_FULL_ROWS, "form must be TB_CELLS_UPPER/TB_CELLS_FULL_ROWS");
    DeviceGuard g(s->device);
    export_csr(s, form, nnz, h_ptr, h_col, h_val);
    return TB_OK;
  });
}

int tb_comm_unique_id(uint8_t id_out[128]) {
  return guarded([&]() -> int {
    TB_REQUIRE(id_out, "null argument");
    comm_unique_id(id_out);
    return TB_OK;
  });
}

int tb_comm_init(tb_store *s, int32_t rank, int32_t world, const uint8_t id[128]) {
  return guarded([&]() -> int {
    TB_REQUIRE(s && (world == 1 || id), "null argument");
    DeviceGuard g(s->device);
    comm_init(s, rank, world, id);
    return TB_OK;
  });
}

int tb_comm_create(int32_t rank, int32_t world, const uint8_t id[128], int device, tb_comm **out) {
  return guarded([&]() -> int {
    TB_REQUIRE(out && (world == 1 || id), "null argument");
    *out = nullptr;
    DeviceGuard g(device);
    *out = group_create(rank, world, id, device);
    return TB_OK;
  });
}

int tb_comm_destroy(tb_comm *c) {
  if (!c) return TB_OK;
  DeviceGuard g(c->device);
  group_destroy(c);
  return TB_OK;
}

int tb_store_attach(tb_store *s, tb_comm *c) {
  return guarded([&]() -> int {
    TB_REQUIRE(s && c, "null argument");
    DeviceGuard g(s->device);
    store_attach(s, c, false);
    return TB_OK;
  });
}

int tb_row_stats(tb_store *s, int64_t *h_nnz_row, int64_t *h_row_sum) {
  return guarded([&]() -> int {
    TB_REQUIRE(s && h_nnz_row && h_row_sum, "null argument");
    DeviceGuard g(s->device);
    row_stats(s, h_nnz_row, h_row_sum);
    return TB_OK;
  });
}

int tb_col_sums_masked(tb_store *s, const uint8_t *h_bad, int64_t *h_col_sum) {
  return guarded([&]() -> int {
    TB_REQUIRE(s && h_col_sum, "null argument");
    DeviceGuard g(s->device);
    col_sums_masked(s, h_bad, h_col_sum);
    return TB_OK;
  });
}

int tb_ice(tb_store *s, const uint8_t *h_bad, int32_t iterations, double max_dev,
           double *h_bias, double *h_trace, int32_t *n_passes) {
  return guarded([&]() -> int {
    TB_REQUIRE(s && h_bias, "null argument");
    DeviceGuard g(s->device);
    return ice(s, h_bad, iterations, max_dev, h_bias, h_trace, n_passes);
  });
}

int tb_probe_rowsum(tb_store *s, int32_t reps, double ms[3]) {
  return guarded([&]() -> int {
    TB_REQUIRE(s, "null argument");
    DeviceGuard g(s->device);
    probe_rowsum(s, reps, ms);
    return TB_OK;
  });
}

int tb_probe_k1(tb_store *s, int32_t reps, double ms[2]) {
  return guarded([&]() -> int {
    TB_REQUIRE(s, "null argument");
    DeviceGuard g(s->device);
    probe_k1(s, reps, ms);
    return TB_OK;
  });
}

int tb_row_work(tb_store *s, int64_t *h_dense_bytes, int64_t *h_tile_cells) {
  return guarded([&]() -> int {
    TB_REQUIRE(s && h_dense_bytes && h_tile_cells, "null argument");
    DeviceGuard g(s->device);
    row_work(s, h_dense_bytes, h_tile_cells);
    return TB_OK;
  });
}

int tb_ice_last_timing(const tb_store *s, double *loop_ms, double *rowsum_ms,
                       int32_t *rowsum_launches, int32_t *total_launches) {
  if (!s) { set_error("null store"); return TB_ERR_ARG; }
  if (loop_ms) *loop_ms = s->loop_ms;
  if (rowsum_ms) *rowsum_ms = s->rowsum_ms;
  if (rowsum_launches) *rowsum_launches = s->rowsum_launches;
  if (total_launches) *total_launches = s->total_launches;
  return TB_OK;
}

int tb_last_timing(const tb_store *s, double out[8]) {
  if (!s || !out) { set_error("null argument"); return TB_ERR_ARG; }
  out[0] = s->loop_ms;
  out[1] = s->rowsum_ms;
  out[2] = s->exchange_ms;
  out[3] = s->last_kernel_ms;
  out[4] = (double)s->rowsum_launches;
  out[5] = (double)s->total_launches;
  out[6] = s->px.mapped ? 2.0 : (s->px.ready ? 1.0 : 0.0);
  out[7] = s->csr_resident ? 1.0 : 0.0;
  return TB_OK;
}

int tb_norm_sum(tb_store *s, const double *h_bias, const uint8_t *h_bad, double *out) {
  return guarded([&]() -> int {
    TB_REQUIRE(s && out, "null argument");
    DeviceGuard g(s->device);
    norm_sum(s, h_bias, h_bad, out);
    return TB_OK;
  });
}

int tb_pair_stats(tb_store *s, const double *h_bias, const uint8_t *h_bad, double center, double out[4]) {
  return guarded([&]() -> int {
    TB_REQUIRE(s && out, "null argument");
    DeviceGuard g(s->device);
    pair_stats(s, h_bias, h_bad, center, out);
    return TB_OK;
  });
}

int tb_pair_values(tb_store *s, const double *h_bias, const uint8_t *h_bad, int32_t zscored, double mean,
                   double std, int64_t *n, int32_t *h_row, int32_t *h_col, double *h_val) {
  return guarded([&]() -> int {
    TB_REQUIRE(s && n, "null argument");
    TB_REQUIRE(!zscored || std > 0.0, "z-scores need a positive standard deviation");
    DeviceGuard g(s->device);
    pair_values(s, h_bias, h_bad, zscored, mean, std, n, h_row, h_col, h_val);
    return TB_OK;
  });
}

int tb_decay_sums(tb_store *s, const double *h_bias, const uint8_t *h_bad, double *h_nrm, int64_t *h_raw,
                  double out_cis_total[2]) {
  return guarded([&]() -> int {
    TB_REQUIRE(s, "null argument");
    DeviceGuard g(s->device);
    decay_sums(s, h_bias, h_bad, h_nrm, h_raw, out_cis_total);
    return TB_OK;
  });
}

int tb_window(tb_store *s, int64_t row_begin, int64_t row_end, int64_t col_begin, int64_t col_end,
              const double *h_bias, double *h_out) {
  return guarded([&]() -> int {
    TB_REQUIRE(s, "null argument");
    DeviceGuard g(s->device);
    window(s, row_begin, row_end, col_begin, col_end, h_bias, h_out);
    return TB_OK;
  });
}

int tb_region_cells(tb_store *s, int64_t row_begin, int64_t row_end, int64_t col_begin, int64_t col_end,
                    const uint8_t *h_bad, const double *h_bias, const double *h_decay, int32_t half, int64_t *n,
                    int32_t *h_row, int32_t *h_col, int32_t *h_raw, double *h_nrm, double *h_dec) {
  return guarded([&]() -> int {
    TB_REQUIRE(s && n, "null argument");
    DeviceGuard g(s->device);
    region_cells(s, row_begin, row_end, col_begin, col_end, h_bad, h_bias, h_decay, half, n, h_row, h_col,
                 h_raw, h_nrm, h_dec);
    return TB_OK;
  });
}

int tb_write_abc_lines(const char *path, int64_t n, const int32_t *h_a, const int32_t *h_b, const int32_t *h_int,
                       const double *h_val) {
  return guarded([&]() -> int {
    write_abc_lines(path, n, h_a, h_b, h_int, h_val);
    return TB_OK;
  });
}

int tb_write_pair_lines(const char *path, int64_t n, const int64_t *h_a, const int64_t *h_b, const int64_t *h_int,
                        const double *h_val, const char *sep, const char *tail) {
  return guarded([&]() -> int {
    write_pair_lines(path, n, h_a, h_b, h_int, h_val, sep, tail);
    return TB_OK;
  });
}

int tb_pool_expected(int64_t size, int64_t ndiag, const int64_t *h_sum_d, const int64_t *h_cells_d, double min_n,
                     double *h_out, int64_t *nkeys) {
  return guarded([&]() -> int {
    pool_expected(size, ndiag, h_sum_d, h_cells_d, min_n, h_out, nkeys);
    return TB_OK;
  });
}

int tb_pool_decay(int64_t size, const double *h_nrm, const int64_t *h_raw, const int64_t *h_ndiag, double min_n,
                  double *h_out, uint8_t *h_state, int32_t *empty) {
  return guarded([&]() -> int {
    pool_decay(size, h_nrm, h_raw, h_ndiag, min_n, h_out, h_state, empty);
    return TB_OK;
  });
}

int tb_corr_build(tb_store *s, int64_t bin_begin, int64_t bin_end, const double *h_bias, const uint8_t *h_bad,
                  const double *h_expected, int32_t stage, int64_t *n_good) {
  return guarded([&]() -> int {
    TB_REQUIRE(s, "null argument");
    DeviceGuard g(s->device);
    corr_build(s, bin_begin, bin_end, h_bias, h_bad, h_expected, stage, n_good);
    return TB_OK;
  });
}

int tb_corr_get(tb_store *s, double *h_out) {
  return guarded([&]() -> int {
    TB_REQUIRE(s, "null argument");
    DeviceGuard g(s->device);
    corr_get(s, h_out);
    return TB_OK;
  });
}

int tb_corr_set(tb_store *s, int64_t n, const double *h_in) {
  return guarded([&]() -> int {
    TB_REQUIRE(s, "null argument");
    DeviceGuard g(s->device);
    corr_set(s, n, h_in);
    return TB_OK;
  });
}

int tb_corr_eig(tb_store *s, int32_t k, double *h_evals, double *h_evecs, int32_t info[8]) {
  return guarded([&]() -> int {
    TB_REQUIRE(s, "null argument");
    DeviceGuard g(s->device);
    corr_eig(s, k, h_evals, h_evecs, info);
    return TB_OK;
  });
}

int tb_corr_timing(const tb_store *s, double out[8]) {
  if (!s || !out) { set_error("null argument"); return TB_ERR_ARG; }
  for (int k = 0; k < 5; ++k) out[k] = s->corr_ms[k];
  out[5] = (double)s->dense_n;
  out[6] = (double)s->dense_stage;
  out[7] = (double)s->corr_gemm_mode;
  return TB_OK;
}

int tb_tridiag_eig(int32_t n, double *diag, const double *sub, double *vecs) {
  return guarded([&]() -> int {
    TB_REQUIRE(n >= 1 && diag && vecs && (sub || n == 1), "null argument");
    std::vector<double> d(diag, diag + n), e(n, 0.0), z;
    for (int i = 0; i + 1 < n; ++i) e[i] = sub[i];
    if (!tridiag_eig(n, d, e, z)) { set_error("tridiag_eig: no convergence"); return TB_ERR_STATE; }
    for (int i = 0; i < n; ++i) diag[i] = d[i];
    for (size_t i = 0; i < z.size(); ++i) vecs[i] = z[i];
    return TB_OK;
  });
}

int tb_diag_sums(tb_store *s, const uint8_t *h_bad, int64_t ndiag, int64_t *h_sum_d) {
  return guarded([&]() -> int {
    TB_REQUIRE(s, "null argument");
    DeviceGuard g(s->device);
    diag_sums(s, h_bad, ndiag, h_sum_d);
    return TB_OK;
  });
}

}  // extern "C"
