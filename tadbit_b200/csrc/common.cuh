// Shared declarations of the tadbit_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/tadbit_b200.h"

namespace tb {

void set_error(const char *fmt, ...);
struct Fail {           // thrown inside the library, mapped to a tb_status at the C-ABI
  int code;
};

#define TB_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      ::tb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                     \
                      cudaGetErrorString(e__));                                        \
      throw ::tb::Fail{TB_ERR_CUDA};                                                   \
    }                                                                                  \
  } while (0)

#define TB_REQUIRE(cond, ...)                                                          \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      ::tb::set_error(__VA_ARGS__);                                                    \
      throw ::tb::Fail{TB_ERR_ARG};                                                    \
    }                                                                                  \
  } while (0)

// Device memory comes from the device's stream-ordered pool with the release threshold
// lifted (api.cu): building a store needs tens of GB of short-lived scratch, and handing
// that back to the driver and mapping it again dominated repeated store creation.
// tb_trim_memory() returns the cached memory.  TB_POOL=0 falls back to cudaMalloc/cudaFree.
void *dev_malloc(size_t bytes);
void dev_free(void *p);

// Device allocation owned by a store; freed in its destructor.
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) p = static_cast<T *>(dev_malloc(count * sizeof(T)));
  }
  void release() {
    if (p) dev_free(p);
    p = nullptr;
    n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

template <typename T>
struct PinnedBuf {
  T *p = nullptr;
  size_t n = 0;
  ~PinnedBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) TB_CUDA(cudaMallocHost(&p, count * sizeof(T)));
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
  }
};

// Device-side state of one ICE call (written by the reduce kernel, polled by the host).
struct IceState {
  double mean_s, s_min, s_max, dev;
  long long n_present;
  int passes_done;   // passes whose statistics were finalised
  int stopped;       // 1 once the reference loop would have left (dev < max_dev, or iterations == 0)
  int all_bad;       // no present row
  unsigned int blocks_arrived;
  int zero_div;      // meanS == 0: the reference raises ZeroDivisionError in _updateDB (normalize_hic.py:148)
  int fault;         // 1 = a peer did not arrive at the exchange in time (k_ice_exchange)
  unsigned int arrive1, arrive2;          // grid barriers of k_ice_exchange (reset by their last arriver)
  unsigned long long fold_epoch;          // last exchange whose statistics are final
  unsigned long long epoch_base;          // exchange number of pass 0 of the current call, minus 1
};

// Row-sum exchange between the GPUs of one box without a collective: every rank holds two
// full-length buffers (alternating by pass) that every peer maps through CUDA IPC and writes
// its own rows into, plus one arrival flag per peer (comm.cu sets this up, k_ice_exchange in
// kernels.cu uses it).  With one GPU the "peers" are the rank itself and nothing is mapped.
constexpr int kMaxPeers = 16;
struct PeerView {                  // passed to the kernel by value; indexed with constants only
  double *s[kMaxPeers][2];         // peer q's buffers (s[rank] = local)
  unsigned long long *flag[kMaxPeers];   // peer q's flag array; this rank writes flag[q][rank]
  unsigned long long *my_flags;    // local flag array, entry q written by peer q
  double *mine[2];                 // the local buffers again (no dynamic index into s[][])
  int rank, world;
};
struct PeerExchange {
  bool ready = false;              // buffers allocated (world == 1: local only)
  bool mapped = false;             // peers' buffers opened (world > 1)
  void *base = nullptr;            // local allocation: 2 x stride doubles, then kMaxPeers flags
  size_t stride = 0;               // doubles per buffer
  void *peer_base[kMaxPeers] = {};
  unsigned long long epoch = 0;    // exchanges issued so far (same on every rank)
  PeerView view = {};
};

// -DTB_CHECK=1: a debug build with bounds assertions on every descriptor / ring / gather index of
// the hot kernels (compute-sanitizer is not available on the GPU pool).  A failed assertion prints
// its source line and traps, which the host sees as a CUDA error.  tools/kernel_coverage.py runs
// under this build in the GPU tests (tadbit_b200/variants/check.so, built by __graft_entry__.build()).
#ifndef TB_CHECK
#define TB_CHECK 0
#endif
#if TB_CHECK
#include <stdio.h>
#define TB_DCHECK(cond)                                                                        \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      printf("TB_CHECK failed %s:%d: %s (block %d thread %d)\n", __FILE__, __LINE__, #cond,    \
             (int)blockIdx.x, (int)threadIdx.x);                                               \
      __trap();                                                                                \
    }                                                                                          \
  } while (0)
#else
#define TB_DCHECK(cond) do { } while (0)
#endif
struct CheckLimits {               // sizes the assertions compare against (one copy per translation unit)
  long long nbins, nrows, nchunks, enc_bytes, n_groups;
  long long t_cell_words, t_ntiles, t_nunits, t_ndesc, t_nlist;
};
#if TB_CHECK
static __constant__ CheckLimits c_lim;
#endif

constexpr int kWarp = 32;
constexpr int kChunk = 4096;      // max stored cells one warp reduces before writing a partial
constexpr int kPanel = 4096;      // column panel of the chunk planner / dense encodings
#ifndef TB_KSUB
#define TB_KSUB 1024
#endif
constexpr int kSub = TB_KSUB;     // bytes per row per pipeline stage of k_rowsum_dense
constexpr int kRows = 4;          // dense chunks (rows) one warp walks together
constexpr int kXPad = 16384;      // zero tail of the gathered vector (panels / windows read past N)
#ifndef TB_TILE_ROWS
#define TB_TILE_ROWS 4096                 // measured on the genome-wide 5 kb matrix: tile kernel 1.067 ms with
#endif                                    // 4096 rows per tile (229.8 KB of shared memory), 1.163 ms with 2048
constexpr int kTileRows = TB_TILE_ROWS;   // rows of a sparse tile (2048 or 4096: per-unit accumulators in shared memory)
constexpr int kTileCols = 4 * kPanel;  // columns of a sparse tile = x window in shared memory (128 KB)
constexpr int kTileValBits = 15;  // packed tile cell: (column offset * 4) << 16 | value, value < 2^15

// chunk encodings of the compact row-sum stream; ENC_T = the cells live in the sparse tiles
enum { ENC_S32 = 0, ENC_S16 = 1, ENC_D32 = 2, ENC_D16 = 3, ENC_D8 = 4, ENC_T = 5 };

// A group of kRows dense chunks that share (encoding, column panel): the unit of work of a
// warp of k_rowsum_dense.  Built on the host in encode.cu.
struct GroupRec {
  uint32_t off16[kRows];   // payload offset / 16 of each seat
  int32_t id[kRows];       // chunk id (index of its partial), -1 = empty seat
  int32_t c0;              // first column of the panel
  int32_t nsub;            // kSub-byte sub-blocks per row
  int32_t enc;             // ENC_D32 / ENC_D16 / ENC_D8 | (some seat carries overflow cells) << 8
  int32_t seg_end;         // first group after this one that sits in another panel
  int32_t novf[kRows];     // overflow cells of each seat
};
static_assert(sizeof(GroupRec) == 64, "GroupRec is read as four int4");
#ifndef TB_DENSE_WARPS
#define TB_DENSE_WARPS 24
#endif
constexpr int kDenseWarps = TB_DENSE_WARPS;   // warps per CTA of k_rowsum_dense (one CTA per SM)

}  // namespace tb

// The opaque handles of include/tadbit_b200.h.
struct tb_comm {
  void *nccl = nullptr;     // ncclComm_t (null when world == 1)
  int rank = 0, world = 1, device = 0;
};

struct tb_store {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;         // K1: the dense kernel runs here next to the tile kernel (rowsum.cu)
  cudaEvent_t k1_fork = nullptr, k1_join = nullptr;
  int k1_streams_mode = -1;               // -1 default (TB_K1_STREAMS, else 2), 1 serial, 2 two streams
  int sm_count = 148;
  int64_t nbins = 0, row_begin = 0, row_end = 0, nrows = 0;
  int64_t nnz_full = 0;     // stored full-matrix cells with row in the block
  int64_t nnz_upper = 0;    // of those, col >= row
  // sections
  int32_t n_chrom = 0;
  std::vector<int64_t> chrom_offsets;     // n_chrom+1 (always >= 2 entries: one section if none given)
  tb::DevBuf<int32_t> chrom_of;           // [N] section id of each bin
  tb::DevBuf<int64_t> chrom_end;          // [sections] one past the last bin of each section
  // CSR of the owned rows of the FULL symmetric matrix, columns ascending
  tb::DevBuf<int64_t> rowptr;             // [nrows+1]
  tb::DevBuf<int32_t> col, val;           // [nnz_full]
  // warp work items: <= kChunk consecutive cells of one row
  int64_t nchunks = 0;
  tb::DevBuf<int64_t> chunk_beg;          // [nchunks+1] chunk c = cells [chunk_beg[c], chunk_beg[c+1])
  tb::DevBuf<int32_t> chunk_row;          // [nchunks] local row
  tb::DevBuf<int64_t> row_chunk;          // [nrows+1] first chunk of each row
  // compact row-sum stream (encode.cu): one record per chunk, same chunk order as above
  tb::DevBuf<uint8_t> enc_data;           // packed chunk payloads, each 16-byte aligned
  tb::DevBuf<int64_t> enc_off;            // [nchunks] payload offset
  tb::DevBuf<int4> enc_desc;              // [nchunks] (first column, slots or cells, encoding, 0)
  tb::DevBuf<int32_t> work_ids;           // [nchunks] chunk ids in work order: sparse, then dense
  int64_t n_sparse = 0, n_dense = 0;
  tb::DevBuf<tb::GroupRec> groups;        // [n_groups] kRows dense chunks of one panel each
  int64_t n_groups = 0;
  tb::DevBuf<int32_t> group_range;        // [n_dense_ctas + 1] contiguous groups of each k_rowsum_dense CTA
  int n_dense_ctas = 0;
  int64_t n_overflow = 0;                 // cells kept in a list behind their dense chunk (value out of range)
  tb::DevBuf<uint16_t> enc_tovf;          // [nchunks] cells of a D8 chunk (255 < v < 2^15) moved to the tiles
  int64_t n_overflow_tiled = 0;
  int64_t enc_bytes = 0;
  int64_t enc_chunks[6] = {0, 0, 0, 0, 0, 0};  // chunks per encoding
  // sparse tiles (tiles.cu): cells of ENC_T chunks, kTileRows x kTileCols tiles, rows of a tile
  // sorted by length and stored lane-interleaved in slices of 32 rows (sliced ELL)
  int64_t t_nrb = 0, t_nwin = 0;          // row blocks, column windows
  int64_t t_ncells = 0, t_npadded = 0;    // real cells / stored cells incl. padding
  tb::DevBuf<uint32_t> t_cells;           // packed cells, tile after tile
  tb::DevBuf<int64_t> t_off;              // [t_nrb*t_nwin + 1] first cell of each tile
  tb::DevBuf<uint16_t> t_perm;            // [tiles * kTileRows] local row handled by thread t
  tb::DevBuf<uint32_t> t_slice;           // [tiles * 65] first cell of each 32-row slice in the tile
  tb::DevBuf<int32_t> t_list;             // non-empty tile ids, grouped by work unit
  tb::DevBuf<int32_t> t_unit;             // [t_nunits * 4] (row block, list begin, list end, part<<8|parts)
  tb::DevBuf<int32_t> t_unit_of_rb;       // [t_nrb + 1] units of a row block are contiguous
  tb::DevBuf<int32_t> t_order;            // [t_nunits] unit ids, most expensive first (fetch order)
  tb::DevBuf<uint2> t_desc;               // copy descriptors: per (unit, warp) the pieces of its cell stream
  tb::DevBuf<int64_t> t_desc_off;         // [t_nunits * 32 + 1]
  tb::DevBuf<unsigned int> t_counter;     // [2] unit fetch counters, used alternately by successive launches
  int t_launches = 0;
  int64_t t_nunits = 0;
  tb::DevBuf<double> t_partial;           // [t_nunits * kTileRows] per-unit row sums
  // scratch
  tb::DevBuf<double> partial_f;           // [nchunks]
  tb::DevBuf<long long> partial_i;        // [nchunks]
  tb::DevBuf<double> x, b, s_part, s_full, block_red;
  tb::DevBuf<long long> i_part, i_full;   // [2N] integer reductions
  tb::DevBuf<double> scal_f;              // [16] scalar results
  tb::DevBuf<long long> scal_i;           // [16]
  tb::DevBuf<uint8_t> bad, present;
  tb::DevBuf<tb::IceState> state;
  tb::DevBuf<double> trace;               // [(iterations+1)*4]
  tb::PinnedBuf<tb::IceState> h_state;
  // what survives of the CSR when it is released after the build (positive matrices)
  bool csr_resident = true;               // rowptr/col/val/chunk_beg still hold the cells
  bool positive = false;                  // every stored value > 0 and the total < 2^53
  int64_t n_zero_cells = 0;               // stored cells with value 0
  tb::DevBuf<long long> row_nnz, row_sum; // [nrows] stored cells / sum of each owned full row (exact)
  tb::DevBuf<double> live;                // [nrows] cells with a good partner (stores that kept the CSR)
  // multi-GPU
  void *comm = nullptr;                   // the group's ncclComm_t
  tb_comm *group = nullptr;
  bool owns_group = false;                // created by tb_comm_init: destroyed with the store
  int rank = 0, world = 1;
  tb::PeerExchange px;                    // peer-mapped row-sum buffers (comm.cu)
  std::vector<cudaEvent_t> events;        // reused by every ICE call (kernels.cu)
  // timing of the last ICE call
  double loop_ms = 0, rowsum_ms = 0, exchange_ms = 0;
  double last_kernel_ms = 0;              // device time of the last one-pass reduction (K4 / K5)
  int rowsum_launches = 0, total_launches = 0;
  int64_t device_bytes = 0;
  // N4 (compartments.cu): the dense matrix of one section over its good bins, resident between calls
  tb::DevBuf<double> dense;               // [dense_rows x dense_ld], zero outside n x n
  int64_t dense_n = 0, dense_ld = 0, dense_rows = 0;
  int dense_stage = -1;                   // 0 observed / expected, 1 correlation, 2 uploaded by the host
  int corr_gemm_mode = 0;                 // 0 DMMA tiles, 1 plain kernel (tb_store_set_option corr_gemm)
  double corr_ms[5] = {0, 0, 0, 0, 0};    // visitor, centring, X X^T, scaling, eigen-solver
  int rowsum_mode = -1;     // -1 default (compact stream), 0 compact, 1 CSR walk (tb_store_set_option)
  int exchange_mode = -1;   // -1 default (TB_EXCHANGE), 1 fused exchange kernel, 0 separate steps / NCCL
};

namespace tb {
#if TB_CHECK
inline void set_check_limits(const tb_store *s) {   // before a checked launch, in that kernel's translation unit
  CheckLimits h;
  h.nbins = s->nbins; h.nrows = s->nrows; h.nchunks = s->nchunks; h.enc_bytes = s->enc_bytes;
  h.n_groups = s->n_groups;
  h.t_cell_words = (long long)s->t_cells.n; h.t_ntiles = s->t_nrb * s->t_nwin; h.t_nunits = s->t_nunits;
  h.t_ndesc = (long long)s->t_desc.n; h.t_nlist = (long long)s->t_list.n;
  cudaMemcpyToSymbolAsync(c_lim, &h, sizeof(h), 0, cudaMemcpyHostToDevice, s->stream);
}
#else
inline void set_check_limits(const tb_store *) {}
#endif
// build.cu
void build_from_coo(tb_store *s, int64_t nnz, const int32_t *row, const int32_t *col,
                    const int32_t *val, int mem);
int dense_fill_factor();            // see build.cu
void setup_sections(tb_store *s);   // bin -> section id table (before any build)
void finish_layout(tb_store *s);     // chunks + scratch once rowptr/col/val are in place
void export_upper(const tb_store *s, int64_t *nnz, int32_t *h_row, int32_t *h_col, int32_t *h_val);
void export_upper_device(const tb_store *s, int32_t *d_row, int32_t *d_col, int32_t *d_val, int32_t bin_off);
void export_csr(const tb_store *s, int form, int64_t *nnz, int64_t *h_ptr, int32_t *h_col, int32_t *h_val);
void build_from_csr(tb_store *s, int form, const int64_t *ptr, const int32_t *col, const int32_t *val, int mem);
void build_batch(tb_store *s, int32_t n, tb_store *const *members);   // block-diagonal concatenation
// encode.cu
void build_encoded(tb_store *s);     // compact row-sum stream from the CSR
// rowsum.cu
void launch_rowsum(tb_store *s, const double *x, const IceState *state);   // K1 over the compact stream
void launch_rowsum_chunks(tb_store *s, const double *x, const IceState *state);   // dense + per-cell part
bool sparse_in_dense_tail(const tb_store *s);
// tiles.cu
void build_tiles(tb_store *s);       // sparse tiles from the ENC_T chunks (after build_encoded)
void launch_rowsum_tiles(tb_store *s, const double *x, const IceState *state, unsigned long long *prof = nullptr);
// synth.cu
void build_synthetic(tb_store *s, const tb_synth_params *p);
void synthetic_row_counts(int64_t nbins, int32_t n_chrom, const int64_t *chrom_offsets,
                          const tb_synth_params *p, int64_t rb, int64_t re, int device,
                          int64_t *h_counts);
// kernels.cu
void row_stats(tb_store *s, int64_t *h_nnz_row, int64_t *h_row_sum);
void col_sums_masked(tb_store *s, const uint8_t *h_bad, int64_t *h_col_sum);
int ice(tb_store *s, const uint8_t *h_bad, int32_t iterations, double max_dev, double *h_bias,
        double *h_trace, int32_t *n_passes);
void norm_sum(tb_store *s, const double *h_bias, const uint8_t *h_bad, double *out);
void probe_rowsum(tb_store *s, int32_t reps, double ms[3]);
void probe_k1(tb_store *s, int32_t reps, double ms[2]);
void row_work(tb_store *s, int64_t *h_dense, int64_t *h_tile);
void ice_batch(tb_store *s, const uint8_t *h_bad, int32_t iterations, double max_dev, double *h_bias,
               double *h_trace, int32_t *n_passes);
void diag_sums(tb_store *s, const uint8_t *h_bad, int64_t ndiag, int64_t *h_sum_d);
// extras.cu
void pair_stats(tb_store *s, const double *h_bias, const uint8_t *h_bad, double center, double out[4]);
void pair_values(tb_store *s, const double *h_bias, const uint8_t *h_bad, int zscored, double mean, double std,
                 int64_t *n, int32_t *h_row, int32_t *h_col, double *h_val);
void decay_sums(tb_store *s, const double *h_bias, const uint8_t *h_bad, double *h_nrm, int64_t *h_raw,
                double out[2]);
void window(tb_store *s, int64_t r0, int64_t r1, int64_t c0, int64_t c1, const double *h_bias, double *h_out);
void write_abc_lines(const char *path, int64_t n, const int32_t *a, const int32_t *b, const int32_t *ival,
                     const double *fval);
void write_pair_lines(const char *path, int64_t n, const int64_t *a, const int64_t *b, const int64_t *ival,
                      const double *fval, const char *sep, const char *tail);
void pool_expected(int64_t size, int64_t ndiag, const int64_t *sums, const int64_t *cells, double min_n,
                   double *out_val, int64_t *nkeys);
void pool_decay(int64_t size, const double *nrm, const int64_t *raw, const int64_t *ndiag, double min_n,
                double *out_val, uint8_t *out_state, int32_t *empty);
void region_cells(tb_store *s, int64_t r0, int64_t r1, int64_t c0, int64_t c1, const uint8_t *h_bad,
                  const double *h_bias, const double *h_decay, int half, int64_t *n, int32_t *h_row,
                  int32_t *h_col, int32_t *h_raw, double *h_nrm, double *h_dec);
// compartments.cu
void corr_build(tb_store *s, int64_t beg, int64_t end, const double *h_bias, const uint8_t *h_bad,
                const double *h_expected, int stage, int64_t *n_good);
void corr_get(tb_store *s, double *h_out);
void corr_set(tb_store *s, int64_t n, const double *h_in);
void corr_eig(tb_store *s, int32_t k, double *h_evals, double *h_evecs, int32_t info[8]);
// comm.cu
void comm_unique_id(uint8_t id[128]);
void comm_init(tb_store *s, int rank, int world, const uint8_t id[128]);
tb_comm *group_create(int rank, int world, const uint8_t id[128], int device);
void group_destroy(tb_comm *g);
void store_attach(tb_store *s, tb_comm *g, bool owned);
void comm_destroy(tb_store *s);
void exchange_setup(tb_store *s);          // (re)allocate the row-sum exchange buffers; maps the peers' when world > 1
void exchange_release(tb_store *s);
void allreduce_f64(tb_store *s, const double *send, double *recv, size_t n);
void allreduce_i64(tb_store *s, const long long *send, long long *recv, size_t n);
}  // namespace tb
