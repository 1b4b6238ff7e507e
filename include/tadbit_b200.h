/*
 * tadbit_b200 -- C-ABI of the B200-native Hi-C matrix-balancing hot path.
 *
 * Drop-in boundary for ONE path of 3DGenomes/TADbit: filter_columns -> ICE
 * (normalize_hic) -> per-diagonal expected.  The reference has no FFI for this path
 * (it is pure Python over a dict); each entry point below names the reference
 * function it replaces (paths relative to the reference root).  The Python host class
 * tadbit_b200.HiC_data binds these with ctypes (tadbit_b200/_capi.py) and keeps the
 * reference's method signatures; INTEGRATION.md shows the stub a TADbit maintainer
 * would add.
 *
 * Conventions
 *   - every function returns a tb_status (0 = ok); tb_last_error() gives the text of
 *     the last failure on the calling thread;
 *   - plain pointers and sizes only.  Pointers named h_* are HOST memory, d_* are
 *     DEVICE memory on the store's device, "mem" arguments say which (TB_MEM_*);
 *   - the library never frees caller memory; a tb_store owns every device buffer it
 *     allocates and releases them in tb_store_destroy;
 *   - a tb_store is not thread-safe; calls block until their result is in the output
 *     buffers (the store's private CUDA stream is synchronised before returning);
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     TB_ERR_CUDA.
 *
 * Contact store.  A symmetric N x N integer matrix given as its stored upper-triangle
 * cells (row <= col, each cell once; stored zeros allowed and counted as stored, as the
 * reference's dict does -- _pytadbit/hic_data.py:54-118).  A store may own only a
 * contiguous block of matrix rows [row_begin, row_end) (multi-GPU row sharding); it then
 * keeps every cell of the FULL symmetric matrix whose row lies in the block.
 */
#ifndef TADBIT_B200_H
#define TADBIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tb_store tb_store;

typedef enum {
  TB_OK = 0,
  TB_ERR_ARG = 1,      /* bad argument (null pointer, row > col, out of range ...)        */
  TB_ERR_CUDA = 2,     /* CUDA runtime failure or no device                               */
  TB_ERR_ALL_BAD = 3,  /* ICE: no row left after masking (reference: ZeroDivisionError
                          'ERROR: normalization failed, all bad columns',
                          _pytadbit/utils/normalize_hic.py:207-208)                      */
  TB_ERR_NCCL = 4,     /* NCCL failure / NCCL not loadable                                */
  TB_ERR_STATE = 5,    /* call not valid in this state (e.g. comm not initialised), or a
                          GPU of the group never reached the row-sum exchange             */
  TB_ERR_ZERO_DIV = 6  /* the reference would divide by zero: mean row sum 0 in ICE
                          (_pytadbit/utils/normalize_hic.py:148), a bias of 0 in the
                          normalised sum (_pytadbit/hic_data.py:371-374)                 */
} tb_status;

enum { TB_MEM_HOST = 0, TB_MEM_DEVICE = 1 };
/* which cells a CSR-form argument holds (tb_store_create_csr, tb_store_export_csr) */
enum { TB_CELLS_UPPER = 0,      /* stored cells with col >= row of the rows concerned             */
       TB_CELLS_FULL_ROWS = 1   /* every stored cell of the rows concerned (both triangles)       */ };

/* ---- library ------------------------------------------------------------------- */
int tb_version(void);
const char *tb_last_error(void);
int tb_device_count(int *count);
/* Device memory is cached in the device's stream-ordered pool between stores (building a
 * store needs tens of GB of scratch); this returns the cached memory to the driver. */
int tb_trim_memory(int device);

/* ---- store construction -------------------------------------------------------- */
/*
 * Replaces building the dict (HiC_data.__init__, _pytadbit/hic_data.py:58-81) and the
 * array export HiC_data.get_hic_data_as_csr (_pytadbit/hic_data.py:166-181).
 *
 * nbins        N, the matrix size (len(hic_data), hic_data.py:124)
 * nnz          number of upper-triangle cells passed
 * row,col,val  int32[nnz], row[k] <= col[k] < N, cells unique, any order
 * mem          TB_MEM_HOST or TB_MEM_DEVICE for row/col/val
 * n_chrom      number of sections (0 = one section [0,N): hic_data.py:76-78)
 * chrom_offsets int64[n_chrom+1] HOST, bin offsets of the sections (section_pos,
 *              hic_data.py:70-75); ignored when n_chrom == 0
 * row_begin,row_end  block of full-matrix rows this store owns (0,N for one GPU)
 * device       CUDA device ordinal
 */
int tb_store_create_coo(int64_t nbins, int64_t nnz, const int32_t *row, const int32_t *col,
                        const int32_t *val, int mem, int32_t n_chrom,
                        const int64_t *chrom_offsets, int64_t row_begin, int64_t row_end,
                        int device, tb_store **out);

/*
 * Synthetic Hi-C matrix generated on the device (SURVEY.md 8d model): per-bin
 * visibility b_i, cis mean a*b_i*b_j*(|i-j|+1)^-alpha, trans mean t*b_i*b_j, counts
 * Poisson(mean) decided per CELL by a counter-based hash of (seed, min(i,j), max(i,j)),
 * so the matrix does not depend on how rows are sharded.  Not in the reference (its
 * test/simulate_genomic_matrix.py is unrelated); used by bench.py and the scale tests.
 */
/* The same matrix in CSR form (what HiC_data.get_hic_data_as_csr holds, _pytadbit/hic_data.py:166-181).
 * form = TB_CELLS_UPPER: upper-triangle rows of the WHOLE matrix, ptr[nbins + 1]; 8 bytes per cell
 *   cross PCIe instead of the 12 of the coordinate form.  row_begin/row_end must be 0/nbins.
 * form = TB_CELLS_FULL_ROWS: every stored cell of the rows [row_begin, row_end) of the full
 *   symmetric matrix, ptr[row_end - row_begin + 1] (starting at 0), columns strictly ascending in
 *   each row: the pre-partitioned input of a row-sharded matrix (each GPU receives only its rows,
 *   nothing has to be mirrored or sorted).
 * ptr is int64; col/val as in tb_store_create_coo; mem says where all three arrays live. */
int tb_store_create_csr(int64_t nbins, int form, const int64_t *ptr, const int32_t *col,
                        const int32_t *val, int mem, int32_t n_chrom, const int64_t *chrom_offsets,
                        int64_t row_begin, int64_t row_end, int device, tb_store **out);

typedef struct {
  uint64_t seed;
  double cis_scale;      /* a      */
  double cis_alpha;      /* alpha  */
  double trans_mean;     /* t      */
  double vis_sigma;      /* sigma of log-normal visibility          */
  double low_frac;       /* fraction of bins scaled by low_scale    */
  double low_scale;
  double empty_frac;     /* fraction of bins with zero visibility   */
} tb_synth_params;

int tb_store_create_synthetic(int64_t nbins, int32_t n_chrom, const int64_t *chrom_offsets,
                              const tb_synth_params *params, int64_t row_begin,
                              int64_t row_end, int device, tb_store **out);
/* full-row stored-cell counts of rows [row_begin,row_end) of the synthetic matrix
 * (int64[row_end-row_begin], HOST) without materialising it -- used to balance shards */
int tb_synthetic_row_counts(int64_t nbins, int32_t n_chrom, const int64_t *chrom_offsets,
                            const tb_synth_params *params, int64_t row_begin,
                            int64_t row_end, int device, int64_t *h_counts);

int tb_store_destroy(tb_store *s);

/* nbins, owned row block, stored upper-triangle cells with row in the block (i<=j),
 * stored full-matrix cells in the block, device bytes held */
int tb_store_info(const tb_store *s, int64_t *nbins, int64_t *row_begin, int64_t *row_end,
                  int64_t *nnz_upper_owned, int64_t *nnz_full_owned, int64_t *device_bytes);

/* Options of a store.
 *   "rowsum" = "compact" (default) or "csr": which copy of the matrix the ICE passes walk -- the
 *     compact stream, or the plain CSR (two independent K1 implementations; the parity tests
 *     compare them).  A store of positive counts releases its CSR after the build unless the
 *     environment has TB_KEEP_CSR=1 at creation; "csr" then fails with TB_ERR_STATE.
 *   "exchange" = "fused" (default) or "nccl": per ICE pass either ONE kernel that stores the owned
 *     row sums into every GPU's buffer (peer-mapped memory), waits on flags and does the
 *     statistics and the bias update, or combine -> ncclAllReduce -> statistics -> update as
 *     separate launches.  All ranks of a group must choose the same.
 *   "k1_streams" = "2" (default) or "1": the dense kernel of a row-sum evaluation on a second
 *     stream next to the tile kernel, or the two kernels back to back on the store's stream
 *     (same bits either way; 0.5 % faster on a whole genome-wide matrix, 4.4 % on a 1/8 shard).
 *   "corr_gemm" = "dmma" (default) or "simple": X X^T of tb_corr_build as fp64 tensor-core tiles
 *     or as a plain kernel (parity reference of the former). */
int tb_store_set_option(tb_store *s, const char *name, const char *value);

/* Layout of the compact row-sum stream (tadbit_b200/csrc/encode.cu, tiles.cu):
 * out[0] chunk payload bytes, out[1..6] chunks encoded as S32, S16, D32, D16, D8, T (cells in
 * the sparse tiles), out[7] chunks, out[8] CSR bytes, out[9] cells in sparse tiles, out[10]
 * the same with padding, out[11] tile work units, out[12] cells kept in overflow lists behind
 * dense chunks (value fits neither the slot nor a tile cell), out[13] bytes of the sparse tiles,
 * out[14] cells of D8 chunks stored in the sparse tiles (255 < count < 2^15). */
int tb_store_layout(const tb_store *s, int64_t out[16]);

/* Export the owned upper-triangle cells (row in block, row <= col), row-major sorted.
 * Call with row == NULL to get the count in *nnz first.  Buffers are HOST. */
int tb_store_export_upper(const tb_store *s, int64_t *nnz, int32_t *h_row, int32_t *h_col,
                          int32_t *h_val);
/* The owned rows in CSR form (see tb_store_create_csr): h_ptr[row_end - row_begin + 1], h_col and
 * h_val [*nnz].  Call with null arrays first to learn *nnz. */
int tb_store_export_csr(const tb_store *s, int form, int64_t *nnz, int64_t *h_ptr, int32_t *h_col,
                        int32_t *h_val);

/* ---- multi-GPU (one process per GPU) ------------------------------------------- */
/* 128-byte NCCL unique id made on rank 0 and passed to every rank by the host
 * (torch.distributed / MPI / files -- plumbing is the caller's). */
int tb_comm_unique_id(uint8_t id_out[128]);
int tb_comm_init(tb_store *s, int32_t rank, int32_t world, const uint8_t id[128]);
/* A group that outlives the stores: creating the NCCL communicator costs most of a second, so a
 * long-running process creates it once per GPU and attaches every matrix it normalises.
 * tb_store_attach replaces tb_comm_init for such a store (the row-sum exchange buffers of the
 * store are set up and mapped by its peers, a collective call: every rank attaches its shard).
 * The group must be destroyed after the stores attached to it. */
typedef struct tb_comm tb_comm;
int tb_comm_create(int32_t rank, int32_t world, const uint8_t id[128], int device, tb_comm **out);
int tb_comm_destroy(tb_comm *c);
int tb_store_attach(tb_store *s, tb_comm *c);

/* ---- K4: filter statistics ----------------------------------------------------- */
/*
 * Replaces the two loops of filter_by_zero_count
 * (_pytadbit/utils/hic_filtering.py:185-190 stored cells per row, :198-200 row sums).
 * Outputs int64[N] HOST, exact; rows outside the owned block are all-reduced in when a
 * communicator is set, otherwise left 0.
 */
int tb_row_stats(tb_store *s, int64_t *h_nnz_row, int64_t *h_row_sum);
/*
 * Replaces the O(N^2) column-sum list comprehensions of filter_by_mean
 * (_pytadbit/utils/hic_filtering.py:36-39 with the bad rows excluded; :141-145 with
 * h_bad == NULL, all rows).  h_bad: uint8[N] HOST (non-zero = bad) or NULL.
 * h_col_sum: int64[N] HOST, sum over NON-bad rows of every column, exact.
 */
int tb_col_sums_masked(tb_store *s, const uint8_t *h_bad, int64_t *h_col_sum);

/* ---- K1 + K2: ICE --------------------------------------------------------------- */
/*
 * Replaces iterative() incl. copy_matrix/_update_S/_updateDB/_update_W
 * (_pytadbit/utils/normalize_hic.py:136-231).
 *
 * h_bad       uint8[N] HOST or NULL (bads dict as a mask)
 * iterations  as the reference: at most iterations+1 passes; 0 = one pass, no test
 * max_dev     stop when max(|Smin/meanS-1|,|Smax/meanS-1|) < max_dev
 * h_bias      float64[N] HOST out: B_i*sqrt(meanS); 1.0 for rows not present or B_i==0
 * h_trace     float64[(iterations+1)*4] HOST out or NULL: per pass Smin, meanS, Smax, dev
 *             (the verbose line of normalize_hic.py:219-220)
 * n_passes    passes executed (= trace lines; for iterations==0 it is 1 with no line)
 * Returns TB_ERR_ALL_BAD when no row is present.
 */
int tb_ice(tb_store *s, const uint8_t *h_bad, int32_t iterations, double max_dev,
           double *h_bias, double *h_trace, int32_t *n_passes);

/*
 * Batch mode (BASELINE configs[4]: many experiments normalised at once on one GPU).
 * tb_store_create_batch concatenates n whole stores of one device into a block-diagonal store
 * whose section k is member k (device-to-device; the members stay valid and independent).
 * tb_ice_batch runs iterative() on every SECTION independently -- own present rows, meanS,
 * deviation test and stop pass; a converged section is frozen while the others continue.
 * In the reference this is a Python loop over experiments (Experiment.normalize_hic,
 * _pytadbit/experiment.py:704-748 -> _pytadbit/utils/normalize_hic.py:180-231).
 *   h_bias   float64[N] of the batch store (sections concatenated)
 *   h_trace  float64[n_sections][iterations+1][4] or NULL
 *   n_passes int32[n_sections]; -1 where no row is present (the reference raises
 *            ZeroDivisionError for that experiment; its biases are left at 1.0)
 */
int tb_store_create_batch(int32_t n, tb_store *const *members, int device, tb_store **out);
int tb_ice_batch(tb_store *s, const uint8_t *h_bad, int32_t iterations, double max_dev,
                 double *h_bias, double *h_trace, int32_t *n_passes);

/* Mean device time (ms) of one row-sum evaluation (K1) of this store alone, measured with CUDA
 * events over `reps` launches: ms[0] total, ms[1] sparse-tile kernel, ms[2] dense (+ per-cell)
 * kernels.  Works on a shard that is not connected yet.  tb_row_work gives, for every owned row,
 * the payload bytes of its dense chunks and its cells in sparse tiles (int64[row_end-row_begin],
 * HOST).  The host combines both to re-cut row blocks so that every GPU takes the same time per
 * pass (tadbit_b200/workloads.py:build_sharded). */
int tb_probe_rowsum(tb_store *s, int32_t reps, double ms[3]);
/* Mean device ms of one K1 evaluation issued as the ICE loop issues it: ms[0] with its two kernels
 * back to back on one stream, ms[1] with the dense kernel on a second stream next to the tile
 * kernel (tb_store_set_option "k1_streams" = "1" | "2").  Timing only, also on an unconnected shard. */
int tb_probe_k1(tb_store *s, int32_t reps, double ms[2]);
int tb_row_work(tb_store *s, int64_t *h_dense_bytes, int64_t *h_tile_cells);

/* timing of the last tb_ice call, measured with CUDA events on the store's stream:
 * total device ms of the pass loop, and the ms spent in the row-sum (K1) launches */
int tb_ice_last_timing(const tb_store *s, double *loop_ms, double *rowsum_ms,
                       int32_t *rowsum_launches, int32_t *total_launches);
/* Device times of the last calls on this store, in ms: out[0] ICE loop, [1] its K1 row-sum
 * launches, [2] everything between them (exchange + statistics + update), [3] the last one-pass
 * reduction (tb_col_sums_masked / tb_diag_sums), [4] row-sum launches, [5] kernels launched,
 * [6] row-sum exchange in use: 2 = peer-mapped NVLink stores, 1 = local / NCCL all-reduce,
 * [7] 1 if the store kept its CSR (matrices with zeros or negative values). */
int tb_last_timing(const tb_store *s, double out[8]);

/* ---- K3: normalised sum ---------------------------------------------------------- */
/*
 * Replaces HiC_data.sum(bias, bads) (_pytadbit/hic_data.py:350-375): sum over BOTH
 * triangles (diagonal once) of v/(bias_i*bias_j) for i,j not bad.  h_bias == NULL gives
 * the raw sum.  Result is all-reduced when a communicator is set.
 */
int tb_norm_sum(tb_store *s, const double *h_bias, const uint8_t *h_bad, double *out);

/* ---- K5: per-diagonal sums -------------------------------------------------------- */
/*
 * Replaces the cell loop of _meandiag (_pytadbit/utils/normalize_hic.py:271-283):
 * h_sum_d[d] = sum of cells (i, i+d), i and i+d in the same section, i not bad, for
 * d in [0, ndiag).  int64[ndiag] HOST, exact.  (Valid-cell counts are closed-form from
 * the section lengths and the mask; the host computes them.)
 */
int tb_diag_sums(tb_store *s, const uint8_t *h_bad, int64_t ndiag, int64_t *h_sum_d);

/* ---- consumers of the biases (N1, N2) ---------------------------------------- */
/* N1 -- Experiment.normalize_hic / get_hic_zscores / tadmaths.zscore
 * (_pytadbit/experiment.py:737-743, :751-799; _pytadbit/utils/tadmaths.py:94-172).
 * Over the stored pairs i < j between bins that are not bad, value = count (h_bias NULL) or
 * count / bias[i] / bias[j]: out[0] = number of pairs with a non-zero value, out[1] = the smallest
 * such value, out[2] = sum of log10(value), out[3] = sum of (log10(value) - center)^2.  Pairs with
 * no stored contact are not visited (the caller accounts for them in closed form). */
int tb_pair_stats(tb_store *s, const double *h_bias, const uint8_t *h_bad, double center,
                  double out[4]);
/* The same pairs, row-major: h_val = value, or (log10(value) - mean) / std when zscored.  Call with
 * NULL arrays to learn *n.  A row-sharded store returns the pairs of its own rows. */
int tb_pair_values(tb_store *s, const double *h_bias, const uint8_t *h_bad, int32_t zscored,
                   double mean, double std, int64_t *n, int32_t *h_row, int32_t *h_col,
                   double *h_val);
/* N2 -- the reductions `tadbit normalize` makes once it has the biases
 * (_pytadbit/tools/tadbit_normalize.py:963-1107; sum_dec_matrix :1110-1136, get_cis_perc :1139-1152).
 * h_bias[nbins] (NaN = bad column, as that tool marks them), h_bad optional.
 * h_nrm / h_raw [nbins]: entry (first bin of section c) + d = sum over the cells (i, i + d) inside
 * section c, both bins good, of count / bias[i] / bias[j] and of count (diagonal included).
 * out_cis_total[0..1] = normalised interactions inside sections / in total, off-diagonal pairs. */
int tb_decay_sums(tb_store *s, const double *h_bias, const uint8_t *h_bad, double *h_nrm,
                  int64_t *h_raw, double out_cis_total[2]);
/* N3 (the HiC_data part) -- HiC_data.get_matrix / yield_matrix / write_matrix
 * (_pytadbit/hic_data.py:578-686, :1514-1582): the dense window rows [row_begin, row_end) x columns
 * [col_begin, col_end) of the full symmetric matrix, row-major into h_out: the count, or
 * count / bias[i] / bias[j] when h_bias is given; cells without contacts are 0.  A row-sharded
 * store fills only the rows it owns. */
int tb_window(tb_store *s, int64_t row_begin, int64_t row_end, int64_t col_begin, int64_t col_end,
              const double *h_bias, double *h_out);

/* N3 (the extraction side) -- get_matrix / write_matrix of the BAM parser
 * (_pytadbit/parsers/hic_bam_parser.py:755-872, :892-1242) with the store in the place of the BAM
 * file: the stored cells of the full symmetric matrix inside rows [row_begin, row_end) x columns
 * [col_begin, col_end), row-major; cells of a bin flagged in h_bad (NULL = none) are dropped, as
 * `if i not in bads1 and j not in bads2` does (:860, :1205).  half != 0 keeps the cells whose offset
 * inside the row range is <= their offset inside the column range (_read_half_bam_frag, :447-458,
 * which write_matrix reads with).  Per cell: h_row / h_col global bin indices, h_raw the count,
 * h_nrm = count / bias[i] / bias[j] (transform_value_norm, :828; h_bias[nbins] or NULL), h_dec =
 * h_nrm / h_decay[first bin of the section + |i - j|] for a cell inside one section, NaN between
 * sections (transform_value_decay, :830-833; h_decay[nbins] in the layout tb_decay_sums uses, or
 * NULL).  Call with NULL arrays to learn *n.  A row-sharded store returns the cells of its rows. */
int tb_region_cells(tb_store *s, int64_t row_begin, int64_t row_end, int64_t col_begin, int64_t col_end,
                    const uint8_t *h_bad, const double *h_bias, const double *h_decay, int32_t half,
                    int64_t *n, int32_t *h_row, int32_t *h_col, int32_t *h_raw, double *h_nrm,
                    double *h_dec);

/* Host-only helper of the same path (no GPU involved): appends n cell lines of an .abc file,
 * "a<TAB>b<TAB><TAB>value<NL>" -- what the write_raw / write_bias / write_expc closures of
 * write_matrix print per cell (_pytadbit/parsers/hic_bam_parser.py:1101-1125,
 * '{}\t{}\n'.format('{}\t{}\t'.format(a, b), value)).  value = h_int[k] when h_int is given, else
 * h_val[k] written as Python's repr(float) writes it (shortest digits that read back as the same
 * double; fixed notation for 1e-4 <= |x| < 1e16, d.ddde+XX outside; nan, inf, -inf). */
int tb_write_abc_lines(const char *path, int64_t n, const int32_t *h_a, const int32_t *h_b,
                       const int32_t *h_int, const double *h_val);

/* Host-only, the same for N1: appends n lines "a<sep>b<sep>value<tail>" of
 * Experiment.write_interaction_pairs (_pytadbit/experiment.py:1171-1186: '%s\t%s\t%s\n' for tsv,
 * '%s,%s,%s\n' and '%s,%s,%s,\n' for json); value = h_int[k] when h_int is given, else h_val[k] as
 * Python prints a float.  sep and tail: at most 8 bytes each. */
int tb_write_pair_lines(const char *path, int64_t n, const int64_t *h_a, const int64_t *h_b,
                        const int64_t *h_int, const double *h_val, const char *sep, const char *tail);

/* Host-only tails of K5 and N2 (no GPU involved): the sequential pooling walks over the per-distance
 * sums.  tb_pool_expected -- the loop of `expected` around `_meandiag`
 * (_pytadbit/utils/normalize_hic.py:262-291): h_sum_d / h_cells_d [ndiag] = reads and visited cells
 * per distance (tb_diag_sums and the closed-form counts), consecutive distances pooled until they
 * hold more than min_n reads; h_out[size + 2] receives the value of key d at h_out[d], *nkeys the
 * number of keys (0 .. *nkeys - 1, the trailing key of the reference included).
 * tb_pool_decay -- the pooling of one chromosome in `tadbit normalize`
 * (_pytadbit/tools/tadbit_normalize.py:1077-1106): h_nrm / h_raw [size] from tb_decay_sums, h_ndiag
 * [size] the cells per distance that touch no bad column; h_out[size] the decay per distance,
 * h_state[size] = 0 no key, 1 a key the walk started with, 2 a key the walk added (the reference's
 * dictionary lists the former first); *empty = 1 when the reference ends with an empty dictionary. */
int tb_pool_expected(int64_t size, int64_t ndiag, const int64_t *h_sum_d, const int64_t *h_cells_d,
                     double min_n, double *h_out, int64_t *nkeys);
int tb_pool_decay(int64_t size, const double *h_nrm, const int64_t *h_raw, const int64_t *h_ndiag,
                  double min_n, double *h_out, uint8_t *h_state, int32_t *empty);

/* ---- N4: compartments --------------------------------------------------------- */
/* The numeric core of HiC_data.find_compartments (_pytadbit/hic_data.py:845-937), per section
 * (chromosome) [bin_begin, bin_end) of the genome, over the bins of the section that are not bad
 * (n of them, in genomic order):
 *   stage 0  the observed / expected matrix the reference builds as nested lists (:851-871):
 *            m[a][b] = float(count(i, j)) / h_expected[j - i] / h_bias[i] / h_bias[j] for i <= j,
 *            mirrored (the same order of divisions: the values equal the reference's bit for bit);
 *   stage 1  numpy.corrcoef of its rows (:873) with the reference's NaN -> 0 (:881): rows centred,
 *            X X^T / (n - 1) as fp64 tensor-core tiles (DMMA), c_ij / sd_i / sd_j, clipped to [-1, 1].
 * h_bias[nbins], h_bad[nbins] or NULL, h_expected[bin_end - bin_begin] (expected count per distance;
 * a zero entry is the caller's to refuse: the reference raises ZeroDivisionError).  *n_good = n.
 * The matrix stays resident in the store (one at a time) for tb_corr_get / tb_corr_eig.  On a
 * row-sharded store every rank ends with the whole matrix (one all-reduce of n x n doubles). */
int tb_corr_build(tb_store *s, int64_t bin_begin, int64_t bin_end, const double *h_bias,
                  const uint8_t *h_bad, const double *h_expected, int32_t stage, int64_t *n_good);
/* The resident matrix, row-major n x n doubles, to the host (the reference's savecorr file, :883-913,
 * and its optional scipy median_filter smoothing, :915-917, stay host work) ... */
int tb_corr_get(tb_store *s, double *h_out);
/* ... and a (smoothed) n x n symmetric matrix from the host back into the store. */
int tb_corr_set(tb_store *s, int64_t n, const double *h_in);
/* Replaces scipy.sparse.linalg.eigsh(matrix, k=max_ev) (:926-937): the k eigenpairs of largest
 * magnitude of the resident symmetric matrix by Lanczos iteration with full re-orthogonalisation
 * (every vector operation a kernel; fixed start vector, so the result is reproducible, which
 * ARPACK's is too), eigenvalues in descending order = the order the reference reads them in;
 * k >= n returns all n eigenpairs by descending value (scipy falls back to a dense solver there).
 * h_evals[k], h_evecs[k * n] (vector v at h_evecs + v * n, unit norm; the sign of an eigenvector
 * is arbitrary here as it is in ARPACK).  info[0] eigenpairs returned, [1] Lanczos vectors built,
 * [2] restarts after an invariant subspace, [3] converged, [4] host round trips (the steps between
 * two convergence tests are queued without one), [5..7] reserved.  TB_ERR_STATE if not converged. */
int tb_corr_eig(tb_store *s, int32_t k, double *h_evals, double *h_evecs, int32_t info[8]);
/* Device times of the last tb_corr_build / tb_corr_eig in ms: out[0] cell visitor, [1] centring,
 * [2] X X^T, [3] scaling, [4] eigen-solver; out[5] n, [6] stage resident, [7] 1 if corr_gemm=simple. */
int tb_corr_timing(const tb_store *s, double out[8]);
/* Host-only helper of tb_corr_eig, exported for tests: eigen-decomposition of the symmetric
 * tridiagonal matrix (diag[n], sub[n - 1]) by implicit QL.  diag is overwritten by the
 * eigenvalues (unordered), vecs[r * n + i] = component r of the eigenvector of diag[i]. */
int tb_tridiag_eig(int32_t n, double *diag, const double *sub, double *vecs);

#ifdef __cplusplus
}
#endif
#endif /* TADBIT_B200_H */
