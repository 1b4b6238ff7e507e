// K1 -- the ICE row sum s_i = sum_j v_ij * x_j over the compact stream (encode.cu).
// Replaces _update_S + _update_W of the reference (_pytadbit/utils/normalize_hic.py:136-162)
// in the non-mutating form S = x * (W0 x).
//
// Two kernels per pass, no atomics, fixed summation order:
//
//   k_rowsum_dense   A warp owns a group of kRows dense chunks that cover the SAME column
//                    panel of different rows.  The payloads are pulled into shared memory by
//                    the bulk-copy engine (cp.async.bulk -> mbarrier complete_tx, SASS UBLKCP):
//                    every warp keeps a private ring of kStages stages, each kRows x kSub
//                    bytes, and runs its producer cursor kStages sub-blocks ahead of its
//                    consumer cursor, across group boundaries.  HBM latency is therefore
//                    covered by bytes in flight in the copy engine (up to 128 KB per SM),
//                    not by resident warps or registers.  The groups are sorted by panel and
//                    a CTA owns a contiguous range of them, so the 32 KB of x of the current
//                    panel sit in shared memory too (one bulk copy per panel change, 1-3 per
//                    CTA and pass) instead of being re-read from L2 by every warp.  Per 128
//                    columns a lane reads one word per row and two 16-byte x values shared by
//                    the kRows rows (all conflict-free shared-memory loads), then 4 x kRows DFMA.
//   k_rowsum_sparse  one per-cell chunk (negative or very large values) per CTA, x gathered
//                    through L1/L2.
//
// Integer -> double by exponent splice (one DADD) instead of I2F.F64.
#include <type_traits>

#include <stdlib.h>

#include "common.cuh"

namespace tb {

namespace {

constexpr int kStages = 2;          // sub-blocks in flight per warp
constexpr int kWarpRing = kStages * kRows * kSub;                 // 8 KB per warp
constexpr int kPanelBytes = kPanel * 8;                           // x of one column panel, 32 KB
constexpr int kDenseSmem = kPanelBytes + kDenseWarps * kWarpRing + (kDenseWarps * kStages + 1) * 8;

__device__ __forceinline__ double u2d(uint32_t v) {      // exact for v < 2^32
  return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}
__device__ __forceinline__ double i2d(int32_t v) {       // exact for any int32
  return __hiloint2double(0x43300000, (int)((uint32_t)v ^ 0x80000000u)) - 4503601774854144.0;
}

__device__ __forceinline__ uint32_t ldg_stream_u32(const void *p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_stream_u64(const void *p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ double warp_sum(double a) {
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  return a;
}

// ---- mbarrier / bulk copy (PTX) -----------------------------------------------------------
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

__device__ __forceinline__ double2 lds_x2(const double *p) {      // x panel in shared memory
  double2 r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(smem_u32(p)));
  return r;
}

// One 512-byte piece of a stage for every row: 512/256/128 columns for D8/D16/D32.
// Straight-line: per column and row one byte-extract (PRMT), one DADD (exponent splice) and
// one DFMA; x is loaded once for the kRows rows.
template <int ENC>
__device__ __forceinline__ void consume(const uint8_t *stage, const double *xb, int lane,
                                        double lo[kRows], double hi[kRows]) {
  constexpr int W = ENC == ENC_D32 ? 4 : (ENC == ENC_D16 ? 2 : 1);
  constexpr int G = 512 / (128 * W);             // 128-column groups per 512-byte piece
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const double2 xa = lds_x2(xb + g * 128), xc = lds_x2(xb + g * 128 + 64);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const uint8_t *p = stage + r * kSub + g * 128 * W;
      double v0, v1, v2, v3;
      if (ENC == ENC_D8) {
        const uint32_t w = reinterpret_cast<const uint32_t *>(p)[lane];
        v0 = u2d(__byte_perm(w, 0, 0x4440)); v1 = u2d(__byte_perm(w, 0, 0x4441));
        v2 = u2d(__byte_perm(w, 0, 0x4442)); v3 = u2d(__byte_perm(w, 0, 0x4443));
      } else if (ENC == ENC_D16) {
        const uint2 w = reinterpret_cast<const uint2 *>(p)[lane];
        v0 = u2d(w.x & 0xffff); v1 = u2d(w.x >> 16); v2 = u2d(w.y & 0xffff); v3 = u2d(w.y >> 16);
      } else {
        const uint4 w = reinterpret_cast<const uint4 *>(p)[lane];
        v0 = i2d((int32_t)w.x); v1 = i2d((int32_t)w.y); v2 = i2d((int32_t)w.z); v3 = i2d((int32_t)w.w);
      }
      lo[r] = fma(v0, xa.x, lo[r]);
      hi[r] = fma(v1, xa.y, hi[r]);
      lo[r] = fma(v2, xc.x, lo[r]);
      hi[r] = fma(v3, xc.y, hi[r]);
    }
  }
}

template <int ENC>
__device__ __forceinline__ void consume_stage(const uint8_t *sp, const double *xb, int lane,
                                              double lo[kRows], double hi[kRows]) {
  constexpr int W = ENC == ENC_D32 ? 4 : (ENC == ENC_D16 ? 2 : 1);
#pragma unroll 1
  for (int h = 0; h < kSub / 512; ++h) consume<ENC>(sp + h * 512, xb + h * (512 / W), lane, lo, hi);
}

__device__ __forceinline__ GroupRec load_rec(const GroupRec *__restrict__ recs, int64_t gi) {
  const int4 *p = reinterpret_cast<const int4 *>(recs + gi);
  const int4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  GroupRec g;                            // novf[] is read where it is needed (load_novf)
  g.off16[0] = a.x; g.off16[1] = a.y; g.off16[2] = a.z; g.off16[3] = a.w;
  g.id[0] = b.x; g.id[1] = b.y; g.id[2] = b.z; g.id[3] = b.w;
  g.c0 = c.x; g.nsub = c.y; g.enc = c.z; g.seg_end = c.w;
  return g;
}

template <int NT>
__device__ __forceinline__ void sparse_chunk(const uint8_t *__restrict__ data, const int64_t *__restrict__ off,
                                             const int4 *__restrict__ desc, int32_t c,
                                             const double *__restrict__ x, double *__restrict__ partial,
                                             double *wsum);

__global__ void __launch_bounds__(kDenseWarps * 32, 1)
k_rowsum_dense(const uint8_t *__restrict__ data, const GroupRec *__restrict__ recs,
               const int32_t *__restrict__ cta_range, const int4 *__restrict__ desc,
               const double *__restrict__ x, double *__restrict__ partial,
               const IceState *__restrict__ state, const int64_t *__restrict__ enc_off,
               const int32_t *__restrict__ sparse_ids, int64_t n_sparse_tail) {
  if (state && state->stopped) return;
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const double *xs = reinterpret_cast<const double *>(smem);           // x of the current panel
  uint8_t *ring = smem + kPanelBytes + wic * kWarpRing;
  const uint32_t ring_u32 = smem_u32(ring);
  const uint32_t bar0 = smem_u32(smem + kPanelBytes + kDenseWarps * kWarpRing);
  const uint32_t bars = bar0 + wic * kStages * 8;
  const uint32_t xbar = bar0 + kDenseWarps * kStages * 8;
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bars + 8 * s, 1);
    if (wic == 0) mbar_init(xbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // this CTA's contiguous range of groups; its warps take them round robin
  const int64_t g0 = cta_range[blockIdx.x], g1 = cta_range[blockIdx.x + 1];

  // Producer cursor (pg, psub) runs kStages sub-blocks ahead of the consumer cursor (cg, csub)
  // over the same sequence of (group, sub-block) pairs, across group and panel boundaries.
  int64_t pgi = g0 + wic, cgi = g0 + wic;
  int psub = 0, csub = 0;
  // producer keeps only what the copies need: payload offsets (0xffffffff = empty seat)
  uint32_t poff[kRows];
  int pnsub = 0;
  GroupRec cg;
  auto load_producer = [&](const GroupRec &g) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) poff[r] = g.id[r] >= 0 ? g.off16[r] : 0xffffffffu;
    pnsub = g.nsub;
  };
  if (pgi < g1) {
    cg = load_rec(recs, pgi);
    load_producer(cg);
  }
  auto issue = [&](int stage) {             // all lanes advance the cursor, lane 0 issues
    if (pgi >= g1) return;
    if (lane == 0) {
      int nvalid = 0;
#pragma unroll
      for (int r = 0; r < kRows; ++r) nvalid += poff[r] != 0xffffffffu;
      const uint32_t bar = bars + 8 * stage;
      mbar_expect_tx(bar, (uint32_t)(nvalid * kSub));
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        if (poff[r] != 0xffffffffu) {
          TB_DCHECK(((long long)poff[r] << 4) + ((long long)psub + 1) * kSub <= c_lim.enc_bytes);
        }
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        if (poff[r] != 0xffffffffu)
          bulk_g2s(ring_u32 + (stage * kRows + r) * kSub,
                   data + ((int64_t)poff[r] << 4) + (int64_t)psub * kSub, kSub, bar);
    }
    if (++psub == pnsub) {
      psub = 0;
      pgi += kDenseWarps;
      if (pgi < g1) load_producer(load_rec(recs, pgi));
    }
  };
#pragma unroll
  for (int s = 0; s < kStages; ++s) issue(s);

  double lo[kRows], hi[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) lo[r] = hi[r] = 0.0;
  uint32_t rparity = 0, xparity = 0;        // rparity bit q = parity of ring stage q
  int stage = 0;
  // one sub-block of group cgi; the ring stage is a compile-time constant (fixed addresses)
  auto step = [&](auto stage_c) {
    constexpr int ST = decltype(stage_c)::value;
    mbar_wait(bars + 8 * ST, (rparity >> ST) & 1u);
    rparity ^= 1u << ST;
    const uint8_t *sp = ring + ST * kRows * kSub;
    const int enc = cg.enc & 0xff;
    if (enc == ENC_D8)
      consume_stage<ENC_D8>(sp, xs + csub * kSub + 2 * lane, lane, lo, hi);
    else if (enc == ENC_D16)
      consume_stage<ENC_D16>(sp, xs + csub * (kSub / 2) + 2 * lane, lane, lo, hi);
    else
      consume_stage<ENC_D32>(sp, xs + csub * (kSub / 4) + 2 * lane, lane, lo, hi);
    __syncwarp();                          // every lane is done with this stage's bytes
    issue(ST);
    if (++csub == cg.nsub) {               // group finished: overflow cells, reduce, publish
      if (cg.enc >> 8) {                     // cells outside [0, 2^15): a list behind the seat (rare)
        const int4 nv = __ldg(reinterpret_cast<const int4 *>(recs + cgi) + 3);
        const int novf[kRows] = {nv.x, nv.y, nv.z, nv.w};
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          const int2 *cells = reinterpret_cast<const int2 *>(
              data + ((int64_t)cg.off16[r] << 4) + (int64_t)cg.nsub * kSub);
          for (int k = lane; k < novf[r]; k += 32) {
            const int2 e = __ldg(cells + k);
            TB_DCHECK(e.x >= cg.c0 && e.x - cg.c0 < kPanel);
            lo[r] = fma(i2d(e.y), xs[e.x - cg.c0], lo[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const double a = warp_sum(lo[r] + hi[r]);
        TB_DCHECK(cg.id[r] < c_lim.nchunks);
        if (lane == 0 && cg.id[r] >= 0) partial[cg.id[r]] = a;
        lo[r] = hi[r] = 0.0;
      }
      csub = 0;
      cgi += kDenseWarps;
      if (cgi < g1) cg = load_rec(recs, cgi);
    }
  };
  for (int64_t g = g0; g < g1;) {            // panel segments of the range
    const int4 head = __ldg(reinterpret_cast<const int4 *>(recs + g) + 2);   // c0, nsub, enc, seg_end
    const int64_t ge = head.w < g1 ? head.w : g1;
    TB_DCHECK(g1 <= c_lim.n_groups && head.x >= 0 && (long long)head.x + kPanel <= c_lim.nbins + kXPad);
    __syncthreads();                         // every warp is done with the previous panel
    if (threadIdx.x == 0) {
      mbar_expect_tx(xbar, kPanelBytes);
      bulk_g2s(smem_u32(smem), x + head.x, kPanelBytes, xbar);
    }
    mbar_wait(xbar, xparity);
    xparity ^= 1u;
    while (cgi < ge) {
      if (stage == 0) {
        step(std::integral_constant<int, 0>());
        stage = 1;
        if (cgi >= ge) break;
      }
      step(std::integral_constant<int, 1>());
      stage = 0;
    }
    g = ge;
  }
  // the few per-cell chunks of the store (values that fit no other encoding), one per CTA that is
  // done with its panels: no launch of their own (10 us per pass for a single chunk otherwise)
  if (n_sparse_tail) {
    __syncthreads();
    double *wsum = reinterpret_cast<double *>(smem);          // the x panel is no longer needed
    for (int64_t i = blockIdx.x; i < n_sparse_tail; i += gridDim.x)
      sparse_chunk<kDenseWarps * 32>(data, enc_off, desc, sparse_ids[i], x, partial, wsum);
  }
}

// One per-cell chunk (negative or very large values: there are few) by the NT threads of a CTA:
// the gathers go through L1/L2, warp sums are added in warp order.
template <int NT>
__device__ __forceinline__ void sparse_chunk(const uint8_t *__restrict__ data, const int64_t *__restrict__ off,
                                             const int4 *__restrict__ desc, int32_t c,
                                             const double *__restrict__ x, double *__restrict__ partial,
                                             double *wsum) {
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5, tid = threadIdx.x;
  TB_DCHECK(c >= 0 && c < c_lim.nchunks);
  const int4 d = desc[c];
  const uint8_t *p = data + off[c];
  const int n = d.y;
  TB_DCHECK(off[c] + (long long)n * 4 <= c_lim.enc_bytes);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if ((d.z & 0xff) == ENC_S16) {
    const double *xb = x + d.x;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p);
    int k = tid;
    for (; k + 3 * NT < n; k += 4 * NT) {
      const uint32_t e0 = ldg_stream_u32(w + k), e1 = ldg_stream_u32(w + k + NT);
      const uint32_t e2 = ldg_stream_u32(w + k + 2 * NT), e3 = ldg_stream_u32(w + k + 3 * NT);
      const double x0 = __ldg(xb + (e0 & 0xffff)), x1 = __ldg(xb + (e1 & 0xffff));
      const double x2 = __ldg(xb + (e2 & 0xffff)), x3 = __ldg(xb + (e3 & 0xffff));
      a0 = fma(u2d(e0 >> 16), x0, a0);
      a1 = fma(u2d(e1 >> 16), x1, a1);
      a2 = fma(u2d(e2 >> 16), x2, a2);
      a3 = fma(u2d(e3 >> 16), x3, a3);
    }
    for (; k < n; k += NT) {
      const uint32_t e0 = ldg_stream_u32(w + k);
      a0 = fma(u2d(e0 >> 16), __ldg(xb + (e0 & 0xffff)), a0);
    }
  } else {                                     // ENC_S32
    const uint2 *w = reinterpret_cast<const uint2 *>(p);
    int k = tid;
    for (; k + 3 * NT < n; k += 4 * NT) {
      const uint2 e0 = ldg_stream_u64(w + k), e1 = ldg_stream_u64(w + k + NT);
      const uint2 e2 = ldg_stream_u64(w + k + 2 * NT), e3 = ldg_stream_u64(w + k + 3 * NT);
      const double x0 = __ldg(x + e0.x), x1 = __ldg(x + e1.x);
      const double x2 = __ldg(x + e2.x), x3 = __ldg(x + e3.x);
      a0 = fma(i2d((int32_t)e0.y), x0, a0);
      a1 = fma(i2d((int32_t)e1.y), x1, a1);
      a2 = fma(i2d((int32_t)e2.y), x2, a2);
      a3 = fma(i2d((int32_t)e3.y), x3, a3);
    }
    for (; k < n; k += NT) {
      const uint2 e0 = ldg_stream_u64(w + k);
      a0 = fma(i2d((int32_t)e0.y), __ldg(x + e0.x), a0);
    }
  }
  const double a = warp_sum((a0 + a1) + (a2 + a3));
  if (lane == 0) wsum[wic] = a;
  __syncthreads();
  if (tid == 0) {
    double t = wsum[0];
#pragma unroll
    for (int q = 1; q < NT / 32; ++q) t += wsum[q];
    partial[c] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
k_rowsum_sparse(const uint8_t *__restrict__ data, const int64_t *__restrict__ off,
                const int4 *__restrict__ desc, const int32_t *__restrict__ ids, int64_t ns,
                const double *__restrict__ x, double *__restrict__ partial,
                const IceState *__restrict__ state) {
  if (state && state->stopped) return;
  __shared__ double wsum[8];
  for (int64_t i = blockIdx.x; i < ns; i += gridDim.x)
    sparse_chunk<256>(data, off, desc, ids[i], x, partial, wsum);
}

}  // namespace

// per-cell chunks ride at the end of the dense kernel when there are only a few of them
bool sparse_in_dense_tail(const tb_store *s) {
  return s->n_groups > 0 && s->n_sparse <= 2 * (int64_t)s->n_dense_ctas;
}

// One K1 evaluation = the dense kernel (static panel ranges per CTA) and the tile kernel (work
// units fetched from a counter), both one CTA per SM with all of its shared memory.  Back to back
// on one stream each pays its own tail: the second kernel cannot start before the slowest CTA of
// the first has left.  They are independent (own partials, same x), so the dense kernel goes to
// a second stream first and the tile kernel follows on the main stream: its CTAs start on an SM
// as soon as the dense CTA there retires, and the dynamic unit fetch absorbs the spread
// (option k1_streams = 1 | 2, environment TB_K1_STREAMS; default 2).  Same bits either way.
// Measured on the genome-wide 5 kb matrix (tb_probe_k1): whole matrix 2.067 -> 2.055 ms per pass
// (0.5 %), a 1/2 shard 1.044 -> 1.029 ms (1.4 %), a 1/8 shard 0.309 -> 0.296 ms (4.4 %): the
// shorter the kernels, the larger the share of their tails.
static bool two_streams(const tb_store *s) {
  if (s->k1_streams_mode >= 0) return s->k1_streams_mode == 2;
  static int env = -1;
  if (env < 0) {
    const char *e = getenv("TB_K1_STREAMS");
    env = (e && e[0] == '1') ? 1 : 2;
  }
  return env == 2;
}

void launch_rowsum(tb_store *s, const double *x, const IceState *state) {
  if (!two_streams(s) || !s->n_groups || !s->t_nunits) {
    launch_rowsum_tiles(s, x, state);
    launch_rowsum_chunks(s, x, state);
    return;
  }
  if (!s->stream2) {
    // highest priority: of two kernels that become runnable together its CTAs are placed first
    int prio_low = 0, prio_high = 0;
    TB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
    TB_CUDA(cudaStreamCreateWithPriority(&s->stream2, cudaStreamNonBlocking, prio_high));
    TB_CUDA(cudaEventCreateWithFlags(&s->k1_fork, cudaEventDisableTiming));
    TB_CUDA(cudaEventCreateWithFlags(&s->k1_join, cudaEventDisableTiming));
  }
  cudaStream_t main_stream = s->stream;
  TB_CUDA(cudaEventRecord(s->k1_fork, main_stream));
  TB_CUDA(cudaStreamWaitEvent(s->stream2, s->k1_fork, 0));
  s->stream = s->stream2;                     // the launch helpers use the store's stream
  try {
    launch_rowsum_chunks(s, x, state);
  } catch (...) {
    s->stream = main_stream;
    throw;
  }
  s->stream = main_stream;
  TB_CUDA(cudaEventRecord(s->k1_join, s->stream2));
  launch_rowsum_tiles(s, x, state);
  TB_CUDA(cudaStreamWaitEvent(main_stream, s->k1_join, 0));
}

void launch_rowsum_chunks(tb_store *s, const double *x, const IceState *state) {
  set_check_limits(s);
  if (s->n_groups) {
    static bool attr_set[64] = {false};          // per device: the attribute is per context
    if (!attr_set[s->device & 63]) {
      TB_CUDA(cudaFuncSetAttribute(k_rowsum_dense, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kDenseSmem));
      attr_set[s->device & 63] = true;
    }
    k_rowsum_dense<<<s->n_dense_ctas, kDenseWarps * 32, kDenseSmem, s->stream>>>(
        s->enc_data.p, s->groups.p, s->group_range.p, s->enc_desc.p, x, s->partial_f.p, state, s->enc_off.p,
        s->work_ids.p, sparse_in_dense_tail(s) ? s->n_sparse : 0);
  }
  if (s->n_sparse && !sparse_in_dense_tail(s)) {
    const int64_t cap = (int64_t)s->sm_count * 8;
    const int grid = (int)(s->n_sparse < cap ? s->n_sparse : cap);
    k_rowsum_sparse<<<grid, 256, 0, s->stream>>>(s->enc_data.p, s->enc_off.p, s->enc_desc.p,
                                                 s->work_ids.p, s->n_sparse, x, s->partial_f.p, state);
  }
}

}  // namespace tb
