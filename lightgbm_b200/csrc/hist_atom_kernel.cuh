// hist_atom_kernel.cuh — per-leaf gradient/hessian histogram construction (sm_100a), round-2 design.
//
// Replaces: Dataset::ConstructHistograms -> MultiValDenseBin::ConstructHistogramInner
// (reference src/io/dataset.cpp:1293-1485, src/io/multi_val_dense_bin.hpp:58-102), the constant-hessian variant that
// counts and scales (src/io/dataset.cpp:1430-1437, src/io/dense_bin.hpp:98-141), and the reference CUDA kernel
// CUDAConstructHistogramDenseKernel (src/treelearner/cuda/cuda_histogram_constructor.cu:20-71), which does two
// shared-memory *fp32* atomicAdds per cell — on sm_100a those are ATOMS.CAST.SPIN compare-and-swap loops.
//
// What the hardware does fast is the INTEGER shared-memory atomic: ATOMS.ADD (int32, no return) retires ~1 warp
// instruction per clock when the 32 lanes hit 32 different banks (tools/mb/hist_mb.cu, profiles/r02_hist_microbench_*:
// 3.3 clk per 32 cells with two atomics per cell, against 10.1 clk for the LDS.64 / FADD / STS.64 warp-private tables
// of round 1).  This kernel is built on that:
//
//   * EXACT fixed point.  Per tree every row's gradient (hessian) is rounded ONCE to a 30-bit integer
//     q = rint(g * 2^k), 2^k = the largest power of two with max|g| * 2^k < 2^30 (k_quant_rows).  A histogram cell is
//     the integer sum of those q: associative, so the result does not depend on warp scheduling, on the work split, on
//     the flush order or on the number of GPUs — bitwise reproducible — and parent - child is exact.  The rounding
//     error per row is max|g| * 2^-31, two orders of magnitude below the fp32 partial sums of round 1.
//   * A 30-bit q does not fit an int32 accumulator for more than two rows, so it is accumulated as TWO int32 cells:
//     hi = q >> 16 (signed, |hi| <= 2^14) and lo = q & 0xffff (unsigned 16 bit).  Either cell can take 65 535 rows
//     before it can wrap; the tables are flushed ( (hi << 16) + lo -> RED.ADD.64 into the int64 pool slot) at least that
//     often.  Cost: 2 ATOMS per cell and component.
//   * Constant hessian (unweighted L2: the BASELINE configs) needs no hessian sums at all: like the reference
//     (dataset.cpp:1430-1437) the kernel COUNTS rows per cell and scales at the flush.  Counts are 16-bit fields, two
//     bins per int32 word.  3 ATOMS per cell (g hi, g lo, count); general hessians: 4 (g hi, g lo, h hi, h lo).
//   * The tables are shared by ALL consumer warps of the CTA ([bin][column] int32, bank = column): one table set per
//     column group per SM instead of one per warp, so 8 consumer warps (2 per SMSP) hide each other's latencies and the
//     shared memory left over becomes a 24..48-stage ring (49..61 KB of rows in flight per SM for gathered leaves).
//   * A lane handles FOUR columns of ONE row: one LDS.32 fetches its four bin bytes from the row-major stage (a flat
//     array of 32-byte segments), lanes (q, j) = (lane & 7, lane >> 3) take bytes 4q..4q+3 of segment 4u+j, and walk
//     their four columns in the rotated order (k + j) & 3 so that in every step the 32 lanes of the warp touch 32
//     different columns = 32 different banks.  No transposed staging, no per-cell byte loads: 1 LDS.32 + 1 LDS.64 per
//     128 cells.
//   * Staging is unchanged in spirit: producer warps gather the leaf's rows (64-byte segments = whole DRAM atoms when a
//     CTA owns two column groups) with cp.async into an mbarrier ring; the root of an un-bagged tree uses one 2-D TMA
//     tile + one bulk copy per stage.  Work is dealt as contiguous ranges of (column-group set, 32-row stage) so that
//     every CTA gets the same number of stages.
#pragma once
#include <cuda.h>

#include "comm.cuh"
#include "hist_common.cuh"
#include "types.cuh"

namespace b200 {

constexpr int kARows = 32;                 // rows per stage
constexpr int kAConsumers = 8;             // consumer warps (2 per SMSP)
constexpr int kAProducers = 4;             // producer warps (1 per SMSP)
constexpr int kAThreads = (kAConsumers + kAProducers) * 32;
constexpr int kATable = kBinsPerColumn * 32 * 4;      // one [bin][column] int32 table = 32 KB
constexpr int kAFlushRows = 65504;         // rows a table set may take between flushes (multiple of 32, <= 65535)

// CH = constant hessian (count + scale), else general hessians
template <bool CH>
struct AShape {
  static constexpr int G = CH ? 2 : 1;                                   // column groups per CTA
  static constexpr int kCgBytes = CH ? (2 * kATable + kATable / 2) : 4 * kATable;   // 80 KB : 128 KB per column group
  static constexpr int kTables = G * kCgBytes;                           // 160 KB : 128 KB
  static constexpr int kRowBytes = kColGroup * G;                        // bin bytes per staged row
  static constexpr int kStageBytes = kARows * kRowBytes + kARows * 8;    // + int2 (g, h) per row: 2304 : 1280
  static constexpr int kStages = CH ? 24 : 48;                           // multiple of lcm(consumers, producers)
  static constexpr int kSmem = kTables + kStages * kStageBytes + kStages * 2 * 8;
};
static_assert(AShape<true>::kSmem <= 232448 && AShape<false>::kSmem <= 232448, "exceeds 227 KB of shared memory per CTA");
static_assert(AShape<true>::kStages % kAConsumers == 0 && AShape<true>::kStages % kAProducers == 0, "ring/warp mapping");
static_assert(AShape<false>::kStages % kAConsumers == 0 && AShape<false>::kStages % kAProducers == 0, "ring/warp mapping");

struct HistAArgs {
  const uint8_t* bins;            // [num_data x pitch] row-major stored values, pitch a multiple of 64
  int64_t pitch;
  const int2* gq;                 // [num_data] fixed-point (g, h) of this tree by row id (k_quant_rows / k_quantize)
  const int2* gqo0;               // leaf-ordered copies parallel to idx0 / idx1 (written by k_part_scatter), or nullptr
  const int2* gqo1;
  const int32_t* idx0;            // ping-pong row-index buffers
  const int32_t* idx1;
  const Leaf* leaves;
  const Ctl* ctl;
  unsigned long long* pool;       // int64 fixed-point histogram pool [slot][column][256][2]
  int64_t slot_stride;            // int64 elements per slot
  int32_t num_colgroups;          // pitch / 32
  int32_t min_rows_per_cta;       // small leaves: do not spread the (column set x rows) work over more CTAs than this allows
  int32_t use_tma;                // 1: contiguous (root, un-bagged) stages are staged by TMA tile copies
  int32_t l2_prefetch;            // > 0: gathered passes prefetch the rows of the stage this many stages ahead into L2
  // explicit mode (stand-alone ConstructHistogram hook): explicit_n >= 0
  int32_t explicit_n;
  int32_t explicit_slot;
  const int32_t* explicit_idx;    // nullptr = identity
};

__device__ __forceinline__ void red_s32(unsigned addr, int v) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(unsigned addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ int2 lds_i2(unsigned addr) { int2 v; asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr)); return v; }
__device__ __forceinline__ void consumer_bar_sync_a() { asm volatile("bar.sync 1, %0;" ::"n"(kAConsumers * 32) : "memory"); }

// The leaf (or explicit row set) this launch works on, and this CTA's share of it.
struct AWork {
  int n, begin, slot;
  const int32_t* idx;            // nullptr = identity (root of an un-bagged tree)
  const int2* gq_ord;            // leaf-ordered (g, h) (position-indexed), or nullptr => gather by row id
  int spc;                       // stages per column-group set = ceil(n / 32)
  long long v_lo, v_hi;          // this CTA's range of virtual stages (set-major, stage-minor)
};

template <bool CH>
__device__ __forceinline__ bool a_work_setup(const HistAArgs& a, AWork* w) {
  if (a.explicit_n >= 0) {
    w->n = a.explicit_n; w->begin = 0; w->slot = a.explicit_slot; w->idx = a.explicit_idx; w->gq_ord = nullptr;
  } else {
    const Ctl* c = a.ctl;
    if (!c->cur_valid || !c->do_find) return false;
    const Leaf& L = a.leaves[c->smaller];
    w->n = L.lcount; w->begin = L.begin; w->slot = L.slot;
    w->idx = (c->num_leaves == 1 && c->root_identity) ? nullptr : (L.buf ? a.idx1 : a.idx0);
    // every non-root segment was written by its parent's scatter, together with its (g, h) copy
    w->gq_ord = (c->num_leaves > 1 && a.gqo0 != nullptr) ? (L.buf ? a.gqo1 : a.gqo0) + L.begin : nullptr;
  }
  if (w->n <= 0) return false;
  const int sets = a.num_colgroups / AShape<CH>::G;
  w->spc = (w->n + kARows - 1) / kARows;
  const long long total = static_cast<long long>(sets) * w->spc;
  // CTAs that take part: all of them unless the leaf is so small that the per-CTA flush (one RED per touched cell)
  // would dominate; never fewer than one per column-group set
  long long P = (static_cast<long long>(sets) * w->n + a.min_rows_per_cta - 1) / a.min_rows_per_cta;
  if (P < sets) P = sets;
  if (P > static_cast<long long>(gridDim.x)) P = gridDim.x;
  if (static_cast<long long>(blockIdx.x) >= P) return false;
  w->v_lo = total * blockIdx.x / P;
  w->v_hi = total * (blockIdx.x + 1) / P;
  return w->v_hi > w->v_lo;
}

// Flush the CTA's tables of column group `cg` into the leaf's pool slot and zero them.  Called by the 256 consumer
// threads between two consumer barriers.  Thread t: column t & 31, bin pairs (t >> 5) + 8 i.
template <bool CH>
__device__ __forceinline__ void a_flush_zero(unsigned tb, unsigned long long* dst_cg, int t, long long hq_const) {
  const int col = t & 31, pg = t >> 5;
  unsigned long long* dst = dst_cg + static_cast<int64_t>(col) * (kBinsPerColumn * 2);
  const unsigned cb = tb + col * 4;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int m = pg + 8 * i;               // bins 2m, 2m+1
    int hi[2], lo[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const unsigned ca = cb + (2 * m + e) * 128;
      hi[e] = static_cast<int>(lds_u32(ca)); lo[e] = static_cast<int>(lds_u32(ca + kATable));
    }
    long long hv[2];
    if (CH) {
      const unsigned wa = cb + 2 * kATable + m * 128;
      const uint32_t cw = lds_u32(wa);
      hv[0] = static_cast<long long>(cw & 0xffffu) * hq_const; hv[1] = static_cast<long long>(cw >> 16) * hq_const;
      if (cw != 0u) asm volatile("st.shared.u32 [%0], %1;" ::"r"(wa), "r"(0u) : "memory");
    } else {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const unsigned ca = cb + (2 * m + e) * 128 + 2 * kATable;
        const int hh = static_cast<int>(lds_u32(ca)), hl = static_cast<int>(lds_u32(ca + kATable));
        hv[e] = static_cast<long long>(hh) * 65536 + static_cast<long long>(static_cast<uint32_t>(hl));
        if (hh != 0 || hl != 0) {
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(ca), "r"(0u) : "memory");
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(ca + kATable), "r"(0u) : "memory");
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const long long gv = static_cast<long long>(hi[e]) * 65536 + static_cast<long long>(static_cast<uint32_t>(lo[e]));
      if (hi[e] != 0 || lo[e] != 0) {
        const unsigned ca = cb + (2 * m + e) * 128;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ca), "r"(0u) : "memory");
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ca + kATable), "r"(0u) : "memory");
      }
      if (gv != 0) atomicAdd(dst + 2 * (2 * m + e), static_cast<unsigned long long>(gv));
      if (hv[e] != 0) atomicAdd(dst + 2 * (2 * m + e) + 1, static_cast<unsigned long long>(hv[e]));
    }
  }
}

template <bool CH>
__global__ void __launch_bounds__(kAThreads, 1) k_hist_a(const HistAArgs a, const __grid_constant__ CUtensorMap tmap) {
  using S = AShape<CH>;
  constexpr int G = S::G, NS = S::kStages;
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_producer = warp >= kAConsumers;
  unsigned char* ring = smem + S::kTables;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + NS * S::kStageBytes);   // producer -> consumer: stage landed
  uint64_t* empty = full + NS;                                                // consumer -> producer: stage consumed
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 32); mbar_init(empty + i, 1); }
  }
  if (!is_producer) {
    // tables start (and, after every flush, return to) all-zero
    float4* z = reinterpret_cast<float4*>(smem);
    for (int i = threadIdx.x; i < S::kTables / 16; i += kAConsumers * 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  pdl_enter();          // everything above overlaps the predecessor's tail; everything below reads its results

  AWork w;
  if (!a_work_setup<CH>(a, &w)) return;
  const int32_t* ip = w.idx ? w.idx + w.begin : nullptr;

  if (is_producer) {
    // ------------------------------------------------------------------ producer warp: stage rows
    const int pw = warp - kAConsumers;
    const int pf = (ip != nullptr) ? a.l2_prefetch : 0;
    long long seq = 0;
    int pf_id = -1;                 // row id whose segment is prefetched into L2 on the next visit
    for (long long v = w.v_lo; v < w.v_hi; ++v, ++seq) {
      if (static_cast<int>(seq % kAProducers) != pw) continue;
      const int set = static_cast<int>(v / w.spc), st = static_cast<int>(v - static_cast<long long>(set) * w.spc);
      const int p0 = st * kARows;
      const int cnt = min(kARows, w.n - p0);
      const uint8_t* colbase = a.bins + static_cast<int64_t>(set) * S::kRowBytes;
      const int slot = static_cast<int>(seq % NS);
      const unsigned par = static_cast<unsigned>((seq / NS) & 1);
      unsigned char* sb = ring + slot * S::kStageBytes;
      if (pf > 0) {
        // walk the index list `pf` of this warp's stages ahead and pull every row's segment into L2 (fire-and-forget)
        if (pf_id >= 0) prefetch_l2(colbase + static_cast<int64_t>(pf_id) * a.pitch);
        const int pp = p0 + pf * kAProducers * kARows + lane;
        pf_id = (pp < w.n && st + pf * kAProducers < w.spc) ? __ldg(ip + pp) : -1;
      }
      if (a.use_tma && ip == nullptr && cnt == kARows) {
        // contiguous rows (root of an un-bagged tree): ONE 2-D TMA tile {32 G columns, 32 rows} of the row-major matrix
        // + one bulk copy of the 32 (g, h) pairs per stage; both complete on the stage's mbarrier
        mbar_wait_parked(empty + slot, par ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(full + slot, S::kStageBytes);
          tma_load_tile_2d(sb, &tmap, set * S::kRowBytes, p0, full + slot);
          tma_load_1d(sb + kARows * S::kRowBytes, a.gq + p0, kARows * 8, full + slot);
        } else {
          mbar_arrive(full + slot);
        }
        continue;
      }
      // gathered: lane l holds the id of row p0 + l; each row is 2 G chunks of 16 bytes
      int rid = -1;
      if (lane < cnt) rid = ip ? __ldg(ip + p0 + lane) : p0 + lane;
      mbar_wait_parked(empty + slot, par ^ 1);           // the consumer released this ring slot
#pragma unroll
      for (int i = 0; i < 2 * G; ++i) {
        const int c = lane + 32 * i;
        const int row = c / (2 * G), part = c % (2 * G);
        const int r = __shfl_sync(0xffffffffu, rid, row);
        if (r >= 0) cp_async16(sb + row * S::kRowBytes + part * 16, colbase + static_cast<int64_t>(r) * a.pitch + part * 16);
      }
      if (rid >= 0) cp_async8(sb + kARows * S::kRowBytes + lane * 8, w.gq_ord != nullptr ? w.gq_ord + p0 + lane : a.gq + rid);
      mbar_arrive_on_cp_async(full + slot);
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warp: accumulate
  const int cw = warp;
  const int q = lane & 7, j = lane >> 3;
  const unsigned tb0 = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  const unsigned ring0 = static_cast<unsigned>(__cvta_generic_to_shared(ring));
  // The lane's four cells of a unit: columns 4q + ((k + j) & 3), k = 0..3 (rotated by j so that the four lane groups
  // never share a bank), in the table set of column group (4u + j) % G = j % G — a per-lane constant.  sel[k] is the
  // PRMT selector that extracts the bin byte, cell[k] the shared-memory address of (bin 0, that column).
  unsigned sel[4], cell[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int kk = (k + j) & 3;
    sel[k] = 0x4440u | static_cast<unsigned>(kk);
    cell[k] = tb0 + static_cast<unsigned>(j % G) * S::kCgBytes + (4u * q + kk) * 4u;
  }
  const long long hq_const = CH ? a.ctl->h_const_q : 0;
  const int t = threadIdx.x;        // 0..255 among the consumers

  long long seq = 0;
  long long v = w.v_lo;
  while (v < w.v_hi) {
    const int set = static_cast<int>(v / w.spc);
    const int s0 = static_cast<int>(v - static_cast<long long>(set) * w.spc);
    const int s1 = static_cast<int>(min(static_cast<long long>(w.spc), s0 + (w.v_hi - v)));
    unsigned long long* dst_set = a.pool + static_cast<int64_t>(w.slot) * a.slot_stride +
                                  static_cast<int64_t>(set) * S::kRowBytes * (kBinsPerColumn * 2);
    int rows_acc = 0;
    for (int st = s0; st < s1; ++st, ++seq) {
      if (static_cast<int>(seq % kAConsumers) == cw) {
        const int slot = static_cast<int>(seq % NS);
        const unsigned par = static_cast<unsigned>((seq / NS) & 1);
        mbar_wait(full + slot, par);
        const int cnt = min(kARows, w.n - st * kARows);
        const unsigned sb = ring0 + slot * S::kStageBytes;
        // The stage is a flat array of 32-byte segments: segment s = (row s / G, column group s % G).  One unit = the
        // 128 contiguous bytes of four segments: lane (q, j) reads bytes 4q..4q+3 of segment 4u + j (one conflict-free
        // LDS.32 per unit) and owns those four cells.
#pragma unroll 2
        for (int u = 0; u < (kARows * G) / 4; ++u) {
          const int seg = 4 * u + j;
          const int row = seg / G;
          const int2 rd = lds_i2(sb + kARows * S::kRowBytes + row * 8);
          const uint32_t wv = lds_u32(sb + seg * kColGroup + q * 4);
          const int ghi = rd.x >> 16, glo = rd.x & 0xffff;
          const int hhi = rd.y >> 16, hlo = rd.y & 0xffff;
          if (row < cnt) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t b = __byte_perm(wv, 0u, sel[k]);
              const unsigned ca = cell[k] + (b << 7);
              red_s32(ca, ghi);
              red_s32(ca + kATable, glo);
              if (CH) {
                // 16-bit count field of bin b: word (b >> 1) of the count table, low or high half
                unsigned cc, inc;
                asm("{\n\t.reg .u32 t;\n\tand.b32 t, %2, 0xfe;\n\tmad.lo.u32 %0, t, 64, %3;\n\tand.b32 t, %2, 1;\n\tmad.lo.u32 %1, t, 0xffff, 1;\n\t}"
                    : "=r"(cc), "=r"(inc) : "r"(b), "r"(cell[k]));
                red_s32(cc + 2 * kATable, static_cast<int>(inc));
              }
              else { red_s32(ca + 2 * kATable, hhi); red_s32(ca + 3 * kATable, hlo); }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);
      }
      rows_acc += kARows;
      if (rows_acc >= kAFlushRows && st + 1 < s1) {
        consumer_bar_sync_a();
#pragma unroll
        for (int gs = 0; gs < G; ++gs)
          a_flush_zero<CH>(tb0 + gs * S::kCgBytes, dst_set + static_cast<int64_t>(gs) * kColGroup * (kBinsPerColumn * 2), t, hq_const);
        consumer_bar_sync_a();
        rows_acc = 0;
      }
    }
    consumer_bar_sync_a();
#pragma unroll
    for (int gs = 0; gs < G; ++gs)
      a_flush_zero<CH>(tb0 + gs * S::kCgBytes, dst_set + static_cast<int64_t>(gs) * kColGroup * (kBinsPerColumn * 2), t, hq_const);
    consumer_bar_sync_a();
    v += (s1 - s0);
  }
}

}  // namespace b200
