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
//     before it can wrap; the tables are flushed (dumped as raw int32 images into a scratch block; k_hist_reduce forms
//     (hi << 16) + lo in int64 and sums the blocks into the pool slot) at least that often.  Cost: 2 ATOMS per cell and
//     component.
//   * Constant hessian (unweighted L2: the BASELINE configs) needs no hessian sums at all: like the reference
//     (dataset.cpp:1430-1437) the kernel COUNTS rows per cell and scales at the flush.  Counts are 16-bit fields, two
//     bins per int32 word.  3 ATOMS per cell (g hi, g lo, count); general hessians: 4 (g hi, g lo, h hi, h lo).
//   * The tables are shared by ALL consumer warps of the CTA ([bin][column] int32, bank = column): one table set per
//     column group per SM instead of one per warp, so 16 consumer warps (4 per SMSP) hide each other's latencies and the
//     shared memory left over becomes a 28..48-stage ring (60 KB of rows in flight per SM for gathered leaves).
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
constexpr int kAConsumers = 16;            // consumer warps (4 per SMSP: the per-cell ALU chains are latency-bound with fewer)
constexpr int kAProducers = 4;             // producer warps (1 per SMSP)
constexpr int kAThreads = (kAConsumers + kAProducers) * 32;
constexpr int kATable = kBinsPerColumn * 32 * 4;      // one [bin][column] int32 table = 32 KB
constexpr int kAFlushRows = 65504;         // rows a table set may take between flushes (multiple of 32, <= 65535)

// CH = constant hessian (count + scale), else general hessians
template <bool CH>
struct AShape {
  static constexpr int G = CH ? 2 : 1;                                   // column groups per CTA
  static constexpr int kLineSets = 128 / (kColGroup * G);                // column-group sets per 128-byte line of a row
  static constexpr int kCgBytes = CH ? (2 * kATable + kATable / 2) : 4 * kATable;   // 80 KB : 128 KB per column group
  static constexpr int kTables = G * kCgBytes;                           // 160 KB : 128 KB
  static constexpr int kRowBytes = kColGroup * G;                        // bin bytes per staged row
  static constexpr int kStageBytes = kARows * kRowBytes + kARows * 8;    // + int2 (g, h) per row: 2304 : 1280
  static constexpr int kStages = CH ? 28 : 48;                           // ring depth
  static constexpr int kSmem = kTables + kStages * kStageBytes + kStages * 2 * 8 + 16;
};
static_assert(AShape<true>::kSmem <= 232448 && AShape<false>::kSmem <= 232448, "exceeds 227 KB of shared memory per CTA");
static_assert(AShape<true>::kStages > kAConsumers && AShape<false>::kStages > kAConsumers, "ring must hold more stages than there are consumers");

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
  unsigned long long* pool;       // int64 fixed-point histogram pool [slot][column][256][2] (written by k_hist_reduce)
  int64_t slot_stride;            // int64 elements per slot
  // Flush path: a CTA never adds into the pool itself.  It DUMPS its raw int32 tables (the shared-memory image of one
  // column-group set) as one block into `scratch[set][blk]`, blk = atomicAdd(blk_count[epoch][set], 1), and
  // k_hist_reduce sums the blocks of every set into the leaf's pool slot with plain loads and stores.  Integer sums:
  // the order in which blocks were claimed does not matter.  (Round-2 measurement: flushing with RED.ADD.64 cost
  // ~0.6 us of L2 atomic time per CTA, 22 % of the 4M x 1024 root pass and most of a 50K-row pass.)
  unsigned char* scratch;         // [sets][blk_cap] blocks of AShape::kTables bytes
  int32_t* blk_count;             // [num_leaves][sets] blocks claimed per (histogram pass of the tree, set); zeroed once per tree
  int32_t blk_cap;
  int32_t num_colgroups;          // pitch / 32
  float split_k;                  // work-split constant: CTA groups = sqrt(split_k * rows * set groups), see a_work_setup
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
  long long v_lo, v_hi;          // this CTA's range of virtual stages (set-GROUP-major, stage-minor)
  int member;                    // this CTA's set inside a set group: set = kLineSets * (v / spc) + member
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
  // The sets that share a 128-byte line of a row (2 sets of 64 columns, 4 of 32) form a set GROUP, and kLineSets
  // adjacent CTAs form a CTA group that walks the same (set group, stage) range in step, each on its own member set:
  // the line a row segment lives in is then fetched from DRAM once and found in L2 by the other members.  (Without
  // this ncu showed DRAM reads of 2.0x the algorithmic bytes: L2 fills whole 128-byte lines, and a CTA working alone on
  // 64 of those bytes had long gone when another CTA came for the other half.)
  constexpr int Q = AShape<CH>::kLineSets;
  const int sets = a.num_colgroups / AShape<CH>::G;
  const int set_groups = (sets + Q - 1) / Q;
  w->spc = (w->n + kARows - 1) / kARows;
  const long long total = static_cast<long long>(set_groups) * w->spc;
  // CTA groups that take part.  Accumulating costs ~t_row per (row, set group) and is divided by P; every group adds a
  // fixed cost t_blk (its table dumps + their share of k_hist_reduce): T(P) = W t_row / P + P t_blk, W = n * set_groups,
  // is smallest at P* = sqrt(W t_row / t_blk).  a.split_k = t_row / t_blk (measured ~0.044 for the constant-hessian
  // kernel: 5.6 ns per row against 128 ns per group).  Never fewer than one group per set group, at most grid / Q.
  long long P = static_cast<long long>(sqrtf(a.split_k * static_cast<float>(w->n) * static_cast<float>(set_groups)) + 0.5f);
  if (P < set_groups) P = set_groups;
  const long long max_groups = static_cast<long long>(gridDim.x) / Q;
  if (P > max_groups) P = max_groups;
  const int grp = static_cast<int>(blockIdx.x) / Q;
  w->member = static_cast<int>(blockIdx.x) % Q;
  if (grp >= P) return false;
  w->v_lo = total * grp / P;
  w->v_hi = total * (grp + 1) / P;
  // the last set group may be partial: a member whose set does not exist skips it
  if (Q * (set_groups - 1) + w->member >= sets) w->v_hi = min(w->v_hi, static_cast<long long>(set_groups - 1) * w->spc);
  return w->v_hi > w->v_lo;
}

// Dump the CTA's tables (the whole shared-memory image of one column-group set) into a scratch block and zero them.
// Called by all consumer threads between two consumer barriers: 16-byte LDS -> STG + STS(0), fully coalesced.
template <bool CH>
__device__ __forceinline__ void a_dump_zero(unsigned char* smem, unsigned char* block, int t) {
  const uint4* src = reinterpret_cast<const uint4*>(smem);
  uint4* dst = reinterpret_cast<uint4*>(block);
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 4
  for (int i = t; i < AShape<CH>::kTables / 16; i += kAConsumers * 32) {
    const uint4 v = src[i];
    __stcg(dst + i, v);
    if ((v.x | v.y | v.z | v.w) != 0u) const_cast<uint4*>(src)[i] = zero;
  }
}

// The epoch of a histogram pass within its tree: one row of blk_count per pass (explicit mode uses row 0).
__device__ __forceinline__ int a_epoch(const HistAArgs& a) { return a.explicit_n >= 0 ? 0 : a.ctl->num_leaves - 1; }

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
    const int total = static_cast<int>(w.v_hi - w.v_lo);          // stages of this CTA, numbered seq = 0..total-1
    int sgrp = static_cast<int>((w.v_lo + pw) / w.spc);                    // set group; this CTA's set = kLineSets * sgrp + member
    int st = static_cast<int>((w.v_lo + pw) - static_cast<long long>(sgrp) * w.spc);
    int slot = pw % NS; unsigned par = 0;
    int pf_id = -1;                 // row id whose segment is prefetched into L2 on the next visit
    for (int seq = pw; seq < total; seq += kAProducers) {
      const int p0 = st * kARows;
      const int cnt = min(kARows, w.n - p0);
      const int set = S::kLineSets * sgrp + w.member;
      const uint8_t* colbase = a.bins + static_cast<int64_t>(set) * S::kRowBytes;
      unsigned char* sb = ring + slot * S::kStageBytes;
      if (pf > 0) {
        // walk the index list `pf` of this warp's stages ahead and pull every row's segment into L2 (fire-and-forget)
        if (pf_id >= 0) prefetch_l2(colbase + static_cast<int64_t>(pf_id) * a.pitch);
        const int pp = p0 + pf * kAProducers * kARows + lane;
        pf_id = (pp < w.n) ? __ldg(ip + pp) : -1;
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
      } else {
        // gathered: lane l holds the id of row p0 + l; each row is 2 G chunks of 16 bytes
        int rid = -1;
        if (lane < cnt) rid = ip ? __ldg(ip + p0 + lane) : p0 + lane;
        mbar_wait_parked(empty + slot, par ^ 1);           // the consumer released this ring slot
        // (TMA tile::gather4 — four row ids per instruction, UTMALDG.2D.GATHER4 — was built, parity-checked and timed here
        // in round 2: 1.84 ms against 1.62 ms for these cp.async on a 2M-row gathered pass of 4M x 1024, 3.34 against
        // 2.70 ms for the general kernel: 32/64-byte rows are too small for the TMA unit.  Not kept.)
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
      // next stage of this warp: kAProducers further along the (set, stage) sequence and the ring
      st += kAProducers;
      while (st >= w.spc) { st -= w.spc; ++sgrp; }
      slot += kAProducers;
      if (slot >= NS) { slot -= NS; par ^= 1; }
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
  const int sets = a.num_colgroups / G, epoch = a_epoch(a);
  volatile int* s_blk = reinterpret_cast<volatile int*>(empty + NS);      // 4 bytes behind the barriers
  const int t = threadIdx.x;        // consumers are warps 0..kAConsumers-1

  // One stage = 8 G units of 128 bytes (four 32-byte segments; segment s = row s / G, column group s % G): lane (q, j)
  // reads bytes 4q..4q+3 of segment 4u + j with one conflict-free LDS.32 and owns those four cells.
  auto accumulate = [&](unsigned sb, int cnt) {
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
          } else {
            red_s32(ca + 2 * kATable, hhi); red_s32(ca + 3 * kATable, hlo);
          }
        }
      }
    }
  };

  // The CTA's stages are numbered seq = 0..total-1 along (set, stage); consumer warp cw takes seq = cw, cw + 16, ...
  // The sequence is cut into segments at set boundaries and every kAFlushRows rows; after each segment all consumer
  // warps meet, dump the tables into a scratch block of the set and zero them.
  const int total = static_cast<int>(w.v_hi - w.v_lo);
  int my_seq = cw, slot = cw % NS; unsigned par = 0;
  int sgrp = static_cast<int>(w.v_lo / w.spc);
  int st0 = static_cast<int>(w.v_lo - static_cast<long long>(sgrp) * w.spc);     // first stage of the current segment
  int seq0 = 0;
  while (seq0 < total) {
    const int len = min(min(w.spc - st0, total - seq0), kAFlushRows / kARows);
    const int seq1 = seq0 + len;
    for (; my_seq < seq1; my_seq += kAConsumers) {
      const int st = st0 + (my_seq - seq0);
      mbar_wait(full + slot, par);
      accumulate(ring0 + slot * S::kStageBytes, min(kARows, w.n - st * kARows));
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + slot);
      slot += kAConsumers;
      if (slot >= NS) { slot -= NS; par ^= 1; }
    }
    // all consumer warps are done with the segment: claim a scratch block of this set, dump the tables, zero them
    const int set = S::kLineSets * sgrp + w.member;
    if (t == 0) *s_blk = atomicAdd(a.blk_count + epoch * sets + set, 1);
    consumer_bar_sync_a();
    {
      const int blk = *s_blk;
      if (blk < a.blk_cap) a_dump_zero<CH>(smem, a.scratch + (static_cast<int64_t>(set) * a.blk_cap + blk) * S::kTables, t);
    }
    consumer_bar_sync_a();
    seq0 = seq1; st0 += len;
    if (st0 >= w.spc) { st0 = 0; ++sgrp; }
  }
}

// Sum the scratch blocks of every column-group set into the leaf's pool slot (plain stores: the slot needs no memset).
// A 32-column row of cells — CH -> (set, column group gs, bin pair m), general -> (set, bin); lane = column — is summed by
// `wpr` warps that deal the set's blocks among themselves (4 load groups in flight each) and meet in shared memory.  (The first version gave every thread ALL blocks of its set: a chain of up to 74 dependent L2 round trips on
// 64 CTAs when a GPU holds only two sets — 10M x 128, the per-GPU shard of C3 at 8 GPUs — which cost more than the
// accumulation itself.)
constexpr int kReduceWarps = 8;
// `wpr` (1, 2, 4 or 8, chosen by the host so that the grid has a few hundred CTAs) warps share one 32-column row of
// cells; a CTA of 8 warps therefore covers 8 / wpr rows.
template <bool CH>
__global__ void __launch_bounds__(kReduceWarps * 32) k_hist_reduce(const HistAArgs a, const int wpr) {
  using S = AShape<CH>;
  constexpr int G = S::G;
  pdl_enter();
  int slot;
  if (a.explicit_n >= 0) slot = a.explicit_slot;
  else {
    const Ctl* c = a.ctl;
    if (!c->cur_valid || !c->do_find) return;
    slot = a.leaves[c->smaller].slot;        // a rank holding no row of the leaf (row-shard) claims no block: zeros are written
  }
  const int sets = a.num_colgroups / G, epoch = a_epoch(a);
  constexpr int kRowsPerSet = CH ? G * 128 : kBinsPerColumn;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = warp % wpr;                                               // this warp's share of the blocks
  const int grow = blockIdx.x * (kReduceWarps / wpr) + warp / wpr;          // global row of cells
  const int set = grow / kRowsPerSet, r = grow % kRowsPerSet;
  const bool live = set < sets;
  const int nblk = live ? min(a.blk_count[epoch * sets + set], a.blk_cap) : 0;
  const uint32_t* base = reinterpret_cast<const uint32_t*>(a.scratch + static_cast<int64_t>(live ? set : 0) * a.blk_cap * S::kTables);
  constexpr int kBlkWords = S::kTables / 4, kTabWords = kATable / 4;
  unsigned long long* dst_slot = a.pool + static_cast<int64_t>(slot) * a.slot_stride;
  __shared__ long long s_part[kReduceWarps][4][32];
  long long v0 = 0, v1 = 0, v2 = 0, v3 = 0;      // CH: g(2m), g(2m+1), count(2m), count(2m+1); general: g, h, -, -
  if (CH) {
    const int gs = r / 128, m = r % 128;
    const uint32_t* p0 = base + gs * (S::kCgBytes / 4) + (2 * m) * 32 + lane;
#pragma unroll 4
    for (int b = sub; b < nblk; b += wpr) {
      const uint32_t* p = p0 + static_cast<int64_t>(b) * kBlkWords;
      const int h0 = static_cast<int>(__ldcg(p)), h1 = static_cast<int>(__ldcg(p + 32));
      const uint32_t l0 = __ldcg(p + kTabWords), l1 = __ldcg(p + kTabWords + 32);
      const uint32_t cw = __ldcg(p + 2 * kTabWords - m * 32);        // count word m of the [128][32] count table
      v0 += static_cast<long long>(h0) * 65536 + l0; v1 += static_cast<long long>(h1) * 65536 + l1;
      v2 += cw & 0xffffu; v3 += cw >> 16;
    }
  } else {
    const uint32_t* p0 = base + r * 32 + lane;
#pragma unroll 4
    for (int b = sub; b < nblk; b += wpr) {
      const uint32_t* p = p0 + static_cast<int64_t>(b) * kBlkWords;
      v0 += static_cast<long long>(static_cast<int>(__ldcg(p))) * 65536 + __ldcg(p + kTabWords);
      v1 += static_cast<long long>(static_cast<int>(__ldcg(p + 2 * kTabWords))) * 65536 + __ldcg(p + 3 * kTabWords);
    }
  }
  if (wpr > 1) {
    s_part[warp][0][lane] = v0; s_part[warp][1][lane] = v1; s_part[warp][2][lane] = v2; s_part[warp][3][lane] = v3;
    __syncthreads();
    if (sub != 0) return;
    for (int w = 1; w < wpr; ++w) { v0 += s_part[warp + w][0][lane]; v1 += s_part[warp + w][1][lane]; v2 += s_part[warp + w][2][lane]; v3 += s_part[warp + w][3][lane]; }
  }
  if (!live) return;
  if (CH) {
    const int gs = r / 128, m = r % 128;
    const long long hq = a.ctl->h_const_q;
    longlong2* d = reinterpret_cast<longlong2*>(dst_slot + (static_cast<int64_t>(set * G + gs) * kColGroup + lane) * (kBinsPerColumn * 2) + 4 * m);
    d[0] = make_longlong2(v0, v2 * hq);
    d[1] = make_longlong2(v1, v3 * hq);
  } else {
    *reinterpret_cast<longlong2*>(dst_slot + (static_cast<int64_t>(set) * kColGroup + lane) * (kBinsPerColumn * 2) + 2 * r) = make_longlong2(v0, v1);
  }
}

}  // namespace b200
