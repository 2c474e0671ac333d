// hist_common.cuh — shared device helpers of the histogram kernels (sm_100a): cp.async / mbarrier / TMA wrappers and
// the work decomposition of the packed-cell quantized kernel (hist_q_kernel.cuh).
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "comm.cuh"
#include "types.cuh"

namespace b200 {

constexpr int kHistWarps = 3;                    // consumer warps (one per SMSP 0..2), each with a partner producer warp
constexpr int kHistThreads = 2 * kHistWarps * 32; // warps 0..2 consume, warps 3..5 produce (warp 3 owns SMSP 3)
constexpr int kStageRows = 32;
constexpr int kHistBatch = 4;                    // rows per in-register RMW batch of k_hist_q
constexpr int kStageBinBytes = kStageRows * kColGroup;            // 1024

struct HistArgs {
  const uint8_t* bins;            // [num_data x pitch] row-major stored values
  int64_t pitch;                  // bytes per row, multiple of 32
  const int32_t* idx0;            // ping-pong row-index buffers
  const int32_t* idx1;
  const Leaf* leaves;
  const Ctl* ctl;
  unsigned long long* pool;       // int64 fixed-point histogram pool [slot][column][256][2]
  int64_t slot_stride;            // int64 elements per slot
  int32_t num_colgroups;          // ceil(num_columns / 32)
  int32_t min_rows_per_item;      // do not split a column group over more warps than n / this
  int32_t use_tma;                // 1: contiguous (root, un-bagged) stages are staged by TMA tile copies
  const int32_t* ghqo0;           // leaf-ordered packed quantized words (k_hist_q), or nullptr
  const int32_t* ghqo1;
  int32_t l2_prefetch;            // > 0: gathered passes prefetch the bin sectors of the stage this many stages ahead into L2
  int32_t map_mode;               // 0: items dealt column-group-major; 1: one CTA = (column group, 3 row parts), adjacent CTAs = adjacent column groups
  // explicit mode (stand-alone ConstructHistogram hook): explicit_n >= 0
  int32_t explicit_n;
  int32_t explicit_slot;
  const int32_t* explicit_idx;    // nullptr = identity
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
constexpr int kPfIdLead = 2;          // producer iterations between loading a future stage's row ids and prefetching its rows
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- mbarrier helpers (shared::cta) -----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
}
// all prior cp.async of this thread arrive on `bar` when they complete (count pre-accounted at init)
__device__ __forceinline__ void mbar_arrive_on_cp_async(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  unsigned done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (spin > (1u << 28)) __trap();          // watchdog: turn a protocol bug into a launch error, not a hang
  }
}

// Producer-side wait: the same protocol, but the try_wait carries a suspend-time hint so that a producer that has
// filled its ring parks in hardware instead of spinning (ncu on k_hist_q: the spin of the three producer warps was
// 40 % of all issued instructions and took issue slots from the consumer warps that share their SMSPs).
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, unsigned parity) {
  const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  unsigned done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
    if (spin > (1u << 24)) __trap();
  }
}

// ---- TMA (bulk async copies, completion by mbarrier complete_tx) ---------------------------------------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(bytes) : "memory");
}
// 2-D tile [32 rows x 32 columns] of the row-major bin matrix -> shared memory (UTMALDG)
__device__ __forceinline__ void tma_load_tile_2d(void* smem_dst, const CUtensorMap* tmap, int col, int row, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(tmap), "r"(col), "r"(row),
                 "r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
}
// contiguous bytes -> shared memory (UBLKCP); size multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc), "r"(bytes),
                 "r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
}

// Work decomposition shared by the producer and the consumer warp of a pair.
// Items are numbered column-group-major (item j -> column group j / splits, row part j % splits) and dealt to
// warp slots in order, so the three consumer warps of a CTA normally work on consecutive row parts of the SAME
// column group and can merge their private tables in shared memory before touching global memory.
struct HistWork {
  int n, begin, slot;
  const int32_t* idx;
  const int32_t* ghq_ord;        // same for the packed quantized word (k_hist_q)
  int CG, splits, per, items, total_warps;
  int mode, rounds_total;        // mode 1: rounds_total = CG * ceil(splits/3) virtual CTAs
};

__device__ __forceinline__ bool hist_work_setup(const HistArgs& a, HistWork* w) {
  if (a.explicit_n >= 0) {
    w->n = a.explicit_n; w->begin = 0; w->slot = a.explicit_slot; w->idx = a.explicit_idx;
    w->ghq_ord = nullptr;
  } else {
    const Ctl* c = a.ctl;
    if (!c->cur_valid || !c->do_find) return false;
    const Leaf& L = a.leaves[c->smaller];
    w->n = L.lcount; w->begin = L.begin; w->slot = L.slot;
    // the root of an un-bagged tree is the identity list: skip the index load altogether
    w->idx = (c->num_leaves == 1 && c->root_identity) ? nullptr : (L.buf ? a.idx1 : a.idx0);
    // every non-root segment was written by its parent's scatter, together with its (g,h) copy
    w->ghq_ord = (c->num_leaves > 1 && a.ghqo0 != nullptr) ? (L.buf ? a.ghqo1 : a.ghqo0) + L.begin : nullptr;
  }
  if (w->n <= 0) return false;
  w->total_warps = gridDim.x * kHistWarps;
  w->CG = a.num_colgroups;
  const int max_splits = max(1, w->total_warps / w->CG);
  // row parts per column group: balance the per-row work (~17 ns per row and warp) against the per-item fixed
  // cost (zeroing + the merged flush, ~16K/3 global atomics per item at ~0.2 G atomics/us):
  // t(s) = n/s * t_row + CG*s * t_item  =>  s* = sqrt(n * t_row / (CG * t_item)) ~ sqrt(0.7 n / CG)
  int sp = static_cast<int>(sqrtf(0.7f * static_cast<float>(w->n) / static_cast<float>(w->CG)));
  sp = max(1, min(sp, (w->n + a.min_rows_per_item - 1) / a.min_rows_per_item));
  // mapping choice: the CTA-per-(column group, row triple) mapping unless it would leave >5 % of the warps idle
  // on a leaf big enough to use them all (e.g. 32 column groups: 4 triples = 12 parts vs 13 parts column-group-major;
  // measured 8-13 % slower on 4M x 1024)
  const int max_triples0 = max(1, static_cast<int>(gridDim.x) / w->CG);
  w->mode = a.map_mode;
  if (w->mode == 1 && sp > max_triples0 * kHistWarps && max_triples0 * kHistWarps * 20 < max_splits * 19) w->mode = 0;
  if (w->mode == 1) {
    // one CTA = one column group x three consecutive row parts (always mergeable); CTAs b, b+1 work on adjacent
    // column groups of the SAME rows at the same time, so the two 32-byte sectors of a 64-byte DRAM atom are
    // fetched once (the second hits in L2).  One wave: at most floor(grid / CG) triples per column group.
    const int max_triples = max(1, static_cast<int>(gridDim.x) / w->CG);
    const int triples = min(max_triples, (sp + kHistWarps - 1) / kHistWarps);
    sp = min(sp, triples * kHistWarps);
    if (sp > kHistWarps) sp = triples * kHistWarps;          // full triples
    w->rounds_total = w->CG * triples;
  } else {
    sp = min(max_splits, sp);
    // whole CTAs per column group => every CTA can merge its tables (not when the warp count is the limit:
    // there the rows per warp matter more than the flush)
    if (sp >= kHistWarps && sp < max_splits) sp -= sp % kHistWarps;
    w->rounds_total = 0;
  }
  w->splits = sp;
  w->per = (((w->n + w->splits - 1) / w->splits) + 31) & ~31;
  w->items = w->CG * w->splits;
  return true;
}

// item of (round r, warp pair p): returns false if this warp has nothing to do in this round
struct HistItem { int cg, part; bool valid, merge; };
__device__ __forceinline__ bool hist_round_valid(const HistWork& w, int round) {
  return w.mode == 1 ? (static_cast<int>(blockIdx.x) + round * static_cast<int>(gridDim.x) < w.rounds_total)
                     : (static_cast<int>(blockIdx.x) * kHistWarps + round * w.total_warps < w.items);
}
__device__ __forceinline__ HistItem hist_item(const HistWork& w, int round, int pair) {
  HistItem it;
  if (w.mode == 1) {
    const int v = blockIdx.x + round * gridDim.x;
    it.cg = v % w.CG;
    const int t = v / w.CG;
    it.part = t * kHistWarps + pair;
    it.valid = it.part < w.splits;
    it.merge = (t * kHistWarps + kHistWarps - 1) < w.splits;
  } else {
    const int base = blockIdx.x * kHistWarps + round * w.total_warps;
    const int item = base + pair;
    it.valid = item < w.items;
    it.cg = item / w.splits; it.part = item % w.splits;
    it.merge = (base + kHistWarps - 1 < w.items) && (base / w.splits == (base + kHistWarps - 1) / w.splits);
  }
  return it;
}

__device__ __forceinline__ void consumer_bar_sync() {          // the 3 consumer warps only (named barrier 1)
  asm volatile("bar.sync 1, %0;" ::"n"(kHistWarps * 32) : "memory");
}

}  // namespace b200
