// hist_q_kernel.cuh — histogram construction for quantized-gradient training with PACKED 32-bit cells (sm_100a).
//
// Replaces, for Config::use_quantized_grad: the reference's 16-bit-per-field integer histograms
// (DenseBin::ConstructHistogramInt16 src/io/dense_bin.hpp:174-222, MultiValDenseBin::ConstructHistogramIntInner
// src/io/multi_val_dense_bin.hpp:104-160; CUDA: cuda_histogram_constructor.cu:252-513).
//
// Same machine as k_hist (hist_kernel.cuh): lane owns a column, warp owns a private [bin][lane] table, producer warps
// stage rows with cp.async / TMA through an mbarrier ring — but a cell is ONE int32 = (gradient sum << 16) + hessian
// sum, and a row contributes one precomputed packed word (k_quantize).  Per 32 cell updates that is LDS.32 + STS.32
// = 2 shared-memory wavefronts instead of the 4 of the 8-byte (g,h) cell, and one broadcast LDS.128 brings the packed
// words of 4 rows.  k_hist is bound by shared-memory wavefronts, so this is the lever quantized
// training was expected to give.  The table is 32 KB per warp: 2 CTAs (6 consumer warps) fit per SM.
//
// Exactness: with hessians in [0, Q] and |gradients| <= Q/2 per row (Q = num_grad_quant_bins), a cell cannot leave
// its 16-bit fields within R = floor(65535 / Q) rows; every consumer warp flushes its table into the int64 pool
// (RED.ADD.64 of the two unpacked fields) after at most R rows.  The hessian field is unsigned, the gradient field
// signed: P = G * 65536 + H  =>  H = P & 0xffff, G = (P - H) >> 16.
#pragma once
#include "hist_common.cuh"

namespace b200 {

constexpr int kQStages = 4;
constexpr int kQTableBytes = kBinsPerColumn * 32 * 4;                 // 32768
constexpr int kQStageBytes = kStageBinBytes + kStageRows * 4;         // 1024 + 128 packed (g,h) words
constexpr int kQWarpSmemBytes = kQTableBytes + kQStages * kQStageBytes;
constexpr int kQBarBytes = kHistWarps * kQStages * 2 * 8;
constexpr int kQSmemBytes = kHistWarps * kQWarpSmemBytes + kQBarBytes;   // 112320 B: two CTAs per SM

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ int ldsi32(unsigned addr) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void stsi32(unsigned addr, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v)); }

__device__ __forceinline__ void q_flush_cell(unsigned long long* dst, int b, int G, int H) {
  if (G != 0) atomicAdd(dst + 2 * b, static_cast<unsigned long long>(static_cast<long long>(G)));
  if (H != 0) atomicAdd(dst + 2 * b + 1, static_cast<unsigned long long>(static_cast<long long>(H)));
}

struct HistQArgs {
  HistArgs h;                     // bins, index buffers, leaves, pool, work mapping (gh unused)
  const int32_t* ghq;             // [num_data] packed (g << 16) + h of the discretized gradients
  int32_t flush_rows;             // R rounded down to a multiple of kStageRows
};

__global__ void __launch_bounds__(kHistThreads, 2) k_hist_q(const HistQArgs qa, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem[];
  const HistArgs& a = qa.h;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_producer = warp >= kHistWarps;
  const int pair = is_producer ? warp - kHistWarps : warp;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kHistWarps * kQWarpSmemBytes);
  uint64_t* full = bars + pair * (2 * kQStages);
  uint64_t* empty = full + kQStages;
  if (!is_producer && lane == 0) {
    for (int i = 0; i < kQStages; ++i) { mbar_init(full + i, 32); mbar_init(empty + i, 1); }
  }
  __syncthreads();
  pdl_enter();

  HistWork w;
  if (!hist_work_setup(a, &w)) return;
  unsigned char* wbase = smem + pair * kQWarpSmemBytes;
  unsigned char* ring = wbase + kQTableBytes;

  if (is_producer) {
    int slot = 0; unsigned phase = 0;
    for (int round = 0; hist_round_valid(w, round); ++round) {
      const HistItem it = hist_item(w, round, pair);
      if (!it.valid) continue;
      const int cg = it.cg, part = it.part;
      const int r0 = part * w.per;
      const int r1 = min(w.n, r0 + w.per);
      if (r0 >= r1) continue;
      const uint8_t* colbase = a.bins + static_cast<int64_t>(cg) * kColGroup;
      const int half = (lane & 1) * 16;
      const int32_t* ip = w.idx ? w.idx + w.begin : nullptr;
      // Gathered passes: the ring holds only ~5 stages (6 KB) per warp, too little to cover the latency tail of 96
      // scattered requests per stage (ncu: consumers starved 28 % of the time, shared pipe 44 % busy vs 66 % at the
      // root).  The producer therefore also walks the index list `l2_prefetch` stages ahead and pulls every row's
      // 32-byte bin sector into L2 (fire-and-forget, no shared memory needed); the later cp.async hits L2.
      const int pf = (ip != nullptr) ? a.l2_prefetch : 0;
      int pfq[kPfIdLead];
#pragma unroll
      for (int d = 0; d < kPfIdLead; ++d) pfq[d] = -1;
      for (int p0 = r0; p0 < r1; p0 += kStageRows) {
        if (pf > 0) {
          if (pfq[0] >= 0) prefetch_l2(colbase + static_cast<int64_t>(pfq[0]) * a.pitch);
#pragma unroll
          for (int d = 0; d + 1 < kPfIdLead; ++d) pfq[d] = pfq[d + 1];
          const int pp = p0 + (pf + kPfIdLead) * kStageRows + lane;
          pfq[kPfIdLead - 1] = (pp < r1) ? __ldg(ip + pp) : -1;
        }
        if (a.use_tma && ip == nullptr && p0 + kStageRows <= r1) {
          mbar_wait_parked(empty + slot, phase ^ 1);
          unsigned char* sb = ring + slot * kQStageBytes;
          if (lane == 0) {
            mbar_arrive_expect_tx(full + slot, kQStageBytes);
            tma_load_tile_2d(sb, &tmap, cg * kColGroup, p0, full + slot);
            tma_load_1d(sb + kStageBinBytes, qa.ghq + p0, kStageRows * 4, full + slot);
          } else {
            mbar_arrive(full + slot);
          }
          if (++slot == kQStages) { slot = 0; phase ^= 1; }
          continue;
        }
        const int pa = p0 + (lane >> 1), pb = pa + 16, pg = p0 + lane;
        int ra = -1, rb = -1, rg = -1;
        if (pa < r1) ra = ip ? __ldg(ip + pa) : pa;
        if (pb < r1) rb = ip ? __ldg(ip + pb) : pb;
        if (pg < r1) rg = (w.ghq_ord != nullptr) ? pg : (ip ? __ldg(ip + pg) : pg);
        mbar_wait_parked(empty + slot, phase ^ 1);
        unsigned char* sb = ring + slot * kQStageBytes;
        if (ra >= 0) cp_async16(sb + (lane >> 1) * kColGroup + half, colbase + static_cast<int64_t>(ra) * a.pitch + half);
        if (rb >= 0) cp_async16(sb + (16 + (lane >> 1)) * kColGroup + half, colbase + static_cast<int64_t>(rb) * a.pitch + half);
        if (rg >= 0) cp_async4(sb + kStageBinBytes + lane * 4, (w.ghq_ord != nullptr ? w.ghq_ord : qa.ghq) + rg);
        mbar_arrive_on_cp_async(full + slot);
        if (++slot == kQStages) { slot = 0; phase ^= 1; }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warp
  const unsigned hbase = static_cast<unsigned>(__cvta_generic_to_shared(reinterpret_cast<int*>(wbase) + lane));
  const unsigned hbase0 = static_cast<unsigned>(__cvta_generic_to_shared(reinterpret_cast<int*>(smem) + lane));
  const int R = qa.flush_rows;
  int slot = 0; unsigned phase = 0;
  for (int round = 0; hist_round_valid(w, round); ++round) {
    const HistItem it = hist_item(w, round, pair);
    const bool merge = it.merge;                    // CTA-uniform
    if (!it.valid) continue;
    const int cg = it.cg, part = it.part;
    const int r0 = part * w.per;
    const int r1 = min(w.n, r0 + w.per);
    if (r0 >= r1 && !merge) continue;
    unsigned long long* dst = a.pool + static_cast<int64_t>(w.slot) * a.slot_stride +
                              (static_cast<int64_t>(cg) * kColGroup + lane) * (kBinsPerColumn * 2);
    // flush intervals: the same count for the three warps of a merging CTA (w.per is common to all parts)
    const int rows_here = max(0, r1 - r0);
    const int n_int = max(1, ((merge ? w.per : rows_here) + R - 1) / R);
    for (int iv = 0; iv < n_int; ++iv) {
      {
        float4* z = reinterpret_cast<float4*>(wbase);
#pragma unroll 8
        for (int i = lane; i < kQTableBytes / 16; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncwarp();
      const int i0 = min(r1, r0 + iv * R), i1 = min(r1, i0 + R);
      for (int p0 = i0; p0 < i1; p0 += kStageRows) {
        mbar_wait(full + slot, phase);
        const int cnt = min(kStageRows, i1 - p0);
        const unsigned char* sb = ring + slot * kQStageBytes;
        const unsigned char* sbin = sb + lane;
        const int* sq = reinterpret_cast<const int*>(sb + kStageBinBytes);
        if (cnt == kStageRows) {
          constexpr int K = kHistBatch;
          uint32_t nb[K]; int nq[K];
          unsigned addrN[K]; int sN[K];
          auto fetch = [&](int r) {
#pragma unroll
            for (int i = 0; i < K; ++i) nb[i] = sbin[(r + i) * 32];
            const int4 t = *reinterpret_cast<const int4*>(sq + r);
            nq[0] = t.x; nq[1] = t.y; nq[2] = t.z; nq[3] = t.w;
          };
          auto prepare = [&]() {
#pragma unroll
            for (int i = 0; i < K; ++i) {
              addrN[i] = hbase + (nb[i] << 7);
              sN[i] = nq[i];
#pragma unroll
              for (int j = 0; j < i; ++j) if (nb[j] == nb[i]) sN[i] += nq[j];
            }
          };
          static_assert(K == 4, "the packed words of a batch are fetched with one LDS.128");
          fetch(0);
          prepare();
          fetch(K);
#pragma unroll
          for (int r = 0; r < kStageRows; r += K) {
            unsigned addrC[K]; int sC[K], v[K];
#pragma unroll
            for (int i = 0; i < K; ++i) { addrC[i] = addrN[i]; sC[i] = sN[i]; }
#pragma unroll
            for (int i = 0; i < K; ++i) v[i] = ldsi32(addrC[i]);
            if (r + K < kStageRows) {
              prepare();
              if (r + 2 * K < kStageRows) fetch(r + 2 * K);
            }
#pragma unroll
            for (int i = 0; i < K; ++i) stsi32(addrC[i], v[i] + sC[i]);
          }
        } else {
          for (int r = 0; r < cnt; ++r) {
            const unsigned addr = hbase + (static_cast<uint32_t>(sbin[r * 32]) << 7);
            stsi32(addr, ldsi32(addr) + sq[r]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);
        if (++slot == kQStages) { slot = 0; phase ^= 1; }
      }

      // flush the interval: unpack the two 16-bit fields, RED.ADD.64 into the leaf's pool slot
      if (merge) {
        consumer_bar_sync();
        const int b_lo = pair * 86, b_hi = min(kBinsPerColumn, b_lo + 86);
#pragma unroll 2
        for (int b = b_lo; b < b_hi; ++b) {
          const int p0 = ldsi32(hbase0 + (b << 7));
          const int p1 = ldsi32(hbase0 + kQWarpSmemBytes + (b << 7));
          const int p2 = ldsi32(hbase0 + 2 * kQWarpSmemBytes + (b << 7));
          const int h0 = p0 & 0xffff, h1 = p1 & 0xffff, h2 = p2 & 0xffff;
          const int G = ((p0 - h0) >> 16) + ((p1 - h1) >> 16) + ((p2 - h2) >> 16);
          q_flush_cell(dst, b, G, h0 + h1 + h2);
        }
        consumer_bar_sync();
      } else {
#pragma unroll 4
        for (int b = 0; b < kBinsPerColumn; ++b) {
          const int p = ldsi32(hbase + (b << 7));
          const int h = p & 0xffff;
          q_flush_cell(dst, b, (p - h) >> 16, h);
        }
        __syncwarp();
      }
    }
  }
}

}  // namespace b200
