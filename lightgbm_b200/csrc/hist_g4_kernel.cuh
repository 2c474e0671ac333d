// hist_g4_kernel.cuh — EXPERIMENTAL (LGBMB200_Config.reserved bit 9, off by default, not yet run on hardware):
// k_hist with the gathered (non-root) stages moved by TMA tile::gather4 instead of per-lane cp.async.
//
// Why: after the L2 prefetch and the leaf-ordered (g,h) copy a gathered pass still costs ~50 % more per row than the
// root pass; the remaining difference between the two paths is HOW the 32-byte row segments reach
// shared memory: one TMA tile per stage at the root, 64 LDGSTS lane requests + 32 arrivals per stage when gathering.
// tile::gather4 (PTX ISA 8.6, sm_100+) takes one column coordinate and four ROW coordinates and lands the four rows
// as one contiguous [4][box_w] tile; the tensor map is encoded with box {32 columns, 1 row} (as CUTLASS does for
// SM100_TMA_LOAD_2D_GATHER4, cute/atom/copy_traits_sm90_tma.hpp).  Everything else (consumer, flush, work mapping) is
// k_hist's, copied verbatim so that the default kernel stays byte-identical while this one is being brought up.
#pragma once
#include "hist_kernel.cuh"

namespace b200 {

constexpr int kOobRow = 0x40000000;      // any row coordinate >= num_data: the TMA unit zero-fills that row

// tx bytes are registered separately from the arrival (the 32 arrivals of a gathered stage come from cp.async)
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(bytes) : "memory");
}
// four rows {r0..r3} x box_w columns starting at column `col` -> [4][box_w] in shared memory (UTMALDG gather4)
__device__ __forceinline__ void tma_gather4_2d(void* smem_dst, const CUtensorMap* tmap, int col, int r0, int r1, int r2, int r3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3),
                 "r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
}

template <bool QUANT>
__global__ void __launch_bounds__(kHistThreads, 1) k_hist_g4(const HistArgs a, const __grid_constant__ CUtensorMap tmap,
                                                                 const __grid_constant__ CUtensorMap tmap_g4) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_producer = warp >= kHistWarps;
  const int pair = is_producer ? warp - kHistWarps : warp;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kHistWarps * kWarpSmemBytes);
  uint64_t* full = bars + pair * (2 * kStages);       // producer -> consumer: stage landed
  uint64_t* empty = full + kStages;                   // consumer -> producer: stage consumed
  if (!is_producer && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 32); mbar_init(empty + i, 1); }
  }
  __syncthreads();
  pdl_enter();          // barrier init above overlaps the predecessor's tail; everything below reads its results

  HistWork w;
  if (!hist_work_setup(a, &w)) return;
  unsigned char* wbase = smem + pair * kWarpSmemBytes;
  unsigned char* ring = wbase + kWarpHistBytes;

  if (is_producer) {
    // ------------------------------------------------------------------ producer warp: stage rows
    int slot = 0; unsigned phase = 0;
    for (int round = 0; hist_round_valid(w, round); ++round) {
      const HistItem it = hist_item(w, round, pair);
      if (!it.valid) continue;
      const int cg = it.cg, part = it.part;
      const int r0 = part * w.per;
      const int r1 = min(w.n, r0 + w.per);
      if (r0 >= r1) continue;
      const uint8_t* colbase = a.bins + static_cast<int64_t>(cg) * kColGroup;
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
          // contiguous rows (root of an un-bagged tree): ONE 2-D TMA tile (32 rows x 32 columns of the row-major
          // matrix) + one bulk copy of the 32 (g,h) pairs per stage; both complete on the stage's mbarrier
          mbar_wait_parked(empty + slot, phase ^ 1);
          unsigned char* sb = ring + slot * kStageBytes;
          if (lane == 0) {
            mbar_arrive_expect_tx(full + slot, kStageBytes);
            tma_load_tile_2d(sb, &tmap, cg * kColGroup, p0, full + slot);
            tma_load_1d(sb + kStageBinBytes, a.gh + p0, kStageRows * 8, full + slot);
          } else {
            mbar_arrive(full + slot);
          }
          if (++slot == kStages) { slot = 0; phase ^= 1; }
          continue;
        }
        // Gathered stage through TMA: lane l < 8 issues ONE tile::gather4 (rows p0+4l .. p0+4l+3, 32 columns each ->
        // 128 contiguous bytes of the stage) instead of the 64 LDGSTS lane requests of k_hist; rows past the part's
        // end use the out-of-bounds coordinate num_rows (TMA zero-fills, the consumer never reads them).  The eight
        // instructions complete on the stage's mbarrier with complete_tx; the (g,h) pairs still travel by cp.async
        // whose completion provides the 32 arrivals.
        const int pg = p0 + lane;
        int rid = kOobRow;
        if (pg < r1) rid = ip ? __ldg(ip + pg) : pg;
        const int q0 = __shfl_sync(0xffffffffu, rid, (4 * lane) & 31), q1 = __shfl_sync(0xffffffffu, rid, (4 * lane + 1) & 31);
        const int q2 = __shfl_sync(0xffffffffu, rid, (4 * lane + 2) & 31), q3 = __shfl_sync(0xffffffffu, rid, (4 * lane + 3) & 31);
        const int rg = (pg < r1) ? (w.gh_ord != nullptr ? pg : rid) : -1;
        mbar_wait_parked(empty + slot, phase ^ 1);           // the consumer released this ring slot
        unsigned char* sb = ring + slot * kStageBytes;
        if (lane < 8) {
          mbar_expect_tx(full + slot, 4 * kColGroup);
          tma_gather4_2d(sb + lane * (4 * kColGroup), &tmap_g4, cg * kColGroup, q0, q1, q2, q3, full + slot);
        }
        if (rg >= 0) cp_async8(sb + kStageBinBytes + lane * 8, (w.gh_ord != nullptr ? w.gh_ord : a.gh) + rg);
        mbar_arrive_on_cp_async(full + slot);
        if (++slot == kStages) { slot = 0; phase ^= 1; }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warp: accumulate
  float2* H = reinterpret_cast<float2*>(wbase) + lane;            // lane's column of the [bin][lane] table
  const unsigned hbase = static_cast<unsigned>(__cvta_generic_to_shared(H));   // + bin*256 = the lane's cell
  const unsigned hbase0 = static_cast<unsigned>(__cvta_generic_to_shared(reinterpret_cast<float2*>(smem) + lane));
  const double gs = a.ctl->g_scale, hs = a.ctl->h_scale;
  int slot = 0; unsigned phase = 0;
  for (int round = 0; hist_round_valid(w, round); ++round) {
    const HistItem it = hist_item(w, round, pair);
    // merge path: all three warps of the CTA hold items of the same column group (CTA-uniform condition)
    const bool merge = it.merge;
    if (!it.valid) continue;                                       // only possible when !merge
    const int cg = it.cg, part = it.part;
    const int r0 = part * w.per;
    const int r1 = min(w.n, r0 + w.per);
    if (r0 >= r1 && !merge) continue;

    // zero the warp-private histogram (the producer is already filling the ring meanwhile)
    {
      float4* z = reinterpret_cast<float4*>(wbase);
#pragma unroll 8
      for (int i = lane; i < kWarpHistBytes / 16; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();

    for (int p0 = r0; p0 < r1; p0 += kStageRows) {
      mbar_wait(full + slot, phase);
      const int cnt = min(kStageRows, r1 - p0);
      const unsigned char* sb = ring + slot * kStageBytes;
      const unsigned char* sbin = sb + lane;
      const float2* sgh = reinterpret_cast<const float2*>(sb + kStageBinBytes);
      if (cnt == kStageRows) {
        constexpr int K = kHistBatch;
        // Software pipeline over the stage's 32/K batches (one warp per SMSP has no other warp to hide latency):
        //   LDS x K of batch k  |  [ALU: addresses + combined increments of batch k+1]  |  FADD + STS x K of batch k
        // and the raw bins / (g,h) of batch k+2 are fetched from the stage one iteration ahead.
        uint32_t nb[K]; float2 nq[K];            // raw inputs of the batch after next
        unsigned addrN[K]; float2 sN[K];         // prepared batch (next to be committed)
        auto fetch = [&](int r) {
#pragma unroll
          for (int i = 0; i < K; ++i) nb[i] = sbin[(r + i) * 32];
#pragma unroll
          for (int i = 0; i < K; i += 2) {
            const float4 t = *reinterpret_cast<const float4*>(sgh + r + i);
            nq[i] = make_float2(t.x, t.y); nq[i + 1] = make_float2(t.z, t.w);
          }
        };
        fetch(0);
        batch_prepare<K, QUANT>(hbase, nb, nq, addrN, sN);
        if (K < kStageRows) fetch(K);
#pragma unroll
        for (int r = 0; r < kStageRows; r += K) {
          unsigned addrC[K]; float2 sC[K], v[K];
#pragma unroll
          for (int i = 0; i < K; ++i) { addrC[i] = addrN[i]; sC[i] = sN[i]; }
#pragma unroll
          for (int i = 0; i < K; ++i) v[i] = lds64(addrC[i]);
          if (r + K < kStageRows) {
            batch_prepare<K, QUANT>(hbase, nb, nq, addrN, sN);     // batch r+K, in the shadow of the loads above
            if (r + 2 * K < kStageRows) fetch(r + 2 * K);
          }
#pragma unroll
          for (int i = 0; i < K; ++i) sts64(addrC[i], acc2<QUANT>(v[i], sC[i]));
        }
      } else {
        for (int r = 0; r < cnt; ++r) {
          const uint32_t b = sbin[r * 32];
          const float2 q = sgh[r];
          const unsigned addr = hbase + (b << 8);
          sts64(addr, acc2<QUANT>(lds64(addr), q));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + slot);
      if (++slot == kStages) { slot = 0; phase ^= 1; }
    }

    // flush: fp32 partials -> int64 fixed point, RED.ADD.64 into the leaf's pool slot.
    unsigned long long* dst = a.pool + static_cast<int64_t>(w.slot) * a.slot_stride +
                              (static_cast<int64_t>(cg) * kColGroup + lane) * (kBinsPerColumn * 2);
    if (merge) {
      // the CTA's three tables cover the same 32 columns: sum them in shared memory (fixed order 0,1,2 =>
      // deterministic) and let each warp flush a third of the bins => 3x fewer global atomics
      consumer_bar_sync();
      const int b_lo = pair * 86, b_hi = min(kBinsPerColumn, b_lo + 86);
#pragma unroll 2
      for (int b = b_lo; b < b_hi; ++b) {
        const float2 v0 = lds64(hbase0 + (b << 8));
        const float2 v1 = lds64(hbase0 + kWarpSmemBytes + (b << 8));
        const float2 v2 = lds64(hbase0 + 2 * kWarpSmemBytes + (b << 8));
        if (QUANT) {
          const int gi = __float_as_int(v0.x) + __float_as_int(v1.x) + __float_as_int(v2.x);
          const int hi = __float_as_int(v0.y) + __float_as_int(v1.y) + __float_as_int(v2.y);
          if (gi != 0 || hi != 0) {
            atomicAdd(dst + 2 * b, static_cast<unsigned long long>(static_cast<long long>(gi)));
            atomicAdd(dst + 2 * b + 1, static_cast<unsigned long long>(static_cast<long long>(hi)));
          }
        } else {
          const float gx = (v0.x + v1.x) + v2.x, hx = (v0.y + v1.y) + v2.y;
          if (gx != 0.f || hx != 0.f) {
            atomicAdd(dst + 2 * b, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(gx) * gs)));
            atomicAdd(dst + 2 * b + 1, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(hx) * hs)));
          }
        }
      }
      consumer_bar_sync();            // the tables are free to be zeroed for the next item
    } else {
#pragma unroll 4
      for (int b = 0; b < kBinsPerColumn; ++b) {
        const float2 v = lds64(hbase + (b << 8));
        if (QUANT) {
          const int gi = __float_as_int(v.x), hi = __float_as_int(v.y);
          if (gi != 0 || hi != 0) {
            atomicAdd(dst + 2 * b, static_cast<unsigned long long>(static_cast<long long>(gi)));
            atomicAdd(dst + 2 * b + 1, static_cast<unsigned long long>(static_cast<long long>(hi)));
          }
        } else if (v.x != 0.f || v.y != 0.f) {
          atomicAdd(dst + 2 * b, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(v.x) * gs)));
          atomicAdd(dst + 2 * b + 1, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(v.y) * hs)));
        }
      }
      __syncwarp();
    }
  }
}

}  // namespace b200
