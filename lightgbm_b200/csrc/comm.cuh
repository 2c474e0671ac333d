// comm.cuh — device-side primitives of the multi-GPU exchange over NVLink peer memory (no NCCL call on the
// per-split path).  Every rank maps every peer's CommBlock (CUDA IPC); a message is "plain stores into the
// peer's block, then st.release.sys of a sequence number by the SAME thread (the release orders its earlier stores at
// system scope: no separate fence, which would cost one more NVLink round trip)"; the receiver spins with
// ld.acquire.sys on its OWN memory and reads the payload around L1.  Messages are double-buffered by the parity
// of the sequence number: a rank can never run two exchanges ahead of a peer, because every exchange needs
// that peer's contribution.  A watchdog turns a missing peer into an error flag instead of a hang.
#pragma once
#include "types.cuh"

namespace b200 {

// Programmatic dependent launch (LGBMB200_Config.reserved bit 4): every kernel of the per-split chain first waits
// for its predecessor grid (a no-op when the launch carries no programmatic edge), then lets ITS successor's CTAs
// be scheduled, so that the next kernel's launch latency overlaps this kernel's execution.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr long long kWatchdogCycles = 20000000000ll;   // ~10 s: a missing peer becomes an error, not a hang

// spin until *p >= target (sequence numbers only grow); returns false on watchdog expiry
__device__ __forceinline__ bool wait_seq(const unsigned long long* p, unsigned long long target) {
  const long long t0 = clock64();
  while (ld_acquire_sys(p) < target) {
    if (clock64() - t0 > kWatchdogCycles) return false;
  }
  return true;
}

// Small all-gather of 8 doubles per rank, called by ALL threads of one block (blockDim.x >= world).
// in8: this rank's payload (shared or global memory, same for all threads); out: shared memory [kMaxRanks][8].
__device__ __forceinline__ void exchange_misc(const CommPeers& P, Ctl* c, const double* in8, double (*out)[8]) {
  const int tid = threadIdx.x, W = P.world, me = P.rank;
  const unsigned long long seq = c->misc_seq + 1;
  const int par = static_cast<int>(seq & 1);
  if (tid < W) {
    CommBlock* dst = P.block[tid];
#pragma unroll
    for (int k = 0; k < 8; ++k) dst->misc[par][me][k] = in8[k];
    st_release_sys(&dst->misc_seq[par][me], seq);
  }
  __syncthreads();
  if (tid < W) {
    CommBlock* mine = P.block[me];
    if (!wait_seq(&mine->misc_seq[par][tid], seq)) c->error = 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) out[tid][k] = __ldcv(&mine->misc[par][tid][k]);
  }
  __syncthreads();
  if (tid == 0) c->misc_seq = seq;
  __syncthreads();
}

}  // namespace b200
