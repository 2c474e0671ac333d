// hist_mb.cu — micro-benchmark of candidate inner loops for the histogram consumer (sm_100a).
// Measures warp-rows (32 cells) per clock per SM for shared-memory scatter-add variants, with the staged rows
// already resident in shared memory (no global traffic in the timed loop): what bounds k_hist once staging is
// hidden.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hist_mb hist_mb.cu ; run on one B200.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kStageRows = 128;                  // rows of synthetic staged data, reused round-robin
constexpr int kStageBytes = kStageRows * 32;     // [row][32 columns] bytes

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ float2 lds64(unsigned a) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(unsigned a, float2 v) { asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(v.x), "f"(v.y)); }
__device__ __forceinline__ float lds32f(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32f(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ void red32(unsigned a, int v) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// smem layout: [tables][stage bins: kStageBytes][stage gh: kStageRows*8]
// ---------------------------------------------------------------------------------------------------------
// P64: warp-private [bin][lane] float2 table (64 KB / warp), LDS.64 + FADD x2 + STS.64, batches of 4 rows with
// in-register duplicate forwarding (the round-1 kernel's inner loop).
template <int TRANSPOSED>
__global__ void __launch_bounds__(128, 1) mb_p64(const uint8_t* gbins, const float2* ggh, int rows_per_warp, float* sink, int nwarps) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* stage = smem + nwarps * 65536;
  float2* sgh = reinterpret_cast<float2*>(stage + kStageBytes);
  for (int i = threadIdx.x; i < kStageBytes; i += blockDim.x) stage[i] = gbins[i];
  for (int i = threadIdx.x; i < kStageRows; i += blockDim.x) sgh[i] = ggh[i];
  for (int i = threadIdx.x; i < nwarps * 65536 / 16; i += blockDim.x) reinterpret_cast<float4*>(smem)[i] = make_float4(0, 0, 0, 0);
  __syncthreads();
  if (warp >= nwarps) return;
  const unsigned hbase = smem_u32(smem + warp * 65536) + lane * 8;
  for (int r = 0; r < rows_per_warp; r += 4) {
    const int rr = r & (kStageRows - 1);
    uint32_t b[4];
    if (TRANSPOSED) {
      // stage laid out [row/4][col][4 rows]: one LDS.32 gives the lane's column for 4 rows
      const uint32_t w = *reinterpret_cast<const uint32_t*>(stage + (rr >> 2) * 128 + lane * 4);
      b[0] = w & 255; b[1] = (w >> 8) & 255; b[2] = (w >> 16) & 255; b[3] = w >> 24;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = stage[(rr + i) * 32 + lane];
    }
    float2 q[4];
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
      const float4 t = *reinterpret_cast<const float4*>(sgh + rr + i);
      q[i] = make_float2(t.x, t.y); q[i + 1] = make_float2(t.z, t.w);
    }
    unsigned addr[4]; float2 s[4], v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      addr[i] = hbase + (b[i] << 8);
      s[i] = q[i];
#pragma unroll
      for (int j = 0; j < i; ++j) if (b[j] == b[i]) { s[i].x += q[j].x; s[i].y += q[j].y; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = lds64(addr[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) sts64(addr[i], make_float2(v[i].x + s[i].x, v[i].y + s[i].y));
  }
  __syncwarp();
  float acc = 0.f;
  for (int b = 0; b < 256; ++b) { const float2 v = lds64(hbase + (b << 8)); acc += v.x + v.y; }
  if (acc == 123.456f) sink[0] = acc;
}

// P32: warp-private [bin][lane] fp32 table (32 KB / warp), LDS.32 + FADD + STS.32 (one component; the regime of a
// 32-bit cell: constant hessian gradient-only, or the packed quantized cell)
template <int TRANSPOSED>
__global__ void __launch_bounds__(256, 1) mb_p32(const uint8_t* gbins, const float2* ggh, int rows_per_warp, float* sink, int nwarps) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* stage = smem + nwarps * 32768;
  float2* sgh = reinterpret_cast<float2*>(stage + kStageBytes);
  for (int i = threadIdx.x; i < kStageBytes; i += blockDim.x) stage[i] = gbins[i];
  for (int i = threadIdx.x; i < kStageRows; i += blockDim.x) sgh[i] = ggh[i];
  for (int i = threadIdx.x; i < nwarps * 32768 / 16; i += blockDim.x) reinterpret_cast<float4*>(smem)[i] = make_float4(0, 0, 0, 0);
  __syncthreads();
  if (warp >= nwarps) return;
  const unsigned hbase = smem_u32(smem + warp * 32768) + lane * 4;
  const float* sg = reinterpret_cast<const float*>(sgh);
  for (int r = 0; r < rows_per_warp; r += 4) {
    const int rr = r & (kStageRows - 1);
    uint32_t b[4];
    if (TRANSPOSED) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(stage + (rr >> 2) * 128 + lane * 4);
      b[0] = w & 255; b[1] = (w >> 8) & 255; b[2] = (w >> 16) & 255; b[3] = w >> 24;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = stage[(rr + i) * 32 + lane];
    }
    const float4 t = *reinterpret_cast<const float4*>(sg + rr);
    const float q[4] = {t.x, t.y, t.z, t.w};
    unsigned addr[4]; float s[4], v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      addr[i] = hbase + (b[i] << 7);
      s[i] = q[i];
#pragma unroll
      for (int j = 0; j < i; ++j) if (b[j] == b[i]) s[i] += q[j];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = lds32f(addr[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) sts32f(addr[i], v[i] + s[i]);
  }
  __syncwarp();
  float acc = 0.f;
  for (int b = 0; b < 256; ++b) acc += lds32f(hbase + (b << 7));
  if (acc == 123.456f) sink[0] = acc;
}

// A: CTA-shared int32 tables, native ATOMS.ADD (RED).  NCOMP = 2: gradient + hessian tables ([bin][lane] each, 32 KB),
// NCOMP = 1: one table.  All warps of the CTA add into the same tables (any number of warps).
template <int NCOMP, int TRANSPOSED>
__global__ void __launch_bounds__(1024, 1) mb_atoms(const uint8_t* gbins, const float2* ggh, int rows_per_warp, float* sink, int ntables) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tbytes = NCOMP * 32768;
  unsigned char* stage = smem + ntables * tbytes;
  int2* sgh = reinterpret_cast<int2*>(stage + kStageBytes);
  for (int i = threadIdx.x; i < kStageBytes; i += blockDim.x) stage[i] = gbins[i];
  for (int i = threadIdx.x; i < kStageRows; i += blockDim.x) sgh[i] = make_int2(static_cast<int>(ggh[i].x * 1000.f), static_cast<int>(ggh[i].y * 1000.f));
  for (int i = threadIdx.x; i < ntables * tbytes / 16; i += blockDim.x) reinterpret_cast<float4*>(smem)[i] = make_float4(0, 0, 0, 0);
  __syncthreads();
  const unsigned gbase = smem_u32(smem + (warp % ntables) * tbytes) + lane * 4;
  for (int r = 0; r < rows_per_warp; r += 4) {
    const int rr = (r + warp * 4) & (kStageRows - 1);
    uint32_t b[4];
    if (TRANSPOSED) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(stage + (rr >> 2) * 128 + lane * 4);
      b[0] = w & 255; b[1] = (w >> 8) & 255; b[2] = (w >> 16) & 255; b[3] = w >> 24;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = stage[(rr + i) * 32 + lane];
    }
    int2 q[4];
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
      const int4 t = *reinterpret_cast<const int4*>(sgh + rr + i);
      q[i] = make_int2(t.x, t.y); q[i + 1] = make_int2(t.z, t.w);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned a = gbase + (b[i] << 7);
      red32(a, q[i].x);
      if (NCOMP == 2) red32(a + 32768, q[i].y);
    }
  }
  __syncthreads();
  int acc = 0;
  for (int b = 0; b < 256; ++b) acc += *reinterpret_cast<int*>(smem + (warp % ntables) * tbytes + b * 128 + lane * 4);
  if (acc == 123456789) sink[0] = static_cast<float>(acc);
}

// A16: packed-pair variant for a constant hessian: gradient int32 table + a count table of 16-bit fields
// (two bins per word) — still one ATOMS per component; measures whether smaller count table matters (it should not).

template <typename F>
static double time_kernel(F launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms / reps;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("device %s, %d SMs, %d MHz\n", prop.name, sms, clk_khz / 1000);
  std::vector<uint8_t> hb(kStageBytes); std::vector<float2> hg(kStageRows);
  srand(1);
  for (auto& x : hb) x = rand() % 255;
  for (auto& g : hg) g = make_float2((rand() % 2000 - 1000) / 1000.f, 1.f);
  uint8_t* dbins; float2* dgh; float* sink;
  CK(cudaMalloc(&dbins, kStageBytes)); CK(cudaMalloc(&dgh, kStageRows * 8)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemcpy(dbins, hb.data(), kStageBytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dgh, hg.data(), kStageRows * 8, cudaMemcpyHostToDevice));
  const int stage_total = kStageBytes + kStageRows * 8;
  const int total_rows_per_sm = 1 << 20;        // warp-rows per SM per launch
  auto report = [&](const char* name, int nwarps, double ms) {
    const double cyc = ms * 1e-3 * clk_khz * 1e3;
    printf("%-34s warps/SM %2d  %8.3f ms  %6.3f clk/warp-row/SM  %6.2f cells/clk/SM  -> %5.1f %% of 6567 GB/s at 1 B/cell\n", name, nwarps, ms,
           cyc / total_rows_per_sm, 32.0 * total_rows_per_sm / cyc, 100.0 * (32.0 * total_rows_per_sm / (ms * 1e-3) * sms) / 6567.4e9);
  };
#define RUN_P64(T, NW) { CK(cudaFuncSetAttribute(mb_p64<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, NW * 65536 + stage_total)); \
    const int rpw = total_rows_per_sm / NW; \
    double ms = time_kernel([&] { mb_p64<T><<<sms, 128, NW * 65536 + stage_total>>>(dbins, dgh, rpw, sink, NW); }, 3); \
    report(T ? "P64 lds64/sts64 private, transposed" : "P64 lds64/sts64 private, u8 bins", NW, ms); }
  RUN_P64(0, 3) RUN_P64(1, 3)
#define RUN_P32(T, NW) { CK(cudaFuncSetAttribute(mb_p32<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, NW * 32768 + stage_total)); \
    const int rpw = total_rows_per_sm / NW; \
    double ms = time_kernel([&] { mb_p32<T><<<sms, 256, NW * 32768 + stage_total>>>(dbins, dgh, rpw, sink, NW); }, 3); \
    report(T ? "P32 lds32/sts32 private, transposed" : "P32 lds32/sts32 private, u8 bins", NW, ms); }
  RUN_P32(0, 3) RUN_P32(0, 4) RUN_P32(0, 6) RUN_P32(1, 4) RUN_P32(1, 6)
#define RUN_A(NC, T, NW, NT) { CK(cudaFuncSetAttribute(mb_atoms<NC, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, NT * NC * 32768 + stage_total)); \
    const int rpw = total_rows_per_sm / NW; \
    double ms = time_kernel([&] { mb_atoms<NC, T><<<sms, NW * 32, NT * NC * 32768 + stage_total>>>(dbins, dgh, rpw, sink, NT); }, 3); \
    char nm[96]; snprintf(nm, sizeof nm, "A%d atoms.add %s, %d table(s)", NC, T ? "transposed" : "u8 bins", NT); report(nm, NW, ms); }
  RUN_A(2, 0, 4, 1) RUN_A(2, 0, 8, 1) RUN_A(2, 0, 16, 1) RUN_A(2, 0, 32, 1)
  RUN_A(2, 1, 8, 1) RUN_A(2, 1, 16, 1) RUN_A(2, 1, 32, 1)
  RUN_A(2, 1, 12, 3) RUN_A(2, 1, 24, 3)
  RUN_A(1, 0, 8, 1) RUN_A(1, 0, 16, 1) RUN_A(1, 1, 8, 1) RUN_A(1, 1, 16, 1) RUN_A(1, 1, 32, 1)
  RUN_A(1, 1, 24, 6)
  return 0;
}
