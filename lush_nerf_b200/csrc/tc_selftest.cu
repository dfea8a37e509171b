// Hardware self-test of the UMMA conventions in tc_common.cuh (not on the product path):
// one CTA, D[128 x N] = A[128 x K] * B[N x K]^T with fp16 operands held in tile-canonical blobs,
// in either operand-major mode.  Called from tests/test_tc_selftest_gpu.py through the C ABI.
#include "tc_common.cuh"

namespace lush {
using namespace tc;

// a_mode / b_mode: 0 = K-major view, 1 = MN-major view of the blob; a_mode = 2: A is first moved to TENSOR MEMORY
// (tcgen05.st, two fp16 per 32-bit column) and the MMA reads it from there (the fused MLP's activation path)
// K-major A: blob rows = 128 (M), cols = K.   MN-major A: blob rows = K, cols = 128 (M).
// K-major B: blob rows = N, cols = K.         MN-major B: blob rows = K, cols = N.
__global__ void __launch_bounds__(128) tc_selftest_kernel(const __half* __restrict__ a_blob,
                                                          const __half* __restrict__ b_blob,
                                                          float* __restrict__ d_out, int N, int K, int a_mode,
                                                          int b_mode, int swap_lbo_sbo) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t a_bytes = 128u * K * 2u, b_bytes = (uint32_t)N * K * 2u;
  uint8_t* sa = smem;
  uint8_t* sb = smem + a_bytes;
  for (uint32_t i = tid * 16; i < a_bytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(a_blob) + i);
  for (uint32_t i = tid * 16; i < b_bytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(b_blob) + i);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (a_mode == 2) {   // row tid of A: K fp16 -> K/2 columns of lane tid, starting at column 256
    for (int k0 = 0; k0 < K; k0 += 32) {
      uint32_t pk[16];
      for (int g = 0; g < 4; ++g) {
        const uint4 q = *reinterpret_cast<const uint4*>(sa + blob_off(128, tid, k0 + 8 * g));
        pk[4 * g] = q.x; pk[4 * g + 1] = q.y; pk[4 * g + 2] = q.z; pk[4 * g + 3] = q.w;
      }
      tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + 256 + k0 / 2, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (tid == 0 && a_mode == 2) {
    const uint32_t idesc = make_idesc_f16(128, N, 0, b_mode);
    const uint32_t b_rows = b_mode == 0 ? (uint32_t)N : (uint32_t)K;
    const uint32_t b_lbo = b_mode == 0 ? b_rows * 16u : 128u, b_sbo = b_mode == 0 ? 128u : b_rows * 16u;
    const uint32_t b_step = b_mode == 0 ? 2u * b_rows * 16u : 256u;
    for (int k = 0; k < K / 16; ++k) {
      uint64_t bd = make_smem_desc(smem_u32(sb) + k * b_step, b_lbo, b_sbo);
      umma_f16_ts(tmem_base, tmem_base + 256 + k * 8, bd, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  } else if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N, a_mode, b_mode);
    // per K=16 step: K-major advances two 8-column groups (2*rows*16 B); MN-major advances 16 rows (256 B)
    const uint32_t a_rows = a_mode == 0 ? 128u : (uint32_t)K;
    const uint32_t b_rows = b_mode == 0 ? (uint32_t)N : (uint32_t)K;
    uint32_t a_lbo = a_mode == 0 ? a_rows * 16u : 128u, a_sbo = a_mode == 0 ? 128u : a_rows * 16u;
    uint32_t b_lbo = b_mode == 0 ? b_rows * 16u : 128u, b_sbo = b_mode == 0 ? 128u : b_rows * 16u;
    if (swap_lbo_sbo) {
      uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t;
      t = b_lbo; b_lbo = b_sbo; b_sbo = t;
    }
    const uint32_t a_step = a_mode == 0 ? 2u * a_rows * 16u : 256u;
    const uint32_t b_step = b_mode == 0 ? 2u * b_rows * 16u : 256u;
    for (int k = 0; k < K / 16; ++k) {
      uint64_t ad = make_smem_desc(smem_u32(sa) + k * a_step, a_lbo, a_sbo);
      uint64_t bd = make_smem_desc(smem_u32(sb) + k * b_step, b_lbo, b_sbo);
      umma_f16(tmem_base, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = tid;  // TMEM lane = 32*(warp%4) + lane = tid for a 128-thread block
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) d_out[row * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}


// ---- tensor-pipe micro-benchmark (tools/bench_umma.py): one CTA, one issuing thread, `iters` back-to-back
// tcgen05.mma (M=128, K=16) with zeroed operands; returns elapsed SM cycles from first issue to completion.
//   N        : MMA N (64 | 128 | 256)
//   nacc     : number of accumulator regions used round-robin (1 = every MMA depends on the previous one)
//   a_tmem   : 1 = A operand from tensor memory (TS mode), 0 = from shared memory (SS mode)
__device__ __forceinline__ uint32_t desc_lo_(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__global__ void __launch_bounds__(512) tc_microbench_kernel(long long* __restrict__ out, int N, int nacc, int a_tmem,
                                                            int iters, int flags) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_flag;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) stop_flag = 0;
  for (uint32_t i = tid * 16; i < 4096u + 256u * 32u; i += 512 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int issuer_warp = (flags & 16) ? 15 : 0;
  if (warp != issuer_warp && (flags & 8)) {   // busy neighbours: dependent FMA chains + shared-memory loads
    float a = (float)tid, b = 1.0001f;
    while (!stop_flag) {
#pragma unroll
      for (int k = 0; k < 64; ++k) a = fmaf(a, b, reinterpret_cast<float*>(smem)[(tid + k) & 1023]);
    }
    if (a == 12345.f) out[3] = 1;
  }
  if (warp == issuer_warp) {   // converged warp, one elected lane issues blocks of 8 MMAs (as the product kernels do)
    const uint32_t idesc = make_idesc_f16(128, N, 0, 0);
    const uint32_t a_lo = desc_lo_(smem_u32(smem), 128u * 16u), b_lo = desc_lo_(smem_u32(smem) + 4096u, (uint32_t)N * 16u);
    const uint32_t hi = (128u >> 4) | (1u << 14);
    const uint32_t d0 = tmem_base, d1 = tmem_base + (nacc > 1 ? (uint32_t)N : 0u);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
      if (flags & 4) mbar_wait(&bar3, 1);      // already-complete phase: pure wait latency
      if (flags & 2) tc_fence_after();
      uint32_t e;
      asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(e));
      if (e) {
        if (a_tmem) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 db, {%2, %3};\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"((k & 1) ? d1 : d0),
                         "r"(tmem_base + 448u), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 da, {%1, %3};\n\t"
                         "mov.b64 db, {%2, %3};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(
                             (k & 1) ? d1 : d0),
                         "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
        }
        if (flags & 1) umma_commit(&bar2);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if ((tid & 31) == 0) umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if ((tid & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    stop_flag = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace lush

extern "C" int lush_tc_selftest(const void* a_blob, const void* b_blob, float* d_out, int N, int K, int a_mode,
                                int b_mode, int swap_lbo_sbo, void* stream) {
  using namespace lush;
  LUSH_REQUIRE(a_blob && b_blob && d_out, "tc_selftest: null pointer");
  LUSH_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K >= 16 && K % 16 == 0, "tc_selftest: bad N/K");
  size_t smem = (size_t)(128 + N) * K * 2;
  LUSH_REQUIRE(smem <= 200 * 1024, "tc_selftest: operands do not fit in shared memory");
  int rc = check_cuda(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem), "tc_selftest: smem attr");
  if (rc) return rc;
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __half*)a_blob, (const __half*)b_blob, d_out,
                                                             N, K, a_mode, b_mode, swap_lbo_sbo);
  LUSH_LAUNCH_CHECK("tc_selftest");
  return 0;
}

extern "C" int lush_tc_microbench(long long* out_cycles, int N, int nacc, int a_tmem, int iters, int flags, void* stream) {
  using namespace lush;
  LUSH_REQUIRE(out_cycles && (N == 64 || N == 128 || N == 256) && nacc >= 1 && nacc <= 2 && nacc * N <= 384 && iters > 0 && iters % 8 == 0,
               "tc_microbench: bad arguments");
  const size_t smem = 4096 + 256 * 32;
  tc_microbench_kernel<<<1, 512, smem, (cudaStream_t)stream>>>(out_cycles, N, nacc, a_tmem, iters, flags);
  LUSH_LAUNCH_CHECK("tc_microbench");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Microbenchmark: every CTA streams the SAME `nblocks` x `block_bytes` weight image from L2 into a 3-stage
// shared-memory ring, `iters` times (the access pattern of the chain kernels' weight producer), either with
// one unicast bulk copy per CTA and block, or -- clusters of CL CTAs -- with each CTA fetching 1/CL of a block
// and multicasting it to the whole cluster.  out[cta] = cycles; tells whether multicast raises the delivered
// weight bandwidth per SM.  Not on the product path.
// ------------------------------------------------------------------------------------------------
namespace lush {
using namespace tc;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                            uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}

template <int CL>
__global__ void __launch_bounds__(128) bcast_bench_kernel(const uint8_t* __restrict__ src, int nblocks, int block_bytes,
                                                          int iters, long long* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[3], empty[3];
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 3; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }
    fence_barrier_init();
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    // producer and consumer in one thread, software-pipelined 3 deep: issue block i, consume block i-2
    const int total = nblocks * iters;
    const uint32_t part = (uint32_t)block_bytes / CL;
    int ps = 0, cs = 0; uint32_t pphase = 0, cphase = 0;
    for (int i = 0; i < total + 2; ++i) {
      if (i < total) {
        if (i >= 3) mbar_wait(&empty[ps], pphase ^ 1);
        mbar_arrive_expect_tx(&full[ps], (uint32_t)block_bytes);
        const uint8_t* g = src + (size_t)(i % nblocks) * block_bytes + rank * part;
        uint8_t* d = smem + ps * block_bytes + rank * part;
        if (CL > 1) bulk_g2s_mc(d, g, part, &full[ps], (uint16_t)((1u << CL) - 1u));
        else bulk_g2s(d, g, part, &full[ps]);
        if (++ps == 3) { ps = 0; pphase ^= 1; }
      }
      if (i >= 2) {
        mbar_wait(&full[cs], cphase);
        if (CL > 1) { for (uint32_t c = 0; c < CL; ++c) mbar_arrive_remote(&empty[cs], c); }
        else mbar_arrive(&empty[cs]);
        if (++cs == 3) { cs = 0; cphase ^= 1; }
      }
    }
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

}  // namespace lush

extern "C" int lush_tc_bcast_bench(const void* src, int nblocks, int block_bytes, int iters, int cluster, int grid,
                                   long long* out_cycles, void* stream) {
  using namespace lush;
  LUSH_REQUIRE(src && out_cycles && nblocks > 0 && iters > 0 && block_bytes > 0 && block_bytes % 64 == 0 &&
                   3 * block_bytes <= 200 * 1024 && (cluster == 1 || cluster == 2 || cluster == 4) && grid > 0 &&
                   grid % cluster == 0,
               "tc_bcast_bench: bad arguments");
  void (*kern)(const uint8_t*, int, int, int, long long*) =
      cluster == 1 ? bcast_bench_kernel<1> : cluster == 2 ? bcast_bench_kernel<2> : bcast_bench_kernel<4>;
  const size_t smem = (size_t)3 * block_bytes;
  LUSH_REQUIRE(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess,
               "tc_bcast_bench: cannot raise the shared memory limit");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const uint8_t* s8 = (const uint8_t*)src;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, s8, nblocks, block_bytes, iters, out_cycles);
  LUSH_REQUIRE(e == cudaSuccess, "tc_bcast_bench: launch failed: %s", cudaGetErrorString(e));
  LUSH_LAUNCH_CHECK("tc_bcast_bench");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// CTA-pair self-test (cta_group::2): D[256 x N] = A[256 x K] * B[N x K]^T with a cluster of two CTAs.  CTA r stages
// rows [128 r, 128 r + 128) of A and rows [N/2 r, N/2 r + N/2) of B (both K-major canonical blobs) in ITS shared
// memory; the leader issues one M=256 MMA per K=16 that reads both CTAs' operands, and each CTA finds its 128 rows of D
// (all N columns) in its own tensor memory.  This is the mode that halves the weight bytes an SM has to stage and read
// (shared-memory traffic); not on the product path.
// ------------------------------------------------------------------------------------------------
namespace lush {
using namespace tc;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
tc_selftest_pair_kernel(const __half* __restrict__ a_blob, const __half* __restrict__ b_blob, float* __restrict__ d_out,
                        int N, int K) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const uint32_t a_bytes = 128u * K * 2u, b_bytes = (uint32_t)(N / 2) * K * 2u;
  uint8_t* sa = smem;
  uint8_t* sb = smem + a_bytes;
  const uint8_t* ga = reinterpret_cast<const uint8_t*>(a_blob) + (size_t)rank * a_bytes;
  const uint8_t* gb = reinterpret_cast<const uint8_t*>(b_blob) + (size_t)rank * b_bytes;
  for (uint32_t i = tid * 16; i < a_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(ga + i);
  for (uint32_t i = tid * 16; i < b_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(gb + i);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {   // the same warp id in both CTAs, the same shared-memory address for the result
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' operands and barriers are in place before the leader issues
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = make_idesc_f16(256, N, 0, 0);
    const uint32_t a_lbo = 128u * 16u, b_lbo = (uint32_t)(N / 2) * 16u;
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t ad = make_smem_desc(smem_u32(sa) + k * 2u * a_lbo, a_lbo, 128u);
      const uint64_t bd = make_smem_desc(smem_u32(sb) + k * 2u * b_lbo, b_lbo, 128u);
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_base),
          "l"(ad), "l"(bd), "r"(idesc), "r"(k > 0 ? 1u : 0u)
          : "memory");
    }
    // completion is signalled on `bar` of BOTH CTAs (same shared-memory offset)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(&bar)),
                 "h"((uint16_t)3)
                 : "memory");
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = tid;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) d_out[(size_t)(128 * rank + row) * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // neither CTA frees tensor memory while its peer may still be reading
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

}  // namespace lush

extern "C" int lush_tc_selftest_pair(const void* a_blob, const void* b_blob, float* d_out, int N, int K, void* stream) {
  using namespace lush;
  LUSH_REQUIRE(a_blob && b_blob && d_out && N >= 32 && N <= 256 && N % 32 == 0 && K >= 16 && K % 16 == 0,
               "tc_selftest_pair: bad arguments");
  const size_t smem = (size_t)128 * K * 2 + (size_t)(N / 2) * K * 2;
  LUSH_REQUIRE(smem <= 200 * 1024, "tc_selftest_pair: operands do not fit shared memory");
  LUSH_REQUIRE(cudaFuncSetAttribute(tc_selftest_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) ==
                   cudaSuccess,
               "tc_selftest_pair: cannot raise the shared memory limit");
  tc_selftest_pair_kernel<<<2, 128, smem, (cudaStream_t)stream>>>((const __half*)a_blob, (const __half*)b_blob, d_out, N, K);
  LUSH_LAUNCH_CHECK("tc_selftest_pair");
  return 0;
}
