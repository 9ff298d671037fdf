// Self-test of the tcgen05 pipeline: Y[128][N] = X[128][K] * W[K][N] (+ a second accumulate pass), used by
// tests/test_gpu_tc.py to pin descriptors / layouts / barrier protocol against torch before any model kernel
// depends on them.  Included by mdb_forward.cu inside its anonymous namespace.
#pragma once
#include "tc_pipe.cuh"

template <int K, int N, bool XF = false>
__global__ void __launch_bounds__(tc::NTHREADS_TC, 1)
tc_selftest_kernel(const float* __restrict__ X, const uint8_t* __restrict__ w_img, float* __restrict__ Y, int twice) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* a_hi = smem_raw;                                   // 128 * K * 2 bytes
  uint8_t* a_lo = a_hi + tc::ROWS * K * 2;
  uint8_t* stages = a_lo + tc::ROWS * K * 2;
  tc::PipeSmem* ps = reinterpret_cast<tc::PipeSmem*>(stages + tc::NSTAGE * tc::STAGE_SLOT);
  const int tid = threadIdx.x, warp = tid >> 5;
  tc::PipeT<tc::NSTAGE, XF> p;
  tc::pipe_init(p, ps, stages);
  if (warp == 4) tc::tmem_alloc<256>(&ps->tmem_base);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (p.role == 0) {
    for (int k0 = 0; k0 < K; k0 += 16) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = X[(size_t)tid * K + k0 + i];
      tc::store_a<K, 16>(a_hi, a_lo, tid, k0, v);
    }
    tc::rows_publish(p);
  }
  tc::gemm<K, N>(p, a_hi, a_lo, w_img, 0, false, true, !twice);
  if (twice) tc::gemm<K, N>(p, a_hi, a_lo, w_img, 0, true, false, true);   // accumulate a second pass: Y = 2 X W
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    const uint32_t taddr = ps->tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tc::tmem_ld32(taddr + c0, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) Y[(size_t)tid * N + c0 + i] = v[i];
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == 4) { __syncwarp(); tc::tmem_dealloc<256>(ps->tmem_base); }
}

template <int K, int N, bool XF = false>
int launch_tc_selftest(const float* X, const void* w_img, float* Y, int twice, cudaStream_t st) {
  const size_t smem = 2 * (size_t)tc::ROWS * K * 2 + tc::NSTAGE * tc::STAGE_SLOT + sizeof(tc::PipeSmem) + 128;
  CUDA_TRY(cudaFuncSetAttribute(tc_selftest_kernel<K, N, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_kernel<K, N, XF><<<1, tc::NTHREADS_TC, smem, st>>>(X, reinterpret_cast<const uint8_t*>(w_img), Y, twice);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return MDB_OK;
}
