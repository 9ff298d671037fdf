// Warp-specialised GEMM pipeline on top of tc_engine.cuh, shared by every tensor-core kernel of the path.
//
// CTA = NRW row warps + 2: the row warps (4 or 8) build the A operand and run every epilogue from TMEM (thread t
// of warp w owns tile row 32 * (w % 4) + t = its TMEM lane; with 8 row warps, warps 4-7 take the upper half of
// the columns of the same rows), lane 0 of the next warp streams weight stages L2 -> smem with cp.async.bulk,
// lane 0 of the last warp issues tcgen05.mma.  All three roles walk the same static sequence of GEMMs; they meet only on
// mbarriers:
//     a_ready (128 arrivals)  rows -> MMA   "A planes written, previous accumulator drained"
//     full[s] / empty[s]      producer <-> MMA, one weight stage each (empty is signalled by tcgen05.commit)
//     done                    MMA -> rows   "accumulator complete"
// The producer runs ahead across GEMM boundaries, so weight fetch overlaps the row threads' epilogues.
#pragma once
#include "tc_engine.cuh"

namespace tc {

constexpr int ROWS = 128;              // tile rows = UMMA M
constexpr int NTHREADS_TC = 192;       // 4 row warps + producer warp + MMA warp
constexpr int KB = 16;                 // K columns per weight stage
constexpr int NSTAGE = 4;
constexpr uint32_t STAGE_SLOT = 256 * KB * 2 * 2;   // bytes reserved per stage (N = 256 worst case): 32 KB

template <int NS>
struct PipeSmemT {                     // lives in shared memory
  uint64_t full[NS], empty[NS], done, a_ready;
  uint32_t tmem_base;
};
using PipeSmem = PipeSmemT<NSTAGE>;

template <int NS>
struct PipeT {
  PipeSmemT<NS>* s;
  uint8_t* stages;      // NSTAGE * STAGE_SLOT, 128-byte aligned
  uint32_t it;          // running stage counter (producer and MMA thread each advance their own copy)
  uint32_t n_done;      // GEMMs completed (MMA thread / row threads)
  uint32_t n_ready;     // a_ready phases consumed (MMA thread)
  int role;             // 0 = row thread, 1 = producer thread, 2 = MMA thread, 3 = idle lane
};
using Pipe = PipeT<NSTAGE>;

template <int NRW = 4, int NS = NSTAGE>   // NRW row warps: 4 = one thread per row, 8 / 16 = two / four threads per row
__device__ __forceinline__ void pipe_init(PipeT<NS>& p, PipeSmemT<NS>* s, uint8_t* stages) {
  p.s = s; p.stages = stages; p.it = 0; p.n_done = 0; p.n_ready = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  p.role = warp < NRW ? 0 : (lane == 0 ? (warp == NRW ? 1 : 2) : 3);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&s->full[i], 1); mbar_init(&s->empty[i], 1); }
    mbar_init(&s->done, 1);
    mbar_init(&s->a_ready, NRW * 32);
    fence_barrier_init();
  }
}

// Row threads: "my part of the A planes is written and I no longer read the accumulator".
template <int NS>
__device__ __forceinline__ void rows_publish(PipeT<NS>& p) {
  fence_proxy_async();
  fence_before_sync();
  mbar_arrive(&p.s->a_ready);
}
// Row threads: wait for the accumulator of the next GEMM in program order.
template <int NS>
__device__ __forceinline__ void rows_wait_acc(PipeT<NS>& p) {
  mbar_wait(&p.s->done, p.n_done & 1);
  ++p.n_done;
  fence_after_sync();
}

// One GEMM: D[tmem col d_col .. +N) (+)= A(a_hi, a_lo planes, K columns) * W (image at w_img).
// Called by ALL threads of the CTA at the same program point; row threads return immediately.
// wait_ready = false chains a second GEMM onto the same A planes / accumulator epoch without a new a_ready phase.
template <int K, int N, int NSLOT>
__device__ __forceinline__ void gemm(PipeT<NSLOT>& p, const uint8_t* a_hi, const uint8_t* a_lo, const uint8_t* w_img,
                                     uint32_t d_col, bool accumulate, bool wait_ready, bool signal_done) {
  using WS = WStage<N, KB>;
  constexpr int NS = K / KB;
  static_assert(K % KB == 0, "K must be a multiple of the stage depth");
  if (p.role == 1) {
    for (int s = 0; s < NS; ++s, ++p.it) {
      const uint32_t slot = p.it % NSLOT, ph = (p.it / NSLOT) & 1;
      mbar_wait(&p.s->empty[slot], ph ^ 1);
      mbar_arrive_expect_tx(&p.s->full[slot], WS::STAGE_BYTES);
      bulk_g2s(p.stages + slot * STAGE_SLOT, w_img + (size_t)s * WS::STAGE_BYTES, WS::STAGE_BYTES, &p.s->full[slot]);
    }
  } else if (p.role == 2) {
    if (wait_ready) {
      mbar_wait(&p.s->a_ready, p.n_ready & 1);
      ++p.n_ready;
      fence_after_sync();
    }
    constexpr uint32_t idesc = make_idesc_f16(ROWS, N, OPERAND_FMT, OPERAND_FMT);
    constexpr uint32_t SBO_A = (K / 8) * 128;
    const uint32_t d_tmem = p.s->tmem_base + d_col;
    const uint32_t ahi = smem_u32(a_hi), alo = smem_u32(a_lo);
    for (int s = 0; s < NS; ++s, ++p.it) {
      const uint32_t slot = p.it % NSLOT, ph = (p.it / NSLOT) & 1;
      mbar_wait(&p.s->full[slot], ph);
      fence_after_sync();
      const uint32_t bhi = smem_u32(p.stages + slot * STAGE_SLOT), blo = bhi + WS::PLANE_BYTES;
#pragma unroll
      for (int j = 0; j < KB / 16; ++j) {
        const uint32_t ks = s * (KB / 16) + j;
        const uint64_t da_hi = make_smem_desc(ahi + ks * 256, 128, SBO_A);
        const uint64_t da_lo = make_smem_desc(alo + ks * 256, 128, SBO_A);
        const uint64_t db_hi = make_smem_desc(bhi + j * 256, 128, WS::SBO);
        const uint64_t db_lo = make_smem_desc(blo + j * 256, 128, WS::SBO);
        mma_bf16_ss(d_tmem, da_hi, db_hi, idesc, (accumulate || ks > 0) ? 1u : 0u);
        mma_bf16_ss(d_tmem, da_lo, db_hi, idesc, 1u);
        mma_bf16_ss(d_tmem, da_hi, db_lo, idesc, 1u);
      }
      mma_commit(&p.s->empty[slot]);
    }
    if (signal_done) mma_commit(&p.s->done);
  }
}

}  // namespace tc
