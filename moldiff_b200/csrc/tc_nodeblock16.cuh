// 16-row-warp variant of tc_nodeblock_fwd_kernel: four threads per tile row (64 columns each) instead of two.
// The row epilogues of the 8-warp kernel are issue/latency bound (IPC ~1 per SM); doubling the resident row
// warps (4 per scheduler) is the cheapest way to overlap their dependent instruction chains.  Same math, same
// pipeline protocol; LayerNorm statistics are merged exactly over the four parts of a row.
//
// Included by mdb_forward.cu inside its anonymous namespace (after tc_bondffn.cuh).
#pragma once
#include "tc_pipe.cuh"

constexpr int NB16_NRW = 16;
constexpr int NB16_THREADS = (NB16_NRW + 4) * 32;   // 640: 4 row warpgroups + {producer, MMA, 2 idle warps}
constexpr int NB16_NS = 3;                           // weight stages (48 KB): leaves room for the 4-way stat buffer
// smem: E planes | X planes | weight stages | pipe barriers (128 B) | LN stat [4][128] float2 | ls [128] int | vecs [8][256]
constexpr size_t NB16_VEC_OFF = 2 * (size_t)tc::ROWS * C * 2 + 2 * (size_t)tc::ROWS * D * 2 + NB16_NS * tc::STAGE_SLOT
                                + 128 + 4 * tc::ROWS * sizeof(float2) + tc::ROWS * sizeof(int);
static_assert(NB16_VEC_OFF % 16 == 0, "vector block must be 16-byte aligned");
using Pipe16 = tc::PipeT<NB16_NS>;
using PipeSmem16 = tc::PipeSmemT<NB16_NS>;

// LayerNorm statistics of a 256-wide row held as 4 x 64 columns by four threads: returns (mean, rstd)
__device__ __forceinline__ float2 ln_stats_quarter(const float (&v)[64], float2* stat, int row, int part) {
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 64; i += 4) { s4[0] += v[i]; s4[1] += v[i + 1]; s4[2] += v[i + 2]; s4[3] += v[i + 3]; }
  const float m_p = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.f / 64.f);
  float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 64; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { const float d = v[i + u] - m_p; q4[u] = fmaf(d, d, q4[u]); }
  }
  stat[part * tc::ROWS + row] = make_float2(m_p, (q4[0] + q4[1]) + (q4[2] + q4[3]));
  asm volatile("bar.sync 1, 512;" ::: "memory");
  const float2 s0 = stat[row], s1 = stat[tc::ROWS + row], s2 = stat[2 * tc::ROWS + row], s3 = stat[3 * tc::ROWS + row];
  asm volatile("bar.sync 1, 512;" ::: "memory");
  const float mean = 0.25f * ((s0.x + s1.x) + (s2.x + s3.x));
  const float d0 = s0.x - mean, d1 = s1.x - mean, d2 = s2.x - mean, d3 = s3.x - mean;
  const float m2 = (s0.y + s1.y) + (s2.y + s3.y) + 64.f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));   // exact merge
  return make_float2(mean, 1.f / sqrtf(m2 * (1.f / 256.f) + LN_EPS));
}

// v[i] = f(acc[i], v[i]) for this thread's 64 accumulator columns, read from TMEM in two 32-column chunks: v[] already
// holds the gathered operand (loaded BEFORE the accumulator wait), so the peak register footprint is 64 + 32.
template <typename F>
__device__ __forceinline__ void combine_cols64(uint32_t taddr, float (&v)[64], F&& f) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tc::tmem_ld32_issue(taddr + c * 32, r);
    tc::tmem_ld32_wait(r);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[c * 32 + i] = f(c * 32 + i, tc::acc_f(r[i]), v[c * 32 + i]);
  }
}

template <bool IS_ROW>
__device__ __forceinline__ void tc_nodeblock_fwd16_body(const TcNbArgs& a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* e_hi = smem_raw;
  uint8_t* e_lo = e_hi + tc::ROWS * C * 2;
  uint8_t* x_hi = e_lo + tc::ROWS * C * 2;
  uint8_t* x_lo = x_hi + tc::ROWS * D * 2;
  uint8_t* stages = x_lo + tc::ROWS * D * 2;
  PipeSmem16* ps = reinterpret_cast<PipeSmem16*>(stages + NB16_NS * tc::STAGE_SLOT);
  float2* stat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(ps) + 128);     // [4][128]
  int* ls = reinterpret_cast<int*>(stat + 4 * tc::ROWS);
  // per-column parameter vectors, copied once from the kernel arguments: register-indexed constant-bank loads (LDC, two
  // values per instruction) were the epilogues' throughput limiter; warp-uniform LDS.128 reads broadcast four per instruction
  float* vecs = reinterpret_cast<float*>(smem_raw + NB16_VEC_OFF);
  const float* v_en1_b = vecs, *v_en1_g = vecs + D, *v_en1_be = vecs + 2 * D, *v_en2_b = vecs + 3 * D, *v_msg_b = vecs + 4 * D,
              *v_g1_g = vecs + 5 * D, *v_g1_be = vecs + 6 * D, *v_g2_b = vecs + 7 * D;
  float* out_tile = reinterpret_cast<float*>(smem_raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * tc::ROWS;
  const Tables& tb = a.tb;
  TC_STAMP(0);
  Pipe16 p;
  tc::pipe_init_split<NB16_NRW, IS_ROW, NB16_NS>(p, ps, stages);
  if (a.dbg) p.dbg = a.dbg + (size_t)blockIdx.x * 32;
  if (warp == NB16_NRW) tc::tmem_alloc<512>(&ps->tmem_base);
  const int row = (warp & 3) * 32 + lane;
  const int part = (warp >> 2) & 3;
  const int pc = part * 64;                   // first column of this thread's quarter
  int my_r = -1;
  float e16[16];                              // this thread's 16 columns of the e tile: in flight across the set-up barrier
  if (IS_ROW) {
    const int q = q0 + row;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < a.n_edges) x = *reinterpret_cast<const float4*>(a.ebuf + (size_t)q * C + part * 16 + i);
      e16[i] = x.x; e16[i + 1] = x.y; e16[i + 2] = x.z; e16[i + 3] = x.w;
    }
    if (q < a.n_edges) { my_r = a.right[q]; if (part == 0) ls[row] = a.left[q]; }
    else if (part == 0) ls[row] = -1;
    static_assert(sizeof(NbVecs) == 8 * D * sizeof(float), "NbVecs layout");
    const float* src = reinterpret_cast<const float*>(&a.v);
#pragma unroll
    for (int i = 0; i < 8 * D / (NB16_NRW * 32); ++i) vecs[i * NB16_NRW * 32 + tid] = src[i * NB16_NRW * 32 + tid];
  }
  tc::fence_before_sync();
  tc::cta_sync();
  tc::fence_after_sync();
  const uint32_t lane_base = ps->tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t D0 = 0, D1 = 256;
  const int rr = my_r < 0 ? 0 : my_r;
  TC_STAMP(1);

  if (IS_ROW) {   // e tile -> E planes: 16 columns per thread
    tc::store_a<C, 16>(e_hi, e_lo, row, part * 16, e16);
    tc::rows_publish(p);
    TC_STAMP(6);
  }
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_EN1), D1, false, true, true);
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    TC_STAMP(2);
    float v[64];
    load_cols_tm<64>(lane_base + D1 + pc, v);
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] += v_en1_b[pc + i];
    const float2 ms = ln_stats_quarter(v, stat, row, part);
    const float* gam = v_en1_g + pc;
    const float* bet = v_en1_be + pc;
    tc::store_a_sliced<64>(p, x_hi, x_lo, row, pc,
                           [&](int i) { return fmaxf((v[i] - ms.x) * ms.y * gam[i] + bet[i], 0.f); });
    TC_STAMP(7);
  }
  tc::gemm<D, D, NB16_NS, 4>(p, x_hi, x_lo, TCW_(NB_EN2), D1, false, true, true);
  if (IS_ROW) {
    const float* hn = tb.hn + (size_t)rr * D + pc;
    float v[64];                                 // gathered node_net(x)[col] row part: requested BEFORE the accumulator wait
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
      const float4 t4 = *reinterpret_cast<const float4*>(hn + i);
      v[i] = t4.x; v[i + 1] = t4.y; v[i + 2] = t4.z; v[i + 3] = t4.w;
    }
    tc::rows_wait_acc(p);
    TC_STAMP(3);
    const float* b2 = v_en2_b + pc;
    combine_cols64(lane_base + D1 + pc, v, [&](int i, float acc, float h) { return (acc + b2[i]) * h; });
    tc::store_a_sliced<64>(p, x_hi, x_lo, row, pc, [&](int i) { return v[i]; });
    TC_STAMP(8);
  }
  tc::gemm<D, D, NB16_NS, 4>(p, x_hi, x_lo, TCW_(NB_MSG), D0, false, true, false);
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_GE), D1, false, false, true);
  if (IS_ROW) {
    const float* gxr = tb.gx + (size_t)rr * D + pc;
    float v[64];                                 // hoisted node / time / bias part of gate.net.0, gathered before the wait
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
      const float4 t4 = *reinterpret_cast<const float4*>(gxr + i);
      v[i] = t4.x; v[i + 1] = t4.y; v[i + 2] = t4.z; v[i + 3] = t4.w;
    }
    tc::rows_wait_acc(p);
    TC_STAMP(4);
    combine_cols64(lane_base + D1 + pc, v, [&](int, float acc, float g) { return acc + g; });
    const float2 ms = ln_stats_quarter(v, stat, row, part);
    const float* gam = v_g1_g + pc;
    const float* bet = v_g1_be + pc;
    tc::store_a_sliced<64>(p, x_hi, x_lo, row, pc,
                           [&](int i) { return fmaxf((v[i] - ms.x) * ms.y * gam[i] + bet[i], 0.f); });
    TC_STAMP(9);
  }
  tc::gemm<D, D, NB16_NS, 4>(p, x_hi, x_lo, TCW_(NB_G2), D1, false, true, true);
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    TC_STAMP(5);
    {
      float g[64], m[64];
      load_cols_tm<64>(lane_base + D1 + pc, g);
      load_cols_tm<64>(lane_base + D0 + pc, m);
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        float4 o;
        o.x = (m[i] + v_msg_b[pc + i]) * tc::fast_sigmoid(g[i] + v_g2_b[pc + i]);
        o.y = (m[i + 1] + v_msg_b[pc + i + 1]) * tc::fast_sigmoid(g[i + 1] + v_g2_b[pc + i + 1]);
        o.z = (m[i + 2] + v_msg_b[pc + i + 2]) * tc::fast_sigmoid(g[i + 2] + v_g2_b[pc + i + 2]);
        o.w = (m[i + 3] + v_msg_b[pc + i + 3]) * tc::fast_sigmoid(g[i + 3] + v_g2_b[pc + i + 3]);
        *reinterpret_cast<float4*>(out_tile + row * OUT_LD + pc + i) = o;
      }
    }
    tc::fence_before_sync();
    asm volatile("bar.sync 1, 512;" ::: "memory");
    // scatter_sum over row (= left): thread (c, h) owns channel c and walks rows [64 h, 64 h + 64)   graph.py:50
    const int c = tid & 255, r0 = (tid >> 8) * 64;
    int cur = ls[r0];
    float s0 = 0.f;
#pragma unroll 1
    for (int r8 = r0; r8 < r0 + 64; r8 += 8) {
      float x[8];
      int n[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { x[u] = out_tile[(r8 + u) * OUT_LD + c]; n[u] = ls[r8 + u]; }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (n[u] != cur) {
          if (cur >= 0) atomicAdd(tb.agg + (size_t)cur * D + c, s0);
          cur = n[u]; s0 = 0.f;
        }
        s0 += x[u];
      }
    }
    if (cur >= 0) atomicAdd(tb.agg + (size_t)cur * D + c, s0);
  }
  TC_STAMP(15);
  tc::cta_sync();
  if (warp == NB16_NRW) { __syncwarp(); tc::tmem_dealloc<512>(ps->tmem_base); }
}

// 640 threads compile to 96 registers; the producer/MMA warpgroup keeps 32, the four row warpgroups get 112.
__global__ void __launch_bounds__(NB16_THREADS, 1) tc_nodeblock_fwd16_kernel(const __grid_constant__ TcNbArgs a) {
  if (threadIdx.x < NB16_NRW * 32) {
    tc::reg_alloc<112>();
    tc_nodeblock_fwd16_body<true>(a);
  } else {
    tc::reg_dealloc<32>();
    tc_nodeblock_fwd16_body<false>(a);
  }
}

constexpr size_t SMEM_TC_NB16 = NB16_VEC_OFF + 8 * D * sizeof(float);
static_assert(sizeof(PipeSmem16) <= 128, "PipeSmem must fit its 128-byte slot");
static_assert(SMEM_TC_NB16 <= 232448, "tc_nodeblock_fwd16_kernel exceeds the shared-memory limit");
