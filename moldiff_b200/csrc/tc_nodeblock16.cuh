// 16-row-warp variant of tc_nodeblock_fwd_kernel: four threads per tile row (64 columns each) instead of two.
// The row epilogues of the 8-warp kernel are issue/latency bound (IPC ~1 per SM); doubling the resident row
// warps (4 per scheduler) is the cheapest way to overlap their dependent instruction chains.  Same math, same
// pipeline protocol; LayerNorm statistics are merged exactly over the four parts of a row.
//
// Included by mdb_forward.cu inside its anonymous namespace (after tc_bondffn.cuh).
#pragma once
#include "tc_pipe.cuh"

constexpr int NB16_NRW = 16;
constexpr int NB16_THREADS = (NB16_NRW + 2) * 32;   // 576
constexpr int NB16_NS = 3;                           // weight stages (48 KB): leaves room for the 4-way stat buffer
using Pipe16 = tc::PipeT<NB16_NS>;
using PipeSmem16 = tc::PipeSmemT<NB16_NS>;

// LayerNorm over a 256-wide row held as 4 x 64 columns by four threads; in place: v <- relu(LN(v) * g + b)
__device__ __forceinline__ void ln_relu_quarter(float (&v)[64], const float* __restrict__ gamma,
                                                const float* __restrict__ beta, float2* stat, int row, int part) {
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
  const float rstd = 1.f / sqrtf(m2 * (1.f / 256.f) + LN_EPS);
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = fmaxf((v[i] - mean) * rstd * gamma[i] + beta[i], 0.f);
}

__global__ void __launch_bounds__(NB16_THREADS, 1) tc_nodeblock_fwd16_kernel(const __grid_constant__ TcNbArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* e_hi = smem_raw;
  uint8_t* e_lo = e_hi + tc::ROWS * C * 2;
  uint8_t* x_hi = e_lo + tc::ROWS * C * 2;
  uint8_t* x_lo = x_hi + tc::ROWS * D * 2;
  uint8_t* stages = x_lo + tc::ROWS * D * 2;
  PipeSmem16* ps = reinterpret_cast<PipeSmem16*>(stages + NB16_NS * tc::STAGE_SLOT);
  float2* stat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(ps) + 128);     // [4][128]
  int* ls = reinterpret_cast<int*>(stat + 4 * tc::ROWS);
  float* out_tile = reinterpret_cast<float*>(smem_raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * tc::ROWS;
  const Tables& tb = a.tb;
  Pipe16 p;
  tc::pipe_init<NB16_NRW, NB16_NS>(p, ps, stages);
  if (warp == NB16_NRW) tc::tmem_alloc<512>(&ps->tmem_base);
  const int row = (warp & 3) * 32 + lane;
  const int part = (warp >> 2) & 3;
  const int pc = part * 64;                   // first column of this thread's quarter
  int my_r = -1;
  if (p.role == 0) {
    const int q = q0 + row;
    if (q < a.n_edges) { my_r = a.right[q]; if (part == 0) ls[row] = a.left[q]; }
    else if (part == 0) ls[row] = -1;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t lane_base = ps->tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t D0 = 0, D1 = 256;
  const int rr = my_r < 0 ? 0 : my_r;

  if (p.role == 0) {   // e tile -> E planes: 16 columns per thread
    const int q = q0 + row;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < a.n_edges) x = *reinterpret_cast<const float4*>(a.ebuf + (size_t)q * C + part * 16 + i);
      v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
    }
    tc::store_a<C, 16>(e_hi, e_lo, row, part * 16, v);
    tc::rows_publish(p);
  }
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_EN1), D1, false, true, true);
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    float v[64];
    load_cols_tm<64>(lane_base + D1 + pc, v);
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] += a.v.en1_b[pc + i];
    ln_relu_quarter(v, a.v.en1_g + pc, a.v.en1_be + pc, stat, row, part);
    tc::store_a<D, 64>(x_hi, x_lo, row, pc, v);
    tc::rows_publish(p);
  }
  tc::gemm<D, D>(p, x_hi, x_lo, TCW_(NB_EN2), D1, false, true, true);
  if (p.role == 0) {
    const float* hn = tb.hn + (size_t)rr * D + pc;
    tc::rows_wait_acc(p);
    float v[64];
    load_cols_tm<64>(lane_base + D1 + pc, v);
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
      const float4 t4 = *reinterpret_cast<const float4*>(hn + i);
      v[i] = (v[i] + a.v.en2_b[pc + i]) * t4.x;
      v[i + 1] = (v[i + 1] + a.v.en2_b[pc + i + 1]) * t4.y;
      v[i + 2] = (v[i + 2] + a.v.en2_b[pc + i + 2]) * t4.z;
      v[i + 3] = (v[i + 3] + a.v.en2_b[pc + i + 3]) * t4.w;
    }
    tc::store_a<D, 64>(x_hi, x_lo, row, pc, v);
    tc::rows_publish(p);
  }
  tc::gemm<D, D>(p, x_hi, x_lo, TCW_(NB_MSG), D0, false, true, false);
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_GE), D1, false, false, true);
  if (p.role == 0) {
    const float* gxr = tb.gx + (size_t)rr * D + pc;
    tc::rows_wait_acc(p);
    float v[64];
    load_cols_tm<64>(lane_base + D1 + pc, v);
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
      const float4 t4 = *reinterpret_cast<const float4*>(gxr + i);
      v[i] += t4.x; v[i + 1] += t4.y; v[i + 2] += t4.z; v[i + 3] += t4.w;
    }
    ln_relu_quarter(v, a.v.g1_g + pc, a.v.g1_be + pc, stat, row, part);
    tc::store_a<D, 64>(x_hi, x_lo, row, pc, v);
    tc::rows_publish(p);
  }
  tc::gemm<D, D>(p, x_hi, x_lo, TCW_(NB_G2), D1, false, true, true);
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    {
      float g[64], m[64];
      load_cols_tm<64>(lane_base + D1 + pc, g);
      load_cols_tm<64>(lane_base + D0 + pc, m);
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        float4 o;
        o.x = (m[i] + a.v.msg_b[pc + i]) * tc::fast_sigmoid(g[i] + a.v.g2_b[pc + i]);
        o.y = (m[i + 1] + a.v.msg_b[pc + i + 1]) * tc::fast_sigmoid(g[i + 1] + a.v.g2_b[pc + i + 1]);
        o.z = (m[i + 2] + a.v.msg_b[pc + i + 2]) * tc::fast_sigmoid(g[i + 2] + a.v.g2_b[pc + i + 2]);
        o.w = (m[i + 3] + a.v.msg_b[pc + i + 3]) * tc::fast_sigmoid(g[i + 3] + a.v.g2_b[pc + i + 3]);
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
  __syncthreads();
  if (warp == NB16_NRW) { __syncwarp(); tc::tmem_dealloc<512>(ps->tmem_base); }
}

constexpr size_t SMEM_TC_NB16 = 2 * (size_t)tc::ROWS * C * 2 + 2 * (size_t)tc::ROWS * D * 2 + NB16_NS * tc::STAGE_SLOT
                                + 128 + 4 * tc::ROWS * sizeof(float2) + tc::ROWS * sizeof(int) + 64;
static_assert(sizeof(PipeSmem16) <= 128, "PipeSmem must fit its 128-byte slot");
static_assert(SMEM_TC_NB16 <= 232448, "tc_nodeblock_fwd16_kernel exceeds the shared-memory limit");
