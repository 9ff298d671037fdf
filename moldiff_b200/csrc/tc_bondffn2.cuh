// tc_bondffn_fwd2_kernel: the work of tc_bondffn_fwd_kernel (RBF, edge_embs, both BondFFNs of the EdgeBlock; reference
// models/graph.py:354-357, 133-141, 278-284) sized for TWO resident CTAs per SM.
//
// Why: a BondFFN tile is a chain of 11 small GEMM phases (<= 1.5 k cycles of MMA each) separated by per-row epilogues; with
// one tile per SM (tc_bondffn_fwd_kernel: 176 KB of shared memory) the tensor pipe idles during every epilogue and the CUDA
// cores during every publish -> issue -> commit -> wake round trip (phase table: 22 k of 49 k cycles per tile are such waits,
// tensor pipe 13 % busy).  Two co-resident CTAs let the hardware interleave the round trips of one tile with the epilogues of
// the other.  To fit twice (<= 113 KB, 256 TMEM columns, <= 96 registers at 320 threads):
//   * ONE 64 KB operand region holds, in turn, the K = 80 input planes, the K = 64 e planes, the K = 128 planes of the two
//     hidden layers and the fp32 output tile; the e planes needed again by the right FFN are rebuilt from `ebuf` (this
//     thread's own 32 floats, an L2 hit) instead of being kept alive;
//   * 3 weight-ring slots of 8 KB (widest N here is 128) instead of 4 x 16 KB;
//   * rolled epilogues over 16-column chunks that re-read TMEM, per-column vectors in shared memory (warp-uniform LDS.128),
//     the gate sigmoid evaluated after the last GEMM from the parked gate accumulator instead of 32 live registers.
// Same arithmetic as tc_bondffn_fwd_kernel (MDB_TC_FFN2=0 selects that one for A/B runs).
//
// Included by mdb_forward.cu inside its anonymous namespace (after tc_nodeblock_bwd16.cuh: RunStat, Row16 helpers).
#pragma once
#include "tc_pipe.cuh"

constexpr int FFN2_NS = 3;
constexpr uint32_t FFN2_SLOT = tc::WStage<128, tc::KB>::STAGE_BYTES;                 // 8 KB
constexpr size_t FFN2_OFF_G = 2 * (size_t)tc::ROWS * 128 * 2;                         // gate planes (K = 32) after the 64 KB region
constexpr size_t FFN2_OFF_RING = FFN2_OFF_G + 2 * (size_t)tc::ROWS * 32 * 2;
constexpr size_t FFN2_OFF_PS = FFN2_OFF_RING + FFN2_NS * FFN2_SLOT;
constexpr size_t FFN2_OFF_STAT = FFN2_OFF_PS + 128;
constexpr size_t FFN2_OFF_LS = FFN2_OFF_STAT + 2 * tc::ROWS * sizeof(float2);
constexpr size_t FFN2_OFF_VEC = FFN2_OFF_LS + tc::ROWS * sizeof(int);
constexpr size_t SMEM_TC_FFN2 = FFN2_OFF_VEC + sizeof(FfnVecs);
static_assert(sizeof(tc::PipeSmemT<FFN2_NS>) <= 128, "PipeSmem must fit its 128-byte slot");
static_assert(FFN2_OFF_VEC % 16 == 0 && sizeof(FfnVecs) % 16 == 0, "vector block must be 16-byte aligned");
static_assert(2 * (SMEM_TC_FFN2 + 1024) <= 233472, "tc_bondffn_fwd2_kernel must fit twice per SM");
static_assert((size_t)tc::ROWS * FFN_O_LD * 4 <= FFN2_OFF_G, "o tile must fit in the operand region");

template <bool XF>   // cross-first accumulation order (tc_pipe.cuh)
__global__ void __launch_bounds__(TC_NB_THREADS, 2) tc_bondffn_fwd2_kernel(const __grid_constant__ TcFfnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int KI = C + G;                                  // 80
  uint8_t* a_hi = smem_raw;                                  // K = 128 planes: 2 x 32 KB
  uint8_t* a_lo = a_hi + tc::ROWS * 128 * 2;
  uint8_t* in_hi = smem_raw;                                 // K = 80 input planes (dead once e is in TMEM)
  uint8_t* in_lo = in_hi + tc::ROWS * KI * 2;
  uint8_t* e_hi = smem_raw;                                  // K = 64 e planes (dead once bond_linear / gate.0 are in TMEM)
  uint8_t* e_lo = e_hi + tc::ROWS * C * 2;
  uint8_t* g_hi = smem_raw + FFN2_OFF_G;                     // gate planes, K = 32
  uint8_t* g_lo = g_hi + tc::ROWS * 32 * 2;
  uint8_t* stages = smem_raw + FFN2_OFF_RING;
  tc::PipeSmemT<FFN2_NS>* ps = reinterpret_cast<tc::PipeSmemT<FFN2_NS>*>(smem_raw + FFN2_OFF_PS);
  float2* stat = reinterpret_cast<float2*>(smem_raw + FFN2_OFF_STAT);
  int* ls = reinterpret_cast<int*>(smem_raw + FFN2_OFF_LS);
  float* vecs = reinterpret_cast<float*>(smem_raw + FFN2_OFF_VEC);
  const FfnVecs& vv = *reinterpret_cast<const FfnVecs*>(vecs);
  float* o_tile = reinterpret_cast<float*>(smem_raw);        // [128][FFN_O_LD] fp32 at the very end (planes dead)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  FFN_STAMP(0);
  const int q0 = blockIdx.x * tc::ROWS;
  const Tables& tb = a.tb;
  tc::PipeT<FFN2_NS, XF> p;
  tc::pipe_init<TC_NRW>(p, ps, stages);
  p.slot_bytes = FFN2_SLOT;
  if (warp == TC_NRW) tc::tmem_alloc<256>(&ps->tmem_base);
  {   // parameter vectors: constant bank -> shared memory, the (slow, register-indexed) LDCs all in flight before the first store
    constexpr int NV = (int)(sizeof(FfnVecs) / sizeof(float)), NIT = (NV + TC_NB_THREADS - 1) / TC_NB_THREADS;
    float t[NIT];
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int i = tid + k * TC_NB_THREADS;
      t[k] = reinterpret_cast<const float*>(&a.v)[i < NV ? i : 0];
    }
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int i = tid + k * TC_NB_THREADS;
      if (i < NV) vecs[i] = t[k];
    }
  }
  const int row = (warp & 3) * 32 + lane;
  const int half = (warp >> 2) & 1;
  const int q = q0 + row;
  const bool valid = p.role == 0 && q < a.n_edges;
  const int ll = valid ? a.left[q] : 0, rr = valid ? a.right[q] : 0;
  const float te = valid ? tb.te[q] : 0.f;
  if (p.role == 0 && half == 0) ls[row] = valid ? ll : -1;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t lane_base = ps->tmem_base + ((uint32_t)((warp & 3) * 32) << 16);

  // ---- input planes [h_edge ; rbf(d)] (K = 80): each thread 32 h_edge columns + 8 of the 16 radial basis values
  if (p.role == 0) {
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) x = *reinterpret_cast<const float4*>(tb.hedge + (size_t)q * C + half * 32 + c * 16 + i);
        v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
      }
      tc::store_a<KI, 16>(in_hi, in_lo, row, half * 32 + c * 16, v);
    }
    const float dx = a.pos[ll * 3 + 0] - a.pos[rr * 3 + 0];
    const float dy = a.pos[ll * 3 + 1] - a.pos[rr * 3 + 1];
    const float dz = a.pos[ll * 3 + 2] - a.pos[rr * 3 + 2];
    const float d = fminf(fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), a.rbf_lo), a.rbf_hi);
    float g[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float u = d - vv.rbf_off[half * 8 + k];
      g[k] = valid ? expf(vv.rbf_coeff[half * 8 + k] * (u * u)) : 0.f;
    }
    tc::store_a<KI, 8>(in_hi, in_lo, row, C + half * 8, g);
    tc::rows_publish(p);
    FFN_STAMP(1);
  }
  // G0: e = edge_embs(cat)                                                        graph.py:354-357
  tc::gemm<KI, C>(p, in_hi, in_lo, TCW_(EE), 0, false, true, true);
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    FFN_STAMP(2);
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      float v[16], b[16];
      tc::tmem_ld16(lane_base + half * 32 + c * 16, v);
      lds16(vv.ee_b + half * 32 + c * 16, b);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += b[i];
      if (valid) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(a.ebuf + (size_t)q * C + half * 32 + c * 16 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      tc::store_a<C, 16>(e_hi, e_lo, row, half * 32 + c * 16, v);     // (the input planes are dead: G0 has completed)
    }
    tc::rows_publish(p);
    FFN_STAMP(3);
  }
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    const FfnSideVecs& sv = vv.side[side];
    const int node = side ? rr : ll;                       // node feeding this FFN        graph.py:278,282
    const uint8_t* w_bl = side ? TCW_(ER_BL) : TCW_(EL_BL);
    const uint8_t* w_gb = side ? TCW_(ER_GB) : TCW_(EL_GB);
    const uint8_t* w_i1 = side ? TCW_(ER_I1) : TCW_(EL_I1);
    const uint8_t* w_g2 = side ? TCW_(ER_G2) : TCW_(EL_G2);
    const uint8_t* w_i2 = side ? TCW_(ER_I2) : TCW_(EL_I2);
    // bond_linear(e) -> D[0:128] ; gate.net.0 bond columns -> D[128:160]          graph.py:134,139
    tc::gemm<C, 128>(p, e_hi, e_lo, w_bl, 0, false, true, false);
    tc::gemm<C, 32>(p, e_hi, e_lo, w_gb, 128, false, false, true);
    if (p.role == 0) {
      const float* nl = (side ? tb.nlr : tb.nll) + (size_t)node * 128 + half * 64;
      const float* gn = (side ? tb.gnr : tb.gnl) + (size_t)node * 32 + half * 16;
      Row16 nx = ld_row16(nl), gq = ld_row16(gn);            // requested before the accumulator wait
      tc::rows_wait_acc(p);
      FFN_STAMP(4 + side * 8);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {                          // * node_linear(h_node)[.] -> K = 128 planes (e planes are dead)
        float x[16], h[16];
        unpack_row16(nx, h);
        if (c < 3) nx = ld_row16(nl + (c + 1) * 16);
        tc::tmem_ld16(lane_base + half * 64 + c * 16, x);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] *= h[i];
        tc::store_a<128, 16>(a_hi, a_lo, row, half * 64 + c * 16, x);
      }
      {
        float g[16], gv[16], w[16];
        unpack_row16(gq, gv);
        tc::tmem_ld16(lane_base + 128 + half * 16, g);
        lds16(sv.gt_w + half * 16, w);
#pragma unroll
        for (int i = 0; i < 16; ++i) g[i] += gv[i] + te * w[i];
        ln_relu_part<16>(g, sv.g1_g + half * 16, sv.g1_be + half * 16, stat, row, half);
        tc::store_a<32, 16>(g_hi, g_lo, row, half * 16, g);
      }
      tc::rows_publish(p);
      FFN_STAMP(5 + side * 8);
    }
    // inter_module.net.0 -> D[0:128] ; gate.net.3 -> D[128:192]                   graph.py:137,139
    tc::gemm<128, 128>(p, a_hi, a_lo, w_i1, 0, false, true, false);
    tc::gemm<32, C>(p, g_hi, g_lo, w_g2, 128, false, false, true);
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      FFN_STAMP(6 + side * 8);
      RunStat rs = {0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float x[16], b[16];
        tc::tmem_ld16(lane_base + half * 64 + c * 16, x);
        lds16(sv.i1_b + half * 64 + c * 16, b);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] += b[i];
        stat_add16(rs, x);
      }
      const float2 o = exchange_half(stat, row, half, rs.mean, rs.m2);
      const float mean = 0.5f * (rs.mean + o.x);
      const float dm = rs.mean - o.x;
      const float rstd = 1.f / sqrtf((rs.m2 + o.y + dm * dm * 32.f) * (1.f / 128.f) + LN_EPS);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float x[16], b[16], ga[16], be[16];
        tc::tmem_ld16(lane_base + half * 64 + c * 16, x);
        lds16(sv.i1_b + half * 64 + c * 16, b); lds16(sv.i1_g + half * 64 + c * 16, ga); lds16(sv.i1_be + half * 64 + c * 16, be);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaxf((x[i] + b[i] - mean) * rstd * ga[i] + be[i], 0.f);
        tc::store_a<128, 16>(a_hi, a_lo, row, half * 64 + c * 16, x);
      }
      tc::rows_publish(p);
      FFN_STAMP(7 + side * 8);
    }
    // inter_module.net.3 -> D[0:64]  (the gate pre-activation stays parked in D[128:192])      graph.py:137
    tc::gemm<128, C>(p, a_hi, a_lo, w_i2, 0, false, true, true);
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      FFN_STAMP(8 + side * 8);
      if (side == 0) {
        // bond_ffn_left output is scattered over right_node (unsorted): one 16-byte RED per 4 channels   graph.py:279
        float* dst = a.sl + (size_t)rr * C + half * 32;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float o[16], g[16], b[16], bg[16];
          tc::tmem_ld16(lane_base + half * 32 + c * 16, o);
          tc::tmem_ld16(lane_base + 128 + half * 32 + c * 16, g);
          lds16(sv.i2_b + half * 32 + c * 16, b); lds16(sv.g2_b + half * 32 + c * 16, bg);
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = (o[i] + b[i]) * tc::fast_sigmoid(g[i] + bg[i]);
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) tc::red_add_v4(dst + c * 16 + i, o[i], o[i + 1], o[i + 2], o[i + 3]);
          }
        }
        // e planes for the right FFN: this thread's own 32 floats of e come back from ebuf (the K = 128 planes are dead)
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) x = *reinterpret_cast<const float4*>(a.ebuf + (size_t)q * C + half * 32 + c * 16 + i);
            v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
          }
          tc::store_a<C, 16>(e_hi, e_lo, row, half * 32 + c * 16, v);
        }
        FFN_STAMP(9);
        tc::rows_publish(p);     // accumulator drained, e planes rebuilt: the right FFN's first GEMM may start
      } else {
        // bond_ffn_right output is scattered over left_node = contiguous CSR runs: reduce in the tile   graph.py:283
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float o[16], g[16], b[16], bg[16];
          tc::tmem_ld16(lane_base + half * 32 + c * 16, o);
          tc::tmem_ld16(lane_base + 128 + half * 32 + c * 16, g);
          lds16(sv.i2_b + half * 32 + c * 16, b); lds16(sv.g2_b + half * 32 + c * 16, bg);
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = (o[i] + b[i]) * tc::fast_sigmoid(g[i] + bg[i]);
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(o_tile + row * FFN_O_LD + half * 32 + c * 16 + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
        }
        tc::fence_before_sync();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float* sr = a.sl + (size_t)a.n_nodes * C;
        const int c = tid & 63, r0 = (tid >> 6) * 32;       // 4 row quarters x 64 channels
        int cur = ls[r0];
        float s = 0.f;
        for (int r = r0; r < r0 + 32; ++r) {
          const int n = ls[r];
          if (n != cur) {
            if (cur >= 0) atomicAdd(sr + (size_t)cur * C + c, s);
            cur = n; s = 0.f;
          }
          s += o_tile[r * FFN_O_LD + c];
        }
        if (cur >= 0) atomicAdd(sr + (size_t)cur * C + c, s);
        FFN_STAMP(17);
      }
    }
  }
  __syncthreads();
  if (warp == TC_NRW) { __syncwarp(); tc::tmem_dealloc<256>(ps->tmem_base); }
}
