// tc_bwd_node_kernel: the per-NODE part of the bond predictor's input-gradient backward on tcgen05 (same math as the fp32
// bwd_node_kernel in mdb_backward.cuh, which it replaces when the transposed operand images are packed):
//
//   phase B (block i+1):  dx += dul W_nfl + dur W_nfr + dnl_L W_nlL + dnl_R W_nlR + dgn_L W_gnL + dgn_R W_gnR + dgx W_gx
//                         node_net backward: hn = W2 relu(LN(W1 x + b1)) + b2           (forward: graph.py:39)
//                            d r = dhn W2 ;  d a1 = LNbwd(d r; a1 = x W1^T + b1) ;  dx += d a1 W1
//   phase A (block i):    NodeBlock node tail backward: x' = x + W_out relu(LN(cen + agg)) + b  (graph.py:51-54, 363)
//                            d r = dx W_out ;  dc = LNbwd(d r; cen + agg) ;  dagg = dc ;  dx += dc W_cen
//   first launch of a backward: dx = ddect W_dec1n                                      (bond_predictor.py:155-160)
//
// 128 nodes per CTA, two threads per row (128 columns each), rolled 16-column epilogues.  Everything that is linear in the
// incoming gradients runs as a chain of five stages (fourteen K stages of seven different operands); within a stage the
// main terms (hi * hi) accumulate in D0 and the cross terms in D1 (tc_pipe.cuh: gemm_split), and after EVERY stage the row
// threads add D0 + D1 to the running gradient on the CUDA cores (round-to-nearest), so that each stage starts from zero.
// The hardware's truncating accumulate would otherwise shrink the running gradient coherently -- it is shared by every
// edge of the node, so the bias does not average out: guidance-gradient median error 1.7e-5 with one chained accumulator,
// 9e-6 with the main / cross split, ~6e-6 = the level of the fp32 node kernel with the per-stage fold.
// D1 is also the scratch accumulator of the two LayerNorm backwards; the normalised activations
// of the node tail are parked, as fp16 hi|lo planes, in the very shared-memory bytes that the dc operand planes overwrite.
// The running gradient `dx` lives in the node-blocked layout (tile_engine.cuh: blk_off) -- it is private to this kernel --
// so a warp's row threads read and write it as 512 contiguous bytes per instruction.
//
// Included by mdb_forward.cu inside its anonymous namespace (after tc_node.cuh).
#pragma once
#include "tc_pipe.cuh"

struct BwdNodeVecs { float nn1_b[D], nn1_g[D], nn1_be[D], ln_g[D], ln_be[D]; };

struct TcBwdNodeArgs {
  const uint8_t* tc_blob;
  TcOff blkB, blkA;          // operand images of block i+1 (phase B) and block i (phase A)
  TcOffHead hd;
  int n_nodes;
  int do_final, do_B, do_A;
  int red_blocked;           // dgx / dhn are node-blocked (written by tc_nodeblock_bwd16_kernel)
  const float* xB;           // [N][256] saved h_node entering block i+1
  const float *cenA, *aggA;  // [N][256] centroid_lin(x) table and aggregated messages of block i
  float* dx;                 // [pad64(N)][256] running d/d h_node, node-blocked (in / out)
  const float* ddect;        // [N][64]
  float *dul, *dur;          // [N][64]      read, then cleared for block i
  float* dnl;                // [2][N][128]
  float* dgn;                // [2][N][32]
  float *dgx, *dhn;          // [pad64(N)][256]
  float* dagg;               // [N][256] out
  BwdNodeVecs v;
};
#define TCBB_(slot) (a.tc_blob + a.blkB.o[MDB_T_##slot])
#define TCBA_(slot) (a.tc_blob + a.blkA.o[MDB_T_##slot])

// 16 values of row r, columns [k0, k0 + 16) of the K = 256 planes -> registers (hi + lo: the 22 bits that were stored)
__device__ __forceinline__ void load_a16(const uint8_t* a_hi, const uint8_t* a_lo, int r, int k0, float (&v)[16]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const uint32_t off = tc::a_chunk_off<256>(r, k0 / 8 + c);
    const uint4 h = *reinterpret_cast<const uint4*>(a_hi + off), l = *reinterpret_cast<const uint4*>(a_lo + off);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float h0, h1, l0, l1;
      tc::unpack_f16x2(hw[i], h0, h1);
      tc::unpack_f16x2(lw[i], l0, l1);
      v[c * 8 + 2 * i] = h0 + l0;
      v[c * 8 + 2 * i + 1] = h1 + l1;
    }
  }
}

// NC (multiple of 16) consecutive floats starting at `src` (16-byte aligned; nullptr -> zeros) -> columns [k0, k0 + NC) of
// row r of operand planes with K columns
template <int K, int NC>
__device__ __forceinline__ void stage_cols(uint8_t* hi, uint8_t* lo, int r, int k0, const float* __restrict__ src) {
#pragma unroll 1
  for (int c = 0; c < NC / 16; ++c) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src) x = *reinterpret_cast<const float4*>(src + c * 16 + i);
      v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
    }
    tc::store_a<K, 16>(hi, lo, r, k0 + c * 16, v);
  }
}
// 16 columns [k0, k0 + 16) of node n of a node-blocked [.][256] table (nullptr -> zeros)
__device__ __forceinline__ void ld_blocked16(const float* __restrict__ table, int n, int k0, bool ok, float (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) x = *reinterpret_cast<const float4*>(table + blk_off(n, k0 / 4 + j));
    v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
  }
}
__device__ __forceinline__ void st_blocked16(float* __restrict__ table, int n, int k0, const float (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<float4*>(table + blk_off(n, k0 / 4 + j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

constexpr size_t BN_OFF_STAGES = 2 * (size_t)tc::ROWS * D * 2;                       // after the 128 KB plane region
constexpr size_t BN_OFF_PS = BN_OFF_STAGES + tc::NSTAGE * tc::STAGE_SLOT;
constexpr size_t BN_OFF_STAT = BN_OFF_PS + 128;
constexpr size_t BN_OFF_VEC = BN_OFF_STAT + 2 * tc::ROWS * sizeof(float2);
constexpr size_t SMEM_TC_BWD_NODE = BN_OFF_VEC + sizeof(BwdNodeVecs);
static_assert(SMEM_TC_BWD_NODE <= 232448, "tc_bwd_node_kernel exceeds the shared-memory limit");
static_assert(BN_OFF_VEC % 16 == 0, "vector block must be 16-byte aligned");

__global__ void __launch_bounds__(TC_NB_THREADS, 1) tc_bwd_node_kernel(const __grid_constant__ TcBwdNodeArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* P = smem_raw;                                     // 128 KB plane region
  uint8_t* p_hi = P;                                         // as K = 256 planes
  uint8_t* p_lo = P + tc::ROWS * D * 2;
  uint8_t* stages = smem_raw + BN_OFF_STAGES;
  tc::PipeSmem* ps = reinterpret_cast<tc::PipeSmem*>(smem_raw + BN_OFF_PS);
  float2* stat = reinterpret_cast<float2*>(smem_raw + BN_OFF_STAT);
  float* vecs = reinterpret_cast<float*>(smem_raw + BN_OFF_VEC);
  const BwdNodeVecs& vv = *reinterpret_cast<const BwdNodeVecs*>(vecs);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * tc::ROWS;
  tc::PipeT<tc::NSTAGE, true> p;       // cross-first order for the stand-alone GEMMs; the chain uses gemm_split
  tc::pipe_init<TC_NRW>(p, ps, stages);
  if (warp == TC_NRW) tc::tmem_alloc<512>(&ps->tmem_base);
  {
    constexpr int NV = (int)(sizeof(BwdNodeVecs) / sizeof(float)), NIT = NV / TC_NB_THREADS;
    static_assert(NV % TC_NB_THREADS == 0, "vector block is a whole number of passes");
    float t[NIT];
#pragma unroll
    for (int k = 0; k < NIT; ++k) t[k] = reinterpret_cast<const float*>(&a.v)[tid + k * TC_NB_THREADS];
#pragma unroll
    for (int k = 0; k < NIT; ++k) vecs[tid + k * TC_NB_THREADS] = t[k];
  }
  const int row = (warp & 3) * 32 + lane;
  const int half = (warp >> 2) & 1;
  const int hc = half * 128;
  const int n = row0 + row;
  const bool valid = p.role == 0 && n < a.n_nodes;
  const size_t nn = valid ? (size_t)n : 0;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t lane_base = ps->tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t T0 = lane_base + hc, T1 = lane_base + 256 + hc;   // this thread's 128 columns of D0 / D1
  constexpr uint32_t D0 = 0, D1 = 256;
  bool have_dx = !a.do_final;            // false: the running gradient starts from zero (first launch of a backward)
  // dx += D0 + D1 on the CUDA cores (round-to-nearest), after EVERY stage of the chain: each stage then accumulates from zero,
  // so the hardware's truncating accumulate (tc_pipe.cuh) acts on that stage's partial sums only instead of on the running
  // total -- the per-node gradient is shared by every edge of the node, so its bias does not average out.
  auto fold = [&]() {
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      float v[16], d[16], x[16];
      tc::tmem_ld16(T0 + c * 16, d);
      tc::tmem_ld16(T1 + c * 16, x);
      ld_blocked16(a.dx, n, hc + c * 16, valid && have_dx, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += d[i] + x[i];
      if (valid) st_blocked16(a.dx, n, hc + c * 16, v);
    }
    have_dx = true;
  };

  if (a.do_B) {
    // ---- node_net backward first: it needs both accumulators, and D0 is free until the chain below starts --------------
    if (p.role == 0) {
      stage_cols<D, 128>(p_hi, p_lo, row, hc, valid ? a.xB + nn * D + hc : nullptr);
      tc::rows_publish(p);
    }
    tc::gemm<D, D>(p, p_hi, p_lo, TCBB_(NB_NN1), D1, false, true, true);              // a1 - b1 -> D1 (recompute)
    if (p.role == 0) {
      tc::rows_wait_acc(p);                                                           // x planes consumed
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float v[16];
        if (a.red_blocked) {
          ld_blocked16(a.dhn, n, hc + c * 16, valid, v);
        } else {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) x = *reinterpret_cast<const float4*>(a.dhn + nn * D + hc + c * 16 + i);
            v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
          }
        }
        store_a16(p_hi, p_lo, row, hc + c * 16, v);
      }
      tc::rows_publish(p);
    }
    tc::gemm<D, D>(p, p_hi, p_lo, TCBB_(BT_NB_NN2), D0, false, true, true);           // d relu1 -> D0
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      RunStat rs = {0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float x[16], b[16];
        tc::tmem_ld16(T1 + c * 16, x);
        lds16(vv.nn1_b + hc + c * 16, b);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] += b[i];
        stat_add16(rs, x);
      }
      const float2 o = exchange_half(stat, row, half, rs.mean, rs.m2);
      const float mean = 0.5f * (rs.mean + o.x);
      const float dm = rs.mean - o.x;
      const float rstd = 1.f / sqrtf((rs.m2 + o.y + dm * dm * 64.f) * (1.f / 256.f) + LN_EPS);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {                 // pass 1: d xhat -> D0, row sums
        float x[16], d[16], b[16], ga[16], be[16];
        tc::tmem_ld16(T1 + c * 16, x);
        tc::tmem_ld16(T0 + c * 16, d);
        lds16(vv.nn1_b + hc + c * 16, b); lds16(vv.nn1_g + hc + c * 16, ga); lds16(vv.nn1_be + hc + c * 16, be);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float xh = (x[i] + b[i] - mean) * rstd;
          const float dxh = (xh * ga[i] + be[i] > 0.f) ? d[i] * ga[i] : 0.f;
          d[i] = dxh;
          s1 += dxh;
          s2 = fmaf(dxh, xh, s2);
        }
        tc::tmem_st16(T0 + c * 16, d);
      }
      tc::tmem_st_wait();
      const float2 o2 = exchange_half(stat, row, half, s1, s2);
      const float m1 = (s1 + o2.x) * (1.f / 256.f), m2 = (s2 + o2.y) * (1.f / 256.f);
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {                 // pass 2: d a1 -> planes
        float x[16], d[16], b[16];
        tc::tmem_ld16(T1 + c * 16, x);
        tc::tmem_ld16(T0 + c * 16, d);
        lds16(vv.nn1_b + hc + c * 16, b);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float xh = (x[i] + b[i] - mean) * rstd;
          d[i] = rstd * (d[i] - m1 - xh * m2);
        }
        store_a16(p_hi, p_lo, row, hc + c * 16, d);
      }
      tc::rows_publish(p);
    }
    // ---- the chain into D0: d a1 W1, then the gradients of every hoisted per-node projection -------------------------
    tc::gemm_split<D, D>(p, p_hi, p_lo, TCBB_(BT_NB_NN1), D0, D1, false, false, true, true);   // (a1 in D1 is dead)
    uint8_t* q0_hi = P;                              // [dul (K = 64) | dur (K = 64) | dnl_L (K = 128)]
    uint8_t* q0_lo = q0_hi + tc::ROWS * C * 2;
    uint8_t* q1_hi = q0_lo + tc::ROWS * C * 2;
    uint8_t* q1_lo = q1_hi + tc::ROWS * C * 2;
    uint8_t* q2_hi = q1_lo + tc::ROWS * C * 2;
    uint8_t* q2_lo = q2_hi + tc::ROWS * 128 * 2;
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      fold();
      stage_cols<C, 32>(q0_hi, q0_lo, row, half * 32, valid ? a.dul + nn * C + half * 32 : nullptr);
      stage_cols<C, 32>(q1_hi, q1_lo, row, half * 32, valid ? a.dur + nn * C + half * 32 : nullptr);
      stage_cols<128, 64>(q2_hi, q2_lo, row, half * 64, valid ? a.dnl + nn * 128 + half * 64 : nullptr);
      tc::rows_publish(p);
    }
    tc::gemm_split<C, D>(p, q0_hi, q0_lo, TCBB_(BT_EB_NFL), D0, D1, false, false, true, false);
    tc::gemm_split<C, D>(p, q1_hi, q1_lo, TCBB_(BT_EB_NFR), D0, D1, true, true, false, false);
    tc::gemm_split<128, D>(p, q2_hi, q2_lo, TCBB_(BT_EL_NL), D0, D1, true, true, false, true);
    uint8_t* r0_hi = P;                              // [dnl_R (K = 128) | dgn_L (K = 32) | dgn_R (K = 32)]
    uint8_t* r0_lo = r0_hi + tc::ROWS * 128 * 2;
    uint8_t* r1_hi = r0_lo + tc::ROWS * 128 * 2;
    uint8_t* r1_lo = r1_hi + tc::ROWS * 32 * 2;
    uint8_t* r2_hi = r1_lo + tc::ROWS * 32 * 2;
    uint8_t* r2_lo = r2_hi + tc::ROWS * 32 * 2;
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      fold();
      stage_cols<128, 64>(r0_hi, r0_lo, row, half * 64, valid ? a.dnl + (size_t)a.n_nodes * 128 + nn * 128 + half * 64 : nullptr);
      stage_cols<32, 16>(r1_hi, r1_lo, row, half * 16, valid ? a.dgn + nn * 32 + half * 16 : nullptr);
      stage_cols<32, 16>(r2_hi, r2_lo, row, half * 16, valid ? a.dgn + (size_t)a.n_nodes * 32 + nn * 32 + half * 16 : nullptr);
      tc::rows_publish(p);
    }
    tc::gemm_split<128, D>(p, r0_hi, r0_lo, TCBB_(BT_ER_NL), D0, D1, false, false, true, false);
    tc::gemm_split<32, D>(p, r1_hi, r1_lo, TCBB_(BT_EL_GN), D0, D1, true, true, false, false);
    tc::gemm_split<32, D>(p, r2_hi, r2_lo, TCBB_(BT_ER_GN), D0, D1, true, true, false, true);
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      fold();
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float v[16];
        if (a.red_blocked) {
          ld_blocked16(a.dgx, n, hc + c * 16, valid, v);
        } else {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) x = *reinterpret_cast<const float4*>(a.dgx + nn * D + hc + c * 16 + i);
            v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
          }
        }
        store_a16(p_hi, p_lo, row, hc + c * 16, v);
      }
      tc::rows_publish(p);
    }
    tc::gemm_split<D, D>(p, p_hi, p_lo, TCBB_(BT_NB_GX), D0, D1, false, false, true, true);
  } else {
    // first launch of a backward: dx = ddect W_dec1n  (nothing has reached the node path yet)
    uint8_t* s_hi = P;
    uint8_t* s_lo = s_hi + tc::ROWS * C * 2;
    if (p.role == 0) {
      stage_cols<C, 32>(s_hi, s_lo, row, half * 32, (valid && a.do_final) ? a.ddect + nn * C + half * 32 : nullptr);
      tc::rows_publish(p);
    }
    tc::gemm_split<C, D>(p, s_hi, s_lo, a.tc_blob + a.hd.o[MDB_TH_BT_EDEC1N], D0, D1, false, false, true, true);
  }

  if (a.do_A) {
    // ---- NodeBlock node tail backward ---------------------------------------------------------------------------------
    float mean = 0.f, rstd = 1.f;
    if (p.role == 0) {
      tc::rows_wait_acc(p);                          // D0 = everything linear so far ; the planes are free
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {                  // dx += D0 + D1 ; dx -> planes
        float v[16], d[16], x[16];
        tc::tmem_ld16(T0 + c * 16, d);
        tc::tmem_ld16(T1 + c * 16, x);
        ld_blocked16(a.dx, n, hc + c * 16, valid && have_dx, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += d[i] + x[i];
        if (valid) st_blocked16(a.dx, n, hc + c * 16, v);
        store_a16(p_hi, p_lo, row, hc + c * 16, v);
      }
      have_dx = true;
      tc::rows_publish(p);
      // statistics of u = cen + agg while the GEMM below runs
      RunStat rs = {0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float u[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
          if (valid) {
            x = *reinterpret_cast<const float4*>(a.cenA + nn * D + hc + c * 16 + i);
            y = *reinterpret_cast<const float4*>(a.aggA + nn * D + hc + c * 16 + i);
          }
          u[i] = x.x + y.x; u[i + 1] = x.y + y.y; u[i + 2] = x.z + y.z; u[i + 3] = x.w + y.w;
        }
        stat_add16(rs, u);
      }
      const float2 o = exchange_half(stat, row, half, rs.mean, rs.m2);
      mean = 0.5f * (rs.mean + o.x);
      const float dm = rs.mean - o.x;
      rstd = 1.f / sqrtf((rs.m2 + o.y + dm * dm * 64.f) * (1.f / 256.f) + LN_EPS);
    }
    tc::gemm<D, D>(p, p_hi, p_lo, TCBA_(BT_NB_OUT), D1, false, true, true);           // d relu -> D1
    if (p.role == 0) {
      tc::rows_wait_acc(p);                          // (dx_new planes consumed)
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {                  // pass 1: xhat -> planes (parked), d xhat -> D1, row sums
        float u[16], d[16], ga[16], be[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
          if (valid) {
            x = *reinterpret_cast<const float4*>(a.cenA + nn * D + hc + c * 16 + i);
            y = *reinterpret_cast<const float4*>(a.aggA + nn * D + hc + c * 16 + i);
          }
          u[i] = x.x + y.x; u[i + 1] = x.y + y.y; u[i + 2] = x.z + y.z; u[i + 3] = x.w + y.w;
        }
        tc::tmem_ld16(T1 + c * 16, d);
        lds16(vv.ln_g + hc + c * 16, ga); lds16(vv.ln_be + hc + c * 16, be);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float xh = (u[i] - mean) * rstd;
          const float dxh = (xh * ga[i] + be[i] > 0.f) ? d[i] * ga[i] : 0.f;
          u[i] = xh;
          d[i] = dxh;
          s1 += dxh;
          s2 = fmaf(dxh, xh, s2);
        }
        tc::tmem_st16(T1 + c * 16, d);
        store_a16(p_hi, p_lo, row, hc + c * 16, u);
      }
      tc::tmem_st_wait();
      const float2 o2 = exchange_half(stat, row, half, s1, s2);
      const float m1 = (s1 + o2.x) * (1.f / 256.f), m2 = (s2 + o2.y) * (1.f / 256.f);
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {                  // pass 2: dc -> dagg and, in place of xhat, the planes
        float xh[16], d[16];
        load_a16(p_hi, p_lo, row, hc + c * 16, xh);
        tc::tmem_ld16(T1 + c * 16, d);
#pragma unroll
        for (int i = 0; i < 16; ++i) d[i] = valid ? rstd * (d[i] - m1 - xh[i] * m2) : 0.f;
        if (valid) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(a.dagg + nn * D + hc + c * 16 + i) = make_float4(d[i], d[i + 1], d[i + 2], d[i + 3]);
        }
        store_a16(p_hi, p_lo, row, hc + c * 16, d);
      }
      tc::rows_publish(p);
    }
    tc::gemm_split<D, D>(p, p_hi, p_lo, TCBA_(BT_NB_CEN), D0, D1, false, false, true, true);   // dc W_cen (d xhat in D1 is dead)
  }

  if (p.role == 0) {
    tc::rows_wait_acc(p);
    fold();                                          // dx_out
    if (a.do_A) {
      // clear the scatter accumulators that the edge kernels of block i add into (this tile's rows; every read is done)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int rt = tid;                            // 0 .. 255
      const int rows_here = min(tc::ROWS, a.n_nodes - row0);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = rt; i < rows_here * (C / 4); i += 256) {
        reinterpret_cast<float4*>(a.dul + (size_t)row0 * C)[i] = z;
        reinterpret_cast<float4*>(a.dur + (size_t)row0 * C)[i] = z;
      }
      for (int i = rt; i < rows_here * (128 / 4); i += 256) {
        reinterpret_cast<float4*>(a.dnl + (size_t)row0 * 128)[i] = z;
        reinterpret_cast<float4*>(a.dnl + (size_t)a.n_nodes * 128 + (size_t)row0 * 128)[i] = z;
      }
      for (int i = rt; i < rows_here * (32 / 4); i += 256) {
        reinterpret_cast<float4*>(a.dgn + (size_t)row0 * 32)[i] = z;
        reinterpret_cast<float4*>(a.dgn + (size_t)a.n_nodes * 32 + (size_t)row0 * 32)[i] = z;
      }
      // dgx / dhn: node-blocked tables are padded to 64 nodes (whole 32-node blocks are contiguous); row-major ones to n_nodes
      const int64_t lim = a.red_blocked ? ((int64_t)a.n_nodes + 63) / 64 * 64 : (int64_t)a.n_nodes;
      const int rows_z = (int)min((int64_t)tc::ROWS, lim - row0);
      for (int i = rt; i < rows_z * (D / 4); i += 256) {
        reinterpret_cast<float4*>(a.dgx + (size_t)row0 * D)[i] = z;
        reinterpret_cast<float4*>(a.dhn + (size_t)row0 * D)[i] = z;
      }
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == TC_NRW) { __syncwarp(); tc::tmem_dealloc<512>(ps->tmem_base); }
}

inline void fill_bwd_node_vecs(BwdNodeVecs& v, const float* bh, const BlkOff* blkB, const BlkOff* blkA) {
  auto cp = [&](float* dst, int off, int n) { memcpy(dst, bh + off, n * sizeof(float)); };
  if (blkB) { cp(v.nn1_b, blkB->o[MDB_S_NB_NN1_B], D); cp(v.nn1_g, blkB->o[MDB_S_NB_NN1_G], D); cp(v.nn1_be, blkB->o[MDB_S_NB_NN1_BE], D); }
  if (blkA) { cp(v.ln_g, blkA->o[MDB_S_NB_LN_G], D); cp(v.ln_be, blkA->o[MDB_S_NB_LN_BE], D); }
}
