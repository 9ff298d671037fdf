// tc_bondffn_bwd3_kernel: tc_bondffn_bwd2_kernel (same arithmetic, TMEM map and GEMM sequence: backward of the two BondFFNs
// of an EdgeBlock + the edge_embs backward; forward: reference models/graph.py:133-141, 278-284, 354-357) with SIXTEEN row
// warps -- four threads per tile row, each owning a quarter of every activation row (32 / 16 columns of the 128- / 64-wide
// tiles; the 32-wide gate tiles stay with parts 0 and 1) -- and the register split of the 16-warp NodeBlock kernels
// (setmaxnreg: 112 for the four row warpgroups, 32 for the producer / MMA warpgroup).
//
// Why: the kernel executes ~120 k warp instructions per tile at IPC ~1 with 8 row warps (ncu: issue slots 30 %, tensor pipe
// 13 %, no unit above 50 %): every epilogue is a chain of dependent TMEM reads, table gathers and barriers per THREAD, and
// TMEM (512 columns parked across the tile) rules out a second resident CTA.  Halving the columns per thread halves every
// chain and doubles the warps that can hide it.
// MDB_TC_FFN_BWD3=0 selects tc_bondffn_bwd2_kernel (A/B runs).
//
// Included by mdb_forward.cu inside its anonymous namespace (after tc_bondffn_bwd2.cuh).
#pragma once
#include "tc_pipe.cuh"

constexpr size_t FB3_OFF_A = 2 * (size_t)tc::ROWS * C * 2;                            // after the E planes (32 KB)
constexpr size_t FB3_OFF_G = FB3_OFF_A + 2 * (size_t)tc::ROWS * 128 * 2;              // after the A planes (64 KB)
constexpr size_t FB3_OFF_RING = FB3_OFF_G + 2 * (size_t)tc::ROWS * 32 * 2;            // after the gate planes (16 KB)
constexpr size_t FB3_OFF_RT = FB3_OFF_RING + tc::NSTAGE * tc::STAGE_SLOT;             // run-reduction tile [128][64] fp32
constexpr size_t FB3_OFF_PS = FB3_OFF_RT + (size_t)tc::ROWS * C * 4;
constexpr size_t FB3_OFF_STAT = FB3_OFF_PS + 128;                                     // [4][128] float2
constexpr size_t FB3_OFF_LS = FB3_OFF_STAT + 4 * tc::ROWS * sizeof(float2);
constexpr size_t FB3_OFF_VEC = FB3_OFF_LS + tc::ROWS * sizeof(int);
constexpr size_t SMEM_TC_FFN_BWD3 = FB3_OFF_VEC + sizeof(FfnVecs);
static_assert(SMEM_TC_FFN_BWD3 <= 232448, "tc_bondffn_bwd3_kernel exceeds the shared-memory limit");
static_assert(FB3_OFF_VEC % 16 == 0, "vector block must be 16-byte aligned");

// sum of (x, y) over the four parts of a row (all 512 row threads call)
__device__ __forceinline__ float2 sum4_parts(float2* stat, int row, int part, float x, float y) {
  stat[part * tc::ROWS + row] = make_float2(x, y);
  asm volatile("bar.sync 1, 512;" ::: "memory");
  const float2 s0 = stat[row], s1 = stat[tc::ROWS + row], s2 = stat[2 * tc::ROWS + row], s3 = stat[3 * tc::ROWS + row];
  asm volatile("bar.sync 1, 512;" ::: "memory");
  return make_float2((s0.x + s1.x) + (s2.x + s3.x), (s0.y + s1.y) + (s2.y + s3.y));
}
// partner exchange between parts 0 and 1 only (the 32-wide gate tiles; warps 0..7 = 256 threads call)
__device__ __forceinline__ float2 exchange_gate(float2* stat, int row, int part, float x, float y) {
  stat[part * tc::ROWS + row] = make_float2(x, y);
  asm volatile("bar.sync 2, 256;" ::: "memory");
  const float2 o = stat[(part ^ 1) * tc::ROWS + row];
  asm volatile("bar.sync 2, 256;" ::: "memory");
  return o;
}

template <bool IS_ROW>
__device__ __forceinline__ void tc_bondffn_bwd3_body(const TcFfnBwdArgs& a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* e_hi = smem_raw;                                  // E planes, K = 64
  uint8_t* e_lo = e_hi + tc::ROWS * C * 2;
  uint8_t* a_hi = smem_raw + FB3_OFF_A;                      // A planes, K = 128 (64 KB) ...
  uint8_t* a_lo = a_hi + tc::ROWS * 128 * 2;
  uint8_t* s1_hi = a_hi;                                     // ... or two K = 64 plane pairs (2 x 32 KB) aliased on them
  uint8_t* s1_lo = s1_hi + tc::ROWS * C * 2;
  uint8_t* s2_hi = s1_lo + tc::ROWS * C * 2;
  uint8_t* s2_lo = s2_hi + tc::ROWS * C * 2;
  uint8_t* g_hi = smem_raw + FB3_OFF_G;                      // gate planes, K = 32
  uint8_t* g_lo = g_hi + tc::ROWS * 32 * 2;
  uint8_t* stages = smem_raw + FB3_OFF_RING;
  float* rt = reinterpret_cast<float*>(smem_raw + FB3_OFF_RT);
  tc::PipeSmem* ps = reinterpret_cast<tc::PipeSmem*>(smem_raw + FB3_OFF_PS);
  float2* stat = reinterpret_cast<float2*>(smem_raw + FB3_OFF_STAT);
  int* ls = reinterpret_cast<int*>(smem_raw + FB3_OFF_LS);
  float* vecs = reinterpret_cast<float*>(smem_raw + FB3_OFF_VEC);
  const FfnVecs& vv = *reinterpret_cast<const FfnVecs*>(vecs);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * tc::ROWS;
  const Tables& tb = a.tb;
  tc::Pipe p;
  tc::pipe_init_split<NB16_NRW, IS_ROW, tc::NSTAGE>(p, ps, stages);
  if (warp == NB16_NRW) tc::tmem_alloc<512>(&ps->tmem_base);
  if (IS_ROW) {
    constexpr int NV = (int)(sizeof(FfnVecs) / sizeof(float)), NIT = (NV + 511) / 512;
    float t[NIT];
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int i = tid + k * 512;
      t[k] = reinterpret_cast<const float*>(&a.v)[i < NV ? i : 0];
    }
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int i = tid + k * 512;
      if (i < NV) vecs[i] = t[k];
    }
  }
  const int row = (warp & 3) * 32 + lane;
  const int part = (warp >> 2) & 3;                          // quarter of the row this thread owns
  const bool gate_thr = IS_ROW && part < 2;                  // the 32-wide gate tiles: parts 0 and 1, 16 columns each
  const int q = q0 + row;
  const bool valid = IS_ROW && q < a.n_edges;
  const int ll = valid ? a.left[q] : 0, rr = valid ? a.right[q] : 0;
  const float te = valid ? tb.te[q] : 0.f;
  if (IS_ROW && part == 0) ls[row] = valid ? ll : -1;
  tc::fence_before_sync();
  tc::cta_sync();
  tc::fence_after_sync();
  const uint32_t lane_base = ps->tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  // TMEM column map (as tc_bondffn_bwd_kernel)
  constexpr uint32_t T_BL = 0, T_A5 = 128, T_I2 = 256, T_G1 = 320, T_GG = 352, T_DR6 = 416, T_DE = 448, T_WIDE = 256;
  const int c32 = part * 32, c16 = part * 16;                // this thread's first column in a 128- / 64-wide (and gate) tile

  if (IS_ROW) {   // e tile -> E planes
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) x = *reinterpret_cast<const float4*>(a.e + (size_t)q * C + c16 + i);
      v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
    }
    tc::store_a<C, 16>(e_hi, e_lo, row, c16, v);
    tc::rows_publish(p);
  }
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    const FfnSideVecs& sv = vv.side[side];
    const int node = side ? rr : ll;                         // node feeding this FFN
    const int other = side ? ll : rr;                        // node its output was scattered to
    const float* nl = (side ? tb.nlr : tb.nll) + (size_t)node * 128 + c32;
    const float* gn = (side ? tb.gnr : tb.gnl) + (size_t)node * 32 + c16;
    float* dnl_tab = a.dnl + (size_t)side * a.n_nodes * 128;
    float* dgn_tab = a.dgn + (size_t)side * a.n_nodes * 32;
    const uint8_t* w_bl = side ? TCW_(ER_BL) : TCW_(EL_BL);
    const uint8_t* w_gb = side ? TCW_(ER_GB) : TCW_(EL_GB);
    const uint8_t* w_i1 = side ? TCW_(ER_I1) : TCW_(EL_I1);
    const uint8_t* w_g2 = side ? TCW_(ER_G2) : TCW_(EL_G2);
    const uint8_t* w_i2 = side ? TCW_(ER_I2) : TCW_(EL_I2);
    const uint8_t* t_g2 = side ? TCW_(BT_ER_G2) : TCW_(BT_EL_G2);
    const uint8_t* t_i2 = side ? TCW_(BT_ER_I2) : TCW_(BT_EL_I2);
    const uint8_t* t_gb = side ? TCW_(BT_ER_GB) : TCW_(BT_EL_GB);
    const uint8_t* t_i1 = side ? TCW_(BT_ER_I1) : TCW_(BT_EL_I1);
    const uint8_t* t_bl = side ? TCW_(BT_ER_BL) : TCW_(BT_EL_BL);

    // ---- forward recompute (side 1 chains onto side 0's last GEMM: same E planes, accumulators already drained)
    tc::gemm<C, 128>(p, e_hi, e_lo, w_bl, T_BL, false, side == 0, false);
    tc::gemm<C, 32>(p, e_hi, e_lo, w_gb, T_G1, false, false, true);
    float xh6[16], rstd6 = 0.f;
    if (IS_ROW) {
      Row16 n0 = ld_row16(nl), n1 = ld_row16(nl + 16);       // requested before the accumulator wait
      Row16 gq = n0;
      if (gate_thr) gq = ld_row16(gn);
      tc::rows_wait_acc(p);
#pragma unroll
      for (int c = 0; c < 2; ++c) {                          // inter0 = bond_linear(e) * node_linear(h_node)[.]
        float x[16], h[16];
        unpack_row16(c ? n1 : n0, h);
        tc::tmem_ld16(lane_base + T_BL + c32 + c * 16, x);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] *= h[i];
        tc::store_a<128, 16>(a_hi, a_lo, row, c32 + c * 16, x);
      }
      if (gate_thr) {
        float gv[16], w[16], ga[16], be[16], r6[16];
        unpack_row16(gq, gv);
        tc::tmem_ld16(lane_base + T_G1 + c16, xh6);
        lds16(sv.gt_w + c16, w);
#pragma unroll
        for (int i = 0; i < 16; ++i) xh6[i] += gv[i] + te * w[i];
        {   // LayerNorm statistics of the 32-wide row: two parts of 16
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) s += xh6[i];
          const float m_h = s * (1.f / 16.f);
          float qq = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float d = xh6[i] - m_h; qq = fmaf(d, d, qq); }
          const float2 o = exchange_gate(stat, row, part, m_h, qq);
          const float mean = 0.5f * (m_h + o.x);
          const float dm = m_h - o.x;
          rstd6 = 1.f / sqrtf((qq + o.y + dm * dm * 8.f) * (1.f / 32.f) + LN_EPS);
#pragma unroll
          for (int i = 0; i < 16; ++i) xh6[i] = (xh6[i] - mean) * rstd6;
        }
        lds16(sv.g1_g + c16, ga); lds16(sv.g1_be + c16, be);
#pragma unroll
        for (int i = 0; i < 16; ++i) r6[i] = fmaxf(xh6[i] * ga[i] + be[i], 0.f);
        tc::store_a<32, 16>(g_hi, g_lo, row, c16, r6);
      }
      tc::rows_publish(p);
    }
    tc::gemm<128, 128>(p, a_hi, a_lo, w_i1, T_A5, false, true, false);
    tc::gemm<32, C>(p, g_hi, g_lo, w_g2, T_GG, false, false, true);
    float mean5 = 0.f, rstd5 = 0.f;
    if (IS_ROW) {
      tc::rows_wait_acc(p);
      RunStat rs = {0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float x[16], b[16];
        tc::tmem_ld16(lane_base + T_A5 + c32 + c * 16, x);
        lds16(sv.i1_b + c32 + c * 16, b);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] += b[i];
        stat_add16(rs, x);
      }
      {   // merge the four 32-column parts (equal counts)
        stat[part * tc::ROWS + row] = make_float2(rs.mean, rs.m2);
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const float2 s0 = stat[row], s1 = stat[tc::ROWS + row], s2 = stat[2 * tc::ROWS + row], s3 = stat[3 * tc::ROWS + row];
        asm volatile("bar.sync 1, 512;" ::: "memory");
        mean5 = 0.25f * ((s0.x + s1.x) + (s2.x + s3.x));
        const float d0 = s0.x - mean5, d1 = s1.x - mean5, d2 = s2.x - mean5, d3 = s3.x - mean5;
        const float m2 = (s0.y + s1.y) + (s2.y + s3.y) + 32.f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
        rstd5 = 1.f / sqrtf(m2 * (1.f / 128.f) + LN_EPS);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {                          // r5 = relu(LN(a5)) -> A planes
        float x[16], b[16], ga[16], be[16];
        tc::tmem_ld16(lane_base + T_A5 + c32 + c * 16, x);
        lds16(sv.i1_b + c32 + c * 16, b); lds16(sv.i1_g + c32 + c * 16, ga); lds16(sv.i1_be + c32 + c * 16, be);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaxf((x[i] + b[i] - mean5) * rstd5 * ga[i] + be[i], 0.f);
        tc::store_a<128, 16>(a_hi, a_lo, row, c32 + c * 16, x);
      }
      tc::rows_publish(p);
    }
    tc::gemm<128, C>(p, a_hi, a_lo, w_i2, T_I2, false, true, true);
    // ---- d o = DUL[r] (left FFN, scattered over right) / DUR[l] (right FFN, scattered over left)
    if (IS_ROW) {
      const float* du = (side ? a.dur : a.dul) + (size_t)other * C + c16;
      Row16 d0 = ld_row16(du);                               // (`other` is 0 for padding rows: a valid address)
      tc::rows_wait_acc(p);
      float i2[16], gg[16], dd[16], b[16], bg[16];
      unpack_row16(d0, dd);
      tc::tmem_ld16(lane_base + T_I2 + c16, i2);
      tc::tmem_ld16(lane_base + T_GG + c16, gg);
      lds16(sv.i2_b + c16, b); lds16(sv.g2_b + c16, bg);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float d = valid ? dd[i] : 0.f;
        const float s = tc::fast_sigmoid(gg[i] + bg[i]);
        const float val = i2[i] + b[i];
        gg[i] = d * val * s * (1.f - s);                      // d gate logit
        i2[i] = d * s;                                        // d i2
      }
      tc::store_a<C, 16>(s1_hi, s1_lo, row, c16, gg);         // (the K = 128 A planes are dead: r5 was consumed)
      tc::store_a<C, 16>(s2_hi, s2_lo, row, c16, i2);
      tc::rows_publish(p);
    }
    tc::gemm<C, 32>(p, s1_hi, s1_lo, t_g2, T_DR6, false, true, false);        // d relu6
    tc::gemm<C, 128>(p, s2_hi, s2_lo, t_i2, T_WIDE, false, false, true);      // d relu5
    if (IS_ROW) {
      tc::rows_wait_acc(p);
      if (gate_thr) {  // gate branch (parts 0, 1)
        float d6[16], ga[16], be[16];
        tc::tmem_ld16(lane_base + T_DR6 + c16, d6);
        lds16(sv.g1_g + c16, ga); lds16(sv.g1_be + c16, be);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float dxh = (xh6[i] * ga[i] + be[i] > 0.f) ? d6[i] * ga[i] : 0.f;
          d6[i] = dxh;
          s1 += dxh;
          s2 = fmaf(dxh, xh6[i], s2);
        }
        const float2 o = exchange_gate(stat, row, part, s1, s2);
        const float m1 = (s1 + o.x) * (1.f / 32.f), m2 = (s2 + o.y) * (1.f / 32.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) d6[i] = rstd6 * (d6[i] - m1 - xh6[i] * m2);
        tc::store_a<32, 16>(g_hi, g_lo, row, c16, d6);
        if (side == 0) {     // node = left endpoint: equal for (almost) every lane of the warp -> reduced over CSR runs below
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(rt + tail_tile_off(row, c16 + i)) =
                valid ? make_float4(d6[i], d6[i + 1], d6[i + 2], d6[i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (valid) {
          float* dgn = dgn_tab + (size_t)node * 32 + c16;
#pragma unroll
          for (int i = 0; i < 16; i += 4) tc::red_add_v4(dgn + i, d6[i], d6[i + 1], d6[i + 2], d6[i + 3]);
        }
      }
      if (side == 0) {       // all 512 row threads: 32 channels x 16 row groups of 8 rows
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const int ch = tid & 31, r0 = (tid >> 5) * 8;
        int cur = ls[r0];
        float acc = 0.f;
        for (int r = r0; r < r0 + 8; ++r) {
          const int n = ls[r];
          if (n != cur) {
            if (cur >= 0) atomicAdd(dgn_tab + (size_t)cur * 32 + ch, acc);
            cur = n; acc = 0.f;
          }
          acc += rt[tail_tile_off(r, ch)];
        }
        if (cur >= 0) atomicAdd(dgn_tab + (size_t)cur * 32 + ch, acc);
        asm volatile("bar.sync 1, 512;" ::: "memory");
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {                          // inter branch, pass 1: d xhat5 -> T_WIDE, row sums
        float x[16], d[16], b[16], ga[16], be[16];
        tc::tmem_ld16(lane_base + T_A5 + c32 + c * 16, x);
        tc::tmem_ld16(lane_base + T_WIDE + c32 + c * 16, d);
        lds16(sv.i1_b + c32 + c * 16, b); lds16(sv.i1_g + c32 + c * 16, ga); lds16(sv.i1_be + c32 + c * 16, be);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float xh = (x[i] + b[i] - mean5) * rstd5;
          const float dxh = (xh * ga[i] + be[i] > 0.f) ? d[i] * ga[i] : 0.f;
          d[i] = dxh;
          s1 += dxh;
          s2 = fmaf(dxh, xh, s2);
        }
        tc::tmem_st16(lane_base + T_WIDE + c32 + c * 16, d);
      }
      tc::tmem_st_wait();
      const float2 tot = sum4_parts(stat, row, part, s1, s2);
      const float m1 = tot.x * (1.f / 128.f), m2 = tot.y * (1.f / 128.f);
#pragma unroll
      for (int c = 0; c < 2; ++c) {                          // pass 2: d a5 -> A planes (the K = 64 planes aliased here are dead)
        float x[16], d[16], b[16];
        tc::tmem_ld16(lane_base + T_A5 + c32 + c * 16, x);
        tc::tmem_ld16(lane_base + T_WIDE + c32 + c * 16, d);
        lds16(sv.i1_b + c32 + c * 16, b);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float xh = (x[i] + b[i] - mean5) * rstd5;
          d[i] = rstd5 * (d[i] - m1 - xh * m2);
        }
        tc::store_a<128, 16>(a_hi, a_lo, row, c32 + c * 16, d);
      }
      tc::rows_publish(p);
    }
    tc::gemm<32, C>(p, g_hi, g_lo, t_gb, T_DE, side != 0, true, false);        // de (+)= d a6 W_gb^T
    tc::gemm<128, 128>(p, a_hi, a_lo, t_i1, T_WIDE, false, false, true);      // d inter0
    if (IS_ROW) {
      Row16 n0 = ld_row16(nl), n1 = ld_row16(nl + 16);
      tc::rows_wait_acc(p);
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {                          // (side 0: one 64-column run-reduction pass per chunk)
        float di[16], bl[16], h[16];
        unpack_row16(c ? n1 : n0, h);
        tc::tmem_ld16(lane_base + T_WIDE + c32 + c * 16, di);
        tc::tmem_ld16(lane_base + T_BL + c32 + c * 16, bl);
#pragma unroll
        for (int i = 0; i < 16; ++i) bl[i] *= di[i];          // d node_linear(h_node)[node]
        if (side == 0) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(rt + tail_tile_off(row, c16 + i)) =
                valid ? make_float4(bl[i], bl[i + 1], bl[i + 2], bl[i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (valid) {
          float* dnl = dnl_tab + (size_t)node * 128 + c32 + c * 16;
#pragma unroll
          for (int i = 0; i < 16; i += 4) tc::red_add_v4(dnl + i, bl[i], bl[i + 1], bl[i + 2], bl[i + 3]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) di[i] *= h[i];           // d bond_linear(e)
        tc::store_a<128, 16>(a_hi, a_lo, row, c32 + c * 16, di);
        if (side == 0) {
          // tile column j of this pass = dnl column (j / 16) * 32 + c * 16 + j % 16
          asm volatile("bar.sync 1, 512;" ::: "memory");
          const int ch = tid & 63, r0 = (tid >> 6) * 16;      // 64 channels x 8 row groups of 16 rows
          const int col = (ch >> 4) * 32 + c * 16 + (ch & 15);
          int cur = ls[r0];
          float acc = 0.f;
          for (int r = r0; r < r0 + 16; ++r) {
            const int n = ls[r];
            if (n != cur) {
              if (cur >= 0) atomicAdd(dnl_tab + (size_t)cur * 128 + col, acc);
              cur = n; acc = 0.f;
            }
            acc += rt[tail_tile_off(r, ch)];
          }
          if (cur >= 0) atomicAdd(dnl_tab + (size_t)cur * 128 + col, acc);
          asm volatile("bar.sync 1, 512;" ::: "memory");
        }
      }
      tc::rows_publish(p);
    }
    tc::gemm<128, C>(p, a_hi, a_lo, t_bl, T_DE, true, true, side == 1);        // de += d bl W_bl^T
  }
  // ---- total d e -> d h_in = de W_EE[:, :64]^T..., d rbf += de W_EE[:, 64:80]
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    float de[16];
    tc::tmem_ld16(lane_base + T_DE + c16, de);
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) x = *reinterpret_cast<const float4*>(a.de_in + (size_t)q * C + c16 + i);
      de[i] += x.x; de[i + 1] += x.y; de[i + 2] += x.z; de[i + 3] += x.w;
    }
    tc::store_a<C, 16>(s1_hi, s1_lo, row, c16, de);
    tc::rows_publish(p);
  }
  tc::gemm<C, C>(p, s1_hi, s1_lo, TCW_(BT_EEH), 0, false, true, false);
  tc::gemm<C, 32>(p, s1_hi, s1_lo, TCW_(BT_EEG), 64, false, false, true);
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    float v[16];
    tc::tmem_ld16(lane_base + c16, v);
    if (valid) {
#pragma unroll
      for (int i = 0; i < 16; i += 4)
        *reinterpret_cast<float4*>(a.dh + (size_t)q * C + c16 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    if (part == 0) {
      float g[16];
      tc::tmem_ld16(lane_base + 64, g);
      if (valid) {
        float* dst = a.dg + (size_t)q * G;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 o = *reinterpret_cast<float4*>(dst + i);
          o.x += g[i]; o.y += g[i + 1]; o.z += g[i + 2]; o.w += g[i + 3];
          *reinterpret_cast<float4*>(dst + i) = o;
        }
      }
    }
    tc::fence_before_sync();
  }
  tc::cta_sync();
  if (warp == NB16_NRW) { __syncwarp(); tc::tmem_dealloc<512>(ps->tmem_base); }
}

__global__ void __launch_bounds__(NB16_THREADS, 1) tc_bondffn_bwd3_kernel(const __grid_constant__ TcFfnBwdArgs a) {
  if (threadIdx.x < NB16_NRW * 32) {
    tc::reg_alloc<112>();
    tc_bondffn_bwd3_body<true>(a);
  } else {
    tc::reg_dealloc<32>();
    tc_bondffn_bwd3_body<false>(a);
  }
}
