// tc_bondffn_bwd2_kernel: backward of the two BondFFNs of an EdgeBlock + the edge_embs backward that closes a block -- the
// arithmetic and TMEM map of tc_bondffn_bwd_kernel (tc_bondffn.cuh; forward: reference models/graph.py:133-141, 278-284,
// 354-357), restructured after its ncu source view (profiles/r01_*: 12.5 k SASS instructions executed once per tile, 10 %
// of the issue slots stalled on instruction fetch, 13 % on sixteen serialised table loads inside the last epilogue, 144 B of
// spills, and -- for the left FFN, whose node is the CSR-sorted left endpoint -- 16-byte REDs from up to 32 lanes of a warp
// onto the SAME address):
//   * every epilogue is a `#pragma unroll 1` loop over 16-column chunks that re-reads TMEM; gathered table rows are
//     software-prefetched one chunk ahead; per-column vectors sit in shared memory (warp-uniform LDS.128);
//   * the sigmoid of the gate is re-evaluated from the parked gate accumulator instead of living in 32 registers;
//   * the left FFN's d node_linear / d gate-node gradients are reduced over the CSR runs of the tile in shared memory
//     (one atomic per (node, channel) per tile, as the forward kernels do) instead of same-address REDs;
//   * the weight ring carries several K stages per slot (packed slots, tc_pipe.cuh); a 32 KB fp32 tile next to it serves the
//     run reduction.
// MDB_TC_FFN_BWD2=0 selects the old kernel (A/B runs).
//
// Included by mdb_forward.cu inside its anonymous namespace (after tc_edge_tail_bwd.cuh: tail_tile_off).
#pragma once
#include "tc_pipe.cuh"

constexpr uint32_t FB2_SLOT = 2 * tc::WStage<128, tc::KB>::STAGE_BYTES;              // 16 KB: two N = 128 stages, four N = 64, eight N = 32 per slot (tc_pipe.cuh: packed slots)
constexpr size_t FB2_OFF_A = 2 * (size_t)tc::ROWS * C * 2;                            // after the E planes (32 KB)
constexpr size_t FB2_OFF_G = FB2_OFF_A + 2 * (size_t)tc::ROWS * 128 * 2;              // after the A planes (64 KB)
constexpr size_t FB2_OFF_RING = FB2_OFF_G + 2 * (size_t)tc::ROWS * 32 * 2;            // after the gate planes (16 KB)
constexpr size_t FB2_OFF_RT = FB2_OFF_RING + tc::NSTAGE * FB2_SLOT;                   // run-reduction tile [128][64] fp32
constexpr size_t FB2_OFF_PS = FB2_OFF_RT + (size_t)tc::ROWS * C * 4;
constexpr size_t FB2_OFF_STAT = FB2_OFF_PS + 128;
constexpr size_t FB2_OFF_LS = FB2_OFF_STAT + 2 * tc::ROWS * sizeof(float2);
constexpr size_t FB2_OFF_VEC = FB2_OFF_LS + tc::ROWS * sizeof(int);
constexpr size_t SMEM_TC_FFN_BWD2 = FB2_OFF_VEC + sizeof(FfnVecs);
static_assert(SMEM_TC_FFN_BWD2 <= 232448, "tc_bondffn_bwd2_kernel exceeds the shared-memory limit");
static_assert(FB2_OFF_VEC % 16 == 0, "vector block must be 16-byte aligned");

// [128][64] fp32 tile, 16-byte chunks XOR-swizzled by row (tc_edge_tail_bwd.cuh: tail_tile_off): NC columns of this thread's
// row are written at tile columns [c0, c0 + NC); then every CSR run of equal `ls` is summed per channel and added to
// table[node * ld + col0 + channel] with one atomic.  All 256 row threads call; nch = number of tile columns in use.
template <int NCH>
__device__ __forceinline__ void run_reduce_tile(float* rt, const int* ls, float* __restrict__ table, int ld, int col0, int tid) {
  asm volatile("bar.sync 1, 256;" ::: "memory");
  constexpr int NQ = 256 / NCH;                // row groups
  constexpr int RPQ = tc::ROWS / NQ;           // rows per group
  const int c = tid % NCH, r0 = (tid / NCH) * RPQ;
  int cur = ls[r0];
  float acc = 0.f;
  for (int r = r0; r < r0 + RPQ; ++r) {
    const int n = ls[r];
    if (n != cur) {
      if (cur >= 0) atomicAdd(table + (size_t)cur * ld + col0 + c, acc);
      cur = n; acc = 0.f;
    }
    acc += rt[tail_tile_off(r, c)];
  }
  if (cur >= 0) atomicAdd(table + (size_t)cur * ld + col0 + c, acc);
  asm volatile("bar.sync 1, 256;" ::: "memory");   // the tile may be rewritten right after
}

__global__ void __launch_bounds__(TC_NB_THREADS, 1) tc_bondffn_bwd2_kernel(const __grid_constant__ TcFfnBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* e_hi = smem_raw;                                  // E planes, K = 64
  uint8_t* e_lo = e_hi + tc::ROWS * C * 2;
  uint8_t* a_hi = smem_raw + FB2_OFF_A;                      // A planes, K = 128 (64 KB) ...
  uint8_t* a_lo = a_hi + tc::ROWS * 128 * 2;
  uint8_t* s1_hi = a_hi;                                     // ... or two K = 64 plane pairs (2 x 32 KB) aliased on them
  uint8_t* s1_lo = s1_hi + tc::ROWS * C * 2;
  uint8_t* s2_hi = s1_lo + tc::ROWS * C * 2;
  uint8_t* s2_lo = s2_hi + tc::ROWS * C * 2;
  uint8_t* g_hi = smem_raw + FB2_OFF_G;                      // gate planes, K = 32
  uint8_t* g_lo = g_hi + tc::ROWS * 32 * 2;
  uint8_t* stages = smem_raw + FB2_OFF_RING;
  float* rt = reinterpret_cast<float*>(smem_raw + FB2_OFF_RT);
  tc::PipeSmem* ps = reinterpret_cast<tc::PipeSmem*>(smem_raw + FB2_OFF_PS);
  float2* stat = reinterpret_cast<float2*>(smem_raw + FB2_OFF_STAT);
  int* ls = reinterpret_cast<int*>(smem_raw + FB2_OFF_LS);
  float* vecs = reinterpret_cast<float*>(smem_raw + FB2_OFF_VEC);
  const FfnVecs& vv = *reinterpret_cast<const FfnVecs*>(vecs);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * tc::ROWS;
  const Tables& tb = a.tb;
  tc::Pipe p;
  tc::pipe_init<TC_NRW>(p, ps, stages);
  p.slot_bytes = FB2_SLOT;
  if (warp == TC_NRW) tc::tmem_alloc<512>(&ps->tmem_base);
  {
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
  // TMEM column map (as tc_bondffn_bwd_kernel)
  constexpr uint32_t T_BL = 0, T_A5 = 128, T_I2 = 256, T_G1 = 320, T_GG = 352, T_DR6 = 416, T_DE = 448, T_WIDE = 256;

  if (p.role == 0) {   // e tile -> E planes
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) x = *reinterpret_cast<const float4*>(a.e + (size_t)q * C + half * 32 + c * 16 + i);
        v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
      }
      tc::store_a<C, 16>(e_hi, e_lo, row, half * 32 + c * 16, v);
    }
    tc::rows_publish(p);
  }
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    const FfnSideVecs& sv = vv.side[side];
    const int node = side ? rr : ll;                         // node feeding this FFN
    const int other = side ? ll : rr;                        // node its output was scattered to
    const float* nl = (side ? tb.nlr : tb.nll) + (size_t)node * 128 + half * 64;
    const float* gn = (side ? tb.gnr : tb.gnl) + (size_t)node * 32 + half * 16;
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
    if (p.role == 0) {
      Row16 nx = ld_row16(nl), gq = ld_row16(gn);            // requested before the accumulator wait
      tc::rows_wait_acc(p);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {                          // inter0 = bond_linear(e) * node_linear(h_node)[.]
        float x[16], h[16];
        unpack_row16(nx, h);
        if (c < 3) nx = ld_row16(nl + (c + 1) * 16);
        tc::tmem_ld16(lane_base + T_BL + half * 64 + c * 16, x);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] *= h[i];
        tc::store_a<128, 16>(a_hi, a_lo, row, half * 64 + c * 16, x);
      }
      {
        float gv[16], w[16], ga[16], be[16], r6[16];
        unpack_row16(gq, gv);
        tc::tmem_ld16(lane_base + T_G1 + half * 16, xh6);
        lds16(sv.gt_w + half * 16, w);
#pragma unroll
        for (int i = 0; i < 16; ++i) xh6[i] += gv[i] + te * w[i];
        rstd6 = ln_xhat_part<16>(xh6, stat, row, half);
        lds16(sv.g1_g + half * 16, ga); lds16(sv.g1_be + half * 16, be);
#pragma unroll
        for (int i = 0; i < 16; ++i) r6[i] = fmaxf(xh6[i] * ga[i] + be[i], 0.f);
        tc::store_a<32, 16>(g_hi, g_lo, row, half * 16, r6);
      }
      tc::rows_publish(p);
    }
    tc::gemm<128, 128>(p, a_hi, a_lo, w_i1, T_A5, false, true, false);
    tc::gemm<32, C>(p, g_hi, g_lo, w_g2, T_GG, false, false, true);
    float mean5 = 0.f, rstd5 = 0.f;
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      RunStat rs = {0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float x[16], b[16];
        tc::tmem_ld16(lane_base + T_A5 + half * 64 + c * 16, x);
        lds16(sv.i1_b + half * 64 + c * 16, b);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] += b[i];
        stat_add16(rs, x);
      }
      const float2 o = exchange_half(stat, row, half, rs.mean, rs.m2);
      mean5 = 0.5f * (rs.mean + o.x);
      const float dm = rs.mean - o.x;
      rstd5 = 1.f / sqrtf((rs.m2 + o.y + dm * dm * 32.f) * (1.f / 128.f) + LN_EPS);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {                          // r5 = relu(LN(a5)) -> A planes
        float x[16], b[16], ga[16], be[16];
        tc::tmem_ld16(lane_base + T_A5 + half * 64 + c * 16, x);
        lds16(sv.i1_b + half * 64 + c * 16, b); lds16(sv.i1_g + half * 64 + c * 16, ga); lds16(sv.i1_be + half * 64 + c * 16, be);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaxf((x[i] + b[i] - mean5) * rstd5 * ga[i] + be[i], 0.f);
        tc::store_a<128, 16>(a_hi, a_lo, row, half * 64 + c * 16, x);
      }
      tc::rows_publish(p);
    }
    tc::gemm<128, C>(p, a_hi, a_lo, w_i2, T_I2, false, true, true);
    // ---- d o = DUL[r] (left FFN, scattered over right) / DUR[l] (right FFN, scattered over left)
    if (p.role == 0) {
      const float* du = (side ? a.dur : a.dul) + (size_t)other * C + half * 32;
      Row16 d0 = ld_row16(du), d1 = ld_row16(du + 16);       // (`other` is 0 for padding rows: a valid address)
      tc::rows_wait_acc(p);
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        float i2[16], gg[16], dd[16], b[16], bg[16];
        unpack_row16(c ? d1 : d0, dd);
        tc::tmem_ld16(lane_base + T_I2 + half * 32 + c * 16, i2);
        tc::tmem_ld16(lane_base + T_GG + half * 32 + c * 16, gg);
        lds16(sv.i2_b + half * 32 + c * 16, b); lds16(sv.g2_b + half * 32 + c * 16, bg);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float d = valid ? dd[i] : 0.f;
          const float s = tc::fast_sigmoid(gg[i] + bg[i]);
          const float val = i2[i] + b[i];
          gg[i] = d * val * s * (1.f - s);                    // d gate logit
          i2[i] = d * s;                                      // d i2
        }
        tc::store_a<C, 16>(s1_hi, s1_lo, row, half * 32 + c * 16, gg);   // (the K = 128 A planes are dead: r5 was consumed)
        tc::store_a<C, 16>(s2_hi, s2_lo, row, half * 32 + c * 16, i2);
      }
      tc::rows_publish(p);
    }
    tc::gemm<C, 32>(p, s1_hi, s1_lo, t_g2, T_DR6, false, true, false);        // d relu6
    tc::gemm<C, 128>(p, s2_hi, s2_lo, t_i2, T_WIDE, false, false, true);      // d relu5
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      {  // gate branch
        float d6[16], ga[16], be[16];
        tc::tmem_ld16(lane_base + T_DR6 + half * 16, d6);
        lds16(sv.g1_g + half * 16, ga); lds16(sv.g1_be + half * 16, be);
        ln_bwd_part<16>(d6, xh6, rstd6, ga, be, stat, row, half);
        tc::store_a<32, 16>(g_hi, g_lo, row, half * 16, d6);
        if (side == 0) {     // node = left endpoint: equal for (almost) every lane of the warp -> reduce the CSR runs first
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(rt + tail_tile_off(row, half * 16 + i)) =
                valid ? make_float4(d6[i], d6[i + 1], d6[i + 2], d6[i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
          run_reduce_tile<32>(rt, ls, dgn_tab, 32, 0, tid);
        } else if (valid) {
          float* dgn = dgn_tab + (size_t)node * 32 + half * 16;
#pragma unroll
          for (int i = 0; i < 16; i += 4) tc::red_add_v4(dgn + i, d6[i], d6[i + 1], d6[i + 2], d6[i + 3]);
        }
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {                          // inter branch, pass 1: d xhat5 -> T_WIDE, row sums
        float x[16], d[16], b[16], ga[16], be[16];
        tc::tmem_ld16(lane_base + T_A5 + half * 64 + c * 16, x);
        tc::tmem_ld16(lane_base + T_WIDE + half * 64 + c * 16, d);
        lds16(sv.i1_b + half * 64 + c * 16, b); lds16(sv.i1_g + half * 64 + c * 16, ga); lds16(sv.i1_be + half * 64 + c * 16, be);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float xh = (x[i] + b[i] - mean5) * rstd5;
          const float dxh = (xh * ga[i] + be[i] > 0.f) ? d[i] * ga[i] : 0.f;
          d[i] = dxh;
          s1 += dxh;
          s2 = fmaf(dxh, xh, s2);
        }
        tc::tmem_st16(lane_base + T_WIDE + half * 64 + c * 16, d);
      }
      tc::tmem_st_wait();
      const float2 o2 = exchange_half(stat, row, half, s1, s2);
      const float m1 = (s1 + o2.x) * (1.f / 128.f), m2 = (s2 + o2.y) * (1.f / 128.f);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {                          // pass 2: d a5 -> A planes (the K = 64 planes aliased here are dead)
        float x[16], d[16], b[16];
        tc::tmem_ld16(lane_base + T_A5 + half * 64 + c * 16, x);
        tc::tmem_ld16(lane_base + T_WIDE + half * 64 + c * 16, d);
        lds16(sv.i1_b + half * 64 + c * 16, b);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float xh = (x[i] + b[i] - mean5) * rstd5;
          d[i] = rstd5 * (d[i] - m1 - xh * m2);
        }
        tc::store_a<128, 16>(a_hi, a_lo, row, half * 64 + c * 16, d);
      }
      tc::rows_publish(p);
    }
    tc::gemm<32, C>(p, g_hi, g_lo, t_gb, T_DE, side != 0, true, false);        // de (+)= d a6 W_gb^T
    tc::gemm<128, 128>(p, a_hi, a_lo, t_i1, T_WIDE, false, false, true);      // d inter0
    if (p.role == 0) {
      Row16 nx = ld_row16(nl);
      tc::rows_wait_acc(p);
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {                 // two 32-column passes (the run-reduction tile is 64 wide)
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c = pass * 2 + cc;
          float di[16], bl[16], h[16];
          unpack_row16(nx, h);
          if (c < 3) nx = ld_row16(nl + (c + 1) * 16);
          tc::tmem_ld16(lane_base + T_WIDE + half * 64 + c * 16, di);
          tc::tmem_ld16(lane_base + T_BL + half * 64 + c * 16, bl);
#pragma unroll
          for (int i = 0; i < 16; ++i) bl[i] *= di[i];        // d node_linear(h_node)[node]
          if (side == 0) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4*>(rt + tail_tile_off(row, half * 32 + cc * 16 + i)) =
                  valid ? make_float4(bl[i], bl[i + 1], bl[i + 2], bl[i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
          } else if (valid) {
            float* dnl = dnl_tab + (size_t)node * 128 + half * 64 + c * 16;
#pragma unroll
            for (int i = 0; i < 16; i += 4) tc::red_add_v4(dnl + i, bl[i], bl[i + 1], bl[i + 2], bl[i + 3]);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) di[i] *= h[i];         // d bond_linear(e)
          tc::store_a<128, 16>(a_hi, a_lo, row, half * 64 + c * 16, di);
        }
        if (side == 0) {
          // tile column j of this pass = dnl column (j / 32) * 64 + pass * 32 + j % 32: two 32-wide column groups
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const int ch = tid & 63, r0 = (tid >> 6) * 32;
          const int col = (ch >> 5) * 64 + pass * 32 + (ch & 31);
          int cur = ls[r0];
          float acc = 0.f;
          for (int r = r0; r < r0 + 32; ++r) {
            const int n = ls[r];
            if (n != cur) {
              if (cur >= 0) atomicAdd(dnl_tab + (size_t)cur * 128 + col, acc);
              cur = n; acc = 0.f;
            }
            acc += rt[tail_tile_off(r, ch)];
          }
          if (cur >= 0) atomicAdd(dnl_tab + (size_t)cur * 128 + col, acc);
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
      tc::rows_publish(p);
    }
    tc::gemm<128, C>(p, a_hi, a_lo, t_bl, T_DE, true, true, side == 1);        // de += d bl W_bl^T
  }
  // ---- total d e -> d h_in = de W_EE[:, :64]^T..., d rbf += de W_EE[:, 64:80]
  if (p.role == 0) {
    tc::rows_wait_acc(p);
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      float de[16];
      tc::tmem_ld16(lane_base + T_DE + half * 32 + c * 16, de);
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) x = *reinterpret_cast<const float4*>(a.de_in + (size_t)q * C + half * 32 + c * 16 + i);
        de[i] += x.x; de[i + 1] += x.y; de[i + 2] += x.z; de[i + 3] += x.w;
      }
      tc::store_a<C, 16>(s1_hi, s1_lo, row, half * 32 + c * 16, de);
    }
    tc::rows_publish(p);
  }
  tc::gemm<C, C>(p, s1_hi, s1_lo, TCW_(BT_EEH), 0, false, true, false);
  tc::gemm<C, 32>(p, s1_hi, s1_lo, TCW_(BT_EEG), 64, false, false, true);
  if (p.role == 0) {
    tc::rows_wait_acc(p);
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      float v[16];
      tc::tmem_ld16(lane_base + half * 32 + c * 16, v);
      if (valid) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(a.dh + (size_t)q * C + half * 32 + c * 16 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
    if (half == 0) {
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
  __syncthreads();
  if (warp == TC_NRW) { __syncwarp(); tc::tmem_dealloc<512>(ps->tmem_base); }
}
