// tc_nodeblock_bwd16_kernel: input-gradient backward of the NodeBlock per-edge path (same math as tc_nodeblock_bwd_kernel /
// bwd_edge_nodeblock_kernel; forward: reference models/graph.py:42-50) with 16 row warps and rolled epilogues.
//
//   forward (recomputed):  a2 = e W1 + b1 ; r2 = relu(LN(a2)) ; he = r2 W2 + b2 ; m = he * hn[r] ; msg = m Wm + bm
//                          a3 = e Wge + gx[r] ; r3 = relu(LN(a3)) ; gt = r3 Wg2 + bg2 ; s = sigmoid(gt) ; out = msg * s
//   backward (d out = dagg[l]):  d msg = d out * s ; d gt = d out * msg * s (1 - s)
//     gate branch:     d r3 = d gt Wg2^T ; d a3 = LNbwd(d r3; a3) ; dgx[r] += d a3 ; d e += d a3 Wge^T
//     message branch:  d m = d msg Wm^T ; dhn[r] += d m * he ; d he = d m * hn[r] ; d r2 = d he W2^T ;
//                      d a2 = LNbwd(d r2; a2) ; d e += d a2 W1^T
//
// What changed against the 8-warp kernel (ncu source view: its 270 KB of unrolled epilogue code stalled 50-60 % of the
// issue slots on instruction fetch, 750 B/thread of spills, 14 GEMMs per tile):
//   * epilogues are `#pragma unroll 1` loops over 16-column chunks that re-read their operands from TMEM; no register
//     arrays beyond one chunk, parameter vectors via warp-uniform LDS.128;
//   * the two big tiles that had to be parked while the other branch ran no longer sit in TMEM/registers: d msg leaves
//     the CTA as ready-made A-operand planes (fp16 hi|lo, canonical UMMA layout) in a global scratch slab and comes back
//     with two 64 KB bulk copies issued by the producer warp; he (needed again for dhn) goes through a fp32 scratch slab.
//     Both slabs stay L2-resident (126 MB L2); HBM bandwidth is idle on this path anyway.  12 GEMMs per tile.
//   * LayerNorm statistics of the two LN layers are kept from the recompute for their backward.
//
// TMEM: two 256-column accumulators A0 / A1 (see tc_nodeblock16.cuh for the sliced-publication hazard rule).
// Included by mdb_forward.cu inside its anonymous namespace (after tc_nodeblock16.cuh).
#pragma once
#include "tc_pipe.cuh"

#undef TC_STAMP
#define TC_STAMP(i) do { if (a.dbg && tid == 0) a.dbg[(size_t)tile * 32 + (i)] = clock64(); } while (0)

struct TcNbBwd16Args {
  const float* blob;
  const uint8_t* tc_blob;
  BlkOff off;
  TcOff tco;
  Tables tb;
  const int *left, *right;
  int n_nodes, n_edges;
  const float* e;        // [E][64] saved e_i
  const float* dagg;     // [N][256] d/d (aggregated messages)
  float *dgx, *dhn;      // [pad64(N)][256] scatter targets (pre-zeroed), NODE-BLOCKED layout (tile_engine.cuh: blk_off)
  float* de;             // [E][64]  d/d e, accumulated (+=)
  float* scr_he;         // [gridDim.x * 128][256] fp32 scratch, one slab per (persistent) CTA
  uint8_t* scr_dm;       // [gridDim.x][2][64 KB] d msg as A-operand planes
  long long* dbg;
};

constexpr uint32_t PLANE256_BYTES = tc::ROWS * D * 2;   // one K = 256 A plane: 64 KB

// 16 floats of a gathered / scratch row, software-prefetched one chunk ahead in the rolled loops
struct Row16 { float4 v[4]; };
__device__ __forceinline__ Row16 ld_row16(const float* __restrict__ p) {
  Row16 r;
#pragma unroll
  for (int j = 0; j < 4; ++j) r.v[j] = *reinterpret_cast<const float4*>(p + 4 * j);
  return r;
}
// chunk c (16 columns = four 16-byte pieces, PS floats apart) of this thread's row in a blocked layout: the tile-blocked
// he scratch (PS = 512) or a node-blocked table (PS = BLK_PIECE_STRIDE)
template <int PS>
__device__ __forceinline__ Row16 ld_blk16(const float* __restrict__ base, int c) {
  Row16 r;
#pragma unroll
  for (int j = 0; j < 4; ++j) r.v[j] = *reinterpret_cast<const float4*>(base + (c * 4 + j) * PS);
  return r;
}
__device__ __forceinline__ Row16 ld_he16(const float* __restrict__ base, int c) { return ld_blk16<tc::ROWS * 4>(base, c); }
__device__ __forceinline__ Row16 ld_tab16(const float* __restrict__ base, int c) { return ld_blk16<BLK_PIECE_STRIDE>(base, c); }
__device__ __forceinline__ void unpack_row16(const Row16& r, float (&x)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) { x[4 * j] = r.v[j].x; x[4 * j + 1] = r.v[j].y; x[4 * j + 2] = r.v[j].z; x[4 * j + 3] = r.v[j].w; }
}
__device__ __forceinline__ void lds16(const float* __restrict__ p, float (&x)[16]) {   // warp-uniform address: broadcast
  Row16 r = ld_row16(p);
  unpack_row16(r, x);
}

template <bool IS_ROW>
__device__ __forceinline__ void tc_nodeblock_bwd16_body(const TcNbBwd16Args& a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* e_hi = smem_raw;
  uint8_t* e_lo = e_hi + tc::ROWS * C * 2;
  uint8_t* x_hi = e_lo + tc::ROWS * C * 2;
  uint8_t* x_lo = x_hi + PLANE256_BYTES;
  uint8_t* stages = x_lo + PLANE256_BYTES;
  PipeSmem16* ps = reinterpret_cast<PipeSmem16*>(stages + NB16_NS * tc::STAGE_SLOT);
  uint64_t* x_free = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ps) + 112);   // MMA -> producer: X planes consumed
  uint64_t* x_full = x_free + 1;                                                          // bulk copies -> MMA: d msg planes landed
  float2* stat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(ps) + 128);        // [4][128]
  float* vecs = reinterpret_cast<float*>(smem_raw + NB16_VEC_OFF);
  const float *v_en1_b = vecs, *v_en1_g = vecs + D, *v_en1_be = vecs + 2 * D, *v_en2_b = vecs + 3 * D, *v_msg_b = vecs + 4 * D,
              *v_g1_g = vecs + 5 * D, *v_g1_be = vecs + 6 * D, *v_g2_b = vecs + 7 * D;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Tables& tb = a.tb;
  const int n_tiles = (a.n_edges + tc::ROWS - 1) / tc::ROWS;
  Pipe16 p;
  tc::pipe_init_split<NB16_NRW, IS_ROW, NB16_NS>(p, ps, stages);
  if (tid == 0) { tc::mbar_init(x_free, 1); tc::mbar_init(x_full, 1); tc::fence_barrier_init(); }
  if (warp == NB16_NRW) tc::tmem_alloc<512>(&ps->tmem_base);
  const int row = (warp & 3) * 32 + lane;
  const int part = (warp >> 2) & 3;
  const int pc = part * 64;
  if (IS_ROW) {
    const int vj = tid >> 6;                    // 64 threads per parameter vector
    int so = a.off.o[MDB_S_NB_EN1_B];
    if (vj == 1) so = a.off.o[MDB_S_NB_EN1_G];
    if (vj == 2) so = a.off.o[MDB_S_NB_EN1_BE];
    if (vj == 3) so = a.off.o[MDB_S_NB_EN2_B];
    if (vj == 4) so = a.off.o[MDB_S_NB_MSG_B];
    if (vj == 5) so = a.off.o[MDB_S_NB_G1_G];
    if (vj == 6) so = a.off.o[MDB_S_NB_G1_BE];
    if (vj == 7) so = a.off.o[MDB_S_NB_G2_B];
    *reinterpret_cast<float4*>(vecs + tid * 4) = *reinterpret_cast<const float4*>(a.blob + so + (tid & 63) * 4);
  }
  tc::fence_before_sync();
  tc::cta_sync();
  tc::fence_after_sync();
  const uint32_t lane_base = ps->tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t A0 = lane_base + pc, A1 = lane_base + 256 + pc;

  // persistent CTA (see tc_nodeblock16.cuh); `it` counts this CTA's tiles: parity of the two X-plane reload barriers
#pragma unroll 1
  for (int tile = blockIdx.x, it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
  const int q0 = tile * tc::ROWS;
  TC_STAMP(0);
  const int q = q0 + row;
  const bool valid = IS_ROW && q < a.n_edges;
  float e16[16];
  if (IS_ROW) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) x = *reinterpret_cast<const float4*>(a.e + (size_t)q * C + part * 16 + i);
      e16[i] = x.x; e16[i + 1] = x.y; e16[i + 2] = x.z; e16[i + 3] = x.w;
    }
  }
  const int ll = valid ? a.left[q] : 0, rr = valid ? a.right[q] : 0;
  TC_STAMP(1);
  const float* hn = tb.hnb + blk_off(rr, pc / 4);      // node-blocked tables (tile_engine.cuh): coalesced gathers / REDs
  const float* gxr = tb.gxb + blk_off(rr, pc / 4);
  // he scratch, tile-blocked so that a warp instruction (32 consecutive rows, one 16-byte piece each) touches 4 lines
  // instead of 32: [tile][16-byte column piece 0..63][row 0..127][4 floats]
  // Both slabs are indexed by CTA, not by tile: the kernel is persistent, a CTA finishes with its slab (same threads write
  // and read he; the d msg planes are read back by the bulk copy long before the next tile's epilogue rewrites them)
  // before its next tile, so gridDim.x x 256 KB (38 MB) stays L2-resident instead of tiles x 256 KB of write-backs.
  float* he_scr = a.scr_he + (size_t)blockIdx.x * tc::ROWS * D + (size_t)(pc / 4) * tc::ROWS * 4 + row * 4;
  uint8_t* dm_scr = a.scr_dm + (size_t)blockIdx.x * 2 * PLANE256_BYTES;
  float2 ms_en1 = make_float2(0.f, 1.f), ms_g1 = make_float2(0.f, 1.f);
  float de16[16];

  if (IS_ROW) {   // e tile -> E planes
    tc::store_a<C, 16>(e_hi, e_lo, row, part * 16, e16);
    tc::rows_publish(p);
    TC_STAMP(2);
  }
  // ---- forward recompute ---------------------------------------------------------------------------------------
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_EN1), 256, false, true, true);                    // a2 -> A1
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    TC_STAMP(3);
    RunStat rs = {0.f, 0.f, 0.f};
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float x[16], b[16];
      tc::tmem_ld16(A1 + c * 16, x);
      lds16(v_en1_b + pc + c * 16, b);
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] += b[i];
      stat_add16(rs, x);
    }
    ms_en1 = ln_merge_quarter(rs.mean, rs.m2, stat, row, part);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float x[16], b[16], g[16], be[16];
      tc::tmem_ld16(A1 + c * 16, x);
      lds16(v_en1_b + pc + c * 16, b); lds16(v_en1_g + pc + c * 16, g); lds16(v_en1_be + pc + c * 16, be);
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaxf((x[i] + b[i] - ms_en1.x) * ms_en1.y * g[i] + be[i], 0.f);
      store_a16(x_hi, x_lo, row, pc + c * 16, x);
      tc::rows_publish_group(p, c);
    }
    TC_STAMP(4);
  }
  tc::gemm<D, D, NB16_NS, 4>(p, x_hi, x_lo, TCW_(NB_EN2), 0, false, true, true);          // he -> A0
  if (IS_ROW) {
    Row16 nx = ld_tab16(hn, 0);                     // first chunk of hn[r]: in flight while the GEMM runs
    tc::rows_wait_acc(p);
    TC_STAMP(5);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float x[16], b[16], h[16];
      unpack_row16(nx, h);
      if (c < 3) nx = ld_tab16(hn, c + 1);
      tc::tmem_ld16(A0 + c * 16, x);
      lds16(v_en2_b + pc + c * 16, b);
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] += b[i];
#pragma unroll
      for (int i = 0; i < 16; i += 4)            // he: needed again by the message branch
        *reinterpret_cast<float4*>(he_scr + (c * 4 + i / 4) * tc::ROWS * 4) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] *= h[i];
      store_a16(x_hi, x_lo, row, pc + c * 16, x);
      tc::rows_publish_group(p, c);
    }
    TC_STAMP(6);
  }
  tc::gemm<D, D, NB16_NS, 4>(p, x_hi, x_lo, TCW_(NB_MSG), 256, false, true, false);        // msg -> A1
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_GE), 0, false, false, true);                       // a3 - gx -> A0
  if (IS_ROW) {
    float gv[64];                                // gx[r] row part, requested before the accumulator wait
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 t4 = *reinterpret_cast<const float4*>(gxr + j * BLK_PIECE_STRIDE);
      gv[4 * j] = t4.x; gv[4 * j + 1] = t4.y; gv[4 * j + 2] = t4.z; gv[4 * j + 3] = t4.w;
    }
    tc::rows_wait_acc(p);
    TC_STAMP(7);
    RunStat rs = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c) {                // unrolled: a3 = acc + gx stays in registers (accumulator drained before
      float x[16];                               // the first slice is published, so the next GEMM may overwrite A0)
      tc::tmem_ld16(A0 + c * 16, x);
#pragma unroll
      for (int i = 0; i < 16; ++i) { x[i] += gv[c * 16 + i]; gv[c * 16 + i] = x[i]; }
      stat_add16(rs, x);
    }
    ms_g1 = ln_merge_quarter(rs.mean, rs.m2, stat, row, part);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float x[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = pc + c * 16 + i;
        x[i] = fmaxf((gv[c * 16 + i] - ms_g1.x) * ms_g1.y * v_g1_g[k] + v_g1_be[k], 0.f);
      }
      store_a16(x_hi, x_lo, row, pc + c * 16, x);
      tc::rows_publish_group(p, c);
    }
    TC_STAMP(8);
  }
  tc::gemm<D, D, NB16_NS, 4>(p, x_hi, x_lo, TCW_(NB_G2), 0, false, true, true);            // gt -> A0
  // ---- d out = dagg[l]:  d gt -> X planes ;  d msg -> scratch planes ------------------------------------------------
  if (IS_ROW) {
    const float* dout = a.dagg + (size_t)ll * D + pc;
    Row16 nx = ld_row16(dout);
    tc::rows_wait_acc(p);
    TC_STAMP(9);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float gt[16], msg[16], dz[16], bg[16], bm[16];
      unpack_row16(nx, dz);
      if (c < 3) nx = ld_row16(dout + (c + 1) * 16);
      tc::tmem_ld16(A0 + c * 16, gt);
      tc::tmem_ld16(A1 + c * 16, msg);
      lds16(v_g2_b + pc + c * 16, bg); lds16(v_msg_b + pc + c * 16, bm);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float s = tc::fast_sigmoid(gt[i] + bg[i]);
        const float d = valid ? dz[i] : 0.f;
        gt[i] = d * (msg[i] + bm[i]) * s * (1.f - s);     // d gate logit
        msg[i] = d * s;                                    // d msg
      }
      store_a16(x_hi, x_lo, row, pc + c * 16, gt);
      store_a16(dm_scr, dm_scr + PLANE256_BYTES, row, pc + c * 16, msg);
    }
    tc::fence_proxy_async_global();              // the scratch planes are read back by the async proxy (bulk copy)
    tc::rows_publish(p);
    TC_STAMP(10);
  }
  tc::gemm<D, D>(p, x_hi, x_lo, TCW_(BT_NB_G2), 256, false, true, false);                  // d r3 -> A1
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_GE), 0, false, false, true);                       // a3 - gx -> A0 (again)
  if (IS_ROW) {
    Row16 nx = ld_tab16(gxr, 0);
    tc::rows_wait_acc(p);
    TC_STAMP(11);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {                // pass 1: a3 = acc + gx (folded back) ; d xhat -> A1 ; partial sums
      float x[16], g[16], d[16], ga[16], be[16];
      unpack_row16(nx, g);
      if (c < 3) nx = ld_tab16(gxr, c + 1);
      tc::tmem_ld16(A0 + c * 16, x);
      tc::tmem_ld16(A1 + c * 16, d);
      lds16(v_g1_g + pc + c * 16, ga); lds16(v_g1_be + pc + c * 16, be);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        x[i] += g[i];
        const float xh = (x[i] - ms_g1.x) * ms_g1.y;
        const float dxh = (xh * ga[i] + be[i] > 0.f) ? d[i] * ga[i] : 0.f;
        d[i] = dxh;
        s1 += dxh;
        s2 = fmaf(dxh, xh, s2);
      }
      tc::tmem_st16(A0 + c * 16, x);
      tc::tmem_st16(A1 + c * 16, d);
    }
    tc::tmem_st_wait();
    // row sums over the four parts
    stat[part * tc::ROWS + row] = make_float2(s1, s2);
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const float2 t0 = stat[row], t1 = stat[tc::ROWS + row], t2 = stat[2 * tc::ROWS + row], t3 = stat[3 * tc::ROWS + row];
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const float m1 = ((t0.x + t1.x) + (t2.x + t3.x)) * (1.f / 256.f), m2 = ((t0.y + t1.y) + (t2.y + t3.y)) * (1.f / 256.f);
    float* dst = a.dgx + blk_off(rr, pc / 4);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {                // pass 2: d a3 -> dgx[r] (RED) and the X planes
      float x[16], d[16];
      tc::tmem_ld16(A0 + c * 16, x);
      tc::tmem_ld16(A1 + c * 16, d);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float xh = (x[i] - ms_g1.x) * ms_g1.y;
        d[i] = ms_g1.y * (d[i] - m1 - xh * m2);
      }
      if (valid) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          tc::red_add_v4(dst + (c * 4 + i / 4) * BLK_PIECE_STRIDE, d[i], d[i + 1], d[i + 2], d[i + 3]);
      }
      store_a16(x_hi, x_lo, row, pc + c * 16, d);
    }
    tc::rows_publish(p);
    TC_STAMP(12);
  }
  tc::gemm<D, C>(p, x_hi, x_lo, TCW_(BT_NB_GE), 0, false, true, true);                     // d e (gate part) -> A0[0:64]
  // ---- d msg planes come back from the scratch slab (no row-thread work) ---------------------------------------
  if (!IS_ROW) {
    if (p.role == 2) {
      tc::mma_commit_elect(x_free);              // arrives when BT_NB_GE has finished reading the X planes
      tc::mbar_wait(x_full, it & 1);
      tc::fence_after_sync();
    } else if (p.role == 1) {
      tc::mbar_wait(x_free, it & 1);
      if (lane == 0) {
        tc::mbar_arrive_expect_tx(x_full, 2 * PLANE256_BYTES);
        tc::bulk_g2s(x_hi, dm_scr, PLANE256_BYTES, x_full);
        tc::bulk_g2s(x_lo, dm_scr + PLANE256_BYTES, PLANE256_BYTES, x_full);
      }
      __syncwarp();
    }
  }
  tc::gemm<D, D>(p, x_hi, x_lo, TCW_(BT_NB_MSG), 256, false, true, true);                  // d m -> A1
  if (IS_ROW) {
    tc::rows_wait_acc(p);                        // BT_NB_GE
    TC_STAMP(13);
    tc::tmem_ld16(lane_base + part * 16, de16);
    tc::rows_publish(p);                         // keeps the `done` barrier at most one phase ahead of the row threads
    Row16 nh = ld_tab16(hn, 0), ne = ld_he16(he_scr, 0);
    tc::rows_wait_acc(p);                        // BT_NB_MSG
    TC_STAMP(14);
    float* dst = a.dhn + blk_off(rr, pc / 4);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float dm[16], h[16], he[16];
      unpack_row16(nh, h); unpack_row16(ne, he);
      if (c < 3) { nh = ld_tab16(hn, c + 1); ne = ld_he16(he_scr, c + 1); }
      tc::tmem_ld16(A1 + c * 16, dm);
      if (valid) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          tc::red_add_v4(dst + (c * 4 + i / 4) * BLK_PIECE_STRIDE, dm[i] * he[i], dm[i + 1] * he[i + 1], dm[i + 2] * he[i + 2],
                         dm[i + 3] * he[i + 3]);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) dm[i] *= h[i];           // d he
      store_a16(x_hi, x_lo, row, pc + c * 16, dm);
      tc::rows_publish_group(p, c);
    }
    TC_STAMP(15);
  }
  tc::gemm<D, D, NB16_NS, 4>(p, x_hi, x_lo, TCW_(BT_NB_EN2), 0, false, true, false);       // d r2 -> A0
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_EN1), 256, false, false, true);                    // a2 - b1 -> A1 (again)
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    TC_STAMP(16);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float x[16], d[16], b[16], ga[16], be[16];
      tc::tmem_ld16(A1 + c * 16, x);
      tc::tmem_ld16(A0 + c * 16, d);
      lds16(v_en1_b + pc + c * 16, b); lds16(v_en1_g + pc + c * 16, ga); lds16(v_en1_be + pc + c * 16, be);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float xh = (x[i] + b[i] - ms_en1.x) * ms_en1.y;
        const float dxh = (xh * ga[i] + be[i] > 0.f) ? d[i] * ga[i] : 0.f;
        d[i] = dxh;
        s1 += dxh;
        s2 = fmaf(dxh, xh, s2);
      }
      tc::tmem_st16(A0 + c * 16, d);
    }
    tc::tmem_st_wait();
    stat[part * tc::ROWS + row] = make_float2(s1, s2);
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const float2 t0 = stat[row], t1 = stat[tc::ROWS + row], t2 = stat[2 * tc::ROWS + row], t3 = stat[3 * tc::ROWS + row];
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const float m1 = ((t0.x + t1.x) + (t2.x + t3.x)) * (1.f / 256.f), m2 = ((t0.y + t1.y) + (t2.y + t3.y)) * (1.f / 256.f);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float x[16], d[16], b[16];
      tc::tmem_ld16(A1 + c * 16, x);
      tc::tmem_ld16(A0 + c * 16, d);
      lds16(v_en1_b + pc + c * 16, b);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float xh = (x[i] + b[i] - ms_en1.x) * ms_en1.y;
        d[i] = ms_en1.y * (d[i] - m1 - xh * m2);
      }
      store_a16(x_hi, x_lo, row, pc + c * 16, d);
    }
    tc::rows_publish(p);
    TC_STAMP(17);
  }
  tc::gemm<D, C>(p, x_hi, x_lo, TCW_(BT_NB_EN1), 0, false, true, true);                    // d e (message part) -> A0[0:64]
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    TC_STAMP(18);
    float v[16];
    tc::tmem_ld16(lane_base + part * 16, v);
    if (valid) {
      float* dst = a.de + (size_t)q * C + part * 16;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 o = *reinterpret_cast<float4*>(dst + i);
        o.x += de16[i] + v[i]; o.y += de16[i + 1] + v[i + 1]; o.z += de16[i + 2] + v[i + 2]; o.w += de16[i + 3] + v[i + 3];
        *reinterpret_cast<float4*>(dst + i) = o;
      }
    }
    tc::fence_before_sync();
  }
  TC_STAMP(19);
  }   // tile loop
  tc::cta_sync();
  if (warp == NB16_NRW) { __syncwarp(); tc::tmem_dealloc<512>(ps->tmem_base); }
}

__global__ void __launch_bounds__(NB16_THREADS, 1) tc_nodeblock_bwd16_kernel(const __grid_constant__ TcNbBwd16Args a) {
  if (threadIdx.x < NB16_NRW * 32) {
    tc::reg_alloc<112>();
    tc_nodeblock_bwd16_body<true>(a);
  } else {
    tc::reg_dealloc<32>();
    tc_nodeblock_bwd16_body<false>(a);
  }
}

constexpr size_t SMEM_TC_NB_BWD16 = NB16_VEC_OFF + 8 * D * sizeof(float);
static_assert(sizeof(PipeSmem16) <= 112, "two extra barriers live at byte 112 of the pipe slot");
