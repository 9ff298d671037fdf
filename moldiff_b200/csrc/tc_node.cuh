// Tensor-core (tcgen05) version of node_kernel: everything per NODE of one block boundary -- NodeBlock node tail
// (graph.py:51-54), h_node residual (:363), PosUpdate node-side MLPs (:387-388), every hoisted per-node
// projection of the next block, and the decoders (model.py:226, bond_predictor.py:155-160).
// 128 nodes per CTA, two threads per row.  At BASELINE config 2 (N = 6 286) that is only 50 CTAs, but each is
// ~8x faster per row than the 64-row FFMA tile (99 CTAs), so the launch is ~3x shorter.
//
// Included by mdb_forward.cu inside its anonymous namespace (after tc_nodeblock16.cuh).
#pragma once
#include "tc_pipe.cuh"

struct PuSideVecs { float l1_b[C], l1_g[C], l1_be[C], l2_b[C]; };
struct NodeVecs {
  float ln_g[D], ln_be[D], out_b[D];                               // mid
  PuSideVecs pu[2];
  float nn1_b[D], nn1_g[D], nn1_be[D], nn2_b[D], gx_b[D], gx_t[D], cen_b[D];   // pre
  float gn_b[2][32], nf_b[2][C];
  float d1_b[D], d1_g[D], d1_be[D], d2_b[32];                      // decode
};

struct TcOffHead { int64_t o[MDB_NUM_TC_HEAD_SLOTS]; };

struct TcNodeArgs {
  const uint8_t* tc_blob;
  TcOff mid, pre;
  TcOffHead hd;
  Tables tb, tb_pre;
  int n_nodes;
  int do_mid, do_pre, do_dec, update_pos, kind, kn;
  float* sl_next;
  float *fl_next, *fr_next;
  float* x_save;
  float* agg_save;
  const float* pos_cur;
  float* pos_nxt;
  float* pred_node;
  NodeVecs v;
};
#define TCM_(slot) (a.tc_blob + a.mid.o[MDB_T_##slot])
#define TCP_(slot) (a.tc_blob + a.pre.o[MDB_T_##slot])
#define TCH_(slot) (a.tc_blob + a.hd.o[MDB_TH_##slot])

// this thread's 128 columns of table row n <- v (row-major table, 16-byte stores)
__device__ __forceinline__ void store_row128(float* __restrict__ dst, const float (&v)[128]) {
#pragma unroll
  for (int i = 0; i < 128; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}
// same into the node-blocked copy of the table (tile_engine.cuh: blk_off); lanes = consecutive nodes -> 512 B per store
__device__ __forceinline__ void store_row128_blocked(float* __restrict__ table, int n, int c0, const float (&v)[128]) {
  float* dst = table + blk_off(n, c0 / 4);
#pragma unroll
  for (int i = 0; i < 128; i += 4)
    *reinterpret_cast<float4*>(dst + (i / 4) * BLK_PIECE_STRIDE) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}
template <int NC>
__device__ __forceinline__ void store_rowN(float* __restrict__ dst, const float (&v)[NC]) {
#pragma unroll
  for (int i = 0; i < NC; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}

template <bool XF>   // cross-first accumulation order (tc_pipe.cuh)
__global__ void __launch_bounds__(TC_NB_THREADS, 1) tc_node_kernel(const __grid_constant__ TcNodeArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* p_hi = smem_raw;                                  // K = 256 planes (h_node / hidden activations)
  uint8_t* p_lo = p_hi + tc::ROWS * D * 2;
  uint8_t* s_hi = p_lo + tc::ROWS * D * 2;                   // K = 64 planes (PosUpdate hidden)
  uint8_t* s_lo = s_hi + tc::ROWS * C * 2;
  uint8_t* stages = s_lo + tc::ROWS * C * 2;
  tc::PipeSmem* ps = reinterpret_cast<tc::PipeSmem*>(stages + tc::NSTAGE * tc::STAGE_SLOT);
  float2* stat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(ps) + 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * tc::ROWS;
  const Tables& tb = a.tb;
  tc::PipeT<tc::NSTAGE, XF> p;
  tc::pipe_init<TC_NRW>(p, ps, stages);
  if (warp == TC_NRW) tc::tmem_alloc<512>(&ps->tmem_base);
  const int row = (warp & 3) * 32 + lane;
  const int half = (warp >> 2) & 1;
  const int hc = half * 128;
  const int n = row0 + row;
  const bool valid = p.role == 0 && n < a.n_nodes;
  const size_t nn = valid ? (size_t)n : 0;
  const float tnv = valid ? tb.tn[n] : 0.f;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t lane_base = ps->tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t D0 = 0, D1 = 256;

  if (a.do_mid) {
    if (p.role == 0) {   // relu(LN(centroid_lin(x) + aggr))                          graph.py:51-54
      float v[128];
#pragma unroll
      for (int i = 0; i < 128; i += 4) {
        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = c4;
        if (valid) {
          c4 = *reinterpret_cast<const float4*>(tb.cen + nn * D + hc + i);
          g4 = *reinterpret_cast<const float4*>(tb.agg + nn * D + hc + i);
          if (a.agg_save) *reinterpret_cast<float4*>(a.agg_save + nn * D + hc + i) = g4;
          *reinterpret_cast<float4*>(tb.agg + nn * D + hc + i) = make_float4(0.f, 0.f, 0.f, 0.f);   // re-arm
        }
        v[i] = c4.x + g4.x; v[i + 1] = c4.y + g4.y; v[i + 2] = c4.z + g4.z; v[i + 3] = c4.w + g4.w;
      }
      ln_relu_half(v, a.v.ln_g + hc, a.v.ln_be + hc, stat, row, half);
      store_half_row_a(p_hi, p_lo, row, hc, v);
      tc::rows_publish(p);
    }
    tc::gemm<D, D>(p, p_hi, p_lo, TCM_(NB_OUT), D0, false, true, true);
    if (p.role == 0) {   // h_node = h_node + out_transform(.)                          graph.py:54,363
      tc::rows_wait_acc(p);
      float v[128];
      load_half_row(lane_base + D0 + hc, v);
      add_vec128(v, a.v.out_b + hc);
      if (valid) add_row128(v, tb.x + nn * D + hc);
      if (valid) {
        store_row128(tb.x + nn * D + hc, v);
        if (a.x_save && a.do_pre) store_row128(a.x_save + nn * D + hc, v);
        if (a.update_pos && half == 0) {   // pos_nxt starts as pos_cur; the edge kernel adds the forces   graph.py:366
          a.pos_nxt[nn * 3 + 0] = a.pos_cur[nn * 3 + 0];
          a.pos_nxt[nn * 3 + 1] = a.pos_cur[nn * 3 + 1];
          a.pos_nxt[nn * 3 + 2] = a.pos_cur[nn * 3 + 2];
        }
      }
      store_half_row_a(p_hi, p_lo, row, hc, v);
      tc::rows_publish(p);
    }
    if (a.update_pos) {   // PosUpdate left / right_lin_edge MLPs on the NEW h_node          graph.py:387-388
#pragma unroll 1
      for (int side = 0; side < 2; ++side) {
        const PuSideVecs& pv = a.v.pu[side];
        tc::gemm<D, C>(p, p_hi, p_lo, side ? TCM_(PU_RL1) : TCM_(PU_LL1), D1, false, true, true);
        if (p.role == 0) {
          tc::rows_wait_acc(p);
          float u[32];
          tc::tmem_ld32(lane_base + D1 + half * 32, u);
#pragma unroll
          for (int i = 0; i < 32; ++i) u[i] += pv.l1_b[half * 32 + i];
          ln_relu_part<32>(u, pv.l1_g + half * 32, pv.l1_be + half * 32, stat, row, half);
          tc::store_a<C, 32>(s_hi, s_lo, row, half * 32, u);
          tc::rows_publish(p);
        }
        tc::gemm<C, C>(p, s_hi, s_lo, side ? TCM_(PU_RL2) : TCM_(PU_LL2), D1 + 64, false, true, true);
        if (p.role == 0) {
          tc::rows_wait_acc(p);
          float u[32];
          tc::tmem_ld32(lane_base + D1 + 64 + half * 32, u);
#pragma unroll
          for (int i = 0; i < 32; ++i) u[i] += pv.l2_b[half * 32 + i];
          if (valid) store_rowN<32>((side ? tb.rf : tb.lf) + nn * C + half * 32, u);
          tc::rows_publish(p);
        }
      }
    }
  } else {
    if (p.role == 0) {   // first call of a forward: h_node planes straight from memory
      float v[128];
#pragma unroll
      for (int i = 0; i < 128; ++i) v[i] = 0.f;
      if (valid) {
        add_row128(v, tb.x + nn * D + hc);
        if (a.x_save && a.do_pre) store_row128(a.x_save + nn * D + hc, v);
      }
      store_half_row_a(p_hi, p_lo, row, hc, v);
      tc::rows_publish(p);
    }
  }

  if (a.do_pre) {
    const Tables& tp = a.tb_pre;
    // gate.net.0 node / time / bias part ; centroid_lin                             graph.py:46,51
    tc::gemm<D, D>(p, p_hi, p_lo, TCP_(NB_GX), D0, false, true, false);
    tc::gemm<D, D>(p, p_hi, p_lo, TCP_(NB_CEN), D1, false, false, true);
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      float v[128];
      load_half_row(lane_base + D0 + hc, v);
#pragma unroll
      for (int i = 0; i < 128; ++i) v[i] += a.v.gx_b[hc + i] + tnv * a.v.gx_t[hc + i];
      if (valid) { store_row128(tp.gx + nn * D + hc, v); store_row128_blocked(tp.gxb, n, hc, v); }
      load_half_row(lane_base + D1 + hc, v);
      add_vec128(v, a.v.cen_b + hc);
      if (valid) store_row128(tp.cen + nn * D + hc, v);
      tc::rows_publish(p);
    }
    // EdgeBlock hoists: node_linear L/R, gate node columns L/R, node_ffn L/R        graph.py:135,139,288-289
    tc::gemm<D, 128>(p, p_hi, p_lo, TCP_(EL_NL), D0, false, true, false);
    tc::gemm<D, 128>(p, p_hi, p_lo, TCP_(ER_NL), D0 + 128, false, false, false);
    tc::gemm<D, 32>(p, p_hi, p_lo, TCP_(EL_GN), D1, false, false, false);
    tc::gemm<D, 32>(p, p_hi, p_lo, TCP_(ER_GN), D1 + 32, false, false, false);
    tc::gemm<D, C>(p, p_hi, p_lo, TCP_(EB_NFL), D1 + 64, false, false, false);
    tc::gemm<D, C>(p, p_hi, p_lo, TCP_(EB_NFR), D1 + 128, false, false, true);
    if (p.role == 0) {
      tc::rows_wait_acc(p);
#pragma unroll 1
      for (int side = 0; side < 2; ++side) {
        float w[64];
        load_cols_tm<64>(lane_base + D0 + side * 128 + half * 64, w);
        if (valid) store_rowN<64>((side ? tp.nlr : tp.nll) + nn * 128 + half * 64, w);
        float g[16];
        load_cols_tm<16>(lane_base + D1 + side * 32 + half * 16, g);
#pragma unroll
        for (int i = 0; i < 16; ++i) g[i] += a.v.gn_b[side][half * 16 + i];
        if (valid) store_rowN<16>((side ? tp.gnr : tp.gnl) + nn * 32 + half * 16, g);
        float f[32];
        tc::tmem_ld32(lane_base + D1 + 64 + side * 64 + half * 32, f);
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] += a.v.nf_b[side][half * 32 + i];
        if (valid) store_rowN<32>((side ? a.fr_next : a.fl_next) + nn * C + half * 32, f);
      }
      if (valid && a.sl_next) {   // clear the SL / SR accumulators of the next block
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          *reinterpret_cast<float4*>(a.sl_next + nn * C + half * 32 + i) = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(a.sl_next + (size_t)a.n_nodes * C + nn * C + half * 32 + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      tc::rows_publish(p);
    }
    // node_net MLP                                                                   graph.py:39
    tc::gemm<D, D>(p, p_hi, p_lo, TCP_(NB_NN1), D0, false, true, true);
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      float v[128];
      load_half_row(lane_base + D0 + hc, v);
      add_vec128(v, a.v.nn1_b + hc);
      ln_relu_half(v, a.v.nn1_g + hc, a.v.nn1_be + hc, stat, row, half);
      store_half_row_a(p_hi, p_lo, row, hc, v);       // h_node planes are dead: every other projection is done
      tc::rows_publish(p);
    }
    tc::gemm<D, D>(p, p_hi, p_lo, TCP_(NB_NN2), D0, false, true, true);
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      float v[128];
      load_half_row(lane_base + D0 + hc, v);
      add_vec128(v, a.v.nn2_b + hc);
      if (valid) { store_row128(tp.hn + nn * D + hc, v); store_row128_blocked(tp.hnb, n, hc, v); }
    }
  }

  if (a.do_dec) {
    if (a.kind == 1) {   // node_decoder MLP(256 -> 256 -> Kn)                          model.py:226
      tc::gemm<D, D>(p, p_hi, p_lo, TCH_(NDEC1), D0, false, true, true);
      if (p.role == 0) {
        tc::rows_wait_acc(p);
        float v[128];
        load_half_row(lane_base + D0 + hc, v);
        add_vec128(v, a.v.d1_b + hc);
        ln_relu_half(v, a.v.d1_g + hc, a.v.d1_be + hc, stat, row, half);
        store_half_row_a(p_hi, p_lo, row, hc, v);
        tc::rows_publish(p);
      }
      tc::gemm<D, 32>(p, p_hi, p_lo, TCH_(NDEC2), D1, false, true, true);
      if (p.role == 0) {
        tc::rows_wait_acc(p);
        float o[16];
        load_cols_tm<16>(lane_base + D1 + half * 16, o);
        if (valid) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (half * 16 + i < a.kn) a.pred_node[nn * a.kn + half * 16 + i] = o[i] + a.v.d2_b[half * 16 + i];
        }
      }
    } else if (a.kind == 2) {   // node half of edge_decoder.net.0                      bond_predictor.py:155-160
      tc::gemm<D, C>(p, p_hi, p_lo, TCH_(EDEC1N), D0, false, true, true);
      if (p.role == 0) {
        tc::rows_wait_acc(p);
        float o[32];
        tc::tmem_ld32(lane_base + D0 + half * 32, o);
        if (valid) store_rowN<32>(tb.dect + nn * C + half * 32, o);
      }
    }
  }
  if (p.role == 0) tc::fence_before_sync();
  __syncthreads();
  if (warp == TC_NRW) { __syncwarp(); tc::tmem_dealloc<512>(ps->tmem_base); }
}

constexpr size_t SMEM_TC_NODE = 2 * (size_t)tc::ROWS * D * 2 + 2 * (size_t)tc::ROWS * C * 2 + tc::NSTAGE * tc::STAGE_SLOT
                                + 128 + 2 * tc::ROWS * sizeof(float2) + 64;
static_assert(SMEM_TC_NODE <= 232448, "tc_node_kernel exceeds the shared-memory limit");

inline void fill_node_vecs(NodeVecs& v, const float* bh, const BlkOff* mid, const BlkOff* pre, const HeadOff* hd, int kind,
                           bool update_pos) {
  auto cp = [&](float* dst, int off, int n) { memcpy(dst, bh + off, n * sizeof(float)); };
  if (mid) {
    cp(v.ln_g, mid->o[MDB_S_NB_LN_G], D); cp(v.ln_be, mid->o[MDB_S_NB_LN_BE], D); cp(v.out_b, mid->o[MDB_S_NB_OUT_B], D);
    if (update_pos) {
      cp(v.pu[0].l1_b, mid->o[MDB_S_PU_LL1_B], C); cp(v.pu[0].l1_g, mid->o[MDB_S_PU_LL1_G], C);
      cp(v.pu[0].l1_be, mid->o[MDB_S_PU_LL1_BE], C); cp(v.pu[0].l2_b, mid->o[MDB_S_PU_LL2_B], C);
      cp(v.pu[1].l1_b, mid->o[MDB_S_PU_RL1_B], C); cp(v.pu[1].l1_g, mid->o[MDB_S_PU_RL1_G], C);
      cp(v.pu[1].l1_be, mid->o[MDB_S_PU_RL1_BE], C); cp(v.pu[1].l2_b, mid->o[MDB_S_PU_RL2_B], C);
    }
  }
  if (pre) {
    cp(v.nn1_b, pre->o[MDB_S_NB_NN1_B], D); cp(v.nn1_g, pre->o[MDB_S_NB_NN1_G], D); cp(v.nn1_be, pre->o[MDB_S_NB_NN1_BE], D);
    cp(v.nn2_b, pre->o[MDB_S_NB_NN2_B], D); cp(v.gx_b, pre->o[MDB_S_NB_G1_B], D); cp(v.gx_t, pre->o[MDB_S_NB_GT_W], D);
    cp(v.cen_b, pre->o[MDB_S_NB_CEN_B], D);
    cp(v.gn_b[0], pre->o[MDB_S_EL_G1_B], 32); cp(v.gn_b[1], pre->o[MDB_S_ER_G1_B], 32);
    cp(v.nf_b[0], pre->o[MDB_S_EB_NFL_B], C); cp(v.nf_b[1], pre->o[MDB_S_EB_NFR_B], C);
  }
  if (hd && kind == 1) {
    cp(v.d1_b, hd->o[MDB_H_NDEC1_B], D); cp(v.d1_g, hd->o[MDB_H_NDEC1_G], D); cp(v.d1_be, hd->o[MDB_H_NDEC1_BE], D);
    cp(v.d2_b, hd->o[MDB_H_NDEC2_B], 32);
  }
}
