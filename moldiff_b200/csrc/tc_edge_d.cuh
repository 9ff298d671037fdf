// Tensor-core (tcgen05) version of edge_kernel_d: EdgeBlock tail (-> new h_edge, graph.py:286-294) and the PosUpdate
// per-edge path (-> force -> pos, graph.py:384-396).  128 CSR-ordered edges per CTA, two threads per row.
//
// Included by mdb_forward.cu inside its anonymous namespace (after tc_bondffn.cuh).
#pragma once
#include "tc_pipe.cuh"

struct EdgeDVecs {
  float self_b[C], ln_g[C], ln_be[C], out_b[C];                    // EdgeBlock tail
  float i1_b[D], i1_g[D], i1_be[D], i2_w[D];                       // PosUpdate edge_lin.inter_module
  float gt_w[32], g1_b[32], g1_g[32], g1_be[32], g2_w[32];         // PosUpdate edge_lin.gate
  float i2_b, g2_b;
};

struct TcEdgeDArgs {
  const uint8_t* tc_blob;
  TcOff tco;
  Tables tb;
  const int *left, *right;
  int n_nodes, n_edges;
  int update_pos;
  const float* ebuf;       // [E][64] e of this block
  const float* sl;         // [2][N][64] SL, SR
  const float *fl, *fr;    // [N][64] node_ffn_{left,right}(h_node)
  const float* pos_cur;
  float* pos_nxt;
  EdgeDVecs v;
};

template <bool XF>   // cross-first accumulation order (tc_pipe.cuh)
__global__ void __launch_bounds__(TC_NB_THREADS, 1) tc_edge_d_kernel(const __grid_constant__ TcEdgeDArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* h_hi = smem_raw;                                  // region A: e planes, then new-h_edge planes (K = 64)
  uint8_t* h_lo = h_hi + tc::ROWS * C * 2;
  uint8_t* p_hi = h_lo + tc::ROWS * C * 2;                   // region B: LN(u) planes, then lf*rf planes (K = 64)
  uint8_t* p_lo = p_hi + tc::ROWS * C * 2;
  uint8_t* x_hi = smem_raw;                                  // K = 256 planes: overlay regions A, B (dead by then) + 64 KB
  uint8_t* x_lo = x_hi + tc::ROWS * D * 2;
  uint8_t* stages = smem_raw + 2 * (size_t)tc::ROWS * D * 2;
  tc::PipeSmem* ps = reinterpret_cast<tc::PipeSmem*>(stages + tc::NSTAGE * tc::STAGE_SLOT);
  float2* stat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(ps) + 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * tc::ROWS;
  const Tables& tb = a.tb;
  tc::PipeT<tc::NSTAGE, XF> p;
  tc::pipe_init<TC_NRW>(p, ps, stages);
  if (warp == TC_NRW) tc::tmem_alloc<512>(&ps->tmem_base);
  const int row = (warp & 3) * 32 + lane;
  const int half = (warp >> 2) & 1;
  const int q = q0 + row;
  const bool valid = p.role == 0 && q < a.n_edges;
  const int ll = valid ? a.left[q] : 0, rr = valid ? a.right[q] : 0;
  const float te = valid ? tb.te[q] : 0.f;
  const int c32 = half * 32;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t lane_base = ps->tmem_base + ((uint32_t)((warp & 3) * 32) << 16);

  float ev[32];                                              // this thread's 32 columns of e (kept for the residual)
  if (p.role == 0) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) x = *reinterpret_cast<const float4*>(a.ebuf + (size_t)q * C + c32 + i);
      ev[i] = x.x; ev[i + 1] = x.y; ev[i + 2] = x.z; ev[i + 3] = x.w;
    }
    tc::store_a<C, 32>(h_hi, h_lo, row, c32, ev);
    tc::rows_publish(p);
  }
  // self_ffn(e)                                                                    graph.py:290
  tc::gemm<C, C>(p, h_hi, h_lo, TCW_(EB_SELF), 0, false, true, true);
  if (p.role == 0) {
    const float* t0 = a.sl + (size_t)ll * C + c32;                          // scatter_sum(msg_left, right)[left]
    const float* t1 = a.sl + (size_t)a.n_nodes * C + (size_t)rr * C + c32;  // scatter_sum(msg_right, left)[right]
    const float* t2 = a.fl + (size_t)ll * C + c32;
    const float* t3 = a.fr + (size_t)rr * C + c32;
    float gsum[32];                                          // the four gathered rows, summed while the GEMM runs
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 x0 = *reinterpret_cast<const float4*>(t0 + i), x1 = *reinterpret_cast<const float4*>(t1 + i);
      const float4 x2 = *reinterpret_cast<const float4*>(t2 + i), x3 = *reinterpret_cast<const float4*>(t3 + i);
      gsum[i] = x0.x + x1.x + x2.x + x3.x; gsum[i + 1] = x0.y + x1.y + x2.y + x3.y;
      gsum[i + 2] = x0.z + x1.z + x2.z + x3.z; gsum[i + 3] = x0.w + x1.w + x2.w + x3.w;
    }
    tc::rows_wait_acc(p);
    float u[32];
    tc::tmem_ld32(lane_base + c32, u);
#pragma unroll
    for (int i = 0; i < 32; ++i) u[i] += a.v.self_b[c32 + i] + gsum[i];
    ln_relu_part<32>(u, a.v.ln_g + c32, a.v.ln_be + c32, stat, row, half);
    tc::store_a<C, 32>(p_hi, p_lo, row, c32, u);
    tc::rows_publish(p);
  }
  // out_transform ; h_edge = e + EdgeBlock(.)                                       graph.py:294,362
  tc::gemm<C, C>(p, p_hi, p_lo, TCW_(EB_OUT), 0, false, true, true);
  float rel = 0.f, dist = 1.f;                               // half 0: x / half 1: y, and z handled by half 0 too
  float relz = 0.f;
  if (p.role == 0) {
    float pv[32];                                            // left_feat * right_feat (PosUpdate): gathered before the wait
    if (a.update_pos) {
      const float* lf = tb.lf + (size_t)ll * C + c32;
      const float* rf = tb.rf + (size_t)rr * C + c32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 x = *reinterpret_cast<const float4*>(lf + i), y = *reinterpret_cast<const float4*>(rf + i);
        pv[i] = x.x * y.x; pv[i + 1] = x.y * y.y; pv[i + 2] = x.z * y.z; pv[i + 3] = x.w * y.w;
      }
    }
    tc::rows_wait_acc(p);
    float h[32];
    tc::tmem_ld32(lane_base + c32, h);
#pragma unroll
    for (int i = 0; i < 32; ++i) h[i] += a.v.out_b[c32 + i] + ev[i];
    if (valid) {
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(tb.hedge + (size_t)q * C + c32 + i) = make_float4(h[i], h[i + 1], h[i + 2], h[i + 3]);
    }
    if (a.update_pos) {
      tc::store_a<C, 32>(h_hi, h_lo, row, c32, h);                          // e planes are dead: reuse for new h_edge
      tc::store_a<C, 32>(p_hi, p_lo, row, c32, pv);
      const float dx = a.pos_cur[ll * 3 + 0] - a.pos_cur[rr * 3 + 0];
      const float dy = a.pos_cur[ll * 3 + 1] - a.pos_cur[rr * 3 + 1];
      const float dz = a.pos_cur[ll * 3 + 2] - a.pos_cur[rr * 3 + 2];
      dist = sqrtf(dx * dx + dy * dy + dz * dz);
      rel = half ? dy : dx;
      relz = dz;
      tc::rows_publish(p);
    }
  }
  if (a.update_pos) {
    // gate.net.0: bond columns + node (= lf*rf) columns -> D[0:32]                  graph.py:139
    tc::gemm<C, 32>(p, h_hi, h_lo, TCW_(PU_GB), 0, false, true, false);
    tc::gemm<C, 32>(p, p_hi, p_lo, TCW_(PU_GN), 0, true, false, true);
    float gsum = 0.f;                                        // this thread's partial of gate.net.3 (32 -> 1)
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      float g[16];
      load_cols_tm<16>(lane_base + half * 16, g);
#pragma unroll
      for (int i = 0; i < 16; ++i) g[i] += a.v.g1_b[half * 16 + i] + te * a.v.gt_w[half * 16 + i];
      ln_relu_part<16>(g, a.v.g1_g + half * 16, a.v.g1_be + half * 16, stat, row, half);
#pragma unroll
      for (int i = 0; i < 16; ++i) gsum = fmaf(g[i], a.v.g2_w[half * 16 + i], gsum);
      tc::rows_publish(p);
    }
    // bond_linear(h_edge) -> D[0:256] ; node_linear(lf*rf) -> D[256:512]            graph.py:134-136
    tc::gemm<C, D>(p, h_hi, h_lo, TCW_(PU_PB), 0, false, true, false);
    tc::gemm<C, D>(p, p_hi, p_lo, TCW_(PU_PN), 256, false, false, true);
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      float v[128];
      load_half_row(lane_base + half * 128, v);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float w[32];
        tc::tmem_ld32(lane_base + 256 + half * 128 + cc * 32, w);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[cc * 32 + i] *= w[i];
      }
      store_half_row_a(x_hi, x_lo, row, half * 128, v);      // overlays the K = 64 planes: their GEMMs are complete
      tc::rows_publish(p);
    }
    // inter_module.net.0                                                             graph.py:137
    tc::gemm<D, D>(p, x_hi, x_lo, TCW_(PU_I1), 0, false, true, true);
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      float v[128];
      load_half_row(lane_base + half * 128, v);
      add_vec128(v, a.v.i1_b + half * 128);
      ln_relu_half(v, a.v.i1_g + half * 128, a.v.i1_be + half * 128, stat, row, half);
      float isum = 0.f;
#pragma unroll
      for (int i = 0; i < 128; ++i) isum = fmaf(v[i], a.v.i2_w[half * 128 + i], isum);
      const float2 o = exchange_half(stat, row, half, isum, gsum);
      const float w_inter = isum + o.x + a.v.i2_b;
      const float gate = gsum + o.y + a.v.g2_b;
      const float w = w_inter * (1.f / (1.f + expf(-gate)));
      // force = w * rel / d / (d + 1) ; delta_pos = scatter_sum(force, left)          graph.py:393-394
      if (valid) {
        atomicAdd(a.pos_nxt + (size_t)ll * 3 + half, w * rel / dist / (dist + 1.f));
        if (half == 0) atomicAdd(a.pos_nxt + (size_t)ll * 3 + 2, w * relz / dist / (dist + 1.f));
      }
    }
  }
  __syncthreads();
  if (warp == TC_NRW) { __syncwarp(); tc::tmem_dealloc<512>(ps->tmem_base); }
}

constexpr size_t SMEM_TC_EDGE_D = 2 * (size_t)tc::ROWS * D * 2 + tc::NSTAGE * tc::STAGE_SLOT + 128
                                  + 2 * tc::ROWS * sizeof(float2) + 64;
static_assert(SMEM_TC_EDGE_D <= 232448, "tc_edge_d_kernel exceeds the shared-memory limit");

inline void fill_edge_d_vecs(EdgeDVecs& v, const float* blob_host, const BlkOff& off, bool update_pos) {
  auto cp = [&](float* dst, int slot, int n) { memcpy(dst, blob_host + off.o[slot], n * sizeof(float)); };
  cp(v.self_b, MDB_S_EB_SELF_B, C); cp(v.ln_g, MDB_S_EB_LN_G, C); cp(v.ln_be, MDB_S_EB_LN_BE, C); cp(v.out_b, MDB_S_EB_OUT_B, C);
  if (update_pos) {
    cp(v.i1_b, MDB_S_PU_I1_B, D); cp(v.i1_g, MDB_S_PU_I1_G, D); cp(v.i1_be, MDB_S_PU_I1_BE, D); cp(v.i2_w, MDB_S_PU_I2_W, D);
    cp(v.gt_w, MDB_S_PU_GT_W, 32); cp(v.g1_b, MDB_S_PU_G1_B, 32); cp(v.g1_g, MDB_S_PU_G1_G, 32); cp(v.g1_be, MDB_S_PU_G1_BE, 32);
    cp(v.g2_w, MDB_S_PU_G2_W, 32);
    v.i2_b = blob_host[off.o[MDB_S_PU_I2_B]];
    v.g2_b = blob_host[off.o[MDB_S_PU_G2_B]];
  }
}
