// Tensor-core (tcgen05) kernel for everything edge_kernel_b does besides the NodeBlock: RBF, edge_embs, and the two
// BondFFNs of the EdgeBlock (reference models/graph.py:354-357, 133-141, 278-284).  128 CSR-ordered edges per CTA,
// two threads per row (column halves), all Linears as split-bf16 MMAs into TMEM; only per-row epilogues run on
// the CUDA cores.  Writes e (for the NodeBlock / tail kernels) and accumulates SL / SR.
//
// Included by mdb_forward.cu inside its anonymous namespace (after tc_nodeblock.cuh).
#pragma once
#include "tc_pipe.cuh"

struct FfnSideVecs { float i1_b[128], i1_g[128], i1_be[128], i2_b[C], gt_w[32], g1_g[32], g1_be[32], g2_b[C]; };
struct FfnVecs {
  float ee_b[C], rbf_off[G], rbf_coeff[G];
  FfnSideVecs side[2];
};

struct TcFfnArgs {
  const uint8_t* tc_blob;
  TcOff tco;
  Tables tb;
  const int *left, *right;
  int n_nodes, n_edges;
  const float* pos;        // [N][3] positions of this block
  float rbf_lo, rbf_hi;
  float* ebuf;             // [E][64] out: e
  float* sl;               // [2][N][64] SL, SR accumulators (pre-zeroed)
  FfnVecs v;
};

template <int NC>
__device__ __forceinline__ void load_cols_tm(uint32_t taddr, float (&v)[NC]) {
  static_assert(NC == 16 || NC == 32 || NC == 64, "unsupported width");
  if constexpr (NC == 16) {
    tc::tmem_ld16(taddr, v);
  } else if constexpr (NC == 32) {
    tc::tmem_ld32(taddr, v);
  } else {
    uint32_t r0[32], r1[32];
    tc::tmem_ld32_issue(taddr, r0);
    tc::tmem_ld32_issue(taddr + 32, r1);
    tc::tmem_ld32_wait(r0); tc::tmem_ld32_wait(r1);
#pragma unroll
    for (int i = 0; i < 32; ++i) { v[i] = __uint_as_float(r0[i]); v[32 + i] = __uint_as_float(r1[i]); }
  }
}

// LayerNorm over a 2*NC-wide row split between two threads (NC columns each) + affine + ReLU, in place.
template <int NC>
__device__ __forceinline__ void ln_relu_part(float (&v)[NC], const float* __restrict__ gamma, const float* __restrict__ beta,
                                             float2* stat, int row, int half) {
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < NC; i += 4) { s4[0] += v[i]; s4[1] += v[i + 1]; s4[2] += v[i + 2]; s4[3] += v[i + 3]; }
  const float m_h = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.f / NC);
  float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < NC; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { const float d = v[i + u] - m_h; q4[u] = fmaf(d, d, q4[u]); }
  }
  const float q = (q4[0] + q4[1]) + (q4[2] + q4[3]);
  const float2 o = exchange_half(stat, row, half, m_h, q);
  const float mean = 0.5f * (m_h + o.x);
  const float dm = m_h - o.x;
  const float rstd = 1.f / sqrtf((q + o.y + dm * dm * (0.5f * NC)) * (1.f / (2 * NC)) + LN_EPS);
#pragma unroll
  for (int i = 0; i < NC; ++i) v[i] = fmaxf((v[i] - mean) * rstd * gamma[i] + beta[i], 0.f);
}

constexpr int FFN_O_LD = 68;    // fp32 row stride of the o tile: 272 B = 16 (mod 128)

__global__ void __launch_bounds__(TC_NB_THREADS, 1) tc_bondffn_fwd_kernel(const __grid_constant__ TcFfnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int KI = C + G;                                  // 80
  uint8_t* e_hi = smem_raw;                                  // E planes, K = 64: 2 x 16 KB
  uint8_t* e_lo = e_hi + tc::ROWS * C * 2;
  uint8_t* a_hi = e_lo + tc::ROWS * C * 2;                   // A planes, K = 128: 2 x 32 KB (also the K = 80 input planes)
  uint8_t* a_lo = a_hi + tc::ROWS * 128 * 2;
  uint8_t* g_hi = a_lo + tc::ROWS * 128 * 2;                 // gate planes, K = 32: 2 x 8 KB
  uint8_t* g_lo = g_hi + tc::ROWS * 32 * 2;
  uint8_t* stages = g_lo + tc::ROWS * 32 * 2;
  tc::PipeSmem* ps = reinterpret_cast<tc::PipeSmem*>(stages + tc::NSTAGE * tc::STAGE_SLOT);
  float2* stat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(ps) + 128);
  int* ls = reinterpret_cast<int*>(stat + 2 * tc::ROWS);
  float* o_tile = reinterpret_cast<float*>(a_hi);            // [128][FFN_O_LD] fp32 at the very end (A planes dead)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * tc::ROWS;
  const Tables& tb = a.tb;
  tc::Pipe p;
  tc::pipe_init<TC_NRW>(p, ps, stages);
  if (warp == TC_NRW) tc::tmem_alloc<256>(&ps->tmem_base);
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
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) x = *reinterpret_cast<const float4*>(tb.hedge + (size_t)q * C + half * 32 + i);
      v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
    }
    tc::store_a<KI, 32>(a_hi, a_lo, row, half * 32, v);
    const float dx = a.pos[ll * 3 + 0] - a.pos[rr * 3 + 0];
    const float dy = a.pos[ll * 3 + 1] - a.pos[rr * 3 + 1];
    const float dz = a.pos[ll * 3 + 2] - a.pos[rr * 3 + 2];
    const float d = fminf(fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), a.rbf_lo), a.rbf_hi);
    float g[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float u = d - a.v.rbf_off[half * 8 + k];
      g[k] = valid ? expf(a.v.rbf_coeff[half * 8 + k] * (u * u)) : 0.f;
    }
    tc::store_a<KI, 8>(a_hi, a_lo, row, C + half * 8, g);
    tc::rows_publish(p);
  }
  // G0: e = edge_embs(cat)                                                        graph.py:354-357
  tc::gemm<KI, C>(p, a_hi, a_lo, TCW_(EE), 0, false, true, true);
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    float v[32];
    tc::tmem_ld32(lane_base + half * 32, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += a.v.ee_b[half * 32 + i];
    if (valid) {
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(a.ebuf + (size_t)q * C + half * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    tc::store_a<C, 32>(e_hi, e_lo, row, half * 32, v);
    tc::rows_publish(p);
  }
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    const FfnSideVecs& sv = a.v.side[side];
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
      tc::rows_wait_acc(p);
      {
        float v[64];
        load_cols_tm<64>(lane_base + half * 64, v);
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(nl + i);
          v[i] *= t4.x; v[i + 1] *= t4.y; v[i + 2] *= t4.z; v[i + 3] *= t4.w;       // * node_linear(h_node)[.]
        }
        tc::store_a<128, 64>(a_hi, a_lo, row, half * 64, v);
      }
      {
        float g[16];
        load_cols_tm<16>(lane_base + 128 + half * 16, g);
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(gn + i);
          g[i] += t4.x + te * sv.gt_w[half * 16 + i];
          g[i + 1] += t4.y + te * sv.gt_w[half * 16 + i + 1];
          g[i + 2] += t4.z + te * sv.gt_w[half * 16 + i + 2];
          g[i + 3] += t4.w + te * sv.gt_w[half * 16 + i + 3];
        }
        ln_relu_part<16>(g, sv.g1_g + half * 16, sv.g1_be + half * 16, stat, row, half);
        tc::store_a<32, 16>(g_hi, g_lo, row, half * 16, g);
      }
      tc::rows_publish(p);
    }
    // inter_module.net.0 -> D[0:128] ; gate.net.3 -> D[128:192]                   graph.py:137,139
    tc::gemm<128, 128>(p, a_hi, a_lo, w_i1, 0, false, true, false);
    tc::gemm<32, C>(p, g_hi, g_lo, w_g2, 128, false, false, true);
    float sgg[32];
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      {
        float v[64];
        load_cols_tm<64>(lane_base + half * 64, v);
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] += sv.i1_b[half * 64 + i];
        ln_relu_part<64>(v, sv.i1_g + half * 64, sv.i1_be + half * 64, stat, row, half);
        tc::store_a<128, 64>(a_hi, a_lo, row, half * 64, v);
      }
      tc::tmem_ld32(lane_base + 128 + half * 32, sgg);
#pragma unroll
      for (int i = 0; i < 32; ++i) sgg[i] = tc::fast_sigmoid(sgg[i] + sv.g2_b[half * 32 + i]);
      tc::rows_publish(p);
    }
    // inter_module.net.3 -> D[0:64]                                                graph.py:137
    tc::gemm<128, C>(p, a_hi, a_lo, w_i2, 0, false, true, true);
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      float o[32];
      tc::tmem_ld32(lane_base + half * 32, o);
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = (o[i] + sv.i2_b[half * 32 + i]) * sgg[i];
      if (side == 0) {
        // bond_ffn_left output is scattered over right_node (unsorted): one 16-byte RED per 4 channels   graph.py:279
        if (valid) {
          float* dst = a.sl + (size_t)rr * C + half * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4) tc::red_add_v4(dst + i, o[i], o[i + 1], o[i + 2], o[i + 3]);
        }
      } else {
        // bond_ffn_right output is scattered over left_node = contiguous CSR runs: reduce in the tile   graph.py:283
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(o_tile + row * FFN_O_LD + half * 32 + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
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
      }
      if (side == 0) tc::rows_publish(p);     // accumulator drained: the right FFN's first GEMM may overwrite it
    }
  }
  __syncthreads();
  if (warp == TC_NRW) { __syncwarp(); tc::tmem_dealloc<256>(ps->tmem_base); }
}

constexpr size_t SMEM_TC_FFN = 2 * (size_t)tc::ROWS * C * 2 + 2 * (size_t)tc::ROWS * 128 * 2 + 2 * (size_t)tc::ROWS * 32 * 2
                               + tc::NSTAGE * tc::STAGE_SLOT + 128 + 2 * tc::ROWS * sizeof(float2) + tc::ROWS * sizeof(int) + 64;
static_assert(SMEM_TC_FFN <= 232448, "tc_bondffn_fwd_kernel exceeds the shared-memory limit");
static_assert((size_t)tc::ROWS * FFN_O_LD * 4 <= 2 * (size_t)tc::ROWS * 128 * 2, "o tile must fit in the A planes");
static_assert(2 * (size_t)tc::ROWS * (C + G) * 2 <= 2 * (size_t)tc::ROWS * 128 * 2, "K = 80 input planes must fit in the A planes");

inline void fill_ffn_vecs(FfnVecs& v, const float* blob_host, const BlkOff& off, const HeadOff& hoff) {
  auto cp = [&](float* dst, int slot, int n) { memcpy(dst, blob_host + off.o[slot], n * sizeof(float)); };
  cp(v.ee_b, MDB_S_EE_B, C);
  memcpy(v.rbf_off, blob_host + hoff.o[MDB_H_RBF_OFFSET], G * sizeof(float));
  memcpy(v.rbf_coeff, blob_host + hoff.o[MDB_H_RBF_COEFF], G * sizeof(float));
  const int base[2] = {MDB_S_EL_BL_W, MDB_S_ER_BL_W};
  for (int s = 0; s < 2; ++s) {
    const int d = base[s] - MDB_S_EL_BL_W;      // slot distance between the left and right groups
    cp(v.side[s].i1_b, MDB_S_EL_I1_B + d, 128); cp(v.side[s].i1_g, MDB_S_EL_I1_G + d, 128); cp(v.side[s].i1_be, MDB_S_EL_I1_BE + d, 128);
    cp(v.side[s].i2_b, MDB_S_EL_I2_B + d, C); cp(v.side[s].gt_w, MDB_S_EL_GT_W + d, 32);
    cp(v.side[s].g1_g, MDB_S_EL_G1_G + d, 32); cp(v.side[s].g1_be, MDB_S_EL_G1_BE + d, 32); cp(v.side[s].g2_b, MDB_S_EL_G2_B + d, C);
  }
}
