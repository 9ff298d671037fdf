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
  long long* dbg;          // optional [grid][32] clock64 stamps of row thread 0 (tools/tc_phase_times.py --ffn)
};
#define FFN_STAMP(i) do { if (a.dbg && tid == 0) a.dbg[(size_t)blockIdx.x * 32 + (i)] = clock64(); } while (0)

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
    for (int i = 0; i < 32; ++i) { v[i] = tc::acc_f(r0[i]); v[32 + i] = tc::acc_f(r1[i]); }
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

// NC floats of a gathered table row -> registers (issued BEFORE the accumulator wait so the L2 latency hides behind the GEMM)
template <int NC>
__device__ __forceinline__ void gather_n(const float* __restrict__ src, float (&v)[NC]) {
#pragma unroll
  for (int i = 0; i < NC; i += 4) {
    const float4 t4 = *reinterpret_cast<const float4*>(src + i);
    v[i] = t4.x; v[i + 1] = t4.y; v[i + 2] = t4.z; v[i + 3] = t4.w;
  }
}

constexpr int FFN_O_LD = 68;    // fp32 row stride of the o tile: 272 B = 16 (mod 128)

template <bool XF>   // cross-first accumulation order (tc_pipe.cuh)
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
  FFN_STAMP(0);
  const int q0 = blockIdx.x * tc::ROWS;
  const Tables& tb = a.tb;
  tc::PipeT<tc::NSTAGE, XF> p;
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
    FFN_STAMP(1);
  }
  // G0: e = edge_embs(cat)                                                        graph.py:354-357
  tc::gemm<KI, C>(p, a_hi, a_lo, TCW_(EE), 0, false, true, true);
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    FFN_STAMP(2);
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
    FFN_STAMP(3);
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
      float nlv[64], gnv[16];
      gather_n<64>(nl, nlv); gather_n<16>(gn, gnv);
      tc::rows_wait_acc(p);
      FFN_STAMP(4 + side * 8);
      {
        float v[64];
        load_cols_tm<64>(lane_base + half * 64, v);
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] *= nlv[i];                                 // * node_linear(h_node)[.]
        tc::store_a<128, 64>(a_hi, a_lo, row, half * 64, v);
      }
      {
        float g[16];
        load_cols_tm<16>(lane_base + 128 + half * 16, g);
#pragma unroll
        for (int i = 0; i < 16; ++i) g[i] += gnv[i] + te * sv.gt_w[half * 16 + i];
        ln_relu_part<16>(g, sv.g1_g + half * 16, sv.g1_be + half * 16, stat, row, half);
        tc::store_a<32, 16>(g_hi, g_lo, row, half * 16, g);
      }
      tc::rows_publish(p);
      FFN_STAMP(5 + side * 8);
    }
    // inter_module.net.0 -> D[0:128] ; gate.net.3 -> D[128:192]                   graph.py:137,139
    tc::gemm<128, 128>(p, a_hi, a_lo, w_i1, 0, false, true, false);
    tc::gemm<32, C>(p, g_hi, g_lo, w_g2, 128, false, false, true);
    float sgg[32];
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      FFN_STAMP(6 + side * 8);
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
      FFN_STAMP(7 + side * 8);
    }
    // inter_module.net.3 -> D[0:64]                                                graph.py:137
    tc::gemm<128, C>(p, a_hi, a_lo, w_i2, 0, false, true, true);
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      FFN_STAMP(8 + side * 8);
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
      FFN_STAMP(9 + side * 8);
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

// =================================================================================================================
// Backward of the two BondFFNs + the edge_embs backward that closes a block (same math as bwd_edge_bondffn_kernel)
// =================================================================================================================
struct TcFfnBwdArgs {
  const uint8_t* tc_blob;
  TcOff tco;
  Tables tb;
  const int *left, *right;
  int n_nodes, n_edges;
  const float* e;          // [E][64] saved e_i
  const float *dul, *dur;  // [N][64] d/d SL (= d/d fl), d/d SR (= d/d fr)
  float* dnl;              // [2][N][128] scatter targets (pre-zeroed)
  float* dgn;              // [2][N][32]
  const float* de_in;      // [E][64] d/d e accumulated by the tail and NodeBlock kernels
  float* dh;               // [E][64] out: d/d h_edge entering the block
  float* dg;               // [E][16] d/d rbf features, accumulated (+=)
  FfnVecs v;
};

// One row of the backward of y = relu(LN(x) * g + b) for a row split between two threads (NC columns each):
// d (grad wrt y) <- grad wrt x; xh = this thread's normalised activations.
template <int NC>
__device__ __forceinline__ void ln_bwd_part(float (&d)[NC], const float (&xh)[NC], float rstd,
                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                            float2* stat, int row, int half) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const float dxh = (xh[i] * gamma[i] + beta[i] > 0.f) ? d[i] * gamma[i] : 0.f;
    d[i] = dxh;
    s1 += dxh;
    s2 = fmaf(dxh, xh[i], s2);
  }
  const float2 o = exchange_half(stat, row, half, s1, s2);
  const float m1 = (s1 + o.x) * (1.f / (2 * NC)), m2 = (s2 + o.y) * (1.f / (2 * NC));
#pragma unroll
  for (int i = 0; i < NC; ++i) d[i] = rstd * (d[i] - m1 - xh[i] * m2);
}

// LayerNorm statistics of a 2*NC-wide row split between two threads; v <- xhat, returns rstd
template <int NC>
__device__ __forceinline__ float ln_xhat_part(float (&v)[NC], float2* stat, int row, int half, float* mean_out = nullptr) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NC; ++i) s += v[i];
  const float m_h = s * (1.f / NC);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NC; ++i) { const float d = v[i] - m_h; q = fmaf(d, d, q); }
  const float2 o = exchange_half(stat, row, half, m_h, q);
  const float mean = 0.5f * (m_h + o.x);
  const float dm = m_h - o.x;
  const float rstd = 1.f / sqrtf((q + o.y + dm * dm * (0.5f * NC)) * (1.f / (2 * NC)) + LN_EPS);
#pragma unroll
  for (int i = 0; i < NC; ++i) v[i] = (v[i] - mean) * rstd;
  if (mean_out) *mean_out = mean;
  return rstd;
}

__global__ void __launch_bounds__(TC_NB_THREADS, 1) tc_bondffn_bwd_kernel(const __grid_constant__ TcFfnBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* e_hi = smem_raw;                                  // E planes, K = 64
  uint8_t* e_lo = e_hi + tc::ROWS * C * 2;
  uint8_t* a_hi = e_lo + tc::ROWS * C * 2;                   // A planes, K = 128 (64 KB) ...
  uint8_t* a_lo = a_hi + tc::ROWS * 128 * 2;
  uint8_t* s1_hi = a_hi;                                     // ... or two K = 64 plane pairs (2 x 32 KB) aliased on them
  uint8_t* s1_lo = s1_hi + tc::ROWS * C * 2;
  uint8_t* s2_hi = s1_lo + tc::ROWS * C * 2;
  uint8_t* s2_lo = s2_hi + tc::ROWS * C * 2;
  uint8_t* g_hi = a_lo + tc::ROWS * 128 * 2;                 // gate planes, K = 32
  uint8_t* g_lo = g_hi + tc::ROWS * 32 * 2;
  uint8_t* stages = g_lo + tc::ROWS * 32 * 2;
  tc::PipeSmem* ps = reinterpret_cast<tc::PipeSmem*>(stages + tc::NSTAGE * tc::STAGE_SLOT);
  float2* stat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(ps) + 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * tc::ROWS;
  const Tables& tb = a.tb;
  tc::Pipe p;
  tc::pipe_init<TC_NRW>(p, ps, stages);
  if (warp == TC_NRW) tc::tmem_alloc<512>(&ps->tmem_base);
  const int row = (warp & 3) * 32 + lane;
  const int half = (warp >> 2) & 1;
  const int q = q0 + row;
  const bool valid = p.role == 0 && q < a.n_edges;
  const int ll = valid ? a.left[q] : 0, rr = valid ? a.right[q] : 0;
  const float te = valid ? tb.te[q] : 0.f;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t lane_base = ps->tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  // TMEM column map
  constexpr uint32_t T_BL = 0, T_A5 = 128, T_I2 = 256, T_G1 = 320, T_GG = 352, T_DR6 = 416, T_DE = 448, T_WIDE = 256;

  if (p.role == 0) {   // e tile -> E planes
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) x = *reinterpret_cast<const float4*>(a.e + (size_t)q * C + half * 32 + i);
      v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
    }
    tc::store_a<C, 32>(e_hi, e_lo, row, half * 32, v);
    tc::rows_publish(p);
  }
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    const FfnSideVecs& sv = a.v.side[side];
    const int node = side ? rr : ll;                         // node feeding this FFN
    const int other = side ? ll : rr;                        // node its output was scattered to
    const float* nl = (side ? tb.nlr : tb.nll) + (size_t)node * 128 + half * 64;
    const float* gn = (side ? tb.gnr : tb.gnl) + (size_t)node * 32 + half * 16;
    float* dnl = a.dnl + (size_t)side * a.n_nodes * 128 + (size_t)node * 128 + half * 64;
    float* dgn = a.dgn + (size_t)side * a.n_nodes * 32 + (size_t)node * 32 + half * 16;
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
      float nlv[64], gnv[16];
      gather_n<64>(nl, nlv); gather_n<16>(gn, gnv);
      tc::rows_wait_acc(p);
      {
        float v[64];
        load_cols_tm<64>(lane_base + T_BL + half * 64, v);
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] *= nlv[i];
        tc::store_a<128, 64>(a_hi, a_lo, row, half * 64, v);
      }
      load_cols_tm<16>(lane_base + T_G1 + half * 16, xh6);
#pragma unroll
      for (int i = 0; i < 16; ++i) xh6[i] += gnv[i] + te * sv.gt_w[half * 16 + i];
      rstd6 = ln_xhat_part<16>(xh6, stat, row, half);
      float r6[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) r6[i] = fmaxf(xh6[i] * sv.g1_g[half * 16 + i] + sv.g1_be[half * 16 + i], 0.f);
      tc::store_a<32, 16>(g_hi, g_lo, row, half * 16, r6);
      tc::rows_publish(p);
    }
    tc::gemm<128, 128>(p, a_hi, a_lo, w_i1, T_A5, false, true, false);
    tc::gemm<32, C>(p, g_hi, g_lo, w_g2, T_GG, false, false, true);
    float sgg[32], mean5 = 0.f, rstd5 = 0.f;
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      {
        float v[64];
        load_cols_tm<64>(lane_base + T_A5 + half * 64, v);
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] += sv.i1_b[half * 64 + i];
        rstd5 = ln_xhat_part<64>(v, stat, row, half, &mean5);
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = fmaxf(v[i] * sv.i1_g[half * 64 + i] + sv.i1_be[half * 64 + i], 0.f);
        tc::store_a<128, 64>(a_hi, a_lo, row, half * 64, v);
      }
      tc::tmem_ld32(lane_base + T_GG + half * 32, sgg);
#pragma unroll
      for (int i = 0; i < 32; ++i) sgg[i] = tc::fast_sigmoid(sgg[i] + sv.g2_b[half * 32 + i]);
      tc::rows_publish(p);
    }
    tc::gemm<128, C>(p, a_hi, a_lo, w_i2, T_I2, false, true, true);
    // ---- d o = DUL[r] (left FFN, scattered over right) / DUR[l] (right FFN, scattered over left)
    if (p.role == 0) {
      const float* du = (side ? a.dur : a.dul) + (size_t)other * C + half * 32;
      float duv[32];
      gather_n<32>(du, duv);                                  // (`other` is 0 for padding rows: a valid address)
      tc::rows_wait_acc(p);
      float i2[32];
      tc::tmem_ld32(lane_base + T_I2 + half * 32, i2);
      float dgg[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float dd = valid ? duv[i] : 0.f;
        const float s = sgg[i];
        const float val = i2[i] + sv.i2_b[half * 32 + i];
        dgg[i] = dd * val * s * (1.f - s);
        i2[i] = dd * s;                                       // d i2
      }
      tc::store_a<C, 32>(s1_hi, s1_lo, row, half * 32, dgg);   // (the K = 128 A planes are dead: r5 was consumed)
      tc::store_a<C, 32>(s2_hi, s2_lo, row, half * 32, i2);
      tc::rows_publish(p);
    }
    tc::gemm<C, 32>(p, s1_hi, s1_lo, t_g2, T_DR6, false, true, false);        // d relu6
    tc::gemm<C, 128>(p, s2_hi, s2_lo, t_i2, T_WIDE, false, false, true);      // d relu5
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      {  // gate branch
        float d6[16];
        load_cols_tm<16>(lane_base + T_DR6 + half * 16, d6);
        ln_bwd_part<16>(d6, xh6, rstd6, sv.g1_g + half * 16, sv.g1_be + half * 16, stat, row, half);
        if (valid) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) tc::red_add_v4(dgn + i, d6[i], d6[i + 1], d6[i + 2], d6[i + 3]);
        }
        tc::store_a<32, 16>(g_hi, g_lo, row, half * 16, d6);
      }
      {  // inter branch: xhat5 recomputed from the parked pre-LN accumulator
        float xh[64], d5[64];
        load_cols_tm<64>(lane_base + T_A5 + half * 64, xh);
#pragma unroll
        for (int i = 0; i < 64; ++i) xh[i] = (xh[i] + sv.i1_b[half * 64 + i] - mean5) * rstd5;
        load_cols_tm<64>(lane_base + T_WIDE + half * 64, d5);
        ln_bwd_part<64>(d5, xh, rstd5, sv.i1_g + half * 64, sv.i1_be + half * 64, stat, row, half);
        tc::store_a<128, 64>(a_hi, a_lo, row, half * 64, d5);     // (the K = 64 planes aliased here are dead)
      }
      tc::rows_publish(p);
    }
    tc::gemm<32, C>(p, g_hi, g_lo, t_gb, T_DE, side != 0, true, false);        // de (+)= d a6 W_gb^T
    tc::gemm<128, 128>(p, a_hi, a_lo, t_i1, T_WIDE, false, false, true);      // d inter0
    if (p.role == 0) {
      tc::rows_wait_acc(p);
      float di[64], bl[64];
      load_cols_tm<64>(lane_base + T_WIDE + half * 64, di);
      load_cols_tm<64>(lane_base + T_BL + half * 64, bl);
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        if (valid) tc::red_add_v4(dnl + i, di[i] * bl[i], di[i + 1] * bl[i + 1], di[i + 2] * bl[i + 2], di[i + 3] * bl[i + 3]);
        const float4 t4 = *reinterpret_cast<const float4*>(nl + i);
        di[i] *= t4.x; di[i + 1] *= t4.y; di[i + 2] *= t4.z; di[i + 3] *= t4.w;       // d bond_linear(e)
      }
      tc::store_a<128, 64>(a_hi, a_lo, row, half * 64, di);
      tc::rows_publish(p);
    }
    tc::gemm<128, C>(p, a_hi, a_lo, t_bl, T_DE, true, true, side == 1);        // de += d bl W_bl^T
  }
  // ---- total d e -> d h_in = de W_EE[:, :64]^T..., d rbf += de W_EE[:, 64:80]
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    float de[32];
    tc::tmem_ld32(lane_base + T_DE + half * 32, de);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) x = *reinterpret_cast<const float4*>(a.de_in + (size_t)q * C + half * 32 + i);
      de[i] += x.x; de[i + 1] += x.y; de[i + 2] += x.z; de[i + 3] += x.w;
    }
    tc::store_a<C, 32>(s1_hi, s1_lo, row, half * 32, de);
    tc::rows_publish(p);
  }
  tc::gemm<C, C>(p, s1_hi, s1_lo, TCW_(BT_EEH), 0, false, true, false);
  tc::gemm<C, 32>(p, s1_hi, s1_lo, TCW_(BT_EEG), 64, false, false, true);
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    float v[32];
    tc::tmem_ld32(lane_base + half * 32, v);
    if (valid) {
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(a.dh + (size_t)q * C + half * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    if (half == 0) {
      float g[16];
      load_cols_tm<16>(lane_base + 64, g);
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

constexpr size_t SMEM_TC_FFN_BWD = 2 * (size_t)tc::ROWS * C * 2 + 2 * (size_t)tc::ROWS * 128 * 2 + 2 * (size_t)tc::ROWS * 32 * 2
                                   + tc::NSTAGE * tc::STAGE_SLOT + 128 + 2 * tc::ROWS * sizeof(float2) + 64;
static_assert(SMEM_TC_FFN_BWD <= 232448, "tc_bondffn_bwd_kernel exceeds the shared-memory limit");
