// Tensor-core (tcgen05) kernels for the NodeBlock per-edge path -- the dominant FLOPs of the denoiser and of the
// bond predictor (reference models/graph.py:42-50: edge_net MLP, msg_net, gate MLP on every directed edge).
//
// One CTA = 128 CSR-ordered edges; thread t of the 4 row warps owns edge row t end to end: its accumulator row
// lives in TMEM lane t, LayerNorm statistics are thread-local (no shuffles), and it writes its own row of the
// next layer's split-bf16 A operand.  msg_net's output never leaves TMEM (columns 0..255) while the gate chain
// runs in columns 256..511; out = msg * sigmoid(gate) is reduced over each node's contiguous run of edges through
// shared memory and leaves the CTA as one RED per (node, channel).
//
// Included by mdb_forward.cu inside its anonymous namespace.
#pragma once
#include "tc_pipe.cuh"

struct TcOff { int64_t o[MDB_NUM_TC_SLOTS]; };   // byte offsets into the tc blob
#define TCW_(slot) (a.tc_blob + a.tco.o[MDB_T_##slot])

// Per-column parameter vectors of the NodeBlock edge path, passed BY VALUE: they then live in the constant bank
// and every row thread reads them with warp-uniform constant loads instead of 96 LDG per LayerNorm epilogue.
struct NbVecs {
  float en1_b[D], en1_g[D], en1_be[D], en2_b[D], msg_b[D], g1_g[D], g1_be[D], g2_b[D];
};

struct TcNbArgs {
  const float* blob;          // fp32 blob: biases / LayerNorm parameters
  const uint8_t* tc_blob;     // split-bf16 weight stage images
  BlkOff off;
  TcOff tco;
  Tables tb;
  const int *left, *right;
  int n_nodes, n_edges;
  const float* ebuf;          // [E][64] e = edge_embs(cat(h_edge, rbf)) (written by edge_kernel_b)
  NbVecs v;
  long long* dbg;             // optional [grid][32] clock64 stamps of row thread 0 (phase timing, tools/tc_phase_times.py)
};
#define TC_STAMP(i) do { if (a.dbg && tid == 0) a.dbg[(size_t)blockIdx.x * 32 + (i)] = clock64(); } while (0)

constexpr int OUT_LD = 260;   // fp32 row stride of the out tile in smem: 1040 B = 16 (mod 128) -> conflict-free rows
constexpr int TC_NRW = 8;     // row warps: thread (w, t) owns row 32 * (w % 4) + t, columns [128 * (w / 4), +128)
constexpr int TC_NB_THREADS = (TC_NRW + 2) * 32;

// Row-thread helpers ------------------------------------------------------------------------------------------
// v[0:128] <- this thread's half row of the accumulator (TMEM lane = row, columns taddr .. +128): four 32-column
// loads in flight, one wait.
__device__ __forceinline__ void load_half_row(uint32_t taddr, float (&v)[128]) {
  uint32_t r0[32], r1[32], r2[32], r3[32];
  tc::tmem_ld32_issue(taddr, r0);
  tc::tmem_ld32_issue(taddr + 32, r1);
  tc::tmem_ld32_issue(taddr + 64, r2);
  tc::tmem_ld32_issue(taddr + 96, r3);
  tc::tmem_ld32_wait(r0); tc::tmem_ld32_wait(r1); tc::tmem_ld32_wait(r2); tc::tmem_ld32_wait(r3);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    v[i] = tc::acc_f(r0[i]); v[32 + i] = tc::acc_f(r1[i]);
    v[64 + i] = tc::acc_f(r2[i]); v[96 + i] = tc::acc_f(r3[i]);
  }
}

// v += vec[0:128]  (vec: kernel-argument array in the constant bank, warp-uniform address)
__device__ __forceinline__ void add_vec128(float (&v)[128], const float* __restrict__ vec) {
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] += vec[i];
}
// v += row[0:128]  (row: this thread's own gathered table row)
__device__ __forceinline__ void add_row128(float (&v)[128], const float* __restrict__ row) {
#pragma unroll
  for (int i = 0; i < 128; i += 4) {
    const float4 b = *reinterpret_cast<const float4*>(row + i);
    v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
  }
}

// LayerNorm over the full 256-wide row whose two halves live in two threads: each half computes (mean, M2) of its
// 128 values two-pass, the halves are merged exactly (Chan et al.) through `stat`, then v <- relu(LN(v) * g + b).
__device__ __forceinline__ void ln_relu_half(float (&v)[128], const float* __restrict__ gamma,
                                             const float* __restrict__ beta, float2* stat, int row, int half) {
  float s4[4] = {0.f, 0.f, 0.f, 0.f};                    // 4 independent chains: 32-deep instead of 128-deep
#pragma unroll
  for (int i = 0; i < 128; i += 4) { s4[0] += v[i]; s4[1] += v[i + 1]; s4[2] += v[i + 2]; s4[3] += v[i + 3]; }
  const float m_h = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.f / 128.f);
  float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 128; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { const float d = v[i + u] - m_h; q4[u] = fmaf(d, d, q4[u]); }
  }
  const float q = (q4[0] + q4[1]) + (q4[2] + q4[3]);
  stat[half * tc::ROWS + row] = make_float2(m_h, q);
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const float2 o = stat[(half ^ 1) * tc::ROWS + row];
  const float mean = 0.5f * (m_h + o.x);
  const float dm = m_h - o.x;
  const float rstd = 1.f / sqrtf((q + o.y + dm * dm * 64.f) * (1.f / 256.f) + LN_EPS);
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] = fmaxf((v[i] - mean) * rstd * gamma[i] + beta[i], 0.f);
}

// statistics only: (mean, rstd) of the full 256-wide row (same merge as ln_relu_half); the caller normalises while it
// stores (sliced protocol, tc_pipe.cuh)
__device__ __forceinline__ float2 ln_stats_half(const float (&v)[128], float2* stat, int row, int half) {
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 128; i += 4) { s4[0] += v[i]; s4[1] += v[i + 1]; s4[2] += v[i + 2]; s4[3] += v[i + 3]; }
  const float m_h = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.f / 128.f);
  float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 128; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { const float d = v[i + u] - m_h; q4[u] = fmaf(d, d, q4[u]); }
  }
  const float q = (q4[0] + q4[1]) + (q4[2] + q4[3]);
  stat[half * tc::ROWS + row] = make_float2(m_h, q);
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const float2 o = stat[(half ^ 1) * tc::ROWS + row];
  const float mean = 0.5f * (m_h + o.x);
  const float dm = m_h - o.x;
  return make_float2(mean, 1.f / sqrtf((q + o.y + dm * dm * 64.f) * (1.f / 256.f) + LN_EPS));
}

// this thread's 128 values -> columns [k0, k0 + 128) of row r of the K = 256 A planes
__device__ __forceinline__ void store_half_row_a(uint8_t* a_hi, uint8_t* a_lo, int r, int k0, const float (&v)[128]) {
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = v[c * 8 + i];
    uint4 hi, lo;
    tc::split8(x, hi, lo);
    const uint32_t off = tc::a_chunk_off<256>(r, k0 / 8 + c);
    *reinterpret_cast<uint4*>(a_hi + off) = hi;
    *reinterpret_cast<uint4*>(a_lo + off) = lo;
  }
}

__global__ void __launch_bounds__(TC_NB_THREADS, 1) tc_nodeblock_fwd_kernel(const __grid_constant__ TcNbArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* e_hi = smem_raw;                               // 128 x 64 bf16 = 16 KB
  uint8_t* e_lo = e_hi + tc::ROWS * C * 2;
  uint8_t* x_hi = e_lo + tc::ROWS * C * 2;                // 128 x 256 bf16 = 64 KB
  uint8_t* x_lo = x_hi + tc::ROWS * D * 2;
  uint8_t* stages = x_lo + tc::ROWS * D * 2;              // 2 x 32 KB
  tc::PipeSmem* ps = reinterpret_cast<tc::PipeSmem*>(stages + tc::NSTAGE * tc::STAGE_SLOT);
  float2* stat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(ps) + 128);     // [2][128]
  int* ls = reinterpret_cast<int*>(stat + 2 * tc::ROWS);
  float* out_tile = reinterpret_cast<float*>(smem_raw);   // [128][OUT_LD] fp32, aliases the E and X planes at the end

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * tc::ROWS;
  const Tables& tb = a.tb;
  TC_STAMP(0);
  tc::Pipe p;
  tc::pipe_init<TC_NRW>(p, ps, stages);
  if (warp == TC_NRW) tc::tmem_alloc<512>(&ps->tmem_base);
  const int row = (warp & 3) * 32 + lane;    // meaningful for row threads
  const int half = (warp >> 2) & 1;
  const int hc = half * 128;                 // first column of this thread's half
  int my_r = -1;
  if (p.role == 0) {
    const int q = q0 + row;
    if (q < a.n_edges) { my_r = a.right[q]; if (half == 0) ls[row] = a.left[q]; }
    else if (half == 0) ls[row] = -1;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t lane_base = ps->tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t D0 = 0, D1 = 256;          // TMEM column bases: msg accumulator / everything else
  const int rr = my_r < 0 ? 0 : my_r;
  TC_STAMP(1);

  // ---- rows: e tile -> E planes (each thread 32 of the 64 columns of its row)
  if (p.role == 0) {
    const int q = q0 + row;
    const int k0 = half * 32;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < a.n_edges) x = *reinterpret_cast<const float4*>(a.ebuf + (size_t)q * C + k0 + i);
      v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
    }
    tc::store_a32<C>(e_hi, e_lo, row, k0, v);
    tc::rows_publish(p);
    TC_STAMP(6);
  }
  // G1: edge_net.net.0                                                            graph.py:42
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_EN1), D1, false, true, true);
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    TC_STAMP(2);
    float v[128];
    load_half_row(lane_base + D1 + hc, v);
    add_vec128(v, a.v.en1_b + hc);
    ln_relu_half(v, a.v.en1_g + hc, a.v.en1_be + hc, stat, row, half);
    store_half_row_a(x_hi, x_lo, row, hc, v);
    tc::rows_publish(p);
    TC_STAMP(7);
  }
  // G2: edge_net.net.3 ; m = he * node_net(x)[col]                               graph.py:43
  tc::gemm<D, D>(p, x_hi, x_lo, TCW_(NB_EN2), D1, false, true, true);
  if (p.role == 0) {
    const float* hn = tb.hn + (size_t)rr * D + hc;          // gathered row: pull it into L1 while the MMAs run
#pragma unroll
    for (int i = 0; i < 128; i += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(hn + i));
    tc::rows_wait_acc(p);
    TC_STAMP(3);
    float v[128];
    load_half_row(lane_base + D1 + hc, v);
    add_vec128(v, a.v.en2_b + hc);
#pragma unroll
    for (int i = 0; i < 128; i += 4) {
      const float4 t4 = *reinterpret_cast<const float4*>(hn + i);
      v[i] *= t4.x; v[i + 1] *= t4.y; v[i + 2] *= t4.z; v[i + 3] *= t4.w;
    }
    store_half_row_a(x_hi, x_lo, row, hc, v);
    tc::rows_publish(p);
    TC_STAMP(8);
  }
  // G3: msg_net -> D0 (stays in TMEM) ; G4: gate.net.0 edge columns -> D1         graph.py:43,46
  tc::gemm<D, D>(p, x_hi, x_lo, TCW_(NB_MSG), D0, false, true, false);
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_GE), D1, false, false, true);
  if (p.role == 0) {
    const float* gxr = tb.gx + (size_t)rr * D + hc;
#pragma unroll
    for (int i = 0; i < 128; i += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(gxr + i));
    tc::rows_wait_acc(p);
    TC_STAMP(4);
    float v[128];
    load_half_row(lane_base + D1 + hc, v);
    add_row128(v, gxr);                                    // hoisted node / time / bias part of gate.net.0
    ln_relu_half(v, a.v.g1_g + hc, a.v.g1_be + hc, stat, row, half);
    store_half_row_a(x_hi, x_lo, row, hc, v);
    tc::rows_publish(p);
    TC_STAMP(9);
  }
  // G5: gate.net.3                                                                graph.py:46
  tc::gemm<D, D>(p, x_hi, x_lo, TCW_(NB_G2), D1, false, true, true);
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    TC_STAMP(5);
    // out = (msg + b) * sigmoid(gate + b)  -> smem tile (all operand planes are dead now)     graph.py:47
    {
      float v[128];
      load_half_row(lane_base + D1 + hc, v);
      add_vec128(v, a.v.g2_b + hc);
#pragma unroll
      for (int i = 0; i < 128; i += 4)                    // park sigmoid(gate) in the tile, then fold msg in
        *reinterpret_cast<float4*>(out_tile + row * OUT_LD + hc + i) =
            make_float4(tc::fast_sigmoid(v[i]), tc::fast_sigmoid(v[i + 1]), tc::fast_sigmoid(v[i + 2]),
                        tc::fast_sigmoid(v[i + 3]));
      load_half_row(lane_base + D0 + hc, v);
      add_vec128(v, a.v.msg_b + hc);
#pragma unroll
      for (int i = 0; i < 128; i += 4) {
        float4* o = reinterpret_cast<float4*>(out_tile + row * OUT_LD + hc + i);
        const float4 g = *o;
        *o = make_float4(v[i] * g.x, v[i + 1] * g.y, v[i + 2] * g.z, v[i + 3] * g.w);
      }
    }
    tc::fence_before_sync();
    asm volatile("bar.sync 1, 256;" ::: "memory");      // row threads only
    // scatter_sum over row (= left): thread c owns channel c and walks the 128 CSR-ordered rows       graph.py:50
    int cur = ls[0];
    float s0 = 0.f;
#pragma unroll 1
    for (int r8 = 0; r8 < tc::ROWS; r8 += 8) {
      float x[8];
      int n[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { x[u] = out_tile[(r8 + u) * OUT_LD + tid]; n[u] = ls[r8 + u]; }   // 8 loads in flight
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (n[u] != cur) {
          if (cur >= 0) atomicAdd(tb.agg + (size_t)cur * D + tid, s0);
          cur = n[u]; s0 = 0.f;
        }
        s0 += x[u];
      }
    }
    if (cur >= 0) atomicAdd(tb.agg + (size_t)cur * D + tid, s0);
  }
  TC_STAMP(15);
  __syncthreads();
  if (warp == TC_NRW) { __syncwarp(); tc::tmem_dealloc<512>(ps->tmem_base); }
}

// host: copy the eight 256-float parameter vectors of block `off` out of the host blob
inline void fill_nb_vecs(NbVecs& v, const float* blob_host, const BlkOff& off) {
  auto cp = [&](float* dst, int slot) { memcpy(dst, blob_host + off.o[slot], D * sizeof(float)); };
  cp(v.en1_b, MDB_S_NB_EN1_B); cp(v.en1_g, MDB_S_NB_EN1_G); cp(v.en1_be, MDB_S_NB_EN1_BE); cp(v.en2_b, MDB_S_NB_EN2_B);
  cp(v.msg_b, MDB_S_NB_MSG_B); cp(v.g1_g, MDB_S_NB_G1_G); cp(v.g1_be, MDB_S_NB_G1_BE); cp(v.g2_b, MDB_S_NB_G2_B);
}

constexpr size_t SMEM_TC_NB = 2 * (size_t)tc::ROWS * C * 2 + 2 * (size_t)tc::ROWS * D * 2 + tc::NSTAGE * tc::STAGE_SLOT
                              + 128 + 2 * tc::ROWS * sizeof(float2) + tc::ROWS * sizeof(int) + 64;
static_assert(sizeof(tc::PipeSmem) <= 128, "PipeSmem must fit its 128-byte slot");
static_assert(SMEM_TC_NB <= 232448, "tc_nodeblock_fwd_kernel exceeds the 227 KB shared-memory limit");
static_assert((size_t)tc::ROWS * OUT_LD * 4 <= 2 * (size_t)tc::ROWS * C * 2 + 2 * (size_t)tc::ROWS * D * 2,
              "out tile must fit in the operand planes it aliases");

// =================================================================================================================
// Backward of the NodeBlock per-edge path on tensor cores (same math as bwd_edge_nodeblock_kernel in mdb_backward.cuh)
// =================================================================================================================
struct TcNbBwdArgs {
  const float* blob;
  const uint8_t* tc_blob;
  BlkOff off;
  TcOff tco;
  Tables tb;
  const int *left, *right;
  int n_nodes, n_edges;
  const float* e;        // [E][64] saved e_i
  const float* dagg;     // [N][256] d/d (aggregated messages)
  float *dgx, *dhn;      // [N][256] scatter targets (pre-zeroed)
  float* de;             // [E][64]  d/d e, accumulated (+=)
  NbVecs v;
  long long* dbg;        // optional [grid][32] clock64 stamps of row thread 0 (tools/tc_phase_times.py --bwd)
};

// cross-half exchange of two partial values per row; returns the partner's pair
__device__ __forceinline__ float2 exchange_half(float2* stat, int row, int half, float x, float y) {
  stat[half * tc::ROWS + row] = make_float2(x, y);
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const float2 o = stat[(half ^ 1) * tc::ROWS + row];
  asm volatile("bar.sync 1, 256;" ::: "memory");     // the buffer may be rewritten right after
  return o;
}

// v[i] = f(i, acc[i], v[i]) over this thread's 128 accumulator columns, read in four 32-column chunks with the next chunk's
// TMEM load in flight; v[] holds an operand gathered BEFORE the accumulator wait (peak registers 128 + 2 x 32).
template <typename F>
__device__ __forceinline__ void combine_cols128(uint32_t taddr, float (&v)[128], F&& f) {
  uint32_t r[2][32];
  tc::tmem_ld32_issue(taddr, r[0]);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c + 1 < 4) tc::tmem_ld32_issue(taddr + (c + 1) * 32, r[(c + 1) & 1]);
    tc::tmem_ld32_wait(r[c & 1]);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[c * 32 + i] = f(c * 32 + i, tc::acc_f(r[c & 1][i]), v[c * 32 + i]);
  }
}
// this thread's gathered table row part (128 floats) -> registers
__device__ __forceinline__ void gather_row128(const float* __restrict__ src, float (&v)[128]) {
#pragma unroll
  for (int i = 0; i < 128; i += 4) {
    const float4 t4 = *reinterpret_cast<const float4*>(src + i);
    v[i] = t4.x; v[i + 1] = t4.y; v[i + 2] = t4.z; v[i + 3] = t4.w;
  }
}

// LayerNorm backward with the row statistics (mean, rstd) SAVED from the forward recompute: two passes over the
// pre-LayerNorm activations a = acc (TMEM) [+ bias] [+ extra row, folded into the accumulator on the first pass], next
// chunk's TMEM load in flight.  d: in = gradient w.r.t. relu(LN(a) g + b), out = gradient w.r.t. a.
template <bool HAS_BIAS, bool FOLD_EXTRA>
__device__ __forceinline__ void ln_bwd_half_ms(uint32_t taddr, const float* __restrict__ bias, const float* __restrict__ extra,
                                               const float* __restrict__ gamma, const float* __restrict__ beta, float2 ms,
                                               float (&d)[128], float2* stat, int row, int half) {
  const float mean = ms.x, rstd = ms.y;
  float s1 = 0.f, s2 = 0.f;
  {
    uint32_t r[2][32];
    tc::tmem_ld32_issue(taddr, r[0]);
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      if (cc + 1 < 4) tc::tmem_ld32_issue(taddr + (cc + 1) * 32, r[(cc + 1) & 1]);
      tc::tmem_ld32_wait(r[cc & 1]);
      float a[32];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (FOLD_EXTRA) x = *reinterpret_cast<const float4*>(extra + cc * 32 + i);
        const float xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float av = tc::acc_f(r[cc & 1][i + u]);
          if (HAS_BIAS) av += bias[cc * 32 + i + u];
          if (FOLD_EXTRA) av += xx[u];
          a[i + u] = av;
          const float gg = gamma[cc * 32 + i + u], bb = beta[cc * 32 + i + u];
          const float xh = (av - mean) * rstd;
          const float dxh = (xh * gg + bb > 0.f) ? d[cc * 32 + i + u] * gg : 0.f;
          d[cc * 32 + i + u] = dxh;
          s1 += dxh;
          s2 = fmaf(dxh, xh, s2);
        }
      }
      if (FOLD_EXTRA) tc::tmem_st32(taddr + cc * 32, a);     // the second pass re-reads acc + extra
    }
    if (FOLD_EXTRA) tc::tmem_st_wait();
  }
  const float2 o2 = exchange_half(stat, row, half, s1, s2);
  const float m1 = (s1 + o2.x) * (1.f / 256.f), m2 = (s2 + o2.y) * (1.f / 256.f);
  {
    uint32_t r[2][32];
    tc::tmem_ld32_issue(taddr, r[0]);
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      if (cc + 1 < 4) tc::tmem_ld32_issue(taddr + (cc + 1) * 32, r[(cc + 1) & 1]);
      tc::tmem_ld32_wait(r[cc & 1]);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float av = tc::acc_f(r[cc & 1][i]);
        if (HAS_BIAS) av += bias[cc * 32 + i];
        const float xh = (av - mean) * rstd;
        d[cc * 32 + i] = rstd * (d[cc * 32 + i] - m1 - xh * m2);
      }
    }
  }
}

template <bool IS_ROW>
__device__ __forceinline__ void tc_nodeblock_bwd_body(const TcNbBwdArgs& a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* e_hi = smem_raw;
  uint8_t* e_lo = e_hi + tc::ROWS * C * 2;
  uint8_t* x_hi = e_lo + tc::ROWS * C * 2;
  uint8_t* x_lo = x_hi + tc::ROWS * D * 2;
  uint8_t* stages = x_lo + tc::ROWS * D * 2;
  tc::PipeSmem* ps = reinterpret_cast<tc::PipeSmem*>(stages + tc::NSTAGE * tc::STAGE_SLOT);
  float2* stat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(ps) + 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TC_STAMP(0);
  const int q0 = blockIdx.x * tc::ROWS;
  const Tables& tb = a.tb;
  tc::Pipe p;
  tc::pipe_init_split<TC_NRW, IS_ROW>(p, ps, stages);
  if (warp == TC_NRW) tc::tmem_alloc<512>(&ps->tmem_base);
  const int row = (warp & 3) * 32 + lane;
  const int half = (warp >> 2) & 1;
  const int hc = half * 128;
  const int dc = half * 32;                   // this thread's 32 of the 64 d/d e columns
  const int q = q0 + row;
  const bool valid = IS_ROW && q < a.n_edges;
  float e32[32];                              // e tile part: requested before the set-up barrier
  if (IS_ROW) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) x = *reinterpret_cast<const float4*>(a.e + (size_t)q * C + dc + i);
      e32[i] = x.x; e32[i + 1] = x.y; e32[i + 2] = x.z; e32[i + 3] = x.w;
    }
  }
  const int ll = valid ? a.left[q] : 0, rr = valid ? a.right[q] : 0;
  tc::fence_before_sync();
  tc::cta_sync();
  tc::fence_after_sync();
  TC_STAMP(1);
  const uint32_t lane_base = ps->tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t D0 = 0, D1 = 256;
  const float* hn = tb.hn + (size_t)rr * D + hc;
  const float* gxr = tb.gx + (size_t)rr * D + hc;
  float de[32];
  float2 ms_en1 = make_float2(0.f, 1.f), ms_g1 = make_float2(0.f, 1.f);   // LayerNorm (mean, rstd) of the two LN layers

  if (IS_ROW) {   // e tile -> E planes
    tc::store_a32<C>(e_hi, e_lo, row, dc, e32);
    tc::rows_publish(p);
    TC_STAMP(2);
  }
  // ---- forward recompute up to msg (D0) and sigmoid(gate)
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_EN1), D1, false, true, true);
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    TC_STAMP(3);
    float v[128];
    load_half_row(lane_base + D1 + hc, v);
    add_vec128(v, a.v.en1_b + hc);
    ms_en1 = ln_stats_half(v, stat, row, half);
    const float2 ms = ms_en1;
    const float* gam = a.v.en1_g + hc;
    const float* bet = a.v.en1_be + hc;
    tc::store_a_sliced<128>(p, x_hi, x_lo, row, hc,
                            [&](int i) { return fmaxf((v[i] - ms.x) * ms.y * gam[i] + bet[i], 0.f); });
    TC_STAMP(4);
  }
  tc::gemm<D, D, tc::NSTAGE, 2>(p, x_hi, x_lo, TCW_(NB_EN2), D1, false, true, true);
  if (IS_ROW) {
    float v[128];
    gather_row128(hn, v);                       // in flight while the GEMM runs
    tc::rows_wait_acc(p);
    TC_STAMP(5);
    const float* b2 = a.v.en2_b + hc;
    combine_cols128(lane_base + D1 + hc, v, [&](int i, float acc, float h) { return (acc + b2[i]) * h; });
    tc::store_a_sliced<128>(p, x_hi, x_lo, row, hc, [&](int i) { return v[i]; });
    TC_STAMP(6);
  }
  tc::gemm<D, D, tc::NSTAGE, 2>(p, x_hi, x_lo, TCW_(NB_MSG), D0, false, true, false);
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_GE), D1, false, false, true);
  if (IS_ROW) {
    float v[128];
    gather_row128(gxr, v);
    tc::rows_wait_acc(p);
    TC_STAMP(7);
    combine_cols128(lane_base + D1 + hc, v, [&](int, float acc, float g) { return acc + g; });
    ms_g1 = ln_stats_half(v, stat, row, half);
    const float2 ms = ms_g1;
    const float* gam = a.v.g1_g + hc;
    const float* bet = a.v.g1_be + hc;
    tc::store_a_sliced<128>(p, x_hi, x_lo, row, hc,
                            [&](int i) { return fmaxf((v[i] - ms.x) * ms.y * gam[i] + bet[i], 0.f); });
    TC_STAMP(8);
  }
  tc::gemm<D, D, tc::NSTAGE, 2>(p, x_hi, x_lo, TCW_(NB_G2), D1, false, true, true);
  // ---- d out = dagg[l]:  d gate-logit -> X planes (A operand of the next GEMM),  d msg -> parked in D0 over msg
  if (IS_ROW) {
    float dd[128];
    if (valid) gather_row128(a.dagg + (size_t)ll * D + hc, dd);
    else {
#pragma unroll
      for (int i = 0; i < 128; ++i) dd[i] = 0.f;
    }
    tc::rows_wait_acc(p);
    TC_STAMP(9);
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t rs[32], rm[32];
      tc::tmem_ld32_issue(lane_base + D1 + hc + cc * 32, rs);
      tc::tmem_ld32_issue(lane_base + D0 + hc + cc * 32, rm);
      tc::tmem_ld32_wait(rs); tc::tmem_ld32_wait(rm);
      float m[32], gl[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float sgm = tc::fast_sigmoid(tc::acc_f(rs[i]) + a.v.g2_b[hc + cc * 32 + i]);
        const float msg = tc::acc_f(rm[i]) + a.v.msg_b[hc + cc * 32 + i];
        const float dz = dd[cc * 32 + i];
        m[i] = dz * sgm;                                  // d msg
        gl[i] = dz * msg * sgm * (1.f - sgm);             // d gate logit
      }
      tc::tmem_st32(lane_base + D0 + hc + cc * 32, m);
      tc::store_a32<D>(x_hi, x_lo, row, hc + cc * 32, gl);   // G2 has completed: the X planes are free
    }
    tc::tmem_st_wait();
    tc::rows_publish(p);
    TC_STAMP(10);
  }
  // ---- gate branch: d relu3 = d gt W_g2^T ; LN backward needs a3 again -> recompute into D1 after draining it
  tc::gemm<D, D>(p, x_hi, x_lo, TCW_(BT_NB_G2), D1, false, true, true);
  float dr[128];
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    TC_STAMP(11);
    load_half_row(lane_base + D1 + hc, dr);
    tc::rows_publish(p);
    TC_STAMP(12);
  }
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_GE), D1, false, true, true);
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    TC_STAMP(13);
    // a3 = acc + gx[r] (folded into the accumulator on the first pass); statistics saved from the recompute above
    ln_bwd_half_ms<false, true>(lane_base + D1 + hc, nullptr, gxr, a.v.g1_g + hc, a.v.g1_be + hc, ms_g1, dr, stat, row, half);
    if (valid) {
      float* dst = a.dgx + (size_t)rr * D + hc;
#pragma unroll
      for (int i = 0; i < 128; i += 4) tc::red_add_v4(dst + i, dr[i], dr[i + 1], dr[i + 2], dr[i + 3]);
    }
    tc::store_a_sliced<128>(p, x_hi, x_lo, row, hc, [&](int i) { return dr[i]; });
    TC_STAMP(14);
  }
  tc::gemm<D, C, tc::NSTAGE, 2>(p, x_hi, x_lo, TCW_(BT_NB_GE), D1, false, true, true);      // d e (gate part) -> D1[0:64]
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    TC_STAMP(15);
    tc::tmem_ld32(lane_base + D1 + dc, de);
    // d msg: D0 -> X planes
    load_half_row(lane_base + D0 + hc, dr);
    tc::store_a_sliced<128>(p, x_hi, x_lo, row, hc, [&](int i) { return dr[i]; });
    TC_STAMP(16);
  }
  // ---- message branch: dm = d msg W_msg^T -> D0 ; he needed again -> recompute (a2 -> r2 -> he) through D1
  tc::gemm<D, D, tc::NSTAGE, 2>(p, x_hi, x_lo, TCW_(BT_NB_MSG), D0, false, true, false);
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_EN1), D1, false, false, true);
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    TC_STAMP(17);
    float v[128];
    load_half_row(lane_base + D1 + hc, v);
    add_vec128(v, a.v.en1_b + hc);
    const float2 ms = ms_en1;                    // same row, same statistics: no second reduction
    const float* gam = a.v.en1_g + hc;
    const float* bet = a.v.en1_be + hc;
    tc::store_a_sliced<128>(p, x_hi, x_lo, row, hc,
                            [&](int i) { return fmaxf((v[i] - ms.x) * ms.y * gam[i] + bet[i], 0.f); });
    TC_STAMP(18);
  }
  tc::gemm<D, D, tc::NSTAGE, 2>(p, x_hi, x_lo, TCW_(NB_EN2), D1, false, true, true);         // he
  if (IS_ROW) {
    float hv[128];
    gather_row128(hn, hv);
    tc::rows_wait_acc(p);
    TC_STAMP(19);
    float* dst = a.dhn + (size_t)rr * D + hc;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t rh[32], rd[32];
      tc::tmem_ld32_issue(lane_base + D1 + hc + cc * 32, rh);
      tc::tmem_ld32_issue(lane_base + D0 + hc + cc * 32, rd);
      tc::tmem_ld32_wait(rh); tc::tmem_ld32_wait(rd);
      float dm[32];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float dmv = tc::acc_f(rd[i + u]);
          o[u] = dmv * (tc::acc_f(rh[i + u]) + a.v.en2_b[hc + cc * 32 + i + u]);     // d node_net(x)[col]
          dm[i + u] = dmv * hv[cc * 32 + i + u];                                      // d he
        }
        if (valid) tc::red_add_v4(dst + cc * 32 + i, o[0], o[1], o[2], o[3]);
      }
      tc::store_a32<D>(x_hi, x_lo, row, hc + cc * 32, dm);
    }
    tc::rows_publish(p);
    TC_STAMP(20);
  }
  // d relu2 = d he W_en2^T -> D1 ; a2 recompute -> D0 (dm is dead) ; one epilogue for both
  tc::gemm<D, D>(p, x_hi, x_lo, TCW_(BT_NB_EN2), D1, false, true, false);
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_EN1), D0, false, false, true);
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    TC_STAMP(21);
    load_half_row(lane_base + D1 + hc, dr);
    ln_bwd_half_ms<true, false>(lane_base + D0 + hc, a.v.en1_b + hc, nullptr, a.v.en1_g + hc, a.v.en1_be + hc, ms_en1,
                                dr, stat, row, half);
    tc::store_a_sliced<128>(p, x_hi, x_lo, row, hc, [&](int i) { return dr[i]; });
    TC_STAMP(22);
  }
  tc::gemm<D, C, tc::NSTAGE, 2>(p, x_hi, x_lo, TCW_(BT_NB_EN1), D1, false, true, true);       // d e (message part) -> D1[0:64]
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    TC_STAMP(23);
    float v[32];
    tc::tmem_ld32(lane_base + D1 + dc, v);
    if (valid) {
      float* dst = a.de + (size_t)q * C + dc;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 o = *reinterpret_cast<float4*>(dst + i);
        o.x += de[i] + v[i]; o.y += de[i + 1] + v[i + 1]; o.z += de[i + 2] + v[i + 2]; o.w += de[i + 3] + v[i + 3];
        *reinterpret_cast<float4*>(dst + i) = o;
      }
    }
    tc::fence_before_sync();
  }
  TC_STAMP(24);
  tc::cta_sync();
  if (warp == TC_NRW) { __syncwarp(); tc::tmem_dealloc<512>(ps->tmem_base); }
}

__global__ void __launch_bounds__(tc::RB_THREADS, 1) tc_nodeblock_bwd_kernel(const __grid_constant__ TcNbBwdArgs a) {
  if (threadIdx.x < TC_NRW * 32) {
    tc::reg_alloc<232>();
    tc_nodeblock_bwd_body<true>(a);
  } else {
    tc::reg_dealloc<40>();
    tc_nodeblock_bwd_body<false>(a);
  }
}

constexpr size_t SMEM_TC_NB_BWD = 2 * (size_t)tc::ROWS * C * 2 + 2 * (size_t)tc::ROWS * D * 2
                                  + tc::NSTAGE * tc::STAGE_SLOT + 128 + 2 * tc::ROWS * sizeof(float2) + 64;
static_assert(SMEM_TC_NB_BWD <= 232448, "tc_nodeblock_bwd_kernel exceeds the 227 KB shared-memory limit");
