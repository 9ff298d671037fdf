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

struct TcNbArgs {
  const float* blob;          // fp32 blob: biases / LayerNorm parameters
  const uint8_t* tc_blob;     // split-bf16 weight stage images
  BlkOff off;
  TcOff tco;
  Tables tb;
  const int *left, *right;
  int n_nodes, n_edges;
  const float* ebuf;          // [E][64] e = edge_embs(cat(h_edge, rbf)) (written by edge_kernel_b)
};

constexpr int OUT_LD = 260;   // fp32 row stride of the out tile in smem: 1040 B = 16 (mod 128) -> conflict-free rows

// Row-thread helpers ------------------------------------------------------------------------------------------
// mean / rstd of (acc[lane row][0:256] + bias[0:256] (+ extra row)) with two TMEM passes (two-pass variance).
template <bool HAS_EXTRA>
__device__ __forceinline__ void row_ln_stats(uint32_t taddr, const float* __restrict__ bias,
                                             const float* __restrict__ extra, float& mean, float& rstd) {
  float s = 0.f;
  for (int c0 = 0; c0 < 256; c0 += 32) {
    float v[32];
    tc::tmem_ld32(taddr + c0, v);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias + c0 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (HAS_EXTRA) x = *reinterpret_cast<const float4*>(extra + c0 + i);
      s += (v[i] + b.x + x.x) + (v[i + 1] + b.y + x.y) + (v[i + 2] + b.z + x.z) + (v[i + 3] + b.w + x.w);
    }
  }
  mean = s * (1.f / 256.f);
  float q = 0.f;
  for (int c0 = 0; c0 < 256; c0 += 32) {
    float v[32];
    tc::tmem_ld32(taddr + c0, v);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias + c0 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (HAS_EXTRA) x = *reinterpret_cast<const float4*>(extra + c0 + i);
      const float d0 = v[i] + b.x + x.x - mean, d1 = v[i + 1] + b.y + x.y - mean;
      const float d2 = v[i + 2] + b.z + x.z - mean, d3 = v[i + 3] + b.w + x.w - mean;
      q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
    }
  }
  rstd = 1.f / sqrtf(q * (1.f / 256.f) + LN_EPS);
}

// A planes (K = 256) <- relu(LN(acc + bias (+ extra)) * gamma + beta), row `r`.
template <bool HAS_EXTRA>
__device__ __forceinline__ void row_ln_relu_to_a(uint32_t taddr, const float* __restrict__ bias,
                                                 const float* __restrict__ extra, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, uint8_t* a_hi, uint8_t* a_lo, int r) {
  float mean, rstd;
  row_ln_stats<HAS_EXTRA>(taddr, bias, extra, mean, rstd);
  for (int c0 = 0; c0 < 256; c0 += 32) {
    float v[32];
    tc::tmem_ld32(taddr + c0, v);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias + c0 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c0 + i));
      const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c0 + i));
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (HAS_EXTRA) x = *reinterpret_cast<const float4*>(extra + c0 + i);
      v[i] = fmaxf((v[i] + b.x + x.x - mean) * rstd * g.x + be.x, 0.f);
      v[i + 1] = fmaxf((v[i + 1] + b.y + x.y - mean) * rstd * g.y + be.y, 0.f);
      v[i + 2] = fmaxf((v[i + 2] + b.z + x.z - mean) * rstd * g.z + be.z, 0.f);
      v[i + 3] = fmaxf((v[i + 3] + b.w + x.w - mean) * rstd * g.w + be.w, 0.f);
    }
    tc::store_a32<256>(a_hi, a_lo, r, c0, v);
  }
}

__global__ void __launch_bounds__(tc::NTHREADS_TC, 1) tc_nodeblock_fwd_kernel(const TcNbArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* e_hi = smem_raw;                               // 128 x 64 bf16 = 16 KB
  uint8_t* e_lo = e_hi + tc::ROWS * C * 2;
  uint8_t* x_hi = e_lo + tc::ROWS * C * 2;                // 128 x 256 bf16 = 64 KB
  uint8_t* x_lo = x_hi + tc::ROWS * D * 2;
  uint8_t* stages = x_lo + tc::ROWS * D * 2;              // 2 x 32 KB
  tc::PipeSmem* ps = reinterpret_cast<tc::PipeSmem*>(stages + tc::NSTAGE * tc::STAGE_SLOT);
  int* ls = reinterpret_cast<int*>(ps + 1);
  int* rs = ls + tc::ROWS;
  float* out_tile = reinterpret_cast<float*>(smem_raw);   // [128][OUT_LD] fp32, aliases the E and X planes at the end

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * tc::ROWS;
  const float* blob = a.blob;
  const BlkOff& off = a.off;
  const Tables& tb = a.tb;
  tc::Pipe p;
  tc::pipe_init(p, ps, stages);
  if (warp == 4) tc::tmem_alloc<512>(&ps->tmem_base);
  int my_l = -1, my_r = -1;
  if (p.role == 0) {
    const int q = q0 + tid;
    if (q < a.n_edges) { my_l = a.left[q]; my_r = a.right[q]; }
    ls[tid] = my_l; rs[tid] = my_r;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t lane_base = ps->tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t D0 = 0, D1 = 256;          // TMEM column bases: msg accumulator / everything else
  const int rr = my_r < 0 ? 0 : my_r;

  // ---- rows: e tile -> E planes
  if (p.role == 0) {
    const int q = q0 + tid;
    for (int k0 = 0; k0 < C; k0 += 32) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < a.n_edges) x = *reinterpret_cast<const float4*>(a.ebuf + (size_t)q * C + k0 + i);
        v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
      }
      tc::store_a32<C>(e_hi, e_lo, tid, k0, v);
    }
    tc::rows_publish(p);
  }
  // G1: edge_net.net.0                                                            graph.py:42
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_EN1), D1, false, true, true);
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    row_ln_relu_to_a<false>(lane_base + D1, W_(NB_EN1_B), nullptr, W_(NB_EN1_G), W_(NB_EN1_BE), x_hi, x_lo, tid);
    tc::rows_publish(p);
  }
  // G2: edge_net.net.3 ; m = he * node_net(x)[col]                               graph.py:43
  tc::gemm<D, D>(p, x_hi, x_lo, TCW_(NB_EN2), D1, false, true, true);
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    const float* hn = tb.hn + (size_t)rr * D;
    const float* b2 = W_(NB_EN2_B);
    for (int c0 = 0; c0 < D; c0 += 32) {
      float v[32];
      tc::tmem_ld32(lane_base + D1 + c0, v);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(b2 + c0 + i));
        const float4 h = *reinterpret_cast<const float4*>(hn + c0 + i);
        v[i] = (v[i] + b.x) * h.x; v[i + 1] = (v[i + 1] + b.y) * h.y;
        v[i + 2] = (v[i + 2] + b.z) * h.z; v[i + 3] = (v[i + 3] + b.w) * h.w;
      }
      tc::store_a32<D>(x_hi, x_lo, tid, c0, v);
    }
    tc::rows_publish(p);
  }
  // G3: msg_net -> D0 (stays in TMEM) ; G4: gate.net.0 edge columns -> D1         graph.py:43,46
  tc::gemm<D, D>(p, x_hi, x_lo, TCW_(NB_MSG), D0, false, true, false);
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_GE), D1, false, false, true);
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    row_ln_relu_to_a<true>(lane_base + D1, nullptr, tb.gx + (size_t)rr * D, W_(NB_G1_G), W_(NB_G1_BE), x_hi, x_lo, tid);
    tc::rows_publish(p);
  }
  // G5: gate.net.3                                                                graph.py:46
  tc::gemm<D, D>(p, x_hi, x_lo, TCW_(NB_G2), D1, false, true, true);
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    // out = (msg + b) * sigmoid(gate + b)  -> smem tile (all operand planes are dead now)     graph.py:47
    const float* bm = W_(NB_MSG_B);
    const float* bg = W_(NB_G2_B);
    for (int c0 = 0; c0 < D; c0 += 32) {
      float m[32], g[32];
      tc::tmem_ld32(lane_base + D0 + c0, m);
      tc::tmem_ld32(lane_base + D1 + c0, g);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bm + c0 + i));
        const float4 b2 = __ldg(reinterpret_cast<const float4*>(bg + c0 + i));
        float4 o;
        o.x = (m[i] + b1.x) * (1.f / (1.f + expf(-(g[i] + b2.x))));
        o.y = (m[i + 1] + b1.y) * (1.f / (1.f + expf(-(g[i + 1] + b2.y))));
        o.z = (m[i + 2] + b1.z) * (1.f / (1.f + expf(-(g[i + 2] + b2.z))));
        o.w = (m[i + 3] + b1.w) * (1.f / (1.f + expf(-(g[i + 3] + b2.w))));
        *reinterpret_cast<float4*>(out_tile + tid * OUT_LD + c0 + i) = o;
      }
    }
    tc::fence_before_sync();
    asm volatile("bar.sync 1, 128;" ::: "memory");      // row threads only
    // scatter_sum over row (= left): thread c owns channels c and c + 128, walks the 128 CSR-ordered rows   graph.py:50
    int cur = ls[0];
    float s0 = 0.f, s1 = 0.f;
    for (int r = 0; r < tc::ROWS; ++r) {
      const int n = ls[r];
      if (n != cur) {
        if (cur >= 0) { atomicAdd(tb.agg + (size_t)cur * D + tid, s0); atomicAdd(tb.agg + (size_t)cur * D + tid + 128, s1); }
        cur = n; s0 = 0.f; s1 = 0.f;
      }
      s0 += out_tile[r * OUT_LD + tid];
      s1 += out_tile[r * OUT_LD + tid + 128];
    }
    if (cur >= 0) { atomicAdd(tb.agg + (size_t)cur * D + tid, s0); atomicAdd(tb.agg + (size_t)cur * D + tid + 128, s1); }
  }
  __syncthreads();
  if (warp == 4) { __syncwarp(); tc::tmem_dealloc<512>(ps->tmem_base); }
}

constexpr size_t SMEM_TC_NB = 2 * (size_t)tc::ROWS * C * 2 + 2 * (size_t)tc::ROWS * D * 2 + tc::NSTAGE * tc::STAGE_SLOT
                              + sizeof(tc::PipeSmem) + 2 * tc::ROWS * sizeof(int) + 64;
static_assert(SMEM_TC_NB <= 232448, "tc_nodeblock_fwd_kernel exceeds the 227 KB shared-memory limit");
static_assert((size_t)tc::ROWS * OUT_LD * 4 <= 2 * (size_t)tc::ROWS * C * 2 + 2 * (size_t)tc::ROWS * D * 2,
              "out tile must fit in the operand planes it aliases");
