// Forward kernels of the NodeEdgeNet hot path (reference: models/graph.py:10-55,122-141,251-396) and the
// C-ABI entry points declared in include/moldiff_b200.h.
//
// Kernel schedule for one network forward (L blocks):
//   node_init, edge_init                       embedders + time features         model.py:210-213
//   node_kernel(pre 0)                         per-node hoisted projections for block 0
//   for i in 0..L-1:
//     edge_kernel_b(i)   RBF, edge_embs[i], NodeBlock edge path -> AGG, BondFFN L/R -> SL/SR
//     node_kernel(mid i, pre i+1 | decode)     NodeBlock node path, h_node residual, PosUpdate node MLPs,
//                                              hoisted projections for block i+1
//     edge_kernel_d(i)   EdgeBlock tail -> h_edge, PosUpdate edge path -> pos
//   edge_decode                                 edge decoder on h_edge[p] + h_edge[p + E/2]
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/moldiff_b200.h"
#include "tile_engine.cuh"

using namespace mdb;

namespace {

constexpr int D = MDB_NODE_DIM;   // 256
constexpr int C = MDB_EDGE_DIM;   // 64
constexpr int G = MDB_NUM_RBF;    // 16

struct BlkOff { int o[MDB_NUM_BLOCK_SLOTS]; };
struct HeadOff { int o[MDB_NUM_HEAD_SLOTS]; };

// Per-node tables produced by node_kernel and gathered by the edge kernels (workspace carve-up).
struct Tables {
  float *x, *agg, *hn, *gx, *cen;       // [N][256]
  float *hnb, *gxb;                     // node-blocked copies of hn / gx (tile_engine.cuh: blk_off), read by the tc kernels
  float *nll, *nlr;                     // [N][128]   bond_ffn_{left,right}.node_linear(h_node)
  float *gnl, *gnr;                     // [N][32]    bond_ffn gate first layer, node + time + bias part
  float *fl, *fr;                       // [2 parities][N][64]  node_ffn_{left,right}(h_node)
  float *lf, *rf, *dect;                // [N][64]
  float *slsr;                          // [2 parities][2 (SL,SR)][N][64]
  float *pos0, *pos1;                   // [N][3]
  float *tn;                            // [N]
  float *hedge, *ebuf;                  // [E][64] sorted order
  float *te;                            // [E]
};

#define W_(slot) (blob + off.o[MDB_S_##slot])
#define H_(slot) (blob + hoff.o[MDB_H_##slot])

// ------------------------------------------------------------------------------------------------
// init kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float smear(float v, float lo, float hi, float offset, float coeff) {
  const float d = fminf(fmaxf(v, lo), hi) - offset;
  return expf(coeff * (d * d));   // coeff * pow(d, 2): same rounding order as common.py:237
}

// kind 1/2: x[n] = [node_embedder(h_node_pert[n]) ; time_emb(t[batch[n]])], tn[n] = t/T.  kind 0: copy.
__global__ void node_init_kernel(int kind, int n_nodes, int kn, int time_dim, float T,
                                 const float* __restrict__ blob, HeadOff hoff,
                                 const float* __restrict__ h_in, const int64_t* __restrict__ batch,
                                 const int64_t* __restrict__ t, const float* __restrict__ node_time,
                                 float* __restrict__ x, float* __restrict__ tn) {
  const int n = blockIdx.x;
  const int c = threadIdx.x;
  if (kind == 0) {
    x[(size_t)n * D + c] = h_in[(size_t)n * D + c];
    if (c == 0) tn[n] = node_time[n];
    return;
  }
  const float tt = (float)t[batch[n]];
  const int de = D - time_dim;
  float v;
  if (c < de) {
    const float* w = H_(NODE_EMB_W);
    v = 0.f;
    for (int k = 0; k < kn; ++k) v = fmaf(h_in[(size_t)n * kn + k], w[k * de + c], v);
  } else {
    const int j = c - de;
    v = smear(tt, 0.f, T, H_(TIME_OFFSET)[j], H_(TIME_COEFF)[j]);
  }
  x[(size_t)n * D + c] = v;
  if (c == 0) tn[n] = tt / T;
}

// Sorted edge q (orig id p = perm[q]).  kind 1: [edge_embedder(h_edge_pert[p]) ; time_emb]; kind 2:
// [edge_embedder(cat(h_node[l], h_node[r])) ; time_emb] (bond_predictor.py:135-140); kind 0: gather.
__global__ void edge_init_kernel(int kind, int n_edges, int kn, int ke, int time_dim, float T,
                                 const float* __restrict__ blob, HeadOff hoff,
                                 const float* __restrict__ h_edge_in, const float* __restrict__ h_node_in,
                                 const int64_t* __restrict__ batch_edge, const int64_t* __restrict__ t,
                                 const float* __restrict__ edge_time,
                                 const int* __restrict__ left, const int* __restrict__ right,
                                 const int* __restrict__ perm,
                                 float* __restrict__ hedge, float* __restrict__ te) {
  const int q = blockIdx.x * 4 + (threadIdx.x >> 6);
  const int c = threadIdx.x & 63;
  if (q >= n_edges) return;
  const int p = perm[q];
  if (kind == 0) {
    hedge[(size_t)q * C + c] = h_edge_in[(size_t)p * C + c];
    if (c == 0) te[q] = edge_time[p];
    return;
  }
  const float tt = (float)t[batch_edge[p]];
  const int de = C - time_dim;
  float v;
  if (c < de) {
    const float* w = H_(EDGE_EMB_W);
    v = 0.f;
    if (kind == 1) {
      for (int k = 0; k < ke; ++k) v = fmaf(h_edge_in[(size_t)p * ke + k], w[k * de + c], v);
    } else {
      const int l = left[q], r = right[q];
      for (int k = 0; k < kn; ++k) v = fmaf(h_node_in[(size_t)l * kn + k], w[k * de + c], v);
      for (int k = 0; k < kn; ++k) v = fmaf(h_node_in[(size_t)r * kn + k], w[(kn + k) * de + c], v);
    }
  } else {
    const int j = c - de;
    v = smear(tt, 0.f, T, H_(TIME_OFFSET)[j], H_(TIME_COEFF)[j]);
  }
  hedge[(size_t)q * C + c] = v;
  if (c == 0) te[q] = tt / T;
}

// ------------------------------------------------------------------------------------------------
// node kernel: one 64-node tile per CTA
// ------------------------------------------------------------------------------------------------
struct NodeArgs {
  const float* blob;
  BlkOff mid, pre;      // offsets of block i (mid) and block i+1 (pre)
  HeadOff head;
  Tables tb;            // tables of block i (mid phase reads cen / agg)
  Tables tb_pre;        // tables the pre phase writes for block i+1 (same buffers unless the forward saves per block)
  int n_nodes;
  int do_mid, do_pre, do_dec;   // phases
  int update_pos;
  int kind, kn;
  float* sl_next;               // [2 (SL,SR)][N][64] accumulators to clear for the `pre` block
  float *fl_next, *fr_next;     // [N][64] node_ffn_{left,right} outputs of the `pre` block
  float* x_save;                // optional [N][256]: copy of the `pre` block's input h_node (backward pass)
  float* agg_save;              // optional [N][256]: copy of the `mid` block's aggregated messages
  const float* pos_cur;
  float* pos_nxt;
  float* pred_node;             // kind 1 decode target [N][kn]
};

template <int N>
__device__ __forceinline__ void store_table(const float (&acc)[8][N / 32], float* __restrict__ table, int row0,
                                            int n_rows, int warp, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = row0 + warp * 8 + i;
    if (n < n_rows) store_cols<N>(acc[i], table + (size_t)n * N, lane);
  }
}
__device__ __forceinline__ void store_table_blocked(const float (&acc)[8][8], float* __restrict__ table, int row0,
                                                    int n_rows, int warp, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = row0 + warp * 8 + i;
    if (n < n_rows) store_cols_blocked256(acc[i], table, n, lane);
  }
}

// Per-node hoisted projections of one block from the h_node tile X (smem): everything a first-layer Linear
// of the reference applies to a gathered node feature is applied per node here and gathered afterwards.
__device__ __forceinline__ void node_pre_phase(const float* __restrict__ blob, const BlkOff& off, const Tables& tb,
                                               const float* X, float* A, float* Ws, const float* tns, int row0,
                                               int n_nodes, float* fl_dst, float* fr_dst, float* sl_clear) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {  // node_net(x)                                                                graph.py:39
    float acc[8][8];
    tile_gemm<D, D>(acc, X, D, W_(NB_NN1_W), Ws);
    add_rowvec<D>(acc, W_(NB_NN1_B), lane);
    layernorm_rows<D, true>(acc, W_(NB_NN1_G), W_(NB_NN1_BE), lane);
    store_smem<D>(acc, A, D, warp, lane);
    tile_gemm<D, D>(acc, A, D, W_(NB_NN2_W), Ws);
    add_rowvec<D>(acc, W_(NB_NN2_B), lane);
    store_table<D>(acc, tb.hn, row0, n_nodes, warp, lane);
    store_table_blocked(acc, tb.hnb, row0, n_nodes, warp, lane);
  }
  {  // node + time + bias part of gate.net.0 (hoisted from the per-edge cat)      graph.py:46
    float acc[8][8];
    tile_gemm<D, D>(acc, X, D, W_(NB_GX_W), Ws);
    add_rowvec<D>(acc, W_(NB_G1_B), lane);
    add_scaled_rowvec<D>(acc, W_(NB_GT_W), tns + warp * 8, lane);
    store_table<D>(acc, tb.gx, row0, n_nodes, warp, lane);
    store_table_blocked(acc, tb.gxb, row0, n_nodes, warp, lane);
  }
  {  // centroid_lin(x)                                                            graph.py:51
    float acc[8][8];
    tile_gemm<D, D>(acc, X, D, W_(NB_CEN_W), Ws);
    add_rowvec<D>(acc, W_(NB_CEN_B), lane);
    store_table<D>(acc, tb.cen, row0, n_nodes, warp, lane);
  }
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {   // EdgeBlock hoists                     graph.py:135,139,288-289
    {
      float acc[8][4];
      tile_gemm<D, 128>(acc, X, D, side ? W_(ER_NL_W) : W_(EL_NL_W), Ws);
      store_table<128>(acc, side ? tb.nlr : tb.nll, row0, n_nodes, warp, lane);
    }
    {
      float acc[8][1];
      tile_gemm<D, 32>(acc, X, D, side ? W_(ER_GN_W) : W_(EL_GN_W), Ws);
      add_rowvec<32>(acc, side ? W_(ER_G1_B) : W_(EL_G1_B), lane);   // the time column uses the EDGE time: added per edge
      store_table<32>(acc, side ? tb.gnr : tb.gnl, row0, n_nodes, warp, lane);
    }
    {
      float acc[8][2];
      tile_gemm<D, C>(acc, X, D, side ? W_(EB_NFR_W) : W_(EB_NFL_W), Ws);
      add_rowvec<C>(acc, side ? W_(EB_NFR_B) : W_(EB_NFL_B), lane);
      // separately buffered per block: edge_kernel_d(i) still reads block i's values after block i+1's were written
      store_table<C>(acc, side ? fr_dst : fl_dst, row0, n_nodes, warp, lane);
    }
  }
  if (sl_clear) {  // clear the SL/SR accumulators the block's edge_kernel_b will add into
    for (int i = tid; i < TM * C; i += NTHREADS) {
      const int n = row0 + i / C;
      if (n < n_nodes) {
        sl_clear[(size_t)row0 * C + i] = 0.f;
        sl_clear[(size_t)n_nodes * C + (size_t)row0 * C + i] = 0.f;
      }
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) node_kernel(const NodeArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* X = smem;                 // [64][256]
  float* A = X + TM * D;           // [64][256]
  float* Ws = A + TM * D;          // 2*WCHUNK
  float* tns = Ws + 2 * WCHUNK;    // [64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * TM;
  const float* blob = a.blob;
  const Tables& tb = a.tb;

  // load x tile (+ node time)
  for (int i = tid; i < TM * D / 4; i += NTHREADS) {
    const int r = i / (D / 4), c4 = i % (D / 4);
    const int n = row0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < a.n_nodes) v = reinterpret_cast<const float4*>(tb.x + (size_t)n * D)[c4];
    reinterpret_cast<float4*>(X + r * D)[c4] = v;
  }
  if (tid < TM) tns[tid] = (row0 + tid < a.n_nodes) ? tb.tn[row0 + tid] : 0.f;
  __syncthreads();

  if (a.do_mid) {
    const BlkOff& off = a.mid;
    {  // out = LN(centroid_lin(x) + aggr) -> ReLU                                   graph.py:51-54
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = row0 + warp * 8 + i;
        if (n < a.n_nodes) {
          float u[8], v[8];
          load_cols<D>(u, tb.cen + (size_t)n * D, lane);
          load_cols<D>(v, tb.agg + (size_t)n * D, lane);
          if (a.agg_save) store_cols<D>(v, a.agg_save + (size_t)n * D, lane);
#pragma unroll
          for (int j = 0; j < 8; ++j) { acc[i][j] = u[j] + v[j]; v[j] = 0.f; }
          store_cols<D>(v, tb.agg + (size_t)n * D, lane);   // re-arm the accumulator for the next block
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        }
      }
      layernorm_rows<D, true>(acc, W_(NB_LN_G), W_(NB_LN_BE), lane);
      store_smem<D>(acc, A, D, warp, lane);
    }
    {  // h_node = h_node + out_transform(.)                                        graph.py:54,363
      float acc[8][8];
      tile_gemm<D, D>(acc, A, D, W_(NB_OUT_W), Ws);
      add_rowvec<D>(acc, W_(NB_OUT_B), lane);
      combine_smem<D, false>(acc, X, D, warp, lane);
      store_smem<D>(acc, X, D, warp, lane);
      store_table<D>(acc, tb.x, row0, a.n_nodes, warp, lane);
    }
    if (a.update_pos) {  // PosUpdate node-side MLPs on the NEW h_node             graph.py:387-388
#pragma unroll 1
      for (int side = 0; side < 2; ++side) {
        float acc[8][2];
        tile_gemm<D, C>(acc, X, D, side ? W_(PU_RL1_W) : W_(PU_LL1_W), Ws);
        add_rowvec<C>(acc, side ? W_(PU_RL1_B) : W_(PU_LL1_B), lane);
        layernorm_rows<C, true>(acc, side ? W_(PU_RL1_G) : W_(PU_LL1_G), side ? W_(PU_RL1_BE) : W_(PU_LL1_BE), lane);
        store_smem<C>(acc, A, C, warp, lane);
        tile_gemm<C, C>(acc, A, C, side ? W_(PU_RL2_W) : W_(PU_LL2_W), Ws);
        add_rowvec<C>(acc, side ? W_(PU_RL2_B) : W_(PU_LL2_B), lane);
        store_table<C>(acc, side ? tb.rf : tb.lf, row0, a.n_nodes, warp, lane);
      }
      // pos_nxt starts as pos_cur; edge_kernel_d accumulates the forces into it     graph.py:366
      for (int i = tid; i < TM * 3; i += NTHREADS) {
        const int n = row0 + i / 3;
        if (n < a.n_nodes) a.pos_nxt[(size_t)row0 * 3 + i] = a.pos_cur[(size_t)row0 * 3 + i];
      }
    }
  }

  if (a.do_pre) {
    if (a.x_save) {
      for (int i = tid; i < TM * D / 4; i += NTHREADS) {
        const int r = i / (D / 4), c4 = i % (D / 4);
        const int n = row0 + r;
        if (n < a.n_nodes) reinterpret_cast<float4*>(a.x_save + (size_t)n * D)[c4] = reinterpret_cast<const float4*>(X + r * D)[c4];
      }
    }
    node_pre_phase(blob, a.pre, a.tb_pre, X, A, Ws, tns, row0, a.n_nodes, a.fl_next, a.fr_next, a.sl_next);
  }

  if (a.do_dec) {
    const HeadOff& hoff = a.head;
    if (a.kind == 1) {  // node_decoder MLP(256 -> 256 -> Kn)                         model.py:226
      float acc[8][8];
      tile_gemm<D, D>(acc, X, D, H_(NDEC1_W), Ws);
      add_rowvec<D>(acc, H_(NDEC1_B), lane);
      layernorm_rows<D, true>(acc, H_(NDEC1_G), H_(NDEC1_BE), lane);
      store_smem<D>(acc, A, D, warp, lane);
      float o[8][1];
      tile_gemm<D, 32>(o, A, D, H_(NDEC2_W), Ws);
      add_rowvec<32>(o, H_(NDEC2_B), lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = row0 + warp * 8 + i;
        if (n < a.n_nodes && lane < a.kn) a.pred_node[(size_t)n * a.kn + lane] = o[i][0];
      }
    } else if (a.kind == 2) {  // node half of edge_decoder.net.0                     bond_predictor.py:155-160
      float acc[8][2];
      tile_gemm<D, C>(acc, X, D, H_(EDEC1N_W), Ws);
      store_table<C>(acc, tb.dect, row0, a.n_nodes, warp, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// edge kernel B: RBF + edge_embs + NodeBlock edge path + BondFFN left/right     (64 sorted edges / CTA)
// ------------------------------------------------------------------------------------------------
struct EdgeArgs {
  const float* blob;
  BlkOff off;
  HeadOff head;
  Tables tb;
  const int *left, *right;
  int n_nodes, n_edges;
  int update_pos;
  float* sl;                  // [2 (SL,SR)][N][64] scatter targets of this block's BondFFNs
  const float *fl, *fr;       // [N][64] node_ffn_{left,right}(h_node) of this block
  float* ebuf;                // [E][64] e = edge_embs(cat(h_edge, rbf)) of this block
  int skip_nodeblock;         // the NodeBlock edge path runs in tc_nodeblock_fwd_kernel (tensor cores) instead
  float rbf_lo, rbf_hi;
  const float* pos_cur;
  float* pos_nxt;
};

__device__ __forceinline__ void load_edge_meta(const EdgeArgs& a, int q0, int* ls, int* rs, float* tes,
                                               float* rel /* [4][64]: x,y,z,d */) {
  const int tid = threadIdx.x;
  if (tid < TM) {
    const int q = q0 + tid;
    int l = -1, r = -1;
    float t = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, d = 1.f;
    if (q < a.n_edges) {
      l = a.left[q]; r = a.right[q]; t = a.tb.te[q];
      dx = a.pos_cur[l * 3 + 0] - a.pos_cur[r * 3 + 0];
      dy = a.pos_cur[l * 3 + 1] - a.pos_cur[r * 3 + 1];
      dz = a.pos_cur[l * 3 + 2] - a.pos_cur[r * 3 + 2];
      d = sqrtf(dx * dx + dy * dy + dz * dz);
    }
    ls[tid] = l; rs[tid] = r; tes[tid] = t;
    rel[tid] = dx; rel[TM + tid] = dy; rel[2 * TM + tid] = dz; rel[3 * TM + tid] = d;
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) edge_kernel_b(const EdgeArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Es = smem;                  // [64][64]   e = edge_embs(cat(h_edge, rbf))
  float* A = Es + TM * C;            // [64][256]
  float* Bf = A + TM * D;            // [64][256]
  float* Ws = Bf + TM * D;           // 2*WCHUNK
  float* tes = Ws + 2 * WCHUNK;      // [64]
  float* rel = tes + TM;             // [4][64]
  int* ls = reinterpret_cast<int*>(rel + 4 * TM);
  int* rs = ls + TM;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * TM;
  const float* blob = a.blob;
  const BlkOff& off = a.off;
  const Tables& tb = a.tb;

  load_edge_meta(a, q0, ls, rs, tes, rel);
  // IN = [h_edge ; rbf(d)]  as A[64][80]
  constexpr int KI = C + G;
  for (int i = tid; i < TM * C / 4; i += NTHREADS) {
    const int r = i / (C / 4), c4 = i % (C / 4);
    const int q = q0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < a.n_edges) v = reinterpret_cast<const float4*>(tb.hedge + (size_t)q * C)[c4];
    *reinterpret_cast<float4*>(A + r * KI + c4 * 4) = v;
  }
  __syncthreads();
  {
    const HeadOff& hoff = a.head;
    const float* ro = H_(RBF_OFFSET);
    const float* rc = H_(RBF_COEFF);
    for (int i = tid; i < TM * G; i += NTHREADS) {
      const int r = i / G, k = i % G;
      A[r * KI + C + k] = smear(rel[3 * TM + r], a.rbf_lo, a.rbf_hi, ro[k], rc[k]);
    }
  }
  __syncthreads();
  const int* lw = ls + warp * 8;
  const int* rw = rs + warp * 8;
  {  // e = edge_embs[i](cat)                                                        graph.py:354-357
    float acc[8][2];
    tile_gemm<KI, C>(acc, A, KI, W_(EE_W), Ws);
    add_rowvec<C>(acc, W_(EE_B), lane);
    store_smem<C>(acc, Es, C, warp, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int q = q0 + warp * 8 + i;
      if (q < a.n_edges) store_cols<C>(acc[i], a.ebuf + (size_t)q * C, lane);
    }
  }
  // ---- NodeBlock edge path                                                       graph.py:42-50
  if (!a.skip_nodeblock) {
    float acc[8][8];
    tile_gemm<C, D>(acc, Es, C, W_(NB_EN1_W), Ws);           // edge_net.net.0
    add_rowvec<D>(acc, W_(NB_EN1_B), lane);
    layernorm_rows<D, true>(acc, W_(NB_EN1_G), W_(NB_EN1_BE), lane);
    store_smem<D>(acc, A, D, warp, lane);
    tile_gemm<D, D>(acc, A, D, W_(NB_EN2_W), Ws);            // edge_net.net.3
    add_rowvec<D>(acc, W_(NB_EN2_B), lane);
    gather_rows<D, true>(acc, tb.hn, rw, lane);              // * node_net(x)[col]
    store_smem<D>(acc, Bf, D, warp, lane);
    tile_gemm<C, D>(acc, Es, C, W_(NB_GE_W), Ws);            // gate.net.0, edge columns
    gather_rows<D, false>(acc, tb.gx, rw, lane);             //   + hoisted node/time/bias part
    layernorm_rows<D, true>(acc, W_(NB_G1_G), W_(NB_G1_BE), lane);
    store_smem<D>(acc, A, D, warp, lane);
    tile_gemm<D, D>(acc, A, D, W_(NB_G2_W), Ws);             // gate.net.3
    add_rowvec<D>(acc, W_(NB_G2_B), lane);
    sigmoid_rows<D>(acc);
    store_smem<D>(acc, A, D, warp, lane);
    tile_gemm<D, D>(acc, Bf, D, W_(NB_MSG_W), Ws);           // msg_net
    add_rowvec<D>(acc, W_(NB_MSG_B), lane);
    combine_smem<D, true>(acc, A, D, warp, lane);
    scatter_add_rows<D, true>(acc, tb.agg, lw, lane);        // scatter_sum over row (= left)
  }
  // ---- EdgeBlock: bond_ffn_left (node = left, scattered to right) and _right      graph.py:278-284
  float* sl = a.sl;
  float* sr = sl + (size_t)a.n_nodes * C;
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    const int* nw = side ? rw : lw;   // node feeding the FFN
    float inter[8][2];
    {
      float acc[8][4];
      tile_gemm<C, 128>(acc, Es, C, side ? W_(ER_BL_W) : W_(EL_BL_W), Ws);      // bond_linear
      gather_rows<128, true>(acc, side ? tb.nlr : tb.nll, nw, lane);            // * node_linear(h_node)[.]
      store_smem<128>(acc, A, 128, warp, lane);
      tile_gemm<128, 128>(acc, A, 128, side ? W_(ER_I1_W) : W_(EL_I1_W), Ws);   // inter_module
      add_rowvec<128>(acc, side ? W_(ER_I1_B) : W_(EL_I1_B), lane);
      layernorm_rows<128, true>(acc, side ? W_(ER_I1_G) : W_(EL_I1_G), side ? W_(ER_I1_BE) : W_(EL_I1_BE), lane);
      store_smem<128>(acc, A, 128, warp, lane);
      tile_gemm<128, C>(inter, A, 128, side ? W_(ER_I2_W) : W_(EL_I2_W), Ws);
      add_rowvec<C>(inter, side ? W_(ER_I2_B) : W_(EL_I2_B), lane);
    }
    {
      float g1[8][1];
      tile_gemm<C, 32>(g1, Es, C, side ? W_(ER_GB_W) : W_(EL_GB_W), Ws);        // gate.net.0 bond columns
      gather_rows<32, false>(g1, side ? tb.gnr : tb.gnl, nw, lane);
      add_scaled_rowvec<32>(g1, side ? W_(ER_GT_W) : W_(EL_GT_W), tes + warp * 8, lane);
      layernorm_rows<32, true>(g1, side ? W_(ER_G1_G) : W_(EL_G1_G), side ? W_(ER_G1_BE) : W_(EL_G1_BE), lane);
      store_smem<32>(g1, Bf, 32, warp, lane);
      float g2[8][2];
      tile_gemm<32, C>(g2, Bf, 32, side ? W_(ER_G2_W) : W_(EL_G2_W), Ws);
      add_rowvec<C>(g2, side ? W_(ER_G2_B) : W_(EL_G2_B), lane);
      sigmoid_rows<C>(g2);
#pragma unroll
      for (int i = 0; i < 8; ++i) { inter[i][0] *= g2[i][0]; inter[i][1] *= g2[i][1]; }
    }
    if (side == 0) scatter_add_rows<C, false>(inter, sl, rw, lane);   // scatter over right_node
    else           scatter_add_rows<C, true>(inter, sr, lw, lane);    // scatter over left_node
  }
}

// ------------------------------------------------------------------------------------------------
// edge kernel D: EdgeBlock tail (-> new h_edge) + PosUpdate edge path (-> pos)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1) edge_kernel_d(const EdgeArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Es = smem;                  // [64][64]  e
  float* E2 = Es + TM * C;           // [64][64]  new h_edge
  float* P = E2 + TM * C;            // [64][64]  left_feat * right_feat
  float* A = P + TM * C;             // [64][256]
  float* Ws = A + TM * D;
  float* tes = Ws + 2 * WCHUNK;
  float* rel = tes + TM;
  int* ls = reinterpret_cast<int*>(rel + 4 * TM);
  int* rs = ls + TM;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * TM;
  const float* blob = a.blob;
  const BlkOff& off = a.off;
  const Tables& tb = a.tb;

  load_edge_meta(a, q0, ls, rs, tes, rel);
  for (int i = tid; i < TM * C / 4; i += NTHREADS) {
    const int r = i / (C / 4), c4 = i % (C / 4);
    const int q = q0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < a.n_edges) v = reinterpret_cast<const float4*>(a.ebuf + (size_t)q * C)[c4];
    reinterpret_cast<float4*>(Es + r * C)[c4] = v;
  }
  __syncthreads();
  const int* lw = ls + warp * 8;
  const int* rw = rs + warp * 8;
  const float* sl = a.sl;
  const float* sr = sl + (size_t)a.n_nodes * C;
  {  // graph.py:286-294
    float acc[8][2];
    tile_gemm<C, C>(acc, Es, C, W_(EB_SELF_W), Ws);
    add_rowvec<C>(acc, W_(EB_SELF_B), lane);
    gather_rows<C, false>(acc, sl, lw, lane);       // scatter_sum(msg_left, right)[left]
    gather_rows<C, false>(acc, sr, rw, lane);       // scatter_sum(msg_right, left)[right]
    gather_rows<C, false>(acc, a.fl, lw, lane);     // node_ffn_left(h_node[left])
    gather_rows<C, false>(acc, a.fr, rw, lane);     // node_ffn_right(h_node[right])
    layernorm_rows<C, true>(acc, W_(EB_LN_G), W_(EB_LN_BE), lane);
    store_smem<C>(acc, A, C, warp, lane);
    tile_gemm<C, C>(acc, A, C, W_(EB_OUT_W), Ws);
    add_rowvec<C>(acc, W_(EB_OUT_B), lane);
    combine_smem<C, false>(acc, Es, C, warp, lane);  // h_edge = h_edge + EdgeBlock(.)  graph.py:362
    store_smem<C>(acc, E2, C, warp, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int q = q0 + warp * 8 + i;
      if (q < a.n_edges) store_cols<C>(acc[i], tb.hedge + (size_t)q * C, lane);
    }
  }
  if (!a.update_pos) return;
  // ---- PosUpdate                                                                 graph.py:384-396
  {
    float p[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { p[i][0] = 1.f; p[i][1] = 1.f; }
    gather_rows<C, true>(p, tb.lf, lw, lane);
    gather_rows<C, true>(p, tb.rf, rw, lane);
    store_smem<C>(p, P, C, warp, lane);
  }
  float w_inter[8], w_gate[8];
  {
    float acc[8][8];
    tile_gemm<C, D>(acc, E2, C, W_(PU_PB_W), Ws);            // edge_lin.bond_linear(h_edge)
    store_smem<D>(acc, A, D, warp, lane);
    tile_gemm<C, D>(acc, P, C, W_(PU_PN_W), Ws);             // edge_lin.node_linear(lf * rf)
    combine_smem<D, true>(acc, A, D, warp, lane);
    store_smem<D>(acc, A, D, warp, lane);
    tile_gemm<D, D>(acc, A, D, W_(PU_I1_W), Ws);             // inter_module
    add_rowvec<D>(acc, W_(PU_I1_B), lane);
    layernorm_rows<D, true>(acc, W_(PU_I1_G), W_(PU_I1_BE), lane);
    dot_rows<D>(w_inter, acc, W_(PU_I2_W), lane);
  }
  {
    float g1[8][1];
    tile_gemm<C, 32>(g1, E2, C, W_(PU_GB_W), Ws);            // gate.net.0: bond | node | time columns
    tile_gemm<C, 32, true>(g1, P, C, W_(PU_GN_W), Ws);
    add_rowvec<32>(g1, W_(PU_G1_B), lane);
    add_scaled_rowvec<32>(g1, W_(PU_GT_W), tes + warp * 8, lane);
    layernorm_rows<32, true>(g1, W_(PU_G1_G), W_(PU_G1_BE), lane);
    dot_rows<32>(w_gate, g1, W_(PU_G2_W), lane);
  }
  const float b_i2 = W_(PU_I2_B)[0], b_g2 = W_(PU_G2_B)[0];
  // force = w * rel / d / (d + 1); delta_pos = scatter_sum(force, left)            graph.py:393-394
  if (lane < 3) {
    int cur = lw[0];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp * 8 + i;
      const int n = lw[i];
      const float w = (w_inter[i] + b_i2) * (1.f / (1.f + expf(-(w_gate[i] + b_g2))));
      const float d = rel[3 * TM + r];
      const float f = w * rel[lane * TM + r] / d / (d + 1.f);
      if (n != cur) {
        if (cur >= 0) atomicAdd(a.pos_nxt + (size_t)cur * 3 + lane, s);
        cur = n; s = 0.f;
      }
      s += f;
    }
    if (cur >= 0) atomicAdd(a.pos_nxt + (size_t)cur * 3 + lane, s);
  }
}

// ------------------------------------------------------------------------------------------------
// edge decode: half-edge p in caller order -> h_edge[inv[p]] + h_edge[inv[p + Eh]]
// ------------------------------------------------------------------------------------------------
struct DecArgs {
  const float* blob;
  HeadOff head;
  Tables tb;
  const int *left, *right, *inv;
  int n_half, kind, ke;
  float* out;   // [Eh][ke]
};

__global__ void __launch_bounds__(NTHREADS, 1) edge_decode_kernel(const DecArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Es = smem;               // [64][64]
  float* A = Es + TM * C;         // [64][64]
  float* Ws = A + TM * C;
  int* ls = reinterpret_cast<int*>(Ws + 2 * WCHUNK);
  int* rs = ls + TM;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p0 = blockIdx.x * TM;
  const float* blob = a.blob;
  const HeadOff& hoff = a.head;
  const Tables& tb = a.tb;
  if (tid < TM) {
    const int p = p0 + tid;
    int l = -1, r = -1;
    if (p < a.n_half) { const int q = a.inv[p]; l = a.left[q]; r = a.right[q]; }
    ls[tid] = l; rs[tid] = r;
  }
  for (int i = tid; i < TM * C / 4; i += NTHREADS) {
    const int r = i / (C / 4), c4 = i % (C / 4);
    const int p = p0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p < a.n_half) {
      const float4 u = reinterpret_cast<const float4*>(tb.hedge + (size_t)a.inv[p] * C)[c4];
      const float4 w = reinterpret_cast<const float4*>(tb.hedge + (size_t)a.inv[p + a.n_half] * C)[c4];
      v = make_float4(u.x + w.x, u.y + w.y, u.z + w.z, u.w + w.w);
    }
    reinterpret_cast<float4*>(Es + r * C)[c4] = v;
  }
  __syncthreads();
  float acc[8][2];
  tile_gemm<C, C>(acc, Es, C, H_(EDEC1_W), Ws);
  add_rowvec<C>(acc, H_(EDEC1_B), lane);
  if (a.kind == 2) {   // + W_node (h_node[l] + h_node[r])                           bond_predictor.py:155-160
    gather_rows<C, false>(acc, tb.dect, ls + warp * 8, lane);
    gather_rows<C, false>(acc, tb.dect, rs + warp * 8, lane);
  }
  layernorm_rows<C, true>(acc, H_(EDEC1_G), H_(EDEC1_BE), lane);
  store_smem<C>(acc, A, C, warp, lane);
  float o[8][1];
  if (a.kind == 2) {
    tile_gemm<C, C>(acc, A, C, H_(EDEC2_W), Ws);
    add_rowvec<C>(acc, H_(EDEC2_B), lane);
    layernorm_rows<C, true>(acc, H_(EDEC3_G), H_(EDEC3_BE), lane);
    store_smem<C>(acc, A, C, warp, lane);
    tile_gemm<C, 32>(o, A, C, H_(EDEC3_W), Ws);
    add_rowvec<32>(o, H_(EDEC3_B), lane);
  } else {
    tile_gemm<C, 32>(o, A, C, H_(EDEC2_W), Ws);
    add_rowvec<32>(o, H_(EDEC2_B), lane);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int p = p0 + warp * 8 + i;
    if (p < a.n_half && lane < a.ke) a.out[(size_t)p * a.ke + lane] = o[i][0];
  }
}

// kind 0 epilogue: h_edge_out[perm[q]] = hedge[q]
__global__ void edge_unsort_kernel(int n_edges, const int* __restrict__ perm, const float* __restrict__ hedge,
                                   float* __restrict__ out) {
  const int q = blockIdx.x * 4 + (threadIdx.x >> 6);
  const int c = threadIdx.x & 63;
  if (q < n_edges) out[(size_t)perm[q] * C + c] = hedge[(size_t)q * C + c];
}

#include "tc_nodeblock.cuh"
#include "tc_bondffn.cuh"
#include "tc_edge_d.cuh"
#include "tc_nodeblock16.cuh"
#include "tc_nodeblock_bwd16.cuh"
#include "tc_bondffn2.cuh"
#include "tc_node.cuh"
#include "tc_edge_tail_bwd.cuh"
#include "tc_bondffn_bwd2.cuh"
#include "tc_bondffn_bwd3.cuh"
#include "tc_bwd_node.cuh"

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
thread_local char g_err[512] = "";
int64_t g_launches = 0;
bool g_profiling = false;
int g_dbg_sel = 0;                    // which kernel fills the stamp buffer: 0 nodeblock fwd, 1 nodeblock bwd, 2 bondffn fwd
long long* g_dbg_stamps = nullptr;   // optional device buffer for in-kernel clock64 phase stamps (mdb_debug_set_buffer)
struct ProfRec { int cls; cudaEvent_t a, b; };
std::vector<ProfRec> g_prof;

// Every launch goes through LAUNCH so that it is counted and, when profiling, timed on its own stream.
#define LAUNCH(kcls_, st, ...)                                             \
  do {                                                                    \
    ProfRec pr_;                                                          \
    if (g_profiling) {                                                    \
      pr_.cls = (kcls_);                                                  \
      cudaEventCreate(&pr_.a); cudaEventCreate(&pr_.b);                   \
      cudaEventRecord(pr_.a, (st));                                       \
    }                                                                     \
    __VA_ARGS__;                                                          \
    ++g_launches;                                                         \
    if (g_profiling) { cudaEventRecord(pr_.b, (st)); g_prof.push_back(pr_); } \
  } while (0)

int fail(int code, const char* fmt, const char* extra = "") {
  snprintf(g_err, sizeof(g_err), fmt, extra);
  return code;
}

#define CUDA_TRY(expr)                                                        \
  do {                                                                        \
    cudaError_t e_ = (expr);                                                  \
    if (e_ != cudaSuccess) return fail(MDB_ECUDA, #expr ": %s", cudaGetErrorString(e_)); \
  } while (0)

inline size_t al(size_t n) { return (n + 31) & ~size_t(31); }
// grid of a persistent kernel (one CTA per SM, each looping over tiles)
inline int persistent_grid(int64_t n_tiles) {
  static const int sms = []() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
  }();
  return (int)(n_tiles < sms ? (n_tiles > 0 ? n_tiles : 1) : sms);
}
// A second, HIGH-PRIORITY stream (per device) for the few-CTA node kernels that may run next to an edge kernel of the main
// chain (their CTAs are then placed first as SMs free up), and the two events that fork / join it.
// Fork / join by events keeps the pair capturable into one CUDA graph (MolDiff.graphed_step).
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
SideStream* side_stream() {
  static SideStream ss[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& x = ss[dev];
  if (!x.s) {
    int lo = 0, hi = 0;                      // (numerically lower = higher priority)
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&x.s, cudaStreamNonBlocking, hi) != cudaSuccess) { x.s = nullptr; return nullptr; }
    cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming);
  }
  return &x;
}
inline int64_t pad64(int64_t n) { return (n + 63) / 64 * 64; }   // node-blocked tables (tile_engine.cuh: blk_off)

size_t carve(Tables& tb, float* base, int64_t N, int64_t E) {
  size_t o = 0;
  auto take = [&](size_t n) { float* p = base ? base + o : nullptr; o += al(n); return p; };
  tb.x = take(N * D); tb.agg = take(N * D); tb.hn = take(N * D); tb.gx = take(N * D); tb.cen = take(N * D);
  tb.hnb = take(pad64(N) * D); tb.gxb = take(pad64(N) * D);
  tb.nll = take(N * 128); tb.nlr = take(N * 128);
  tb.gnl = take(N * 32); tb.gnr = take(N * 32);
  tb.fl = take(2 * N * C); tb.fr = take(2 * N * C); tb.lf = take(N * C); tb.rf = take(N * C); tb.dect = take(N * C);
  tb.slsr = take(4 * N * C);
  tb.pos0 = take(N * 3); tb.pos1 = take(N * 3);
  tb.tn = take(N);
  tb.hedge = take(E * C); tb.ebuf = take(E * C);
  tb.te = take(E);
  return o;
}

// Extra workspace of a forward that will be followed by the input-gradient backward (bond predictor
// guidance): per-block inputs that are cheap to keep (per-node tensors and e = edge_embs output); every
// other per-edge activation is recomputed tile by tile in the backward kernels.
struct Saved {
  float *e;                  // [L][E][64]   e_i
  float *x;                  // [L][N][256]  h_node entering block i
  float *agg;                // [L][N][256]  aggregated NodeBlock messages of block i
  float *slsr;               // [L][2][N][64]
  float *tabs;               // [L][N * TAB_FLOATS]  per-node hoisted tables of every block (hn, gx, cen, nl, gn, fl, fr)
  float *dx;                 // [pad64(N)][256]  running d/d h_node (row-major: bwd_node_kernel; node-blocked: tc_bwd_node_kernel)
  float *dh, *de;            // [E][64]   d/d h_edge (block output -> block input), d/d e
  float *dg;                 // [E][16]   d/d rbf features, summed over blocks
  float *dul, *dur;          // [2][N][64]   sum_{p: l_p = n} du_p, sum_{p: r_p = n} du_p; parity = block index & 1
  float *dagg, *dgx, *dhn;   // [N][256]
  float *dnl;                // [2][N][128]
  float *dgn;                // [2][N][32]
  float *ddect;              // [N][64]
  float *gamax;              // [32]: word 0 = bit pattern of max |d_logits| of the current backward call
  float *hnb, *gxb;          // [L][pad64(N)][256]  node-blocked copies of the saved hn / gx tables
  float *scr_he, *scr_dm;    // [persistent CTAs * 128][256] each: tc_nodeblock_bwd16 scratch slabs (he fp32 / d msg operand planes)
};

constexpr int TAB_FLOATS = 3 * D + 2 * 128 + 2 * 32 + 2 * C;   // per node

// Tables whose per-block members (hn, gx, cen, nl, gn, fl, fr) point into one saved slab
Tables with_block_tables(const Tables& tb, float* slab, int64_t N, float* hnb, float* gxb) {
  Tables t = tb;
  t.hnb = hnb; t.gxb = gxb;
  size_t o = 0;
  auto take = [&](size_t n) { float* p = slab + o; o += n; return p; };
  t.hn = take(N * D); t.gx = take(N * D); t.cen = take(N * D);
  t.nll = take(N * 128); t.nlr = take(N * 128); t.gnl = take(N * 32); t.gnr = take(N * 32);
  t.fl = take(N * C); t.fr = take(N * C);
  return t;
}

size_t carve_saved(Saved& sv, float* base, int64_t N, int64_t E, int64_t L) {
  size_t o = 0;
  auto take = [&](size_t n) { float* p = base ? base + o : nullptr; o += al(n); return p; };
  sv.e = take(L * E * C); sv.x = take(L * N * D); sv.agg = take(L * N * D); sv.slsr = take(L * 2 * N * C);
  sv.tabs = take(L * N * TAB_FLOATS);
  sv.dx = take(pad64(N) * D); sv.dh = take(E * C); sv.de = take(E * C); sv.dg = take(E * G);
  sv.dul = take(2 * N * C); sv.dur = take(2 * N * C);   // two parity buffers each (block i / block i + 1)
  sv.dagg = take(N * D); sv.dgx = take(pad64(N) * D); sv.dhn = take(pad64(N) * D);   // dgx / dhn: row-major or node-blocked
  sv.hnb = take(L * pad64(N) * D); sv.gxb = take(L * pad64(N) * D);
  sv.dnl = take(2 * N * 128); sv.dgn = take(2 * N * 32); sv.ddect = take(N * C);
  sv.gamax = take(32);
  const int64_t scr_rows = (int64_t)persistent_grid((E + 127) / 128) * 128;   // one 128-row slab per persistent CTA
  sv.scr_he = take(scr_rows * D); sv.scr_dm = take(scr_rows * D);
  return o;
}

constexpr size_t SMEM_NODE = (2 * TM * D + 2 * WCHUNK + TM) * sizeof(float);
constexpr size_t SMEM_EDGE_B = (TM * C + 2 * TM * D + 2 * WCHUNK + TM + 4 * TM + 2 * TM) * sizeof(float);
constexpr size_t SMEM_EDGE_D = (3 * TM * C + TM * D + 2 * WCHUNK + TM + 4 * TM + 2 * TM) * sizeof(float);
constexpr size_t SMEM_DEC = (2 * TM * C + 2 * WCHUNK + 2 * TM) * sizeof(float);

int ensure_attrs() {
  static bool done[64] = {};       // per device: the max-dynamic-smem attribute is a property of (function, device)
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && done[dev]) return MDB_OK;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(MDB_EARCH, "moldiff_b200 kernels are built for sm_100a only%s");
  CUDA_TRY(cudaFuncSetAttribute(node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_NODE));
  CUDA_TRY(cudaFuncSetAttribute(edge_kernel_b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_EDGE_B));
  CUDA_TRY(cudaFuncSetAttribute(edge_kernel_d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_EDGE_D));
  CUDA_TRY(cudaFuncSetAttribute(edge_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_DEC));
  CUDA_TRY(cudaFuncSetAttribute(tc_nodeblock_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_NB));
  CUDA_TRY(cudaFuncSetAttribute(tc_bondffn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_FFN));
  CUDA_TRY(cudaFuncSetAttribute(tc_bondffn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_FFN));
  CUDA_TRY(cudaFuncSetAttribute(tc_bondffn_fwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_FFN2));
  CUDA_TRY(cudaFuncSetAttribute(tc_bondffn_fwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_FFN2));
  CUDA_TRY(cudaFuncSetAttribute(tc_bondffn_fwd2_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CUDA_TRY(cudaFuncSetAttribute(tc_bondffn_fwd2_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CUDA_TRY(cudaFuncSetAttribute(tc_edge_d_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_EDGE_D));
  CUDA_TRY(cudaFuncSetAttribute(tc_edge_d_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_EDGE_D));
  CUDA_TRY(cudaFuncSetAttribute(tc_nodeblock_fwd16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_NB16));
  CUDA_TRY(cudaFuncSetAttribute(tc_nodeblock_fwd16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_NB16));
  CUDA_TRY(cudaFuncSetAttribute(tc_node_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_NODE));
  CUDA_TRY(cudaFuncSetAttribute(tc_node_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_NODE));
  if (dev >= 0 && dev < 64) done[dev] = true;
  return MDB_OK;
}

void fill_blk(BlkOff& b, const mdb_net_desc* net, int i) {
  for (int s = 0; s < MDB_NUM_BLOCK_SLOTS; ++s) b.o[s] = (int)net->block_off[i][s];
}

struct FwdIn {
  const float *h_node, *pos, *h_edge;          // inputs (meaning depends on kind)
  const float *node_time, *edge_time;          // kind 0
  const int64_t *batch_node, *batch_edge, *t;  // kind 1/2
  float *out_node, *out_pos, *out_edge;        // kind 0: h_node/pos/h_edge; kind 1: preds; kind 2: logits in out_edge
  int save;                                    // keep per-block inputs for the backward pass
};

// Accumulation order of a network's forward GEMMs (tc_pipe.cuh: gemm): cross-first for the bond predictor, whose forward
// activations feed the guidance gradient (the hardware's truncating accumulate costs 1.4e-4 there in the interleaved order,
// 1e-5 in this one); MDB_CROSS_FIRST=0 / 1 forces it off / on for every network (A/B runs).
bool cross_first(const mdb_net_desc* net) {
  static const int env = []() { const char* e = getenv("MDB_CROSS_FIRST"); return e == nullptr ? -1 : (e[0] != '0' ? 1 : 0); }();
  return env >= 0 ? env == 1 : net->kind == 2;
}

// node kernel launch: tensor-core version when its operand images are packed, FFMA version otherwise
int launch_node(const mdb_net_desc* net, const NodeArgs& na, int blk_mid, int blk_pre, int node_tiles, cudaStream_t st) {
  // Per-node tables are gathered by every edge of the node, so an error in them is COHERENT over ~25 edges instead of averaging
  // out: with the bond predictor's node path on tensor cores the guidance gradient sits at 4e-5..6e-5 of fp64 (crossed-path
  // runs, profiles/r02_tc_numerics.txt), on the fp32 FFMA node kernel at 5e-6 -- so the bond predictor (kind 2) takes the
  // FFMA node kernel (N rows only: ~10 % of its forward) and the denoiser the tensor-core one.  MDB_TC_NODE=0 / 1 forces
  // FFMA / tensor cores for every network (A/B runs).
  static const int tc_node_env = []() { const char* e = getenv("MDB_TC_NODE"); return e == nullptr ? -1 : (e[0] != '0' ? 1 : 0); }();
  const bool tc_node_on = tc_node_env >= 0 ? tc_node_env == 1 : net->kind != 2;
  const int ref_blk = blk_pre >= 0 ? blk_pre : blk_mid;
  const bool tc = tc_node_on && net->tc_blob != nullptr && net->blob_host != nullptr && ref_blk >= 0 &&
                  net->tc_block_off[ref_blk][MDB_T_NB_OUT] >= 0;
  if (!tc) {
    LAUNCH(MDB_K_node, st, (node_kernel<<<node_tiles, NTHREADS, SMEM_NODE, st>>>(na)));
    return MDB_OK;
  }
  TcNodeArgs ta;
  memset(&ta, 0, sizeof(ta));
  ta.tc_blob = reinterpret_cast<const uint8_t*>(net->tc_blob);
  for (int s = 0; s < MDB_NUM_TC_SLOTS; ++s) {
    ta.mid.o[s] = blk_mid >= 0 ? net->tc_block_off[blk_mid][s] : -1;
    ta.pre.o[s] = blk_pre >= 0 ? net->tc_block_off[blk_pre][s] : -1;
  }
  for (int s = 0; s < MDB_NUM_TC_HEAD_SLOTS; ++s) ta.hd.o[s] = net->tc_head_off[s];
  ta.tb = na.tb; ta.tb_pre = na.tb_pre; ta.n_nodes = na.n_nodes;
  ta.do_mid = na.do_mid; ta.do_pre = na.do_pre; ta.do_dec = na.do_dec; ta.update_pos = na.update_pos;
  ta.kind = na.kind; ta.kn = na.kn;
  ta.sl_next = na.sl_next; ta.fl_next = na.fl_next; ta.fr_next = na.fr_next; ta.x_save = na.x_save; ta.agg_save = na.agg_save;
  ta.pos_cur = na.pos_cur; ta.pos_nxt = na.pos_nxt; ta.pred_node = na.pred_node;
  fill_node_vecs(ta.v, net->blob_host, na.do_mid ? &na.mid : nullptr, na.do_pre ? &na.pre : nullptr,
                 na.do_dec ? &na.head : nullptr, na.kind, na.update_pos != 0);
  const int grid = (na.n_nodes + tc::ROWS - 1) / tc::ROWS;
  if (cross_first(net)) LAUNCH(MDB_K_tc_node, st, (tc_node_kernel<true><<<grid, TC_NB_THREADS, SMEM_TC_NODE, st>>>(ta)));
  else LAUNCH(MDB_K_tc_node, st, (tc_node_kernel<false><<<grid, TC_NB_THREADS, SMEM_TC_NODE, st>>>(ta)));
  return MDB_OK;
}

int run_forward(const mdb_net_desc* net, const mdb_plan* plan, const FwdIn& in, float* workspace,
                size_t workspace_bytes, cudaStream_t st) {
  if (!net || !plan || !workspace) return fail(MDB_EINVAL, "null argument%s");
  const int N = plan->n_nodes, E = plan->n_edges, L = net->num_blocks;
  if (N <= 0 || E < 0) return fail(MDB_EINVAL, "empty graph%s");
  if (L < 1 || L > MDB_MAX_BLOCKS) return fail(MDB_EINVAL, "num_blocks out of range%s");
  if (net->kind != 0 && (plan->n_half * 2 != E)) return fail(MDB_EINVAL, "edges must be (half, flipped half)%s");
  if (net->num_node_types > 32 || net->num_edge_types > 32) return fail(MDB_EINVAL, "too many types%s");
  if (net->time_dim < 0 || net->time_dim >= C) return fail(MDB_EINVAL, "time_dim out of range%s");
  if (mdb_workspace_bytes(N, E, in.save ? 1 : 0, L) > workspace_bytes) return fail(MDB_EINVAL, "workspace too small%s");
  int rc = ensure_attrs();
  if (rc) return rc;

  Tables tb;
  const size_t tb_floats = carve(tb, workspace, N, E);
  Saved sv;
  memset(&sv, 0, sizeof(sv));
  if (in.save) carve_saved(sv, workspace + tb_floats, N, E, L);
  const size_t NC = (size_t)N * C, ND = (size_t)N * D, EC = (size_t)E * C;
  auto sl_of = [&](int i) { return in.save ? sv.slsr + (size_t)i * 2 * NC : tb.slsr + (size_t)(i & 1) * 2 * NC; };
  // per-block view of the hoisted tables: one shared set normally, one slab per block when saving for the backward
  auto tb_of = [&](int i) {
    return in.save ? with_block_tables(tb, sv.tabs + (size_t)i * N * TAB_FLOATS, N, sv.hnb + (size_t)i * pad64(N) * D,
                                       sv.gxb + (size_t)i * pad64(N) * D)
                   : tb;
  };
  auto fl_of = [&](int i) { return in.save ? tb_of(i).fl : tb.fl + (size_t)(i & 1) * NC; };
  auto fr_of = [&](int i) { return in.save ? tb_of(i).fr : tb.fr + (size_t)(i & 1) * NC; };
  HeadOff head;
  for (int s = 0; s < MDB_NUM_HEAD_SLOTS; ++s) head.o[s] = (int)net->head_off[s];
  const int node_tiles = (N + TM - 1) / TM, edge_tiles = (E + TM - 1) / TM;
  const bool xf = cross_first(net);
  // side-stream overlap of the node kernels with the EdgeBlock tail / PosUpdate kernel (see the block loop); not while the
  // per-kernel event profile is being taken (mdb_profile_begin), MDB_OVERLAP=0 disables it
  static const bool overlap_env = []() { const char* e = getenv("MDB_OVERLAP"); return e == nullptr || e[0] != '0'; }();
  const bool overlap = overlap_env && !g_profiling;
  SideStream* side = overlap ? side_stream() : nullptr;

  LAUNCH(MDB_K_node_init, st,
         (node_init_kernel<<<N, D, 0, st>>>(net->kind, N, net->num_node_types, net->time_dim, net->num_timesteps,
                                            net->blob, head, in.h_node, in.batch_node, in.t, in.node_time, tb.x, tb.tn)));
  // pre(0) -- the hoisted tables of block 0, a 50...99-CTA node kernel that needs node_init only -- runs on the side stream
  // beside the edge initialisation (an E-sized elementwise kernel) and the two small memory operations
  const bool fork0 = overlap && side != nullptr && E > 0;
  if (fork0) {
    CUDA_TRY(cudaEventRecord(side->fork, st));
    CUDA_TRY(cudaStreamWaitEvent(side->s, side->fork, 0));
  }
  if (E > 0) {
    LAUNCH(MDB_K_edge_init, st,
           (edge_init_kernel<<<(E + 3) / 4, 256, 0, st>>>(net->kind, E, net->num_node_types, net->num_edge_types,
                                                          net->time_dim, net->num_timesteps, net->blob, head,
                                                          in.h_edge, in.h_node, in.batch_edge, in.t, in.edge_time,
                                                          plan->left, plan->right, plan->perm, tb.hedge, tb.te)));
  }
  CUDA_TRY(cudaMemsetAsync(tb.agg, 0, (size_t)N * D * sizeof(float), st));
  CUDA_TRY(cudaMemcpyAsync(tb.pos0, in.pos, (size_t)N * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));

  const float* pos_cur = tb.pos0;
  float* pos_nxt = tb.pos1;
  NodeArgs na;
  memset(&na, 0, sizeof(na));
  na.blob = net->blob; na.head = head; na.tb = tb; na.tb_pre = tb_of(0); na.n_nodes = N; na.update_pos = net->update_pos;
  na.kind = net->kind; na.kn = net->num_node_types; na.pred_node = in.out_node;
  // pre(0)
  fill_blk(na.pre, net, 0);
  na.do_mid = 0; na.do_pre = 1; na.do_dec = 0; na.pos_cur = pos_cur; na.pos_nxt = pos_nxt;
  na.sl_next = sl_of(0); na.fl_next = fl_of(0); na.fr_next = fr_of(0);
  na.x_save = in.save ? sv.x : nullptr; na.agg_save = nullptr;
  rc = launch_node(net, na, -1, 0, node_tiles, fork0 ? side->s : st);
  if (rc) return rc;
  if (fork0) {
    CUDA_TRY(cudaEventRecord(side->join, side->s));
    CUDA_TRY(cudaStreamWaitEvent(st, side->join, 0));
  }

  EdgeArgs ea;
  memset(&ea, 0, sizeof(ea));
  ea.blob = net->blob; ea.head = head; ea.tb = tb; ea.left = plan->left; ea.right = plan->right;
  ea.n_nodes = N; ea.n_edges = E; ea.update_pos = net->update_pos;
  ea.rbf_lo = net->rbf_start; ea.rbf_hi = net->rbf_stop;

  for (int i = 0; i < L; ++i) {
    fill_blk(ea.off, net, i);
    const Tables tbi = tb_of(i);
    ea.tb = tbi;
    ea.sl = sl_of(i); ea.fl = fl_of(i); ea.fr = fr_of(i); ea.ebuf = in.save ? sv.e + (size_t)i * EC : tb.ebuf;
    ea.pos_cur = pos_cur; ea.pos_nxt = pos_nxt;
    const bool tc_nb = net->tc_blob != nullptr && net->blob_host != nullptr && net->tc_block_off[i][MDB_T_NB_EN1] >= 0;
    ea.skip_nodeblock = tc_nb ? 1 : 0;
    const bool tc_ffn = tc_nb && net->tc_block_off[i][MDB_T_EE] >= 0;
    if (E > 0 && tc_ffn) {
      TcFfnArgs fa;
      memset(&fa, 0, sizeof(fa));
      fa.tc_blob = reinterpret_cast<const uint8_t*>(net->tc_blob); fa.tb = tbi;
      for (int s = 0; s < MDB_NUM_TC_SLOTS; ++s) fa.tco.o[s] = net->tc_block_off[i][s];
      fa.left = plan->left; fa.right = plan->right; fa.n_nodes = N; fa.n_edges = E;
      fa.pos = pos_cur; fa.rbf_lo = net->rbf_start; fa.rbf_hi = net->rbf_stop; fa.ebuf = ea.ebuf; fa.sl = ea.sl;
      fill_ffn_vecs(fa.v, net->blob_host, ea.off, head);
      fa.dbg = g_dbg_sel == 2 ? g_dbg_stamps : nullptr;
      // default: the two-CTAs-per-SM kernel (tc_bondffn2.cuh); MDB_TC_FFN2=0 -> the one-tile-per-SM kernel (A/B runs)
      static const bool ffn2 = []() { const char* e = getenv("MDB_TC_FFN2"); return e == nullptr || e[0] != '0'; }();
      const int ffn_grid = (E + tc::ROWS - 1) / tc::ROWS;
      if (ffn2) {
        if (xf) LAUNCH(MDB_K_tc_bondffn, st, (tc_bondffn_fwd2_kernel<true><<<ffn_grid, TC_NB_THREADS, SMEM_TC_FFN2, st>>>(fa)));
        else LAUNCH(MDB_K_tc_bondffn, st, (tc_bondffn_fwd2_kernel<false><<<ffn_grid, TC_NB_THREADS, SMEM_TC_FFN2, st>>>(fa)));
      } else {
        if (xf) LAUNCH(MDB_K_tc_bondffn, st, (tc_bondffn_fwd_kernel<true><<<ffn_grid, TC_NB_THREADS, SMEM_TC_FFN, st>>>(fa)));
        else LAUNCH(MDB_K_tc_bondffn, st, (tc_bondffn_fwd_kernel<false><<<ffn_grid, TC_NB_THREADS, SMEM_TC_FFN, st>>>(fa)));
      }
    } else if (E > 0) {
      LAUNCH(MDB_K_edge_b, st, (edge_kernel_b<<<edge_tiles, NTHREADS, SMEM_EDGE_B, st>>>(ea)));
    }
    // EdgeBlock tail (+ PosUpdate) of block i.  Without a position update (bond predictor) it depends on the BondFFN kernel
    // above only -- not on the NodeBlock / node kernels below (h_edge' = f(e, SL, SR, fl, fr); the tables of block i + 1 go
    // to the other parity buffers) -- so the two run side by side: the node kernel (50 ... 99 long-running CTAs) on the
    // high-priority side stream, this kernel's 1 228 short CTAs on the SMs it leaves free.
    auto launch_edge_d = [&](cudaStream_t s_) {
      if (E > 0 && tc_nb && net->tc_block_off[i][MDB_T_EB_SELF] >= 0) {
        TcEdgeDArgs da;
        memset(&da, 0, sizeof(da));
        da.tc_blob = reinterpret_cast<const uint8_t*>(net->tc_blob); da.tb = tbi;
        for (int s = 0; s < MDB_NUM_TC_SLOTS; ++s) da.tco.o[s] = net->tc_block_off[i][s];
        da.left = plan->left; da.right = plan->right; da.n_nodes = N; da.n_edges = E; da.update_pos = net->update_pos;
        da.ebuf = ea.ebuf; da.sl = ea.sl; da.fl = ea.fl; da.fr = ea.fr; da.pos_cur = pos_cur; da.pos_nxt = pos_nxt;
        fill_edge_d_vecs(da.v, net->blob_host, ea.off, net->update_pos != 0);
        if (xf) LAUNCH(MDB_K_tc_edge_d, s_,
                       (tc_edge_d_kernel<true><<<(E + tc::ROWS - 1) / tc::ROWS, TC_NB_THREADS, SMEM_TC_EDGE_D, s_>>>(da)));
        else LAUNCH(MDB_K_tc_edge_d, s_,
                    (tc_edge_d_kernel<false><<<(E + tc::ROWS - 1) / tc::ROWS, TC_NB_THREADS, SMEM_TC_EDGE_D, s_>>>(da)));
      } else if (E > 0) {
        LAUNCH(MDB_K_edge_d, s_, (edge_kernel_d<<<edge_tiles, NTHREADS, SMEM_EDGE_D, s_>>>(ea)));
      }
    };
    const bool forked = overlap && side != nullptr && E > 0 && tc_ffn && !net->update_pos;
    const bool forked_pre = overlap && side != nullptr && E > 0 && tc_ffn && net->update_pos;
    if (E > 0 && tc_nb) {
      TcNbArgs ta;
      memset(&ta, 0, sizeof(ta));
      ta.blob = net->blob; ta.tc_blob = reinterpret_cast<const uint8_t*>(net->tc_blob); ta.off = ea.off; ta.tb = tbi;
      for (int s = 0; s < MDB_NUM_TC_SLOTS; ++s) ta.tco.o[s] = net->tc_block_off[i][s];
      ta.left = plan->left; ta.right = plan->right; ta.n_nodes = N; ta.n_edges = E; ta.ebuf = ea.ebuf;
      ta.dbg = g_dbg_sel == 0 ? g_dbg_stamps : nullptr;
      fill_nb_vecs(ta.v, net->blob_host, ea.off);
      static const bool nb16 = []() { const char* e = getenv("MDB_TC_NB16"); return e == nullptr || e[0] != '0'; }();
      if (nb16) {
        const int grid = persistent_grid((E + tc::ROWS - 1) / tc::ROWS);
        if (xf) LAUNCH(MDB_K_tc_nodeblock, st, (tc_nodeblock_fwd16_kernel<true><<<grid, NB16_THREADS, SMEM_TC_NB16, st>>>(ta)));
        else LAUNCH(MDB_K_tc_nodeblock, st, (tc_nodeblock_fwd16_kernel<false><<<grid, NB16_THREADS, SMEM_TC_NB16, st>>>(ta)));
      } else {
        LAUNCH(MDB_K_tc_nodeblock, st,
               (tc_nodeblock_fwd_kernel<<<(E + tc::ROWS - 1) / tc::ROWS, TC_NB_THREADS, SMEM_TC_NB, st>>>(ta)));
      }
    }
    fill_blk(na.mid, net, i);
    na.tb = tbi; na.tb_pre = tb_of(i + 1 < L ? i + 1 : i);
    na.do_mid = 1; na.do_pre = (i + 1 < L); na.do_dec = (i + 1 == L) && net->kind != 0;
    if (na.do_pre) fill_blk(na.pre, net, i + 1);
    na.sl_next = sl_of(i + 1 < L ? i + 1 : i); na.fl_next = fl_of(i + 1 < L ? i + 1 : i); na.fr_next = fr_of(i + 1 < L ? i + 1 : i);
    na.x_save = (in.save && i + 1 < L) ? sv.x + (size_t)(i + 1) * ND : nullptr;
    na.agg_save = in.save ? sv.agg + (size_t)i * ND : nullptr;
    na.pos_cur = pos_cur; na.pos_nxt = pos_nxt;
    if (forked) {                                    // node kernel on the high-priority side stream, tail kernel on the main one
      CUDA_TRY(cudaEventRecord(side->fork, st));
      CUDA_TRY(cudaStreamWaitEvent(side->s, side->fork, 0));
      rc = launch_node(net, na, i, na.do_pre ? i + 1 : -1, node_tiles, side->s);
      if (rc) return rc;
      launch_edge_d(st);
      CUDA_TRY(cudaEventRecord(side->join, side->s));
      CUDA_TRY(cudaStreamWaitEvent(st, side->join, 0));       // join: the next block needs both
    } else if (forked_pre && na.do_pre) {
      // With a position update the tail / PosUpdate kernel needs this block's `mid` half (new h_node -> lf / rf, pos_nxt) but
      // not the hoisted tables of block i + 1: `mid` stays in the chain, `pre` runs beside the edge kernel.
      NodeArgs nm = na;
      nm.do_pre = 0;
      rc = launch_node(net, nm, i, -1, node_tiles, st);
      if (rc) return rc;
      CUDA_TRY(cudaEventRecord(side->fork, st));
      CUDA_TRY(cudaStreamWaitEvent(side->s, side->fork, 0));
      NodeArgs np = na;
      np.do_mid = 0; np.do_dec = 0;
      rc = launch_node(net, np, -1, i + 1, node_tiles, side->s);
      if (rc) return rc;
      launch_edge_d(st);
      CUDA_TRY(cudaEventRecord(side->join, side->s));
      CUDA_TRY(cudaStreamWaitEvent(st, side->join, 0));
    } else {
      rc = launch_node(net, na, i, na.do_pre ? i + 1 : -1, node_tiles, st);
      if (rc) return rc;
      launch_edge_d(st);
    }
    if (net->update_pos) { const float* t_ = pos_cur; pos_cur = pos_nxt; pos_nxt = const_cast<float*>(t_); }
  }

  if (net->kind == 0) {
    CUDA_TRY(cudaMemcpyAsync(in.out_node, tb.x, (size_t)N * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (E > 0) LAUNCH(MDB_K_edge_unsort, st, (edge_unsort_kernel<<<(E + 3) / 4, 256, 0, st>>>(E, plan->perm, tb.hedge, in.out_edge)));
  } else if (plan->n_half > 0) {
    DecArgs da;
    memset(&da, 0, sizeof(da));
    da.blob = net->blob; da.head = head; da.tb = tb; da.left = plan->left; da.right = plan->right;
    da.inv = plan->inv; da.n_half = plan->n_half; da.kind = net->kind; da.ke = net->num_edge_types;
    da.out = in.out_edge;
    LAUNCH(MDB_K_edge_decode, st, (edge_decode_kernel<<<(plan->n_half + TM - 1) / TM, NTHREADS, SMEM_DEC, st>>>(da)));
  }
  if (in.out_pos)
    CUDA_TRY(cudaMemcpyAsync(in.out_pos, pos_cur, (size_t)N * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaGetLastError());
  return MDB_OK;
}

#include "mdb_backward.cuh"
#include "tc_selftest.cuh"
#include "mdb_transition.cuh"
#include "mdb_graph_build.cuh"
#include "mdb_decode.cuh"

// max |v| over a buffer -> atomicMax on the bit pattern (non-negative floats order like unsigned ints); NaN / Inf map to Inf
__global__ void amax_kernel(const float* __restrict__ v, size_t n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float x = fabsf(v[i]);
    m = (x > m || x != x) ? (x != x ? INFINITY : x) : m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

}  // namespace

extern "C" {

size_t mdb_workspace_bytes(int64_t n_nodes, int64_t n_edges, int32_t with_backward, int32_t num_blocks) {
  Tables tb;
  size_t n = carve(tb, nullptr, n_nodes, n_edges);
  if (with_backward) {
    Saved sv;
    n += carve_saved(sv, nullptr, n_nodes, n_edges, num_blocks);
  }
  return n * sizeof(float);
}

int mdb_net_forward(const mdb_net_desc* net, const mdb_plan* plan, const float* h_node_in, const float* pos_in,
                    const float* h_edge_in, const float* node_time, const float* edge_time, float* h_node_out,
                    float* pos_out, float* h_edge_out, float* workspace, size_t workspace_bytes, void* stream) {
  if (!net || net->kind != 0) return fail(MDB_EINVAL, "mdb_net_forward needs a kind-0 descriptor%s");
  FwdIn in;
  memset(&in, 0, sizeof(in));
  in.h_node = h_node_in; in.pos = pos_in; in.h_edge = h_edge_in; in.node_time = node_time; in.edge_time = edge_time;
  in.out_node = h_node_out; in.out_pos = pos_out; in.out_edge = h_edge_out;
  return run_forward(net, plan, in, workspace, workspace_bytes, (cudaStream_t)stream);
}

int mdb_moldiff_forward(const mdb_net_desc* net, const mdb_plan* plan, const float* h_node_pert,
                        const float* pos_pert, const float* h_edge_pert, const int64_t* batch_node,
                        const int64_t* batch_edge, const int64_t* t, float* pred_node, float* pred_pos,
                        float* pred_halfedge, float* workspace, size_t workspace_bytes, void* stream) {
  if (!net || net->kind != 1) return fail(MDB_EINVAL, "mdb_moldiff_forward needs a kind-1 descriptor%s");
  FwdIn in;
  memset(&in, 0, sizeof(in));
  in.h_node = h_node_pert; in.pos = pos_pert; in.h_edge = h_edge_pert;
  in.batch_node = batch_node; in.batch_edge = batch_edge; in.t = t;
  in.out_node = pred_node; in.out_pos = pred_pos; in.out_edge = pred_halfedge;
  return run_forward(net, plan, in, workspace, workspace_bytes, (cudaStream_t)stream);
}

int mdb_transition_step(int32_t n_nodes, int32_t n_half, int32_t kn, int32_t ke, const int64_t* batch_node,
                        const int64_t* batch_half, const int64_t* t, const float* pos, const float* pred_pos,
                        const float* z_pos, const float* coef_x0, const float* coef_xt, const float* std_, float* pos_out,
                        const float* pred_node, const float* log_node, const float* u_node, const float* qn_cum,
                        const float* qn_stepT, float* log_node_out, float* h_node_out, const float* pred_half,
                        const float* log_half, const float* u_half, const float* qe_cum, const float* qe_stepT,
                        float* log_half_out, float* h_edge_out, int64_t* half_type_out, void* stream) {
  if (n_nodes < 0 || n_half < 0 || kn < 1 || ke < 1 || kn > 16 || ke > 16) return fail(MDB_EINVAL, "mdb_transition_step: bad sizes%s");
  if (n_nodes + n_half == 0) return MDB_OK;
  TransArgs a;
  memset(&a, 0, sizeof(a));
  a.n_nodes = n_nodes; a.n_half = n_half; a.kn = kn; a.ke = ke;
  a.batch_node = batch_node; a.batch_half = batch_half; a.t = t;
  a.pos = pos; a.pred_pos = pred_pos; a.z_pos = z_pos; a.coef_x0 = coef_x0; a.coef_xt = coef_xt; a.stdv = std_; a.pos_out = pos_out;
  a.pred_node = pred_node; a.log_node = log_node; a.u_node = u_node; a.qn_cum = qn_cum; a.qn_stepT = qn_stepT;
  a.log_node_out = log_node_out; a.h_node_out = h_node_out;
  a.pred_half = pred_half; a.log_half = log_half; a.u_half = u_half; a.qe_cum = qe_cum; a.qe_stepT = qe_stepT;
  a.log_half_out = log_half_out; a.h_edge_out = h_edge_out; a.half_type_out = half_type_out;
  cudaStream_t st = (cudaStream_t)stream;
  const int total = n_nodes + n_half;
  if (kn == 8 && ke == 6) {          // the reference's type counts (utils/transforms.py:22-23)
    LAUNCH(MDB_K_transition, st, (transition_step_kernel<8, 6><<<(total + 255) / 256, 256, 0, st>>>(a)));
  } else {
    LAUNCH(MDB_K_transition, st, (transition_step_kernel<0, 0><<<(total + 255) / 256, 256, 0, st>>>(a)));
  }
  return MDB_OK;
}

int mdb_decode_rows(int32_t n_nodes, int32_t kn, const float* pred_node, int32_t n_half, int32_t ke, const float* pred_half,
                    uint8_t* node_type, float* node_prob, uint8_t* half_type, float* half_prob, void* stream) {
  if (n_nodes < 0 || n_half < 0 || kn < 1 || ke < 1 || kn > 255 || ke > 255) return fail(MDB_EINVAL, "mdb_decode_rows: bad sizes%s");
  if (n_nodes + n_half == 0) return MDB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  LAUNCH(MDB_K_decode_rows, st,
         (decode_rows_kernel<<<(n_nodes + n_half + 255) / 256, 256, 0, st>>>(n_nodes, kn, pred_node, n_half, ke, pred_half,
                                                                             node_type, node_prob, half_type, half_prob)));
  return MDB_OK;
}

int mdb_radius_graph(int32_t n_nodes, const float* pos, const int32_t* seg_lo, const int32_t* seg_hi, float radius,
                     int32_t loop, int32_t max_num_neighbors, int32_t* counts, int32_t* neighbors, void* stream) {
  if (n_nodes < 0 || max_num_neighbors < 1 || !(radius > 0.f)) return fail(MDB_EINVAL, "mdb_radius_graph: bad arguments%s");
  if (n_nodes == 0) return MDB_OK;
  if (!pos || !seg_lo || !seg_hi || !counts || !neighbors) return fail(MDB_EINVAL, "null argument%s");
  cudaStream_t st = (cudaStream_t)stream;
  LAUNCH(MDB_K_graph_build, st,
         (radius_graph_kernel<<<(n_nodes + 127) / 128, 128, 0, st>>>(n_nodes, pos, seg_lo, seg_hi, radius * radius, loop,
                                                                    max_num_neighbors, counts, neighbors)));
  return MDB_OK;
}

int mdb_operand_amax(int64_t n_nodes, int64_t n_edges, const float* workspace, float* amax4, void* stream) {
  if (!workspace || !amax4 || n_nodes <= 0 || n_edges < 0) return fail(MDB_EINVAL, "mdb_operand_amax: bad arguments%s");
  Tables tb;
  carve(tb, const_cast<float*>(workspace), n_nodes, n_edges);
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(amax4, 0, 4 * sizeof(float), st));
  const float* src[4] = {tb.x, tb.hedge, tb.ebuf, tb.hn};
  const size_t cnt[4] = {(size_t)n_nodes * D, (size_t)n_edges * C, (size_t)n_edges * C, (size_t)n_nodes * D};
  for (int i = 0; i < 4; ++i) {
    if (cnt[i] == 0) continue;
    amax_kernel<<<(int)std::min<size_t>((cnt[i] + 255) / 256, 592), 256, 0, st>>>(src[i], cnt[i], reinterpret_cast<unsigned*>(amax4) + i);
    ++g_launches;
  }
  CUDA_TRY(cudaGetLastError());
  return MDB_OK;
}

int mdb_knn_graph(int32_t n_nodes, const float* pos, const int32_t* seg_lo, const int32_t* seg_hi, int32_t k, int32_t loop,
                  int32_t* counts, int32_t* neighbors, void* stream) {
  if (n_nodes < 0 || k < 1 || k > KNN_MAX_K) return fail(MDB_EINVAL, "mdb_knn_graph: k must be in 1..32%s");
  if (n_nodes == 0) return MDB_OK;
  if (!pos || !seg_lo || !seg_hi || !counts || !neighbors) return fail(MDB_EINVAL, "null argument%s");
  cudaStream_t st = (cudaStream_t)stream;
  LAUNCH(MDB_K_graph_build, st,
         (knn_graph_kernel<<<(n_nodes + 127) / 128, 128, 0, st>>>(n_nodes, pos, seg_lo, seg_hi, k, loop, counts, neighbors)));
  return MDB_OK;
}

int mdb_bondpred_forward(const mdb_net_desc* net, const mdb_plan* plan, const float* h_node, const float* pos,
                         const int64_t* batch_node, const int64_t* batch_edge, const int64_t* t, float* logits,
                         int32_t save_for_backward, float* workspace, size_t workspace_bytes, void* stream) {
  if (!net || net->kind != 2) return fail(MDB_EINVAL, "mdb_bondpred_forward needs a kind-2 descriptor%s");
  FwdIn in;
  memset(&in, 0, sizeof(in));
  in.h_node = h_node; in.pos = pos; in.h_edge = nullptr;
  in.batch_node = batch_node; in.batch_edge = batch_edge; in.t = t;
  in.out_edge = logits;
  in.save = save_for_backward;
  return run_forward(net, plan, in, workspace, workspace_bytes, (cudaStream_t)stream);
}

int mdb_bondpred_backward(const mdb_net_desc* net, const mdb_plan* plan, const float* h_node, const float* pos,
                          const int64_t* batch_node, const int64_t* batch_edge, const int64_t* t,
                          const float* d_logits, float* d_pos, float* workspace, size_t workspace_bytes, void* stream) {
  (void)h_node; (void)batch_node; (void)batch_edge; (void)t;   // everything they determine was saved by the forward
  return run_bondpred_backward(net, plan, pos, d_logits, d_pos, workspace, workspace_bytes, (cudaStream_t)stream);
}

int mdb_tc_selftest(const float* x, const void* w_img, float* y, int32_t k, int32_t n, int32_t twice, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (twice & 2) {   // bit 1: cross-first accumulation order
    twice &= 1;
    if (k == 64 && n == 256) return launch_tc_selftest<64, 256, true>(x, w_img, y, twice, st);
    if (k == 256 && n == 256) return launch_tc_selftest<256, 256, true>(x, w_img, y, twice, st);
    if (k == 256 && n == 64) return launch_tc_selftest<256, 64, true>(x, w_img, y, twice, st);
    if (k == 128 && n == 128) return launch_tc_selftest<128, 128, true>(x, w_img, y, twice, st);
    if (k == 80 && n == 64) return launch_tc_selftest<80, 64, true>(x, w_img, y, twice, st);
    if (k == 64 && n == 32) return launch_tc_selftest<64, 32, true>(x, w_img, y, twice, st);
    return fail(MDB_EINVAL, "mdb_tc_selftest: unsupported (k, n) for the cross-first order%s");
  }
  if (k == 64 && n == 256) return launch_tc_selftest<64, 256>(x, w_img, y, twice, st);
  if (k == 256 && n == 256) return launch_tc_selftest<256, 256>(x, w_img, y, twice, st);
  if (k == 256 && n == 64) return launch_tc_selftest<256, 64>(x, w_img, y, twice, st);
  if (k == 128 && n == 128) return launch_tc_selftest<128, 128>(x, w_img, y, twice, st);
  if (k == 64 && n == 32) return launch_tc_selftest<64, 32>(x, w_img, y, twice, st);
  if (k == 80 && n == 64) return launch_tc_selftest<80, 64>(x, w_img, y, twice, st);
  if (k == 64 && n == 128) return launch_tc_selftest<64, 128>(x, w_img, y, twice, st);
  if (k == 32 && n == 64) return launch_tc_selftest<32, 64>(x, w_img, y, twice, st);
  if (k == 128 && n == 64) return launch_tc_selftest<128, 64>(x, w_img, y, twice, st);
  return fail(MDB_EINVAL, "mdb_tc_selftest: unsupported (k, n)%s");
}

void mdb_debug_set_buffer(void* device_i64_buffer) { g_dbg_stamps = reinterpret_cast<long long*>(device_i64_buffer); }
void mdb_debug_select(int32_t kernel) { g_dbg_sel = kernel; }

void mdb_profile_begin(void) {
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_profiling = true;
}

int mdb_profile_end(double* ms_per_class, int64_t* launches_per_class) {
  g_profiling = false;
  for (int i = 0; i < MDB_NUM_KERNEL_CLASSES; ++i) { ms_per_class[i] = 0.0; launches_per_class[i] = 0; }
  for (auto& r : g_prof) {
    CUDA_TRY(cudaEventSynchronize(r.b));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_per_class[r.cls] += ms;
    launches_per_class[r.cls] += 1;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return MDB_OK;
}

const char* mdb_kernel_class_name(int cls) {
  static const char* names[] = {
#define MDB_X(name) #name,
      MDB_KERNEL_CLASSES(MDB_X)
#undef MDB_X
  };
  return (cls >= 0 && cls < MDB_NUM_KERNEL_CLASSES) ? names[cls] : "?";
}
int mdb_num_kernel_classes(void) { return MDB_NUM_KERNEL_CLASSES; }

const char* mdb_last_error(void) { return g_err; }
int mdb_version(void) { return 1; }
int64_t mdb_launch_count(void) { return g_launches; }

}  // extern "C"
