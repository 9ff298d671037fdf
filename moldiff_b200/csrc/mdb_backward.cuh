// Input-gradient backward of the bond predictor (BondPredictor.forward -> d sum(logits * d_logits) / d pos),
// i.e. what torch.autograd.grad(objective, pos_in) computes inside the guidance block of MolDiff.sample
// (reference models/model.py:312-325) -- written by hand so that nothing per-edge is saved by the forward
// except e_i = edge_embs[i](.) [E,64]; every other per-edge activation is recomputed tile by tile.
//
// pos enters the network only through the RBF features g (graph.py:351-357, update_pos = False), which feed
// edge_embs[i] of every block, so  d/dpos = sum_i (d/d e_i) W_EE,i[:, 64:80]  chained through exp/clamp/norm.
// That needs d/d e_i for every block, hence the full input-gradient chain through both residual streams.
//
// Included by mdb_forward.cu inside its anonymous namespace (shares Tables / Saved / BlkOff / LAUNCH).
//
// Kernel schedule (blocks in reverse):
//   bwd_decode            edge_decoder backward -> dh (grad of final h_edge), ddect
//   bwd_node(final)       dx = ddect W_dec1n^T ; then phase A(L-1)
//   for i = L-1 .. 0:
//     [phase A(i) inside the node kernel]  recompute block i's per-node tables from x_i; NodeBlock node-tail
//                                          backward -> dagg ; dx += dc W_cen^T ; clear scatter accumulators
//     bwd_edge_tail(i)      EdgeBlock tail backward: du -> DUL/DUR, de = dh' + du W_self^T
//     bwd_edge_nodeblock(i) NodeBlock edge path backward: de += ..., dgx[r], dhn[r]
//     bwd_edge_bondffn(i)   BondFFN L/R backward: de += ..., dnl, dgn ; then dh = de W_EEh^T, dG += de W_EEg^T
//     bwd_node: phase B(i)  dx += all per-node table gradients ; phase A(i-1)
//   bwd_pos               dG -> d dist -> d pos
#pragma once

// ------------------------------------------------------------------------------------------------
// edge decoder backward (half-edge tiles, caller order)
// ------------------------------------------------------------------------------------------------
// The backward is linear in d_logits, so the whole chain runs on d_logits * 2^k with k chosen per call such that
// max |d_logits * 2^k| lies in [8, 16): every gradient tile then sits inside the fp16 operand range of the tensor-core
// GEMMs whatever the caller's objective scale is, and bwd_pos multiplies the result by 2^-k.  Both factors are exact.
__global__ void grad_amax_kernel(const float* __restrict__ d, size_t n, unsigned* __restrict__ amax_bits) {
  float m = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(d[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && m < INFINITY) atomicMax(amax_bits, __float_as_uint(m));
}
__device__ __forceinline__ int grad_shift(const float* gamax) {      // k: scale 2^k
  const float m = *gamax;
  return m > 0.f ? 3 - ilogbf(m) : 0;
}

struct BwdDecArgs {
  const float* blob;
  HeadOff head;
  Tables tb;
  const int *left, *right, *inv;
  int n_half, ke;
  const float* d_logits;   // [Eh][ke]
  const float* gamax;      // max |d_logits| (grad_amax_kernel)
  float* dh;               // [E][64] sorted order
  float* ddect;            // [N][64] (pre-zeroed)
};

__global__ void __launch_bounds__(NTHREADS, 1) bwd_decode_kernel(const BwdDecArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Es = smem;               // [64][64] hs
  float* A = Es + TM * C;         // [64][64]
  float* DL = A + TM * C;         // [64][32] d_logits, zero padded
  float* Ws = DL + TM * 32;
  int* ls = reinterpret_cast<int*>(Ws + 2 * WCHUNK);
  int* rs = ls + TM;
  int* q1 = rs + TM;
  int* q2 = q1 + TM;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p0 = blockIdx.x * TM;
  const float* blob = a.blob;
  const HeadOff& hoff = a.head;
  const Tables& tb = a.tb;
  if (tid < TM) {
    const int p = p0 + tid;
    int l = -1, r = -1, qa = -1, qb = -1;
    if (p < a.n_half) { qa = a.inv[p]; qb = a.inv[p + a.n_half]; l = a.left[qa]; r = a.right[qa]; }
    ls[tid] = l; rs[tid] = r; q1[tid] = qa; q2[tid] = qb;
  }
  const float gscale = scalbnf(1.f, grad_shift(a.gamax));
  for (int i = tid; i < TM * 32; i += NTHREADS) {
    const int r = i >> 5, c = i & 31, p = p0 + r;
    DL[i] = (p < a.n_half && c < a.ke) ? a.d_logits[(size_t)p * a.ke + c] * gscale : 0.f;
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
  // forward recompute, keeping xhat / rstd of both LayerNorms
  float xh1[8][2], rstd1[8], xh2[8][2], rstd2[8];
  tile_gemm<C, C>(xh1, Es, C, H_(EDEC1_W), Ws);
  add_rowvec<C>(xh1, H_(EDEC1_B), lane);
  gather_rows<C, false>(xh1, tb.dect, ls + warp * 8, lane);
  gather_rows<C, false>(xh1, tb.dect, rs + warp * 8, lane);
  ln_xhat<C>(xh1, rstd1);
  {
    float r1[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { r1[i][0] = xh1[i][0]; r1[i][1] = xh1[i][1]; }
    affine_relu<C>(r1, H_(EDEC1_G), H_(EDEC1_BE), lane);
    store_smem<C>(r1, A, C, warp, lane);
  }
  tile_gemm<C, C>(xh2, A, C, H_(EDEC2_W), Ws);
  add_rowvec<C>(xh2, H_(EDEC2_B), lane);
  ln_xhat<C>(xh2, rstd2);
  // backward
  float d[8][2];
  tile_gemm<32, C>(d, DL, 32, H_(T_EDEC3), Ws);                       // d relu2 = d_logits W3
  ln_relu_bwd<C>(d, xh2, rstd2, H_(EDEC3_G), H_(EDEC3_BE), lane);      // -> d a2
  store_smem<C>(d, A, C, warp, lane);
  tile_gemm<C, C>(d, A, C, H_(T_EDEC2), Ws);                          // d relu1
  ln_relu_bwd<C>(d, xh1, rstd1, H_(EDEC1_G), H_(EDEC1_BE), lane);      // -> d a1
  scatter_add_rows<C, false>(d, a.ddect, ls + warp * 8, lane);
  scatter_add_rows<C, false>(d, a.ddect, rs + warp * 8, lane);
  store_smem<C>(d, A, C, warp, lane);
  tile_gemm<C, C>(d, A, C, H_(T_EDEC1), Ws);                          // d hs -> both directed edges of the pair
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = warp * 8 + i;
    if (q1[r] >= 0) {
      store_cols<C>(d[i], a.dh + (size_t)q1[r] * C, lane);
      store_cols<C>(d[i], a.dh + (size_t)q2[r] * C, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// node kernel of the backward pass
// ------------------------------------------------------------------------------------------------
struct BwdNodeArgs {
  const float* blob;
  BlkOff blkB, blkA;     // block i (phase B) and block i-1 (phase A)
  HeadOff head;
  Tables tb;
  Saved sv;
  int n_nodes;
  int do_final, do_B, do_A;
  const float* xB;       // saved h_node entering block i      (phase B: node_net backward)
  const float* xA;       // saved h_node entering block i-1    (phase A: table recompute)
  const float* aggA;     // saved aggregated messages of block i-1
  float *flA, *frA;      // table destinations for phase A
  int red_blocked;       // dgx / dhn are in the node-blocked layout (written by tc_nodeblock_bwd16_kernel)
};

// load a [64][W] tile of a per-node table into smem (zero padded rows)
template <int W>
__device__ __forceinline__ void load_node_tile(float* dst, const float* __restrict__ src, int row0, int n_nodes) {
  for (int i = threadIdx.x; i < TM * W / 4; i += NTHREADS) {
    const int r = i / (W / 4), c4 = i % (W / 4);
    const int n = row0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < n_nodes) v = reinterpret_cast<const float4*>(src + (size_t)n * W)[c4];
    reinterpret_cast<float4*>(dst + r * W)[c4] = v;
  }
}

// same for a node-blocked [.][256] table (tile_engine.cuh: blk_off); rows beyond n_nodes hold zeros there
__device__ __forceinline__ void load_node_tile_blocked(float* dst, const float* __restrict__ src, int row0) {
  for (int i = threadIdx.x; i < TM * D / 4; i += NTHREADS) {
    const int r = i & (TM - 1), c4 = i / TM;          // consecutive threads -> consecutive nodes: contiguous 16-byte pieces
    reinterpret_cast<float4*>(dst + r * D)[c4] = *reinterpret_cast<const float4*>(src + blk_off(row0 + r, c4));
  }
}
__device__ __forceinline__ void zero_node_rows_blocked(float* __restrict__ dst, int row0) {   // TM = 64 nodes = 2 whole blocks
  for (int i = threadIdx.x; i < TM * D / 4; i += NTHREADS)
    reinterpret_cast<float4*>(dst + (size_t)row0 * D)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int W>
__device__ __forceinline__ void zero_node_rows(float* __restrict__ dst, int row0, int n_nodes) {
  for (int i = threadIdx.x; i < TM * W / 4; i += NTHREADS) {
    const int n = row0 + i / (W / 4);
    if (n < n_nodes) reinterpret_cast<float4*>(dst + (size_t)row0 * W)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) bwd_node_kernel(const BwdNodeArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* X = smem;                 // [64][256]
  float* A = X + TM * D;           // [64][256]
  float* Ws = A + TM * D;
  float* tns = Ws + 2 * WCHUNK;    // [64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * TM;
  const float* blob = a.blob;
  const Tables& tb = a.tb;
  const Saved& sv = a.sv;
  if (tid < TM) tns[tid] = (row0 + tid < a.n_nodes) ? tb.tn[row0 + tid] : 0.f;

  float dx[8][8];   // running gradient w.r.t. h_node rows of this tile (lives in registers across phases)
  if (a.do_final) {   // d h_node(final) = ddect W_dec1n^T                     bond_predictor.py:155-160
    const HeadOff& hoff = a.head;
    load_node_tile<C>(A, sv.ddect, row0, a.n_nodes);
    __syncthreads();
    tile_gemm<C, D>(dx, A, C, H_(T_EDEC1N), Ws);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = row0 + warp * 8 + i;
      if (n < a.n_nodes) load_cols<D>(dx[i], sv.dx + (size_t)n * D, lane);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) dx[i][j] = 0.f;
      }
    }
  }

  if (a.do_B) {
    const BlkOff& off = a.blkB;
    // gradients of the hoisted per-node projections, each  dx += dTable * W^T
    __syncthreads();
    load_node_tile<C>(A, sv.dul, row0, a.n_nodes);  __syncthreads();
    tile_gemm<C, D, true>(dx, A, C, W_(T_EB_NFL), Ws);                // node_ffn_left
    load_node_tile<C>(A, sv.dur, row0, a.n_nodes);  __syncthreads();
    tile_gemm<C, D, true>(dx, A, C, W_(T_EB_NFR), Ws);                // node_ffn_right
    load_node_tile<128>(A, sv.dnl, row0, a.n_nodes);  __syncthreads();
    tile_gemm<128, D, true>(dx, A, 128, W_(T_EL_NL), Ws);             // bond_ffn_left.node_linear
    load_node_tile<128>(A, sv.dnl + (size_t)a.n_nodes * 128, row0, a.n_nodes);  __syncthreads();
    tile_gemm<128, D, true>(dx, A, 128, W_(T_ER_NL), Ws);
    load_node_tile<32>(A, sv.dgn, row0, a.n_nodes);  __syncthreads();
    tile_gemm<32, D, true>(dx, A, 32, W_(T_EL_GN), Ws);               // bond_ffn gate.net.0 node columns
    load_node_tile<32>(A, sv.dgn + (size_t)a.n_nodes * 32, row0, a.n_nodes);  __syncthreads();
    tile_gemm<32, D, true>(dx, A, 32, W_(T_ER_GN), Ws);
    if (a.red_blocked) load_node_tile_blocked(A, sv.dgx, row0); else load_node_tile<D>(A, sv.dgx, row0, a.n_nodes);
    __syncthreads();
    tile_gemm<D, D, true>(dx, A, D, W_(T_NB_GX), Ws);                 // NodeBlock gate.net.0 node columns
    // node_net backward: hn = W2 relu(LN(W1 x + b1)) + b2                        graph.py:39
    load_node_tile<D>(X, a.xB, row0, a.n_nodes);
    if (a.red_blocked) load_node_tile_blocked(A, sv.dhn, row0); else load_node_tile<D>(A, sv.dhn, row0, a.n_nodes);
    __syncthreads();
    {
      float dr[8][8];
      tile_gemm<D, D>(dr, A, D, W_(T_NB_NN2), Ws);                    // d relu1
      store_smem<D>(dr, A, D, warp, lane);
      float xh[8][8], rstd[8];
      tile_gemm<D, D>(xh, X, D, W_(NB_NN1_W), Ws);
      add_rowvec<D>(xh, W_(NB_NN1_B), lane);
      ln_xhat<D>(xh, rstd);
      float g[8], b[8];
      load_cols<D>(g, W_(NB_NN1_G), lane);
      load_cols<D>(b, W_(NB_NN1_BE), lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float d[8];
        load_cols<D>(d, A + (warp * 8 + i) * D, lane);
        ln_relu_bwd_row<D>(d, xh[i], rstd[i], g, b);
        store_cols<D>(d, A + (warp * 8 + i) * D, lane);
      }
    }
    tile_gemm<D, D, true>(dx, A, D, W_(T_NB_NN1), Ws);
  }

  if (a.do_A) {
    const BlkOff& off = a.blkA;
    __syncthreads();
    // (the per-node tables of every block were saved by the forward -- nothing to recompute here)
    // NodeBlock node tail backward: dn = W_out relu(LN(cen + agg)) + b ; h_node' = h_node + dn   graph.py:51-54,363
    store_smem<D>(dx, A, D, warp, lane);            // d(dn) = dx'  (dx itself is parked in A until reloaded below)
    {
      float dr[8][8];
      tile_gemm<D, D>(dr, A, D, W_(T_NB_OUT), Ws);
      float xh[8][8], rstd[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = row0 + warp * 8 + i;
        if (n < a.n_nodes) {
          float u[8], v[8];
          load_cols<D>(u, tb.cen + (size_t)n * D, lane);   // written by this thread just above (same rows / cols)
          load_cols<D>(v, a.aggA + (size_t)n * D, lane);
#pragma unroll
          for (int j = 0; j < 8; ++j) xh[i][j] = u[j] + v[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) xh[i][j] = 0.f;
        }
      }
      ln_xhat<D>(xh, rstd);
      ln_relu_bwd<D>(dr, xh, rstd, W_(NB_LN_G), W_(NB_LN_BE), lane);   // -> dc = d(cen + agg)
      store_table<D>(dr, sv.dagg, row0, a.n_nodes, warp, lane);
      store_smem<D>(dr, X, D, warp, lane);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) load_cols<D>(dx[i], A + (warp * 8 + i) * D, lane);
    tile_gemm<D, D, true>(dx, X, D, W_(T_NB_CEN), Ws);                  // dx += dc W_cen^T
    // clear the scatter accumulators the edge kernels of block i-1 add into
    zero_node_rows<C>(sv.dul, row0, a.n_nodes);
    zero_node_rows<C>(sv.dur, row0, a.n_nodes);
    if (a.red_blocked) { zero_node_rows_blocked(sv.dgx, row0); zero_node_rows_blocked(sv.dhn, row0); }
    else { zero_node_rows<D>(sv.dgx, row0, a.n_nodes); zero_node_rows<D>(sv.dhn, row0, a.n_nodes); }
    zero_node_rows<128>(sv.dnl, row0, a.n_nodes);
    zero_node_rows<128>(sv.dnl + (size_t)a.n_nodes * 128, row0, a.n_nodes);
    zero_node_rows<32>(sv.dgn, row0, a.n_nodes);
    zero_node_rows<32>(sv.dgn + (size_t)a.n_nodes * 32, row0, a.n_nodes);
  }
  store_table<D>(dx, sv.dx, row0, a.n_nodes, warp, lane);
}

// ------------------------------------------------------------------------------------------------
// edge kernels of the backward pass
// ------------------------------------------------------------------------------------------------
struct BwdEdgeArgs {
  const float* blob;
  BlkOff off;
  Tables tb;
  Saved sv;
  const int *left, *right;
  int n_nodes, n_edges;
  const float* e;          // [E][64] saved e_i
  const float* sl;         // [2][N][64] saved SL / SR of block i
  const float *fl, *fr;    // recomputed node_ffn tables of block i
};

__device__ __forceinline__ void bwd_edge_meta(const BwdEdgeArgs& a, int q0, int* ls, int* rs, float* tes) {
  const int tid = threadIdx.x;
  if (tid < TM) {
    const int q = q0 + tid;
    const bool ok = q < a.n_edges;
    ls[tid] = ok ? a.left[q] : -1;
    rs[tid] = ok ? a.right[q] : -1;
    tes[tid] = ok ? a.tb.te[q] : 0.f;
  }
}

template <int W>
__device__ __forceinline__ void load_edge_tile(float* dst, const float* __restrict__ src, int q0, int n_edges) {
  for (int i = threadIdx.x; i < TM * W / 4; i += NTHREADS) {
    const int r = i / (W / 4), c4 = i % (W / 4);
    const int q = q0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < n_edges) v = reinterpret_cast<const float4*>(src + (size_t)q * W)[c4];
    reinterpret_cast<float4*>(dst + r * W)[c4] = v;
  }
}

// EdgeBlock tail: h' = e + W_out relu(LN(u)) + b,  u = SL[l] + SR[r] + fl[l] + fr[r] + W_self e + b    graph.py:286-294
__global__ void __launch_bounds__(NTHREADS, 1) bwd_edge_tail_kernel(const BwdEdgeArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Es = smem;               // [64][64] e
  float* DH = Es + TM * C;        // [64][64] dh'
  float* A = DH + TM * C;         // [64][64]
  float* Ws = A + TM * C;
  float* tes = Ws + 2 * WCHUNK;
  int* ls = reinterpret_cast<int*>(tes + TM);
  int* rs = ls + TM;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * TM;
  const float* blob = a.blob;
  const BlkOff& off = a.off;
  bwd_edge_meta(a, q0, ls, rs, tes);
  load_edge_tile<C>(Es, a.e, q0, a.n_edges);
  load_edge_tile<C>(DH, a.sv.dh, q0, a.n_edges);
  __syncthreads();
  const int* lw = ls + warp * 8;
  const int* rw = rs + warp * 8;
  float xh[8][2], rstd[8];
  tile_gemm<C, C>(xh, Es, C, W_(EB_SELF_W), Ws);
  add_rowvec<C>(xh, W_(EB_SELF_B), lane);
  gather_rows<C, false>(xh, a.sl, lw, lane);
  gather_rows<C, false>(xh, a.sl + (size_t)a.n_nodes * C, rw, lane);
  gather_rows<C, false>(xh, a.fl, lw, lane);
  gather_rows<C, false>(xh, a.fr, rw, lane);
  ln_xhat<C>(xh, rstd);
  float d[8][2];
  tile_gemm<C, C>(d, DH, C, W_(T_EB_OUT), Ws);                          // d relu
  ln_relu_bwd<C>(d, xh, rstd, W_(EB_LN_G), W_(EB_LN_BE), lane);          // -> du
  scatter_add_rows<C, true>(d, a.sv.dul, lw, lane);                      // d SL[l], d fl[l]
  scatter_add_rows<C, false>(d, a.sv.dur, rw, lane);                     // d SR[r], d fr[r]
  store_smem<C>(d, A, C, warp, lane);
  tile_gemm<C, C>(d, A, C, W_(T_EB_SELF), Ws);
  combine_smem<C, false>(d, DH, C, warp, lane);                          // + dh' (residual)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = q0 + warp * 8 + i;
    if (q < a.n_edges) store_cols<C>(d[i], a.sv.de + (size_t)q * C, lane);
  }
}

// NodeBlock edge path backward                                                   graph.py:42-50
__global__ void __launch_bounds__(NTHREADS, 1) bwd_edge_nodeblock_kernel(const BwdEdgeArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Es = smem;               // [64][64]
  float* B1 = Es + TM * C;        // [64][256]
  float* B2 = B1 + TM * D;        // [64][256]
  float* Ws = B2 + TM * D;
  float* tes = Ws + 2 * WCHUNK;
  int* ls = reinterpret_cast<int*>(tes + TM);
  int* rs = ls + TM;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * TM;
  const float* blob = a.blob;
  const BlkOff& off = a.off;
  const Tables& tb = a.tb;
  bwd_edge_meta(a, q0, ls, rs, tes);
  load_edge_tile<C>(Es, a.e, q0, a.n_edges);
  __syncthreads();
  const int* lw = ls + warp * 8;
  const int* rw = rs + warp * 8;
  float de[8][2];
  {
    float acc[8][8];
    // ---- forward recompute: msg -> B1, sigmoid(gate) in registers
    tile_gemm<C, D>(acc, Es, C, W_(NB_EN1_W), Ws);
    add_rowvec<D>(acc, W_(NB_EN1_B), lane);
    layernorm_rows<D, true>(acc, W_(NB_EN1_G), W_(NB_EN1_BE), lane);
    store_smem<D>(acc, B1, D, warp, lane);
    tile_gemm<D, D>(acc, B1, D, W_(NB_EN2_W), Ws);
    add_rowvec<D>(acc, W_(NB_EN2_B), lane);
    gather_rows<D, true>(acc, tb.hn, rw, lane);
    store_smem<D>(acc, B1, D, warp, lane);                                // m = he * hn[r]
    tile_gemm<D, D>(acc, B1, D, W_(NB_MSG_W), Ws);
    add_rowvec<D>(acc, W_(NB_MSG_B), lane);
    store_smem<D>(acc, B1, D, warp, lane);                                // msg
    tile_gemm<C, D>(acc, Es, C, W_(NB_GE_W), Ws);
    gather_rows<D, false>(acc, tb.gx, rw, lane);
    layernorm_rows<D, true>(acc, W_(NB_G1_G), W_(NB_G1_BE), lane);
    store_smem<D>(acc, B2, D, warp, lane);
    tile_gemm<D, D>(acc, B2, D, W_(NB_G2_W), Ws);
    add_rowvec<D>(acc, W_(NB_G2_B), lane);
    sigmoid_rows<D>(acc);                                                 // sg
    // ---- out = msg * sg, d out = dagg[l]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp * 8 + i;
      int n = lw[i]; n = n < 0 ? 0 : n;
      float dout[8], msg[8], dgt[8], dmsg[8];
      load_cols<D>(dout, a.sv.dagg + (size_t)n * D, lane);
      load_cols<D>(msg, B1 + r * D, lane);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s = acc[i][j];
        dgt[j] = dout[j] * msg[j] * s * (1.f - s);
        dmsg[j] = dout[j] * s;
      }
      store_cols<D>(dgt, B2 + r * D, lane);
      store_cols<D>(dmsg, B1 + r * D, lane);
    }
    // ---- gate branch
    tile_gemm<D, D>(acc, B2, D, W_(T_NB_G2), Ws);                         // d relu3
    store_smem<D>(acc, B2, D, warp, lane);
    float rstd[8];
    tile_gemm<C, D>(acc, Es, C, W_(NB_GE_W), Ws);                         // recompute a3 -> xhat3
    gather_rows<D, false>(acc, tb.gx, rw, lane);
    ln_xhat<D>(acc, rstd);
    {
      float g[8], b[8];
      load_cols<D>(g, W_(NB_G1_G), lane);
      load_cols<D>(b, W_(NB_G1_BE), lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = warp * 8 + i;
        float d[8];
        load_cols<D>(d, B2 + r * D, lane);
        ln_relu_bwd_row<D>(d, acc[i], rstd[i], g, b);                     // d a3
        store_cols<D>(d, B2 + r * D, lane);
        const int n = rw[i];
        if (n >= 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) atomicAdd(a.sv.dgx + (size_t)n * D + col_of<D>(lane, j), d[j]);
        }
      }
    }
    tile_gemm<D, C>(de, B2, D, W_(T_NB_GE), Ws);                          // de  = d a3 W_ge^T
    // ---- message branch
    tile_gemm<D, D>(acc, B1, D, W_(T_NB_MSG), Ws);                        // dm
    store_smem<D>(acc, B1, D, warp, lane);
    tile_gemm<C, D>(acc, Es, C, W_(NB_EN1_W), Ws);                        // recompute he
    add_rowvec<D>(acc, W_(NB_EN1_B), lane);
    layernorm_rows<D, true>(acc, W_(NB_EN1_G), W_(NB_EN1_BE), lane);
    store_smem<D>(acc, B2, D, warp, lane);
    tile_gemm<D, D>(acc, B2, D, W_(NB_EN2_W), Ws);
    add_rowvec<D>(acc, W_(NB_EN2_B), lane);                               // he
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp * 8 + i;
      const int n = rw[i];
      float dm[8], hn[8];
      load_cols<D>(dm, B1 + r * D, lane);
      load_cols<D>(hn, tb.hn + (size_t)(n < 0 ? 0 : n) * D, lane);
      if (n >= 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(a.sv.dhn + (size_t)n * D + col_of<D>(lane, j), dm[j] * acc[i][j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) dm[j] *= hn[j];                         // d he
      store_cols<D>(dm, B1 + r * D, lane);
    }
    tile_gemm<D, D>(acc, B1, D, W_(T_NB_EN2), Ws);                        // d relu2
    store_smem<D>(acc, B1, D, warp, lane);
    tile_gemm<C, D>(acc, Es, C, W_(NB_EN1_W), Ws);                        // recompute a2 -> xhat2
    add_rowvec<D>(acc, W_(NB_EN1_B), lane);
    ln_xhat<D>(acc, rstd);
    {
      float g[8], b[8];
      load_cols<D>(g, W_(NB_EN1_G), lane);
      load_cols<D>(b, W_(NB_EN1_BE), lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = warp * 8 + i;
        float d[8];
        load_cols<D>(d, B1 + r * D, lane);
        ln_relu_bwd_row<D>(d, acc[i], rstd[i], g, b);                     // d a2
        store_cols<D>(d, B1 + r * D, lane);
      }
    }
  }
  tile_gemm<D, C, true>(de, B1, D, W_(T_NB_EN1), Ws);                     // de += d a2 W_en1^T
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = q0 + warp * 8 + i;
    if (q < a.n_edges) {
      float v[2];
      load_cols<C>(v, a.sv.de + (size_t)q * C, lane);
      v[0] += de[i][0]; v[1] += de[i][1];
      store_cols<C>(v, a.sv.de + (size_t)q * C, lane);
    }
  }
}

// BondFFN left / right backward + the edge_embs backward that closes the block            graph.py:133-141,278-284,357
__global__ void __launch_bounds__(NTHREADS, 1) bwd_edge_bondffn_kernel(const BwdEdgeArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Es = smem;               // [64][64]
  float* A = Es + TM * C;         // [64][128]
  float* Bx = A + TM * 128;       // [64][128] xhat of inter_module LN
  float* Cs = Bx + TM * 128;      // [64][64]
  float* Ws = Cs + TM * C;
  float* tes = Ws + 2 * WCHUNK;
  int* ls = reinterpret_cast<int*>(tes + TM);
  int* rs = ls + TM;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * TM;
  const float* blob = a.blob;
  const BlkOff& off = a.off;
  const Tables& tb = a.tb;
  bwd_edge_meta(a, q0, ls, rs, tes);
  load_edge_tile<C>(Es, a.e, q0, a.n_edges);
  __syncthreads();
  const int* lw = ls + warp * 8;
  const int* rw = rs + warp * 8;
  float de[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { de[i][0] = 0.f; de[i][1] = 0.f; }
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    const int* nw = side ? rw : lw;          // node feeding this FFN
    const int* ow = side ? lw : rw;          // node the output was scattered to
    const float* nl = side ? tb.nlr : tb.nll;
    float* dnl = a.sv.dnl + (size_t)side * a.n_nodes * 128;
    float* dgn = a.sv.dgn + (size_t)side * a.n_nodes * 32;
    float i2[8][2], rstd5[8];
    {  // forward recompute
      float acc[8][4];
      tile_gemm<C, 128>(acc, Es, C, side ? W_(ER_BL_W) : W_(EL_BL_W), Ws);
      gather_rows<128, true>(acc, nl, nw, lane);
      store_smem<128>(acc, A, 128, warp, lane);
      tile_gemm<128, 128>(acc, A, 128, side ? W_(ER_I1_W) : W_(EL_I1_W), Ws);
      add_rowvec<128>(acc, side ? W_(ER_I1_B) : W_(EL_I1_B), lane);
      ln_xhat<128>(acc, rstd5);
      store_smem<128>(acc, Bx, 128, warp, lane);
      affine_relu<128>(acc, side ? W_(ER_I1_G) : W_(EL_I1_G), side ? W_(ER_I1_BE) : W_(EL_I1_BE), lane);
      store_smem<128>(acc, A, 128, warp, lane);
      tile_gemm<128, C>(i2, A, 128, side ? W_(ER_I2_W) : W_(EL_I2_W), Ws);
      add_rowvec<C>(i2, side ? W_(ER_I2_B) : W_(EL_I2_B), lane);
    }
    float xh6[8][1], rstd6[8];
    tile_gemm<C, 32>(xh6, Es, C, side ? W_(ER_GB_W) : W_(EL_GB_W), Ws);
    gather_rows<32, false>(xh6, side ? tb.gnr : tb.gnl, nw, lane);
    add_scaled_rowvec<32>(xh6, side ? W_(ER_GT_W) : W_(EL_GT_W), tes + warp * 8, lane);
    ln_xhat<32>(xh6, rstd6);
    {
      float r6[8][1];
#pragma unroll
      for (int i = 0; i < 8; ++i) r6[i][0] = xh6[i][0];
      affine_relu<32>(r6, side ? W_(ER_G1_G) : W_(EL_G1_G), side ? W_(ER_G1_BE) : W_(EL_G1_BE), lane);
      store_smem<32>(r6, Cs, 32, warp, lane);
    }
    float sgg[8][2];
    tile_gemm<32, C>(sgg, Cs, 32, side ? W_(ER_G2_W) : W_(EL_G2_W), Ws);
    add_rowvec<C>(sgg, side ? W_(ER_G2_B) : W_(EL_G2_B), lane);
    sigmoid_rows<C>(sgg);
    // d o = DUL[r] (left FFN, scattered over right) / DUR[l] (right FFN, scattered over left)
    float dout[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dout[i][0] = 0.f; dout[i][1] = 0.f; }
    gather_rows<C, false>(dout, side ? a.sv.dur : a.sv.dul, ow, lane);
    float dgg[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        dgg[i][j] = dout[i][j] * i2[i][j] * sgg[i][j] * (1.f - sgg[i][j]);
        dout[i][j] *= sgg[i][j];                                            // d i2
      }
    // gate branch
    store_smem<C>(dgg, Cs, C, warp, lane);
    {
      float d6[8][1];
      tile_gemm<C, 32>(d6, Cs, C, side ? W_(T_ER_G2) : W_(T_EL_G2), Ws);
      ln_relu_bwd<32>(d6, xh6, rstd6, side ? W_(ER_G1_G) : W_(EL_G1_G), side ? W_(ER_G1_BE) : W_(EL_G1_BE), lane);
      if (side == 0) scatter_add_rows<32, true>(d6, dgn, nw, lane);
      else           scatter_add_rows<32, false>(d6, dgn, nw, lane);
      store_smem<32>(d6, Cs, 32, warp, lane);
    }
    tile_gemm<32, C, true>(de, Cs, 32, side ? W_(T_ER_GB) : W_(T_EL_GB), Ws);
    // inter branch
    store_smem<C>(dout, Cs, C, warp, lane);
    {
      float d5[8][4];
      tile_gemm<C, 128>(d5, Cs, C, side ? W_(T_ER_I2) : W_(T_EL_I2), Ws);
      float g[4], b[4];
      load_cols<128>(g, side ? W_(ER_I1_G) : W_(EL_I1_G), lane);
      load_cols<128>(b, side ? W_(ER_I1_BE) : W_(EL_I1_BE), lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float xh[4];
        load_cols<128>(xh, Bx + (warp * 8 + i) * 128, lane);
        ln_relu_bwd_row<128>(d5[i], xh, rstd5[i], g, b);
      }
      store_smem<128>(d5, A, 128, warp, lane);
      tile_gemm<128, 128>(d5, A, 128, side ? W_(T_ER_I1) : W_(T_EL_I1), Ws);       // d inter0
      float bl[8][4];
      tile_gemm<C, 128>(bl, Es, C, side ? W_(ER_BL_W) : W_(EL_BL_W), Ws);           // recompute bond_linear(e)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = nw[i];
        float nlv[4];
        load_cols<128>(nlv, nl + (size_t)(n < 0 ? 0 : n) * 128, lane);
        if (n >= 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) atomicAdd(dnl + (size_t)n * 128 + col_of<128>(lane, j), d5[i][j] * bl[i][j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) d5[i][j] *= nlv[j];                              // d bond_linear(e)
      }
      store_smem<128>(d5, A, 128, warp, lane);
    }
    tile_gemm<128, C, true>(de, A, 128, side ? W_(T_ER_BL) : W_(T_EL_BL), Ws);
  }
  // total d e of this block, then through edge_embs: d h_in = de W_EE[:, :64], d rbf += de W_EE[:, 64:80]
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = q0 + warp * 8 + i;
    float v[2] = {0.f, 0.f};
    if (q < a.n_edges) load_cols<C>(v, a.sv.de + (size_t)q * C, lane);
    de[i][0] += v[0]; de[i][1] += v[1];
  }
  store_smem<C>(de, Cs, C, warp, lane);
  {
    float dh[8][2];
    tile_gemm<C, C>(dh, Cs, C, W_(T_EEH), Ws);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int q = q0 + warp * 8 + i;
      if (q < a.n_edges) store_cols<C>(dh[i], a.sv.dh + (size_t)q * C, lane);
    }
    float dg[8][1];
    tile_gemm<C, 32>(dg, Cs, C, W_(T_EEG), Ws);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int q = q0 + warp * 8 + i;
      if (q < a.n_edges && lane < G) a.sv.dg[(size_t)q * G + lane] += dg[i][0];
    }
  }
}

// d rbf -> d dist -> d pos                                   common.py:233-237, graph.py:369-374
__global__ void bwd_pos_kernel(int n_edges, const int* __restrict__ left, const int* __restrict__ right,
                               const float* __restrict__ pos, const float* __restrict__ dg,
                               const float* __restrict__ rbf_off, const float* __restrict__ rbf_coeff,
                               float lo, float hi, const float* __restrict__ gamax, float* __restrict__ d_pos) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_edges) return;
  const int l = left[q], r = right[q];
  const float dx = pos[l * 3 + 0] - pos[r * 3 + 0];
  const float dy = pos[l * 3 + 1] - pos[r * 3 + 1];
  const float dz = pos[l * 3 + 2] - pos[r * 3 + 2];
  const float d = sqrtf(dx * dx + dy * dy + dz * dz);
  if (!(d >= lo && d <= hi) || d == 0.f) return;     // clamp passes no gradient outside [lo, hi]; norm'(0) := 0
  float dd = 0.f;
#pragma unroll
  for (int k = 0; k < G; ++k) {
    const float u = d - rbf_off[k];
    const float g = expf(rbf_coeff[k] * (u * u));
    dd = fmaf(dg[(size_t)q * G + k], g * 2.f * rbf_coeff[k] * u, dd);
  }
  const float s = scalbnf(dd / d, -grad_shift(gamax));
  atomicAdd(d_pos + l * 3 + 0, s * dx); atomicAdd(d_pos + l * 3 + 1, s * dy); atomicAdd(d_pos + l * 3 + 2, s * dz);
  atomicAdd(d_pos + r * 3 + 0, -s * dx); atomicAdd(d_pos + r * 3 + 1, -s * dy); atomicAdd(d_pos + r * 3 + 2, -s * dz);
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
constexpr size_t SMEM_BWD_DEC = (2 * TM * C + TM * 32 + 2 * WCHUNK + 4 * TM) * sizeof(float);
constexpr size_t SMEM_BWD_NODE = SMEM_NODE;
constexpr size_t SMEM_BWD_TAIL = (3 * TM * C + 2 * WCHUNK + 3 * TM) * sizeof(float);
constexpr size_t SMEM_BWD_NB = (TM * C + 2 * TM * D + 2 * WCHUNK + 3 * TM) * sizeof(float);
constexpr size_t SMEM_BWD_FFN = (2 * TM * C + 2 * TM * 128 + 2 * WCHUNK + 3 * TM) * sizeof(float);

int ensure_bwd_attrs() {
  static bool done[64] = {};
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && done[dev]) return MDB_OK;
  CUDA_TRY(cudaFuncSetAttribute(bwd_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BWD_DEC));
  CUDA_TRY(cudaFuncSetAttribute(bwd_node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BWD_NODE));
  CUDA_TRY(cudaFuncSetAttribute(bwd_edge_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BWD_TAIL));
  CUDA_TRY(cudaFuncSetAttribute(bwd_edge_nodeblock_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BWD_NB));
  CUDA_TRY(cudaFuncSetAttribute(bwd_edge_bondffn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BWD_FFN));
  CUDA_TRY(cudaFuncSetAttribute(tc_nodeblock_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_NB_BWD));
  CUDA_TRY(cudaFuncSetAttribute(tc_nodeblock_bwd16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_NB_BWD16));
  CUDA_TRY(cudaFuncSetAttribute(tc_bondffn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_FFN_BWD));
  CUDA_TRY(cudaFuncSetAttribute(tc_edge_tail_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_EDGE_TAIL_BWD));
  CUDA_TRY(cudaFuncSetAttribute(tc_bondffn_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_FFN_BWD2));
  CUDA_TRY(cudaFuncSetAttribute(tc_bondffn_bwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_FFN_BWD3));
  CUDA_TRY(cudaFuncSetAttribute(tc_bwd_node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC_BWD_NODE));
  if (dev >= 0 && dev < 64) done[dev] = true;
  return MDB_OK;
}

int run_bondpred_backward(const mdb_net_desc* net, const mdb_plan* plan, const float* pos, const float* d_logits,
                          float* d_pos, float* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (!net || !plan || !workspace || !d_logits || !d_pos) return fail(MDB_EINVAL, "null argument%s");
  if (net->kind != 2 || net->update_pos) return fail(MDB_EINVAL, "backward is implemented for the bond predictor%s");
  const int N = plan->n_nodes, E = plan->n_edges, L = net->num_blocks;
  if (plan->n_half * 2 != E || E <= 0) return fail(MDB_EINVAL, "edges must be (half, flipped half)%s");
  if (mdb_workspace_bytes(N, E, 1, L) > workspace_bytes) return fail(MDB_EINVAL, "workspace too small%s");
  int rc = ensure_attrs();
  if (rc) return rc;
  rc = ensure_bwd_attrs();
  if (rc) return rc;
  Tables tb;
  const size_t tb_floats = carve(tb, workspace, N, E);
  Saved sv;
  carve_saved(sv, workspace + tb_floats, N, E, L);
  HeadOff head;
  for (int s = 0; s < MDB_NUM_HEAD_SLOTS; ++s) head.o[s] = (int)net->head_off[s];
  const size_t NC = (size_t)N * C, ND = (size_t)N * D, EC = (size_t)E * C;
  const int node_tiles = (N + TM - 1) / TM, edge_tiles = (E + TM - 1) / TM;

  CUDA_TRY(cudaMemsetAsync(sv.ddect, 0, NC * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(sv.dg, 0, (size_t)E * G * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(d_pos, 0, (size_t)N * 3 * sizeof(float), st));

  // Scatter accumulators of block L-1 (every later block's are re-zeroed by bwd_node_kernel after it consumed them).  They
  // are left zero by a completed backward, but a FRESH workspace (torch.empty, recycled allocator blocks) is not: zero them
  // here instead of relying on it.
  CUDA_TRY(cudaMemsetAsync(sv.dul, 0, 2 * NC * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(sv.dur, 0, 2 * NC * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(sv.dgx, 0, (size_t)pad64(N) * D * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(sv.dhn, 0, (size_t)pad64(N) * D * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(sv.dnl, 0, (size_t)2 * N * 128 * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(sv.dgn, 0, (size_t)2 * N * 32 * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(sv.gamax, 0, 32 * sizeof(float), st));
  const size_t n_dl = (size_t)plan->n_half * net->num_edge_types;
  if (n_dl > 0) {
    LAUNCH(MDB_K_bwd_decode, st,
           (grad_amax_kernel<<<(int)std::min<size_t>((n_dl + 255) / 256, 592), 256, 0, st>>>(
               d_logits, n_dl, reinterpret_cast<unsigned*>(sv.gamax))));
  }
  BwdDecArgs da;
  memset(&da, 0, sizeof(da));
  da.gamax = sv.gamax;
  da.blob = net->blob; da.head = head; da.tb = tb; da.left = plan->left; da.right = plan->right; da.inv = plan->inv;
  da.n_half = plan->n_half; da.ke = net->num_edge_types; da.d_logits = d_logits; da.dh = sv.dh; da.ddect = sv.ddect;
  LAUNCH(MDB_K_bwd_decode, st,
         (bwd_decode_kernel<<<(plan->n_half + TM - 1) / TM, NTHREADS, SMEM_BWD_DEC, st>>>(da)));

  BwdNodeArgs na;
  memset(&na, 0, sizeof(na));
  na.blob = net->blob; na.head = head; na.tb = tb; na.sv = sv; na.n_nodes = N;
  BwdEdgeArgs ea;
  memset(&ea, 0, sizeof(ea));
  ea.blob = net->blob; ea.tb = tb; ea.sv = sv; ea.left = plan->left; ea.right = plan->right;
  ea.n_nodes = N; ea.n_edges = E;

  static const bool overlap_env = []() { const char* e = getenv("MDB_OVERLAP"); return e == nullptr || e[0] != '0'; }();
  const bool overlap = overlap_env && !g_profiling;
  SideStream* side = overlap ? side_stream() : nullptr;
  static const bool bwd16_env = []() { const char* e = getenv("MDB_TC_NB_BWD16"); return e == nullptr || e[0] != '0'; }();
  auto nb_tc = [&](int blk) { return net->tc_blob != nullptr && net->blob_host != nullptr && net->tc_block_off[blk][MDB_T_BT_NB_G2] >= 0; };
  // per-node part on tensor cores when every block carries the transposed per-node images (MDB_TC_BWD_NODE=0: fp32 FFMA kernel)
  static const bool tc_bn_env = []() { const char* e = getenv("MDB_TC_BWD_NODE"); return e == nullptr || e[0] != '0'; }();
  bool tc_bn = tc_bn_env && net->tc_blob != nullptr && net->blob_host != nullptr && net->tc_head_off[MDB_TH_BT_EDEC1N] >= 0;
  for (int b = 0; b < L && tc_bn; ++b)
    tc_bn = net->tc_block_off[b][MDB_T_BT_NB_CEN] >= 0 && net->tc_block_off[b][MDB_T_BT_EB_NFL] >= 0 && net->tc_block_off[b][MDB_T_NB_NN1] >= 0;
  for (int i = L - 1; i >= -1; --i) {
    // node kernel: [final] + phase B(i+1) + phase A(i)
    na.do_final = (i == L - 1);
    na.do_B = (i + 1 <= L - 1);
    // layout of dgx / dhn: read (block i+1 scattered into them) AND re-zeroed (for block i) by this launch, so it must be
    // the same for every block of the network -- packing.py packs the tensor-core images for all blocks or for none
    na.red_blocked = (bwd16_env && nb_tc(L - 1)) ? 1 : 0;
    na.do_A = (i >= 0);
    if (na.do_B) { fill_blk(na.blkB, net, i + 1); na.xB = sv.x + (size_t)(i + 1) * ND; }
    const int ib = i < 0 ? 0 : i;
    const Tables tbi = with_block_tables(tb, sv.tabs + (size_t)ib * N * TAB_FLOATS, N, sv.hnb + (size_t)ib * pad64(N) * D,
                                         sv.gxb + (size_t)ib * pad64(N) * D);
    if (na.do_A) {
      fill_blk(na.blkA, net, i);
      na.xA = sv.x + (size_t)i * ND; na.aggA = sv.agg + (size_t)i * ND;
      na.tb = tbi;                      // phase A reads block i's saved centroid_lin(x) table
    }
    // dul / dur are parity-buffered by block: this launch reads block i + 1's sums and re-zeroes that buffer (block i - 1 will
    // accumulate into it); block i's tail kernel accumulates into the other one -- so the node kernel and the tail kernel of
    // block i are independent and run side by side: the node kernel (50 long CTAs) on the high-priority side stream.
    float* dul_rd = sv.dul + (size_t)((i + 1) & 1) * NC;
    float* dur_rd = sv.dur + (size_t)((i + 1) & 1) * NC;
    float* dul_cur = sv.dul + (size_t)(i & 1) * NC;
    float* dur_cur = sv.dur + (size_t)(i & 1) * NC;
    na.sv.dul = dul_rd; na.sv.dur = dur_rd;
    ea.sv.dul = dul_cur; ea.sv.dur = dur_cur;
    const bool forked = overlap && side != nullptr && i >= 0;
    cudaStream_t nst = forked ? side->s : st;
    if (forked) {
      CUDA_TRY(cudaEventRecord(side->fork, st));
      CUDA_TRY(cudaStreamWaitEvent(side->s, side->fork, 0));
    }
    if (tc_bn) {
      // tensor-core node kernel (tc_bwd_node.cuh): 128 nodes per CTA, `dx` node-blocked (private to it)
      TcBwdNodeArgs ba;
      memset(&ba, 0, sizeof(ba));
      ba.tc_blob = reinterpret_cast<const uint8_t*>(net->tc_blob);
      for (int s = 0; s < MDB_NUM_TC_SLOTS; ++s) {
        ba.blkB.o[s] = na.do_B ? net->tc_block_off[i + 1][s] : -1;
        ba.blkA.o[s] = na.do_A ? net->tc_block_off[i][s] : -1;
      }
      for (int s = 0; s < MDB_NUM_TC_HEAD_SLOTS; ++s) ba.hd.o[s] = net->tc_head_off[s];
      ba.n_nodes = N; ba.do_final = na.do_final; ba.do_B = na.do_B; ba.do_A = na.do_A; ba.red_blocked = na.red_blocked;
      ba.xB = na.xB; ba.cenA = tbi.cen; ba.aggA = na.aggA; ba.dx = sv.dx; ba.ddect = sv.ddect;
      ba.dul = dul_rd; ba.dur = dur_rd; ba.dnl = sv.dnl; ba.dgn = sv.dgn; ba.dgx = sv.dgx; ba.dhn = sv.dhn; ba.dagg = sv.dagg;
      fill_bwd_node_vecs(ba.v, net->blob_host, na.do_B ? &na.blkB : nullptr, na.do_A ? &na.blkA : nullptr);
      LAUNCH(MDB_K_tc_node_bwd, nst,
             (tc_bwd_node_kernel<<<(N + tc::ROWS - 1) / tc::ROWS, TC_NB_THREADS, SMEM_TC_BWD_NODE, nst>>>(ba)));
    } else {
      LAUNCH(MDB_K_bwd_node, nst, (bwd_node_kernel<<<node_tiles, NTHREADS, SMEM_BWD_NODE, nst>>>(na)));
    }
    if (forked) CUDA_TRY(cudaEventRecord(side->join, side->s));
    if (i < 0) break;
    fill_blk(ea.off, net, i);
    ea.tb = tbi;
    ea.e = sv.e + (size_t)i * EC; ea.sl = sv.slsr + (size_t)i * 2 * NC; ea.fl = tbi.fl; ea.fr = tbi.fr;
    static const bool tc_tail_env = []() { const char* e = getenv("MDB_TC_TAIL_BWD"); return e == nullptr || e[0] != '0'; }();
    if (tc_tail_env && net->tc_blob != nullptr && net->blob_host != nullptr && net->tc_block_off[i][MDB_T_BT_EB_OUT] >= 0) {
      TcEdgeTailBwdArgs ta;
      memset(&ta, 0, sizeof(ta));
      ta.tc_blob = reinterpret_cast<const uint8_t*>(net->tc_blob);
      for (int s = 0; s < MDB_NUM_TC_SLOTS; ++s) ta.tco.o[s] = net->tc_block_off[i][s];
      ta.left = plan->left; ta.right = plan->right; ta.n_nodes = N; ta.n_edges = E;
      ta.e = ea.e; ta.dh = sv.dh; ta.sl = ea.sl; ta.fl = ea.fl; ta.fr = ea.fr; ta.dul = dul_cur; ta.dur = dur_cur; ta.de = sv.de;
      fill_edge_tail_bwd_vecs(ta.v, net->blob_host, ea.off);
      LAUNCH(MDB_K_tc_edge_tail_bwd, st,
             (tc_edge_tail_bwd_kernel<<<(E + tc::ROWS - 1) / tc::ROWS, TC_NB_THREADS, SMEM_TC_EDGE_TAIL_BWD, st>>>(ta)));
    } else {
      LAUNCH(MDB_K_bwd_edge_tail, st, (bwd_edge_tail_kernel<<<edge_tiles, NTHREADS, SMEM_BWD_TAIL, st>>>(ea)));
    }
    if (forked) CUDA_TRY(cudaStreamWaitEvent(st, side->join, 0));     // the NodeBlock backward needs dagg (node kernel) and de (tail)
    if (nb_tc(i)) {
      if (bwd16_env) {
        TcNbBwd16Args ta;
        memset(&ta, 0, sizeof(ta));
        ta.blob = net->blob; ta.tc_blob = reinterpret_cast<const uint8_t*>(net->tc_blob); ta.off = ea.off; ta.tb = tbi;
        for (int s = 0; s < MDB_NUM_TC_SLOTS; ++s) ta.tco.o[s] = net->tc_block_off[i][s];
        ta.left = plan->left; ta.right = plan->right; ta.n_nodes = N; ta.n_edges = E;
        ta.e = ea.e; ta.dagg = sv.dagg; ta.dgx = sv.dgx; ta.dhn = sv.dhn; ta.de = sv.de; ta.dbg = g_dbg_sel == 1 ? g_dbg_stamps : nullptr;
        ta.scr_he = sv.scr_he; ta.scr_dm = reinterpret_cast<uint8_t*>(sv.scr_dm);
        LAUNCH(MDB_K_tc_nodeblock_bwd, st,
               (tc_nodeblock_bwd16_kernel<<<persistent_grid((E + tc::ROWS - 1) / tc::ROWS), NB16_THREADS, SMEM_TC_NB_BWD16, st>>>(ta)));
      } else {
        TcNbBwdArgs ta;
        memset(&ta, 0, sizeof(ta));
        ta.blob = net->blob; ta.tc_blob = reinterpret_cast<const uint8_t*>(net->tc_blob); ta.off = ea.off; ta.tb = tbi;
        for (int s = 0; s < MDB_NUM_TC_SLOTS; ++s) ta.tco.o[s] = net->tc_block_off[i][s];
        ta.left = plan->left; ta.right = plan->right; ta.n_nodes = N; ta.n_edges = E;
        ta.e = ea.e; ta.dagg = sv.dagg; ta.dgx = sv.dgx; ta.dhn = sv.dhn; ta.de = sv.de; ta.dbg = g_dbg_sel == 1 ? g_dbg_stamps : nullptr;
        fill_nb_vecs(ta.v, net->blob_host, ea.off);
        LAUNCH(MDB_K_tc_nodeblock_bwd, st,
               (tc_nodeblock_bwd_kernel<<<(E + tc::ROWS - 1) / tc::ROWS, tc::RB_THREADS, SMEM_TC_NB_BWD, st>>>(ta)));
      }
    } else {
      LAUNCH(MDB_K_bwd_edge_nodeblock, st, (bwd_edge_nodeblock_kernel<<<edge_tiles, NTHREADS, SMEM_BWD_NB, st>>>(ea)));
    }
    if (net->tc_blob != nullptr && net->blob_host != nullptr && net->tc_block_off[i][MDB_T_BT_EEH] >= 0) {
      TcFfnBwdArgs fa;
      memset(&fa, 0, sizeof(fa));
      fa.tc_blob = reinterpret_cast<const uint8_t*>(net->tc_blob); fa.tb = tbi;
      for (int s = 0; s < MDB_NUM_TC_SLOTS; ++s) fa.tco.o[s] = net->tc_block_off[i][s];
      fa.left = plan->left; fa.right = plan->right; fa.n_nodes = N; fa.n_edges = E;
      fa.e = ea.e; fa.dul = dul_cur; fa.dur = dur_cur; fa.dnl = sv.dnl; fa.dgn = sv.dgn;
      fa.de_in = sv.de; fa.dh = sv.dh; fa.dg = sv.dg;
      fill_ffn_vecs(fa.v, net->blob_host, ea.off, head);
      const bool ffn_bwd2 = []() { const char* e = getenv("MDB_TC_FFN_BWD2"); return e == nullptr || e[0] != '0'; }();   // (read per call: A/B inside one process)
      const bool ffn_bwd3 = []() { const char* e = getenv("MDB_TC_FFN_BWD3"); return e == nullptr || e[0] != '0'; }();
      if (ffn_bwd3)             // 16 row warps (tc_bondffn_bwd3.cuh); MDB_TC_FFN_BWD3=0 -> the 8-row-warp kernels below
        LAUNCH(MDB_K_tc_bondffn_bwd, st,
               (tc_bondffn_bwd3_kernel<<<(E + tc::ROWS - 1) / tc::ROWS, NB16_THREADS, SMEM_TC_FFN_BWD3, st>>>(fa)));
      else if (ffn_bwd2)
        LAUNCH(MDB_K_tc_bondffn_bwd, st,
               (tc_bondffn_bwd2_kernel<<<(E + tc::ROWS - 1) / tc::ROWS, TC_NB_THREADS, SMEM_TC_FFN_BWD2, st>>>(fa)));
      else
        LAUNCH(MDB_K_tc_bondffn_bwd, st,
               (tc_bondffn_bwd_kernel<<<(E + tc::ROWS - 1) / tc::ROWS, TC_NB_THREADS, SMEM_TC_FFN_BWD, st>>>(fa)));
    } else {
      LAUNCH(MDB_K_bwd_edge_bondffn, st, (bwd_edge_bondffn_kernel<<<edge_tiles, NTHREADS, SMEM_BWD_FFN, st>>>(ea)));
    }
  }
  LAUNCH(MDB_K_bwd_pos, st,
         (bwd_pos_kernel<<<(E + 255) / 256, 256, 0, st>>>(E, plan->left, plan->right, pos, sv.dg,
                                                         net->blob + net->head_off[MDB_H_RBF_OFFSET],
                                                         net->blob + net->head_off[MDB_H_RBF_COEFF],
                                                         net->rbf_start, net->rbf_stop, sv.gamax, d_pos)));
  CUDA_TRY(cudaGetLastError());
  return MDB_OK;
}
