// Tensor-core (tcgen05) backward of the EdgeBlock tail (forward: reference models/graph.py:286-294; same math as the fp32
// bwd_edge_tail_kernel it replaces):
//     u = W_self e + b + SL[l] + SR[r] + fl[l] + fr[r] ;  h' = e + W_out relu(LN(u)) + b_out
//     d relu = dh' W_out^T ;  du = LNbwd(d relu; u) ;  DUL[l] += du ;  DUR[r] += du ;  de = dh' + du W_self^T
// Three 64 x 64 GEMMs per 128-edge tile -- far too little tensor work to fill an SM from one tile, so the kernel is sized for
// TWO resident CTAs per SM (83 KB of shared memory, 128 TMEM columns, <= 80 registers): the hardware interleaves the
// GEMM round trips of one tile with the epilogue of the other.
//
// Included by mdb_forward.cu inside its anonymous namespace (after tc_bondffn.cuh).
#pragma once
#include "tc_pipe.cuh"

struct EdgeTailBwdVecs { float self_b[C], ln_g[C], ln_be[C]; };

struct TcEdgeTailBwdArgs {
  const uint8_t* tc_blob;
  TcOff tco;
  const int *left, *right;
  int n_nodes, n_edges;
  const float* e;          // [E][64] saved e_i
  const float* dh;         // [E][64] d/d h_edge leaving the block (sorted order)
  const float* sl;         // [2][N][64] saved SL / SR of block i
  const float *fl, *fr;    // [N][64] node_ffn_{left,right}(h_node) of block i
  float *dul, *dur;        // [N][64] scatter targets (pre-zeroed)
  float* de;               // [E][64] out: d/d e (tail part; the NodeBlock / BondFFN kernels add theirs)
  EdgeTailBwdVecs v;
};

constexpr uint32_t TAIL_SLOT_BYTES = tc::WStage<C, tc::KB>::STAGE_BYTES;   // 4 KB: every GEMM here has N = 64

// [128][64] fp32 tile in shared memory, 16-byte chunks XOR-swizzled by row so that both the row-major writer (lane = row)
// and the column reader (lane = channel) are bank-conflict free
__device__ __forceinline__ int tail_tile_off(int r, int c) { return r * C + ((((c >> 2) ^ r) & 15) << 2) + (c & 3); }

__global__ void __launch_bounds__(TC_NB_THREADS, 2) tc_edge_tail_bwd_kernel(const __grid_constant__ TcEdgeTailBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* e_hi = smem_raw;                                  // e planes (K = 64); the du tile for the run reduction later
  uint8_t* e_lo = e_hi + tc::ROWS * C * 2;
  uint8_t* d_hi = e_lo + tc::ROWS * C * 2;                   // dh' planes, then du planes
  uint8_t* d_lo = d_hi + tc::ROWS * C * 2;
  uint8_t* stages = d_lo + tc::ROWS * C * 2;
  tc::PipeSmem* ps = reinterpret_cast<tc::PipeSmem*>(stages + tc::NSTAGE * TAIL_SLOT_BYTES);
  float2* stat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(ps) + 128);
  int* ls = reinterpret_cast<int*>(stat + 2 * tc::ROWS);
  float* du_tile = reinterpret_cast<float*>(e_hi);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * tc::ROWS;
  tc::Pipe p;
  tc::pipe_init<TC_NRW>(p, ps, stages);
  p.slot_bytes = TAIL_SLOT_BYTES;
  if (warp == TC_NRW) tc::tmem_alloc<128>(&ps->tmem_base);
  const int row = (warp & 3) * 32 + lane;
  const int half = (warp >> 2) & 1;
  const int c32 = half * 32;
  const int q = q0 + row;
  const bool valid = p.role == 0 && q < a.n_edges;
  const int ll = valid ? a.left[q] : 0, rr = valid ? a.right[q] : 0;
  if (p.role == 0 && half == 0) ls[row] = valid ? ll : -1;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t lane_base = ps->tmem_base + ((uint32_t)((warp & 3) * 32) << 16);

  if (p.role == 0) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {                   // e -> E planes, dh' -> D planes (one 32-float array live)
      const float* src = (pass ? a.dh : a.e) + (size_t)q * C + c32;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) x = *reinterpret_cast<const float4*>(src + i);
        v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
      }
      tc::store_a<C, 32>(pass ? d_hi : e_hi, pass ? d_lo : e_lo, row, c32, v);
    }
    tc::rows_publish(p);
  }
  // recompute W_self e -> D[0:64] ;  d relu = dh' W_out^T -> D[64:128]
  tc::gemm<C, C>(p, e_hi, e_lo, TCW_(EB_SELF), 0, false, true, false);
  tc::gemm<C, C>(p, d_hi, d_lo, TCW_(BT_EB_OUT), 64, false, false, true);
  if (p.role == 0) {
    const float* t0 = a.sl + (size_t)ll * C + c32;
    const float* t1 = a.sl + (size_t)a.n_nodes * C + (size_t)rr * C + c32;
    const float* t2 = a.fl + (size_t)ll * C + c32;
    const float* t3 = a.fr + (size_t)rr * C + c32;
    float u[32];                                             // gathers requested before the accumulator wait
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 x0 = *reinterpret_cast<const float4*>(t0 + i), x1 = *reinterpret_cast<const float4*>(t1 + i);
      const float4 x2 = *reinterpret_cast<const float4*>(t2 + i), x3 = *reinterpret_cast<const float4*>(t3 + i);
      u[i] = x0.x + x1.x + x2.x + x3.x; u[i + 1] = x0.y + x1.y + x2.y + x3.y;
      u[i + 2] = x0.z + x1.z + x2.z + x3.z; u[i + 3] = x0.w + x1.w + x2.w + x3.w;
    }
    tc::rows_wait_acc(p);
    {
      float acc[32];
      tc::tmem_ld32(lane_base + c32, acc);
#pragma unroll
      for (int i = 0; i < 32; ++i) u[i] += acc[i] + a.v.self_b[c32 + i];
    }
    // LayerNorm statistics of the 64-wide row (two halves merged exactly, as ln_relu_part)
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += u[i];
    const float m_h = s * (1.f / 32.f);
    float qq = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { const float d = u[i] - m_h; qq = fmaf(d, d, qq); }
    const float2 o = exchange_half(stat, row, half, m_h, qq);
    const float mean = 0.5f * (m_h + o.x);
    const float dm = m_h - o.x;
    const float rstd = 1.f / sqrtf((qq + o.y + dm * dm * 16.f) * (1.f / 64.f) + LN_EPS);
    float d[32];
    tc::tmem_ld32(lane_base + 64 + c32, d);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float xh = (u[i] - mean) * rstd;
      const float g = a.v.ln_g[c32 + i];
      const float dxh = (xh * g + a.v.ln_be[c32 + i] > 0.f) ? d[i] * g : 0.f;
      u[i] = xh;
      d[i] = dxh;
      s1 += dxh;
      s2 = fmaf(dxh, xh, s2);
    }
    const float2 o2 = exchange_half(stat, row, half, s1, s2);
    const float m1 = (s1 + o2.x) * (1.f / 64.f), m2 = (s2 + o2.y) * (1.f / 64.f);
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = valid ? rstd * (d[i] - m1 - u[i] * m2) : 0.f;      // du
    // DUR[r] += du: right nodes are unsorted -> one 16-byte RED per 4 channels
    if (valid) {
      float* dst = a.dur + (size_t)rr * C + c32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) tc::red_add_v4(dst + i, d[i], d[i + 1], d[i + 2], d[i + 3]);
    }
    // du -> A planes (dh' planes are dead: BT_EB_OUT has completed) and -> the fp32 tile over the dead e planes
    tc::store_a<C, 32>(d_hi, d_lo, row, c32, d);
    tc::rows_publish(p);
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      *reinterpret_cast<float4*>(du_tile + tail_tile_off(row, c32 + i)) = make_float4(d[i], d[i + 1], d[i + 2], d[i + 3]);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // DUL[l] += du: left nodes form contiguous CSR runs -> reduce each run inside the tile, one atomic per (node, channel)
    {
      const int c = tid & 63, r0 = (tid >> 6) * 32;            // 4 row quarters x 64 channels
      int cur = ls[r0];
      float acc = 0.f;
      for (int r = r0; r < r0 + 32; ++r) {
        const int n = ls[r];
        if (n != cur) {
          if (cur >= 0) atomicAdd(a.dul + (size_t)cur * C + c, acc);
          cur = n; acc = 0.f;
        }
        acc += du_tile[tail_tile_off(r, c)];
      }
      if (cur >= 0) atomicAdd(a.dul + (size_t)cur * C + c, acc);
    }
  }
  // de = dh' + du W_self^T
  tc::gemm<C, C>(p, d_hi, d_lo, TCW_(BT_EB_SELF), 0, false, true, true);
  if (p.role == 0) {
    tc::rows_wait_acc(p);
    float v[32];
    tc::tmem_ld32(lane_base + c32, v);
    if (valid) {                                             // + dh' (residual; re-read: an L2 hit is cheaper than 32 registers)
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 r4 = *reinterpret_cast<const float4*>(a.dh + (size_t)q * C + c32 + i);
        *reinterpret_cast<float4*>(a.de + (size_t)q * C + c32 + i) =
            make_float4(v[i] + r4.x, v[i + 1] + r4.y, v[i + 2] + r4.z, v[i + 3] + r4.w);
      }
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == TC_NRW) { __syncwarp(); tc::tmem_dealloc<128>(ps->tmem_base); }
}

constexpr size_t SMEM_TC_EDGE_TAIL_BWD = 4 * (size_t)tc::ROWS * C * 2 + tc::NSTAGE * TAIL_SLOT_BYTES + 128
                                         + 2 * tc::ROWS * sizeof(float2) + tc::ROWS * sizeof(int) + 64;
static_assert(SMEM_TC_EDGE_TAIL_BWD <= 113 * 1024, "tc_edge_tail_bwd_kernel must fit twice per SM");

inline void fill_edge_tail_bwd_vecs(EdgeTailBwdVecs& v, const float* blob_host, const BlkOff& off) {
  auto cp = [&](float* dst, int slot, int n) { memcpy(dst, blob_host + off.o[slot], n * sizeof(float)); };
  cp(v.self_b, MDB_S_EB_SELF_B, C); cp(v.ln_g, MDB_S_EB_LN_G, C); cp(v.ln_be, MDB_S_EB_LN_BE, C);
}
