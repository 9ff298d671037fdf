// tc_nodeblock_fwd16_kernel: the NodeBlock per-edge path (reference models/graph.py:42-50) on tcgen05 with 16 row warps
// (four threads per tile row, 64 columns each) and ROLLED epilogues.
//
// Why rolled: the fully unrolled register-resident epilogues (tc_nodeblock_fwd_kernel) are 100-270 KB of straight-line
// SASS that every tile executes exactly once -- ncu's source view showed 50-60 % of their issue slots stalled on
// instruction fetch (no_inst).  Here an epilogue is one or two `#pragma unroll 1` loops over 16-column chunks that re-read
// the accumulator from TMEM (cheap) instead of holding 64-128 values in registers: the hot code fits the instruction
// cache, needs < 100 registers, and chunk c of an epilogue is K-slice c of the next GEMM (sliced publication, tc_pipe.cuh).
//
// TMEM: two 256-column accumulators A0 / A1.  An epilogue that re-reads its accumulator while it publishes slices must not
// feed a GEMM that overwrites the same accumulator:  G1 -> A1, G2 -> A0, G3 (msg) -> A1, G4 -> A0, G5 -> A0 (its epilogue
// keeps the 64 pre-LayerNorm values in registers instead of re-reading A0, so it can publish in slices too).
//
// Included by mdb_forward.cu inside its anonymous namespace (after tc_bondffn.cuh).
#pragma once
#include "tc_pipe.cuh"

// phase stamps of the persistent kernels are indexed by tile, not by CTA
#undef TC_STAMP
#define TC_STAMP(i) do { if (a.dbg && tid == 0) a.dbg[(size_t)tile * 32 + (i)] = clock64(); } while (0)

constexpr int NB16_NRW = 16;
constexpr int NB16_THREADS = (NB16_NRW + 4) * 32;   // 640: 4 row warpgroups + {producer, MMA, 2 idle warps}
#ifndef MDB_NB16_SLICED
#define MDB_NB16_SLICED 1   // 1: publish the A planes in 4 K-slices (MMA overlaps the remaining stores); 0: one publication
#endif
constexpr int NB16_PARTS = MDB_NB16_SLICED ? 4 : 0;
constexpr int NB16_NS = 3;                           // weight stages (48 KB): leaves room for the 4-way stat buffer
// smem: E planes | X planes | weight stages | pipe barriers (128 B) | LN stat [4][128] float2 | ls [128] int | vecs [8][256]
constexpr size_t NB16_VEC_OFF = 2 * (size_t)tc::ROWS * C * 2 + 2 * (size_t)tc::ROWS * D * 2 + NB16_NS * tc::STAGE_SLOT
                                + 128 + 4 * tc::ROWS * sizeof(float2) + tc::ROWS * sizeof(int);
static_assert(NB16_VEC_OFF % 16 == 0, "vector block must be 16-byte aligned");
using Pipe16 = tc::PipeT<NB16_NS>;
using PipeSmem16 = tc::PipeSmemT<NB16_NS>;

// running (count, mean, M2) of a row part, merged chunk by chunk (Chan et al.: exact, no cancellation)
struct RunStat { float n, mean, m2; };
__device__ __forceinline__ void stat_add16(RunStat& s, const float (&x)[16]) {
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 16; i += 4) { s4[0] += x[i]; s4[1] += x[i + 1]; s4[2] += x[i + 2]; s4[3] += x[i + 3]; }
  const float mc = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.f / 16.f);
  float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { const float d = x[i + u] - mc; q4[u] = fmaf(d, d, q4[u]); }
  }
  const float nt = s.n + 16.f, delta = mc - s.mean, w = 16.f / nt;
  s.mean = fmaf(delta, w, s.mean);
  s.m2 += (q4[0] + q4[1]) + (q4[2] + q4[3]) + delta * delta * s.n * w;
  s.n = nt;
}
// merge the four 64-column parts of a row (equal counts): returns (mean, rstd) of the 256-wide row
__device__ __forceinline__ float2 ln_merge_quarter(float m_p, float q_p, float2* stat, int row, int part) {
  stat[part * tc::ROWS + row] = make_float2(m_p, q_p);
  asm volatile("bar.sync 1, 512;" ::: "memory");
  const float2 s0 = stat[row], s1 = stat[tc::ROWS + row], s2 = stat[2 * tc::ROWS + row], s3 = stat[3 * tc::ROWS + row];
  asm volatile("bar.sync 1, 512;" ::: "memory");
  const float mean = 0.25f * ((s0.x + s1.x) + (s2.x + s3.x));
  const float d0 = s0.x - mean, d1 = s1.x - mean, d2 = s2.x - mean, d3 = s3.x - mean;
  const float m2 = (s0.y + s1.y) + (s2.y + s3.y) + 64.f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
  return make_float2(mean, 1.f / sqrtf(m2 * (1.f / 256.f) + LN_EPS));
}
// 16 values of row r, columns [k0, k0 + 16) -> the K = 256 A planes (two 16-byte chunks per plane)
__device__ __forceinline__ void store_a16(uint8_t* a_hi, uint8_t* a_lo, int r, int k0, const float (&v)[16]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
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

template <bool IS_ROW, bool XF>
__device__ __forceinline__ void tc_nodeblock_fwd16_body(const TcNbArgs& a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* e_hi = smem_raw;
  uint8_t* e_lo = e_hi + tc::ROWS * C * 2;
  uint8_t* x_hi = e_lo + tc::ROWS * C * 2;
  uint8_t* x_lo = x_hi + tc::ROWS * D * 2;
  uint8_t* stages = x_lo + tc::ROWS * D * 2;
  PipeSmem16* ps = reinterpret_cast<PipeSmem16*>(stages + NB16_NS * tc::STAGE_SLOT);
  float2* stat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(ps) + 128);     // [4][128]
  int* ls = reinterpret_cast<int*>(stat + 4 * tc::ROWS);
  // Per-column parameter vectors in shared memory (copied from the fp32 blob with one coalesced 16-byte load per row
  // thread): the epilogues read them with warp-uniform LDS.128 -- four values per instruction.  Register-indexed
  // constant-bank loads (LDC, one or two values each) made the rolled LayerNorm epilogue MIO-bound (ncu: 46 % of its
  // samples on LDC).
  float* vecs = reinterpret_cast<float*>(smem_raw + NB16_VEC_OFF);
  const float *v_en1_b = vecs, *v_en1_g = vecs + D, *v_en1_be = vecs + 2 * D, *v_en2_b = vecs + 3 * D, *v_msg_b = vecs + 4 * D,
              *v_g1_g = vecs + 5 * D, *v_g1_be = vecs + 6 * D, *v_g2_b = vecs + 7 * D;
  float* out_tile = reinterpret_cast<float*>(smem_raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Tables& tb = a.tb;
  const int n_tiles = (a.n_edges + tc::ROWS - 1) / tc::ROWS;
  tc::PipeT<NB16_NS, XF> p;
  tc::pipe_init_split<NB16_NRW, IS_ROW, NB16_NS>(p, ps, stages);
  if (warp == NB16_NRW) tc::tmem_alloc<512>(&ps->tmem_base);
  const int row = (warp & 3) * 32 + lane;
  const int part = (warp >> 2) & 3;
  const int pc = part * 64;                   // first column of this thread's quarter
  if (IS_ROW) {
    const int vj = tid >> 6;                    // 64 threads per vector
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
  const uint32_t A0 = lane_base + pc, A1 = lane_base + 256 + pc;     // this thread's part of the two accumulators

  // Persistent CTA: TMEM, barriers and the parameter vectors are set up once; the pipeline counters and mbarrier phases
  // simply keep running across tiles, and the producer prefetches the next tile's first weight stages during the tail.
  // this thread's inputs of a tile: 16 columns of e and the row's endpoints.  Loaded one tile AHEAD (during the previous
  // tile's last epilogue), so the tile starts without an exposed L2 round trip.
  float e16[16];
  int nxt_r = -1, nxt_l = -1;
  auto load_tile_inputs = [&](int t) {
    const int q = t * tc::ROWS + row;
    const bool in = t < n_tiles && q < a.n_edges;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in) x = *reinterpret_cast<const float4*>(a.ebuf + (size_t)q * C + part * 16 + i);
      e16[i] = x.x; e16[i + 1] = x.y; e16[i + 2] = x.z; e16[i + 3] = x.w;
    }
    nxt_r = in ? a.right[q] : -1;
    nxt_l = in ? a.left[q] : -1;
  };
  if (IS_ROW) load_tile_inputs(blockIdx.x);
#pragma unroll 1
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
  TC_STAMP(0);
  if (a.dbg) p.dbg = a.dbg + (size_t)tile * 32;
  const int my_r = nxt_r;
  if (IS_ROW && part == 0) ls[row] = nxt_l;
  const int rr = my_r < 0 ? 0 : my_r;
  TC_STAMP(1);

  if (IS_ROW) {   // e tile -> E planes: 16 columns per thread
    tc::store_a<C, 16>(e_hi, e_lo, row, part * 16, e16);
    tc::rows_publish(p);
    TC_STAMP(6);
  }
  // G1: edge_net.net.0 -> A1                                                      graph.py:42
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_EN1), 256, false, true, true);
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    TC_STAMP(2);
    RunStat rs = {0.f, 0.f, 0.f};
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float x[16];
      tc::tmem_ld16(A1 + c * 16, x);
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] += v_en1_b[pc + c * 16 + i];
      stat_add16(rs, x);
    }
    const float2 ms = ln_merge_quarter(rs.mean, rs.m2, stat, row, part);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float x[16];
      tc::tmem_ld16(A1 + c * 16, x);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = pc + c * 16 + i;
        x[i] = fmaxf((x[i] + v_en1_b[k] - ms.x) * ms.y * v_en1_g[k] + v_en1_be[k], 0.f);
      }
      store_a16(x_hi, x_lo, row, pc + c * 16, x);
      if (MDB_NB16_SLICED) tc::rows_publish_group(p, c);
    }
    if (!MDB_NB16_SLICED) tc::rows_publish(p);
    TC_STAMP(7);
  }
  // G2: edge_net.net.3 -> A0 ; m = he * node_net(x)[col]                           graph.py:43
  tc::gemm<D, D, NB16_NS, NB16_PARTS>(p, x_hi, x_lo, TCW_(NB_EN2), 0, false, true, true);
  if (IS_ROW) {
    const float* hn = tb.hnb + blk_off(rr, pc / 4);     // node-blocked copy: ~4 lines per warp load instead of 32
    float hv[64];                                // gathered node_net(x)[col] row part: requested BEFORE the accumulator wait
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
      const float4 t4 = *reinterpret_cast<const float4*>(hn + (i / 4) * BLK_PIECE_STRIDE);
      hv[i] = t4.x; hv[i + 1] = t4.y; hv[i + 2] = t4.z; hv[i + 3] = t4.w;
    }
    tc::rows_wait_acc(p);
    TC_STAMP(3);
#pragma unroll
    for (int c = 0; c < 4; ++c) {                // unrolled: hv[] must stay in registers
      float x[16];
      tc::tmem_ld16(A0 + c * 16, x);
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = (x[i] + v_en2_b[pc + c * 16 + i]) * hv[c * 16 + i];
      store_a16(x_hi, x_lo, row, pc + c * 16, x);
      if (MDB_NB16_SLICED) tc::rows_publish_group(p, c);
    }
    if (!MDB_NB16_SLICED) tc::rows_publish(p);
    TC_STAMP(8);
  }
  // G3: msg_net -> A1 (stays in TMEM) ; G4: gate.net.0 edge columns -> A0          graph.py:43,46
  tc::gemm<D, D, NB16_NS, NB16_PARTS>(p, x_hi, x_lo, TCW_(NB_MSG), 256, false, true, false);
  tc::gemm<C, D>(p, e_hi, e_lo, TCW_(NB_GE), 0, false, false, true);
  if (IS_ROW) {
    const float* gxr = tb.gxb + blk_off(rr, pc / 4);
    float gv[64];                                // hoisted node / time / bias part of gate.net.0, gathered before the wait
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
      const float4 t4 = *reinterpret_cast<const float4*>(gxr + (i / 4) * BLK_PIECE_STRIDE);
      gv[i] = t4.x; gv[i + 1] = t4.y; gv[i + 2] = t4.z; gv[i + 3] = t4.w;
    }
    tc::rows_wait_acc(p);
    TC_STAMP(4);
    RunStat rs = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c) {                // unrolled: gv[] <- acc + gathered row stays in registers, so the
      float x[16];                               // accumulator is drained before the first slice is published
      tc::tmem_ld16(A0 + c * 16, x);
#pragma unroll
      for (int i = 0; i < 16; ++i) { x[i] += gv[c * 16 + i]; gv[c * 16 + i] = x[i]; }
      stat_add16(rs, x);
    }
    const float2 ms = ln_merge_quarter(rs.mean, rs.m2, stat, row, part);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float x[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = pc + c * 16 + i;
        x[i] = fmaxf((gv[c * 16 + i] - ms.x) * ms.y * v_g1_g[k] + v_g1_be[k], 0.f);
      }
      store_a16(x_hi, x_lo, row, pc + c * 16, x);
      if (MDB_NB16_SLICED) tc::rows_publish_group(p, c);
    }
    if (!MDB_NB16_SLICED) tc::rows_publish(p);
    TC_STAMP(9);
  }
  // G5: gate.net.3 -> A0                                                          graph.py:46
  tc::gemm<D, D, NB16_NS, NB16_PARTS>(p, x_hi, x_lo, TCW_(NB_G2), 0, false, true, true);
  if (IS_ROW) {
    tc::rows_wait_acc(p);
    TC_STAMP(5);
    load_tile_inputs(tile + gridDim.x);          // next tile's inputs: in flight under this epilogue and the reduction
    // out = (msg + b) * sigmoid(gate + b)  -> smem tile (all operand planes are dead now)     graph.py:47
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float g[16], m[16];
      tc::tmem_ld16(A0 + c * 16, g);
      tc::tmem_ld16(A1 + c * 16, m);
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const int k = pc + c * 16 + i;
        float4 o;
        o.x = (m[i] + v_msg_b[k]) * tc::fast_sigmoid(g[i] + v_g2_b[k]);
        o.y = (m[i + 1] + v_msg_b[k + 1]) * tc::fast_sigmoid(g[i + 1] + v_g2_b[k + 1]);
        o.z = (m[i + 2] + v_msg_b[k + 2]) * tc::fast_sigmoid(g[i + 2] + v_g2_b[k + 2]);
        o.w = (m[i + 3] + v_msg_b[k + 3]) * tc::fast_sigmoid(g[i + 3] + v_g2_b[k + 3]);
        *reinterpret_cast<float4*>(out_tile + row * OUT_LD + k) = o;
      }
    }
    tc::fence_before_sync();
    asm volatile("bar.sync 1, 512;" ::: "memory");
    // scatter_sum over row (= left): thread (c, h) owns channel c and walks rows [64 h, 64 h + 64)   graph.py:50
    const int c = tid & 255, r0 = (tid >> 8) * 64;
    int cur = ls[r0];
    float s0 = 0.f;
#pragma unroll 1
    for (int r8 = r0; r8 < r0 + 64; r8 += 8) {
      float x[8];
      int n[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { x[u] = out_tile[(r8 + u) * OUT_LD + c]; n[u] = ls[r8 + u]; }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (n[u] != cur) {
          if (cur >= 0) atomicAdd(tb.agg + (size_t)cur * D + c, s0);
          cur = n[u]; s0 = 0.f;
        }
        s0 += x[u];
      }
    }
    if (cur >= 0) atomicAdd(tb.agg + (size_t)cur * D + c, s0);
    asm volatile("bar.sync 1, 512;" ::: "memory");   // the next tile overwrites ls / the planes the out tile aliases
  }
  TC_STAMP(15);
  }   // tile loop
  tc::cta_sync();
  if (warp == NB16_NRW) { __syncwarp(); tc::tmem_dealloc<512>(ps->tmem_base); }
}

// 640 threads compile to 96 registers; the producer/MMA warpgroup keeps 32, the four row warpgroups get 112 (inc and dec must balance inside the CTA pool: 512 x 16 = 128 x 64 -- an unbalanced inc blocks forever).
// XF: cross-first accumulation order (tc_pipe.cuh) -- the bond predictor's forward, whose activations feed the guidance gradient
template <bool XF>
__global__ void __launch_bounds__(NB16_THREADS, 1) tc_nodeblock_fwd16_kernel(const __grid_constant__ TcNbArgs a) {
  if (threadIdx.x < NB16_NRW * 32) {
    tc::reg_alloc<112>();
    tc_nodeblock_fwd16_body<true, XF>(a);
  } else {
    tc::reg_dealloc<32>();
    tc_nodeblock_fwd16_body<false, XF>(a);
  }
}

constexpr size_t SMEM_TC_NB16 = NB16_VEC_OFF + 8 * D * sizeof(float);
static_assert(sizeof(PipeSmem16) <= 128, "PipeSmem must fit its 128-byte slot");
static_assert(SMEM_TC_NB16 <= 232448, "tc_nodeblock_fwd16_kernel exceeds the shared-memory limit");
