// Tile engine shared by every kernel of the path: a 64-row activation tile lives in shared memory,
// each of the CTA's 8 warps owns 8 rows of it end-to-end (so chained layers need no CTA barrier for
// the activations), and every Linear of the reference becomes  acc[8][N/32] = X_tile[8 rows][K] * W[K][N]
// with the weight panel streamed L2 -> smem by cp.async in 16 KB double-buffered chunks.
//
// fp32 FFMA path (parity mode, ~1e-6 of the fp32 reference).  The tcgen05 path replaces tile_gemm only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mdb {

constexpr int TM = 64;            // rows (edges or nodes) per CTA tile
constexpr int NTHREADS = 256;     // 8 warps x 8 rows
constexpr int WCHUNK = 4096;      // floats per weight stage (16 KB), two stages
constexpr float LN_EPS = 1e-5f;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Column owned by `lane` in slot j of its N/32 columns.  Layouts are chosen so that a warp's smem /
// global accesses are contiguous 16 B (N>=128), 8 B (N=64) or 4 B (N=32) per lane.
template <int N>
__device__ __forceinline__ int col_of(int lane, int j) {
  if constexpr (N == 256) return (j < 4) ? (4 * lane + j) : (128 + 4 * lane + (j - 4));
  else if constexpr (N == 128) return 4 * lane + j;
  else if constexpr (N == 64) return 2 * lane + j;
  else return lane;
}

// Node-blocked layout of a [n_nodes][256] per-node table: [node / 32][16-byte column piece 0..63][node % 32][4 floats].
// A warp whose lanes hold (mostly) consecutive nodes -- the tensor-core edge kernels: one thread per CSR-ordered edge,
// gathering by `right` -- then touches ~4 cache lines per 16-byte load / RED instead of 32 (row-major rows are 1 KB
// apart), which is what bounded their epilogues (LSU tag stage: one line per cycle).  Tables are padded to 64 nodes.
__host__ __device__ __forceinline__ size_t blk_off(int n, int piece) {
  return ((size_t)(n >> 5) * 64 + piece) * 128 + (size_t)(n & 31) * 4;
}
constexpr int BLK_PIECE_STRIDE = 128;   // floats between consecutive 16-byte pieces of the same node
// this lane's 8 columns (pieces lane and 32 + lane) of node n
__device__ __forceinline__ void store_cols_blocked256(const float (&v)[8], float* table, int n, int lane) {
  *reinterpret_cast<float4*>(table + blk_off(n, lane)) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(table + blk_off(n, 32 + lane)) = make_float4(v[4], v[5], v[6], v[7]);
}

// Load / store this lane's N/32 columns of one row (row pointer p, 16-byte aligned rows).
template <int N>
__device__ __forceinline__ void load_cols(float (&v)[N / 32], const float* p, int lane) {
  if constexpr (N == 256) {
    float4 a = *reinterpret_cast<const float4*>(p + 4 * lane);
    float4 b = *reinterpret_cast<const float4*>(p + 128 + 4 * lane);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else if constexpr (N == 128) {
    float4 a = *reinterpret_cast<const float4*>(p + 4 * lane);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else if constexpr (N == 64) {
    float2 a = *reinterpret_cast<const float2*>(p + 2 * lane);
    v[0] = a.x; v[1] = a.y;
  } else {
    v[0] = p[lane];
  }
}
template <int N>
__device__ __forceinline__ void store_cols(const float (&v)[N / 32], float* p, int lane) {
  if constexpr (N == 256) {
    *reinterpret_cast<float4*>(p + 4 * lane) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 128 + 4 * lane) = make_float4(v[4], v[5], v[6], v[7]);
  } else if constexpr (N == 128) {
    *reinterpret_cast<float4*>(p + 4 * lane) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (N == 64) {
    *reinterpret_cast<float2*>(p + 2 * lane) = make_float2(v[0], v[1]);
  } else {
    p[lane] = v[0];
  }
}

// acc (+)= X[warp rows][0:K] * W[K][N].  Xs: smem, leading dim ldx (multiple of 4 floats).  Wg: global,
// 16-byte aligned.  Ws: smem staging, 2*WCHUNK floats.  Contains CTA barriers: all 256 threads must call.
template <int K, int N, bool ACCUM = false>
__device__ __forceinline__ void tile_gemm(float (&acc)[8][N / 32], const float* __restrict__ Xs, int ldx,
                                          const float* __restrict__ Wg, float* __restrict__ Ws) {
  constexpr int CPL = N / 32;
  constexpr int KC = (WCHUNK / N) < K ? (WCHUNK / N) : K;
  constexpr int NCH = (K + KC - 1) / KC;
  static_assert(K % 4 == 0 && KC % 4 == 0, "K must be a multiple of 4");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if constexpr (!ACCUM) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[i][j] = 0.f;
  }
  auto stage = [&](int c) {
    const int k0 = c * KC;
    const int rows = (K - k0) < KC ? (K - k0) : KC;
    const float4* src = reinterpret_cast<const float4*>(Wg + (size_t)k0 * N);
    float4* dst = reinterpret_cast<float4*>(Ws + (c & 1) * WCHUNK);
    const int n4 = rows * N / 4;
    for (int i = tid; i < n4; i += NTHREADS) cp_async16(dst + i, src + i);
    cp_async_commit();
  };
  stage(0);
  const float* xbase = Xs + (warp * 8) * ldx;
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    if (c + 1 < NCH) { stage(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const int k0 = c * KC;
    const int rows = (K - k0) < KC ? (K - k0) : KC;
    const float* Wb = Ws + (c & 1) * WCHUNK;
#pragma unroll 1
    for (int kk = 0; kk < rows; kk += 4) {
      float4 xv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(xbase + i * ldx + k0 + kk);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float w[CPL];
        load_cols<N>(w, Wb + (kk + u) * N, lane);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = (u == 0) ? xv[i].x : (u == 1) ? xv[i].y : (u == 2) ? xv[i].z : xv[i].w;
#pragma unroll
          for (int j = 0; j < CPL; ++j) acc[i][j] = fmaf(x, w[j], acc[i][j]);
        }
      }
    }
    __syncthreads();
  }
}

// ---- epilogue helpers on the warp's acc[8][N/32] --------------------------------------------------
template <int N>
__device__ __forceinline__ void add_rowvec(float (&acc)[8][N / 32], const float* __restrict__ b, int lane) {
  float v[N / 32];
  load_cols<N>(v, b, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < N / 32; ++j) acc[i][j] += v[j];
}

// acc[i] += s[i] * vec   (time column of a split first layer: w_t * t_row)
template <int N>
__device__ __forceinline__ void add_scaled_rowvec(float (&acc)[8][N / 32], const float* __restrict__ vec,
                                                  const float* s_rows /* smem [8] of this warp */, int lane) {
  float v[N / 32];
  load_cols<N>(v, vec, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float s = s_rows[i];
#pragma unroll
    for (int j = 0; j < N / 32; ++j) acc[i][j] = fmaf(s, v[j], acc[i][j]);
  }
}

// acc[i] (op)= table[idx[i]][:]; idx from smem (this warp's 8 rows), negative idx (padding rows) read row 0.
template <int N, bool MUL>
__device__ __forceinline__ void gather_rows(float (&acc)[8][N / 32], const float* __restrict__ table,
                                            const int* idx_rows, int lane) {
  float v[8][N / 32];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int n = idx_rows[i];
    n = n < 0 ? 0 : n;
    load_cols<N>(v[i], table + (size_t)n * N, lane);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < N / 32; ++j) acc[i][j] = MUL ? acc[i][j] * v[i][j] : acc[i][j] + v[i][j];
}

template <int N, bool MUL>
__device__ __forceinline__ void combine_smem(float (&acc)[8][N / 32], const float* Ys, int ldy, int warp, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v[N / 32];
    load_cols<N>(v, Ys + (warp * 8 + i) * ldy, lane);
#pragma unroll
    for (int j = 0; j < N / 32; ++j) acc[i][j] = MUL ? acc[i][j] * v[j] : acc[i][j] + v[j];
  }
}

template <int N>
__device__ __forceinline__ void store_smem(const float (&acc)[8][N / 32], float* Ys, int ldy, int warp, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) store_cols<N>(acc[i], Ys + (warp * 8 + i) * ldy, lane);
}

// Row-wise LayerNorm over the N columns (two-pass, biased variance, eps 1e-5 as torch) + optional ReLU.
template <int N, bool RELU>
__device__ __forceinline__ void layernorm_rows(float (&acc)[8][N / 32], const float* __restrict__ gamma,
                                               const float* __restrict__ beta, int lane) {
  float g[N / 32], b[N / 32];
  load_cols<N>(g, gamma, lane);
  load_cols<N>(b, beta, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < N / 32; ++j) s += acc[i][j];
    const float mean = warp_sum(s) * (1.f / N);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < N / 32; ++j) { const float d = acc[i][j] - mean; q = fmaf(d, d, q); }
    const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / N) + LN_EPS);
#pragma unroll
    for (int j = 0; j < N / 32; ++j) {
      float y = (acc[i][j] - mean) * rstd * g[j] + b[j];
      acc[i][j] = RELU ? fmaxf(y, 0.f) : y;
    }
  }
}

// LayerNorm forward that keeps what the backward needs: acc <- xhat = (x - mean) * rstd, rstd[i] per row.
template <int N>
__device__ __forceinline__ void ln_xhat(float (&acc)[8][N / 32], float (&rstd)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < N / 32; ++j) s += acc[i][j];
    const float mean = warp_sum(s) * (1.f / N);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < N / 32; ++j) { const float d = acc[i][j] - mean; q = fmaf(d, d, q); }
    rstd[i] = 1.f / sqrtf(warp_sum(q) * (1.f / N) + LN_EPS);
#pragma unroll
    for (int j = 0; j < N / 32; ++j) acc[i][j] = (acc[i][j] - mean) * rstd[i];
  }
}

// acc (= xhat) <- relu(xhat * gamma + beta)
template <int N>
__device__ __forceinline__ void affine_relu(float (&acc)[8][N / 32], const float* __restrict__ gamma,
                                            const float* __restrict__ beta, int lane) {
  float g[N / 32], b[N / 32];
  load_cols<N>(g, gamma, lane);
  load_cols<N>(b, beta, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < N / 32; ++j) acc[i][j] = fmaxf(acc[i][j] * g[j] + b[j], 0.f);
}

// One row of the backward of y = relu(LN(x) * gamma + beta):  d (grad wrt y) <- grad wrt x.
//   dxhat = d * 1[xhat*g + b > 0] * g ;  dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))
template <int N>
__device__ __forceinline__ void ln_relu_bwd_row(float (&d)[N / 32], const float (&xh)[N / 32], float rstd,
                                                const float (&g)[N / 32], const float (&b)[N / 32]) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < N / 32; ++j) {
    const float y = xh[j] * g[j] + b[j];
    d[j] = (y > 0.f) ? d[j] * g[j] : 0.f;
    s1 += d[j];
    s2 = fmaf(d[j], xh[j], s2);
  }
  const float m1 = warp_sum(s1) * (1.f / N), m2 = warp_sum(s2) * (1.f / N);
#pragma unroll
  for (int j = 0; j < N / 32; ++j) d[j] = rstd * (d[j] - m1 - xh[j] * m2);
}

// All 8 rows, d and xhat both in registers.
template <int N>
__device__ __forceinline__ void ln_relu_bwd(float (&d)[8][N / 32], const float (&xh)[8][N / 32], const float (&rstd)[8],
                                            const float* __restrict__ gamma, const float* __restrict__ beta, int lane) {
  float g[N / 32], b[N / 32];
  load_cols<N>(g, gamma, lane);
  load_cols<N>(b, beta, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) ln_relu_bwd_row<N>(d[i], xh[i], rstd[i], g, b);
}

template <int N>
__device__ __forceinline__ void sigmoid_rows(float (&acc)[8][N / 32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < N / 32; ++j) acc[i][j] = 1.f / (1.f + expf(-acc[i][j]));
}

// out[i] = dot(acc[i][:], vec) reduced over the warp's lanes.
template <int N>
__device__ __forceinline__ void dot_rows(float (&out)[8], const float (&acc)[8][N / 32],
                                         const float* __restrict__ vec, int lane) {
  float v[N / 32];
  load_cols<N>(v, vec, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < N / 32; ++j) s = fmaf(acc[i][j], v[j], s);
    out[i] = warp_sum(s);
  }
}

// Scatter-add rows to table[idx[i]][:].  SORTED: consecutive equal indices are pre-summed in registers
// (CSR order makes the left index of a warp's 8 rows at most ~2 distinct values), one RED per run.
template <int N, bool SORTED>
__device__ __forceinline__ void scatter_add_rows(const float (&acc)[8][N / 32], float* __restrict__ table,
                                                 const int* idx_rows, int lane) {
  constexpr int CPL = N / 32;
  if constexpr (SORTED) {
    int cur = idx_rows[0];
    float s[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) s[j] = acc[0][j];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      const int n = idx_rows[i];
      if (n == cur) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) s[j] += acc[i][j];
      } else {
        if (cur >= 0) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) atomicAdd(table + (size_t)cur * N + col_of<N>(lane, j), s[j]);
        }
        cur = n;
#pragma unroll
        for (int j = 0; j < CPL; ++j) s[j] = acc[i][j];
      }
    }
    if (cur >= 0) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) atomicAdd(table + (size_t)cur * N + col_of<N>(lane, j), s[j]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = idx_rows[i];
      if (n >= 0) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) atomicAdd(table + (size_t)n * N + col_of<N>(lane, j), acc[i][j]);
      }
    }
  }
}

}  // namespace mdb
