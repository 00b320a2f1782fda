// brute.cu — the brute-force FP32 1-NN validator (north star (2)): no grid, every query against
// every target point, target tiles staged in shared memory.  Same float L2_Simple distance
// ([FLANN] algorithms/dist.h) and the same tie rule (lowest original index) as the grid search,
// so the two must agree bit for bit; this is the on-device ground truth of the parity tests.
// FP32-pipe bound: n_q * n_t * (3 FADD + 3 FMUL + 2 FADD) with no tensor cores, because the
// distance test has to be exact.
#include <algorithm>

#include "nn_search.cuh"

namespace peb {

namespace {

constexpr int kBruteThreads = 256;
constexpr int kBruteTile = 2048;        // target points per shared-memory tile (32 KB)
constexpr int kBruteQPerThread = 4;     // queries per thread: four independent min chains per LDS.128
constexpr int kBruteChunk = 8;          // points whose distances are reduced with FMNMX before the (rare) index update

// Grid = (query blocks, target segments): at the 50 k queries of configs[1] the query blocks alone would fill a
// third of the 148 SMs, so the target is cut into segments and the per-segment winners are merged with a 64-bit
// atomicMin on (distance bits << 32 | index) — non-negative floats order like their bit patterns, so the minimum of
// the keys is the smallest distance and, among equal distances, the lowest index: the tie rule of the grid search.
// Inner loop per (query, target) pair: 3 FADD + 3 FMUL + 2 FADD (l2_simple, unfused) + 1 FMNMX; the index is only
// looked for when a chunk of 8 points improves the query's best distance.
__global__ void __launch_bounds__(kBruteThreads) nn_bruteforce_kernel(const float4* __restrict__ tgt, int n_tgt, int seg_len,
                                                                      const float4* __restrict__ q, int nq,
                                                                      unsigned long long* __restrict__ keys) {
  __shared__ float4 tile[kBruteTile];
  const int q0 = (blockIdx.x * kBruteThreads + threadIdx.x) * kBruteQPerThread;
  const int seg0 = blockIdx.y * seg_len, seg1 = min(seg0 + seg_len, n_tgt);
  float qx[kBruteQPerThread], qy[kBruteQPerThread], qz[kBruteQPerThread], bd[kBruteQPerThread];
  int bi[kBruteQPerThread];
#pragma unroll
  for (int k = 0; k < kBruteQPerThread; ++k) {
    const float4 p = (q0 + k < nq) ? q[q0 + k] : make_float4(0.f, 0.f, 0.f, 0.f);
    qx[k] = p.x;
    qy[k] = p.y;
    qz[k] = p.z;
    bd[k] = pos_inf();
    bi[k] = -1;
  }
  const float qnan = __int_as_float(0x7fc00000);
  for (int base = seg0; base < seg1; base += kBruteTile) {
    const int cnt = min(kBruteTile, seg1 - base);
    const int cnt_pad = (cnt + kBruteChunk - 1) / kBruteChunk * kBruteChunk;
    __syncthreads();
    for (int j = threadIdx.x; j < cnt_pad; j += kBruteThreads) {
      float4 t = make_float4(qnan, qnan, qnan, 0.f);
      if (j < cnt) {
        t = tgt[base + j];
        // non-finite targets are not part of the index (KdTreeFLANN leaves them out): park them (and the padding) at
        // NaN so that every comparison against them fails
        if (!finite3(t.x, t.y, t.z)) t.x = t.y = t.z = qnan;
      }
      tile[j] = t;
    }
    __syncthreads();
    for (int j = 0; j < cnt_pad; j += kBruteChunk) {
      float d[kBruteQPerThread][kBruteChunk];
#pragma unroll
      for (int c = 0; c < kBruteChunk; ++c) {
        const float4 t = tile[j + c];
#pragma unroll
        for (int k = 0; k < kBruteQPerThread; ++k) d[k][c] = l2_simple(qx[k], qy[k], qz[k], t.x, t.y, t.z);
      }
#pragma unroll
      for (int k = 0; k < kBruteQPerThread; ++k) {
        float m = d[k][0];  // (fminf drops NaN operands)
#pragma unroll
        for (int c = 1; c < kBruteChunk; ++c) m = fminf(m, d[k][c]);
        if (m < bd[k]) {  // ascending index order and strict <: the lowest index wins exact ties
          bd[k] = m;
          int first = kBruteChunk - 1;
#pragma unroll
          for (int c = kBruteChunk - 2; c >= 0; --c) first = (d[k][c] == m) ? c : first;
          bi[k] = base + j + first;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kBruteQPerThread; ++k) {
    if (q0 + k < nq && bi[k] >= 0 && finite3(qx[k], qy[k], qz[k]))
      atomicMin(keys + q0 + k, (static_cast<unsigned long long>(__float_as_uint(bd[k])) << 32) | static_cast<unsigned>(bi[k]));
  }
}

__global__ void nn_bruteforce_unpack_kernel(const unsigned long long* __restrict__ keys, int nq, int32_t* __restrict__ out_idx,
                                            float* __restrict__ out_d2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const unsigned long long k = keys[i];
  const bool ok = k != ~0ull;
  out_idx[i] = ok ? static_cast<int32_t>(k & 0xFFFFFFFFull) : -1;
  out_d2[i] = ok ? __uint_as_float(static_cast<unsigned>(k >> 32)) : pos_inf();
}

}  // namespace

int nn_bruteforce_device(peb_ctx* ctx, const float4* d_q, int nq, int32_t* d_idx, float* d_d2) {
  if (!ctx->tgt_grid.valid) return fail(ctx, PEB_E_NO_TARGET, "nn_search_bruteforce: no target set");
  if (nq == 0) return PEB_OK;
  const int n_tgt = static_cast<int>(ctx->n_tgt);
  const int qblocks = ceil_div(nq, kBruteThreads * kBruteQPerThread);
  // two blocks per SM (measured: four are slower, 3.60 against 3.43 ms at 50 k x 203 k); segments are whole tiles
  int segs = std::max(1, std::min((2 * kSmCount + qblocks / 2) / qblocks, ceil_div(n_tgt, kBruteTile)));
  const int seg_len = ceil_div(ceil_div(n_tgt, segs), kBruteTile) * kBruteTile;
  segs = std::max(1, ceil_div(n_tgt, seg_len));
  PEB_CUDA(ctx, ctx->brute_keys.ensure(static_cast<size_t>(nq) * 8));
  unsigned long long* keys = ctx->brute_keys.as<unsigned long long>();
  PEB_CUDA(ctx, cudaMemsetAsync(keys, 0xFF, static_cast<size_t>(nq) * 8, ctx->stream));
  PEB_LAUNCH(ctx, nn_bruteforce_kernel, dim3(qblocks, segs), kBruteThreads, 0, ctx->tgt_raw.as<float4>(), n_tgt, seg_len,
             d_q, nq, keys);
  PEB_LAUNCH(ctx, nn_bruteforce_unpack_kernel, ceil_div(nq, 256), 256, 0, keys, nq, d_idx, d_d2);
  return PEB_OK;
}

}  // namespace peb
