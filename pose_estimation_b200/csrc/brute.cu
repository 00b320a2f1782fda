// brute.cu — the brute-force FP32 1-NN validator (north star (2)): no grid, every query against
// every target point, target tiles staged in shared memory.  Same float L2_Simple distance
// ([FLANN] algorithms/dist.h) and the same tie rule (lowest original index) as the grid search,
// so the two must agree bit for bit; this is the on-device ground truth of the parity tests.
// FP32-pipe bound: n_q * n_t * (3 FADD + 3 FMUL + 2 FADD) with no tensor cores, because the
// distance test has to be exact.
#include "nn_search.cuh"

namespace peb {

namespace {

constexpr int kBruteThreads = 256;
constexpr int kBruteTile = 2048;        // target points per shared-memory tile (32 KB)
constexpr int kBruteQPerThread = 2;     // queries per thread: two independent min chains

__global__ void __launch_bounds__(kBruteThreads) nn_bruteforce_kernel(const float4* __restrict__ tgt, int n_tgt,
                                                                      const float4* __restrict__ q, int nq,
                                                                      int32_t* __restrict__ out_idx,
                                                                      float* __restrict__ out_d2) {
  __shared__ float4 tile[kBruteTile];
  const int q0 = (blockIdx.x * kBruteThreads + threadIdx.x) * kBruteQPerThread;
  float qx[kBruteQPerThread], qy[kBruteQPerThread], qz[kBruteQPerThread], bd[kBruteQPerThread];
  int bi[kBruteQPerThread];
#pragma unroll
  for (int k = 0; k < kBruteQPerThread; ++k) {
    float4 p = (q0 + k < nq) ? q[q0 + k] : make_float4(0.f, 0.f, 0.f, 0.f);
    qx[k] = p.x;
    qy[k] = p.y;
    qz[k] = p.z;
    bd[k] = pos_inf();
    bi[k] = -1;
  }
  for (int base = 0; base < n_tgt; base += kBruteTile) {
    const int cnt = min(kBruteTile, n_tgt - base);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += kBruteThreads) {
      float4 t = tgt[base + j];
      // non-finite targets are not part of the index (KdTreeFLANN leaves them out): park them at
      // NaN so that every comparison against them fails
      if (!finite3(t.x, t.y, t.z)) t.x = t.y = t.z = __int_as_float(0x7fc00000);
      tile[j] = t;
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const float4 t = tile[j];
#pragma unroll
      for (int k = 0; k < kBruteQPerThread; ++k) {
        const float d2 = l2_simple(qx[k], qy[k], qz[k], t.x, t.y, t.z);
        if (d2 < bd[k]) {  // ascending index order => the lowest index wins exact ties
          bd[k] = d2;
          bi[k] = base + j;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kBruteQPerThread; ++k) {
    if (q0 + k < nq) {
      const bool ok = finite3(qx[k], qy[k], qz[k]);
      out_idx[q0 + k] = ok ? bi[k] : -1;
      out_d2[q0 + k] = ok ? bd[k] : pos_inf();
    }
  }
}

}  // namespace

int nn_bruteforce_device(peb_ctx* ctx, const float4* d_q, int nq, int32_t* d_idx, float* d_d2) {
  if (!ctx->tgt_grid.valid) return fail(ctx, PEB_E_NO_TARGET, "nn_search_bruteforce: no target set");
  if (nq == 0) return PEB_OK;
  const int blocks = ceil_div(nq, kBruteThreads * kBruteQPerThread);
  PEB_LAUNCH(ctx, nn_bruteforce_kernel, blocks, kBruteThreads, 0, ctx->tgt_raw.as<float4>(),
             static_cast<int>(ctx->n_tgt), d_q, nq, d_idx, d_d2);
  return PEB_OK;
}

}  // namespace peb
