// TEST: the staged searches of csrc/nn_upfront.cuh that no kernel of the library instantiates yet must still compile as
// DEVICE code for sm_100a (tests/test_host_fuzz.py runs `nvcc -c` on this file; their arithmetic is checked on the host).
#include "../../pose_estimation_b200/csrc/core_math.cuh"
#include "../../pose_estimation_b200/csrc/nn_upfront.cuh"
using namespace peb;
template <int RW>
__global__ void __launch_bounds__(128, 8) warm_bounded(GridView g, const float4* __restrict__ work, const float4* __restrict__ prev, int n, float lim, int* out, float* od) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = work[i], o = prev[i];
  const float mx = p.x - o.x, my = p.y - o.y, mz = p.z - o.z;
  NnBest b = grid_nn_bounded_upfront<RW>(g, p.x, p.y, p.z, warm_bound_d2(o.w, sqrtf(mx * mx + my * my + mz * mz), g.h), lim);
  out[i] = b.j; od[i] = b.d2;
}
template __global__ void warm_bounded<2>(GridView, const float4*, const float4*, int, float, int*, float*);
template __global__ void warm_bounded<3>(GridView, const float4*, const float4*, int, float, int*, float*);
