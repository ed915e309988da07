#include "../fhe_rs_b200/csrc/ntt.cuh"
using namespace fhe_b200;
__global__ void kbf(u64* d, const ulonglong2* tw, u64 p) {
  u64 x = d[threadIdx.x], y = d[threadIdx.x + 512];
  ulonglong2 w = tw[blockIdx.x];
  bf_fwd<false>(x, y, w.x, w.y, p, 2 * p, 0);
  d[threadIdx.x] = x; d[threadIdx.x + 512] = y;
}
__global__ void kbi(u64* d, const ulonglong2* tw, u64 p) {
  u64 x = d[threadIdx.x], y = d[threadIdx.x + 512];
  ulonglong2 w = tw[blockIdx.x];
  bf_inv<false>(x, y, w.x, w.y, p, 2 * p, 0);
  d[threadIdx.x] = x; d[threadIdx.x + 512] = y;
}
