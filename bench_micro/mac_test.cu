// SASS experiment: 62x62-bit multiply-accumulate into a wide lazy accumulator with even/odd column chains
#include <cstdint>
typedef unsigned long long u64;
typedef unsigned int u32;

struct AccEO {
  u32 e0, e1, e2, e3, e4, o1, o2, o3;
  __device__ __forceinline__ void clear() { e0 = e1 = e2 = e3 = e4 = o1 = o2 = o3 = 0; }
  __device__ __forceinline__ void mac(u64 a, u64 b) {
    asm("{\n\t"
        ".reg .u32 a0, a1, b0, b1;\n\t"
        "mov.b64 {a0, a1}, %8;\n\t"
        "mov.b64 {b0, b1}, %9;\n\t"
        "mad.lo.cc.u32 %0, a0, b0, %0;\n\t"
        "madc.hi.cc.u32 %1, a0, b0, %1;\n\t"
        "madc.lo.cc.u32 %2, a1, b1, %2;\n\t"
        "madc.hi.cc.u32 %3, a1, b1, %3;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %5, a0, b1, %5;\n\t"
        "madc.hi.cc.u32 %6, a0, b1, %6;\n\t"
        "addc.u32 %7, %7, 0;\n\t"
        "mad.lo.cc.u32 %5, a1, b0, %5;\n\t"
        "madc.hi.cc.u32 %6, a1, b0, %6;\n\t"
        "addc.u32 %7, %7, 0;\n\t"
        "}"
        : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4), "+r"(o1), "+r"(o2), "+r"(o3)
        : "l"(a), "l"(b));
  }
};

__global__ void k(const u64* __restrict__ x, const u64* __restrict__ w, u32* out, int n) {
  AccEO acc[4];
  for (int k = 0; k < 4; k++) acc[k].clear();
  u64 r = x[threadIdx.x];
#pragma unroll 2
  for (int i = 0; i < n; i++) {
    u64 rr = x[threadIdx.x + i * 128];
    acc[0].mac(rr, w[i * 4]);
    acc[1].mac(rr, w[i * 4 + 1]);
    acc[2].mac(rr, w[i * 4 + 2]);
    acc[3].mac(rr, w[i * 4 + 3]);
  }
  for (int k = 0; k < 4; k++) {
    u32* o = out + (threadIdx.x * 4 + k) * 8;
    o[0] = acc[k].e0; o[1] = acc[k].e1; o[2] = acc[k].e2; o[3] = acc[k].e3; o[4] = acc[k].e4;
    o[5] = acc[k].o1; o[6] = acc[k].o2; o[7] = acc[k].o3;
  }
}
