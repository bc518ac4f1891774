// seqsum.cuh -- primitives of the exact PARALLEL emulation of the sequential f32 accumulation
//   c_i = fl32(c_{i-1} + w_i), w_i >= 0   (pu:430 numba np.sum, pu:436-443 the running sum of the walk).
// While c stays inside one binade [2^e, 2^(e+1)) it is an integer multiple K of u = 2^(e-23) and adding w is the
// integer map K -> K + A(K & 1) (the parity only matters for exact ties, round-to-even); such maps compose
// associatively.  Shared by resample.cu (stand-alone kernels) and tail.cu (persistent step tail).
#pragma once
#include "common.cuh"

// increment if the incoming K is even / odd.  Increments saturate at SEQ_SAT (>= 2^25 > any in-binade K):
// a saturated value means "the sum has left the binade", which is absorbing, so the composition stays
// associative while everything fits in 32 bits (half the shuffles and registers of a 64-bit scan).
#define SEQ_SAT 0x40000000u
struct Pair64 { unsigned a0, a1; };

__device__ __forceinline__ unsigned sat_add(unsigned a, unsigned b) { return min(a + b, SEQ_SAT); }   // a, b <= SEQ_SAT
__device__ __forceinline__ Pair64 pair_compose(const Pair64 &f1, const Pair64 &f2) {   // first f1, then f2
    Pair64 r;
    r.a0 = sat_add(f1.a0, (f1.a0 & 1u) ? f2.a1 : f2.a0);
    r.a1 = sat_add(f1.a1, ((1u + f1.a1) & 1u) ? f2.a1 : f2.a0);
    return r;
}
__device__ __forceinline__ Pair64 pair_shfl_up(const Pair64 &v, int o) {
    Pair64 r;
    r.a0 = __shfl_up_sync(0xffffffffu, v.a0, o);
    r.a1 = __shfl_up_sync(0xffffffffu, v.a1, o);
    return r;
}
// element map for weight w when the running sum has unit exponent e (u = 2^(e-23), e >= -126)
__device__ __forceinline__ Pair64 seq_decode(float w, int e) {
    const unsigned b = __float_as_uint(w);
    const unsigned ef = (b >> 23) & 0xffu, mf = b & 0x7fffffu;
    Pair64 r; r.a0 = 0; r.a1 = 0;
    if ((b & 0x7fffffffu) == 0u) return r;
    const long long M = ef ? (long long)(mf | 0x800000u) : (long long)mf;
    const int Ew = ef ? (int)ef - 127 : -126;
    const int sh = e - Ew;
    if (sh <= 0) {
        const unsigned a = (-sh > 6) ? SEQ_SAT : (unsigned)min((long long)SEQ_SAT, M << (-sh));
        r.a0 = a; r.a1 = a;
    } else if (sh <= 24) {
        const unsigned a = (unsigned)(M >> sh), rem = (unsigned)(M & ((1ll << sh) - 1)), half = 1u << (sh - 1);
        if (rem < half) { r.a0 = a; r.a1 = a; }
        else if (rem > half) { r.a0 = a + 1; r.a1 = a + 1; }
        else { r.a0 = a + (a & 1u); r.a1 = a + ((1u + a) & 1u); }
    }   // sh >= 25: w < u/2, the sum does not move
    return r;
}
__device__ __forceinline__ int seq_exponent(float c) {      // unit exponent of c (denormals share e = -126)
    const unsigned ef = (__float_as_uint(c) >> 23) & 0xffu;
    return ef ? (int)ef - 127 : -126;
}
__device__ __forceinline__ long long seq_K(float c) {        // c = K * 2^(e-23)
    const unsigned b = __float_as_uint(c);
    const unsigned ef = (b >> 23) & 0xffu, mf = b & 0x7fffffu;
    return ef ? (long long)(mf | 0x800000u) : (long long)mf;
}
__device__ __forceinline__ float seq_value(long long K, int e) {   // K < 2^24
    return __uint_as_float((unsigned)(((long long)(e + 126) << 23) + K));
}
