// 32-bit limb primitives with carry: PTX add.cc / mad.lo.cc / madc.hi.cc ... on the device.
//
// ptxas fuses each adjacent (mad[c].lo.cc , madc.hi.cc) pair on the same operands into ONE
// IMAD.WIDE.U32[.X] (32x32+64 -> 64 with carry-in/out in a predicate), which is what makes the
// Montgomery kernels in fp.cuh run on IMAD carry chains (check: cuobjdump -sass | grep IMAD.WIDE).
//
// When compiled for the host (unit tests of the limb schedule, and the small amount of host-side
// curve arithmetic the library does at the end of an MSM) the same names are backed by a
// software carry flag, so the exact instruction schedule can be exercised without a GPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SWB_HD __host__ __device__ __forceinline__
#define SWB_D __device__ __forceinline__
#else
#define SWB_HD inline
#define SWB_D inline
#endif

namespace swb {
namespace ptx {

#if defined(__CUDA_ARCH__)

SWB_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SWB_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SWB_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SWB_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SWB_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SWB_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SWB_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SWB_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SWB_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SWB_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SWB_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SWB_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SWB_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

#else  // host emulation, one carry flag per thread

inline uint32_t& cc_flag() { static thread_local uint32_t cc = 0; return cc; }
inline uint32_t add3_(uint32_t a, uint32_t b, uint32_t cin, bool set) {
    uint64_t t = (uint64_t)a + b + cin;
    if (set) cc_flag() = (uint32_t)(t >> 32);
    return (uint32_t)t;
}
inline uint32_t sub3_(uint32_t a, uint32_t b, uint32_t bin, bool set) {
    uint64_t t = (uint64_t)a - b - bin;
    if (set) cc_flag() = (uint32_t)((t >> 32) & 1);   // borrow kept in the same flag, as PTX does
    return (uint32_t)t;
}
inline uint32_t add_cc(uint32_t a, uint32_t b) { return add3_(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return add3_(a, b, cc_flag(), true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return add3_(a, b, cc_flag(), false); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return sub3_(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return sub3_(a, b, cc_flag(), true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return sub3_(a, b, cc_flag(), false); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3_(mul_lo(a, b), c, 0, true); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3_(mul_lo(a, b), c, cc_flag(), true); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3_(mul_hi(a, b), c, 0, true); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3_(mul_hi(a, b), c, cc_flag(), true); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return add3_(mul_hi(a, b), c, cc_flag(), false); }

#endif

}  // namespace ptx
}  // namespace swb
