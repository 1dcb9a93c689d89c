// simt.cuh -- the handful of warp / memory primitives the kernels use, behind one spelling.
//
// Device build (nvcc, sm_100a): each lab_* is the CUDA intrinsic.
// Host build (tests/simt/): tests/simt/simt_host.hh is included INSTEAD of this file's device
// half and implements the same names on a 32-fibre lock-step warp emulator, so the kernel
// bodies in engine.cuh can be executed and checked against the oracle on a machine without a
// GPU.  That emulator is test infrastructure; the product only ever runs the device half.
#pragma once
#include <cstdint>

#if defined(LAB_SIMT_HOST)
#include "simt_host.hh"
#else

#define LAB_DEV __device__ __forceinline__
#define LAB_FULL 0xFFFFFFFFu

LAB_DEV uint32_t lab_shfl_up(uint32_t v, int d) { return __shfl_up_sync(LAB_FULL, v, d); }
LAB_DEV uint32_t lab_shfl_down(uint32_t v, int d) { return __shfl_down_sync(LAB_FULL, v, d); }
LAB_DEV uint32_t lab_shfl(uint32_t v, int src) { return __shfl_sync(LAB_FULL, v, src); }
LAB_DEV uint32_t lab_ballot(bool p) { return __ballot_sync(LAB_FULL, p); }
LAB_DEV bool lab_any(bool p) { return __any_sync(LAB_FULL, p); }
LAB_DEV uint32_t lab_reduce_min(uint32_t v) { return __reduce_min_sync(LAB_FULL, v); }
LAB_DEV uint32_t lab_reduce_max(uint32_t v) { return __reduce_max_sync(LAB_FULL, v); }
LAB_DEV uint32_t lab_reduce_or(uint32_t v) { return __reduce_or_sync(LAB_FULL, v); }
LAB_DEV uint32_t lab_reduce_add(uint32_t v) { return __reduce_add_sync(LAB_FULL, v); }
LAB_DEV void lab_syncwarp() { __syncwarp(); }
LAB_DEV uint32_t lab_match_any(uint32_t v) { return __match_any_sync(LAB_FULL, v); } // lanes holding the same value

LAB_DEV int lab_ffs64(uint64_t v) { return __ffsll(static_cast<long long>(v)); } // 1-based, 0 if none
LAB_DEV int lab_popc64(uint64_t v) { return __popcll(v); }
LAB_DEV int lab_clz64(uint64_t v) { return __clzll(static_cast<long long>(v)); } // 64 if none
LAB_DEV int lab_ffs32(uint32_t v) { return __ffs(static_cast<int>(v)); } // 1-based, 0 if none
LAB_DEV int lab_clz32(uint32_t v) { return __clz(static_cast<int>(v)); }  // 32 if none

// loads that must observe other SMs' stores (L1 is not coherent inside a persistent kernel)
LAB_DEV uint32_t lab_ld_cg(uint32_t const *p) { return __ldcg(p); }
LAB_DEV uint64_t lab_ld_cg(uint64_t const *p) {
    return static_cast<uint64_t>(__ldcg(reinterpret_cast<unsigned long long const *>(p)));
}
LAB_DEV uint32_t lab_ld_volatile(uint32_t const *p) { return *reinterpret_cast<uint32_t const volatile *>(p); }
LAB_DEV uint64_t lab_ld_volatile(uint64_t const *p) { return *reinterpret_cast<uint64_t const volatile *>(p); }
LAB_DEV uint16_t lab_ld_volatile(uint16_t const *p) { return *reinterpret_cast<uint16_t const volatile *>(p); }
// read-only data (path masks): non-coherent path is fine
LAB_DEV uint64_t lab_ld_ro(uint64_t const *p) {
    return static_cast<uint64_t>(__ldg(reinterpret_cast<unsigned long long const *>(p)));
}
LAB_DEV uint32_t lab_ld_ro(uint32_t const *p) { return __ldg(p); }
LAB_DEV uint16_t lab_ld_ro(uint16_t const *p) { return __ldg(p); }
LAB_DEV uint8_t lab_ld_ro(uint8_t const *p) { return __ldg(p); }
LAB_DEV uint8_t lab_ld_cg(uint8_t const *p) { return __ldcg(p); }
LAB_DEV uint16_t lab_ld_cg(uint16_t const *p) { return __ldcg(p); }
LAB_DEV void lab_st_cg(uint32_t *p, uint32_t v) { __stcg(p, v); }
LAB_DEV void lab_st_cg(uint64_t *p, uint64_t v) {
    __stcg(reinterpret_cast<unsigned long long *>(p), static_cast<unsigned long long>(v));
}
// 16 bytes at a time (p is 16-byte aligned), streaming: read once / written once, nothing to keep in the caches
struct lab_u32x4 {
    uint32_t x, y, z, w;
};
LAB_DEV lab_u32x4 lab_ld_stream_v4(void const *p) {
    uint4 const t = __ldcs(reinterpret_cast<uint4 const *>(p));
    return {t.x, t.y, t.z, t.w};
}
LAB_DEV lab_u32x4 lab_ld_ro_v4(void const *p) { // read-only data
    uint4 const t = __ldg(reinterpret_cast<uint4 const *>(p));
    return {t.x, t.y, t.z, t.w};
}
LAB_DEV void lab_st_v4(void *p, lab_u32x4 v) { *reinterpret_cast<uint4 *>(p) = make_uint4(v.x, v.y, v.z, v.w); }
LAB_DEV void lab_st_stream_v4(void *p, lab_u32x4 v) { __stcs(reinterpret_cast<uint4 *>(p), make_uint4(v.x, v.y, v.z, v.w)); }
// a (pointer, value) pair of the pointer-jumping lists: ONE 8-byte access (a random 4-byte access costs a 32-byte sector
// either way, so two separate arrays would move twice the sectors)
struct lab_u32x2 {
    uint32_t x, y;
};
LAB_DEV lab_u32x2 lab_ld_cg_x2(uint32_t const *base, uint32_t i) {
    uint2 const t = __ldcg(reinterpret_cast<uint2 const *>(base) + i);
    return {t.x, t.y};
}
LAB_DEV void lab_st_x2(uint32_t *base, uint32_t i, uint32_t x, uint32_t y) { reinterpret_cast<uint2 *>(base)[i] = make_uint2(x, y); }
LAB_DEV void lab_st_pair(uint32_t *p, uint32_t x, uint32_t y) { *reinterpret_cast<uint2 *>(p) = make_uint2(x, y); } // p on an 8-byte boundary
LAB_DEV void lab_st_volatile(uint32_t *p, uint32_t v) { *reinterpret_cast<uint32_t volatile *>(p) = v; }
LAB_DEV void lab_st_volatile(uint16_t *p, uint16_t v) { *reinterpret_cast<uint16_t volatile *>(p) = v; }

LAB_DEV uint32_t lab_atomic_add(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
LAB_DEV uint64_t lab_atomic_add(uint64_t *p, uint64_t v) {
    return static_cast<uint64_t>(atomicAdd(reinterpret_cast<unsigned long long *>(p), static_cast<unsigned long long>(v)));
}
LAB_DEV uint32_t lab_atomic_sub(uint32_t *p, uint32_t v) { return atomicSub(p, v); }
LAB_DEV uint32_t lab_atomic_or(uint32_t *p, uint32_t v) { return atomicOr(p, v); }
LAB_DEV uint64_t lab_atomic_or(uint64_t *p, uint64_t v) {
    return static_cast<uint64_t>(atomicOr(reinterpret_cast<unsigned long long *>(p), static_cast<unsigned long long>(v)));
}
LAB_DEV uint32_t lab_atomic_and(uint32_t *p, uint32_t v) { return atomicAnd(p, v); }
LAB_DEV uint32_t lab_atomic_min(uint32_t *p, uint32_t v) { return atomicMin(p, v); }
LAB_DEV uint32_t lab_atomic_max(uint32_t *p, uint32_t v) { return atomicMax(p, v); }
LAB_DEV uint32_t lab_atomic_cas(uint32_t *p, uint32_t expect, uint32_t v) { return atomicCAS(p, expect, v); }
LAB_DEV uint32_t lab_atomic_exch(uint32_t *p, uint32_t v) { return atomicExch(p, v); }

LAB_DEV void lab_threadfence() { __threadfence(); }
LAB_DEV void lab_nanosleep(unsigned ns) { __nanosleep(ns); }
LAB_DEV uint64_t lab_cycles() { return static_cast<uint64_t>(clock64()); }
LAB_DEV uint64_t lab_clock_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

#endif // LAB_SIMT_HOST
