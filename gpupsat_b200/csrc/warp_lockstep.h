// Warp-lockstep vocabulary used by cdcl_warp.inl.
//
// The per-cube solver is written as warp-synchronous code: control flow is warp-uniform, and everything that
// differs per lane lives inside LANES { } phases, with ballots / shuffles between phases.  On the GPU (the product)
// these macros are the CUDA warp intrinsics and per-lane variables are registers.
//
// GPSAT_WARP_EMU is a TEST-ONLY build mode (tests/emu/, never part of libgpsat.so): the same source compiled by
// g++ with the 32 lanes of a phase run one after the other, so the kernel logic can be stepped and compared with
// the oracle on a machine without a GPU.  It is a debugging aid for the kernel source, not a CPU fallback: nothing
// in the product links or dispatches to it, and the C ABI fails with GPSAT_E_NO_DEVICE when there is no GPU.
//
// Phase discipline (what makes both readings equivalent): inside one LANES phase a lane never reads memory that
// another lane writes in that same phase, except through the commutative atomics below; phases that communicate
// through memory are separated by SYNCWARP().
#pragma once
#include <stdint.h>

#if defined(GPSAT_WARP_EMU)

#include <cstring>
#define GPSAT_DEV inline
#define GPSAT_DEV_OUTLINE inline
#define GPSAT_NOUNROLL
#define GPSAT_LANE_DECL
#define GPSAT_LANE_DECL_S
#define LANEVAR(T, name) T name[32]
#define LV(name) name[lane]
#define LANES for (int lane = 0; lane < 32; ++lane)
#define BALLOT(expr)                                      \
    ([&]() -> unsigned {                                  \
        unsigned m_ = 0;                                  \
        for (int lane = 0; lane < 32; ++lane)             \
            if (expr) m_ |= 1u << lane;                   \
        return m_;                                        \
    }())
#define SHFL(name, src) (name[(src)])
#define MATCH_ANY(name)                                   \
    ([&]() -> unsigned {                                  \
        unsigned m_ = 0;                                  \
        for (int j_ = 0; j_ < 32; ++j_)                   \
            if (name[j_] == name[lane]) m_ |= 1u << j_;   \
        return m_;                                        \
    }())
#define SETLANE(name, src, value) (name[(src)] = (value))
#define SYNCWARP() ((void)0)
#define LANE0 for (int lane = 0; lane < 1; ++lane)
#define LANE_SUM_I64(name)                                \
    ([&]() -> long long {                                 \
        long long s_ = 0;                                 \
        for (int l_ = 0; l_ < 32; ++l_) s_ += name[l_];   \
        return s_;                                        \
    }())
#define LANE_MAX_I64(name)                                \
    ([&]() -> long long {                                 \
        long long s_ = name[0];                           \
        for (int l_ = 1; l_ < 32; ++l_)                   \
            if (name[l_] > s_) s_ = name[l_];             \
        return s_;                                        \
    }())
#define LANE_SUM_U64(name)                                \
    ([&]() -> unsigned long long {                        \
        unsigned long long s_ = 0;                        \
        for (int l_ = 0; l_ < 32; ++l_) s_ += name[l_];   \
        return s_;                                        \
    }())
static inline void gpsat_atomic_or(uint32_t *p, uint32_t v) { *p |= v; }
static inline void gpsat_atomic_and(uint32_t *p, uint32_t v) { *p &= v; }
static inline int gpsat_atomic_add(int *p, int v) { int o = *p; *p += v; return o; }
static inline int gpsat_atomic_cas(int *p, int cmp, int v) { int o = *p; if (o == cmp) *p = v; return o; }
static inline void gpsat_threadfence() {}
static inline int gpsat_atomic_max(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline void gpsat_atomic_add_ll(long long *p, long long v) { *p += v; }
static inline int gpsat_ld_volatile(const int *p) { return *(const volatile int *)p; }
static inline int gpsat_ld_cg(const int *p) { return *p; }
static inline void gpsat_nanosleep(unsigned) {}
// mesh (several GPUs as one work pool): system-scope variants; the emulator runs one warp on one "GPU"
static inline int gpsat_atomic_cas_sys(int *p, int cmp, int v) { return gpsat_atomic_cas(p, cmp, v); }
static inline int gpsat_atomic_add_sys(int *p, int v) { return gpsat_atomic_add(p, v); }
static inline void gpsat_threadfence_sys() {}
static inline unsigned long long gpsat_ld_volatile64(const int *p) { unsigned long long v; std::memcpy(&v, p, 8); return v; }
static inline void gpsat_copy_cg(int *dst, const int *src, int n_words, int lane) { if (lane == 0 && dst && src) std::memcpy(dst, src, (size_t)n_words * 4); }
// emulated clock: one tick per query, so tests can make budgeted steps expire deterministically
static unsigned long long g_gpsat_emu_clock = 0;
static inline unsigned long long gpsat_now_ns() { return ++g_gpsat_emu_clock; }
static inline int gpsat_popc(unsigned m) { return __builtin_popcount(m); }
static inline int gpsat_ffs(unsigned m) { return __builtin_ffs((int)m); }
struct gpsat_int2 { int x, y; };
typedef gpsat_int2 gint2;
static inline gint2 gpsat_ld2(const gint2 *p) { return *p; }
static inline int gpsat_ld(const int *p) { return *p; }

#else  // ---- CUDA device build ----

#define GPSAT_DEV __device__ __forceinline__
#define GPSAT_DEV_OUTLINE __device__ __noinline__
// every warp of a block is at a different point of a large program: code size (instruction-cache footprint), not
// loop overhead, is what costs issue slots here, so loops are kept rolled
#define GPSAT_NOUNROLL _Pragma("unroll 1")
// the lane index is read ONCE per thread through a volatile asm (gpsat_read_lane) and carried in a register: the
// compiler otherwise re-reads %tid.x (S2R, short-scoreboard latency) wherever `lane` is needed — 3 % of the stall
// samples of the CDCL kernel (profiles/r01_cdcl_lines_h.txt, kernels.cu:38)
#define GPSAT_LANE_DECL const int lane = gpsat_lane_of(this);
#define GPSAT_LANE_DECL_S const int lane = S.lane_id;
__device__ __forceinline__ int gpsat_read_lane()
{
    int l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}
template <class T>
__device__ __forceinline__ int gpsat_lane_of(const T *s) { return s->lane_id; }
#define LANEVAR(T, name) T name
#define LV(name) name
#define LANES
#define BALLOT(expr) __ballot_sync(0xffffffffu, (expr))
#define SHFL(name, src) __shfl_sync(0xffffffffu, name, (src))
#define MATCH_ANY(name) __match_any_sync(0xffffffffu, name)
#define SETLANE(name, src, value) \
    do {                          \
        if (lane == (src)) name = (value); \
    } while (0)
#define SYNCWARP() __syncwarp()
#define LANE0 if (lane == 0)
__device__ __forceinline__ long long gpsat_warp_sum_i64(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long gpsat_warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long gpsat_warp_max_i64(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}
#define LANE_SUM_I64(name) gpsat_warp_sum_i64((long long)(name))
#define LANE_SUM_U64(name) gpsat_warp_sum_u64((unsigned long long)(name))
#define LANE_MAX_I64(name) gpsat_warp_max_i64((long long)(name))
__device__ __forceinline__ void gpsat_atomic_or(uint32_t *p, uint32_t v) { atomicOr(p, v); }
__device__ __forceinline__ void gpsat_atomic_and(uint32_t *p, uint32_t v) { atomicAnd(p, v); }
__device__ __forceinline__ int gpsat_atomic_add(int *p, int v) { return atomicAdd(p, v); }
__device__ __forceinline__ int gpsat_atomic_cas(int *p, int cmp, int v) { return atomicCAS(p, cmp, v); }
__device__ __forceinline__ void gpsat_threadfence() { __threadfence(); }
__device__ __forceinline__ int gpsat_atomic_max(int *p, int v) { return atomicMax(p, v); }
__device__ __forceinline__ void gpsat_atomic_add_ll(long long *p, long long v)
{
    atomicAdd((unsigned long long *)p, (unsigned long long)v);
}
__device__ __forceinline__ int gpsat_ld_volatile(const int *p) { return *(const volatile int *)p; }
// L2-coherent load for data another SM wrote during this launch (queue slots, hand-off blocks, pool slots): a line
// cached in this SM's L1 before the writer published could otherwise be served stale
__device__ __forceinline__ int gpsat_ld_cg(const int *p) { return __ldcg(p); }
__device__ __forceinline__ void gpsat_nanosleep(unsigned ns) { __nanosleep(ns); }
// mesh (the GPUs of one box as one work pool over NVLink peer memory): words that another GPU reads or updates are
// accessed with system-scope atomics / fences; peer memory is never cached in this GPU's L2, .cg keeps it out of L1
__device__ __forceinline__ int gpsat_atomic_cas_sys(int *p, int cmp, int v) { return atomicCAS_system(p, cmp, v); }
__device__ __forceinline__ int gpsat_atomic_add_sys(int *p, int v) { return atomicAdd_system(p, v); }
__device__ __forceinline__ void gpsat_threadfence_sys() { __threadfence_system(); }
__device__ __forceinline__ unsigned long long gpsat_ld_volatile64(const int *p)
{
    return *(const volatile unsigned long long *)p;   // 8-byte aligned: one atomic snapshot of two adjacent words
}
// warp-cooperative copy of n_words (both sides 16-byte aligned, n_words rounded up to a multiple of 4 by the caller's
// layout) with L2-coherent 128-bit loads: four loads in flight per lane
__device__ __forceinline__ void gpsat_copy_cg(int *dst, const int *src, int n_words, int lane)
{
    const int n4 = (n_words + 3) >> 2;
    const int4 *s4 = reinterpret_cast<const int4 *>(src);
    int4 *d4 = reinterpret_cast<int4 *>(dst);
    int i = lane;
    for (; i + 96 < n4; i += 128) {
        const int4 a = __ldcg(s4 + i), b = __ldcg(s4 + i + 32), c = __ldcg(s4 + i + 64), d = __ldcg(s4 + i + 96);
        d4[i] = a;
        d4[i + 32] = b;
        d4[i + 64] = c;
        d4[i + 96] = d;
    }
    for (; i < n4; i += 32) d4[i] = __ldcg(s4 + i);
}
__device__ __forceinline__ unsigned long long gpsat_now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int gpsat_popc(unsigned m) { return __popc(m); }
__device__ __forceinline__ int gpsat_ffs(unsigned m) { return __ffs((int)m); }
typedef int2 gint2;
// read-only formula index: plain loads, so the same code reads it from shared memory (staged once per block when it
// fits) or from global memory through L1/L2
__device__ __forceinline__ gint2 gpsat_ld2(const gint2 *p) { return *p; }
__device__ __forceinline__ int gpsat_ld(const int *p) { return *p; }

#endif

#define GPSAT_LANEMASK_LT ((1u << lane) - 1u)
// the clause walk of propagate(): rolled by default (code size); -DGPSAT_HOTLOOP_UNROLL=n to measure unrolling
#if defined(GPSAT_WARP_EMU) || !defined(GPSAT_HOTLOOP_UNROLL)
#define GPSAT_HOTLOOP GPSAT_NOUNROLL
#else
#define GPSAT_HOTLOOP_STR2(x) #x
#define GPSAT_HOTLOOP_STR(x) GPSAT_HOTLOOP_STR2(unroll x)
#define GPSAT_HOTLOOP _Pragma(GPSAT_HOTLOOP_STR(GPSAT_HOTLOOP_UNROLL))
#endif
