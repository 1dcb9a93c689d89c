// simt_host.hh -- TEST INFRASTRUCTURE: a lock-step warp emulator for machines without a GPU.
//
// csrc/device/simt.cuh includes this file instead of its CUDA half when LAB_SIMT_HOST is defined,
// so the kernel bodies in engine.cuh / post.cuh compile unchanged as ordinary C++.  Each lane is
// a ucontext fibre; a warp-wide primitive (shuffle, ballot, reduce, syncwarp) parks the calling
// lane until all 32 lanes of the warp have reached it, exactly as *_sync intrinsics with a full
// mask do.  Several warps are interleaved (optionally in a seeded random order) on ONE OS thread,
// so global atomics are trivially atomic and every interleaving at primitive / nanosleep
// granularity is reproducible.  What it does not model: the L1's incoherence, memory reordering,
// performance.  It exists to find logic errors before GPU minutes are spent.
#pragma once
#include <ucontext.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <random>
#include <type_traits>
#include <vector>

#define LAB_DEV inline
#define LAB_FULL 0xFFFFFFFFu

namespace simt {

enum Lane_state { RUNNABLE, AT_BARRIER, FINISHED };

struct Warp;
struct Lane {
    ucontext_t ctx;
    std::unique_ptr<char[]> stack;
    Lane_state state = RUNNABLE;
    unsigned pc = 0; // warp primitives executed so far (selects the exchange buffer)
    Warp *warp = nullptr;
    int id = 0;
};

struct Warp {
    Lane lanes[32];
    uint64_t xchg[2][32];
    int site[32]; // source line of the primitive each lane is waiting in
    std::function<void(int)> body;
    int finished = 0;
};

inline ucontext_t g_sched;
inline Lane *g_cur = nullptr;
inline constexpr size_t STACK = 256 * 1024;

inline void lane_entry(unsigned lo, unsigned hi) {
    Lane *l = reinterpret_cast<Lane *>((static_cast<uintptr_t>(hi) << 32) | lo);
    l->warp->body(l->id);
    l->state = FINISHED;
    l->warp->finished++;
    swapcontext(&l->ctx, &g_sched);
}

inline void park(Lane_state s) {
    Lane *l = g_cur;
    l->state = s;
    swapcontext(&l->ctx, &g_sched);
}

// all 32 lanes exchange one value; returns the buffer every lane can read after the barrier
inline int g_site = 0; // set by the lab_* macros below just before the primitive runs
inline uint64_t const *exchange(uint64_t v) {
    Lane *l = g_cur;
    unsigned const b = l->pc & 1u;
    l->pc++;
    l->warp->xchg[b][l->id] = v;
    l->warp->site[l->id] = g_site;
    park(AT_BARRIER);
    return l->warp->xchg[b];
}

// seeded chaos: loads that observe other warps' stores sometimes let other warps run first, so
// that two lanes of one warp can see different values -- as on the real machine
inline std::mt19937 g_chaos(12345);
inline unsigned g_chaos_on = 0;
inline void maybe_yield() {
    if (g_chaos_on && (g_chaos() & 7u) == 0u) {
        park(RUNNABLE);
    }
}

// run one scheduling round of a warp: every runnable lane runs until it parks
inline void run_round(Warp &w) {
    for (int i = 0; i < 32; i++) {
        Lane &l = w.lanes[i];
        if (l.state == RUNNABLE) {
            g_cur = &l;
            swapcontext(&g_sched, &l.ctx);
        }
    }
    int at = 0;
    for (auto &l : w.lanes) {
        at += l.state == AT_BARRIER;
    }
    if (at == 32) {
        for (int i = 1; i < 32; i++) {
            if (w.site[i] != w.site[0]) {
                std::fprintf(stderr, "simt: lanes 0 and %d of one warp sit in DIFFERENT primitives (lines %d and %d)\n", i,
                             w.site[0], w.site[i]);
                std::abort();
            }
        }
        for (auto &l : w.lanes) {
            l.state = RUNNABLE;
        }
    } else if (at != 0 && at + w.finished == 32) {
        std::fprintf(stderr, "simt: divergent warp primitive (%d lanes waiting, %d finished)\n", at, w.finished);
        std::abort();
    }
}

// Launch `nwarps` warps of `body(warp, lane)` and interleave them until all have returned.
// sched_seed == 0: round-robin; otherwise warps are picked in a seeded random order.
inline void launch(int nwarps, std::function<void(int, int)> body, unsigned sched_seed = 0) {
    std::vector<std::unique_ptr<Warp>> warps;
    for (int w = 0; w < nwarps; w++) {
        warps.emplace_back(new Warp());
        Warp &W = *warps.back();
        W.body = [body, w](int lane) { body(w, lane); };
        for (int i = 0; i < 32; i++) {
            Lane &l = W.lanes[i];
            l.warp = &W;
            l.id = i;
            l.stack.reset(new char[STACK]);
            getcontext(&l.ctx);
            l.ctx.uc_stack.ss_sp = l.stack.get();
            l.ctx.uc_stack.ss_size = STACK;
            l.ctx.uc_link = nullptr;
            uintptr_t const p = reinterpret_cast<uintptr_t>(&l);
            makecontext(&l.ctx, reinterpret_cast<void (*)()>(lane_entry), 2, static_cast<unsigned>(p & 0xFFFFFFFFu),
                        static_cast<unsigned>(p >> 32));
        }
    }
    std::mt19937 rng(sched_seed);
    g_chaos.seed(sched_seed * 7919u + 1u);
    g_chaos_on = sched_seed;
    int live = nwarps;
    while (live > 0) {
        if (sched_seed == 0) {
            for (auto &W : warps) {
                if (W->finished < 32) run_round(*W);
            }
        } else {
            Warp &W = *warps[rng() % static_cast<unsigned>(nwarps)];
            if (W.finished < 32) {
                int const rounds = 1 + static_cast<int>(rng() % 8);
                for (int r = 0; r < rounds && W.finished < 32; r++) run_round(W);
            }
        }
        live = 0;
        for (auto &W : warps) {
            live += W->finished < 32;
        }
    }
    g_cur = nullptr;
    g_chaos_on = 0;
}

} // namespace simt

LAB_DEV int lab_lane() { return simt::g_cur->id; }
LAB_DEV uint32_t lab_shfl_up(uint32_t v, int d) {
    int const me = lab_lane();
    uint64_t const *b = simt::exchange(v);
    return static_cast<uint32_t>(me >= d ? b[me - d] : v);
}
LAB_DEV uint32_t lab_shfl_down(uint32_t v, int d) {
    int const me = lab_lane();
    uint64_t const *b = simt::exchange(v);
    return static_cast<uint32_t>(me + d < 32 ? b[me + d] : v);
}
LAB_DEV uint32_t lab_shfl(uint32_t v, int src) {
    uint64_t const *b = simt::exchange(v);
    return static_cast<uint32_t>(b[src & 31]);
}
LAB_DEV uint32_t lab_ballot(bool p) {
    uint64_t const *b = simt::exchange(p ? 1 : 0);
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) m |= static_cast<uint32_t>(b[i] & 1) << i;
    return m;
}
LAB_DEV bool lab_any(bool p) { return lab_ballot(p) != 0; }
LAB_DEV uint32_t lab_reduce_min(uint32_t v) {
    uint64_t const *b = simt::exchange(v);
    uint32_t m = 0xFFFFFFFFu;
    for (int i = 0; i < 32; i++) m = static_cast<uint32_t>(b[i]) < m ? static_cast<uint32_t>(b[i]) : m;
    return m;
}
LAB_DEV uint32_t lab_reduce_max(uint32_t v) {
    uint64_t const *b = simt::exchange(v);
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) m = static_cast<uint32_t>(b[i]) > m ? static_cast<uint32_t>(b[i]) : m;
    return m;
}
LAB_DEV uint32_t lab_reduce_or(uint32_t v) {
    uint64_t const *b = simt::exchange(v);
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) m |= static_cast<uint32_t>(b[i]);
    return m;
}
LAB_DEV uint32_t lab_reduce_add(uint32_t v) {
    uint64_t const *b = simt::exchange(v);
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) m += static_cast<uint32_t>(b[i]);
    return m;
}
LAB_DEV void lab_syncwarp() { (void)simt::exchange(0); }
LAB_DEV uint32_t lab_match_any(uint32_t v) {
    uint64_t const *b = simt::exchange(v);
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) m |= (static_cast<uint32_t>(b[i]) == v ? 1u : 0u) << i;
    return m;
}

LAB_DEV int lab_ffs64(uint64_t v) { return v ? __builtin_ctzll(v) + 1 : 0; }
LAB_DEV int lab_popc64(uint64_t v) { return __builtin_popcountll(v); }
LAB_DEV int lab_clz64(uint64_t v) { return v ? __builtin_clzll(v) : 64; }
LAB_DEV int lab_ffs32(uint32_t v) { return v ? __builtin_ctz(v) + 1 : 0; }
LAB_DEV int lab_clz32(uint32_t v) { return v ? __builtin_clz(v) : 32; }

template <class T> LAB_DEV T lab_ld_cg(T const *p) {
    simt::maybe_yield();
    return *p;
}
template <class T> LAB_DEV T lab_ld_volatile(T const *p) {
    simt::maybe_yield();
    return *p;
}
template <class T> LAB_DEV T lab_ld_ro(T const *p) { return *p; }
template <class T> LAB_DEV void lab_st_cg(T *p, typename std::common_type<T>::type v) { *p = v; }
struct lab_u32x4 {
    uint32_t x, y, z, w;
};
LAB_DEV lab_u32x4 lab_ld_stream_v4(void const *p) {
    lab_u32x4 v;
    std::memcpy(&v, p, sizeof v);
    return v;
}
LAB_DEV lab_u32x4 lab_ld_ro_v4(void const *p) { return lab_ld_stream_v4(p); }
LAB_DEV void lab_st_stream_v4(void *p, lab_u32x4 v) { std::memcpy(p, &v, sizeof v); }
LAB_DEV void lab_st_v4(void *p, lab_u32x4 v) { std::memcpy(p, &v, sizeof v); }
struct lab_u32x2 {
    uint32_t x, y;
};
LAB_DEV lab_u32x2 lab_ld_cg_x2(uint32_t const *base, uint32_t i) {
    simt::maybe_yield();
    return {base[2 * static_cast<size_t>(i)], base[2 * static_cast<size_t>(i) + 1]};
}
LAB_DEV void lab_st_x2(uint32_t *base, uint32_t i, uint32_t x, uint32_t y) {
    base[2 * static_cast<size_t>(i)] = x;
    base[2 * static_cast<size_t>(i) + 1] = y;
}
LAB_DEV void lab_st_pair(uint32_t *p, uint32_t x, uint32_t y) {
    if (reinterpret_cast<uintptr_t>(p) & 7u) {
        std::fprintf(stderr, "simt: lab_st_pair to an address that is not on an 8-byte boundary\n");
        std::abort();
    }
    p[0] = x;
    p[1] = y;
}
LAB_DEV void lab_st_volatile(uint32_t *p, uint32_t v) { *p = v; }
LAB_DEV void lab_st_volatile(uint16_t *p, uint16_t v) { *p = v; }

template <class T> LAB_DEV T lab_atomic_add(T *p, typename std::common_type<T>::type v) {
    T const o = *p;
    *p = o + v;
    return o;
}
LAB_DEV uint32_t lab_atomic_sub(uint32_t *p, uint32_t v) {
    uint32_t const o = *p;
    *p = o - v;
    return o;
}
template <class T> LAB_DEV T lab_atomic_or(T *p, typename std::common_type<T>::type v) {
    T const o = *p;
    *p = o | v;
    return o;
}
LAB_DEV uint32_t lab_atomic_and(uint32_t *p, uint32_t v) {
    uint32_t const o = *p;
    *p = o & v;
    return o;
}
LAB_DEV uint32_t lab_atomic_min(uint32_t *p, uint32_t v) {
    uint32_t const o = *p;
    *p = v < o ? v : o;
    return o;
}
LAB_DEV uint32_t lab_atomic_max(uint32_t *p, uint32_t v) {
    uint32_t const o = *p;
    *p = v > o ? v : o;
    return o;
}
LAB_DEV uint32_t lab_atomic_cas(uint32_t *p, uint32_t expect, uint32_t v) {
    uint32_t const o = *p;
    if (o == expect) *p = v;
    return o;
}
LAB_DEV uint32_t lab_atomic_exch(uint32_t *p, uint32_t v) {
    uint32_t const o = *p;
    *p = v;
    return o;
}
LAB_DEV void lab_threadfence() {}
LAB_DEV void lab_nanosleep(unsigned) { simt::park(simt::RUNNABLE); }
LAB_DEV uint64_t lab_cycles() { return 0; }
LAB_DEV uint64_t lab_clock_ns() {
    return static_cast<uint64_t>(
        std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count());
}

// every warp-wide primitive records the source line it was called from; run_round() aborts if
// the 32 lanes of a warp meet in primitives at different lines (divergent *_sync = deadlock on HW)
#define lab_shfl_up(...) (simt::g_site = __LINE__, lab_shfl_up(__VA_ARGS__))
#define lab_shfl_down(...) (simt::g_site = __LINE__, lab_shfl_down(__VA_ARGS__))
#define lab_shfl(...) (simt::g_site = __LINE__, lab_shfl(__VA_ARGS__))
#define lab_ballot(...) (simt::g_site = __LINE__, lab_ballot(__VA_ARGS__))
#define lab_any(...) (simt::g_site = __LINE__, lab_any(__VA_ARGS__))
#define lab_reduce_min(...) (simt::g_site = __LINE__, lab_reduce_min(__VA_ARGS__))
#define lab_reduce_max(...) (simt::g_site = __LINE__, lab_reduce_max(__VA_ARGS__))
#define lab_reduce_or(...) (simt::g_site = __LINE__, lab_reduce_or(__VA_ARGS__))
#define lab_reduce_add(...) (simt::g_site = __LINE__, lab_reduce_add(__VA_ARGS__))
#define lab_syncwarp() (simt::g_site = __LINE__, lab_syncwarp())
#define lab_match_any(...) (simt::g_site = __LINE__, lab_match_any(__VA_ARGS__))
