/*
 * simt_emu.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A tiny CPU emulation of one CUDA warp so that the *kernel source itself* (lz4net_b200/csrc/*.cuh) can be
 * compiled with g++ and exercised by `pytest -m "not gpu"` in a container that has no GPU: each lane runs as a
 * ucontext coroutine; every warp collective (shfl / ballot / match_any / syncwarp) is a rendezvous of the lanes
 * named in its mask.  Lanes are scheduled in a seeded pseudo-random order between collectives, so code that relies
 * on lock-step execution without a __syncwarp() is likely to show up as a mismatch here.
 * It is never linked into liblz4b200.so and is not a product code path (there is no CPU fallback).
 */
#pragma once
#include <stdint.h>
#include <string.h>

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };

namespace simt_emu {
// run `fn(lane, arg)` on 32 lanes as one warp
void run_warp(void (*fn)(int lane, void* arg), void* arg, uint64_t sched_seed);
uint32_t collective(int op, uint32_t mask, uint32_t a, uint32_t b);
int current_lane();
void yield();          // let the other lanes run (used by emulated spin-waits)
enum { OP_SYNC = 0, OP_BALLOT = 1, OP_SHFL = 2, OP_MATCH = 3, OP_RMAX = 4 };
}

namespace simt {
static inline uint32_t shfl(uint32_t mask, uint32_t v, int src) { return simt_emu::collective(simt_emu::OP_SHFL, mask, v, (uint32_t)src); }
static inline uint32_t ballot(uint32_t mask, bool p) { return simt_emu::collective(simt_emu::OP_BALLOT, mask, p ? 1u : 0u, 0); }
static inline uint32_t match_any(uint32_t mask, uint32_t v) { return simt_emu::collective(simt_emu::OP_MATCH, mask, v, 0); }
static inline uint32_t reduce_max(uint32_t mask, uint32_t v) { return simt_emu::collective(simt_emu::OP_RMAX, mask, v, 0); }
static inline void syncwarp(uint32_t mask) { simt_emu::collective(simt_emu::OP_SYNC, mask, 0, 0); }
static inline int ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline int popc(uint32_t v) { return __builtin_popcount(v); }
static inline uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { uint64_t t = ((uint64_t)hi << 32) | lo; return (uint32_t)(t >> (sh & 31)); }
static inline uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t sh) { uint64_t t = ((uint64_t)hi << 32) | lo; return (uint32_t)((t << (sh & 31)) >> 32); }

template <class T> static inline T* keep(T* p) { return p; }
static inline uint32_t keep(uint32_t v) { return v; }
static inline uint8_t  ldg_u8(const uint8_t* p) { return *p; }
static inline uint32_t ldg_u32(const void* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint4    ldg_v4(const void* p) { uint4 v; memcpy(&v, p, 16); return v; }
static inline uint8_t  ldg_nc_u8(const uint8_t* p) { return *p; }
static inline uint32_t ldg_nc_u32(const void* p) { return ldg_u32(p); }
static inline uint4    ldg_nc_v4(const void* p) { return ldg_v4(p); }
static inline uint32_t ldg_u16(const uint16_t* p) { return *p; }
static inline void stg_u16(uint16_t* p, uint32_t v) { *p = (uint16_t)v; }
static inline void stg_u8(uint8_t* p, uint8_t v) { *p = v; }
static inline void stg_u32(void* p, uint32_t v) { memcpy(p, &v, 4); }
static inline void stg_v4(void* p, uint4 v) { memcpy(p, &v, 16); }
static inline uint64_t l2_policy_keep() { return 0; }
static inline uint64_t l2_policy_stream() { return 0; }
static inline uint32_t ldg_nc_hint_u32(const void* p, uint64_t) { return ldg_u32(p); }
static inline void stg_hint_u8(uint8_t* p, uint32_t v, uint64_t) { *p = (uint8_t)v; }
static inline void tie(uint32_t&, uint32_t) {}
struct smem_ref { uint8_t* p; };
static inline smem_ref smem_ref_of(const void* p) { return smem_ref{(uint8_t*)p}; }
static inline uint32_t lds_u16(smem_ref r, uint32_t off) { uint16_t v; memcpy(&v, r.p + off, 2); return v; }
static inline uint32_t lds_u32(smem_ref r, uint32_t off) { uint32_t v; memcpy(&v, r.p + off, 4); return v; }
static inline uint32_t lds_u8(smem_ref r, uint32_t off) { return r.p[off]; }
static inline uint4 lds_v4(smem_ref r, uint32_t off) { uint4 v; memcpy(&v, r.p + off, 16); return v; }
static inline void sts_u8(smem_ref r, uint32_t off, uint32_t v) { r.p[off] = (uint8_t)v; }
static inline void sts_u16(smem_ref r, uint32_t off, uint32_t v) { uint16_t t = (uint16_t)v; memcpy(r.p + off, &t, 2); }
static inline void sts_u32(smem_ref r, uint32_t off, uint32_t v) { memcpy(r.p + off, &v, 4); }
static inline void sts_v4(smem_ref r, uint32_t off, uint4 v) { memcpy(r.p + off, &v, 16); }
// a lane is only descheduled inside a collective, so a plain read-modify-write is atomic here; lanes reach it in the
// scheduler's pseudo-random order, like the hardware's unspecified order among the lanes of one ATOMS
static inline uint32_t atoms_add(smem_ref r, uint32_t off, uint32_t v) { uint32_t o; memcpy(&o, r.p + off, 4); uint32_t n = o + v; memcpy(r.p + off, &n, 4); return o; }
// cp.async: the copy is DEFERRED until the wait that covers its group (per lane), so that code which reads a unit before
// waiting for it, or overwrites a ring slot that is still to be read, fails in the emulator as well
void cp_async16_emu(void* sdst, const void* gsrc);
void cp_async_commit_emu();
void cp_async_wait_emu(int n);
static inline void cp_async16(smem_ref r, uint32_t off, const void* g) { cp_async16_emu(r.p + off, g); }
static inline void cp_async_commit() { cp_async_commit_emu(); }
template <int N> static inline void cp_async_wait() { cp_async_wait_emu(N); }
static inline uint32_t atomic_inc(uint32_t* p) { return (*p)++; }
static inline uint32_t atomg_add(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline uint32_t ldg_cg_u32(const uint32_t* p) { return *p; }
static inline void prefetch_l1(const void*) {}
static inline void prefetch_l2(const void*) {}

// async bulk copy global -> shared with an mbarrier: immediate in the emulator
struct mbar_t { uint64_t v; };
static inline void mbar_init(mbar_t* b, int) { b->v = 0; }
// the copy "lands" at once and completes one barrier phase; waiters poll the phase parity like mbarrier.try_wait.parity
// (the issuing lane runs expect + copies without yielding, so waiters never see a half-filled phase)
static inline void mbar_expect(mbar_t* b, uint32_t) { b->v++; }
static inline void bulk_copy(void* sdst, const void* gsrc, uint32_t bytes, mbar_t*) { memcpy(sdst, gsrc, bytes); }
static inline void mbar_wait(mbar_t* b, uint32_t parity) { while ((b->v & 1) == (parity & 1)) simt_emu::yield(); }
static inline void fence_mbar_init() {}
}
#define SIMT_DEV static inline
#define SIMT_MEM inline
#define SIMT_NOINLINE static
