// TEST INFRASTRUCTURE ONLY -- see simt_emu.h
#include "simt_emu.h"
#include <ucontext.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <deque>

namespace simt_emu {

static const int NL = 32;
static const size_t STACK = 512 * 1024;

struct Slot { uint32_t arrived; uint32_t phase; uint32_t a[NL], b[NL], res[NL]; int op; uint32_t mask; };

static ucontext_t g_main, g_ctx[NL];
static bool g_done[NL];
static int g_cur = -1;
static Slot g_slot[NL];
static void (*g_fn)(int, void*);
static void* g_arg;
static uint64_t g_progress;

int current_lane() { return g_cur; }

static void yield_to_main() { int me = g_cur; swapcontext(&g_ctx[me], &g_main); }
static uint64_t g_spins;
void yield() { if (++g_spins > 100000000ull) { fprintf(stderr, "simt_emu: livelock in a spin-wait\n"); abort(); } g_progress++; yield_to_main(); }

static void finish(Slot& s)
{
    for (int l = 0; l < NL; l++) {
        if (!(s.mask >> l & 1)) continue;
        switch (s.op) {
        case OP_SYNC: s.res[l] = 0; break;
        case OP_BALLOT: { uint32_t r = 0; for (int k = 0; k < NL; k++) if ((s.mask >> k & 1) && s.a[k]) r |= 1u << k; s.res[l] = r; } break;
        case OP_SHFL: { int src = (int)(s.b[l] & 31); s.res[l] = (s.mask >> src & 1) ? s.a[src] : s.a[l]; } break;
        case OP_RMAX: { uint32_t r = 0; for (int k = 0; k < NL; k++) if ((s.mask >> k & 1) && s.a[k] > r) r = s.a[k]; s.res[l] = r; } break;
        case OP_MATCH: { uint32_t r = 0; for (int k = 0; k < NL; k++) if ((s.mask >> k & 1) && s.a[k] == s.a[l]) r |= 1u << k; s.res[l] = r; } break;
        }
    }
}

uint32_t collective(int op, uint32_t mask, uint32_t a, uint32_t b)
{
    int me = g_cur;
    if (!(mask >> me & 1)) { fprintf(stderr, "simt_emu: lane %d not in its own mask %08x\n", me, mask); abort(); }
    Slot& s = g_slot[__builtin_ffs((int)mask) - 1];
    if (s.arrived == 0) { s.op = op; s.mask = mask; }
    else if (s.op != op || s.mask != mask) { fprintf(stderr, "simt_emu: divergent collective (op %d/%d mask %08x/%08x)\n", s.op, op, s.mask, mask); abort(); }
    uint32_t my_phase = s.phase;
    s.a[me] = a; s.b[me] = b; s.arrived |= 1u << me;
    g_progress++;
    if (s.arrived == mask) { finish(s); s.arrived = 0; s.phase++; }
    else while (s.phase == my_phase) yield_to_main();
    return s.res[me];
}

static void trampoline(int lane)
{
    g_fn(lane, g_arg);
    g_done[lane] = true;
    g_progress++;
    yield_to_main();
}

void run_warp(void (*fn)(int, void*), void* arg, uint64_t seed)
{
    static char* stacks = 0;
    if (!stacks) stacks = (char*)malloc(STACK * NL);
    g_fn = fn; g_arg = arg;
    memset(g_slot, 0, sizeof g_slot);
    for (int l = 0; l < NL; l++) {
        g_done[l] = false;
        getcontext(&g_ctx[l]);
        g_ctx[l].uc_stack.ss_sp = stacks + STACK * l;
        g_ctx[l].uc_stack.ss_size = STACK;
        g_ctx[l].uc_link = &g_main;
        makecontext(&g_ctx[l], (void (*)())trampoline, 1, l);
    }
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 1;
    for (;;) {
        int alive = 0; uint64_t before = g_progress;
        // one pass over all lanes in a rotated order (seeded) -- no lock-step guarantee between collectives
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        int start = (int)(st % NL), dir = (st >> 8) & 1 ? 1 : NL - 1;
        for (int i = 0; i < NL; i++) {
            int l = (start + i * dir) % NL;
            if (g_done[l]) continue;
            alive++;
            g_cur = l;
            swapcontext(&g_main, &g_ctx[l]);
        }
        g_cur = -1;
        if (!alive) break;
        if (g_progress == before) {
            fprintf(stderr, "simt_emu: deadlock (a collective is missing lanes)\n");
            for (int k = 0; k < NL; k++) if (g_slot[k].arrived)
                fprintf(stderr, "  slot %d: op %d mask %08x arrived %08x\n", k, g_slot[k].op, g_slot[k].mask, g_slot[k].arrived);
            for (int k = 0; k < NL; k++) fprintf(stderr, "%d", (int)g_done[k]);
            fprintf(stderr, " (done flags)\n");
            abort();
        }
    }
}
}

// ---- deferred cp.async (per lane) -------------------------------------------------------------------------------
namespace simt {
struct PendingCopy { void* dst; const void* src; };
static std::deque<std::vector<PendingCopy>> g_groups[32];
static std::vector<PendingCopy> g_open[32];
void cp_async16_emu(void* sdst, const void* gsrc)
{
    const int l = simt_emu::current_lane();
    // two copies in flight to the same shared-memory unit land in no defined order on the device: a bug, whatever the
    // emulator's own (in-order) completion would make of it
    for (const auto& g : g_groups[l]) for (const PendingCopy& c : g) if (c.dst == sdst) { fprintf(stderr, "simt_emu: two cp.async in flight to one shared-memory unit (lane %d)\n", l); abort(); }
    for (const PendingCopy& c : g_open[l]) if (c.dst == sdst) { fprintf(stderr, "simt_emu: two cp.async in flight to one shared-memory unit (lane %d)\n", l); abort(); }
    g_open[l].push_back(PendingCopy{sdst, gsrc});
}
void cp_async_commit_emu() { int l = simt_emu::current_lane(); g_groups[l].push_back(g_open[l]); g_open[l].clear(); }
void cp_async_wait_emu(int n)
{
    int l = simt_emu::current_lane();
    while ((int)g_groups[l].size() > n) {
        for (const PendingCopy& c : g_groups[l].front()) memcpy(c.dst, c.src, 16);
        g_groups[l].pop_front();
    }
}
}
