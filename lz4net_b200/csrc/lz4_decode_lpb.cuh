// lz4_decode_lpb.cuh -- batched LZ4 block decoder, ONE LANE per block ("lane-per-block"): for sequence-dense data.
//
// Same contract as lz4_decode.cuh (LZ4_uncompress, original/lz4.c:812-914, and LZ4_uncompress_unknownOutputSize,
// :916-1044; lz4net: src/LZ4ps/LZ4Codec.Safe64.Dirty.cs:533-659,665-798), different mapping onto the machine.
//
// Why: on token-dense streams (text: 3 compressed bytes -> 8 decoded bytes per sequence) the group-per-block decoder
// is bound by warp instructions per SEQUENCE -- every header costs the whole group a dependent parse, whatever its
// width.  Here every lane parses and copies the sequences of its OWN block, so one warp instruction serves 32
// sequences of 32 different blocks; the per-sequence chain (token -> lengths -> offset -> copy) is SIMT-parallel
// across blocks instead of being repeated per block.
//
// Data flow per lane (all of it private to the lane, no warp collective on the common path):
//   compressed stream --cp.async, 16 B per request, issued a full iteration ahead--> a 256 B input ring in shared memory
//   -> header bytes by LDS; literals ring -> output ring, matches output ring -> output ring, in 32-bit words with the
//   source funnel-shifted into the destination's alignment (SHF), bytes only for heads, tails and overlaps < 8;
//   output ring (the last OUT bytes of the block: the window that serves every match with offset <= OUT - 8)
//   --LDS.128 + STG.128, 64 B at a time--> global memory.  Matches further back read the lane's own earlier output
//   from global memory (same thread: program order is enough).
// The lanes' rings are laid out 16 bytes past a multiple of 128 apart, so that the 128-bit accesses of the fills and
// flushes are conflict-free and word accesses of lanes in lock step are spread over eight bank groups.
// Runs that a single lane would take too long over (literal runs and matches longer than 64 bytes -- incompressible
// stretches, RLE) are handed to the whole warp: the lane publishes (source, destination, length, offset), the warp
// copies with coalesced 128-bit moves (lz4_copy.cuh), and the lane carries on behind the run.
//
// Accept / reject decisions are those of the reference's 64-bit flavour, restated per lane exactly as in
// decode_careful (lz4_decode.cuh); a malformed stream yields a negative result and never an access outside
// [src, src+isize) (rounded out to 16-byte units) or [dst, dst+cap).
#pragma once
#include "simt.cuh"
#include "lz4_copy.cuh"

namespace lz4b200 {

template <int IN_, int OUT_>
struct LpbGeom {
    static constexpr int IN = IN_, OUT = OUT_;
    static constexpr int LANE_BYTES = IN + OUT + 16;               // == 16 (mod 128)
    static constexpr int MAXL = 64, MAXM = 64;                     // longest literal run / match a lane copies by itself
    static constexpr int LOOK = 1 + 1 + MAXL + 2 + 1 + 4;          // stream bytes such a sequence can touch (+ word over-read)
    static constexpr int WIN = OUT - 8;                            // matches up to this far back are served by the output ring
    static constexpr int FLUSH = 64;                               // bytes written out per flush step
    static_assert((IN & (IN - 1)) == 0 && (OUT & (OUT - 1)) == 0 && (IN + OUT) % 128 == 0, "ring sizes");
    static_assert(IN >= 3 * LOOK + 16 && OUT >= FLUSH + MAXL + MAXM + 16, "ring capacity");     // (unflushed bytes never exceed FLUSH - 1 + one sequence)
};

template <class GEO> struct alignas(16) LpbShared { uint8_t lane[32][GEO::LANE_BYTES]; };

struct LpbBatch {
    const uint8_t* src; const int64_t* src_off; const int32_t* src_len;
    uint8_t* dst; const int64_t* dst_off; const int32_t* dst_cap;
    int32_t* out_len; int32_t n_blocks;
};

// One warp: every lane decodes blocks of its own, taken from the global counter, until the batch is exhausted.
template <bool KNOWN, class GEO>
SIMT_DEV void lpb_decode_warp(LpbShared<GEO>* sh, const LpbBatch& a, uint32_t* counter, int lane)
{
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr uint32_t IMASK = GEO::IN - 1, OMASK = GEO::OUT - 1;
    constexpr int LEN_LIMIT = 0x3FFFFFFF;
    const simt::smem_ref ir = simt::smem_ref_of(sh->lane[lane]);
    const simt::smem_ref orr = simt::smem_ref_of(sh->lane[lane] + GEO::IN);

    // ---- lane state -------------------------------------------------------------------------------------------------
    bool active = false, drained = false;
    uint32_t blk = 0;
    const uint8_t* abase = nullptr;     // compressed stream, rounded down to 16 bytes
    uint8_t* gbase = nullptr;           // output block, rounded down to 16 bytes
    uint32_t skew = 0, total = 0;       // src - abase; round_up16(skew + isize)
    uint32_t a0 = 0;                    // dst - gbase: "virtual" output position v = op + a0, so that gbase + v is the address
    int isize = 0, cap = 0;
    int ip = 0;                         // read cursor (stream position)
    uint32_t vop = 0;                   // write cursor (virtual)
    uint32_t fpos = 0;                  // virtual position up to which the output is in global memory
    uint32_t slo = 0;                   // lowest virtual position whose byte is (still) valid in the output ring
    uint32_t ifill = 0;                 // aligned stream offset (from abase) of the next 16-byte unit to request
    uint32_t f1 = 0;                    // ifill at the last commit
    uint32_t iland = 0;                 // every unit below this offset has landed in the ring
    int phase = 0;                      // 0: a sequence header comes next; 1: offset + match of the current sequence
    uint32_t token = 0;
    int coop = 0;                       // 0 none, 1 literal run, 2 match: a copy the whole warp makes for this lane
    int clen = 0; uint32_t coff = 0;
    bool fresh = false;                 // the block was set up in this iteration: its first units have not landed yet

    auto ib = [&](int p) -> uint32_t { return simt::lds_u8(ir, ((uint32_t)p + skew) & IMASK); };
    auto ob = [&](uint32_t v) -> uint32_t { return simt::lds_u8(orr, v & OMASK); };
    // request every unit that fits: the ring may hold [need & ~15, (need & ~15) + IN) where need = oldest offset still wanted
    auto refill = [&](uint32_t need_off) {
        const uint32_t lim = (need_off & ~15u) + (uint32_t)GEO::IN;
        while (ifill < total && ifill < lim) {
            simt::cp_async16(ir, ifill & IMASK, abase + ifill);
            ifill += 16;
        }
    };
    // make stream byte p readable NOW; lo = the oldest stream position still wanted (p - lo < IN - 16).  Slow path: the
    // first units of a block, long length runs, the bytes behind a run the warp copied, an iteration that consumed more
    // than the look-ahead covers
    auto fetch = [&](int lo, int p) {
        const uint32_t q = (uint32_t)p + skew;
        if (q < iland) return;
        // Units requested earlier may still be in flight towards slots the new requests will use (the ring window moves up
        // to `lo`, possibly past units that were requested but never read): two copies in flight to one slot land in no
        // defined order, so what is in flight is drained first.
        simt::cp_async_commit(); simt::cp_async_wait<0>();
        const uint32_t start = ((uint32_t)lo + skew) & ~15u;
        if (start > ifill) ifill = start;                          // units below are never read
        refill(start);
        simt::cp_async_commit(); simt::cp_async_wait<0>();
        f1 = iland = ifill;
    };
    auto ibx = [&](int p) -> uint32_t { fetch(p, p); return ib(p); };
    // bytes [fpos, hi) of the output ring -> global memory: bytes up to a 16-byte boundary, then 128-bit units
    auto flush_to = [&](uint32_t hi) {
        while ((fpos & 15u) && fpos < hi) { simt::stg_u8(gbase + fpos, (uint8_t)ob(fpos)); fpos++; }
        while (fpos + 64 <= hi) {
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = simt::lds_v4(orr, (fpos + 16u * k) & OMASK);
#pragma unroll
            for (int k = 0; k < 4; k++) simt::stg_v4(gbase + fpos + 16u * k, v[k]);
            fpos += 64;
        }
        while (fpos + 16 <= hi) { simt::stg_v4(gbase + fpos, simt::lds_v4(orr, fpos & OMASK)); fpos += 16; }
    };
    auto flush_all = [&]() {
        flush_to(vop);
        while (fpos < vop) { simt::stg_u8(gbase + fpos, (uint8_t)ob(fpos)); fpos++; }
    };
    // n bytes -> output ring at virtual position v.  SRC: 0 input ring (stream position s), 1 output ring (virtual position
    // s, at least 8 back), 2 global memory (virtual position s of the own output, not overlapping the destination)
    auto src_byte = [&](int kind, uint32_t s) -> uint32_t {
        return kind == 0 ? simt::lds_u8(ir, (s + skew) & IMASK) : (kind == 1 ? simt::lds_u8(orr, s & OMASK) : (uint32_t)simt::ldg_u8(gbase + s));
    };
    auto src_word = [&](int kind, uint32_t s4) -> uint32_t {       // aligned word holding source index s4 (ring offset / address rounded down)
        return kind == 0 ? simt::lds_u32(ir, ((s4 + skew) & ~3u) & IMASK)
             : (kind == 1 ? simt::lds_u32(orr, (s4 & ~3u) & OMASK) : simt::ldg_u32(gbase + (s4 & ~3u)));
    };
    auto copy_in = [&](int kind, uint32_t s, uint32_t v, uint32_t n) {
        uint32_t i = 0;
        if (n >= 8) {
            while ((v + i) & 3u) { simt::sts_u8(orr, (v + i) & OMASK, src_byte(kind, s + i)); i++; }
            // the source as aligned words, funnel-shifted: bytes s+i .. s+i+3 = (lo, hi) >> 8 * misalignment
            const uint32_t mis = kind == 0 ? ((s + i + skew) & 3u) : ((s + i) & 3u);
            const uint32_t sh8 = mis * 8u;
            uint32_t lo = src_word(kind, s + i);
            for (; i + 4 <= n; i += 4) {
                const uint32_t hi = mis ? src_word(kind, s + i + 4) : 0u;
                simt::sts_u32(orr, (v + i) & OMASK, simt::funnel_r(lo, hi, sh8));
                lo = mis ? hi : src_word(kind, s + i + 4);
            }
        }
        for (; i < n; i++) simt::sts_u8(orr, (v + i) & OMASK, src_byte(kind, s + i));
    };

    for (;;) {
        // ---------------- idle lanes take the next block ----------------
        if (!active && !drained) {
            blk = simt::atomic_inc(counter);
            if (blk >= (uint32_t)a.n_blocks) drained = true;
            else {
                const uint8_t* src = a.src + a.src_off[blk];
                uint8_t* dst = a.dst + a.dst_off[blk];
                isize = a.src_len[blk]; cap = a.dst_cap[blk];
                if (isize <= 0 || cap < 0) simt::stg_u32(a.out_len + blk, (uint32_t)-1);       // original/lz4.c:949; a block has >= 1 token
                else {
                    skew = (uint32_t)((uintptr_t)src & 15); abase = src - skew;
                    total = (skew + (uint32_t)isize + 15u) & ~15u;
                    a0 = (uint32_t)((uintptr_t)dst & 15); gbase = dst - a0;
                    ip = 0; vop = fpos = slo = a0; phase = 0; coop = 0;
                    simt::cp_async_commit(); simt::cp_async_wait<0>();     // (units the previous block requested but never read)
                    ifill = 0; refill(0);
                    active = true; fresh = true;
                }
            }
        }
        if (!simt::ballot(FULL, active)) break;
        // everything requested up to the last commit but one has landed; a fresh block waits for its first units
        simt::cp_async_commit();
        if (simt::ballot(FULL, fresh)) { simt::cp_async_wait<0>(); f1 = iland = ifill; }
        else { simt::cp_async_wait<1>(); iland = f1; f1 = ifill; }  // (all groups but the one just committed are complete: everything below the previous commit's ifill)
        fresh = false;

        int result = 0; bool done = false;
        if (active) {
            const int op = (int)(vop - a0);
            // ---------------- header + literals ----------------
            if (phase == 0) {
                if (ip < isize) fetch(ip, ip + GEO::LOOK < isize ? ip + GEO::LOOK : isize - 1);
                if (ip >= isize) { result = -ip - 1; done = true; }
                else {
                    token = ib(ip); ip++;
                    int L = (int)(token >> 4);
                    if (L == 15) {                                          // :843 / :959-963
                        uint32_t s = 255;
                        if (KNOWN) { do { if (ip >= isize || L > LEN_LIMIT) { L = -1; break; } s = ibx(ip++); L += (int)s; } while (s == 255); }
                        else       { while (ip < isize && s == 255 && L <= LEN_LIMIT) { s = ibx(ip++); L += (int)s; } }
                        if (L < 0 || L > LEN_LIMIT) { result = -ip - 1; done = true; }
                    }
                    if (!done && (L > isize - ip || L > cap - op)) { result = -ip - 1; done = true; }
                    if (!done) {
                        const int end = op + L;
                        bool last;
                        if (KNOWN) last = end > cap - 8;                    // :847
                        else       last = end > cap - 12 || ip + L > isize - 8;    // :968
                        bool ok = true;
                        if (last) ok = KNOWN ? (end == cap && ip + L <= isize) : (end <= cap && ip + L == isize);   // :849-857 / :974-975
                        else      ok = ip + L + 2 <= isize;
                        if (!ok) { result = -ip - 1; done = true; }
                        else {
                            if (L > GEO::MAXL) { coop = 1; clen = L; }     // the warp copies it: src + ip -> dst + op
                            else {
                                copy_in(0, (uint32_t)ip, vop, (uint32_t)L);
                                ip += L; vop += (uint32_t)L;
                            }
                            phase = last ? 2 : 1;
                        }
                    }
                }
            }
            // ---------------- offset + match ----------------
            if (!done && coop == 0 && phase == 1) {
                const int op1 = (int)(vop - a0);
                fetch(ip, ip + 8 < isize ? ip + 8 : isize - 1);
                uint32_t off = ib(ip) | (ib(ip + 1) << 8); ip += 2;         // :862 / :982 (ip + 2 <= isize was checked with the literals)
                if (off == 0 || off > (uint32_t)op1) { result = -ip - 1; done = true; }          // :863 / :983 (offset 0 rejected by design)
                else {
                    int M = (int)(token & 15);
                    if (M == 15) {                                          // :866 / :986-999
                        uint32_t s = 255;
                        if (KNOWN) { do { if (ip >= isize || M > LEN_LIMIT) { M = -1; break; } s = ibx(ip++); M += (int)s; } while (s == 255); }
                        else       { while (ip < isize - 6 && M <= LEN_LIMIT) { s = ibx(ip++); M += (int)s; if (s != 255) break; } }
                        if (M < 0 || M > LEN_LIMIT) { result = -ip - 1; done = true; }
                    }
                    if (!done && M > cap - op1) { result = -ip - 1; done = true; }
                    if (!done) {
                        const int n = M + 4;
                        if (op1 + n > cap - 5) { result = -ip - 1; done = true; }               // :893 / :1025 -- the last 5 bytes are literals
                        else if (n > GEO::MAXM) { coop = 2; clen = n; coff = off; }
                        else {
                            const uint32_t s = vop - off;
                            if (off <= (uint32_t)GEO::WIN && s >= slo) {
                                if (off >= 8) copy_in(1, s, vop, (uint32_t)n);
                                else for (int i = 0; i < n; i++) simt::sts_u8(orr, (vop + (uint32_t)i) & OMASK, ob(s + (uint32_t)i));   // overlap: byte by byte, in order
                            } else if (off > (uint32_t)GEO::WIN) {
                                // beyond the window: in global memory for good (off > WIN >= n: no overlap with the destination,
                                // and everything below vop - WIN + n was flushed long ago)
                                copy_in(2, s, vop, (uint32_t)n);
                            } else {
                                // just behind a run the warp copied: the source straddles the ring's valid range
                                for (int i = 0; i < n; i++) {
                                    const uint32_t p = s + (uint32_t)i;
                                    simt::sts_u8(orr, (vop + (uint32_t)i) & OMASK, p >= slo ? ob(p) : (uint32_t)simt::ldg_u8(gbase + p));
                                }
                            }
                            vop += (uint32_t)n;
                            phase = 0;
                        }
                    }
                }
            }
            if (!done && coop == 0 && phase == 2) {                        // the last literals are in: finished
                result = KNOWN ? ip : (int)(vop - a0);
                done = true;
            }
            if (done || coop) flush_all();                                  // the warp's copy / the end of the block: everything out
            else if (vop - fpos >= (uint32_t)GEO::FLUSH) flush_to(vop & ~15u);
        }
        // ---------------- runs the whole warp copies ----------------
        uint32_t req = simt::ballot(FULL, active && coop != 0);
        while (req) {
            const int k = simt::ffs(req) - 1; req &= req - 1;
            const int kind = (int)simt::shfl(FULL, (uint32_t)coop, k);
            const uint32_t n = simt::shfl(FULL, (uint32_t)clen, k);
            const uint32_t o = simt::shfl(FULL, coff, k);
            const uint64_t dpk = (uint64_t)(uintptr_t)(gbase + vop);
            const uint64_t spk = (uint64_t)(uintptr_t)(abase + skew + (uint32_t)ip);
            uint8_t* const d = (uint8_t*)(uintptr_t)(((uint64_t)simt::shfl(FULL, (uint32_t)(dpk >> 32), k) << 32) | simt::shfl(FULL, (uint32_t)dpk, k));
            const uint8_t* const s = (const uint8_t*)(uintptr_t)(((uint64_t)simt::shfl(FULL, (uint32_t)(spk >> 32), k) << 32) | simt::shfl(FULL, (uint32_t)spk, k));
            simt::syncwarp(FULL);                                           // lane k's flush -> everybody's loads
            if (kind == 1) { InputSrc sp{s}; group_copy<32, false>(d, sp, n, lane, FULL); }
            else group_copy_match<32>(d, o, n, lane, FULL);
            simt::syncwarp(FULL);                                           // the copy -> lane k's later loads
        }
        if (active && coop != 0) {
            if (coop == 1) {                                                // the input ring jumps behind the run (drains what is in flight first)
                ip += clen;
                if (phase == 1) fetch(ip, ip + 8 < isize ? ip + 8 : isize - 1);
            }
            else phase = 0;
            vop += (uint32_t)clen; fpos = slo = vop;                        // the ring holds nothing of the block any more
            coop = 0;
            if (phase == 2) { flush_all(); result = KNOWN ? ip : (int)(vop - a0); done = true; }
        }
        if (active && done) { simt::stg_u32(a.out_len + blk, (uint32_t)result); active = false; }
        // ---------------- ask for the units the next iterations will read ----------------
        if (active) refill((uint32_t)ip + skew);
    }
    simt::cp_async_commit(); simt::cp_async_wait<0>();
}

}  // namespace lz4b200
