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
// Data flow per lane:
//   compressed stream --cp.async, 16 B per request, issued a full iteration ahead--> a private 256 B input ring
//   (shared memory, rows 16 bytes past a multiple of 128 apart) -> header bytes by LDS;
//   literals (input ring) and matches (output ring) are APPENDED to the output ring in aligned 32-bit words only: the
//   source is read as aligned words and funnel-shifted (SHF) to the destination's byte phase, the bytes that do not
//   fill a word yet are carried in a register and written through -- no byte loops for heads and tails;
//   the output ring holds the last OUT bytes of the block (the window that serves every match up to OUT - 8 back).
//   The lanes' output rings lie OUT + 4 bytes apart, so word w of lane l sits in bank (l + w) mod 32: the lanes' own
//   accesses are bank-conflict free when they run in lock step AND consecutive words of one lane can be read by
//   neighbouring lanes without conflicts: finished 128-byte chunks leave for global memory through quarter-warp
//   transposed reads and fully coalesced 128-bit stores (four lanes' chunks per step).
//   Matches further back than the window read the lane's earlier output from global memory.
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

// DEPTH: how many iterations ahead the input is requested (1: a unit is waited for one full iteration after its request,
// needs IN >= 3 LOOK + 16; 0: requests are waited for at the top of the next iteration, IN >= 2 LOOK + 16 -- the smaller
// rings let more warps share an SM, which is what hides the latency then).
template <int IN_, int OUT_, int MAXRUN_ = 64, int DEPTH_ = 1, int SEQS_ = 1>
struct LpbGeom {
    static constexpr int IN = IN_, OUT = OUT_, DEPTH = DEPTH_;
    static constexpr int SEQS = SEQS_;                             // sequences a lane takes per iteration of the warp loop
    static constexpr int IN_STRIDE = IN + 16;                      // == 16 (mod 128): the 128-bit fills of the eight lanes of a phase never collide
    static constexpr int MAXL = MAXRUN_, MAXM = MAXRUN_;           // longest literal run / match a lane copies by itself
    static constexpr int LOOK = 1 + 1 + MAXL + 2 + 1 + 4;          // stream bytes such a sequence can touch (+ word over-read)
    static constexpr int WIN = OUT - 8;                            // matches up to this far back are served by the output ring
    static constexpr int CHUNK = 128;                              // flush unit (bytes, aligned in the output buffer)
    static constexpr int HIGH = OUT - 256 > CHUNK ? OUT - 256 : CHUNK;   // a lane with this many unflushed bytes forces a flush step
    static_assert((IN & (IN - 1)) == 0 && (OUT & (OUT - 1)) == 0 && OUT >= 256, "ring sizes");
    static_assert(IN >= (2 + DEPTH) * LOOK + 16 && HIGH + MAXL + MAXM + 8 <= OUT, "ring capacity");
};

template <class GEO> struct alignas(128) LpbShared {
    uint8_t in[32][GEO::IN_STRIDE];                                // the lanes' input rings
    uint8_t out[32 * (GEO::OUT + 4) + 12];                         // the lanes' output rings, OUT + 4 bytes apart (see above)
};

struct LpbBatch {
    const uint8_t* src; const int64_t* src_off; const int32_t* src_len;
    uint8_t* dst; const int64_t* dst_off; const int32_t* dst_cap;
    int32_t* out_len; int32_t n_blocks;
};

template <int K> struct LpbKind { static constexpr int value = K; };

// One warp: every lane decodes blocks of its own, taken from the global counter, until the batch is exhausted.
template <bool KNOWN, class GEO>
SIMT_DEV void lpb_decode_warp(LpbShared<GEO>* sh, const LpbBatch& a, uint32_t* counter, int lane)
{
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr uint32_t IMASK = GEO::IN - 1;
    constexpr int LEN_LIMIT = 0x3FFFFFFF;
    const simt::smem_ref ir = simt::smem_ref_of(sh->in[lane]);
    const simt::smem_ref ow = simt::smem_ref_of(sh->out);
    const uint32_t l4 = (uint32_t)lane * (uint32_t)(GEO::OUT + 4);
    // byte offset (inside `out`) of the aligned word holding virtual position v4 (a multiple of 4) of the lane whose ring starts at lane4
    auto oword = [](uint32_t lane4, uint32_t v4) -> uint32_t { return lane4 + (v4 & (uint32_t)(GEO::OUT - 4)); };

    // ---- lane state -------------------------------------------------------------------------------------------------
    bool active = false, drained = false;
    uint32_t blk = 0;
    const uint8_t* abase = nullptr;     // compressed stream, rounded down to 16 bytes
    uint8_t* gbase = nullptr;           // output block, rounded down to 128 bytes
    int32_t g128 = 0;                   // the same as a (signed) count of 128-byte units from the batch's rounded base: one shuffle instead of two
    uint8_t* const dbase128 = (uint8_t*)((uintptr_t)a.dst & ~(uintptr_t)127);
    uint32_t skew = 0, total = 0;       // src - abase; round_up16(skew + isize)
    uint32_t a0 = 0;                    // dst - gbase: "virtual" output position v = op + a0, so that gbase + v is the address
    int isize = 0, cap = 0;
    int ip = 0;                         // read cursor (stream position)
    uint32_t vop = 0;                   // write cursor (virtual)
    uint32_t acc = 0;                   // the bytes of the word at vop & ~3 written so far (low vop & 3 bytes, the rest zero)
    uint32_t fpos = 0;                  // virtual position up to which the output is in global memory
    uint32_t slo = 0;                   // lowest virtual position whose byte is (still) valid in the output ring
    uint32_t ifill = 0;                 // aligned stream offset (from abase) of the next 16-byte unit to request
    uint32_t f1 = 0;                    // ifill at the last commit
    uint32_t iland = 0;                 // every unit below this offset has landed in the ring
    int phase = 0;                      // 0: a sequence header comes next; 1: offset + match of the current sequence; 2: finished
    uint32_t token = 0;
    int coop = 0;                       // 0 none, 1 literal run, 2 match: a copy the whole warp makes for this lane
    int clen = 0; uint32_t coff = 0;
    bool fresh = false;                 // the block was set up in this iteration: its first units have not landed yet
    bool need_all = false;              // everything of this lane must be in global memory before its next step (end, warp copy, far match)
    int result = 0; bool done = false;

    auto ib = [&](int p) -> uint32_t { return simt::lds_u8(ir, ((uint32_t)p + skew) & IMASK); };
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

    // Aligned source word at position p4 (a multiple of 4) of: K = 0 the input ring (p4 = stream offset from abase),
    // 1 the lane's output ring (virtual position), 2 the lane's output in global memory (virtual position).
    auto load_src = [&](auto kind, uint32_t p4) -> uint32_t {
        constexpr int K = decltype(kind)::value;
        if (K == 0) return simt::lds_u32(ir, p4 & IMASK);
        if (K == 1) return simt::lds_u32(ow, oword(l4, p4));
        return simt::ldg_u32(gbase + p4);
    };
    // Append n >= 1 bytes to the output ring: source bytes s .. s+n-1 (K as above; s = stream position for K = 0).
    // K = 1 requires the source to start at least 8 bytes back (the loop reads one aligned source word ahead).
    auto append = [&](auto kind, uint32_t s, uint32_t n) {
        constexpr int K = decltype(kind)::value;
        const uint32_t sq = K == 0 ? s + skew : s;
        const uint32_t ssh = (sq & 3u) * 8u;
        uint32_t sp = sq & ~3u;
        uint32_t lo = load_src(kind, sp);
        const uint32_t k = vop & 3u, dsh = k * 8u;
        uint32_t v4 = vop & ~3u;
        uint32_t c = dsh ? acc << (32u - dsh) : 0u;                // the carried bytes, moved to the top of a word
        uint32_t rem = n;
        // Up to four words per round trip when the source allows reading that far ahead (always for the input ring and
        // global memory; in the output ring the words read must lie below the write cursor).  When neither ring wraps inside
        // the group, every access is base + constant (LDS / STS with immediate offsets: no address arithmetic per word).
        uint32_t full = (rem - 1u) >> 2;                           // whole 4-byte groups before the last group (1..4 bytes)
        const bool ahead = K != 1 || vop - s >= 24u;
        while (full > 0) {
            const uint32_t cnt = ahead ? (full < 4u ? full : 4u) : 1u;
            uint32_t q[4];
            const uint32_t so = K == 0 ? (sp & IMASK) : (sp & (uint32_t)(GEO::OUT - 4));
            const uint32_t dof = v4 & (uint32_t)(GEO::OUT - 4);
            const bool flat = ahead && dof <= (uint32_t)(GEO::OUT - 16) && (K == 2 || so <= (uint32_t)((K == 0 ? GEO::IN : GEO::OUT) - 20));
            if (flat) {
                if (K == 2) {
#pragma unroll
                    for (int i = 0; i < 4; i++) q[i] = simt::ldg_u32(gbase + sp + 4u + 4u * i);
                } else {
                    const simt::smem_ref sr = K == 0 ? ir : ow;
                    const uint32_t sb = K == 0 ? so : l4 + so;
#pragma unroll
                    for (int i = 0; i < 4; i++) q[i] = simt::lds_u32(sr, sb + 4u + 4u * i);
                }
                const uint32_t db = l4 + dof;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t x = simt::funnel_r(lo, q[i], ssh);
                    if ((uint32_t)i < cnt) { simt::sts_u32(ow, db + 4u * i, simt::funnel_l(c, x, dsh)); lo = q[i]; c = x; }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) if ((uint32_t)i < cnt) q[i] = load_src(kind, sp + 4u + 4u * i);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if ((uint32_t)i < cnt) {
                        const uint32_t x = simt::funnel_r(lo, q[i], ssh);
                        simt::sts_u32(ow, oword(l4, v4 + 4u * i), simt::funnel_l(c, x, dsh));
                        lo = q[i]; c = x;
                    }
                }
            }
            sp += 4u * cnt; v4 += 4u * cnt; rem -= 4u * cnt; full -= cnt;
        }
        {                                                          // the last group: 1..4 bytes
            const uint32_t hi = load_src(kind, sp + 4);
            uint32_t x = simt::funnel_r(lo, hi, ssh);
            if (rem < 4) x &= (1u << (8u * rem)) - 1u;
            const uint32_t carried = dsh ? c >> (32u - dsh) : 0u;
            const uint32_t w = carried | (x << dsh);
            simt::sts_u32(ow, oword(l4, v4), w);
            if (k + rem >= 4) {                                    // the word is finished; what spilled over starts the next one
                acc = dsh ? x >> (32u - dsh) : 0u;
                if (k + rem > 4) simt::sts_u32(ow, oword(l4, v4 + 4), acc);      // written through
            } else acc = w;
        }
        vop += n;
    };
    // The same for a match whose source is at least 4 bytes back and lies, word by word, either in the output ring
    // (positions >= rmin: still valid when the append ends) or in global memory (written out earlier).  Every lane of the
    // warp can take this one path whatever its offset: no read-ahead, two fresh source words per group.
    auto append_mixed = [&](uint32_t s, uint32_t n, int rmin) {
        const uint32_t ssh = (s & 3u) * 8u;
        uint32_t sp = s & ~3u;
        const uint32_t k = vop & 3u, dsh = k * 8u;
        uint32_t v4 = vop & ~3u;
        uint32_t c = dsh ? acc << (32u - dsh) : 0u;
        uint32_t rem = n;
        auto ld = [&](uint32_t p4) -> uint32_t { return (int)p4 >= rmin ? simt::lds_u32(ow, oword(l4, p4)) : simt::ldg_u32(gbase + p4); };
        // closer than 8 bytes, the next group reads bytes this group produced: the part of them that spills into the next
        // word is written through at once (further back, the carried bytes are stored with the next group early enough)
        const bool wt = dsh != 0u && vop - s < 8u;
        while (rem > 4) {
            const uint32_t x = simt::funnel_r(ld(sp), ld(sp + 4), ssh);
            simt::sts_u32(ow, oword(l4, v4), simt::funnel_l(c, x, dsh));
            if (wt) simt::sts_u32(ow, oword(l4, v4 + 4), x >> (32u - dsh));
            c = x; sp += 4; v4 += 4; rem -= 4;
        }
        {
            uint32_t x = simt::funnel_r(ld(sp), ld(sp + 4), ssh);
            if (rem < 4) x &= (1u << (8u * rem)) - 1u;
            const uint32_t carried = dsh ? c >> (32u - dsh) : 0u;
            const uint32_t w = carried | (x << dsh);
            simt::sts_u32(ow, oword(l4, v4), w);
            if (k + rem >= 4) {
                acc = dsh ? x >> (32u - dsh) : 0u;
                if (k + rem > 4) simt::sts_u32(ow, oword(l4, v4 + 4), acc);
            } else acc = w;
        }
        vop += n;
    };
    // one byte (overlapping matches closer than 4 bytes; sources that straddle the ring's valid range)
    auto append_byte = [&](uint32_t b) {
        const uint32_t k = vop & 3u;
        acc |= b << (8u * k);
        simt::sts_u32(ow, oword(l4, vop & ~3u), acc);
        if (k == 3) acc = 0;
        vop++;
    };
    auto ring_byte = [&](uint32_t v) -> uint32_t { return (simt::lds_u32(ow, oword(l4, v & ~3u)) >> (8u * (v & 3u))) & 255u; };

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
                    a0 = (uint32_t)((uintptr_t)dst & 127); gbase = dst - a0;
                    g128 = (int32_t)(((intptr_t)gbase - (intptr_t)dbase128) >> 7);
                    ip = 0; vop = fpos = slo = a0; acc = 0; phase = 0; coop = 0; need_all = false; done = false;
                    simt::cp_async_commit(); simt::cp_async_wait<0>();     // (units the previous block requested but never read)
                    ifill = 0; refill(0);
                    active = true; fresh = true;
                }
            }
        }
        if (!simt::ballot(FULL, active)) break;
        // everything requested up to the last commit but one has landed; a fresh block waits for its first units
        simt::cp_async_commit();
        if (GEO::DEPTH == 0 || simt::ballot(FULL, fresh)) { simt::cp_async_wait<0>(); f1 = iland = ifill; }
        else { simt::cp_async_wait<1>(); iland = f1; f1 = ifill; }  // (all groups but the one just committed are complete: everything below the previous commit's ifill)
        fresh = false;

        for (int rep = 0; rep < GEO::SEQS; rep++) {
          // (a further sequence in the same iteration only while the unflushed bytes plus one more sequence -- and the word
          // that is written through behind it -- still fit the ring)
          const bool go = active && !need_all && !done && coop == 0 &&
                          (rep == 0 || (vop - fpos) + (uint32_t)(GEO::MAXL + GEO::MAXM + 8) <= (uint32_t)GEO::OUT);
          bool mcan = false, ring8 = false; uint32_t m_s = 0, m_n = 0; int m_q = 0, m_rmin = 0;
          if (go) {
            // ---------------- header + literals ----------------
            // The ordinary sequence -- far from the end of both buffers, its bytes resident, runs a lane copies itself -- needs
            // none of the end-of-block tests of the reference: one guard replaces them (everything else takes the general
            // code below, which restates them one by one).
            bool fast_h = false;
            if (phase == 0 && (uint32_t)(ip + GEO::LOOK) + skew <= iland && ip + GEO::LOOK + 16 <= isize &&
                (int)(vop - a0) + GEO::MAXL + 12 <= cap) {
                const uint32_t t = ib(ip);
                uint32_t L = t >> 4; int p = ip + 1;
                if (L == 15) { L += ib(p); p++; }                           // (an extension byte of 255 makes L > MAXL: general code)
                if (L <= (uint32_t)GEO::MAXL) {
                    token = t;
                    if (L) append(LpbKind<0>(), (uint32_t)p, L);
                    ip = p + (int)L; phase = 1; fast_h = true;
                }
            }
            if (phase == 0 && !fast_h) {
                const int op = (int)(vop - a0);
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
                            else if (L > 0) { append(LpbKind<0>(), (uint32_t)ip, (uint32_t)L); ip += L; }
                            phase = last ? 2 : 1;
                        }
                    }
                }
            }
            // ---------------- offset + match ----------------
            // ---------------- offset + match: the ordinary case is parsed here and copied below, on ONE path for the warp ------
            if (!done && coop == 0 && phase == 1 && (uint32_t)(ip + 8) + skew <= iland && ip + 16 <= isize &&
                (int)(vop - a0) + GEO::MAXM + 5 <= cap) {
                const uint32_t off = ib(ip) | (ib(ip + 1) << 8);
                uint32_t M = token & 15u; int q = ip + 2;
                if (M == 15) { M += ib(q); q++; }
                const uint32_t n = M + 4u, sv = vop - off;
                // ring words: positions >= rmin (valid until the append ends); anything below must be in global memory already
                int rmin = (int)(vop + n + 4u) - GEO::OUT; if (rmin < (int)slo) rmin = (int)slo;
                rmin = (rmin + 3) & ~3;
                if (off >= 4u && n <= (uint32_t)GEO::MAXM && off <= vop - a0 && ((int)(sv & ~3u) >= rmin || (int)fpos >= rmin)) {
                    mcan = true; m_s = sv; m_n = n; m_q = q; m_rmin = rmin;
                    ring8 = off >= 8u && (int)(sv & ~3u) >= rmin;
                }
            }
          }
          {
            const uint32_t anym = simt::ballot(FULL, mcan), mixed = simt::ballot(FULL, mcan && !ring8);
            if (anym) {
                if (!mixed) { if (mcan) { append(LpbKind<1>(), m_s, m_n); ip = m_q; phase = 0; } }
                else if (mcan) { append_mixed(m_s, m_n, m_rmin); ip = m_q; phase = 0; }
            }
          }
          if (go) {
            const bool fast_m = mcan;
            if (!done && coop == 0 && phase == 1 && !fast_m) {
                const int op1 = (int)(vop - a0);
                fetch(ip, ip + 8 < isize ? ip + 8 : isize - 1);
                const uint32_t off = ib(ip) | (ib(ip + 1) << 8);            // :862 / :982 (ip + 2 <= isize was checked with the literals)
                int q = ip + 2;
                if (off == 0 || off > (uint32_t)op1) { result = -q - 1; done = true; }          // :863 / :983 (offset 0 rejected by design)
                else {
                    int M = (int)(token & 15);
                    if (M == 15) {                                          // :866 / :986-999
                        uint32_t s = 255;
                        if (KNOWN) { do { if (q >= isize || M > LEN_LIMIT) { M = -1; break; } s = ibx(q++); M += (int)s; } while (s == 255); }
                        else       { while (q < isize - 6 && M <= LEN_LIMIT) { s = ibx(q++); M += (int)s; if (s != 255) break; } }
                        if (M < 0 || M > LEN_LIMIT) { result = -q - 1; done = true; }
                    }
                    if (!done && M > cap - op1) { result = -q - 1; done = true; }
                    if (!done) {
                        const int n = M + 4;
                        const uint32_t s = vop - off;
                        if (op1 + n > cap - 5) { result = -q - 1; done = true; }                // :893 / :1025 -- the last 5 bytes are literals
                        else if (n > GEO::MAXM) { coop = 2; clen = n; coff = off; ip = q; }
                        else if (off <= (uint32_t)GEO::WIN) {
                            if (s >= slo && off >= 8) append(LpbKind<1>(), s, (uint32_t)n);
                            else {
                                // overlap closer than 8, or a source that straddles the start of the ring's valid range (just
                                // behind a run the warp copied -- what lies below is in global memory): byte by byte, in order
                                for (int i = 0; i < n; i++) {
                                    const uint32_t p = s + (uint32_t)i;
                                    append_byte(p >= slo ? ring_byte(p) : (uint32_t)simt::ldg_u8(gbase + p));
                                }
                            }
                            ip = q; phase = 0;
                        } else if (s + (uint32_t)n > fpos) {
                            // a source behind the window that is not in global memory yet: write this lane out completely
                            // first and take the sequence again in the next iteration
                            need_all = true;
                        } else {
                            append(LpbKind<2>(), s, (uint32_t)n);         // off > WIN >= n: no overlap with the destination
                            ip = q; phase = 0;
                        }
                    }
                }
            }
            if (!done && coop == 0 && phase == 2) {                        // the last literals are in: finished
                result = KNOWN ? ip : (int)(vop - a0);
                done = true;
            }
            if (done || coop) need_all = true;                              // the end of the block / the warp's copy: everything out first
          }
        }
        // ---------------- finished chunks -> global memory, four lanes per step --------------------------------------------
        // A step serves, per quarter warp, the lowest lane of that quarter that wants it: its eight lanes read 16 bytes each
        // of that lane's ring (transposed: conflict free) and store 128 contiguous bytes.  Steps run when enough lanes have
        // a finished chunk (lock-step data: all of them at once), when a lane's backlog nears the ring's capacity, or when
        // a lane must be written out completely.
        {
            const uint32_t pend = active ? vop - fpos : 0u;
            const uint32_t ready = simt::ballot(FULL, active && (vop >> 7) > (fpos >> 7));
            const uint32_t urgent = simt::ballot(FULL, active && (pend >= (uint32_t)GEO::HIGH || (need_all && pend > 0)));
            if (urgent || simt::popc(ready) >= 16) {
                const uint32_t qbase = (uint32_t)lane & 24u, j = (uint32_t)lane & 7u;
                for (;;) {
                    const bool want = active && (need_all ? vop > fpos : (vop >> 7) > (fpos >> 7));
                    const uint32_t m = simt::ballot(FULL, want);
                    if (!m) break;
                    const uint32_t mq = (m >> qbase) & 255u;
                    // the lane this quarter serves (none: itself, nothing to do).  Quarter q starts looking at its lane q: when
                    // all lanes want (lock step), the four lanes served in one step differ mod 4, so the four quarters' reads
                    // (each spread over the eight banks of one residue class mod 4) do not collide
                    const uint32_t qi = qbase >> 3;
                    const uint32_t rot = ((mq >> qi) | (mq << (8u - qi))) & 255u;
                    const int k = mq ? (int)(qbase + (((uint32_t)simt::ffs(rot) - 1u + qi) & 7u)) : lane;
                    // two shuffles describe the served lane's chunk: its 128-byte unit index from the batch's (rounded) base, and
                    // (ring offset of fpos) << 10 | unflushed bytes -- everything below is chunk relative
                    const int32_t kunit = (int32_t)simt::shfl(FULL, (uint32_t)(g128 + (int32_t)(fpos >> 7)), k);
                    const uint32_t kpk = simt::shfl(FULL, ((fpos & (uint32_t)(GEO::OUT - 1)) << 10) | (vop - fpos), k);
                    if (mq) {
                        const uint32_t fo = kpk >> 10, pend_k = kpk & 1023u;
                        const uint32_t klo = fo & 127u;                               // first byte of the chunk still to be written
                        uint32_t khi = klo + pend_k; if (khi > 128u) khi = 128u;      // (only a lane that is written out completely ends inside a chunk)
                        uint8_t* const kg = dbase128 + (intptr_t)kunit * 128;
                        const uint32_t kr = (uint32_t)k * (uint32_t)(GEO::OUT + 4) + (fo & ~127u);   // the chunk in lane k's ring (no wrap inside a chunk)
                        const uint32_t b = 16u * j;
                        if (b >= klo && b + 16u <= khi) {
                            uint4 v;
                            v.x = simt::lds_u32(ow, kr + b); v.y = simt::lds_u32(ow, kr + b + 4);
                            v.z = simt::lds_u32(ow, kr + b + 8); v.w = simt::lds_u32(ow, kr + b + 12);
                            simt::stg_v4(kg + b, v);
                        } else if (b + 16u > klo && b < khi) {                      // the edges of the range: words, then bytes
                            for (uint32_t t = 0; t < 16; t += 4) {
                                const uint32_t ws = b + t;
                                if (ws + 4 <= klo || ws >= khi) continue;
                                const uint32_t w = simt::lds_u32(ow, kr + ws);
                                if (ws >= klo && ws + 4 <= khi) simt::stg_u32(kg + ws, w);
                                else for (uint32_t e = 0; e < 4; e++) if (ws + e >= klo && ws + e < khi) simt::stg_u8(kg + ws + e, (uint8_t)(w >> (8u * e)));
                            }
                        }
                        if (lane == k) fpos += khi - klo;
                    }
                }
                simt::syncwarp(FULL);                                       // the stores above -> the served lanes' later loads of their own output
            }
        }
        if (active && need_all && !done && coop == 0) need_all = false;     // (a far match waited for its source: take the sequence again)
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
            if (kind == 1) { InputSrc sp{s}; group_copy<32, false>(d, sp, n, lane, FULL); }
            else group_copy_match<32>(d, o, n, lane, FULL);
            simt::syncwarp(FULL);                                           // the copy -> lane k's later loads
        }
        if (active && coop != 0) {
            if (coop == 1) {                                                // the input ring jumps behind the run (drains what is in flight first)
                ip += clen;
                if (phase == 1) fetch(ip, ip + 8 < isize ? ip + 8 : isize - 1);
            } else phase = 0;
            vop += (uint32_t)clen; fpos = slo = vop; acc = 0;               // the ring holds nothing of the block any more ...
            if (vop & 3u) {                                                 // ... except the started word, which later appends complete
                acc = simt::ldg_u32(gbase + (vop & ~3u)) & ((1u << (8u * (vop & 3u))) - 1u);
                simt::sts_u32(ow, oword(l4, vop & ~3u), acc);
                slo = vop & ~3u;
            }
            coop = 0; need_all = false;
            if (phase == 2) { result = KNOWN ? ip : (int)(vop - a0); done = true; }
        }
        if (active && done) { simt::stg_u32(a.out_len + blk, (uint32_t)result); active = false; done = false; }
        // ---------------- ask for the units the next iterations will read ----------------
        if (active) refill((uint32_t)ip + skew);
    }
    simt::cp_async_commit(); simt::cp_async_wait<0>();
}

}  // namespace lz4b200
