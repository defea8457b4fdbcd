// lz4_decode.cuh -- batched LZ4 block decoder: G lanes (a warp or a sub-warp group of 16 / 8 lanes) own one block.
//
// Replaces LZ4_uncompress (original/lz4.c:812-914) and LZ4_uncompress_unknownOutputSize (:916-1044), i.e. lz4net's
// LZ4_uncompress_safe64 / _unknownOutputSize_safe64 (src/LZ4ps/LZ4Codec.Safe64.Dirty.cs:533-659,665-798).
//
// Data flow per block:
//   compressed stream  --cp.async.bulk (TMA engine, 16 B aligned chunks, mbarrier completion)-->  a 4-slot ring in
//   shared memory owned by the group  -->  tokens / lengths / offsets parsed from the ring (LDS, group-uniform);
//   literals: ring -> global; matches: global (own earlier output, L1/L2 hits) -> global.
//
// The kernel is instruction-issue bound on token-dense data (one warp instruction serves one sequence of one block),
// so the per-sequence work is split in two:
//   * a FAST path for the common short sequence (<= 32 literals, <= 64 match bytes, far from the ends of both
//     buffers): no per-byte ring checks (the ring keeps a 64-byte mirror of its head behind its tail, so a header and
//     its literals are always contiguous), one availability test per sequence, byte-wide predicated copies;
//   * the CAREFUL path (every other sequence, and always the last ones): the fully checked, piecewise-through-the-ring
//     implementation with 128-bit copies for long runs.
// Sub-warp groups (G = 16 or 8) let one warp instruction serve 2 or 4 blocks, which is what raises throughput on short
// sequences; G = 32 is best for long literal runs / long matches (incompressible or RLE data).
//
// Accept / reject decisions are those of the reference's 64-bit flavour (see oracle/lz4_oracle.c, pinned against the
// reference sources); a malformed stream yields a negative result and never an access outside [src, src+isize) or
// [dst, dst+cap).
#pragma once
#include "simt.cuh"
#include "lz4_copy.cuh"

#ifndef LZ4B200_DEC_UNIFIED_MATCH
// Staged variant, match source older than the sequence's own literals: 1 = one predicated copy loop whether the source is
// still staged, already in global memory, or both (fewer distinct paths for the groups of a warp to diverge over:
// ETEXT +3.5 % at 4 lanes, +6 % at 8; E50 unchanged); 0 = a separate loop for sources that are entirely in global memory.
#define LZ4B200_DEC_UNIFIED_MATCH 1
#endif

namespace lz4b200 {

constexpr int DEC_SLOTS = 4;
constexpr int DEC_MIRROR = 64;                                  // bytes of the ring head replicated behind its tail
constexpr int DEC_AHEAD = 64;                                   // the fast path may read this far past the cursor

template <int G> struct DecGeom {
    static constexpr int LOG_CHUNK = G == 32 ? 10 : (G == 16 ? 9 : (G == 8 ? 8 : 7));     // 1 KiB / 512 / 256 / 128 B bulk copies
    static constexpr int CHUNK = 1 << LOG_CHUNK;
    static constexpr int RING = CHUNK * DEC_SLOTS;
};

template <int G>
struct alignas(16) DecRing {
    uint8_t buf[DecGeom<G>::RING + DEC_MIRROR];
    simt::mbar_t bar[DEC_SLOTS];
};

// Per-group view of the compressed stream through the ring.
template <int G>
struct DecStream {
    static constexpr int LOG_CHUNK = DecGeom<G>::LOG_CHUNK, CHUNK = DecGeom<G>::CHUNK, RING = DecGeom<G>::RING;
    DecRing<G>* ring;
    const uint8_t* abase;        // src rounded down to 16 B
    uint32_t skew;               // src - abase
    uint32_t total;              // round_up16(skew + isize): bytes that may be fetched
    int nchunks;
    int issued, ready;           // chunks whose bulk copy has been launched / has completed
    uint32_t gbase;              // chunks this group has ever launched before this block: chunk k of the block is global
                                 // chunk gbase+k, lives in slot (gbase+k)&3 and completes phase ((gbase+k)>>2)&1 of it
    uint32_t rot;                // ring offset of stream byte i is (i + rot) & (RING-1), rot = skew + gbase*CHUNK
    int lane; uint32_t gmask;

    SIMT_MEM void begin(const uint8_t* src, int isize)
    {
        skew = (uint32_t)((uintptr_t)src & 15);
        abase = src - skew;
        total = (skew + (uint32_t)isize + 15u) & ~15u;
        nchunks = (int)((total + CHUNK - 1) >> LOG_CHUNK);
        issued = ready = 0;
        rot = (skew + (gbase << LOG_CHUNK)) & (uint32_t)(RING - 1);
        fill(0);
    }
    // launch every chunk that may be resident while the read cursor is in chunk c (chunks c .. c+3)
    SIMT_MEM void fill(int c)
    {
        int want = c + DEC_SLOTS; if (want > nchunks) want = nchunks;
        if (issued >= want) return;
        simt::syncwarp(gmask);                                     // every lane is done with the slots being recycled
        if (lane == 0) {
            for (int k = issued; k < want; k++) {
                const uint32_t b0 = (uint32_t)k << LOG_CHUNK;
                uint32_t nb = total - b0; if (nb > (uint32_t)CHUNK) nb = CHUNK;
                const int s = (int)((gbase + (uint32_t)k) & (DEC_SLOTS - 1));
                const uint32_t mb = (s == 0) ? (nb < (uint32_t)DEC_MIRROR ? nb : (uint32_t)DEC_MIRROR) : 0u;
                simt::mbar_expect(&ring->bar[s], nb + mb);
                simt::bulk_copy(ring->buf + s * CHUNK, abase + b0, nb, &ring->bar[s]);
                if (mb) simt::bulk_copy(ring->buf + RING, abase + b0, mb, &ring->bar[s]);   // head mirror behind the tail
            }
        }
        issued = want;
    }
    SIMT_MEM void wait_chunk(int c)                               // make chunks <= c readable
    {
        while (ready <= c) {
            const uint32_t g = gbase + (uint32_t)ready;
            simt::mbar_wait(&ring->bar[g & (DEC_SLOTS - 1)], (g >> 2) & 1);
            ready++;
        }
    }
    SIMT_MEM int chunk_of(uint32_t p) const { return (int)((p + skew) >> LOG_CHUNK); }
    // one byte at stream position p (the read cursor): group-uniform broadcast LDS, fully checked
    SIMT_MEM uint32_t byte_at(uint32_t p)
    {
        const int c = chunk_of(p);
        if (c + DEC_SLOTS > issued) fill(c);
        if (c >= ready) wait_chunk(c);
        return ring->buf[(p + rot) & (RING - 1)];
    }
    // Fast-path window: after this call every stream byte in [ip, limit) is resident and, thanks to the mirror,
    // contiguous from ring offset (ip + rot) & (RING-1); limit >= ip + DEC_AHEAD unless the stream ends first.
    // Only the chunk that the next DEC_AHEAD bytes reach into is waited for, so a chunk is first touched about three
    // chunk-times after its bulk copy was launched (fill keeps chunks c .. c+3 in flight).
    // `extra` > 0 (<= CHUNK) looks that much further ahead, so that groups sliding together stay together for a while.
    SIMT_MEM int window(int ip, int isize, int extra = 0)
    {
        const int c = chunk_of((uint32_t)ip);
        fill(c);
        int last = chunk_of((uint32_t)(ip + DEC_AHEAD + extra)); if (last > nchunks - 1) last = nchunks - 1;
        wait_chunk(last);
        const int lim = (int)(((uint32_t)(last + 1) << LOG_CHUNK) - skew);
        return lim < isize ? lim : isize;
    }
    // finish the block: every launched copy must have landed before the slots / barriers are reused
    SIMT_MEM void end()
    {
        if (issued > 0) wait_chunk(issued - 1);
        gbase += (uint32_t)issued;
        simt::syncwarp(gmask);
    }
    // copy n stream bytes starting at p to dst; advances through the ring piecewise
    SIMT_MEM void copy_out(uint8_t* dst, uint32_t p, uint32_t n)
    {
        while (n) {
            const int c = chunk_of(p);
            if (c + DEC_SLOTS > issued) fill(c);
            int lim = c + DEC_SLOTS; if (lim > issued) lim = issued;
            const uint32_t avail = ((uint32_t)lim << LOG_CHUNK) - skew - p;   // resident bytes from p on
            const uint32_t m = n < avail ? n : avail;
            wait_chunk(chunk_of(p + m - 1));
            RingSrc<RING> s{ring->buf, p + rot};
            group_copy<G, false>(dst, s, m, lane, gmask);
            dst += m; p += m; n -= m;
        }
    }
};

struct DecCursor { int ip, op; };

// The careful path: one fully checked sequence.  Returns 0 = continue, 1 = finished (*result set), -1 = malformed.
template <int G, bool KNOWN>
SIMT_DEV int decode_careful(DecStream<G>& st, DecCursor& cur, int isize, uint8_t* dst, int cap, int* result)
{
    constexpr int LEN_LIMIT = 0x3FFFFFFF;                          // a run this long cannot fit any int-sized buffer
    int ip = cur.ip, op = cur.op;
    if (ip >= isize) { *result = -ip - 1; return -1; }
    const uint32_t token = st.byte_at(ip++);
    int L = (int)(token >> 4);
    if (L == 15) {                                                 // :843 / :959-963
        uint32_t s = 255;
        if (KNOWN) { do { if (ip >= isize || L > LEN_LIMIT) { L = -1; break; } s = st.byte_at(ip++); L += (int)s; } while (s == 255); }
        else       { while (ip < isize && s == 255 && L <= LEN_LIMIT) { s = st.byte_at(ip++); L += (int)s; } }
        if (L < 0 || L > LEN_LIMIT) { *result = -ip - 1; return -1; }
    }
    // a run longer than the rest of either buffer fails every end test below; rejecting it here keeps op + L and ip + L
    // inside int for blocks of any size
    if (L > isize - ip || L > cap - op) { *result = -ip - 1; return -1; }
    int end = op + L;
    bool last;
    if (KNOWN) last = end > cap - 8;                               // :847
    else       last = end > cap - 12 || ip + L > isize - 8;        // :968
    if (last) {
        const bool ok = KNOWN ? (end == cap && ip + L <= isize)    // :849-857
                              : (end <= cap && ip + L == isize);   // :974-975
        if (!ok) { *result = -ip - 1; return -1; }
        st.copy_out(dst + op, ip, L);
        ip += L; op = end;
        *result = KNOWN ? ip : op;
        return 1;
    }
    if (ip + L + 2 > isize) { *result = -ip - 1; return -1; }      // (KNOWN: bounded reads; !KNOWN: implied by :968)
    st.copy_out(dst + op, ip, L);
    ip += L; op = end;
    uint32_t off = st.byte_at(ip);                                 // :862 / :982 (two sequenced reads: the ring may turn over)
    off |= st.byte_at(ip + 1) << 8; ip += 2;
    if (off == 0 || off > (uint32_t)op) { *result = -ip - 1; return -1; }          // :863 / :983 (offset 0 rejected by design)
    int M = (int)(token & 15);
    if (M == 15) {                                                 // :866 / :986-999
        uint32_t s = 255;
        if (KNOWN) { do { if (ip >= isize || M > LEN_LIMIT) { M = -1; break; } s = st.byte_at(ip++); M += (int)s; } while (s == 255); }
        else       { while (ip < isize - 6 && M <= LEN_LIMIT) { s = st.byte_at(ip++); M += (int)s; if (s != 255) break; } }
        if (M < 0 || M > LEN_LIMIT) { *result = -ip - 1; return -1; }
    }
    if (M > cap - op) { *result = -ip - 1; return -1; }            // (keeps op + M + 4 inside int; fails :893 / :1025 anyway)
    end = op + M + 4;
    if (end > cap - 5) { *result = -ip - 1; return -1; }           // :893 / :1025 -- the last 5 bytes are literals
    simt::syncwarp(st.gmask);                                      // literal (and earlier match) stores -> match loads
    group_copy_match<G>(dst + op, off, (uint32_t)(M + 4), st.lane, st.gmask);
    cur.ip = ip; cur.op = end;
    return 0;
}

// k mod off for small k (< 128) without an integer division
SIMT_DEV uint32_t small_mod(uint32_t k, uint32_t off, float rcp)
{
    const uint32_t q = (uint32_t)(((float)k + 0.5f) * rcp);
    return k - q * off;
}

// Fast-path byte movers: n <= MAX bytes, lane i moves bytes i, i+G, i+2G, ...  `d` and `s` are PER-LANE pointers
// (already offset by the lane), so every step is one predicated load + one predicated store with an immediate offset.
template <int G, int MAX, class LD>
SIMT_DEV void fast_steps(uint8_t* d, const uint8_t* s, uint32_t n, uint32_t lane, LD ld)
{
    if (lane < n) simt::stg_u8(d, ld(s));
    if (MAX > G && n > (uint32_t)G) {
#pragma unroll
        for (int i = 1; i < MAX / G && i < 4; i++)
            if (lane + i * G < n) simt::stg_u8(d + i * G, ld(s + i * G));
        if (MAX > 4 * G && n > 4u * G) {
#pragma unroll
            for (int i = 4; i < MAX / G; i++)
                if (lane + i * G < n) simt::stg_u8(d + i * G, ld(s + i * G));
        }
    }
}

// Decode one block.  KNOWN: cap is the exact decoded size, result = bytes read (LZ4_uncompress).
// !KNOWN: cap is the capacity, result = bytes written (LZ4_uncompress_unknownOutputSize).  < 0 = malformed.
template <int G, bool KNOWN>
SIMT_DEV int decode_block(DecStream<G>& st, const uint8_t* src, int isize, uint8_t* dst, int cap)
{
    constexpr int RING = DecGeom<G>::RING;
    constexpr int FAST_L = 32, FAST_M = 64;                        // longest literal run / match the fast path takes
    const uint32_t lane = (uint32_t)st.lane; const uint32_t gmask = st.gmask;
    if (isize <= 0 || cap < 0) return -1;                          // original/lz4.c:949; a block has >= 1 token
    st.begin(src, isize);
    const uint8_t* const rb = st.ring->buf;
    const uint32_t rot = simt::keep(st.rot);
    uint8_t* const dl = simt::keep(dst + lane);                    // this lane's column of the output (one register pair)
    DecCursor cur{0, 0};
    int result = 0;
    // fast path preconditions: DEC_AHEAD readable bytes after ip, and every end test of the reference trivially passes
    const int out_fast = cap - (FAST_L + FAST_M + 16);
    int in_fast = st.window(0, isize) - DEC_AHEAD;
    for (;;) {
        if (cur.ip > in_fast) in_fast = st.window(cur.ip, isize) - DEC_AHEAD;      // slide the window (about once per chunk)
        if (cur.ip <= in_fast && cur.op <= out_fast) {
            // ---------------- fast path: the sequence header and its literals are contiguous at h ----------------
            const uint8_t* const h = rb + (((uint32_t)cur.ip + rot) & (RING - 1));
            // branch-free header parse: both possible length bytes are loaded unconditionally (they are inside the
            // 64-byte window either way) and selected, so the loads of one header overlap instead of chaining
            const uint32_t token = h[0], e1 = h[1];
            const bool lx = (token >> 4) == 15;
            const uint32_t L = lx ? 15 + e1 : token >> 4, hdr = lx ? 2u : 1u;
            if (L <= (uint32_t)FAST_L) {                            // (an e1 of 255 gives L >= 270 and falls to the careful path)
                const uint8_t* const q = h + hdr + L;
                const uint32_t off = q[0] | ((uint32_t)q[1] << 8), e2 = q[2];
                const bool mx = (token & 15) == 15;
                const uint32_t M = (mx ? 15 + e2 : token & 15) + 4, adv = hdr + L + (mx ? 3u : 2u);
                if (M <= (uint32_t)FAST_M) {                        // (255 -> M >= 274 -> careful path)
                    const uint32_t opl = (uint32_t)cur.op + L;
                    if (off - 1u >= opl) { result = -(cur.ip + (int)adv) - 1; break; }            // :863 / :983 (0 or too far)
                    uint8_t* const d = dl + cur.op;
                    const uint8_t* const lit = h + hdr;            // literal j of this sequence is ring byte lit[j]
                    auto lds = [](const uint8_t* p) { return *p; };
                    auto ldg = [](const uint8_t* p) { return simt::ldg_u8(p); };
                    fast_steps<G, FAST_L>(d, lit + lane, L, lane, lds);
                    uint8_t* const m = d + L;
                    if (off <= L) {
                        // The match starts inside this sequence's own literal run: every source byte is (a periodic
                        // repetition of) literal bytes that are still in the ring -- no global load, no store->load
                        // round trip through L2, no group synchronisation.
                        const uint8_t* const s = lit + (L - off);
                        if (off >= M) fast_steps<G, FAST_M>(m, s + lane, M, lane, lds);
                        else {
                            const float rcp = 1.0f / (float)off;
#pragma unroll 1
                            for (uint32_t k = lane; k < M; k += G) simt::stg_u8(dst + opl + k, s[small_mod(k, off, rcp)]);
                        }
                    } else {
                        simt::syncwarp(gmask);                     // earlier stores of other lanes -> match loads
                        if (off >= M) fast_steps<G, FAST_M>(m, m - off, M, lane, ldg);
                        else {                                     // overlapping: period `off`, only finished bytes are read
                            const float rcp = 1.0f / (float)off;
                            const uint8_t* const base = dst + opl - off;
#pragma unroll 1
                            for (uint32_t k = lane; k < M; k += G) simt::stg_u8(dst + opl + k, simt::ldg_u8(base + small_mod(k, off, rcp)));
                        }
                    }
                    cur.ip += (int)adv; cur.op = (int)(opl + M);
                    continue;
                }
            }
        }
        // ---------------- careful path ----------------
        const int r = decode_careful<G, KNOWN>(st, cur, isize, dst, cap, &result);
        if (r != 0) break;
    }
    st.end();
    return result;
}

// ---------------------------------------------------------------------------------------------------------------------
// Output-staged variant.  The fast path appends its output to a small per-group buffer in shared memory instead of
// issuing byte-wide global stores; the buffer is flushed with 128-bit, 16-byte-aligned, fully coalesced stores.
// Per-sequence global traffic drops from (litLen + matchLen) byte stores in 32/G different cache lines to 1/16 of a
// 128-bit store per byte, and a match whose source is still staged is served from shared memory.
//   ob index i  <->  output position ostart + i;   (dst + ostart) is 16-byte aligned;   valid bytes are [olo, ohi).
// Everything below output position ostart + olo is in global memory and visible to the whole group (each flush and each
// careful-path sequence ends with a group synchronisation).
// ---------------------------------------------------------------------------------------------------------------------
template <int G> struct DecStageGeom {
    static constexpr int FLUSH_AT = G <= 8 ? 256 : 512;           // flush once this many bytes are staged.  (A flush by one group stalls the
                                                                   // other groups of its warp; 1 KiB for G = 8 was measured slower: it costs a third of the resident warps.)
    static constexpr int SIZE = FLUSH_AT + 32 + 64 + 16;           // + one more fast sequence + slack
};
template <int G> struct alignas(16) DecStage { uint8_t ob[DecStageGeom<G>::SIZE]; };

template <int G, int MAX, class LD>
SIMT_DEV void fast_steps_smem(uint8_t* d, const uint8_t* s, uint32_t n, uint32_t lane, LD ld)
{
    if (lane < n) *d = ld(s);
    if (MAX > G && n > (uint32_t)G) {
#pragma unroll
        for (int i = 1; i < MAX / G && i < 4; i++)
            if (lane + i * G < n) d[i * G] = ld(s + i * G);
        if (MAX > 4 * G && n > 4u * G) {
#pragma unroll
            for (int i = 4; i < MAX / G; i++)
                if (lane + i * G < n) d[i * G] = ld(s + i * G);
        }
    }
}

template <int G, bool KNOWN>
SIMT_DEV int decode_block_staged(DecStream<G>& st, DecStage<G>* stage, const uint8_t* src, int isize, uint8_t* dst, int cap)
{
    constexpr int RING = DecGeom<G>::RING;
    constexpr int FAST_L = 32, FAST_M = 64;
    constexpr int FLUSH_AT = DecStageGeom<G>::FLUSH_AT;
    const uint32_t lane = (uint32_t)st.lane; const uint32_t gmask = st.gmask;
    if (isize <= 0 || cap < 0) return -1;
    st.begin(src, isize);
    const uint8_t* const rb = st.ring->buf;
    uint8_t* const ob = stage->ob;
    const uint32_t rot = simt::keep(st.rot);
    DecCursor cur{0, 0};
    int result = 0;
    int ostart, olo, ohi;
    auto stage_reset = [&](int op) { const int a = (int)(((uintptr_t)dst + (uintptr_t)op) & 15); ostart = op - a; olo = ohi = a; };
    // write out the staged bytes.  all == false: whole 16-byte vectors only, the tail (< 16 bytes) moves to the front.
    // all == true: the tail goes out as bytes too and the stage is left empty (before the careful path / at the end).
    auto flush = [&](bool all) {
        simt::syncwarp(gmask);                                     // every lane's staged bytes are in place
        uint8_t* const g = dst + ostart;
        const int nv = ohi >> 4;
        for (int v = (int)lane; v < nv; v += G) {
            if (v == 0 && olo > 0) { for (int j = olo; j < 16; j++) simt::stg_u8(g + j, ob[j]); }   // first, unaligned vector
            else simt::stg_v4(g + 16 * v, *(const uint4*)(ob + 16 * v));
        }
        const int done = nv << 4, rem = ohi - done;
        const int first = done > olo ? done : olo;                 // (nv == 0: nothing was flushed, olo stays)
        if (all) {
            for (int j = first + (int)lane; j < ohi; j += G) simt::stg_u8(g + j, ob[j]);
            simt::syncwarp(gmask);
            return;
        }
        constexpr int TSTEPS = (15 + G) / G;                       // the tail is < 16 bytes: 1 step (G >= 16) or 2 (G = 8)
        uint8_t t[TSTEPS];
#pragma unroll
        for (int i = 0; i < TSTEPS; i++) { const int j = (int)lane + i * G; t[i] = (nv > 0 && j < rem) ? ob[done + j] : (uint8_t)0; }
        simt::syncwarp(gmask);                                     // tail read before the front is overwritten; stores ordered
        if (nv > 0) {
#pragma unroll
            for (int i = 0; i < TSTEPS; i++) { const int j = (int)lane + i * G; if (j < rem) ob[j] = t[i]; }
            ostart += done; olo = 0; ohi = rem;
            simt::syncwarp(gmask);
        }
    };
    stage_reset(0);
    const int out_fast = cap - (FAST_L + FAST_M + 16);
    int in_fast = st.window(0, isize) - DEC_AHEAD;
    for (;;) {
        // Slide the input window when this group needs it (one chunk further ahead than strictly needed, so that it
        // happens about once per chunk).  The decision is the group's own: a vote over exactly the group's lanes.
        if (cur.ip > in_fast) in_fast = st.window(cur.ip, isize, DecGeom<G>::CHUNK) - DEC_AHEAD;
        if (cur.ip <= in_fast && cur.op <= out_fast) {
            const uint8_t* const h = rb + (((uint32_t)cur.ip + rot) & (RING - 1));
            // branch-free header parse: both possible length bytes are loaded unconditionally (they are inside the
            // 64-byte window either way) and selected, so the loads of one header overlap instead of chaining
            const uint32_t token = h[0], e1 = h[1];
            const bool lx = (token >> 4) == 15;
            const uint32_t L = lx ? 15 + e1 : token >> 4, hdr = lx ? 2u : 1u;
            if (L <= (uint32_t)FAST_L) {                            // (an e1 of 255 gives L >= 270 and falls to the careful path)
                const uint8_t* const q = h + hdr + L;
                const uint32_t off = q[0] | ((uint32_t)q[1] << 8), e2 = q[2];
                const bool mx = (token & 15) == 15;
                const uint32_t M = (mx ? 15 + e2 : token & 15) + 4, adv = hdr + L + (mx ? 3u : 2u);
                if (M <= (uint32_t)FAST_M) {                        // (255 -> M >= 274 -> careful path)
                    const uint32_t opl = (uint32_t)cur.op + L;
                    if (off - 1u >= opl) { result = -(cur.ip + (int)adv) - 1; break; }            // :863 / :983
                    const uint8_t* const lit = h + hdr;
                    auto lds = [](const uint8_t* p) { return *p; };
                    uint8_t* const o = ob + ohi;                   // == staged image of output position cur.op
                    fast_steps_smem<G, FAST_L>(o + lane, lit + lane, L, lane, lds);
                    uint8_t* const mo = o + L;
                    const uint32_t span = off < M ? off : M;       // distinct source bytes
                    if (off <= L) {
                        // source inside this sequence's own literals: still in the input ring, no ordering needed
                        const uint8_t* const s = lit + (L - off);
                        if (off >= M) fast_steps_smem<G, FAST_M>(mo + lane, s + lane, M, lane, lds);
                        else {
                            const float rcp = 1.0f / (float)off;
#pragma unroll 1
                            for (uint32_t k = lane; k < M; k += G) mo[k] = s[small_mod(k, off, rcp)];
                        }
                    } else {
                        const int spos = (int)opl - (int)off;      // output position of the first source byte
                        const int glim = ostart + olo;             // positions below are in global memory, visible
                        if (!LZ4B200_DEC_UNIFIED_MATCH && spos + (int)span <= glim) {
                            const uint8_t* const s = dst + spos;
                            auto ldg = [](const uint8_t* p) { return simt::ldg_u8(p); };
                            if (off >= M) fast_steps_smem<G, FAST_M>(mo + lane, s + lane, M, lane, ldg);
                            else {
                                const float rcp = 1.0f / (float)off;
#pragma unroll 1
                                for (uint32_t k = lane; k < M; k += G) mo[k] = simt::ldg_u8(s + small_mod(k, off, rcp));
                            }
                        } else {
                            // (partly) staged source: bytes other lanes wrote in this or earlier sequences
                            simt::syncwarp(gmask);
                            const float rcp = 1.0f / (float)off;
#pragma unroll 1
                            for (uint32_t k = lane; k < M; k += G) {
                                const int p = spos + (int)(off >= M ? k : small_mod(k, off, rcp));
                                mo[k] = p >= glim ? ob[p - ostart] : simt::ldg_u8(dst + p);
                            }
                        }
                    }
                    ohi += (int)(L + M);
                    cur.ip += (int)adv; cur.op = (int)(opl + M);
                    // Flush when this group's stage is full (group-uniform state: every lane of the group takes the same
                    // branch, and the collectives inside flush() name exactly the group's lanes).
                    if (ohi > FLUSH_AT) flush(false);
                    continue;
                }
            }
        }
        // ---------------- careful path: works on global memory directly ----------------
        flush(true);
        const int r = decode_careful<G, KNOWN>(st, cur, isize, dst, cap, &result);
        if (r != 0) { ohi = olo; break; }
        simt::syncwarp(gmask);                                     // its stores are visible before anything reads them back
        stage_reset(cur.op);
    }
    if (ohi > olo) flush(true);
    st.end();
    return result;
}

}  // namespace lz4b200
