// lz4_decode.cuh -- batched LZ4 block decoder: G lanes (a warp or a sub-warp group) own one block.
//
// Replaces LZ4_uncompress (original/lz4.c:812-914) and LZ4_uncompress_unknownOutputSize (:916-1044), i.e. lz4net's
// LZ4_uncompress_safe64 / _unknownOutputSize_safe64 (src/LZ4ps/LZ4Codec.Safe64.Dirty.cs:533-659,665-798).
//
// Data flow per block:
//   compressed stream  --cp.async.bulk (TMA engine, 16 B aligned chunks, mbarrier completion)-->  a small ring in
//   shared memory owned by the group  -->  tokens / lengths / offsets parsed from the ring (LDS, group-uniform);
//   literals: ring -> global, 128-bit funnel-shifted stores; matches: global (own earlier output, L1/L2 hits) ->
//   global.  The stream is consumed strictly sequentially, so the ring only needs to hide one HBM round trip.
// The accept / reject decisions are those of the reference's 64-bit flavour (see oracle/lz4_oracle.c, which is
// pinned against the reference sources); a malformed stream yields a negative result and never an access outside
// [src, src+isize) or [dst, dst+cap).
#pragma once
#include "simt.cuh"
#include "lz4_copy.cuh"

namespace lz4b200 {

constexpr int DEC_LOG_CHUNK = 10;                       // 1 KiB per bulk copy
constexpr int DEC_CHUNK = 1 << DEC_LOG_CHUNK;
constexpr int DEC_SLOTS = 2;
constexpr int DEC_RING = DEC_CHUNK * DEC_SLOTS;

struct alignas(16) DecRing {
    uint8_t buf[DEC_RING];
    simt::mbar_t bar[DEC_SLOTS];
};

// Per-group view of the compressed stream through the ring.
template <int G>
struct DecStream {
    DecRing* ring;
    const uint8_t* abase;        // src rounded down to 16 B
    uint32_t skew;               // src - abase
    uint32_t total;              // round_up16(skew + isize): bytes that may be fetched
    int nchunks;
    int issued, ready;           // chunks whose bulk copy has been launched / has completed
    uint32_t uses[DEC_SLOTS];    // bulk copies ever launched on each slot (mbarrier phase parity), kept across blocks
    int lane; uint32_t gmask;

    SIMT_MEM void begin(const uint8_t* src, int isize)
    {
        skew = (uint32_t)((uintptr_t)src & 15);
        abase = src - skew;
        total = (skew + (uint32_t)isize + 15u) & ~15u;
        nchunks = (int)((total + DEC_CHUNK - 1) >> DEC_LOG_CHUNK);
        issued = ready = 0;
        fill(0);
    }
    // launch every chunk that may be resident while the read cursor is in chunk c
    SIMT_MEM void fill(int c)
    {
        int want = c + DEC_SLOTS; if (want > nchunks) want = nchunks;
        if (issued >= want) return;
        simt::syncwarp(gmask);                                     // every lane is done with the slots being recycled
        if (lane == 0) {
            for (int k = issued; k < want; k++) {
                uint32_t b0 = (uint32_t)k << DEC_LOG_CHUNK;
                uint32_t nb = total - b0; if (nb > DEC_CHUNK) nb = DEC_CHUNK;
                simt::bulk_g2s(ring->buf + (k % DEC_SLOTS) * DEC_CHUNK, abase + b0, nb, &ring->bar[k % DEC_SLOTS]);
            }
        }
        issued = want;
    }
    SIMT_MEM void wait_chunk(int c)                               // make chunks <= c readable
    {
        while (ready <= c) {
            int s = ready % DEC_SLOTS;
            simt::mbar_wait(&ring->bar[s], (uses[s] + (uint32_t)(ready / DEC_SLOTS)) & 1);
            ready++;
        }
    }
    SIMT_MEM int chunk_of(uint32_t p) const { return (int)((p + skew) >> DEC_LOG_CHUNK); }
    // one byte at stream position p (the read cursor): group-uniform broadcast LDS
    SIMT_MEM uint32_t byte_at(uint32_t p)
    {
        int c = chunk_of(p);
        if (c + DEC_SLOTS > issued) fill(c);
        if (c >= ready) wait_chunk(c);
        return ring->buf[(p + skew) & (DEC_RING - 1)];
    }
    // finish the block: every launched copy must have landed before the slots / barriers are reused
    SIMT_MEM void end()
    {
        if (issued > 0) wait_chunk(issued - 1);
        for (int s = 0; s < DEC_SLOTS; s++) uses[s] += (uint32_t)((issued - s + DEC_SLOTS - 1) / DEC_SLOTS);
        simt::syncwarp(gmask);
    }
    // copy n stream bytes starting at p to dst; advances through the ring piecewise
    SIMT_MEM void copy_out(uint8_t* dst, uint32_t p, uint32_t n)
    {
        while (n) {
            int c = chunk_of(p);
            if (c + DEC_SLOTS > issued) fill(c);
            int lim = c + DEC_SLOTS; if (lim > issued) lim = issued;
            uint32_t avail = ((uint32_t)lim << DEC_LOG_CHUNK) - skew - p;   // resident bytes from p on
            uint32_t m = n < avail ? n : avail;
            wait_chunk(chunk_of(p + m - 1));
            RingSrc<DEC_RING> s{ring->buf, p + skew};
            group_copy<G, false>(dst, s, m, lane, gmask);
            dst += m; p += m; n -= m;
        }
    }
};

// Decode one block.  KNOWN: cap is the exact decoded size, result = bytes read (LZ4_uncompress).
// !KNOWN: cap is the capacity, result = bytes written (LZ4_uncompress_unknownOutputSize).  < 0 = malformed.
template <int G, bool KNOWN>
SIMT_DEV int decode_block(DecStream<G>& st, const uint8_t* src, int isize, uint8_t* dst, int cap)
{
    const int lane = st.lane; const uint32_t gmask = st.gmask;
    if (isize <= 0 || cap < 0) return -1;                          // original/lz4.c:949; a block has >= 1 token
    st.begin(src, isize);
    constexpr int LEN_LIMIT = 0x3FFFFFFF;                          // a run this long cannot fit any int-sized buffer
    int ip = 0, op = 0, result;
    for (;;) {
        if (ip >= isize) { result = -ip - 1; break; }
        const uint32_t token = st.byte_at(ip++);
        int L = (int)(token >> 4);
        if (L == 15) {                                             // :843 / :959-963
            uint32_t s = 255;
            if (KNOWN) { do { if (ip >= isize || L > LEN_LIMIT) { L = -1; break; } s = st.byte_at(ip++); L += (int)s; } while (s == 255); }
            else       { while (ip < isize && s == 255 && L <= LEN_LIMIT) { s = st.byte_at(ip++); L += (int)s; } }
            if (L < 0 || L > LEN_LIMIT) { result = -ip - 1; break; }
        }
        int end = op + L;
        bool last;
        if (KNOWN) last = end > cap - 8;                           // :847
        else       last = end > cap - 12 || ip + L > isize - 8;    // :968
        if (last) {
            bool ok = KNOWN ? (end == cap && ip + L <= isize)      // :849-857
                            : (end <= cap && ip + L == isize);     // :974-975
            if (!ok) { result = -ip - 1; break; }
            st.copy_out(dst + op, ip, L);
            ip += L; op = end;
            result = KNOWN ? ip : op;
            break;
        }
        if (ip + L + 2 > isize) { result = -ip - 1; break; }       // (KNOWN: bounded reads; !KNOWN: implied by :968)
        st.copy_out(dst + op, ip, L);
        ip += L; op = end;
        uint32_t off = st.byte_at(ip);                             // :862 / :982 (two sequenced reads: the ring may turn over)
        off |= st.byte_at(ip + 1) << 8; ip += 2;
        if (off == 0 || off > (uint32_t)op) { result = -ip - 1; break; }            // :863 / :983 (offset 0 rejected by design)
        int M = (int)(token & 15);
        if (M == 15) {                                             // :866 / :986-999
            uint32_t s = 255;
            if (KNOWN) { do { if (ip >= isize || M > LEN_LIMIT) { M = -1; break; } s = st.byte_at(ip++); M += (int)s; } while (s == 255); }
            else       { while (ip < isize - 6 && M <= LEN_LIMIT) { s = st.byte_at(ip++); M += (int)s; if (s != 255) break; } }
            if (M < 0 || M > LEN_LIMIT) { result = -ip - 1; break; }
        }
        end = op + M + 4;
        if (end > cap - 5) { result = -ip - 1; break; }            // :893 / :1025 -- the last 5 bytes are literals
        simt::syncwarp(gmask);                                     // literal (and earlier match) stores -> match loads
        group_copy_match<G>(dst + op, off, (uint32_t)(M + 4), lane, gmask);
        op = end;
    }
    st.end();
    return result;
}

}  // namespace lz4b200
