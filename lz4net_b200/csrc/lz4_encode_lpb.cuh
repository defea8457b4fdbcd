// lz4_encode_lpb.cuh -- byte-exact LZ4 r93 fast encoder, ONE LANE per block ("lane-per-block"), K probes per step.
//
// The same contract as lz4_encode.cuh: LZ4_compress64kCtx (original/lz4.c:573-771; lz4net: LZ4_compress64kCtx_safe64,
// src/LZ4ps/LZ4Codec.Safe64.Dirty.cs:306-527) for inputs < 65547 bytes and LZ4_compressCtx (original/lz4.c:345-562)
// above that, bytes identical to lz4net's LZ4Codec.Encode.
//
// Why a second mapping: the warp-per-block encoder is a per-warp latency chain and an SM holds 14 of its 16 KiB
// position tables -- 14 blocks in flight, issue slots half idle.  This kernel's warp adds 16 more blocks in flight per SM:
// every one of its first 16 LANES runs the greedy parse of a block of its own, with its position table in global memory
// (an arena of 32 KiB per lane, kept L2-resident with an evict-last policy: a table access is an L2 round trip).
// A table entry holds the position AND a 16-bit tag of the 4 input bytes at that position (the bits of the hash product
// below the bucket index): a candidate whose tag differs cannot match, so its 4 bytes are never fetched -- in
// incompressible stretches nearly every probe ends at the table, where the first version of this kernel fetched a random
// earlier word of the block per probe (measured: 10x DRAM read amplification, every step waiting on the slowest of 8 DRAM
// round trips, L2 hit rate of the whole encoder from 84 % to 59 %).  The tag never changes a decision: equal 4 bytes imply
// equal tags, and the empty bucket's implicit candidate, position 0 (:583 / :651), is entered explicitly with its tag.
// To keep the
// lane's chain short, one step evaluates K consecutive iterations of the reference's find-match loop at once
// (:415-429 / :642-654): the probe positions follow from the attempt counter alone, so the K input words, hashes and
// table entries are loaded together (one memory round trip instead of K), iterations that share a bucket inside the step
// are resolved in registers (the later one's candidate is the earlier one's position), and the table is written in
// iteration order up to the first hit -- exactly the state the serial loop leaves.  The two table operations that
// follow every match (insert ip-2, probe ip, :519-531 / :739-751) ride at the front of the next step.
//
// All lanes of the warp run the same step loop (a step, then -- for the lanes that found a match -- catch-up, match
// length and emission), so the warp stays converged on the step even though every lane is in a different block.
// Literal runs longer than 64 bytes are copied by the whole warp (lz4_copy.cuh).
#pragma once
#include "simt.cuh"
#include "lz4_copy.cuh"

namespace lz4b200 {

struct EncLpbBatch {
    const uint8_t* src; const int64_t* src_off; const int32_t* src_len;
    uint8_t* dst; const int64_t* dst_off; const int32_t* dst_cap;
    int32_t* out_len; int32_t n_blocks;
};

constexpr int ENC_LPB_K = 8;                                        // iterations of the find-match loop per step
constexpr int ENC_LPB_LANES = 16;                                   // lanes of the warp that own a block (the others only help with long copies)
constexpr int ENC_LPB_TABLE = 32768;                                // bytes of table per lane: 8192 x (u16 position, u16 tag) or 4096 x (u32, u32)
constexpr int ENC_LPB_COOP = 64;                                    // literal runs longer than this are copied by the warp
constexpr int ENC_LPB_64KLIMIT = 65547;                             // original/lz4.c:565

// little-endian 32-bit read at an arbitrary byte position of the input (two aligned words, funnel shift)
SIMT_DEV uint32_t lpb_in32(const uint8_t* src, int p) { return in32(src, p); }

// Block hand-out shared with the warp-per-block encoders: ONE 64-bit word = (blocks taken from the front) << 32 | (blocks
// taken from the back).  Warps take from the front, lanes from the back; a lane holds a block ~30x longer than a warp, so
// lanes stop taking when fewer than `reserve` blocks are left -- the warps finish those while the lanes' last blocks run.
SIMT_DEV int64_t take_front(unsigned long long* q, uint32_t n)
{
    for (;;) {
        const unsigned long long old = simt::atomic_load_u64(q);
        const uint32_t f = (uint32_t)(old >> 32), b = (uint32_t)old;
        if (f + b >= n) return -1;
        if (simt::atomic_cas_u64(q, old, old + (1ull << 32)) == old) return (int64_t)f;
    }
}
SIMT_DEV int64_t take_back(unsigned long long* q, uint32_t n, uint32_t reserve)
{
    for (;;) {
        const unsigned long long old = simt::atomic_load_u64(q);
        const uint32_t f = (uint32_t)(old >> 32), b = (uint32_t)old;
        if (f + b + reserve >= n) return -1;
        if (simt::atomic_cas_u64(q, old, old + 1ull) == old) return (int64_t)(n - 1u - b);
    }
}

// `tables`: ENC_LPB_LANES x 32 KiB of global memory owned by this warp.
SIMT_DEV void lpb_encode_warp(uint8_t* tables, const EncLpbBatch& a, unsigned long long* queue, uint32_t reserve, int lane)
{
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr int K = ENC_LPB_K;
    uint8_t* const T = tables + (size_t)(lane < ENC_LPB_LANES ? lane : 0) * ENC_LPB_TABLE;
    const uint64_t keep = simt::l2_policy_keep();

    bool active = false, drained = lane >= ENC_LPB_LANES;
    uint32_t blk = 0;
    const uint8_t* src = nullptr; uint8_t* dst = nullptr;
    int n = 0, cap = 0, mflimit = 0, matchlimit = 0;
    bool general = false;
    int anchor = 0, op = 0;
    int pos = 0;                        // find-match loop: position of the next iteration; after a match: the match end (ip)
    uint32_t att = 0;                   // its attempt counter (findMatchAttempts, :409 / :636)
    bool special = false;               // the step starts with the post-match table operations on `pos`
    int state = 0;                      // 0 stepping, 1 last literals pending, 2 failed (output too small), 3 done
    int coop_len = 0; int coop_src = 0, coop_dst = 0;      // a literal run the warp copies: src + coop_src -> dst + coop_dst
    int after_coop = 0;                 // what follows the warp's copy: 1 = the match part of the sequence, 2 = the block is finished
    int m_ip = 0, m_ref = 0, m_tok = 0, m_L = 0;           // the sequence whose literals the warp is copying

    // table entry: position + tag.  64 K variant: one u32 = position | tag << 16; general variant: two u32 (position, tag).
    auto tag_of = [&](uint32_t v) -> uint32_t { return ((v * 2654435761u) >> 3) & 0xFFFFu; };
    auto tget = [&](uint32_t h, uint32_t* tag) -> int {
        if (general) { const uint2 e = simt::ldt_hint_v2(T + h * 8u, keep); *tag = e.y; return (int)e.x; }
        const uint32_t e = simt::ldt_hint_u32(T + h * 4u, keep); *tag = e >> 16; return (int)(e & 0xFFFFu);
    };
    auto tput = [&](uint32_t h, int p, uint32_t tag) {
        if (general) simt::stt_hint_v2(T + h * 8u, uint2{(uint32_t)p, tag}, keep); else simt::stt_hint_u32(T + h * 4u, (uint32_t)p | (tag << 16), keep);
    };
    auto put = [&](int at, uint32_t v) { if (at < cap) simt::stg_u8(dst + at, (uint8_t)v); };       // never outside [dst, dst + cap)
    auto put_len = [&](int at, int v) -> int { while (v > 254) { put(at++, 255); v -= 255; } put(at++, (uint32_t)v); return at; };

    // offset + match length + token of a sequence whose literals are already written; returns false if the output is full.
    // On return `pos` is the match end; the lane goes on with the post-match step or with the last literals.
    auto finish_sequence = [&](int ip, int ref, int tok, int L) -> bool {
        put(op, (uint32_t)(ip - ref) & 255u); put(op + 1, ((uint32_t)(ip - ref) >> 8) & 255u); op += 2;     // :470 / :695
        int mp = ip + 4, mr = ref + 4;
        bool counted = false;
        while (!counted && mp + 4 <= matchlimit) {                      // == the 8/4/2/1 scheme of :475-494 / :701-716
            const uint32_t x = lpb_in32(src, mp) ^ lpb_in32(src, mr);
            if (x) { mp += (simt::ffs(x) - 1) >> 3; counted = true; }
            else { mp += 4; mr += 4; }
        }
        if (!counted) while (mp < matchlimit && simt::ldg_nc_u8(src + mp) == simt::ldg_nc_u8(src + mr)) { mp++; mr++; }
        const int M = mp - (ip + 4);
        if (op + (M >> 8) > cap - 6) return false;                      // :501 / :728
        uint32_t tv = (uint32_t)(L < 15 ? L : 15) << 4;
        if (M >= 15) { tv |= 15u; op = put_len(op, M - 15); } else tv |= (uint32_t)M;
        if (op > cap) return false;                                     // (only reachable with run lengths beyond any 64 KiB block)
        put(tok, tv);
        anchor = mp; pos = mp;
        if (mp > mflimit) state = 1;                                    // :516 / :736
        else special = true;
        return true;
    };

    for (;;) {
        // ---------------- idle lanes take the next block ----------------
        if (!active && !drained) {
            const int64_t got = take_back(queue, (uint32_t)a.n_blocks, reserve);
            if (got < 0) drained = true;
            else {
                blk = (uint32_t)got;
                src = a.src + a.src_off[blk]; dst = a.dst + a.dst_off[blk];
                n = a.src_len[blk]; cap = a.dst_cap[blk];
                if (n < 0 || cap < 0) simt::stg_u32(a.out_len + blk, 0u);
                else {
                    general = n >= ENC_LPB_64KLIMIT;
                    mflimit = n - 12; matchlimit = n - 5;               // :361,:366 / :590,:596
                    // zero == "candidate at position 0" (:583 / :651) with tag 0; the bucket of the block's first 4 bytes
                    // gets position 0's real tag, so that a later position with the same 4 bytes still finds it
                    for (int i = 0; i < ENC_LPB_TABLE / 16; i++) simt::stg_v4(T + 16 * i, uint4{0, 0, 0, 0});
                    anchor = 0; op = 0; pos = 1; att = 67; special = false; coop_len = 0;       // :404,:409 / :631,:636
                    state = n < 13 ? 1 : 0;                             // :387 / :615 (MINLENGTH)
                    if (state == 0) { const uint32_t v0 = lpb_in32(src, 0); tput((v0 * 2654435761u) >> (general ? 20 : 19), 0, tag_of(v0)); }
                    active = true;
                }
            }
        }
        if (!simt::ballot(FULL, active)) break;

        // ---------------- one step: K iterations of the find-match loop (preceded by the post-match operations) ----------
        if (active && state == 0) {
            const int hshift = general ? 20 : 19;                       // :185-187 / :566-569
            int P[K]; bool valid[K]; uint32_t v[K], h[K], tg[K]; int t[K];
            int hs_pos = 0; uint32_t hs = 0, vs = 0;
            {
                int p = pos; uint32_t at = att;
#pragma unroll
                for (int e = 0; e < K; e++) {
                    if (special && e == 0) { P[e] = pos; valid[e] = true; p = pos + 1; at = 67; }   // probe ip (:523 / :743); then a fresh find-match loop from ip + 1
                    else { const int step = (int)(at >> 6); P[e] = p; valid[e] = p + step <= mflimit; p += step; at++; }   // the bounds test precedes the probe (:420 / :648)
                }
            }
#pragma unroll
            for (int e = 0; e < K; e++) v[e] = valid[e] ? lpb_in32(src, P[e]) : 0u;
            if (special) { hs_pos = pos - 2; vs = lpb_in32(src, hs_pos); hs = (vs * 2654435761u) >> hshift; }
#pragma unroll
            for (int e = 0; e < K; e++) { h[e] = (v[e] * 2654435761u) >> hshift; tg[e] = 0; t[e] = valid[e] ? tget(h[e], &tg[e]) : 0; }
            // candidates: the table entry, unless an earlier operation of this step wrote the bucket
            int cand[K]; uint32_t w[K]; bool dup[K];
#pragma unroll
            for (int e = 0; e < K; e++) {
                cand[e] = t[e]; dup[e] = false; w[e] = 0;
                if (special && hs == h[e]) { cand[e] = hs_pos; w[e] = vs; dup[e] = true; }
#pragma unroll
                for (int j = 0; j < e; j++) if (valid[j] && h[j] == h[e]) { cand[e] = P[j]; w[e] = v[j]; dup[e] = true; }
            }
#pragma unroll
            for (int e = 0; e < K; e++) {
                if (!valid[e] || dup[e]) continue;
                if (tg[e] == tag_of(v[e])) w[e] = lpb_in32(src, cand[e]);      // (a different tag: different bytes, nothing to fetch)
                else w[e] = ~v[e];
            }
            int f = K;                                                   // the first iteration that hits or runs past mflimit
#pragma unroll
            for (int e = K - 1; e >= 0; e--) {
                const bool hit = valid[e] && (!general || cand[e] >= P[e] - 65535) && w[e] == v[e];   // :429 / :654, :531 / :751
                if (hit || !valid[e]) f = e;
            }
            if (special) tput(hs, hs_pos, tag_of(vs));                   // :519 / :739
#pragma unroll
            for (int e = 0; e < K; e++) if (e <= f && valid[e]) tput(h[e], P[e], tag_of(v[e]));      // in iteration order: the last writer of a bucket wins
            if (f == K) {                                                // no hit: the loop goes on
                int p = pos; uint32_t at = att;
                if (special) { p = pos + 1; at = 67; for (int e = 1; e < K; e++) { p += (int)(at >> 6); at++; } }
                else for (int e = 0; e < K; e++) { p += (int)(at >> 6); at++; }
                pos = p; att = at; special = false;
            } else {
                int fp = 0, fr = 0; bool fv = false;
#pragma unroll
                for (int e = 0; e < K; e++) if (e == f) { fp = P[e]; fr = cand[e]; fv = valid[e]; }
                if (!fv) state = 1;                                      // ran past mflimit -> last literals (:420 / :648)
                else {
                    int ip = fp, ref = fr;
                    const bool zero_lit = special && f == 0;             // `goto _next_match` with no literals and no :663 check (:531 / :751)
                    special = false;
                    if (!zero_lit) while (ip > anchor && ref > 0 && simt::ldg_nc_u8(src + ip - 1) == simt::ldg_nc_u8(src + ref - 1)) { ip--; ref--; }   // :432 / :657
                    const int L = ip - anchor, tok = op++;
                    if (!zero_lit && op + L + (L >> 8) > cap - 8) state = 2;                      // :438 / :663
                    else {
                        if (L >= 15) op = put_len(op, L - 15);
                        if (L > ENC_LPB_COOP) {                          // the warp copies the run; the match part follows it
                            coop_len = L; coop_src = anchor; coop_dst = op; op += L; after_coop = 1;
                            m_ip = ip; m_ref = ref; m_tok = tok; m_L = L;
                        } else {
                            for (int i = 0; i < L; i++) put(op + i, simt::ldg_nc_u8(src + anchor + i));   // :466 / :691
                            op += L;
                            if (!finish_sequence(ip, ref, tok, L)) state = 2;
                        }
                    }
                }
            }
        }
        // ---------------- last literals (:540-551 / :760-767) ----------------
        if (active && state == 1 && coop_len == 0) {
            const int R = n - anchor;
            if (op + R + 1 + (R - 15 + 255) / 255 > cap) state = 2;
            else {
                if (R >= 15) { put(op++, 0xF0); op = put_len(op, R - 15); } else put(op++, (uint32_t)R << 4);
                if (R > ENC_LPB_COOP) { coop_len = R; coop_src = anchor; coop_dst = op; op += R; after_coop = 2; }
                else { for (int i = 0; i < R; i++) put(op + i, simt::ldg_nc_u8(src + anchor + i)); op += R; state = 3; }
            }
        }
        // ---------------- literal runs the whole warp copies ----------------
        uint32_t req = simt::ballot(FULL, active && coop_len != 0);
        while (req) {
            const int k = simt::ffs(req) - 1; req &= req - 1;
            const uint32_t len = simt::shfl(FULL, (uint32_t)coop_len, k);
            const uint64_t dpk = (uint64_t)(uintptr_t)(dst + coop_dst), spk = (uint64_t)(uintptr_t)(src + coop_src);
            uint8_t* const d = (uint8_t*)(uintptr_t)(((uint64_t)simt::shfl(FULL, (uint32_t)(dpk >> 32), k) << 32) | simt::shfl(FULL, (uint32_t)dpk, k));
            const uint8_t* const s = (const uint8_t*)(uintptr_t)(((uint64_t)simt::shfl(FULL, (uint32_t)(spk >> 32), k) << 32) | simt::shfl(FULL, (uint32_t)spk, k));
            InputSrc sp{s};
            group_copy<32, false>(d, sp, len, lane, FULL);
        }
        if (active && coop_len != 0) {
            coop_len = 0;
            if (after_coop == 2) state = 3;
            else if (!finish_sequence(m_ip, m_ref, m_tok, m_L)) state = 2;
        }
        if (active && state >= 2) { simt::stg_u32(a.out_len + blk, state == 3 ? (uint32_t)op : 0u); active = false; }
    }
}

}  // namespace lz4b200
