// lz4_encode.cuh -- byte-exact LZ4 r93 fast encoder, one warp per block.
//
// Replaces LZ4_compress64kCtx (original/lz4.c:573-771; lz4net: LZ4_compress64kCtx_safe64,
// src/LZ4ps/LZ4Codec.Safe64.Dirty.cs:306-527) for inputs < 65547 bytes and LZ4_compressCtx (original/lz4.c:345-562;
// Safe64.Dirty.cs:77-300) above that, behind the LZ4_compress_limitedOutput dispatch (original/lz4.c:774-792).
// The emitted bytes must equal lz4net's LZ4Codec.Encode, so the greedy parse is reproduced exactly -- but as a
// warp-wide formulation of the serial state machine:
//
//   * one "round" evaluates 32 consecutive probe positions of the reference's find-match loop at once: the probe
//     positions follow from the attempt counter alone (step = attempts >> 6, :636/:644), their hashes only depend on
//     the input, and the table state each serial iteration would have seen is reconstructed inside the round with
//     MATCH.ANY (a lower lane with the same hash supplies the candidate instead of the table);
//   * VOTE + FFS picks the first lane that either hits or runs past mflimit (:648); only lanes up to it commit their
//     table update, highest lane per bucket winning -- exactly the serial write order;
//   * the backward catch-up (:657), the match-length count (:701-716) and all copies are lane-parallel.
//
// The 16 KiB position table (u16[8192] for the 64 K variant, u32[4096] for the general one -- the same footprint)
// lives in shared memory, zero-filled per block (zero == "candidate at position 0", :583/:651).
#pragma once
#include "simt.cuh"
#include "lz4_copy.cuh"

namespace lz4b200 {

struct alignas(16) EncShared { uint32_t table[4096]; };

constexpr int LZ4_64KLIMIT = 65547;                                // original/lz4.c:565

// token length extension: `v` as a run of 255s plus a remainder byte, written lane-parallel; returns bytes written
SIMT_DEV int put_len_ext(uint8_t* dst, int op, int cap, int v, int lane)
{
    if (v < 255) {                                                 // the common case: one byte
        if (lane == 0 && op < cap) simt::stg_u8(dst + op, (uint8_t)v);
        return 1;
    }
    const int nff = v / 255;
    for (int i = lane; i < nff; i += 32) if (op + i < cap) simt::stg_u8(dst + op + i, 255);
    if (lane == 0 && op + nff < cap) simt::stg_u8(dst + op + nff, (uint8_t)(v - nff * 255));
    return nff + 1;
}

template <bool GENERAL>
SIMT_DEV int encode_block_t(EncShared* sh, const uint8_t* src, int n, uint8_t* dst, int cap, int lane)
{
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr int HSHIFT = GENERAL ? 20 : 19;                      // :185-187 / :566-569
    uint16_t* const T16 = (uint16_t*)sh->table;
    uint32_t* const T32 = sh->table;
    const int mflimit = n - 12, matchlimit = n - 5;                // :361,:366 / :590,:596
    int ip = 0, anchor = 0, op = 0, ref = 0, tok = 0;
    uint32_t tokval = 0;
    InWords in; in.init(src);

    for (int i = lane; i < 1024; i += 32) ((uint4*)sh->table)[i] = uint4{0, 0, 0, 0};
    simt::syncwarp(FULL);

    if (n >= 13) {                                                 // :387 / :615 (MINLENGTH)
        ip = 1;                                                    // :404 / :631
        for (;;) {
            // ---------------- find a match: rounds of 32 serial probe iterations (:415-429 / :642-654) ----------
            uint32_t A = (1u << 6) + 3;                            // findMatchAttempts, :409 / :636
            int p0 = ip;
            bool finished = false;
            for (;;) {
                int pos, nxt;
                if (A == (1u << 6) + 3) { pos = p0 + lane; nxt = pos + 1; }      // first round: 32 attempts, all with step 1
                else {
                    const int q = (int)(A >> 6), cross = 64 - (int)(A & 63);
                    const int bump = lane - cross;                 // attempts before `lane` that already use step q+1
                    pos = p0 + q * lane + (bump > 0 ? bump : 0);
                    nxt = pos + q + (lane >= cross ? 1 : 0);
                }
                const bool valid = nxt <= mflimit;                 // the bounds test precedes the probe (:420 / :648)
                uint32_t v = 0, h = 0x80000000u | (uint32_t)lane;  // invalid lanes get a key nobody shares
                if (valid) { v = in.at(pos); h = (v * 2654435761u) >> HSHIFT; }
                const uint32_t same = simt::match_any(FULL, h);
                const uint32_t lower = same & ((1u << lane) - 1u);
                const int from = lower ? 31 - simt::clz(lower) : lane;
                const int fwd = (int)simt::shfl(FULL, (uint32_t)pos, from);
                int cand = 0; bool hit = false;
                if (valid) {
                    cand = lower ? fwd : (GENERAL ? (int)T32[h] : (int)T16[h]);
                    hit = (!GENERAL || cand >= pos - 65535) && in.at(cand) == v;        // :429 / :654
                }
                const uint32_t stop = simt::ballot(FULL, !valid || hit);
                const int f = stop ? simt::ffs(stop) - 1 : 32;
                // commit the table writes of the serial iterations that really happened, last writer per bucket
                const bool commit = valid && lane <= f;
                const uint32_t cmask = simt::ballot(FULL, commit);
                const uint32_t above = same & cmask & ~((2u << lane) - 1u);
                if (commit && !above) { if (GENERAL) T32[h] = (uint32_t)pos; else T16[h] = (uint16_t)pos; }
                simt::syncwarp(FULL);
                if (f < 32) {
                    if (!((cmask >> f) & 1u)) { finished = true; break; }       // lane f ran past mflimit -> last literals
                    ip = (int)simt::shfl(FULL, (uint32_t)pos, f);
                    ref = (int)simt::shfl(FULL, (uint32_t)cand, f);
                    break;
                }
                p0 = (int)simt::shfl(FULL, (uint32_t)nxt, 31);
                A += 32;
            }
            if (finished) break;

            // ---------------- catch up (:432 / :657) ----------------
            for (;;) {
                const int k = lane + 1;
                const bool eq = ip - k >= anchor && ref - k >= 0 &&
                                simt::ldg_nc_u8(src + ip - k) == simt::ldg_nc_u8(src + ref - k);
                const uint32_t ne = ~simt::ballot(FULL, eq);
                const int cnt = ne ? simt::ffs(ne) - 1 : 32;
                ip -= cnt; ref -= cnt;
                if (cnt < 32) break;
            }

            // ---------------- literal run (:435-466 / :660-691) ----------------
            {
                const int L = ip - anchor;
                tok = op++;
                if (op + L + (L >> 8) > cap - 8) return 0;         // :438 / :663
                if (L >= 15) { tokval = 0xF0; op += put_len_ext(dst, op, cap, L - 15, lane); }
                else tokval = (uint32_t)L << 4;
                InputSrc s{src + anchor};
                group_copy<32, false>(dst + op, s, (uint32_t)L, lane, FULL);
                op += L;
            }

            // ---------------- match(es) ----------------
            bool again;
            do {
                if (lane == 0 && (!GENERAL || op + 2 <= cap)) {    // :470 / :695
                    simt::stg_u8(dst + op, (uint8_t)(ip - ref)); simt::stg_u8(dst + op + 1, (uint8_t)((ip - ref) >> 8));
                }
                op += 2;
                ip += 4; ref += 4; anchor = ip;
                for (;;) {                                         // count equal bytes up to matchlimit (:475-494 / :701-716)
                    const int a = ip + 4 * lane;
                    int room = matchlimit - a; room = room < 0 ? 0 : (room > 4 ? 4 : room);
                    int cnt = 0;
                    if (room > 0) {
                        const uint32_t x = in.at(a) ^ in.at(ref + 4 * lane);
                        cnt = x ? (simt::ffs(x) - 1) >> 3 : 4;
                        if (cnt > room) cnt = room;
                    }
                    const uint32_t part = simt::ballot(FULL, cnt != 4);
                    if (part) {
                        const int fl = simt::ffs(part) - 1;
                        ip += 4 * fl + (int)simt::shfl(FULL, (uint32_t)cnt, fl);
                        break;
                    }
                    ip += 128; ref += 128;
                }
                const int M = ip - anchor;
                if (op + (M >> 8) > cap - 6) return 0;             // :501 / :728
                if (M >= 15) { tokval |= 15; op += put_len_ext(dst, op, cap, M - 15, lane); }
                else tokval |= (uint32_t)M;
                if (lane == 0) simt::stg_u8(dst + tok, (uint8_t)tokval);

                again = false;
                if (ip > mflimit) { anchor = ip; finished = true; break; }      // :516 / :736
                // table fix-up for ip-2, then probe ip itself (:519-531 / :739-751); reads first, then lane 0 writes
                const int p2 = ip - 2;
                const uint32_t h2 = (in.at(p2) * 2654435761u) >> HSHIFT;
                const uint32_t vi = in.at(ip);
                const uint32_t h = (vi * 2654435761u) >> HSHIFT;
                ref = (h == h2) ? p2 : (GENERAL ? (int)T32[h] : (int)T16[h]);
                simt::syncwarp(FULL);
                if (lane == 0) {
                    if (GENERAL) { T32[h2] = (uint32_t)p2; T32[h] = (uint32_t)ip; }
                    else         { T16[h2] = (uint16_t)p2; T16[h] = (uint16_t)ip; }
                }
                simt::syncwarp(FULL);
                if ((!GENERAL || ref > ip - 65536) && in.at(ref) == vi) {
                    tok = op++; tokval = 0; again = true;          // zero-literal sequence (:531 / :751)
                }
            } while (again);
            if (finished) break;
            anchor = ip++;                                         // :534 / :754
        }
    }

    // ---------------- last literals (:540-551 / :760-767) ----------------
    {
        const int R = n - anchor;
        if (op + R + 1 + (R - 15 + 255) / 255 > cap) return 0;
        tok = op++;
        if (R >= 15) { if (lane == 0) simt::stg_u8(dst + tok, 0xF0); op += put_len_ext(dst, op, cap, R - 15, lane); }
        else if (lane == 0) simt::stg_u8(dst + tok, (uint8_t)(R << 4));
        InputSrc s{src + anchor};
        group_copy<32, false>(dst + op, s, (uint32_t)R, lane, FULL);
        op += R;
    }
    return op;
}

// LZ4_compress_limitedOutput dispatch (original/lz4.c:774-792)
SIMT_DEV int encode_block(EncShared* sh, const uint8_t* src, int n, uint8_t* dst, int cap, int lane)
{
    if (n < 0 || cap < 0) return 0;
    simt::syncwarp(0xFFFFFFFFu);                                   // previous block's table users are done
    return n < LZ4_64KLIMIT ? encode_block_t<false>(sh, src, n, dst, cap, lane)
                            : encode_block_t<true>(sh, src, n, dst, cap, lane);
}

}  // namespace lz4b200
