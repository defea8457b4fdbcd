/*
 * lz4_oracle.h -- TEST INFRASTRUCTURE ONLY (see lz4_oracle.c).
 *
 * CPU restatement ("port" oracle) of the LZ4 r93 block codec that lz4net ships.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; nothing under lz4net_b200/ (the product) links it.
 */
#ifndef LZ4_ORACLE_H
#define LZ4_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* original/lz4.h:85 (LZ4_compressBound) == src/LZ4/LZ4Codec.cs:313-316 (MaximumOutputLength) */
int lz4o_bound(int n);

/* original/lz4.c:774-792 LZ4_compress_limitedOutput (dispatch on n < 65547) ; returns bytes written, 0 = failed */
int lz4o_encode(const uint8_t* src, int n, uint8_t* dst, int cap);

/* original/lz4hc.c:745-755 LZ4_compressHC_limitedOutput ; returns bytes written, 0 = failed */
int lz4o_encode_hc(const uint8_t* src, int n, uint8_t* dst, int cap);

/* original/lz4.c:812-914 LZ4_uncompress ; returns bytes READ, <0 on malformed input.
 * isize bounds the reads (the managed callers always know it: src/LZ4cc/LZ4Codec.64.cpp:81-100);
 * pass a negative isize for the reference's unbounded-read behaviour. */
int lz4o_decode_known(const uint8_t* src, int isize, uint8_t* dst, int osize);

/* original/lz4.c:916-1044 LZ4_uncompress_unknownOutputSize ; returns bytes WRITTEN, <0 on malformed input */
int lz4o_decode_unknown(const uint8_t* src, int isize, uint8_t* dst, int max_out);

/* ---- multi-threaded batch driver used only for the CPU baseline timing (oracle_mt.c) ---- */
typedef int (*lz4o_enc_fn)(const char*, char*, int, int);            /* LZ4_compress[HC]_limitedOutput shape */
typedef int (*lz4o_dec_fn)(const char*, char*, int);                 /* LZ4_uncompress shape */
/* kind: 0 = encode-shaped fn, 1 = LZ4_uncompress-shaped fn. Blocks i in [0,n): src+src_off[i] (src_len[i] bytes)
 * -> dst+dst_off[i] (dst_cap[i] bytes); out[i] = return value. Returns wall seconds of the parallel region. */
double lz4o_mt_run(int kind, void* fn, const uint8_t* src, const int64_t* src_off, const int32_t* src_len,
                   uint8_t* dst, const int64_t* dst_off, const int32_t* dst_cap, int32_t* out, int n, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
