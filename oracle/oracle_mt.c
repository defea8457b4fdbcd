/*
 * oracle_mt.c -- TEST/BENCH INFRASTRUCTURE ONLY: runs a per-block CPU codec function (the reference's own,
 * loaded from oracle/_ref, or the port in lz4_oracle.c) over a batch of independent blocks on `nthreads`
 * host threads with a static block partition, the way SURVEY.md 8(d) asks the CPU baseline to be timed
 * (one block per task, like original/bench.c:402-443 times fixed-size chunks).  Used by bench.py's
 * cpu_baseline / --impl reference legs; the product never calls it.
 */
#define _GNU_SOURCE
#include "lz4_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <time.h>

typedef struct {
    int kind; void* fn;
    const uint8_t* src; const int64_t* src_off; const int32_t* src_len;
    uint8_t* dst; const int64_t* dst_off; const int32_t* dst_cap; int32_t* out;
    int lo, hi;
} job_t;

static void* worker(void* arg)
{
    job_t* j = (job_t*)arg;
    for (int i = j->lo; i < j->hi; i++) {
        const char* s = (const char*)(j->src + j->src_off[i]);
        char* d = (char*)(j->dst + j->dst_off[i]);
        if (j->kind == 0) j->out[i] = ((lz4o_enc_fn)j->fn)(s, d, j->src_len[i], j->dst_cap[i]);
        else              j->out[i] = ((lz4o_dec_fn)j->fn)(s, d, j->dst_cap[i]);
    }
    return 0;
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

double lz4o_mt_run(int kind, void* fn, const uint8_t* src, const int64_t* src_off, const int32_t* src_len,
                   uint8_t* dst, const int64_t* dst_off, const int32_t* dst_cap, int32_t* out, int n, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n > 0 ? n : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)nthreads);
    double t0 = now_s();
    for (int t = 0; t < nthreads; t++) {
        job_t j = { kind, fn, src, src_off, src_len, dst, dst_off, dst_cap, out,
                    (int)((int64_t)n * t / nthreads), (int)((int64_t)n * (t + 1) / nthreads) };
        jobs[t] = j;
        if (t + 1 < nthreads) pthread_create(&th[t], 0, worker, &jobs[t]);
    }
    worker(&jobs[nthreads - 1]);
    for (int t = 0; t + 1 < nthreads; t++) pthread_join(th[t], 0);
    double t1 = now_s();
    free(th); free(jobs);
    return t1 - t0;
}
