/*
 * zso_bench.c -- native multi-threaded TIMING helpers for bench.py's CPU legs (cpu_baseline, --impl reference).
 *
 * TEST / MEASUREMENT INFRASTRUCTURE like the rest of oracle/ (see zso.h): the product never links or calls it.
 * Threads are created here (no Python in the timed region), pinned to one core each, start together on a barrier and the
 * wall time of the slowest is returned.  Deflate = the oracle's restatement of the reference's managed algorithm
 * (zso_deflate_segments: one stream per contiguous range of segments, which is what a reference user would run per core);
 * inflate = system zlib (the reference's own inflater cannot serve as a baseline) or the oracle's.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>

#include "zso.h"

typedef struct {
    int kind;   /* 0 deflate, 1 inflate (system zlib), 2 inflate (oracle) */
    int core, reps;
    pthread_barrier_t* bar;
    /* deflate */
    const uint8_t* in;
    size_t n, seg_size;
    int level, strategy, hv;
    uint64_t out_bytes;
    /* inflate */
    const uint8_t* comp;
    const uint64_t* seg_ends;
    size_t seg_begin, seg_end, first_offset;
    uint8_t* scratch;
    uint64_t decoded;
    int rc;
    double t0, t1;
} job_t;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void* worker(void* arg)
{
    job_t* j = (job_t*)arg;
    if (j->core >= 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(j->core, &set);
        pthread_setaffinity_np(pthread_self(), sizeof set, &set);   /* best effort */
    }
    pthread_barrier_wait(j->bar);
    j->t0 = now_s();
    for (int r = 0; r < j->reps; r++) {
        if (j->kind == 0) {
            size_t out_len = 0;
            uint32_t ad = 0;
            /* out == NULL: measure only (the sink counts bytes), same work as writing to a MemoryStream minus the copy */
            int rc = zso_deflate_segments(j->in, j->n, j->seg_size, j->level, j->strategy, j->hv, NULL, 0, &out_len, NULL, &ad);
            if (rc != ZSO_OK && rc != ZSO_BUF_ERROR) j->rc = rc;
            j->out_bytes = out_len;
        } else {
            uint64_t dec = 0;
            for (size_t k = j->seg_begin; k < j->seg_end; k++) {
                const size_t lo = k ? (size_t)j->seg_ends[k - 1] : j->first_offset, hi = (size_t)j->seg_ends[k];
                if (j->kind == 1) {
                    z_stream zs;
                    memset(&zs, 0, sizeof zs);
                    if (inflateInit2(&zs, -15) != Z_OK) { j->rc = -1; break; }
                    zs.next_in = (Bytef*)(j->comp + lo);
                    zs.avail_in = (uInt)(hi - lo);
                    zs.next_out = j->scratch;
                    zs.avail_out = (uInt)j->seg_size;
                    int rc = inflate(&zs, Z_SYNC_FLUSH);
                    if (rc != Z_OK && rc != Z_STREAM_END && rc != Z_BUF_ERROR) j->rc = rc;
                    dec += zs.total_out;
                    inflateEnd(&zs);
                } else {
                    size_t out_len = 0, used = 0;
                    const char* msg = NULL;
                    int rc = zso_inflate_raw(j->comp + lo, hi - lo, j->scratch, j->seg_size, &out_len, &used, &msg);
                    if (rc < 0 && rc != ZSO_BUF_ERROR) j->rc = rc;   /* a segment ends at its marker, not at a final block */
                    dec += out_len;
                }
            }
            j->decoded = dec;
        }
    }
    j->t1 = now_s();
    return NULL;
}

static int run_jobs(job_t* jobs, int threads, double* seconds)
{
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof *th);
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)threads);
    for (int t = 0; t < threads; t++) {
        jobs[t].bar = &bar;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    double t0 = 1e300, t1 = 0;
    int rc = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].t0 < t0) t0 = jobs[t].t0;
        if (jobs[t].t1 > t1) t1 = jobs[t].t1;
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    pthread_barrier_destroy(&bar);
    free(th);
    *seconds = t1 - t0;
    return rc;
}

/* `reps` passes over `in` on `threads` pinned threads, thread t compressing its contiguous share of the segments as its own
 * full-flush-segmented stream.  seconds = first start to last finish; out_bytes = compressed bytes of one pass. */
int zso_bench_deflate_mt(const uint8_t* in, size_t n, size_t seg_size, int level, int strategy, int hash_variant, int threads,
                         int reps, double* seconds, uint64_t* out_bytes)
{
    if (threads < 1 || reps < 1 || !seg_size) return ZSO_STREAM_ERROR;
    const size_t nseg = (n + seg_size - 1) / seg_size;
    if ((size_t)threads > nseg) threads = (int)(nseg ? nseg : 1);
    job_t* jobs = (job_t*)calloc((size_t)threads, sizeof *jobs);
    const long ncores = sysconf(_SC_NPROCESSORS_ONLN);
    for (int t = 0; t < threads; t++) {
        const size_t b = nseg * (size_t)t / (size_t)threads, e = nseg * (size_t)(t + 1) / (size_t)threads;
        jobs[t].kind = 0;
        jobs[t].core = ncores > 0 ? (int)(t % ncores) : -1;
        jobs[t].reps = reps;
        jobs[t].in = in + b * seg_size;
        jobs[t].n = (e * seg_size < n ? e * seg_size : n) - b * seg_size;
        jobs[t].seg_size = seg_size;
        jobs[t].level = level;
        jobs[t].strategy = strategy;
        jobs[t].hv = hash_variant;
    }
    int rc = run_jobs(jobs, threads, seconds);
    uint64_t total = 0;
    for (int t = 0; t < threads; t++) total += jobs[t].out_bytes;
    if (out_bytes) *out_bytes = total;
    free(jobs);
    return rc;
}

/* `reps` passes of segment-by-segment raw inflate (every segment of a full-flush stream starts byte-aligned with an empty
 * history) on `threads` pinned threads; which = 1 system zlib, 2 the oracle's inflater.  decoded = bytes of one pass. */
int zso_bench_inflate_mt(const uint8_t* comp, const uint64_t* seg_ends, size_t nseg, size_t first_offset, size_t seg_size,
                         int which, int threads, int reps, double* seconds, uint64_t* decoded)
{
    if (threads < 1 || reps < 1 || !seg_size || (which != 1 && which != 2)) return ZSO_STREAM_ERROR;
    if ((size_t)threads > nseg) threads = (int)(nseg ? nseg : 1);
    job_t* jobs = (job_t*)calloc((size_t)threads, sizeof *jobs);
    const long ncores = sysconf(_SC_NPROCESSORS_ONLN);
    for (int t = 0; t < threads; t++) {
        jobs[t].kind = which;
        jobs[t].core = ncores > 0 ? (int)(t % ncores) : -1;
        jobs[t].reps = reps;
        jobs[t].comp = comp;
        jobs[t].seg_ends = seg_ends;
        jobs[t].seg_begin = nseg * (size_t)t / (size_t)threads;
        jobs[t].seg_end = nseg * (size_t)(t + 1) / (size_t)threads;
        jobs[t].first_offset = first_offset;
        jobs[t].seg_size = seg_size;
        jobs[t].scratch = (uint8_t*)malloc(seg_size + 64);
    }
    int rc = run_jobs(jobs, threads, seconds);
    uint64_t total = 0;
    for (int t = 0; t < threads; t++) {
        total += jobs[t].decoded;
        free(jobs[t].scratch);
    }
    if (decoded) *decoded = total;
    free(jobs);
    return rc;
}
