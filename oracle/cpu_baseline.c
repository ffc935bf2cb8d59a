/*
 * oracle/cpu_baseline.c -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Multi-threaded CPU harness over the literal per-pair oracle, shaped like the
 * reference's "parallelize reference set" map:
 *   /root/reference/src/sw/Distribution.java:403-436 (MapRef.call: for every read,
 *   OptAlignments.call; wrapping int total; addAll; stable sort by beginning)
 *   /root/reference/src/sw/Distribution.java:337-338 (parallelize + mapToPair)
 *   /root/reference/src/sw/Distribution.java:714-724 (CombineReadsToRef: one element per ref)
 * It is a C restatement, not the JVM: the Java code allocates ~12 objects per
 * cell (SmithWaterman.java:162-174, 227-251) and would be slower than this.
 */
#include "sw_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
    const char *ref_bytes; const int64_t *ref_off; int64_t n_refs;
    const char *read_bytes; const int64_t *read_off; int64_t n_reads;
    int32_t match, mismatch, gap;
    int32_t *ref_totals; int32_t *pair_scores;
    /* work assignment */
    int mode; int64_t lo, hi;           /* mode 0: contiguous slice */
    int64_t *next;                      /* mode 1: shared counter   */
    pthread_mutex_t *mu;
    uint64_t checksum;
} job_t;

typedef struct { int32_t beginning; uint64_t digest; } site_t;

/* stable merge sort on beginning (java.util.Collections.sort is a stable merge
 * sort; comparator Distribution.java:691-694) */
static void site_sort(site_t *a, site_t *tmp, int64_t n)
{
    if (n < 2) return;
    int64_t h = n / 2;
    site_sort(a, tmp, h); site_sort(a + h, tmp, n - h);
    int64_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].beginning < a[i].beginning) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, (size_t)n * sizeof(site_t));
}

static uint64_t one_ref(job_t *jb, int64_t r)
{
    const char *ref = jb->ref_bytes + jb->ref_off[r];
    const int64_t n = jb->ref_off[r + 1] - jb->ref_off[r];
    uint32_t total = 0;
    site_t *sites = NULL; int64_t ns = 0, cap = 0;
    for (int64_t q = 0; q < jb->n_reads; ++q) {
        const char *read = jb->read_bytes + jb->read_off[q];
        const int64_t m = jb->read_off[q + 1] - jb->read_off[q];
        sw_oracle_result res;
        if (sw_oracle_align(ref, n, read, m, jb->match, jb->mismatch, jb->gap, &res)) continue;
        total += (uint32_t)res.score;                       /* Distribution.java:424 */
        if (jb->pair_scores) jb->pair_scores[r * jb->n_reads + q] = res.score;
        if (ns + res.n_cells > cap) {
            cap = (ns + res.n_cells) * 2 + 16;
            sites = (site_t *)realloc(sites, (size_t)cap * sizeof(site_t));
        }
        for (int64_t k = 0; k < res.n_cells; ++k) {          /* :425 addAll */
            uint64_t h = 0xcbf29ce484222325ULL;
            const int64_t len = res.aln_off[k + 1] - res.aln_off[k];
            const char *p = res.ref_aln + res.aln_off[k], *s = res.read_aln + res.aln_off[k];
            for (int64_t c = 0; c < len; ++c) { h ^= (unsigned char)p[c]; h *= 0x100000001b3ULL; }
            for (int64_t c = 0; c < len; ++c) { h ^= (unsigned char)s[c]; h *= 0x100000001b3ULL; }
            sites[ns].beginning = res.beginning[k]; sites[ns].digest = h; ++ns;
        }
        sw_oracle_free(&res);
    }
    uint64_t cs = 0;
    if (ns) {
        site_t *tmp = (site_t *)malloc((size_t)ns * sizeof(site_t));
        site_sort(sites, tmp, ns);                            /* :428 */
        free(tmp);
        for (int64_t k = 0; k < ns; ++k)
            cs = cs * 0x9E3779B97F4A7C15ULL + sites[k].digest + (uint64_t)(uint32_t)sites[k].beginning;
    }
    free(sites);
    if (jb->ref_totals) jb->ref_totals[r] = (int32_t)total;
    return cs ^ ((uint64_t)total << 32) ^ (uint64_t)r;
}

static void *worker(void *arg)
{
    job_t *jb = (job_t *)arg;
    uint64_t cs = 0;
    if (jb->mode == 0) {
        for (int64_t r = jb->lo; r < jb->hi; ++r) cs += one_ref(jb, r);
    } else {
        for (;;) {
            pthread_mutex_lock(jb->mu);
            int64_t r = (*jb->next)++;
            pthread_mutex_unlock(jb->mu);
            if (r >= jb->n_refs) break;
            cs += one_ref(jb, r);
        }
    }
    jb->checksum = cs;
    return NULL;
}

double sw_cpu_baseline_run(const char *ref_bytes, const int64_t *ref_off, int64_t n_refs,
                           const char *read_bytes, const int64_t *read_off, int64_t n_reads,
                           int32_t match, int32_t mismatch, int32_t gap,
                           int32_t n_threads, int32_t mode,
                           int32_t *ref_totals, int32_t *pair_scores, uint64_t *checksum)
{
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc((size_t)n_threads * sizeof(pthread_t));
    job_t *jobs = (job_t *)calloc((size_t)n_threads, sizeof(job_t));
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    int64_t next = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 0; k < n_threads; ++k) {
        job_t *jb = &jobs[k];
        jb->ref_bytes = ref_bytes; jb->ref_off = ref_off; jb->n_refs = n_refs;
        jb->read_bytes = read_bytes; jb->read_off = read_off; jb->n_reads = n_reads;
        jb->match = match; jb->mismatch = mismatch; jb->gap = gap;
        jb->ref_totals = ref_totals; jb->pair_scores = pair_scores;
        jb->mode = mode; jb->next = &next; jb->mu = &mu;
        /* ParallelCollectionRDD slicing: slice k = [k*len/N, (k+1)*len/N) */
        jb->lo = (int64_t)k * n_refs / n_threads;
        jb->hi = (int64_t)(k + 1) * n_refs / n_threads;
        pthread_create(&th[k], NULL, worker, jb);
    }
    uint64_t cs = 0;
    for (int k = 0; k < n_threads; ++k) { pthread_join(th[k], NULL); cs += jobs[k].checksum; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (checksum) *checksum = cs;
    free(th); free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
