/* pair_bench.c -- pairs/s of the UNCHANGED-driver path measured natively (pthreads on the C ABI, no interpreter in the
 * way): T host threads each issue `OptAlignments.call`-shaped single-pair requests (Distribution.java:419-426: one
 * MapRef task per reference, reads in file order) through
 *   mode "queued":   swb_align_pair           (the context's coalescing submission queue)
 *   mode "single":   swb_refset_load(1) + swb_align(1) + frees, one launch sequence per pair
 * and, for scale, one batched swb_align over the same pair set ("file").  Synthetic RefSeq-shaped references
 * (log-normal lengths, median 1,609) and 150 bp reads from a fixed LCG; every result is consumed (score, cell count,
 * first alignment materialised) so the marshalling a host shim would do is inside the clock.
 *
 *   gcc -O2 -pthread -Iinclude tools/pair_bench.c -Lsparksmithwaterman_b200/_lib -lswb200 -lm -o /tmp/pair_bench
 *   LD_LIBRARY_PATH=sparksmithwaterman_b200/_lib /tmp/pair_bench [refs=256] [reads=16] [threads=1,4,16,64]
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "swb200.h"

static uint64_t lcg_state = 20151001u;
static uint32_t lcg(void) { lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(lcg_state >> 33); }
static double unif(void) { return (lcg() + 0.5) / 2147483648.0; }
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static int n_refs = 256, n_reads = 16;
static char **refs, **reads;
static int64_t *ref_len, *read_len;
static swb_ctx *ctx;
static int mode_queued;
static volatile int64_t next_ref;
static int64_t checksum_total;
static pthread_mutex_t sum_mu = PTHREAD_MUTEX_INITIALIZER;

static int64_t consume(const swb_result *res, const char *ref, int64_t rl, const char *read, int64_t ql)
{
    int64_t sum = swb_result_scores(res)[0];
    const int64_t cnt = swb_result_pair_cell_count(res, 0);
    sum += cnt;
    if (swb_result_scores(res)[0] > 0 && cnt > 0) {
        int32_t i, j, b, len;
        static __thread char a[4096], c[4096];
        if (swb_result_pair_cell(res, 0, 0, &i, &j, &b, &len) == 0 && len < 4095 &&
            swb_result_materialize(res, swb_result_cell_offsets(res)[0], ref, rl, read, ql, a, c, 4096) == 0)
            sum += b + a[0];
    }
    return sum;
}

static void *worker(void *arg)
{
    (void)arg;
    int64_t sum = 0;
    for (;;) {
        const int64_t r = __sync_fetch_and_add(&next_ref, 1);         /* one MapRef task per reference */
        if (r >= n_refs) break;
        for (int q = 0; q < n_reads; ++q) {
            swb_result *res = 0;
            if (mode_queued) {
                if (swb_align_pair(ctx, refs[r], ref_len[r], reads[q], read_len[q], 5, -3, -4, 0, &res)) { fprintf(stderr, "%s\n", swb_last_error()); exit(1); }
            } else {
                swb_refset *rs = 0;
                const int64_t ro[2] = {0, ref_len[r]}, qo[2] = {0, read_len[q]};
                if (swb_refset_load(ctx, 1, refs[r], ro, &rs) || swb_align(ctx, rs, 1, reads[q], qo, 5, -3, -4, 0, &res)) { fprintf(stderr, "%s\n", swb_last_error()); exit(1); }
                swb_refset_free(rs);
            }
            sum += consume(res, refs[r], ref_len[r], reads[q], read_len[q]);
            swb_result_free(res);
        }
    }
    pthread_mutex_lock(&sum_mu); checksum_total += sum; pthread_mutex_unlock(&sum_mu);
    return 0;
}

static double run(int threads, int queued, int64_t *checksum)
{
    pthread_t th[256];
    mode_queued = queued; next_ref = 0; checksum_total = 0;
    const double t0 = now_s();
    for (int k = 0; k < threads; ++k) pthread_create(&th[k], 0, worker, 0);
    for (int k = 0; k < threads; ++k) pthread_join(th[k], 0);
    const double dt = now_s() - t0;
    *checksum = checksum_total;
    return dt;
}

int main(int argc, char **argv)
{
    if (argc > 1) n_refs = atoi(argv[1]);
    if (argc > 2) n_reads = atoi(argv[2]);
    const char *tl = argc > 3 ? argv[3] : "1,4,16,64";
    if (swb_create(0, 0, &ctx)) { fprintf(stderr, "%s\n", swb_last_error()); return 1; }
    refs = malloc(sizeof(char *) * n_refs); ref_len = malloc(8 * n_refs);
    reads = malloc(sizeof(char *) * n_reads); read_len = malloc(8 * n_reads);
    for (int r = 0; r < n_refs; ++r) {
        const double g = sqrt(-2.0 * log(unif())) * cos(6.283185307179586 * unif());
        int64_t n = (int64_t)(exp(7.3834 + 0.7674 * g) + 0.5);
        if (n < 50) n = 50;
        if (n > 200000) n = 200000;
        refs[r] = malloc(n + 1); ref_len[r] = n;
        for (int64_t k = 0; k < n; ++k) refs[r][k] = "ACGT"[lcg() & 3];
    }
    for (int q = 0; q < n_reads; ++q) {
        reads[q] = malloc(151); read_len[q] = 150;
        const int r = lcg() % n_refs;
        const int64_t at = ref_len[r] > 150 ? lcg() % (ref_len[r] - 150) : 0;
        for (int k = 0; k < 150; ++k)
            reads[q][k] = ((q & 1) && ref_len[r] >= 150 && lcg() % 50) ? refs[r][at + k] : "ACGT"[lcg() & 3];   /* half planted, 2 % substitutions */
    }
    const int64_t n_pairs = (int64_t)n_refs * n_reads;
    int64_t sum_q = 0, sum_s = 0, sum_f = 0;
    run(4, 1, &sum_q);                                                 /* warm: pools, pinned buffers */
    printf("{\"refs\": %d, \"reads\": %d, \"pairs\": %lld, \"runs\": [", n_refs, n_reads, (long long)n_pairs);
    char *list = strdup(tl);
    int first = 1;
    for (char *tok = strtok(list, ","); tok; tok = strtok(0, ",")) {
        const int T = atoi(tok);
        if (T < 1 || T > 256) continue;
        int64_t q0[3], q1[3];
        swb_queue_stats(ctx, q0, 3);
        const double dq = run(T, 1, &sum_q);
        swb_queue_stats(ctx, q1, 3);
        const double ds = run(T, 0, &sum_s);
        printf("%s{\"threads\": %d, \"queued_pairs_per_s\": %.1f, \"queued_batches\": %lld, \"queued_largest_batch\": %lld, \"single_pairs_per_s\": %.1f, \"equal\": %s}",
               first ? "" : ", ", T, n_pairs / dq, (long long)(q1[1] - q0[1]), (long long)q1[2], n_pairs / ds, sum_q == sum_s ? "true" : "false");
        first = 0;
        fflush(stdout);
    }
    /* the batched call over the same pair set, results consumed the same way */
    {
        int64_t tot = 0, *ro = malloc(8 * (n_refs + 1)), *qo = malloc(8 * (n_reads + 1));
        ro[0] = 0; for (int r = 0; r < n_refs; ++r) ro[r + 1] = ro[r] + ref_len[r];
        qo[0] = 0; for (int q = 0; q < n_reads; ++q) qo[q + 1] = qo[q] + read_len[q];
        char *rb = malloc(ro[n_refs] + 1), *qb = malloc(qo[n_reads] + 1);
        for (int r = 0; r < n_refs; ++r) memcpy(rb + ro[r], refs[r], ref_len[r]);
        for (int q = 0; q < n_reads; ++q) memcpy(qb + qo[q], reads[q], read_len[q]);
        double best = 1e30;
        for (int rep = 0; rep < 3; ++rep) {
            const double t0 = now_s();
            swb_refset *rs = 0; swb_result *res = 0;
            if (swb_refset_load(ctx, n_refs, rb, ro, &rs) || swb_align(ctx, rs, n_reads, qb, qo, 5, -3, -4, 0, &res)) { fprintf(stderr, "%s\n", swb_last_error()); return 1; }
            tot = 0;
            const int32_t *sc = swb_result_scores(res);
            const int64_t *co = swb_result_cell_offsets(res);
            char a[4096], c[4096];
            for (int64_t p = 0; p < n_pairs; ++p) {
                const int64_t cnt = swb_result_pair_cell_count(res, p);
                tot += sc[p] + cnt;
                if (sc[p] > 0 && cnt > 0) {
                    int32_t i, j, b, len;
                    const int64_t r = p / n_reads, q = p % n_reads;
                    if (swb_result_pair_cell(res, p, 0, &i, &j, &b, &len) == 0 && len < 4095 &&
                        swb_result_materialize(res, co[p], refs[r], ref_len[r], reads[q], read_len[q], a, c, 4096) == 0)
                        tot += b + a[0];
                }
            }
            swb_result_free(res); swb_refset_free(rs);
            const double dt = now_s() - t0;
            if (dt < best) best = dt;
        }
        sum_f = tot;
        printf("], \"file_pairs_per_s\": %.1f, \"file_equal\": %s}\n", n_pairs / best, sum_f == sum_q ? "true" : "false");
    }
    swb_destroy(ctx);
    return 0;
}
