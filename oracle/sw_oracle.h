/*
 * oracle/sw_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's per-pair Smith-Waterman operator
 *   /root/reference/src/sw/SmithWaterman.java:62-92   (OptAlignments.call)
 *   /root/reference/src/sw/SmithWaterman.java:129-190 (ScoreMatrix.call)
 *   /root/reference/src/sw/SmithWaterman.java:217-252 (GetCellScore.call, ">=" cascade)
 *   /root/reference/src/sw/SmithWaterman.java:277-280 (InsDelScore.call)
 *   /root/reference/src/sw/SmithWaterman.java:309-318 (AlignmentScore.call)
 *   /root/reference/src/sw/SmithWaterman.java:354-436 (GetAlignment.call)
 * and of the per-reference reduction
 *   /root/reference/src/sw/Distribution.java:403-436  (MapRef.call)
 *   /root/reference/src/sw/Distribution.java:691-694  (MatchSiteComp)
 *
 * PARITY PINNING: the reference is Java 8 + Spark; there is no JVM in this image
 * and the reference ships no tests, golden vectors or fixtures for this path
 * (SURVEY.md section 8c).  The oracle is therefore pinned only by (1) the
 * hand-derivable known-answer vectors of SURVEY.md section 8c, re-derived here,
 * and (2) agreement with an independently written Python twin
 * (oracle/sw_twin.py).  "parity unpinned" by reference-run outputs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product library (libswb200.so)
 * never links or calls it.
 */
#ifndef SW_ORACLE_H
#define SW_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sw_oracle_result {
    int32_t  score;        /* maximum cell score (0 if a sequence is empty)        */
    int64_t  n_cells;      /* number of maximum-score cells, reference list order   */
    int32_t *cells;        /* 2*n_cells ints: (i, j) 1-based, i = read row, j = ref column */
    int32_t *beginning;    /* n_cells: 1-based ref column of first aligned column, 0 if none */
    int64_t *aln_off;      /* n_cells+1: offsets into ref_aln / read_aln            */
    char    *ref_aln;      /* concatenated aligned reference strings ('_' = gap)    */
    char    *read_aln;     /* concatenated aligned read strings                     */
} sw_oracle_result;

/* Literal restatement: full (m+1)x(n+1) score and alignment-type matrices. */
int sw_oracle_align(const char *ref, int64_t n, const char *read, int64_t m,
                    int32_t match, int32_t mismatch, int32_t gap,
                    sw_oracle_result *out);

/* Low-memory variant for pairs whose int matrices do not fit: two score rows + a 2-bit
 * plane per cell (0 = score is zero, 1/2/3 = a/i/d of a positive cell) -- all the walk
 * ever reads.  Must agree with sw_oracle_align wherever both run. Returns 0 on success. */
int sw_oracle_align_lowmem(const char *ref, int64_t n, const char *read, int64_t m,
                           int32_t match, int32_t mismatch, int32_t gap,
                           sw_oracle_result *out);

/* DistributedSW.OptAlignments semantics (strict '>' ties, diagonal-major list, stable sort by beginning). */
int sw_oracle_align_gt(const char *ref, int64_t n, const char *read, int64_t m,
                       int32_t match, int32_t mismatch, int32_t gap, sw_oracle_result *out);

/* Score + max-cell count only (no traceback), two-row memory. */
int sw_oracle_score(const char *ref, int64_t n, const char *read, int64_t m,
                    int32_t match, int32_t mismatch, int32_t gap,
                    int32_t *score, int64_t *n_cells);

void sw_oracle_free(sw_oracle_result *r);

/* FNV-1a style 64-bit digest of one pair's complete result
 * (score, every cell, every beginning, both strings). Used for sampled parity. */
uint64_t sw_oracle_digest(const sw_oracle_result *r);

/* ---- CPU baseline harness (oracle/cpu_baseline.c) ------------------------------
 * Runs reads x refs through the literal per-pair function with the reference's
 * work shape (fresh matrices per pair, init pass, fill, traceback of every max
 * cell, per-ref wrapping int32 total + stable sort of sites by beginning).
 * mode 0: "spark local[N]"-shaped  -- ref list cut in N contiguous slices
 *         [k*len/N,(k+1)*len/N), one thread per slice (Distribution.java:337-338)
 * mode 1: "threadedMetrics"-shaped -- dynamic queue of refs over N threads
 * Returns wall seconds; writes per-ref totals and a checksum of all results. */
double sw_cpu_baseline_run(const char *ref_bytes, const int64_t *ref_off, int64_t n_refs,
                           const char *read_bytes, const int64_t *read_off, int64_t n_reads,
                           int32_t match, int32_t mismatch, int32_t gap,
                           int32_t n_threads, int32_t mode,
                           int32_t *ref_totals /* n_refs, may be NULL */,
                           int32_t *pair_scores /* n_refs*n_reads, may be NULL */,
                           uint64_t *checksum /* may be NULL */);

#ifdef __cplusplus
}
#endif
#endif
