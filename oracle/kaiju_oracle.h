/* kaiju_oracle.h -- TEST INFRASTRUCTURE ONLY (see kaiju_oracle.c). */
#ifndef KAIJU_ORACLE_H
#define KAIJU_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ko_index ko_index;
typedef struct ko_tax ko_tax;

typedef struct {
    int mode;                 /* 0 = MEM, 1 = GREEDY               (Config.hpp:29,33) */
    uint32_t min_fragment_length; /* -m, default 11                (Config.hpp:41) */
    uint32_t mismatches;      /* -e, default 3                     (Config.hpp:42) */
    uint32_t min_score;       /* -s, default 65                    (Config.hpp:43) */
    uint32_t seed_length;     /* -l, default 7                     (Config.hpp:44) */
    int use_evalue;           /* greedy default 1, mem forces 0    (kaiju.cpp:76-86) */
    double min_evalue;        /* -E, default 0.01                  (Config.hpp:46) */
    int seg;                  /* -x/-X, default 1                  (Config.hpp:39) */
    int input_is_protein;     /* -p                                (Config.hpp:40) */
} ko_params;

/* work counters of the REFERENCE algorithm (roofline numerator) */
typedef struct {
    uint64_t reads, initial_si, update_si, fmindex, scanned_bytes, get_suffix, lf_steps;
    uint64_t seg_calls, seg_hits, fragments_initial, fragments_searched, bases, classified;
} ko_counters;

ko_index *ko_index_load(const char *fmi_path);
void ko_index_free(ko_index *);
int64_t ko_index_bwtlen(const ko_index *);
int ko_index_alen(const ko_index *);
int ko_index_nseq(const ko_index *);
uint64_t ko_index_seq_taxon(const ko_index *, int iseq);
/* FMindex restatement: C[c] + rank_c(BWT[0..k))            (compactfmi.c:267-307) */
int64_t ko_fmindex(const ko_index *, int c, int64_t k);
/* get_suffix restatement -> sequence number + position      (bwt.c:105-121) */
void ko_get_suffix(const ko_index *, int64_t k, int *iseq, int64_t *pos);

ko_tax *ko_tax_load(const char *nodes_dmp_path);
void ko_tax_free(ko_tax *);
/* lca_from_ids restatement over a sorted id set             (util.cpp:194-263) */
uint64_t ko_lca(const ko_tax *, const uint64_t *sorted_ids, int n);

/* SEG restatement (SeqBufferSeg with SegParametersNewAa + overlaps=TRUE; blast_seg.c:2278-2332,
 * Config.cpp:24-27).  aa = upper-case one-letter residues.  Writes up to max [left,right] pairs in
 * list order; returns the number of regions. */
int ko_seg(const char *aa, int len, int *left, int *right, int max);
/* ln(n!) table exactly as the reference literals (blast_seg.c:53-1306): "%.6f" rounding of lgamma */
double ko_lnfact(int n);

/* Classify one read item (ConsumerThread::doWork body, ConsumerThread.cpp:630-749).
 * seq2 == NULL -> single-end.  Returns taxon id (0 = unclassified); *best = match length (MEM) or
 * score (Greedy), 0 if unclassified.  ids_out (optional, >= 32 slots) receives the sorted match id set. */
uint64_t ko_classify(const ko_index *, const ko_tax *, const ko_params *,
                     const char *seq1, int len1, const char *seq2, int len2,
                     uint32_t *best, uint64_t *ids_out, int *n_ids_out, ko_counters *ctr);

/* Batch form over concatenated sequences + offsets (n+1), as the product C ABI takes them. */
void ko_classify_batch(const ko_index *, const ko_tax *, const ko_params *,
                       const char *seq1, const uint64_t *off1, const char *seq2, const uint64_t *off2,
                       uint64_t n, uint64_t *taxon_out, uint32_t *best_out, ko_counters *ctr);

/* Six-frame fragments of one mate in queue-insertion order (getAllFragmentsBits, ConsumerThread.cpp:190-270).
 * Writes NUL-separated fragments into buf; returns the count. */
int ko_fragments(const char *seq, int len, int min_len, char *buf, int bufsize);

#ifdef __cplusplus
}
#endif
#endif
