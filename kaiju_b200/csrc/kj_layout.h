// kj_layout.h -- HBM data layout of the index, taxonomy and tables (shared by host builder and kernels).
//
// The reference keeps the BWT as one byte per letter that also encodes a distance to the nearest 256-
// checkpoint (bwt/compactfmi.c:44-65) plus two checkpoint tables behind row pointers
// (bwt/fmicommon.h:60-73).  A rank query there costs two dependent pointer derefs plus a byte scan of
// ~18-30 bytes.  Rank values are a pure function of the decoded letters (SURVEY.md 8a exactness note a),
// so the device uses its own layout, built at load time from the decoded letters:
//
//   rank[c][b] : one record per (letter c, block b of 64 or 192 BWT positions; two layouts, below); the 192-row one:
//                { hdr = C[c] + #{p < 192 b : L[p] == c} (40 bit) + the popcounts of w0 and w0|w1 (2 x 8 bit),
//                  w0,w1,w2 = one-hot bitmap of L[p]==c }
//                -> FMindex(c,k) touches exactly ONE 32-byte DRAM sector and needs ONE 64-bit popcount.
//   letters    : the decoded BWT packed 12 letters (5 bit) per 64-bit word, read only by the SA walk.
//   sa_tax     : the sampled suffix array reduced to what classification needs -- the (compact) taxon of
//                the sequence each sampled suffix belongs to (bwt/suffixArray.h:47-51 + ConsumerThread.cpp:812-832).
//   seq_tax    : compact taxon per sequence number (used when the LF walk runs into a terminator, bwt.c:119).
//   tax_*      : taxonomy re-indexed densely: ids present in nodes.dmp get indices [0,n_present), DB taxa
//                absent from nodes.dmp get the indices after that (depth 0 marks "absent", util.cpp:206).
#pragma once
#include <stdint.h>

// Two record layouts, chosen per index at load time:
//   narrow (bwtlen < 2^32, 32-bit interval kernels): 16-byte records per 64 rows  { hdr = C[c] + #c before the block, w0 = one-hot bitmap }
//          -> one 16-byte load + one popcount per rank query, no division; 5.25 B per BWT row (measured +6 % over the 192-row layout)
//   wide   (bwtlen >= 2^32, 64-bit interval kernels): 32-byte records per 192 rows { hdr (40-bit count | popc(w0) | popc(w0)+popc(w1)), w0, w1, w2 }
//          -> 3.5 B per row: a refseq_ref-scale index (2.7e10 rows) takes 95 GB of the 180 GB HBM instead of 142 GB
#define KJ_RANK_ROWS_NARROW 64
#define KJ_RANK_ROWS_WIDE 192
#define KJ_RANK_WORDS_NARROW 2
#define KJ_RANK_WORDS_WIDE 4
static inline uint32_t kj_rank_rows(int wide) { return wide ? KJ_RANK_ROWS_WIDE : KJ_RANK_ROWS_NARROW; }
static inline uint32_t kj_rank_words(int wide) { return wide ? KJ_RANK_WORDS_WIDE : KJ_RANK_WORDS_NARROW; }
#define KJ_LETTERS_PER_WORD 12
#define KJ_MAX_ALEN 24
#define KJ_MAX_IDS 21              // max_match_ids = 20 -> the set holds at most 21 (ConsumerThread.cpp:805)
#define KJ_MAX_BEST_SI 20          // max_matches_SI (Config.hpp:35)
#define KJ_TAX_BAD 0xffffffffu     // "bad number" database name (ConsumerThread.cpp:817-820): skipped
#define KJ_MAX_MM 8                // max supported -e
#define KJ_SEG_WINDOW 12

// wide records: hdr = cnt (40 bit: C[c] + #c before the block) | popc(w0) << 40 | (popc(w0)+popc(w1)) << 48 ; w0..w2 = one-hot bitmap.
#define KJ_CNT_MASK 0xffffffffffull
#define KJ_P1_SHIFT 40
#define KJ_P2_SHIFT 48

struct KjKmer { uint64_t lo, hi; };    // SA interval of a k-mer (empty if lo >= hi), indexes >= 2^32
struct KjKmer32 { uint32_t lo, hi; };  // same, for indexes with bwtlen < 2^32

struct KjTables {
    uint8_t codon_aa[64];            // (n0<<4|n1<<2|n2) -> alphabet index, 0 = stop  (ConsumerThread.cpp:117-181 + sequence.c:68-97)
    int8_t b62[KJ_MAX_ALEN][KJ_MAX_ALEN];   // BLOSUM62 indexed by ALPHABET index        (ConsumerThread.cpp:88-107)
    uint8_t subst[KJ_MAX_ALEN][20];  // substitution try-order per residue, alphabet indices (ConsumerThread.cpp:10-30)
    int32_t seg_logfix[KJ_SEG_WINDOW + 1];  // round(2^24 log2(12/c)) : entropy of a 12-window in fixed point
    int32_t seg_locut_fix, seg_hicut_fix;   // 12*2.2*2^24, 12*2.5*2^24 (margins checked on host against FP64)
    char letters[KJ_MAX_ALEN];       // alphabet index -> letter (fragment strings of the verbose output)
    uint8_t aa_index[32];            // protein input: upper-case letter - 'A' -> alphabet index, 0 = splits the read (ConsumerThread.cpp:664)
};

struct KjDevIndex {
    const uint64_t* rank; uint64_t nb;          // [alen][nb] records of 2 (narrow) or 4 (wide) 64-bit words
    const uint64_t* rank_base[KJ_MAX_ALEN];     // records of letter c (saves the multiply in the inner loop)
    const uint64_t* letters;
    uint64_t bwtlen; int alen;
    uint64_t C[KJ_MAX_ALEN + 1];                // C[c] = first SA row of letter c; C[alen] = bwtlen
    const uint32_t* sa_tax; const uint32_t* seq_tax;
    const uint32_t* sa_acc; const uint32_t* seq_acc;     // accession rank per sampled suffix / sequence (NULL unless the index view carried seq_accession)
    uint64_t sa_check; int sa_exp; int64_t sa_bias; uint64_t n_sa; uint32_t nseq;
    const uint32_t* tax_parent; const uint32_t* tax_depth; const uint64_t* tax_id; uint32_t n_tax;
    const double* lnfact; int n_lnfact;
    const void* kmer; int kmer_k;               // direct-address table of k-mer intervals (KjKmer32 if !wide else KjKmer; 0 = off)
    int wide;                                   // 1 if bwtlen >= 2^32 (64-bit interval arithmetic in the kernels)
    int mono;                                   // 1: true FM index (match starts monotone in the end position); 0: the reference's checkpoint quirk applies (no chain bounds)
    uint64_t quirk_lo; const uint64_t* quirk_d;         // rows k >= quirk_lo: FMindex(c,k) -= quirk_d[c] (reference checkpoint quirk, ~0 = none; [KJ_MAX_ALEN] in global memory)
    const KjTables* tables;
};

struct KjRunParams {
    int mode;                       // 0 MEM, 1 GREEDY
    uint32_t m, e, min_score, seed_length;
    int use_evalue, seg, protein, name_mode;
    // E-value gate as an integer threshold: ev_breaks[k] = the largest query length (double) for which score k passes,
    // so the minimal passing score of a read is the number of breaks below its query length (kj_build_evalue_breaks)
    const double* ev_breaks; uint32_t n_ev_breaks;
    // per-warp scratch geometry
    uint32_t max_len;               // longest read (bases) in the batch, rounded up
    uint32_t max_frag;              // longest fragment (residues)
    uint32_t item_cap;              // fragment-queue capacity per warp
    uint32_t kept_cap_smem;         // winners kept in shared memory before spilling
    uint32_t scratch_entries;       // global spill entries per warp
    uint32_t variant_cap;           // Greedy: entries of the per-warp substituted-variant ring
    uint32_t ws_global;             // 1: the per-warp work space lives in global memory (reads too long for shared memory)
    uint32_t stage;                 // 1: the bases of the reads a warp has claimed are brought into shared memory with one bulk copy per mate (kj_device.cu)
};
