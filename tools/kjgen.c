/* kjgen -- deterministic synthetic workload generator for the Kaiju hot path.
 *
 * NOT product code and NOT the oracle: it only makes inputs (protein DB FASTA + nodes.dmp, and reads).
 * Everything is a pure function of (seed, index), so any slice of the workload can be regenerated
 * anywhere (this container, the GPU box) and in parallel.
 *
 *   DB    : nprot proteins in families (~23 members, 0-30 % substitution divergence, every 997th family
 *           120 identical copies -> exercises the 21-id cap), log-normal lengths (mean ~290 aa),
 *           UniProt-like residue frequencies, ~3 % of families carry a low-complexity insert (SEG),
 *           header ">P<i>_<taxid>";  taxonomy: 7 levels, root 1 (self-parent), ~nprot/23 leaves,
 *           every 4999th protein gets a taxid that is absent from nodes.dmp.
 *   reads : insert 350 bp; 70 % back-translated DB windows (random synonymous codons, random frame
 *           offset, 1 % per-base substitutions), 30 % uniform random DNA, 2 % carry one 'N',
 *           50 % reverse-complemented, mate2 = revcomp of the insert's last L bases.
 *           ~0.1 % adversarial tail: homopolymer / dinucleotide repeats, very short reads,
 *           lower-case reads, one short mate.
 *
 * Build:  gcc -O2 -fPIC -shared -fopenmp -o libkjgen.so kjgen.c -lm      (library, used via ctypes)
 *         gcc -O2 -fopenmp -DKJGEN_MAIN -o kjgen kjgen.c -lm             (CLI)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r) {            /* splitmix64 */
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline rng_t rng_make(uint64_t seed, uint64_t stream) {
    rng_t r; r.s = seed * 0xD1342543DE82EF95ull + stream * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    rng_next(&r); rng_next(&r); return r;
}
static inline uint32_t rng_below(rng_t *r, uint32_t n) { return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32); }
static inline double rng_unif(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline double rng_norm(rng_t *r) {
    double u1 = rng_unif(r), u2 = rng_unif(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

static const char AA[21] = "ARNDCQEGHILKMFPSTWYV";
static const double AAFREQ[20] = {8.25,5.53,4.06,5.45,1.37,3.93,6.75,7.07,2.27,5.96,9.66,5.84,2.42,3.86,4.70,6.56,5.34,1.08,2.92,6.87};
static double aacdf[20];
static void init_cdf(void) {
    double s = 0, t = 0; for (int i = 0; i < 20; i++) s += AAFREQ[i];
    for (int i = 0; i < 20; i++) { t += AAFREQ[i] / s; aacdf[i] = t; }
    aacdf[19] = 1.0;
}
static inline char rand_aa(rng_t *r) {
    double u = rng_unif(r); int i = 0; while (u > aacdf[i]) i++; return AA[i];
}

/* standard genetic code, codons per residue */
static const char *CODONS[26];
static void init_codons(void) {
    memset(CODONS, 0, sizeof(CODONS));
    CODONS['A'-'A'] = "GCAGCCGCGGCT"; CODONS['R'-'A'] = "CGACGCCGGCGTAGAAGG"; CODONS['N'-'A'] = "AACAAT";
    CODONS['D'-'A'] = "GACGAT"; CODONS['C'-'A'] = "TGCTGT"; CODONS['Q'-'A'] = "CAACAG"; CODONS['E'-'A'] = "GAAGAG";
    CODONS['G'-'A'] = "GGAGGCGGGGGT"; CODONS['H'-'A'] = "CACCAT"; CODONS['I'-'A'] = "ATAATCATT";
    CODONS['L'-'A'] = "TTATTGCTACTCCTGCTT"; CODONS['K'-'A'] = "AAAAAG"; CODONS['M'-'A'] = "ATG";
    CODONS['F'-'A'] = "TTCTTT"; CODONS['P'-'A'] = "CCACCCCCGCCT"; CODONS['S'-'A'] = "TCATCCTCGTCTAGCAGT";
    CODONS['T'-'A'] = "ACAACCACGACT"; CODONS['W'-'A'] = "TGG"; CODONS['Y'-'A'] = "TACTAT"; CODONS['V'-'A'] = "GTAGTCGTGGTT";
}

/* ------------------------------------------------------------------ DB ---- */
typedef struct {
    uint64_t seed;
    int64_t nprot;
    char *seq;          /* concatenated residues */
    int64_t *off;       /* nprot+1 */
    uint64_t *taxid;    /* per protein */
    /* taxonomy */
    int64_t nnodes;
    uint64_t *node_id, *node_parent;
} kjgen_db;

#define NLEVEL 7
#define FAM_SIZE 23

void kjgen_db_free(kjgen_db *db) {
    if (!db) return;
    free(db->seq); free(db->off); free(db->taxid); free(db->node_id); free(db->node_parent); free(db);
}

kjgen_db *kjgen_db_create(int64_t nprot, uint64_t seed) {
    init_cdf(); init_codons();
    kjgen_db *db = (kjgen_db *)calloc(1, sizeof(kjgen_db));
    db->seed = seed; db->nprot = nprot;
    int64_t nfam = (nprot + FAM_SIZE - 1) / FAM_SIZE;
    /* ---- taxonomy: level sizes grow geometrically to ~nfam*? leaves */
    int64_t nleaf = nfam > 8 ? nfam : 8;
    int64_t lsize[NLEVEL]; lsize[0] = 1; lsize[NLEVEL-1] = nleaf;
    for (int l = 1; l < NLEVEL-1; l++) {
        double f = pow((double)nleaf, (double)l / (NLEVEL-1));
        lsize[l] = (int64_t)(f < 2 ? 2 : f);
    }
    int64_t lstart[NLEVEL+1]; lstart[0] = 0;
    for (int l = 0; l < NLEVEL; l++) lstart[l+1] = lstart[l] + lsize[l];
    db->nnodes = lstart[NLEVEL];
    db->node_id = (uint64_t *)malloc(sizeof(uint64_t) * db->nnodes);
    db->node_parent = (uint64_t *)malloc(sizeof(uint64_t) * db->nnodes);
    rng_t tr = rng_make(seed, 0x7A78);
    /* ids: root = 1, then increasing with random gaps (like NCBI) */
    uint64_t id = 1;
    for (int64_t i = 0; i < db->nnodes; i++) { db->node_id[i] = id; id += 1 + rng_below(&tr, 3); }
    db->node_parent[0] = 1;
    for (int l = 1; l < NLEVEL; l++)
        for (int64_t i = lstart[l]; i < lstart[l+1]; i++) {
            /* children are spread evenly over the previous level, jittered */
            int64_t p = lstart[l-1] + (int64_t)(((double)(i - lstart[l]) / lsize[l]) * lsize[l-1]);
            if (rng_below(&tr, 8) == 0) p = lstart[l-1] + rng_below(&tr, (uint32_t)lsize[l-1]);
            db->node_parent[i] = db->node_id[p];
        }
    /* ---- proteins */
    db->off = (int64_t *)malloc(sizeof(int64_t) * (nprot + 1));
    db->taxid = (uint64_t *)malloc(sizeof(uint64_t) * nprot);
    /* pass 1: lengths (per family root length, members share it) */
    int64_t *famlen = (int64_t *)malloc(sizeof(int64_t) * nfam);
    for (int64_t f = 0; f < nfam; f++) {
        rng_t r = rng_make(seed, 0x1000000ull + (uint64_t)f);
        double L = exp(5.49 + 0.6 * rng_norm(&r));
        if (L < 30) L = 30; if (L > 3000) L = 3000;
        famlen[f] = (int64_t)L;
    }
    db->off[0] = 0;
    for (int64_t p = 0; p < nprot; p++) db->off[p+1] = db->off[p] + famlen[p / FAM_SIZE];
    db->seq = (char *)malloc(db->off[nprot] + 1);
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t f = 0; f < nfam; f++) {
        rng_t r = rng_make(seed, 0x2000000ull + (uint64_t)f);
        int64_t L = famlen[f];
        char *root = (char *)malloc(L);
        for (int64_t k = 0; k < L; k++) root[k] = rand_aa(&r);
        if (rng_below(&r, 33) == 0 && L > 60) {          /* low-complexity insert */
            int64_t s = rng_below(&r, (uint32_t)(L - 40)), n = 15 + rng_below(&r, 20);
            char a = rand_aa(&r), b = rand_aa(&r), c = rand_aa(&r); int mode = rng_below(&r, 3);
            for (int64_t k = 0; k < n; k++) root[s+k] = mode == 0 ? a : mode == 1 ? ((k & 1) ? a : b) : (rng_below(&r, 3) == 0 ? c : (k & 1) ? a : b);
        }
        int identical = (f % 997 == 5);
        /* family home: a node one level above the leaves; members go to leaves nearby */
        int64_t home_leaf = lstart[NLEVEL-1] + (int64_t)(((double)f / nfam) * lsize[NLEVEL-1]);
        for (int m = 0; m < FAM_SIZE; m++) {
            int64_t p = f * FAM_SIZE + m; if (p >= nprot) break;
            char *dst = db->seq + db->off[p];
            memcpy(dst, root, L);
            if (!identical && m > 0) {
                double d = 0.3 * rng_unif(&r);
                for (int64_t k = 0; k < L; k++) if (rng_unif(&r) < d) dst[k] = rand_aa(&r);
            }
            int64_t leaf = home_leaf + (int64_t)rng_below(&r, 6) - 2;
            if (rng_below(&r, 10) == 0) leaf = lstart[NLEVEL-1] + rng_below(&r, (uint32_t)lsize[NLEVEL-1]);
            if (leaf < lstart[NLEVEL-1]) leaf = lstart[NLEVEL-1];
            if (leaf >= lstart[NLEVEL]) leaf = lstart[NLEVEL] - 1;
            uint64_t t = db->node_id[leaf];
            if (rng_below(&r, 12) == 0) {                  /* some proteins annotated at inner nodes */
                int up = 1 + rng_below(&r, 3);
                int64_t n = leaf; int lvl = NLEVEL-1;
                while (up-- > 0 && lvl > 1) {               /* find parent index: ids are sorted -> bsearch */
                    uint64_t pid = db->node_parent[n]; int64_t lo = lstart[lvl-1], hi = lstart[lvl]-1;
                    while (lo < hi) { int64_t mid = (lo + hi) / 2; if (db->node_id[mid] < pid) lo = mid + 1; else hi = mid; }
                    n = lo; lvl--;
                }
                t = db->node_id[n];
            }
            if (p % 4999 == 17) t = 9000000ull + (uint64_t)p;   /* absent from nodes.dmp */
            db->taxid[p] = t;
        }
        free(root);
    }
    free(famlen);
    return db;
}

int64_t kjgen_db_nprot(const kjgen_db *db) { return db->nprot; }
int64_t kjgen_db_nletters(const kjgen_db *db) { return db->off[db->nprot]; }

int kjgen_db_write(const kjgen_db *db, const char *faa_path, const char *nodes_path) {
    FILE *fp = fopen(faa_path, "w"); if (!fp) return -1;
    setvbuf(fp, NULL, _IOFBF, 1 << 22);
    for (int64_t p = 0; p < db->nprot; p++) {
        fprintf(fp, ">P%lld_%llu\n", (long long)p, (unsigned long long)db->taxid[p]);
        fwrite(db->seq + db->off[p], 1, (size_t)(db->off[p+1] - db->off[p]), fp);
        fputc('\n', fp);
    }
    fclose(fp);
    fp = fopen(nodes_path, "w"); if (!fp) return -1;
    for (int64_t i = 0; i < db->nnodes; i++)
        fprintf(fp, "%llu\t|\t%llu\t|\tno rank\t|\n", (unsigned long long)db->node_id[i], (unsigned long long)db->node_parent[i]);
    fclose(fp);
    return 0;
}

/* --------------------------------------------------------------- reads ---- */
static inline char comp(char c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
                 case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; default: return c; }
}
static void revcomp_inplace(char *s, int n) {
    for (int i = 0, j = n - 1; i <= j; i++, j--) { char a = comp(s[i]), b = comp(s[j]); s[i] = b; s[j] = a; }
}
#define INSERT 350

/* Generate read item `idx` (global index). out1/out2 must hold readlen bytes. Returns lengths via len1/len2. */
static void gen_item(const kjgen_db *db, uint64_t seed, uint64_t idx, int readlen, int paired,
                     char *out1, int *len1, char *out2, int *len2) {
    rng_t r = rng_make(seed, 0x5000000000ull + idx);
    char ins[INSERT + 8];
    int ilen = INSERT; if (ilen < readlen) ilen = readlen;
    if (ilen > INSERT) ilen = INSERT;                       /* readlen <= INSERT enforced by caller */
    static const char NUC[4] = {'A','C','G','T'};
    int from_db = rng_unif(&r) < 0.70;
    for (int k = 0; k < ilen; k++) ins[k] = NUC[rng_below(&r, 4)];
    if (from_db) {
        int64_t p = (int64_t)(rng_unif(&r) * db->nprot); if (p >= db->nprot) p = db->nprot - 1;
        int64_t L = db->off[p+1] - db->off[p];
        int naa = (ilen - 2) / 3;                            /* window in residues */
        int64_t s = 0; int n = naa;
        if (L > naa) s = (int64_t)(rng_unif(&r) * (L - naa + 1)); else n = (int)L;
        int frame = rng_below(&r, 3);
        int start = frame + (L > naa ? 0 : (int)rng_below(&r, (uint32_t)((ilen - frame - 3 * n) / 3 + 1)) * 3);
        const char *prot = db->seq + db->off[p] + s;
        for (int k = 0; k < n; k++) {
            const char *cs = CODONS[prot[k] - 'A']; int nc = (int)strlen(cs) / 3; const char *c = cs + 3 * rng_below(&r, nc);
            int o = start + 3 * k; if (o + 3 > ilen) break;
            ins[o] = c[0]; ins[o+1] = c[1]; ins[o+2] = c[2];
        }
        for (int k = 0; k < ilen; k++) if (rng_unif(&r) < 0.01) ins[k] = NUC[rng_below(&r, 4)];
    }
    if (rng_below(&r, 2)) revcomp_inplace(ins, ilen);
    int l1 = readlen, l2 = paired ? readlen : 0;
    memcpy(out1, ins, l1);
    if (paired) { memcpy(out2, ins + ilen - l2, l2); revcomp_inplace(out2, l2); }
    if (rng_unif(&r) < 0.02) out1[rng_below(&r, l1)] = 'N';
    if (paired && rng_unif(&r) < 0.02) out2[rng_below(&r, l2)] = 'N';
    /* adversarial tail */
    uint32_t adv = rng_below(&r, 5000);
    if (adv < 5) {
        switch (adv) {
        case 0: { char a = NUC[rng_below(&r, 4)]; int s = rng_below(&r, l1 / 2), n = l1 / 2; for (int k = 0; k < n; k++) out1[s+k] = a; } break;
        case 1: { char a = NUC[rng_below(&r, 4)], b = NUC[rng_below(&r, 4)]; for (int k = 0; k < l1; k++) out1[k] = (k & 1) ? a : b;
                  if (paired) { char t3[3] = {NUC[rng_below(&r,4)], NUC[rng_below(&r,4)], NUC[rng_below(&r,4)]}; for (int k = 0; k < l2; k++) out2[k] = t3[k % 3]; } } break;
        case 2: l1 = 5 + rng_below(&r, 30); if (paired) l2 = 5 + rng_below(&r, 30); break;
        case 3: for (int k = 0; k < l1; k++) out1[k] = (char)(out1[k] | 0x20); break;
        case 4: if (paired) l2 = 10 + rng_below(&r, 25); else l1 = 20 + rng_below(&r, 40); break;
        }
    }
    *len1 = l1; *len2 = l2;
}

/* Fill caller buffers with items [first, first+n): seq1/seq2 are n*readlen bytes (item i at i*readlen,
 * only the first len bytes valid), len1/len2 int32[n]. Deterministic and thread-parallel. */
void kjgen_reads(const kjgen_db *db, uint64_t seed, uint64_t first, int64_t n, int readlen, int paired,
                 char *seq1, int32_t *len1, char *seq2, int32_t *len2) {
    if (readlen > INSERT) readlen = INSERT;
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        int a = 0, b = 0; char tmp2[INSERT + 8];
        gen_item(db, seed, first + (uint64_t)i, readlen, paired, seq1 + i * readlen, &a, paired ? seq2 + i * readlen : tmp2, &b);
        len1[i] = a; if (paired) len2[i] = b;
    }
}

/* Compact variant: concatenated sequences + uint64 offsets (n+1 each), as the C ABI takes them.
 * Buffers must hold n*readlen bytes. */
void kjgen_reads_packed(const kjgen_db *db, uint64_t seed, uint64_t first, int64_t n, int readlen, int paired,
                        char *seq1, uint64_t *off1, char *seq2, uint64_t *off2) {
    int32_t *l1 = (int32_t *)malloc(sizeof(int32_t) * n), *l2 = (int32_t *)malloc(sizeof(int32_t) * n);
    char *t1 = (char *)malloc((size_t)n * readlen), *t2 = paired ? (char *)malloc((size_t)n * readlen) : NULL;
    kjgen_reads(db, seed, first, n, readlen, paired, t1, l1, t2, l2);
    uint64_t o1 = 0, o2 = 0;
    for (int64_t i = 0; i < n; i++) {
        off1[i] = o1; memcpy(seq1 + o1, t1 + i * readlen, l1[i]); o1 += l1[i];
        if (paired) { off2[i] = o2; memcpy(seq2 + o2, t2 + i * readlen, l2[i]); o2 += l2[i]; }
    }
    off1[n] = o1; if (paired) off2[n] = o2;
    free(l1); free(l2); free(t1); free(t2);
}

int kjgen_reads_write_fastq(const kjgen_db *db, uint64_t seed, uint64_t first, int64_t n, int readlen, int paired,
                            const char *fq1, const char *fq2) {
    FILE *f1 = fopen(fq1, "w"), *f2 = paired ? fopen(fq2, "w") : NULL;
    if (!f1 || (paired && !f2)) return -1;
    setvbuf(f1, NULL, _IOFBF, 1 << 22); if (f2) setvbuf(f2, NULL, _IOFBF, 1 << 22);
    const int64_t CH = 1 << 16;
    char *s1 = (char *)malloc(CH * readlen), *s2 = (char *)malloc(CH * readlen), *q = (char *)malloc(readlen + 1);
    int32_t *l1 = (int32_t *)malloc(sizeof(int32_t) * CH), *l2 = (int32_t *)malloc(sizeof(int32_t) * CH);
    memset(q, 'I', readlen);
    for (int64_t b = 0; b < n; b += CH) {
        int64_t m = n - b < CH ? n - b : CH;
        kjgen_reads(db, seed, first + b, m, readlen, paired, s1, l1, s2, l2);
        for (int64_t i = 0; i < m; i++) {
            fprintf(f1, "@r%llu/1\n%.*s\n+\n%.*s\n", (unsigned long long)(first + b + i), l1[i], s1 + i * readlen, l1[i], q);
            if (paired) fprintf(f2, "@r%llu/2\n%.*s\n+\n%.*s\n", (unsigned long long)(first + b + i), l2[i], s2 + i * readlen, l2[i], q);
        }
    }
    fclose(f1); if (f2) fclose(f2);
    free(s1); free(s2); free(q); free(l1); free(l2);
    return 0;
}

/* ---- variable-length inputs for the "any read length" and protein-input (-p) paths -------------------------------
 * Long DNA read: a concatenation of segments until the drawn length is reached; a segment is a back-translated DB
 * protein window (60 %), random DNA (25 %), a codon repeat (low complexity after translation, 10 %) or a homopolymer
 * (5 %); then 1 % substitutions, rare N, lower-case, and reverse-complement with probability 1/2. */
static int gen_long_item(const kjgen_db *db, uint64_t seed, uint64_t idx, int minlen, int maxlen, char *out) {
    rng_t r = rng_make(seed, 0x6000000000ull + idx);
    static const char NUC[4] = {'A','C','G','T'};
    int target = minlen + (int)rng_below(&r, (uint32_t)(maxlen - minlen + 1)), n = 0;
    while (n < target) {
        double u = rng_unif(&r); int room = target - n;
        if (u < 0.60) {
            int64_t p = (int64_t)(rng_unif(&r) * db->nprot); if (p >= db->nprot) p = db->nprot - 1;
            int64_t L = db->off[p+1] - db->off[p];
            int naa = 20 + (int)rng_below(&r, 400); if (naa > L) naa = (int)L;
            int64_t s0 = (int64_t)(rng_unif(&r) * (L - naa + 1));
            int shift = (int)rng_below(&r, 3);
            for (int k = 0; k < shift && room > 0; k++, room--) out[n++] = NUC[rng_below(&r, 4)];
            const char *prot = db->seq + db->off[p] + s0;
            for (int k = 0; k < naa && room >= 3; k++, room -= 3) {
                const char *cs = CODONS[prot[k] - 'A']; int nc = (int)strlen(cs) / 3; const char *c = cs + 3 * rng_below(&r, nc);
                out[n++] = c[0]; out[n++] = c[1]; out[n++] = c[2];
            }
        } else if (u < 0.85) {
            int m = 10 + (int)rng_below(&r, 300); if (m > room) m = room;
            for (int k = 0; k < m; k++) out[n++] = NUC[rng_below(&r, 4)];
        } else if (u < 0.95) {
            /* codon repeat of period 3 or 6: translates to a long run of one or two residues in some frames */
            int per = rng_below(&r, 2) ? 3 : 6; char unit[6]; for (int k = 0; k < per; k++) unit[k] = NUC[rng_below(&r, 4)];
            int m = 36 + (int)rng_below(&r, 900); if (m > room) m = room;
            for (int k = 0; k < m; k++) out[n++] = unit[k % per];
        } else {
            char a = NUC[rng_below(&r, 4)]; int m = 20 + (int)rng_below(&r, 200); if (m > room) m = room;
            for (int k = 0; k < m; k++) out[n++] = a;
        }
    }
    for (int k = 0; k < n; k++) if (rng_unif(&r) < 0.01) out[k] = NUC[rng_below(&r, 4)];
    if (rng_below(&r, 2)) revcomp_inplace(out, n);
    if (rng_unif(&r) < 0.3) { int c = 1 + (int)rng_below(&r, 3); for (int k = 0; k < c; k++) out[rng_below(&r, (uint32_t)n)] = 'N'; }
    if (rng_unif(&r) < 0.05) for (int k = 0; k < n; k++) out[k] = (char)(out[k] | 0x20);
    return n;
}
/* seq must hold n*maxlen bytes; off has n+1 entries */
void kjgen_long_reads_packed(const kjgen_db *db, uint64_t seed, uint64_t first, int64_t n, int minlen, int maxlen, char *seq, uint64_t *off) {
    char *tmp = (char *)malloc((size_t)maxlen + 8); uint64_t o = 0;
    for (int64_t i = 0; i < n; i++) { int l = gen_long_item(db, seed, first + (uint64_t)i, minlen, maxlen, tmp); off[i] = o; memcpy(seq + o, tmp, (size_t)l); o += (uint64_t)l; }
    off[n] = o; free(tmp);
}
/* Protein input: DB protein windows with substitutions (70 %) or random residues, low-complexity inserts, and the
 * letters that split a protein read in the reference (X, B, Z, J, U, O -- the reader strips non-letters before the
 * consumer sees the read, kaiju.cpp:331, so only letters can split), some lower-case. */
static int gen_protein_item(const kjgen_db *db, uint64_t seed, uint64_t idx, int minlen, int maxlen, char *out) {
    rng_t r = rng_make(seed, 0x7000000000ull + idx);
    int target = minlen + (int)rng_below(&r, (uint32_t)(maxlen - minlen + 1)), n = 0;
    while (n < target) {
        double u = rng_unif(&r); int room = target - n;
        if (u < 0.70) {
            int64_t p = (int64_t)(rng_unif(&r) * db->nprot); if (p >= db->nprot) p = db->nprot - 1;
            int64_t L = db->off[p+1] - db->off[p];
            int naa = 8 + (int)rng_below(&r, 500); if (naa > L) naa = (int)L; if (naa > room) naa = room;
            int64_t s0 = (int64_t)(rng_unif(&r) * (L - naa + 1));
            memcpy(out + n, db->seq + db->off[p] + s0, (size_t)naa); n += naa;
        } else if (u < 0.85) {
            int m = 5 + (int)rng_below(&r, 80); if (m > room) m = room;
            for (int k = 0; k < m; k++) out[n++] = rand_aa(&r);
        } else {
            char a = rand_aa(&r), b = rand_aa(&r); int m = 12 + (int)rng_below(&r, 300); if (m > room) m = room;
            int two = (int)rng_below(&r, 2);
            for (int k = 0; k < m; k++) out[n++] = (two && (k & 1)) ? b : a;
        }
    }
    for (int k = 0; k < n; k++) if (rng_unif(&r) < 0.03) out[k] = rand_aa(&r);
    static const char BAD[8] = {'X', 'X', 'B', 'Z', 'J', 'X', 'U', 'O'};
    if (rng_unif(&r) < 0.5) { int c = 1 + (int)rng_below(&r, 4); for (int k = 0; k < c; k++) out[rng_below(&r, (uint32_t)n)] = BAD[rng_below(&r, 8)]; }
    if (rng_unif(&r) < 0.1) for (int k = 0; k < n; k++) if (out[k] >= 'A' && out[k] <= 'Z') out[k] = (char)(out[k] | 0x20);
    return n;
}
void kjgen_protein_reads_packed(const kjgen_db *db, uint64_t seed, uint64_t first, int64_t n, int minlen, int maxlen, char *seq, uint64_t *off) {
    char *tmp = (char *)malloc((size_t)maxlen + 8); uint64_t o = 0;
    for (int64_t i = 0; i < n; i++) { int l = gen_protein_item(db, seed, first + (uint64_t)i, minlen, maxlen, tmp); off[i] = o; memcpy(seq + o, tmp, (size_t)l); o += (uint64_t)l; }
    off[n] = o; free(tmp);
}

#ifdef KJGEN_MAIN
static void usage(void) {
    fprintf(stderr, "kjgen db    <nprot> <seed> <out.faa> <out.nodes.dmp>\n"
                    "kjgen reads <nprot> <dbseed> <readseed> <first> <n> <readlen> <paired 0|1> <out1.fq> [out2.fq]\n");
    exit(2);
}
int main(int argc, char **argv) {
    if (argc < 2) usage();
    if (!strcmp(argv[1], "db") && argc == 6) {
        kjgen_db *db = kjgen_db_create(atoll(argv[2]), strtoull(argv[3], 0, 10));
        fprintf(stderr, "kjgen: %lld proteins, %lld letters, %lld taxonomy nodes\n", (long long)db->nprot, (long long)kjgen_db_nletters(db), (long long)db->nnodes);
        return kjgen_db_write(db, argv[4], argv[5]) ? 1 : 0;
    }
    if (!strcmp(argv[1], "reads") && argc >= 10) {
        kjgen_db *db = kjgen_db_create(atoll(argv[2]), strtoull(argv[3], 0, 10));
        int paired = atoi(argv[8]);
        return kjgen_reads_write_fastq(db, strtoull(argv[4], 0, 10), strtoull(argv[5], 0, 10), atoll(argv[6]), atoi(argv[7]), paired,
                                       argv[9], paired ? argv[10] : NULL) ? 1 : 0;
    }
    usage(); return 2;
}
#endif
