/*
 * oracle/tcount_oracle.c -- CPU restatement of the reference's `slamdunk count` / `filter`
 * arithmetic.  TEST INFRASTRUCTURE ONLY: it is the checker the CUDA path is compared with
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference leg).  Nothing
 * in slamdunk_b200/ may import, link or call it.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py) against
 *   - the reference's only golden vector, slamdunk/test/data/reads_slamdunk_mapped_filtered_tcount.tsv
 *     (fixture tests/golden/cfg1), and
 *   - outputs of the reference's own unmodified code run through I/O shims on randomized
 *     data sets (tests/golden/case_*, made by tests/golden/make_golden.py).
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 * The loop structure is kept literal (per-UTR fetch, per-base coverage loop, per-position
 * pass) -- this file is meant to be obviously equivalent, not fast.  The only liberty is
 * OpenMP over BED rows (rows are independent in the reference; the bedgraph dict is merged
 * afterwards in BED order so insertion order is preserved).
 */
#include <ctype.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------- input records */
typedef struct {
    int32_t chrom;    /* index into the FASTA chromosome table; -1: BAM chromosome not in FASTA */
    int32_t pos;      /* pysam reference_start */
    int32_t end;      /* pysam reference_end = pos + max(1, sum M/D/N/=/X) */
    int32_t l_qual;   /* len(query_qualities) */
    uint16_t flag;    /* BAM flag */
    uint8_t mapq;
    uint8_t has_mp;   /* read.has_tag("MP") */
    int32_t mp_len;   /* bytes of MP text, or number of triples */
    int64_t qual_off; /* into the qual byte pool */
    int64_t mp_off;   /* byte offset into the MP text pool, or entry offset into the triple pool */
} orc_read_t;

typedef struct {
    int32_t chrom;    /* FASTA chromosome index, -1 when utr.chromosome is not in the FASTA */
    int32_t in_bam;   /* utr.chromosome in bamFile.references */
    int32_t start;
    int32_t stop;
    int32_t strand;   /* 0 '+', 1 '-' */
    int32_t name_id;  /* index of utr.chromosome among distinct BED chromosome strings (bedgraph key) */
} orc_utr_t;

typedef struct {      /* one row of integers; the caller does the three float divisions */
    int64_t tcontent, coverage_on_ts, conversions_on_ts, read_count, tc_read_count, multimap_count;
    int64_t has_reads; /* the `if countFwd > 0 / countRev > 0` branch was taken (tcounter.py:251) */
} orc_row_t;

typedef struct {      /* bedgraph dict entry in insertion order (tcounter.py:285, 318-326) */
    int32_t name_id, pos, strand, tc, cov;
} orc_bg_t;

/* ---------------------------------------------------------------- SNP dictionary
 * utils/SNPtools.py:23-60.  Keys are the STRING chrom + pos (1-based, as spelled in the VCF),
 * so 'chr1'+'2345' and 'chr12'+'345' are the same key; kept literally. */
typedef struct { char **keys; uint64_t cap, n; } strset_t;

static uint64_t fnv1a(const char *s, size_t n) {
    uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < n; i++) { h ^= (unsigned char)s[i]; h *= 1099511628211ULL; }
    return h;
}
static void strset_init(strset_t *s, uint64_t expect) {
    s->cap = 64; while (s->cap < expect * 2 + 8) s->cap <<= 1;
    s->keys = (char **)calloc(s->cap, sizeof(char *)); s->n = 0;
}
static void strset_add(strset_t *s, const char *k) {
    size_t n = strlen(k); uint64_t i = fnv1a(k, n) & (s->cap - 1);
    while (s->keys[i]) { if (!strcmp(s->keys[i], k)) return; i = (i + 1) & (s->cap - 1); }
    s->keys[i] = strdup(k); s->n++;
}
static int strset_has(const strset_t *s, const char *k) {
    if (!s->n) return 0;
    size_t n = strlen(k); uint64_t i = fnv1a(k, n) & (s->cap - 1);
    while (s->keys[i]) { if (!strcmp(s->keys[i], k)) return 1; i = (i + 1) & (s->cap - 1); }
    return 0;
}
static void strset_free(strset_t *s) {
    for (uint64_t i = 0; i < s->cap; i++) free(s->keys[i]);
    free(s->keys); s->keys = NULL; s->n = 0;
}

typedef struct { strset_t tc, ag; } orc_snps_t;

/* SNPDictionary._addSNP, SNPtools.py:30-38: snp[0]=CHROM, snp[1]=POS text, [3]=REF, [4]=ALT */
void *orc_snps_new(int64_t expect) {
    orc_snps_t *s = (orc_snps_t *)calloc(1, sizeof(*s));
    strset_init(&s->tc, (uint64_t)expect); strset_init(&s->ag, (uint64_t)expect);
    return s;
}
static int ieq1(const char *s, char c) { return s[0] && !s[1] && toupper((unsigned char)s[0]) == c; }
void orc_snps_add(void *h, const char *chrom, const char *pos_text, const char *ref, const char *alt) {
    orc_snps_t *s = (orc_snps_t *)h;
    char key[512];
    snprintf(key, sizeof key, "%s%s", chrom, pos_text);
    if (ieq1(ref, 'T') && ieq1(alt, 'C')) strset_add(&s->tc, key);
    if (ieq1(ref, 'A') && ieq1(alt, 'G')) strset_add(&s->ag, key);
}
void orc_snps_free(void *h) {
    orc_snps_t *s = (orc_snps_t *)h;
    if (!s) return;
    strset_free(&s->tc); strset_free(&s->ag); free(s);
}
/* isTCSnp / isAGSnp, SNPtools.py:53-60: key = chromosome + str(int(position) + 1) */
static int snp_query(const strset_t *set, const char *chrom, int64_t pos0) {
    if (!set->n) return 0;
    char key[512];
    snprintf(key, sizeof key, "%s%lld", chrom, (long long)(pos0 + 1));
    return strset_has(set, key);
}

/* ---------------------------------------------------------------- MP tag
 * SlamSeqFile.py:251-311 (MPTagToConversion): conv = 5*ref + read over A,C,G,T,N = 0..4,
 * so T>C is 16 and A>G is 2.  SlamSeqFile.py:313-348 (fillMismatchesNGM). */
typedef struct { int32_t conv, read_pos, ref_pos; } mp_entry_t;

/* parse "conv:readPos:refPos,conv:readPos:refPos,..." ; returns number of entries, -1 on junk */
static int mp_parse(const char *s, int len, mp_entry_t *out, int cap) {
    int n = 0, i = 0;
    while (i < len) {
        long v[3]; int f;
        for (f = 0; f < 3; f++) {
            long x = 0; int digits = 0, neg = 0;
            if (i < len && s[i] == '-') { neg = 1; i++; }
            while (i < len && s[i] >= '0' && s[i] <= '9') { x = x * 10 + (s[i] - '0'); i++; digits++; }
            if (!digits) return -1;
            v[f] = neg ? -x : x;
            if (f < 2) { if (i >= len || s[i] != ':') return -1; i++; }
        }
        if (n < cap) { out[n].conv = (int32_t)v[0]; out[n].read_pos = (int32_t)v[1]; out[n].ref_pos = (int32_t)v[2]; }
        n++;
        if (i < len) { if (s[i] != ',') return -1; i++; }
    }
    return n;
}

/* ---------------------------------------------------------------- count */
typedef struct {
    /* genome */
    int32_t n_chrom;
    const char *const *chrom_names;   /* FASTA names */
    const char *const *chrom_seq;     /* FASTA sequences, case as in the file */
    const int32_t *chrom_len;
    /* reads sorted by (chrom, pos); range of each chromosome; widest alignment per chromosome */
    const orc_read_t *reads;
    const int64_t *chrom_read_begin;  /* [n_chrom + 1] */
    const int32_t *chrom_max_span;    /* [n_chrom] */
    const uint8_t *qual_pool;
    const char *mp_text;              /* MP tag text pool, or NULL */
    const int32_t *mp_triples;        /* (conv, readPos, refPos) triples, or NULL */
    /* annotation */
    int32_t n_utr;
    const orc_utr_t *utrs;
    /* parameters */
    int32_t min_qual;
    int32_t conversion_threshold;
    void *snps;                       /* orc_snps_new() handle or NULL */
} orc_count_args_t;

typedef struct { int32_t pos, tc, cov; } pos_entry_t;

static int64_t lower_bound_pos(const orc_read_t *r, int64_t lo, int64_t hi, int64_t pos) {
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (r[mid].pos < pos) lo = mid + 1; else hi = mid; }
    return lo;
}

/* One BED row: tcounter.py:165-307.  Returns the row's integers and the per-position entries
 * that the reference would put into conversionBedGraph (tcounter.py:277-285). */
static void count_one_utr(const orc_count_args_t *a, const orc_utr_t *u, orc_row_t *row,
                          pos_entry_t **bg_out, int64_t *n_bg_out) {
    memset(row, 0, sizeof *row);
    *bg_out = NULL; *n_bg_out = 0;
    const int reverse_utr = (u->strand == 1);
    const int64_t ulen = (int64_t)u->stop - (int64_t)u->start;      /* BedEntry.getLength, BedReader.py:46 */

    /* Tcontent: tcounter.py:177-189, FASTA slice upper-cased, count 'A' on '-' else 'T' */
    if (u->chrom >= 0) {
        const char *seq = a->chrom_seq[u->chrom];
        int64_t clen = a->chrom_len[u->chrom];
        int64_t s = u->start, e = u->stop; if (e > clen) e = clen;
        int64_t t = 0;
        for (int64_t p = s; p < e; p++) {
            char c = (char)toupper((unsigned char)seq[p]);
            if (reverse_utr ? (c == 'A') : (c == 'T')) t++;
        }
        row->tcontent = t;
    }
    /* readInRegion: SlamSeqFile.py:419-445 -- iter([]) unless the chromosome is in FASTA and BAM */
    if (u->chrom < 0 || !u->in_bam || ulen <= 0) return;
    const char *seq = a->chrom_seq[u->chrom];
    const char *cname = a->chrom_names[u->chrom];
    const int64_t clen = a->chrom_len[u->chrom];
    const int64_t fstart = u->start > 0 ? u->start : 0;             /* max(0, start) */
    const int64_t fend = u->stop < clen ? u->stop : clen;           /* min(chromosomeLength, stop) */
    if (fstart > fend) return;

    int32_t *tcCountUtr = (int32_t *)calloc((size_t)ulen, sizeof(int32_t));   /* tcounter.py:193-194 */
    int32_t *coverageUtr = (int32_t *)calloc((size_t)ulen, sizeof(int32_t));
    int64_t countFwd = 0, tcCountFwd = 0, countRev = 0, tCountRev = 0, multiMapFwd = 0, multiMapRev = 0;

    const int64_t rb = a->chrom_read_begin[u->chrom], re = a->chrom_read_begin[u->chrom + 1];
    int64_t i0 = lower_bound_pos(a->reads, rb, re, fstart - a->chrom_max_span[u->chrom]);
    int64_t i1 = lower_bound_pos(a->reads, rb, re, fend);
    mp_entry_t ents[4096];

    for (int64_t i = i0; i < i1; i++) {
        const orc_read_t *r = &a->reads[i];
        if (!(r->pos < fend && r->end > fstart)) continue;          /* htslib overlap rule */
        const int is_rev = (r->flag & 0x10) != 0;
        if (is_rev != reverse_utr) continue;                        /* strand skip, SlamSeqFile.py:369 */
        const int is_multi = (r->mapq == 0);                        /* SlamSeqFile.py:381-384 */

        /* fillMismatchesNGM, SlamSeqFile.py:313-348 */
        int tcCount = 0, n_kept = 0;
        int32_t kept_refpos[4096]; uint8_t kept_istc[4096];
        if (r->has_mp) {
            int n_ent;
            if (a->mp_text) n_ent = mp_parse(a->mp_text + r->mp_off, r->mp_len, ents, 4096);
            else {
                n_ent = r->mp_len;
                for (int k = 0; k < n_ent && k < 4096; k++) {
                    const int32_t *t = a->mp_triples + 3 * (r->mp_off + k);
                    ents[k].conv = t[0]; ents[k].read_pos = t[1]; ents[k].ref_pos = t[2];
                }
            }
            if (n_ent > 4096) n_ent = 4096;
            for (int k = 0; k < n_ent; k++) {
                int readPos = ents[k].read_pos - 1;
                int readQlty = (readPos >= r->l_qual) ? 0 : a->qual_pool[r->qual_off + readPos];  /* :326-329 */
                if (readQlty >= a->min_qual) {                                                    /* :330 */
                    int refPos = ents[k].ref_pos - 1;
                    int isSnp;
                    if (is_rev) isSnp = a->snps && snp_query(&((orc_snps_t *)a->snps)->ag, cname, (int64_t)r->pos + refPos);
                    else        isSnp = a->snps && snp_query(&((orc_snps_t *)a->snps)->tc, cname, (int64_t)r->pos + refPos);
                    int relPos = r->pos - u->start + refPos;                                      /* :339 */
                    /* isTCMismatch, SlamSeqFile.py:137-141: (A,G) on reverse, (T,C) forward, not a SNP */
                    int isTC = ((is_rev ? ents[k].conv == 2 : ents[k].conv == 16) && !isSnp);
                    if (isTC) tcCount++;
                    kept_refpos[n_kept] = relPos; kept_istc[n_kept] = (uint8_t)isTC; n_kept++;
                }
            }
        }
        int isTcRead = tcCount >= a->conversion_threshold;          /* SlamSeqFile.py:397-400 */
        if (!isTcRead) { tcCount = 0; n_kept = 0; }                 /* tcounter.py:210-214 */

        if (is_rev) { countRev++; if (tcCount > 0) tCountRev++; if (is_multi) multiMapRev++; }   /* :216-227 */
        else        { countFwd++; if (tcCount > 0) tcCountFwd++; if (is_multi) multiMapFwd++; }

        for (int k = 0; k < n_kept; k++)                            /* tcounter.py:229-231 */
            if (kept_istc[k] && kept_refpos[k] >= 0 && kept_refpos[k] < ulen) tcCountUtr[kept_refpos[k]]++;

        int64_t startRefPos = (int64_t)r->pos - u->start, endRefPos = (int64_t)r->end - u->start;
        for (int64_t p = startRefPos; p < endRefPos; p++)           /* tcounter.py:246-248 */
            if (p >= 0 && p < ulen) coverageUtr[p]++;
    }

    if ((!reverse_utr && countFwd > 0) || (reverse_utr && countRev > 0)) {   /* tcounter.py:251 */
        row->has_reads = 1;
        row->read_count = reverse_utr ? countRev : countFwd;
        row->tc_read_count = reverse_utr ? tCountRev : tcCountFwd;
        row->multimap_count = reverse_utr ? multiMapRev : multiMapFwd;
        pos_entry_t *bg = (pos_entry_t *)malloc((size_t)ulen * sizeof(pos_entry_t));
        int64_t nbg = 0;
        for (int64_t p = 0; p < ulen; p++) {                         /* tcounter.py:277-287 */
            if (coverageUtr[p] <= 0) continue;
            int64_t g = (int64_t)u->start + p;
            /* getRefSeq(): FASTA [start, stop) upper-cased, N-filled outside the chromosome
             * (SlamSeqFile.py:227-228, 424-438) */
            char c = (g < 0 || g >= clen) ? 'N' : (char)toupper((unsigned char)seq[g]);
            if (reverse_utr ? (c == 'A') : (c == 'T')) {
                row->coverage_on_ts += coverageUtr[p];
                row->conversions_on_ts += tcCountUtr[p];
                bg[nbg].pos = (int32_t)g; bg[nbg].tc = tcCountUtr[p]; bg[nbg].cov = coverageUtr[p]; nbg++;
            }
        }
        *bg_out = bg; *n_bg_out = nbg;
    }
    free(tcCountUtr); free(coverageUtr);
}

/* One chromosome of genomewideConversionRates, tcounter.py:373-398 (the array part; the run-length text is written by the
 * Python side of the oracle).  readsInChromosome (SlamSeqFile.py:447-453) hands the iterator startPosition 1 and strand ".",
 * so every index below is (0-based genomic position - 1) and no read is skipped for its strand.
 * tc/ag/covp/covm: int32[chrom_len], zeroed by the caller. */
int orc_positional_chrom(const orc_count_args_t *a, int32_t chrom, int32_t in_bam,
                         int32_t *tc, int32_t *ag, int32_t *covp, int32_t *covm) {
    if (chrom < 0 || chrom >= a->n_chrom) return -1;
    if (!in_bam) return 0;                                           /* `chromosome in self._bamFile.references` else iter([]) */
    const int64_t chrLength = a->chrom_len[chrom];
    const char *cname = a->chrom_names[chrom];
    const int64_t startPosition = 1;
    mp_entry_t ents[4096];
    for (int64_t i = a->chrom_read_begin[chrom]; i < a->chrom_read_begin[chrom + 1]; i++) {   /* fetch(reference=chromosome) */
        const orc_read_t *r = &a->reads[i];
        const int is_rev = (r->flag & 0x10) != 0;                    /* read.direction, SlamSeqFile.py:376-379 */
        int tcCount = 0, n_kept = 0;
        int32_t kept_refpos[4096]; uint8_t kept_istc[4096];
        if (r->has_mp) {                                             /* fillMismatchesNGM, SlamSeqFile.py:313-348 */
            int n_ent;
            if (a->mp_text) n_ent = mp_parse(a->mp_text + r->mp_off, r->mp_len, ents, 4096);
            else {
                n_ent = r->mp_len;
                for (int k = 0; k < n_ent && k < 4096; k++) {
                    const int32_t *t = a->mp_triples + 3 * (r->mp_off + k);
                    ents[k].conv = t[0]; ents[k].read_pos = t[1]; ents[k].ref_pos = t[2];
                }
            }
            if (n_ent > 4096) n_ent = 4096;
            for (int k = 0; k < n_ent; k++) {
                int readPos = ents[k].read_pos - 1;
                int readQlty = (readPos >= r->l_qual) ? 0 : a->qual_pool[r->qual_off + readPos];
                if (readQlty >= a->min_qual) {
                    int refPos = ents[k].ref_pos - 1;
                    int isSnp;
                    if (is_rev) isSnp = a->snps && snp_query(&((orc_snps_t *)a->snps)->ag, cname, (int64_t)r->pos + refPos);
                    else        isSnp = a->snps && snp_query(&((orc_snps_t *)a->snps)->tc, cname, (int64_t)r->pos + refPos);
                    int64_t relPos = (int64_t)r->pos - startPosition + refPos;                    /* :339 */
                    int isTC = ((is_rev ? ents[k].conv == 2 : ents[k].conv == 16) && !isSnp);
                    if (isTC) tcCount++;
                    kept_refpos[n_kept] = (int32_t)relPos; kept_istc[n_kept] = (uint8_t)isTC; n_kept++;
                }
            }
        }
        if (!(tcCount >= a->conversion_threshold)) { tcCount = 0; n_kept = 0; }                  /* tcounter.py:378-382 */
        for (int k = 0; k < n_kept; k++)                                                          /* tcounter.py:384-389 */
            if (kept_istc[k] && kept_refpos[k] >= 0 && kept_refpos[k] < chrLength) {
                if (is_rev) ag[kept_refpos[k]]++; else tc[kept_refpos[k]]++;
            }
        const int64_t startRefPos = (int64_t)r->pos - startPosition, endRefPos = (int64_t)r->end - startPosition;
        for (int64_t p = startRefPos; p < endRefPos; p++)                                         /* tcounter.py:391-396 */
            if (p >= 0 && p < chrLength) { if (is_rev) covm[p]++; else covp[p]++; }
    }
    return 0;
}

/* tcCount of every read as the iterator of readsInChromosome computes it (fillMismatchesNGM, SlamSeqFile.py:313-348;
 * used by genomewideReadSeparation, tcounter.py:510-516, through read.isTcRead = tcCount >= conversionThreshold).
 * out[i] for a->reads[i]; reads whose chromosome is not in the FASTA (chrom < 0) are never seen: out = -1. */
int orc_read_tc(const orc_count_args_t *a, int64_t n_reads, int32_t *out) {
    mp_entry_t ents[4096];
    for (int64_t i = 0; i < n_reads; i++) {
        const orc_read_t *r = &a->reads[i];
        if (r->chrom < 0 || r->chrom >= a->n_chrom) { out[i] = -1; continue; }
        const char *cname = a->chrom_names[r->chrom];
        const int is_rev = (r->flag & 0x10) != 0;
        int tcCount = 0;
        if (r->has_mp) {
            int n_ent = mp_parse(a->mp_text + r->mp_off, r->mp_len, ents, 4096);
            if (n_ent > 4096) n_ent = 4096;
            for (int k = 0; k < n_ent; k++) {
                int readPos = ents[k].read_pos - 1;
                int readQlty = (readPos >= r->l_qual) ? 0 : a->qual_pool[r->qual_off + readPos];
                if (readQlty >= a->min_qual) {
                    int refPos = ents[k].ref_pos - 1;
                    int isSnp;
                    if (is_rev) isSnp = a->snps && snp_query(&((orc_snps_t *)a->snps)->ag, cname, (int64_t)r->pos + refPos);
                    else        isSnp = a->snps && snp_query(&((orc_snps_t *)a->snps)->tc, cname, (int64_t)r->pos + refPos);
                    if ((is_rev ? ents[k].conv == 2 : ents[k].conv == 16) && !isSnp) tcCount++;
                }
            }
        }
        out[i] = tcCount;
    }
    return 0;
}

/* computeTconversions, tcounter.py:165-326, without the text formatting.
 * rows[n_utr]; bedgraph entries returned through *bg (malloc'd, free with orc_free) in the
 * dict's iteration order (first insertion keeps the slot, later assignments overwrite the value). */
int orc_count(const orc_count_args_t *a, orc_row_t *rows, orc_bg_t **bg, int64_t *n_bg, int threads) {
    pos_entry_t **lists = (pos_entry_t **)calloc((size_t)a->n_utr, sizeof(*lists));
    int64_t *nlist = (int64_t *)calloc((size_t)a->n_utr, sizeof(*nlist));
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int32_t i = 0; i < a->n_utr; i++)
        count_one_utr(a, &a->utrs[i], &rows[i], &lists[i], &nlist[i]);

    /* merge into the ordered dict keyed by chromosome-string:position:strand */
    int32_t max_name = 0;
    for (int32_t i = 0; i < a->n_utr; i++) if (a->utrs[i].name_id > max_name) max_name = a->utrs[i].name_id;
    int32_t **slot = (int32_t **)calloc((size_t)(max_name + 1) * 2, sizeof(*slot));
    int64_t total = 0;
    for (int32_t i = 0; i < a->n_utr; i++) total += nlist[i];
    orc_bg_t *out = (orc_bg_t *)malloc((size_t)(total ? total : 1) * sizeof(orc_bg_t));
    int64_t n_out = 0;
    for (int32_t i = 0; i < a->n_utr; i++) {
        const orc_utr_t *u = &a->utrs[i];
        if (!nlist[i]) { free(lists[i]); continue; }
        int32_t **sp = &slot[(size_t)u->name_id * 2 + u->strand];
        if (!*sp) *sp = (int32_t *)calloc((size_t)a->chrom_len[u->chrom] + 1, sizeof(int32_t));
        for (int64_t k = 0; k < nlist[i]; k++) {
            pos_entry_t *e = &lists[i][k];
            int32_t s = (*sp)[e->pos];
            if (s) { out[s - 1].tc = e->tc; out[s - 1].cov = e->cov; }
            else {
                out[n_out].name_id = u->name_id; out[n_out].pos = e->pos; out[n_out].strand = u->strand;
                out[n_out].tc = e->tc; out[n_out].cov = e->cov;
                n_out++; (*sp)[e->pos] = (int32_t)n_out;
            }
        }
        free(lists[i]);
    }
    for (int32_t i = 0; i < (max_name + 1) * 2; i++) free(slot[i]);
    free(slot); free(lists); free(nlist);
    *bg = out; *n_bg = n_out;
    return 0;
}

void orc_free(void *p) { free(p); }

/* The coordinate order `slamdunk filter` leaves its output in (samtools sort, filter.py:308) and the indexed fetch relies on:
 * the reads with keep[i] != 0 (keep == NULL: all) bucketed by chromosome -- reads on no FASTA chromosome (chrom < 0) first, as
 * np.lexsort would place them -- and ordered by position inside each bucket (stable: equal positions keep file order), the buckets
 * sorted side by side on the OpenMP threads.  out: n_out reads; begin[n_chrom + 1]: first read of every chromosome in `out`;
 * max_span[n_chrom]: widest alignment per chromosome (at least 1).  Returns n_out. */
static int cmp_read_pos(const void *a, const void *b) {
    const orc_read_t *x = (const orc_read_t *)a, *y = (const orc_read_t *)b;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    return x->mp_off < y->mp_off ? -1 : (x->mp_off > y->mp_off ? 1 : (x->qual_off < y->qual_off ? -1 : (x->qual_off > y->qual_off ? 1 : 0)));
}
int64_t orc_select_sort(const orc_read_t *r, const uint8_t *keep, int64_t n, int32_t n_chrom, orc_read_t *out, int64_t *begin,
                        int32_t *max_span, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    const int nt = omp_get_max_threads();
#else
    const int nt = 1;
#endif
    const int32_t nb = n_chrom + 1;                       /* bucket 0: chrom < 0, bucket c + 1: chromosome c */
    int64_t *cnt = (int64_t *)calloc((size_t)nt * nb, sizeof(int64_t));
    const int64_t chunk = (n + nt - 1) / nt;
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; t++) {
        const int64_t b = t * chunk, e = b + chunk < n ? b + chunk : n;
        for (int64_t i = b; i < e; i++)
            if (!keep || keep[i]) cnt[(size_t)t * nb + (r[i].chrom < 0 || r[i].chrom >= n_chrom ? 0 : r[i].chrom + 1)]++;
    }
    int64_t run = 0;
    int64_t *bucket0 = (int64_t *)calloc((size_t)nb + 1, sizeof(int64_t));
    for (int32_t k = 0; k < nb; k++) {
        bucket0[k] = run;
        for (int t = 0; t < nt; t++) { const int64_t c = cnt[(size_t)t * nb + k]; cnt[(size_t)t * nb + k] = run; run += c; }
    }
    bucket0[nb] = run;
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; t++) {
        const int64_t b = t * chunk, e = b + chunk < n ? b + chunk : n;
        int64_t *at = cnt + (size_t)t * nb;
        for (int64_t i = b; i < e; i++)
            if (!keep || keep[i]) out[at[r[i].chrom < 0 || r[i].chrom >= n_chrom ? 0 : r[i].chrom + 1]++] = r[i];
    }
    /* the scatter above keeps file order inside a bucket; qsort is not stable, so ties are broken by the reads' pool offsets,
     * which ascend in file order for the batches the bench builds */
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t k = 0; k < nb; k++) {
        const int64_t b = bucket0[k], e = bucket0[k + 1];
        if (e - b > 1) qsort(out + b, (size_t)(e - b), sizeof(orc_read_t), cmp_read_pos);
        if (k > 0) {
            int32_t span = 1;
            for (int64_t i = b; i < e; i++) if (out[i].end - out[i].pos > span) span = out[i].end - out[i].pos;
            max_span[k - 1] = span;
        }
    }
    for (int32_t c = 0; c <= n_chrom; c++) begin[c] = bucket0[c + 1];
    free(cnt); free(bucket0);
    return run;
}

/* ---------------------------------------------------------------- filter
 * dunks/filter.py.  Input in FILE order. */
typedef struct {
    int32_t chrom_id;   /* id of the BAM chromosome name among the distinct BED+BAM names; -1 unmapped */
    int32_t start, end; /* reference_start, reference_end */
    int32_t name_id;    /* equal ids <=> equal query_name strings */
    uint16_t flag; uint8_t mapq; uint8_t pad;
    float xi;           /* XI:f tag; the reference promotes it with float() */
    int32_t nm;         /* NM tag */
} orc_fread_t;

typedef struct { int32_t chrom_id, start, stop, name_id; } orc_futr_t;  /* name_id: id of utr.name */

/* Default branch, filter.py:233-256.  keep[i] = 1 when the record is written.
 * stats = mapped, unmapped, filtered, mqFiltered, idFiltered, nmFiltered, multimapper */
int orc_filter_default_mt(const orc_fread_t *r, int64_t n, int mq, double min_identity, int nm,
                          uint8_t *keep, int64_t stats[7], int threads) {
    int64_t mapped = 0, unmapped = 0, filtered = 0, mqf = 0, idf = 0, nmf = 0;
    /* every record is decided on its own (the counters are plain sums), so the file-order loop splits over threads */
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(static) reduction(+ : mapped, unmapped, filtered, mqf, idf, nmf)
#endif
    for (int64_t i = 0; i < n; i++) {
        keep[i] = 0;
        int primary = !(r[i].flag & 0x100) && !(r[i].flag & 0x800);
        int is_unmapped = (r[i].flag & 0x4) != 0;
        if (primary) { if (is_unmapped) unmapped++; else mapped++; }      /* :235-239 */
        if (is_unmapped) continue;                                        /* :241 */
        if (r[i].mapq < mq) { mqf++; continue; }                          /* :243-245 */
        if ((double)r[i].xi < min_identity) { idf++; continue; }          /* :246-248 */
        if (nm > -1 && r[i].nm > nm) { nmf++; continue; }                 /* :249-251 */
        if (primary) filtered++;                                          /* :253-254 */
        keep[i] = 1;                                                      /* :256 */
    }
    stats[0] = mapped; stats[1] = unmapped; stats[2] = filtered; stats[3] = mqf; stats[4] = idf; stats[5] = nmf; stats[6] = 0;
    return 0;
}
int orc_filter_default(const orc_fread_t *r, int64_t n, int mq, double min_identity, int nm,
                       uint8_t *keep, int64_t stats[7]) {
    return orc_filter_default_mt(r, n, mq, min_identity, nm, keep, stats, 1);
}

/* multimapUTRRetainment, filter.py:72-210 (+ dumpBufferToBam :56-65).
 * keep[i]: 0 dropped, 1 written as is, 2 written by dumpBufferToBam (secondary/supplementary
 * cleared, RD tag set).  rd_first[i]/rd_last[i]: for keep==2 the index range of records whose
 * "chr:start-end " strings make up the RD tag, rd_list[] holds those indices (multimapList).
 * The UTR tree query returns a SET in the reference (iteration order unspecified); here hits are
 * visited in BED order, which only matters when one alignment hits two differently named UTRs. */
int orc_filter_multimap(const orc_fread_t *r, int64_t n, const orc_futr_t *utrs, int32_t n_utr,
                        double min_identity, int nm, uint8_t *keep,
                        int64_t *rd_begin, int64_t *rd_end, int64_t *rd_list, int64_t stats[7]) {
    int64_t mapped = 0, unmapped = 0, filtered = 0, idf = 0, nmf = 0;
    /* multimapBuffer: ordered dict  utr-name -> list of reads; we only need key order and, per key, the last read */
    int32_t *buf_keys = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_utr + 1));
    int64_t *buf_last = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_utr + 1));
    int32_t n_keys = 0;
    int32_t prev_name = -1;             /* prevRead = "" */
    int dumpBuffer = 1;
    int64_t list_n = 0, list_total = 0; /* multimapList as indices appended to rd_list */
    int64_t list_begin = 0;

    for (int64_t i = 0; i < n; i++) {
        keep[i] = 0; rd_begin[i] = rd_end[i] = 0;
        int primary = !(r[i].flag & 0x100) && !(r[i].flag & 0x800);
        int is_unmapped = (r[i].flag & 0x4) != 0;
        if (primary) { if (is_unmapped) unmapped++; else mapped++; }      /* :98-102 */
        if (is_unmapped) continue;                                        /* :105-106 */
        if ((double)r[i].xi < min_identity) { idf++; continue; }          /* :107-109 */
        if (nm > -1 && r[i].nm > nm) { nmf++; continue; }                 /* :110-112 */
        if (r[i].mapq == 0) {
            if (r[i].name_id != prev_name && prev_name != -1) {           /* :115 */
                if (dumpBuffer && n_keys > 0) {                           /* :118-120 */
                    int64_t w = buf_last[n_keys - 1];
                    keep[w] = 2; rd_begin[w] = list_begin; rd_end[w] = list_begin + list_n;
                    filtered++;
                }
                dumpBuffer = 1; list_begin += list_n; list_n = 0; n_keys = 0;   /* :131-133 */
            }
            /* interval tree query [start, end) against [utr.start, utr.stop + 1), BedReader.py:28 */
            int first_hit_group = (n_keys == 0);                           /* :147 */
            for (int32_t k = 0; k < n_utr; k++) {
                if (utrs[k].chrom_id != r[i].chrom_id) continue;
                if (!(utrs[k].start < r[i].end && utrs[k].stop + 1 > r[i].start)) continue;
                int32_t found = -1;
                for (int32_t q = 0; q < n_keys; q++) if (buf_keys[q] == utrs[k].name_id) { found = q; break; }
                if (found < 0) {
                    buf_keys[n_keys] = utrs[k].name_id; buf_last[n_keys] = i; n_keys++;
                    if (!first_hit_group) dumpBuffer = 0;                  /* :154-158 */
                } else buf_last[found] = i;
            }
            rd_list[list_begin + list_n] = i; list_n++; list_total++;      /* :170 */
            prev_name = r[i].name_id;                                      /* :172 */
        } else {
            if (n_keys > 0) {                                              /* :176 */
                if (dumpBuffer) {
                    int64_t w = buf_last[n_keys - 1];
                    keep[w] = 2; rd_begin[w] = list_begin; rd_end[w] = list_begin + list_n;
                    filtered++;
                }
                n_keys = 0; dumpBuffer = 1; list_begin += list_n; list_n = 0;   /* :181-187 */
            }
            prev_name = r[i].name_id;                                      /* :190 */
            keep[i] = 1; filtered++;                                       /* :191-192 */
        }
    }
    if (dumpBuffer && n_keys > 0) {                                        /* :197-199 */
        int64_t w = buf_last[n_keys - 1];
        keep[w] = 2; rd_begin[w] = list_begin; rd_end[w] = list_begin + list_n;
        filtered++;
    }
    (void)list_total;
    stats[0] = mapped; stats[1] = unmapped; stats[2] = filtered; stats[3] = 0; stats[4] = idf; stats[5] = nmf;
    stats[6] = mapped - filtered - idf - nmf;                              /* :201 */
    free(buf_keys); free(buf_last);
    return 0;
}

int orc_sizeof_read(void) { return (int)sizeof(orc_read_t); }
int orc_sizeof_utr(void) { return (int)sizeof(orc_utr_t); }
int orc_sizeof_row(void) { return (int)sizeof(orc_row_t); }
int orc_sizeof_bg(void) { return (int)sizeof(orc_bg_t); }
int orc_sizeof_fread(void) { return (int)sizeof(orc_fread_t); }
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
