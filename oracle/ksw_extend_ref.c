/*
 * oracle/ksw_extend_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, int32) of the seed-extension hot path that
 * peterpengwei/bwa-mem-sw implements in RTL.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file's
 * shared object.  The product library (libbsw.so) never links, loads or calls it.
 *
 * PARITY STATUS: "parity unpinned" by the reference itself -- /root/reference is
 * RTL only (no tests, no golden vectors, no simulator in this image, and the
 * software twin bwa-mem-quickassist/bwa-0.7.8 is an un-vendored separate repo).
 * This restatement follows, line by line:
 *   - sw_pe_array_sw_extend.v           (the recurrence; cites below as "sx:N")
 *   - sw_pe_array_proc_element.v        (left/right sequencing + clip rule; "pe:N")
 *   - the published ksw_extend2 algorithm of BWA 0.7.x for the two pieces the RTL
 *     does not contain: the z-drop rule and the max_ins/max_del band clamp
 *     (the RTL has no zdrop port, sx:96-116, and takes max_ins/max_del from the
 *     host, pe:924-934).
 * It is pinned by hand-derivable known-answer tests (tests/test_oracle_kat.py) and
 * an independently structured full-matrix model (oracle/matrix_model.py).
 *
 * Two recurrence policies:
 *   variant 1 (V1) = what the RTL computes = BWA-0.7.8-era ksw_extend2
 *   variant 2 (V2) = current upstream BWA ksw_extend2
 * They differ at the four places marked [V1/V2] below.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef struct { int32_t h, e; } eh_t;

typedef struct {
    int8_t  mat[25];                 /* 5x5, row = target base, col = query base (sx:1915-1940: sel = 5*t+q) */
    int32_t o_del, e_del, o_ins, e_ins, zdrop, end_bonus;
} bswref_params;

typedef struct { int32_t score, qle, tle, gtle, gscore, max_off; } bswref_result;

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* ksw_extend2's band clamp (not in the RTL: max_ins/max_del arrive precomputed, pe:924-934;
 * applied at sx:1763-1765,1881,1890). */
int bswref_clamp_w(const bswref_params *p, int qlen, int w, int end_bonus)
{
    int i, max = 0, max_ins, max_del;
    for (i = 0; i < 25; ++i) max = max > p->mat[i] ? max : p->mat[i];
    max_ins = (int)((double)(qlen * max + end_bonus - p->o_ins) / p->e_ins + 1.);
    max_ins = max_ins > 1 ? max_ins : 1;
    w = w < max_ins ? w : max_ins;
    max_del = (int)((double)(qlen * max + end_bonus - p->o_del) / p->e_del + 1.);
    max_del = max_del > 1 ? max_del : 1;
    w = w < max_del ? w : max_del;
    return w;
}

/*
 * One banded affine-gap extension (one band try).  Preconditions: qlen>=1, tlen>=1, h0>0
 * (the PE never calls sw_extend with qlen==0: pe:1670).
 * `scratch` must hold (qlen+1) eh_t.  Returns score; fills *out; adds executed cells to *cells.
 */
/* `carry` (normally NULL): the running maxima of a previous band try.  ksw_extend2 re-initialises them on every call;
 * the RTL initialises max / max_i / max_j / max_ie / gscore / max_off once per sw_extend invocation, i.e. BEFORE the
 * band-try loop (sx:885-890,913-930,957-959,1003-1031), so its second try starts from the first try's values
 * (SURVEY appendix C row 5).  Only bswref_sw_extend_rtl() passes a carry; it exists to compare against the translated
 * RTL (oracle/_ref) on tasks where the second try runs. */
typedef struct { int32_t valid, max, max_i, max_j, max_ie, gscore, max_off; } bswref_carry;

static int bswref_extend_core_x(const bswref_params *p, int variant, int qlen, const uint8_t *query,
                       int tlen, const uint8_t *target, int w, int h0,
                       bswref_result *out, int64_t *cells, eh_t *eh, bswref_carry *carry, int clamp_w);

int bswref_extend_core(const bswref_params *p, int variant, int qlen, const uint8_t *query,
                       int tlen, const uint8_t *target, int w, int h0,
                       bswref_result *out, int64_t *cells, eh_t *eh)
{
    return bswref_extend_core_x(p, variant, qlen, query, tlen, target, w, h0, out, cells, eh, NULL, 1);
}

static int bswref_extend_core_x(const bswref_params *p, int variant, int qlen, const uint8_t *query,
                       int tlen, const uint8_t *target, int w, int h0,
                       bswref_result *out, int64_t *cells, eh_t *eh, bswref_carry *carry, int clamp_w)
{
    const int o_del = p->o_del, e_del = p->e_del, o_ins = p->o_ins, e_ins = p->e_ins;
    const int oe_del = o_del + e_del;            /* sx:1860 */
    const int oe_ins = o_ins + e_ins;            /* sx:1823 */
    const int zdrop = p->zdrop;
    int i, j, beg, end, max, max_i, max_j, max_ie, gscore, max_off;
    int64_t ncell = 0;

    if (clamp_w) w = bswref_clamp_w(p, qlen, w, p->end_bonus);

    /* first row: eh[j].h = H(-1, j-1), all e = 0  (sx:1818; 1979,1957,1974; 1975-1978,1821) */
    memset(eh, 0, sizeof(eh_t) * (size_t)(qlen + 1));
    eh[0].h = h0;
    eh[1].h = h0 > oe_ins ? h0 - oe_ins : 0;
    for (j = 2; j <= qlen && eh[j - 1].h > e_ins; ++j) eh[j].h = eh[j - 1].h - e_ins;

    max = h0; max_i = max_j = -1; max_ie = -1; gscore = -1; max_off = 0;   /* sx:889,1009,919,1019,1029,929 */
    if (carry && carry->valid) {                                            /* RTL second band try: see bswref_carry */
        max = carry->max; max_i = carry->max_i; max_j = carry->max_j;
        max_ie = carry->max_ie; gscore = carry->gscore; max_off = carry->max_off;
    }
    beg = 0; end = qlen;                                                    /* sx:769,779 */
    for (i = 0; i < tlen; ++i) {                                            /* sx:1891 */
        int f = 0, h1, m = 0, mj = -1;                                      /* sx:789,879,909 */
        const int8_t *srow = &p->mat[5 * target[i]];
        if (beg < i - w) beg = i - w;                                       /* sx:1846,1894,1895,1803 */
        if (end > i + w + 1) end = i + w + 1;                               /* sx:1980,1843,1897 */
        if (end > qlen) end = qlen;                                         /* sx:1898,1842 */
        /* first column */
        if (variant == 1 || beg == 0) {                                     /* [V1/V2] V1 unconditional: sx:1796,1795,1880,1835,849 */
            h1 = h0 - (o_del + e_del * (i + 1));
            if (h1 < 0) h1 = 0;
        } else h1 = 0;
        for (j = beg; j < end; ++j) {                                       /* sx:1901 */
            eh_t *c = &eh[j];
            int h, t, M = c->h, e = c->e;                                   /* sx:1799,1772-1773 */
            c->h = h1;                                                      /* sx:1776 */
            if (variant == 1) M = M + srow[query[j]];                       /* [V1/V2] sx:1836,1871,1956,1797 */
            else              M = M ? M + srow[query[j]] : 0;
            h = M > e ? M : e;                                              /* sx:1943,1798 */
            h = h > f ? h : f;                                              /* sx:1944,1809 */
            h1 = h;                                                         /* sx:847 */
            mj = m > h ? mj : j;                                            /* sx:1945,1816  (ties move mj right) */
            m = m > h ? m : h;                                              /* sx:1808 */
            t = (variant == 1 ? h : M) - oe_del;                            /* [V1/V2] gap open from h (V1: sx:1866) or M (V2) */
            t = t > 0 ? t : 0;                                              /* sx:1981,1862 */
            e -= e_del;                                                     /* sx:1770 */
            e = e > t ? e : t;                                              /* sx:1966,1771 */
            c->e = e;                                                       /* sx:1776 */
            t = (variant == 1 ? h : M) - oe_ins;                            /* sx:1863 */
            t = t > 0 ? t : 0;                                              /* sx:1967,1865 */
            f -= e_ins;                                                     /* sx:1780 */
            f = f > t ? f : t;                                              /* sx:1968,1781 */
        }
        if (end > beg) ncell += end - beg;
#ifdef BSWREF_ROWHIST
        { extern long long bswref_rowhist[64]; int wdt = end > beg ? end - beg : 0; bswref_rowhist[wdt / 16 < 63 ? wdt / 16 : 63]++; }
#endif
        eh[end].h = h1; eh[end].e = 0;                                      /* sx:1775,1904,1494-1495,1538 */
        if (j == qlen) {                                                    /* sx:1768,1913 */
            max_ie = gscore > h1 ? max_ie : i;                              /* sx:1941,1829 */
            gscore = gscore > h1 ? gscore : h1;                             /* sx:1831 */
        }
        if (m == 0) break;                                                  /* sx:1942,1686-1687 */
        if (m > max) {                                                      /* sx:1959 */
            max = m; max_i = i; max_j = mj;                                 /* sx:1810,1833,1801 */
            max_off = max_off > abs(mj - i) ? max_off : abs(mj - i);        /* sx:1845,1708,1707,1964,1812 */
        } else if (zdrop > 0) {                                             /* NOT IN RTL: ksw_extend2's z-drop rule */
            if (i - max_i > mj - max_j) {
                if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break;
            } else {
                if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break;
            }
        }
        /* narrowing for the next row */
        if (variant == 1) {                                                 /* [V1/V2] run of non-zero h around mj */
            for (j = mj; j >= beg && eh[j].h; --j) ;                        /* sx:1767,1766,1838,1965 */
            beg = j + 1;                                                    /* sx:1769 */
            for (j = mj + 2; j <= end && eh[j].h; ++j) ;                    /* sx:1949,1960,1782-1789,1905-1911 */
            end = j;                                                        /* sx:1826,1872,1779 */
        } else {
            for (j = beg; j < end && eh[j].h == 0 && eh[j].e == 0; ++j) ;
            beg = j;
            for (j = end; j >= beg && eh[j].h == 0 && eh[j].e == 0; --j) ;
            end = j + 2 < qlen ? j + 2 : qlen;
        }
    }
    out->score = max;                                                       /* sx:1315-1375 return tuple */
    out->qle = max_j + 1;                                                   /* sx:1841 */
    out->tle = max_i + 1;                                                   /* sx:1868 */
    out->gtle = max_ie + 1;                                                 /* sx:1794 */
    out->gscore = gscore;                                                   /* sx:1792 */
    out->max_off = max_off;                                                 /* sx:1815 */
    if (cells) *cells += ncell;
    if (carry) {
        carry->valid = 1; carry->max = max; carry->max_i = max_i; carry->max_j = max_j;
        carry->max_ie = max_ie; carry->gscore = gscore; carry->max_off = max_off;
    }
    return max;
}

/* One invocation of the RTL's sw_extend as a whole (sx:1639-1705): band clamp from the host-supplied max_ins / max_del
 * (sx:1763-1765,1881,1890), up to two band tries with `prev` seeded from regScore (sx:1069,1822,1859) and the maxima
 * carried from try to try, no z-drop.  out7 = ap_return_0..6 = score, aw, qle, tle, gtle, gscore, max_off
 * (sx:117-123,1315-1375).  int32 arithmetic: equal to the RTL wherever the RTL's 8-bit datapath does not wrap. */
void bswref_sw_extend_rtl(const bswref_params *p, int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                          int w, int h0, int reg_score, int max_ins, int max_del, int32_t *out7, int64_t *cells)
{
    eh_t *eh = (eh_t *)malloc(sizeof(eh_t) * (size_t)(qlen + 1));
    bswref_params pp = *p;
    bswref_carry carry;
    bswref_result res;
    int k, aw = w, score = reg_score, prev;
    memset(&carry, 0, sizeof carry);
    memset(&res, 0, sizeof res);
    pp.zdrop = 0;
    for (k = 0; k < 2; ++k) {                                               /* sx:1963,1878 */
        int wk;
        prev = score;                                                       /* sx:1822,1859 */
        aw = w << k;                                                        /* sx:1765 */
        wk = imin(aw, imin(max_ins, max_del));                              /* sx:1764,1881 ; 1763,1890 */
        score = bswref_extend_core_x(&pp, 1, qlen, query, tlen, target, wk, h0, &res, cells, eh, &carry, 0);
        if (score == prev || res.max_off < (aw >> 1) + (aw >> 2)) break;    /* sx:1824-1825,1969-1970,1837 */
    }
    out7[0] = score; out7[1] = aw; out7[2] = res.qle; out7[3] = res.tle;
    out7[4] = res.gtle; out7[5] = res.gscore; out7[6] = res.max_off;
    free(eh);
}

int bswref_extend(const bswref_params *p, int variant, int qlen, const uint8_t *query,
                  int tlen, const uint8_t *target, int w, int h0,
                  bswref_result *out, int64_t *cells)
{
    eh_t *eh = (eh_t *)malloc(sizeof(eh_t) * (size_t)(qlen + 1));
    int r = bswref_extend_core(p, variant, qlen, query, tlen, target, w, h0, out, cells, eh);
    free(eh);
    return r;
}

/* ---- tiny pthread work-sharing helper (dynamic chunks, like omp schedule(dynamic,chunk)) ---- */
typedef struct {
    void (*fn)(void *arg, int64_t lo, int64_t hi, int tid);
    void *arg; int64_t n, chunk; atomic_llong next; int tid_seq;
} bswref_pool;
typedef struct { bswref_pool *pool; int tid; } bswref_worker;

static void *bswref_worker_main(void *v)
{
    bswref_worker *w = (bswref_worker *)v;
    bswref_pool *P = w->pool;
    for (;;) {
        int64_t lo = atomic_fetch_add(&P->next, P->chunk), hi;
        if (lo >= P->n) break;
        hi = lo + P->chunk < P->n ? lo + P->chunk : P->n;
        P->fn(P->arg, lo, hi, w->tid);
    }
    return NULL;
}

int bswref_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void bswref_parallel_for(int64_t n, int64_t chunk, int nthreads,
                                void (*fn)(void *, int64_t, int64_t, int), void *arg)
{
    bswref_pool P; pthread_t th[256]; bswref_worker wk[256]; int t;
    if (nthreads <= 0) nthreads = bswref_max_threads();
    if (nthreads > 256) nthreads = 256;
    P.fn = fn; P.arg = arg; P.n = n; P.chunk = chunk; atomic_init(&P.next, 0);
    if (nthreads == 1) { wk[0].pool = &P; wk[0].tid = 0; bswref_worker_main(&wk[0]); return; }
    for (t = 0; t < nthreads; ++t) { wk[t].pool = &P; wk[t].tid = t; pthread_create(&th[t], NULL, bswref_worker_main, &wk[t]); }
    for (t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
}

/* ---- level 1 batch (flat layout: task i's query = qbuf[qoff[i]..qoff[i+1]) etc.) ---- */
typedef struct {
    const bswref_params *p; int variant;
    const uint8_t *qbuf; const int64_t *qoff; const uint8_t *tbuf; const int64_t *toff;
    const int32_t *h0, *w; bswref_result *out; int64_t *cells; int64_t maxq;
} l1_args;

static void l1_range(void *v, int64_t lo, int64_t hi, int tid)
{
    l1_args *a = (l1_args *)v; int64_t i;
    eh_t *eh = (eh_t *)malloc(sizeof(eh_t) * (size_t)(a->maxq + 1));
    (void)tid;
    for (i = lo; i < hi; ++i) {
        int64_t c = 0;
        int ql = (int)(a->qoff[i + 1] - a->qoff[i]), tl = (int)(a->toff[i + 1] - a->toff[i]);
        bswref_extend_core(a->p, a->variant, ql, a->qbuf + a->qoff[i], tl, a->tbuf + a->toff[i],
                           a->w[i], a->h0[i], &a->out[i], &c, eh);
        if (a->cells) a->cells[i] = c;
    }
    free(eh);
}

void bswref_extend_batch(const bswref_params *p, int variant, int64_t n,
                         const uint8_t *qbuf, const int64_t *qoff,
                         const uint8_t *tbuf, const int64_t *toff,
                         const int32_t *h0, const int32_t *w,
                         bswref_result *out, int64_t *cells_per_task, int nthreads)
{
    l1_args a = { p, variant, qbuf, qoff, tbuf, toff, h0, w, out, cells_per_task, 0 };
    int64_t i;
    for (i = 0; i < n; ++i) if (qoff[i + 1] - qoff[i] > a.maxq) a.maxq = qoff[i + 1] - qoff[i];
    bswref_parallel_for(n, 256, nthreads, l1_range, &a);
}

/* ---- level 2: one seed task = what one FPGA PE does (pe:1593-1685) ---- */
typedef struct {
    bswref_params p;
    int32_t w, pen_clip5, pen_clip3;        /* TBB header words 0-1 (pe:815-820, 915-918) */
} bswref_params2;

typedef struct {
    const uint8_t *q_left, *q_right, *t_left, *t_right;   /* left flanks already reversed (pe reads both forward) */
    int32_t qlen[2], tlen[2];                               /* pe:880-892 */
    int32_t init_score, qbeg, h0;                           /* regScore, qBeg_ori, h0: pe:871-874, 826-828 */
    uint32_t id;                                            /* pe:807 */
} bswref_seed_task;

typedef struct { uint32_t id; int32_t qb, qe, rb, re, score, truesc, w; } bswref_aln_record;

#define BSWREF_MAX_BAND_TRY 2

/* rtl_gaps == NULL: BWA's mem_chain2aln loop (every band try is a fresh ksw_extend2 call, clamp from the formula).
 * rtl_gaps != NULL: {max_ins_left, max_del_left, max_ins_right, max_del_right} as the TBB carries them (pe:924-934);
 * each side is then ONE invocation of the RTL's sw_extend (bswref_sw_extend_rtl: maxima carried across the two tries,
 * no z-drop) -- the mode compared against the translated RTL in oracle/_ref. */
static void chain2aln_impl(const bswref_params2 *P, int variant, const bswref_seed_task *s,
                           bswref_aln_record *r, int64_t *cells, const int32_t *rtl_gaps)
{
    /* initial state: pe:471-475,581-583,605-607,649-651,673-675,707-709,717-719,757-759,783-797 */
    int qb = 0, rb = 0, qe = s->qlen[1], re = 0, score = 0;
    int sc0 = s->init_score, truesc = s->init_score;
    int aw[2] = { P->w, P->w };
    int side;
    for (side = 0; side < 2; ++side) {                                      /* pe:1597,1622 */
        const uint8_t *q = side ? s->q_right : s->q_left;
        const uint8_t *t = side ? s->t_right : s->t_left;
        int ql = s->qlen[side], tl = s->tlen[side], k;
        bswref_params pp = P->p;
        bswref_result res;
        int h0, prev, pen_clip = side ? P->pen_clip3 : P->pen_clip5, a_score;
        if (ql == 0) continue;                                              /* pe:1670,430,443-445 */
        h0 = side ? sc0 : s->h0;                                            /* pe:1671,1652 */
        pp.end_bonus = pen_clip;                                            /* BWA passes pen_clip5/3 as ksw_extend2's end_bonus */
        a_score = sc0;
        if (rtl_gaps) {
            int32_t o7[7];
            bswref_sw_extend_rtl(&pp, ql, q, tl, t, P->w, h0, sc0, rtl_gaps[2 * side], rtl_gaps[2 * side + 1], o7, cells);
            a_score = o7[0]; aw[side] = o7[1];
            res.score = o7[0]; res.qle = o7[2]; res.tle = o7[3]; res.gtle = o7[4]; res.gscore = o7[5]; res.max_off = o7[6];
        } else
        for (k = 0; k < BSWREF_MAX_BAND_TRY; ++k) {                         /* sx:1963,1878 */
            prev = a_score;                                                 /* sx:1822,1859 */
            aw[side] = P->w << k;                                           /* sx:1765 */
            a_score = bswref_extend(&pp, variant, ql, q, tl, t, aw[side], h0, &res, cells);
            if (a_score == prev || res.max_off < (aw[side] >> 1) + (aw[side] >> 2)) break;   /* sx:1824-1825,1969-1970,1837 */
        }
        if (res.gscore <= 0 || res.gscore <= a_score - pen_clip) {          /* local: pe:1672,1674-1675,1667 */
            if (side == 0) { qb = s->qbeg - res.qle; rb = -res.tle; truesc = a_score; }        /* pe:591-599,659-667,767-777 */
            else           { qe = res.qle; re = res.tle; truesc += a_score - sc0; }            /* pe:615-623,683-691,1679-1680 */
        } else {                                                            /* to-end */
            if (side == 0) { qb = 0; rb = -res.gtle; truesc = res.gscore; }
            else           { qe = s->qlen[1]; re = res.gtle; truesc += res.gscore - sc0; }
        }
        score = sc0 = a_score;                                              /* pe:697-700,727-728,1594,1685 */
    }
    r->id = s->id;                                                          /* pe:1187-1205,1662-1665 */
    r->qb = qb; r->qe = qe; r->rb = rb; r->re = re;
    r->score = score; r->truesc = truesc;
    r->w = aw[0] > aw[1] ? aw[0] : aw[1];                                   /* pe:1669,1684 */
}

void bswref_chain2aln(const bswref_params2 *P, int variant, const bswref_seed_task *s,
                      bswref_aln_record *r, int64_t *cells)
{
    chain2aln_impl(P, variant, s, r, cells, NULL);
}

/* One FPGA processing element's task exactly as the RTL sequences it (see chain2aln_impl). */
void bswref_chain2aln_rtl(const bswref_params2 *P, const bswref_seed_task *s, const int32_t *gaps4,
                          bswref_aln_record *r, int64_t *cells)
{
    chain2aln_impl(P, 1, s, r, cells, gaps4);
}

typedef struct {
    const bswref_params2 *P; int variant; const bswref_seed_task *tasks; bswref_aln_record *out;
    atomic_llong cells;
} l2_args;

static void l2_range(void *v, int64_t lo, int64_t hi, int tid)
{
    l2_args *a = (l2_args *)v; int64_t i, c = 0;
    (void)tid;
    for (i = lo; i < hi; ++i) bswref_chain2aln(a->P, a->variant, &a->tasks[i], &a->out[i], &c);
    atomic_fetch_add(&a->cells, c);
}

void bswref_chain2aln_batch(const bswref_params2 *P, int variant, int64_t n,
                            const bswref_seed_task *tasks, bswref_aln_record *out,
                            int64_t *cells_total, int nthreads)
{
    l2_args a; a.P = P; a.variant = variant; a.tasks = tasks; a.out = out; atomic_init(&a.cells, 0);
    bswref_parallel_for(n, 64, nthreads, l2_range, &a);
    if (cells_total) *cells_total = (int64_t)atomic_load(&a.cells);
}

/* =====================================================================================================================
 * ksw_global2 -- banded global alignment with affine gaps and traceback, the DP that follows the extension in BWA-MEM
 * (bwa_gen_cigar2 calls it once per reported alignment).  NOT part of /root/reference and not
 * derivable from the RTL: this restates the published BWA algorithm (ksw.c, 0.7.x) from its description -- query in the
 * inner loop, eh[j] = {H(i-1,j-1), E(i,j)}, M separated from H so that the direction is recorded correctly, one byte
 * per cell holding the H direction (bits 0-1) and the "E / F extends" flags for the next cell (bit 2, bit 5).
 * PARITY UNPINNED for this function: no source, no vectors, no RTL -- it is the oracle of SURVEY row f.4 only.
 * cigar: BAM encoding (len << 4 | op), op 0 = M, 1 = I, 2 = D; returns the score, *n_cigar = number of operations
 * (or -1 if more than max_cigar would be needed).
 * ===================================================================================================================== */
#define BSWREF_MINUS_INF (-0x40000000)

static int push_op(uint32_t *cigar, int n, int max_cigar, int op, int len)
{
    if (n > 0 && (cigar[n - 1] & 0xf) == (uint32_t)op) { cigar[n - 1] += (uint32_t)len << 4; return n; }
    if (n >= max_cigar) return -1;
    cigar[n] = (uint32_t)len << 4 | (uint32_t)op;
    return n + 1;
}

int bswref_global(const bswref_params *p, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int w,
                  int max_cigar, uint32_t *cigar, int *n_cigar)
{
    const int o_del = p->o_del, e_del = p->e_del, o_ins = p->o_ins, e_ins = p->e_ins;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    int32_t *eh_h = (int32_t *)malloc(sizeof(int32_t) * (size_t)(qlen + 1));
    int32_t *eh_e = (int32_t *)malloc(sizeof(int32_t) * (size_t)(qlen + 1));
    uint8_t *z = (uint8_t *)malloc((size_t)n_col * (size_t)tlen + 1);
    int i, j, k, score, n = 0, which = 0;
    eh_h[0] = 0; eh_e[0] = BSWREF_MINUS_INF;
    for (j = 1; j <= qlen && j <= w; ++j) { eh_h[j] = -(o_ins + e_ins * j); eh_e[j] = BSWREF_MINUS_INF; }
    for (; j <= qlen; ++j) eh_h[j] = eh_e[j] = BSWREF_MINUS_INF;                  /* everything outside the band */
    for (i = 0; i < tlen; ++i) {
        int32_t f = BSWREF_MINUS_INF, h1;
        const int8_t *srow = &p->mat[5 * target[i]];
        const int beg = i > w ? i - w : 0;
        const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
        uint8_t *zi = &z[(size_t)i * (size_t)n_col];
        h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : BSWREF_MINUS_INF;
        for (j = beg; j < end; ++j) {
            int32_t h, m = eh_h[j], e = eh_e[j], t;
            uint8_t d;
            eh_h[j] = h1;
            m += srow[query[j]];
            d = m >= e ? 0 : 1;
            h = m >= e ? m : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            h1 = h;
            t = m - oe_del;
            e -= e_del;
            d |= e > t ? 1 << 2 : 0;
            e = e > t ? e : t;
            eh_e[j] = e;
            t = m - oe_ins;
            f -= e_ins;
            d |= f > t ? 2 << 4 : 0;
            f = f > t ? f : t;
            zi[j - beg] = d;
        }
        eh_h[end] = h1; eh_e[end] = BSWREF_MINUS_INF;
    }
    score = eh_h[qlen];
    i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;                  /* (i,k): the last cell */
    while (i >= 0 && k >= 0 && n >= 0) {
        which = z[(size_t)i * (size_t)n_col + (size_t)(k - (i > w ? i - w : 0))] >> (which << 1) & 3;
        if (which == 0) { n = push_op(cigar, n, max_cigar, 0, 1); --i; --k; }
        else if (which == 1) { n = push_op(cigar, n, max_cigar, 2, 1); --i; }
        else { n = push_op(cigar, n, max_cigar, 1, 1); --k; }
    }
    if (n >= 0 && i >= 0) n = push_op(cigar, n, max_cigar, 2, i + 1);
    if (n >= 0 && k >= 0) n = push_op(cigar, n, max_cigar, 1, k + 1);
    for (i = 0; n > 0 && i < n >> 1; ++i) { const uint32_t tmp = cigar[i]; cigar[i] = cigar[n - 1 - i]; cigar[n - 1 - i] = tmp; }
    *n_cigar = n;
    free(eh_h); free(eh_e); free(z);
    return score;
}
