/* A C caller of libbsw.so, compiled against include/bsw.h by tests/test_chain_builder.py (gcc, not ctypes): the way
 * BWA's mem_chain2aln would use the host task builder and the level-2 batch call.
 *   stdin : l_query l_pac n w   then the read, the forward reference (l_pac bases) as digit strings, then n x (rbeg qbeg len)
 *   argv[1] = "host": print the window and the seed tasks only (no GPU needed)
 *   argv[1] = "gpu" : also run bsw_chain2aln_batch and print the records and the finished alignments */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bsw.h"

static void put(const char *tag, const uint8_t *p, int n)
{
    printf("%s ", tag);
    for (int i = 0; i < n; ++i) putchar('0' + p[i]);
    if (!n) putchar('-');
    putchar('\n');
}

int main(int argc, char **argv)
{
    int l_query, n, w;
    long long l_pac;
    if (scanf("%d %lld %d %d", &l_query, &l_pac, &n, &w) != 4) return 2;
    char *rd = malloc((size_t)l_query + 2), *rf = malloc((size_t)l_pac + 2);
    if (scanf("%s %s", rd, rf) != 2) return 2;
    uint8_t *query = malloc((size_t)l_query), *ref = malloc((size_t)l_pac);
    for (int i = 0; i < l_query; ++i) query[i] = (uint8_t)(rd[i] - '0');
    for (long long i = 0; i < l_pac; ++i) ref[i] = (uint8_t)(rf[i] - '0');
    bsw_chain_seed *seeds = malloc(sizeof(*seeds) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        long long rb; int qb, len;
        if (scanf("%lld %d %d", &rb, &qb, &len) != 3) return 2;
        seeds[i].rbeg = rb; seeds[i].qbeg = qb; seeds[i].len = len;
    }
    bsw_chain_opt opt = { 1, 6, 1, 6, 1, w };
    int64_t rmax[2];
    int rc = bsw_chain_window(&opt, l_query, seeds, n, l_pac, rmax);
    if (rc) { printf("window error %d\n", rc); return 1; }
    printf("rmax %lld %lld\n", (long long)rmax[0], (long long)rmax[1]);
    const uint8_t *rseq = ref + rmax[0];                       /* the test keeps the chain on the forward strand */
    size_t sb = bsw_seed_scratch_bytes(l_query, rmax[0], rmax[1], n);
    uint8_t *scratch = malloc(sb);
    bsw_seed_task *tasks = malloc(sizeof(*tasks) * (size_t)n);
    rc = bsw_build_seed_tasks(&opt, query, l_query, rseq, rmax[0], rmax[1], seeds, n, scratch, sb, tasks);
    if (rc) { printf("build error %d\n", rc); return 1; }
    for (int i = 0; i < n; ++i) {
        const bsw_seed_task *t = &tasks[i];
        printf("task %u qlen %d %d tlen %d %d init %d qbeg %d h0 %d\n", t->id, t->qlen[0], t->qlen[1], t->tlen[0], t->tlen[1],
               t->init_score, t->qbeg, t->h0);
        put("ql", t->q_left, t->qlen[0]); put("tl", t->t_left, t->tlen[0]);
        put("qr", t->q_right, t->qlen[1]); put("tr", t->t_right, t->tlen[1]);
    }
    if (argc > 1 && !strcmp(argv[1], "gpu")) {
        bsw_ctx *ctx = NULL;
        if (bsw_init(&ctx, NULL, 0, 0)) { printf("no device\n"); return 3; }
        bsw_params2 P;
        memset(&P, 0, sizeof P);
        for (int t = 0; t < 5; ++t)
            for (int q = 0; q < 5; ++q) P.p.mat[5 * t + q] = (int8_t)((t == 4 || q == 4) ? -1 : (t == q ? 1 : -4));
        P.p.o_del = 6; P.p.e_del = 1; P.p.o_ins = 6; P.p.e_ins = 1; P.p.zdrop = 100; P.p.end_bonus = 5;
        P.w = w; P.pen_clip5 = 5; P.pen_clip3 = 5;
        bsw_aln_record *rec = malloc(sizeof(*rec) * (size_t)n);
        rc = bsw_chain2aln_batch(ctx, &P, tasks, (size_t)n, rec);
        if (rc) { printf("batch error %d: %s\n", rc, bsw_last_error(ctx)); return 1; }
        for (int i = 0; i < n; ++i) {
            bsw_seed_aln a;
            bsw_finish_seed(&opt, &seeds[i], l_query, &rec[i], &a);
            printf("rec %u %d %d %d %d %d %d %d\n", rec[i].id, rec[i].qb, rec[i].qe, rec[i].rb, rec[i].re, rec[i].score, rec[i].truesc, rec[i].w);
            printf("aln %d %d %lld %lld %d %d %d\n", a.qb, a.qe, (long long)a.rb, (long long)a.re, a.score, a.truesc, a.w);
        }
        bsw_destroy(ctx);
    }
    return 0;
}
