/*
 * bsw.h -- C ABI of the B200-native batched seed-extension library (libbsw.so).
 *
 * Drop-in boundary for the one hot path peterpengwei/bwa-mem-sw accelerates on an FPGA:
 * BWA-MEM's banded affine-gap Smith-Waterman seed extension (the ksw_extend2 recurrence).
 * The reference "plugin API" is the AFU batch contract between the BWA host and the FPGA
 * (task-batch buffer in, result-batch buffer out, start/poll by CSR/DSM); this header replaces
 * it with plain pointers + sizes.  All citations are into /root/reference (read-only RTL).
 *
 *   level 0  lifecycle                     replaces AAL/CCI session + CSR setup (batch_manager.v:208-213,313-351)
 *   level 1  bsw_extend_batch[_flat]       replaces sw_extend = one ksw_extend2 call per task
 *                                          (ports sw_pe_array_sw_extend.v:96-123; outputs in its return order :117-123)
 *   level 2  bsw_chain2aln_batch           replaces one proc_element task = left+right extension, band retry,
 *                                          clip decision (sw_pe_array_proc_element.v:1593-1685), record order of
 *                                          fill_resulBuf (sw_pe_array_fill_resulBuf.v:377-429; pe:1187-1205)
 *   level 3  bsw_fpga_batch                consumes a TBB image / produces an RBB image bit-for-bit in the FPGA layout
 *                                          (tbb.v:163-194, rbb.v:117-167, sw_pe_array_task_parse.v:924-948)
 *   async    bsw_submit / bsw_poll / bsw_wait   mirrors "write REQ_PEARRAY / poll DSM busy bit"
 *                                          (batch_manager.v:347,851-854)
 *
 * Ownership: the caller owns every in/out array; the library copies into its own pinned staging and
 * never keeps a caller pointer after a blocking call returns (async: until bsw_wait returns).
 * Errors: 0 on success, negative BSW_E* otherwise; never aborts; there is NO CPU fallback -- without a
 * usable CUDA device bsw_init fails with BSW_ECUDA.
 * Threading: blocking calls are re-entrant; a context may be shared by many host threads.
 */
#ifndef BSW_H
#define BSW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSW_OK        0
#define BSW_EINVAL   -1   /* bad argument: null pointer, qlen<1, tlen<1, h0<1, e_ins<1, e_del<1, base code >4, ... */
#define BSW_ECUDA    -2   /* CUDA runtime error (text via bsw_last_error) */
#define BSW_ENOMEM   -3   /* host or device allocation failed */
#define BSW_ERANGE   -4   /* outside the numeric envelope: qlen > 2^20, tlen > 2^22, or h0 + qlen*(max(mat)+e_ins) >= 2^30
                             (with option "wide" = 0: outside the 16-bit envelope h0 + qlen*max(mat) <= 32767, qlen <= 40000) */
#define BSW_EWIRE    -5   /* malformed TBB image */
#define BSW_EBUSY    -6   /* no free async slot / ticket not finished */

#define BSW_VARIANT_V1 1  /* recurrence of the reference RTL (= BWA 0.7.8 era); default */
#define BSW_VARIANT_V2 2  /* recurrence of current upstream BWA ksw_extend2 */

typedef struct bsw_ctx bsw_ctx;

/* ---------------- level 0: lifecycle ---------------- */
/* device_ids == NULL or n_devices <= 0: use the current CUDA device only.  streams_per_device <= 0: default (2). */
int  bsw_init(bsw_ctx **ctx, const int *device_ids, int n_devices, int streams_per_device);
void bsw_destroy(bsw_ctx *ctx);
const char *bsw_last_error(const bsw_ctx *ctx);          /* thread-unsafe convenience: last error text of this ctx */
const char *bsw_version(void);
/* Options (all optional): "variant" {1,2}; "host_threads" N; "chunk_tasks" N (pipeline granularity, default 16384);
 * "slots" N (chunks one host worker keeps in flight, default 3); "raw_inputs" how a flat batch's bases reach the GPU:
 * 0 = 4-bit staged by host threads (round-1 path), 1 = copied in place from registered buffers (1 byte per base, no host
 * pass), 3 = packed 2 bit per base by host threads, 2 = auto (default: 1 for registered buffers when the GPU has fewer
 * than 12 host threads, else 3);
 * "force_kernel" {0 auto, 1 inter-task K1, 2 intra-task K2}; "k2_min_qlen" N (tasks with qlen >= N use K2 in auto mode);
 * "k2_warps" {1,4} warps per K2 task; "fused_l2" {0,1} level 2 as one fused kernel (default 1);
 * "device_plan" {1,0} sort + tile building on the device / on the host; "k2_narrow" {1,0} register path for narrow K2 rows;
 * "wide" {1,0,2} tasks outside the 16-bit envelope of K1 / K2 (h0 + qlen*max(mat) > 32767, qlen > 40000 or tlen > 500000): 1 = the
 * batch is split and they run on the 32-bit kernel K5 (default), 0 = the batch is refused with BSW_ERANGE, 2 = every task on K5;
 * "fpga_strict" {0,1} bsw_fpga_batch refuses tasks outside the FPGA's 8-bit envelope;
 * "kernel_timing" {0,1} record CUDA events around each chunk's kernels for bsw_stats.kernel_ms (default 0:
 *                       the batch calls then leave kernel_ms at 0; bsw_resident_run always times its launches) */
int  bsw_set_option(bsw_ctx *ctx, const char *key, int64_t value);
int  bsw_num_devices(const bsw_ctx *ctx);

/* ---------------- level 1: raw ksw_extend2 batch ---------------- */
typedef struct {
    int8_t  mat[25];          /* 5x5 score matrix, index 5*target_base+query_base (sw_pe_array_sw_extend.v:1915-1940) */
    int32_t o_del, e_del;     /* gap open/extend, deletion (TBB header word 0: proc_element.v:815-820)          */
    int32_t o_ins, e_ins;     /* gap open/extend, insertion                                                     */
    int32_t zdrop;            /* ksw_extend2 z-drop threshold, <=0 disables (not in the RTL)                    */
    int32_t end_bonus;        /* enters only the band clamp max_ins/max_del (host-side in the RTL: pe:924-934)  */
} bsw_params;

typedef struct {
    const uint8_t *query;     /* qlen bases, 1 per byte, codes 0..3 = ACGT, 4 = N; caller-owned, read-only */
    const uint8_t *target;    /* tlen bases */
    int32_t qlen, tlen;       /* >= 1 (the PE never calls sw_extend with qlen==0: proc_element.v:1670) */
    int32_t h0;               /* > 0 */
    int32_t w;                /* band width for this (single) band try */
} bsw_task;

/* sw_extend's return tuple without aw (sw_pe_array_sw_extend.v:117-123,1315-1375) */
typedef struct { int32_t score, qle, tle, gtle, gscore, max_off; } bsw_result;

/* out[i] <-> tasks[i].  ONE band try per task with the given w (the MAX_BAND_TRY loop lives in level 2). */
int bsw_extend_batch(bsw_ctx *ctx, const bsw_params *params, const bsw_task *tasks, size_t n, bsw_result *out);

/* Registered host memory (optional).  A flat batch whose qbuf and tbuf ranges lie inside buffers registered here is
 * not staged by host threads: the copy engine reads the bases in place (one code per byte) and the device packs them.
 * The call page-locks the range (cudaHostRegister); register long-lived batch buffers once, not per call.  Tasks that
 * hold an N are rerun on the staged path automatically; results and error codes are the same either way. */
int bsw_host_register(bsw_ctx *ctx, const void *ptr, size_t bytes);
int bsw_host_unregister(bsw_ctx *ctx, const void *ptr);

/* Same, flat layout: task i's query is qbuf[qoff[i] .. qoff[i+1]), target likewise (n+1 offsets each).
 * cells (optional, may be NULL): per-task number of DP cells actually evaluated (sum over rows of end-beg). */
int bsw_extend_batch_flat(bsw_ctx *ctx, const bsw_params *params,
                          const uint8_t *qbuf, const int64_t *qoff,
                          const uint8_t *tbuf, const int64_t *toff,
                          const int32_t *h0, const int32_t *w, size_t n,
                          bsw_result *out, uint32_t *cells);

/* ---------------- level 2: fused seed task (one FPGA PE task) ---------------- */
typedef struct { bsw_params p; int32_t w, pen_clip5, pen_clip3; } bsw_params2;   /* TBB header words 0-1 */

typedef struct {
    const uint8_t *q_left;    /* left query flank, already reversed (the PE reads both flanks forward) */
    const uint8_t *q_right;
    const uint8_t *t_left;    /* left target flank, already reversed */
    const uint8_t *t_right;
    int32_t qlen[2], tlen[2]; /* [0]=left, [1]=right; qlen may be 0 = no extension on that side (pe:1670) */
    int32_t init_score;       /* regScore: a->score before the task (pe:871-874) */
    int32_t qbeg;             /* qBeg_ori: seed start on the query */
    int32_t h0;               /* seed_len * a (left extension only; right uses the left score) */
    uint32_t id;              /* opaque, echoed (pe:807,1199) */
} bsw_seed_task;

/* fill_resulBuf's 5-word record, unpacked: [id][qe<<16|qb][re<<16|rb][truesc<<16|score][w]; rb/re/qe are
 * relative to the seed exactly as in the RTL (pe:1662-1665). */
typedef struct { uint32_t id; int32_t qb, qe, rb, re, score, truesc, w; } bsw_aln_record;

int bsw_chain2aln_batch(bsw_ctx *ctx, const bsw_params2 *params, const bsw_seed_task *tasks, size_t n,
                        bsw_aln_record *out);

/* ---------------- host task builder: mem_chain2aln either side of the kernel ---------------- */
/* What feeds param words 0-7 of a PE task (sw_pe_array_proc_element.v:815-934) from a read, a reference window and a
 * chain of seeds -- host work in the (unmounted) quickassist port of BWA 0.7.8; restated from the published BWA-MEM
 * algorithm (mem_chain2aln, cal_max_gap).  Coordinates: query positions in the read, rbeg in the 2*l_pac forward+reverse
 * reference space, rseq = the reference bases of [rmax0, rmax1). */
typedef struct { int32_t a, o_del, e_del, o_ins, e_ins, w; } bsw_chain_opt;       /* the mem_opt_t fields involved */
typedef struct { int64_t rbeg; int32_t qbeg, len; } bsw_chain_seed;               /* mem_seed_t */
typedef struct { int64_t rb, re; int32_t qb, qe, score, truesc, w, seedcov; } bsw_seed_aln;   /* the mem_alnreg_t fields the PE decides */
/* rmax[0], rmax[1]: the reference span the chain's extensions may touch (flank + cal_max_gap either side, clipped to
 * [0, 2*l_pac) and to the seed's side of the forward/reverse boundary). */
int bsw_chain_window(const bsw_chain_opt *opt, int l_query, const bsw_chain_seed *seeds, int n, int64_t l_pac, int64_t rmax[2]);
/* One bsw_seed_task per seed: left flanks reversed into `scratch` (bsw_seed_scratch_bytes gives a sufficient size), right
 * flanks point into query / rseq; h0 = len * a; init_score = -1 with a left flank, len * a without; id = seed index. */
size_t bsw_seed_scratch_bytes(int l_query, int64_t rmax0, int64_t rmax1, int n);
int bsw_build_seed_tasks(const bsw_chain_opt *opt, const uint8_t *query, int l_query, const uint8_t *rseq, int64_t rmax0,
                         int64_t rmax1, const bsw_chain_seed *seeds, int n, uint8_t *scratch, size_t scratch_bytes,
                         bsw_seed_task *out);
/* Record (relative to the seed, as the PE returns it) -> absolute coordinates, with BWA's rule for a seed without flanks. */
void bsw_finish_seed(const bsw_chain_opt *opt, const bsw_chain_seed *seed, int l_query, const bsw_aln_record *rec, bsw_seed_aln *out);

/* ---------------- after the extension: banded global alignment with traceback (ksw_global2) ---------------- */
/* the DP BWA-MEM runs once per reported alignment (bwa_gen_cigar2 -> ksw_global2) to turn the extension's
 * end points into a CIGAR.  Not part of the reference tree (the FPGA stops at the extension); follows the published BWA
 * algorithm.  Scores are int32.  cigar[i*max_ops .. +n_cigar[i]) = task i's operations in BAM encoding (len << 4 | op;
 * op 0 = M, 1 = I, 2 = D).  BSW_ERANGE if an alignment needs more than max_ops operations. */
typedef struct { const uint8_t *query, *target; int32_t qlen, tlen, w; } bsw_global_task;
int bsw_global_batch(bsw_ctx *ctx, const bsw_params *params, const bsw_global_task *tasks, size_t n, int max_ops,
                     int32_t *score, int32_t *n_cigar, uint32_t *cigar);

/* ---------------- level 3: FPGA wire format ---------------- */
#define BSW_TBB_WORDS 65536   /* 4096 x 64 B (bwa_mem_sw.v:163-166) */
#define BSW_RBB_WORDS 4096    /*  256 x 64 B (bwa_mem_sw.v:167-170) */
/* tbb_words: BSW_TBB_WORDS u32 (layout SURVEY.md App. A.1); rbb_words: BSW_RBB_WORDS u32, records written densely
 * from word 0 in task order (a valid completion order); words past 5*n_results are left untouched, like the RBB. */
int bsw_fpga_batch(bsw_ctx *ctx, const uint32_t *tbb_words, uint32_t *rbb_words, int *n_results);
/* Helpers for hosts/tests that build or read the images (host side of the AFU contract). */
int bsw_tbb_encode(const bsw_params2 *params, const bsw_seed_task *tasks, size_t n, uint32_t *tbb_words);
/* How many tasks of the image lie outside the envelope in which the FPGA's 8-bit datapath equals ksw_extend2 (scores
 * <= 127, flanks <= 127 bases, w <= 63, no wrap of the first column / first row generators: sw_pe_array_sw_extend.v:155-159,
 * 1795,1975,770).  bsw_fpga_batch computes ksw_extend2's answer for those tasks, which the FPGA does
 * not; with option "fpga_strict" = 1 it returns BSW_ERANGE instead.  first_outside may be NULL. */
int bsw_fpga_envelope(const uint32_t *tbb_words, int *n_outside, int *first_outside);
int bsw_rbb_decode(const uint32_t *rbb_words, size_t n, bsw_aln_record *out);

/* ---------------- async pair (level 1) ---------------- */
typedef struct { int32_t slot; uint32_t seq; } bsw_ticket;
int bsw_submit(bsw_ctx *ctx, const bsw_params *params, const bsw_task *tasks, size_t n, bsw_result *out,
               bsw_ticket *ticket);
int bsw_poll(bsw_ctx *ctx, const bsw_ticket *ticket);    /* 1 = finished, 0 = still running, <0 = error */
int bsw_wait(bsw_ctx *ctx, const bsw_ticket *ticket);    /* blocks; returns the batch status */

/* ---------------- device-resident batches (measurement with inputs already in HBM) ---------------- */
typedef struct bsw_resident bsw_resident;
int  bsw_resident_create(bsw_ctx *ctx, const bsw_params *params,
                         const uint8_t *qbuf, const int64_t *qoff, const uint8_t *tbuf, const int64_t *toff,
                         const int32_t *h0, const int32_t *w, size_t n, bsw_resident **r);   /* pack + schedule + H2D */
/* kernels only, inputs already in HBM.  kernel_ms = CUDA-event time of the launches on their own stream,
 * cells = DP cells evaluated (device-counted), launches = kernel launches issued; each may be NULL. */
int  bsw_resident_run(bsw_ctx *ctx, bsw_resident *r, double *kernel_ms, uint64_t *cells, uint64_t *launches);
int  bsw_resident_fetch(bsw_ctx *ctx, bsw_resident *r, bsw_result *out, uint32_t *cells);   /* D2H */
void bsw_resident_free(bsw_ctx *ctx, bsw_resident *r);

/* ---------------- statistics of the calls since the last reset ---------------- */
typedef struct {
    uint64_t tasks;            /* extensions executed */
    uint64_t cells_band;       /* DP cells evaluated (device-counted, same definition as the oracle's cells) */
    uint64_t kernel_launches;  /* extension-kernel launches */
    uint64_t h2d_bytes, d2h_bytes;
    double   kernel_ms;        /* sum of CUDA-event durations of the extension kernels (on their own streams) */
    double   pack_ms;          /* host: validate + schedule + pack */
    double   wall_ms;          /* host wall time inside blocking calls */
} bsw_stats;
int bsw_get_stats(bsw_ctx *ctx, bsw_stats *out);
int bsw_reset_stats(bsw_ctx *ctx);

/* ---------------- INT-pipe micro-benchmark (roofline denominator) ---------------- */
typedef struct {
    double iadd_tops;          /* independent IADD3 streams, scalar int32 ops/s */
    double vimnmx_tops;        /* independent VIMNMX (max) streams */
    double dpx_tops;           /* VIADDMNMX streams, counted as 2 ops/instruction */
    double mix_tops;           /* the DP's own add/max mix (13-op cell body without memory), ops/s */
    double dual_tops;          /* adds issued alternately as IADD3 (ALU pipe) and IMAD (FMA pipe) */
    double sm_clock_mhz;       /* clock observed during the run */
    int    sm_count;
} bsw_int_peak;
int bsw_measure_int_peak(bsw_ctx *ctx, int device_index, bsw_int_peak *out);

#ifdef __cplusplus
}
#endif
#endif /* BSW_H */
