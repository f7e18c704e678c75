/*
 * oracle/rtl_width_model.c -- TEST INFRASTRUCTURE (part of oracle/libbswref.so).
 *
 * The same recurrence as ksw_extend_ref.c, but at the WIDTHS of the mounted RTL: 8-bit H / E / F / h0 / max compared
 * signed, positives clipped through 7-bit slices, running 8-bit subtractions for the first column and the first row,
 * an 8-bit band, maxima initialised once per invocation.  Its job is to EXPLAIN
 * what the FPGA returns outside the envelope in which it agrees with ksw_extend2 (oracle.rtl_envelope): the
 * translated RTL (oracle/_ref, tests/golden/rtl_sw_extend_wide.npz) is the judge, this file is the explanation.
 * Column indices are plain ints here: the RTL sign-extends its 8-bit mj in the max_off and narrowing compares
 * (sw_pe_array_sw_extend.v:1654,1336,1547), which this model does not follow, so it is only claimed for qlen <= 127.
 * `sx:N` = sw_pe_array_sw_extend.v line N.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int s8(int x) { return (int8_t)(x & 0xff); }          /* 8-bit two's complement value of x */
static inline int u8(int x) { return x & 0xff; }
static inline int sx_bits(int x, int bits) { int m = 1 << (bits - 1); x &= (1 << bits) - 1; return (x ^ m) - m; }
/* "tmp > 0 ? tmp[6:0] : 0" -- the RTL's relu on an 8-bit signed value (sx:1862,1865,1978,1835) */
static inline int clip7(int x8) { return s8(x8) > 0 ? (x8 & 0x7f) : 0; }

/* mat: the RTL's hard-wired +1 / -4 / -1 (sx:1915-1940); bases use the low 3 bits of a nibble (sx:1883,1885). */
static inline int rtl_score(int t, int q)
{
    t &= 7; q &= 7;
    if (t > 4 || q > 4) return 0;                 /* mux inputs 26..32 do not exist: sel >= 25 reads as 0 in the cycle model */
    if (t == 4 || q == 4) return -1;
    return t == q ? 1 : -4;
}

/* One invocation of sw_extend at RTL widths.  out7 = ap_return_0..6 (score, aw, qle, tle, gtle, gscore, max_off). */
void bswref_sw_extend_rtl8(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int o_ins, int e_ins,
                           int o_del, int e_del, int w_in, int h0, int reg_score, int max_ins, int max_del, int32_t *out7)
{
    int eh_h[257], eh_e[257];
    const int oe_ins = u8(e_ins + o_ins);                               /* sx:1823 */
    const int sum = u8(o_del + e_del);                                  /* sx:1860 */
    int max = u8(h0), max_i = 0xfff, max_j = 0xff, max_off = 0;         /* sx:889,919,929; p_2 (max_i) starts at -1 */
    int gscore12 = 0xfff, max_ie12 = 0xfff;                             /* 12-bit registers, -1 */
    int val = reg_score & 0xffff, aw_ret = 0;
    int k, cont = 1;
    memset(eh_h, 0, sizeof eh_h); memset(eh_e, 0, sizeof eh_e);
    for (k = 0; k < 2 && cont; ++k) {                                   /* sx:1963,1878 */
        const int prev = val & 0x3ff;                                   /* sx:1839 */
        const int aw_tmp = u8(w_in << k);                               /* sx:1765 */
        const int aw2 = (s8(aw_tmp) < sx_bits(max_ins, 16)) ? aw_tmp : u8(max_ins);       /* sx:1764,1881 */
        const int aw1 = (s8(aw2) < sx_bits(max_del, 16)) ? aw2 : u8(max_del);             /* sx:1763,1890 */
        const int awx = s8(aw1);                                        /* sign-extended where it meets i (sx:1850-1851) */
        int beg = 0, end = u8(qlen), i, h1run = u8(h0 - o_del);         /* sx:769,779,1796 */
        aw_ret = aw_tmp;
        for (i = 0; i < tlen; ++i) {                                    /* sx:1891 */
            int f = 0, m = 0, mj = 0xff, j, jbeg, jend, h1, hrow, lastj;
            h1run = u8(h1run - e_del);                                  /* sx:1795: a running 8-bit subtraction ... */
            h1 = (h1run & 0x80) ? 0 : (h1run & 0x7f);                   /* ... clipped by its sign bit (sx:890-907) */
            jbeg = (beg < i - awx) ? u8(i - aw1) : beg;                 /* sx:1846,1894,1895,1803 */
            jend = (end > i + awx + 1) ? u8(1 + i + aw1) : end;         /* sx:1843,1897,1869,1778 */
            if (jend > qlen) jend = u8(qlen);                           /* sx:1898,1842 (unsigned 8-bit compare) */
            hrow = clip7(u8(h0 - oe_ins));                              /* first-row generator: starts from the clipped eh[1].h (sx:1059,688),
                                                                           then runs on as an UNCLIPPED 8-bit subtraction (sx:1975,857) */
            lastj = jbeg > jend ? jbeg : jend;                          /* sx:1768: value of j after the loop */
            for (j = jbeg; j < jend; ++j) {                             /* sx:1901 */
                int M, e, h, t, sc;
                if (i == 0) {                                           /* sx:1900,1819-1820,1773: generated, not read */
                    if (j == 0) M = u8(h0);
                    else if (j == 1) M = clip7(u8(h0 - oe_ins));        /* sx:1979,1957,1974 */
                    else { hrow = u8(hrow - e_ins); M = clip7(hrow); }  /* sx:1975-1978,1821 */
                    e = 0;
                } else { M = eh_h[j]; e = eh_e[j]; }                    /* sx:1799,1772 */
                eh_h[j] = h1;                                           /* sx:1776 */
                sc = rtl_score(target[i], query[j]);
                h = u8(M + sc);                                         /* sx:1797 8-bit add */
                h = s8(h) > s8(e) ? h : e;                              /* sx:1798 */
                h = s8(h) > s8(f) ? h : f;                              /* sx:1809 */
                h1 = h;                                                 /* sx:847 */
                if (!(s8(m) > s8(h))) mj = j;                           /* sx:1816 */
                m = s8(m) > s8(h) ? m : h;                              /* sx:1808 */
                t = clip7(u8(h - sum));                                 /* sx:1866,1862 */
                e = u8(e - e_del);                                      /* sx:1770 */
                e = s8(e) > t ? e : t;                                  /* sx:1771 */
                eh_e[j] = e;
                t = clip7(u8(h - oe_ins));                              /* sx:1863,1865 */
                f = u8(f - e_ins);                                      /* sx:1780 */
                f = s8(f) > t ? f : t;                                  /* sx:1781 */
            }
            eh_h[jend] = h1; eh_e[jend] = 0;                            /* sx:1775,1904 */
            if (lastj == u8(qlen)) {                                    /* sx:1913 */
                if (!(sx_bits(gscore12, 12) > s8(h1))) { max_ie12 = i & 0xfff; gscore12 = s8(h1) & 0xfff; }   /* sx:1941,1829,1831 */
            }
            if (m == 0) break;                                          /* sx:1942 */
            if (s8(m) > s8(max)) {                                      /* sx:1959 */
                int d = sx_bits(mj, 8) - i, a;                          /* sx:1654: mj sign-extended */
                max = m; max_i = i & 0xfff; max_j = mj;
                a = d > 0 ? (d & 0x3ff) : ((0 - (d & 0x3ff)) & 0x3ff);  /* sx:1663-1679 */
                if (sx_bits(max_off, 12) < sx_bits(a, 10)) max_off = sx_bits(a, 10) & 0xfff;      /* sx:1786,1792 */
            }
            for (j = mj; j >= jbeg && eh_h[j]; --j) ;                   /* narrowing as in V1 (plain indices, see header) */
            beg = j + 1;
            for (j = mj + 2; j <= jend && eh_h[j]; ++j) ;
            end = j;
        }
        val = sx_bits(max, 8) & 0xffff;                                 /* sx:1067,1716 */
        {
            const int op2 = sx_bits(aw_tmp >> 1, 7) + sx_bits(aw_tmp >> 2, 6);             /* sx:1724-1750 */
            const int differs = (sx_bits(max, 8) & 0x3ff) != prev;                          /* sx:1760 */
            cont = differs && !(sx_bits(max_off, 12) < op2);                                 /* sx:1765-1777 */
        }
    }
    out7[0] = sx_bits(max, 8);
    out7[1] = sx_bits(aw_ret, 8);
    out7[2] = sx_bits(max_j, 8) + 1;                                    /* sx:1841 */
    out7[3] = sx_bits(max_i, 12) + 1;                                   /* sx:1868 */
    out7[4] = sx_bits(max_ie12, 12) + 1;                                /* sx:1794 */
    out7[5] = sx_bits(gscore12, 12);
    out7[6] = sx_bits(max_off, 12);
}
