#!/usr/bin/env python3
"""One-off refactoring helper (kept for the record): moves the ENCODER-side restatements out of oracle/
into datagen/ so that oracle/ is decoder-side checker only.  Already applied; running it again fails."""
import os
import re

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) + "/"


def rd(p):
    return open(R + p).read()


def wr(p, s):
    os.makedirs(os.path.dirname(R + p), exist_ok=True)
    open(R + p, "w").write(s)


def cut(src, start_pat, end_pat):
    i = src.index(start_pat)
    j = src.index(end_pat, i)
    return src[:i] + src[j:], src[i:j]


# ---------- MQ ----------
mq = rd("oracle/orc_mq.c")
i_enc = mq.index("/* ---- encoder: NewMQEncoder")
i_dec = mq.index("/* ---- decoder: NewMQDecoder")
i_flat = mq.index("/* ---- flat test entry points")
enc_part, dec_part, flat, head = mq[i_enc:i_dec], mq[i_dec:i_flat], mq[i_flat:], mq[:i_enc]
i_fd = flat.index("void orc_mq_decode(")
flat_enc, flat_dec = flat[:i_fd], flat[i_fd:]
wr("oracle/orc_mq.c", head + dec_part +
   "/* ---- flat test entry point ---------------------------------------------------- */\n" + flat_dec)
tbl = head[head.index("uint32_t orc_mq_qe[94];"):]
gen_mq = ('''/*
 * gen_mq.c -- restatement of the reference MQ ENCODER (internal/entropy/mqc.go:169-349).
 * Part of datagen/: the synthetic-input generator (reference encoder side).  Not the oracle,
 * not the product: it only manufactures code-block bitstreams for tests and bench inputs.
 */
#include "datagen.h"
#include "gen_mq.h"
#include <string.h>

''' + tbl + enc_part + flat_enc.replace(
    "/* ---- flat test entry points -------------------------------------------------- */\n",
    "/* ---- flat entry point ------------------------------------------------------------ */\n"))
for a, b in [("orc_mq_tables_init", "gen_mq_tables_init"), ("orc_mqenc", "gen_mqenc"), ("orc_mq_qe", "gen_mq_qe"),
             ("orc_mq_nmps", "gen_mq_nmps"), ("orc_mq_nlps", "gen_mq_nlps"), ("ORC_CTX_UNI", "GEN_CTX_UNI"),
             ("orc_mq_encode", "gen_mq_encode")]:
    gen_mq = gen_mq.replace(a, b)
wr("datagen/gen_mq.c", gen_mq)
wr("datagen/gen_mq.h", '''/* gen_mq.h -- MQ encoder state for datagen (reference encoder side; mqc.go:169-349). */
#ifndef GEN_MQ_H
#define GEN_MQ_H
#include <stdint.h>
enum { GEN_CTX_ZC0 = 0, GEN_CTX_SC0 = 9, GEN_CTX_MAG0 = 14, GEN_CTX_RL = 17, GEN_CTX_UNI = 18, GEN_NUM_CTX = 19 };
extern uint32_t gen_mq_qe[94];
extern uint8_t  gen_mq_nmps[94];
extern uint8_t  gen_mq_nlps[94];
void gen_mq_tables_init(void);
typedef struct {
    uint32_t A, C, CT;
    uint8_t *buf; int cap; int bp; int overflow;
    uint8_t ctx[GEN_NUM_CTX];
} gen_mqenc;
void gen_mqenc_init(gen_mqenc *e, uint8_t *buf, int cap);
void gen_mqenc_encode(gen_mqenc *e, int ctx, int d);
int  gen_mqenc_flush(gen_mqenc *e, const uint8_t **start);
#endif
''')
h = rd("oracle/orc_mq.h")
h = re.sub(r"typedef struct \{\n    uint32_t A, C, CT;\n    uint8_t \*buf;.*?\} orc_mqenc;\n\n", "", h, flags=re.S)
h = h.replace("void orc_mqenc_init(orc_mqenc *e, uint8_t *buf, int cap);\n"
              "void orc_mqenc_encode(orc_mqenc *e, int ctx, int d);\n"
              "int  orc_mqenc_flush(orc_mqenc *e, const uint8_t **start);   /* returns length, *start = first payload byte */\n", "")
wr("oracle/orc_mq.h", h)

# ---------- T1 ----------
t1 = rd("oracle/orc_t1.c")
i_e = t1.index("/* ============================ encoder")
enc, dec = t1[i_e:], t1[:i_e]
wr("oracle/orc_t1.c", dec)
i_d = dec.index("/* ============================ decoder")
shared = dec[dec.index("/* flag bits, t1.go:72-91"):i_d]
gen_t1 = ('''/*
 * gen_t1.c -- restatement of the reference EBCOT tier-1 ENCODER: T1.SetData (t1.go:292-304) and
 * T1.Encode = EncodeFast5 (t1_fast5.go:10-899; same decisions as EncodeSafe t1.go:923-947), with the
 * context rules of t1.go:349-479 / t1_luts.go:35-110.  Part of datagen/ (synthetic-input generator).
 */
#include "datagen.h"
#include "gen_mq.h"
#include <stdlib.h>
#include <string.h>

''' + shared + enc)
for a, b in [("orc_mqenc", "gen_mqenc"), ("orc_t1_encode", "gen_t1_encode"), ("ORC_CTX_", "GEN_CTX_"),
             ("ORC_BAND_", "GEN_BAND_"),
             ("const uint8_t *orc_t1_zc_lut(void) { zc_lut_init(); return g_zc_lut; }\n", "")]:
    gen_t1 = gen_t1.replace(a, b)
wr("datagen/gen_t1.c", gen_t1)

# ---------- HT ----------
ht = rd("oracle/orc_ht.c")
i_e = ht.index("/* ------------------------------ encoder")
wr("oracle/orc_ht.c", ht[:i_e])
wr("datagen/gen_ht.c", '''/*
 * gen_ht.c -- restatement of the reference "HT" block ENCODER, HTEncoder.Encode
 * (internal/entropy/ht.go:942-1391).  Part of datagen/ (synthetic-input generator).  The reference
 * encoder is not ISO/IEC 15444-15 and its output does not round-trip through the reference decoder
 * (zero-filled MEL segment of maxSize/4 bytes, byte-reversed VLC segment, ht.go:978,1019,1036-1038);
 * it is restated as it is because its bytes are what the reference would hand its own decoder.
 */
#include "datagen.h"
#include <stdlib.h>
#include <string.h>

#include "ht_vlc_tables.inc"
static const uint16_t k_vlc_tbl0[1024] = HT_VLC_TBL0_INIT;
static const uint16_t k_vlc_tbl1[1024] = HT_VLC_TBL1_INIT;

static inline uint32_t shl32(uint32_t v, uint32_t n) { return n >= 32 ? 0u : v << n; }
static inline uint64_t shl64(uint64_t v, uint32_t n) { return n >= 64 ? 0u : v << n; }

''' + ht[i_e:].replace("orc_ht_encode", "gen_ht_encode"))

# ---------- forward transforms ----------
dwt = rd("oracle/orc_dwt.c")
dwt, f_deint_i = cut(dwt, "/* deinterleave / interleave dwt.go:265-306 */\nstatic void deinterleave_i", "static void interleave_i")
dwt = dwt.replace("static void interleave_i", "/* interleave dwt.go:287-306, 329-346 */\nstatic void interleave_i", 1)
dwt, f_deint_f = cut(dwt, "static void deinterleave_f", "static void interleave_f")
dwt, f_fwd53 = cut(dwt, "static void fwd53_t", "static void inv53_t")
dwt, f_fwd97 = cut(dwt, "static void fwd97_t", "static void inv97_t")
dwt = dwt.replace("void orc_fwd53(int32_t *d, int n) { if (n < 2) return; int32_t *t = malloc(sizeof(int32_t) * (size_t)n); fwd53_t(d, n, t); free(t); }\n", "")
dwt = dwt.replace("void orc_fwd97(double *d, int n) { if (n < 2) return; double *t = malloc(sizeof(double) * (size_t)n); fwd97_t(d, n, t); free(t); }\n", "")
dwt, f_fwd2d53 = cut(dwt, "void orc_fwd2d53", "void orc_inv2d53")
dwt, f_fwd2d97 = cut(dwt, "void orc_fwd2d97", "void orc_inv2d97")
dwt, f_dec53 = cut(dwt, "void orc_decompose53", "void orc_reconstruct53")
dwt, f_dec97 = cut(dwt, "void orc_decompose97", "void orc_reconstruct97")
dwt, f_quant = cut(dwt, "void orc_quantize", "void orc_dequantize")
wr("oracle/orc_dwt.c", dwt)
tail = rd("oracle/orc_tail.c")
tail, f_frct = cut(tail, "void orc_fwd_rct", "void orc_inv_rct")
tail, f_fict = cut(tail, "void orc_fwd_ict", "void orc_inv_ict")
tail, f_dcf = cut(tail, "void orc_dc_shift_forward", "void orc_dc_shift_inverse")
wr("oracle/orc_tail.c", tail)
consts = dwt[dwt.index("static const double kAlpha"):dwt.index("/* interleave dwt.go")]
gen_fwd = ('''/*
 * gen_fwd.c -- restatement of the reference ENCODER-side transforms: Forward53/Forward97
 * (internal/dwt/dwt.go:73-118, 161-210), Forward2D53/97 (dwt.go:356-407, 432-451),
 * DecomposeMultiLevel53/97 (dwt.go:524-531, 551-558; dense-prefix layout), Quantize (dwt.go:500-511),
 * ForwardRCT/ForwardICT (internal/mct/mct.go:14-38) and DCLevelShiftForward (mct.go:96-101).
 * Part of datagen/ (synthetic-input generator).  float64 without FMA (-ffp-contract=off), int32 wraps.
 */
#include "datagen.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define WADD(a, b) ((int32_t)((uint32_t)(a) + (uint32_t)(b)))
#define WSUB(a, b) ((int32_t)((uint32_t)(a) - (uint32_t)(b)))
#define WMUL(a, b) ((int32_t)((uint32_t)(a) * (uint32_t)(b)))

''' + consts + f_deint_i.replace("/* deinterleave / interleave dwt.go:265-306 */", "/* deinterleave dwt.go:265-284, 309-326 */")
    + f_deint_f + f_fwd53 + f_fwd97 +
    '''void gen_fwd53(int32_t *d, int n) { if (n < 2) return; int32_t *t = malloc(sizeof(int32_t) * (size_t)n); fwd53_t(d, n, t); free(t); }
void gen_fwd97(double *d, int n) { if (n < 2) return; double *t = malloc(sizeof(double) * (size_t)n); fwd97_t(d, n, t); free(t); }

''' + f_fwd2d53 + f_fwd2d97 + f_dec53 + f_dec97 + f_quant + f_frct + f_fict + f_dcf)
gen_fwd = re.sub(r"\borc_(fwd2d53|fwd2d97|decompose53|decompose97|quantize|fwd_rct|fwd_ict|dc_shift_forward)\b", r"gen_\1", gen_fwd)
wr("datagen/gen_fwd.c", gen_fwd)

oh = rd("oracle/oracle.h")
for pat in [r"/\* Encode n \(ctx,bit\).*?\nint  orc_mq_encode\(.*?\);\n", r"/\* T1\.SetData \+ T1\.Encode:.*?\nint  orc_t1_encode\(.*?\);\n",
            r"int  orc_ht_encode\(.*?\);\n", r"void orc_fwd53\(.*?\n", r"void orc_fwd97\(.*?\n", r"void orc_fwd2d53\(.*?\n",
            r"void orc_fwd2d97\(.*?\n", r"void orc_decompose53\(.*?\n", r"void orc_decompose97\(.*?\n", r"void orc_quantize\(.*?\n",
            r"void orc_fwd_rct\(.*?\n", r"void orc_fwd_ict\(.*?\n", r"void orc_dc_shift_forward\(.*?\n"]:
    oh, n = re.subn(pat, "", oh, count=1, flags=re.S)
    assert n == 1, pat
wr("oracle/oracle.h", oh)
print("split done")
