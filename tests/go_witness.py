"""go_witness.py -- SECOND WITNESS for the REF-mode block decoders: a mechanical, statement-by-statement Python
transliteration of the reference's Go code, written from the Go source alone (not from oracle/*.c), so that the C
restatement the GPU kernels are checked against is itself checked by an independent reading of the same source:

    HTDecoder.Decode        internal/entropy/ht.go:93-150   (initMEL :153-195, initVLC :276-314, revRead :317-378,
                            revFetch/revAdvance :381-396, initMagSgn :399-429, frwdRead :432-501, frwdFetch/Advance :504-519,
                            decodeCleanup :583-713, decodeInitUVLC :716-805, decodeNonInitUVLC :808-864)
    T1.Decode               internal/entropy/t1.go:1261-1410 (flag helpers :307-345, contexts :349-479, :1087-1092,
                            canUseRunLength :1195-1208) with MQDecoder internal/entropy/mqc.go:370-497
    lutZCCtx                internal/entropy/t1_luts.go:32-110

Go semantics are spelled out where Python differs: fixed-width wrap-around (u32 / u64 masks), shifts by >= the operand
width give 0, uint32 counters that go "negative" wrap.  Test infrastructure only (tests/test_second_witness.py).
mqStates below was extracted mechanically from mqc.go:21-116 (Qe, MPS, NMPS, NLPS); the two CxtVLC tables are read from
oracle/ht_vlc_tables.inc, whose equality with ht_luts.go:18-281 tools/check_tables_vs_reference.py establishes."""
import os
import re

M32, M64 = 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF

MQ_STATES = [
    (0x5601, 0, 2, 3), (0x5601, 1, 3, 2), (0x3401, 0, 4, 12), (0x3401, 1, 5, 13),
    (0x1801, 0, 6, 18), (0x1801, 1, 7, 19), (0x0AC1, 0, 8, 24), (0x0AC1, 1, 9, 25),
    (0x0521, 0, 10, 58), (0x0521, 1, 11, 59), (0x0221, 0, 76, 66), (0x0221, 1, 77, 67),
    (0x5601, 0, 14, 13), (0x5601, 1, 15, 12), (0x5401, 0, 16, 28), (0x5401, 1, 17, 29),
    (0x4801, 0, 18, 28), (0x4801, 1, 19, 29), (0x3801, 0, 20, 28), (0x3801, 1, 21, 29),
    (0x3001, 0, 22, 34), (0x3001, 1, 23, 35), (0x2401, 0, 24, 36), (0x2401, 1, 25, 37),
    (0x1C01, 0, 26, 40), (0x1C01, 1, 27, 41), (0x1601, 0, 58, 42), (0x1601, 1, 59, 43),
    (0x5601, 0, 30, 29), (0x5601, 1, 31, 28), (0x5401, 0, 32, 28), (0x5401, 1, 33, 29),
    (0x5101, 0, 34, 30), (0x5101, 1, 35, 31), (0x4801, 0, 36, 32), (0x4801, 1, 37, 33),
    (0x3801, 0, 38, 34), (0x3801, 1, 39, 35), (0x3401, 0, 40, 36), (0x3401, 1, 41, 37),
    (0x3001, 0, 42, 38), (0x3001, 1, 43, 39), (0x2801, 0, 44, 38), (0x2801, 1, 45, 39),
    (0x2401, 0, 46, 40), (0x2401, 1, 47, 41), (0x2201, 0, 48, 42), (0x2201, 1, 49, 43),
    (0x1C01, 0, 50, 44), (0x1C01, 1, 51, 45), (0x1801, 0, 52, 46), (0x1801, 1, 53, 47),
    (0x1601, 0, 54, 48), (0x1601, 1, 55, 49), (0x1401, 0, 56, 50), (0x1401, 1, 57, 51),
    (0x1201, 0, 58, 52), (0x1201, 1, 59, 53), (0x1101, 0, 60, 54), (0x1101, 1, 61, 55),
    (0x0AC1, 0, 62, 56), (0x0AC1, 1, 63, 57), (0x09C1, 0, 64, 58), (0x09C1, 1, 65, 59),
    (0x08A1, 0, 66, 60), (0x08A1, 1, 67, 61), (0x0521, 0, 68, 62), (0x0521, 1, 69, 63),
    (0x0441, 0, 70, 64), (0x0441, 1, 71, 65), (0x02A1, 0, 72, 66), (0x02A1, 1, 73, 67),
    (0x0221, 0, 74, 68), (0x0221, 1, 75, 69), (0x0141, 0, 76, 70), (0x0141, 1, 77, 71),
    (0x0111, 0, 78, 72), (0x0111, 1, 79, 73), (0x0085, 0, 80, 74), (0x0085, 1, 81, 75),
    (0x0049, 0, 82, 76), (0x0049, 1, 83, 77), (0x0025, 0, 84, 78), (0x0025, 1, 85, 79),
    (0x0015, 0, 86, 80), (0x0015, 1, 87, 81), (0x0009, 0, 88, 82), (0x0009, 1, 89, 83),
    (0x0005, 0, 90, 84), (0x0005, 1, 91, 85), (0x0001, 0, 90, 86), (0x0001, 1, 91, 87),
    (0x5601, 0, 92, 92), (0x5601, 1, 93, 93)]
MQ_QE = [s[0] for s in MQ_STATES]
MQ_NMPS = [s[2] for s in MQ_STATES]
MQ_NLPS = [s[3] for s in MQ_STATES]
(CTX_ZC0, CTX_SC0, CTX_MAG0, CTX_RL, CTX_UNI, NUM_CONTEXTS) = (0, 9, 14, 17, 18, 19)       # mqc.go:135-166
T1_SIG, T1_VISIT, T1_REFINE, T1_SIGN_NEG, T1_SIG_N, T1_SIG_S, T1_SIG_E, T1_SIG_W = 1, 2, 4, 8, 16, 32, 64, 128   # t1.go:72-91
BAND_LL, BAND_HL, BAND_LH, BAND_HH = 0, 1, 2, 3                                           # t1.go:125-130


def shl32(v, n):
    return (v << n) & M32 if n < 32 else 0


def shl64(v, n):
    return (v << n) & M64 if n < 64 else 0


def shr64(v, n):
    return v >> n if n < 64 else 0


# ---- t1_luts.go:32-110 ----------------------------------------------------------------------------------------------
def _make_zc_lut():
    lut = [0] * 1024
    for band in range(4):
        for packed in range(256):
            w, e, n, s = packed & 1, (packed >> 1) & 1, (packed >> 2) & 1, (packed >> 3) & 1
            nw, ne, sw, se = (packed >> 4) & 1, (packed >> 5) & 1, (packed >> 6) & 1, (packed >> 7) & 1
            h, v, d = w + e, n + s, nw + ne + sw + se
            ctx = 0
            if band in (BAND_HL, BAND_LL, BAND_LH):
                if band == BAND_HL:
                    h, v = v, h
                if h == 2:
                    ctx = 8
                elif h == 1:
                    ctx = 7 if v >= 1 else (6 if d >= 1 else 5)
                elif v == 2:
                    ctx = 4
                elif v == 1:
                    ctx = 3 if d >= 1 else 2
                elif d >= 2:
                    ctx = 1
                else:
                    ctx = 0
            else:
                hv = h + v
                if hv >= 3:
                    ctx = 8
                elif hv == 2:
                    ctx = 7 if d >= 2 else (6 if d >= 1 else 5)
                elif hv == 1:
                    ctx = 4 if d >= 2 else 3
                else:
                    ctx = 2 if d >= 2 else (1 if d >= 1 else 0)
            lut[band * 256 + packed] = ctx
    return lut


LUT_ZC = _make_zc_lut()


# ---- mqc.go:352-497 -------------------------------------------------------------------------------------------------
class MQDecoder:
    def __init__(self, data):                       # NewMQDecoder :370-399
        self.A, self.C, self.CT, self.data, self.bp = 0x8000, 0, 0, data, -1
        self.contexts = [0] * NUM_CONTEXTS
        self.contexts[CTX_UNI] = 92
        if len(data) == 0:
            self.C = 0xFF << 16
        else:
            self.bp = 0
            self.C = data[0] << 16
        self.byte_in()
        self.C = shl32(self.C, 7)
        self.CT = (self.CT - 7) & M32
        self.A = 0x8000

    def byte_in(self):                              # :402-439
        if self.bp < 0:
            self.bp = 0
        if self.bp >= len(self.data):
            self.C = (self.C + 0xFF00) & M32
            self.CT = 8
            return
        nxt = self.data[self.bp + 1] if self.bp + 1 < len(self.data) else 0xFF
        if self.data[self.bp] == 0xFF:
            if nxt > 0x8F:
                self.C = (self.C + 0xFF00) & M32
                self.CT = 8
            else:
                self.bp += 1
                self.C = (self.C + (nxt << 9)) & M32
                self.CT = 7
        else:
            self.bp += 1
            self.C = (self.C + (nxt << 8)) & M32
            self.CT = 8

    def renorm(self):                               # :488-497
        while (self.A & 0x8000) == 0:
            if self.CT == 0:
                self.byte_in()
            self.A = shl32(self.A, 1)
            self.C = shl32(self.C, 1)
            self.CT = (self.CT - 1) & M32

    def decode(self, ctx):                          # :443-485
        st = self.contexts[ctx]
        qe = MQ_QE[st]
        mps = st & 1
        self.A = (self.A - qe) & M32
        if (self.C >> 16) < qe:
            if self.A < qe:
                self.A = qe
                decision = mps
                self.contexts[ctx] = MQ_NMPS[st]
            else:
                self.A = qe
                decision = 1 - mps
                self.contexts[ctx] = MQ_NLPS[st]
            self.renorm()
            return decision
        self.C = (self.C - shl32(qe, 16)) & M32
        if (self.A & 0x8000) == 0:
            if self.A < qe:
                decision = 1 - mps
                self.contexts[ctx] = MQ_NLPS[st]
            else:
                decision = mps
                self.contexts[ctx] = MQ_NMPS[st]
            self.renorm()
            return decision
        return mps


# ---- t1.go -----------------------------------------------------------------------------------------------------------
class T1:
    def __init__(self, width, height):              # NewT1
        self.width, self.height = width, height
        self.data = [0] * (width * height)
        self.flags = [0] * ((width + 2) * (height + 2))

    def flag_index(self, x, y):                     # :307-309
        return (y + 1) * (self.width + 2) + (x + 1)

    def set_flag(self, x, y, f):
        self.flags[self.flag_index(x, y)] |= f

    def has_flag(self, x, y, f):
        return (self.flags[self.flag_index(x, y)] & f) != 0

    def clear_flag(self, x, y, f):
        self.flags[self.flag_index(x, y)] &= ~f & 0xFF

    def update_neighbor_flags(self, x, y):          # :328-345
        idx, stride = self.flag_index(x, y), self.width + 2
        if y > 0:
            self.flags[idx - stride] |= T1_SIG_S
        if y < self.height - 1:
            self.flags[idx + stride] |= T1_SIG_N
        if x > 0:
            self.flags[idx - 1] |= T1_SIG_E
        if x < self.width - 1:
            self.flags[idx + 1] |= T1_SIG_W

    def get_zc_context(self, x, y, band):           # :349-383
        idx, stride, f = self.flag_index(x, y), self.width + 2, self.flags
        packed = 0
        if f[idx - 1] & T1_SIG:
            packed |= 0x01
        if f[idx + 1] & T1_SIG:
            packed |= 0x02
        if f[idx - stride] & T1_SIG:
            packed |= 0x04
        if f[idx + stride] & T1_SIG:
            packed |= 0x08
        if f[idx - stride - 1] & T1_SIG:
            packed |= 0x10
        if f[idx - stride + 1] & T1_SIG:
            packed |= 0x20
        if f[idx + stride - 1] & T1_SIG:
            packed |= 0x40
        if f[idx + stride + 1] & T1_SIG:
            packed |= 0x80
        return LUT_ZC[band * 256 + packed]

    def get_sc_context(self, x, y):                 # :387-460
        idx, stride, f = self.flag_index(x, y), self.width + 2, self.flags
        hc = 0
        if f[idx - 1] & T1_SIG:
            hc += -1 if f[idx - 1] & T1_SIGN_NEG else 1
        if f[idx + 1] & T1_SIG:
            hc += -1 if f[idx + 1] & T1_SIGN_NEG else 1
        vc = 0
        if f[idx - stride] & T1_SIG:
            vc += -1 if f[idx - stride] & T1_SIGN_NEG else 1
        if f[idx + stride] & T1_SIG:
            vc += -1 if f[idx + stride] & T1_SIGN_NEG else 1
        pred = 0
        if hc < 0:
            pred = 1
            hc = -hc
        if hc == 0:
            if vc < 0:
                pred = 1
                vc = -vc
        ctx = CTX_SC0
        if hc == 1:
            if vc == 1:
                ctx = CTX_SC0 + 4
            elif vc == 0:
                ctx = CTX_SC0 + 2
            else:
                ctx = CTX_SC0 + 1
        elif hc == 0:
            if vc == 1:
                ctx = CTX_SC0 + 1
            elif vc == 0:
                ctx = CTX_SC0
        elif hc == 2:
            ctx = CTX_SC0 + 3
        return ctx, pred

    def get_mr_context(self, x, y):                 # :463-479
        idx, stride, f = self.flag_index(x, y), self.width + 2, self.flags
        if (f[idx] & T1_REFINE) == 0:
            nb = (f[idx - 1] | f[idx + 1] | f[idx - stride] | f[idx + stride] | f[idx - stride - 1] | f[idx - stride + 1] |
                  f[idx + stride - 1] | f[idx + stride + 1]) & T1_SIG
            return CTX_MAG0 + 1 if nb else CTX_MAG0
        return CTX_MAG0 + 2

    def has_significant_neighbor(self, x, y):       # :1087-1092
        idx, stride, f = self.flag_index(x, y), self.width + 2, self.flags
        return ((f[idx - 1] | f[idx + 1] | f[idx - stride] | f[idx + stride] | f[idx - stride - 1] | f[idx - stride + 1] |
                 f[idx + stride - 1] | f[idx + stride + 1]) & T1_SIG) != 0

    def can_use_run_length(self, x, y):             # :1195-1208
        if y + 4 > self.height:
            return False
        for yy in range(y, y + 4):
            if self.has_flag(x, yy, T1_SIG | T1_VISIT):
                return False
            if self.has_significant_neighbor(x, yy):
                return False
        return True

    def decode_sign(self, x, y):                    # :1322-1328
        ctx, pred = self.get_sc_context(x, y)
        if self.mq.decode(ctx) ^ pred:
            self.set_flag(x, y, T1_SIGN_NEG)

    def decode(self, data, num_bps, band):          # :1261-1292
        self.band = band
        self.mq = MQDecoder(data)
        self.data = [0] * len(self.data)
        self.flags = [0] * len(self.flags)
        for bp in range(num_bps - 1, -1, -1):
            self.significance_pass(bp)
            self.refinement_pass(bp)
            self.cleanup_pass(bp)
        out = []
        for i, v in enumerate(self.data):
            neg = self.flags[self.flag_index(i % self.width, i // self.width)] & T1_SIGN_NEG
            out.append(_i32(-v) if neg else v)
        return out

    def _became_significant(self, x, y, bit):
        self.data[y * self.width + x] = bit
        self.decode_sign(x, y)
        self.set_flag(x, y, T1_SIG)
        self.update_neighbor_flags(x, y)

    def significance_pass(self, bp):                # :1295-1319 (raster order)
        bit = _i32(1 << bp)
        for y in range(self.height):
            for x in range(self.width):
                if self.has_flag(x, y, T1_SIG):
                    continue
                if not self.has_significant_neighbor(x, y):
                    continue
                if self.mq.decode(self.get_zc_context(x, y, self.band)):
                    self._became_significant(x, y, bit)
                self.set_flag(x, y, T1_VISIT)

    def refinement_pass(self, bp):                  # :1331-1347 (raster order)
        bit = _i32(1 << bp)
        for y in range(self.height):
            for x in range(self.width):
                if not self.has_flag(x, y, T1_SIG) or self.has_flag(x, y, T1_VISIT):
                    continue
                if self.mq.decode(self.get_mr_context(x, y)):
                    self.data[y * self.width + x] = _i32(self.data[y * self.width + x] | bit)
                self.set_flag(x, y, T1_REFINE)

    def cleanup_pass(self, bp):                     # :1350-1381
        bit = _i32(1 << bp)
        for y in range(0, self.height, 4):
            for x in range(self.width):
                if self.can_use_run_length(x, y):
                    self.decode_run_length(x, y, bit)
                    continue
                yy = y
                while yy < y + 4 and yy < self.height:
                    if self.has_flag(x, yy, T1_VISIT):
                        self.clear_flag(x, yy, T1_VISIT)
                    elif not self.has_flag(x, yy, T1_SIG):
                        if self.mq.decode(self.get_zc_context(x, yy, self.band)):
                            self._became_significant(x, yy, bit)
                    yy += 1

    def decode_run_length(self, x, y, bit):         # :1384-1410
        if self.mq.decode(CTX_RL) == 0:
            return
        pos = self.mq.decode(CTX_UNI) << 1
        pos |= self.mq.decode(CTX_UNI)
        self._became_significant(x, y + pos, bit)
        i = pos + 1
        while i < 4 and y + i < self.height:
            if self.mq.decode(self.get_zc_context(x, y + i, self.band)):
                self._became_significant(x, y + i, bit)
            i += 1


def _i32(v):
    v &= M32
    return v - (1 << 32) if v & 0x80000000 else v


def t1_decode(data, w, h, num_bps, band):
    return T1(w, h).decode(bytes(data), num_bps, band)


# ---- ht.go -----------------------------------------------------------------------------------------------------------
def _load_vlc_tables():
    txt = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "ht_vlc_tables.inc")).read()
    out = []
    for name in ("HT_VLC_TBL0_INIT", "HT_VLC_TBL1_INIT"):
        m = re.search(r"#define\s+" + name + r"\s*\\?\s*\{(.*?)\}", txt, re.S)
        vals = [int(t, 0) for t in re.findall(r"0x[0-9A-Fa-f]+|\d+", m.group(1))]
        assert len(vals) == 1024, (name, len(vals))
        out.append(vals)
    return out


VLC_TBL0, VLC_TBL1 = _load_vlc_tables()
UVLC_DEC = [3 | (5 << 2) | (5 << 5), 1 | (0 << 2) | (1 << 5), 2 | (0 << 2) | (2 << 5), 1 | (0 << 2) | (1 << 5),
            3 | (1 << 2) | (3 << 5), 1 | (0 << 2) | (1 << 5), 2 | (0 << 2) | (2 << 5), 1 | (0 << 2) | (1 << 5)]   # ht.go:718-727


class _Rev:          # revBitstream ht.go:55-64: bits is uint32
    pass


class _Fwd:          # frwdBitstream ht.go:66-75
    pass


class HTDecoder:
    def __init__(self, width, height):              # NewHTDecoder :78-88
        self.width, self.height = width, height
        self.data = [0] * (width * height)
        qc = (width + 3) // 4
        self.sigma1, self.sigma2, self.line_state = [0] * (qc + 1), [0] * (qc + 1), [0] * (qc + 1)

    def init_mel(self, data, lcup, scup):           # :153-195 (only the boolean result is observable)
        pos, bits, tmp, unstuff, size = lcup - scup, 0, 0, False, scup - 1
        num = 4 - (pos & 3)
        if num > 4:
            num = 4
        i = 0
        while i < num and size > 0:
            if unstuff and pos < len(data) and data[pos] > 0x8F:
                return False
            if size > 0 and pos < len(data):
                b = data[pos]
                pos += 1
                size -= 1
            else:
                b = 0xFF
            if size == 1:
                b |= 0x0F
            d_bits = 7 if unstuff else 8
            tmp = ((tmp << d_bits) & M64) | b
            bits += d_bits
            unstuff = b == 0xFF
            i += 1
        return True

    def init_vlc(self, data, lcup, scup):           # :276-314
        v = self.vlc = _Rev()
        v.data, v.pos, v.size, v.tmp, v.bits, v.unstuff = data, lcup - 2, scup - 2, 0, 0, False
        if 0 <= v.pos < len(data):
            b = data[v.pos]
            v.pos -= 1
            v.tmp = b >> 4
            v.bits = (4 - ((v.tmp & 7) >> 2)) & M32
            v.unstuff = (b | 0x0F) > 0x8F
        num = 1 + (v.pos & 3)
        if num > v.size:
            num = v.size
        for _ in range(max(num, 0)):
            b = 0
            if 0 <= v.pos < len(data):
                b = data[v.pos]
                v.pos -= 1
            d_bits = 7 if (v.unstuff and (b & 0x7F) == 0x7F) else 8
            v.tmp = (v.tmp | shl64(b, v.bits)) & M64
            v.bits = (v.bits + d_bits) & M32
            v.unstuff = b > 0x8F
        v.size -= num
        self.rev_read(v)

    def rev_read(self, v):                          # :317-378
        if v.bits > 32:
            return
        val = 0
        if v.size > 3:
            p = v.pos - 3
            if p >= 0 and p + 3 < len(v.data):
                val = v.data[p] | (v.data[p + 1] << 8) | (v.data[p + 2] << 16) | (v.data[p + 3] << 24)
            v.pos -= 4
            v.size -= 4
        elif v.size > 0:
            i = 24
            while v.size > 0:
                if 0 <= v.pos < len(v.data):
                    val |= shl32(v.data[v.pos], i) if i >= 0 else 0
                    v.pos -= 1
                v.size -= 1
                i -= 8
        tmp = val >> 24
        bits = 7 if (v.unstuff and ((val >> 24) & 0x7F) == 0x7F) else 8
        unstuff = (val >> 24) > 0x8F
        tmp |= shl32((val >> 16) & 0xFF, bits)
        bits += 7 if (unstuff and ((val >> 16) & 0x7F) == 0x7F) else 8
        unstuff = ((val >> 16) & 0xFF) > 0x8F
        tmp |= shl32((val >> 8) & 0xFF, bits)
        bits += 7 if (unstuff and ((val >> 8) & 0x7F) == 0x7F) else 8
        unstuff = ((val >> 8) & 0xFF) > 0x8F
        tmp |= shl32(val & 0xFF, bits)
        bits += 7 if (unstuff and (val & 0x7F) == 0x7F) else 8
        v.unstuff = (val & 0xFF) > 0x8F
        v.tmp = (v.tmp | shl64(tmp, v.bits)) & M64
        v.bits = (v.bits + bits) & M32

    def rev_fetch(self, v):                         # :381-389
        if v.bits < 32:
            self.rev_read(v)
            if v.bits < 32:
                self.rev_read(v)
        return v.tmp & M32

    def rev_advance(self, v, n):                    # :392-396
        v.tmp = shr64(v.tmp, n)
        v.bits = (v.bits - n) & M32

    def init_magsgn(self, data, size):              # :399-429
        f = self.ms = _Fwd()
        f.data, f.pos, f.size, f.tmp, f.bits, f.unstuff, f.x = data, 0, size, 0, 0, False, 0xFF
        num = 4 - (f.pos & 3)
        for _ in range(num):
            if f.size > 0 and f.pos < len(data):
                b = data[f.pos]
                f.pos += 1
                f.size -= 1
            else:
                b = f.x & 0xFF
            d_bits = 7 if f.unstuff else 8
            f.tmp = (f.tmp | shl64(b, f.bits)) & M64
            f.bits = (f.bits + d_bits) & M32
            f.unstuff = b == 0xFF
        self.frwd_read(f)

    def frwd_read(self, f):                         # :432-501
        if f.bits > 32:
            return
        val = 0
        if f.size > 3:
            if f.pos + 3 < len(f.data):
                val = f.data[f.pos] | (f.data[f.pos + 1] << 8) | (f.data[f.pos + 2] << 16) | (f.data[f.pos + 3] << 24)
            f.pos += 4
            f.size -= 4
        elif f.size > 0:
            if f.x != 0:
                val = 0xFFFFFFFF
            i = 0
            while f.size > 0:
                if f.pos < len(f.data):
                    b = f.data[f.pos]
                    m = ~shl32(0xFF, i) & M32
                    val = (val & m) | shl32(b, i)
                    f.pos += 1
                f.size -= 1
                i += 8
        else:
            if f.x != 0:
                val = 0xFFFFFFFF
        bits = 7 if f.unstuff else 8
        t = val & 0xFF
        unstuff = (val & 0xFF) == 0xFF
        t |= shl32((val >> 8) & 0xFF, bits)
        bits += 7 if unstuff else 8
        unstuff = ((val >> 8) & 0xFF) == 0xFF
        t |= shl32((val >> 16) & 0xFF, bits)
        bits += 7 if unstuff else 8
        unstuff = ((val >> 16) & 0xFF) == 0xFF
        t |= shl32((val >> 24) & 0xFF, bits)
        bits += 7 if unstuff else 8
        f.unstuff = ((val >> 24) & 0xFF) == 0xFF
        f.tmp = (f.tmp | shl64(t, f.bits)) & M64
        f.bits = (f.bits + bits) & M32

    def frwd_fetch(self, f):                        # :504-512
        if f.bits < 32:
            self.frwd_read(f)
            if f.bits < 32:
                self.frwd_read(f)
        return f.tmp & M32

    def frwd_advance(self, f, n):                   # :515-519
        f.tmp = shr64(f.tmp, n)
        f.bits = (f.bits - n) & M32

    def decode(self, data):                         # :93-150 (numBitplanes, bandType are never read)
        data = bytes(data)
        if len(data) < 2:
            return [0] * len(self.data)
        scup = data[-1] + ((data[-2] & 0x0F) << 8)
        if scup < 2 or scup > len(data):
            return [0] * len(self.data)
        lcup = len(data)
        if not self.init_mel(data, lcup, scup):
            return [0] * len(self.data)
        self.init_vlc(data, lcup, scup)
        self.init_magsgn(data, lcup - scup)
        self.decode_cleanup()
        return self.data

    def _uvlc(self, vlc, mode, initial):            # decodeInitUVLC :716-805 / decodeNonInitUVLC :808-864
        u = [0, 0]
        consumed = 0
        if mode == 0:
            u = [1, 1]
        elif mode <= 2:
            t = UVLC_DEC[vlc & 7]
            pl = t & 3
            vlc >>= pl
            consumed += pl
            sl = (t >> 2) & 7
            consumed += sl
            val = (t >> 5) + (vlc & ((1 << sl) - 1))
            u = [val + 1, 1] if mode == 1 else [1, val + 1]
        elif mode == 3:
            t1 = UVLC_DEC[vlc & 7]
            p1 = t1 & 3
            vlc >>= p1
            consumed += p1
            if initial and p1 > 2:
                u[1] = (vlc & 1) + 2
                consumed += 1
                vlc >>= 1
                sl = (t1 >> 2) & 7
                consumed += sl
                u[0] = (t1 >> 5) + (vlc & ((1 << sl) - 1)) + 1
            else:
                t2 = UVLC_DEC[vlc & 7]
                p2 = t2 & 3
                vlc >>= p2
                consumed += p2
                s1 = (t1 >> 2) & 7
                consumed += s1
                u[0] = (t1 >> 5) + (vlc & ((1 << s1) - 1)) + 1
                vlc >>= s1
                s2 = (t2 >> 2) & 7
                consumed += s2
                u[1] = (t2 >> 5) + (vlc & ((1 << s2) - 1)) + 1
        return consumed, u

    def _sample(self, emb):                         # :664-684
        mag_val = self.frwd_fetch(self.ms)
        mag = ((mag_val & ((shl32(1, emb) - 1) & M32)) + shl32(1, (emb - 1) & M32)) & M32
        self.frwd_advance(self.ms, emb)
        sign = self.frwd_fetch(self.ms) & 1
        self.frwd_advance(self.ms, 1)
        mag = _i32(mag)
        return _i32(-mag) if sign else mag

    def decode_cleanup(self):                       # :583-713
        width, height = self.width, self.height
        quad_cols = (width + 3) // 4
        for y in range(0, height, 4):
            initial = y == 0
            for qx in range(0, quad_cols, 2):
                vlc_val = self.rev_fetch(self.vlc)
                context = 0
                if initial:
                    if qx > 0:
                        context = self.sigma1[qx - 1] >> 4
                else:
                    context = (self.sigma1[qx] >> 4) | (self.line_state[qx] >> 4)
                tbl = VLC_TBL0 if initial else VLC_TBL1
                qinf = tbl[(context << 7) | (vlc_val & 0x7F)]
                vlc_len, rho, u_off1 = qinf & 0x0F, (qinf >> 4) & 0x0F, (qinf >> 3) & 1
                self.rev_advance(self.vlc, vlc_len)
                vlc_val = self.rev_fetch(self.vlc)
                context2 = ((rho >> 2) | (self.sigma1[qx + 1] >> 4)) & 0xFF
                qinf2 = tbl[(context2 << 7) | (vlc_val & 0x7F)]
                vlc_len2, rho2, u_off2 = qinf2 & 0x0F, (qinf2 >> 4) & 0x0F, (qinf2 >> 3) & 1
                self.rev_advance(self.vlc, vlc_len2)
                self.sigma1[qx], self.sigma1[qx + 1] = rho & 0xFF, rho2 & 0xFF
                mode = (u_off1 << 1) | u_off2
                if mode > 0:
                    vlc_val = self.rev_fetch(self.vlc)
                    consumed, u = self._uvlc(vlc_val, mode, initial)
                    self.rev_advance(self.vlc, consumed)
                else:
                    u = [1, 1]
                for q, r in ((qx, rho), (qx + 1, rho2)):
                    i = 0
                    while i < 4 and q * 4 + i < width:
                        if r & (1 << i):
                            v = self._sample(u[0] if q == qx else u[1])
                            idx = y * width + q * 4 + i
                            if idx < len(self.data):
                                self.data[idx] = v
                        i += 1


def ht_decode(data, w, h):
    return HTDecoder(w, h).decode(data)
