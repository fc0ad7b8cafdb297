"""datagen.codestream -- minimal ISO/IEC 15444-1 / -15 codestream WRITER and tier-2 PARSER (host-side test
harness; in the product the reference's Go packages internal/codestream + a real tier-2 do this job).

Writer: samples -> DC shift -> RCT -> Mallat multi-level 5-3 DWT per tile -> 64x64 code blocks -> block coder
(conformant HT cleanup encoder of gen_iso_ht.c) -> packets (LRCP, one layer, one precinct per resolution,
tag trees, Lblock lengths) -> SOC SIZ [CAP] COD QCD (SOT SOD packets)* EOC.  Its purpose is to let OpenJPEG
(Pillow) decode our HT encoder's output -- the ISO conformance pin -- and to build ISO-mode job tables.

Parser: main header + tile-parts + packet headers of a codestream (as produced by OpenJPEG or by the writer)
-> per code block: concatenated segment bytes, number of coding passes, missing MSBs, band geometry.
Supported: LRCP/RLCP, any number of layers, default (maximal) precincts, no SOP/EPH, no PPM/PPT, one or more
tile-parts per tile in order, no sub-sampling, tile/image origins at 0.
"""
import struct

import numpy as np

from . import fwd2d53, iso_ht_encode, iso_ht_encode_passes

SOC, SIZ, CAP, COD, QCD, SOT, SOD, EOC, COM = 0xFF4F, 0xFF51, 0xFF50, 0xFF52, 0xFF5C, 0xFF90, 0xFF93, 0xFFD9, 0xFF64


def cdiv(a, b):
    return -(-a // b)


# ------------------------------------------------------------------------------------------------ geometry
def band_list(nlevels):
    """[(res, band, level)] in codestream order: band 0 LL, 1 HL, 2 LH, 3 HH; level = decomposition level"""
    out = [(0, 0, nlevels)]
    for r in range(1, nlevels + 1):
        lvl = nlevels - r + 1
        out += [(r, 1, lvl), (r, 2, lvl), (r, 3, lvl)]
    return out


def band_rect(tx0, ty0, tx1, ty1, band, lvl):
    """band bounds in band coordinates (B.15): HL/HH are offset in x, LH/HH in y"""
    if band == 0:
        s = 1 << lvl
        return cdiv(tx0, s), cdiv(ty0, s), cdiv(tx1, s), cdiv(ty1, s)
    xo, yo = (band & 1), (band >> 1)
    s, hs = 1 << lvl, 1 << (lvl - 1)
    return (cdiv(tx0 - hs * xo, s), cdiv(ty0 - hs * yo, s), cdiv(tx1 - hs * xo, s), cdiv(ty1 - hs * yo, s))


def band_origin_in_plane(tx0, ty0, tx1, ty1, band, lvl):
    """top-left of the band inside the tile-component's Mallat plane (LL_lvl top-left, HL right, LH below)"""
    s = 1 << lvl
    lw = cdiv(tx1, s) - cdiv(tx0, s)     # width of LL at this level
    lh = cdiv(ty1, s) - cdiv(ty0, s)
    return (lw if band & 1 else 0), (lh if band >> 1 else 0)


def cblk_grid(bx0, by0, bx1, by1, cbw, cbh):
    """code blocks of a band: cells of the cbw x cbh grid anchored at 0 intersected with the band; raster order.
    -> list of (x0, y0, x1, y1) in band coordinates, plus grid dims"""
    if bx1 <= bx0 or by1 <= by0:
        return [], 0, 0
    gx0, gy0 = bx0 // cbw, by0 // cbh
    gx1, gy1 = cdiv(bx1, cbw), cdiv(by1, cbh)
    blocks = []
    for gy in range(gy0, gy1):
        for gx in range(gx0, gx1):
            blocks.append((max(bx0, gx * cbw), max(by0, gy * cbh), min(bx1, (gx + 1) * cbw), min(by1, (gy + 1) * cbh)))
    return blocks, gx1 - gx0, gy1 - gy0


# ------------------------------------------------------------------------------------------------ bit I/O
class BitWriter:
    """packet-header bit writer: MSB first, a byte after 0xFF carries 7 bits (B.10.1)"""

    def __init__(self):
        self.out = bytearray()
        self.cur, self.free = 0, 8

    def put(self, bit):
        self.free -= 1
        self.cur |= (bit & 1) << self.free
        if self.free == 0:
            self._flush()

    def _flush(self):
        self.out.append(self.cur)
        self.free = 7 if self.cur == 0xFF else 8
        self.cur = 0

    def bits(self, v, n):
        for i in range(n - 1, -1, -1):
            self.put((v >> i) & 1)

    def finish(self):
        full = 7 if (self.out and self.out[-1] == 0xFF) else 8
        if self.free != full:                  # pending bits: pad the byte with zeros
            self._flush()
        if self.out and self.out[-1] == 0xFF:  # a header may not end in 0xFF: the stuffed byte follows
            self.out.append(0)
        return bytes(self.out)


class BitReader:
    def __init__(self, data, pos):
        self.d, self.pos = data, pos
        self.cur, self.left, self.last = 0, 0, 0

    def get(self):
        if self.left == 0:
            self.last, self.cur = self.cur if self.pos else 0, self.d[self.pos]
            self.left = 7 if self._prev_ff else 8
            self._prev_ff = self.cur == 0xFF
            self.pos += 1
        self.left -= 1
        return (self.cur >> self.left) & 1

    _prev_ff = False

    def bits(self, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | self.get()
        return v

    def align(self):
        """end of packet header: skip to the byte boundary; if the last byte was 0xFF one stuffed byte follows"""
        if self._prev_ff and self.left == 0:
            self.pos += 1
        elif self._prev_ff and self.left > 0:
            pass
        self.left = 0
        self._prev_ff = False
        return self.pos


class TagTree:
    def __init__(self, w, h):
        self.dims = []
        while True:
            self.dims.append((w, h))
            if w <= 1 and h <= 1:
                break
            w, h = cdiv(w, 2), cdiv(h, 2)
        self.val = [np.full((hh, ww), 1 << 30, np.int64) for ww, hh in self.dims]
        self.low = [np.zeros((hh, ww), np.int64) for ww, hh in self.dims]
        self.known = [np.zeros((hh, ww), bool) for ww, hh in self.dims]

    def set_values(self, leaves):
        self.val[0][:, :] = leaves
        for l in range(1, len(self.dims)):
            ww, hh = self.dims[l]
            for y in range(hh):
                for x in range(ww):
                    self.val[l][y, x] = self.val[l - 1][2 * y:2 * y + 2, 2 * x:2 * x + 2].min()

    def encode(self, bw, x, y, threshold):
        low = 0
        for l in range(len(self.dims) - 1, -1, -1):
            xx, yy = x >> l, y >> l
            if low > self.low[l][yy, xx]:
                self.low[l][yy, xx] = low
            else:
                low = int(self.low[l][yy, xx])
            while low < threshold:
                if low >= self.val[l][yy, xx]:
                    if not self.known[l][yy, xx]:
                        bw.put(1)
                        self.known[l][yy, xx] = True
                    break
                bw.put(0)
                low += 1
            self.low[l][yy, xx] = low

    def decode(self, br, x, y, threshold):
        """-> True when value(x, y) < threshold is established (value then in self.val)"""
        low = 0
        for l in range(len(self.dims) - 1, -1, -1):
            xx, yy = x >> l, y >> l
            if low > self.low[l][yy, xx]:
                self.low[l][yy, xx] = low
            else:
                low = int(self.low[l][yy, xx])
            while low < threshold and not self.known[l][yy, xx]:
                if br.get():
                    self.known[l][yy, xx] = True
                    self.val[l][yy, xx] = low
                else:
                    low += 1
            self.low[l][yy, xx] = low
            if self.known[l][yy, xx]:
                low = max(low, int(self.val[l][yy, xx])) if l else low
        return bool(self.known[0][y, x]) and self.val[0][y, x] < threshold


# ------------------------------------------------------------------------------------------------ transforms
def mallat53(plane, nlevels):
    """multi-level forward 5-3, ISO order (columns then rows), Mallat layout (LL recursively top-left); even origin"""
    out = plane.astype(np.int32).copy()
    h, w = out.shape
    for _ in range(nlevels):
        if w < 1 or h < 1:
            break
        # ISO/IEC 15444-1 F.4: vertical analysis first, then horizontal (the reference's Forward2D53 does rows
        # first -- with integer lifting the order changes results), so run it on the transpose
        sub = fwd2d53(np.ascontiguousarray(out[:h, :w].T), h, w).reshape(w, h).T
        out[:h, :w] = sub
        w, h = (w + 1) // 2, (h + 1) // 2
    return out


def forward_tile_iso(samples, prec, nlevels, mct):
    ncomp = samples.shape[0]
    c = [samples[i].astype(np.int32) - (1 << (prec - 1)) for i in range(ncomp)]
    if mct and ncomp >= 3:
        y = (c[0] + 2 * c[1] + c[2]) >> 2
        u, v = c[2] - c[1], c[0] - c[1]
        c[0], c[1], c[2] = y, u, v
    return [mallat53(p, nlevels) for p in c]


def fwd97_1d(x, axis):
    """one line direction of the forward 9-7 (Annex F.4.8, float64), interleaved in -> band order out (L first)"""
    x = np.moveaxis(np.asarray(x, np.float64), axis, 0).copy()
    n = x.shape[0]
    if n < 2:
        return np.moveaxis(x, 0, axis)
    al, be, ga, de, K = -1.586134342059924, -0.052980118572961, 0.882911075530934, 0.443506852043971, 1.230174104914001
    for par, c in ((1, al), (0, be), (1, ga), (0, de)):
        idx = np.arange(par, n, 2)
        l = np.where(idx - 1 >= 0, idx - 1, idx + 1)
        r = np.where(idx + 1 < n, idx + 1, idx - 1)
        x[idx] = x[idx] + c * (x[l] + x[r])
    out = np.concatenate([x[0::2] / K, x[1::2] * K], axis=0)
    return np.moveaxis(out, 0, axis)


def forward_tile_iso_97(samples, prec, nlevels, mct):
    """DC shift -> forward ICT -> multi-level forward 9-7 in Mallat layout (float64; vertical then horizontal)"""
    ncomp = samples.shape[0]
    c = [samples[i].astype(np.float64) - (1 << (prec - 1)) for i in range(ncomp)]
    if mct and ncomp >= 3:
        r, g, b = c[0], c[1], c[2]
        c[0] = 0.299 * r + 0.587 * g + 0.114 * b
        c[1] = -0.168736 * r - 0.331264 * g + 0.5 * b
        c[2] = 0.5 * r - 0.418688 * g - 0.081312 * b
    planes = []
    for p in c:
        out = p.copy()
        h, w = out.shape
        for _ in range(nlevels):
            if w < 1 or h < 1:
                break
            out[:h, :w] = fwd97_1d(fwd97_1d(out[:h, :w], 0), 1)
            w, h = (w + 1) // 2, (h + 1) // 2
        planes.append(out)
    return planes


# ------------------------------------------------------------------------------------------------ writer
def _put_npasses(bw, n):
    """number of coding passes codeword (B.10.6), n = 1..5"""
    if n == 1:
        bw.put(0)
    elif n == 2:
        bw.bits(2, 2)
    else:
        bw.bits(0b1100 + (n - 3), 4)


def write_htj2k(samples, prec, tile_w=None, tile_h=None, nlevels=5, mct=1, cb=64, guard=2, extra_eps=1, lossy_step=None,
                ht_passes=1, ht_plane=0):
    """samples int [ncomp, H, W] -> (codestream bytes, info dict with per-block tables for ISO jobs).
    lossy_step = None: lossless 5-3 + RCT.  lossy_step = s (a power of two, e.g. 2.0 or 0.5): irreversible 9-7 + ICT with
    the dead-zone quantiser of Annex E and the same step s in every band (scalar expounded QCD, mantissa 0).
    ht_passes = 2 or 3 with ht_plane = P >= 1: every block is one HT set of a cleanup pass at bit-plane P plus SigProp
    (and MagRef) passes at bit-plane P - 1; the packet header then carries two lengths (cleanup segment, refinement
    segment) and P fewer missing bit-planes.  Such a stream is lossy by up to 2^(P-1) in the coefficient domain (isolated
    small coefficients belong to no pass of the set)."""
    ncomp, H, W = samples.shape
    tile_w, tile_h = tile_w or W, tile_h or H
    ntx, nty = cdiv(W, tile_w), cdiv(H, tile_h)
    bands = band_list(nlevels)
    gain = {0: 0, 1: 1, 2: 1, 3: 2}
    eps = [prec + gain[b] + extra_eps for (_, b, _) in bands]
    irrev = lossy_step is not None
    if irrev:
        k = int(round(np.log2(lossy_step)))
        assert 2.0 ** k == lossy_step, "lossy_step must be a power of two"
        eps = [prec + gain[b] - k for (_, b, _) in bands]            # step = 2^(Rb - eps) = 2^k, Rb = prec + gain
        guard = max(guard, 3)
    out = bytearray()
    out += struct.pack(">H", SOC)
    out += struct.pack(">HHHIIIIIIIIH", SIZ, 38 + 3 * ncomp, 0x4000, W, H, 0, 0, tile_w, tile_h, 0, 0, ncomp)
    for _ in range(ncomp):
        out += struct.pack(">BBB", prec - 1, 1, 1)
    out += struct.pack(">HHIH", CAP, 8, 0x00020000, 0)
    xcb = cb.bit_length() - 1
    out += struct.pack(">HHBBHBBBBBB", COD, 12, 0, 0, 1, 1 if (mct and ncomp >= 3) else 0, nlevels, xcb - 2, xcb - 2, 0x40,
                       0 if irrev else 1)
    if irrev:
        out += struct.pack(">HHB", QCD, 3 + 2 * len(bands), (guard << 5) | 2)
        for e in eps:
            out += struct.pack(">H", e << 11)
    else:
        out += struct.pack(">HHB", QCD, 3 + len(bands), guard << 5)
        for e in eps:
            out += struct.pack(">B", e << 3)
    blocks_info = []
    for ty in range(nty):
        for tx in range(ntx):
            tidx = ty * ntx + tx
            x0, y0, x1, y1 = tx * tile_w, ty * tile_h, min((tx + 1) * tile_w, W), min((ty + 1) * tile_h, H)
            if irrev:
                fp = forward_tile_iso_97(samples[:, y0:y1, x0:x1], prec, nlevels, mct)
                planes = [(np.sign(p) * np.floor(np.abs(p) / lossy_step)).astype(np.int32) for p in fp]
            else:
                planes = forward_tile_iso(samples[:, y0:y1, x0:x1], prec, nlevels, mct)
            body = bytearray()
            for r in range(nlevels + 1):
                for c in range(ncomp):
                    pkt_blocks = []
                    rb = [(bi, b, l) for bi, (rr, b, l) in enumerate(bands) if rr == r]
                    per_band = []
                    for bi, b, l in rb:
                        bx0, by0, bx1, by1 = band_rect(x0, y0, x1, y1, b, l)
                        cbs, gw, gh = cblk_grid(bx0, by0, bx1, by1, cb, cb)
                        ox, oy = band_origin_in_plane(x0, y0, x1, y1, b, l)
                        mb = eps[bi] + guard - 1
                        entries = []
                        for (cx0, cy0, cx1, cy1) in cbs:
                            px, py = ox + cx0 - bx0, oy + cy0 - by0
                            blk = planes[c][py:py + cy1 - cy0, px:px + cx1 - cx0]
                            lcup = 0
                            if ht_passes > 1 or ht_plane:
                                data, lcup, _ = iso_ht_encode_passes(blk, cx1 - cx0, cy1 - cy0, ht_plane, ht_passes)
                            else:
                                data = iso_ht_encode(blk, cx1 - cx0, cy1 - cy0)
                            entries.append(dict(data=data, px=px, py=py, w=cx1 - cx0, h=cy1 - cy0, band=b, level=l,
                                                mb=mb, comp=c, tile=tidx, res=r, lcup=lcup or len(data),
                                                passes=(ht_passes if len(data) > (lcup or len(data)) else 1) if data else 0))
                        per_band.append((entries, gw, gh, mb, bx0 // cb, by0 // cb, cbs))
                    bw = BitWriter()
                    if not any(e["data"] for pb in per_band for e in pb[0]):
                        bw.put(0)
                        body += bw.finish()
                        for pb in per_band:
                            blocks_info += pb[0]
                        continue
                    bw.put(1)
                    for entries, gw, gh, mb, gx0, gy0, cbs in per_band:
                        if not entries:
                            continue
                        incl, imsb = TagTree(gw, gh), TagTree(gw, gh)
                        iv = np.zeros((gh, gw), np.int64)
                        mv = np.zeros((gh, gw), np.int64)
                        for k, e in enumerate(entries):
                            gx, gy = k % gw, k // gw
                            iv[gy, gx] = 0 if e["data"] else 1
                            mv[gy, gx] = mb - 1 - ht_plane     # HT cleanup at bit-plane P carries every bit above it: Mb - 1 - P missing
                        incl.set_values(iv)
                        imsb.set_values(mv)
                        for k, e in enumerate(entries):
                            gx, gy = k % gw, k // gw
                            incl.encode(bw, gx, gy, 1)
                            if not e["data"]:
                                continue
                            imsb.encode(bw, gx, gy, 1 << 20)
                            npass = e["passes"]
                            _put_npasses(bw, npass)
                            # HT: the cleanup pass is one codeword segment, SigProp + MagRef together the second (T.814 B.10.7)
                            n1, n2 = e["lcup"], len(e["data"]) - e["lcup"]
                            x2 = (npass - 1).bit_length() - 1 if npass > 1 else 0
                            lblock = 3
                            while n1 >= (1 << lblock) or (npass > 1 and n2 >= (1 << (lblock + x2))):
                                bw.put(1)
                                lblock += 1
                            bw.put(0)
                            bw.bits(n1, lblock)
                            if npass > 1:
                                bw.bits(n2, lblock + x2)
                            pkt_blocks.append(e["data"])
                        blocks_info += entries
                    body += bw.finish()
                    for d in pkt_blocks:
                        body += d
            out += struct.pack(">HHHIBB", SOT, 10, tidx, 14 + len(body), 0, 1)
            out += struct.pack(">H", SOD)
            out += body
    out += struct.pack(">H", EOC)
    info = dict(width=W, height=H, ncomp=ncomp, prec=prec, nlevels=nlevels, mct=mct, tile_w=tile_w, tile_h=tile_h,
                blocks=blocks_info, guard=guard, eps=eps, reversible=0 if irrev else 1, ht=1, lossy_step=lossy_step)
    return bytes(out), info


# ------------------------------------------------------------------------------------------------ parser
def _npasses(br):
    """number of coding passes codeword (B.10.6)"""
    if not br.get():
        return 1
    if not br.get():
        return 2
    v = br.bits(2)
    if v != 3:
        return 3 + v
    v = br.bits(5)
    if v != 31:
        return 6 + v
    return 37 + br.bits(7)


def parse_codestream(data, keep_packets=False):
    """main header + tile-parts + packet headers -> dict(width, height, ncomp, prec, sgnd, nlevels, reversible, mct,
    layers, guard, tile_w, tile_h, blocks=[dict(tile, comp, res, band, level, px, py, w, h, data, passes, zbp, mb,
    expn, mant)]).  Restrictions: see the module docstring; raises ValueError on anything else.
    keep_packets: also return main_segments = [(marker, segment bytes)] and packets[tile] = [(header bytes, body bytes)]
    for reassemble()."""
    d = bytes(data)
    pos = 0
    main_segments, packets = [], {}

    def u16(p):
        return struct.unpack(">H", d[p:p + 2])[0]

    if u16(0) != SOC:
        raise ValueError("no SOC")
    pos = 2
    hdr = {}
    qcd = None
    while True:
        m = u16(pos)
        if m == SOT:
            break
        L = u16(pos + 2)
        seg = d[pos + 4:pos + 2 + L]
        main_segments.append((m, seg))
        if m == SIZ:
            rsiz, xs, ys, xo, yo, xt, yt, xto, yto, nc = struct.unpack(">HIIIIIIIIH", seg[:36])
            if xo or yo or xto or yto:
                raise ValueError("non-zero image / tile origin")
            comps = [struct.unpack(">BBB", seg[36 + 3 * i:39 + 3 * i]) for i in range(nc)]
            if any(c[1] != 1 or c[2] != 1 for c in comps) or len({c[0] for c in comps}) != 1:
                raise ValueError("sub-sampled or mixed-depth components")
            hdr.update(width=xs, height=ys, tile_w=xt, tile_h=yt, ncomp=nc, prec=(comps[0][0] & 0x7F) + 1,
                       sgnd=comps[0][0] >> 7)
        elif m == COD:
            scod, prog, layers, mct, nl, xcb, ycb, sty, xf = struct.unpack(">BBHBBBBBB", seg[:10])
            if scod & 1:
                raise ValueError("user precincts")
            if scod & 6:
                raise ValueError("SOP / EPH markers")
            if prog not in (0, 1):
                raise ValueError("progression order %d" % prog)
            hdr.update(prog=prog, layers=layers, mct=mct, nlevels=nl, cbw=1 << (xcb + 2), cbh=1 << (ycb + 2),
                       cblk_style=sty, reversible=int(xf == 1))
        elif m == CAP:
            hdr["ht"] = 1
        elif m == QCD:
            sq = seg[0]
            style, guard = sq & 31, sq >> 5
            if style == 0:
                qcd = [(b >> 3, 0) for b in seg[1:]]
            elif style == 2:
                qcd = [(u16(pos + 5 + 2 * i) >> 11, u16(pos + 5 + 2 * i) & 0x7FF) for i in range((L - 3) // 2)]
            else:
                raise ValueError("derived quantisation")
            hdr.update(guard=guard, qstyle=style)
        elif m in (0xFF53, 0xFF5D, 0xFF5E, 0xFF5F, 0xFF60, 0xFF61):       # COC QCC RGN POC PPM PPT
            raise ValueError("marker %04X not supported" % m)
        pos += 2 + L
    hdr.setdefault("ht", 0)
    nl, nc, W, H = hdr["nlevels"], hdr["ncomp"], hdr["width"], hdr["height"]
    if hdr["cblk_style"] & ~0x7A or (hdr["cblk_style"] & 0x40 and hdr["cblk_style"] & 0x3F):     # not BYPASS / TERMALL
        raise ValueError("code-block style %02X" % hdr["cblk_style"])
    bands = band_list(nl)
    ntx, nty = cdiv(W, hdr["tile_w"]), cdiv(H, hdr["tile_h"])
    # per tile: concatenated packet bytes of its tile-parts
    bodies = {}
    while pos < len(d) and u16(pos) == SOT:
        lsot, isot, psot, tp, tn = struct.unpack(">HHIBB", d[pos + 2:pos + 12])
        end = pos + psot if psot else len(d) - 2
        p = pos + 12
        while u16(p) != SOD:
            p += 2 + u16(p + 2)
        bodies.setdefault(isot, bytearray()).extend(d[p + 2:end])
        pos = end
    blocks = []
    for tidx in range(ntx * nty):
        body = bytes(bodies.get(tidx, b""))
        tx, ty = tidx % ntx, tidx // ntx
        x0, y0 = tx * hdr["tile_w"], ty * hdr["tile_h"]
        x1, y1 = min(x0 + hdr["tile_w"], W), min(y0 + hdr["tile_h"], H)
        # code-block state per (comp, band index)
        state = {}
        for c in range(nc):
            for bi, (r, b, lvl) in enumerate(bands):
                bx0, by0, bx1, by1 = band_rect(x0, y0, x1, y1, b, lvl)
                cbs, gw, gh = cblk_grid(bx0, by0, bx1, by1, hdr["cbw"], hdr["cbh"])
                ox, oy = band_origin_in_plane(x0, y0, x1, y1, b, lvl)
                expn, mant = qcd[bi]
                ents = [dict(tile=tidx, comp=c, res=r, band=b, level=lvl, px=ox + cx0 - bx0, py=oy + cy0 - by0,
                             w=cx1 - cx0, h=cy1 - cy0, data=bytearray(), passes=0, zbp=0, lblock=3, included=False,
                             mb=hdr["guard"] + expn - 1, expn=expn, mant=mant)
                        for (cx0, cy0, cx1, cy1) in cbs]
                state[(c, bi)] = dict(ents=ents, gw=gw, gh=gh, incl=TagTree(gw, gh) if ents else None,
                                      imsb=TagTree(gw, gh) if ents else None)
        p = 0
        order = [(l, r, c) for l in range(hdr["layers"]) for r in range(nl + 1) for c in range(nc)] if hdr["prog"] == 0 else \
                [(l, r, c) for r in range(nl + 1) for l in range(hdr["layers"]) for c in range(nc)]
        for (layer, r, c) in order:
            if p >= len(body):
                break
            br = BitReader(body, p)
            segs = []
            if br.get():
                for bi, (rr, b, lvl) in enumerate(bands):
                    if rr != r:
                        continue
                    st = state[(c, bi)]
                    for k, e in enumerate(st["ents"]):
                        gx, gy = k % st["gw"], k // st["gw"]
                        if not e["included"]:
                            inc = st["incl"].decode(br, gx, gy, layer + 1)
                        else:
                            inc = bool(br.get())
                        if not inc:
                            continue
                        if not e["included"]:
                            t = 1
                            while not st["imsb"].decode(br, gx, gy, t):
                                t += 1
                            e["zbp"] = int(st["imsb"].val[0][gy, gx])
                            e["included"] = True
                        n = _npasses(br)
                        while br.get():
                            e["lblock"] += 1
                        if hdr["ht"]:
                            # HT (T.814 B.10.7): the cleanup pass is a codeword segment of its own, the SigProp and MagRef
                            # passes share the second one
                            if e["passes"] or n > 3:
                                raise ValueError("more than one HT set per code block")
                            ln = br.bits(e["lblock"])
                            e["lcup"] = ln
                            if n > 1:
                                ln += br.bits(e["lblock"] + ((n - 1).bit_length() - 1))
                        else:
                            nbits = e["lblock"] + (n.bit_length() - 1)
                            ln = br.bits(nbits)
                        e["passes"] += n
                        segs.append((e, ln))
            p0, p = p, br.align()
            p1 = p
            for e, ln in segs:
                e["data"] += body[p:p + ln]
                p += ln
            if keep_packets:
                packets.setdefault(tidx, []).append((body[p0:p1], body[p1:p]))
        for key in sorted(state):
            blocks += state[key]["ents"]
    for e in blocks:
        e["data"] = bytes(e["data"])
        e["num_bps"] = max(e["mb"] - e["zbp"], 0)
    hdr["blocks"] = blocks
    hdr["qcd"] = qcd
    if keep_packets:
        hdr["main_segments"], hdr["packets"] = main_segments, packets
    return hdr


def reassemble(h, sop=False, eph=False, tlm=False, split=False, plt=False):
    """Re-emit a codestream parsed with keep_packets=True in another legal shape: SOP marker segments before and / or EPH
    markers after every packet header (A.8), a TLM marker segment in the main header (A.7.1), PLT marker segments in the
    tile-part headers (A.7.3), every tile cut into two tile-parts at a packet boundary.  Packet bytes are untouched."""
    ntiles = cdiv(h["width"], h["tile_w"]) * cdiv(h["height"], h["tile_h"])
    parts = []                                               # (tile, tile-part index, number of tile-parts, [packet bytes])
    for t in range(ntiles):
        pk = []
        for i, (hd, body) in enumerate(h["packets"].get(t, [])):
            b = bytearray()
            if sop:
                b += struct.pack(">HHH", 0xFF91, 4, i & 0xFFFF)
            b += hd
            if eph:
                b += struct.pack(">H", 0xFF92)
            b += body
            pk.append(bytes(b))
        if split and len(pk) >= 2:
            k = len(pk) // 2
            parts += [(t, 0, 2, pk[:k]), (t, 1, 2, pk[k:])]
        else:
            parts.append((t, 0, 1, pk))
    tps = []
    for (t, tp, ntp, pk) in parts:
        hdrs = bytearray()
        if plt:
            lens = bytearray()
            for b in pk:
                n, enc = len(b), []
                while True:
                    enc.append(n & 0x7F)
                    n >>= 7
                    if not n:
                        break
                for j, v in enumerate(reversed(enc)):
                    lens.append(v | (0x80 if j < len(enc) - 1 else 0))
            z = 0
            while lens or z == 0:                            # a PLT segment holds at most 65535 - 3 bytes of lengths
                chunk, lens = lens[:60000], lens[60000:]
                while lens and chunk[-1] & 0x80:             # never cut inside a length
                    chunk.append(lens.pop(0))
                hdrs += struct.pack(">HHB", 0xFF58, 3 + len(chunk), z) + chunk
                z += 1
        body = b"".join(pk)
        tps.append(struct.pack(">HHHIBB", SOT, 10, t, 12 + len(hdrs) + 2 + len(body), tp, ntp) + hdrs + struct.pack(">H", SOD) + body)
    out = bytearray(struct.pack(">H", SOC))
    for (m, seg) in h["main_segments"]:
        if m in (0xFF55, 0xFF57):                            # existing TLM / PLM: lengths change
            continue
        if m == COD:
            seg = bytes([(seg[0] & ~6) | (2 if sop else 0) | (4 if eph else 0)]) + seg[1:]
        out += struct.pack(">HH", m, len(seg) + 2) + seg
    if tlm:                                                  # Stlm: ST = 2 (16-bit tile index), SP = 1 (32-bit lengths)
        ent = b"".join(struct.pack(">HI", t, len(b)) for (t, _, _, _), b in zip(parts, tps))
        out += struct.pack(">HHBB", 0xFF55, 4 + len(ent), 0, 0x60) + ent
    for b in tps:
        out += b
    out += struct.pack(">H", EOC)
    return bytes(out)


def wrap_jp2(codestream, enumcs=16, width=0, height=0, ncomp=3, bpc=7, colr_method=1, extra_colr=None, header_after=False):
    """TEST HARNESS: a minimal JP2 file (ISO/IEC 15444-1 Annex I) around a codestream: signature, file type, JP2 header
    (image header + colour specification box[es]) and contiguous codestream box.  extra_colr: EnumCS values of further colr
    boxes (the reference keeps the last, box.go:427-431); header_after: the header box behind the codestream box."""
    def box(t, payload):
        return struct.pack(">I4s", 8 + len(payload), t) + payload
    def colr(cs):
        if colr_method == 1:
            return box(b"colr", struct.pack(">BBBI", 1, 0, 0, cs))
        return box(b"colr", struct.pack(">BBB", colr_method, 0, 0) + bytes(16))          # an (empty) ICC profile
    ihdr = box(b"ihdr", struct.pack(">IIHBBBB", height, width, ncomp, bpc, 7, 0, 0))
    jp2h = box(b"jp2h", ihdr + b"".join(colr(c) for c in [enumcs] + list(extra_colr or [])))
    head = box(b"jP  ", b"\r\n\x87\n") + box(b"ftyp", b"jp2 " + struct.pack(">I", 0) + b"jp2 ")
    body = box(b"jp2c", bytes(codestream))
    return head + (body + jp2h if header_after else jp2h + body)


def with_qcc(data, scramble_qcd=True):
    """TEST HARNESS: the same codestream with a QCC marker segment per component that repeats what QCD says, and (scramble_qcd)
    a QCD rewritten to other guard bits / exponents -- every component then takes its quantisation from its QCC, so a decoder
    that honours QCC (A.6.5) reproduces the original image and one that does not cannot."""
    data = bytes(data)
    pos, ncomp, out, qcd_at = 2, None, None, None
    while True:
        m, L = struct.unpack(">HH", data[pos:pos + 4])
        if m == SOT:
            break
        if m == SIZ:
            ncomp = struct.unpack(">H", data[pos + 38:pos + 40])[0]
        if m == QCD:
            qcd_at = (pos, L)
        pos += 2 + L
    p, L = qcd_at
    body = data[p + 4:p + 2 + L]                                     # Sqcd + SPqcd
    qccs = b"".join(struct.pack(">HH", 0xFF5D, 2 + 1 + len(body)) + bytes([c]) + body for c in range(ncomp))
    qcd = bytearray(data[p:p + 2 + L])
    if scramble_qcd:
        style = body[0] & 31
        qcd[4] = (((body[0] >> 5) + 1) & 7) << 5 | style              # another number of guard bits
        if style == 0:
            for i in range(5, len(qcd)):
                qcd[i] = (((qcd[i] >> 3) + 2) & 31) << 3
        else:
            for i in range(5, len(qcd) - 1, 2):
                v = struct.unpack(">H", qcd[i:i + 2])[0]
                qcd[i:i + 2] = struct.pack(">H", ((((v >> 11) + 1) & 31) << 11) | ((v + 77) & 0x7FF))
    return data[:p] + bytes(qcd) + qccs + data[p + 2 + L:]


def with_coc(data, scramble_cod=True):
    """TEST HARNESS: the same codestream with a COC marker segment per component that repeats COD's SPcod, and (scramble_cod) a
    COD that announces other code-block dimensions -- every component then takes its coding parameters from its COC (A.6.2)."""
    data = bytes(data)
    pos, ncomp, cod_at = 2, None, None
    while True:
        m, L = struct.unpack(">HH", data[pos:pos + 4])
        if m == SOT:
            break
        if m == SIZ:
            ncomp = struct.unpack(">H", data[pos + 38:pos + 40])[0]
        if m == COD:
            cod_at = (pos, L)
        pos += 2 + L
    p, L = cod_at
    scod, spcod = data[p + 4], data[p + 9:p + 2 + L]                 # SPcod: levels, xcb, ycb, style, transform[, precincts]
    cocs = b"".join(struct.pack(">HH", 0xFF53, 2 + 1 + 1 + len(spcod)) + bytes([c, scod & 1]) + spcod for c in range(ncomp))
    cod = bytearray(data[p:p + 2 + L])
    if scramble_cod:
        cod[10] = cod[11] = 3 if cod[10] != 3 else 2                  # 32 x 32 (or 16 x 16) code blocks, says COD
    return data[:p] + bytes(cod) + cocs + data[p + 2 + L:]
