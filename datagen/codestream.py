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

from . import fwd2d53, iso_ht_encode

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


# ------------------------------------------------------------------------------------------------ writer
def write_htj2k(samples, prec, tile_w=None, tile_h=None, nlevels=5, mct=1, cb=64, guard=2, extra_eps=1):
    """samples int [ncomp, H, W] -> (codestream bytes, info dict with per-block tables for ISO jobs)"""
    ncomp, H, W = samples.shape
    tile_w, tile_h = tile_w or W, tile_h or H
    ntx, nty = cdiv(W, tile_w), cdiv(H, tile_h)
    bands = band_list(nlevels)
    gain = {0: 0, 1: 1, 2: 1, 3: 2}
    eps = [prec + gain[b] + extra_eps for (_, b, _) in bands]
    out = bytearray()
    out += struct.pack(">H", SOC)
    out += struct.pack(">HHHIIIIIIIIH", SIZ, 38 + 3 * ncomp, 0x4000, W, H, 0, 0, tile_w, tile_h, 0, 0, ncomp)
    for _ in range(ncomp):
        out += struct.pack(">BBB", prec - 1, 1, 1)
    out += struct.pack(">HHIH", CAP, 8, 0x00020000, 0)
    xcb = cb.bit_length() - 1
    out += struct.pack(">HHBBHBBBBBB", COD, 12, 0, 0, 1, 1 if (mct and ncomp >= 3) else 0, nlevels, xcb - 2, xcb - 2, 0x40, 1)
    out += struct.pack(">HHB", QCD, 3 + len(bands), guard << 5)
    for e in eps:
        out += struct.pack(">B", e << 3)
    blocks_info = []
    for ty in range(nty):
        for tx in range(ntx):
            tidx = ty * ntx + tx
            x0, y0, x1, y1 = tx * tile_w, ty * tile_h, min((tx + 1) * tile_w, W), min((ty + 1) * tile_h, H)
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
                            data = iso_ht_encode(blk, cx1 - cx0, cy1 - cy0)
                            entries.append(dict(data=data, px=px, py=py, w=cx1 - cx0, h=cy1 - cy0, band=b, level=l,
                                                mb=mb, comp=c, tile=tidx, res=r))
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
                            mv[gy, gx] = mb - 1                # HT cleanup carries every magnitude bit: P = Mb - 1
                        incl.set_values(iv)
                        imsb.set_values(mv)
                        for k, e in enumerate(entries):
                            gx, gy = k % gw, k // gw
                            incl.encode(bw, gx, gy, 1)
                            if not e["data"]:
                                continue
                            imsb.encode(bw, gx, gy, 1 << 20)
                            bw.put(0)                          # one coding pass
                            n = len(e["data"])
                            lblock = 3
                            while n >= (1 << lblock):
                                bw.put(1)
                                lblock += 1
                            bw.put(0)
                            bw.bits(n, lblock)
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
                blocks=blocks_info, guard=guard, eps=eps, reversible=1, ht=1)
    return bytes(out), info
