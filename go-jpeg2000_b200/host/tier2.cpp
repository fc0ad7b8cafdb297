// tier2.cpp -- see tier2.h.  Main header (SIZ, CAP, COD, COC, QCD, QCC, TLM, COM), tile-part index (SOT / Psot, TLM cross-check), packet
// headers (tag trees, number of passes, Lblock, segment lengths; SOP / EPH; PLT cross-check) for all five progression orders
// with maximal or user-defined precincts, any number of quality layers, classic and HT code blocks (one HT set: the cleanup length and the
// SigProp + MagRef length are separate codeword segments, T.814 B.10.7).  Tiles are parsed concurrently.
// Everything is bounds-checked: malformed bytes give an error code, never a fault (the reference's fuzz contract).
#include "tier2.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <new>
#include <thread>

namespace {

constexpr uint16_t SOC = 0xFF4F, SIZ = 0xFF51, CAP = 0xFF50, COD = 0xFF52, COC = 0xFF53, QCD = 0xFF5C, QCC = 0xFF5D, RGN = 0xFF5E,
                   POC = 0xFF5F, TLM = 0xFF55, PLM = 0xFF57, PLT = 0xFF58, PPM = 0xFF60, PPT = 0xFF61, SOT = 0xFF90, SOP = 0xFF91,
                   EPH = 0xFF92, SOD = 0xFF93, EOC = 0xFFD9;

struct Err {
    int code = 0;
    std::string msg;
    int fail(int c, const char *fmt, ...)
    {
        if (code) return code;
        char buf[256];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        code = c; msg = buf;
        return c;
    }
};

inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }

struct Header {
    uint32_t W = 0, H = 0, tile_w = 0, tile_h = 0, ncomp = 0, prec = 0, sgnd = 0;
    uint32_t prog = 0, layers = 0, mct = 0, nlevels = 0, cbw = 0, cbh = 0, style = 0, reversible = 0, sop = 0, eph = 0, ht = 0;
    uint32_t guard = 0;
    uint32_t colorspace = 0;                             // J2KGPU_CS_* from the JP2 colour specification box (0: none / raw codestream)
    uint8_t ppx[33], ppy[33];                            // precinct size exponents per resolution (15 = maximal)
    bool have_siz = false, have_cod = false, have_qcd = false;
    std::vector<std::pair<uint32_t, uint32_t>> q;      // per band in codestream order: exponent, mantissa
    // QCC (A.6.5): a component's own guard bits and step sizes override QCD's for that component
    struct QComp { bool set = false; uint32_t guard = 0; std::vector<std::pair<uint32_t, uint32_t>> q; };
    std::vector<QComp> qc;                             // per component (empty: no QCC in the main header)
    // COC (A.6.2): a component's own decomposition levels, code-block size / style, transform and precinct sizes.  The device
    // path wants the components of a tile to share one geometry, so COC is accepted when every component ends up with the same
    // parameters (each from its COC, or from COD); they then replace COD's.
    struct CComp { bool set = false; uint32_t nlevels = 0, cbw = 0, cbh = 0, style = 0, reversible = 0; uint8_t ppx[33], ppy[33]; };
    std::vector<CComp> cc;
    uint32_t guard_of(uint32_t c) const { return c < qc.size() && qc[c].set ? qc[c].guard : guard; }
    const std::vector<std::pair<uint32_t, uint32_t>> &q_of(uint32_t c) const { return c < qc.size() && qc[c].set ? qc[c].q : q; }
};

struct BandId { uint32_t res, band, lvl; };

std::vector<BandId> band_list(uint32_t nl)
{
    std::vector<BandId> out;
    out.push_back({0, 0, nl});
    for (uint32_t r = 1; r <= nl; r++) {
        const uint32_t lvl = nl - r + 1;
        out.push_back({r, 1, lvl}); out.push_back({r, 2, lvl}); out.push_back({r, 3, lvl});
    }
    return out;
}

struct Rect { int64_t x0, y0, x1, y1; };

// band bounds in band coordinates (ISO/IEC 15444-1 B.15): HL / HH are offset in x, LH / HH in y
Rect band_rect(int64_t tx0, int64_t ty0, int64_t tx1, int64_t ty1, uint32_t band, uint32_t lvl)
{
    const int64_t s = (int64_t)1 << lvl;
    if (band == 0 || lvl == 0) return {cdiv64(tx0, s), cdiv64(ty0, s), cdiv64(tx1, s), cdiv64(ty1, s)};
    const int64_t hs = (int64_t)1 << (lvl - 1), xo = band & 1, yo = band >> 1;
    return {cdiv64(tx0 - hs * xo, s), cdiv64(ty0 - hs * yo, s), cdiv64(tx1 - hs * xo, s), cdiv64(ty1 - hs * yo, s)};
}

// packet-header bit reader: MSB first, the byte after 0xFF carries 7 bits (B.10.1); reads past `end` give zeros and set bad
struct BitReader {
    const uint8_t *d; size_t pos, end;
    uint32_t cur = 0; int left = 0; bool prev_ff = false, bad = false;
    BitReader(const uint8_t *d_, size_t pos_, size_t end_) : d(d_), pos(pos_), end(end_) {}
    uint32_t get()
    {
        if (left == 0) {
            if (pos >= end) { bad = true; return 0; }
            cur = d[pos++];
            left = prev_ff ? 7 : 8;
            prev_ff = cur == 0xFF;
        }
        left--;
        return (cur >> left) & 1u;
    }
    uint32_t bits(int n) { uint32_t v = 0; for (int i = 0; i < n; i++) v = (v << 1) | get(); return v; }
    size_t align()
    {
        if (prev_ff && left == 0) pos++;                 // a header may not end in 0xFF: the stuffed byte follows
        left = 0; prev_ff = false;
        return pos;
    }
};

// Tag tree (B.10.2) whose nodes live in an arena shared by all the trees of a tile (a 4K frame has some 13 000 of them:
// one heap block per tree and array made the parser spend its time in malloc)
struct TagArena {
    std::vector<int32_t> val, low;
    std::vector<uint8_t> known;
    size_t grab(size_t n)
    {
        const size_t off = val.size();
        val.resize(off + n, 1 << 30); low.resize(off + n, 0); known.resize(off + n, 0);
        return off;
    }
};

struct TagTree {
    struct Lvl { uint32_t w, h; uint32_t off; };
    Lvl lv[17];
    uint32_t nlv = 0;
    void init(TagArena &ar, uint32_t w, uint32_t h)
    {
        nlv = 0;
        if (!w || !h) return;
        size_t n = 0;
        uint32_t ww = w, hh = h;
        for (;;) {
            lv[nlv++] = {ww, hh, (uint32_t)n};
            n += (size_t)ww * hh;
            if ((ww <= 1 && hh <= 1) || nlv == 17) break;
            ww = cdiv(ww, 2); hh = cdiv(hh, 2);
        }
        const size_t base = ar.grab(n);
        for (uint32_t l = 0; l < nlv; l++) lv[l].off += (uint32_t)base;
    }
    // -> true when value(x, y) < threshold is established
    bool decode(TagArena &ar, BitReader &br, uint32_t x, uint32_t y, int32_t threshold)
    {
        int32_t lo = 0;
        for (int l = (int)nlv - 1; l >= 0; l--) {
            const size_t i = lv[l].off + (size_t)(y >> l) * lv[l].w + (x >> l);
            if (lo > ar.low[i]) ar.low[i] = lo; else lo = ar.low[i];
            while (lo < threshold && !ar.known[i]) {
                if (br.bad) return false;
                if (br.get()) { ar.known[i] = 1; ar.val[i] = lo; }
                else lo++;
            }
            ar.low[i] = lo;
            if (ar.known[i] && l) lo = std::max(lo, ar.val[i]);
        }
        const size_t i0 = lv[0].off + (size_t)y * lv[0].w + x;
        return ar.known[i0] && ar.val[i0] < threshold;
    }
    int32_t value(const TagArena &ar, uint32_t x, uint32_t y) const { return ar.val[lv[0].off + (size_t)y * lv[0].w + x]; }
};

uint32_t read_npasses(BitReader &br)                      // B.10.6
{
    if (!br.get()) return 1;
    if (!br.get()) return 2;
    uint32_t v = br.bits(2);
    if (v != 3) return 3 + v;
    v = br.bits(5);
    if (v != 31) return 6 + v;
    return 37 + br.bits(7);
}

inline int floorlog2(uint32_t v) { int r = 0; while (v >>= 1) r++; return r; }

struct Blk {
    uint32_t comp, bidx;               // component, index into the band list
    uint32_t px, py, w, h;             // placement in the tile-component's Mallat plane
    uint32_t passes = 0, zbp = 0, lblock = 3, lcup = 0;
    bool included = false;
    uint64_t off = 0; uint32_t len = 0; // first contribution (offset into the tile body)
    uint32_t npieces = 0;
    int32_t more_head = -1, more_tail = -1;            // further contributions (quality layers): list in the tile's `pieces`
    std::vector<uint32_t> segl;                        // BYPASS / TERMALL blocks: bytes of each codeword segment so far
};

// Codeword segments of a classic block whose style terminates inside the block (B.10.7.2, D.4, D.6): with TERMALL every coding
// pass is a segment; with selective bypass alone the first ten passes share one, then each (significance + refinement) pair
// is a raw segment and each cleanup pass an MQ segment.  -> segment index of pass i, and how many passes a segment may hold.
inline uint32_t seg_of_pass(uint32_t style, uint32_t i)
{
    if (style & 0x04u) return i;
    if (i < 10) return 0;
    const uint32_t j = i - 10;
    return 1 + 2 * (j / 3) + (j % 3 == 2 ? 1 : 0);
}
inline uint32_t seg_capacity(uint32_t style, uint32_t seg)
{
    if (style & 0x04u) return 1;
    if (seg == 0) return 10;
    return (seg & 1) ? 2 : 1;
}
struct Piece { uint64_t off; uint32_t len; int32_t next; };

// the code blocks of one band that lie in one precinct: their own grid and tag trees (B.10.2)
struct PrecBand {
    uint32_t gw = 0, gh = 0, first = 0, count = 0;       // blocks [first, first + count) of the tile's block list, raster order
    TagTree incl, imsb;
};
struct Prec {
    uint32_t c, r, idx;                // component, resolution, raster index among the resolution's precincts
    int64_t x, y;                      // top-left corner on the reference grid (clamped to the tile): position of its packets
    uint32_t nb = 0;                   // bands at this resolution: 1 (LL) or 3
    PrecBand pb[3];
};

struct TilePart { uint64_t body, end; };

struct TileOut {
    std::vector<j2k_tilecomp_t> tcs;
    std::vector<j2k_cblk_t> cbs;
    std::vector<uint8_t> extra;                       // concatenated bytes of blocks that are not contiguous in the codestream
    std::vector<uint8_t> is_extra;                    // per block: data_off is relative to `extra`
    uint32_t packets = 0, plt_checked = 0;
    Err err;
};

// truncated: the codestream ends inside this tile's data -- its last, incomplete packet is dropped instead of being an error
void parse_tile(const uint8_t *cs, const Header &h, uint32_t tidx, const std::vector<TilePart> &parts,
                const std::vector<uint32_t> &plt, uint32_t reduce, bool truncated, TileOut &out)
{
    const uint32_t ntx = cdiv(h.W, h.tile_w);
    const uint32_t tx = tidx % ntx, ty = tidx / ntx;
    const int64_t x0 = (int64_t)tx * h.tile_w, y0 = (int64_t)ty * h.tile_h;
    const int64_t x1 = std::min<int64_t>(x0 + h.tile_w, h.W), y1 = std::min<int64_t>(y0 + h.tile_h, h.H);
    const std::vector<BandId> bands = band_list(h.nlevels);
    const uint32_t nl = h.nlevels, nc = h.ncomp;
    // the tile body: one tile-part -> parsed in place; several -> concatenated (their blocks then go through `extra`)
    std::vector<uint8_t> joined;
    const uint8_t *body;
    size_t blen;
    int64_t abs_off;
    if (parts.size() == 1) { body = cs + parts[0].body; blen = (size_t)(parts[0].end - parts[0].body); abs_off = (int64_t)parts[0].body; }
    else {
        for (const TilePart &p : parts) joined.insert(joined.end(), cs + p.body, cs + p.end);
        body = joined.data(); blen = joined.size(); abs_off = -1;
    }
    // precincts of every (component, resolution) and, per precinct and band, the code blocks inside it (B.6, B.7)
    std::vector<Blk> blks;
    std::vector<Prec> precs;
    std::vector<Piece> pieces;
    TagArena arena;
    for (uint32_t c = 0; c < nc; c++)
        for (uint32_t r = 0; r <= nl; r++) {
            const int64_t rs = (int64_t)1 << (nl - r);
            const int64_t trx0 = cdiv64(x0, rs), try0 = cdiv64(y0, rs), trx1 = cdiv64(x1, rs), try1 = cdiv64(y1, rs);
            if (trx1 <= trx0 || try1 <= try0) continue;
            const uint32_t ppx = h.ppx[r], ppy = h.ppy[r];
            const int64_t pw = (int64_t)1 << ppx, ph = (int64_t)1 << ppy;
            const int64_t pxa0 = trx0 >> ppx, pya0 = try0 >> ppy, pxa1 = cdiv64(trx1, pw), pya1 = cdiv64(try1, ph);
            if ((uint64_t)(pxa1 - pxa0) * (uint64_t)(pya1 - pya0) > 65536) { out.err.fail(J2KGPU_E_UNSUPPORTED, "tile %u: too many precincts", tidx); return; }
            uint32_t pidx = 0;
            for (int64_t pya = pya0; pya < pya1; pya++)
                for (int64_t pxa = pxa0; pxa < pxa1; pxa++, pidx++) {
                    Prec pr;
                    pr.c = c; pr.r = r; pr.idx = pidx;
                    pr.x = std::max(x0, pxa * pw * rs); pr.y = std::max(y0, pya * ph * rs);
                    pr.nb = r ? 3 : 1;
                    for (uint32_t k = 0; k < pr.nb; k++) {
                        const size_t bi = r ? 1 + 3 * (r - 1) + k : 0;
                        const BandId &b = bands[bi];
                        PrecBand &pb = pr.pb[k];
                        pb.first = (uint32_t)blks.size();
                        const Rect br = band_rect(x0, y0, x1, y1, b.band, b.lvl);
                        // the precinct in band coordinates (half the resolution's above resolution 0) and the block size it allows
                        const uint32_t sx = r ? ppx - 1 : ppx, sy = r ? ppy - 1 : ppy;
                        const int64_t cbw = std::min<int64_t>(h.cbw, (int64_t)1 << sx), cbh = std::min<int64_t>(h.cbh, (int64_t)1 << sy);
                        const int64_t qx0 = std::max(br.x0, pxa << sx), qy0 = std::max(br.y0, pya << sy);
                        const int64_t qx1 = std::min(br.x1, (pxa + 1) << sx), qy1 = std::min(br.y1, (pya + 1) << sy);
                        if (qx1 > qx0 && qy1 > qy0) {
                            const int64_t gx0 = qx0 / cbw, gy0 = qy0 / cbh, gx1 = cdiv64(qx1, cbw), gy1 = cdiv64(qy1, cbh);
                            pb.gw = (uint32_t)(gx1 - gx0); pb.gh = (uint32_t)(gy1 - gy0);
                            // origin of the band inside the Mallat plane: LL_lvl top-left, HL to the right, LH below
                            const int64_t s = (int64_t)1 << b.lvl;
                            const int64_t lw = cdiv64(x1, s) - cdiv64(x0, s), lh = cdiv64(y1, s) - cdiv64(y0, s);
                            const int64_t ox = (b.band & 1) ? lw : 0, oy = (b.band >> 1) ? lh : 0;
                            for (int64_t gy = gy0; gy < gy1; gy++)
                                for (int64_t gx = gx0; gx < gx1; gx++) {
                                    const int64_t cx0 = std::max(qx0, gx * cbw), cy0 = std::max(qy0, gy * cbh);
                                    const int64_t cx1 = std::min(qx1, (gx + 1) * cbw), cy1 = std::min(qy1, (gy + 1) * cbh);
                                    Blk k2;
                                    k2.comp = c; k2.bidx = (uint32_t)bi;
                                    k2.px = (uint32_t)(ox + cx0 - br.x0); k2.py = (uint32_t)(oy + cy0 - br.y0);
                                    k2.w = (uint32_t)(cx1 - cx0); k2.h = (uint32_t)(cy1 - cy0);
                                    blks.push_back(k2);
                                }
                        }
                        pb.count = (uint32_t)blks.size() - pb.first;
                        pb.incl.init(arena, pb.gw, pb.gh); pb.imsb.init(arena, pb.gw, pb.gh);
                    }
                    precs.push_back(pr);
                }
        }
    // packet sequence (B.12): one packet per layer, resolution, component and precinct; the position-driven orders visit a
    // precinct when the scan over the reference grid reaches its top-left corner, i.e. in the order of (y, x)
    struct Pk { uint32_t l, pi; };
    std::vector<Pk> order;
    order.reserve((size_t)h.layers * precs.size());
    for (uint32_t l = 0; l < h.layers; l++)
        for (uint32_t i = 0; i < precs.size(); i++) order.push_back({l, i});
    auto key = [&](const Pk &k, int64_t out5[5]) {
        const Prec &q = precs[k.pi];
        switch (h.prog) {
        case 0: out5[0] = k.l; out5[1] = q.r; out5[2] = q.c; out5[3] = q.idx; out5[4] = 0; break;            // LRCP
        case 1: out5[0] = q.r; out5[1] = k.l; out5[2] = q.c; out5[3] = q.idx; out5[4] = 0; break;            // RLCP
        case 2: out5[0] = q.r; out5[1] = q.y; out5[2] = q.x; out5[3] = q.c; out5[4] = k.l; break;            // RPCL
        case 3: out5[0] = q.y; out5[1] = q.x; out5[2] = q.c; out5[3] = q.r; out5[4] = k.l; break;            // PCRL
        default: out5[0] = q.c; out5[1] = q.y; out5[2] = q.x; out5[3] = q.r; out5[4] = k.l; break;           // CPRL
        }
    };
    std::stable_sort(order.begin(), order.end(), [&](const Pk &a2, const Pk &b2) {
        int64_t ka[5], kb[5];
        key(a2, ka); key(b2, kb);
        for (int i = 0; i < 5; i++) if (ka[i] != kb[i]) return ka[i] < kb[i];
        return false;
    });
    size_t p = 0;
    std::vector<std::pair<uint32_t, uint32_t>> segs;   // (block, length) of this packet
    struct Undo { uint32_t blk, passes, zbp, lblock, lcup; bool included; std::vector<uint32_t> segl; };
    std::vector<Undo> undo;                            // state of the blocks this packet's header touched (truncated tiles only)
    bool cut = false;
    for (const Pk &pk : order) {
        if (p >= blen) break;                            // truncated codestream: the remaining packets are absent
        const size_t pk_start = p;
        if (h.sop && p + 6 <= blen && body[p] == 0xFF && body[p + 1] == 0x91) p += 6;
        BitReader br(body, p, blen);
        segs.clear();
        undo.clear();
        if (br.get()) {
            Prec &pq = precs[pk.pi];
            for (uint32_t kb = 0; kb < pq.nb; kb++) {
                PrecBand &st = pq.pb[kb];
                for (uint32_t k = 0; k < st.count; k++) {
                    Blk &e = blks[st.first + k];
                    const uint32_t gx = k % st.gw, gy = k / st.gw;
                    bool inc;
                    if (!e.included) inc = st.incl.decode(arena, br, gx, gy, (int32_t)pk.l + 1);
                    else inc = br.get() != 0;
                    if (br.bad) { if (truncated) { cut = true; goto packet_done; } out.err.fail(J2KGPU_E_RANGE, "tile %u: packet header runs past the tile data", tidx); return; }
                    if (!inc) continue;
                    if (truncated) undo.push_back({st.first + k, e.passes, e.zbp, e.lblock, e.lcup, e.included, e.segl});
                    if (!e.included) {
                        int32_t t = 1;
                        while (!st.imsb.decode(arena, br, gx, gy, t)) {
                            if (br.bad && truncated) { cut = true; goto packet_done; }
                            if (br.bad || t > 64) { out.err.fail(J2KGPU_E_RANGE, "tile %u: bad zero-bit-plane tag tree", tidx); return; }
                            t++;
                        }
                        e.zbp = (uint32_t)st.imsb.value(arena, gx, gy);
                        e.included = true;
                    }
                    const uint32_t n = read_npasses(br);
                    while (br.get()) {
                        if (++e.lblock > 32 || br.bad) {
                            if (br.bad && truncated) { cut = true; goto packet_done; }
                            out.err.fail(J2KGPU_E_RANGE, "tile %u: bad Lblock", tidx); return;
                        }
                    }
                    uint32_t ln;
                    if (h.ht) {
                        // HT (T.814 B.10.7): the cleanup pass is a codeword segment of its own, SigProp + MagRef share the second
                        if (e.passes || n > 3) { out.err.fail(J2KGPU_E_UNSUPPORTED, "tile %u: more than one HT set per code block", tidx); return; }
                        ln = br.bits((int)e.lblock);
                        e.lcup = ln;
                        if (n > 1) ln += br.bits((int)e.lblock + floorlog2(n - 1));
                    } else if (h.style & 0x05u) {
                        // several codeword segments: one length per segment the new passes touch, each Lblock + floor(log2(passes
                        // of that segment in this packet)) bits wide (B.10.7.2)
                        ln = 0;
                        for (uint32_t done = 0; done < n;) {
                            const uint32_t i = e.passes + done, sg = seg_of_pass(h.style, i);
                            uint32_t first = i;                                    // passes of segment sg that came before
                            while (first > 0 && seg_of_pass(h.style, first - 1) == sg) first--;
                            const uint32_t m = std::min(n - done, seg_capacity(h.style, sg) - (i - first));
                            const int nb = (int)e.lblock + floorlog2(m);
                            if (nb > 32) { out.err.fail(J2KGPU_E_RANGE, "tile %u: segment length field too wide", tidx); return; }
                            const uint32_t l = br.bits(nb);
                            if (e.segl.size() <= sg) e.segl.resize(sg + 1, 0);
                            e.segl[sg] += l;
                            ln += l;
                            done += m;
                        }
                    } else {
                        const int nb = (int)e.lblock + floorlog2(n);
                        if (nb > 32) { out.err.fail(J2KGPU_E_RANGE, "tile %u: segment length field too wide", tidx); return; }
                        ln = br.bits(nb);
                    }
                    if (br.bad) { if (truncated) { cut = true; goto packet_done; } out.err.fail(J2KGPU_E_RANGE, "tile %u: packet header runs past the tile data", tidx); return; }
                    e.passes += n;
                    segs.push_back({st.first + k, ln});
                }
            }
        }
        p = br.align();
        if (h.eph && p + 2 <= blen && body[p] == 0xFF && body[p + 1] == 0x92) p += 2;
        {
            uint64_t total = 0;
            for (auto &sg : segs) total += sg.second;
            if (total > blen - std::min(p, blen)) {
                if (truncated) cut = true;
                else { out.err.fail(J2KGPU_E_RANGE, "tile %u: code-block data runs past the tile data", tidx); return; }
            }
        }
    packet_done:
        if (cut) {                                       // the codestream ended inside this packet: forget it and stop
            for (const Undo &u : undo) { Blk &e = blks[u.blk]; e.passes = u.passes; e.zbp = u.zbp; e.lblock = u.lblock; e.lcup = u.lcup; e.included = u.included; e.segl = u.segl; }
            break;
        }
        for (auto &sg : segs) {
            Blk &e = blks[sg.first];
            if (e.npieces == 0) { e.off = p; e.len = sg.second; }
            else {
                pieces.push_back({p, sg.second, -1});
                if (e.more_tail >= 0) pieces[e.more_tail].next = (int32_t)pieces.size() - 1; else e.more_head = (int32_t)pieces.size() - 1;
                e.more_tail = (int32_t)pieces.size() - 1;
            }
            e.npieces++;
            p += sg.second;
        }
        if (out.packets < plt.size()) {                  // PLT announced this packet's length: it must agree
            if (plt[out.packets] != p - pk_start) { out.err.fail(J2KGPU_E_RANGE, "tile %u packet %u: PLT length %u, parsed %zu", tidx, out.packets, plt[out.packets], p - pk_start); return; }
            out.plt_checked++;
        }
        out.packets++;
    }
    // ---- tables ----
    const uint32_t sc = reduce;
    auto red = [&](int64_t v) { return (uint32_t)cdiv64(v, (int64_t)1 << sc); };
    for (uint32_t c = 0; c < nc; c++) {
        j2k_tilecomp_t tc{};
        tc.comp = c; tc.x0 = red(x0); tc.y0 = red(y0); tc.x1 = red(x1); tc.y1 = red(y1);
        out.tcs.push_back(tc);
    }
    static const int gain[4] = {0, 1, 1, 2};
    for (const Blk &e : blks) {
        const BandId &b = bands[e.bidx];
        if (b.res > nl - reduce) continue;                 // ReduceResolution: the finest resolutions are not handed over
        const uint32_t expn = h.q_of(e.comp)[e.bidx].first, mant = h.q_of(e.comp)[e.bidx].second;
        const int mb = (int)h.guard_of(e.comp) + (int)expn - 1;
        j2k_cblk_t cb{};
        cb.tilecomp = e.comp; cb.x0 = (uint16_t)e.px; cb.y0 = (uint16_t)e.py; cb.w = (uint16_t)e.w; cb.h = (uint16_t)e.h;
        cb.band = (uint8_t)b.band; cb.level = (uint8_t)(b.lvl > reduce ? b.lvl - reduce : 0);
        // Annex E.1: step = 2^(Rb - eps) * (1 + mu / 2^11), Rb = precision + band gain
        cb.step = h.reversible ? 1.0f : (float)(std::ldexp(1.0, (int)h.prec + gain[b.band] - (int)expn) * (1.0 + mant / 2048.0));
        const int nb = mb - (int)e.zbp;
        uint32_t total = e.len;
        for (int32_t m = e.more_head; m >= 0; m = pieces[m].next) total += pieces[m].len;
        if (!e.passes || nb <= 0 || total == 0) { cb.num_bps = 0; cb.num_passes = 0; cb.data_len = 0; cb.data_off = 0; out.is_extra.push_back(0); out.cbs.push_back(cb); continue; }
        if (nb > 31) { out.err.fail(J2KGPU_E_UNSUPPORTED, "tile %u: %d magnitude bit-planes", tidx, nb); return; }
        cb.num_bps = (uint8_t)nb;
        cb.num_passes = (uint8_t)std::min<uint32_t>(e.passes, 255);
        cb.data_len = total;
        cb.len_cleanup = (h.ht && e.passes > 1) ? e.lcup : 0;
        if (e.more_head < 0 && abs_off >= 0 && e.segl.empty()) { cb.data_off = (uint64_t)abs_off + e.off; out.is_extra.push_back(0); }
        else {
            cb.data_off = out.extra.size();
            out.extra.insert(out.extra.end(), body + e.off, body + e.off + e.len);
            for (int32_t m = e.more_head; m >= 0; m = pieces[m].next) out.extra.insert(out.extra.end(), body + pieces[m].off, body + pieces[m].off + pieces[m].len);
            if (h.style & 0x05u) {
                // the segment lengths follow the block's bytes as little-endian 32-bit words, one per segment its passes touch
                // (include/j2kgpu.h, j2k_image_t.cblk_style)
                const uint32_t nseg = seg_of_pass(h.style, e.passes - 1) + 1;
                for (uint32_t s = 0; s < nseg; s++) {
                    const uint32_t l = s < e.segl.size() ? e.segl[s] : 0;
                    for (int k = 0; k < 4; k++) out.extra.push_back((uint8_t)(l >> (8 * k)));
                }
            }
            out.is_extra.push_back(1);
        }
        out.cbs.push_back(cb);
    }
}

inline uint32_t be16(const uint8_t *p) { return ((uint32_t)p[0] << 8) | p[1]; }
inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

}  // namespace

// a codestream whose main header and tile-part index have been read: its tiles can then be parsed independently
struct j2k_t2_frame {
    const uint8_t *d = nullptr; uint64_t len = 0; uint32_t reduce = 0;
    Header h;
    std::vector<BandId> bands;
    uint32_t ntiles = 0, ntp = 0, ntlm = 0;
    int64_t truncated_tile = -1;
    std::vector<std::vector<TilePart>> parts;
    std::vector<std::vector<uint32_t>> plt;
    std::vector<TileOut> touts;
};

int j2k_tier2_begin(const uint8_t *d, uint64_t len, uint32_t reduce, j2k_t2_frame **fout, std::string &errmsg)
{
    *fout = nullptr;
    j2k_t2_frame *F = new (std::nothrow) j2k_t2_frame();
    if (!F) { errmsg = "out of memory"; return J2KGPU_E_NOMEM; }
    std::unique_ptr<j2k_t2_frame> guard(F);
    Err err;
    Header &h = F->h;
#define T2_FAIL(...) do { err.fail(__VA_ARGS__); errmsg = err.msg; return err.code; } while (0)
    // a JP2 file (ISO/IEC 15444-1 Annex I: signature box first): walk the top-level boxes to the contiguous codestream box
    // (decoder.readJP2, decoder.go:206-253).  Of the JP2 header box only the colour specification is read here: its enumerated
    // colour space selects the conversion to sRGB the pixel epilogue applies (decoder.getColorSpace decoder.go:135-178 ->
    // getColorConversion colorspace.go:54-88); the last colr box wins and a header box behind the codestream box is never
    // seen, as in box.ParseJP2Header / readJP2.  Palette, channel definition, resolution stay with the caller (box.go).
    if (d && len >= 12 && be32(d) == 12 && be32(d + 4) == 0x6A502020u) {
        uint64_t bp = 0;
        bool found = false;
        while (bp + 8 <= len) {
            uint64_t bl = be32(d + bp), hl = 8;
            const uint32_t bt = be32(d + bp + 4);
            if (bl == 1) { if (bp + 16 > len) break; bl = ((uint64_t)be32(d + bp + 8) << 32) | be32(d + bp + 12); hl = 16; }
            else if (bl == 0) bl = len - bp;              // the box runs to the end of the file
            if (bl < hl || bl > len - bp) break;
            if (bt == 0x6A703263u) { d += bp + hl; len = bl - hl; found = true; break; }   // 'jp2c'
            if (bt == 0x6A703268u) {                                                       // 'jp2h': a superbox
                uint64_t sp = bp + hl;
                const uint64_t se = bp + bl;
                while (sp + 8 <= se) {
                    uint64_t sl = be32(d + sp), shl = 8;
                    const uint32_t st = be32(d + sp + 4);
                    if (sl == 1) { if (sp + 16 > se) break; sl = ((uint64_t)be32(d + sp + 8) << 32) | be32(d + sp + 12); shl = 16; }
                    else if (sl == 0) sl = se - sp;
                    if (sl < shl || sl > se - sp) break;
                    if (st == 0x636F6C72u) {                                               // 'colr' (box.go:287-307)
                        const uint8_t *c = d + sp + shl;
                        const uint64_t cl = sl - shl;
                        if (cl < 3) T2_FAIL(J2KGPU_E_RANGE, "color specification box too short");
                        uint32_t enumcs = 0;                                               // ICC methods leave it 0 (bi-level: no conversion)
                        if (c[0] == 1) {
                            if (cl < 7) T2_FAIL(J2KGPU_E_RANGE, "color specification box too short for enumerated CS");
                            enumcs = be32(c + 3);
                        }
                        switch (enumcs) {                                                  // box.go:265-283 -> decoder.go:140-177 -> colorspace.go:54-88
                        case 1: case 18: case 22: case 23: case 24: h.colorspace = J2KGPU_CS_YCC709; break;   // YCbCr(1), sYCC, YPbPr 1125 / 1250, e-sYCC
                        case 3: case 4: h.colorspace = J2KGPU_CS_YCC601; break;            // YCbCr(2), YCbCr(3)
                        case 9: h.colorspace = J2KGPU_CS_PHOTOYCC; break;
                        case 11: h.colorspace = J2KGPU_CS_CMY; break;
                        case 12: h.colorspace = J2KGPU_CS_CMYK; break;
                        case 13: h.colorspace = J2KGPU_CS_YCCK; break;
                        case 14: h.colorspace = J2KGPU_CS_CIELAB; break;
                        case 19: h.colorspace = J2KGPU_CS_CIEJAB; break;
                        case 20: h.colorspace = J2KGPU_CS_ESRGB; break;
                        case 21: h.colorspace = J2KGPU_CS_ROMM; break;
                        default: h.colorspace = J2KGPU_CS_NONE; break;                    // bi-level, sRGB, grey, unknown
                        }
                    }
                    sp += sl;
                }
            }
            bp += bl;
        }
        if (!found) T2_FAIL(J2KGPU_E_RANGE, "JP2 file without a codestream box");
    }
    if (!d || len < 4 || be16(d) != SOC) T2_FAIL(J2KGPU_E_ARG, "not a codestream (no SOC)");
    uint64_t pos = 2;
    std::vector<uint32_t> tlm;                          // tile-part lengths announced by TLM, in codestream order
    // ---- main header: codestream.Parser.ReadHeader (parser.go:44-124) ----
    for (;;) {
        if (pos + 4 > len) T2_FAIL(J2KGPU_E_RANGE, "main header truncated");
        const uint32_t m = be16(d + pos);
        if (m == SOT) break;
        if (m == EOC) T2_FAIL(J2KGPU_E_RANGE, "no tile-part");
        const uint32_t L = be16(d + pos + 2);
        if (L < 2 || pos + 2 + L > len) T2_FAIL(J2KGPU_E_RANGE, "marker %04X: bad length", m);
        const uint8_t *seg = d + pos + 4;
        const uint32_t sl = L - 2;
        if (m == SIZ) {                                   // readSIZ parser.go:193-285
            if (sl < 36) T2_FAIL(J2KGPU_E_RANGE, "SIZ too short");
            const uint32_t xs = be32(seg + 2), ys = be32(seg + 6), xo = be32(seg + 10), yo = be32(seg + 14), xt = be32(seg + 18),
                           yt = be32(seg + 22), xto = be32(seg + 26), yto = be32(seg + 30), nc = be16(seg + 34);
            if (xo || yo || xto || yto) T2_FAIL(J2KGPU_E_UNSUPPORTED, "non-zero image / tile origin");
            if (nc < 1 || nc > 4 || sl < 36 + 3 * nc) T2_FAIL(J2KGPU_E_UNSUPPORTED, "unsupported number of components: %u", nc);   // decoder.go:585
            for (uint32_t c = 0; c < nc; c++) {
                const uint8_t *q = seg + 36 + 3 * c;
                if (q[1] != 1 || q[2] != 1) T2_FAIL(J2KGPU_E_UNSUPPORTED, "sub-sampled components");
                if (q[0] != seg[36]) T2_FAIL(J2KGPU_E_UNSUPPORTED, "components of different depth");
            }
            if (!xs || !ys || !xt || !yt || xs > 32768 || ys > 32768) T2_FAIL(J2KGPU_E_UNSUPPORTED, "image %ux%u", xs, ys);
            h.W = xs; h.H = ys; h.tile_w = xt; h.tile_h = yt; h.ncomp = nc; h.prec = (seg[36] & 0x7F) + 1; h.sgnd = seg[36] >> 7;
            if (h.prec > 16) T2_FAIL(J2KGPU_E_UNSUPPORTED, "precision %u", h.prec);
            h.have_siz = true;
        } else if (m == COD) {                            // readCOD parser.go:288-370
            if (sl < 10) T2_FAIL(J2KGPU_E_RANGE, "COD too short");
            const uint32_t scod = seg[0];
            h.sop = (scod >> 1) & 1; h.eph = (scod >> 2) & 1;
            h.prog = seg[1]; h.layers = be16(seg + 2); h.mct = seg[4]; h.nlevels = seg[5];
            h.cbw = 1u << (seg[6] + 2); h.cbh = 1u << (seg[7] + 2); h.style = seg[8]; h.reversible = seg[9] == 1;
            if (h.prog > 4) T2_FAIL(J2KGPU_E_RANGE, "progression order %u", h.prog);
            if (h.nlevels > 10) T2_FAIL(J2KGPU_E_UNSUPPORTED, "%u decomposition levels", h.nlevels);
            if (h.cbw > 64 || h.cbh > 64 || h.cbw * h.cbh > 4096) T2_FAIL(J2KGPU_E_UNSUPPORTED, "code blocks %ux%u", h.cbw, h.cbh);
            // RESET, VCAUSAL, PREDTERM, SEGSYM keep one codeword segment per block: the block decoder handles them;
            // BYPASS and TERMALL change the length signalling of the packet headers (B.10.7.2) and are refused
            if ((h.style & ~0x7Fu) || ((h.style & 0x40u) && (h.style & 0x3Fu)))
                T2_FAIL(J2KGPU_E_UNSUPPORTED, "code-block style %02X (HT with a classic style bit)", h.style);
            if (!h.layers) T2_FAIL(J2KGPU_E_RANGE, "zero quality layers");
            for (uint32_t r = 0; r <= 32; r++) h.ppx[r] = h.ppy[r] = 15;
            if (scod & 1) {                               // user-defined precincts: one byte per resolution, PPx | PPy << 4 (A.6.1)
                if (sl < 10 + h.nlevels + 1) T2_FAIL(J2KGPU_E_RANGE, "COD too short for its precinct sizes");
                for (uint32_t r = 0; r <= h.nlevels; r++) {
                    h.ppx[r] = seg[10 + r] & 15; h.ppy[r] = seg[10 + r] >> 4;
                    if (r && (h.ppx[r] == 0 || h.ppy[r] == 0)) T2_FAIL(J2KGPU_E_RANGE, "precinct size 1 above resolution 0");
                }
            }
            h.have_cod = true;
        } else if (m == CAP) {
            h.ht = 1;                                     // Part 15 capability (Header.IsHTJ2K header.go:241-257)
        } else if (m == QCD) {                            // readQCD parser.go:459-519
            if (sl < 1) T2_FAIL(J2KGPU_E_RANGE, "QCD too short");
            const uint32_t style = seg[0] & 31;
            h.guard = seg[0] >> 5;
            h.q.clear();
            if (style == 0) for (uint32_t i = 1; i < sl; i++) h.q.push_back({(uint32_t)seg[i] >> 3, 0u});
            else if (style == 1 || style == 2) for (uint32_t i = 1; i + 1 < sl; i += 2) h.q.push_back({be16(seg + i) >> 11, be16(seg + i) & 0x7FFu});
            else T2_FAIL(J2KGPU_E_RANGE, "quantisation style %u", style);
            if (style == 1) h.q.resize(1);
            h.have_qcd = true;
            if (style == 1) h.q.push_back({0xFFFFFFFFu, 0});   // marks "derived": expanded once nlevels is known
        } else if (m == TLM) {                            // parser.go:671-775: Stlm, then (Ttlm, Ptlm) pairs
            if (sl >= 2) {
                const uint32_t st = (seg[1] >> 4) & 3, sp = (seg[1] >> 6) & 1;
                const uint32_t esz = st + (sp ? 4 : 2);
                if (st <= 2) for (uint32_t i = 2; i + esz <= sl; i += esz) tlm.push_back(sp ? be32(seg + i + st) : be16(seg + i + st));
            }
        } else if (m == QCC) {                            // A.6.5: Cqcc (1 byte below 257 components, else 2), then as QCD
            if (!h.have_siz) T2_FAIL(J2KGPU_E_RANGE, "QCC before SIZ");
            const uint32_t cw = h.ncomp < 257 ? 1 : 2;
            if (sl < cw + 1) T2_FAIL(J2KGPU_E_RANGE, "QCC too short");
            const uint32_t c = cw == 1 ? seg[0] : be16(seg);
            if (c >= h.ncomp) T2_FAIL(J2KGPU_E_RANGE, "QCC for component %u of %u", c, h.ncomp);
            const uint8_t *qs = seg + cw;
            const uint32_t ql = sl - cw, style = qs[0] & 31;
            if (h.qc.empty()) h.qc.resize(h.ncomp);
            Header::QComp &qc = h.qc[c];
            qc.set = true; qc.guard = qs[0] >> 5; qc.q.clear();
            if (style == 0) for (uint32_t i = 1; i < ql; i++) qc.q.push_back({(uint32_t)qs[i] >> 3, 0u});
            else if (style == 1 || style == 2) for (uint32_t i = 1; i + 1 < ql; i += 2) qc.q.push_back({be16(qs + i) >> 11, be16(qs + i) & 0x7FFu});
            else T2_FAIL(J2KGPU_E_RANGE, "quantisation style %u", style);
            if (style == 1) { qc.q.resize(1); qc.q.push_back({0xFFFFFFFFu, 0}); }
        } else if (m == COC) {                            // A.6.2: Ccoc, Scoc, then SPcoc laid out like SPcod
            if (!h.have_siz) T2_FAIL(J2KGPU_E_RANGE, "COC before SIZ");
            const uint32_t cw = h.ncomp < 257 ? 1 : 2;
            if (sl < cw + 6) T2_FAIL(J2KGPU_E_RANGE, "COC too short");
            const uint32_t c = cw == 1 ? seg[0] : be16(seg);
            if (c >= h.ncomp) T2_FAIL(J2KGPU_E_RANGE, "COC for component %u of %u", c, h.ncomp);
            const uint8_t *sp = seg + cw + 1;
            if (h.cc.empty()) h.cc.resize(h.ncomp);
            Header::CComp &q = h.cc[c];
            q.set = true; q.nlevels = sp[0]; q.cbw = 1u << ((sp[1] & 15) + 2); q.cbh = 1u << ((sp[2] & 15) + 2); q.style = sp[3]; q.reversible = sp[4] == 1;
            if (q.nlevels > 10) T2_FAIL(J2KGPU_E_UNSUPPORTED, "%u decomposition levels", q.nlevels);
            for (uint32_t r = 0; r <= 32; r++) q.ppx[r] = q.ppy[r] = 15;
            if (seg[cw] & 1) {
                if (sl < cw + 6 + q.nlevels + 1) T2_FAIL(J2KGPU_E_RANGE, "COC too short for its precinct sizes");
                for (uint32_t r = 0; r <= q.nlevels; r++) {
                    q.ppx[r] = sp[5 + r] & 15; q.ppy[r] = sp[5 + r] >> 4;
                    if (r && (q.ppx[r] == 0 || q.ppy[r] == 0)) T2_FAIL(J2KGPU_E_RANGE, "precinct size 1 above resolution 0");
                }
            }
        } else if (m == RGN || m == POC || m == PPM || m == PLM) {
            T2_FAIL(J2KGPU_E_UNSUPPORTED, "marker %04X", m);
        }
        pos += 2 + L;
    }
    if (!h.have_siz || !h.have_cod || !h.have_qcd) T2_FAIL(J2KGPU_E_RANGE, "SIZ / COD / QCD missing");
    if (!h.cc.empty()) {                                  // COC: every component must end up with the same coding parameters
        Header::CComp cod;
        cod.nlevels = h.nlevels; cod.cbw = h.cbw; cod.cbh = h.cbh; cod.style = h.style; cod.reversible = h.reversible;
        memcpy(cod.ppx, h.ppx, sizeof cod.ppx); memcpy(cod.ppy, h.ppy, sizeof cod.ppy);
        const Header::CComp &e0 = h.cc[0].set ? h.cc[0] : cod;
        for (uint32_t c = 1; c < h.ncomp; c++) {
            const Header::CComp &e = h.cc[c].set ? h.cc[c] : cod;
            if (e.nlevels != e0.nlevels || e.cbw != e0.cbw || e.cbh != e0.cbh || e.style != e0.style || e.reversible != e0.reversible ||
                memcmp(e.ppx, e0.ppx, e0.nlevels + 1) || memcmp(e.ppy, e0.ppy, e0.nlevels + 1))
                T2_FAIL(J2KGPU_E_UNSUPPORTED, "COC: component %u is coded with other parameters than component 0", c);
        }
        if (e0.cbw > 64 || e0.cbh > 64 || e0.cbw * e0.cbh > 4096) T2_FAIL(J2KGPU_E_UNSUPPORTED, "code blocks %ux%u", e0.cbw, e0.cbh);
        if ((e0.style & ~0x7Fu) || ((e0.style & 0x40u) && (e0.style & 0x3Fu))) T2_FAIL(J2KGPU_E_UNSUPPORTED, "code-block style %02X", e0.style);
        h.nlevels = e0.nlevels; h.cbw = e0.cbw; h.cbh = e0.cbh; h.style = e0.style; h.reversible = e0.reversible;
        memcpy(h.ppx, e0.ppx, sizeof h.ppx); memcpy(h.ppy, e0.ppy, sizeof h.ppy);
    }
    if (h.style & 0x40) h.ht = 1;
    F->bands = band_list(h.nlevels);
    const std::vector<BandId> &bands = F->bands;
    if (h.q.size() == 2 && h.q[1].first == 0xFFFFFFFFu) {   // scalar derived (E-5): eps_b = eps_0 - N_L + n_b, mu_b = mu_0
        const auto q0 = h.q[0];
        h.q.clear();
        for (const BandId &b : bands) h.q.push_back({q0.first + b.lvl >= h.nlevels ? q0.first + b.lvl - h.nlevels : 0u, q0.second});
    }
    if (h.q.size() < bands.size()) T2_FAIL(J2KGPU_E_RANGE, "QCD lists %zu bands, %zu needed", h.q.size(), bands.size());
    for (Header::QComp &qc : h.qc) {
        if (!qc.set) continue;
        if (qc.q.size() == 2 && qc.q[1].first == 0xFFFFFFFFu) {
            const auto q0 = qc.q[0];
            qc.q.clear();
            for (const BandId &b : bands) qc.q.push_back({q0.first + b.lvl >= h.nlevels ? q0.first + b.lvl - h.nlevels : 0u, q0.second});
        }
        if (qc.q.size() < bands.size()) T2_FAIL(J2KGPU_E_RANGE, "QCC lists %zu bands, %zu needed", qc.q.size(), bands.size());
    }
    if (reduce > h.nlevels) T2_FAIL(J2KGPU_E_ARG, "reduce %u > %u decomposition levels", reduce, h.nlevels);
    const uint32_t ntx = cdiv(h.W, h.tile_w), nty = cdiv(h.H, h.tile_h);
    if ((uint64_t)ntx * nty > 65535) T2_FAIL(J2KGPU_E_UNSUPPORTED, "too many tiles");
    if (ntx * nty > 1 && ((h.tile_w | h.tile_h) & ((1u << h.nlevels) - 1)))
        T2_FAIL(J2KGPU_E_UNSUPPORTED, "tile size %ux%u is not a multiple of 2^nlevels", h.tile_w, h.tile_h);
    // ---- tile-part index: ReadTilePartHeader (parser.go:894-982), hopping by Psot ----
    const uint32_t ntiles = ntx * nty;
    std::vector<std::vector<TilePart>> &parts = F->parts;
    std::vector<std::vector<uint32_t>> &plt = F->plt;
    parts.assign(ntiles, {}); plt.assign(ntiles, {});
    uint32_t ntp = 0;
    int64_t truncated_tile = -1;
    while (pos + 12 <= len && be16(d + pos) == SOT) {
        const uint32_t isot = be16(d + pos + 4), psot = be32(d + pos + 6);
        if (isot >= ntiles) T2_FAIL(J2KGPU_E_RANGE, "tile index %u out of range", isot);
        uint64_t end = psot ? pos + psot : len - 2;
        if (end > len) { end = len; truncated_tile = (int64_t)isot; }     // the file ends inside this tile-part: decode what is there
        if (end < pos + 14) T2_FAIL(J2KGPU_E_RANGE, "tile-part %u: bad Psot", ntp);
        if (ntp < tlm.size() && psot && tlm[ntp] != psot) T2_FAIL(J2KGPU_E_RANGE, "tile-part %u: TLM says %u bytes, Psot %u", ntp, tlm[ntp], psot);
        uint64_t p = pos + 12;
        for (;;) {
            if (p + 2 > end) T2_FAIL(J2KGPU_E_RANGE, "tile-part %u: no SOD", ntp);
            const uint32_t m = be16(d + p);
            if (m == SOD) break;
            if (p + 4 > end) T2_FAIL(J2KGPU_E_RANGE, "tile-part header truncated");
            const uint32_t L = be16(d + p + 2);
            if (L < 2 || p + 2 + L > end) T2_FAIL(J2KGPU_E_RANGE, "tile-part marker %04X: bad length", m);
            if (m == PLT) {                               // packet lengths, 7 bits per byte, MSB = continuation
                uint32_t v = 0;
                for (uint32_t i = 5; i < 2 + L; i++) {
                    const uint8_t b = d[p + i];
                    v = (v << 7) | (b & 0x7F);
                    if (!(b & 0x80)) { plt[isot].push_back(v); v = 0; }
                }
            } else if (m == COD || m == COC || m == QCD || m == QCC || m == RGN || m == POC || m == PPT) {
                T2_FAIL(J2KGPU_E_UNSUPPORTED, "marker %04X in a tile-part header", m);
            }
            p += 2 + L;
        }
        parts[isot].push_back({p + 2, end});
        pos = end;
        ntp++;
    }
    F->d = d; F->len = len; F->reduce = reduce; F->ntiles = ntiles; F->ntp = ntp; F->ntlm = (uint32_t)tlm.size();
    F->truncated_tile = truncated_tile;
    F->touts.assign(ntiles, TileOut());
    *fout = guard.release();
    return J2KGPU_OK;
#undef T2_FAIL
}

uint32_t j2k_tier2_tiles(const j2k_t2_frame *F) { return F ? F->ntiles : 0; }
void j2k_tier2_free(j2k_t2_frame *F) { delete F; }

// tier-2 of tile t (callable concurrently for distinct tiles)
void j2k_tier2_tile(j2k_t2_frame *F, uint32_t t)
{
    if (t >= F->ntiles) return;
    if (F->parts[t].empty()) { std::vector<TilePart> none{{0, 0}}; parse_tile(F->d, F->h, t, none, F->plt[t], F->reduce, false, F->touts[t]); }
    else parse_tile(F->d, F->h, t, F->parts[t], F->plt[t], F->reduce, (int64_t)t == F->truncated_tile, F->touts[t]);
}

// every tile parsed: merge into the flat tables; the frame is freed
int j2k_tier2_finish(j2k_t2_frame *F, j2kgpu_parsed &out)
{
    std::unique_ptr<j2k_t2_frame> guard(F);
    const Header &h = F->h;
    const uint8_t *d = F->d;
    const uint64_t len = F->len;
    const uint32_t ntiles = F->ntiles, reduce = F->reduce;
    std::vector<TileOut> &touts = F->touts;
    const std::vector<BandId> &bands = F->bands;
    // ---- merge ----
    uint64_t extra_total = 0;
    size_t ncb = 0;
    for (uint32_t t = 0; t < ntiles; t++) {
        if (touts[t].err.code) { out.err = touts[t].err.msg; return touts[t].err.code; }
        extra_total += touts[t].extra.size();
        ncb += touts[t].cbs.size();
    }
    out.tilecomps.clear(); out.cblks.clear(); out.owned.clear();
    out.cblks.reserve(ncb);
    if (extra_total) { out.owned.reserve(len + extra_total + 8); out.owned.assign(d, d + len); }
    for (uint32_t t = 0; t < ntiles; t++) {
        TileOut &to = touts[t];
        const uint32_t tc_base = (uint32_t)out.tilecomps.size();
        const uint64_t ex_base = out.owned.size();
        out.tilecomps.insert(out.tilecomps.end(), to.tcs.begin(), to.tcs.end());
        if (!to.extra.empty()) out.owned.insert(out.owned.end(), to.extra.begin(), to.extra.end());
        for (size_t i = 0; i < to.cbs.size(); i++) {
            j2k_cblk_t cb = to.cbs[i];
            cb.tilecomp += tc_base;
            if (to.is_extra[i]) cb.data_off += ex_base;
            out.cblks.push_back(cb);
        }
        out.packets += to.packets; out.plt_packets += to.plt_checked;
    }
    if (extra_total) { out.blob = out.owned.data(); out.blob_len = out.owned.size(); }
    else { out.blob = d; out.blob_len = len; }
    j2k_image_t &im = out.image;
    memset(&im, 0, sizeof im);
    im.width = (uint32_t)cdiv64(h.W, (int64_t)1 << reduce); im.height = (uint32_t)cdiv64(h.H, (int64_t)1 << reduce);
    im.ncomp = (uint16_t)h.ncomp;
    for (uint32_t c = 0; c < h.ncomp; c++) { im.prec[c] = (uint8_t)h.prec; im.sgnd[c] = (uint8_t)h.sgnd; }
    im.mct = (h.mct && h.ncomp >= 3) ? 1 : 0; im.reversible = (uint8_t)h.reversible; im.nlevels = (uint8_t)(h.nlevels - reduce);
    im.ht = (uint8_t)h.ht; im.mode = J2KGPU_MODE_ISO; im.out_fmt = J2KGPU_FMT_AUTO;
    uint32_t cbits = 0;
    for (uint32_t c = 0; c < h.ncomp; c++)
        for (size_t bi = 0; bi < bands.size(); bi++) cbits = std::max(cbits, h.q_of(c)[bi].first + h.guard_of(c) - 1);
    im.coef_bits = (uint8_t)std::min(cbits, 255u);
    im.cblk_style = h.ht ? 0 : (uint8_t)(h.style & 0x3Fu);
    im.colorspace = (uint8_t)h.colorspace;
    out.layers = h.layers; out.tiles = ntiles; out.tile_parts = F->ntp; out.progression = h.prog; out.tlm_tile_parts = F->ntlm;
    return J2KGPU_OK;
}

int j2k_tier2_parse(const uint8_t *d, uint64_t len, uint32_t reduce, uint32_t threads, j2kgpu_parsed &out)
{
    j2k_t2_frame *F = nullptr;
    const int rc = j2k_tier2_begin(d, len, reduce, &F, out.err);
    if (rc) return rc;
    const uint32_t ntiles = F->ntiles;
    uint32_t nthreads = threads ? threads : std::max(1u, std::thread::hardware_concurrency());
    nthreads = std::min<uint32_t>({nthreads, ntiles, 64u});
    std::atomic<uint32_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const uint32_t t = next.fetch_add(1);
            if (t >= ntiles) break;
            j2k_tier2_tile(F, t);
        }
    };
    if (nthreads <= 1) worker();
    else {
        std::vector<std::thread> th;
        for (uint32_t i = 0; i < nthreads; i++) th.emplace_back(worker);
        for (auto &t : th) t.join();
    }
    return j2k_tier2_finish(F, out);
}
