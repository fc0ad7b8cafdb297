"""CPU: the forward-path checker (oracle/orc_enc.c) is pinned the way the reference pins its encoder -- by exact round trips
(t1_test.go / dwt_test.go / mct_test.go: decode(encode(x)) == x) through the decoder-side oracle, which the golden vectors
pin -- plus a frozen digest of one tile, so that a change of either side of the checker shows."""
import hashlib

import numpy as np
import pytest

import oracle_lib as O
from enc_cases import CASES, go_image


def params(case):
    w, h, nc, bits, ll, nr, cbx, cby, q, prec = case
    return O.EncodeParams(width=w, height=h, ncomp=nc, pix_bits=bits, precision=prec, lossless=ll, num_resolutions=nr,
                          cb_x=cbx, cb_y=cby, quality=q)


def block_list(p):
    """encodeTile's job list (encoder.go:615-673), restated once more in Python"""
    num_res = p.num_resolutions or 6
    cbw, cbh = 1 << (p.cb_x + 2), 1 << (p.cb_y + 2)
    out = []
    for c in range(p.ncomp):
        for r in range(num_res):
            for b in range(1 if r == 0 else 3):
                band = 0 if r == 0 else b + 1
                scale = 1 << (num_res - 1 - r)
                bw, bh = (p.width + scale - 1) // scale, (p.height + scale - 1) // scale
                if r > 0:
                    bw, bh = (bw + 1) // 2, (bh + 1) // 2
                for cby in range((bh + cbh - 1) // cbh):
                    for cbx in range((bw + cbw - 1) // cbw):
                        out.append((c, cbx * cbw, cby * cbh, min(cbw, bw - cbx * cbw), min(cbh, bh - cby * cbh), band))
    return out


@pytest.mark.parametrize("case", CASES)
def test_blocks_decode_back_to_the_planes(case):
    p = params(case)
    pix = go_image(p.width, p.height, p.ncomp, p.pix_bits, seed=p.width + p.height)
    planes = O.encode_preprocess(p, pix)
    data, lens, bps = O.encode_tile(p, pix)
    blocks = block_list(p)
    assert len(blocks) == O.encode_block_count(p) == len(lens)
    assert int(lens.sum()) == len(data)
    off = 0
    for (c, sx, sy, w, h, band), n, nb in zip(blocks, lens, bps):
        want = np.zeros((h, w), np.int32)                          # extractCodeBlockData, encoder.go:778-793
        src = planes[c, sy:sy + h, sx:sx + w]
        want[:src.shape[0], :src.shape[1]] = src
        got = O.t1_decode(data[off:off + n], w, h, int(nb), band).reshape(h, w)
        assert np.array_equal(got, want), (c, sx, sy, w, h, band)
        assert (n == 0) == (not want.any())
        off += int(n)


@pytest.mark.parametrize("case", [c for c in CASES if c[4] == 1 and c[9] == 0])
def test_lossless_planes_invert_to_the_pixels(case):
    p = params(case)
    pix = go_image(p.width, p.height, p.ncomp, p.pix_bits, seed=3)
    planes = O.encode_preprocess(p, pix)
    levels = p.num_resolutions - 1 if p.num_resolutions > 1 else 5
    comps = [O.reconstruct53(planes[c], p.width, p.height, levels) for c in range(p.ncomp)]
    if p.ncomp >= 3:
        comps[:3] = O.inv_rct(*comps[:3])
    comps = [O.dc_shift_inverse(c, p.pix_bits) for c in comps]
    ch = 1 if p.ncomp == 1 else 4
    a = pix.reshape(p.height, p.width, ch, p.pix_bits // 8).astype(np.int32)
    src = a[..., 0] if p.pix_bits == 8 else (a[..., 0] << 8) | a[..., 1]
    for c in range(p.ncomp):
        assert np.array_equal(comps[c].reshape(p.height, p.width), src[:, :, c])


def test_frozen_digest():
    p = params(CASES[0])
    data, lens, bps = O.encode_tile(p, go_image(96, 64, 3, 8, seed=1))
    digest = hashlib.sha256(data.tobytes() + lens.tobytes() + bps.tobytes()).hexdigest()
    assert digest == FROZEN, digest


FROZEN = "9c92a3286d026c0f301176f080922f2708af3821d6bcf61e41e3ddf48abd6b0e"


def test_product_block_count_without_a_gpu(j2k):
    """j2kgpu_encode_block_count is host logic (encodeTile's job list, encoder.go:615-673): same count as the checker and as the
    Python restatement above, for the fixed cases and a sweep of sizes / block shapes; 0 for options the library refuses"""
    import ctypes as C
    L = j2k.lib()
    rng = np.random.default_rng(9)
    cases = list(CASES) + [(int(rng.integers(1, 5000)), int(rng.integers(1, 5000)), int(rng.choice([1, 3, 4])), 8, 1,
                            int(rng.integers(0, 9)), int(rng.integers(0, 7)), int(rng.integers(0, 7)), 0, 0) for _ in range(40)]
    for case in cases:
        p = params(case)
        q = j2k.EncodeParams(width=p.width, height=p.height, ncomp=p.ncomp, pix_bits=p.pix_bits, lossless=p.lossless,
                             num_resolutions=p.num_resolutions, cb_x=p.cb_x, cb_y=p.cb_y)
        assert L.j2kgpu_encode_block_count(C.byref(q)) == O.encode_block_count(p) == len(block_list(p)), case
    for bad in (dict(width=0, height=4, ncomp=1), dict(width=4, height=4, ncomp=2), dict(width=4, height=4, ncomp=1, cb_x=7)):
        assert L.j2kgpu_encode_block_count(C.byref(j2k.EncodeParams(pix_bits=8, **bad))) == 0
