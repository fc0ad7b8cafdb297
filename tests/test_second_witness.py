"""CPU: the second witness (tests/go_witness.py, a Python transliteration written from the Go source alone) against the C
restatement in oracle/ (orc_ht.c, orc_t1.c) that the GPU kernels are checked against -- differential, 10^4 random streams:
encoder-made streams (the reference's own encoders, restated in datagen/), truncated and corrupted ones, and pure garbage.
For the reference's HT coder, whose results no reference test pins, two independent readings of ht.go that
agree on every byte string are the strongest pin this image allows (no Go toolchain)."""
import multiprocessing as mp
import os

import numpy as np

import go_witness as G
import oracle_lib as O


def _ht_case(seed):
    rng = np.random.default_rng(seed)
    bad = []
    for k in range(125):
        w, h = int(rng.integers(1, 33)), int(rng.integers(1, 33))
        kind = int(rng.integers(0, 4))
        if kind < 2:                                            # a stream of the reference's HT encoder, maybe damaged
            nb = int(rng.integers(1, 12))
            d = rng.integers(-(1 << nb) + 1, 1 << nb, w * h).astype(np.int32)
            d[rng.random(w * h) < rng.uniform(0, 0.9)] = 0
            try:
                data = bytearray(O.ht_encode(d, w, h, 0))
            except Exception:
                continue
            if kind == 1 and len(data) > 2:
                for _ in range(int(rng.integers(1, 6))):
                    data[int(rng.integers(0, len(data)))] = int(rng.integers(0, 256))
        else:                                                   # garbage with a plausible trailer, 0xFF-heavy half of the time
            n = int(rng.integers(0, 300))
            a = rng.integers(0, 256, n).astype(np.uint8)
            if kind == 3:
                a[rng.random(n) < 0.3] = 0xFF
            if n >= 2 and rng.random() < 0.8:
                scup = int(rng.integers(2, min(n, 4095) + 1))
                a[-1], a[-2] = scup & 0xFF, (a[-2] & 0xF0) | ((scup >> 8) & 0x0F)
            data = bytearray(a.tobytes())
        want = O.ht_decode(bytes(data), w, h)
        got = np.array(G.ht_decode(bytes(data), w, h), np.int64).astype(np.int32)
        if not np.array_equal(got, want):
            bad.append((seed, k, w, h, kind))
    return 125, bad


def _t1_case(seed):
    rng = np.random.default_rng(seed)
    bad = []
    for k in range(40):
        w, h = int(rng.integers(1, 13)), int(rng.integers(1, 13))
        band = int(rng.integers(0, 4))
        kind = int(rng.integers(0, 3))
        nb = int(rng.integers(1, 9))
        if kind < 2:
            d = rng.integers(-(1 << nb) + 1, 1 << nb, w * h).astype(np.int32)
            d[rng.random(w * h) < rng.uniform(0, 0.8)] = 0
            data, nbps = O.t1_encode(d, w, h, band)
            data = bytearray(data)
            if kind == 1 and len(data) > 1:
                data = data[: int(rng.integers(0, len(data)))]                       # truncated
                for _ in range(int(rng.integers(0, 3))):
                    if data:
                        data[int(rng.integers(0, len(data)))] = int(rng.integers(0, 256))
            nbps = max(nbps, 1)
        else:
            data = bytearray(rng.integers(0, 256, int(rng.integers(0, 40))).astype(np.uint8).tobytes())
            nbps = nb
        want = O.t1_decode(bytes(data), w, h, nbps, band)
        got = np.array(G.t1_decode(bytes(data), w, h, nbps, band), np.int64).astype(np.int32)
        if not np.array_equal(got, want):
            bad.append((seed, k, w, h, band, kind))
    return 40, bad


def _run(fn, seeds):
    with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(fn, seeds, chunksize=1)
    return sum(r[0] for r in res), [b for r in res for b in r[1]]


def test_ht_witness_agrees_with_c_restatement_on_7000_streams():
    n, bad = _run(_ht_case, range(9000, 9056))
    assert n == 7000 and not bad, bad[:5]


def test_t1_witness_agrees_with_c_restatement_on_3000_streams():
    n, bad = _run(_t1_case, range(7000, 7075))
    assert n == 3000 and not bad, bad[:5]


def test_witness_tables():
    """the witness's own ZC table (from t1_luts.go's rules) equals the C restatement's; MQ states are 47 x 2 with UNI at 92"""
    lut = np.ctypeslib.as_array(O.lib().orc_t1_zc_lut(), shape=(1024,))
    assert list(lut) == G.LUT_ZC
    assert len(G.MQ_STATES) == 94 and G.MQ_STATES[92][2] == 92 and G.MQ_STATES[93][3] == 93
