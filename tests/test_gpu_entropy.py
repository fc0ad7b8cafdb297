"""GPU parity, entropy stage: j2kgpu_t1_decode_blocks / j2kgpu_ht_decode_blocks (C ABI) against the oracle's
T1.Decode / HTDecoder.Decode restatement on the same bytes -- bit-exact.  Cases: the reference's own test
patterns (frozen in tests/golden/entropy_ref.npz), random blocks of every shape, truncated / garbage streams
(the reference fuzz contract, fuzz_test.go:9-70)."""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "entropy_ref.npz")


def test_golden_vectors_bit_exact(gpu_ctx):
    g = np.load(GOLD)
    t1, ht, want_t1, want_ht = [], [], [], []
    for i in range(int(g["n_cases"])):
        w, h, band, nbps, kind = (int(v) for v in g["meta_%d" % i])
        data = g["bytes_%d" % i].tobytes()
        (t1 if kind == 0 else ht).append((data, w, h, nbps, band))
        (want_t1 if kind == 0 else want_ht).append(g["out_%d" % i])
    for got, want in zip(gpu_ctx.t1_decode_blocks(t1), want_t1):
        assert np.array_equal(got, want)
    for got, want in zip(gpu_ctx.ht_decode_blocks(ht), want_ht):
        assert np.array_equal(got, want)


def test_t1_reference_roundtrip_cases(gpu_ctx):
    """t1_test.go:7-94 and coverage_test.go:464-534, 814-841: GPU Decode(Encode(x)) == x"""
    cases = [
        (4, 4, 0, list(range(1, 17))),
        (4, 4, 1, [-1, 2, -3, 4, 5, -6, 7, -8, -9, 10, -11, 12, 13, -14, 15, -16]),
        (4, 4, 3, [1, -1, 1, -1, -1, 1, -1, 1, 1, -1, 1, -1, -1, 1, -1, 1]),
        (8, 8, 0, [i * 2 for i in range(64)]),
        (1, 1, 0, [42]), (8, 1, 0, list(range(1, 9))), (1, 8, 0, list(range(1, 9))), (8, 5, 0, list(range(1, 41))),
        (64, 64, 3, [(-((i * 17) % 512) if i % 7 == 0 else (i * 17) % 512) for i in range(4096)]),
        (16, 16, 0, [-(i + 1) for i in range(256)]),
    ]
    blocks = []
    for w, h, band, d in cases:
        enc, nbps = O.t1_encode(d, w, h, band)
        blocks.append((enc, w, h, nbps, band))
    for got, (w, h, band, d) in zip(gpu_ctx.t1_decode_blocks(blocks), cases):
        assert got.tolist() == d


def test_t1_random_blocks_vs_oracle(gpu_ctx):
    rng = np.random.default_rng(21)
    blocks, want = [], []
    for _ in range(300):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        band = int(rng.integers(0, 4))
        nb = int(rng.integers(1, 17))
        d = rng.integers(-(1 << nb) + 1, 1 << nb, w * h).astype(np.int32)
        d[rng.random(w * h) < rng.uniform(0, 0.95)] = 0
        enc, nbps = O.t1_encode(d, w, h, band)
        blocks.append((enc, w, h, nbps, band))
        want.append(d if d.any() else O.t1_decode(enc, w, h, nbps, band))
    for i, (got, w_) in enumerate(zip(gpu_ctx.t1_decode_blocks(blocks), want)):
        assert np.array_equal(got, w_), i


def test_t1_garbage_and_truncated_vs_oracle(gpu_ctx):
    """FuzzT1Decode: arbitrary bytes, sizes 4..64, numBPS 8 -- must not fault and must equal the oracle"""
    rng = np.random.default_rng(22)
    seeds = [b"", b"\x00", b"\xff", b"\x00\x01\x02\x03", b"\xff\xff\xff\xff", bytes(range(16)), b"\xff\x90\x00\xff"]
    seeds += [rng.integers(0, 256, int(rng.integers(1, 400))).astype(np.uint8).tobytes() for _ in range(60)]
    blocks = []
    for i, s in enumerate(seeds):
        sz = [4, 8, 16, 32, 64][i % 5]
        blocks.append((s, sz, sz, 8, i % 4))
        blocks.append((s, int(rng.integers(1, 65)), int(rng.integers(1, 65)), int(rng.integers(1, 14)), i % 4))
    for (s, w, h, nbps, band), got in zip(blocks, gpu_ctx.t1_decode_blocks(blocks)):
        assert np.array_equal(got, O.t1_decode(s, w, h, nbps, band))


@pytest.mark.parametrize("group", [4, 8, 16, 32])
def test_t1_every_lanes_per_block_variant(gpu_ctx, group):
    """k_t1_ref<OT, G>: 32 / G blocks share a warp (one MQ chain per group of G lanes, finished bit-planes merged into the
    samples in place): every G against the oracle on encoder streams of all shapes, garbage, and a block count that
    leaves the last warp partly empty"""
    rng = np.random.default_rng(40 + group)
    blocks, want = [], []
    for k in range(75):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        band, nb = int(rng.integers(0, 4)), int(rng.integers(1, 15))
        if k % 5 == 4:                                               # arbitrary bytes
            s = rng.integers(0, 256, int(rng.integers(1, 300))).astype(np.uint8).tobytes()
            blocks.append((s, w, h, nb, band))
            want.append(O.t1_decode(s, w, h, nb, band))
            continue
        d = rng.integers(-(1 << nb) + 1, 1 << nb, w * h).astype(np.int32)
        d[rng.random(w * h) < rng.uniform(0, 0.95)] = 0
        enc, nbps = O.t1_encode(d, w, h, band)
        blocks.append((enc, w, h, nbps, band))
        want.append(d if d.any() else O.t1_decode(enc, w, h, nbps, band))
    with gpu_ctx.options(t1_group=group):
        got = gpu_ctx.t1_decode_blocks(blocks)
    for i, (g, w_) in enumerate(zip(got, want)):
        assert np.array_equal(g, w_), i


def test_t1_wide_dynamic_range(gpu_ctx):
    """num_bps up to 31 (int32 magnitudes)"""
    rng = np.random.default_rng(23)
    blocks, want = [], []
    for nb in (17, 24, 30, 31):
        d = rng.integers(-(1 << nb) + 1, 1 << nb, 16 * 12).astype(np.int64).astype(np.int32)
        enc, nbps = O.t1_encode(d, 16, 12, 2)
        assert nbps == nb
        blocks.append((enc, 16, 12, nbps, 2))
        want.append(d)
    for got, w_ in zip(gpu_ctx.t1_decode_blocks(blocks), want):
        assert np.array_equal(got, w_)


def test_ht_encoder_streams_vs_oracle(gpu_ctx):
    """ht_test.go:29-37 pattern and random data through HTEncoder.Encode, decoded by both"""
    rng = np.random.default_rng(24)
    blocks = []
    for sz in (4, 8, 16, 32, 64):
        d = np.array([((i % 256) - 128) * 4 if i % 7 == 0 else 0 for i in range(sz * sz)], np.int32)
        blocks.append((O.ht_encode(d, sz, sz), sz, sz, 0, 0))
    for _ in range(120):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        nb = int(rng.integers(1, 12))
        d = rng.integers(-(1 << nb) + 1, 1 << nb, w * h).astype(np.int32)
        d[rng.random(w * h) < rng.uniform(0, 0.9)] = 0
        try:
            blocks.append((O.ht_encode(d, w, h), w, h, 0, 0))
        except OverflowError:
            pass
    for (s, w, h, _, _), got in zip(blocks, gpu_ctx.ht_decode_blocks(blocks)):
        assert np.array_equal(got, O.ht_decode(s, w, h))


def test_ht_garbage_vs_oracle(gpu_ctx):
    """FuzzHTDecode: arbitrary bytes; most with a plausible scup so the cleanup body actually runs"""
    rng = np.random.default_rng(25)
    blocks = [(b"", 8, 8, 0, 0), (b"\x01", 8, 8, 0, 0), (b"\x00\x00\x00\x01", 8, 8, 0, 0), (b"\x00\x00\x0f\xff", 8, 8, 0, 0)]
    for _ in range(400):
        n = int(rng.integers(2, 700))
        s = rng.integers(0, 256, n).astype(np.uint8)
        if rng.random() < 0.5:
            s[rng.random(n) < 0.3] = 0xFF                # stress the unstuffing paths
        if rng.random() < 0.8:
            scup = int(rng.integers(2, min(n, 4095) + 1))
            s[-1] = scup & 0xFF
            s[-2] = (s[-2] & 0xF0) | (scup >> 8)
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        blocks.append((s.tobytes(), w, h, 0, 0))
    for i, ((s, w, h, _, _), got) in enumerate(zip(blocks, gpu_ctx.ht_decode_blocks(blocks))):
        assert np.array_equal(got, O.ht_decode(s, w, h)), i


def test_ht_corrupted_magsgn_vs_oracle(gpu_ctx):
    """valid VLC segments (so the U-VLC values stay small and the blocks take the two-kernel path) over damaged MagSgn
    segments: 0xFF runs (stuffing), random bytes, truncation (the exhausted stream continues with ones), empty"""
    rng = np.random.default_rng(26)
    blocks = []
    for _ in range(300):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        nb = int(rng.integers(1, 14))
        d = rng.integers(-(1 << nb) + 1, 1 << nb, w * h).astype(np.int32)
        d[rng.random(w * h) < rng.uniform(0, 0.9)] = 0
        try:
            s = np.frombuffer(O.ht_encode(d, w, h), np.uint8).copy()
        except OverflowError:
            continue
        if s.size < 2:
            continue
        scup = int(s[-1]) + ((int(s[-2]) & 0x0F) << 8)
        L = s.size - scup
        if L > 0:
            ms, tail = s[:L].copy(), s[L:]
            mode = int(rng.integers(0, 4))
            if mode == 0:
                ms[rng.random(L) < 0.3] = 0xFF
            elif mode == 1:
                ms = rng.integers(0, 256, L).astype(np.uint8)
            elif mode == 2:
                ms = ms[: int(rng.integers(0, L))]            # scup stays: Lcup shrinks with the segment
            else:
                ms = np.concatenate([ms[: L // 2], np.full(L - L // 2, 0xFF, np.uint8)])
            s = np.concatenate([ms, tail])
        blocks.append((s.tobytes(), w, h, 0, 0))
    assert len(blocks) > 200
    for i, ((s, w, h, _, _), got) in enumerate(zip(blocks, gpu_ctx.ht_decode_blocks(blocks))):
        assert np.array_equal(got, O.ht_decode(s, w, h)), i


def test_block_argument_errors(gpu_ctx, j2k):
    with pytest.raises(j2k.J2KError) as e:
        gpu_ctx.t1_decode_blocks([(b"\x00", 65, 4, 8, 0)])
    assert e.value.code == j2k.E_UNSUPPORTED
    with pytest.raises(j2k.J2KError) as e:
        gpu_ctx.t1_decode_blocks([(b"\x00", 4, 4, 40, 0)])
    assert e.value.code == j2k.E_UNSUPPORTED
    assert gpu_ctx.t1_decode_blocks([]) == []
