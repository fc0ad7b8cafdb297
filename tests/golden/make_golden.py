#!/usr/bin/env python3
"""Regenerates tests/golden/entropy_ref.npz from the oracle (oracle/liboracle.so).

The reference is Go and cannot run in this image, and it ships no fixtures, so the
golden vectors are the ORACLE's outputs on the reference's own test patterns, frozen so that
(a) the oracle cannot drift and (b) the GPU parity tests have fixed byte strings to decode.
Patterns: t1_test.go:15-40, coverage_test.go:414-446/464-534/814-841/883-901/1058-1085,
ht_test.go:29-37, bench_512_test.go:13-17 (first block), plus truncated / garbage streams.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import oracle_lib as O  # noqa: E402

cases = []  # (kind, w, h, band, src or None, bytes or None, nbps)


def t1(src, w, h, band):
    cases.append((0, w, h, band, np.asarray(src, np.int32), None, None))


def ht(src, w, h):
    cases.append((1, w, h, 0, np.asarray(src, np.int32), None, None))


t1(list(range(1, 17)), 4, 4, 0)
t1([-1, 2, -3, 4, 5, -6, 7, -8, -9, 10, -11, 12, 13, -14, 15, -16], 4, 4, 1)
t1([1, -1, 1, -1, -1, 1, -1, 1, 1, -1, 1, -1, -1, 1, -1, 1], 4, 4, 3)
t1([i * 2 for i in range(64)], 8, 8, 0)
for band in range(4):
    t1([(-(i % 128) if i % 3 == 0 else i % 128) for i in range(1024)], 32, 32, band)
t1([42], 1, 1, 0)
t1(list(range(1, 9)), 8, 1, 0)
t1(list(range(1, 9)), 1, 8, 0)
t1(list(range(1, 41)), 8, 5, 0)
t1([(-((i * 17) % 512) if i % 7 == 0 else (i * 17) % 512) for i in range(4096)], 64, 64, 3)
t1([-(i + 1) for i in range(256)], 16, 16, 0)
sp = np.zeros(1024, np.int32); sp[0], sp[100], sp[500], sp[900] = 100, -50, 200, -150
t1(sp, 32, 32, 0)
rng = np.random.default_rng(42)
t1(rng.integers(0, 256, 4096), 64, 64, 2)
t1(rng.integers(-2048, 2048, 64 * 37), 64, 37, 1)
for sz in (4, 8, 16, 32, 64):
    ht([((i % 256) - 128) * 4 if i % 7 == 0 else 0 for i in range(sz * sz)], sz, sz)
ht(rng.integers(-300, 300, 64 * 64), 64, 64)
ht(rng.integers(-5, 5, 48 * 20), 48, 20)

out = {}
n = 0
for kind, w, h, band, src, _, _ in cases:
    if kind == 0:
        data, nbps = O.t1_encode(src, w, h, band)
        dec = O.t1_decode(data, w, h, nbps, band)
    else:
        data, nbps = O.ht_encode(src, w, h), 0
        dec = O.ht_decode(data, w, h)
    out["meta_%d" % n] = np.array([w, h, band, nbps, kind], np.int32)
    out["bytes_%d" % n] = np.frombuffer(data, np.uint8)
    out["src_%d" % n] = src.reshape(-1)
    out["out_%d" % n] = dec
    n += 1
    # a truncated and a corrupted variant of the same stream (decoder-only cases)
    if len(data) > 8:
        for variant in (data[: len(data) // 2], bytes(b ^ 0x5A if i % 5 == 0 else b for i, b in enumerate(data))):
            dec = O.t1_decode(variant, w, h, nbps, band) if kind == 0 else O.ht_decode(variant, w, h)
            out["meta_%d" % n] = np.array([w, h, band, nbps, kind], np.int32)
            out["bytes_%d" % n] = np.frombuffer(variant, np.uint8)
            out["out_%d" % n] = dec
            n += 1
out["n_cases"] = np.array(n)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "entropy_ref.npz")
np.savez_compressed(path, **out)
print("wrote", path, n, "cases", os.path.getsize(path), "bytes")
