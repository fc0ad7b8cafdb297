"""CPU: the product's codestream front door (go-jpeg2000_b200/host/tier2.cpp behind j2kgpu_parse_codestream) -- main header,
tile-part index, tier-2 packet headers -> the job tables of include/j2kgpu.h.  No device is needed for parsing.

Pins: (1) its tables equal, block for block and byte for byte, the tables of the harness's independent Python tier-2
(datagen/codestream.py) on codestreams written by OpenJPEG 2.5.4 and by our HTJ2K writer; (2) fed to the CPU checker
(oracle/iso_path.c) those tables reproduce OpenJPEG's own decode of the same bytes -- so the parser is pinned by an
independent decoder, not only by a second reading of the standard; (3) progression orders, SOP / EPH, PLT and TLM marker
cross-checks, several tile-parts per tile, ReduceResolution; (4) the reference's fuzz contract (fuzz_test.go: never
panic): truncations and byte flips return an error code or decode, never fault."""
import io

import numpy as np
import pytest

import oracle_lib as O
from datagen import codestream as cs, jobs

PIL_Image = pytest.importorskip("PIL.Image")


def opj_decode(data, reduce=0):
    im = PIL_Image.open(io.BytesIO(data))
    if reduce:
        im.reduce = reduce
    im.load()
    return np.array(im)


def opj_encode(s, **kw):
    a = np.moveaxis(s, 0, 2).astype(np.uint8) if s.shape[0] == 3 else s[0].astype(np.uint8)
    buf = io.BytesIO()
    PIL_Image.fromarray(a).save(buf, format="JPEG2000", no_jp2=True, **kw)
    return buf.getvalue()


def job_from_parsed(p, data):
    """Parsed (C++ tier-2) -> the job dict O.iso_decode_job takes"""
    im = p.image
    tcs, cbs, blob = p.tables()
    return dict(width=im.width, height=im.height, ncomp=im.ncomp, prec=im.prec[0], sgnd=im.sgnd[0], mct=im.mct,
                reversible=im.reversible, nlevels=im.nlevels, ht=im.ht, mode=1,
                tilecomps=np.frombuffer(tcs.tobytes(), jobs.TILECOMP_DT), cblks=np.frombuffer(cbs.tobytes(), jobs.CBLK_DT),
                blob=np.concatenate([blob, np.zeros(8, np.uint8)]), coef_bits=im.coef_bits, codestream=bytes(data))


def canon(job):
    """order-independent view of a job's blocks: geometry + coding parameters + the block's bytes"""
    out = {}
    cb, blob = job["cblks"], job["blob"]
    for b in cb:
        key = (int(b["tilecomp"]), int(b["level"]), int(b["band"]), int(b["y0"]), int(b["x0"]))
        nb = int(b["num_bps"])
        data = bytes(blob[int(b["data_off"]):int(b["data_off"]) + int(b["data_len"])]) if nb and b["data_len"] else b""
        coded = nb > 0 and len(data) > 0 and int(b["num_passes"]) > 0
        out[key] = (int(b["w"]), int(b["h"]), nb if coded else 0, int(b["num_passes"]) if coded else 0,
                    float(b["step"]), int(b["len_cleanup"]) if coded and b["num_passes"] > 1 else 0,     # read for multi-pass HT blocks only
                    data if coded else b"")
    assert len(out) == len(cb)
    return out


def check_against_python_tier2(j2k, data, reduce=0):
    p = j2k.Parsed(data, reduce)
    mine = job_from_parsed(p, data)
    ref = jobs.build_iso_job_from_codestream(data, reduce)
    for k in ("width", "height", "ncomp", "prec", "sgnd", "mct", "reversible", "nlevels", "ht", "coef_bits"):
        assert mine[k] == ref[k], k
    assert np.array_equal(mine["tilecomps"], ref["tilecomps"])
    a, b = canon(mine), canon(ref)
    assert a.keys() == b.keys()
    for k in a:
        assert a[k] == b[k], (k, a[k][:6], b[k][:6])
    return p, mine


OPJ_CASES = [
    dict(w=96, h=64, nc=1, kw=dict(num_resolutions=3)),
    dict(w=200, h=150, nc=3, kw=dict(num_resolutions=4, mct=1)),
    dict(w=333, h=211, nc=3, kw=dict(num_resolutions=5, mct=1, tile_size=(128, 128))),
    dict(w=256, h=256, nc=3, kw=dict(num_resolutions=6, mct=1, quality_mode="rates", quality_layers=[40, 20, 10, 5, 1])),
    dict(w=300, h=200, nc=3, kw=dict(num_resolutions=4, mct=1, irreversible=True, quality_mode="rates", quality_layers=[30, 10])),
    dict(w=256, h=192, nc=3, kw=dict(num_resolutions=4, mct=1, progression="RLCP", quality_layers=[20, 5, 1])),
    dict(w=256, h=192, nc=3, kw=dict(num_resolutions=4, mct=1, progression="RPCL", quality_layers=[20, 5, 1])),
    dict(w=256, h=192, nc=3, kw=dict(num_resolutions=4, mct=1, progression="PCRL", quality_layers=[20, 1])),
    dict(w=256, h=192, nc=3, kw=dict(num_resolutions=4, mct=1, progression="CPRL", quality_layers=[20, 1], tile_size=(128, 64))),
    dict(w=200, h=150, nc=3, kw=dict(num_resolutions=4, mct=1, plt=True, tile_size=(64, 64))),
    dict(w=200, h=150, nc=1, kw=dict(num_resolutions=3, codeblock_size=(32, 32))),
    dict(w=130, h=70, nc=3, kw=dict(num_resolutions=3, mct=1, codeblock_size=(16, 64), quality_layers=[10, 1])),
]


@pytest.mark.parametrize("case", OPJ_CASES, ids=lambda c: "%dx%dx%d-%s" % (c["w"], c["h"], c["nc"], "-".join(
    "%s" % (v if not isinstance(v, (list, tuple)) else len(v)) for v in c["kw"].values())))
def test_openjpeg_codestreams(j2k, case):
    s = jobs.synth_image(case["w"], case["h"], case["nc"], 8, seed=case["w"])
    data = opj_encode(s, **case["kw"])
    prog = {"LRCP": 0, "RLCP": 1, "RPCL": 2, "PCRL": 3, "CPRL": 4}[case["kw"].get("progression", "LRCP")]
    if prog <= 1:                                                  # the Python harness reads LRCP / RLCP only
        p, mine = check_against_python_tier2(j2k, data)
    else:
        p = j2k.Parsed(data)
        mine = job_from_parsed(p, data)
    assert p.info["progression"] == prog
    assert p.info["layers"] == len(case["kw"].get("quality_layers", [0]))
    if case["kw"].get("plt"):
        assert p.info["plt_packets"] == p.info["packets"] > 0      # every packet length announced and confirmed
    got = O.iso_decode_job(mine).reshape(case["h"], case["w"], -1)[:, :, :case["nc"]]
    ref = opj_decode(data).reshape(case["h"], case["w"], case["nc"])
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("w,h,nc,tw,nl,passes,P", [
    (200, 150, 1, None, 3, 1, 0), (256, 256, 3, None, 5, 3, 2), (333, 211, 3, 128, 4, 2, 3), (333, 211, 3, 128, 4, 3, 1),
    (70, 5, 1, None, 2, 3, 1), (5, 90, 3, None, 3, 1, 0), (64, 64, 1, None, 0, 1, 0),
])
def test_our_htj2k_codestreams(j2k, w, h, nc, tw, nl, passes, P):
    s = jobs.synth_image(w, h, nc, 8, seed=w + h)
    job = jobs.build_iso_job(s, 8, tw, tw, nl, ht_passes=passes, ht_plane=P)
    data = job["codestream"]
    p, mine = check_against_python_tier2(j2k, data)
    assert p.image.ht == 1 and p.info["zero_copy"] == 1            # one layer: every block's bytes are contiguous in the codestream
    got = O.iso_decode_job(mine).reshape(h, w, -1)[:, :, :nc]
    assert np.array_equal(got, opj_decode(data).reshape(h, w, nc))
    if passes == 1 and P == 0:
        assert np.array_equal(got, np.moveaxis(s, 0, 2))           # lossless


@pytest.mark.parametrize("w,h,reduce", [(333, 211, 1), (330, 210, 2), (320, 200, 3), (384, 256, 2)])   # (sizes Pillow's reduce accepts)
def test_reduce_resolution(j2k, w, h, reduce):
    """Config.ReduceResolution (jpeg2000.go:205-207): the tables of the reduced decode = OpenJPEG's reduced decode"""
    s = jobs.synth_image(w, h, 3, 8, seed=9)
    data = opj_encode(s, num_resolutions=5, mct=1, tile_size=(128, 128), quality_layers=[10, 1])
    p, mine = check_against_python_tier2(j2k, data, reduce)
    ref = opj_decode(data, reduce)
    assert (p.image.height, p.image.width) == ref.shape[:2]
    assert np.array_equal(O.iso_decode_job(mine).reshape(ref.shape[0], ref.shape[1], -1)[:, :, :3], ref)
    with pytest.raises(j2k.J2KError):
        j2k.Parsed(data, 6)


def rewrite(data, sop=False, eph=False, tlm=False, split=False, plt=False):
    """re-emit a codestream of our writer / OpenJPEG with SOP / EPH marker segments in every packet, a TLM marker in the main
    header, PLT markers, or every tile split into two tile-parts at a packet boundary (test-side transformation; uses the
    harness parser's packet map)"""
    h = cs.parse_codestream(data, keep_packets=True)
    return cs.reassemble(h, sop=sop, eph=eph, tlm=tlm, split=split, plt=plt)


@pytest.mark.parametrize("opts", [dict(sop=True), dict(eph=True), dict(sop=True, eph=True), dict(tlm=True), dict(plt=True),
                                  dict(split=True), dict(split=True, tlm=True, plt=True, sop=True, eph=True)],
                         ids=lambda o: "+".join(sorted(o)))
def test_marker_variants(j2k, opts):
    s = jobs.synth_image(200, 150, 3, 8, seed=21)
    base = opj_encode(s, num_resolutions=4, mct=1, tile_size=(128, 128), quality_layers=[20, 5, 1])
    data = rewrite(base, **opts)
    assert data != base
    ref = opj_decode(data).reshape(150, 200, 3)
    assert np.array_equal(ref, opj_decode(base).reshape(150, 200, 3))   # OpenJPEG accepts the rewritten stream: it is well-formed
    p = j2k.Parsed(data)
    mine = job_from_parsed(p, data)
    assert canon(mine) == canon(job_from_parsed(j2k.Parsed(base), base))
    if opts.get("plt"):
        assert p.info["plt_packets"] == p.info["packets"]
    if opts.get("tlm"):
        assert p.info["tlm_tile_parts"] == p.info["tile_parts"]
    if opts.get("split"):
        assert p.info["tile_parts"] == 2 * p.info["tiles"] and p.info["zero_copy"] == 0
    assert np.array_equal(O.iso_decode_job(mine).reshape(150, 200, -1)[:, :, :3], ref)


def test_plt_and_tlm_disagreements_are_errors(j2k):
    s = jobs.synth_image(128, 128, 1, 8, seed=2)
    base = opj_encode(s, num_resolutions=3)
    good = bytearray(rewrite(base, plt=True, tlm=True))
    i = good.index(b"\xff\x58")                                    # PLT: corrupt the first packet length
    bad = bytearray(good)
    bad[i + 5] ^= 0x01
    with pytest.raises(j2k.J2KError) as e:
        j2k.Parsed(bytes(bad))
    assert e.value.code == j2k.E_RANGE and "PLT" in str(e.value)
    i = good.index(b"\xff\x55")                                    # TLM: corrupt the tile-part length
    bad = bytearray(good)
    bad[i + 4 + 2 + 1 + 3] ^= 0x01
    with pytest.raises(j2k.J2KError) as e:
        j2k.Parsed(bytes(bad))
    assert "TLM" in str(e.value)


PRECINCT_CASES = [
    dict(w=256, h=192, nc=3, kw=dict(num_resolutions=4, mct=1, precinct_size=(64, 64))),
    dict(w=333, h=211, nc=3, kw=dict(num_resolutions=5, mct=1, precinct_size=(128, 128), quality_layers=[20, 5, 1])),
    dict(w=333, h=211, nc=3, kw=dict(num_resolutions=4, mct=1, precinct_size=(64, 128), tile_size=(128, 128), quality_layers=[10, 1])),
    dict(w=256, h=256, nc=1, kw=dict(num_resolutions=6, precinct_size=(64, 64), codeblock_size=(32, 32))),
    dict(w=200, h=150, nc=3, kw=dict(num_resolutions=3, mct=1, precinct_size=(32, 32), codeblock_size=(16, 16))),
    dict(w=300, h=200, nc=3, kw=dict(num_resolutions=4, mct=1, precinct_size=(64, 64), irreversible=True, quality_layers=[30, 10])),
    dict(w=256, h=192, nc=3, kw=dict(num_resolutions=4, mct=1, precinct_size=(64, 64), progression="RLCP", quality_layers=[20, 1])),
    dict(w=256, h=192, nc=3, kw=dict(num_resolutions=4, mct=1, precinct_size=(64, 64), progression="RPCL", quality_layers=[20, 1])),
    dict(w=333, h=211, nc=3, kw=dict(num_resolutions=4, mct=1, precinct_size=(64, 32), progression="PCRL", quality_layers=[20, 1])),
    dict(w=333, h=211, nc=3, kw=dict(num_resolutions=5, mct=1, precinct_size=(32, 64), progression="CPRL", tile_size=(128, 128))),
    dict(w=256, h=192, nc=3, kw=dict(num_resolutions=4, mct=1, precinct_size=(128, 128), progression="RPCL", plt=True)),
]


@pytest.mark.parametrize("case", PRECINCT_CASES, ids=lambda c: "%dx%dx%d-%s" % (c["w"], c["h"], c["nc"], "-".join(
    "%s" % (v if not isinstance(v, (list, tuple)) else "x".join(map(str, v))) for v in c["kw"].values())))
def test_user_defined_precincts(j2k, case):
    """COD with precinct sizes (A.6.1, B.6): blocks are cut by the precinct grid, every precinct-band has its own tag trees,
    a layer has one packet per precinct, and the position-driven progressions visit precincts by their reference-grid
    corner -- the tables decode (CPU checker) to OpenJPEG's pixels of the same bytes"""
    s = jobs.synth_image(case["w"], case["h"], case["nc"], 8, seed=case["w"] + 7)
    data = opj_encode(s, **case["kw"])
    p = j2k.Parsed(data)
    assert p.info["packets"] > p.info["layers"] * (case["kw"]["num_resolutions"]) * case["nc"] * p.info["tiles"]   # more than one precinct somewhere
    if case["kw"].get("plt"):
        assert p.info["plt_packets"] == p.info["packets"]
    mine = job_from_parsed(p, data)
    got = O.iso_decode_job(mine).reshape(case["h"], case["w"], -1)[:, :, :case["nc"]]
    assert np.array_equal(got, opj_decode(data).reshape(case["h"], case["w"], case["nc"]))
    if not case["kw"].get("irreversible") and len(case["kw"].get("quality_layers", [1])) == 1:
        assert np.array_equal(got, np.moveaxis(s, 0, 2).reshape(case["h"], case["w"], case["nc"]))


def test_jp2_container_and_rgba(j2k):
    """a JP2 file is accepted as it is (the codestream box is located; the other boxes stay with the Go side), and a fourth
    component becomes the alpha channel as in createImage (decoder.go:489-523)"""
    s = jobs.synth_image(200, 150, 4, 8, seed=5)
    for no_jp2 in (True, False):
        buf = io.BytesIO()
        PIL_Image.fromarray(np.moveaxis(s, 0, 2).astype(np.uint8), "RGBA").save(buf, format="JPEG2000", no_jp2=no_jp2, num_resolutions=4)
        data = buf.getvalue()
        assert (data[4:8] == b"jP  ") == (not no_jp2)
        p = j2k.Parsed(data)
        assert p.image.ncomp == 4 and p.info["zero_copy"] == 1
        got = O.iso_decode_job(job_from_parsed(p, data)).reshape(150, 200, 4)
        assert np.array_equal(got, np.moveaxis(s, 0, 2)) and np.array_equal(got, opj_decode(data))
    with pytest.raises(j2k.J2KError) as e:
        j2k.Parsed(data[:40])                                      # signature + file type boxes only
    assert "codestream box" in str(e.value)


@pytest.mark.parametrize("rate", [0, 150])
def test_16_bit_rgb_written_by_opencv(j2k, rate):
    """three 16-bit components in a JP2 file written by OpenCV's OpenJPEG encoder (lossless, and truncated by a rate target):
    tier-2 + CPU checker give OpenCV's own decode (RGBA64, big-endian channels)"""
    cv2 = pytest.importorskip("cv2")
    s = jobs.synth_image(300, 200, 3, 16, seed=6)
    bgr = np.ascontiguousarray(np.moveaxis(s, 0, 2).astype(np.uint16)[:, :, ::-1])
    ok, enc = cv2.imencode(".jp2", bgr, [cv2.IMWRITE_JPEG2000_COMPRESSION_X1000, rate] if rate else [])
    assert ok
    data = enc.tobytes()
    p = j2k.Parsed(data)
    assert p.image.prec[0] == 16 and p.image.ncomp == 3 and p.image.coef_bits > 14
    got = O.iso_decode_job(job_from_parsed(p, data)).reshape(200, 300, 4, 2)
    val = (got[..., 0].astype(np.uint16) << 8) | got[..., 1]
    ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED)
    assert np.array_equal(val[:, :, :3], ref[:, :, ::-1]) and (val[:, :, 3] == 65535).all()   # (OpenCV returns B G R)


def test_unsupported_features_are_reported(j2k):
    s = jobs.synth_image(128, 128, 3, 8, seed=3)
    good = opj_encode(s, num_resolutions=3, mct=1)
    i = good.index(b"\xff\x5c")                                   # a COC marker segment in front of QCD: component 0 with one level less
    coc = b"\xff\x53" + (9).to_bytes(2, "big") + bytes([0, 0, 1, 4, 4, 0, 1])
    with pytest.raises(j2k.J2KError) as e:
        j2k.Parsed(good[:i] + coc + good[i:])
    assert e.value.code == j2k.E_UNSUPPORTED and "COC" in str(e.value)
    same = b"\xff\x53" + (9).to_bytes(2, "big") + bytes([0, 0, 2, 4, 4, 0, 1])   # what COD says anyway: accepted
    j2k.Parsed(good[:i] + same + good[i:]).close()
    poc = b"\xff\x5f" + (9).to_bytes(2, "big") + bytes([0, 0, 0, 1, 3, 3, 0])       # progression order change
    with pytest.raises(j2k.J2KError) as e:
        j2k.Parsed(good[:i] + poc + good[i:])
    assert e.value.code == j2k.E_UNSUPPORTED and "FF5F" in str(e.value)
    with pytest.raises(j2k.J2KError) as e:
        j2k.Parsed(b"\x00\x01\x02\x03")
    assert e.value.code == j2k.E_ARG
    with pytest.raises(j2k.J2KError):
        j2k.Parsed(b"")


def test_fuzz_contract_never_faults(j2k):
    """fuzz_test.go: malformed input must not panic -- every truncation and 2000 byte-flip mutants either parse or error"""
    s = jobs.synth_image(96, 80, 3, 8, seed=4)
    streams = [opj_encode(s, num_resolutions=3, mct=1, quality_layers=[10, 1], tile_size=(64, 64)),
               jobs.build_iso_job(s, 8, 64, 64, 2, ht_passes=3, ht_plane=1)["codestream"]]
    try:                                                  # a stream with several codeword segments per block (BYPASS | TERMALL)
        from datagen import opj_direct
        streams.append(opj_direct.encode(s, mode=0x05, num_resolutions=3, tile=(64, 64), rates=[10, 1]))
    except OSError:
        pass
    rng = np.random.default_rng(11)
    for data in streams:
        ok = 0
        for n in range(0, len(data), 7):
            try:
                j2k.Parsed(data[:n], threads=1).close()
                ok += 1
            except j2k.J2KError:
                pass
        for _ in range(1000):
            b = bytearray(data)
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            try:
                p = j2k.Parsed(bytes(b), threads=1)
                tcs, cbs, blob = p.tables()
                assert (cbs["data_off"] + cbs["data_len"] <= blob.size).all()   # tables never point outside the blob
                p.close()
            except j2k.J2KError:
                pass
        assert ok > 0                                              # truncated tile data is tolerated (remaining packets absent)


def test_many_tiles_parsed_concurrently_equal_serial(j2k):
    s = jobs.synth_image(512, 384, 3, 8, seed=5)
    data = opj_encode(s, num_resolutions=4, mct=1, tile_size=(64, 64), quality_layers=[8, 1])
    a, b = j2k.Parsed(data, threads=1), j2k.Parsed(data, threads=8)
    ta, tb = a.tables(), b.tables()
    assert a.info["tiles"] == 48 and all(np.array_equal(x, y) for x, y in zip(ta, tb))


# ---- JP2 container: the colour specification box selects the conversion to sRGB (decoder.go:135-178, 350-356) ------------
ENUMCS_TO_CS = {0: 0, 15: 0, 16: 0, 17: 0, 99: 0,              # bi-level, sRGB, grey, unknown: no conversion
                1: 1, 18: 1, 22: 1, 23: 1, 24: 1,              # YCbCr(1), sYCC, YPbPr 1125/60 and 1250/50, e-sYCC: BT.709 matrix
                3: 2, 4: 2, 9: 3, 11: 4, 12: 5, 13: 6, 14: 7, 19: 8, 20: 9, 21: 10}


def test_jp2_colour_specification_box(j2k):
    s = jobs.synth_image(64, 48, 3, 8, seed=4)
    data = opj_encode(s, num_resolutions=3, mct=0)
    for enumcs, want in ENUMCS_TO_CS.items():
        p = j2k.Parsed(cs.wrap_jp2(data, enumcs, 64, 48))
        assert p.image.colorspace == want, enumcs
        assert (p.image.width, p.image.height, p.image.ncomp) == (64, 48, 3)
        p.close()
    assert j2k.Parsed(data).image.colorspace == 0                                          # raw codestream: unspecified
    assert j2k.Parsed(cs.wrap_jp2(data, 18, 64, 48, extra_colr=[16])).image.colorspace == 0  # the last colr box wins (box.go:427-431)
    assert j2k.Parsed(cs.wrap_jp2(data, 16, 64, 48, extra_colr=[14])).image.colorspace == 7
    assert j2k.Parsed(cs.wrap_jp2(data, 18, 64, 48, colr_method=2)).image.colorspace == 0    # ICC profile: EnumCS stays 0
    assert j2k.Parsed(cs.wrap_jp2(data, 18, 64, 48, header_after=True)).image.colorspace == 0  # readJP2 stops at the codestream box
    bad = cs.wrap_jp2(data, 18, 64, 48)
    k = bad.index(b"colr")
    short = bad[:k - 4] + (8 + 5).to_bytes(4, "big") + b"colr" + bytes([1, 0, 0, 0, 0]) + bad[k + 4 + 7:]
    short = short[:short.index(b"jp2h") - 4] + (int.from_bytes(bad[bad.index(b"jp2h") - 4:bad.index(b"jp2h")], "big") - 2).to_bytes(4, "big") + short[short.index(b"jp2h"):]
    with pytest.raises(j2k.J2KError) as e:
        j2k.Parsed(short)
    assert "color specification box too short" in str(e.value)


def test_pillow_jp2_file_parses_like_its_codestream(j2k):
    s = jobs.synth_image(80, 60, 3, 8, seed=6)
    a = np.moveaxis(s, 0, 2).astype(np.uint8)
    buf = io.BytesIO()
    PIL_Image.fromarray(a).save(buf, format="JPEG2000", num_resolutions=3)
    p = j2k.Parsed(buf.getvalue())
    assert p.image.colorspace == 0 and (p.image.width, p.image.height) == (80, 60)
    buf = io.BytesIO()
    PIL_Image.fromarray(a).convert("YCbCr").save(buf, format="JPEG2000", num_resolutions=3)
    data = buf.getvalue()
    if b"colr" in data and data[data.index(b"colr") + 4] == 1 and int.from_bytes(data[data.index(b"colr") + 7:data.index(b"colr") + 11], "big") == 18:
        assert j2k.Parsed(data).image.colorspace == 1                                       # Pillow writes YCbCr images as sYCC


@pytest.mark.parametrize("kw", [dict(num_resolutions=4, mct=1), dict(num_resolutions=4, mct=1, irreversible=True, quality_layers=[20, 5]),
                                dict(num_resolutions=3, mct=1, irreversible=True, tile_size=(64, 64))])
def test_qcc_overrides_qcd(j2k, kw):
    """QCC (A.6.5): datagen.codestream.with_qcc gives every component a QCC that repeats the stream's QCD and rewrites QCD to
    other guard bits / exponents.  OpenJPEG decodes both streams to the same pixels, and the product's tier-2 fills the same
    tables (bit-plane counts, steps, coef_bits) from both; the rewritten QCD alone gives other tables."""
    s = jobs.synth_image(200, 150, 3, 8, seed=3)
    d0 = opj_encode(s, **kw)
    d1 = cs.with_qcc(d0)
    assert np.array_equal(opj_decode(d0), opj_decode(d1))
    p0, p1 = j2k.Parsed(d0), j2k.Parsed(d1)
    (_, c0, _), (_, c1, _) = p0.tables(), p1.tables()
    for f in ("data_len", "num_bps", "num_passes", "step", "band", "level", "x0", "y0", "w", "h", "tilecomp"):
        assert np.array_equal(c0[f], c1[f]), f
    assert p0.image.coef_bits == p1.image.coef_bits
    k = d1.index(b"\xff\x5d")
    no_qcc = d1[:k] + d1[k + 3 * (d1[k + 2] * 256 + d1[k + 3] + 2):]                 # the scrambled QCD without the three QCCs
    p2 = j2k.Parsed(no_qcc)
    c2 = p2.tables()[1]
    assert not (np.array_equal(c0["num_bps"], c2["num_bps"]) and np.array_equal(c0["step"], c2["step"]))
    for p in (p0, p1, p2):
        p.close()


@pytest.mark.parametrize("kw", [dict(num_resolutions=4, mct=1), dict(num_resolutions=4, mct=1, irreversible=True, quality_layers=[20, 5], precinct_size=(64, 64)),
                                dict(num_resolutions=3, mct=1, tile_size=(64, 64))])
def test_coc_overrides_cod(j2k, kw):
    """COC (A.6.2): datagen.codestream.with_coc gives every component a COC that repeats the stream's SPcod while COD announces
    other code-block dimensions.  OpenJPEG decodes both streams to the same pixels and the product's tier-2 fills the same
    tables from both; components that end up with different parameters are refused by name."""
    s = jobs.synth_image(200, 150, 3, 8, seed=3)
    d0 = opj_encode(s, **kw)
    d1 = cs.with_coc(d0)
    assert np.array_equal(opj_decode(d0), opj_decode(d1))
    p0, p1 = j2k.Parsed(d0), j2k.Parsed(d1)
    (_, c0, _), (_, c1, _) = p0.tables(), p1.tables()
    assert len(c0) == len(c1)
    for f in ("data_len", "num_bps", "num_passes", "step", "band", "level", "x0", "y0", "w", "h", "tilecomp"):
        assert np.array_equal(c0[f], c1[f]), f
    k = d1.index(b"\xff\x53")
    for _ in range(2):                                                    # hop to the COC of component 2
        k += 2 + d1[k + 2] * 256 + d1[k + 3]
    assert d1[k:k + 2] == b"\xff\x53" and d1[k + 4] == 2
    mixed = d1[:k] + d1[k + 2 + d1[k + 2] * 256 + d1[k + 3]:]             # component 2 falls back to the (other) COD parameters
    with pytest.raises(j2k.J2KError) as e:
        j2k.Parsed(mixed)
    assert e.value.code == j2k.E_UNSUPPORTED and "component 2" in str(e.value)
    p0.close(); p1.close()
