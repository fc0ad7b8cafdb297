"""GPU: the forward path (j2kgpu_encode_preprocess / j2kgpu_encode_tile; encoder.go:79-281, 597-743) against the CPU checker
(oracle/orc_enc.c) -- planes, tile bytes, bytes per block and bit planes per block, all bit-exact -- and back through the
GPU's own block decoder."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from enc_cases import CASES, go_image

pytestmark = pytest.mark.gpu


def params(j2k, case, flags=0):
    w, h, nc, bits, ll, nr, cbx, cby, q, prec = case
    return j2k.EncodeParams(width=w, height=h, ncomp=nc, pix_bits=bits, precision=prec, lossless=ll, num_resolutions=nr,
                            cb_x=cbx, cb_y=cby, quality=q, flags=flags)


@pytest.mark.parametrize("case", CASES)
def test_forward_path_equals_the_checker(j2k, gpu_ctx, case):
    p = params(j2k, case)
    pix = go_image(p.width, p.height, p.ncomp, p.pix_bits, seed=p.width * 3 + p.height)
    assert np.array_equal(gpu_ctx.encode_preprocess(p, pix), O.encode_preprocess(p, pix))
    want, wlens, wbps = O.encode_tile(p, pix)
    for enc_bytes in (0, 1):                                       # row-mask coder (blocks <= 64 wide) / flag-byte coder
        with gpu_ctx.options(enc_bytes=enc_bytes):
            data, lens, bps = gpu_ctx.encode_tile(p, pix)
        assert np.array_equal(lens, wlens) and np.array_equal(bps, wbps)
        assert np.array_equal(data, want)


@pytest.mark.parametrize("seed", range(24))
def test_random_options_and_contents(j2k, gpu_ctx, seed):
    """random sizes, component counts, block shapes, qualities; noise / sparse / ramps / near-constant contents"""
    rng = np.random.default_rng(1000 + seed)
    w, h = int(rng.integers(1, 200)), int(rng.integers(1, 200))
    nc, bits = int(rng.choice([1, 3, 4])), int(rng.choice([8, 16]))
    case = (w, h, nc, bits, int(rng.integers(0, 2)), int(rng.integers(0, 7)), int(rng.integers(0, 7)), int(rng.integers(0, 7)),
            int(rng.choice([0, 1, 30, 75, 100])), int(rng.choice([0, 0, 5, 12])))
    ch, m = (1 if nc == 1 else 4), (1 << bits) - 1
    kind = seed % 4
    if kind == 0:
        v = rng.integers(0, m + 1, (h, w, ch))
    elif kind == 1:
        v = (rng.random((h, w, ch)) < 0.02) * rng.integers(0, m + 1, (h, w, ch))
    elif kind == 2:
        yy, xx = np.mgrid[0:h, 0:w]
        v = np.stack([(xx * 3 + yy * 5 + c * 17) % (m + 1) for c in range(ch)], axis=2)
    else:
        v = np.full((h, w, ch), m // 2 + 1) + rng.integers(-2, 3, (h, w, ch))
    v = np.clip(v, 0, m).astype(np.uint32)
    if bits == 8:
        pix = v.astype(np.uint8).reshape(-1)
    else:
        o = np.zeros((h, w, ch, 2), np.uint8)
        o[..., 0], o[..., 1] = v >> 8, v & 255
        pix = o.reshape(-1)
    p = params(j2k, case)
    want = O.encode_tile(p, pix)
    assert all(np.array_equal(x, y) for x, y in zip(gpu_ctx.encode_tile(p, pix), want)), case
    with gpu_ctx.options(enc_bytes=1):
        assert all(np.array_equal(x, y) for x, y in zip(gpu_ctx.encode_tile(p, pix), want)), case


@pytest.mark.parametrize("seed", range(12))
def test_blocks_wider_than_64(j2k, gpu_ctx, seed):
    """128- and 256-wide blocks: the row-mask coder with several words per row (neighbourhoods and run-length columns
    across word boundaries), noise / sparse / word-aligned steps / near-constant contents"""
    rng = np.random.default_rng(2000 + seed)
    w, h = int(rng.integers(60, 400)), int(rng.integers(1, 90))
    nc, bits = int(rng.choice([1, 3])), int(rng.choice([8, 16]))
    case = (w, h, nc, bits, int(rng.integers(0, 2)), int(rng.integers(1, 4)), int(rng.choice([5, 6])), int(rng.integers(0, 5)),
            int(rng.choice([0, 3, 75])), 0)
    ch, m = (1 if nc == 1 else 4), (1 << bits) - 1
    kind = seed % 4
    if kind == 0:
        v = rng.integers(0, m + 1, (h, w, ch))
    elif kind == 1:
        v = (rng.random((h, w, ch)) < 0.05) * rng.integers(0, m + 1, (h, w, ch))
    elif kind == 2:
        yy, xx = np.mgrid[0:h, 0:w]
        v = np.stack([((xx // 63) * 977 + yy * 5 + c * 17) % (m + 1) for c in range(ch)], axis=2)
    else:
        v = np.full((h, w, ch), m // 2 + 1) + rng.integers(-3, 4, (h, w, ch))
    v = np.clip(v, 0, m).astype(np.uint32)
    if bits == 8:
        pix = v.astype(np.uint8).reshape(-1)
    else:
        o = np.zeros((h, w, ch, 2), np.uint8)
        o[..., 0], o[..., 1] = v >> 8, v & 255
        pix = o.reshape(-1)
    p = params(j2k, case)
    want = O.encode_tile(p, pix)
    assert all(np.array_equal(x, y) for x, y in zip(gpu_ctx.encode_tile(p, pix), want)), case


def test_default_options_4k(j2k, gpu_ctx):
    """DefaultOptions() (jpeg2000.go:305-320) on a 4K RGB frame: lossy, Quality 75, 6 resolutions, 256 x 256 blocks"""
    case = (3840, 2160, 3, 8, 0, 6, 6, 6, 75, 0)
    p = params(j2k, case)
    pix = go_image(3840, 2160, 3, 8, seed=10)
    a = gpu_ctx.encode_tile(p, pix)
    b = O.encode_tile(p, pix, threads=16)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_padded_rows_and_constant_images(j2k, gpu_ctx):
    case = (90, 50, 3, 8, 1, 5, 4, 4, 0, 0)
    p = params(j2k, case)
    pix = go_image(90, 50, 3, 8, seed=9).reshape(50, 360)
    padded = np.full((50, 512), 0xA5, np.uint8)
    padded[:, :360] = pix
    a = gpu_ctx.encode_tile(p, padded, stride=512)
    b = O.encode_tile(p, pix)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # mid-grey: every coefficient zero after the DC shift -> every block nil, an empty tile (t1_fast5.go:20-22)
    data, lens, bps = gpu_ctx.encode_tile(p, np.full(90 * 50 * 4, 128, np.uint8))
    assert len(data) == 0 and not lens.any() and not bps.any()
    # white: only the blocks that reach the plane's top-left corner values are coded
    a = gpu_ctx.encode_tile(p, np.full(90 * 50 * 4, 255, np.uint8))
    b = O.encode_tile(p, np.full(90 * 50 * 4, 255, np.uint8))
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and len(a[0]) > 0


def test_blocks_decode_on_the_gpu(j2k, gpu_ctx):
    """encode on the GPU, decode the blocks with the GPU's reference-mode EBCOT decoder (given the bit-plane counts the
    reference leaves out of its codestream): the planes come back"""
    case = (192, 128, 3, 8, 1, 6, 4, 4, 0, 0)
    p = params(j2k, case)
    pix = go_image(192, 128, 3, 8, seed=4)
    planes = gpu_ctx.encode_preprocess(p, pix)
    data, lens, bps = gpu_ctx.encode_tile(p, pix)
    import test_oracle_encode as T
    blocks = T.block_list(p)
    jobs, off = [], 0
    for (c, sx, sy, w, h, band), n, nb in zip(blocks, lens, bps):
        jobs.append((data[off:off + int(n)].tobytes(), w, h, int(nb), band))
        off += int(n)
    got = gpu_ctx.t1_decode_blocks(jobs)
    for (c, sx, sy, w, h, band), g in zip(blocks, got):
        want = np.zeros((h, w), np.int32)
        src = planes[c, sy:sy + h, sx:sx + w]
        want[:src.shape[0], :src.shape[1]] = src
        assert np.array_equal(np.asarray(g).reshape(h, w), want)


def test_device_pointers(j2k, gpu_ctx):
    import torch
    case = (160, 96, 3, 8, 0, 5, 4, 4, 60, 0)
    p = params(j2k, case, flags=j2k.ENC_DEVICE_PTRS)
    pix = go_image(160, 96, 3, 8, seed=5)
    want, wlens, wbps = O.encode_tile(params(j2k, case), pix)
    d_pix = torch.from_numpy(pix).cuda()
    d_out = torch.zeros(len(want) + 64, dtype=torch.uint8, device="cuda")
    n = j2k.lib().j2kgpu_encode_block_count(C.byref(p))
    lens, bps = np.zeros(n, np.uint32), np.zeros(n, np.uint8)
    got = C.c_uint64(0)
    torch.cuda.synchronize()
    rc = j2k.lib().j2kgpu_encode_tile(gpu_ctx._h, C.byref(p), d_pix.data_ptr(), 160 * 4, d_out.data_ptr(), d_out.numel(), C.byref(got),
                                      lens.ctypes.data_as(C.POINTER(C.c_uint32)), bps.ctypes.data_as(j2k.u8p), n)
    assert rc == 0 and got.value == len(want)
    assert np.array_equal(d_out[: got.value].cpu().numpy(), want)
    assert np.array_equal(lens, wlens) and np.array_equal(bps, wbps)
    d_planes = torch.zeros(3 * 160 * 96, dtype=torch.int32, device="cuda")
    rc = j2k.lib().j2kgpu_encode_preprocess(gpu_ctx._h, C.byref(p), d_pix.data_ptr(), 160 * 4, d_planes.data_ptr())
    assert rc == 0
    assert np.array_equal(d_planes.cpu().numpy().reshape(3, 96, 160), O.encode_preprocess(params(j2k, case), pix))


def test_argument_errors(j2k, gpu_ctx):
    pix = go_image(64, 64, 3, 8, seed=6)
    ok = (64, 64, 3, 8, 1, 6, 4, 4, 0, 0)
    for bad, code in [((0, 64) + ok[2:], j2k.E_ARG), ((64, 64, 2) + ok[3:], j2k.E_ARG), ((64, 64, 3, 12) + ok[4:], j2k.E_ARG),
                      (ok[:6] + (7, 4, 0, 0), j2k.E_UNSUPPORTED), ((20000, 4) + ok[2:], j2k.E_UNSUPPORTED)]:
        with pytest.raises(j2k.J2KError) as e:
            gpu_ctx.encode_tile(params(j2k, bad), pix)
        assert e.value.code == code, bad
    p = params(j2k, ok)
    with pytest.raises(j2k.J2KError):
        gpu_ctx.encode_tile(p, pix, stride=100)                    # stride below the row size
    got = C.c_uint64(0)
    out = np.zeros(16, np.uint8)
    rc = j2k.lib().j2kgpu_encode_tile(gpu_ctx._h, C.byref(p), pix.ctypes.data, 256, out.ctypes.data, 16, C.byref(got), None, None, 0)
    assert rc == j2k.E_ARG and got.value == len(O.encode_tile(p, pix)[0])   # too small: the size needed is reported
    assert j2k.lib().j2kgpu_encode_block_count(C.byref(params(j2k, (64, 64, 2) + ok[3:]))) == 0
    a = gpu_ctx.encode_tile(p, pix)                                 # the context still works
    assert np.array_equal(a[0], O.encode_tile(p, pix)[0])


def test_full_size_4k_rgb_lossless(j2k, gpu_ctx):
    """BASELINE configs[1]'s image through the forward path: 3840 x 2160 RGB 8-bit, lossless, 6 resolutions, 64 x 64 blocks"""
    case = (3840, 2160, 3, 8, 1, 6, 4, 4, 0, 0)
    p = params(j2k, case)
    pix = go_image(3840, 2160, 3, 8, seed=7)
    data, lens, bps = gpu_ctx.encode_tile(p, pix)
    want, wlens, wbps = O.encode_tile(p, pix, threads=16)
    assert np.array_equal(lens, wlens) and np.array_equal(bps, wbps) and np.array_equal(data, want)
    assert np.array_equal(gpu_ctx.encode_preprocess(p, pix), O.encode_preprocess(p, pix))


def test_full_size_1080p_lossy(j2k, gpu_ctx):
    """BASELINE configs[4]'s frame shape, 9-7 + ICT + Quality 75 in float64"""
    case = (1920, 1080, 3, 8, 0, 6, 4, 4, 75, 0)
    p = params(j2k, case)
    pix = go_image(1920, 1080, 3, 8, seed=8)
    a = gpu_ctx.encode_tile(p, pix)
    b = O.encode_tile(p, pix, threads=16)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_full_size_8192_grey16(j2k, gpu_ctx):
    """BASELINE configs[3]'s image through the forward path: 8192 x 8192 Gray16, lossless (two waves of code blocks)"""
    w = h = 8192
    rng = np.random.default_rng(84)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    v = ((np.sin(xx / 37.0) + np.cos(yy / 53.0) + 2.0) * 12000.0 + rng.integers(0, 600, (h, w))).astype(np.uint32)
    pix = np.zeros((h, w, 2), np.uint8)
    pix[..., 0], pix[..., 1] = v >> 8, v & 255
    p = params(j2k, (w, h, 1, 16, 1, 6, 4, 4, 0, 0))
    a = gpu_ctx.encode_tile(p, pix)
    b = O.encode_tile(p, pix, threads=16)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("w,bits,q,cb", [(64, 16, 30000, 4), (64, 16, 2000000, 4), (40, 8, 2000000000, 3), (200, 16, 40000, 6)])
def test_out_of_range_quality(j2k, gpu_ctx, w, bits, q, cb):
    """Options.Quality is not validated by the reference: huge values drive the quantised coefficients to 31 bit planes and
    past int32 (Go's int32(float64) gives 0x80000000 there); noise input, the densest blocks the coder can meet"""
    rng = np.random.default_rng(q % 1000 + w)
    m = (1 << bits) - 1
    v = rng.integers(0, m + 1, (w, w, 1)).astype(np.uint32)
    if bits == 8:
        pix = v.astype(np.uint8).reshape(-1)
    else:
        o = np.zeros((w, w, 1, 2), np.uint8)
        o[..., 0], o[..., 1] = v >> 8, v & 255
        pix = o.reshape(-1)
    p = params(j2k, (w, w, 1, bits, 0, 1, cb, cb, q, 0))
    a, b = gpu_ctx.encode_tile(p, pix), O.encode_tile(p, pix)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert int(a[2].max()) == 31
