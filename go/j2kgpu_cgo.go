//go:build cuda

// Package jpeg2000: cgo binding of libj2kgpu.so (include/j2kgpu.h).
//
// Drop this file next to decoder.go in mrjoshuak/go-jpeg2000 and build with `-tags cuda`; it replaces the
// body of decoder.decodeTiles from tcd.NewTileDecoder onward (decoder.go:311-359) with ONE blocking C call.
// Everything above it (readFormat, parseCodestream, the tier-2 walk that fills the job tables) stays in Go.
// UNCOMPILED in this repository: there is no Go toolchain in the build image (see INTEGRATION.md).
package jpeg2000

/*
#cgo CFLAGS: -I${SRCDIR}/include
#cgo LDFLAGS: -L${SRCDIR}/lib -lj2kgpu
#include <stdlib.h>
#include "j2kgpu.h"
*/
import "C"

import (
	"fmt"
	"image"
	"os"
	"runtime"
	"strconv"
	"unsafe"
)

// gpuCtx is one j2kgpu_ctx (one GPU, one stream, pooled device buffers, pinned staging).  A ctx serialises its calls, so
// a bounded free list keeps one per concurrently decoding goroutine; J2KGPU_DEVICE selects the device.  A ctx that the
// list has no room for is destroyed at once, and a finalizer destroys whatever the collector finds unreferenced: a
// sync.Pool would drop contexts at any GC without ever calling j2kgpu_destroy.
type gpuCtx struct{ h *C.j2kgpu_ctx }

var gpuFree = make(chan *gpuCtx, 16)

func (c *gpuCtx) Close() {
	if c.h != nil {
		C.j2kgpu_destroy(c.h)
		c.h = nil
	}
}

func getGPUCtx() (*gpuCtx, error) {
	select {
	case c := <-gpuFree:
		return c, nil
	default:
	}
	dev, _ := strconv.Atoi(os.Getenv("J2KGPU_DEVICE"))
	var h *C.j2kgpu_ctx
	if rc := C.j2kgpu_create(C.int(dev), &h); rc != 0 {
		return nil, fmt.Errorf("j2kgpu_create: %s", C.GoString(C.j2kgpu_strerror(rc)))
	}
	c := &gpuCtx{h}
	runtime.SetFinalizer(c, (*gpuCtx).Close)
	return c, nil
}

func putGPUCtx(c *gpuCtx) {
	select {
	case gpuFree <- c:
	default:
		c.Close()
	}
}

// gpuJob is what the Go tier-2 walk produces: flat tables (no Go pointers inside) plus one blob.
type gpuJob struct {
	img       C.j2k_image_t
	tilecomps []C.j2k_tilecomp_t
	cblks     []C.j2k_cblk_t
	blob      []byte
}

// gpuColorSpace maps the colour spaces whose conversion to sRGB the library applies in its pixel epilogue
// (decoder.go:350-356, colorspace.go:54-88): all of getColorConversion's cases.
func gpuColorSpace(cs ColorSpace) (C.uint8_t, bool) {
	switch cs {
	case ColorSpaceSYCC, ColorSpaceEYCC, ColorSpaceYPbPr60, ColorSpaceYPbPr50:
		return C.J2KGPU_CS_YCC709, true
	case ColorSpaceYCbCr2, ColorSpaceYCbCr3:
		return C.J2KGPU_CS_YCC601, true
	case ColorSpacePhotoYCC:
		return C.J2KGPU_CS_PHOTOYCC, true
	case ColorSpaceCMY:
		return C.J2KGPU_CS_CMY, true
	case ColorSpaceCMYK:
		return C.J2KGPU_CS_CMYK, true
	case ColorSpaceYCCK:
		return C.J2KGPU_CS_YCCK, true
	case ColorSpaceCIELab: // the four math.Pow conversions: CUDA's pow(), 1 LSB tolerance (include/j2kgpu.h)
		return C.J2KGPU_CS_CIELAB, true
	case ColorSpaceCIEJab:
		return C.J2KGPU_CS_CIEJAB, true
	case ColorSpaceESRGB:
		return C.J2KGPU_CS_ESRGB, true
	case ColorSpaceROMMRGB:
		return C.J2KGPU_CS_ROMM, true
	}
	return C.J2KGPU_CS_NONE, getColorConversion(cs) == nil // false: a conversion this library does not know
}

// decodeTilesGPU is decoder.decodeTiles on the GPU: it returns the same image types createImage would
// (decoder.go:417-588) with Pix filled by the library; errors are wrapped like decoder.go:47.
func (d *decoder) decodeTilesGPU(job *gpuJob) (image.Image, error) {
	ctx, err := getGPUCtx()
	if err != nil {
		return nil, fmt.Errorf("decoding tiles: %w", err)
	}
	defer putGPUCtx(ctx)

	w, h := int(job.img.width), int(job.img.height)
	if w <= 0 || h <= 0 {
		return nil, fmt.Errorf("decoding tiles: empty image") // &pix[0] below needs at least one pixel
	}
	var out image.Image
	var pix []uint8
	var stride int
	prec8 := job.img.prec[0] <= 8
	switch {
	case job.img.ncomp == 1 && prec8:
		im := image.NewGray(image.Rect(0, 0, w, h))
		out, pix, stride = im, im.Pix, im.Stride
	case job.img.ncomp == 1:
		im := image.NewGray16(image.Rect(0, 0, w, h))
		out, pix, stride = im, im.Pix, im.Stride
	case prec8:
		im := image.NewRGBA(image.Rect(0, 0, w, h))
		out, pix, stride = im, im.Pix, im.Stride
	default:
		im := image.NewRGBA64(image.Rect(0, 0, w, h))
		out, pix, stride = im, im.Pix, im.Stride
	}
	var tc *C.j2k_tilecomp_t
	var cb *C.j2k_cblk_t
	var blob *C.uint8_t
	if len(job.tilecomps) > 0 {
		tc = &job.tilecomps[0]
	}
	if len(job.cblks) > 0 {
		cb = &job.cblks[0]
	}
	if len(job.blob) > 0 {
		blob = (*C.uint8_t)(unsafe.Pointer(&job.blob[0]))
	}
	// Page-lock the pixel buffer for the call (ABI v3): the device->host copy of the pixels is the slowest step of the
	// path and runs at link speed, overlapped with the kernels, only into page-locked memory.  The driver keeps the
	// address while the buffer is registered, so the buffer is pinned in the Go sense too (runtime.Pinner, Go 1.21) until
	// it is unregistered; a decoder that recycles its images would take them from j2kgpu_host_alloc instead.  Failure
	// to register is not an error: the copy is then staged by the driver.
	var pin runtime.Pinner
	pin.Pin(&pix[0])
	defer pin.Unpin()
	if C.j2kgpu_host_register(ctx.h, unsafe.Pointer(&pix[0]), C.uint64_t(len(pix))) == 0 {
		defer C.j2kgpu_host_unregister(ctx.h, unsafe.Pointer(&pix[0]))
	}
	rc := C.j2kgpu_decode(ctx.h, &job.img, tc, C.uint32_t(len(job.tilecomps)), cb, C.uint32_t(len(job.cblks)),
		blob, C.uint64_t(len(job.blob)), (*C.uint8_t)(unsafe.Pointer(&pix[0])), C.uint64_t(stride))
	if rc != 0 {
		return nil, fmt.Errorf("decoding tiles: %s: %s", C.GoString(C.j2kgpu_strerror(rc)),
			C.GoString(C.j2kgpu_last_error(ctx.h)))
	}
	return out, nil
}

// decodeCodestreamGPU hands the raw codestream (what readJP2 or the SOC check found) to the library's own front door:
// main header, tile-part index and tier-2 run on host threads inside libj2kgpu.so (host/tier2.cpp), so no Go tier-2 is
// needed.  pix / stride belong to the image.* the caller allocated as createImage would; reduce = Config.ReduceResolution.
func decodeCodestreamGPU(cs []byte, reduce int, pix []uint8, stride int) error {
	if len(cs) == 0 || len(pix) == 0 {
		return fmt.Errorf("decoding tiles: empty input")
	}
	ctx, err := getGPUCtx()
	if err != nil {
		return fmt.Errorf("decoding tiles: %w", err)
	}
	defer putGPUCtx(ctx)
	rc := C.j2kgpu_decode_codestream(ctx.h, (*C.uint8_t)(unsafe.Pointer(&cs[0])), C.uint64_t(len(cs)), C.uint32_t(reduce),
		(*C.uint8_t)(unsafe.Pointer(&pix[0])), C.uint64_t(stride))
	if rc != 0 {
		return fmt.Errorf("decoding tiles: %s: %s", C.GoString(C.j2kgpu_strerror(rc)), C.GoString(C.j2kgpu_last_error(ctx.h)))
	}
	return nil
}

// encodeTileGPU replaces encoder.preprocess (encoder.go:216-281) and the body of encoder.encodeTile up to
// createTileHeader (encoder.go:597-743): it takes the Pix bytes of the image.Gray / Gray16 / RGBA / RGBA64 / NRGBA /
// NRGBA64 the encoder was given (extractImageData, encoder.go:79-214, happens on the device as well) and returns
// tileData, byte for byte what the goroutine pool produces.  The caller keeps generateSIZ / COD / QCD, createTileHeader
// and writeJP2.  ncomp / pixBits follow the type switch of extractImageData (RGBA: 3 components, alpha ignored).
func encodeTileGPU(pix []uint8, stride, width, height, ncomp, pixBits int, o *Options) ([]byte, error) {
	if len(pix) == 0 || width <= 0 || height <= 0 {
		return nil, fmt.Errorf("preprocessing: empty image")
	}
	ctx, err := getGPUCtx()
	if err != nil {
		return nil, fmt.Errorf("preprocessing: %w", err)
	}
	defer putGPUCtx(ctx)
	var p C.j2k_encode_t
	p.width, p.height = C.uint32_t(width), C.uint32_t(height)
	p.ncomp, p.pix_bits = C.uint16_t(ncomp), C.uint8_t(pixBits)
	if o.Precision > 0 && o.Precision <= 16 {
		p.precision = C.uint8_t(o.Precision)
	}
	if o.Lossless {
		p.lossless = 1
	}
	if o.NumResolutions > 0 && o.NumResolutions < 256 {
		p.num_resolutions = C.uint8_t(o.NumResolutions)
	}
	p.cb_x, p.cb_y = C.uint8_t(o.CodeBlockSize.X), C.uint8_t(o.CodeBlockSize.Y)
	p.quality = C.int32_t(o.Quality)
	out := make([]byte, width*height*ncomp*2+65536)
	var n C.uint64_t
	for {
		rc := C.j2kgpu_encode_tile(ctx.h, &p, (*C.uint8_t)(unsafe.Pointer(&pix[0])), C.uint64_t(stride),
			(*C.uint8_t)(unsafe.Pointer(&out[0])), C.uint64_t(len(out)), &n, nil, nil, 0)
		if rc == C.J2KGPU_E_ARG && uint64(n) > uint64(len(out)) { // the size needed was reported: once more
			out = make([]byte, n)
			continue
		}
		if rc != 0 {
			return nil, fmt.Errorf("encoding tile: %s: %s", C.GoString(C.j2kgpu_strerror(rc)), C.GoString(C.j2kgpu_last_error(ctx.h)))
		}
		return out[:n], nil
	}
}
