// rgb_expand.h -- host half of the packed-RGB pixel transfer of host-buffer runs (pure host code, no CUDA).
//
// createImage gives a three-component 8-bit image the Go type image.RGBA (decoder.go:468-487): 4 bytes per pixel whose
// fourth is the constant 255.  The device->host copy of the pixels is what bounds the end-to-end rate of the path, so for
// such images the device writes packed R G B (idwt_wide.cu, RGB24), the copy engine moves 3 bytes per pixel into a
// page-locked staging block of the library, and the worker threads of this pool widen finished rows into the caller's
// RGBA buffer while the next chunk is decoded and copied.  The caller sees exactly the bytes it would have received.
// Whether it pays is a property of the host: the threads must store 4 bytes per pixel faster than the link delivers 3
// (option host_alpha; off on hosts with fewer than 32 hardware threads per GPU, where it measured slower: DESIGN.md 5).
#pragma once
#include <stdint.h>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

struct J2kExpandTask {
    const uint8_t *src; uint8_t *dst;        // first row of the band: packed R G B -> R G B 255
    uint64_t sstride, dstride;
    uint32_t width, rows;
};

// widen `rows` rows of `width` pixels (SSSE3 / AVX2 when the CPU has them)
void j2k_expand_rgb24(const J2kExpandTask &t);

class J2kExpandPool {
public:
    ~J2kExpandPool() { stop(); }
    void start(unsigned threads);            // idempotent
    void submit(const J2kExpandTask &t);     // cut into bands of rows; callable from any thread (also from a CUDA host callback)
    void wait_idle();                        // every submitted band has been written
    void stop();
    unsigned threads() const { return (unsigned)th_.size(); }
private:
    void run();
    std::vector<std::thread> th_;
    std::deque<J2kExpandTask> q_;
    std::mutex mu_;
    std::condition_variable cv_, idle_;
    uint64_t pending_ = 0;
    bool quit_ = false;
};
