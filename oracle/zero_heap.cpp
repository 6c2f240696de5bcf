// ORACLE / TEST INFRASTRUCTURE ONLY.
//
// The reference reads one element past the end of two std::vector<float> in
// PiecewiseLinearSpectrum::operator() (spectrum.cpp:101-109, when lambda lies beyond the
// second-to-last knot -- e.g. the Canon sensor curves for 720 < lambda <= 830 nm, which feed
// every PixelSensor and its imaging ratio).  That is undefined behaviour: the value read is
// whatever the heap holds after the vector's last element, so the reference's sensor output
// depends on allocation history.  To give the oracle ONE answer without touching the
// reference's sources, this file replaces the global allocation functions of the oracle
// library with versions that zero-fill the whole usable block, which pins those reads to 0 --
// the value the reference sees on a fresh heap and the value the product defines
// (quetzalcoatlus_b200/host/color/spectrum.hpp).  Bound into the library with
// -Wl,-Bsymbolic-functions so that only the oracle's own code uses it.
#include <malloc.h>

#include <cstdlib>
#include <cstring>
#include <new>

// calloc zeroes the requested bytes; the slack up to the usable size (where the stray read
// lands) is cleared byte by byte (a plain memset past n trips _FORTIFY_SOURCE)
static void* zeroed_nothrow(std::size_t n) noexcept {
    if (!n) n = 1;
    void* p = std::calloc(1, n);
    if (!p) return nullptr;
    volatile char* c = static_cast<volatile char*>(p);
    for (std::size_t i = n, e = malloc_usable_size(p); i < e; i++) c[i] = 0;
    return p;
}

static void* zeroed(std::size_t n) {
    void* p = zeroed_nothrow(n);
    if (!p) throw std::bad_alloc();
    return p;
}

void* operator new(std::size_t n) { return zeroed(n); }
void* operator new[](std::size_t n) { return zeroed(n); }
void* operator new(std::size_t n, const std::nothrow_t&) noexcept { return zeroed_nothrow(n); }
void* operator new[](std::size_t n, const std::nothrow_t& t) noexcept { return operator new(n, t); }
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, std::size_t) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }
