// clover_b200/containers.hpp - drop-in host containers for the Clover hot path, backed by libclover_b200.so.
//
// Same class names, method names and argument meaning as the reference's header-only containers
// (include/CloverVector32.h, CloverVector4.h, CloverVector8.h, CloverMatrix32.h, CloverMatrix4.h, CloverMatrix8.h),
// so user code and the reference's own harness templates (test/validate/03_matrix.cpp:248-276,
// test/performance/01_measure.h:699-720) compile against them unchanged for the methods on the hot path:
//
//     quantize / restore / dot            (vectors)         + the `_scalar` / `_parallel` spellings
//     quantize / mvm                      (matrices)          (all three forward to the same GPU kernel)
//     getData / getScales / get / getBits / getBytes / size / size_pad / getRows / getCols / setRandomKeys
//
// Where the data lives: every container owns ONE device allocation in the reference's exact byte layout
// ([values | scales], include/CloverVector4.h:68-103) plus a host mirror of the same bytes. getData()/getScales()
// return HOST pointers like the reference does; the mirror is synchronised lazily (device -> host before a host
// read, host -> device before the next kernel). Once a WRITABLE pointer has been handed out, the caller may keep it
// (the reference's idiom): the host image is then re-uploaded before every kernel that reads the container and
// refreshed after every kernel that writes it, until commit() declares the pointer dead - always correct, one PCIe
// copy per call only for containers whose raw pointers escaped.
//
// Errors keep the reference's behaviour: a message on std::cout and exit(1) (include/CloverMatrix4.h:779-782).
// Stochastic rounding is a run-time switch: containers start WITHOUT a key (= the reference built with
// CLOVER_STOCHASTIC_ROUNDING_DISABLED); setRandomKeys()/seed() enables the reference's XORShift128+ stream.
//
// The borrowing view constructor CloverVector4/8(n, values, scales) (include/CloverVector4.h:114-119) works on the caller's
// host buffers: they are re-read before every kernel that reads the view and written back after every kernel that writes
// it (one PCIe copy per call - keep hot operands in owning containers).
// Not provided (outside the hot path): 16-bit containers, fp32 BLAS on CloverVector32/CloverMatrix32.
#ifndef CLOVER_B200_CONTAINERS_HPP
#define CLOVER_B200_CONTAINERS_HPP

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>
#include <vector>
#if defined(__AVX2__)
#include <immintrin.h>      // the reference's __m256i key arguments (include/CloverRandom.h:90-94) are accepted as well
#endif

#include "../clover_b200.h"

#define CLOVER_VECTOR_BLOCK 64
#define CLOVER_VECTOR_SIZE_PAD (CLOVER_VECTOR_BLOCK * 2)

namespace clover_b200_detail {

inline void check(int status, const char *what) {
    if (status != CLOVER_OK) {
        std::cout << what << " failed: " << clover_last_error() << ". Exiting ..." << std::endl;
        exit(1);
    }
}

// One device buffer + host mirror with lazy synchronisation. A VIEW (the reference's borrowing constructor,
// include/CloverVector4.h:114-119) has its bytes in two host regions owned by the caller (values, scales): those
// regions are authoritative whenever the host side is, so they are re-read before every kernel that reads the
// container (the caller may have written through its own pointers) and written back right after every kernel that
// writes it (the caller reads its own memory without asking) - a compatibility path that pays a PCIe copy per call.
class Mirror {
    void *dev_ = nullptr;
    std::vector<unsigned char> host_;
    mutable bool host_fresh_ = true, dev_fresh_ = true;
    unsigned char *ext_a_ = nullptr, *ext_b_ = nullptr;     // view: the caller's values / scales
    size_t ext_a_bytes_ = 0;
    // An owning container whose WRITABLE host pointer was handed out (getData() / getScales()): the reference's idiom is
    // to fetch that pointer once and keep writing / reading through it, so from then on the host image is treated like a
    // view's memory - re-uploaded before every kernel that reads the container, refreshed after every kernel that writes
    // it. Correct but one PCIe copy per call; commit() ends that mode once the caller is done with the raw pointer.
    bool escaped_ = false;
    void pull_view() const {
        std::memcpy(const_cast<unsigned char *>(host_.data()), ext_a_, ext_a_bytes_);
        std::memcpy(const_cast<unsigned char *>(host_.data()) + ext_a_bytes_, ext_b_, host_.size() - ext_a_bytes_);
    }
    void push_view() const {
        std::memcpy(ext_a_, host_.data(), ext_a_bytes_);
        std::memcpy(ext_b_, host_.data() + ext_a_bytes_, host_.size() - ext_a_bytes_);
    }
public:
    explicit Mirror(size_t bytes) : host_(bytes, 0) { check(clover_malloc(&dev_, bytes), "clover_malloc"); dev_fresh_ = false; }
    Mirror(unsigned char *a, size_t a_bytes, unsigned char *b, size_t b_bytes)
        : host_(a_bytes + b_bytes, 0), ext_a_(a), ext_b_(b), ext_a_bytes_(a_bytes) {
        check(clover_malloc(&dev_, host_.size()), "clover_malloc");
        dev_fresh_ = false;
    }
    Mirror(const Mirror &o) : host_(o.host_.size()) {        // deep copy (of a view too: the copy owns its bytes)
        check(clover_malloc(&dev_, host_.size()), "clover_malloc");
        host_.assign(o.host_ro(), o.host_ro() + o.host_.size());
        dev_fresh_ = false;
    }
    Mirror &operator=(const Mirror &) = delete;
    ~Mirror() {
        if (is_view() && !host_fresh_) to_host();
        if (dev_) clover_free(dev_);
    }
    bool is_view() const { return ext_a_ != nullptr; }
    size_t bytes() const { return host_.size(); }
    void to_host() const {
        if (!host_fresh_) {
            check(clover_copy_d2h(const_cast<unsigned char *>(host_.data()), dev_, host_.size(), nullptr), "clover_copy_d2h");
            check(clover_stream_sync(nullptr), "clover_stream_sync");
            host_fresh_ = true;
            if (is_view()) push_view();
        }
    }
    // view / escaped host pointer: write a kernel's result back into the caller-visible memory now
    void flush_view() const { if (is_view() || escaped_) to_host(); }
    void commit() { if (!is_view()) { (void)dev_in(); escaped_ = false; } }
    // host pointer the caller may write through: the device copy becomes stale
    unsigned char *host_rw() { to_host(); dev_fresh_ = false; if (!is_view()) escaped_ = true; return host_.data(); }
    // the same for the container's own code: the pointer does not outlive the call
    unsigned char *host_rw_internal() { to_host(); dev_fresh_ = false; return host_.data(); }
    // the same for the part that starts at `offset` (0 = values, value bytes = scales): a view hands out the caller's own regions
    unsigned char *host_rw_part(size_t offset) {
        if (!is_view()) return host_rw() + offset;
        to_host();
        dev_fresh_ = false;
        return offset < ext_a_bytes_ ? ext_a_ + offset : ext_b_ + (offset - ext_a_bytes_);
    }
    // contiguous read-only image [values | scales]
    const unsigned char *host_ro() const {
        const bool host_was_authoritative = host_fresh_;
        to_host();
        if (is_view() && host_was_authoritative) pull_view();
        return host_.data();
    }
    // device pointer for a kernel that READS the buffer
    const void *dev_in() {
        if (is_view()) {
            if (host_fresh_) {                                 // the caller's memory is authoritative: always re-read it
                pull_view();
                check(clover_copy_h2d(dev_, host_.data(), host_.size(), nullptr), "clover_copy_h2d");
            }
            dev_fresh_ = true;
            return dev_;
        }
        // escaped: the caller may have written through a retained pointer at any time while the host image is current
        if (!dev_fresh_ || (escaped_ && host_fresh_)) { check(clover_copy_h2d(dev_, host_.data(), host_.size(), nullptr), "clover_copy_h2d"); dev_fresh_ = true; }
        return dev_;
    }
    // device pointer for a kernel that OVERWRITES (part of) the buffer
    void *dev_out() { (void)dev_in(); host_fresh_ = false; return dev_; }
};

inline uint64_t pad128(uint64_t n) { return (n % CLOVER_VECTOR_SIZE_PAD) ? n + CLOVER_VECTOR_SIZE_PAD - (n % CLOVER_VECTOR_SIZE_PAD) : n; }

class Keyed {   // include/CloverRandom.h: per-object XORShift128+ state
protected:
    uint64_t key_[8];
    bool has_key_ = false;
    uint64_t *key_ptr() { return has_key_ ? key_ : nullptr; }
public:
    void setRandomKeys(const uint64_t key1[4], const uint64_t key2[4]) {   // include/CloverRandom.h:90-94
        std::memcpy(key_, key1, 32); std::memcpy(key_ + 4, key2, 32); has_key_ = true;
    }
    void seed(uint64_t k1, uint64_t k2) { check(clover_prng_init(k1, k2, key_), "clover_prng_init"); has_key_ = true; }
    void disableStochasticRounding() { has_key_ = false; }
    const uint64_t *getRandomKeys() const { return has_key_ ? key_ : nullptr; }
#if defined(__AVX2__)
    void setRandomKeys(const __m256i &key1, const __m256i &key2) {         // the reference's exact signature
        std::memcpy(key_, &key1, 32); std::memcpy(key_ + 4, &key2, 32); has_key_ = true;
    }
#endif
protected:
    // the object's own key, seeded from the OS on first use like the reference seeds from RDRAND (include/CloverRandom.h:96-114)
    uint64_t *own_key() {
        if (!has_key_) { std::random_device rd; seed(((uint64_t)rd() << 32) | rd() | 1u, ((uint64_t)rd() << 32) | rd() | 1u); }
        return key_;
    }
};

// setRandomFloats / setRandomInteger of CloverVector32 and CloverMatrix32 (include/CloverVector32.h:712-783,
// include/CloverMatrix32.h:252-323) on the device: `Derived` provides device_out() and generator_length().
template <class Derived>
class Generators : public Keyed {
    void fill(bool integer, float lo, float hi, uint64_t *key) {
        Derived &d = static_cast<Derived &>(*this);
        check((integer ? clover_v32_set_random_integers : clover_v32_set_random_floats)(d.device_out(), d.generator_length(), lo, hi, key, nullptr),
              "setRandom");
        d.sync_view();
    }
    void fill_pair(bool integer, float lo, float hi, void *key1, void *key2) {          // two separate 32-byte halves, advanced in place
        uint64_t k[8];
        std::memcpy(k, key1, 32); std::memcpy(k + 4, key2, 32);
        fill(integer, lo, hi, k);
        std::memcpy(key1, k, 32); std::memcpy(key2, k + 4, 32);
    }
public:
    void setRandomFloats(float min_value, float max_value) { fill(false, min_value, max_value, own_key()); }
    void setRandomInteger(float min_value, float max_value) { fill(true, min_value, max_value, own_key()); }
    void setRandomFloats(float min_value, float max_value, uint64_t key1[4], uint64_t key2[4]) { fill_pair(false, min_value, max_value, key1, key2); }
    void setRandomInteger(float min_value, float max_value, uint64_t key1[4], uint64_t key2[4]) { fill_pair(true, min_value, max_value, key1, key2); }
#if defined(__AVX2__)
    void setRandomFloats(float min_value, float max_value, __m256i &key1, __m256i &key2) { fill_pair(false, min_value, max_value, &key1, &key2); }
    void setRandomInteger(float min_value, float max_value, __m256i &key1, __m256i &key2) { fill_pair(true, min_value, max_value, &key1, &key2); }
#endif
};

}  // namespace clover_b200_detail

// ---------------------------------------------------------------------------------------------------------------
class CloverVector32 : public clover_b200_detail::Generators<CloverVector32> {   // include/CloverVector32.h:53-70 - fp32, length padded to x128, pad zeroed
    uint64_t length, length_pad;
    mutable clover_b200_detail::Mirror buf;
public:
    uint64_t generator_length() const { return length; }        // the generators fill `length` elements, the pad stays 0
    explicit CloverVector32(uint64_t s) : length(s), length_pad(clover_b200_detail::pad128(s)), buf(length_pad * sizeof(float)) {}
    uint64_t size() const { return length; }
    uint64_t size_pad() const { return length_pad; }
    uint64_t getBitsLength() const { return 32; }
    uint64_t getBytes() const { return length_pad * sizeof(float); }
    float *getData() const { return reinterpret_cast<float *>(buf.host_rw()); }
    float get(uint64_t i) const { return reinterpret_cast<const float *>(buf.host_ro())[i]; }
    void set(uint64_t i, float v) { getData()[i] = v; }
    const float *device_in() const { return static_cast<const float *>(buf.dev_in()); }
    float *device_out() { return static_cast<float *>(buf.dev_out()); }
    void sync_view() const { buf.flush_view(); }
    void commit() { buf.commit(); }                          // the caller is done with pointers obtained from getData()
};

// ---------------------------------------------------------------------------------------------------------------
template <int BITS>
class CloverQuantizedVector : public clover_b200_detail::Keyed {
protected:
    uint64_t length, length_pad;
    mutable clover_b200_detail::Mirror buf;                 // [values | scales], one allocation (:76-79)
    uint64_t value_bytes() const { return length_pad * BITS / 8; }
    uint64_t scale_count() const { return length_pad / 64; }
    void init_padding() {                                   // pad values 0 (already), pad scales 1 (:86-94)
        float *s = reinterpret_cast<float *>(buf.host_rw_internal() + value_bytes());
        for (uint64_t i = length / 64; i < scale_count(); ++i) s[i] = 1.0f;
    }
public:
    explicit CloverQuantizedVector(uint64_t s)
        : length(s), length_pad(clover_b200_detail::pad128(s)), buf(length_pad * BITS / 8 + (length_pad / 64) * sizeof(float)) { init_padding(); }
    CloverQuantizedVector(const CloverVector32 &other) : CloverQuantizedVector(other.size()) { quantize(other); }
    // borrowing view over the caller's host buffers, not freed here (include/CloverVector4.h:114-119, CloverVector8.h:104-109)
    CloverQuantizedVector(uint64_t s, int8_t *values, float *scales)
        : length(s), length_pad(clover_b200_detail::pad128(s)),
          buf(reinterpret_cast<unsigned char *>(values), length_pad * BITS / 8, reinterpret_cast<unsigned char *>(scales),
              (length_pad / 64) * sizeof(float)) {}
    void sync_view() const { buf.flush_view(); }            // a kernel wrote this container: hand the bytes to a view's owner
    void commit() { buf.commit(); }                         // the caller is done with pointers obtained from getData() / getScales()
    uint64_t size() const { return length; }
    uint64_t size_pad() const { return length_pad; }
    uint64_t getBitsLength() const { return BITS; }
    uint64_t getBytes() const { return value_bytes() + scale_count() * sizeof(float); }
    int8_t *getData() const { return reinterpret_cast<int8_t *>(buf.host_rw_part(0)); }
    float *getScales() const { return reinterpret_cast<float *>(buf.host_rw_part(value_bytes())); }
    const int8_t *device_values() const { return static_cast<const int8_t *>(buf.dev_in()); }
    const float *device_scales() const { return reinterpret_cast<const float *>(static_cast<const char *>(buf.dev_in()) + value_bytes()); }
    const int8_t *host_values() const { return reinterpret_cast<const int8_t *>(buf.host_ro()); }      // read-only host image
    const float *host_scales() const { return reinterpret_cast<const float *>(buf.host_ro() + value_bytes()); }
    int8_t *device_values_out() { return static_cast<int8_t *>(buf.dev_out()); }
    float *device_scales_out() { return reinterpret_cast<float *>(static_cast<char *>(buf.dev_out()) + value_bytes()); }

    void quantize(const CloverVector32 &other) {
        if (other.size_pad() != length_pad) { std::cout << "Vectors do not have the same size. Exiting ..." << std::endl; exit(1); }
        int8_t *v = device_values_out();
        float *s = device_scales_out();
        const int rc = BITS == 4 ? clover_v4_quantize(other.device_in(), length_pad, v, s, key_ptr(), nullptr)
                                 : clover_v8_quantize(other.device_in(), length_pad, v, s, key_ptr(), nullptr);
        clover_b200_detail::check(rc, "quantize");
        sync_view();
    }
    void quantize_scalar(const CloverVector32 &o) { quantize(o); }
    void quantize_parallel(const CloverVector32 &o) { quantize(o); }

    // all values 0, all scales 1 (include/CloverVector4.h:306-318) - the start vector of the IHT / GD loops
    void clear() {
        if (buf.is_view()) {
            std::memset(getData(), 0, value_bytes());
            float *s = getScales();
            for (uint64_t i = 0; i < scale_count(); ++i) s[i] = 1.0f;
            return;
        }
        unsigned char *h = buf.host_rw_internal();
        std::memset(h, 0, value_bytes());
        float *s = reinterpret_cast<float *>(h + value_bytes());
        for (uint64_t i = 0; i < scale_count(); ++i) s[i] = 1.0f;
    }

    void restore(CloverVector32 &other) const {
        if (other.size_pad() != length_pad) { std::cout << "Vectors do not have the same size. Exiting ..." << std::endl; exit(1); }
        const int rc = BITS == 4 ? clover_v4_restore(device_values(), device_scales(), length_pad, other.device_out(), nullptr)
                                 : clover_v8_restore(device_values(), device_scales(), length_pad, other.device_out(), nullptr);
        clover_b200_detail::check(rc, "restore");
        other.sync_view();
    }
    void restore_scalar(CloverVector32 &o) const { restore(o); }

    float dot(const CloverQuantizedVector &other, int mode = CLOVER_DOT_AUTO) const {
        if (other.length_pad != length_pad) { std::cout << "Vectors do not have the same size. Exiting ..." << std::endl; exit(1); }
        static thread_local float *d_result = nullptr;
        if (!d_result) clover_b200_detail::check(clover_malloc(reinterpret_cast<void **>(&d_result), sizeof(float)), "clover_malloc");
        const int rc = BITS == 4 ? clover_v4_dot(device_values(), device_scales(), other.device_values(), other.device_scales(), length_pad, d_result, mode, nullptr)
                                 : clover_v8_dot(device_values(), device_scales(), other.device_values(), other.device_scales(), length_pad, d_result, mode, nullptr);
        clover_b200_detail::check(rc, "dot");
        float h = 0.f;
        clover_b200_detail::check(clover_copy_d2h(&h, d_result, sizeof(float), nullptr), "clover_copy_d2h");
        clover_b200_detail::check(clover_stream_sync(nullptr), "clover_stream_sync");
        return h;
    }
    float dot_scalar(const CloverQuantizedVector &o) const { return dot(o, CLOVER_DOT_EXACT); }
    float dot_parallel(const CloverQuantizedVector &o) const { return dot(o); }

    // this = this + a * other, re-quantized (include/CloverVector4.h:1195-1204, CloverVector8.h:1063-1072)
    void scaleAndAdd(const CloverQuantizedVector &other, float a) { scaleAndAdd(other, a, *this); }
    // result = this + a * other (include/CloverVector4.h:1206-1218, CloverVector8.h:1074-1086)
    void scaleAndAdd(const CloverQuantizedVector &other, float a, CloverQuantizedVector &result) {
        if (other.length_pad != length_pad || result.length_pad != length_pad) { std::cout << "Vectors do not have the same size. Exiting ..." << std::endl; exit(1); }
        const int8_t *u = device_values(), *v = other.device_values();
        const float *su = device_scales(), *sv = other.device_scales();
        int8_t *r = result.device_values_out();
        float *sr = result.device_scales_out();
        const int rc = BITS == 4 ? clover_v4_scale_and_add(u, su, v, sv, a, length_pad, r, sr, key_ptr(), nullptr)
                                 : clover_v8_scale_and_add(u, su, v, sv, a, length_pad, r, sr, key_ptr(), nullptr);
        clover_b200_detail::check(rc, "scaleAndAdd");
        result.sync_view();
    }
    // hard thresholding in place: only the k largest magnitudes survive (include/CloverVector4.h:1913-1973, CloverVector8.h:1680-1740)
    void threshold(uint64_t k, int mode = CLOVER_THRESHOLD_AUTO) {
        int8_t *v = device_values_out();
        const int rc = BITS == 4 ? clover_v4_threshold(v, device_scales(), length, k, mode, nullptr)
                                 : clover_v8_threshold(v, device_scales(), length, k, mode, nullptr);
        clover_b200_detail::check(rc, "threshold");
        sync_view();
    }
    void threshold_parallel(uint64_t k) { threshold(k); }
    void scaleAndAdd_scalar(const CloverQuantizedVector &o, float a) { scaleAndAdd(o, a); }
    void scaleAndAdd_parallel(const CloverQuantizedVector &o, float a) { scaleAndAdd(o, a); }
    void scaleAndAdd_scalar(const CloverQuantizedVector &o, float a, CloverQuantizedVector &r) { scaleAndAdd(o, a, r); }
    void scaleAndAdd_parallel(const CloverQuantizedVector &o, float a, CloverQuantizedVector &r) { scaleAndAdd(o, a, r); }
};

class CloverVector4 : public CloverQuantizedVector<4> {
public:
    using CloverQuantizedVector<4>::CloverQuantizedVector;
    int8_t getBits(uint64_t pos) const {                    // include/CloverVector4.h:154-160
        const int8_t b = reinterpret_cast<const int8_t *>(buf.host_ro())[pos >> 1];
        return (pos & 1) ? (int8_t)((int8_t)(b << 4) >> 4) : (int8_t)(b >> 4);
    }
    float get(uint64_t pos) const {                         // include/CloverVector4.h:179-188
        const float *s = reinterpret_cast<const float *>(buf.host_ro() + value_bytes());
        return (s[pos >> 6] / 7.0f) * (float)getBits(pos);
    }
};

class CloverVector8 : public CloverQuantizedVector<8> {
public:
    using CloverQuantizedVector<8>::CloverQuantizedVector;
    int8_t getBits(uint64_t pos) const { return reinterpret_cast<const int8_t *>(buf.host_ro())[pos]; }
    float get(uint64_t pos) const {                         // include/CloverVector8.h:136-139
        const float *s = reinterpret_cast<const float *>(buf.host_ro() + value_bytes());
        return getBits(pos) * s[pos >> 6] / 127.0f;
    }
};

// ---------------------------------------------------------------------------------------------------------------
class CloverMatrix32 : public clover_b200_detail::Generators<CloverMatrix32> {   // include/CloverMatrix32.h:43-66
    uint64_t rows, cols;
    mutable clover_b200_detail::Mirror buf;
public:
    uint64_t generator_length() const { return rows * cols; }   // size() of the padded matrix (CloverMatrix32.h:254)
    float *device_out() { return static_cast<float *>(buf.dev_out()); }
    void sync_view() const { buf.flush_view(); }
    void commit() { buf.commit(); }
    float get(uint64_t i, uint64_t j) const { return reinterpret_cast<const float *>(buf.host_ro())[i * cols + j]; }
    CloverMatrix32(uint64_t h, uint64_t w) : rows(clover_b200_detail::pad128(h)), cols(clover_b200_detail::pad128(w)), buf(rows * cols * sizeof(float)) {}
    uint64_t getRows() const { return rows; }
    uint64_t getCols() const { return cols; }
    uint64_t size() const { return rows * cols; }
    uint64_t getBytes() const { return rows * cols * sizeof(float); }
    float *getData() const { return reinterpret_cast<float *>(buf.host_rw()); }
    const float *device_in() const { return static_cast<const float *>(buf.dev_in()); }
};

template <int BITS, class QVector>
class CloverQuantizedMatrix : public clover_b200_detail::Keyed {
protected:
    uint64_t rows, cols;
    mutable clover_b200_detail::Mirror buf;                 // [values | scales] (include/CloverMatrix4.h:77-93)
    uint64_t value_bytes() const { return rows * cols * BITS / 8; }
    uint64_t scale_count() const { return (rows >> 6) * (cols >> 6); }
public:
    CloverQuantizedMatrix(uint64_t h, uint64_t w)
        : rows(clover_b200_detail::pad128(h)), cols(clover_b200_detail::pad128(w)),
          buf(rows * cols * BITS / 8 + (rows >> 6) * (cols >> 6) * sizeof(float)) {}
    uint64_t getRows() const { return rows; }
    uint64_t getCols() const { return cols; }
    uint64_t size() const { return rows * cols; }
    uint64_t getBitsLength() const { return BITS; }
    uint64_t getBytes() const { return value_bytes() + scale_count() * sizeof(float); }
    int8_t *getData() const { return reinterpret_cast<int8_t *>(buf.host_rw()); }
    float *getScales() const { return reinterpret_cast<float *>(buf.host_rw() + value_bytes()); }
    const int8_t *device_values() const { return static_cast<const int8_t *>(buf.dev_in()); }
    const float *device_scales() const { return reinterpret_cast<const float *>(static_cast<const char *>(buf.dev_in()) + value_bytes()); }
    const int8_t *host_values() const { return reinterpret_cast<const int8_t *>(buf.host_ro()); }      // read-only host image
    const float *host_scales() const { return reinterpret_cast<const float *>(buf.host_ro() + value_bytes()); }

    void quantize(const CloverMatrix32 &m) {
        if (m.getRows() != rows || m.getCols() != cols) { std::cout << "Matrices do not have the same size. Exiting ..." << std::endl; exit(1); }
        char *d = static_cast<char *>(buf.dev_out());
        const int rc = BITS == 4 ? clover_m4_quantize(m.device_in(), rows, cols, reinterpret_cast<int8_t *>(d), reinterpret_cast<float *>(d + value_bytes()), key_ptr(), nullptr)
                                 : clover_m8_quantize(m.device_in(), rows, cols, reinterpret_cast<int8_t *>(d), reinterpret_cast<float *>(d + value_bytes()), key_ptr(), nullptr);
        clover_b200_detail::check(rc, "quantize");
        buf.flush_view();
    }
    void quantize_scalar(const CloverMatrix32 &m) { quantize(m); }
    void commit() { buf.commit(); }                         // the caller is done with pointers obtained from getData() / getScales()

    // include/CloverMatrix4.h:266-301 (restore_scalar); 8-bit: other(i, j) = get(i, j) (include/CloverMatrix8.h:117-129, :1300)
    void restore(CloverMatrix32 &other) const {
        if (other.getRows() != rows || other.getCols() != cols) { std::cout << "Matrices do not have the same size. Exiting ..." << std::endl; exit(1); }
        const int rc = BITS == 4 ? clover_m4_restore(device_values(), device_scales(), rows, cols, other.device_out(), nullptr)
                                 : clover_m8_restore(device_values(), device_scales(), rows, cols, other.device_out(), nullptr);
        clover_b200_detail::check(rc, "restore");
        other.sync_view();
    }
    void restore_scalar(CloverMatrix32 &other) const { restore(other); }

    // fp32 vectors: include/CloverMatrix4.h:1451-1547, include/CloverMatrix8.h:558-661
    void mvm(const CloverVector32 &productVector, CloverVector32 &resultVector) {
        if (productVector.size() != getCols() || resultVector.size_pad() < getRows()) {
            std::cout << "MVM can not be performed. Exiting ..." << std::endl;
            exit(1);
        }
        const int rc = BITS == 4 ? clover_m4_mvm_f32(device_values(), device_scales(), rows, cols, productVector.device_in(), resultVector.device_out(), nullptr)
                                 : clover_m8_mvm_f32(device_values(), device_scales(), rows, cols, productVector.device_in(), resultVector.device_out(), nullptr);
        clover_b200_detail::check(rc, "mvm");
        resultVector.sync_view();
    }

    // include/CloverMatrix4.h:777-1083 / include/CloverMatrix8.h:1002-1298
    void mvm(const QVector &productVector, QVector &resultVector) {
        if (productVector.size() != getCols() || resultVector.size_pad() != getRows()) {
            std::cout << "MVM can not be performed. Exiting ..." << std::endl;
            exit(1);
        }
        int8_t *yv = resultVector.device_values_out();
        float *ys = resultVector.device_scales_out();
        const int rc = BITS == 4 ? clover_m4_mvm(device_values(), device_scales(), rows, cols, productVector.device_values(), productVector.device_scales(), yv, ys, nullptr, key_ptr(), nullptr)
                                 : clover_m8_mvm(device_values(), device_scales(), rows, cols, productVector.device_values(), productVector.device_scales(), yv, ys, nullptr, key_ptr(), nullptr);
        clover_b200_detail::check(rc, "mvm");
        resultVector.sync_view();
    }
    void mvm_scalar(const QVector &x, QVector &y) { mvm(x, y); }
    void mvm_parallel(const QVector &x, QVector &y) { mvm(x, y); }

    // include/CloverMatrix4.h:1549-1663 / include/CloverMatrix8.h:1359-1385: other(j, i) = this(i, j), scales included
    void transpose(CloverQuantizedMatrix &other) {
        if (other.rows != cols || other.cols != rows) { std::cout << "Matrix can not be transposed. Exiting ..." << std::endl; exit(1); }
        char *d = static_cast<char *>(other.buf.dev_out());
        int8_t *ov = reinterpret_cast<int8_t *>(d);
        float *os = reinterpret_cast<float *>(d + other.value_bytes());
        const int rc = BITS == 4 ? clover_m4_transpose(device_values(), device_scales(), rows, cols, ov, os, nullptr)
                                 : clover_m8_transpose(device_values(), device_scales(), rows, cols, ov, os, nullptr);
        clover_b200_detail::check(rc, "transpose");
        other.buf.flush_view();
    }
    void transpose_scalar(CloverQuantizedMatrix &other) { transpose(other); }
    void transpose_parallel(CloverQuantizedMatrix &other) { transpose(other); }
};

class CloverMatrix4 : public CloverQuantizedMatrix<4, CloverVector4> {
public:
    using CloverQuantizedMatrix<4, CloverVector4>::CloverQuantizedMatrix;
    using CloverQuantizedMatrix<4, CloverVector4>::mvm;
    float get(uint64_t i, uint64_t j) const {               // include/CloverMatrix4.h:123-139
        const uint64_t pos = i * cols + j;
        const int8_t b = reinterpret_cast<const int8_t *>(buf.host_ro())[pos >> 1];
        const int8_t q = (pos & 1) ? (int8_t)((int8_t)(b << 4) >> 4) : (int8_t)(b >> 4);
        const float *s = reinterpret_cast<const float *>(buf.host_ro() + value_bytes());
        return (s[(i >> 6) * (cols >> 6) + (j >> 6)] / 7.0f) * (float)q;
    }
    // mixed precision, include/CloverMatrix4.h:1093-1441 (mvm_parallel :2017): 4-bit matrix x 8-bit vector -> 8-bit vector
    void mvm(const CloverVector8 &productVector, CloverVector8 &resultVector) {
        if (productVector.size() != getCols() || resultVector.size_pad() != getRows()) {
            std::cout << "MVM can not be performed. Exiting ..." << std::endl;
            exit(1);
        }
        int8_t *yv = resultVector.device_values_out();
        float *ys = resultVector.device_scales_out();
        clover_b200_detail::check(clover_m4_mvm_v8(device_values(), device_scales(), rows, cols, productVector.device_values(),
                                                   productVector.device_scales(), yv, ys, nullptr, key_ptr(), nullptr), "mvm");
        resultVector.sync_view();
    }
    void mvm_parallel(const CloverVector8 &x, CloverVector8 &y) { mvm(x, y); }
    using CloverQuantizedMatrix<4, CloverVector4>::mvm_parallel;
    // extension: C = A * Bt^T, C[i][j] = rowView(A,i).dot(rowView(Bt,j)); c_dev: rows x Bt.rows fp32 on the device
    void gemm(const CloverMatrix4 &Bt, float *c_dev, uint64_t ldc) const {
        if (Bt.getCols() != getCols()) { std::cout << "GEMM can not be performed. Exiting ..." << std::endl; exit(1); }
        clover_b200_detail::check(clover_m4_gemm(device_values(), device_scales(), Bt.device_values(), Bt.device_scales(),
                                                 rows, Bt.getRows(), cols, c_dev, ldc, nullptr), "gemm");
    }
};

class CloverMatrix8 : public CloverQuantizedMatrix<8, CloverVector8> {
public:
    using CloverQuantizedMatrix<8, CloverVector8>::CloverQuantizedMatrix;
    using CloverQuantizedMatrix<8, CloverVector8>::mvm;
    float get(uint64_t i, uint64_t j) const {               // include/CloverMatrix8.h:117-129
        const int8_t q = reinterpret_cast<const int8_t *>(buf.host_ro())[i * cols + j];
        const float *s = reinterpret_cast<const float *>(buf.host_ro() + value_bytes());
        return (s[(i >> 6) * (cols >> 6) + (j >> 6)] / 127.0f) * (float)q;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Row-sharded CloverMatrix4 over several GPUs of the node, driven by this one process: mvm() returns the
// same CloverVector4 bytes as CloverMatrix4::mvm on one GPU. No counterpart in the reference.
class ShardedCloverMatrix4 : public clover_b200_detail::Keyed {
    clover_m4_sharded *h = nullptr;
    uint64_t rows, cols;
public:
    ShardedCloverMatrix4(uint64_t r, uint64_t c, int ngpus, const int *devices = nullptr)
        : rows(clover_b200_detail::pad128(r)), cols(clover_b200_detail::pad128(c)) {
        clover_b200_detail::check(clover_m4_sharded_create(&h, rows, cols, ngpus, devices), "clover_m4_sharded_create");
    }
    ShardedCloverMatrix4(const ShardedCloverMatrix4 &) = delete;
    ShardedCloverMatrix4 &operator=(const ShardedCloverMatrix4 &) = delete;
    ~ShardedCloverMatrix4() { clover_m4_sharded_destroy(h); }
    uint64_t getRows() const { return rows; }
    uint64_t getCols() const { return cols; }
    int getGpus() const { return clover_m4_sharded_world(h); }
    // scatter a quantized matrix (host image in the reference layout) to the shards
    void load(const CloverMatrix4 &m) {
        if (m.getRows() != rows || m.getCols() != cols) { std::cout << "Matrices do not have the same size. Exiting ..." << std::endl; exit(1); }
        clover_b200_detail::check(clover_m4_sharded_load_host(h, m.host_values(), m.host_scales()), "clover_m4_sharded_load_host");
    }
    void mvm(const CloverVector4 &productVector, CloverVector4 &resultVector) {
        if (productVector.size() != getCols() || resultVector.size_pad() != getRows()) {
            std::cout << "MVM can not be performed. Exiting ..." << std::endl;
            exit(1);
        }
        // host images: x is read, y is overwritten entirely (values and scales of every block)
        clover_b200_detail::check(clover_m4_sharded_mvm_host(h, productVector.host_values(), productVector.host_scales(),
                                                             resultVector.getData(), resultVector.getScales(), key_ptr()), "mvm");
    }
    void mvm_parallel(const CloverVector4 &x, CloverVector4 &y) { mvm(x, y); }
};

#endif  // CLOVER_B200_CONTAINERS_HPP
