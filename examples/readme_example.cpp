// The reference README's example (README.md:70-107) compiled against the drop-in containers, followed by a
// reference-style validation loop (test/validate/03_matrix.cpp:248-326) for mvm.
//   g++ -std=c++17 -Iinclude examples/readme_example.cpp -Lclover_b200 -lclover_b200 -Wl,-rpath,$PWD/clover_b200 -o /tmp/readme_example
#include <clover_b200/containers.hpp>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

static void example() {
    const int n = 128;
    CloverVector32 a_vector_32bit(n);
    CloverVector32 b_vector_32bit(n);
    float *a = a_vector_32bit.getData();
    float *b = b_vector_32bit.getData();
    for (int i = 0; i < n; i += 1) {
        a[i] = 1;
        b[i] = 2;
    }
    CloverVector4 a_vector_4bit(128);
    CloverVector4 b_vector_4bit(128);
    a_vector_4bit.quantize(a_vector_32bit);
    b_vector_4bit.quantize(b_vector_32bit);
    float dot = a_vector_4bit.dot(b_vector_4bit);
    std::cout << "The dot product is: " << dot << std::endl;
    if (dot != 256.0f) { std::cout << "unexpected dot" << std::endl; exit(1); }
}

template <class QMatrix, class QVector>
static void validate_mvm(uint64_t m, uint64_t n) {
    CloverMatrix32 a32(m, n);
    CloverVector32 x32(n);
    float *a = a32.getData();
    float *x = x32.getData();
    uint32_t lcg = 12345;
    auto next = [&]() { lcg = lcg * 1664525u + 1013904223u; return (float)((int)(lcg >> 24) % 21 - 10); };
    for (uint64_t i = 0; i < m * n; ++i) a[i] = next();
    for (uint64_t i = 0; i < n; ++i) x[i] = next();
    QMatrix qa(m, n);
    QVector qx(n), qy(m), qy2(m);
    qa.quantize(a32);
    qx.quantize(x32);
    qa.mvm(qx, qy);
    qa.mvm_parallel(qx, qy2);
    for (uint64_t k = 0; k < m; ++k)
        if (qy.get(k) != qy2.get(k)) { std::cout << "mvm mismatch at " << k << std::endl; exit(1); }
    // consistency against a double-precision evaluation of the quantized operands (03_matrix.cpp:328-416 style):
    // the result is a truncating re-quantization of the fp32 row sums, so it may differ from the exact value by
    // up to one quantum of its block (scale / 7 or scale / 127) plus the fp32 summation error
    const double levels = qa.getBitsLength() == 4 ? 7.0 : 127.0;
    double worst = 0;
    for (uint64_t i = 0; i < m; ++i) {
        double s = 0;
        for (uint64_t j = 0; j < n; ++j) s += (double)qa.get(i, j) * (double)qx.get(j);
        const double quantum = (double)qy.getScales()[i >> 6] / levels;
        worst = std::fmax(worst, std::fabs(s - (double)qy.get(i)) / quantum);
    }
    std::cout << "mvm " << m << "x" << n << " bits=" << qa.getBitsLength() << " max error vs fp64: " << worst << " quanta" << std::endl;
    if (worst > 1.001) exit(1);
}

// The shape of the reference's application loop (test/performance/01_measure.h:924-946): gradient step on
// ||y - Phi x||^2 followed by hard thresholding, all operands quantized - exercises transpose, scaleAndAdd (both
// forms), threshold and (QVector = CloverVector8 with a 4-bit matrix) the mixed-precision mvm.
template <class QMatrix, class QVector>
static void iterate(uint64_t m, uint64_t n, uint64_t K) {
    CloverMatrix32 phi32(m, n);
    CloverVector32 y32(m);
    uint32_t lcg = 777;
    auto next = [&]() { lcg = lcg * 1664525u + 1013904223u; return (float)((int)(lcg >> 24) % 21 - 10) * 0.01f; };
    for (uint64_t i = 0; i < m * n; ++i) phi32.getData()[i] = next();
    for (uint64_t i = 0; i < m; ++i) y32.getData()[i] = next();
    QMatrix Phi(m, n), PhiT(n, m);
    Phi.quantize(phi32);
    Phi.transpose(PhiT);
    for (uint64_t i = 0; i < m; i += 37)
        for (uint64_t j = 0; j < n; j += 41)
            if (Phi.get(i, j) != PhiT.get(j, i)) { std::cout << "transpose mismatch" << std::endl; exit(1); }
    QVector x(n), y(m), t1(m), t2(m), t3(n);
    y.quantize(y32);
    x.clear();
    for (int it = 0; it < 3; ++it) {
        Phi.mvm(x, t1);
        y.scaleAndAdd(t1, -1.0f, t2);
        PhiT.mvm(t2, t3);
        x.scaleAndAdd(t3, 0.5f);
        x.threshold(K);
    }
    uint64_t nonzero = 0;
    for (uint64_t i = 0; i < n; ++i) nonzero += x.getBits(i) != 0;
    std::cout << "loop " << m << "x" << n << " K=" << K << " bits=" << x.getBitsLength() << ": " << nonzero << " survivors" << std::endl;
    if (nonzero == 0 || nonzero > K) exit(1);
}

// The borrowing view constructor (include/CloverVector4.h:114-119): the container works on the caller's own host
// buffers - quantize INTO a view, read the bytes through the caller's pointers, use views as operands.
template <class QVector>
static void views(uint64_t n) {
    CloverVector32 a32(n), b32(n);
    for (uint64_t i = 0; i < n; ++i) { a32.getData()[i] = (float)((int)(i % 13) - 6); b32.getData()[i] = (float)((int)(i % 7) - 3); }
    QVector qa(n), qb(n);
    qa.quantize(a32);
    qb.quantize(b32);
    const uint64_t n_pad = qa.size_pad(), vbytes = n_pad * qa.getBitsLength() / 8;
    std::vector<int8_t> va(vbytes), vb(vbytes);
    std::vector<float> sa(n_pad / 64), sb(n_pad / 64);
    QVector wa(n, va.data(), sa.data()), wb(n, vb.data(), sb.data());
    wa.quantize(a32);                                           // lands in va / sa
    if (std::memcmp(va.data(), qa.getData(), vbytes) != 0 || std::memcmp(sa.data(), qa.getScales(), sa.size() * sizeof(float)) != 0) {
        std::cout << "view quantize mismatch" << std::endl; exit(1);
    }
    std::memcpy(vb.data(), qb.getData(), vbytes);               // the caller fills the view's memory itself
    std::memcpy(sb.data(), qb.getScales(), sb.size() * sizeof(float));
    if (wa.dot(wb) != qa.dot(qb) || wb.get(5) != qb.get(5)) { std::cout << "view dot mismatch" << std::endl; exit(1); }
    QVector copy(wa);                                           // deep copy of a view owns its bytes
    va[0] = 0;                                                  // ... so changing the caller's buffer does not reach it
    if (copy.dot(qb) != qa.dot(qb)) { std::cout << "view copy mismatch" << std::endl; exit(1); }
    std::cout << "views n=" << n << " bits=" << qa.getBitsLength() << " ok" << std::endl;
}

// The reference's own generators and the retained-pointer idiom of its harnesses (test/performance/01_measure.h:699-720):
// keys from avx_xorshift128plus_init(445560390295639063, 2935984234003016713) (test/random/00_random.cpp:42), inputs
// drawn with setRandomFloats(-1, 1, key1, key2) - the first draws must be the known answers of SURVEY.md 8c - then
// quantize -> restore / mvm(V32) consistency for both matrix widths.
template <class QMatrix>
static void reference_harness(uint64_t m, uint64_t n) {
    uint64_t key[8];
    if (clover_prng_init(445560390295639063ULL, 2935984234003016713ULL, key) != CLOVER_OK) exit(1);
    CloverVector32 a(4096);
    a.setRandomFloats(-1.0f, 1.0f, key, key + 4);
    if (a.get(0) != 0x1.ff90ap-4f || a.get(1) != -0x1.604778p-3f || a.get(2) != 0x1.8b518p-2f || a.get(3) != 0x1.31567p-2f) {
        std::cout << "generator mismatch: " << a.get(0) << " " << a.get(1) << std::endl; exit(1);
    }
    CloverMatrix32 M(m, n), R(m, n);
    CloverVector32 x(n), y(m);
    M.setRandomFloats(-1.0f, 1.0f, key, key + 4);
    x.setRandomInteger(-10.0f, 10.0f, key, key + 4);
    for (uint64_t i = 0; i < n; ++i) if (x.get(i) != (float)(int)x.get(i) || x.get(i) < -10.f || x.get(i) > 10.f) { std::cout << "setRandomInteger" << std::endl; exit(1); }
    QMatrix Q(m, n);
    Q.quantize(M);
    Q.restore(R);
    for (uint64_t i = 0; i < m; i += 7)
        for (uint64_t j = 0; j < n; j += 5)
            if (R.get(i, j) != Q.get(i, j)) { std::cout << "matrix restore mismatch" << std::endl; exit(1); }
    Q.mvm(x, y);                                                // fp32 vectors (CloverMatrix4.h:1451, CloverMatrix8.h:558)
    for (uint64_t i = 0; i < m; i += 11) {
        double want = 0;
        for (uint64_t j = 0; j < n; ++j) want += (double)R.get(i, j) * (double)x.get(j);
        if (std::fabs(want - (double)y.get(i)) > 0.01) { std::cout << "mvm(V32) mismatch " << want << " " << y.get(i) << std::endl; exit(1); }   // 03_matrix.cpp:419-491
    }
    // fetch the raw pointer ONCE, refill it before each quantize (the reference's idiom): the device copy must follow
    CloverVector32 v(256);
    float *raw = v.getData();
    CloverVector4 q(256);
    for (int round = 1; round <= 3; ++round) {
        for (uint64_t i = 0; i < 256; ++i) raw[i] = (float)round * (float)((int)(i % 15) - 7);
        q.quantize(v);
        if (q.get(14) != (float)round * 7.0f || q.get(0) != (float)round * -7.0f) { std::cout << "retained pointer not re-read" << std::endl; exit(1); }
    }
    std::cout << "reference harness " << m << "x" << n << " bits=" << Q.getBitsLength() << " ok" << std::endl;
}

// The sharded matrix over every GPU this process can see gives the bytes of the single-GPU mvm.
static void sharded(uint64_t m, uint64_t n) {
    const int gpus = std::min(clover_device_count(), 8);
    CloverMatrix32 a32(m, n);
    CloverVector32 x32(n);
    uint64_t key[8];
    clover_prng_init(7, 9, key);
    a32.setRandomFloats(-1.0f, 1.0f, key, key + 4);
    x32.setRandomFloats(-1.0f, 1.0f, key, key + 4);
    CloverMatrix4 A(m, n);
    A.quantize(a32);
    CloverVector4 x(n), y(m), ys(m);
    x.quantize(x32);
    A.mvm(x, y);
    for (int g = 1; g <= gpus; g *= 2) {
        ShardedCloverMatrix4 S(m, n, g);
        S.load(A);
        for (int round = 0; round < 3; ++round) {          // epochs / alternating result buffers
            S.mvm(x, ys);
            if (std::memcmp(y.host_values(), ys.host_values(), y.size_pad() / 2) != 0 ||
                std::memcmp(y.host_scales(), ys.host_scales(), y.size_pad() / 64 * sizeof(float)) != 0) {
                std::cout << "sharded mvm mismatch on " << g << " GPUs" << std::endl; exit(1);
            }
        }
        std::cout << "sharded mvm " << m << "x" << n << " on " << g << " GPU(s) ok" << std::endl;
    }
}

int main() {
    example();
    sharded(1024, 2048);
    sharded(4096, 16384 + 128);
    reference_harness<CloverMatrix4>(256, 384);
    reference_harness<CloverMatrix8>(256, 384);
    validate_mvm<CloverMatrix4, CloverVector4>(256, 384);
    validate_mvm<CloverMatrix8, CloverVector8>(256, 384);
    views<CloverVector4>(1000);
    views<CloverVector8>(1000);
    iterate<CloverMatrix4, CloverVector4>(256, 512, 40);
    iterate<CloverMatrix8, CloverVector8>(256, 512, 40);
    iterate<CloverMatrix4, CloverVector8>(256, 512, 40);      // mixed precision (include/CloverMatrix4.h:1093)
    std::cout << "OK" << std::endl;
    return 0;
}
