// The reference README's example (README.md:70-107) compiled against the drop-in containers, followed by a
// reference-style validation loop (test/validate/03_matrix.cpp:248-326) for mvm.
//   g++ -std=c++17 -Iinclude examples/readme_example.cpp -Lclover_b200 -lclover_b200 -Wl,-rpath,$PWD/clover_b200 -o /tmp/readme_example
#include <clover_b200/containers.hpp>

#include <cmath>

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

int main() {
    example();
    validate_mvm<CloverMatrix4, CloverVector4>(256, 384);
    validate_mvm<CloverMatrix8, CloverVector8>(256, 384);
    std::cout << "OK" << std::endl;
    return 0;
}
