/*
 * Minimal stand-in for Intel IPP, used ONLY to compile the untouched reference headers
 * (/root/reference/include) into oracle/_ref/. The 4/8-bit quantize/dot/mvm path never
 * calls IPP; the reference only uses it for an init banner and for transposes
 * (CloverBase.h:250-350, CloverMatrix4.h:496). TEST INFRASTRUCTURE - not product code.
 */
#ifndef CLOVER_B200_ORACLE_IPP_STUB_H
#define CLOVER_B200_ORACLE_IPP_STUB_H
#include <stdint.h>
typedef uint64_t Ipp64u;
typedef uint8_t  Ipp8u;
typedef uint16_t Ipp16u;
typedef float    Ipp32f;
typedef int      IppStatus;
typedef struct { int width; int height; } IppiSize;
typedef struct { const char *Name; const char *Version; } IppLibraryVersion;
enum { ippStsNoErr = 0 };
enum {
    ippCPUID_MMX = 1 << 0, ippCPUID_SSE = 1 << 1, ippCPUID_SSE2 = 1 << 2, ippCPUID_SSE3 = 1 << 3,
    ippCPUID_SSSE3 = 1 << 4, ippCPUID_MOVBE = 1 << 5, ippCPUID_SSE41 = 1 << 6, ippCPUID_SSE42 = 1 << 7,
    ippCPUID_AVX = 1 << 8, ippAVX_ENABLEDBYOS = 1 << 9, ippCPUID_AES = 1 << 10, ippCPUID_CLMUL = 1 << 11,
    ippCPUID_RDRAND = 1 << 13, ippCPUID_F16C = 1 << 14, ippCPUID_AVX2 = 1 << 15, ippCPUID_ADCOX = 1 << 16,
    ippCPUID_RDSEED = 1 << 17, ippCPUID_PREFETCHW = 1 << 18, ippCPUID_SHA = 1 << 19, ippCPUID_AVX512F = 1 << 20,
    ippCPUID_AVX512CD = 1 << 21, ippCPUID_AVX512ER = 1 << 22, ippCPUID_KNC = 1 << 23
};
static inline IppStatus ippInit(void) { return ippStsNoErr; }
static inline IppStatus ippSetNumThreads(int) { return ippStsNoErr; }
static inline const IppLibraryVersion *ippGetLibVersion(void) {
    static const IppLibraryVersion v = { "ipp-stub (clover_b200 oracle)", "0" };
    return &v;
}
static inline IppStatus ippGetCpuFeatures(Ipp64u *mask, void *) { *mask = 0; return ippStsNoErr; }
static inline Ipp64u ippGetEnabledCpuFeatures(void) { return 0; }
template <typename T>
static inline IppStatus ipp_stub_transpose(const T *src, int srcStep, T *dst, int dstStep, IppiSize roi) {
    for (int y = 0; y < roi.height; ++y) {
        const T *s = (const T *)((const char *)src + (int64_t)y * srcStep);
        for (int x = 0; x < roi.width; ++x) {
            T *d = (T *)((char *)dst + (int64_t)x * dstStep);
            d[y] = s[x];
        }
    }
    return ippStsNoErr;
}
static inline IppStatus ippiTranspose_8u_C1R(const Ipp8u *s, int ss, Ipp8u *d, int ds, IppiSize r) { return ipp_stub_transpose(s, ss, d, ds, r); }
static inline IppStatus ippiTranspose_16u_C1R(const Ipp16u *s, int ss, Ipp16u *d, int ds, IppiSize r) { return ipp_stub_transpose(s, ss, d, ds, r); }
static inline IppStatus ippiTranspose_32f_C1R(const Ipp32f *s, int ss, Ipp32f *d, int ds, IppiSize r) { return ipp_stub_transpose(s, ss, d, ds, r); }
#endif
