/*
 * zso_adler32.c -- CPU ORACLE (test infrastructure only, see zso.h).
 * Adler-32 as Adler32.cs:270-326 (scalar path; the SSSE3/AVX2 paths compute the same function) and the
 * standard adler32_combine, which the reference does not have.
 */
#include "zso.h"

#define BASE 65521u /* Adler32.cs:27 */
#define NMAX 5552u  /* Adler32.cs:30 */

uint32_t zso_adler32(uint32_t seed, const uint8_t* p, size_t n)
{
    uint32_t s1 = seed & 0xFFFF, s2 = seed >> 16;
    if (n == 0) return seed; /* Adler32.cs:63-66 */
    while (n > 0) {
        size_t k = n < NMAX ? n : NMAX;
        n -= k;
        while (k--) {
            s1 += *p++;
            s2 += s1;
        }
        s1 %= BASE;
        s2 %= BASE;
    }
    return (s2 << 16) | s1;
}

uint32_t zso_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2)
{
    uint32_t rem = (uint32_t)(len2 % BASE);
    uint32_t sum1 = a1 & 0xFFFF;
    uint32_t sum2 = (rem * sum1) % BASE;
    sum1 += (a2 & 0xFFFF) + BASE - 1;
    sum2 += (a1 >> 16) + (a2 >> 16) + BASE - rem;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum2 >= (BASE << 1)) sum2 -= (BASE << 1);
    if (sum2 >= BASE) sum2 -= BASE;
    return sum1 | (sum2 << 16);
}
