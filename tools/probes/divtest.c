#include <stdio.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <omp.h>
static inline float u2f (uint32_t u) { float f; memcpy (&f, &u, 4); return f; }
static inline uint32_t f2u (float f) { uint32_t u; memcpy (&u, &f, 4); return u; }
static inline uint64_t mix (uint64_t z) { z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static inline float fastdiv (float x, float d, float r)
{
    volatile float q = x * r;
    float rem = fmaf (-d, q, x);
    return fmaf (rem, r, q);
}
int main (void)
{
    long bad = 0, total = 0;
    /* exhaustive over the mantissa of d in [1,2), NX random x in [1,4) with random sign each */
    const int NX = 768;
    #pragma omp parallel for schedule(dynamic,4096) reduction(+:bad,total)
    for (uint32_t md = 0; md < (1u << 23); md++)
    {
        float d = u2f (0x3f800000u | md);
        float r = 1.0f / d;
        uint64_t s = mix (md);
        for (int k = 0; k < NX; k++)
        {
            s = mix (s);
            uint32_t mx = (uint32_t) s & 0x7fffffu;
            uint32_t ex = 127 + ((s >> 23) & 1);
            float x = u2f ((ex << 23) | mx);
            if ((s >> 24) & 1) x = -x;
            float want = x / d, got = fastdiv (x, d, r);
            total++;
            if (f2u (want) != f2u (got)) { bad++; if (bad < 10) printf ("mismatch x=%a d=%a want=%a got=%a\n", x, d, want, got); }
        }
    }
    printf ("random-x: %ld mismatches of %ld\n", bad, total);
    /* exhaustive x for a few special d: all-ones mantissa, 1.0, 1+ulp, 1.5, sqrt2-ish */
    uint32_t special[] = {0x7fffffu, 0x7ffffeu, 0x000000u, 0x000001u, 0x400000u, 0x3504f3u, 0x3504f4u, 0x2aaaabu, 0x555555u, 0x7ffffdu};
    for (unsigned i = 0; i < sizeof special / sizeof *special; i++)
    {
        float d = u2f (0x3f800000u | special[i]);
        float r = 1.0f / d;
        long b = 0;
        for (uint32_t mx = 0; mx < (1u << 24); mx++)
        {
            float x = u2f (0x3f800000u + mx);   /* [1,4) */
            if (f2u (x / d) != f2u (fastdiv (x, d, r))) { b++; if (b < 3) printf ("  special d=%a x=%a want %a got %a\n", d, x, x / d, fastdiv (x, d, r)); }
        }
        printf ("d=%a: %ld mismatches of %d\n", d, b, 1 << 24);
    }
    return 0;
}
