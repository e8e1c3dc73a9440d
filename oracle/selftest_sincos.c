/* CPU ORACLE self-test (test infrastructure): so_sinf/so_cosf vs the host libm over every finite
 * float (or a stride of them: argv[1]).  Prints mismatch counts; exit 0 iff none. */
#include "stitch_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int main(int argc, char **argv)
{
    unsigned stride = argc > 1 ? (unsigned)atoi(argv[1]) : 1;
    unsigned long long n = 0, bad_s = 0, bad_c = 0;
    for (unsigned long long b = 0; b < 0x100000000ULL; b += stride) {
        unsigned u = (unsigned)b; float x, a, r; unsigned ua, ur;
        memcpy(&x, &u, 4);
        if (!isfinite(x)) continue;
        ++n;
        a = so_sinf(x); r = sinf(x); memcpy(&ua, &a, 4); memcpy(&ur, &r, 4);
        if (ua != ur) { if (bad_s < 5) printf("sinf(%a): oracle %a libm %a\n", x, a, r); ++bad_s; }
        a = so_cosf(x); r = cosf(x); memcpy(&ua, &a, 4); memcpy(&ur, &r, 4);
        if (ua != ur) { if (bad_c < 5) printf("cosf(%a): oracle %a libm %a\n", x, a, r); ++bad_c; }
    }
    printf("checked %llu floats: sinf mismatches %llu, cosf mismatches %llu\n", n, bad_s, bad_c);
    return (bad_s || bad_c) ? 1 : 0;
}
