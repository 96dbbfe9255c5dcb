/* Test infrastructure: checks that the division-free `v / 255` of yolo_v3_tf2_b200/csrc/preprocess.cuh (div255:
 * q0 = v * RN(1/255); e = fma(-q0, 255, v); q = fma(e, RN(1/255), q0)) equals the IEEE quotient the reference computes
 * (`tf.image.resize(...) / 255`, core/load_tfrecords.py:46) for every float in [0, 256].
 * usage: div255_check [stride]   -- stride 1 = all 1 132 462 081 values (34 s on one core, 0 mismatches). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline float asf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t asu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
int main(int argc, char** argv) {
    const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1u;
    const float r = 0.003921568859368563f;
    if (asu(r) != 0x3B808081u || asu(1.0f / 255.0f) != 0x3B808081u) { printf("unexpected reciprocal bits\n"); return 2; }
    const uint32_t hi = asu(256.0f);
    unsigned long long bad = 0, n = 0;
    for (uint32_t u = 0; u <= hi; u += stride, ++n) {
        const float v = asf(u);
        const float q0 = v * r;
        const float e = fmaf(-q0, 255.0f, v);
        const float q = fmaf(e, r, q0);
        if (asu(q) != asu(v / 255.0f)) ++bad;
    }
    printf("checked %llu values, mismatches %llu\n", n, bad);
    return bad != 0;
}
