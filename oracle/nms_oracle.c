/* C port of oracle/nms_oracle.py::nms_padded_greedy (and decode) -- TEST INFRASTRUCTURE ONLY, parity unpinned
 * (see oracle/__init__.py).  Restates tf.image.non_max_suppression_padded (TF 2.8.1 image_ops_impl.py) as called by
 * reference core/yolo_nms.py:26-33, and core/yolo_decode_layer.py:4-36.  Build: oracle/Makefile (gcc -O2
 * -ffp-contract=off so every float operation is separately rounded, like TF's elementwise kernels). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float iou_tf(const float* a, const float* b) {
    float ix1 = fmaxf(a[0], b[0]), iy1 = fmaxf(a[1], b[1]);
    float ix2 = fminf(a[2], b[2]), iy2 = fminf(a[3], b[3]);
    float iw = fmaxf(ix2 - ix1, 0.0f), ih = fmaxf(iy2 - iy1, 0.0f);
    float inter = iw * ih;
    float area_a = (a[3] - a[1]) * (a[2] - a[0]);
    float area_b = (b[3] - b[1]) * (b[2] - b[0]);
    float uni = area_a + area_b;
    uni = uni - inter;
    uni = uni + 1e-8f;
    return inter / uni;
}

typedef struct { float s; int32_t i; } item_t;

static int cmp_desc(const void* pa, const void* pb) {
    const item_t* a = (const item_t*)pa; const item_t* b = (const item_t*)pb;
    if (a->s > b->s) return -1;
    if (a->s < b->s) return 1;
    return (a->i > b->i) - (a->i < b->i);   /* ties: lower index first */
}

/* boxes [B,N,4], scores [B,N] -> selected [B,max] (zero padded), num_valid [B] */
int y3o_nms(const float* boxes, const float* scores, int B, int N, int max_boxes, float iou_thr, float score_thr,
            int32_t* selected, int32_t* num_valid) {
    item_t* items = (item_t*)malloc(sizeof(item_t) * (size_t)N);
    float* mb = (float*)malloc(sizeof(float) * 4 * (size_t)N);
    float* kept = (float*)malloc(sizeof(float) * 4 * (size_t)N);
    if (!items || !mb || !kept) return 1;
    for (int b = 0; b < B; ++b) {
        const float* bx = boxes + (size_t)b * N * 4;
        const float* sc = scores + (size_t)b * N;
        int32_t* sel = selected + (size_t)b * max_boxes;
        memset(sel, 0, sizeof(int32_t) * (size_t)max_boxes);
        for (int i = 0; i < N; ++i) {
            int pass = sc[i] > score_thr;
            float m = pass ? 1.0f : 0.0f;
            items[i].s = sc[i] * m + 0.0f;   /* -0 -> +0 is irrelevant for the comparison */
            items[i].i = i;
            for (int k = 0; k < 4; ++k) mb[4 * i + k] = bx[4 * i + k] * m;
        }
        qsort(items, (size_t)N, sizeof(item_t), cmp_desc);
        int nk = 0, ns = 0;
        for (int r = 0; r < N && ns < max_boxes; ++r) {
            const float* c = mb + 4 * (size_t)items[r].i;
            int dead = 0;
            for (int k = 0; k < nk && !dead; ++k) dead = iou_tf(kept + 4 * k, c) >= iou_thr;
            if (dead) continue;
            memcpy(kept + 4 * nk, c, sizeof(float) * 4);
            ++nk;
            if (c[0] > 0.0f || c[1] > 0.0f || c[2] > 0.0f || c[3] > 0.0f) sel[ns++] = items[r].i;
        }
        num_valid[b] = ns;
    }
    free(items); free(mb); free(kept);
    return 0;
}

static float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

/* one scale: grid [B,gh,gw,3,5+C] -> writes records off..off+gh*gw*3 of bboxes [B,N,4], conf [B,N], probs [B,N,C] */
int y3o_decode_scale(const float* grid, int B, int gh, int gw, int C, const float* anchors3x2, int N, int off,
                     float* bboxes, float* conf, float* probs) {
    const int F = 5 + C;
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < gh; ++i)
            for (int j = 0; j < gw; ++j)
                for (int a = 0; a < 3; ++a) {
                    const float* r = grid + ((((size_t)b * gh + i) * gw + j) * 3 + a) * F;
                    size_t o = (size_t)b * N + off + ((size_t)i * gw + j) * 3 + a;
                    float x = (sigmoidf_(r[0]) + (float)j) / (float)gh;   /* reference divides (x,y) by (gh,gw) */
                    float y = (sigmoidf_(r[1]) + (float)i) / (float)gw;
                    float w = expf(r[2]) * anchors3x2[2 * a], h = expf(r[3]) * anchors3x2[2 * a + 1];
                    bboxes[4 * o + 0] = x - w / 2.0f; bboxes[4 * o + 1] = y - h / 2.0f;
                    bboxes[4 * o + 2] = x + w / 2.0f; bboxes[4 * o + 3] = y + h / 2.0f;
                    conf[o] = sigmoidf_(r[4]);
                    for (int c = 0; c < C; ++c) probs[o * C + c] = sigmoidf_(r[5 + c]);
                }
    return 0;
}

/* class_indices = argmax (first max wins), scores = conf * max  (reference core/yolo_nms.py:18-24) */
int y3o_class_reduce(const float* probs, const float* conf, long long nrec, int C, float* scores, int64_t* cls) {
    for (long long r = 0; r < nrec; ++r) {
        const float* p = probs + r * C;
        float best = p[0]; int bi = 0;
        for (int c = 1; c < C; ++c) if (p[c] > best) { best = p[c]; bi = c; }
        scores[r] = conf[r] * best;
        cls[r] = bi;
    }
    return 0;
}
