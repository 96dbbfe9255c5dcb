// Detection-vs-ground-truth counters of the reference's evaluation:
// EvaluateDetections.evaluate / calc_iou / process_decisions / update_counters, evaluate_detections.py:39-48, 57-135.
// Per image (one CTA): iou[p, g] = overlap / (area_p + area_g - overlap); every prediction picks its best ground-truth
// box (first maximum); it is a true positive when that IoU > iou_thresh and the classes match -- the reference evaluates
// all predictions of an image at once against an all-False "assigned" list, so several predictions may claim the same
// ground-truth box and all count; a ground-truth box is a false negative when no prediction claimed it.
// Counters are per class: preds, gts, tp, fp, fn (int32, accumulated with atomics across images and calls).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace y3 {

struct EvalArgs {
    const float* det_boxes;      // [B, max_det, 4]
    const long long* det_cls;    // [B, max_det]
    const int* num_det;          // [B]
    const float* gt_boxes;       // [B, max_gt, 4]
    const int* gt_cls;           // [B, max_gt]
    const int* num_gt;           // [B]
    int B, max_det, max_gt, nclasses;
    float iou_thresh;
    int* preds; int* gts; int* tp; int* fp; int* fn;   // [nclasses] each
    int* examples;               // [1]
    int* errors;                 // [1] images skipped because of a class id outside [0, nclasses)
};

__device__ __forceinline__ float iou_eval(const float4 a, const float4 b) {
    const float ow = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
    const float oh = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
    const float ov = __fmul_rn(ow, oh);
    const float aa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float ab = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    return __fdiv_rn(ov, __fsub_rn(__fadd_rn(aa, ab), ov));
}

__global__ void __launch_bounds__(128) evaluate_kernel(const EvalArgs a) {
    extern __shared__ int esm[];
    int* assigned = esm;                 // [max_gt]
    __shared__ int bad;
    const int img = blockIdx.x;
    const int nd = min(a.num_det[img], a.max_det), ng = min(a.num_gt[img], a.max_gt);
    const float4* db = reinterpret_cast<const float4*>(a.det_boxes) + (long long)img * a.max_det;
    const float4* gb = reinterpret_cast<const float4*>(a.gt_boxes) + (long long)img * a.max_gt;
    const long long* dc = a.det_cls + (long long)img * a.max_det;
    const int* gc = a.gt_cls + (long long)img * a.max_gt;
    if (threadIdx.x == 0) bad = 0;
    for (int g = threadIdx.x; g < ng; g += blockDim.x) assigned[g] = 0;
    __syncthreads();
    // the reference skips the whole sample when a class id cannot index the counters (update_counters' except branch)
    for (int g = threadIdx.x; g < ng; g += blockDim.x)
        if (gc[g] < 0 || gc[g] >= a.nclasses) bad = 1;
    for (int p = threadIdx.x; p < nd; p += blockDim.x)
        if (dc[p] < 0 || dc[p] >= a.nclasses) bad = 1;
    __syncthreads();
    if (bad) {
        if (threadIdx.x == 0) atomicAdd(a.errors, 1);
        return;
    }
    for (int p = threadIdx.x; p < nd; p += blockDim.x) {
        const float4 pb = db[p];
        float best = -INFINITY;
        int bi = 0;
        for (int g = 0; g < ng; ++g) {
            const float v = iou_eval(pb, gb[g]);
            if (v > best) { best = v; bi = g; }     // first maximum wins (tf.math.argmax)
        }
        const int pc = (int)dc[p];
        const bool hit = ng > 0 && best > a.iou_thresh && gc[bi] == pc;
        if (hit) { atomicAdd(&a.tp[pc], 1); atomicOr(&assigned[bi], 1); }
        else atomicAdd(&a.fp[pc], 1);
        atomicAdd(&a.preds[pc], 1);
    }
    __syncthreads();
    for (int g = threadIdx.x; g < ng; g += blockDim.x) {
        atomicAdd(&a.gts[gc[g]], 1);
        if (!assigned[g]) atomicAdd(&a.fn[gc[g]], 1);
    }
    if (threadIdx.x == 0) atomicAdd(a.examples, 1);
}

}  // namespace y3
