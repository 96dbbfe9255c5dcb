/* y3b200 -- C ABI of the B200-native YOLOv3 inference hot path.
 *
 * The reference (ronen-halevy/yolo-v3-tf2) has no FFI: its hot path is three Python/Keras call sites.  Each entry point
 * below replaces one of them; the Python package yolo_v3_tf2_b200.core.* binds these symbols with ctypes and mirrors
 * the reference signatures (see INTEGRATION.md for the reference-side stub).
 *
 *   reference interface                                              replaced by
 *   ---------------------------------------------------------------  -------------------------------------------
 *   ParseModel.build_model       core/parse_model.py:279-314         y3_net_create / y3_net_load_conv / y3_net_plan_*
 *   model(x) / model.predict(x)  inference.py:109,127,162            y3_net_forward
 *   yolo_decode                  core/yolo_decode_layer.py:15-36     y3_decode
 *   yolo_nms (argmax/max/score)  core/yolo_nms.py:18-24              y3_class_reduce
 *   tf.image.non_max_suppression_padded via yolo_nms :26-33          y3_nms
 *   Inference.gather_valid_detections_results inference.py:21-28     y3_gather_detections
 *   one Conv2D(+BN+LeakyReLU+Add+UpSampling2D) core/parse_model.py:13-75,143-160   y3_conv2d_bf16 (unit-test entry)
 *
 * Conventions: every function returns 0 on success or a Y3_ERR_* code; y3_last_error() returns a thread-local message.
 * All tensor pointers are DEVICE pointers unless the parameter name ends in _host.  The caller owns every input and
 * output buffer; the library owns the context, packed weights, the activation arena and its tensor maps.  Calls are
 * asynchronous on the given stream (cudaStream_t passed as void*; NULL = default stream).  A context is bound to one
 * GPU and is not thread-safe.  There is no CPU fallback anywhere: on a machine without an sm_100 GPU every compute
 * entry point fails with Y3_ERR_CUDA / Y3_ERR_UNSUPPORTED.
 */
#ifndef Y3B200_H
#define Y3B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Y3_OK 0
#define Y3_ERR_INVALID 1      /* bad argument / malformed graph (Python raises ValueError) */
#define Y3_ERR_UNSUPPORTED 2  /* valid but not implemented on this path (e.g. maxpool) */
#define Y3_ERR_CUDA 3         /* CUDA runtime / driver failure, message has the CUDA error string */
#define Y3_ERR_STATE 4        /* call order violated (weights missing, net not planned, ...) */

typedef struct y3_ctx y3_ctx;
typedef struct y3_net y3_net;

/* ---- layer list: the flattened form of the reference's per-sub-model yaml layer lists ---------------------------
 * Tensor ids: 0 is the network input image; layer i (0-based) produces tensor i+1.
 * Y3_OP_CONV     src0            conv(k, stride, filters) [+BN] [+leaky]   parse_model.py:13-56
 * Y3_OP_SHORTCUT src0=x src1=from    x + from                              parse_model.py:143-160
 * Y3_OP_UPSAMPLE src0            nearest x2                                parse_model.py:59-75
 * Y3_OP_CONCAT   src0,src1       channel concat [src0, src1]               parse_model.py:102-140
 * Y3_OP_YOLO     src0            [B,g,g,3*(5+C)] -> [B,g,g,3,5+C] view; marks a network output   parse_model.py:163-213
 * Y3_OP_MAXPOOL  src0            max pool ksize x ksize, stride, pad: 1 'same' / 0 'valid' (yolov3-tiny)   parse_model.py:78-99
 */
enum { Y3_OP_CONV = 0, Y3_OP_SHORTCUT = 1, Y3_OP_UPSAMPLE = 2, Y3_OP_CONCAT = 3, Y3_OP_YOLO = 4, Y3_OP_MAXPOOL = 5 };

typedef struct {
    int32_t op;
    int32_t src0;
    int32_t src1;             /* -1 when unused */
    int32_t ksize;            /* conv: 1 or 3 */
    int32_t stride;           /* conv: 1 or 2 ; upsample: 2 */
    int32_t filters;          /* conv: output channels */
    int32_t pad;              /* conv: yaml 'pad'; stride 1 and pad==1 -> 'same', stride 1 and pad==0 -> 'valid',
                                 stride 2 -> ZeroPadding2D(((1,0),(1,0))) + 'valid' (parse_model.py:29-35) */
    int32_t batch_normalize;  /* conv: 1 = BN follows (no bias), 0 = conv bias */
    int32_t activation;       /* conv: 0 linear, 1 leaky(0.1) */
} y3_layer_desc;

/* what the planner decided for one layer; lets CPU-only tests check the host logic without a GPU */
typedef struct {
    int32_t H, W, C;          /* output shape of the layer's tensor */
    int32_t kernel;           /* 0 none (fused/view), 1 tcgen05 conv, 2 direct conv, 3 add, 4 upsample, 5 copy, 6 maxpool */
    int32_t fused_add;        /* conv: residual tensor id added in the epilogue, else -1 */
    int32_t fused_upsample;   /* conv: 1 if the epilogue writes the 2x upsampled tensor */
    int32_t buffer;           /* arena buffer id holding this tensor, -1 for views of fused ops / outputs */
    int32_t chan_offset;      /* channel offset inside the buffer (concat slices) */
    int32_t pix_stride;       /* channels between consecutive pixels of the buffer */
    int32_t block_n, swizzle, stages;   /* tcgen05 tile configuration */
    int32_t flat;             /* conv: 1 = flat-patch 3x3 kernel (input staged once per channel block, taps = row shifts) */
    int32_t padded;           /* tensor stored as [B, H+1, W+1, C] with a zero last row / column */
    int64_t arena_offset;     /* byte offset of the buffer inside the activation arena */
} y3_layer_plan;

const char* y3_last_error(void);
int y3_version(void);

/* CRC-32C (Castagnoli) of data[0..n), continuing from crc (0 to start): host helper of the TensorFlow checkpoint
 * reader / writer that stands in for model.load_weights / save_weights (inference.py:102, train.py:93-104). */
uint32_t y3_crc32c(uint32_t crc, const void* data, int64_t n);

/* device < 0 creates a planning-only context (no CUDA calls; usable on a CPU-only machine for y3_net_plan_*). */
int y3_ctx_create(int device, y3_ctx** out);
void y3_ctx_destroy(y3_ctx* ctx);
int y3_ctx_sm_count(y3_ctx* ctx);

/* Build the network plan for images of H x W, up to max_batch images, nclasses classes.  Validates the graph
 * (Y3_ERR_INVALID mirrors the reference's ValueError / AssertionError cases) and, on a GPU context, allocates the
 * activation arena and weight storage. */
int y3_net_create(y3_ctx* ctx, const y3_layer_desc* layers, int n_layers, int H, int W, int max_batch, int nclasses,
                  y3_net** out);
void y3_net_destroy(y3_net* net);
int y3_net_num_convs(y3_net* net);
int y3_net_num_outputs(y3_net* net);
int y3_net_get_plan(y3_net* net, y3_layer_plan* plans_host, int n_layers);

/* Layer chaining (csrc/conv_tc.cuh, ChainArgs): how each launch ("step", y3_net_num_steps of them) of a forward pass
 * of B images synchronises with its producers and in which order it walks its output tiles.  A chained step waits tile
 * by tile for the 128-row blocks of its input instead of for its whole predecessor (griddepcontrol.wait).  Planning
 * only; works on a context without a device. */
typedef struct y3_chain_step {
    int32_t layer;           /* layer index of the step */
    int32_t posts;           /* 1: its epilogue posts per-tile completion flags */
    int32_t chained;         /* 1: waits on those flags (dep_step, res_step) */
    int32_t dep_step;        /* step that produces its input, -1 */
    int32_t res_step;        /* step that produces its fused residual, -1 */
    int32_t tiles;           /* output tiles of the launch */
    int32_t ctas;            /* CTAs (or CTA pairs) walking them */
    int32_t tiles_n;         /* N tiles per M group */
    int32_t rows_per_group;  /* output rows per M group: 128, or 256 for CTA pairs */
    int32_t rev, rot;        /* sequence position q is tile (q + rot) mod tiles, counted from the end when rev */
    int32_t run_first;       /* persistent run (csrc/conv_chain.cuh): first step of the launch this step belongs to, -1 */
    int32_t run_len;         /* steps in that launch */
    int32_t vshift;          /* CTA pair c walks the tile sequence of pair (c + vshift) mod ctas */
} y3_chain_step;
int y3_net_chain_plan(y3_net* net, int B, y3_chain_step* steps_host, int n_steps);
int64_t y3_net_arena_bytes(y3_net* net);
/* Debugging / parity aid: copy the activation tensor that layer `layer` materialised during the LAST forward pass of B
 * images to host memory as dense bf16 [B,H,W,C] (H, W, C as in y3_layer_plan; stored padding channels are dropped).
 * Arena buffers are recycled, so the tensor is only intact if no later layer of the pass reused its buffer (true for
 * every tensor that is still read by one of the last three launches).  Synchronises the device.  What the reference
 * offers through Keras sub-model outputs (core/parse_model.py:279-314). */
int y3_net_read_layer(y3_net* net, int layer, int B, void* host_bf16);
/* output k: grid height/width and channel count 3*(5+C) */
int y3_net_output_shape(y3_net* net, int k, int* gh, int* gw, int* ch);

/* Weights of conv number conv_idx (creation order, = Keras conv2d_<idx> and Darknet file order, convert.py:93-137) in
 * the reference layout: kernel HWIO fp32 (kh,kw,Cin,Cout); either bias[Cout] or the four BN vectors (Keras order
 * gamma, beta, moving_mean, moving_variance) with epsilon (Keras default 1e-3).  BN is folded, the result rounded to
 * bf16 and packed for the kernel.  Host pointers. */
int y3_net_load_conv(y3_net* net, int conv_idx, const float* kernel_hwio_host, const float* bias_host,
                     const float* bn_gamma_host, const float* bn_beta_host, const float* bn_mean_host,
                     const float* bn_var_host, float bn_eps);

/* x: [B,H,W,3] fp32 NHWC in [0,1].  outs[k]: [B,gh_k,gw_k,3,5+C] fp32 for every output in model order. */
int y3_net_forward(y3_net* net, const float* x, int B, float* const* outs, int n_outs, void* stream);

/* Same, with a pixel pitch (in floats) per output: out k is [B, gh, gw, out_pitch[k]] with the 3*(5+C) logits of a pixel
 * at its start.  A pitch that is a multiple of 4 floats (e.g. 256 for C = 80) lets the head convs write through the
 * TMA-store epilogue instead of unaligned 255-float rows; y3_decode_pitched reads that layout.  pitch == 3*(5+C) is
 * y3_net_forward. */
int y3_net_forward_pitched(y3_net* net, const float* x, int B, float* const* outs, const int* out_pitch, int n_outs,
                           void* stream);

/* Same for a uint8 image x [B,H,W,3] already at the network resolution: computes the network on float32(x) / 255, the
 * serving input of the reference (inference.py:157-158 resize(...) / 255, core/load_tfrecords.py:46), bit-identical to
 * y3_net_forward_pitched on that float tensor, with a quarter of the input bytes.  out_pitch may be NULL (dense
 * outputs as y3_net_forward). */
int y3_net_forward_u8(y3_net* net, const uint8_t* x, int B, float* const* outs, const int* out_pitch, int n_outs,
                      void* stream);

/* Profiling aid: same as y3_net_forward with a CUDA event between kernels; after the call ms_host[i] is the device time
 * of kernel i and layer_host[i] the layer index it implements (n_steps = y3_net_num_steps). Synchronises the stream. */
int y3_net_num_steps(y3_net* net);
int y3_net_forward_timed(y3_net* net, const float* x, int B, float* const* outs, int n_outs, void* stream,
                         float* ms_host, int32_t* layer_host, int n_steps);

/* grids[s]: [B,gh[s],gw[s],3,5+C] fp32; anchors_host: 3x3x2 fp32 (scale, anchor, (w,h)) image fractions.
 * bboxes [B,N,4], conf [B,N,1], probs [B,N,C]; scores [B,N] and class_idx [B,N] (int64) optional (both or neither).
 * "Compact" decode: conf and probs may both be NULL when scores / class_idx are given -- what yolo_nms (core/yolo_nms.py:
 * 18-33) needs of the decode output is boxes, scores and class ids only. */
int y3_decode(y3_ctx* ctx, const float* const* grids, const int* gh, const int* gw, int n_scales,
              const float* anchors_host, int B, int nclasses, float* bboxes, float* conf, float* probs, float* scores,
              int64_t* class_idx, void* stream);

/* y3_decode on grids stored with a pixel pitch (floats) per scale, as written by y3_net_forward_pitched. */
int y3_decode_pitched(y3_ctx* ctx, const float* const* grids, const int* gh, const int* gw, const int* pix_pitch,
                      int n_scales, const float* anchors_host, int B, int nclasses, float* bboxes, float* conf,
                      float* probs, float* scores, int64_t* class_idx, void* stream);

int y3_class_reduce(y3_ctx* ctx, const float* probs, const float* conf, int B, int N, int nclasses, float* scores,
                    int64_t* class_idx, void* stream);

/* tf.image.non_max_suppression_padded(pad_to_max_output_size=True) as called at core/yolo_nms.py:26-33.
 * selected [B,max_boxes] int32 zero padded, num_valid [B] int32, status [B] int32 (0 ok, 1 = kept-list overflow).
 * Preconditions (Y3_ERR_UNSUPPORTED otherwise where checkable): max_boxes <= 768; iou_thr > 0 (TF suppresses on
 * iou >= thr only where iou > 0); boxes are canonical corners (x1 <= x2, y1 <= y2) -- TF's whole-batch coordinate
 * swap keyed on the first box of the first image is NOT reproduced (yolo_decode always emits canonical boxes). */
int y3_nms(y3_ctx* ctx, const float* bboxes, const float* scores, int B, int N, int max_boxes, float iou_thr,
           float score_thr, int32_t* selected, int32_t* num_valid, int32_t* status, void* stream);

/* Evaluation counters (reference evaluate_detections.py:39-135, EvaluateDetections.evaluate) for a batch: detections
 * as produced by y3_gather_detections ([B,max_det,4], [B,max_det] int64, num_det [B]), ground truth [B,max_gt,4] (same
 * corner order), [B,max_gt] int32, num_gt [B].  counters: int32 [5*nclasses + 2] = preds, gts, tp, fp, fn per class,
 * then examples, errors; accumulated (atomics), the caller zeroes them once. */
int y3_evaluate(y3_ctx* ctx, const float* det_boxes, const int64_t* det_classes, const int32_t* num_det, int max_det,
                const float* gt_boxes, const int32_t* gt_classes, const int32_t* num_gt, int max_gt, int B, int nclasses,
                float iou_thresh, int32_t* counters, void* stream);

/* Input pre-processing (reference inference.py:157-158, core/load_tfrecords.py:46, core/utils.py:17-28): for each of B
 * images, tf.image.resize-compatible bilinear resampling (half-pixel centres, no antialias) of a uint8 / float32
 * [H, W, 3] device image to out_h x out_w, placed at (off_y, off_x) of a zero-filled dst_h x dst_w canvas
 * (pad_to_bounding_box), optionally divided by 255.  image_descs_dev: device array of B records of 10 int64
 * {src pointer, H, W, dtype (0 uint8, 1 float32), out_h, out_w, off_y, off_x, bits of float32(H)/float32(out_h),
 * bits of float32(W)/float32(out_w)}.  out: [B, dst_h, dst_w, 3] float32. */
int y3_preprocess(y3_ctx* ctx, const void* image_descs_dev, int B, int dst_h, int dst_w, int divide_by_255, float* out,
                  void* stream);

int y3_gather_detections(y3_ctx* ctx, const float* bboxes, const int64_t* class_idx, const float* scores,
                         const int32_t* selected, const int32_t* num_valid, int B, int N, int max_boxes,
                         float* out_boxes, int64_t* out_classes, float* out_scores, void* stream);

/* Same, additionally writing the packed float32 records [B, max_boxes*6 + 1] (x1,y1,x2,y2,score,class per slot, then
 * num_valid) that the multi-GPU detection gather sends (packed may be NULL). */
int y3_gather_detections_packed(y3_ctx* ctx, const float* bboxes, const int64_t* class_idx, const float* scores,
                                const int32_t* selected, const int32_t* num_valid, int B, int N, int max_boxes,
                                float* out_boxes, int64_t* out_classes, float* out_scores, float* packed, void* stream);

/* One fused conv layer on bf16 NHWC views (unit-test entry; the net executor runs the same kernel).
 * x: [B,H,W,Cin] bf16 with pixel stride x_stride elements.  w_packed: [Cout_pad][k][k][Cin] bf16 (Cout_pad = Cout
 * rounded up to the tile width returned by y3_conv_block_n).  bias: [Cout_pad] fp32.  residual optional.
 * out: bf16 (or fp32) view with pixel stride out_stride; upsample=1 writes the (2Ho,2Wo) nearest-upsampled tensor. */
int y3_conv_block_n(int cin, int cout);
int y3_conv2d_bf16(y3_ctx* ctx, const void* x, int B, int H, int W, int Cin, int64_t x_stride, const void* w_packed,
                   const float* bias, int ksize, int stride, int Cout, int leaky, const void* residual,
                   int64_t res_stride, void* out, int64_t out_stride, int out_fp32, int upsample, void* stream);

/* 3x3 stride-1 'same' conv on a haloed-flat input x_padded [B, H+1, W+1, Cin] (zero last row / column), weights packed
 * [Cout_pad][Cin/BK][3][3][BK] (BK = 64 if Cin % 64 == 0 else 32); output / residual dense.  Unit-test entry. */
int y3_conv2d_flat_bf16(y3_ctx* ctx, const void* x_padded, int B, int H, int W, int Cin, const void* w_packed,
                        const float* bias, int Cout, int leaky, const void* residual, int64_t res_stride, void* out,
                        int64_t out_stride, void* stream);

/* The 3-channel stem conv (3x3, 32 filters) on the tensor cores: x fp32 [B,H,W,3]; w_packed bf16 [32][64] with the
 * 27 BN-folded weights of output o at columns 0..26 AND 32..58 (the kernel multiplies bf16(x) and the bf16 remainder
 * x - bf16(x) against the same weights, so the fp32 image keeps its precision).  For stride 1 the library re-orders
 * the columns internally for its column-sharing producer; the caller's layout is the one above either way.
 * Unit-test entry. */
int y3_conv2d_stem_f32(y3_ctx* ctx, const float* x, int B, int H, int W, const void* w_packed, const float* bias,
                       int stride, int leaky, void* out, int64_t out_stride, void* stream);

/* Debug: fetch one 128-pixel x (swizzle/2)-channel A tile through the conv kernel's TMA path and return the raw
 * (swizzled) shared-memory image, 128*swizzle bytes. */
int y3_dbg_tma_tile(y3_ctx* ctx, const void* x, int B, int H, int W, int Cin, int64_t x_stride, int ksize, int stride,
                    int swizzle, int tap_r, int tap_s, int c0, int m0, void* out_bytes, void* stream);

/* Debug: D[128,64] = X[shift .. shift+128, :] * W^T with X (rows x swizzle/2, bf16) staged once by TMA and the UMMA
 * A descriptor advanced by `shift` swizzle rows (base_off_mode 1 also sets the descriptor's base-offset field). */
int y3_dbg_umma_shift(y3_ctx* ctx, const void* x, int rows, const void* w, int swizzle, int shift, int base_off_mode,
                      float* out, void* stream);

/* Profiling: while dev_u64_buffer (device memory, 32 x grid-size uint64) is non-null, every CTA of the CTA-pair conv
 * kernel records %globaltimer at 12 points of its life (entry, prologue done, predecessor complete, last load issued,
 * first operands landed, last MMA issued, first accumulator complete, epilogue hand-off / stores complete per group,
 * exit).  Pass NULL to switch it off. */
int y3_dbg_timestamps(void* dev_u64_buffer);
/* The same for every launch of y3_net_forward*: launch k writes at uint64 offset k * 32 * 160 of the buffer
 * (y3_net_num_steps x 32 x 160 uint64).  Profiling build only (the release kernels carry no stamps). */
int y3_dbg_timestamps_net(void* dev_u64_buffer);
/* Measurement aid: 0 makes later y3_net_forward* calls launch every layer separately again (the persistent multi-layer
 * launches of csrc/conv_chain.cuh are the default), 1 restores the default.  Results are bit-identical either way. */
int y3_dbg_set_chain_runs(int on);

/* last device-side watchdog code (0 = none) */
int y3_watchdog_code(y3_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
