/* libbyolo C ABI - the drop-in boundary of the B200 detection hot path.
 *
 * The reference (flkraus/bayesian-yolov3) is pure Python on TensorFlow 1.x: it has no FFI of its own.  Its hot path
 * is reached through `sess.run` on a graph built by lib_yolo/yolov3.py; these entry points are what a ctypes binding
 * calls INSTEAD of that sess.run (see INTEGRATION.md for the binding).  Each function cites the reference code whose
 * execution it replaces (paths relative to /root/reference).
 *
 * Conventions: every call returns 0 on success, <0 on error (-1 bad argument, -2 CUDA runtime error, -3 driver/TMA
 * error); byolo_last_error() returns the message of the calling thread's last failure.  No exceptions cross the
 * boundary.  `stream` is a cudaStream_t passed as void*; calls only enqueue work on it (no host sync) unless stated.
 * "dev" pointers are CUDA device memory owned by the caller (e.g. torch tensors' data_ptr()); the handle owns
 * weights, workspace and cached activations.  One handle = one thread at a time; handles are independent.
 */
#ifndef BYOLO_H_
#define BYOLO_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct byolo_engine* byolo_handle;

enum { BYOLO_STANDARD = 0, BYOLO_ALEATORIC = 1, BYOLO_EPISTEMIC = 2 };   /* yolov3.py:176, :313, :454 */
enum {
    BYOLO_PREC_FP32 = 0,        /* CUDA-core fp32 convs, fp32 activations: the exact reference-precision path      */
    BYOLO_PREC_FP16_SIMT = 1,   /* CUDA-core convs on fp16-stored activations (debug twin of the tensor-core path) */
    BYOLO_PREC_FP16 = 2,        /* tcgen05 tensor-core convs, fp16 operands, fp32 accumulate (fastest path)        */
    BYOLO_PREC_FP16X3 = 3       /* tcgen05 split-fp16: every operand a hi+lo fp16 pair, hi*hi + hi*lo + lo*hi into one
                                   fp32 accumulator (3 MMAs per K step) - fp32-grade results, meets the 1e-3 parity
                                   bound element-wise (tests/test_gpu_fullsize.py)                                 */
};

typedef struct byolo_config {
    int32_t variant;            /* BYOLO_*; which model class of lib_yolo/yolov3.py                              */
    int32_t height, width;      /* config['full_img_size'][:2]; multiples of 32 (yolov3.py:207-211)              */
    int32_t cls_cnt;            /* config['cls_cnt']                                                             */
    int32_t max_batch;          /* largest B a forward call may pass                                             */
    int32_t T;                  /* config['T'] MC samples per image (epistemic only, else 1)                     */
    int32_t precision;          /* BYOLO_PREC_*                                                                  */
    int32_t standard_test_dropout; /* config['standard_test_dropout']: true = no dropout at all (layers.py:567)  */
    float drop_prob;            /* 0.1 in the reference (yolov3.py:462)                                          */
    float prior_h[9], prior_w[9]; /* config['priors'] as fractions of the image, stride 32 first (yolov3.py:29-61) */
} byolo_config;

/* Returns the ABI version (2: byolo_conv_layer takes t1 / t2, BYOLO_PREC_FP16X3, count row in gathered detections). */
int byolo_version(void);
const char* byolo_last_error(void);

/* Replaces model construction: yolov3.*.__init__ + init_model (yolov3.py:176-230, 455-516) and the prior rescale
 * of model.img_size_and_priors_if_crop (model.py:6-17) is the caller's job. */
int byolo_create(const byolo_config* cfg, byolo_handle* out);
int byolo_destroy(byolo_handle h);

/* Replaces tf.train.Saver().restore / load_darknet53_weights (detect.py:96-107, darknet.py:42-122).
 * `blob` is host memory in the BYW1 format (byolo/weights.py): per conv beta,gamma,mean,var + HWIO kernel, or
 * bias + kernel.  Folds BN (eps 1e-5, layers.py:511) into the weights and uploads them.  Synchronises the device. */
int byolo_load_weights(byolo_handle h, const void* blob, size_t bytes);

/* Geometry of the detection output: rows per image N (22743 at 608x608) and row width D (7 / 16 / 23 at cls_cnt 2);
 * obj_idx / cls_start_idx are Model.obj_idx / Model.cls_start_idx (yolov3.py:183-184, 321-322, 464-465). */
int byolo_output_shape(byolo_handle h, int32_t* N, int32_t* D, int32_t* obj_idx, int32_t* cls_start_idx);

/* Replaces one sess.run of [det_layer.bbox ...] + concat_bbox (inference_epistemic.py:173-184): backbone once,
 * head x T with Philox dropout masks (stream spec: oracle/philox.py), decode.  img_dev: [B,H,W,3] fp32 in [0,1).
 * rows_dev: [B,N,D] fp32.  image_index0 = global index of image 0 (keeps masks identical when images are sharded
 * over ranks).  seed selects the dropout stream. */
int byolo_forward(byolo_handle h, const float* img_dev, int32_t B, uint64_t seed, int32_t image_index0,
                  float* rows_dev, void* stream);

/* Replaces tf.image.non_max_suppression(boxes[:, :4], boxes[:, obj_idx], max_out) + tf.gather
 * (inference_epistemic.py:101-102; batched: inference_aleatoric.py:104-145).  out_rows_dev [B,max_out,D] in selection
 * order, zero padded; out_idx_dev [B,max_out] (-1 padded, may be NULL); out_count_dev [B]. Needs no handle. */
int byolo_nms(const float* rows_dev, int32_t B, int32_t N, int32_t D, int32_t obj_idx, float iou_thr, int32_t max_out,
              float* out_rows_dev, int32_t* out_idx_dev, int32_t* out_count_dev, void* stream);

/* byolo_nms with the options the tests and the multi-GPU gather need.  packed != 0: out_rows_dev is [B, max_out + 1, D] and
 * row max_out of every image holds (count, 0, ...) - detections and count leave as ONE fp32 block, the message of the
 * single all-gather of SURVEY.md 8e (out_count_dev may then be NULL).  force_cluster_size (0 = automatic | 1 | 2 | 4 | 8)
 * and force_chunked (the N > 32768 path for any N) are test hooks: results are identical for every setting. */
int byolo_nms_ex(const float* rows_dev, int32_t B, int32_t N, int32_t D, int32_t obj_idx, float iou_thr, int32_t max_out,
                 float* out_rows_dev, int32_t* out_idx_dev, int32_t* out_count_dev, int32_t packed, int32_t force_cluster_size,
                 int32_t force_chunked, void* stream);

/* Per-class NMS (the commented variant of inference_epistemic.py:104-126, "used to produce the results for the paper"): copies
 * rows_dev [B,N,D] to out_rows_dev with every row whose class `cls` score is NOT strictly greater than all other class scores
 * neutralised (score -inf, zero-area box).  byolo_nms on the copy then selects exactly what NMS over the class's subset
 * selects, followed - only if fewer than max_out of them survive - by neutral rows (score -inf), which the caller drops. */
int byolo_class_filter(const float* rows_dev, int32_t B, int32_t N, int32_t D, int32_t obj_idx, int32_t cls_start_idx, int32_t cls_cnt,
                       int32_t cls, float* out_rows_dev, void* stream);

/* forward + nms on device buffers (what Inference.nms fetches, inference_epistemic.py:50-54). rows_dev may be NULL
 * (internal scratch is used). */
int byolo_detect(byolo_handle h, const float* img_dev, int32_t B, uint64_t seed, int32_t image_index0, float iou_thr,
                 int32_t max_out, float* rows_dev, float* out_rows_dev, int32_t* out_idx_dev, int32_t* out_count_dev,
                 void* stream);

/* byolo_detect writing the packed layout of byolo_nms_ex: out_packed_dev [B, max_out + 1, D].  What each rank of the
 * image-sharded multi-GPU path hands to ncclAllGather (byolo/dist.py) - no packing kernels or host work in between. */
int byolo_detect_packed(byolo_handle h, const float* img_dev, int32_t B, uint64_t seed, int32_t image_index0, float iou_thr,
                        int32_t max_out, float* out_packed_dev, int32_t* out_idx_dev, void* stream);

/* The call a user of the reference makes (detect.py:124 `sess.run(box_op, {img_tensor: img})`): HOST buffers in and
 * out; copies the images host->device, runs byolo_detect, copies results back and synchronises `stream`.
 * img_host should be pinned for full copy bandwidth. */
int byolo_detect_host(byolo_handle h, const float* img_host, int32_t B, uint64_t seed, int32_t image_index0, float iou_thr,
                      int32_t max_out, float* out_rows_host, int32_t* out_count_host, void* stream);

/* Pipelined form of byolo_detect_host for streams of batches (the reference overlaps the JSON writer thread with the
 * next sess.run, inference_epistemic.py:78-83; here the copies overlap too): submit enqueues H2D(images) -> detect ->
 * D2H(results) for slot 0|1 and returns at once; wait blocks until that slot's results are in the host buffers passed
 * to submit.  Alternate the slots: the H2D of batch i+1 then overlaps the compute of batch i.  Pinned host memory. */
int byolo_submit_host(byolo_handle h, const float* img_host, int32_t B, uint64_t seed, int32_t image_index0, float iou_thr,
                      int32_t max_out, float* out_rows_host, int32_t* out_count_host, int32_t slot, void* stream);
int byolo_wait_host(byolo_handle h, int32_t slot);

/* Head decode alone on raw head outputs (layers.py:191-502 + concat): raw{0,1,2}_dev dense fp32 [B*T, g, g, ch],
 * stride 32/16/8.  Test hook and the `DetLayer.raw_output -> bbox` step of model.py:107-185. */
int byolo_decode(byolo_handle h, const float* raw0_dev, const float* raw1_dev, const float* raw2_dev, int32_t B,
                 float* rows_dev, void* stream);

/* Per-layer test hook: one conv (+dropout)+BN+leaky(+residual) (layers.py:545-575, 505-507) on dense fp32 NHWC device
 * arrays, run through the chosen precision path.  in2 (channel concat partner, 1x1 only) and residual may be NULL.
 * kernel HWIO [k,k,cin1+cin2,cout]; bn = {beta,gamma,mean,var}[cout] or NULL with bias[cout] (linear, no leaky).
 * upsample: store with the nearest x2 of layers.py:578-580 ([S,2H,2W,cout]).  dropout_layer < 0: no dropout.
 * cin1 == 3 selects the stem kernels (darknet.py:10: 3x3, stride 1, 32 filters, BN; H, W multiples of 32).
 * t1 / t2 > 1 (1x1 convs only): stack_feature_map without the copy (layers.py:595-597) - in1 (t1, no in2) or in2 (t2) holds
 * S / t samples and sample s of the conv reads sample s / t of it; pass 1 otherwise. */
int byolo_conv_layer(int32_t precision, const float* in1_dev, const float* in2_dev, int32_t S, int32_t H, int32_t W,
                     int32_t cin1, int32_t cin2, int32_t k, int32_t stride, int32_t cout, const float* kernel_host,
                     const float* bn_host, const float* bias_host, const float* residual_dev, int32_t upsample,
                     int32_t dropout_layer, int32_t T, uint64_t seed, int32_t image_index0, float drop_prob,
                     int32_t t1, int32_t t2, float* out_dev, void* stream);

/* Debug read-back of ModelBuilder layer outputs (model.py:40-41 `layers` list semantics): conv index 0..74 in weight
 * order; writes dense fp32 [S,H,W,C] to dst_dev (capacity in floats) and reports the shape.  Valid after a forward. */
int byolo_get_activation(byolo_handle h, int32_t conv_index, float* dst_dev, size_t capacity, int32_t shape[4], void* stream);

/* Per-launch device timing of byolo_detect (CUDA events on the launch stream).  byolo_profile(h, 1) makes every
 * following byolo_detect record an event before each launch (mode 1); byolo_profile_read returns, for the most recent one, one
 * entry per launch in order: duration [ms], kind (0 stem, 1 conv, 3 decode, 4 nms; 2 is unused), conv index (or -1)
 * and that launch's algorithmic FLOPs (2*MAC), plus for conv launches the effective SM clock in MHz during the launch
 * (clock64 / globaltimer read by CTA 0; 0 elsewhere).  Returns the number of entries.  Waits for the last event. */
int byolo_profile(byolo_handle h, int32_t enable);
/* byolo_profile(h, 2): coarse mode - four events per byolo_detect (start, after the stem launch, before the decode
 * launch, end) kept in a ring of 256 calls.  Nothing is recorded between the 74 conv launches, so they overlap
 * (programmatic dependent launch) exactly as in an unprofiled run.  Reads the most recent calls, oldest first: duration
 * [ms] of the stem launch, of the whole conv stack, and of decode + NMS.  Returns the number of entries. */
int byolo_profile_read_coarse(byolo_handle h, float* stem_ms, float* conv_ms, float* tail_ms, int32_t capacity);
int byolo_profile_read(byolo_handle h, float* ms, int32_t* kind, int32_t* layer, double* flops, float* sm_mhz, int32_t capacity);

/* Number of kernels one byolo_detect launches for batch B (bench.py reports it as gpu_launches). */
int byolo_launch_count(byolo_handle h, int32_t B);

/* Algorithmic FLOPs (2*MAC over all 75 convs, backbone once + head x T) of one image, SURVEY.md 8d. */
double byolo_flops_per_image(byolo_handle h);
/* FLOPs the tensor cores actually execute per image: conv "75" of the Bayesian head reads the MC-stacked backbone map,
 * identical for the T samples (yolov3.py:538-544), so its GEMM runs once per image (the T dropout masks are applied in the
 * epilogue); BYOLO_PREC_FP16X3 executes three MMA passes per product. */
double byolo_flops_per_image_executed(byolo_handle h);

#ifdef __cplusplus
}
#endif
#endif /* BYOLO_H_ */
