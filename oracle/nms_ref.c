/* ORACLE (test infrastructure, never shipped): greedy class-agnostic NMS, CPU restatement.
 *
 * Restates tf.image.non_max_suppression(boxes, scores, max_output_size) with its defaults
 * (iou_threshold = 0.5, no score threshold) as the reference calls it:
 *   /root/reference/inference_epistemic.py:101-102, inference_aleatoric.py:107-108,
 *   inference_standard_yolov3.py:107-108, lib_yolo/utils.py:38-39.
 * TensorFlow is an un-vendored dependency (version unpinned, 1.x); this follows the published algorithm of
 * tensorflow/core/kernels/non_max_suppression_op.cc:
 *   candidates ordered by score descending (ties: lower index first), pop next, reject iff
 *   IoU(next, s) > thr for any already selected s, stop at max_out or exhaustion.
 * IoU in fp32: corners re-ordered with min/max, area = (y2-y1)*(x2-x1), 0 if either area <= 0,
 * inter / (a_i + a_j - inter).   Compile with -ffp-contract=off so no FMA changes a rounding.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float score; int32_t idx; } cand_t;

static int cmp_cand(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

static float fminf_(float a, float b) { return a < b ? a : b; }
static float fmaxf_(float a, float b) { return a > b ? a : b; }

float byolo_oracle_iou(const float* bi, const float* bj) {
    const float ymin_i = fminf_(bi[0], bi[2]), xmin_i = fminf_(bi[1], bi[3]);
    const float ymax_i = fmaxf_(bi[0], bi[2]), xmax_i = fmaxf_(bi[1], bi[3]);
    const float ymin_j = fminf_(bj[0], bj[2]), xmin_j = fminf_(bj[1], bj[3]);
    const float ymax_j = fmaxf_(bj[0], bj[2]), xmax_j = fmaxf_(bj[1], bj[3]);
    const float area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i);
    const float area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j);
    if (area_i <= 0.0f || area_j <= 0.0f) return 0.0f;
    const float iymin = fmaxf_(ymin_i, ymin_j), ixmin = fmaxf_(xmin_i, xmin_j);
    const float iymax = fminf_(ymax_i, ymax_j), ixmax = fminf_(xmax_i, xmax_j);
    const float inter = fmaxf_(iymax - iymin, 0.0f) * fmaxf_(ixmax - ixmin, 0.0f);
    return inter / (area_i + area_j - inter);
}

/* rows: [n, d] fp32 row-major, box = columns 0..3 ([y0,x0,y1,x1]), score = column obj_idx.
 * out_idx: [max_out] selected indices in selection order. Returns the number selected. */
int byolo_oracle_nms(const float* rows, int n, int d, int obj_idx, float iou_thr, int max_out, int32_t* out_idx) {
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) { c[i].score = rows[(size_t)i * d + obj_idx]; c[i].idx = i; }
    qsort(c, (size_t)n, sizeof(cand_t), cmp_cand);
    int cnt = 0;
    for (int k = 0; k < n && cnt < max_out; ++k) {
        const float* b = rows + (size_t)c[k].idx * d;
        int keep = 1;
        for (int j = cnt - 1; j >= 0; --j) {           /* most recently selected first, as TF does */
            if (byolo_oracle_iou(b, rows + (size_t)out_idx[j] * d) > iou_thr) { keep = 0; break; }
        }
        if (keep) out_idx[cnt++] = c[k].idx;
    }
    free(c);
    return cnt;
}
