/* p24.h — C ABI of libp24_b200: hand-written sm_100a kernels for the YOLOX-24p loss, SimOTA
 * assignment and inference postprocess hot path.
 *
 * The reference (IN2-ViAUn/Exploration-of-Potential) has no FFI for this path: its boundary is
 * plain Python attribute lookup on torch eager code.  Each entry point below
 * names the reference function it replaces (paths relative to /root/reference/yolox_24p); the
 * Python host side in exploration-of-potential_b200/p24/ keeps the reference signatures and
 * calls these through ctypes (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with `h_`;
 *   - all floats are IEEE fp32; strides are in ELEMENTS (floats), so strided views of the head's
 *     [B, A, 27+nc] buffer are passed without copies (losses.py:185-187);
 *   - `stream` is a cudaStream_t passed as void*; functions only enqueue work: they never
 *     allocate, never synchronise and may be called from different host threads on different
 *     streams with disjoint buffers;
 *   - return value: 0 on success, a positive cudaError_t from the launch, or a negative
 *     P24_E_* argument error.  p24_error_string() explains either.
 */
#ifndef P24_H_
#define P24_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P24_ABI_VERSION 4
#define P24_RAYS 24
#define P24_TOPK 10        /* n_candidate_k cap, losses.py:452 */

#define P24_E_BADARG (-1)
#define P24_E_WORKSPACE (-2)
#define P24_E_UNSUPPORTED (-3)

#define P24_MAX_RANKS 16    /* GPUs of one box that can share the fused all-reduce */

/* p24_simota_loss_batch flags */
#define P24_F_NO_PRUNE 1u      /* evaluate every polygon angle sum exactly (self-check of the pruning) */
#define P24_F_NO_FILTER 2u     /* evaluate every pair value exactly (self-check of the top-k filter) */
#define P24_F_NO_PDL 8u        /* plain stream-ordered launches instead of programmatic dependent launch */
#define P24_F_ALL_ROWS 4u      /* every label row is a GT (per-image API: the caller passes num_gt rows) */
#define P24_F_EARLY_PREP 16u   /* the caller guarantees that `labels` and the head output of THIS call were complete before
                                  the previous call on this workspace and stream was enqueued (back-to-back loss steps on
                                  resident inputs): k_prep then prepares the step beside the previous step's last kernel and
                                  waits for it only before it exits.  Without the flag k_prep waits first (always safe). */

int p24_abi_version(void);
const char* p24_error_string(int code);

/* Bytes of scratch p24_assign_batch / p24_loss_sums need for a batch (B images, A anchors,
 * label rows Lmax).  The caller allocates it once (device memory, 256-byte aligned). */
size_t p24_workspace_bytes(int B, int A, int Lmax);
/* Clears a fresh workspace (cudaMemsetAsync on `stream`); needed once per allocation, or after a failed call. */
int p24_workspace_init(void* workspace, size_t workspace_bytes, void* stream);

/* The fused training hot path: Loss_Function.get_assignments + get_in_boxes_info + pts_in_poly +
 * cost + dynamic_k_matching (models/losses.py:359-592), utils.boxes.bboxes_iou (utils/boxes.py:166-243)
 * and the loss sums of Loss_Function.forward (models/losses.py:246-302) for a whole batch in one call;
 * the G x A cost matrix is never written to memory.
 *   outputs  [B, A, 27+nc] decoded head output (strides img_stride / row_stride, channel stride 1)
 *   labels   [B, Lmax, 51]  (cls, cx, cy, 24 x (x, y)); valid rows first, zero padded
 *   x_shifts, y_shifts, strides: [A] (the head's 3-lists concatenated, losses.py:193-195)
 *   h_levels [n_levels][4] HOST int32 (anchor offset, grid width, grid height, the bits of the level's fp32 stride): the
 *            level structure of the grid as
 *            the head builds it (yolo_head_24p.py:222-230): anchors [offset, offset + W*H) of a level are its W x H cells,
 *            row-major, x_shifts = column, y_shifts = row, one stride value.  The levels must tile [0, A) in order
 *            (n_levels <= 4).  The centre-window anchors of a GT are enumerated from it instead of being searched.
 * writes
 *   fg_mask    [B, A] uint8   final foreground mask (losses.py:486)
 *   matched_gt [B, A] int32   GT row matched to the anchor, -1 for background (losses.py:488)
 *   pred_iou   [B, A] fp32    pair value of the match, 0 for background (losses.py:491)
 *   num_fg     [B]    int32   foreground anchors per image (losses.py:481)
 *   num_gt     [B]    int32   nlabel (losses.py:190)
 *   dyn_k      [B, Lmax] int32 dynamic k per GT, 0 beyond num_gt (losses.py:456)
 *   sums28     [28] fp32 (may be NULL: assignment only):
 *              sums[0..23] = sum_f loss24[f, k]            (IOUloss.forward on the matched pairs, losses.py:283)
 *              sums[24]    = sum BCEWithLogits(obj, fg)     (losses.py:294)
 *              sums[25]    = sum BCEWithLogits(cls[fg], onehot * pred_iou)   (losses.py:298)
 *              sums[26]    = sum_b num_fg, sums[27] = sum_b num_gt
 *              This 28-float vector is what is all-reduced across GPUs.
 *   state26 / result54 / weights_n27 (all three or none): when given (single GPU), the last CTA also applies
 *              p24_loss_finalize (one launch less); with several GPUs they are ignored here: p24_comm_finish (fused peer
 *              exchange) or an all-reduce of sums28 followed by p24_loss_finalize apply them.
 * The workspace (p24_workspace_bytes, 256-byte aligned) must have been cleared once with p24_workspace_init;
 * a successful call leaves it ready for the next one.  One workspace serves one call at a time.
 */
int p24_simota_loss_batch(const float* outputs, int64_t img_stride, int64_t row_stride, int B, int A, int num_classes,
                          const float* labels, int64_t lab_img_stride, int64_t lab_row_stride, int Lmax,
                          const float* x_shifts, const float* y_shifts, const float* strides,
                          const int32_t* h_levels, int n_levels,
                          uint8_t* fg_mask, int32_t* matched_gt, float* pred_iou,
                          int32_t* num_fg, int32_t* num_gt, int32_t* dyn_k, float* sums28,
                          float* state26, float* result54, float* weights_n27,
                          void* workspace, size_t workspace_bytes, uint32_t flags,
                          void* const* h_mailboxes, int rank, int nranks, uint32_t epoch, void* stream);

/* The same path on the head's RAW conv outputs (head decode fused into the kernels).  Replaces
 * YOLOXHead.get_output_and_grid + the cat / view / permute / reshape copies of YOLOXHead.forward(train=True)
 * (models/yolo_head_24p.py:167-197, 212-237): the kernels read the per-level NCHW tensors
 *   reg [B, 26, H, W]   obj [B, 1, H, W]   cls [B, nc, H, W]      (yolo_head_24p.py:160-164; H x W planes dense)
 * and decode on load -- centre (v + grid) * stride, radii exp(v) * stride (yolo_head_24p.py:233-235) -- so the decoded
 * [B, A, 27 + nc] buffer is never written.
 *   h_raw               HOST array of 3 * n_levels device pointers: reg of level 0 .. n_levels-1, then obj, then cls
 *   h_raw_batch_stride  HOST array of 3 * n_levels batch strides (elements); channel stride = H * W
 * Grids, labels, outputs and everything else as in p24_simota_loss_batch; 27 + nc <= 108. */
int p24_simota_loss_batch_raw(const float* const* h_raw, const int64_t* h_raw_batch_stride, int B, int A, int num_classes,
                              const float* labels, int64_t lab_img_stride, int64_t lab_row_stride, int Lmax,
                              const float* x_shifts, const float* y_shifts, const float* strides,
                              const int32_t* h_levels, int n_levels,
                              uint8_t* fg_mask, int32_t* matched_gt, float* pred_iou,
                              int32_t* num_fg, int32_t* num_gt, int32_t* dyn_k, float* sums28,
                              float* state26, float* result54, float* weights_n27,
                              void* workspace, size_t workspace_bytes, uint32_t flags,
                              void* const* h_mailboxes, int rank, int nranks, uint32_t epoch, void* stream);

/* Fused all-reduce of the 28 loss sums over NVLink / NVSwitch peer memory (the one collective of the path).
 * Every rank (one process per GPU) owns a small mailbox allocated with p24_comm_alloc and maps its peers' mailboxes
 * through CUDA IPC (p24_comm_export on the owner, p24_comm_import on the peers).  p24_simota_loss_batch then takes
 *   h_mailboxes  HOST array of nranks device pointers, entry r = the mailbox of rank r as mapped in this process
 *   rank, nranks this process's rank and the number of ranks (<= P24_MAX_RANKS); nranks <= 1 or h_mailboxes == NULL:
 *                no exchange
 *   epoch        call counter, the same on every rank, starting at 1 and incremented by the caller for every call
 * and the last CTA of the chain stores its 28 sums into every peer's mailbox (P2P stores + a flag with the epoch): the
 * PUBLISH half of the exchange, fused into the compute kernel.  The COLLECT half is p24_comm_finish: one warp waits for the
 * flags of all peers in the rank's own mailbox, adds the contributions in rank order (bit-identical on all ranks), writes
 * sums28 and applies the normalisation / re-weighting (state26 / result54 / weights_n27 as in p24_loss_finalize): no NCCL
 * launch.  Enqueue it on ANY stream of the device after the call (it waits for the chain's publication through a flag
 * in the mailbox, not through the stream): on a side stream the next step does not wait for the peers -- nothing of
 * the next step's chain depends on the global sums.  Four slot sets alternate with the epoch; the chain itself holds
 * back the publication of epoch e until this rank's p24_comm_finish of epoch e - 2 has completed, so the calls of
 * p24_comm_finish must be enqueued in epoch order and none may be skipped.
 * All ranks must make the same sequence of calls.  `workspace` (+ B, A, Lmax) is optional: the wait is recorded in its
 * status word 4. */
size_t p24_comm_mailbox_bytes(void);
int p24_comm_alloc(void** d_mailbox);
int p24_comm_free(void* d_mailbox);
int p24_comm_export(void* d_mailbox, void* h_handle64);
int p24_comm_import(const void* h_handle64, void** d_peer_mailbox);
int p24_comm_close(void* d_peer_mailbox);
int p24_comm_finish(void* d_own_mailbox, int nranks, uint32_t epoch, float* sums28, float* state26, float* result54,
                    float* weights_n27, void* workspace, int B, int A, int Lmax, void* stream);

/* Normalisation and stateful re-weighting (models/losses.py:280-345).
 * state[26] = last_iou_loss[24], last_obj_loss, last_cls_loss (initially 1.0), updated in place.
 * result[54] = loss, reg_w*loss_iou[24], loss_obj, loss_cls, num_fg/max(num_gts,1),
 *              reg_w[24], obj_w, cls_w;  weights_n[27] = reg_w[24], obj_w, cls_w, max(num_fg,1)
 *              (what the backward needs). */
int p24_loss_finalize(const float* sums28, float* state26, float* result54, float* weights_n27, void* stream);

/* IOUloss.circle_inter (models/losses.py:23-78), element-wise over n pairs.
 * gt_r / pd_r: [n, 24] with row strides; outputs res_inter, dist: dense [n, 24]. */
int p24_circle_inter_fwd(const float* gt_cx, const float* gt_cy, const float* gt_r, int64_t gt_r_stride,
                         const float* pd_cx, const float* pd_cy, const float* pd_r, int64_t pd_r_stride,
                         int n, float* res_inter, float* dist, void* stream);

/* IOUloss.forward (models/losses.py:80-157): pred [n,26], target [n,50] -> loss24 [n,24] dense. */
int p24_iou_loss_fwd(const float* pred, int64_t pred_stride, const float* target, int64_t target_stride,
                     int n, float* loss24, void* stream);

/* Backward of IOUloss.forward w.r.t. pred (autograd of losses.py:80-157; clipped acos and branch
 * masks have zero gradient).  grad_loss24 dense [n,24]; grad_pred dense [n,26] (overwritten). */
int p24_iou_loss_bwd(const float* pred, int64_t pred_stride, const float* target, int64_t target_stride,
                     const float* grad_loss24, int n, float* grad_pred, void* stream);

/* utils.boxes.bboxes_iou (utils/boxes.py:166-243): gt [G,50], pred [P,26] -> out dense [G,P]
 * (the mean ray LOSS / 2 that SimOTA consumes as an "iou", SURVEY.md note 3). */
int p24_pair_iou(const float* gt50, int64_t gt_stride, int G, const float* pred26, int64_t pred_stride, int P,
                 float* out, void* stream);

/* Backward of the whole loss w.r.t. outputs (autograd of models/losses.py:283-341; the weights are constants,
 * losses.py:312-314): grad_outputs [B, A, 27+nc] dense, fully overwritten.  weights_n27 from p24_loss_finalize /
 * p24_simota_loss_batch; grad_scale: device scalar (upstream gradient of the loss) or NULL for 1. */
int p24_loss_bwd(const float* outputs, int64_t img_stride, int64_t row_stride, int B, int A, int num_classes,
                 const float* labels, int64_t lab_img_stride, int64_t lab_row_stride,
                 const uint8_t* fg_mask, const int32_t* matched_gt, const float* pred_iou,
                 const float* weights_n27, const float* grad_scale, float* grad_outputs, void* stream);

/* The same backward w.r.t. the head's RAW per-level conv outputs (p24_simota_loss_batch_raw; the decode of
 * models/yolo_head_24p.py:233-235 folded in).  h_raw / h_raw_batch_stride as for p24_simota_loss_batch_raw;
 * h_grad_raw: HOST array of 3 * n_levels DEVICE pointers in the same order, each a dense [B, C, H, W] tensor
 * (C = 26 / 1 / nc), fully overwritten; h_levels: HOST (anchor offset, W, H, stride bits) per level. */
int p24_loss_bwd_raw(const float* const* h_raw, const int64_t* h_raw_batch_stride, float* const* h_grad_raw,
                     const int32_t* h_levels, int n_levels, int B, int A, int num_classes,
                     const float* labels, int64_t lab_img_stride, int64_t lab_row_stride,
                     const uint8_t* fg_mask, const int32_t* matched_gt, const float* pred_iou,
                     const float* weights_n27, const float* grad_scale, void* stream);

/* Label packing of the dataset transform for a whole batch (TrainTransform.__call__, datasets/data_augment.py:131-174;
 * the wire format COCO24PDataset.__getitem__ hands to the loss, datasets/coco24p.py:107-131).
 *   targets    DEVICE float64 [sum_b n_b, 51]: the images' label-file rows [cls, cx, cy, 24 x (x, y)], normalised to
 *              [0, 1] (np.loadtxt), image after image (may be NULL when the batch has no target at all)
 *   offsets    DEVICE int32 [B + 1]: first row of image b (offsets[B] = total)
 *   shapes_hw  DEVICE int32 [B, 2]: (height, width) of the RESIZED, unpadded image the transform receives
 *   in_h, in_w the padded network input (input_dim)
 *   labels     DEVICE fp32 [B, max_labels, 51], fully overwritten: pixel coordinates x * width * r, y * height * r with
 *              r = min(in_h / height, in_w / width), computed in float64 and cast once like the reference; rows past
 *              the image's targets are zero, targets past max_labels are dropped
 *   nlabel     DEVICE int32 [B] or NULL: rows with a positive sum (models/losses.py:190) */
int p24_pack_labels(const double* targets, const int32_t* offsets, const int32_t* shapes_hw, int in_h, int in_w,
                    int B, int max_labels, float* labels, int32_t* nlabel, void* stream);

/* Loss_Function.dynamic_k_matching (models/losses.py:444-494) on a materialised cost matrix.
 * cost, ious: dense [G, P];  fg_in [P] uint8, matched [P] int32 (-1 when not fg), matched_iou [P] fp32,
 * dyn_k [G] int32, num_fg [1] int32.  torch.topk leaves tie order unspecified; ties go to the lower index here. */
size_t p24_dynamic_k_workspace_bytes(int G, int P);
int p24_dynamic_k_matching(const float* cost, const float* ious, int G, int P,
                           uint8_t* fg_in, int32_t* matched, float* matched_iou, int32_t* dyn_k, int32_t* num_fg,
                           void* workspace, size_t workspace_bytes, void* stream);

/* utils.boxes.postprocess (utils/boxes.py:29-99; twin show_24p.py:212-264), per image.
 *   prediction [B, A, 27+nc] (obj / cls already sigmoid)
 *   h_coef_x, h_coef_y [24] HOST arrays: theta_k*cos(theta_k), theta_k*sin(theta_k) as the reference's own torch
 *     expression evaluates them (boxes.py:30-33), passed in so libm differences cannot enter
 *   class_agnostic != 0 -> torchvision nms, else batched_nms with the coordinate trick
 * writes, per image b (capacity A rows each):
 *   cand_count [B]          rows that passed the score filter
 *   det_count  [B]          rows kept after NMS
 *   det_rows   [B, A, 29]   cx, cy, r0..r23, obj, class_conf, class_pred in NMS (score) order
 *   keep_idx   [B, A]       anchor index of every kept row, same order
 *   rect_debug [B, A, 4]    (may be NULL) rectangles of the candidates in anchor order */
size_t p24_postprocess_workspace_bytes(int B, int A);
int p24_postprocess(const float* prediction, int64_t img_stride, int64_t row_stride, int B, int A, int num_classes,
                    const float* h_coef_x, const float* h_coef_y, float conf_thre, float nms_thre, int class_agnostic,
                    int32_t* cand_count, int32_t* det_count, float* det_rows, int32_t* keep_idx, float* rect_debug,
                    void* workspace, size_t workspace_bytes, void* stream);

/* The same on the head's RAW per-level conv outputs (YOLOXHead.forward(train=False) after the prediction convs:
 * sigmoid on obj / cls, flatten + cat + permute, decode_outputs, models/yolo_head_24p.py:191, 201-211, 239-256 -- folded
 * into the filter pass; the decoded [B, A, 27+nc] prediction is never written).  h_raw / h_raw_batch_stride / h_levels as
 * for p24_simota_loss_batch_raw (reg [B,26,H,W], obj [B,1,H,W], cls [B,nc,H,W] logits per level).  det_rows carry the
 * decoded [cx, cy, r0..r23, sigmoid(obj)] like the reference's rows. */
int p24_postprocess_raw(const float* const* h_raw, const int64_t* h_raw_batch_stride, const int32_t* h_levels,
                        int n_levels, int B, int A, int num_classes, const float* h_coef_x, const float* h_coef_y,
                        float conf_thre, float nms_thre, int class_agnostic, int32_t* cand_count,
                        int32_t* det_count, float* det_rows, int32_t* keep_idx, float* rect_debug, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Status of a workspace, read back to the HOST (one small device-to-host copy + a synchronisation of `stream`):
 *   h_status8[0]  error bits (P24_ERR_*), sticky; 0 = none.  The reference raises on its failures (losses.py:81-82); the
 *                 Python host side raises P24Error when a bit is set.  (Channel and bit values are part of the ABI; the
 *                 kernels of this version have no failure left to report: the window-pair list is sized for the worst case
 *                 and the collect kernel of the fused all-reduce waits without a time-out.)
 *   counters since the previous read (they restart with every read):
 *   h_status8[1]  GTs whose dynamic k took the brute-force path
 *   h_status8[2]  GTs that spilled into the penalised regime
 *   h_status8[3]  longest top-10 candidate list seen
 *   h_status8[4]  clock cycles between the first and the last rank's contribution to the last fused all-reduce
 *                 arriving at this rank (nranks > 1)
 *   h_status8[5]  list entries seen, over h_status8[6] GTs
 *   h_status8[7]  GTs whose dynamic k needed exact pair values (the bracket of the bounds straddled an integer) */
#define P24_ERR_WINDOW_OVERFLOW 1
#define P24_ERR_PEER_TIMEOUT 4
int p24_read_status(void* workspace, int B, int A, int Lmax, int32_t* h_status8, void* stream);

/* Profiling aid (bench.py): when enabled, p24_simota_loss_batch and p24_postprocess record CUDA events around their
 * kernels on the launching stream and launch them in plain stream order; p24_profile_read waits for the last call and
 * returns the durations in milliseconds into a HOST array: h_ms8[0..2] = k_prep, k_pass, k_tail (+ k_fin),
 * h_ms8[4..5] = k_post_filter, k_post_nms; 0 for kernels that did not run since the last read.  Process-global. */
int p24_profile_enable(int on);
int p24_profile_read(float* h_ms8);

#ifdef __cplusplus
}
#endif
#endif /* P24_H_ */
