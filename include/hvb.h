/*
 * hvb.h — C ABI of libhvb.so, the B200 (sm_100a) implementation of the per-frame hot path of
 * JetJadeja/hockey-vision-analytics.
 *
 * The reference has no FFI of its own (it is pure Python gluing ultralytics / supervision /
 * OpenCV / torchvision / scikit-learn); every entry point below names the reference call site
 * (file:line under the reference tree) whose arithmetic it replaces.  The Python host layer
 * (hockey-vision-analytics_b200/hvb/) binds these with ctypes and presents the reference's own
 * call surface (InferenceSlicer callback, sv.Detections, TeamClassifier/HybridTeamClassifier);
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative hvb_status otherwise; the message is in
 *     hvb_last_error() (thread-local);
 *   - pointers named *_dev are device pointers on the context's GPU (allocated by the caller —
 *     e.g. a torch tensor's data_ptr() — or by hvb_malloc); pointers named *_host are host
 *     pointers; nothing is retained after a call returns unless stated;
 *   - all work is enqueued on the context's stream (hvb_ctx_set_stream) and is asynchronous
 *     unless the function name ends in _host or the comment says it synchronises;
 *   - a context is not re-entrant: one caller at a time per hvb_ctx (the Python layer locks).
 *   - there is NO CPU fallback anywhere in this library.
 */
#ifndef HVB_H_
#define HVB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HVB_VERSION 100

#if defined(__GNUC__)
#define HVB_API __attribute__((visibility("default")))
#else
#define HVB_API
#endif

typedef enum {
    HVB_OK = 0,
    HVB_ERR_CUDA = -1,         /* a CUDA runtime call failed (message has the CUDA error string)   */
    HVB_ERR_ARG = -2,          /* invalid argument                                                  */
    HVB_ERR_NO_DEVICE = -3,    /* no usable CUDA device / not an sm_100 part                        */
    HVB_ERR_CAPACITY = -4,     /* a fixed on-chip capacity was exceeded (e.g. NMS candidates)       */
    HVB_ERR_UNSUPPORTED = -5
} hvb_status;

typedef struct hvb_ctx hvb_ctx;

/* ---------------------------------------------------------------- context / memory plumbing */
HVB_API int hvb_version(void);
HVB_API const char* hvb_last_error(void);
HVB_API int hvb_device_count(int* out_count);                       /* never fails hard: count 0 without a GPU */
HVB_API int hvb_ctx_create(int device, hvb_ctx** out_ctx);
HVB_API int hvb_ctx_destroy(hvb_ctx* ctx);
HVB_API int hvb_ctx_set_stream(hvb_ctx* ctx, void* cuda_stream);    /* cudaStream_t; NULL = the legacy default stream */
HVB_API int hvb_ctx_use_own_stream(hvb_ctx* ctx);                   /* back to the context's private non-blocking stream (the initial state) */
HVB_API int hvb_ctx_get_stream(hvb_ctx* ctx, void** out_stream);
HVB_API int hvb_ctx_synchronize(hvb_ctx* ctx);
/* CUDA-graph support.  libhvb launches may be captured into a CUDA graph (bind the capturing stream with
 * hvb_ctx_set_stream).  Call this with on=1 BEFORE capturing: from then on a work buffer that has to grow is
 * replaced without freeing the old one, so pointers baked into captured graphs stay valid until hvb_ctx_destroy.
 * A call that would have to (re)allocate while its stream is being captured fails with HVB_ERR_UNSUPPORTED —
 * run the step once eagerly first. */
HVB_API int hvb_ctx_retain_buffers(hvb_ctx* ctx, int on);
HVB_API int hvb_ctx_sm_count(hvb_ctx* ctx, int* out_sms);
HVB_API int hvb_malloc(hvb_ctx* ctx, size_t bytes, void** out_dev);
HVB_API int hvb_free(hvb_ctx* ctx, void* ptr_dev);
HVB_API int hvb_host_alloc(hvb_ctx* ctx, size_t bytes, void** out_host);   /* pinned */
HVB_API int hvb_host_free(hvb_ctx* ctx, void* ptr_host);
HVB_API int hvb_memcpy_h2d(hvb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);  /* async */
HVB_API int hvb_memcpy_d2h(hvb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);  /* async */
/* Host-side staging: n separately allocated frames of bytes_each bytes (hockey/main.py:321 yields one array per frame)
 * -> one contiguous (pinned) buffer, copied by n_threads workers.  No CUDA call; the source of the chunk's one H2D copy. */
HVB_API int hvb_stage_frames(const void* const* src_host, int n, size_t bytes_each, void* dst_host, int n_threads);
HVB_API int hvb_memset(hvb_ctx* ctx, void* dst_dev, int value, size_t bytes);
/* Launch counter: number of libhvb kernels launched on this context since the last reset. */
HVB_API int hvb_ctx_launch_count(hvb_ctx* ctx, int reset, uint64_t* out_launches);
/* Device event timing helpers (events recorded on the context's stream). */
HVB_API int hvb_timer_start(hvb_ctx* ctx);
HVB_API int hvb_timer_stop_ms(hvb_ctx* ctx, float* out_ms);          /* synchronises on the stop event      */

/* ---------------------------------------------------------------- K1: letterbox / slicing
 * Replaces ultralytics LetterBox + BasePredictor.preprocess reached from
 * hockey/main.py:179-184 (whole frame, imgsz 1280) and, per tile, the documented
 * sv.InferenceSlicer(callback, slice_wh=(640,640), overlap 0.2) path (README.md:25, CLAUDE.md:55):
 * crop_image -> LetterBox(auto) -> cv2.resize(INTER_LINEAR) -> copyMakeBorder(114) ->
 * BGR->RGB, HWC->CHW, float32, /255.
 *
 * A plan fixes the geometry for a chunk of `n_frames` equally-sized frames; running it is ONE
 * kernel launch that writes every tile of every frame into one output buffer laid out as a
 * sequence of shape-class batches:  for class c: float32[n_frames * tiles_per_frame_c, 3, out_h_c, out_w_c]
 * starting at element offset class.out_offset.
 */
typedef struct hvb_lb_plan hvb_lb_plan;

typedef enum {
    HVB_LB_WHOLE = 0,          /* one job per frame: letterbox the whole frame to imgsz             */
    HVB_LB_SLICE_EXACT = 1,    /* InferenceSlicer tiles, each LetterBox(auto=True) -> per-shape-class batches */
    HVB_LB_SLICE_UNIFORM = 2   /* InferenceSlicer tiles, each LetterBox(auto=False) -> uniform imgsz x imgsz  */
} hvb_lb_mode;

typedef struct {
    int32_t out_h, out_w;        /* letterboxed (padded) shape of this class                        */
    int32_t tiles_per_frame;     /* how many tiles of one frame fall in this class                  */
    int32_t batch;               /* n_frames * tiles_per_frame, batch index = frame * tiles_per_frame + k */
    int64_t out_offset;          /* float offset of the class batch inside the output buffer        */
} hvb_lb_class;

typedef struct {
    int32_t frame;               /* source frame index inside the chunk                              */
    int32_t tile;                /* tile index inside the frame (slicer row-major order), 0 for WHOLE */
    int32_t cls;                 /* shape class                                                       */
    int32_t batch_index;         /* index inside the class batch                                      */
    int32_t src_x, src_y, src_w, src_h;   /* source rectangle in the frame (slicer offset + clipped size) */
    int32_t new_w, new_h;        /* resized, unpadded size                                            */
    int32_t top, left;           /* padding before the image                                          */
    int32_t out_h, out_w;
    float gain, pad_x, pad_y;    /* ultralytics scale_boxes() constants for (out_h,out_w)->(src_h,src_w) */
} hvb_lb_tile;

HVB_API int hvb_lb_plan_create(hvb_ctx* ctx, int n_frames, int frame_h, int frame_w, int mode, int imgsz,
                       int auto_pad, int stride, int slice_w, int slice_h, int overlap_w, int overlap_h,
                       hvb_lb_plan** out_plan);
HVB_API int hvb_lb_plan_destroy(hvb_lb_plan* plan);
HVB_API int hvb_lb_plan_num_classes(const hvb_lb_plan* plan, int* out_n);
HVB_API int hvb_lb_plan_get_class(const hvb_lb_plan* plan, int cls, hvb_lb_class* out_class);
HVB_API int hvb_lb_plan_num_tiles(const hvb_lb_plan* plan, int* out_n);              /* n_frames * tiles per frame */
HVB_API int hvb_lb_plan_get_tiles(const hvb_lb_plan* plan, hvb_lb_tile* out_tiles_host, int capacity);
HVB_API int hvb_lb_plan_out_floats(const hvb_lb_plan* plan, int64_t* out_floats);    /* size of the output buffer */
HVB_API int hvb_lb_plan_bytes(const hvb_lb_plan* plan, int64_t* out_read_bytes, int64_t* out_write_bytes); /* algorithmic */
/* frames_dev: uint8[n_frames, frame_h, frame_w, 3] BGR, dense.  out_dev: float32[out_floats].
 * One launch per call, grid = (blocks of a frame, n_frames): a plan holds at most 65535 frames (HVB_ERR_CAPACITY beyond). */
HVB_API int hvb_lb_plan_run(hvb_lb_plan* plan, const uint8_t* frames_dev, float* out_dev);
/* Test hook: the uint8 stage only (resize + 114 padding, still BGR/HWC), same batch layout in bytes. */
HVB_API int hvb_lb_plan_run_u8(hvb_lb_plan* plan, const uint8_t* frames_dev, uint8_t* out_dev);

/* ---------------------------------------------------------------- K2a: YOLOv8 head decode + NMS
 * Replaces ultralytics Detect._inference (DFL softmax-expectation, dist2bbox, x stride, sigmoid),
 * ops.non_max_suppression (conf gate, per-image class-aware NMS through torchvision.ops.nms with
 * the +cls*7680 offset, max_det) and ops.scale_boxes/clip_boxes, reached from
 * hockey/main.py:179-186, plus sv.move_detections for sliced tiles.
 *
 * level_dev[i]: float32 raw Detect output of level i (stride 8,16,32): channel c of anchor a of
 * image b at  b*batch_stride[i] + c*chan_stride[i] + a*anchor_stride[i]  (NCHW: chan_stride=H*W,
 * anchor_stride=1).  Channels: 64 box bins (side-major, 16 bins each: l,t,r,b) then nc class logits.
 * meta_dev: one hvb_img_meta per image.  Outputs hold max_det rows per slot (row block = meta.out_slot),
 * score order; out_count[slot] = kept count, or -1 if the image had more than 1024 candidates.
 */
typedef struct {
    float gain, pad_x, pad_y;    /* scale_boxes: (x - pad) / gain                                    */
    float clip_w, clip_h;        /* clip x to [0,clip_w], y to [0,clip_h] (original tile/frame shape) */
    float off_x, off_y;          /* slice offset (sv.move_detections), applied by hvb_gather_tiles, not here   */
    int32_t out_slot;            /* row of the output arrays this image writes (e.g. frame * tiles + tile)     */
} hvb_img_meta;                  /* 32 bytes */

HVB_API int hvb_decode_nms(hvb_ctx* ctx, const float* const level_dev[3], const int32_t level_h[3],
                   const int32_t level_w[3], const int64_t batch_stride[3], const int64_t chan_stride[3],
                   const int64_t anchor_stride[3], int batch, int nc, float conf_thres, float iou_thres,
                   int max_det, int agnostic, const hvb_img_meta* meta_dev,
                   float* out_xyxy_dev /*[batch,max_det,4]*/, float* out_conf_dev /*[batch,max_det]*/,
                   int32_t* out_cls_dev /*[batch,max_det]*/, int32_t* out_count_dev /*[batch]*/);
/* Images whose out_count came back -1 had more than 1024 candidates above conf_thres; re-run just
 * those (images_dev: int32[n_images] batch indices) with the 8192-candidate tier.  Still -1 after
 * that means the on-chip capacity (hvb_nms_capacity) is exceeded. */
HVB_API int hvb_decode_nms_large(hvb_ctx* ctx, const float* const level_dev[3], const int32_t level_h[3],
                         const int32_t level_w[3], const int64_t batch_stride[3], const int64_t chan_stride[3],
                         const int64_t anchor_stride[3], const int32_t* images_dev, int n_images, int nc,
                         float conf_thres, float iou_thres, int max_det, int agnostic,
                         const hvb_img_meta* meta_dev, float* out_xyxy_dev, float* out_conf_dev,
                         int32_t* out_cls_dev, int32_t* out_count_dev);
/* The same with the class logits in their own tensors (box_level_dev: the 64 box-bin channels; cls_level_dev: the nc
 * class channels, channel c of anchor a of image b at  b*cls_batch_stride + c*cls_chan_stride + a*cls_anchor_stride).
 * A dense [B, H, W, nc] class tensor makes the confidence scan — the only pass that touches every anchor — read
 * contiguous memory (a channels-last [B, H, W, 64+nc] head costs a 32-byte sector per anchor for nc*4 useful bytes).
 * images_dev: NULL = all `batch` images with the 1024-candidate tier; else int32[batch] image indices to (re)run with
 * the 8192-candidate tier (what hvb_decode_nms_large does for the combined layout). */
HVB_API int hvb_decode_nms_split(hvb_ctx* ctx, const float* const box_level_dev[3], const float* const cls_level_dev[3],
                         const int32_t level_h[3], const int32_t level_w[3], const int64_t box_batch_stride[3],
                         const int64_t box_chan_stride[3], const int64_t box_anchor_stride[3],
                         const int64_t cls_batch_stride[3], const int64_t cls_chan_stride[3],
                         const int64_t cls_anchor_stride[3], const int32_t* images_dev, int batch, int nc,
                         float conf_thres, float iou_thres, int max_det, int agnostic, const hvb_img_meta* meta_dev,
                         float* out_xyxy_dev, float* out_conf_dev, int32_t* out_cls_dev, int32_t* out_count_dev);
/* Test hooks: decode only (float32[batch, 4+nc, A] like Detect._inference) and NMS only on
 * caller-provided candidates (boxes xyxy float32[n,4], scores, classes; one image). */
HVB_API int hvb_decode_only(hvb_ctx* ctx, const float* const level_dev[3], const int32_t level_h[3],
                    const int32_t level_w[3], const int64_t batch_stride[3], const int64_t chan_stride[3],
                    const int64_t anchor_stride[3], int batch, int nc, float* out_pred_dev);
HVB_API int hvb_nms_f32(hvb_ctx* ctx, const float* boxes_dev, const float* scores_dev, const int32_t* cls_dev,
                int n, float iou_thres, int max_det, int agnostic,
                int32_t* out_keep_idx_dev /*[max_det]*/, int32_t* out_count_dev /*[1]*/);
HVB_API int hvb_nms_capacity(int* out_max_candidates);

/* ---------------------------------------------------------------- K2b: cross-slice merge NMS
 * Replaces sv.Detections.with_nms -> box_non_max_suppression / box_iou_batch (float64, keep mask
 * in input order, class-aware unless class_agnostic) applied to the merged slicer result.
 * Segments: detections of segment s are rows seg_offsets[s] .. seg_offsets[s+1]-1 (one frame each).
 */
HVB_API int hvb_merge_nms(hvb_ctx* ctx, const double* xyxy_dev, const float* conf_dev, const int32_t* cls_dev,
                  const int32_t* seg_offsets_dev, int n_segments, int n_total, double iou_thres,
                  int class_agnostic, uint8_t* out_keep_dev);

/* sv.move_detections + Detections.merge for a chunk of frames: compacts the per-slot K2a results
 * (slot = frame * slots_per_frame + tile; rows 0..count[slot]-1 of each slot) into per-frame merged
 * lists in slicer tile order, float64 boxes moved by the slot's offset (slot_off_xy_dev float32[n_slots,2]).
 * out_seg_offsets_dev: int32[n_frames+1]; out_slot_dev (optional) records the source slot of each row.
 * Output arrays must hold n_slots*max_det rows. */
HVB_API int hvb_gather_tiles(hvb_ctx* ctx, const float* xyxy_dev, const float* conf_dev, const int32_t* cls_dev,
                     const int32_t* count_dev, const float* slot_off_xy_dev, int n_slots, int slots_per_frame,
                     int max_det, double* out_xyxy_dev, float* out_conf_dev, int32_t* out_cls_dev,
                     int32_t* out_slot_dev, int32_t* out_seg_offsets_dev);

/* ---------------------------------------------------------------- K3: per-detection crop features
 * A crop is described by where its pixels live inside one device byte buffer (`pixels_dev`):
 * either a view into a resident frame (offset of its first pixel, pitch = frame_w*3) or a packed
 * copy of a host crop (pitch = w*3).
 */
typedef struct {
    int64_t offset;              /* byte offset of pixel (0,0) of the crop inside pixels_dev         */
    int32_t pitch;               /* bytes between rows                                                */
    int32_t h, w;                /* crop size in pixels (may be 0)                                    */
    int32_t reserved;
} hvb_crop_desc;

/* sv.crop_image (hockey/main.py:324-326): np.round(xyxy).astype(int) (half-to-even) then the numpy
 * slice frame[y0:y1, x0:x1] (negative indices wrap, ends clamp).  frame_idx_dev may be NULL (frame 0). */
HVB_API int hvb_crops_from_boxes(hvb_ctx* ctx, const float* xyxy_dev, const int32_t* frame_idx_dev, int n,
                         int frame_h, int frame_w, hvb_crop_desc* out_crops_dev);

typedef enum {
    HVB_ROI_HYBRID = 0,          /* HybridTeamClassifier.extract_jersey_region, team_hybrid.py:49-64  */
    HVB_ROI_SIMPLE = 1,          /* TeamClassifier.extract_jersey_region, team.py:76-99               */
    HVB_ROI_WHOLE = 2,
    HVB_ROI_SEGMENT = 3          /* the rectangle SegmentationTeamClassifier.segment_player falls back to
                                  * when GrabCut is unavailable (team_segmentation.py:87-96): rows
                                  * [int(0.2h), int(0.6h)), cols [int(0.3w), int(0.7w)); only
                                  * hvb_jersey_color_stats accepts it                                  */
} hvb_roi_mode;

typedef struct {
    uint32_t hist[34];           /* H 18 bins of 10, S 8 bins of 32, V 8 bins of 32 (cv2.calcHist)    */
    uint32_t counts[3];          /* S<30, S>100, (V>200)&(S<30)                                       */
    uint32_t n;                  /* ROI pixel count                                                   */
    uint32_t roi[4];             /* top, bottom, left, right actually used                            */
    uint32_t pad_[2];
    uint64_t sums[6];            /* sum of H,S,V,L,a,b                                                */
    uint64_t sumsq[6];           /* sum of squares of H,S,V,L,a,b                                     */
} hvb_color_raw;                 /* 272 bytes */

/* HybridTeamClassifier.extract_color_features, team_hybrid.py:89-142: cv2.cvtColor BGR2HSV/BGR2LAB
 * (bit-exact), calcHist x3, per-channel mean / population std (/255), S<30, S>100, white ratios.
 * out_feat_dev: float64[n,49] (layout as the reference's np.concatenate) or NULL;
 * out_raw_dev: hvb_color_raw[n] or NULL.  An empty ROI yields NaN features (the reference raises). */
HVB_API int hvb_color_features(hvb_ctx* ctx, const uint8_t* pixels_dev, const hvb_crop_desc* crops_dev, int n,
                       int roi_mode, double* out_feat_dev, int64_t feat_row_stride /*doubles, >=49*/,
                       hvb_color_raw* out_raw_dev);
/* SegmentationTeamClassifier.extract_jersey_colors (team_segmentation.py:98-148) over a rectangular
 * mask (the colour features without GrabCut).  Same bit-exact HSV / LAB pass as
 * hvb_color_features, different statistics; the four dictionary values are one division each on the
 * host (hvb/team_segmentation.py). */
typedef struct {
    uint32_t n;                  /* pixels under the mask                                             */
    uint32_t white;              /* L>200 & 0<=a-128<10 & 0<=b-128<10: the reference subtracts 128 from
                                  * UINT8 arrays, so a,b below 128 wrap and never count as white        */
    uint32_t hue_hist[18];       /* np.histogram(H, 18, (0,180)) of the NON-white pixels              */
    uint64_t sat_colored;        /* sum of S over the non-white pixels                                */
    uint64_t sat_all;            /* sum of S over all pixels                                          */
    uint64_t val_all;            /* sum of V over all pixels                                          */
    int32_t roi[4];              /* top, bottom, left, right actually used                            */
} hvb_jersey_raw;                /* 120 bytes */
HVB_API int hvb_jersey_color_stats(hvb_ctx* ctx, const uint8_t* pixels_dev, const hvb_crop_desc* crops_dev, int n,
                           int roi_mode, hvb_jersey_raw* out_raw_dev);
/* Test hook: plain per-pixel conversion of n_px BGR pixels (the 2^24 colour-cube test). */
HVB_API int hvb_cvt_hsv_lab(hvb_ctx* ctx, const uint8_t* bgr_dev, int64_t n_px, uint8_t* out_hsv_dev,
                    uint8_t* out_lab_dev);

/* self.preprocess of team_hybrid.py:31-36 applied to the jersey ROI of each crop: Pillow-exact
 * antialiased bilinear resize to 128x64 (uint8), /255, (x-mean)/std with the RGB statistics applied
 * to the BGR channels unswapped, CHW float32.  out_dev: float32[n,3,128,64].
 * out_u8_dev (optional, test hook): the resized uint8[n,128,64,3].  Empty ROI -> zeros + flag
 * (the reference's except: -> zeros(576) path); out_valid_dev uint8[n] may be NULL. */
HVB_API int hvb_mnv3_preprocess(hvb_ctx* ctx, const uint8_t* pixels_dev, const hvb_crop_desc* crops_dev, int n,
                        int roi_mode, float* out_dev, uint8_t* out_u8_dev, uint8_t* out_valid_dev);

/* ---------------------------------------------------------------- K4a: standardise + Gram / RBF affinity
 * StandardScaler.fit_transform (team_hybrid.py:166) and the affinity SpectralClustering(affinity='rbf',
 * gamma=1.0) builds inside fit (team_hybrid.py:185-193 -> sklearn pairwise rbf_kernel):
 *   d2 = max(|x|^2 + |y|^2 - 2 x.y, 0), diag 0;  A = exp(-gamma d2).
 * x_dev: float64[N,D] row-major.  mode 0: X.X^T on the tensor cores (tcgen05, split-TF32 operands,
 * fp32 accumulate in TMEM) + float64 refinement of the pairs whose affinity does not underflow;
 * mode 1: float64 CUDA-core path only.  Either output may be NULL.
 */
HVB_API int hvb_standardize(hvb_ctx* ctx, const double* x_dev, int n, int d, double* out_mean_dev,
                    double* out_scale_dev, double* out_xs_dev);
HVB_API int hvb_scale_transform(hvb_ctx* ctx, const double* x_dev, int n, int d, const double* mean_dev,
                        const double* scale_dev, double* out_xs_dev);
HVB_API int hvb_gram_affinity(hvb_ctx* ctx, const double* x_dev, int n, int d, double gamma, int mode,
                      double* out_d2_dev, double* out_a_dev);
/* The tensor-core Gram alone: G = X.X^T as float32[N,N] (bench / roofline hook). */
HVB_API int hvb_gram_tc(hvb_ctx* ctx, const double* x_dev, int n, int d, float* out_g_dev);

/* ---------------------------------------------------------------- K4b: ByteTrack IoU cost
 * matching.iou_distance (+ fuse_score) of supervision's ByteTrack reached from hockey/main.py:228,265:
 * cost[t,d] = 1 - IoU(a[t], b[d])  (nan -> IoU 0), optionally 1 - IoU*score[d].  float64 like numpy.
 * Batched over independent problems (clips): problem p uses rows a_off[p]..a_off[p+1]-1 of a,
 * b_off[p]..b_off[p+1]-1 of b and writes a dense [na_p, nb_p] block at out_off[p].
 * flags bit 0 / bit 1: the a / b boxes came from a float32 array — numpy then computes that side's
 * area in float32 before promoting (detections are float32, Kalman track boxes float64).
 */
HVB_API int hvb_iou_cost(hvb_ctx* ctx, const double* a_dev, const double* b_dev, const double* scores_dev,
                 const int32_t* a_off_dev, const int32_t* b_off_dev, const int64_t* out_off_dev,
                 int n_problems, int max_na, int max_nb, int flags, double* out_dev);

/* ---------------------------------------------------------------- K5: backbone glue
 * The element-wise work BETWEEN the library convolutions of the YOLOv8 forward the reference runs at
 * hockey/main.py:179-184 (ultralytics nn/modules: Conv.forward_fuse = act(conv(x)+b), Bottleneck's
 * x + cv2(cv1(x)), the torch.cat of C2f / SPPF / Detect, nn.Upsample + Concat of the neck) and of the
 * MobileNetV3 forward at common/team_hybrid.py:73-81.  Tensors are float32 NHWC ("channels_last").
 *
 * hvb_bias_act:  y[p,c] = act(x[p,c] + bias[c]) (+ residual[p,c]);  act: 0 none, 1 SiLU (expf + IEEE division,
 *   the formula of torch's kernel), 2 ReLU, 3 hardswish, 4 SiLU with ex2.approx / rcp.approx (<= 1e-6 relative error; the
 *   exact formula makes the in-place pass ALU-bound: 5.5 vs 7.0 TB/s).
 *   x_dev is the dense [npix, channels] raw convolution output; y goes to out1 (all channels, row pitch
 *   out1_ld floats, starting at channel out1_off of each row; may alias x_dev) and/or to out2 (only
 *   channels [c2_begin, c2_begin+c2_count), written at out2_off of rows of pitch out2_ld) — i.e. the
 *   contiguous tensor the next convolution reads and the slice of a concat buffer, in one pass.
 *   out2_up2_h/w > 0: the rows of x are the pixels of [n, h, w] images and out2 is a [n, 2h, 2w] tensor;
 *   every pixel is written to its 2x2 nearest-upsampled positions (nn.Upsample(2) + Concat of the neck).
 * hvb_concat_nhwc: out[n,y,x,:] = cat_s src_s[n, y >> shift_s, x >> shift_s, :]  (nearest 2^shift upsample).
 * hvb_stem_conv: layer 0 (3 -> c_out, 3x3, stride 2, pad 1) + bias + SiLU, reading K1's NCHW output and
 *   writing NHWC; exact fp32.  weight_host: float32[c_out,3,3,3] (PyTorch layout), bias_host: [c_out] or NULL;
 *   both are host pointers (they travel as kernel parameters).  c_out in {16,32,48,64}.
 */
HVB_API int hvb_bias_act(hvb_ctx* ctx, const float* x_dev, const float* bias_dev, const float* residual_dev,
                 int64_t npix, int channels, int act, float* out1_dev, int64_t out1_ld, int64_t out1_off,
                 float* out2_dev, int64_t out2_ld, int64_t out2_off, int c2_begin, int c2_count,
                 int out2_up2_h, int out2_up2_w);
HVB_API int hvb_concat_nhwc(hvb_ctx* ctx, const float* const src_dev[4], const int32_t src_channels[4],
                    const int32_t src_shift[4], int n_src, int n, int h, int w, float* out_dev);
/* SPPF.forward's pooling + concat (ultralytics nn/modules/block.py): out[n,y,x,:] = [y0, m(y0), m(m(y0)), m(m(m(y0)))]
 * with m = MaxPool2d(5, stride 1, pad 2), computed on chip from one read of y0.  y0: [n,h,w,channels] NHWC,
 * out: [n,h,w,4*channels].  HVB_ERR_UNSUPPORTED when an h x w map does not fit shared memory (h*w > ~6400). */
HVB_API int hvb_sppf_pool_concat(hvb_ctx* ctx, const float* y0_dev, int n, int h, int w, int channels, float* out_cat_dev);
HVB_API int hvb_stem_conv(hvb_ctx* ctx, const float* in_nchw_dev, const float* weight_host, const float* bias_host,
                  int n, int h, int w, int c_out, float* out_nhwc_dev);

/* ---------------------------------------------------------------- K6: pointwise convolution + epilogue
 * A 1x1 convolution of the YOLO forward (ultralytics Conv / C2f.cv1 / C2f.cv2 / SPPF / the last Conv2d of a Detect
 * branch; hockey/main.py:179-184) on NHWC float32 as one tcgen05 TF32 GEMM with the K5 epilogue fused in:
 *   out[p, co] = act(sum_ci x[p * x_ld + ci] * w[co * c_in + ci] + bias[co])
 * written to out1[p * out1_ld + out1_off + co] and, for co in [c2_begin, c2_begin + c2_count), also to
 * out2[p * out2_ld + out2_off + co - c2_begin] (out2 may be NULL).  act: 0 none, 1 SiLU, 4 fast SiLU (as hvb_bias_act).
 * c_in must be a multiple of 32 and c_out of 96 or 64 (HVB_ERR_UNSUPPORTED otherwise: the caller keeps the
 * cuDNN convolution + hvb_bias_act for that layer); pointers 16-byte aligned, pitches / offsets multiples of 4. */
HVB_API int hvb_pointwise_conv(hvb_ctx* ctx, const float* x_dev, int x_ld, const float* w_dev, const float* bias_dev,
                       int64_t npix, int c_in, int c_out, int act, float* out1_dev, int out1_ld, int out1_off,
                       float* out2_dev, int out2_ld, int out2_off, int c2_begin, int c2_count);

/* ---------------------------------------------------------------- K7: device ByteTrack
 * sv.ByteTrack as the reference constructs it (hockey/main.py:162-168: the main tracker; :207-211: the temporary one of
 * initialize_team_classifier) and steps it once per frame with update_with_detections (:228, :265).  One tracker object
 * holds the state of n_clips independent clips (tracking is sequential per clip, clips are independent).
 *   track_activation_threshold, minimum_matching_threshold, minimum_consecutive_frames: the constructor arguments;
 *   det_threshold = track_activation_threshold + 0.1 and max_time_lost = int(frame_rate / 30 * lost_track_buffer),
 *   both evaluated by the caller in double / Python arithmetic exactly like supervision does.
 * hvb_bytetrack_update consumes the detector's per-image outputs in K2a's layout (xyxy [images, max_det, 4], conf and
 * cls [images, max_det], count [images]; image of (clip c, frame f) = c * clip_stride + f * frame_stride) for n_frames
 * consecutive frames of every clip in ONE launch.  A detection enters the tracker iff conf > min_conf and bit cls of
 * class_mask is set (the mask of main.py:189-193; cls_dev may be NULL when class_mask is all ones).  Per image it
 * writes the detections supervision's `detections[tracker_id != -1]` keeps, in detection order: their row index
 * (out_row [images, max_det]), tracker id (out_tid) and number (out_count [images]).  out_count = -1: a capacity was
 * exceeded (256 live tracks per clip, 320 detections per frame; sticky until hvb_bytetrack_reset); -2: the chunk was
 * rejected and the clip's state NOT advanced — either one of the clip's input counts was negative (K2a candidate overflow
 * pending a retry) or `seq` is not the sequence number the clip expects next (0, 1, 2, ... since create / reset: a chunk
 * queued behind a rejected one); resubmit rejected chunks in order with their own seq.  Stream-ordered on the context's
 * stream, no host synchronisation.
 */
typedef struct hvb_bytetrack hvb_bytetrack;
HVB_API int hvb_bytetrack_create(hvb_ctx* ctx, int n_clips, double track_activation_threshold, double det_threshold,
                         double minimum_matching_threshold, int max_time_lost, int minimum_consecutive_frames,
                         hvb_bytetrack** out_tracker);
HVB_API int hvb_bytetrack_destroy(hvb_ctx* ctx, hvb_bytetrack* tracker);
HVB_API int hvb_bytetrack_reset(hvb_ctx* ctx, hvb_bytetrack* tracker);
HVB_API int hvb_bytetrack_update(hvb_ctx* ctx, hvb_bytetrack* tracker, const float* xyxy_dev, const float* conf_dev,
                         const int32_t* cls_dev /*or NULL*/, const int32_t* count_dev, int n_frames, int max_det,
                         int64_t clip_stride, int64_t frame_stride, float min_conf, uint32_t class_mask, int seq,
                         int32_t* out_row_dev, int32_t* out_tid_dev, int32_t* out_count_dev);

/* ---------------------------------------------------------------- K8: device spectral clustering (opt-in)
 * What SpectralClustering(affinity='precomputed', n_init=10, random_state=42).fit_predict does with K4a's affinity in
 * HybridTeamClassifier.fit (hockey/common/team_hybrid.py:185-193): sklearn.manifold.spectral_embedding(norm_laplacian=True,
 * drop_first=False) then KMeans on the N x k embedding.  SURVEY.md §8f rank 3; opt-in because the solver differs from
 * the reference's ARPACK / LOBPCG.  All float64; vectors come in blocks of hvb_spectral_block() = 8, stored [8][n].
 *   hvb_laplacian_normalize  scipy csgraph.laplacian(normed=True) pieces: dd = sqrt(column sums of A with a zero diagonal)
 *                            (1 for isolated nodes), M = D^-1/2 A0 D^-1/2 (L = I - M)
 *   hvb_sym_block_matvec     Y = (M + shift I) X — the one pass over the N x N matrix of a subspace-iteration step
 *   hvb_block_gram           mode 0: out[64] = A^T B;  mode 1 (a == b): out[0..63] = R^-1 with A^T A = R^T R (Cholesky QR),
 *                            out[64] = 1.0 when A^T A was not positive definite
 *   hvb_block_rotate         X <- X Q (and Y <- Y Q); with lambda: out_res[8] = |Y q_j - lambda_j X q_j|^2
 *   hvb_kmeans_lloyd         sklearn _kmeans_single_lloyd for n_init initialisations in one launch (one CTA each) on the
 *                            mean-centred x [n][d]: labels [n_init][n], centres [n_init][k][d], inertia, iterations, flags
 *                            (1 = a cluster went empty: the caller redoes that fit with sklearn's relocation); k, d <= 8 */
HVB_API int hvb_spectral_block(int* out_block);
HVB_API int hvb_laplacian_normalize(hvb_ctx* ctx, const double* a_dev, int n, double* m_dev, double* dd_dev);
HVB_API int hvb_sym_block_matvec(hvb_ctx* ctx, const double* m_dev, int n, const double* x_dev, double shift, double* y_dev);
HVB_API int hvb_block_gram(hvb_ctx* ctx, const double* a_dev, const double* b_dev, int n, int mode, double* out_dev /*[65]*/);
HVB_API int hvb_block_rotate(hvb_ctx* ctx, double* x_dev, double* y_dev /*or NULL*/, int n, const double* q_dev /*[8][8]*/,
                     const double* lambda_dev /*[8] or NULL*/, double* out_res_dev /*[8] or NULL*/);
HVB_API int hvb_kmeans_lloyd(hvb_ctx* ctx, const double* x_dev, int n, int d, int k, const double* init_centers_dev, int n_init,
                     int max_iter, double tol, int32_t* out_labels_dev, double* out_centers_dev, double* out_inertia_dev,
                     int32_t* out_n_iter_dev, int32_t* out_flags_dev);

/* ---------------------------------------------------------------- feature exchange for the global team fit
 * SURVEY.md §8b `hvb_allgather_features` / §8e: the one collective on the path.  Before HybridTeamClassifier.fit
 * (hockey/common/team_hybrid.py:155-196) standardises the crop features, every rank needs the float64[n_g, 625] rows of
 * all ranks in rank order, bit-identical.  NCCL is bound at run time (dlopen libnccl.so.2, or $HVB_NCCL_LIB): without it
 * these calls return HVB_ERR_UNSUPPORTED and nothing else in the library is affected.
 *   hvb_comm_unique_id  rank 0 makes the 128-byte id; the host ships it to the other ranks over its own channel
 *   hvb_comm_create     collective: ncclCommInitRank on the context's device
 *   hvb_comm_wrap       adopt a communicator the host already owns (ncclComm_t); not destroyed by hvb_comm_destroy
 *   hvb_allgather_counts    phase 1, synchronises: every rank's row count (the host sizes out_dev from them)
 *   hvb_allgather_features  phase 2, stream-ordered: an all-gather-v (one ncclGroup of per-rank broadcasts straight into
 *                           the compacted output, no padding pass); out_dev: float64[sum counts, d] */
#define HVB_COMM_ID_BYTES 128
typedef struct hvb_comm hvb_comm;
HVB_API int hvb_comm_unique_id(uint8_t* out_id128);
HVB_API int hvb_comm_create(hvb_ctx* ctx, const uint8_t* id128, int world, int rank, hvb_comm** out_comm);
HVB_API int hvb_comm_wrap(hvb_ctx* ctx, void* nccl_comm, hvb_comm** out_comm);
HVB_API int hvb_comm_destroy(hvb_ctx* ctx, hvb_comm* comm);
HVB_API int hvb_comm_info(hvb_comm* comm, int* out_world, int* out_rank);
HVB_API int hvb_allgather_counts(hvb_ctx* ctx, hvb_comm* comm, int n_local, int32_t* out_counts_host /*[world]*/,
                         int64_t* out_total /*or NULL*/);
HVB_API int hvb_allgather_features(hvb_ctx* ctx, hvb_comm* comm, const double* local_dev, int n_local, int d,
                           const int32_t* counts_host /*[world]*/, double* out_dev);

/* ---------------------------------------------------------------- host-buffer entry points
 * What a non-Python binding (cgo / JNI / N-API) would call: host in, host out, synchronous.
 * They stage through context-owned pinned + device scratch and run the same kernels. */
HVB_API int hvb_color_features_host(hvb_ctx* ctx, const uint8_t* pixels_host, size_t pixel_bytes,
                            const hvb_crop_desc* crops_host, int n, int roi_mode,
                            double* out_feat_host /*[n,49]*/, hvb_color_raw* out_raw_host /*or NULL*/);
HVB_API int hvb_jersey_color_stats_host(hvb_ctx* ctx, const uint8_t* pixels_host, size_t pixel_bytes,
                                const hvb_crop_desc* crops_host, int n, int roi_mode,
                                hvb_jersey_raw* out_raw_host /*[n]*/);
HVB_API int hvb_mnv3_preprocess_host(hvb_ctx* ctx, const uint8_t* pixels_host, size_t pixel_bytes,
                             const hvb_crop_desc* crops_host, int n, int roi_mode,
                             float* out_host /*[n,3,128,64]*/, uint8_t* out_valid_host /*or NULL*/);
HVB_API int hvb_merge_nms_host(hvb_ctx* ctx, const double* xyxy_host, const float* conf_host,
                       const int32_t* cls_host, int n, double iou_thres, int class_agnostic,
                       uint8_t* out_keep_host);
HVB_API int hvb_iou_cost_host(hvb_ctx* ctx, const double* a_host, int na, const double* b_host, int nb,
                      const double* scores_host /*or NULL*/, int flags, double* out_host /*[na,nb]*/);
HVB_API int hvb_gram_affinity_host(hvb_ctx* ctx, const double* x_host, int n, int d, double gamma, int mode,
                           double* out_d2_host /*or NULL*/, double* out_a_host /*or NULL*/);

#ifdef __cplusplus
}
#endif
#endif /* HVB_H_ */
