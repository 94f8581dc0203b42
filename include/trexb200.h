/*
 * trexb200.h -- C ABI of libtrexb200.so: TRex's per-frame segmentation + identification hot path
 * on one NVIDIA B200 (sm_100a).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Each entry point names the reference interface it replaces (paths relative to the TRex
 * checkout; C/ = Application/src/commons/common/, T/ = Application/src/tracker/).
 * INTEGRATION.md shows the reference-side shim that binds these.
 *
 * Conventions: every function returns TB_OK (0) or a negative tb_status; the message of the last
 * failure on the calling thread is tb_last_error().  No exceptions cross the ABI.  A handle is
 * single-producer (like the reference's `pipeline_async` thread, T/core/TaskPipeline.h:226-259).
 * There is NO CPU fallback: without a CUDA device tb_*_create fails with TB_ERR_CUDA.
 */
#ifndef TREXB200_H
#define TREXB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TB_ABI_VERSION 6

#if defined(__GNUC__)
#define TB_API __attribute__((visibility("default")))
#else
#define TB_API
#endif

typedef enum tb_status {
    TB_OK = 0,
    TB_ERR_INVALID = -1,      /* bad argument / unsupported configuration          */
    TB_ERR_CUDA = -2,         /* CUDA runtime failure (no device, OOM, launch)     */
    TB_ERR_STATE = -3,        /* call order (no background, no weights, no submit) */
    TB_ERR_CAPACITY = -4      /* a per-frame capacity was exceeded (see tb_frame_info.status) */
} tb_status;

/* Same memory layout as cmn::HorizontalLine (C/misc/detail.h:73-76): 8 bytes. */
typedef struct tb_line { uint16_t x0, x1, y, pad; } tb_line;

/* The settings keys RawProcessing::generate_binary caches (C/processing/RawProcessing.cpp:266-327)
 * plus those BackgroundSubtraction::apply reads (T/python/BackgroundSubtraction.cpp:137-139).
 * Defaults (tb_seg_default_params) are the reference's. */
typedef struct tb_seg_params {
    int32_t detect_threshold;             /* 15;  <0 inverts the mask (RawProcessing.cpp:537)        */
    int32_t threshold_maximum;            /* 255; <255 selects inRange[T,Tmax] (:529-531)            */
    int32_t enable_difference;            /* 1                                                        */
    int32_t detect_threshold_is_absolute; /* 1: |in-bg| ; 0: saturate(bg-in) (:393-399)              */
    int32_t image_invert;                 /* 0                                                        */
    int32_t use_closing;                  /* 0; dilate+erode with the elliptical (2k+1)^2 element (:438-448,504-505) */
    int32_t closing_size;                 /* 3 (1..7)                                                 */
    int32_t dilation_size;                /* 0; >0 dilate ones(n,n), <0 erode + re-threshold (:541-550); |n| <= 15 */
    float   cm_per_pixel;                 /* 1                                                        */
    int32_t n_size_ranges;                /* detect_size_filter: 0 = keep all; ranges are [lo,hi)    */
    double  size_lo[4], size_hi[4];
    int32_t color_channel;                /* -1 (none); 0..3: with gray encoding take this plane of a colour frame
                                             instead of cvtColor (T/python/BackgroundSubtraction.cpp:161-173)      */
    int32_t blur_difference;              /* 0; 1: difference -> zero values <= |T| -> cv::blur 25x25 -> > |T| (:371-387); replaces the
                                             rest of the mask pipeline; 1-channel backgrounds only                     */
    int32_t use_adaptive_threshold;       /* 0; 1: cv::adaptiveThreshold(MEAN_C, BINARY, n, -T) on the difference image instead of the
                                             plain threshold (:487,526), n = int(width * adaptive_threshold_scale) made odd, >= 3 (:427-434) */
    float   adaptive_threshold_scale;     /* 2 (T/core/default_config.cpp:1161)                        */
    int32_t open_size;                    /* 0; NOT a reference setting (the reference has no opening stage): the optional "2x2 morphological
                                             open" BASELINE.json's north_star names, applied to the threshold mask before use_closing /
                                             dilation_size: cv::morphologyEx(mask, MORPH_OPEN, ones(n,n)) = erode then dilate, default anchor
                                             (n/2, n/2), outside pixels ignored; 0 / 1 = off, 2..15                                      */
} tb_seg_params;

typedef struct tb_seg_config {
    int32_t device;               /* CUDA ordinal                                                    */
    int32_t width, height;        /* frame size                                                      */
    int32_t max_batch;            /* frames per submit                                               */
    int32_t max_runs_per_frame;   /* capacity of the run list of one frame (0 -> 32768)              */
    int32_t max_pixels_per_frame; /* capacity of blob pixel bytes of one frame (0 -> width*height/4) */
    int32_t max_crops_per_frame;  /* crops rendered per frame, first K blobs in canonical order (0 = no crops) */
    int32_t crop_width, crop_height; /* individual_image_size, 80x80 (T/core/default_config.cpp:1091) */
    int32_t crop_method;          /* 0 grey, 1 |bg-px|, 2 max(0,bg-px)  (Background.h:231-294)       */
    int32_t channels;             /* bytes per pixel of the submitted frames: 0/1 gray, 3 BGR, 4 BGRA (interleaved;
                                     TileImage.images[0], T/python/BackgroundSubtraction.cpp:151-188)               */
    int32_t encoding;             /* meta_encoding: 0 gray (colour frames -> cv::cvtColor(BGR[A]2GRAY) or color_channel;
                                     1 byte per blob pixel), 1 rgb8 (mask from the grey images, B,G,R per blob pixel,
                                     blob flag is_rgb; RawProcessing.cpp:355-358,557-593; needs channels 3 or 4),
                                     2 r3g3b2 (frames become 1-byte codes, convert_to_r3g3b2, C/misc/detail.h:508-555 +
                                     T/python/BackgroundSubtraction.cpp:151-158; the 1-channel path then runs on the codes against a
                                     1-channel background of codes; 1 byte per blob pixel, blob flag is_r3g3b2; crops are rendered as
                                     B,G,R like imageFromLines does for such blobs, C/processing/Background.cpp:134-139)          */
    int32_t crop_normalize;       /* individual_image_normalization (T/tracking/FilterCache.cpp:318-346): 0 none (centre pad /
                                     crop, :158-235), 1 moments (rotation by the blob's second-moment orientation through
                                     cv::warpAffine, :329-341 + :21-115), 2 posture (the reference's default, T/core/default_config.cpp:1089),
                                     3 legacy: the blob image warped by Midline::transform (:266-276 + :21-115) -- the crops of a batch
                                     are the `none` crops until tb_seg_posture has run on it with normalize = 1, which re-renders them
                                     from the normalised midlines.  1..3: gray encoding */
    float   crop_scale;           /* individual_image_scale (T/tracking/FilterCache.cpp:178-180): the masked blob image is resized with
                                     cv::resize(INTER_NEAREST) before the pad / crop; 0 or 1 = no scaling; gray encoding, crop_normalize 0 */
} tb_seg_config;

/* Per-frame result header. status bit0: run capacity exceeded (frame dropped, n_blobs=0),
 * bit1: pixel capacity exceeded, bit2: more blobs than max_crops_per_frame (crops truncated).
 * px_begin / n_pixels count BYTES of the pixel arena (= pixels for gray, 3 per pixel for rgb8). */
typedef struct tb_frame_info {
    uint32_t blob_begin, n_blobs;     /* range in the batch's blob-record array   */
    uint32_t line_begin, n_lines;     /* range in the batch's line arena           */
    uint32_t px_begin, n_pixels;      /* range in the batch's pixel arena          */
    uint32_t n_runs, status;
} tb_frame_info;

/* Per-blob record (32 bytes, fixed stride: also the unit of the multi-GPU metadata all-gather).
 * Blobs of a frame are in canonical order: by (y,x0) of their first line.  bid is
 * pv::bid::from_data (C/misc/bid.h:87-94). */
typedef struct tb_blob_rec {
    uint32_t line_off, px_off;        /* absolute offsets into the batch arenas (px_off in bytes) */
    uint32_t n_lines, n_pixels;       /* n_pixels counts pixels (payload = n_pixels * bytes per pixel) */
    uint16_t x0, y0, x1, y1;          /* inclusive bounding box                     */
    uint32_t bid;
    uint32_t frame;                   /* index of the frame inside the batch        */
} tb_blob_rec;

/* Host view of one frame's blobs: what CPULabeling::run returns as blobs_t
 * (C/processing/CPULabeling.cpp:189-343) after the size filter of
 * BackgroundSubtraction::apply (T/python/BackgroundSubtraction.cpp:259,306).
 * Blob k owns lines[recs[k].line_off - line_base ...] etc.; pointers are valid until the next submit. */
typedef struct tb_blob_view {
    tb_frame_info info;
    const tb_blob_rec *recs;          /* n_blobs records                                            */
    const tb_line *lines;             /* frame's lines; blob k: lines + (recs[k].line_off - info.line_begin) */
    const uint8_t *pixels;            /* frame's pixel bytes; blob k: pixels + (recs[k].px_off - info.px_begin) */
} tb_blob_view;

typedef struct tb_seg tb_seg;

TB_API const char *tb_last_error(void);
TB_API int tb_abi_version(void);
TB_API int tb_device_count(void);

TB_API void tb_seg_default_params(tb_seg_params *p);

/* BackgroundSubtraction::BackgroundSubtraction / deinit (T/python/BackgroundSubtraction.cpp:50-84,118-120) */
TB_API int tb_seg_create(const tb_seg_config *cfg, tb_seg **out);
TB_API void tb_seg_destroy(tb_seg *h);

/* settings callbacks of generate_binary (RawProcessing.cpp:283-327) */
TB_API int tb_seg_set_params(tb_seg *h, const tb_seg_params *p);

/* BackgroundSubtraction::set_background / Data::set (T/python/BackgroundSubtraction.cpp:86-101).
 * stride in bytes between rows.  1-channel image (gray encoding). */
TB_API int tb_seg_set_background(tb_seg *h, const uint8_t *bg, int width, int height, int64_t stride);
/* The same with an explicit channel count: 1 for gray encoding, 3 (B,G,R interleaved) for rgb8, where the grey
 * background the threshold runs against is derived once on the device (_grey_average, RawProcessing.cpp:356-357). */
TB_API int tb_seg_set_background_c(tb_seg *h, const uint8_t *bg, int width, int height, int channels, int64_t stride);

/* BackgroundSubtraction::apply(std::vector<TileImage>&&) (T/python/BackgroundSubtraction.cpp:126-347):
 * n host frames (width*height*channels bytes each, `stride` bytes between rows) are copied to the device,
 * segmented, labelled and (optionally) cropped; results are copied back asynchronously.
 * tb_seg_wait blocks until they are on the host.  fetch: 0 = results stay on the device (only
 * per-frame headers and totals are fetched), 1 = blob records + lines + pixels, 2 = also the crops. */
TB_API int tb_seg_submit(tb_seg *h, const uint8_t *const *frames, int n, int64_t stride, int fetch);

/* Same for n packed frames already resident in device memory (n*width*height*channels bytes).
 * stream: a cudaStream_t (NULL = the handle's own stream); work is ordered on it. */
TB_API int tb_seg_submit_device(tb_seg *h, const void *frames_dev, int n, void *stream, int fetch);

TB_API int tb_seg_wait(tb_seg *h);

/* Stream (cudaStream_t) on which tb_seg_submit / tb_seg_set_background enqueue their copies and kernels
 * (default: a private stream).  Lets a caller chain the identification network behind the segmentation
 * of the same batch and overlap the H2D copy of the next batch on a second handle. */
TB_API int tb_seg_set_stream(tb_seg *h, void *stream);

/* blobs_t of frame i of the last batch (after tb_seg_wait, fetch=1). */
TB_API int tb_seg_result(tb_seg *h, int i, tb_blob_view *out);

/* Totals of the last batch (after tb_seg_wait): blobs, lines, pixel bytes, crops. */
TB_API int tb_seg_totals(tb_seg *h, uint32_t out[4]);

/* Device-side results of the last batch, for chaining without a host round trip:
 *  crops  u8 [n_crops][crop_h][crop_w][C] (NHWC, C = 1 gray / 3 rgb8; image::calculate_diff_image,
 *         T/tracking/FilterCache.cpp:158-235), n_crops_dev -> uint32 on the device,
 *  recs   tb_blob_rec array, infos tb_frame_info[max_batch]. */
TB_API int tb_seg_device_results(tb_seg *h, void **crops, void **n_crops_dev, void **crop_blob_index,
                          void **recs, void **infos);

/* Multi-GPU metadata unit: the per-frame headers, the identification outputs and the blob records of a batch live in
 * ONE device block laid out for the all-gather -- K2 / K3 write headers and records in place, the identification head writes the
 * arg-max identity and its probability per crop (pass top_id / top_p to tb_vi_set_top1) -- so the collective sends the first
 * gather_bytes of `base` without any packing step:
 *   base + off_infos   tb_frame_info[batch]
 *   base + off_top_id  uint32[batch * kmax]   identity of crop n (crop n = blob n while no frame exceeds kmax blobs)
 *   base + off_top_p   float [batch * kmax]
 *   base + off_recs    tb_blob_rec[...]       the first batch * kmax records are inside the gathered prefix
 * batch = max_batch, kmax = max_crops_per_frame of the handle (128 at 100 individuals, 256 for BASELINE config 4). */
typedef struct tb_meta_layout {
    void *base;
    uint64_t gather_bytes;
    uint64_t off_infos, off_top_id, off_top_p, off_recs;
    uint32_t batch, kmax;
} tb_meta_layout;
TB_API int tb_seg_metadata(tb_seg *h, tb_meta_layout *out);

/* Host copy of the crops of the last batch (after tb_seg_wait with fetch=2 and crops enabled). */
TB_API int tb_seg_crops(tb_seg *h, const uint8_t **crops, const uint32_t **crop_blob_index, uint32_t *n);

/* Tracker-side re-threshold ("next" row N3a): pixel::threshold_blob (C/processing/PixelTree.cpp:186-291) applied to
 * every blob of det's last batch, on the device.  trk is a second handle of the same frame size whose params carry
 * the tracker settings: detect_threshold = track_threshold (comparison >=, Background.h:415-427),
 * enable_difference = track_background_subtraction, detect_threshold_is_absolute = track_threshold_is_absolute,
 * size ranges = track_size_filter (or none).  Gray encoding: trk is a channels = 1 handle (it runs on the grey plane);
 * rgb8: trk has det's channels and encoding, pixels are compared through cmn::bgr2gray (Background.h:76-81) against
 * the background's grey image and keep their B,G,R bytes (test_pixels.cpp:1073-1166).  Like the entry the tracker calls (:344-356, size_range
 * (-1, -1)), sub-blobs with a payload of ONE byte are not handed on (`pixels->size() > 1`: a lone grey pixel goes, a lone rgb8 pixel stays).
 * Afterwards tb_seg_wait / tb_seg_result / tb_seg_crops / tb_seg_device_results on trk return the tracker-side blobs (and their crops: what the
 * reference feeds the CNN). */
TB_API int tb_seg_rethreshold(tb_seg *det, tb_seg *trk, int fetch);

/* pv::Blob::recount(threshold, background) (C/processing/PVBlob.cpp:934-1027) through Background::count_above_threshold
 * (C/processing/Background.h:430-489) for every blob of the handle's last batch (after tb_seg_wait): the number of the blob's pixels
 * whose difference to the background -- the handle's method: enable_difference = 0 none, else absolute / sign -- is >= threshold,
 * times SQR(cm_per_pixel); threshold 0 = num_pixels (:942-949).  What PrefilterBlobs compares with track_size_filter
 * (T/tracking/PrefilterBlobs.cpp:227).  out: n >= n_blobs floats on the host, blob order of the batch.  (pv::Blob::threshold, :1040-1116, is
 * only called by another detector, T/python/PrecomuptedDetection.cpp:729, and is not built; the tracker's own re-threshold is
 * tb_seg_rethreshold.) */
TB_API int tb_seg_recount(tb_seg *h, int threshold, float *out, uint32_t n);

/* Outlines ("next" row N4, first stage): pixel::find_outer_points (C/processing/PixelTree.cpp:497-651) for every blob of the
 * handle's last batch (after tb_seg_wait with fetch >= 1), the outline calculate_posture selects (the first of maximal
 * size, T/tracking/Posture.cpp:341-348), and Outline::resample(outline_resample) of it (T/tracking/Outline.cpp:724-766;
 * <= 0: no resampling).  Points are x,y float pairs relative to the blob's bounding-box origin (tb_blob_rec.x0, y0), pixel
 * centres at +0.5 -- the frame of the blob after add_offset(-bounds.pos()), Posture.cpp:337.  Record k belongs to blob k of
 * the batch (tb_blob_rec order).  The result pointers are valid until the next tb_seg_outlines call on the handle. */
typedef struct tb_outline_rec {
    uint32_t raw_off, n_raw;          /* range in raw_points (in points): the outline as find_outer_points returns it */
    uint32_t res_off, n_res;          /* range in points: after Outline::resample (ranges of different blobs do not
                                         overlap but need not be contiguous)                                          */
} tb_outline_rec;
TB_API int tb_seg_outlines(tb_seg *h, float outline_resample);
TB_API int tb_seg_outline_result(tb_seg *h, const tb_outline_rec **recs, const float **raw_points, const float **points, uint32_t *n_blobs);

/* Midlines ("next" row N4, second stage), after tb_seg_outlines on the same batch: for every blob Outline::calculate_midline
 * (T/tracking/Outline.cpp:768-868) on its resampled outline -- Outline::smooth (:330-452), offset_to_middle (:454-718: clockwise
 * orientation, the elliptic-Fourier approximation periodic::eft / ieft with outline_approximate harmonics, curvature, find_peaks;
 * tail = the highest curvature peak (peak_mode pointy) or the middle of the broadest one (broad, :621-650), head = the peak farthest
 * from it; the outline is rotated to start at the tail) and the pairing walk.
 * points: the outline as the midline saw it (smoothed, approximated, rotated), same ranges as tb_outline_rec.res_off / n_res;
 * segments: {pos.x, pos.y, height, l_length} per midline segment from the tail on, blob k's at [seg_off, seg_off + n_seg);
 * n_seg = 0 when the reference would return "Too few midline segments calculated." */
typedef struct tb_posture_params {
    int32_t outline_smooth_samples;        /* 4     T/core/default_config.cpp:890 */
    int32_t outline_smooth_step;           /* 1     :889 */
    int32_t outline_approximate;           /* 3     :888 (0..8) */
    float   outline_curvature_range_ratio; /* 0.03  :891 */
    float   midline_walk_offset;           /* 0.025 :892 */
    int32_t peak_mode;                     /* 0 pointy (:902), 1 broad */
    int32_t midline_start_with_head;       /* 0     :900 */
    int32_t midline_invert;                /* 0     :901 */
    int32_t midline_resolution;            /* 25    :894 (2..256): points of a normalised midline */
    float   midline_stiff_percentage;      /* 0.15  :893 */
} tb_posture_params;
typedef struct tb_midline_rec { uint32_t seg_off, n_seg; int32_t tail, head; } tb_midline_rec;
TB_API void tb_posture_default_params(tb_posture_params *p);
TB_API int tb_seg_midlines(tb_seg *h, const tb_posture_params *p);
TB_API int tb_seg_midline_result(tb_seg *h, const tb_midline_rec **recs, const float **points, const float **segments, uint32_t *n_blobs);

/* The whole posture chain of a batch in one asynchronous call ("next" row N4): outlines -> raw midlines -> (normalize = 1)
 * Midline::post_process (T/tracking/Outline.cpp:895-1062) and Midline::normalize (:1268-1456) per blob, i.e. what
 * Individual::calculate_midline_for returns (T/tracking/Individual.cpp:1348-1383) -> (crop_normalize 2 / 3) the crops of the batch
 * re-rendered through Midline::transform (Outline.cpp:1238-1256) and image::normalize_image (T/tracking/FilterCache.cpp:21-115).
 * Everything is enqueued behind the batch's kernels on the batch's stream -- tb_seg_posture may be called right after
 * tb_seg_submit[_device], before tb_seg_wait; the blob count is read on the device -- so the identification network can be
 * chained behind it.  tb_seg_posture_wait blocks until the requested results are on the host.
 * A normalised midline has midline_resolution segments {pos.x, pos.y, height, l_length} (blob k's at norm_points + 4 * k *
 * midline_resolution), pos rotated / translated so that the head end is the origin; len / angle / offset are Midline::len(),
 * angle(), offset().  n_points = 0 where the reference has no (normalised) midline; flags: bit0 inverted because of the movement
 * direction, bit1 post_process threw (std::out_of_range), bit2 normalize() returned nullptr. */
typedef struct tb_midline_norm { float len, angle, offx, offy; uint32_t n_points, flags; int32_t tail, head; } tb_midline_norm;
typedef struct tb_posture_request {
    tb_posture_params params;
    float   outline_resample;              /* 1  T/core/default_config.cpp:898; <= 0: no resampling */
    int32_t normalize;                     /* 0: raw midlines only; 1: + post_process + normalize (+ crops for crop_normalize 2 / 3) */
    int32_t fetch;                         /* 0: results stay on the device; 1: midline records + normalised midlines (+ crop validity);
                                              2: also outline records, outline points and raw midline segments */
    float   median_midline_length_px;      /* FilterCache::median_midline_length_px of the tracklet (FilterCache.cpp:24,54-58); used when
                                              median_midline_length_dev is NULL */
    float   individual_image_scale;        /* 1 (FilterCache.cpp:45); 0 = 1 */
    const float *move_direction_dev;       /* optional device array, 2 floats per blob: MovementInformation::direction
                                              (posture_direction_smoothing > 1, Individual.cpp:1365-1368); NULL = (0, 0) */
    const float *fix_length_dev;           /* optional device array, 1 float per blob: Midline::normalize(fix_length) as
                                              Individual::fixed_midline calls it (Individual.cpp:507-522); NULL = -1 (none) */
    const float *median_midline_length_dev;/* optional device array, 1 float per blob */
} tb_posture_request;
typedef struct tb_posture_view {
    uint32_t n_blobs, midline_resolution;
    const tb_midline_rec *midlines;        /* fetch >= 1 */
    const tb_midline_norm *normalized;     /* fetch >= 1, normalize = 1 */
    const float *norm_points;              /* fetch >= 1, normalize = 1 */
    const uint8_t *crop_valid;             /* fetch >= 1, crop_normalize 2 / 3: 1 per crop that has a posture-normalised image */
    const tb_outline_rec *outlines;        /* fetch >= 2 */
    const float *raw_points, *points, *segments;   /* fetch >= 2: find_outer_points' outline, the outline the midline saw, raw segments */
} tb_posture_view;
TB_API void tb_posture_default_request(tb_posture_request *r);
TB_API int tb_seg_posture(tb_seg *h, const tb_posture_request *r);
TB_API int tb_seg_posture_wait(tb_seg *h);
TB_API int tb_seg_posture_result(tb_seg *h, tb_posture_view *out);
/* posture::calculate_posture (T/tracking/Posture.cpp:305-400) for every blob of src's last batch (the tracker-side blobs): starting at
 * track_posture_threshold the blob is re-thresholded (pixel::threshold_get_biggest_blob, C/processing/PixelTree.cpp:297-340: the
 * sub-blob with the most pixels, the first among equals in canonical order), the longest outline of that sub-blob is taken in the
 * frame of the ORIGINAL blob's bounds (:337), resampled, and calculate_midline is tried; blobs without a midline go another round
 * with the threshold raised by 2 until the sub-blob keeps fewer than max(1, pixels / 10) pixels or the threshold reaches the
 * start + 100; a blob that never yields a midline keeps its first resampled outline (:381-397).  pst is a second handle of src's
 * geometry configured like the tracker side of tb_seg_rethreshold (its detect_threshold is driven by this call and restored); the
 * results live in pst but are indexed by SRC's blobs: tb_seg_posture_wait / _result / _device on pst.  With crop_normalize 2 / 3 on
 * src, src's crops are re-rendered from these midlines.  One stream synchronisation per round; returns the number of rounds (>= 1)
 * or a negative tb_status. */
TB_API int tb_seg_posture_thresholded(tb_seg *src, tb_seg *pst, const tb_posture_request *r, int track_posture_threshold);
/* Device pointers of the last tb_seg_posture call (valid until the next one): records per blob of the batch. */
TB_API int tb_seg_posture_device(tb_seg *h, void **outline_recs, void **midline_recs, void **normalized, void **norm_points,
                                 void **points, void **segments);
/* Measurement hook: summed durations (ms) of {outlines, midlines (+ normalisation), posture crops} and the number of calls. */
TB_API int tb_seg_posture_ms(tb_seg *h, double out_ms[3], uint64_t *n_calls);

/* Debug / parity: generate_binary's output image (mask & input) of one device- or host-resident
 * frame, RawProcessing.cpp:597-600.  out is width*height (gray) or width*height*3 (rgb8) host bytes. */
TB_API int tb_seg_debug_binary(tb_seg *h, const uint8_t *frame_host, uint8_t *out_host);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
TB_API uint64_t tb_seg_launch_count(tb_seg *h);

/* Measurement hooks (no reference counterpart; BackgroundSubtraction::fps() only averages wall time,
 * T/python/BackgroundSubtraction.cpp:21-35).  With profiling enabled every kernel launch is bracketed
 * by CUDA events on the launching stream; tb_seg_kernel_ms synchronises, then returns the summed
 * durations (ms) of {seg_rle, ccl_label, blob_emit} and the number of batches they cover, and resets. */
TB_API int tb_seg_profile(tb_seg *h, int enable);
TB_API int tb_seg_kernel_ms(tb_seg *h, double out_ms[3], uint64_t *n_batches);

/* ----------------------------------------------------------------------------------------------
 * Visual identification: Python::VINetwork (T/ml/VisualIdentification.h:104-133) +
 * predict_numpy (T/python/visual_recognition_torch.py:290-352) + V118_3
 * (T/python/visual_identification_network_torch.py:184-258) or one of the other custom networks
 * (tb_vi_config.arch), inference only.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tb_vi tb_vi;

typedef struct tb_vi_config {
    int32_t device;
    int32_t width, height, channels;   /* 80, 80, 1                                          */
    int32_t num_classes;               /* track_max_individuals                              */
    int32_t max_images;                /* capacity of one predict call                       */
    int32_t precision;                 /* 0: fp32 CUDA cores (parity mode); 1: bf16x3 split on tcgen05 tensor cores;
                                          2: fp16 operands, one MMA per k-step in conv2/conv3 (max|dlogit| ~3e-4) */
    int32_t arch;                      /* visual_identification_version (ModelFetcher, T/python/visual_identification_network_torch.py:537-567):
                                          0 v118_3 (default, :184-258), 1 v100 (:328-386), 2 v110 (:262-325), 3 v119 (:106-181),
                                          4 v200 (:30-103).  v119 / v200 run on fp32 CUDA cores (precision must be 0); v100 / v110 do
                                          with precision 0 and share v118_3's tensor-core kernels with precision 1 / 2 (conv3 and
                                          fc1 zero-padded to 128 channels; v110 then needs positive BatchNorm2d scales: its
                                          BatchNorm follows the max-pool and is folded into the filters)                */
} tb_vi_config;

TB_API int tb_vi_create(const tb_vi_config *cfg, tb_vi **out);
TB_API void tb_vi_destroy(tb_vi *h);

/* VINetwork::load_weights (T/ml/VisualIdentification.cpp:306-319): one call per state_dict entry,
 * name exactly as in the reference's state_dict ("model.conv1.weight", "model.bn1.running_var",
 * "model.bn4.bias", ...; torch layouts), then tb_vi_commit folds BatchNorm and uploads. */
TB_API int tb_vi_set_tensor(tb_vi *h, const char *name, const float *data, int64_t count);
TB_API int tb_vi_commit(tb_vi *h);

/* VINetwork::probabilities (sync form, VisualIdentification.h:115-133): n images of
 * height*width*channels u8 (NHWC) -> probs[n][num_classes] (softmax rows); logits optional (NULL). */
TB_API int tb_vi_predict(tb_vi *h, const uint8_t *images, int n, float *probs, float *logits);

/* Device-resident variant: images_dev u8 NHWC, n_dev optional device uint32 holding the actual
 * count (<= n_max); outputs are device pointers ([n_max][num_classes]); stream NULL = own stream. */
TB_API int tb_vi_predict_device(tb_vi *h, const void *images_dev, int n_max, const void *n_dev,
                         void *probs_dev, void *logits_dev, void *stream);
TB_API int tb_vi_wait(tb_vi *h);
/* Optional device outputs of the following predict calls: arg-max class (uint32) and its probability (float) per image,
 * i.e. the identity the tracker assigns (Tracker::predicted, T/tracking/Tracker.cpp:237-246); part of the per-blob
 * metadata gathered across GPUs.  NULL, NULL disables. */
TB_API int tb_vi_set_top1(tb_vi *h, void *ids_dev, void *probs_dev);
TB_API uint64_t tb_vi_launch_count(tb_vi *h);
/* As tb_seg_profile / tb_seg_kernel_ms for {conv1, conv2, conv3, fc1, head}; n = chunks covered. */
TB_API int tb_vi_profile(tb_vi *h, int enable);
TB_API int tb_vi_kernel_ms(tb_vi *h, double out_ms[5], uint64_t *n_chunks);

/* ----------------------------------------------------------------------------------------------
 * Background image generation ("next" row N2b): cmn::AveragingAccumulator
 * (C/video/AveragingAccumulator.{h,cpp}); method = averaging_method_t: 0 mean, 1 mode, 2 max, 3 min.
 * add() takes n packed width*height u8 frames; finalize() writes the width*height background that
 * tb_seg_set_background consumes (VideoSource::generate_average, C/video/VideoSource.cpp:940-1030).
 * ---------------------------------------------------------------------------------------------- */
typedef struct tb_avg tb_avg;
TB_API int tb_avg_create(int device, int width, int height, int method, tb_avg **out);
TB_API void tb_avg_destroy(tb_avg *h);
TB_API int tb_avg_add(tb_avg *h, const uint8_t *frames, int n);
TB_API int tb_avg_add_device(tb_avg *h, const void *frames_dev, int n, void *stream);
TB_API int tb_avg_finalize(tb_avg *h, uint8_t *out);

/* ----------------------------------------------------------------------------------------------
 * Host frame buffers.  The reference hands BackgroundSubtraction::apply frames that live in a pool of reusable
 * Image::Ptr buffers (buffers::TileBuffers, T/core/TileBuffers.h:13-22: ImageBuffers<Image::Ptr, ImageMaker, 16>, filled by
 * TileImage's constructor, T/core/TileImage.h:27-45, and returned with TileBuffers::get().move_back,
 * T/python/BackgroundSubtraction.cpp:336-339).  tb_seg_submit copies from the caller's pointer with cudaMemcpyAsync: from
 * pageable memory the driver stages through its own bounce buffers (about a third of the PCIe rate, and the call is not
 * asynchronous), so the pool's buffers should be page-locked -- either allocated here (an ImageMaker that calls
 * tb_host_alloc) or registered in place once (tb_host_register on the pool's 16 buffers).
 * ---------------------------------------------------------------------------------------------- */
TB_API int tb_host_alloc(size_t bytes, void **out);          /* page-locked, mapped for every device; 4096-byte aligned */
TB_API int tb_host_free(void *p);
TB_API int tb_host_register(void *p, size_t bytes);          /* page-lock an existing allocation (cudaHostRegisterPortable) */
TB_API int tb_host_unregister(void *p);

/* ----------------------------------------------------------------------------------------------
 * Plug-in entry, mirroring the reference's own dlopen convention (SURVEY.md s8b A''): the host resolves ONE C symbol --
 * the reference's is `extern "C" void trex_python_register()` (T/python/PythonEntryPoint.cpp:142-179), which fills a table
 * of function pointers (PythonImplInterface) and registers the detection back ends as
 * detect::BackendHooks{init, deinit, is_initializing, fps, apply, set_background} (T/python/BackendRegistry.h:10-22,
 * register_yolo_backend, T/python/YOLO.cpp:1738-1747) and the identification service as RecTaskBackend{init, deinit, predict}
 * (T/python/PythonBackendRegistry.cpp:52-58).  Ours is trex_b200_register(): it fills the pure-C table below (one process-wide
 * detection handle and one identification handle, like the reference's static BackgroundSubtraction::data()) and, when the
 * host passes a tb_host_table, hands it to the host's register_backend callback for detect type "background_subtraction".
 * The C++ shim of INTEGRATION.md s1 wraps the six detection pointers into a detect::BackendHooks.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tb_backend_table {
    uint32_t abi_version, size;                  /* TB_ABI_VERSION, sizeof(tb_backend_table) */
    /* detect::BackendHooks */
    int    (*init)(const tb_seg_config *cfg, const tb_seg_params *params);     /* BackgroundSubtraction::init (.cpp:56-84) */
    void   (*deinit)(void);                                                    /* ::deinit (:118-120) */
    int    (*is_initializing)(void);                                           /* hooks.is_initializing: 1 until a background is set */
    double (*fps)(void);                                                       /* ::fps (:21-35): mean of frames / elapsed per apply */
    int    (*apply)(const uint8_t *const *frames, int n, int64_t stride, tb_blob_view *views /* n entries */);   /* ::apply (:126-347) */
    int    (*set_background)(const uint8_t *bg, int width, int height, int channels, int64_t stride);           /* ::set_background (:86-90) */
    int    (*update_params)(const tb_seg_params *params);                      /* the settings callbacks (RawProcessing.cpp:283-327) */
    /* RecTaskBackend{init, deinit, predict} + VINetwork::load_weights */
    int    (*vi_init)(const tb_vi_config *cfg);
    void   (*vi_deinit)(void);
    int    (*vi_set_tensor)(const char *name, const float *data, int64_t count);
    int    (*vi_commit)(void);
    int    (*vi_predict)(const uint8_t *images, int n, float *probs);
    const char *(*last_error)(void);
} tb_backend_table;

typedef struct tb_host_table {                   /* what the host hands in (the reference passes GlobalSettings / the tile-buffer pool
                                                    through PythonImplInterface::set_settings, T/python/PythonEntryPoint.cpp) */
    uint32_t abi_version, size;                  /* TB_ABI_VERSION the host was compiled against, sizeof(tb_host_table) */
    void *user;
    void (*register_backend)(void *user, const char *detect_type, const tb_backend_table *table);
    void (*log)(void *user, int level, const char *message);      /* optional (may be NULL) */
} tb_host_table;

/* Returns TB_OK, or TB_ERR_INVALID when the host's abi_version differs.  host may be NULL (the table is then only reachable
 * through tb_backend()). */
TB_API int trex_b200_register(const tb_host_table *host);
TB_API const tb_backend_table *tb_backend(void);

#ifdef TB_DEBUG_EXPORTS
/* Debug / bring-up (not part of the product ABI; declared only with -DTB_DEBUG_EXPORTS): one tcgen05 "shifted GEMM"
 * D[128][N] = A[shift+m][:] . B[n][:]^T on bf16 operands in the channel-group-planar layout the convolution kernels use
 * (a: [n_cg][n_pos][8] bf16, b: [n_cg][N][8] bf16, d: [128][N] f32; all host pointers).  Exercised by tests/test_gpu_umma.py
 * through the symbol name tbdbg_umma_shifted_gemm. */
int tbdbg_umma_shifted_gemm(const void *a, int n_pos, int n_cg, int shift, const void *b, int N, float *d);
#endif

#ifdef __cplusplus
}
#endif
#endif /* TREXB200_H */
