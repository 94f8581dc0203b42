/*
 * trex_oracle.c -- CPU restatement of the TRex segmentation / crop hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path
 * in trex_b200/csrc.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * path never links or calls anything in oracle/.
 *
 * Parity status: PINNED for segmentation + labeling (reproduces every blob of
 * the reference's own videos/test.pv fixture bit-exactly, see
 * tests/golden/make_golden.py and tests/test_oracle_golden.py) and for
 * imageFromLines (known-answer vector of Application/Tests/test_pixels.cpp:
 * 1381-1466); generate_binary and BackgroundSubtraction::apply additionally on
 * the reference's own RawProcessing.cpp / BackgroundSubtraction.cpp compiled
 * unmodified and run with the real OpenCV (tests/test_oracle_ref_detect.py).  The pad/crop-to-80x80 geometry has no golden vector in the
 * reference; it is pinned on the reference's own FilterCache.cpp, compiled
 * unmodified (oracle/build_ref.py, tests/test_oracle_ref_filtercache.py).
 *
 * Every function cites the reference file:line it follows.  Paths are relative
 * to the reference checkout:
 *   C/ = Application/src/commons/common/    T/ = Application/src/tracker/
 *
 * Written from the reference's behaviour, not copied from its source: plain C,
 * flat arrays, union-find bookkeeping instead of the reference's
 * Brototype/DLList object graph.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include <pthread.h>
#include <unistd.h>

/* Same memory layout as HorizontalLine, C/misc/detail.h:73-76. */
typedef struct { uint16_t x0, x1, y, pad; } to_line_t;

/* Settings read by RawProcessing::generate_binary (C/processing/RawProcessing.cpp:266-327)
 * and BackgroundSubtraction::apply (T/python/BackgroundSubtraction.cpp:137-139). */
typedef struct {
    int32_t detect_threshold;         /* grabber default_config.cpp:98  (15)   */
    int32_t threshold_maximum;        /* :99 (255)                              */
    int32_t enable_difference;        /* :126 (true)                            */
    int32_t detect_threshold_is_absolute; /* T/core/default_config.cpp:1168 (true) */
    int32_t image_invert;             /* :1159 (false)                          */
    int32_t use_closing;              /* :1164 (false)                          */
    int32_t closing_size;             /* :1165 (3)                              */
    int32_t dilation_size;            /* :1163 (0)                              */
    float   cm_per_pixel;             /* BackgroundSubtraction.cpp:137          */
    int32_t n_size_ranges;            /* detect_size_filter: up to 4 half-open ranges */
    double  size_lo[4], size_hi[4];
    int32_t blur_difference;          /* grabber default_config.cpp:125 (false) */
    int32_t use_adaptive_threshold;   /* T/core/default_config.cpp:1162 (false) */
    float   adaptive_threshold_scale; /* :1161 (2)                              */
    int32_t open_size;                /* NOT a reference setting: BASELINE north_star's optional "2x2 morphological open" of the
                                         threshold mask, cv::morphologyEx(MORPH_OPEN, ones(n,n)); 0 / 1 = off (default)  */
} to_params_t;

/* Blob emission order (SURVEY.md s7 "Blob order"):
 *  0 canonical: by (y,x0) of the blob's first run (what the CUDA path emits)
 *  1 reference, current source: survivor chosen by OWN run count, children attached lazily
 *    (C/processing/CPULabeling.cpp:121-136, Brototype.cpp:84-137)
 *  2 reference, older immediate-merge build that wrote videos/test.pv: survivor absorbs the
 *    loser's runs (the commented-out variant CPULabeling.cpp:132-140) */
enum { TO_ORDER_CANONICAL = 0, TO_ORDER_REF_LAZY = 1, TO_ORDER_REF_ABSORB = 2 };

/* ------------------------------------------------------------------------------------------
 * Morphology helpers: cv::dilate / cv::erode with a binary structuring element, anchor at the
 * centre (k/2), one iteration, default border = pixels outside the image are ignored
 * (RawProcessing.cpp:504-518,541-550; element shapes :419,:442).
 * ------------------------------------------------------------------------------------------ */
static void morph(const uint8_t *src, uint8_t *dst, int w, int h,
                  const uint8_t *el, int kw, int kh, int ax, int ay, int is_dilate)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int v = is_dilate ? 0 : 255;
            for (int j = 0; j < kh; ++j) {
                int yy = y + j - ay;
                if (yy < 0 || yy >= h) continue;
                for (int i = 0; i < kw; ++i) {
                    if (!el[j * kw + i]) continue;
                    int xx = x + i - ax;
                    if (xx < 0 || xx >= w) continue;
                    int s = src[(size_t)yy * w + xx];
                    if (is_dilate) { if (s > v) v = s; } else { if (s < v) v = s; }
                }
            }
            dst[(size_t)y * w + x] = (uint8_t)v;
        }
}

/* cv::getStructuringElement(MORPH_ELLIPSE, (2k+1,2k+1)) as OpenCV computes it (imgproc/morph:
 * row i spans |dx| <= round(k * sqrt(1 - (dy/k)^2))).  Used at RawProcessing.cpp:442. */
static void ellipse_element(uint8_t *el, int k)
{
    int n = 2 * k + 1;
    double inv_r2 = k ? 1.0 / ((double)k * k) : 0.0;
    for (int i = 0; i < n; ++i) {
        int dy = i - k, j1 = 0, j2 = 0;
        if (abs(dy) <= k) {
            int dx = (int)lrint(k * sqrt((k * k - dy * dy) * inv_r2));
            j1 = k - dx; if (j1 < 0) j1 = 0;
            j2 = k + dx + 1; if (j2 > n) j2 = n;
        }
        for (int j = 0; j < n; ++j) el[i * n + j] = (j >= j1 && j < j2) ? 1 : 0;
    }
}

/* ------------------------------------------------------------------------------------------
 * generate_binary, gray (1-channel) input.  C/processing/RawProcessing.cpp:263-600:
 *   :364-369 image_invert (INPUT = 255 - input) | plain copy
 *   :390-399 enable_difference: absdiff(INPUT, avg)  |  saturating subtract(avg, INPUT)
 *   :529-534 threshold_maximum<255 ? inRange(d, T, Tmax) : threshold(d, |T|, 255, BINARY)  (strict >)
 *   :537-539 detect_threshold<0 -> mask = 255 - mask
 *   :504-505 use_closing: dilate, erode with the ellipse element
 *   :541-550 dilation_size>0: dilate ones(n,n);  <0: erode, then re-threshold diff under the mask
 *   :597-599 output = mask & input  (ORIGINAL input, grey values under the mask)
 *   :371-387 blur_difference: difference -> THRESH_TOZERO(|T|) -> cv::blur 25x25 -> THRESH_BINARY(|T|); nothing else applies
 *   :427-434,487,526 use_adaptive_threshold: cv::adaptiveThreshold(MEAN_C, BINARY, n, -T) on the difference image replaces
 *            the plain threshold, n = int(cols * adaptive_threshold_scale) made odd, at least 3
 * Not restated: tags (out of scope).
 * ------------------------------------------------------------------------------------------ */

/* Normalised box filter of an 8-bit image as cv::boxFilter / cv::blur compute it (OpenCV is a third-party dependency of
 * the reference, not vendored): window k x k around the pixel, border 0 = BORDER_REPLICATE (what cv::adaptiveThreshold
 * passes), 1 = BORDER_REFLECT_101 (cv::blur's default); result = the window sum divided by k*k, rounded to nearest
 * (k odd: never a tie).  Checked against cv2 4.13 for k = 3 .. 3841 in tests/test_oracle_golden.py.  OpenCV accumulates in
 * int32; a window whose sum reaches 2^31 (k*k*mean >= 2^31) would wrap there and does not here. */
static int border_index(int i, int n, int border)
{
    if (border == 0) return i < 0 ? 0 : (i >= n ? n - 1 : i);
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; if (i >= n) i = 2 * (n - 1) - i; }
    return i;
}
int to_box_mean(const uint8_t *src, int w, int h, int k, int border, uint8_t *dst)
{
    const int p = k / 2;
    int64_t *hs = (int64_t *)malloc(sizeof(int64_t) * (size_t)w * h);
    if (!hs) return -1;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int64_t s = 0;                             /* direct sum for the first pixel of a row, sliding update afterwards */
            if (x == 0) { for (int i = -p; i <= p; ++i) s += src[(size_t)y * w + border_index(i, w, border)]; }
            else s = hs[(size_t)y * w + x - 1] - src[(size_t)y * w + border_index(x - 1 - p, w, border)]
                                               + src[(size_t)y * w + border_index(x + p, w, border)];
            hs[(size_t)y * w + x] = s;
        }
    const int64_t kk = (int64_t)k * k;
    for (int x = 0; x < w; ++x) {
        int64_t s = 0;
        for (int i = -p; i <= p; ++i) s += hs[(size_t)border_index(i, h, border) * w + x];
        for (int y = 0; y < h; ++y) {
            if (y) s += hs[(size_t)border_index(y + p, h, border) * w + x] - hs[(size_t)border_index(y - 1 - p, h, border) * w + x];
            dst[(size_t)y * w + x] = (uint8_t)((2 * s + kk) / (2 * kk));
        }
    }
    free(hs);
    return 0;
}
/* neighbourhood of cv::adaptiveThreshold as generate_binary derives it, RawProcessing.cpp:427-434 */
int to_adaptive_neighbourhood(int cols, float scale)
{
    int n = (int)((float)cols * scale);
    if (n % 2 == 0) n++;
    if (n < 3) n = 3;
    return n;
}
/* the threshold mask (255 / 0) of generate_binary, before it is ANDed with the input (:597-599) */
static int gen_mask(const uint8_t *frame, const uint8_t *bg, int w, int h, const to_params_t *p, uint8_t *mask)
{
    const size_t n = (size_t)w * h;
    const int T = p->detect_threshold, aT = abs(T);
    const int need_morph = p->use_closing || p->dilation_size != 0 || p->open_size > 1;
    uint8_t *diff = NULL, *tmp = NULL;
    if (p->blur_difference) {          /* RawProcessing.cpp:371-387; the difference is taken regardless of enable_difference */
        uint8_t *tz = (uint8_t *)malloc(n), *bl = (uint8_t *)malloc(n);
        if (!tz || !bl) { free(tz); free(bl); return -1; }
        for (size_t i = 0; i < n; ++i) {
            int in = p->image_invert ? 255 - frame[i] : frame[i], d;
            if (p->detect_threshold_is_absolute) d = abs(in - (int)bg[i]);
            else { d = (int)bg[i] - in; if (d < 0) d = 0; }
            tz[i] = d > aT ? (uint8_t)d : 0;                          /* THRESH_TOZERO */
        }
        if (to_box_mean(tz, w, h, 25, 1, bl)) { free(tz); free(bl); return -1; }
        for (size_t i = 0; i < n; ++i) mask[i] = bl[i] > aT ? 255 : 0;
        free(tz); free(bl);
        return 0;
    }
    if (need_morph || p->use_adaptive_threshold) {
        diff = (uint8_t *)malloc(n); tmp = (uint8_t *)malloc(n);
        if (!diff || !tmp) { free(diff); free(tmp); return -1; }
    }
    for (size_t i = 0; i < n; ++i) {
        int in = p->image_invert ? 255 - frame[i] : frame[i];
        int d = in;
        if (p->enable_difference) {
            if (p->detect_threshold_is_absolute) d = abs(in - (int)bg[i]);
            else { d = (int)bg[i] - in; if (d < 0) d = 0; }
        }
        if (diff) diff[i] = (uint8_t)d;
        int m;
        if (p->threshold_maximum < 255) m = (d >= T && d <= p->threshold_maximum);
        else m = d > aT;
        if (T < 0) m = !m;
        mask[i] = m ? 255 : 0;
    }
    if (p->use_adaptive_threshold) {   /* dst = src > mean - C ? 255 : 0 with C = -T (:487,526), then the T < 0 inversion (:498,537) */
        if (to_box_mean(diff, w, h, to_adaptive_neighbourhood(w, p->adaptive_threshold_scale), 0, tmp)) { free(diff); free(tmp); return -1; }
        for (size_t i = 0; i < n; ++i) {
            int m = (int)diff[i] - (int)tmp[i] > T;
            if (T < 0) m = !m;
            mask[i] = m ? 255 : 0;
        }
    }
    if (p->open_size > 1) {            /* optional open of the threshold mask: erode, then dilate, ones(n,n), OpenCV's default anchor n/2 */
        int k = p->open_size;
        uint8_t *el = (uint8_t *)malloc((size_t)k * k);
        memset(el, 1, (size_t)k * k);
        morph(mask, tmp, w, h, el, k, k, k / 2, k / 2, 0);
        morph(tmp, mask, w, h, el, k, k, k / 2, k / 2, 1);
        free(el);
    }
    if (p->use_closing) {
        int k = p->closing_size, kn = 2 * k + 1;
        uint8_t *el = (uint8_t *)malloc((size_t)kn * kn);
        ellipse_element(el, k);
        morph(mask, tmp, w, h, el, kn, kn, k, k, 1);
        morph(tmp, mask, w, h, el, kn, kn, k, k, 0);
        free(el);
    }
    if (p->dilation_size != 0) {
        int k = abs(p->dilation_size);
        uint8_t *el = (uint8_t *)malloc((size_t)k * k);
        memset(el, 1, (size_t)k * k);
        if (p->dilation_size > 0) {
            morph(mask, tmp, w, h, el, k, k, k / 2, k / 2, 1);
            memcpy(mask, tmp, n);
        } else {
            /* erode; diff.copyTo(OUTPUT, mask) into a buffer that still holds older
             * contents is order-dependent in the reference (CALLCV ping-pong); we restate
             * the intended semantics: keep diff under the eroded mask, threshold again. */
            morph(mask, tmp, w, h, el, k, k, k / 2, k / 2, 0);
            for (size_t i = 0; i < n; ++i) mask[i] = (tmp[i] && diff[i] > aT) ? 255 : 0;
            if (p->use_closing) {
                int c = p->closing_size, kn = 2 * c + 1;
                uint8_t *ce = (uint8_t *)malloc((size_t)kn * kn);
                ellipse_element(ce, c);
                morph(mask, tmp, w, h, ce, kn, kn, c, c, 1);
                morph(tmp, mask, w, h, ce, kn, kn, c, c, 0);
                free(ce);
            }
        }
        free(el);
    }
    free(diff); free(tmp);
    return 0;
}

int to_generate_binary(const uint8_t *frame, const uint8_t *bg, int w, int h,
                       const to_params_t *p, uint8_t *out)
{
    const size_t n = (size_t)w * h;
    uint8_t *mask = (uint8_t *)malloc(n);
    if (!mask) return -1;
    if (gen_mask(frame, bg, w, h, p, mask)) { free(mask); return -1; }
    for (size_t i = 0; i < n; ++i) out[i] = mask[i] & frame[i];
    free(mask);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Colour inputs.  The arithmetic of cv::cvtColor lives in OpenCV (third party, not vendored in the
 * reference; conda/meta.yaml asks for opencv 4.x): COLOR_BGR2GRAY / COLOR_BGRA2GRAY on 8-bit data is
 * the fixed-point  Y = (B*3735 + G*19235 + R*9798 + 16384) >> 15  (imgproc color_rgb: RGB2Gray<uchar>,
 * BY15 / GY15 / RY15, shift 15); checked against cv2 4.13 in tests/test_oracle_color.py.
 * Call sites: T/python/BackgroundSubtraction.cpp:165-169 (gray / binary encoding),
 * C/processing/RawProcessing.cpp:355-358 (3-channel input of generate_binary).
 * ------------------------------------------------------------------------------------------ */
void to_bgr2gray(const uint8_t *src, int64_t npx, int cn, uint8_t *dst)
{
    for (int64_t i = 0; i < npx; ++i) {
        const uint8_t *q = src + (size_t)i * cn;
        dst[i] = (uint8_t)((q[0] * 3735 + q[1] * 19235 + q[2] * 9798 + 16384) >> 15);
    }
}

/* The tracker-side grey value of a BGR triple, C/processing/Background.h:76-81:
 * saturate(float(B)*0.114 + float(G)*0.587 + float(R)*0.299 + 0.5, 0, 255) in double, truncated. */
static uint8_t bgr2gray_tracker(const uint8_t *q)
{
    double v = (double)(float)q[0] * 0.114 + (double)(float)q[1] * 0.587 + (double)(float)q[2] * 0.299 + 0.5;
    if (v < 0) v = 0;
    if (v > 255) v = 255;
    return (uint8_t)v;
}
void to_bgr2gray_tracker(const uint8_t *src, int64_t npx, uint8_t *dst)
{
    for (int64_t i = 0; i < npx; ++i) dst[i] = bgr2gray_tracker(src + 3 * (size_t)i);
}

/* meta_encoding r3g3b2: vec_to_r3g3b2 / convert_to_r3g3b2 (C/misc/detail.h:508-555) and r3g3b2_to_vec /
 * convert_from_r3g3b2 (:515-531,557-590).  Known answers: Application/Tests/test_pixels.cpp:629-795. */
void to_convert_to_r3g3b2(const uint8_t *src, int64_t npx, int cn, uint8_t *dst)
{
    for (int64_t i = 0; i < npx; ++i) {
        const uint8_t *q = src + (size_t)i * cn;
        dst[i] = (uint8_t)(((uint8_t)(q[0] / 64) << 6) | ((uint8_t)(q[1] / 32) << 3) | ((uint8_t)(q[2] / 32) << 0));
    }
}
void to_convert_from_r3g3b2(const uint8_t *src, int64_t npx, uint8_t *dst3)
{
    for (int64_t i = 0; i < npx; ++i) {
        const uint8_t c = src[i];
        dst3[3 * i + 0] = (uint8_t)(((c >> 6) & 3) * 64);
        dst3[3 * i + 1] = (uint8_t)(((c >> 3) & 7) * 32);
        dst3[3 * i + 2] = (uint8_t)((c & 7) * 32);
    }
}

/* The detect-side colour handling of BackgroundSubtraction::apply (T/python/BackgroundSubtraction.cpp:151-188)
 * followed by generate_binary (C/processing/RawProcessing.cpp:355-358,557-599):
 *   encoding 0 gray: cn 3/4 -> cvtColor(BGR[A]2GRAY), or the plane `color_channel` (0..cn-1) when set (:171-173);
 *                    then the 1-channel generate_binary against the 1-channel background.  out: w*h bytes.
 *   encoding 1 rgb8: cn 4 -> BGRA2BGR (:177-178); mask from gray(input) vs gray(background) (bg3 is 3-channel,
 *                    _grey_average :356-357); out = mask & each of B,G,R (:581-589).  out: w*h*3 bytes.
 *   encoding 2 r3g3b2: cn 3/4 -> convert_to_r3g3b2 (:151-158); the 1-channel generate_binary then runs on the CODES
 *                    against the 1-channel background of codes (RawProcessing.cpp:344: average has input.channels()).
 *                    out: w*h bytes (mask & code); blob flag is_r3g3b2 (BackgroundSubtraction.cpp:221).
 * gray_out (optional, w*h): the grey plane the threshold ran on. */
int to_generate_binary_color(const uint8_t *frame, int cn, int encoding, int color_channel, const uint8_t *bg,
                             int w, int h, const to_params_t *p, uint8_t *out, uint8_t *gray_out)
{
    const size_t n = (size_t)w * h;
    uint8_t *g = (uint8_t *)malloc(n), *mask = (uint8_t *)malloc(n), *bgg = NULL;
    int ret = -1;
    if (!g || !mask) goto done;
    if (cn == 1) memcpy(g, frame, n);
    else if (encoding == 2) to_convert_to_r3g3b2(frame, (int64_t)n, cn, g);
    else if (encoding == 0 && color_channel >= 0 && color_channel < 4) {
        if (color_channel >= cn) goto done;
        for (size_t i = 0; i < n; ++i) g[i] = frame[i * cn + color_channel];
    } else to_bgr2gray(frame, (int64_t)n, cn, g);
    if (gray_out) memcpy(gray_out, g, n);
    if (encoding == 0 || encoding == 2) {
        if (gen_mask(g, bg, w, h, p, mask)) goto done;
        for (size_t i = 0; i < n; ++i) out[i] = mask[i] & g[i];
    } else {
        if (cn < 3) goto done;
        bgg = (uint8_t *)malloc(n);
        if (!bgg) goto done;
        to_bgr2gray(bg, (int64_t)n, 3, bgg);
        if (gen_mask(g, bgg, w, h, p, mask)) goto done;
        for (size_t i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) out[i * 3 + k] = mask[i] & frame[i * cn + k];
    }
    ret = 0;
done:
    free(g); free(mask); free(bgg);
    return ret;
}

/* ------------------------------------------------------------------------------------------
 * Source::extract_lines, C/processing/Source.cpp:156-255: per row, maximal runs of pixels
 * with any channel non-zero.  Returns the number of runs (written up to cap).
 * ------------------------------------------------------------------------------------------ */
int64_t to_extract_lines(const uint8_t *img, int w, int h, int c, to_line_t *runs, int64_t cap)
{
    int64_t n = 0;
    for (int y = 0; y < h; ++y) {
        const uint8_t *row = img + (size_t)y * w * c;
        int prev = 0, x0 = 0;
        for (int x = 0; x < w; ++x) {
            int set = 0;
            for (int k = 0; k < c; ++k) if (row[(size_t)x * c + k]) { set = 1; break; }
            if (set && !prev) { x0 = x; prev = 1; }
            else if (!set && prev) {
                if (n < cap) { runs[n].x0 = (uint16_t)x0; runs[n].x1 = (uint16_t)(x - 1); runs[n].y = (uint16_t)y; runs[n].pad = 0; }
                ++n; prev = 0;
            }
        }
        if (prev) {
            if (n < cap) { runs[n].x0 = (uint16_t)x0; runs[n].x1 = (uint16_t)(w - 1); runs[n].y = (uint16_t)y; runs[n].pad = 0; }
            ++n;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Labeling.  merge_lines two-pointer sweep, C/processing/CPULabeling.cpp:44-187, over the runs
 * of the last non-empty row and the current one; 8-connectivity test HLine.h:90-92 /
 * CPULabeling.cpp:60-62,91.  The reference's Brototype parent/child graph is replaced by a
 * union-find whose representative is the reference's *survivor*, so the DLList emission order
 * (creation order of surviving roots, CPULabeling.cpp:256-323) can be reproduced.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t parent;      /* union-find over blob ids                                      */
    int64_t size;        /* run count used by the survivor rule (own / absorbed)          */
    int32_t min_run;     /* smallest run index of the component (canonical order key)     */
} blob_t;

static int32_t bfind(blob_t *b, int32_t i)
{
    while (b[i].parent != i) { b[i].parent = b[b[i].parent].parent; i = b[i].parent; }
    return i;
}

/* In: runs sorted (y,x0).  Out: label[i] = dense blob index in the requested emission order.
 * Returns number of blobs, <0 on allocation failure. */
int64_t to_label_runs(const to_line_t *runs, int64_t n, int order, int32_t *label)
{
    if (n == 0) return 0;
    blob_t *bl = (blob_t *)malloc(sizeof(blob_t) * (size_t)n);
    int32_t *node = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    if (!bl || !node) { free(bl); free(node); return -1; }
    int32_t nb = 0;
    for (int64_t i = 0; i < n; ++i) node[i] = -1;

#define NEW_BLOB(i) do { bl[nb].parent = nb; bl[nb].size = 1; bl[nb].min_run = (int32_t)(i); node[i] = nb; ++nb; } while (0)

    /* first non-empty row: one blob per run (CPULabeling.cpp:213-226) */
    int64_t ps = 0, pe = 0;
    while (pe < n && runs[pe].y == runs[0].y) { NEW_BLOB(pe); ++pe; }
    int64_t cs = pe;
    while (cs < n) {
        int64_t ce = cs;
        while (ce < n && runs[ce].y == runs[cs].y) ++ce;
        int64_t cur = cs, prev = ps;
        while (cur < ce) {
            const to_line_t *c = &runs[cur];
            const to_line_t *q = prev < pe ? &runs[prev] : NULL;
            if (!q || (int)c->y > (int)q->y + 1 || (int)c->x1 + 1 < (int)q->x0) {
                if (node[cur] < 0) NEW_BLOB(cur);            /* :60-86 */
                ++cur;
            } else if ((int)c->x0 > (int)q->x1 + 1) {
                ++prev;                                       /* :91-95 */
            } else {
                int32_t pb = bfind(bl, node[prev]);
                if (node[cur] < 0) {                          /* :103-107 */
                    node[cur] = pb; bl[pb].size += 1;
                    if ((int32_t)cur < bl[pb].min_run) bl[pb].min_run = (int32_t)cur;
                } else {
                    int32_t cb = bfind(bl, node[cur]);
                    if (cb != pb) {                           /* :109-136 */
                        int32_t win = pb, lose = cb;
                        if (bl[pb].size <= bl[cb].size) { win = cb; lose = pb; }   /* :121-123 */
                        bl[lose].parent = win;
                        if (bl[lose].min_run < bl[win].min_run) bl[win].min_run = bl[lose].min_run;
                        if (order == TO_ORDER_REF_ABSORB) bl[win].size += bl[lose].size;
                    }
                }
                if (c->x1 <= q->x1) ++cur; else ++prev;        /* :181-184 */
            }
        }
        ps = cs; pe = ce; cs = ce;
    }
#undef NEW_BLOB

    /* dense numbering of surviving roots */
    int32_t *dense = (int32_t *)malloc(sizeof(int32_t) * (size_t)nb);
    if (!dense) { free(bl); free(node); return -1; }
    int64_t k = 0;
    if (order == TO_ORDER_CANONICAL) {
        /* roots ordered by their smallest run index == (y,x0) of the first run */
        for (int32_t b = 0; b < nb; ++b) dense[b] = -1;
        for (int64_t i = 0; i < n; ++i) {
            int32_t r = bfind(bl, node[i]);
            if (bl[r].min_run == (int32_t)i) dense[r] = (int32_t)k++;
        }
    } else {
        /* creation order of the survivors (DLList order) */
        for (int32_t b = 0; b < nb; ++b) dense[b] = (bl[b].parent == b) ? (int32_t)k++ : -1;
    }
    for (int64_t i = 0; i < n; ++i) label[i] = dense[bfind(bl, node[i])];
    free(dense); free(bl); free(node);
    return k;
}

/* size filter: T/python/BackgroundSubtraction.cpp:259 -> SizeFilters::in_range_of_one
 * (T/core/SizeFilters.cpp:36-53), Range<double>::contains half-open (C/misc/ranges.h:162-168);
 * num_pixels * SQR(cm_per_pixel) is evaluated in float (Float2_t, C/misc/vec2.h:6).
 * num_pixels = pixels->size() (:247-251): the BYTES of the blob's pixel payload, i.e. 3 per pixel for
 * rgb8 blobs (CPULabeling.cpp:302-311 stores `channels` bytes per pixel), 1 for gray / r3g3b2. */
static int size_ok(const to_params_t *p, uint64_t npx)
{
    if (p->n_size_ranges <= 0) return 1;
    float sq = p->cm_per_pixel * p->cm_per_pixel;
    float v = (float)npx * sq;
    for (int i = 0; i < p->n_size_ranges; ++i)
        if ((double)v >= p->size_lo[i] && (double)v < p->size_hi[i]) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * One frame through BackgroundSubtraction::apply's per-image body
 * (T/python/BackgroundSubtraction.cpp:209-313): generate_binary -> CPULabeling::run ->
 * materialise (CPULabeling.cpp:256-323: lines in (y,x0) order, pixel bytes concatenated) ->
 * size filter -> drop blobs with >= 65535 lines (:306).
 * Output is SoA:  line_off[k]..line_off[k+1] / px_off[k]..px_off[k+1] index lines[] / pixels[].
 * Returns number of kept blobs, or -(needed) if a capacity is too small (-1 on alloc failure).
 * ------------------------------------------------------------------------------------------ */
/* CPULabeling::run + materialisation + size filter on generate_binary's output image `bin` (w*h*c bytes,
 * c = 1 or 3: a pixel is set when any channel is non-zero, Source.cpp:200-206,221-231; pixel payload is
 * c bytes per pixel, CPULabeling.cpp:302-311).  px_off counts BYTES. */
static int64_t segment_binary(const uint8_t *bin, int w, int h, int c,
                              const to_params_t *p, int order,
                              to_line_t *lines, int64_t cap_lines,
                              uint8_t *pixels, int64_t cap_px,
                              int64_t *line_off, int64_t *px_off, int64_t cap_blobs)
{
    int64_t ret = -1;
    to_line_t *runs = NULL; int32_t *label = NULL;
    int64_t *cnt_l = NULL, *cnt_p = NULL, *cur_l = NULL, *cur_p = NULL; int32_t *remap = NULL;

    int64_t nr = to_extract_lines(bin, w, h, c, NULL, 0);
    runs = (to_line_t *)malloc(sizeof(to_line_t) * (size_t)(nr + 1));
    label = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nr + 1));
    if (!runs || !label) goto done;
    to_extract_lines(bin, w, h, c, runs, nr);
    int64_t nb = to_label_runs(runs, nr, order, label);
    if (nb < 0) goto done;

    cnt_l = (int64_t *)calloc((size_t)nb + 1, sizeof(int64_t));
    cnt_p = (int64_t *)calloc((size_t)nb + 1, sizeof(int64_t));
    cur_l = (int64_t *)calloc((size_t)nb + 1, sizeof(int64_t));
    cur_p = (int64_t *)calloc((size_t)nb + 1, sizeof(int64_t));
    remap = (int32_t *)malloc(sizeof(int32_t) * ((size_t)nb + 1));
    if (!cnt_l || !cnt_p || !cur_l || !cur_p || !remap) goto done;
    for (int64_t i = 0; i < nr; ++i) {
        cnt_l[label[i]] += 1;
        cnt_p[label[i]] += (int64_t)runs[i].x1 - runs[i].x0 + 1;
    }
    int64_t kept = 0, tl = 0, tp = 0;
    for (int64_t b = 0; b < nb; ++b) {
        if (size_ok(p, (uint64_t)cnt_p[b] * (uint64_t)c) && cnt_l[b] < 65535) {
            remap[b] = (int32_t)kept;
            if (kept < cap_blobs) { line_off[kept] = tl; px_off[kept] = tp; }
            cur_l[b] = tl; cur_p[b] = tp;
            tl += cnt_l[b]; tp += cnt_p[b] * c; ++kept;
        } else remap[b] = -1;
    }
    if (kept > cap_blobs || tl > cap_lines || tp > cap_px) {
        int64_t need = kept > tl ? kept : tl; if (tp > need) need = tp;
        ret = -(need + 2); goto done;
    }
    line_off[kept] = tl; px_off[kept] = tp;
    for (int64_t i = 0; i < nr; ++i) {            /* runs are visited in (y,x0) order */
        int32_t b = label[i];
        if (remap[b] < 0) continue;
        lines[cur_l[b]++] = runs[i];
        int64_t len = (int64_t)runs[i].x1 - runs[i].x0 + 1;
        memcpy(pixels + cur_p[b], bin + ((size_t)runs[i].y * w + runs[i].x0) * c, (size_t)len * c);  /* :307 */
        cur_p[b] += len * c;
    }
    ret = kept;
done:
    free(runs); free(label); free(cnt_l); free(cnt_p); free(cur_l); free(cur_p); free(remap);
    return ret;
}

int64_t to_segment_frame(const uint8_t *frame, const uint8_t *bg, int w, int h,
                         const to_params_t *p, int order,
                         to_line_t *lines, int64_t cap_lines,
                         uint8_t *pixels, int64_t cap_px,
                         int64_t *line_off, int64_t *px_off, int64_t cap_blobs,
                         uint8_t *binary_out /* optional w*h, may be NULL */)
{
    const size_t n = (size_t)w * h;
    uint8_t *bin = binary_out ? binary_out : (uint8_t *)malloc(n);
    if (!bin) return -1;
    int64_t ret = -1;
    if (!to_generate_binary(frame, bg, w, h, p, bin))
        ret = segment_binary(bin, w, h, 1, p, order, lines, cap_lines, pixels, cap_px, line_off, px_off, cap_blobs);
    if (!binary_out) free(bin);
    return ret;
}

/* The same for colour frames (cn = 3 BGR / 4 BGRA interleaved; see to_generate_binary_color): encoding 0 gray
 * yields 1 byte per pixel, encoding 1 rgb8 yields B,G,R per pixel (blob flag is_rgb, CPULabeling.cpp:193). */
int64_t to_segment_frame_color(const uint8_t *frame, int cn, int encoding, int color_channel, const uint8_t *bg,
                               int w, int h, const to_params_t *p, int order,
                               to_line_t *lines, int64_t cap_lines, uint8_t *pixels, int64_t cap_px,
                               int64_t *line_off, int64_t *px_off, int64_t cap_blobs,
                               uint8_t *binary_out /* optional w*h*(encoding ? 3 : 1) */)
{
    const int c = encoding == 1 ? 3 : 1;
    const size_t n = (size_t)w * h * c;
    uint8_t *bin = binary_out ? binary_out : (uint8_t *)malloc(n);
    if (!bin) return -1;
    int64_t ret = -1;
    if (!to_generate_binary_color(frame, cn, encoding, color_channel, bg, w, h, p, bin, NULL))
        ret = segment_binary(bin, w, h, c, p, order, lines, cap_lines, pixels, cap_px, line_off, px_off, cap_blobs);
    if (!binary_out) free(bin);
    return ret;
}

/* CPULabeling::run on an already-binary image (C/processing/CPULabeling.cpp:358-371), no
 * size filter: used for the render -> relabel idempotence property of
 * Application/Tests/test_matching.cpp:1556-1602. */
int64_t to_label_image(const uint8_t *img, int w, int h, int order,
                       to_line_t *lines, int64_t cap_lines, uint8_t *pixels, int64_t cap_px,
                       int64_t *line_off, int64_t *px_off, int64_t cap_blobs)
{
    to_params_t p; memset(&p, 0, sizeof p);
    p.detect_threshold = 0; p.threshold_maximum = 255; p.enable_difference = 0;
    p.cm_per_pixel = 1.f; p.n_size_ranges = 0;
    /* with enable_difference off and T=0, mask = (img > 0) and out = img: identity */
    return to_segment_frame(img, img, w, h, &p, order, lines, cap_lines, pixels, cap_px,
                            line_off, px_off, cap_blobs, NULL);
}

/* pv::bid::from_data, C/misc/bid.h:87-94 via blob_bid, C/processing/BlobIdentity.cpp:7-16 */
uint32_t to_blob_id(const to_line_t *first, int64_t n_lines)
{
    return ((uint32_t)(((uint32_t)first->x0 + first->x1 + 1) / 2) << 19)
         | (((uint32_t)first->y & 0x1FFFu) << 6)
         | ((uint32_t)((uint8_t)n_lines) % 64u);
}

/* ------------------------------------------------------------------------------------------
 * imageFromLines, gray input, C/processing/Background.cpp:113-299, restricted to what the
 * crop stage and the reference test use: mask + grey + difference images of bbox size,
 * threshold >= (Background.h:415-427), difference method per Background.h:231-294:
 *   method 0 none: value   1 absolute: |bg - v|   2 sign: max(0, bg - v)
 * bbox = lines_dimensions (C/misc/detail.cpp:440-460).  Outputs may be NULL.
 * Returns recount (pixels set); bbox in rect[4] = x,y,w,h.
 * ------------------------------------------------------------------------------------------ */
int64_t to_image_from_lines(const to_line_t *lines, int64_t n_lines, const uint8_t *px,
                            const uint8_t *bg, int bg_w, int method, int base_threshold,
                            int32_t rect[4], uint8_t *mask, uint8_t *grey, uint8_t *diffimg)
{
    int mx = 1 << 30, my = 1 << 30, Mx = -1, My = -1;
    for (int64_t i = 0; i < n_lines; ++i) {
        if (lines[i].x0 < mx) mx = lines[i].x0;
        if (lines[i].y < my) my = lines[i].y;
        if (lines[i].x1 > Mx) Mx = lines[i].x1;
        if (lines[i].y > My) My = lines[i].y;
    }
    int bw = Mx - mx + 1, bh = My - my + 1;
    rect[0] = mx; rect[1] = my; rect[2] = bw; rect[3] = bh;
    size_t n = (size_t)bw * bh;
    if (mask) memset(mask, 0, n);
    if (grey) memset(grey, 0, n);
    if (diffimg) memset(diffimg, 0, n);
    int64_t recount = 0;
    for (int64_t i = 0; i < n_lines; ++i)
        for (int x = lines[i].x0; x <= lines[i].x1; ++x, ++px) {
            int v = *px, d = v;
            if (method == 1) d = abs((int)bg[(size_t)lines[i].y * bg_w + x] - v);
            else if (method == 2) { d = (int)bg[(size_t)lines[i].y * bg_w + x] - v; if (d < 0) d = 0; }
            int set = base_threshold == 0 || d >= base_threshold;
            if (!set) continue;
            size_t o = (size_t)(lines[i].y - my) * bw + (x - mx);
            if (mask) mask[o] = 255;
            if (grey) grey[o] = (uint8_t)v;
            if (diffimg) diffimg[o] = (uint8_t)d;
            ++recount;
        }
    return recount;
}

/* ------------------------------------------------------------------------------------------
 * image::calculate_diff_image, T/tracking/FilterCache.cpp:158-235 ("none" normalisation,
 * individual_image_scale == 1): render the blob (difference image when
 * track_background_subtraction, else grey; :165-174), copy under its mask (:176), then centre
 * pad (left = d - d/2, :187-190) or centre crop (start = d - d/2, :211-227) to out_w x out_h.
 * method as in to_image_from_lines.  out is out_w*out_h bytes, zero filled.
 * ------------------------------------------------------------------------------------------ */
void to_crop_blob(const to_line_t *lines, int64_t n_lines, const uint8_t *px,
                  const uint8_t *bg, int bg_w, int method, int out_w, int out_h, uint8_t *out)
{
    int32_t r[4];
    int mx = 1 << 30, my = 1 << 30, Mx = -1, My = -1;
    for (int64_t i = 0; i < n_lines; ++i) {
        if (lines[i].x0 < mx) mx = lines[i].x0;
        if (lines[i].y < my) my = lines[i].y;
        if (lines[i].x1 > Mx) Mx = lines[i].x1;
        if (lines[i].y > My) My = lines[i].y;
    }
    int bw = Mx - mx + 1, bh = My - my + 1;
    uint8_t *img = (uint8_t *)malloc((size_t)bw * bh), *mask = (uint8_t *)malloc((size_t)bw * bh);
    if (method == 0) to_image_from_lines(lines, n_lines, px, bg, bg_w, 0, 0, r, mask, img, NULL);
    else             to_image_from_lines(lines, n_lines, px, bg, bg_w, method, 0, r, mask, NULL, img);
    /* offsets of the bbox image inside the output canvas (negative = cropped away) */
    int offx, offy;
    if (bw < out_w) { int d = out_w - bw; offx = d - d / 2; } else { int d = bw - out_w; offx = -(d - d / 2); }
    if (bh < out_h) { int d = out_h - bh; offy = d - d / 2; } else { int d = bh - out_h; offy = -(d - d / 2); }
    memset(out, 0, (size_t)out_w * out_h);
    for (int y = 0; y < bh; ++y) {
        int oy = y + offy; if (oy < 0 || oy >= out_h) continue;
        for (int x = 0; x < bw; ++x) {
            int ox = x + offx; if (ox < 0 || ox >= out_w) continue;
            if (mask[(size_t)y * bw + x]) out[(size_t)oy * out_w + ox] = img[(size_t)y * bw + x];
        }
    }
    free(img); free(mask);
}

/* imageFromLines for rgb8 blobs (input 3 channels -> output 3 channels, Background.cpp:134-139,209-221):
 * value = the B,G,R bytes, diff = per-channel difference against the 3-channel background
 * (DifferenceImpl on RGBArray, Background.h:238-262), a pixel is set when base_threshold == 0 or
 * bgr2gray(diff) >= base_threshold (is_value_different, Background.h:415-427, tracker grey formula :76-81).
 * Pinned by Application/Tests/test_pixels.cpp:1381-1466.  Images are bbox-sized, 3 bytes per pixel. */
int64_t to_image_from_lines_rgb(const to_line_t *lines, int64_t n_lines, const uint8_t *px,
                                const uint8_t *bg3, int bg_w, int method, int base_threshold,
                                int32_t rect[4], uint8_t *mask, uint8_t *image, uint8_t *diffimg)
{
    int mx = 1 << 30, my = 1 << 30, Mx = -1, My = -1;
    for (int64_t i = 0; i < n_lines; ++i) {
        if (lines[i].x0 < mx) mx = lines[i].x0;
        if (lines[i].y < my) my = lines[i].y;
        if (lines[i].x1 > Mx) Mx = lines[i].x1;
        if (lines[i].y > My) My = lines[i].y;
    }
    int bw = Mx - mx + 1, bh = My - my + 1;
    rect[0] = mx; rect[1] = my; rect[2] = bw; rect[3] = bh;
    size_t n = (size_t)bw * bh;
    if (mask) memset(mask, 0, n);
    if (image) memset(image, 0, n * 3);
    if (diffimg) memset(diffimg, 0, n * 3);
    int64_t recount = 0;
    for (int64_t i = 0; i < n_lines; ++i)
        for (int x = lines[i].x0; x <= lines[i].x1; ++x, px += 3) {
            uint8_t d[3];
            for (int k = 0; k < 3; ++k) {
                int v = px[k], b = method ? bg3[((size_t)lines[i].y * bg_w + x) * 3 + k] : 0, q = v;
                if (method == 1) q = abs(b - v);
                else if (method == 2) { q = b - v; if (q < 0) q = 0; }
                d[k] = (uint8_t)q;
            }
            int set = base_threshold == 0 || (int)bgr2gray_tracker(d) >= base_threshold;
            if (!set) continue;
            size_t o = (size_t)(lines[i].y - my) * bw + (x - mx);
            if (mask) mask[o] = 255;
            if (image) memcpy(image + o * 3, px, 3);
            if (diffimg) memcpy(diffimg + o * 3, d, 3);
            ++recount;
        }
    return recount;
}

/* calculate_diff_image for rgb8 blobs: as to_crop_blob with 3 bytes per pixel (out: out_w*out_h*3). */
void to_crop_blob_rgb(const to_line_t *lines, int64_t n_lines, const uint8_t *px,
                      const uint8_t *bg3, int bg_w, int method, int out_w, int out_h, uint8_t *out)
{
    int32_t r[4];
    int mx = 1 << 30, my = 1 << 30, Mx = -1, My = -1;
    for (int64_t i = 0; i < n_lines; ++i) {
        if (lines[i].x0 < mx) mx = lines[i].x0;
        if (lines[i].y < my) my = lines[i].y;
        if (lines[i].x1 > Mx) Mx = lines[i].x1;
        if (lines[i].y > My) My = lines[i].y;
    }
    int bw = Mx - mx + 1, bh = My - my + 1;
    uint8_t *img = (uint8_t *)malloc((size_t)bw * bh * 3), *mask = (uint8_t *)malloc((size_t)bw * bh);
    if (method == 0) to_image_from_lines_rgb(lines, n_lines, px, bg3, bg_w, 0, 0, r, mask, img, NULL);
    else             to_image_from_lines_rgb(lines, n_lines, px, bg3, bg_w, method, 0, r, mask, NULL, img);
    int offx, offy;
    if (bw < out_w) { int d = out_w - bw; offx = d - d / 2; } else { int d = bw - out_w; offx = -(d - d / 2); }
    if (bh < out_h) { int d = out_h - bh; offy = d - d / 2; } else { int d = bh - out_h; offy = -(d - d / 2); }
    memset(out, 0, (size_t)out_w * out_h * 3);
    for (int y = 0; y < bh; ++y) {
        int oy = y + offy; if (oy < 0 || oy >= out_h) continue;
        for (int x = 0; x < bw; ++x) {
            int ox = x + offx; if (ox < 0 || ox >= out_w) continue;
            if (mask[(size_t)y * bw + x]) memcpy(out + ((size_t)oy * out_w + ox) * 3, img + ((size_t)y * bw + x) * 3, 3);
        }
    }
    free(img); free(mask);
}

/* ------------------------------------------------------------------------------------------
 * individual_image_normalization = moments ("next" row N3b): constraints::diff_image
 * (T/tracking/FilterCache.cpp:329-341) -> calculate_normalized_diff_image (:133-154) -> normalize_image (:21-115).
 *
 * (1) pv::Blob::calculate_moments, C/processing/PVBlob.cpp:111-214:
 *     float accumulators, pixels visited line by line; centre = (m10/m00, m01/m00); central moments from INTEGER offsets
 *     vx = int(x0 - centre.x) ..., products in float; angle = 0.5 * fast_atan2(2 mu'11, mu'20 - mu'02)
 *     (C/misc/math.h:34-59: 3rd-order polynomial, no libm).
 * ------------------------------------------------------------------------------------------ */
static float fast_atan_f(float z) { const float n1 = 0.97239411f, n2 = -0.19194795f; return (n1 + n2 * z * z) * z; }
static float fast_atan2_f(float y, float x)
{
    if (x == 0.0f) return copysignf((float)M_PI_2, y);
    float abs_y = fabsf(y), r, angle;
    if (abs_y < fabsf(x)) { r = abs_y / fabsf(x); angle = fast_atan_f(r); }
    else { r = fabsf(x) / abs_y; angle = (float)(M_PI_2 - fast_atan_f(r)); }      /* M_PI_2 is a double constant */
    if (x < 0.0f) angle = (float)(M_PI - angle);
    if (y < 0.0f) angle = -angle;
    return angle;
}
/* The runs are visited in `packages` consecutive packages with their own float accumulators, merged in package order afterwards: 1 package up to 1000 runs,
 * min(runs, 4) beyond (PVBlob.cpp:118-122 num_threads; distribute_indexes, C/misc/ThreadPool.h:126-195: runs / packages per package, the last one takes the rest).
 * The split changes the float rounding of the sums once they pass 2^24, so it is part of the result. */
float to_blob_orientation(const to_line_t *lines, int64_t n_lines, float centre[2])
{
    const int packages = n_lines > 1000 ? (int)(n_lines < 4 ? n_lines : 4) : 1;
    const int64_t per = n_lines / packages > 1 ? n_lines / packages : 1;
    float m00 = 0, m01 = 0, m10 = 0;
    for (int p = 0; p < packages; ++p) {
        const int64_t b = p * per, e = p + 1 == packages ? n_lines : (p + 1) * per;
        float l00 = 0, l01 = 0, l10 = 0;
        for (int64_t i = b; i < e; ++i) {
            const unsigned my = lines[i].y;
            int mx = lines[i].x0;
            for (int x = lines[i].x0; x <= lines[i].x1; ++x, ++mx) { l00 += 1; l01 += 1 * my; l10 += mx * 1; }
        }
        m00 += l00; m01 += l01; m10 += l10;
    }
    const float cx = m10 / m00, cy = m01 / m00;
    if (centre) { centre[0] = cx; centre[1] = cy; }
    float mu00 = 0, mu02 = 0, mu11 = 0, mu20 = 0;
    for (int p = 0; p < packages; ++p) {
        const int64_t b = p * per, e = p + 1 == packages ? n_lines : (p + 1) * per;
        float l00 = 0, l02 = 0, l11 = 0, l20 = 0;
        for (int64_t i = b; i < e; ++i) {
            const int vy = (int)((lines[i].y) - cy);
            const int vy2 = vy * vy;
            int vx = (int)((lines[i].x0) - cx);
            for (int x = lines[i].x0; x <= lines[i].x1; ++x, ++vx) {
                const int vx2 = vx * vx;
                l00 += 1; l02 += 1 * vy2; l11 += (float)vx * (float)vy; l20 += vx2 * 1;
            }
        }
        mu00 += l00; mu02 += l02; mu11 += l11; mu20 += l20;
    }
    const float inv = 1.0f / mu00;
    const float n11 = mu11 * inv, n20 = mu20 * inv, n02 = mu02 * inv;
    return (float)(0.5 * fast_atan2_f(2 * n11, n20 - n02));
}

/* (2) the affine map handed to cv::warpAffine: FilterCache.cpp:329-339 (rotate by DEGREE(-orientation + pi/4), translate by
 *     -bounds.size()/2) combined behind normalize_image's translate(size/2) . scale(individual_image_scale = 1)
 *     (:47-62); gui::Transform arithmetic in double (C/gui/Transform.cpp:118-176), toCV (Transform.h:65-78).
 *     M = {m00, m01, m02, m10, m11, m12} row-major 2x3. */
void to_moments_matrix(float orientation, int bw, int bh, int out_w, int out_h, double M[6])
{
    const float deg = (-orientation + (float)M_PI * 0.25f) * (1.0f / (float)M_PI * 180.0f);       /* DEGREE(), float */
    const double rad = (double)deg * 3.141592654 / 180.0;
    const double c = cos(rad), s = sin(rad);
    const float txf = -((float)bw * 0.5f), tyf = -((float)bh * 0.5f);                               /* -bounds.size() * 0.5 (Vec2 of float) */
    const double tx = txf, ty = tyf;
    /* tr = rotation . translation */
    const double r02 = c * tx + (-s) * ty + 0.0 * 1.0, r12 = s * tx + c * ty + 0.0 * 1.0;
    /* normalize_image: translate(size * 0.5) . scale(1) . translate(0) . tr */
    const double ox = (double)((float)out_w * 0.5f), oy = (double)((float)out_h * 0.5f);
    M[0] = 1.0 * c + 0.0 * s + ox * 0.0;  M[1] = 1.0 * (-s) + 0.0 * c + ox * 0.0;  M[2] = 1.0 * r02 + 0.0 * r12 + ox * 1.0;
    M[3] = 0.0 * c + 1.0 * s + oy * 0.0;  M[4] = 0.0 * (-s) + 1.0 * c + oy * 0.0;  M[5] = 0.0 * r02 + 1.0 * r12 + oy * 1.0;
}

/* (3) cv::warpAffine, 8-bit 1 channel, INTER_LINEAR, BORDER_CONSTANT(0), forward matrix M (OpenCV 4.x imgproc
 *     imgwarp.cpp: the matrix is inverted in double, coordinates run in 10-bit fixed point with 5 fractional bits kept
 *     (AB_BITS 10, INTER_BITS 5), the four taps are weighted with 15-bit coefficients (32-fx)(32-fy)*32 ... and the sum
 *     is rounded with +2^14 >> 15).  OpenCV is third party (not vendored in the reference); checked bit-exactly against
 *     cv2 4.13 in tests/test_oracle_moments.py. */
void to_warp_affine_u8(const uint8_t *src, int sw, int sh, const double Min[6], uint8_t *dst, int dw, int dh)
{
    double M[6];
    memcpy(M, Min, sizeof(M));
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
    const double b1 = -M[0] * M[2] - M[1] * M[5], b2 = -M[3] * M[2] - M[4] * M[5];
    M[2] = b1; M[5] = b2;
    for (int y = 0; y < dh; ++y) {
        const int X0 = (int)lrint((M[1] * y + M[2]) * 1024) + 16, Y0 = (int)lrint((M[4] * y + M[5]) * 1024) + 16;
        for (int x = 0; x < dw; ++x) {
            const int X = (X0 + (int)lrint(M[0] * x * 1024)) >> 5, Y = (Y0 + (int)lrint(M[3] * x * 1024)) >> 5;
            const int sx = X >> 5, sy = Y >> 5, fx = X & 31, fy = Y & 31;
            int acc = 0;
            for (int k = 0; k < 4; ++k) {
                const int yy = sy + (k >> 1), xx = sx + (k & 1);
                const int w = ((k & 1) ? fx : 32 - fx) * ((k >> 1) ? fy : 32 - fy) * 32;
                if (yy >= 0 && yy < sh && xx >= 0 && xx < sw) acc += src[(size_t)yy * sw + xx] * w;
            }
            dst[(size_t)y * dw + x] = (uint8_t)((acc + (1 << 14)) >> 15);
        }
    }
}

/* (4) the crop: render the blob (difference image when track_background_subtraction, else grey; FilterCache.cpp:141-153),
 *     warp it into the zero-filled out_w x out_h canvas (:39-73; no pad / crop step remains because the canvas already
 *     has the output size).  method as in to_image_from_lines. */
void to_crop_blob_moments(const to_line_t *lines, int64_t n_lines, const uint8_t *px,
                          const uint8_t *bg, int bg_w, int method, int out_w, int out_h, uint8_t *out)
{
    int32_t r[4];
    int mx = 1 << 30, my = 1 << 30, Mx = -1, My = -1;
    for (int64_t i = 0; i < n_lines; ++i) {
        if (lines[i].x0 < mx) mx = lines[i].x0;
        if (lines[i].y < my) my = lines[i].y;
        if (lines[i].x1 > Mx) Mx = lines[i].x1;
        if (lines[i].y > My) My = lines[i].y;
    }
    const int bw = Mx - mx + 1, bh = My - my + 1;
    uint8_t *img = (uint8_t *)malloc((size_t)bw * bh);
    if (method == 0) to_image_from_lines(lines, n_lines, px, bg, bg_w, 0, 0, r, NULL, img, NULL);
    else             to_image_from_lines(lines, n_lines, px, bg, bg_w, method, 0, r, NULL, NULL, img);
    double M[6];
    to_moments_matrix(to_blob_orientation(lines, n_lines, NULL), bw, bh, out_w, out_h, M);
    to_warp_affine_u8(img, bw, bh, M, out, out_w, out_h);
    free(img);
}

/* ------------------------------------------------------------------------------------------
 * Tracker-side re-threshold of one frame's blobs ("next" row N3a): pixel::threshold_blob
 * (C/processing/PixelTree.cpp:186-291 behind the public entry :344-356) applied to every blob: line_without_grid keeps the pixels whose
 * difference to the background is >= threshold (Background::is_value_different, Background.h:415-427;
 * method 0 none: value, 1 absolute: |bg - v|, 2 sign: max(0, bg - v), Background.h:231-294), cutting each
 * line into sub-lines, then CPULabeling::run(lines, pixels) (CPULabeling.cpp:378-414) relabels them.
 * Input / output are SoA blob lists; sub-blobs of parent k come out consecutively (canonical order inside
 * a parent).  Returns the number of sub-blobs, or -(needed) if a capacity is too small.
 * ------------------------------------------------------------------------------------------ */
static int64_t rethreshold_frame_c(const to_line_t *lines, const int64_t *line_off, const uint8_t *px, const int64_t *px_off,
                             int64_t n_blobs, const uint8_t *bg, int bg_w, int method, int threshold,
                             to_line_t *olines, int64_t cap_lines, uint8_t *opx, int64_t cap_px,
                             int64_t *oline_off, int64_t *opx_off, int64_t cap_blobs, int c);

int64_t to_rethreshold_frame(const to_line_t *lines, const int64_t *line_off, const uint8_t *px, const int64_t *px_off,
                             int64_t n_blobs, const uint8_t *bg, int bg_w, int method, int threshold,
                             to_line_t *olines, int64_t cap_lines, uint8_t *opx, int64_t cap_px,
                             int64_t *oline_off, int64_t *opx_off, int64_t cap_blobs)
{
    return rethreshold_frame_c(lines, line_off, px, px_off, n_blobs, bg, bg_w, method, threshold,
                               olines, cap_lines, opx, cap_px, oline_off, opx_off, cap_blobs, 1);
}

/* The same for rgb8 blobs (3 bytes per pixel; px_off counts bytes): line_without_grid<{3, rgb8}, {1, gray}> compares the
 * pixel's tracker grey value (cmn::bgr2gray, Background.h:76-81, via diffable_pixel_value :84-160) with the background's
 * grey image, which Background's constructor derives with cv::cvtColor(BGR2GRAY) (Background.cpp:71-77): bg here is that
 * grey image.  Pinned by Application/Tests/test_pixels.cpp:1073-1166 and :1289-1379. */
int64_t to_rethreshold_frame_rgb(const to_line_t *lines, const int64_t *line_off, const uint8_t *px, const int64_t *px_off,
                                 int64_t n_blobs, const uint8_t *bg, int bg_w, int method, int threshold,
                                 to_line_t *olines, int64_t cap_lines, uint8_t *opx, int64_t cap_px,
                                 int64_t *oline_off, int64_t *opx_off, int64_t cap_blobs)
{
    return rethreshold_frame_c(lines, line_off, px, px_off, n_blobs, bg, bg_w, method, threshold,
                               olines, cap_lines, opx, cap_px, oline_off, opx_off, cap_blobs, 3);
}

/* pixel::threshold_blob as the tracker calls it (PixelTree.cpp:344-356, size_range = Rangel(-1, -1)) hands on only the sub-blobs with
 * pixels->size() > 1, i.e. more than ONE BYTE of payload: a single grey pixel is dropped, a single rgb8 pixel (3 bytes) is kept.
 * pixel::threshold_get_biggest_blob (:297-340, the posture loop) has no such rule: to_rethreshold_keep_single(1) switches it off. */
static int g_rt_keep_single = 0;
void to_rethreshold_keep_single(int on) { g_rt_keep_single = on; }

static int64_t rethreshold_frame_c(const to_line_t *lines, const int64_t *line_off, const uint8_t *px, const int64_t *px_off,
                             int64_t n_blobs, const uint8_t *bg, int bg_w, int method, int threshold,
                             to_line_t *olines, int64_t cap_lines, uint8_t *opx, int64_t cap_px,
                             int64_t *oline_off, int64_t *opx_off, int64_t cap_blobs, int c)
{
    int64_t kept = 0, tl = 0, tp = 0;
    for (int64_t k = 0; k < n_blobs; ++k) {
        const int64_t nl = line_off[k + 1] - line_off[k], np_ = (px_off[k + 1] - px_off[k]) / c;
        to_line_t *sub = (to_line_t *)malloc(sizeof(to_line_t) * (size_t)(np_ + 1));
        int64_t *src = (int64_t *)malloc(sizeof(int64_t) * (size_t)(np_ + 1));      /* pixel offset of each sub-line */
        int32_t *label = (int32_t *)malloc(sizeof(int32_t) * (size_t)(np_ + 1));
        int64_t ns = 0;
        const uint8_t *p = px + px_off[k];
        int64_t o = 0;
        for (int64_t i = 0; i < nl; ++i) {                      /* PixelTree.cpp:108-166 */
            const to_line_t *l = &lines[line_off[k] + i];
            int start = -1;
            for (int x = l->x0; x <= l->x1; ++x, ++o) {
                int v = c == 1 ? p[o] : (int)bgr2gray_tracker(p + 3 * o), d = v;
                if (method == 1) d = abs((int)bg[(size_t)l->y * bg_w + x] - v);
                else if (method == 2) { d = (int)bg[(size_t)l->y * bg_w + x] - v; if (d < 0) d = 0; }
                if (d >= threshold) { if (start < 0) start = x; }
                else if (start >= 0) {
                    sub[ns].x0 = (uint16_t)start; sub[ns].x1 = (uint16_t)(x - 1); sub[ns].y = l->y; sub[ns].pad = 0;
                    src[ns++] = o - (x - start); start = -1;
                }
            }
            if (start >= 0) {
                sub[ns].x0 = (uint16_t)start; sub[ns].x1 = l->x1; sub[ns].y = l->y; sub[ns].pad = 0;
                src[ns++] = o - (l->x1 + 1 - start);
            }
        }
        int64_t nb = to_label_runs(sub, ns, TO_ORDER_CANONICAL, label);
        if (nb < 0) { free(sub); free(src); free(label); return -1; }
        for (int64_t b = 0; b < nb; ++b) {
            if (!g_rt_keep_single) {                                /* :350-352: pixels->size() > 1 */
                int64_t bytes = 0;
                for (int64_t i = 0; i < ns; ++i) if (label[i] == b) bytes += ((int64_t)sub[i].x1 - sub[i].x0 + 1) * c;
                if (bytes <= 1) continue;
            }
            if (kept < cap_blobs) { oline_off[kept] = tl; opx_off[kept] = tp; }
            for (int64_t i = 0; i < ns; ++i) {
                if (label[i] != b) continue;
                const int64_t len = (int64_t)sub[i].x1 - sub[i].x0 + 1;
                if (tl < cap_lines) olines[tl] = sub[i];
                if (tp + len * c <= cap_px) memcpy(opx + tp, p + src[i] * c, (size_t)len * c);
                ++tl; tp += len * c;
            }
            ++kept;
        }
        free(sub); free(src); free(label);
    }
    if (kept > cap_blobs || tl > cap_lines || tp > cap_px) {
        int64_t need = kept > tl ? kept : tl; if (tp > need) need = tp;
        return -(need + 2);
    }
    oline_off[kept] = tl; opx_off[kept] = tp;
    return kept;
}

/* ------------------------------------------------------------------------------------------
 * AveragingAccumulator, C/video/AveragingAccumulator.cpp:23-196 (gray): method 0 mean (float sum,
 * cv::divide = true float division, convertTo = round half to even + saturate), 1 mode (per-pixel
 * histogram, first maximum = smallest value among ties, :168-171), 2 max, 3 min.
 * ------------------------------------------------------------------------------------------ */
int to_average(const uint8_t *frames, int n, int w, int h, int method, uint8_t *out)
{
    const size_t px = (size_t)w * h;
    for (size_t i = 0; i < px; ++i) {
        if (method == 0) {
            float s = 0.f;
            for (int f = 0; f < n; ++f) s += (float)frames[(size_t)f * px + i];
            float q = s / (float)n;
            long r = lrintf(q);
            out[i] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
        } else if (method == 1) {
            uint32_t hist[256]; memset(hist, 0, sizeof hist);
            for (int f = 0; f < n; ++f) hist[frames[(size_t)f * px + i]]++;
            int best = 0;
            for (int b = 1; b < 256; ++b) if (hist[b] > hist[best]) best = b;
            out[i] = (uint8_t)best;
        } else {
            int v = method == 3 ? 255 : 0;
            for (int f = 0; f < n; ++f) { int q = frames[(size_t)f * px + i]; v = method == 2 ? (q > v ? q : v) : (q < v ? q : v); }
            out[i] = (uint8_t)v;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Batch driver for the CPU baseline: n frames through to_segment_frame + to_crop_blob, frames
 * handed to `threads` pthreads through a shared counter (frames are independent:
 * BackgroundSubtraction::apply keeps no cross-frame state).  Returns total kept blobs; per-frame
 * counts in n_blobs[].  crops (optional): n * max_crops * out_w*out_h bytes.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const uint8_t *frames, *bg; int n, w, h; const to_params_t *p;
    int crop_method, out_w, out_h, max_crops; uint8_t *crops; int32_t *n_blobs;
    int next; pthread_mutex_t mu; int64_t total;
} batch_t;

static void *batch_worker(void *arg)
{
    batch_t *b = (batch_t *)arg;
    const int w = b->w, h = b->h;
    const int64_t capL = 1 << 20, capP = (int64_t)w * h, capB = 1 << 18;
    to_line_t *lines = (to_line_t *)malloc(sizeof(to_line_t) * (size_t)capL);
    uint8_t *px = (uint8_t *)malloc((size_t)capP);
    int64_t *lo = (int64_t *)malloc(sizeof(int64_t) * (size_t)(capB + 1));
    int64_t *po = (int64_t *)malloc(sizeof(int64_t) * (size_t)(capB + 1));
    int64_t local = 0;
    for (;;) {
        pthread_mutex_lock(&b->mu);
        int f = b->next++;
        pthread_mutex_unlock(&b->mu);
        if (f >= b->n) break;
        int64_t k = to_segment_frame(b->frames + (size_t)f * w * h, b->bg, w, h, b->p, TO_ORDER_CANONICAL,
                                     lines, capL, px, capP, lo, po, capB, NULL);
        if (k < 0) k = 0;
        b->n_blobs[f] = (int32_t)k;
        if (b->crops)
            for (int64_t i = 0; i < k && i < b->max_crops; ++i)
                to_crop_blob(lines + lo[i], lo[i + 1] - lo[i], px + po[i], b->bg, w, b->crop_method,
                             b->out_w, b->out_h,
                             b->crops + ((size_t)f * b->max_crops + (size_t)i) * b->out_w * b->out_h);
        local += k;
    }
    pthread_mutex_lock(&b->mu); b->total += local; pthread_mutex_unlock(&b->mu);
    free(lines); free(px); free(lo); free(po);
    return NULL;
}

int to_num_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int64_t to_segment_batch(const uint8_t *frames, int n, const uint8_t *bg, int w, int h,
                         const to_params_t *p, int crop_method, int out_w, int out_h,
                         int max_crops, uint8_t *crops, int32_t *n_blobs, int threads)
{
    batch_t b = { frames, bg, n, w, h, p, crop_method, out_w, out_h, max_crops, crops, n_blobs,
                  0, PTHREAD_MUTEX_INITIALIZER, 0 };
    if (threads <= 0) threads = to_num_threads();
    if (threads > n) threads = n > 0 ? n : 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (int i = 1; i < threads; ++i) pthread_create(&th[i], NULL, batch_worker, &b);
    batch_worker(&b);
    for (int i = 1; i < threads; ++i) pthread_join(th[i], NULL);
    return b.total;
}

/* ------------------------------------------------------------------------------------------
 * "Next" row N4, first stage: the outline of a blob.
 *   pixel::find_outer_points  C/processing/PixelTree.cpp:497-651  (+ finalize :393-495)
 *   Tree::add / generate_edges / add_edge / walk   :783-1130, :657-781
 *   Outline::resample         T/tracking/Outline.cpp:724-766
 * Restated as a literal emulation of the reference's bookkeeping: border pixels become nodes sorted by
 * leaf_index (x << 32 | y, PixelTree.h Node::leaf_index: column-major), every missing 4-neighbour side
 * (order TOP, LEFT, RIGHT, BOTTOM = direction_from_bool) emits one edge from its side midpoint to the next side midpoint
 * walking with the blob on the right hand (generate_edges :876-965), add_edge keeps the `non_full_nodes` / `_sides`
 * vectors with the reference's linear search, and walk() is the reference's deque walk (push_front of edges[0], edges[1]).
 * One deliberate simplification: the node set is "pixels with a missing 4-neighbour" -- what the streaming three-row scan
 * (:543-633) is built to find; a node without missing sides emits no edge, so a superset is harmless.
 * Pinned on the reference's own code: PixelTree.cpp is compiled unmodified by oracle/build_ref.py and tests/test_oracle_ref_pixeltree.py holds
 * every outline of 334 blobs (holes, diagonal contacts, single pixels, a 300-pixel-wide blob) to it point for point, in the reference's order.
 * ------------------------------------------------------------------------------------------ */
typedef struct { float x, y; int32_t e[2]; int walked; uint64_t index; } subnode_t;

static int px_set(const to_line_t *lines, int64_t n, int x, int y)
{
    if (x < 0 || y < 0) return 0;
    for (int64_t i = 0; i < n; ++i)
        if (lines[i].y == y && lines[i].x0 <= x && x <= lines[i].x1) return 1;
    return 0;
}

typedef struct { int x, y; uint8_t nb[8]; } onode_t;       /* nb[d]: pixel in Direction d set (TOP, TOPR, RIGHT, BOTTOMR, BOTTOM, BOTTOML, LEFT, TOPL) */
static int onode_cmp(const void *a, const void *b)
{
    const onode_t *p = (const onode_t *)a, *q = (const onode_t *)b;
    if (p->x != q->x) return p->x < q->x ? -1 : 1;
    return p->y < q->y ? -1 : (p->y > q->y);
}

/* All outlines of a blob in the order Tree::generate_edges returns them.  Coordinates are relative to the blob's
 * bounding box origin (the blob after add_offset(-bounds.pos()), Posture.cpp:337): pixel centres at +0.5.
 * pts: x,y pairs; loop_off[k]..loop_off[k+1] = points of outline k.  Returns the number of outlines, or -1 when a
 * capacity is too small. */
int64_t to_find_outer_points(const to_line_t *lines_in, int64_t n_lines, float *pts, int64_t cap_pts, int64_t *loop_off, int64_t cap_loops)
{
    static const int VX[8] = {0, 1, 1, 1, 0, -1, -1, -1}, VY[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
    static const int FROM_BOOL[4] = {0, 6, 2, 4};            /* direction_from_bool: TOP, LEFT, RIGHT, BOTTOM */
    if (n_lines <= 0) return 0;
    int mx = 1 << 30, my = 1 << 30;
    int64_t npx = 0;
    for (int64_t i = 0; i < n_lines; ++i) {
        if (lines_in[i].x0 < mx) mx = lines_in[i].x0;
        if (lines_in[i].y < my) my = lines_in[i].y;
        npx += lines_in[i].x1 - lines_in[i].x0 + 1;
    }
    to_line_t *lines = (to_line_t *)malloc(sizeof(to_line_t) * (size_t)n_lines);
    for (int64_t i = 0; i < n_lines; ++i) { lines[i] = lines_in[i]; lines[i].x0 -= mx; lines[i].x1 -= mx; lines[i].y -= my; }
    onode_t *nodes = (onode_t *)malloc(sizeof(onode_t) * (size_t)npx);
    int64_t nn = 0;
    for (int64_t i = 0; i < n_lines; ++i)
        for (int x = lines[i].x0; x <= lines[i].x1; ++x) {
            onode_t nd; nd.x = x; nd.y = lines[i].y;
            for (int d = 0; d < 8; ++d) nd.nb[d] = (uint8_t)px_set(lines, n_lines, x + VX[d], nd.y + VY[d]);
            if (!nd.nb[0] || !nd.nb[6] || !nd.nb[2] || !nd.nb[4]) nodes[nn++] = nd;
        }
    qsort(nodes, (size_t)nn, sizeof(onode_t), onode_cmp);
    const int64_t cap_sub = 4 * nn + 4;
    subnode_t *sub = (subnode_t *)malloc(sizeof(subnode_t) * (size_t)cap_sub);
    int32_t *non_full = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap_sub), *sides = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap_sub);
    int64_t n_sub = 0, n_nf = 0, n_sides = 0;
    for (int64_t k = 0; k < nn; ++k) {
        const onode_t *nd = &nodes[k];
        for (int bi = 0; bi < 4; ++bi) {
            const int border = FROM_BOOL[bi];
            if (nd->nb[border]) continue;                     /* node->border[i]: this 4-neighbour is missing */
            const int left = (border + 7) & 7, left_left = (border + 6) & 7, opposite = (border + 2) & 7;
            int bx, by, out_dir = border, in_dir;             /* Edge(out_direction, in_direction, A = node, B) */
            if (nd->nb[left]) { bx = nd->x + VX[left]; by = nd->y + VY[left]; in_dir = opposite; }
            else if (nd->nb[left_left]) { bx = nd->x + VX[left_left]; by = nd->y + VY[left_left]; in_dir = border; }
            else { bx = nd->x; by = nd->y; in_dir = left_left; }
            /* add_edge (:657-781) */
            const float ox = ((float)nd->x + 0.5f) + (float)VX[out_dir] * 0.5f, oy = ((float)nd->y + 0.5f) + (float)VY[out_dir] * 0.5f;
            const float ix = ((float)bx + 0.5f) + (float)VX[in_dir] * 0.5f, iy = ((float)by + 0.5f) + (float)VY[in_dir] * 0.5f;
            const uint64_t out_idx = (((uint64_t)(int64_t)(ox * 10)) << 32) | ((uint64_t)(int32_t)(oy * 10) & 0xFFFFFFFFull);
            const uint64_t in_idx = (((uint64_t)(int64_t)(ix * 10)) << 32) | ((uint64_t)(int32_t)(iy * 10) & 0xFFFFFFFFull);
            int32_t in_node = -1, out_node = -1;
            for (int64_t it = 0; it < n_nf;) {
                if (sub[non_full[it]].index == in_idx) {
                    in_node = non_full[it]; sides[n_sides++] = non_full[it];
                    memmove(non_full + it, non_full + it + 1, sizeof(int32_t) * (size_t)(n_nf - it - 1)); --n_nf;
                    if (in_node >= 0 && out_node >= 0) break;
                } else if (sub[non_full[it]].index == out_idx) {
                    out_node = non_full[it]; sides[n_sides++] = non_full[it];
                    memmove(non_full + it, non_full + it + 1, sizeof(int32_t) * (size_t)(n_nf - it - 1)); --n_nf;
                    if (in_node >= 0 && out_node >= 0) break;
                } else ++it;
            }
            const int found_out = out_node >= 0;
            if (out_node < 0) {
                subnode_t s; s.x = ox; s.y = oy; s.e[0] = in_node; s.e[1] = -1; s.walked = 0; s.index = out_idx;
                out_node = (int32_t)n_sub; sub[n_sub++] = s; non_full[n_nf++] = out_node;
            }
            if (in_node < 0) {
                subnode_t s; s.x = ix; s.y = iy; s.e[0] = out_node; s.e[1] = -1; s.walked = 0; s.index = in_idx;
                in_node = (int32_t)n_sub; sub[n_sub++] = s; non_full[n_nf++] = in_node;
                sub[out_node].e[found_out ? 1 : 0] = in_node;
            } else if (found_out) sub[out_node].e[1] = in_node;
        }
    }
    for (int64_t i = 0; i < n_nf; ++i) sides[n_sides++] = non_full[i];
    /* walk (:1034-1130): deque with push_front / pop_front */
    int32_t *dq = (int32_t *)malloc(sizeof(int32_t) * (size_t)(cap_sub + 2));
    int64_t n_loops = 0, n_pts = 0, ret = 0;
    loop_off[0] = 0;
    for (int64_t si = 0; si < n_sides && ret == 0; ++si) {
        if (sub[sides[si]].walked) continue;
        if (n_loops >= cap_loops) { ret = -1; break; }
        int64_t top = 0;                                      /* dq[top-1] is the front */
        dq[top++] = sides[si]; sub[sides[si]].walked = 1;
        while (top > 0) {
            const int32_t nd = dq[--top];
            if (n_pts >= cap_pts) { ret = -1; break; }
            pts[2 * n_pts] = sub[nd].x; pts[2 * n_pts + 1] = sub[nd].y; ++n_pts;
            for (int e = 0; e < 2; ++e) {
                const int32_t t = sub[nd].e[e];
                if (t < 0 || sub[t].walked) continue;
                sub[t].walked = 1;
                dq[top++] = t;
            }
        }
        loop_off[++n_loops] = n_pts;
    }
    free(lines); free(nodes); free(sub); free(non_full); free(sides); free(dq);
    return ret ? ret : n_loops;
}

/* Outline::resample (T/tracking/Outline.cpp:724-766), float arithmetic as written there (Float2_t = float, the step
 * fraction offset * 1.0 / percent in double, then narrowed by Vector2D::operator*).  Returns the number of points. */
int64_t to_outline_resample(const float *pts, int64_t L, float rd, float *out, int64_t cap)
{
    if (rd <= 0 || L <= 1) { if (L > cap) return -1; memcpy(out, pts, sizeof(float) * 2 * (size_t)L); return L; }
    float walked = 0.0f;
    int64_t n = 0;
    for (int64_t i = 0; i < L; ++i) {
        const int64_t i1 = i + 1 >= L ? i + 1 - L : i + 1;
        const float x0 = pts[2 * i], y0 = pts[2 * i + 1];
        const float lx = pts[2 * i1] - x0, ly = pts[2 * i1 + 1] - y0;
        const float len = sqrtf(lx * lx + ly * ly);
        walked += len;
        const float percent = len / rd;
        float walked_percent = walked / rd;
        int offset = 0;
        while (walked_percent >= 1.0) {
            const float f = (float)(offset * 1.0 / percent);
            if (n >= cap) return -1;
            out[2 * n] = x0 + lx * f; out[2 * n + 1] = y0 + ly * f; ++n;
            offset++;
            walked -= rd;
            walked_percent = (float)(walked_percent - 1.0);
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * "Next" row N4, second stage (ORACLE ONLY so far -- no CUDA counterpart yet): from the resampled outline to
 * the raw midline.
 *   Outline::smooth / smooth_outline         T/tracking/Outline.cpp:330-452
 *   Outline::offset_to_middle                T/tracking/Outline.cpp:454-718   (the non-"old" method)
 *   periodic::differentiate(_and_test_clockwise), curvature, find_peaks, eft, ieft, fast::cos
 *                                            C/misc/CircularGraph.cpp:12-30,49-113,115-407,409-482,484-606
 *   Outline::calculate_midline               T/tracking/Outline.cpp:768-868
 * Float2_t = scalar_t = float; every double literal / M_PI in the reference promotes exactly as written below.
 * Pinned on the reference's own code: CircularGraph.cpp and Outline.cpp are compiled unmodified by oracle/build_ref.py, and
 * tests/test_oracle_ref_circular_graph.py / tests/test_oracle_ref_outline.py hold these functions to them bit for bit.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t outline_smooth_samples;        /* 4   T/core/default_config.cpp:890 */
    int32_t outline_smooth_step;           /* 1   :889 */
    int32_t outline_approximate;           /* 3   :888 */
    float   outline_curvature_range_ratio; /* 0.03 :891 */
    float   midline_walk_offset;           /* 0.025 :892 */
    int32_t peak_mode;                     /* 0 pointy (default, :902), 1 broad */
    int32_t midline_start_with_head;       /* 0   :900 */
    int32_t midline_invert;                /* 0   :901 */
    int32_t midline_resolution;            /* 25  :894 */
    float   midline_stiff_percentage;      /* 0.15 :893 */
} to_posture_params_t;

void to_posture_default_params(to_posture_params_t *p)
{
    p->outline_smooth_samples = 4; p->outline_smooth_step = 1; p->outline_approximate = 3;
    p->outline_curvature_range_ratio = 0.03f; p->midline_walk_offset = 0.025f;
    p->peak_mode = 0; p->midline_start_with_head = 0; p->midline_invert = 0;
    p->midline_resolution = 25; p->midline_stiff_percentage = 0.15f;
}

/* fast::cos<float> / fast::sin<float>, CircularGraph.cpp:12-30 */
static float fast_cosf(float x)
{
    const float tp = (float)(1. / (2. * 3.14159265358979323846264338327950288));
    x *= tp;
    x -= 0.25f + floorf(x + 0.25f);
    x *= 16.f * (fabsf(x) - 0.5f);
    x += 0.225f * x * (fabsf(x) - 1.f);
    return x;
}
static float fast_sinf(float x) { return fast_cosf(x - (float)1.57079632679489661923132169163975144); }
float to_fast_cos(float x) { return fast_cosf(x); }

/* smooth_outline (:330-378) as Outline::smooth calls it (:380-389).  Returns 1 when the points were replaced. */
int to_outline_smooth(const float *pts, int64_t L, int samples, int step_i, float *out)
{
    const float range = (float)samples;
    if (!((float)L > range)) return 0;
    const long step = step_i;
    const float step_row = range * (float)step;
    float *weights = (float *)malloc(sizeof(float) * (size_t)(2 * (int64_t)step_row + 4));
    int nw = 0;
    float sum = 0;
    for (int i = (int)(-step_row); (float)i <= step_row; i += (int)step) {
        const float val = (step_row - (float)abs(i)) / step_row;
        sum += val;
        weights[nw++] = val;
    }
    for (int i = 0; i < nw; ++i) weights[i] /= sum;
    for (long i = 0; i < L; ++i) {
        long samples_n = 0;
        float px = 0, py = 0;
        for (long j = (long)((float)i - step_row); (float)j <= (float)i + step_row; j += step) {
            long idx = j;
            while (idx < 0) idx += L;
            while (idx >= L) idx -= L;
            const float wgt = weights[samples_n++];
            px += pts[2 * idx] * wgt; py += pts[2 * idx + 1] * wgt;
        }
        out[2 * i] = px; out[2 * i + 1] = py;
    }
    free(weights);
    return 1;
}

/* periodic::curvature (CircularGraph.cpp:49-113); out must be zero-initialised by the caller (std::vector::resize). */
void to_periodic_curvature(const float *p, int64_t N, int r, int absolute, float *out)
{
    for (int64_t i = 0; i < N; ++i) {
        const int64_t i1 = ((i - r) % N + N) % N, i3 = (i + r) % N;
        const float x1 = p[2 * i1], y1 = p[2 * i1 + 1], x2 = p[2 * i], y2 = p[2 * i + 1], x3 = p[2 * i3], y3 = p[2 * i3 + 1];
        const int e12 = x1 == x2 && y1 == y2, e13 = x1 == x3 && y1 == y3, e23 = x2 == x3 && y2 == y3;
        if (!e12 && !e13 && !e23) {
            const float cross = (x2 - x1) * (y3 - y2) - (y2 - y1) * (x3 - x2);
            const float d12 = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1), d23 = (x3 - x2) * (x3 - x2) + (y3 - y2) * (y3 - y2),
                        d13 = (x3 - x1) * (x3 - x1) + (y3 - y1) * (y3 - y1);
            out[i] = 2.f * (absolute ? fabsf(cross) : cross) / sqrtf(d12 * d23 * d13);
        } else out[i] = 0.f;
    }
}

/* _differentiate<true> on points (:409-463): dxy[n] = a[n+1] - a[n] (cyclic) and the orientation sum as the reference
 * accumulates it (the wrap-around term enters with the operands swapped, :458-459). */
static float differentiate_points(const float *p, int64_t N, float *dxy)
{
    float sum = 0;
    for (int64_t i = 1; i < N; ++i) {
        dxy[2 * (i - 1)] = p[2 * i] - p[2 * (i - 1)]; dxy[2 * (i - 1) + 1] = p[2 * i + 1] - p[2 * (i - 1) + 1];
        sum += p[2 * (i - 1)] * p[2 * i + 1] - p[2 * i] * p[2 * (i - 1) + 1];
    }
    dxy[2 * (N - 1)] = p[0] - p[2 * (N - 1)]; dxy[2 * (N - 1) + 1] = p[1] - p[2 * (N - 1) + 1];
    sum += p[0] * p[2 * (N - 1) + 1] - p[2 * (N - 1)] * p[1];
    return sum;
}
float to_orientation_sum(const float *p, int64_t N)
{
    float *d = (float *)malloc(sizeof(float) * 2 * (size_t)N);
    const float s = differentiate_points(p, N, d);
    free(d);
    return s;
}

/* periodic::eft (:484-561) with EFT::dt (:466-482); coeffs: order x {a, b, c, d}. */
void to_eft(const float *p, int64_t N, int order, float *coeffs)
{
    float *dxy = (float *)malloc(sizeof(float) * 2 * (size_t)N);
    differentiate_points(p, N, dxy);
    const int64_t nd = N - 1;                               /* dt has dxy.size() - 1 entries */
    float *dt = (float *)malloc(sizeof(float) * (size_t)(nd + 1)), *phi = (float *)calloc((size_t)(nd + 2), sizeof(float));
    float *cx = (float *)malloc(sizeof(float) * (size_t)(nd + 1)), *cy = (float *)malloc(sizeof(float) * (size_t)(nd + 1));
    float *cum = (float *)calloc((size_t)(nd + 2), sizeof(float));
    float sum = 0;
    for (int64_t i = 0; i < nd; ++i) {
        const float x = dxy[2 * i], y = dxy[2 * i + 1];
        dt[i] = (float)((double)sqrtf(x * x + y * y) + 1e-10);
        sum += dt[i];
        cum[i + 1] = sum;
        phi[i + 1] = (float)(2 * 3.14159265358979323846 * (double)cum[i + 1]);
        cx[i] = x / dt[i]; cy[i] = y / dt[i];
    }
    const float T = cum[nd];
    const float norm_base = (float)((double)T / (2 * (3.14159265358979323846 * 3.14159265358979323846)));
    const int64_t np = nd + 1;                              /* phi.size() */
    float *cs = (float *)malloc(sizeof(float) * 2 * (size_t)np);
    for (int n = 1; n < order + 1; ++n) {
        const float norm = norm_base / (float)((size_t)n * (size_t)n);
        float cnx = 0, cny = 0, snx = 0, sny = 0;
        for (int64_t i = 0; i < np; ++i) {
            const float phi_n = phi[i] * (float)n / T;
            cs[2 * i] = fast_cosf(phi_n); cs[2 * i + 1] = fast_sinf(phi_n);
        }
        for (int64_t i = 0; i < np - 1; ++i) {
            const float dc = cs[2 * (i + 1)] - cs[2 * i], ds = cs[2 * (i + 1) + 1] - cs[2 * i + 1];
            cnx += cx[i] * dc; cny += cy[i] * dc;
            snx += cx[i] * ds; sny += cy[i] * ds;
        }
        cnx *= norm; cny *= norm; snx *= norm; sny *= norm;
        coeffs[4 * (n - 1) + 0] = cnx; coeffs[4 * (n - 1) + 1] = snx; coeffs[4 * (n - 1) + 2] = cny; coeffs[4 * (n - 1) + 3] = sny;
    }
    free(dxy); free(dt); free(phi); free(cx); free(cy); free(cum); free(cs);
}

/* periodic::ieft (:563-606), save_steps = false, scale = 1 */
void to_ieft(const float *coeffs, int order, int64_t n_points, float offx, float offy, float *out)
{
    for (int64_t j = 0; j < n_points; ++j) { out[2 * j] = offx; out[2 * j + 1] = offy; }
    for (int i = 0; i < order; ++i)
        for (int64_t j = 0; j < n_points; ++j) {
            const float t = (float)((double)j / (double)(n_points - 1) * 3.14159265358979323846 * 2.0);
            const float ct = fast_cosf(t * (float)(i + 1)), st = fast_sinf(t * (float)(i + 1));
            out[2 * j] += 1.f * (coeffs[4 * i] * ct + coeffs[4 * i + 1] * st);
            out[2 * j + 1] += 1.f * (coeffs[4 * i + 2] * ct + coeffs[4 * i + 3] * st);
        }
}

/* periodic::find_peaks (:115-407) for the call in offset_to_middle: diffs = {first difference, first difference}
 * (differentiate(curv, 2) writes the same difference into both outputs, :441-444). */
typedef struct { float x, y, width, integral, r0, r1, max_y_extrema, max_y; int n_pts; } to_peak_t;
typedef struct { float i; int is_max; } extremum_t;

int64_t to_find_peaks(const float *pt, int64_t N, int broad, to_peak_t *maxima, int64_t cap)
{
    if (N <= 0) return 0;
    float *diff = (float *)malloc(sizeof(float) * (size_t)N);
    for (int64_t i = 0; i + 1 < N; ++i) diff[i] = pt[i + 1] - pt[i];
    diff[N - 1] = pt[0] - pt[N - 1];
    const float *second = diff;
    extremum_t *ext = (extremum_t *)malloc(sizeof(extremum_t) * (size_t)N);
    int64_t n_ext = 0, n_max = 0;
    int sign = diff[N - 1] < 0;
    float minimum = 3.402823466e+38f;
    for (int64_t i = 0; i < N; ++i) {
        const int c = diff[i] < 0;
        if (pt[i] < minimum) minimum = pt[i];
        if (c != sign) {
            if (second[i] != 0) {
                if (!sign) {
                    if (n_max >= cap) { free(diff); free(ext); return -1; }
                    to_peak_t pk; memset(&pk, 0, sizeof pk);
                    pk.x = (float)i; pk.y = pt[i]; pk.r0 = -1; pk.r1 = -1;
                    maxima[n_max++] = pk;
                }
                ext[n_ext].i = (float)i; ext[n_ext].is_max = !sign; ++n_ext;        /* std::set ordered by i: inserted ascending */
            }
            sign = c;
        }
    }
    const float min_y = minimum;
    /* maxima in descending (y, index) order: std::set<tuple<y, idx>, greater<>> */
    int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_max + 1));
    for (int64_t i = 0; i < n_max; ++i) order[i] = i;
    for (int64_t i = 1; i < n_max; ++i) {
        const int64_t k = order[i]; int64_t j = i - 1;
        while (j >= 0 && (maxima[order[j]].y < maxima[k].y || (maxima[order[j]].y == maxima[k].y && order[j] < k))) { order[j + 1] = order[j]; --j; }
        order[j + 1] = k;
    }
    float *rng = (float *)malloc(sizeof(float) * 2 * (size_t)(n_max + 1));
    int64_t n_rng = 0;
    for (int64_t oi = 0; oi < n_max; ++oi) {
        to_peak_t *peak = &maxima[order[oi]];
        int64_t after = 0, prev = n_ext - 1;
        for (; after != n_ext; ++after) {
            if (ext[after].i == peak->x) { ++after; if (after == n_ext) after = 0; break; }
            prev = after;
        }
        if (after == n_ext) after = 0;       /* cannot happen: the peak is in extrema */
        float minimum_left = 3.402823466e+38f, index_left = ext[prev].i, minimum_right = 3.402823466e+38f, index_right = ext[after].i;
        float left_border = 0, right_border = (float)N;
        for (int64_t k = 0; k < n_rng; ++k) {
            const float rs = rng[2 * k], re = rng[2 * k + 1];
            if (re > left_border && re < peak->x) left_border = re;
            if (rs < right_border && rs > peak->x) right_border = rs;
        }
        int64_t cl = prev;
        float last_y = peak->y, offset = 0;
        while (ext[cl].i + offset >= left_border) {
            const float x = ext[cl].i, y = pt[(size_t)x];
            if (ext[cl].is_max) { if ((double)y > (double)last_y * 1.05) break; last_y = y; }
            if (y < minimum_left) { minimum_left = y; index_left = x; }
            if (y > peak->max_y_extrema) peak->max_y_extrema = y;
            if (cl == 0) { offset = -(float)N; cl = n_ext - 1; }
            if (cl == after) break;
            --cl;
            if (cl < 0) break;                /* decrementing begin() is undefined in the reference; only reachable with one extremum */
        }
        offset = 0; cl = after; last_y = peak->y;
        while (ext[cl].i + offset <= right_border) {
            const float x = ext[cl].i, y = pt[(size_t)x];
            if (ext[cl].is_max) { if ((double)y > (double)last_y * 1.05) break; last_y = y; }
            if (y < minimum_right) { minimum_right = y; index_right = x; }
            if (y > peak->max_y_extrema) peak->max_y_extrema = y;
            if (++cl == n_ext) { offset = (float)N; cl = 0; }
            if (cl == prev) break;
        }
        while (index_left > peak->x) index_left -= (float)N;
        while (index_right < peak->x) index_right += (float)N;
        peak->width = right_border - left_border;
        peak->r0 = index_left; peak->r1 = index_right;
        rng[2 * n_rng] = index_left; rng[2 * n_rng + 1] = index_right; ++n_rng;
    }
    /* points above half height inside the (periodically split) range, then the integral (:336-402) */
    for (int64_t k = 0; k < n_max; ++k) {
        to_peak_t *peak = &maxima[k];
        float chk[3][2]; int nchk = 1;
        chk[0][0] = peak->r0; chk[0][1] = peak->r1;
        float x0 = peak->r0, x1 = peak->r1;
        if (x0 < 0) { x0 += (float)N; chk[0][0] = 0; chk[nchk][0] = x0; chk[nchk][1] = (float)(N - 1); ++nchk; }
        if (x1 >= (float)N) { x1 -= (float)N; chk[0][1] = (float)(N - 1); chk[nchk][0] = 0; chk[nchk][1] = x1; ++nchk; }
        /* first pass: max_y and count; second pass: the integral needs max_y of ALL points first */
        float max_y = 0; int n_pts = 0;
        for (int c = 0; c < nchk; ++c) {
            const float first = chk[c][0], last = chk[c][1];
            if (!(last - first >= 0)) continue;
            const size_t steps = (size_t)((last - first) / 1.f);
            for (size_t s = 0; s < steps; ++s) {
                const float i = first + (float)s * 1.f;
                const float y = pt[(size_t)i] - min_y;
                if ((double)(y / (peak->y - min_y)) >= 0.5) { ++n_pts; if (y > max_y) max_y = y; }
            }
        }
        peak->max_y = max_y; peak->n_pts = n_pts;
        const double median = (double)max_y * 0.5;
        const float factor = broad ? 1.f : peak->y;
        float integral = 0;
        for (int c = 0; c < nchk; ++c) {
            const float first = chk[c][0], last = chk[c][1];
            if (!(last - first >= 0)) continue;
            const size_t steps = (size_t)((last - first) / 1.f);
            for (size_t s = 0; s < steps; ++s) {
                const float i = first + (float)s * 1.f;
                const float y = pt[(size_t)i] - min_y;
                if ((double)(y / (peak->y - min_y)) >= 0.5) integral = (float)((double)integral + ((double)y - median) * (double)factor);
            }
        }
        peak->integral = integral;
    }
    free(diff); free(ext); free(order); free(rng);
    return n_max;
}

static void rotate_points(float *p, int64_t N, int64_t k)      /* std::rotate(begin, begin + k, end) */
{
    if (N <= 0 || k <= 0 || k >= N) return;
    float *t = (float *)malloc(sizeof(float) * 2 * (size_t)N);
    memcpy(t, p + 2 * k, sizeof(float) * 2 * (size_t)(N - k));
    memcpy(t + 2 * (N - k), p, sizeof(float) * 2 * (size_t)k);
    memcpy(p, t, sizeof(float) * 2 * (size_t)N);
    free(t);
}

/* is_in_periodic_range (T/tracking/Outline.cpp:442-452) */
static int in_periodic_range(float period, float r0, float r1, float x, float *cx)
{
    if (r0 < 0) {
        if (x - period >= r0) { *cx = x - period; return 1; }
        *cx = x; return x <= r1;
    } else if (r1 >= period) {
        if (x + period <= r1) { *cx = x + period; return 1; }
        *cx = x; return x >= r0;
    }
    *cx = x; return x >= r0 && x < r1;           /* Range::contains is half open (C/misc/ranges.h:163-168) */
}

/* The broad-tail index of offset_to_middle (:621-650): the peaks whose integral equals the largest one are merged, and the tail is
 * the middle of the span of their points that reach 90 % of the highest one.  (The reference collects the points in a std::set
 * first; only their extent is used, so the order and duplicates do not matter.) */
static float broad_tail_index(const float *curv, int64_t N, const to_peak_t *mx, int64_t nm)
{
    float max_int = -1, minimum = 3.402823466e+38f;
    for (int64_t i = 0; i < N; ++i) if (curv[i] < minimum) minimum = curv[i];
    for (int64_t k = 0; k < nm; ++k) if (mx[k].integral > max_int) max_int = mx[k].integral;
    float m0 = 0, m1 = 0, max_y = 0; int first = 1;
    for (int64_t k = 0; k < nm; ++k) {
        if (!((double)fabsf(mx[k].integral - max_int) <= 1e-5)) continue;
        if (first) { m0 = mx[k].r0; m1 = mx[k].r1; first = 0; }
        else { if (mx[k].r0 < m0) m0 = mx[k].r0; if (mx[k].r1 > m1) m1 = mx[k].r1; }
        if (mx[k].max_y > max_y) max_y = mx[k].max_y;
    }
    float start = m1, end = m0;
    for (int64_t k = 0; k < nm; ++k) {
        if (!((double)fabsf(mx[k].integral - max_int) <= 1e-5)) continue;
        float chk[3][2]; int nchk = 1;
        chk[0][0] = mx[k].r0; chk[0][1] = mx[k].r1;
        float x0 = mx[k].r0, x1 = mx[k].r1;
        if (x0 < 0) { x0 += (float)N; chk[0][0] = 0; chk[nchk][0] = x0; chk[nchk][1] = (float)(N - 1); ++nchk; }
        if (x1 >= (float)N) { x1 -= (float)N; chk[0][1] = (float)(N - 1); chk[nchk][0] = 0; chk[nchk][1] = x1; ++nchk; }
        for (int c = 0; c < nchk; ++c) {
            if (!(chk[c][1] - chk[c][0] >= 0)) continue;
            const size_t steps = (size_t)(chk[c][1] - chk[c][0]);
            for (size_t s2 = 0; s2 < steps; ++s2) {
                const float i = chk[c][0] + (float)s2;
                const float y = curv[(size_t)i] - minimum;
                if (!((double)(y / (mx[k].y - minimum)) >= 0.5)) continue;          /* a point of this peak (:373-389) */
                float cx;
                const int in_range = in_periodic_range((float)(size_t)N, m0, m1, i, &cx);
                if ((double)y >= (double)max_y * 0.9 && in_range) { if (start > cx) start = cx; if (end < cx) end = cx; }
            }
        }
    }
    float idx = (float)round((double)start + (double)(end - start) * 0.5);
    if (idx < 0) idx += (float)(size_t)N;
    if (idx >= (float)(size_t)N) idx -= (float)(size_t)N;
    return idx;
}

/* Outline::offset_to_middle (T/tracking/Outline.cpp:454-718), peak_mode pointy (the default) or broad.  Works in place on the N points (they may be reversed, replaced by their elliptic-Fourier approximation and
 * rotated); curv_out (optional, N floats) receives the curvature the peaks were searched in.
 * Returns 0 and tail / head indices, or a negative code: -1 empty. */
int to_offset_to_middle(float *p, int64_t N, const to_posture_params_t *P, int64_t *tail_out, int64_t *head_out, float *curv_out)
{
    if (N <= 0) return -1;
    if (to_orientation_sum(p, N) < 0)                         /* make it clockwise (:496-498) */
        for (int64_t i = 0, j = N - 1; i < j; ++i, --j) {
            float tx = p[2 * i], ty = p[2 * i + 1];
            p[2 * i] = p[2 * j]; p[2 * i + 1] = p[2 * j + 1]; p[2 * j] = tx; p[2 * j + 1] = ty;
        }
    if (P->outline_approximate > 0) {
        float cx = 0, cy = 0;
        for (int64_t i = 0; i < N; ++i) { cx += p[2 * i]; cy += p[2 * i + 1]; }
        cx /= (float)(size_t)N; cy /= (float)(size_t)N;
        float *coeffs = (float *)malloc(sizeof(float) * 4 * (size_t)P->outline_approximate);
        to_eft(p, N, P->outline_approximate, coeffs);
        to_ieft(coeffs, P->outline_approximate, N, cx, cy, p);
        free(coeffs);
    }
    float rf = P->outline_curvature_range_ratio * (float)(size_t)N;          /* max(1, ratio * size) as int (:514) */
    if (rf < 1.f) rf = 1.f;
    const int r = (int)rf;
    float *curv = (float *)calloc((size_t)N, sizeof(float));
    to_periodic_curvature(p, N, r, P->outline_approximate > 0, curv);
    if (curv_out) memcpy(curv_out, curv, sizeof(float) * (size_t)N);
    to_peak_t *mx = (to_peak_t *)malloc(sizeof(to_peak_t) * (size_t)(N + 1));
    const int64_t nm = to_find_peaks(curv, N, P->peak_mode != 0, mx, N + 1);
    float max_y = -1, max_y_idx = 0;
    for (int64_t k = 0; k < nm; ++k)
        if (mx[k].y > max_y) { max_y = mx[k].y; max_y_idx = mx[k].x; }
    const float idx = P->peak_mode == 0 ? max_y_idx           /* FIND_POINTY (:617-619) */
                                        : broad_tail_index(curv, N, mx, nm);
    int64_t tail = (int64_t)idx, head = -1;
    float max_d = 0;
    for (int64_t k = 0; k < nm; ++k) {
        float d;
        const float px = mx[k].x, sz = (float)(size_t)N;
        if (px >= idx) { const float a = fabsf(px - idx), b = fabsf(px - idx - sz); d = a < b ? a : b; }
        else { const float a = fabsf(idx - px), b = fabsf(idx - px - sz); d = a < b ? a : b; }
        if (d > max_d) { max_d = d; head = (int64_t)px; }
    }
    if (P->midline_start_with_head && head != -1) {
        if (tail != -1) { tail -= head; if (tail < 0) tail += N; }
        rotate_points(p, N, head);
        head = 0;
    } else {
        if (head != -1) { head -= tail; if (head < 0) head += N; }
        rotate_points(p, N, tail);
        tail = 0;
    }
    if (P->midline_invert) { const int64_t t = tail; tail = head; head = t; }
    *tail_out = tail; *head_out = head;
    free(curv); free(mx);
    return 0;
}

/* Outline::calculate_midline (T/tracking/Outline.cpp:768-868) on a resampled outline: smooth -> offset_to_middle -> the pairing
 * walk.  p (N points) is modified in place; segments: {pos.x, pos.y, height, l_length} per midline segment.
 * Returns the number of segments (> 2), or a negative code: -1 empty, -2 too few segments, -3 broad mode, -4 capacity. */
int64_t to_calculate_midline(float *p, int64_t N, const to_posture_params_t *P, float *segments, int64_t cap,
                             int64_t *tail_out, int64_t *head_out)
{
    if (N <= 0) return -1;
    if (P->outline_smooth_samples > 0) {
        float *sm = (float *)malloc(sizeof(float) * 2 * (size_t)N);
        if (to_outline_smooth(p, N, P->outline_smooth_samples, P->outline_smooth_step, sm)) memcpy(p, sm, sizeof(float) * 2 * (size_t)N);
        free(sm);
    }
    int rc = to_offset_to_middle(p, N, P, tail_out, head_out, NULL);
    if (rc) return rc;
    if (N <= 1) return -1;
    const long L = (long)N;
    int idx_r = 1, idx_l = -1;
    float mo = P->midline_walk_offset * (float)L;
    if (mo < 3.f) mo = 3.f;
    const int max_offset = (int)mo;
    int64_t ns = 0;
    while (idx_r < L + idx_l) {
        float prx = 0, pry = 0, plx = p[2 * (L + idx_l)], ply = p[2 * (L + idx_l) + 1];
        float min_d = 3.402823466e+38f; int min_idx = -1;
        for (int i = 0; i < max_offset; ++i) {
            if (idx_r + i >= L) break;
            const float dx = p[2 * (idx_r + i)] - plx, dy = p[2 * (idx_r + i) + 1] - ply;
            const float len = sqrtf(dx * dx + dy * dy);
            if (len < min_d) { min_d = len; min_idx = idx_r + i; }
        }
        if (min_idx != -1) { prx = p[2 * min_idx]; pry = p[2 * min_idx + 1]; idx_r = min_idx; }
        min_d = 3.402823466e+38f; min_idx = 1;
        for (int i = 0; i < max_offset; ++i) {
            if (idx_l - i <= -L) break;
            const float dx = prx - p[2 * (L + idx_l - i)], dy = pry - p[2 * (L + idx_l - i) + 1];
            const float len = sqrtf(dx * dx + dy * dy);
            if (len < min_d) { min_d = len; min_idx = idx_l - i; }
        }
        if (min_idx != 1) { plx = p[2 * (L + min_idx)]; ply = p[2 * (L + min_idx) + 1]; idx_l = min_idx; }
        const float lx = prx - plx, ly = pry - ply;
        const float mx_ = plx + lx * 0.5f, my_ = ply + ly * 0.5f;
        if (ns >= cap) return -4;
        segments[4 * ns + 0] = mx_; segments[4 * ns + 1] = my_;
        segments[4 * ns + 2] = sqrtf((prx - plx) * (prx - plx) + (pry - ply) * (pry - ply));
        segments[4 * ns + 3] = sqrtf((plx - mx_) * (plx - mx_) + (ply - my_) * (ply - my_));
        ++ns;
        idx_r++; idx_l--;
    }
    if (ns <= 2) return -2;
    return ns;
}

/* ------------------------------------------------------------------------------------------
 * "Next" row N4, third stage: Midline::post_process / normalize / fix_length / calculate_angle / transform
 * (T/tracking/Outline.cpp:870-1085, 1113-1456) and the `posture` / `legacy` crop normalisation built on it
 * (T/tracking/FilterCache.cpp:21-115,133-154,266-276; call site T/tracking/ImageExtractor.cpp:244-270:
 * Individual::calculate_midline_for = post_process + normalize(), then Midline::transform(type)).
 * Segments are {pos.x, pos.y, height, l_length}.  Float2_t = float; the double promotions are the reference's.
 * Third-party libm calls: the reference calls ::atan2f (cmn::atan2, C/misc/math.h:168-170), cos / sin / acos.  They are
 * restated as the double-precision function rounded to float ((float)atan2((double)y, (double)x)): that is the correctly
 * rounded float result except for inputs within 2^-29 of a rounding boundary, it is what glibc >= 2.41's atan2f returns,
 * and it is within 1 ulp of older atan2f versions (tests/test_oracle_posture.py measures the agreement with the local libm).
 * Pinned on the reference's own code (Outline.cpp and gui/Transform.cpp compiled unmodified, oracle/build_ref.py): post_process is bit-exact,
 * normalize is bit-exact when this file calls the local libm like the compiled reference does (to_use_local_libm) and within 3e-5 px with
 * the rounding above, the crop matrix is bit-exact in double (tests/test_oracle_ref_outline.py); the normalised length is also corroborated
 * against the midline_length column TRex itself exported (tests/golden/posture_golden.npz).
 * ------------------------------------------------------------------------------------------ */
/* to_use_local_libm(1): call the C library's own atan2f / cosf / sinf instead, as the reference binary does on this machine -- only for
 * tests/test_oracle_ref_outline.py, which compares with the reference's code compiled here (oracle/build_ref.py) bit for bit. */
static int g_local_libm = 0;
void to_use_local_libm(int on) { g_local_libm = on; }
static float atan2_f(float y, float x) { return g_local_libm ? atan2f(y, x) : (float)atan2((double)y, (double)x); }
static float cos_f(float a) { return g_local_libm ? cosf(a) : (float)cos((double)a); }
static float sin_f(float a) { return g_local_libm ? sinf(a) : (float)sin((double)a); }

static void vnormalize(float x, float y, float *ox, float *oy)      /* Vector2D::normalize, C/misc/vec2.h:159-162 */
{
    const float L = sqrtf(x * x + y * y);
    const float s = (float)(L != 0), d = (float)(L == 0) + L;
    *ox = s * (x / d); *oy = s * (y / d);
}

/* Midline::midline_direction (:870-887) */
static void midline_direction(const float *seg, int64_t n, float stiff, float *dx_out, float *dy_out)
{
    const int32_t samples = (int32_t)fmaxf(1.f, (float)(size_t)n * stiff);       /* cmn::max(int, float) -> float */
    float dx = 0, dy = 0;
    int32_t counted = 0;
    for (int32_t i = 0; i < samples && i + 1 < (int32_t)n; i++, counted++) {
        dx += seg[4 * (i + 1)] - seg[4 * i]; dy += seg[4 * (i + 1) + 1] - seg[4 * i + 1];
    }
    if (counted > 0) {
        dx /= (float)counted; dy /= (float)counted;
        vnormalize(dx, dy, &dx, &dy);
    }
    *dx_out = dx; *dy_out = dy;
}

static void reverse_segments(float *seg, int64_t n)
{
    for (int64_t i = 0, j = n - 1; i < j; ++i, --j)
        for (int k = 0; k < 4; ++k) { const float t = seg[4 * i + k]; seg[4 * i + k] = seg[4 * j + k]; seg[4 * j + k] = t; }
}

/* Midline::post_process (:895-1062).  seg: n segments, modified in place; move_dir: MovementInformation::direction ((0,0): none,
 * the default -- posture_direction_smoothing = 0, T/tracking/Individual.cpp:1365-1368); tail / head are swapped when the
 * movement says so.  Returns 0, 1 when the midline was inverted because of the movement direction, or -5 where the reference
 * throws (segments().at(i + 1) past the end in the axis loop, :981-984). */
int to_midline_post_process(float *seg, int64_t n, const to_posture_params_t *P, const float move_dir[2], int64_t *tail, int64_t *head)
{
    if (n <= 2) return 0;
    float dx, dy;
    midline_direction(seg, n, P->midline_stiff_percentage, &dx, &dy);
    int needs_invert = !P->midline_invert, inverted_prev = 0;
    if (!needs_invert) { dx = -dx; dy = -dy; }
    if (move_dir && (move_dir[0] != 0 || move_dir[1] != 0)) {
        const float a = (-dx) * move_dir[0] + (-dy) * move_dir[1], b = dx * move_dir[0] + dy * move_dir[1];
        if (acos((double)a) < acos((double)b)) {
            needs_invert = !needs_invert; inverted_prev = 1;
            if (tail && head) { const int64_t t = *tail; *tail = *head; *head = t; }
        }
    }
    if (needs_invert) { if (!P->midline_start_with_head) reverse_segments(seg, n); }
    else if (P->midline_start_with_head) reverse_segments(seg, n);
    const float stiff = P->midline_stiff_percentage;
    if (stiff > 0) {
        const size_t center = (size_t)fminf((float)(size_t)n - 1, roundf((float)(size_t)n * stiff) + 1);
        const float cpx = seg[4 * center], cpy = seg[4 * center + 1];
        float ax = 0, ay = 0;
        uint32_t count = 0;
        const size_t extra = (size_t)fmin((double)(int)n, (double)center + fmax(0.0, (double)(size_t)n * 0.1));
        for (size_t i = center; i < extra; ++i) {
            if (i + 1 >= (size_t)n) return -5;                      /* segments().at(i + 1) throws std::out_of_range */
            float nx, ny;
            vnormalize(seg[4 * i] - seg[4 * (i + 1)], seg[4 * i + 1] - seg[4 * (i + 1) + 1], &nx, &ny);
            ax += nx; ay += ny; ++count;
        }
        if (count > 0) { ax /= (float)count; ay /= (float)count; }
        float *copy = (float *)malloc(sizeof(float) * 4 * (size_t)n);
        memcpy(copy, seg, sizeof(float) * 4 * (size_t)n);
        for (size_t i = center; i > 0; --i) {
            const float p1x = seg[4 * i], p1y = seg[4 * i + 1];
            const float lx = copy[4 * i] - copy[4 * (i - 1)], ly = copy[4 * i + 1] - copy[4 * (i - 1) + 1];
            const float L = sqrtf(lx * lx + ly * ly);
            float dcx, dcy, tx, ty;
            vnormalize(seg[4 * (i - 1)] - cpx, seg[4 * (i - 1) + 1] - cpy, &dcx, &dcy);
            vnormalize((dcx + ax) * 0.5f, (dcy + ay) * 0.5f, &tx, &ty);
            seg[4 * (i - 1)] = p1x + L * tx; seg[4 * (i - 1) + 1] = p1y + L * ty;
        }
        free(copy);
    }
    reverse_segments(seg, n);
    return inverted_prev;
}

/* Midline::calculate_angle (:1113-1123) */
static float midline_calculate_angle(const float *seg, int64_t n, float stiff)
{
    if (n < 2) return 0;
    const float center = fmaxf(0.f, (float)((size_t)n - 2) - (float)(size_t)n * stiff);
    const size_t start = (size_t)center;
    const float rest = center - (float)start;
    const float lx = seg[4 * (n - 1)] - (seg[4 * start] * (1 - rest) + seg[4 * (start + 1)] * rest);
    const float ly = seg[4 * (n - 1) + 1] - (seg[4 * start + 1] * (1 - rest) + seg[4 * (start + 1) + 1] * rest);
    return atan2_f(ly, lx);
}

/* t_circle_line (C/misc/math.h:287-310) */
static void t_circle_line(float x0, float y0, float x1, float y1, float h, float k, float r, float *t0, float *t1)
{
    const float a = (x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0);
    const float b = 2 * (x1 - x0) * (x0 - h) + 2 * (y1 - y0) * (y0 - k);
    const float c = (x0 - h) * (x0 - h) + (y0 - k) * (y0 - k) - r * r;
    float disc = b * b - 4 * a * c;
    if (disc < 0) { *t0 = -1; *t1 = -1; return; }
    disc = sqrtf(disc);
    *t0 = (-b + disc) / (2 * a); *t1 = (-b - disc) / (2 * a);
}

/* Midline::fix_length (:1125-1236) on `resolution` points; pts is replaced by the re-spaced points (count returned). */
static int64_t midline_fix_length(float len, float *pts, int64_t n_pts, uint32_t resolution)
{
    const float step = len / (float)resolution;
    float *out = (float *)malloc(sizeof(float) * 4 * (size_t)(resolution + 2));
    int64_t n_out = 0;
    float seg[4]; memcpy(seg, pts, sizeof seg);
    memcpy(out + 4 * n_out++, seg, sizeof seg);
    uint32_t j = 1;
    float last_t = -1;
    for (uint32_t i = 1; i < resolution; i++) {
        int found = 0;
        float mx = 0, my = 0;
        for (; j < resolution && j < (uint32_t)n_pts; j++) {
            const float v0x = pts[4 * (j - 1)], v0y = pts[4 * (j - 1) + 1], v1x = pts[4 * j], v1y = pts[4 * j + 1];
            const float h0 = pts[4 * (j - 1) + 2], h1 = pts[4 * j + 2];
            float t0, t1;
            t_circle_line(v0x, v0y, v1x, v1y, seg[0], seg[1], step, &t0, &t1);
            if (t0 >= 0 && t0 <= 1 && t0 > last_t) {
                found = 1; mx = v0x + (v1x - v0x) * t0; my = v0y + (v1y - v0y) * t0;
                seg[2] = t0 * h1 + (1 - t0) * h0; last_t = t0;
                break;
            } else if (t1 >= 0 && t1 <= 1 && t1 > last_t) {
                found = 1; mx = v0x + (v1x - v0x) * t1; my = v0y + (v1y - v0y) * t1;
                seg[2] = t1 * h1 + (1 - t1) * h0; last_t = t1;
                break;
            }
            last_t = -1;
        }
        if (found) {
            seg[0] = mx; seg[1] = my;
            memcpy(out + 4 * n_out++, seg, sizeof seg);
        } else if (j >= resolution) {
            if (n_pts >= 3) {
                const float lx = pts[4 * (n_pts - 1)] - pts[4 * (n_pts - 2)], ly = pts[4 * (n_pts - 1) + 1] - pts[4 * (n_pts - 2) + 1];
                const float l1x = pts[4 * (n_pts - 2)] - pts[4 * (n_pts - 3)], l1y = pts[4 * (n_pts - 2) + 1] - pts[4 * (n_pts - 3) + 1];
                const float angle0 = atan2_f(ly, lx), angle1 = atan2_f(l1y, l1x);
                const float change = angle0 - angle1;
                float angle = angle0;
                while ((uint64_t)n_out < resolution) {
                    seg[0] += cos_f(angle) * step; seg[1] += sin_f(angle) * step;
                    angle += change;
                    seg[2] *= 0.5f;
                    memcpy(out + 4 * n_out++, seg, sizeof seg);
                    i++;
                }
            }
            break;
        }
    }
    memcpy(pts, out, sizeof(float) * 4 * (size_t)n_out);
    free(out);
    return n_out;
}

/* Midline::normalize (:1268-1456).  seg: n post-processed segments.  out: `resolution` normalised segments (rotated so that the
 * midline points along -x from the origin); info = {len, angle, offset.x, offset.y}.  fix_length <= 0: none (the default call,
 * Individual.cpp:1372); > 0: Individual::fixed_midline (:507-522).  Returns resolution, or 0 where the reference returns nullptr. */
int64_t to_midline_normalize(const float *seg, int64_t n, const to_posture_params_t *P, float fix_length, float *out, float info[4])
{
    if (n < 2) return 0;
    double len = 0.0f;
    for (int64_t i = 1; i < n; i++) {
        const float lx = seg[4 * i] - seg[4 * (i - 1)], ly = seg[4 * i + 1] - seg[4 * (i - 1) + 1];
        len += sqrtf(lx * lx + ly * ly);
    }
    if (len == 0.0) return 0;
    const uint32_t resolution = (uint32_t)P->midline_resolution;
    const int max_segments = (int)(resolution - 1);
    const double step = len / (double)max_segments;
    if (step < 0) return 0;                                          /* the reference throws */
    size_t index = 0;
    float *red = (float *)malloc(sizeof(float) * 4 * (size_t)(n + resolution + 4));
    int64_t nr = 0;
    memcpy(red + 4 * nr++, seg, sizeof(float) * 4);
    double last_pt_distance = 0.0, distance;
    for (distance = 0.0; distance <= len && index < (size_t)n - 1;) {
        while (distance - last_pt_distance < step && index < (size_t)n - 1) {
            const float lx = seg[4 * (index + 1)] - seg[4 * index], ly = seg[4 * (index + 1) + 1] - seg[4 * index + 1];
            distance += sqrtf(lx * lx + ly * ly);
            index++;
        }
        float off = (float)(distance - last_pt_distance);
        if (off < step) break;
        while (off >= step) {
            off = (float)((double)off - step);
            if (index > 0) {
                const float *s0 = seg + 4 * (index - 1), *s1 = seg + 4 * index;
                const float lx = s1[0] - s0[0], ly = s1[1] - s0[1];
                const float local_d = sqrtf(lx * lx + ly * ly);
                float percent = off;
                if (local_d > 0) percent /= local_d;
                percent = 1.f - percent;
                float *o = red + 4 * nr++;
                o[0] = s0[0] + lx * percent; o[1] = s0[1] + ly * percent;
                o[2] = (float)((double)(s0[2] * percent) + (double)s1[2] * (1.0 - (double)percent));
                o[3] = s0[3] > s1[3] ? s0[3] : s1[3];                 /* cmn::max -> std::max(a, b): b only when a < b */
                const float q = (float)(1.0 - (double)percent);       /* line * (1.0 - percent): the double is cast to Scalar first (vec2.h:191-192) */
                const float qx = lx * q, qy = ly * q;
                last_pt_distance = distance - (double)sqrtf(qx * qx + qy * qy);
            } else {
                float *o = red + 4 * nr++;
                o[0] = seg[4 * index]; o[1] = seg[4 * index + 1]; o[2] = seg[4 * index + 2]; o[3] = 0;   /* l_length is indeterminate in the reference; unreachable (index > 0 after the walk) */
                last_pt_distance = distance;
            }
            if ((uint64_t)nr > (uint64_t)n + resolution) { free(red); return 0; }   /* cannot be normalised to `resolution` points anyway */
        }
    }
    {
        const float lx = red[4 * (nr - 1)] - seg[4 * (n - 1)], ly = red[4 * (nr - 1) + 1] - seg[4 * (n - 1) + 1];
        if ((double)sqrtf(lx * lx + ly * ly) >= 0.01) memcpy(red + 4 * nr++, seg + 4 * (n - 1), sizeof(float) * 4);
    }
    if ((uint64_t)nr != resolution) { free(red); return 0; }
    {
        const float lx = red[4] - red[0], ly = red[5] - red[1];
        float percent = sqrtf(lx * lx + ly * ly);
        if (len > 0) percent = (float)((double)percent / len);
        red[2] = (float)((double)(red[4 + 2] * percent) + (double)red[2] * (1.0 - (double)percent));
    }
    if (fix_length > 0) {
        reverse_segments(red, nr);
        nr = midline_fix_length(fix_length, red, nr, resolution);
        reverse_segments(red, nr);
    }
    len = 0.0f;
    for (int64_t i = 1; i < nr; i++) {
        const float lx = red[4 * i] - red[4 * (i - 1)], ly = red[4 * i + 1] - red[4 * (i - 1) + 1];
        len += sqrtf(lx * lx + ly * ly);
    }
    const float ang = midline_calculate_angle(red, nr, P->midline_stiff_percentage);
    const float angle = (float)((double)(-ang) + M_PI);
    const float offx = red[4 * (nr - 1)], offy = red[4 * (nr - 1) + 1];
    /* gui::Transform tf; tf.rotate(DEGREE(angle)); tf.translate(-offx, -offy)  (C/gui/Transform.cpp:118-160, doubles) */
    const float deg = angle * (1.0f / (float)M_PI * 180.0f);
    const double rad = (double)deg * 3.141592654 / 180.0;
    const double c = cos(rad), s = sin(rad);
    const double tx = (double)(-offx), ty = (double)(-offy);
    const double m0 = 1.0 * c + 0.0 * s + 0.0 * 0.0, m4 = 1.0 * (-s) + 0.0 * c + 0.0 * 0.0, m1 = 0.0 * c + 1.0 * s + 0.0 * 0.0, m5 = 0.0 * (-s) + 1.0 * c + 0.0 * 0.0;
    const double m12 = m0 * tx + m4 * ty + 0.0 * 1.0, m13 = m1 * tx + m5 * ty + 0.0 * 1.0;
    for (int64_t i = nr - 1, k = 0; i >= 0; i--, k++) {
        const double x = red[4 * i], y = red[4 * i + 1];
        out[4 * k] = (float)(m0 * x + m4 * y + m12); out[4 * k + 1] = (float)(m1 * x + m5 * y + m13);
        out[4 * k + 2] = red[4 * i + 2]; out[4 * k + 3] = red[4 * i + 3];
    }
    const float fx = out[0], fy = out[1];
    if (fx != 0 || fy != 0)
        for (int64_t k = 0; k < nr; ++k) { out[4 * k] -= fx; out[4 * k + 1] -= fy; }
    info[0] = (float)len; info[1] = ang; info[2] = offx; info[3] = offy;
    free(red);
    return nr;
}

/* The affine map of the `posture` (legacy = 0) / `legacy` (1) crop: Midline::transform(type) (:1238-1256: translate(-front())
 * with front() never set = (0,0), rotate(DEGREE(-angle + pi/4 | pi)), translate(-offset())) combined behind normalize_image's
 * translate(size / 2) . scale(individual_image_scale) . translate(midline_length * 0.4 | (-midline_length / 2, 0))
 * (FilterCache.cpp:47-62).  M = {m00, m01, m02, m10, m11, m12}, as toCV() hands it to cv::warpAffine. */
static void tf_combine(double a[9], const double b[9])               /* row-major 3x3: a = a . b (Transform::combine) */
{
    double r[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
    memcpy(a, r, sizeof r);
}
void to_posture_matrix(float midline_angle, float offx, float offy, float midline_length, float image_scale, int legacy,
                       int out_w, int out_h, double M[6])
{
    double mt[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    {
        const float angle = legacy ? (float)((double)(-midline_angle) + M_PI) : (float)((double)(-midline_angle) + M_PI * (double)0.25f);
        const double t0[9] = {1, 0, (double)(-0.f), 0, 1, (double)(-0.f), 0, 0, 1};
        tf_combine(mt, t0);
        const float deg = angle * (1.0f / (float)M_PI * 180.0f);
        const double rad = (double)deg * 3.141592654 / 180.0, c = cos(rad), s = sin(rad);
        const double rot[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
        tf_combine(mt, rot);
        const double t1[9] = {1, 0, (double)(-offx), 0, 1, (double)(-offy), 0, 0, 1};
        tf_combine(mt, t1);
    }
    double tr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double t0[9] = {1, 0, (double)((float)out_w * 0.5f), 0, 1, (double)((float)out_h * 0.5f), 0, 0, 1};
    tf_combine(tr, t0);
    const double sc[9] = {(double)image_scale, 0, 0, 0, (double)image_scale, 0, 0, 0, 1};
    tf_combine(tr, sc);
    if (legacy) {
        const double t1[9] = {1, 0, (double)(-midline_length * 0.5f) /* float * double 0.5 is exact either way */, 0, 1, 0.0, 0, 0, 1};
        tf_combine(tr, t1);
    } else {
        const float v = (float)((double)midline_length * 0.4);        /* Vec2(midline_length * 0.4): double product cast to Scalar */
        const double t1[9] = {1, 0, (double)v, 0, 1, (double)v, 0, 0, 1};
        tf_combine(tr, t1);
    }
    tf_combine(tr, mt);
    M[0] = tr[0]; M[1] = tr[1]; M[2] = tr[2]; M[3] = tr[3]; M[4] = tr[4]; M[5] = tr[5];
}

/* The crop: calculate_normalized_diff_image (FilterCache.cpp:133-154) -> normalize_image (:21-115): the blob rendered as for the
 * `moments` crop, warped (INTER_LINEAR) into the zero-filled out_w x out_h canvas.  midline_length < 0: no image (:32-39) -> returns 0. */
int to_crop_blob_posture(const to_line_t *lines, int64_t n_lines, const uint8_t *px, const uint8_t *bg, int bg_w, int method,
                         float midline_angle, float offx, float offy, float midline_length, float image_scale, int legacy,
                         int out_w, int out_h, uint8_t *out)
{
    if (midline_length < 0) return 0;
    int32_t r[4];
    int mx = 1 << 30, my = 1 << 30, Mx = -1, My = -1;
    for (int64_t i = 0; i < n_lines; ++i) {
        if (lines[i].x0 < mx) mx = lines[i].x0;
        if (lines[i].y < my) my = lines[i].y;
        if (lines[i].x1 > Mx) Mx = lines[i].x1;
        if (lines[i].y > My) My = lines[i].y;
    }
    const int bw = Mx - mx + 1, bh = My - my + 1;
    uint8_t *img = (uint8_t *)malloc((size_t)bw * bh);
    if (method == 0) to_image_from_lines(lines, n_lines, px, bg, bg_w, 0, 0, r, NULL, img, NULL);
    else             to_image_from_lines(lines, n_lines, px, bg, bg_w, method, 0, r, NULL, NULL, img);
    double M[6];
    to_posture_matrix(midline_angle, offx, offy, midline_length, image_scale, legacy, out_w, out_h, M);
    to_warp_affine_u8(img, bw, bh, M, out, out_w, out_h);
    free(img);
    return 1;
}
