/* ORACLE (test infrastructure, not product): CPU restatement of the reference post-processing.
 *
 * Plain C restatement of `process` / `__proc_np_hv` (reference models/hovernet/post_proc.py:26-90,
 * 94-186) and of the two helpers it calls (reference misc/utils.py:18-28 get_bounding_box,
 * misc/utils.py:142-182 remove_small_objects), with the third-party calls on the path restated
 * stage by stage:
 *   scipy.ndimage.label (post_proc.py:45,85), binary_fill_holes (:82), cv2.normalize (:49-68),
 *   cv2.Sobel ksize=21 (:56-57), cv2.GaussianBlur 3x3 (:76), cv2.morphologyEx OPEN ellipse5 (:84),
 *   skimage.segmentation.watershed (:88).
 * Every stage except the watershed is pinned bit-for-bit against the cv2 / scipy installed in this
 * image by tests/test_oracle_postproc.py.  The watershed restates scikit-image 0.17.2
 * (requirements.txt:10; _watershed.py / _watershed_cy.pyx / heap_general.pxi), which is NOT
 * installed and not vendored: PARITY UNPINNED for that one step (known-answer grids + invariants
 * only, see tests/test_oracle_watershed.py and DESIGN.md).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may load
 * this library.  The product path never does.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

/* ---------------------------------------------------------------- label (scipy.ndimage.label) */
/* 4-connectivity; labels contiguous, numbered in raster order of each component's first pixel. */
API int hvo_label4(const int32_t *bin, int H, int W, int32_t *lab)
{
    int n = H * W, next = 0;
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    memset(lab, 0, sizeof(int32_t) * (size_t)n);
    for (int p = 0; p < n; ++p) {
        if (!bin[p] || lab[p]) continue;
        ++next;
        int sp = 0;
        stack[sp++] = p;
        lab[p] = next;
        while (sp) {
            int q = stack[--sp];
            int y = q / W, x = q - y * W;
            if (y > 0 && bin[q - W] && !lab[q - W]) { lab[q - W] = next; stack[sp++] = q - W; }
            if (x > 0 && bin[q - 1] && !lab[q - 1]) { lab[q - 1] = next; stack[sp++] = q - 1; }
            if (x < W - 1 && bin[q + 1] && !lab[q + 1]) { lab[q + 1] = next; stack[sp++] = q + 1; }
            if (y < H - 1 && bin[q + W] && !lab[q + W]) { lab[q + W] = next; stack[sp++] = q + W; }
        }
    }
    free(stack);
    return next;
}

/* misc/utils.py:169-180 : zero every label whose pixel count < min_size; survivors keep ids. */
API void hvo_remove_small(int32_t *lab, int n, int nlab, int min_size)
{
    int64_t *cnt = (int64_t *)calloc((size_t)nlab + 1, sizeof(int64_t));
    for (int p = 0; p < n; ++p) cnt[lab[p]]++;
    for (int p = 0; p < n; ++p)
        if (cnt[lab[p]] < min_size) lab[p] = 0;
    free(cnt);
}

/* ------------------------------------------------------------- cv2.normalize(NORM_MINMAX, 0..1) */
static void minmax_scale(double smin, double smax, double *scale, double *shift)
{
    double sc = (smax - smin > DBL_EPSILON) ? 1.0 / (smax - smin) : 0.0;
    sc = (double)(float)sc;
    *scale = sc;
    *shift = 0.0 - (double)(float)(smin * sc);
}

API void hvo_normalize_f32(const float *src, int n, float *dst)
{
    double smin = src[0], smax = src[0], scale, shift;
    for (int i = 1; i < n; ++i) {
        if (src[i] < smin) smin = src[i];
        if (src[i] > smax) smax = src[i];
    }
    minmax_scale(smin, smax, &scale, &shift);
    for (int i = 0; i < n; ++i) dst[i] = (float)fma((double)src[i], scale, shift);
}

API void hvo_normalize_f64(const double *src, int n, float *dst)
{
    double smin = src[0], smax = src[0], scale, shift;
    for (int i = 1; i < n; ++i) {
        if (src[i] < smin) smin = src[i];
        if (src[i] > smax) smax = src[i];
    }
    minmax_scale(smin, smax, &scale, &shift);
    for (int i = 0; i < n; ++i) dst[i] = (float)fma(src[i], scale, shift);
}

/* ---------------------------------------------------------------------- cv2.Sobel(ksize = 21) */
static const double K_DERIV[21] = {-1, -18, -152, -798, -2907, -7752, -15504, -23256, -25194, -16796, 0,
                                   16796, 25194, 23256, 15504, 7752, 2907, 798, 152, 18, 1};
static const double K_SMOOTH[21] = {1, 20, 190, 1140, 4845, 15504, 38760, 77520, 125970, 167960, 184756,
                                    167960, 125970, 77520, 38760, 15504, 4845, 1140, 190, 20, 1};

static int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

/* dx=1: derivative along x (row pass), smoothing along y.  dx=0: the transpose roles. */
API void hvo_sobel21(const float *src, int H, int W, int dx, double *dst)
{
    const double *kx = dx ? K_DERIV : K_SMOOTH;
    const double *ky = dx ? K_SMOOTH : K_DERIV;
    int ky_sym = dx ? 1 : 0;
    double *row = (double *)malloc(sizeof(double) * (size_t)H * (size_t)W);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            double acc = 0.0;
            for (int k = 0; k < 21; ++k)
                acc += kx[k] * (double)src[y * W + reflect101(x + k - 10, W)];
            row[y * W + x] = acc;
        }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            double acc = ky[10] * row[y * W + x];
            for (int k = 1; k <= 10; ++k) {
                double a = row[reflect101(y + k, H) * W + x];
                double b = row[reflect101(y - k, H) * W + x];
                acc += ky[10 + k] * (ky_sym ? (a + b) : (a - b));
            }
            dst[y * W + x] = acc;
        }
    free(row);
}

/* -------------------------------------------------------------- cv2.GaussianBlur((3,3), 0) f64 */
API void hvo_gauss3_f64(const double *src, int H, int W, double *dst)
{
    double *row = (double *)malloc(sizeof(double) * (size_t)H * (size_t)W);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            double a = src[y * W + reflect101(x - 1, W)], b = src[y * W + x],
                   c = src[y * W + reflect101(x + 1, W)];
            row[y * W + x] = (a * 0.25 + b * 0.5) + c * 0.25; /* row pass: plain left-to-right sum */
        }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            double a = row[reflect101(y - 1, H) * W + x], b = row[y * W + x],
                   c = row[reflect101(y + 1, H) * W + x];
            dst[y * W + x] = b * 0.5 + (a + c) * 0.25; /* column pass: symmetric form */
        }
    free(row);
}

/* ------------------------------------------------------- scipy.ndimage.binary_fill_holes (4-conn) */
API void hvo_fill_holes(const int32_t *bin, int H, int W, uint8_t *out)
{
    int n = H * W, sp = 0;
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    uint8_t *outside = (uint8_t *)calloc((size_t)n, 1);
#define PUSH_IF(q) do { if (!bin[q] && !outside[q]) { outside[q] = 1; stack[sp++] = (q); } } while (0)
    for (int x = 0; x < W; ++x) { PUSH_IF(x); PUSH_IF((H - 1) * W + x); }
    for (int y = 0; y < H; ++y) { PUSH_IF(y * W); PUSH_IF(y * W + W - 1); }
    while (sp) {
        int q = stack[--sp];
        int y = q / W, x = q - y * W;
        if (y > 0) PUSH_IF(q - W);
        if (x > 0) PUSH_IF(q - 1);
        if (x < W - 1) PUSH_IF(q + 1);
        if (y < H - 1) PUSH_IF(q + W);
    }
#undef PUSH_IF
    for (int p = 0; p < n; ++p) out[p] = outside[p] ? 0 : 1;
    free(stack);
    free(outside);
}

/* ------------------------------------- cv2.morphologyEx(MORPH_OPEN, getStructuringElement(ELLIPSE,(5,5))) */
static const uint8_t ELLIPSE5[5][5] = {
    {0, 0, 1, 0, 0}, {1, 1, 1, 1, 1}, {1, 1, 1, 1, 1}, {1, 1, 1, 1, 1}, {0, 0, 1, 0, 0}};

API void hvo_open_ellipse5(const uint8_t *src, int H, int W, uint8_t *dst)
{
    uint8_t *er = (uint8_t *)malloc((size_t)H * (size_t)W);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            uint8_t v = 1; /* out-of-image taps are ignored (border = +inf for erode) */
            for (int j = -2; j <= 2; ++j)
                for (int i = -2; i <= 2; ++i) {
                    if (!ELLIPSE5[j + 2][i + 2]) continue;
                    int yy = y + j, xx = x + i;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    if (!src[yy * W + xx]) v = 0;
                }
            er[y * W + x] = v;
        }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            uint8_t v = 0; /* border = -inf for dilate */
            for (int j = -2; j <= 2; ++j)
                for (int i = -2; i <= 2; ++i) {
                    if (!ELLIPSE5[j + 2][i + 2]) continue;
                    int yy = y + j, xx = x + i;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    if (er[yy * W + xx]) v = 1;
                }
            dst[y * W + x] = v;
        }
    free(er);
}

/* --------------------------------------- skimage.segmentation.watershed(image, markers, mask=mask)
 * scikit-image 0.17.2, connectivity=1, compactness=0, watershed_line=False  [UNPINNED restatement].
 * One global binary heap ordered by (value, age); markers pushed in raster order with age 0; a
 * neighbour is labelled at push time; neighbour order up, left, right, down. */
typedef struct { double value; int64_t age; int32_t index; } hitem_t;

static int smaller(const hitem_t *a, const hitem_t *b)
{
    if (a->value != b->value) return a->value < b->value;
    return a->age < b->age;
}

typedef struct { hitem_t *d; int64_t n, cap; } heap_t;

static void heap_push(heap_t *h, hitem_t e)
{
    if (h->n == h->cap) {
        h->cap = h->cap ? h->cap * 2 : 1024;
        h->d = (hitem_t *)realloc(h->d, sizeof(hitem_t) * (size_t)h->cap);
    }
    int64_t child = h->n++;
    h->d[child] = e;
    while (child > 0) {
        int64_t parent = (child + 1) / 2 - 1;
        if (smaller(&h->d[child], &h->d[parent])) {
            hitem_t t = h->d[parent]; h->d[parent] = h->d[child]; h->d[child] = t;
            child = parent;
        } else break;
    }
}

static hitem_t heap_pop(heap_t *h)
{
    hitem_t top = h->d[0];
    h->d[0] = h->d[--h->n];
    int64_t i = 0;
    for (;;) {
        int64_t l = 2 * i + 1, r = 2 * i + 2, s = i;
        if (l < h->n && smaller(&h->d[l], &h->d[s])) s = l;
        if (r < h->n && smaller(&h->d[r], &h->d[s])) s = r;
        if (s == i) break;
        hitem_t t = h->d[s]; h->d[s] = h->d[i]; h->d[i] = t;
        i = s;
    }
    return top;
}

API void hvo_watershed(const double *image, const int32_t *markers, const int32_t *mask, int H, int W,
                       int32_t *out)
{
    int n = H * W;
    heap_t hp = {0, 0, 0};
    int64_t age = 1;
    for (int p = 0; p < n; ++p) out[p] = mask[p] ? markers[p] : 0; /* markers * mask */
    for (int p = 0; p < n; ++p)
        if (out[p]) { hitem_t e = {image[p], 0, p}; heap_push(&hp, e); }
    while (hp.n > 0) {
        hitem_t e = heap_pop(&hp);
        int y = e.index / W, x = e.index - y * W;
        int nb[4], k = 0;
        nb[k++] = (y > 0) ? e.index - W : -1;
        nb[k++] = (x > 0) ? e.index - 1 : -1;
        nb[k++] = (x < W - 1) ? e.index + 1 : -1;
        nb[k++] = (y < H - 1) ? e.index + W : -1;
        for (int i = 0; i < 4; ++i) {
            int q = nb[i];
            if (q < 0 || !mask[q] || out[q]) continue;
            age += 1;
            out[q] = out[e.index];
            hitem_t ne = {image[q], age, q};
            heap_push(&hp, ne);
        }
    }
    free(hp.d);
}

/* ------------------------------------------------------------------ __proc_np_hv (post_proc.py:26-90)
 * pred: float32 [H,W,3] = (np_prob, hv_x, hv_y) with pixel stride `cs` floats (3 or 4).
 * stage pointers (any may be NULL) receive the intermediates for per-stage parity tests. */
typedef struct {
    int32_t *blb;      /* after CCL#1 + remove_small + binarise   (:43-47) */
    double *sobelh;    /* cv2.Sobel outputs                        (:56-57) */
    double *sobelv;
    float *overall32;  /* max(1-norm(sobelh), 1-norm(sobelv))      (:59-70) */
    double *dist;      /* -GaussianBlur(...)                        (:74-76) */
    int32_t *marker;   /* labelled markers after remove_small       (:78-86) */
} hvo_stages_t;

API void hvo_proc_np_hv(const float *pred, int cs, int H, int W, int32_t *inst, hvo_stages_t *st)
{
    int n = H * W;
    int32_t *blb = (int32_t *)calloc((size_t)n, sizeof(int32_t));
    int32_t *lab = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    float *hraw = (float *)malloc(sizeof(float) * (size_t)n);
    float *vraw = (float *)malloc(sizeof(float) * (size_t)n);
    float *hn = (float *)malloc(sizeof(float) * (size_t)n);
    float *vn = (float *)malloc(sizeof(float) * (size_t)n);
    double *sh = (double *)malloc(sizeof(double) * (size_t)n);
    double *sv = (double *)malloc(sizeof(double) * (size_t)n);
    double *ov = (double *)malloc(sizeof(double) * (size_t)n);
    double *dist = (double *)malloc(sizeof(double) * (size_t)n);
    double *blur = (double *)malloc(sizeof(double) * (size_t)n);
    uint8_t *filled = (uint8_t *)malloc((size_t)n);
    uint8_t *opened = (uint8_t *)malloc((size_t)n);

    for (int p = 0; p < n; ++p) {
        blb[p] = pred[(size_t)p * cs + 0] >= 0.5f ? 1 : 0;
        hraw[p] = pred[(size_t)p * cs + 1];
        vraw[p] = pred[(size_t)p * cs + 2];
    }
    int nl = hvo_label4(blb, H, W, lab);
    hvo_remove_small(lab, n, nl, 10);
    for (int p = 0; p < n; ++p) blb[p] = lab[p] > 0 ? 1 : 0;

    hvo_normalize_f32(hraw, n, hn);
    hvo_normalize_f32(vraw, n, vn);
    hvo_sobel21(hn, H, W, 1, sh);
    hvo_sobel21(vn, H, W, 0, sv);
    hvo_normalize_f64(sh, n, hn); /* reuse hn/vn as the normalised Sobel maps */
    hvo_normalize_f64(sv, n, vn);
    for (int p = 0; p < n; ++p) {
        float a = 1.0f - hn[p], b = 1.0f - vn[p];
        float o32 = a > b ? a : b;                 /* np.maximum, float32 */
        if (st && st->overall32) st->overall32[p] = o32;
        double o = (double)o32 - (double)(1 - blb[p]); /* float32 - int32 -> float64 (:71) */
        if (o < 0) o = 0;
        ov[p] = o;
        dist[p] = (1.0 - o) * (double)blb[p];
    }
    hvo_gauss3_f64(dist, H, W, blur);
    for (int p = 0; p < n; ++p) blur[p] = -blur[p];

    for (int p = 0; p < n; ++p) {
        int m = blb[p] - (ov[p] >= 0.4 ? 1 : 0);
        lab[p] = m < 0 ? 0 : m;
    }
    hvo_fill_holes(lab, H, W, filled);
    hvo_open_ellipse5(filled, H, W, opened);
    {
        int32_t *mk = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
        for (int p = 0; p < n; ++p) mk[p] = opened[p];
        int nm = hvo_label4(mk, H, W, lab);
        hvo_remove_small(lab, n, nm, 10);
        free(mk);
    }
    hvo_watershed(blur, lab, blb, H, W, inst);

    if (st) {
        if (st->blb) memcpy(st->blb, blb, sizeof(int32_t) * (size_t)n);
        if (st->sobelh) memcpy(st->sobelh, sh, sizeof(double) * (size_t)n);
        if (st->sobelv) memcpy(st->sobelv, sv, sizeof(double) * (size_t)n);
        if (st->dist) memcpy(st->dist, blur, sizeof(double) * (size_t)n);
        if (st->marker) memcpy(st->marker, lab, sizeof(int32_t) * (size_t)n);
    }
    free(blb); free(lab); free(hraw); free(vraw); free(hn); free(vn); free(sh); free(sv);
    free(ov); free(dist); free(blur); free(filled); free(opened);
}

/* ------------------------------------------------------------------ process (post_proc.py:94-186)
 * pred_map float32 [H,W,C]: C==3 -> (np,hx,hy); C==4 -> (tp,np,hx,hy) and nr_types>0.
 * Table row per instance id present in inst (ascending id), 16 x int64/double slots:
 *   id, rmin, cmin, rmax, cmax (max exclusive), area, sum_x, sum_y, type, type_count
 * Centroid (cv2.moments m10/m00, m01/m00 of the bbox crop + offset) == sum_x/area, sum_y/area.
 * The <3-point-contour drop (:140-143) needs cv2.findContours and is applied by the Python side. */
#define HVO_ROW 10

API int hvo_process(const float *pred_map, int H, int W, int C, int nr_types, int32_t *inst,
                    int64_t *table, int max_rows)
{
    int n = H * W;
    const float *np_hv = (C == 4) ? pred_map + 1 : pred_map;
    hvo_proc_np_hv(np_hv, C, H, W, inst, NULL);
    int32_t maxid = 0;
    for (int p = 0; p < n; ++p) if (inst[p] > maxid) maxid = inst[p];
    if (maxid == 0) return 0;
    int ntp = nr_types > 0 ? nr_types : 1;
    int64_t *acc = (int64_t *)calloc((size_t)(maxid + 1) * 7, sizeof(int64_t));
    int64_t *tcnt = (int64_t *)calloc((size_t)(maxid + 1) * (size_t)ntp, sizeof(int64_t));
    for (int i = 0; i <= maxid; ++i) { acc[i * 7 + 0] = H; acc[i * 7 + 1] = W; acc[i * 7 + 2] = -1; acc[i * 7 + 3] = -1; }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int id = inst[y * W + x];
            if (!id) continue;
            int64_t *a = acc + (size_t)id * 7;
            if (y < a[0]) a[0] = y;
            if (x < a[1]) a[1] = x;
            if (y > a[2]) a[2] = y;
            if (x > a[3]) a[3] = x;
            a[4] += 1; a[5] += x; a[6] += y;
            if (nr_types > 0) {
                int t = (int)pred_map[(size_t)(y * W + x) * C]; /* astype(int32) truncation (:111) */
                if (t >= 0 && t < nr_types) tcnt[(size_t)id * ntp + t]++;
            }
        }
    int rows = 0;
    for (int id = 1; id <= maxid; ++id) {
        int64_t *a = acc + (size_t)id * 7;
        if (a[4] == 0) continue;
        if (rows >= max_rows) { rows = -1; break; }
        int64_t *r = table + (size_t)rows * HVO_ROW;
        r[0] = id; r[1] = a[0]; r[2] = a[1]; r[3] = a[2] + 1; r[4] = a[3] + 1;
        r[5] = a[4]; r[6] = a[5]; r[7] = a[6];
        r[8] = -1; r[9] = 0;
        if (nr_types > 0) {
            /* sorted by count desc, stable => ties to the smaller type id (:172-177) */
            int best = -1, second = -1;
            for (int t = 0; t < nr_types; ++t) {
                int64_t c = tcnt[(size_t)id * ntp + t];
                if (c == 0) continue;
                if (best < 0 || c > tcnt[(size_t)id * ntp + best]) { second = best; best = t; }
                else if (second < 0 || c > tcnt[(size_t)id * ntp + second]) second = t;
            }
            int pick = best;
            if (best == 0 && second >= 0) pick = second;
            r[8] = pick;
            r[9] = pick >= 0 ? tcnt[(size_t)id * ntp + pick] : 0;
        }
        ++rows;
    }
    free(acc); free(tcnt);
    return rows;
}

/* ------------------------------------------------------------------------------------------------
 * Contour of one instance: the first contour of cv2.findContours(crop, RETR_TREE, CHAIN_APPROX_SIMPLE)
 * on the bbox crop `inst[rmin:rmax, cmin:cmax] == id` (reference post_proc.py:133-137), plus the bbox
 * offset (:146-147).  Restates OpenCV's Suzuki-Abe border following (modules/imgproc/src/contours.cpp,
 * icvFetchContour; 4.x `contours_new.cpp` yields the same sequence -- pinned against cv2 4.13 in
 * tests/test_oracle_contour.py): start at the first pixel in raster order, direction codes
 * 0..7 = E,NE,N,NW,W,SW,S,SE, the previous pixel is the first non-zero neighbour clockwise from NW,
 * each step searches counter-clockwise from the previous pixel, and CHAIN_APPROX_SIMPLE keeps a point
 * only where the step direction changes.  An instance is one 8-connected component (flood regions are
 * 4-connected), so the first contour IS this outer border.
 * Returns the number of points (x, y pairs written to pts up to cap points). */
static const int HVO_DX[8] = {1, 1, 0, -1, -1, -1, 0, 1};
static const int HVO_DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};

API int hvo_contour(const int32_t *inst, int H, int W, int32_t id, int rmin, int cmin, int rmax, int cmax,
                    int32_t *pts, int cap)
{
    (void)H;
#define F(y, x) ((y) >= rmin && (y) < rmax && (x) >= cmin && (x) < cmax && inst[(size_t)(y) * W + (x)] == id)
    int y0 = -1, x0 = -1;
    for (int y = rmin; y < rmax && y0 < 0; ++y)
        for (int x = cmin; x < cmax; ++x)
            if (inst[(size_t)y * W + x] == id) { y0 = y; x0 = x; break; }
    if (y0 < 0) return 0;
    int n = 0;
    int s_end = 4, s = 4;
    do {
        s = (s - 1) & 7;
        if (F(y0 + HVO_DY[s], x0 + HVO_DX[s])) break;
    } while (s != s_end);
    if (s == s_end) { /* isolated pixel */
        if (n < cap) { pts[0] = x0; pts[1] = y0; }
        return 1;
    }
    const int y1 = y0 + HVO_DY[s], x1 = x0 + HVO_DX[s];
    int y3 = y0, x3 = x0, prev_s = s ^ 4;
    for (;;) {
        int y4, x4;
        for (;;) {
            ++s;
            y4 = y3 + HVO_DY[s & 7]; x4 = x3 + HVO_DX[s & 7];
            if (F(y4, x4)) break;
        }
        s &= 7;
        if (s != prev_s) {
            if (n < cap) { pts[2 * n] = x3; pts[2 * n + 1] = y3; }
            ++n;
            prev_s = s;
        }
        if (y4 == y0 && x4 == x0 && y3 == y1 && x3 == x1) break;
        y3 = y4; x3 = x4;
        s = (s + 4) & 7;
    }
#undef F
    return n;
}
