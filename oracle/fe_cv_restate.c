/*
 * oracle/fe_cv_restate.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the three OpenCV algorithms that Ground-Fusion's front end calls
 * (reference: vins_estimator/src/featureTracker/feature_tracker.cpp:118-153 calcOpticalFlowPyrLK,
 * :198 goodFeaturesToTrack).  The arithmetic lives in OpenCV (un-vendored third-party dependency;
 * pinned here to the opencv-python-headless 4.13.0 wheel of this image, SSE3 baseline build).
 * Every function below is pinned against cv2 4.13.0 by tests/test_fe_oracle.py (bit-exact).
 *
 * Build: gcc -O2 -ffp-contract=off -msse2 -mfpmath=sse -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off is REQUIRED: every float op below must round individually.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#define GF_EXPORT __attribute__((visibility("default")))

static inline int reflect101(int i, int n)
{
    /* BORDER_REFLECT_101: ...cba|abc...; callers guarantee |overshoot| < n */
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

/* round-half-to-even of a float (cvRound on SSE2 = cvtss2si under default MXCSR) */
static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_floor_f(float v) { int i = (int)v; return i - (i > v); }

/* ---------------------------------------------------------------------------------------------
 * cv::pyrDown for CV_8UC1, BORDER_REFLECT_101.
 * dst is ((w+1)/2) x ((h+1)/2), tightly packed.
 * ------------------------------------------------------------------------------------------- */
GF_EXPORT void gfo_pyr_down_u8(const uint8_t* src, int w, int h, uint8_t* dst)
{
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    int* row = (int*)malloc(sizeof(int) * (size_t)dw * 5);
    for (int y = 0; y < dh; y++) {
        for (int k = 0; k < 5; k++) {
            int sy = reflect101(2 * y + k - 2, h);
            const uint8_t* s = src + (size_t)sy * w;
            int* r = row + (size_t)k * dw;
            for (int x = 0; x < dw; x++) {
                int x0 = reflect101(2 * x - 2, w), x1 = reflect101(2 * x - 1, w), x2 = 2 * x;
                int x3 = reflect101(2 * x + 1, w), x4 = reflect101(2 * x + 2, w);
                r[x] = s[x0] + 4 * s[x1] + 6 * s[x2] + 4 * s[x3] + s[x4];
            }
        }
        for (int x = 0; x < dw; x++) {
            int v = row[x] + 4 * row[dw + x] + 6 * row[2 * dw + x] + 4 * row[3 * dw + x] + row[4 * dw + x];
            dst[(size_t)y * dw + x] = (uint8_t)((v + 128) >> 8);
        }
    }
    free(row);
}

/* ---------------------------------------------------------------------------------------------
 * Scharr derivative image used by LK (calcScharrDeriv): int16 (Ix, Iy)
 * interleaved, computed on the un-padded level with REFLECT_101.
 * ------------------------------------------------------------------------------------------- */
static void scharr_deriv(const uint8_t* img, int w, int h, int16_t* d /* w*h*2 */)
{
    int* t0 = (int*)malloc(sizeof(int) * (size_t)(w + 2));
    int* t1 = (int*)malloc(sizeof(int) * (size_t)(w + 2));
    for (int y = 0; y < h; y++) {
        const uint8_t* up = img + (size_t)reflect101(y - 1, h) * w;
        const uint8_t* md = img + (size_t)y * w;
        const uint8_t* dn = img + (size_t)reflect101(y + 1, h) * w;
        for (int x = 0; x < w; x++) {
            t0[x + 1] = (up[x] + dn[x]) * 3 + md[x] * 10;
            t1[x + 1] = dn[x] - up[x];
        }
        t0[0] = t0[2]; t0[w + 1] = t0[w - 1];
        t1[0] = t1[2]; t1[w + 1] = t1[w - 1];
        for (int x = 0; x < w; x++) {
            d[((size_t)y * w + x) * 2 + 0] = (int16_t)(t0[x + 2] - t0[x]);
            d[((size_t)y * w + x) * 2 + 1] = (int16_t)((t1[x + 2] + t1[x]) * 3 + t1[x + 1] * 10);
        }
    }
    free(t0); free(t1);
}

#define WIN 21
#define HALF 10.0f
#define W_BITS 14
#define DESCALE(v, n) (((v) + (1 << ((n) - 1))) >> (n))

static inline int img_at(const uint8_t* img, int w, int h, int x, int y)
{
    return img[(size_t)reflect101(y, h) * w + reflect101(x, w)];
}
static inline int der_at(const int16_t* d, int w, int h, int x, int y, int c)
{
    if (x < 0 || y < 0 || x >= w || y >= h) return 0; /* zero-padded derivative image */
    return d[((size_t)y * w + x) * 2 + c];
}

/*
 * Float accumulation order of OpenCV's SSE (CV_SIMD128) LK code, as pinned against cv2 4.13:
 *   order_mode 0: plain sequential (row-major) -- the naive restatement, NOT what cv2 does
 *   order_mode 1: 4 SIMD lanes over columns 0..15 (lane l <- columns l, l+4, l+8, l+12 per row),
 *                 scalar tail for columns 16..20, final  tail + ((l0+l1)+(l2+l3))   [hadd order]
 *   order_mode 2: same, final tail + ((l0+l2)+(l1+l3))                              [movehl order]
 * For the mismatch vector b, the SIMD part accumulates pmaddwd pairs (col c, col c+4) converted to
 * float, in 4 chains per component, see lk_b().
 */
static float reduce4(const float q[4], int mode)
{
    if (mode == 1) return (q[0] + q[1]) + (q[2] + q[3]);
    return (q[0] + q[2]) + (q[1] + q[3]);
}

typedef struct { int16_t I[WIN * WIN]; int16_t dx[WIN * WIN]; int16_t dy[WIN * WIN]; } lk_patch;

static void lk_A(const lk_patch* p, int mode, float* A11, float* A12, float* A22)
{
    if (mode == 0) {
        float a11 = 0, a12 = 0, a22 = 0;
        for (int i = 0; i < WIN * WIN; i++) {
            int ix = p->dx[i], iy = p->dy[i];
            a11 += (float)(ix * ix); a12 += (float)(ix * iy); a22 += (float)(iy * iy);
        }
        *A11 = a11; *A12 = a12; *A22 = a22; return;
    }
    float q11[4] = {0, 0, 0, 0}, q12[4] = {0, 0, 0, 0}, q22[4] = {0, 0, 0, 0};
    float t11 = 0, t12 = 0, t22 = 0;
    for (int y = 0; y < WIN; y++) {
        for (int x = 0; x < 16; x++) {
            float fx = (float)p->dx[y * WIN + x], fy = (float)p->dy[y * WIN + x];
            int l = x & 3;
            q22[l] = fy * fy + q22[l]; q12[l] = fx * fy + q12[l]; q11[l] = fx * fx + q11[l];
        }
        for (int x = 16; x < WIN; x++) {
            int ix = p->dx[y * WIN + x], iy = p->dy[y * WIN + x];
            t11 += (float)(ix * ix); t12 += (float)(ix * iy); t22 += (float)(iy * iy);
        }
    }
    *A11 = t11 + reduce4(q11, mode); *A12 = t12 + reduce4(q12, mode); *A22 = t22 + reduce4(q22, mode);
}

static void lk_b(const lk_patch* p, const int* diff /* WIN*WIN */, int mode, float* b1, float* b2)
{
    if (mode == 0) {
        float s1 = 0, s2 = 0;
        for (int i = 0; i < WIN * WIN; i++) { s1 += (float)(diff[i] * p->dx[i]); s2 += (float)(diff[i] * p->dy[i]); }
        *b1 = s1; *b2 = s2; return;
    }
    /* qb0 = [x(0,4) y(0,4) x(1,5) y(1,5)], qb1 = [x(2,6) y(2,6) x(3,7) y(3,7)] per 8-column step */
    float cx[4] = {0, 0, 0, 0}, cy[4] = {0, 0, 0, 0}; /* chain k <- column pairs (k, k+4) */
    float t1 = 0, t2 = 0;
    for (int y = 0; y < WIN; y++) {
        for (int x0 = 0; x0 < 16; x0 += 8)
            for (int k = 0; k < 4; k++) {
                int a = y * WIN + x0 + k, b = a + 4;
                int sx = diff[a] * p->dx[a] + diff[b] * p->dx[b]; /* pmaddwd: exact int32 */
                int sy = diff[a] * p->dy[a] + diff[b] * p->dy[b];
                cx[k] += (float)sx; cy[k] += (float)sy;
            }
        for (int x = 16; x < WIN; x++) {
            int i = y * WIN + x;
            t1 += (float)(diff[i] * p->dx[i]); t2 += (float)(diff[i] * p->dy[i]);
        }
    }
    /* (qb0+qb1) -> [x02 y02 x13 y13]; reduce_sum([x02 x13 0 0]) */
    float x02 = cx[0] + cx[2], x13 = cx[1] + cx[3], y02 = cy[0] + cy[2], y13 = cy[1] + cy[3];
    float z = 0.f;
    float rx, ry;
    if (mode == 1) { rx = (x02 + x13) + (z + z); ry = (y02 + y13) + (z + z); }
    else { rx = (x02 + z) + (x13 + z); ry = (y02 + z) + (y13 + z); }
    *b1 = t1 + rx; *b2 = t2 + ry;
}

/* One pyramid level of cv::detail::LKTrackerInvoker for one point. */
static void lk_level(const uint8_t* I, const int16_t* dI, const uint8_t* J, int w, int h,
                     float px, float py, float* nx, float* ny, int level, uint8_t* status, int mode)
{
    const float FLT_SCALE = 1.f / (1 << 20);
    float ppx = px - HALF, ppy = py - HALF;
    int ipx = cv_floor_f(ppx), ipy = cv_floor_f(ppy);
    if (ipx < -WIN || ipx >= w || ipy < -WIN || ipy >= h) { if (level == 0) *status = 0; return; }
    float a = ppx - ipx, b = ppy - ipy;
    int iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS));
    int iw01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
    int iw10 = cv_round_f((1.f - a) * b * (1 << W_BITS));
    int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
    lk_patch p;
    for (int y = 0; y < WIN; y++)
        for (int x = 0; x < WIN; x++) {
            int X = ipx + x, Y = ipy + y;
            int iv = img_at(I, w, h, X, Y) * iw00 + img_at(I, w, h, X + 1, Y) * iw01 +
                     img_at(I, w, h, X, Y + 1) * iw10 + img_at(I, w, h, X + 1, Y + 1) * iw11;
            p.I[y * WIN + x] = (int16_t)DESCALE(iv, W_BITS - 5);
            for (int c = 0; c < 2; c++) {
                int dv = der_at(dI, w, h, X, Y, c) * iw00 + der_at(dI, w, h, X + 1, Y, c) * iw01 +
                         der_at(dI, w, h, X, Y + 1, c) * iw10 + der_at(dI, w, h, X + 1, Y + 1, c) * iw11;
                (c ? p.dy : p.dx)[y * WIN + x] = (int16_t)DESCALE(dv, W_BITS);
            }
        }
    float A11, A12, A22;
    lk_A(&p, mode, &A11, &A12, &A22);
    A11 *= FLT_SCALE; A12 *= FLT_SCALE; A22 *= FLT_SCALE;
    float D = A11 * A22 - A12 * A12;
    float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * WIN * WIN);
    if (minEig < 1e-4f || D < FLT_EPSILON) { if (level == 0) *status = 0; return; }
    D = 1.f / D;
    float qx = *nx - HALF, qy = *ny - HALF;
    float pdx = 0, pdy = 0;
    int diff[WIN * WIN];
    for (int j = 0; j < 30; j++) {
        int iqx = cv_floor_f(qx), iqy = cv_floor_f(qy);
        if (iqx < -WIN || iqx >= w || iqy < -WIN || iqy >= h) { if (level == 0) *status = 0; break; }
        a = qx - iqx; b = qy - iqy;
        iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS));
        iw01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
        iw10 = cv_round_f((1.f - a) * b * (1 << W_BITS));
        iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
        for (int y = 0; y < WIN; y++)
            for (int x = 0; x < WIN; x++) {
                int X = iqx + x, Y = iqy + y;
                int jv = img_at(J, w, h, X, Y) * iw00 + img_at(J, w, h, X + 1, Y) * iw01 +
                         img_at(J, w, h, X, Y + 1) * iw10 + img_at(J, w, h, X + 1, Y + 1) * iw11;
                diff[y * WIN + x] = DESCALE(jv, W_BITS - 5) - p.I[y * WIN + x];
            }
        float b1, b2;
        lk_b(&p, diff, mode, &b1, &b2);
        b1 *= FLT_SCALE; b2 *= FLT_SCALE;
        float dx = (A12 * b2 - A22 * b1) * D, dy = (A12 * b1 - A11 * b2) * D;
        qx += dx; qy += dy;
        *nx = qx + HALF; *ny = qy + HALF;
        if ((double)dx * dx + (double)dy * dy <= 0.01 * 0.01) break;
        if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
            *nx -= dx * 0.5f; *ny -= dy * 0.5f; break;
        }
        pdx = dx; pdy = dy;
    }
    /* epilogue of the `err` computation (always requested by the tracker): at level 0 a point whose
     * final window origin left [-21, cols) x [-21, rows) loses its status. */
    if (level == 0 && *status) {
        int fx = cv_floor_f(*nx - HALF), fy = cv_floor_f(*ny - HALF);
        if (fx < -WIN || fx >= w || fy < -WIN || fy >= h) *status = 0;
    }
}

/*
 * cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, Size(21,21), max_level,
 *                          TermCriteria(COUNT+EPS,30,0.01), use_initial_flow ? OPTFLOW_USE_INITIAL_FLOW : 0)
 * next_pts is in/out (read only when use_initial_flow).  Points are (x,y) float pairs.
 */
GF_EXPORT void gfo_lk(const uint8_t* prev, const uint8_t* next, int w, int h, const float* prev_pts,
                      float* next_pts, int n, int max_level, int use_initial_flow, uint8_t* status,
                      int order_mode)
{
    const uint8_t* P[8]; const uint8_t* N[8]; int W[8], H[8];
    uint8_t* own[16]; int nown = 0;
    P[0] = prev; N[0] = next; W[0] = w; H[0] = h;
    for (int l = 1; l <= max_level; l++) {
        W[l] = (W[l - 1] + 1) / 2; H[l] = (H[l - 1] + 1) / 2;
        uint8_t* a = (uint8_t*)malloc((size_t)W[l] * H[l]); uint8_t* b = (uint8_t*)malloc((size_t)W[l] * H[l]);
        gfo_pyr_down_u8(P[l - 1], W[l - 1], H[l - 1], a); gfo_pyr_down_u8(N[l - 1], W[l - 1], H[l - 1], b);
        P[l] = a; N[l] = b; own[nown++] = a; own[nown++] = b;
    }
    for (int i = 0; i < n; i++) status[i] = 1;
    for (int l = max_level; l >= 0; l--) {
        int16_t* d = (int16_t*)malloc(sizeof(int16_t) * 2 * (size_t)W[l] * H[l]);
        scharr_deriv(P[l], W[l], H[l], d);
        float sc = (float)(1. / (1 << l));
        for (int i = 0; i < n; i++) {
            float px = prev_pts[2 * i] * sc, py = prev_pts[2 * i + 1] * sc, nx, ny;
            if (l == max_level) {
                if (use_initial_flow) { nx = next_pts[2 * i] * sc; ny = next_pts[2 * i + 1] * sc; }
                else { nx = px; ny = py; }
            } else { nx = next_pts[2 * i] * 2.f; ny = next_pts[2 * i + 1] * 2.f; }
            next_pts[2 * i] = nx; next_pts[2 * i + 1] = ny;
            lk_level(P[l], d, N[l], W[l], H[l], px, py, &next_pts[2 * i], &next_pts[2 * i + 1], l, &status[i], order_mode);
        }
        free(d);
    }
    for (int i = 0; i < nown; i++) free(own[i]);
}

/* ---------------------------------------------------------------------------------------------
 * cv::cornerMinEigenVal(img, eig, blockSize=3, ksize=3) as reached from goodFeaturesToTrack
 * (SURVEY Appendix A.4).  variant selects the two build-dependent op orders probed by the tests:
 *   bit0: Sobel op order  (0: the cv2-4.13 order, see inline comment ; 1: plain (integer sum)*s)
 *   bit1: min-eig expression  (0: (a-c)^2 + b*b separately rounded ; 1: fma(b,b,(a-c)^2))
 * ------------------------------------------------------------------------------------------- */
GF_EXPORT void gfo_min_eig(const uint8_t* img, int w, int h, float* eig, int variant)
{
    const float s = (float)(1.0 / (4.0 * 3.0 * 255.0)) ; /* scale = 1/((1<<(ksize-1))*blockSize*255) */
    const float s2 = 2.f * s;
    size_t n = (size_t)w * h;
    float* dxx = (float*)malloc(sizeof(float) * n);
    float* dxy = (float*)malloc(sizeof(float) * n);
    float* dyy = (float*)malloc(sizeof(float) * n);
    for (int y = 0; y < h; y++) {
        int y0 = reflect101(y - 1, h), y2 = reflect101(y + 1, h);
        for (int x = 0; x < w; x++) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            const uint8_t *r0 = img + (size_t)y0 * w, *r1 = img + (size_t)y * w, *r2 = img + (size_t)y2 * w;
            int d0 = r0[xp] - r0[xm], d1 = r1[xp] - r1[xm], d2 = r2[xp] - r2[xm];
            int sm0 = r0[xm] + 2 * r0[x] + r0[xp], sm2 = r2[xm] + 2 * r2[x] + r2[xp];
            float dx, dy;
            if (!(variant & 1)) {
                /* dx: row [-1 0 1] exact, column [s 2s s] vectorised with v_muladd.
                 * dy: Sobel() scales the ROW smoothing kernel when dx==0 (kx=[s 2s s]); the generic
                 * RowFilter<uchar,float> is FMA-contracted in the AVX2 dispatch; column [-1 0 1] is a
                 * plain subtraction.  Pinned against cv2.Sobel / cv2.cornerMinEigenVal 4.13.0. */
                dx = fmaf((float)(d0 + d2), s, s2 * (float)d1);
                float rr0, rr2;
                if (x < (w / 32) * 32) {   /* vectorised body of the row filter (32-column groups): FMA chain */
                    rr0 = fmaf(s, (float)r0[xp], fmaf(s2, (float)r0[x], s * (float)r0[xm]));
                    rr2 = fmaf(s, (float)r2[xp], fmaf(s2, (float)r2[x], s * (float)r2[xm]));
                } else {                   /* scalar tail of the row filter: separately rounded */
                    rr0 = (s * (float)r0[xm] + s2 * (float)r0[x]) + s * (float)r0[xp];
                    rr2 = (s * (float)r2[xm] + s2 * (float)r2[x]) + s * (float)r2[xp];
                }
                dy = rr2 - rr0;
            } else {
                dx = (float)(d0 + 2 * d1 + d2) * s;
                dy = (float)(sm2 - sm0) * s;
            }
            dxx[(size_t)y * w + x] = dx * dx; dxy[(size_t)y * w + x] = dx * dy; dyy[(size_t)y * w + x] = dy * dy;
        }
    }
    /* boxFilter(cov, 3x3, normalize=false) on CV_32FC3: RowSum<float,double> (ksize==3 special case:
     * (S0+S1)+S2 in double) then ColumnSum<double,float>: ONE running double sum per column and
     * channel down the whole image (s0 = SUM + D[y+1]; out = (float)s0; SUM = s0 - D[y-2]) -- the
     * running sum is not always exact, so it has to be replayed in this order to be bit-equal. */
    {
        float* cov[3] = {dxx, dxy, dyy};
        double* D = (double*)malloc(sizeof(double) * (size_t)(h + 2) * w);   /* padded row sums */
        double* SUM = (double*)malloc(sizeof(double) * (size_t)w);
        float* box[3];
        for (int ch = 0; ch < 3; ch++) {
            box[ch] = (float*)malloc(sizeof(float) * n);
            for (int yp = 0; yp < h + 2; yp++) {
                const float* r = cov[ch] + (size_t)reflect101(yp - 1, h) * w;
                for (int x = 0; x < w; x++)
                    D[(size_t)yp * w + x] = ((double)r[reflect101(x - 1, w)] + (double)r[x]) + (double)r[reflect101(x + 1, w)];
            }
            for (int x = 0; x < w; x++) SUM[x] = (0.0 + D[x]) + D[(size_t)w + x];
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++) {
                    double s0 = SUM[x] + D[(size_t)(y + 2) * w + x];
                    box[ch][(size_t)y * w + x] = (float)s0;
                    SUM[x] = s0 - D[(size_t)y * w + x];
                }
        }
        for (size_t k = 0; k < n; k++) {
            float a = box[0][k] * 0.5f, b = box[1][k], c = box[2][k] * 0.5f;
            float t = a - c;
            float r = (variant & 2) ? fmaf(b, b, t * t) : (t * t + b * b);
            eig[k] = (a + c) - sqrtf(r);
        }
        for (int ch = 0; ch < 3; ch++) free(box[ch]);
        free(D); free(SUM);
    }
    free(dxx); free(dxy); free(dyy);
}

typedef struct { float v; int idx; } cand_t;
static int cand_cmp(const void* pa, const void* pb)
{
    const cand_t* a = (const cand_t*)pa; const cand_t* b = (const cand_t*)pb;
    if (a->v > b->v) return -1;
    if (a->v < b->v) return 1;
    return (a->idx > b->idx) ? -1 : (a->idx < b->idx);   /* greaterThanPtr: address descending */
}

/*
 * cv::goodFeaturesToTrack(img, corners, max_corners, 0.01, min_distance, mask) with blockSize 3,
 * useHarris false.  mask may be NULL.  Returns the number of corners written (x,y float pairs).
 */
GF_EXPORT int gfo_gftt(const uint8_t* img, int w, int h, const uint8_t* mask, int max_corners,
                       double quality, double min_distance, float* corners, int variant)
{
    size_t n = (size_t)w * h;
    float* eig = (float*)malloc(sizeof(float) * n);
    float* dil = (float*)malloc(sizeof(float) * n);
    gfo_min_eig(img, w, h, eig, variant);
    double maxVal = 0; int any = 0;
    for (size_t i = 0; i < n; i++)
        if (!mask || mask[i]) { if (!any || eig[i] > maxVal) { maxVal = eig[i]; any = 1; } }
    if (!any) maxVal = 0;
    float thr = (float)(maxVal * quality);
    for (size_t i = 0; i < n; i++) eig[i] = eig[i] > thr ? eig[i] : 0.f;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float m = -FLT_MAX;
            for (int j = -1; j <= 1; j++)
                for (int i = -1; i <= 1; i++) {
                    int yy = y + j, xx = x + i;
                    if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue; /* dilate border = -inf */
                    float v = eig[(size_t)yy * w + xx]; if (v > m) m = v;
                }
            dil[(size_t)y * w + x] = m;
        }
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * n);
    int nc = 0;
    for (int y = 1; y < h - 1; y++)
        for (int x = 1; x < w - 1; x++) {
            size_t k = (size_t)y * w + x;
            float v = eig[k];
            if (v != 0 && v == dil[k] && (!mask || mask[k])) { c[nc].v = v; c[nc].idx = (int)k; nc++; }
        }
    qsort(c, (size_t)nc, sizeof(cand_t), cand_cmp);
    int out = 0;
    double md2 = min_distance * min_distance;
    for (int i = 0; i < nc && !(max_corners > 0 && out == max_corners); i++) {
        int y = c[i].idx / w, x = c[i].idx - y * w;
        int good = 1;
        if (min_distance >= 1)
            for (int j = 0; j < out; j++) {
                float dx = (float)x - corners[2 * j], dy = (float)y - corners[2 * j + 1];
                if (dx * dx + dy * dy < md2) { good = 0; break; }
            }
        if (good) { corners[2 * out] = (float)x; corners[2 * out + 1] = (float)y; out++; }
    }
    free(eig); free(dil); free(c);
    return out;
}
