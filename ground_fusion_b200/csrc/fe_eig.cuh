// fe_eig.cuh -- cv::pyrDown (u8) and cv::cornerMinEigenVal(blockSize 3, Sobel 3), bit-exact with cv2 4.13.
//
// Replaces the dense part of cv::goodFeaturesToTrack (reference feature_tracker.cpp:198) and the
// pyramid construction inside cv::calcOpticalFlowPyrLK (:118-153).  Arithmetic: oracle/fe_cv_restate.c.
//
// Min-eig, one fused kernel (u8 image in, float map out; no Sobel/cov image in HBM):
//   dx = fma(d0+d2, s, 2s*d1)                 d  = P[x+1]-P[x-1] per row           (float32)
//   dy = r2 - r0,  r = fma(s,P[x+1], fma(2s,P[x], s*P[x-1]))   (columns >= (w/32)*32: (s*P[x-1]+2s*P[x])+s*P[x+1])
//   cov = (dx*dx, dx*dy, dy*dy)  float32
//   box: row sums (S0+S1)+S2 in double; column sums are ONE RUNNING double sum per column and channel
//        down the whole image in OpenCV (s0 = SUM + D[y+1]; out = (float)s0; SUM = s0 - D[y-1]).
//        That chain is inherently sequential, so it is cut into bands that each start from the
//        speculated state fl(D[y0-1]+D[y0]) (the exact value whenever no rounding happened above,
//        which is the case for >99.9 % of the states); every band records its start and end state and
//        k_eig_verify replays the rare bands whose start differs from the true end of the band above.
//   eig = (a+c) - sqrt((a-c)^2 + b*b), a = c0/2, b = c1, c = c2/2                 (float32)
// Compile with -fmad=false.
#pragma once
#include "gf_common.cuh"

namespace gf {

// ------------------------------------------------------------------------------------------------
// pyrDown: out(y,x) = (sum_ij k_i k_j P[2y+i-2][2x+j-2] + 128) >> 8, k=[1 4 6 4 1], REFLECT_101
// ------------------------------------------------------------------------------------------------
constexpr int PD_TX = 32, PD_TY = 8;
__global__ void __launch_bounds__(PD_TX* PD_TY) k_pyr_down(Level src, uint8_t* dst, int dw, int dh, int dpitch)
{
    __shared__ uint8_t tile[2 * PD_TY + 3][2 * PD_TX + 4];
    __shared__ int rowf[2 * PD_TY + 3][PD_TX];
    const int ox = blockIdx.x * PD_TX, oy = blockIdx.y * PD_TY;
    const int tid = threadIdx.y * PD_TX + threadIdx.x;
    for (int i = tid; i < (2 * PD_TY + 3) * (2 * PD_TX + 3); i += PD_TX * PD_TY) {
        int r = i / (2 * PD_TX + 3), c = i - r * (2 * PD_TX + 3);
        int yy = reflect101(2 * oy - 2 + r, src.h), xx = reflect101(2 * ox - 2 + c, src.w);
        // tiles at the right/bottom edge may index beyond the last needed pixel: clamp keeps it legal
        yy = min(max(yy, 0), src.h - 1);
        xx = min(max(xx, 0), src.w - 1);
        tile[r][c] = __ldg(src.ptr + (size_t)yy * src.pitch + xx);
    }
    __syncthreads();
    for (int i = tid; i < (2 * PD_TY + 3) * PD_TX; i += PD_TX * PD_TY) {
        int r = i / PD_TX, c = i - r * PD_TX;
        const uint8_t* p = &tile[r][2 * c];
        rowf[r][c] = p[0] + 4 * p[1] + 6 * p[2] + 4 * p[3] + p[4];
    }
    __syncthreads();
    int x = ox + threadIdx.x, y = oy + threadIdx.y;
    if (x < dw && y < dh) {
        int r = 2 * threadIdx.y, c = threadIdx.x;
        int v = rowf[r][c] + 4 * rowf[r + 1][c] + 6 * rowf[r + 2][c] + 4 * rowf[r + 3][c] + rowf[r + 4][c];
        dst[(size_t)y * dpitch + x] = (uint8_t)((v + 128) >> 8);
    }
}

// ------------------------------------------------------------------------------------------------
// min-eig
// ------------------------------------------------------------------------------------------------
constexpr int EIG_TX = 128;   // columns per CTA (= threads per CTA)
constexpr int EIG_BAND = 16;  // rows per speculative band

// tail: column >= (w/32)*32, where cv2's row filter runs its scalar (non-FMA) remainder loop
__device__ __forceinline__ void sobel_dxdy(int p00, int p01, int p02, int p10, int p11, int p12, int p20,
                                           int p21, int p22, bool tail, float& dx, float& dy)
{
    const float s = (float)(1.0 / (4.0 * 3.0 * 255.0));
    const float s2 = 2.f * s;
    (void)p11;
    int d0 = p02 - p00, d1 = p12 - p10, d2 = p22 - p20;
    dx = __fmaf_rn((float)(d0 + d2), s, s2 * (float)d1);
    float r0, r2;
    if (!tail) {
        r0 = __fmaf_rn(s, (float)p02, __fmaf_rn(s2, (float)p01, s * (float)p00));
        r2 = __fmaf_rn(s, (float)p22, __fmaf_rn(s2, (float)p21, s * (float)p20));
    } else {
        r0 = __fadd_rn(__fadd_rn(__fmul_rn(s, (float)p00), __fmul_rn(s2, (float)p01)), __fmul_rn(s, (float)p02));
        r2 = __fadd_rn(__fadd_rn(__fmul_rn(s, (float)p20), __fmul_rn(s2, (float)p21)), __fmul_rn(s, (float)p22));
    }
    dy = r2 - r0;
}

__device__ __forceinline__ float eig_from_box(double c0, double c1, double c2)
{
    float a = (float)c0 * 0.5f, b = (float)c1, c = (float)c2 * 0.5f;
    float t = a - c;
    return (a + c) - sqrtf(t * t + b * b);
}

// spec_start/spec_end: [band][3][w] doubles
__global__ void __launch_bounds__(EIG_TX) k_min_eig(Level img, float* eig, int epitch /*floats*/,
                                                    double* spec_start, double* spec_end)
{
    constexpr int IH = EIG_BAND + 4, IW = EIG_TX + 4;     // image tile rows y0-2.., cols x0-2..
    constexpr int CH = EIG_BAND + 2, CW = EIG_TX + 2;     // dx/dy tile rows y0-1.., cols x0-1..
    __shared__ uint8_t timg[IH][IW + 4];
    __shared__ float tdx[CH][CW + 2];
    __shared__ float tdy[CH][CW + 2];
    const int w = img.w, h = img.h;
    const int x0 = blockIdx.x * EIG_TX, y0 = blockIdx.y * EIG_BAND;
    const int tid = threadIdx.x;
    for (int i = tid; i < IH * IW; i += EIG_TX) {
        int r = i / IW, c = i - r * IW;
        int yy = y0 - 2 + r, xx = x0 - 2 + c;
        // clamp far-out indices (only reached by slots that are never used), reflect the border ring
        yy = min(max(yy, -1), h);
        xx = min(max(xx, -1), w);
        yy = reflect101(yy, h);
        xx = reflect101(xx, w);
        timg[r][c] = __ldg(img.ptr + (size_t)yy * img.pitch + xx);
    }
    __syncthreads();
    // Sobel at slots whose (row, col) is inside the image; ring slots are filled by reflection below
    for (int i = tid; i < CH * CW; i += EIG_TX) {
        int r = i / CW, c = i - r * CW;
        int yy = y0 - 1 + r, xx = x0 - 1 + c;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
            const uint8_t* p = &timg[r][c];  // timg row r <-> image row yy-1, col c <-> xx-1
            float dx, dy;
            sobel_dxdy(p[0], p[1], p[2], p[IW + 4], p[IW + 5], p[IW + 6], p[2 * (IW + 4)], p[2 * (IW + 4) + 1],
                       p[2 * (IW + 4) + 2], xx >= (w / 32) * 32, dx, dy);
            tdx[r][c] = dx;
            tdy[r][c] = dy;
        }
    }
    __syncthreads();
    // cov is extended by REFLECT_101 at the *cov* level: slot (-1) := slot (+1), slot (n) := slot (n-2)
    for (int i = tid; i < CH * CW; i += EIG_TX) {
        int r = i / CW, c = i - r * CW;
        int yy = y0 - 1 + r, xx = x0 - 1 + c;
        bool in = (yy >= 0 && yy < h && xx >= 0 && xx < w);
        if (!in && yy >= -1 && yy <= h && xx >= -1 && xx <= w) {
            int ry = reflect101(yy, h) - (y0 - 1), rx = reflect101(xx, w) - (x0 - 1);
            if (ry >= 0 && ry < CH && rx >= 0 && rx < CW) {
                tdx[r][c] = tdx[ry][rx];
                tdy[r][c] = tdy[ry][rx];
            }
        }
    }
    __syncthreads();
    const int x = x0 + tid;
    if (x >= w) return;
    const int nrows = min(EIG_BAND, h - y0);
    // D for tile row r (image row y0-1+r) at this thread's column
    auto rowsum = [&](int r, double& d0, double& d1, double& d2) {
        const float* ax = &tdx[r][tid];  // cols x-1, x, x+1 <-> tid, tid+1, tid+2
        const float* ay = &tdy[r][tid];
        float xl = ax[0], xc = ax[1], xr = ax[2], yl = ay[0], yc = ay[1], yr = ay[2];
        d0 = ((double)(xl * xl) + (double)(xc * xc)) + (double)(xr * xr);
        d1 = ((double)(xl * yl) + (double)(xc * yc)) + (double)(xr * yr);
        d2 = ((double)(yl * yl) + (double)(yc * yc)) + (double)(yr * yr);
    };
    double a0, a1, a2, b0, b1, b2;  // D[y-1], D[y]
    rowsum(0, a0, a1, a2);
    rowsum(1, b0, b1, b2);
    double S0 = (0.0 + a0) + b0, S1 = (0.0 + a1) + b1, S2 = (0.0 + a2) + b2;
    const size_t so = ((size_t)blockIdx.y * 3) * w + x;
    spec_start[so] = S0;
    spec_start[so + w] = S1;
    spec_start[so + 2 * (size_t)w] = S2;
    for (int k = 0; k < nrows; k++) {
        double c0, c1, c2;
        rowsum(k + 2, c0, c1, c2);
        double s0 = S0 + c0, s1 = S1 + c1, s2 = S2 + c2;
        eig[(size_t)(y0 + k) * epitch + x] = eig_from_box(s0, s1, s2);
        S0 = s0 - a0; S1 = s1 - a1; S2 = s2 - a2;
        a0 = b0; a1 = b1; a2 = b2;
        b0 = c0; b1 = c1; b2 = c2;
    }
    spec_end[so] = S0;
    spec_end[so + w] = S1;
    spec_end[so + 2 * (size_t)w] = S2;
}

// One thread per column: check every band's speculated start against the end of the band above (independent
// loads), and replay the rare mismatching bands from the true state.  The replay first pulls the band's
// (BAND+4) x 5 pixel neighbourhood into local memory with independent loads, then runs without touching HBM.
__global__ void k_eig_verify(Level img, float* eig, int epitch, const double* spec_start,
                             const double* spec_end, int nbands, int* fixups)
{
    const int w = img.w, h = img.h;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    {   // common case: nothing to do
        bool any = false;
#pragma unroll 4
        for (int b = 1; b < nbands; b++) {
            size_t so = ((size_t)b * 3) * w + x, sp = ((size_t)(b - 1) * 3) * w + x;
            any |= (spec_start[so] != spec_end[sp]) | (spec_start[so + w] != spec_end[sp + w]) | (spec_start[so + 2 * (size_t)w] != spec_end[sp + 2 * (size_t)w]);
        }
        if (!any) return;
    }
    double t0 = spec_end[x], t1 = spec_end[(size_t)w + x], t2 = spec_end[2 * (size_t)w + x];
    for (int b = 1; b < nbands; b++) {
        size_t so = ((size_t)b * 3) * w + x;
        double s0 = spec_start[so], s1 = spec_start[so + w], s2 = spec_start[so + 2 * (size_t)w];
        if (s0 == t0 && s1 == t1 && s2 == t2) {
            t0 = spec_end[so]; t1 = spec_end[so + w]; t2 = spec_end[so + 2 * (size_t)w];
            continue;
        }
        atomicAdd(fixups, 1);
        const int y0 = b * EIG_BAND, nrows = min(EIG_BAND, h - y0);
        // pixels of image rows y0-2 .. y0+BAND+1 and columns x-2 .. x+2 (image-level REFLECT_101 applied here)
        uint8_t P[EIG_BAND + 4][5];
#pragma unroll
        for (int r = 0; r < EIG_BAND + 4; r++) {
            int yy = min(max(y0 - 2 + r, -1), h); yy = reflect101(yy, h);
#pragma unroll
            for (int c = 0; c < 5; c++) {
                int xx = min(max(x - 2 + c, -1), w); xx = reflect101(xx, w);
                P[r][c] = __ldg(img.ptr + (size_t)yy * img.pitch + xx);
            }
        }
        // D at image row r (REFLECT_101 at the cov level) for this column
        auto rowsum = [&](int r, double& d0, double& d1, double& d2) {
            const int rr = reflect101(min(max(r, -1), h), h);
            float cx[3], cy[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int cc = reflect101(min(max(x - 1 + k, -1), w), w);
                const int pr = rr - (y0 - 2), pc = cc - (x - 2);          // centre of the 3x3 neighbourhood inside P
                sobel_dxdy(P[pr - 1][pc - 1], P[pr - 1][pc], P[pr - 1][pc + 1], P[pr][pc - 1], P[pr][pc], P[pr][pc + 1],
                           P[pr + 1][pc - 1], P[pr + 1][pc], P[pr + 1][pc + 1], cc >= (w / 32) * 32, cx[k], cy[k]);
            }
            d0 = ((double)(cx[0] * cx[0]) + (double)(cx[1] * cx[1])) + (double)(cx[2] * cx[2]);
            d1 = ((double)(cx[0] * cy[0]) + (double)(cx[1] * cy[1])) + (double)(cx[2] * cy[2]);
            d2 = ((double)(cy[0] * cy[0]) + (double)(cy[1] * cy[1])) + (double)(cy[2] * cy[2]);
        };
        double a0, a1, a2, b0, b1, b2;
        rowsum(y0 - 1, a0, a1, a2);
        rowsum(y0, b0, b1, b2);
        for (int k = 0; k < nrows; k++) {
            double c0, c1, c2;
            rowsum(y0 + k + 1, c0, c1, c2);
            double u0 = t0 + c0, u1 = t1 + c1, u2 = t2 + c2;
            eig[(size_t)(y0 + k) * epitch + x] = eig_from_box(u0, u1, u2);
            t0 = u0 - a0; t1 = u1 - a1; t2 = u2 - a2;
            a0 = b0; a1 = b1; a2 = b2;
            b0 = c0; b1 = c1; b2 = c2;
        }
    }
}

}  // namespace gf
