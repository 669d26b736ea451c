// ba_chol.cuh -- dense FP64 linear algebra of the sliding-window solve on 8x8 tiles with FP64 tensor-core MMAs
// (mma.sync m8n8k4 f64 = SASS DMMA.8x8x4; measured on B200: 64 FMA/clk/SM, 26 cycles dependent latency, tools/dmma_probe.cu).
//
// Replaces Ceres' DENSE_SCHUR linear solver (SchurEliminator + dense Cholesky; reference call site estimator.cpp:3303-3318):
//   schur_tile      one 8x8 tile of the reduced camera system  S = H' + mu D^2 - W'^T C W'  (rank-L update on DMMA)
//   chol_factor     left-looking tiled Cholesky of the (nc+1)-row augmented system (row nc = right-hand side) by one CTA:
//                   tile (I,J) = A_IJ - sum_p L_Ip L_Jp^T accumulated in registers by DMMAs, 8x8 diagonal factorisation
//                   + inverse by one warp in registers, TRSM as two DMMAs against the published inverse
//   chol_backsubst  L^T y = z by one warp
// Tile storage ("fragment order"): tile element [r][c] at  (c/4)*32 + r*4 + c%4 , so that the k-half h of a tile is the
// contiguous run [h*32, h*32+32): lane l of an m8n8k4 operand fragment reads element h*32 + l.
#pragma once
#include "gf_common.cuh"

namespace gfba {

constexpr int ST_THREADS = 256;                 // k_ba_step block size: 8 warps x 255 registers (the 8x8 diagonal factorisation lives in registers)
constexpr int ST_WARPS = ST_THREADS / 32;
constexpr int TILE_CAP = 384;                   // factor tiles resident in shared memory (512 B each); the rest spills to L2
constexpr int MAX_N8 = 48;                      // block rows of the augmented system
constexpr int MAX_NC = 8 * MAX_N8 - 1;          // reduced dimension supported by the solver (383)
constexpr int MAXR = (MAX_N8 + ST_WARPS - 1) / ST_WARPS;   // block rows owned by one warp (kernels are instantiated for MAXR / 2 and MAXR)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__host__ __device__ __forceinline__ int tix(int I, int J) { return I * (I + 1) / 2 + J; }

// Factor tiles live in shared memory up to TILE_CAP, beyond that in a global (L2-resident) spill buffer.
struct TileStore {
    double* Ls;
    double* Lg;
    int cap;                  // tiles [0, cap) in shared memory (<= TILE_CAP; smaller only in tests of the spill path)
    __device__ __forceinline__ double frag(int t, int h, int lane) const
    {
        return t < cap ? Ls[t * 64 + h * 32 + lane] : __ldcg(Lg + (size_t)(t - cap) * 64 + h * 32 + lane);
    }
    __device__ __forceinline__ double at(int t, int r, int c) const
    {
        const int o = (c >> 2) * 32 + r * 4 + (c & 3);
        return t < cap ? Ls[t * 64 + o] : __ldcg(Lg + (size_t)(t - cap) * 64 + o);
    }
    // C-fragment (row l/4, columns 2(l%4), 2(l%4)+1) of an MMA result into fragment order
    __device__ __forceinline__ void store_c(int t, int lane, double x0, double x1) const
    {
        const int q = lane & 3, o = (q >> 1) * 32 + (lane >> 2) * 4 + 2 * (q & 1);
        double* p = t < cap ? Ls + t * 64 + o : Lg + (size_t)(t - cap) * 64 + o;
        *reinterpret_cast<double2*>(p) = make_double2(x0, x1);
    }
};

// ------------------------------------------------------------------------------------------------
// Cholesky of one 8x8 diagonal tile and the inverse of its factor, computed redundantly by every lane of a warp in
// registers (no shuffles: the chain per column is rsqrt -> mul -> fma).  S8: the updated tile, row-major, lower triangle.
// Columns >= ncol are padding and behave as identity columns (the right-hand-side row of the augmented system keeps its
// entries in the genuine columns).  Outputs: inverse of the factor in fragment order (the B operand of the TRSM MMAs and
// the diagonal solve of the back substitution); the factor itself row-major when Lout is given (last tile: it holds z).
__device__ __forceinline__ bool chol8_inv(const double* __restrict__ S8, int ncol, double* __restrict__ Linv, double* __restrict__ Lout)
{
    double a[8][8];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c2 = 0; c2 <= r / 2; c2++) {
            const double2 v = *reinterpret_cast<const double2*>(S8 + r * 8 + 2 * c2);
            a[r][2 * c2] = v.x; a[r][2 * c2 + 1] = v.y;
        }
    bool ok = true;
    double inv[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (k < ncol) {
            const double dk = a[k][k];
            ok = ok && (dk > 0.0);
            const double r = rsqrt(dk);
            inv[k] = r;
            a[k][k] = dk * r;
#pragma unroll
            for (int i = k + 1; i < 8; i++) a[i][k] *= r;
#pragma unroll
            for (int i = k + 1; i < 8; i++)
#pragma unroll
                for (int j = k + 1; j <= i; j++) a[i][j] = fma(-a[i][k], a[j][k], a[i][j]);
        } else {
            inv[k] = 1.0; a[k][k] = 1.0;
#pragma unroll
            for (int i = k + 1; i < 8; i++) a[i][k] = 0.0;
        }
    }
    if (Lout) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c2 = 0; c2 < 4; c2++)
                *reinterpret_cast<double2*>(Lout + r * 8 + 2 * c2) = make_double2(2 * c2 <= r ? a[r][2 * c2] : 0.0, 2 * c2 + 1 <= r ? a[r][2 * c2 + 1] : 0.0);
    }
    // inverse of the lower-triangular factor, row by row: li[r][c] = -(sum_{c<=m<r} l[r][m] li[m][c]) / l[r][r]
    double li[8][8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
        for (int c = 0; c < r; c++) {
            double s = 0.0;
#pragma unroll
            for (int m = c; m < r; m++) s = fma(a[r][m], li[m][c], s);
            li[r][c] = -s * inv[r];
        }
        li[r][r] = inv[r];
    }
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c2 = 0; c2 < 2; c2++) {
                const int c = 4 * h + 2 * c2;
                *reinterpret_cast<double2*>(Linv + h * 32 + r * 4 + 2 * c2) = make_double2(c <= r ? li[r][c] : 0.0, c + 1 <= r ? li[r][c + 1] : 0.0);
            }
    return ok;
}

// ------------------------------------------------------------------------------------------------
// Left-looking tiled Cholesky of the augmented system held as row-major 8x8 tiles in global memory (Ag, tile tix(I,J)),
// run by the whole CTA (ST_WARPS warps).  Block row I belongs to warp I % ST_WARPS; per panel J:
//   A  every owner finishes its tile (I,J): adds the term of panel J-1 to the partial sums it carried over, c = A_IJ - sum
//   B  the owner of the diagonal tile factors it (chol8_inv) while the other warps already accumulate the sums of panel
//      J+1 over the panels < J (software pipelining: the sqrt chain of the diagonal overlaps the bulk of the MMAs)
//   C  after a barrier: L_IJ = c L_JJ^-T as two MMAs against the published inverse
// nr = nc + 1 rows; n8 = ceil(nr / 8).  Linv: n8 tiles (fragment order), kept for the back substitution.  Ld: the factor of
// the last diagonal tile, row-major (it contains the tail of z).  Returns false (uniformly) if a pivot was not positive.
template <int R>
__device__ __forceinline__ bool chol_factor(const double* __restrict__ Ag, const TileStore& T, double* Linv, double* S8, double* Ld,
                                            int nc, int n8, int* s_fail)
{
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    double acc[R][4][2];      // partial sums of the NEXT panel: [row slot][k-half + 2 * panel parity][2]
    double cc[R][2];          // finished tiles of the CURRENT panel (C-fragments)
    double ar[R][2];          // A tiles of the next panel (prefetched)
    const int coff = (lane >> 2) * 8 + 2 * (lane & 3);
#pragma unroll
    for (int s = 0; s < R; s++) {
        const int row = w + ST_WARPS * s;
#pragma unroll
        for (int k = 0; k < 4; k++) acc[s][k][0] = acc[s][k][1] = 0.0;
        ar[s][0] = ar[s][1] = 0.0;
        if (row < n8) { const double2 v = __ldcg(reinterpret_cast<const double2*>(Ag + (size_t)tix(row, 0) * 64 + coff)); ar[s][0] = v.x; ar[s][1] = v.y; }
    }
    bool ok = true;
    for (int J = 0; J < n8; J++) {
        const int dw = J % ST_WARPS;
        // ---- A: finish the tiles of panel J ----
#pragma unroll
        for (int s = 0; s < R; s++) {
            const int row = w + ST_WARPS * s;
            if (row >= J && row < n8) {
                if (J > 0) {
                    const int ta = tix(row, J - 1), tb = tix(J, J - 1);
                    const double a0 = T.frag(ta, 0, lane), a1 = T.frag(ta, 1, lane), b0 = T.frag(tb, 0, lane), b1 = T.frag(tb, 1, lane);
                    dmma884(acc[s][0][0], acc[s][0][1], a0, b0);
                    dmma884(acc[s][1][0], acc[s][1][1], a1, b1);
                }
                cc[s][0] = ar[s][0] - ((acc[s][0][0] + acc[s][1][0]) + (acc[s][2][0] + acc[s][3][0]));
                cc[s][1] = ar[s][1] - ((acc[s][0][1] + acc[s][1][1]) + (acc[s][2][1] + acc[s][3][1]));
            }
        }
        // ---- B: diagonal factorisation on its owner; everyone starts on panel J+1 ----
        if (w == dw) {
            const int sd = J / ST_WARPS;
            double c0 = 0, c1 = 0;
#pragma unroll
            for (int s = 0; s < R; s++) if (s == sd) { c0 = cc[s][0]; c1 = cc[s][1]; }
            *reinterpret_cast<double2*>(S8 + coff) = make_double2(c0, c1);
            __syncwarp();
            const bool okd = chol8_inv(S8, min(8, nc - 8 * J), Linv + 64 * J, J == n8 - 1 ? Ld : nullptr);
            if (!okd && lane == 0) *s_fail = 1;
            __syncwarp();
        }
        if (J + 1 < n8) {
#pragma unroll
            for (int s = 0; s < R; s++) {
                const int row = w + ST_WARPS * s;
#pragma unroll
                for (int k = 0; k < 4; k++) acc[s][k][0] = acc[s][k][1] = 0.0;
                if (row > J && row < n8) { const double2 v = __ldcg(reinterpret_cast<const double2*>(Ag + (size_t)tix(row, J + 1) * 64 + coff)); ar[s][0] = v.x; ar[s][1] = v.y; }
            }
            const int tb0 = tix(J + 1, 0);
            for (int p = 0; p + 1 < J; p += 2) {              // panels p, p+1 < J: final
                const double b00 = T.frag(tb0 + p, 0, lane), b01 = T.frag(tb0 + p, 1, lane), b10 = T.frag(tb0 + p + 1, 0, lane), b11 = T.frag(tb0 + p + 1, 1, lane);
#pragma unroll
                for (int s = 0; s < R; s++) {
                    const int row = w + ST_WARPS * s;
                    if (row > J && row < n8) {
                        const int ta = tix(row, p);
                        const double a00 = T.frag(ta, 0, lane), a01 = T.frag(ta, 1, lane), a10 = T.frag(ta + 1, 0, lane), a11 = T.frag(ta + 1, 1, lane);
                        dmma884(acc[s][0][0], acc[s][0][1], a00, b00);
                        dmma884(acc[s][1][0], acc[s][1][1], a01, b01);
                        dmma884(acc[s][2][0], acc[s][2][1], a10, b10);
                        dmma884(acc[s][3][0], acc[s][3][1], a11, b11);
                    }
                }
            }
            if (J & 1) {                                        // odd number of final panels: the last one, p = J - 1
                const int p = J - 1;
                const double b00 = T.frag(tb0 + p, 0, lane), b01 = T.frag(tb0 + p, 1, lane);
#pragma unroll
                for (int s = 0; s < R; s++) {
                    const int row = w + ST_WARPS * s;
                    if (row > J && row < n8) {
                        const int ta = tix(row, p);
                        const double a00 = T.frag(ta, 0, lane), a01 = T.frag(ta, 1, lane);
                        dmma884(acc[s][2][0], acc[s][2][1], a00, b00);
                        dmma884(acc[s][3][0], acc[s][3][1], a01, b01);
                    }
                }
            }
        }
        __syncthreads();                                        // inverse of L_JJ published
        if (*reinterpret_cast<volatile int*>(s_fail)) { ok = false; break; }
        // ---- C: L_IJ = c L_JJ^-T ----
        {
            const double b0 = Linv[64 * J + lane], b1 = Linv[64 * J + 32 + lane];
#pragma unroll
            for (int s = 0; s < R; s++) {
                const int row = w + ST_WARPS * s;
                if (row > J && row < n8) {
                    // C-fragment -> A-fragments: lane (row r, k) needs column 4h+k, held by lane 4r + 2h + k/2, element k%2
                    const int src0 = (lane & ~3) + ((lane & 3) >> 1), src1 = src0 + 2;
                    const double v00 = __shfl_sync(0xffffffffu, cc[s][0], src0), v01 = __shfl_sync(0xffffffffu, cc[s][1], src0);
                    const double v10 = __shfl_sync(0xffffffffu, cc[s][0], src1), v11 = __shfl_sync(0xffffffffu, cc[s][1], src1);
                    const double a0 = (lane & 1) ? v01 : v00, a1 = (lane & 1) ? v11 : v10;
                    double x0 = 0.0, x1 = 0.0;
                    dmma884(x0, x1, a0, b0);
                    dmma884(x0, x1, a1, b1);
                    T.store_c(tix(row, J), lane, x0, x1);
                }
            }
        }
        __syncthreads();                                        // panel J published
    }
    return ok;
}

// ------------------------------------------------------------------------------------------------
// Back substitution  L^T y = z  over the leading nc x nc part of the factor, by ONE warp.  z = row nc of the factor.
// Lane (c & 31) keeps z_c / the running right-hand side of column c in a register (slot c / 32).  Blocks are solved from
// the bottom: y_J = L_JJ^-T rhs_J through the stored inverse (no dependent divide chain), then every lane subtracts
// L_{J,K}^T y_J from its columns of the blocks K < J.
__device__ __noinline__ void chol_backsubst(const TileStore& T, const double* __restrict__ Linv, const double* __restrict__ Ld,
                                            double* __restrict__ y, int nc, int lane)
{
    constexpr int NS = (MAX_NC + 31) / 32;
    const int ir = nc >> 3, rr = nc & 7;              // tile row / local row of the right-hand-side row
    double z[NS];
#pragma unroll
    for (int q = 0; q < NS; q++) {
        const int c = 32 * q + lane;
        z[q] = 0.0;
        if (c < nc) z[q] = (c >> 3) == ir ? Ld[rr * 8 + (c & 7)] : T.at(tix(ir, c >> 3), rr, c & 7);
    }
    const int jtop = (nc - 1) >> 3;
#pragma unroll
    for (int q = NS - 1; q >= 0; q--) {
        if (32 * q < nc) {
            for (int jj = min(3, jtop - 4 * q); jj >= 0; jj--) {
                const int J = 4 * q + jj;
                // rhs of block J to every lane, y_J = Linv_J^T rhs (rows r >= k; rows beyond nc carry zeros in z)
                double rh[8], yj[8];
#pragma unroll
                for (int r = 0; r < 8; r++) rh[r] = __shfl_sync(0xffffffffu, z[q], 8 * jj + r);
                const double* Li = Linv + 64 * J;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    double s = 0.0;
#pragma unroll
                    for (int r = k; r < 8; r++) s = fma(Li[(k >> 2) * 32 + r * 4 + (k & 3)], (8 * J + r < nc) ? rh[r] : 0.0, s);
                    yj[k] = s;
                }
                if (lane < 8 && 8 * J + lane < nc) {
                    double v = yj[0];
#pragma unroll
                    for (int k = 1; k < 8; k++) v = lane == k ? yj[k] : v;
                    y[8 * J + lane] = v;
                }
                // columns of the blocks K < J
#pragma unroll
                for (int q2 = 0; q2 <= q; q2++) {
                    const int c = 32 * q2 + lane;
                    if (c < 8 * J) {
                        const int t = tix(J, c >> 3);
                        double s = z[q2];
#pragma unroll
                        for (int r = 0; r < 8; r++) s = fma(-T.at(t, r, c & 7), (8 * J + r < nc) ? yj[r] : 0.0, s);
                        z[q2] = s;
                    }
                }
            }
        }
    }
}

}  // namespace gfba
