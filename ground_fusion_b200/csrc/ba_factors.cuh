// ba_factors.cuh -- FP64 device restatement of the reference's analytic factors used by
// Estimator::optimization() (estimator.cpp:3102-3297):
//   ProjectionTwoFrameOneCamFactor::Evaluate   factor/projectionTwoFrameOneCamFactor.cpp:43-151
//   IMUFactor::Evaluate + IntegrationBase::evaluate   factor/imu_factor.h:28-191, integration_base.h:169-195
//   MarginalizationFactor::Evaluate            factor/marginalization_factor.cpp:344-392
//   PoseLocalParameterization::Plus            factor/pose_local_parameterization.cpp:12-26
// Same formulas and operation order as oracle/ba_oracle.c (the CPU oracle the parity tests compare with).
// Quaternions are stored x,y,z,w.  Compiled with -fmad=false like the rest of the library.
#pragma once
#include "gf_common.cuh"

namespace gfba {

__device__ __forceinline__ void m3_mul(const double* a, const double* b, double* c)
{
    double t[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) t[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
#pragma unroll
    for (int i = 0; i < 9; i++) c[i] = t[i];
}
__device__ __forceinline__ void m3_T(const double* a, double* c)
{
    double t[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) t[i * 3 + j] = a[j * 3 + i];
#pragma unroll
    for (int i = 0; i < 9; i++) c[i] = t[i];
}
__device__ __forceinline__ void m3_v(const double* a, const double* b, double* c)
{
    double t0 = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    double t1 = a[3] * b[0] + a[4] * b[1] + a[5] * b[2];
    double t2 = a[6] * b[0] + a[7] * b[1] + a[8] * b[2];
    c[0] = t0; c[1] = t1; c[2] = t2;
}
__device__ __forceinline__ void skew(const double* q, double* s)
{
    s[0] = 0; s[1] = -q[2]; s[2] = q[1];
    s[3] = q[2]; s[4] = 0; s[5] = -q[0];
    s[6] = -q[1]; s[7] = q[0]; s[8] = 0;
}
__device__ __forceinline__ void q_mul(const double* a, const double* b, double* c)
{
    double t3 = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    double t0 = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    double t1 = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    double t2 = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    c[0] = t0; c[1] = t1; c[2] = t2; c[3] = t3;
}
__device__ __forceinline__ void q_inv(const double* a, double* c)
{
    double n2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];
    c[0] = -a[0] / n2; c[1] = -a[1] / n2; c[2] = -a[2] / n2; c[3] = a[3] / n2;
}
__device__ __forceinline__ void q_normalize(double* a)
{
    double n = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]);
    a[0] /= n; a[1] /= n; a[2] /= n; a[3] /= n;
}
__device__ __forceinline__ void q_to_R(const double* q, double* R)
{
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
__device__ __forceinline__ void q_rot(const double* q, const double* v, double* out)
{
    double R[9];
    q_to_R(q, R);
    m3_v(R, v, out);
}
__device__ __forceinline__ void delta_q(const double* theta, double* dq)
{   // Utility::deltaQ (utility/utility.h:23-36)
    dq[0] = theta[0] / 2.0; dq[1] = theta[1] / 2.0; dq[2] = theta[2] / 2.0; dq[3] = 1.0;
    q_normalize(dq);
}
__device__ __forceinline__ void q_left_br(const double* q, double* out)
{
    double s[9]; skew(q, s);
#pragma unroll
    for (int i = 0; i < 9; i++) out[i] = s[i];
    out[0] += q[3]; out[4] += q[3]; out[8] += q[3];
}
__device__ __forceinline__ void q_right_br(const double* q, double* out)
{
    double s[9]; skew(q, s);
#pragma unroll
    for (int i = 0; i < 9; i++) out[i] = -s[i];
    out[0] += q[3]; out[4] += q[3]; out[8] += q[3];
}
__device__ __forceinline__ void pose_plus(const double* x, const double* d, double* out)
{
    out[0] = x[0] + d[0]; out[1] = x[1] + d[1]; out[2] = x[2] + d[2];
    double dq[4], q[4];
    delta_q(d + 3, dq); q_mul(x + 3, dq, q); q_normalize(q);
    out[3] = q[0]; out[4] = q[1]; out[5] = q[2]; out[6] = q[3];
}

// Visual factor.  J = [Ji(2x6) | Jj(2x6) | Jex(2x6) | Jtd(2x1) | Jf(2x1)] as 2 rows of 20 doubles; raw (no loss).
// want_jac = false: residual only.
__device__ __forceinline__ void eval_visual(const gf_ba_visual_factor& f, double sqrt_info, const double* pose_i, const double* pose_j,
                                            const double* ex, double inv_dep, double td, double* res, double* J, bool want_jac)
{
    const double *Pi = pose_i, *Qi = pose_i + 3, *Pj = pose_j, *Qj = pose_j + 3, *tic = ex, *qic = ex + 3;
    double vi[3] = {f.vel_i[0], f.vel_i[1], 0}, vj[3] = {f.vel_j[0], f.vel_j[1], 0};
    double pts_i_td[3], pts_j_td[3], pc_i[3], pimu_i[3], pw[3], pimu_j[3], pc_j[3], t[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { pts_i_td[k] = f.pts_i[k] - (td - f.td_i) * vi[k]; pts_j_td[k] = f.pts_j[k] - (td - f.td_j) * vj[k]; }
#pragma unroll
    for (int k = 0; k < 3; k++) pc_i[k] = pts_i_td[k] / inv_dep;
    double Ri[9], Rj[9], ric[9], RjT[9], ricT[9];
    q_to_R(Qi, Ri); q_to_R(Qj, Rj); q_to_R(qic, ric); m3_T(Rj, RjT); m3_T(ric, ricT);
    m3_v(ric, pc_i, pimu_i);
#pragma unroll
    for (int k = 0; k < 3; k++) pimu_i[k] += tic[k];
    m3_v(Ri, pimu_i, pw);
#pragma unroll
    for (int k = 0; k < 3; k++) { pw[k] += Pi[k]; t[k] = pw[k] - Pj[k]; }
    m3_v(RjT, t, pimu_j);
#pragma unroll
    for (int k = 0; k < 3; k++) t[k] = pimu_j[k] - tic[k];
    m3_v(ricT, t, pc_j);
    double dep_j = pc_j[2];
    res[0] = sqrt_info * (pc_j[0] / dep_j - pts_j_td[0]);
    res[1] = sqrt_info * (pc_j[1] / dep_j - pts_j_td[1]);
    if (!want_jac) return;
    double reduce[6] = {1. / dep_j, 0, -pc_j[0] / (dep_j * dep_j), 0, 1. / dep_j, -pc_j[1] / (dep_j * dep_j)};
#pragma unroll
    for (int k = 0; k < 6; k++) reduce[k] *= sqrt_info;
    double A[9], B[9], S[9], C[9];
    m3_mul(ricT, RjT, A);
    {   // pose i
        m3_mul(A, Ri, B); skew(pimu_i, S); m3_mul(B, S, C);
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                J[r * 20 + c] = reduce[r * 3] * A[c] + reduce[r * 3 + 1] * A[3 + c] + reduce[r * 3 + 2] * A[6 + c];
                J[r * 20 + 3 + c] = reduce[r * 3] * -C[c] + reduce[r * 3 + 1] * -C[3 + c] + reduce[r * 3 + 2] * -C[6 + c];
            }
    }
    {   // pose j
        skew(pimu_j, S); m3_mul(ricT, S, C);
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                J[r * 20 + 6 + c] = reduce[r * 3] * -A[c] + reduce[r * 3 + 1] * -A[3 + c] + reduce[r * 3 + 2] * -A[6 + c];
                J[r * 20 + 9 + c] = reduce[r * 3] * C[c] + reduce[r * 3 + 1] * C[3 + c] + reduce[r * 3 + 2] * C[6 + c];
            }
    }
    double tmp_r[9];
    m3_mul(A, Ri, B); m3_mul(B, ric, tmp_r);
    {   // extrinsic
        double RjTRi[9], M[9], L[9], S1[9], T1[9], S2[9], S3[9], u[3], w2[3], x[3];
        m3_mul(RjT, Ri, RjTRi);
#pragma unroll
        for (int k = 0; k < 9; k++) M[k] = RjTRi[k];
        M[0] -= 1; M[4] -= 1; M[8] -= 1;
        m3_mul(ricT, M, L);
        skew(pc_i, S1); m3_mul(tmp_r, S1, T1);
        m3_v(tmp_r, pc_i, u); skew(u, S2);
        m3_v(Ri, tic, w2);
#pragma unroll
        for (int k = 0; k < 3; k++) w2[k] += Pi[k] - Pj[k];
        m3_v(RjT, w2, x);
#pragma unroll
        for (int k = 0; k < 3; k++) x[k] -= tic[k];
        m3_v(ricT, x, w2); skew(w2, S3);
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                J[r * 20 + 12 + c] = reduce[r * 3] * L[c] + reduce[r * 3 + 1] * L[3 + c] + reduce[r * 3 + 2] * L[6 + c];
                double e0 = -T1[c] + S2[c] + S3[c], e1 = -T1[3 + c] + S2[3 + c] + S3[3 + c], e2 = -T1[6 + c] + S2[6 + c] + S3[6 + c];
                J[r * 20 + 15 + c] = reduce[r * 3] * e0 + reduce[r * 3 + 1] * e1 + reduce[r * 3 + 2] * e2;
            }
    }
    {   // td and inverse depth
        double u[3], v[3];
        m3_v(tmp_r, vi, u); m3_v(tmp_r, pts_i_td, v);
#pragma unroll
        for (int r = 0; r < 2; r++) {
            J[r * 20 + 18] = (reduce[r * 3] * u[0] + reduce[r * 3 + 1] * u[1] + reduce[r * 3 + 2] * u[2]) / inv_dep * -1.0 + sqrt_info * vj[r];
            J[r * 20 + 19] = (reduce[r * 3] * v[0] + reduce[r * 3 + 1] * v[1] + reduce[r * 3 + 2] * v[2]) * -1.0 / (inv_dep * inv_dep);
        }
    }
}

// HuberLoss(1.0) + Ceres Corrector with rho'' <= 0 (in-tree restatement marginalization_factor.cpp:46-77):
// returns rho(s); scale = sqrt(rho'(s)) multiplies residual and Jacobian.
__device__ __forceinline__ double huber(double sq, double& scale)
{
    if (sq > 1.0) {
        double r = sqrt(sq), rho1 = 1.0 / r;
        if (rho1 < 2.2250738585072014e-308) rho1 = 2.2250738585072014e-308;
        scale = sqrt(rho1);
        return 2.0 * r - 1.0;
    }
    scale = 1.0;
    return sq;
}

constexpr int O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12;

// IMU residual (15) before sqrt_info, plus the quantities the Jacobians need.  J (15 x 30, row-major, columns
// [pose_i 6 | sb_i 9 | pose_j 6 | sb_j 9], raw = before sqrt_info) is filled when want_jac.
__device__ inline void eval_imu_raw(const gf_ba_imu_factor& f, const double* G, const double* pose_i, const double* sb_i,
                                    const double* pose_j, const double* sb_j, double* r, double* J, bool want_jac)
{
    const double *Pi = pose_i, *Qi = pose_i + 3, *Vi = sb_i, *Bai = sb_i + 3, *Bgi = sb_i + 6;
    const double *Pj = pose_j, *Qj = pose_j + 3, *Vj = sb_j, *Baj = sb_j + 3, *Bgj = sb_j + 6;
    auto blk = [&](int r0, int c0, double* out) {
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) out[a * 3 + b] = f.jacobian[(r0 + a) * 15 + c0 + b];
    };
    double dp_dba[9], dp_dbg[9], dq_dbg[9], dv_dba[9], dv_dbg[9];
    blk(O_P, O_BA, dp_dba); blk(O_P, O_BG, dp_dbg); blk(O_R, O_BG, dq_dbg); blk(O_V, O_BA, dv_dba); blk(O_V, O_BG, dv_dbg);
    double dba[3], dbg[3], t[3], u[3];
    for (int k = 0; k < 3; k++) { dba[k] = Bai[k] - f.linearized_ba[k]; dbg[k] = Bgi[k] - f.linearized_bg[k]; }
    double dq[4], corr_q[4], Qi_inv[4], tq[4], tq2[4];
    m3_v(dq_dbg, dbg, t); delta_q(t, dq); q_mul(f.delta_q, dq, corr_q);
    double corr_v[3], corr_p[3];
    m3_v(dv_dba, dba, t); m3_v(dv_dbg, dbg, u);
    for (int k = 0; k < 3; k++) corr_v[k] = f.delta_v[k] + t[k] + u[k];
    m3_v(dp_dba, dba, t); m3_v(dp_dbg, dbg, u);
    for (int k = 0; k < 3; k++) corr_p[k] = f.delta_p[k] + t[k] + u[k];
    q_inv(Qi, Qi_inv);
    const double dt = f.sum_dt;
    double a[3], b[3];
    for (int k = 0; k < 3; k++) a[k] = 0.5 * G[k] * dt * dt + Pj[k] - Pi[k] - Vi[k] * dt;
    q_rot(Qi_inv, a, t);
    for (int k = 0; k < 3; k++) r[O_P + k] = t[k] - corr_p[k];
    q_inv(corr_q, tq); q_mul(Qi_inv, Qj, tq2); q_mul(tq, tq2, tq);
    for (int k = 0; k < 3; k++) r[O_R + k] = 2 * tq[k];
    for (int k = 0; k < 3; k++) b[k] = G[k] * dt + Vj[k] - Vi[k];
    q_rot(Qi_inv, b, u);
    for (int k = 0; k < 3; k++) r[O_V + k] = u[k] - corr_v[k];
    for (int k = 0; k < 3; k++) { r[O_BA + k] = Baj[k] - Bai[k]; r[O_BG + k] = Bgj[k] - Bgi[k]; }
    if (!want_jac) return;
    for (int k = 0; k < 450; k++) J[k] = 0.0;
    auto set_blk = [&](int r0, int c0, const double* M, double s) {
        for (int x = 0; x < 3; x++) for (int y = 0; y < 3; y++) J[(r0 + x) * 30 + c0 + y] = s * M[x * 3 + y];
    };
    double RiT[9], S[9], M[9], N[9];
    q_to_R(Qi_inv, RiT);
    // pose_i: columns 0..5
    set_blk(O_P, 0 + O_P, RiT, -1.0);
    skew(t, S); set_blk(O_P, 0 + O_R, S, 1.0);          // t = Qi^-1 (0.5 G dt^2 + Pj - Pi - Vi dt)
    {
        double qji[4]; q_inv(Qj, tq); q_mul(tq, Qi, qji);
        q_left_br(qji, M); q_right_br(corr_q, N);
        double L4[16], R4[16];
        L4[0] = qji[3]; L4[1] = -qji[0]; L4[2] = -qji[1]; L4[3] = -qji[2];
        for (int rr = 0; rr < 3; rr++) { L4[(rr + 1) * 4] = qji[rr]; for (int cc = 0; cc < 3; cc++) L4[(rr + 1) * 4 + 1 + cc] = M[rr * 3 + cc]; }
        R4[0] = corr_q[3]; R4[1] = -corr_q[0]; R4[2] = -corr_q[1]; R4[3] = -corr_q[2];
        for (int rr = 0; rr < 3; rr++) { R4[(rr + 1) * 4] = corr_q[rr]; for (int cc = 0; cc < 3; cc++) R4[(rr + 1) * 4 + 1 + cc] = N[rr * 3 + cc]; }
        for (int rr = 0; rr < 3; rr++)
            for (int cc = 0; cc < 3; cc++) {
                double s = 0;
                for (int k = 0; k < 4; k++) s += L4[(rr + 1) * 4 + k] * R4[k * 4 + 1 + cc];
                J[(O_R + rr) * 30 + 0 + O_R + cc] = -s;
            }
    }
    skew(u, S); set_blk(O_V, 0 + O_R, S, 1.0);          // u = Qi^-1 (G dt + Vj - Vi)
    // speed/bias i: columns 6..14
    set_blk(O_P, 6 + 0, RiT, -dt);
    set_blk(O_P, 6 + 3, dp_dba, -1.0);
    set_blk(O_P, 6 + 6, dp_dbg, -1.0);
    {
        double q3[4]; q_inv(Qj, tq); q_mul(tq, Qi, q3); q_mul(q3, f.delta_q, q3);
        q_left_br(q3, M); m3_mul(M, dq_dbg, N);
        set_blk(O_R, 6 + 6, N, -1.0);
    }
    set_blk(O_V, 6 + 0, RiT, -1.0);
    set_blk(O_V, 6 + 3, dv_dba, -1.0);
    set_blk(O_V, 6 + 6, dv_dbg, -1.0);
    for (int k = 0; k < 3; k++) { J[(O_BA + k) * 30 + 6 + 3 + k] = -1.0; J[(O_BG + k) * 30 + 6 + 6 + k] = -1.0; }
    // pose_j: columns 15..20
    set_blk(O_P, 15 + O_P, RiT, 1.0);
    {
        double q3[4]; q_inv(corr_q, tq); q_mul(tq, Qi_inv, q3); q_mul(q3, Qj, q3);
        q_left_br(q3, M);
        set_blk(O_R, 15 + O_R, M, 1.0);
    }
    // speed/bias j: columns 21..29
    set_blk(O_V, 21 + 0, RiT, 1.0);
    for (int k = 0; k < 3; k++) { J[(O_BA + k) * 30 + 21 + 3 + k] = 1.0; J[(O_BG + k) * 30 + 21 + 6 + k] = 1.0; }
}

// sqrt_info = LLT(cov^-1).matrixL().transpose() for an n x n covariance (n <= 15); scratch M is n x 2n.
__device__ inline bool sqrt_info_from_cov(const double* cov, int n, double* out, double* M)
{
    const int w2 = 2 * n;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { M[i * w2 + j] = cov[i * n + j]; M[i * w2 + n + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < n; c++) {           // Gauss-Jordan with partial pivoting (same as the oracle)
        int piv = c;
        for (int r = c + 1; r < n; r++) if (fabs(M[r * w2 + c]) > fabs(M[piv * w2 + c])) piv = r;
        if (M[piv * w2 + c] == 0.0) return false;
        if (piv != c) for (int j = 0; j < w2; j++) { double t = M[c * w2 + j]; M[c * w2 + j] = M[piv * w2 + j]; M[piv * w2 + j] = t; }
        double d = M[c * w2 + c];
        for (int j = 0; j < w2; j++) M[c * w2 + j] /= d;
        for (int r = 0; r < n; r++) if (r != c) {
            double fct = M[r * w2 + c];
            if (fct != 0.0) for (int j = 0; j < w2; j++) M[r * w2 + j] -= fct * M[c * w2 + j];
        }
    }
    // Cholesky (lower) of the inverse, in place in the right half
    double* A = M + n;                       // element (i,j) at A[i*w2 + j]
    for (int j = 0; j < n; j++) {
        double d = A[j * w2 + j];
        for (int k = 0; k < j; k++) d -= A[j * w2 + k] * A[j * w2 + k];
        if (!(d > 0.0)) return false;
        d = sqrt(d);
        A[j * w2 + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = A[i * w2 + j];
            for (int k = 0; k < j; k++) s -= A[i * w2 + k] * A[j * w2 + k];
            A[i * w2 + j] = s / d;
        }
    }
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) out[i * n + j] = (j >= i) ? A[j * w2 + i] : 0.0;
    return true;
}

}  // namespace gfba
