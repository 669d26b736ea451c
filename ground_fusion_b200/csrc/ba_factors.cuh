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

// ------------------------------------------------------------------------------------------------
// Wheel odometry factor.  SO3 exp/log as in Sophus (upstream, un-vendored), right Jacobians as vendored in the reference's
// utility/sophus_utils.hpp:155-236; the arithmetic mirrors oracle/ba_oracle.c (gfo_eval_wheel).
constexpr double kSophusEps = 1e-10;
constexpr double kPi = 3.14159265358979323846;
__device__ inline void so3_exp_q(const double* w, double* q)
{
    const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double imag, real;
    if (t2 < kSophusEps * kSophusEps) {
        const double t4 = t2 * t2;
        imag = 0.5 - (1.0 / 48.0) * t2 + (1.0 / 3840.0) * t4;
        real = 1.0 - (1.0 / 8.0) * t2 + (1.0 / 384.0) * t4;
    } else {
        const double t = sqrt(t2), h = 0.5 * t;
        imag = sin(h) / t; real = cos(h);
    }
    q[0] = imag * w[0]; q[1] = imag * w[1]; q[2] = imag * w[2]; q[3] = real;
}
__device__ inline void so3_log_q(const double* qin, double* out)
{
    double q[4] = {qin[0], qin[1], qin[2], qin[3]};
    q_normalize(q);
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2], w = q[3];
    double f;
    if (n2 < kSophusEps * kSophusEps) f = 2.0 / w - (2.0 / 3.0) * n2 / (w * w * w);
    else {
        const double n = sqrt(n2);
        if (fabs(w) < kSophusEps) f = (w > 0 ? kPi : -kPi) / n;
        else f = 2.0 * atan(n / w) / n;
    }
    out[0] = f * q[0]; out[1] = f * q[1]; out[2] = f * q[2];
}
__device__ inline void so3_exp_R(const double* w, double* R) { double q[4]; so3_exp_q(w, q); q_to_R(q, R); }
__device__ inline void so3_Jr(const double* phi, double* J)
{
    const double n2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
    double h[9], h2[9]; skew(phi, h); m3_mul(h, h, h2);
    double a, b;
    if (n2 > kSophusEps) { const double n = sqrt(n2); a = (1.0 - cos(n)) / n2; b = (n - sin(n)) / (n2 * n); }
    else { a = 0.5; b = 1.0 / 6.0; }
    for (int k = 0; k < 9; k++) J[k] = ((k % 4 == 0) ? 1.0 : 0.0) - a * h[k] + b * h2[k];
}
__device__ inline void so3_Jr_inv(const double* phi, double* J)
{
    const double n2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
    double h[9], h2[9]; skew(phi, h); m3_mul(h, h, h2);
    double b;
    if (n2 > kSophusEps) {
        const double n = sqrt(n2);
        if (n < kPi - 1e-5) b = 1.0 / n2 - (1.0 + cos(n)) / (2.0 * n * sin(n));
        else b = 1.0 / (kPi * kPi);
    } else b = 1.0 / 12.0;
    for (int k = 0; k < 9; k++) J[k] = ((k % 4 == 0) ? 1.0 : 0.0) + 0.5 * h[k] + b * h2[k];
}

constexpr int WHEEL_COLS = 22;   // local columns of one wheel factor: pose_i 6 | pose_j 6 | ex_wheel 6 | sx | sy | sw | td
// WheelFactor::Evaluate (reference factor/wheel_factor.h:28-247) + WheelIntegrationBase::evaluate
// (wheel_integration_base.h:179-219).  res[6] and J[6][22] come out multiplied by sqrt_info = LLT(cov^-1).L^T.
// M: scratch of 6*12 doubles.  Returns false if the covariance is not positive definite.
__device__ __noinline__ bool eval_wheel(const gf_ba_wheel_factor& f, const double* pose_i, const double* pose_j, const double* exw,
                                  double sx, double sy, double sw, double td, double* res, double* J, bool jac, double* M)
{
    const double* Pi = pose_i; const double* Qi = pose_i + 3; const double* Pj = pose_j; const double* Qj = pose_j + 3;
    const double* tio = exw; const double* qio = exw + 3;
    double dp_dsx[3], dp_dsy[3], dp_dsw[3], dq_dsw[3];
    for (int k = 0; k < 3; k++) { dp_dsx[k] = f.jacobian[k * 3]; dp_dsy[k] = f.jacobian[k * 3 + 1]; dp_dsw[k] = f.jacobian[k * 3 + 2]; dq_dsw[k] = f.jacobian[(3 + k) * 3 + 2]; }
    const double dsx = sx - f.linearized_sx, dsy = sy - f.linearized_sy, dsw = sw - f.linearized_sw, dtd = td - f.linearized_td;
    const double sv[3] = {sx, sy, 1.0};
    double Ri[9], Rj[9], rio[9]; q_to_R(Qi, Ri); q_to_R(Qj, Rj); q_to_R(qio, rio);
    double cp[3]; for (int k = 0; k < 3; k++) cp[k] = f.delta_p[k] + dp_dsx[k] * dsx + dp_dsy[k] * dsy + dp_dsw[k] * dsw;
    double dq0[4] = {f.delta_q[0], f.delta_q[1], f.delta_q[2], f.delta_q[3]}, e[4], cq[4], t3[3];
    q_normalize(dq0);
    for (int k = 0; k < 3; k++) t3[k] = dq_dsw[k] * dsw;
    so3_exp_q(t3, e); q_mul(dq0, e, cq); q_normalize(cq);
    double Rcq[9]; q_to_R(cq, Rcq);
    double fcw[3], fcv[3], bcv[3], bcw[3], nbcw[3];
    for (int k = 0; k < 3; k++) { fcw[k] = sw * f.linearized_gyr[k] * dtd; fcv[k] = sv[k] * f.linearized_vel[k] * dtd; bcv[k] = sv[k] * f.vel_1[k] * dtd; bcw[k] = sw * f.gyr_1[k] * dtd; nbcw[k] = -bcw[k]; }
    double E1[4], E2[4], qt[4], dqt[4]; so3_exp_q(fcw, E1); so3_exp_q(nbcw, E2);
    q_mul(E1, cq, qt); q_mul(qt, E2, dqt); q_normalize(dqt);
    double RE1[9]; q_to_R(E1, RE1);
    double u[3], inner[3], dpt[3]; m3_v(Rcq, bcv, u);
    for (int k = 0; k < 3; k++) inner[k] = fcv[k] + cp[k] - u[k];
    m3_v(RE1, inner, dpt);
    double Rio[9], RioT[9]; m3_mul(Ri, rio, Rio); m3_T(Rio, RioT);
    double a1[3], a2[3], dw[3], dpos[3]; m3_v(Rj, tio, a1); m3_v(Ri, tio, a2);
    for (int k = 0; k < 3; k++) dw[k] = a1[k] + Pj[k] - a2[k] - Pi[k];
    m3_v(RioT, dw, dpos);
    double raw[6];
    for (int k = 0; k < 3; k++) raw[k] = dpos[k] - dpt[k];
    double qiqio[4], inv1[4], inv2[4], qjqio[4], t1[4], t2[4];
    q_mul(Qi, qio, qiqio); q_inv(qiqio, inv1); q_inv(dqt, inv2); q_mul(Qj, qio, qjqio);
    q_mul(inv2, inv1, t1); q_mul(t1, qjqio, t2);
    double rq[3]; so3_log_q(t2, rq);
    for (int k = 0; k < 3; k++) raw[3 + k] = rq[k];
    double U[36];
    if (!sqrt_info_from_cov(f.covariance, 6, U, M)) return false;
    for (int i = 0; i < 6; i++) { double v = 0; for (int k = 0; k < 6; k++) v += U[i * 6 + k] * raw[k]; res[i] = v; }
    if (!jac) return true;
    double R[6 * WHEEL_COLS];                         // raw Jacobian
    for (int k = 0; k < 6 * WHEEL_COLS; k++) R[k] = 0.0;
    auto put = [&](int r0, int c0, const double* B, double sgn) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[(r0 + r) * WHEEL_COLS + c0 + c] = sgn * B[r * 3 + c]; };
    double Jri[9]; so3_Jr_inv(rq, Jri);
    double drdsw[3]; for (int k = 0; k < 3; k++) drdsw[k] = dq_dsw[k] * dsw;
    double Jr_drdsw[9]; so3_Jr(drdsw, Jr_drdsw);
    double S[9], A[9], B[9], C[9], RiT[9], rioT[9]; m3_T(Ri, RiT); m3_T(rio, rioT);
    {   // pose_i
        put(0, 0, RioT, -1.0);
        skew(tio, S); m3_mul(Ri, S, A); m3_mul(RioT, A, B);
        double w1[3]; m3_v(RiT, dw, w1); skew(w1, S); m3_mul(rioT, S, C);
        for (int k = 0; k < 9; k++) B[k] += C[k];
        put(0, 3, B, 1.0);
        double qa[4], qb[4]; q_inv(qjqio, qa); q_mul(qa, Qi, qb); q_to_R(qb, A); m3_mul(Jri, A, B);
        put(3, 3, B, -1.0);
    }
    {   // pose_j
        put(0, 6, RioT, 1.0);
        skew(tio, S); m3_mul(RioT, Rj, A); m3_mul(A, S, B);
        put(0, 9, B, -1.0);
        m3_mul(Jri, rioT, B);
        put(3, 9, B, 1.0);
    }
    {   // wheel extrinsic
        for (int k = 0; k < 9; k++) A[k] = Rj[k] - Ri[k];
        m3_mul(RioT, A, B);
        put(0, 12, B, 1.0);
        skew(dpos, S);
        put(0, 15, S, 1.0);
        double qa[4], qb[4], qc[4]; q_inv(qjqio, qa); q_mul(qa, Qi, qb); q_mul(qb, qio, qc); q_to_R(qc, A);
        for (int k = 0; k < 9; k++) A[k] = ((k % 4 == 0) ? 1.0 : 0.0) - A[k];
        m3_mul(Jri, A, B);
        put(3, 15, B, 1.0);
    }
    double Jrtd[9], Jrmtd[9], Rfcv[9], Rfcw[9], Rnr[9], Rbcw[9], RcqT[9], nfcw[3], nrq[3];
    for (int k = 0; k < 3; k++) { nfcw[k] = -fcw[k]; nrq[k] = -rq[k]; }
    so3_Jr(fcw, Jrtd); so3_Jr(nfcw, Jrmtd);
    so3_exp_R(fcv, Rfcv); so3_exp_R(fcw, Rfcw); so3_exp_R(nrq, Rnr); so3_exp_R(bcw, Rbcw); m3_T(Rcq, RcqT);
    for (int axis = 0; axis < 2; axis++) {            // sx, sy (the reference rotates by exp(forward_compensate_v) here)
        const double* dpds = axis == 0 ? dp_dsx : dp_dsy;
        double e1[3] = {0, 0, 0}, e2[3] = {0, 0, 0}, w1[3], w2[3], w3[3];
        e1[axis] = f.linearized_vel[axis] * dtd;
        e2[axis] = f.vel_1[axis] * dtd;
        m3_v(Rcq, e2, w1);
        for (int k = 0; k < 3; k++) w2[k] = e1[k] + dpds[k] - w1[k];
        m3_v(Rfcv, w2, w3);
        for (int k = 0; k < 3; k++) R[k * WHEEL_COLS + 18 + axis] = -w3[k];
    }
    {   // sw
        double w1[3], w2[3], w3[3], w4[3], w5[3], lg[3], svv1[3], tot[3], outp[3], r1[3], r2[3], r3[3], r4[3], r5[3];
        m3_v(Jr_drdsw, dq_dsw, w1); skew(w1, S);
        for (int k = 0; k < 3; k++) svv1[k] = sv[k] * f.vel_1[k] * dtd;
        m3_v(S, svv1, w2); m3_v(Rcq, w2, w3);
        for (int k = 0; k < 3; k++) lg[k] = f.linearized_gyr[k] * dtd;
        m3_v(Jrtd, lg, w4); skew(w4, S); m3_v(S, inner, w5);
        for (int k = 0; k < 3; k++) tot[k] = dp_dsw[k] - w3[k] + w5[k];
        m3_v(Rfcw, tot, outp);
        m3_v(RcqT, w4, r1);
        for (int k = 0; k < 3; k++) r2[k] = r1[k] + w1[k];
        m3_v(Rbcw, r2, r3); m3_v(Rnr, r3, r4); m3_v(Jri, r4, r5);
        for (int k = 0; k < 3; k++) { R[k * WHEEL_COLS + 20] = -outp[k]; R[(3 + k) * WHEEL_COLS + 20] = -r5[k]; }
    }
    {   // td
        double w1[3], w2[3], w3[3], w4[3], w5[3], svl[3], svv[3], swg[3], swg1[3], r1[3], r2[3], r3[3], r4[3], r5[3], r6[3];
        for (int k = 0; k < 3; k++) { svl[k] = sv[k] * f.linearized_vel[k]; svv[k] = sv[k] * f.vel_1[k]; swg[k] = sw * f.linearized_gyr[k]; swg1[k] = sw * f.gyr_1[k]; }
        m3_v(Rcq, svv, w1);
        m3_v(Jrtd, swg, w2); skew(w2, S); m3_v(S, inner, w3);
        for (int k = 0; k < 3; k++) w4[k] = svl[k] - w1[k] + w3[k];
        m3_v(Rfcw, w4, w5);
        m3_v(RcqT, w2, r1); m3_v(Rbcw, r1, r2);
        m3_v(Jrmtd, swg1, r3);
        for (int k = 0; k < 3; k++) r4[k] = r2[k] - r3[k];
        m3_v(Rnr, r4, r5); m3_v(Jri, r5, r6);
        for (int k = 0; k < 3; k++) { R[k * WHEEL_COLS + 21] = -w5[k]; R[(3 + k) * WHEEL_COLS + 21] = -r6[k]; }
    }
    for (int i = 0; i < 6; i++)
        for (int c = 0; c < WHEEL_COLS; c++) { double v = 0; for (int k = 0; k < 6; k++) v += U[i * 6 + k] * R[k * WHEEL_COLS + c]; J[i * WHEEL_COLS + c] = v; }
    return true;
}

constexpr int PLANE_COLS = 16;   // local columns of one plane factor: pose_i 6 | wheel extrinsic 6 | plane rotation 3 | plane height 1
// PlaneFactor::Evaluate (reference factor/plane_factor.h:24-118); mirrors oracle/ba_oracle.c (gfo_eval_plane).
__device__ inline void eval_plane(const double* pose_i, const double* exw, const double* qpw, double zpw, const double* sinfo,
                                  double* res, double* J, bool jac)
{
    const double* Pi = pose_i; const double* Qi = pose_i + 3; const double* tio = exw; const double* qio = exw + 3;
    double Ri[9], rio[9], Rpw[9], RiT[9], rioT[9], RpwT[9];
    q_to_R(Qi, Ri); q_to_R(qio, rio); q_to_R(qpw, Rpw); m3_T(Ri, RiT); m3_T(rio, rioT); m3_T(Rpw, RpwT);
    const double e3[3] = {0, 0, 1};
    double a[3], b[3], c[3], t[3], u[3];
    m3_v(RpwT, e3, a); m3_v(RiT, a, b); m3_v(rioT, b, c);
    m3_v(Ri, tio, t); for (int k = 0; k < 3; k++) t[k] += Pi[k];
    m3_v(Rpw, t, u);
    res[0] = sinfo[0] * c[0]; res[1] = sinfo[1] * c[1]; res[2] = sinfo[2] * (zpw + u[2]);
    if (!jac) return;
    for (int k = 0; k < 3 * PLANE_COLS; k++) J[k] = 0.0;
    double S[9], A[9], B[9];
    skew(b, S); m3_mul(rioT, S, A);
    for (int r = 0; r < 2; r++) for (int k = 0; k < 3; k++) J[r * PLANE_COLS + 3 + k] = sinfo[r] * A[r * 3 + k];
    for (int k = 0; k < 3; k++) J[2 * PLANE_COLS + k] = sinfo[2] * Rpw[6 + k];
    skew(tio, S); m3_mul(Rpw, Ri, A); m3_mul(A, S, B);
    for (int k = 0; k < 3; k++) J[2 * PLANE_COLS + 3 + k] = -sinfo[2] * B[6 + k];
    skew(c, S);
    for (int r = 0; r < 2; r++) for (int k = 0; k < 3; k++) J[r * PLANE_COLS + 9 + k] = sinfo[r] * S[r * 3 + k];
    for (int k = 0; k < 3; k++) J[2 * PLANE_COLS + 6 + k] = sinfo[2] * A[6 + k];          // e3^T Rpw Ri
    skew(a, S); m3_mul(RiT, S, A); m3_mul(rioT, A, B);
    for (int r = 0; r < 2; r++) for (int k = 0; k < 3; k++) J[r * PLANE_COLS + 12 + k] = sinfo[r] * B[r * 3 + k];
    skew(t, S); m3_mul(Rpw, S, A);
    for (int k = 0; k < 3; k++) J[2 * PLANE_COLS + 12 + k] = -sinfo[2] * A[6 + k];
    J[2 * PLANE_COLS + 15] = sinfo[2];
}

}  // namespace gfba
