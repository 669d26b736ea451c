// ba_solver.cu -- gf_ba_* (C ABI): Estimator::optimization()'s ceres::Solve (DENSE_SCHUR + DOGLEG,
// estimator.cpp:3303-3318) on one B200, FP64.  Algorithm and constants: oracle/ba_oracle.c (restatement of
// Ceres 1.14's trust_region_minimizer.cc / dogleg_strategy.cc; parity with Ceres itself is unpinned).
//
// Data flow of one solve (everything stays on the device between the upload and the final download):
//   k_ba_setup      IMU sqrt-information matrices, H_prior = J0^T J0, solver state
//   k_ba_eval       mode 0 "linearise": one CTA per pose pair (i,j) evaluates its visual factors, stages the
//                   Huber-corrected Jacobian slab [Ji|Jj|Jex|Jtd] in shared memory and reduces it to block
//                   Hessians; one CTA per IMU factor; one CTA for the marginalisation prior.  Landmark terms
//                   (h_ll, g_l, W = H_landmark,camera) go to their own arrays: J is never materialised.
//                   mode 1 "candidate cost": residuals only at x (+) delta.
//   k_ba_step       single CTA: Jacobi scaling, Schur complement of the free landmarks into the packed reduced
//                   system, Cholesky (rhs carried as an extra row), traditional dogleg, model cost change,
//                   candidate x (+) delta.
//   k_ba_decide     step acceptance, trust-region / mu update, convergence tests (the Ceres state machine).
// All four are enqueued for every iteration up front; kernels return immediately once the state says "done",
// so the host synchronises exactly once per solve.
#include <new>
#include <vector>

#include "ba_factors.cuh"

using namespace gf;
using namespace gfba;

namespace gfba {

constexpr int MAXF = GF_BA_MAX_FRAMES;
constexpr int X_POSE = 0, X_SB = 7 * MAXF, X_EX = X_SB + 9 * MAXF, X_TD = X_EX + 7, X_FEAT = X_TD + 1;
constexpr int PAIR_THREADS = 256, PAIR_CHUNK = 64;   // factors staged per pass (2*64 rows x 20 cols in smem)
constexpr int MAX_NC = 208;                          // reduced (camera-side) dimension supported by k_ba_step

struct BaState {
    double x_cost, cand_cost, radius, mu, alpha, dogleg_norm, model_change, x_norm, step_norm, grad_max;
    double cost_hist[GF_BA_MAX_ITERATIONS + 1], radius_hist[GF_BA_MAX_ITERATIONS + 1];
    double acc_cost[2];          // cost accumulated by k_ba_eval into buffer 0/1
    int it, reuse, done, termination, n_success, invalid_streak, need_linearize, step_valid, cur, first, max_iter, solver_failed;
};

struct BaDev {
    int F, nfeat, n_vis, n_imu, n_pairs, nc, L, n;
    int col_pose[MAXF], col_sb[MAXF], col_ex, col_td;
    const int* col_feat;
    double *X, *Xc;
    const gf_ba_visual_factor* vis;
    const int *pair_start, *pair_ij;
    const gf_ba_imu_factor* imu;
    double* imu_sqrt;
    int pn, pnb;
    int pkind[64], pindex[64], pidx[64], pxoff[64];
    const double *pJ, *pr0, *px0;
    const int* pcol;          // [pn] prior column -> layout column or -1
    double* Hp;               // [nc*nc]
    double* acc[2];           // accumulators: [H nc*nc | g n | W L*nc | hll L]
    double *scale, *diag, *gs, *gn, *step, *delta;
    double gravity[3], vis_sqrt_info;
    BaState* st;
};

__device__ __forceinline__ double* acc_H(const BaDev& d, int b) { return d.acc[b]; }
__device__ __forceinline__ double* acc_g(const BaDev& d, int b) { return d.acc[b] + (size_t)d.nc * d.nc; }
__device__ __forceinline__ double* acc_W(const BaDev& d, int b) { return d.acc[b] + (size_t)d.nc * d.nc + d.n; }
__device__ __forceinline__ double* acc_hll(const BaDev& d, int b) { return d.acc[b] + (size_t)d.nc * d.nc + d.n + (size_t)d.L * d.nc; }
__host__ __device__ inline size_t acc_size(int nc, int L) { return (size_t)nc * nc + (nc + L) + (size_t)L * nc + L; }

__device__ __forceinline__ double block_reduce_sum(double v, double* sh)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0;
    if (w == 0) {
        t = (l < nw) ? sh[l] : 0.0;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (l == 0) sh[0] = t;
    }
    __syncthreads();
    t = sh[0];
    __syncthreads();
    return t;
}

// ------------------------------------------------------------------------------------------------
__global__ void k_ba_setup(BaDev d)
{
    int gid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    // H_prior = J0^T J0 mapped to the layout's columns
    const int nc = d.nc, pn = d.pn;
    for (int e = gid; e < nc * nc; e += gsz) d.Hp[e] = 0.0;
    if (gid < d.n_imu) {
        double M[15 * 30];
        if (!sqrt_info_from_cov(d.imu[gid].covariance, 15, d.imu_sqrt + 225 * gid, M)) d.st->termination = GF_BA_FAILURE;
    }
    if (gid == 0) {
        BaState& s = *d.st;
        s.radius = 1e4; s.mu = 1e-8; s.reuse = 0; s.done = 0; s.it = 0; s.n_success = 0; s.invalid_streak = 0;
        s.need_linearize = 1; s.step_valid = 0; s.cur = 1; s.first = 1; s.acc_cost[0] = s.acc_cost[1] = 0.0;   // first linearisation goes to buffer 0
        s.x_cost = 0; s.cand_cost = 0;
    }
}
__global__ void k_ba_prior_hessian(BaDev d)
{
    const int nc = d.nc, pn = d.pn;
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= pn * pn) return;
    int pa = e / pn, pb = e - pa * pn;
    int ca = d.pcol[pa], cb = d.pcol[pb];
    if (ca < 0 || cb < 0) return;
    double s = 0;
    for (int k = 0; k < pn; k++) s += d.pJ[(size_t)k * pn + pa] * d.pJ[(size_t)k * pn + pb];
    d.Hp[(size_t)ca * nc + cb] = s;
}

// ------------------------------------------------------------------------------------------------
// mode 0: linearise at X into the inactive accumulator (if state.need_linearize); mode 1: cost at Xc
__global__ void __launch_bounds__(PAIR_THREADS) k_ba_eval(BaDev d, int mode)
{
    __shared__ double sJ[2 * PAIR_CHUNK][20];
    __shared__ double sR[2 * PAIR_CHUNK];
    __shared__ double sred[32];
    __shared__ double simu_J[15 * 30], simu_JU[15 * 30], simu_r[15], simu_ru[15];
    const BaState& st = *d.st;
    if (st.done) return;
    const int tid = threadIdx.x;
    const int tgt = st.cur ^ 1;            // inactive buffer
    if (mode == 0) { if (!st.need_linearize) return; }
    else {
        if (!st.step_valid) return;
        // zero the inactive accumulator for the linearisation that may follow (safe: nothing reads it now)
        size_t tot = acc_size(d.nc, d.L);
        for (size_t e = (size_t)blockIdx.x * blockDim.x + tid; e < tot; e += (size_t)gridDim.x * blockDim.x) d.acc[tgt][e] = 0.0;
        if (blockIdx.x == 0 && tid == 0) d.st->acc_cost[tgt] = 0.0;
    }
    const double* X = mode == 0 ? d.X : d.Xc;
    double* costp = mode == 0 ? &d.st->acc_cost[tgt] : &d.st->cand_cost;
    const bool jac = (mode == 0);
    const int b = blockIdx.x;
    if (b < d.n_pairs) {
        // ---------------- visual factors of one pose pair ----------------
        const int f0 = d.pair_start[b], f1 = d.pair_start[b + 1];
        const int pi = d.pair_ij[2 * b], pj = d.pair_ij[2 * b + 1];
        int cols[4] = {d.col_pose[pi], d.col_pose[pj], d.col_ex, d.col_td};
        const int bsz[4] = {6, 6, 6, 1}, boff[4] = {0, 6, 12, 18};
        double out0 = 0, out1 = 0;          // this thread's two entries of the 19x19 (+19 gradient) product
        double cost = 0;
        for (int c0 = f0; c0 < f1; c0 += PAIR_CHUNK) {
            int nf = min(PAIR_CHUNK, f1 - c0);
            if (tid < nf) {
                const gf_ba_visual_factor& f = d.vis[c0 + tid];
                double res[2], J[40];
                eval_visual(f, d.vis_sqrt_info, X + X_POSE + 7 * f.imu_i, X + X_POSE + 7 * f.imu_j, X + X_EX, X[X_FEAT + f.feature], X[X_TD], res, J, jac);
                double sc, rho = huber(res[0] * res[0] + res[1] * res[1], sc);
                cost += 0.5 * rho;
                if (jac) {
                    for (int r = 0; r < 2; r++) { for (int c = 0; c < 20; c++) sJ[2 * tid + r][c] = sc * J[r * 20 + c]; sR[2 * tid + r] = sc * res[r]; }
                    int cf = d.col_feat[f.feature];
                    if (cf >= 0) {      // landmark terms: h_ll, g_l, W[l][camera cols]
                        int l = cf - d.nc;
                        double j0 = sc * J[19], j1 = sc * J[39], r0 = sc * res[0], r1 = sc * res[1];
                        atomicAdd(&acc_hll(d, tgt)[l], j0 * j0 + j1 * j1);
                        atomicAdd(&acc_g(d, tgt)[cf], j0 * r0 + j1 * r1);
                        double* Wl = acc_W(d, tgt) + (size_t)l * d.nc;
                        for (int q = 0; q < 4; q++) if (cols[q] >= 0)
                            for (int k = 0; k < bsz[q]; k++) atomicAdd(&Wl[cols[q] + k], j0 * sc * J[boff[q] + k] + j1 * sc * J[20 + boff[q] + k]);
                    }
                }
            }
            if (jac) {
                __syncthreads();
                // 19x19 J^T J entries + 19 J^T r entries = 380 outputs, two per thread
                for (int o = tid, slot = 0; o < 380; o += PAIR_THREADS, slot++) {
                    double s = 0;
                    if (o < 361) { int a = o / 19, c = o - a * 19; for (int r = 0; r < 2 * nf; r++) s += sJ[r][a] * sJ[r][c]; }
                    else { int a = o - 361; for (int r = 0; r < 2 * nf; r++) s += sJ[r][a] * sR[r]; }
                    if (slot == 0) out0 += s; else out1 += s;
                }
                __syncthreads();
            }
        }
        if (jac) {
            auto col_of = [&](int a) { int q = a < 6 ? 0 : a < 12 ? 1 : a < 18 ? 2 : 3; return cols[q] < 0 ? -1 : cols[q] + (a - boff[q]); };
            for (int o = tid, slot = 0; o < 380; o += PAIR_THREADS, slot++) {
                double v = slot == 0 ? out0 : out1;
                if (o < 361) { int a = o / 19, c = o - a * 19; int ca = col_of(a), cc = col_of(c); if (ca >= 0 && cc >= 0) atomicAdd(&acc_H(d, tgt)[(size_t)ca * d.nc + cc], v); }
                else { int ca = col_of(o - 361); if (ca >= 0) atomicAdd(&acc_g(d, tgt)[ca], v); }
            }
        }
        cost = block_reduce_sum(cost, sred);
        if (tid == 0 && cost != 0.0) atomicAdd(costp, cost);
    } else if (b < d.n_pairs + d.n_imu) {
        // ---------------- one IMU factor ----------------
        const int m = b - d.n_pairs;
        const gf_ba_imu_factor& f = d.imu[m];
        const double* U = d.imu_sqrt + 225 * m;
        if (tid == 0) eval_imu_raw(f, d.gravity, X + X_POSE + 7 * f.i, X + X_SB + 9 * f.i, X + X_POSE + 7 * f.j, X + X_SB + 9 * f.j, simu_r, simu_J, jac);
        __syncthreads();
        if (tid < 15) { double s = 0; for (int k = 0; k < 15; k++) s += U[tid * 15 + k] * simu_r[k]; simu_ru[tid] = s; }
        if (jac) for (int o = tid; o < 450; o += PAIR_THREADS) { int r = o / 30, c = o - r * 30; double s = 0; for (int k = 0; k < 15; k++) s += U[r * 15 + k] * simu_J[k * 30 + c]; simu_JU[o] = s; }
        __syncthreads();
        if (tid == 0) { double c = 0; for (int k = 0; k < 15; k++) c += 0.5 * simu_ru[k] * simu_ru[k]; atomicAdd(costp, c); }
        if (jac) {
            int cols[4] = {d.col_pose[f.i], d.col_sb[f.i], d.col_pose[f.j], d.col_sb[f.j]};
            const int boff[4] = {0, 6, 15, 21};
            auto col_of = [&](int a) { int q = a < 6 ? 0 : a < 15 ? 1 : a < 21 ? 2 : 3; return cols[q] < 0 ? -1 : cols[q] + (a - boff[q]); };
            for (int o = tid; o < 930; o += PAIR_THREADS) {
                if (o < 900) {
                    int a = o / 30, c = o - a * 30; int ca = col_of(a), cc = col_of(c);
                    if (ca < 0 || cc < 0) continue;
                    double s = 0; for (int r = 0; r < 15; r++) s += simu_JU[r * 30 + a] * simu_JU[r * 30 + c];
                    atomicAdd(&acc_H(d, tgt)[(size_t)ca * d.nc + cc], s);
                } else {
                    int a = o - 900, ca = col_of(a);
                    if (ca < 0) continue;
                    double s = 0; for (int r = 0; r < 15; r++) s += simu_JU[r * 30 + a] * simu_ru[r];
                    atomicAdd(&acc_g(d, tgt)[ca], s);
                }
            }
        }
    } else if (b == d.n_pairs + d.n_imu && d.pn > 0) {
        // ---------------- marginalisation prior: r = r0 + J0 dx, g += J0^T r (H_prior is constant) ----------------
        extern __shared__ double sdyn[];    // dx[pn], r[pn]
        double* dx = sdyn; double* rr = sdyn + d.pn;
        const int pn = d.pn;
        for (int blk = tid; blk < d.pnb; blk += PAIR_THREADS) {
            int kind = d.pkind[blk], idx = d.pidx[blk];
            const double* x0 = d.px0 + d.pxoff[blk];
            const double* x = kind == GF_BA_BLOCK_POSE ? X + X_POSE + 7 * d.pindex[blk] : kind == GF_BA_BLOCK_SPEEDBIAS ? X + X_SB + 9 * d.pindex[blk]
                              : kind == GF_BA_BLOCK_EX_POSE ? X + X_EX : X + X_TD;
            int size = (kind == GF_BA_BLOCK_POSE || kind == GF_BA_BLOCK_EX_POSE) ? 7 : kind == GF_BA_BLOCK_SPEEDBIAS ? 9 : 1;
            if (size != 7) for (int k = 0; k < size; k++) dx[idx + k] = x[k] - x0[k];
            else {
                for (int k = 0; k < 3; k++) dx[idx + k] = x[k] - x0[k];
                double qi[4], dq[4]; q_inv(x0 + 3, qi); q_mul(qi, x + 3, dq);
                double sgn = (dq[3] >= 0) ? 1.0 : -1.0;
                for (int k = 0; k < 3; k++) dx[idx + 3 + k] = 2.0 * sgn * dq[k];
            }
        }
        __syncthreads();
        double cost = 0;
        for (int i = tid; i < pn; i += PAIR_THREADS) {
            double v = d.pr0[i];
            for (int k = 0; k < pn; k++) v += d.pJ[(size_t)i * pn + k] * dx[k];
            rr[i] = v; cost += 0.5 * v * v;
        }
        cost = block_reduce_sum(cost, sred);
        if (tid == 0) atomicAdd(costp, cost);
        if (jac)
            for (int c = tid; c < pn; c += PAIR_THREADS) {
                int lc = d.pcol[c];
                if (lc < 0) continue;
                double s = 0;
                for (int k = 0; k < pn; k++) s += d.pJ[(size_t)k * pn + c] * rr[k];
                atomicAdd(&acc_g(d, tgt)[lc], s);
            }
    }
}

// ------------------------------------------------------------------------------------------------
// ambient-space helpers over the non-constant blocks
__device__ __forceinline__ void for_each_free_block(const BaDev& d, int t, int nt, double (*fn)(const BaDev&, int off, int size, void* ctx), void* ctx, double& acc);

__device__ inline void plus_all(const BaDev& d, const double* X, const double* delta, double* Y, int tid, int nt)
{
    // copy everything, then overwrite the free blocks
    const int tot = X_FEAT + d.nfeat;
    for (int e = tid; e < tot; e += nt) Y[e] = X[e];
    __syncthreads();
    for (int f = tid; f < d.F; f += nt) {
        if (d.col_pose[f] >= 0) pose_plus(X + X_POSE + 7 * f, delta + d.col_pose[f], Y + X_POSE + 7 * f);
        if (d.col_sb[f] >= 0) for (int k = 0; k < 9; k++) Y[X_SB + 9 * f + k] = X[X_SB + 9 * f + k] + delta[d.col_sb[f] + k];
    }
    if (tid == 0) {
        if (d.col_ex >= 0) pose_plus(X + X_EX, delta + d.col_ex, Y + X_EX);
        if (d.col_td >= 0) Y[X_TD] = X[X_TD] + delta[d.col_td];
    }
    for (int k = tid; k < d.nfeat; k += nt) { int c = d.col_feat[k]; if (c >= 0) Y[X_FEAT + k] = X[X_FEAT + k] + delta[c]; }
    __syncthreads();
}
// sum of squares / max abs of (A - B) over the ambient coordinates of the free blocks (B may be null)
__device__ inline void diff_norms(const BaDev& d, const double* A, const double* B, double& s2, double& mx, int tid, int nt)
{
    s2 = 0; mx = 0;
    auto acc = [&](int off, int size) { for (int k = 0; k < size; k++) { double v = A[off + k] - (B ? B[off + k] : 0.0); s2 += v * v; mx = fmax(mx, fabs(v)); } };
    for (int f = tid; f < d.F; f += nt) { if (d.col_pose[f] >= 0) acc(X_POSE + 7 * f, 7); if (d.col_sb[f] >= 0) acc(X_SB + 9 * f, 9); }
    if (tid == 0) { if (d.col_ex >= 0) acc(X_EX, 7); if (d.col_td >= 0) acc(X_TD, 1); }
    for (int k = tid; k < d.nfeat; k += nt) if (d.col_feat[k] >= 0) acc(X_FEAT + k, 1);
}
__device__ __forceinline__ double block_reduce_max(double v, double* sh)
{
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0;
    if (w == 0) {
        t = (l < nw) ? sh[l] : 0.0;
        for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (l == 0) sh[0] = t;
    }
    __syncthreads();
    t = sh[0];
    __syncthreads();
    return t;
}

// DoglegStrategy::ComputeStep + TrustRegionMinimizer::ComputeTrustRegionStep + candidate point.
// Dynamic shared memory: packed lower triangle of the (nc+1) x (nc+1) augmented reduced system.
__global__ void __launch_bounds__(1024) k_ba_step(BaDev d)
{
    extern __shared__ double S[];                 // packed: S[i*(i+1)/2 + j], j <= i ; row nc = rhs
    __shared__ double sred[32];
    __shared__ double colbuf[MAX_NC + 1];
    __shared__ double yc[MAX_NC + 1];
    __shared__ int s_fail;
    BaState& st = *d.st;
    if (st.done) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nc = d.nc, L = d.L, n = d.n;
    if (st.need_linearize) {                      // a fresh linearisation landed in the inactive buffer: adopt it
        __syncthreads();
        if (tid == 0) { st.cur ^= 1; st.need_linearize = 0; st.x_cost = st.acc_cost[st.cur]; }
        __syncthreads();
        const int cur = st.cur;
        double* H = acc_H(d, cur); const double* Hp = d.Hp;
        // total Hessian diagonal includes the prior
        if (st.first) {                           // Jacobi scaling, fixed at iteration 0
            for (int c = tid; c < nc; c += nt) d.scale[c] = 1.0 / (1.0 + sqrt(H[(size_t)c * nc + c] + Hp[(size_t)c * nc + c]));
            for (int l = tid; l < L; l += nt) d.scale[nc + l] = 1.0 / (1.0 + sqrt(acc_hll(d, cur)[l]));
        }
        __syncthreads();
        // gradient_max_norm = |x - Plus(x, -g)|_inf with the unscaled gradient
        for (int c = tid; c < n; c += nt) d.delta[c] = -acc_g(d, cur)[c];
        __syncthreads();
        plus_all(d, d.X, d.delta, d.Xc, tid, nt);
        double s2, mx; diff_norms(d, d.X, d.Xc, s2, mx, tid, nt);
        mx = block_reduce_max(mx, sred);
        double xs2, xmx; diff_norms(d, d.X, nullptr, xs2, xmx, tid, nt);
        xs2 = block_reduce_sum(xs2, sred);
        if (tid == 0) {
            st.grad_max = mx; st.x_norm = sqrt(xs2);
            if (st.first) { st.cost_hist[0] = st.x_cost; st.radius_hist[0] = st.radius; st.first = 0; }
            else { st.cost_hist[st.it] = st.x_cost; }      // cost after the accepted step of iteration `it`
            if (mx <= 1e-10 || n == 0) { st.done = 1; st.termination = GF_BA_CONVERGENCE_GRADIENT; }
            st.reuse = 0;
        }
        __syncthreads();
        if (st.done) return;
    }
    // ---- TrustRegionMinimizer: iteration bookkeeping ----
    if (st.it >= st.max_iter || st.radius < 1e-32) { if (tid == 0) { st.done = 1; st.termination = GF_BA_NO_CONVERGENCE; } return; }
    __syncthreads();
    if (tid == 0) { st.it++; st.step_valid = 0; st.solver_failed = 0; }
    __syncthreads();
    const int cur = st.cur;
    const double* H = acc_H(d, cur); const double* Hp = d.Hp; const double* g = acc_g(d, cur);
    const double* W = acc_W(d, cur); const double* hll = acc_hll(d, cur);
    const double* sc = d.scale;
    if (!st.reuse) {
        // diag = sqrt(clamp(diag(H'), 1e-6, 1e32)); gs = g'/D ; Cauchy alpha = |gs|^2 / (v^T H' v), v = gs / D
        for (int c = tid; c < n; c += nt) {
            double hd = (c < nc) ? (H[(size_t)c * nc + c] + Hp[(size_t)c * nc + c]) : hll[c - nc];
            hd *= sc[c] * sc[c];
            hd = hd < 1e-6 ? 1e-6 : (hd > 1e32 ? 1e32 : hd);
            double D = sqrt(hd);
            d.diag[c] = D;
            d.gs[c] = g[c] * sc[c] / D;
        }
        __syncthreads();
        double num = 0, den = 0;
        for (int c = tid; c < n; c += nt) num += d.gs[c] * d.gs[c];
        // v^T H' v = vc^T H'cc vc + 2 vl^T W' vc + sum h'll vl^2,  v = gs / D
        for (int a = tid; a < nc; a += nt) {
            double va = d.gs[a] / d.diag[a] * sc[a], s = 0;
            for (int b = 0; b < nc; b++) s += (H[(size_t)a * nc + b] + Hp[(size_t)a * nc + b]) * (d.gs[b] / d.diag[b] * sc[b]);
            den += va * s;
        }
        for (int l = tid; l < L; l += nt) {
            double vl = d.gs[nc + l] / d.diag[nc + l] * sc[nc + l], s = 0;
            for (int b = 0; b < nc; b++) s += W[(size_t)l * nc + b] * (d.gs[b] / d.diag[b] * sc[b]);
            den += 2.0 * vl * s + hll[l] * vl * vl;
        }
        num = block_reduce_sum(num, sred);
        den = block_reduce_sum(den, sred);
        if (tid == 0) st.alpha = num / den;
        // ---- ComputeGaussNewtonStep: (H' + mu D^2) y = g' by Schur complement on the landmarks + Cholesky ----
        while (true) {
            const double mu = st.mu;
            __syncthreads();
            if (tid == 0) s_fail = 0;
            __syncthreads();
            // e_l = 1 / (h'll + mu D_l^2)   (kept in gn[nc + l] for now)
            for (int l = tid; l < L; l += nt) {
                double v = hll[l] * sc[nc + l] * sc[nc + l] + mu * d.diag[nc + l] * d.diag[nc + l];
                if (!(v > 0)) s_fail = 1;
                d.gn[nc + l] = 1.0 / v;
            }
            __syncthreads();
            const int tri = (nc + 1) * (nc + 2) / 2;
            for (int e = tid; e < tri; e += nt) {
                // unpack (a, b), a >= b
                int a = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
                while ((a + 1) * (a + 2) / 2 <= e) a++;
                while (a * (a + 1) / 2 > e) a--;
                int b = e - a * (a + 1) / 2;
                double v;
                if (a < nc) {
                    v = (H[(size_t)a * nc + b] + Hp[(size_t)a * nc + b]) * sc[a] * sc[b];
                    if (a == b) v += mu * d.diag[a] * d.diag[a];
                    double s = 0;
                    for (int l = 0; l < L; l++) s += d.gn[nc + l] * sc[nc + l] * sc[nc + l] * W[(size_t)l * nc + a] * W[(size_t)l * nc + b];
                    v -= s * sc[a] * sc[b];
                } else if (b < nc) {                  // rhs row
                    v = g[b] * sc[b];
                    double s = 0;
                    for (int l = 0; l < L; l++) s += d.gn[nc + l] * sc[nc + l] * sc[nc + l] * W[(size_t)l * nc + b] * g[nc + l];
                    v -= s * sc[b];
                } else v = 0.0;
                S[e] = v;
            }
            __syncthreads();
            // right-looking Cholesky on the packed lower triangle; row nc (the rhs) is carried along: after the
            // loop S[nc][0..nc) = L^-1 rhs
            for (int j = 0; j < nc && !s_fail; j++) {
                double dj = S[j * (j + 1) / 2 + j];
                if (!(dj > 0.0)) { s_fail = 1; break; }      // same value seen by all threads
                double inv = 1.0 / sqrt(dj);
                for (int i = j + tid; i <= nc; i += nt) colbuf[i] = S[i * (i + 1) / 2 + j] * inv;
                __syncthreads();
                for (int i = j + tid; i <= nc; i += nt) S[i * (i + 1) / 2 + j] = colbuf[i];
                const int m = nc - j;                        // trailing rows j+1..nc
                for (int e = tid; e < m * (m + 1) / 2; e += nt) {
                    int a = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
                    while ((a + 1) * (a + 2) / 2 <= e) a++;
                    while (a * (a + 1) / 2 > e) a--;
                    int b = e - a * (a + 1) / 2;
                    int i = j + 1 + a, k = j + 1 + b;
                    if (i == nc && k == nc) continue;
                    S[i * (i + 1) / 2 + k] -= colbuf[i] * colbuf[k];
                }
                __syncthreads();
            }
            __syncthreads();
            if (!s_fail) {
                // back substitution L^T y = z (z = row nc), one warp, column sweeps over contiguous packed rows
                if (tid < 32) {
                    for (int c = tid; c < nc; c += 32) yc[c] = S[nc * (nc + 1) / 2 + c];
                    __syncwarp();
                    for (int j = nc - 1; j >= 0; j--) {
                        double xj = yc[j] / S[j * (j + 1) / 2 + j];
                        __syncwarp();
                        if (tid == 0) yc[j] = xj;
                        for (int k = tid; k < j; k += 32) yc[k] -= S[j * (j + 1) / 2 + k] * xj;
                        __syncwarp();
                    }
                }
                __syncthreads();
                for (int c = tid; c < nc; c += nt) if (!isfinite(yc[c])) s_fail = 1;
                __syncthreads();
            }
            if (!s_fail) {
                // y_l = e_l (g'_l - w'_l . y_c) ; gn = -D y
                for (int l = tid; l < L; l += nt) {
                    double s = g[nc + l] * sc[nc + l];
                    double t = 0;
                    for (int b = 0; b < nc; b++) t += W[(size_t)l * nc + b] * sc[b] * yc[b];
                    double yl = d.gn[nc + l] * (s - t * sc[nc + l]);
                    d.step[nc + l] = yl;                       // temp
                }
                __syncthreads();
                for (int c = tid; c < n; c += nt) { double y = (c < nc) ? yc[c] : d.step[c]; d.gn[c] = -d.diag[c] * y; }
                __syncthreads();
                break;
            }
            __syncthreads();
            if (tid == 0) st.mu *= 10.0;
            __syncthreads();
            if (!(st.mu < 1.0)) { if (tid == 0) st.solver_failed = 1; break; }
        }
        __syncthreads();
        if (tid == 0) st.reuse = 1;
        __syncthreads();
    }
    if (!st.solver_failed) {
        // ---- ComputeTraditionalDoglegStep ----
        double g2 = 0, n2 = 0, ga = 0;
        for (int c = tid; c < n; c += nt) { g2 += d.gs[c] * d.gs[c]; n2 += d.gn[c] * d.gn[c]; ga += d.gs[c] * d.gn[c]; }
        g2 = block_reduce_sum(g2, sred); n2 = block_reduce_sum(n2, sred); ga = block_reduce_sum(ga, sred);
        const double gnorm = sqrt(g2), gnn = sqrt(n2), radius = st.radius, alpha = st.alpha;
        double ca, cb, dn;                      // step = ca * gs + cb * gn  (D-space)
        if (gnn <= radius) { ca = 0; cb = 1; dn = gnn; }
        else if (gnorm * alpha >= radius) { ca = -(radius / gnorm); cb = 0; dn = radius; }
        else {
            double b_dot_a = -alpha * ga;
            double a2 = pow(alpha * gnorm, 2.0);
            double bma2 = a2 - 2 * b_dot_a + pow(gnn, 2);
            double cc = b_dot_a - a2;
            double dd = sqrt(cc * cc + bma2 * (pow(radius, 2.0) - a2));
            double beta = (cc <= 0) ? (dd - cc) / bma2 : (radius * radius - a2) / (dd + cc);
            ca = -alpha * (1.0 - beta); cb = beta; dn = -1.0;
        }
        double nn = 0;
        for (int c = tid; c < n; c += nt) { double s = ca * d.gs[c] + cb * d.gn[c]; nn += s * s; d.step[c] = s / d.diag[c]; }
        nn = block_reduce_sum(nn, sred);
        if (dn < 0) dn = sqrt(nn);
        // model_cost_change = -(s^T g' + s^T H' s / 2)
        double sg = 0, sHs = 0;
        for (int c = tid; c < n; c += nt) sg += d.step[c] * g[c] * sc[c];
        for (int a = tid; a < nc; a += nt) {
            double s = 0;
            for (int b = 0; b < nc; b++) s += (H[(size_t)a * nc + b] + Hp[(size_t)a * nc + b]) * sc[b] * d.step[b];
            sHs += d.step[a] * sc[a] * s;
        }
        for (int l = tid; l < L; l += nt) {
            double sl = d.step[nc + l] * sc[nc + l], s = 0;
            for (int b = 0; b < nc; b++) s += W[(size_t)l * nc + b] * sc[b] * d.step[b];
            sHs += 2.0 * sl * s + hll[l] * sl * sl;
        }
        sg = block_reduce_sum(sg, sred); sHs = block_reduce_sum(sHs, sred);
        const double model_change = -(sg + 0.5 * sHs);
        for (int c = tid; c < n; c += nt) d.delta[c] = d.step[c] * sc[c];
        __syncthreads();
        plus_all(d, d.X, d.delta, d.Xc, tid, nt);
        double s2, mx; diff_norms(d, d.X, d.Xc, s2, mx, tid, nt);
        s2 = block_reduce_sum(s2, sred);
        if (tid == 0) {
            st.dogleg_norm = dn; st.model_change = model_change; st.step_norm = sqrt(s2);
            st.step_valid = model_change > 0.0 ? 1 : 0;
            st.cand_cost = 0.0;
        }
    }
}

// TrustRegionMinimizer: HandleInvalidStep / tolerances / IsStepSuccessful / HandleSuccessfulStep / HandleUnsuccessfulStep
__global__ void k_ba_decide(BaDev d)
{
    BaState& st = *d.st;
    if (st.done) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ int accept;
    if (tid == 0) {
        accept = 0;
        const int it = st.it;
        if (!st.step_valid) {
            if (++st.invalid_streak >= 5) { st.done = 1; st.termination = GF_BA_FAILURE; }
            st.mu *= 10.0; st.reuse = 0;
            st.cost_hist[it] = st.x_cost; st.radius_hist[it] = st.radius;
        } else {
            st.invalid_streak = 0;
            const double x_cost = st.x_cost, cand = st.cand_cost;
            if (st.step_norm <= 1e-8 * (st.x_norm + 1e-8)) { st.done = 1; st.termination = GF_BA_CONVERGENCE_PARAMETER; st.cost_hist[it] = x_cost; st.radius_hist[it] = st.radius; }
            else if (fabs(x_cost - cand) <= 1e-6 * x_cost) { st.done = 1; st.termination = GF_BA_CONVERGENCE_FUNCTION; st.cost_hist[it] = x_cost; st.radius_hist[it] = st.radius; }
            else {
                double rel = (x_cost - cand) / st.model_change;
                if (rel > 1e-3) {
                    accept = 1; st.n_success++;
                    if (rel < 0.25) st.radius *= 0.5;
                    if (rel > 0.75) st.radius = fmax(st.radius, 3.0 * st.dogleg_norm);
                    st.mu = fmax(1e-8, 2.0 * st.mu / 10.0);
                    st.reuse = 0; st.need_linearize = 1;
                    st.radius_hist[it] = st.radius;          // cost_hist[it] is written when the new linearisation is adopted
                } else {
                    st.radius *= 0.5; st.reuse = 1;
                    st.cost_hist[it] = x_cost; st.radius_hist[it] = st.radius;
                }
            }
        }
    }
    __syncthreads();
    if (accept) { const int tot = X_FEAT + d.nfeat; for (int e = tid; e < tot; e += nt) d.X[e] = d.Xc[e]; }
}

}  // namespace gfba

// ------------------------------------------------------------------------------------------------
struct gf_ba {
    int device;
    cudaStream_t s;
    cudaEvent_t e0, e1;
    // growable device buffers
    void* dbuf; size_t dcap;
    void* hbuf; size_t hcap;     // pinned staging
};

static int ensure(gf_ba* s, size_t dbytes)
{
    if (dbytes > s->dcap) {
        if (s->dbuf) cudaFree(s->dbuf);
        if (s->hbuf) cudaFreeHost(s->hbuf);
        s->dcap = dbytes * 2;
        GF_CUDA(cudaMalloc(&s->dbuf, s->dcap));
        GF_CUDA(cudaHostAlloc(&s->hbuf, s->dcap, cudaHostAllocDefault));
        s->hcap = s->dcap;
    }
    return GF_OK;
}

extern "C" {

int gf_ba_create(gf_ba** out, int device)
{
    if (!out) return set_err(GF_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { snprintf(g_err, sizeof(g_err), "no CUDA device visible (%s); libgf_b200 has no CPU fallback", cudaGetErrorString(e)); return GF_ERR_NO_DEVICE; }
    if (device < 0 || device >= n) return set_err(GF_ERR_INVALID_ARG, "device index out of range");
    GF_CUDA(cudaSetDevice(device));
    gf_ba* s = new (std::nothrow) gf_ba();
    if (!s) return set_err(GF_ERR_CUDA, "out of host memory");
    memset(s, 0, sizeof(*s));
    s->device = device;
    GF_CUDA(cudaStreamCreateWithFlags(&s->s, cudaStreamNonBlocking));
    GF_CUDA(cudaEventCreate(&s->e0)); GF_CUDA(cudaEventCreate(&s->e1));
    GF_CUDA(cudaFuncSetAttribute(k_ba_step, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    *out = s;
    return GF_OK;
}

void gf_ba_destroy(gf_ba* s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->s);
    if (s->dbuf) cudaFree(s->dbuf);
    if (s->hbuf) cudaFreeHost(s->hbuf);
    cudaEventDestroy(s->e0); cudaEventDestroy(s->e1);
    cudaStreamDestroy(s->s);
    delete s;
}

int gf_ba_solve(gf_ba* s, const gf_ba_problem* p, gf_ba_summary* sum)
{
    if (!s || !p || !sum) return set_err(GF_ERR_INVALID_ARG, "null argument");
    if (p->n_frames < 1 || p->n_frames > GF_BA_MAX_FRAMES) return set_err(GF_ERR_INVALID_ARG, "n_frames out of range");
    if (p->n_wheel > 0) return set_err(GF_ERR_UNSUPPORTED, "wheel factors are not implemented yet (SURVEY 8a BA-6)");
    if (p->max_num_iterations < 0 || p->max_num_iterations > GF_BA_MAX_ITERATIONS) return set_err(GF_ERR_INVALID_ARG, "max_num_iterations out of range");
    GF_CUDA(cudaSetDevice(s->device));
    memset(sum, 0, sizeof(*sum));
    const int F = p->n_frames, nfeat = p->n_features;
    // ---- layout (same rules as ceres::Problem construction, estimator.cpp:2950-3100, 3233-3246, 3291) ----
    BaDev d; memset(&d, 0, sizeof(d));
    d.F = F; d.nfeat = nfeat; d.n_vis = p->n_visual; d.n_imu = p->n_imu;
    const bool use_sb = p->para_speed_bias && !p->pose0_const;
    int c = 0;
    for (int f = 0; f < MAXF; f++) { d.col_pose[f] = -1; d.col_sb[f] = -1; }
    for (int f = 0; f < F; f++) { bool k = p->frames_const || (f == 0 && p->pose0_const); if (!k) { d.col_pose[f] = c; c += 6; } }
    for (int f = 0; f < F; f++) { bool k = p->frames_const || !use_sb; if (!k) { d.col_sb[f] = c; c += 9; } }
    d.col_ex = p->ex_pose_const ? -1 : c; if (!p->ex_pose_const) c += 6;
    d.col_td = p->td_const ? -1 : c; if (!p->td_const) c += 1;
    d.nc = c;
    std::vector<int> col_feat(nfeat > 0 ? nfeat : 1, -1);
    for (int v = 0; v < p->n_visual; v++) {
        int k = p->visual[v].feature;
        if (k < 0 || k >= nfeat || p->visual[v].imu_i < 0 || p->visual[v].imu_i >= F || p->visual[v].imu_j < 0 || p->visual[v].imu_j >= F)
            return set_err(GF_ERR_INVALID_ARG, "visual factor index out of range");
        if (!p->feature_const[k] && col_feat[k] == -1) col_feat[k] = -2;
    }
    for (int k = 0; k < nfeat; k++) if (col_feat[k] == -2) col_feat[k] = c++;
    d.L = c - d.nc; d.n = c;
    if (d.nc > MAX_NC) return set_err(GF_ERR_CAPACITY, "reduced system larger than 208");
    // ---- sort visual factors by pose pair ----
    std::vector<int> pair_id(F * F, -1), pair_cnt;
    std::vector<int> pair_ij;
    for (int v = 0; v < p->n_visual; v++) {
        int key = p->visual[v].imu_i * F + p->visual[v].imu_j;
        if (pair_id[key] < 0) { pair_id[key] = (int)pair_cnt.size(); pair_cnt.push_back(0); pair_ij.push_back(p->visual[v].imu_i); pair_ij.push_back(p->visual[v].imu_j); }
        pair_cnt[pair_id[key]]++;
    }
    const int n_pairs = (int)pair_cnt.size();
    std::vector<int> pair_start(n_pairs + 1, 0);
    for (int k = 0; k < n_pairs; k++) pair_start[k + 1] = pair_start[k] + pair_cnt[k];
    d.n_pairs = n_pairs;
    const gf_ba_prior* pr = (p->prior && p->prior->n > 0) ? p->prior : nullptr;
    const int pn = pr ? pr->n : 0;
    std::vector<int> pcol(pn > 0 ? pn : 1, -1);
    size_t px0_len = 0;
    if (pr) {
        if (pr->n_blocks > 64) return set_err(GF_ERR_CAPACITY, "more than 64 prior blocks");
        d.pn = pn; d.pnb = pr->n_blocks;
        for (int b = 0; b < pr->n_blocks; b++) {
            int kind = pr->block_kind[b], idx = pr->block_index[b];
            d.pkind[b] = kind; d.pindex[b] = idx; d.pidx[b] = pr->block_idx[b]; d.pxoff[b] = (int)px0_len;
            int gs = (kind == GF_BA_BLOCK_POSE || kind == GF_BA_BLOCK_EX_POSE) ? 7 : kind == GF_BA_BLOCK_SPEEDBIAS ? 9 : 1;
            int ls = gs == 7 ? 6 : gs;
            int lc = kind == GF_BA_BLOCK_POSE ? d.col_pose[idx] : kind == GF_BA_BLOCK_SPEEDBIAS ? d.col_sb[idx] : kind == GF_BA_BLOCK_EX_POSE ? d.col_ex : kind == GF_BA_BLOCK_TD ? d.col_td : -1;
            if (kind > GF_BA_BLOCK_TD) return set_err(GF_ERR_UNSUPPORTED, "prior on wheel blocks not implemented yet");
            if (lc >= 0) for (int k = 0; k < ls; k++) pcol[pr->block_idx[b] + k] = lc + k;
            px0_len += gs;
        }
    }
    // ---- pack one upload buffer ----
    const int nc = d.nc, L = d.L, n = d.n;
    auto al = [](size_t v) { return (v + 15) / 16 * 16; };
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += al(bytes); return o; };
    const size_t o_X = take(sizeof(double) * (X_FEAT + nfeat)), o_vis = take(sizeof(gf_ba_visual_factor) * (size_t)p->n_visual),
                 o_imu = take(sizeof(gf_ba_imu_factor) * (size_t)p->n_imu), o_ps = take(sizeof(int) * (n_pairs + 1)), o_pij = take(sizeof(int) * 2 * (size_t)(n_pairs > 0 ? n_pairs : 1)),
                 o_cf = take(sizeof(int) * (size_t)(nfeat > 0 ? nfeat : 1)), o_pJ = take(sizeof(double) * (size_t)pn * pn), o_pr0 = take(sizeof(double) * pn),
                 o_px0 = take(sizeof(double) * px0_len), o_pcol = take(sizeof(int) * (size_t)(pn > 0 ? pn : 1));
    const size_t upload_bytes = off;
    const size_t o_Xc = take(sizeof(double) * (X_FEAT + nfeat)), o_sq = take(sizeof(double) * 225 * (size_t)(p->n_imu > 0 ? p->n_imu : 1)),
                 o_Hp = take(sizeof(double) * (size_t)nc * nc), o_a0 = take(sizeof(double) * acc_size(nc, L)), o_a1 = take(sizeof(double) * acc_size(nc, L)),
                 o_vec = take(sizeof(double) * 6 * (size_t)(n > 0 ? n : 1)), o_st = take(sizeof(BaState));
    int rc = ensure(s, off);
    if (rc) return rc;
    char* hb = (char*)s->hbuf; char* db = (char*)s->dbuf;
    double* hX = (double*)(hb + o_X);
    memset(hX, 0, sizeof(double) * (X_FEAT + nfeat));
    memcpy(hX + X_POSE, p->para_pose, sizeof(double) * 7 * F);
    if (p->para_speed_bias) memcpy(hX + X_SB, p->para_speed_bias, sizeof(double) * 9 * F);
    memcpy(hX + X_EX, p->para_ex_pose, sizeof(double) * 7);
    hX[X_TD] = p->para_td[0];
    memcpy(hX + X_FEAT, p->para_feature, sizeof(double) * nfeat);
    {   // factors sorted by pair
        gf_ba_visual_factor* hv = (gf_ba_visual_factor*)(hb + o_vis);
        std::vector<int> fill(pair_start.begin(), pair_start.end() - 1);
        for (int v = 0; v < p->n_visual; v++) { int k = pair_id[p->visual[v].imu_i * F + p->visual[v].imu_j]; hv[fill[k]++] = p->visual[v]; }
    }
    if (p->n_imu) memcpy(hb + o_imu, p->imu, sizeof(gf_ba_imu_factor) * (size_t)p->n_imu);
    for (int k = 0; k < p->n_imu; k++) if (p->imu[k].i < 0 || p->imu[k].i >= F || p->imu[k].j < 0 || p->imu[k].j >= F) return set_err(GF_ERR_INVALID_ARG, "imu factor index out of range");
    memcpy(hb + o_ps, pair_start.data(), sizeof(int) * (n_pairs + 1));
    if (n_pairs) memcpy(hb + o_pij, pair_ij.data(), sizeof(int) * 2 * n_pairs);
    memcpy(hb + o_cf, col_feat.data(), sizeof(int) * (size_t)(nfeat > 0 ? nfeat : 1));
    if (pr) {
        memcpy(hb + o_pJ, pr->linearized_jacobians, sizeof(double) * (size_t)pn * pn);
        memcpy(hb + o_pr0, pr->linearized_residuals, sizeof(double) * pn);
        memcpy(hb + o_px0, pr->x0, sizeof(double) * px0_len);
        memcpy(hb + o_pcol, pcol.data(), sizeof(int) * pn);
    }
    d.col_feat = (const int*)(db + o_cf); d.X = (double*)(db + o_X); d.Xc = (double*)(db + o_Xc);
    d.vis = (const gf_ba_visual_factor*)(db + o_vis); d.pair_start = (const int*)(db + o_ps); d.pair_ij = (const int*)(db + o_pij);
    d.imu = (const gf_ba_imu_factor*)(db + o_imu); d.imu_sqrt = (double*)(db + o_sq);
    d.pJ = (const double*)(db + o_pJ); d.pr0 = (const double*)(db + o_pr0); d.px0 = (const double*)(db + o_px0); d.pcol = (const int*)(db + o_pcol);
    d.Hp = (double*)(db + o_Hp); d.acc[0] = (double*)(db + o_a0); d.acc[1] = (double*)(db + o_a1);
    double* vec = (double*)(db + o_vec);
    const size_t nn = (size_t)(n > 0 ? n : 1);
    d.scale = vec; d.diag = vec + nn; d.gs = vec + 2 * nn; d.gn = vec + 3 * nn; d.step = vec + 4 * nn; d.delta = vec + 5 * nn;
    d.st = (BaState*)(db + o_st);
    for (int k = 0; k < 3; k++) d.gravity[k] = p->gravity[k];
    d.vis_sqrt_info = p->visual_sqrt_info;

    cudaStream_t st = s->s;
    GF_CUDA(cudaEventRecord(s->e0, st));
    GF_CUDA(cudaMemcpyAsync(db, hb, upload_bytes, cudaMemcpyHostToDevice, st));
    GF_CUDA(cudaMemsetAsync(db + o_a0, 0, al(sizeof(double) * acc_size(nc, L)) * 2, st));
    GF_CUDA(cudaMemsetAsync(db + o_st, 0, sizeof(BaState), st));
    k_ba_setup<<<(nc * nc + 255) / 256 + 1, 256, 0, st>>>(d); GF_LAUNCHED();
    {   // max_iter into the state (after setup zeroed/initialised it)
        int mi = p->max_num_iterations;
        GF_CUDA(cudaMemcpyAsync((char*)d.st + offsetof(BaState, max_iter), &mi, sizeof(int), cudaMemcpyHostToDevice, st));
    }
    if (pn) { k_ba_prior_hessian<<<(pn * pn + 255) / 256, 256, 0, st>>>(d); GF_LAUNCHED(); }
    const int eval_blocks = n_pairs + p->n_imu + (pn ? 1 : 0);
    const size_t prior_smem = sizeof(double) * 2 * (size_t)pn;
    const size_t step_smem = sizeof(double) * (size_t)(nc + 1) * (nc + 2) / 2;
    const int iters = p->max_num_iterations;
    if (eval_blocks > 0) { k_ba_eval<<<eval_blocks, PAIR_THREADS, prior_smem, st>>>(d, 0); GF_LAUNCHED(); }
    for (int it = 0; it <= iters; it++) {
        k_ba_step<<<1, 1024, step_smem, st>>>(d); GF_LAUNCHED();
        if (it == iters) break;                  // the extra k_ba_step adopts the last linearisation and closes the run
        if (eval_blocks > 0) { k_ba_eval<<<eval_blocks, PAIR_THREADS, prior_smem, st>>>(d, 1); GF_LAUNCHED(); }
        k_ba_decide<<<1, 256, 0, st>>>(d); GF_LAUNCHED();
        if (eval_blocks > 0) { k_ba_eval<<<eval_blocks, PAIR_THREADS, prior_smem, st>>>(d, 0); GF_LAUNCHED(); }
    }
    GF_CUDA(cudaGetLastError());
    GF_CUDA(cudaMemcpyAsync(hb + o_X, db + o_X, sizeof(double) * (X_FEAT + nfeat), cudaMemcpyDeviceToHost, st));
    GF_CUDA(cudaMemcpyAsync(hb + o_st, db + o_st, sizeof(BaState), cudaMemcpyDeviceToHost, st));
    GF_CUDA(cudaEventRecord(s->e1, st));
    GF_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    GF_CUDA(cudaEventElapsedTime(&ms, s->e0, s->e1));
    const BaState* hs = (const BaState*)(hb + o_st);
    memcpy(p->para_pose, hX + X_POSE, sizeof(double) * 7 * F);
    if (p->para_speed_bias) memcpy(p->para_speed_bias, hX + X_SB, sizeof(double) * 9 * F);
    memcpy(p->para_ex_pose, hX + X_EX, sizeof(double) * 7);
    p->para_td[0] = hX[X_TD];
    memcpy(p->para_feature, hX + X_FEAT, sizeof(double) * nfeat);
    sum->iterations = hs->it; sum->num_successful_steps = hs->n_success; sum->termination = hs->termination;
    sum->reduced_dim = nc; sum->n_free_landmarks = L; sum->n_residuals = pn + 15 * p->n_imu + 2 * p->n_visual;
    sum->initial_cost = hs->cost_hist[0]; sum->final_cost = hs->x_cost;
    for (int k = 0; k <= GF_BA_MAX_ITERATIONS; k++) { sum->cost[k] = hs->cost_hist[k]; sum->radius[k] = hs->radius_hist[k]; }
    sum->device_ms = ms;
    return GF_OK;
}

}  // extern "C"
