// lbfgs.cuh — device-resident L-BFGS (HLBFGS) vectors, reductions and line-search state.
//
// Replaces HLBFGS() as configured by HLBFGSOptimizer::optimize
// (geogram/third_party/HLBFGS/HLBFGS.cpp:281-587, geogram/numerics/lbfgs_optimizers.cpp:159-196):
// standard L-BFGS two-loop recursion (INFO[3]=0, INFO[7]=0, INFO[10]=0) with the
// More-Thuente MCSRCH/MCSTEP line search (LineSearch.cpp:10-465, SAFE_SEARCH variant).
// All vectors (x, g, q, s/y history, line-search base point) and all scalars (f, step,
// rho/alpha, the line-search state) live in device memory; the scalar state machine runs
// inside the two kernels below; the host reads back one small record per function evaluation (the
// line-search status that drives control flow and the (f, |g|) pair reported to the iteration callback).
//
//   lbfgs_direction_kernel : ONE cooperative launch per iteration = history update (s, y, rho), two-loop
//                            recursion, Hessian scaling, saves of x and g, g.d, the start of MCSRCH and the
//                            first trial point x = wa + stp d. Every level of the recursion is one fused
//                            pass (axpy of the previous level + dot product of the next) ended by a grid
//                            barrier: 2 bound + 4 barriers instead of ~45 launches.
//   lbfgs_post_eval_kernel : after each function evaluation: f = sum of the per-seed energies, g.d, |g|, |x|,
//                            then the MCSRCH decision (last block).
#pragma once
#include "common.cuh"
#include "comm.cuh"
#include <cooperative_groups.h>

#define LBFGS_MAXM 32
#define LBFGS_GRAM_MAXM 8        // history sizes the one-reduction direction kernel serves (5 (M - 1) + 5 <= 40 dot products)
#define LBFGS_GRAM_MAXK 40

struct McsState {
    double dg, dgm, dginit, dgtest, dgx, dgxm, dgy, dgym, finit, fm, ftest1, fx, fxm, fy, fym;
    double stmax, stmin, stx, sty, width, width1;
    int infoc, brackt, stage1;
};

struct LbfgsScalars {
    // first 48 bytes = the record the host reads back after every evaluation
    double f;              // current function value
    double dot;            // g.d (dginit at the start of a line search, dg afterwards)
    double stp;
    double gnorm, xnorm;
    int info, nfev;        // MCSRCH status (-1: evaluate x = wa + stp d and come back), evaluations of this search
    double rho[LBFGS_MAXM];
    McsState L;
    unsigned int red_counter;
    // Gram matrices of the history, indexed by slot: SY[a][b] = s_a . y_b, YY[a][b] = y_a . y_b (lbfgs_direction_gram_kernel)
    double SY[LBFGS_GRAM_MAXM][LBFGS_GRAM_MAXM], YY[LBFGS_GRAM_MAXM][LBFGS_GRAM_MAXM];
};

__device__ __forceinline__ double block_sum(double v, double* sm) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = (lane < (int)(blockDim.x >> 5)) ? sm[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// x = wa + stp * s  (LineSearch.cpp:107-108)
__global__ void step_kernel(u32 n, const LbfgsScalars* sc, const double* wa, const double* s, double* x) {
    const double stp = sc->stp;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double t = wa[i];
        t += stp * s[i];
        x[i] = t;
    }
}

__device__ inline double dmin_(double a, double b) { return a < b ? a : b; }
__device__ inline double dmax_(double a, double b) { return a > b ? a : b; }

// MCSTEP, SAFE_SEARCH variant — LineSearch.cpp:232-465
__device__ inline void mcstep_dev(double* stx, double* fx, double* dx, double* sty, double* fy, double* dy,
                                  double* stp, const double* fp, const double* dp, int* brackt,
                                  const double* stpmin, const double* stpmax, int* info) {
    double p, q, r, s, gama, sgnd, stpc, stpf, stpq, theta, t;
    int bound;
    const double xsafe = .001;
    *info = 0;
    if ((*brackt && (*stp <= dmin_(*stx, *sty) || *stp >= dmax_(*stx, *sty)))
        || *dx * (*stp - *stx) >= 0. || *stpmax < *stpmin) return;
    sgnd = *dp * (*dx / fabs(*dx));
    if (*fp > *fx) {
        *info = 1; bound = 1;
        theta = (*fx - *fp) * 3 / (*stp - *stx) + *dx + *dp;
        s = dmax_(dmax_(fabs(theta), fabs(*dx)), fabs(*dp));
        t = theta / s;
        gama = s * sqrt(t * t - *dx / s * (*dp / s));
        if (*stp < *stx) gama = -gama;
        p = gama - *dx + theta;
        q = gama - *dx + gama + *dp;
        r = p / q;
        stpc = *stx + r * (*stp - *stx);
        stpq = *stx + *dx / ((*fx - *fp) / (*stp - *stx) + *dx) / 2 * (*stp - *stx);
        if (fabs(stpc - *stx) < fabs(stpq - *stx)) stpf = stpc;
        else stpf = stpc + (stpq - stpc) / 2;
        if (*stp > *stx) stpf = dmax_(*stx + xsafe * (*stp - *stx), stpf);
        else stpf = dmin_(*stx + xsafe * (*stp - *stx), stpf);
        *brackt = 1;
    } else if (sgnd < 0.) {
        *info = 2; bound = 0;
        theta = (*fx - *fp) * 3 / (*stp - *stx) + *dx + *dp;
        s = dmax_(dmax_(fabs(theta), fabs(*dx)), fabs(*dp));
        t = theta / s;
        gama = s * sqrt(t * t - *dx / s * (*dp / s));
        if (*stp > *stx) gama = -gama;
        p = gama - *dp + theta;
        q = gama - *dp + gama + *dx;
        r = p / q;
        stpc = *stp + r * (*stx - *stp);
        stpq = *stp + *dp / (*dp - *dx) * (*stx - *stp);
        if (fabs(stpc - *stp) > fabs(stpq - *stp)) stpf = stpc; else stpf = stpq;
        *brackt = 1;
    } else if (fabs(*dp) < fabs(*dx)) {
        *info = 3; bound = 1;
        theta = (*fx - *fp) * 3 / (*stp - *stx) + *dx + *dp;
        s = dmax_(dmax_(fabs(theta), fabs(*dx)), fabs(*dp));
        t = theta / s;
        gama = s * sqrt(dmax_(0., t * t - *dx / s * (*dp / s)));
        if (*stp > *stx) gama = -gama;
        p = gama - *dp + theta;
        q = gama + (*dx - *dp) + gama;
        r = p / q;
        if (r < 0. && gama != 0.) stpc = *stp + r * (*stx - *stp);
        else if (*stp > *stx) stpc = *stpmax;
        else stpc = *stpmin;
        stpq = *stp + *dp / (*dp - *dx) * (*stx - *stp);
        if (*brackt) { if (fabs(*stp - stpc) < fabs(*stp - stpq)) stpf = stpc; else stpf = stpq; }
        else { if (fabs(*stp - stpc) > fabs(*stp - stpq)) stpf = stpc; else stpf = stpq; }
    } else {
        *info = 4; bound = 0;
        if (*brackt) {
            theta = (*fp - *fy) * 3 / (*sty - *stp) + *dy + *dp;
            s = dmax_(dmax_(fabs(theta), fabs(*dy)), fabs(*dp));
            t = theta / s;
            gama = s * sqrt(t * t - *dy / s * (*dp / s));
            if (*stp > *sty) gama = -gama;
            p = gama - *dp + theta;
            q = gama - *dp + gama + *dy;
            r = p / q;
            stpc = *stp + r * (*sty - *stp);
            stpf = stpc;
        } else if (*stp > *stx) stpf = *stpmax;
        else stpf = *stpmin;
    }
    sgnd = *dp * (*stx - *stp);
    if (*fp > *fx) { *sty = *stp; *fy = *fp; *dy = *dp; }
    else {
        if (sgnd < 0.) { *sty = *stx; *fy = *fx; *dy = *dx; }
        *stx = *stp; *fx = *fp; *dx = *dp;
    }
    stpf = dmin_(*stpmax, stpf);
    stpf = dmax_(*stpmin, stpf);
    *stp = stpf;
    if (*brackt && bound) {
        if (*sty > *stx) *stp = dmin_(*stx + (*sty - *stx) * .66, *stp);
        else *stp = dmax_(*stx + (*sty - *stx) * .66, *stp);
    }
}

// MCSRCH scalar part — LineSearch.cpp:10-230. sc->dot holds g.s (dginit at start, dg on resume).
// sc->info: in -1 = resume after an evaluation, else start. Out: -1 asks for x = wa + stp*s and an
// evaluation; any other value ends the line search. The vector parts (wa = x at start; x = wa + stp*s)
// are separate launches.
__device__ inline void mcsrch_dev(LbfgsScalars* sc, u32 n) {
    const double ftol = 1.0e-4, xtol = 1.0e-16, gtol = 0.9, stpmin = 1.0e-20, stpmax = 1.0e+20;
    const int maxfev = 20;
    McsState* L = &sc->L;
    double* stp = &sc->stp;
    const double f = sc->f;
    if (sc->info != -1) {
        L->infoc = 1;
        if (n == 0 || *stp <= 0.) return;
        L->dginit = sc->dot;
        if (L->dginit >= 0.) return;
        L->brackt = 0; L->stage1 = 1; sc->nfev = 0;
        L->finit = f; L->dgtest = ftol * L->dginit;
        L->width = stpmax - stpmin; L->width1 = L->width / .5;
        L->stx = 0.; L->fx = L->finit; L->dgx = L->dginit;
        L->sty = 0.; L->fy = L->finit; L->dgy = L->dginit;
    } else {
        int info = 0;
        ++sc->nfev;
        L->dg = sc->dot;
        L->ftest1 = L->finit + *stp * L->dgtest;
        if ((L->brackt && (*stp <= L->stmin || *stp >= L->stmax)) || L->infoc == 0) info = 6;
        if (*stp == stpmax && f <= L->ftest1 && L->dg <= L->dgtest) info = 5;
        if (*stp == stpmin && (f > L->ftest1 || L->dg >= L->dgtest)) info = 4;
        if (sc->nfev >= maxfev) info = 3;
        if (L->brackt && L->stmax - L->stmin <= xtol * L->stmax) info = 2;
        if (f <= L->ftest1 && fabs(L->dg) <= gtol * (-L->dginit)) info = 1;
        sc->info = info;
        if (info != 0) return;
        if (L->stage1 && f <= L->ftest1 && L->dg >= dmin_(ftol, gtol) * L->dginit) L->stage1 = 0;
        if (L->stage1 && f <= L->fx && f > L->ftest1) {
            L->fm = f - *stp * L->dgtest;
            L->fxm = L->fx - L->stx * L->dgtest;
            L->fym = L->fy - L->sty * L->dgtest;
            L->dgm = L->dg - L->dgtest;
            L->dgxm = L->dgx - L->dgtest;
            L->dgym = L->dgy - L->dgtest;
            mcstep_dev(&L->stx, &L->fxm, &L->dgxm, &L->sty, &L->fym, &L->dgym, stp, &L->fm, &L->dgm,
                       &L->brackt, &L->stmin, &L->stmax, &L->infoc);
            L->fx = L->fxm + L->stx * L->dgtest;
            L->fy = L->fym + L->sty * L->dgtest;
            L->dgx = L->dgxm + L->dgtest;
            L->dgy = L->dgym + L->dgtest;
        } else {
            double fp = f;
            mcstep_dev(&L->stx, &L->fx, &L->dgx, &L->sty, &L->fy, &L->dgy, stp, &fp, &L->dg,
                       &L->brackt, &L->stmin, &L->stmax, &L->infoc);
        }
        if (L->brackt) {
            if (fabs(L->sty - L->stx) >= L->width1 * .66) *stp = L->stx + (L->sty - L->stx) * .5;
            L->width1 = L->width;
            L->width = fabs(L->sty - L->stx);
        }
    }
    if (L->brackt) { L->stmin = dmin_(L->stx, L->sty); L->stmax = dmax_(L->stx, L->sty); }
    else { L->stmin = L->stx; L->stmax = *stp + (*stp - L->stx) * 4.; }
    *stp = dmax_(*stp, stpmin);
    *stp = dmin_(*stp, stpmax);
    if ((L->brackt && (*stp <= L->stmin || *stp >= L->stmax)) || sc->nfev >= maxfev - 1 || L->infoc == 0
        || (L->brackt && L->stmax - L->stmin <= xtol * L->stmax)) *stp = L->stx;
    sc->info = -1;
}

// ---------------------------------------------------------------------------------------
// fused direction kernel
// ---------------------------------------------------------------------------------------
#ifndef LBFGS_DIR_THREADS
#define LBFGS_DIR_THREADS 512
#endif
#ifndef LBFGS_DIR_BLOCKS_PER_SM
#define LBFGS_DIR_BLOCKS_PER_SM 2
#endif

struct LbfgsDirArgs {
    u32 N;
    int first;               // 1: first iteration (no history pair to add, step 1/|g|)
    int M;                   // history size (0: steepest descent)
    int cur_pos;             // slot that receives the new (s, y) pair
    int bound;               // levels of the recursion - 1 (-1: none)
    int st1[LBFGS_MAXM];     // history slot of level i in the first loop (i = bound .. 0)
    int st2[LBFGS_MAXM];     // history slot of level i in the second loop (i = 0 .. bound)
    double* x; double* g; double* q; double* px; double* pg; double* wa;
    double* s; double* y;    // [M][N]
    LbfgsScalars* sc;
    double* partials;        // [2][3 * gridDim.x]
    // Sharded runs: N and all vectors are this rank's slice (a contiguous range of original seed indices); every dot
    // product is the local partial followed by a sum over the ranks through the peer mailboxes (comm.cuh).
    PeerComm pc;
    u32 N_global;
};

// Sum over the grid of up to three per-thread values: fixed partition (grid-stride), block tree, then every block
// adds the per-block partials in the same order, so all blocks hold bit-identical totals. One grid barrier.
// `slot` alternates between consecutive calls (a block may enter the next reduction while others still read this one).
template <int K>
__device__ __forceinline__ void grid_sums(cooperative_groups::grid_group& grid, double* v, double* partials, int slot,
                                          double* sm, double* out, const PeerComm& pc) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* mine = partials + (size_t)slot * 3 * gridDim.x;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const double r = block_sum(v[j], sm);
        if (threadIdx.x == 0) mine[(size_t)j * gridDim.x + blockIdx.x] = r;
        __syncthreads();
    }
    __threadfence();
    grid.sync();
    if (w == 0) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            double p = 0.0;
            for (u32 b = lane; b < gridDim.x; b += 32) p += __ldcg(mine + (size_t)j * gridDim.x + b);
            p = warp_sum(p);
            if (lane == 0) sm[32 + j] = p;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K; ++j) out[j] = sm[32 + j];
    __syncthreads();
    if (pc.nranks > 1) {
        // the local totals (identical in every block) become global ones: block 0 exchanges them with the peers and
        // publishes the sums; one more grid barrier
        if (blockIdx.x == 0 && w == 0) {
            double tot[K];
            peer_allreduce<K>(pc, out, tot);
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < K; ++j) pc.gtot[slot * 8 + j] = tot[j];
                __threadfence();
            }
        }
        grid.sync();
#pragma unroll
        for (int j = 0; j < K; ++j) out[j] = __ldcg(pc.gtot + slot * 8 + j);
    }
}

// HLBFGS.cpp:356-533 between two line searches, on the device (see the header of this file)
__global__ void __launch_bounds__(LBFGS_DIR_THREADS)
lbfgs_direction_kernel(const __grid_constant__ LbfgsDirArgs a) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double sm[40];
    const u32 N = a.N;
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    LbfgsScalars* sc = a.sc;
    int slot = 0;
    double v[3], tot[3];
    const bool hist = !a.first && a.M > 0;
    if (!hist) {
        // q = -g; |g|^2 (first step 1/|g|, HLBFGS.cpp:520-523) and g.q
        v[0] = 0.0; v[1] = 0.0;
        for (u32 i = tid; i < N; i += nth) {
            const double gi = a.g[i], qi = -gi;
            a.q[i] = qi; a.px[i] = a.x[i]; a.pg[i] = gi; a.wa[i] = a.x[i];
            v[0] += gi * gi; v[1] += gi * qi;
        }
        grid_sums<2>(grid, v, a.partials, slot, sm, tot, a.pc); slot ^= 1;
        if (tid == 0) {
            if (a.first) { sc->gnorm = sqrt(tot[0]); sc->stp = 1.0 / sc->gnorm; } else sc->stp = 1.0;
            sc->dot = tot[1];
        }
    } else {
        double* s_cur = a.s + (size_t)a.cur_pos * N;
        double* y_cur = a.y + (size_t)a.cur_pos * N;
        const int bound = a.bound;
        // new pair s = x - prev_x, y = g - prev_g (HLBFGS.cpp:375-379); y.s (rho, :380; Hessian scaling ys, :103),
        // y.y (:108); q = -g and the first dot product of the first loop, s_st.q with st = st1[bound]
        {
            const double* s0 = a.s + (size_t)a.st1[bound] * N;     // == s_cur: the newest pair comes first
            v[0] = 0.0; v[1] = 0.0; v[2] = 0.0;
            for (u32 i = tid; i < N; i += nth) {
                const double gi = a.g[i];
                const double si = a.x[i] - a.px[i], yi = gi - a.pg[i];
                s_cur[i] = si; y_cur[i] = yi;
                const double qi = -gi;
                a.q[i] = qi;
                v[0] += yi * si; v[1] += yi * yi;
                v[2] += qi * ((s0 == s_cur) ? si : s0[i]);
            }
            grid_sums<3>(grid, v, a.partials, slot, sm, tot, a.pc); slot ^= 1;
        }
        const double ys = tot[0], yy = tot[1];
        const double rho_cur = 1.0 / ys;
        if (tid == 0) sc->rho[a.cur_pos] = rho_cur;
        // first loop (HLBFGS_UPDATE_First_Step, :160-176): alpha_i = rho_st (s_st.q); q -= alpha_i y_st
        double alpha[LBFGS_MAXM];
        double dotv = tot[2];
        for (int i = bound; i >= 0; --i) {
            const int st = a.st1[i];
            const double rho = (st == a.cur_pos) ? rho_cur : sc->rho[st];
            alpha[i] = rho * dotv;
            const double c = -alpha[i];
            const double* yv = a.y + (size_t)st * N;
            v[0] = 0.0;
            if (i > 0) {
                const double* sn = a.s + (size_t)a.st1[i - 1] * N;
                for (u32 k = tid; k < N; k += nth) {
                    double qk = a.q[k];
                    qk += c * yv[k];
                    a.q[k] = qk;
                    v[0] += qk * sn[k];
                }
            } else {
                // last level, then the Hessian scaling q *= ys/yy (HLBFGS_UPDATE_Hessian, :90-110) and the first dot
                // product of the second loop, y_st.q
                const double factor = ys / yy;
                const double* yn = a.y + (size_t)a.st2[0] * N;
                for (u32 k = tid; k < N; k += nth) {
                    double qk = a.q[k];
                    qk += c * yv[k];
                    qk *= factor;
                    a.q[k] = qk;
                    v[0] += yn[k] * qk;
                }
            }
            grid_sums<1>(grid, v, a.partials, slot, sm, tot, a.pc); slot ^= 1;
            dotv = tot[0];
        }
        // second loop (HLBFGS_UPDATE_Second_Step, :178-196): q += (alpha_i - rho_st (y_st.q)) s_st
        for (int i = 0; i <= bound; ++i) {
            const int st = a.st2[i];
            const double rho = (st == a.cur_pos) ? rho_cur : sc->rho[st];
            const double c = alpha[i] - rho * dotv;
            const double* sv = a.s + (size_t)st * N;
            v[0] = 0.0;
            if (i < bound) {
                const double* yn = a.y + (size_t)a.st2[i + 1] * N;
                for (u32 k = tid; k < N; k += nth) {
                    double qk = a.q[k];
                    qk += c * sv[k];
                    a.q[k] = qk;
                    v[0] += yn[k] * qk;
                }
            } else {
                // direction complete: save x and g (:502-503), wa = x (LineSearch.cpp:84), g.q
                for (u32 k = tid; k < N; k += nth) {
                    double qk = a.q[k];
                    qk += c * sv[k];
                    a.q[k] = qk;
                    const double gk = a.g[k], xk = a.x[k];
                    a.px[k] = xk; a.pg[k] = gk; a.wa[k] = xk;
                    v[0] += gk * qk;
                }
            }
            grid_sums<1>(grid, v, a.partials, slot, sm, tot, a.pc); slot ^= 1;
            dotv = tot[0];
        }
        if (tid == 0) { sc->stp = 1.0; sc->dot = dotv; }
    }
    // start of MCSRCH (LineSearch.cpp:60-100): info becomes -1 when a trial point is wanted
    if (tid == 0) {
        sc->info = 0;
        mcsrch_dev(sc, a.N_global);
        __threadfence();
    }
    grid.sync();
    const int info = *((volatile int*)&sc->info);
    if (info != -1) return;
    const double stp = *((volatile double*)&sc->stp);
    for (u32 i = tid; i < N; i += nth) {
        double t = a.wa[i];
        t += stp * a.q[i];
        a.x[i] = t;
    }
}

// ---------------------------------------------------------------------------------------
// The same update with ONE reduction (history of at most LBFGS_GRAM_MAXM pairs). The two-loop recursion needs 2 (bound + 1)
// dot products in sequence, each a grid barrier here and a round trip over NVLink on several GPUs (117 us per iteration on one
// B200, 300 us on two). All of them are linear in quantities that are known up front: with q0 = -g and the Gram matrices
// SY[a][b] = s_a.y_b, YY[a][b] = y_a.y_b of the stored pairs,
//     alpha_i = rho_i (s_i.q0 - sum_{j > i} alpha_j SY[i][j])                                             (first loop, newest to oldest)
//     beta_i  = rho_i (gamma (y_i.q0 - sum_j alpha_j YY[i][j]) + sum_{j < i} (alpha_j - beta_j) SY[j][i])  (second loop)
//     d       = gamma q0 - sum_j gamma alpha_j y_j + sum_j (alpha_j - beta_j) s_j,   gamma = ys / yy of the newest pair
//     g.d     = -(gamma |g|^2 - gamma sum_j alpha_j (y_j.q0) + sum_j (alpha_j - beta_j) (s_j.q0))
// The Gram entries of old pairs never change; an iteration adds the row and column of the new pair. So: one pass computes the
// 5 (bound) + 5 dot products that involve g or the new pair, one reduction (one exchange between GPUs), the scalars are
// evaluated by every block, one pass writes d, the saves and the first trial point. Same mathematics as
// HLBFGS_UPDATE_First_Step / Hessian / Second_Step (HLBFGS.cpp:90-196); the rounding differs from the level-by-level
// evaluation the way a different summation order does.
// ---------------------------------------------------------------------------------------
#ifndef LBFGS_GRAM_MINBLK
#define LBFGS_GRAM_MINBLK 1
#endif
__global__ void __launch_bounds__(LBFGS_DIR_THREADS, LBFGS_GRAM_MINBLK)
lbfgs_direction_gram_kernel(const __grid_constant__ LbfgsDirArgs a) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double sm_w[LBFGS_GRAM_MAXK][LBFGS_DIR_THREADS / 32];
    __shared__ double sm_tot[LBFGS_GRAM_MAXK];
    __shared__ double sm_ga[LBFGS_GRAM_MAXM], sm_c[LBFGS_GRAM_MAXM], sm_misc[2];
    const u32 N = a.N;
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    LbfgsScalars* sc = a.sc;
    const int bound = a.bound;                       // levels 0 .. bound, oldest to newest; level bound is the new pair
    const int K = 5 + 5 * bound;
    double* s_cur = a.s + (size_t)a.cur_pos * N;
    double* y_cur = a.y + (size_t)a.cur_pos * N;
    // pass A: the new pair and every dot product that involves it or g
    double acc[LBFGS_GRAM_MAXK];
#pragma unroll
    for (int j = 0; j < LBFGS_GRAM_MAXK; ++j) acc[j] = 0.0;
    for (u32 i = tid; i < N; i += nth) {
        const double gi = a.g[i];
        const double si = a.x[i] - a.px[i], yi = gi - a.pg[i];
        s_cur[i] = si; y_cur[i] = yi;
        const double q0 = -gi;
        acc[0] += yi * si; acc[1] += yi * yi; acc[2] += gi * gi; acc[3] += si * q0; acc[4] += yi * q0;
#pragma unroll
        for (int j = 0; j < LBFGS_GRAM_MAXM - 1; ++j) {
            if (j < bound) {
                const double sj = a.s[(size_t)a.st1[j] * N + i], yj = a.y[(size_t)a.st1[j] * N + i];
                acc[5 + 5 * j] += sj * q0; acc[6 + 5 * j] += yj * q0;
                acc[7 + 5 * j] += si * yj; acc[8 + 5 * j] += sj * yi; acc[9 + 5 * j] += yi * yj;
            }
        }
    }
    // block sums (fixed tree), one partial per block and value
#pragma unroll
    for (int j = 0; j < LBFGS_GRAM_MAXK; ++j) {
        if (j < K) {
            const double r = warp_sum(acc[j]);
            if (lane == 0) sm_w[j][w] = r;
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < K) {
        double r = 0.0;
        for (int i = 0; i < LBFGS_DIR_THREADS / 32; ++i) r += sm_w[threadIdx.x][i];
        a.partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = r;
    }
    __threadfence();
    grid.sync();
    // block j adds the partials of value j in a fixed order
    double* totals = a.partials + (size_t)LBFGS_GRAM_MAXK * gridDim.x;
    if ((int)blockIdx.x < K) {
        double p = 0.0;
        for (u32 b = threadIdx.x; b < gridDim.x; b += blockDim.x) p += __ldcg(a.partials + (size_t)blockIdx.x * gridDim.x + b);
        const double r = block_sum(p, &sm_w[0][0]);
        if (threadIdx.x == 0) { totals[blockIdx.x] = r; __threadfence(); }
    }
    grid.sync();
    if (a.pc.nranks > 1) {
        // the local totals become global ones: one exchange with the peers for all of them
        if (blockIdx.x == 0 && w == 0) {
            peer_allreduce_n(a.pc, totals, K, totals);
            __threadfence();
        }
        grid.sync();
    }
    if ((int)threadIdx.x < K) sm_tot[threadIdx.x] = __ldcg(totals + threadIdx.x);
    __syncthreads();
    // scalars, by every block
    if (threadIdx.x == 0) {
        const double ys = sm_tot[0], yy = sm_tot[1], gg = sm_tot[2];
        const int cur = a.cur_pos;
        double SY[LBFGS_GRAM_MAXM][LBFGS_GRAM_MAXM], YY[LBFGS_GRAM_MAXM][LBFGS_GRAM_MAXM];   // by level
        double u[LBFGS_GRAM_MAXM], wq[LBFGS_GRAM_MAXM], rho[LBFGS_GRAM_MAXM], alpha[LBFGS_GRAM_MAXM], beta[LBFGS_GRAM_MAXM];
        for (int i = 0; i < bound; ++i) {
            for (int j = 0; j < bound; ++j) { SY[i][j] = sc->SY[a.st1[i]][a.st1[j]]; YY[i][j] = sc->YY[a.st1[i]][a.st1[j]]; }
            u[i] = sm_tot[5 + 5 * i]; wq[i] = sm_tot[6 + 5 * i];
            SY[bound][i] = sm_tot[7 + 5 * i]; SY[i][bound] = sm_tot[8 + 5 * i];
            YY[bound][i] = sm_tot[9 + 5 * i]; YY[i][bound] = sm_tot[9 + 5 * i];
            rho[i] = sc->rho[a.st1[i]];
        }
        SY[bound][bound] = ys; YY[bound][bound] = yy; u[bound] = sm_tot[3]; wq[bound] = sm_tot[4]; rho[bound] = 1.0 / ys;
        for (int i = bound; i >= 0; --i) {
            double t = u[i];
            for (int j = bound; j > i; --j) t -= alpha[j] * SY[i][j];
            alpha[i] = rho[i] * t;
        }
        const double gamma = ys / yy;
        double dotq = gamma * gg;                    // q0 . d, term by term
        for (int i = 0; i <= bound; ++i) {
            double t = wq[i];
            for (int j = bound; j >= 0; --j) t -= alpha[j] * YY[i][j];
            t *= gamma;
            for (int j = 0; j < i; ++j) t += (alpha[j] - beta[j]) * SY[j][i];
            beta[i] = rho[i] * t;
        }
        for (int i = 0; i <= bound; ++i) {
            sm_ga[i] = gamma * alpha[i]; sm_c[i] = alpha[i] - beta[i];
            dotq -= sm_ga[i] * wq[i]; dotq += sm_c[i] * u[i];
        }
        sm_misc[0] = gamma; sm_misc[1] = -dotq;      // g . d
        if (blockIdx.x == 0) {
            // the new pair's row and column stay for the next iterations; start of MCSRCH (LineSearch.cpp:60-100)
            for (int i = 0; i < bound; ++i) {
                sc->SY[cur][a.st1[i]] = SY[bound][i]; sc->SY[a.st1[i]][cur] = SY[i][bound];
                sc->YY[cur][a.st1[i]] = YY[bound][i]; sc->YY[a.st1[i]][cur] = YY[i][bound];
            }
            sc->SY[cur][cur] = ys; sc->YY[cur][cur] = yy;
            sc->rho[cur] = 1.0 / ys;
            sc->stp = 1.0; sc->dot = -dotq;
            sc->info = 0;
            mcsrch_dev(sc, a.N_global);
        }
    }
    __syncthreads();
    // pass B: the direction, the saves (HLBFGS.cpp:502-503, LineSearch.cpp:84) and the first trial point x = wa + stp d.
    // The search starts (info = -1, stp = 1) exactly when g.d < 0 (mcsrch_dev, start branch)
    const double gamma = sm_misc[0];
    const bool started = sm_misc[1] < 0.0 && a.N_global != 0;
    for (u32 k = tid; k < N; k += nth) {
        const double gk = a.g[k], xk = a.x[k];
        double d = gamma * (-gk);
#pragma unroll
        for (int j = 0; j < LBFGS_GRAM_MAXM; ++j) {
            if (j <= bound) {
                d -= sm_ga[j] * a.y[(size_t)a.st1[j] * N + k];
                d += sm_c[j] * a.s[(size_t)a.st1[j] * N + k];
            }
        }
        a.q[k] = d; a.px[k] = xk; a.pg[k] = gk; a.wa[k] = xk;
        if (started) { double t = xk; t += 1.0 * d; a.x[k] = t; }
    }
}

// ---------------------------------------------------------------------------------------
// after a function evaluation
// ---------------------------------------------------------------------------------------
#define LBFGS_POST_BLOCKS 592
#define LBFGS_POST_THREADS 256

// f = sum_i fs[i] (per-seed energies, ns entries); g.q, g.g, x.x over n entries; then (resume = 1) the MCSRCH decision.
// Deterministic: fixed partition, per-block partials added in order by the last block.
__global__ void __launch_bounds__(LBFGS_POST_THREADS)
lbfgs_post_eval_kernel(u32 ns, const double* fs, u32 n, const double* g, const double* q, const double* x,
                       double* partials, LbfgsScalars* sc, int resume, PeerComm pc, u32 n_global, const u32* pending) {
    __shared__ double sm[32];
    __shared__ bool last;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) v[0] += fs[i];
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double gi = g[i], xi = x[i];
        v[1] += gi * q[i]; v[2] += gi * gi; v[3] += xi * xi;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double r = block_sum(v[j], sm);
        if (threadIdx.x == 0) partials[(size_t)j * gridDim.x + blockIdx.x] = r;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&sc->red_counter, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double tot[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double p = 0.0;
        for (u32 i = threadIdx.x; i < gridDim.x; i += blockDim.x) p += __ldcg(partials + (size_t)j * gridDim.x + i);
        __syncthreads();
        tot[j] = block_sum(p, sm);
        __syncthreads();
    }
    if (pc.nranks > 1) {
        // sharded: fs / g, q, x are this rank's seeds / slice; totals over the ranks through the peer mailboxes
        if (threadIdx.x < 32) {
            double gt[4];
            peer_allreduce<4>(pc, tot, gt);
#pragma unroll
            for (int j = 0; j < 4; ++j) tot[j] = gt[j];
        }
    }
    if (threadIdx.x == 0) {
        sc->red_counter = 0;
        sc->f = tot[0];
        sc->dot = tot[1];
        sc->gnorm = sqrt(tot[2]);
        sc->xnorm = sqrt(tot[3]);
        // *pending != 0: some seeds of this evaluation still wait for longer neighbour lists (the host reads that count with this
        // record): the evaluation is not complete, the search is left untouched and marked (info = -2); the kernel runs again
        // when the lists have been enlarged
        const bool incomplete = pending && *pending != 0u;
        if (incomplete) { if (sc->info == -1) sc->info = -2; }
        else {
            if (sc->info == -2) sc->info = -1;
            // a line search that could not start (info != -1 after the direction kernel) stays as it is
            if (resume && sc->info == -1) mcsrch_dev(sc, n_global);
        }
    }
}
