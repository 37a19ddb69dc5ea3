// lbfgs.cuh — device-resident L-BFGS (HLBFGS) vectors, reductions and line-search state.
//
// Replaces HLBFGS() as configured by HLBFGSOptimizer::optimize
// (geogram/third_party/HLBFGS/HLBFGS.cpp:281-587, geogram/numerics/lbfgs_optimizers.cpp:159-196):
// standard L-BFGS two-loop recursion (INFO[3]=0, INFO[7]=0, INFO[10]=0) with the
// More-Thuente MCSRCH/MCSTEP line search (LineSearch.cpp:10-465, SAFE_SEARCH variant).
// All vectors (x, g, q, s/y history, line-search base point) and all scalars (f, step,
// rho/alpha, the line-search state) live in device memory; the scalar state machine runs
// in a single-thread kernel; the host only reads back the 4-byte status that drives control
// flow and the (f, |g|) pair reported to the iteration callback.
#pragma once
#include "common.cuh"

#define LBFGS_MAXM 32
#define LBFGS_RED_BLOCKS 1024
#define LBFGS_RED_THREADS 256

struct McsState {
    double dg, dgm, dginit, dgtest, dgx, dgxm, dgy, dgym, finit, fm, ftest1, fx, fxm, fy, fym;
    double stmax, stmin, stx, sty, width, width1;
    int infoc, brackt, stage1;
};

struct LbfgsScalars {
    double f;              // current function value (written by the evaluation)
    double dot;            // last reduction result
    double stp;
    double gnorm, xnorm;
    double coef;           // coefficient for the next axpy / scale
    double rho[LBFGS_MAXM];
    double alpha[LBFGS_MAXM];
    McsState L;
    int info, nfev;
    unsigned int red_counter;
};

// epilogues executed by the last block of a reduction
enum { RED_STORE = 0, RED_RHO, RED_ALPHA, RED_BETA, RED_YS, RED_FACTOR, RED_GNORM, RED_XNORM, RED_F, RED_STP0 };

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

// sum_i a[i]*b[i] (b == NULL: sum a[i]); deterministic: fixed partition, partials summed in order.
__global__ void __launch_bounds__(LBFGS_RED_THREADS)
reduce_kernel(u32 n, const double* a, const double* b, double* partials, LbfgsScalars* sc, int op, int i0, int i1) {
    __shared__ double sm[32];
    __shared__ bool last;
    double v = 0.0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        v += b ? a[i] * b[i] : a[i];
    double r = block_sum(v, sm);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = r;
        __threadfence();
        unsigned int t = atomicAdd(&sc->red_counter, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double p = 0.0;
        for (u32 i = threadIdx.x; i < gridDim.x; i += blockDim.x) p += partials[i];
        __syncthreads();
        double tot = block_sum(p, sm);
        if (threadIdx.x == 0) {
            sc->red_counter = 0;
            sc->dot = tot;
            switch (op) {
            case RED_RHO:    sc->rho[i0] = 1.0 / tot; break;                       // HLBFGS.cpp:380
            case RED_ALPHA:  sc->alpha[i0] = sc->rho[i1] * tot; sc->coef = -sc->alpha[i0]; break;   // :171-173
            case RED_BETA:   sc->coef = sc->alpha[i0] - sc->rho[i1] * tot; break;   // :192
            case RED_YS:     sc->coef = tot; break;                                // ys, :103
            case RED_FACTOR: sc->coef = sc->coef / tot; break;                     // ys/yy, :108
            case RED_GNORM:  sc->gnorm = sqrt(tot); break;
            case RED_XNORM:  sc->xnorm = sqrt(tot); break;
            case RED_F:      sc->f = tot; break;
            case RED_STP0:   sc->gnorm = sqrt(tot); sc->stp = 1.0 / sc->gnorm; break; // :520-523
            default: break;
            }
        }
    }
}

// y += coef * x with coef read from the device scalars (HLBFGS_DAXPY, HLBFGS_BLAS.cpp:42-52)
__global__ void axpy_dev_kernel(u32 n, const LbfgsScalars* sc, const double* x, double* y) {
    const double c = sc->coef;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] += c * x[i];
}
__global__ void scale_dev_kernel(u32 n, const LbfgsScalars* sc, double* x) {
    const double c = sc->coef;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] *= c;
}
__global__ void neg_kernel(u32 n, const double* g, double* q) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) q[i] = -g[i];
}
// s = x - prev_x ; y = g - prev_g  (HLBFGS.cpp:375-379)
__global__ void diff_kernel(u32 n, const double* x, const double* px, const double* g, const double* pg, double* s, double* y) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        s[i] = x[i] - px[i];
        y[i] = g[i] - pg[i];
    }
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
__global__ void mcsrch_kernel(LbfgsScalars* sc, u32 n) {
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

__global__ void set_info_kernel(LbfgsScalars* sc, int info, double stp, int set_stp) {
    sc->info = info;
    if (set_stp) sc->stp = stp;
}
