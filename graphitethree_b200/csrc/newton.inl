// newton.inl — b200cvt_newton: CentroidalVoronoiTesselation::Newton_iterations
// (geogram/voronoi/CVT.cpp:272-338) = HLBFGS (geogram/third_party/HLBFGS/HLBFGS.cpp:281-587) over
// funcgrad = set_vertices + compute_CVT_func_grad(check_SR=true) + constrain_points.
// Included at the end of b200cvt.cu.

static void lb_reduce(b200cvt_ctx* h, u32 n, const double* a, const double* b, int op, int i0, int i1) {
    u32 blocks = std::min<u32>(div_up(n, LBFGS_RED_THREADS), LBFGS_RED_BLOCKS);
    LAUNCH(h, reduce_kernel, blocks, LBFGS_RED_THREADS, 0, n, a, b, h->lb_part.p, h->lb_sc.p, op, i0, i1);
}
static u32 lb_blocks(u32 n) { return std::min<u32>(div_up(n, 256), 148u * 8u); }

// funcgrad (CVT.cpp:323-338) on the device: seeds = h->x, result f -> scalars.f, g -> h->lb_g.
// With partitioned seeds every rank evaluates its Morton slice, the (g, f_seed) slices are
// all-gathered, and every rank then holds the full gradient and the same energy.
template <int D>
static void newton_eval_t(b200cvt_ctx* h) {
    h->grid_valid = false; h->knn_valid = false;
    evaluate(h, 1, 1);
    const u32 S = h->S;
    if (h->nranks == 1) {
        LAUNCH(h, scatter_results_kernel<D>, div_up(S, 256), 256, 0, (const SeedRec<D>*)h->xs.p, 0u, S, h->out_s.p, h->out_v.p,
               h->flags.p, h->pair_cnt.p, h->locked.p, 1, (double*)nullptr, h->lb_g.p, h->flags_orig.p, h->cnt_orig.p);
        lb_reduce(h, S, h->out_s.p, nullptr, RED_F, 0, 0);
    } else {
        if (!h->x_slice) throw StateError("seeds are partitioned but no exchange was set (b200cvt_set_exchange)");
        pack_slice<D>(h, h->out_s.p, h->out_v.p);
        run_exchange(h);
        h->s_orig.ensure(S);
        LAUNCH(h, unpack_all_kernel<D>, div_up(S, 256), 256, 0, (const SeedRec<D>*)h->xs.p, h->x_all, S, h->slice_len(),
               h->locked.p, 1, h->lb_g.p, h->s_orig.p);
        lb_reduce(h, S, h->s_orig.p, nullptr, RED_F, 0, 0);
    }
}
static void newton_eval(b200cvt_ctx* h) {
    if (h->dim == 3) newton_eval_t<3>(h); else newton_eval_t<6>(h);
}

// HLBFGS main loop (HLBFGS.cpp:356-586) on the device-resident seeds h->x
static void newton_loop(b200cvt_ctx* h, u32 nb_iter, u32 m, b200cvt_progress_cb cb, void* user, uint32_t* info_out) {
    if (m > LBFGS_MAXM) throw ArgError("m too large");
    if (nb_iter < 1) return;                 // HLBFGS: INFO[4] < 1 -> "check your input parameters", no work
    const int D = h->dim;
    const u32 S = h->S;
    const u32 N = S * (u32)D;
    const int M = (int)m;
    h->flags_orig.ensure(S); h->cnt_orig.ensure(S);
    h->lb_g.ensure(N); h->lb_q.ensure(N); h->lb_px.ensure(N); h->lb_pg.ensure(N); h->lb_wa.ensure(N);
    h->lb_s.ensure((size_t)std::max(M, 1) * N); h->lb_y.ensure((size_t)std::max(M, 1) * N);
    h->lb_part.ensure(LBFGS_RED_BLOCKS); h->lb_sc.ensure(1);
    CUDA_CHECK(cudaMemsetAsync(h->lb_sc.p, 0, sizeof(LbfgsScalars), h->stream));
    double* x = h->x.p; double* g = h->lb_g.p; double* q = h->lb_q.p;
    double* px = h->lb_px.p; double* pg = h->lb_pg.p; double* wa = h->lb_wa.p;
    LbfgsScalars* sc = h->lb_sc.p;
    const u32 nb = lb_blocks(N);
    const double stpmin = 1.0e-20, stpmax = 1.0e+20;

    u32 iter = 0, nfev_total = 0;
    int cur_pos = 0, bound = 0, ls_info = 0;
    bool canceled = false;
    struct { double f, dot, stp, gnorm, xnorm; } hs;
    for (;;) {
        if (iter == 0) { newton_eval(h); nfev_total++; }
        if (iter > 0 && M > 0) {
            double* s_cur = h->lb_s.p + (size_t)cur_pos * N;
            double* y_cur = h->lb_y.p + (size_t)cur_pos * N;
            LAUNCH(h, diff_kernel, nb, 256, 0, N, x, px, g, pg, s_cur, y_cur);
            lb_reduce(h, N, y_cur, s_cur, RED_RHO, cur_pos, 0);
        }
        LAUNCH(h, neg_kernel, nb, 256, 0, N, g, q);
        if (iter > 0 && M > 0) {
            bound = (int)iter > M ? M - 1 : (int)iter - 1;
            for (int i = bound; i >= 0; --i) {       // HLBFGS_UPDATE_First_Step
                int st = (int)iter <= M ? cur_pos - bound + i : (cur_pos - (bound - i) + M) % M;
                lb_reduce(h, N, q, h->lb_s.p + (size_t)st * N, RED_ALPHA, i, st);
                LAUNCH(h, axpy_dev_kernel, nb, 256, 0, N, sc, h->lb_y.p + (size_t)st * N, q);
            }
            {                                         // HLBFGS_UPDATE_Hessian: q *= ys/yy
                double* s_cur = h->lb_s.p + (size_t)cur_pos * N;
                double* y_cur = h->lb_y.p + (size_t)cur_pos * N;
                lb_reduce(h, N, y_cur, s_cur, RED_YS, 0, 0);
                lb_reduce(h, N, y_cur, y_cur, RED_FACTOR, 0, 0);
                LAUNCH(h, scale_dev_kernel, nb, 256, 0, N, sc, q);
            }
            for (int i = 0; i <= bound; ++i) {        // HLBFGS_UPDATE_Second_Step
                int st = (int)iter <= M ? i : (cur_pos + 1 + i) % M;
                lb_reduce(h, N, h->lb_y.p + (size_t)st * N, q, RED_BETA, i, st);
                LAUNCH(h, axpy_dev_kernel, nb, 256, 0, N, sc, h->lb_s.p + (size_t)st * N, q);
            }
            cur_pos = (cur_pos + 1) % M;
        }
        CUDA_CHECK(cudaMemcpyAsync(px, x, sizeof(double) * N, cudaMemcpyDeviceToDevice, h->stream));
        CUDA_CHECK(cudaMemcpyAsync(pg, g, sizeof(double) * N, cudaMemcpyDeviceToDevice, h->stream));
        if (iter == 0) {
            lb_reduce(h, N, g, g, RED_STP0, 0, 0);
            LAUNCH(h, set_info_kernel, 1, 1, 0, sc, 0, 0.0, 0);
        } else {
            LAUNCH(h, set_info_kernel, 1, 1, 0, sc, 0, 1.0, 1);
        }
        // MCSRCH: wa = x at the start of the line search (LineSearch.cpp:84)
        CUDA_CHECK(cudaMemcpyAsync(wa, x, sizeof(double) * N, cudaMemcpyDeviceToDevice, h->stream));
        for (;;) {
            lb_reduce(h, N, g, q, RED_STORE, 0, 0);           // g.s: dginit / dg
            LAUNCH(h, mcsrch_kernel, 1, 1, 0, sc, N);
            CUDA_CHECK(cudaMemcpyAsync(&ls_info, &sc->info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            sync_stream(h);
            if (ls_info != -1) break;
            LAUNCH(h, step_kernel, nb, 256, 0, N, sc, wa, q, x);
            newton_eval(h);
            nfev_total++;
        }
        lb_reduce(h, N, g, g, RED_GNORM, 0, 0);
        lb_reduce(h, N, x, x, RED_XNORM, 0, 0);
        iter++;
        CUDA_CHECK(cudaMemcpyAsync(&hs, sc, sizeof(hs), cudaMemcpyDeviceToHost, h->stream));
        sync_stream(h);
        if (cb && cb(user, iter, hs.f, hs.gnorm)) { canceled = true; break; }
        double xnorm = hs.xnorm < 1.0 ? 1.0 : hs.xnorm;
        if (ls_info != 1) break;                          // "Linesearch has failed"
        if (hs.gnorm / xnorm <= 0.0) break;               // PARAMETERS[5] = 0
        if (hs.gnorm < 0.0) break;                        // PARAMETERS[6] = epsg = 0
        if (hs.stp < stpmin || hs.stp > stpmax) break;
        if (iter > nb_iter) break;
    }
    if (info_out) { info_out[0] = iter; info_out[1] = nfev_total; info_out[2] = (u32)ls_info; info_out[3] = 0; }
    h->grid_valid = false; h->knn_valid = false; h->has_results = true;
    if (canceled) throw CanceledError("canceled by the progress callback");
}

extern "C" int b200cvt_newton_device(b200cvt_handle h, uint32_t nb_iter, uint32_t m, b200cvt_progress_cb cb, void* user,
                                     uint32_t* info_out) {
    return guarded([&] {
        if (!h) throw ArgError("null handle");
        if (!h->has_seeds) throw StateError("no seeds");
        CUDA_CHECK(cudaSetDevice(h->device));
        newton_loop(h, nb_iter, m, cb, user, info_out);
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
    });
}

extern "C" int b200cvt_newton(b200cvt_handle h, uint32_t nb_iter, uint32_t m, const uint8_t* locked, double* x_inout,
                              uint32_t S, b200cvt_progress_cb cb, void* user, uint32_t* info_out) {
    return guarded([&] {
        if (!h || !x_inout) throw ArgError("null argument");
        if (S == 0) throw ArgError("no seeds");
        CUDA_CHECK(cudaSetDevice(h->device));
        const u32 N = S * (u32)h->dim;
        h->x.ensure(N);
        CUDA_CHECK(cudaMemcpyAsync(h->x.p, x_inout, sizeof(double) * N, cudaMemcpyHostToDevice, h->stream));
        if (S != h->S && h->facet_guess.p) LAUNCH(h, fill_u32_kernel, 1024, 256, 0, h->facet_guess.p, (size_t)h->T, B200_NONE);
        set_seeds_common(h, S);
        upload_locked(h, locked, S);
        bool canceled = false;
        try { newton_loop(h, nb_iter, m, cb, user, info_out); } catch (const CanceledError&) { canceled = true; }
        CUDA_CHECK(cudaMemcpyAsync(x_inout, h->x.p, sizeof(double) * N, cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (canceled) throw CanceledError("canceled by the progress callback");
    });
}
