// newton.inl — b200cvt_newton: CentroidalVoronoiTesselation::Newton_iterations
// (geogram/voronoi/CVT.cpp:272-338) = HLBFGS (geogram/third_party/HLBFGS/HLBFGS.cpp:281-587) over
// funcgrad = set_vertices + compute_CVT_func_grad(check_SR=true) + constrain_points.
// Included at the end of b200cvt.cu.

static u32 lb_blocks(u32 n) { return std::max<u32>(1u, std::min<u32>(div_up(n, 256), 148u * 8u)); }

// funcgrad (CVT.cpp:323-338) on the device: seeds = h->x, result g -> h->lb_g, per-seed energies -> newton_fs(h)
// (summed by lbfgs_post_eval_kernel).
// With partitioned seeds every rank evaluates its Morton slice, the (g, f_seed) slices are
// all-gathered, and every rank then holds the full gradient and the same energy.
// Sharded runs with the in-library communicator: L-BFGS vectors are sharded by contiguous ranges of ORIGINAL seed indices
// (rank r holds seeds [r L, (r+1) L), L = ceil(S / nranks)), independent of the Morton ranges the evaluation is sharded by
// (those are re-cut at every evaluation). The gradient travels from its evaluator to its L-BFGS owner with one
// reduce-scatter (every entry has exactly one non-zero contribution), the new positions come back with one in-place
// all-gather on the seed array; everything else of the optimiser is local + a few doubles through the peer mailboxes.
static u32 lb_local_seeds(b200cvt_ctx* h) {
    const u64 L = h->slice_len(), b = std::min<u64>((u64)h->rank * L, h->S);
    return (u32)(std::min<u64>(b + L, h->S) - b);
}

template <int D>
static void newton_scatter_one_gpu(b200cvt_ctx* h) {
    LAUNCH(h, scatter_results_kernel<D>, div_up(h->S, 256), 256, 0, (const SeedRec<D>*)h->xs.p, 0u, h->S, h->out_s.p, h->out_v.p,
           h->flags.p, h->pair_cnt.p, h->locked.p, 1, (double*)nullptr, h->lb_g.p, h->flags_orig.p, h->cnt_orig.p);
}

// the enlargement loop an evaluation left for later (evaluate_t, defer_redo), then the gradient again
static void newton_finish_deferred(b200cvt_ctx* h) {
    if (h->dim == 3) { surface_redo_loop<3>(h, h->pending_clip); newton_scatter_one_gpu<3>(h); }
    else { surface_redo_loop<6>(h, h->pending_clip); newton_scatter_one_gpu<6>(h); }
    CUDA_CHECK(cudaMemsetAsync(h->redo_n.p, 0, 4 * sizeof(u32), h->stream));
    h->redo_deferred = false;
}

template <int D>
static void newton_eval_t(b200cvt_ctx* h, bool defer) {
    h->grid_valid = false; h->knn_valid = false;
    h->redo_deferred = false;
    h->defer_redo = defer && h->nranks == 1 && !h->volumetric;
    evaluate(h, 1, 1);
    h->defer_redo = false;
    const u32 S = h->S;
    if (h->has_comm && h->nranks > 1 && h->pb_valid) {
        // peer-memory path: every gradient row is stored into its owner's slice, then a barrier
        const u32 nown = h->qend() - h->qbegin();
        if (nown > 0)
            LAUNCH(h, scatter_gradient_peer_kernel<D>, div_up(nown, 256), 256, 0, (const SeedRec<D>*)h->xs.p, h->qbegin(), h->qend(),
                   h->out_v.p, h->flags.p, h->pair_cnt.p, h->locked.p, h->pb, (u32)h->slice_len(), h->flags_orig.p, h->cnt_orig.p);
        LAUNCH(h, peer_barrier_kernel, 1, 32, 0, h->pc);
        h->exchanges++;
        return;
    }
    if (h->has_comm && h->nranks > 1) {
        const size_t L = h->slice_len();
        const size_t padded = L * h->nranks * D;
        h->g_full.ensure(padded);
        CUDA_CHECK(cudaMemsetAsync(h->g_full.p, 0, sizeof(double) * padded, h->stream));
        const u32 nown = h->qend() - h->qbegin();
        if (nown > 0)
            LAUNCH(h, scatter_results_kernel<D>, div_up(nown, 256), 256, 0, (const SeedRec<D>*)h->xs.p, h->qbegin(), h->qend(), h->out_s.p,
                   h->out_v.p, h->flags.p, h->pair_cnt.p, h->locked.p, 1, (double*)nullptr, h->g_full.p, h->flags_orig.p, h->cnt_orig.p);
        NCCL_CHECK(nccl_api().ReduceScatter(h->g_full.p, h->lb_g.p, L * D, ncclDouble, ncclSum, h->nccl, h->stream));
        h->exchanges++;
        return;
    }
    if (h->nranks == 1) {
        newton_scatter_one_gpu<D>(h);
    } else {
        if (!h->x_slice) throw StateError("seeds are partitioned but no exchange was set (b200cvt_set_exchange)");
        pack_slice<D>(h, h->out_s.p, h->out_v.p);
        run_exchange(h);
        h->s_orig.ensure(S);
        LAUNCH(h, unpack_all_kernel<D>, div_up(S, 256), 256, 0, (const SeedRec<D>*)h->xs.p, h->x_all, S, h->slice_len(),
               h->locked.p, 1, h->lb_g.p, h->s_orig.p);
    }
}
static void newton_eval(b200cvt_ctx* h, bool defer = false) {
    if (h->dim == 3) newton_eval_t<3>(h, defer); else newton_eval_t<6>(h, defer);
}

static const double* newton_fs(b200cvt_ctx* h) {
    if (h->has_comm && h->nranks > 1) return h->out_s.p + h->qbegin();     // this rank's evaluated seeds (sorted range)
    return h->nranks == 1 ? h->out_s.p : h->s_orig.p;
}

// the seed array with the padding the in-place all-gather needs (nranks * L rows), contents kept
static void newton_pad_seeds(b200cvt_ctx* h) {
    const size_t padded = (size_t)h->slice_len() * h->nranks * h->dim;
    h->x.grow_keep(padded, (size_t)h->S * h->dim, h->stream);
}

// the seed arrays and gradient slices of all ranks, mapped into this process (collective; called once per Newton call
// after the buffers have their final size)
static void newton_map_peer_buffers(b200cvt_ctx* h) {
    h->pb_valid = false;
    if (!h->use_peer_exchange) return;
    const u32 n = h->nranks;
    memset(&h->pb, 0, sizeof(h->pb));
    if (h->group) {
        // one process: peer access is enabled, the pointers are valid on every GPU
        h->group->barrier();
        for (u32 p = 0; p < n; ++p) { h->pb.x[p] = h->group->members[p]->x.p; h->pb.g[p] = h->group->members[p]->lb_g.p; }
        h->group->barrier();
        h->pb_valid = true;
        return;
    }
    // one process per GPU: CUDA IPC handles of the two allocations travel with one all-gather; a mapping is reused as long as
    // the peer's handle does not change
    cudaIpcMemHandle_t mine[2];
    CUDA_CHECK(cudaIpcGetMemHandle(&mine[0], h->x.p));
    CUDA_CHECK(cudaIpcGetMemHandle(&mine[1], h->lb_g.p));
    DevBuf<unsigned char> d_handles;
    d_handles.ensure(sizeof(mine) * n);
    CUDA_CHECK(cudaMemcpyAsync(d_handles.p + sizeof(mine) * h->rank, mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
    NCCL_CHECK(nccl_api().AllGather(d_handles.p + sizeof(mine) * h->rank, d_handles.p, sizeof(mine), ncclUint8, h->nccl, h->stream));
    std::vector<cudaIpcMemHandle_t> all(2 * (size_t)n);
    CUDA_CHECK(cudaMemcpyAsync(all.data(), d_handles.p, sizeof(mine) * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    for (u32 p = 0; p < n; ++p)
        for (int b = 0; b < 2; ++b) {
            double** slot = b == 0 ? &h->pb.x[p] : &h->pb.g[p];
            if (p == h->rank) { *slot = b == 0 ? h->x.p : h->lb_g.p; continue; }
            const cudaIpcMemHandle_t& hd = all[2 * (size_t)p + b];
            if (h->pb_ipc_ptr[b][p] && memcmp(&hd, &h->pb_ipc_handle[b][p], sizeof(hd)) != 0) {
                cudaIpcCloseMemHandle(h->pb_ipc_ptr[b][p]);
                h->pb_ipc_ptr[b][p] = nullptr;
            }
            if (!h->pb_ipc_ptr[b][p]) {
                void* ptr = nullptr;
                CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
                h->pb_ipc_ptr[b][p] = ptr; h->pb_ipc_handle[b][p] = hd;
            }
            *slot = (double*)h->pb_ipc_ptr[b][p];
        }
    h->pb_valid = true;
}

static void newton_gather_seeds(b200cvt_ctx* h) {
    const size_t L = h->slice_len();
    if (h->pb_valid) {
        // the new trial point of this rank's slice -> every rank's seed array, then a barrier
        const size_t n = (size_t)lb_local_seeds(h) * h->dim;
        if (n > 0)
            LAUNCH(h, peer_push_slice_kernel, std::min<u32>(div_up(n, 256), (u32)h->num_sms * 8u), 256, 0, h->pb, (int)h->rank, (int)h->nranks,
                   (size_t)h->rank * L * h->dim, n);
        LAUNCH(h, peer_barrier_kernel, 1, 32, 0, h->pc);
        h->exchanges++;
        return;
    }
    NCCL_CHECK(nccl_api().AllGather(h->x.p + (size_t)h->rank * L * h->dim, h->x.p, L * h->dim, ncclDouble, h->nccl, h->stream));
    h->exchanges++;
}

// HLBFGS main loop (HLBFGS.cpp:356-586) on the device-resident seeds h->x
static void newton_loop(b200cvt_ctx* h, u32 nb_iter, u32 m, b200cvt_progress_cb cb, void* user, uint32_t* info_out) {
    if (m > LBFGS_MAXM) throw ArgError("m too large");
    if (nb_iter < 1) return;                 // HLBFGS: INFO[4] < 1 -> "check your input parameters", no work
    h->rdt_valid = false; h->rdt_valid_mn = false;                    // the seeds move: a cached triangulation is stale
    const int D = h->dim;
    const u32 S = h->S;
    const u32 N_global = S * (u32)D;
    const bool sharded = h->has_comm && h->nranks > 1;
    // sharded: this rank's slice of every vector (a range of original seed indices); else the whole vectors
    const u32 N = sharded ? lb_local_seeds(h) * (u32)D : N_global;
    const u32 Nalloc = sharded ? h->slice_len() * (u32)D : N_global;
    const size_t xoff = sharded ? (size_t)h->rank * h->slice_len() * D : 0;
    if (sharded) newton_pad_seeds(h);
    const int M = (int)m;
    h->flags_orig.ensure(S); h->cnt_orig.ensure(S);
    h->lb_g.ensure(Nalloc); h->lb_q.ensure(Nalloc); h->lb_px.ensure(Nalloc); h->lb_pg.ensure(Nalloc); h->lb_wa.ensure(Nalloc);
    h->lb_s.ensure((size_t)std::max(M, 1) * Nalloc); h->lb_y.ensure((size_t)std::max(M, 1) * Nalloc);
    PeerComm pc = h->pc;
    if (!sharded) { memset(&pc, 0, sizeof(pc)); pc.nranks = 1; }
    if (sharded) newton_map_peer_buffers(h);
    // the direction kernel needs all its blocks resident (grid barriers): one cooperative launch
    if (h->lb_dir_blocks == 0) {
        int per_sm = 0, coop = 0;
        CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
        if (!coop) throw std::runtime_error("device does not support cooperative launches");
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbfgs_direction_kernel, LBFGS_DIR_THREADS, 0));
        { int per_sm2 = 0; CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, lbfgs_direction_gram_kernel, LBFGS_DIR_THREADS, 0)); per_sm = std::min(per_sm, per_sm2); }
        if (per_sm < 1) throw std::runtime_error("lbfgs_direction_kernel does not fit on an SM");
        h->lb_dir_blocks = (u32)h->num_sms * (u32)std::min(per_sm, LBFGS_DIR_BLOCKS_PER_SM);
    }
    h->lb_part.ensure(std::max<size_t>((LBFGS_GRAM_MAXK + 1) * (size_t)h->lb_dir_blocks + LBFGS_GRAM_MAXK, 4 * (size_t)LBFGS_POST_BLOCKS)); h->lb_sc.ensure(1);
    CUDA_CHECK(cudaMemsetAsync(h->lb_sc.p, 0, sizeof(LbfgsScalars), h->stream));
    double* x = h->x.p + xoff; double* g = h->lb_g.p; double* q = h->lb_q.p;
    LbfgsScalars* sc = h->lb_sc.p;
    const u32 nb = lb_blocks(N);
    const double stpmin = 1.0e-20, stpmax = 1.0e+20;

    u32 iter = 0, nfev_total = 0;
    int cur_pos = 0, ls_info = 0;
    bool canceled = false;
    struct { double f, dot, stp, gnorm, xnorm; int info, nfev; } hs;
    auto post_eval = [&](int resume) {
        const u32 nfs = sharded ? h->qend() - h->qbegin() : S;
        LAUNCH(h, lbfgs_post_eval_kernel, LBFGS_POST_BLOCKS, LBFGS_POST_THREADS, 0, nfs, newton_fs(h), N, g, q, x, h->lb_part.p, sc, resume,
               pc, N_global, h->redo_deferred ? (const u32*)h->redo_n.p : (const u32*)nullptr);
    };
    newton_eval(h); nfev_total++;
    post_eval(0);
    for (;;) {
        NvtxRange nvtx_it("b200cvt:L-BFGS iteration");
        LbfgsDirArgs da;
        memset(&da, 0, sizeof(da));
        da.N = N; da.first = (iter == 0) ? 1 : 0; da.M = M; da.cur_pos = cur_pos; da.bound = -1;
        if (iter > 0 && M > 0) {
            const int bound = (int)iter > M ? M - 1 : (int)iter - 1;
            da.bound = bound;
            for (int i = 0; i <= bound; ++i) {
                da.st1[i] = (int)iter <= M ? cur_pos - bound + i : (cur_pos - (bound - i) + M) % M;   // HLBFGS.cpp:160-176
                da.st2[i] = (int)iter <= M ? i : (cur_pos + 1 + i) % M;                              // :178-196
            }
        }
        da.x = x; da.g = g; da.q = q; da.px = h->lb_px.p; da.pg = h->lb_pg.p; da.wa = h->lb_wa.p;
        da.s = h->lb_s.p; da.y = h->lb_y.p; da.sc = sc; da.partials = h->lb_part.p;
        da.pc = pc; da.N_global = N_global;
        if (!h->evl[0]) for (int i = 0; i < 2; ++i) CUDA_CHECK(cudaEventCreate(&h->evl[i]));
        CUDA_CHECK(cudaEventRecord(h->evl[0], h->stream));
        {
            void* kargs[] = {(void*)&da};
            // with a history of at most LBFGS_GRAM_MAXM pairs: all dot products in one reduction (B200CVT_LBFGS_GRAM=0: level by level)
            const bool gram = h->use_lbfgs_gram && da.bound >= 0 && M <= LBFGS_GRAM_MAXM && h->lb_dir_blocks >= (u32)LBFGS_GRAM_MAXK;
            CUDA_CHECK(cudaLaunchCooperativeKernel(gram ? (const void*)lbfgs_direction_gram_kernel : (const void*)lbfgs_direction_kernel,
                                                   dim3(h->lb_dir_blocks), dim3(LBFGS_DIR_THREADS), kargs, 0, h->stream));
            h->launches++;
        }
        if (sharded) newton_gather_seeds(h);          // the first trial point (written by the direction kernel)
        CUDA_CHECK(cudaEventRecord(h->evl[1], h->stream));
        h->evl_used = true;
        if (iter > 0 && M > 0) cur_pos = (cur_pos + 1) % M;
        // MCSRCH (LineSearch.cpp:100-230): the direction kernel has moved x to the first trial point, unless the search
        // could not start (info != -1; rare: the evaluation below is then wasted and not counted)
        u32 nfev_ls = 0;
        for (;;) {
            newton_eval(h, true);
            nfev_ls++;
            post_eval(1);
            CUDA_CHECK(cudaMemcpyAsync(&hs, sc, sizeof(hs), cudaMemcpyDeviceToHost, h->stream));
            sync_stream(h);
            if (hs.info == -2) {
                // some seeds needed longer neighbour lists: finish the evaluation, then the line-search decision
                newton_finish_deferred(h);
                post_eval(1);
                CUDA_CHECK(cudaMemcpyAsync(&hs, sc, sizeof(hs), cudaMemcpyDeviceToHost, h->stream));
                sync_stream(h);
            }
            h->redo_deferred = false;
            ls_info = hs.info;
            if (ls_info != -1) break;
            LAUNCH(h, step_kernel, nb, 256, 0, N, sc, h->lb_wa.p, q, x);
            if (sharded) newton_gather_seeds(h);
        }
        if (ls_info == 0) nfev_ls = 0;                    // the search never started (a finished search reports 1..6)
        nfev_total += nfev_ls;
        iter++;
        if (agree_cancel(h, cb && cb(user, iter, hs.f, hs.gnorm))) { canceled = true; break; }
        double xnorm = hs.xnorm < 1.0 ? 1.0 : hs.xnorm;
        if (ls_info != 1) break;                          // "Linesearch has failed"
        if (hs.gnorm / xnorm <= 0.0) break;               // PARAMETERS[5] = 0
        if (hs.gnorm < 0.0) break;                        // PARAMETERS[6] = epsg = 0
        if (hs.stp < stpmin || hs.stp > stpmax) break;
        if (iter > nb_iter) break;
    }
    if (sharded) {
        unsigned int perr = 0;
        CUDA_CHECK(cudaMemcpyAsync(&perr, h->pc_err.p, sizeof(perr), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (perr) throw std::runtime_error("a peer GPU did not answer a mailbox reduction (timeout)");
    }
    if (info_out) { info_out[0] = iter; info_out[1] = nfev_total; info_out[2] = (u32)ls_info; info_out[3] = 0; }
    if (sharded && h->pb_valid) {
        // nobody may leave (and free or move its buffers) while a peer still stores into them
        LAUNCH(h, peer_barrier_kernel, 1, 32, 0, h->pc);
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        h->pb_valid = false;
    }
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
