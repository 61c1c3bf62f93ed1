// lda_api.inl -- host side of the LDA entry points (included by mmsig_api.cu)

#define LDA_TILE_DISPATCH(K, V, DENSE_RT, EXPR)                                                          \
    do {                                                                                                 \
        if (DENSE_RT) { constexpr bool DENSE = true; THETA_DISPATCH(K, TILE_DISPATCH_NW(V, EXPR)); }     \
        else { constexpr bool DENSE = false; THETA_DISPATCH(K, TILE_DISPATCH_NW(V, EXPR)); }             \
    } while (0)

template <typename F>
static int pick_lda_plan(mmsig_handle *h, F kernel, int KV /* V * (KP + 2) */, long long D, int *W_out, int *grid_out,
                         size_t *smem_out) {
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, kernel));
    CU(allow_max_smem(h, kernel));
    int bestW = 0, best_warps = 0, best_blocks = 0;
    size_t best_smem = 0;
    for (int W : {8, 4, 2, 1}) {
        size_t smem = (size_t)(1 + W) * KV * sizeof(double);
        if (smem + fa.sharedSizeBytes > h->smem_optin) continue;
        int nb = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, W * 32, smem));
        if (nb * W > best_warps) { best_warps = nb * W; bestW = W; best_blocks = nb; best_smem = smem; }
    }
    if (!bestW) return fail(h, MMSIG_ELIMIT, "K*V topic-term table does not fit in shared memory");
    *W_out = bestW;
    *smem_out = best_smem;
    long long want = (long long)h->numSM * best_blocks, cap = (D + bestW - 1) / bestW;
    *grid_out = (int)std::max<long long>(1, std::min(want, cap));
    return 0;
}

// job != nullptr: the counts are already on the device as a counted and scanned dense matrix
// (mmsig_lda_set_data_dense); else CSR from the host
static int lda_set_data_impl(mmsig_handle *h, int64_t D, int64_t D_total, int32_t K, int32_t V, const int64_t *rowptr,
                             const int32_t *term, const int32_t *count, const DenseJob *job, const double *job_N,
                             int elem_bytes, int layout) {
    NEED(h, "null handle");
    NEED(rowptr || job, "null rowptr");
    NEED(D >= 1 && D_total >= D, "need 1 <= D <= D_total");
    NEED(h->nranks > 1 || D_total == D, "D_total != D without mmsig_comm_init");
    NEED(K >= 1 && V >= 1, "K, V must be >= 1");
    if (K > 32) return fail(h, MMSIG_ELIMIT, "K <= 32 supported");
    if (V > 65535) return fail(h, MMSIG_ELIMIT, "V <= 65535 supported");
    CU(cudaSetDevice(h->device));
    // same shape as what is resident (a repeated fit! on the same corpus, e.g. every mmsig_lda_fit_host call of a
    // session): keep every allocation and launch plan, only the counts travel again
    {
        const long long nnz_in = job ? job->nnz : rowptr[D];
        LdaHost &L0 = h->lda;
        if (L0.has_data && !L0.p.factored && L0.p.K == K && L0.p.V == V && L0.p.D == D && L0.p.D_total == D_total && L0.cb.nnz == nnz_in) {
            L0.has_data = false;
            L0.has_state = false;
            L0.iterated = false;
            double *dN0 = const_cast<double *>(L0.p.N);
            long long ntot0 = 0;
            int rc0;
            if (job) {
                CU(cudaMemcpyAsync(dN0, job_N, (size_t)D * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
                CU(cudaMemcpyAsync(L0.cb.rowptr, job->rowptr, (D + 1) * sizeof(long long), cudaMemcpyDeviceToDevice, h->stream));
                dense_fill(h, *job, D, V, elem_bytes, layout, L0.cb.rowptr, L0.cb.rec, 1);
                densify_launch(h, L0.cb, 0, D);
                CU(cudaStreamSynchronize(h->stream));
                CU(cudaGetLastError());
                ntot0 = job->total;
            } else if ((rc0 = upload_counts(h, h->allocs_lda, L0.cb, D, V, 1, 0, rowptr, term, count, dN0, &ntot0, 1))) return rc0;
            if ((rc0 = allsum_ll(h, &ntot0, 1))) return rc0;
            L0.p.Ntot = (double)ntot0;
            L0.has_data = true;
            return 0;
        }
    }
    free_pool(h->allocs_lda);
    h->lda = LdaHost();
    LdaHost &L = h->lda;
    LdaDev &p = L.p;
    p.K = K;
    p.V = V;
    p.D = D;
    p.D_total = D_total;
    int rc;
    double *dN = nullptr;
    if ((rc = dev_alloc(h, h->allocs_lda, &dN, (size_t)D))) return rc;
    p.N = dN;
    long long ntot = 0;
    if (job) {
        if ((rc = ensure_countbuf(h, h->allocs_lda, L.cb, D, job->nnz, V))) return rc;
        CU(cudaMemcpyAsync(dN, job_N, (size_t)D * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CU(cudaMemcpyAsync(L.cb.rowptr, job->rowptr, (D + 1) * sizeof(long long), cudaMemcpyDeviceToDevice, h->stream));
        dense_fill(h, *job, D, V, elem_bytes, layout, L.cb.rowptr, L.cb.rec, 1);
        densify_launch(h, L.cb, 0, D);
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaGetLastError());
        ntot = job->total;
    } else if ((rc = upload_counts(h, h->allocs_lda, L.cb, D, V, 1, 0, rowptr, term, count, dN, &ntot, 1))) return rc;
    p.rowptr = L.cb.rowptr;
    p.rec = L.cb.rec;
    L.nnz = L.cb.nnz;
    if ((rc = allsum_ll(h, &ntot, 1))) return rc;
    p.Ntot = (double)ntot;
    const size_t KV = (size_t)K * V, DK = (size_t)D * K;
    for (double **t : {&p.lam, &p.Elnbeta, &p.Elnbeta_prev, &p.beta, &p.expElnbeta, &p.expElnbeta_prev})
        if ((rc = dev_alloc(h, h->allocs_lda, t, KV))) return rc;
    if ((rc = dev_alloc(h, h->allocs_lda, &L.gamA, DK))) return rc;
    if ((rc = dev_alloc(h, h->allocs_lda, &L.gamB, DK))) return rc;
    p.gamma = L.gamA;
    p.gamma_next = L.gamB;
    THETA_DISPATCH(K, rc = pick_lda_plan(h, k_lda_estep<KP, NP>, V * (KP + 2), D, &L.W, &L.grid, &L.smem));
    if (rc) return rc;
    L.grid_row = L.grid;
    // tiled (skinny-product) E pass when the tile fits: NW = ceil(V/32) warps, 32 NW samples per tile
    {
        const char *e = getenv("MMSIG_LDA");                 // "row": force the per-nonzero kernel (A/B)
        L.NW = (V + 31) / 32;
        L.tile = L.NW <= 8 && !(e && !strcmp(e, "row"));
        if (L.tile) {
            int KPv = 0;
            THETA_DISPATCH(K, KPv = KP);
            const int TS = 32 * L.NW, VP = V | 1;
            L.smem_tile = (size_t)(V * (KPv + 2) + TS * VP + TS * KPv) * sizeof(double);
            if (L.smem_tile > h->smem_optin) L.tile = false;
        }
        if (L.tile) {
            int nb = 0;
            THETA_DISPATCH(K, CU(allow_max_smem(h, k_lda_estep_tile<KP>)));
            THETA_DISPATCH(K, CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_lda_estep_tile<KP>, L.NW * 32, L.smem_tile)));
            const long long ntiles = (D + 32 * L.NW - 1) / (32 * L.NW);
            L.grid_tile = (int)std::max<long long>(1, std::min<long long>((long long)h->numSM * std::max(nb, 1), ntiles));
            L.grid = std::max(L.grid, L.grid_tile);          // the partial buffer serves both kernels
        }
    }
    // 32-sample tiles, one thread per term (lda_tile.cuh): the default when they fit
    {
        const char *e = getenv("MMSIG_LDA");                 // "tile96": the 32·NW-sample tile kernel (A/B)
        int KPv = 0;
        THETA_DISPATCH(K, KPv = KP);
        const int VP = V | 1, NW = (V + 31) / 32;
        const bool dn = L.cb.cnt != nullptr;
        {
            const size_t b1 = (size_t)V * KPv + (size_t)LDA_TS * VP + (size_t)LDA_TS * KPv + LDA_TS;
            const size_t b2 = (size_t)LDA_TS * VP + (size_t)LDA_TS * KPv + LDA_TS + (size_t)NW * 32;
            L.smem_t32 = (dn ? tile_stage_offset(b1) + tile_stage_doubles(V) : b1 + 4) * sizeof(double);
            L.smem_llt = (dn ? tile_stage_offset(b2) + tile_stage_doubles(V) : b2 + 4) * sizeof(double);
        }
        if (h->precision) {
            L.smem_t32 = f32_theta_smem(KPv, V);          // the LDA's FP32 E pass has the θ pass's layout
            L.smem_llt = f32_ll_smem(KPv, V);
            if (!(V <= 1024 && KPv <= 24 && L.smem_t32 <= h->smem_optin && dn))
                return fail(h, MMSIG_ELIMIT, "the FP32 mode of the LDA needs V <= 1024, K <= 24 and a tile that fits shared memory");
        }
        L.t32 = V <= 1024 && KPv <= 24 && L.smem_t32 <= h->smem_optin && (h->precision || !(e && (!strcmp(e, "row") || !strcmp(e, "tile96"))));
        if (L.t32) {
            int nb = 0, nb2 = 0;
            if (h->precision) {
                THETA_DISPATCH(K, TILE_DISPATCH_NW(V, CU(allow_max_smem(h, k_lda_estep_f32<KP, NWT>))));
                THETA_DISPATCH(K, TILE_DISPATCH_NW(V, CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_lda_estep_f32<KP, NWT>, NW * 32, L.smem_t32))));
                THETA_DISPATCH(K, TILE_DISPATCH_NW(V, CU(allow_max_smem(h, k_lda_ll_f32<KP, NWT>))));
                THETA_DISPATCH(K, TILE_DISPATCH_NW(V, CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, k_lda_ll_f32<KP, NWT>, NW * 32, L.smem_llt))));
            } else {
            LDA_TILE_DISPATCH(K, V, dn, CU(allow_max_smem(h, k_lda_estep_t32<KP, NWT, DENSE>)));
            LDA_TILE_DISPATCH(K, V, dn, CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_lda_estep_t32<KP, NWT, DENSE>, NW * 32, L.smem_t32)));
            LDA_TILE_DISPATCH(K, V, dn, CU(allow_max_smem(h, k_lda_ll_tile<KP, NWT, DENSE>)));
            LDA_TILE_DISPATCH(K, V, dn, CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, k_lda_ll_tile<KP, NWT, DENSE>, NW * 32, L.smem_llt)));
            }
            if (nb < 1 || nb2 < 1) L.t32 = false;
            else {
                const long long ntiles = (D + LDA_TS - 1) / LDA_TS;
                L.grid_t32 = (int)std::max<long long>(1, std::min<long long>((long long)h->numSM * nb, ntiles));
                L.grid_llt = (int)std::max<long long>(1, std::min<long long>((long long)h->numSM * nb2, ntiles));
                L.grid = std::max(L.grid, L.grid_t32);
            }
        }
        p.cnt = L.t32 ? L.cb.cnt : nullptr;
    }
    {
        int nb = 0;
        L.smem_ll = (KV + 256) * sizeof(double);
        L.smem_elbo = (2 * KV + 512) * sizeof(double);
        CU(allow_max_smem(h, k_lda_ll));
        CU(allow_max_smem(h, k_lda_elbo));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_lda_ll, 256, L.smem_ll));
        L.grid_ll = (int)std::max<long long>(1, std::min<long long>((long long)h->numSM * std::max(nb, 1), (D + 7) / 8));
    }
    if ((rc = dev_alloc(h, h->allocs_lda, &L.part, (size_t)L.grid * KV))) return rc;
    if ((rc = dev_alloc(h, h->allocs_lda, &L.part_ll, (size_t)std::max(L.grid_ll, L.grid_llt) * 8))) return rc;
    if ((rc = dev_alloc(h, h->allocs_lda, &L.rank_p, KV + 16))) return rc;
    if ((rc = dev_alloc(h, h->allocs_lda, &L.gath_p, (KV + 16) * h->nranks))) return rc;
    if ((rc = dev_alloc(h, h->allocs_lda, &L.rank_ll, (size_t)16))) return rc;
    if ((rc = dev_alloc(h, h->allocs_lda, &L.gath_ll, (size_t)16 * h->nranks))) return rc;
    if ((rc = dev_alloc(h, h->allocs_lda, &L.d_ll, (size_t)8))) return rc;
    L.has_data = true;
    return 0;
}

extern "C" int32_t mmsig_lda_set_data(mmsig_handle *h, int64_t D, int64_t D_total, int32_t K, int32_t V,
                                      const int64_t *rowptr, const int32_t *term, const int32_t *count) {
    return lda_set_data_impl(h, D, D_total, K, V, rowptr, term, count, nullptr, nullptr, 0, 0);
}

// LDA(K, α, η, format_counts_lda(df, cols)) without the host-side CSR (src/utils.jl:9-18)
extern "C" int32_t mmsig_lda_set_data_dense(mmsig_handle *h, int64_t D, int64_t D_total, int32_t K, int32_t V,
                                            const void *dense, int32_t elem_bytes, int32_t layout) {
    NEED(h, "null handle");
    NEED(D >= 1 && V >= 1, "D, V must be >= 1");
    CU(cudaSetDevice(h->device));
    DenseJob j;
    double *tmpN = nullptr;
    if (cudaMalloc(&tmpN, (size_t)D * sizeof(double)) != cudaSuccess) return fail(h, MMSIG_ENOMEM, "cudaMalloc (N)");
    int rc = dense_count_scan(h, j, D, V, dense, elem_bytes, layout, tmpN, 1, 0);
    if (!rc) rc = lda_set_data_impl(h, D, D_total, K, V, nullptr, nullptr, nullptr, &j, tmpN, elem_bytes, layout);
    free_job(j);
    cudaFree(tmpN);
    return rc;
}

extern "C" int32_t mmsig_lda_set_state(mmsig_handle *h, double alpha, double eta, const double *lambda,
                                       const double *gamma_next) {
    NEED(h, "null handle");
    LdaHost &L = h->lda;
    NEED(L.has_data, "mmsig_lda_set_data first");
    NEED(lambda, "lambda is required");
    NEED(alpha > 0 && eta > 0, "alpha, eta must be > 0");
    CU(cudaSetDevice(h->device));
    LdaDev &p = L.p;
    p.alpha = alpha;
    p.eta = eta;
    const size_t KV = (size_t)p.K * p.V, DK = (size_t)p.D * p.K;
    CU(cudaMemcpyAsync(p.lam, lambda, KV * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    {
        LaunchScope ls(h, "k_lda_elnbeta");
        k_lda_elnbeta<<<1, 1024, 0, h->stream>>>(p);
    }
    if (gamma_next) CU(cudaMemcpyAsync(p.gamma_next, gamma_next, DK * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    else {
        LaunchScope ls(h, "k_lda_gamma_init");
        k_lda_gamma_init<<<L.grid_ll, 256, 0, h->stream>>>(p);
    }
    fill(h, p.gamma, DK, 1.0);                       // model.γ = 1 (src/LDA.jl:41)
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    L.has_state = true;
    L.iterated = false;
    return 0;
}

static int lda_iterate_async(mmsig_handle *h, uint32_t flags) {
    LdaHost &L = h->lda;
    LdaDev &p = L.p;
    const int KV = p.K * p.V;
    const bool freeze = (flags & MMSIG_FLAG_FREEZE_TOPICS) != 0, unsm = (flags & MMSIG_FLAG_UNSMOOTHED) != 0;
    L.last_unsmoothed = unsm;
    L.last_frozen = freeze;
    std::swap(p.gamma, p.gamma_next);               // γ_t <- what the previous pass (or init) produced
    int nparts = L.grid_row;
    if (L.t32) {
        LaunchScope ls(h, "k_lda_estep_t32");
        const int nthr = 32 * ((p.V + 31) / 32);
        if (h->precision)
            THETA_DISPATCH(p.K, TILE_DISPATCH_NW(p.V, (k_lda_estep_f32<KP, NWT><<<L.grid_t32, nthr, L.smem_t32, h->stream>>>(
                                                          p, L.part, unsm ? p.beta : p.expElnbeta, !freeze))));
        else
        LDA_TILE_DISPATCH(p.K, p.V, p.cnt != nullptr, (k_lda_estep_t32<KP, NWT, DENSE><<<L.grid_t32, nthr, L.smem_t32, h->stream>>>(
                                                          p, L.part, unsm ? p.beta : p.expElnbeta, !freeze)));
        nparts = L.grid_t32;
    } else if (L.tile) {
        LaunchScope ls(h, "k_lda_estep_tile");
        THETA_DISPATCH(p.K, (k_lda_estep_tile<KP><<<L.grid_tile, L.NW * 32, L.smem_tile, h->stream>>>(
                                p, L.part, unsm ? p.beta : p.expElnbeta, !freeze, L.NW)));
        nparts = L.grid_tile;
    } else {
        LaunchScope ls(h, "k_lda_estep");
        THETA_DISPATCH(p.K, (k_lda_estep<KP, NP><<<L.grid_row, L.W * 32, L.smem, h->stream>>>(p, L.part, L.W,
                                                                                             unsm ? p.beta : p.expElnbeta, !freeze)));
    }
    const double2 *g = nullptr;
    int rc;
    if (!freeze) {
        {
            CombineSegs s{};
            s.nseg = 1;
            s.src[0] = L.part;
            s.nparts[0] = nparts;
            s.n[0] = KV;
            s.dst_off[0] = 0;
            LaunchScope ls(h, "k_combine");
            k_combine<<<(KV + 7) / 8, 256, 0, h->stream>>>(s, L.rank_p);
        }
        if ((rc = gather(h, L.rank_p, L.gath_p, KV, &g))) return rc;
        LaunchScope ls(h, "k_lda_mstep");
        if (p.factored) k_ilda_mstep<<<1, 1024, (size_t)2 * p.R * sizeof(double), h->stream>>>(p, g, h->nranks);
        else k_lda_mstep<<<1, 1024, 0, h->stream>>>(p, g, h->nranks);
    }
    if (L.t32) {
        LaunchScope ls(h, "k_lda_ll_tile");
        const int nthr = 32 * ((p.V + 31) / 32);
        if (h->precision)
            THETA_DISPATCH(p.K, TILE_DISPATCH_NW(p.V, (k_lda_ll_f32<KP, NWT><<<L.grid_llt, nthr, L.smem_llt, h->stream>>>(p, L.part_ll))));
        else
        LDA_TILE_DISPATCH(p.K, p.V, p.cnt != nullptr, (k_lda_ll_tile<KP, NWT, DENSE><<<L.grid_llt, nthr, L.smem_llt, h->stream>>>(p, L.part_ll)));
    } else {
        LaunchScope ls(h, "k_lda_ll");
        k_lda_ll<<<L.grid_ll, 256, L.smem_ll, h->stream>>>(p, L.part_ll);
    }
    {
        CombineSegs s{};
        s.nseg = 1;
        s.src[0] = L.part_ll;
        s.nparts[0] = L.t32 ? L.grid_llt : L.grid_ll;
        s.n[0] = 1;
        s.dst_off[0] = 0;
        LaunchScope ls(h, "k_combine");
        k_combine<<<1, 256, 0, h->stream>>>(s, L.rank_ll);
    }
    if ((rc = gather(h, L.rank_ll, L.gath_ll, 1, &g))) return rc;
    {
        LaunchScope ls(h, "k_lda_ll_final");
        k_lda_ll_final<<<1, 32, 0, h->stream>>>(g, h->nranks, p.Ntot, L.d_ll);
    }
    L.iterated = true;
    return 0;
}

extern "C" int32_t mmsig_lda_set_beta(mmsig_handle *h, const double *beta) {
    NEED(h && beta, "null argument");
    NEED(h->lda.has_state, "mmsig_lda_set_state first");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->lda.p.beta, beta, (size_t)h->lda.p.K * h->lda.p.V * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int32_t mmsig_lda_iterate(mmsig_handle *h, double *ll_out) { return mmsig_lda_iterate_flags(h, 0, ll_out); }

extern "C" int32_t mmsig_lda_iterate_flags(mmsig_handle *h, uint32_t flags, double *ll_out) {
    NEED(h, "null handle");
    NEED(h->lda.has_state, "mmsig_lda_set_state first");
    CU(cudaSetDevice(h->device));
    int rc = lda_iterate_async(h, flags);
    if (rc) return rc;
    double ll = 0.0;
    CU(cudaMemcpyAsync(&ll, h->lda.d_ll, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    if (ll_out) *ll_out = ll;
    return 0;
}

extern "C" int32_t mmsig_lda_fit(mmsig_handle *h, int32_t maxiter, double tol, double *ll_hist, int32_t *n_iter,
                                 int32_t *converged) {
    NEED(h, "null handle");
    NEED(h->lda.has_state, "mmsig_lda_set_state first");
    NEED(maxiter >= 1 && ll_hist, "maxiter >= 1 and ll_hist required");
    CU(cudaSetDevice(h->device));
    int it = 0, conv = 0, iter = 1;
    {
        // the convergence rule needs more than 10 log-likelihoods (src/LDA.jl:215): the iterations before that are
        // enqueued back to back, their log-likelihoods landing in page-locked memory, one synchronisation
        const int nfree = std::min<int>(maxiter, 10);
        if (nfree > 1) {
            for (int i = 0; i < nfree; ++i) {
                int rc = lda_iterate_async(h, 0);
                if (rc) return rc;
                CU(cudaMemcpyAsync(h->ll_pinned + i, h->lda.d_ll, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            }
            CU(cudaStreamSynchronize(h->stream));
            CU(cudaGetLastError());
            memcpy(ll_hist, h->ll_pinned, nfree * sizeof(double));
            iter = nfree + 1;
            it = nfree;
        }
    }
    for (; iter <= maxiter; ++iter) {
        int rc = mmsig_lda_iterate(h, &ll_hist[iter - 1]);
        if (rc) return rc;
        it = iter;
        if (iter > 10) {                                       // src/LDA.jl:215, src/common.jl:53-56
            double r = std::fabs(ll_hist[iter - 2] - ll_hist[iter - 1]) / std::fabs(ll_hist[iter - 1]);
            if (r < tol) { conv = 1; break; }
        }
    }
    if (n_iter) *n_iter = it;
    if (converged) *converged = conv;
    return 0;
}

extern "C" int32_t mmsig_lda_get_state(mmsig_handle *h, double *lambda, double *Elnbeta, double *beta, double *gamma,
                                       double *Elntheta, double *theta);

// fit!(model::LDA) from and to HOST buffers in one call (src/LDA.jl:198-224): set_data + set_state + fit + get_state.
// An LDA iteration (2.6 ms at a million samples) is far shorter than the upload of the counts, so nothing is gained by
// chunking the first E pass behind the copies as mmsig_mmctm_fit_host does; the call saves the host round trips.
extern "C" int32_t mmsig_lda_fit_host(mmsig_handle *h, int64_t D, int64_t D_total, int32_t K, int32_t V, const int64_t *rowptr,
                                      const int32_t *term, const int32_t *count, double alpha, double eta, const double *lambda,
                                      const double *gamma_next, int32_t maxiter, double tol, double *ll_hist, int32_t *n_iter,
                                      int32_t *converged, double *lambda_out, double *Elnbeta_out, double *beta_out,
                                      double *gamma_out, double *Elntheta_out, double *theta_out) {
    int rc = mmsig_lda_set_data(h, D, D_total, K, V, rowptr, term, count);
    if (rc) return rc;
    if ((rc = mmsig_lda_set_state(h, alpha, eta, lambda, gamma_next))) return rc;
    if ((rc = mmsig_lda_fit(h, maxiter, tol, ll_hist, n_iter, converged))) return rc;
    return mmsig_lda_get_state(h, lambda_out, Elnbeta_out, beta_out, gamma_out, Elntheta_out, theta_out);
}

static int lda_elbo_pass(mmsig_handle *h, double *phi_dev, double2 *host_parts /*[7]*/, double *tab /*[4]*/) {
    LdaHost &L = h->lda;
    LdaDev &p = L.p;
    const int nb = L.grid_ll;
    double2 *parts = nullptr;
    double *d_tab = nullptr;
    CU(cudaMalloc(&parts, (size_t)nb * 8 * sizeof(double2)));
    CU(cudaMalloc(&d_tab, 4 * sizeof(double)));
    CU(cudaMemsetAsync(parts, 0, (size_t)nb * 8 * sizeof(double2), h->stream));
    {
        LaunchScope ls(h, "k_lda_elbo");
        // ϕ_T as the last E pass computed it: from β (unsmoothed), the current e^{Elnβ} (topics frozen)
        // or the e^{Elnβ} that preceded the last M-step
        const double *Etab = L.last_unsmoothed ? p.beta : (L.last_frozen ? p.expElnbeta : p.expElnbeta_prev);
        k_lda_elbo<<<nb, 256, L.smem_elbo, h->stream>>>(p, parts, phi_dev, Etab);
    }
    {
        LaunchScope ls(h, "k_lda_elbo_tables");
        k_lda_elbo_tables<<<1, 256, 0, h->stream>>>(p, d_tab);
    }
    {
        CombineSegs s{};
        s.nseg = 1;
        s.src[0] = parts;
        s.nparts[0] = nb;
        s.n[0] = 8;
        s.dst_off[0] = 0;
        LaunchScope ls(h, "k_combine");
        k_combine<<<1, 256, 0, h->stream>>>(s, L.rank_ll);
    }
    const double2 *g = nullptr;
    int rc = gather(h, L.rank_ll, L.gath_ll, 8, &g);
    if (rc) return rc;
    std::vector<double2> hp((size_t)8 * h->nranks);
    CU(cudaMemcpyAsync(hp.data(), g, hp.size() * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(tab, d_tab, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(parts);
    cudaFree(d_tab);
    CU(cudaGetLastError());
    for (int i = 0; i < 7; ++i) {
        long double s = 0;
        for (int r = 0; r < h->nranks; ++r) s += (long double)hp[(size_t)r * 8 + i].x + (long double)hp[(size_t)r * 8 + i].y;
        host_parts[i] = make_double2((double)s, 0.0);
    }
    return 0;
}

extern "C" int32_t mmsig_lda_elbo(mmsig_handle *h, double *elbo, double *terms) {
    NEED(h, "null handle");
    LdaHost &L = h->lda;
    NEED(L.has_state && L.iterated, "mmsig_lda_elbo needs at least one iteration");
    CU(cudaSetDevice(h->device));
    LdaDev &p = L.p;
    double2 s[7];
    double tab[4];
    int rc = lda_elbo_pass(h, nullptr, s, tab);
    if (rc) return rc;
    const double K = p.K, V = p.V, Dt = (double)p.D_total;
    double t[7];
    t[0] = K * (std::lgamma(V * p.eta) - V * std::lgamma(p.eta)) + (p.eta - 1) * tab[0];          // :114-118
    double t4_factored = 0.0;
    if (p.factored) {
        // ILDA: the two table terms over the feature tables, on the host (a few hundred numbers).
        // ElnPβ src/ILDA.jl:131-140; ElnQβ :174-181 AS WRITTEN: `lnq = ...` inside the loop over the
        // features, so only the last feature contributes (reproduced; the ELBO is only reported).
        std::vector<double> lf(p.T), ef(p.T);
        CU(cudaMemcpyAsync(lf.data(), p.lambdaf, p.T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(ef.data(), p.Elnbetaf, p.T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        const int nf = p.nfeat, rowlen = p.T / p.K;
        int fo = 0;
        t[0] = 0.0;
        for (int f = 0; f < nf; ++f) {
            const int J = L.J_host[f];
            const double eta = L.etaf_host[f];
            double se = 0.0, a = 0.0, b = 0.0, c = 0.0;
            for (int k = 0; k < p.K; ++k) {
                double cs = 0.0;
                for (int j = 0; j < J; ++j) {
                    const double l = lf[(size_t)k * rowlen + fo + j], e = ef[(size_t)k * rowlen + fo + j];
                    se += e;
                    a += std::lgamma(l);
                    cs += l;
                    c += (l - 1) * e;
                }
                b += std::lgamma(cs);
            }
            t[0] += K * (std::lgamma(J * eta) - J * std::lgamma(eta)) + (eta - 1) * se;
            t4_factored = a - b - c;
            fo += J;
        }
    }
    t[1] = Dt * (std::lgamma(K * p.alpha) - K * std::lgamma(p.alpha)) + (p.alpha - 1) * s[0].x;   // :120-124
    t[2] = s[1].x;                                                                               // :126-132
    t[3] = s[2].x;                                                                               // :134-140
    t[4] = p.factored ? t4_factored : tab[1] - tab[2] - tab[3];                                  // :142-146
    t[5] = s[4].x - s[5].x - s[6].x;                                                             // :148-152
    t[6] = s[3].x;                                                                               // :154-160
    if (terms) memcpy(terms, t, sizeof(t));
    if (elbo) *elbo = t[0] + t[1] + t[2] + t[3] - t[4] - t[5] - t[6];
    return 0;
}

extern "C" int32_t mmsig_lda_get_state(mmsig_handle *h, double *lambda, double *Elnbeta, double *beta, double *gamma,
                                       double *Elntheta, double *theta) {
    NEED(h, "null handle");
    LdaHost &L = h->lda;
    NEED(L.has_state, "mmsig_lda_set_state first");
    CU(cudaSetDevice(h->device));
    LdaDev &p = L.p;
    const size_t KV = (size_t)p.K * p.V, DK = (size_t)p.D * p.K;
    auto d2h = [&](double *dst, const double *src, size_t n) -> cudaError_t {
        return dst ? cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream) : cudaSuccess;
    };
    CU(d2h(lambda, p.lam, KV));
    CU(d2h(Elnbeta, p.Elnbeta, KV));
    CU(d2h(beta, p.beta, KV));
    CU(d2h(gamma, p.gamma, DK));
    if (Elntheta || theta) {
        double *dE = nullptr, *dT = nullptr;
        CU(cudaMalloc(&dE, DK * sizeof(double)));
        CU(cudaMalloc(&dT, DK * sizeof(double)));
        {
            LaunchScope ls(h, "k_lda_theta_out");
            k_lda_theta_out<<<L.grid_ll, 256, 0, h->stream>>>(p, dE, dT);
        }
        CU(d2h(Elntheta, dE, DK));
        CU(d2h(theta, dT, DK));
        CU(cudaStreamSynchronize(h->stream));
        cudaFree(dE);
        cudaFree(dT);
    }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return 0;
}

extern "C" int32_t mmsig_lda_get_phi(mmsig_handle *h, double *phi_out) {
    NEED(h && phi_out, "null argument");
    LdaHost &L = h->lda;
    NEED(L.has_state && L.iterated, "mmsig_lda_get_phi needs at least one iteration");
    CU(cudaSetDevice(h->device));
    const size_t n = (size_t)L.nnz * L.p.K;
    double *d = nullptr;
    CU(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(double)));
    double2 s[7];
    double tab[4];
    int rc = lda_elbo_pass(h, d, s, tab);
    if (rc) { cudaFree(d); return rc; }
    CU(cudaMemcpyAsync(phi_out, d, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(d);
    return 0;
}


// ---- ILDA (reference src/ILDA.jl) on the LDA path: composite K x V tables for the per-sample kernels, the
// M-step over the feature tables (k_ilda_mstep).  Call order: mmsig_lda_set_data, mmsig_ilda_set_features,
// mmsig_ilda_set_state, then mmsig_lda_iterate / _fit / _elbo / _get_state, mmsig_ilda_get_tables.
extern "C" int32_t mmsig_ilda_set_features(mmsig_handle *h, int32_t nfeat, const int32_t *features) {
    NEED(h && features, "null argument");
    LdaHost &L = h->lda;
    NEED(L.has_data, "mmsig_lda_set_data first");
    NEED(!L.p.factored, "features are already set for this corpus");
    NEED(nfeat >= 1 && nfeat <= 16, "1 <= features <= 16");
    CU(cudaSetDevice(h->device));
    LdaDev &p = L.p;
    const int V = p.V, K = p.K;
    std::vector<int> J(nfeat, 0), ent_row, row_off, row_len, row_eta;
    for (int v = 0; v < V; ++v)
        for (int f = 0; f < nfeat; ++f) {
            const int x = features[(size_t)v * nfeat + f];
            NEED(x >= 0, "feature values are 0-based and non-negative");
            J[f] = std::max(J[f], x + 1);                       // J = maximum(features, dims=1), src/ILDA.jl:35
        }
    int t = 0;
    for (int k = 0; k < K; ++k)
        for (int f = 0; f < nfeat; ++f) {
            row_off.push_back(t);
            row_len.push_back(J[f]);
            row_eta.push_back(f);
            for (int j = 0; j < J[f]; ++j) ent_row.push_back((int)row_off.size() - 1);
            t += J[f];
        }
    p.T = t;
    p.R = (int)row_off.size();
    p.nfeat = nfeat;
    if ((size_t)2 * p.R * sizeof(double) > 48 * 1024) return fail(h, MMSIG_ELIMIT, "too many feature-table rows");
    int rc;
    auto up = [&](const int *src, size_t n, const int **dst) -> int {
        int *d = nullptr;
        int r = dev_alloc(h, h->allocs_lda, &d, n);
        if (r) return r;
        if (cudaMemcpyAsync(d, src, n * sizeof(int), cudaMemcpyHostToDevice, h->stream) != cudaSuccess)
            return fail(h, MMSIG_ECUDA, "cudaMemcpyAsync (feature index tables)");
        *dst = d;
        return 0;
    };
    if ((rc = up(features, (size_t)V * nfeat, &p.feat)) || (rc = up(ent_row.data(), ent_row.size(), &p.ent_row)) ||
        (rc = up(row_off.data(), row_off.size(), &p.row_off)) || (rc = up(row_len.data(), row_len.size(), &p.row_len)) ||
        (rc = up(row_eta.data(), row_eta.size(), &p.row_eta)))
        return rc;
    if ((rc = dev_alloc(h, h->allocs_lda, &p.lambdaf, (size_t)p.T))) return rc;
    if ((rc = dev_alloc(h, h->allocs_lda, &p.Elnbetaf, (size_t)p.T))) return rc;
    if ((rc = dev_alloc(h, h->allocs_lda, &p.etaf, (size_t)nfeat))) return rc;
    CU(cudaStreamSynchronize(h->stream));
    L.J_host = J;
    p.factored = 1;
    L.has_state = false;
    return 0;
}

// model.α, η ([i]), λ ([k][i][j] flat: model.λ[i][j, k]); gamma [d][k] or NULL => constructor state (src/ILDA.jl:38-52)
extern "C" int32_t mmsig_ilda_set_state(mmsig_handle *h, double alpha, const double *etaf, const double *lambdaf,
                                        const double *gamma_next) {
    NEED(h && etaf && lambdaf, "eta and lambda are required");
    LdaHost &L = h->lda;
    NEED(L.has_data && L.p.factored, "mmsig_lda_set_data and mmsig_ilda_set_features first");
    CU(cudaSetDevice(h->device));
    LdaDev &p = L.p;
    for (int f = 0; f < p.nfeat; ++f) NEED(etaf[f] > 0, "eta must be > 0");
    L.etaf_host.assign(etaf, etaf + p.nfeat);
    std::vector<double> l0((size_t)p.K * p.V, 1.0);            // placeholder K x V table: k_ilda_compose replaces what it seeds
    p.factored = 0;
    int rc = mmsig_lda_set_state(h, alpha, etaf[0], l0.data(), gamma_next);
    p.factored = 1;
    if (rc) return rc;
    L.has_state = false;
    CU(cudaMemcpyAsync(p.etaf, etaf, p.nfeat * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(p.lambdaf, lambdaf, p.T * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    {
        LaunchScope ls(h, "k_ilda_compose");
        k_ilda_compose<<<1, 1024, (size_t)2 * p.R * sizeof(double), h->stream>>>(p);
    }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    L.has_state = true;
    return 0;
}

extern "C" int32_t mmsig_ilda_get_tables(mmsig_handle *h, double *lambdaf, double *Elnbetaf) {
    NEED(h, "null handle");
    LdaHost &L = h->lda;
    NEED(L.has_state && L.p.factored, "mmsig_ilda_set_state first");
    CU(cudaSetDevice(h->device));
    LdaDev &p = L.p;
    if (lambdaf) CU(cudaMemcpyAsync(lambdaf, p.lambdaf, p.T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (Elnbetaf) CU(cudaMemcpyAsync(Elnbetaf, p.Elnbetaf, p.T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}
