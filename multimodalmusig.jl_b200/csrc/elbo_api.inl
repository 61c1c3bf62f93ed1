// elbo_api.inl -- host side of mmsig_mmctm_elbo (included by mmsig_api.cu)

__global__ void k_sum_qz(const double2 *parts, int nblocks, int nmod, double2 *out4) {
    // parts: [nmod][nblocks][4]; slot 3 of each -> out4[3]; slots 0..2 from modality 0's buffer
    if (threadIdx.x == 0) {
        double hi = 0.0, lo = 0.0;
        for (int m = 0; m < nmod; ++m)
            for (int b = 0; b < nblocks; ++b) {
                const double2 v = parts[((size_t)m * nblocks + b) * 4 + 3];
                dd_merge(hi, lo, v.x, v.y);
            }
        out4[3] = make_double2(hi, lo);
    }
    if (threadIdx.x >= 1 && threadIdx.x < 4) {
        const int s = threadIdx.x - 1;
        double hi = 0.0, lo = 0.0;
        for (int b = 0; b < nblocks; ++b) {
            const double2 v = parts[(size_t)b * 4 + s];
            dd_merge(hi, lo, v.x, v.y);
        }
        out4[s] = make_double2(hi, lo);
    }
}

static int mmctm_elbo_impl(mmsig_handle *h, double *elbo, double *terms) {
    MmctmHost &mm = h->mm;
    MmctmDev &p = mm.p;
    const int nb = mm.grid_post;
    // scratch lives with the plan (no cudaMalloc / cudaFree -- an implicit device synchronisation -- per call)
    double2 *parts = mm.part_elbo;                                       // [M][nb][4]
    double *d_tab = reinterpret_cast<double *>(mm.part_elbo + (size_t)p.M * nb * 4);
    CU(cudaMemsetAsync(parts, 0, (size_t)p.M * nb * 4 * sizeof(double2), h->stream));
    {
        LaunchScope ls(h, "k_elbo_tables");
        const size_t lus = (size_t)2 * p.MK * p.MK * sizeof(double) + p.MK * sizeof(int);
        CU(allow_max_smem(h, k_elbo_tables));
        k_elbo_tables<<<1, 256, lus, h->stream>>>(p, d_tab);
    }
    {
        LaunchScope ls(h, "k_elbo_samples");
        const size_t ess = (size_t)(p.MK * p.MK + 8 * p.MK) * sizeof(double);
        CU(allow_max_smem(h, k_elbo_samples));
        k_elbo_samples<<<nb, 256, ess, h->stream>>>(p, parts);
    }
    for (int m = 0; m < p.M; ++m) {
        const size_t smem = (size_t)p.K[m] * p.V[m] * sizeof(double);
        CU(allow_max_smem(h, k_elbo_qz));
        LaunchScope ls(h, "k_elbo_qz");
        k_elbo_qz<<<nb, 256, smem, h->stream>>>(p, m, parts + (size_t)m * nb * 4);
    }
    {
        LaunchScope ls(h, "k_sum_qz");
        k_sum_qz<<<1, 32, 0, h->stream>>>(parts, nb, p.M, mm.rank_p2);
    }
    const double2 *g = nullptr;
    int rc = gather(h, mm.rank_p2, mm.gath_p2, 4, &g);
    if (rc) return rc;
    std::vector<double2> hp((size_t)4 * h->nranks);
    double tab[4];
    CU(cudaMemcpyAsync(hp.data(), g, hp.size() * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(tab, d_tab, sizeof(tab), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    long double s[4] = {0, 0, 0, 0};
    for (int r = 0; r < h->nranks; ++r)
        for (int i = 0; i < 4; ++i) s[i] += (long double)hp[(size_t)r * 4 + i].x + (long double)hp[(size_t)r * 4 + i].y;
    const double Dt = (double)p.D_total, MK = (double)p.MK;
    const double log2pi = std::log(2.0 * M_PI);
    double t[7];
    t[0] = tab[0];
    t[1] = 0.5 * (Dt * (tab[3] - MK * log2pi) - (double)s[0]);
    t[2] = (double)s[1];
    t[3] = tab[2];
    t[4] = tab[1];
    t[5] = -0.5 * ((double)s[2] + Dt * MK * (log2pi + 1.0));
    t[6] = (double)s[3];
    if (terms) memcpy(terms, t, sizeof(t));
    if (elbo) *elbo = t[0] + t[1] + t[2] + t[3] - t[4] - t[5] - t[6];
    return 0;
}
