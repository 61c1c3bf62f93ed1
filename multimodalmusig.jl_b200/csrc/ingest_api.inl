// ingest_api.inl -- host side of the count ingest (included by mmsig_api.cu):
// format_counts_* (reference src/utils.jl:1-36) on the device, and set_data straight from a dense
// count matrix without a CSR round trip through the host.

struct DenseJob {                 // one modality's dense matrix on the device, counted and scanned
    void *dense = nullptr;        // D * V elements
    long long *rowptr = nullptr;  // D + 1
    int *flags = nullptr;         // [0] flag bits, [2..3] Σ counts
    long long nnz = 0, total = 0;
};
static void free_job(DenseJob &j) {
    cudaFree(j.dense);
    cudaFree(j.rowptr);
    cudaFree(j.flags);
    j = DenseJob();
}

// panel size (samples staged per block) for k_dense_panel, 0 if V is too large to stage
// Term-major panels take 128 samples (512-byte / 1 KB runs of int32 / int64 per term row).  Measured
// at D = 1e6, V = 96: int32 0.40 ms with 128 samples (4 blocks / SM), 0.66 ms with 256 (2 blocks / SM).
static int panel_samples(int V, int layout, int elem_bytes, size_t *smem_out) {
    const int VP = V | 1;
    (void)elem_bytes;
    const int first = layout == MMSIG_DENSE_TERM_MAJOR ? 128 : 64;
    for (int PS : {first, 128, 64, 32, 16}) {
        if (PS > first) continue;
        const size_t bytes = ((size_t)(PS * VP + 1) & ~(size_t)1) * sizeof(int) + (size_t)(PS + 1) * sizeof(long long);
        const size_t cap = PS >= 256 ? 112 * 1024 : (PS >= 128 ? 56 * 1024 : 50 * 1024);
        if (bytes <= cap && (size_t)PS * V <= 49152) { *smem_out = bytes; return PS; }
    }
    return 0;
}
template <typename T, bool FILL>
static void launch_panel(mmsig_handle *h, const T *dense, long long D, int V, int layout, int PS, size_t smem, long long *rowptr,
                         double *N, int n_stride, int n_off, int *flags, int2 *rec, int tag = 0) {
    auto kernel = k_dense_panel<T, FILL>;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, 256, smem);
    const long long npanels = (D + PS - 1) / PS;
    const int grid = (int)std::max<long long>(1, std::min<long long>(npanels, (long long)h->numSM * std::max(nb, 1)));
    kernel<<<grid, 256, smem, h->stream>>>(dense, D, V, layout, PS, rowptr, N, n_stride, n_off, flags,
                                           flags ? (unsigned long long *)(flags + 2) : nullptr, rec, tag);
}

static int launch_scan(mmsig_handle *h, long long *x, long long n) {
    const int nb = (int)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
    long long *bs = nullptr;
    CU(cudaMalloc(&bs, std::max(nb, 1) * sizeof(long long)));
    {
        LaunchScope ls(h, "k_scan");
        k_scan_local<<<std::max(nb, 1), 256, 0, h->stream>>>(x, n, bs);
        k_scan_sums<<<1, 1024, 0, h->stream>>>(bs, nb);
        k_scan_add<<<std::max(nb, 1), 256, 0, h->stream>>>(x, n, bs);
        h->launches += 2;
    }
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(bs);
    return 0;
}

// H2D of the dense matrix, row sizes + row totals, prefix sum; leaves job.nnz / job.total on the host
static int dense_count_scan(mmsig_handle *h, DenseJob &j, long long D, int V, const void *dense_host, int elem_bytes,
                            int layout, double *d_N, int n_stride, int n_off) {
    NEED(dense_host, "null dense matrix");
    NEED(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 (int32) or 8 (int64)");
    NEED(layout == MMSIG_DENSE_TERM_MAJOR || layout == MMSIG_DENSE_SAMPLE_MAJOR, "bad layout");
    const size_t bytes = (size_t)D * V * elem_bytes;
    if (cudaMalloc(&j.dense, bytes) != cudaSuccess || cudaMalloc(&j.rowptr, (D + 1) * sizeof(long long)) != cudaSuccess ||
        cudaMalloc(&j.flags, 4 * sizeof(int)) != cudaSuccess) {
        free_job(j);
        return fail(h, MMSIG_ENOMEM, "cudaMalloc (dense counts)");
    }
    CU(cudaMemcpyAsync(j.dense, dense_host, bytes, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemsetAsync(j.flags, 0, 4 * sizeof(int), h->stream));
    CU(cudaMemsetAsync(j.rowptr, 0, sizeof(long long), h->stream));
    {
        LaunchScope ls(h, "k_dense_count");
        size_t smem = 0;
        // term-major counting needs no staging: thread <-> sample reads 1 KB runs per term with V independent loads
        const int PS = layout == MMSIG_DENSE_TERM_MAJOR ? 0 : panel_samples(V, layout, elem_bytes, &smem);
        const long long units = layout == 0 ? (D + 255) / 256 : (D + 7) / 8;
        const int grid = (int)std::max<long long>(1, std::min<long long>(units, (long long)h->numSM * 8));
        unsigned long long *tot = (unsigned long long *)(j.flags + 2);
        if (elem_bytes == 4) {
            if (PS) launch_panel<int32_t, false>(h, (const int32_t *)j.dense, D, V, layout, PS, smem, j.rowptr, d_N, n_stride, n_off, j.flags, nullptr);
            else k_dense_count<int32_t><<<grid, 256, 0, h->stream>>>((const int32_t *)j.dense, D, V, layout, j.rowptr, d_N, n_stride, n_off, j.flags, tot);
        } else {
            if (PS) launch_panel<long long, false>(h, (const long long *)j.dense, D, V, layout, PS, smem, j.rowptr, d_N, n_stride, n_off, j.flags, nullptr);
            else k_dense_count<long long><<<grid, 256, 0, h->stream>>>((const long long *)j.dense, D, V, layout, j.rowptr, d_N, n_stride, n_off, j.flags, tot);
        }
    }
    int rc = launch_scan(h, j.rowptr, D);
    if (rc) return rc;
    int hf[4];
    CU(cudaMemcpyAsync(hf, j.flags, sizeof(hf), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(&j.nnz, j.rowptr + D, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    if (hf[0] & 8) return fail(h, MMSIG_ELIMIT, "a count exceeds 2^31-1");
    unsigned long long nt;
    memcpy(&nt, hf + 2, 8);
    j.total = (long long)nt;
    return 0;
}
template <typename T>
static void dense_fill_t(mmsig_handle *h, const T *dense, long long D, int V, int layout, const long long *rowptr, int2 *rec, int tag) {
    size_t smem = 0;
    const int PS = panel_samples(V, layout, (int)sizeof(T), &smem);
    if (PS) {
        launch_panel<T, true>(h, dense, D, V, layout, PS, smem, const_cast<long long *>(rowptr), nullptr, 0, 0, nullptr, rec, tag);
        return;
    }
    const long long units = layout == 0 ? (D + 255) / 256 : (D + 7) / 8;
    const int grid = (int)std::max<long long>(1, std::min<long long>(units, (long long)h->numSM * 8));
    k_dense_fill<T><<<grid, 256, 0, h->stream>>>(dense, D, V, layout, rowptr, rec, tag);
}
// tag != 0: records carry the sample's slot in its 32-sample tile (MMCTM's tile kernels)
static void dense_fill(mmsig_handle *h, const DenseJob &j, long long D, int V, int elem_bytes, int layout,
                       const long long *rowptr, int2 *rec, int tag = 0) {
    LaunchScope ls(h, "k_dense_fill");
    if (elem_bytes == 4) dense_fill_t<int32_t>(h, (const int32_t *)j.dense, D, V, layout, rowptr, rec, tag);
    else dense_fill_t<long long>(h, (const long long *)j.dense, D, V, layout, rowptr, rec, tag);
}

// ---- format_counts_lda / one modality of format_counts_mmctm -> CSR on the host ----------------
extern "C" int32_t mmsig_format_counts(mmsig_handle *h, int64_t D, int32_t V, const void *dense, int32_t elem_bytes,
                                       int32_t layout, int64_t *rowptr_out, int64_t *nnz_out) {
    NEED(h, "null handle");
    NEED(D >= 1 && V >= 1, "D, V must be >= 1");
    NEED(rowptr_out && nnz_out, "null output");
    CU(cudaSetDevice(h->device));
    cudaFree(h->fmt_rec);
    h->fmt_rec = nullptr;
    h->fmt_nnz = -1;
    DenseJob j;
    int rc = dense_count_scan(h, j, D, V, dense, elem_bytes, layout, nullptr, 0, 0);
    if (rc) { free_job(j); return rc; }
    if (cudaMalloc(&h->fmt_rec, std::max<long long>(j.nnz, 1) * sizeof(int2)) != cudaSuccess) {
        free_job(j);
        return fail(h, MMSIG_ENOMEM, "cudaMalloc (records)");
    }
    dense_fill(h, j, D, V, elem_bytes, layout, j.rowptr, h->fmt_rec);
    cudaError_t e = cudaMemcpyAsync(rowptr_out, j.rowptr, (D + 1) * sizeof(long long), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    const long long nnz = j.nnz;
    free_job(j);
    if (e != cudaSuccess) return fail(h, MMSIG_ECUDA, std::string("format_counts: ") + cudaGetErrorString(e));
    h->fmt_nnz = nnz;
    *nnz_out = nnz;
    return 0;
}

extern "C" int32_t mmsig_format_counts_fetch(mmsig_handle *h, int32_t *term_out, int32_t *count_out) {
    NEED(h, "null handle");
    NEED(h->fmt_nnz >= 0, "mmsig_format_counts first");
    NEED(h->fmt_nnz == 0 || (term_out && count_out), "null output");
    CU(cudaSetDevice(h->device));
    const long long n = h->fmt_nnz;
    if (n) {
        int *t = nullptr, *c = nullptr;
        if (cudaMalloc(&t, n * sizeof(int)) != cudaSuccess || cudaMalloc(&c, n * sizeof(int)) != cudaSuccess) {
            cudaFree(t);
            return fail(h, MMSIG_ENOMEM, "cudaMalloc (term / count)");
        }
        {
            LaunchScope ls(h, "k_unpack_rec");
            k_unpack_rec<<<(int)std::min<long long>((n + 255) / 256, (long long)h->numSM * 8), 256, 0, h->stream>>>(h->fmt_rec, n, t, c);
        }
        cudaError_t e = cudaMemcpyAsync(term_out, t, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(count_out, c, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cudaFree(t);
        cudaFree(c);
        if (e != cudaSuccess) return fail(h, MMSIG_ECUDA, std::string("format_counts_fetch: ") + cudaGetErrorString(e));
    }
    cudaFree(h->fmt_rec);
    h->fmt_rec = nullptr;
    h->fmt_nnz = -1;
    return 0;
}

// ---- MMCTM(K, α, V, format_counts_mmctm(dfs, cols)) without the host-side CSR -------------------
extern "C" int32_t mmsig_mmctm_set_data_dense(mmsig_handle *h, int64_t D, int64_t D_total, int32_t M, const int32_t *K,
                                              const int32_t *V, const void *const *dense, int32_t elem_bytes,
                                              int32_t layout) {
    NEED(h, "null handle");
    NEED(K && V && dense, "null argument");
    NEED(D >= 1, "D must be >= 1");
    if (M < 1 || M > MAXM) return fail(h, MMSIG_ELIMIT, "1 <= M <= 8 modalities supported");
    for (int m = 0; m < M; ++m) NEED(V[m] >= 1, "V[m] must be >= 1");
    CU(cudaSetDevice(h->device));
    DenseJob jobs[MAXM];
    double *tmpN = nullptr;
    auto cleanup = [&]() {
        for (int m = 0; m < M; ++m) free_job(jobs[m]);
        cudaFree(tmpN);
    };
    if (cudaMalloc(&tmpN, (size_t)D * M * sizeof(double)) != cudaSuccess) return fail(h, MMSIG_ENOMEM, "cudaMalloc (N)");
    int rc = 0;
    long long nnz[MAXM], ntot[MAXM];
    for (int m = 0; m < M && !rc; ++m) {
        rc = dense_count_scan(h, jobs[m], D, V[m], dense[m], elem_bytes, layout, tmpN, M, m);
        nnz[m] = jobs[m].nnz;
        ntot[m] = jobs[m].total;
    }
    bool same = false;
    if (!rc) rc = mmctm_prepare(h, D, D_total, M, K, V, nnz, &same);
    if (rc) { cleanup(); return rc; }
    MmctmHost &mm = h->mm;
    mm.has_data = false;
    cudaError_t e = cudaMemcpyAsync(const_cast<double *>(mm.p.N), tmpN, (size_t)D * M * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
    for (int m = 0; m < M && e == cudaSuccess; ++m) {
        e = cudaMemcpyAsync(mm.cb[m].rowptr, jobs[m].rowptr, (D + 1) * sizeof(long long), cudaMemcpyDeviceToDevice, h->stream);
        dense_fill(h, jobs[m], D, V[m], elem_bytes, layout, mm.cb[m].rowptr, mm.cb[m].rec, 1);
        densify_launch(h, mm.cb[m], 0, D);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
    if (e != cudaSuccess) return fail(h, MMSIG_ECUDA, std::string("set_data_dense: ") + cudaGetErrorString(e));
    if ((rc = allsum_ll(h, ntot, M))) return rc;
    for (int m = 0; m < M; ++m) mm.p.Ntot[m] = (double)ntot[m];
    mm.has_data = true;
    return 0;
}


// ---- count TSV files (data/*.tsv: header `term<TAB>sample...`, one line per term) ----------------
// Host-side reader for the numeric body, in the layout the ingest kernels take directly
// (MMSIG_DENSE_TERM_MAJOR).  The names (first column, header) stay with the caller's language.
static int tsv_scan(const char *path, long long *V, long long *D, std::vector<int32_t> *out, std::string &err) {
    FILE *f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return MMSIG_EINVAL; }
    std::vector<char> buf;
    {
        fseek(f, 0, SEEK_END);
        const long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        buf.resize((size_t)std::max<long>(n, 0) + 1);
        const size_t got = fread(buf.data(), 1, buf.size() - 1, f);
        buf[got] = 0;
        buf.resize(got + 1);
        fclose(f);
    }
    const char *p = buf.data(), *end = buf.data() + buf.size() - 1;
    // header: count the tabs
    long long ncol = 0;
    while (p < end && *p != '\n') { if (*p == '\t') ++ncol; ++p; }
    if (p < end) ++p;
    long long rows = 0;
    while (p < end) {
        if (*p == '\n' || *p == '\r') { ++p; continue; }                 // blank line
        while (p < end && *p != '\t' && *p != '\n') ++p;                  // the term name
        long long c = 0;
        while (p < end && *p == '\t') {
            ++p;
            bool neg = false;
            if (p < end && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
            if (p >= end || *p < '0' || *p > '9') { err = "non-integer count in row " + std::to_string(rows + 1); return MMSIG_EINVAL; }
            long long x = 0;
            while (p < end && *p >= '0' && *p <= '9') {
                x = x * 10 + (*p - '0');
                if (x > 2147483647LL) { err = "a count exceeds 2^31-1 in row " + std::to_string(rows + 1); return MMSIG_ELIMIT; }
                ++p;
            }
            if (out) {
                // the buffer was sized from an earlier scan of the same path: never write past it if the file changed
                const size_t at = (size_t)rows * ncol + c;
                if (c >= ncol || at >= out->size()) { err = "the file changed between the sizing scan and the read"; return MMSIG_EINVAL; }
                (*out)[at] = (int32_t)(neg ? -x : x);
            }
            ++c;
        }
        if (p < end && *p == '\r') ++p;
        if (p < end && *p != '\n') { err = "unexpected character in row " + std::to_string(rows + 1); return MMSIG_EINVAL; }
        if (c != ncol) { err = "row " + std::to_string(rows + 1) + " has " + std::to_string(c) + " counts, the header names " + std::to_string(ncol); return MMSIG_EINVAL; }
        ++rows;
    }
    *V = rows;
    *D = ncol;
    return 0;
}

extern "C" int32_t mmsig_tsv_dims(const char *path, int64_t *V, int64_t *D) {
    mmsig_handle *h = nullptr;
    NEED(path && V && D, "null argument");
    std::string err;
    long long v = 0, d = 0;
    const int rc = tsv_scan(path, &v, &d, nullptr, err);
    if (rc) return fail(h, rc, err);
    *V = v;
    *D = d;
    return 0;
}

extern "C" int32_t mmsig_tsv_read(const char *path, int64_t V, int64_t D, int32_t *dense_term_major) {
    mmsig_handle *h = nullptr;
    NEED(path && dense_term_major && V >= 0 && D >= 0, "bad argument");
    std::string err;
    long long v = 0, d = 0;
    int rc = tsv_scan(path, &v, &d, nullptr, err);
    if (rc) return fail(h, rc, err);
    NEED(v == V && d == D, "dimensions differ from mmsig_tsv_dims");
    std::vector<int32_t> tmp((size_t)V * D);
    if ((rc = tsv_scan(path, &v, &d, &tmp, err))) return fail(h, rc, err);
    memcpy(dense_term_major, tmp.data(), tmp.size() * sizeof(int32_t));
    return 0;
}
