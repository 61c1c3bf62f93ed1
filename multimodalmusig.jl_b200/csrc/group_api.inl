// group_api.inl -- several GPUs from ONE process (include/mmsig.h "mmsig_group"); included by mmsig_api.cu.
//
// The reference's only parallelism is `addprocs` + `pmap` over independent restarts
// (scripts/run_mmctm.jl:8-11,99-111); `fit!` itself is serial (src/MMCTM.jl:463-465).  A group gives a single
// caller (the Julia shim's fit!(model; devices=0:7)) both: samples sharded over N devices for one fit, and
// restarts dealt over the devices with no communication.
//
// A group owns one mmsig_handle per device and runs every call on one host thread per device: each thread
// executes the ordinary single-handle entry point on its shard, and the places where ranks exchange data
// (gather, allsum_ll) take the group branch -- peer stores into the members' exchange arenas, an event, a
// host barrier, cross-stream event waits.  All members then reduce the gathered double-double partials in
// rank order, so every device holds bit-identical globals, exactly as the one-process-per-GPU path over NCCL.

static int group_gather(mmsig_handle *h, const double2 *rank_buf, size_t n, const double2 **out) {
    mmsig_group *g = h->grp;
    const int R = g->n, me = h->rank;
    // arena: two halves (call parity), each holding the members' buffers of THIS call packed in rank order
    // ([src rank][n]); (re)grown by every member at the same call (n is the same on all ranks)
    const size_t need = (size_t)2 * R * n;
    if (g->arena_cap[me] < need) {
        double2 *a = nullptr;
        const size_t cap = std::max<size_t>(need, (size_t)2 * R * 4096);
        if (cudaMalloc(&a, cap * sizeof(double2)) != cudaSuccess) {
            g->bar->abort();
            return fail(h, MMSIG_ENOMEM, "cudaMalloc (group exchange arena)");
        }
        if (g->arena[me]) h->old_arenas.push_back(g->arena[me]);     // peers may still be reading it: freed with the handle
        g->arena[me] = a;
        g->arena_cap[me] = cap;
        if (!g->bar->wait()) return fail(h, MMSIG_EINVAL, "another device of the group failed");
    }
    const int par = h->xparity;
    h->xparity ^= 1;
    PeerSlots dst;
    dst.n = R;
    for (int r = 0; r < R; ++r) dst.p[r] = g->arena[r] + (size_t)par * (g->arena_cap[r] / 2) + (size_t)me * n;
    {
        LaunchScope ls(h, "k_group_push");
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, 32));
        k_group_push<<<grid, 256, 0, h->stream>>>(rank_buf, n, dst);
    }
    cudaEvent_t mine = g->ev[(size_t)me * 2 + par];
    if (cudaEventRecord(mine, h->stream) != cudaSuccess) {
        g->bar->abort();
        return fail(h, MMSIG_ECUDA, "cudaEventRecord (group exchange)");
    }
    if (!g->bar->wait()) return fail(h, MMSIG_EINVAL, "another device of the group failed");      // every member has recorded
    for (int r = 0; r < R; ++r)
        if (r != me && cudaStreamWaitEvent(h->stream, g->ev[(size_t)r * 2 + par], 0) != cudaSuccess) {
            g->bar->abort();
            return fail(h, MMSIG_ECUDA, "cudaStreamWaitEvent (group exchange)");
        }
    double2 *base = g->arena[me] + (size_t)par * (g->arena_cap[me] / 2);
    *out = base;
    return 0;
}

static int group_allsum(mmsig_handle *h, long long *vals, int n) {
    mmsig_group *g = h->grp;
    for (int i = 0; i < n; ++i) g->scratch[(size_t)h->rank * MAXM + i] = vals[i];
    if (!g->bar->wait()) return fail(h, MMSIG_EINVAL, "another device of the group failed");
    for (int i = 0; i < n; ++i) {
        long long s = 0;
        for (int r = 0; r < g->n; ++r) s += g->scratch[(size_t)r * MAXM + i];
        vals[i] = s;
    }
    if (!g->bar->wait()) return fail(h, MMSIG_EINVAL, "another device of the group failed");      // scratch may be reused
    return 0;
}

static int gfail(mmsig_group *g, int code, const std::string &msg) {
    if (g) g->err = msg;
    g_last_error = msg;
    return code;
}

extern "C" const char *mmsig_group_last_error(const mmsig_group *g) { return g ? g->err.c_str() : g_last_error.c_str(); }
extern "C" int32_t mmsig_group_size(const mmsig_group *g) { return g ? g->n : 0; }
extern "C" mmsig_handle *mmsig_group_member(mmsig_group *g, int32_t i) { return (g && i >= 0 && i < g->n) ? g->h[i] : nullptr; }

extern "C" int32_t mmsig_group_destroy(mmsig_group *g) {
    if (!g) return 0;
    for (size_t i = 0; i < g->ev.size(); ++i)
        if (g->ev[i]) { cudaSetDevice(g->devices[i / 2]); cudaEventDestroy(g->ev[i]); }
    for (int r = 0; r < g->n; ++r) {
        if (!g->h[r]) continue;
        cudaSetDevice(g->devices[r]);
        cudaStreamSynchronize(g->h[r]->stream);
    }
    for (int r = 0; r < g->n; ++r) {
        if (!g->h[r]) continue;
        cudaSetDevice(g->devices[r]);
        for (void *p : g->h[r]->old_arenas) cudaFree(p);
        cudaFree(g->arena[r]);
        g->h[r]->grp = nullptr;
        mmsig_destroy(g->h[r]);
    }
    delete g->bar;
    delete g;
    return 0;
}

extern "C" int32_t mmsig_group_create(const mmsig_config *cfg, int32_t n, const int32_t *device_ids, mmsig_group **out) {
    if (!cfg || !out || !device_ids) return gfail(nullptr, MMSIG_EINVAL, "mmsig_group_create: null argument");
    if (n < 1 || n > 16) return gfail(nullptr, MMSIG_ELIMIT, "mmsig_group_create: 1 <= n_devices <= 16");
    mmsig_group *g = new mmsig_group();
    g->n = n;
    g->h.assign(n, nullptr);
    g->devices.assign(device_ids, device_ids + n);
    g->arena.assign(n, nullptr);
    g->arena_cap.assign(n, 0);
    g->ev.assign((size_t)2 * n, nullptr);
    g->scratch.assign((size_t)n * MAXM, 0);
    g->status.assign(n, 0);
    g->bar = new GroupBarrier();
    for (int r = 0; r < n; ++r) {
        mmsig_config c = *cfg;
        c.device = device_ids[r];
        int rc = mmsig_create(&c, &g->h[r]);
        if (rc) {
            const std::string msg = g_last_error;
            mmsig_group_destroy(g);
            return gfail(nullptr, rc, msg);
        }
        g->h[r]->grp = g;
        g->h[r]->rank = r;
        g->h[r]->nranks = n;
    }
    // peer access between every pair of distinct devices (the same device may be listed twice: two shards on
    // one GPU, which is how a one-GPU box tests the group path)
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            if (device_ids[a] == device_ids[b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, device_ids[a], device_ids[b]);
            if (!can) {
                mmsig_group_destroy(g);
                return gfail(nullptr, MMSIG_ENODEV, "devices of a group need peer access to one another (NVLink / NVSwitch)");
            }
            cudaSetDevice(device_ids[a]);
            cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[b], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) {
                mmsig_group_destroy(g);
                return gfail(nullptr, MMSIG_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            }
        }
    for (int r = 0; r < n; ++r) {
        cudaSetDevice(device_ids[r]);
        for (int par = 0; par < 2; ++par)
            if (cudaEventCreateWithFlags(&g->ev[(size_t)r * 2 + par], cudaEventDisableTiming) != cudaSuccess) {
                mmsig_group_destroy(g);
                return gfail(nullptr, MMSIG_ECUDA, "cudaEventCreate (group)");
            }
    }
    *out = g;
    return 0;
}

// run f(rank, handle) on one host thread per member; a failing member releases the others from any barrier
template <typename F>
static int group_run(mmsig_group *g, F f) {
    g->bar->reset(g->n);
    std::vector<std::thread> th;
    for (int r = 0; r < g->n; ++r)
        th.emplace_back([g, r, &f]() {
            cudaSetDevice(g->devices[r]);
            int rc = f(r, g->h[r]);
            g->status[r] = rc;
            if (rc) g->bar->abort();
        });
    for (auto &t : th) t.join();
    // report the first failure that is not the echo of another member's
    int first = 0, pick = -1;
    for (int r = 0; r < g->n && pick < 0; ++r)
        if (g->status[r] && g->h[r]->err.find("another device of the group failed") == std::string::npos) pick = r;
    for (int r = 0; r < g->n && pick < 0; ++r)
        if (g->status[r]) pick = r;
    if (pick >= 0) {
        first = g->status[pick];
        g->err = "device " + std::to_string(g->devices[pick]) + " (member " + std::to_string(pick) + "): " + g->h[pick]->err;
    }
    if (first) g_last_error = g->err;
    return first;
}

// contiguous shards of the rows, balanced by nonzeros (sum over modalities)
static int group_shard(mmsig_group *g, int64_t D, int32_t M, const int64_t *const *rowptr) {
    if (D < g->n) return gfail(g, MMSIG_EINVAL, "fewer samples than devices in the group");
    long long total = 0;
    for (int m = 0; m < M; ++m) {
        if (!rowptr[m] || rowptr[m][0] != 0) return gfail(g, MMSIG_EINVAL, "rowptr[0] must be 0");
        total += rowptr[m][D];
    }
    g->cut.assign(g->n + 1, 0);
    g->cut[g->n] = D;
    auto cum = [&](long long d) { long long s = 0; for (int m = 0; m < M; ++m) s += rowptr[m][d]; return s; };
    for (int r = 1; r < g->n; ++r) {
        const long long target = (long long)((double)total * r / g->n);
        long long lo = g->cut[r - 1] + 1, hi = D - (g->n - r);             // every shard keeps at least one row
        while (lo < hi) {
            const long long mid = (lo + hi) / 2;
            if (cum(mid) < target) lo = mid + 1; else hi = mid;
        }
        g->cut[r] = lo;
    }
    return 0;
}

// shard r of a CSR corpus: row pointers rebased to 0, term / count pointers advanced
struct ShardView {
    std::vector<std::vector<int64_t>> rp;
    std::vector<const int64_t *> rowptr;
    std::vector<const int32_t *> term, count;
    long long d0 = 0, d1 = 0;
};
static void make_shard(ShardView &sv, long long d0, long long d1, int32_t M, const int64_t *const *rowptr,
                       const int32_t *const *term, const int32_t *const *count) {
    sv.d0 = d0;
    sv.d1 = d1;
    sv.rp.assign(M, {});
    sv.rowptr.assign(M, nullptr);
    sv.term.assign(M, nullptr);
    sv.count.assign(M, nullptr);
    for (int m = 0; m < M; ++m) {
        const int64_t base = rowptr[m][d0];
        sv.rp[m].resize(d1 - d0 + 1);
        for (long long d = d0; d <= d1; ++d) sv.rp[m][d - d0] = rowptr[m][d] - base;
        sv.rowptr[m] = sv.rp[m].data();
        sv.term[m] = term[m] ? term[m] + base : nullptr;
        sv.count[m] = count[m] ? count[m] + base : nullptr;
    }
}

// a member running alone on a whole corpus (replica): no exchange with the other members
struct Detach {
    mmsig_handle *h;
    mmsig_group *g;
    int rank, nranks;
    explicit Detach(mmsig_handle *h_) : h(h_), g(h_->grp), rank(h_->rank), nranks(h_->nranks) {
        h->grp = nullptr;
        h->rank = 0;
        h->nranks = 1;
    }
    ~Detach() {
        h->grp = g;
        h->rank = rank;
        h->nranks = nranks;
    }
};

#define GNEED(cond, msg)                                  \
    do {                                                  \
        if (!(cond)) return gfail(g, MMSIG_EINVAL, msg);  \
    } while (0)

extern "C" int32_t mmsig_group_mmctm_set_data(mmsig_group *g, int64_t D, int32_t M, const int32_t *K, const int32_t *V,
                                              const int64_t *const *rowptr, const int32_t *const *term,
                                              const int32_t *const *count) {
    GNEED(g && K && V && rowptr && term && count, "null argument");
    if (M < 1 || M > MAXM) return gfail(g, MMSIG_ELIMIT, "1 <= M <= 8 modalities supported");
    int rc = group_shard(g, D, M, rowptr);
    if (rc) return rc;
    g->replica_best = -1;
    return group_run(g, [&](int r, mmsig_handle *h) {
        ShardView sv;
        make_shard(sv, g->cut[r], g->cut[r + 1], M, rowptr, term, count);
        return (int)mmsig_mmctm_set_data(h, sv.d1 - sv.d0, D, M, K, V, sv.rowptr.data(), sv.term.data(), sv.count.data());
    });
}

extern "C" int32_t mmsig_group_mmctm_set_state(mmsig_group *g, const double *alpha, const double *gamma, const double *lambda,
                                               const double *nu, const double *mu, const double *Sigma, const double *invSigma) {
    GNEED(g && alpha && gamma, "alpha and gamma are required");
    GNEED(g->cut.size() == (size_t)g->n + 1 && g->h[0]->mm.has_data, "mmsig_group_mmctm_set_data first");
    const int MK = g->h[0]->mm.p.MK;
    return group_run(g, [&](int r, mmsig_handle *h) {
        const size_t off = (size_t)g->cut[r] * MK;
        return (int)mmsig_mmctm_set_state(h, alpha, gamma, lambda ? lambda + off : nullptr, nu ? nu + off : nullptr, mu, Sigma, invSigma);
    });
}

extern "C" int32_t mmsig_group_mmctm_iterate(mmsig_group *g, uint32_t flags, double *ll_out) {
    GNEED(g, "null group");
    std::vector<double> ll((size_t)g->n * MAXM, 0.0);
    int rc = group_run(g, [&](int r, mmsig_handle *h) { return (int)mmsig_mmctm_iterate(h, flags, ll.data() + (size_t)r * MAXM); });
    if (!rc && ll_out) memcpy(ll_out, ll.data(), g->h[0]->mm.p.M * sizeof(double));
    return rc;
}

extern "C" int32_t mmsig_group_mmctm_fit(mmsig_group *g, int32_t maxiter, double tol, uint32_t flags, double *ll_hist,
                                         int32_t *n_iter, int32_t *converged) {
    GNEED(g && maxiter >= 1 && ll_hist, "maxiter >= 1 and ll_hist required");
    const int M = g->h[0]->mm.p.M;
    std::vector<std::vector<double>> hist(g->n, std::vector<double>((size_t)maxiter * std::max(M, 1)));
    std::vector<int> nit(g->n, 0), conv(g->n, 0);
    int rc = group_run(g, [&](int r, mmsig_handle *h) {
        // every member sees the same log-likelihoods, bit for bit, so all leave the loop at the same iteration
        return (int)mmsig_mmctm_fit(h, maxiter, tol, flags, r == 0 ? ll_hist : hist[r].data(), &nit[r], &conv[r]);
    });
    if (rc) return rc;
    if (n_iter) *n_iter = nit[0];
    if (converged) *converged = conv[0];
    return 0;
}

extern "C" int32_t mmsig_group_mmctm_elbo(mmsig_group *g, double *elbo, double *terms) {
    GNEED(g, "null group");
    if (g->replica_best >= 0) {
        Detach dt(g->h[g->replica_best]);
        int rc = mmsig_mmctm_elbo(g->h[g->replica_best], elbo, terms);
        if (rc) g->err = g->h[g->replica_best]->err;
        return rc;
    }
    std::vector<double> e(g->n, 0.0), t((size_t)g->n * 7, 0.0);
    int rc = group_run(g, [&](int r, mmsig_handle *h) { return (int)mmsig_mmctm_elbo(h, &e[r], t.data() + (size_t)r * 7); });
    if (rc) return rc;
    if (elbo) *elbo = e[0];
    if (terms) memcpy(terms, t.data(), 7 * sizeof(double));
    return 0;
}

extern "C" int32_t mmsig_group_mmctm_get_state(mmsig_group *g, double *lambda, double *nu, double *zeta, double *mu,
                                               double *Sigma, double *invSigma, double *gamma, double *Elnphi, double *phi,
                                               double *props) {
    GNEED(g, "null group");
    if (g->replica_best >= 0) {          // after mmsig_group_mmctm_restarts: the best restart lives on one device, whole
        Detach dt(g->h[g->replica_best]);
        int rc = mmsig_mmctm_get_state(g->h[g->replica_best], lambda, nu, zeta, mu, Sigma, invSigma, gamma, Elnphi, phi, props);
        if (rc) g->err = g->h[g->replica_best]->err;
        return rc;
    }
    GNEED(g->cut.size() == (size_t)g->n + 1, "mmsig_group_mmctm_set_data first");
    const int MK = g->h[0]->mm.p.MK, M = g->h[0]->mm.p.M;
    return group_run(g, [&](int r, mmsig_handle *h) {
        const size_t off = (size_t)g->cut[r] * MK, offz = (size_t)g->cut[r] * M;
        const bool z = r == 0;           // the tables are identical on every member: rank 0 writes them
        return (int)mmsig_mmctm_get_state(h, lambda ? lambda + off : nullptr, nu ? nu + off : nullptr, zeta ? zeta + offz : nullptr,
                                          z ? mu : nullptr, z ? Sigma : nullptr, z ? invSigma : nullptr, z ? gamma : nullptr,
                                          z ? Elnphi : nullptr, z ? phi : nullptr, props ? props + off : nullptr);
    });
}

extern "C" int32_t mmsig_group_mmctm_get_evals(mmsig_group *g, int32_t *nev_nu, int32_t *nev_lambda) {
    GNEED(g && g->cut.size() == (size_t)g->n + 1, "mmsig_group_mmctm_set_data first");
    return group_run(g, [&](int r, mmsig_handle *h) {
        return (int)mmsig_mmctm_get_evals(h, nev_nu ? nev_nu + g->cut[r] : nullptr, nev_lambda ? nev_lambda + g->cut[r] : nullptr);
    });
}

// model.θ[d][m] of modality m for all samples (nnz_m x K_m, [w][k]): every member recomputes its shard's rows
extern "C" int32_t mmsig_group_mmctm_get_theta(mmsig_group *g, int32_t m, double *theta_out) {
    GNEED(g && theta_out, "null argument");
    if (g->replica_best >= 0) {
        Detach dt(g->h[g->replica_best]);
        return mmsig_mmctm_get_theta(g->h[g->replica_best], m, theta_out);
    }
    GNEED(g->cut.size() == (size_t)g->n + 1 && g->h[0]->mm.has_state, "mmsig_group_mmctm_set_data / _set_state first");
    GNEED(m >= 0 && m < g->h[0]->mm.p.M, "bad modality");
    std::vector<size_t> off(g->n + 1, 0);
    for (int r = 0; r < g->n; ++r) off[r + 1] = off[r] + (size_t)g->h[r]->mm.nnz[m] * g->h[r]->mm.p.K[m];
    return group_run(g, [&](int r, mmsig_handle *h) { return (int)mmsig_mmctm_get_theta(h, m, theta_out + off[r]); });
}

// fit! from and to host buffers over all devices of the group: mmsig_mmctm_fit_host per shard, each device
// pipelining its own uploads behind its own E-step
static int group_fit_host_impl(mmsig_group *g, int64_t D, int32_t M, const int32_t *K, const int32_t *V,
                                              const int64_t *const *rowptr, const int32_t *const *term,
                                              const int32_t *const *count, const double *alpha, const double *gamma,
                                              const double *lambda, const double *nu, const double *mu, const double *Sigma,
                                              const double *invSigma, int32_t maxiter, double tol, uint32_t flags,
                                              double *ll_hist, int32_t *n_iter, int32_t *converged, double *lambda_out,
                                              double *nu_out, double *zeta_out, double *mu_out, double *Sigma_out,
                                              double *invSigma_out, double *gamma_out, double *Elnphi_out, double *phi_out,
                                              double *props_out, bool packed) {
    GNEED(g && K && V && rowptr && term && (count || packed) && alpha && gamma, "null argument");
    GNEED(maxiter >= 1 && ll_hist, "maxiter >= 1 and ll_hist required");
    if (M < 1 || M > MAXM) return gfail(g, MMSIG_ELIMIT, "1 <= M <= 8 modalities supported");
    int rc = group_shard(g, D, M, rowptr);
    if (rc) return rc;
    g->replica_best = -1;
    int MK = 0;
    for (int m = 0; m < M; ++m) MK += K[m];
    std::vector<std::vector<double>> hist(g->n, std::vector<double>((size_t)maxiter * M));
    std::vector<int> nit(g->n, 0), conv(g->n, 0);
    rc = group_run(g, [&](int r, mmsig_handle *h) {
        ShardView sv;
        std::vector<const int32_t *> nocount(M, nullptr);
        make_shard(sv, g->cut[r], g->cut[r + 1], M, rowptr, term, packed ? nocount.data() : count);
        const size_t off = (size_t)sv.d0 * MK, offz = (size_t)sv.d0 * M;
        const bool z = r == 0;
        return (int)mmctm_fit_host_impl(h, sv.d1 - sv.d0, D, M, K, V, sv.rowptr.data(), sv.term.data(), packed ? nullptr : sv.count.data(), alpha, gamma,
                                         lambda ? lambda + off : nullptr, nu ? nu + off : nullptr, mu, Sigma, invSigma, maxiter, tol, flags,
                                         z ? ll_hist : hist[r].data(), &nit[r], &conv[r], lambda_out ? lambda_out + off : nullptr,
                                         nu_out ? nu_out + off : nullptr, zeta_out ? zeta_out + offz : nullptr, z ? mu_out : nullptr,
                                         z ? Sigma_out : nullptr, z ? invSigma_out : nullptr, z ? gamma_out : nullptr,
                                         z ? Elnphi_out : nullptr, z ? phi_out : nullptr, props_out ? props_out + off : nullptr, packed);
    });
    if (rc) return rc;
    if (n_iter) *n_iter = nit[0];
    if (converged) *converged = conv[0];
    return 0;
}

extern "C" int32_t mmsig_group_mmctm_fit_host(mmsig_group *g, int64_t D, int32_t M, const int32_t *K, const int32_t *V,
                                              const int64_t *const *rowptr, const int32_t *const *term,
                                              const int32_t *const *count, const double *alpha, const double *gamma,
                                              const double *lambda, const double *nu, const double *mu, const double *Sigma,
                                              const double *invSigma, int32_t maxiter, double tol, uint32_t flags,
                                              double *ll_hist, int32_t *n_iter, int32_t *converged, double *lambda_out,
                                              double *nu_out, double *zeta_out, double *mu_out, double *Sigma_out,
                                              double *invSigma_out, double *gamma_out, double *Elnphi_out, double *phi_out,
                                              double *props_out) {
    return group_fit_host_impl(g, D, M, K, V, rowptr, term, count, alpha, gamma, lambda, nu, mu, Sigma, invSigma, maxiter, tol, flags,
                               ll_hist, n_iter, converged, lambda_out, nu_out, zeta_out, mu_out, Sigma_out, invSigma_out, gamma_out,
                               Elnphi_out, phi_out, props_out, false);
}
extern "C" int32_t mmsig_group_mmctm_fit_host_packed(mmsig_group *g, int64_t D, int32_t M, const int32_t *K, const int32_t *V,
                                                     const int64_t *const *rowptr, const uint32_t *const *rec, const double *alpha,
                                                     const double *gamma, const double *lambda, const double *nu, const double *mu,
                                                     const double *Sigma, const double *invSigma, int32_t maxiter, double tol,
                                                     uint32_t flags, double *ll_hist, int32_t *n_iter, int32_t *converged,
                                                     double *lambda_out, double *nu_out, double *zeta_out, double *mu_out,
                                                     double *Sigma_out, double *invSigma_out, double *gamma_out, double *Elnphi_out,
                                                     double *phi_out, double *props_out) {
    return group_fit_host_impl(g, D, M, K, V, rowptr, reinterpret_cast<const int32_t *const *>(rec), nullptr, alpha, gamma, lambda, nu, mu,
                               Sigma, invSigma, maxiter, tol, flags, ll_hist, n_iter, converged, lambda_out, nu_out, zeta_out, mu_out,
                               Sigma_out, invSigma_out, gamma_out, Elnphi_out, phi_out, props_out, true);
}

// Independent restarts dealt over the devices (scripts/run_mmctm.jl:99-111 `pmap(fit_restart, ...)`; README.md:42):
// every device holds the WHOLE corpus and fits restarts r = rank, rank + n, ... with no communication; the
// arg-max of the ELBOs is taken on the host (first restart wins ties).  Afterwards the group's get_state / elbo
// read the device that holds the best restart.
extern "C" int32_t mmsig_group_mmctm_restarts(mmsig_group *g, int64_t D, int32_t M, const int32_t *K, const int32_t *V,
                                              const int64_t *const *rowptr, const int32_t *const *term,
                                              const int32_t *const *count, const double *alpha, int32_t R, const double *gamma0,
                                              int32_t maxiter, double tol, uint32_t flags, double *elbo_out, double *ll_out,
                                              int32_t *n_iter_out, int32_t *best) {
    GNEED(g && K && V && rowptr && term && count && alpha && gamma0, "null argument");
    GNEED(R >= 1 && maxiter >= 1, "R >= 1 and maxiter >= 1 required");
    if (M < 1 || M > MAXM) return gfail(g, MMSIG_ELIMIT, "1 <= M <= 8 modalities supported");
    size_t G = 0;
    for (int m = 0; m < M; ++m) G += (size_t)K[m] * V[m];
    std::vector<double> e(R, 0.0), ll((size_t)R * M, 0.0);
    std::vector<int> nit(R, 0), lbest(g->n, -1);
    g->cut.clear();
    int rc = group_run(g, [&](int r, mmsig_handle *h) {
        std::vector<int> mine;
        for (int i = r; i < R; i += g->n) mine.push_back(i);
        if (mine.empty()) return 0;
        Detach dt(h);                        // a replica: this member runs alone on the whole corpus
        int rcl = mmsig_mmctm_set_data(h, D, D, M, K, V, rowptr, term, count);
        if (!rcl) rcl = mmsig_mmctm_set_state(h, alpha, gamma0 + (size_t)mine[0] * G, nullptr, nullptr, nullptr, nullptr, nullptr);
        if (!rcl) {
            std::vector<double> g0(mine.size() * G), el(mine.size()), l2(mine.size() * M);
            std::vector<int> ni(mine.size());
            for (size_t i = 0; i < mine.size(); ++i) memcpy(g0.data() + i * G, gamma0 + (size_t)mine[i] * G, G * sizeof(double));
            int b = -1;
            rcl = mmsig_mmctm_restarts(h, (int)mine.size(), g0.data(), maxiter, tol, flags, el.data(), l2.data(), ni.data(), &b);
            if (!rcl) {
                for (size_t i = 0; i < mine.size(); ++i) {
                    e[mine[i]] = el[i];
                    nit[mine[i]] = ni[i];
                    memcpy(ll.data() + (size_t)mine[i] * M, l2.data() + i * M, M * sizeof(double));
                }
                lbest[r] = mine[b];
            }
        }
        return rcl;
    });
    if (rc) return rc;
    int bi = 0;
    for (int i = 1; i < R; ++i)
        if (e[i] > e[bi] || e[bi] != e[bi]) bi = i;
    g->replica_best = bi % g->n;
    // the owning device holds ITS best restart; that is the global best by construction (bi is among its restarts
    // and no restart of that device has a larger ELBO)
    if (elbo_out) memcpy(elbo_out, e.data(), R * sizeof(double));
    if (ll_out) memcpy(ll_out, ll.data(), (size_t)R * M * sizeof(double));
    if (n_iter_out) memcpy(n_iter_out, nit.data(), R * sizeof(int));
    if (best) *best = bi;
    return 0;
}

// ---- LDA over a group: samples sharded as above (src/LDA.jl:198-224) ---------------------------------------
extern "C" int32_t mmsig_group_lda_set_data(mmsig_group *g, int64_t D, int32_t K, int32_t V, const int64_t *rowptr,
                                            const int32_t *term, const int32_t *count) {
    GNEED(g && rowptr && term && count, "null argument");
    const int64_t *rps[1] = {rowptr};
    const int32_t *ts[1] = {term}, *cs[1] = {count};
    int rc = group_shard(g, D, 1, rps);
    if (rc) return rc;
    return group_run(g, [&](int r, mmsig_handle *h) {
        ShardView sv;
        make_shard(sv, g->cut[r], g->cut[r + 1], 1, rps, ts, cs);
        return (int)mmsig_lda_set_data(h, sv.d1 - sv.d0, D, K, V, sv.rowptr[0], sv.term[0], sv.count[0]);
    });
}
extern "C" int32_t mmsig_group_lda_set_state(mmsig_group *g, double alpha, double eta, const double *lambda,
                                             const double *gamma_next) {
    GNEED(g && lambda && g->cut.size() == (size_t)g->n + 1 && g->h[0]->lda.has_data, "mmsig_group_lda_set_data first; lambda required");
    const int K = g->h[0]->lda.p.K;
    return group_run(g, [&](int r, mmsig_handle *h) {
        return (int)mmsig_lda_set_state(h, alpha, eta, lambda, gamma_next ? gamma_next + (size_t)g->cut[r] * K : nullptr);
    });
}
extern "C" int32_t mmsig_group_lda_fit(mmsig_group *g, int32_t maxiter, double tol, double *ll_hist, int32_t *n_iter,
                                       int32_t *converged) {
    GNEED(g && maxiter >= 1 && ll_hist, "maxiter >= 1 and ll_hist required");
    std::vector<std::vector<double>> hist(g->n, std::vector<double>((size_t)maxiter));
    std::vector<int> nit(g->n, 0), conv(g->n, 0);
    int rc = group_run(g, [&](int r, mmsig_handle *h) {
        return (int)mmsig_lda_fit(h, maxiter, tol, r == 0 ? ll_hist : hist[r].data(), &nit[r], &conv[r]);
    });
    if (rc) return rc;
    if (n_iter) *n_iter = nit[0];
    if (converged) *converged = conv[0];
    return 0;
}
extern "C" int32_t mmsig_group_lda_elbo(mmsig_group *g, double *elbo, double *terms) {
    GNEED(g, "null group");
    std::vector<double> e(g->n, 0.0), t((size_t)g->n * 7, 0.0);
    int rc = group_run(g, [&](int r, mmsig_handle *h) { return (int)mmsig_lda_elbo(h, &e[r], t.data() + (size_t)r * 7); });
    if (rc) return rc;
    if (elbo) *elbo = e[0];
    if (terms) memcpy(terms, t.data(), 7 * sizeof(double));
    return 0;
}
extern "C" int32_t mmsig_group_lda_get_state(mmsig_group *g, double *lambda, double *Elnbeta, double *beta, double *gamma,
                                             double *Elntheta, double *theta) {
    GNEED(g && g->cut.size() == (size_t)g->n + 1, "mmsig_group_lda_set_data first");
    const int K = g->h[0]->lda.p.K;
    return group_run(g, [&](int r, mmsig_handle *h) {
        const size_t off = (size_t)g->cut[r] * K;
        const bool z = r == 0;
        return (int)mmsig_lda_get_state(h, z ? lambda : nullptr, z ? Elnbeta : nullptr, z ? beta : nullptr, gamma ? gamma + off : nullptr,
                                        Elntheta ? Elntheta + off : nullptr, theta ? theta + off : nullptr);
    });
}
