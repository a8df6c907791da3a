// neigh_host.inl -- host side of the explicit neighbour list (included by amofb.cu)

struct NeighState {
    Batcher bt;
    int n_species = 0, nkeys = 0;
    double r2search = 0.0;
    double *d_cnthr2 = nullptr;
    uint16_t *d_keyidx = nullptr;
    int *d_count = nullptr;
    long long *d_offset = nullptr;
    BatchSlot *slot = nullptr;
    long long total = 0;
};

static void neigh_release(amofb_ctx *ctx) {
    NeighState *p = ctx->neigh;
    if (!p) return;
    cudaStreamSynchronize(ctx->s_copy);
    cudaStreamSynchronize(ctx->s_compute);
    batcher_release(ctx, p->bt);
    pool_put(ctx, p->d_cnthr2); pool_put(ctx, p->d_keyidx); pool_put(ctx, p->d_count); pool_put(ctx, p->d_offset);
    delete p;
    ctx->neigh = nullptr;
}

static NeighArgs neigh_args(NeighState *p) {
    NeighArgs a;
    a.sorted = p->slot->d_sorted; a.geom = p->slot->d_geom; a.cell_start = p->slot->d_cell_start; a.orig = p->slot->d_orig;
    a.cn_thr2 = p->d_cnthr2; a.keyidx = p->d_keyidx; a.count = p->d_count; a.offset = p->d_offset; a.nbr = nullptr;
    a.wraps = p->slot->d_wraps; a.dist = nullptr; a.shifts = nullptr;
    a.r2search = p->r2search; a.n_atoms = p->bt.n_atoms; a.n_keep = p->bt.n_keep; a.n_species = p->n_species;
    return a;
}

extern "C" int amofb_neigh_count(amofb_ctx *ctx, int n_atoms, int n_species, const uint8_t *species, const double *cutoff,
                                 const double *pos, const double *cell, int64_t *offsets) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->neigh) neigh_release(ctx);           // a count that was never followed by its fill
    if (n_atoms < 0 || n_species < 1 || n_species > AMOFB_MAX_SPECIES || (n_atoms > 0 && (!species || !pos)) || !cutoff || !cell || !offsets)
        return amofb_fail(ctx, AMOFB_ERR_ARG, "bad neighbour-list arguments");
    for (int i = 0; i < n_atoms; ++i)
        if (species[i] >= n_species) return amofb_fail(ctx, AMOFB_ERR_ARG, "species[%d] = %d out of range", i, species[i]);
    const int S = n_species;
    double cut_max = 0.0;
    for (int a = 0; a < S; ++a)
        for (int b = 0; b < S; ++b) {
            double c = cutoff[a * S + b];
            if (!(c >= 0.0) || !isfinite(c)) return amofb_fail(ctx, AMOFB_ERR_ARG, "cutoff[%d][%d] must be finite and >= 0", a, b);
            if (c != cutoff[b * S + a]) return amofb_fail(ctx, AMOFB_ERR_ARG, "cutoff matrix must be symmetric");
            cut_max = std::max(cut_max, c);
        }
    offsets[0] = 0;
    if (n_atoms == 0 || !(cut_max > 0.0)) {       // nothing can be a neighbour: empty rows, and amofb_neigh_fill has nothing to do
        for (int i = 0; i < n_atoms; ++i) offsets[i + 1] = 0;
        return AMOFB_OK;
    }
    NeighState *p = new (std::nothrow) NeighState();
    if (!p) return AMOFB_ERR_MEMORY;
    ctx->neigh = p;
    p->n_species = S; p->nkeys = S * (S + 1) / 2;
    std::vector<double> cnthr((size_t)p->nkeys, 0.0);
    for (int a = 0; a < S; ++a)
        for (int b = a; b < S; ++b) {
            double c = cutoff[a * S + b];
            cnthr[fold_key(a, b, S)] = c > 0.0 ? host_threshold(c * c, [&](double t) { return sqrt(t) >= c; }) : 0.0;     // P5
        }
    for (double t : cnthr) p->r2search = std::max(p->r2search, t);
    std::vector<uint16_t> keyidx((size_t)S * S);
    for (int a = 0; a < S; ++a)
        for (int b = 0; b < S; ++b) keyidx[a * S + b] = (uint16_t)fold_key(a, b, S);
    int rc = AMOFB_OK;
    auto fail = [&](int code) { neigh_release(ctx); return code; };
    p->bt.want_orig = true;
    // one frame per call: the batch buffers are sized for one frame
    if ((rc = batcher_init(ctx, p->bt, n_atoms, species, cut_max, 1, 0, 1))) return fail(rc);
    {
        uint8_t keep[AMOFB_MAX_SPECIES];
        memset(keep, 0, sizeof keep);
        for (int x = 0; x < S; ++x)
            for (int y = 0; y < S; ++y)
                if (cutoff[x * S + y] > 0.0) keep[x] = 1;
        if ((rc = batcher_set_filter(ctx, p->bt, species, keep))) return fail(rc);
    }
    if ((rc = dev_alloc(ctx, &p->d_cnthr2, cnthr.size()))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_keyidx, keyidx.size()))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_count, (size_t)n_atoms))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_offset, (size_t)n_atoms + 1))) return fail(rc);
    cudaMemcpy(p->d_cnthr2, cnthr.data(), sizeof(double) * cnthr.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p->d_keyidx, keyidx.data(), sizeof(uint16_t) * keyidx.size(), cudaMemcpyHostToDevice);
    cudaMemsetAsync(p->d_count, 0, sizeof(int) * (size_t)n_atoms, ctx->s_compute);
    const double *raw = nullptr;
    if ((rc = batcher_stage(ctx, p->bt, 1, pos, false, cell, &p->slot, &raw))) return fail(rc);
    if (p->bt.n_keep > 0) {
        k_neigh<false><<<(p->bt.n_keep + 127) / 128, 128, 0, ctx->s_compute>>>(neigh_args(p));
        ctx->launches += 1;
    }
    if ((rc = batcher_commit(ctx, p->bt, *p->slot, 1))) return fail(rc);
    std::vector<int> count((size_t)n_atoms);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(count.data(), p->d_count, sizeof(int) * (size_t)n_atoms, cudaMemcpyDeviceToHost, ctx->s_compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_compute);
    if (e != cudaSuccess) { amofb_fail(ctx, AMOFB_ERR_CUDA, "neigh_count: %s", cudaGetErrorString(e)); return fail(AMOFB_ERR_CUDA); }
    for (int i = 0; i < n_atoms; ++i) offsets[i + 1] = offsets[i] + count[i];
    p->total = offsets[n_atoms];
    std::vector<long long> off64((size_t)n_atoms + 1);
    for (int i = 0; i <= n_atoms; ++i) off64[i] = offsets[i];
    e = cudaMemcpy(p->d_offset, off64.data(), sizeof(long long) * off64.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { amofb_fail(ctx, AMOFB_ERR_CUDA, "neigh_count: %s", cudaGetErrorString(e)); return fail(AMOFB_ERR_CUDA); }
    if (p->total == 0) neigh_release(ctx);
    return AMOFB_OK;
}

extern "C" int amofb_neigh_fill_ex(amofb_ctx *ctx, int32_t *neighbors, double *distances, int32_t *shifts, int64_t capacity) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    NeighState *p = ctx->neigh;
    if (!p) return capacity >= 0 ? AMOFB_OK : amofb_fail(ctx, AMOFB_ERR_ARG, "negative capacity");      // count found no pair at all
    auto fail = [&](int code) { neigh_release(ctx); return code; };
    if (!neighbors || capacity < p->total) {
        amofb_fail(ctx, AMOFB_ERR_ARG, "neighbour buffer holds %lld entries, %lld are needed", (long long)capacity, p->total);
        return fail(AMOFB_ERR_ARG);
    }
    int *d_nbr = nullptr, *d_shifts = nullptr;
    double *d_dist = nullptr;
    int rc = dev_alloc(ctx, &d_nbr, (size_t)p->total);
    if (!rc && distances) rc = dev_alloc(ctx, &d_dist, (size_t)p->total);
    if (!rc && shifts) rc = dev_alloc(ctx, &d_shifts, (size_t)p->total * 3);
    if (rc) { pool_put(ctx, d_nbr); pool_put(ctx, d_dist); pool_put(ctx, d_shifts); return fail(rc); }
    NeighArgs a = neigh_args(p);
    a.nbr = d_nbr; a.dist = d_dist; a.shifts = d_shifts;
    k_neigh<true><<<(p->bt.n_keep + 127) / 128, 128, 0, ctx->s_compute>>>(a);
    ctx->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(neighbors, d_nbr, sizeof(int) * (size_t)p->total, cudaMemcpyDeviceToHost, ctx->s_compute);
    if (e == cudaSuccess && distances) e = cudaMemcpyAsync(distances, d_dist, sizeof(double) * (size_t)p->total, cudaMemcpyDeviceToHost, ctx->s_compute);
    if (e == cudaSuccess && shifts) e = cudaMemcpyAsync(shifts, d_shifts, sizeof(int) * 3 * (size_t)p->total, cudaMemcpyDeviceToHost, ctx->s_compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_compute);
    pool_put(ctx, d_nbr); pool_put(ctx, d_dist); pool_put(ctx, d_shifts);
    if (e != cudaSuccess) { amofb_fail(ctx, AMOFB_ERR_CUDA, "neigh_fill: %s", cudaGetErrorString(e)); return fail(AMOFB_ERR_CUDA); }
    neigh_release(ctx);
    return AMOFB_OK;
}

extern "C" int amofb_neigh_fill(amofb_ctx *ctx, int32_t *neighbors, int64_t capacity) {
    return amofb_neigh_fill_ex(ctx, neighbors, nullptr, nullptr, capacity);
}
