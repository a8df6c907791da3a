// bad_host.inl -- host side of the bond-angle analysis (included by amofb.cu)

struct BadState {
    Batcher bt;
    int n_species = 0, nkeys = 0, n_triples = 0, nbins = 0;
    double dtheta = 0.0, rcut = 0.0, r2search = 0.0;
    double *d_cnthr2 = nullptr, *d_tthr = nullptr;
    uint16_t *d_keyidx = nullptr;
    int2 *d_triples = nullptr;
    unsigned long long *d_hist = nullptr, *d_dropped = nullptr;
    int *d_flags = nullptr;
    BadNb *d_pool = nullptr;          // unit vectors of the neighbours of every centre of a batch
    BadCentre *d_centres = nullptr;
    unsigned *d_counters = nullptr;   // [0] centres, [1] pool entries
    unsigned long long *d_centre_mask = nullptr;
    unsigned pool_cap = 0;
    int n_slots = 0, tthr_smem = 0, angle_grid = 0;
    size_t angle_smem = 0;
    unsigned long long centre_mask[AMOFB_MAX_SPECIES];
    uint16_t partner_mask[AMOFB_MAX_SPECIES];     // species with a positive cutoff to this one
};

static void bad_release(amofb_ctx *ctx) {
    BadState *p = ctx->bad;
    if (!p) return;
    cudaStreamSynchronize(ctx->s_copy);
    cudaStreamSynchronize(ctx->s_compute);
    batcher_release(ctx, p->bt);
    pool_put(ctx, p->d_cnthr2); pool_put(ctx, p->d_tthr); pool_put(ctx, p->d_keyidx); pool_put(ctx, p->d_triples);
    pool_put(ctx, p->d_hist); pool_put(ctx, p->d_dropped); pool_put(ctx, p->d_flags);
    pool_put(ctx, p->d_pool); pool_put(ctx, p->d_centres); pool_put(ctx, p->d_counters); pool_put(ctx, p->d_centre_mask);
    delete p;
    ctx->bad = nullptr;
}

// P7: np.histogram(theta, bins=edges) with edges e_k = k*dtheta, k = 0..nbins.
// Returns the bin, nbins for "above the last edge" (dropped), never called with NaN.
static inline int host_theta_index(double theta, double dtheta, int nbins) {
    const double last = (double)nbins * dtheta;
    if (theta > last) return nbins;
    if (theta == last) return nbins - 1;          // the last bin is closed on the right
    if (theta < 0.0) return 0;
    long k = (long)(theta / dtheta);
    if (k > nbins - 1) k = nbins - 1;
    while (k > 0 && theta < (double)k * dtheta) --k;                  // settle on e_k <= theta < e_{k+1}
    while (k < nbins - 1 && theta >= (double)(k + 1) * dtheta) ++k;
    return (int)k;
}

// order-preserving map double -> uint64 over the whole real line
static inline uint64_t dbl_key(double d) {
    uint64_t u;
    memcpy(&u, &d, 8);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
static inline double key_dbl(uint64_t k) {
    uint64_t u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double d;
    memcpy(&d, &u, 8);
    return d;
}

// tthr[k], k = 1..nbins: smallest t = -x in [-1, 1] whose angle acos(-t)*(180/pi) (P6, host libm) falls in
// bin >= k; 2.0 when no t does.  tthr[0] = -2, tthr[nbins+1] = +2 are sentinels.
static void host_angle_thresholds(double dtheta, int nbins, std::vector<double> &tthr) {
    const double deg = 180.0 / 3.14159265358979323846;
    tthr.assign((size_t)nbins + 2, 2.0);
    tthr[0] = -2.0;
    auto index_of = [&](double t) { return host_theta_index(acos(-t) * deg, dtheta, nbins); };
    const uint64_t klo = dbl_key(-1.0), khi = dbl_key(1.0);
    const int top = index_of(1.0);
    uint64_t prev = klo;   // thresholds are non-decreasing: start each bisection at the previous one
    for (int k = 1; k <= nbins; ++k) {
        if (top < k) break;                       // unreachable bins keep the +2 sentinel
        if (index_of(key_dbl(prev)) >= k) { tthr[k] = key_dbl(prev); continue; }
        uint64_t lo = prev, hi = khi;             // index(lo) < k <= index(hi)
        while (hi - lo > 1) {
            uint64_t mid = lo + (hi - lo) / 2;
            if (index_of(key_dbl(mid)) >= k) hi = mid; else lo = mid;
        }
        tthr[k] = key_dbl(hi);
        prev = hi;
    }
}

extern "C" int amofb_bad_begin(amofb_ctx *ctx, int n_atoms, int n_species, const uint8_t *species, const double *cutoff,
                               int n_triples, const int *triples, double dtheta, int nbins) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->bad) return amofb_fail(ctx, AMOFB_ERR_STATE, "bond-angle analysis already open; call amofb_bad_finish first");
    if (n_atoms < 0 || n_species < 1 || n_species > AMOFB_MAX_SPECIES || (n_atoms > 0 && !species) || !cutoff)
        return amofb_fail(ctx, AMOFB_ERR_ARG, "bad n_atoms/n_species/cutoff");
    if (n_triples < 1 || n_triples > BAD_MAX_TRIPLES || !triples)
        return amofb_fail(ctx, AMOFB_ERR_ARG, "n_triples must be 1..%d", BAD_MAX_TRIPLES);
    if (nbins < 1 || !(dtheta > 0.0) || !isfinite(dtheta)) return amofb_fail(ctx, AMOFB_ERR_ARG, "bad dtheta/nbins");
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
    for (int t = 0; t < n_triples; ++t)
        for (int k = 0; k < 2; ++k)
            if (triples[2 * t + k] < -1 || triples[2 * t + k] >= S)
                return amofb_fail(ctx, AMOFB_ERR_ARG, "triples[%d][%d] out of range", t, k);
    BadState *p = new (std::nothrow) BadState();
    if (!p) return AMOFB_ERR_MEMORY;
    ctx->bad = p;
    p->n_species = S; p->nkeys = S * (S + 1) / 2; p->n_triples = n_triples; p->nbins = nbins; p->dtheta = dtheta;
    p->rcut = cut_max;
    std::vector<double> cnthr((size_t)p->nkeys, 0.0), tthr;
    for (int a = 0; a < S; ++a)
        for (int b = a; b < S; ++b) {
            double c = cutoff[a * S + b];
            cnthr[fold_key(a, b, S)] = c > 0.0 ? host_threshold(c * c, [&](double t) { return sqrt(t) >= c; }) : 0.0;
        }
    p->r2search = 0.0;
    for (double t : cnthr) p->r2search = std::max(p->r2search, t);
    if (ctx->tthr_nbins == nbins && ctx->tthr_dtheta == dtheta && (int)ctx->tthr_cache.size() == nbins + 2) tthr = ctx->tthr_cache;
    else {
        host_angle_thresholds(dtheta, nbins, tthr);
        ctx->tthr_cache = tthr; ctx->tthr_dtheta = dtheta; ctx->tthr_nbins = nbins;
    }
    std::vector<uint16_t> keyidx((size_t)S * S);
    for (int a = 0; a < S; ++a)
        for (int b = 0; b < S; ++b) keyidx[a * S + b] = (uint16_t)fold_key(a, b, S);
    std::vector<int2> tr((size_t)n_triples);
    memset(p->centre_mask, 0, sizeof p->centre_mask);
    for (int t = 0; t < n_triples; ++t) {
        tr[t] = make_int2(triples[2 * t], triples[2 * t + 1]);
        for (int s = 0; s < S; ++s)
            if (tr[t].x < 0 || tr[t].x == s) p->centre_mask[s] |= 1ull << t;
    }
    int rc = AMOFB_OK;
    auto fail = [&](int code) { bad_release(ctx); return code; };
    double rcut = cut_max > 0.0 ? cut_max : 1e-3;
    int cell_div = env_int("AMOFB_BAD_CELL_DIV", 1);
    if (cell_div < 1) cell_div = 1;
    {
        // cell width in cutoffs (measured on C4: 1.0 -> 3.99, 1.25 -> 4.4, 1.5 -> 4.71 us per frame): the filtered frames are sparse (0.3 atoms per cutoff-wide cell on the ZIF 'Zn-N'
        // analysis), wider cells mean a third of the cells to count, scan and look up, for a few more distance evaluations
        const char *w = getenv("AMOFB_BAD_CELL_WIDEN");
        p->bt.cell_widen = w && *w ? atof(w) : 1.0;
        if (!(p->bt.cell_widen >= 1.0)) p->bt.cell_widen = 1.0;
    }
    uint8_t keep_pre[AMOFB_MAX_SPECIES];
    memset(keep_pre, 0, sizeof keep_pre);
    for (int x = 0; x < S; ++x)
        for (int y = 0; y < S; ++y)
            if (cutoff[x * S + y] > 0.0) keep_pre[x] = 1;
    int n_work = 0;
    for (int i = 0; i < n_atoms; ++i) n_work += keep_pre[species[i]];
    if ((rc = batcher_init(ctx, p->bt, n_atoms, species, rcut, cell_div, 0, 0, n_work))) return fail(rc);
    {   // species filter: an atom whose species has no positive cutoff with any species can neither be a centre with
        // neighbours nor a neighbour, so it never enters the cell list (ZIF-4 'Zn-N': 71 % of the atoms drop out)
        uint8_t keep[AMOFB_MAX_SPECIES];
        memset(keep, 0, sizeof keep);
        for (int x = 0; x < S; ++x)
            for (int y = 0; y < S; ++y)
                if (cutoff[x * S + y] > 0.0) keep[x] = 1;
        if ((rc = batcher_set_filter(ctx, p->bt, species, keep, env_int("AMOFB_BAD_NO_CENTRE_LIST", 0) ? nullptr : p->centre_mask))) return fail(rc);
        // one cell list per kept species, and for every species the partners it has a positive cutoff with
        memset(p->partner_mask, 0, sizeof p->partner_mask);
        int nl = 0;
        for (int x = 0; x < S; ++x) {
            if (keep[x]) p->bt.list_of[x] = (uint8_t)nl++;
            for (int y = 0; y < S; ++y)
                if (cutoff[x * S + y] > 0.0) p->partner_mask[x] |= (uint16_t)(1u << y);
        }
        p->bt.n_lists = (nl > 1 && !env_int("AMOFB_BAD_ONE_LIST", 0)) ? nl : 1;
        if (p->bt.n_lists == 1) memset(p->bt.list_of, 0, sizeof p->bt.list_of);
    }
    const size_t hist_n = (size_t)n_triples * (AMOFB_BAD_MAX_CN + 1) * nbins;
    if ((rc = dev_alloc(ctx, &p->d_cnthr2, cnthr.size()))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_tthr, tthr.size()))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_keyidx, keyidx.size()))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_triples, tr.size()))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_hist, hist_n))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_dropped, (size_t)n_triples))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_flags, 1))) return fail(rc);
    {
        // neighbour pool of one batch: 16 neighbours per filtered atom, at most 512 MB (an overflow is reported, not dropped)
        const size_t want = (size_t)std::max(p->bt.n_keep, 1) * (size_t)p->bt.cap_frames * 16;
        p->pool_cap = (unsigned)std::min<size_t>(std::max<size_t>(want, 4096), (size_t)16 << 20);
        if ((rc = dev_alloc(ctx, &p->d_pool, (size_t)p->pool_cap))) return fail(rc);
        if ((rc = dev_alloc(ctx, &p->d_centres, (size_t)std::max(p->bt.n_keep, 1) * (size_t)p->bt.cap_frames))) return fail(rc);
        if ((rc = dev_alloc(ctx, &p->d_counters, 2))) return fail(rc);
        if ((rc = dev_alloc(ctx, &p->d_centre_mask, (size_t)AMOFB_MAX_SPECIES))) return fail(rc);
        cudaMemcpy(p->d_centre_mask, p->centre_mask, sizeof p->centre_mask, cudaMemcpyHostToDevice);
        // shared memory of the angle kernel: threshold table + as many histogram rows as fit (two blocks per SM)
        const size_t budget = ((size_t)ctx->max_smem_optin + 1024) / 2 - 1024 - 256;
        const size_t tab = sizeof(double) * ((size_t)nbins + 2), row = sizeof(uint32_t) * (size_t)nbins;
        p->tthr_smem = tab <= budget ? 1 : 0;
        size_t left = budget - (p->tthr_smem ? tab : 0);
        p->n_slots = (int)std::min<size_t>(BAD_SLOTS, left / row);
        p->angle_smem = (p->tthr_smem ? tab : 0) + row * (size_t)p->n_slots;
        cudaError_t e1 = cudaFuncSetAttribute(k_bad_angles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->angle_smem);
        int per_sm = 0;
        if (e1 == cudaSuccess) e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bad_angles, 256, p->angle_smem);
        if (e1 != cudaSuccess || per_sm < 1) { amofb_fail(ctx, AMOFB_ERR_CUDA, "bad_begin: angle kernel does not fit on an SM"); return fail(AMOFB_ERR_CUDA); }
        p->angle_grid = ctx->num_sms * per_sm;
    }
    cudaMemcpy(p->d_cnthr2, cnthr.data(), sizeof(double) * cnthr.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p->d_tthr, tthr.data(), sizeof(double) * tthr.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p->d_keyidx, keyidx.data(), sizeof(uint16_t) * keyidx.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p->d_triples, tr.data(), sizeof(int2) * tr.size(), cudaMemcpyHostToDevice);
    cudaMemset(p->d_hist, 0, sizeof(unsigned long long) * hist_n);
    cudaMemset(p->d_dropped, 0, sizeof(unsigned long long) * n_triples);
    cudaMemset(p->d_flags, 0, sizeof(int));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { amofb_fail(ctx, AMOFB_ERR_CUDA, "bad_begin: %s", cudaGetErrorString(e)); return fail(AMOFB_ERR_CUDA); }
    return AMOFB_OK;
}

static int bad_push_impl(amofb_ctx *ctx, int n_frames, const double *pos, bool on_device, const double *cell) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    BadState *p = ctx->bad;
    if (!p) return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_bad_push before amofb_bad_begin");
    if (n_frames < 0 || (n_frames > 0 && (!cell || (!pos && p->bt.n_atoms > 0))))
        return amofb_fail(ctx, AMOFB_ERR_ARG, "bad push arguments");
    Batcher &b = p->bt;
    // P6 precondition: the neighbour image vector is the minimum image only below half the smallest height
    if (p->rcut > 0.0)
        for (int f = 0; f < n_frames; ++f) {
            double inv[9], h[3];
            if (!host_cell_inverse(cell + 9 * (size_t)f, inv))
                return amofb_fail(ctx, AMOFB_ERR_GEOMETRY, "frame %lld: singular cell", (long long)(b.frames_seen + f));
            host_cell_heights(inv, h);
            for (int k = 0; k < 3; ++k)
                if (!(p->rcut < 0.5 * h[k]))
                    return amofb_fail(ctx, AMOFB_ERR_GEOMETRY, "frame %lld: cutoff %g is not below half the cell height %g",
                                      (long long)(b.frames_seen + f), p->rcut, h[k]);
        }
    for (int done = 0; done < n_frames;) {
        int nf = std::min(b.cap_frames, n_frames - done);
        BatchSlot *s = nullptr;
        const double *raw = nullptr;
        AMOFB_TRY(batcher_stage(ctx, b, nf, pos + 3 * (size_t)done * b.n_atoms, on_device, cell + 9 * (size_t)done, &s, &raw));
        BadArgs a;
        a.sorted = s->d_sorted; a.geom = s->d_geom; a.cell_start = s->d_cell_start;
        a.cn_thr2 = p->d_cnthr2; a.keyidx = p->d_keyidx; a.triples = p->d_triples; a.tthr = p->d_tthr;
        a.hist = p->d_hist; a.dropped = p->d_dropped; a.flags = p->d_flags;
        a.n_keep = b.n_keep;
        a.centre_mask = p->d_centre_mask;
        a.centre_list = b.d_centre_list; a.n_centres = b.n_centres;
        a.n_lists = b.n_lists;
        memcpy(a.list_of, b.list_of, sizeof a.list_of);
        memcpy(a.partner_mask, p->partner_mask, sizeof a.partner_mask);
        a.r2search = p->r2search; a.inv_dtheta_f = (float)(1.0 / p->dtheta);
        a.n_atoms = b.n_atoms; a.n_frames = nf; a.n_species = p->n_species; a.nkeys = p->nkeys;
        a.n_triples = p->n_triples; a.nbins = p->nbins;
        a.pool = p->d_pool; a.centres = p->d_centres; a.counters = p->d_counters; a.pool_cap = p->pool_cap;
        a.n_slots = p->n_slots; a.tthr_smem = p->tthr_smem;
        long long total = (long long)nf * b.n_keep;
        if (total > 0) {
            CUDA_TRY(ctx, cudaMemsetAsync(p->d_counters, 0, sizeof(unsigned) * 2, ctx->s_compute));
            // one thread per possible centre (compact list written by the scatter kernel), or per atom of the filtered, cell-sorted frames
            const long long nsearch = b.d_centre_list ? (long long)nf * b.n_centres : total;
            if (nsearch > 0) k_bad_search<<<(unsigned)((nsearch + 127) / 128), 128, 0, ctx->s_compute>>>(a);
            k_bad_angles<<<p->angle_grid, 256, p->angle_smem, ctx->s_compute>>>(a);         // one thread per centre found
            ctx->launches += 2;
            CUDA_TRY(ctx, cudaGetLastError());
        }
        AMOFB_TRY(batcher_commit(ctx, b, *s, nf));
        done += nf;
    }
    return AMOFB_OK;
}

extern "C" int amofb_bad_push(amofb_ctx *ctx, int n_frames, const double *pos, const double *cell) {
    nvtx_range rng("amofb_bad_push");
    return bad_push_impl(ctx, n_frames, pos, false, cell);
}
extern "C" int amofb_bad_push_device(amofb_ctx *ctx, int n_frames, const double *pos_device, const double *cell) {
    nvtx_range rng("amofb_bad_push_device");
    return bad_push_impl(ctx, n_frames, pos_device, true, cell);
}

extern "C" int amofb_bad_finish(amofb_ctx *ctx, uint64_t *hist, uint64_t *dropped, int64_t *n_frames_out) {
    nvtx_range rng("amofb_bad_finish");
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    BadState *p = ctx->bad;
    if (!p) return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_bad_finish before amofb_bad_begin");
    auto body = [&]() -> int {
        AMOFB_TRY(batcher_drain(ctx, p->bt));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->s_compute));
        int flags = 0;
        CUDA_TRY(ctx, cudaMemcpy(&flags, p->d_flags, sizeof(int), cudaMemcpyDeviceToHost));
        if (flags & 1) return amofb_fail(ctx, AMOFB_ERR_GEOMETRY, "a centre has more than %d neighbours under the cutoffs", BAD_NB_MAX);
        if (flags & 2) return amofb_fail(ctx, AMOFB_ERR_GEOMETRY, "a centre has more than %d B-neighbours", AMOFB_BAD_MAX_CN);
        if (flags & 4) return amofb_fail(ctx, AMOFB_ERR_MEMORY, "neighbour pool of a batch overflowed (%u entries); rerun with a smaller AMOFB_BATCH_ATOMS", p->pool_cap);
        const size_t hist_n = (size_t)p->n_triples * (AMOFB_BAD_MAX_CN + 1) * p->nbins;
        if (hist) CUDA_TRY(ctx, cudaMemcpy(hist, p->d_hist, sizeof(uint64_t) * hist_n, cudaMemcpyDeviceToHost));
        if (dropped) CUDA_TRY(ctx, cudaMemcpy(dropped, p->d_dropped, sizeof(uint64_t) * p->n_triples, cudaMemcpyDeviceToHost));
        if (n_frames_out) *n_frames_out = p->bt.frames_seen;
        return AMOFB_OK;
    };
    int rc = body();
    bad_release(ctx);
    return rc;
}
