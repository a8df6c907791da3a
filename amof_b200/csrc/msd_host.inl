// msd_host.inl -- host side of the mean-squared-displacement analysis (included by amofb.cu)

struct MsdState {
    int T = 0, n = 0, S = 0;
    double *d_P = nullptr;            // [n][T][3] atom-major
    MsdGeom *d_geom = nullptr;        // [T]
    double *d_masses = nullptr;       // [n]
    uint8_t *d_species = nullptr;     // [n]
    double *d_com = nullptr;          // [T][3]
    double *d_stage[3] = {nullptr, nullptr, nullptr};      // staging slots of host slabs (three: two slabs may await their commit while the next one is copied)
    cudaEvent_t ev_stage[3] = {nullptr, nullptr, nullptr};
    int stage_frames = 0, next_stage = 0;
    bool have_com = false, prepared = false, consumed = false, fixed_cell = false, diag_cell = false;
    std::vector<double> cell;         // host copy [T][9]
    double mass_sum = 0.0;
    // streaming path (amofb_msd_slab_*): prepared series, three arrays of Tp doubles per atom
    int Tp = 0;                       // T rounded up to even (16-byte aligned rows)
    bool soa = false;                 // d_P holds (or is being filled with) the prepared SoA series
    int slab_next = 0;                // next frame the streaming path expects
    struct PendingSlab {              // a slab whose sums were enqueued and that awaits its commit
        const double *ptr = nullptr;  // device pointer
        int first = 0, count = 0, slot = -1;
        double *d_partial = nullptr, *d_out = nullptr, *h_out = nullptr;
        cudaEvent_t ev = nullptr;     // the sums have landed in h_out
        bool waited = false;
    } pend[2];
    int n_pend = 0;                   // FIFO: pend[0] is the oldest
    int sums_next = 0;                // next frame slab_sums expects
    double *d_carry = nullptr;        // [n][6]
};

static void msd_release(amofb_ctx *ctx) {
    MsdState *p = ctx->msd;
    if (!p) return;
    cudaStreamSynchronize(ctx->s_copy);
    cudaStreamSynchronize(ctx->s_compute);
    pool_put(ctx, p->d_P); pool_put(ctx, p->d_geom); pool_put(ctx, p->d_masses); pool_put(ctx, p->d_species); pool_put(ctx, p->d_com); pool_put(ctx, p->d_carry);
    for (int i = 0; i < 3; ++i) {
        pool_put(ctx, p->d_stage[i]);
        if (p->ev_stage[i]) cudaEventDestroy(p->ev_stage[i]);
    }
    for (int i = 0; i < 2; ++i) {
        pool_put(ctx, p->pend[i].d_partial); pool_put(ctx, p->pend[i].d_out); pool_put(ctx, p->pend[i].h_out);
        if (p->pend[i].ev) cudaEventDestroy(p->pend[i].ev);
    }
    delete p;
    ctx->msd = nullptr;
}

extern "C" int amofb_msd_begin(amofb_ctx *ctx, int n_frames, int n_atoms, const double *masses, const uint8_t *species,
                               int n_species, const double *cell) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->msd) return amofb_fail(ctx, AMOFB_ERR_STATE, "MSD analysis already open; call amofb_msd_end first");
    if (n_frames < 1 || n_atoms < 1 || n_species < 1 || n_species > AMOFB_MAX_SPECIES || !masses || !species || !cell)
        return amofb_fail(ctx, AMOFB_ERR_ARG, "bad MSD arguments");
    for (int i = 0; i < n_atoms; ++i)
        if (species[i] >= n_species) return amofb_fail(ctx, AMOFB_ERR_ARG, "species[%d] = %d out of range", i, species[i]);
    MsdState *p = new (std::nothrow) MsdState();
    if (!p) return AMOFB_ERR_MEMORY;
    ctx->msd = p;
    p->T = n_frames; p->n = n_atoms; p->S = n_species;
    p->Tp = (n_frames + 1) & ~1;
    p->cell.assign(cell, cell + 9 * (size_t)n_frames);
    std::vector<MsdGeom> geom((size_t)n_frames);
    int rc = AMOFB_OK;
    auto fail = [&](int code) { msd_release(ctx); return code; };
    for (int k = 0; k < n_frames; ++k) {
        memcpy(geom[k].cell, cell + 9 * (size_t)k, sizeof(double) * 9);
        if (!host_cell_inverse(cell + 9 * (size_t)k, geom[k].inv)) {
            amofb_fail(ctx, AMOFB_ERR_GEOMETRY, "frame %d: singular cell", k);
            return fail(AMOFB_ERR_GEOMETRY);
        }
    }
    for (int i = 0; i < n_atoms; ++i) p->mass_sum += masses[i];
    p->fixed_cell = true;
    for (int k = 1; k < n_frames && p->fixed_cell; ++k)
        if (memcmp(cell + 9 * (size_t)k, cell, sizeof(double) * 9) != 0) p->fixed_cell = false;
    // orthorhombic box along the axes: the off-diagonal entries of the cell AND of its inverse are all zero
    p->diag_cell = p->fixed_cell;
    for (int q = 0; q < 9 && p->diag_cell; ++q)
        if (q % 4 != 0 && (geom[0].cell[q] != 0.0 || geom[0].inv[q] != 0.0)) p->diag_cell = false;
    if ((rc = dev_alloc(ctx, &p->d_P, (size_t)n_atoms * p->Tp * 3))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_carry, (size_t)n_atoms * 6))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_geom, (size_t)n_frames))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_masses, (size_t)n_atoms))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_species, (size_t)n_atoms))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_com, (size_t)n_frames * 3))) return fail(rc);
    // staging slots: a multiple of 32 frames (the rounds of the commit kernel, 256-byte output runs) when 32 frames stay
    // below 1.5 GiB per slot, at most 256 frames
    size_t frame_bytes = sizeof(double) * 3 * (size_t)n_atoms;
    long long sf = (long long)((3ull << 29) / frame_bytes);
    sf = std::min<long long>(sf, 256);
    if (sf >= 32) sf -= sf % 32;
    p->stage_frames = (int)std::max<long long>(1, std::min<long long>(sf, n_frames));
    for (int i = 0; i < 3; ++i) {
        if ((rc = dev_alloc(ctx, &p->d_stage[i], (size_t)p->stage_frames * n_atoms * 3))) return fail(rc);
        cudaEventCreateWithFlags(&p->ev_stage[i], cudaEventDisableTiming);
    }
    if (p->Tp > p->T)           // the pad frame of every series (T odd): the window kernels copy whole series and rely on zeros after frame T-1
        cudaMemset2D(p->d_P + p->T, sizeof(double) * (size_t)p->Tp, 0, sizeof(double), 3 * (size_t)n_atoms);
    cudaMemcpy(p->d_geom, geom.data(), sizeof(MsdGeom) * geom.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p->d_masses, masses, sizeof(double) * n_atoms, cudaMemcpyHostToDevice);
    cudaMemcpy(p->d_species, species, (size_t)n_atoms, cudaMemcpyHostToDevice);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { amofb_fail(ctx, AMOFB_ERR_CUDA, "msd_begin: %s", cudaGetErrorString(e)); return fail(AMOFB_ERR_CUDA); }
    return AMOFB_OK;
}

static int msd_state(amofb_ctx *ctx, MsdState **out, const char *what) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->msd) return amofb_fail(ctx, AMOFB_ERR_STATE, "%s before amofb_msd_begin", what);
    *out = ctx->msd;
    return AMOFB_OK;
}

static void msd_transpose_launch(amofb_ctx *ctx, MsdState *p, const double *src, int first, int count) {
    dim3 grid((p->n + 31) / 32, (count + 31) / 32);
    k_msd_transpose<<<grid, 256, 0, ctx->s_compute>>>(src, p->d_P, p->n, p->T, first, count);
    ctx->launches += 1;
}

static int msd_load_impl(amofb_ctx *ctx, int first_frame, int count, const double *pos, bool on_device) {
    MsdState *p = nullptr;
    AMOFB_TRY(msd_state(ctx, &p, "amofb_msd_load"));
    if (p->prepared || p->consumed || p->soa) return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_msd_load after the positions were transformed");
    if (first_frame < 0 || count < 0 || first_frame + (long long)count > p->T || (count > 0 && !pos))
        return amofb_fail(ctx, AMOFB_ERR_ARG, "frames [%d, %d) outside [0, %d)", first_frame, first_frame + count, p->T);
    p->have_com = false;
    const size_t fr = 3 * (size_t)p->n;
    for (int done = 0; done < count;) {
        int nf = std::min(p->stage_frames, count - done);
        const double *src = pos + fr * done;
        if (!on_device) {
            int sl = p->next_stage;
            p->next_stage = (p->next_stage + 1) % 3;
            CUDA_TRY(ctx, cudaEventSynchronize(p->ev_stage[sl]));   // the transpose that last read this slot is done
            CUDA_TRY(ctx, cudaMemcpyAsync(p->d_stage[sl], src, sizeof(double) * fr * nf, cudaMemcpyHostToDevice, ctx->s_copy));
            cudaEvent_t ev;
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventRecord(ev, ctx->s_copy));
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->s_compute, ev, 0));
            CUDA_TRY(ctx, cudaEventDestroy(ev));
            msd_transpose_launch(ctx, p, p->d_stage[sl], first_frame + done, nf);
            CUDA_TRY(ctx, cudaGetLastError());
            CUDA_TRY(ctx, cudaEventRecord(p->ev_stage[sl], ctx->s_compute));
        } else {
            msd_transpose_launch(ctx, p, src, first_frame + done, nf);
            CUDA_TRY(ctx, cudaGetLastError());
        }
        done += nf;
    }
    return AMOFB_OK;
}

extern "C" int amofb_msd_load(amofb_ctx *ctx, int first_frame, int count, const double *pos) {
    return msd_load_impl(ctx, first_frame, count, pos, false);
}
extern "C" int amofb_msd_load_device(amofb_ctx *ctx, int first_frame, int count, const double *pos_device) {
    return msd_load_impl(ctx, first_frame, count, pos_device, true);
}

static int msd_scan_grid(amofb_ctx *ctx, int n) {
    long long warps = n;
    long long blocks = (warps + SCAN_WARPS - 1) / SCAN_WARPS;
    return (int)std::max<long long>(1, std::min<long long>(blocks, (long long)ctx->num_sms * 4));
}

extern "C" int amofb_msd_unwrap(amofb_ctx *ctx) {
    MsdState *p = nullptr;
    AMOFB_TRY(msd_state(ctx, &p, "amofb_msd_unwrap"));
    if (p->prepared || p->consumed || p->soa) return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_msd_unwrap after the positions were transformed");
    if (p->fixed_cell) k_msd_scan<false, true><<<msd_scan_grid(ctx, p->n), 32 * SCAN_WARPS, 0, ctx->s_compute>>>(p->d_P, p->d_geom, nullptr, p->n, p->T);
    else k_msd_scan<false, false><<<msd_scan_grid(ctx, p->n), 32 * SCAN_WARPS, 0, ctx->s_compute>>>(p->d_P, p->d_geom, nullptr, p->n, p->T);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    p->have_com = false;
    return AMOFB_OK;
}

// out[len] = per-frame sums over atoms with weights d_w, NC components
template <int NC>
static int msd_frame_sums(amofb_ctx *ctx, MsdState *p, const double *d_w, double *d_out) {
    int groups = std::max(1, std::min(64, p->n / 256));
    double *d_partial = nullptr;
    AMOFB_TRY(dev_alloc(ctx, &d_partial, (size_t)groups * p->T * NC));
    dim3 grid((p->T + 31) / 32, groups);
    k_msd_frame_sums<NC><<<grid, 256, 0, ctx->s_compute>>>(p->d_P, d_w, p->n, p->T, d_partial);
    k_msd_sum_groups<<<std::max(1, std::min((p->T * NC + 255) / 256, ctx->num_sms * 4)), 256, 0, ctx->s_compute>>>(d_partial, groups, p->T * NC, d_out);
    ctx->launches += 2;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_compute);
    pool_put(ctx, d_partial);
    CUDA_TRY(ctx, e);
    return AMOFB_OK;
}

extern "C" int amofb_msd_com_sums(amofb_ctx *ctx, double *sums) {
    MsdState *p = nullptr;
    AMOFB_TRY(msd_state(ctx, &p, "amofb_msd_com_sums"));
    if (p->prepared || p->consumed || p->soa) return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_msd_com_sums after the positions were transformed");
    if (!sums) return amofb_fail(ctx, AMOFB_ERR_ARG, "null output");
    AMOFB_TRY(msd_frame_sums<3>(ctx, p, p->d_masses, p->d_com));   // d_com temporarily holds the weighted sums
    std::vector<double> tmp((size_t)p->T * 3);
    CUDA_TRY(ctx, cudaMemcpy(tmp.data(), p->d_com, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
    for (int k = 0; k < p->T; ++k) {
        sums[4 * (size_t)k] = tmp[3 * (size_t)k];
        sums[4 * (size_t)k + 1] = tmp[3 * (size_t)k + 1];
        sums[4 * (size_t)k + 2] = tmp[3 * (size_t)k + 2];
        sums[4 * (size_t)k + 3] = p->mass_sum;
    }
    return AMOFB_OK;
}

extern "C" int amofb_msd_set_com(amofb_ctx *ctx, const double *com) {
    MsdState *p = nullptr;
    AMOFB_TRY(msd_state(ctx, &p, "amofb_msd_set_com"));
    if (p->prepared || p->consumed || p->soa) return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_msd_set_com after the positions were transformed");
    if (!com) return amofb_fail(ctx, AMOFB_ERR_ARG, "null centre of mass");
    CUDA_TRY(ctx, cudaMemcpy(p->d_com, com, sizeof(double) * 3 * (size_t)p->T, cudaMemcpyHostToDevice));
    p->have_com = true;
    return AMOFB_OK;
}

// k_msd_scan<PREPARE>: centre-of-mass shift, displacement wrap and running sum written back over the trajectory
static int msd_prepare(amofb_ctx *ctx, MsdState *p) {
    if (!p->have_com) return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_msd_window needs amofb_msd_set_com first");
    if (p->fixed_cell) k_msd_scan<true, true><<<msd_scan_grid(ctx, p->n), 32 * SCAN_WARPS, 0, ctx->s_compute>>>(p->d_P, p->d_geom, p->d_com, p->n, p->T);
    else k_msd_scan<true, false><<<msd_scan_grid(ctx, p->n), 32 * SCAN_WARPS, 0, ctx->s_compute>>>(p->d_P, p->d_geom, p->d_com, p->n, p->T);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    p->prepared = true;
    return AMOFB_OK;
}

static int msd_window_soa(amofb_ctx *ctx, MsdState *p, int n_window, int ap_delta, double *sums);

extern "C" int amofb_msd_window(amofb_ctx *ctx, int n_window, const int *window, double *sums) {
    MsdState *p = nullptr;
    AMOFB_TRY(msd_state(ctx, &p, "amofb_msd_window"));
    if (p->consumed) return amofb_fail(ctx, AMOFB_ERR_STATE, "positions were consumed by amofb_msd_direct");
    if (n_window < 0 || (n_window > 0 && (!window || !sums))) return amofb_fail(ctx, AMOFB_ERR_ARG, "bad window arguments");
    const int S = p->S;
    if (p->soa) {
        if (p->n_pend > 0 || p->slab_next != p->T) return amofb_fail(ctx, AMOFB_ERR_STATE, "the streaming path has committed %d of %d frames", p->slab_next, p->T);
        if (n_window == 0) return AMOFB_OK;
        int d = n_window >= 2 ? window[1] : 1;
        bool ap = window[0] == 0 && d > 0;
        for (int w = 2; w < n_window && ap; ++w) ap = (long long)window[w] == (long long)w * d;
        if (!ap || 4LL * d >= p->T)
            return amofb_fail(ctx, AMOFB_ERR_ARG, "the streaming path takes window lengths 0, D, 2D, ... with 4 D < n_frames; use amofb_msd_load for others");
        nvtx_range rng("amofb_msd_window");
        return msd_window_soa(ctx, p, n_window, d, sums);
    }
    if (n_window == 0) {                                    // nothing to sum: only bring the trajectory into the prepared state
        if (!p->prepared) AMOFB_TRY(msd_prepare(ctx, p));
        return AMOFB_OK;
    }
    int threads = MSD_THREADS;
    // window lengths 0, D, 2D, ... (what WindowMsd always asks for, msd.py:176-178) -> register-tiled kernel
    int ap_delta = 0;
    if (n_window >= 2 && window[0] == 0 && window[1] > 0 && !env_int("AMOFB_MSD_NO_AP", 0)) {
        ap_delta = window[1];
        for (int w = 2; w < n_window && ap_delta; ++w)
            if ((long long)window[w] != (long long)w * ap_delta) ap_delta = 0;
        if (ap_delta && (long long)MSD_AP_KB * ap_delta >= p->T) ap_delta = 0;      // not a single full tile: nothing to gain
    }
    // the register-tiled kernel prepares the series itself (centre-of-mass shift, wrap, running sum) while it is staged;
    // every other path needs k_msd_scan<PREPARE> to rewrite the trajectory first
    {
        size_t need = sizeof(double) * (3 * (size_t)p->T + 4 + (size_t)S * n_window + (size_t)MSD_NW * (MSD_AP_THREADS / 32) + 96);
        size_t budget0 = (size_t)ctx->max_smem_optin > 1024 ? (size_t)ctx->max_smem_optin - 1024 : 0;
        if (need > budget0 || env_int("AMOFB_MSD_NO_SMEM", 0)) ap_delta = 0;
    }
    const bool fuse_prepare = ap_delta && !p->prepared && !env_int("AMOFB_MSD_NO_FUSE", 0);
    if (!p->prepared && !fuse_prepare) AMOFB_TRY(msd_prepare(ctx, p));
    if (fuse_prepare && !p->have_com) return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_msd_window needs amofb_msd_set_com first");
    int ap_ng = 1, ap_nwt = 1;
    const int force_ng = env_int("AMOFB_MSD_AP_NG", 0), force_wpg = env_int("AMOFB_MSD_AP_WPG", 0), force_nwt = env_int("AMOFB_MSD_AP_NWT", 0);
    if (ap_delta) {
        // groups of <= MSD_AP_NWT windows bound to warps, and the block size that wastes the fewest thread slots:
        // efficiency = (balance of the windows over the groups) x (fill of the last round of tasks)
        const int span = MSD_AP_KB * ap_delta, nsr = (p->T - 1) / span;
        const long long ntask = (long long)nsr * ap_delta, rem = (p->T - 1) - (long long)nsr * span;
        const int max_warps = MSD_AP_THREADS / 32;
        double best = -1.0;
        for (int nwt = 5; nwt <= MSD_AP_NWT_MAX; nwt += 2) {
            if (force_nwt > 0 && nwt != force_nwt) continue;
            for (int ng = 1; ng <= max_warps; ++ng) {
                if (force_ng > 0 && ng != force_ng) continue;
                if (ng > 1 && (ng - 1) * nwt >= n_window) break;          // a group without a single requested window
                const int passes = (n_window + ng * nwt - 1) / (ng * nwt);
                for (int wpg = 1; wpg * ng <= max_warps; ++wpg) {
                    if (force_wpg > 0 && wpg != force_wpg) continue;
                    const int gs = 32 * wpg;
                    const double slots = (double)((ntask + gs - 1) / gs) * MSD_AP_KB + (double)((rem + gs - 1) / gs);
                    const double fill = ((double)ntask * MSD_AP_KB + (double)rem) / (slots * gs);
                    const double balance = (double)n_window / ((double)passes * ng * nwt);
                    // resident warps hide the FP64 latency: below 3/4 of the warp budget the score drops in proportion
                    const double occ = std::min(1.0, (double)(wpg * ng) / (0.75 * max_warps));
                    // frame pairs formed per shared-memory read: below ~2.2 the shared-memory pipe, not FP64, sets the pace
                    const double ppl = (double)(MSD_AP_KB * nwt) / (double)(2 * MSD_AP_KB + nwt - 1);
                    const double eff = fill * balance * occ * std::min(1.0, ppl / 2.2);
                    if (eff > best + 1e-9) { best = eff; threads = 32 * wpg * ng; ap_ng = ng; ap_nwt = nwt; }
                }
            }
        }
    }
    const int nwarp = threads / 32;
    int *d_window = nullptr;
    double *d_partial = nullptr;
    size_t extra = sizeof(double) * ((size_t)S * n_window + (size_t)MSD_NW * (size_t)std::max(MSD_AP_THREADS / 32, nwarp) + 96);
    size_t smem_full = sizeof(double) * (3 * (size_t)p->T + 4) + extra;
    size_t budget = (size_t)ctx->max_smem_optin > 1024 ? (size_t)ctx->max_smem_optin - 1024 : 0;
    bool use_smem = smem_full <= budget && !env_int("AMOFB_MSD_NO_SMEM", 0);
    if (!use_smem && ap_delta) return amofb_fail(ctx, AMOFB_ERR_STATE, "internal: shared-memory budget of the tiled window kernel");
    size_t smem = use_smem ? smem_full : extra;
    if (extra > budget) return amofb_fail(ctx, AMOFB_ERR_ARG, "too many window lengths (%d) for one pass", n_window);
    int per_sm = 1;
    const void *ap_kernel = nullptr;
    if (ap_delta) {
        const int cm = p->diag_cell ? 2 : (p->fixed_cell ? 1 : 0);
#define AMOFB_AP_KERNEL(NWT_) (cm == 2 ? (const void *)k_msd_window_ap<MSD_AP_KB, NWT_, 2> : cm == 1 ? (const void *)k_msd_window_ap<MSD_AP_KB, NWT_, 1> \
                                        : (const void *)k_msd_window_ap<MSD_AP_KB, NWT_, 0>)
        switch (ap_nwt) {
            case 5: ap_kernel = AMOFB_AP_KERNEL(5); break;
            case 7: ap_kernel = AMOFB_AP_KERNEL(7); break;
            case 9: ap_kernel = AMOFB_AP_KERNEL(9); break;
            case 11: ap_kernel = AMOFB_AP_KERNEL(11); break;
            default: ap_kernel = AMOFB_AP_KERNEL(13); ap_nwt = 13; break;
        }
#undef AMOFB_AP_KERNEL
        CUDA_TRY(ctx, cudaFuncSetAttribute(ap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ap_kernel, threads, smem));
    } else if (use_smem) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(k_msd_window<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_msd_window<true>, threads, smem));
    } else {
        CUDA_TRY(ctx, cudaFuncSetAttribute(k_msd_window<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_msd_window<false>, threads, smem));
    }
    if (per_sm < 1) return amofb_fail(ctx, AMOFB_ERR_CUDA, "window kernel does not fit on an SM");
    int grid = std::max(1, std::min(p->n, ctx->num_sms * per_sm));
    int rc = AMOFB_OK;
    if ((rc = dev_alloc(ctx, &d_window, (size_t)n_window))) return rc;
    if ((rc = dev_alloc(ctx, &d_partial, (size_t)grid * S * n_window))) { pool_put(ctx, d_window); return rc; }
    std::vector<double> part((size_t)grid * S * n_window);
    cudaError_t e = cudaMemcpyAsync(d_window, window, sizeof(int) * n_window, cudaMemcpyHostToDevice, ctx->s_compute);
    if (e == cudaSuccess) {
        if (ap_delta) {
            const double *a_P = p->d_P; const uint8_t *a_sp = p->d_species;
            int a_n = p->n, a_T = p->T, a_S = S;
            const MsdGeom *a_geom = p->d_geom; const double *a_com = p->d_com;
            int a_prep = fuse_prepare ? 1 : 0;
            void *kargs[] = {(void *)&a_P, (void *)&a_sp, (void *)&a_n, (void *)&a_T, (void *)&ap_delta, (void *)&n_window, (void *)&ap_ng, (void *)&a_S, (void *)&d_partial,
                             (void *)&a_geom, (void *)&a_com, (void *)&a_prep};
            e = cudaLaunchKernel(ap_kernel, dim3(grid), dim3(threads), kargs, smem, ctx->s_compute);
        } else if (use_smem) k_msd_window<true><<<grid, threads, smem, ctx->s_compute>>>(p->d_P, p->d_species, p->n, p->T, d_window, n_window, S, d_partial);
        else k_msd_window<false><<<grid, threads, smem, ctx->s_compute>>>(p->d_P, p->d_species, p->n, p->T, d_window, n_window, S, d_partial);
        ctx->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(part.data(), d_partial, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, ctx->s_compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_compute);
    pool_put(ctx, d_window);
    pool_put(ctx, d_partial);
    CUDA_TRY(ctx, e);
    for (int i = 0; i < S * n_window; ++i) {
        double s = 0.0;
        for (int b = 0; b < grid; ++b) s += part[(size_t)b * S * n_window + i];   // fixed order: deterministic
        sums[i] = s;
    }
    return AMOFB_OK;
}


// ---- streaming path ---------------------------------------------------------------------------------------------
extern "C" int amofb_msd_slab_frames(amofb_ctx *ctx, int *frames) {
    MsdState *p = nullptr;
    AMOFB_TRY(msd_state(ctx, &p, "amofb_msd_slab_frames"));
    if (!frames) return amofb_fail(ctx, AMOFB_ERR_ARG, "null output");
    *frames = p->stage_frames;
    return AMOFB_OK;
}

// enqueue the copy (host slabs) and the mass sums of one slab; the result is fetched by msd_slab_sums_wait
static int msd_slab_sums_begin_impl(amofb_ctx *ctx, int first_frame, int count, const double *pos, bool on_device, int64_t frame_stride = 0) {
    MsdState *p = nullptr;
    AMOFB_TRY(msd_state(ctx, &p, "amofb_msd_slab_sums"));
    if (p->prepared || p->consumed) return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_msd_slab_sums after the positions were transformed");
    if (p->n_pend >= 2) return amofb_fail(ctx, AMOFB_ERR_STATE, "two slabs already await amofb_msd_slab_commit");
    if (count < 1 || !pos || first_frame != p->sums_next || first_frame + (long long)count > p->T)
        return amofb_fail(ctx, AMOFB_ERR_ARG, "slabs must cover the frames in order: expected first_frame %d, got [%d, %d) of %d", p->sums_next,
                          first_frame, first_frame + count, p->T);
    if (!on_device && count > p->stage_frames)
        return amofb_fail(ctx, AMOFB_ERR_ARG, "a host slab holds at most %d frames (amofb_msd_slab_frames)", p->stage_frames);
    p->soa = true;                    // from here on d_P is the prepared SoA store
    const size_t fr = 3 * (size_t)p->n;
    const double *src = pos;
    int sl = -1;
    nvtx_range rng("amofb_msd_slab_sums");
    if (!on_device) {
        sl = p->next_stage;
        p->next_stage = (p->next_stage + 1) % 3;
        CUDA_TRY(ctx, cudaEventSynchronize(p->ev_stage[sl]));   // the commit that last read this slot is done
        if (frame_stride > (int64_t)fr)      // the local atoms are a column block of wider frames: strided copy, no host-side packing
            CUDA_TRY(ctx, cudaMemcpy2DAsync(p->d_stage[sl], sizeof(double) * fr, pos, sizeof(double) * (size_t)frame_stride, sizeof(double) * fr, (size_t)count,
                                            cudaMemcpyHostToDevice, ctx->s_copy));
        else
            CUDA_TRY(ctx, cudaMemcpyAsync(p->d_stage[sl], pos, sizeof(double) * fr * count, cudaMemcpyHostToDevice, ctx->s_copy));
        cudaEvent_t ev;
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventRecord(ev, ctx->s_copy));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->s_compute, ev, 0));
        CUDA_TRY(ctx, cudaEventDestroy(ev));
        src = p->d_stage[sl];
    }
    MsdState::PendingSlab &q = p->pend[p->n_pend];
    const int groups = std::max(1, std::min(64, p->n / 4096));
    pool_put(ctx, q.d_partial); pool_put(ctx, q.d_out); pool_put(ctx, q.h_out);
    q.d_partial = q.d_out = q.h_out = nullptr;
    AMOFB_TRY(dev_alloc(ctx, &q.d_partial, (size_t)groups * count * 3));
    AMOFB_TRY(dev_alloc(ctx, &q.d_out, (size_t)count * 3));
    AMOFB_TRY(pinned_alloc(ctx, &q.h_out, (size_t)count * 3));
    if (!q.ev) CUDA_TRY(ctx, cudaEventCreateWithFlags(&q.ev, cudaEventDisableTiming));
    k_msd_slab_sums<<<dim3(count, groups), 256, 0, ctx->s_compute>>>(src, p->d_masses, p->n, count, q.d_partial);
    k_msd_sum_groups<<<std::max(1, std::min((count * 3 + 255) / 256, ctx->num_sms * 4)), 256, 0, ctx->s_compute>>>(q.d_partial, groups, count * 3, q.d_out);
    ctx->launches += 2;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(q.h_out, q.d_out, sizeof(double) * 3 * (size_t)count, cudaMemcpyDeviceToHost, ctx->s_compute));
    CUDA_TRY(ctx, cudaEventRecord(q.ev, ctx->s_compute));
    q.ptr = src; q.first = first_frame; q.count = count; q.slot = sl; q.waited = false;
    p->n_pend += 1;
    p->sums_next = first_frame + count;
    return AMOFB_OK;
}

// sums of the oldest slab whose sums have not been fetched yet
static int msd_slab_sums_wait_impl(amofb_ctx *ctx, double *sums) {
    MsdState *p = nullptr;
    AMOFB_TRY(msd_state(ctx, &p, "amofb_msd_slab_sums_wait"));
    if (!sums) return amofb_fail(ctx, AMOFB_ERR_ARG, "null output");
    for (int i = 0; i < p->n_pend; ++i) {
        MsdState::PendingSlab &q = p->pend[i];
        if (q.waited) continue;
        CUDA_TRY(ctx, cudaEventSynchronize(q.ev));
        for (int k = 0; k < q.count; ++k) {
            sums[4 * (size_t)k] = q.h_out[3 * (size_t)k];
            sums[4 * (size_t)k + 1] = q.h_out[3 * (size_t)k + 1];
            sums[4 * (size_t)k + 2] = q.h_out[3 * (size_t)k + 2];
            sums[4 * (size_t)k + 3] = p->mass_sum;
        }
        q.waited = true;
        return AMOFB_OK;
    }
    return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_msd_slab_sums_wait without a slab whose sums are outstanding");
}

extern "C" int amofb_msd_slab_sums_begin(amofb_ctx *ctx, int first_frame, int count, const double *pos) {
    return msd_slab_sums_begin_impl(ctx, first_frame, count, pos, false);
}
extern "C" int amofb_msd_slab_sums_begin_strided(amofb_ctx *ctx, int first_frame, int count, const double *pos, int64_t frame_stride) {
    if (ctx && frame_stride < 0) return amofb_fail(ctx, AMOFB_ERR_ARG, "negative frame stride");
    return msd_slab_sums_begin_impl(ctx, first_frame, count, pos, false, frame_stride);
}
extern "C" int amofb_msd_slab_sums_begin_device(amofb_ctx *ctx, int first_frame, int count, const double *pos_device) {
    return msd_slab_sums_begin_impl(ctx, first_frame, count, pos_device, true);
}
extern "C" int amofb_msd_slab_sums_wait(amofb_ctx *ctx, double *sums) { return msd_slab_sums_wait_impl(ctx, sums); }
extern "C" int amofb_msd_slab_sums(amofb_ctx *ctx, int first_frame, int count, const double *pos, double *sums) {
    AMOFB_TRY(msd_slab_sums_begin_impl(ctx, first_frame, count, pos, false));
    return msd_slab_sums_wait_impl(ctx, sums);
}
extern "C" int amofb_msd_slab_sums_device(amofb_ctx *ctx, int first_frame, int count, const double *pos_device, double *sums) {
    AMOFB_TRY(msd_slab_sums_begin_impl(ctx, first_frame, count, pos_device, true));
    return msd_slab_sums_wait_impl(ctx, sums);
}

extern "C" int amofb_msd_slab_commit(amofb_ctx *ctx, const double *com) {
    MsdState *p = nullptr;
    AMOFB_TRY(msd_state(ctx, &p, "amofb_msd_slab_commit"));
    if (p->n_pend < 1) return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_msd_slab_commit before amofb_msd_slab_sums");
    if (!p->pend[0].waited) return amofb_fail(ctx, AMOFB_ERR_STATE, "the sums of the oldest slab were never fetched");
    if (!com) return amofb_fail(ctx, AMOFB_ERR_ARG, "null centre of mass");
    nvtx_range rng("amofb_msd_slab_commit");
    MsdState::PendingSlab q = p->pend[0];
    const int first = q.first, count = q.count;
    double *d_com = p->d_com + 3 * (size_t)first;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_com, com, sizeof(double) * 3 * (size_t)count, cudaMemcpyHostToDevice, ctx->s_compute));   // pageable: staged before return
    const int cm = p->diag_cell ? 2 : (p->fixed_cell ? 1 : 0);
    {
        const double *a_slab = q.ptr; double *a_P = p->d_P; const MsdGeom *a_geom = p->d_geom; const double *a_com = d_com; double *a_carry = p->d_carry;
        int a_n = p->n, a_tp = p->Tp, a_first = first, a_count = count;
        void *kargs[] = {(void *)&a_slab, (void *)&a_P, (void *)&a_geom, (void *)&a_com, (void *)&a_carry, (void *)&a_n, (void *)&a_tp, (void *)&a_first, (void *)&a_count};
        {
            const void *kfn = cm == 2 ? (const void *)k_msd_slab_commit<2> : cm == 1 ? (const void *)k_msd_slab_commit<1> : (const void *)k_msd_slab_commit<0>;
            CUDA_TRY(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COMMIT_SMEM));
            CUDA_TRY(ctx, cudaLaunchKernel(kfn, dim3((p->n + COMMIT_A - 1) / COMMIT_A), dim3(COMMIT_THREADS), kargs, COMMIT_SMEM, ctx->s_compute));
        }
    }
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    if (q.slot >= 0) CUDA_TRY(ctx, cudaEventRecord(p->ev_stage[q.slot], ctx->s_compute));
    // pop the FIFO (the entries swap so that the buffers of the popped one are reused)
    p->pend[0] = p->pend[1];
    p->pend[1] = q;
    p->pend[1].ptr = nullptr;
    p->n_pend -= 1;
    p->slab_next = first + count;
    return AMOFB_OK;
}

// window sums from the prepared SoA store (lengths 0, D, 2D, ...): autocorrelation form first; a window whose result is
// small against the squares it was taken from (loss of digits) sends the whole request through the difference form
static int msd_window_soa(amofb_ctx *ctx, MsdState *p, int n_window, int ap_delta, double *sums) {
    const int S = p->S;
    int threads = MSD_SOA_THREADS, ap_ng = 1, ap_nwt = 13;
    {
        // groups of <= 13 window sums bound to warps: the shape that wastes the fewest thread slots
        const int span = MSD_SOA_KB * ap_delta, nsr = (p->T - 1) / span;
        const long long ntask = (long long)nsr * ap_delta, rem = (p->T - 1) - (long long)nsr * span;
        const int nwarp = MSD_SOA_THREADS / 32;
        const int force_ng = env_int("AMOFB_MSD_AP_NG", 0), force_nwt = env_int("AMOFB_MSD_AP_NWT", 0);
        double best = -1.0;
        for (int nwt = 5; nwt <= MSD_AP_NWT_MAX; nwt += 2) {
            if (force_nwt > 0 && nwt != force_nwt) continue;
            for (int ng = 1; ng <= nwarp; ng *= 2) {          // the warps split evenly over the groups
                if (force_ng > 0 && ng != force_ng) continue;
                if (ng > 1 && (ng - 1) * nwt >= n_window) break;
                const int passes = (n_window + ng * nwt - 1) / (ng * nwt);
                const int gs = 32 * (nwarp / ng);
                const double slots = (double)((ntask + gs - 1) / gs) * MSD_SOA_KB + (double)((rem + gs - 1) / gs);
                const double fill = ((double)ntask * MSD_SOA_KB + (double)rem) / (slots * gs);
                const double balance = (double)n_window / ((double)passes * ng * nwt);
                const double ppl = (double)(MSD_SOA_KB * nwt) / (double)(2 * MSD_SOA_KB + nwt - 1);      // pairs per shared-memory read
                const double eff = fill * balance * std::min(1.0, ppl / 2.6) / passes;                  // every pass stages the atoms again
                if (eff > best + 1e-9) { best = eff; ap_ng = ng; ap_nwt = nwt; }
            }
        }
    }
    if (env_int("AMOFB_MSD_DEBUG", 0)) fprintf(stderr, "[amofb msd] window kernel shape: %d sums per thread, %d groups, %d warps per group (%d threads)\n", ap_nwt, ap_ng, threads / 32 / ap_ng, threads);
    const size_t smem = sizeof(double) * ((size_t)p->Tp + 2 * (size_t)S * n_window + (size_t)MSD_AP_NWT_MAX * (MSD_SOA_THREADS / 32) + MSD_SOA_THREADS + 32 + 8);
    if (smem + 64 > (size_t)ctx->max_smem_optin) return amofb_fail(ctx, AMOFB_ERR_ARG, "%d frames and %d window lengths do not fit the shared-memory series buffer", p->T, n_window);
    // the autocorrelation form runs in the wide-tile kernel (two series buffers) whenever they fit; shape: the smallest
    // instantiated number of sums per thread covering the request (or 32 per pass), and the frames per tile that waste the fewest slots
    int wide_kb = 0, wide_nwt = 0, wide_rowcap = 0, wide_nbuf = 0, wide_threads = 256;
    size_t wide_smem = 0;
    if (!env_int("AMOFB_MSD_NO_WIDE", 0)) {
        wide_nwt = n_window <= 13 ? 13 : n_window <= 25 ? 25 : 32;
        if (int f = env_int("AMOFB_MSD_WIDE_NWT", 0)) wide_nwt = f == 13 || f == 25 ? f : 32;
        // (one block of 512 threads per SM with up to four series buffers was measured too: 4.8 against 4.2 ms per 100 000 atoms x 5 000 frames)
        double best = -1.0;
        for (int kb : {6, 8, 10}) {
            if (int f = env_int("AMOFB_MSD_WIDE_KB", 0)) if (kb != f) continue;
            const long long span = (long long)kb * ap_delta, nsr = (p->T - 2 + span) / span, ntask = nsr * ap_delta;
            const long long rounds = (ntask + wide_threads - 1) / wide_threads;
            const double fill = (double)(p->T - 1) / (double)(rounds * wide_threads * kb);
            const double ppl = (double)(kb * (wide_nwt - 1)) / (double)(kb + wide_nwt - 1);              // products per shared-memory read
            const double eff = fill * std::min(1.0, ppl / 6.0);
            if (eff > best + 1e-9) { best = eff; wide_kb = kb; wide_rowcap = (int)((std::max<long long>(p->Tp, nsr * span + 1) + 1) & ~1LL); }
        }
        // 512 threads: one block per SM; 256 threads: two, each with half of the shared memory
        const size_t fixed = sizeof(double) * (2 * (size_t)S * n_window + (size_t)wide_nwt * (wide_threads / 32) + wide_threads + 8);
        const size_t sm_total = (size_t)ctx->max_smem_optin + 1024;
        const size_t per_block = wide_threads == 512 ? (size_t)ctx->max_smem_optin : sm_total / 2 - 1024;
        const size_t room = per_block > fixed + 256 ? per_block - fixed - 256 : 0;
        wide_nbuf = wide_kb ? (int)std::min<size_t>(MSD_WIDE_MAX_BUF, room / (sizeof(double) * (size_t)wide_rowcap)) : 0;
        if (wide_nbuf < 2 && wide_threads == 256 && wide_kb) {          // two buffers do not fit twice: one block of 256 threads per SM
            const size_t room1 = (size_t)ctx->max_smem_optin > fixed + 256 ? (size_t)ctx->max_smem_optin - fixed - 256 : 0;
            wide_nbuf = (int)std::min<size_t>(MSD_WIDE_MAX_BUF, room1 / (sizeof(double) * (size_t)wide_rowcap));
        }
        if (int f = env_int("AMOFB_MSD_WIDE_NBUF", 0)) wide_nbuf = std::min(wide_nbuf, std::max(2, f));
        if (wide_nbuf < 2) wide_kb = 0;                       // not even two series buffers: the narrow kernel stages one at a time
        wide_smem = fixed + sizeof(double) * (size_t)wide_nbuf * wide_rowcap;
    }
    for (int form = env_int("AMOFB_MSD_NO_DOT", 0) ? 1 : 0; form < 2; ++form) {
        const bool dot = form == 0;
        const bool wide = dot && wide_kb > 0;
        const void *kfn = nullptr;
        if (wide) {
#define AMOFB_WIDE_KERNEL2(KB_, TH_) (wide_nwt == 13 ? (const void *)k_msd_window_wide<KB_, 13, TH_> : wide_nwt == 25 ? (const void *)k_msd_window_wide<KB_, 25, TH_> \
                                                     : (const void *)k_msd_window_wide<KB_, 32, TH_>)
#define AMOFB_WIDE_KERNEL(KB_) AMOFB_WIDE_KERNEL2(KB_, 256)
            kfn = wide_kb == 6 ? AMOFB_WIDE_KERNEL(6) : wide_kb == 8 ? AMOFB_WIDE_KERNEL(8) : AMOFB_WIDE_KERNEL(10);
#undef AMOFB_WIDE_KERNEL2
#undef AMOFB_WIDE_KERNEL
            if (env_int("AMOFB_MSD_DEBUG", 0)) fprintf(stderr, "[amofb msd] wide window kernel: %d threads, %d frames x %d sums per tile, %d series buffers of %d doubles, %zu bytes of shared memory\n", wide_threads, wide_kb, wide_nwt, wide_nbuf, wide_rowcap, wide_smem);
        } else {
#define AMOFB_SOA_KERNEL(NWT_) (dot ? (const void *)k_msd_window_soa<MSD_SOA_KB, NWT_, true> : (const void *)k_msd_window_soa<MSD_SOA_KB, NWT_, false>)
        switch (ap_nwt) {
            case 5: kfn = AMOFB_SOA_KERNEL(5); break;
            case 7: kfn = AMOFB_SOA_KERNEL(7); break;
            case 9: kfn = AMOFB_SOA_KERNEL(9); break;
            case 11: kfn = AMOFB_SOA_KERNEL(11); break;
            default: kfn = AMOFB_SOA_KERNEL(13); ap_nwt = 13; break;
        }
#undef AMOFB_SOA_KERNEL
        }
        const size_t smem_use = wide ? wide_smem : smem;
        if (wide) threads = wide_threads; else threads = MSD_SOA_THREADS;
        int per_sm = 0;
        CUDA_TRY(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_use));
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, threads, smem_use));
        if (per_sm < 1) return amofb_fail(ctx, AMOFB_ERR_CUDA, "window kernel does not fit on an SM");
        const int grid = std::max(1, std::min(p->n, ctx->num_sms * per_sm));
        double *d_partial = nullptr;
        int *d_perm = nullptr;
        AMOFB_TRY(dev_alloc(ctx, &d_partial, (size_t)grid * 2 * S * n_window));
        if (int rc2 = dev_alloc(ctx, &d_perm, (size_t)p->n)) { pool_put(ctx, d_partial); return rc2; }
        {
            // every block visits its range of atoms in species order (counting sort per range)
            std::vector<uint8_t> spec((size_t)p->n);
            std::vector<int> perm((size_t)p->n);
            cudaMemcpy(spec.data(), p->d_species, (size_t)p->n, cudaMemcpyDeviceToHost);
            const int per = (p->n + grid - 1) / grid;
            for (int lo = 0; lo < p->n; lo += per) {
                const int hi = std::min(p->n, lo + per);
                int at = lo;
                for (int sp = 0; sp < S; ++sp)
                    for (int a = lo; a < hi; ++a)
                        if (spec[a] == sp) perm[at++] = a;
            }
            cudaMemcpy(d_perm, perm.data(), sizeof(int) * (size_t)p->n, cudaMemcpyHostToDevice);
        }
        std::vector<double> part((size_t)grid * 2 * S * n_window);
        const double *a_P = p->d_P; const uint8_t *a_sp = p->d_species; const int *a_perm = d_perm;
        int a_n = p->n, a_T = p->T, a_tp = p->Tp, a_S = S, a_nw = n_window, a_delta = ap_delta;
        void *kargs[] = {(void *)&a_P, (void *)&a_sp, (void *)&a_perm, (void *)&a_n, (void *)&a_T, (void *)&a_tp, (void *)&a_delta, (void *)&a_nw, (void *)&ap_ng, (void *)&a_S, (void *)&d_partial};
        int *d_pieces = nullptr;
        if (wide) {
            // ranges of frames for the sums of squares, per pass: cut at m + 1 and T - m for every window length m of the pass, then
            // into pieces of at most c frames, with the smallest c that leaves no more pieces than threads
            const int npass = (n_window + wide_nwt - 1) / wide_nwt;
            std::vector<int> pieces((size_t)npass * (threads + 1));
            for (int ps = 0; ps < npass; ++ps) {
                std::vector<int> cut = {1, p->T};
                for (int w = ps * wide_nwt; w < std::min(n_window, (ps + 1) * wide_nwt); ++w) {
                    const long long m = (long long)w * ap_delta;
                    if (m + 1 > 1 && m + 1 < p->T) cut.push_back((int)(m + 1));
                    if (p->T - m > 1 && p->T - m < p->T) cut.push_back((int)(p->T - m));
                }
                std::sort(cut.begin(), cut.end());
                cut.erase(std::unique(cut.begin(), cut.end()), cut.end());
                int c = std::max(1, (p->T + threads - 1) / threads) | 1;      // odd: neighbouring threads read different banks
                for (;; c += 2) {
                    long long np = 0;
                    for (size_t i = 0; i + 1 < cut.size(); ++i) np += (cut[i + 1] - cut[i] + c - 1) / c;
                    if (np <= threads) break;
                }
                int *out = pieces.data() + (size_t)ps * (threads + 1);
                int at = 0;
                for (size_t i = 0; i + 1 < cut.size(); ++i)
                    for (int k = cut[i]; k < cut[i + 1]; k += c) out[at++] = k;
                for (; at <= threads; ++at) out[at] = p->T;
            }
            if (int rc2 = dev_alloc(ctx, &d_pieces, pieces.size())) { pool_put(ctx, d_partial); pool_put(ctx, d_perm); return rc2; }
            cudaMemcpy(d_pieces, pieces.data(), sizeof(int) * pieces.size(), cudaMemcpyHostToDevice);
        }
        const int *a_pieces = d_pieces;
        void *wargs[] = {(void *)&a_P, (void *)&a_sp, (void *)&a_perm, (void *)&a_pieces, (void *)&a_n, (void *)&a_T, (void *)&a_tp, (void *)&a_delta, (void *)&a_nw, (void *)&a_S, (void *)&wide_rowcap, (void *)&wide_nbuf, (void *)&d_partial};
        cudaError_t e = cudaLaunchKernel(kfn, dim3(grid), dim3(threads), wide ? wargs : kargs, smem_use, ctx->s_compute);
        ctx->launches += 1;
        if (e == cudaSuccess) e = cudaMemcpyAsync(part.data(), d_partial, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, ctx->s_compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_compute);
        pool_put(ctx, d_partial);
        pool_put(ctx, d_perm);
        pool_put(ctx, d_pieces);
        CUDA_TRY(ctx, e);
        bool redo = false;
        for (int i = 0; i < S * n_window; ++i) {
            double c = 0.0, ss = 0.0;
            for (int b = 0; b < grid; ++b) {                  // fixed order: deterministic
                c += part[(size_t)b * 2 * S * n_window + i];
                ss += part[(size_t)b * 2 * S * n_window + (size_t)S * n_window + i];
            }
            if (i % n_window == 0) { sums[i] = 0.0; continue; }            // m = 0: exactly zero (msd.py:197-204 with m = 0)
            if (!dot) { sums[i] = c; continue; }
            const double v = ss - 2.0 * c;
            // the cross term and the squares each carry ~1e-16 of ss: keep 1e-13 of the result
            if (ss > 500.0 * fabs(v)) redo = true;
            sums[i] = v;
        }
        if (!redo) break;
    }
    return AMOFB_OK;
}

extern "C" int amofb_msd_direct(amofb_ctx *ctx, double *sums) {
    MsdState *p = nullptr;
    AMOFB_TRY(msd_state(ctx, &p, "amofb_msd_direct"));
    if (p->prepared || p->consumed || p->soa) return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_msd_direct needs the untouched positions");
    if (!sums) return amofb_fail(ctx, AMOFB_ERR_ARG, "null output");
    for (int k = 0; k < p->T; ++k) {
        const double *c = p->cell.data() + 9 * (size_t)k;
        if (c[1] != 0.0 || c[2] != 0.0 || c[3] != 0.0 || c[5] != 0.0 || c[6] != 0.0 || c[7] != 0.0)
            return amofb_fail(ctx, AMOFB_ERR_GEOMETRY, "frame %d: DirectMsd only works for orthogonal cells", k);
    }
    k_msd_direct<<<(p->n + 127) / 128, 128, 0, ctx->s_compute>>>(p->d_P, p->d_geom, p->n, p->T);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    p->consumed = true;
    std::vector<uint8_t> spec((size_t)p->n);
    CUDA_TRY(ctx, cudaMemcpy(spec.data(), p->d_species, (size_t)p->n, cudaMemcpyDeviceToHost));
    std::vector<double> w((size_t)p->n);
    double *d_w = nullptr, *d_out = nullptr;
    AMOFB_TRY(dev_alloc(ctx, &d_w, (size_t)p->n));
    int rc = dev_alloc(ctx, &d_out, (size_t)p->T);
    for (int s = 0; s < p->S && rc == AMOFB_OK; ++s) {
        for (int i = 0; i < p->n; ++i) w[i] = spec[i] == s ? 1.0 : 0.0;
        cudaError_t e = cudaMemcpy(d_w, w.data(), sizeof(double) * p->n, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { rc = amofb_fail(ctx, AMOFB_ERR_CUDA, "msd_direct: %s", cudaGetErrorString(e)); break; }
        rc = msd_frame_sums<1>(ctx, p, d_w, d_out);
        if (rc) break;
        e = cudaMemcpy(sums + (size_t)s * p->T, d_out, sizeof(double) * p->T, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { rc = amofb_fail(ctx, AMOFB_ERR_CUDA, "msd_direct: %s", cudaGetErrorString(e)); break; }
    }
    pool_put(ctx, d_w);
    pool_put(ctx, d_out);
    return rc;
}

extern "C" int amofb_msd_get_positions(amofb_ctx *ctx, double *pos) {
    MsdState *p = nullptr;
    AMOFB_TRY(msd_state(ctx, &p, "amofb_msd_get_positions"));
    if (p->prepared || p->consumed || p->soa) return amofb_fail(ctx, AMOFB_ERR_STATE, "positions were already transformed in place");
    if (!pos) return amofb_fail(ctx, AMOFB_ERR_ARG, "null output");
    const size_t fr = 3 * (size_t)p->n;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->s_copy));
    for (int done = 0; done < p->T;) {
        int nf = std::min(p->stage_frames, p->T - done);
        dim3 grid((p->n + 31) / 32, (nf + 31) / 32);
        k_msd_untranspose<<<grid, 256, 0, ctx->s_compute>>>(p->d_P, p->d_stage[0], p->n, p->T, done, nf);
        ctx->launches += 1;
        CUDA_TRY(ctx, cudaGetLastError());
        CUDA_TRY(ctx, cudaMemcpyAsync(pos + fr * done, p->d_stage[0], sizeof(double) * fr * nf, cudaMemcpyDeviceToHost, ctx->s_compute));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->s_compute));
        done += nf;
    }
    return AMOFB_OK;
}

extern "C" int amofb_msd_end(amofb_ctx *ctx) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->msd) return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_msd_end before amofb_msd_begin");
    msd_release(ctx);
    return AMOFB_OK;
}
