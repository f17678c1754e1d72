// pfx_api.cu -- context management and the C ABI of libpfx (include/pfx.h).
//
// Execution model.  One candidate is scored by one warp (Bates groups) or by 8 lanes (Lyon
// statistics).  Every Bates kernel is persistent: the grid is sized to SM count x resident blocks
// and each warp pulls candidate indices from a global atomic counter, because the work per
// candidate is heavy-tailed (a Levenberg-Marquardt solve takes 30 to 1800 function evaluations).
// All per-candidate working data (profile, histogram, Jacobian, R factor) live in shared memory;
// HBM is touched once to read the candidate and once to write its feature row.
#include "pfx_internal.cuh"

using namespace pfx;


int pfx::set_err(pfx_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

extern "C" int pfx_version(void) { return PFX_VERSION; }

extern "C" int pfx_create(int device, pfx_ctx** out) {
    if (!out) return PFX_ERR_INVALID;
    *out = nullptr;
    pfx_ctx* ctx = (pfx_ctx*)calloc(1, sizeof(pfx_ctx));
    if (!ctx) return PFX_ERR_NOMEM;
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        free(ctx);
        return PFX_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (cudaMalloc(&ctx->counters, sizeof(u64) * N_SETS * COUNTERS_PER_SET) != cudaSuccess) {
        free(ctx);
        return PFX_ERR_NOMEM;
    }
    ctx->scratch_bytes = sizeof(u64) * N_SETS * COUNTERS_PER_SET;
    for (int i = 0; i < N_SETS; ++i) {
        ctx->sets[i].counters = ctx->counters + (size_t)i * COUNTERS_PER_SET;
        cudaEventCreateWithFlags(&ctx->sets[i].done, cudaEventDisableTiming);
    }
    for (int i = 0; i < N_EVENTS; ++i) cudaEventCreate(&ctx->ev[i]);
    for (int i = 0; i < 2; ++i) {
        cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&ctx->side2[i], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&ctx->join2_ev[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->fork_ev[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->join_ev[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->hist_ev[i], cudaEventDisableTiming);
        int lo_pr = 0, hi_pr = 0;
        cudaDeviceGetStreamPriorityRange(&lo_pr, &hi_pr);
        cudaStreamCreateWithPriority(&ctx->hi[i], cudaStreamNonBlocking, hi_pr);
        {
            // stream of the final-fit kernel when it runs next to the rounds kernel (PFX_PIPE_PRIO=0: lowest priority)
            const char* pe = getenv("PFX_PIPE_PRIO");
            cudaStreamCreateWithPriority(&ctx->pf[i], cudaStreamNonBlocking, (pe && atoi(pe) == 0) ? lo_pr : hi_pr);
        }
        cudaEventCreateWithFlags(&ctx->pipe_ev[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->pipe_done_ev[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->fill_ev[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->main_ev[i], cudaEventDisableTiming);
    }
    for (int i = 0; i < MAX_SUB; ++i) {
        cudaStreamCreateWithFlags(&ctx->sub[i], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&ctx->sub_ev[i], cudaEventDisableTiming);
    }
    // sub-batches of the histogram -> Gaussian chain per call (see score_dev_impl); PFX_SUBBATCH
    // overrides the default for experiments
    ctx->subbatch = 0;
    ctx->dm_exact = 1;
    ctx->gauss_hybrid = 0;
    ctx->gauss_wave = 0;
    if (const char* e = getenv("PFX_WAVE")) ctx->gauss_wave = atoi(e) != 0;
    ctx->hist_u8 = 1;
    if (const char* e = getenv("PFX_HIST_U8")) ctx->hist_u8 = atoi(e) != 0;
    if (const char* e = getenv("PFX_HYBRID")) ctx->gauss_hybrid = atoi(e) != 0;
    ctx->gauss_slots = 1;      // the rounds kernel in use; PFX_SLOTS=0 selects the lane kernel (or PFX_HYBRID / PFX_WAVE)
    if (const char* e = getenv("PFX_SLOTS")) ctx->gauss_slots = atoi(e) != 0;
    else if (getenv("PFX_HYBRID") || getenv("PFX_WAVE")) ctx->gauss_slots = 0;
    ctx->final_slots = 1;      // PFX_FINAL_SLOTS=0: the lane-pair final-fit kernel
    if (const char* e = getenv("PFX_FINAL_SLOTS")) ctx->final_slots = atoi(e) != 0;
    ctx->fill = 0;
    if (const char* e = getenv("PFX_FILL")) ctx->fill = atoi(e);
    ctx->sub_global = 1;
    ctx->sub_blocks = 0;
    if (const char* e = getenv("PFX_SUB_GLOBAL")) ctx->sub_global = atoi(e);
    if (const char* e = getenv("PFX_SUB_BLOCKS")) ctx->sub_blocks = atoi(e);
    ctx->sin_per_sm = 0;
    if (const char* e = getenv("PFX_SIN_PER_SM")) ctx->sin_per_sm = atoi(e);
    ctx->pipe = 1;             // PFX_PIPE=0: the final-fit kernel starts when the rounds kernel has ended
    if (const char* e = getenv("PFX_PIPE")) ctx->pipe = atoi(e) != 0;
    ctx->chain_on_hi = 0;
    ctx->chain_par = 0;
    ctx->lmpar_budget = 2;
    if (const char* e = getenv("PFX_LMPAR_BUDGET")) ctx->lmpar_budget = atoi(e);
    ctx->lmpar_budget8 = ctx->lmpar_budget;
    if (const char* e = getenv("PFX_LMPAR_BUDGET8")) ctx->lmpar_budget8 = atoi(e);
    ctx->fs_k = 0;
    if (const char* e = getenv("PFX_FS_K")) ctx->fs_k = atoi(e);
    ctx->pipe_hold = 8;
    if (const char* e = getenv("PFX_PIPE_HOLD")) ctx->pipe_hold = atoi(e) > 0 ? atoi(e) : 1;
    ctx->hy_k = 0;
    ctx->hy_warps = 8;
    ctx->hy_pad = 8;
    if (const char* e = getenv("PFX_HY_PAD")) ctx->hy_pad = atoi(e);
    if (const char* e = getenv("PFX_HY_K")) ctx->hy_k = atoi(e);
    if (const char* e = getenv("PFX_HY_WARPS")) ctx->hy_warps = atoi(e);
    if (const char* e = getenv("PFX_DM_EXACT")) ctx->dm_exact = atoi(e) != 0;
    if (const char* e = getenv("PFX_SUBBATCH")) ctx->subbatch = atoi(e);
    *out = ctx;
    return PFX_OK;
}

extern "C" void pfx_destroy(pfx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cudaFree(ctx->counters);
    for (int i = 0; i < N_SETS; ++i) {
        pfx_scratch_set& t = ctx->sets[i];
        pfx_buf* bufs[] = {&t.carry, &t.lane[0], &t.lane[1], &t.hist_lane, &t.ovf, &t.hist_warp, &t.wave_state, &t.wave_ws, &t.wave_list, &t.ready, &t.sub_scratch};
        for (pfx_buf* b : bufs)
            if (b->p) cudaFree(b->p);
        cudaEventDestroy(t.done);
    }
    for (int i = 0; i < 2; ++i) {
        if (ctx->stage[i]) cudaFree(ctx->stage[i]);
        cudaStreamDestroy(ctx->streams[i]);
        cudaStreamDestroy(ctx->side[i]);
        cudaStreamDestroy(ctx->side2[i]);
        cudaStreamDestroy(ctx->hi[i]);
        cudaStreamDestroy(ctx->pf[i]);
        cudaEventDestroy(ctx->pipe_ev[i]);
        cudaEventDestroy(ctx->pipe_done_ev[i]);
        cudaEventDestroy(ctx->fill_ev[i]);
        cudaEventDestroy(ctx->main_ev[i]);
        cudaEventDestroy(ctx->join2_ev[i]);
        cudaEventDestroy(ctx->fork_ev[i]);
        cudaEventDestroy(ctx->join_ev[i]);
        cudaEventDestroy(ctx->hist_ev[i]);
    }
    for (int i = 0; i < MAX_SUB; ++i) {
        cudaStreamDestroy(ctx->sub[i]);
        cudaEventDestroy(ctx->sub_ev[i]);
    }
    for (int i = 0; i < N_EVENTS; ++i) cudaEventDestroy(ctx->ev[i]);
    free(ctx);
}

extern "C" const char* pfx_last_error(const pfx_ctx* ctx) { return ctx ? ctx->err : "null context"; }

extern "C" int pfx_device_info(const pfx_ctx* ctx, int* sm_count, int* smem_per_block, size_t* scratch_bytes) {
    if (!ctx) return PFX_ERR_INVALID;
    if (sm_count) *sm_count = ctx->sm_count;
    if (smem_per_block) *smem_per_block = ctx->smem_optin;
    if (scratch_bytes) *scratch_bytes = ctx->scratch_bytes;
    return PFX_OK;
}

// ---- launch helpers ------------------------------------------------------------------------
static int check_shape(pfx_ctx* ctx, const pfx_shape* s, int64_t n) {
    if (!ctx) return PFX_ERR_INVALID;
    if (!s) return set_err(ctx, PFX_ERR_INVALID, "shape is NULL");
    if (n < 0) return set_err(ctx, PFX_ERR_INVALID, "n < 0");
    if (s->dtype < PFX_U8 || s->dtype > PFX_F64 || s->dm_dtype < PFX_U8 || s->dm_dtype > PFX_F64 ||
        s->dmstat_dtype < PFX_U8 || s->dmstat_dtype > PFX_F64)
        return set_err(ctx, PFX_ERR_INVALID, "bad dtype");
    if (s->family != PFX_FAMILY_PHCX && s->family != PFX_FAMILY_PFD)
        return set_err(ctx, PFX_ERR_INVALID, "bad family");
    return PFX_OK;
}

// Everything the Bates kernels refuse, checked before anything is launched: an error in the middle
// of pfx_score_dev would leave kernels of the side streams in flight on buffers the caller may free.
static int check_bates_shape(pfx_ctx* ctx, const pfx_shape* s) {
    if (s->nbin < 8 || s->nbin > 65535)
        return set_err(ctx, PFX_ERR_UNSUPPORTED, "nbin = %d: the Bates scores need 8 <= nbin <= 65535", s->nbin);
    if (s->nsub < 1) return set_err(ctx, PFX_ERR_UNSUPPORTED, "nsub < 1");
    if (s->n_dm < 4) return set_err(ctx, PFX_ERR_UNSUPPORTED, "n_dm < 4");
    const size_t sub_bytes = (size_t)align_up(sub_work_doubles(s->nbin, s->nsub), 2) * 8;
    if (sub_bytes > (size_t)ctx->smem_optin)
        return set_err(ctx, PFX_ERR_UNSUPPORTED, "sub-bands %d x %d need %zu bytes of shared memory per candidate, limit %d",
                       s->nsub, s->nbin, sub_bytes, ctx->smem_optin);
    return PFX_OK;
}

pfx_scratch_set* pfx::take_set(pfx_ctx* ctx, cudaStream_t st) {
    pfx_scratch_set* set = &ctx->sets[ctx->set_rr];
    ctx->set_rr = (ctx->set_rr + 1) % N_SETS;
    if (set->used) cudaStreamWaitEvent(st, set->done, 0);      // behind the set's previous user, on the device
    set->used = true;
    return set;
}

void pfx::release_set(pfx_ctx*, pfx_scratch_set* set, cudaStream_t st) { cudaEventRecord(set->done, st); }

int pfx::ensure_buf(pfx_ctx* ctx, pfx_buf* b, size_t need) {
    if (b->bytes >= need) return PFX_OK;
    if (b->p) {
        // the old block may be in use by a call in flight: freeing it waits for the device
        cudaDeviceSynchronize();
        cudaFree(b->p);
        ctx->scratch_bytes -= b->bytes;
        b->p = nullptr;
        b->bytes = 0;
    }
    need = (need + 255) / 256 * 256;
    PFX_CUDA(ctx, cudaMalloc(&b->p, need));
    b->bytes = need;
    ctx->scratch_bytes += need;
    return PFX_OK;
}

int pfx::persistent_grid_impl(pfx_ctx* ctx, const void* fn, int threads, size_t smem, int* grid) {
    for (int i = 0; i < ctx->n_grid_cache; ++i) {
        const pfx_grid_cache_entry& e = ctx->grid_cache[i];
        if (e.fn == fn && e.threads == threads && e.smem == smem) {
            *grid = e.grid;
            return PFX_OK;
        }
    }
    PFX_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PFX_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
    if (per_sm < 1)
        return set_err(ctx, PFX_ERR_UNSUPPORTED,
                       "kernel needs %zu bytes of shared memory per block, limit %d", smem, ctx->smem_optin);
    *grid = per_sm * ctx->sm_count;
    // the attribute is per function, not per launch: the cache keeps the last shared-memory size
    // set for a function valid by dropping entries of the same function with another size
    int w = 0;
    for (int i = 0; i < ctx->n_grid_cache; ++i)
        if (ctx->grid_cache[i].fn != fn) ctx->grid_cache[w++] = ctx->grid_cache[i];
    ctx->n_grid_cache = w;
    if (ctx->n_grid_cache < (int)(sizeof(ctx->grid_cache) / sizeof(ctx->grid_cache[0]))) {
        pfx_grid_cache_entry& e = ctx->grid_cache[ctx->n_grid_cache++];
        e.fn = fn;
        e.threads = threads;
        e.smem = smem;
        e.grid = *grid;
    }
    return PFX_OK;
}

// Number of sub-batches the Gaussian chain of a call is cut into (pfx_set_subbatches).  The idea: the
// persistent lane kernels end in a tail set by the slowest candidates, and with S sub-batches on S
// streams the blocks freed by the tail of one sub-batch could be taken by the body of the next.
// Measured on B200 (profiles/r02_subbatch_sweep.txt): it LOSES -- 727 / 787 / 898 / 968 ms per 100 k
// candidates for S = 1 / 2 / 3 / 4 -- because a block leaves only when its last lane is done, so every
// sub-batch pays a full tail of its own and its successor starts block by block.  Default: off.
static int pick_subbatches(const pfx_ctx* ctx, int64_t n) {
    int s = ctx->subbatch > 0 ? ctx->subbatch : 1;
    if (s > MAX_SUB) s = MAX_SUB;
    while (s > 1 && n / s < 16384) --s;
    return s;
}

// =============================================================================================
// device-pointer entry points
// =============================================================================================
extern "C" int pfx_lyon8_dev(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, const void* profile,
                             const void* dmstat, double* feat, int64_t feat_ld, void* stream) {
    int rc = check_shape(ctx, shape, n);
    if (rc) return rc;
    if (n == 0) return PFX_OK;
    if (!profile || !dmstat || !feat || feat_ld < PFX_NLYON) return set_err(ctx, PFX_ERR_INVALID, "bad argument");
    PFX_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_lyon(ctx, shape, n, profile, dmstat, feat, feat_ld, (cudaStream_t)stream);
}

extern "C" int pfx_sinusoid_dev(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, const void* profile,
                                double* feat, int64_t feat_ld, int8_t* ier, int32_t* status, void* stream) {
    (void)status;
    int rc = check_shape(ctx, shape, n);
    if (rc) return rc;
    if (n == 0) return PFX_OK;
    if (!profile || !feat || feat_ld < 4) return set_err(ctx, PFX_ERR_INVALID, "bad argument");
    PFX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    pfx_scratch_set* set = take_set(ctx, st);
    rc = launch_sinusoid(ctx, shape, n, profile, feat, feat_ld, ier, st, set, 0);
    release_set(ctx, set, st);
    return rc;
}

extern "C" int pfx_gaussian_dev(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, const void* profile,
                                double* feat, int64_t feat_ld, int8_t* ier, int32_t* status, void* stream) {
    int rc = check_shape(ctx, shape, n);
    if (rc) return rc;
    if (n == 0) return PFX_OK;
    if (!profile || !feat || !status || feat_ld < 11) return set_err(ctx, PFX_ERR_INVALID, "bad argument");
    PFX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    pfx_scratch_set* set = take_set(ctx, st);
    rc = launch_hist(ctx, shape, n, profile, feat, feat_ld, ier, status, st, set, 0);
    if (!rc) rc = launch_gauss(ctx, shape, n, profile, feat, feat_ld, ier, status, st, set, 1);
    release_set(ctx, set, st);
    return rc;
}

extern "C" int pfx_dmfit_dev(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, const void* profile,
                             const void* dmcurve, const double* params, double* feat, int64_t feat_ld,
                             int8_t* ier, int32_t* status, void* stream) {
    (void)profile;
    int rc = check_shape(ctx, shape, n);
    if (rc) return rc;
    if (n == 0) return PFX_OK;
    if (!dmcurve || !params || !feat || !status || feat_ld < 19) return set_err(ctx, PFX_ERR_INVALID, "bad argument");
    PFX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    pfx_scratch_set* set = take_set(ctx, st);
    rc = launch_dmfit(ctx, shape, n, dmcurve, params, feat, feat_ld, ier, status, st, set, 0);
    release_set(ctx, set, st);
    return rc;
}

extern "C" int pfx_pfd_params_dev(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, const void* profile,
                                  double* params, void* stream) {
    int rc = check_shape(ctx, shape, n);
    if (rc) return rc;
    if (n == 0) return PFX_OK;
    if (!profile || !params) return set_err(ctx, PFX_ERR_INVALID, "bad argument");
    PFX_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_pfd_params(ctx, shape, n, profile, params, (cudaStream_t)stream);
}

extern "C" int pfx_subband_dev(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, const void* profile,
                               const void* subbands, const double* params, double* feat, int64_t feat_ld,
                               int32_t* status, void* stream) {
    int rc = check_shape(ctx, shape, n);
    if (rc) return rc;
    if (n == 0) return PFX_OK;
    if (!profile || !subbands || !params || !feat || !status || feat_ld < 22)
        return set_err(ctx, PFX_ERR_INVALID, "bad argument");
    PFX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    pfx_scratch_set* set = take_set(ctx, st);
    rc = launch_subband(ctx, shape, n, profile, subbands, params, feat, feat_ld, status, st, set, 0);
    release_set(ctx, set, st);
    return rc;
}

// The histogram -> Gaussian chain of one call: the histogram stage over the whole batch on `st`,
// then the two Gaussian kernels per sub-batch, each sub-batch on its own stream with its own scratch
// set, joined back into `st`.  Returns the first error; the streams are joined in any case.
static int launch_hist_gauss(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, const void* profile, double* feat,
                             int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, pfx_scratch_set* set,
                             cudaEvent_t after_hist, int* launches) {
    int rc = launch_hist(ctx, shape, n, profile, feat, ld, ier, status, st, set, 1);
    if (rc) return rc;
    *launches += 2;
    if (after_hist) cudaEventRecord(after_hist, st);
    const int nsub = pick_subbatches(ctx, n);
    if (nsub <= 1) {
        const int rcg = launch_gauss(ctx, shape, n, profile, feat, ld, ier, status, st, set, 2);
        *launches += ctx->gauss_launches;
        return rcg;
    }
    const int par = ctx->fork_rr & 1;
    ctx->chain_on_hi = 0;        // (sub-batches run on streams of their own: no hand-over between the Gaussian kernels)
    cudaEventRecord(ctx->hist_ev[par], st);
    const size_t es = elem_size(shape->dtype);
    for (int k = 0; k < nsub; ++k) {
        const int64_t lo = n * k / nsub, hi = n * (k + 1) / nsub;
        cudaStream_t sk = ctx->sub[k];
        cudaStreamWaitEvent(sk, ctx->hist_ev[par], 0);
        pfx_scratch_set* sset = take_set(ctx, sk);
        if (!rc)
            rc = launch_gauss(ctx, shape, hi - lo, (const char*)profile + (size_t)lo * shape->nbin * es,
                              feat + lo * ld, ld, ier ? ier + lo * PFX_NIER : nullptr, status + lo, sk, sset, 2);
        release_set(ctx, sset, sk);
        cudaEventRecord(ctx->sub_ev[k], sk);
        cudaStreamWaitEvent(st, ctx->sub_ev[k], 0);
        *launches += 2;
    }
    return rc;
}

static int score_dev_impl(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, uint32_t features,
                          const void* profile, const void* subbands, const void* dmcurve,
                          const void* dmstat, const double* params, double* feat, int64_t feat_ld,
                          int8_t* ier, int32_t* status, cudaStream_t st, int timing) {
    // timing 1: stages serialised on one stream and timed each; 2: streams as in production, the
    // events mark when each group completed since the start of the call (overlap trace)
    const bool timed = timing == 1, trace = timing == 2;
    const bool bates = (features & PFX_FEAT_BATES) != 0, lyon = (features & PFX_FEAT_LYON) != 0;
    const int ncol = (bates ? PFX_NBATES : 0) + (lyon ? PFX_NLYON : 0);
    if (ncol == 0) return set_err(ctx, PFX_ERR_INVALID, "no feature set selected");
    if (!profile || !feat || feat_ld < ncol) return set_err(ctx, PFX_ERR_INVALID, "bad argument");
    if (bates && (!subbands || !dmcurve || !params || !status)) return set_err(ctx, PFX_ERR_INVALID, "Bates scores need subbands, dmcurve, params and status");
    if (lyon && !dmstat) return set_err(ctx, PFX_ERR_INVALID, "Lyon statistics need dmstat");
    if (bates) {
        const int rcs = check_bates_shape(ctx, shape);
        if (rcs) return rcs;
    }
    int launches = 0;
    int e = 0;
    if (timed || trace) cudaEventRecord(ctx->ev[e++], st);
    int rc;
    if (lyon) {
        rc = launch_lyon(ctx, shape, n, profile, dmstat, feat + (bates ? PFX_NBATES : 0), feat_ld, st);
        if (rc) return rc;
        ++launches;
    }
    if (timed || trace) cudaEventRecord(ctx->ev[e++], st);
    if (bates) {
        PFX_CUDA(ctx, cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)n, st));
        if (ier) PFX_CUDA(ctx, cudaMemsetAsync(ier, 0, (size_t)n * PFX_NIER, st));
        rc = launch_fill_nan(ctx, feat, n, feat_ld, PFX_NBATES, st);
        if (rc) return rc;
        ++launches;
        pfx_scratch_set* set = take_set(ctx, st);
        ctx->fork_rr++;
        const int par = ctx->fork_rr & 1;
        if (!timed) {
            // The histogram -> Gaussian chain is the long pole and ends in a tail of slow candidates;
            // scores 1-4, 12-19 and 20-22 do not depend on it and run on side streams next to it,
            // launched first: the DM and sinusoid kernels have a short body and a long tail of slow
            // candidates (a few lingering blocks) that hides behind the body of the chain.
            // Whatever happens, the side streams are joined again before this function returns: the
            // caller's stream must cover every kernel that touches the caller's buffers.
            cudaStream_t sd = ctx->side[par], sd2 = ctx->side2[par];
            rc = PFX_OK;
            if (cudaEventRecord(ctx->fork_ev[par], st) != cudaSuccess ||
                cudaStreamWaitEvent(sd, ctx->fork_ev[par], 0) != cudaSuccess ||
                cudaStreamWaitEvent(sd2, ctx->fork_ev[par], 0) != cudaSuccess)
                rc = set_err(ctx, PFX_ERR_CUDA, "fork of the side streams failed: %s", cudaGetErrorString(cudaGetLastError()));
            if (!rc) rc = launch_dmfit(ctx, shape, n, dmcurve, params, feat, feat_ld, ier, status, sd, set, 3);
            if (!rc && trace) cudaEventRecord(ctx->ev[5], sd);
            if ((ctx->fill || (ctx->pipe && !ctx->pipe_suspend)) && !trace) {
                // The chain runs on a high-priority stream and the sinusoid kernel, cut into short blocks, is
                // released when the histogram stage is done: the block scheduler serves the chain's kernels
                // first, so the sinusoid blocks land on the SMs the tails of the Gaussian kernels leave idle.
                cudaStream_t hs = ctx->hi[par];
                cudaStreamWaitEvent(hs, ctx->fork_ev[par], 0);
                if (!rc) rc = launch_subband(ctx, shape, n, profile, subbands, params, feat, feat_ld, status, sd, set, 4);
                if (!ctx->fill && !rc) rc = launch_sinusoid(ctx, shape, n, profile, feat, feat_ld, ier, sd2, set, 0);
                ctx->chain_on_hi = 1;
                ctx->chain_par = par;
                if (!rc) rc = launch_hist_gauss(ctx, shape, n, profile, feat, feat_ld, ier, status, hs, set, ctx->fill_ev[par], &launches);
                ctx->chain_on_hi = 0;
                cudaEventRecord(ctx->main_ev[par], hs);
                if (ctx->fill) cudaStreamWaitEvent(sd2, ctx->fill_ev[par], 0);
                if (ctx->fill && !rc) rc = launch_sinusoid(ctx, shape, n, profile, feat, feat_ld, ier, sd2, set, 0, ctx->fill > 1 ? ctx->fill : 64);
                cudaStreamWaitEvent(st, ctx->main_ev[par], 0);
            } else {
            if (!rc) rc = launch_sinusoid(ctx, shape, n, profile, feat, feat_ld, ier, sd2, set, 0);
            if (!rc && trace) cudaEventRecord(ctx->ev[2], sd2);
            if (!rc) rc = launch_subband(ctx, shape, n, profile, subbands, params, feat, feat_ld, status, sd, set, 4);
            if (!rc) rc = launch_hist_gauss(ctx, shape, n, profile, feat, feat_ld, ier, status, st, set,
                                            trace ? ctx->ev[3] : nullptr, &launches);
            if (!rc && trace) cudaEventRecord(ctx->ev[4], st);
            }
            cudaEventRecord(ctx->join_ev[par], sd);
            cudaStreamWaitEvent(st, ctx->join_ev[par], 0);
            cudaEventRecord(ctx->join2_ev[par], sd2);
            cudaStreamWaitEvent(st, ctx->join2_ev[par], 0);
            release_set(ctx, set, st);
            if (rc) return rc;
            PFX_CUDA(ctx, cudaGetLastError());
            launches += 3;
        } else {
            rc = launch_sinusoid(ctx, shape, n, profile, feat, feat_ld, ier, st, set, 0);
            if (!rc) cudaEventRecord(ctx->ev[e++], st);
            if (!rc) rc = launch_hist_gauss(ctx, shape, n, profile, feat, feat_ld, ier, status, st, set, ctx->ev[e++], &launches);
            if (!rc) cudaEventRecord(ctx->ev[e++], st);
            if (!rc) rc = launch_dmfit(ctx, shape, n, dmcurve, params, feat, feat_ld, ier, status, st, set, 3);
            if (!rc) cudaEventRecord(ctx->ev[e++], st);
            if (!rc) rc = launch_subband(ctx, shape, n, profile, subbands, params, feat, feat_ld, status, st, set, 4);
            if (!rc) cudaEventRecord(ctx->ev[e++], st);
            release_set(ctx, set, st);
            if (rc) return rc;
            launches += 3;
        }
    }
    ctx->last_launches = launches;
    if (trace && bates) cudaEventRecord(ctx->ev[6], st);
    if (timed || trace) ctx->ev_valid = bates;
    ctx->ev_trace = trace;
    return PFX_OK;
}

extern "C" int pfx_score_dev(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, uint32_t features,
                             const void* profile, const void* subbands, const void* dmcurve,
                             const void* dmstat, const double* params, double* feat, int64_t feat_ld,
                             int8_t* ier, int32_t* status, void* stream) {
    int rc = check_shape(ctx, shape, n);
    if (rc) return rc;
    if (n == 0) return PFX_OK;
    PFX_CUDA(ctx, cudaSetDevice(ctx->device));
    return score_dev_impl(ctx, shape, n, features, profile, subbands, dmcurve, dmstat, params, feat, feat_ld,
                          ier, status, (cudaStream_t)stream, ctx->stage_timing);
}

extern "C" int pfx_reserve(pfx_ctx* ctx, const pfx_shape* shape, int64_t n_max) {
    int rc = check_shape(ctx, shape, n_max);
    if (rc) return rc;
    if (n_max == 0) return PFX_OK;
    PFX_CUDA(ctx, cudaSetDevice(ctx->device));
    rc = check_bates_shape(ctx, shape);
    if (rc) return rc;
    // every set is sized for a whole call of n_max candidates, so whichever set a call or one of
    // its sub-batches draws is large enough
    for (int i = 0; i < N_SETS; ++i) {
        rc = reserve_hist(ctx, &ctx->sets[i], shape, n_max);
        if (!rc) rc = reserve_gauss(ctx, &ctx->sets[i], shape, n_max);
        if (!rc) rc = reserve_subband(ctx, &ctx->sets[i], shape);
        if (rc) return rc;
    }
    ctx->reserved = true;
    return PFX_OK;
}

extern "C" int pfx_set_dm_exact(pfx_ctx* ctx, int on) {
    if (!ctx) return PFX_ERR_INVALID;
    ctx->dm_exact = on != 0;
    return PFX_OK;
}

extern "C" int pfx_set_subbatches(pfx_ctx* ctx, int nsub) {
    if (!ctx || nsub < 0) return PFX_ERR_INVALID;
    ctx->subbatch = nsub;
    return PFX_OK;
}

extern "C" int pfx_set_stage_timing(pfx_ctx* ctx, int on) {
    if (!ctx) return PFX_ERR_INVALID;
    ctx->stage_timing = on;
    return PFX_OK;
}

extern "C" int pfx_last_timing(pfx_ctx* ctx, pfx_timing* out) {
    if (!ctx || !out) return PFX_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    out->launches = ctx->last_launches;
    if (!ctx->ev_valid) return PFX_OK;
    PFX_CUDA(ctx, cudaSetDevice(ctx->device));
    PFX_CUDA(ctx, cudaEventSynchronize(ctx->ev[6]));
    if (ctx->ev_trace) {
        // completion times since the start of the call, streams as in production
        cudaEventElapsedTime(&out->lyon, ctx->ev[0], ctx->ev[1]);
        cudaEventElapsedTime(&out->sinusoid, ctx->ev[0], ctx->ev[2]);
        cudaEventElapsedTime(&out->histgauss, ctx->ev[0], ctx->ev[3]);
        cudaEventElapsedTime(&out->gauss, ctx->ev[0], ctx->ev[4]);
        cudaEventElapsedTime(&out->dmfit, ctx->ev[0], ctx->ev[5]);
        out->subband = 0.f;
        cudaEventElapsedTime(&out->total, ctx->ev[0], ctx->ev[6]);
        return PFX_OK;
    }
    cudaEventElapsedTime(&out->lyon, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&out->sinusoid, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&out->histgauss, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&out->gauss, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&out->dmfit, ctx->ev[4], ctx->ev[5]);
    cudaEventElapsedTime(&out->subband, ctx->ev[5], ctx->ev[6]);
    cudaEventElapsedTime(&out->total, ctx->ev[0], ctx->ev[6]);
    return PFX_OK;
}

// =============================================================================================
// host-pointer entry point: chunked, double-buffered H2D -> kernels -> D2H
// =============================================================================================
namespace {
struct StageLayout {
    size_t profile, subbands, dmcurve, dmstat, params, feat, ier, status, total;
};
inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }
StageLayout stage_layout(const pfx_shape* s, int64_t chunk, bool bates, bool lyon, int ncol) {
    StageLayout L;
    size_t off = 0;
    L.profile = off; off += up256((size_t)chunk * s->nbin * elem_size(s->dtype));
    L.subbands = off; if (bates) off += up256((size_t)chunk * s->nsub * s->nbin * elem_size(s->dtype));
    L.dmcurve = off; if (bates) off += up256((size_t)chunk * s->n_dm * elem_size(s->dm_dtype));
    L.dmstat = off; if (lyon) off += up256((size_t)chunk * s->n_dmstat * elem_size(s->dmstat_dtype));
    L.params = off; if (bates) off += up256((size_t)chunk * PFX_NPARAM * 8);
    L.feat = off; off += up256((size_t)chunk * ncol * 8);
    L.ier = off; if (bates) off += up256((size_t)chunk * PFX_NIER);
    L.status = off; if (bates) off += up256((size_t)chunk * 4);
    L.total = off;
    return L;
}
}  // namespace

extern "C" int pfx_score_host(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, uint32_t features,
                              const void* profile, const void* subbands, const void* dmcurve,
                              const void* dmstat, const double* params, double* feat, int64_t feat_ld,
                              int8_t* ier, int32_t* status, int64_t chunk) {
    int rc = check_shape(ctx, shape, n);
    if (rc) return rc;
    if (n == 0) return PFX_OK;
    const bool bates = (features & PFX_FEAT_BATES) != 0, lyon = (features & PFX_FEAT_LYON) != 0;
    const int ncol = (bates ? PFX_NBATES : 0) + (lyon ? PFX_NLYON : 0);
    if (ncol == 0) return set_err(ctx, PFX_ERR_INVALID, "no feature set selected");
    if (!profile || !feat || feat_ld < ncol) return set_err(ctx, PFX_ERR_INVALID, "bad argument");
    if (bates && (!subbands || !dmcurve || !params || !status)) return set_err(ctx, PFX_ERR_INVALID, "Bates scores need subbands, dmcurve, params and status");
    if (lyon && !dmstat) return set_err(ctx, PFX_ERR_INVALID, "Lyon statistics need dmstat");
    PFX_CUDA(ctx, cudaSetDevice(ctx->device));
    // Bates scoring is compute bound (about 5 us of GPU time per candidate against 2.7 KB of
    // input), so copies need no overlap and every chunk pays one tail of slow candidates: use few,
    // large chunks (end to end on 2 Mi candidates: 226 k / 240 k / 244 k candidates/s with chunks of
    // 256 Ki / 512 Ki / 1 Mi; staging 2 x 3.2 GB).  Lyon-only batches are PCIe bound: 1 Mi candidates
    // per chunk, double buffered.
    if (chunk <= 0) {
        chunk = 1 << 20;
        if (const char* e = getenv("PFX_CHUNK")) chunk = atoll(e) > 0 ? atoll(e) : chunk;      // (development)
    }
    if (chunk > n) chunk = n;
    const StageLayout L = stage_layout(shape, chunk, bates, lyon, ncol);
    for (int b = 0; b < 2; ++b) {
        pfx_buf sb = {ctx->stage[b], ctx->stage_bytes[b]};
        rc = ensure_buf(ctx, &sb, L.total);
        ctx->stage[b] = sb.p;
        ctx->stage_bytes[b] = sb.bytes;
        if (rc) return rc;
    }
    const size_t es = elem_size(shape->dtype), eds = elem_size(shape->dm_dtype), ess = elem_size(shape->dmstat_dtype);
    int launches = 0;
    int64_t done = 0;
    int b = 0;
    // every exit of the loop goes through the two synchronisations below: nothing may still be
    // copying into the caller's arrays when this function returns, error or not
#define PFX_TRY(call)                                                                                  \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            rc = set_err(ctx, PFX_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),    \
                         __FILE__, __LINE__);                                                          \
            goto out;                                                                                  \
        }                                                                                              \
    } while (0)
    while (done < n) {
        const int64_t cn = (n - done < chunk) ? n - done : chunk;
        cudaStream_t st = ctx->streams[b];
        char* base = (char*)ctx->stage[b];
        PFX_TRY(cudaMemcpyAsync(base + L.profile, (const char*)profile + (size_t)done * shape->nbin * es,
                                (size_t)cn * shape->nbin * es, cudaMemcpyHostToDevice, st));
        if (bates) {
            PFX_TRY(cudaMemcpyAsync(base + L.subbands,
                                    (const char*)subbands + (size_t)done * shape->nsub * shape->nbin * es,
                                    (size_t)cn * shape->nsub * shape->nbin * es, cudaMemcpyHostToDevice, st));
            PFX_TRY(cudaMemcpyAsync(base + L.dmcurve, (const char*)dmcurve + (size_t)done * shape->n_dm * eds,
                                    (size_t)cn * shape->n_dm * eds, cudaMemcpyHostToDevice, st));
            PFX_TRY(cudaMemcpyAsync(base + L.params, (const char*)(params + (size_t)done * PFX_NPARAM),
                                    (size_t)cn * PFX_NPARAM * 8, cudaMemcpyHostToDevice, st));
        }
        if (lyon)
            PFX_TRY(cudaMemcpyAsync(base + L.dmstat, (const char*)dmstat + (size_t)done * shape->n_dmstat * ess,
                                    (size_t)cn * shape->n_dmstat * ess, cudaMemcpyHostToDevice, st));
        // (several chunks in flight: the hand-over between the Gaussian kernels is left out -- with the next chunk's
        // kernels arriving behind it, it measured 3 % slower end to end; a batch of one chunk keeps it)
        ctx->pipe_suspend = n > chunk;
        rc = score_dev_impl(ctx, shape, cn, features, base + L.profile, base + L.subbands, base + L.dmcurve,
                            base + L.dmstat, (const double*)(base + L.params), (double*)(base + L.feat), ncol,
                            bates && ier ? (int8_t*)(base + L.ier) : nullptr, (int32_t*)(base + L.status), st, 0);
        ctx->pipe_suspend = 0;
        if (rc) goto out;
        launches += ctx->last_launches;
        if (feat_ld == ncol) {
            PFX_TRY(cudaMemcpyAsync(feat + (size_t)done * feat_ld, base + L.feat, (size_t)cn * ncol * 8,
                                    cudaMemcpyDeviceToHost, st));
        } else {
            PFX_TRY(cudaMemcpy2DAsync(feat + (size_t)done * feat_ld, (size_t)feat_ld * 8, base + L.feat,
                                      (size_t)ncol * 8, (size_t)ncol * 8, (size_t)cn, cudaMemcpyDeviceToHost, st));
        }
        if (bates) {
            if (ier)
                PFX_TRY(cudaMemcpyAsync(ier + (size_t)done * PFX_NIER, base + L.ier, (size_t)cn * PFX_NIER,
                                        cudaMemcpyDeviceToHost, st));
            PFX_TRY(cudaMemcpyAsync(status + done, base + L.status, (size_t)cn * 4, cudaMemcpyDeviceToHost, st));
        }
        done += cn;
        b ^= 1;
    }
#undef PFX_TRY
out:
    {
        const cudaError_t e0 = cudaStreamSynchronize(ctx->streams[0]), e1 = cudaStreamSynchronize(ctx->streams[1]);
        if (rc) return rc;
        if (e0 != cudaSuccess || e1 != cudaSuccess)
            return set_err(ctx, PFX_ERR_CUDA, "pfx_score_host: %s", cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
    }
    ctx->last_launches = launches;
    ctx->ev_valid = false;
    return PFX_OK;
}
