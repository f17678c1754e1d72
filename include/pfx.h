/*
 * pfx.h -- C ABI of libpfx: batched pulsar-candidate scoring on NVIDIA B200 (sm_100a).
 *
 * Drop-in boundary for the per-candidate scoring path of scienceguyrob/PulsarFeatureExtractor.
 * Every entry point cites the reference interface it replaces; paths are relative to
 * PulsarFeatureExtractor/src/ of the reference.  The reference API is per candidate and
 * returns Python lists; here one call scores a whole structure-of-arrays batch.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  No hidden cross-call state: the values the
 *     reference keeps on its operator object (self.snr/dm/period/width, PHCXOperations.py:107-110)
 *     travel in the batch (`params`).
 *   - `*_dev` entry points take DEVICE pointers, enqueue work on `stream` (a cudaStream_t
 *     passed as void*) and return without synchronising, provided the context's scratch is large
 *     enough: it grows on demand (a device allocation, and a device synchronisation when an older
 *     block has to be freed), so call pfx_reserve() once with the largest batch to keep every later
 *     call free of allocations and host synchronisation.  Any number of calls may be in flight on
 *     any streams of one context: scratch is handed out in sets whose reuse is ordered on the device
 *     by events (more calls in flight than sets serialise, they never share scratch).  The host-side
 *     entry points of one context are not re-entrant: one host thread per context.
 *     `pfx_score_host` takes HOST pointers,
 *     streams chunks through context-owned device staging buffers (H2D copy, kernels, D2H
 *     copy) and returns when the results are in the caller's host arrays.
 *   - All data buffers are caller-owned.  The context owns only scratch.
 *   - Return value: 0 on success, negative pfx_error on failure; pfx_last_error() gives text.
 *   - Per-candidate failures are NOT call failures: they are reported in `status[i]`
 *     (which reference exception would have fired) and the feature row is left as computed
 *     so far (NaN for the groups not reached), mirroring DataProcessor.py:517-523 where the
 *     row is skipped and the path logged.
 */
#ifndef PFX_H
#define PFX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFX_VERSION 100

/* ---- error codes ------------------------------------------------------------------- */
typedef enum {
    PFX_OK = 0,
    PFX_ERR_INVALID = -1,      /* bad argument */
    PFX_ERR_CUDA = -2,         /* CUDA runtime error, see pfx_last_error */
    PFX_ERR_UNSUPPORTED = -3,  /* shape beyond the compiled limits */
    PFX_ERR_NOMEM = -4
} pfx_error;

/* ---- element types of the input arrays ------------------------------------------- */
typedef enum {
    PFX_U8 = 0,   /* PHCX family: the file stores two hex digits per sample (PHCXFile.py:169-176) */
    PFX_F32 = 1,
    PFX_F64 = 2   /* PFD family: float64 profiles (PFDFile.py:96-252) */
} pfx_dtype;

/* ---- candidate family: selects the DM-curve model and constants -------------------- */
typedef enum {
    PFX_FAMILY_PHCX = 0,  /* PHCXOperations.py:121-233: 3-parameter fit, df=400, f=1374 */
    PFX_FAMILY_PFD = 1    /* PFDOperations.py:274-393:  4-parameter fit, df=32,  f=135  */
} pfx_family;

/* ---- per-candidate status bits (same values as oracle/pfx_oracle.py) ---------------- */
#define PFX_ST_HIST_ZERO_BINS   (1u << 0)  /* numpy.histogram(bins=0) ValueError, ProfileOperations.py:657,678 */
#define PFX_ST_LSQ_IMPROPER     (1u << 1)  /* leastsq N > M TypeError, ProfileOperations.py:1045 */
#define PFX_ST_INDEX_OOB        (1u << 2)  /* IndexError outside the guarded block, ProfileOperations.py:972-974 */
#define PFX_ST_SUBBAND_NO_PAIRS (1u << 3)  /* ZeroDivisionError, ProfileOperations.py:1681,1684 */
#define PFX_ST_DM_NDEG_ZERO     (1u << 4)  /* ZeroDivisionError, PFDOperations.py:385 */
#define PFX_ST_DG_SENTINEL      (1u << 5)  /* soft: 1e6 sentinels written, ProfileOperations.py:762-764 */
#define PFX_ST_SUBBAND_BAD_WIDTH (1u << 6) /* width_bins > nbin, ProfileOperations.py:1630 */
#define PFX_ST_HIST_OVERFLOW    (1u << 7)  /* histogram larger than the engine's scratch (engine limit) */
#define PFX_ST_OTHER            (1u << 8)
#define PFX_ST_EMPTY_SEQUENCE   (1u << 9)  /* argmax of an empty sequence, ProfileOperations.py:1361 */
#define PFX_ST_FAIL_MASK (PFX_ST_HIST_ZERO_BINS | PFX_ST_LSQ_IMPROPER | PFX_ST_INDEX_OOB | \
                          PFX_ST_SUBBAND_NO_PAIRS | PFX_ST_DM_NDEG_ZERO | PFX_ST_SUBBAND_BAD_WIDTH | \
                          PFX_ST_HIST_OVERFLOW | PFX_ST_OTHER | PFX_ST_EMPTY_SEQUENCE)

/* ---- least-squares slots: ier[i][slot] = MINPACK info of that leastsq call site ------ */
#define PFX_FIT_SINE        0   /* ProfileOperations.py:449 */
#define PFX_FIT_SINESQ      1   /* ProfileOperations.py:549 */
#define PFX_FIT_DYHIST      2   /* ProfileOperations.py:949 (derivative histogram, last retry) */
#define PFX_FIT_PROFHIST    3   /* ProfileOperations.py:949 (profile histogram, last retry) */
#define PFX_FIT_FIXED       4   /* ProfileOperations.py:1045 */
#define PFX_FIT_T1          5   /* ProfileOperations.py:1246 */
#define PFX_FIT_DG_ROUND0   6   /* ProfileOperations.py:1379, rounds 1..8 -> slots 6..13 */
#define PFX_FIT_DG_FINAL    14  /* ProfileOperations.py:1473 */
#define PFX_FIT_DM          15  /* PHCXOperations.py:212 / PFDOperations.py:369 */
#define PFX_NFIT            16
#define PFX_FIT_DYHIST_COUNT   16  /* number of fits the retry loop did (1..6) */
#define PFX_FIT_PROFHIST_COUNT 17
#define PFX_NIER            18

#define PFX_NPARAM 8   /* params[i] = snr, dm, bary_period_s, width, dm_start, dm_end, dm_len_all, reserved */
#define PFX_NBATES 22
#define PFX_NLYON  8
#define PFX_NALL   30

/* feature selection bits for pfx_score_* */
#define PFX_FEAT_BATES (1u << 0)
#define PFX_FEAT_LYON  (1u << 1)
#define PFX_FEAT_ALL   (PFX_FEAT_BATES | PFX_FEAT_LYON)

/* Shape and typing of one SoA batch.  Arrays are C-contiguous:
 *   profile  [n][nbin]         dtype
 *   subbands [n][nsub][nbin]   dtype
 *   dmcurve  [n][n_dm]         dm_dtype   DM-SNR curve (PHCX: after the max-of-127 reduction of
 *                                          PHCXOperations.py:237-259; PFD: chi-square curve, PFDFile.py:378-423)
 *   dmstat   [n][n_dmstat]     PFX_U8/..  bytes of DataBlock[0] for the Lyon DM statistics
 *                                          (PHCXOperations.py:528-539); may alias dmcurve
 *   params   [n][PFX_NPARAM]   float64
 */
typedef struct {
    int32_t nbin;
    int32_t nsub;
    int32_t n_dm;
    int32_t n_dmstat;
    int32_t dtype;        /* pfx_dtype of profile and subbands */
    int32_t dm_dtype;     /* pfx_dtype of dmcurve */
    int32_t dmstat_dtype; /* pfx_dtype of dmstat */
    int32_t family;       /* pfx_family */
} pfx_shape;

typedef struct pfx_ctx pfx_ctx;

/* Per-stage device timings of the last pfx_score_* call, milliseconds (CUDA events). */
typedef struct {
    float lyon;
    float sinusoid;
    float histgauss;
    float gauss;
    float dmfit;
    float subband;
    float total;
    int32_t launches;   /* kernels launched by the call */
} pfx_timing;

/* ---- context ------------------------------------------------------------------------- */
int pfx_version(void);
/* One context per (host thread, device).  Contexts are independent: 8-GPU sharding = 8 contexts. */
int pfx_create(int device, pfx_ctx** out);
void pfx_destroy(pfx_ctx* ctx);
const char* pfx_last_error(const pfx_ctx* ctx);
/* SM count, shared memory per block and scratch bytes held by the context. */
int pfx_device_info(const pfx_ctx* ctx, int* sm_count, int* smem_per_block, size_t* scratch_bytes);

/* Size the context's scratch for batches of up to n_max candidates of this shape (work-queue
 * counters, the records passed between the two Gaussian kernels, histogram buffers; staging of
 * pfx_score_host is sized by its first call).  After it no *_dev call with n <= n_max and this shape
 * allocates or synchronises.  Optional: without it scratch grows lazily inside the calls. */
int pfx_reserve(pfx_ctx* ctx, const pfx_shape* shape, int64_t n_max);
/* The histogram -> Gaussian chain of pfx_score_dev (the long pole, a persistent kernel pair that ends
 * in a tail of slow candidates) is cut into `nsub` sub-batches on separate streams so that the tail
 * of one overlaps the body of the next; 0 = default = 1 = off (measured slower on B200, see
 * DESIGN.md; kept as a knob).  Results do not depend on it: a candidate's row is a function of the
 * candidate alone. */
int pfx_set_subbatches(pfx_ctx* ctx, int nsub);
/* DM-curve fit (scores 17, 18; PHCXOperations.py:212, PFDOperations.py:369).  on = 1 (default): every sum
 * inside lmdif runs in MINPACK's own order without contraction, and scores 16-19 and ier equal scipy's
 * bit for bit (the model needs only correctly rounded operations).  on = 0: the faster streaming
 * factorisation of the other fits, equal to scipy at the rounding level only. */
int pfx_set_dm_exact(pfx_ctx* ctx, int on);

/* ---- Lyon statistics -------------------------------------------------------------------
 * Replaces PHCX.computeProfileStatScores / computeDMCurveStatScores (PHCXFile.py:320-379,
 * PFDFile.py:522-583) called from Candidate.calculateProfileStatScores / calculateDMCurveStatScores
 * (Candidate.py:190-286).  feat[i] = mean, std, skew, kurtosis of the profile, then of dmstat. */
int pfx_lyon8_dev(pfx_ctx* ctx, const pfx_shape* shape, int64_t n,
                  const void* profile, const void* dmstat,
                  double* feat, int64_t feat_ld, void* stream);

/* ---- Bates scores, one entry per operator-API method ---------------------------------
 * Each writes its columns of feat[i] (row stride feat_ld doubles, column = score number - 1),
 * ORs failure bits into status[i] and stores MINPACK info codes into ier[i][PFX_NIER].
 * `ier` may be NULL. */

/* scores 1-4: ProfileOperationsInterface.getSinusoidFittings (ProfileOperations.py:190-587) */
int pfx_sinusoid_dev(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, const void* profile,
                     double* feat, int64_t feat_ld, int8_t* ier, int32_t* status, void* stream);

/* scores 5-11: ProfileOperationsInterface.getGaussianFittings (ProfileOperations.py:595-1483) */
int pfx_gaussian_dev(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, const void* profile,
                     double* feat, int64_t feat_ld, int8_t* ier, int32_t* status, void* stream);

/* scores 12-19: getCandidateParameters + getDMFittings (PHCXOperations.py:81-233,
 * PFDOperations.py:93-393) with CandidateFileInterface.filterScore (CandidateFileInterface.py:84-112) */
int pfx_dmfit_dev(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, const void* profile,
                  const void* dmcurve, const double* params,
                  double* feat, int64_t feat_ld, int8_t* ier, int32_t* status, void* stream);

/* PFD family only: derive S/N and width from the profile the way PFDOperations.getCandidateParameters
 * does (PFDOperations.py:130-231: 3-sigma clipped S/N, half-maximum width of the centred profile)
 * and store them in params[i][0] and params[i][3]; the other parameters (period, DM, DM range)
 * come from the file header and are left untouched.  Call before pfx_dmfit_dev / pfx_score_dev. */
int pfx_pfd_params_dev(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, const void* profile,
                       double* params, void* stream);

/* scores 20-22: getSubbandParameters (PHCXOperations.py:305-415, PFDOperations.py:401-466,
 * ProfileOperations.py:1585-1686) */
int pfx_subband_dev(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, const void* profile,
                    const void* subbands, const double* params,
                    double* feat, int64_t feat_ld, int32_t* status, void* stream);

/* ---- whole candidates ------------------------------------------------------------------
 * Replaces CandidateFileInterface.compute (PHCXFile.py:383-409, PFDFile.py:587-613) =
 * Candidate.calculateScores (Candidate.py:116-150), optionally followed by the Lyon statistics.
 * feat row layout: Bates 1..22 in columns 0..21 when PFX_FEAT_BATES is set, followed by the 8 Lyon
 * columns when PFX_FEAT_LYON is set (Lyon only: columns 0..7). */
int pfx_score_dev(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, uint32_t features,
                  const void* profile, const void* subbands, const void* dmcurve,
                  const void* dmstat, const double* params,
                  double* feat, int64_t feat_ld, int8_t* ier, int32_t* status, void* stream);

/* Same call with HOST buffers (pinned memory gives asynchronous copies): the batch is cut into
 * chunks of `chunk` candidates (0 = default), each chunk is copied H2D on a copy stream,
 * scored, and its feature rows / ier / status copied back, double-buffered so copies overlap
 * kernels.  Returns after the last D2H copy completed. */
int pfx_score_host(pfx_ctx* ctx, const pfx_shape* shape, int64_t n, uint32_t features,
                   const void* profile, const void* subbands, const void* dmcurve,
                   const void* dmstat, const double* params,
                   double* feat, int64_t feat_ld, int8_t* ier, int32_t* status, int64_t chunk);

/* Stage timing.  By default pfx_score_dev overlaps independent score groups on a side stream.
 * With stage timing on, the groups run back to back on the caller's stream with a CUDA event
 * between them, and pfx_last_timing (which synchronises) returns their durations.
 * on == 2 keeps the production stream layout and records when each group completed: the fields
 * of pfx_timing then hold completion times since the start of the call (overlap trace; subband
 * is reported with dmfit, which shares its stream). */
int pfx_set_stage_timing(pfx_ctx* ctx, int on);
int pfx_last_timing(pfx_ctx* ctx, pfx_timing* out);

#ifdef __cplusplus
}
#endif
#endif /* PFX_H */
