/* pfx_io.h -- C ABI of libpfxio.so: PHCX candidate files -> structure-of-arrays batch rows, and
 * feature rows -> output text.
 *
 * Host-only companion of libpfx (include/pfx.h).  It replaces, for a whole list of files at once,
 * the reading half of the reference's candidate objects:
 *
 *   PHCX.load / getprofile                      PHCXFile.py:90-186      (gzip + minidom + hex loop)
 *   PHCXOperations.getDM_FFT / hexToDec         PHCXOperations.py:263-297,353-383
 *   PHCXOperations.dm_curve                     PHCXOperations.py:237-259 (max of the first 127 of each 128)
 *   DmIndex token rule, Snr/Dm/BaryPeriod/Width PHCXOperations.py:107-110,172-182
 *   section index 1 for HTRU .phcx.gz, 0 for SUPERB .phcx   PHCXFile.py:80, SUPERBPHCXFile.py:80
 *   Lyon DM statistics always from DataBlock[0] PHCXOperations.py:538
 *
 * A file is inflated and scanned once (the reference parses it twice for the Lyon statistics,
 * Candidate.py:218,277); files are spread over worker threads.  The decoded arrays stay inside a
 * scan handle until the caller has gathered them into its own (pinned) batch buffers, grouped by
 * shape, which are then handed to pfx_score_host.
 */
#ifndef PFX_IO_H
#define PFX_IO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pfx_scan pfx_scan;

/* Per file: ok = 1 when the reference's reader would have produced the arrays, 0 when it would have
 * raised (unreadable file, missing section or tag, malformed number, sub-band size mismatch).
 * params = snr, dm, bary_period_s, width, dm_start, dm_end, len(DataBlock[section]), 0
 * (the layout pfx_score_* expects, include/pfx.h). */
typedef struct pfx_phcx_info {
    int32_t ok;
    int32_t nbin;      /* decoded length of <Profile>              */
    int32_t nsub;      /* nSub attribute of <SubBands>             */
    int32_t n_dm;      /* number of complete 128-blocks            */
    int32_t n_dmstat;  /* decoded length of DataBlock[0]           */
    int32_t lyon_ok;   /* 1: Profile[section] and DataBlock[0] were read -- enough for the Lyon statistics
                          (--dmprof, PHCXFile.py:320-379) even when ok == 0; nbin and n_dmstat are then valid */
    double params[8];
} pfx_phcx_info;

/* Parse paths[0..n).  section: 0 or 1 for every file, -1 = by name as the reference's dispatch does
 * (1 when the name contains ".gz", else 0; Candidate.py:141-150).  threads <= 0: hardware
 * concurrency.  info[n] is filled; *out owns the decoded arrays.  Returns 0, or -1 on bad arguments
 * / out of memory (per-file problems are reported through info[i].ok). */
int pfx_phcx_scan(const char* const* paths, int64_t n, int section, int threads,
                  pfx_phcx_info* info, pfx_scan** out);

/* Copy the arrays of files idx[0..count) -- all of the same shape (nbin, nsub, n_dm, n_dmstat) --
 * into rows 0..count-1 of caller-owned C-contiguous buffers:
 *   profile[count][nbin] u8, subbands[count][nsub][nbin] u8, dmcurve[count][n_dm] u8,
 *   dmstat[count][n_dmstat] u8, params[count][8] f64.  Any pointer may be NULL (skipped).
 * Returns 0, -1 on bad arguments, -2 when a file is not ok or its shape differs from the first. */
int pfx_phcx_gather(const pfx_scan* scan, const int64_t* idx, int64_t count, uint8_t* profile,
                    uint8_t* subbands, uint8_t* dmcurve, uint8_t* dmstat, double* params);

void pfx_phcx_free(pfx_scan* scan);

/* ---- PRESTO .pfd files (PFDFile.py:96-456, PFDOperations.py:242-252,474-525) --------------------
 * Per file: header and fold cube profs[npart][nsub][proflen]; avgprof (numpy's pairwise sum of
 * profs / proflen) and varprof before any rotation; de-dispersion at the best DM by whole-bin left
 * rotations; profile = (sumprof - min) / mean scaled to [0, 255] with the reference's formula;
 * sub-bands = profs.sum(0); the reduced chi-square versus DM curve on 100 trial DMs between dms[0]
 * and dms[-1] with rotations accumulated from trial to trial, rounded to float32.
 * Outputs per file: profile f64[proflen], subbands f64[nsub][proflen], dmcurve f32[100] (also the
 * array of the Lyon DM statistics), params = 0 (S/N: pfx_pfd_params_dev), best DM, bary_p1 (s),
 * 0 (width: pfx_pfd_params_dev), DMs[1], DMs[99], 100, 0.
 * ok = 0: unreadable or truncated file, impossible dimensions, fewer than two DMs in the header. */
typedef struct pfx_pfd_info {
    int32_t ok;
    int32_t nbin;      /* proflen */
    int32_t nsub;
    int32_t n_dm;      /* 100 */
    int32_t reserved0, reserved1;
    double params[8];
} pfx_pfd_info;

int pfx_pfd_scan(const char* const* paths, int64_t n, int threads, pfx_pfd_info* info, pfx_scan** out);
/* rows of files idx[0..count) (same nbin, nsub): profile f64, subbands f64, dmcurve f32, params f64 */
int pfx_pfd_gather(const pfx_scan* scan, const int64_t* idx, int64_t count, double* profile,
                   double* subbands, float* dmcurve, double* params);
void pfx_pfd_free(pfx_scan* scan);

/* Output rows (DataProcessor.storeScore / storeScoreARFF, DataProcessor.py:305-325,405-425).
 * Formats rows i < count with skip[i] == 0 (skip may be NULL) of feat[count][ld] (first ncol columns):
 *   mode 0  CSV   "<name>,s1,...,sN\n"
 *   mode 1  ARFF  "s1,...,sN,?%<name>\n"
 *   mode 2  bare  "s1,...,sN\n"            (names may be NULL)
 * Numbers as Python 2's str(float) when py2 != 0 ('%.12g', integer-valued results get '.0'), else as
 * Python 3's repr (shortest digits that round-trip).  As in the reference, every "nan" and then every
 * "inf" of the finished line -- the candidate name included -- is replaced by "0" (so "-inf" becomes
 * "-0").  Rows are formatted by `threads` workers (<= 0: hardware concurrency) and concatenated in
 * order.  *out is allocated by the library (free with pfx_io_free), *out_len is its length in bytes
 * (no terminating NUL counted).  Returns 0, -1 on bad arguments / out of memory. */
int pfx_format_rows(const double* feat, int64_t count, int64_t ld, int ncol, const char* const* names,
                    const int8_t* skip, int mode, int py2, int threads, char** out, int64_t* out_len);
void pfx_io_free(void* p);

const char* pfx_io_version(void);

#ifdef __cplusplus
}
#endif
#endif
