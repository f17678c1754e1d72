"""
DataProcessor -- directory walk, batch submission and the output writers of the reference
(DataProcessor.py:51-995), with the per-candidate loop replaced by batches on the GPU.

Kept from the reference: file patterns (DataProcessor.py:67-70), os.walk + fnmatch order, the
CSV / ARFF / .dat row formats (DataProcessor.py:305-325,405-447), the 'nan' and 'inf' -> '0'
substitution over the whole line, ARFF headers (DataProcessor.py:329-401), the per-candidate
failure handling (row omitted, path appended to CandidateErrorLog.txt, DataProcessor.py:517-523),
the counters printed at the end, rows buffered and appended once (DataProcessor.py:590-594).
Numbers are written with Python 2's str(float) = '%.12g' (SURVEY.md 8c D3); pass py2_format=False for
repr().
"""
import datetime
import fnmatch
import os
import sys

import numpy as np

from . import _lib, pfd_io, phcx_io
from .Candidate import Candidate


def fmt_float(v, py2=True):
    v = float(v)
    if not py2:
        return repr(v)
    if v != v:
        return "nan"
    if v in (float("inf"), float("-inf")):
        return "inf" if v > 0 else "-inf"
    s = "%.12g" % v
    if "." not in s and "e" not in s:
        s += ".0"
    return s


class DataProcessor(object):
    def __init__(self, debugFlag=False, engine=None, batch=32768, features=None, py2_format=True,
                 error_log="CandidateErrorLog.txt"):
        self.debug = debugFlag
        self.phcxRegex = "*.phcx.gz"
        self.superbRegex = "*.phcx"
        self.pfdRegex = "*.pfd"
        self.pfdScrunchedRegex = "*.pfd.36scrunch"
        self.scoreStore = []
        self.candidateErrorLog = error_log
        self.engine = engine
        self.batch = int(batch)
        self.features = features          # None: what the reference mode implies
        self.py2 = py2_format
        self.pfd = self.phcx = self.superb = False
        if not os.path.isfile(self.candidateErrorLog):
            self.appendToFile(self.candidateErrorLog, "")

    # ---- Utilities.py:60-115 -------------------------------------------------------------------
    @staticmethod
    def appendToFile(path, text):
        with open(path, "a") as fh:
            fh.write(str(text))

    @staticmethod
    def fileExists(path):
        return os.path.isfile(path)

    def _engine(self):
        if self.engine is None:
            from .engine import Engine
            self.engine = Engine(0)
        return self.engine

    # ---- row formats ----------------------------------------------------------------------------
    def storeScore(self, candidate, scores, outPath=None):
        entry = candidate + "," + ",".join(fmt_float(s, self.py2) for s in scores)
        self.scoreStore.append(entry.replace("nan", "0").replace("inf", "0"))

    def storeScoreARFF(self, candidate, scores, outPath=None):
        entry = ",".join(fmt_float(s, self.py2) for s in scores) + ",?%" + candidate
        self.scoreStore.append(entry.replace("nan", "0").replace("inf", "0"))

    def outputScores(self, scores, outputFile):
        with open(outputFile + ".dat", "w") as fh:
            fh.write(",".join(fmt_float(s, self.py2) for s in scores).replace("nan", "0").replace("inf", "0"))

    def prepareARFFFile(self, path, genProfileData, n_scores=22):
        header = "@relation PulsarCandidates_" + datetime.datetime.now().isoformat() + "\n"
        attributes = n_scores + 1
        if genProfileData and self.superb:
            attributes = 65
        elif genProfileData and self.phcx:
            attributes = 129
        for n in range(1, attributes):
            header += "@attribute Score" + str(n) + " numeric\n"
        header += "@attribute class {0,1}\n@data\n"
        self.appendToFile(path, header)

    def prepareDMProfileARFFFile(self, path):
        header = "@relation PulsarCandidates_" + datetime.datetime.now().isoformat() + "\n"
        for name in ("Profile_mean", "Profile_stdev", "Profile_skewness", "Profile_kurtosis",
                     "DM_mean", "DM_stdev", "DM_skewness", "DM_kurtosis"):
            header += "@attribute " + name + " numeric\n"
        header += "@attribute class {0,1}\n@data\n"
        self.appendToFile(path, header)

    # ---- candidate discovery (DataProcessor.py:491-497) --------------------------------------------
    def find(self, directory, fileTypeRegexes, processSingleCandidate):
        if processSingleCandidate:
            return [directory]
        if directory == "":
            directory = os.path.dirname(os.path.realpath(__file__))
            print("User has not provided a search directory - searching local directory")
        found = []
        for filetype in fileTypeRegexes:
            for root, _sub, filenames in os.walk(directory):
                for filename in fnmatch.filter(filenames, filetype):
                    found.append(os.path.join(root, filename))
        return found

    # ---- the batched core ------------------------------------------------------------------------
    def _score_chunks(self, paths, features):
        """Yields (chunk of paths, feature matrix [len(chunk)][ncol], ok mask) in input order; rows of
        candidates the reference would have raised for are NaN with ok = False."""
        eng = self._engine()
        ncol = {"bates": 22, "lyon": 8, "all": 30}[features]
        for lo in range(0, len(paths), self.batch):
            chunk = paths[lo:lo + self.batch]
            feat = np.full((len(chunk), ncol), np.nan, dtype=np.float64)
            okm = np.zeros(len(chunk), dtype=bool)

            def score_group(ks, prof, sub, dmc, dms, par):
                out = eng.score_host(prof, sub, dmc, dms, par, features=features)
                ok = out.ok() if features != "lyon" else np.ones(len(ks), dtype=bool)
                ks = np.asarray(ks)
                feat[ks] = out.features
                okm[ks] = ok

            if phcx_io.native_available() and all(".pfd" in os.path.basename(p) for p in chunk):
                # PRESTO .pfd files: native reader (de-dispersion, profile, sub-bands, chi-square curve),
                # S/N and width on the device, PFD family constants
                scan = pfd_io.PfdScan(chunk)
                groups = {}
                for k in range(len(chunk)):
                    if scan.ok(k):
                        groups.setdefault(scan.shape_key(k), []).append(k)
                    else:
                        print("Error reading profile data :\n\t", chunk[k])
                for _key, ks in groups.items():
                    prof, sub, dmc, par = scan.gather(ks)
                    out = eng.score_pfd_host(prof, sub, dmc, par, features=features)
                    ok = out.ok() if features != "lyon" else np.ones(len(ks), dtype=bool)
                    ks = np.asarray(ks)
                    feat[ks] = out.features
                    okm[ks] = ok
                scan.close()
            elif phcx_io.native_available() and not any(".pfd" in os.path.basename(p) for p in chunk):
                # bulk path: native scanner, rows gathered per shape into pinned buffers
                scan = phcx_io.Scan(chunk)
                lyon_only = features == "lyon"      # --dmprof: only the profile and DataBlock[0] are needed
                groups = {}
                for k in range(len(chunk)):
                    if scan.ok(k, lyon_only):
                        groups.setdefault(scan.shape_key(k, lyon_only), []).append(k)
                    else:
                        print("Error reading profile data :\n\t", chunk[k])
                for _key, ks in groups.items():
                    if lyon_only:
                        prof, dms = scan.gather_lyon(ks)
                        score_group(ks, prof, None, None, dms, None)
                    else:
                        score_group(ks, *scan.gather(ks))
                scan.close()
            else:
                groups = {}
                lyon_files = features == "lyon"
                for k, p in enumerate(chunk):
                    try:
                        cd = Candidate(p, p)
                        if lyon_files and cd.family() != "pfd":
                            c = phcx_io.read_phcx(p, section=1 if cd.family() == "phcx" else 0, lyon_only=True)
                            key = (c.profile.size, None, None, c.dmstat.size)
                        else:
                            c = cd.load()
                            key = (c.profile.size, c.subbands.shape, c.dmcurve.size, c.dmstat.size)
                    except Exception as e:      # noqa: BLE001 - mirrors DataProcessor.py:517
                        print("Error reading profile data :\n\t", type(e).__name__, e)
                        continue
                    groups.setdefault(key, []).append((k, c))
                for _key, items in groups.items():
                    stack = lambda name: (None if getattr(items[0][1], name) is None else
                                          np.ascontiguousarray(np.stack([getattr(c, name) for _k, c in items])))
                    if isinstance(items[0][1], pfd_io.PFDCandidate):
                        out = eng.score_pfd_host(stack("profile"), stack("subbands"), stack("dmcurve"), stack("params"),
                                                 features=features)
                        ks = np.asarray([k for k, _c in items])
                        feat[ks] = out.features
                        okm[ks] = out.ok() if features != "lyon" else True
                        continue
                    score_group([k for k, _c in items], stack("profile"), stack("subbands"), stack("dmcurve"),
                                stack("dmstat"), stack("params"))
            yield chunk, feat, okm

    def _score_files(self, paths, features):
        """Yields (path, scores or None) in input order.  None = the reference would have raised."""
        for chunk, feat, okm in self._score_chunks(paths, features):
            for k, p in enumerate(chunk):
                yield p, (feat[k] if okm[k] else None)

    def _run(self, directory, fileTypeRegexes, outPath, arff, features, processSingleCandidate, separate=False):
        start = datetime.datetime.now()
        paths = self.find(directory, fileTypeRegexes, processSingleCandidate)
        successes = failures = 0
        if not separate and phcx_io.native_available():
            # rows of a chunk are formatted at once by the native formatter (include/pfx_io.h);
            # all text is buffered and appended at the end, as the reference does (DataProcessor.py:590-594)
            blobs = []
            for chunk, feat, okm in self._score_chunks(paths, features):
                for k in np.nonzero(~okm)[0]:
                    print(chunk[k], " did not have scores generated.")
                    self.appendToFile(self.candidateErrorLog, chunk[k] + "\n")
                blobs.append(phcx_io.format_rows(feat, chunk, skip=~okm, mode="arff" if arff else "csv", py2=self.py2))
                successes += int(okm.sum())
                failures += int((~okm).sum())
            end = datetime.datetime.now()
            self.appendToFile(outPath, "".join(blobs))
        else:
            for cand, scores in self._score_files(paths, features):
                if scores is None:
                    print(cand, " did not have scores generated.")
                    self.appendToFile(self.candidateErrorLog, cand + "\n")
                    failures += 1
                    continue
                if separate:
                    self.outputScores(scores, cand)
                elif arff:
                    self.storeScoreARFF(cand, scores, outPath)
                else:
                    self.storeScore(cand, scores, outPath)
                successes += 1
            end = datetime.datetime.now()
            if not separate:
                self.appendToFile(outPath, "".join(s + "\n" for s in self.scoreStore))
                self.scoreStore = []
        print("\nCandidates processed:\t", len(paths))
        print("Successes:\t", successes)
        print("Failures:\t", failures)
        print("Execution time: ", str(end - start))
        return successes, failures

    # ---- public methods, names as in the reference ---------------------------------------------------
    def processCollectively(self, directory, verbose, fileTypeRegexes, outPath, arff, genProfileData,
                            processSingleCandidate):
        feats = self.features or "bates"
        if arff:
            self.prepareARFFFile(outPath, genProfileData, {"bates": 22, "lyon": 8, "all": 30}[feats])
        if genProfileData:
            return self._run_profile(directory, fileTypeRegexes, outPath, arff, processSingleCandidate)
        return self._run(directory, fileTypeRegexes, outPath, arff, feats, processSingleCandidate)

    def _run_profile(self, directory, fileTypeRegexes, outPath, arff, processSingleCandidate):
        paths = self.find(directory, fileTypeRegexes, processSingleCandidate)
        ok = bad = 0
        for p in paths:
            try:
                scores = Candidate(p, p).calculateProfileScores()
                (self.storeScoreARFF if arff else self.storeScore)(p, scores, outPath)
                ok += 1
            except Exception:       # noqa: BLE001
                self.appendToFile(self.candidateErrorLog, p + "\n")
                bad += 1
        self.appendToFile(outPath, "".join(s + "\n" for s in self.scoreStore))
        self.scoreStore = []
        return ok, bad

    def processSeparately(self, directory, verbose, fileTypeRegexes, processSingleCandidate):
        return self._run(directory, fileTypeRegexes, None, False, self.features or "bates",
                         processSingleCandidate, separate=True)

    def dmprof(self, directory, verbose, fileTypeRegexes, outPath, arff, processSingleCandidate):
        if arff:
            self.prepareDMProfileARFFFile(outPath)
        return self._run(directory, fileTypeRegexes, outPath, arff, "lyon", processSingleCandidate)

    def processPHCXCollectively(self, directory, verbose, outPath, arff, genProfileData, processSingleCandidate):
        self.phcx = True
        return self.processCollectively(directory, verbose, [self.phcxRegex], outPath, arff, genProfileData,
                                        processSingleCandidate)

    def processPHCXSeparately(self, directory, verbose, processSingleCandidate):
        self.phcx = True
        return self.processSeparately(directory, verbose, [self.phcxRegex], processSingleCandidate)

    def processSUPERBCollectively(self, directory, verbose, outPath, arff, genProfileData, processSingleCandidate):
        self.superb = True
        return self.processCollectively(directory, verbose, [self.superbRegex], outPath, arff, genProfileData,
                                        processSingleCandidate)

    def dmprofPHCX(self, directory, verbose, outPath, arff, processSingleCandidate):
        return self.dmprof(directory, verbose, [self.phcxRegex], outPath, arff, processSingleCandidate)

    def dmprofSUPERB(self, directory, verbose, outPath, arff, processSingleCandidate):
        return self.dmprof(directory, verbose, [self.superbRegex], outPath, arff, processSingleCandidate)

    # ---- PFD entry points (DataProcessor.py:150-227,271-288) ---------------------------------------
    def processPFDSeparately(self, directory, verbose, processSingleCandidate):
        self.pfd = True
        return self.processSeparately(directory, verbose, [self.pfdRegex, self.pfdScrunchedRegex], processSingleCandidate)

    def processPFDCollectively(self, directory, verbose, outPath, arff, genProfileData, processSingleCandidate):
        self.pfd = True
        return self.processCollectively(directory, verbose, [self.pfdRegex, self.pfdScrunchedRegex], outPath, arff,
                                        False, processSingleCandidate)

    def processPFDAndPHCXSeparately(self, directory, verbose, processSingleCandidate):
        self.pfd = self.phcx = True
        return self.processSeparately(directory, verbose, [self.phcxRegex, self.pfdRegex, self.pfdScrunchedRegex],
                                      processSingleCandidate)

    def processPFDAndPHCXCollectively(self, directory, verbose, outPath, arff, genProfileData, processSingleCandidate):
        self.pfd = self.phcx = True
        return self.processCollectively(directory, verbose, [self.phcxRegex, self.pfdRegex, self.pfdScrunchedRegex],
                                        outPath, arff, False, processSingleCandidate)

    def dmprofPFD(self, directory, verbose, outPath, arff, processSingleCandidate):
        return self.dmprof(directory, verbose, [self.pfdRegex, self.pfdScrunchedRegex], outPath, arff,
                           processSingleCandidate)
