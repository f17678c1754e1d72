"""
Candidate -- facade kept from the reference (Candidate.py:42-547), re-pointed at the batch engine.

The reference builds a reader object per call and scores one candidate on the CPU.  Here a
Candidate is a thin record: it knows its file, which reader applies (same file-name dispatch as
Candidate.py:136-150: '.pfd' -> PFD, '.gz' -> HTRU PHCX section 1, anything else -> SUPERB PHCX
section 0) and can score itself through a (shared) Engine as a batch of one.  Bulk work goes
through DataProcessor, which packs many candidates into one structure-of-arrays batch.
"""
import numpy as np

from . import pfd_io, phcx_io


class Candidate(object):
    def __init__(self, name="Unknown", path=""):
        self.candidateName = name      # Candidate.py:66
        self.candidatePath = path
        self.scores = []
        self.label = "Unknown"
        self.special = "None"

    # ---- reader dispatch (Candidate.py:136-150) ------------------------------------------------
    def family(self):
        # the reference tests the whole path for the substrings (Candidate.py:136-150); the file name
        # is what is meant, and a directory called 'x.pfd' must not re-route PHCX files
        import os
        name = os.path.basename(self.candidateName)
        if ".pfd" in name:
            return "pfd"
        if ".gz" in name:
            return "phcx"
        return "superb"

    def load(self):
        fam = self.family()
        if fam == "pfd":
            return pfd_io.read_pfd(self.candidateName)
        return phcx_io.read_phcx(self.candidateName, section=1 if fam == "phcx" else 0)

    # ---- scoring entry points, same names as the reference ----------------------------------------
    def _score(self, engine, features):
        if features == "lyon" and self.family() != "pfd":
            # the reference touches only the profile and DataBlock[0] here (PHCXFile.py:320-379)
            c = phcx_io.read_phcx(self.candidateName, section=1 if self.family() == "phcx" else 0, lyon_only=True)
            out = engine.score_host(np.ascontiguousarray(c.profile[None]), None, None,
                                    np.ascontiguousarray(c.dmstat[None]), None, features="lyon")
            return [float(v) for v in out.features[0]]
        c = self.load()
        if self.family() == "pfd":
            out = engine.score_pfd_host(c.profile[None], c.subbands[None], c.dmcurve[None], c.params[None],
                                        features=features)
            if features != "lyon" and not out.ok()[0]:
                raise Exception("candidate scoring failed, status=0x%x" % int(out.status[0]))
            return [float(v) for v in out.features[0]]
        out = engine.score_host(np.ascontiguousarray(c.profile[None]), np.ascontiguousarray(c.subbands[None]),
                                np.ascontiguousarray(c.dmcurve[None]), np.ascontiguousarray(c.dmstat[None]),
                                np.ascontiguousarray(c.params[None]), features=features)
        if features != "lyon" and not out.ok()[0]:
            raise Exception("candidate scoring failed, status=0x%x" % int(out.status[0]))
        return [float(v) for v in out.features[0]]

    def calculateScores(self, engine):
        """22 Bates scores (Candidate.py:116-150)."""
        self.scores = self._score(engine, "bates")
        return self.scores

    def calculateProfileStatScores(self, engine):
        """mean, std, skew, kurtosis of the profile (Candidate.py:190-227)."""
        return self._score(engine, "lyon")[0:4]

    def calculateDMCurveStatScores(self, engine):
        """mean, std, skew, kurtosis of the DM curve (Candidate.py:257-286)."""
        return self._score(engine, "lyon")[4:8]

    def calculateProfileScores(self, engine=None):
        """Raw profile bins as floats (PHCXFile.py:291-304)."""
        return [float(v) for v in self.load().profile]

    def getName(self):
        return self.candidateName

    def getPath(self):
        return self.candidatePath
