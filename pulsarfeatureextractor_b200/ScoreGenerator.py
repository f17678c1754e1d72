"""
ScoreGenerator -- command line of the reference (ScoreGenerator.py:99-292), flag compatible:

    python -m pulsarfeatureextractor_b200.ScoreGenerator -c DIR -o OUT --phcx [--arff] [--dmprof]

  -c PATH     directory with candidates, or a single candidate file
  -o PATH     output file; when it cannot be created, scores go to <candidate>.dat files
  -v          verbose
  --pfd / --phcx / --superb    which candidate files to look for
  --arff      ARFF rows and header instead of CSV
  --profile   raw profile bins instead of scores
  --label     (reference: interactive labelling; not part of the scoring path)
  --dmprof    the 8 Lyon statistics instead of the 22 Bates scores

Additive options of this build: --features {bates,lyon,all}, --gpu N (first device), --gpus G (score
on G devices of this box: candidates are dealt to one context and host thread per device, rows come
back in input order), --batch N, --repr (shortest round-trip float formatting instead of Python 2's
12 significant digits).
"""
import os
import sys
from optparse import OptionParser

from .DataProcessor import DataProcessor


class ScoreGenerator(object):
    def main(self, argv=None):
        parser = OptionParser()
        parser.add_option("-c", action="store", dest="candDir", default="",
                          help="the path to the directory containing candidates (required).")
        parser.add_option("-v", action="store_true", dest="verbose", default=False, help="Verbose debugging flag (optional).")
        parser.add_option("-o", action="store", dest="outputPath", type="string", default="",
                          help="The path to write scores to (required).")
        parser.add_option("--pfd", action="store_true", dest="pfd", default=False)
        parser.add_option("--phcx", action="store_true", dest="phcx", default=False)
        parser.add_option("--superb", action="store_true", dest="superb", default=False)
        parser.add_option("--arff", action="store_true", dest="arff", default=False)
        parser.add_option("--profile", action="store_true", dest="profile", default=False)
        parser.add_option("--label", action="store_true", dest="label", default=False)
        parser.add_option("--dmprof", action="store_true", dest="dmprof", default=False)
        parser.add_option("--features", action="store", dest="features", default=None,
                          help="bates (default), lyon or all (22 Bates followed by 8 Lyon columns)")
        parser.add_option("--gpu", action="store", dest="gpu", type="int", default=0)
        parser.add_option("--gpus", action="store", dest="gpus", type="int", default=1,
                          help="number of GPUs of this box to shard the candidates over (no collective)")
        parser.add_option("--batch", action="store", dest="batch", type="int", default=32768,
                          help="candidates per GPU batch (the kernels end in a tail of slow candidates, so few large batches win)")
        parser.add_option("--repr", action="store_true", dest="repr", default=False)
        args, _ = parser.parse_args(argv)

        single_file = os.path.isfile(args.outputPath)
        if not single_file:
            try:
                open(args.outputPath, "w").close()
            except IOError:
                pass
            single_file = os.path.isfile(args.outputPath)
        process_single = False
        if os.path.isdir(args.candDir):
            search = args.candDir + "/"
        elif os.path.isfile(args.candDir):
            search = args.candDir
            process_single = True
        else:
            search = ""

        print("\n***********************************")
        print("| Executing score generation code |")
        print("***********************************")
        print("\tOutput path:", args.outputPath, "\n\tCandidate directory:", args.candDir,
              "\n\tWrite to single file:", single_file, "\n\tProcess single candidate:", process_single, "\n")

        from .engine import make_engine
        batch = args.batch * max(1, args.gpus)
        dp = DataProcessor(args.verbose, engine=make_engine(args.gpus, args.gpu), batch=batch, features=args.features,
                           py2_format=not args.repr)
        if args.label:
            print("Labelling is interactive and outside the scoring path; nothing to do.")
        elif args.dmprof:
            if args.phcx and not args.pfd and not args.superb:
                dp.dmprofPHCX(search, args.verbose, args.outputPath, args.arff, process_single)
            elif args.pfd and not args.phcx and not args.superb:
                dp.dmprofPFD(search, args.verbose, args.outputPath, args.arff, process_single)
            elif args.superb and not args.phcx and not args.pfd:
                dp.dmprofSUPERB(search, args.verbose, args.outputPath, args.arff, process_single)
        elif args.phcx and not args.pfd and not args.superb:
            if not single_file:
                dp.processPHCXSeparately(search, args.verbose, process_single)
            else:
                dp.processPHCXCollectively(search, args.verbose, args.outputPath, args.arff, args.profile, process_single)
        elif args.pfd and not args.phcx and not args.superb:
            if not single_file:         # ScoreGenerator.py:205-211 of the reference
                dp.processPFDSeparately(search, args.verbose, process_single)
            else:
                dp.processPFDCollectively(search, args.verbose, args.outputPath, args.arff, args.profile, process_single)
        elif args.superb and not args.pfd and not args.phcx:
            dp.processSUPERBCollectively(search, args.verbose, args.outputPath, args.arff, args.profile, process_single)
        elif not args.superb:
            if not single_file:
                dp.processPFDAndPHCXSeparately(search, args.verbose, process_single)
            else:
                dp.processPFDAndPHCXCollectively(search, args.verbose, args.outputPath, args.arff, args.profile, process_single)
        else:
            print("Didn't know what to do with your input.")
        print("Done.")


if __name__ == "__main__":
    ScoreGenerator().main(sys.argv[1:])
