"""Drop the B200 classes into the reference's own harness without touching the reference tree.

    python -m rgbd_b200.dropin --reference /path/to/Learning-based-RGB-D-Image-Compression \
        -- -m ELIC_united --channel 4 -q 2_2 --dataset /data/nyuv2 ...

`testing/tester.py:11` binds `from models import modelZoo`, i.e. the same dict object as
`models.modelZoo`, so re-assigning its two entries in place (which keeps the dict order: the R2D key
must stay ahead of `ELIC_united`, models/__init__.py:11-20) is enough for
`Tester.get_net` (testing/tester.py:55-59) to construct our classes.  The tester writes to cwd-relative
`../experiments/...`, so run from a writable scratch directory such as `<scratch>/playground/`.
"""
import argparse
import os
import sys


def install(reference_root, compressai_root=None):
    """Put the reference on sys.path and swap the two model classes. Returns the patched dict."""
    import rgbd_b200
    compressai_root = compressai_root or os.path.join(reference_root, "CompressAI")
    for p in (compressai_root, reference_root):
        if p not in sys.path:
            sys.path.insert(0, p)
    import models  # the reference's package
    models.modelZoo["ELIC_united_R2D"] = rgbd_b200.ELIC_united_R2D
    models.modelZoo["ELIC_united"] = rgbd_b200.ELIC_united
    models.modelZoo["ELIC"] = rgbd_b200.ELIC            # single-modality baselines (testing/tester_single.py)
    models.modelZoo["STF_united"] = rgbd_b200.SymmetricalTransFormerUnited
    return models.modelZoo


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", required=True, help="root of the reference checkout")
    ap.add_argument("--compressai", default=None, help="root that contains the `compressai` package")
    ap.add_argument("rest", nargs=argparse.REMAINDER, help="arguments for playground/test.py after `--`")
    args = ap.parse_args(argv)
    install(args.reference, args.compressai)
    rest = [a for a in args.rest if a != "--"]
    sys.argv = ["test.py"] + rest
    from playground import test as ref_test
    ref_test.main(rest) if "argv" in ref_test.main.__code__.co_varnames else ref_test.main()


if __name__ == "__main__":
    main()
