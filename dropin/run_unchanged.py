#!/usr/bin/env python
"""Runs one of the reference's UNMODIFIED scripts on the B200 modules:

    python dropin/run_unchanged.py --reference /path/to/AdaptSegNet train -- --level multi-level --gan Vanilla ...
    python dropin/run_unchanged.py --reference /path/to/AdaptSegNet evaluate -- --level multi-level ...

sys.path is arranged as  [dropin/, <reference checkout>, ..., dropin/_shims/]  so that the scripts' own
`from model.deeplab_multi import DeeplabMulti`, `from model.discriminator import FCDiscriminator`,
`from utils.loss import CrossEntropy2d`, `from dataset.* import ...` statements bind to this directory, everything else
(`model.warper`, `model.deeplab`, ...) to the checkout, and `tensorboardX` / `matplotlib` to import shims only when the
real packages are missing.  The script file itself is executed as it is on disk: its text is not
touched.  The one thing a launcher has to do from outside is what SURVEY.md Q2 describes -- the fork hard-codes
`SOURCE_ONLY = True` as a module constant, so the adversarial branches are reachable only by setting the constant
after import (`--source-only 0`, the default here) and then calling its `main()`.
"""
import argparse
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SCRIPTS = {"train": "train_gta2cityscapes_multi", "evaluate": "evaluate_cityscapes"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("ASN_REFERENCE", "/root/reference"))
    ap.add_argument("--source-only", type=int, default=0, help="value for the script's SOURCE_ONLY module constant")
    ap.add_argument("script", choices=sorted(SCRIPTS))
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    ref = os.path.abspath(a.reference)
    if not os.path.exists(os.path.join(ref, SCRIPTS[a.script] + ".py")):
        sys.exit(f"{ref} does not hold {SCRIPTS[a.script]}.py")
    os.environ["ASN_REFERENCE"] = ref
    sys.path[:] = [HERE, ref, os.path.dirname(HERE)] + [p for p in sys.path if p not in (HERE, ref, "")] + \
        [os.path.join(HERE, "_shims")]
    rest = a.rest[1:] if a.rest and a.rest[0] == "--" else a.rest
    sys.argv = [os.path.join(ref, SCRIPTS[a.script] + ".py")] + rest
    mod = importlib.import_module(SCRIPTS[a.script])     # train...: `args = get_arguments()` runs here, at import (Q2)
    assert os.path.abspath(mod.__file__).startswith(ref), mod.__file__
    mod.SOURCE_ONLY = bool(a.source_only)
    mod.main()


if __name__ == "__main__":
    main()
