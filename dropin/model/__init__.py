"""Drop-in `model` package: deeplab_multi / discriminator / deeplab_vgg are the B200 implementations in this directory;
every other sub-module the reference's scripts import (`model.warper`, `model.deeplab`, `model.custom_layers`) is
resolved from the reference checkout itself -- found as the sibling `model/` directory of any later sys.path entry or
of $ASN_REFERENCE -- so this directory only has to sit IN FRONT of the checkout on sys.path."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _root in [os.environ.get("ASN_REFERENCE", "")] + list(sys.path):
    _cand = os.path.join(_root or ".", "model")
    if _root and os.path.isdir(_cand) and os.path.abspath(_cand) != _here and os.path.exists(os.path.join(_cand, "warper.py")):
        if _cand not in __path__:
            __path__.append(_cand)
        break
