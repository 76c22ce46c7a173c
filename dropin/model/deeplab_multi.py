"""Drop-in for the reference's model/deeplab_multi.py: same names, B200 kernels underneath."""
from adaptsegnet_b200.model.deeplab_multi import (Bottleneck, Classifier_Module, DeeplabMulti,  # noqa: F401
                                                   ResNetMulti)

affine_par = True
