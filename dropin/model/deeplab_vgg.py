"""Drop-in for the reference's model/deeplab_vgg.py (its constructor does not run on Python 3)."""
from adaptsegnet_b200.model.deeplab_vgg import Classifier_Module, DeeplabVGG  # noqa: F401
