"""adaptsegnet_b200: B200 (sm_100a) implementation of AdaptSegNet's output-space-adaptation hot
path behind the reference's own module API.  See DESIGN.md; C ABI in include/asn_b200.h."""
__version__ = "0.1.0"
