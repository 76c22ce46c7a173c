"""Drop-in for the reference's model/discriminator.py."""
from adaptsegnet_b200.model.discriminator import FCDiscriminator  # noqa: F401
