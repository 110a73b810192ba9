"""Decoders that consume the channel-last feature render."""
from .networks import CNN_decoder, CNN_scale_decoder  # noqa: F401
