"""tensorboard stand-in -- TEST INFRASTRUCTURE ONLY."""
from . import program  # noqa: F401
