"""Empty stand-in: the reference imports matplotlib.pyplot at module scope but the step path never uses it."""
