"""Empty stand-in (rendering only; not on the step path)."""
class Surface:  # referenced in a type annotation of renderer_utils
    pass
