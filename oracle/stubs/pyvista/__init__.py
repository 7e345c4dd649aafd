"""Empty stand-in (rendering only; not on the step path)."""
