class Figure:  # referenced only in a type annotation (wrappers.py:177)
    pass
