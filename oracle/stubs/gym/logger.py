def warn(*a, **k):
    pass
info = debug = error = warn
