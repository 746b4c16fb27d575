import contextlib


def annotate_function(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


@contextlib.contextmanager
def TraceAnnotation(*a, **k):
    yield
