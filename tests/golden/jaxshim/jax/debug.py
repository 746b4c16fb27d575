def print(*a, **k):   # noqa: A001
    return None


def callback(*a, **k):
    return None
