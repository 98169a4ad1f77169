"""Import shim so the UNMODIFIED reference can be imported in the build container, where nltk is
not installed (used only by tests/golden/make_golden.py).  The two calls the reference makes are
forwarded to the repo's own restatement of them."""


def download(*_args, **_kwargs):
    return True
