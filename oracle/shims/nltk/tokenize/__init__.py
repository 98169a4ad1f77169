from shim_backend import word_tokenize  # noqa: F401
