"""Backs the nltk / rapidfuzz shims.  Loaded by file path so that it does not depend on the
product package being importable under the same name as the reference's package."""
import importlib.util
import pathlib

_ROOT = pathlib.Path(__file__).resolve().parents[2]


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_port = _load("_nsm_oracle_port", _ROOT / "oracle" / "reference_port.py")
_tok_src = (_ROOT / "napkon-string-matching_b200" / "napkon_string_matching" / "text" /
            "tokenize.py").read_text(encoding="utf-8")
_ns: dict = {"__name__": "_nsm_tok"}
exec(compile(_tok_src, "tokenize.py", "exec"), _ns)  # noqa: S102 - our own file
GERMAN_STOP_WORDS = list(_ns["_GERMAN_STOP_WORDS"])
word_tokenize = _ns["word_tokenize"]


def qratio_percent(s1, s2):
    a, b = _port.default_process(s1), _port.default_process(s2)
    if not a or not b:
        return 0
    lcs = _port.lcs_length(a, b)
    lensum = len(a) + len(b)
    norm_dist = (lensum - 2 * lcs) / lensum if lensum else 0.0
    return (1.0 - norm_dist) * 100
