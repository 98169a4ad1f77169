"""
Host-side tokenisation used before packing.

The reference tokenises with ``nltk.word_tokenize`` and drops NLTK's German stop
words (/root/reference/napkon_string_matching/types/comparable_data.py:287-299).
``nltk`` is a third-party dependency (pinned ``nltk~=3.7``, requirements.txt:3).
When it is importable we call it, so the preprocessing is exactly the
reference's.  When it is absent (as in the build image) we use the restatement
below of NLTK's published word-tokeniser rules (Penn-Treebank style regex
cascade) and of its German stop-word list.  Sentence splitting (punkt) needs a
trained model and is approximated; this only matters for text with periods in
the middle.  Tokenisation runs once per item on the host, never per pair.
"""
from __future__ import annotations

import logging
from bisect import insort
import re
from functools import lru_cache
from typing import Iterable, List

PREPARE_REMOVE_SYMBOLS = "!?,.()[]:;*"

_GERMAN_STOP_WORDS = """
aber alle allem allen aller alles als also am an ander andere anderem anderen anderer anderes
anderm andern anderr anders auch auf aus bei bin bis bist da damit dann der den des dem die das
dass daß derselbe derselben denselben desselben demselben dieselbe dieselben dasselbe dazu dein
deine deinem deinen deiner deines denn derer dessen dich dir du dies diese diesem diesen dieser
dieses doch dort durch ein eine einem einen einer eines einig einige einigem einigen einiger
einiges einmal er ihn ihm es etwas euer eure eurem euren eurer eures für gegen gewesen hab habe
haben hat hatte hatten hier hin hinter ich mich mir ihr ihre ihrem ihren ihrer ihres euch im in
indem ins ist jede jedem jeden jeder jedes jene jenem jenen jener jenes jetzt kann kein keine
keinem keinen keiner keines können könnte machen man manche manchem manchen mancher manches mein
meine meinem meinen meiner meines mit muss musste nach nicht nichts noch nun nur ob oder ohne sehr
sein seine seinem seinen seiner seines selbst sich sie ihnen sind so solche solchem solchen
solcher solches soll sollte sondern sonst über um und uns unsere unserem unseren unser unseres
unter viel vom von vor während war waren warst was weg weil weiter welche welchem welchen welcher
welches wenn werde werden wie wieder will wir wird wirst wo wollen wollte würde würden zu zum zur
zwar zwischen
""".split()


@lru_cache(maxsize=4)
def stop_words(language: str = "german") -> frozenset:
    """NLTK stop-word list for ``language`` (only German is restated)."""
    try:  # exact reference behaviour when nltk + corpus are installed
        from nltk.corpus import stopwords  # type: ignore

        return frozenset(stopwords.words(language))
    except Exception:
        if language != "german":
            raise LookupError(f"stop words for {language!r} need nltk")
        return frozenset(_GERMAN_STOP_WORDS)


# --- restatement of the Treebank-style cascade -------------------------------------------

_OPENING_QUOTES = [
    (re.compile("([«“‘„]|[`]+)"), r" \1 "),
    (re.compile(r'^"'), r"``"),
    (re.compile(r"(``)"), r" \1 "),
    (re.compile(r"([ \(\[{<])(\"|'{2})"), r"\1 `` "),
    (re.compile(r"(?i)(')(?!re|ve|ll|m|t|s|d|n)(\w)\b"), r"\1 \2"),
]
_PUNCT = [
    (re.compile(r"([^\.])(\.)([\]\)}>\"'»”’ ]*)\s*$"), r"\1 \2 \3 "),
    (re.compile(r"([:,])([^\d])"), r" \1 \2"),
    (re.compile(r"([:,])$"), r" \1 "),
    (re.compile(r"\.{2,}"), r" \g<0> "),
    (re.compile(r"[;@#$%&]"), r" \g<0> "),
    (re.compile(r"([^\.])(\.)([\]\)}>\"']*)\s*$"), r"\1 \2\3 "),
    (re.compile(r"[?!]"), r" \g<0> "),
    (re.compile(r"([^'])' "), r"\1 ' "),
    (re.compile(r"[*]"), r" \g<0> "),
]
_BRACKETS = (re.compile(r"[\]\[\(\)\{\}\<\>]"), r" \g<0> ")
_DASHES = (re.compile(r"--"), r" -- ")
_CLOSING_QUOTES = [
    (re.compile("([»”’])"), r" \1 "),
    (re.compile(r"''"), " '' "),
    (re.compile(r'"'), " '' "),
    (re.compile(r"([^' ])('[sS]|'[mM]|'[dD]|') "), r"\1 \2 "),
    (re.compile(r"([^' ])('ll|'LL|'re|'RE|'ve|'VE|n't|N'T) "), r"\1 \2 "),
]
_CONTRACT2 = [
    re.compile(p)
    for p in (
        r"(?i)\b(can)(not)\b",
        r"(?i)\b(d)('ye)\b",
        r"(?i)\b(gim)(me)\b",
        r"(?i)\b(gon)(na)\b",
        r"(?i)\b(got)(ta)\b",
        r"(?i)\b(lem)(me)\b",
        r"(?i)\b(more)('n)\b",
        r"(?i)\b(wan)(na)(?=\s)",
    )
]
_CONTRACT3 = [re.compile(p) for p in (r"(?i) ('t)(is)\b", r"(?i) ('t)(was)\b")]

# crude stand-in for punkt: ". " followed by an upper-case letter ends a sentence unless the
# word before the period is a single letter or contains another period (abbreviation-like).
_SENT_END = re.compile(r"(?<=[.?!])\s+(?=[A-ZÄÖÜ0-9\"'(\[])")


def _split_sentences(text: str) -> List[str]:
    parts, start = [], 0
    for m in _SENT_END.finditer(text):
        head = text[start : m.start()]
        last = head.rsplit(None, 1)[-1] if head.strip() else ""
        if last.endswith(".") and (len(last) <= 2 or "." in last[:-1]):
            continue  # "z.B." / "B." style abbreviation: keep going
        parts.append(head)
        start = m.end()
    parts.append(text[start:])
    return [p for p in parts if p.strip()]


def _treebank(sentence: str) -> List[str]:
    text = sentence
    for rx, sub in _OPENING_QUOTES:
        text = rx.sub(sub, text)
    for rx, sub in _PUNCT:
        text = rx.sub(sub, text)
    text = _BRACKETS[0].sub(_BRACKETS[1], text)
    text = _DASHES[0].sub(_DASHES[1], text)
    text = " " + text + " "
    for rx, sub in _CLOSING_QUOTES:
        text = rx.sub(sub, text)
    for rx in _CONTRACT2:
        text = rx.sub(r" \1 \2 ", text)
    for rx in _CONTRACT3:
        text = rx.sub(r" \1 \2 ", text)
    return text.split()


_SIMPLE = re.compile(r"^[\w \t\n\r\-/+]*$")  # nothing any rule above would touch ...
# ... except the apostrophe-free contractions of _CONTRACT2, which the Treebank rules split in two
_CONTRACTION_WORD = re.compile(r"(?i)\b(?:cannot|gimme|gonna|gotta|lemme|wanna)\b")


def _is_simple(text: str) -> bool:
    return (_SIMPLE.match(text) is not None and "--" not in text
            and _CONTRACTION_WORD.search(text) is None)


_nltk_word_tokenize = None   # nltk's tokeniser once resolved; False when nltk cannot be used


def _resolve_nltk():
    """Looks for nltk ONCE (a failing import costs half a millisecond, per call it was three
    quarters of the host time of a comparison) and checks that its punkt data is usable."""
    global _nltk_word_tokenize
    if _nltk_word_tokenize is None:
        try:
            from nltk.tokenize import word_tokenize as fn  # type: ignore

            fn("Probe. Satz?")
            _nltk_word_tokenize = fn
        except Exception:  # noqa: BLE001 - not installed, or its data files are missing
            _nltk_word_tokenize = False
            logging.getLogger(__name__).warning(
                "nltk (with its punkt data) is not usable: tokenising with the built-in restatement "
                "of the Treebank word tokeniser and a crude sentence splitter; token sets of texts "
                "with abbreviations or unusual punctuation can differ from the reference's. "
                "Install nltk~=3.7 and its 'punkt' + 'stopwords' data for exact reference behaviour.")
    return _nltk_word_tokenize


def word_tokenize(text: str) -> List[str]:
    """``nltk.word_tokenize(text)`` (default language), or its restatement."""
    fn = _resolve_nltk()
    if fn:
        return fn(text)
    if _is_simple(text):
        return text.split()
    out: List[str] = []
    for sent in _split_sentences(text):
        out.extend(_treebank(sent))
    return out


def flatten_list(list_: Iterable) -> List[str]:
    """One level of flattening; str input is iterated per character
    (comparable_data.py:567-574, Q2)."""
    result: List[str] = []
    for part in list_:
        if isinstance(part, list):
            result.extend(part)
        else:
            result.append(part)
    return result


def tokenize(parts, language: str = "german") -> List[str]:
    """Token set of ``parts``: stop words and bare punctuation removed, de-duplicated,
    ordered by ``str.casefold`` (comparable_data.py:287-299)."""
    words = word_tokenize(" ".join(flatten_list(parts)))
    stops = stop_words(language)
    kept = {w for w in words if w.casefold() not in stops and w not in PREPARE_REMOVE_SYMBOLS}
    # ties under casefold ('Haus'/'haus') come out in set order in the reference, i.e. in no
    # defined order; break them by the raw string so packing is reproducible
    return sorted(kept, key=lambda w: (w.casefold(), w))


def gen_comp_value(items) -> List[List[str]]:
    """Level j = token set of the last j+1 parts (comparable_data.py:283-285).

    The reference tokenises every suffix from scratch (K (K + 1) / 2 part tokenisations per item).
    When the restated tokeniser is in use and no string of the item contains anything its rules
    would touch, the words of a suffix are just the words of its parts, so the levels are built
    incrementally: each part is split and filtered once, level j is level j-1 plus one part."""
    if _resolve_nltk() is False:
        fast = _gen_comp_value_simple(items)
        if fast is not None:
            return fast
    return [tokenize(items[-i:]) for i in range(1, len(items) + 1)]


_PART_WORDS: dict = {}          # text of a part -> its kept (casefold, word) pairs, or None: not simple
_PART_WORDS_CAP = 1 << 20       # distinct parts remembered (columns repeat their parts: ids, headers)
_MISS = object()


def _part_words(text):
    """The words of one simple part that survive the stop-word / punctuation filter, as
    (casefold, word) pairs in the order of the text; None when the part is not a simple str."""
    if not isinstance(text, str):
        return None
    if _is_simple(text):
        stops = stop_words("german")
        hit = tuple((w.casefold(), w) for w in text.split()
                    if w.casefold() not in stops and w not in PREPARE_REMOVE_SYMBOLS)
    else:
        hit = None
    if len(_PART_WORDS) < _PART_WORDS_CAP:
        _PART_WORDS[text] = hit
    return hit


def _gen_comp_value_simple(items):
    kept: set = set()            # (casefold, word) of the suffix so far ...
    ordered: list = []           # ... and the same pairs in sorted order (they sort without a key call)
    levels = []
    cached = _PART_WORDS.get
    for part in reversed(items):                       # the part that enters at the next level
        for text in (part if part.__class__ is list else (part,)):
            try:
                words = cached(text, _MISS)
            except TypeError:                          # an unhashable cell
                return None
            if words is _MISS:
                words = _part_words(text)
            if words is None:
                return None
            for fw in words:
                if fw not in kept:
                    kept.add(fw)
                    insort(ordered, fw)
        levels.append([w for _, w in ordered])
    return levels
