"""CPU suite: the host tokeniser (text/tokenize.py) — restatement of nltk's word_tokenize used when
nltk is not installed, stop-word / punctuation filter, and the incremental gen_comp_value."""
import random

import pytest

from napkon_string_matching.text import tokenize as T

# outputs of nltk 3.7's word_tokenize for inputs without sentence-boundary ambiguity
KNOWN = [
    ("Hatte Sie Dialyse oder sonstiges?", ["Hatte", "Sie", "Dialyse", "oder", "sonstiges", "?"]),
    ("Gewicht (kg): 70,5", ["Gewicht", "(", "kg", ")", ":", "70,5"]),
    ("Ja/Nein", ["Ja/Nein"]),
    ('Er sagte "ja".', ["Er", "sagte", "``", "ja", "''", "."]),
    ("Blutdruck: systolisch", ["Blutdruck", ":", "systolisch"]),
    ("can't cannot", ["ca", "n't", "can", "not"]),
    ("Alter [Jahre]", ["Alter", "[", "Jahre", "]"]),
    ("'tis gut", ["'t", "is", "gut"]),
    ("50% der Fälle", ["50", "%", "der", "Fälle"]),
    ("COVID-19 -- schwer", ["COVID-19", "--", "schwer"]),
    ("", []),
    # apostrophe-free contractions (Treebank CONTRACTIONS2): two tokens, also without punctuation
    ("cannot", ["can", "not"]),
    ("Ich wanna gehen", ["Ich", "wan", "na", "gehen"]),
    ("gonna gotta lemme gimme", ["gon", "na", "got", "ta", "lem", "me", "gim", "me"]),
    ("Cannot CANNOT cannots", ["Can", "not", "CAN", "NOT", "cannots"]),
]


@pytest.mark.skipif(T._resolve_nltk() is not False, reason="nltk itself is in use")
@pytest.mark.parametrize("text,want", KNOWN)
def test_restated_word_tokenize(text, want):
    assert T.word_tokenize(text) == want


def test_tokenize_filters_and_orders_like_the_reference():
    # comparable_data.py:287-299: German stop words by casefold, bare punctuation, set, casefold order
    assert T.tokenize(["Hatte Sie Dialyse oder sonstiges?"]) == ["Dialyse", "sonstiges"]   # "hatte", "sie", "oder": stop words
    assert T.tokenize(["Der die DAS", "Haus haus"]) == ["Haus", "haus"]
    assert T.tokenize([["Nieren", "Dialyse"], "nicht"]) == ["Dialyse", "Nieren"]
    assert T.tokenize("abc") == ["a", "b", "c"]       # a str is iterated per character (Q2)


def test_gen_comp_value_levels_are_suffix_token_sets():
    assert T.gen_comp_value(["Kopf", "Frage eins", "Antwort"]) == \
        [["Antwort"], ["Antwort", "eins", "Frage"], ["Antwort", "eins", "Frage", "Kopf"]]
    assert T.gen_comp_value("gec_abc")[:3] == [["c"], ["b", "c"], ["a", "b", "c"]]
    assert T.gen_comp_value([]) == []


def test_incremental_gen_comp_value_equals_tokenising_every_suffix():
    rnd = random.Random(1)
    words = ["Haus", "haus", "der", "Die", "und", "Dialyse", "x-y", "a/b", "B12", "+", "nicht", "Nieren", "cannot", "wanna",
             "über", "ÄRZTE", "1.5", "z.B.", "(ja)", "nein?", "--", "a--b", "", "ist"]
    text = lambda: " ".join(rnd.choice(words) for _ in range(rnd.randint(0, 6)))  # noqa: E731
    taken = 0
    for _ in range(4000):
        if rnd.random() < 0.15:
            item = "".join(rnd.choice("gec_ab1-/ ") for _ in range(rnd.randint(0, 9)))
        else:
            item = [[text() for _ in range(rnd.randint(0, 3))] if rnd.random() < 0.3 else text()
                    for _ in range(rnd.randint(0, 5))]
        want = [T.tokenize(item[-i:]) for i in range(1, len(item) + 1)]
        assert T.gen_comp_value(item) == want, item
        taken += T._gen_comp_value_simple(item) is not None
    assert 0 < taken < 4000   # both the incremental and the generic path were exercised
