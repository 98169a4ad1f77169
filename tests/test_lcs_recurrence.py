"""CPU suite: the bit-parallel LCS recurrence of csrc/qratio.cu restated word by word in Python —
u = S & M[c]; S = (S + u) | (S & ~M[c]) on W 64-bit words with the carry of the add chained from
word to word — against the textbook dynamic programme of the oracle port."""
import random

from oracle import reference_port as port

MASK64 = (1 << 64) - 1


def lcs_words(pattern: str, text: str) -> int:
    words = max(1, -(-len(pattern) // 64))
    masks = {}
    for i, ch in enumerate(pattern):
        m = masks.setdefault(ch, [0] * words)
        m[i >> 6] |= 1 << (i & 63)
    zero = [0] * words
    S = [MASK64] * words
    for ch in text:
        M = masks.get(ch, zero)
        carry = 0
        for x in range(words):
            u = S[x] & M[x]
            total = S[x] + u + carry          # add.cc / addc.cc over the 32-bit halves
            carry = total >> 64
            S[x] = (total & MASK64) | (S[x] & ~M[x] & MASK64)
    return sum(64 - bin(s).count("1") for s in S)


def test_word_chained_recurrence_equals_dynamic_programming():
    rnd = random.Random(5)
    for alphabet, max_len in (("ab", 40), ("abcdefgh ", 200), ("abcdefghijklmnopqrstuvwxyzäöü0123456789 ", 520)):
        for _ in range(150):
            a = "".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, max_len)))
            b = "".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, max_len)))
            assert lcs_words(a, b) == port.lcs_length(a, b), (a, b)
    # lengths on the word boundaries the kernel classes use
    for n in (63, 64, 65, 127, 128, 129, 191, 192, 193, 255, 256, 257, 383, 384, 511, 512):
        a = "".join(rnd.choice("abc") for _ in range(n))
        b = a[::-1]
        assert lcs_words(a, b) == port.lcs_length(a, b)
        assert lcs_words(a, a) == n


def test_score_map_known_answer():
    # rapidfuzz docs: ratio("this is a test", "this is a test!") = 96.55172413793103
    a, b = "this is a test", "this is a test!"
    lcs = lcs_words(a, b)
    dist, lensum = len(a) + len(b) - 2 * lcs, len(a) + len(b)
    assert (dist, lensum) == (1, 29)
    assert port.indel_ratio_from_counts(dist, lensum) * 100 == 96.55172413793103
