"""CPU: the engine's C++ text front end (tokenizer, sentence splitter, word count) is bit-exact against the upstream
SentencePiece wheel and against the Python restatement of the reference's text.h."""
import json
import os
import random

import pytest

from conftest import REPO

GOLD = os.path.join(REPO, "tests", "golden")
TOK = os.path.join(GOLD, "tokenizer.model")


@pytest.fixture(scope="module")
def front(P):
    return P.TextFrontEnd(TOK)


def test_token_ids_golden(front):
    g = json.load(open(os.path.join(GOLD, "text_golden.json")))
    for text, ids in g["token_ids"].items():
        assert front.encode(text) == ids, text


def test_token_ids_vs_sentencepiece(front):
    import sentencepiece as spm
    sp = spm.SentencePieceProcessor(model_file=TOK)
    rnd = random.Random(3)
    alphabet = list("abcdefghijklmnopqrstuvwxyzABCDEFGH    ...,!?'-0123456789") + ["é", "ß", "½", "™", "ﬁ", "日本", "😀", "\t", "\n", "　", "ｶ", "㍿"]
    cases = ["", " ", "   ", ".", "...", "....", ". . .", "a", "A.", "x" * 500]
    for _ in range(400):
        n = rnd.randint(1, 80)
        cases.append("".join(rnd.choice(alphabet) for _ in range(n)))
    for c in cases:
        assert front.encode(c) == list(sp.encode(c)), repr(c)


def test_count_words(P):
    import oracle
    L = P.lib()
    for t in ["", " ", "a", " a  b\tc\n", "The quick brown fox.", "x  ", "été deux"]:
        assert L.ptts_c_count_words(t.encode()) == oracle.count_words(t), repr(t)
        assert oracle.count_words(t) == len(t.split()) or any(ord(ch) > 127 for ch in t)


def test_sentence_splitter_matches_reference_restatement(front):
    import oracle
    rnd = random.Random(5)
    texts = ["hello there.  how are you? fine", "one. two! three?four...five", "   ", "no terminator", "a.b.c.", "trailing dot. ",
             "multi\nline\ttext. with   spaces !  ok"]
    for _ in range(100):
        n = rnd.randint(1, 120)
        texts.append("".join(rnd.choice("ab c.d!e?  \n") for _ in range(n)))
    for text in texts:
        for chunk in (1, 7, 15, 1000):                      # demos/pocket-tts.cpp feeds 15 chars at a time (:466-470)
            sp = oracle.StrProcessor()
            front.reset()
            for i in range(0, len(text), chunk):
                sp.ingest(text[i:i + chunk]); front.send(text[i:i + chunk])
            got_mid = front.pop_all()
            assert got_mid == list(sp.sentences), (text, chunk)
            sp.sentences.clear()
            sp.flush(); front.flush()
            assert front.pop_all() == list(sp.sentences), (text, chunk)
