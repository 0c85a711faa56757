"""Text front-end oracle (TEST INFRASTRUCTURE): pure-Python restatement of the reference's
sentence splitter and word counter (src/pocket_tts/conditioners/text.h:52-94,181-251), and the
tokenizer oracle = the upstream SentencePiece core through its Python wheel (the reference links the
same C++ core: src/pocket_tts/conditioners/text.h:10-27)."""
from __future__ import annotations

from collections import deque

_C_SPACE = b" \t\n\v\f\r"


def _isspace(ch: int) -> bool:          # C isspace() in the "C" locale, byte-wise like the reference
    return bytes([ch]) in [bytes([c]) for c in _C_SPACE]


def _islower(ch: int) -> bool:
    return 0x61 <= ch <= 0x7A


def _isalnum(ch: int) -> bool:
    return (0x30 <= ch <= 0x39) or (0x41 <= ch <= 0x5A) or (0x61 <= ch <= 0x7A)


def count_words(text) -> int:
    """text.h:81-94 — number of maximal non-whitespace runs."""
    b = text.encode() if isinstance(text, str) else bytes(text)
    n, i, words = len(b), 0, 0
    while i < n and _isspace(b[i]):
        i += 1
    while i < n:
        words += 1
        while i < n and not _isspace(b[i]):
            i += 1
        if i == n:
            return words
        while i < n and _isspace(b[i]):
            i += 1
    return words


def frames_after_eos_guess(text) -> int:
    """src/pocket_tts.cpp:504-506."""
    return (3 if count_words(text) <= 4 else 1) + 2


def max_gen_len_for(text) -> int:
    """src/pocket_tts.cpp:429-430: int((count_words + 2.0f) * 12.5f)."""
    return int((count_words(text) + 2.0) * 12.5)


class StrProcessor:
    """text.h:181-251 — streaming sentence splitter (byte-wise, like the C++ char loop)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.tail = bytearray()
        self.sentences = deque()
        self.was_whitespace = True
        self.was_eos = False
        self.leading_char = True

    def ingest(self, chunk):
        b = chunk.encode() if isinstance(chunk, str) else bytes(chunk)
        for c in b:
            is_eos = c in b".!?"
            if not is_eos and self.was_eos:
                self.sentences.append(bytes(self.tail))
                self.tail = bytearray()
                self.was_whitespace = True
                self.leading_char = True
            is_ws = _isspace(c)
            if is_ws and not self.was_whitespace:
                self.tail.append(0x20)
            elif not is_ws:
                if self.leading_char:
                    if _islower(c):
                        c = c - 32
                    self.leading_char = False
                self.tail.append(c)
            self.was_whitespace = is_ws
            self.was_eos = is_eos

    def flush(self):
        if len(self.tail):
            if _isalnum(self.tail[-1]):
                self.tail.append(ord("."))
            self.sentences.append(bytes(self.tail))
            self.tail = bytearray()
        self.was_whitespace = True
        self.was_eos = False
        self.leading_char = True


class SentencePieceOracle:
    def __init__(self, model_path: str):
        import sentencepiece as spm
        self.sp = spm.SentencePieceProcessor(model_file=model_path)

    def encode(self, text) -> list[int]:
        if isinstance(text, (bytes, bytearray)):
            text = bytes(text).decode("utf-8", errors="replace")
        return list(self.sp.encode(text))
