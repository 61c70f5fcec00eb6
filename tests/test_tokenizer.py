"""CLIP tokenizer row: crog_b200.utils.simple_tokenizer against golden vectors produced by the
unmodified reference tokenizer (oracle/make_golden_tokenizer.py).  The merge table is data that ships with the
reference, not with this repository, so the id-level checks run wherever it can be found (CROG_BPE_PATH, a copy next
to the module, or the reference checkout) and are skipped elsewhere; the table-independent parts always run."""
import json
import os

import pytest
import torch

from crog_b200.utils import simple_tokenizer as ST

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tokenizer_cases.json")))


def _have_table():
    try:
        return os.path.exists(ST.default_bpe())
    except FileNotFoundError:
        return False


needs_table = pytest.mark.skipif(not _have_table(), reason="bpe_simple_vocab_16e6.txt.gz not available")


def test_byte_alphabet_is_the_gpt2_table():
    a = ST.byte_alphabet()
    assert len(a) == 256 and len(set(a)) == 256
    assert a[ord("a")] == "a" and a[ord("!")] == "!" and a[0xA1] == "¡" and a[0xFF] == "ÿ"
    assert a[0] == "Ā" and a[ord(" ")] == "Ġ" and a[0x7F] == "ġ" and a[0xAD] == "Ń"
    order = ST.vocabulary_order(a)
    assert order[0] == "!" and order[93] == "~" and order[94] == "¡" and order[188] == "Ā" and len(order) == 256


def test_cleaning_without_table():
    assert ST.whitespace_clean("  a \t b\n\nc ") == "a b c"
    assert ST.basic_clean(" &amp;lt;x&amp;gt; ") == "<x>"


def test_missing_table_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setenv("CROG_BPE_PATH", str(tmp_path / "nope.txt.gz"))
    with pytest.raises(FileNotFoundError):
        ST.SimpleTokenizer()


@needs_table
def test_vocabulary_matches_reference():
    tk = ST.SimpleTokenizer()
    assert len(tk.encoder) == GOLD["vocab_size"]
    assert tk.sot_token == GOLD["specials"]["sot"] == 49406 and tk.eot_token == GOLD["specials"]["eot"] == 49407
    for s, i in GOLD["vocab_probe"].items():
        assert tk.encoder[s] == i, s


@needs_table
def test_encode_decode_match_reference():
    tk = ST.SimpleTokenizer()
    for case in GOLD["cases"]:
        ids = tk.encode(case["text"])
        assert ids == case["ids"], case["text"]
        assert tk.decode(ids) == case["decoded"], case["text"]


@needs_table
def test_tokenize_matches_reference():
    tk = ST.SimpleTokenizer()
    for case in GOLD["cases"]:
        for L, trunc in ((77, False), (20, True), (17, True)):
            want = case[f"tokenize_{L}"]
            if want == "RuntimeError":
                with pytest.raises(RuntimeError):
                    ST.tokenize(case["text"], L, trunc, tokenizer=tk)
                continue
            got = ST.tokenize(case["text"], L, trunc, tokenizer=tk)
            assert got.dtype == torch.int64 and got.shape == (1, L)
            assert got[0].tolist() == want, (case["text"], L)
    assert GOLD["too_long_raises"]
    with pytest.raises(RuntimeError):
        ST.tokenize("the " * 100, 77, False, tokenizer=tk)
    batch = ST.tokenize([c["text"] for c in GOLD["cases"][:4]], 20, True, tokenizer=tk)
    assert batch.tolist() == GOLD["batch_20"]
    # the EOT id is the arg-max of every row, which is what the text encoder's gather relies on (clip.py:450-451)
    assert (batch.argmax(1) == (batch == tk.eot_token).int().argmax(1)).all()
