"""Golden vectors for the CLIP tokenizer row, produced by the UNMODIFIED reference
(utils/simple_tokenizer.py + utils/dataset.py:tokenize) imported from /root/reference in this container.

`ftfy` is not installed here; the reference imports it unconditionally, so it is stubbed with the identity
(`fix_text(x) = x`), which is what ftfy does on text without mojibake - every case below is such text.
`utils/dataset.py` pulls cv2 / lmdb / pyarrow at import time, so `tokenize` is exec'd from its source lines
(read in place, not copied into this repository) against the reference tokenizer instance.

    python oracle/make_golden_tokenizer.py        # writes tests/golden/tokenizer_cases.json
"""
import inspect
import json
import os
import sys
import types

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tokenizer_cases.json")

CASES = [
    "the red mug on the left",
    "The banana behind the cereal box.",
    "pick up the second stapler from the right",
    "a green apple next to the keyboard's corner",
    "the bowl that is in front of the tissue box and to the left of the orange",
    "lemon",
    "  the   ball \t on  top\n of the towel  ",
    "it's the one they're holding; don't take the child's toy, I'll get it, we've seen it, I'm sure he'd agree",
    "object number 12 of 340, shelf 7b",
    "glue-stick (small) & marker -> tray #3!!!",
    "café crème brûlée naïve über straße",
    "红色的杯子 在 左边",
    "красная кружка слева",
    "emoji \U0001f34c and \U0001f9f4 bottle",
    "&lt;b&gt;bold&lt;/b&gt; &amp;amp; double-escaped &amp;quot;quotes&amp;quot;",
    "<|startoftext|> literal specials <|endoftext|> inside",
    "supercalifragilisticexpialidocious pneumonoultramicroscopicsilicovolcanoconiosis",
    "UPPER lower MiXeD CASE Words",
    "",
    "...",
    "the " * 40 + "end",
]


def main():
    sys.path.insert(0, REF)
    ftfy = types.ModuleType("ftfy")
    ftfy.fix_text = lambda x: x
    sys.modules.setdefault("ftfy", ftfy)
    import torch
    from utils.simple_tokenizer import SimpleTokenizer

    tk = SimpleTokenizer()
    src = open(os.path.join(REF, "utils", "dataset.py")).read().split("\n")
    start = next(i for i, ln in enumerate(src) if ln.startswith("def tokenize("))
    end = next(i for i in range(start + 1, len(src)) if src[i].startswith("def "))
    ns = {"torch": torch, "Union": __import__("typing").Union, "List": __import__("typing").List, "_tokenizer": tk}
    exec("\n".join(src[start:end]), ns)
    tokenize = ns["tokenize"]

    out = {"cases": [], "specials": {"sot": tk.encoder["<|startoftext|>"], "eot": tk.encoder["<|endoftext|>"]},
           "vocab_size": len(tk.encoder),
           "vocab_probe": {s: tk.encoder[s] for s in ["!", "a", "z</w>", "Ġ", "Ā", "the</w>", "in", "th"] if s in tk.encoder}}
    for text in CASES:
        ids = tk.encode(text)
        row = {"text": text, "ids": ids, "decoded": tk.decode(ids)}
        for L, trunc in ((77, False), (20, True), (17, True)):
            try:
                row[f"tokenize_{L}"] = tokenize(text, L, trunc)[0].tolist()
            except RuntimeError as e:
                row[f"tokenize_{L}"] = "RuntimeError"
        out["cases"].append(row)
    try:
        tokenize("the " * 100, 77, False)
        out["too_long_raises"] = False
    except RuntimeError:
        out["too_long_raises"] = True
    out["batch_20"] = tokenize(CASES[:4], 20, True).tolist()
    with open(OUT, "w") as f:
        json.dump(out, f, ensure_ascii=True, indent=0)
    print("wrote", OUT, len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
