"""Golden vectors for mmvid_b200.tokenizer: ids produced by the UNMODIFIED reference tokenizer
(/root/reference/mmvid_pytorch/tokenizer.py) on a fixed caption list.  `ftfy` is absent from this image; it is shimmed as
the identity, which it is on the clean ASCII / UTF-8 captions used here.  Run in the build container only:

    python tests/golden/gen_tokenizer_golden.py      # writes tests/golden/tokenizer.json
"""
import json
import os
import sys
import types

REF = "/root/reference"
CAPTIONS = [
    "A person is doing push ups on the floor.",
    "an object moving to the left, then the right; a large red sphere and a small blue cube",
    "She has wavy hair, arched eyebrows and is wearing lipstick & earrings.",
    "two dogs running in the park... it's sunny, they're happy!",
    "Résumé of a naïve café owner — 3 cats, 12 dogs, 100% déjà vu",
    "  multiple   spaces\tand\nnewlines   ",
    "<|startoftext|>a video of fireworks<|endoftext|>",
    "supercalifragilisticexpialidocious antidisestablishmentarianism 1234567890",
    "&amp;lt;tag&amp;gt; HTML &quot;entities&quot; don't survive",
    "",
    "x",
    "日本語のテキスト and emoji 🎆",
]

if __name__ == "__main__":
    shim = types.ModuleType("ftfy")
    shim.fix_text = lambda s: s
    sys.modules["ftfy"] = shim
    sys.path.insert(0, REF)
    from mmvid_pytorch.tokenizer import SimpleTokenizer
    tok = SimpleTokenizer()
    ids = [tok.encode(c) for c in CAPTIONS]
    out = {"captions": CAPTIONS, "ids": ids, "decoded": [tok.decode(i) for i in ids],
           "tokenize_64": tok.tokenize(CAPTIONS[:4], context_length=64).tolist(),
           "vocab_size": tok.vocab_size, "n_vocab_entries": len(tok.encoder)}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tokenizer.json")
    with open(path, "w") as f:
        json.dump(out, f, ensure_ascii=True, indent=0)
    print("wrote", path, sum(len(i) for i in ids), "ids")
