"""Host input pipeline, text side: the byte-level BPE tokenizer the reference feeds `BERT` / `DALLE` with
(`mmvid_pytorch/tokenizer.py:61-171`, itself OpenAI CLIP's `simple_tokenizer`), restated from the published algorithm, plus
pinned staging of token batches.  Same constructor / `encode` / `decode` / `tokenize` surface and the same ids.

The merge table is DATA the user supplies (`bpe_simple_vocab_16e6.txt`, shipped with the reference under
`mmvid_pytorch/data/` and with OpenAI CLIP); it is not copied into this repository.  Path resolution: the `bpe_path`
argument, else `$MMVID_BPE_PATH`.

Differences kept deliberately small:
* `ftfy.fix_text` (mojibake repair) is applied when `ftfy` is importable, exactly like the reference; without it the text is
  used as is (identical for clean UTF-8 / ASCII captions, which is what the golden vectors in `tests/golden/tokenizer.json`
  cover);
* `tokenize(..., pin_memory=True)` returns the `[B, context_length]` int64 batch in pinned host memory so that
  `.to(device, non_blocking=True)` overlaps with compute (the reference's `.cuda()` of a pageable tensor, `test.py:227`,
  synchronises).
"""
import html
import os
from functools import lru_cache

import regex as re
import torch

try:  # optional, like every other third-party dependency of the reference that is absent from this image
    import ftfy
except ImportError:  # pragma: no cover
    ftfy = None

VOCAB_SIZE = 49408
N_MERGES = 49152 - 256 - 2
SOT, EOT = "<|startoftext|>", "<|endoftext|>"
END = "</w>"


@lru_cache()
def byte_alphabet():
    """Printable stand-in character for each of the 256 byte values (GPT-2 / CLIP byte-level BPE): the printable
    Latin-1 ranges map to themselves, the remaining bytes to code points 256, 257, ... in byte order."""
    keep = list(range(0x21, 0x7F)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    table, extra = {}, 0
    for b in keep:
        table[b] = chr(b)
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + extra)
            extra += 1
    return table


def default_bpe():
    p = os.environ.get("MMVID_BPE_PATH")
    if not p:
        raise FileNotFoundError("SimpleTokenizer needs the BPE merge table: pass bpe_path= or set MMVID_BPE_PATH "
                                "(bpe_simple_vocab_16e6.txt from the reference's mmvid_pytorch/data/ or OpenAI CLIP)")
    return p


def _clean(text):
    if ftfy is not None:
        text = ftfy.fix_text(text)
    text = html.unescape(html.unescape(text)).strip()
    return re.sub(r"\s+", " ", text).strip()


class SimpleTokenizer(object):
    def __init__(self, bpe_path=None):
        bpe_path = bpe_path or default_bpe()
        alphabet = byte_alphabet()
        self.byte_encoder = alphabet
        self.byte_decoder = {c: b for b, c in alphabet.items()}
        with open(bpe_path, encoding="utf8") as f:
            lines = f.read().split("\n")
        pairs = [tuple(ln.split()) for ln in lines[1:1 + N_MERGES]]  # line 0 is the "#version" header
        pairs = [p for p in pairs if len(p) == 2]  # (a short hand-made table may end in a blank line)
        # vocabulary order (tokenizer.py:66-72): 256 byte symbols in the alphabet's insertion order, the same with the
        # end-of-word marker, every merge in rank order, the two specials
        symbols = list(alphabet.values())
        vocab = symbols + [s + END for s in symbols] + [a + b for a, b in pairs] + [SOT, EOT]
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.decoder = {i: tok for tok, i in self.encoder.items()}
        self.rank = {p: r for r, p in enumerate(pairs)}
        self.vocab_size = VOCAB_SIZE
        self._word_ids = {SOT: (self.encoder[SOT],), EOT: (self.encoder[EOT],)}
        self.pat = re.compile(r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+",
                              re.IGNORECASE)

    # ------------------------------------------------------------------ BPE
    def _merge_word(self, word):
        """ids of one pre-token: greedy lowest-rank merges over the symbol list, in place (tokenizer.py:87-128 computes the
        same fixed point by rebuilding the word once per merge)."""
        syms = list(word[:-1]) + [word[-1] + END]
        rank = self.rank
        while len(syms) > 1:
            best, best_rank = None, None
            for i in range(len(syms) - 1):
                r = rank.get((syms[i], syms[i + 1]))
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = (syms[i], syms[i + 1]), r
            if best is None:
                break
            a, b = best
            out, i, n = [], 0, len(syms)
            while i < n:  # every non-overlapping occurrence, left to right
                if i + 1 < n and syms[i] == a and syms[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(syms[i])
                    i += 1
            syms = out
        return tuple(self.encoder[s] for s in syms)

    def encode(self, text):
        ids = []
        enc = self.byte_encoder
        for tok in self.pat.findall(_clean(text).lower()):
            got = self._word_ids.get(tok)
            if got is None:
                got = self._merge_word("".join(enc[b] for b in tok.encode("utf-8")))
                self._word_ids[tok] = got
            ids.extend(got)
        return ids

    def decode(self, tokens, remove_start_end=True):
        if torch.is_tensor(tokens):
            tokens = tokens.tolist()
        if remove_start_end:
            tokens = [t for t in tokens if t not in (49406, 40407, 0)]  # (sic) the reference's constants, tokenizer.py:145
        text = "".join(self.decoder[t] for t in tokens)
        return bytearray(self.byte_decoder[c] for c in text).decode("utf-8", errors="replace").replace(END, " ")

    def tokenize(self, texts, context_length=256, truncate_text=False, pin_memory=False):
        if isinstance(texts, str):
            texts = [texts]
        result = torch.zeros(len(texts), context_length, dtype=torch.long, pin_memory=pin_memory)
        for i, text in enumerate(texts):
            ids = self.encode(text)
            if len(ids) > context_length:
                if not truncate_text:
                    raise RuntimeError(f"Input {text} is too long for context length {context_length}")
                ids = ids[:context_length]
            result[i, :len(ids)] = torch.tensor(ids, dtype=torch.long)
        return result


class PinnedStager:
    """Double-buffered pinned host staging for the per-batch inputs of `train.py:265-272` / `test.py:225-227` (`text`, `frames`,
    `visuals` -> `.cuda()`): `put` copies a batch into the next pinned slot and enqueues the H2D copies on a side stream;
    `get` makes the compute stream wait for them.  With two slots the copy of batch i+1 overlaps the step on batch i."""

    def __init__(self, device, slots=2):
        self.device = torch.device(device)
        self.slots = [dict() for _ in range(slots)]
        self.events = [None] * slots
        self.stream = torch.cuda.Stream(device=self.device)
        self.head = 0
        self.pending = []

    def put(self, **tensors):
        slot = self.head % len(self.slots)
        self.head += 1
        host = self.slots[slot]
        if self.events[slot] is not None:
            self.events[slot].synchronize()  # the slot's previous copies have left host memory
        out = {}
        with torch.cuda.stream(self.stream):
            for name, t in tensors.items():
                if t is None:
                    out[name] = None
                    continue
                buf = host.get(name)
                if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                    buf = host[name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                buf.copy_(t)
                out[name] = buf.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.events[slot] = ev
        self.pending.append((ev, out))

    def get(self):
        ev, out = self.pending.pop(0)
        torch.cuda.current_stream(self.device).wait_event(ev)
        for t in out.values():
            if t is not None:
                t.record_stream(torch.cuda.current_stream(self.device))
        return out
