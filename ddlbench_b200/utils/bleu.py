"""Corpus BLEU (Papineni et al. 2002; brevity penalty, uniform 4-gram weights, optional +1
smoothing) — the metric of the reference's ``runtime/translation/compute_bleu_scores.py``, which
shells out to sacrebleu (not available here)."""
from __future__ import annotations

import math
from collections import Counter
from typing import List, Sequence, Tuple


def _ngrams(tokens: Sequence[str], n: int) -> Counter:
    return Counter(tuple(tokens[i:i + n]) for i in range(len(tokens) - n + 1))


def corpus_bleu(hypotheses: Sequence[str], references: Sequence[str], max_n: int = 4,
                smooth: bool = False, lowercase: bool = False) -> Tuple[float, List[float], float]:
    """-> (BLEU in [0, 100], n-gram precisions, brevity penalty)."""
    match, total = [0] * max_n, [0] * max_n
    hyp_len = ref_len = 0
    for h, r in zip(hypotheses, references):
        if lowercase:
            h, r = h.lower(), r.lower()
        ht, rt = h.split(), r.split()
        hyp_len += len(ht)
        ref_len += len(rt)
        for n in range(1, max_n + 1):
            hn, rn = _ngrams(ht, n), _ngrams(rt, n)
            match[n - 1] += sum(min(c, rn[g]) for g, c in hn.items())
            total[n - 1] += max(len(ht) - n + 1, 0)
    precisions = []
    for m, t in zip(match, total):
        if smooth:
            precisions.append((m + 1.0) / (t + 1.0))
        else:
            precisions.append(m / t if t > 0 else 0.0)
    if min(precisions) <= 0.0:
        return 0.0, precisions, 0.0
    bp = 1.0 if hyp_len > ref_len else math.exp(1.0 - ref_len / max(hyp_len, 1))
    score = bp * math.exp(sum(math.log(p) for p in precisions) / max_n)
    return 100.0 * score, precisions, bp


def cli(argv=None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description="corpus BLEU of a hypothesis file against a reference file")
    ap.add_argument("hypothesis")
    ap.add_argument("reference")
    ap.add_argument("--lc", action="store_true")
    a = ap.parse_args(argv)
    hyp = [l.rstrip("\n") for l in open(a.hypothesis, encoding="utf-8")]
    ref = [l.rstrip("\n") for l in open(a.reference, encoding="utf-8")]
    score, prec, bp = corpus_bleu(hyp, ref, lowercase=a.lc)
    print("BLEU = %.2f %s (BP = %.3f)" % (score, "/".join("%.1f" % (100 * p) for p in prec), bp))
    return 0


if __name__ == "__main__":
    raise SystemExit(cli())
