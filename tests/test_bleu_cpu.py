"""sentence_bleu(method4) restatement (campaign.sentence_bleu_method4) against closed forms evaluated by hand from nltk 3.8's
algorithm (nltk.translate.bleu_score: corpus_bleu, modified_precision, brevity_penalty, SmoothingFunction.method4 with k = 5):

    p_n   = clipped n-gram matches / max(1, #hyp n-grams)                   n = 1..4
    p_1.numerator == 0                 -> 0
    p_n.numerator == 0 and len(hyp) > 1 -> p_n = (1 / (2^i * 5 / ln(len(hyp)))) / denominator,  i = 1, 2, ... per smoothed order
    bleu  = bp * exp(sum_{p_n > 0} 0.25 * ln p_n),   bp = 1 if len(hyp) > len(ref) else exp(1 - len(ref) / len(hyp))
"""
import math

import pytest

from onnx_transformer_b200 import campaign as C


def test_identical_long_sentence_is_one():
    g = [5, 6, 7, 8, 9, 10]
    assert C.sentence_bleu_method4(g, g) == pytest.approx(1.0, abs=1e-12)


def test_one_token_hypothesis_is_not_zero():
    # len(hyp) == 1: method4 does not smooth, orders 2-4 have p = 0 and are skipped by the p_i > 0 filter
    assert C.sentence_bleu_method4([5], [5]) == pytest.approx(1.0)
    assert C.sentence_bleu_method4([5, 6], [5]) == pytest.approx(math.exp(1 - 2 / 1))          # brevity penalty only
    assert C.sentence_bleu_method4([5, 6, 7], [6]) == pytest.approx(math.exp(1 - 3 / 1))
    assert C.sentence_bleu_method4([5], [9]) == 0.0                                           # no unigram match


def test_short_sentences_with_themselves_are_smoothed_not_one():
    # len 2: p1 = 2/2, p2 = 1/1, p3 and p4 have no n-grams (denominator max(1, 0) = 1) and are smoothed with i = 1, 2
    ln2 = math.log(2)
    want2 = math.exp(0.25 * (math.log(ln2 / 10) + math.log(ln2 / 20)))
    assert C.sentence_bleu_method4([5, 6], [5, 6]) == pytest.approx(want2, rel=1e-12)
    # len 3: only p4 is smoothed (i = 1)
    want3 = math.exp(0.25 * math.log(math.log(3) / 10))
    assert C.sentence_bleu_method4([5, 6, 7], [5, 6, 7]) == pytest.approx(want3, rel=1e-12)
    # len 4: all four orders match
    assert C.sentence_bleu_method4([5, 6, 7, 8], [5, 6, 7, 8]) == pytest.approx(1.0)


def test_zero_higher_order_matches():
    ref, hyp = [5, 6, 7, 8, 9, 10], [5, 7, 9, 6, 8, 10]      # all unigrams match, no bigram does
    l6 = math.log(6)
    # p1 = 6/6 ; p2 = smoothed(i=1)/5 ; p3 = smoothed(i=2)/4 ; p4 = smoothed(i=3)/3
    want = math.exp(0.25 * (math.log(l6 / 10 / 5) + math.log(l6 / 20 / 4) + math.log(l6 / 40 / 3)))
    assert C.sentence_bleu_method4(ref, hyp) == pytest.approx(want, rel=1e-12)


def test_one_substitution_and_brevity_penalty():
    ref = [5, 6, 7, 8, 9, 10]
    hyp = [5, 6, 7, 30, 9, 10]
    # unigrams 5/6, bigrams (5,6),(6,7),(9,10) = 3/5, trigrams (5,6,7) = 1/4, 4-grams 0/3 -> smoothed i = 1
    want = math.exp(0.25 * (math.log(5 / 6) + math.log(3 / 5) + math.log(1 / 4) + math.log(math.log(6) / 10 / 3)))
    assert C.sentence_bleu_method4(ref, hyp) == pytest.approx(want, rel=1e-12)
    hyp_short = [5, 6, 7, 8]                                   # prefix of the reference: all orders match, bp = exp(1 - 6/4)
    assert C.sentence_bleu_method4(ref, hyp_short) == pytest.approx(math.exp(1 - 6 / 4), rel=1e-12)
    hyp_long = ref + [11]                                      # longer than the reference: bp = 1 ; p = 6/7, 5/6, 4/5, 3/4
    want = math.exp(0.25 * (math.log(6 / 7) + math.log(5 / 6) + math.log(4 / 5) + math.log(3 / 4)))
    assert C.sentence_bleu_method4(ref, hyp_long) == pytest.approx(want, rel=1e-12)


def test_clipping_and_empty():
    # modified precision clips repeated n-grams by the reference count: hyp "5 5 5 5" vs ref "5 6 7 8" -> p1 = 1/4
    l4 = math.log(4)
    want = math.exp(0.25 * (math.log(1 / 4) + math.log(l4 / 10 / 3) + math.log(l4 / 20 / 2) + math.log(l4 / 40 / 1)))
    assert C.sentence_bleu_method4([5, 6, 7, 8], [5, 5, 5, 5]) == pytest.approx(want, rel=1e-12)
    assert C.sentence_bleu_method4([5, 6], []) == 0.0
    assert C.sentence_bleu_method4([5, 6, 7, 8], [40, 41]) == 0.0


def test_classification_rows():
    import numpy as np
    ys = np.array([0, 5, 6, 1, 9, 9]); ys2 = np.array([0, 5, 7, 1, 9, 9]); ys3 = np.array([0, 5, 6, 7, 8, 9])
    assert C.classify(ys, ys)["outcome"] == "masked" and C.classify(ys, ys2)["outcome"] == "changed"
    assert C.classify(ys, ys3)["outcome"] == "no-EOS"
    one = np.array([0, 5, 1, 7, 7, 7]); other = np.array([0, 6, 1, 7, 7, 7])
    r = C.classify(one, other)          # 1-token golden sentence: golden BLEU is 1.0 (not 0), so a changed output is "changed"
    assert r["golden_bleu"] == pytest.approx(1.0) and r["faulty_bleu"] == 0.0 and r["outcome"] == "changed"
