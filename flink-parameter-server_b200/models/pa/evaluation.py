"""Accuracy of a trained passive-aggressive model on labelled examples -- the helpers the reference keeps
next to its tests (T/test/utils/PassiveAggressiveBinaryModelEvaluation.scala:14-44,
T/test/utils/PassiveAggressiveMultiModelEvaluation.scala:14-30).  Both return a PERCENTAGE and refuse
unlabelled examples, like the originals; ``confusion`` adds the binary confusion counts the original only logs."""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple


def _labelled(testLines: Iterable[Tuple[object, Optional[object]]]):
    for vector, label in testLines:
        if label is None:
            raise ValueError("Labels should not be missing.")
        yield vector, label


class PassiveAggressiveBinaryModelEvaluation:
    @staticmethod
    def confusion(model, testLines, pac) -> Dict[str, int]:
        c = {"tt": 0, "ff": 0, "tf": 0, "ft": 0}
        for vector, label in _labelled(testLines):
            c[("t" if bool(label) else "f") + ("t" if bool(pac.predict(vector, model)) else "f")] += 1
        return c

    @staticmethod
    def accuracy(model, testLines, featureCount: int = 0, pac=None) -> float:
        """``model``: ``{featureId: weight}`` (what the PS dumps at close); ``featureCount`` is unused, kept
        for signature parity."""
        c = PassiveAggressiveBinaryModelEvaluation.confusion(model, testLines, pac)
        n = sum(c.values())
        return 100.0 * (c["tt"] + c["ff"]) / n if n else float("nan")


class PassiveAggressiveMultiModelEvaluation:
    @staticmethod
    def accuracy(model, testLines, featureCount: int = 0, pac=None) -> float:
        hit = cnt = 0
        for vector, label in _labelled(testLines):
            hit += int(pac.predict(vector, model) == label)
            cnt += 1
        return 100.0 * hit / cnt if cnt else float("nan")
