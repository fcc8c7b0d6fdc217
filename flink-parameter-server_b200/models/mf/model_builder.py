"""``OnlineFactorModelBuilder`` (M/matrix/factorization/OnlineFactorModelBuilder.scala:5-12): an
abstraction the reference declares but never uses; kept for API parity."""
from __future__ import annotations

from typing import Any


class OnlineFactorModelBuilder:
    def buildModel(self, ratings: Any, factorInit: Any, factorUpdate: Any, parameters: dict) -> Any:
        raise NotImplementedError
