"""Fail-fast exceptions shared by the host and device tiers (the reference has no fault
tolerance, only these fail-fast checks)."""


class FactorIsNotANumberException(FloatingPointError):
    """A NaN / Inf appeared in a factor vector or was pushed to the parameter server
    (M/matrix/factorization/utils/Vector.scala:78-80).  Subclass of ``FloatingPointError`` (hence of
    ``ArithmeticError``), so callers may catch either."""
