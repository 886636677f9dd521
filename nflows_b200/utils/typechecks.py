"""Argument checks used by constructors (same predicates as reference nflows/utils/typechecks.py)."""


def is_bool(x):
    return isinstance(x, bool)


def is_int(x):
    return isinstance(x, int)


def is_positive_int(x):
    return is_int(x) and x > 0


def is_nonnegative_int(x):
    return is_int(x) and x >= 0


def is_power_of_two(n):
    return is_positive_int(n) and not n & (n - 1)
