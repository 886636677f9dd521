"""Import shim: the reference imports the third-party `UMNN` package unconditionally
(nflows/transforms/UMNN/MonotonicNormalizer.py:2). It is not installed and off the hot path."""


class NeuralIntegral:  # pragma: no cover - placeholder
    pass


class ParallelNeuralIntegral:  # pragma: no cover - placeholder
    pass
