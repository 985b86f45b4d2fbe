"""B200-native common-nearest-neighbour clustering: the ``fit`` hot path of ``cnnclustering``
(radius-neighbour search -> common-neighbour similarity -> density-connected components ->
canonical relabel) as hand-written sm_100a CUDA behind a C-ABI (``include/cnnb200.h``).

The package mirrors the reference's component roles (``_types``: InputData, NeighboursGetter,
Metric, SimilarityChecker; ``_fit``: Fitter; ``cluster``: Clustering) with bulk, device-side
implementations.  There is no CPU fallback.
"""
__version__ = "0.1.0"
